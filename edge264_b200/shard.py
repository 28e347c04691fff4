"""Stream sharding for multi-GPU runs: rank 0 owns the Annex-B inputs, one NCCL (or gloo, in CPU tests)
broadcast ships the concatenated buffer + size table, every rank keeps its own contiguous shard.
This is the only collective of the whole path: streams are independent."""
import torch


def shard_range(n_streams_per_rank, rank):
    return range(rank * n_streams_per_rank, (rank + 1) * n_streams_per_rank)


def broadcast_streams(bufs_all, per_rank, world, rank, dist=None, device="cpu"):
    """bufs_all: list of bytes on rank 0 (len = per_rank*world), ignored elsewhere.  Returns this rank's list."""
    n = per_rank * world
    if world == 1 or dist is None:
        return list(bufs_all[:per_rank])
    if rank == 0:
        assert len(bufs_all) == n
        sizes = torch.tensor([len(b) for b in bufs_all], dtype=torch.int64, device=device)
        blob = torch.frombuffer(bytearray(b"".join(bufs_all)), dtype=torch.uint8).to(device)
        total = torch.tensor([blob.numel()], dtype=torch.int64, device=device)
    else:
        sizes = torch.zeros(n, dtype=torch.int64, device=device)
        total = torch.zeros(1, dtype=torch.int64, device=device)
    dist.broadcast(sizes, 0)
    dist.broadcast(total, 0)
    if rank != 0:
        blob = torch.empty(int(total.item()), dtype=torch.uint8, device=device)
    dist.broadcast(blob, 0)
    host = blob.cpu().numpy().tobytes()
    offs = [0]
    for s in sizes.tolist():
        offs.append(offs[-1] + s)
    return [host[offs[i]:offs[i + 1]] for i in shard_range(per_rank, rank)]


def max_over_ranks(x, dist=None, device="cpu"):
    if dist is None:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _nal_units(data):
    """(offset of the start code, offset of the NAL header byte) of every NAL unit of an Annex-B byte string."""
    out, i, n = [], 0, len(data)
    while True:
        j = data.find(b"\x00\x00\x01", i)
        if j < 0:
            break
        start = j - 1 if j > 0 and data[j - 1] == 0 else j
        out.append((start, j + 3))
        i = j + 3
    return out


def _first_ue(b):
    """first Exp-Golomb code of a byte string (pic_parameter_set_id of a PPS payload)"""
    bits = "".join(f"{x:08b}" for x in b)
    z = len(bits) - len(bits.lstrip("0"))
    return int(bits[z:2 * z + 1] or "1", 2) - 1 if 2 * z + 1 <= len(bits) else -1


def split_closed_gops(data):
    """Cut one Annex-B stream into independently decodable pieces at its IDR pictures (closed GOPs: an IDR picture
    empties the decoded picture buffer, nothing after it refers to anything before it — 8.2.1, 8.2.5.1; the reference
    decodes such a stream serially, SURVEY.md section 8e names the GOP as the unit that may go to another GPU).
    Every piece starts with the parameter sets in force at its IDR picture (re-inserted when the stream does not
    repeat them), so a fresh decoder instance can take it."""
    units = _nal_units(data)
    sps, pps = None, {}
    cuts = []          # (start offset, parameter sets to prepend)
    run_start = None   # first of the non-VCL NAL units directly in front of the current one
    seen_sets_in_run = False
    for k, (start, hdr) in enumerate(units):
        if hdr >= len(data):
            break
        typ = data[hdr] & 31
        end = units[k + 1][0] if k + 1 < len(units) else len(data)
        if typ in (6, 7, 8, 9):
            if run_start is None:
                run_start, seen_sets_in_run = start, False
            if typ == 7:
                sps = data[start:end]; seen_sets_in_run = True
            elif typ == 8:
                pps.pop(_first_ue(data[hdr + 1:hdr + 6]), None)               # the latest PPS of an id, in arrival order
                pps[_first_ue(data[hdr + 1:hdr + 6])] = data[start:end]
            continue
        if typ == 5 and hdr + 1 < len(data) and (data[hdr + 1] & 0x80):      # first_mb_in_slice == 0: first slice of an IDR picture
            at = run_start if run_start is not None else start
            prefix = b"" if seen_sets_in_run else (sps or b"") + b"".join(pps.values())
            cuts.append((at, prefix))
        run_start, seen_sets_in_run = None, False
    if not cuts:
        return [bytes(data)]
    pieces = []
    for i, (at, prefix) in enumerate(cuts):
        end = cuts[i + 1][0] if i + 1 < len(cuts) else len(data)
        body = data[at:end]
        if i == 0 and at > 0:
            body = data[:end]                                        # whatever precedes the first IDR picture stays with it
            prefix = b""
        pieces.append(bytes(prefix + body))
    return pieces
