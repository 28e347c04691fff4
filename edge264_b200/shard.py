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
