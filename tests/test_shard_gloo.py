"""world_size-2 gloo test of the multi-GPU host logic (no GPU): broadcast of the input buffer, sharding,
max-over-ranks timing reduction, and that both ranks' shards decode to the golden output on the CPU oracle."""
import hashlib, os, sys
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import ROOT, STREAMS, make_stream


def _worker(rank, world, port, paths, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from edge264_b200.shard import broadcast_streams, max_over_ranks
    from checkers import decode_bytes
    bufs_all = [open(p, "rb").read() for p in paths] if rank == 0 else None
    mine = broadcast_streams(bufs_all, len(paths) // world, world, rank, dist)
    digests = []
    for b in mine:
        frames, _ = decode_bytes(b, "port")
        digests.append([hashlib.md5(f[3]).hexdigest() for f in frames])
    t = max_over_ranks(float(rank + 1), dist)
    torch.save({"digests": digests, "sizes": [len(b) for b in mine], "tmax": t}, os.path.join(out, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_broadcast_and_shard_two_ranks(workdir, tmp_path):
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "streams.json")))
    sel = [STREAMS[0], STREAMS[4], STREAMS[6], STREAMS[8]]
    paths = [make_stream(workdir, *s) for s in sel]
    mp.spawn(_worker, args=(2, 29517, paths, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        r = torch.load(os.path.join(str(tmp_path), f"r{rank}.pt"))
        assert r["tmax"] == 2.0
        for k in range(2):
            name = sel[rank * 2 + k][0]
            assert r["sizes"][k] == os.path.getsize(paths[rank * 2 + k])
            assert r["digests"][k] == gold[name]["md5"]


def _gop_worker(rank, world, port, path, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from edge264_b200.shard import broadcast_streams, split_closed_gops
    from checkers import decode_bytes
    gops = split_closed_gops(open(path, "rb").read()) if rank == 0 else None     # rank 0 cuts the one stream at its IDR pictures
    mine = broadcast_streams(gops, 2, world, rank, dist)
    digests = [[hashlib.md5(f[3]).hexdigest() for f in decode_bytes(g, "port")[0]] for g in mine]
    torch.save({"digests": digests}, os.path.join(out, f"g{rank}.pt"))
    dist.destroy_process_group()


def test_one_stream_sharded_by_closed_gop(workdir, tmp_path):
    """SURVEY section 8e / BASELINE configs[4]: ONE stream goes to several GPUs GOP by GOP.  Cutting at the IDR pictures
    and decoding every piece with a fresh decoder must give exactly the frames of the serial decode, in order."""
    from checkers import decode_bytes
    from edge264_b200.shard import split_closed_gops
    path = make_stream(workdir, "gop_shard", 9, 7, "-n 24 -s 5 --gop IPB --idr 6 --refs 2 --deblock 0 --wp 1")
    data = open(path, "rb").read()
    whole = [hashlib.md5(f[3]).hexdigest() for f in decode_bytes(data, "port")[0]]
    assert len(split_closed_gops(data)) == 4
    mp.spawn(_gop_worker, args=(2, 29519, path, str(tmp_path)), nprocs=2, join=True)
    got = []
    for rank in range(2):
        for d in torch.load(os.path.join(str(tmp_path), f"g{rank}.pt"))["digests"]:
            got += d
    assert got == whole
    # a stream that repeats no parameter sets at its later IDR pictures gets them re-inserted
    nals = data.split(b"\0\0\0\1")[1:]
    first_idr = next(i for i, n in enumerate(nals) if (n[0] & 31) == 5)
    stripped = b"".join(b"\0\0\0\1" + n for i, n in enumerate(nals) if i < first_idr or (n[0] & 31) not in (7, 8))
    pieces = split_closed_gops(stripped)
    assert len(pieces) == 4
    serial = [hashlib.md5(f[3]).hexdigest() for f in decode_bytes(stripped, "port")[0]]    # (the generator varies its PPS: another stream than `data`)
    again = []
    for p in pieces:
        again += [hashlib.md5(f[3]).hexdigest() for f in decode_bytes(p, "port")[0]]
    assert again == serial and len(serial) == 24
