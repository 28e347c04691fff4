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
