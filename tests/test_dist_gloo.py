"""N>1 host plumbing on CPU (gloo, world_size 2): stream sharding and the all_gather of per-stream
{size, checksum}. The per-stream payload here comes from the CPU oracle (tiny streams) because this
test runs without a GPU; the GPU path uses the same shard/gather code in bench.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gmix_b200 import shard  # noqa: E402


def _streams():
    dic = open(os.path.join(ROOT, "tests", "data", "english.dic"), "rb").read()
    return [dic[i * 211:i * 211 + 20 + 37 * (i % 4)] for i in range(7)] + [b""]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib
    oracle = oracle_lib.load()
    streams = _streams()
    ranges = shard.shard_ranges([len(s) for s in streams], world)
    lo, hi = ranges[rank]
    comp = [oracle.compress(s) for s in streams[lo:hi]]
    sizes = torch.tensor([len(c) for c in comp], dtype=torch.int64)
    sums = torch.tensor([np.int64(np.uint64(shard.fnv1a64(c))) for c in comp], dtype=torch.int64)
    all_sizes, all_sums = shard.gather_sizes_checksums(sizes, sums, [b - a for a, b in ranges])
    if rank == 0:
        q.put((ranges, all_sizes.tolist(), all_sums.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_and_balance():
    lens = [100, 5, 5, 5, 100, 1, 0, 90]
    for world in (1, 2, 3, 8):
        r = shard.shard_ranges(lens, world)
        assert r[0][0] == 0 and r[-1][1] == len(lens)
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
    two = shard.shard_ranges(lens, 2)
    assert abs(sum(lens[two[0][0]:two[0][1]]) - sum(lens[two[1][0]:two[1][1]])) <= 100
    assert shard.shard_ranges([], 2) == [(0, 0), (0, 0)]


def test_gather_sizes_and_checksums_world2():
    import oracle_lib
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ranges, sizes, sums = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    oracle = oracle_lib.load()
    comp = [oracle.compress(s) for s in _streams()]
    assert sizes == [len(c) for c in comp]
    assert [x & 0xFFFFFFFFFFFFFFFF for x in sums] == [shard.fnv1a64(c) for c in comp]
    assert ranges[0][1] == ranges[1][0]
