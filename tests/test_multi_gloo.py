"""CPU, world_size 2, gloo: the N > 1 data path of bench.py / compress_multi -- each rank owns one shard (the
reference's get_range split), compresses it seeing the left window halo, rank 0 gathers the byte-aligned shard
streams and the concatenation must decode to the input.  The per-shard compute here is the CPU model (the CUDA
library needs a GPU); what is under test is the sharding, seam flags and gather logic."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rust_brotli_b200  # noqa: F401
    from rust_brotli_b200 import sharding
    from tools.model_harness import Model
    data = open(os.path.join(ROOT, "tests", "golden", "random_then_unicode"), "rb").read()
    a, b = sharding.get_range(rank, world, len(data))
    first, last, align = sharding.shard_flags(rank, world)
    part, _ = Model().compress_range(data, a, b - a, 5, 22, first, last, align)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(part)], dtype=torch.int64))
    cap = int(max(s.item() for s in sizes))
    buf = torch.zeros(cap, dtype=torch.uint8)
    buf[: len(part)] = torch.frombuffer(bytearray(part), dtype=torch.uint8)
    gathered = [torch.zeros(cap, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0)
    if rank == 0:
        parts = [bytes(gathered[i][: int(sizes[i].item())].numpy()) for i in range(world)]
        from oracle.harness import sys_decompress
        out = sharding.concat_shards(parts)
        q.put((sys_decompress(out, len(data)) == data, len(out)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_compress():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, size = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok
    assert size <= 151857  # test_threading.rs bound for 2 threads


def test_get_range_matches_reference_rule():
    sys.path.insert(0, ROOT)
    import rust_brotli_b200  # noqa: F401
    from rust_brotli_b200 import sharding
    total = 272666
    for n in (1, 2, 3, 5, 16):
        rs = [sharding.get_range(i, n, total) for i in range(n)]
        assert rs[0][0] == 0 and rs[-1][1] == total
        assert all(rs[i][1] == rs[i + 1][0] for i in range(n - 1))
