"""Multi-rank host logic (chunk sharding, size exchange, assembly) on CPU: world_size 2 over gloo.
The per-rank encoder is the CPU oracle here; the GPU twin of this test is
tests/test_gpu_multi.py (NCCL, needs >= 2 GPUs)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, stream_type, data, count, log2c, out_q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import torch.distributed as dist
    from checkers import Oracle
    from trico_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = Oracle()
    lay = oracle.layout(stream_type)
    ncomp = lay["ncomp"] if lay["codec"] == 1 else 1
    nsub = lay["ncomp"] if lay["codec"] == 1 else lay["wordsize"]
    piece_type = shard.PIECE_TYPE.get(stream_type, stream_type)

    def encode_range(lo, hi):
        return oracle.v1_write_stream(piece_type, data[lo * ncomp:hi * ncomp], hi - lo, log2c, 2, 4)

    stream, counts, base = shard.encode_sharded(dist, stream_type, count, lay["per_count"], log2c, encode_range, nsub)
    assert base == sum(counts[:rank])
    if rank == 0:
        out_q.put(stream)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("stream_type,n,log2c", [(1, 5000, 7), (3, 4099, 10), (16, 1000, 6), (13, 70001, 12), (1, 100, 9)])
def test_sharded_stream_equals_single_rank_stream(oracle, stream_type, n, log2c):
    import torch.multiprocessing as mp
    from trico_b200 import STREAM_DTYPES
    rng = np.random.default_rng(n)
    lay = oracle.layout(stream_type)
    arity = lay["ncomp"] if lay["codec"] == 1 else lay["per_count"]
    dt = np.dtype(STREAM_DTYPES[stream_type])
    if dt.kind == "f":
        data = (np.cumsum(rng.standard_normal(n * arity)) * 0.01).astype(dt)
    else:
        data = (np.repeat(np.arange(n), arity) + rng.integers(0, 9, n * arity)).astype(dt)
    want = oracle.v1_write_stream(stream_type, data, n, log2c, 2, 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, stream_type, data, n, log2c, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == want
    t, c, arr, used = oracle.v1_read_stream(b"Trco\x01\0\0\0" + got, 8)
    assert (t, c) == (stream_type, n) and arr.tobytes() == data.tobytes()


def test_plan_ranges_are_chunk_aligned():
    from trico_b200.shard import plan_ranges
    for n, l2, w in ((0, 9, 4), (1, 9, 8), (512, 9, 2), (513, 9, 2), (10**9, 9, 8), (300007, 14, 3)):
        r = plan_ranges(n, l2, w)
        assert len(r) == w and r[0][0] == 0 and r[-1][1] == n
        for (lo, hi), (lo2, _) in zip(r, r[1:]):
            assert hi == lo2 and (hi % (1 << l2) == 0 or hi == n)
