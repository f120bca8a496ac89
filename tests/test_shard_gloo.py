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


# ---- the C library's own split (tb200_shard_range, device_api_comm.inc): host logic, no GPU ----
def test_c_library_shard_ranges_cover_the_stream():
    import trico_b200
    d = trico_b200.Device.__new__(trico_b200.Device)          # no context: tb200_shard_range is pure host arithmetic
    d.lib = trico_b200.load()
    d._ck = lambda ok: (_ for _ in ()).throw(AssertionError("tb200_shard_range failed")) if not ok else None
    for ty, count, l2 in ((1, 0, 9), (1, 1, 9), (1, 513, 9), (3, 200003, 14), (13, 70001, 13), (17, 10**6, 14), (2, 125000000 * 8, 8), (4, 99, 13)):
        lay_pc = 3 if ty in (3, 4) else 1
        for world in (1, 2, 3, 8):
            pos = 0
            for r in range(world):
                f, n = d.shard_range(ty, count, r, world, l2)
                assert f == pos, (ty, count, world, r)
                pos += n
                if r + 1 < world and pos < count:
                    assert (pos * lay_pc) % (1 << l2) == 0          # a share ends on a chunk boundary
            assert pos == count


def _worker_c_split(rank, world, port, stream_type, data, count, log2c, out_q):
    """every rank: its share from the C library's split, encoded by the CPU oracle; rank 0 assembles
    [header | size tables in rank order | payloads in rank order] exactly as device_api_comm.inc does"""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    import torch
    import torch.distributed as dist
    import trico_b200
    from checkers import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = Oracle()
    lib = trico_b200.load()
    import ctypes as C
    f, n = C.c_uint32(), C.c_uint32()
    assert lib.tb200_shard_range(stream_type, count, log2c, rank, world, C.byref(f), C.byref(n))
    lay = oracle.layout(stream_type)
    arity = lay["ncomp"] if lay["codec"] == 1 else lay["per_count"]
    piece = oracle.v1_write_stream(stream_type, data[f.value * arity:(f.value + n.value) * arity], n.value, log2c, 2, 4)
    nch = lib.tb200_v1_nchunks(stream_type, n.value, log2c)
    table, payload = piece[15:15 + 2 * nch], piece[15 + 2 * nch:]
    # size exchange (the all-gather of device_api_comm.inc), then the gather of the shares on rank 0
    mine = torch.tensor([len(payload), nch], dtype=torch.int64)
    allp = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allp, mine)
    gathered = [None] * world
    dist.gather_object((table, payload), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        assert sum(int(p[1]) for p in allp) == lib.tb200_v1_nchunks(stream_type, count, log2c)
        total = sum(int(p[0]) for p in allp)
        info = 0x12 if lay["codec"] == 1 else 0
        head = bytes([stream_type]) + count.to_bytes(4, "little") + bytes([info, log2c]) + total.to_bytes(8, "little")
        out_q.put(head + b"".join(t for t, _ in gathered) + b"".join(p for _, p in gathered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("stream_type,n,log2c", [(1, 5000, 7), (3, 4099, 10), (13, 70001, 12), (20, 3000, 9)])
def test_c_library_split_assembles_to_the_single_rank_stream(oracle, stream_type, n, log2c):
    import torch.multiprocessing as mp
    from trico_b200 import STREAM_DTYPES
    rng = np.random.default_rng(n + 1)
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
    procs = [ctx.Process(target=_worker_c_split, args=(r, 2, port, stream_type, data, n, log2c, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == want
