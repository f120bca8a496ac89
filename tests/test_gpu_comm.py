"""The C library's multi-GPU path (device_api_comm.inc): chunk-sharded streams over NCCL.
World size 1 exercises the whole code path (NCCL communicator, size exchange, assembly copies) on a
one-GPU box; with >= 2 GPUs the same test runs one process per GPU and the stream assembled on rank 0
must be byte-for-byte the stream one GPU produces.  tb200_decode_stream_range is checked on any box."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(stream_type, n, seed):
    import trico_b200
    from trico_b200 import STREAM_DTYPES
    dev = trico_b200.Device(0)
    lay = dev.layout(stream_type)
    dev.close()
    arity = lay["ncomp"] if lay["codec"] == 1 else lay["per_count"]
    rng = np.random.default_rng(seed)
    dt = np.dtype(STREAM_DTYPES[stream_type])
    if dt.kind == "f":
        return (np.cumsum(rng.standard_normal(n * arity)) * 0.01).astype(dt), arity
    return (np.repeat(np.arange(n), arity) + rng.integers(0, 9, n * arity)).astype(dt), arity


def _sharded_worker(rank, world, uid_path, stream_type, data, arity, count, log2c, out_q):
    import sys
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import trico_b200
    dev = trico_b200.Device(rank)
    # the 128-byte NCCL id travels through a file: the library needs no torch.distributed
    if rank == 0:
        uid = dev.comm_unique_id()
        with open(uid_path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(uid_path + ".tmp", uid_path)
    else:
        for _ in range(600):
            if os.path.exists(uid_path):
                break
            time.sleep(0.05)
        uid = open(uid_path, "rb").read()
    comm = dev.comm_create(rank, world, uid)
    f, n = dev.shard_range(stream_type, count, rank, world, log2c)
    local = np.ascontiguousarray(data[f * arity:(f + n) * arity])
    d_local = dev.upload(local)
    cap = dev.lib.tb200_v1_stream_bound(stream_type, count, log2c)
    d_out = dev.alloc(cap)
    d_nb = dev.alloc(64)
    ms = dev.encode_stream_sharded(comm, stream_type, d_local.ptr, n, count, 0, True, d_out.ptr, cap, d_nb.ptr, log2c)
    assert ms[0] >= 0 and ms[1] >= 0
    # every rank decodes its own share again from what it holds
    ps, ts, pp, pb = dev.comm_local_share(comm)
    lay = dev.layout(stream_type)
    back = dev.alloc(local.nbytes + 64)
    import ctypes as C
    if n:
        if lay["codec"] == 1:
            ok = dev.lib.tb200_fpc_decode(dev.ctx, lay["wordsize"], lay["ncomp"], C.c_void_p(ps), C.c_void_p(pp), pb, n * lay["per_count"], log2c, 2, 4, C.c_void_p(back.ptr))
        else:
            ok = dev.lib.tb200_lz4_decode(dev.ctx, lay["wordsize"], C.c_void_p(ps), C.c_void_p(pp), pb, n * lay["per_count"], log2c, C.c_void_p(back.ptr))
        assert ok, dev.lib.tb200_last_error()
        assert dev.download(back.ptr, local.nbytes).tobytes() == local.tobytes()
    if rank == 0:
        nb = int(dev.download(d_nb.ptr, 8).view(np.uint64)[0])
        out_q.put(dev.download(d_out.ptr, nb).tobytes())
    dev.sync()
    dev.comm_destroy(comm)
    dev.close()


@pytest.mark.parametrize("stream_type,n,log2c", [(1, 300007, 9), (3, 200003, 14), (13, 150001, 13), (2, 70001, 8)])
def test_sharded_stream_through_the_c_library(stream_type, n, log2c, tmp_path):
    import torch
    import torch.multiprocessing as mp
    import trico_b200
    world = max(1, min(torch.cuda.device_count(), 4))
    data, arity = _data(stream_type, n, n)
    dev = trico_b200.Device(0)
    want = dev.encode_stream(stream_type, data, n, log2c)
    dev.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    uid_path = str(tmp_path / "nccl_uid")
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, uid_path, stream_type, data, arity, n, log2c, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == want, f"stream assembled from {world} rank(s) differs from the single-GPU stream"


@pytest.mark.parametrize("stream_type,n,log2c", [(1, 100003, 9), (3, 70001, 14), (18, 50000, 13), (16, 9001, 8)])
def test_decode_of_chunk_ranges(stream_type, n, log2c):
    """a rank that is handed units [first, first + k) of a whole stream decodes exactly those"""
    import trico_b200
    data, arity = _data(stream_type, n, n + 7)
    dev = trico_b200.Device(0)
    stream = dev.encode_stream(stream_type, data, n, log2c)
    d_s = dev.upload(np.frombuffer(stream, np.uint8))
    for world in (1, 2, 3, 5):
        for r in range(world):
            f, k = dev.shard_range(stream_type, n, r, world, log2c)
            if k == 0:
                continue
            want = data[f * arity:(f + k) * arity]
            d_o = dev.alloc(want.nbytes + 64)
            dev.decode_stream_range(stream, d_s.ptr, len(stream), f, k, d_o.ptr)
            assert dev.download(d_o.ptr, want.nbytes).tobytes() == want.tobytes(), (world, r)
            d_o.free()
    # a range that is not chunk aligned is refused
    with pytest.raises(trico_b200.TB200Error):
        dev.decode_stream_range(stream, d_s.ptr, len(stream), 1, 10, d_s.ptr)
    dev.close()
