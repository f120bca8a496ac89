"""Chunk sharding across real GPUs (NCCL): the stream assembled from N ranks' pieces is byte-for-
byte the stream one GPU produces.  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, stream_type, data, count, log2c, out_q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    import trico_b200
    from trico_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = trico_b200.Device(rank)
    lay = dev.layout(stream_type)
    ncomp = lay["ncomp"] if lay["codec"] == 1 else 1
    nsub = lay["ncomp"] if lay["codec"] == 1 else lay["wordsize"]
    piece_type = shard.PIECE_TYPE.get(stream_type, stream_type)

    def encode_range(lo, hi):
        return dev.encode_stream(piece_type, data[lo * ncomp:hi * ncomp], hi - lo, log2c)

    stream, counts, base = shard.encode_sharded(dist, stream_type, count, lay["per_count"], log2c, encode_range, nsub,
                                                device=torch.device("cuda", rank))
    if rank == 0:
        out_q.put(stream)
    dist.barrier()
    dist.destroy_process_group()
    dev.close()


@pytest.mark.parametrize("stream_type,n,log2c", [(1, 300007, 9), (3, 200003, 14)])
def test_sharded_equals_single_gpu(stream_type, n, log2c):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    import trico_b200
    from trico_b200 import STREAM_DTYPES
    world = min(torch.cuda.device_count(), 4)
    rng = np.random.default_rng(n)
    dev = trico_b200.Device(0)
    lay = dev.layout(stream_type)
    arity = lay["ncomp"] if lay["codec"] == 1 else lay["per_count"]
    dt = np.dtype(STREAM_DTYPES[stream_type])
    if dt.kind == "f":
        data = (np.cumsum(rng.standard_normal(n * arity)) * 0.01).astype(dt)
    else:
        data = (np.repeat(np.arange(n), arity) + rng.integers(0, 9, n * arity)).astype(dt)
    want = dev.encode_stream(stream_type, data, n, log2c)
    dev.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, stream_type, data, n, log2c, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == want
