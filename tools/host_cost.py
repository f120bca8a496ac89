#!/usr/bin/env python
"""Host time per tb200_encode_stream / tb200_decode_stream call on small streams (no synchronisation
inside the loop): what a batch of thousands of small streams pays per stream on the submitting thread."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trico_b200
dev = trico_b200.Device(0)
rng = np.random.default_rng(0)
n = 20000
cases = [("float vec3 (FPC)", 1, rng.standard_normal(n * 3).astype(np.float32), n),
         ("u32 triangles (LZ4)", 3, (np.arange(n * 3) // 3).astype(np.uint32), n),
         ("u8 list (LZ4)", 17, (np.arange(n) // 16 % 256).astype(np.uint8), n)]
for name, ty, data, cnt in cases:
    log2c = dev.lib.tb200_default_log2_chunk(ty, cnt)
    bound = dev.lib.tb200_v1_stream_bound(ty, cnt, log2c)
    d_in, d_out, d_sz, d_back = dev.upload(data), dev.alloc(bound), dev.alloc(64), dev.alloc(data.nbytes + 64)
    dev.encode_stream_device(ty, d_in.ptr, cnt, d_out.ptr, bound, d_sz.ptr, log2c); dev.sync()
    nbytes = int(dev.download(d_sz.ptr, 8).view(np.uint64)[0])
    hdr = dev.download(d_out.ptr, 16).tobytes()
    reps = 2000
    l0 = dev.launches
    t0 = time.perf_counter()
    for _ in range(reps): dev.encode_stream_device(ty, d_in.ptr, cnt, d_out.ptr, bound, d_sz.ptr, log2c)
    t1 = time.perf_counter(); dev.sync(); t2 = time.perf_counter()
    le = dev.launches - l0
    t3 = time.perf_counter()
    for _ in range(reps): dev.decode_stream_device(hdr, d_out.ptr, nbytes, d_back.ptr)
    t4 = time.perf_counter(); dev.sync(); t5 = time.perf_counter()
    print(f"{name:22s} encode: {1e6 * (t1 - t0) / reps:6.1f} us host per call ({le / reps:.0f} kernels), {1e6 * (t2 - t0) / reps:6.1f} us with the GPU drained; decode: {1e6 * (t4 - t3) / reps:6.1f} us host, {1e6 * (t5 - t3) / reps:6.1f} us drained")
