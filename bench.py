#!/usr/bin/env python
"""bench.py - encode/decode throughput of the trico hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2|C1|C3|C4|C5|bunny]   # our sm_100a path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]                   # reference CPU library

A "step" = encode every stream of the workload and decode it again.  The headline line is always
BASELINE.json configs[1] ("C2": a synthetic 100,010,000-vertex / 199,980,000-triangle float mesh
with uint32 indices) unless --config names another one; with N > 1 every rank holds
its own mesh of that size (the path shards by chunk with no data-path collective; only the
compressed-size exchange is a collective), so scaling is weak and `value` is the aggregate.

value  = uncompressed GB moved per second over the step, inputs and outputs resident in HBM:
         2 * raw_bytes / (t_encode + t_decode); encode_gbs / decode_gbs are reported beside it.
e2e    = the same metric through the drop-in C API (trico_write_* / trico_read_*) with pinned HOST
         buffers, host<->device copies inside the timed region.
roofline     = the dominant kernel's algorithmic bytes (raw + compressed of its stream) / its CUDA-
               event time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
cpu_baseline = the unmodified reference (oracle/_ref/libtrico_ref.so) on this box's host cores over
               the SAME arrays (cut into one shard per thread), plus a single-thread figure.
configs      = (N = 1) the other BASELINE configurations, device resident: C1 bunny, C3 double mesh,
               C4 one GPU's share of the point cloud, C5 the 1024-mesh batch, and the bunny tiled
               to 100 M triangles (a real mesh's index planes at bench size).
multi_gpu    = (N > 1) the sharded paths north_star names: C4 (one logical stream cut by chunk
               range over the ranks, sizes exchanged and shares gathered over NCCL; assembly timed
               separately) and C5 (whole meshes dealt to the ranks), and a byte-identity check of a
               sharded stream against the same stream encoded on one GPU.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "encode/decode GB/s (uncompressed bytes) at 1/2/4/8 B200 + compression ratio"
HBM_FALLBACK_GBS = 6650.0              # B200_PROFILING.md fallback if MEASURED_PEAKS.json is absent
C2_WORKLOAD = "C2 synthetic float mesh: 100010000 vertices + 199980000 uint32 triangles per GPU (10000x10001 jittered grid, ids shuffled in blocks of 64)"
TRAFFIC_CMD = ("ncu --set full --clock-control none -k regex:'fpc_|lz4_' -c 12 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-extras "
               "(dram__bytes_read.sum + dram__bytes_write.sum per launch; summaries under profiles/)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5", "bunny"])
    ap.add_argument("--grid", default=None, help="WxH override of the C2 grid (debugging only; invalidates the number)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload")
    return ap.parse_args()


def config_dict(world, devname=None, l2v=9, l2t=14):
    return {"workload": C2_WORKLOAD, "raw_bytes_per_gpu": 3599880000, "fpc_chunk_values": 1 << l2v, "lz4_block_bytes": 1 << l2t,
            "fpc_exponents": [2, 4], "lz4_hash_entries": 1024, "e2e_host_pipeline": "32 MiB slabs, H2D / kernels / D2H on three streams",
            "l2": "inputs (3.6 GB) are larger than L2; no flush needed",
            "sharding": "every rank a whole mesh, chunk-sharded, size exchange only", "gpu": devname or "NVIDIA B200"}


# ------------------------------------------------------------------------------------------------
# reference CPU arm (TEST INFRASTRUCTURE: the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def _ref_lib():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from checkers import Ref, have_ref
    if not have_ref():
        raise RuntimeError("oracle/_ref/libtrico_ref.so is missing (build it with `make -C oracle` where /root/reference is mounted)")
    return Ref().lib


def cpu_reference_shards(verts, tris, threads: int, passes: int = 1):
    """The unmodified reference over host arrays cut into `threads` contiguous shards, one thread
    each (the library is single-threaded; shards are independent archives):
    trico_write_vertices + trico_write_triangles, then trico_read_* of the same archive.
    -> dict(enc_s, dec_s (slowest thread), raw (bytes per pass), archive (bytes))"""
    import numpy as np
    L = _ref_lib()
    nv, nt = verts.shape[0], tris.shape[0]
    vb = [nv * i // threads for i in range(threads + 1)]
    tb = [nt * i // threads for i in range(threads + 1)]
    enc, dec, size = [0.0] * threads, [0.0] * threads, [0] * threads
    ok = [True] * threads

    def work(i):
        v = verts[vb[i]:vb[i + 1]]
        t = tris[tb[i]:tb[i + 1]]
        vout, tout = np.empty_like(v), np.empty_like(t)
        for _ in range(passes):
            t0 = time.perf_counter()
            a = L.trico_open_archive_for_writing((v.nbytes + t.nbytes) // 2 + 64)
            r1 = L.trico_write_vertices(a, v.ctypes.data_as(C.c_void_p), v.shape[0])
            r2 = L.trico_write_triangles(a, t.ctypes.data_as(C.c_void_p), t.shape[0])
            size[i] = L.trico_get_size(a)
            t1 = time.perf_counter()
            blob = C.string_at(L.trico_get_buffer_pointer(a), size[i])
            L.trico_close_archive(a)
            buf = np.frombuffer(blob, np.uint8)
            t2 = time.perf_counter()
            r = L.trico_open_archive_for_reading(buf.ctypes.data_as(C.c_void_p), size[i])
            pv, pt = C.c_void_p(vout.ctypes.data), C.c_void_p(tout.ctypes.data)
            r3 = L.trico_read_vertices(r, C.byref(pv))
            r4 = L.trico_read_triangles(r, C.byref(pt))
            L.trico_close_archive(r)
            t3 = time.perf_counter()
            enc[i] += t1 - t0
            dec[i] += t3 - t2
            ok[i] = ok[i] and r1 == 1 and r2 == 1 and r3 == 1 and r4 == 1
        ok[i] = ok[i] and vout.tobytes() == v.tobytes() and tout.tobytes() == t.tobytes()

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert all(ok), "the reference library failed its own round trip"
    raw = (verts.nbytes + tris.nbytes) * passes
    return dict(enc_s=max(enc), dec_s=max(dec), raw=raw, archive=sum(size) - 8 * threads, kind="reference")


def _gbs(r):
    return 2 * r["raw"] / (r["enc_s"] + r["dec_s"]) / 1e9, r["raw"] / r["enc_s"] / 1e9, r["raw"] / r["dec_s"] / 1e9


def host_c2_arrays(grid=None):
    """the C2 arrays on the host: generated on the GPU when there is one (seconds), else a smaller
    grid from the numpy twin of the generator"""
    try:
        import torch
        if torch.cuda.is_available():
            from trico_b200 import workloads as W
            s = W.c2(torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))), seed=1, grid=grid or W.C2_GRID)
            v, t = s[0][2].cpu().numpy(), s[1][2].cpu().numpy().view("uint32")
            del s
            torch.cuda.empty_cache()
            return v, t, "whole C2 mesh"
    except Exception:
        pass
    from trico_b200.synth import grid_mesh
    v, t = grid_mesh(4000, 2001, jitter=1.0, seed=1)
    return v, t, "8 M-vertex grid of the C2 generator (no GPU to generate the full mesh)"


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    import numpy as np
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    verts, tris, what = host_c2_arrays()
    nv, nt = verts.shape[0], tris.shape[0]
    # every step: `threads` shards of 2 M vertices / 4 M triangles, consecutive slices of the C2 arrays
    sv, st = min(2_000_000, nv // threads), min(4_000_000, nt // threads)
    groups = max(1, min(nv // (sv * threads), nt // (st * threads)))

    def step(s):
        g = s % groups
        return cpu_reference_shards(verts[g * sv * threads:(g + 1) * sv * threads], tris[g * st * threads:(g + 1) * st * threads], threads)

    for s in range(args.warmup):
        step(s)
    vals, encs, decs, raw, arch = [], [], [], 0, 0
    t0 = time.perf_counter()
    for s in range(args.steps):
        r = step(args.warmup + s)
        v, e, d = _gbs(r)
        vals.append(v); encs.append(e); decs.append(d); raw += r["raw"]; arch += r["archive"]
    wall = time.perf_counter() - t0
    val = sum(vals) / len(vals)
    sample = (f"every step {threads} threads x one shard of {sv} vertices / {st} triangles each: consecutive slices of the {what} "
              f"({groups} steps cover {100.0 * min(1.0, groups * sv * threads / nv):.0f} % of it)")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(wall / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config_dict(args.gpus),
        "encode_gbs": round(sum(encs) / len(encs), 4), "decode_gbs": round(sum(decs) / len(decs), 4), "ratio": round(raw / arch, 4),
        "cpu_baseline": {"value": round(val, 4), "unit": "GB/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import trico_b200
        self.torch, self.tb = torch, trico_b200
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist_mod
            self.dist = dist_mod
            # NCCL may print a banner on stdout while it initialises; stdout must carry exactly one JSON
            # line, so fd 1 points at stderr until the first collective has run
            sys.stdout.flush()
            saved_fd = os.dup(1)
            os.dup2(2, 1)
            try:
                self.dist.init_process_group("nccl", device_id=self.dev)
                self.dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_fd, 1)
                os.close(saved_fd)
        self.devname = torch.cuda.get_device_name(self.local)
        self.lib = trico_b200.load()
        # a dedicated (non-default) torch stream: the library launches on it and torch.cuda.Event records
        # on it, so the events bracket exactly our kernels
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        assert self.stream.cuda_stream != 0
        self.d = trico_b200.Device(self.local, self.stream.cuda_stream)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def sum_over_ranks(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t]

    # ---- one workload, device resident -----------------------------------------------------------
    def prepare(self, streams):
        torch, lib = self.torch, self.lib
        P = []
        for name, ty, ten, cnt in streams:
            l2 = lib.tb200_default_log2_chunk(ty, cnt)
            cap = lib.tb200_v1_stream_bound(ty, cnt, l2)
            P.append(dict(name=name, type=ty, ten=ten, count=cnt, l2=l2, cap=cap, raw=ten.numel() * ten.element_size(),
                          enc=torch.empty(cap, dtype=torch.uint8, device=self.dev), out=torch.empty_like(ten)))
        sizes = torch.zeros(max(len(P), 1), dtype=torch.int64, device=self.dev)
        for i, p in enumerate(P):
            p["sz_ptr"] = sizes.data_ptr() + 8 * i
        return P, sizes

    def encode_all(self, P):
        d = self.d
        for p in P:
            d.encode_stream_device(p["type"], p["ten"].data_ptr(), p["count"], p["enc"].data_ptr(), p["cap"], p["sz_ptr"], p["l2"])

    def decode_all(self, P):
        d = self.d
        for p in P:
            d.decode_stream_device(p["hdr"], p["enc"].data_ptr(), p["bytes"], p["out"].data_ptr())

    def finish_prepare(self, P, sizes):
        torch = self.torch
        self.encode_all(P)
        torch.cuda.synchronize()
        sz = sizes.cpu().numpy()
        for i, p in enumerate(P):
            p["bytes"] = int(sz[i])
            p["hdr"] = bytes(p["enc"][:15].cpu().numpy())
        self.decode_all(P)
        torch.cuda.synchronize()
        for p in P:
            assert torch.equal(p["out"].reshape(-1).view(torch.uint8), p["ten"].reshape(-1).view(torch.uint8)), f"round trip mismatch in stream {p['name']}"
            p["out"].zero_()

    def measure(self, streams, steps, warmup, exchange=None):
        """-> dict: device-timed encode / decode of the whole workload, max over ranks"""
        torch = self.torch
        P, sizes = self.prepare(streams)
        self.finish_prepare(P, sizes)
        for _ in range(warmup):
            self.encode_all(P)
            if exchange: exchange(sizes)
            self.decode_all(P)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * steps)]
        self.barrier()
        l0 = self.d.launches
        t_wall0 = time.perf_counter()
        for s in range(steps):
            ev[3 * s].record()
            self.encode_all(P)
            if exchange: exchange(sizes)
            ev[3 * s + 1].record()
            self.decode_all(P)
            ev[3 * s + 2].record()
        torch.cuda.synchronize()
        self.barrier()
        t_wall1 = time.perf_counter()
        launches = self.d.launches - l0
        t_enc = sum(ev[3 * s].elapsed_time(ev[3 * s + 1]) for s in range(steps)) / 1e3
        t_dec = sum(ev[3 * s + 1].elapsed_time(ev[3 * s + 2]) for s in range(steps)) / 1e3
        t_enc, t_dec = self.max_over_ranks([t_enc, t_dec])
        for p in P:
            assert torch.equal(p["out"].reshape(-1).view(torch.uint8), p["ten"].reshape(-1).view(torch.uint8)), f"round trip mismatch after timing in stream {p['name']}"
        raw = sum(p["raw"] for p in P)
        comp = sum(p["bytes"] for p in P)
        raw_all, comp_all = self.sum_over_ranks([raw, comp])
        return dict(P=P, sizes=sizes, raw=raw, comp=comp, raw_all=raw_all, comp_all=comp_all, t_enc=t_enc, t_dec=t_dec, steps=steps,
                    launches=launches, wall=(t_wall0, t_wall1))

    @staticmethod
    def summary(m, extra=None):
        tot = m["raw_all"] * m["steps"]
        out = {"encode_gbs": round(tot / m["t_enc"] / 1e9, 2), "decode_gbs": round(tot / m["t_dec"] / 1e9, 2),
               "value": round(2 * tot / (m["t_enc"] + m["t_dec"]) / 1e9, 2), "ratio": round(m["raw_all"] / m["comp_all"], 4),
               "raw_bytes": int(m["raw_all"]), "steps": m["steps"],
               "streams": {}}
        seen = {}
        for p in m["P"]:
            k = p["name"]
            r, c = seen.get(k, (0, 0))
            seen[k] = (r + p["raw"], c + p["bytes"])
        out["streams"] = {k: {"raw_bytes": r, "ratio": round(r / c, 4)} for k, (r, c) in seen.items()}
        if extra:
            out.update(extra)
        return out

    def time_one(self, fn, n):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(); torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n / 1e3

    def per_stream_kernels(self, P, nrep, names):
        """every stream's encode and decode timed alone (CUDA events on the launching stream)"""
        d, K = self.d, {}
        for p in P:
            te = self.time_one(lambda: d.encode_stream_device(p["type"], p["ten"].data_ptr(), p["count"], p["enc"].data_ptr(), p["cap"], p["sz_ptr"], p["l2"]), nrep)
            td = self.time_one(lambda: d.decode_stream_device(p["hdr"], p["enc"].data_ptr(), p["bytes"], p["out"].data_ptr()), nrep)
            alg = p["raw"] + p["bytes"]
            for kind, t in (("encode", te), ("decode", td)):
                K[names.get((p["name"], kind), f"{p['name']} {kind}")] = {"ms": round(t * 1e3, 4), "algorithmic_bytes": alg, "achieved_gbs": round(alg / t / 1e9, 2),
                                                                           "uncompressed_gbs": round(p["raw"] / t / 1e9, 2)}
        return K

    def free(self, m):
        del m["P"], m["sizes"]
        self.torch.cuda.empty_cache()


def e2e_c2(B, streams, steps):
    """the headline metric through trico_write_* / trico_read_* with pinned host buffers"""
    import torch
    from trico_b200 import LIB_PATH
    L = C.CDLL(LIB_PATH)
    L.trico_open_archive_for_writing.restype = C.c_void_p
    L.trico_open_archive_for_writing.argtypes = [C.c_uint64]
    L.trico_open_archive_for_reading.restype = C.c_void_p
    L.trico_open_archive_for_reading.argtypes = [C.c_void_p, C.c_uint64]
    L.trico_get_buffer_pointer.restype = C.c_void_p
    L.trico_get_buffer_pointer.argtypes = [C.c_void_p]
    L.trico_get_size.restype = C.c_uint64
    L.trico_get_size.argtypes = [C.c_void_p]
    L.trico_close_archive.argtypes = [C.c_void_p]
    L.trico_b200_launch_count.restype = C.c_uint64
    L.trico_b200_launch_count.argtypes = [C.c_void_p]
    L.trico_b200_last_error.restype = C.c_char_p
    for fn in ("trico_write_vertices", "trico_write_triangles"):
        getattr(L, fn).restype = C.c_int
        getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    for fn in ("trico_read_vertices", "trico_read_triangles"):
        getattr(L, fn).restype = C.c_int
        getattr(L, fn).argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    os.environ["TRICO_B200_DEVICE"] = str(B.local)
    verts, tris = streams[0][2], streams[1][2]
    nv, nt = verts.shape[0], tris.shape[0]
    raw = verts.numel() * 4 + tris.numel() * 4
    hv, ht = verts.cpu().pin_memory(), tris.cpu().pin_memory()
    hov, hot = torch.empty_like(hv).pin_memory(), torch.empty_like(ht).pin_memory()
    te = td = 0.0
    arch_bytes = launches = 0
    for s in range(1 + steps):            # first pass = warm-up (context + buffer growth)
        B.barrier()
        t0 = time.perf_counter()
        a = L.trico_open_archive_for_writing(raw // 2)
        ok = L.trico_write_vertices(a, hv.data_ptr(), nv) and L.trico_write_triangles(a, ht.data_ptr(), nt)
        assert ok, L.trico_b200_last_error()
        arch_bytes = L.trico_get_size(a)
        ptr = L.trico_get_buffer_pointer(a)
        t1 = time.perf_counter()
        r = L.trico_open_archive_for_reading(ptr, arch_bytes)
        pv, pt = C.c_void_p(hov.data_ptr()), C.c_void_p(hot.data_ptr())
        ok = L.trico_read_vertices(r, C.byref(pv)) and L.trico_read_triangles(r, C.byref(pt))
        assert ok, L.trico_b200_last_error()
        t2 = time.perf_counter()
        if s == steps:
            launches = L.trico_b200_launch_count(a) + L.trico_b200_launch_count(r)
        L.trico_close_archive(r)
        L.trico_close_archive(a)
        if s > 0:
            te += t1 - t0; td += t2 - t1
    assert torch.equal(hov.view(torch.int32), hv.view(torch.int32)) and torch.equal(hot, ht)
    te, td = B.max_over_ranks([te, td])
    tot = raw * B.world * steps
    return {"value": round(2 * tot / (te + td) / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": raw + arch_bytes, "d2h_bytes_per_step": arch_bytes + raw,
            "encode_gbs": round(tot / te / 1e9, 3), "decode_gbs": round(tot / td / 1e9, 3), "steps": steps, "host_memory": "pinned",
            "launches_per_step": int(launches)}


def c1_through_api(B):
    """C1: the bundled bunny through the archive API (trico_write_* / trico_read_*), next to the
    reference's own archive of the same mesh (tests/golden/bunny_full.npz, md5-pinned in the tests)"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from checkers import TricoCApi
    from trico_b200 import LIB_PATH
    z = np.load(os.path.join(ROOT, "tests", "golden", "bunny_full.npz"))
    v, t = np.ascontiguousarray(z["vertices"]), np.ascontiguousarray(z["triangles"])
    api = TricoCApi(LIB_PATH)
    raw = v.nbytes + t.nbytes
    blob = api.encode([(1, v, v.shape[0]), (3, t, t.shape[0])])
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        blob = api.encode([(1, v, v.shape[0]), (3, t, t.shape[0])])
    t1 = time.perf_counter()
    return {"workload": "C1 Stanford bunny (34834 vertices, 69451 triangles) through trico_write_* with pageable host buffers",
            "archive_bytes": len(blob), "ratio": round(raw / (len(blob) - 8), 4), "reference_archive_bytes": int(z["v0_archive"].size),
            "ratio_reference": round(raw / (int(z["v0_archive"].size) - 8), 4), "encode_ms_per_archive": round((t1 - t0) / reps * 1e3, 3)}


def stl_frontend(B):
    """the tools' STL reader on the GPU - facets of the bunny, replicated with offsets, parsed and
    de-duplicated (sort + unique + index remap, trico_io/iostl.c:70-138) device resident; wall clock around
    tb200_stl_dedup, which allocates its scratch and synchronises.  tools/stl_speed.py times the reference beside it."""
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "bunny_full.npz"))
    v, t = z["vertices"], z["triangles"]
    reps = 40
    corners = v[t.reshape(-1)].reshape(-1, 9)
    rec = np.zeros((reps, corners.shape[0], 50), np.uint8)
    for r in range(reps):
        c = corners.copy()
        c[:, 0::3] += np.float32(0.25 * r)
        rec[r, :, 12:48] = c.view(np.uint8).reshape(-1, 36)
    nt = reps * corners.shape[0]
    d = B.d
    d_f, d_v, d_t = d.upload(rec.reshape(-1)), d.alloc(nt * 36), d.alloc(nt * 12)
    best, nv = 1e9, 0
    for it in range(4):
        t0 = time.perf_counter()
        nv = d.stl_dedup_device(d_f.ptr, nt, d_v.ptr, d_t.ptr)
        if it:
            best = min(best, time.perf_counter() - t0)
    # the first copy is the bunny itself: its vertices and indices must come back where the reference put them
    tri = d.download(d_t.ptr, nt * 12).view(np.uint32).reshape(reps, -1, 3)
    ver = d.download(d_v.ptr, nv * 12).view(np.float32).reshape(-1, 3)
    same = bool(np.array_equal(ver[tri[0]], v[t]))
    return {"workload": "binary-STL facets of the bunny x %d (%d triangles) -> indexed mesh, device resident" % (reps, nt),
            "triangles": nt, "vertices": int(nv), "ms": round(best * 1e3, 3), "mtriangles_per_s": round(nt / best / 1e6, 1),
            "stl_gbs": round(nt * 50 / best / 1e9, 2), "sort_passes": int(d.lib.tb200_stl_last_sort_passes()),
            "first_copy_matches_reference_mesh": same}


def extras_single(B, steps):
    """the other BASELINE configurations on one GPU, device resident"""
    from trico_b200 import workloads as W
    out = {}
    try:
        out["C1"] = c1_through_api(B)
    except Exception as ex:
        out["C1"] = {"error": str(ex)}
    for name, make, what in (
            ("C3", lambda: W.c3(B.dev), "50 M-vertex double mesh: double vertices + double normals + double uv + uint64 indices"),
            ("C4_one_shard", lambda: W.c4_shard(B.dev, 0), "one GPU's share of the 1 B-point cloud: 125 M float points + uint32 RGBA"),
            ("bunny_tiled", lambda: W.bunny_tiled(B.dev), "the Stanford bunny replicated with vertex offsets to 100 M triangles (a real mesh's index planes)")):
        try:
            s = make()
            m = B.measure(s, steps, 3)
            out[name] = B.summary(m, {"workload": what})
            B.free(m)
            del s
            B.torch.cuda.empty_cache()
        except Exception as ex:
            out[name] = {"error": str(ex)}
    try:
        out["C5"] = c5_batch(B, max(2, steps // 2))
    except Exception as ex:
        out["C5"] = {"error": str(ex)}
    try:
        out["stl_frontend"] = stl_frontend(B)
    except Exception as ex:
        out["stl_frontend"] = {"error": str(ex)}
    try:
        out["C2_reference_format"] = c2_reference_format(B, steps)
    except Exception as ex:
        out["C2_reference_format"] = {"error": str(ex)}
    return out


def c2_reference_format(B, steps):
    """The C2 mesh written in the REFERENCE's own format on the device (what trico_write_* do under
    trico_b200_set_format(archive, 0)): every vertex component ONE FPC chain (tile-parallel encoder,
    bytes identical to trico_compress), every index plane ONE LZ4 block (merged from the per-block results)."""
    import ctypes as C
    torch, d, lib = B.torch, B.d, B.lib
    from trico_b200 import workloads as W
    (_, _, v, nv), (_, _, t, nt) = W.c2(B.dev)
    vp = C.c_void_p
    bf = (lib.tb200_fpc_v0_bound(4, nv) + 255) & ~255
    bl = (lib.tb200_lz4_v0_bound(t.numel()) + 255) & ~255
    of = torch.empty(3 * bf, dtype=torch.uint8, device=B.dev)
    ol = torch.empty(4 * bl, dtype=torch.uint8, device=B.dev)
    nb = torch.zeros(16, dtype=torch.int64, device=B.dev)

    def enc_f():
        assert lib.tb200_fpc_encode_v0(d.ctx, 4, vp(v.data_ptr()), nv, 3, 3, 4, 10, vp(of.data_ptr()), bf, vp(nb.data_ptr()))

    def enc_l():
        assert lib.tb200_lz4_encode_v0(d.ctx, 4, vp(t.data_ptr()), t.numel(), vp(ol.data_ptr()), bl, vp(nb.data_ptr() + 64))

    for _ in range(2):
        enc_f(); enc_l()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * steps)]
    for s in range(steps):
        ev[3 * s].record(); enc_f(); ev[3 * s + 1].record(); enc_l(); ev[3 * s + 2].record()
    torch.cuda.synchronize()
    tf = sum(ev[3 * s].elapsed_time(ev[3 * s + 1]) for s in range(steps)) / steps / 1e3
    tl = sum(ev[3 * s + 1].elapsed_time(ev[3 * s + 2]) for s in range(steps)) / steps / 1e3
    h = nb.cpu().numpy()
    fbytes = int(h[:2].view("uint32")[:3].sum())
    lbytes = int(h[8:12].view("uint64").sum())
    rawf, rawl = v.numel() * 4, t.numel() * 4
    return {"workload": "C2 in the reference's own (version 0) format, device resident: 3 FPC chains of 100 M floats + 4 whole-plane LZ4 blocks of 600 M bytes",
            "vertices_encode_gbs": round(rawf / tf / 1e9, 1), "triangles_encode_gbs": round(rawl / tl / 1e9, 1),
            "encode_gbs": round((rawf + rawl) / (tf + tl) / 1e9, 1), "ratio": round((rawf + rawl) / (fbytes + lbytes + 8 + 10 + 28), 4),
            "ratio_streams": {"vertices": round(rawf / fbytes, 4), "triangles": round(rawl / lbytes, 4)},
            "steps": steps, "decode": "v0 streams are one serial chain per component / plane: read by the legacy kernels (K4L / K6L), not timed here"}


def c5_batch(B, steps):
    """C5: 1024 meshes with float / u8 / u16 / u64 attribute lists, whole meshes dealt to the ranks by a
    size-balanced greedy.  Every stream of every mesh is its own v1 stream; the rank's streams go
    through the batched entry points (tb200_encode_streams / tb200_decode_streams: concurrent CUDA
    streams, one size read-back, no per-stream synchronisation)."""
    torch, d, lib = B.torch, B.d, B.lib
    from trico_b200 import workloads as W
    import numpy as np
    owner = W.c5_assign(B.world)
    mine = [m for m in range(W.C5_MESHES) if owner[m] == B.rank]
    meshes = W.c5_meshes(B.dev, mine)
    streams = [s for mesh in meshes for s in mesh]
    n = len(streams)
    batch = d.Batch([s[1] for s in streams], [s[2].data_ptr() for s in streams], [s[3] for s in streams])
    raw = sum(s[2].numel() * s[2].element_size() for s in streams)
    arena_cap = d.batch_arena_bytes(batch)
    arena = torch.empty(arena_cap, dtype=torch.uint8, device=B.dev)
    packed = torch.empty(arena_cap, dtype=torch.uint8, device=B.dev)
    sizes = torch.zeros(n, dtype=torch.int64, device=B.dev)
    prefix = torch.zeros(2 * n + 2, dtype=torch.int64, device=B.dev)
    status = torch.zeros(2, dtype=torch.int32, device=B.dev)
    outs = [torch.empty_like(s[2]) for s in streams]

    def encode():
        d.encode_streams(batch, arena.data_ptr(), arena_cap, packed.data_ptr(), arena_cap, sizes.data_ptr(), prefix.data_ptr())

    encode()
    torch.cuda.synchronize()
    h_sizes = sizes.cpu().numpy().astype(np.uint64)
    h_off = prefix[:n].cpu().numpy().astype(np.uint64)
    total = int(prefix[n])
    hp = packed[:total].cpu().numpy()
    headers = b"".join(hp[int(o):int(o) + 15].tobytes() for o in h_off)
    out_ptrs = [o.data_ptr() for o in outs]

    def decode():
        d.decode_streams(headers, packed.data_ptr(), [int(x) for x in h_off], [int(x) for x in h_sizes], out_ptrs, status.data_ptr())

    decode()
    torch.cuda.synchronize()
    assert int(status[0]) == 0, "malformed block in the batch"
    for s, o in zip(streams, outs):
        assert torch.equal(o.reshape(-1).view(torch.uint8), s[2].reshape(-1).view(torch.uint8)), f"batch round trip mismatch in {s[0]}"
    # the packed batch is byte for byte what the per-stream entry point produces (first mesh checked here, all of them in tests/)
    one = torch.empty(int(lib.tb200_v1_stream_bound(streams[0][1], streams[0][3], lib.tb200_default_log2_chunk(streams[0][1], streams[0][3]))), dtype=torch.uint8, device=B.dev)
    nb = torch.zeros(1, dtype=torch.int64, device=B.dev)
    d.encode_stream_device(streams[0][1], streams[0][2].data_ptr(), streams[0][3], one.data_ptr(), one.numel(), nb.data_ptr(), 0)
    torch.cuda.synchronize()
    same = int(nb[0]) == int(h_sizes[0]) and torch.equal(one[:int(nb[0])], packed[:int(nb[0])])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * steps)]
    B.barrier()
    l0 = d.launches
    for s in range(steps):
        ev[3 * s].record(); encode(); ev[3 * s + 1].record(); decode(); ev[3 * s + 2].record()
    torch.cuda.synchronize()
    launches = d.launches - l0
    t_enc = sum(ev[3 * s].elapsed_time(ev[3 * s + 1]) for s in range(steps)) / 1e3
    t_dec = sum(ev[3 * s + 1].elapsed_time(ev[3 * s + 2]) for s in range(steps)) / 1e3
    t_enc, t_dec = B.max_over_ranks([t_enc, t_dec])
    raw_all, comp_all, n_all = B.sum_over_ranks([raw, total, n])
    per_type = {}
    for s, z in zip(streams, h_sizes):
        r, c = per_type.get(s[0], (0, 0))
        per_type[s[0]] = (r + s[2].numel() * s[2].element_size(), c + int(z))
    out = {"workload": f"C5: 1024 meshes (1 K - 287 K vertices, 105 M in all) x 6 streams, {len(mine)} meshes on this rank, one v1 stream per mesh and type, batched entry points",
           "encode_gbs": round(raw_all * steps / t_enc / 1e9, 2), "decode_gbs": round(raw_all * steps / t_dec / 1e9, 2),
           "value": round(2 * raw_all * steps / (t_enc + t_dec) / 1e9, 2), "ratio": round(raw_all / comp_all, 4), "raw_bytes": int(raw_all), "steps": steps,
           "streams_per_step": int(n_all), "launches_per_step": int(launches // steps), "batch_bytes_identical_to_per_stream_path": bool(same),
           "streams": {k: {"raw_bytes": r, "ratio": round(r / c, 4)} for k, (r, c) in per_type.items()}}
    del arena, packed, outs, streams, meshes
    torch.cuda.empty_cache()
    return out


def c4_sharded(B, steps):
    """C4 over N ranks: ONE logical stream per attribute, cut by chunk range (tb200_shard_range); every
    rank encodes its share, sizes are exchanged with ncclAllGather, the shares are gathered on rank 0
    device to device (assembly timed separately), every rank decodes its own share again."""
    torch, d, lib = B.torch, B.d, B.lib
    from trico_b200 import workloads as W
    uid = torch.zeros(128, dtype=torch.uint8, device=B.dev)
    if B.rank == 0:
        uid.copy_(torch.frombuffer(bytearray(d.comm_unique_id()), dtype=torch.uint8))
    B.dist.broadcast(uid, 0)
    comm = d.comm_create(B.rank, B.world, bytes(uid.cpu().numpy()))
    res = {}
    try:
        # ---- byte identity: a small stream sharded over the ranks == the same stream on one GPU ----
        ident = True
        small = W.c4_shard(B.dev, 99, grid=(1500, 1400))
        for name, ty, ten, cnt in small:
            l2 = lib.tb200_default_log2_chunk(ty, cnt)
            f, n = d.shard_range(ty, cnt, B.rank, B.world, l2)
            cap = lib.tb200_v1_stream_bound(ty, cnt, l2)
            out = torch.zeros(cap, dtype=torch.uint8, device=B.dev)
            nb = torch.zeros(2, dtype=torch.int64, device=B.dev)
            esz = ten.element_size() * (ten.shape[1] if ten.dim() > 1 else 1)
            d.encode_stream_sharded(comm, ty, ten.data_ptr() + f * esz, n, cnt, 0, True, out.data_ptr(), cap, nb.data_ptr(), l2)
            if B.rank == 0:
                one = torch.zeros(cap, dtype=torch.uint8, device=B.dev)
                d.encode_stream_device(ty, ten.data_ptr(), cnt, one.data_ptr(), cap, nb.data_ptr() + 8, l2)
                torch.cuda.synchronize()
                a, b = int(nb[0]), int(nb[1])
                ident = ident and a == b and torch.equal(out[:a], one[:b])
        ident = B.sum_over_ranks([0.0 if ident else 1.0])[0] == 0.0
        del small
        # ---- the timed workload: 125 M points + RGBA per rank ----
        shard = W.c4_shard(B.dev, B.rank)
        per = []
        for name, ty, ten, cnt_gen in shard:
            total = 125_000_000 * B.world                 # 1.0 B points at 8 GPUs; every rank generated a few thousand more than its share
            l2 = lib.tb200_default_log2_chunk(ty, total)
            f, n = d.shard_range(ty, total, B.rank, B.world, l2)
            assert n <= cnt_gen, "share larger than what was generated"
            cap = lib.tb200_v1_stream_bound(ty, total, l2) if B.rank == 0 else 0
            per.append(dict(name=name, type=ty, ten=ten, n=n, total=total, l2=l2, cap=cap, esz=ten.element_size() * (ten.shape[1] if ten.dim() > 1 else 1),
                            out=torch.empty(max(cap, 16), dtype=torch.uint8, device=B.dev), back=torch.empty_like(ten)))
        # all ranks must agree on the shares adding up: recompute n from shard_range exactly
        nb = torch.zeros(len(per), dtype=torch.int64, device=B.dev)
        for p in per:
            lay = d.layout(p["type"])
            p["lay"] = lay
        t_enc = t_asm = t_dec = 0.0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for s in range(1 + steps):
            B.barrier()
            e = a = 0.0
            for i, p in enumerate(per):
                me, ma = d.encode_stream_sharded(comm, p["type"], p["ten"].data_ptr(), p["n"], p["total"], 0, True,
                                                 p["out"].data_ptr() if B.rank == 0 else 0, p["cap"], nb.data_ptr() + 8 * i if B.rank == 0 else 0, p["l2"])
                e += me; a += ma
                p["share"] = d.comm_local_share(comm)
                # every rank decodes the share it holds (what it would be handed for decoding)
                ps, ts, pp, pb = p["share"]
                lay = p["lay"]
                ne = p["n"] * lay["per_count"]
                ev0.record()
                if lay["codec"] == 1:
                    ok = lib.tb200_fpc_decode(d.ctx, lay["wordsize"], lay["ncomp"], C.c_void_p(ps), C.c_void_p(pp), pb, ne, p["l2"], 2, 4, C.c_void_p(p["back"].data_ptr()))
                else:
                    ok = lib.tb200_lz4_decode(d.ctx, lay["wordsize"], C.c_void_p(ps), C.c_void_p(pp), pb, ne, p["l2"], C.c_void_p(p["back"].data_ptr()))
                assert ok, lib.tb200_last_error()
                ev1.record()
                torch.cuda.synchronize()
                if s > 0:
                    t_dec += ev0.elapsed_time(ev1) / 1e3
            if s > 0:
                t_enc += e / 1e3; t_asm += a / 1e3
        for p in per:
            k = p["n"] * (p["ten"].shape[1] if p["ten"].dim() > 1 else 1)
            assert torch.equal(p["back"].reshape(-1)[:k].view(torch.uint8), p["ten"].reshape(-1)[:k].view(torch.uint8)), "sharded round trip mismatch"
        raw_local = sum(p["n"] * p["esz"] for p in per)
        t_enc, t_asm, t_dec = B.max_over_ranks([t_enc, t_asm, t_dec])
        raw_all = B.sum_over_ranks([raw_local])[0]
        comp_all = 0
        if B.rank == 0:
            torch.cuda.synchronize()
            comp_all = int(nb.sum())
        res = {"workload": f"C4: {int(per[0]['total'])} float points + uint32 RGBA as ONE stream each, chunk-sharded over {B.world} GPUs (tb200_encode_stream_sharded)",
               "encode_gbs": round(raw_all * steps / t_enc / 1e9, 2), "decode_gbs": round(raw_all * steps / t_dec / 1e9, 2),
               "assemble_ms": round(t_asm / steps * 1e3, 3), "encode_ms": round(t_enc / steps * 1e3, 3),
               "assembled_gbs": round(raw_all * steps / (t_enc + t_asm) / 1e9, 2),
               "ratio": round(raw_all / comp_all, 4) if comp_all else None, "raw_bytes": int(raw_all), "steps": steps,
               "collectives": "ncclAllGather of 2 x u64 per rank and stream (sizes); grouped ncclSend/ncclRecv of the shares to rank 0 (assembly)",
               "sharded_stream_byte_identical_to_single_gpu": bool(ident)}
    finally:
        torch.cuda.synchronize()
        d.comm_destroy(comm)
    return res


def run_b200(args):
    import numpy as np
    from trico_b200 import workloads as W
    from trico_b200.clocks import ClockSampler
    B = Bench(args)
    torch, d, lib = B.torch, B.d, B.lib
    rank, world = B.rank, B.world

    grid = W.C2_GRID
    if args.grid:
        grid = tuple(int(x) for x in args.grid.lower().split("x"))
    makers = {"C2": lambda: W.c2(B.dev, seed=1 + rank, grid=grid), "C3": lambda: W.c3(B.dev, seed=2 + rank),
              "C4": lambda: W.c4_shard(B.dev, rank), "bunny": lambda: W.bunny_tiled(B.dev), "C1": lambda: W.bunny_tiled(B.dev, 1)}
    if args.config == "C5":
        owner = W.c5_assign(world)
        streams = [s for mesh in W.c5_meshes(B.dev, [m for m in range(W.C5_MESHES) if owner[m] == rank]) for s in mesh]
    else:
        streams = makers[args.config]()
    torch.cuda.synchronize()

    def exchange(sizes):
        # the one collective of the replicated path: every rank learns every rank's compressed byte counts
        if B.dist is not None:
            gathered = torch.empty(world * sizes.numel(), dtype=torch.int64, device=B.dev)
            B.dist.all_gather_into_tensor(gathered, sizes.contiguous())

    sampler = ClockSampler(B.local) if rank == 0 else None      # polls from here on; only the samples inside the timed region count
    m = B.measure(streams, args.steps, max(args.warmup, 3), exchange)
    clocks = sampler.stop(*m["wall"]) if sampler else None
    main = B.summary(m)

    # ---- per-kernel timing for the roofline (rank 0, same stream, CUDA events) ----
    kernels = {}
    if rank == 0:
        names = {("vertices", "encode"): "fpc_encode_lanes_kernel<u32,3,1,32>", ("vertices", "decode"): "fpc_decode_kernel<u32,3,1,32>",
                 ("triangles", "encode"): "lz4_encode_kernel<4,10>+lz4_assemble_kernel", ("triangles", "decode"): "lz4_decode_kernel<4>"} if args.config == "C2" else {}
        if len(m["P"]) <= 8:
            kernels = B.per_stream_kernels(m["P"], max(3, min(args.steps, 10)), names)
    P = m["P"]
    l2v, l2t = P[0]["l2"], P[-1]["l2"]
    per_stream_ratio = {p["name"]: round(p["raw"] / p["bytes"], 4) for p in P} if len(P) <= 8 else {}
    B.free(m)

    # ---- e2e through the drop-in C API with pinned host buffers (rank-local, max over ranks) ----
    e2e = None
    if not args.no_e2e and args.config == "C2":
        e2e = e2e_c2(B, streams, max(3, min(args.steps, 6)))

    # ---- CPU baseline on the same arrays (rank 0; the arrays leave the device first) ----
    host_arrays = None
    if rank == 0 and not args.no_cpu and args.config == "C2":
        host_arrays = (streams[0][2].cpu().numpy(), streams[1][2].cpu().numpy().view(np.uint32))
    del streams, P
    torch.cuda.empty_cache()

    # ---- the other configurations / the sharded multi-GPU paths ----
    configs = multi = None
    if not args.no_extras and args.config == "C2" and not args.grid:
        xs = max(2, min(args.steps, 4))
        if world == 1:
            configs = extras_single(B, xs)
        else:
            multi = {}
            for key, fn in (("C4_sharded", c4_sharded), ("C5_batch", c5_batch)):
                try:
                    multi[key] = fn(B, xs)
                except Exception as ex:
                    multi[key] = {"error": str(ex)}
                torch.cuda.empty_cache()

    # C5 is a batch of thousands of small streams: its figure is the batched entry points' (the per-stream
    # loop above is kept beside it)
    batched = None
    if args.config == "C5":
        batched = c5_batch(B, max(2, args.steps // 2))

    if rank != 0:
        B.barrier()
        if B.dist is not None:
            B.dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
    roofline = None
    if kernels:
        dom = max(kernels, key=lambda k: kernels[k]["ms"])
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(dom)
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": round(kernels[dom]["achieved_gbs"] / peak, 4), "traffic": traffic,
                    "traffic_source": "profiles/traffic.json: " + TRAFFIC_CMD, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                    "encode_frac_of_peak": round(main["encode_gbs"] / world * (1 + 1 / main["ratio"]) / peak, 4),
                    "decode_frac_of_peak": round(main["decode_gbs"] / world * (1 + 1 / main["ratio"]) / peak, 4),
                    "all_kernels": {k: {"ms": v["ms"], "achieved_gbs": v["achieved_gbs"], "frac": round(v["achieved_gbs"] / peak, 4)} for k, v in kernels.items()}}

    # ---- CPU baseline: the unmodified reference on this box's cores, over the same arrays ----
    cpu = None
    ratio_ref = None
    if host_arrays is not None:
        try:
            cores = os.cpu_count() or 1
            threads = max(1, min(cores, 64))
            hv, ht = host_arrays
            r = cpu_reference_shards(hv, ht, threads)
            v, e, dd = _gbs(r)
            ratio_ref = r["raw"] / r["archive"]
            one = cpu_reference_shards(hv[:2_000_000], ht[:4_000_000], 1)
            v1, e1, d1 = _gbs(one)
            cpu = {"value": round(v, 4), "unit": "GB/s", "cores": threads, "kind": "reference",
                   "encode_gbs": round(e, 4), "decode_gbs": round(dd, 4),
                   "single_thread": {"value": round(v1, 4), "encode_gbs": round(e1, 4), "decode_gbs": round(d1, 4),
                                     "sample": "one thread over the first 2,000,000 vertices / 4,000,000 triangles of the same arrays"},
                   "sample": f"the whole C2 mesh of rank 0 (the arrays the GPU arm encoded), cut into {threads} contiguous shards, one thread and one archive each, one pass"}
        except Exception as ex:      # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}

    cfg = config_dict(world, B.devname, l2v, l2t) if args.config == "C2" else {"workload": args.config, "gpu": B.devname}
    line = {
        "metric": METRIC, "value": main["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round((m["t_enc"] + m["t_dec"]) / args.steps * 1e3, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": cfg,
        "encode_gbs": main["encode_gbs"], "decode_gbs": main["decode_gbs"],
        "ratio": main["ratio"], "ratio_streams": per_stream_ratio,
        "ratio_reference_same_arrays": None if ratio_ref is None else round(ratio_ref, 4),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(m["launches"]), "clocks": clocks,
    }
    if batched is not None:
        line["per_stream_loop"] = {"value": main["value"], "encode_gbs": main["encode_gbs"], "decode_gbs": main["decode_gbs"]}
        line["value"], line["encode_gbs"], line["decode_gbs"] = batched["value"], batched["encode_gbs"], batched["decode_gbs"]
        line["ms_per_step"] = round(2 * batched["raw_bytes"] / batched["value"] / 1e9 * 1e3, 4)
        line["gpu_launches"] = int(batched["launches_per_step"] * batched["steps"])
        line["config"] = dict(cfg, batch=batched["workload"])
    if configs is not None:
        line["configs"] = configs
    if multi is not None:
        line["multi_gpu"] = multi
    print(json.dumps(line))
    B.barrier()
    if B.dist is not None:
        B.dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
