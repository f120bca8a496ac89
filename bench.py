#!/usr/bin/env python
"""bench.py - encode/decode throughput of the trico hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                  # our sm_100a path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # reference CPU library

A "step" = encode the whole mesh (vertices + triangle indices) and decode it again.  At N = 1 the
workload is BASELINE.json configs[1]: a synthetic 100,010,000-vertex / 199,980,000-triangle float
mesh with uint32 indices (SURVEY.md 8d "C2").  With N > 1 every rank holds its own mesh of that
size (the path shards by chunk with no data-path collective; only the compressed-size exchange is
a collective), so scaling is weak and `value` is the aggregate over ranks.

value  = uncompressed GB moved per second over the step, inputs and outputs resident in HBM:
         2 * raw_bytes / (t_encode + t_decode); encode_gbs / decode_gbs are reported beside it.
e2e    = the same metric through the drop-in C API (trico_write_* / trico_read_*) with pinned HOST
         buffers, host<->device copies inside the timed region.
roofline    = the dominant kernel's algorithmic bytes (raw + compressed of its stream) / its CUDA-
              event time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
cpu_baseline = the unmodified reference (oracle/_ref/libtrico_ref.so) on this box's host cores,
               on a bounded sample of the same generator.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "encode/decode GB/s (uncompressed bytes) at 1/2/4/8 B200 + compression ratio"
FULL_W, FULL_H = 10000, 10001          # 100,010,000 vertices / 199,980,000 triangles
HBM_FALLBACK_GBS = 6650.0              # B200_PROFILING.md fallback if MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", default=None, help="WxH override of the mesh grid (debugging only; invalidates the number)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# reference CPU arm (TEST INFRASTRUCTURE: the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(verts, tris, reps: int, threads: int):
    """Times the unmodified reference library: `threads` workers, each encoding and decoding the
    same sample mesh `reps` times through trico_write_* / trico_read_*.
    -> dict(enc_s, dec_s, raw_bytes (per worker per rep), archive_bytes, kind)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    from checkers import Oracle, Ref, have_ref
    if have_ref():
        lib, kind = Ref(), "reference"
    else:
        raise RuntimeError("oracle/_ref/libtrico_ref.so is missing (build it with `make -C oracle` where /root/reference is mounted)")
    L = lib.lib
    nv, nt = verts.shape[0], tris.shape[0]
    raw = verts.nbytes + tris.nbytes
    res = {"enc": [0.0] * threads, "dec": [0.0] * threads, "size": 0}

    def work(tid):
        vout = np.empty_like(verts)
        tout = np.empty_like(tris)
        for _ in range(reps):
            t0 = time.perf_counter()
            a = L.trico_open_archive_for_writing(raw // 2)
            assert L.trico_write_vertices(a, verts.ctypes.data_as(C.c_void_p), nv) == 1
            assert L.trico_write_triangles(a, tris.ctypes.data_as(C.c_void_p), nt) == 1
            size = L.trico_get_size(a)
            t1 = time.perf_counter()
            blob = C.string_at(L.trico_get_buffer_pointer(a), size)
            L.trico_close_archive(a)
            buf = np.frombuffer(blob, np.uint8)
            t2 = time.perf_counter()
            r = L.trico_open_archive_for_reading(buf.ctypes.data_as(C.c_void_p), size)
            pv, pt = C.c_void_p(vout.ctypes.data), C.c_void_p(tout.ctypes.data)
            assert L.trico_read_vertices(r, C.byref(pv)) == 1
            assert L.trico_read_triangles(r, C.byref(pt)) == 1
            L.trico_close_archive(r)
            t3 = time.perf_counter()
            res["enc"][tid] += t1 - t0
            res["dec"][tid] += t3 - t2
            res["size"] = size
        assert vout.tobytes() == verts.tobytes() and tout.tobytes() == tris.tobytes()

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    wall = time.perf_counter() - t0
    enc_s, dec_s = max(res["enc"]), max(res["dec"])
    return dict(enc_s=enc_s, dec_s=dec_s, wall=wall, raw=raw, reps=reps, threads=threads, archive=res["size"], kind=kind)


def cpu_summary(r):
    total_raw = r["raw"] * r["reps"] * r["threads"]
    enc = total_raw / r["enc_s"] / 1e9
    dec = total_raw / r["dec_s"] / 1e9
    val = 2 * total_raw / (r["enc_s"] + r["dec_s"]) / 1e9
    return val, enc, dec, r["raw"] / (r["archive"] - 8)


def host_sample(W, H):
    from trico_b200.synth import grid_mesh
    return grid_mesh(W, H, jitter=1.0, seed=1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    sw, sh = 2000, 1000
    verts, tris = host_sample(sw, sh)
    reps = 2
    for _ in range(args.warmup):
        cpu_reference_run(verts, tris, 1, threads)
    t0 = time.perf_counter()
    vals, encs, decs, ratio = [], [], [], 0.0
    for _ in range(args.steps):
        r = cpu_reference_run(verts, tris, reps, threads)
        v, e, d, ratio = cpu_summary(r)
        vals.append(v); encs.append(e); decs.append(d)
    wall = time.perf_counter() - t0
    val = sum(vals) / len(vals)
    sample = f"{threads} threads x {reps} passes over a {verts.shape[0]}-vertex / {tris.shape[0]}-triangle mesh from the C2 generator per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(wall / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "C2 synthetic float mesh + uint32 indices (bounded CPU sample)", "sample": sample},
        "encode_gbs": round(sum(encs) / len(encs), 4), "decode_gbs": round(sum(decs) / len(decs), 4), "ratio": round(ratio, 4),
        "cpu_baseline": {"value": round(val, 4), "unit": "GB/s", "cores": threads, "kind": r["kind"], "sample": sample},
        "e2e": {"value": round(val, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 2 ms from a
    thread that is started before the warm-up (the timed region is ~50 ms; `nvidia-smi -lms` needs
    longer than that to print its first line); nvidia-smi is the fallback."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.rows = []            # (time, sm_mhz, set of reasons)
        self.max_mhz = None
        self.stop_flag = False
        self.proc = None
        self.mode = None
        try:
            import pynvml as N
            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = gpu_index
            if vis:
                ids = [v for v in vis.split(",") if v.strip() != ""]
                if gpu_index < len(ids) and ids[gpu_index].strip().isdigit():
                    idx = int(ids[gpu_index])
            self.N, self.h = N, N.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
            self.mode = "nvml"
            self.th = threading.Thread(target=self._poll_nvml, daemon=True)
            self.th.start()
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "nvidia-smi"
            self.th = threading.Thread(target=self._pump_smi, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        N = self.N
        names = (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"))
        bits = [(n, getattr(N, a)) for n, a in names if hasattr(N, a)]
        while not self.stop_flag:
            try:
                clk = float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM))
                r = N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), clk, {n for n, b in bits if r & b}))
            except Exception:
                pass
            time.sleep(0.002)

    def _pump_smi(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.strip().split(",")]
            if len(parts) < 9:
                continue
            try:
                clk = float(parts[1]); self.max_mhz = float(parts[2])
            except ValueError:
                continue
            rs = {name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]) if v.lower().startswith("active")}
            self.rows.append((time.perf_counter(), clk, rs))

    def stop(self, t0, t1):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (NVML and nvidia-smi unavailable)"]}
        if self.mode == "nvidia-smi":
            time.sleep(0.15)
            self.proc.terminate()
        self.stop_flag = True
        inside = [(c, r) for t, c, r in self.rows if t0 <= t <= t1]
        if not inside:                                   # region shorter than the sampling period: the samples around it
            inside = [(c, r) for t, c, r in self.rows if t0 - 0.05 <= t <= t1 + 0.05] or [(c, r) for _, c, r in self.rows[-3:]] or [(0.0, set())]
        sm = sorted(c for c, _ in inside)
        reasons = set().union(*[r for _, r in inside])
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons), "samples": len(sm), "source": self.mode}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import trico_b200
    from trico_b200.synth import grid_mesh_torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # NCCL may print a version banner on stdout while it initialises; stdout must carry exactly one
        # JSON line, so fd 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    devname = torch.cuda.get_device_name(local)

    W, H = (FULL_W, FULL_H)
    if args.grid:
        W, H = (int(x) for x in args.grid.lower().split("x"))
    dev = torch.device("cuda", local)
    verts, tris = grid_mesh_torch(W, H, dev, jitter=1.0, seed=1 + rank)
    torch.cuda.synchronize()
    nv, nt = verts.shape[0], tris.shape[0]
    raw_v, raw_t = verts.numel() * 4, tris.numel() * 4
    raw = raw_v + raw_t

    lib = trico_b200.load()
    # a dedicated (non-default) torch stream: its handle is non-zero, the library launches on it and
    # torch.cuda.Event records on it, so the events bracket exactly our kernels
    bench_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(bench_stream)
    stream = bench_stream.cuda_stream
    assert stream != 0
    d = trico_b200.Device(local, stream)
    l2v = lib.tb200_default_log2_chunk(1, nv)
    l2t = lib.tb200_default_log2_chunk(3, nt)
    cap_v = lib.tb200_v1_stream_bound(1, nv, l2v)
    cap_t = lib.tb200_v1_stream_bound(3, nt, l2t)
    enc_v = torch.empty(cap_v, dtype=torch.uint8, device=dev)
    enc_t = torch.empty(cap_t, dtype=torch.uint8, device=dev)
    sizes = torch.zeros(8, dtype=torch.int64, device=dev)       # [0] vertex stream bytes, [1] triangle stream bytes
    out_v = torch.empty_like(verts)
    out_t = torch.empty_like(tris)

    def encode():
        d.encode_stream_device(1, verts.data_ptr(), nv, enc_v.data_ptr(), cap_v, sizes.data_ptr(), l2v)
        d.encode_stream_device(3, tris.data_ptr(), nt, enc_t.data_ptr(), cap_t, sizes.data_ptr() + 8, l2t)

    encode()
    torch.cuda.synchronize()
    sz = sizes.cpu().numpy()
    bytes_v, bytes_t = int(sz[0]), int(sz[1])
    hdr_v = bytes(enc_v[:15].cpu().numpy())
    hdr_t = bytes(enc_t[:15].cpu().numpy())

    def decode():
        d.decode_stream_device(hdr_v, enc_v.data_ptr(), bytes_v, out_v.data_ptr())
        d.decode_stream_device(hdr_t, enc_t.data_ptr(), bytes_t, out_t.data_ptr())

    def exchange():
        # the one collective of the path: every rank learns every rank's compressed byte counts
        if dist is not None:
            gathered = torch.empty(world * 2, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(gathered, sizes[:2].contiguous())
            return gathered
        return None

    # correctness of what is being timed
    decode()
    torch.cuda.synchronize()
    assert torch.equal(out_v.view(torch.int32), verts.view(torch.int32)) and torch.equal(out_t, tris), "round trip mismatch"
    out_v.zero_(); out_t.zero_()

    sampler = ClockSampler(local) if rank == 0 else None      # polls from here on; only the samples inside the timed region count
    for _ in range(max(args.warmup, 3)):
        encode(); exchange(); decode()
    torch.cuda.synchronize()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps)]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = d.launches
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        ev[3 * s].record()
        encode(); exchange()
        ev[3 * s + 1].record()
        decode()
        ev[3 * s + 2].record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    launches = d.launches - launches0
    t_enc = sum(ev[3 * s].elapsed_time(ev[3 * s + 1]) for s in range(args.steps)) / 1e3
    t_dec = sum(ev[3 * s + 1].elapsed_time(ev[3 * s + 2]) for s in range(args.steps)) / 1e3
    tt = torch.tensor([t_enc, t_dec], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_enc, t_dec = float(tt[0]), float(tt[1])
    assert torch.equal(out_v.view(torch.int32), verts.view(torch.int32)) and torch.equal(out_t, tris), "round trip mismatch after timing"

    total_raw = raw * world * args.steps
    enc_gbs = total_raw / t_enc / 1e9
    dec_gbs = total_raw / t_dec / 1e9
    value = 2 * total_raw / (t_enc + t_dec) / 1e9
    ratio = raw / (bytes_v + bytes_t)

    # ---- per-kernel timing for the roofline (rank 0, same stream, CUDA events) ----
    kernels = {}
    if rank == 0:
        def timeit(fn, n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fn(); torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n / 1e3
        nrep = max(3, min(args.steps, 10))
        specs = {
            "fpc_encode_lanes_kernel<u32,3,1,32>": (lambda: d.encode_stream_device(1, verts.data_ptr(), nv, enc_v.data_ptr(), cap_v, sizes.data_ptr(), l2v), raw_v, bytes_v),
            "lz4_encode_kernel<4,10>+lz4_assemble_kernel": (lambda: d.encode_stream_device(3, tris.data_ptr(), nt, enc_t.data_ptr(), cap_t, sizes.data_ptr() + 8, l2t), raw_t, bytes_t),
            "fpc_decode_kernel<u32,3,1,32>": (lambda: d.decode_stream_device(hdr_v, enc_v.data_ptr(), bytes_v, out_v.data_ptr()), raw_v, bytes_v),
            "lz4_decode_kernel<4>": (lambda: d.decode_stream_device(hdr_t, enc_t.data_ptr(), bytes_t, out_t.data_ptr()), raw_t, bytes_t),
        }
        for name, (fn, r, c) in specs.items():
            t = timeit(fn, nrep)
            kernels[name] = {"ms": round(t * 1e3, 4), "algorithmic_bytes": r + c, "achieved_gbs": round((r + c) / t / 1e9, 2),
                             "uncompressed_gbs": round(r / t / 1e9, 2)}

    # ---- e2e through the drop-in C API with pinned host buffers (rank-local, max over ranks) ----
    e2e = None
    if not args.no_e2e:
        L = C.CDLL(trico_b200.LIB_PATH)
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
        for fn in ("trico_write_vertices", "trico_write_triangles"):
            getattr(L, fn).restype = C.c_int
            getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        for fn in ("trico_read_vertices", "trico_read_triangles"):
            getattr(L, fn).restype = C.c_int
            getattr(L, fn).argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        os.environ["TRICO_B200_DEVICE"] = str(local)
        hv = verts.cpu().pin_memory()
        ht = tris.cpu().pin_memory()
        hov = torch.empty_like(hv).pin_memory()
        hot = torch.empty_like(ht).pin_memory()
        e2e_steps = max(2, min(args.steps, 4))
        te = td = 0.0
        arch_bytes = 0
        e2e_launches = 0
        for s in range(1 + e2e_steps):            # first pass = warm-up (context + buffer growth)
            if dist is not None:
                dist.barrier()
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
            assert ok
            t2 = time.perf_counter()
            if s == e2e_steps:
                e2e_launches = L.trico_b200_launch_count(a) + L.trico_b200_launch_count(r)
            L.trico_close_archive(r)
            L.trico_close_archive(a)
            if s > 0:
                te += t1 - t0; td += t2 - t1
        assert torch.equal(hov.view(torch.int32), hv.view(torch.int32)) and torch.equal(hot, ht)
        t2e = torch.tensor([te, td], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t2e, op=dist.ReduceOp.MAX)
        te, td = float(t2e[0]), float(t2e[1])
        e2e_raw = raw * world * e2e_steps
        e2e = {"value": round(2 * e2e_raw / (te + td) / 1e9, 3), "unit": "GB/s",
               "h2d_bytes_per_step": raw + arch_bytes, "d2h_bytes_per_step": arch_bytes + raw,
               "encode_gbs": round(e2e_raw / te / 1e9, 3), "decode_gbs": round(e2e_raw / td / 1e9, 3),
               "steps": e2e_steps, "host_memory": "pinned", "launches_per_step": int(e2e_launches)}
        del hv, ht, hov, hot

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": round(kernels[dom]["achieved_gbs"] / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                "all_kernels": {k: {"ms": v["ms"], "achieved_gbs": v["achieved_gbs"], "frac": round(v["achieved_gbs"] / peak, 4)} for k, v in kernels.items()}}

    # ---- CPU baseline: the unmodified reference on this box's cores, bounded sample ----
    cpu = None
    ratio_ref = None
    if not args.no_cpu:
        try:
            cores = os.cpu_count() or 1
            threads = max(1, min(cores, 64))
            hvs, hts = host_sample(2000, 1000)
            cpu_reference_run(hvs, hts, 1, threads)
            r = cpu_reference_run(hvs, hts, 2, threads)
            v, e, dd, ratio_ref = cpu_summary(r)
            cpu = {"value": round(v, 4), "unit": "GB/s", "cores": threads, "kind": r["kind"],
                   "encode_gbs": round(e, 4), "decode_gbs": round(dd, 4),
                   "sample": f"{threads} threads x 2 passes over a {hvs.shape[0]}-vertex / {hts.shape[0]}-triangle mesh from the C2 generator"}
        except Exception as ex:      # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round((t_enc + t_dec) / args.steps * 1e3, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"C2 synthetic float mesh: {nv} vertices + {nt} uint32 triangles per GPU ({W}x{H} jittered grid, ids shuffled in blocks of 64)",
                   "raw_bytes_per_gpu": raw, "fpc_chunk_values": 1 << l2v, "lz4_block_bytes": 1 << l2t, "fpc_exponents": [2, 4],
                   "lz4_hash_entries": 1024, "e2e_host_pipeline": "32 MiB slabs, H2D / kernels / D2H on three streams",
                   "l2": "inputs (3.6 GB) are larger than L2; no flush needed", "sharding": f"{world} x whole mesh, chunk-sharded, size exchange only",
                   "gpu": devname},
        "encode_gbs": round(enc_gbs, 2), "decode_gbs": round(dec_gbs, 2),
        "ratio": round(ratio, 4), "ratio_vertices": round(raw_v / bytes_v, 4), "ratio_triangles": round(raw_t / bytes_t, 4),
        "ratio_reference_sample": None if ratio_ref is None else round(ratio_ref, 4),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
