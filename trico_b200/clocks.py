"""SM clock / throttle-reason sampling during a timed region (B200_PROFILING.md recipe), used by bench.py."""
from __future__ import annotations

import os
import subprocess
import threading
import time


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
