#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list.

    tools/launch_list_summary.py launches.csv [--traffic profiles/traffic.json] [--md out.md]

Per kernel: launches, mean duration, share of the listed GPU time, mean DRAM bytes per launch.
The times are cold-cache and serialised (ncu replays every launch): compare SHARES, not absolutes.
"""
import argparse, collections, csv, json, re

ap = argparse.ArgumentParser()
ap.add_argument("csv"); ap.add_argument("--traffic"); ap.add_argument("--md")
a = ap.parse_args()
rows = []
with open(a.csv, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
per = collections.OrderedDict()
for r in rd:
    k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("tb200::", "")
    k = k.replace("unsigned int", "u32").replace("unsigned long", "u64").replace("(int)", "").replace(" ", "")
    d = per.setdefault(k, {"n": set(), "t": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    d["n"].add(r["ID"])
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["t"] += v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1.0)
    else:
        b = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d["rd" if "read" in r["Metric Name"] else "wr"] += b
tot = sum(d["t"] for d in per.values())
out = ["| kernel | launches | mean us | share of listed GPU time | DRAM read MB / launch | DRAM write MB / launch |", "|---|---|---|---|---|---|"]
traffic = {}
for k, d in per.items():
    n = len(d["n"])
    out.append(f"| `{k}` | {n} | {d['t'] / n:.1f} | {100 * d['t'] / tot:.1f} % | {d['rd'] / n / 1e6:.1f} | {d['wr'] / n / 1e6:.1f} |")
    traffic[k] = int((d["rd"] + d["wr"]) / n)
txt = "\n".join(out)
print(txt)
if a.md:
    open(a.md, "w").write(txt + "\n")
if a.traffic:
    # keys = the kernel labels bench.py uses (the LZ4 encoder is two launches: compress + assemble)
    def find(sub):
        return sum(v for k, v in traffic.items() if sub in k)
    out_t = {
        "fpc_encode_lanes_kernel<u32,3,1,32>": find("fpc_encode_lanes_kernel"),
        "lz4_encode_kernel<4,10>+lz4_assemble_kernel": find("lz4_encode_kernel") + find("lz4_assemble_kernel"),
        "fpc_decode_kernel<u32,3,1,32>": find("fpc_decode_kernel"),
        "lz4_decode_kernel<4>": find("lz4_decode_kernel"),
        "_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, bench.py C2 full size (" + a.csv + ")",
    }
    json.dump(out_t, open(a.traffic, "w"), indent=1)
