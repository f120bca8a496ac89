#!/usr/bin/env python
"""Per-source-line hot spots from an ncu report, without the GUI.

    tools/ncu_lines.py report.ncu-rep kernel_substring [--top 25] [--lib trico_b200/lib/libtrico_b200.so]

Joins the SASS page of the report (stall samples + executed instructions per instruction) with
nvdisasm's line info of the same cubin (the k-th instruction of the kernel in both listings),
and prints the source lines ranked by stall samples.  The library must be the build that was profiled.
"""
import argparse, csv, os, re, subprocess, sys, tempfile, collections

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("kernel")
ap.add_argument("--top", type=int, default=25)
ap.add_argument("--lib", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "trico_b200", "lib", "libtrico_b200.so"))
ap.add_argument("--sass", action="store_true", help="also print the hottest SASS instructions")
args = ap.parse_args()

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(args.lib)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("device_api")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()

# nvdisasm: sections per kernel
sections, cur, name, line = {}, None, None, None
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        name = m.group(1); cur = []; sections[name] = cur; line = ("?", 0); continue
    if l.startswith("//---") or l.strip().startswith(".section"):
        if not l.strip().startswith(".section\t.text"):
            cur = None if l.startswith("//---") and ".text." not in l else cur
        continue
    if cur is None:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        cur.append((int(m.group(1), 16), m.group(2).strip(), line))

out = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
done = set()
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        kname = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if rows[j]:
                body.append(rows[j])
            j += 1
        i = j
        if args.kernel not in kname or kname in done:
            continue
        done.add(kname)
        # find the matching nvdisasm section by instruction count + mangled-name hint
        import re as _re
        base = _re.sub(r"^void ", "", kname).split("<")[0].split("(")[0].split("::")[-1]
        cands = [(n, s) for n, s in sections.items() if len(s) == len(body) and base in n]
        if len(cands) > 1:
            # several template instances of the same size: keep the one whose opcodes line up
            def score(sec):
                return sum(1 for r, (a, ins, ln) in zip(body, sec) if r[1].split()[0:1] == ins.split()[0:1] or ins.split()[0] in r[1])
            cands.sort(key=lambda c: -score(c[1]))
        if not cands:
            print(f"## {kname}: no disassembly with {len(body)} instructions (library differs from the profiled build?)"); continue
        sec = cands[0][1]
        cS, cI = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [(k, h) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
        tot_s = tot_i = 0
        hot = []
        for r, (addr, ins, ln) in zip(body, sec):
            s, n = int(r[cS] or 0), int(r[cI] or 0)
            per[ln][0] += s; per[ln][1] += n
            for k, h in stall_cols:
                v = int(r[k] or 0)
                if v: per[ln][2][h[6:]] += v
            tot_s += s; tot_i += n
            hot.append((s, n, addr, ins, ln))
        print(f"## {kname}\n   {len(body)} SASS instructions, {tot_i} warp-instructions executed, {tot_s} stall samples ({cands[0][0]})")
        print(f"{'file:line':28s} {'samples':>8s} {'%':>6s} {'inst':>11s} {'%':>6s}  top stalls")
        for ln, (s, n, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:args.top]:
            tops = ", ".join(f"{a}:{b}" for a, b in st.most_common(3))
            print(f"{ln[0] + ':' + str(ln[1]):28s} {s:8d} {100.0 * s / max(tot_s, 1):6.1f} {n:11d} {100.0 * n / max(tot_i, 1):6.1f}  {tops}")
        if args.sass:
            print("-- hottest instructions")
            for s, n, addr, ins, ln in sorted(hot, key=lambda t: -t[0])[:args.top]:
                print(f"  {addr:06x} {ln[0]}:{ln[1]:<5d} {s:7d} {n:10d}  {ins}")
    else:
        i += 1
