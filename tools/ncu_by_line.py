#!/usr/bin/env python
"""Aggregate an ncu SASS-level source page by CUDA source line.

usage: ncu_by_line.py <report.ncu-rep> <kernel-mangled-substring> [--so mop_b200/libmop_b200.so] [--top 40]
The ncu CSV source page lists SASS instructions in address order; nvdisasm -g on the cubin embedded in the .so
gives the same instruction sequence annotated with `//## File "...", line N`.  The two are zipped by order.
"""
import argparse, csv, os, re, subprocess, sys, tempfile

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("kernel")
ap.add_argument("--so", default=os.path.join(os.path.dirname(__file__), "..", "mop_b200", "libmop_b200.so"))
ap.add_argument("--top", type=int, default=45)
ap.add_argument("--name", default=None, help="regex for ncu's --kernel-name filter (default: the function name inside `kernel`)")
a = ap.parse_args()
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.so)], cwd=tmp, capture_output=True)
# one cubin per translation unit: take the one that holds the kernel
sass, start = None, None
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    cand = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    hit = [i for i, l in enumerate(cand) if l.startswith(".text.") and a.kernel in l]
    if hit:
        sass, start = cand, hit[0]
        break
if sass is None:
    sys.exit(f"kernel {a.kernel!r} not found in {a.so}")
lines = []  # (file, line) per instruction
cur = ("?", 0)
for l in sass[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
# a report may hold several kernels: keep the launches whose (demangled) name contains the function name of `kernel`
fname = re.sub(r"^_ZN?\d*", "", a.kernel)
fname = re.search(r"[A-Za-z_][A-Za-z0-9_]*kernel[A-Za-z0-9_]*", a.kernel)
flt = ["--kernel-name", "regex:" + (a.name or fname.group(0))] if (a.name or fname) else []
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", *flt], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
ins = [r for r in rows[2:] if len(r) >= len(hdr)]
if len(ins) != len(lines):
    print(f"warning: {len(ins)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
agg = {}
ti = ts = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r, key in zip(ins, lines):
    n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    d = agg.setdefault(key, [0, 0, {}])
    d[0] += n; d[1] += s
    for h in stall_cols:
        v = int(r[ix[h]] or 0)
        if v: d[2][h] = d[2].get(h, 0) + v
    ti += n; ts += s
print(f"total warp-instructions {ti}, samples {ts}")
srcs = {}
for (f, ln), (n, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
    if f not in srcs:
        try: srcs[f] = open(os.path.join(os.path.dirname(__file__), "..", "mop_b200", "csrc", f)).read().splitlines()
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    top = ",".join(f"{k[6:]}:{100*v//max(s,1)}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{f}:{ln:<4d} inst {100*n/ti:5.1f}%  samp {100*s/ts:5.1f}%  [{top}] | {text}")
