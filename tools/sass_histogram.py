#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libmop_b200.so: the mnemonics that prove tcgen05 / TMA / packed-math use.

usage: python tools/sass_histogram.py [path/to/libmop_b200.so] > profiles/rNN_sass_opcodes.txt
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG = cp.async.bulk.tensor (TMA tile load),
UBLKCP = cp.async.bulk (non-tensor bulk copy), FFMA2 / FMUL2 / FADD2 = packed fp32 pairs, MUFU = transcendental unit.
"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "mop_b200", "libmop_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "FFMA2", "FMUL2", "FADD2", "MUFU", "FFMA", "SHFL", "BAR", "SYNCS"]
cur, counts, total = None, {}, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for w in WATCH:
            if op == w or (w in ("MUFU", "BAR", "SYNCS", "SHFL") and op.startswith(w)):
                counts[cur][w] += 1
print(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{w:>7s}" for w in WATCH))
for k in sorted(counts, key=lambda k: -total[k]):
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
    name = re.sub(r"\(.*", "", name)[:70]
    print(f"{name:70s} {total[k]:7d} " + " ".join(f"{counts[k][w]:7d}" for w in WATCH))
