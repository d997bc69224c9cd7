#!/usr/bin/env python
"""Print kernel name and duration (us) per launch from `ncu --csv --metrics gpu__time_duration.sum` output on stdin."""
import csv, sys
rows = [r for r in csv.reader(sys.stdin) if r]
hdr = next((r for r in rows if "Kernel Name" in r), None)
if not hdr:
    sys.exit("no ncu csv header found")
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rows[rows.index(hdr) + 1:]:
    if len(r) > iv:
        v = float(r[iv].replace(",", ""))
        if r[iu] in ("ns", "nsecond"): v /= 1e3
        elif r[iu] in ("ms", "msecond"): v *= 1e3
        print(f"{v:10.1f} us  {r[ik][:90]}")
