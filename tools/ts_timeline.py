#!/usr/bin/env python
"""Print the per-tile clock64 timeline written by a -DMOP_FWD_TIMELINE build (development aid)."""
import re, sys
L = [l.split() for l in open(sys.argv[1]) if re.match(r'(S-lane|PV-lane|softmax) t=', l)]
base = min(int(x) for l in L for x in l[3:] if int(x) > 0)
for name in ('softmax', 'S-lane', 'PV-lane'):
    for l in L:
        if l[0] == name and 8 <= int(l[2]) <= 13:
            print(l[0], l[2], ' '.join(str(int(x) - base) if int(x) > 0 else '-' for x in l[3:]))
