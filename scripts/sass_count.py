#!/usr/bin/env python3
"""static SASS instruction counts per kernel: python scripts/sass_count.py file.o [name-substring]"""
import re, subprocess, sys, collections
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
pat = sys.argv[2] if len(sys.argv) > 2 else ""
name = None; cnt = collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1); cnt[name] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and name:
        cnt[name][m.group(1)] += 1
for k, c in cnt.items():
    if pat in k:
        tot = sum(c.values()); f64 = sum(c[x] for x in ("DFMA", "DMUL", "DADD", "DSETP"))
        print(f"{k[:110]}\n   total {tot} fp64 {f64} (DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']} DSETP {c['DSETP']}) MUFU {c['MUFU']} FSEL {c['FSEL']} IMAD {c['IMAD']} BRA {c['BRA']}")
