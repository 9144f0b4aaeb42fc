#!/usr/bin/env python3
"""static SASS instruction mix per kernel: python scripts/sass_mix.py <file.sass> <name filter>..."""
import re, collections, subprocess, sys
txt = open(sys.argv[1]).read()
filters = sys.argv[2:]
funcs = re.split(r'\n\s*Function : ', txt)[1:]
for f in funcs:
    name = f.split('\n')[0]
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r'\(anonymous namespace\)::', '', dem).split('(')[0]
    if filters and not any(k in dem for k in filters):
        continue
    ops = collections.Counter()
    for m in re.finditer(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', f, re.M):
        ops[m.group(1).split('.')[0]] += 1
    tot = sum(ops.values())
    fp64 = sum(v for k, v in ops.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX'))
    print(dem)
    print('  total', tot, 'fp64pipe', fp64, dict(ops.most_common(16)))
