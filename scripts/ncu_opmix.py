#!/usr/bin/env python3
"""dynamic opcode mix + hottest stall lines from `ncu --page source --csv`: python scripts/ncu_opmix.py src.csv [nwarps]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
# the source page concatenates kernels: keep the first one ("Kernel Name" row, header row, instructions...)
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = rows[starts[which]:(starts[which + 1] if which + 1 < len(starts) else len(rows))]
print(rows[0][1][:120])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samples = collections.Counter(); tot = 0; stot = 0
lines = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix['Source']].strip()
    m = re.match(r'(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    if not m: continue
    op = m.group(1).split('.')[0]
    n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    ops[op] += n; samples[op] += s; tot += n; stot += s
    lines.append((s, n, src, r[ix['stall_long_sb']], r[ix['stall_wait']], r[ix['stall_math']] if 'stall_math' in ix else ''))
nw = float(sys.argv[2]) if len(sys.argv) > 2 else None
print(f"total warp instr {tot}  samples {stot}" + (f"  per warp {tot/nw:.0f}" if nw else ""))
fp64 = sum(v for k, v in ops.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
print(f"fp64 pipe {fp64} = {100*fp64/tot:.1f}%" + (f"  per warp {fp64/nw:.0f}" if nw else ""))
for k, v in ops.most_common(28):
    print(f"  {k:10s} {v:12d} {100*v/tot:5.1f}%  samples {100*samples[k]/max(stot,1):5.1f}%" + (f"  per warp {v/nw:7.1f}" if nw else ""))
print("hottest instructions by stall samples:")
for s, n, src, lsb, w, mth in sorted(lines, reverse=True)[:25]:
    print(f"  {s:6d} ({100*s/max(stot,1):4.1f}%) exec {n:9d}  long_sb {lsb:>5} wait {w:>5} math {mth:>5}  {src[:90]}")
