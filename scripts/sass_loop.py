#!/usr/bin/env python3
"""Static SASS opcode mix of a kernel's main loop (largest backward-branch span), the offline proxy for the issue-bound sweeps:
    python scripts/sass_loop.py file.o <mangled-name-substring> [...]     (or a .sass dump)
Prints total instructions, the FP64-pipe share and the largest non-FP64 classes of the loop body."""
import collections
import re
import subprocess
import sys


def parse(text):
    ins = []
    for ln in text.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    return ins


def loop_span(ins):
    best = None
    for a, op, args in ins:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", args)
            if m:
                t = int(m.group(1), 16)
                if t < a and (best is None or a - t > best[1] - best[0]):
                    best = (t, a)
    return best


FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU")


def report(name, ins):
    span = loop_span(ins)
    if not span:
        print(name, "no loop")
        return
    body = [(a, op, args) for a, op, args in ins if span[0] <= a <= span[1]]
    c = collections.Counter(op.split(".")[0] for _, op, _ in body)
    full = collections.Counter(op for _, op, _ in body)
    tot = len(body)
    f64 = sum(c[x] for x in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"{name[:120]}\n  loop 0x{span[0]:x}..0x{span[1]:x}: {tot} instr, FP64 pipe {f64} ({100*f64/tot:.1f} %)  [DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']} DSETP {c['DSETP']}] MUFU {c['MUFU']}")
    print("  top:", ", ".join(f"{k} {v}" for k, v in c.most_common(16)))
    print("  IMAD.MOV*", sum(v for k, v in full.items() if k.startswith("IMAD.MOV")), " MOV", c["MOV"], " IMAD.WIDE*", sum(v for k, v in full.items() if k.startswith("IMAD.WIDE")),
          " BRA", c["BRA"], " BSSY/BSYNC", c["BSSY"] + c["BSYNC"], " LDS", c["LDS"], " STG", c["STG"], " LDG", c["LDG"], " SHFL", c["SHFL"])


def main():
    src = sys.argv[1]
    if src.endswith(".sass"):
        report(src, parse(open(src).read()))
        return
    out = subprocess.run(["cuobjdump", "-sass", src], capture_output=True, text=True).stdout
    chunks = re.split(r"\n\s+Function : ", out)
    for ch in chunks[1:]:
        name = ch.split("\n", 1)[0].strip()
        if all(p in name for p in sys.argv[2:]):
            report(name, parse(ch))


if __name__ == "__main__":
    main()
