#!/usr/bin/env python3
"""profiles/ncu_traffic.json: measured DRAM traffic per launch of the sweep kernels (dram__bytes_read.sum + dram__bytes_write.sum from one
`ncu --set full` capture, 256^3 per GPU), keyed by a SHA-256 of the kernel sources so that bench.py reports `roofline.traffic` only while the
kernels are the ones that were profiled (it prints null + "stale" otherwise).
    python scripts/make_ncu_traffic.py <mode: relaxed|exact> <summary.csv made by scripts/ncu_summary.py> [<mode> <summary.csv> ...]"""
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ["qk_sweep_kernels.cuh", "qk_march.cuh", "qk_fast.cuh", "qk_relaxed.cuh", "qk_tma.cuh", "qk_div.cuh", "qk_physics.cuh"]


def sources_sha():
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "quokka_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def parse(summary):
    """launches in order: x, y, z (stage 1), x, y, z (stage 2) -> average bytes per class"""
    acc = {}
    name = None
    nxt = "sweep_y"
    rd = 0.0
    for ln in open(summary):
        if ln.startswith("Kernel Name"):
            name = ln
        elif ln.startswith("dram__bytes_read.sum"):
            rd = float(ln.split(",")[1]) * (1e9 if "Gbyte" in ln else 1e6 if "Mbyte" in ln else 1.0)
        elif ln.startswith("dram__bytes_write.sum"):
            wr = float(ln.split(",")[1]) * (1e9 if "Gbyte" in ln else 1e6 if "Mbyte" in ln else 1.0)
            # a stage launches x, then the y march, then the z march
            if "k_sweep_x" in name:
                cls, nxt = "sweep_x", "sweep_y"
            elif "k_march_t" in name:
                cls, nxt = nxt, ("sweep_z" if nxt == "sweep_y" else "sweep_y")
            else:
                cls = None
            if cls:
                acc.setdefault(cls, []).append(rd + wr)
    return {k: sum(v) / len(v) for k, v in acc.items()}


if __name__ == "__main__":
    out = {"sources_sha256": sources_sha(), "sources": KERNEL_SOURCES, "cells_per_launch": 256 ** 3}
    for mode, path in zip(sys.argv[1::2], sys.argv[2::2]):
        out[mode] = {"bytes_per_launch": parse(path), "from": os.path.relpath(path, ROOT)}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
