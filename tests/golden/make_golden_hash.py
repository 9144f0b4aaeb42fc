#!/usr/bin/env python3
"""Generate tests/golden/sedov_hashes.json: SHA-256 digests (and per-component sums) of the state the UNMODIFIED reference
(oracle/_ref/test_hydro3d_blast) holds after 100 coarse steps of the Sedov blast at 32^3, 64^3 and 128^3 -- the "after 100 steps" point
of BASELINE.json's tolerance, at sizes whose full dumps (12.6 MB, 101 MB) are too large to commit.  Zeros are canonicalised
(-0.0 -> +0.0) before hashing; everything else is compared bit for bit through the digest.

Run in the build container only (needs oracle/_ref):   python tests/golden/make_golden_hash.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import run_reference  # noqa: E402


def digest(state):
    a = np.ascontiguousarray(state, dtype="<f8") + 0.0  # -0.0 + 0.0 = +0.0
    return hashlib.sha256(a.tobytes()).hexdigest()


def main():
    # `--full` adds the benchmarked size itself, Sedov 256^3 in 128^3 boxes x 100 steps (configs[1]; 36 minutes of the reference on 8 cores)
    cases = [(32, 16, 100), (64, 32, 100), (128, 64, 100)] + ([(256, 128, 100)] if "--full" in sys.argv else [])
    out = {}
    path = os.path.join(HERE, "sedov_hashes.json")
    if os.path.exists(path):  # keep entries this invocation does not regenerate
        with open(path) as f:
            out = json.load(f)
    for ncell, box, nsteps in cases:
        state, time, dts, retries = run_reference(ncell, box, nsteps, threads=os.cpu_count() or 8)
        out[f"sedov{ncell}_b{box}_s{nsteps}"] = {"ncell": ncell, "box": box, "nsteps": nsteps, "time": repr(float(time)), "retries": retries,
                                                   "sha256": digest(state), "sums": [repr(float(state[c].sum())) for c in range(6)],
                                                   "shape": list(state.shape)}
        print(ncell, time, retries, out[f"sedov{ncell}_b{box}_s{nsteps}"]["sha256"])
    with open(os.path.join(HERE, "sedov_hashes.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
