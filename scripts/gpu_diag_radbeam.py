#!/usr/bin/env python3
"""diagnostic: relaxed radiation sweeps vs the oracle on 'beam' fields (inadmissible states), error distribution after 1..3 substeps"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_radiation as t  # noqa: E402
from quokka_b200 import capi  # noqa: E402
from quokka_b200.capi import rad_params  # noqa: E402

lib = capi.load()
for case in ("plm_reflect", "ppm_single_ragged"):
    cfg = t.CASES[case]
    tr = t.TRAITS[cfg["traits"]]
    for nsub in (1, 2, 3):
        prm = rad_params(recon_order=cfg["order"], arith=capi.QK_ARITH_FAST, **tr)
        p = t.RadProblem(cfg["ncell"], cfg["grid"], cfg["periodic"], prm)
        st = p.states("beam")
        dt = 0.3 * min(p.dx) / prm.c_hat
        got = t.gpu_rad_steps(lib, p, st, dt, nsub)
        p.prm = rad_params(recon_order=cfg["order"], **tr)
        ex = t.gpu_rad_steps(lib, p, st, dt, nsub)
        want = t.oracle_rad_steps(p, st, dt, nsub)
        ng = p.nghost
        for name, ref in (("oracle", want), ("exact GPU", ex)):
            v = np.concatenate([g[prm.nstart:, ng:-ng, ng:-ng, ng:-ng].reshape(4, -1) for g in got], axis=1)
            r = np.concatenate([g[prm.nstart:, ng:-ng, ng:-ng, ng:-ng].reshape(4, -1) for g in ref], axis=1)
            rel = np.abs(v - r) / np.maximum(np.abs(r), 1e-300)
            scale = np.array([1.0, prm.c_light, prm.c_light, prm.c_light])[:, None] * np.abs(r[0]).max()
            abss = np.abs(v - r) / scale
            print(case, "substeps", nsub, "vs", name, "| pointwise rel: median %.2e p99 %.2e max %.2e | frac rel>1e-12: %.4f  frac abs/scale>1e-10: %.4f max abs/scale %.2e"
                  % (np.median(rel), np.quantile(rel, 0.99), rel.max(), (rel > 1e-12).mean(), (abss > 1e-10).mean(), abss.max()), flush=True)
