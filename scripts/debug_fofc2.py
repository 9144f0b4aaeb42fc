import ctypes as C, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from quokka_b200 import capi
from quokka_b200.capi import check
from test_gpu_level import GenericProblem, oracle_level, oracle_state, run_stage_pair
lib = capi.load()
p = GenericProblem((32, 32, 32), 16, (1, 1, 1), "periodic")
for dt in (1e-3, 1.5e-3, 2e-3, 4e-3):
    prm = p.params(); prm.abort_on_fofc_failure = 0
    st = p.states(seed=9, kind="shocked")
    L, keep = oracle_level(p, st)
    o = ol.oracle()
    bo1, bo2 = C.c_int64(), C.c_int64()
    o.orc_advance_hydro_level(L, C.byref(prm), dt, 1e9, C.byref(bo1), C.byref(bo2))
    got, b1, b2 = run_stage_pair(lib, p, prm, st, dt, lib.qk_hydro_advance_stage_faithful)
    tot = 0; per = np.zeros(6, int)
    for b in range(len(p.boxes)):
        ref = oracle_state(p, L, 0, b)[:, 4:-4, 4:-4, 4:-4]; g = got[b][:, 4:-4, 4:-4, 4:-4]
        bad = ~((g == ref) | (np.isnan(g) & np.isnan(ref)))
        tot += bad.sum(); per += bad.reshape(6, -1).sum(1)
        if bad.any() and b == 0:
            idx = np.argwhere(bad)
            print("  first", idx[:4].tolist(), g[bad][:3], ref[bad][:3], "nan ref", np.isnan(ref).sum(), "nan got", np.isnan(g).sum(), "neg rho ref", (ref[0] <= 0).sum())
    print(f"dt={dt}: oracle bad1={bo1.value} bad2={bo2.value}; gpu after-fofc {b1} {b2}; mismatches={tot} per comp {per.tolist()}")
    o.orc_level_destroy(L)
