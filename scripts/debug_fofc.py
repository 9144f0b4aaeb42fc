import ctypes as C, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from quokka_b200 import capi
from quokka_b200.capi import check
from quokka_b200.device import DevMultiFab
from test_gpu_level import GenericProblem, oracle_level, oracle_state, level_desc
lib = capi.load()
p = GenericProblem((32, 32, 32), 16, (1, 1, 1), "periodic")
for dt in (1e-3, 2e-3, 4e-3):
    prm = p.params(); prm.abort_on_fofc_failure = 0; prm.integrator_order = 1
    st = p.states(seed=9, kind="shocked")
    L, keep = oracle_level(p, st)
    o = ol.oracle()
    bo1, bo2 = C.c_int64(), C.c_int64()
    o.orc_advance_hydro_level(L, C.byref(prm), dt, 1e9, C.byref(bo1), C.byref(bo2))
    desc, keep2 = level_desc(p)
    lev = C.c_void_p(); check(lib.qk_level_create(C.byref(desc), C.byref(lev)))
    U0 = DevMultiFab(p.boxes, p.ncomp, ngrow=4, host=st); U1 = DevMultiFab(p.boxes, p.ncomp, ngrow=4)
    b1 = C.c_int64(-1)
    check(lib.qk_fill_boundary(lev, U0.descs, 0, p.ncomp, None))
    check(lib.qk_hydro_advance_stage_faithful(lev, C.byref(prm), 1, U0.descs, U0.descs, U1.descs, dt, C.byref(b1), None))
    got = U1.numpy()
    tot = 0
    for b in range(len(p.boxes)):
        ref = oracle_state(p, L, 0, b)[:, 4:-4, 4:-4, 4:-4]; g = got[b][:, 4:-4, 4:-4, 4:-4]
        bad = ~((g == ref) | (np.isnan(g) & np.isnan(ref)))
        tot += bad.sum()
        if bad.any() and b == 0:
            idx = np.argwhere(bad)
            print("  first mismatches", idx[:5].tolist(), g[bad][:3], ref[bad][:3], "nan ref", np.isnan(ref).sum(), "nan got", np.isnan(g).sum())
    print(f"dt={dt}: oracle bad(first check)={bo1.value} gpu bad(after fofc)={b1.value} mismatches={tot}")
    lib.qk_level_destroy(lev); o.orc_level_destroy(L)
