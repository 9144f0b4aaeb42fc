import ctypes as C, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from quokka_b200 import capi
from quokka_b200.capi import check, QK_HLLC, QK_MINMOD
from test_gpu_operators import make_cons, params, dev, one, VALID, NG
lib = capi.load()
def cmp(a, b, what):
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    print(f"{what}: mismatches {bad.sum()} of {bad.size}; nan gpu {np.isnan(a).sum()} nan ref {np.isnan(b).sum()}")
    if bad.any():
        idx = np.argwhere(bad)[:3]
        print("   ", idx.tolist(), a[bad][:3], b[bad][:3])
prm = params(0)
cons = make_cons(0, "shocked")
v = cons.view(VALID)
v[0, 5, 5, 5] = -0.3      # negative density
v[0, 10, 12, 20] = -2.0
v[4, 8, 8, 8] = -50.0     # negative energy
gb = VALID.grown(NG)
po = ol.HostFab(gb, 6)
ol.oracle().orc_conserved_to_primitive(one(prm), one(cons.desc()), one(po.desc()), one(gb))
dc, dp = dev(cons), dev(ol.HostFab(gb, 6))
check(lib.qk_hydro_conserved_to_primitive(one(prm), 1, one(VALID), one(dc.desc()), one(dp.desc()), NG, None))
cmp(dp.numpy(), po.a, "prim")
g1, g2 = VALID.grown(1), VALID.grown(2)
chis = []; dch = []
for d in range(3):
    c = ol.HostFab(g2, 1)
    ol.oracle().orc_flattening_coefficients(one(prm), d, one(po.desc()), one(c.desc()), one(g2))
    dcc = dev(ol.HostFab(g2, 1))
    check(lib.qk_hydro_flattening_coefficients(one(prm), d, 1, one(VALID), one(dp.desc()), one(dcc.desc()), 2, None))
    cmp(dcc.numpy(), c.a, f"chi{d}")
    chis.append(c); dch.append(dev(c))
dpo = dev(po)
for d in range(3):
    fb = ol.face_box(VALID, d, 1); fb0 = ol.face_box(VALID, d, 0)
    L, R = ol.HostFab(fb, 6), ol.HostFab(fb, 6)
    ol.oracle().orc_reconstruct_states(3, QK_MINMOD, d, one(po.desc()), one(L.desc()), one(R.desc()), one(g1), 6)
    ol.oracle().orc_flatten_shocks(d, one(po.desc()), one(chis[0].desc()), one(chis[1].desc()), one(chis[2].desc()), one(L.desc()), one(R.desc()), one(g1), 6)
    Fo, Vo = ol.HostFab(fb0, 6), ol.HostFab(fb0, 1)
    ol.oracle().orc_compute_fluxes(one(prm), QK_HLLC, d, one(Fo.desc()), one(Vo.desc()), one(L.desc()), one(R.desc()), one(po.desc()), one(fb0))
    dF, dV = dev(ol.HostFab(fb0, 6)), dev(ol.HostFab(fb0, 1))
    check(lib.qk_hydro_flux_function(one(prm), 0, d, 1, one(VALID), one(dpo.desc()), one(dch[0].desc()), one(dch[1].desc()), one(dch[2].desc()), one(dF.desc()), one(dV.desc()), None))
    cmp(dF.numpy(), Fo.a, f"flux{d} fused")
    cmp(dV.numpy(), Vo.a, f"fvel{d} fused")
    dl, dr = dev(L), dev(R)
    dF2, dV2 = dev(ol.HostFab(fb0, 6)), dev(ol.HostFab(fb0, 1))
    check(lib.qk_hydro_compute_fluxes(one(prm), QK_HLLC, d, 1, one(VALID), one(dF2.desc()), one(dV2.desc()), one(dl.desc()), one(dr.desc()), one(dpo.desc()), None))
    cmp(dF2.numpy(), Fo.a, f"flux{d} unfused")
