import ctypes as C, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from quokka_b200 import capi
from quokka_b200.capi import check, qk_box
from quokka_b200.device import DevMultiFab
from test_gpu_level import GenericProblem, oracle_level, oracle_state, level_desc
lib = capi.load()
p = GenericProblem((32, 32, 32), 32, (1, 1, 1), "periodic")   # ONE box: no redo exchange except periodic self
dt = 2e-3
prm = p.params(); prm.abort_on_fofc_failure = 0
st = p.states(seed=9, kind="shocked")
L, keep = oracle_level(p, st)
o = ol.oracle()
bo1, bo2 = C.c_int64(), C.c_int64()
o.orc_advance_hydro_level(L, C.byref(prm), dt, 1e9, C.byref(bo1), C.byref(bo2))
ref = oracle_state(p, L, 0, 0)[:, 4:-4, 4:-4, 4:-4]
desc, keep2 = level_desc(p)
lev = C.c_void_p(); check(lib.qk_level_create(C.byref(desc), C.byref(lev)))
B = p.boxes; nb = len(B); nv = 6
U0 = DevMultiFab(B, 6, ngrow=4, host=st); U1 = DevMultiFab(B, 6, ngrow=4); U2 = DevMultiFab(B, 6, ngrow=4)
b1, b2 = C.c_int64(-1), C.c_int64(-1)
check(lib.qk_fill_boundary(lev, U0.descs, 0, 6, None))
check(lib.qk_hydro_advance_stage_faithful(lev, C.byref(prm), 1, U0.descs, U0.descs, U1.descs, dt, C.byref(b1), None))
check(lib.qk_fill_boundary(lev, U1.descs, 0, 6, None))
check(lib.qk_hydro_advance_stage_faithful(lev, C.byref(prm), 2, U0.descs, U1.descs, U2.descs, dt, C.byref(b2), None))
got = U2.numpy()[0][:, 4:-4, 4:-4, 4:-4]
bad = ~((got == ref) | (np.isnan(got) & np.isnan(ref)))
print("one-box periodic: oracle bad", bo1.value, bo2.value, "gpu", b1.value, b2.value, "mismatches", bad.sum())
# ---- python chain with GPU per-op calls, stage 2 from the same U0, U1 ----
vb = (qk_box * nb)(*B)
dx = (C.c_double * 3)(*p.dx)
def mf(nc, ng=0, fd=None, dtype="f64"): return DevMultiFab(B, nc, ngrow=ng, face_dir=fd, dtype=dtype)
def fluxes(U):
    prim = mf(6, 4); chi = [mf(1, 2) for _ in range(3)]
    check(lib.qk_hydro_conserved_to_primitive(C.byref(prm), nb, vb, U.descs, prim.descs, 4, None))
    for d in range(3): check(lib.qk_hydro_flattening_coefficients(C.byref(prm), d, nb, vb, prim.descs, chi[d].descs, 2, None))
    F = [mf(6, 0, d) for d in range(3)]; V = [mf(1, 0, d) for d in range(3)]
    for d in range(3): check(lib.qk_hydro_flux_function(C.byref(prm), 0, d, nb, vb, prim.descs, chi[0].descs, chi[1].descs, chi[2].descs, F[d].descs, V[d].descs, None))
    return F, V, prim
def fo_fluxes(U):
    prim = mf(6, 4)
    check(lib.qk_hydro_conserved_to_primitive(C.byref(prm), nb, vb, U.descs, prim.descs, 4, None))
    F = [mf(6, 0, d) for d in range(3)]; V = [mf(1, 0, d) for d in range(3)]
    for d in range(3): check(lib.qk_hydro_flux_function(C.byref(prm), 1, d, nb, vb, prim.descs, None, None, None, F[d].descs, V[d].descs, None))
    return F, V
def facebox(d): return (qk_box * nb)(*[b.grown(0, d) for b in B])
def update(F, V, Uout, redo):
    rhs = mf(6)
    check(lib.qk_hydro_rhs_from_fluxes(nb, vb, rhs.descs, F[0].descs, F[1].descs, F[2].descs, dx, 6, None))
    check(lib.qk_hydro_add_internal_energy_pdv(C.byref(prm), nb, vb, rhs.descs, U0.descs, dx, V[0].descs, V[1].descs, V[2].descs, redo.descs, None))
    n = C.c_int64()
    check(lib.qk_hydro_predict_step(C.byref(prm), nb, vb, U0.descs, Uout.descs, rhs.descs, dt, 6, redo.descs, C.byref(n), None))
    return n.value
FO, FOV = fo_fluxes(U0)
F0, V0, _ = fluxes(U0)
frk = [mf(6, 0, d) for d in range(3)]; avg = [mf(1, 0, d) for d in range(3)]
for d in range(3):
    check(lib.qk_saxpy(nb, facebox(d), frk[d].descs, 0.5, F0[d].descs, 6, None)); check(lib.qk_saxpy(nb, facebox(d), avg[d].descs, 0.5, V0[d].descs, 1, None))
F1, V1, _ = fluxes(U1)
for d in range(3):
    check(lib.qk_saxpy(nb, facebox(d), frk[d].descs, 0.5, F1[d].descs, 6, None)); check(lib.qk_saxpy(nb, facebox(d), avg[d].descs, 0.5, V1[d].descs, 1, None))
redo = mf(1, 1, None, "i32")
U2c = DevMultiFab(B, 6, ngrow=4)
n1 = update(frk, avg, U2c, redo)
print("chain stage-2 first-check bad", n1)
# periodic self exchange of the flags (one box): do it on the host
r = redo.fabs[0].t
r[:, 0, :, :] = r[:, -2, :, :]; r[:, -1, :, :] = r[:, 1, :, :]
r[:, :, 0, :] = r[:, :, -2, :]; r[:, :, -1, :] = r[:, :, 1, :]
r[:, :, :, 0] = r[:, :, :, -2]; r[:, :, :, -1] = r[:, :, :, 1]
for d in range(3):
    check(lib.qk_hydro_replace_fluxes(d, nb, vb, frk[d].descs, FO[d].descs, redo.descs, 6, None))
    check(lib.qk_hydro_replace_fluxes(d, nb, vb, avg[d].descs, FOV[d].descs, redo.descs, 1, None))
n2 = update(frk, avg, U2c, redo)
check(lib.qk_hydro_enforce_limits(C.byref(prm), nb, vb, U2c.descs, None)); check(lib.qk_hydro_sync_dual_energy(C.byref(prm), nb, vb, U2c.descs, None, None))
chain = U2c.numpy()[0][:, 4:-4, 4:-4, 4:-4]
bad_c = ~((chain == ref) | (np.isnan(chain) & np.isnan(ref)))
bad_f = ~((chain == got) | (np.isnan(chain) & np.isnan(got)))
print("chain after-fofc bad", n2, "chain vs oracle mismatches", bad_c.sum(), "chain vs faithful mismatches", bad_f.sum())
