import ctypes as C, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from quokka_b200 import capi
from quokka_b200.capi import qk_box, QK_HLLC, QK_LLF, QK_MINMOD, check
from quokka_b200.device import DevFab
from test_gpu_level import GenericProblem
lib = capi.load()
o = ol.oracle()
p = GenericProblem((32, 32, 32), 32, (1, 1, 1), "periodic")
dt = 2e-3
prm = p.params(); prm.abort_on_fofc_failure = 0
st = p.states(seed=9, kind="shocked")
one = C.byref
vb = p.boxes[0]; g4 = vb.grown(4); g2 = vb.grown(2); g1 = vb.grown(1)
dx = (C.c_double * 3)(*p.dx)
def dev(hf): return DevFab(hf.box, hf.ncomp, dtype="f64" if hf.a.dtype == np.float64 else "i32", host=hf.a)
def cmp(g, r, what):
    bad = ~((g == r) | (np.isnan(g) & np.isnan(r)))
    print(f"{what}: mismatches {bad.sum()} nan(gpu,ref)=({np.isnan(g).sum()},{np.isnan(r).sum()}) inf=({np.isinf(g).sum()},{np.isinf(r).sum()})")
    if bad.any():
        idx = np.argwhere(bad)[:4]; print("    at", idx.tolist(), g[bad][:4], r[bad][:4])
    return bad
def fill(U):
    a = U.a; n = 32
    for ax in (1, 2, 3):
        sl = [slice(None)] * 4
        lo = sl.copy(); lo[ax] = slice(0, 4); src = sl.copy(); src[ax] = slice(n, n + 4); a[tuple(lo)] = a[tuple(src)]
        hi = sl.copy(); hi[ax] = slice(n + 4, n + 8); src = sl.copy(); src[ax] = slice(4, 8); a[tuple(hi)] = a[tuple(src)]
def fluxes(U, fo=False, tag=""):
    prim = ol.HostFab(g4, 6)
    o.orc_conserved_to_primitive(one(prm), one(U.desc()), one(prim.desc()), one(g4))
    dU = dev(U); dprim = dev(ol.HostFab(g4, 6))
    check(lib.qk_hydro_conserved_to_primitive(one(prm), 1, one(vb), one(dU.desc()), one(dprim.desc()), 4, None))
    cmp(dprim.numpy(), prim.a, tag + " prim")
    dprim = dev(prim)
    chi = [ol.HostFab(g2, 1) for _ in range(3)]; dchi = []
    for d in range(3):
        o.orc_flattening_coefficients(one(prm), d, one(prim.desc()), one(chi[d].desc()), one(g2))
        dc = dev(ol.HostFab(g2, 1))
        check(lib.qk_hydro_flattening_coefficients(one(prm), d, 1, one(vb), one(dprim.desc()), one(dc.desc()), 2, None))
        cmp(dc.numpy(), chi[d].a, tag + f" chi{d}")
        dchi.append(dev(chi[d]))
    F = []; V = []
    for d in range(3):
        fb = ol.face_box(vb, d, 1); fb0 = ol.face_box(vb, d, 0)
        Ls, Rs = ol.HostFab(fb, 6), ol.HostFab(fb, 6)
        o.orc_reconstruct_states(1 if fo else 3, QK_MINMOD, d, one(prim.desc()), one(Ls.desc()), one(Rs.desc()), one(g1), 6)
        if not fo:
            o.orc_flatten_shocks(d, one(prim.desc()), one(chi[0].desc()), one(chi[1].desc()), one(chi[2].desc()), one(Ls.desc()), one(Rs.desc()), one(g1), 6)
        f, v = ol.HostFab(fb0, 6), ol.HostFab(fb0, 1)
        o.orc_compute_fluxes(one(prm), QK_LLF if fo else QK_HLLC, d, one(f.desc()), one(v.desc()), one(Ls.desc()), one(Rs.desc()), one(prim.desc()), one(fb0))
        dF, dV = dev(ol.HostFab(fb0, 6)), dev(ol.HostFab(fb0, 1))
        check(lib.qk_hydro_flux_function(one(prm), 1 if fo else 0, d, 1, one(vb), one(dprim.desc()), one(dchi[0].desc()), one(dchi[1].desc()), one(dchi[2].desc()), one(dF.desc()), one(dV.desc()), None))
        b = cmp(dF.numpy(), f.a, tag + f" flux{d}")
        cmp(dV.numpy(), v.a, tag + f" fvel{d}")
        if b.any():
            i = np.argwhere(b)[0]; n_, k_, j_, i_ = i
            print("     L", Ls.a[:, k_+1 if d != 2 else k_+1, j_+1, i_+1] if False else "", "prim around:", prim.a[:, k_+4, j_+4, i_+2:i_+7].tolist() if d == 0 else "")
        F.append(f); V.append(v)
    return F, V
def update(F, V, U0, Uout, redo, tag):
    rhs = ol.HostFab(vb, 6)
    o.orc_rhs_from_fluxes(one(rhs.desc()), one(F[0].desc()), one(F[1].desc()), one(F[2].desc()), dx, one(vb), 6)
    dF = [dev(f) for f in F]; dV = [dev(v) for v in V]; drhs = dev(ol.HostFab(vb, 6)); dU0 = dev(U0); dredo = dev(redo)
    check(lib.qk_hydro_rhs_from_fluxes(1, one(vb), one(drhs.desc()), one(dF[0].desc()), one(dF[1].desc()), one(dF[2].desc()), dx, 6, None))
    cmp(drhs.numpy(), rhs.a, tag + " rhs")
    o.orc_add_internal_energy_pdv(one(prm), one(rhs.desc()), one(U0.desc()), dx, one(V[0].desc()), one(V[1].desc()), one(V[2].desc()), one(redo.desc()), one(vb))
    drhs = dev(ol.HostFab(vb, 6)); drhs.t.copy_(__import__("torch").from_numpy(rhs.a * 0)); 
    # redo the GPU pdv from the oracle's pre-pdv rhs
    rhs0 = ol.HostFab(vb, 6)
    o.orc_rhs_from_fluxes(one(rhs0.desc()), one(F[0].desc()), one(F[1].desc()), one(F[2].desc()), dx, one(vb), 6)
    drhs = dev(rhs0)
    check(lib.qk_hydro_add_internal_energy_pdv(one(prm), 1, one(vb), one(drhs.desc()), one(dU0.desc()), dx, one(dV[0].desc()), one(dV[1].desc()), one(dV[2].desc()), one(dredo.desc()), None))
    cmp(drhs.numpy(), rhs.a, tag + " rhs+pdv")
    dUout = dev(ol.HostFab(g4, 6)); n = C.c_int64()
    check(lib.qk_hydro_predict_step(one(prm), 1, one(vb), one(dU0.desc()), one(dUout.desc()), one(drhs.desc()), dt, 6, one(dredo.desc()), C.byref(n), None))
    nb_ = o.orc_predict_step(one(prm), one(U0.desc()), one(Uout.desc()), one(rhs.desc()), dt, 6, one(redo.desc()), one(vb))
    cmp(dUout.numpy()[:, 4:-4, 4:-4, 4:-4], Uout.view(vb), tag + f" predict (bad gpu {n.value} ref {nb_})")
    cmp(dredo.numpy(), redo.a, tag + " redo flags")
    return nb_
def exch(redo):
    r = redo.a
    r[:, 0, :, :] = r[:, -2, :, :]; r[:, -1, :, :] = r[:, 1, :, :]
    r[:, :, 0, :] = r[:, :, -2, :]; r[:, :, -1, :] = r[:, :, 1, :]
    r[:, :, :, 0] = r[:, :, :, -2]; r[:, :, :, -1] = r[:, :, :, 1]
def stage(F, V, U0, Uout, tag):
    redo = ol.HostFab(g1, 1, np.int32)
    n1 = update(F, V, U0, Uout, redo, tag + " pass1")
    n2 = n1
    if n1 > 0:
        exch(redo)
        for d in range(3):
            o.orc_replace_fluxes(d, one(F[d].desc()), one(FO[d].desc()), one(redo.desc()), one(vb), 6)
            o.orc_replace_fluxes(d, one(V[d].desc()), one(FOV[d].desc()), one(redo.desc()), one(vb), 1)
        n2 = update(F, V, U0, Uout, redo, tag + " pass2")
    st_ = ol.HostFab(vb, 6); st_.a[...] = Uout.view(vb); dst = dev(st_)
    o.orc_enforce_limits(one(prm), one(st_.desc()), one(vb)); check(lib.qk_hydro_enforce_limits(one(prm), 1, one(vb), one(dst.desc()), None))
    cmp(dst.numpy(), st_.a, tag + " enforce")
    o.orc_sync_dual_energy(one(prm), one(st_.desc()), one(vb)); check(lib.qk_hydro_sync_dual_energy(one(prm), 1, one(vb), one(dst.desc()), None, None))
    cmp(dst.numpy(), st_.a, tag + " sync")
    Uout.view(vb)[...] = st_.a
    return n1, n2
U0 = ol.HostFab(g4, 6); U0.a[...] = st[0]; fill(U0)
FO, FOV = fluxes(U0, fo=True, tag="FO")
F0, V0 = fluxes(U0, tag="S1")
frk = [ol.HostFab(f.box, 6) for f in F0]; avg = [ol.HostFab(v.box, 1) for v in V0]
for d in range(3):
    o.orc_saxpy(one(frk[d].desc()), 0.5, one(F0[d].desc()), one(F0[d].box), 6); o.orc_saxpy(one(avg[d].desc()), 0.5, one(V0[d].desc()), one(V0[d].box), 1)
U1 = ol.HostFab(g4, 6); U2 = ol.HostFab(g4, 6)
print("stage1", stage(F0, V0, U0, U1, "S1"))
fill(U1)
print("U1: neg rho", (U1.view(vb)[0] <= 0).sum(), "nan", np.isnan(U1.a).sum(), "neg E", (U1.view(vb)[4] <= 0).sum())
F1, V1 = fluxes(U1, tag="S2")
for d in range(3):
    o.orc_saxpy(one(frk[d].desc()), 0.5, one(F1[d].desc()), one(F1[d].box), 6); o.orc_saxpy(one(avg[d].desc()), 0.5, one(V1[d].desc()), one(V1[d].box), 1)
print("stage2", stage(frk, avg, U0, U2, "S2"))
