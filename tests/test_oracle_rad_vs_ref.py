"""Pin the radiation part of the C oracle (oracle/quokka_oracle.c: orc_rad_*) against the REFERENCE's own RadSystem<problem_t>
templates compiled from /root/reference (oracle/_ref/libquokka_ref.so, oracle/ref_build/ref_harness.cpp).  Bar: bit-exact."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from quokka_b200.capi import QK_MC, qk_box, qk_rad_params

pytestmark = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libquokka_ref.so not built")

VALID = qk_box.make((2, -3, 4), (13, 6, 11))
NG = 4
DX = (C.c_double * 3)(0.1, 0.07, 0.13)


def exact(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    assert not bad.any(), f"{bad.sum()} mismatches of {a.size}"


def rparams(problem):
    p = qk_rad_params()
    assert ol.ref().ref_rad_params(problem, C.byref(p)) == 0
    return p


def make_cons(prm, kind, seed=777):
    gb = VALID.grown(NG)
    f = ol.HostFab(gb, prm.nstart + 4 * prm.ngroups)
    f.a[...] = ol.random_rad_cons(gb, prm, seed, kind)
    return f


def prim_of(prm, cons):
    gb = VALID.grown(NG)
    q = ol.HostFab(gb, 4 * prm.ngroups)
    with np.errstate(all="ignore"):
        ol.oracle().orc_rad_conserved_to_primitive(C.byref(prm), C.byref(cons.desc()), C.byref(q.desc()), C.byref(gb))
    return q


@pytest.mark.parametrize("problem", [0, 1])
@pytest.mark.parametrize("kind", ["smooth", "beam"])
def test_rad_cons_to_prim(problem, kind):
    prm = rparams(problem)
    cons = make_cons(prm, kind)
    qo = prim_of(prm, cons)
    qr = ol.HostFab(VALID.grown(NG), 4 * prm.ngroups)
    ol.ref().ref_rad_cons_to_prim(problem, C.byref(VALID), C.byref(cons.desc()), C.byref(qr.desc()), NG)
    exact(qo.a, qr.a)


def recon(prm, q, order, d):
    g1 = VALID.grown(1)
    fb = ol.face_box(VALID, d, 1)
    l, r = ol.HostFab(fb, 4 * prm.ngroups), ol.HostFab(fb, 4 * prm.ngroups)
    ol.oracle().orc_reconstruct_states(order, QK_MC, d, C.byref(q.desc()), C.byref(l.desc()), C.byref(r.desc()), C.byref(g1 if order == 3 else fb),
                                       4 * prm.ngroups)
    return l, r


@pytest.mark.parametrize("problem", [0, 1])
@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("d", [0, 1, 2])
@pytest.mark.parametrize("kind", ["smooth", "beam"])
def test_rad_compute_fluxes(problem, order, d, kind):
    prm = rparams(problem)
    cons = make_cons(prm, kind)
    q = prim_of(prm, cons)
    l, r = recon(prm, q, order, d)
    fb = ol.face_box(VALID, d)
    fo, fdo, fr, fdr = (ol.HostFab(fb, 4 * prm.ngroups) for _ in range(4))
    with np.errstate(all="ignore"):
        ol.oracle().orc_rad_compute_fluxes(C.byref(prm), d, C.byref(fo.desc()), C.byref(fdo.desc()), C.byref(l.desc()), C.byref(r.desc()),
                                           C.byref(cons.desc()), C.byref(fb))
    ol.ref().ref_rad_compute_fluxes(problem, d, C.byref(VALID), C.byref(fr.desc()), C.byref(fdr.desc()), C.byref(l.desc()), C.byref(r.desc()),
                                    C.byref(cons.desc()), NG)
    exact(fo.a, fr.a)
    exact(fdo.a, fdr.a)
    if kind == "beam":  # the inadmissible-state fallback was exercised
        bad = (l.view(fb)[0] <= 0) | (r.view(fb)[0] <= 0)
        assert bad.any()


@pytest.mark.parametrize("problem", [2, 3])
@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_rad_compute_fluxes_with_wavespeed_correction(problem, order, d):
    """radiation.use_wavespeed_correction = true (radiation_system.hpp:1018-1022,1100-1109 + ComputeCellOpticalDepth :803-871): the energy
    flux's diffusive term is scaled by min(1, 1 / tau_cell) on faces with even i + j + k.  Problem types with opacities (the source-term
    harness types R2 = RadhydroShell cgs, kappa_F = 20; R3 dimensionless, kappa_F = 2); gas densities chosen so that tau straddles 1."""
    from quokka_b200.capi import qk_hydro_params, qk_rad_source_params

    hp, prm, sp = qk_hydro_params(), qk_rad_params(), qk_rad_source_params()
    assert ol.ref().ref_rad_source_params(problem, C.byref(hp), C.byref(prm), C.byref(sp)) == 0
    prm.use_wavespeed_correction = 1
    prm.kappa_F = sp.kappa_F
    prm.cell_dx[:] = list(DX)
    cons = make_cons(prm, "smooth")
    rng = np.random.default_rng(5)
    cons.a[0] = rng.uniform(0.05, 5.0, cons.a[0].shape) * (20.0 / sp.kappa_F)  # tau = dl rho kappa_F in (0.07, 13)
    cons.a[4] = 10.0 * cons.a[0]  # a positive internal energy for the reference's (unused) gas temperature
    q = prim_of(prm, cons)
    l, r = recon(prm, q, order, d)
    fb = ol.face_box(VALID, d)
    fo, fdo, fr, fdr = (ol.HostFab(fb, 4 * prm.ngroups) for _ in range(4))
    ol.oracle().orc_rad_compute_fluxes(C.byref(prm), d, C.byref(fo.desc()), C.byref(fdo.desc()), C.byref(l.desc()), C.byref(r.desc()),
                                       C.byref(cons.desc()), C.byref(fb))
    ol.ref().ref_rad_compute_fluxes_wsc(problem, d, C.byref(VALID), C.byref(fr.desc()), C.byref(fdr.desc()), C.byref(l.desc()), C.byref(r.desc()),
                                        C.byref(cons.desc()), NG, DX)
    exact(fo.a, fr.a)
    exact(fdo.a, fdr.a)
    changed = (fo.a[0] != fdo.a[0])
    assert changed.any() and not changed.all(), "the correction must act on some (even, optically thick) faces only"
    exact(fo.a[1:], fdo.a[1:])  # only the energy component carries epsilon


def fluxes_of(prm, cons, order):
    q = prim_of(prm, cons)
    out = []
    for d in range(3):
        l, r = recon(prm, q, order, d)
        fb = ol.face_box(VALID, d)
        f = ol.HostFab(fb, 4 * prm.ngroups)
        with np.errstate(all="ignore"):
            ol.oracle().orc_rad_compute_fluxes(C.byref(prm), d, C.byref(f.desc()), None, C.byref(l.desc()), C.byref(r.desc()), C.byref(cons.desc()),
                                               C.byref(fb))
        out.append(f)
    return out


@pytest.mark.parametrize("problem", [0, 1])
@pytest.mark.parametrize("kind", ["smooth", "beam"])
@pytest.mark.parametrize("dt", [0.02, 3.0])
def test_rad_predict_step_and_rk2(problem, kind, dt):
    """dt = 3 (far beyond CFL) drives E_r negative and |F| > cE: isStateValid / amendRadState paths"""
    prm = rparams(problem)
    dt = dt / prm.c_hat
    nc = prm.nstart + 4 * prm.ngroups
    u0 = make_cons(prm, kind, 777)
    u1 = make_cons(prm, kind, 778)
    f0 = fluxes_of(prm, u0, 3)
    f1 = fluxes_of(prm, u1, 2)
    # PredictStep
    a, b = ol.HostFab(VALID, nc, fill=-7.0), ol.HostFab(VALID, nc, fill=-7.0)
    with np.errstate(all="ignore"):
        ol.oracle().orc_rad_predict_step(C.byref(prm), C.byref(u0.desc()), C.byref(a.desc()), C.byref(f0[0].desc()), C.byref(f0[1].desc()),
                                         C.byref(f0[2].desc()), dt, DX, C.byref(VALID))
    z = ol.HostFab(VALID, nc)
    ol.ref().ref_rad_update(problem, 0, C.byref(VALID), C.byref(b.desc()), C.byref(u0.desc()), C.byref(z.desc()), None, None, None,
                            C.byref(f0[0].desc()), C.byref(f0[1].desc()), C.byref(f0[2].desc()), dt, DX)
    exact(a.a, b.a)
    assert (a.a[:prm.nstart] == -7.0).all()  # hydro components untouched
    # AddFluxesRK2
    a, b = ol.HostFab(VALID, nc, fill=-7.0), ol.HostFab(VALID, nc, fill=-7.0)
    with np.errstate(all="ignore"):
        ol.oracle().orc_rad_add_fluxes_rk2(C.byref(prm), C.byref(a.desc()), C.byref(u0.desc()), C.byref(u1.desc()), *[C.byref(f.desc()) for f in f0],
                                           *[C.byref(f.desc()) for f in f1], dt, DX, C.byref(VALID))
    ol.ref().ref_rad_update(problem, 1, C.byref(VALID), C.byref(b.desc()), C.byref(u0.desc()), C.byref(u1.desc()), *[C.byref(f.desc()) for f in f0],
                            *[C.byref(f.desc()) for f in f1], dt, DX)
    exact(a.a, b.a)
