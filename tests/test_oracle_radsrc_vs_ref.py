"""Pin the matter-radiation source terms of the C oracle (oracle/quokka_oracle.c: orc_rad_add_source_terms) against the
REFERENCE's own RadSystem<problem_t>::AddSourceTermsSingleGroup (src/radiation/source_terms_single_group.hpp:9-565) compiled
from /root/reference (oracle/_ref/libquokka_ref.so, problems R2..R7 of oracle/ref_build/ref_harness.cpp).  Both run on the
host with the same libm, so the bar is bit-exact, iteration counters included."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from quokka_b200.capi import QK_RAD_SOURCE_NCOUNTERS, qk_box, qk_hydro_params, qk_rad_params, qk_rad_source_params

pytestmark = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libquokka_ref.so not built")

VALID = qk_box.make((3, -2, 5), (18, 9, 12))

# per problem: generator scales (T0, rho0, vmax) and a list of dt_radiation values spanning optically thin -> thick steps
CASES = {
    2: dict(T0=100.0, rho0=1e-19, vmax=3e5, dts=[4e6, 4e8, 4e10]),  # RadhydroShell traits (cgs)
    3: dict(T0=1.0, rho0=1.0, vmax=1.0, dts=[1e-3, 0.1, 10.0]),  # kappa_F != kappa_E, beta_order 2
    4: dict(T0=1.0, rho0=1.0, vmax=1.0, dts=[1e-3, 0.1, 10.0]),  # beta_order 0
    5: dict(T0=1.0, rho0=1.0, vmax=1.0, dts=[1e-3, 0.1, 10.0]),  # beta_order 3, Erad_floor > 0
    6: dict(T0=1.0, rho0=1.0, vmax=1.0, dts=[0.1]),  # gamma = 1
    7: dict(T0=1.0, rho0=1.0, vmax=1.0, dts=[1e-3, 0.1, 10.0]),  # kappa_P = kappa_E = 0: tau = 0 branches
}


def params(problem):
    hp, rp, sp = qk_hydro_params(), qk_rad_params(), qk_rad_source_params()
    assert ol.ref().ref_rad_source_params(problem, C.byref(hp), C.byref(rp), C.byref(sp)) == 0
    return hp, rp, sp


def exact(a, b):
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    assert not bad.any(), f"{bad.sum()} mismatches of {a.size}, max rel {np.nanmax(np.abs(a - b) / (np.abs(b) + 1e-300))}"


@pytest.mark.parametrize("problem", sorted(CASES))
@pytest.mark.parametrize("stage", [1, 2])
@pytest.mark.parametrize("with_source", [False, True])
def test_source_terms_bit_exact(problem, stage, with_source):
    hp, rp, sp = params(problem)
    cs = CASES[problem]
    for n, dt in enumerate(cs["dts"]):
        st = ol.random_radhydro_cons(VALID, hp, rp, sp, seed=100 * problem + n, T0=cs["T0"], rho0=cs["rho0"], vmax=cs["vmax"])
        a, b = ol.HostFab(VALID, rp.nstart + 4), ol.HostFab(VALID, rp.nstart + 4)
        a.a[...] = st
        b.a[...] = st
        src = None
        if with_source:
            src = ol.HostFab(VALID, 1)
            rng = np.random.default_rng(7)
            src.a[...] = rng.uniform(0.0, 2.0, src.a.shape) * st[rp.nstart] / (dt * rp.c_hat)
        ca = (C.c_int64 * QK_RAD_SOURCE_NCOUNTERS)()
        cb = (C.c_int64 * QK_RAD_SOURCE_NCOUNTERS)()
        psrc = C.byref(src.desc()) if src is not None else None
        with np.errstate(all="ignore"):
            ol.oracle().orc_rad_add_source_terms(C.byref(hp), C.byref(rp), C.byref(sp), C.byref(a.desc()), psrc, C.byref(VALID), dt, stage, ca)
            assert ol.ref().ref_rad_add_source_terms(problem, C.byref(VALID), C.byref(b.desc()), psrc, dt, stage, cb) == 0
        exact(a.a, b.a)
        assert list(ca) == list(cb), (list(ca), list(cb))
        if hp.gamma != 1.0:
            ncell = int(np.prod(VALID.shape()))
            assert cb[0] >= ncell  # every cell solved at least once (cb[4] counts cells whose Newton-Raphson loop ran out: compared above)
            assert not np.array_equal(a.a[4], st[4])  # the gas energy did change


def test_energy_and_momentum_are_exchanged_conservatively():
    """known-answer property of the scheme (He, Wibking & Krumholz 2024): with beta_order = 0 and the full step (stage 2 uses
    (1 - a32) dt for both), E_gas + (c/c_hat) E_rad and p_gas + F/(c c_hat) are conserved by the exchange."""
    hp, rp, sp = params(4)
    st = ol.random_radhydro_cons(VALID, hp, rp, sp, seed=5)
    a = ol.HostFab(VALID, rp.nstart + 4)
    a.a[...] = st
    ol.oracle().orc_rad_add_source_terms(C.byref(hp), C.byref(rp), C.byref(sp), C.byref(a.desc()), None, C.byref(VALID), 0.1, 2, None)
    cs = rp.c_light / rp.c_hat
    ns = rp.nstart
    eint0, eint1 = st[5], a.a[5]
    tot0 = eint0 + cs * st[ns]
    tot1 = eint1 + cs * a.a[ns]
    assert np.abs(tot1 - tot0).max() / np.abs(tot0).max() < 1e-10  # Newton residual tolerance 1e-11
    for m in range(3):
        p0 = st[1 + m] + st[ns + 1 + m] / (rp.c_light * rp.c_hat)
        p1 = a.a[1 + m] + a.a[ns + 1 + m] / (rp.c_light * rp.c_hat)
        assert np.abs(p1 - p0).max() <= 1e-14 * np.abs(p0).max()
