"""The fused sweeps replace `a / b` by a shared-reciprocal sequence (csrc/qk_div.cuh) that must return the SAME
bits as the compiler's IEEE-754 division.  Checked here on the device over 2^32 pairs: random bit patterns,
moderate exponents, and special values."""
import ctypes as C

import pytest

from quokka_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,npairs", [(0, 1 << 31), (1, 1 << 31), (2, 1 << 28)])
def test_shared_reciprocal_division_is_ieee_exact(mode, npairs):
    lib = capi.load()
    bad_div, bad_rcp = C.c_int64(-1), C.c_int64(-1)
    capi.check(lib.qk_selftest_division(20261017 + mode, mode, npairs, C.byref(bad_div), C.byref(bad_rcp)))
    print(f"mode {mode}: {npairs} pairs, quotient mismatches {bad_div.value}, refined reciprocal != 1/b: {bad_rcp.value}")
    assert bad_div.value == 0
