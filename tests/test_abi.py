"""The C-ABI boundary without a GPU: libquokka_b200.so loads, exports exactly the functions that
include/quokka_b200.h declares, its structs have the layout the header (and amrex::Array4<double>,
extern/amrex/Src/Base/AMReX_Array4.H:59-68) promise, and -- on a machine without a CUDA device -- every
compute entry refuses with QK_ERR_NO_DEVICE instead of falling back to the CPU.
"""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from quokka_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "quokka_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|int64_t|double|char|qk_level)\s*\**\s*(qk_[a-z0-9_]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_and_binding_table_agree():
    decl = declared_functions()
    assert len(decl) > 50
    assert sorted(capi.SYMBOLS) == decl


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    for name in declared_functions():
        assert getattr(lib, name) is not None
    assert lib.qk_abi_version() == 1
    assert lib.qk_error_string(capi.QK_ERR_NO_DEVICE)


def test_struct_layout_matches_the_header():
    """compile a probe against the header with gcc and compare sizeof/offsetof with the ctypes mirror"""
    fields = {
        "qk_array4": ["p", "jstride", "kstride", "nstride", "begin", "end", "ncomp"],
        "qk_box": ["lo", "hi"],
        "qk_hydro_params": [f for f, _ in capi.qk_hydro_params._fields_],
        "qk_level_desc": [f for f, _ in capi.qk_level_desc._fields_],
        "qk_copy_tag": [f for f, _ in capi.qk_copy_tag._fields_],
        "qk_rad_params": [f for f, _ in capi.qk_rad_params._fields_],
        "qk_rad_source_params": [f for f, _ in capi.qk_rad_source_params._fields_],
    }
    lines = ["#include <stdio.h>", "#include <stddef.h>", f'#include "{HEADER}"', "int main(void){"]
    for st, fl in fields.items():
        lines.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for f in fl:
            lines.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "probe.c")
        open(src, "w").write("\n".join(lines))
        exe = os.path.join(tmp, "probe")
        subprocess.check_call(["gcc", "-std=c99", "-o", exe, src])
        out = subprocess.check_output([exe], text=True)
    got = dict(l.split() for l in out.splitlines())
    for st, fl in fields.items():
        cls = getattr(capi, st)
        assert int(got[st]) == C.sizeof(cls), st
        for f in fl:
            assert int(got[f"{st}.{f}"]) == getattr(cls, f).offset, (st, f)
    # amrex::Array4<double>: {T* p; Long jstride, kstride, nstride; Dim3 begin, end; int ncomp;} = 64 bytes
    assert C.sizeof(capi.qk_array4) == 64


def test_compute_entries_refuse_without_a_device():
    lib = capi.load()
    if lib.qk_device_count() > 0:
        pytest.skip("a CUDA device is present")
    prm = capi.hydro_params()
    box = capi.qk_box.make((0, 0, 0), (7, 7, 7))
    a = capi.qk_array4()
    rc = lib.qk_hydro_conserved_to_primitive(C.byref(prm), 1, C.byref(box), C.byref(a), C.byref(a), 0, None)
    assert rc == capi.QK_ERR_NO_DEVICE
    m = C.c_double()
    assert lib.qk_hydro_max_signal_speed(C.byref(prm), 0, 1, C.byref(box), C.byref(a), C.byref(m), None) == capi.QK_ERR_NO_DEVICE
    rp, sp = capi.rad_params(), capi.rad_source_params()
    assert lib.qk_rad_add_source_terms(C.byref(prm), C.byref(rp), C.byref(sp), 1, 1, C.byref(box), C.byref(a), None, 1.0, None, None) == capi.QK_ERR_NO_DEVICE
    r3 = (C.c_int * 3)(2, 2, 2)
    bcs = (C.c_int32 * 18)()
    assert lib.qk_amr_average_down(1, C.byref(a), 0, C.byref(a), 0, 1, C.byref(box), r3, None) == capi.QK_ERR_NO_DEVICE
    assert lib.qk_amr_interp_cons_lin_minmax(1, C.byref(a), 0, C.byref(a), 0, 6, C.byref(box), C.byref(box), C.byref(box), r3, bcs, bcs,
                                             None) == capi.QK_ERR_NO_DEVICE
    assert lib.qk_amr_interp_cons_lin_minmax(1, C.byref(a), 0, C.byref(a), 0, 17, C.byref(box), C.byref(box), C.byref(box), r3, bcs, bcs,
                                             None) == capi.QK_ERR_UNSUPPORTED  # more than QK_AMR_MAXCOMP components
    bad, rcp = C.c_int64(), C.c_int64()
    assert lib.qk_selftest_division(1, 0, 16, C.byref(bad), C.byref(rcp)) == capi.QK_ERR_NO_DEVICE
    with pytest.raises(RuntimeError):
        from quokka_b200.problems import SedovProblem
        from quokka_b200.simulation import HydroSimulation

        HydroSimulation(SedovProblem(16, 16))


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.load(str(tmp_path / "libquokka_b200.so"))


def test_bench_reference_arm_runs_the_reference_executable():
    """bench.py --impl reference times the reference's own CPU build (oracle/_ref) and prints the contract's JSON line"""
    import json
    import sys

    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "test_hydro3d_blast")):
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "reference"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "Mcell-updates/s"
