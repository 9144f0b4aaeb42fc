"""include/quokka_b200_amrex.hpp (the C++ shim a Quokka maintainer would add) compiles against the reference's own headers and
offers the call signatures QuokkaSimulation<problem_t> uses for HydroSystem<problem_t>.  Compile-only; needs /root/reference."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src") or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "gen", "extern_parameters.H")),
                    reason="reference headers not present")
def test_shim_compiles_against_reference_headers():
    r = subprocess.run([os.path.join(ROOT, "tests", "integration", "check.sh")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHIM_OK" in r.stdout, r.stderr[-4000:]
