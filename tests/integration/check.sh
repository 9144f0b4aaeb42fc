#!/bin/bash
# g++ -fsyntax-only of tests/integration/shim_check.cpp against the reference's headers (needs /root/reference and oracle/_ref/gen)
set -e
cd "$(dirname "$0")/../.."
REF=${REF:-/root/reference}; AM=$REF/extern/amrex/Src; M=$REF/extern/Microphysics
g++ -std=c++17 -fsyntax-only -w -DNDEBUG -DFMT_HEADER_ONLY -DNAUX_NET -DSTRANG -DAMREX_SPACEDIM=3 -Iinclude -Ioracle/ref_build -Ioracle/_ref/gen \
  -Ioracle/ref_build/stub_hdf5 -I$REF/src $(for d in Base Base/Parser Boundary AmrCore Particle LinearSolvers LinearSolvers/MLMG LinearSolvers/OpenBC; do echo -I$AM/$d; done) \
  -I$REF/extern/fmt/include -I$REF/extern/yaml-cpp/include -I$M/util -I$M/util/gcem/include -I$M/interfaces -I$M/EOS -I$M/EOS/gamma_law \
  -I$M/networks -I$M/networks/general_null -I$M/constants tests/integration/shim_check.cpp
echo SHIM_OK
