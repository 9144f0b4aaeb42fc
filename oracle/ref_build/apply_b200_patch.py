#!/usr/bin/env python3
"""Write a patched copy of the reference's src/QuokkaSimulation.hpp that routes QuokkaSimulation::advanceHydroAtLevel and
subcycleRadiationAtLevel through libquokka_b200 (include/quokka_b200_driver.hpp).

    python3 apply_b200_patch.py /root/reference/src oracle/_ref/cuda/patched

TEST / MEASUREMENT INFRASTRUCTURE.  The output is a build product (oracle/_ref is git-ignored); nothing of the reference
is stored in this repository.  The patch is four anchored insertions, each matched on ONE line of the reference and
checked to match exactly once, so that an upstream change makes the build fail loudly instead of silently mis-patching:

  1. after `#include "simulation.hpp"`                      -> #include "quokka_b200_driver.hpp"   (DriverState)
  2. before `void addStrangSplitSources(` in the class      -> three member declarations + the DriverState member
  3. after `BL_PROFILE("QuokkaSimulation::advanceHydroAtLevel()");`
                                                            -> hook: forward to advanceHydroAtLevelB200
  4. after `// compute radiation timestep` (first line of subcycleRadiationAtLevel's body)
                                                            -> hook: forward to subcycleRadiationAtLevelB200
  5. before the final `#endif`                              -> second include with QUOKKA_B200_DRIVER_IMPL (definitions)
"""
import os
import sys

INCLUDE_ANCHOR = '#include "simulation.hpp"'
DECL_ANCHOR = "\tvoid addStrangSplitSources(amrex::MultiFab &state, int lev, amrex::Real time, amrex::Real dt_lev);"
HYDRO_ANCHOR = '\tBL_PROFILE("QuokkaSimulation::advanceHydroAtLevel()");'
RAD_SIG = "void QuokkaSimulation<problem_t>::subcycleRadiationAtLevel(int lev, amrex::Real time, amrex::Real dt_lev_hydro, amrex::YAFluxRegister *fr_as_crse,"
RAD_ANCHOR = "\t// compute radiation timestep"
END_ANCHOR = "#endif // RADIATION_SIMULATION_HPP_"

DECLS = """\t// ---- libquokka_b200 driver (inserted by oracle/ref_build/apply_b200_patch.py; defined in quokka_b200_driver.hpp) ----
\tauto advanceHydroAtLevelB200(amrex::MultiFab &state_old_cc_tmp, amrex::YAFluxRegister *fr_as_crse, amrex::YAFluxRegister *fr_as_fine, int lev,
\t\t\t\t     amrex::Real time, amrex::Real dt_lev) -> bool;
\tvoid subcycleRadiationAtLevelB200(int lev, amrex::Real time, amrex::Real dt_lev_hydro);
\tauto b200CanFill(int lev) -> bool;
\tquokka::b200::DriverState b200_;
"""
HYDRO_HOOK = """\tif (b200_.on() && do_tracers == 0) {
\t\treturn advanceHydroAtLevelB200(state_old_cc_tmp, fr_as_crse, fr_as_fine, lev, time, dt_lev);
\t}
"""
RAD_HOOK = """\tif (b200_.on() && b200_.radiation != 0 && fr_as_crse == nullptr && fr_as_fine == nullptr && Physics_Traits<problem_t>::is_hydro_enabled && !(constantDt_ > 0.) &&
\t    Physics_Traits<problem_t>::nGroups == 1 && b200CanFill(lev)) {
\t\tsubcycleRadiationAtLevelB200(lev, time, dt_lev_hydro);
\t\treturn;
\t}
"""


def insert(lines, anchor, text, after=True, start=0, what=""):
    hits = [i for i in range(start, len(lines)) if lines[i].rstrip("\n") == anchor]
    if len(hits) != 1:
        raise SystemExit(f"apply_b200_patch: anchor for {what} matched {len(hits)} times (expected 1): {anchor!r}")
    at = hits[0] + (1 if after else 0)
    lines[at:at] = [text if text.endswith("\n") else text + "\n"]
    return at


def main():
    src, out = sys.argv[1], sys.argv[2]
    with open(os.path.join(src, "QuokkaSimulation.hpp")) as f:
        lines = f.readlines()
    insert(lines, INCLUDE_ANCHOR, '#include "quokka_b200_driver.hpp"', what="include")
    insert(lines, DECL_ANCHOR, DECLS, after=False, what="member declarations")
    insert(lines, HYDRO_ANCHOR, HYDRO_HOOK, what="advanceHydroAtLevel hook")
    sig = [i for i, l in enumerate(lines) if l.rstrip("\n") == RAD_SIG]
    if len(sig) != 1:
        raise SystemExit(f"apply_b200_patch: subcycleRadiationAtLevel definition matched {len(sig)} times")
    first = [i for i in range(sig[0], len(lines)) if lines[i].rstrip("\n") == RAD_ANCHOR][:1]
    if not first:
        raise SystemExit("apply_b200_patch: subcycleRadiationAtLevel body anchor not found")
    lines[first[0]:first[0]] = [RAD_HOOK]
    insert(lines, END_ANCHOR, '#define QUOKKA_B200_DRIVER_IMPL\n#include "quokka_b200_driver.hpp"\n', after=False, what="definitions include")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "QuokkaSimulation.hpp"), "w") as f:
        f.writelines(lines)
    print(f"apply_b200_patch: wrote {os.path.join(out, 'QuokkaSimulation.hpp')} (5 insertions)")


if __name__ == "__main__":
    main()
