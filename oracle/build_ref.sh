#!/usr/bin/env bash
# oracle/build_ref.sh -- builds the UNMODIFIED reference (TREKIS-3, Fortran 2003) with gfortran into oracle/_ref/ and,
# with `run`, produces the level-1 golden tables and the CPU timing of the Monte-Carlo section.
#
# Status (round 2): no Fortran compiler exists in the build container NOR on the B200 box
# (probe committed as profiles/r2_probe_fortran.txt: gfortran, mpif90, ifx, flang, nvfortran, f951 all absent), so this
# recipe has never been executed.  It is committed so that any machine with gfortran >= 10 can pin the oracle:
#
#   oracle/build_ref.sh build   [/path/to/TREKIS-3]      -> oracle/_ref/TREKIS.x (OpenMP build)
#   oracle/build_ref.sh run C1  [/path/to/TREKIS-3]      -> oracle/_ref/run_C1/ (cache tables + MC timing)
#   oracle/build_ref.sh golden C1                        -> tests/golden/ref/C1/ (the cache files tests/test_reference_golden.py reads)
#
# Nothing is copied into the repository from the reference except the GENERATED outputs under tests/golden/ref/.
# Sources are patched in a scratch copy (the reference tree is read-only); the patch set is the one BASELINE.md section 3.1
# lists plus a stub of Check_EPICS_files for materials whose .cdf is complete (the EPICS2023 blobs are not redistributable).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
cmd="${1:-build}"
out="$here/_ref"
mkdir -p "$out"

need_fortran() {
    FC="${FC:-gfortran}"
    if ! command -v "$FC" >/dev/null 2>&1; then
        echo "build_ref.sh: no Fortran compiler ($FC) on this machine -- the reference cannot be built here" >&2
        exit 3
    fi
}

patch_sources() {   # $1 = reference root -> scratch copy in $out/src
    local ref="$1"
    rm -rf "$out/src"; mkdir -p "$out/src"
    cp "$ref"/Source_files/*.f90 "$out/src/"
    # (a) cpp wants double quotes (Universal_MC_for_SHI_MAIN.f90:48-59)
    sed -i "s/^#include '\(.*\)'/#include \"\1\"/" "$out/src/Universal_MC_for_SHI_MAIN.f90"
    # (b) typo inside the __GFORTRAN__ branch (Reading_files_and_parameters.f90:441)
    sed -i 's/inquire(FILE==trim/inquire(FILE=trim/' "$out/src/Reading_files_and_parameters.f90"
    # (c) EPICS2023 blobs absent: report "found" so that complete .cdf files run (Dealing_with_EADL.f90:867-892);
    #     with the blobs present (TRK3_HAVE_EPICS=1) the routine is left alone
    if [ "${TRK3_HAVE_EPICS:-0}" != "1" ]; then
        python3 - "$out/src/Dealing_with_EADL.f90" <<'PY'
import re, sys
p = sys.argv[1]; s = open(p, encoding="latin-1").read()
s = re.sub(r"(inquire\(file=trim\(adjustl\(File_name2?\)\),exist=file_exist\)[^\n]*\n)", r"\1    file_exist = .true. ! build_ref.sh stub\n", s)
open(p, "w", encoding="latin-1").write(s)
PY
    fi
    # (d) event counter for events/s (5 lines inside grid_do is enough; the oracle counts the same way): optional, off by default
}

case "$cmd" in
build)
    need_fortran
    ref="${2:-/root/reference}"
    patch_sources "$ref"
    ( cd "$out/src" && "$FC" -O3 -march=native -cpp -ffree-line-length-none -fdec -fdec-format-defaults \
        -fdefault-real-8 -fdefault-double-8 -std=legacy -fopenmp -w Universal_MC_for_SHI_MAIN.f90 -o "$out/TREKIS.x" )
    echo "built $out/TREKIS.x"
    ;;
run)
    need_fortran
    cfg="${2:?config C1..C5}"; ref="${3:-/root/reference}"
    [ -x "$out/TREKIS.x" ] || "$0" build "$ref"
    run="$out/run_$cfg"; rm -rf "$run"; mkdir -p "$run"
    python3 - "$cfg" "$run" "$here/.." <<'PY'
import sys, os
cfg, run, repo = sys.argv[1:4]
sys.path.insert(0, repo)
import trekis3_b200 as tk
tk.make_run_dir(run, cfg, extra_lines=("gnuplot no", "grid 1", "verbose"))
PY
    for d in INPUT_CDF INPUT_DOS INPUT_EADL; do rm -f "$run/$d"; cp -r "$ref/$d" "$run/$d"; done
    ( cd "$run" && ulimit -s unlimited && export OMP_STACKSIZE=1G && "$out/TREKIS.x" > first.log 2>&1 || true
      "$out/TREKIS.x" > second.log 2>&1 )       # the second run takes all tables from the cache: its MC section is the timing
    python3 - "$run/second.log" <<'PY'
import re, sys
# stamps "Starting MC iterations:" / "Preparing MC output data:" (Universal_MC_for_SHI_MAIN.f90:267, 278)
t = {}
for line in open(sys.argv[1], errors="replace"):
    for key in ("Starting MC iterations", "Preparing MC output data"):
        if key in line:
            m = re.search(r"(\d+):(\d+):(\d+)[.:](\d+)", line)
            if m:
                h, mi, s, ms = map(int, m.groups()); t[key] = h * 3600 + mi * 60 + s + ms / 1000.0
if len(t) == 2:
    print("MC section: %.3f s" % (t["Preparing MC output data"] - t["Starting MC iterations"]))
else:
    print("stamps not found; inspect", sys.argv[1])
PY
    ;;
golden)
    cfg="${2:?config}"; run="$out/run_$cfg"
    dst="$here/../tests/golden/ref/$cfg"; mkdir -p "$dst"
    # level-1 goldens: the reference's own cache tables (Analytical_IMFPs.f90:590-760, 2657-2707)
    find "$run" -maxdepth 2 -name 'OUTPUT_*_IMFPs_*.dat' -o -maxdepth 2 -name 'OUTPUT_*_EMFPs_*.dat' | while read -r f; do cp "$f" "$dst/"; done
    find "$run" -maxdepth 2 -type d -name 'OUTPUT_*_in_*' | while read -r d; do cp "$d"/OUTPUT_*_IMFP.dat "$d"/OUTPUT_*_dEdx.dat "$dst/" 2>/dev/null || true; done
    echo "goldens in $dst -- commit them; tests/test_reference_golden.py compares build_tables() with them at 1e-12"
    ;;
*)
    echo "usage: $0 build|run <cfg>|golden <cfg> [reference root]" >&2; exit 2;;
esac
