#!/usr/bin/env bash
# compute-sanitizer over small runs of every kernel family (memcheck; racecheck for the shared-memory closing of the blocks)
set -u
O=gpurun_out; T=${1:-r2s}
mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_target.py 2 > $O/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 $O/${T}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_target.py 1 > $O/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 $O/${T}_racecheck.log
