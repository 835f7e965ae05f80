#!/usr/bin/env bash
# compute-sanitizer memcheck over the table re-binding path (k_companions, k_monotone_rows, staged DMAs)
set -u
O=gpurun_out; T=${1:-r2H}
mkdir -p $O
[ -z "${SKIP_MEMCHECK:-}" ] && timeout 50 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_rebind.py > $O/${T}_memcheck_rebind.log 2>&1; echo "memcheck rc=$?"; tail -9 $O/${T}_memcheck_rebind.log
timeout 50 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_rebind.py > $O/${T}_racecheck_rebind.log 2>&1; echo "racecheck rc=$?"; tail -9 $O/${T}_racecheck_rebind.log
