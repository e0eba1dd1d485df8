#!/bin/bash
# compute-sanitizer over every kernel of libjmpc.so: memcheck and racecheck on small batches of
# the step kernel (whole-warp and half-warp instantiations, generic horizon), the flag kernel, the episode kernels
# (incl. the scripted obstacles) and the planner.  Run on the GPU box from the repo root:
#     bash tests/tools/sanitize.sh [outdir]          (default gpurun_out/sanitize)
# NOTE (round 2): on this build's GPU pool compute-sanitizer is closed ("runs under it have left GPUs needing a reset"),
# every invocation exits 86 with that message -- profiles/r2_sanitizer_refused.txt.  The script is kept for a box where
# it is allowed.
# The step kernel hands its Hessian from the generic proxy to the async proxy (TMA bulk copy) and back by hand
# (jmpc_step.cuh: tma_load_1d / mbar_wait), which is what racecheck is here for.
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
CS=${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
python -c "import __graft_entry__ as g; g.build()" > /dev/null
rc=0
for tool in memcheck racecheck; do
  for case in step_T20 step_T13 step_T8 step_T25 step_T10 collision episodes planner; do
    log="$OUT/${tool}_${case}.log"
    timeout 600 "$CS" --tool $tool --error-exitcode 9 --print-limit 20 python tests/tools/sanitize_driver.py $case > "$log" 2>&1
    code=$?
    summary=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
    echo "$tool $case: exit $code  $summary"
    [ $code -ne 0 ] && rc=1
  done
done
echo "sanitize: overall rc $rc"
exit $rc
