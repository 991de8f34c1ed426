#!/bin/bash
# r02_sanitize_gpu.sh [tag]: compute-sanitizer over scripts/sanitize_small.py (every kernel of the library, small
# sizes): memcheck on the full tour, racecheck and synccheck on the short one.  Every run in its own process group
# with a hard limit, so that a run that stalls under the tool cannot outlive the call.
tag=${1:-r02z}
out=gpurun_out; mkdir -p $out
run() {   # run <limit-seconds> <log> <command...>
  local tmo=$1 log=$2; shift 2
  setsid "$@" > $log 2>&1 &
  local pid=$!
  ( sleep $tmo; kill -KILL -- -$pid 2>/dev/null ) &
  local w=$!
  wait $pid; local rc=$?
  kill $w 2>/dev/null
  echo "exit code $rc" >> $log
}
run 45 $out/${tag}_sanitize_plain.log python scripts/sanitize_small.py all
run 120 $out/${tag}_memcheck.log compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python scripts/sanitize_small.py all
run 110 $out/${tag}_racecheck.log compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --print-limit 20 python scripts/sanitize_small.py small
run 60 $out/${tag}_synccheck.log compute-sanitizer --tool synccheck --error-exitcode 7 --print-limit 20 python scripts/sanitize_small.py small
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv > $out/${tag}_sanitize_after.txt 2>&1
