#!/bin/bash
# r02_bounds_gpu.sh [tag]: the bounds-checked build of the library (-DCVO_BOUNDS: every index of the list machinery is
# tested on the device; violations are counted, reported on stderr with the phase cycles) on the bench workload
# (1 024 pairs, both modes), the dense pair and the single pair.  compute-sanitizer is not available on this pool.
tag=${1:-r02z}
out=gpurun_out; mkdir -p $out
export CVO_B200_LIB=$PWD/gpurun_variants/bounds.so
{
  for m in 0 1; do
    echo "== bench.py --frames 128 --exp-mode $m (1 024 pairs)"
    timeout 150 python bench.py --frames 128 --steps 2 --warmup 1 --no-cpu-baseline --side-legs 0 --sequence-frames 0 --exp-mode $m 2>&1 >/dev/null | grep -v "^$"
  done
  echo "== scripts/bench_dense.py (C3, cooperative grid)"
  timeout 100 python scripts/bench_dense.py 2>&1 >/dev/null
  echo "== scripts/time_single.py (C1, cluster of 16)"
  timeout 100 python scripts/time_single.py 3 2>&1 | grep "bounds"
  echo "== scripts/sanitize_small.py all"
  timeout 100 python scripts/sanitize_small.py all 2>&1 | tail -3
} > $out/${tag}_bounds_build.log 2>&1
