#!/bin/bash
# ab_bench.sh <tag> <variant names...>: runs the short bench (exact and fast) for each variant library on the GPU box
tag=$1; shift
for v in "$@"; do
  for m in 0 1; do
    CVO_B200_LIB=$PWD/gpurun_variants/$v.so python bench.py --steps 2 --warmup 2 --sequence-frames 0 --no-cpu-baseline --exp-mode $m > gpurun_out/${tag}_${v}_m$m.log 2> gpurun_out/${tag}_${v}_m$m.err
  done
done
