#!/bin/bash
# r02_ab_dyn_gpu.sh [tag]: the in-tree library (P1b / P2 tiles pulled from a queue in the exact mode) against
# gpurun_variants/base.so (static snake schedule, -DCVO_DYN_TILES=0): the GPU parity tests on the in-tree build, then
# the same short bench, the single pair and the dense pair on both, same box.
tag=${1:-r02w}
out=gpurun_out; mkdir -p $out
( time timeout 240 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider ) > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc $?" >> $out/${tag}_pytest_gpu.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --side-legs 0 --sequence-frames 0"
for v in base dyn; do
  if [ $v = base ]; then export CVO_B200_LIB=$PWD/gpurun_variants/base.so; else unset CVO_B200_LIB; fi
  timeout 120 $B > $out/${tag}_ab_${v}.json.log 2> $out/${tag}_ab_${v}.err
  timeout 60 python scripts/time_single.py 5 > $out/${tag}_ab_${v}_c1.log 2>&1
  timeout 60 python scripts/bench_dense.py > $out/${tag}_ab_${v}_c3.log 2>&1
done
