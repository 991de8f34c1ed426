#!/bin/bash
# r02_final_gpu.sh [tag]: the round's evidence run on the GPU box, most important first — GPU parity tests, the default
# bench line, the CPU reference arm, the ncu launch list of the bench command, one `ncu --set full` capture of
# k_align_batch per mode (full size: 8 192 pairs), summarised on the box (the reports themselves are kept only
# when they are small).  Everything goes to gpurun_out/<tag>_*.
tag=${1:-r02z}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/${tag}_env.txt 2>&1
( time timeout 780 python -m pytest tests -m gpu -q --timeout 400 --durations=8 -p no:cacheprovider ) > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc $?" >> $out/${tag}_pytest_gpu.log
( time timeout 420 python bench.py ) > $out/${tag}_bench_n1.json.log 2> $out/${tag}_bench_n1.err
( time timeout 240 python bench.py --impl reference --steps 2 --warmup 1 ) > $out/${tag}_bench_reference_n1.json.log 2> $out/${tag}_bench_reference_n1.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --side-legs 0 --sequence-frames 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_launches_bench.log 2>&1
python scripts/summarise_launches.py $out/${tag}_launches.csv "ncu launch list of: $B (3 warm-up + 2 timed device steps, then the end-to-end leg)" > $out/${tag}_launches.txt 2>&1
for m in 0 1; do
  name=$([ $m = 0 ] && echo exact || echo fast)
  timeout 420 ncu --set full --clock-control none --import-source on -k regex:k_align_batch --launch-skip 3 -c 1 -f -o $out/${tag}_align_$name \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --side-legs 0 --sequence-frames 0 --exp-mode $m > $out/${tag}_ncu_$name.log 2>&1
  rep=$out/${tag}_align_$name.ncu-rep
  if [ -f $rep ]; then
    python scripts/ncu_summary.py $rep "ncu --set full --clock-control none, k_align_batch<$name>, the bench's launch: 8192 pairs over 1024 frames (launch 4: after 3 warm-up steps)" > $out/${tag}_align_batch_${name}_full.txt 2>&1
    python scripts/ncu_phases.py $rep k_align_batchILb$((1 - m)) "P1b" "P2 entries" > $out/${tag}_align_batch_${name}_phases.txt 2>&1
    sz=$(stat -c %s $rep)
    if [ $sz -gt 25000000 ]; then rm -f $rep; fi
  fi
done
ls -la $out > $out/${tag}_ls.txt
