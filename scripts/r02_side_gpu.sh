#!/bin/bash
# r02_side_gpu.sh [tag]: one GPU — the ncu launch list of the bench command (this library's kernels only: the
# synthetic renderer's torch kernels are not part of the step), the per-rank share of the 8-GPU job on one GPU
# (1 024 pairs: what the tail of the pair queue costs), the dense C3 pair with one `ncu --set full` capture of
# k_align_coop, and the single-pair timing with its phase cycles.
tag=${1:-r02z}
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --side-legs 0 --sequence-frames 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 4000 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_launches_bench.log 2>&1
python scripts/summarise_launches.py $out/${tag}_launches.csv "ncu launch list (kernels of libcvo_b200.so) of: $B (3 warm-up + 2 timed device steps, then 2 + 2 end-to-end steps)" > $out/${tag}_launches.txt 2>&1
for m in 0 1; do
  timeout 200 python bench.py --frames 128 --steps 5 --warmup 3 --no-cpu-baseline --side-legs 0 --sequence-frames 0 --exp-mode $m > $out/${tag}_bench_1024pairs_m$m.json.log 2> $out/${tag}_bench_1024pairs_m$m.err
done
timeout 200 python scripts/bench_dense.py > $out/${tag}_bench_dense_c3.json.log 2>&1
timeout 200 python scripts/time_single.py 7 > $out/${tag}_time_single_c1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_align_coop --launch-skip 2 -c 1 -f -o $out/${tag}_c3_coop python scripts/bench_dense.py > $out/${tag}_ncu_c3.log 2>&1
rep=$out/${tag}_c3_coop.ncu-rep
if [ -f $rep ]; then
  python scripts/ncu_summary.py $rep "ncu --set full --clock-control none, k_align_coop<exact>: C3 dense pair (18 k points per cloud), one alignment on a cooperative grid of 128 CTAs" > $out/${tag}_c3_coop_full.txt 2>&1
  python scripts/ncu_phases.py $rep k_align_coopILb1 > $out/${tag}_c3_coop_phases.txt 2>&1
  if [ $(stat -c %s $rep) -gt 25000000 ]; then rm -f $rep; fi
fi
