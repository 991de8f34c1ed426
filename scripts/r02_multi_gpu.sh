#!/bin/bash
# r02_multi_gpu.sh [tag] [N]: the sharded (strong-scaling) job on N GPUs of one box: the multi-device tests, then the
# bench under torchrun exactly as the driver launches it, and the CPU reference arm under the same launcher.
tag=${1:-r02z}; N=${2:-2}
out=gpurun_out; mkdir -p $out
nvidia-smi -L > $out/${tag}_n${N}_env.txt 2>&1
( time timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_device or sharded" -p no:cacheprovider ) > $out/${tag}_n${N}_pytest.log 2>&1
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 3 --warmup 3 ) > $out/${tag}_bench_n${N}.json.log 2> $out/${tag}_bench_n${N}.err
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 0 ) > $out/${tag}_bench_reference_n${N}.json.log 2> $out/${tag}_bench_reference_n${N}.err
