#!/bin/bash
# build_variant.sh <name> <nvcc -D flags...> : builds gpurun_variants/<name>.so (kernel-variant experiments; git-ignored)
set -e
name=$1; shift
mkdir -p gpurun_variants/obj_$name
CVO_B200_OUT=$PWD/gpurun_variants/$name.so CVO_B200_OBJDIR=$PWD/gpurun_variants/obj_$name CVO_NVCC_EXTRA="$*" python -m cvo_slam_b200.build --force > gpurun_variants/$name.buildlog 2>&1 || { tail -20 gpurun_variants/$name.buildlog; exit 1; }
rm -rf gpurun_variants/obj_$name
echo built gpurun_variants/$name.so
