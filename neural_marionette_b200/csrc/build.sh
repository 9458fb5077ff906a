#!/bin/bash
# Build libnm_b200.so for sm_100a (in-tree; the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr ${NM_NVCC_EXTRA}"
mkdir -p _obj
pids=()
for f in api voxelize pointwise conv_direct conv_pw conv_tc heatmap dynamics eval conv_wgrad gn_backward conv_wgrad_gather conv_wgrad_tc backward; do
  if [ ! -f _obj/$f.o ] || [ $f.cu -nt _obj/$f.o ] || [ common.cuh -nt _obj/$f.o ] || [ tc_ptx.cuh -nt _obj/$f.o ] || [ ../../include/nm_b200.h -nt _obj/$f.o ]; then
    $NVCC $FLAGS -c $f.cu -o _obj/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libnm_b200.so _obj/*.o
echo "built $(pwd)/libnm_b200.so"
