#!/bin/bash
# tuning build of libuavsim.so with extra -D switches:  tools/build_variant.sh <name> [-DFAST_... ...]
# -> variants/libuavsim_<name>.so (git-ignored, travels to the GPU box; selected with UAVSIM_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../marl_uavs_targets_tracking_b200/csrc"
mkdir -p ../../variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static \
  "$@" -o ../../variants/libuavsim_${name}.so uavsim.cu
echo built variants/libuavsim_${name}.so
