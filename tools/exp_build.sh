#!/bin/bash
# Scratch build of libntt_b200.so for kernel A/B timing: tools/exp_build.sh <name> [-DFLAG ...]
# -> build_exp/libntt_b200_<name>.so (only the L = 14 FP64 ring kernels; the rest of the library as usual).
# Time it on the GPU with: python tools/time_lib.py build_exp/libntt_b200_<name>.so 14 check
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=$ROOT/optimized-number-theoretic-transform-implementations_b200
name=$1; shift
OUT=$ROOT/build_exp; mkdir -p $OUT/obj_$name
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DNTT_EXPERIMENT $*"
for f in ntt_ring_fp_10 ntt_ring_fp_11 ntt_ring_fp_12 ntt_ring_fp_13 ntt_ring_fp_14; do
  $NV -c $PKG/csrc/$f.cu -o $OUT/obj_$name/$f.o &
done
wait
# objects that do not depend on the experiment flags come from the in-tree build
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libntt_b200_$name.so $OUT/obj_$name/*.o \
  $PKG/csrc/ntt_kernels.o $PKG/csrc/ntt_ring_int.o $PKG/csrc/ntt_polymul_fp.o $PKG/csrc/ntt_strided_fp.o $PKG/host/ntt_math.o $PKG/host/ntt_plan.o $PKG/host/ntt_dropin.o \
  $PKG/host/ntt_multi.o -Xlinker --version-script=$PKG/exports.map -cudart static -lpthread
echo built $OUT/libntt_b200_$name.so
