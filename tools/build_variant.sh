#!/bin/bash
# usage: tools/build_variant.sh NAME "extra nvcc flags" [--split]
# Builds gpurun_scratch/variants/NAME/libtmpc_cuda.so: k_f32.o recompiled with the extra flags, every other object shared with
# the regular build (development A/B helper; run with TMPC_LIB_PATH=<that file>).
set -e
here=$(cd "$(dirname "$0")/.." && pwd)
pk=$here/accelerated-tinympc_b200
out=$here/gpurun_scratch/variants/$1
mkdir -p $out
split=""
[ "$3" == "--split" ] && split="-split-compile 0"   # (nondeterministic code generation: A/B only)
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false $split -Xcompiler -fPIC -Xcompiler -pthread \
     -I$here/include -I$pk/csrc $2 -c -o $out/k_f32.o $pk/csrc/k_f32.cu
nvcc -shared -o $out/libtmpc_cuda.so $pk/build/tmpc_api.o $out/k_f32.o $pk/build/k_generic.o $pk/build/k_small_warp.o $pk/build/k_sys.o $pk/build/k_f64p.o -lcudart
echo built $out/libtmpc_cuda.so
