#!/bin/bash
# scripts/build_variant.sh <name> [-DKNOB=value ...]  ->  build_variants/<name>.so
# (compile-time experiments; scripts/variant_bench*.py time every .so in build_variants/)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build_variants/obj_$name
for s in kernels extended host_pipeline lightcurve peer; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
       -c caustics_b200/csrc/$s.cu -o build_variants/obj_$name/$s.o &
done
wait
nvcc -shared -o build_variants/$name.so build_variants/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a
echo build_variants/$name.so
