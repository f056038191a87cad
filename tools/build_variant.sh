#!/bin/bash
# developer A/B: build_variant.sh <name> <source.cu> [-DFLAG ...]  ->  build_ab/libfa_<name>.so
set -e
cd "$(dirname "$0")/../tf_flash_attention_b200/csrc"
name=$1; src=$2; shift 2
mkdir -p ../../build_ab/obj
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c $src -o ../../build_ab/obj/${name}.o
objs=""
for f in fa_api fa_generic fa_partial fa_layout fa_pack fa_ring fa_plan fa_f64_dmma fa_fwd_f16_sm100 fa_bwd_f16_sm100 fa_fwd_f32_sm100 fa_bwd_f32_sm100; do
  if [ "$f.cu" == "$src" ]; then objs="$objs ../../build_ab/obj/${name}.o"; else objs="$objs build/$f.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build_ab/libfa_${name}.so $objs -lcudart
echo built build_ab/libfa_${name}.so
