#!/bin/bash
# A/B builds: profiles/build_variant.sh NAME -DPLB_X=...   ->  profiles/variants/libplb_NAME.so  (use with PLB_LIB=...)
set -e
cd "$(dirname "$0")/../petlion.jl_b200/csrc"
name=$1; shift
out=../../profiles/variants; mkdir -p $out/obj_$name
for u in plb_kernels plb_variant_iso plb_variant_th plb_variant_sei plb_variant_wide plb_variant_wsei plb_variant_wth plb_variant_thsei plb_variant_wthsei plb_variant_isodc plb_variant_widedc plb_variant_isomhc plb_variant_thmhc plb_variant_seimhc plb_variant_isolgm plb_variant_thlgm; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o $out/obj_$name/$u.o $u.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libplb_$name.so $out/obj_$name/*.o
rm -rf $out/obj_$name
ls -la $out/libplb_$name.so
