#!/bin/bash
# A/B builds: profiles/build_variant.sh NAME -DPLB_X=...   ->  profiles/variants/libplb_NAME.so  (use with PLB_LIB=...)
# PLB_AB_UNITS="plb_variant_iso plb_variant_th": only these units are compiled with the extra flags, the others are taken from
# the default build's objects (petlion.jl_b200/csrc/*.o)
set -e
cd "$(dirname "$0")/../petlion.jl_b200/csrc"
name=$1; shift
out=../../profiles/variants; mkdir -p $out/obj_$name
all=$(ls plb_kernels.cu plb_variant_*.cu | sed 's/\.cu$//')
units=$(echo ${PLB_AB_UNITS:-$all})
for u in $all; do
  if echo " $units " | grep -q " $u "; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o $out/obj_$name/$u.o $u.cu &
  else
    cp $u.o $out/obj_$name/$u.o
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libplb_$name.so $out/obj_$name/*.o
rm -rf $out/obj_$name
ls -la $out/libplb_$name.so
