#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG ..." : experiment build of the library into heat_b200/variants/libhk_NAME.so
set -e
name=$1; flags=$2
mkdir -p heat_b200/variants/obj_$name
objs=""
for f in heat_b200/csrc/*.cu; do
  o=heat_b200/variants/obj_$name/$(basename ${f%.cu}).o
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $flags -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -shared -o heat_b200/variants/libhk_$name.so $objs -ldl -gencode arch=compute_100a,code=sm_100a
echo built heat_b200/variants/libhk_$name.so
