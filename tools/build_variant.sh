#!/bin/bash
# Experiment build of the library with extra compiler flags, next to the product build:
#   tools/build_variant.sh NAME -DB2_MBAR_NS=0   ->  variants/libb2dt_NAME.so   (select with B2DT_LIB=variants/libb2dt_NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
pkg=$(ls -d yolo*_b200)
mkdir -p variants/obj_$name
for f in $pkg/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I include -I $pkg/csrc "$@" -c $f -o variants/obj_$name/$(basename $f .cu).o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libb2dt_$name.so variants/obj_$name/*.o
echo variants/libb2dt_$name.so
