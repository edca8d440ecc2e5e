#!/bin/bash
# usage: scripts/build_variant.sh NAME [-DFLAG ...]   ->  variants/libofb200_NAME.so (development aid)
set -e
name=$1; shift
mkdir -p variants/$name
for f in $(cd openfermion_b200/csrc && ls *.cu | sed 's/\.cu$//'); do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default "$@" -Xptxas -v -I include -c openfermion_b200/csrc/$f.cu -o variants/$name/$f.o 2> variants/$name/$f.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libofb200_$name.so variants/$name/*.o -lcudart
