#!/bin/bash
# developer builds of the library with extra -D flags:  tools/build_variant.sh NAME -DFLAG ...  -> build_tl/libamgb200_NAME.so
name=$1; shift
mkdir -p build_tl/obj_$name
for f in amg_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off,-fopenmp -Iinclude "$@" -c $f -o build_tl/obj_$name/$b.o &
done; wait
hostobjs=$(ls amg_b200/csrc/*.cpp | grep -v debug_host | sed 's/\.cpp$/.o/')
nvcc -shared -gencode arch=compute_100a,code=sm_100a -Xcompiler -fopenmp -o build_tl/libamgb200_$name.so build_tl/obj_$name/*.o $hostobjs -lcudart
