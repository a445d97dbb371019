#!/bin/bash
# developer builds of the library with extra -D flags:  tools/build_variant.sh NAME -DFLAG ...  -> build_tl/libamgb200_NAME.so
name=$1; shift
mkdir -p build_tl/obj_$name
for f in dropin hier setup_dev; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off,-fopenmp -Iinclude "$@" -c amg_b200/csrc/$f.cu -o build_tl/obj_$name/$f.o &
done; wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -Xcompiler -fopenmp -o build_tl/libamgb200_$name.so build_tl/obj_$name/dropin.o build_tl/obj_$name/hier.o build_tl/obj_$name/setup_dev.o amg_b200/csrc/analysis.o amg_b200/csrc/host_gen.o amg_b200/csrc/host_setup.o amg_b200/csrc/host_mtx.o -lcudart
