#!/bin/bash
# build_variant.sh NAME [-DFLAG=VALUE ...]: an A/B build of the library into build/ab/NAME.so (git-ignored; travels
# to the GPU box), selected at run time with RSPL_BA_LIB=build/ab/NAME.so
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
mkdir -p build/ab
env -u CXX -u CC nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -shared "$@" \
  -o build/ab/$name.so rspl_slam_b200/csrc/capi.cu
echo built build/ab/$name.so "$@"
