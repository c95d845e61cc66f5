#!/bin/bash
# Developer tool: build an A/B variant of the library with extra -D flags:  tools_dev/build_variant.sh <name> -DFOO=1 ...
# -> build/var/libspef_<name>.so (travels to the GPU box; select it with SPEF_DEV_LIB=build/var/libspef_<name>.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/var
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC,-pthread "$@" \
  -o build/var/libspef_$name.so spacecraft-pose-estimation-framework_b200/csrc/spef_api.cu spacecraft-pose-estimation-framework_b200/csrc/host_pack.cpp
echo build/var/libspef_$name.so
