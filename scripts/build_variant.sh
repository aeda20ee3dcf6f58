#!/bin/bash
# build_variant.sh NAME [-DFLAG ...]: A/B build of the CUDA library into wembed_b200/lib/variants/libwb_NAME.so
set -e
cd "$(dirname "$0")/.."
mkdir -p wembed_b200/lib/variants
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC,-pthread -shared \
  "$@" -o wembed_b200/lib/variants/libwb_$name.so wembed_b200/csrc/wb_api.cu -ldl
echo built $name
