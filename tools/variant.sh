#!/bin/bash
# tools/variant.sh NAME [nvcc -D flags...]: builds the CUDA core with extra flags into kuafu_b200/lib_NAME/
# (with a copy of the host library beside it) for A/B runs: KFRT_LIB_DIR=kuafu_b200/lib_NAME python ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p kuafu_b200/lib_$name
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  -shared "$@" -o kuafu_b200/lib_$name/libkfrt.so ${SRC:-kuafu_b200/csrc/kf_rt.cu}
cp kuafu_b200/lib/libkuafu.so kuafu_b200/lib_$name/
