#!/bin/bash
# Build libhirest_b200.so (sm_100a only) in-tree: hirest_b200/libhirest_b200.so
set -e
cd "$(dirname "$0")"
SRC="hirest_b200/csrc/hb_gemm.cu hirest_b200/csrc/hb_attn.cu hirest_b200/csrc/hb_attn2.cu hirest_b200/csrc/hb_attn_small.cu hirest_b200/csrc/hb_elem.cu hirest_b200/csrc/hb_moment.cu hirest_b200/csrc/hb_api.cu"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  -o hirest_b200/libhirest_b200.so $SRC "$@"
