#!/bin/bash
# Build libhirest_b200.so (sm_100a only) in-tree: hirest_b200/libhirest_b200.so
# Sources compile in parallel into build/ (git-ignored); extra arguments go to every nvcc compile.
set -e
cd "$(dirname "$0")"
SRC="hb_gemm hb_attn hb_attn2 hb_attn3 hb_attn_tc hb_attn_small hb_elem hb_moment hb_preproc hb_tokenize hb_api"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
OUT="${HB_OUT:-hirest_b200/libhirest_b200.so}"
BDIR="${HB_BUILD_DIR:-build}"
mkdir -p "$BDIR"
pids=""
for f in $SRC; do
  nvcc $FLAGS -c hirest_b200/csrc/$f.cu -o $BDIR/$f.o "$@" &
  pids="$pids $!"
done
for p in $pids; do wait $p; done
OBJS=""
for f in $SRC; do OBJS="$OBJS $BDIR/$f.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $OBJS
