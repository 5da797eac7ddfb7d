#!/bin/bash
# Build the standalone GPU probes (no torch dependency). Output: tools/_build/ (git-ignored, travels with gpurun).
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_build
NVCC_FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17"
nvcc $NVCC_FLAGS -o tools/_build/gemm_selftest tools/gemm_selftest.cu hirest_b200/csrc/hb_gemm.cu
nvcc $NVCC_FLAGS -DHB_ATTN_TIMING -o tools/_build/attn3_timing tools/attn3_timing.cu hirest_b200/csrc/hb_attn3.cu hirest_b200/csrc/hb_gemm.cu
# host-only emulation of the resize kernel v2 (tables + arithmetic vs the two-pass evaluation of the v1 tables); needs ./build.sh first
nvcc -std=c++17 -O1 -Wno-deprecated-gpu-targets -o tools/_build/preproc_host_check tools/preproc_host_check.cu build/hb_preproc.o
