#!/bin/bash
# Reproduces the ncu evidence under profiles/ on a B200 box (run from the repo root, library already built).
#   1. launch list of a short bench run            -> gpurun_out/launches.csv  (summarised in profiles/r01_launch_list_summary.txt)
#   2. ncu --set full of one ViT layer (5 kernels) -> gpurun_out/layer.ncu-rep (summarised by tools/ncu_summary.py)
#   3. power-capped GEMM throughput vs cuBLAS      -> gpurun_out/sustained.json
# A number printed by a run under ncu is never a bench value.
set -e
mkdir -p gpurun_out
KERNELS='gemm_kernel|vit_attn|layernorm|row_stats|small_attn|im2col|pool_norm|text_embed|split_bf16|cls_row'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERNELS" -c 1400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
# with the filter below the patch-embed GEMM is launch 98 of the run (text tower first), layer 0 starts at 99, 5 launches per layer
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|vit_attn2" -s 119 -c 5 -o gpurun_out/layer -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/layer.ncu-rep "one ViT layer, see profiles/r01_one_layer_ncu_full.txt" > gpurun_out/layer_summary.txt
python tools/sustained_gemm.py > gpurun_out/sustained.json
