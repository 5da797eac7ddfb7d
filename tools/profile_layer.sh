#!/bin/bash
# Reproduces the ncu evidence under profiles/ on a B200 box (run from the repo root, library already built).
#   1. launch list of a short bench run            -> gpurun_out/launches.csv  (summarised in profiles/r01_launch_list_summary.txt)
#   2. ncu --set full of one ViT layer (5 kernels) -> gpurun_out/layer.ncu-rep (summarised by tools/ncu_summary.py)
#   3. power-capped GEMM throughput vs cuBLAS      -> gpurun_out/sustained.json
# A number printed by a run under ncu is never a bench value.
set -e
mkdir -p gpurun_out
KERNELS='gemm_kernel|vit_attn|layernorm|row_stats|small_attn|im2col|pool_norm|text_embed|split_bf16|cls_row|split3|attn_split'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERNELS" -c 1400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
# one ViT layer (layer 4) of an encode_image call: the library brackets it with cudaProfilerStart / Stop (HB_DEBUG_PROFILE_LAYER)
HB_DEBUG_PROFILE_LAYER=4 ncu --profile-from-start off --set full --clock-control none --import-source on -c 5 -o gpurun_out/layer -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/layer.ncu-rep --json gpurun_out/gemm_traffic.json "one ViT layer (layer 4), EVA-CLIP-g/14, 1024 frames" > gpurun_out/layer_summary.txt
python tools/sustained_gemm.py > gpurun_out/sustained.json
