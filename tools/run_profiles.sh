#!/bin/bash
# Round-end evidence run on one B200 (from the repo root): new decoder test, ncu launch list of the bench command, one-layer
# ncu --set full, sustained GEMM vs cuBLAS.  Outputs under gpurun_out/ (copied into profiles/ by hand).
set -x
timeout 600 python -m pytest tests/test_gpu_moment.py -m gpu -x -q 2>&1 | tail -3
timeout 1500 bash tools/profile_layer.sh
python tools/ncu_launch_agg.py gpurun_out/launches.csv > gpurun_out/launch_list_summary.txt
head -20 gpurun_out/launch_list_summary.txt
head -40 gpurun_out/layer_summary.txt | grep -E "kernel|duration|tensor_cycles|dram_throughput"
