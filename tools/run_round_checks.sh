set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.txt
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
timeout 600 python bench.py --config moment > gpurun_out/r02_moment_1gpu.json 2> gpurun_out/r02_moment_1gpu.err
timeout 600 python bench.py --config e2e > gpurun_out/r02_e2e_1gpu.json 2> gpurun_out/r02_e2e_1gpu.err
timeout 300 python tools/profile_moment.py > gpurun_out/r02_profile_moment.json 2>&1
tail -3 gpurun_out/r02_gpu_tests.txt; tail -2 gpurun_out/r02_smoke.txt; cut -c1-400 gpurun_out/r02_bench_1gpu.json; cut -c1-300 gpurun_out/r02_moment_1gpu.json; cut -c1-300 gpurun_out/r02_e2e_1gpu.json
