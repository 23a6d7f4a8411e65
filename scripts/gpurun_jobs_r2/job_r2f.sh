set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_programs.py tests/test_gpu_kernels.py -q --timeout 200 2>&1 | grep -v " err " | tail -8
timeout 200 python scripts/concurrent_eval.py 16384 24 1 2 3 > gpurun_out/concurrent_r2f.log 2>&1; cat gpurun_out/concurrent_r2f.log
timeout 200 python scripts/quick_eval.py 16384 6 base 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cov_tile_kernel -c 2 -o gpurun_out/prof_covtile_r2f -f python scripts/one_eval.py 16384 1 > gpurun_out/ncu_covtile_r2f.log 2>&1; tail -1 gpurun_out/ncu_covtile_r2f.log
