set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_programs.py tests/test_gpu_multi.py -q --timeout 300 -k "varying or ragged or routed" 2>&1 | tail -12
