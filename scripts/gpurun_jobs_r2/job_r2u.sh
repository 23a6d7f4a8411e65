set -x
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_gpu_ep.py -q --timeout 300 2>&1 | grep -v " err " | tail -4
