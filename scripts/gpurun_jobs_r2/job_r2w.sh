set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -q --timeout 100 -x 2>&1 | grep -v " err " | tail -8
timeout 60 python scripts/diag_clk.py 2>&1 | tail -2
GPK_DIAG_OVL=0 timeout 60 python scripts/diag_clk.py 2>&1 | tail -1
for v in "0 0" "1 0" "0 1" "1 1" "2 1" "0 0" "1 1"; do
  set -- $v
  GPK_POTRF_HEADK=$1 GPK_DIAG_OVL=$2 timeout 100 python scripts/quick_eval.py 16384 8 "headk$1-ovl$2" | tail -1
done
GPK_POTRF_HEADK=1 GPK_DIAG_OVL=1 timeout 100 python scripts/quick_eval.py 8192 8 "headk1-ovl1" | tail -1
GPK_POTRF_HEADK=0 GPK_DIAG_OVL=0 timeout 100 python scripts/quick_eval.py 8192 8 "headk0-ovl0" | tail -1
