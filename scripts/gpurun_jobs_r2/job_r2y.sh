set -x
mkdir -p gpurun_out
GPK_DIAG_OVL=2 GPK_SMALL_BREG=1 timeout 200 python -m pytest tests/test_gpu_kernels.py -q --timeout 100 -k "diag or chain or potrf or factor" 2>&1 | grep -v " err " | tail -4
GPK_DIAG_OVL=2 timeout 60 python scripts/diag_clk.py 2>&1 | tail -1
GPK_DIAG_OVL=1 timeout 60 python scripts/diag_clk.py 2>&1 | tail -1
for v in "1 1" "3 1" "4 1" "1 2" "3 2" "4 2" "1 1" "3 2"; do
  set -- $v
  GPK_POTRF_HEADK=$1 GPK_DIAG_OVL=$2 timeout 100 python scripts/quick_eval.py 16384 8 "headk$1-ovl$2" | tail -1
done
GPK_POTRF_HEADK=3 GPK_DIAG_OVL=2 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -q --timeout 200 2>&1 | grep -v " err " | tail -4
