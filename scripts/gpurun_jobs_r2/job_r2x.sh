set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -q --timeout 100 -x -k "diag or chain or potrf or factor" 2>&1 | grep -v " err " | tail -4
timeout 60 python scripts/diag_clk.py 2>&1 | tail -1
timeout 100 python scripts/chain_dump.py 16384 > gpurun_out/chain_r2x.log 2>&1; tail -1 gpurun_out/chain_r2x.log
timeout 100 python scripts/quick_eval.py 16384 8 "default" | tail -1
for v in "3 48" "6 48" "4 32" "4 40" "4 64" "5 48"; do
  set -- $v
  GPK_POTRF_W2B=$1 GPK_POTRF_W1_MINREM=$2 timeout 100 python scripts/quick_eval.py 16384 8 "w2b$1-minrem$2" | tail -1
done
timeout 100 python scripts/quick_eval.py 16384 8 "default" | tail -1
