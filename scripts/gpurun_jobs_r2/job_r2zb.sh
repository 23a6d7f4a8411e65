set -x
mkdir -p gpurun_out
for v in 0 1 0 1; do
  GPK_DIAG_SKIPZ=$v timeout 60 python scripts/quick_eval.py 16384 10 "skipz$v" | tail -1
done
for v in 0 1; do GPK_DIAG_SKIPZ=$v timeout 30 python scripts/diag_clk.py 2>&1 | tail -1; done
GPK_DIAG_SKIPZ=1 timeout 120 python -m pytest tests -m gpu -q --timeout 100 -x > gpurun_out/pytest_r2zb.log 2>&1; echo "pytest rc=$?"
grep -v " err " gpurun_out/pytest_r2zb.log | tail -5
