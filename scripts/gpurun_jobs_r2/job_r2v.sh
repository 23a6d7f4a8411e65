set -x
mkdir -p gpurun_out
timeout 100 python scripts/chain_dump.py 16384 > gpurun_out/chain_r2v.log 2>&1; tail -3 gpurun_out/chain_r2v.log
timeout 60 python scripts/diag_clk.py 2>&1 | tail -3
