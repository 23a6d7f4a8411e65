set -x
mkdir -p gpurun_out
timeout 300 python scripts/sweep_potrf.py 16384 3,9,48 1,9,48 1,12,48 2,12,48 2,10,48 3,12,48 3,12,40 4,12,48 > gpurun_out/sweep_r2j.log 2>&1; cat gpurun_out/sweep_r2j.log
for m in 8 12 16 24; do GPK_OZAKI_MIN=$m timeout 100 python scripts/quick_eval.py 16384 6 "ozmin$m" 2>&1 | tail -1; done
for w in 2 3 4 6; do GPK_POTRF_W2B=$w timeout 100 python scripts/quick_eval.py 16384 6 "w2b$w" 2>&1 | tail -1; done
