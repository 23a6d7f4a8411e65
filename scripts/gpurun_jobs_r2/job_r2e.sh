set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_r2e.log 2>&1; echo "pytest rc=$?"
grep -v " err " gpurun_out/pytest_r2e.log | tail -15
timeout 300 python scripts/sweep_potrf.py 16384 3,9,48 3,12,48 3,6,48 2,8,48 4,12,48 4,8,48 3,9,32 3,9,64 3,15,48 2,10,48 > gpurun_out/sweep_r2e.log 2>&1; cat gpurun_out/sweep_r2e.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cov_tile_kernel -c 2 -o gpurun_out/prof_covtile_r2e -f python scripts/one_eval.py 16384 1 > gpurun_out/ncu_covtile_r2e.log 2>&1; tail -2 gpurun_out/ncu_covtile_r2e.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2e.csv python scripts/one_eval.py 16384 3 > gpurun_out/ncu_launches_r2e.log 2>&1; tail -2 gpurun_out/ncu_launches_r2e.log
