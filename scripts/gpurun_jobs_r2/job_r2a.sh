set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dist.py > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$?"
tail -60 gpurun_out/pytest_r2a.log
python __graft_entry__.py --smoke > gpurun_out/smoke_r2a.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r2a.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; cat gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
python bench.py --impl reference --problem-n 4096 --steps 1 > gpurun_out/bench_ref_r2a.json 2>&1; cat gpurun_out/bench_ref_r2a.json
ncu --set full --clock-control none --import-source on -k regex:cov_kernel -c 2 -o gpurun_out/prof_cov_r2a -f python scripts/one_eval.py 16384 1 > gpurun_out/ncu_cov_r2a.log 2>&1; tail -3 gpurun_out/ncu_cov_r2a.log
GPK_PROFILE_DUMP=1 python scripts/step_profile.py > gpurun_out/step_profile_r2a.log 2>&1; tail -25 gpurun_out/step_profile_r2a.log
