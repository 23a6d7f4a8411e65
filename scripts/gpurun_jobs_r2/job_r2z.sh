set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_r2z.log 2>&1; echo "pytest rc=$?"
grep -v " err " gpurun_out/pytest_r2z.log | tail -6
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2z.json')); print(d['value'], d['e2e']['value'], d['clocks'], d['stage_ms_per_eval'], d['roofline']['frac'], d['roofline']['int8_issue_peak']['frac'], d['parity'])"
timeout 100 python scripts/chain_dump.py 16384 > gpurun_out/chain_r2z.log 2>&1; tail -1 gpurun_out/chain_r2z.log
for o in 0 1 2; do GPK_DIAG_OVL=$o timeout 60 python scripts/diag_clk.py 2>&1 | tail -1; done > gpurun_out/diag_clk_r2z.log 2>&1; cat gpurun_out/diag_clk_r2z.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:potrf_diag_ovl -s 100 -c 1 -o gpurun_out/prof_diag_ovl_r2 -f python scripts/one_eval.py 16384 1 > gpurun_out/ncu_diag_r2z.log 2>&1; tail -1 gpurun_out/ncu_diag_r2z.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:small_nt -s 8 -c 2 -o gpurun_out/prof_small_nt_r2 -f python scripts/one_eval.py 16384 1 > gpurun_out/ncu_small_r2z.log 2>&1; tail -1 gpurun_out/ncu_small_r2z.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2z.csv python scripts/one_eval.py 16384 3 > gpurun_out/ncu_launches_r2z.log 2>&1; tail -1 gpurun_out/ncu_launches_r2z.log
