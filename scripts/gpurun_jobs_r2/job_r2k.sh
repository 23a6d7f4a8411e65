set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_r2k.log 2>&1; echo "pytest rc=$?"
grep -v " err " gpurun_out/pytest_r2k.log | tail -6
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2k.json')); print(d['value'], d['e2e']['value'], d['clocks'], d['stage_ms_per_eval'], d['concurrent_streams'], d['cpu_baseline']['value'])"
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_ref_r2k.json 2> gpurun_out/bench_ref_r2k.err; cat gpurun_out/bench_ref_r2k.json | cut -c1-500; tail -4 gpurun_out/bench_ref_r2k.err
for s in 6 7; do GPK_OZAKI_SLICES=$s timeout 100 python scripts/quick_eval.py 16384 6 "S$s" 2>&1 | tail -1; done
