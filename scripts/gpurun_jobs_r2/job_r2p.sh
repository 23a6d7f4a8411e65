set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_r2p.log 2>&1; echo "pytest rc=$?"
grep -v " err " gpurun_out/pytest_r2p.log | tail -5
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2p.json')); print(d['value'], d['e2e']['value'], d['clocks'], d['stage_ms_per_eval'], d['concurrent_streams'], d['cpu_baseline'])"
