set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q --timeout 150 > gpurun_out/pytest_r2c_multi.log 2>&1; echo "multi rc=$?"
grep -v " err " gpurun_out/pytest_r2c_multi.log | tail -30
timeout 600 python -m pytest tests/test_gpu_programs.py tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_edge.py -q --timeout 300 > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?"
grep -v " err " gpurun_out/pytest_r2c.log | tail -60
timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2c.json')); print(d['value'], d['stage_ms_per_eval'], d['der_eval_ms'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2c_n2.json 2> gpurun_out/bench_r2c_n2.err; echo "bench2 rc=$?"; cat gpurun_out/bench_r2c_n2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], json.dumps(d.get('sharded'), indent=1))"; tail -5 gpurun_out/bench_r2c_n2.err
