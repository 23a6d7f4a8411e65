set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?"
grep -v " err " gpurun_out/pytest_r2b.log | tail -40
python bench.py --steps 5 --warmup 3 --no-extra --no-cpu > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2b.json')); print(d['value'], d['stage_ms_per_eval'], d['der_eval_ms'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2b_n2.json 2> gpurun_out/bench_r2b_n2.err; echo "bench2 rc=$?"; cat gpurun_out/bench_r2b_n2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], json.dumps(d.get('sharded'), indent=1))"; tail -5 gpurun_out/bench_r2b_n2.err
