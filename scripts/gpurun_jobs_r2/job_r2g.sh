set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2g_n2.json 2> gpurun_out/bench_r2g_n2.err; echo "bench2 rc=$?"; cat gpurun_out/bench_r2g_n2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], json.dumps(d.get('sharded'), indent=1))"; tail -5 gpurun_out/bench_r2g_n2.err
timeout 300 python -m pytest tests/test_gpu_dist.py -q --timeout 200 2>&1 | tail -3
