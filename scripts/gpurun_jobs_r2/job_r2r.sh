set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r2r_n8.json 2> gpurun_out/bench_r2r_n8.err; echo "bench8 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2r_n8.json')); s=d['sharded']; print(d['value'], d['e2e']['value'], s['c3']['ms_per_eval'], s['c3']['rel_err_vs_single_gpu'], s['sharded_der_predict']['der_eval_ms'], s['c4']['ms_per_eval'])"; tail -3 gpurun_out/bench_r2r_n8.err
