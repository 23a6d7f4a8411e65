set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r2h_n8.json 2> gpurun_out/bench_r2h_n8.err; echo "bench8 rc=$?"; cat gpurun_out/bench_r2h_n8.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], json.dumps(d.get('sharded'), indent=1))"; tail -3 gpurun_out/bench_r2h_n8.err
for wd in 4 16; do GPK_DIST_WD=$wd timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 scripts/bench_dist.py 65536 32 2 2>/dev/null | tail -1 | sed "s/^/WD=$wd /"; done
GPK_DIST_OZAKI=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 scripts/bench_dist.py 65536 32 1 2>/dev/null | tail -1 | sed "s/^/DMMA /"
