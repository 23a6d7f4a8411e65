set -x
timeout 200 python -m pytest tests/test_gpu_dist.py -q --timeout 150 2>&1 | tail -2
GPK_DIST_SPLIT=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 scripts/bench_dist.py 65536 32 1 2>/dev/null | tail -1 | sed "s/^/SPLIT=1 /"
