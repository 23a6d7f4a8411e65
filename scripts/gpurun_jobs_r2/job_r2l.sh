set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py tests/test_gpu_multi.py -q --timeout 200 2>&1 | grep -v " err " | tail -4
for b in 0 1; do GPK_DIST_SPLIT=$b timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$b scripts/bench_dist.py 65536 32 1 2>/dev/null | tail -1 | sed "s/^/SPLIT=$b /"; done
GPK_DIST_SPLIT=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 scripts/bench_dist.py 16384 8 2 2>/dev/null | tail -1 | sed "s/^/N16384 SPLIT=1 /"
