set -x
i=0
for cfg in "1 16" "0 16" "1 8"; do set -- $cfg; i=$((i+1)); GPK_DIST_SPLIT=$1 GPK_DIST_WD=$2 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2958$i scripts/bench_dist.py 65536 32 2 2>/dev/null | tail -1 | sed "s/^/SPLIT=$1 WD=$2 /"; done
