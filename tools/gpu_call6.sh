#!/bin/bash
# 8-GPU session, final code of the round: headline workload, the 10 Gbp / 1 B fragment configuration, ATAC -q
set +e
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export GR_BENCH_CACHE=/tmp/grcache
run() {  # N workload steps
  SECONDS=0
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $1 --steps $3 --warmup 3 --workload $2 > $O/c9_bench$1_$2.json 2> $O/c9_bench$1_$2.err
  echo "bench N=$1 $2 rc=$? ${SECONDS}s"
}
run 8 hg38_chip_50M_50M 10
run 8 hg38_atac_100M_q 5
run 8 g10_multimap_1B_q 5
ls -la $O | grep c9_ | head
