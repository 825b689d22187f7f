#!/bin/bash
# 2-GPU session, final code of the round: headline workload and ATAC -q on two ranks
set +e
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export GR_BENCH_CACHE=/tmp/grcache
for w in hg38_chip_50M_50M hg38_atac_100M_q; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 --workload $w > $O/c10_bench2_$w.json 2> $O/c10_bench2_$w.err
  echo "bench N=2 $w rc=$?"
done
