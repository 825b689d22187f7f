#!/bin/bash
# 8-GPU session (second run; the first one also ran the host program on 8 devices: c6_pytest_cli.log): BASELINE configs 2-5 and the headline workload on 8 ranks (one process per GPU, NCCL), the host
# program on 8 devices, the headline workload on 4 ranks
set +e
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export GR_BENCH_CACHE=/tmp/grcache
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $O/c7_gpu.txt; nproc >> $O/c7_gpu.txt; nvidia-smi topo -m >> $O/c7_gpu.txt 2>&1
run() {  # N workload steps
  SECONDS=0
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $1 --steps $3 --warmup 3 --workload $2 > $O/c7_bench$1_$2.json 2> $O/c7_bench$1_$2.err
  echo "bench N=$1 $2 rc=$? ${SECONDS}s"
}
run 8 hg38_chip_50M_50M 10
run 8 g10_multimap_1B_q 5
run 8 hg38_fisher3 5
run 8 hg38_atac_100M_q 5
run 4 hg38_chip_50M_50M 10
ls -la $O | grep c7_ | head -40
