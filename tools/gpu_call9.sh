#!/bin/bash
# 2-GPU check of the histogram exchange with reused buffers (ATAC -q on two ranks; the dist tests)
set +e
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export GR_BENCH_CACHE=/tmp/grcache
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --workload hg38_atac_100M_q > $O/c12_bench2_hg38_atac_100M_q.json 2> $O/c12_bench2_hg38_atac_100M_q.err
echo "bench N=2 atac rc=$?"
timeout 300 python -m pytest tests/test_gpu_dist.py -q > $O/c12_pytest_dist.log 2>&1
echo "pytest dist rc=$?"; tail -3 $O/c12_pytest_dist.log
