#!/bin/bash
# 2-GPU session: the whole GPU suite (incl. the NCCL parity tests and genrich-b200 --gpus 2), the 1-GPU workloads
# with every scan form timed, the N = 2 lines (BH histogram all-gather on the path with -q)
set +e
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export GR_BENCH_CACHE=/tmp/grcache
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > $O/c3_gpu.txt; nproc >> $O/c3_gpu.txt; nvidia-smi topo -m >> $O/c3_gpu.txt 2>&1
timeout 300 python bench.py --workload mini --steps 3 --no-cpu-baseline > $O/c3_bench_mini.json 2> $O/c3_bench_mini.err
echo "mini rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > $O/c3_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/c3_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/c3_bench_hg38_chip_50M_50M.json 2> $O/c3_bench_hg38_chip_50M_50M.err
echo "bench chip rc=$?"
for w in hg38_atac_100M_q hg38_fisher3 g10_shard_125M_q; do
  timeout 600 python bench.py --workload $w --steps 5 --no-cpu-baseline > $O/c3_bench_$w.json 2> $O/c3_bench_$w.err
  echo "bench $w rc=$?"
done
for w in hg38_chip_50M_50M hg38_atac_100M_q mini; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 --workload $w > $O/c3_bench2_$w.json 2> $O/c3_bench2_$w.err
  echo "bench N=2 $w rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c3_launches_mini.csv \
    python bench.py --profile --workload mini > /dev/null 2> $O/c3_launches_mini.err
cp $O/profile_meta.json $O/c3_meta_mini.json
ls -la $O | grep c3_ | head -40
