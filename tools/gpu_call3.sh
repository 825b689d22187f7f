#!/bin/bash
# 1-GPU session: the five workloads with the device-chosen scan form (every form timed beside it), GPU tests,
# DRAM bytes per kernel of one step of each workload
set +e
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export GR_BENCH_CACHE=/tmp/grcache
timeout 300 python bench.py --workload mini --steps 3 --no-cpu-baseline > $O/c5_bench_mini.json 2> $O/c5_bench_mini.err
echo "mini rc=$?"
SECONDS=0
timeout 600 python bench.py > $O/c5_bench_hg38_chip_50M_50M.json 2> $O/c5_bench_hg38_chip_50M_50M.err
echo "bench chip (default command line) rc=$? ${SECONDS}s"
for w in hg38_atac_100M_q g10_shard_125M_q hg38_fisher3; do
  GR_BENCH_FD=1 timeout 400 python bench.py --workload $w --steps 5 --no-cpu-baseline > $O/c5_bench_$w.json 2> $O/c5_bench_$w.err
  echo "bench $w rc=$?"
done
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dist.py > $O/c5_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/c5_pytest.log
for w in hg38_chip_50M_50M hg38_atac_100M_q hg38_fisher3 g10_shard_125M_q; do
  timeout 240 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      --csv --log-file $O/c5_dram_$w.csv python bench.py --profile --workload $w > /dev/null 2> $O/c5_dram_$w.err
  echo "ncu dram $w rc=$?"
  cp $O/profile_meta.json $O/c5_meta_$w.json
done
ls -la $O | grep c5_ | head -40
