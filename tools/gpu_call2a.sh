#!/bin/bash
# 1-GPU session: GPU tests, the five 1-GPU workloads, DRAM passes (short limits: a hung ncu run cost 15 minutes once)
set +e
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export GR_BENCH_CACHE=/tmp/grcache
timeout 300 python bench.py --workload mini --steps 3 --no-cpu-baseline > $O/c2_bench_mini.json 2> $O/c2_bench_mini.err
echo "mini rc=$?"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/c2_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/c2_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/c2_bench_hg38_chip_50M_50M.json 2> $O/c2_bench_hg38_chip_50M_50M.err
echo "bench chip rc=$?"
for w in hg38_atac_100M_q hg38_fisher3 g10_shard_125M_q; do
  timeout 600 python bench.py --workload $w --steps 5 > $O/c2_bench_$w.json 2> $O/c2_bench_$w.err
  echo "bench $w rc=$?"
done
for w in hg38_chip_50M_50M hg38_atac_100M_q hg38_fisher3 g10_shard_125M_q; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      --csv --log-file $O/c2_dram_$w.csv python bench.py --profile --workload $w > /dev/null 2> $O/c2_dram_$w.err
  echo "ncu dram $w rc=$?"
  cp $O/profile_meta.json $O/c2_meta_$w.json
done
ls -la $O | grep c2_ | head -40
