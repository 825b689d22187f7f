#!/bin/bash
# first GPU session of round 2 (1 GPU): functional check, GPU tests, the four 1-GPU workloads, ncu passes
set +e
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export GR_BENCH_CACHE=/tmp/grcache
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/c1_gpu.txt; nproc >> $O/c1_gpu.txt; free -g >> $O/c1_gpu.txt
timeout 600 python bench.py --workload mini --steps 3 > $O/c1_bench_mini.json 2> $O/c1_bench_mini.err
echo "mini rc=$?"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/c1_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/c1_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/c1_bench_hg38_chip_50M_50M.json 2> $O/c1_bench_hg38_chip_50M_50M.err
echo "bench chip rc=$?"
for w in hg38_atac_100M_q hg38_fisher3 g10_shard_125M_q; do
  timeout 900 python bench.py --workload $w --steps 5 > $O/c1_bench_$w.json 2> $O/c1_bench_$w.err
  echo "bench $w rc=$?"
done
for w in hg38_chip_50M_50M hg38_atac_100M_q hg38_fisher3 g10_shard_125M_q; do
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      --csv --log-file $O/c1_dram_$w.csv python bench.py --profile --workload $w > /dev/null 2> $O/c1_dram_$w.err
  echo "ncu dram $w rc=$?"
  cp $O/profile_meta.json $O/c1_meta_$w.json
done
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'k_fr_scan|k_fb_move|k_fb_count|k_union_emit|k_ctrl_clamp|k_scan_place|k_union_rank|k_rle_moment|k_pair_insert|k_peak_events' \
    --launch-skip 45 --launch-count 15 -f -o $O/c1_full python bench.py --profile --workload hg38_chip_50M_50M > /dev/null 2> $O/c1_full.err
echo "ncu full rc=$?"
ls -la $O | tail -30
