#!/bin/bash
# 2-GPU session, second try: diagnostics first (a crash in the N = 2 path stops the session early)
set +e
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export GR_BENCH_CACHE=/tmp/grcache
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 --workload mini > $O/c4_bench2_mini.json 2> $O/c4_bench2_mini.err
rc=$?; echo "bench N=2 mini rc=$rc"
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_cli.py -q -k "two_gpus or sharded_over" > $O/c4_pytest_dist.log 2>&1
echo "pytest dist rc=$?"; tail -5 $O/c4_pytest_dist.log
if [ $rc -ne 0 ]; then grep -v "^\[W" $O/c4_bench2_mini.err | head -60; exit 0; fi
timeout 300 python bench.py --workload mini --steps 3 --no-cpu-baseline > $O/c4_bench_mini.json 2> $O/c4_bench_mini.err
echo "mini rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dist.py > $O/c4_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/c4_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/c4_bench_hg38_chip_50M_50M.json 2> $O/c4_bench_hg38_chip_50M_50M.err
echo "bench chip rc=$?"
for w in hg38_atac_100M_q hg38_fisher3 g10_shard_125M_q; do
  timeout 600 python bench.py --workload $w --steps 5 --no-cpu-baseline > $O/c4_bench_$w.json 2> $O/c4_bench_$w.err
  echo "bench $w rc=$?"
done
for w in hg38_chip_50M_50M hg38_atac_100M_q; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 --workload $w > $O/c4_bench2_$w.json 2> $O/c4_bench2_$w.err
  echo "bench N=2 $w rc=$?"
done
GR_BUCKET_ATOMIC=1 timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-dense --no-e2e > $O/c4_bench_chip_atomic_buckets.json 2> $O/c4_bench_chip_atomic_buckets.err
echo "bench chip atomic buckets rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c4_launches_mini.csv \
    python bench.py --profile --workload mini > /dev/null 2> $O/c4_launches_mini.err
cp $O/profile_meta.json $O/c4_meta_mini.json
ls -la $O | grep c4_ | head -40
