#!/bin/bash
# 1-GPU session, final code of the round: GPU tests, the default bench command line, the two -q workloads
set +e
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export GR_BENCH_CACHE=/tmp/grcache
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dist.py > $O/c8_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/c8_pytest.log
SECONDS=0
timeout 600 python bench.py > $O/c8_bench_hg38_chip_50M_50M.json 2> $O/c8_bench_hg38_chip_50M_50M.err
echo "bench chip (default command line) rc=$? ${SECONDS}s"
for w in hg38_atac_100M_q g10_shard_125M_q hg38_fisher3 mini; do
  timeout 300 python bench.py --workload $w --steps 5 --no-cpu-baseline --no-dense > $O/c8_bench_$w.json 2> $O/c8_bench_$w.err
  echo "bench $w rc=$?"
done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/c8_smoke.log 2>&1; echo "smoke rc=$?"
ls -la $O | grep c8_ | head
