#!/bin/bash
# 1-GPU session, last of the round: GPU tests and smoke on the final tree; ncu --set full of the largest kernels of
# the two -q workloads (the kernels DESIGN.md section 7 names as next)
set +e
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
export GR_BENCH_CACHE=/tmp/grcache
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dist.py > $O/c11_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/c11_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/c11_smoke.log 2>&1; echo "smoke rc=$?"
for w in hg38_atac_100M_q g10_shard_125M_q; do
  timeout 300 ncu --set full --clock-control none --import-source on \
      -k regex:'k_slot_hist|k_fb_scan|k_fb_move|k_fb_count|k_union_emit' --launch-skip 15 --launch-count 5 \
      -f -o $O/c11_full_$w python bench.py --profile --workload $w > /dev/null 2> $O/c11_full_$w.err
  echo "ncu full $w rc=$?"
done
ls -la $O | grep c11_
