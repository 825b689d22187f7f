#!/bin/bash
# First GPU call of round 2 (one B200):  gpurun --timeout 1500 -- 'bash tools/r02_first_call.sh'
# 1. the GPU test-suite (default path gates; the knob-gated variants report XPASS / xfail)
# 2. bench.py at N = 1 (its "variants" key times every variant in a child process)
# 3. launch lists (ncu, serialised: shares only) of the default path and of the all-knobs path
# 4. one `ncu --set full` capture each of the new kernels
# Everything lands in gpurun_out/r02_*; summarise into profiles/ with tools/launch_summary.py / ncu_summary.py.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02_pytest.log
cat gpurun_out/r02_pytest.log
( timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02_bench.err | tail -1 ) > gpurun_out/r02_bench_1gpu.json
cat gpurun_out/r02_bench_1gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d.get('variants'), indent=1)[:6000])"
ALL="GR_FUSED_RANK=1 GR_FB_P2=1 GR_UE_WARP=1 GR_UR_GROUPS=4 GR_CL_TILES=4"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense --no-variants > gpurun_out/r02_ncu_default.log 2>&1
env $ALL timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_all_p2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense --no-variants > gpurun_out/r02_ncu_all_p2.log 2>&1
for k in k_fr_scan k_p1_move k_p2 k_union_emit_w; do
  env $ALL timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 2 -o gpurun_out/r02_$k -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dense --no-variants > gpurun_out/r02_ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -20
