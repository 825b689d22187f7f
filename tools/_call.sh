mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused or packed or edge or error" 2>&1 | tail -15 ) > gpurun_out/c5_pytest.log
cat gpurun_out/c5_pytest.log
for sh in 11 12; do
GR_FUSED_SHIFT=$sh timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c5_bench_sh$sh.json 2> gpurun_out/c5_bench_sh$sh.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c5_bench_sh$sh.json').read().strip().split('\n')[-1])
print('sh=$sh step %.2f ms e2e %.2f peaks %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['peaks']), d['stage_ms_per_step'])
PY
done
