mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/c11_pytest.log
cat gpurun_out/c11_pytest.log
timeout 400 python bench.py > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c11_bench.json').read().strip().split('\n')[-1])
print('step %.2f ms e2e %.2f peaks %d launches %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['peaks'], d['gpu_launches']), d['stage_ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'])
PY
tail -3 gpurun_out/c11_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/c11_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense > gpurun_out/c11_ncu_bench.log 2>&1
tail -2 gpurun_out/c11_ncu_bench.log | cut -c1-300
