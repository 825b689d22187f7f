mkdir -p gpurun_out
( timeout 700 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 ) > gpurun_out/c18_pytest.log
cat gpurun_out/c18_pytest.log
timeout 400 python bench.py > gpurun_out/c18_bench.json 2> gpurun_out/c18_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c18_bench.json').read().strip().split('\n')[-1])
print('step %.2f ms e2e %.2f (%s) peaks %d launches %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e'].get('record_format'), d['config']['peaks'], d['gpu_launches']), d['stage_ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])
PY
tail -3 gpurun_out/c18_bench.err
timeout 300 python bench.py --no-cpu-baseline --no-dense --prefetch-depth 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('depth1: step %.2f e2e %.2f' % (d['ms_per_step'], d['e2e']['ms_per_step']))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file gpurun_out/c18_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense > gpurun_out/c18_ncu_bench.log 2>&1
tail -1 gpurun_out/c18_ncu_bench.log | cut -c1-200
