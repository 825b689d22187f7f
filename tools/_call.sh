timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/t29.log; cat gpurun_out/t29.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench29.json 2>> gpurun_out/bench29.err
python -c "
import json
d=json.load(open('gpurun_out/bench29.json'))
print('step %.2f ms  e2e %.2f (wall %.2f)  scan %.3f ms/launch frac %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['wall_ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac']), d['stage_ms_per_step'])"
tail -3 gpurun_out/bench29.err
