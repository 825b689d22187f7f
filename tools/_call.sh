mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/c8_pytest.log
cat gpurun_out/c8_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dense 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('step', d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms_per_step'])"
