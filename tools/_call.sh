timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -8 > gpurun_out/t24.log; cat gpurun_out/t24.log
for st in 3 2; do
GR_SCATTER_BIN=0 GR_SCAN_STAGES=$st timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench24_$st.json 2>> gpurun_out/bench24.err
python -c "
import json
d=json.load(open('gpurun_out/bench24_$st.json'))
print('stages $st: step %.2f ms  e2e %.2f  scan %.3f ms/launch frac %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac']), d['stage_ms_per_step'])"
done
tail -3 gpurun_out/bench24.err
