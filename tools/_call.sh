timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -8 > gpurun_out/t26.log; cat gpurun_out/t26.log
for cfg in "2 2" "3 2" "2 3"; do set -- $cfg
GR_SCATTER_BIN=0 GR_SCAN_STAGES=$1 GR_SCAN_CPS=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench26_$1_$2.json 2>> gpurun_out/bench26.err
python -c "
import json
d=json.load(open('gpurun_out/bench26_$1_$2.json'))
print('stages $1 cps $2: step %.2f ms  e2e %.2f  scan %.3f ms/launch frac %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac']))"
done
tail -3 gpurun_out/bench26.err
GR_SCAN_CPS=3 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
