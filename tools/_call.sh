timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py -q -x 2>&1 | tail -4 > gpurun_out/t45.log; cat gpurun_out/t45.log
n=2
GR_DIST_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench45_${n}gpu.json 2> gpurun_out/bench45_$n.err
grep "rank 0 host-side" gpurun_out/bench45_$n.err | head -1
python -c "
import json
d=json.loads(open('gpurun_out/bench45_${n}gpu.json').read().strip().split('\n')[-1])
print('${n}gpu: step %.2f ms  value %.1f e2e %.2f ms  peaks %d' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['peaks']), d['stage_ms_per_step'])"
