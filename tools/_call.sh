mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dist.py -q -x 2>&1 | tail -6 ) > gpurun_out/c10_pytest.log
cat gpurun_out/c10_pytest.log
n=2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench_${n}gpu.json 2> gpurun_out/c10_bench_$n.err
python -c "
import json
d=json.loads(open('gpurun_out/c10_bench_${n}gpu.json').read().strip().split('\n')[-1])
print('${n}gpu: step %.2f ms  value %.1f e2e %.2f ms  peaks %d' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['peaks']), d['stage_ms_per_step'])"
tail -3 gpurun_out/c10_bench_$n.err
