mkdir -p gpurun_out
GR_BENCH_TRACE=1 GR_GAP_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dense > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
grep -n "host +\|gr_call_peaks\|host ms" gpurun_out/c14_bench.err | sed -n '150,260p'
