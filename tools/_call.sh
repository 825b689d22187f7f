mkdir -p gpurun_out
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
tail -c 4000 gpurun_out/c7_bench.json; tail -3 gpurun_out/c7_bench.err
