mkdir -p gpurun_out
( timeout 700 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/c1_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c1_ncu_bench.log 2>&1
cat gpurun_out/c1_pytest.log; tail -c 3000 gpurun_out/c1_bench.json
