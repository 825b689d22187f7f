timeout 900 python -m pytest tests/test_gpu_dist.py -q 2>&1 | tail -40 > gpurun_out/t43.log; grep -v "^$" gpurun_out/t43.log | tail -40
