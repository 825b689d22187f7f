mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_cli.py -q -x 2>&1 | tail -6 ) > gpurun_out/c19_pytest.log
cat gpurun_out/c19_pytest.log
