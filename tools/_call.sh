mkdir -p gpurun_out
( timeout 45 python -m pytest tests/test_gpu_cli.py -q -x -k "c2_ctrl_q or bed_fisher_q or threaded or host_r_y" 2>&1 | tail -4 ) > gpurun_out/c20_pytest.log
cat gpurun_out/c20_pytest.log
