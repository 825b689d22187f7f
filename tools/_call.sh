mkdir -p gpurun_out
( timeout 700 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/c3_pytest.log
cat gpurun_out/c3_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fb_scan|k_fb_move|k_fb_count" -s 8 -c 3 -o gpurun_out/c3_fb -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c3_ncu.log 2>&1
tail -3 gpurun_out/c3_ncu.log
