set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r11_pytest.log
python tools/bench_kernel.py > gpurun_out/r11_kern.json 2> gpurun_out/r11_kern.err
RDP_PFN_ROWS=1 python tools/bench_kernel.py > gpurun_out/r11_kern_rows.json 2>> gpurun_out/r11_kern.err
RDP_PFN_ROWS=1 RDP_LIB_PATH=$PWD/radardistill_b200/librdp_r1.so python tools/bench_kernel.py > gpurun_out/r11_kern_r1.json 2>> gpurun_out/r11_kern.err
RDP_PFN_ROWS=1 RDP_LIB_PATH=$PWD/radardistill_b200/librdp_r1g6.so python tools/bench_kernel.py > gpurun_out/r11_kern_r1g6.json 2>> gpurun_out/r11_kern.err
RDP_PFN_ROWS=1 RDP_LIB_PATH=$PWD/radardistill_b200/librdp_r1.so python -m pytest tests -m gpu -x -q > gpurun_out/r11_pytest_r1.log 2>&1
RDP_NO_FLUSH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pfn_tile" --launch-skip 56 -c 2 -o gpurun_out/r11_prof_bwd -f python tools/bench_kernel.py > gpurun_out/r11_ncu.log 2>&1
tail -3 gpurun_out/r11_pytest.log gpurun_out/r11_pytest_r1.log; cat gpurun_out/r11_kern*.json
