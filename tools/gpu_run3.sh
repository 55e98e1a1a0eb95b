set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4_pytest.log
python tools/bench_kernel.py > gpurun_out/r4_kern.json 2> gpurun_out/r4_kern.err
RDP_LIB_PATH=$PWD/radardistill_b200/librdp_g4.so python tools/bench_kernel.py > gpurun_out/r4_kern_g4.json 2>> gpurun_out/r4_kern.err
RDP_NO_FLUSH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pfn_rows" -c 3 -o gpurun_out/r4_prof -f python tools/bench_kernel.py > gpurun_out/r4_ncu.log 2>&1
tail -15 gpurun_out/r4_pytest.log; cat gpurun_out/r4_kern.json gpurun_out/r4_kern_g4.json
