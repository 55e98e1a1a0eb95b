set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r9_pytest.log
python tools/bench_kernel.py > gpurun_out/r9_kern.json 2> gpurun_out/r9_kern.err
RDP_PFN_LEGACY=1 python tools/bench_kernel.py > gpurun_out/r9_kern_legacy.json 2>> gpurun_out/r9_kern.err
RDP_PFN_ROWS=1 python tools/bench_kernel.py > gpurun_out/r9_kern_rows.json 2>> gpurun_out/r9_kern.err
RDP_LIB_PATH=$PWD/radardistill_b200/librdp_u6.so python tools/bench_kernel.py > gpurun_out/r9_kern_u6.json 2>> gpurun_out/r9_kern.err
tail -3 gpurun_out/r9_pytest.log; cat gpurun_out/r9_kern.json gpurun_out/r9_kern_legacy.json gpurun_out/r9_kern_rows.json gpurun_out/r9_kern_u6.json
