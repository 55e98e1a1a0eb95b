set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r8_pytest.log
python tools/bench_kernel.py > gpurun_out/r8_kern.json 2> gpurun_out/r8_kern.err
RDP_LIB_PATH=$PWD/radardistill_b200/librdp_u6.so python tools/bench_kernel.py > gpurun_out/r8_kern_u6.json 2>> gpurun_out/r8_kern.err
RDP_LIB_PATH=$PWD/radardistill_b200/librdp_u2.so python tools/bench_kernel.py > gpurun_out/r8_kern_u2.json 2>> gpurun_out/r8_kern.err
tail -3 gpurun_out/r8_pytest.log; cat gpurun_out/r8_kern.json gpurun_out/r8_kern_u6.json gpurun_out/r8_kern_u2.json
