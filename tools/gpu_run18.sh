set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r18_pytest.log
python tools/bench_kernel.py > gpurun_out/r18_kern.json 2> gpurun_out/r18_kern.err
RDP_LIB_PATH=$PWD/radardistill_b200/librdp_k1d.so python tools/bench_kernel.py > gpurun_out/r18_kern_k1d.json 2>> gpurun_out/r18_kern.err
RDP_LIB_PATH=$PWD/radardistill_b200/librdp_k1d.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r18_pytest_k1d.log 2>&1
tail -n 3 gpurun_out/r18_pytest.log gpurun_out/r18_pytest_k1d.log; cat gpurun_out/r18_kern.json gpurun_out/r18_kern_k1d.json
