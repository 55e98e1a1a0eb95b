set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r7_pytest.log
python tools/bench_kernel.py > gpurun_out/r7_kern.json 2> gpurun_out/r7_kern.err
RDP_LIB_PATH=$PWD/radardistill_b200/librdp_np.so python tools/bench_kernel.py > gpurun_out/r7_kern_np.json 2>> gpurun_out/r7_kern.err
RDP_NO_FLUSH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bwd_stream" -c 2 -o gpurun_out/r7_prof -f python tools/bench_kernel.py > gpurun_out/r7_ncu.log 2>&1
tail -3 gpurun_out/r7_pytest.log; cat gpurun_out/r7_kern.json gpurun_out/r7_kern_np.json
