set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3_pytest.log
python tools/bench_kernel.py > gpurun_out/r3_kern.json 2> gpurun_out/r3_kern.err
RDP_PFN_LEGACY=1 python tools/bench_kernel.py > gpurun_out/r3_kern_legacy.json 2>> gpurun_out/r3_kern.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
RDP_NO_FLUSH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pfn_rows|pfn_tile" -c 8 -o gpurun_out/r3_prof -f python tools/bench_kernel.py > gpurun_out/r3_ncu.log 2>&1
tail -15 gpurun_out/r3_pytest.log; cat gpurun_out/r3_kern.json gpurun_out/r3_kern_legacy.json; cat gpurun_out/r3_bench.json
