set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r10_pytest.log
python tools/bench_kernel.py > gpurun_out/r10_kern.json 2> gpurun_out/r10_kern.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err
RDP_NO_FLUSH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pfn_tile" -c 4 -o gpurun_out/r10_prof -f python tools/bench_kernel.py > gpurun_out/r10_ncu.log 2>&1
tail -3 gpurun_out/r10_pytest.log; cat gpurun_out/r10_kern.json; cat gpurun_out/r10_bench.json
