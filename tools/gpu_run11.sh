set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r12_pytest.log
python tools/bench_kernel.py > gpurun_out/r12_kern.json 2> gpurun_out/r12_kern.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r12_bench.json 2> gpurun_out/r12_bench.err
RDP_NO_FLUSH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pfn_tile" --launch-skip 56 -c 2 -o gpurun_out/r12_prof_bwd -f python tools/bench_kernel.py > gpurun_out/r12_ncu.log 2>&1
tail -n 3 gpurun_out/r12_pytest.log; cat gpurun_out/r12_kern.json; cat gpurun_out/r12_bench.json; tail -n 5 gpurun_out/r12_bench.err
