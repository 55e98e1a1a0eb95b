set -x
RDP_PFN_ROWS=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r15_pytest_rows.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r15_pytest_rows.log
RDP_PFN_ROWS=1 timeout 300 python tools/bench_kernel.py > gpurun_out/r15_kern_rows.json 2> gpurun_out/r15_kern.err
RDP_PFN_ROWS=1 RDP_NO_FLUSH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pfn_rows" --launch-skip 6 -c 1 -o gpurun_out/r15_prof -f python tools/bench_kernel.py > gpurun_out/r15_ncu.log 2>&1
tail -n 8 gpurun_out/r15_pytest_rows.log; cat gpurun_out/r15_kern_rows.json
