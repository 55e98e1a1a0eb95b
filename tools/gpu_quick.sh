set -x
python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
RDP_BENCH_ONLY_EVAL=1 python tools/bench_kernel.py > gpurun_out/q_kern.json 2> gpurun_out/q_kern.err
RDP_BENCH_ONLY_EVAL=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"group_rows|quantize" -c 12 --csv --log-file gpurun_out/q_index_launches.csv python tools/bench_kernel.py > gpurun_out/q_ncu.log 2>&1
tail -n 4 gpurun_out/q_pytest.log; cat gpurun_out/q_kern.json
