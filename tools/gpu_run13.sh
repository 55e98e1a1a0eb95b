set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r14_pytest.log
python tools/dbg_timeline.py > gpurun_out/r14_timeline.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r14_bench.json 2> gpurun_out/r14_bench.err
tail -n 12 gpurun_out/r14_pytest.log; tail -n 12 gpurun_out/r14_timeline.log; cat gpurun_out/r14_bench.json; tail -n 5 gpurun_out/r14_bench.err
