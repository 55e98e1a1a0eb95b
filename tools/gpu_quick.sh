set -x
python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
python tools/dbg_timeline.py > gpurun_out/q_timeline.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
tail -n 12 gpurun_out/q_pytest.log; tail -n 6 gpurun_out/q_timeline.log; cut -c1-260 gpurun_out/q_bench.json; tail -n 3 gpurun_out/q_bench.err
