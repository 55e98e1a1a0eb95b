set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r17_pytest.log
python tools/dbg_timeline.py > gpurun_out/r17_timeline.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r17_bench.json 2> gpurun_out/r17_bench.err
tail -n 4 gpurun_out/r17_pytest.log; tail -n 12 gpurun_out/r17_timeline.log; cat gpurun_out/r17_bench.json | cut -c1-300
