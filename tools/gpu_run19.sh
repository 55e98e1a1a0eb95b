set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r19_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r19_bench.json 2> gpurun_out/r19_bench.err
tail -n 12 gpurun_out/r19_pytest.log; cat gpurun_out/r19_bench.json | cut -c1-1500; tail -n 3 gpurun_out/r19_bench.err
