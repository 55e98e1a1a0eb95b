#!/bin/bash
# One GPU-box pass over everything the round-end driver runs: parity tests, smoke, bench (both arms).
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/check_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/check_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/check_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/check_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/check_bench_ref.json 2> gpurun_out/check_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err
tail -n 3 gpurun_out/check_pytest.log gpurun_out/check_smoke.log; cut -c1-400 gpurun_out/check_bench_ref.json; cut -c1-600 gpurun_out/check_bench.json
