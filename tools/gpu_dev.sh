#!/bin/bash
# Development pass on a GPU box: parity tests (stop at first failure), then kernel timings of the LiDAR batch.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/dev_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/dev_pytest.log
tail -n 40 gpurun_out/dev_pytest.log
timeout 300 python tools/bench_kernel.py > gpurun_out/dev_kern.json 2> gpurun_out/dev_kern.err; cat gpurun_out/dev_kern.json; tail -n 5 gpurun_out/dev_kern.err
