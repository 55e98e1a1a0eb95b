#!/bin/bash
# ncu --set full on kernels of one LiDAR batch-8 pass (driver: tools/bench_kernel.py with few repetitions)
set -x
mkdir -p gpurun_out
export RDP_BENCH_REPS=1
python tools/bench_kernel.py > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:"${NCU_REGEX:-pfn_}" -s ${NCU_SKIP:-5} -c ${NCU_COUNT:-10} \
    -o gpurun_out/${NCU_OUT:-r2c_pfn} -f python tools/bench_kernel.py > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_run.log
