#!/bin/bash
# Per-kernel device times of one LiDAR batch-8 pass (tools/bench_kernel.py, one repetition) as an ncu launch list.
#   gpurun --timeout 600 -- 'OUT=r02x bash tools/gpu_launches.sh'
set -x
mkdir -p gpurun_out
export RDP_BENCH_REPS=1
OUT=${OUT:-launches}
python tools/bench_kernel.py > gpurun_out/${OUT}_plain.json 2> gpurun_out/${OUT}_plain.err || { tail gpurun_out/${OUT}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${OUT}.csv \
    python tools/bench_kernel.py > gpurun_out/${OUT}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${OUT}.csv "${OUT}: tools/bench_kernel.py, LiDAR batch 8, RDP_BENCH_REPS=1" | tee gpurun_out/${OUT}_summary.txt
