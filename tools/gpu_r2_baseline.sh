#!/bin/bash
# Round-2 "before" evidence on the round-1 kernels: new full-size parity tests + DRAM counters of every kernel of one
# LiDAR batch-8 train step (ncu --set full), incl. the 7 index kernels the r1 verdict asked for.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -n 5 gpurun_out/r2a_pytest.log
python tools/bench_kernel.py > gpurun_out/r2a_kern.json 2> gpurun_out/r2a_kern.err && \
ncu --set full --clock-control none -k regex:'quantize_mark|bitmap_rank|zero_counts|rank_count|count_scan|group_rows|pillar_table' -s 14 -c 7 \
    -o gpurun_out/r2a_index python tools/bench_kernel.py > gpurun_out/r2a_ncu.log 2>&1
cat gpurun_out/r2a_kern.json
