#!/bin/bash
# Round-end evidence for profiles/: (1) ncu --set full, one launch of every librdp kernel of a LiDAR batch-8 pass
# (tools/bench_kernel.py); (2) the ncu launch list (gpu__time_duration) of the bench command itself.
mkdir -p gpurun_out
export RDP_BENCH_REPS=1
python tools/bench_kernel.py > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || { tail gpurun_out/prof_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"rdp|quantize|bitmap|rank_count|count_scan|group_rows|pillar_table" -s 0 -c 40 \
    -o gpurun_out/r02_full -f python tools/bench_kernel.py > gpurun_out/prof_ncu.log 2>&1
python bench.py --steps 2 --warmup 3 --no-configs > gpurun_out/prof_bench.json 2> gpurun_out/prof_bench.err || tail gpurun_out/prof_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-configs > gpurun_out/prof_bench_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02_bench_launches.csv "r02: python bench.py --steps 2 --warmup 3 --no-configs (ncu --metrics gpu__time_duration.sum --clock-control none)" > gpurun_out/r02_bench_launches_summary.txt
head -20 gpurun_out/r02_bench_launches_summary.txt
