#!/bin/bash
# Quick iteration pass: kernel timings (CUDA events) + instruction counts / issue rate of selected kernels.
#   gpurun --timeout 600 -- 'K=pfn_apply bash tools/gpu_quick.sh'
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$TESTS" 2>&1 | tail -3; fi
timeout 300 python tools/bench_kernel.py 2> gpurun_out/quick.err || tail gpurun_out/quick.err
RDP_BENCH_REPS=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"${K:-pfn_}" -s ${SKIP:-2} -c ${COUNT:-4} --csv --log-file gpurun_out/quick.csv python tools/bench_kernel.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/quick.csv')) if r]
hi=next(i for i,r in enumerate(rows) if 'Kernel Name' in r); H={h:k for k,h in enumerate(rows[hi])}
out={}
for r in rows[hi+1:]:
    if len(r)<=H['Metric Value']: continue
    out.setdefault((r[H['ID']],r[H['Kernel Name']][:70]),{})[r[H['Metric Name']]]=r[H['Metric Value']]
seen=set()
for (i,k),m in out.items():
    if k in seen: continue
    seen.add(k); print(k); print('   ',{a.split('.')[0].replace('smsp__','').replace('sm__','').replace('launch__',''):b for a,b in m.items()})
PY
