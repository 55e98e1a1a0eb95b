python tools/dbg_hostprof.py > gpurun_out/r16_hostprof.log 2>&1
head -c 7000 gpurun_out/r16_hostprof.log
