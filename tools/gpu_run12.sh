python tools/dbg_timeline.py > gpurun_out/r13_timeline.log 2>&1
tail -n 30 gpurun_out/r13_timeline.log
