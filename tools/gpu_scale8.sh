#!/bin/bash
# 8-GPU box: weak scaling (8 frames per GPU) and strong scaling (global batch 64) of the resident step, N = 1 and N = 8
# back to back on the same box.   gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_scale8.sh'
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs > gpurun_out/z1.json 2> gpurun_out/z1.err
$T --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/z8.json 2> gpurun_out/z8.err
if [ -n "$STRONG" ]; then
python bench.py --gpus 1 --steps 10 --warmup 3 --scaling strong --no-configs > gpurun_out/zs1.json 2> gpurun_out/zs1.err
$T --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --scaling strong > gpurun_out/zs8.json 2> gpurun_out/zs8.err
fi
if [ -n "$MID" ]; then
$T --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/z2.json 2> gpurun_out/z2.err
$T --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/z4.json 2> gpurun_out/z4.err
fi
for f in z1 z2 z4 z8 zs1 zs8; do [ -s gpurun_out/$f.json ] && python - <<PY
import json
d=json.load(open("gpurun_out/$f.json"))
print("$f", "N=%d"%d["n_gpus"], d["scaling"], "value %.3f Gpts/s"%(d["value"]/1e9), "ms/step %.3f"%d["ms_per_step"], "e2e %.3f Gpts/s (%.3f ms)"%(d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"]))
PY
done
tail -2 gpurun_out/z8.err
