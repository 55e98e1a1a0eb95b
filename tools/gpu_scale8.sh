#!/bin/bash
# 8-GPU box: weak scaling of the resident step (N = 1 and N = 8 back to back on the same box).
set -x
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/z1.json 2> gpurun_out/z1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/z8.json 2> gpurun_out/z8.err
cut -c1-230 gpurun_out/z1.json gpurun_out/z8.json
