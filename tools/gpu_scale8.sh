#!/bin/bash
# 8-GPU box: weak scaling of the resident step, with and without per-rank CPU pinning.
set -x
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/z1.json 2> gpurun_out/z1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/z8.json 2> gpurun_out/z8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --pin-cpus 1 > gpurun_out/z8pin.json 2> gpurun_out/z8pin.err
nproc; lscpu | grep -E "Thread|Core|Socket|NUMA" ; cut -c1-230 gpurun_out/z1.json gpurun_out/z8.json gpurun_out/z8pin.json
