set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/t2_lean.json 2> gpurun_out/t2_lean.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --dp ddp > gpurun_out/t2_ddp.json 2> gpurun_out/t2_ddp.err
cat gpurun_out/t2_lean.json gpurun_out/t2_ddp.json | cut -c1-330; tail -n 3 gpurun_out/t2_lean.err
