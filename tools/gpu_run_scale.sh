set -x
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/s1.json 2> gpurun_out/s1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/s8.json 2> gpurun_out/s8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/s4.json 2> gpurun_out/s4.err
nproc; cat gpurun_out/s1.json gpurun_out/s4.json gpurun_out/s8.json | cut -c1-330; tail -n 3 gpurun_out/s8.err
