set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/p_bench_ref.json 2> gpurun_out/p_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 20 --warmup 5 > gpurun_out/p_ncu_bench.log 2>&1
RDP_NO_FLUSH=1 ncu --set full --clock-control none --import-source on -k regex:"pfn_tile" --launch-skip 10 -c 1 -o gpurun_out/p_apply_eval -f python tools/bench_kernel.py > gpurun_out/p_ncu1.log 2>&1
RDP_NO_FLUSH=1 ncu --set full --clock-control none --import-source on -k regex:"pfn_tile" --launch-skip 40 -c 2 -o gpurun_out/p_train -f python tools/bench_kernel.py > gpurun_out/p_ncu2.log 2>&1
RDP_NO_FLUSH=1 ncu --set full --clock-control none --import-source on -k regex:"pfn_tile" --launch-skip 60 -c 1 -o gpurun_out/p_bwd -f python tools/bench_kernel.py > gpurun_out/p_ncu3.log 2>&1
python __graft_entry__.py > gpurun_out/p_smoke.log 2>&1 || python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p_smoke.log 2>&1
cat gpurun_out/p_bench.json | cut -c1-200; cat gpurun_out/p_bench_ref.json | cut -c1-300; tail -n 3 gpurun_out/p_smoke.log
