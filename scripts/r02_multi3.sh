#!/bin/bash
# 2-GPU sanity pass after the round-2 kernel changes: the default bench under torchrun (extras at every N)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err
echo "rc=$?"
tail -c 1500 gpurun_out/r02_bench_n2_final.json | head -c 1500
tail -5 gpurun_out/r02_bench_n2_final.err
