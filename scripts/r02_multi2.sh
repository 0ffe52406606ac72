#!/bin/bash
# 2-GPU pass: sharded single-NTT test (NCCL all-to-all), default bench at N=2 (extras at every N), ntt_sharded at N=1,2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded_ntt.py tests/test_gpu_fullsize.py -q -m gpu -k "sharded or two_ranks or config2" > gpurun_out/r02_pytest_multi2.log 2>&1
echo "multi2 tests rc=$?" > gpurun_out/r02_multi2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_ref_n2.json 2>> gpurun_out/r02_bench_n2.err
timeout 600 python bench.py --workload ntt_sharded --steps 10 --no-cpu --no-extras > gpurun_out/r02_ntt_sharded_n1.json 2>> gpurun_out/r02_multi2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --workload ntt_sharded --steps 10 --no-cpu --no-extras > gpurun_out/r02_ntt_sharded_n2.json 2>> gpurun_out/r02_multi2.log
timeout 300 python bench.py --workload gfp_inv --steps 20 --no-cpu --no-extras --no-ref-e2e 2>>gpurun_out/r02_multi2.log | tail -1 > gpurun_out/r02_bench_gfp_inv_b.json
timeout 300 python bench.py --workload gfp_div --steps 20 --no-cpu --no-extras --no-ref-e2e 2>>gpurun_out/r02_multi2.log | tail -1 > gpurun_out/r02_bench_gfp_div_b.json
