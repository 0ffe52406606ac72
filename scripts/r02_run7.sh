#!/bin/bash
mkdir -p gpurun_out /tmp/prof
timeout 600 python -m pytest tests/test_gpu_binding.py -q -m gpu -k "other_codes or reed" > gpurun_out/r02_pytest_binding4.log 2>&1
echo "binding4 rc=$?" > gpurun_out/r02_run7.log
GFB_RS_FAST=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_rs_decode_fast -s 3 -c 1 -o /tmp/prof/rsfast \
    python bench.py --workload rs_decode --batch 200000 --steps 3 --warmup 3 --no-cpu --no-extras --no-ref-e2e > gpurun_out/r02_ncu_rsfast2.log 2>&1
ncu -i /tmp/prof/rsfast.ncu-rep --page raw --csv > gpurun_out/r02_rsfast2_raw.csv 2>/dev/null
ncu -i /tmp/prof/rsfast.ncu-rep --page source --csv > gpurun_out/r02_rsfast2_source.csv 2>/dev/null
timeout 300 python bench.py --workload gf256_mul_1e6 --steps 50 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_run7.log | tail -1 > gpurun_out/r02_bench_gf256_mul_1e6_graph.json
