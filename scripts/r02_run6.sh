#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_binding.py -q -m gpu -o faulthandler_timeout=600 > gpurun_out/r02_pytest_binding3.log 2>&1
echo "binding rc=$?" > gpurun_out/r02_run6.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_binding.py -x > gpurun_out/r02_pytest_gpu_all.log 2>&1
echo "gpu suite rc=$?" >> gpurun_out/r02_run6.log
for f in 1 0; do
  GFB_RS_FAST=$f timeout 300 python bench.py --workload rs_decode --steps 10 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_run6.log | tail -1 > gpurun_out/r02_rs2_fast$f.json
done
GFB_RS_FAST=1 timeout 300 python bench.py --workload rs_decode_erasures --steps 10 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_run6.log | tail -1 > gpurun_out/r02_rs2_era_fast1.json
GFB_RS_FAST=1 timeout 300 python bench.py --workload bch_decode --steps 10 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_run6.log | tail -1 > gpurun_out/r02_bch_fast1.json
GFB_COPY_THREADS=32 timeout 300 python bench.py --steps 20 --no-extras --no-cpu 2>>gpurun_out/r02_run6.log | tail -1 > gpurun_out/r02_bench_copy32.json
GFB_COPY_THREADS=8 timeout 300 python bench.py --steps 20 --no-extras --no-cpu 2>>gpurun_out/r02_run6.log | tail -1 > gpurun_out/r02_bench_copy8.json
