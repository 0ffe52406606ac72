#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/debug_binding.py > gpurun_out/r02_debug_binding.log 2>&1
echo "debug rc=$?" > gpurun_out/r02_run4.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_ntt_fused.py tests/test_gpu_sharded_ntt.py -q -m gpu --timeout 300 -k "rs or bch or reed or fused or sharded or config" > gpurun_out/r02_pytest_rsfast.log 2>&1
echo "rsfast rc=$?" >> gpurun_out/r02_run4.log
GFB_NTT_TILE=8 timeout 600 python -m pytest tests/test_gpu_ntt_fused.py -q -m gpu --timeout 300 > gpurun_out/r02_pytest_fused_w8.log 2>&1
echo "fused w8 rc=$?" >> gpurun_out/r02_run4.log
for tile in 16 8; do
for cfg in "3 8" "5 12" "7 16" "9 22"; do
  set -- $cfg
  GFB_NTT_TILE=$tile GFB_NTT_LAG=$1 GFB_NTT_RING=$2 timeout 300 python bench.py --workload ntt --batch 4096 --steps 5 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_run4.log | tail -1 > gpurun_out/r02_ntt_w${tile}_lag$1_ring$2.json
done
done
for f in 1 0; do
  GFB_RS_FAST=$f timeout 300 python bench.py --workload rs_decode --steps 10 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_run4.log | tail -1 > gpurun_out/r02_rs_fast$f.json
done
GFB_RS_FAST=1 timeout 300 python bench.py --workload rs_decode_erasures --steps 10 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_run4.log | tail -1 > gpurun_out/r02_rs_era_fast1.json
timeout 300 python bench.py --workload gf256_mul_1e6 --steps 50 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_run4.log | tail -1 > gpurun_out/r02_bench_gf256_mul_1e6_b.json
