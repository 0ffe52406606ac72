#!/bin/bash
# round-2 third GPU call: remaining binding tests with stack dumps on timeout, new kernel tests, fused-NTT ncu capture, benches
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests/test_gpu_binding.py -q -m gpu --timeout 150 -k "convolve or linalg or reed or other_codes or poly or fullsize" > gpurun_out/r02_pytest_binding2.log 2>&1
echo "binding2 rc=$?" > gpurun_out/r02_run3.log
timeout 900 python -m pytest tests/test_gpu_tab8.py tests/test_gpu_parity.py tests/test_gpu_next.py -q -m gpu --timeout 300 > gpurun_out/r02_pytest_new.log 2>&1
echo "new rc=$?" >> gpurun_out/r02_run3.log
# fused NTT: one full capture of the persistent kernel (592 transforms = 4 per CTA-pair of tiles) + launch list
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_fused -s 1 -c 1 -o /tmp/prof/ntt_fused \
    python bench.py --workload ntt --batch 592 --steps 2 --warmup 3 --no-cpu --no-extras --no-ref-e2e > gpurun_out/r02_ncu_fused.log 2>&1
ncu -i /tmp/prof/ntt_fused.ncu-rep --page raw --csv > gpurun_out/r02_ntt_fused_raw.csv 2>/dev/null
ncu -i /tmp/prof/ntt_fused.ncu-rep --page details --csv > gpurun_out/r02_ntt_fused_details.csv 2>/dev/null
ncu -i /tmp/prof/ntt_fused.ncu-rep --page source --csv > gpurun_out/r02_ntt_fused_source.csv 2>/dev/null
for wl in gf256_mul gf256_mul_1e6 convolve convolve_gf256 matmul_rs; do
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-extras 2>>gpurun_out/r02_run3.log | tail -1 > gpurun_out/r02_bench_$wl.json
done
timeout 300 python bench.py --impl reference --workload convolve --steps 3 --warmup 1 2>>gpurun_out/r02_run3.log | tail -1 > gpurun_out/r02_bench_convolve_ref.json
