#!/bin/bash
# round-2 second GPU call: fused NTT correctness first (short timeouts: a hang must not hold the box), then binding, then A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ntt_fused.py -x -q -m gpu > gpurun_out/r02_pytest_fused.log 2>&1
echo "fused rc=$?" > gpurun_out/r02_run2.log
timeout 900 python -m pytest tests/test_gpu_binding.py -q -m gpu > gpurun_out/r02_pytest_binding.log 2>&1
echo "binding rc=$?" >> gpurun_out/r02_run2.log
for f in 1 0; do
  GFB_NTT_FUSED=$f timeout 300 python bench.py --workload ntt --batch 4096 --steps 5 --no-cpu --no-ref-e2e > gpurun_out/r02_ntt_fused$f.json 2>> gpurun_out/r02_run2.log
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r02_bench_default2.json 2>> gpurun_out/r02_run2.log
