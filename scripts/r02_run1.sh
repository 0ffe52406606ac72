#!/bin/bash
# round-2 first GPU call: binding tests, full GPU suite, default bench (both arms), NTT chunk experiment
mkdir -p gpurun_out
python -c "import numba, llvmlite; print('numba', numba.__version__, 'llvmlite', llvmlite.__version__)" > gpurun_out/r02_probe.log 2>&1
nproc >> gpurun_out/r02_probe.log; free -g >> gpurun_out/r02_probe.log; nvidia-smi -L >> gpurun_out/r02_probe.log
timeout 1500 python -m pytest tests/test_gpu_binding.py -x -q -m gpu > gpurun_out/r02_pytest_binding.log 2>&1
echo "binding rc=$?" >> gpurun_out/r02_probe.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_binding.py > gpurun_out/r02_pytest_gpu.log 2>&1
echo "gpu suite rc=$?" >> gpurun_out/r02_probe.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
for c in 74 37 18; do
  GFB_NTT_CHUNK=$c timeout 300 python bench.py --workload ntt --batch 1036 --steps 10 --no-cpu --no-ref-e2e > gpurun_out/r02_ntt_chunk$c.json 2>> gpurun_out/r02_probe.log
done
