#!/bin/bash
# final measurement pass, part 1: full GPU suite, smoke, every workload (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_final.log 2>&1
echo "gpu suite rc=$?" > gpurun_out/r02_final1.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02_final1.log
bash scripts/bench_all.sh r02 >> gpurun_out/r02_final1.log 2>&1
for f in 0 1; do
  GFB_RS_ENCODE_LFSR=$f timeout 300 python bench.py --workload rs_encode --steps 10 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_final1.log | tail -1 > gpurun_out/r02_rs_encode_lfsr$f.json
done
