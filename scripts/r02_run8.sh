#!/bin/bash
mkdir -p gpurun_out
GFB_RS_DECODER=bs timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_next.py -q -m gpu --timeout 300 -k "rs or bch or reed" > gpurun_out/r02_pytest_rsbs.log 2>&1
echo "rs bs tests rc=$?" > gpurun_out/r02_run8.log
for dec in bs log; do
  GFB_RS_DECODER=$dec timeout 300 python bench.py --workload rs_decode --steps 10 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_run8.log | tail -1 > gpurun_out/r02_rs3_$dec.json
done
GFB_RS_DECODER=bs timeout 300 python bench.py --workload rs_decode_erasures --steps 10 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_run8.log | tail -1 > gpurun_out/r02_rs3_era_bs.json
GFB_RS_DECODER=bs timeout 300 python bench.py --workload bch_decode --steps 10 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_run8.log | tail -1 > gpurun_out/r02_rs3_bch_bs.json
timeout 300 python bench.py --workload gfp_inv --steps 20 --no-cpu --no-extras --no-ref-e2e 2>>gpurun_out/r02_run8.log | tail -1 > gpurun_out/r02_bench_gfp_inv_c.json
