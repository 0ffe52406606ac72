#!/bin/bash
# cluster elimination panel: A/B + oracle tests, the linalg tests, matinv bench + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_elim_cluster.py -m gpu -x -q > gpurun_out/r15_cluster_tests.log 2>&1
tail -15 gpurun_out/r15_cluster_tests.log
timeout 900 python -m pytest tests/test_gpu_next.py tests/test_gpu_parity.py tests/test_gpu_binding.py -m gpu -q -k "linalg or solve or inv or row_reduce or det or rank or null or lu or space" > gpurun_out/r15_linalg_tests.log 2>&1
tail -5 gpurun_out/r15_linalg_tests.log
timeout 300 python bench.py --workload matinv --steps 10 --warmup 3 --no-ref-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/r15_matinv.json
GFB_ELIM_CLUSTER=0 timeout 300 python bench.py --workload matinv --steps 10 --warmup 3 --no-ref-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/r15_matinv_old.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r15_matinv_launches.csv \
    python bench.py --workload matinv --steps 1 --warmup 1 --no-cpu --no-extras --no-ref-e2e > /dev/null 2>&1
python - <<'PY'
import json
for w in ("matinv","matinv_old"):
    try:
        d=json.loads(open(f"gpurun_out/r15_{w}.json").read().strip().splitlines()[-1])
        print(w, d["value"], d.get("ms_per_step"))
    except Exception as e: print(w, "ERR", e)
PY
