#!/bin/bash
# cluster elimination panel, second pass: tests, bench, source-level capture of the panel kernel
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests/test_gpu_elim_cluster.py -m gpu -x -q > gpurun_out/r16_cluster_tests.log 2>&1
tail -5 gpurun_out/r16_cluster_tests.log
timeout 300 python bench.py --workload matinv --steps 10 --warmup 3 --no-ref-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/r16_matinv.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rr_panel_cluster -s 40 -c 1 -o /tmp/prof/rrc \
    python bench.py --workload matinv --steps 1 --warmup 1 --no-cpu --no-extras --no-ref-e2e > /dev/null 2>&1
ncu -i /tmp/prof/rrc.ncu-rep --page source --csv > gpurun_out/r16_rrc_source.csv 2>/dev/null
ncu -i /tmp/prof/rrc.ncu-rep --page raw --csv > gpurun_out/r16_rrc_raw.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r16_matinv_launches.csv \
    python bench.py --workload matinv --steps 1 --warmup 1 --no-cpu --no-extras --no-ref-e2e > /dev/null 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r16_matinv.json").read().strip().splitlines()[-1])
print("matinv", d["value"], d.get("ms_per_step"))
PY
