#!/bin/bash
# default decoder = bit-sliced: full GPU suite, RS rows, ncu capture of the decoder, launch list of the matrix inverse
mkdir -p gpurun_out /tmp/prof
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r12_gpu_tests.log 2>&1
tail -3 gpurun_out/r12_gpu_tests.log
for w in rs_decode rs_decode_erasures; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-ref-e2e 2>/dev/null | tail -1 > gpurun_out/r12_$w.json
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_rs_decode_bs -s 3 -c 1 -o /tmp/prof/rsbs \
    python bench.py --workload rs_decode --batch 200000 --steps 3 --warmup 3 --no-cpu --no-extras --no-ref-e2e > /dev/null 2>&1
ncu -i /tmp/prof/rsbs.ncu-rep --page raw --csv > gpurun_out/r12_rsbs_raw.csv 2>/dev/null
ncu -i /tmp/prof/rsbs.ncu-rep --page details --csv > gpurun_out/r12_rsbs_details.csv 2>/dev/null
ncu -i /tmp/prof/rsbs.ncu-rep --page source --csv > gpurun_out/r12_rsbs_source.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r12_matinv_launches.csv \
    python bench.py --workload matinv --steps 1 --warmup 1 --no-cpu --no-extras --no-ref-e2e > /dev/null 2>&1
python - <<'PY'
import json
for w in ("rs_decode","rs_decode_erasures"):
    try:
        d=json.loads(open(f"gpurun_out/r12_{w}.json").read().strip().splitlines()[-1])
        print(w, d["value"], d.get("e2e"), d.get("secondary_roofline"))
    except Exception as e: print(w, "ERR", e)
PY
