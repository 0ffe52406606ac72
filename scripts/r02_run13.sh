#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_next.py tests/test_gpu_binding.py -m gpu -x -q -k "rs or bch or reed or RS or BCH" > gpurun_out/r13_rs_tests.log 2>&1
tail -3 gpurun_out/r13_rs_tests.log
for w in rs_decode rs_decode_erasures; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-ref-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/r13_$w.json
done
python - <<'PY'
import json
for w in ("rs_decode","rs_decode_erasures"):
    try:
        d=json.loads(open(f"gpurun_out/r13_{w}.json").read().strip().splitlines()[-1])
        print(w, d["value"], d.get("roofline_secondary",{}).get("frac"))
    except Exception as e: print(w, "ERR", e)
PY
