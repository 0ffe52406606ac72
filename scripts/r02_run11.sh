#!/bin/bash
mkdir -p gpurun_out
timeout 900 env GFB_RS_DECODER=bs python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_next.py -m gpu -x -q -k "rs or bch or reed or RS or BCH" > gpurun_out/r11_bs_tests.log 2>&1
tail -3 gpurun_out/r11_bs_tests.log
for w in rs_decode rs_decode_erasures bch_decode; do
  GFB_RS_DECODER=bs timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-ref-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/r11_bs_$w.json
done
timeout 900 python -m pytest tests -m gpu -x -q -k "linalg or solve or inv or row_reduce or det or rank or null or lu" > gpurun_out/r11_linalg_tests.log 2>&1
tail -3 gpurun_out/r11_linalg_tests.log
timeout 300 python bench.py --workload matinv --steps 10 --warmup 3 --no-ref-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/r11_matinv.json
python - <<'PY'
import json
for w in ("bs_rs_decode","bs_rs_decode_erasures","bs_bch_decode","matinv"):
    try:
        d=json.loads(open(f"gpurun_out/r11_{w}.json").read().strip().splitlines()[-1])
        print(w, d["value"], d.get("ms_per_step"))
    except Exception as e: print(w, "ERR", e)
PY
