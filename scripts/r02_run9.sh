#!/bin/bash
# bit-sliced RS decoder with thread-per-codeword key equation / Forney: parity, then A/B against the log-domain kernel
mkdir -p gpurun_out
export GFB_RS_DECODER=bs
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_next.py -m gpu -x -q -k "rs or bch or reed or RS or BCH" > gpurun_out/r9_bs_tests.log 2>&1
tail -3 gpurun_out/r9_bs_tests.log
for w in rs_decode; do
  GFB_RS_DECODER=bs timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-ref-e2e > gpurun_out/r9_bs_$w.json 2>gpurun_out/r9_bs_$w.err
  GFB_RS_DECODER=log timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-ref-e2e > gpurun_out/r9_log_$w.json 2>gpurun_out/r9_log_$w.err
done
python - <<'PY'
import json
for k in ("bs","log"):
    try:
        d=json.loads(open(f"gpurun_out/r9_{k}_rs_decode.json").read().strip().splitlines()[-1])
        print(k, d["value"], d.get("extras",{}))
    except Exception as e: print(k, "ERR", e)
PY
