#!/bin/bash
# final pass of round 2: full GPU suite, smoke, the rows whose kernels changed (RS decode, elimination), the default bench line
# (both arms), the ncu pass
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_final3.log 2>&1
echo "gpu suite rc=$?" > gpurun_out/r02_final3.log
tail -3 gpurun_out/r02_pytest_gpu_final3.log >> gpurun_out/r02_final3.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke3.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02_final3.log
OUT=gpurun_out/r02_bench_changed3.jsonl; : > $OUT
for wl in rs_decode rs_decode_erasures bch_decode matinv; do
    timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-extras 2>gpurun_out/r02_err3_$wl.log | tail -1 >> $OUT
done
GFB_ELIM_CLUSTER=0 timeout 300 python bench.py --workload matinv --steps 10 --warmup 3 --no-extras --no-cpu --no-ref-e2e 2>/dev/null | tail -1 > gpurun_out/r02_matinv_single_cta.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_default_final3.json 2> gpurun_out/r02_bench_default_final3.err
bash scripts/r02_profile.sh r02 >> gpurun_out/r02_final3.log 2>&1
cat gpurun_out/r02_final3.log | head -20
python - <<'PY'
import json
for l in open("gpurun_out/r02_bench_changed3.jsonl"):
    try:
        d=json.loads(l); print(d["metric"], d["value"], d.get("e2e",{}).get("value"))
    except Exception as e: print("ERR", e)
PY
