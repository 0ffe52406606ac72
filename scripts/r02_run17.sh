#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_elim_cluster.py tests/test_gpu_next.py -m gpu -x -q > gpurun_out/r17_tests.log 2>&1
tail -4 gpurun_out/r17_tests.log
timeout 300 python bench.py --workload matinv --steps 10 --warmup 3 --no-ref-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/r17_matinv.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r17_matinv_launches.csv \
    python bench.py --workload matinv --steps 1 --warmup 1 --no-cpu --no-extras --no-ref-e2e > /dev/null 2>&1
python - <<'PY'
import json,csv,collections
d=json.loads(open("gpurun_out/r17_matinv.json").read().strip().splitlines()[-1])
print("matinv", d["value"], d.get("ms_per_step"))
rows=list(csv.reader(open('gpurun_out/r17_matinv_launches.csv')))
hdr=None;agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if 'Kernel Name' in r: hdr={h:i for i,h in enumerate(r)}; continue
    if hdr and len(r)>=len(hdr):
        try:
            n=r[hdr['Kernel Name']].split('(')[0][:70]; v=float(r[hdr['Metric Value']].replace(',',''))
            agg[n][0]+=1; agg[n][1]+=v
        except: pass
for n,(c,v) in sorted(agg.items(),key=lambda x:-x[1][1])[:4]: print(n,c,round(v/1e3,1),'us total', round(v/c/1e3,2),'us avg')
PY
