#!/bin/bash
# source-level captures of the elimination kernels (1024 x 2048 row reduction over GF(998244353))
mkdir -p gpurun_out /tmp/prof
for k in k_rr_panel_w k_rr_panel_update; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^${k}\$|${k}<" -s 40 -c 1 -o /tmp/prof/$k \
      python bench.py --workload matinv --steps 1 --warmup 1 --no-cpu --no-extras --no-ref-e2e > /dev/null 2>&1
  ncu -i /tmp/prof/$k.ncu-rep --page source --csv > gpurun_out/r14_${k}_source.csv 2>/dev/null
  ncu -i /tmp/prof/$k.ncu-rep --page raw --csv > gpurun_out/r14_${k}_raw.csv 2>/dev/null
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_rr_panel<" -s 40 -c 1 -o /tmp/prof/k_rr_panel \
    python bench.py --workload matinv --steps 1 --warmup 1 --no-cpu --no-extras --no-ref-e2e > /dev/null 2>&1
ncu -i /tmp/prof/k_rr_panel.ncu-rep --page source --csv > gpurun_out/r14_k_rr_panel_source.csv 2>/dev/null
ncu -i /tmp/prof/k_rr_panel.ncu-rep --page raw --csv > gpurun_out/r14_k_rr_panel_raw.csv 2>/dev/null
ls -la gpurun_out | grep r14_
