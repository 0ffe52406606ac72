#!/bin/bash
mkdir -p gpurun_out /tmp/prof
GFB_RS_FAST=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_rs_decode_fast -s 3 -c 1 -o /tmp/prof/rsfast \
    python bench.py --workload rs_decode --batch 200000 --steps 3 --warmup 3 --no-cpu --no-extras --no-ref-e2e > gpurun_out/r02_ncu_rsfast.log 2>&1
ncu -i /tmp/prof/rsfast.ncu-rep --page raw --csv > gpurun_out/r02_rsfast_raw.csv 2>/dev/null
ncu -i /tmp/prof/rsfast.ncu-rep --page source --csv > gpurun_out/r02_rsfast_source.csv 2>/dev/null
