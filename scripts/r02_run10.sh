#!/bin/bash
# source-level ncu capture of the bit-sliced RS decoder
mkdir -p gpurun_out /tmp/prof
export GFB_RS_DECODER=bs
timeout 300 python bench.py --workload rs_decode --steps 10 --warmup 3 --no-ref-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/r10_bs.json
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_rs_decode_bs -s 3 -c 1 -o /tmp/prof/rsbs \
    python bench.py --workload rs_decode --batch 200000 --steps 3 --warmup 3 --no-cpu --no-extras --no-ref-e2e > /dev/null 2>&1
ncu -i /tmp/prof/rsbs.ncu-rep --page raw --csv > gpurun_out/r10_rsbs_raw.csv 2>/dev/null
ncu -i /tmp/prof/rsbs.ncu-rep --page details --csv > gpurun_out/r10_rsbs_details.csv 2>/dev/null
ncu -i /tmp/prof/rsbs.ncu-rep --page source --csv > gpurun_out/r10_rsbs_source.csv 2>/dev/null
ls -la gpurun_out | grep r10_
cat gpurun_out/r10_bs.json | cut -c1-200
