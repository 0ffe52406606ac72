#!/bin/bash
# source-level capture of the current cluster panel kernel
mkdir -p gpurun_out /tmp/prof
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rr_panel_cluster -s 12 -c 1 -o /tmp/prof/rrc2 \
    python bench.py --workload matinv --steps 1 --warmup 1 --no-cpu --no-extras --no-ref-e2e > /dev/null 2>&1
ncu -i /tmp/prof/rrc2.ncu-rep --page source --csv > gpurun_out/r18_rrc_source.csv 2>/dev/null
ncu -i /tmp/prof/rrc2.ncu-rep --page raw --csv > gpurun_out/r18_rrc_raw.csv 2>/dev/null
ncu -i /tmp/prof/rrc2.ncu-rep --page details --csv > gpurun_out/r18_rrc_details.csv 2>/dev/null
ls -la gpurun_out | grep r18_
