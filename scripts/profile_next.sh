#!/bin/bash
# Run on the GPU box (under gpurun): one `ncu --set full` capture per hot kernel of the §8(f) rows
# (blocked elimination, root search, hybrid mixed-radix DFT).  CSV exports only (see profile_gpu.sh).
set -u
OUT=gpurun_out
mkdir -p $OUT /tmp/prof
TAG=${1:-r01}
capture() {  # name, kernel regex, skip, count, bench args...
    local name=$1 regex=$2 skip=$3 count=$4; shift 4
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count \
        -o /tmp/prof/$name python bench.py "$@" --no-cpu --no-extras > /dev/null 2>&1
    ncu -i /tmp/prof/$name.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
    ncu -i /tmp/prof/$name.ncu-rep --page details --csv > $OUT/${TAG}_${name}_details.csv 2>/dev/null
}
capture elim "k_rr_panel" 45 3 --workload matinv --steps 1 --warmup 3
capture roots "k_roots" 6 1 --workload poly_roots --steps 1 --warmup 3
capture fftmixed "k_mixed_stage_v|k_transpose_rows" 10 3 --workload fft_mixed --steps 1 --warmup 3
ls -la $OUT | tail -8
