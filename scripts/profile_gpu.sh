#!/bin/bash
# Run on the GPU box (under gpurun): ncu launch list + one --set full capture per hot kernel.
# Leaves only small CSV exports in gpurun_out/ (the .ncu-rep files exceed the gpurun return quota).
set -u
OUT=gpurun_out
mkdir -p $OUT /tmp/prof
TAG=${1:-r01}

# every launch of the default bench command with its device time (shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file $OUT/${TAG}_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_bench_under_ncu.log 2>&1

capture() {  # name, kernel regex, skip, count, bench args...
    local name=$1 regex=$2 skip=$3 count=$4; shift 4
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count \
        -o /tmp/prof/$name python bench.py "$@" --no-cpu --no-extras > /dev/null 2>&1
    ncu -i /tmp/prof/$name.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
    ncu -i /tmp/prof/$name.ncu-rep --page details --csv > $OUT/${TAG}_${name}_details.csv 2>/dev/null
    ncu -i /tmp/prof/$name.ncu-rep --page source --csv > $OUT/${TAG}_${name}_source.csv 2>/dev/null
    # keep the source page small: drop it if it is huge
    if [ $(stat -c %s $OUT/${TAG}_${name}_source.csv 2>/dev/null || echo 0) -gt 6000000 ]; then
        head -c 6000000 $OUT/${TAG}_${name}_source.csv > $OUT/${TAG}_${name}_source_head.csv; rm $OUT/${TAG}_${name}_source.csv
    fi
}
capture mul k_binary 3 1 --steps 4 --warmup 3
capture ntt k_ntt_pass 6 2 --workload ntt --batch 64 --steps 3 --warmup 3
capture rs k_rs_decode 3 1 --workload rs_decode --batch 200000 --steps 3 --warmup 3
capture gf128 k_binary 3 1 --workload gf2_128_mul --steps 3 --warmup 3
ls -la $OUT
