#!/bin/bash
# Round-2 ncu pass (run under gpurun, one GPU): launch list of the default bench command + one --set full capture per hot
# kernel of this round.  Leaves CSV exports in gpurun_out/ (scripts/summarize_ncu.py r02 copies them to profiles/).
set -u
OUT=gpurun_out
TAG=${1:-r02}
mkdir -p $OUT /tmp/prof
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/${TAG}_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-ref-e2e > $OUT/${TAG}_bench_under_ncu.log 2>&1
capture() {  # name, kernel regex, skip, count, bench args...
    local name=$1 regex=$2 skip=$3 count=$4; shift 4
    timeout 500 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count \
        -o /tmp/prof/$name python bench.py "$@" --no-cpu --no-extras --no-ref-e2e > /dev/null 2>&1
    ncu -i /tmp/prof/$name.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
    ncu -i /tmp/prof/$name.ncu-rep --page details --csv > $OUT/${TAG}_${name}_details.csv 2>/dev/null
}
capture mul k_binary 3 1 --steps 4 --warmup 3
capture ntt k_ntt_fused 1 1 --workload ntt --batch 592 --steps 2 --warmup 3
capture rs k_rs_decode 3 1 --workload rs_decode --batch 200000 --steps 3 --warmup 3
capture gf256 k_tab8_aa 3 1 --workload gf256_mul --steps 3 --warmup 3
capture gf128 k_binary 3 1 --workload gf2_128_mul --steps 3 --warmup 3
capture elim k_rr_panel_cluster 12 1 --workload matinv --steps 1 --warmup 1
ncu -i /tmp/prof/elim.ncu-rep --page source --csv > $OUT/${TAG}_elim_source.csv 2>/dev/null
ncu -i /tmp/prof/rs.ncu-rep --page source --csv > $OUT/${TAG}_rs_source.csv 2>/dev/null
ls -la $OUT | grep ${TAG}_
