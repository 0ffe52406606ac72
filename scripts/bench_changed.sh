#!/bin/bash
# Run on the GPU box (under gpurun): re-measure a subset of bench.py workloads (our arm; the CPU reference arm only
# for the names after `--ref`) and merge the lines into an existing <tag>_bench_all.jsonl by metric + impl.
#   scripts/bench_changed.sh r01 "rs_decode ntt" --ref "matinv"
TAG=${1:-r01}; GPU_LIST=$2; REF_LIST=""
[ "$3" = "--ref" ] && REF_LIST=$4
OUT=gpurun_out/${TAG}_bench_changed.jsonl
mkdir -p gpurun_out; : > $OUT
for wl in $GPU_LIST; do
    extra=""
    case $wl in ntt|intt) extra="--batch 1036";; esac
    timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-extras --no-cpu $extra 2>>gpurun_out/${TAG}_err_changed.log | tail -1 >> $OUT
done
for wl in $REF_LIST; do
    timeout 600 python bench.py --impl reference --workload $wl --steps 3 --warmup 1 2>>gpurun_out/${TAG}_err_changed.log | tail -1 >> $OUT
done
[ -s gpurun_out/${TAG}_err_changed.log ] || rm -f gpurun_out/${TAG}_err_changed.log
wc -l $OUT
