#!/bin/bash
# final measurement pass, part 2: rows that failed in part 1 (bench bugs), then the ncu pass
mkdir -p gpurun_out
OUT=gpurun_out/r02_bench_changed.jsonl; : > $OUT
for wl in rs_encode convolve convolve_gf256 ntt_sharded; do
    timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-extras 2>gpurun_out/r02_err2_$wl.log | tail -1 >> $OUT
    timeout 400 python bench.py --impl reference --workload $wl --steps 3 --warmup 1 2>>gpurun_out/r02_err2_$wl.log | tail -1 >> $OUT
done
GFB_RS_ENCODE_LFSR=0 timeout 300 python bench.py --workload rs_encode --steps 10 --no-cpu --no-ref-e2e 2>>gpurun_out/r02_final2.log | tail -1 > gpurun_out/r02_rs_encode_lfsr0.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_default_final.json 2> gpurun_out/r02_bench_default_final.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_ref_final.json 2>> gpurun_out/r02_bench_default_final.err
bash scripts/r02_profile.sh r02 >> gpurun_out/r02_final2.log 2>&1
