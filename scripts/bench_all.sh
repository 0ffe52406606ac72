#!/bin/bash
# Run on the GPU box (under gpurun): every workload of bench.py, our arm and the reference arm (the unmodified reference
# from baseline/_ref where a galois-driven leg exists, the oracle port otherwise).
# Output: gpurun_out/<tag>_bench_all.jsonl (one JSON line per run).
TAG=${1:-r02}
OUT=gpurun_out/${TAG}_bench_all.jsonl
mkdir -p gpurun_out; : > $OUT
for wl in gfp_mul gfp_add gfp_inv gfp_div gfp_mul_u32 gfp_mul_barrett gf256_mul_1e6 gf256_mul gf2_32_mul ntt intt gf2_128_mul gf2_128_inv rs_encode rs_decode rs_decode_erasures bch_decode convolve convolve_gf256 matmul_rs matinv poly_roots fft_mixed ntt_sharded; do
    extra=""
    case $wl in ntt|intt) extra="--batch 4096";; esac
    timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-extras $extra 2>gpurun_out/${TAG}_err_$wl.log | tail -1 >> $OUT
    timeout 400 python bench.py --impl reference --workload $wl --steps 3 --warmup 1 2>>gpurun_out/${TAG}_err_$wl.log | tail -1 >> $OUT
    [ -s gpurun_out/${TAG}_err_$wl.log ] || rm -f gpurun_out/${TAG}_err_$wl.log
done
wc -l $OUT
