// gfb_ntt_fused.cuh — persistent, warp-specialised four-step NTT for N = 2^20 = 1024 x 1024, uint32 elements.
// (included by gfb_ntt.cu after Mod<>, radixR<> and gfb_ntt_plan)
//
// Same mathematics as k_ntt_pass1 / k_ntt_pass2 (reference: fft_jit radix-2 butterflies, _domains/_function.py:337-352;
// any exact factorisation gives the reference's X), different execution model:
//
//   * ONE launch per batch.  148 persistent CTAs (one per SM) pull tile tasks from a global ticket counter.  A task is
//     either a pass-1 tile (16 columns of one transform: 1024-point column NTTs + four-step twiddle) or a pass-2 tile
//     (16 rows of the intermediate: 1024-point row NTTs + transposed store).  Pass-2 tiles of transform T are queued
//     kLag transforms behind its pass-1 tiles and wait on a per-transform completion counter, so both passes overlap in
//     one grid with no wave quantisation and no host round trip.
//   * The intermediate lives in a ring of kRing transforms (32 MiB) that stays resident in the 126 MB L2: written by
//     pass 1, read once by pass 2 a few microseconds later, then overwritten — it does not travel to HBM.  A second
//     counter per transform gates the reuse of a ring slot.
//   * Inside a CTA: 16 compute warps + 1 load warp + 1 store warp around a 3-deep ring of 64 KB shared-memory tiles.
//       load warp   draws the ticket, waits for the task's dependency, and fills the next tile asynchronously:
//                   pass 1: 4096 x 16-byte cp.async (LDGSTS) gathers of the strided column tile,
//                   pass 2: one 64 KB cp.async.bulk (TMA) of the contiguous intermediate block;
//                   completion is signalled on an mbarrier (cp.async.mbarrier.arrive / complete_tx).
//       compute     two in-register radix-32 rounds per tile with one XOR-swizzled exchange through the dense tile
//                   (bank-conflict free without padding, which is what lets TMA write the tile);
//                   pass 1 parks its outputs in the tile in the layout pass 2 wants to read, pass 2 stores straight
//                   to y (64-byte segments, streaming).
//       store warp  pass 1: 64 x 1 KB cp.async.bulk shared->global of the parked tile, then the release of the
//                   transform's counter; recycles the tile buffer (empty mbarrier).
//   * Intermediate layout W[slot][kb = k/16][c][k%16 ^ (c & 14)]: a pass-1 tile (16 c, all k) is 64 runs of 1 KB, a
//     pass-2 tile (16 k, all c) is one contiguous 64 KB block that lands in shared memory exactly as the compute
//     warps index it ([c][16], lanes along the 16 transforms of the tile).
//
// Every spin loop carries a clock watchdog that traps instead of hanging the GPU.
#pragma once

namespace fused {  // (inside namespace gfb)

// W = transforms per tile (tile = W x 1024 words): 16 -> one 576-thread CTA per SM (3 x 64 KB tiles), 8 -> two 320-thread
// CTAs per SM (3 x 32 KB tiles each) whose phases drift apart, which mixes the ALU / FMA / LSU demand of an SM.
constexpr int kStages = 3;
template <int W>
struct Cfg {
    static constexpr int kComputeThreads = W * 32;
    static constexpr int kThreads = kComputeThreads + 64;  // + load warp + store warp
    static constexpr int kTileWords = W * 1024;
    static constexpr unsigned kTilesPerPass = 1024 / W;
    static constexpr int kTPerWarp = 32 / W;                // sub-transform indices t that share a warp
    static constexpr int kSigmaMask = kTPerWarp - 1;        // exchange swizzle: row (t ^ (k1 & mask))
    static constexpr size_t kSmemTiles = (size_t)kStages * kTileWords * sizeof(uint32_t);
    static constexpr size_t kSmemBytes = kSmemTiles + 1024 * sizeof(uint2) + 256;
    // column permutation inside a W x W block of the intermediate (keeps the parking stores conflict free)
    static __device__ __forceinline__ int f(int w) { return W == 16 ? (w & 14) : (w & 4); }
};
// Pass 2 of transform T is queued `lag` transforms behind its pass 1, and the intermediate lives in a ring of `ring`
// transforms.  Every CTA runs up to kStages tickets ahead of its compute warps, so a few hundred tickets are in flight:
// the lag must cover that window or the load warps sit on the dependency counter, and the ring must cover lag + window
// so pass 1 never waits for a slot.  GFB_NTT_LAG / GFB_NTT_RING override.
constexpr unsigned kLagDefault = 5;
constexpr unsigned kRingDefault = 12;
constexpr unsigned kRingMax = 32;
constexpr long long kWatchdogCycles = 4000000000ll;  // ~2 s at 1.9 GHz

enum : int { TASK_P1 = 0, TASK_P2 = 1, TASK_DONE = 2 };

struct Ctrl {  // device memory, zeroed before every launch; followed by cnt1[batch], cnt2[batch]
    unsigned int ticket;
    unsigned int error;
    unsigned int pad[2];
};

struct Desc {
    int type;
    int tile;
    unsigned transform;
    unsigned slot;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void watchdog_trap(unsigned int* error, unsigned int code) {
    atomicExch(error, code);
    __threadfence_system();
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned int* error, unsigned int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > kWatchdogCycles) watchdog_trap(error, code);
    }
}
__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
template <int THREADS>
__device__ __forceinline__ void compute_barrier() {
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_counter(const unsigned int* cnt, unsigned int want, unsigned int* error,
                                             unsigned int code) {
    if (ld_acquire(cnt) >= want) return;
    const long long t0 = clock64();
    while (ld_acquire(cnt) < want) {
        __nanosleep(64);
        if (clock64() - t0 > kWatchdogCycles) watchdog_trap(error, code);
    }
}

// block number -> (pass, transform): P1(0..L-1), then P1(L) P2(0) P1(L+1) P2(1) ..., then the last L pass-2 blocks
__device__ __forceinline__ void decode_block(unsigned n, unsigned batch, unsigned lag, int& type, unsigned& transform) {  // n = ticket / tiles-per-pass
    const unsigned L = lag < batch ? lag : batch;
    if (n < L) {
        type = TASK_P1;
        transform = n;
        return;
    }
    const unsigned mid = 2u * (batch - L);
    unsigned m = n - L;
    if (m < mid) {
        type = (m & 1u) ? TASK_P2 : TASK_P1;
        transform = (m & 1u) ? (m >> 1) : L + (m >> 1);
        return;
    }
    type = TASK_P2;
    transform = (batch - L) + (m - mid);
}

// ---- the compute rounds ------------------------------------------------------------------------------------------------
// tile[j*W + col]: j = index along the 1024-point transform, col = which of the tile's W transforms.
// Round A: thread (w, t) transforms x[32 a + t], a < 32, multiplies Y[k1] by zeta^(t k1) and parks it at
// row 32 k1 + (t ^ sigma(k1)), sigma(k1) = k1 & (32/W - 1) (the XOR keeps round B's column walk on distinct banks and
// only permutes rows among the t's of one warp).  `col_read` is the column the inputs sit in (pass 2 tiles arrive with
// the columns XOR-permuted by the row index), `w` the column the outputs use.
template <bool LAZY, int W>
__device__ __forceinline__ void round_a(uint32_t* __restrict__ tile, int t, int col_read, int w, const Mod<LAZY>& md,
                                        const Radix32Roots& rt, const uint2* __restrict__ zt_s) {
    constexpr int M = Cfg<W>::kSigmaMask;
    uint32_t v[32];
    const uint32_t* in = tile + t * W + col_read;
#pragma unroll
    for (int a = 0; a < 32; ++a) v[brev5(a)] = in[a * 32 * W];
    radixR<LAZY, 5, true>(v, md, rt);
    __syncwarp();  // the partner lanes (t ^ sigma, any column) have consumed the rows this thread is about to overwrite
    uint32_t* out[M + 1];
#pragma unroll
    for (int sg = 0; sg <= M; ++sg) out[sg] = tile + (t ^ sg) * W + w;
    out[0][0] = LAZY ? Mod<LAZY>::csub(v[0], md.p2) : v[0];
    const uint2* z = zt_s + t;
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) {
        const uint2 zz = z[k1 * 32];
        out[k1 & M][k1 * 32 * W] = md.mul(v[k1], zz.x, zz.y);
    }
}

// Round B loads: thread (w, k1) gathers Y[b], b < 32, of its inner transform k1 (bit-reversed into v)
template <int W>
__device__ __forceinline__ void round_b_load(const uint32_t* __restrict__ tile, int k1, int w, uint32_t (&v)[32]) {
    constexpr int M = Cfg<W>::kSigmaMask;
    const int sg = k1 & M;
    const uint32_t* base = tile + k1 * 32 * W + w;
    const uint32_t* p[M + 1];  // p[r] addresses row (r ^ sigma) of the group of M + 1 rows
#pragma unroll
    for (int r = 0; r <= M; ++r) p[r] = base + (r ^ sg) * W;
#pragma unroll
    for (int b = 0; b < 32; ++b) v[brev5(b)] = p[b & M][(b & ~M) * W];
}

template <bool LAZY, int W>
__global__ void __launch_bounds__(Cfg<W>::kThreads, W == 16 ? 1 : 2)
    k_ntt_fused(Mod<LAZY> md, Radix32Roots rt, const uint2* __restrict__ zt, const uint2* __restrict__ tw4,
                const uint32_t* __restrict__ x, uint32_t* __restrict__ y, uint32_t* __restrict__ ws, Ctrl* ctrl,
                unsigned batch, unsigned ring, unsigned lag) {
    using C = Cfg<W>;
    constexpr int kTileWords = C::kTileWords;
    constexpr unsigned TP = C::kTilesPerPass;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* bufs = reinterpret_cast<uint32_t*>(smem_raw);
    uint2* zt_s = reinterpret_cast<uint2*>(smem_raw + C::kSmemTiles);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + C::kSmemTiles + 1024 * sizeof(uint2));
    uint64_t* full = bars;         // [kStages] tile landed (32 arrivals of the load warp + tx bytes)
    uint64_t* staged = bars + 3;   // [kStages] compute finished with the tile (W*32 arrivals)
    uint64_t* empty = bars + 6;    // [kStages] tile buffer recycled (1 arrival of the store warp)
    Desc* desc = reinterpret_cast<Desc*>(bars + 12);
    unsigned int* cnt1 = reinterpret_cast<unsigned int*>(ctrl + 1);
    unsigned int* cnt2 = cnt1 + batch;
    unsigned int* err = &ctrl->error;
    const int tid = threadIdx.x;
    const unsigned total = batch * 2u * TP;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 32);
            mbar_init(&staged[s], C::kComputeThreads);
            mbar_init(&empty[s], 1);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    for (int i = tid; i < 1024; i += C::kThreads) zt_s[i] = __ldg(zt + i);
    __syncthreads();

    if (tid >= C::kComputeThreads + 32) {
        // ================================================================== store warp
        const int lane = tid & 31;
        for (unsigned it = 0;; ++it) {
            const int b = it % kStages;
            const uint32_t ph = (it / kStages) & 1u;
            mbar_wait(&full[b], ph, err, 0x10u + b);
            const Desc d = desc[b];
            if (d.type == TASK_DONE) break;
            mbar_wait(&staged[b], ph, err, 0x20u + b);
            if (d.type == TASK_P1) {
                // the parked tile is TP blocks of W x W words; block kb goes to W[slot][kb][16 tile .. ][.]
                const uint32_t src = smem_u32(bufs + (size_t)b * kTileWords);
                uint32_t* dst = ws + ((size_t)d.slot << 20) + (size_t)d.tile * (W * W);
#pragma unroll
                for (int h = 0; h < (int)TP / 32; ++h) {
                    const int kb = lane + 32 * h;
                    bulk_s2g(dst + (size_t)kb * (1024 * W), src + kb * (W * W * 4), W * W * 4);
                }
                bulk_commit();
                bulk_wait_all();
                __syncwarp();
                if (lane == 0) {
                    __threadfence();
                    atomicAdd(&cnt1[d.transform], 1u);
                }
            } else if (lane == 0) {
                __threadfence();
                atomicAdd(&cnt2[d.transform], 1u);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[b]);
        }
        return;
    }

    if (tid >= C::kComputeThreads) {
        // ================================================================== load warp
        const int lane = tid & 31;
        for (unsigned it = 0;; ++it) {
            const int b = it % kStages;
            const uint32_t ph = (it / kStages) & 1u;
            mbar_wait(&empty[b], ph ^ 1u, err, 0x30u + b);
            unsigned ticket = 0;
            if (lane == 0) ticket = atomicAdd(&ctrl->ticket, 1u);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            if (ticket >= total) {
                if (lane == 0) desc[b].type = TASK_DONE;
                __syncwarp();
                mbar_arrive(&full[b]);
                break;
            }
            int type;
            unsigned transform;
            decode_block(ticket / TP, batch, lag, type, transform);
            const int tile = (int)(ticket % TP);
            const unsigned slot = transform % ring;
            if (lane == 0) {
                if (type == TASK_P1) {
                    if (transform >= ring) wait_counter(&cnt2[transform - ring], TP, err, 0x40u);
                } else {
                    wait_counter(&cnt1[transform], TP, err, 0x41u);
                    fence_proxy_async();  // the intermediate was written through the async proxy of other SMs
                }
                desc[b] = Desc{type, tile, transform, slot};
                __threadfence_block();
            }
            __syncwarp();
            const uint32_t dst = smem_u32(bufs + (size_t)b * kTileWords);
            if (type == TASK_P1) {
                // 16-byte gathers: W/4 per row of the tile, 8 W per lane, immediate offsets only
                constexpr int PARTS = W / 4, ROWS_PER_IT = 32 / PARTS;
                const int row0 = lane / PARTS, part = lane % PARTS;
                const uint32_t* src = x + ((size_t)transform << 20) + (size_t)tile * W + (size_t)row0 * 1024 + part * 4;
                const uint32_t d0 = dst + (uint32_t)((row0 * W + part * 4) * 4);
#pragma unroll 16
                for (int i = 0; i < 8 * W; ++i)
                    cp_async_16(d0 + (uint32_t)i * (ROWS_PER_IT * W * 4), src + (size_t)i * (ROWS_PER_IT * 1024));
                cp_async_mbar_arrive_noinc(&full[b]);
            } else {
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[b], (uint32_t)(kTileWords * sizeof(uint32_t)));
                    bulk_g2s(dst, ws + ((size_t)slot << 20) + (size_t)tile * kTileWords,
                             (uint32_t)(kTileWords * sizeof(uint32_t)), &full[b]);
                } else {
                    mbar_arrive(&full[b]);
                }
            }
        }
        return;
    }

    // ====================================================================== compute warps
    const int w = tid % W, t = tid / W;  // t = 0..31: sub-transform in round A, inner transform k1 in round B
    for (unsigned it = 0;; ++it) {
        const int b = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        mbar_wait(&full[b], ph, err, 0x50u + b);
        const Desc d = desc[b];
        if (d.type == TASK_DONE) break;
        uint32_t* tile = bufs + (size_t)b * kTileWords;
        uint32_t v[32];
        if (d.type == TASK_P1) {
            round_a<LAZY, W>(tile, t, w, w, md, rt, zt_s);
            compute_barrier<C::kComputeThreads>();
            round_b_load<W>(tile, t, w, v);
            compute_barrier<C::kComputeThreads>();  // every read of the tile is done: it becomes the parking area
            radixR<LAZY, 5, true>(v, md, rt);
            // k = t + 32 k0; multiply by scale * omega^(c k), c = W tile + w, park at [kb = k/W][w][k%W ^ f(w)]
            const uint2* twp = tw4 + (size_t)t * 1024 + (size_t)d.tile * W + w;
            uint32_t* park = tile + (t / W) * (W * W) + w * W + ((t % W) ^ C::f(w));
#pragma unroll
            for (int k0 = 0; k0 < 32; ++k0) {
                const uint2 tw = __ldg(twp + (size_t)k0 * 32768);
                park[k0 * 32 * W] = md.mul(v[k0], tw.x, tw.y);  // lazy: below 2p, which is all pass 2 needs
            }
            fence_proxy_async_smem();  // generic-proxy writes of the tile before the store warp's bulk copies
            mbar_arrive(&staged[b]);
        } else {
            round_a<LAZY, W>(tile, t, w ^ C::f(t % W), w, md, rt, zt_s);
            compute_barrier<C::kComputeThreads>();
            round_b_load<W>(tile, t, w, v);
            radixR<LAZY, 5, true>(v, md, rt);
            // X[k + 1024 k2], k = W tile + w (this tile's rows), k2 = t + 32 k0
            uint32_t* out = y + ((size_t)d.transform << 20) + (size_t)t * 1024 + (size_t)d.tile * W + w;
#pragma unroll
            for (int k0 = 0; k0 < 32; ++k0) __stcs(out + (size_t)k0 * 32768, md.canon(v[k0]));
            // the arrive releases these stores to the store warp at CTA scope; its single gpu-scope fence before the
            // counter increment is cumulative over them (the __syncthreads + thread-0 fence + flag idiom)
            mbar_arrive(&staged[b]);
        }
    }
}

// bytes of workspace for `batch` transforms: ring + control block + two counters per transform
static inline size_t workspace_bytes(size_t batch) {
    const size_t ring = batch < kRingMax ? batch : kRingMax;
    return (ring << 20) * sizeof(uint32_t) + sizeof(Ctrl) + 2 * batch * sizeof(unsigned int) + 64;
}

}  // namespace fused
