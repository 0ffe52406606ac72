// gfb_ewise_tab8.cu — GF(2^m), m <= 8, byte elements: multiply / divide through ONE shared-memory lookup.
//
// The reference's lookup ufunc does three dependent table reads per product (EXP[LOG[a] + LOG[b]],
// /root/reference/src/galois/_domains/_lookup.py:143-168; divide :201-233).  The first CUDA version mirrored that
// with uint16 EXP/LOG tables in shared memory and ran at 0.28 of HBM: three dependent, randomly conflicting 16-bit
// lookups per byte.  Here the whole 256 x 256 product (or quotient) table — 64 KiB, built once per field from the
// same EXP/LOG tables, so every entry is the reference's value — is pulled into shared memory by a single
// cp.async.bulk (TMA) per CTA and each element costs one byte load: index = a << 8 | b.
//   array (op) array : the 64 KiB table;  array (op) scalar / scalar (op) array : its 256-entry row or column.
// Algorithmic bytes per element: 3.  Division flags a zero divisor with the has-zero-byte test
// on the packed operand words.
#include <atomic>
#include <mutex>

#include "gfb_common.cuh"
#include "gfb_ewise_api.cuh"

namespace gfb {

constexpr int kTabThreads = 256;
constexpr int kTabUnroll = 2;              // 128-bit vectors per operand in flight per thread
constexpr size_t kTabBytes = 65536;

// tab[0..65535] = a * b, tab[65536..] = a / b (0 where b == 0), index a << 8 | b;
// tab[131072..] = a * b again at index (a << 8) | (b ^ ((a & 31) << 2)): rows start on different banks, for kernels whose
// lanes hold different constant multipliers a and equal operands b (the decoder's Horner chains all start at b = 0)
__global__ void k_build_tab8(BinaryLut f, uint8_t* tab) {
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. 65535
    const uint32_t a = i >> 8, b = i & 255u;
    uint8_t prod = 0, quot = 0;
    if (a < f.order && b < f.order) {
        prod = (uint8_t)f.mul(a, b);
        if (b != 0u) quot = (uint8_t)(a == 0u ? 0u : f.mul(a, f.inv1(b)));
    }
    tab[i] = prod;
    tab[kTabBytes + i] = quot;
    tab[2 * kTabBytes + ((a << 8) | (b ^ ((a & 31u) << 2)))] = prod;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tab_fetch(uint8_t* tab_s, const uint8_t* tab_g, uint64_t* bar) {
    // one thread arms the barrier and issues the 64 KiB bulk copy; everybody waits on the barrier before the first lookup
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_addr(bar)),
                     "r"((uint32_t)kTabBytes)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(tab_s)),
                     "l"(tab_g), "r"((uint32_t)kTabBytes), "r"(smem_addr(bar))
                     : "memory");
    }
    __syncthreads();  // the barrier is initialised before anyone polls it
}

__device__ __forceinline__ void tab_wait(uint64_t* bar) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_addr(bar))
            : "memory");
    }
}

// four products of the packed bytes of wa and wb
__device__ __forceinline__ uint32_t lookup4(const uint8_t* __restrict__ tab, uint32_t wa, uint32_t wb) {
    const uint32_t i01 = __byte_perm(wb, wa, 0x5140);  // [b0, a0, b1, a1]: two 16-bit indices a << 8 | b
    const uint32_t i23 = __byte_perm(wb, wa, 0x7362);  // [b2, a2, b3, a3]
    const uint32_t r0 = tab[i01 & 0xFFFFu], r1 = tab[i01 >> 16];
    const uint32_t r2 = tab[i23 & 0xFFFFu], r3 = tab[i23 >> 16];
    return __byte_perm(__byte_perm(r0, r1, 0x0040), __byte_perm(r2, r3, 0x0040), 0x5410);
}

__device__ __forceinline__ bool has_zero_byte(uint32_t w) { return ((w - 0x01010101u) & ~w & 0x80808080u) != 0u; }

// out[i] = tab[a[i] << 8 | b[i]]
template <bool DIV>
__global__ void __launch_bounds__(kTabThreads)
    k_tab8_aa(const uint8_t* __restrict__ tab_g, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
              uint8_t* __restrict__ out, size_t n, int* err) {
    extern __shared__ __align__(128) uint8_t tab_s[];
    __shared__ uint64_t bar;
    tab_fetch(tab_s, tab_g + (DIV ? kTabBytes : 0), &bar);
    const size_t nvec = n / 16;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint4* a4 = reinterpret_cast<const uint4*>(a);
    const uint4* b4 = reinterpret_cast<const uint4*>(b);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    bool zero_seen = false;
    bool ready = false;
    for (; i < nvec; i += kTabUnroll * stride) {
        uint4 pa[kTabUnroll], pb[kTabUnroll];
#pragma unroll
        for (int u = 0; u < kTabUnroll; ++u) {
            const size_t j = i + u * stride;
            if (j < nvec) {
                pa[u] = __ldcs(a4 + j);
                pb[u] = __ldcs(b4 + j);
            }
        }
        if (!ready) {  // the first operand loads are already in flight while the table streams in
            tab_wait(&bar);
            ready = true;
        }
#pragma unroll
        for (int u = 0; u < kTabUnroll; ++u) {
            const size_t j = i + u * stride;
            if (j < nvec) {
                if (DIV)
                    zero_seen |= has_zero_byte(pb[u].x) | has_zero_byte(pb[u].y) | has_zero_byte(pb[u].z) |
                                 has_zero_byte(pb[u].w);
                uint4 r;
                r.x = lookup4(tab_s, pa[u].x, pb[u].x);
                r.y = lookup4(tab_s, pa[u].y, pb[u].y);
                r.z = lookup4(tab_s, pa[u].z, pb[u].z);
                r.w = lookup4(tab_s, pa[u].w, pb[u].w);
                __stcs(o4 + j, r);
            }
        }
    }
    if (!ready) tab_wait(&bar);
    // tail (n % 16 elements): first threads of the first CTA
    const size_t t = nvec * 16 + threadIdx.x;
    if (blockIdx.x == 0 && t < n) {
        const uint32_t x = a[t], y = b[t];
        if (DIV && y == 0u) zero_seen = true;
        out[t] = tab_s[(x << 8) | y];
    }
    if (zero_seen) *err = 1;
}

// out[i] = row[a[i]] with row[v] = tab[v << 8 | s] (array (op) scalar) or tab[s << 8 | v] (scalar (op) array)
template <bool DIV>
__global__ void __launch_bounds__(kTabThreads)
    k_tab8_scalar(const uint8_t* __restrict__ tab_g, const uint8_t* __restrict__ a, uint32_t s, int scalar_on_left,
                  uint8_t* __restrict__ out, size_t n, int* err) {
    __shared__ __align__(16) uint8_t row[256];
    const uint8_t* tg = tab_g + (DIV ? kTabBytes : 0);
    row[threadIdx.x] = scalar_on_left ? tg[(s << 8) | threadIdx.x] : tg[((uint32_t)threadIdx.x << 8) | s];
    __syncthreads();
    bool zero_seen = DIV && !scalar_on_left && s == 0u;
    const size_t nvec = n / 16;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint4* a4 = reinterpret_cast<const uint4*>(a);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    auto look = [&](uint32_t w) {
        const uint32_t r0 = row[w & 255u], r1 = row[(w >> 8) & 255u], r2 = row[(w >> 16) & 255u], r3 = row[w >> 24];
        return r0 | (r1 << 8) | (r2 << 16) | (r3 << 24);
    };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 p = __ldcs(a4 + i);
        if (DIV && scalar_on_left) zero_seen |= has_zero_byte(p.x) | has_zero_byte(p.y) | has_zero_byte(p.z) | has_zero_byte(p.w);
        uint4 r;
        r.x = look(p.x);
        r.y = look(p.y);
        r.z = look(p.z);
        r.w = look(p.w);
        __stcs(o4 + i, r);
    }
    const size_t t = nvec * 16 + threadIdx.x;
    if (blockIdx.x == 0 && t < n) {
        const uint32_t x = a[t];
        if (DIV && scalar_on_left && x == 0u) zero_seen = true;
        out[t] = row[x];
    }
    if (zero_seen) *err = 1;
}

// build (once per field, on `st`) and return the device table
gfb_status ensure_tab8(gfb_field* f, cudaStream_t st) {
    if (f->d_tab8) return GFB_OK;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (f->d_tab8) return GFB_OK;
    uint8_t* tab = nullptr;
    GFB_CUDA_TRY(cudaMalloc((void**)&tab, 3 * kTabBytes));
    BinaryLut lut = f->bin_lut();
    k_build_tab8<<<256, 256, lut.smem_bytes(), st>>>(lut, tab);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFree(tab);
        set_last_error("k_build_tab8", e);
        return GFB_ERR_CUDA;
    }
    // the table is produced and consumed on the caller's stream; other streams of the same field synchronise here once
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(tab);
        set_last_error("k_build_tab8 sync", e);
        return GFB_ERR_CUDA;
    }
    f->d_tab8 = tab;
    return GFB_OK;
}

bool tab8_applies(const gfb_field* f, int op, int eb, const void* a, const void* b, const void* out) {
    if (f->impl != IMPL_BIN_LUT || f->degree > 8 || eb != 1) return false;
    if (op != GFB_MUL && op != GFB_DIV) return false;
    return ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)out)) & 15) == 0;  // 128-bit operand vectors
}

gfb_status ewise_binary_tab8(gfb_field* f, int op, int mode, const void* a, const void* b, uint64_t s, void* out, size_t n,
                             cudaStream_t st) {
    gfb_status rc = ensure_tab8(f, st);
    if (rc != GFB_OK) return rc;
    const size_t nvec = n / 16;
    size_t blocks = (nvec + (size_t)kTabThreads * kTabUnroll - 1) / ((size_t)kTabThreads * kTabUnroll);
    if (blocks == 0) blocks = 1;
    const uint8_t* A = (const uint8_t*)a;
    uint8_t* O = (uint8_t*)out;
    if (mode == MODE_AA) {
        const size_t cap = (size_t)sm_count() * 3;  // three 64 KiB tables fit one SM
        if (blocks > cap) blocks = cap;
        // the 64 KiB opt-in is a per-device function attribute: set it once per device (the call costs microseconds of host
        // time, which is what a 1e6-element launch is made of)
        static std::atomic<unsigned long long> attr_done{0};
        const unsigned long long bit = 1ull << (f->device & 63);
        if (!(attr_done.load(std::memory_order_acquire) & bit)) {
            GFB_CUDA_TRY(cudaFuncSetAttribute(k_tab8_aa<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTabBytes));
            GFB_CUDA_TRY(cudaFuncSetAttribute(k_tab8_aa<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTabBytes));
            attr_done.fetch_or(bit, std::memory_order_release);
        }
        if (op == GFB_MUL)
            k_tab8_aa<false><<<(unsigned)blocks, kTabThreads, kTabBytes, st>>>(f->d_tab8, A, (const uint8_t*)b, O, n, f->d_err);
        else
            k_tab8_aa<true><<<(unsigned)blocks, kTabThreads, kTabBytes, st>>>(f->d_tab8, A, (const uint8_t*)b, O, n, f->d_err);
    } else {
        const size_t cap = (size_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        const int left = mode == MODE_SA;
        if (op == GFB_MUL)
            k_tab8_scalar<false><<<(unsigned)blocks, kTabThreads, 0, st>>>(f->d_tab8, A, (uint32_t)s, left, O, n, f->d_err);
        else
            k_tab8_scalar<true><<<(unsigned)blocks, kTabThreads, 0, st>>>(f->d_tab8, A, (uint32_t)s, left, O, n, f->d_err);
    }
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

}  // namespace gfb
