// gfb_rs_fast.cuh — errors-and-erasures decoder for codes with at most 32 parity symbols over GF(2^m), m <= 8
// (RS(255,223), RS(255,239), RS(63,55) ..., binary BCH with d - 1 <= 32), included by gfb_rs.cu.
//
// Same algorithm and the same decisions as bch_decode_jit.implementation
// (/root/reference/src/galois/_codes/_bch.py:1337-1578; Berlekamp-Massey _lfsr.py:1647-1702), restructured so that a
// warp spends its issue slots on field operations instead of table plumbing (round 1: 9.3 k warp instructions per
// 16-error codeword, half of them in Berlekamp-Massey's branches and log/antilog lookups):
//   * every product is ONE shared-memory byte load from the field's 256 x 256 product table (64 KiB per CTA, pulled in
//     with one cp.async.bulk, shared by 16 warps): no LOG/EXP round trip, no zero tests (T[0][x] = 0);
//   * polynomials live in registers, coefficient i in lane i (degree 32 terms in a warp-uniform register); shifts and
//     broadcasts are warp shuffles;
//   * syndromes: the received word is first reduced modulo the generator polynomial, rho = r mod g (an LFSR: lane j holds
//     rho_j, per symbol one shuffle pair and ONE conflict-free load from the per-code table G[f][j] = f g_j), and the
//     syndromes are S_l = rho(alpha^(c+l)) — 32-term Horner chains instead of 255-term ones (g(alpha^(c+l)) = 0).
//     Codes with more than 32 parity symbols per 32 roots (binary BCH) keep the direct Horner evaluation;
//   * the product table is stored with row a rotated by (a mod 32) banks, so lanes holding different constant
//     multipliers and equal operands (every Horner chain starts at 0) hit different banks;
//   * Berlekamp-Massey in the inversion-free form  C <- b C - d x^m B  (C stays a non-zero scalar multiple of the
//     reference's connection polynomial: same L, same degree, same roots, and Forney's quotient Omega / Lambda' is scale
//     invariant), branch-free apart from warp-uniform selects;
//   * Chien search: each lane walks 8 positions with 8 independent Horner chains.
// Decoder failures (more than d-1 erasures, 2v + u > d - 1, wrong root count) return -1 and the received row, as the
// reference does.
#pragma once

namespace fastrs {  // (inside namespace gfb)

constexpr int kWarps = 16;
constexpr int kThreadsPerCta = kWarps * 32;
constexpr int kTab = 65536;
constexpr int kRecStride = 256;
constexpr int kGTab = 256 * 32;  // G[f][j] = f * g_j for the remainder LFSR
constexpr size_t kSmemBytes = kTab + kGTab + 512 + 256 + (size_t)kWarps * (kRecStride + 64 + 64);

__device__ __forceinline__ uint32_t sm_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// row offset of multiplier a in the bank-skewed product table: entry a * b sits at row(a) ^ b
__device__ __forceinline__ uint32_t row(uint32_t a) { return (a << 8) ^ ((a & 31u) << 2); }

struct Tables {
    const uint8_t* T;    // product table, a * b at row(a) ^ b
    const uint8_t* exp;  // alpha0^i, i < 2 q1 (alpha0 = the field's primitive element)
    const uint8_t* inv;  // multiplicative inverses, inv[0] = 0
    int q1;
    __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) const { return T[row(a) ^ b]; }
    // alpha0^(la * x mod q1) for any integer x
    __device__ __forceinline__ uint32_t pow_alpha(int la, int x) const {
        int r = (int)(((long long)la * x) % q1);
        if (r < 0) r += q1;
        return exp[r];
    }
};

__device__ __forceinline__ uint32_t warp_xor32(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// value of lane (lane - by), 0 for the first `by` lanes
__device__ __forceinline__ uint32_t shift_up(uint32_t v, int by, int lane) {
    const uint32_t s = __shfl_up_sync(0xffffffffu, v, by);
    return lane >= by ? s : 0u;
}

__global__ void __launch_bounds__(kThreadsPerCta, 2)
    k_rs_decode_fast(RsDev rs, const uint8_t* __restrict__ tab_g, const uint8_t* __restrict__ received,
                     const uint8_t* __restrict__ erasures, int ns, uint8_t* __restrict__ decoded,
                     long long* __restrict__ n_errors, size_t N) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    uint8_t* T = smem;
    uint8_t* G = smem + kTab;
    uint8_t* s_exp = G + kGTab;
    uint8_t* s_inv = s_exp + 512;
    uint8_t* s_warp = s_inv + 256;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* rec = s_warp + warp * (kRecStride + 128);  // received word, ascending degree, erasures zeroed
    uint8_t* pos = rec + kRecStride;                     // root positions found by the Chien search
    uint8_t* epos = pos + 64;                            // erasure positions
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm_addr(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(sm_addr(&bar)),
                     "r"((uint32_t)kTab)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm_addr(T)),
                     "l"(tab_g), "r"((uint32_t)kTab), "r"(sm_addr(&bar))
                     : "memory");
    }
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_exp[i] = i < 2 * rs.order ? rs.exp8[i] : 0;
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_inv[i] = (i > 0 && i < rs.order) ? rs.exp8[rs.q1 - rs.log8[i]] : 0;
    __syncthreads();
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(sm_addr(&bar))
                : "memory");
        }
    }
    const Tables gf{T, s_exp, s_inv, rs.q1};
    const int d1 = rs.d - 1;  // <= 32
    const int R = rs.n - rs.k;                  // parity symbols = degree of g(x)
    const bool lfsr = R <= 32 && rs.gen != nullptr;
    if (lfsr) {
        // G[f][j] = f * g_j, g_j = coefficient of x^j of the monic generator polynomial (rs.gen is degree-descending)
        for (int i = threadIdx.x; i < kGTab; i += blockDim.x) {
            const uint32_t f = i >> 5, j = i & 31;
            G[i] = (int)j < R ? T[row(f) ^ rs.gen[R - j]] : 0;
        }
        __syncthreads();
    }
    const int la = rs.log_alpha;
    const int q1 = rs.q1;
    const size_t warps_total = (size_t)gridDim.x * kWarps;

    for (size_t cwi = (size_t)blockIdx.x * kWarps + warp; cwi < N; cwi += warps_total) {
        const uint8_t* rrow = received + cwi * ns;
        const uint8_t* erow = erasures ? erasures + cwi * ns : nullptr;
        uint8_t* drow = decoded + cwi * ns;
        __syncwarp();
        // ---- load (received ascending degree, erasures zeroed, erasure positions ascending: _bch.py:1344-1355)
        int u = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = lane + 32 * k;
            uint32_t v = 0, er = 0;
            if (i < ns) {
                v = rrow[ns - 1 - i];
                er = erow ? (erow[ns - 1 - i] != 0) : 0u;
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, er);
            if (er) {
                const int slot = u + __popc(bal & ((1u << lane) - 1u));
                if (slot < 64) epos[slot] = (uint8_t)i;
                v = 0;
            }
            u += __popc(bal);
            rec[i] = (uint8_t)v;
        }
        __syncwarp();
        long long result = 0;
        bool fail = u > d1;  // :1357-1360
        bool clean = false;
        uint32_t S = 0;
        if (!fail) {
            // ---- syndromes S_l = r(alpha^(c+l)), l = lane < d1 (:1370)
            const int lroot = (int)(((long long)la * (rs.c + lane)) % q1);  // log of the root alpha^(c + l)
            const uint32_t xg = row(gf.exp[lroot]);
            if (lfsr) {
                // rho(x) = r(x) mod g(x): rho <- rho x + r_i (mod g), symbols in descending degree
                uint32_t rho = 0;
                const int top = R - 1;
#pragma unroll 4
                for (int i = ns - 1; i >= 0; --i) {
                    const uint32_t f = __shfl_sync(0xffffffffu, rho, top);
                    const uint32_t up = __shfl_up_sync(0xffffffffu, rho, 1);
                    rho = (lane == 0 ? (uint32_t)rec[i] : up) ^ (uint32_t)G[f * 32 + lane];
                }
                // S_l = rho(alpha^(c+l)): Horner over the R <= 32 coefficients
                uint32_t acc = 0;
                for (int j = top; j >= 0; --j) acc = (uint32_t)T[xg ^ acc] ^ __shfl_sync(0xffffffffu, rho, j);
                S = acc;
            } else {
                // direct evaluation, two Horner chains over the halves of the word: r(x) = A(x) x^h + B(x)
                const int h = ns >> 1;
                uint32_t a = 0, bq = 0;
                int i = ns - 1;
                if (ns & 1) {
                    a = rec[i];
                    --i;
                }
                for (; i >= h; --i) {
                    a = (uint32_t)T[xg ^ a] ^ rec[i];
                    bq = (uint32_t)T[xg ^ bq] ^ rec[i - h];
                }
                S = gf.mul(a, gf.exp[(int)(((long long)lroot * h) % q1)]) ^ bq;
            }
            if (lane >= d1) S = 0;
            const uint32_t any = __ballot_sync(0xffffffffu, S != 0);
            if (any == 0 && u == 0) clean = true;  // :1373-1376
        }
        int v = 0, v_total = 0;
        if (!fail && !clean) {
            // ---- Gamma(x) = prod (1 - Y_k x) (:1389-1393): coefficient i in lane i, coefficient 32 in Gtop
            uint32_t Gm = lane == 0 ? 1u : 0u, Gtop = 0;
            for (int k = 0; k < u; ++k) {
                const uint32_t Yrow = row(gf.pow_alpha(la, (int)epos[k]));
                const uint32_t g31 = __shfl_sync(0xffffffffu, Gm, 31);
                Gtop ^= T[Yrow ^ g31];
                Gm ^= T[Yrow ^ shift_up(Gm, 1, lane)];
            }
            // ---- S'(x) = Gamma(x) S(x) mod x^(d-1) (:1408-1409)
            uint32_t Sp = S;
            if (u > 0) {
                Sp = 0;
                for (int i = 0; i <= u && i < 32; ++i) {
                    const uint32_t g = __shfl_sync(0xffffffffu, Gm, i);
                    Sp ^= T[row(g) ^ shift_up(S, i, lane)];
                }
                if (lane >= d1) Sp = 0;
            }
            // ---- Berlekamp-Massey on S'[u:], inversion-free (:1421-1428, _lfsr.py:1647-1702)
            uint32_t C = lane == 0 ? 1u : 0u;  // connection polynomial (scaled)
            uint32_t Bs = lane == 1 ? 1u : 0u;  // x^m B
            uint32_t Sw = 0;                    // Sw[i] = Sq[n - i]
            uint32_t b = 1;
            int L = 0;
            const int ls = d1 - u;
            for (int nn = 0; nn < ls; ++nn) {
                const uint32_t s_new = __shfl_sync(0xffffffffu, Sp, (u + nn) & 31);
                const uint32_t up = __shfl_up_sync(0xffffffffu, Sw, 1);
                Sw = lane == 0 ? s_new : up;
                const uint32_t dsc = warp_xor32(T[row(C) ^ Sw]);
                const uint32_t Cnew = (uint32_t)T[row(b) ^ C] ^ (uint32_t)T[row(dsc) ^ Bs];
                const bool grow = dsc != 0 && 2 * L <= nn;
                const uint32_t src = grow ? C : Bs;  // B <- old C (m = 1) or m += 1
                if (dsc != 0) C = Cnew;
                if (grow) {
                    L = nn + 1 - L;
                    b = dsc;
                }
                Bs = shift_up(src, 1, lane);
            }
            // Lambda = C[:L+1] with high-degree zeros trimmed
            const int len = L + 1 < ls ? L + 1 : ls;
            const uint32_t nzmask = __ballot_sync(0xffffffffu, lane < len && C != 0);
            const int lam_len = nzmask ? 32 - __clz(nzmask) : 1;
            if (lane >= lam_len) C = 0;
            if (u >= d1) C = lane == 0 ? 1u : 0u;  // no syndromes left for BM: Lambda = 1
            v = (u < d1 ? lam_len : 1) - 1;
            if (2 * v + u > d1) fail = true;  // :1431-1433
            if (!fail) {
                // ---- Lambda_total = Gamma Lambda (:1449), degree u + v <= d - 1 <= 32
                uint32_t Lt = C, Lt_top = 0;
                if (u > 0) {
                    Lt = 0;
                    for (int i = 0; i <= u; ++i) {
                        const uint32_t g = i < 32 ? __shfl_sync(0xffffffffu, Gm, i) : Gtop;
                        if (i < 32) Lt ^= T[row(g) ^ shift_up(C, i, lane)];
                        if (i >= 1) Lt_top ^= T[row(g) ^ __shfl_sync(0xffffffffu, C, (32 - i) & 31)];
                    }
                }
                const int lt_len = u + v + 1;
                // ---- Chien search over i = 0 .. n-1: Lambda_total(alpha^-i) == 0 (:1464-1475), 8 positions per lane
                uint32_t xr[8], acc[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = lane + 32 * k;
                    int lx = (int)(((long long)la * i) % q1);
                    lx = lx ? q1 - lx : 0;
                    xr[k] = row(gf.exp[lx]);
                    acc[k] = 0;
                }
                for (int j = lt_len - 1; j >= 0; --j) {
                    const uint32_t cj = j < 32 ? __shfl_sync(0xffffffffu, Lt, j) : Lt_top;
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] = (uint32_t)T[xr[k] ^ acc[k]] ^ cj;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = lane + 32 * k;
                    const bool root = i < rs.n && i < ns && acc[k] == 0;  // roots at i >= ns are not recorded (:1469-1472)
                    const uint32_t bal = __ballot_sync(0xffffffffu, root);
                    if (root) {
                        const int slot = v_total + __popc(bal & ((1u << lane) - 1u));
                        if (slot < 64) pos[slot] = (uint8_t)i;
                    }
                    v_total += __popc(bal);
                }
                __syncwarp();
                if (v_total != v + u) fail = true;  // :1479-1481
                if (!fail) {
                    // ---- Omega' = Lambda S' mod x^(d-1) (:1498-1499)
                    uint32_t Om = 0;
                    for (int i = 0; i <= v; ++i) {
                        const uint32_t li = __shfl_sync(0xffffffffu, C, i);
                        Om ^= T[row(li) ^ shift_up(Sp, i, lane)];
                    }
                    if (lane >= d1) Om = 0;
                    // ---- Forney (:1512-1556): lane r handles root r.  E = Omega'(X^-1) / Lambda_total'(X^-1) (X^-1)^(c-1);
                    //      in characteristic 2 the derivative keeps the odd-degree terms.
                    const int my = lane < v_total ? pos[lane] : 0;
                    int lx = (int)(((long long)la * my) % q1);
                    lx = lx ? q1 - lx : 0;
                    const uint32_t xrow = row(gf.exp[lx]);
                    uint32_t num = 0, den = 0;
                    for (int j = d1 - 1; j >= 0; --j) num = (uint32_t)T[xrow ^ num] ^ __shfl_sync(0xffffffffu, Om, j);
                    for (int j = lt_len - 1; j >= 1; --j) {
                        const uint32_t cj = j < 32 ? __shfl_sync(0xffffffffu, Lt, j) : Lt_top;
                        den = (uint32_t)T[xrow ^ den] ^ ((j & 1) ? cj : 0u);
                    }
                    if (lane < v_total) {
                        uint32_t E = 0;
                        if (den != 0) {
                            E = gf.mul(num, gf.inv[den]);
                            int ex = (int)(((long long)lx * (rs.c - 1)) % q1);
                            if (ex < 0) ex += q1;
                            E = gf.mul(E, gf.exp[ex]);
                        }
                        // ---- correct (:1567-1576): xor for RS; GF(2)'s subtract_modular rule for binary BCH
                        if (rs.base2) {
                            const int a = rec[my], bb = (int)E;
                            rec[my] = (uint8_t)(a >= bb ? a - bb : 2 + a - bb);
                        } else {
                            rec[my] ^= (uint8_t)E;
                        }
                    }
                    __syncwarp();
                    result = v;
                }
            }
        }
        if (fail) result = -1;
        // ---- store: corrected word, or the received row untouched when nothing was (or could be) corrected
        if (fail || clean) {
            for (int i = lane; i < ns; i += 32) drow[i] = rrow[i];
        } else {
            for (int i = lane; i < ns; i += 32) drow[ns - 1 - i] = rec[i];
        }
        if (lane == 0) n_errors[cwi] = result;
    }
}

}  // namespace fastrs
