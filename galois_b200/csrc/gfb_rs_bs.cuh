// gfb_rs_bs.cuh — bit-sliced errors-and-erasures decoder: 32 codewords per warp, codes with at most 32 parity symbols over
// GF(2^m), m <= 8 (included by gfb_rs.cu).
//
// Same algorithm and the same decisions as bch_decode_jit.implementation
// (/root/reference/src/galois/_codes/_bch.py:1337-1578; Berlekamp-Massey _lfsr.py:1647-1702).  ncu on the two
// warp-per-codeword decoders (profiles/r02_summary.md) showed where the time goes: ~12 k field multiply-adds per codeword sit
// in the syndromes (d-1 Horner chains over n symbols) and the Chien search (n Horner chains over the locator), every one of
// them a multiplication by a CONSTANT (a root alpha^(c+l), a position alpha^-i), and a table lookup per product costs either
// ~7 instructions (log / antilog plumbing) or a 3.4-way bank-conflicted byte load.  A constant multiply over GF(2^m) is an
// m x m bit matrix, so with the symbols of 32 codewords held as m bit-planes (bit k of plane j = bit j of codeword k's
// symbol) one Horner step for 32 codewords is 64 LOP3 (`acc_b & mask_jb ^ ...`) and 8 broadcast loads — no tables, no
// conflicts, 0.08 instructions per multiply-add.
//   1. load: lane k reads codeword k's symbols (erasures zeroed) four to a word; a 32 x 32 bit transpose across the warp
//      (five shuffle stages) turns each word into the planes R[i][j] of four positions; the received rows are copied to the
//      output as they are (16-byte accesses);
//   2. syndromes: lane l holds the 64 masks of alpha^(c+l) and runs Horner over the planes; the 8 result planes are
//      transposed back to one byte per (codeword, syndrome);
//   3. key equation, ONE THREAD PER CODEWORD, products through the bank-skewed product table of gfb_rs_fast.cuh, polynomials
//      byte-packed in registers: Gamma and S' = Gamma S when the group has erasures, inversion-free Berlekamp-Massey (17
//      coefficients: a longer connection polynomial is redone by the warp-per-codeword formulation of gfb_rs_fast.cuh),
//      Lambda_total = Gamma Lambda, Omega' = Lambda S' — 32 codewords per warp in lock step, no cross-lane reductions;
//   4. Chien search: Lambda_total of the 32 codewords as planes; lane = position, 8 rounds of 32 positions, Horner with the
//      masks of alpha^-i; a zero across the 8 result planes marks a root of codeword k in bit k;
//   5. one thread per codeword: root list from the (transposed) bit masks, Forney for each root with the Horner chains cut at
//      the group's highest degree, corrected symbols patched into the
//      output row.
// Decoder failures (more than d-1 erasures, 2v + u > d - 1, wrong root count) return -1 and leave the received row.
#pragma once

namespace bsrs {  // (inside namespace gfb)

constexpr int kWarps = 12;
constexpr int kThreadsPerCta = kWarps * 32;
constexpr int kTab = 65536;
constexpr int kOffPlanes = 0;                  // uint32 R[256][8]
constexpr int kOffMask = 8192;                 // uint32 [256]: erasure mask per position, later root mask per position
constexpr int kRow = 36;                       // bytes per codeword row (9 words: a lane reads its row without conflicts)
constexpr int kOffS = kOffMask + 1024;         // uint8 S[32][36], later Omega'[32][36]
constexpr int kOffLt = kOffS + 32 * kRow;      // uint8 Lambda_total[32][36] (33 used)
constexpr int kOffLP = kOffLt + 32 * kRow;     // uint32 LP[33][8]: coefficient planes for the Chien search; before that the
                                               // erasure positions, after it the root positions of each codeword ([32][33])
constexpr int kOffMeta = kOffLP + 1056;        // int32 meta[32][4] (state, u, v, lt_len)
constexpr int kWarpBytes = kOffMeta + 512;
constexpr size_t kSmemBytes = kTab + 512 + 256 + (size_t)kWarps * kWarpBytes;
static_assert(kSmemBytes <= 232448, "bit-sliced decoder exceeds the shared memory of an SM");

enum : int { ST_FAIL = 0, ST_CLEAN = 1, ST_OK = 2, ST_SLOW = 3 };

using fastrs::row;
using fastrs::shift_up;
using fastrs::sm_addr;
using fastrs::warp_xor32;

// masks of the constant multiplier x: mk[j * 8 + b] = all-ones iff bit j of x * 2^b is set
__device__ __forceinline__ void build_masks(const uint8_t* __restrict__ T, uint32_t x, uint32_t (&mk)[64]) {
    const uint32_t xr = row(x);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const uint32_t cb = T[xr ^ (1u << b)];
#pragma unroll
        for (int j = 0; j < 8; ++j) mk[j * 8 + b] = 0u - ((cb >> j) & 1u);
    }
    // keep the masks as opaque register values: left to itself the compiler re-derives them as predicates inside the Horner
    // loop (one SEL plus one LOP3 per term instead of one LOP3)
#pragma unroll
    for (int i = 0; i < 64; ++i) asm volatile("" : "+r"(mk[i]));
}

// the aligned 32-bit word at p, with the bytes outside [lo, hi) read as zero (only the first and the last word of a buffer)
__device__ __forceinline__ uint32_t load_word_within(const uint8_t* p, const uint8_t* lo, const uint8_t* hi) {
    if (p >= lo && p + 4 <= hi) return __ldg(reinterpret_cast<const uint32_t*>(p));
    uint32_t w = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t)
        if (p + t >= lo && p + t < hi) w |= (uint32_t)p[t] << (8 * t);
    return w;
}

// 32 x 32 bit transpose across the warp: on return bit k of lane m is bit m of lane k's input
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t m0 = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
        const bool hi = (lane & s) != 0;
        const uint32_t keep = hi ? ~m0 : m0;
        const uint32_t t = hi ? (y >> s) : (y << s);
        x = (x & keep) | (t & ~keep);
    }
    return x;
}

// acc <- x * acc + add on 8 bit-planes (32 independent field elements)
__device__ __forceinline__ void horner_step(const uint32_t (&mk)[64], uint32_t (&acc)[8], const uint32_t* __restrict__ add) {
    uint32_t nw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t v = add[j];
#pragma unroll
        for (int b = 0; b < 8; ++b) v ^= acc[b] & mk[j * 8 + b];
        nw[j] = v;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = nw[j];
}

// ---- byte-packed polynomials in registers (coefficient i = byte i of a little-endian word array) -----------------------
template <int W>
__device__ __forceinline__ uint32_t byte_of(const uint32_t (&w)[W], int i) {  // i is a compile-time constant at every call site
    return (w[i >> 2] >> (8 * (i & 3))) & 255u;
}
template <int W>
__device__ __forceinline__ void shift_in_low(uint32_t (&w)[W], uint32_t in) {  // p(x) <- p(x) x + in
#pragma unroll
    for (int j = W - 1; j >= 1; --j) w[j] = __funnelshift_l(w[j - 1], w[j], 8);
    w[0] = (w[0] << 8) | in;
}
template <int W>
__device__ __forceinline__ void shift_out_low(uint32_t (&w)[W]) {  // drop coefficient 0
#pragma unroll
    for (int j = 0; j < W - 1; ++j) w[j] = __funnelshift_r(w[j], w[j + 1], 8);
    w[W - 1] >>= 8;
}

// Warp-per-codeword key equation for codeword k of the group (the general path: erasures, long locators).  Reads the
// syndromes from Ss[k], writes Lambda_total to LtS[k], Omega' to Ss[k] and the codeword's state to meta[k].
__device__ __forceinline__ void key_equation_warp(const fastrs::Tables& gf, const uint8_t* __restrict__ T, int k, int lane, int d1,
                                                  int la, int ns, bool any_era, const uint32_t* __restrict__ PM,
                                                  uint8_t* __restrict__ Ss, uint8_t* __restrict__ LtS, uint8_t* __restrict__ epos,
                                                  int* __restrict__ meta) {
    const uint32_t S = Ss[k * kRow + lane];
    int u = 0;
    if (any_era) {  // erasure positions of codeword k, ascending
#pragma unroll 1
        for (int r = 0; r < 8; ++r) {
            const int i = lane + 32 * r;
            const uint32_t er = (i < ns) ? ((PM[i] >> k) & 1u) : 0u;
            const uint32_t bal = __ballot_sync(0xffffffffu, er);
            if (er) {
                const int slot = u + __popc(bal & ((1u << lane) - 1u));
                if (slot < 64) epos[slot] = (uint8_t)i;
            }
            u += __popc(bal);
        }
        __syncwarp();
    }
    int state = ST_OK, v = 0, lt_len = 1;
    uint32_t Lt = 0, Lt_top = 0, Om = 0;
    if (u > d1) {  // :1357-1360
        state = ST_FAIL;
    } else if (__ballot_sync(0xffffffffu, S != 0) == 0 && u == 0) {
        state = ST_CLEAN;  // :1373-1376
    } else {
        // Gamma(x) = prod (1 - Y_k x) (:1389-1393): coefficient i in lane i, coefficient 32 in Gtop
        uint32_t Gm = lane == 0 ? 1u : 0u, Gtop = 0;
        for (int e = 0; e < u; ++e) {
            const uint32_t Yrow = row(gf.pow_alpha(la, (int)epos[e]));
            const uint32_t g31 = __shfl_sync(0xffffffffu, Gm, 31);
            Gtop ^= T[Yrow ^ g31];
            Gm ^= T[Yrow ^ shift_up(Gm, 1, lane)];
        }
        // S'(x) = Gamma(x) S(x) mod x^(d-1) (:1408-1409)
        uint32_t Sp = S;
        if (u > 0) {
            Sp = 0;
            for (int i = 0; i <= u && i < 32; ++i) {
                const uint32_t g = __shfl_sync(0xffffffffu, Gm, i);
                Sp ^= T[row(g) ^ shift_up(S, i, lane)];
            }
            if (lane >= d1) Sp = 0;
        }
        // Berlekamp-Massey on S'[u:], inversion-free (:1421-1428, _lfsr.py:1647-1702)
        uint32_t C = lane == 0 ? 1u : 0u, Bs = lane == 1 ? 1u : 0u, Sw = 0, b = 1;
        int L = 0;
        const int ls = d1 - u;
        for (int nn = 0; nn < ls; ++nn) {
            const uint32_t s_new = __shfl_sync(0xffffffffu, Sp, (u + nn) & 31);
            const uint32_t up = __shfl_up_sync(0xffffffffu, Sw, 1);
            Sw = lane == 0 ? s_new : up;
            const uint32_t dsc = warp_xor32(T[row(C) ^ Sw]);
            const uint32_t Cnew = (uint32_t)T[row(b) ^ C] ^ (uint32_t)T[row(dsc) ^ Bs];
            const bool grow = dsc != 0 && 2 * L <= nn;
            const uint32_t src = grow ? C : Bs;
            if (dsc != 0) C = Cnew;
            if (grow) {
                L = nn + 1 - L;
                b = dsc;
            }
            Bs = shift_up(src, 1, lane);
        }
        const int len = L + 1 < ls ? L + 1 : ls;
        const uint32_t nzmask = __ballot_sync(0xffffffffu, lane < len && C != 0);
        const int lam_len = nzmask ? 32 - __clz(nzmask) : 1;
        if (lane >= lam_len) C = 0;
        if (u >= d1) C = lane == 0 ? 1u : 0u;
        v = (u < d1 ? lam_len : 1) - 1;
        if (2 * v + u > d1) {  // :1431-1433
            state = ST_FAIL;
        } else {
            Lt = C;  // Lambda_total = Gamma Lambda (:1449)
            if (u > 0) {
                Lt = 0;
                for (int i = 0; i <= u; ++i) {
                    const uint32_t g = i < 32 ? __shfl_sync(0xffffffffu, Gm, i) : Gtop;
                    if (i < 32) Lt ^= T[row(g) ^ shift_up(C, i, lane)];
                    if (i >= 1) Lt_top ^= T[row(g) ^ __shfl_sync(0xffffffffu, C, (32 - i) & 31)];
                }
            }
            lt_len = u + v + 1;
            for (int i = 0; i <= v; ++i) {  // Omega' = Lambda S' mod x^(d-1) (:1498-1499)
                const uint32_t li = __shfl_sync(0xffffffffu, C, i);
                Om ^= T[row(li) ^ shift_up(Sp, i, lane)];
            }
            if (lane >= d1) Om = 0;
        }
    }
    __syncwarp();
    LtS[k * kRow + lane] = state == ST_OK ? (uint8_t)Lt : (lane == 0 ? 1 : 0);  // a unit locator is never searched
    if (lane == 0) {
        LtS[k * kRow + 32] = state == ST_OK ? (uint8_t)Lt_top : 0;
        meta[k * 4 + 0] = state;
        meta[k * 4 + 1] = u;
        meta[k * 4 + 2] = v;
        meta[k * 4 + 3] = lt_len;
    }
    Ss[k * kRow + lane] = (uint8_t)Om;  // S is consumed: the slot now holds Omega'
    __syncwarp();
}

// Inversion-free Berlekamp-Massey (_lfsr.py:1647-1702) on the sequence in Sq (one element per byte, `ls` of them), one
// thread per sequence; `ls_max` is the loop bound shared by the warp.  C receives the connection polynomial (17
// coefficients); returns L (a value above 16 means C overflowed and is not usable).
__device__ __forceinline__ int bm_thread(const uint8_t* __restrict__ T, uint32_t (&Sq)[8], int ls, int ls_max, uint32_t (&C)[5]) {
    uint32_t Bs[5] = {0x100u, 0, 0, 0, 0};  // x^1 * B, B = 1
    uint32_t Sw[5] = {0, 0, 0, 0, 0};       // Sw_i = S[n - i]
#pragma unroll
    for (int j = 0; j < 5; ++j) C[j] = 0;
    C[0] = 1;
    uint32_t b = 1;
    int L = 0;
    for (int nn = 0; nn < ls_max; ++nn) {
        if (nn < ls) {
            shift_in_low(Sw, Sq[0] & 255u);
            shift_out_low(Sq);
            // C, x^m B and the new C have degree <= nn + 1: the words above that are zero and are skipped (uniform branches)
            uint32_t dsc = 0;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                if (4 * j <= nn + 1) {
#pragma unroll
                    for (int i = 4 * j; i < 4 * j + 4 && i < 17; ++i) dsc ^= T[row(byte_of(C, i)) ^ byte_of(Sw, i)];
                }
            }
            const uint32_t rb = row(b), rd = row(dsc);
            uint32_t Cn[5] = {0, 0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                if (4 * j <= nn + 1) {
#pragma unroll
                    for (int i = 4 * j; i < 4 * j + 4 && i < 17; ++i) {
                        const uint32_t cn = (uint32_t)T[rb ^ byte_of(C, i)] ^ (uint32_t)T[rd ^ byte_of(Bs, i)];
                        Cn[j] |= cn << (8 * (i & 3));
                    }
                }
            }
            const bool grow = dsc != 0 && 2 * L <= nn;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const uint32_t src = grow ? C[j] : Bs[j];  // B <- old C (m = 1) or m += 1
                if (dsc != 0) C[j] = Cn[j];
                Bs[j] = src;
            }
            shift_in_low(Bs, 0u);
            if (grow) {
                L = nn + 1 - L;
                b = dsc;
            }
        }
    }
    return L;
}

// Lambda = C[:min(L + 1, ls)] with high-degree zeros trimmed; returns its length
__device__ __forceinline__ int trim_locator(uint32_t (&C)[5], int L, int ls) {
    const int len = L + 1 < ls ? L + 1 : ls;
    int lam_len = 1;
#pragma unroll
    for (int i = 0; i < 17; ++i)
        if (i < len && byte_of(C, i) != 0) lam_len = i + 1;
#pragma unroll
    for (int i = 0; i < 17; ++i)
        if (i >= lam_len) C[i >> 2] &= ~(255u << (8 * (i & 3)));
    return lam_len;
}

// Omega'_l = sum_{i <= l} Lambda_i S_{l - i}, l < d1 (:1498-1499); Lambda_i = 0 beyond the degree and T[row(0) ^ s] = 0
__device__ __forceinline__ void omega_thread(const uint8_t* __restrict__ T, const uint32_t (&C)[5], const uint32_t (&S)[8], int d1,
                                             uint32_t (&Om)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) Om[j] = 0;
#pragma unroll
    for (int i = 0; i < 17; ++i) {
        const uint32_t ri = row(byte_of(C, i));
#pragma unroll
        for (int l = i; l < 32; ++l) Om[l >> 2] ^= (uint32_t)T[ri ^ byte_of(S, l - i)] << (8 * (l & 3));
    }
#pragma unroll
    for (int l = 0; l < 32; ++l)
        if (l >= d1) Om[l >> 2] &= ~(255u << (8 * (l & 3)));
}

// Thread-per-codeword key equation for a group WITHOUT erasures: lane k works on codeword k.  S: the d1 syndromes, one per
// byte.  Returns the state; Lt receives Lambda (17 coefficients), Om the coefficients of Omega' = Lambda S mod x^d1.
__device__ __forceinline__ int key_equation_thread(const uint8_t* __restrict__ T, const uint32_t (&S)[8], int d1,
                                                   uint32_t (&Lt)[9], uint32_t (&Om)[8], int& v_out) {
    uint32_t nzS = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) nzS |= S[j];
    v_out = 0;
    if (nzS == 0) return ST_CLEAN;  // :1373-1376
    uint32_t Sq[8], C[5];
#pragma unroll
    for (int j = 0; j < 8; ++j) Sq[j] = S[j];
    const int L = bm_thread(T, Sq, d1, d1, C);
    if (L > 16) return ST_SLOW;  // the locator may not fit 17 coefficients: the general path decides
    const int v = trim_locator(C, L, d1) - 1;
    v_out = v;
    if (2 * v > d1) return ST_FAIL;  // :1431-1433
    omega_thread(T, C, S, d1, Om);
#pragma unroll
    for (int j = 0; j < 9; ++j) Lt[j] = j < 5 ? C[j] : 0u;
    return ST_OK;
}

// The same with erasures: lane k holds the u erasure positions of codeword k in epos (ascending, at most 33 stored); u_max is
// the largest u of the group.  Lt receives Lambda_total = Gamma Lambda (33 coefficients), Om the coefficients of Omega'.
__device__ __forceinline__ int key_equation_thread_era(const fastrs::Tables& gf, const uint8_t* __restrict__ T, const uint32_t (&S)[8],
                                                       int d1, int la, int u, int u_max, const uint8_t* __restrict__ epos,
                                                       uint32_t (&Lt)[9], uint32_t (&Om)[8], int& v_out) {
    v_out = 0;
    if (u > d1) return ST_FAIL;  // :1357-1360
    uint32_t nzS = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) nzS |= S[j];
    if (nzS == 0 && u == 0) return ST_CLEAN;  // :1373-1376
    // Gamma(x) = prod (1 - Y_e x) (:1389-1393), 33 coefficients
    uint32_t G[9] = {1u, 0, 0, 0, 0, 0, 0, 0, 0};
    const int ub = u_max < 32 ? u_max : 32;
    for (int e = 0; e < ub; ++e) {
        if (e < u) {
            const uint32_t Yrow = row(gf.exp[((uint32_t)la * (uint32_t)epos[e]) % (uint32_t)gf.q1]);
            uint32_t Gn[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) Gn[j] = G[j];
#pragma unroll
            for (int i = 1; i <= 32; ++i) Gn[i >> 2] ^= (uint32_t)T[Yrow ^ byte_of(G, i - 1)] << (8 * (i & 3));
#pragma unroll
            for (int j = 0; j < 9; ++j) G[j] = Gn[j];
        }
    }
    // S'(x) = Gamma(x) S(x) mod x^(d-1) (:1408-1409)
    uint32_t Sp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if (i <= u_max) {  // uniform: Gamma_i = 0 beyond the largest erasure count
            const uint32_t ri = row(byte_of(G, i));
#pragma unroll
            for (int l = i; l < 32; ++l) Sp[l >> 2] ^= (uint32_t)T[ri ^ byte_of(S, l - i)] << (8 * (l & 3));
        }
    }
#pragma unroll
    for (int l = 0; l < 32; ++l)
        if (l >= d1) Sp[l >> 2] &= ~(255u << (8 * (l & 3)));
    // Berlekamp-Massey on S'[u:] (:1421-1428)
    uint32_t Sq[8], C[5];
#pragma unroll
    for (int j = 0; j < 8; ++j) Sq[j] = Sp[j];
    {  // drop the u low coefficients: whole words, then bytes
        const int wq = u >> 2, bq = 8 * (u & 3);
        if (wq & 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j) Sq[j] = j + 4 < 8 ? Sq[j + 4] : 0u;
        }
        if (wq & 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) Sq[j] = j + 2 < 8 ? Sq[j + 2] : 0u;
        }
        if (wq & 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) Sq[j] = j + 1 < 8 ? Sq[j + 1] : 0u;
        }
        if (wq & 8) {  // u = 32
#pragma unroll
            for (int j = 0; j < 8; ++j) Sq[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 7; ++j) Sq[j] = __funnelshift_r(Sq[j], Sq[j + 1], bq);
        Sq[7] >>= bq;
    }
    const int ls = d1 - u;
    const int L = bm_thread(T, Sq, ls, d1, C);
    if (L > 16) return ST_SLOW;
    int v = 0;
    if (u >= d1) {  // no syndromes left for the errors: Lambda = 1
#pragma unroll
        for (int j = 0; j < 5; ++j) C[j] = j == 0 ? 1u : 0u;
    } else {
        v = trim_locator(C, L, ls) - 1;
    }
    v_out = v;
    if (2 * v + u > d1) return ST_FAIL;  // :1431-1433
    // Lambda_total = Gamma Lambda (:1449), degree u + v <= 32
#pragma unroll
    for (int j = 0; j < 9; ++j) Lt[j] = 0;
#pragma unroll
    for (int i = 0; i < 17; ++i) {
        const uint32_t ri = row(byte_of(C, i));
#pragma unroll
        for (int l = i; l <= 32; ++l) Lt[l >> 2] ^= (uint32_t)T[ri ^ byte_of(G, l - i)] << (8 * (l & 3));
    }
    omega_thread(T, C, Sp, d1, Om);
    return ST_OK;
}

__global__ void __launch_bounds__(kThreadsPerCta, 1)
    k_rs_decode_bs(RsDev rs, const uint8_t* __restrict__ tab_g, const uint8_t* __restrict__ received,
                   const uint8_t* __restrict__ erasures, int ns, uint8_t* __restrict__ decoded,
                   long long* __restrict__ n_errors, size_t N) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    uint8_t* T = smem;
    uint8_t* s_exp = smem + kTab;
    uint8_t* s_inv = s_exp + 512;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* wbase = s_inv + 256 + (size_t)warp * kWarpBytes;
    uint32_t* R = reinterpret_cast<uint32_t*>(wbase + kOffPlanes);
    uint32_t* PM = reinterpret_cast<uint32_t*>(wbase + kOffMask);
    uint8_t* Ss = wbase + kOffS;
    uint8_t* LtS = wbase + kOffLt;
    uint32_t* LP = reinterpret_cast<uint32_t*>(wbase + kOffLP);
    int* meta = reinterpret_cast<int*>(wbase + kOffMeta);
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
    const fastrs::Tables gf{T, s_exp, s_inv, rs.q1};
    const int d1 = rs.d - 1;  // <= 32
    const int la = rs.log_alpha;
    const int q1 = rs.q1;
    const int cm1 = (int)((((long long)rs.c - 1) % q1 + q1) % q1);  // c - 1 as an exponent of the multiplicative group
    const size_t groups = (N + 31) / 32;
    const size_t warps_total = (size_t)gridDim.x * kWarps;
    uint32_t mk[64];

    for (size_t grp = (size_t)blockIdx.x + (size_t)gridDim.x * warp; grp < groups; grp += warps_total) {
        const size_t cw0 = grp * 32;
        const int ncw = (int)(N - cw0 < 32 ? N - cw0 : 32);
        const bool mine = lane < ncw;
        const uint8_t* rrow = received + (cw0 + lane) * ns;
        const uint8_t* erow = erasures ? erasures + (cw0 + lane) * ns : nullptr;
        __syncwarp();
        // ---- 1. output = received (the rows of the group are contiguous); planes of the received symbols
        {
            const uint8_t* src = received + cw0 * ns;
            uint8_t* dst = decoded + cw0 * ns;
            const size_t bytes = (size_t)ncw * ns;
            if ((((uintptr_t)src | (uintptr_t)dst | bytes) & 15) == 0) {  // full groups: 8160 bytes
                const uint4* s16 = reinterpret_cast<const uint4*>(src);
                uint4* d16 = reinterpret_cast<uint4*>(dst);
                const int n16 = (int)(bytes / 16);
                for (int i0 = lane; i0 < n16; i0 += 8 * 32) {
                    uint4 v[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        if (i0 + 32 * t < n16) v[t] = __ldg(s16 + i0 + 32 * t);
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        if (i0 + 32 * t < n16) d16[i0 + 32 * t] = v[t];
                }
            } else if ((((uintptr_t)src | (uintptr_t)dst) & 3) == 0) {
                const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
                uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
                for (size_t i = lane; i < bytes / 4; i += 32) d4[i] = s4[i];
                for (size_t i = (bytes & ~(size_t)3) + lane; i < bytes; i += 32) dst[i] = src[i];
            } else {
                for (size_t i = lane; i < bytes; i += 32) dst[i] = src[i];
            }
        }
        // Each lane walks its own row in ALIGNED 32-bit words (one load per four symbols; the row starts at any byte offset, so
        // neighbouring words are funnel-shifted together); byte m of the row is the symbol of degree ns - 1 - m.  A 32 x 32 bit
        // transpose across the warp turns one word per lane into the planes of four positions (lane 8 t + j: plane j of byte t).
        uint32_t any_era = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) PM[32 * r + lane] = 0;
        __syncwarp();
        {
            const int nwords = (ns + 3) >> 2;
            const int sh = 8 * (int)((uintptr_t)rrow & 3);
            const uint8_t* pa = rrow - ((uintptr_t)rrow & 3);
            const uint8_t* ea = erow ? erow - ((uintptr_t)erow & 3) : nullptr;
            const int esh = erow ? 8 * (int)((uintptr_t)erow & 3) : 0;
            const uint8_t* r_hi = received + N * (size_t)ns;
            const uint8_t* e_hi = erasures ? erasures + N * (size_t)ns : nullptr;
            for (int j0 = 0; j0 < nwords; j0 += 8) {
                uint32_t W[9], Ew[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) {  // the loads of a batch are in flight together
                    W[t] = (mine && j0 + t <= nwords) ? load_word_within(pa + 4 * (j0 + t), received, r_hi) : 0u;
                    Ew[t] = (erow && mine && j0 + t <= nwords) ? load_word_within(ea + 4 * (j0 + t), erasures, e_hi) : 0u;
                }
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int j = j0 + t;
                    if (j < nwords) {  // uniform
                        uint32_t w = __funnelshift_r(W[t], W[t + 1], sh);
                        const int m = 4 * j + (lane >> 3);  // the row byte this lane holds after the transpose
                        if (erow) {  // uniform: erased symbols are zeroed (_bch.py:1344-1355)
                            const uint32_t e = __funnelshift_r(Ew[t], Ew[t + 1], esh);
                            const uint32_t nz = (((e & 0x7f7f7f7fu) + 0x7f7f7f7fu) | e) & 0x80808080u;  // bit 7 of every non-zero byte
                            const uint32_t em8 = (nz >> 7) * 255u;
                            w &= ~em8;
                            const uint32_t em = transpose32(em8, lane);
                            const bool hit = m < ns && em != 0;
                            any_era |= __ballot_sync(0xffffffffu, hit);
                            if (m < ns && (lane & 7) == 0) PM[ns - 1 - m] = em;
                        }
                        const uint32_t planes = transpose32(w, lane);
                        if (m < ns) R[(ns - 1 - m) * 8 + (lane & 7)] = planes;
                    }
                }
            }
        }
        __syncwarp();
        // ---- 2. syndromes S_l = r(alpha^(c+l)), lane l, all 32 codewords at once (:1370)
        {
            build_masks(T, gf.pow_alpha(la, rs.c + lane), mk);
            uint32_t acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0;
            for (int i = ns - 1; i >= 0; --i) horner_step(mk, acc, R + i * 8);
            for (int k = 0; k < 32; ++k) {
                uint32_t byte = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) byte |= ((acc[j] >> k) & 1u) << j;
                Ss[k * kRow + lane] = lane < d1 ? (uint8_t)byte : 0;
            }
        }
        __syncwarp();
        // ---- 3. key equation, one thread per codeword
        uint8_t* epos = reinterpret_cast<uint8_t*>(LP);  // the LP region is free until step 4
        {
            uint32_t S[8], Lt[9], Om[8];
            const uint32_t* srow = reinterpret_cast<const uint32_t*>(Ss + lane * kRow);
#pragma unroll
            for (int j = 0; j < 8; ++j) S[j] = srow[j];
            int v = 0, u = 0, state;
            if (any_era == 0) {
                state = key_equation_thread(T, S, d1, Lt, Om, v);
            } else {
                uint8_t* ep = epos + lane * 33;  // erasure positions of codeword `lane`, ascending
#pragma unroll 1
                for (int r = 0; r < 8; ++r) {
                    for (uint32_t m = transpose32(PM[32 * r + lane], lane); m; m &= m - 1) {
                        if (u < 33) ep[u] = (uint8_t)(32 * r + __ffs(m) - 1);
                        ++u;
                    }
                }
                int u_max = u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int other = __shfl_xor_sync(0xffffffffu, u_max, o);
                    u_max = other > u_max ? other : u_max;
                }
                state = key_equation_thread_era(gf, T, S, d1, la, u, u_max, ep, Lt, Om, v);
            }
            if (!mine) state = ST_CLEAN;
            const uint32_t slow = __ballot_sync(0xffffffffu, state == ST_SLOW);
            if (state != ST_SLOW) {
                uint32_t* lrow = reinterpret_cast<uint32_t*>(LtS + lane * kRow);
                uint32_t* orow = reinterpret_cast<uint32_t*>(Ss + lane * kRow);
                const bool ok = state == ST_OK;
#pragma unroll
                for (int j = 0; j < 9; ++j) lrow[j] = ok ? Lt[j] : (j == 0 ? 1u : 0u);  // a unit locator is never searched
#pragma unroll
                for (int j = 0; j < 8; ++j) orow[j] = ok ? Om[j] : 0u;
                meta[lane * 4 + 0] = state;
                meta[lane * 4 + 1] = u;
                meta[lane * 4 + 2] = v;
                meta[lane * 4 + 3] = u + v + 1;
            }
            __syncwarp();
            for (uint32_t m = slow; m; m &= m - 1)  // long locators: the warp-per-codeword formulation
                key_equation_warp(gf, T, __ffs(m) - 1, lane, d1, la, ns, any_era != 0, PM, Ss, LtS, epos, meta);
        }
        __syncwarp();
        int max_len = (mine && meta[lane * 4 + 0] == ST_OK) ? meta[lane * 4 + 3] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int other = __shfl_xor_sync(0xffffffffu, max_len, o);
            max_len = other > max_len ? other : max_len;
        }
        // ---- 4. Chien search over i = 0 .. n-1 for the whole group (:1464-1475)
        if (max_len > 0) {
            for (int j = 0; j < max_len; ++j) {
                const uint32_t byte = mine ? LtS[lane * kRow + j] : (j == 0 ? 1u : 0u);
                uint32_t keep = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const uint32_t pb = __ballot_sync(0xffffffffu, (byte >> b) & 1u);
                    if (lane == b) keep = pb;
                }
                if (lane < 8) LP[j * 8 + lane] = keep;
            }
            __syncwarp();
            for (int r = 0; r < 8 && 32 * r < rs.n; ++r) {
                const int i = lane + 32 * r;
                int lx = (int)(((uint32_t)la * (uint32_t)i) % (uint32_t)q1);  // la < q1 <= 255, i < 256
                lx = lx ? q1 - lx : 0;
                build_masks(T, gf.exp[lx], mk);
                uint32_t acc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = 0;
                for (int j = max_len - 1; j >= 0; --j) horner_step(mk, acc, LP + j * 8);
                uint32_t nz = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) nz |= acc[j];
                PM[i] = (i < rs.n && i < ns) ? ~nz : 0u;  // roots at i >= ns are not recorded (:1469-1472)
            }
            for (int i = 32 * ((rs.n + 31) / 32) + lane; i < 256; i += 32) PM[i] = 0;
            __syncwarp();
        }
        // ---- 5. one thread per codeword: roots, Forney, correction
        {
            const int state = mine ? meta[lane * 4 + 0] : ST_CLEAN;
            const int u = meta[lane * 4 + 1], v = meta[lane * 4 + 2];
            long long result = state == ST_FAIL ? -1 : 0;
            uint8_t* pos = reinterpret_cast<uint8_t*>(LP) + lane * 33;  // LP is free again: root positions per codeword
            int cnt = 0;
            uint32_t rootw[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) rootw[r] = max_len > 0 ? transpose32(PM[32 * r + lane], lane) : 0u;  // bit m: position 32 r + m
            if (state == ST_OK) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    for (uint32_t m = rootw[r]; m; m &= m - 1) {
                        if (cnt < 33) pos[cnt] = (uint8_t)(32 * r + __ffs(m) - 1);
                        ++cnt;
                    }
                }
                if (cnt != v + u) {  // :1479-1481
                    result = -1;
                    cnt = 0;
                } else {
                    result = v;
                }
            }
            int max_cnt = cnt;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int other = __shfl_xor_sync(0xffffffffu, max_cnt, o);
                max_cnt = other > max_cnt ? other : max_cnt;
            }
            if (max_cnt > 0) {
                uint32_t Lt[9], Om[8];
                const uint32_t* lrow = reinterpret_cast<const uint32_t*>(LtS + lane * kRow);
                const uint32_t* orow = reinterpret_cast<const uint32_t*>(Ss + lane * kRow);
#pragma unroll
                for (int j = 0; j < 9; ++j) Lt[j] = lrow[j];
#pragma unroll
                for (int j = 0; j < 8; ++j) Om[j] = orow[j];
                // highest coefficient any lane of the group needs: the Horner chains below stop there
                int dO = 0, dL = 1;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (Om[j]) dO = 4 * j + 3 - (__clz(Om[j]) >> 3);
#pragma unroll
                for (int j = 0; j < 9; ++j)
                    if (Lt[j]) dL = 4 * j + 3 - (__clz(Lt[j]) >> 3);
                if (cnt == 0) dO = 0, dL = 1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int a = __shfl_xor_sync(0xffffffffu, dO, o), c = __shfl_xor_sync(0xffffffffu, dL, o);
                    dO = a > dO ? a : dO;
                    dL = c > dL ? c : dL;
                }
                for (int r = 0; r < max_cnt; ++r) {
                    if (r < cnt) {
                        // Forney (:1512-1556): E = Omega'(X^-1) / Lambda_total'(X^-1) * (X^-1)^(c-1); characteristic 2
                        const int my = pos[r];
                        int lx = (int)(((uint32_t)la * (uint32_t)my) % (uint32_t)q1);
                        lx = lx ? q1 - lx : 0;
                        const uint32_t xrow = row(gf.exp[lx]);
                        uint32_t num = 0, den = 0;
#pragma unroll
                        for (int j = 31; j >= 0; --j)
                            if (j <= dO) num = (uint32_t)T[xrow ^ num] ^ byte_of(Om, j);
#pragma unroll
                        for (int j = 32; j >= 1; --j)
                            if (j <= dL) den = (uint32_t)T[xrow ^ den] ^ ((j & 1) ? byte_of(Lt, j) : 0u);
                        uint32_t E = 0;
                        if (den != 0) {
                            E = gf.mul(num, gf.inv[den]);
                            const int ex = (int)(((uint32_t)lx * (uint32_t)cm1) % (uint32_t)q1);
                            E = gf.mul(E, gf.exp[ex]);
                        }
                        uint32_t orig = 0;  // the received symbol (0 where erased), read back from the planes
#pragma unroll
                        for (int j = 0; j < 8; ++j) orig |= ((R[my * 8 + j] >> lane) & 1u) << j;
                        uint32_t fixed;
                        if (rs.base2) {  // GF(2)'s subtract_modular rule for binary BCH (:1567-1576)
                            const int a = (int)orig, bb = (int)E;
                            fixed = (uint32_t)(a >= bb ? a - bb : 2 + a - bb) & 255u;
                        } else {
                            fixed = orig ^ E;
                        }
                        decoded[(cw0 + lane) * ns + (ns - 1 - my)] = (uint8_t)fixed;
                    }
                }
            }
            if (mine) n_errors[cw0 + lane] = result;
        }
        __syncwarp();
    }
}

}  // namespace bsrs
