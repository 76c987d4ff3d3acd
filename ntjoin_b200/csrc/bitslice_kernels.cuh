// bitslice_kernels.cuh -- bit-sliced candidate kernel (cand_variant 2).
//
// Same windowless prefilter as cand31_kernel (see sketch_kernels.cuh), different data layout: one thread runs 32
// independent streams, bit j of every register belongs to stream j.  The 31-bit rotation groups of fwd and rev live
// in 31+31 registers, one per bit; a rotation is a renaming of registers (resolved at compile time by unrolling 31
// steps), and injecting a base is one 3-input XOR (LOP3) per state bit:
//       F'[b] = F[b-1] ^ g_in,b(I1,I0) ^ g_out,b(O1,O0)
// where every g is an affine function of {x0, x1, x0&x1} of the 2-bit base code, i.e. one of 8 precomputed combos.
// The threshold test adds the top H bits of F and R with a bit-sliced ripple carry (carry-in 1 covers both the
// "+1" of the superset test and every carry from the dropped low bits), and compares the top HS sum bits with the
// runtime threshold by propagating the carry of  S + (2^HS-1-Q): no carry-out <=> S <= Q.
// ~125 LOP3 per 32 positions (3.9 ALU ops/base) against 9 per base for the word-parallel kernel.
//
// Input is the bit-plane transposed sequence produced by plane_kernel from pk:
//   tile T = 32 streams x L k-mer starts (32*L consecutive positions); stream j covers positions S_T + j*L + [0, L+k-1)
//   PL[(((T>>2) * R + t) * 4 + (T&3)) * 2 + pl]   bit j = code bit `pl` of base (S_T + j*L + t),  t in [0, R)
// Four tiles share a 32-byte row so that both the transposing stores and the per-step loads move whole sectors.
#pragma once
#include "sketch_kernels.cuh"

namespace mxe {

constexpr int BS_L = 512;            // k-mer starts per stream
constexpr int BS_TILE = 32 * BS_L;   // positions per thread
constexpr int BS_H = 16;             // adder width (top bits of the 31-bit lanes)
constexpr int BS_HS = 12;            // compared sum bits

__host__ __device__ constexpr int bs_iters(int k) { return (BS_L + k - 1 + 30) / 31; }
__host__ __device__ constexpr int bs_rows(int k) { return ((bs_iters(k) * 31 + 15) / 16) * 16; }

// ---------------------------------------------------------------- plane_kernel: pk -> transposed bit planes
// One CTA (4 warps) per group of four tiles, one warp per tile; lane j owns stream j.  A 32x32 bit transpose across
// lanes (5 shuffle rounds) turns "lane = stream, bit = (time, plane)" into "lane = (time, plane), bit = stream".
// Both sides go through shared memory so that global traffic is coalesced: the tile's pk slice (contiguous) is loaded
// with unit stride into a padded buffer (lane stride 32 words + 1 -> conflict-free strided reads), and the 16 rows x
// 4 tiles x 2 planes produced per step are gathered into one 512-byte contiguous store.
constexpr int PLANE_SLICE = BS_TILE / 16;                 // pk words per tile (without halo)

__global__ void __launch_bounds__(128) plane_kernel(const uint32_t* __restrict__ pk, uint64_t pk_words, uint64_t n_tiles, int rows,
                                                    uint32_t* __restrict__ PL)
{
    extern __shared__ uint32_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t g = blockIdx.x;
    const uint64_t T = g * 4 + warp;
    const int n_in = PLANE_SLICE + rows / 16;             // words of this tile's slice incl. the halo of the last stream
    uint32_t* in = sm + warp * (n_in + n_in / 32 + 1);
    uint32_t* outb = sm + 4 * (n_in + n_in / 32 + 1);     // [16 rows][4 tiles][2 planes]
    const uint64_t q0 = (T * BS_TILE) >> 4;
    for (int i = lane; i < n_in; i += 32) {
        const uint64_t wi = q0 + i;
        in[i + (i >> 5)] = (T < n_tiles && wi < pk_words) ? __ldg(pk + wi) : 0u;
    }
    __syncwarp();
    // pk bit layout inside a word: bit 8b + 2q + pl  <->  time u = 4q + b, plane pl
    const int b = lane >> 3, q = (lane >> 1) & 3, pl = lane & 1;
    const int u = 4 * q + b;
    uint32_t* dst = PL + (g * (uint64_t)rows) * 8;
    for (int m = 0; m < rows / 16; m++) {
        const int i = lane * (BS_L / 16) + m;
        uint32_t x = in[i + (i >> 5)];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const uint32_t md = d == 16 ? 0x0000FFFFu : d == 8 ? 0x00FF00FFu : d == 4 ? 0x0F0F0F0Fu : d == 2 ? 0x33333333u : 0x55555555u;
            const uint32_t y = __shfl_xor_sync(0xffffffffu, x, d);
            x = (lane & d) ? ((x & ~md) | ((y & ~md) >> d)) : ((x & md) | ((y & md) << d));
        }
        outb[(u * 4 + warp) * 2 + pl] = x;
        __syncthreads();
        dst[(uint64_t)m * 128 + threadIdx.x] = outb[threadIdx.x];
        __syncthreads();
    }
}

// ---------------------------------------------------------------- compile-time injection tables
struct BsTab {
    // per logical bit: combo index (0..7) and constant for the in- and out-terms of fwd and rev
    int f_in[31], f_out[31], f_neg[31];
    int r_in[31], r_out[31], r_neg[31];
};

__host__ __device__ constexpr uint32_t bs_shi(int code)
{
    return code == 0 ? (uint32_t)(SEED_A >> 33) : code == 1 ? (uint32_t)(SEED_C >> 33) : code == 2 ? (uint32_t)(SEED_T >> 33) : (uint32_t)(SEED_G >> 33);
}
// algebraic normal form over (x0, x1) of  code -> bit `bit` of shi[code ^ cx] : c | a0<<1 | a1<<2 | a01<<3
__host__ __device__ constexpr int bs_anf(int bit, int cx)
{
    const int f0 = (bs_shi(0 ^ cx) >> bit) & 1, f1 = (bs_shi(1 ^ cx) >> bit) & 1, f2 = (bs_shi(2 ^ cx) >> bit) & 1, f3 = (bs_shi(3 ^ cx) >> bit) & 1;
    return f0 | ((f0 ^ f1) << 1) | ((f0 ^ f2) << 2) | ((f0 ^ f1 ^ f2 ^ f3) << 3);
}
__host__ __device__ constexpr int bs_mod31(int x) { return ((x % 31) + 31) % 31; }

template <int KMOD>
__host__ __device__ constexpr BsTab bs_make_tab()
{
    BsTab t{};
    for (int b = 0; b < 31; b++) {
        // fwd' = rol(fwd) ^ rol^k(seed[out]) ^ seed[in]            (bit b of rol^k(s) = bit b-k of s)
        const int fi = bs_anf(b, 0), fo = bs_anf(bs_mod31(b - KMOD), 0);
        t.f_in[b] = fi >> 1; t.f_out[b] = fo >> 1; t.f_neg[b] = (fi ^ fo) & 1;
        // rev' = ror(rev ^ rol^k(seed[~in]) ^ seed[~out])          (XOR happens at the pre-rotation bit index)
        const int ri = bs_anf(bs_mod31(b - KMOD), 2), ro = bs_anf(b, 2);
        t.r_in[b] = ri >> 1; t.r_out[b] = ro >> 1; t.r_neg[b] = (ri ^ ro) & 1;
    }
    return t;
}

// plain expressions: the compiler folds constants / negations into a single LOP3 lookup table
__host__ __device__ __forceinline__ uint32_t bs_maj(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (a & c) | (b & c); }
__host__ __device__ __forceinline__ uint32_t bs_xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }

struct BsParams {
    uint64_t n;            // bases
    uint64_t n_tiles;
    int k, rows, iters;
    uint32_t kmask[BS_HS]; // bit i of (2^HS - 1 - Q) broadcast to a word, i = 0 (lsb) .. HS-1
    uint32_t f0, r0;       // 31-bit lane hashes of the all-A k-mer
};

template <int KMOD>
__global__ void __launch_bounds__(128) cand_bs_kernel(const uint32_t* __restrict__ PL, BsParams P, uint32_t* __restrict__ C)
{
    constexpr BsTab TB = bs_make_tab<KMOD>();
    const uint64_t T = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (T >= P.n_tiles) return;
    const uint2* pl = reinterpret_cast<const uint2*>(PL) + ((T >> 2) * (uint64_t)P.rows) * 4 + (T & 3);
    const uint64_t S_T = T * BS_TILE;
    const int k = P.k;

    uint32_t F[31], R[31];
#pragma unroll
    for (int r = 0; r < 31; r++) {      // phase 0: physical r holds logical bit r for both lanes
        F[r] = ((P.f0 >> r) & 1u) ? 0xFFFFFFFFu : 0u;
        R[r] = ((P.r0 >> r) & 1u) ? 0xFFFFFFFFu : 0u;
    }

    for (int it = 0; it < P.iters; it++) {
#pragma unroll
        for (int ph = 0; ph < 31; ph++) {
            const int t = it * 31 + ph;
            const uint2 in = __ldg(pl + (uint64_t)t * 4);
            uint2 ou = make_uint2(0u, 0u);                       // out base = A while warming up
            if (t >= k) ou = __ldg(pl + (uint64_t)(t - k) * 4);
            uint32_t I[8], O[8];
            I[0] = 0u; I[1] = in.x; I[2] = in.y; I[3] = in.x ^ in.y; I[4] = in.x & in.y; I[5] = I[1] ^ I[4]; I[6] = I[2] ^ I[4]; I[7] = I[3] ^ I[4];
            O[0] = 0u; O[1] = ou.x; O[2] = ou.y; O[3] = ou.x ^ ou.y; O[4] = ou.x & ou.y; O[5] = O[1] ^ O[4]; O[6] = O[2] ^ O[4]; O[7] = O[3] ^ O[4];
#pragma unroll
            for (int r = 0; r < 31; r++) {
                // fwd: physical r holds logical (r + ph) before, (r + ph + 1) after the rotation; inject at the new index
                const int bf = bs_mod31(r + ph + 1);
                F[r] = TB.f_neg[bf] ? ~bs_xor3(F[r], I[TB.f_in[bf]], O[TB.f_out[bf]]) : bs_xor3(F[r], I[TB.f_in[bf]], O[TB.f_out[bf]]);
                // rev: physical r holds logical (r - ph) before, (r - ph - 1) after; inject at the old index, then rotate
                const int br = bs_mod31(r - ph);
                R[r] = TB.r_neg[br] ? ~bs_xor3(R[r], I[TB.r_in[br]], O[TB.r_out[br]]) : bs_xor3(R[r], I[TB.r_in[br]], O[TB.r_out[br]]);
            }
            if (t < k - 1 || t - (k - 1) >= BS_L) continue;      // not yet a full k-mer / start owned by the next stream (uniform)
            // S = top BS_H bits of F + R + 1 ; compare its top BS_HS bits with Q through the carry of S + (2^HS-1-Q)
            uint32_t carry = 0xFFFFFFFFu, cmp = 0u;
#pragma unroll
            for (int i = 0; i < BS_H; i++) {
                const int b = 31 - BS_H + i;                     // logical bit, lsb of the adder first
                const uint32_t a = F[bs_mod31(b - (ph + 1))], c = R[bs_mod31(b + (ph + 1))];
                if (i >= BS_H - BS_HS) {
                    const uint32_t s = bs_xor3(a, c, carry);
                    cmp = bs_maj(s, P.kmask[i - (BS_H - BS_HS)], cmp);
                }
                carry = bs_maj(a, c, carry);
            }
            uint32_t cw = ~cmp;                                  // no carry-out  <=>  S <= Q  <=> candidate
            if (cw) {
                const uint64_t base = S_T + (uint64_t)(t - (k - 1));
                do {
                    const int j = __ffs(cw) - 1;
                    cw &= cw - 1;
                    const uint64_t p = base + (uint64_t)j * BS_L;
                    if (p < P.n) atomicOr(&C[p >> 5], 1u << (p & 31));       // fire-and-forget; V is applied by the mask pass
                } while (cw);
            }
        }
    }
}

// C &= V and the rank-directory counts of C (one warp per 1024-bit block)
__global__ void __launch_bounds__(256) cand_mask_count_kernel(uint32_t* __restrict__ C, const uint32_t* __restrict__ V, uint64_t n_words,
                                                               uint32_t* __restrict__ ccounts)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cw = 0;
    if (t < n_words) { cw = C[t] & V[t]; C[t] = cw; }
    uint32_t c = __popc(cw);
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && (t >> 5) * RANK_BLOCK_WORDS < n_words) ccounts[t >> 5] = c;
}

}  // namespace mxe
