// scan_kernels.cuh -- the sketch front end, second generation: pack2_kernel + scan_bs2_kernel.
//
// Same contract as pack_kernel + cand31_kernel (sketch_kernels.cuh): ASCII -> pk (2-bit codes), V (valid k-mer starts),
// C (candidate bitmap, superset of { hash0 >> 33 <= T }) and the rank-directory counts of V and C.  What changes is
// the instruction count per base: cand31_kernel rolls one stream per thread in 32-bit words (9.1 ALU-pipe + 4.9
// FMA-pipe instructions per base, 0.26 of the HBM roofline); here one thread runs 32 streams bit-sliced, so a
// rotation is a renaming and injecting a base is one LOP3 per state bit (62 LOP3 per 32 positions).
//
// Geometry.  The assembly is cut into TILES of 32 streams x L k-mer starts (L = 16 * Lw, Lw odd); stream j of tile T
// covers positions T*32L + j*L + [0, L + k - 1).  pack2_kernel (one CTA per tile, coalesced ASCII reads) writes, next
// to the position-ordered pk / V, the tile's codes once more in the order the bit-sliced kernel reads them:
//       pkT[(T * R + g) * 32 + j] = pk word g of stream j   (16 positions, same byte-interleaved bit layout as pk),
//       g in [0, R), R = ceil((L + k - 1) / 16): the last rows are the first words of stream j + 1 (the k-1 halo).
// One row = 128 contiguous bytes = what one thread of scan_bs2_kernel needs for 16 steps: eight 16-byte loads, every
// sector used completely, no shared-memory staging.  The thread transposes the 32 words in registers (32 x 32 bits:
// word = stream, bit = (step, plane)  ->  word = (step, plane), bit = stream), keeps the planes of the last k + 16
// steps in a shared-memory ring (the base that leaves the k-mer), and rolls.
//
// Rotation bookkeeping: after 16 steps the physical registers are rotated by 16 (31 moves per lane group), so every
// group of 16 steps starts at phase 0 and the unrolled body is the same for all groups (16 phases instead of 31: the
// transposition code exists once).
//
// V needs no second pass and C no mask pass: V is all ones except near invalid bases, the sequence end and record
// boundaries; pack2_kernel / boundary_kernel push those words on a DIRTY LIST and dirty_fix_kernel applies
// C &= V there only (falling back to a full pass if the list overflows -- decided on the device, no host round trip).
#pragma once
#include "sketch_kernels.cuh"
#include "bitslice_kernels.cuh"

namespace mxe {

struct ScanGeom {
    uint64_t n;          // bases
    uint64_t n_words;    // ceil(n / 32)
    uint64_t n_tiles;
    uint32_t Lw;         // pk words per stream (odd -> conflict-free transposed reads in pack2); L = 16 * Lw
    uint32_t R;          // pkT rows per tile
    uint32_t ext;        // 32-base units packed past the tile end (halo of V and of the last stream)
    int k;
};

struct DirtyList {
    uint32_t* items;     // word indices (V word != all ones)
    uint32_t* count;     // attempted pushes
    uint32_t capacity;
};

__device__ __forceinline__ void dirty_push(const DirtyList& D, uint64_t word)
{
    const uint32_t slot = atomicAdd(D.count, 1u);
    if (slot < D.capacity) D.items[slot] = (uint32_t)word;
}

// ---------------------------------------------------------------- 32 x 32 bit transpose in registers
// after the call: bit j of x[r] = bit r of the old x[j]
#ifdef __CUDA_ARCH__
#define MXE_PRMT(a, b, s) __byte_perm((a), (b), (s))
#else
static inline uint32_t mxe_prmt_host(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
#define MXE_PRMT(a, b, s) mxe_prmt_host((a), (b), (s))
#endif

__host__ __device__ __forceinline__ void transpose32(uint32_t x[32])
{
#pragma unroll
    for (int i = 0; i < 16; i++) {               // d = 16: swap high half of x[i] with low half of x[i + 16]
        const uint32_t a = x[i], b = x[i + 16];
        x[i] = MXE_PRMT(a, b, 0x5410);
        x[i + 16] = MXE_PRMT(a, b, 0x7632);
    }
#pragma unroll
    for (int i = 0; i < 32; i++) {               // d = 8: swap bytes 1,3 of x[i] with bytes 0,2 of x[i + 8]
        if (i & 8) continue;
        const uint32_t a = x[i], b = x[i + 8];
        x[i] = MXE_PRMT(a, b, 0x6240);
        x[i + 8] = MXE_PRMT(a, b, 0x7351);
    }
#pragma unroll
    for (int s = 0; s < 3; s++) {
        const int d = 4 >> s;
        const uint32_t m = d == 4 ? 0x0F0F0F0Fu : d == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
        for (int i = 0; i < 32; i++) {
            if (i & d) continue;
            const uint32_t t = ((x[i] >> d) ^ x[i + d]) & m;
            x[i + d] ^= t;
            x[i] ^= t << d;
        }
    }
}

// ---------------------------------------------------------------- pack2
// Validity of a byte B (ACGTacgt): B7 = 0, B6 = 1, B3 = 0, B0 != B4, B4 == (B2 & ~B1).  The last two conditions are
// evaluated at bit 4 of every byte from three left shifts (IMADs on the FMA pipe); the first three are accumulated
// over the eight words and tested once.
__host__ __device__ __forceinline__ uint32_t bad_screen(const uint32_t wv[8])
{
    uint32_t accA = 0u, accO = 0u, accN = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t w = wv[j];
        const uint32_t m16 = w * 16u, m4 = w * 4u, m8 = w * 8u;
        const uint32_t u = w ^ (m4 & ~m8);
        accA |= u | ~(w ^ m16);
        accO |= w;
        accN &= w;
    }
    return (accA & 0x10101010u) | (accO & 0x88888888u) | (~accN & 0x40404040u);
}

__host__ __device__ __forceinline__ uint32_t pk_from_ascii(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3)
{
    uint32_t r = x0 >> 1;                                        // bits 1:0 of every byte = (B2, B1) of x0
    r = ((x1 * 2u) & 0x0C0C0C0Cu) | (r & ~0x0C0C0C0Cu);
    r = ((x2 * 8u) & 0x30303030u) | (r & ~0x30303030u);
    r = ((x3 * 32u) & 0xC0C0C0C0u) | (r & ~0xC0C0C0C0u);
    return r;
}

// exact bad-base mask of 32 bases held in eight words (rare path of pack2_kernel)
__device__ __forceinline__ uint32_t bad_exact(const uint32_t wv[8])
{
    uint32_t bad = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t bb = bad_bytes((wv[j] & 0xDFDFDFDFu) ^ 0x41414141u);
#pragma unroll
        for (int q = 0; q < 4; q++)
            if ((bb >> (8 * q)) & 0xFFu) bad |= 1u << (4 * j + q);
    }
    return bad;
}

// the unit that contains the end of the sequence: bases at or beyond n are bad, code 0
__device__ __noinline__ uint32_t pack_unit_tail(const uint8_t* __restrict__ seq, uint64_t n, uint64_t gu, uint32_t pkw[2])
{
    const uint64_t p0 = gu << 5;
    uint32_t bad = 0;
    pkw[0] = pkw[1] = 0;
    for (int i = 0; i < 32; i++) {
        const uint64_t p = p0 + i;
        uint32_t code = 0;
        bool ok = false;
        if (p < n) {
            const uint32_t x = ((uint32_t)seq[p] & 0xDFu) ^ 0x41u;
            ok = (x == 0x00u) | (x == 0x02u) | (x == 0x06u) | (x == 0x15u);
            code = (x >> 1) & 3u;
        }
        if (!ok) { bad |= 1u << i; code = 0; }
        const uint32_t ii = i & 15;
        pkw[i >> 4] |= code << (((ii & 3u) << 3) | ((ii >> 2) << 1));
    }
    return bad;
}

constexpr int PACK2_THREADS = 256;      // most; short tiles are packed by 128-thread CTAs (blockDim.x is what the kernel uses)
constexpr int PACK2_INFLIGHT = 4;      // 32-base units per thread whose loads are issued before any is consumed
__host__ __device__ inline size_t pack2_smem_bytes(uint32_t Lw, uint32_t ext) { return (size_t)3 * (16u * Lw + ext) * sizeof(uint32_t); }

// One CTA per tile (tiles [T0, T1)).  vcounts must be zeroed before the first launch (blocks straddle tiles).
// PL: the tile's codes as bit planes, row g = steps 16 g .. 16 g + 15: word 2 u + pl = plane pl of step u, bit j = stream j.
__global__ void __launch_bounds__(PACK2_THREADS) pack2_kernel(const uint8_t* __restrict__ seq, ScanGeom G, uint32_t* __restrict__ pk,
                                                               uint32_t* __restrict__ PL, uint32_t* __restrict__ V,
                                                               uint32_t* __restrict__ vcounts, DirtyList D, uint64_t T0)
{
    extern __shared__ uint32_t p2s[];
    const uint32_t L = 16u * G.Lw;                 // 32-base units per tile ( = k-mer starts per stream)
    const uint32_t U = L + G.ext;
    uint32_t* pkS = p2s;                           // 2 * U words
    uint32_t* badS = p2s + 2 * U;                  // U words
    const uint64_t T = T0 + blockIdx.x;
    const uint64_t u_base = T * (uint64_t)L;       // first global unit of the tile
    int seen_bad = 0;
    for (uint32_t ub = 0; ub < U; ub += blockDim.x * PACK2_INFLIGHT) {
        uint4 a[PACK2_INFLIGHT], b[PACK2_INFLIGHT];
#pragma unroll
        for (int i = 0; i < PACK2_INFLIGHT; i++) {           // all loads of this thread first: the kernel lives on bytes in flight
            const uint32_t u = ub + i * blockDim.x + threadIdx.x;
            const uint64_t gu = u_base + u;
            if (u < U && (gu << 5) + 32 <= G.n) {
                const uint4* src = reinterpret_cast<const uint4*>(seq + (gu << 5));
                a[i] = __ldg(src);
                b[i] = __ldg(src + 1);
            }
        }
#pragma unroll
        for (int i = 0; i < PACK2_INFLIGHT; i++) {
            const uint32_t u = ub + i * blockDim.x + threadIdx.x;
            const uint64_t gu = u_base + u;
            if (u >= U) continue;
            uint32_t pkw[2] = {0u, 0u};
            uint32_t bad = 0xFFFFFFFFu;                      // units past the sequence: all bad
            if ((gu << 5) + 32 <= G.n) {
                const uint32_t wv[8] = {a[i].x, a[i].y, a[i].z, a[i].w, b[i].x, b[i].y, b[i].z, b[i].w};
                pkw[0] = pk_from_ascii(wv[0], wv[1], wv[2], wv[3]);
                pkw[1] = pk_from_ascii(wv[4], wv[5], wv[6], wv[7]);
                bad = bad_screen(wv) ? bad_exact(wv) : 0u;
            } else if (gu < G.n_words) {
                bad = pack_unit_tail(seq, G.n, gu, pkw);
            }
            if (u < L && gu < G.n_words) reinterpret_cast<uint2*>(pk)[gu] = make_uint2(pkw[0], pkw[1]);
            pkS[2 * u] = pkw[0];
            pkS[2 * u + 1] = pkw[1];
            badS[u] = bad;
            seen_bad |= bad != 0u;
        }
    }
    const int lane = threadIdx.x & 31;
    if (!__syncthreads_or(seen_bad)) {
        // the usual tile: every base valid and the sequence continues past the halo -> V is all ones
        for (uint32_t u0 = 0; u0 < L; u0 += blockDim.x) {
            const uint32_t u = u0 + threadIdx.x;
            if (u < L) V[u_base + u] = 0xFFFFFFFFu;
            const uint32_t w0 = u0 + (threadIdx.x & ~31u);                     // first unit of this warp
            if (lane == 0 && w0 < L) {
                const uint64_t g0 = u_base + w0;
                const uint32_t n_act = L - w0 < 32u ? L - w0 : 32u;
                const uint32_t na = 32u - (uint32_t)(g0 & 31u) < n_act ? 32u - (uint32_t)(g0 & 31u) : n_act;
                atomicAdd(&vcounts[g0 >> 5], 32u * na);
                if (n_act > na) atomicAdd(&vcounts[(g0 >> 5) + 1], 32u * (n_act - na));
            }
        }
    } else {
        // V + rank counts (a warp's 32 consecutive units touch at most two 1024-bit rank blocks)
        const int reach = (31 + G.k - 1) / 32;
        for (uint32_t u0 = 0; u0 < L; u0 += blockDim.x) {
            const uint32_t u = u0 + threadIdx.x;
            const uint64_t gu = u_base + u;
            const bool act = u < L && gu < G.n_words;
            uint32_t v = 0;
            if (act) {
                uint32_t any = 0;
                for (int j = 0; j <= reach; j++) any |= badS[u + j];
                v = any ? vmask_from([&](int j) { return badS[u + j]; }, reach, G.k) : 0xFFFFFFFFu;
                V[gu] = v;
                if (v != 0xFFFFFFFFu) dirty_push(D, gu);
            }
            const uint32_t c = __popc(v);
            const uint64_t blk = gu >> 5;
            const uint64_t blk0 = __shfl_sync(0xffffffffu, blk, 0);
            const uint32_t sa = __reduce_add_sync(0xffffffffu, blk == blk0 ? c : 0u);
            const uint32_t sb = __reduce_add_sync(0xffffffffu, blk == blk0 ? 0u : c);
            if (lane == 0) {
                if (sa) atomicAdd(&vcounts[blk0], sa);
                if (sb) atomicAdd(&vcounts[blk0 + 1], sb);
            }
        }
    }
    // Bit planes for scan_bs2_kernel: thread g takes word g of every stream (stream j starts at pk word j * Lw; Lw odd
    // and consecutive g -> conflict-free), transposes 32 x 32 bits in registers and writes one 128-byte row.  pack2 is
    // bound by memory latency with the ALU pipe half idle, scan_bs2 by the ALU pipe: the transposition costs less here.
    for (uint32_t g = threadIdx.x; g < G.R; g += blockDim.x) {
        uint32_t x[32];
#pragma unroll
        for (int j = 0; j < 32; j++) x[j] = pkS[(uint32_t)j * G.Lw + g];
        transpose32(x);
        // pk bit 8b + 2q + pl  <->  step u = 4q + b of the group, plane pl
        uint4* dst = reinterpret_cast<uint4*>(PL + (T * (uint64_t)G.R + g) * 32u);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int u0 = 2 * q, u1 = 2 * q + 1;
            dst[q] = make_uint4(x[8 * (u0 & 3) + 2 * (u0 >> 2)], x[8 * (u0 & 3) + 2 * (u0 >> 2) + 1],
                                x[8 * (u1 & 3) + 2 * (u1 >> 2)], x[8 * (u1 & 3) + 2 * (u1 >> 2) + 1]);
        }
    }
}

// record boundaries: a k-mer may not straddle two records (as boundary_kernel) + dirty list
__global__ void boundary2_kernel(const uint64_t* __restrict__ offsets, uint32_t n_contigs, uint64_t n, int k, uint32_t* __restrict__ V,
                                 uint32_t* __restrict__ vcounts, DirtyList D)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (c >= n_contigs) return;
    const uint64_t q = offsets[c];
    if (q == 0 || q >= n) return;
    const uint64_t lo = q >= (uint64_t)(k - 1) ? q - (k - 1) : 0;
    for (uint64_t wi = lo >> 5; wi <= ((q - 1) >> 5) && lo < q; wi++) {
        const uint64_t a = wi << 5;
        const uint32_t b0 = lo > a ? (uint32_t)(lo - a) : 0u, b1 = q - a < 32 ? (uint32_t)(q - a) : 32u;   // clear bits [b0, b1)
        const uint32_t m = (b1 >= 32 ? 0xFFFFFFFFu : ((1u << b1) - 1u)) & ~((1u << b0) - 1u);
        const uint32_t old = atomicAnd(&V[wi], ~m);
        const uint32_t cleared = old & m;
        if (cleared) atomicSub(&vcounts[wi >> 5], (uint32_t)__popc(cleared));
        dirty_push(D, wi);
    }
}

// C &= V where V is not all ones (after the candidate kernel; the rank counts of C are taken afterwards)
__global__ void __launch_bounds__(256) dirty_fix_kernel(DirtyList D, const uint32_t* __restrict__ V, uint32_t* __restrict__ C, uint64_t n_words)
{
    const uint32_t cnt = *D.count;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cnt <= D.capacity) {
        for (uint64_t i = i0; i < cnt; i += stride) {
            const uint32_t t = D.items[i];
            atomicAnd(&C[t], V[t]);                          // the list may hold a word twice
        }
    } else {
        for (uint64_t t = i0; t < n_words; t += stride) C[t] &= V[t];    // list overflowed: every word once
    }
}

// opaque logic primitives (see bs2_tile): plain expressions on the host (tools/scan_emul.cu), single PTX instructions on the device
__host__ __device__ __forceinline__ uint32_t bs2_xor(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    uint32_t d; asm("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
#else
    return a ^ b;
#endif
}
__host__ __device__ __forceinline__ uint32_t bs2_and(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    uint32_t d; asm("and.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
#else
    return a & b;
#endif
}
// state ^= I[CI] ^ O[CO] ^ NEG   (combo 0 = absent)
__host__ __device__ __forceinline__ int bs2_ffs(uint32_t v)       // index of the lowest set bit (v != 0)
{
#ifdef __CUDA_ARCH__
    return __ffs((int)v) - 1;
#else
    int i = 0;
    while (!((v >> i) & 1u)) i++;
    return i;
#endif
}
// (ci, co, neg are compile-time constants after unrolling: the branches fold to one instruction)
__host__ __device__ __forceinline__ uint32_t bs2_inject(uint32_t st, const uint32_t (&I)[8], const uint32_t (&O)[8], const int CI, const int CO, const int NEG)
{
#ifdef __CUDA_ARCH__
    uint32_t d;
    if (CI && CO) {
        if (NEG) asm("lop3.b32 %0, %1, %2, %3, 0x69;" : "=r"(d) : "r"(st), "r"(I[CI]), "r"(O[CO]));
        else asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(st), "r"(I[CI]), "r"(O[CO]));
        return d;
    }
    if (CI || CO) {
        const uint32_t v = CI ? I[CI] : O[CO];
        if (NEG) asm("lop3.b32 %0, %1, %2, 0, 0xc3;" : "=r"(d) : "r"(st), "r"(v));      // ~(a ^ b)
        else asm("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(st), "r"(v));
        return d;
    }
    return NEG ? ~st : st;
#else
    const uint32_t v = st ^ I[CI] ^ O[CO];
    return NEG ? ~v : v;
#endif
}

// ---------------------------------------------------------------- bit-sliced candidate scan of one tile
struct Bs2Params {
    uint32_t f0, r0;          // 31-bit lane hashes of the all-A k-mer
    uint32_t kmask[16];       // bit i of (2^HS - 1 - Q) broadcast to a word
};

constexpr int BS2_THREADS = 128;
__host__ __device__ constexpr int bs2_ring_slots(int k) { return k + 16 <= 48 ? 48 : 64; }   // steps of plane history (>= k + 16, multiple of 16)
constexpr int BS2_MAX_K = 48;

__host__ __device__ __forceinline__ void bs2_load_row(const uint32_t* __restrict__ row, uint32_t x[32])
{
#ifdef __CUDA_ARCH__
    // four 32-byte loads (sm_100 256-bit LDG), streaming: evict_first in L2 so that they do not push out C's open sectors
#pragma unroll
    for (int q = 0; q < 4; q++)
        asm("ld.global.nc.L2::evict_first.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
            : "=r"(x[8 * q]), "=r"(x[8 * q + 1]), "=r"(x[8 * q + 2]), "=r"(x[8 * q + 3]), "=r"(x[8 * q + 4]), "=r"(x[8 * q + 5]),
              "=r"(x[8 * q + 6]), "=r"(x[8 * q + 7])
            : "l"(row + 8 * q));
#else
    for (int i = 0; i < 32; i++) x[i] = row[i];
#endif
}

// rows: the tile's R rows of 32 plane words (written by pack2_kernel).
// ring: planes of the last RING steps, get(slot, lo, hi) / put(slot, lo, hi);  q: the 16 candidate words of the group
// just finished, get(ph) / put(ph, w) (bit j = stream j);  emit(j, s): k-mer start s of stream j is a candidate.
template <int KMOD, int H, int HS, int RING, typename Ring, typename Queue, typename Emit>
__host__ __device__ __forceinline__ void bs2_tile(const uint32_t* __restrict__ rows, int R, int L, int k, const Bs2Params& P, Ring& ring, Queue& q, Emit& emit)
{
    constexpr BsTab TB = bs_make_tab<KMOD>();
    uint32_t F[31], Rv[31];
#pragma unroll
    for (int r = 0; r < 31; r++) {
        F[r] = ((P.f0 >> r) & 1u) ? 0xFFFFFFFFu : 0u;
        Rv[r] = ((P.r0 >> r) & 1u) ? 0xFFFFFFFFu : 0u;
    }
    for (int s = 0; s < RING; s++) ring.put(s, 0u, 0u);          // out base = A while warming up
    int slot_in = 0;                                             // (16 g) mod RING
    int slot_out = RING - k;                                     // (16 g - k) mod RING, kept in [0, RING)
    while (slot_out < 0) slot_out += RING;
    for (int g = 0; g < R; g++) {
        uint32_t x[32];                                           // x[2 u + pl] = plane pl of step u (bit j = stream j)
        bs2_load_row(rows + (size_t)g * 32, x);
        const int t0 = g * 16;
#pragma unroll
        for (int u = 0; u < 16; u++) ring.put(slot_in + u, x[2 * u], x[2 * u + 1]);
        uint32_t cwv[16];
#pragma unroll
        for (int ph = 0; ph < 16; ph++) {
            const uint32_t ix = x[2 * ph], iy = x[2 * ph + 1];
            uint32_t ox, oy;
            int so = slot_out + ph;
            if (so >= RING) so -= RING;
            ring.get(so, ox, oy);                                // zero while t < k (ring starts zeroed, k + 16 <= RING)
            // the seven non-trivial combos of each base are computed ONCE and every state bit is one 3-input XOR of
            // (state, in-combo, out-combo); opaque to the compiler, which otherwise re-derives the combos inside two
            // LOP3s per state bit (F ^ g(ix, iy), then ^ h(ox, oy)): 96 instead of 72 logic instructions per step
            uint32_t I[8], O[8];
            I[0] = 0u; I[1] = ix; I[2] = iy; I[3] = bs2_xor(ix, iy); I[4] = bs2_and(ix, iy);
            I[5] = bs2_xor(I[1], I[4]); I[6] = bs2_xor(I[2], I[4]); I[7] = bs2_xor(I[3], I[4]);
            O[0] = 0u; O[1] = ox; O[2] = oy; O[3] = bs2_xor(ox, oy); O[4] = bs2_and(ox, oy);
            O[5] = bs2_xor(O[1], O[4]); O[6] = bs2_xor(O[2], O[4]); O[7] = bs2_xor(O[3], O[4]);
#pragma unroll
            for (int r = 0; r < 31; r++) {
                const int bf = bs_mod31(r + ph + 1);            // fwd: physical r holds logical (r + ph) before, (r + ph + 1) after
                F[r] = bs2_inject(F[r], I, O, TB.f_in[bf], TB.f_out[bf], TB.f_neg[bf]);
                const int br = bs_mod31(r - ph);                // rev: physical r holds logical (r - ph) before, (r - ph - 1) after
                Rv[r] = bs2_inject(Rv[r], I, O, TB.r_in[br], TB.r_out[br], TB.r_neg[br]);
            }
            // S = top H bits of F + R + 1 ; its top HS bits <= Q  <=>  no carry out of S_top + (2^HS - 1 - Q).
            // No branch inside the 16 steps, so the serial carry chain of this step overlaps the independent state
            // updates of the next one.
            uint32_t carry = 0xFFFFFFFFu, cmp = 0u;
#pragma unroll
            for (int i = 0; i < H; i++) {
                const int b = 31 - H + i;
                const uint32_t a = F[bs_mod31(b - (ph + 1))], c = Rv[bs_mod31(b + (ph + 1))];
                if (i >= H - HS) {
                    const uint32_t sm = bs_xor3(a, c, carry);
                    cmp = bs_maj(sm, P.kmask[i - (H - HS)], cmp);
                }
                if (i + 1 < H) carry = bs_maj(a, c, carry);
            }
            cwv[ph] = ~cmp;
        }
        // Emission, once per group: the candidate words (~1 % of the bits set) go through a per-thread queue and ONE
        // loop pops one candidate per trip, so a warp makes max-over-lanes(candidates of the group) ~ 10 trips per 16
        // steps; a loop per step would make 16 x (max-over-lanes per word ~ 2.5) trips.
        const int s_lo = t0 - (k - 1);
        uint32_t nz = 0u;
#pragma unroll
        for (int ph = 0; ph < 16; ph++) {
            const uint32_t w = (s_lo + ph >= 0 && s_lo + ph < L) ? cwv[ph] : 0u;      // starts that exist (uniform)
            q.put(ph, w);
            if (w) nz |= 1u << ph;
        }
        uint32_t cur = 0u;
        int s_cur = 0;
        for (;;) {
            if (!cur) {
                if (!nz) break;
                const int ph = bs2_ffs(nz);
                nz &= nz - 1u;
                cur = q.get(ph);
                s_cur = s_lo + ph;
            }
            const int j = bs2_ffs(cur);
            cur &= cur - 1u;
            emit(j, s_cur);
        }
        // rotate the physical registers back to phase 0
        uint32_t Fn[31], Rn[31];
#pragma unroll
        for (int r = 0; r < 31; r++) { Fn[(r + 16) % 31] = F[r]; Rn[(r + 15) % 31] = Rv[r]; }
#pragma unroll
        for (int r = 0; r < 31; r++) { F[r] = Fn[r]; Rv[r] = Rn[r]; }
        slot_in += 16; if (slot_in >= RING) slot_in -= RING;
        slot_out += 16; if (slot_out >= RING) slot_out -= RING;
    }
}

struct Bs2RingSmem {
    uint2* base;            // [RING][BS2_THREADS], this thread's column
    __device__ __forceinline__ void put(int slot, uint32_t lo, uint32_t hi) { base[slot * BS2_THREADS] = make_uint2(lo, hi); }
    __device__ __forceinline__ void get(int slot, uint32_t& lo, uint32_t& hi) const { const uint2 v = base[slot * BS2_THREADS]; lo = v.x; hi = v.y; }
};
struct Bs2QueueSmem {
    uint32_t* base;         // [16][BS2_THREADS]
    __device__ __forceinline__ void put(int ph, uint32_t w) { base[ph * BS2_THREADS] = w; }
    __device__ __forceinline__ uint32_t get(int ph) const { return base[ph * BS2_THREADS]; }
};
struct Bs2EmitGlobal {
    uint32_t* c;            // C at the tile's first position (a multiple of 32)
    uint32_t L;
    __device__ __forceinline__ void operator()(int j, int s) const
    {
        const uint32_t rel = (uint32_t)j * L + (uint32_t)s;
        atomicOr(c + (rel >> 5), 1u << (rel & 31u));          // every position belongs to one (stream, step): set once
    }
};

__host__ __device__ inline size_t bs2_smem_bytes(int k) { return (size_t)BS2_THREADS * (bs2_ring_slots(k) * sizeof(uint2) + 16 * sizeof(uint32_t)); }

// C must hold n_tiles * 16 * Lw words, zeroed (the last tile may set bits at or beyond n; they are cleared by
// dirty_fix_kernel like every other position without a valid k-mer, words past n_words are never read).
template <int KMOD, int H, int HS, int RING>
__global__ void __launch_bounds__(BS2_THREADS) scan_bs2_kernel(const uint32_t* __restrict__ PL, ScanGeom G, Bs2Params P, uint32_t* __restrict__ C)
{
    extern __shared__ uint2 bs2_smem[];
    const uint64_t T = (uint64_t)blockIdx.x * BS2_THREADS + threadIdx.x;
    if (T >= G.n_tiles) return;
    const uint32_t L = 16u * G.Lw;
    Bs2RingSmem ring{bs2_smem + threadIdx.x};
    Bs2QueueSmem q{reinterpret_cast<uint32_t*>(bs2_smem + RING * BS2_THREADS) + threadIdx.x};
    Bs2EmitGlobal emit{C + T * (uint64_t)L, L};
    bs2_tile<KMOD, H, HS, RING>(PL + T * (uint64_t)G.R * 32u, (int)G.R, (int)L, G.k, P, ring, q, emit);
}

}  // namespace mxe
