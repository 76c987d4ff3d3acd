// sketch_kernels.cuh -- step 1 on the device: ordered ntHash minimizer sketch.
//
// Replaces btllib indexlr as invoked at ntJoin:204-205 (algorithm: SURVEY.md Appendix A).
//
// Pipeline (all position-indexed bitmaps are 1 bit per base of the concatenated assembly):
//   pack        ASCII -> 2-bit codes (pk) + bad-byte bitmap B                 [HBM-bound stream]
//   vmask       V = positions where a valid k-mer starts (no bad byte, no contig boundary inside)
//   rank(V)     valid-k-mer ordinals (windows are over ordinals)
//   cand        C = V & { top 31 bits of hash0 <= T }      windowless prefilter [the hot kernel]
//   extract(C), cand_eval (exact 64-bit hash0, ordinal, contig)
//   select      sparse sliding-window arg-min over candidates -> minimizer bitmap M ; windows whose
//               minimum cannot be decided from candidates become "gaps"
//   gap         dense exact evaluation of gap windows -> M
//   extract(M), final_eval -> (out_hash, min_hash, pos, contig, forward) sorted by position
//
// Why the prefilter is exact: let t(p) = hash0(p) >> 33.  Every window whose candidate arg-min has
// t <= T has its true arg-min among the candidates (all k-mers with t <= T are candidates); every
// other window is re-evaluated densely.  Ties on hash0 resolve to the rightmost k-mer (btllib `<=`).
#pragma once
#include "common.cuh"
#include <type_traits>

namespace mxe {

struct SketchTables {
    uint64_t seed[4];      // by device code (A C T G)
    uint64_t seed_rolk[4]; // srol^k(seed)
    uint2 t16[16];         // 31-bit lane roll table: idx = out<<2|in -> {fwd term, rev term}
    uint32_t shi[4];       // seed >> 33 by code
    uint32_t f0, r0;       // 31-bit lane hashes of the all-A k-mer (warm-up seed of cand31)
};

struct SketchParams {
    uint64_t n;            // bases
    uint64_t n_words;      // ceil(n/32)
    int k, w;
    int canon_min;         // 0 sum, 1 min
    uint32_t T;            // candidate threshold on hash0>>33
    int chunk;             // positions per thread in cand kernels
    uint64_t pk_words;     // allocated words of pk (zero padded past the sequence)
    uint32_t mul1, mul2;   // the constants 1 and 2 as run-time values (IMAD operands in cand31_kernel)
};

// ---------------------------------------------------------------- bit helpers
__host__ __device__ __forceinline__ uint64_t srol1(uint64_t x)
{
    uint64_t m = ((x & 0x8000000000000000ULL) >> 30) | ((x & 0x100000000ULL) >> 32);
    return ((x << 1) & 0xFFFFFFFDFFFFFFFFULL) | m;
}
__host__ __device__ __forceinline__ uint64_t sror1(uint64_t x)
{
    uint64_t m = ((x & 0x200000000ULL) << 30) | ((x & 1ULL) << 32);
    return ((x >> 1) & 0xFFFFFFFEFFFFFFFFULL) | m;
}
__host__ __device__ __forceinline__ uint32_t rol31(uint32_t v) { return ((v << 1) | (v >> 30)) & 0x7FFFFFFFu; }
__host__ __device__ __forceinline__ uint32_t ror31(uint32_t v) { return (v >> 1) | ((v & 1u) << 30); }

__host__ __device__ __forceinline__ uint64_t mix_out_hash(uint64_t h0, int k)
{
    uint64_t t = h0 * (1ULL ^ ((uint64_t)k * MULTISEED));
    return t ^ (t >> MULTISHIFT);
}

// pk layout: word q holds positions 16q..16q+15; position i = 4*j + b sits at bit 8*b + 2*j
// (byte-interleaved so that packing four ASCII words costs two ops per word).
__device__ __forceinline__ uint32_t pk_code(const uint32_t* __restrict__ pk, uint64_t p)
{
    uint32_t wv = pk[p >> 4];
    uint32_t i = (uint32_t)p & 15u;
    uint32_t sh = ((i & 3u) << 3) | ((i >> 2) << 1);
    return (wv >> sh) & 3u;
}

__device__ __forceinline__ uint64_t sel4(uint32_t c, const uint64_t* s)
{
    uint64_t lo = (c & 1u) ? s[1] : s[0];
    uint64_t hi = (c & 1u) ? s[3] : s[2];
    return (c & 2u) ? hi : lo;
}

// exact base hashes of the k-mer at p (all bases assumed valid)
__device__ __forceinline__ void kmer_hash64(const uint32_t* __restrict__ pk, uint64_t p, int k, const uint64_t* seed,
                                            uint64_t& fwd, uint64_t& rev)
{
    uint64_t f = 0, r = 0;
    for (int i = 0; i < k; i++) {
        f = srol1(f) ^ sel4(pk_code(pk, p + i), seed);
        r = srol1(r) ^ sel4(pk_code(pk, p + k - 1 - i) ^ 2u, seed);
    }
    fwd = f;
    rev = r;
}

__device__ __forceinline__ uint64_t canon(uint64_t f, uint64_t r, int canon_min)
{
    return canon_min ? (r < f ? r : f) : f + r;
}

// ---------------------------------------------------------------- table-driven exact k-mer hash
// srol^n / sror^n for 0 < n < 31 on the 33|31 split word
__host__ __device__ __forceinline__ uint64_t sroln(uint64_t x, int n)
{
    uint64_t g33 = x & 0x1FFFFFFFFULL, g31 = x >> 33;
    g33 = ((g33 << n) | (g33 >> (33 - n))) & 0x1FFFFFFFFULL;
    g31 = ((g31 << n) | (g31 >> (31 - n))) & 0x7FFFFFFFULL;
    return g33 | (g31 << 33);
}
__host__ __device__ __forceinline__ uint64_t srorn(uint64_t x, int n)
{
    uint64_t g33 = x & 0x1FFFFFFFFULL, g31 = x >> 33;
    g33 = ((g33 >> n) | (g33 << (33 - n))) & 0x1FFFFFFFFULL;
    g31 = ((g31 >> n) | (g31 << (31 - n))) & 0x7FFFFFFFULL;
    return g33 | (g31 << 33);
}

// pk word (byte-interleaved) -> natural order (base i at bits 2i): 4x4 transpose of 2-bit fields
__device__ __forceinline__ uint32_t pk_to_natural(uint32_t x)
{
    uint32_t t = ((x >> 12) ^ x) & 0x0000F0F0u;
    x ^= t ^ (t << 12);
    t = ((x >> 6) ^ x) & 0x00CC00CCu;
    x ^= t ^ (t << 6);
    return x;
}

// Shared-memory tables for 4 bases per lookup (index = natural byte, base j at bits 2j):
//   f4[v]  = XOR_j srol^(3-j)(seed[c_j])                    fwd:  f = srol^4(f) ^ f4[v]
//   r4[v]  = srol^(k-4)( XOR_j srol^j(seed[c_j ^ 2]) )      rev:  r = sror^4(r) ^ r4[v]
//   s1[c]  = seed[c] ,  s1[4+c] = srol^(k-1)(seed[c ^ 2])   single-base steps for k % 4 != 0
struct HashTabs { uint64_t f4[256]; uint64_t r4[256]; uint64_t s1[8]; };

__device__ __forceinline__ void build_hash_tabs(HashTabs* H, const SketchTables& Tb, int k)
{
    for (int v = threadIdx.x; v < 256; v += blockDim.x) {
        uint64_t f = 0, r = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t c = (v >> (2 * j)) & 3;
            f = srol1(f) ^ sel4(c, Tb.seed);
            uint64_t sc = sel4(c ^ 2u, Tb.seed);
            for (int q = 0; q < j; q++) sc = srol1(sc);
            r ^= sc;
        }
        for (int q = 0; q < k - 4; q++) r = srol1(r);      // k >= 4 whenever r4 is used
        H->f4[v] = f;
        H->r4[v] = r;
    }
    if (threadIdx.x < 4) {
        H->s1[threadIdx.x] = Tb.seed[threadIdx.x];
        uint64_t sc = Tb.seed[threadIdx.x ^ 2];
        for (int q = 0; q < k - 1; q++) sc = srol1(sc);
        H->s1[4 + threadIdx.x] = sc;
    }
    __syncthreads();
}

// exact base hashes of the k-mer at p (all bases assumed valid), 4 bases per table lookup
__device__ __forceinline__ void kmer_hash64_tab(const uint32_t* __restrict__ pk, uint64_t p, int k, const HashTabs* H,
                                                uint64_t& fwd, uint64_t& rev)
{
    const uint64_t q = p >> 4;
    const uint32_t sh = ((uint32_t)p & 15u) * 2u;
    uint64_t f = 0, r = 0;
    uint32_t cur = pk_to_natural(__ldg(pk + q));
    int left = k;
    for (int m = 0; left > 0; m++) {
        uint32_t nxt = pk_to_natural(__ldg(pk + q + m + 1));
        uint32_t N = __funnelshift_r(cur, nxt, sh);      // 16 bases starting at p + 16 m, natural order
        cur = nxt;
        int nb = left < 16 ? left : 16;
        left -= nb;
        for (; nb >= 4; nb -= 4) {
            uint32_t v = N & 0xFFu;
            N >>= 8;
            f = sroln(f, 4) ^ H->f4[v];
            r = srorn(r, 4) ^ H->r4[v];
        }
        for (; nb > 0; nb--) {
            uint32_t c = N & 3u;
            N >>= 2;
            f = srol1(f) ^ H->s1[c];
            r = sror1(r) ^ H->s1[4 + c];
        }
    }
    fwd = f;
    rev = r;
}

// rank of position p in bitmap V (number of set bits strictly before p)
__device__ __forceinline__ uint64_t bitmap_rank(const uint32_t* __restrict__ bits, const uint64_t* __restrict__ prefix, uint64_t p)
{
    uint64_t wi = p >> 5;
    uint64_t blk = wi / RANK_BLOCK_WORDS;
    uint64_t r = prefix[blk];
    for (uint64_t j = blk * RANK_BLOCK_WORDS; j < wi; j++) r += __popc(bits[j]);
    uint32_t b = (uint32_t)p & 31u;
    if (b) r += __popc(bits[wi] & ((1u << b) - 1u));
    return r;
}

// position of the set bit with rank o (o < total)
__device__ __forceinline__ uint64_t bitmap_select(const uint32_t* __restrict__ bits, const uint64_t* __restrict__ prefix,
                                                  uint64_t n_blocks, uint64_t o)
{
    uint64_t lo = 0, hi = n_blocks;   // find last blk with prefix[blk] <= o
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (prefix[mid] <= o) lo = mid; else hi = mid;
    }
    uint64_t r = o - prefix[lo];
    uint64_t wi = lo * RANK_BLOCK_WORDS;
    for (;;) {
        uint32_t wv = bits[wi];
        uint32_t c = __popc(wv);
        if (r < c) return (wi << 5) + __fns(wv, 0, (int)r + 1);
        r -= c;
        wi++;
    }
}

// last c with offsets[c] <= p   (offsets has n_contigs+1 entries, offsets[0] == 0)
__device__ __forceinline__ uint32_t contig_of(const uint64_t* __restrict__ offsets, uint32_t n_contigs, uint64_t p)
{
    uint32_t lo = 0, hi = n_contigs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= p) lo = mid; else hi = mid;
    }
    return lo;
}

// warp-cooperative version: 32-ary search, two or three dependent loads instead of log2(n) (all lanes get the result)
__device__ __forceinline__ uint32_t contig_of_warp(const uint64_t* __restrict__ offsets, uint32_t n_contigs, uint64_t p, int lane)
{
    uint32_t lo = 0, hi = n_contigs;          // answer in [lo, hi): last c with offsets[c] <= p
    while (hi - lo > 1) {
        const uint32_t span = hi - lo;
        const uint32_t step = (span + 31) / 32;
        const uint32_t idx = lo + (uint32_t)(lane + 1) * step;      // probes lo+step, lo+2*step, ...
        const bool le = idx < hi && offsets[idx] <= p;
        const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, le));   // probes are monotone: first cnt are true
        const uint32_t nlo = lo + cnt * step;
        const uint32_t nhi = nlo + step < hi ? nlo + step : hi;
        lo = nlo;
        hi = nhi;
    }
    return lo;
}

// ---------------------------------------------------------------- pack: ASCII -> pk, V (+ rank counts of V)
// One thread per 32 bases.  Valid bytes are ACGTacgt; everything else is a "bad" base.
// Byte-SIMD validity: x = (byte & 0xDF) ^ 0x41 is one of 00 02 06 15 for A C G T.
__device__ __forceinline__ uint32_t bad_bytes(uint32_t x)
{
    uint32_t m = x & ~(x << 1) & 0x04040404u;          // T pattern: bit2 set, bit1 clear
    return (x & 0xF9F9F9F9u) ^ (m >> 2) ^ (m << 2);    // non-zero byte <=> invalid base
}

// 32 bases at word t: 2-bit codes (two pk words) and the exact bad-base mask (bases beyond n are bad)
__device__ __forceinline__ uint32_t pack_word(const uint8_t* __restrict__ seq, uint64_t n, uint64_t t, uint32_t pkw[2])
{
    const uint64_t p0 = t << 5;
    uint32_t bad = 0;
    pkw[0] = pkw[1] = 0;
    if (p0 + 32 <= n) {
        const uint4* src = reinterpret_cast<const uint4*>(seq + p0);
        uint4 a = __ldg(src), b = __ldg(src + 1);
        uint32_t wv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t anybad = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t x0 = (wv[4 * h + 0] & 0xDFDFDFDFu) ^ 0x41414141u;
            uint32_t x1 = (wv[4 * h + 1] & 0xDFDFDFDFu) ^ 0x41414141u;
            uint32_t x2 = (wv[4 * h + 2] & 0xDFDFDFDFu) ^ 0x41414141u;
            uint32_t x3 = (wv[4 * h + 3] & 0xDFDFDFDFu) ^ 0x41414141u;
            anybad |= bad_bytes(x0) | bad_bytes(x1) | bad_bytes(x2) | bad_bytes(x3);
            pkw[h] = ((x0 >> 1) & 0x03030303u) | ((x1 << 1) & 0x0C0C0C0Cu) | ((x2 << 3) & 0x30303030u) | ((x3 << 5) & 0xC0C0C0C0u);
        }
        if (anybad) {   // rare: exact per-byte mask
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t bb = bad_bytes((wv[j] & 0xDFDFDFDFu) ^ 0x41414141u);
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if ((bb >> (8 * q)) & 0xFFu) bad |= 1u << (4 * j + q);
            }
        }
    } else {
        for (int i = 0; i < 32; i++) {
            uint64_t p = p0 + i;
            uint32_t code = 0;
            bool ok = false;
            if (p < n) {
                uint32_t x = ((uint32_t)seq[p] & 0xDFu) ^ 0x41u;
                ok = (x == 0x00u) | (x == 0x02u) | (x == 0x06u) | (x == 0x15u);
                code = (x >> 1) & 3u;
            }
            if (!ok) { bad |= 1u << i; code = 0; }
            uint32_t ii = i & 15;
            pkw[i >> 4] |= code << (((ii & 3u) << 3) | ((ii >> 2) << 1));
        }
    }
    return bad;
}

// V word = positions p of this word with no bad base in [p, p+k); `nb(j)` returns the bad mask of word t+j
template <typename F>
__device__ __forceinline__ uint32_t vmask_from(F nb, int reach, int k)
{
    uint32_t any = 0;
    for (int j = 0; j <= reach; j++) any |= nb(j);
    if (!any) return 0xFFFFFFFFu;
    uint32_t v = 0;
    long long next_bad = -1;     // bit index (relative to this word) of the nearest bad base at or after the cursor
    for (int j = reach; j >= 0; j--) {
        const uint32_t bw = nb(j);
        for (int b = 31; b >= 0; b--) {
            const int idx = j * 32 + b;
            if ((bw >> b) & 1u) next_bad = idx;
            if (j == 0 && (next_bad < 0 || next_bad >= idx + k)) v |= 1u << b;
        }
    }
    return v;
}

// Words [t0, t1) (t0 a multiple of 32).  Each warp also classifies the `reach` words that follow its 32 words, so V
// needs no second pass over a bad-base bitmap; the host path shifts chunk ranges by one warp so that this halo has
// always landed.  Emits the rank-directory block counts of V (one warp = one 1024-bit block).
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* __restrict__ seq, SketchParams P,
                                                   uint32_t* __restrict__ pk, uint32_t* __restrict__ V, uint32_t* __restrict__ vcounts,
                                                   uint64_t t0, uint64_t t1)
{
    const uint64_t t = t0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint64_t tw = t - lane;                       // first word of this warp
    if (tw >= t1) return;                               // warp-uniform
    const int reach = (31 + P.k - 1) / 32;              // V of word t looks at words t .. t+reach   (reach <= 32)
    uint32_t pkw[2], tmp[2];
    uint32_t bad = 0xFFFFFFFFu, bad_halo = 0xFFFFFFFFu; // words beyond the sequence are all bad
    if (t < P.n_words) {
        bad = pack_word(seq, P.n, t, pkw);
        if (t < t1) reinterpret_cast<uint2*>(pk)[t] = make_uint2(pkw[0], pkw[1]);
    }
    {
        // halo = the `reach` words after this warp's 32: cooperative screening (one u32 per lane and round); only if a bad
        // base shows up (or the sequence ends inside the halo) do the first `reach` lanes classify their word exactly
        const uint64_t hb = (tw + 32) << 5;                 // first halo base
        bool suspicious = false;
        for (int i = lane; i < reach * 8; i += 32) {
            const uint64_t p = hb + 4 * (uint64_t)i;
            if (p + 4 <= P.n) {
                const uint32_t x = (__ldg(reinterpret_cast<const uint32_t*>(seq + p)) & 0xDFDFDFDFu) ^ 0x41414141u;
                suspicious |= bad_bytes(x) != 0;
            } else suspicious = true;
        }
        if (__any_sync(0xffffffffu, suspicious)) {
            if (lane < reach && tw + 32 + lane < P.n_words) bad_halo = pack_word(seq, P.n, tw + 32 + lane, tmp);
        } else {
            bad_halo = 0;
        }
    }
    uint32_t any = bad;
    for (int j = 1; j <= reach; j++) {
        const int src = lane + j;
        const uint32_t a = __shfl_sync(0xffffffffu, bad, src & 31), h = __shfl_sync(0xffffffffu, bad_halo, src & 31);
        any |= src < 32 ? a : h;
    }
    uint32_t v = 0xFFFFFFFFu;
    const bool slow = any != 0;
    if (__any_sync(0xffffffffu, slow)) {                // rare: some lane has a bad base within reach
        uint32_t nbw[33];
        nbw[0] = bad;
        for (int j = 1; j <= reach; j++) {
            const int src = lane + j;
            const uint32_t a = __shfl_sync(0xffffffffu, bad, src & 31), h = __shfl_sync(0xffffffffu, bad_halo, src & 31);
            nbw[j] = src < 32 ? a : h;
        }
        if (slow) v = vmask_from([&](int j) { return nbw[j]; }, reach, P.k);
    }
    if (t >= P.n_words) v = 0;
    uint32_t c = __popc(v);
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (t < P.n_words && t < t1) V[t] = v;
    if (lane == 0 && tw < P.n_words) vcounts[tw >> 5] = c;
}

// contig boundaries: a k-mer may not straddle two records
__global__ void boundary_kernel(const uint64_t* __restrict__ offsets, uint32_t n_contigs, SketchParams P, uint32_t* __restrict__ V,
                                uint32_t* __restrict__ vcounts)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (c >= n_contigs) return;
    uint64_t q = offsets[c];
    if (q == 0 || q >= P.n) return;
    uint64_t lo = q >= (uint64_t)(P.k - 1) ? q - (P.k - 1) : 0;
    for (uint64_t p = lo; p < q; p++) {
        uint32_t bit = 1u << (p & 31);
        if (atomicAnd(&V[p >> 5], ~bit) & bit) atomicSub(&vcounts[p >> 10], 1u);   // keep the block count exact
    }
}

// ostart[c] = ordinal of the first valid k-mer at or after offsets[c]  (n_contigs+1 entries)
__global__ void contig_bounds_kernel(const uint64_t* __restrict__ offsets, uint32_t n_contigs, SketchParams P,
                                     const uint32_t* __restrict__ V, const uint64_t* __restrict__ vprefix,
                                     uint64_t* __restrict__ ostart)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_contigs) return;
    uint64_t q = offsets[c];
    ostart[c] = q >= P.n ? vprefix[(P.n_words + RANK_BLOCK_WORDS - 1) / RANK_BLOCK_WORDS] : bitmap_rank(V, vprefix, q);
}

// ---------------------------------------------------------------- cand (generic): exact 64-bit test, any k
__global__ void __launch_bounds__(128) cand_generic_kernel(const uint32_t* __restrict__ pk, const uint32_t* __restrict__ V,
                                                            SketchParams P, SketchTables Tb, uint32_t* __restrict__ C)
{
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t p0 = tid * (uint64_t)P.chunk;
    if (p0 >= P.n) return;
    const int k = P.k;
    uint64_t f, r;
    kmer_hash64(pk, p0, k, Tb.seed, f, r);
    const int n_w = P.chunk / 32;
    for (int wi = 0; wi < n_w; wi++) {
        uint64_t pw = p0 + (uint64_t)wi * 32;
        if (pw >= P.n) break;
        uint32_t bits = 0;
        for (int b = 0; b < 32; b++) {
            uint64_t p = pw + b;
            uint64_t h0 = canon(f, r, P.canon_min);
            if ((uint32_t)(h0 >> 33) <= P.T) bits |= 1u << b;
            uint32_t o = pk_code(pk, p), in = pk_code(pk, p + k);
            f = srol1(f) ^ sel4(o, Tb.seed_rolk) ^ sel4(in, Tb.seed);
            r = sror1(r ^ sel4(in ^ 2u, Tb.seed_rolk) ^ sel4(o ^ 2u, Tb.seed));
        }
        C[pw >> 5] = bits & V[pw >> 5];
    }
}

// ---------------------------------------------------------------- cand31: 31-bit lane prefilter, k % 4 == 0
// Only the upper rotation group (bits 63:33) of fwd and rev is rolled, in 32-bit registers.
//   sum mode: t = (f31 + r31 + carry) mod 2^31, carry in {0,1}  =>  superset test (f31+r31+1) mod 2^31 <= T+1
//   min mode: t = min(f31, r31)
// Register layouts chosen so that each roll is two ALU ops + one table XOR:
//   F  "low aligned": value in bits 30:0, bit 31 is don't-care ("dirty")
//   R  "top aligned": value in bits 31:1, bit 0 is dirty
//   a = F << 1 is both the top-aligned clean copy of F (for the sum) and the funnel-shift source.
// The 16-entry table (idx = out<<2 | in) holds {fwd term low aligned, rev term top aligned}; it is 128 B,
// one entry per bank pair, so any mix of indices in a warp is conflict free.
// FMA_OFFLOAD: the kernel is bound by the ALU pipe (LOP3/SHF/IADD3/ISETP/PRMT) while the FMA pipe idles, so the two
// additions of the test are issued as IMADs with run-time multipliers (1 and 2, opaque to ptxas, which would otherwise
// turn them back into IADD3/VIADD): key = F*2 + R, key += 2, and the predicated accumulation gb += 2^i.
#define MXE_CAND_STEP(i)                                                                              \
    {                                                                                                 \
        const uint32_t a = F << 1;                                                                    \
        if (FMA_OFFLOAD && !CANON_MIN) {                                                              \
            uint32_t key;                                                                             \
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(key) : "r"(F), "r"(m_two), "r"(R));               \
            asm("mad.lo.u32 %0, %1, %2, 2;" : "=r"(key) : "r"(key), "r"(m_one));                      \
            asm("{ .reg .pred p; setp.le.u32 p, %1, %2; @p mad.lo.u32 %0, %3, %4, %0; }"              \
                : "+r"(gb) : "r"(key), "r"(Tt), "r"(m_one), "n"(1u << (i)));                          \
        } else {                                                                                      \
            const uint32_t key = CANON_MIN ? min(a, R) : a + R + 2u;                                  \
            asm("{ .reg .pred p; setp.le.u32 p, %1, %2; @p or.b32 %0, %0, %3; }"                      \
                : "+r"(gb) : "r"(key), "r"(Tt), "n"(1u << (i)));                                      \
        }                                                                                             \
        const uint32_t addr = __byte_perm(zq[(i) >> 2], 0, 0x4440 | ((i) & 3));                       \
        const uint2 e = *reinterpret_cast<const uint2*>(tb + addr);                                   \
        F = __funnelshift_l(a, F, 1) ^ e.x;                                                           \
        const uint32_t u = R ^ e.y;                                                                   \
        R = __funnelshift_r(u, u >> 1, 1);                                                            \
    }

// Shared-memory staging: every thread walks its own run of `chunk` consecutive positions, which as
// direct global loads is one scattered 4-byte access per lane (measured: L1TEX 99 % busy, 10x DRAM
// over-fetch).  The CTA instead copies its contiguous slice of pk (plus a k-base halo) and of V into
// shared memory with coalesced loads; rows are padded by one word per thread-run so that the per-thread
// walk (lane stride = run length + 1 words, odd) is bank-conflict free.  C is written back through the
// same buffer, coalesced.
constexpr int CAND_THREADS = 128;

__host__ __device__ inline size_t cand31_smem_bytes(int chunk, int k)
{
    const int wpt = chunk / 16, wv = chunk / 32;
    const size_t n_pk = (size_t)CAND_THREADS * wpt + (size_t)(k / 16) + 2;
    const size_t n_v = (size_t)CAND_THREADS * wv;
    return (32 + n_pk + n_pk / wpt + 1 + n_v + n_v / wv + 1) * sizeof(uint32_t);
}

template <int CANON_MIN, int FMA_OFFLOAD>
__global__ void __launch_bounds__(CAND_THREADS) cand31_kernel(const uint32_t* __restrict__ pk, const uint32_t* __restrict__ V,
                                                               SketchParams P, SketchTables Tb, uint32_t* __restrict__ C,
                                                               uint32_t* __restrict__ ccounts)
{
    extern __shared__ uint32_t smem[];
    uint2* tab = reinterpret_cast<uint2*>(smem);                 // 16 entries
    const int wpt = P.chunk >> 4, wv = P.chunk >> 5;             // pk / V words per thread-run (powers of two)
    const int lw = 31 - __clz(wpt), lv = 31 - __clz(wv);
    const int k = P.k;
    const int kq = k >> 4, ks = (k & 15) >> 2;
    const uint32_t n_pk = CAND_THREADS * wpt + kq + 2;
    uint32_t* pkS = smem + 32;
    uint32_t* vcS = pkS + n_pk + (n_pk >> lw) + 1;
    const uint32_t n_v = CAND_THREADS * wv;

    if (threadIdx.x < 16) { uint2 e = Tb.t16[threadIdx.x]; e.y <<= 1; tab[threadIdx.x] = e; }
    const uint64_t cta_p0 = (uint64_t)blockIdx.x * CAND_THREADS * (uint64_t)P.chunk;
    const uint64_t q_cta = cta_p0 >> 4, v_cta = cta_p0 >> 5;
    for (uint32_t i = threadIdx.x; i < n_pk; i += CAND_THREADS) {
        uint64_t gw = q_cta + i;
        pkS[i + (i >> lw)] = gw < P.pk_words ? __ldg(pk + gw) : 0u;
    }
    for (uint32_t i = threadIdx.x; i < n_v; i += CAND_THREADS) {
        uint64_t gv = v_cta + i;
        vcS[i + (i >> lv)] = gv < P.n_words ? __ldg(V + gv) : 0u;
    }
    __syncthreads();

    const uint64_t p0 = cta_p0 + (uint64_t)threadIdx.x * P.chunk;
    if (p0 < P.n) {
        // warm-up by rolling, without a special case: start from the all-A k-mer and roll ceil(k/16) groups with out = A;
        // when k is not a multiple of 16 the first group is fed 16 - k % 16 further A's in front of the first real bases
        // (rolling A in and A out leaves the all-A k-mer unchanged)
        const int kqc = kq + (ks ? 1 : 0);
        uint32_t F = Tb.f0, R = Tb.r0;
        int g = -kqc;
        R <<= 1;
        const uint32_t lowmask = ks == 1 ? 0x3F3F3F3Fu : ks == 2 ? 0x0F0F0F0Fu : 0x03030303u;
        const char* tb = reinterpret_cast<const char*>(tab);
        const uint32_t Tt = CANON_MIN ? ((P.T << 1) | 1u) : (((P.T + 1u) << 1) | 1u);
        const uint32_t m_one = P.mul1, m_two = P.mul2;       // 1 and 2 as run-time values (see MXE_CAND_STEP)
        const int n_g = P.chunk / 16;
        const uint32_t row = threadIdx.x * wpt;
        const uint32_t vrow = threadIdx.x * wv + threadIdx.x;   // padded index of this run's first V word
        uint32_t bits = 0;
        for (; g < n_g; g++) {
            const uint32_t io = row + g, ii = row + g + kq;
            uint32_t o = g < 0 ? 0u : pkS[io + (io >> lw)];
            uint32_t in;
            if (ks) {
                const uint32_t lo = g == -kqc ? 0u : pkS[ii + (ii >> lw)];      // first warm-up group: the padding A's
                const uint32_t b2 = pkS[ii + 1 + ((ii + 1) >> lw)];
                in = ((lo >> (2 * ks)) & lowmask) | ((b2 << (8 - 2 * ks)) & ~lowmask);
            } else {
                in = pkS[ii + (ii >> lw)];
            }
            // table byte offsets (idx << 3), one per byte: zq[j] serves positions 4j..4j+3 of the group
            uint32_t z1 = ((o << 2) & 0xCCCCCCCCu) | (in & 0x33333333u);
            uint32_t z2 = (o & 0xCCCCCCCCu) | ((in >> 2) & 0x33333333u);
            uint32_t zq[4];
            zq[0] = (z1 << 3) & 0x78787878u;
            zq[1] = (z2 << 3) & 0x78787878u;
            zq[2] = (z1 >> 1) & 0x78787878u;
            zq[3] = (z2 >> 1) & 0x78787878u;
            uint32_t gb = 0;
            MXE_CAND_STEP(0) MXE_CAND_STEP(1) MXE_CAND_STEP(2) MXE_CAND_STEP(3)
            MXE_CAND_STEP(4) MXE_CAND_STEP(5) MXE_CAND_STEP(6) MXE_CAND_STEP(7)
            MXE_CAND_STEP(8) MXE_CAND_STEP(9) MXE_CAND_STEP(10) MXE_CAND_STEP(11)
            MXE_CAND_STEP(12) MXE_CAND_STEP(13) MXE_CAND_STEP(14) MXE_CAND_STEP(15)
            if (g < 0) continue;
            if (g & 1) {
                bits |= gb << 16;
                vcS[vrow + (g >> 1)] &= bits;        // C = candidates & valid starts (bits beyond n are not valid starts)
            } else {
                bits = gb;
            }
        }
    } else {
        for (int j = 0; j < wv; j++) vcS[threadIdx.x * wv + threadIdx.x + j] = 0;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_v; i += CAND_THREADS) {     // n_v is a multiple of 128: warps stay converged
        uint64_t gv = v_cta + i;
        uint32_t cw = gv < P.n_words ? vcS[i + (i >> lv)] : 0u;
        if (gv < P.n_words) C[gv] = cw;
        uint32_t c = __popc(cw);
#pragma unroll
        for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if ((threadIdx.x & 31) == 0 && gv < P.n_words) ccounts[gv >> 5] = c;   // rank-directory block count of C
    }
}
#undef MXE_CAND_STEP

// Sizes that are only known on the device.  The host sizes arrays and grids from upper bounds and passes the pointer to the
// exact count; kernels clamp it to the bound (an overflow is detected at the single host round trip at the end of the sketch,
// which then repeats the call with exact sizes).  d_n == nullptr: n_max is the exact count.
__device__ __forceinline__ uint64_t dev_count(const uint64_t* __restrict__ d_n, uint64_t n_max)
{
    if (!d_n) return n_max;
    const uint64_t v = *d_n;
    return v < n_max ? v : n_max;
}

// ---------------------------------------------------------------- candidate extraction + evaluation
// One warp per 8192 bits of C (8 rank blocks), one lane per 8 consecutive words: two 16-byte loads of C and of V per
// lane are in flight together and one warp scan serves eight rank blocks (the one-warp-per-block version was latency
// bound: a chain of dependent loads per 9 candidates).  Emits ordered positions, valid-k-mer ordinals (popcounts of V,
// no per-candidate rank walk) and record ids (one 32-ary search per warp, then a monotone walk).  Ordinals are stored
// PADDED: ordinal + record * w, so that candidates of different records are never within w of each other and the
// window scans of select_kernel need no record test.
constexpr int XW = 8;                        // words per lane
constexpr int XBLOCKS = XW * 32 / RANK_BLOCK_WORDS;   // rank blocks per warp (8)

__device__ __forceinline__ void load_words8(const uint32_t* __restrict__ bits, uint64_t w0, uint64_t n_words, uint32_t out[XW])
{
    if (w0 + XW <= n_words) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(bits + w0)), b = __ldg(reinterpret_cast<const uint4*>(bits + w0) + 1);
        out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < XW; j++) out[j] = w0 + j < n_words ? __ldg(bits + w0 + j) : 0u;
    }
}

__global__ void __launch_bounds__(256) cand_extract_kernel(const uint32_t* __restrict__ C, const uint32_t* __restrict__ V, uint64_t n_words,
                                                            const uint64_t* __restrict__ cprefix, const uint64_t* __restrict__ vprefix, uint64_t n_blocks,
                                                            const uint64_t* __restrict__ offsets, uint32_t n_contigs, uint64_t w,
                                                            uint64_t* __restrict__ cpos, uint64_t* __restrict__ cord, uint32_t* __restrict__ cctg, uint64_t cap)
{
    const uint64_t sb = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const uint64_t blk0 = sb * XBLOCKS;
    if (blk0 >= n_blocks) return;
    const uint64_t blk1 = blk0 + XBLOCKS < n_blocks ? blk0 + XBLOCKS : n_blocks;
    const uint64_t c_base = cprefix[blk0];
    if (cprefix[blk1] == c_base) return;             // warp-uniform
    const uint64_t w0 = sb * (32 * XW) + (uint64_t)lane * XW;
    uint32_t cw[XW], vw[XW];
    load_words8(C, w0, n_words, cw);
    load_words8(V, w0, n_words, vw);
    uint32_t cc = 0, vc = 0;
#pragma unroll
    for (int j = 0; j < XW; j++) { cc += __popc(cw[j]); vc += __popc(vw[j]); }
    uint32_t cx = cc, vx = vc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y1 = __shfl_up_sync(0xffffffffu, cx, d), y2 = __shfl_up_sync(0xffffffffu, vx, d);
        if (lane >= d) { cx += y1; vx += y2; }
    }
    uint32_t c = contig_of_warp(offsets, n_contigs, sb * (uint64_t)(32 * XW * 32), lane);
    if (cc == 0) return;                              // after the collectives
    uint64_t o = c_base + (cx - cc);
    uint64_t nextb = c + 1 < n_contigs ? offsets[c + 1] : ~0ULL;     // start of the next record
    uint64_t vb = vprefix[blk0] + (vx - vc) + (uint64_t)c * w;       // padded ordinal base (see select_kernel)
#pragma unroll
    for (int j = 0; j < XW; j++) {
        uint32_t wv = cw[j];
        const uint64_t base = (w0 + j) << 5;
        while (wv) {
            const int b = __ffs(wv) - 1;
            wv &= wv - 1;
            const uint64_t p = base + b;
            while (p >= nextb) { c++; vb += w; nextb = c + 1 < n_contigs ? offsets[c + 1] : ~0ULL; }
            if (o < cap) {
                cpos[o] = p;
                cord[o] = vb + __popc(vw[j] & ((1u << b) - 1u));
                cctg[o] = c;
            }
            o++;
        }
        vb += __popc(vw[j]);
    }
}

__global__ void __launch_bounds__(256) cand_hash_kernel(const uint64_t* __restrict__ cpos, const uint64_t* __restrict__ d_n, uint64_t n_max,
                                                         const uint32_t* __restrict__ pk, SketchParams P, SketchTables Tb, uint64_t* __restrict__ h0)
{
    __shared__ HashTabs H;
    build_hash_tabs(&H, Tb, P.k);
    const uint64_t n_cand = dev_count(d_n, n_max);
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cand) return;
    uint64_t f, r;
    kmer_hash64_tab(pk, cpos[i], P.k, &H, f, r);
    h0[i] = canon(f, r, P.canon_min);
}

// ---------------------------------------------------------------- exact hash with position-specific tables
// hash = XOR over 4-base groups of a table entry that already carries the group's rotation, so the per-candidate work
// is one byte extract, two LDS.64 and two 64-bit XORs per group (no rotations).  Tables: PF[g][v] = srol^(4(G-1-g))(f4[v]),
// PR[g][v] = sror^(4(G-1-g))(r4[v]), G = k/4 groups; the k % 4 trailing bases use the single-base recurrences.
constexpr int HASHPOS_MAX_GROUPS = 10;       // 2 * 10 * 2 KB = 40 KB of shared memory (k <= 43)

__global__ void __launch_bounds__(256) hash_pos_tables_kernel(SketchTables Tb, int k, uint64_t* __restrict__ PF, uint64_t* __restrict__ PR)
{
    __shared__ HashTabs H;
    build_hash_tabs(&H, Tb, k);
    const int G = k / 4;
    for (int idx = threadIdx.x + blockIdx.x * blockDim.x; idx < G * 256; idx += blockDim.x * gridDim.x) {
        const int g = idx >> 8, v = idx & 255;
        uint64_t f = H.f4[v], r = H.r4[v];
        for (int q = 0; q < 4 * (G - 1 - g); q++) { f = srol1(f); r = sror1(r); }
        PF[idx] = f;
        PR[idx] = r;
    }
}

// exact base hashes of the k-mer at p from the position-specific tables staged in shared memory (pf, pr: G x 256).
// k <= 4 * HASHPOS_MAX_GROUPS + 3 = 43: the k-mer spans at most four pk words, which are all requested up front so that
// the (scattered) loads of one candidate overlap instead of following each other.
__device__ __forceinline__ void kmer_hash64_pos(const uint32_t* __restrict__ pk, uint64_t p, int k, const uint64_t* pf, const uint64_t* pr,
                                                const uint64_t* s1, uint64_t& fwd, uint64_t& rev)
{
    const uint64_t q = p >> 4;
    const uint32_t sh = ((uint32_t)p & 15u) * 2u;
    const int nw = ((k + 15) >> 4) + 1;
    uint32_t wd[4];
#pragma unroll
    for (int m = 0; m < 4; m++) wd[m] = m < nw ? __ldg(pk + q + m) : 0u;
#pragma unroll
    for (int m = 0; m < 4; m++) wd[m] = pk_to_natural(wd[m]);
    uint64_t f = 0, r = 0;
    int g = 0, left = k;
#pragma unroll
    for (int m = 0; m < 3; m++) {
        if (left <= 0) break;
        uint32_t N = __funnelshift_r(wd[m], wd[m + 1], sh);      // 16 bases starting at p + 16 m, natural order
        int nb = left < 16 ? left : 16;
        left -= nb;
        for (; nb >= 4; nb -= 4, g++) {
            const uint32_t v = N & 0xFFu;
            N >>= 8;
            f ^= pf[g * 256 + v];
            r ^= pr[g * 256 + v];
        }
        for (; nb > 0; nb--) {                               // k % 4 trailing bases (last word only)
            const uint32_t c = N & 3u;
            N >>= 2;
            f = srol1(f) ^ s1[c];
            r = sror1(r) ^ s1[4 + c];
        }
    }
    fwd = f;
    rev = r;
}

// stage the position-specific tables (and the single-base terms for k % 4 trailing bases) in shared memory.
// The two tables (G x 2 KB each) come in as two asynchronous bulk copies (cp.async.bulk, the 1-D form of TMA) that one
// thread issues against an mbarrier; everybody waits on the barrier's phase.  No load / store instruction per element,
// and the copy engine of the SM does the address arithmetic (the loop this replaces: 16 x (LDG.64 + STS.64) per thread).
// pf, pr: 16-byte aligned shared memory; PF, PR: 16-byte aligned global memory (arena blocks are 512-byte aligned).
__device__ __forceinline__ void stage_pos_tables(const SketchTables& Tb, int k, const uint64_t* __restrict__ PF, const uint64_t* __restrict__ PR,
                                                 uint64_t* pf, uint64_t* pr, uint64_t* s1)
{
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bytes = (uint32_t)(k / 4) * 256u * (uint32_t)sizeof(uint64_t);
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(2u * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(pf)), "l"(PF), "r"(bytes), "r"(bar_a) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(pr)), "l"(PR), "r"(bytes), "r"(bar_a) : "memory");
    }
    if (threadIdx.x < 4) {
        s1[threadIdx.x] = Tb.seed[threadIdx.x];
        uint64_t sc = Tb.seed[threadIdx.x ^ 2];
        for (int q = 0; q < k - 1; q++) sc = srol1(sc);
        s1[4 + threadIdx.x] = sc;
    }
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; spin++) {              // phase 0 of the barrier completes when both copies have landed
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_a), "r"(0u) : "memory");
        if (spin > (1u << 22)) __trap();                  // a copy that never lands must not hang the GPU
    }
    __syncthreads();                                       // s1
}

__global__ void __launch_bounds__(256) cand_hash_pos_kernel(const uint64_t* __restrict__ cpos, const uint64_t* __restrict__ d_n, uint64_t n_max,
                                                             const uint32_t* __restrict__ pk, SketchParams P, SketchTables Tb, const uint64_t* __restrict__ PF,
                                                             const uint64_t* __restrict__ PR, uint64_t* __restrict__ h0)
{
    extern __shared__ __align__(16) uint64_t hp[];
    const int G = P.k / 4;
    uint64_t* pf = hp;
    uint64_t* pr = hp + G * 256;
    __shared__ uint64_t s1[8];
    stage_pos_tables(Tb, P.k, PF, PR, pf, pr, s1);
    const uint64_t n_cand = dev_count(d_n, n_max);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cand; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t f, r;
        kmer_hash64_pos(pk, cpos[i], P.k, pf, pr, s1, f, r);
        h0[i] = canon(f, r, P.canon_min);
    }
}

// ---------------------------------------------------------------- candidate pruning on 31-bit bounds
// Most candidates (~1 % of positions) are not minimizers (~0.2 %).  Before paying for exact 64-bit hashes, each
// candidate gets cheap bounds lo <= t <= hi on t = hash0 >> 33 from the 31-bit lanes alone (sum: t in {s1-1, s1},
// s1 = (f31+r31+1) mod 2^31; min: t = min(f31,r31) exactly).  Candidate j DOMINATES i when hi_j < lo_i (then
// hash0_j < hash0_i whatever the low bits are).  A candidate that has a dominator inside every window containing it
// can never be a window arg-min and is dropped; every window's true candidate arg-min survives (nothing dominates
// it), so the exact selection over the survivors returns the same arg-min for every window.
struct HashTabs31 { uint32_t f4[256]; uint32_t r4[256]; uint32_t s1[8]; };

__device__ __forceinline__ void build_hash_tabs31(HashTabs31* H, const SketchTables& Tb, int k)
{
    for (int v = threadIdx.x; v < 256; v += blockDim.x) {
        uint32_t f = 0, r = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t c = (v >> (2 * j)) & 3;
            f = rol31(f) ^ Tb.shi[c];
            uint32_t sc = Tb.shi[c ^ 2u];
            for (int q = 0; q < j; q++) sc = rol31(sc);
            r ^= sc;
        }
        for (int q = 0; q < k - 4; q++) r = rol31(r);
        H->f4[v] = f;
        H->r4[v] = r;
    }
    if (threadIdx.x < 4) {
        H->s1[threadIdx.x] = Tb.shi[threadIdx.x];
        uint32_t sc = Tb.shi[threadIdx.x ^ 2];
        for (int q = 0; q < k - 1; q++) sc = rol31(sc);
        H->s1[4 + threadIdx.x] = sc;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) cand_key31_kernel(const uint64_t* __restrict__ cpos, uint64_t n_cand, const uint32_t* __restrict__ pk,
                                                          SketchParams P, SketchTables Tb, uint32_t* __restrict__ klo, uint32_t* __restrict__ khi)
{
    __shared__ HashTabs31 H;
    build_hash_tabs31(&H, Tb, P.k);
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cand) return;
    const uint64_t p = cpos[i];
    const uint64_t q = p >> 4;
    const uint32_t sh = ((uint32_t)p & 15u) * 2u;
    uint32_t f = 0, r = 0;
    uint32_t cur = pk_to_natural(__ldg(pk + q));
    int left = P.k;
    for (int m = 0; left > 0; m++) {
        uint32_t nxt = pk_to_natural(__ldg(pk + q + m + 1));
        uint32_t N = __funnelshift_r(cur, nxt, sh);
        cur = nxt;
        int nb = left < 16 ? left : 16;
        left -= nb;
        for (; nb >= 4; nb -= 4) {
            uint32_t v = N & 0xFFu;
            N >>= 8;
            f = (((f << 4) | (f >> 27)) & 0x7FFFFFFFu) ^ H.f4[v];
            r = (((r >> 4) | (r << 27)) & 0x7FFFFFFFu) ^ H.r4[v];
        }
        for (; nb > 0; nb--) {
            uint32_t c = N & 3u;
            N >>= 2;
            f = rol31(f) ^ H.s1[c];
            r = ror31(r) ^ H.s1[4 + c];
        }
    }
    uint32_t lo, hi;
    if (P.canon_min) { lo = hi = min(f, r); }
    else {
        uint32_t s1 = (f + r + 1u) & 0x7FFFFFFFu;
        lo = s1 ? s1 - 1u : 0u;
        hi = s1 ? s1 : 0x7FFFFFFFu;
    }
    klo[i] = lo;
    khi[i] = hi;
}

__global__ void __launch_bounds__(256) prune_kernel(const uint32_t* __restrict__ klo, const uint32_t* __restrict__ khi,
                                                     const uint64_t* __restrict__ gord, const uint32_t* __restrict__ ctg, uint64_t n_cand,
                                                     const uint64_t* __restrict__ ostart, SketchParams P, uint32_t* __restrict__ flag)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cand) return;
    const uint32_t c = ctg[i];
    const uint64_t w = (uint64_t)P.w, pad = (uint64_t)c * w;      // gord holds padded ordinals
    const uint64_t os = ostart[c] + pad, oe = ostart[c + 1] + pad;
    uint32_t keep = 0;
    if (oe - os >= w) {
        const uint64_t o = gord[i];
        const uint32_t lo = klo[i];
        const uint64_t jmin = os + w - 1, jmax = oe - 1;
        uint64_t jlo = o > jmin ? o : jmin;
        uint64_t jhi = o + w - 1 < jmax ? o + w - 1 : jmax;
        for (uint64_t j = i; j-- > 0;) {
            const uint64_t oj = gord[j];
            if (oj + w <= o) break;
            if (khi[j] < lo) { uint64_t b = oj + w; if (b > jlo) jlo = b; break; }
        }
        if (jlo <= jhi)
            for (uint64_t j = i + 1; j < n_cand; j++) {
                const uint64_t oj = gord[j];
                if (oj >= o + w) break;
                if (khi[j] < lo) { uint64_t b = oj - 1; if (b < jhi) jhi = b; break; }
            }
        keep = jlo <= jhi;
    }
    flag[i] = keep;
}

__global__ void __launch_bounds__(256) cand_compact_kernel(const uint32_t* __restrict__ flag, const uint64_t* __restrict__ prefix, uint64_t n_cand,
                                                            const uint64_t* __restrict__ cpos, const uint64_t* __restrict__ gord, const uint32_t* __restrict__ ctg,
                                                            uint64_t* __restrict__ cpos2, uint64_t* __restrict__ gord2, uint32_t* __restrict__ ctg2)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cand || !flag[i]) return;
    uint64_t o = prefix[i];
    cpos2[o] = cpos[i];
    gord2[o] = gord[i];
    ctg2[o] = ctg[i];
}

// ---------------------------------------------------------------- select
struct Gap { uint64_t ja, jb; };

struct GapList {
    Gap* items;
    unsigned long long* count;      // attempted pushes
    unsigned long long* windows;    // sum of gap lengths
    uint64_t capacity;
};

__device__ __forceinline__ void push_gap(const GapList& G, uint64_t ja, uint64_t jb)
{
    unsigned long long slot = atomicAdd(G.count, 1ULL);
    atomicAdd(G.windows, (unsigned long long)(jb - ja + 1));
    if (slot < G.capacity) G.items[slot] = Gap{ja, jb};
}

// gord holds PADDED ordinals (ordinal + record * w): two candidates of different records are at least w apart, so the
// window scans stop at record boundaries by themselves and only the candidate's own record id is looked up.
// The kernel is instruction bound (ALU pipe 75 %), so when every padded ordinal (+ 2w) fits 32 bits -- any input below
// ~4 G valid k-mers -- the NARROW variant reads only the low halves of the ordinals and does all window arithmetic and
// indexing in 32 bits.
template <bool NARROW>
__global__ void __launch_bounds__(256) select_kernel(const uint64_t* __restrict__ cpos, const uint64_t* __restrict__ h0,
                                                      const uint64_t* __restrict__ gord, const uint32_t* __restrict__ ctg,
                                                      const uint64_t* __restrict__ d_n, uint64_t n_max, const uint64_t* __restrict__ ostart,
                                                      SketchParams P, uint32_t* __restrict__ M, GapList G)
{
    const uint64_t n_cand = dev_count(d_n, n_max);
    typedef typename std::conditional<NARROW, uint32_t, uint64_t>::type ord_t;   // ordinals
    typedef typename std::conditional<NARROW, uint32_t, uint64_t>::type idx_t;   // candidate indices
    const idx_t i = (idx_t)((uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
    const idx_t n = (idx_t)n_cand;
    if ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x >= n_cand) return;
    auto ord = [&](idx_t j) -> ord_t {
        if (NARROW) return (ord_t) reinterpret_cast<const uint32_t*>(gord)[2 * (size_t)j];   // low half (little endian)
        return (ord_t)gord[j];
    };
    const uint32_t c = ctg[i];
    const uint64_t pad64 = (uint64_t)c * (uint64_t)P.w;
    const ord_t w = (ord_t)P.w, pad = (ord_t)pad64;
    const ord_t os = (ord_t)(ostart[c] + pad64), oe = (ord_t)(ostart[c + 1] + pad64);
    if ((ord_t)(oe - os) < w) return;             // record has fewer than w valid k-mers: no window
    const ord_t o = ord(i);
    const uint64_t h = h0[i];
    const ord_t jmin = os + w - 1, jmax = oe - 1;

    ord_t jlo = o > jmin ? o : jmin;
    ord_t jhi = o + w - 1 < jmax ? o + w - 1 : jmax;
    // nearest strictly smaller to the left within the window span
    for (idx_t j = i; j-- > 0;) {
        const ord_t oj = ord(j);
        if (oj + w <= o) break;
        if (h0[j] < h) { const ord_t b = oj + w; if (b > jlo) jlo = b; break; }
    }
    // nearest smaller-or-equal to the right (rightmost wins ties)
    const ord_t o_next = i + 1 < n ? ord(i + 1) : (ord_t)~(ord_t)0;
    for (idx_t j = i + 1; j < n; j++) {
        const ord_t oj = ord(j);
        if (oj >= o + w) break;
        if (h0[j] <= h) { const ord_t b = oj - 1; if (b < jhi) jhi = b; break; }
    }
    if (jlo <= jhi) {
        if ((uint32_t)(h >> 33) <= P.T && h != ~0ULL) {
            const uint64_t p = cpos[i];
            atomicOr(&M[p >> 5], 1u << (p & 31));
        } else {
            push_gap(G, (uint64_t)(ord_t)(jlo - pad), (uint64_t)(ord_t)(jhi - pad));
        }
    }
    // candidate-free windows to the right of this candidate (a candidate of the next record lies beyond jmax)
    {
        const ord_t ga = o + w > jmin ? o + w : jmin;
        const ord_t gb = (ord_t)(o_next - 1) < jmax ? (ord_t)(o_next - 1) : jmax;
        if (ga <= gb) push_gap(G, (uint64_t)(ord_t)(ga - pad), (uint64_t)(ord_t)(gb - pad));
    }
    // ... and to the left of the first candidate of the record (the previous candidate lies before os)
    if (i == 0 || ord(i - 1) < os) {
        if (o > jmin) push_gap(G, (uint64_t)(ord_t)(jmin - pad), (uint64_t)(ord_t)(((ord_t)(o - 1) < jmax ? (ord_t)(o - 1) : jmax) - pad));
    }
}

// records that have windows but no candidate at all
__global__ void empty_contig_gap_kernel(const uint64_t* __restrict__ cpos, const uint64_t* __restrict__ d_n, uint64_t n_max,
                                        const uint64_t* __restrict__ offsets, uint32_t n_contigs,
                                        const uint64_t* __restrict__ ostart, SketchParams P, GapList G)
{
    const uint64_t n_cand = dev_count(d_n, n_max);
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_contigs) return;
    uint64_t os = ostart[c], oe = ostart[c + 1];
    if (oe - os < (uint64_t)P.w) return;
    uint64_t a = offsets[c], b = offsets[c + 1];
    uint64_t lo = 0, hi = n_cand;      // lower_bound(cpos, a)
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (cpos[mid] < a) lo = mid + 1; else hi = mid;
    }
    if (lo == n_cand || cpos[lo] >= b) push_gap(G, os + P.w - 1, oe - 1);
}

// ---------------------------------------------------------------- gap: dense exact windows
constexpr int GAP_CHUNK = 1024;   // window ends per chunk
constexpr int GAP_RUN = 16;       // consecutive ordinals hashed by one thread (rolling)

__global__ void __launch_bounds__(256) gap_kernel(const Gap* __restrict__ gaps, const unsigned long long* __restrict__ d_n_gaps, uint64_t n_gaps_max,
                                                   const uint32_t* __restrict__ pk, const uint32_t* __restrict__ V,
                                                   const uint64_t* __restrict__ vprefix, uint64_t n_vblocks,
                                                   SketchParams P, SketchTables Tb,
                                                   uint64_t* __restrict__ scratch_h, uint64_t* __restrict__ scratch_p,
                                                   uint64_t* __restrict__ scratch_bm, uint64_t* __restrict__ scratch_bi,
                                                   uint32_t* __restrict__ M)
{
    __shared__ HashTabs Ht;
    build_hash_tabs(&Ht, Tb, P.k);
    const uint64_t w = (uint64_t)P.w;
    const uint64_t stride = (uint64_t)GAP_CHUNK + w;
    uint64_t* H = scratch_h + (uint64_t)blockIdx.x * stride;
    uint64_t* Q = scratch_p + (uint64_t)blockIdx.x * stride;
    uint64_t* BM = scratch_bm + (uint64_t)blockIdx.x * (stride / 32 + 2);
    uint64_t* BI = scratch_bi + (uint64_t)blockIdx.x * (stride / 32 + 2);
    const uint64_t n_gaps = dev_count(reinterpret_cast<const uint64_t*>(d_n_gaps), n_gaps_max);
    for (uint64_t g = blockIdx.x; g < n_gaps; g += gridDim.x) {
        const uint64_t ja = gaps[g].ja, jb = gaps[g].jb;
        for (uint64_t ca = ja; ca <= jb; ca += GAP_CHUNK) {
            uint64_t cb = ca + GAP_CHUNK - 1 < jb ? ca + GAP_CHUNK - 1 : jb;
            uint64_t o0 = ca - (w - 1);
            uint64_t m = cb - o0 + 1;
            __syncthreads();
            // each thread fills a run of GAP_RUN consecutive ordinals: one select + one table hash, then rolling
            for (uint64_t e0 = (uint64_t)threadIdx.x * GAP_RUN; e0 < m; e0 += (uint64_t)blockDim.x * GAP_RUN) {
                uint64_t p = bitmap_select(V, vprefix, n_vblocks, o0 + e0);
                uint64_t f, r;
                kmer_hash64_tab(pk, p, P.k, &Ht, f, r);
                H[e0] = canon(f, r, P.canon_min);
                Q[e0] = p;
                const uint64_t e1 = e0 + GAP_RUN < m ? e0 + GAP_RUN : m;
                for (uint64_t e = e0 + 1; e < e1; e++) {
                    // next valid k-mer start after p
                    uint64_t q = p + 1;
                    uint32_t wv = V[q >> 5] >> (q & 31);
                    while (!wv) { q = ((q >> 5) + 1) << 5; wv = V[q >> 5]; }
                    q += __ffs(wv) - 1;
                    if (q == p + 1) {
                        const uint32_t o = pk_code(pk, p), in = pk_code(pk, p + P.k);
                        f = srol1(f) ^ Tb.seed_rolk[o] ^ Tb.seed[in];
                        r = sror1(r ^ Tb.seed_rolk[in ^ 2u] ^ Tb.seed[o ^ 2u]);
                    } else {
                        kmer_hash64_tab(pk, q, P.k, &Ht, f, r);
                    }
                    p = q;
                    H[e] = canon(f, r, P.canon_min);
                    Q[e] = p;
                }
            }
            __syncthreads();
            // block minima (32 elements, rightmost on ties), then one thread per window end: partial right block,
            // whole blocks right to left, partial left block; strict '<' while moving left keeps the rightmost minimum
            const uint64_t nb = (m + 31) >> 5;
            for (uint64_t b = threadIdx.x; b < nb; b += blockDim.x) {
                const uint64_t s0 = b << 5, s1 = s0 + 32 < m ? s0 + 32 : m;
                uint64_t best = H[s0], bi = s0;
                for (uint64_t x = s0 + 1; x < s1; x++) { uint64_t hx = H[x]; if (hx <= best) { best = hx; bi = x; } }
                BM[b] = best;
                BI[b] = bi;
            }
            __syncthreads();
            for (uint64_t je = w - 1 + threadIdx.x; je < m; je += blockDim.x) {
                const uint64_t lo = je - (w - 1);
                uint64_t best = H[je], bi = je;
                uint64_t x = je;
                while (x > lo && (x & 31) != 0) { x--; uint64_t hx = H[x]; if (hx < best) { best = hx; bi = x; } }   // down to a block boundary
                while (x >= lo + 32) { uint64_t bb = (x >> 5) - 1; if (BM[bb] < best) { best = BM[bb]; bi = BI[bb]; } x -= 32; }
                while (x > lo) { x--; uint64_t hx = H[x]; if (hx < best) { best = hx; bi = x; } }
                if (best != ~0ULL) {
                    uint64_t p = Q[bi];
                    atomicOr(&M[p >> 5], 1u << (p & 31));
                }
            }
        }
    }
}

// ---------------------------------------------------------------- final_eval: one thread per minimizer
__device__ __forceinline__ void emit_minimizer(uint64_t i, uint64_t p, uint64_t f, uint64_t r, const uint64_t* __restrict__ offsets,
                                               uint32_t n_contigs, const SketchParams& P, uint64_t* __restrict__ out_hash,
                                               uint64_t* __restrict__ min_hash, uint32_t* __restrict__ pos, uint32_t* __restrict__ contig,
                                               uint8_t* __restrict__ forward)
{
    const uint64_t h0 = canon(f, r, P.canon_min);
    const uint32_t c = contig_of(offsets, n_contigs, p);
    out_hash[i] = mix_out_hash(h0, P.k);
    min_hash[i] = h0;
    pos[i] = (uint32_t)(p - offsets[c]);
    contig[i] = c;
    forward[i] = f <= r;
}

__global__ void __launch_bounds__(256) final_eval_kernel(const uint64_t* __restrict__ mpos, const uint64_t* __restrict__ d_n, uint64_t n_max,
                                                          const uint32_t* __restrict__ pk,
                                                          const uint64_t* __restrict__ offsets, uint32_t n_contigs,
                                                          SketchParams P, SketchTables Tb,
                                                          uint64_t* __restrict__ out_hash, uint64_t* __restrict__ min_hash,
                                                          uint32_t* __restrict__ pos, uint32_t* __restrict__ contig,
                                                          uint8_t* __restrict__ forward)
{
    __shared__ HashTabs H;
    build_hash_tabs(&H, Tb, P.k);
    const uint64_t n_mx = dev_count(d_n, n_max);
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mx) return;
    const uint64_t p = mpos[i];
    uint64_t f, r;
    kmer_hash64_tab(pk, p, P.k, &H, f, r);
    emit_minimizer(i, p, f, r, offsets, n_contigs, P, out_hash, min_hash, pos, contig, forward);
}

// The same with the position-specific tables of the candidate stage (k / 4 <= HASHPOS_MAX_GROUPS), persistent grid: the
// tables are staged once per CTA and every thread loops over minimizers.  final_eval_kernel rebuilds its 4-base tables
// in every CTA (~300 instructions per thread for ONE minimizer per thread: measured 200 M warp instructions, ALU pipe
// 92 %, 0.25 ms per 6 M minimizers).
__global__ void __launch_bounds__(256) final_eval_pos_kernel(const uint64_t* __restrict__ mpos, const uint64_t* __restrict__ d_n, uint64_t n_max,
                                                              const uint32_t* __restrict__ pk,
                                                              const uint64_t* __restrict__ offsets, uint32_t n_contigs,
                                                              SketchParams P, SketchTables Tb, const uint64_t* __restrict__ PF,
                                                              const uint64_t* __restrict__ PR,
                                                              uint64_t* __restrict__ out_hash, uint64_t* __restrict__ min_hash,
                                                              uint32_t* __restrict__ pos, uint32_t* __restrict__ contig,
                                                              uint8_t* __restrict__ forward)
{
    extern __shared__ __align__(16) uint64_t hp[];
    const int G = P.k / 4;
    uint64_t* pf = hp;
    uint64_t* pr = hp + G * 256;
    __shared__ uint64_t s1[8];
    stage_pos_tables(Tb, P.k, PF, PR, pf, pr, s1);
    const uint64_t n_mx = dev_count(d_n, n_max);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_mx; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t p = mpos[i];
        uint64_t f, r;
        kmer_hash64_pos(pk, p, P.k, pf, pr, s1, f, r);
        emit_minimizer(i, p, f, r, offsets, n_contigs, P, out_hash, min_hash, pos, contig, forward);
    }
}

}  // namespace mxe
