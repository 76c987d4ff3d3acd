// p2p.cu -- steps 2-3, second generation: bucket partition by hash owner, exchanges as direct peer stores.
//
// Replaces (reference, Python): per-assembly uniqueness (bin/ntjoin_utils.py:182-192), the found-in-all intersection
// and ordered filtering (:155-162), the adjacent-pair edge dictionary with its support lists (:94-115) and the edge
// weights (:54-56) -- same results as filter.cu, different machine mapping:
//
//   * no global sort.  The 64-bit out_hash is uniformly mixed, so its top B bits cut the minimizers of all assemblies
//     into 2^B BUCKETS of ~200 records; a bucket is sorted and analysed by one CTA in shared memory (uniqueness per
//     assembly, found-in-all runs, vertex ranks), the kernels before and after it are single passes.
//   * one code path for 1..16 GPUs.  Bucket b belongs to rank (b * world) >> B (monotone in the hash, so the ranks'
//     vertex lists concatenate to the ascending single-GPU order).  Every rank holds the same SYMMETRIC workspace; the
//     producing kernels write straight into the consumer's copy over NVLink (plain stores into peer memory mapped
//     through CUDA IPC, a few atomics for the first-source tables), and a device-side barrier (one flag store per peer,
//     one spinning warp) separates the stages.  No collective library, no host round trip between the stages:
//
//       scatter    (source)  record {hash, asm | rank | local index} -> the OWNER of its bucket (one GPU: straight into
//                            the bucket's slot; several: appended to the owner's segment of this source, which the owner
//                            cuts into buckets)
//       --- barrier 1 (the record counts and the per-rank minimizer counts ride along)
//       buckets    (owner)   sort + run analysis in shared memory; the mark of every record goes back to its SOURCE,
//                            the bucket's survivor count to EVERY rank
//       --- barrier 2
//       adjacency  (source)  vertex ids from the bucket prefix, ordered survivors of own records, adjacent pairs;
//                            every sighting -> the OWNERS of its two vertices (successor / predecessor record)
//       --- barrier 3
//       edges      (owner)   successor / predecessor tables of the own vertices, support masks, edge ownership,
//                            first-source index: local loads only
//       finish     (owner)   world == 1: edges placed directly in the reference's formatted_edges order (a prefix sum
//                            over first edges; no sort).  world > 1: this rank's shard (flags of its own minimizers,
//                            vertices of its hash range, the edges whose source vertex it owns) with 64-bit order keys.
//
// Uniqueness is per ASSEMBLY, not per GPU (bin/ntjoin_utils.py:182-187): the owner of a bucket sees the full multiset.
// Repeated sequence puts thousands of copies of one hash into one bucket.  world > 1: the owner cuts the received
// records into EXACTLY sized bucket ranges (count, prefix sum, place), and a bucket larger than the shared memory of
// its CTA is reduced while it is loaded: at most two copies of every (hash, assembly) pair are kept -- two already
// prove "not unique", the dropped copies keep the mark 0 = "not unique, not kept" their slots were cleared to.
// world == 1 keeps fixed sub-slots (one pass); an overflow there raises an error flag and mxe_filter_and_edges falls
// back to the sort-based formulation of filter.cu by itself.
#include "engine.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace mxe {

constexpr int P2P_MAX_WORLD = 16;
constexpr int P2P_BK_MAX = 1024;          // most records a bucket CTA can hold in shared memory (P2PLayout::bk_max: 512 or 1024)
constexpr int P2P_BK_THREADS = 128;       // = digits of the in-bucket counting sort (7 bits)
constexpr int P2P_N_BARRIERS = 8;

struct PeerPtrs { char* base[P2P_MAX_WORLD]; };
struct PtrTab { const void* p[32]; };

struct P2PLayout {
    int world, rank, n_asm_max, B;
    int n_asm;                             // assemblies of the call in flight (set by mxe_p2p_scatter): stride of the vertex tables
    uint32_t n_buckets;                    // 2^B
    uint32_t fb[P2P_MAX_WORLD + 1];        // first bucket of every rank
    uint32_t nb_own_max;                   // most buckets one rank owns
    uint32_t cap_sub;                      // records per bucket sub-slot (world == 1; world > 1 places buckets at exact offsets)
    int bk_max;                            // records a bucket CTA holds in shared memory: 256, 512 or 1024
    uint64_t L_cap;                        // minimizers of one rank, all assemblies
    uint64_t nv_cap;                       // vertices of one owner
    // byte offsets inside every rank's workspace
    uint64_t off_flags, off_err, off_counts, off_rec_cnt, off_nkeep, off_rec, off_mk, off_tab, off_vgid, off_pred, off_cnt2, off_rec2, off_cnt1, off_seg1, bytes;
    uint64_t cap1;                         // minimizer records per (owner, source) segment (world > 1)
    uint64_t cap2;                         // sighting records per (owner, source) segment (world > 1)
};

struct P2PRecord { uint64_t key, tag; };   // tag = asm << 40 | source rank << 32 | local index

__host__ __device__ __forceinline__ uint32_t p2p_owner(uint32_t b, int world, int B) { return (uint32_t)(((uint64_t)b * (uint64_t)world) >> B); }

// Vertex tables of an owner, INTERLEAVED by assembly: the entries of one vertex in all assemblies share a sector, so the
// support test of a sighting (the successors of v and of x in every assembly) costs two sector reads instead of
// 2 * n_asm (the per-assembly planes [a][vertex] of the first version made these kernels the most expensive of steps 2-3:
// ~146 M random 4-byte accesses per step of configs[2]).
//   tab[vloc * n_asm + a]   = 1 + successor vertex of vloc in assembly a (0 = none); bit 31 = "this sighting created
//                              an edge" (ownership mark, p2p_edge_owner_kernel)
//   vgid[vloc * n_asm + a]  = creation index of the survivor (a, vloc): order keys of result shards (world > 1 only; on
//                              one GPU the first edge of a source is the one of the lowest assembly and nothing reads it)
//   pred[vloc * n_asm + a]  = 1 + predecessor vertex (world > 1, records mode)
// The successor table is kept on its own: 8 bytes per vertex at n_asm = 2, so the whole table of configs[2] (45 MB) stays
// in L2 between the kernel that fills it and the kernels that look things up in it.
__device__ __forceinline__ uint32_t* p2p_tab(const PeerPtrs& P, const P2PLayout& Y, int o, uint64_t vloc)
{
    return reinterpret_cast<uint32_t*>(P.base[o] + Y.off_tab) + vloc * (uint64_t)Y.n_asm;
}
__device__ __forceinline__ uint32_t* p2p_vgid(const PeerPtrs& P, const P2PLayout& Y, int o, uint64_t vloc)
{
    return reinterpret_cast<uint32_t*>(P.base[o] + Y.off_vgid) + vloc * (uint64_t)Y.n_asm;
}
__device__ __forceinline__ uint32_t* p2p_pred(const PeerPtrs& P, const P2PLayout& Y, int o, uint64_t vloc)
{
    return reinterpret_cast<uint32_t*>(P.base[o] + Y.off_pred) + vloc * (uint64_t)Y.n_asm;
}

__device__ __forceinline__ int p2p_slice_of(const LocalSlices& S, uint64_t l)
{
    int a = 0;
    while (a + 1 < S.n && l >= S.lofs[a + 1]) a++;
    return a;
}

// ---------------------------------------------------------------- device-side barrier
// signal: after everything this rank issued before it (stream order; the stage kernels have completed, their peer
// stores are performed), store the epoch into slot [bar][rank] of every peer.  wait: spin until all peers' epochs have
// arrived here.  The spin is bounded (~2 s): a lost peer sets the error flag instead of hanging the GPU.
__global__ void p2p_signal_kernel(const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, int bar, uint32_t epoch)
{
    __threadfence_system();
    if ((int)threadIdx.x < Y.world) {
        volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(P.base[threadIdx.x] + Y.off_flags) + bar * P2P_MAX_WORLD + Y.rank;
        *f = epoch;
    }
    __threadfence_system();
}

__global__ void p2p_wait_kernel(const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, int bar, uint32_t epoch)
{
    if ((int)threadIdx.x < Y.world) {
        volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(P.base[Y.rank] + Y.off_flags) + bar * P2P_MAX_WORLD + threadIdx.x;
        const long long t0 = clock64();
        while ((int32_t)(*f - epoch) < 0) {
            if (clock64() - t0 > 4000000000LL) { *reinterpret_cast<volatile uint32_t*>(P.base[Y.rank] + Y.off_err) = 0x100u + bar; break; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

// ---------------------------------------------------------------- stage 1: scatter to the bucket owners
__global__ void __launch_bounds__(256) p2p_scatter_kernel(PtrTab H, LocalSlices S, uint64_t L, const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, uint32_t* __restrict__ cursor)
{
    const uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const int a = p2p_slice_of(S, l);
    const uint64_t key = reinterpret_cast<const uint64_t*>(H.p[a])[l - S.lofs[a]];
    const uint32_t b = (uint32_t)(key >> (64 - Y.B));
    const uint32_t o = p2p_owner(b, Y.world, Y.B);
    const uint32_t slot = atomicAdd(&cursor[b], 1u);
    if (slot >= Y.cap_sub) { *reinterpret_cast<uint32_t*>(P.base[Y.rank] + Y.off_err) = 1u; return; }
    P2PRecord* dst = reinterpret_cast<P2PRecord*>(P.base[o] + Y.off_rec) + ((uint64_t)(b - Y.fb[o]) * Y.world + Y.rank) * Y.cap_sub + slot;
    *reinterpret_cast<ulonglong2*>(dst) = make_ulonglong2(key, ((uint64_t)a << 40) | ((uint64_t)Y.rank << 32) | l);
}

// per-(bucket, source) counts to the owners; this rank's minimizer counts to everybody
struct AsmCounts { uint64_t n[32]; int n_asm; };
__global__ void __launch_bounds__(256) p2p_push_counts_kernel(const uint32_t* __restrict__ cursor, AsmCounts C, const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < Y.n_buckets) {
        const uint32_t o = p2p_owner(b, Y.world, Y.B);
        const uint32_t c = cursor[b] < Y.cap_sub ? cursor[b] : Y.cap_sub;
        reinterpret_cast<uint32_t*>(P.base[o] + Y.off_rec_cnt)[(uint64_t)(b - Y.fb[o]) * Y.world + Y.rank] = c;
    }
    if (b < (uint32_t)(Y.world * C.n_asm)) {
        const int p = b / C.n_asm, a = b % C.n_asm;
        reinterpret_cast<uint64_t*>(P.base[p] + Y.off_counts)[Y.rank * Y.n_asm_max + a] = C.n[a];
    }
}

// ---------------------------------------------------------------- CTA-staged appends to peer segments (world > 1)
// A remote store is a NVLink packet: 4- or 16-byte stores to random addresses of a peer cost ~0.1 us each and, over a
// window of hundreds of MB, thrash the TLB of the peer mapping (measured: 45 ms per step at 55 M records capacity).
// So nothing is scattered remotely: every CTA groups the records of its threads by destination rank in shared memory,
// reserves one contiguous range per destination in that rank's segment [source] (cursor of the SOURCE: no remote
// atomics) and copies each range with consecutive threads on consecutive 8-byte words.
struct CtaAppendState { uint32_t cnt[P2P_MAX_WORLD], off[P2P_MAX_WORLD + 1], base[P2P_MAX_WORLD]; };

template <int WORDS, int SLOTS>
__device__ __forceinline__ void cta_append(CtaAppendState& sh, uint64_t* stage, const int (&dest)[SLOTS], const uint64_t (&rec)[SLOTS][WORDS],
                                           uint32_t* __restrict__ cur, const PeerPtrs& P, uint64_t off_seg, uint64_t cap, const P2PLayout& Y, uint32_t err_code)
{
    if (threadIdx.x < P2P_MAX_WORLD) sh.cnt[threadIdx.x] = 0u;
    __syncthreads();
    uint32_t local[SLOTS];
#pragma unroll
    for (int q = 0; q < SLOTS; q++) local[q] = dest[q] >= 0 ? atomicAdd(&sh.cnt[dest[q]], 1u) : 0u;
    __syncthreads();
    if ((int)threadIdx.x < Y.world) sh.base[threadIdx.x] = sh.cnt[threadIdx.x] ? atomicAdd(&cur[threadIdx.x], sh.cnt[threadIdx.x]) : 0u;
    if (threadIdx.x == 0) {
        uint32_t at = 0;
        for (int d = 0; d < Y.world; d++) { sh.off[d] = at; at += sh.cnt[d]; }
        sh.off[Y.world] = at;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < SLOTS; q++)
        if (dest[q] >= 0) {
            uint64_t* dstw = stage + (size_t)(sh.off[dest[q]] + local[q]) * WORDS;
#pragma unroll
            for (int w = 0; w < WORDS; w++) dstw[w] = rec[q][w];
        }
    __syncthreads();
    const uint32_t total = sh.off[Y.world] * WORDS;
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
        const uint32_t r = i / WORDS, w = i - r * WORDS;
        int d = 0;
        while (d + 1 < Y.world && r >= sh.off[d + 1]) d++;
        const uint64_t pos = (uint64_t)sh.base[d] + (r - sh.off[d]);
        if (pos < cap) reinterpret_cast<uint64_t*>(P.base[d] + off_seg)[((uint64_t)Y.rank * cap + pos) * WORDS + w] = stage[i];
        else *reinterpret_cast<uint32_t*>(P.base[Y.rank] + Y.off_err) = err_code;
    }
}

// stage 1 (world > 1): own minimizers -> segment [this rank] of the owner of their bucket
__global__ void __launch_bounds__(256) p2p_scatter_seg_kernel(PtrTab H, LocalSlices S, uint64_t L, const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, uint32_t* __restrict__ cur1)
{
    __shared__ CtaAppendState sh;
    __shared__ uint64_t stage[256 * 2];
    const uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int dest[1] = {-1};
    uint64_t rec[1][2] = {{0, 0}};
    if (l < L) {
        const int a = p2p_slice_of(S, l);
        const uint64_t key = reinterpret_cast<const uint64_t*>(H.p[a])[l - S.lofs[a]];
        dest[0] = (int)p2p_owner((uint32_t)(key >> (64 - Y.B)), Y.world, Y.B);
        rec[0][0] = key;
        rec[0][1] = ((uint64_t)a << 40) | ((uint64_t)Y.rank << 32) | l;
    }
    cta_append<2, 1>(sh, stage, dest, rec, cur1, P, Y.off_seg1, Y.cap1, Y, 4u);
}

// segment counts to the owners; this rank's minimizer counts to everybody
__global__ void p2p_push_cnt1_kernel(const uint32_t* __restrict__ cur1, AsmCounts C, const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y)
{
    if ((int)threadIdx.x < Y.world) {
        const uint32_t c = cur1[threadIdx.x] < Y.cap1 ? cur1[threadIdx.x] : (uint32_t)Y.cap1;
        reinterpret_cast<uint32_t*>(P.base[threadIdx.x] + Y.off_cnt1)[Y.rank] = c;
    }
    if ((int)threadIdx.x < Y.world * C.n_asm) {
        const int p = threadIdx.x / C.n_asm, a = threadIdx.x % C.n_asm;
        reinterpret_cast<uint64_t*>(P.base[p] + Y.off_counts)[Y.rank * Y.n_asm_max + a] = C.n[a];
    }
}

// owner: the received records of every source -> bucket ranges of EXACT size (local stores, local atomics): count per
// bucket, exclusive prefix, place.  No per-bucket capacity: a hash repeated 50,000 times makes one long bucket instead
// of an error (p2p_bucket_kernel reduces it while loading).
__device__ __forceinline__ bool p2p_seg_record(const PeerPtrs& P, const P2PLayout& Y, uint64_t idx, ulonglong2* r)
{
    const uint32_t seg = (uint32_t)(idx / Y.cap1);
    if (seg >= (uint32_t)Y.world) return false;
    const uint64_t i = idx - (uint64_t)seg * Y.cap1;
    const char* me = P.base[Y.rank];
    if (i >= reinterpret_cast<const uint32_t*>(me + Y.off_cnt1)[seg]) return false;
    *r = reinterpret_cast<const ulonglong2*>(me + Y.off_seg1)[idx];
    return true;
}

__global__ void __launch_bounds__(256) p2p_part_count_kernel(const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, uint32_t* __restrict__ bcount)
{
    ulonglong2 r;
    if (!p2p_seg_record(P, Y, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, &r)) return;
    atomicAdd(&bcount[(uint32_t)(r.x >> (64 - Y.B)) - Y.fb[Y.rank]], 1u);
}

// one CTA: bstart[bl] = exclusive prefix of the bucket counts (nb + 1 entries, in the workspace: the bucket and vertex
// kernels read it); the counts are replaced by the same values = placement cursors
__global__ void __launch_bounds__(1024) p2p_part_start_kernel(const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, uint32_t* __restrict__ bcount)
{
    __shared__ uint32_t sw[32];
    __shared__ uint32_t carry_s;
    uint32_t* bstart = reinterpret_cast<uint32_t*>(P.base[Y.rank] + Y.off_rec_cnt);
    const uint32_t nb = Y.fb[Y.rank + 1] - Y.fb[Y.rank];
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? bcount[i] : 0u;
        uint32_t x = v;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) sw[warp] = x;
        __syncthreads();
        uint32_t wb = 0;
        for (int wv = 0; wv < warp; wv++) wb += sw[wv];
        const uint32_t carry = carry_s;
        if (i < nb) { bstart[i] = carry + wb + x - v; bcount[i] = carry + wb + x - v; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wb + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) bstart[nb] = carry_s;
}

__global__ void __launch_bounds__(256) p2p_part_place_kernel(const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, uint32_t* __restrict__ bcursor)
{
    ulonglong2 r;
    if (!p2p_seg_record(P, Y, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, &r)) return;
    const uint32_t pos = atomicAdd(&bcursor[(uint32_t)(r.x >> (64 - Y.B)) - Y.fb[Y.rank]], 1u);
    reinterpret_cast<ulonglong2*>(P.base[Y.rank] + Y.off_rec)[pos] = r;      // pos < sum of the segment counts <= world * cap1
}

// the records of local bucket bl: one contiguous range.  world == 1: the bucket's fixed sub-slot (filled by
// p2p_scatter_kernel, count pushed by p2p_push_counts_kernel); world > 1: [bstart[bl], bstart[bl + 1]).
__device__ __forceinline__ P2PRecord* p2p_bucket_range(const PeerPtrs& P, const P2PLayout& Y, uint32_t bl, uint32_t* n)
{
    char* me = P.base[Y.rank];
    const uint32_t* tab = reinterpret_cast<const uint32_t*>(me + Y.off_rec_cnt);
    P2PRecord* rec = reinterpret_cast<P2PRecord*>(me + Y.off_rec);
    if (Y.world == 1) {
        *n = tab[bl] < Y.cap_sub ? tab[bl] : Y.cap_sub;
        return rec + (uint64_t)bl * Y.cap_sub;
    }
    *n = tab[bl + 1] - tab[bl];
    return rec + tab[bl];
}

// ---------------------------------------------------------------- stage 2: one CTA per owned bucket
__device__ __forceinline__ bool rec_greater(uint64_t ka, uint64_t ta, uint64_t kb, uint64_t tb) { return ka > kb || (ka == kb && ta > tb); }

constexpr uint64_t P2P_EMPTY = ~0ULL;      // free slot of the reduction table of an oversized bucket

template <int BKMAX>
__global__ void __launch_bounds__(P2P_BK_THREADS) p2p_bucket_kernel(const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, int n_asm)
{
    extern __shared__ uint64_t bk_smem[];
    uint64_t* skey = bk_smem;                          // sorted records
    uint64_t* stag = bk_smem + BKMAX;
    uint64_t* ukey = bk_smem + 2 * BKMAX;              // as loaded
    uint64_t* utag = bk_smem + 3 * BKMAX;
    uint32_t* srank = reinterpret_cast<uint32_t*>(ukey);   // (after the sort) head flag -> exclusive rank of the kept runs
    __shared__ uint32_t dstart[P2P_BK_THREADS + 1];
    __shared__ uint32_t dfill[P2P_BK_THREADS];
    __shared__ uint32_t swarp[P2P_BK_THREADS / 32];
    __shared__ uint32_t n_red_s, fail_s;
    const uint32_t bl = blockIdx.x;                                  // local bucket
    const uint32_t b = Y.fb[Y.rank] + bl;
    char* me = P.base[Y.rank];
    uint32_t n_in;
    P2PRecord* region = p2p_bucket_range(P, Y, bl, &n_in);
    dfill[threadIdx.x] = 0u;
    if (threadIdx.x == 0) { n_red_s = 0u; fail_s = 0u; }
    // Sort by (hash, tag): equal hashes end up grouped by assembly, then source rank, then position.  The hash is
    // uniformly mixed, so a counting sort on its next 7 bits leaves ~3 records per digit, finished by one thread per
    // digit with an insertion sort (a bitonic network moves every record log^2(n) / 2 times through shared memory:
    // measured 1.3 ms for 12 M records; this is four passes).
    const int dshift = 64 - Y.B - 7;
    uint32_t n = n_in;
    if (n_in <= (uint32_t)BKMAX) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_in; i += P2P_BK_THREADS) {
            const ulonglong2 r = *reinterpret_cast<const ulonglong2*>(region + i);
            ukey[i] = r.x;
            utag[i] = r.y;
            atomicAdd(&dfill[(uint32_t)(r.x >> dshift) & (P2P_BK_THREADS - 1)], 1u);
        }
    } else {
        // Oversized bucket = repeated sequence: thousands of copies of a few hashes beside the usual ~200 records.  Only
        // "once" or "more than once" per (hash, assembly) matters (bin/ntjoin_utils.py:182-187), so at most TWO copies
        // of every pair are loaded; the others are dropped here and keep mark 0 (not unique, not kept).  Table in the
        // space of the sorted arrays: skey = hash (open addressing), stag = two bits per assembly (n_asm <= 32).
        for (uint32_t i = threadIdx.x; i < (uint32_t)BKMAX; i += P2P_BK_THREADS) { skey[i] = P2P_EMPTY; stag[i] = 0ULL; }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_in; i += P2P_BK_THREADS) {
            const ulonglong2 r = *reinterpret_cast<const ulonglong2*>(region + i);
            if (r.x == P2P_EMPTY) { fail_s = 1u; continue; }        // cannot be told from a free slot (one hash value in 2^64)
            uint32_t sl = ((uint32_t)r.x ^ (uint32_t)(r.x >> 29)) & (BKMAX - 1);
            int probes = 0;
            for (; probes < BKMAX; probes++, sl = (sl + 1) & (BKMAX - 1)) {
                const unsigned long long cur = atomicCAS(reinterpret_cast<unsigned long long*>(&skey[sl]), (unsigned long long)P2P_EMPTY, (unsigned long long)r.x);
                if (cur == P2P_EMPTY || cur == r.x) break;
            }
            if (probes == BKMAX) { fail_s = 1u; continue; }         // more distinct hashes than the table holds
            const unsigned long long first = 1ULL << (2 * (int)((r.y >> 40) & 0x1F));
            unsigned long long old = atomicOr(reinterpret_cast<unsigned long long*>(&stag[sl]), first);
            bool keep = !(old & first);
            if (!keep) {
                old = atomicOr(reinterpret_cast<unsigned long long*>(&stag[sl]), first << 1);
                keep = !(old & (first << 1));
            }
            if (keep) {
                const uint32_t pos = atomicAdd(&n_red_s, 1u);
                if (pos < (uint32_t)BKMAX) {
                    ukey[pos] = r.x;
                    utag[pos] = r.y;
                    atomicAdd(&dfill[(uint32_t)(r.x >> dshift) & (P2P_BK_THREADS - 1)], 1u);
                }
            }
        }
        __syncthreads();
        n = n_red_s;
        if (fail_s) n = BKMAX + 1;
    }
    if (n > (uint32_t)BKMAX) {                                       // block-uniform
        if (threadIdx.x == 0) {
            *reinterpret_cast<uint32_t*>(me + Y.off_err) = 2u;
            for (int p = 0; p < Y.world; p++) reinterpret_cast<uint32_t*>(P.base[p] + Y.off_nkeep)[b] = 0u;
        }
        return;
    }
    __syncthreads();
    {   // exclusive scan of the digit counts (one per thread)
        const uint32_t v = dfill[threadIdx.x];
        uint32_t x = v;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) swarp[warp] = x;
        __syncthreads();
        uint32_t base = 0;
        for (int wv = 0; wv < warp; wv++) base += swarp[wv];
        dstart[threadIdx.x] = base + x - v;
        if (threadIdx.x == P2P_BK_THREADS - 1) dstart[P2P_BK_THREADS] = base + x;
        dfill[threadIdx.x] = 0u;
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < n; i += P2P_BK_THREADS) {
        const uint64_t k = ukey[i];
        const uint32_t d = (uint32_t)(k >> dshift) & (P2P_BK_THREADS - 1);
        const uint32_t pos = dstart[d] + atomicAdd(&dfill[d], 1u);
        skey[pos] = k;
        stag[pos] = utag[i];
    }
    __syncthreads();
    {
        const uint32_t s0 = dstart[threadIdx.x], s1 = dstart[threadIdx.x + 1];
        for (uint32_t i = s0 + 1; i < s1; i++) {
            const uint64_t kk = skey[i], tt = stag[i];
            uint32_t j = i;
            while (j > s0 && rec_greater(skey[j - 1], stag[j - 1], kk, tt)) { skey[j] = skey[j - 1]; stag[j] = stag[j - 1]; j--; }
            skey[j] = kk; stag[j] = tt;
        }
    }
    __syncthreads();
    uint32_t Pn = n;
    // run analysis (as mark_kernel): unique inside its assembly; run = exactly one element of every assembly
    uint32_t marks[BKMAX / P2P_BK_THREADS];
#pragma unroll
    for (int q = 0; q < BKMAX / P2P_BK_THREADS; q++) {
        const uint32_t i = threadIdx.x + q * P2P_BK_THREADS;
        uint32_t m = 0, head = 0;
        if (i < n) {
            const uint64_t K = skey[i];
            const int a = (int)((stag[i] >> 40) & 0xFF);
            const bool prev_same = i > 0 && skey[i - 1] == K && (int)((stag[i - 1] >> 40) & 0xFF) == a;
            const bool next_same = i + 1 < n && skey[i + 1] == K && (int)((stag[i + 1] >> 40) & 0xFF) == a;
            const uint32_t uniq = !(prev_same || next_same);
            bool in_all = false;
            if ((uint32_t)a <= i && i - a + n_asm <= n) {
                const uint32_t s0 = i - a;
                in_all = (s0 == 0 || skey[s0 - 1] != K) && (s0 + n_asm == n || skey[s0 + n_asm] != K);
                for (int j = 0; in_all && j < n_asm; j++) in_all = skey[s0 + j] == K && (int)((stag[s0 + j] >> 40) & 0xFF) == j;
            }
            m = (uniq << 31) | (in_all ? 1u : 0u);
            head = in_all && a == 0;
        }
        marks[q] = m;
        if (i < BKMAX) srank[i] = head;
    }
    __syncthreads();
    // exclusive scan of the head flags over [0, Pn): thread t owns the contiguous chunk [t * per, (t + 1) * per)
    {
        const uint32_t per = (Pn + P2P_BK_THREADS - 1) / P2P_BK_THREADS;
        const uint32_t i0 = threadIdx.x * per;
        uint32_t sum = 0;
        for (uint32_t i = i0; i < i0 + per && i < n; i++) sum += srank[i];
        uint32_t x = sum;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) swarp[warp] = x;
        __syncthreads();
        uint32_t base = 0;
        for (int wv = 0; wv < warp; wv++) base += swarp[wv];
        uint32_t run = base + x - sum;
        for (uint32_t i = i0; i < i0 + per && i < n; i++) { const uint32_t h = srank[i]; srank[i] = run; run += h; }
        __syncthreads();
    }
    uint32_t nk = 0;
    for (int wv = 0; wv < P2P_BK_THREADS / 32; wv++) nk += swarp[wv];
    // marks back to the sources; the kept hashes (ascending) stay at the start of the bucket's own region
    uint64_t* kept = reinterpret_cast<uint64_t*>(region);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < BKMAX / P2P_BK_THREADS; q++) {
        const uint32_t i = threadIdx.x + q * P2P_BK_THREADS;
        if (i >= n) continue;
        const uint64_t t = stag[i];
        const int a = (int)((t >> 40) & 0xFF);
        const uint32_t src = (uint32_t)((t >> 32) & 0xFF);
        const uint32_t keep = marks[q] & 1u;
        const uint32_t rk = keep ? srank[i - a] : 0u;
        reinterpret_cast<uint32_t*>(P.base[src] + Y.off_mk)[(uint32_t)t] = (marks[q] & 0x80000000u) | (keep ? rk + 1u : 0u);
        if (keep && a == 0) kept[rk] = skey[i];
    }
    if (threadIdx.x == 0)
        for (int p = 0; p < Y.world; p++) reinterpret_cast<uint32_t*>(P.base[p] + Y.off_nkeep)[b] = nk;
}

// ---------------------------------------------------------------- stage 3: home rank
// exclusive prefix of the per-bucket survivor counts (all ranks' buckets) -> global vertex id base of every bucket;
// global index of this rank's first minimizer of every assembly
struct HomeTabs {
    uint32_t* vbase;        // n_buckets + 1
    uint32_t* vown;         // world + 1: first vertex id of every owner
    uint64_t* goff;         // n_asm: global index of this rank's first minimizer of assembly a  (+ [n_asm] = N)
};

__global__ void __launch_bounds__(1024) p2p_vbase_kernel(const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, int n_asm, HomeTabs T)
{
    __shared__ uint32_t sw[32];
    __shared__ uint32_t carry_s;
    const uint32_t* nkeep = reinterpret_cast<const uint32_t*>(P.base[Y.rank] + Y.off_nkeep);
    // every WARP owns a contiguous chunk of buckets and walks it 32 at a time with coalesced loads, twice: chunk totals
    // first, one block-wide scan over the 32 totals, then the prefixes (a warp scan per 32 buckets, the carry in a
    // register).  Two barriers in all; the first version took three per 1024 buckets (76 us for 2^16 buckets on one CTA).
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = ((Y.n_buckets + 31u) / 32u + 31u) & ~31u;          // buckets per warp, a multiple of 32
    const uint32_t c0 = (uint32_t)warp * chunk;
    const uint32_t c1 = c0 + chunk < Y.n_buckets ? c0 + chunk : Y.n_buckets;
    uint32_t sum = 0;
    for (uint32_t i = c0 + lane; i < c1; i += 32) sum += nkeep[i];
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) sw[warp] = sum;
    __syncthreads();
    uint32_t run = 0;
    for (int wv = 0; wv < warp; wv++) run += sw[wv];
    for (uint32_t i0 = c0; i0 < c1; i0 += 32) {                              // warp-uniform bounds
        const uint32_t i = i0 + lane;
        const uint32_t v = i < c1 ? nkeep[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (i < c1) T.vbase[i] = run + x - v;
        run += __shfl_sync(0xffffffffu, x, 31);
    }
    if (threadIdx.x == 1023) carry_s = run;                                  // the last warp ends at the grand total
    __syncthreads();
    if (threadIdx.x == 0) T.vbase[Y.n_buckets] = carry_s;
    __syncthreads();
    if ((int)threadIdx.x <= Y.world) T.vown[threadIdx.x] = T.vbase[Y.fb[threadIdx.x]];
    if ((int)threadIdx.x <= n_asm) {
        // global index space: assemblies in order, inside an assembly the ranks in order
        const uint64_t* cnt = reinterpret_cast<const uint64_t*>(P.base[Y.rank] + Y.off_counts);
        const int t = (int)threadIdx.x;
        uint64_t g = 0;
        for (int a = 0; a < t && a < n_asm; a++)
            for (int r = 0; r < Y.world; r++) g += cnt[r * Y.n_asm_max + a];
        if (t < n_asm)
            for (int r = 0; r < Y.rank; r++) g += cnt[r * Y.n_asm_max + t];
        T.goff[t] = g;
    }
}

// the owner's vertices (ascending hash): bucket by bucket from the kept lists, one warp per bucket (~90 hashes)
__global__ void __launch_bounds__(128) p2p_vertices_kernel(const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, HomeTabs T, uint64_t* __restrict__ vertices)
{
    const uint32_t bl = blockIdx.x * 4u + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
    if (bl >= Y.fb[Y.rank + 1] - Y.fb[Y.rank]) return;
    const uint32_t b = Y.fb[Y.rank] + bl;
    const uint32_t v0 = T.vbase[b] - T.vown[Y.rank], nk = T.vbase[b + 1] - T.vbase[b];
    uint32_t n_in;
    const uint64_t* kept = reinterpret_cast<const uint64_t*>(p2p_bucket_range(P, Y, bl, &n_in));
    for (uint32_t i = lane; i < nk; i += 32) vertices[v0 + i] = kept[i];
}

__global__ void __launch_bounds__(256) p2p_flags_kernel(const uint32_t* __restrict__ mk, uint64_t L, uint32_t* __restrict__ kflag,
                                                         uint8_t* __restrict__ luniq, uint8_t* __restrict__ lkeep)
{
    const uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const uint32_t m = mk[l];
    const uint32_t k = (m & 0x7FFFFFFFu) != 0;
    kflag[l] = k;
    luniq[l] = (uint8_t)(m >> 31);
    lkeep[l] = (uint8_t)k;
}

// ordered survivors: global vertex id, global (creation) index, local index
__global__ void __launch_bounds__(256) p2p_compact_kernel(const uint32_t* __restrict__ mk, PtrTab H, LocalSlices S, uint64_t L,
                                                           const uint64_t* __restrict__ kprefix, const __grid_constant__ P2PLayout Y, HomeTabs T,
                                                           uint32_t* __restrict__ cvid, uint32_t* __restrict__ cg, uint32_t* __restrict__ cloc)
{
    const uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const uint32_t k = mk[l] & 0x7FFFFFFFu;
    if (!k) return;
    const int a = p2p_slice_of(S, l);
    const uint64_t key = reinterpret_cast<const uint64_t*>(H.p[a])[l - S.lofs[a]];
    const uint64_t j = kprefix[l];
    cvid[j] = T.vbase[(uint32_t)(key >> (64 - Y.B))] + k - 1u;
    cg[j] = (uint32_t)(T.goff[a] + (l - S.lofs[a]));
    cloc[j] = (uint32_t)l;
}

__device__ __forceinline__ int p2p_vowner(const uint32_t* vown, int world, uint32_t v)
{
    int o = 0;
    while (o + 1 < world && v >= vown[o + 1]) o++;
    return o;
}

// adjacent survivors of the same record and assembly: sighting flag; the table entry (a, v) at the owner of v.  Every
// vertex has exactly one survivor in every assembly, so every entry of every vertex is written here: the tables need
// no clearing in this mode.
__global__ void __launch_bounds__(256) p2p_succ_kernel(const uint32_t* __restrict__ cvid, const uint32_t* __restrict__ cg,
                                                        const uint32_t* __restrict__ cloc,
                                                        const uint64_t* __restrict__ kprefix, uint64_t L, LocalSlices S, PtrTab Ctg,
                                                        const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, HomeTabs T, uint32_t* __restrict__ eflag)
{
    const uint64_t n_keep = kprefix[L];
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_keep) return;
    const uint32_t l1 = cloc[j];
    const int a = p2p_slice_of(S, l1);
    const uint32_t v = cvid[j];
    uint32_t f = 0, x = 0;
    if (j + 1 < n_keep) {
        const uint32_t l2 = cloc[j + 1];
        if (a == p2p_slice_of(S, l2)) {
            const uint32_t* ctg = reinterpret_cast<const uint32_t*>(Ctg.p[a]);
            f = ctg[l1 - S.lofs[a]] == ctg[l2 - S.lofs[a]];
            x = cvid[j + 1];
        }
    }
    const int o = p2p_vowner(T.vown, Y.world, v);
    p2p_tab(P, Y, o, v - T.vown[o])[a] = f ? x + 1u : 0u;
    if (Y.world > 1) p2p_vgid(P, Y, o, v - T.vown[o])[a] = cg[j];
    eflag[j] = f;
}

// Edge de-duplication without sorting (as filter.cu): a surviving hash occurs once per assembly, so a vertex has at
// most one successor per assembly; assembly b supports {v, x} iff succ_b[v] == x or succ_b[x] == v; the sighting in
// the first supporting assembly owns the edge (first-seen orientation, bin/ntjoin_utils.py:101-108).
__global__ void __launch_bounds__(256) p2p_edge_owner_kernel(const uint32_t* __restrict__ cvid, const uint32_t* __restrict__ cg,
                                                              const uint32_t* __restrict__ cloc, const uint32_t* __restrict__ eflag,
                                                              const uint64_t* __restrict__ kprefix, uint64_t L, LocalSlices S, int n_asm,
                                                              const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, HomeTabs T, uint32_t* __restrict__ own, uint32_t* __restrict__ mask_out)
{
    const uint64_t n_keep = kprefix[L];
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= L) return;
    if (j >= n_keep) { own[j] = 0u; return; }        // the scan over own[] runs over L entries
    uint32_t is_owner = 0;
    if (eflag[j]) {
        const int a = p2p_slice_of(S, cloc[j]);
        const uint32_t v = cvid[j], x = cvid[j + 1];
        const int ov = p2p_vowner(T.vown, Y.world, v), ox = p2p_vowner(T.vown, Y.world, x);
        uint32_t* tv = p2p_tab(P, Y, ov, v - T.vown[ov]);
        const uint32_t* tx = p2p_tab(P, Y, ox, x - T.vown[ox]);
        uint32_t mask = 0;
        for (int b = 0; b < n_asm; b++)       // bit 31 of an entry is its ownership mark (below): compare the low 31 bits
            if ((tv[b] & 0x7FFFFFFFu) == x + 1u || (tx[b] & 0x7FFFFFFFu) == v + 1u) mask |= 1u << b;
        is_owner = (__ffs(mask) - 1) == a;
        mask_out[j] = mask;
        // "vertex v is the source of an edge created in assembly a": entry (a, v) has this sighting as its only writer,
        // so a plain store marks it (no atomics; readers of the successor ignore the bit)
        if (is_owner) tv[a] = (x + 1u) | 0x80000000u;
    }
    own[j] = is_owner;
}

// in which assemblies vertex v (local index at its owner o) is the source of a created edge (a vertex occurs once per
// assembly; the edges of one source are created in assembly order, so the FIRST edge of a source is the one of the lowest
// assembly in the mask)
__device__ __forceinline__ uint32_t p2p_source_mask(const PeerPtrs& P, const P2PLayout& Y, int o, uint32_t vloc, int n_asm)
{
    const uint32_t* tv = p2p_tab(P, Y, o, vloc);
    uint32_t smask = 0;
    for (int b = 0; b < n_asm; b++) smask |= (tv[b] >> 31) << b;
    return smask;
}
// ... and the creation index of its first one (the order key of its block of edges)
__device__ __forceinline__ uint32_t p2p_source_info(const PeerPtrs& P, const P2PLayout& Y, int o, uint32_t vloc, int n_asm, uint32_t* first_gid)
{
    const uint32_t smask = p2p_source_mask(P, Y, o, vloc, n_asm);
    *first_gid = smask ? p2p_vgid(P, Y, o, vloc)[__ffs(smask) - 1] : 0xFFFFFFFFu;
    return smask;
}

// formatted_edges order (bin/ntjoin_utils.py:115): sources in order of their first edge, the edges of a source in creation
// order (= assembly order: a source owns at most one edge per assembly).  Owned edges are compacted in creation order
// (uprefix); the first edge of a source carries the number of edges of that source, a prefix sum over them gives every
// source's block -> each edge is PLACED, not sorted.  (world == 1)
// The source mask is read from the vertex table ONCE per edge (here) and kept by sighting (smask_j); the position of a
// first edge is its own prefix entry, so only the later edges of a source that owns several go through vstart[] --
// p2p_first_start_kernel and p2p_edge_emit_kernel are sequential passes for everything else.
__global__ void __launch_bounds__(256) p2p_first_count_kernel(const uint32_t* __restrict__ cvid, const uint32_t* __restrict__ cloc,
                                                               const uint32_t* __restrict__ own, const uint64_t* __restrict__ uprefix,
                                                               const uint64_t* __restrict__ kprefix, uint64_t L, LocalSlices S, const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, int n_asm,
                                                               uint32_t* __restrict__ fcount, uint32_t* __restrict__ smask_j)
{
    const uint64_t n_keep = kprefix[L];
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_keep || !own[j]) return;
    const int a = p2p_slice_of(S, cloc[j]);
    const uint32_t smask = p2p_source_mask(P, Y, 0, cvid[j], n_asm);        // bit a is set: this sighting owns an edge
    smask_j[j] = smask;
    fcount[uprefix[j]] = (__ffs(smask) - 1) == a ? (uint32_t)__popc(smask) : 0u;
}

__global__ void __launch_bounds__(256) p2p_first_start_kernel(const uint32_t* __restrict__ cvid, const uint32_t* __restrict__ cloc,
                                                               const uint32_t* __restrict__ own, const uint32_t* __restrict__ smask_j,
                                                               const uint64_t* __restrict__ uprefix, const uint64_t* __restrict__ fprefix,
                                                               const uint64_t* __restrict__ kprefix, uint64_t L, LocalSlices S,
                                                               uint32_t* __restrict__ vstart)
{
    const uint64_t n_keep = kprefix[L];
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_keep || !own[j]) return;
    const uint32_t smask = smask_j[j];
    if (__popc(smask) > 1 && (__ffs(smask) - 1) == p2p_slice_of(S, cloc[j])) vstart[cvid[j]] = (uint32_t)fprefix[uprefix[j]];
}

struct EdgeOut { uint64_t *eu, *ev, *ekey; uint32_t* emask; double* ew; };

__global__ void __launch_bounds__(256) p2p_edge_emit_kernel(const uint32_t* __restrict__ cvid, const uint32_t* __restrict__ cg,
                                                             const uint32_t* __restrict__ cloc, const uint32_t* __restrict__ own,
                                                             const uint32_t* __restrict__ mask_in, const uint32_t* __restrict__ smask_j,
                                                             const uint64_t* __restrict__ uprefix, const uint64_t* __restrict__ fprefix,
                                                             const uint64_t* __restrict__ kprefix, uint64_t L, LocalSlices S, PtrTab H, AsmOffsets A,
                                                             const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, HomeTabs T, const uint32_t* __restrict__ vstart, EdgeOut E)
{
    const uint64_t n_keep = kprefix[L];
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_keep || !own[j]) return;
    const uint32_t l1 = cloc[j], l2 = cloc[j + 1];
    const int a = p2p_slice_of(S, l1);
    const uint64_t* hs = reinterpret_cast<const uint64_t*>(H.p[a]);
    const uint32_t v = cvid[j];
    uint64_t o;
    if (vstart) {       // single GPU: final position (first edge of its source: the start of the source's block)
        const uint32_t smask = smask_j[j];
        o = (__ffs(smask) - 1) == a ? fprefix[uprefix[j]] : (uint64_t)vstart[v] + (uint32_t)__popc(smask & ((1u << a) - 1u));
    } else {            // shard: creation order + global order key (first creation index of the source, creation index)
        const int ov = p2p_vowner(T.vown, Y.world, v);
        uint32_t smin;
        p2p_source_info(P, Y, ov, v - T.vown[ov], A.n, &smin);
        o = uprefix[j];
        E.ekey[o] = ((uint64_t)smin << 32) | (uint64_t)cg[j];
    }
    const uint32_t mask = mask_in[j];
    E.eu[o] = hs[l1 - S.lofs[a]];
    E.ev[o] = hs[l2 - S.lofs[a]];
    E.emask[o] = mask;
    double wsum = 0.0;   // Python: sum() starts at int 0 and adds in support-list (= assembly) order
    for (int b = 0; b < A.n; b++)
        if (mask & (1u << b)) wsum += A.weight[b];
    E.ew[o] = wsum;
}

// ---------------------------------------------------------------- world > 1: sightings travel to the vertex owners
// Two 24-byte records per sighting (a: v -> x, creation index g), written into segment [source rank] of the owner:
//   SUCC -> owner(v): { hash(x), local(v) << 32 | x, g << 8 | a << 1 | 0 }
//   PRED -> owner(x): { hash(v), local(x) << 32 | v,          a << 1 | 1 }
// so that the owner of a vertex holds its successor AND predecessor in every assembly and decides support masks, edge
// ownership, first-source index and order keys with local loads only (a remote 4-byte load costs a NVLink round trip;
// the stores are fire-and-forget).  Slots are reserved per CTA: one shared-memory count per destination, one global
// atomic per (CTA, destination).
__global__ void __launch_bounds__(256) p2p_sight_kernel(const uint32_t* __restrict__ cvid, const uint32_t* __restrict__ cg,
                                                         const uint32_t* __restrict__ cloc, const uint64_t* __restrict__ kprefix, uint64_t L,
                                                         LocalSlices S, PtrTab H, PtrTab Ctg, const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, HomeTabs T,
                                                         uint32_t* __restrict__ cur2)
{
    __shared__ CtaAppendState sh;
    __shared__ uint64_t stage[256 * 2 * 3];
    const uint64_t n_keep = kprefix[L];
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int dest[2] = {-1, -1};
    uint64_t rec[2][3] = {{0, 0, 0}, {0, 0, 0}};
    if (j + 1 < n_keep) {
        const uint32_t l1 = cloc[j], l2 = cloc[j + 1];
        const int a = p2p_slice_of(S, l1);
        if (a == p2p_slice_of(S, l2)) {
            const uint32_t* ctg = reinterpret_cast<const uint32_t*>(Ctg.p[a]);
            if (ctg[l1 - S.lofs[a]] == ctg[l2 - S.lofs[a]]) {
                const uint64_t* hs = reinterpret_cast<const uint64_t*>(H.p[a]);
                const uint32_t v = cvid[j], x = cvid[j + 1];
                const int ov = p2p_vowner(T.vown, Y.world, v), ox = p2p_vowner(T.vown, Y.world, x);
                dest[0] = ov; dest[1] = ox;
                rec[0][0] = hs[l2 - S.lofs[a]]; rec[0][1] = ((uint64_t)(v - T.vown[ov]) << 32) | x; rec[0][2] = ((uint64_t)cg[j] << 8) | ((uint64_t)a << 1);
                rec[1][0] = hs[l1 - S.lofs[a]]; rec[1][1] = ((uint64_t)(x - T.vown[ox]) << 32) | v; rec[1][2] = ((uint64_t)a << 1) | 1ULL;
            }
        }
    }
    cta_append<3, 2>(sh, stage, dest, rec, cur2, P, Y.off_rec2, Y.cap2, Y, 3u);
}

__global__ void p2p_push_cnt2_kernel(const uint32_t* __restrict__ cur2, const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y)
{
    if ((int)threadIdx.x < Y.world) {
        const uint32_t c = cur2[threadIdx.x] < Y.cap2 ? cur2[threadIdx.x] : (uint32_t)Y.cap2;
        reinterpret_cast<uint32_t*>(P.base[threadIdx.x] + Y.off_cnt2)[Y.rank] = c;
    }
}

// owner side.  Record index space: segment s (source rank) x cap2; entries at or beyond the segment's count are skipped.
__device__ __forceinline__ const uint64_t* p2p_rec2(const PeerPtrs& P, const P2PLayout& Y, uint64_t idx, bool* valid)
{
    const uint32_t seg = (uint32_t)(idx / Y.cap2);
    const uint64_t i = idx - (uint64_t)seg * Y.cap2;
    *valid = seg < (uint32_t)Y.world && i < reinterpret_cast<const uint32_t*>(P.base[Y.rank] + Y.off_cnt2)[seg];
    return reinterpret_cast<const uint64_t*>(P.base[Y.rank] + Y.off_rec2) + idx * 3;
}

// successor (+ creation index) / predecessor of every own vertex in every assembly (entries 1 + vertex id, 0 = none)
__global__ void __launch_bounds__(256) p2p_table_kernel(const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok;
    const uint64_t* r = p2p_rec2(P, Y, idx, &ok);
    if (!ok) return;
    const uint64_t w1 = r[1], w2 = r[2];
    const uint64_t a = (w2 >> 1) & 0x7F, vloc = w1 >> 32;
    const uint32_t other = (uint32_t)w1 + 1u;
    if (w2 & 1ULL) p2p_pred(P, Y, Y.rank, vloc)[a] = other;
    else { p2p_tab(P, Y, Y.rank, vloc)[a] = other; p2p_vgid(P, Y, Y.rank, vloc)[a] = (uint32_t)(w2 >> 8); }
}

__global__ void __launch_bounds__(256) p2p_rec_owner_kernel(const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, int n_asm, uint32_t* __restrict__ own, uint32_t* __restrict__ mask_out)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)Y.world * Y.cap2) return;
    bool ok;
    const uint64_t* r = p2p_rec2(P, Y, idx, &ok);
    uint32_t is_owner = 0;
    if (ok && !(r[2] & 1ULL)) {
        const uint64_t w1 = r[1], w2 = r[2];
        const int a = (int)((w2 >> 1) & 0x7F);
        const uint64_t vloc = w1 >> 32;
        const uint32_t x1 = (uint32_t)w1 + 1u;
        uint32_t* tv = p2p_tab(P, Y, Y.rank, vloc);
        const uint32_t* pv = p2p_pred(P, Y, Y.rank, vloc);
        uint32_t mask = 0;
        for (int b = 0; b < n_asm; b++)
            if ((tv[b] & 0x7FFFFFFFu) == x1 || pv[b] == x1) mask |= 1u << b;
        is_owner = (__ffs(mask) - 1) == a;
        mask_out[idx] = mask;
        if (is_owner) tv[a] = x1 | 0x80000000u;      // ownership mark (see p2p_edge_owner_kernel)
    }
    own[idx] = is_owner;
}

__global__ void __launch_bounds__(256) p2p_rec_emit_kernel(const __grid_constant__ PeerPtrs P, const __grid_constant__ P2PLayout Y, AsmOffsets A, const uint32_t* __restrict__ own,
                                                            const uint32_t* __restrict__ mask_in, const uint64_t* __restrict__ uprefix,
                                                            const uint64_t* __restrict__ vertices, EdgeOut E)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)Y.world * Y.cap2 || !own[idx]) return;
    const uint64_t* r = reinterpret_cast<const uint64_t*>(P.base[Y.rank] + Y.off_rec2) + idx * 3;
    const uint32_t vloc = (uint32_t)(r[1] >> 32);
    uint32_t smin;
    p2p_source_info(P, Y, Y.rank, vloc, A.n, &smin);
    const uint64_t o = uprefix[idx];
    const uint32_t mask = mask_in[idx];
    E.eu[o] = vertices[vloc];
    E.ev[o] = r[0];
    E.emask[o] = mask;
    E.ekey[o] = ((uint64_t)smin << 32) | (r[2] >> 8);
    double wsum = 0.0;
    for (int b = 0; b < A.n; b++)
        if (mask & (1u << b)) wsum += A.weight[b];
    E.ew[o] = wsum;
}

}  // namespace mxe

using namespace mxe;

// ====================================================================================================
// host side
// ====================================================================================================
struct mxe_p2p {
    mxe_engine* eng = nullptr;
    P2PLayout Y;
    PeerPtrs P;
    char* ws = nullptr;                    // this rank's symmetric workspace (cudaMalloc: exportable through CUDA IPC)
    std::vector<void*> opened;             // peer mappings opened through IPC
    uint32_t epoch = 0;
    // per-call local workspace (grow-only)
    char* lws = nullptr; size_t lws_bytes = 0, lws_used = 0;
    // state of the call in flight
    int n_asm = 0;
    uint64_t L = 0;
    LocalSlices S;
    PtrTab H, Ctg;
    AsmOffsets A;
    uint32_t *cursor = nullptr, *kflag = nullptr, *cvid = nullptr, *cg = nullptr, *cloc = nullptr, *eflag = nullptr, *own = nullptr, *emask_j = nullptr;
    uint64_t *kprefix = nullptr, *uprefix = nullptr;
    uint32_t* cur2 = nullptr;
    uint32_t* cur1 = nullptr;
    bool records = true;                   // world > 1: sightings travel to the owners of their vertices as records (CTA-staged, coalesced);
                                           // false (MXE_P2P_RECORDS=0) = successor tables at the owners written / read with fine-grained peer accesses
    uint8_t *luniq = nullptr, *lkeep = nullptr;
    uint64_t* vertices = nullptr;
    HomeTabs T;
    void* lalloc(size_t bytes)
    {
        bytes = (bytes + 255) & ~(size_t)255;
        if (lws_used + bytes > lws_bytes) return nullptr;
        void* p = lws + lws_used;
        lws_used += bytes;
        return p;
    }
};

namespace mxe {
void p2p_release(mxe_p2p* X, bool engine_is_alive)
{
    if (engine_is_alive) {
        cudaSetDevice(X->eng->device);
        cudaStreamSynchronize(X->eng->stream);
        for (void* p : X->opened) cudaIpcCloseMemHandle(p);
        if (X->ws) cudaFree(X->ws);
        if (X->lws) cudaFree(X->lws);
        if (X->luniq) cudaFree(X->luniq);
        if (X->lkeep) cudaFree(X->lkeep);
    }
    delete X;
}
}

static inline unsigned p2p_grid(uint64_t n) { return (unsigned)((n + 255) / 256); }

static int p2p_barrier_signal(mxe_p2p* X, int bar)
{
    if (X->Y.world == 1) return MXE_OK;
    MXE_LAUNCH(X->eng, p2p_signal_kernel, 1, 32, 0, X->P, X->Y, bar, X->epoch);
    return MXE_OK;
}
static int p2p_barrier_wait(mxe_p2p* X, int bar)
{
    if (X->Y.world == 1) return MXE_OK;
    static const char* const names[P2P_N_BARRIERS] = {"p2p_wait0", "p2p_wait1", "p2p_wait2", "p2p_wait3", "p2p_wait4", "p2p_wait5", "p2p_wait6", "p2p_wait7"};
    Span sp(X->eng, names[bar]);        // time spent waiting for the slowest peer (timing option only)
    MXE_LAUNCH(X->eng, p2p_wait_kernel, 1, 32, 0, X->P, X->Y, bar, X->epoch);
    return MXE_OK;
}

extern "C" {

int mxe_p2p_create(mxe_t* e, int rank, int world, uint64_t cap_total, int n_asm_max, mxe_p2p_t** out)
{
    if (!e || !out || world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world || n_asm_max < 1 || n_asm_max > 32) {
        set_error("bad rank/world/n_asm_max (world <= %d)", P2P_MAX_WORLD);
        return MXE_ERR_ARG;
    }
    if (cap_total < 1024) cap_total = 1024;
    if (cap_total >= (1ULL << 32)) { set_error("capacity beyond 2^32 minimizers"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(e->device));
    mxe_p2p* X = new mxe_p2p();
    X->eng = e;
    if (const char* sv = getenv("MXE_P2P_RECORDS")) X->records = atoi(sv) != 0;
    P2PLayout& Y = X->Y;
    memset(&Y, 0, sizeof(Y));
    Y.world = world; Y.rank = rank; Y.n_asm_max = n_asm_max;
    int B = 6;
    uint64_t per_bucket = 400;                                    // <= 400 records per bucket on average at full capacity
    if (const char* sv = getenv("MXE_P2P_BUCKET_AVG")) { const long v = atol(sv); if (v >= 32 && v <= P2P_BK_MAX / 2) per_bucket = (uint64_t)v; }
    while (B < 20 && (cap_total >> B) > per_bucket) B++;
    Y.B = B;
    Y.n_buckets = 1u << B;
    for (int r = 0; r <= world; r++) Y.fb[r] = (uint32_t)((((uint64_t)r << B) + world - 1) / world);
    Y.nb_own_max = 0;
    for (int r = 0; r < world; r++) Y.nb_own_max = std::max(Y.nb_own_max, Y.fb[r + 1] - Y.fb[r]);
    const double per_sub = (double)cap_total / ((double)Y.n_buckets * world);
    Y.cap_sub = (uint32_t)(per_sub * 1.5 + 6.0 * std::sqrt(per_sub + 1.0) + 32.0);
    Y.L_cap = (uint64_t)((double)cap_total / world * 1.25) + 4096;
    if (Y.L_cap > cap_total + 4096) Y.L_cap = cap_total + 4096;
    Y.nv_cap = (uint64_t)((double)cap_total / world * 1.1) + 4096;
    uint64_t at = 0;
    auto take = [&](uint64_t bytes) { uint64_t o = at; at += (bytes + 255) & ~255ULL; return o; };
    Y.off_flags = take(P2P_N_BARRIERS * P2P_MAX_WORLD * 4);
    Y.off_err = take(256);
    Y.off_counts = take((uint64_t)world * n_asm_max * 8);
    Y.off_rec_cnt = take(((uint64_t)Y.nb_own_max + 1) * 4);        // world == 1: records per bucket; world > 1: first record of every bucket (+ end)
    Y.off_nkeep = take((uint64_t)Y.n_buckets * 4);
    // world > 1: minimizer records arrive per source in one sequential segment (coalesced remote stores, one TLB-friendly
    // stream per peer) and the OWNER cuts them into buckets with local stores
    Y.cap1 = world > 1 ? (uint64_t)((double)cap_total / world / world * 1.3) + 8192 : 0;
    Y.off_rec = take(std::max<uint64_t>((uint64_t)Y.nb_own_max * world * Y.cap_sub, (uint64_t)world * Y.cap1) * sizeof(P2PRecord));
    // shared memory of a bucket CTA: 4 x bk_max x 8 bytes.  512 records (16 KB, 14 CTAs per SM) where the average bucket
    // at full capacity leaves room for its fluctuations, else 1024
    const double bk_need = (double)(cap_total >> B) * 1.5 + 64.0;
    Y.bk_max = bk_need <= 256.0 ? 256 : bk_need <= 512.0 ? 512 : 1024;
    if (const char* sv = getenv("MXE_P2P_BKMAX")) { const int v = atoi(sv); if (v == 256 || v == 512 || v == 1024) Y.bk_max = v; }
    while (world == 1 && Y.cap_sub > (uint32_t)Y.bk_max && Y.bk_max < 1024) Y.bk_max *= 2;
    Y.off_mk = take(Y.L_cap * 4);
    Y.off_tab = take((uint64_t)n_asm_max * Y.nv_cap * 4);          // vertex tables, interleaved by assembly (p2p_tab, p2p_vgid)
    Y.off_vgid = take(world > 1 ? (uint64_t)n_asm_max * Y.nv_cap * 4 : 0);
    // world > 1: every sighting travels to the owners of its two vertices as a 24-byte record (no remote loads)
    Y.cap2 = world > 1 ? (uint64_t)((double)cap_total / world / world * 2.0 * 1.3) + 8192 : 0;
    Y.off_pred = take(world > 1 ? (uint64_t)n_asm_max * Y.nv_cap * 4 : 0);
    Y.off_cnt2 = take(256);
    Y.off_rec2 = take((uint64_t)world * Y.cap2 * 24);
    Y.off_cnt1 = take(256);
    Y.off_seg1 = take((uint64_t)world * Y.cap1 * 16);
    Y.bytes = at;
    {
        // Load every kernel of this file now.  With lazy module loading the FIRST launch of a kernel can wait for the
        // device to go idle; a rank whose stream holds a spinning barrier kernel would then stall its own host thread
        // (and, when several ranks share a process, everybody) until the barrier times out.
        cudaFuncAttributes fa;
        const void* kernels[] = {(const void*)p2p_signal_kernel, (const void*)p2p_wait_kernel, (const void*)p2p_scatter_kernel,
                                 (const void*)p2p_push_counts_kernel, (const void*)p2p_scatter_seg_kernel, (const void*)p2p_push_cnt1_kernel,
                                 (const void*)p2p_part_count_kernel, (const void*)p2p_part_start_kernel, (const void*)p2p_part_place_kernel,
                                 (const void*)p2p_bucket_kernel<256>, (const void*)p2p_bucket_kernel<512>, (const void*)p2p_bucket_kernel<1024>, (const void*)p2p_vbase_kernel,
                                 (const void*)p2p_vertices_kernel, (const void*)p2p_flags_kernel, (const void*)p2p_compact_kernel,
                                 (const void*)p2p_succ_kernel, (const void*)p2p_edge_owner_kernel, (const void*)p2p_first_count_kernel,
                                 (const void*)p2p_first_start_kernel, (const void*)p2p_edge_emit_kernel,
                                 (const void*)p2p_sight_kernel, (const void*)p2p_push_cnt2_kernel, (const void*)p2p_table_kernel,
                                 (const void*)p2p_rec_owner_kernel, (const void*)p2p_rec_emit_kernel};
        for (const void* k : kernels) cudaFuncGetAttributes(&fa, k);
        cudaGetLastError();
    }
    cudaError_t err = cudaMalloc((void**)&X->ws, Y.bytes);
    if (err != cudaSuccess) { set_error("symmetric workspace of %llu bytes: %s", (unsigned long long)Y.bytes, cudaGetErrorString(err)); delete X; return MXE_ERR_NOMEM; }
    cudaMemset(X->ws, 0, Y.off_rec);      // flags, error word, count tables
    for (int r = 0; r < P2P_MAX_WORLD; r++) X->P.base[r] = nullptr;
    X->P.base[rank] = X->ws;
    *out = X;
    return MXE_OK;
}

/* CUDA IPC handle of this rank's symmetric workspace (64 bytes) */
int mxe_p2p_handle(mxe_p2p_t* X, void* handle64, uint64_t* bytes)
{
    if (!X || !handle64) { set_error("null argument"); return MXE_ERR_ARG; }
    cudaIpcMemHandle_t h;
    MXE_CUDA(cudaIpcGetMemHandle(&h, X->ws));
    memcpy(handle64, &h, sizeof(h));
    if (bytes) *bytes = X->Y.bytes;
    return MXE_OK;
}

/* handles: world x 64 bytes, in rank order (one process per GPU) */
int mxe_p2p_connect(mxe_p2p_t* X, const void* handles)
{
    if (!X || !handles) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(X->eng->device));
    for (int r = 0; r < X->Y.world; r++) {
        if (r == X->Y.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess) { set_error("cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(err)); return MXE_ERR_CUDA; }
        X->opened.push_back(p);
        X->P.base[r] = (char*)p;
    }
    return MXE_OK;
}

/* same-process ranks (tests: several engines on one device): raw workspace pointers instead of IPC */
int mxe_p2p_workspace(mxe_p2p_t* X, void** ptr)
{
    if (!X || !ptr) { set_error("null argument"); return MXE_ERR_ARG; }
    *ptr = X->ws;
    return MXE_OK;
}
int mxe_p2p_connect_pointers(mxe_p2p_t* X, void* const* bases)
{
    if (!X || !bases) { set_error("null argument"); return MXE_ERR_ARG; }
    for (int r = 0; r < X->Y.world; r++) X->P.base[r] = (char*)bases[r];
    return MXE_OK;
}

void mxe_p2p_free(mxe_p2p_t* X)
{
    if (!X) return;
    mxe::p2p_release(X, mxe::engine_alive(X->eng));
}

/* stage 1: scatter this rank's minimizers to the bucket owners.  d_hash[a] / d_contig[a]: n[a] uint64 out_hash in
 * (record, pos) order and their uint32 record ids (device arrays that stay alive until mxe_p2p_finish). */
int mxe_p2p_scatter(mxe_p2p_t* X, const void* const* d_hash, const void* const* d_contig, const uint64_t* n, int n_asm, const double* weights)
{
    if (!X || !d_hash || !d_contig || !n || n_asm < 1 || n_asm > X->Y.n_asm_max) { set_error("bad arguments"); return MXE_ERR_ARG; }
    mxe_engine* e = X->eng;
    MXE_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    X->Y.n_asm = n_asm;
    const P2PLayout& Y = X->Y;
    Span whole(e, "filter");
    Span part(e, "p2p_scatter");
    X->n_asm = n_asm;
    X->S.n = n_asm; X->S.lofs[0] = 0;
    X->A.n = n_asm;
    for (int a = 0; a < n_asm; a++) {
        X->S.lofs[a + 1] = X->S.lofs[a] + n[a]; X->S.goff[a] = 0;
        X->H.p[a] = d_hash[a]; X->Ctg.p[a] = d_contig[a];
        X->A.weight[a] = weights[a];
    }
    const uint64_t L = X->S.lofs[n_asm];
    X->L = L;
    if (L > Y.L_cap) { set_error("%llu minimizers on this rank exceed the workspace capacity %llu", (unsigned long long)L, (unsigned long long)Y.L_cap); return MXE_ERR_ARG; }
    // local workspace for this call
    const uint64_t n_own = Y.world > 1 ? std::max<uint64_t>(L + 1, (uint64_t)Y.world * Y.cap2 + 1) : L + 1;      // entries of own[] / emask_j[] / uprefix[]
    const size_t need = (size_t)Y.n_buckets * 8 + (L + 1) * (5 * 4 + 8 + 2) + n_own * (2 * 4 + 8) + (Y.n_buckets + 64) * 4 + 4096 * 4 + 64 * 1024;
    if (X->lws_bytes < need) {
        if (X->lws) { MXE_CUDA(cudaStreamSynchronize(st)); MXE_CUDA(cudaFree(X->lws)); X->lws = nullptr; }
        MXE_CUDA(cudaMalloc((void**)&X->lws, need + need / 8));
        X->lws_bytes = need + need / 8;
    }
    X->lws_used = 0;
    X->cursor = (uint32_t*)X->lalloc((size_t)Y.n_buckets * 4);
    X->kflag = (uint32_t*)X->lalloc((L + 1) * 4); X->cvid = (uint32_t*)X->lalloc((L + 1) * 4); X->cg = (uint32_t*)X->lalloc((L + 1) * 4);
    X->cloc = (uint32_t*)X->lalloc((L + 1) * 4); X->eflag = (uint32_t*)X->lalloc((L + 1) * 4); X->own = (uint32_t*)X->lalloc(n_own * 4);
    X->emask_j = (uint32_t*)X->lalloc(n_own * 4);
    X->kprefix = (uint64_t*)X->lalloc((L + 2) * 8); X->uprefix = (uint64_t*)X->lalloc((n_own + 1) * 8);
    X->cur2 = (uint32_t*)X->lalloc(64 * 4);
    X->cur1 = (uint32_t*)X->lalloc(64 * 4);
    X->T.vbase = (uint32_t*)X->lalloc((size_t)(Y.n_buckets + 1) * 4); X->T.vown = (uint32_t*)X->lalloc(64 * 4); X->T.goff = (uint64_t*)X->lalloc(64 * 8);
    if (!X->T.goff) { set_error("local workspace too small"); return MXE_ERR_INTERNAL; }
    X->epoch++;
    MXE_CUDA(cudaMemsetAsync(X->cursor, 0, (size_t)Y.n_buckets * 4, st));
    MXE_CUDA(cudaMemsetAsync(X->cur2, 0, 64 * 4, st));
    MXE_CUDA(cudaMemsetAsync(X->cur1, 0, 64 * 4, st));
    // barrier 0: every rank has finished the previous call (its reads of peer tables included) before anybody clears
    // its own tables or writes into a peer's workspace again
    MXE_TRY(p2p_barrier_signal(X, 0));
    MXE_TRY(p2p_barrier_wait(X, 0));
    // vertex tables: in records mode (world > 1) only the vertices with a successor / predecessor get their entries
    // written, so the tables start from zero; otherwise p2p_succ_kernel writes every entry of every vertex
    const bool tables_by_records = Y.world > 1 && X->records;
    if (tables_by_records) MXE_CUDA(cudaMemsetAsync(X->ws + Y.off_tab, 0, (size_t)n_asm * Y.nv_cap * 4, st));
    // marks of own minimizers: every record that reaches its bucket gets one; a record dropped by an overflowing bucket
    // (the call then fails and falls back) must not leave an unwritten word behind
    if (L) MXE_CUDA(cudaMemsetAsync(X->ws + Y.off_mk, 0, L * 4, st));
    if (tables_by_records) MXE_CUDA(cudaMemsetAsync(X->ws + Y.off_pred, 0, (size_t)n_asm * Y.nv_cap * 4, st));
    AsmCounts C;
    C.n_asm = n_asm;
    for (int a = 0; a < n_asm; a++) C.n[a] = n[a];
    if (Y.world == 1) {
        if (L) { Span k_(e, "k_p2p_scatter_kernel"); MXE_LAUNCH(e, p2p_scatter_kernel, p2p_grid(L), 256, 0, X->H, X->S, L, X->P, Y, X->cursor); }
        { Span k_(e, "k_p2p_push_counts_kernel"); MXE_LAUNCH(e, p2p_push_counts_kernel, p2p_grid(std::max<uint64_t>(Y.n_buckets, (uint64_t)Y.world * n_asm)), 256, 0, X->cursor, C, X->P, Y); }
    } else {
        // (the bucket counts of the owner-side partition of stage 2 are taken in X->cursor, cleared above)
        if (L) { Span k_(e, "k_p2p_scatter_seg_kernel"); MXE_LAUNCH(e, p2p_scatter_seg_kernel, p2p_grid(L), 256, 0, X->H, X->S, L, X->P, Y, X->cur1); }
        { Span k_(e, "k_p2p_push_cnt1_kernel"); MXE_LAUNCH(e, p2p_push_cnt1_kernel, 1, 512, 0, X->cur1, C, X->P, Y); }
    }
    MXE_TRY(p2p_barrier_signal(X, 1));
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

/* stage 2: the owned buckets */
int mxe_p2p_buckets(mxe_p2p_t* X)
{
    mxe_engine* e = X->eng;
    MXE_CUDA(cudaSetDevice(e->device));
    const P2PLayout& Y = X->Y;
    Span whole(e, "filter");
    Span part(e, "p2p_buckets");
    MXE_TRY(p2p_barrier_wait(X, 1));
    const uint32_t nb = Y.fb[Y.rank + 1] - Y.fb[Y.rank];
    const size_t bk_smem = (size_t)4 * Y.bk_max * sizeof(uint64_t);
    if (Y.world > 1) {
        const unsigned grid = p2p_grid((uint64_t)Y.world * Y.cap1);
        { Span k_(e, "k_p2p_partition_kernel"); MXE_LAUNCH(e, p2p_part_count_kernel, grid, 256, 0, X->P, Y, X->cursor); }
        MXE_LAUNCH(e, p2p_part_start_kernel, 1, 1024, 0, X->P, Y, X->cursor);
        { Span k_(e, "k_p2p_partition_kernel"); MXE_LAUNCH(e, p2p_part_place_kernel, grid, 256, 0, X->P, Y, X->cursor); }
    }
    if (nb) {
        Span k_(e, "k_p2p_bucket_kernel");
        if (Y.bk_max == 256) {
            MXE_LAUNCH(e, p2p_bucket_kernel<256>, nb, P2P_BK_THREADS, bk_smem, X->P, Y, X->n_asm);
        } else if (Y.bk_max == 512) {
            MXE_CUDA(cudaFuncSetAttribute(p2p_bucket_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_smem));
            MXE_LAUNCH(e, p2p_bucket_kernel<512>, nb, P2P_BK_THREADS, bk_smem, X->P, Y, X->n_asm);
        } else {
            MXE_CUDA(cudaFuncSetAttribute(p2p_bucket_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bk_smem));
            MXE_LAUNCH(e, p2p_bucket_kernel<1024>, nb, P2P_BK_THREADS, bk_smem, X->P, Y, X->n_asm);
        }
    }
    MXE_TRY(p2p_barrier_signal(X, 2));
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

/* stage 3: vertex ids, ordered survivors of own records, successor entries to the vertex owners */
int mxe_p2p_adjacency(mxe_p2p_t* X)
{
    mxe_engine* e = X->eng;
    MXE_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    const P2PLayout& Y = X->Y;
    Span whole(e, "filter");
    Span part(e, "p2p_adjacency");
    const uint64_t L = X->L;
    MXE_TRY(p2p_barrier_wait(X, 2));
    { Span k_(e, "k_p2p_vbase_kernel"); MXE_LAUNCH(e, p2p_vbase_kernel, 1, 1024, 0, X->P, Y, X->n_asm, X->T); }
    // shard outputs that do not depend on sizes read back: flags of own minimizers
    X->luniq = nullptr; X->lkeep = nullptr;
    MXE_CUDA(cudaMallocAsync((void**)&X->luniq, L ? L : 1, st));
    MXE_CUDA(cudaMallocAsync((void**)&X->lkeep, L ? L : 1, st));
    const uint32_t* mk = reinterpret_cast<const uint32_t*>(X->ws + Y.off_mk);
    if (L) { Span k_(e, "k_p2p_flags_kernel"); MXE_LAUNCH(e, p2p_flags_kernel, p2p_grid(L), 256, 0, mk, L, X->kflag, X->luniq, X->lkeep); }
    {
        ArenaScope scope(e);
        MXE_TRY(exclusive_scan_u32_u64(e, X->kflag, X->kprefix, L));
    }
    if (L) {
        { Span k_(e, "k_p2p_compact_kernel"); MXE_LAUNCH(e, p2p_compact_kernel, p2p_grid(L), 256, 0, mk, X->H, X->S, L, X->kprefix, Y, X->T, X->cvid, X->cg, X->cloc); }
        if (Y.world == 1 || !X->records)
            { Span k_(e, "k_p2p_succ_kernel"); MXE_LAUNCH(e, p2p_succ_kernel, p2p_grid(L), 256, 0, X->cvid, X->cg, X->cloc, X->kprefix, L, X->S, X->Ctg, X->P, Y, X->T, X->eflag); }
        else
            { Span k_(e, "k_p2p_sight_kernel"); MXE_LAUNCH(e, p2p_sight_kernel, p2p_grid(L), 256, 0, X->cvid, X->cg, X->cloc, X->kprefix, L, X->S, X->H, X->Ctg, X->P, Y, X->T, X->cur2); }
    }
    if (Y.world > 1 && X->records) { Span k_(e, "k_p2p_push_cnt2_kernel"); MXE_LAUNCH(e, p2p_push_cnt2_kernel, 1, 32, 0, X->cur2, X->P, Y); }
    MXE_TRY(p2p_barrier_signal(X, 3));
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

/* stage 4: support masks, edge ownership, first-source tables at the vertex owners */
int mxe_p2p_edges(mxe_p2p_t* X)
{
    mxe_engine* e = X->eng;
    MXE_CUDA(cudaSetDevice(e->device));
    const P2PLayout& Y = X->Y;
    Span whole(e, "filter");
    Span part(e, "p2p_edges");
    const uint64_t L = X->L;
    MXE_TRY(p2p_barrier_wait(X, 3));
    if (Y.world == 1 || !X->records) {
        if (L) { Span k_(e, "k_p2p_edge_owner_kernel"); MXE_LAUNCH(e, p2p_edge_owner_kernel, p2p_grid(L), 256, 0, X->cvid, X->cg, X->cloc, X->eflag, X->kprefix, L, X->S, X->n_asm, X->P, Y, X->T, X->own, X->emask_j); }
        MXE_TRY(p2p_barrier_signal(X, 4));
    } else {
        const uint64_t n_idx = (uint64_t)Y.world * Y.cap2;
        { Span k_(e, "k_p2p_table_kernel"); MXE_LAUNCH(e, p2p_table_kernel, p2p_grid(n_idx), 256, 0, X->P, Y); }
        { Span k_(e, "k_p2p_rec_owner_kernel"); MXE_LAUNCH(e, p2p_rec_owner_kernel, p2p_grid(n_idx), 256, 0, X->P, Y, X->n_asm, X->own, X->emask_j); }
    }
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

/* stage 5: sizes (the one host round trip), outputs.  world == 1: the complete result in final order;
 * world > 1: this rank's shard (flags of own minimizers, vertices of the owned hash range, own edges + order keys). */
int mxe_p2p_finish(mxe_p2p_t* X, mxe_result_t** out)
{
    if (!X || !out) { set_error("null argument"); return MXE_ERR_ARG; }
    mxe_engine* e = X->eng;
    MXE_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    const P2PLayout& Y = X->Y;
    Span whole(e, "filter");
    Span part(e, "p2p_finish");
    const uint64_t L = X->L;
    const int n_asm = X->n_asm;
    const bool home = Y.world == 1 || !X->records;       // edges emitted by the rank that sketched the sighting
    if (home) MXE_TRY(p2p_barrier_wait(X, 4));
    // owned edges: world == 1 in creation order over the survivors (p2p_edge_owner_kernel clears own[] past n_keep: the scan runs
    // over L entries); world > 1 over the record index space of this owner
    const uint64_t n_scan = home ? L : (uint64_t)Y.world * Y.cap2;
    {
        ArenaScope scope(e);
        MXE_TRY(exclusive_scan_u32_u64(e, X->own, X->uprefix, n_scan));
    }
    uint64_t sizes[2] = {0, 0};
    uint32_t verts[2] = {0, 0}, errw = 0;
    MXE_CUDA(cudaMemcpyAsync(&sizes[0], X->kprefix + L, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaMemcpyAsync(&sizes[1], X->uprefix + n_scan, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaMemcpyAsync(&verts[0], X->T.vown + Y.rank, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaMemcpyAsync(&errw, X->ws + Y.off_err, 4, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));
    if (errw) {
        cudaMemsetAsync(X->ws + Y.off_err, 0, 4, st);
        cudaFreeAsync(X->luniq, st); cudaFreeAsync(X->lkeep, st);
        X->luniq = X->lkeep = nullptr;
        set_error(errw >= 0x100 ? "device barrier %u timed out (a peer rank is gone)" : errw == 1 ? "bucket sub-slot overflow (code %u)" : errw == 3 ? "sighting segment overflow (code %u)" : "bucket overflow (code %u)",
                  errw >= 0x100 ? errw - 0x100 : errw);
        return errw >= 0x100 ? MXE_ERR_CUDA : MXE_ERR_INTERNAL;
    }
    const uint64_t nE = sizes[1], nV = verts[1] - verts[0];
    mxe_result* R = new mxe_result();
    R->eng = e; R->n_asm = n_asm; R->N = L; R->nV = nV; R->nE = nE;
    for (int a = 0; a <= n_asm; a++) R->asm_off[a] = X->S.lofs[a];
    R->d_uniq = X->luniq; R->d_keep = X->lkeep;
    X->luniq = X->lkeep = nullptr;
    if (nV) {
        MXE_CUDA(cudaMallocAsync((void**)&R->d_vertices, nV * 8, st));
        { Span k_(e, "k_p2p_vertices_kernel"); MXE_LAUNCH(e, p2p_vertices_kernel, (Y.fb[Y.rank + 1] - Y.fb[Y.rank] + 3) / 4, 128, 0, X->P, Y, X->T, R->d_vertices); }
    }
    if (nE) {
        EdgeOut E;
        memset(&E, 0, sizeof(E));
        MXE_CUDA(cudaMallocAsync((void**)&R->d_eu, nE * 8, st)); MXE_CUDA(cudaMallocAsync((void**)&R->d_ev, nE * 8, st));
        MXE_CUDA(cudaMallocAsync((void**)&R->d_emask, nE * 4, st)); MXE_CUDA(cudaMallocAsync((void**)&R->d_ew, nE * 8, st));
        E.eu = R->d_eu; E.ev = R->d_ev; E.emask = R->d_emask; E.ew = R->d_ew;
        uint32_t* vstart = nullptr;
        ArenaScope scope(e);
        DBuf<uint32_t> fcount, vst;
        DBuf<uint64_t> fprefix;
        if (Y.world == 1) {
            MXE_TRY(fcount.alloc(nE, st)); MXE_TRY(fprefix.alloc(nE + 1, st)); MXE_TRY(vst.alloc(nV ? nV : 1, st));
            uint32_t* smask_j = X->eflag;       // the sighting flags are dead after p2p_edge_owner_kernel: the array now holds the source masks
            { Span k_(e, "k_p2p_first_count_kernel"); MXE_LAUNCH(e, p2p_first_count_kernel, p2p_grid(L), 256, 0, X->cvid, X->cloc, X->own, X->uprefix, X->kprefix, L, X->S, X->P, Y, n_asm, fcount.p, smask_j); }
            MXE_TRY(exclusive_scan_u32_u64(e, fcount.p, fprefix.p, nE));
            { Span k_(e, "k_p2p_first_start_kernel"); MXE_LAUNCH(e, p2p_first_start_kernel, p2p_grid(L), 256, 0, X->cvid, X->cloc, X->own, smask_j, X->uprefix, fprefix.p, X->kprefix, L, X->S, vst.p); }
            vstart = vst.p;
            { Span k_(e, "k_p2p_edge_emit_kernel"); MXE_LAUNCH(e, p2p_edge_emit_kernel, p2p_grid(L), 256, 0, X->cvid, X->cg, X->cloc, X->own, X->emask_j, smask_j, X->uprefix, fprefix.p, X->kprefix, L, X->S, X->H, X->A,
                       X->P, Y, X->T, vstart, E); }
        } else if (home) {
            MXE_CUDA(cudaMallocAsync((void**)&R->d_ekey, nE * 8, st));
            E.ekey = R->d_ekey;
            { Span k_(e, "k_p2p_edge_emit_kernel"); MXE_LAUNCH(e, p2p_edge_emit_kernel, p2p_grid(L), 256, 0, X->cvid, X->cg, X->cloc, X->own, X->emask_j, (const uint32_t*)nullptr, X->uprefix, (const uint64_t*)nullptr, X->kprefix, L, X->S, X->H, X->A,
                       X->P, Y, X->T, vstart, E); }
        } else {
            MXE_CUDA(cudaMallocAsync((void**)&R->d_ekey, nE * 8, st));
            E.ekey = R->d_ekey;
            { Span k_(e, "k_p2p_rec_emit_kernel"); MXE_LAUNCH(e, p2p_rec_emit_kernel, p2p_grid(n_scan), 256, 0, X->P, Y, X->A, X->own, X->emask_j, X->uprefix, R->d_vertices, E); }
        }
    }
    MXE_CUDA(cudaGetLastError());
    *out = R;
    return MXE_OK;
}

}  // extern "C"
