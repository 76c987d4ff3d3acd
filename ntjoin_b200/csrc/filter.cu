// filter.cu -- steps 2-3 on the device.
//
// Replaces (reference, Python):
//   bin/ntjoin_utils.py:182-192  per-assembly uniqueness of out_hash           (read_minimizers)
//   bin/ntjoin_utils.py:155-162  found-in-all intersection + ordered filtering  (filter_minimizers)
//   bin/ntjoin_utils.py:94-115   adjacent-pair edge dictionary, support lists   (build_graph)
//   bin/ntjoin_utils.py:54-56    edge weight = sum of assembly weights          (calc_total_weight)
//
// Method: one stable radix sort of all assemblies' out_hash (input concatenated in assembly
// order, so equal hashes stay grouped by assembly) -> run analysis gives `uniq`, `keep` and a
// dense vertex id per surviving hash; ordered compaction of survivors -> adjacent pairs ->
// per-assembly successor/predecessor tables indexed by vertex id give every edge's support
// mask and its first sighting without sorting (a surviving hash occurs once per assembly);
// a stable sort on the source's first-creation index then reproduces the insertion order of
// the reference's dict-of-dicts (formatted_edges, :115).
#include "engine.cuh"

#include <algorithm>

namespace mxe {

__device__ __forceinline__ int asm_of(const AsmOffsets& A, uint64_t idx)
{
    int a = 0;
    while (a + 1 < A.n && idx >= A.off[a + 1]) a++;
    return a;
}

__global__ void __launch_bounds__(256) concat_kernel(const uint64_t* __restrict__ src, uint64_t n, uint64_t base,
                                                      uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[base + i] = src[i];
    vals[base + i] = (uint32_t)(base + i);
}

__global__ void __launch_bounds__(256) copy_contig_kernel(const uint32_t* __restrict__ src, uint64_t n, uint64_t base, uint32_t* __restrict__ dst)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[base + i] = src[i];
}

// The hash sort only covers the top SORT_BITS bits (out_hash is uniformly mixed, so ties in the top 32 bits between
// DIFFERENT hashes are rare: ~n^2/2^33 pairs).  This pass finishes the job: inside every run of equal top bits it
// orders the elements by full key (stable insertion sort; the input order inside a run is the concatenation order,
// i.e. by assembly).  Runs whose keys are all equal (true duplicates, any length) need nothing.  A mixed run longer
// than FIXUP_MAX raises the fallback flag and the caller redoes the sort on all 64 bits.
constexpr int FIXUP_MAX = 64;

// number of low key bits left to the fix-up: sort 24 top bits (3 radix passes) while the expected run of equal top bits
// stays short (N <= 2^25 -> about two elements per occupied bucket), else 32 (4 passes)
static inline int sort_low_bit(const Engine* e, uint64_t N)
{
    if (e->sort_bits == 24 || e->sort_bits == 32 || e->sort_bits == 40) return 64 - e->sort_bits;
    return N <= (1ULL << 25) ? 40 : 32;
}

__global__ void __launch_bounds__(256) fixup_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint64_t N, int SORT_LOW_BIT,
                                                     int* __restrict__ fallback)
{
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const uint64_t top = keys[s] >> SORT_LOW_BIT;
    if (s > 0 && (keys[s - 1] >> SORT_LOW_BIT) == top) return;      // not a run head
    if (s + 1 >= N || (keys[s + 1] >> SORT_LOW_BIT) != top) return; // run of one
    const uint64_t k0 = keys[s];
    uint64_t e = s + 1;
    bool mixed = false;
    while (e < N && (keys[e] >> SORT_LOW_BIT) == top) { mixed |= keys[e] != k0; e++; }
    if (!mixed) return;
    if (e - s > FIXUP_MAX) { *fallback = 1; return; }
    for (uint64_t i = s + 1; i < e; i++) {
        uint64_t kk = keys[i]; uint32_t vv = vals[i];
        uint64_t j = i;
        while (j > s && keys[j - 1] > kk) { keys[j] = keys[j - 1]; vals[j] = vals[j - 1]; j--; }
        keys[j] = kk; vals[j] = vv;
    }
}

// per sorted element: uniqueness inside its assembly, membership in a found-in-all run, run head
__global__ void __launch_bounds__(256) mark_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t N,
                                                    AsmOffsets A, uint8_t* __restrict__ uniq, uint8_t* __restrict__ keep,
                                                    uint32_t* __restrict__ head)
{
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const uint64_t K = keys[s];
    const uint32_t idx = vals[s];
    const int a = asm_of(A, idx);
    bool prev_same = s > 0 && keys[s - 1] == K && asm_of(A, vals[s - 1]) == a;
    bool next_same = s + 1 < N && keys[s + 1] == K && asm_of(A, vals[s + 1]) == a;
    uniq[idx] = !(prev_same || next_same);
    bool in_all = false;
    if ((uint64_t)a <= s && s - a + A.n <= N) {
        uint64_t s0 = s - a;
        in_all = (s0 == 0 || keys[s0 - 1] != K) && (s0 + A.n == N || keys[s0 + A.n] != K);
        for (int j = 0; in_all && j < A.n; j++)
            in_all = keys[s0 + j] == K && asm_of(A, vals[s0 + j]) == j;
    }
    keep[idx] = in_all;
    head[s] = in_all && a == 0;
}

// vertex ids: heads get consecutive ids in ascending hash order; every member of the run shares it
__global__ void __launch_bounds__(256) vertex_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t N,
                                                      AsmOffsets A, const uint8_t* __restrict__ keep, const uint64_t* __restrict__ hprefix,
                                                      uint32_t* __restrict__ vid, uint64_t* __restrict__ vertices)
{
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const uint32_t idx = vals[s];
    if (!keep[idx]) return;
    const int a = asm_of(A, idx);
    uint32_t id = (uint32_t)hprefix[s - a];
    vid[idx] = id;
    if (a == 0) vertices[id] = keys[s];
}

__global__ void __launch_bounds__(256) widen_flags_kernel(const uint8_t* __restrict__ f, uint64_t N, uint32_t* __restrict__ out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i] = f[i];
}

// ordered compaction of survivors
__global__ void __launch_bounds__(256) compact_kernel(const uint8_t* __restrict__ keep, const uint64_t* __restrict__ kprefix, uint64_t N,
                                                       const uint32_t* __restrict__ vid, uint32_t* __restrict__ cvid, uint32_t* __restrict__ cidx)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || !keep[i]) return;
    uint64_t j = kprefix[i];
    cvid[j] = vid[i];
    cidx[j] = (uint32_t)i;
}

// adjacent survivors of the same record and assembly form an edge sighting
__global__ void __launch_bounds__(256) pair_flag_kernel(const uint32_t* __restrict__ cidx, uint64_t n_keep, AsmOffsets A,
                                                         const uint32_t* __restrict__ contig, uint32_t* __restrict__ eflag)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_keep) return;
    uint32_t f = 0;
    if (j + 1 < n_keep) {
        uint32_t i1 = cidx[j], i2 = cidx[j + 1];
        f = asm_of(A, i1) == asm_of(A, i2) && contig[i1] == contig[i2];
    }
    eflag[j] = f;
}

// Edge de-duplication without sorting.  Every surviving hash occurs exactly once per assembly, so a vertex has at
// most one successor per assembly.  Assembly b supports the undirected edge {v,x} iff succ_b[v] == x or succ_b[x] == v
// (one table per assembly: the predecessor of v is x exactly when the successor of x is v); the sighting in the first supporting assembly owns the edge
// (that is the reference's first-seen orientation, bin/ntjoin_utils.py:101-108), so ordered compaction of the owning
// sightings yields the distinct edges already in first-seen order.
__global__ void __launch_bounds__(256) adjacency_kernel(const uint32_t* __restrict__ cvid, const uint32_t* __restrict__ cidx,
                                                         const uint32_t* __restrict__ eflag, uint64_t n_keep, AsmOffsets A, uint64_t nV,
                                                         uint32_t* __restrict__ succ)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_keep || !eflag[j]) return;
    const uint64_t a = (uint64_t)asm_of(A, cidx[j]);
    succ[a * nV + cvid[j]] = cvid[j + 1];
}

__global__ void __launch_bounds__(256) edge_owner_kernel(const uint32_t* __restrict__ cvid, const uint32_t* __restrict__ cidx,
                                                          const uint32_t* __restrict__ eflag, uint64_t n_keep, AsmOffsets A, uint64_t nV,
                                                          const uint32_t* __restrict__ succ,
                                                          uint32_t* __restrict__ own, uint32_t* __restrict__ mask_out)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_keep) return;
    uint32_t is_owner = 0;
    if (eflag[j]) {
        const int a = asm_of(A, cidx[j]);
        const uint32_t v = cvid[j], x = cvid[j + 1];
        uint32_t mask = 0;
        for (int b = 0; b < A.n; b++)
            if (succ[(uint64_t)b * nV + v] == x || succ[(uint64_t)b * nV + x] == v) mask |= 1u << b;
        is_owner = (__ffs(mask) - 1) == a;
        mask_out[j] = mask;
    }
    own[j] = is_owner;
}

__global__ void __launch_bounds__(256) edge_compact_kernel(const uint32_t* __restrict__ own, const uint64_t* __restrict__ uprefix,
                                                            const uint32_t* __restrict__ mask_in, uint64_t n_keep, const uint32_t* __restrict__ cvid,
                                                            uint32_t* __restrict__ ue_q0, uint32_t* __restrict__ ue_mask, uint32_t* __restrict__ srcmin)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_keep || !own[j]) return;
    uint64_t t = uprefix[j];
    ue_q0[t] = (uint32_t)j;
    ue_mask[t] = mask_in[j];
    atomicMin(&srcmin[cvid[j]], (uint32_t)j);     // first time this vertex becomes the source of a new edge
}

// formatted_edges order (bin/ntjoin_utils.py:115): sources in order of their first edge, edges of a source in creation
// order.  Edges are already in creation order, so a STABLE sort on the source's first-creation index suffices.
__global__ void __launch_bounds__(256) edge_order_key_kernel(const uint32_t* __restrict__ ue_q0, uint64_t n_edges, const uint32_t* __restrict__ cvid,
                                                              const uint32_t* __restrict__ srcmin, uint64_t* __restrict__ okey, uint32_t* __restrict__ oval)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_edges) return;
    okey[t] = srcmin[cvid[ue_q0[t]]];
    oval[t] = (uint32_t)t;
}

__global__ void __launch_bounds__(256) edge_gather_kernel(const uint32_t* __restrict__ oval, uint64_t n_edges, const uint32_t* __restrict__ ue_q0,
                                                           const uint32_t* __restrict__ ue_mask, const uint32_t* __restrict__ cvid,
                                                           const uint64_t* __restrict__ vertices, AsmOffsets A,
                                                           uint64_t* __restrict__ eu, uint64_t* __restrict__ ev, uint32_t* __restrict__ emask,
                                                           double* __restrict__ ew)
{
    uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_edges) return;
    uint32_t t = oval[o];
    uint32_t q0 = ue_q0[t], mask = ue_mask[t];
    eu[o] = vertices[cvid[q0]];
    ev[o] = vertices[cvid[q0 + 1]];
    emask[o] = mask;
    double wsum = 0.0;   // Python: sum() starts at int 0 and adds in support-list (= assembly) order
    for (int a = 0; a < A.n; a++)
        if (mask & (1u << a)) wsum += A.weight[a];
    ew[o] = wsum;
}

static inline unsigned gridf(uint64_t n) { return (unsigned)((n + 255) / 256); }
static inline int bits_for(uint64_t v) { int b = 1; while (b < 32 && (1ULL << b) <= v) b++; return ((b + 7) / 8) * 8; }


int filter_and_edges_impl(mxe_engine* e, const uint64_t* const* d_hash, const uint32_t* const* d_contig,
                          const uint64_t* n, int n_asm, const double* weights, mxe_result* R)
{
    if (n_asm < 1 || n_asm > 32) { set_error("n_asm must be 1..32"); return MXE_ERR_ARG; }
    cudaStream_t st = e->stream;
    Span whole(e, "filter");
    AsmOffsets A;
    A.n = n_asm;
    A.off[0] = 0;
    for (int a = 0; a < n_asm; a++) { A.off[a + 1] = A.off[a] + n[a]; A.weight[a] = weights[a]; }
    const uint64_t N = A.off[n_asm];
    if (N >= (1ULL << 32)) { set_error("too many minimizers (%llu)", (unsigned long long)N); return MXE_ERR_ARG; }
    R->eng = e; R->n_asm = n_asm; R->N = N;
    for (int a = 0; a <= n_asm; a++) R->asm_off[a] = A.off[a];
    if (N == 0) return MXE_OK;

    DBuf<uint64_t> keys, keys2, hprefix, kprefix, vertices;
    DBuf<uint32_t> vals, vals2, contig, head, vid, kflag;
    DBuf<uint8_t> uniq, keep;
    MXE_TRY(keys.alloc(N, st)); MXE_TRY(keys2.alloc(N, st));
    MXE_TRY(vals.alloc(N, st)); MXE_TRY(vals2.alloc(N, st));
    MXE_TRY(contig.alloc(N, st)); MXE_TRY(head.alloc(N, st)); MXE_TRY(vid.alloc(N, st)); MXE_TRY(kflag.alloc(N, st));
    MXE_TRY(uniq.alloc(N, st)); MXE_TRY(keep.alloc(N, st));
    uint8_t* const keep_p = keep.p;
    MXE_TRY(hprefix.alloc(N + 1, st)); MXE_TRY(kprefix.alloc(N + 1, st));
    for (int a = 0; a < n_asm; a++) {
        if (!n[a]) continue;
        MXE_LAUNCH(e, concat_kernel, gridf(n[a]), 256, 0, d_hash[a], n[a], A.off[a], keys.p, vals.p);
        MXE_LAUNCH(e, copy_contig_kernel, gridf(n[a]), 256, 0, d_contig[a], n[a], A.off[a], contig.p);
    }
    {
        // sort the top 32 bits, finish inside the (tiny) tied runs; full 64-bit sort only if that is not enough
        DBuf<int> fallback;
        MXE_TRY(fallback.alloc(1, st));
        MXE_CUDA(cudaMemsetAsync(fallback.p, 0, sizeof(int), st));
        const int SORT_LOW_BIT = sort_low_bit(e, N);
        MXE_TRY(radix_sort_pairs(e, keys.p, vals.p, keys2.p, vals2.p, N, SORT_LOW_BIT, 64));
        MXE_LAUNCH(e, fixup_kernel, gridf(N), 256, 0, keys.p, vals.p, N, SORT_LOW_BIT, fallback.p);
        int fb = 0;
        MXE_CUDA(cudaMemcpyAsync(&fb, fallback.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaStreamSynchronize(st));
        if (fb) MXE_TRY(radix_sort_pairs(e, keys.p, vals.p, keys2.p, vals2.p, N, 0, SORT_LOW_BIT));   // LSD: low bits, then...
        if (fb) MXE_TRY(radix_sort_pairs(e, keys.p, vals.p, keys2.p, vals2.p, N, SORT_LOW_BIT, 64));  // ...high bits again (stable)
    }
    MXE_LAUNCH(e, mark_kernel, gridf(N), 256, 0, keys.p, vals.p, N, A, uniq.p, keep_p, head.p);
    MXE_TRY(exclusive_scan_u32_u64(e, head.p, hprefix.p, N));
    MXE_LAUNCH(e, widen_flags_kernel, gridf(N), 256, 0, keep_p, N, kflag.p);
    MXE_TRY(exclusive_scan_u32_u64(e, kflag.p, kprefix.p, N));
    uint64_t tot[2];
    MXE_CUDA(cudaMemcpyAsync(&tot[0], hprefix.p + N, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaMemcpyAsync(&tot[1], kprefix.p + N, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));
    const uint64_t nV = tot[0], n_keep = tot[1];

    R->d_uniq = uniq.detach(); R->d_keep = keep.detach();
    if (nV == 0) return MXE_OK;

    MXE_TRY(vertices.alloc(nV, st));
    MXE_LAUNCH(e, vertex_kernel, gridf(N), 256, 0, keys.p, vals.p, N, A, keep_p, hprefix.p, vid.p, vertices.p);
    R->nV = nV;

    DBuf<uint32_t> cvid, cidx, eflag;
    DBuf<uint64_t> eprefix;
    MXE_TRY(cvid.alloc(n_keep + 1, st)); MXE_TRY(cidx.alloc(n_keep + 1, st)); MXE_TRY(eflag.alloc(n_keep, st));
    MXE_TRY(eprefix.alloc(n_keep + 1, st));
    MXE_LAUNCH(e, compact_kernel, gridf(N), 256, 0, keep_p, kprefix.p, N, vid.p, cvid.p, cidx.p);
    MXE_LAUNCH(e, pair_flag_kernel, gridf(n_keep), 256, 0, cidx.p, n_keep, A, contig.p, eflag.p);
    MXE_TRY(exclusive_scan_u32_u64(e, eflag.p, eprefix.p, n_keep));
    uint64_t n_pairs = 0;
    MXE_CUDA(cudaMemcpyAsync(&n_pairs, eprefix.p + n_keep, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));
    if (n_pairs == 0) { R->d_vertices = vertices.detach(); return MXE_OK; }

    DBuf<uint32_t> succ, own, emask_j, srcmin;
    DBuf<uint64_t> uprefix;
    MXE_TRY(succ.alloc((uint64_t)n_asm * nV, st));
    MXE_TRY(own.alloc(n_keep, st)); MXE_TRY(emask_j.alloc(n_keep, st)); MXE_TRY(uprefix.alloc(n_keep + 1, st));
    MXE_TRY(srcmin.alloc(nV, st));
    MXE_CUDA(cudaMemsetAsync(succ.p, 0xFF, (uint64_t)n_asm * nV * sizeof(uint32_t), st));
    MXE_CUDA(cudaMemsetAsync(srcmin.p, 0xFF, nV * sizeof(uint32_t), st));
    MXE_LAUNCH(e, adjacency_kernel, gridf(n_keep), 256, 0, cvid.p, cidx.p, eflag.p, n_keep, A, nV, succ.p);
    MXE_LAUNCH(e, edge_owner_kernel, gridf(n_keep), 256, 0, cvid.p, cidx.p, eflag.p, n_keep, A, nV, succ.p, own.p, emask_j.p);
    MXE_TRY(exclusive_scan_u32_u64(e, own.p, uprefix.p, n_keep));
    uint64_t nE = 0;
    MXE_CUDA(cudaMemcpyAsync(&nE, uprefix.p + n_keep, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));

    DBuf<uint32_t> ue_q0, ue_mask, oval, oval2, emask;
    DBuf<uint64_t> okey, okey2, eu, ev;
    DBuf<double> ew;
    MXE_TRY(ue_q0.alloc(nE, st)); MXE_TRY(ue_mask.alloc(nE, st)); MXE_TRY(oval.alloc(nE, st)); MXE_TRY(oval2.alloc(nE, st));
    MXE_TRY(okey.alloc(nE, st)); MXE_TRY(okey2.alloc(nE, st));
    MXE_TRY(eu.alloc(nE, st)); MXE_TRY(ev.alloc(nE, st)); MXE_TRY(emask.alloc(nE, st)); MXE_TRY(ew.alloc(nE, st));
    MXE_LAUNCH(e, edge_compact_kernel, gridf(n_keep), 256, 0, own.p, uprefix.p, emask_j.p, n_keep, cvid.p, ue_q0.p, ue_mask.p, srcmin.p);
    MXE_LAUNCH(e, edge_order_key_kernel, gridf(nE), 256, 0, ue_q0.p, nE, cvid.p, srcmin.p, okey.p, oval.p);
    MXE_TRY(radix_sort_pairs(e, okey.p, oval.p, okey2.p, oval2.p, nE, 0, bits_for(n_keep)));
    MXE_LAUNCH(e, edge_gather_kernel, gridf(nE), 256, 0, oval.p, nE, ue_q0.p, ue_mask.p, cvid.p, vertices.p, A, eu.p, ev.p, emask.p, ew.p);
    R->nE = nE;
    R->d_vertices = vertices.detach();
    R->d_eu = eu.detach(); R->d_ev = ev.detach(); R->d_emask = emask.detach(); R->d_ew = ew.detach();
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}


// ====================================================================================================
// Multi-GPU steps 2-3 (one process per GPU).  The collectives between the stages belong
// to the caller (torch.distributed / NCCL over NVLink); every stage works on GLOBALLY indexed arrays,
// so each rank does 1/world of the single-GPU work and the partial tables are combined exactly by
// integer all-reduces:
//   global index space  : assemblies in order, inside an assembly the ranks in order (each rank holds
//                         a contiguous range of records), N = all minimizers of all assemblies
//   hash ownership      : rank r owns the hashes h with ((h >> 32) * world) >> 32 == r  (monotone in h,
//                         so per-rank vertex lists concatenate to the ascending single-GPU order)
//   stage 1 mark        : uniqueness / found-in-all / local vertex ids for the owned hash range -> mk[N]
//                         (zero outside the owned entries)                        -> all-reduce(sum) mk
//   stage 2 adjacency   : this rank's survivors (its own records) -> successor table indexed by
//                         (assembly, GLOBAL vertex id), zero outside own sightings -> all-reduce(sum)
//   stage 3 edges       : support mask + ownership of the local sightings, srcmin -> all-reduce(min)
//   stage 4 finish      : local edge shard ordered by (first-creation index of the source, creation
//                         index); merging the shards by that key reproduces formatted_edges order.
// Uniqueness is per ASSEMBLY, not per GPU (bin/ntjoin_utils.py:182-187): stage 1 sees the full multiset.
// ====================================================================================================
__device__ __forceinline__ uint32_t hash_owner(uint64_t h, int world) { return (uint32_t)(((h >> 32) * (uint64_t)world) >> 32); }

__global__ void __launch_bounds__(256) owner_flag_kernel(const uint64_t* __restrict__ keys, uint64_t N, int rank, int world, uint32_t* __restrict__ flag)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) flag[i] = hash_owner(keys[i], world) == (uint32_t)rank;
}

__global__ void __launch_bounds__(256) owner_compact_kernel(const uint64_t* __restrict__ keys, uint64_t N, const uint32_t* __restrict__ flag,
                                                             const uint64_t* __restrict__ prefix, uint64_t* __restrict__ skeys, uint32_t* __restrict__ svals)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || !flag[i]) return;
    uint64_t j = prefix[i];
    skeys[j] = keys[i];
    svals[j] = (uint32_t)i;
}

// mk[idx] = uniq << 31 | (keep ? 1 + local vertex id : 0) for the owned entries (the rest of mk stays zero)
__global__ void __launch_bounds__(256) pack_mk_kernel(const uint32_t* __restrict__ svals, uint64_t n_sel, const uint8_t* __restrict__ uniq,
                                                       const uint8_t* __restrict__ keep, const uint32_t* __restrict__ vid, uint32_t* __restrict__ mk)
{
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sel) return;
    const uint32_t idx = svals[s];
    mk[idx] = ((uint32_t)uniq[idx] << 31) | (keep[idx] ? vid[idx] + 1u : 0u);
}

__device__ __forceinline__ int slice_of(const LocalSlices& S, uint64_t l)
{
    int a = 0;
    while (a + 1 < S.n && l >= S.lofs[a + 1]) a++;
    return a;
}

__global__ void __launch_bounds__(256) local_flag_kernel(const uint32_t* __restrict__ mk, LocalSlices S, uint64_t L,
                                                          uint32_t* __restrict__ kflag, uint8_t* __restrict__ luniq, uint8_t* __restrict__ lkeep)
{
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const int a = slice_of(S, l);
    const uint32_t m = mk[S.goff[a] + (l - S.lofs[a])];
    const uint32_t k = (m & 0x7FFFFFFFu) != 0;
    kflag[l] = k;
    luniq[l] = (uint8_t)(m >> 31);
    lkeep[l] = (uint8_t)k;
}

__global__ void __launch_bounds__(256) local_compact_kernel(const uint32_t* __restrict__ mk, const uint64_t* __restrict__ keys, LocalSlices S, uint64_t L,
                                                             const uint32_t* __restrict__ kflag, const uint64_t* __restrict__ kprefix, DistInfo D,
                                                             uint32_t* __restrict__ cvid, uint32_t* __restrict__ cidx, uint32_t* __restrict__ cloc)
{
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L || !kflag[l]) return;
    const int a = slice_of(S, l);
    const uint64_t g = S.goff[a] + (l - S.lofs[a]);
    const uint64_t j = kprefix[l];
    cvid[j] = (uint32_t)(D.vbase[hash_owner(keys[g], D.world)] + ((mk[g] & 0x7FFFFFFFu) - 1u));
    cidx[j] = (uint32_t)g;
    cloc[j] = (uint32_t)l;
}

__global__ void __launch_bounds__(256) local_pair_flag_kernel(const uint32_t* __restrict__ cloc, uint64_t n_keep, LocalSlices S,
                                                               const uint32_t* __restrict__ lcontig, uint32_t* __restrict__ eflag)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_keep) return;
    uint32_t f = 0;
    if (j + 1 < n_keep) {
        const uint32_t l1 = cloc[j], l2 = cloc[j + 1];
        f = slice_of(S, l1) == slice_of(S, l2) && lcontig[l1] == lcontig[l2];
    }
    eflag[j] = f;
}

// successor entries are 1 + vertex id (0 = none) so that the tables of all ranks combine by summation
__global__ void __launch_bounds__(256) dist_adjacency_kernel(const uint32_t* __restrict__ cvid, const uint32_t* __restrict__ cidx,
                                                              const uint32_t* __restrict__ eflag, uint64_t n_keep, AsmOffsets A, uint64_t nV,
                                                              uint32_t* __restrict__ succ)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_keep || !eflag[j]) return;
    const uint64_t a = (uint64_t)asm_of(A, cidx[j]);
    succ[a * nV + cvid[j]] = cvid[j + 1] + 1u;
}

__global__ void __launch_bounds__(256) dist_edge_owner_kernel(const uint32_t* __restrict__ cvid, const uint32_t* __restrict__ cidx,
                                                               const uint32_t* __restrict__ eflag, uint64_t n_keep, AsmOffsets A, uint64_t nV,
                                                               const uint32_t* __restrict__ succ,
                                                               uint32_t* __restrict__ own, uint32_t* __restrict__ mask_out)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_keep) return;
    uint32_t is_owner = 0;
    if (eflag[j]) {
        const int a = asm_of(A, cidx[j]);
        const uint32_t v = cvid[j], x = cvid[j + 1];
        uint32_t mask = 0;
        for (int b = 0; b < A.n; b++)
            if (succ[(uint64_t)b * nV + v] == x + 1u || succ[(uint64_t)b * nV + x] == v + 1u) mask |= 1u << b;
        is_owner = (__ffs(mask) - 1) == a;
        mask_out[j] = mask;
    }
    own[j] = is_owner;
}

// creation index of an edge = global index of its source element (monotone in the single-GPU sighting order)
__global__ void __launch_bounds__(256) dist_edge_compact_kernel(const uint32_t* __restrict__ own, const uint64_t* __restrict__ uprefix,
                                                                 const uint32_t* __restrict__ mask_in, uint64_t n_keep, const uint32_t* __restrict__ cvid,
                                                                 const uint32_t* __restrict__ cidx, uint32_t* __restrict__ ue_q0, uint32_t* __restrict__ ue_mask,
                                                                 uint32_t* __restrict__ srcmin)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_keep || !own[j]) return;
    uint64_t t = uprefix[j];
    ue_q0[t] = (uint32_t)j;
    ue_mask[t] = mask_in[j];
    atomicMin(&srcmin[cvid[j]], cidx[j]);
}

__global__ void __launch_bounds__(256) dist_edge_key_kernel(const uint32_t* __restrict__ ue_q0, uint64_t n_edges, const uint32_t* __restrict__ cvid,
                                                             const uint32_t* __restrict__ srcmin, uint64_t* __restrict__ okey, uint32_t* __restrict__ oval)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_edges) return;
    okey[t] = srcmin[cvid[ue_q0[t]]];
    oval[t] = (uint32_t)t;
}

__global__ void __launch_bounds__(256) dist_edge_gather_kernel(const uint32_t* __restrict__ oval, const uint64_t* __restrict__ okey, uint64_t n_edges,
                                                                const uint32_t* __restrict__ ue_q0, const uint32_t* __restrict__ ue_mask,
                                                                const uint32_t* __restrict__ cidx, const uint64_t* __restrict__ keys, AsmOffsets A,
                                                                uint64_t* __restrict__ eu, uint64_t* __restrict__ ev, uint32_t* __restrict__ emask,
                                                                double* __restrict__ ew, uint64_t* __restrict__ ekey)
{
    uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_edges) return;
    const uint32_t t = oval[o];
    const uint32_t q0 = ue_q0[t], mask = ue_mask[t];
    eu[o] = keys[cidx[q0]];
    ev[o] = keys[cidx[q0 + 1]];
    emask[o] = mask;
    double wsum = 0.0;
    for (int a = 0; a < A.n; a++)
        if (mask & (1u << a)) wsum += A.weight[a];
    ew[o] = wsum;
    ekey[o] = (okey[o] << 32) | (uint64_t)cidx[q0];
}

int dist_mark_impl(mxe_engine* e, const uint64_t* d_keys, const uint64_t* asm_off, int n_asm, int rank, int world,
                   uint32_t* d_mk, mxe_dist* X, uint64_t* nv_local)
{
    if (n_asm < 1 || n_asm > 32 || world < 1 || world > 32 || rank < 0 || rank >= world) { set_error("bad n_asm/rank/world"); return MXE_ERR_ARG; }
    cudaStream_t st = e->stream;
    Span whole(e, "filter");
    Span part(e, "dist_mark");
    AsmOffsets& A = X->A;
    A.n = n_asm;
    for (int a = 0; a <= n_asm; a++) A.off[a] = asm_off[a];
    for (int a = 0; a < n_asm; a++) A.weight[a] = 0.0;
    const uint64_t N = A.off[n_asm];
    if (N >= 0x7F000000ULL) { set_error("too many minimizers for the multi-GPU path (%llu)", (unsigned long long)N); return MXE_ERR_ARG; }
    X->eng = e; X->rank = rank; X->world = world; X->N = N; X->d_keys = d_keys;
    *nv_local = 0;
    MXE_CUDA(cudaMemsetAsync(d_mk, 0, (N ? N : 1) * sizeof(uint32_t), st));
    if (N == 0) return MXE_OK;

    DBuf<uint32_t> flag, svals, svals2, head, vid;
    DBuf<uint64_t> prefix, skeys, skeys2, hprefix;
    DBuf<uint8_t> uniq, keep;
    MXE_TRY(flag.alloc(N, st)); MXE_TRY(prefix.alloc(N + 1, st));
    MXE_LAUNCH(e, owner_flag_kernel, gridf(N), 256, 0, d_keys, N, rank, world, flag.p);
    MXE_TRY(exclusive_scan_u32_u64(e, flag.p, prefix.p, N));
    uint64_t n_sel = 0;
    MXE_CUDA(cudaMemcpyAsync(&n_sel, prefix.p + N, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));
    if (n_sel == 0) return MXE_OK;
    MXE_TRY(skeys.alloc(n_sel, st)); MXE_TRY(skeys2.alloc(n_sel, st)); MXE_TRY(svals.alloc(n_sel, st)); MXE_TRY(svals2.alloc(n_sel, st));
    MXE_LAUNCH(e, owner_compact_kernel, gridf(N), 256, 0, d_keys, N, flag.p, prefix.p, skeys.p, svals.p);
    {
        DBuf<int> fallback;
        MXE_TRY(fallback.alloc(1, st));
        MXE_CUDA(cudaMemsetAsync(fallback.p, 0, sizeof(int), st));
        // the owned range spans 1/world of the key space: its top log2(world) bits carry (almost) no information
        const int SORT_LOW_BIT = sort_low_bit(e, N);
        MXE_TRY(radix_sort_pairs(e, skeys.p, svals.p, skeys2.p, svals2.p, n_sel, SORT_LOW_BIT, 64));
        MXE_LAUNCH(e, fixup_kernel, gridf(n_sel), 256, 0, skeys.p, svals.p, n_sel, SORT_LOW_BIT, fallback.p);
        int fb = 0;
        MXE_CUDA(cudaMemcpyAsync(&fb, fallback.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaStreamSynchronize(st));
        if (fb) MXE_TRY(radix_sort_pairs(e, skeys.p, svals.p, skeys2.p, svals2.p, n_sel, 0, SORT_LOW_BIT));
        if (fb) MXE_TRY(radix_sort_pairs(e, skeys.p, svals.p, skeys2.p, svals2.p, n_sel, SORT_LOW_BIT, 64));
    }
    // the compaction kept the global order, so equal hashes are still grouped by assembly after the stable sort
    MXE_TRY(uniq.alloc(N, st)); MXE_TRY(keep.alloc(N, st)); MXE_TRY(head.alloc(n_sel, st)); MXE_TRY(hprefix.alloc(n_sel + 1, st));
    MXE_TRY(vid.alloc(N, st));
    MXE_LAUNCH(e, mark_kernel, gridf(n_sel), 256, 0, skeys.p, svals.p, n_sel, A, uniq.p, keep.p, head.p);
    MXE_TRY(exclusive_scan_u32_u64(e, head.p, hprefix.p, n_sel));
    uint64_t nV = 0;
    MXE_CUDA(cudaMemcpyAsync(&nV, hprefix.p + n_sel, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));
    MXE_TRY(X->alloc(&X->vertices, nV));
    MXE_LAUNCH(e, vertex_kernel, gridf(n_sel), 256, 0, skeys.p, svals.p, n_sel, A, keep.p, hprefix.p, vid.p, X->vertices);
    MXE_LAUNCH(e, pack_mk_kernel, gridf(n_sel), 256, 0, svals.p, n_sel, uniq.p, keep.p, vid.p, d_mk);
    X->nV_local = nV;
    *nv_local = nV;
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

int dist_adjacency_impl(mxe_dist* X, const uint32_t* d_mk, const uint64_t* vbase, const uint64_t* loc_off, const uint64_t* loc_n,
                        const uint32_t* const* d_contig, uint32_t* d_succ)
{
    mxe_engine* e = X->eng;
    cudaStream_t st = e->stream;
    Span whole(e, "filter");
    Span part(e, "dist_adjacency");
    const int n_asm = X->A.n;
    LocalSlices& S = X->S;
    S.n = n_asm; S.lofs[0] = 0;
    for (int a = 0; a < n_asm; a++) {
        if (loc_off[a] + loc_n[a] > X->A.off[a + 1] || loc_off[a] < X->A.off[a]) { set_error("local slice of assembly %d outside its global range", a); return MXE_ERR_ARG; }
        S.goff[a] = loc_off[a]; S.lofs[a + 1] = S.lofs[a] + loc_n[a];
    }
    const uint64_t L = S.lofs[n_asm];
    X->L = L;
    X->D.world = X->world;
    for (int r = 0; r <= X->world; r++) X->D.vbase[r] = vbase[r];
    const uint64_t nV = vbase[X->world];
    X->nV = nV;
    if (nV >= 0x7FFFFFFFULL) { set_error("too many vertices"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaMemsetAsync(d_succ, 0, (size_t)n_asm * (nV ? nV : 1) * sizeof(uint32_t), st));
    MXE_TRY(X->alloc(&X->luniq, L)); MXE_TRY(X->alloc(&X->lkeep, L));
    X->n_keep = 0;
    if (L == 0) return MXE_OK;

    DBuf<uint32_t> kflag, lcontig;
    DBuf<uint64_t> kprefix;
    MXE_TRY(kflag.alloc(L, st)); MXE_TRY(kprefix.alloc(L + 1, st)); MXE_TRY(lcontig.alloc(L, st));
    for (int a = 0; a < n_asm; a++)
        if (loc_n[a]) MXE_LAUNCH(e, copy_contig_kernel, gridf(loc_n[a]), 256, 0, d_contig[a], loc_n[a], S.lofs[a], lcontig.p);
    MXE_LAUNCH(e, local_flag_kernel, gridf(L), 256, 0, d_mk, S, L, kflag.p, X->luniq, X->lkeep);
    MXE_TRY(exclusive_scan_u32_u64(e, kflag.p, kprefix.p, L));
    uint64_t n_keep = 0;
    MXE_CUDA(cudaMemcpyAsync(&n_keep, kprefix.p + L, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));
    X->n_keep = n_keep;
    MXE_TRY(X->alloc(&X->cvid, n_keep + 1)); MXE_TRY(X->alloc(&X->cidx, n_keep + 1)); MXE_TRY(X->alloc(&X->cloc, n_keep + 1));
    MXE_TRY(X->alloc(&X->eflag, n_keep + 1));
    if (n_keep == 0) return MXE_OK;
    MXE_LAUNCH(e, local_compact_kernel, gridf(L), 256, 0, d_mk, X->d_keys, S, L, kflag.p, kprefix.p, X->D, X->cvid, X->cidx, X->cloc);
    MXE_LAUNCH(e, local_pair_flag_kernel, gridf(n_keep), 256, 0, X->cloc, n_keep, S, lcontig.p, X->eflag);
    MXE_LAUNCH(e, dist_adjacency_kernel, gridf(n_keep), 256, 0, X->cvid, X->cidx, X->eflag, n_keep, X->A, nV, d_succ);
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

int dist_edges_impl(mxe_dist* X, const uint32_t* d_succ, uint32_t* d_srcmin, uint64_t* n_edges_local)
{
    mxe_engine* e = X->eng;
    cudaStream_t st = e->stream;
    Span whole(e, "filter");
    Span part(e, "dist_edges");
    const uint64_t n_keep = X->n_keep, nV = X->nV;
    *n_edges_local = 0;
    X->nE = 0;
    MXE_CUDA(cudaMemsetAsync(d_srcmin, 0x7F, (nV ? nV : 1) * sizeof(uint32_t), st));     // 0x7f7f7f7f > any global index (N < 0x7f000000)
    if (n_keep == 0) return MXE_OK;
    DBuf<uint32_t> own, emask_j;
    DBuf<uint64_t> uprefix;
    MXE_TRY(own.alloc(n_keep, st)); MXE_TRY(emask_j.alloc(n_keep, st)); MXE_TRY(uprefix.alloc(n_keep + 1, st));
    MXE_LAUNCH(e, dist_edge_owner_kernel, gridf(n_keep), 256, 0, X->cvid, X->cidx, X->eflag, n_keep, X->A, nV, d_succ, own.p, emask_j.p);
    MXE_TRY(exclusive_scan_u32_u64(e, own.p, uprefix.p, n_keep));
    uint64_t nE = 0;
    MXE_CUDA(cudaMemcpyAsync(&nE, uprefix.p + n_keep, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));
    X->nE = nE;
    *n_edges_local = nE;
    MXE_TRY(X->alloc(&X->ue_q0, nE)); MXE_TRY(X->alloc(&X->ue_mask, nE));
    if (nE) MXE_LAUNCH(e, dist_edge_compact_kernel, gridf(n_keep), 256, 0, own.p, uprefix.p, emask_j.p, n_keep, X->cvid, X->cidx, X->ue_q0, X->ue_mask, d_srcmin);
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

int dist_finish_impl(mxe_dist* X, const uint32_t* d_srcmin, const double* weights, mxe_result* R)
{
    mxe_engine* e = X->eng;
    cudaStream_t st = e->stream;
    Span whole(e, "filter");
    Span part(e, "dist_finish");
    const int n_asm = X->A.n;
    for (int a = 0; a < n_asm; a++) X->A.weight[a] = weights[a];
    R->eng = e; R->n_asm = n_asm; R->N = X->L; R->nV = X->nV_local; R->nE = X->nE;
    for (int a = 0; a <= n_asm; a++) R->asm_off[a] = X->S.lofs[a];
    // hand the per-rank shard over to the result object (ownership moves out of X)
    auto take = [&](void* p) { for (auto& q : X->owned) if (q == p) q = nullptr; return p; };
    R->d_uniq = (uint8_t*)take(X->luniq); R->d_keep = (uint8_t*)take(X->lkeep);
    R->d_vertices = (uint64_t*)take(X->vertices);
    const uint64_t nE = X->nE;
    if (nE == 0) return MXE_OK;
    DBuf<uint32_t> oval, oval2, emask;
    DBuf<uint64_t> okey, okey2, eu, ev, ekey;
    DBuf<double> ew;
    MXE_TRY(oval.alloc(nE, st)); MXE_TRY(oval2.alloc(nE, st)); MXE_TRY(okey.alloc(nE, st)); MXE_TRY(okey2.alloc(nE, st));
    MXE_TRY(eu.alloc(nE, st)); MXE_TRY(ev.alloc(nE, st)); MXE_TRY(emask.alloc(nE, st)); MXE_TRY(ew.alloc(nE, st)); MXE_TRY(ekey.alloc(nE, st));
    MXE_LAUNCH(e, dist_edge_key_kernel, gridf(nE), 256, 0, X->ue_q0, nE, X->cvid, d_srcmin, okey.p, oval.p);
    MXE_TRY(radix_sort_pairs(e, okey.p, oval.p, okey2.p, oval2.p, nE, 0, bits_for(X->N)));
    MXE_LAUNCH(e, dist_edge_gather_kernel, gridf(nE), 256, 0, oval.p, okey.p, nE, X->ue_q0, X->ue_mask, X->cidx, X->d_keys, X->A,
               eu.p, ev.p, emask.p, ew.p, ekey.p);
    R->d_eu = eu.detach(); R->d_ev = ev.detach(); R->d_emask = emask.detach(); R->d_ew = ew.detach(); R->d_ekey = ekey.detach();
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}


// ====================================================================================================
// Multi-GPU steps 2-3, all-to-all formulation: per-rank work and traffic shrink with 1/world (the all-reduce
// formulation above keeps O(N) arrays on every rank and is the faster one while a step is latency-bound, i.e. at
// BASELINE configs[2] scale on 2-8 GPUs).  Every exchange moves each item exactly once, to the rank that needs it:
//
//   A partition   (source)   own minimizers grouped by HASH OWNER            -> all-to-all(keys, 8 B each)
//   B mark        (owner)    unique / found-in-all / vertex ids of the range  -> all-to-all(marks, 4 B, same order back)
//   C sightings   (source)   ordered survivors of own records, adjacent pairs;
//                            each sighting (a: v -> x) goes to owner(v) as SUCC
//                            and to owner(x) as PRED                         -> all-to-all(records, 24 B)
//   D finish      (owner)    succ/pred tables of OWN vertices, support masks, edge ownership, first-source
//                            index, order keys: all local.  Result shard = edges whose source vertex is owned.
//
// A vertex is named (owner << 27 | local id) everywhere, so no global vertex numbering is needed on the
// device.  Creation index of a sighting = GLOBAL index of its source element (assemblies in order, ranks
// in order inside an assembly), which is what makes the merged edge order independent of the GPU count.
// ====================================================================================================
constexpr int A2A_VBITS = 27;                    // local vertex id bits (world <= 16)
struct PtrTable { const void* p[32]; };
struct SegTable { uint64_t dst[1025]; uint64_t src[1024]; int n; };   // segments sorted by dst; dst[n] = total

__global__ void __launch_bounds__(256) a2a_owner_key_kernel(PtrTable H, LocalSlices S, uint64_t L, int world, int n_asm,
                                                             uint64_t* __restrict__ lhash, uint64_t* __restrict__ okey, uint32_t* __restrict__ oval,
                                                             unsigned long long* __restrict__ counts)
{
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = l < L;
    uint32_t bucket = 0xFFFFFFFFu;
    if (ok) {
        const int a = slice_of(S, l);
        const uint64_t h = reinterpret_cast<const uint64_t*>(H.p[a])[l - S.lofs[a]];
        const uint32_t o = hash_owner(h, world);
        lhash[l] = h;
        okey[l] = o;
        oval[l] = (uint32_t)l;
        bucket = o * (uint32_t)n_asm + (uint32_t)a;
    }
    // warp-aggregated count per (owner, assembly)
    const uint32_t peers = __match_any_sync(0xffffffffu, bucket);
    if (ok && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&counts[bucket], (unsigned long long)__popc(peers));
}

__global__ void __launch_bounds__(256) gather_u64_kernel(const uint64_t* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n, uint64_t* __restrict__ dst)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[j] = src[idx[j]];
}

__device__ __forceinline__ int seg_of(const SegTable& T, uint64_t p)
{
    int lo = 0, hi = T.n;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (T.dst[mid] <= p) lo = mid; else hi = mid; }
    return lo;
}

// received order (source rank, assembly, i) -> assembly-major order (assembly, source rank, i): equal hashes must stay
// grouped by assembly through the stable sort (mark_kernel)
__global__ void __launch_bounds__(256) a2a_regroup_kernel(const uint64_t* __restrict__ recv, uint64_t n, const SegTable* __restrict__ Tp,
                                                           uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ srcpos)
{
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const SegTable& T = *Tp;
    const int sg = seg_of(T, p);
    const uint64_t src = T.src[sg] + (p - T.dst[sg]);
    keys[p] = recv[src];
    vals[p] = (uint32_t)p;
    srcpos[p] = (uint32_t)src;
}

__global__ void __launch_bounds__(256) a2a_marks_return_kernel(const uint32_t* __restrict__ srcpos, uint64_t n, const uint8_t* __restrict__ uniq,
                                                                const uint8_t* __restrict__ keep, const uint32_t* __restrict__ vid, uint32_t* __restrict__ ret)
{
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    ret[srcpos[p]] = ((uint32_t)uniq[p] << 31) | (keep[p] ? vid[p] + 1u : 0u);
}

// marks come back in send order; perm[j] = local index of send position j
__global__ void __launch_bounds__(256) a2a_local_marks_kernel(const uint32_t* __restrict__ marks_in, const uint32_t* __restrict__ perm, uint64_t L,
                                                               uint32_t* __restrict__ kflag, uint32_t* __restrict__ lmark,
                                                               uint8_t* __restrict__ luniq, uint8_t* __restrict__ lkeep)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= L) return;
    const uint32_t l = perm[j], m = marks_in[j];
    const uint32_t k = (m & 0x7FFFFFFFu) != 0;
    lmark[l] = m;
    kflag[l] = k;
    luniq[l] = (uint8_t)(m >> 31);
    lkeep[l] = (uint8_t)k;
}

struct GlobalOffsets { uint64_t g[32]; };

__global__ void __launch_bounds__(256) a2a_compact_kernel(const uint32_t* __restrict__ lmark, const uint64_t* __restrict__ lhash, LocalSlices S, uint64_t L,
                                                           const uint32_t* __restrict__ kflag, const uint64_t* __restrict__ kprefix, GlobalOffsets G, int world,
                                                           uint32_t* __restrict__ cid, uint64_t* __restrict__ chash, uint32_t* __restrict__ cg, uint32_t* __restrict__ cloc)
{
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L || !kflag[l]) return;
    const int a = slice_of(S, l);
    const uint64_t j = kprefix[l];
    const uint64_t h = lhash[l];
    cid[j] = (hash_owner(h, world) << A2A_VBITS) | ((lmark[l] & 0x7FFFFFFFu) - 1u);
    chash[j] = h;
    cg[j] = (uint32_t)(G.g[a] + (l - S.lofs[a]));
    cloc[j] = (uint32_t)l;
}

// two records per sighting (a: v -> x, creation index g):
//   SUCC -> owner(v): { hash(x), local(v) << 32 | id(x), g << 8 | a << 1 | 0 }
//   PRED -> owner(x): { hash(v), local(x) << 32 | id(v),          a << 1 | 1 }
// written at 2*e, 2*e+1 (e = rank of the sighting among this rank's sightings) with their destination as sort key
__global__ void __launch_bounds__(256) a2a_records_kernel(const uint32_t* __restrict__ cid, const uint64_t* __restrict__ chash, const uint32_t* __restrict__ cg,
                                                           const uint32_t* __restrict__ cloc, const uint32_t* __restrict__ eflag, const uint64_t* __restrict__ eprefix,
                                                           uint64_t n_keep, LocalSlices S, uint64_t* __restrict__ rec, uint64_t* __restrict__ dkey,
                                                           uint32_t* __restrict__ dval, unsigned long long* __restrict__ counts)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = j < n_keep && eflag[j];
    uint32_t d0 = 0xFFFFFFFFu, d1 = 0xFFFFFFFFu;
    if (ok) {
        const uint64_t e = eprefix[j];
        const uint32_t v = cid[j], x = cid[j + 1];
        const uint64_t a = (uint64_t)slice_of(S, cloc[j]);
        d0 = v >> A2A_VBITS;
        d1 = x >> A2A_VBITS;
        const uint32_t vmask = (1u << A2A_VBITS) - 1u;
        uint64_t* r0 = rec + 6 * e;
        r0[0] = chash[j + 1]; r0[1] = ((uint64_t)(v & vmask) << 32) | x; r0[2] = ((uint64_t)cg[j] << 8) | (a << 1);
        r0[3] = chash[j];     r0[4] = ((uint64_t)(x & vmask) << 32) | v; r0[5] = (a << 1) | 1ULL;
        dkey[2 * e] = d0; dval[2 * e] = (uint32_t)(2 * e);
        dkey[2 * e + 1] = d1; dval[2 * e + 1] = (uint32_t)(2 * e + 1);
    }
    uint32_t peers = __match_any_sync(0xffffffffu, d0);
    if (ok && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&counts[d0], (unsigned long long)__popc(peers));
    peers = __match_any_sync(0xffffffffu, d1);
    if (ok && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&counts[d1], (unsigned long long)__popc(peers));
}

__global__ void __launch_bounds__(256) gather_rec_kernel(const uint64_t* __restrict__ rec, const uint32_t* __restrict__ idx, uint64_t n, uint64_t* __restrict__ dst)
{
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t* r = rec + 3 * (uint64_t)idx[j];
    dst[3 * j] = r[0]; dst[3 * j + 1] = r[1]; dst[3 * j + 2] = r[2];
}

// owner side: successor / predecessor of every own vertex in every assembly (entries 1 + vertex name, 0 = none)
__global__ void __launch_bounds__(256) a2a_table_kernel(const uint64_t* __restrict__ rec, uint64_t n_rec, uint64_t nV,
                                                         uint32_t* __restrict__ succ, uint32_t* __restrict__ pred)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rec) return;
    const uint64_t w1 = rec[3 * i + 1], w2 = rec[3 * i + 2];
    const uint64_t a = (w2 >> 1) & 0x7F, vloc = w1 >> 32;
    const uint32_t other = (uint32_t)w1 + 1u;
    if (w2 & 1ULL) pred[a * nV + vloc] = other; else succ[a * nV + vloc] = other;
}

__global__ void __launch_bounds__(256) a2a_edge_owner_kernel(const uint64_t* __restrict__ rec, uint64_t n_rec, uint64_t nV, int n_asm,
                                                              const uint32_t* __restrict__ succ, const uint32_t* __restrict__ pred,
                                                              uint32_t* __restrict__ own, uint32_t* __restrict__ mask_out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rec) return;
    const uint64_t w1 = rec[3 * i + 1], w2 = rec[3 * i + 2];
    uint32_t is_owner = 0;
    if (!(w2 & 1ULL)) {
        const int a = (int)((w2 >> 1) & 0x7F);
        const uint64_t vloc = w1 >> 32;
        const uint32_t x1 = (uint32_t)w1 + 1u;
        uint32_t mask = 0;
        for (int b = 0; b < n_asm; b++)
            if (succ[(uint64_t)b * nV + vloc] == x1 || pred[(uint64_t)b * nV + vloc] == x1) mask |= 1u << b;
        is_owner = (__ffs(mask) - 1) == a;
        mask_out[i] = mask;
    }
    own[i] = is_owner;
}

__global__ void __launch_bounds__(256) a2a_edge_compact_kernel(const uint64_t* __restrict__ rec, uint64_t n_rec, const uint32_t* __restrict__ own,
                                                                const uint64_t* __restrict__ uprefix, const uint32_t* __restrict__ mask_in,
                                                                uint32_t* __restrict__ e_rec, uint32_t* __restrict__ e_mask, uint32_t* __restrict__ srcmin)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rec || !own[i]) return;
    const uint64_t t = uprefix[i];
    e_rec[t] = (uint32_t)i;
    e_mask[t] = mask_in[i];
    atomicMin(&srcmin[rec[3 * i + 1] >> 32], (uint32_t)(rec[3 * i + 2] >> 8));
}

// a source owns at most one edge per assembly and its edges are created in assembly order, so (first creation index of
// the source, assembly) orders the shard exactly like the reference's dict-of-dicts
__global__ void __launch_bounds__(256) a2a_edge_key_kernel(const uint64_t* __restrict__ rec, const uint32_t* __restrict__ e_rec, uint64_t n_edges,
                                                            const uint32_t* __restrict__ srcmin, uint64_t* __restrict__ okey, uint32_t* __restrict__ oval)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_edges) return;
    const uint64_t i = e_rec[t];
    okey[t] = ((uint64_t)srcmin[rec[3 * i + 1] >> 32] << 5) | ((rec[3 * i + 2] >> 1) & 0x1F);
    oval[t] = (uint32_t)t;
}

__global__ void __launch_bounds__(256) a2a_edge_gather_kernel(const uint32_t* __restrict__ oval, const uint64_t* __restrict__ okey, uint64_t n_edges,
                                                               const uint64_t* __restrict__ rec, const uint32_t* __restrict__ e_rec, const uint32_t* __restrict__ e_mask,
                                                               const uint64_t* __restrict__ vertices, AsmOffsets A,
                                                               uint64_t* __restrict__ eu, uint64_t* __restrict__ ev, uint32_t* __restrict__ emask,
                                                               double* __restrict__ ew, uint64_t* __restrict__ ekey)
{
    uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_edges) return;
    const uint32_t t = oval[o];
    const uint64_t i = e_rec[t];
    const uint32_t mask = e_mask[t];
    eu[o] = vertices[rec[3 * i + 1] >> 32];
    ev[o] = rec[3 * i];
    emask[o] = mask;
    double wsum = 0.0;
    for (int a = 0; a < A.n; a++)
        if (mask & (1u << a)) wsum += A.weight[a];
    ew[o] = wsum;
    ekey[o] = ((okey[o] >> 5) << 32) | (rec[3 * i + 2] >> 8);
}

int a2a_partition_impl(mxe_engine* e, const uint64_t* const* d_hash, const uint64_t* n, int n_asm, int rank, int world,
                       mxe_a2a* X, uint64_t* counts, const void** d_send_keys)
{
    if (n_asm < 1 || n_asm > 32 || world < 1 || world > 16 || rank < 0 || rank >= world) { set_error("bad n_asm/rank/world (world <= 16)"); return MXE_ERR_ARG; }
    cudaStream_t st = e->stream;
    Span whole(e, "filter");
    Span part(e, "a2a_partition");
    X->eng = e; X->rank = rank; X->world = world; X->n_asm = n_asm;
    LocalSlices& S = X->S;
    S.n = n_asm; S.lofs[0] = 0;
    PtrTable H;
    for (int a = 0; a < n_asm; a++) { S.lofs[a + 1] = S.lofs[a] + n[a]; S.goff[a] = 0; H.p[a] = d_hash[a]; }
    const uint64_t L = S.lofs[n_asm];
    X->L = L;
    if (L >= (1ULL << 31)) { set_error("too many local minimizers"); return MXE_ERR_ARG; }
    for (int i = 0; i < world * n_asm; i++) counts[i] = 0;
    *d_send_keys = nullptr;
    MXE_TRY(X->alloc(&X->lhash, L)); MXE_TRY(X->alloc(&X->perm, L)); MXE_TRY(X->alloc(&X->send_keys, L));
    *d_send_keys = X->send_keys;
    if (L == 0) return MXE_OK;
    DBuf<uint64_t> okey, okey2;
    DBuf<uint32_t> oval2;
    DBuf<unsigned long long> cnt;
    MXE_TRY(okey.alloc(L, st)); MXE_TRY(okey2.alloc(L, st)); MXE_TRY(oval2.alloc(L, st)); MXE_TRY(cnt.alloc((size_t)world * n_asm, st));
    MXE_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)world * n_asm * sizeof(unsigned long long), st));
    MXE_LAUNCH(e, a2a_owner_key_kernel, gridf(L), 256, 0, H, S, L, world, n_asm, X->lhash, okey.p, X->perm, cnt.p);
    MXE_TRY(radix_sort_pairs(e, okey.p, X->perm, okey2.p, oval2.p, L, 0, 8));        // stable partition by owner
    MXE_LAUNCH(e, gather_u64_kernel, gridf(L), 256, 0, X->lhash, X->perm, L, X->send_keys);
    MXE_CUDA(cudaMemcpyAsync(counts, cnt.p, (size_t)world * n_asm * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

int a2a_mark_impl(mxe_a2a* X, const uint64_t* d_recv, const uint64_t* recv_counts, uint32_t* d_ret, uint64_t* nv_local)
{
    mxe_engine* e = X->eng;
    cudaStream_t st = e->stream;
    Span whole(e, "filter");
    Span part(e, "a2a_mark");
    const int n_asm = X->n_asm, world = X->world;
    // segment table: destination order (assembly, source rank), source order (source rank, assembly)
    SegTable T;
    T.n = 0;
    AsmOffsets& A = X->A;
    A.n = n_asm;
    uint64_t at = 0;
    std::vector<uint64_t> srcoff((size_t)world * n_asm);
    {
        uint64_t s = 0;
        for (int r = 0; r < world; r++) for (int a = 0; a < n_asm; a++) { srcoff[(size_t)r * n_asm + a] = s; s += recv_counts[(size_t)r * n_asm + a]; }
    }
    for (int a = 0; a < n_asm; a++) {
        A.off[a] = at;
        for (int r = 0; r < world; r++) {
            const uint64_t c = recv_counts[(size_t)r * n_asm + a];
            if (!c) continue;
            T.dst[T.n] = at; T.src[T.n] = srcoff[(size_t)r * n_asm + a]; T.n++;
            at += c;
        }
    }
    A.off[n_asm] = at;
    T.dst[T.n] = at;
    const uint64_t n_recv = at;
    X->n_recv = n_recv;
    *nv_local = 0;
    X->nV_local = 0;
    if (n_recv >= (1ULL << 31)) { set_error("too many minimizers in one hash range"); return MXE_ERR_ARG; }
    if (n_recv == 0) return MXE_OK;
    DBuf<SegTable> dT;
    MXE_TRY(dT.alloc(1, st));
    MXE_CUDA(cudaMemcpyAsync(dT.p, &T, sizeof(SegTable), cudaMemcpyHostToDevice, st));
    DBuf<uint64_t> keys, keys2, hprefix;
    DBuf<uint32_t> vals, vals2, srcpos, head, vid;
    DBuf<uint8_t> uniq, keep;
    MXE_TRY(keys.alloc(n_recv, st)); MXE_TRY(keys2.alloc(n_recv, st)); MXE_TRY(vals.alloc(n_recv, st)); MXE_TRY(vals2.alloc(n_recv, st));
    MXE_TRY(srcpos.alloc(n_recv, st)); MXE_TRY(head.alloc(n_recv, st)); MXE_TRY(vid.alloc(n_recv, st));
    MXE_TRY(uniq.alloc(n_recv, st)); MXE_TRY(keep.alloc(n_recv, st)); MXE_TRY(hprefix.alloc(n_recv + 1, st));
    MXE_LAUNCH(e, a2a_regroup_kernel, gridf(n_recv), 256, 0, d_recv, n_recv, dT.p, keys.p, vals.p, srcpos.p);
    {
        DBuf<int> fallback;
        MXE_TRY(fallback.alloc(1, st));
        MXE_CUDA(cudaMemsetAsync(fallback.p, 0, sizeof(int), st));
        const int SORT_LOW_BIT = sort_low_bit(e, n_recv * (uint64_t)world);
        MXE_TRY(radix_sort_pairs(e, keys.p, vals.p, keys2.p, vals2.p, n_recv, SORT_LOW_BIT, 64));
        MXE_LAUNCH(e, fixup_kernel, gridf(n_recv), 256, 0, keys.p, vals.p, n_recv, SORT_LOW_BIT, fallback.p);
        int fb = 0;
        MXE_CUDA(cudaMemcpyAsync(&fb, fallback.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaStreamSynchronize(st));        // also orders the host-side SegTable upload before T goes out of scope
        if (fb) MXE_TRY(radix_sort_pairs(e, keys.p, vals.p, keys2.p, vals2.p, n_recv, 0, SORT_LOW_BIT));
        if (fb) MXE_TRY(radix_sort_pairs(e, keys.p, vals.p, keys2.p, vals2.p, n_recv, SORT_LOW_BIT, 64));
    }
    MXE_LAUNCH(e, mark_kernel, gridf(n_recv), 256, 0, keys.p, vals.p, n_recv, A, uniq.p, keep.p, head.p);
    MXE_TRY(exclusive_scan_u32_u64(e, head.p, hprefix.p, n_recv));
    uint64_t nV = 0;
    MXE_CUDA(cudaMemcpyAsync(&nV, hprefix.p + n_recv, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));
    if (nV >= (1ULL << A2A_VBITS)) { set_error("too many vertices in one hash range (%llu)", (unsigned long long)nV); return MXE_ERR_ARG; }
    MXE_TRY(X->alloc(&X->vertices, nV));
    MXE_LAUNCH(e, vertex_kernel, gridf(n_recv), 256, 0, keys.p, vals.p, n_recv, A, keep.p, hprefix.p, vid.p, X->vertices);
    MXE_LAUNCH(e, a2a_marks_return_kernel, gridf(n_recv), 256, 0, srcpos.p, n_recv, uniq.p, keep.p, vid.p, d_ret);
    X->nV_local = nV;
    *nv_local = nV;
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

int a2a_sightings_impl(mxe_a2a* X, const uint32_t* d_marks, const uint32_t* const* d_contig, const uint64_t* goff,
                       uint64_t* rec_counts, const void** d_send_records)
{
    mxe_engine* e = X->eng;
    cudaStream_t st = e->stream;
    Span whole(e, "filter");
    Span part(e, "a2a_sightings");
    const int n_asm = X->n_asm, world = X->world;
    const LocalSlices& S = X->S;
    const uint64_t L = X->L;
    for (int r = 0; r < world; r++) rec_counts[r] = 0;
    *d_send_records = nullptr;
    MXE_TRY(X->alloc(&X->luniq, L)); MXE_TRY(X->alloc(&X->lkeep, L));
    if (L == 0) return MXE_OK;
    GlobalOffsets G;
    for (int a = 0; a < n_asm; a++) G.g[a] = goff[a];
    DBuf<uint32_t> kflag, lmark, lcontig;
    DBuf<uint64_t> kprefix;
    MXE_TRY(kflag.alloc(L, st)); MXE_TRY(lmark.alloc(L, st)); MXE_TRY(lcontig.alloc(L, st)); MXE_TRY(kprefix.alloc(L + 1, st));
    for (int a = 0; a < n_asm; a++) {
        const uint64_t na = S.lofs[a + 1] - S.lofs[a];
        if (na) MXE_LAUNCH(e, copy_contig_kernel, gridf(na), 256, 0, d_contig[a], na, S.lofs[a], lcontig.p);
    }
    MXE_LAUNCH(e, a2a_local_marks_kernel, gridf(L), 256, 0, d_marks, X->perm, L, kflag.p, lmark.p, X->luniq, X->lkeep);
    MXE_TRY(exclusive_scan_u32_u64(e, kflag.p, kprefix.p, L));
    uint64_t n_keep = 0;
    MXE_CUDA(cudaMemcpyAsync(&n_keep, kprefix.p + L, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));
    if (n_keep < 2) return MXE_OK;
    DBuf<uint32_t> cid, cg, cloc, eflag, dval, dval2;
    DBuf<uint64_t> chash, eprefix, rec, dkey, dkey2;
    DBuf<unsigned long long> cnt;
    MXE_TRY(cid.alloc(n_keep + 1, st)); MXE_TRY(cg.alloc(n_keep + 1, st)); MXE_TRY(cloc.alloc(n_keep + 1, st)); MXE_TRY(chash.alloc(n_keep + 1, st));
    MXE_TRY(eflag.alloc(n_keep, st)); MXE_TRY(eprefix.alloc(n_keep + 1, st));
    MXE_LAUNCH(e, a2a_compact_kernel, gridf(L), 256, 0, lmark.p, X->lhash, S, L, kflag.p, kprefix.p, G, world, cid.p, chash.p, cg.p, cloc.p);
    MXE_LAUNCH(e, local_pair_flag_kernel, gridf(n_keep), 256, 0, cloc.p, n_keep, S, lcontig.p, eflag.p);
    MXE_TRY(exclusive_scan_u32_u64(e, eflag.p, eprefix.p, n_keep));
    // at most n_keep - 1 sightings: size the record buffers for the worst case instead of waiting for the count
    const uint64_t cap = 2 * (n_keep - 1);
    MXE_TRY(rec.alloc(3 * cap, st)); MXE_TRY(dkey.alloc(cap, st)); MXE_TRY(dkey2.alloc(cap, st)); MXE_TRY(dval.alloc(cap, st)); MXE_TRY(dval2.alloc(cap, st));
    MXE_TRY(cnt.alloc(world, st));
    MXE_CUDA(cudaMemsetAsync(cnt.p, 0, world * sizeof(unsigned long long), st));
    MXE_LAUNCH(e, a2a_records_kernel, gridf(n_keep), 256, 0, cid.p, chash.p, cg.p, cloc.p, eflag.p, eprefix.p, n_keep, S, rec.p, dkey.p, dval.p, cnt.p);
    uint64_t n_sight = 0;
    MXE_CUDA(cudaMemcpyAsync(&n_sight, eprefix.p + n_keep, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaMemcpyAsync(rec_counts, cnt.p, world * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));
    const uint64_t n_rec = 2 * n_sight;
    if (n_rec == 0) return MXE_OK;
    MXE_TRY(X->alloc(&X->send_rec, 3 * n_rec));
    MXE_TRY(radix_sort_pairs(e, dkey.p, dval.p, dkey2.p, dval2.p, n_rec, 0, 8));       // group by destination
    MXE_LAUNCH(e, gather_rec_kernel, gridf(n_rec), 256, 0, rec.p, dval.p, n_rec, X->send_rec);
    *d_send_records = X->send_rec;
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

int a2a_finish_impl(mxe_a2a* X, const uint64_t* d_rec, uint64_t n_rec, uint64_t N_global, const double* weights, mxe_result* R)
{
    mxe_engine* e = X->eng;
    cudaStream_t st = e->stream;
    Span whole(e, "filter");
    Span part(e, "a2a_finish");
    const int n_asm = X->n_asm;
    AsmOffsets A = X->A;
    A.n = n_asm;
    for (int a = 0; a < n_asm; a++) A.weight[a] = weights[a];
    R->eng = e; R->n_asm = n_asm; R->N = X->L; R->nV = X->nV_local; R->nE = 0;
    for (int a = 0; a <= n_asm; a++) R->asm_off[a] = X->S.lofs[a];
    auto take = [&](void* p) { for (auto& q : X->owned) if (q == p) q = nullptr; return p; };
    R->d_uniq = (uint8_t*)take(X->luniq); R->d_keep = (uint8_t*)take(X->lkeep);
    R->d_vertices = (uint64_t*)take(X->vertices);
    const uint64_t nV = X->nV_local;
    if (n_rec == 0 || nV == 0) return MXE_OK;
    if (N_global >= (1ULL << 32)) { set_error("too many minimizers"); return MXE_ERR_ARG; }
    DBuf<uint32_t> succ, pred, own, mask_i, srcmin;
    DBuf<uint64_t> uprefix;
    MXE_TRY(succ.alloc((uint64_t)n_asm * nV, st)); MXE_TRY(pred.alloc((uint64_t)n_asm * nV, st)); MXE_TRY(srcmin.alloc(nV, st));
    MXE_TRY(own.alloc(n_rec, st)); MXE_TRY(mask_i.alloc(n_rec, st)); MXE_TRY(uprefix.alloc(n_rec + 1, st));
    MXE_CUDA(cudaMemsetAsync(succ.p, 0, (uint64_t)n_asm * nV * sizeof(uint32_t), st));
    MXE_CUDA(cudaMemsetAsync(pred.p, 0, (uint64_t)n_asm * nV * sizeof(uint32_t), st));
    MXE_CUDA(cudaMemsetAsync(srcmin.p, 0xFF, nV * sizeof(uint32_t), st));
    MXE_LAUNCH(e, a2a_table_kernel, gridf(n_rec), 256, 0, d_rec, n_rec, nV, succ.p, pred.p);
    MXE_LAUNCH(e, a2a_edge_owner_kernel, gridf(n_rec), 256, 0, d_rec, n_rec, nV, n_asm, succ.p, pred.p, own.p, mask_i.p);
    MXE_TRY(exclusive_scan_u32_u64(e, own.p, uprefix.p, n_rec));
    uint64_t nE = 0;
    MXE_CUDA(cudaMemcpyAsync(&nE, uprefix.p + n_rec, 8, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaStreamSynchronize(st));
    R->nE = nE;
    if (nE == 0) return MXE_OK;
    DBuf<uint32_t> e_rec, e_mask, oval, oval2, emask;
    DBuf<uint64_t> okey, okey2, eu, ev, ekey;
    DBuf<double> ew;
    MXE_TRY(e_rec.alloc(nE, st)); MXE_TRY(e_mask.alloc(nE, st)); MXE_TRY(oval.alloc(nE, st)); MXE_TRY(oval2.alloc(nE, st));
    MXE_TRY(okey.alloc(nE, st)); MXE_TRY(okey2.alloc(nE, st));
    MXE_TRY(eu.alloc(nE, st)); MXE_TRY(ev.alloc(nE, st)); MXE_TRY(emask.alloc(nE, st)); MXE_TRY(ew.alloc(nE, st)); MXE_TRY(ekey.alloc(nE, st));
    MXE_LAUNCH(e, a2a_edge_compact_kernel, gridf(n_rec), 256, 0, d_rec, n_rec, own.p, uprefix.p, mask_i.p, e_rec.p, e_mask.p, srcmin.p);
    MXE_LAUNCH(e, a2a_edge_key_kernel, gridf(nE), 256, 0, d_rec, e_rec.p, nE, srcmin.p, okey.p, oval.p);
    int kb = 1;
    while (kb < 32 && (1ULL << kb) <= N_global) kb++;
    kb = ((kb + 5 + 7) / 8) * 8;
    MXE_TRY(radix_sort_pairs(e, okey.p, oval.p, okey2.p, oval2.p, nE, 0, kb));
    MXE_LAUNCH(e, a2a_edge_gather_kernel, gridf(nE), 256, 0, oval.p, okey.p, nE, d_rec, e_rec.p, e_mask.p, R->d_vertices, A,
               eu.p, ev.p, emask.p, ew.p, ekey.p);
    R->d_eu = eu.detach(); R->d_ev = ev.detach(); R->d_emask = emask.detach(); R->d_ew = ew.detach(); R->d_ekey = ekey.detach();
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

}  // namespace mxe
