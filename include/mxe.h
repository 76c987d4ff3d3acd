/*
 * mxe.h -- C ABI of the B200 minimizer sketch-and-filter engine (libmxe.so).
 *
 * This is the drop-in boundary for ntJoin's step-1/2/3 hot path.  The reference has no FFI of
 * its own for this path; it crosses a process seam and three Python functions.  Each entry
 * point below names the reference interface it replaces (paths relative to the ntJoin repo):
 *
 *   mxe_sketch_file / mxe_sketch_buffers / mxe_write_tsv
 *        replace the `indexlr --seq --long --pos -k K -w W -t T FASTA > FASTA.kK.wW.tsv`
 *        subprocess (ntJoin:204-205, bin/ntjoin_utils.py:195-202) and btllib.Indexlr
 *        (bin/ntjoin_assemble.py:478-481).
 *   mxe_filter_and_edges
 *        replaces read_minimizers' uniqueness filter (bin/ntjoin_utils.py:167-193),
 *        filter_minimizers (bin/ntjoin_utils.py:152-165) and the edge stage of build_graph
 *        (bin/ntjoin_utils.py:94-115) with calc_total_weight (:54-56).
 *
 * Conventions: plain C, every call returns 0 on success or a negative code; the message is in
 * mxe_last_error() (thread local).  Objects returned by the engine are owned by the engine and
 * released with the matching *_free.  Calls on one engine handle are not re-entrant.  There is
 * no CPU fallback: without a CUDA device mxe_create fails.
 */
#ifndef MXE_H
#define MXE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mxe_engine mxe_t;
typedef struct mxe_sketch mxe_sketch_t;
typedef struct mxe_result mxe_result_t;
typedef struct mxe_dist mxe_dist_t;
typedef struct mxe_a2a mxe_a2a_t;
typedef struct mxe_p2p mxe_p2p_t;
typedef struct mxe_fasta mxe_fasta_t;

enum {
    MXE_OK = 0,
    MXE_ERR_ARG = -1,      /* bad argument                         */
    MXE_ERR_IO = -2,       /* file could not be read / written     */
    MXE_ERR_CUDA = -3,     /* CUDA runtime error                   */
    MXE_ERR_NOMEM = -4,
    MXE_ERR_INTERNAL = -5
};

/* flags for the sketch calls */
enum {
    MXE_CANON_SUM = 0,     /* hash0 = fwd + rev   (current btllib / ntHash2; default)           */
    MXE_CANON_MIN = 1,     /* hash0 = min(fwd,rev) (legacy ntHash1; matches the shipped goldens) */
    MXE_KEEP_DEVICE = 2    /* keep the minimizer arrays resident on the device for              */
                           /* mxe_filter_and_edges / mxe_sketch_device_view                     */
};

const char* mxe_version(void);
const char* mxe_last_error(void);

/* One engine per process and device.  device = CUDA ordinal. */
int  mxe_create(int device, mxe_t** out);
void mxe_destroy(mxe_t* e);

/* Run all engine work on a caller-owned stream (cudaStream_t passed as a pointer-sized integer),
 * e.g. torch.cuda.current_stream().cuda_stream.  NULL restores the engine's own stream. */
int  mxe_set_stream(mxe_t* e, void* cuda_stream);

/* Tunables (name = "tau" candidate threshold multiplier, "chunk" positions per thread, ...). */
int  mxe_set_option(mxe_t* e, const char* name, double value);

/* ---- step 1: ordered minimizer sketch --------------------------------------------------- */

/* FASTA / FASTQ file (multi-line records, '>' headers; id = header up to first whitespace); gzip, bzip2, xz and zstd
 * files (recognised by their first bytes) are read through a pipe from `gzip -dc` etc., as btllib does.  Records of
 * 2^32 bases or more are rejected (MXE_ERR_ARG): positions are 32-bit in the TSV arrays.
 * Replaces: indexlr subprocess, ntJoin:204-205. */
int mxe_sketch_file(mxe_t* e, const char* fasta_path, int k, int w, int flags, mxe_sketch_t** out);

/* The same reader on its own, host only (no engine, no GPU): what btllib.SeqReader gives ntJoin
 * (bin/ntjoin_assemble.py:313-316).  FASTA or FASTQ; record id = header up to the first blank; sequence lines
 * joined and upper-cased.  `seq` holds all records back to back (offsets has n_records + 1 entries). */
int mxe_fasta_read(const char* path, mxe_fasta_t** out);
int mxe_fasta_view(mxe_fasta_t* f, uint32_t* n_records, const uint64_t** offsets, const char** seq);
int mxe_fasta_name(mxe_fasta_t* f, uint32_t idx, const char** name);
void mxe_fasta_free(mxe_fasta_t* f);

/* Host buffers: `seq` = all records concatenated (ASCII, either case, no separators),
 * `offsets` = n_contigs+1 starts, `names` may be NULL.  Host->device copy is inside the call. */
int mxe_sketch_buffers(mxe_t* e, const uint8_t* seq, const uint64_t* offsets, uint32_t n_contigs,
                       const char* const* names, int k, int w, int flags, mxe_sketch_t** out);

/* Optional: start the host->device copy of a buffer that a later mxe_sketch_buffers call (same
 * pointer and length) will sketch.  Returns at once; the copy runs on the engine's copy streams
 * while earlier work is still computing (at most two buffers in flight; further calls are no-ops).
 * Like any asynchronous copy: the buffer must stay allocated and UNCHANGED until the sketch call that consumes it has
 * returned -- the staged copy is recognised by (pointer, length) alone, a buffer rewritten or recycled in between would
 * be sketched in its old state.  A prefetch that is never consumed is dropped when its slot is needed again. */
int mxe_prefetch_buffers(mxe_t* e, const uint8_t* seq, uint64_t n);

/* Same, with `d_seq` already resident in device memory of this engine's device (16-byte aligned).
 * `offsets` stays a host array.  `stream` = cudaStream_t as an integer (0 = engine stream). */
int mxe_sketch_device(mxe_t* e, const void* d_seq, const uint64_t* offsets, uint32_t n_contigs,
                      const char* const* names, int k, int w, int flags, mxe_sketch_t** out);

/* Several device-resident assemblies (d_seq[a], offsets[a] with n_contigs[a] + 1 entries) in one call: their kernels are
 * enqueued on two streams and run concurrently -- the bandwidth- and latency-bound kernels of one assembly share the SMs
 * with the ALU-bound candidate scan of the other -- and the sizes are read in one host round trip per assembly at the end.
 * Same results as n_asm calls of mxe_sketch_device; out[a] are ordinary sketch objects (names = record indices). */
int mxe_sketch_device_many(mxe_t* e, int n_asm, const void* const* d_seq, const uint64_t* const* offsets, const uint32_t* n_contigs,
                           int k, int w, int flags, mxe_sketch_t** out);

/* Host only (no engine): a sketch object over caller-provided arrays (copied) -- e.g. the minimizers of one assembly
 * gathered from several GPUs -- so that mxe_write_tsv / mxe_sketch_view serve them like a sketch computed here.
 * contig[] ascending; min_hash / forward may be NULL; offsets (n_contigs + 1 record starts in `seq`) and `seq`
 * (borrowed, must outlive the object) are only needed for --seq output. */
int mxe_sketch_from_arrays(const uint64_t* out_hash, const uint64_t* min_hash, const uint32_t* pos, const uint32_t* contig,
                           const uint8_t* forward, uint64_t n, const char* const* names, const uint64_t* offsets,
                           uint32_t n_contigs, int k, const uint8_t* seq, mxe_sketch_t** out);

/* Load a sketch back from an `indexlr` TSV written earlier (make's resume path: ntJoin keeps
 * <fasta>.k<k>.w<w>.tsv as .SECONDARY, ntJoin:202).  Parses id \t hash[:pos[:...]] ... exactly as
 * bin/ntjoin_utils.py:173-185 does; the arrays are uploaded for mxe_filter_and_edges. */
int mxe_sketch_load_tsv(mxe_t* e, const char* tsv_path, mxe_sketch_t** out);

/* Host view (SoA, sorted by (contig, pos); contigs in input order).  Pointers stay valid until
 * mxe_sketch_free.  Any output pointer may be NULL. */
int mxe_sketch_view(mxe_sketch_t* s, uint64_t* n,
                    const uint64_t** out_hash,   /* hash1: the minimizer's identity in ntJoin   */
                    const uint64_t** min_hash,   /* hash0: window selection key                 */
                    const uint32_t** pos,        /* 0-based offset in the record                */
                    const uint32_t** contig,     /* record index                                */
                    const uint8_t**  forward);   /* fwd <= rev                                  */

/* Start the device->host copy behind mxe_sketch_view now, on a copy stream of the engine, and return at once: the copy
 * runs beside the next assembly's host->device copy and sketch (PCIe is full duplex) instead of after them.  A later
 * mxe_sketch_view only waits for it.  No-op for host-only sketches and when the host copy already exists. */
int mxe_sketch_prefetch_host(mxe_sketch_t* s);

/* Device view (valid only with MXE_KEEP_DEVICE): raw device pointers of the same SoA. */
int mxe_sketch_device_view(mxe_sketch_t* s, uint64_t* n, const void** d_out_hash,
                           const void** d_pos, const void** d_contig);

/* Index (in the SoA above) of the first minimizer whose record id is >= `record`: where the minimizers of a record
 * range start.  Several assemblies can be sketched in ONE call as one list of records (windows never cross records, so
 * the result is the same as sketching them separately) and split afterwards at the record where each assembly starts. */
int mxe_sketch_record_start(mxe_sketch_t* s, uint32_t record, uint64_t* index);

/* Record name (header up to the first whitespace) of record `idx`; valid until mxe_sketch_free. */
int mxe_sketch_contig_name(mxe_sketch_t* s, uint32_t idx, const char** name);

int mxe_sketch_counts(mxe_sketch_t* s, uint64_t* n_bases, uint64_t* n_valid_kmers,
                      uint64_t* n_candidates, uint64_t* n_gap_windows, uint32_t* n_contigs);

/* `indexlr` text output: id \t hash[:pos][:strand][:seq] ...   path "-" = stdout.
 * with_seq needs the sequence: available after mxe_sketch_file / mxe_sketch_buffers only. */
int mxe_write_tsv(mxe_sketch_t* s, const char* path, int with_pos, int with_strand, int with_seq);

void mxe_sketch_free(mxe_sketch_t* s);

/* ---- steps 2-3: uniqueness, found-in-all intersection, adjacent-pair edge list ----------- */

/* Sketches in assembly order (references in CLI order, target LAST: bin/ntjoin.py:181-185,
 * bin/ntjoin_assemble.py:804-807).  weights[a] as passed to build_graph. n_asm <= 32. */
int mxe_filter_and_edges(mxe_t* e, mxe_sketch_t* const* sketches, int n_asm, const double* weights,
                         mxe_result_t** out);

/* Same on raw device arrays (used by the multi-GPU path after the NCCL all-gather):
 * d_hash[a] -> n[a] uint64 out_hash in (contig,pos) order; d_contig[a] -> n[a] uint32. */
int mxe_filter_and_edges_device(mxe_t* e, const void* const* d_hash, const void* const* d_contig,
                                const uint64_t* n, int n_asm, const double* weights,
                                mxe_result_t** out);

/* Sizes only (no device->host copy of the arrays). */
int mxe_result_counts(mxe_result_t* r, uint64_t* n_minimizers, uint64_t* n_vertices, uint64_t* n_edges);

/* Per assembly: flags over its minimizers in sketch order.
 * uniq[i]=1: out_hash occurs once in the assembly (survives read_minimizers);
 * keep[i]=1: additionally found (unique) in every assembly (survives filter_minimizers). */
int mxe_result_flags(mxe_result_t* r, int asm_idx, uint64_t* n, const uint8_t** uniq, const uint8_t** keep);

/* Graph: vertices = surviving out_hash values (ascending); edges in the order
 * bin/ntjoin_utils.py:115 lists them, (u,v) in first-seen orientation, support_mask bit a =
 * assembly a, weight = sum of weights in assembly order (IEEE double, bit-identical to Python). */
int mxe_result_graph(mxe_result_t* r, uint64_t* n_vertices, const uint64_t** vertices,
                     uint64_t* n_edges, const uint64_t** edge_u, const uint64_t** edge_v,
                     const uint32_t** support_mask, const double** weight);

void mxe_result_free(mxe_result_t* r);

/* `<prefix>.mx.dot` exactly as Ntjoin.print_graph writes it (bin/ntjoin.py:25-67), from arrays instead of a
 * walk over igraph objects.  Host only (no engine, no GPU).  vertices: n_v hashes in graph.vs order; the label
 * of vertex i carries one line "<asm_keys[a]>_(<ctg_reprs[a][v_ctg[a][i]]>, <v_pos[a][i]>)" per assembly (the
 * caller passes Python's repr() of every record name); edge t joins vertices[e_src[t]] -- vertices[e_dst[t]]
 * (indices, as igraph reports source/target) and ends with attr_text[e_attr[t]], e.g.
 * " [weight=3.0 color=lightgrey]\n" (one text per distinct support set: float formatting stays Python's). */
int mxe_write_dot(const char* path, uint64_t n_v, const uint64_t* vertices, int n_asm, const char* const* asm_keys,
                  const char* const* const* ctg_reprs, const uint32_t* const* v_ctg, const uint32_t* const* v_pos,
                  uint64_t n_e, const uint32_t* e_src, const uint32_t* e_dst, const uint32_t* e_attr,
                  const char* const* attr_text);

/* ---- steps 2-3 across GPUs (one process per GPU) ------------------------------------------
 *
 * Uniqueness is per ASSEMBLY, not per GPU (bin/ntjoin_utils.py:182-187), and the intersection is
 * over all assemblies (:155-157), so the path has real exchange steps.  The collectives belong to
 * the caller (torch.distributed / NCCL over NVLink); these four stages run between them on
 * GLOBALLY indexed device arrays, each rank doing 1/world of the single-GPU work:
 *
 *   all-gather(hashes)   -> mxe_dist_mark       (owned hash range: unique / found-in-all / vertex ids)
 *   all-reduce(sum, mk)  -> mxe_dist_adjacency  (own records: survivors, adjacent pairs -> successor table)
 *   all-reduce(sum, suc) -> mxe_dist_edges      (support masks, edge ownership, first-source table)
 *   all-reduce(min, src) -> mxe_dist_finish     (local edge shard with global order keys)
 *
 * Global index space: assemblies in order (references ..., target last); inside an assembly the
 * ranks in order (rank r holds a contiguous range of records); N = asm_off[n_asm] < 0x7f000000.
 * Rank r owns the hashes h with ((h >> 32) * world) >> 32 == r.
 */

/* d_keys: N uint64 out_hash in global order (must stay alive until mxe_dist_finish).
 * d_mk (out): N uint32, zero outside the owned entries; bit 31 = unique in its assembly
 * (read_minimizers), bits 30:0 = 1 + local vertex id if found unique in every assembly. */
int mxe_dist_mark(mxe_t* e, const void* d_keys, const uint64_t* asm_off, int n_asm, int rank, int world,
                  void* d_mk, mxe_dist_t** out, uint64_t* n_vertices_local);

/* d_mk: summed over ranks.  vbase: world+1 exclusive prefix of the per-rank vertex counts.
 * loc_off[a] / loc_n[a]: this rank's slice of assembly a in the global index space;
 * d_contig[a]: its loc_n[a] uint32 record ids.  d_succ (out): n_asm * nV uint32, entry [a * nV + v] =
 * 1 + global vertex id of the successor of vertex v in assembly a, zero outside this rank's
 * sightings (so the tables of all ranks combine by summation; the predecessor of v is x exactly
 * when the successor of x is v, so one table serves both directions). */
int mxe_dist_adjacency(mxe_dist_t* d, const void* d_mk, const uint64_t* vbase, const uint64_t* loc_off,
                       const uint64_t* loc_n, const void* const* d_contig, void* d_succ);

/* d_succ: summed over ranks.  d_srcmin (out): nV uint32, for every vertex the creation
 * index of the first edge this rank owns with that vertex as source (0x7f7f7f7f = none). */
int mxe_dist_edges(mxe_dist_t* d, const void* d_succ, void* d_srcmin, uint64_t* n_edges_local);

/* d_srcmin: minimum over ranks.  The result is this rank's SHARD: flags of its own slices
 * (mxe_result_flags), its vertices (owned hash range, ascending), its edges ordered by
 * mxe_result_edge_keys; concatenating flags / vertices in rank order and merging the edge shards
 * by key gives exactly the single-GPU result. */
int mxe_dist_finish(mxe_dist_t* d, const void* d_srcmin, const double* weights, mxe_result_t** out);
void mxe_dist_free(mxe_dist_t* d);

/* ---- the same across GPUs with all-to-all exchanges (work and traffic ~ 1/world) -------------
 *
 * Every item travels once, to the rank that needs it (NVSwitch: uniform all-to-all bandwidth):
 *
 *   mxe_a2a_partition  own minimizers grouped by hash owner      -> all-to-all (keys, 8 B)
 *   mxe_a2a_mark       owner: unique / found-in-all / vertex ids  -> all-to-all (marks, 4 B, same order back)
 *   mxe_a2a_sightings  source: ordered survivors, adjacent pairs; every sighting goes to the owner
 *                      of its source vertex (successor record) and of its target (predecessor)
 *                                                                 -> all-to-all (records, 24 B)
 *   mxe_a2a_finish     owner: succ/pred tables of its own vertices, support masks, edge ownership,
 *                      order keys.  Result shard: flags of this rank's minimizers, vertices of its
 *                      hash range, the edges whose source vertex it owns (mxe_result_edge_keys).
 *
 * Buffers returned through d_send_* belong to the handle (valid until mxe_a2a_free / mxe_a2a_finish);
 * receive buffers are the caller's.  world <= 16.
 */

/* d_hash[a], n[a]: this rank's out_hash lists.  counts (out, host): world * n_asm, entry [o * n_asm + a] =
 * minimizers of assembly a sent to owner o; *d_send_keys: sum(n) uint64 grouped by owner, assembly order inside. */
int mxe_a2a_partition(mxe_t* e, const void* const* d_hash, const uint64_t* n, int n_asm, int rank, int world,
                      mxe_a2a_t** out, uint64_t* counts, const void** d_send_keys);

/* d_recv_keys: keys received, grouped by source rank (assembly order inside); recv_counts[r * n_asm + a].
 * d_ret_marks (out, caller buffer of as many uint32): bit 31 = unique in its assembly, bits 30:0 = 1 + vertex id
 * local to this owner if found unique in every assembly; same order as d_recv_keys. */
int mxe_a2a_mark(mxe_a2a_t* x, const void* d_recv_keys, const uint64_t* recv_counts, void* d_ret_marks,
                 uint64_t* n_vertices_local);

/* d_marks: the marks of this rank's minimizers, in the order of d_send_keys.  d_contig[a]: record ids;
 * goff[a]: global index of this rank's first minimizer of assembly a.  rec_counts (out, host): world entries;
 * *d_send_records: 3 uint64 per record, grouped by destination. */
int mxe_a2a_sightings(mxe_a2a_t* x, const void* d_marks, const void* const* d_contig, const uint64_t* goff,
                      uint64_t* rec_counts, const void** d_send_records);

/* d_recv_records: n_records records addressed to this owner (any order).  n_global = all minimizers of all
 * assemblies and ranks.  Frees the handle's send buffers; the handle itself is released by mxe_a2a_free. */
int mxe_a2a_finish(mxe_a2a_t* x, const void* d_recv_records, uint64_t n_records, uint64_t n_global,
                   const double* weights, mxe_result_t** out);
void mxe_a2a_free(mxe_a2a_t* x);

/* ---- steps 2-3 across GPUs, exchanges as direct peer stores over NVLink (the default formulation) ----------------
 *
 * Same replaced reference code and same result shards as mxe_dist_* / mxe_a2a_* (bin/ntjoin_utils.py:182-192, :155-162,
 * :94-115, :54-56), no collective library: every rank owns a SYMMETRIC device workspace that its peers map through
 * CUDA IPC; minimizers are partitioned into hash buckets owned by ranks (bucket b -> rank (b * world) >> B, monotone
 * in the hash), the producing kernels store into the consumer's workspace and a device-side barrier separates the
 * five stages (csrc/p2p.cu).  One host round trip per call (sizes, in mxe_p2p_finish).  world = 1 is the single-GPU
 * path of mxe_filter_and_edges.  world <= 16; capacity < 2^32 minimizers.
 *
 *   h = mxe_p2p_create(e, rank, world, cap_total, n_asm_max)     cap_total: upper bound of all minimizers of all
 *                                                                assemblies and ranks (sizes the workspace)
 *   mxe_p2p_handle(h, handle64)  -> exchange the 64-byte handles (e.g. torch.distributed.all_gather_object)
 *   mxe_p2p_connect(h, handles)                                  one process per GPU
 *   mxe_p2p_connect_pointers(h, bases)                           several ranks in one process (tests), from mxe_p2p_workspace
 *   per job, on every rank, in this order (each call only enqueues work on the engine stream):
 *   mxe_p2p_scatter(h, d_hash, d_contig, n, n_asm, weights)      own minimizers -> bucket owners
 *   mxe_p2p_buckets(h)                                           owned buckets: uniqueness, found-in-all, vertex ranks
 *   mxe_p2p_adjacency(h)                                         own records: survivors, adjacent pairs -> successor tables
 *   mxe_p2p_edges(h)                                             support masks, edge ownership, first-source tables
 *   mxe_p2p_finish(h, &result)                                   world 1: full result in the reference's order;
 *                                                                world > 1: shard (own flags, owned vertices, own edges
 *                                                                in creation order + mxe_result_edge_keys)
 * Repeated sequence (one hash thousands of times in one bucket): world > 1 places the buckets at exact offsets at the
 * owner and a bucket larger than a CTA's shared memory is reduced to two copies per (hash, assembly) while it is loaded,
 * so the call answers; world = 1 uses fixed bucket slots, an overflow there makes mxe_p2p_finish fail with
 * MXE_ERR_INTERNAL and mxe_filter_and_edges falls back to the sort-based formulation by itself.  What remains an error
 * (MXE_ERR_INTERNAL from mxe_p2p_finish on the rank that saw it): more than ~500 / 1000 DISTINCT hashes in one bucket, or
 * a rank that receives more than its segment capacity (cap_total / world^2 * 1.3 + 8192 records per source rank).
 * Environment (read by mxe_p2p_create): MXE_P2P_RECORDS=0 keeps the successor tables at the vertex owners and accesses
 * them with fine-grained peer loads / stores; MXE_P2P_BUCKET_AVG / MXE_P2P_BKMAX size the buckets (default 400 records on
 * average at full capacity; 512 or 1024 records of shared memory per bucket CTA).
 */
int mxe_p2p_create(mxe_t* e, int rank, int world, uint64_t cap_total, int n_asm_max, mxe_p2p_t** out);
int mxe_p2p_handle(mxe_p2p_t* h, void* handle64, uint64_t* workspace_bytes);
int mxe_p2p_connect(mxe_p2p_t* h, const void* handles);
int mxe_p2p_workspace(mxe_p2p_t* h, void** ptr);
int mxe_p2p_connect_pointers(mxe_p2p_t* h, void* const* bases);
int mxe_p2p_scatter(mxe_p2p_t* h, const void* const* d_hash, const void* const* d_contig, const uint64_t* n, int n_asm,
                    const double* weights);
int mxe_p2p_buckets(mxe_p2p_t* h);
int mxe_p2p_adjacency(mxe_p2p_t* h);
int mxe_p2p_edges(mxe_p2p_t* h);
int mxe_p2p_finish(mxe_p2p_t* h, mxe_result_t** out);
void mxe_p2p_free(mxe_p2p_t* h);

/* Global order key of every edge of a multi-GPU shard (mxe_dist / mxe_a2a: ascending inside the shard; mxe_p2p: the
 * shard is in creation order -- merge by key either way). */
int mxe_result_edge_keys(mxe_result_t* r, uint64_t* n_edges, const uint64_t** keys);

/* ---- measurement hooks (bench.py) -------------------------------------------------------- */

/* Device time in milliseconds and launch count of the engine's kernels since the last reset,
 * measured with CUDA events on the engine stream.  name = "cand" (dominant sketch kernel),
 * "pack", "sketch" (whole step 1), "filter" (steps 2-3), "all". */
int mxe_timing(mxe_t* e, const char* name, double* ms, uint64_t* launches);
int mxe_timing_reset(mxe_t* e);
uint64_t mxe_kernel_launches(mxe_t* e);   /* total kernels launched by this engine */

#ifdef __cplusplus
}
#endif
#endif
