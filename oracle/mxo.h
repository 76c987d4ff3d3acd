/*
 * mxo.h -- CPU ORACLE for the ntJoin step-1/2/3 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call into this library.  The product (ntjoin_b200/) never links or loads it.
 *
 * What it restates
 *   step 1  btllib `indexlr` (ntHash + windowed minimizers).  btllib is a third-party,
 *           un-vendored, UNPINNED dependency of the reference (requirements.txt:4); it is
 *           not under /root/reference.  The algorithm is restated from its published
 *           description and anchored on the reference's call sites
 *           ntJoin:204-205, bin/ntjoin_utils.py:195-202 and golden files
 *           tests/expected_outputs/{ref.fa,scaf.f-f.fa}.k32.w1000.tsv.
 *   step 2  bin/ntjoin_utils.py:167-193 read_minimizers (per-assembly uniqueness)
 *           bin/ntjoin_utils.py:152-165 filter_minimizers (found-in-all intersection)
 *   step 3  bin/ntjoin_utils.py:94-115  build_graph edge stage, :54-56 calc_total_weight
 *
 * Parity status (see DESIGN.md "Oracle pinning")
 *   PINNED   byte-for-byte by the two golden TSVs under canonical=MIN (legacy ntHash1):
 *            seeds, split-rotate, rolling, multi-hash constants, window width, emission
 *            rule, decimal text format.
 *   PINNED   by tests/ntjoin_test.py coordinates under canonical=SUM (k=32,w=500 positions;
 *            k=15,w=10 overlap cut points), and by golden vectors generated from the
 *            reference's own Python steps 2-3 (tests/golden/make_golden.py).
 *   UNPINNED ("parity unpinned" for these items only): out_hash VALUES under SUM,
 *            tie-break direction, N-skip vs N-slot, lower-case handling.  They follow the
 *            documented btllib behaviour and sit behind the switches below.
 */
#ifndef MXO_H
#define MXO_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { MXO_CANON_SUM = 0, MXO_CANON_MIN = 1 };   /* hash0 = fwd+rev (ntHash2) | min(fwd,rev) (ntHash1) */
enum { MXO_TIE_RIGHT = 0, MXO_TIE_LEFT = 1 };    /* equal hash0 inside a window: rightmost | leftmost */

typedef struct {
    uint64_t out_hash;  /* hash1: printed identity of the minimizer          */
    uint64_t min_hash;  /* hash0: selection key                               */
    uint64_t pos;       /* 0-based offset of the k-mer in the record          */
    uint32_t contig;    /* record index in input order                        */
    uint32_t forward;   /* fwd <= rev                                         */
} mxo_mx_t;

/* Base hashes of one k-mer (no validity check; caller guarantees ACGTacgt). */
void mxo_kmer_hashes(const char* kmer, unsigned k, int canonical,
                     uint64_t* fwd, uint64_t* rev, uint64_t* hash0, uint64_t* hash1);

/* Windowed minimizers of ONE record.  Appends to *out (realloc'd); returns count appended,
 * or (size_t)-1 on allocation failure. */
size_t mxo_minimize(const char* seq, size_t len, unsigned k, unsigned w, int canonical, int tie,
                    uint32_t contig_idx, mxo_mx_t** out, size_t* n_out, size_t* cap_out);

/* Whole assembly held in memory: concatenated sequence + offsets (n_contigs+1).
 * threads>1 processes records in parallel (one record per worker, like indexlr -t). */
int mxo_sketch_buffers(const char* seq, const uint64_t* offsets, uint32_t n_contigs,
                       unsigned k, unsigned w, int canonical, int tie, int threads,
                       mxo_mx_t** out, size_t* n_out);

/* FASTA (multi-line, '>' records; id = header up to first whitespace; sequence upper-cased). */
typedef struct {
    char*     seq;       /* concatenated, no separators */
    uint64_t* offsets;   /* n_contigs + 1               */
    char**    names;
    uint32_t  n_contigs;
} mxo_fasta_t;
int  mxo_read_fasta(const char* path, mxo_fasta_t* out);
void mxo_free_fasta(mxo_fasta_t* f);

/* TSV as `indexlr [--pos] [--strand] [--seq]` prints it. path "-" = stdout. */
int mxo_write_tsv(const char* path, const mxo_fasta_t* fa, const mxo_mx_t* mx, size_t n,
                  unsigned k, int with_pos, int with_strand, int with_seq);

/* Steps 2-3 on arrays.  Assemblies in order refs..., target.
 * In:  per assembly a: hashes[a][0..n[a]) in (contig,pos) order, contig[a][i] record index.
 * Out: keep[a][i] = 1 iff hash survives uniqueness + intersection;  uniq[a][i] = 1 iff unique
 *      within its assembly; edges in the order bin/ntjoin_utils.py:115 would list them. */
typedef struct {
    uint64_t u, v;          /* first-seen orientation                 */
    uint32_t support_mask;  /* bit a set: assembly a supports edge    */
    double   weight;        /* sum of weights in assembly order       */
} mxo_edge_t;
int mxo_filter_and_edges(int n_asm, const uint64_t* const* hashes, const uint32_t* const* contig,
                         const size_t* n, const double* weights,
                         uint8_t** uniq, uint8_t** keep,
                         mxo_edge_t** edges, size_t* n_edges,
                         uint64_t** vertices, size_t* n_vertices);

void mxo_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
