/*
 * mxo.c -- CPU ORACLE (test infrastructure; see mxo.h for the scope and pinning notes).
 *
 * Plain C restatement of the reference's step-1/2/3 path:
 *   step 1   btllib indexlr as invoked at ntJoin:204-205 / bin/ntjoin_utils.py:195-202
 *            (algorithm: SURVEY.md Appendix A.1-A.5; btllib itself is not in /root/reference)
 *   step 2   bin/ntjoin_utils.py:167-193 (read_minimizers), :152-165 (filter_minimizers)
 *   step 3   bin/ntjoin_utils.py:94-115 (build_graph edge stage), :54-56 (calc_total_weight)
 *
 * Deliberately written the "slow obvious" way (ring buffer with rescans, chained hash maps)
 * so that it shares no structure with the CUDA kernels it checks.
 */
#include "mxo.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ---------------------------------------------------------------- A.1 constants */
#define SEED_A 0x3c8bfbb395c60474ULL
#define SEED_C 0x3193c18562a02b4cULL
#define SEED_G 0x20323ed082572324ULL
#define SEED_T 0x295549f54be24456ULL
#define MULTISEED 0x90b45d39fb6da1faULL
#define MULTISHIFT 27

/* 0..3 = A,C,G,T (either case); 4 = anything else */
static int base_code(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}
static const uint64_t SEED[4] = { SEED_A, SEED_C, SEED_G, SEED_T };

/* ---------------------------------------------------------------- A.2 split rotate */
static uint64_t srol1(uint64_t x)
{
    uint64_t m = ((x & 0x8000000000000000ULL) >> 30) | ((x & 0x100000000ULL) >> 32);
    return ((x << 1) & 0xFFFFFFFDFFFFFFFFULL) | m;
}
static uint64_t sror1(uint64_t x)
{
    uint64_t m = ((x & 0x200000000ULL) << 30) | ((x & 1ULL) << 32);
    return ((x >> 1) & 0xFFFFFFFEFFFFFFFFULL) | m;
}
static uint64_t sroln(uint64_t x, unsigned n)
{
    while (n--) x = srol1(x);
    return x;
}

static uint64_t mix_hash1(uint64_t hash0, unsigned k)
{
    uint64_t t = hash0 * (1ULL ^ ((uint64_t)k * MULTISEED));
    return t ^ (t >> MULTISHIFT);
}

static uint64_t canonical_of(uint64_t fwd, uint64_t rev, int canonical)
{
    if (canonical == MXO_CANON_MIN) return rev < fwd ? rev : fwd;
    return fwd + rev;
}

static void kmer_base_hashes(const char* s, unsigned k, uint64_t* fwd, uint64_t* rev)
{
    uint64_t f = 0, r = 0;
    for (unsigned i = 0; i < k; i++) {
        f = srol1(f) ^ SEED[base_code((unsigned char)s[i])];
        r = srol1(r) ^ SEED[3 - base_code((unsigned char)s[k - 1 - i])];
    }
    *fwd = f;
    *rev = r;
}

void mxo_kmer_hashes(const char* kmer, unsigned k, int canonical,
                     uint64_t* fwd, uint64_t* rev, uint64_t* hash0, uint64_t* hash1)
{
    uint64_t f, r;
    kmer_base_hashes(kmer, k, &f, &r);
    uint64_t h0 = canonical_of(f, r, canonical);
    if (fwd) *fwd = f;
    if (rev) *rev = r;
    if (hash0) *hash0 = h0;
    if (hash1) *hash1 = mix_hash1(h0, k);
}

/* ---------------------------------------------------------------- A.3 / A.4 minimizers */
typedef struct { uint64_t h0, h1; uint64_t pos; uint32_t forward; } hk_t;

static int push_mx(mxo_mx_t** out, size_t* n, size_t* cap, const hk_t* e, uint32_t contig)
{
    if (*n == *cap) {
        size_t nc = *cap ? *cap * 2 : 1024;
        mxo_mx_t* p = (mxo_mx_t*)realloc(*out, nc * sizeof(mxo_mx_t));
        if (!p) return -1;
        *out = p;
        *cap = nc;
    }
    mxo_mx_t* m = &(*out)[(*n)++];
    m->out_hash = e->h1;
    m->min_hash = e->h0;
    m->pos = e->pos;
    m->contig = contig;
    m->forward = e->forward;
    return 0;
}

size_t mxo_minimize(const char* seq, size_t len, unsigned k, unsigned w, int canonical, int tie,
                    uint32_t contig_idx, mxo_mx_t** out, size_t* n_out, size_t* cap_out)
{
    size_t n_before = *n_out;
    if (k == 0 || w == 0 || (size_t)k > len || (size_t)w > len - k + 1) return 0;

    uint64_t rolk[4]; /* srol^k(seed[x]) */
    for (int x = 0; x < 4; x++) rolk[x] = sroln(SEED[x], k);

    hk_t* ring = (hk_t*)malloc((size_t)w * sizeof(hk_t));
    if (!ring) return (size_t)-1;

    uint64_t fwd = 0, rev = 0;
    size_t run = 0;           /* consecutive ACGT bytes ending at i          */
    int have_prev = 0;        /* k-mer starting at i-k was valid (can roll)  */
    size_t idx = 0;           /* ordinal of the next valid k-mer             */
    long long cur = -1;       /* ordinal of the current window minimum       */
    long long last_pos = -1;  /* position of the last emitted minimizer      */

    for (size_t i = 0; i < len; i++) {
        int c = base_code((unsigned char)seq[i]);
        if (c > 3) { run = 0; have_prev = 0; continue; }
        run++;
        if (run < k) continue;
        size_t p = i + 1 - k;
        if (have_prev) {
            int o = base_code((unsigned char)seq[p - 1]);
            fwd = srol1(fwd) ^ rolk[o] ^ SEED[c];
            rev = sror1(rev ^ rolk[3 - c] ^ SEED[3 - o]);
        } else {
            kmer_base_hashes(seq + p, k, &fwd, &rev);
            have_prev = 1;
        }
        hk_t* e = &ring[idx % w];
        e->h0 = canonical_of(fwd, rev, canonical);
        e->h1 = mix_hash1(e->h0, k);
        e->pos = p;
        e->forward = fwd <= rev;

        if (idx + 1 >= w) {
            long long left = (long long)(idx + 1 - w);
            if (cur < left) {
                /* minimum left the window: rescan all w entries */
                cur = left;
                for (long long j = left + 1; j <= (long long)idx; j++) {
                    uint64_t hj = ring[j % w].h0, hc = ring[cur % w].h0;
                    if (tie == MXO_TIE_LEFT ? hj < hc : hj <= hc) cur = j;
                }
            } else {
                uint64_t hn = e->h0, hc = ring[cur % w].h0;
                if (tie == MXO_TIE_LEFT ? hn < hc : hn <= hc) cur = (long long)idx;
            }
            const hk_t* m = &ring[cur % w];
            if ((long long)m->pos > last_pos && m->h0 != UINT64_MAX) {
                last_pos = (long long)m->pos;
                if (push_mx(out, n_out, cap_out, m, contig_idx)) { free(ring); return (size_t)-1; }
            }
        }
        idx++;
    }
    free(ring);
    return *n_out - n_before;
}

typedef struct {
    const char* seq; const uint64_t* offsets; uint32_t n_contigs;
    unsigned k, w; int canonical, tie;
    mxo_mx_t** part; size_t* pn;
    long long next; int err; pthread_mutex_t mu;
} sk_job_t;

static void* sk_worker(void* arg)
{
    sk_job_t* j = (sk_job_t*)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        long long c = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (c >= (long long)j->n_contigs) break;
        size_t cap = 0;
        size_t r = mxo_minimize(j->seq + j->offsets[c], (size_t)(j->offsets[c + 1] - j->offsets[c]),
                                j->k, j->w, j->canonical, j->tie, (uint32_t)c, &j->part[c], &j->pn[c], &cap);
        if (r == (size_t)-1) { pthread_mutex_lock(&j->mu); j->err = -ENOMEM; pthread_mutex_unlock(&j->mu); }
    }
    return NULL;
}

/* one record per worker, like `indexlr -t T` */
int mxo_sketch_buffers(const char* seq, const uint64_t* offsets, uint32_t n_contigs,
                       unsigned k, unsigned w, int canonical, int tie, int threads,
                       mxo_mx_t** out, size_t* n_out)
{
    mxo_mx_t** part = (mxo_mx_t**)calloc(n_contigs ? n_contigs : 1, sizeof(*part));
    size_t* pn = (size_t*)calloc(n_contigs ? n_contigs : 1, sizeof(size_t));
    if (!part || !pn) { free(part); free(pn); return -ENOMEM; }
    sk_job_t job = { seq, offsets, n_contigs, k, w, canonical, tie, part, pn, 0, 0, PTHREAD_MUTEX_INITIALIZER };
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((uint32_t)threads > n_contigs) threads = n_contigs ? (int)n_contigs : 1;
    pthread_t tid[256];
    int started = 0;
    for (int t = 1; t < threads; t++)
        if (pthread_create(&tid[started], NULL, sk_worker, &job) == 0) started++;
    sk_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
    int err = job.err;
    size_t total = 0;
    for (uint32_t c = 0; c < n_contigs; c++) total += pn[c];
    mxo_mx_t* all = (mxo_mx_t*)malloc((total ? total : 1) * sizeof(mxo_mx_t));
    if (!all) err = -ENOMEM;
    size_t at = 0;
    for (uint32_t c = 0; c < n_contigs; c++) {
        if (all && pn[c]) memcpy(all + at, part[c], pn[c] * sizeof(mxo_mx_t));
        at += pn[c];
        free(part[c]);
    }
    free(part);
    free(pn);
    if (err) { free(all); return err; }
    *out = all;
    *n_out = total;
    return 0;
}

/* ---------------------------------------------------------------- FASTA */
int mxo_read_fasta(const char* path, mxo_fasta_t* out)
{
    memset(out, 0, sizeof(*out));
    FILE* f = fopen(path, "rb");
    if (!f) return -errno;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* buf = (char*)malloc((size_t)sz + 1);
    if (!buf) { fclose(f); return -ENOMEM; }
    if (sz && fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); return -EIO; }
    fclose(f);
    buf[sz] = 0;

    size_t cap_c = 16;
    out->seq = (char*)malloc((size_t)sz + 1);
    out->offsets = (uint64_t*)malloc((cap_c + 1) * sizeof(uint64_t));
    out->names = (char**)malloc(cap_c * sizeof(char*));
    size_t at = 0;
    long i = 0;
    while (i < sz) {
        if (buf[i] == '>') {
            long s = i + 1, e = s;
            while (e < sz && buf[e] != '\n' && buf[e] != '\r' && buf[e] != ' ' && buf[e] != '\t') e++;
            if (out->n_contigs == cap_c) {
                cap_c *= 2;
                out->offsets = (uint64_t*)realloc(out->offsets, (cap_c + 1) * sizeof(uint64_t));
                out->names = (char**)realloc(out->names, cap_c * sizeof(char*));
            }
            char* nm = (char*)malloc((size_t)(e - s) + 1);
            memcpy(nm, buf + s, (size_t)(e - s));
            nm[e - s] = 0;
            out->names[out->n_contigs] = nm;
            out->offsets[out->n_contigs] = at;
            out->n_contigs++;
            while (i < sz && buf[i] != '\n') i++;
            i++;
        } else {
            while (i < sz && buf[i] != '\n') {
                char c = buf[i++];
                if (c == '\r') continue;
                if (c >= 'a' && c <= 'z') c = (char)(c - 32);
                if (out->n_contigs) out->seq[at++] = c;
            }
            i++;
        }
    }
    out->offsets[out->n_contigs] = at;
    free(buf);
    return 0;
}

void mxo_free_fasta(mxo_fasta_t* f)
{
    if (!f) return;
    for (uint32_t i = 0; i < f->n_contigs; i++) free(f->names[i]);
    free(f->names);
    free(f->offsets);
    free(f->seq);
    memset(f, 0, sizeof(*f));
}

/* ---------------------------------------------------------------- A.5 text format */
int mxo_write_tsv(const char* path, const mxo_fasta_t* fa, const mxo_mx_t* mx, size_t n,
                  unsigned k, int with_pos, int with_strand, int with_seq)
{
    FILE* f = (path[0] == '-' && path[1] == 0) ? stdout : fopen(path, "wb");
    if (!f) return -errno;
    size_t i = 0;
    for (uint32_t c = 0; c < fa->n_contigs; c++) {
        fputs(fa->names[c], f);
        fputc('\t', f);
        int first = 1;
        while (i < n && mx[i].contig == c) {
            if (!first) fputc(' ', f);
            first = 0;
            fprintf(f, "%llu", (unsigned long long)mx[i].out_hash);
            if (with_pos) fprintf(f, ":%llu", (unsigned long long)mx[i].pos);
            if (with_strand) fprintf(f, ":%c", mx[i].forward ? '+' : '-');
            if (with_seq) {
                fputc(':', f);
                fwrite(fa->seq + fa->offsets[c] + mx[i].pos, 1, k, f);
            }
            i++;
        }
        fputc('\n', f);
    }
    if (f != stdout) fclose(f); else fflush(f);
    return 0;
}

/* ---------------------------------------------------------------- steps 2-3 (A.6) */
/* chained map u64 -> slot index, insertion ordered */
typedef struct { uint64_t key; uint64_t val; uint64_t val2; long long next; } ment_t;
typedef struct { long long* head; size_t nb; ment_t* e; size_t n, cap; } map_t;

static int map_init(map_t* m, size_t expect)
{
    m->nb = 16;
    while (m->nb < expect * 2) m->nb <<= 1;
    m->head = (long long*)malloc(m->nb * sizeof(long long));
    m->cap = expect ? expect : 16;
    m->e = (ment_t*)malloc(m->cap * sizeof(ment_t));
    m->n = 0;
    if (!m->head || !m->e) return -1;
    for (size_t i = 0; i < m->nb; i++) m->head[i] = -1;
    return 0;
}
static void map_free(map_t* m) { free(m->head); free(m->e); }
static size_t map_bucket(const map_t* m, uint64_t k)
{
    k ^= k >> 29; k *= 0xbf58476d1ce4e5b9ULL; k ^= k >> 32;
    return (size_t)k & (m->nb - 1);
}
static ment_t* map_find(const map_t* m, uint64_t k)
{
    for (long long i = m->head[map_bucket(m, k)]; i >= 0; i = m->e[i].next)
        if (m->e[i].key == k) return &m->e[i];
    return NULL;
}
static ment_t* map_put(map_t* m, uint64_t k)
{
    if (m->n == m->cap) {
        m->cap *= 2;
        m->e = (ment_t*)realloc(m->e, m->cap * sizeof(ment_t));
    }
    size_t b = map_bucket(m, k);
    ment_t* e = &m->e[m->n];
    e->key = k; e->val = 0; e->val2 = 0; e->next = m->head[b];
    m->head[b] = (long long)m->n++;
    return e;
}

/* ordered-pair map for edges[s][t] */
typedef struct { uint64_t s, t; size_t edge; long long next; } pent_t;
typedef struct { long long* head; size_t nb; pent_t* e; size_t n, cap; } pmap_t;
static size_t pmap_bucket(const pmap_t* m, uint64_t s, uint64_t t)
{
    uint64_t k = s * 0x9e3779b97f4a7c15ULL ^ (t + 0x7f4a7c15ULL + (s << 6) + (s >> 2));
    k ^= k >> 31; k *= 0x94d049bb133111ebULL; k ^= k >> 29;
    return (size_t)k & (m->nb - 1);
}
static pent_t* pmap_find(const pmap_t* m, uint64_t s, uint64_t t)
{
    for (long long i = m->head[pmap_bucket(m, s, t)]; i >= 0; i = m->e[i].next)
        if (m->e[i].s == s && m->e[i].t == t) return &m->e[i];
    return NULL;
}

typedef struct { size_t src_rank, seq; } ekey_t;
static const ekey_t* g_ekeys;
static int cmp_edge_order(const void* a, const void* b)
{
    const ekey_t* x = &g_ekeys[*(const size_t*)a];
    const ekey_t* y = &g_ekeys[*(const size_t*)b];
    if (x->src_rank != y->src_rank) return x->src_rank < y->src_rank ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}
static int cmp_u64(const void* a, const void* b)
{
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : x > y;
}

int mxo_filter_and_edges(int n_asm, const uint64_t* const* hashes, const uint32_t* const* contig,
                         const size_t* n, const double* weights,
                         uint8_t** uniq, uint8_t** keep,
                         mxo_edge_t** edges_out, size_t* n_edges,
                         uint64_t** vertices_out, size_t* n_vertices)
{
    if (n_asm < 1 || n_asm > 32) return -EINVAL;
    size_t total = 0;
    for (int a = 0; a < n_asm; a++) total += n[a];

    /* read_minimizers (:182-192): a hash seen more than once inside one assembly is dropped */
    /* filter_minimizers (:155-157): keep hashes unique in EVERY assembly                    */
    map_t inall;  /* key -> val = number of assemblies where unique */
    if (map_init(&inall, total / (size_t)n_asm + 16)) return -ENOMEM;
    for (int a = 0; a < n_asm; a++) {
        map_t cnt;
        if (map_init(&cnt, n[a] + 16)) return -ENOMEM;
        for (size_t i = 0; i < n[a]; i++) {
            ment_t* e = map_find(&cnt, hashes[a][i]);
            if (!e) e = map_put(&cnt, hashes[a][i]);
            e->val++;
        }
        for (size_t i = 0; i < n[a]; i++) {
            int u = map_find(&cnt, hashes[a][i])->val == 1;
            uniq[a][i] = (uint8_t)u;
            if (u) {
                ment_t* e = map_find(&inall, hashes[a][i]);
                if (!e) { if (a == 0) e = map_put(&inall, hashes[a][i]); else continue; }
                if (e->val == (uint64_t)a) e->val++;
            }
        }
        map_free(&cnt);
    }
    size_t n_keep = 0;
    for (int a = 0; a < n_asm; a++)
        for (size_t i = 0; i < n[a]; i++) {
            ment_t* e = uniq[a][i] ? map_find(&inall, hashes[a][i]) : NULL;
            keep[a][i] = (uint8_t)(e && e->val == (uint64_t)n_asm);
            n_keep += keep[a][i];
        }

    /* build_graph (:94-115) */
    pmap_t pm;
    pm.nb = 16;
    while (pm.nb < n_keep * 2 + 16) pm.nb <<= 1;
    pm.head = (long long*)malloc(pm.nb * sizeof(long long));
    pm.cap = n_keep + 16;
    pm.e = (pent_t*)malloc(pm.cap * sizeof(pent_t));
    pm.n = 0;
    for (size_t i = 0; i < pm.nb; i++) pm.head[i] = -1;
    map_t srcs;   /* key -> val = creation rank as a key of `edges` */
    map_init(&srcs, n_keep / (size_t)n_asm + 16);
    map_t verts;
    map_init(&verts, n_keep / (size_t)n_asm + 16);

    size_t ecap = n_keep + 16, ne = 0;
    mxo_edge_t* E = (mxo_edge_t*)malloc(ecap * sizeof(mxo_edge_t));
    ekey_t* K = (ekey_t*)malloc(ecap * sizeof(ekey_t));

    for (int a = 0; a < n_asm; a++) {
        int have = 0;
        uint64_t prev = 0;
        uint32_t prev_c = 0;
        for (size_t i = 0; i < n[a]; i++) {
            if (!keep[a][i]) continue;
            uint64_t h = hashes[a][i];
            if (!map_find(&verts, h)) map_put(&verts, h);
            if (have && prev_c == contig[a][i]) {
                pent_t* pe = pmap_find(&pm, prev, h);
                if (!pe) pe = pmap_find(&pm, h, prev);
                if (pe) {
                    E[pe->edge].support_mask |= 1u << a;
                } else {
                    ment_t* se = map_find(&srcs, prev);
                    if (!se) { se = map_put(&srcs, prev); se->val = srcs.n - 1; }
                    size_t b = pmap_bucket(&pm, prev, h);
                    pent_t* ne_ = &pm.e[pm.n];
                    ne_->s = prev; ne_->t = h; ne_->edge = ne; ne_->next = pm.head[b];
                    pm.head[b] = (long long)pm.n++;
                    E[ne].u = prev; E[ne].v = h; E[ne].support_mask = 1u << a; E[ne].weight = 0;
                    K[ne].src_rank = (size_t)se->val; K[ne].seq = ne;
                    ne++;
                }
            }
            have = 1; prev = h; prev_c = contig[a][i];
        }
    }
    /* formatted_edges order (:115): sources in creation order, targets in insertion order */
    size_t* order = (size_t*)malloc((ne ? ne : 1) * sizeof(size_t));
    for (size_t i = 0; i < ne; i++) order[i] = i;
    g_ekeys = K;
    qsort(order, ne, sizeof(size_t), cmp_edge_order);
    mxo_edge_t* Es = (mxo_edge_t*)malloc((ne ? ne : 1) * sizeof(mxo_edge_t));
    for (size_t i = 0; i < ne; i++) {
        Es[i] = E[order[i]];
        /* calc_total_weight (:54-56): sum() starts at int 0, adds in support-list order */
        double wsum = 0;
        for (int a = 0; a < n_asm; a++)
            if (Es[i].support_mask & (1u << a)) wsum += weights[a];
        Es[i].weight = wsum;
    }
    uint64_t* V = (uint64_t*)malloc((verts.n ? verts.n : 1) * sizeof(uint64_t));
    for (size_t i = 0; i < verts.n; i++) V[i] = verts.e[i].key;
    qsort(V, verts.n, sizeof(uint64_t), cmp_u64);

    *edges_out = Es; *n_edges = ne;
    *vertices_out = V; *n_vertices = verts.n;
    free(order); free(E); free(K);
    free(pm.head); free(pm.e);
    map_free(&srcs); map_free(&verts); map_free(&inall);
    return 0;
}

void mxo_free(void* p) { free(p); }
