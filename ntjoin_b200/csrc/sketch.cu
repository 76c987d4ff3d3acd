// sketch.cu -- host orchestration of step 1 (see sketch_kernels.cuh for the algorithm).
#include "engine.cuh"
#include "sketch_kernels.cuh"
#include "bitslice_kernels.cuh"
#include "scan_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace mxe {

static void build_tables(int k, SketchTables* T)
{
    const uint64_t seed[4] = {SEED_A, SEED_C, SEED_T, SEED_G};   // device code order A C T G
    uint32_t shi_rolk[4];
    for (int c = 0; c < 4; c++) {
        T->seed[c] = seed[c];
        uint64_t x = seed[c];
        for (int i = 0; i < k; i++) x = srol1(x);
        T->seed_rolk[c] = x;
        T->shi[c] = (uint32_t)(seed[c] >> 33);
        shi_rolk[c] = (uint32_t)(x >> 33);
    }
    T->f0 = 0; T->r0 = 0;
    for (int i = 0; i < k; i++) {
        T->f0 = rol31(T->f0) ^ T->shi[0];      // fwd of A^k
        T->r0 = rol31(T->r0) ^ T->shi[2];      // rev of A^k = fwd-style fold of T^k
    }
    for (int o = 0; o < 4; o++)
        for (int in = 0; in < 4; in++) {
            uint2 e;
            e.x = shi_rolk[o] ^ T->shi[in];             // fwd' = rol(fwd) ^ rol^k(seed[out]) ^ seed[in]
            e.y = shi_rolk[in ^ 2] ^ T->shi[o ^ 2];     // rev' = ror(rev ^ rol^k(seed[~in]) ^ seed[~out])
            T->t16[(o << 2) | in] = e;
        }
}

static inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

constexpr int BS2_H = 12, BS2_HS = 12;      // adder width / compared bits of the bit-sliced threshold test

// Tile geometry of the bit-sliced front end: L = 16 * Lw starts per stream, Lw odd.  Long streams amortise the k-1
// halo, short ones give more threads; pick the Lw that minimises (waves of resident threads) x (steps per thread).
static void scan_geometry(mxe_engine* e, uint64_t n, int k, ScanGeom* G)
{
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        const size_t rsm = bs2_smem_bytes(32);
        cudaFuncSetAttribute(scan_bs2_kernel<1, BS2_H, BS2_HS, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, scan_bs2_kernel<1, BS2_H, BS2_HS, 48>, BS2_THREADS, rsm) != cudaSuccess || blocks_per_sm < 1) {
            cudaGetLastError();
            blocks_per_sm = 3;
        }
    }
    const uint64_t slots = (uint64_t)e->sm_count * blocks_per_sm * BS2_THREADS;
    uint32_t best_lw = 9;
    uint64_t best_cost = ~0ULL;
    int lw_lo = 9, lw_hi = 63;
    if (e->scan_lw > 0) lw_lo = lw_hi = e->scan_lw | 1;
    for (int lw = lw_lo; lw <= lw_hi; lw += 2) {
        const uint64_t tiles = (n + 512ull * lw - 1) / (512ull * lw);
        const uint64_t waves = (tiles + slots - 1) / slots;
        const uint64_t R = (16ull * lw + k - 1 + 15) / 16;
        const uint64_t cost = waves * R;
        if (cost <= best_cost) { best_cost = cost; best_lw = (uint32_t)lw; }
    }
    G->n = n; G->n_words = (n + 31) / 32; G->k = k;
    G->Lw = best_lw;
    G->n_tiles = (n + 512ull * best_lw - 1) / (512ull * best_lw);
    G->R = (uint32_t)((16ull * best_lw + k - 1 + 15) / 16);
    G->ext = (uint32_t)std::max(2, (31 + k - 1) / 32);
}

// Start the chunked host->device copy of `h` into slot `s` (two copy streams = two copy engines).  Does not wait for
// the compute stream, only for the previous reader of the slot.
int h2d_issue(mxe_engine* e, H2DSlot& s, const uint8_t* h, uint64_t n)
{
    if (s.cap < n + 64) {
        if (s.d) { MXE_CUDA(cudaDeviceSynchronize()); MXE_CUDA(cudaFree(s.d)); s.d = nullptr; s.cap = 0; }
        MXE_CUDA(cudaMalloc((void**)&s.d, n + 64));
        s.cap = n + 64;
    }
    if (!s.consumed) MXE_CUDA(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
    const uint64_t CH = (uint64_t)e->h2d_chunk_mb << 20;
    const int n_chunks = (int)((n + CH - 1) / CH);
    while ((int)s.ev.size() < n_chunks) {
        cudaEvent_t ev;
        MXE_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        s.ev.push_back(ev);
    }
    if (s.consumed_valid)
        for (int c = 0; c < 2; c++) MXE_CUDA(cudaStreamWaitEvent(e->copy_stream[c], s.consumed, 0));
    int i = 0;
    for (uint64_t off = 0; off < n; off += CH, i++) {
        const uint64_t len = std::min<uint64_t>(CH, n - off);
        cudaStream_t cs = e->copy_stream[i & 1];
        MXE_CUDA(cudaMemcpyAsync(s.d + off, h + off, len, cudaMemcpyHostToDevice, cs));
        MXE_CUDA(cudaEventRecord(s.ev[i], cs));
    }
    s.h = h; s.n = n; s.chunk = CH; s.n_chunks = n_chunks;
    return MXE_OK;
}

// A sketch whose kernels are all enqueued but whose counts have not been read yet: lets the sketches of several
// assemblies run concurrently on two streams (sketch_device_many_impl) -- the bandwidth- and latency-bound kernels of
// one assembly fill the SMs that the ALU-bound candidate scan of the other leaves idle, and the other way round.
struct SketchPending {
    bool active = false;
    cudaStream_t stream = nullptr;
    uint64_t* host = nullptr;              // pinned: n_valid, n_cand, n_mx, gap count, gap windows
    uint64_t cand_cap = 0, gap_cap = 0, mx_cap = 0;
};

static int sketch_device_pass(mxe_engine* e, const uint8_t* d_seq, uint64_t n, const uint64_t* offsets, uint32_t n_contigs,
                              int k, int w, int flags, mxe_sketch* S, H2DSlot* staged, bool force_exact, bool* redo_exact,
                              SketchPending* defer = nullptr);

// second half of a deferred sketch: the one host round trip; *redo = a size bound was exceeded
static int sketch_complete(mxe_engine* e, mxe_sketch* S, SketchPending* D, bool* redo)
{
    *redo = false;
    if (!D->active) return MXE_OK;                     // empty input: nothing was enqueued
    MXE_CUDA(cudaStreamSynchronize(D->stream));
    const uint64_t n_valid = D->host[0], n_cand = D->host[1], n_mx = D->host[2], g0 = D->host[3], g1 = D->host[4];
    e->pinned_release(D->host, 64);
    D->host = nullptr;
    if (n_cand > D->cand_cap || g0 > D->gap_cap || n_mx > D->mx_cap) {
        void* outs[] = {S->d_out_hash, S->d_min_hash, S->d_pos, S->d_contig, S->d_forward};
        for (void* q : outs) if (q) cudaFreeAsync(q, D->stream);
        S->d_out_hash = S->d_min_hash = nullptr; S->d_pos = S->d_contig = nullptr; S->d_forward = nullptr;
        *redo = true;
        return MXE_OK;
    }
    S->n_valid = n_valid; S->n_cand = n_cand; S->n_gaps = g0; S->n_gap_windows = g1; S->n = n_mx;
    return MXE_OK;
}

// Several device-resident assemblies: assembly a is enqueued on stream (a & 1) -- the engine stream and its auxiliary
// stream -- and the counts are read after everything has been issued.
int sketch_device_many_impl(mxe_engine* e, int n_asm, const uint8_t* const* d_seq, const uint64_t* const* offsets, const uint32_t* n_contigs,
                            int k, int w, int flags, mxe_sketch* const* S)
{
    if (!e->aux_stream) MXE_CUDA(cudaStreamCreateWithFlags(&e->aux_stream, cudaStreamNonBlocking));
    if (!e->aux_event) MXE_CUDA(cudaEventCreateWithFlags(&e->aux_event, cudaEventDisableTiming));
    cudaStream_t main_stream = e->stream;
    MXE_CUDA(cudaEventRecord(e->aux_event, main_stream));              // inputs are ready in engine-stream order
    MXE_CUDA(cudaStreamWaitEvent(e->aux_stream, e->aux_event, 0));
    std::vector<SketchPending> pend((size_t)n_asm);
    int rc = MXE_OK;
    for (int a = 0; a < n_asm && rc == MXE_OK; a++) {
        const uint64_t n = n_contigs[a] ? offsets[a][n_contigs[a]] : 0;
        bool redo = false;
        e->stream = ((a & 1) && e->many_streams >= 2) ? e->aux_stream : main_stream;
        rc = sketch_device_pass(e, d_seq[a], n, offsets[a], n_contigs[a], k, w, flags, S[a], nullptr, false, &redo, &pend[a]);
        e->stream = main_stream;
    }
    for (int a = 0; a < n_asm; a++) {
        bool redo = false;
        int rc2 = sketch_complete(e, S[a], &pend[a], &redo);
        if (rc == MXE_OK) rc = rc2;
        if (rc == MXE_OK && redo) {                                    // rare: bounds exceeded -> exact sizes, on the engine stream
            const uint64_t n = n_contigs[a] ? offsets[a][n_contigs[a]] : 0;
            MXE_CUDA(cudaStreamSynchronize(e->aux_stream));
            rc = sketch_device_pass(e, d_seq[a], n, offsets[a], n_contigs[a], k, w, flags, S[a], nullptr, true, &redo);
        }
    }
    // later work on the engine stream (steps 2-3, copies) sees the results of the auxiliary stream
    MXE_CUDA(cudaEventRecord(e->aux_event, e->aux_stream));
    MXE_CUDA(cudaStreamWaitEvent(main_stream, e->aux_event, 0));
    return rc;
}

int sketch_device_impl(mxe_engine* e, const uint8_t* d_seq, uint64_t n, const uint64_t* offsets, uint32_t n_contigs,
                       int k, int w, int flags, mxe_sketch* S, H2DSlot* staged)
{
    bool redo = false;
    int rc = sketch_device_pass(e, d_seq, n, offsets, n_contigs, k, w, flags, S, staged, false, &redo);
    if (rc == MXE_OK && redo) {
        // a size bound of the asynchronous path was exceeded (rare: dense candidates).  The input is still resident --
        // a staged host buffer stays in its slot until the next copy is issued -- so the whole sketch is simply repeated
        // with exact sizes read back at every stage.
        e->arena.begin(e->stream);
        rc = sketch_device_pass(e, d_seq, n, offsets, n_contigs, k, w, flags, S, nullptr, true, &redo);
    }
    return rc;
}

static int sketch_device_pass(mxe_engine* e, const uint8_t* d_seq, uint64_t n, const uint64_t* offsets, uint32_t n_contigs,
                              int k, int w, int flags, mxe_sketch* S, H2DSlot* staged, bool force_exact, bool* redo_exact,
                              SketchPending* defer)
{
    if (k < 1 || k > 1024 || w < 1) { set_error("bad k/w (k=%d w=%d)", k, w); return MXE_ERR_ARG; }
    if (n_contigs && offsets[0] != 0) { set_error("offsets[0] must be 0"); return MXE_ERR_ARG; }
    for (uint32_t c = 0; c < n_contigs; c++) {
        if (offsets[c + 1] < offsets[c]) { set_error("offsets not monotone at %u", c); return MXE_ERR_ARG; }
        if (offsets[c + 1] - offsets[c] >= (1ULL << 32)) { set_error("record %u longer than 2^32-1 bases", c); return MXE_ERR_ARG; }
    }
    if (n_contigs && offsets[n_contigs] != n) { set_error("offsets[n_contigs] != n"); return MXE_ERR_ARG; }
    if (n && ((uintptr_t)d_seq & 15) != 0) { set_error("device sequence pointer must be 16-byte aligned"); return MXE_ERR_ARG; }

    cudaStream_t st = e->stream;
    S->eng = e; S->k = k; S->w = w; S->flags = flags;
    S->n_bases = n; S->n_contigs = n_contigs;
    S->offsets.assign(offsets, offsets + n_contigs + 1);
    S->n = 0;
    if (n == 0 || n_contigs == 0 || n < (uint64_t)k) return MXE_OK;

    Span whole(e, "sketch");

    SketchParams P;
    P.n = n; P.n_words = (n + 31) / 32; P.k = k; P.w = w;
    P.canon_min = (flags & MXE_CANON_MIN) ? 1 : 0;
    P.mul1 = 1u; P.mul2 = 2u;
    {
        double t = e->tau * 2147483648.0 / (double)w;
        P.T = t >= 2147483646.0 ? 2147483646u : (uint32_t)t;
    }
    if (e->chunk > 0) P.chunk = std::max(32, (e->chunk / 32) * 32);
    else {   // auto: long runs amortise the k-step warm-up, but keep >= ~4 waves of threads in flight
        uint64_t want = n / ((uint64_t)e->sm_count * 2048 * 2) + 1;
        P.chunk = want >= 512 ? 512 : want >= 256 ? 256 : 128;
    }
    SketchTables Tb;
    build_tables(k, &Tb);

    const uint64_t nW = P.n_words;
    const uint64_t n_vblocks = (nW + RANK_BLOCK_WORDS - 1) / RANK_BLOCK_WORDS;
    const uint64_t pk_words = 2 * nW + (uint64_t)(k / 16) + 16;
    P.pk_words = pk_words;

    // second-generation front end (scan_kernels.cuh): bit-sliced candidate scan fed by a tile-transposed copy of pk
    const int kmod = k % 31;
    const bool bs2 = e->cand_variant >= 4 && !P.canon_min && (kmod == 1 || kmod == 9 || kmod == 24) && k >= 16 && k <= BS2_MAX_K;
    ScanGeom G;
    memset(&G, 0, sizeof(G));
    if (bs2) scan_geometry(e, n, k, &G);

    DBuf<uint32_t> pk, V, C, M, PL, dirty;
    DBuf<uint64_t> vprefix, cprefix, mprefix, d_offsets, ostart;
    DBuf<uint32_t> vcounts, ccounts;
    MXE_TRY(vcounts.alloc(n_vblocks + 1, st));
    MXE_TRY(ccounts.alloc(n_vblocks + 1, st));
    MXE_TRY(pk.alloc(pk_words, st));
    MXE_TRY(V.alloc(nW, st));
    MXE_TRY(C.alloc(bs2 ? std::max<uint64_t>(nW, G.n_tiles * 16ull * G.Lw) : nW, st));
    MXE_TRY(M.alloc(nW, st));
    MXE_TRY(vprefix.alloc(n_vblocks + 1, st));
    MXE_TRY(cprefix.alloc(n_vblocks + 1, st));
    MXE_TRY(mprefix.alloc(n_vblocks + 1, st));
    MXE_TRY(d_offsets.alloc(n_contigs + 1, st));
    MXE_TRY(ostart.alloc(n_contigs + 1, st));
    MXE_CUDA(cudaMemcpyAsync(d_offsets.p, offsets, (n_contigs + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    MXE_CUDA(cudaMemsetAsync(pk.p + 2 * nW, 0, (pk_words - 2 * nW) * sizeof(uint32_t), st));
    MXE_CUDA(cudaMemsetAsync(M.p, 0, nW * sizeof(uint32_t), st));
    DirtyList DL{nullptr, nullptr, 0};
    if (bs2) {
        const uint32_t cap = (uint32_t)std::min<uint64_t>(nW / 16 + 4096, 0x7FFFFFFFu);
        MXE_TRY(PL.alloc(G.n_tiles * (uint64_t)G.R * 32, st));
        MXE_TRY(dirty.alloc((size_t)cap + 1, st));
        DL = DirtyList{dirty.p + 1, dirty.p, cap};
        MXE_CUDA(cudaMemsetAsync(dirty.p, 0, sizeof(uint32_t), st));
        MXE_CUDA(cudaMemsetAsync(vcounts.p, 0, (n_vblocks + 1) * sizeof(uint32_t), st));
        MXE_CUDA(cudaMemsetAsync(C.p, 0, std::max<uint64_t>(nW, G.n_tiles * 16ull * G.Lw) * sizeof(uint32_t), st));
    }

    // ---- pack + validity
    if (bs2) {
        Span sp(e, "pack");
        const size_t psm = pack2_smem_bytes(G.Lw, G.ext);
        const unsigned pthreads = 16u * G.Lw + G.ext <= 320 ? 128 : PACK2_THREADS;      // short tiles: fewer idle lanes per CTA
        const uint64_t tile = 512ull * G.Lw;
        if (!staged) {
            Span k1(e, "k_pack2");
            MXE_LAUNCH(e, pack2_kernel, (unsigned)G.n_tiles, pthreads, psm, d_seq, G, pk.p, PL.p, V.p, vcounts.p, DL, (uint64_t)0);
        } else {
            // host input: a tile is packed once its bytes and its halo (ext units) have landed
            const uint64_t CH = staged->chunk;
            uint64_t T0 = 0;
            int i = 0;
            for (uint64_t off = 0; off < n; off += CH, i++) {
                const uint64_t len = std::min<uint64_t>(CH, n - off);
                MXE_CUDA(cudaStreamWaitEvent(st, staged->ev[i], 0));
                const bool last = off + len >= n;
                const uint64_t have = off + len;
                uint64_t T1 = last ? G.n_tiles : (have > 32ull * G.ext ? (have - 32ull * G.ext) / tile : 0);
                if (T1 > G.n_tiles) T1 = G.n_tiles;
                if (T1 > T0) MXE_LAUNCH(e, pack2_kernel, (unsigned)(T1 - T0), pthreads, psm, d_seq, G, pk.p, PL.p, V.p, vcounts.p, DL, T0);
                T0 = std::max(T0, T1);
            }
            MXE_CUDA(cudaEventRecord(staged->consumed, st));
            staged->consumed_valid = true;
            staged->h = nullptr;
        }
        if (n_contigs > 1) MXE_LAUNCH(e, boundary2_kernel, grid_for(n_contigs - 1, 128), 128, 0, d_offsets.p, n_contigs, n, k, V.p, vcounts.p, DL);
    } else {
        Span sp(e, "pack");
        if (!staged) {
            MXE_LAUNCH(e, pack_kernel, grid_for(nW, 256), 256, 0, d_seq, P, pk.p, V.p, vcounts.p, (uint64_t)0, nW);
        } else {
            // host input: the chunks land one after the other on the copy streams (h2d_issue); each is packed as soon as
            // its arrival event has fired
            const uint64_t CH = staged->chunk;
            int i = 0;
            for (uint64_t off = 0; off < n; off += CH, i++) {
                const uint64_t len = std::min<uint64_t>(CH, n - off);
                MXE_CUDA(cudaStreamWaitEvent(st, staged->ev[i], 0));
                // ranges lag one warp (32 words) behind the copied bytes: the V halo of a warp reads the next words
                const bool last = off + len >= n;
                const uint64_t t0 = off ? (off >> 5) - 32 : 0;
                const uint64_t t1 = last ? nW : ((off + len) >> 5) - 32;
                if (t1 > t0) MXE_LAUNCH(e, pack_kernel, grid_for(t1 - t0, 256), 256, 0, d_seq, P, pk.p, V.p, vcounts.p, t0, t1);
            }
            // the sequence itself is not read after pack: the slot may be refilled from here on
            MXE_CUDA(cudaEventRecord(staged->consumed, st));
            staged->consumed_valid = true;
            staged->h = nullptr;
        }
        if (n_contigs > 1) MXE_LAUNCH(e, boundary_kernel, grid_for(n_contigs - 1, 128), 128, 0, d_offsets.p, n_contigs, P, V.p, vcounts.p);
    }
    {
        Span sp(e, "rank");
        MXE_TRY(exclusive_scan_u32_u64(e, vcounts.p, vprefix.p, n_vblocks));
        MXE_LAUNCH(e, contig_bounds_kernel, grid_for(n_contigs + 1, 128), 128, 0, d_offsets.p, n_contigs, P, V.p, vprefix.p, ostart.p);
    }

    // ---- candidates
    bool c_counted = false;
    {
        Span sp(e, "cand");
        uint64_t n_threads = (n + P.chunk - 1) / P.chunk;
        bool fast = e->cand_variant >= 1 && (k % 4 == 0) && k >= 4;
        bool sliced = !bs2 && e->cand_variant >= 2 && !P.canon_min && (kmod == 1 || kmod == 9 || kmod == 24) && k >= 16 && k <= 256 &&
                      (e->cand_variant >= 3 || n >= ((uint64_t)1 << 29));
        if (bs2) {
            Bs2Params BP;
            BP.f0 = Tb.f0; BP.r0 = Tb.r0;
            const uint32_t TH = P.T >> (31 - BS2_H);
            uint32_t Q = (TH + 1) >> (BS2_H - BS2_HS);
            if (Q > (1u << BS2_HS) - 1) Q = (1u << BS2_HS) - 1;
            const uint32_t K = ((1u << BS2_HS) - 1) - Q;
            for (int i = 0; i < 16; i++) BP.kmask[i] = (i < BS2_HS && ((K >> i) & 1u)) ? 0xFFFFFFFFu : 0u;
            const size_t rsm = bs2_smem_bytes(k);
            const unsigned grid = grid_for(G.n_tiles, BS2_THREADS);
#define MXE_BS2(KM, RG)                                                                                                               \
    do {                                                                                                                              \
        MXE_CUDA(cudaFuncSetAttribute(scan_bs2_kernel<KM, BS2_H, BS2_HS, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm)); \
        MXE_LAUNCH(e, (scan_bs2_kernel<KM, BS2_H, BS2_HS, RG>), grid, BS2_THREADS, rsm, PL.p, G, BP, C.p);                            \
    } while (0)
            {
                Span k2(e, "k_scan");
                if (kmod == 1) MXE_BS2(1, 48);           // k = 32
                else if (kmod == 9) MXE_BS2(9, 64);      // k = 40
                else MXE_BS2(24, 48);                    // k = 24
            }
#undef MXE_BS2
            MXE_LAUNCH(e, dirty_fix_kernel, (unsigned)(e->sm_count * 8), 256, 0, DL, V.p, C.p, nW);
        } else if (sliced) {
            // bit-sliced kernel: transpose pk into bit planes, run 32 streams per thread, then C &= V with counts
            MXE_CUDA(cudaMemsetAsync(C.p, 0, nW * sizeof(uint32_t), st));      // it sets candidate bits with atomics; the others write whole words
            BsParams BP;
            BP.n = n; BP.k = k; BP.iters = bs_iters(k); BP.rows = bs_rows(k);
            BP.n_tiles = (n + BS_TILE - 1) / BS_TILE;
            BP.f0 = Tb.f0; BP.r0 = Tb.r0;
            const uint32_t TH = P.T >> (31 - BS_H);
            uint32_t Q = (TH + 1) >> (BS_H - BS_HS);
            if (Q > (1u << BS_HS) - 1) Q = (1u << BS_HS) - 1;
            const uint32_t K = ((1u << BS_HS) - 1) - Q;
            for (int i = 0; i < BS_HS; i++) BP.kmask[i] = ((K >> i) & 1u) ? 0xFFFFFFFFu : 0u;
            DBuf<uint32_t> PL;
            const uint64_t n_groups = (BP.n_tiles + 3) / 4;
            MXE_TRY(PL.alloc(n_groups * (uint64_t)BP.rows * 8, st));
            const int n_in = PLANE_SLICE + BP.rows / 16;
            const size_t psm = (4 * (size_t)(n_in + n_in / 32 + 1) + 128) * sizeof(uint32_t);
            MXE_LAUNCH(e, plane_kernel, (unsigned)n_groups, 128, psm, pk.p, pk_words, BP.n_tiles, BP.rows, PL.p);
            if (kmod == 1) MXE_LAUNCH(e, cand_bs_kernel<1>, grid_for(BP.n_tiles, 128), 128, 0, PL.p, BP, C.p);
            else if (kmod == 9) MXE_LAUNCH(e, cand_bs_kernel<9>, grid_for(BP.n_tiles, 128), 128, 0, PL.p, BP, C.p);
            else MXE_LAUNCH(e, cand_bs_kernel<24>, grid_for(BP.n_tiles, 128), 128, 0, PL.p, BP, C.p);
            MXE_LAUNCH(e, cand_mask_count_kernel, grid_for(nW, 256), 256, 0, C.p, V.p, nW, ccounts.p);
            c_counted = true;
        } else if (fast) {
            // the staged kernel needs a power-of-two run length between 128 and 1024
            int c = 128;
            while (c * 2 <= P.chunk && c < 1024) c *= 2;
            P.chunk = c;
            n_threads = (n + P.chunk - 1) / P.chunk;
            size_t smem = cand31_smem_bytes(P.chunk, k);
            c_counted = true;
#define MXE_CAND31(CM, FO)                                                                                                  \
    do {                                                                                                                    \
        MXE_CUDA(cudaFuncSetAttribute(cand31_kernel<CM, FO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        MXE_LAUNCH(e, (cand31_kernel<CM, FO>), grid_for(n_threads, CAND_THREADS), CAND_THREADS, smem, pk.p, V.p, P, Tb, C.p, ccounts.p); \
    } while (0)
            if (P.canon_min) MXE_CAND31(1, 0);
            else if (e->fma_offload) MXE_CAND31(0, 1);
            else MXE_CAND31(0, 0);
#undef MXE_CAND31
        } else {
            MXE_LAUNCH(e, cand_generic_kernel, grid_for(n_threads, 128), 128, 0, pk.p, V.p, P, Tb, C.p);
        }
    }
    {
        Span sp(e, "rank");
        if (c_counted) MXE_TRY(exclusive_scan_u32_u64(e, ccounts.p, cprefix.p, n_vblocks));
        else MXE_TRY(bitmap_rank_build(e, C.p, nW, cprefix.p));
    }

    // ---- sizes.  Default: arrays and grids from upper bounds, the exact counts stay on the device (kernels read them through
    // pointers) and come back in ONE host round trip at the end of the sketch; if a bound turns out too small (dense
    // candidates: low-complexity sequence) the tail is repeated with exact sizes.  `exact` = the round-1 path: a round trip
    // after the candidate count, after the gap count and after the minimizer count.
    const bool exact = force_exact || !e->async_sizes || e->prune;
    uint64_t n_valid = 0, n_cand = 0;
    const uint64_t* d_ncand = cprefix.p + n_vblocks;
    const uint64_t* d_nmx = mprefix.p + n_vblocks;
    uint64_t cand_cap, mx_cap;
    if (exact) {
        uint64_t totals[2] = {0, 0};
        MXE_CUDA(cudaMemcpyAsync(&totals[0], vprefix.p + n_vblocks, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaMemcpyAsync(&totals[1], cprefix.p + n_vblocks, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaStreamSynchronize(st));
        n_valid = totals[0]; n_cand = totals[1];
        cand_cap = n_cand;
        d_ncand = nullptr;
        mx_cap = 0;
    } else {
        const double dens = std::min(1.0, (e->tau * 1.08 + 0.5) / (double)w);                // candidate density incl. the superset margin
        cand_cap = std::min<uint64_t>(n, (uint64_t)(((double)n * dens * 1.25 + (double)n_contigs + (double)(1u << 18)) * e->bound_scale) + 1);
        mx_cap = std::min<uint64_t>(n, (uint64_t)(((double)n * 2.0 / ((double)w + 1.0) * 1.25 + 2.0 * n_contigs + (double)(1u << 16)) * e->bound_scale) + 1);
    }

    // ---- candidate evaluation + sparse window selection
    DBuf<uint64_t> cpos, ch0, cord;
    DBuf<uint32_t> cctg;
    DBuf<Gap> gaps;
    DBuf<unsigned long long> gcount;
    MXE_TRY(cpos.alloc(cand_cap, st));
    MXE_TRY(ch0.alloc(cand_cap, st));
    MXE_TRY(cord.alloc(cand_cap, st));
    MXE_TRY(cctg.alloc(cand_cap, st));
    MXE_TRY(gcount.alloc(2, st));
    uint64_t gap_cap = std::max<uint64_t>(65536, cand_cap / 8 + n_contigs);   // grown on demand below (exact path)
    unsigned long long gc[2] = {0, 0};
    uint64_t n_sel = cand_cap;      // candidates that reach the exact stages (upper bound on the async path)
    const uint64_t* d_nsel = d_ncand;
    uint64_t *s_pos = cpos.p, *s_ord = cord.p;
    uint32_t* s_ctg = cctg.p;
    DBuf<uint64_t> cpos2, cord2, pprefix;
    DBuf<uint32_t> cctg2, klo, khi, pflag;
    DBuf<uint64_t> PF, PR;                 // position-specific hash tables (k/4 groups x 256), shared by the exact-hash kernels
    const int G4 = k / 4;
    const bool pos_tables = G4 >= 1 && G4 <= HASHPOS_MAX_GROUPS;
    const size_t hsm = (size_t)2 * G4 * 256 * sizeof(uint64_t);
    {
        Span sp(e, "eval");
        if (cand_cap) {
            MXE_LAUNCH(e, cand_extract_kernel, grid_for((n_vblocks + XBLOCKS - 1) / XBLOCKS * 32, 256), 256, 0, C.p, V.p, nW, cprefix.p, vprefix.p, n_vblocks,
                       d_offsets.p, n_contigs, (uint64_t)w, cpos.p, cord.p, cctg.p, cand_cap);
            if (e->prune && exact) {
                MXE_TRY(klo.alloc(n_cand, st)); MXE_TRY(khi.alloc(n_cand, st)); MXE_TRY(pflag.alloc(n_cand, st));
                MXE_TRY(pprefix.alloc(n_cand + 1, st));
                MXE_LAUNCH(e, cand_key31_kernel, grid_for(n_cand, 256), 256, 0, cpos.p, n_cand, pk.p, P, Tb, klo.p, khi.p);
                MXE_LAUNCH(e, prune_kernel, grid_for(n_cand, 256), 256, 0, klo.p, khi.p, cord.p, cctg.p, n_cand, ostart.p, P, pflag.p);
                MXE_TRY(exclusive_scan_u32_u64(e, pflag.p, pprefix.p, n_cand));
                MXE_CUDA(cudaMemcpyAsync(&n_sel, pprefix.p + n_cand, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
                MXE_CUDA(cudaStreamSynchronize(st));
                MXE_TRY(cpos2.alloc(n_sel, st)); MXE_TRY(cord2.alloc(n_sel, st)); MXE_TRY(cctg2.alloc(n_sel, st));
                MXE_LAUNCH(e, cand_compact_kernel, grid_for(n_cand, 256), 256, 0, pflag.p, pprefix.p, n_cand, cpos.p, cord.p, cctg.p,
                           cpos2.p, cord2.p, cctg2.p);
                s_pos = cpos2.p; s_ord = cord2.p; s_ctg = cctg2.p;
            }
            if (n_sel) {
                if (pos_tables) {
                    // position-specific tables (built once per sketch) -> no rotations per candidate
                    MXE_TRY(PF.alloc((size_t)G4 * 256, st)); MXE_TRY(PR.alloc((size_t)G4 * 256, st));
                    MXE_LAUNCH(e, hash_pos_tables_kernel, G4, 256, 0, Tb, k, PF.p, PR.p);
                    MXE_CUDA(cudaFuncSetAttribute(cand_hash_pos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
                    const unsigned grid = (unsigned)std::min<uint64_t>(grid_for(n_sel, 256), (uint64_t)e->sm_count * 5);
                    MXE_LAUNCH(e, cand_hash_pos_kernel, grid, 256, hsm, s_pos, d_nsel, n_sel, pk.p, P, Tb, PF.p, PR.p, ch0.p);
                } else {
                    MXE_LAUNCH(e, cand_hash_kernel, grid_for(n_sel, 256), 256, 0, s_pos, d_nsel, n_sel, pk.p, P, Tb, ch0.p);
                }
            }
        }
    }
    {
        Span sp(e, "select");
        for (int attempt = 0; attempt < 2; attempt++) {
            MXE_TRY(gaps.alloc(gap_cap, st));
            MXE_CUDA(cudaMemsetAsync(gcount.p, 0, 2 * sizeof(unsigned long long), st));
            GapList G{gaps.p, gcount.p, gcount.p + 1, gap_cap};
            // 32-bit window arithmetic whenever every padded ordinal (+ 2w) fits (see select_kernel); the number of valid
            // k-mers is bounded by the number of bases where it is not known on the host
            const uint64_t nv_bound = exact ? n_valid : n;
            const bool narrow = e->select_narrow && n_sel < 0xFFFFFFFFULL &&
                                nv_bound + ((uint64_t)n_contigs + 2) * (uint64_t)w < 0xFFFFFFFFULL;
            if (n_sel && narrow)
                MXE_LAUNCH(e, select_kernel<true>, grid_for(n_sel, 256), 256, 0, s_pos, ch0.p, s_ord, s_ctg, d_nsel, n_sel, ostart.p, P, M.p, G);
            else if (n_sel)
                MXE_LAUNCH(e, select_kernel<false>, grid_for(n_sel, 256), 256, 0, s_pos, ch0.p, s_ord, s_ctg, d_nsel, n_sel, ostart.p, P, M.p, G);
            MXE_LAUNCH(e, empty_contig_gap_kernel, grid_for(n_contigs, 128), 128, 0, s_pos, d_nsel, n_sel, d_offsets.p, n_contigs, ostart.p, P, G);
            if (!exact) break;
            MXE_CUDA(cudaMemcpyAsync(gc, gcount.p, sizeof(gc), cudaMemcpyDeviceToHost, st));
            MXE_CUDA(cudaStreamSynchronize(st));
            if (gc[0] <= gap_cap) break;
            gap_cap = gc[0];   // rerun with exact capacity (M updates are idempotent)
        }
    }

    // ---- dense gap windows
    if (gc[0] || !exact) {
        Span sp(e, "gap");
        unsigned grid = exact ? (unsigned)std::min<uint64_t>(gc[0], (uint64_t)e->sm_count * 4) : (unsigned)(e->sm_count * 4);
        DBuf<uint64_t> sh, sq, sbm, sbi;
        uint64_t stride = (uint64_t)GAP_CHUNK + (uint64_t)w;
        MXE_TRY(sh.alloc(stride * grid, st));
        MXE_TRY(sq.alloc(stride * grid, st));
        MXE_TRY(sbm.alloc((stride / 32 + 2) * grid, st));
        MXE_TRY(sbi.alloc((stride / 32 + 2) * grid, st));
        MXE_LAUNCH(e, gap_kernel, grid, 256, 0, gaps.p, exact ? (const unsigned long long*)nullptr : gcount.p, exact ? (uint64_t)gc[0] : gap_cap,
                   pk.p, V.p, vprefix.p, n_vblocks, P, Tb, sh.p, sq.p, sbm.p, sbi.p, M.p);
    }

    // ---- ordered emission
    {
        Span sp(e, "rank");
        MXE_TRY(bitmap_rank_build(e, M.p, nW, mprefix.p));
    }
    uint64_t n_mx = 0;
    if (exact) {
        MXE_CUDA(cudaMemcpyAsync(&n_mx, mprefix.p + n_vblocks, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaStreamSynchronize(st));
        mx_cap = n_mx;
        d_nmx = nullptr;
    }
    if (mx_cap) {
        Span sp(e, "emit");
        MXE_CUDA(cudaMallocAsync((void**)&S->d_out_hash, mx_cap * sizeof(uint64_t), st));
        MXE_CUDA(cudaMallocAsync((void**)&S->d_min_hash, mx_cap * sizeof(uint64_t), st));
        MXE_CUDA(cudaMallocAsync((void**)&S->d_pos, mx_cap * sizeof(uint32_t), st));
        MXE_CUDA(cudaMallocAsync((void**)&S->d_contig, mx_cap * sizeof(uint32_t), st));
        MXE_CUDA(cudaMallocAsync((void**)&S->d_forward, mx_cap * sizeof(uint8_t), st));
        DBuf<uint64_t> mpos;
        MXE_TRY(mpos.alloc(mx_cap, st));
        MXE_TRY(bitmap_extract(e, M.p, nW, mprefix.p, mpos.p, mx_cap));
        if (pos_tables && PF.p) {
            // the position-specific tables of the candidate stage, staged once per CTA of a persistent grid
            MXE_CUDA(cudaFuncSetAttribute(final_eval_pos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
            const unsigned grid = (unsigned)std::min<uint64_t>(grid_for(mx_cap, 256), (uint64_t)e->sm_count * 4);
            MXE_LAUNCH(e, final_eval_pos_kernel, grid, 256, hsm, mpos.p, d_nmx, mx_cap, pk.p, d_offsets.p, n_contigs, P, Tb, PF.p, PR.p,
                       S->d_out_hash, S->d_min_hash, S->d_pos, S->d_contig, S->d_forward);
        } else {
            MXE_LAUNCH(e, final_eval_kernel, grid_for(mx_cap, 256), 256, 0, mpos.p, d_nmx, mx_cap, pk.p, d_offsets.p, n_contigs, P, Tb,
                       S->d_out_hash, S->d_min_hash, S->d_pos, S->d_contig, S->d_forward);
        }
    }
    if (!exact && defer) {
        // counts into pinned memory; the caller reads them after it has enqueued the other assemblies (sketch_complete)
        defer->host = (uint64_t*)e->pinned_alloc(64);
        if (!defer->host) { set_error("pinned allocation failed"); return MXE_ERR_NOMEM; }
        MXE_CUDA(cudaMemcpyAsync(&defer->host[0], vprefix.p + n_vblocks, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaMemcpyAsync(&defer->host[1], cprefix.p + n_vblocks, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaMemcpyAsync(&defer->host[2], mprefix.p + n_vblocks, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaMemcpyAsync(&defer->host[3], gcount.p, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        defer->active = true; defer->stream = st;
        defer->cand_cap = cand_cap; defer->gap_cap = gap_cap; defer->mx_cap = mx_cap;
        MXE_CUDA(cudaGetLastError());
        return MXE_OK;
    }
    if (!exact) {
        // the one host round trip of the sketch: every count at once
        uint64_t back[3] = {0, 0, 0};
        MXE_CUDA(cudaMemcpyAsync(&back[0], vprefix.p + n_vblocks, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaMemcpyAsync(&back[1], cprefix.p + n_vblocks, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaMemcpyAsync(&back[2], mprefix.p + n_vblocks, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaMemcpyAsync(gc, gcount.p, sizeof(gc), cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaStreamSynchronize(st));
        n_valid = back[0]; n_cand = back[1]; n_mx = back[2];
        if (n_cand > cand_cap || gc[0] > gap_cap || n_mx > mx_cap) {
            // a bound was too small: drop what was produced and do the tail again with exact sizes
            void* outs[] = {S->d_out_hash, S->d_min_hash, S->d_pos, S->d_contig, S->d_forward};
            for (void* q : outs) if (q) cudaFreeAsync(q, st);
            S->d_out_hash = S->d_min_hash = nullptr; S->d_pos = S->d_contig = nullptr; S->d_forward = nullptr;
            *redo_exact = true;
            return MXE_OK;
        }
    }
    S->n_valid = n_valid; S->n_cand = n_cand;
    S->n_gaps = gc[0]; S->n_gap_windows = gc[1];
    S->n = n_mx;
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

}  // namespace mxe
