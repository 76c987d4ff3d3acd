// api.cu -- the C ABI declared in include/mxe.h, plus host-side FASTA ingest and TSV output.
#include "engine.cuh"

#include <errno.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

namespace mxe {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static thread_local Arena* g_arena = nullptr;
Arena* current_arena() { return g_arena; }
ArenaScope::ArenaScope(Engine* e_) : e(e_) { e->arena.begin(e->stream); g_arena = &e->arena; }
ArenaScope::~ArenaScope() { g_arena = nullptr; e->arena.end(); }

void Arena::begin(cudaStream_t st)
{
    active = true;
    if (chunks.size() > 1) {
        size_t total = 0;
        for (auto& c : chunks) total += c.bytes;
        cudaStreamSynchronize(st);
        for (auto& c : chunks) cudaFree(c.p);
        chunks.clear();
        char* p = nullptr;
        if (cudaMalloc((void**)&p, total) == cudaSuccess) chunks.push_back({p, total, 0});
        else cudaGetLastError();
    }
    for (auto& c : chunks) c.used = 0;
}
void* Arena::alloc(size_t bytes)
{
    bytes = (bytes + 511) & ~(size_t)511;
    for (auto& c : chunks)
        if (c.bytes - c.used >= bytes) { void* q = c.p + c.used; c.used += bytes; return q; }
    size_t want = bytes > ((size_t)64 << 20) ? bytes : ((size_t)64 << 20);
    char* p = nullptr;
    if (cudaMalloc((void**)&p, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    chunks.push_back({p, want, bytes});
    return p;
}
void Arena::destroy()
{
    for (auto& c : chunks) cudaFree(c.p);
    chunks.clear();
}

cudaEvent_t Engine::get_event()
{
    if (!event_pool.empty()) { cudaEvent_t ev = event_pool.back(); event_pool.pop_back(); return ev; }
    cudaEvent_t ev = nullptr;
    cudaEventCreate(&ev);
    return ev;
}
void Engine::span_begin(const char*, cudaEvent_t* a)
{
    *a = get_event();
    cudaEventRecord(*a, stream);
}
void Engine::span_end(const char* name, cudaEvent_t a, uint64_t n_launch)
{
    cudaEvent_t b = get_event();
    cudaEventRecord(b, stream);
    PhaseTimer& t = timers[name];
    t.spans.push_back({a, b});
    t.launches += n_launch;
}
}  // namespace mxe

using namespace mxe;

// Engines that are still alive.  Objects handed out by an engine (sketches, results, stage handles) may be released
// after it -- a garbage-collected host language frees in any order at interpreter exit -- and must then not touch it:
// their device memory went with the engine's context, only the host struct is left to delete.
#include <mutex>
#include <set>
static std::mutex g_live_mu;
static std::set<const void*> g_live_engines;
namespace mxe {
bool engine_alive(const void* e)
{
    std::lock_guard<std::mutex> lk(g_live_mu);
    return g_live_engines.count(e) != 0;
}
}

void* mxe_engine::pinned_alloc(size_t bytes)
{
    if (bytes == 0) bytes = 1;
    for (size_t i = 0; i < pinned_free.size(); i++) {
        if (pinned_free[i].bytes >= bytes && pinned_free[i].bytes <= 2 * bytes + 4096) {
            void* p = pinned_free[i].p;
            pinned_free.erase(pinned_free.begin() + i);
            return p;
        }
    }
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void mxe_engine::pinned_release(void* p, size_t bytes)
{
    if (!p) return;
    if (pinned_free.size() >= 16) { cudaFreeHost(pinned_free[0].p); pinned_free.erase(pinned_free.begin()); }
    pinned_free.push_back({p, bytes});
}

extern "C" {

const char* mxe_version(void) { return "ntjoin_b200 mxe 0.1 (sm_100a)"; }
const char* mxe_last_error(void) { return g_err; }

int mxe_create(int device, mxe_t** out)
{
    if (!out) { set_error("out is NULL"); return MXE_ERR_ARG; }
    *out = nullptr;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(err));
        return MXE_ERR_CUDA;
    }
    if (device < 0 || device >= count) { set_error("device %d out of range (have %d)", device, count); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(device));
    mxe_engine* e = new mxe_engine();
    e->device = device;
    // on any failure below the half-built engine is released again (streams included)
    auto fail = [&](const char* what, cudaError_t err) {
        set_error("%s: %s", what, cudaGetErrorString(err));
        for (int c = 0; c < 2; c++) if (e->copy_stream[c]) cudaStreamDestroy(e->copy_stream[c]);
        if (e->own_stream) cudaStreamDestroy(e->own_stream);
        delete e;
        return MXE_ERR_CUDA;
    };
    cudaError_t err0;
    if ((err0 = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", err0);
    e->stream = e->own_stream;
    for (int c = 0; c < 2; c++)
        if ((err0 = cudaStreamCreateWithFlags(&e->copy_stream[c], cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", err0);
    if (const char* s = getenv("MXE_H2D_CHUNK_MB")) e->h2d_chunk_mb = std::max(1, atoi(s));
    cudaDeviceProp prop;
    if ((err0 = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", err0);
    e->sm_count = prop.multiProcessorCount;
    cudaMemPool_t pool;
    if ((err0 = cudaDeviceGetDefaultMemPool(&pool, device)) != cudaSuccess) return fail("cudaDeviceGetDefaultMemPool", err0);
    uint64_t thresh = ~0ULL;   // keep freed blocks cached in the pool across steps
    if ((err0 = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh)) != cudaSuccess) return fail("cudaMemPoolSetAttribute", err0);
    if (const char* s = getenv("MXE_TAU")) e->tau = atof(s);
    if (const char* s = getenv("MXE_CHUNK")) e->chunk = atoi(s);
    if (const char* s = getenv("MXE_CAND_VARIANT")) e->cand_variant = atoi(s);
    if (const char* s = getenv("MXE_PRUNE")) e->prune = atoi(s) != 0;
    if (const char* s = getenv("MXE_SORT_BITS")) e->sort_bits = atoi(s);
    if (const char* s = getenv("MXE_FILTER_VARIANT")) e->filter_variant = atoi(s);
    if (const char* s = getenv("MXE_FMA_OFFLOAD")) e->fma_offload = atoi(s) != 0;
    if (const char* s = getenv("MXE_SELECT_NARROW")) e->select_narrow = atoi(s) != 0;
    { std::lock_guard<std::mutex> lk(g_live_mu); g_live_engines.insert(e); }
    *out = e;
    return MXE_OK;
}

void mxe_destroy(mxe_t* e)
{
    if (!e || !engine_alive(e)) return;
    { std::lock_guard<std::mutex> lk(g_live_mu); g_live_engines.erase(e); }
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    for (auto& kv : e->timers)
        for (auto& sp : kv.second.spans) { cudaEventDestroy(sp.first); cudaEventDestroy(sp.second); }
    for (auto ev : e->event_pool) cudaEventDestroy(ev);
    if (e->local_p2p) { p2p_release(e->local_p2p, true); e->local_p2p = nullptr; }
    for (auto& p : e->pinned_free) cudaFreeHost(p.p);
    e->arena.destroy();
    for (int c = 0; c < 2; c++) if (e->copy_stream[c]) { cudaStreamSynchronize(e->copy_stream[c]); cudaStreamDestroy(e->copy_stream[c]); }
    for (auto& sl : e->slot) {
        for (auto ev : sl.ev) cudaEventDestroy(ev);
        if (sl.consumed) cudaEventDestroy(sl.consumed);
        if (sl.d) cudaFree(sl.d);
    }
    if (e->aux_stream) { cudaStreamSynchronize(e->aux_stream); cudaStreamDestroy(e->aux_stream); }
    if (e->aux_event) cudaEventDestroy(e->aux_event);
    if (e->d2h_stream) { cudaStreamSynchronize(e->d2h_stream); cudaStreamDestroy(e->d2h_stream); }
    if (e->d2h_event) cudaEventDestroy(e->d2h_event);
    cudaStreamDestroy(e->own_stream);
    delete e;
}

int mxe_set_stream(mxe_t* e, void* cuda_stream)
{
    if (!e) { set_error("null engine"); return MXE_ERR_ARG; }
    cudaStreamSynchronize(e->stream);
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return MXE_OK;
}

int mxe_set_option(mxe_t* e, const char* name, double value)
{
    if (!e || !name) { set_error("null argument"); return MXE_ERR_ARG; }
    if (!strcmp(name, "tau")) { if (value <= 0) { set_error("tau must be > 0"); return MXE_ERR_ARG; } e->tau = value; }
    else if (!strcmp(name, "chunk")) { if (value != 0 && value < 32) { set_error("chunk must be 0 (auto) or >= 32"); return MXE_ERR_ARG; } e->chunk = (int)value; }
    else if (!strcmp(name, "cand_variant")) e->cand_variant = (int)value;
    else if (!strcmp(name, "scan_lw")) e->scan_lw = (int)value;
    else if (!strcmp(name, "prune")) e->prune = value != 0;
    else if (!strcmp(name, "sort_bits")) e->sort_bits = (int)value;
    else if (!strcmp(name, "filter_variant")) e->filter_variant = (int)value;
    else if (!strcmp(name, "async_sizes")) e->async_sizes = value != 0;
    else if (!strcmp(name, "many_streams")) e->many_streams = value >= 2 ? 2 : 1;
    else if (!strcmp(name, "bound_scale")) { if (value <= 0) { set_error("bound_scale must be > 0"); return MXE_ERR_ARG; } e->bound_scale = value; }
    else if (!strcmp(name, "fma_offload")) e->fma_offload = value != 0;
    else if (!strcmp(name, "select_narrow")) e->select_narrow = value != 0;
    else if (!strcmp(name, "timing")) { e->timing = value != 0; e->timing_fine = value >= 2; }
    else { set_error("unknown option %s", name); return MXE_ERR_ARG; }
    return MXE_OK;
}

// ------------------------------------------------------------------ sketch entry points
static int sketch_from_host(mxe_t* e, const uint8_t* seq, uint64_t n, const uint64_t* offsets, uint32_t n_contigs,
                            int k, int w, int flags, mxe_sketch* S)
{
    ArenaScope scope(e);
    if (n == 0) return sketch_device_impl(e, nullptr, 0, offsets, n_contigs, k, w, flags, S);
    // a copy started by mxe_prefetch_buffers for exactly this buffer, else start one now in a free slot
    H2DSlot* sl = nullptr;
    for (auto& c : e->slot) if (c.h == seq && c.n == n) sl = &c;
    if (!sl) {
        for (auto& c : e->slot) if (!c.h) { sl = &c; break; }
        if (!sl) { sl = &e->slot[0]; cudaStreamSynchronize(e->copy_stream[0]); cudaStreamSynchronize(e->copy_stream[1]); }   // drop a stale prefetch
        MXE_TRY(h2d_issue(e, *sl, seq, n));
    }
    int rc = sketch_device_impl(e, sl->d, n, offsets, n_contigs, k, w, flags, S, sl);   // chunk-wise pack as the copy lands
    sl->h = nullptr;
    return rc;
}

static void set_names(mxe_sketch* S, const char* const* names, uint32_t n_contigs)
{
    S->names.resize(n_contigs);
    for (uint32_t c = 0; c < n_contigs; c++) S->names[c] = names ? names[c] : std::to_string(c);
}

int mxe_sketch_buffers(mxe_t* e, const uint8_t* seq, const uint64_t* offsets, uint32_t n_contigs,
                       const char* const* names, int k, int w, int flags, mxe_sketch_t** out)
{
    if (!e || !out || !offsets || (!seq && n_contigs && offsets[n_contigs])) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(e->device));
    mxe_sketch* S = new mxe_sketch();
    set_names(S, names, n_contigs);
    S->seq_borrowed = seq;
    int rc = sketch_from_host(e, seq, n_contigs ? offsets[n_contigs] : 0, offsets, n_contigs, k, w, flags, S);
    if (rc != MXE_OK) { mxe_sketch_free(S); return rc; }
    *out = S;
    return MXE_OK;
}

int mxe_prefetch_buffers(mxe_t* e, const uint8_t* seq, uint64_t n)
{
    if (!e || (!seq && n)) { set_error("null argument"); return MXE_ERR_ARG; }
    if (n == 0) return MXE_OK;
    MXE_CUDA(cudaSetDevice(e->device));
    for (auto& c : e->slot) if (c.h == seq && c.n == n) return MXE_OK;      // already in flight
    for (auto& c : e->slot)
        if (!c.h) return h2d_issue(e, c, seq, n);
    return MXE_OK;                                                           // both slots busy: the sketch call will copy
}

int mxe_sketch_device(mxe_t* e, const void* d_seq, const uint64_t* offsets, uint32_t n_contigs,
                      const char* const* names, int k, int w, int flags, mxe_sketch_t** out)
{
    if (!e || !out || !offsets || (!d_seq && n_contigs && offsets[n_contigs])) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(e->device));
    mxe_sketch* S = new mxe_sketch();
    set_names(S, names, n_contigs);
    int rc;
    {
        ArenaScope scope(e);
        rc = sketch_device_impl(e, (const uint8_t*)d_seq, n_contigs ? offsets[n_contigs] : 0, offsets, n_contigs, k, w, flags, S);
    }
    if (rc != MXE_OK) { mxe_sketch_free(S); return rc; }
    *out = S;
    return MXE_OK;
}

int mxe_sketch_device_many(mxe_t* e, int n_asm, const void* const* d_seq, const uint64_t* const* offsets, const uint32_t* n_contigs,
                           int k, int w, int flags, mxe_sketch_t** out)
{
    if (!e || !d_seq || !offsets || !n_contigs || !out || n_asm < 1 || n_asm > 32) { set_error("bad arguments"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(e->device));
    std::vector<mxe_sketch*> S((size_t)n_asm);
    for (int a = 0; a < n_asm; a++) { S[a] = new mxe_sketch(); set_names(S[a], nullptr, n_contigs[a]); }
    int rc;
    {
        ArenaScope scope(e);
        rc = sketch_device_many_impl(e, n_asm, (const uint8_t* const*)d_seq, offsets, n_contigs, k, w, flags, S.data());
    }
    if (rc != MXE_OK) { for (auto* s : S) mxe_sketch_free(s); return rc; }
    for (int a = 0; a < n_asm; a++) out[a] = S[a];
    return MXE_OK;
}

// host-only face of the reader (no engine, no GPU): what btllib.SeqReader gives ntJoin (bin/ntjoin_assemble.py:313-316)
struct mxe_fasta {
    mxe::HostText seq;
    std::vector<uint64_t> offsets;
    std::vector<std::string> names;
};

int mxe_fasta_read(const char* path, mxe_fasta_t** out)
{
    if (!path || !out) { set_error("null argument"); return MXE_ERR_ARG; }
    mxe_fasta* F = new mxe_fasta();
    int rc = read_fasta(path, F->seq, F->offsets, F->names);
    if (rc != MXE_OK) { delete F; return rc; }
    *out = F;
    return MXE_OK;
}

int mxe_fasta_view(mxe_fasta_t* F, uint32_t* n_records, const uint64_t** offsets, const char** seq)
{
    if (!F) { set_error("null argument"); return MXE_ERR_ARG; }
    if (n_records) *n_records = (uint32_t)F->names.size();
    if (offsets) *offsets = F->offsets.data();
    if (seq) *seq = F->seq.data();
    return MXE_OK;
}

int mxe_fasta_name(mxe_fasta_t* F, uint32_t idx, const char** name)
{
    if (!F || !name || idx >= F->names.size()) { set_error("bad record index"); return MXE_ERR_ARG; }
    *name = F->names[idx].c_str();
    return MXE_OK;
}

void mxe_fasta_free(mxe_fasta_t* F) { delete F; }

int mxe_sketch_file(mxe_t* e, const char* fasta_path, int k, int w, int flags, mxe_sketch_t** out)
{
    if (!e || !out || !fasta_path) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(e->device));
    mxe_sketch* S = new mxe_sketch();
    std::vector<uint64_t> offsets;
    int rc = read_fasta(fasta_path, S->seq_text, offsets, S->names);
    if (rc == MXE_OK)
        rc = sketch_from_host(e, (const uint8_t*)S->seq_text.data(), offsets.back(), offsets.data(), (uint32_t)S->names.size(), k, w, flags, S);
    if (rc != MXE_OK) { mxe_sketch_free(S); return rc; }
    *out = S;
    return MXE_OK;
}

int mxe_sketch_load_tsv(mxe_t* e, const char* tsv_path, mxe_sketch_t** out)
{
    if (!e || !tsv_path || !out) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(e->device));
    FILE* f = fopen(tsv_path, "rb");
    if (!f) { set_error("cannot open %s: %s", tsv_path, strerror(errno)); return MXE_ERR_IO; }
    fseek(f, 0, SEEK_END);
    long long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)sz + 1);
    if (sz && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); set_error("short read on %s", tsv_path); return MXE_ERR_IO; }
    fclose(f);
    buf[sz] = '\n';
    std::vector<uint64_t> hashes;
    std::vector<uint32_t> pos, contig;
    mxe_sketch* S = new mxe_sketch();
    S->eng = e;
    const char* p = buf.data();
    const char* end = p + sz;
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', end - p + 1);
        const char* tab = (const char*)memchr(p, '\t', nl - p);
        const char* idend = tab ? tab : nl;
        while (idend > p && (idend[-1] == '\r' || idend[-1] == ' ')) idend--;
        uint32_t c = (uint32_t)S->names.size();
        S->names.emplace_back(p, idend - p);
        if (tab) {
            const char* q = tab + 1;
            while (q < nl) {
                while (q < nl && (*q == ' ' || *q == '\r')) q++;
                if (q >= nl) break;
                uint64_t h = 0;
                const char* q0 = q;
                while (q < nl && *q >= '0' && *q <= '9') h = h * 10 + (uint64_t)(*q++ - '0');
                if (q == q0) { mxe_sketch_free(S); set_error("%s: malformed minimizer entry in record %u", tsv_path, c); return MXE_ERR_IO; }
                uint64_t ps = 0;
                if (q < nl && *q == ':') { q++; while (q < nl && *q >= '0' && *q <= '9') ps = ps * 10 + (uint64_t)(*q++ - '0'); }
                while (q < nl && *q != ' ') q++;   // :strand / :seq fields
                hashes.push_back(h); pos.push_back((uint32_t)ps); contig.push_back(c);
            }
        }
        p = nl + 1;
    }
    S->n_contigs = (uint32_t)S->names.size();
    S->offsets.assign(S->n_contigs + 1, 0);
    S->n = hashes.size();
    if (S->n) {
        cudaStream_t st = e->stream;
        size_t n = S->n;
        cudaError_t err = cudaMallocAsync((void**)&S->d_out_hash, n * 8, st);
        if (err == cudaSuccess) err = cudaMallocAsync((void**)&S->d_min_hash, n * 8, st);
        if (err == cudaSuccess) err = cudaMallocAsync((void**)&S->d_pos, n * 4, st);
        if (err == cudaSuccess) err = cudaMallocAsync((void**)&S->d_contig, n * 4, st);
        if (err == cudaSuccess) err = cudaMallocAsync((void**)&S->d_forward, n, st);
        if (err == cudaSuccess) err = cudaMemcpyAsync(S->d_out_hash, hashes.data(), n * 8, cudaMemcpyHostToDevice, st);
        if (err == cudaSuccess) err = cudaMemsetAsync(S->d_min_hash, 0, n * 8, st);
        if (err == cudaSuccess) err = cudaMemcpyAsync(S->d_pos, pos.data(), n * 4, cudaMemcpyHostToDevice, st);
        if (err == cudaSuccess) err = cudaMemcpyAsync(S->d_contig, contig.data(), n * 4, cudaMemcpyHostToDevice, st);
        if (err == cudaSuccess) err = cudaMemsetAsync(S->d_forward, 0, n, st);
        if (err == cudaSuccess) err = cudaStreamSynchronize(st);
        if (err != cudaSuccess) { mxe_sketch_free(S); set_error("upload of %s failed: %s", tsv_path, cudaGetErrorString(err)); return MXE_ERR_CUDA; }
    }
    *out = S;
    return MXE_OK;
}

// pinned block behind the host view of a sketch
static int host_block(mxe_sketch* S)
{
    mxe_engine* e = S->eng;
    size_t n = S->n;
    size_t bytes = n * (8 + 8 + 4 + 4 + 1) + 64;
    char* blk = (char*)e->pinned_alloc(bytes);
    if (!blk) { set_error("pinned host allocation of %zu bytes failed", bytes); return MXE_ERR_NOMEM; }
    S->h_block = blk; S->h_bytes = bytes;
    S->h_out_hash = (uint64_t*)blk;
    S->h_min_hash = (uint64_t*)(blk + 8 * n);
    S->h_pos = (uint32_t*)(blk + 16 * n);
    S->h_contig = (uint32_t*)(blk + 20 * n);
    S->h_forward = (uint8_t*)(blk + 24 * n);
    return MXE_OK;
}

static int host_copies(mxe_sketch* S, cudaStream_t st)
{
    const size_t n = S->n;
    MXE_CUDA(cudaMemcpyAsync(S->h_out_hash, S->d_out_hash, 8 * n, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaMemcpyAsync(S->h_min_hash, S->d_min_hash, 8 * n, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaMemcpyAsync(S->h_pos, S->d_pos, 4 * n, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaMemcpyAsync(S->h_contig, S->d_contig, 4 * n, cudaMemcpyDeviceToHost, st));
    MXE_CUDA(cudaMemcpyAsync(S->h_forward, S->d_forward, n, cudaMemcpyDeviceToHost, st));
    return MXE_OK;
}

static int ensure_host(mxe_sketch* S)
{
    if (S->h_block) {
        if (S->h_pending) {                        // started by mxe_sketch_prefetch_host: wait for it, nothing else
            MXE_CUDA(cudaEventSynchronize(S->h_ready));
            S->h_pending = false;
        }
        return MXE_OK;
    }
    if (S->n == 0) return MXE_OK;
    mxe_engine* e = S->eng;
    MXE_CUDA(cudaSetDevice(e->device));
    MXE_TRY(host_block(S));
    MXE_TRY(host_copies(S, e->stream));
    MXE_CUDA(cudaStreamSynchronize(e->stream));
    return MXE_OK;
}

int mxe_sketch_prefetch_host(mxe_sketch_t* S)
{
    if (!S) { set_error("null sketch"); return MXE_ERR_ARG; }
    if (!S->eng || S->h_block || S->n == 0) return MXE_OK;      // host-only sketch, copy already there or in flight, nothing to copy
    mxe_engine* e = S->eng;
    MXE_CUDA(cudaSetDevice(e->device));
    if (!e->d2h_stream) MXE_CUDA(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
    if (!e->d2h_event) MXE_CUDA(cudaEventCreateWithFlags(&e->d2h_event, cudaEventDisableTiming));
    if (!S->h_ready) MXE_CUDA(cudaEventCreateWithFlags(&S->h_ready, cudaEventDisableTiming));
    MXE_TRY(host_block(S));
    // the arrays were produced in engine-stream order; the copies run on their own stream, beside whatever the engine
    // stream and the host->device copy streams do next (PCIe is full duplex)
    MXE_CUDA(cudaEventRecord(e->d2h_event, e->stream));
    MXE_CUDA(cudaStreamWaitEvent(e->d2h_stream, e->d2h_event, 0));
    MXE_TRY(host_copies(S, e->d2h_stream));
    MXE_CUDA(cudaEventRecord(S->h_ready, e->d2h_stream));
    S->h_pending = true;
    return MXE_OK;
}

int mxe_sketch_view(mxe_sketch_t* S, uint64_t* n, const uint64_t** out_hash, const uint64_t** min_hash,
                    const uint32_t** pos, const uint32_t** contig, const uint8_t** forward)
{
    if (!S) { set_error("null sketch"); return MXE_ERR_ARG; }
    MXE_TRY(ensure_host(S));
    if (n) *n = S->n;
    if (out_hash) *out_hash = S->h_out_hash;
    if (min_hash) *min_hash = S->h_min_hash;
    if (pos) *pos = S->h_pos;
    if (contig) *contig = S->h_contig;
    if (forward) *forward = S->h_forward;
    return MXE_OK;
}

int mxe_sketch_device_view(mxe_sketch_t* S, uint64_t* n, const void** d_out_hash, const void** d_pos, const void** d_contig)
{
    if (!S) { set_error("null sketch"); return MXE_ERR_ARG; }
    if (n) *n = S->n;
    if (d_out_hash) *d_out_hash = S->d_out_hash;
    if (d_pos) *d_pos = S->d_pos;
    if (d_contig) *d_contig = S->d_contig;
    return MXE_OK;
}

// index of the first minimizer whose record id is >= `record` (one thread, log2(n) dependent loads)
__global__ void record_start_kernel(const uint32_t* __restrict__ contig, uint64_t n, uint32_t record, uint64_t* __restrict__ out)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (contig[mid] < record) lo = mid + 1; else hi = mid;
    }
    *out = lo;
}

int mxe_sketch_record_start(mxe_sketch_t* s, uint32_t record, uint64_t* index)
{
    if (!s || !index) { set_error("null argument"); return MXE_ERR_ARG; }
    if (!s->eng || !s->n) { *index = 0; return MXE_OK; }
    mxe_engine* e = s->eng;
    MXE_CUDA(cudaSetDevice(e->device));
    uint64_t* d = nullptr;
    MXE_CUDA(cudaMallocAsync((void**)&d, 8, e->stream));
    record_start_kernel<<<1, 1, 0, e->stream>>>(s->d_contig, s->n, record, d);
    e->launches++;
    MXE_CUDA(cudaMemcpyAsync(index, d, 8, cudaMemcpyDeviceToHost, e->stream));
    MXE_CUDA(cudaFreeAsync(d, e->stream));
    MXE_CUDA(cudaStreamSynchronize(e->stream));
    return MXE_OK;
}

int mxe_sketch_contig_name(mxe_sketch_t* S, uint32_t idx, const char** name)
{
    if (!S || !name || idx >= S->names.size()) { set_error("bad record index"); return MXE_ERR_ARG; }
    *name = S->names[idx].c_str();
    return MXE_OK;
}

int mxe_sketch_counts(mxe_sketch_t* S, uint64_t* n_bases, uint64_t* n_valid_kmers, uint64_t* n_candidates,
                      uint64_t* n_gap_windows, uint32_t* n_contigs)
{
    if (!S) { set_error("null sketch"); return MXE_ERR_ARG; }
    if (n_bases) *n_bases = S->n_bases;
    if (n_valid_kmers) *n_valid_kmers = S->n_valid;
    if (n_candidates) *n_candidates = S->n_cand;
    if (n_gap_windows) *n_gap_windows = S->n_gap_windows;
    if (n_contigs) *n_contigs = S->n_contigs;
    return MXE_OK;
}

int mxe_write_tsv(mxe_sketch_t* S, const char* path, int with_pos, int with_strand, int with_seq)
{
    if (!S || !path) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_TRY(ensure_host(S));
    const char* text = !S->seq_text.empty() ? S->seq_text.data() : (const char*)S->seq_borrowed;
    if (with_seq && !text && S->n) { set_error("sequence text not available for --seq output"); return MXE_ERR_ARG; }
    FILE* f = (path[0] == '-' && !path[1]) ? stdout : fopen(path, "wb");
    if (!f) { set_error("cannot open %s: %s", path, strerror(errno)); return MXE_ERR_IO; }
    int rc = write_tsv_text(S, text, f, with_pos, with_strand, with_seq);
    bool ok = rc == MXE_OK;
    if (f != stdout) ok = (fclose(f) == 0) && ok; else fflush(f);
    if (!ok) { set_error("write to %s failed", path); return MXE_ERR_IO; }
    return MXE_OK;
}

// Host only: a sketch object over caller-provided arrays (copied), e.g. the minimizers of one assembly gathered from
// several GPUs, so that mxe_write_tsv / mxe_sketch_view serve them like a sketch computed here.
int mxe_sketch_from_arrays(const uint64_t* out_hash, const uint64_t* min_hash, const uint32_t* pos, const uint32_t* contig,
                           const uint8_t* forward, uint64_t n, const char* const* names, const uint64_t* offsets,
                           uint32_t n_contigs, int k, const uint8_t* seq, mxe_sketch_t** out)
{
    if (!out || (n && (!out_hash || !pos || !contig)) || (n_contigs && !names)) { set_error("null argument"); return MXE_ERR_ARG; }
    for (uint64_t i = 0; i < n; i++)
        if (contig[i] >= n_contigs || (i && contig[i] < contig[i - 1])) { set_error("record indices must be ascending and below n_contigs"); return MXE_ERR_ARG; }
    mxe_sketch* S = new mxe_sketch();
    S->k = k; S->n = n; S->n_contigs = n_contigs;
    S->names.assign(names, names + n_contigs);
    if (offsets) S->offsets.assign(offsets, offsets + n_contigs + 1); else S->offsets.assign((size_t)n_contigs + 1, 0);
    S->seq_borrowed = offsets ? seq : nullptr;
    const size_t bytes = n * (8 + 8 + 4 + 4 + 1) + 64;
    char* blk = (char*)malloc(bytes);
    if (!blk) { delete S; set_error("host allocation of %zu bytes failed", bytes); return MXE_ERR_NOMEM; }
    S->h_block = blk; S->h_bytes = bytes;
    S->h_out_hash = (uint64_t*)blk;
    S->h_min_hash = (uint64_t*)(blk + 8 * n);
    S->h_pos = (uint32_t*)(blk + 16 * n);
    S->h_contig = (uint32_t*)(blk + 20 * n);
    S->h_forward = (uint8_t*)(blk + 24 * n);
    if (n) {
        memcpy(S->h_out_hash, out_hash, 8 * n);
        if (min_hash) memcpy(S->h_min_hash, min_hash, 8 * n); else memset(S->h_min_hash, 0, 8 * n);
        memcpy(S->h_pos, pos, 4 * n);
        memcpy(S->h_contig, contig, 4 * n);
        if (forward) memcpy(S->h_forward, forward, n); else memset(S->h_forward, 0, n);
    }
    *out = S;
    return MXE_OK;
}

void mxe_sketch_free(mxe_sketch_t* S)
{
    if (!S) return;
    if (S->eng && !engine_alive(S->eng)) { delete S; return; }
    if (S->eng) {
        cudaSetDevice(S->eng->device);
        cudaStream_t st = S->eng->stream;
        if (S->h_pending) cudaEventSynchronize(S->h_ready);      // the pinned block and the arrays are still being read
        if (S->h_ready) cudaEventDestroy(S->h_ready);
        if (S->d_out_hash) cudaFreeAsync(S->d_out_hash, st);
        if (S->d_min_hash) cudaFreeAsync(S->d_min_hash, st);
        if (S->d_pos) cudaFreeAsync(S->d_pos, st);
        if (S->d_contig) cudaFreeAsync(S->d_contig, st);
        if (S->d_forward) cudaFreeAsync(S->d_forward, st);
        S->eng->pinned_release(S->h_block, S->h_bytes);
    } else {
        free(S->h_block);              // mxe_sketch_from_arrays
    }
    delete S;
}

// ------------------------------------------------------------------ .mx.dot from arrays (host only)
// Text format of Ntjoin.print_graph (bin/ntjoin.py:25-67):
//   graph G {
//   "<mx>" [label="<mx>\n<asm key>_<(ctg, pos)>\n..."]         one label line per assembly, literal newlines
//   "<u>" --"<v>" [weight=<float> color=<colour>]
//   }
int mxe_write_dot(const char* path, uint64_t n_v, const uint64_t* vertices, int n_asm, const char* const* asm_keys,
                  const char* const* const* ctg_reprs, const uint32_t* const* v_ctg, const uint32_t* const* v_pos,
                  uint64_t n_e, const uint32_t* e_src, const uint32_t* e_dst, const uint32_t* e_attr,
                  const char* const* attr_text)
{
    if (!path || (n_v && !vertices) || n_asm < 0 || (n_asm && (!asm_keys || !ctg_reprs || !v_ctg || !v_pos)) ||
        (n_e && (!e_src || !e_dst || !e_attr || !attr_text))) { set_error("null argument"); return MXE_ERR_ARG; }
    FILE* f = fopen(path, "wb");
    if (!f) { set_error("cannot open %s: %s", path, strerror(errno)); return MXE_ERR_IO; }
    std::vector<char> buf(1 << 22);
    char* p = buf.data();
    bool ok = true;
    auto flush = [&]() { ok = ok && fwrite(buf.data(), 1, p - buf.data(), f) == (size_t)(p - buf.data()); p = buf.data(); };
    auto put_str = [&](const char* t) {
        size_t len = strlen(t);
        if (len > buf.size() / 2) { flush(); ok = ok && fwrite(t, 1, len, f) == len; return; }
        if ((size_t)(buf.data() + buf.size() - p) < len + 64) flush();
        memcpy(p, t, len); p += len;
    };
    put_str("graph G {\n");
    std::vector<size_t> key_len(n_asm);
    for (uint64_t i = 0; i < n_v && ok; i++) {
        if ((size_t)(buf.data() + buf.size() - p) < 128) flush();
        *p++ = '"'; p = put_u64(p, vertices[i]); memcpy(p, "\" [label=\"", 10); p += 10; p = put_u64(p, vertices[i]);
        for (int a = 0; a < n_asm; a++) {
            *p++ = '\n';
            put_str(asm_keys[a]);
            *p++ = '_'; *p++ = '(';
            put_str(ctg_reprs[a][v_ctg[a][i]]);
            if ((size_t)(buf.data() + buf.size() - p) < 64) flush();
            *p++ = ','; *p++ = ' '; p = put_u64(p, v_pos[a][i]); *p++ = ')';
        }
        *p++ = '"'; *p++ = ']'; *p++ = '\n';
    }
    for (uint64_t t = 0; t < n_e && ok; t++) {
        if ((size_t)(buf.data() + buf.size() - p) < 128) flush();
        if (e_src[t] >= n_v || e_dst[t] >= n_v) { ok = false; set_error("edge %llu refers to a vertex out of range", (unsigned long long)t); fclose(f); return MXE_ERR_ARG; }
        *p++ = '"'; p = put_u64(p, vertices[e_src[t]]); memcpy(p, "\" --\"", 5); p += 5; p = put_u64(p, vertices[e_dst[t]]); *p++ = '"';
        put_str(attr_text[e_attr[t]]);
    }
    put_str("}\n");
    flush();
    ok = (fclose(f) == 0) && ok;
    if (!ok) { set_error("write to %s failed", path); return MXE_ERR_IO; }
    return MXE_OK;
}

// ------------------------------------------------------------------ steps 2-3
int mxe_filter_and_edges_device(mxe_t* e, const void* const* d_hash, const void* const* d_contig,
                                const uint64_t* n, int n_asm, const double* weights, mxe_result_t** out)
{
    if (!e || !d_hash || !d_contig || !n || !weights || !out) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(e->device));
    if (e->filter_variant >= 1 && n_asm >= 1 && n_asm <= 32) {
        // hash buckets in shared memory (csrc/p2p.cu with world = 1); a bucket overflow falls back to the global sort
        uint64_t N = 0;
        for (int a = 0; a < n_asm; a++) N += n[a];
        if (N > 0 && N < (1ULL << 32) - 8192) {
            if (!e->local_p2p || e->local_p2p_cap < N || e->local_p2p_asm < n_asm) {
                if (e->local_p2p) { mxe_p2p_free(e->local_p2p); e->local_p2p = nullptr; }
                const uint64_t cap = std::min<uint64_t>(N + N / 8 + 65536, (1ULL << 32) - 1);
                int rc0 = mxe_p2p_create(e, 0, 1, cap, std::max(n_asm, 4), &e->local_p2p);
                if (rc0 != MXE_OK) return rc0;
                e->local_p2p_cap = cap; e->local_p2p_asm = std::max(n_asm, 4);
            }
            mxe_p2p* X = e->local_p2p;
            int rc1 = mxe_p2p_scatter(X, d_hash, d_contig, n, n_asm, weights);
            if (rc1 == MXE_OK) rc1 = mxe_p2p_buckets(X);
            if (rc1 == MXE_OK) rc1 = mxe_p2p_adjacency(X);
            if (rc1 == MXE_OK) rc1 = mxe_p2p_edges(X);
            if (rc1 == MXE_OK) rc1 = mxe_p2p_finish(X, out);
            if (rc1 == MXE_OK) return MXE_OK;
            if (rc1 != MXE_ERR_INTERNAL) return rc1;
        }
    }
    mxe_result* R = new mxe_result();
    int rc;
    {
        ArenaScope scope(e);
        rc = filter_and_edges_impl(e, (const uint64_t* const*)d_hash, (const uint32_t* const*)d_contig, n, n_asm, weights, R);
    }
    if (rc != MXE_OK) { mxe_result_free(R); return rc; }
    *out = R;
    return MXE_OK;
}

int mxe_filter_and_edges(mxe_t* e, mxe_sketch_t* const* sketches, int n_asm, const double* weights, mxe_result_t** out)
{
    if (!e || !sketches || !weights || !out) { set_error("null argument"); return MXE_ERR_ARG; }
    if (n_asm < 1 || n_asm > 32) { set_error("n_asm must be 1..32"); return MXE_ERR_ARG; }
    const void* dh[32]; const void* dc[32]; uint64_t n[32];
    for (int a = 0; a < n_asm; a++) {
        if (!sketches[a] || sketches[a]->eng != e) { set_error("sketch %d does not belong to this engine", a); return MXE_ERR_ARG; }
        dh[a] = sketches[a]->d_out_hash; dc[a] = sketches[a]->d_contig; n[a] = sketches[a]->n;
    }
    return mxe_filter_and_edges_device(e, dh, dc, n, n_asm, weights, out);
}

static int result_ensure_host(mxe_result* R)
{
    if (R->h_block) return MXE_OK;
    mxe_engine* e = R->eng;
    MXE_CUDA(cudaSetDevice(e->device));
    size_t bytes = 2 * R->N + 8 * R->nV + R->nE * (8 + 8 + 4 + 8 + 8) + 256;
    char* blk = (char*)e->pinned_alloc(bytes);
    if (!blk) { set_error("pinned host allocation of %zu bytes failed", bytes); return MXE_ERR_NOMEM; }
    R->h_block = blk; R->h_bytes = bytes;
    char* p = blk;
    R->h_vertices = (uint64_t*)p; p += 8 * R->nV;
    R->h_eu = (uint64_t*)p; p += 8 * R->nE;
    R->h_ev = (uint64_t*)p; p += 8 * R->nE;
    R->h_ew = (double*)p; p += 8 * R->nE;
    R->h_ekey = (uint64_t*)p; p += 8 * R->nE;
    R->h_emask = (uint32_t*)p; p += 4 * R->nE;
    R->h_uniq = (uint8_t*)p; p += R->N;
    R->h_keep = (uint8_t*)p;
    cudaStream_t st = e->stream;
    if (R->N && R->d_uniq) {
        MXE_CUDA(cudaMemcpyAsync(R->h_uniq, R->d_uniq, R->N, cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaMemcpyAsync(R->h_keep, R->d_keep, R->N, cudaMemcpyDeviceToHost, st));
    }
    if (R->nV) MXE_CUDA(cudaMemcpyAsync(R->h_vertices, R->d_vertices, 8 * R->nV, cudaMemcpyDeviceToHost, st));
    if (R->nE) {
        MXE_CUDA(cudaMemcpyAsync(R->h_eu, R->d_eu, 8 * R->nE, cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaMemcpyAsync(R->h_ev, R->d_ev, 8 * R->nE, cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaMemcpyAsync(R->h_ew, R->d_ew, 8 * R->nE, cudaMemcpyDeviceToHost, st));
        MXE_CUDA(cudaMemcpyAsync(R->h_emask, R->d_emask, 4 * R->nE, cudaMemcpyDeviceToHost, st));
        if (R->d_ekey) MXE_CUDA(cudaMemcpyAsync(R->h_ekey, R->d_ekey, 8 * R->nE, cudaMemcpyDeviceToHost, st));
    }
    MXE_CUDA(cudaStreamSynchronize(st));
    return MXE_OK;
}

int mxe_result_counts(mxe_result_t* r, uint64_t* n_minimizers, uint64_t* n_vertices, uint64_t* n_edges)
{
    if (!r) { set_error("null result"); return MXE_ERR_ARG; }
    if (r->eng) MXE_CUDA(cudaStreamSynchronize(r->eng->stream));
    if (n_minimizers) *n_minimizers = r->N;
    if (n_vertices) *n_vertices = r->nV;
    if (n_edges) *n_edges = r->nE;
    return MXE_OK;
}

int mxe_result_flags(mxe_result_t* r, int a, uint64_t* n, const uint8_t** uniq, const uint8_t** keep)
{
    if (!r || a < 0 || a >= r->n_asm) { set_error("bad assembly index"); return MXE_ERR_ARG; }
    MXE_TRY(result_ensure_host(r));
    if (n) *n = r->asm_off[a + 1] - r->asm_off[a];
    if (uniq) *uniq = r->h_uniq + r->asm_off[a];
    if (keep) *keep = r->h_keep + r->asm_off[a];
    return MXE_OK;
}

int mxe_result_graph(mxe_result_t* r, uint64_t* n_vertices, const uint64_t** vertices, uint64_t* n_edges,
                     const uint64_t** edge_u, const uint64_t** edge_v, const uint32_t** support_mask, const double** weight)
{
    if (!r) { set_error("null result"); return MXE_ERR_ARG; }
    MXE_TRY(result_ensure_host(r));
    if (n_vertices) *n_vertices = r->nV;
    if (vertices) *vertices = r->h_vertices;
    if (n_edges) *n_edges = r->nE;
    if (edge_u) *edge_u = r->h_eu;
    if (edge_v) *edge_v = r->h_ev;
    if (support_mask) *support_mask = r->h_emask;
    if (weight) *weight = r->h_ew;
    return MXE_OK;
}

void mxe_result_free(mxe_result_t* r)
{
    if (!r) return;
    if (r->eng && !engine_alive(r->eng)) { delete r; return; }
    if (r->eng) {
        cudaSetDevice(r->eng->device);
        cudaStream_t st = r->eng->stream;
        void* ptrs[] = {r->d_uniq, r->d_keep, r->d_vertices, r->d_eu, r->d_ev, r->d_emask, r->d_ew, r->d_ekey};
        for (void* p : ptrs) if (p) cudaFreeAsync(p, st);
        r->eng->pinned_release(r->h_block, r->h_bytes);
    }
    delete r;
}

int mxe_result_edge_keys(mxe_result_t* r, uint64_t* n_edges, const uint64_t** keys)
{
    if (!r) { set_error("null result"); return MXE_ERR_ARG; }
    if (r->nE && !r->d_ekey) { set_error("edge order keys exist only on multi-GPU result shards"); return MXE_ERR_ARG; }
    MXE_TRY(result_ensure_host(r));
    if (n_edges) *n_edges = r->nE;
    if (keys) *keys = r->h_ekey;
    return MXE_OK;
}

// ------------------------------------------------------------------ multi-GPU steps 2-3 (stages between the caller's collectives)
int mxe_dist_mark(mxe_t* e, const void* d_keys, const uint64_t* asm_off, int n_asm, int rank, int world,
                  void* d_mk, mxe_dist_t** out, uint64_t* n_vertices_local)
{
    if (!e || !asm_off || !out || !n_vertices_local || !d_mk || (!d_keys && n_asm > 0 && asm_off[n_asm])) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(e->device));
    mxe_dist* X = new mxe_dist();
    X->eng = e;
    int rc;
    {
        ArenaScope scope(e);
        rc = dist_mark_impl(e, (const uint64_t*)d_keys, asm_off, n_asm, rank, world, (uint32_t*)d_mk, X, n_vertices_local);
    }
    if (rc != MXE_OK) { mxe_dist_free(X); return rc; }
    *out = X;
    return MXE_OK;
}

int mxe_dist_adjacency(mxe_dist_t* X, const void* d_mk, const uint64_t* vbase, const uint64_t* loc_off, const uint64_t* loc_n,
                       const void* const* d_contig, void* d_succ)
{
    if (!X || !d_mk || !vbase || !loc_off || !loc_n || !d_contig || !d_succ) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(X->eng->device));
    ArenaScope scope(X->eng);
    return dist_adjacency_impl(X, (const uint32_t*)d_mk, vbase, loc_off, loc_n, (const uint32_t* const*)d_contig, (uint32_t*)d_succ);
}

int mxe_dist_edges(mxe_dist_t* X, const void* d_succ, void* d_srcmin, uint64_t* n_edges_local)
{
    if (!X || !d_succ || !d_srcmin || !n_edges_local) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(X->eng->device));
    ArenaScope scope(X->eng);
    return dist_edges_impl(X, (const uint32_t*)d_succ, (uint32_t*)d_srcmin, n_edges_local);
}

int mxe_dist_finish(mxe_dist_t* X, const void* d_srcmin, const double* weights, mxe_result_t** out)
{
    if (!X || !d_srcmin || !weights || !out) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(X->eng->device));
    mxe_result* R = new mxe_result();
    int rc;
    {
        ArenaScope scope(X->eng);
        rc = dist_finish_impl(X, (const uint32_t*)d_srcmin, weights, R);
    }
    if (rc != MXE_OK) { mxe_result_free(R); return rc; }
    *out = R;
    return MXE_OK;
}

void mxe_dist_free(mxe_dist_t* X)
{
    if (!X) return;
    if (X->eng && engine_alive(X->eng)) {
        cudaSetDevice(X->eng->device);
        for (void* p : X->owned) if (p) cudaFreeAsync(p, X->eng->stream);
    }
    delete X;
}

// ------------------------------------------------------------------ multi-GPU steps 2-3, all-to-all formulation
int mxe_a2a_partition(mxe_t* e, const void* const* d_hash, const uint64_t* n, int n_asm, int rank, int world,
                      mxe_a2a_t** out, uint64_t* counts, const void** d_send_keys)
{
    if (!e || !d_hash || !n || !out || !counts || !d_send_keys) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(e->device));
    mxe_a2a* X = new mxe_a2a();
    X->eng = e;
    int rc;
    {
        ArenaScope scope(e);
        rc = a2a_partition_impl(e, (const uint64_t* const*)d_hash, n, n_asm, rank, world, X, counts, d_send_keys);
    }
    if (rc != MXE_OK) { mxe_a2a_free(X); return rc; }
    *out = X;
    return MXE_OK;
}

int mxe_a2a_mark(mxe_a2a_t* X, const void* d_recv_keys, const uint64_t* recv_counts, void* d_ret_marks, uint64_t* n_vertices_local)
{
    if (!X || !recv_counts || !n_vertices_local) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(X->eng->device));
    ArenaScope scope(X->eng);
    return a2a_mark_impl(X, (const uint64_t*)d_recv_keys, recv_counts, (uint32_t*)d_ret_marks, n_vertices_local);
}

int mxe_a2a_sightings(mxe_a2a_t* X, const void* d_marks, const void* const* d_contig, const uint64_t* goff,
                      uint64_t* rec_counts, const void** d_send_records)
{
    if (!X || !d_contig || !goff || !rec_counts || !d_send_records) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(X->eng->device));
    ArenaScope scope(X->eng);
    return a2a_sightings_impl(X, (const uint32_t*)d_marks, (const uint32_t* const*)d_contig, goff, rec_counts, d_send_records);
}

int mxe_a2a_finish(mxe_a2a_t* X, const void* d_recv_records, uint64_t n_records, uint64_t n_global, const double* weights, mxe_result_t** out)
{
    if (!X || !weights || !out) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(X->eng->device));
    mxe_result* R = new mxe_result();
    int rc;
    {
        ArenaScope scope(X->eng);
        rc = a2a_finish_impl(X, (const uint64_t*)d_recv_records, n_records, n_global, weights, R);
    }
    if (rc != MXE_OK) { mxe_result_free(R); return rc; }
    *out = R;
    return MXE_OK;
}

void mxe_a2a_free(mxe_a2a_t* X)
{
    if (!X) return;
    if (X->eng && engine_alive(X->eng)) {
        cudaSetDevice(X->eng->device);
        for (void* p : X->owned) if (p) cudaFreeAsync(p, X->eng->stream);
    }
    delete X;
}

// ------------------------------------------------------------------ measurement
int mxe_timing(mxe_t* e, const char* name, double* ms, uint64_t* launches)
{
    if (!e || !name) { set_error("null argument"); return MXE_ERR_ARG; }
    MXE_CUDA(cudaSetDevice(e->device));
    MXE_CUDA(cudaStreamSynchronize(e->stream));
    double total = 0; uint64_t nl = 0;
    auto it = e->timers.find(name);
    if (it != e->timers.end()) {
        for (auto& sp : it->second.spans) { float t = 0; cudaEventElapsedTime(&t, sp.first, sp.second); total += t; }
        nl = it->second.launches;
    }
    if (ms) *ms = total;
    if (launches) *launches = nl;
    return MXE_OK;
}

int mxe_timing_reset(mxe_t* e)
{
    if (!e) { set_error("null engine"); return MXE_ERR_ARG; }
    cudaStreamSynchronize(e->stream);
    for (auto& kv : e->timers)
        for (auto& sp : kv.second.spans) { e->event_pool.push_back(sp.first); e->event_pool.push_back(sp.second); }
    e->timers.clear();
    return MXE_OK;
}

uint64_t mxe_kernel_launches(mxe_t* e) { return e ? e->launches : 0; }

}  // extern "C"
