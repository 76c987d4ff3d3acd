// engine.cuh -- object definitions behind the opaque C handles.
#pragma once
#include "common.cuh"
#include <memory>

namespace mxe {
// allocator that leaves new elements uninitialised: a 3 GB text buffer is not zero-filled before it is overwritten
template <typename T>
struct DefaultInit : std::allocator<T> {
    template <typename U> struct rebind { using other = DefaultInit<U>; };
    template <typename U> void construct(U* p) noexcept { ::new ((void*)p) U; }
    template <typename U, typename... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};
typedef std::vector<char, DefaultInit<char>> HostText;

inline char* put_u64(char* p, uint64_t v)      // decimal digits, returns the end
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
}

// Host-input path: a host buffer is copied in chunks on two copy streams into an engine-owned staging slot; the pack
// kernel of each chunk waits for that chunk's arrival event only.  Two slots, so that the copy of the next assembly
// (mxe_prefetch_buffers) runs while the previous one is still being sketched.
struct H2DSlot {
    uint8_t* d = nullptr; size_t cap = 0;            // device staging buffer (grow-only)
    const uint8_t* h = nullptr; uint64_t n = 0;      // host buffer staged or in flight (nullptr = slot free)
    uint64_t chunk = 0; int n_chunks = 0;
    std::vector<cudaEvent_t> ev;                     // arrival of chunk i
    cudaEvent_t consumed = nullptr;                  // last reader of the slot (pack) has finished
    bool consumed_valid = false;
};

struct mxe_engine : public mxe::Engine {
    cudaStream_t copy_stream[2] = {nullptr, nullptr};
    cudaStream_t aux_stream = nullptr;      // second compute stream: sketches of several assemblies run concurrently
    cudaEvent_t aux_event = nullptr;
    cudaStream_t d2h_stream = nullptr;      // device->host copies that run beside the next assembly's host->device copy (mxe_sketch_prefetch_host)
    cudaEvent_t d2h_event = nullptr;
    H2DSlot slot[2];
    int h2d_chunk_mb = 256;
    // pinned host block pool (grow-only, reused across steps)
    struct Pinned { void* p; size_t bytes; };
    std::vector<Pinned> pinned_free;
    struct mxe_p2p* local_p2p = nullptr;    // world-1 instance behind mxe_filter_and_edges (grow-only capacity)
    uint64_t local_p2p_cap = 0; int local_p2p_asm = 0;
    void* pinned_alloc(size_t bytes);
    void pinned_release(void* p, size_t bytes);
};

struct mxe_sketch {
    mxe_engine* eng = nullptr;
    int k = 0, w = 0, flags = 0;
    uint64_t n_bases = 0, n_valid = 0, n_cand = 0, n_gap_windows = 0, n_gaps = 0;
    uint32_t n_contigs = 0;
    std::vector<std::string> names;
    std::vector<uint64_t> offsets;
    // device-resident result (sorted by global position = (contig, pos))
    uint64_t n = 0;
    uint64_t* d_out_hash = nullptr;
    uint64_t* d_min_hash = nullptr;
    uint32_t* d_pos = nullptr;
    uint32_t* d_contig = nullptr;
    uint8_t* d_forward = nullptr;
    // lazy host copy (one pinned block)
    void* h_block = nullptr; size_t h_bytes = 0;
    uint64_t* h_out_hash = nullptr; uint64_t* h_min_hash = nullptr;
    uint32_t* h_pos = nullptr; uint32_t* h_contig = nullptr; uint8_t* h_forward = nullptr;
    cudaEvent_t h_ready = nullptr; bool h_pending = false;      // host copy in flight on the engine's d2h stream (mxe_sketch_prefetch_host)
    // sequence text for --seq output: either an owned copy of the whole input or nothing
    mxe::HostText seq_text;           // upper-cased concatenated sequence (mxe_sketch_file)
    const uint8_t* seq_borrowed = nullptr;   // caller buffer (mxe_sketch_buffers), valid while caller keeps it
};

namespace mxe {
struct AsmOffsets { uint64_t off[33]; int n; double weight[32]; };
struct DistInfo { uint64_t vbase[33]; int world; };                     // exclusive prefix of the per-rank vertex counts
struct LocalSlices { uint64_t lofs[33]; uint64_t goff[32]; int n; };    // local concat offsets, global index of each local slice
}

struct mxe_result {
    mxe_engine* eng = nullptr;
    int n_asm = 0;
    uint64_t N = 0, nV = 0, nE = 0;
    uint64_t asm_off[33] = {0};
    // device-resident result
    uint8_t* d_uniq = nullptr; uint8_t* d_keep = nullptr;      // N flags, concatenated in assembly order
    uint64_t* d_vertices = nullptr;                             // nV
    uint64_t* d_eu = nullptr; uint64_t* d_ev = nullptr;         // nE
    uint32_t* d_emask = nullptr; double* d_ew = nullptr;        // nE
    uint64_t* d_ekey = nullptr;                                 // nE, multi-GPU shards only: global order key of each edge
    // lazy host copy (one pinned block)
    void* h_block = nullptr; size_t h_bytes = 0;
    uint8_t* h_uniq = nullptr; uint8_t* h_keep = nullptr; uint64_t* h_vertices = nullptr;
    uint64_t* h_eu = nullptr; uint64_t* h_ev = nullptr; uint32_t* h_emask = nullptr; double* h_ew = nullptr;
    uint64_t* h_ekey = nullptr;
};

// multi-GPU steps 2-3: state carried between the stages (stream-ordered allocations: they outlive one API call)
struct mxe_dist {
    mxe_engine* eng = nullptr;
    mxe::AsmOffsets A;
    mxe::LocalSlices S;
    mxe::DistInfo D;
    int rank = 0, world = 1;
    uint64_t N = 0, L = 0, nV_local = 0, nV = 0, n_keep = 0, nE = 0;
    const uint64_t* d_keys = nullptr;      // caller-owned, must stay alive until mxe_dist_finish
    std::vector<void*> owned;              // everything below, released by mxe_dist_free
    uint64_t* vertices = nullptr;          // nV_local, ascending
    uint8_t* luniq = nullptr; uint8_t* lkeep = nullptr;                               // L
    uint32_t *cvid = nullptr, *cidx = nullptr, *cloc = nullptr, *eflag = nullptr;     // n_keep (+1)
    uint32_t *ue_q0 = nullptr, *ue_mask = nullptr;                                    // nE
    template <typename T> int alloc(T** p, size_t n)
    {
        *p = nullptr;
        cudaError_t err = cudaMallocAsync((void**)p, (n ? n : 1) * sizeof(T), eng->stream);
        if (err != cudaSuccess) { mxe::set_error("cudaMallocAsync(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(err)); return MXE_ERR_NOMEM; }
        owned.push_back(*p);
        return MXE_OK;
    }
};

// multi-GPU steps 2-3, all-to-all formulation: state carried between the stages
struct mxe_a2a {
    mxe_engine* eng = nullptr;
    mxe::AsmOffsets A;                     // owner side: assembly offsets of the received (regrouped) keys
    mxe::LocalSlices S;                    // source side: local concat offsets
    int rank = 0, world = 1, n_asm = 0;
    uint64_t L = 0, n_recv = 0, nV_local = 0;
    std::vector<void*> owned;
    uint64_t *lhash = nullptr, *send_keys = nullptr, *vertices = nullptr, *send_rec = nullptr;
    uint32_t* perm = nullptr;
    uint8_t *luniq = nullptr, *lkeep = nullptr;
    template <typename T> int alloc(T** p, size_t n)
    {
        *p = nullptr;
        cudaError_t err = cudaMallocAsync((void**)p, (n ? n : 1) * sizeof(T), eng->stream);
        if (err != cudaSuccess) { mxe::set_error("cudaMallocAsync(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(err)); return MXE_ERR_NOMEM; }
        owned.push_back(*p);
        return MXE_OK;
    }
};

namespace mxe {
bool engine_alive(const void* e);
void p2p_release(struct ::mxe_p2p* X, bool engine_is_alive);
int a2a_partition_impl(mxe_engine* e, const uint64_t* const* d_hash, const uint64_t* n, int n_asm, int rank, int world,
                       mxe_a2a* X, uint64_t* counts, const void** d_send_keys);
int a2a_mark_impl(mxe_a2a* X, const uint64_t* d_recv, const uint64_t* recv_counts, uint32_t* d_ret, uint64_t* nv_local);
int a2a_sightings_impl(mxe_a2a* X, const uint32_t* d_marks, const uint32_t* const* d_contig, const uint64_t* goff,
                       uint64_t* rec_counts, const void** d_send_records);
int a2a_finish_impl(mxe_a2a* X, const uint64_t* d_rec, uint64_t n_rec, uint64_t N_global, const double* weights, mxe_result* out);
int sketch_device_impl(mxe_engine* e, const uint8_t* d_seq, uint64_t n, const uint64_t* offsets, uint32_t n_contigs,
                       int k, int w, int flags, mxe_sketch* out, H2DSlot* staged = nullptr);
int h2d_issue(mxe_engine* e, H2DSlot& s, const uint8_t* h, uint64_t n);
int sketch_device_many_impl(mxe_engine* e, int n_asm, const uint8_t* const* d_seq, const uint64_t* const* offsets, const uint32_t* n_contigs,
                            int k, int w, int flags, mxe_sketch* const* S);
// host side of the file seam (hostio.cu): multi-threaded FASTA/FASTQ ingest and TSV text
int host_threads();
int read_fasta(const char* path, HostText& seq, std::vector<uint64_t>& offsets, std::vector<std::string>& names);
int write_tsv_text(const mxe_sketch* S, const char* text, FILE* f, int with_pos, int with_strand, int with_seq);
int filter_and_edges_impl(mxe_engine* e, const uint64_t* const* d_hash, const uint32_t* const* d_contig,
                          const uint64_t* n, int n_asm, const double* weights, mxe_result* out);
int dist_mark_impl(mxe_engine* e, const uint64_t* d_keys, const uint64_t* asm_off, int n_asm, int rank, int world,
                   uint32_t* d_mk, mxe_dist* X, uint64_t* nv_local);
int dist_adjacency_impl(mxe_dist* X, const uint32_t* d_mk, const uint64_t* vbase, const uint64_t* loc_off, const uint64_t* loc_n,
                        const uint32_t* const* d_contig, uint32_t* d_succ);
int dist_edges_impl(mxe_dist* X, const uint32_t* d_succ, uint32_t* d_srcmin, uint64_t* n_edges_local);
int dist_finish_impl(mxe_dist* X, const uint32_t* d_srcmin, const double* weights, mxe_result* out);
}
