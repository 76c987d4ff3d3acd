// common.cuh -- shared declarations of the B200 minimizer engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>

#include "../../include/mxe.h"

namespace mxe {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);

#define MXE_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::mxe::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return MXE_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define MXE_TRY(call)                \
    do {                             \
        int _r = (call);             \
        if (_r != MXE_OK) return _r; \
    } while (0)

// ---------------------------------------------------------------- ntHash constants
// Base codes used on the device: bits (2,1) of the ASCII byte -> A=0 C=1 T=2 G=3 ; complement = code^2.
constexpr uint64_t SEED_A = 0x3c8bfbb395c60474ULL;
constexpr uint64_t SEED_C = 0x3193c18562a02b4cULL;
constexpr uint64_t SEED_G = 0x20323ed082572324ULL;
constexpr uint64_t SEED_T = 0x295549f54be24456ULL;
constexpr uint64_t MULTISEED = 0x90b45d39fb6da1faULL;
constexpr int MULTISHIFT = 27;

// ---------------------------------------------------------------- engine
struct PhaseTimer {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
    uint64_t launches = 0;
};

// Grow-only device workspace for per-call temporaries: bump allocation, no allocator calls in steady state.
// Chunks are kept until the next call starts; if a call needed more than one chunk they are merged into one
// chunk of the total size, so from the second step of a fixed workload on there is exactly one chunk.
struct Arena {
    struct Chunk { char* p; size_t bytes, used; };
    std::vector<Chunk> chunks;
    bool active = false;
    void* alloc(size_t bytes);          // nullptr on out-of-memory
    void begin(cudaStream_t st);        // start of an API call: recycle (and merge) the chunks
    void end() { active = false; }
    void destroy();
};

struct Engine {
    int device = 0;
    Arena arena;
    cudaStream_t stream = nullptr;      // stream all work is issued on
    cudaStream_t own_stream = nullptr;  // engine-owned default
    int sm_count = 148;
    // tunables
    double tau = 9.0;        // candidate threshold: hash0>>33 <= tau * 2^31 / w
    int chunk = 0;           // positions per thread in the candidate kernel (multiple of 32; 0 = auto)
    int cand_variant = 4;    // 0 = generic 64-bit, 1 = 31-bit lane prefilter, 2/3 = bit-sliced via plane_kernel, 4 = pack2 + bit-sliced scan (scan_kernels.cuh)
    int scan_lw = 0;         // variant 4: pk words per stream (odd; 0 = by size)
    bool prune = false;      // drop dominated candidates on 31-bit bounds before the exact stages
    bool fma_offload = true; // cand31: additions of the threshold test as IMADs on the FMA pipe
    bool select_narrow = true;   // window selection in 32-bit arithmetic when the ordinals allow it
    int filter_variant = 1;  // steps 2-3: 0 = global radix sort (filter.cu), 1 = hash buckets in shared memory (p2p.cu)
    int sort_bits = 0;       // steps 2-3: top hash bits covered by the radix sort (24/32/40; 0 = by size), rest by the fix-up
    double bound_scale = 1.0;   // scales the size bounds of the asynchronous path (tests: < 1 forces the overflow / repeat path)
    int many_streams = 2;    // mxe_sketch_device_many: 2 = assemblies alternate between the engine stream and an auxiliary one, 1 = all on the engine stream (only the host round trips between them go away)
    bool async_sizes = true; // sketch: arrays sized from bounds, counts stay on the device, one host round trip per sketch
    bool timing = false;
    bool timing_fine = false;   // also time every kernel of steps 2-3 and the barrier waits (option timing = 2)
    // accounting
    uint64_t launches = 0;
    std::map<std::string, PhaseTimer> timers;
    std::vector<cudaEvent_t> event_pool;

    cudaEvent_t get_event();
    void span_begin(const char* name, cudaEvent_t* a);
    void span_end(const char* name, cudaEvent_t a, uint64_t n_launch);
};

// RAII span: records CUDA events on the engine stream around a group of launches when timing is on.
struct Span {
    Engine* e; const char* name; cudaEvent_t a = nullptr; uint64_t l0;
    static bool fine(const char* n) { return (n[0] == 'k' && n[1] == '_' && n[2] == 'p' && n[3] == '2') || (n[0] == 'p' && n[1] == '2' && n[2] == 'p' && n[3] == '_' && n[4] == 'w'); }
    Span(Engine* e_, const char* n) : e(e_), name(n), l0(e_->launches) { if (e->timing && (e->timing_fine || !fine(n))) e->span_begin(name, &a); }
    ~Span() { if (a) e->span_end(name, a, e->launches - l0); }
};

Arena* current_arena();                 // arena of the API call running on this thread, or nullptr
struct ArenaScope {
    Engine* e;
    ArenaScope(Engine* e_);
    ~ArenaScope();
};

// device buffer: workspace arena inside an API call, stream-ordered allocation otherwise
template <typename T>
struct DBuf {
    T* p = nullptr; size_t n = 0; cudaStream_t s = nullptr;
    DBuf() {}
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    bool from_arena = false;
    int alloc(size_t count, cudaStream_t st) {
        release();
        s = st; n = count;
        if (count == 0) count = 1;
        if (Arena* a = current_arena()) {
            p = (T*)a->alloc(count * sizeof(T));
            from_arena = true;
            if (!p) { set_error("device workspace allocation of %zu bytes failed", count * sizeof(T)); return MXE_ERR_NOMEM; }
            return MXE_OK;
        }
        from_arena = false;
        cudaError_t err = cudaMallocAsync((void**)&p, count * sizeof(T), st);
        if (err != cudaSuccess) { p = nullptr; set_error("cudaMallocAsync(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(err)); return MXE_ERR_NOMEM; }
        return MXE_OK;
    }
    void release() { if (p) { if (!from_arena) cudaFreeAsync(p, s); p = nullptr; n = 0; } }
    // hand the buffer to a longer-lived owner: arena memory is copied out into a stream-ordered allocation
    T* detach() {
        T* q = p;
        if (p && from_arena) {
            q = nullptr;
            if (cudaMallocAsync((void**)&q, (n ? n : 1) * sizeof(T), s) == cudaSuccess)
                cudaMemcpyAsync(q, p, n * sizeof(T), cudaMemcpyDeviceToDevice, s);
        }
        p = nullptr; n = 0;
        return q;
    }
    ~DBuf() { release(); }
};

#define MXE_LAUNCH(eng, kernel, grid, block, smem, ...)                         \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), (eng)->stream>>>(__VA_ARGS__);        \
        (eng)->launches++;                                                      \
    } while (0)

// ---------------------------------------------------------------- primitives (sort_scan.cu)
// Exclusive scan of n uint32 counts into uint64 prefixes (out[n] = total).  In-stream.
int exclusive_scan_u32_u64(Engine* e, const uint32_t* d_in, uint64_t* d_out, size_t n);
// Stable LSD radix sort of (key u64, value u32) pairs on bits [begin_bit, end_bit).
// Result lands in d_keys/d_vals; the *_alt buffers are scratch of the same size.
int radix_sort_pairs(Engine* e, uint64_t* d_keys, uint32_t* d_vals, uint64_t* d_keys_alt, uint32_t* d_vals_alt,
                     size_t n, int begin_bit, int end_bit);

// Bitmap rank directory: counts per 1024-bit block -> exclusive prefix (n_blocks+1 entries).
constexpr int RANK_BLOCK_BITS = 1024;
constexpr int RANK_BLOCK_WORDS = RANK_BLOCK_BITS / 32;
int bitmap_rank_build(Engine* e, const uint32_t* d_bits, size_t n_words, uint64_t* d_prefix /* n_blocks+1 */);
// Sorted positions of the set bits; d_out holds `cap` entries (bits beyond it are dropped: the caller detects the overflow
// from the prefix total).
int bitmap_extract(Engine* e, const uint32_t* d_bits, size_t n_words, const uint64_t* d_prefix, uint64_t* d_out, uint64_t cap);

}  // namespace mxe
