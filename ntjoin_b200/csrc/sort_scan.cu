// sort_scan.cu -- device primitives: exclusive scan, stable LSD radix sort, bitmap rank / extract.
// Hand-written for sm_100a; all launches go to the engine stream.
#include "common.cuh"

namespace mxe {

// ------------------------------------------------------------------------------------------
// block-wide exclusive scan of one u64 per thread (blockDim.x == 256); returns the exclusive
// prefix and leaves the block total in *total (valid for all threads after the call).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t block_exscan_256(uint64_t v, uint64_t* total, uint64_t* smem /* 9 */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) smem[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint64_t s = lane < 8 ? smem[lane] : 0;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += y;
        }
        if (lane < 8) smem[lane] = s;   // inclusive warp totals
    }
    __syncthreads();
    uint64_t base = warp ? smem[warp - 1] : 0;
    *total = smem[7];
    __syncthreads();
    return base + x - v;
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in, size_t n, uint64_t* __restrict__ tile_sums)
{
    __shared__ uint64_t sm[9];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) s += in[base + i];
    uint64_t total;
    block_exscan_256(s, &total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single CTA: in-place exclusive scan of the tile sums, total -> tile_sums[n_tiles]
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(uint64_t* tile_sums, size_t n_tiles)
{
    __shared__ uint64_t sm[9];
    uint64_t carry = 0;
    for (size_t base = 0; base < n_tiles; base += SCAN_THREADS) {
        size_t i = base + threadIdx.x;
        uint64_t v = i < n_tiles ? tile_sums[i] : 0;
        uint64_t total;
        uint64_t ex = block_exscan_256(v, &total, sm);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sums[n_tiles] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* __restrict__ in, size_t n,
                                                                   const uint64_t* __restrict__ tile_sums,
                                                                   uint64_t* __restrict__ out, size_t n_tiles)
{
    __shared__ uint64_t sm[9];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    uint64_t total;
    uint64_t ex = block_exscan_256(s, &total, sm) + tile_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = tile_sums[n_tiles];
}

int exclusive_scan_u32_u64(Engine* e, const uint32_t* d_in, uint64_t* d_out, size_t n)
{
    size_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (n_tiles == 0) n_tiles = 1;
    DBuf<uint64_t> sums;
    MXE_TRY(sums.alloc(n_tiles + 1, e->stream));
    MXE_LAUNCH(e, scan_reduce_kernel, (unsigned)n_tiles, SCAN_THREADS, 0, d_in, n, sums.p);
    MXE_LAUNCH(e, scan_tiles_kernel, 1, SCAN_THREADS, 0, sums.p, n_tiles);
    MXE_LAUNCH(e, scan_apply_kernel, (unsigned)n_tiles, SCAN_THREADS, 0, d_in, n, sums.p, d_out, n_tiles);
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

// ------------------------------------------------------------------------------------------
// bitmap rank directory + ordered extraction
// ------------------------------------------------------------------------------------------
// One warp per 8192 bits (8 rank blocks), one lane per 8 consecutive words (two 16-byte loads in flight per lane).
constexpr int BW = 8;                                    // words per lane
constexpr int BBLOCKS = BW * 32 / RANK_BLOCK_WORDS;      // rank blocks per warp

__device__ __forceinline__ void bm_load8(const uint32_t* __restrict__ bits, size_t w0, size_t n_words, uint32_t out[BW])
{
    if (w0 + BW <= n_words) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(bits + w0)), b = __ldg(reinterpret_cast<const uint4*>(bits + w0) + 1);
        out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < BW; j++) out[j] = w0 + j < n_words ? __ldg(bits + w0 + j) : 0u;
    }
}

__global__ void __launch_bounds__(256) bitmap_block_count_kernel(const uint32_t* __restrict__ bits, size_t n_words,
                                                                  uint32_t* __restrict__ counts, size_t n_blocks)
{
    const size_t sb = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const size_t blk0 = sb * BBLOCKS;
    if (blk0 >= n_blocks) return;
    uint32_t wv[BW];
    bm_load8(bits, sb * (32 * BW) + (size_t)lane * BW, n_words, wv);
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < BW; j++) c += __popc(wv[j]);
    c += __shfl_xor_sync(0xffffffffu, c, 1);             // four lanes = one 1024-bit rank block
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    const size_t blk = blk0 + (lane >> 2);
    if ((lane & 3) == 0 && blk < n_blocks) counts[blk] = c;
}

int bitmap_rank_build(Engine* e, const uint32_t* d_bits, size_t n_words, uint64_t* d_prefix)
{
    size_t n_blocks = (n_words + RANK_BLOCK_WORDS - 1) / RANK_BLOCK_WORDS;
    DBuf<uint32_t> counts;
    MXE_TRY(counts.alloc(n_blocks, e->stream));
    if (n_blocks) {
        unsigned grid = (unsigned)(((n_blocks + BBLOCKS - 1) / BBLOCKS * 32 + 255) / 256);
        MXE_LAUNCH(e, bitmap_block_count_kernel, grid, 256, 0, d_bits, n_words, counts.p, n_blocks);
    }
    MXE_TRY(exclusive_scan_u32_u64(e, counts.p, d_prefix, n_blocks));
    return MXE_OK;
}

__global__ void __launch_bounds__(256) bitmap_extract_kernel(const uint32_t* __restrict__ bits, size_t n_words,
                                                              const uint64_t* __restrict__ prefix, size_t n_blocks,
                                                              uint64_t* __restrict__ out, uint64_t cap)
{
    const size_t sb = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const size_t blk0 = sb * BBLOCKS;
    if (blk0 >= n_blocks) return;
    const size_t blk1 = blk0 + BBLOCKS < n_blocks ? blk0 + BBLOCKS : n_blocks;
    const uint64_t b0 = prefix[blk0];
    if (prefix[blk1] == b0) return;      // empty (warp-uniform)
    const size_t w0 = sb * (32 * BW) + (size_t)lane * BW;
    uint32_t wv[BW];
    bm_load8(bits, w0, n_words, wv);
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < BW; j++) c += __popc(wv[j]);
    uint32_t x = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    uint64_t o = b0 + (x - c);
#pragma unroll
    for (int j = 0; j < BW; j++) {
        uint32_t v = wv[j];
        const uint64_t base = (uint64_t)(w0 + j) << 5;
        while (v) {
            int b = __ffs(v) - 1;
            v &= v - 1;
            if (o < cap) out[o] = base + b;
            o++;
        }
    }
}

int bitmap_extract(Engine* e, const uint32_t* d_bits, size_t n_words, const uint64_t* d_prefix, uint64_t* d_out, uint64_t cap)
{
    size_t n_blocks = (n_words + RANK_BLOCK_WORDS - 1) / RANK_BLOCK_WORDS;
    if (!n_blocks) return MXE_OK;
    unsigned grid = (unsigned)(((n_blocks + BBLOCKS - 1) / BBLOCKS * 32 + 255) / 256);
    MXE_LAUNCH(e, bitmap_extract_kernel, grid, 256, 0, d_bits, n_words, d_prefix, n_blocks, d_out, cap);
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

// ------------------------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass, (u64 key, u32 value)
//   histogram: per-tile digit counts -> counts[digit * n_tiles + tile]
//   scan     : exclusive_scan_u32_u64 over that array (digit-major => stable global order)
//   scatter  : each warp owns a contiguous 32*ITEMS slice of the tile; ranks are computed
//              with __match_any_sync so equal digits keep their input order.
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 4096 keys per CTA

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t* __restrict__ keys, size_t n, int shift,
                                                              uint32_t* __restrict__ counts, size_t n_tiles)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int i = 0; i < RS_ITEMS; i++) {
        size_t idx = base + (size_t)i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                                 uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                 size_t n, int shift, const uint64_t* __restrict__ offsets, size_t n_tiles)
{
    __shared__ uint32_t whist[RS_WARPS][256];   // per-warp digit counts -> per-warp running offsets
    __shared__ uint64_t gbase[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&whist[0][0])[i] = 0;
    gbase[threadIdx.x] = offsets[(size_t)threadIdx.x * n_tiles + blockIdx.x];
    __syncthreads();

    size_t wbase = (size_t)blockIdx.x * RS_TILE + (size_t)warp * (32 * RS_ITEMS);
    uint64_t k[RS_ITEMS];
    uint32_t v[RS_ITEMS];
    uint32_t rank_in_warp[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        size_t idx = wbase + (size_t)i * 32 + lane;
        bool ok = idx < n;
        k[i] = ok ? keys[idx] : ~0ULL;
        v[i] = ok ? vals[idx] : 0;
        uint32_t d = ok ? (uint32_t)((k[i] >> shift) & 0xFF) : 256u;   // 256: invalid, never matches a real digit
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t before = __popc(peers & ((1u << lane) - 1));
        uint32_t prev = 0;
        if (ok) prev = whist[warp][d];           // all peers read the same value
        __syncwarp();
        if (ok && before == 0) whist[warp][d] = prev + __popc(peers);
        __syncwarp();
        rank_in_warp[i] = prev + before;
    }
    __syncthreads();
    // exclusive scan across warps for each digit (thread d handles digit d)
    {
        uint32_t run = 0;
#pragma unroll
        for (int wv = 0; wv < RS_WARPS; wv++) {
            uint32_t c = whist[wv][threadIdx.x];
            whist[wv][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        size_t idx = wbase + (size_t)i * 32 + lane;
        if (idx < n) {
            uint32_t d = (uint32_t)((k[i] >> shift) & 0xFF);
            uint64_t dst = gbase[d] + whist[warp][d] + rank_in_warp[i];
            keys_out[dst] = k[i];
            vals_out[dst] = v[i];
        }
    }
}

int radix_sort_pairs(Engine* e, uint64_t* d_keys, uint32_t* d_vals, uint64_t* d_keys_alt, uint32_t* d_vals_alt,
                     size_t n, int begin_bit, int end_bit)
{
    if (n == 0) return MXE_OK;
    size_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
    DBuf<uint32_t> counts;
    DBuf<uint64_t> offsets;
    MXE_TRY(counts.alloc(256 * n_tiles, e->stream));
    MXE_TRY(offsets.alloc(256 * n_tiles + 1, e->stream));
    uint64_t* kin = d_keys; uint32_t* vin = d_vals;
    uint64_t* kout = d_keys_alt; uint32_t* vout = d_vals_alt;
    int passes = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        MXE_LAUNCH(e, rs_hist_kernel, (unsigned)n_tiles, RS_THREADS, 0, kin, n, shift, counts.p, n_tiles);
        MXE_TRY(exclusive_scan_u32_u64(e, counts.p, offsets.p, 256 * n_tiles));
        MXE_LAUNCH(e, rs_scatter_kernel, (unsigned)n_tiles, RS_THREADS, 0, kin, vin, kout, vout, n, shift, offsets.p, n_tiles);
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
        passes++;
    }
    if (passes & 1) {
        MXE_CUDA(cudaMemcpyAsync(d_keys, kin, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, e->stream));
        MXE_CUDA(cudaMemcpyAsync(d_vals, vin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, e->stream));
    }
    MXE_CUDA(cudaGetLastError());
    return MXE_OK;
}

}  // namespace mxe
