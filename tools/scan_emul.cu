// scan_emul.cu -- CPU emulation of scan_kernels.cuh (runs without a GPU): the per-thread core of scan_bs2_kernel
// (bs2_tile: transposition, phase bookkeeping, ring, threshold test) and the byte-validity screen of pack2_kernel are
// __host__ __device__, so they are checked here against plain integer arithmetic before any GPU time is spent.
//   nvcc -std=c++17 -O1 -I ntjoin_b200/csrc -o /tmp/scan_emul tools/scan_emul.cu && /tmp/scan_emul
#include "scan_kernels.cuh"
#include <cstdlib>
#include <cstring>
#include <set>
#include <vector>
namespace mxe { void set_error(const char*, ...) {} Arena* current_arena() { return nullptr; } }
using namespace mxe;

struct HostRing {
    uint32_t lo[64], hi[64];
    void put(int s, uint32_t a, uint32_t b) { lo[s] = a; hi[s] = b; }
    void get(int s, uint32_t& a, uint32_t& b) const { a = lo[s]; b = hi[s]; }
};
struct HostQueue {
    uint32_t w[16];
    void put(int ph, uint32_t v) { w[ph] = v; }
    uint32_t get(int ph) const { return w[ph]; }
};
struct HostEmit {     // as Bs2EmitGlobal; also checks that no position is emitted twice
    std::set<uint64_t>* out; uint64_t base; uint32_t L;
    void operator()(int j, int s) const
    {
        if (!out->insert(base + (uint64_t)j * L + (uint64_t)s).second) printf("position emitted twice\n");
    }
};

template <int KMOD, int H, int HS, int RING>
static int run(int k, uint32_t Lw, double frac)
{
    const uint32_t L = 16 * Lw, R = (L + k - 1 + 15) / 16;
    const uint64_t tile = 32ull * L, n_tiles = 2, n = n_tiles * tile + 1000;
    std::vector<uint8_t> code(n + 4096, 0);
    for (uint64_t i = 0; i < n; i++) code[i] = rand() & 3;
    // position-ordered pk (byte-interleaved bit layout), zero past n
    std::vector<uint32_t> pk((n + 4096) / 16 + 8, 0);
    for (uint64_t p = 0; p < n; p++) { const uint32_t i = p & 15; pk[p >> 4] |= (uint32_t)code[p] << (((i & 3) << 3) | ((i >> 2) << 1)); }
    // bit planes as pack2_kernel writes them
    std::vector<uint32_t> pkT(n_tiles * R * 32);
    for (uint64_t T = 0; T < n_tiles; T++)
        for (uint32_t g = 0; g < R; g++)
        {
            uint32_t x[32];
            for (uint32_t j = 0; j < 32; j++) x[j] = pk[T * 2 * L + j * Lw + g];
            transpose32(x);
            for (int u = 0; u < 16; u++)
                for (int pl = 0; pl < 2; pl++) pkT[(T * R + g) * 32 + 2 * u + pl] = x[8 * (u & 3) + 2 * (u >> 2) + pl];
        }
    // threshold
    const uint32_t Tthr = (uint32_t)(frac * 2147483648.0);
    const uint32_t TH = Tthr >> (31 - H);
    uint32_t Q = (TH + 1) >> (H - HS);
    if (Q > (1u << HS) - 1) Q = (1u << HS) - 1;
    const uint32_t K = ((1u << HS) - 1) - Q;
    Bs2Params P;
    for (int i = 0; i < HS; i++) P.kmask[i] = ((K >> i) & 1u) ? 0xFFFFFFFFu : 0u;
    uint32_t shi[4];
    for (int c = 0; c < 4; c++) shi[c] = bs_shi(c);
    P.f0 = P.r0 = 0;
    for (int i = 0; i < k; i++) { P.f0 = rol31(P.f0) ^ shi[0]; P.r0 = rol31(P.r0) ^ shi[2]; }
    std::set<uint64_t> got, want, exact;
    for (uint64_t T = 0; T < n_tiles; T++) {
        HostRing ring;
        HostQueue q;
        HostEmit emit{&got, T * tile, L};
        bs2_tile<KMOD, H, HS, RING>(pkT.data() + T * R * 32, (int)R, (int)L, k, P, ring, q, emit);
    }
    for (uint64_t p = 0; p < n_tiles * tile; p++) {
        uint32_t f = 0, r = 0;
        for (int i = 0; i < k; i++) { f = rol31(f) ^ shi[code[p + i]]; r = rol31(r) ^ shi[code[p + k - 1 - i] ^ 2]; }
        const uint32_t S = ((f >> (31 - H)) + (r >> (31 - H)) + 1u) & ((1u << H) - 1u);
        if ((S >> (H - HS)) <= Q) want.insert(p);
        if (((f + r) & 0x7FFFFFFFu) <= Tthr || ((f + r + 1) & 0x7FFFFFFFu) <= Tthr) exact.insert(p);   // t = (f31 + r31 + carry) mod 2^31
    }
    int bad = 0;
    if (got != want) { bad = 1; printf("MISMATCH k=%d Lw=%u H=%d HS=%d: got %zu want %zu\n", k, Lw, H, HS, got.size(), want.size()); }
    for (uint64_t p : exact) if (!want.count(p)) { bad = 1; printf("NOT A SUPERSET at %llu\n", (unsigned long long)p); break; }
    printf("k=%d Lw=%u H=%d HS=%d frac=%.4f: %zu candidates (%.3f %%), exact set %zu -> %s\n", k, Lw, H, HS, frac, got.size(),
           100.0 * got.size() / (double)(n_tiles * tile), exact.size(), bad ? "FAIL" : "ok");
    return bad;
}

static int check_bytes()
{
    int bad = 0;
    for (int v = 0; v < 256; v++) {
        for (int slot = 0; slot < 32; slot++) {
            uint8_t bytes[32];
            memset(bytes, "ACGTacgt"[(v + slot) & 7], 32);
            bytes[slot] = (uint8_t)v;
            uint32_t wv[8];
            memcpy(wv, bytes, 32);
            const bool valid = strchr("ACGTacgt", v) != nullptr && v != 0;
            const bool flagged = bad_screen(wv) != 0;
            if (flagged == valid) { printf("bad_screen wrong for byte 0x%02x at %d\n", v, slot); bad = 1; }
        }
    }
    // pk_from_ascii against the definition
    for (int it = 0; it < 1000; it++) {
        uint8_t bytes[16];
        for (int i = 0; i < 16; i++) bytes[i] = "ACGTacgt"[rand() & 7];
        uint32_t x[4];
        memcpy(x, bytes, 16);
        uint32_t want = 0;
        for (int i = 0; i < 16; i++) {
            const uint32_t c = ((bytes[i] & 0xDF) ^ 0x41) >> 1 & 3;
            want |= c << (((i & 3) << 3) | ((i >> 2) << 1));
        }
        if (pk_from_ascii(x[0], x[1], x[2], x[3]) != want) { printf("pk_from_ascii mismatch\n"); bad = 1; break; }
    }
    // transpose32
    uint32_t a[32], b[32];
    for (int i = 0; i < 32; i++) a[i] = b[i] = (uint32_t)rand() * 2654435761u + rand();
    transpose32(b);
    for (int r = 0; r < 32; r++) for (int j = 0; j < 32; j++) if (((b[r] >> j) & 1u) != ((a[j] >> r) & 1u)) { printf("transpose wrong\n"); return 1; }
    printf("byte screen, pk_from_ascii, transpose32: %s\n", bad ? "FAIL" : "ok");
    return bad;
}

int main()
{
    srand(12345);
    int bad = check_bytes();
    bad |= run<1, 16, 12, 48>(32, 9, 0.009);
    bad |= run<1, 12, 12, 48>(32, 11, 0.009);
    bad |= run<1, 12, 12, 48>(32, 13, 0.036);
    bad |= run<1, 12, 12, 64>(32, 9, 0.0018);
    bad |= run<9, 12, 12, 64>(40, 9, 0.009);
    bad |= run<24, 12, 12, 48>(24, 9, 0.018);
    bad |= run<1, 11, 11, 48>(32, 15, 0.009);
    bad |= run<1, 12, 12, 48>(32, 63, 0.009);
    return bad;
}
