// hostio.cu -- host side of the file seam (S1/S2): FASTA/FASTQ ingest and `indexlr` TSV text, multi-threaded.
//
// Around ~60 ms of GPU work per 3 Gbp assembly the file seam used to spend seconds in a single-threaded parser (two
// zero-filled 3 GB buffers, a copy per line, an upper-casing pass) and a single-threaded text writer.  Here the file is
// memory-mapped and cut into chunks at line starts; one parallel pass counts sequence bytes and finds the headers, a
// prefix sum places every chunk, a second parallel pass copies and upper-cases straight into the (uninitialised) text
// buffer.  The TSV is formatted by several threads over contiguous record ranges and written in order.
// Reference behaviour kept: btllib SeqReader / indexlr as ntJoin uses them: record id = header up
// to the first blank, sequence lines joined, one trailing CR per line dropped, lines before the first header ignored,
// FASTQ by 4-line records, sequence upper-cased; TSV = id \t hash[:pos][:strand][:seq] ... \n.
#include "engine.cuh"

#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <thread>

namespace mxe {

int host_threads()
{
    if (const char* s = getenv("MXE_HOST_THREADS")) { int t = atoi(s); if (t >= 1) return std::min(t, 256); }
    unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(hc ? hc : 1u, 16u));
}

template <typename F>
static void parallel_for(size_t n_tasks, int threads, F f)
{
    if (threads <= 1 || n_tasks <= 1) { for (size_t i = 0; i < n_tasks; i++) f(i); return; }
    std::atomic<size_t> next(0);
    std::vector<std::thread> pool;
    const int nt = (int)std::min<size_t>((size_t)threads, n_tasks);
    for (int t = 0; t < nt; t++)
        pool.emplace_back([&]() { for (size_t i = next.fetch_add(1); i < n_tasks; i = next.fetch_add(1)) f(i); });
    for (auto& th : pool) th.join();
}

static inline void copy_upper(char* dst, const char* src, size_t n)
{
    for (size_t i = 0; i < n; i++) { char c = src[i]; dst[i] = (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }
}

static inline std::string header_id(const char* s, const char* end)
{
    const char* q = s;
    while (q < end && *q != ' ' && *q != '\t' && *q != '\r' && *q != '\n') q++;
    return std::string(s, q - s);
}

struct Chunk {
    size_t b, e;                                        // byte range, starts at a line start
    uint64_t nseq = 0;                                  // sequence bytes in the chunk
    std::vector<std::pair<size_t, uint64_t>> hdr;       // header line offset, sequence bytes of the chunk before it
    uint64_t out = 0;                                   // where the chunk's sequence goes
};

static int parse_fastq(const char* base, size_t size, HostText& seq, std::vector<uint64_t>& offsets, std::vector<std::string>& names)
{
    const char* p = base;
    const char* end = base + size;
    seq.resize(size + 64);
    size_t at = 0;
    auto eol = [&](const char* q) { const char* nl = q < end ? (const char*)memchr(q, '\n', end - q) : nullptr; return nl ? nl : end; };
    while (p < end) {
        const char* nl = eol(p);
        if (nl == p || (nl == p + 1 && *p == '\r')) { p = nl < end ? nl + 1 : end; continue; }     // blank line (e.g. at the end of the file): no record
        names.push_back(header_id(p + 1, nl));
        offsets.push_back(at);
        p = nl < end ? nl + 1 : end;
        nl = eol(p);
        size_t len = nl - p;
        if (len && p[len - 1] == '\r') len--;
        copy_upper(&seq[at], p, len);
        at += len;
        p = nl < end ? nl + 1 : end;
        for (int i = 0; i < 2 && p < end; i++) { nl = eol(p); p = nl < end ? nl + 1 : end; }
    }
    offsets.push_back(at);
    memset(&seq[at], 0, 64);
    seq.resize(at + 64);
    return MXE_OK;
}

static int parse_text(const char* path, const char* base, size_t size, HostText& seq, std::vector<uint64_t>& offsets, std::vector<std::string>& names);

// btllib reads gzip / bzip2 / xz / zstd input through external decompressors (a pipe from `gzip -dc` and friends); so does
// this reader: the decompressed text is collected in memory and parsed like a mapped file.
static int read_compressed(const char* path, const char* tool, HostText& seq, std::vector<uint64_t>& offsets, std::vector<std::string>& names)
{
    std::string cmd = std::string(tool) + " -dc -- '";
    for (const char* q = path; *q; q++) { if (*q == '\'') cmd += "'\\''"; else cmd += *q; }      // the path inside single quotes
    cmd += "' 2>/dev/null";
    FILE* pipe = popen(cmd.c_str(), "r");
    if (!pipe) { set_error("cannot start `%s` for %s: %s", tool, path, strerror(errno)); return MXE_ERR_IO; }
    HostText text;
    size_t n = 0;
    for (;;) {
        if (text.size() < n + (1 << 20)) text.resize(std::max<size_t>(text.size() * 2, (size_t)64 << 20));
        const size_t r = fread(text.data() + n, 1, text.size() - n, pipe);
        if (r == 0) break;
        n += r;
    }
    const int status = pclose(pipe);
    if (status != 0) {
        set_error("`%s -dc %s` failed (exit status %d): the file is damaged or the tool is not installed", tool, path, status);
        return MXE_ERR_IO;
    }
    return parse_text(path, text.data(), n, seq, offsets, names);
}

int read_fasta(const char* path, HostText& seq, std::vector<uint64_t>& offsets, std::vector<std::string>& names)
{
    offsets.clear(); names.clear();
    int fd = open(path, O_RDONLY);
    if (fd < 0) { set_error("cannot open %s: %s", path, strerror(errno)); return MXE_ERR_IO; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) {
        // not a regular file (pipe, process substitution): read it all, then parse the copy sequentially
        std::vector<char> tmp;
        char blk[1 << 16];
        ssize_t r;
        while ((r = read(fd, blk, sizeof blk)) > 0) tmp.insert(tmp.end(), blk, blk + r);
        close(fd);
        if (r < 0) { set_error("read error on %s: %s", path, strerror(errno)); return MXE_ERR_IO; }
        const char* base = tmp.data();
        const size_t size = tmp.size();
        if (size && base[0] == '@') return parse_fastq(base, size, seq, offsets, names);
        bool have = false;
        seq.resize(size + 64);
        size_t at = 0;
        const char* p = base;
        const char* end = base + size;
        while (p < end) {
            const char* nl = (const char*)memchr(p, '\n', end - p);
            if (!nl) nl = end;
            if (*p == '>') { names.push_back(header_id(p + 1, nl)); offsets.push_back(at); have = true; }
            else if (have) { size_t len = nl - p; if (len && p[len - 1] == '\r') len--; copy_upper(&seq[at], p, len); at += len; }
            p = nl < end ? nl + 1 : end;
        }
        offsets.push_back(at);
        memset(&seq[at], 0, 64);
        seq.resize(at + 64);
        return MXE_OK;
    }
    const size_t size = (size_t)sb.st_size;
    if (size == 0) { close(fd); offsets.push_back(0); seq.resize(64); memset(seq.data(), 0, 64); return MXE_OK; }
    const char* base = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (base == (const char*)MAP_FAILED) { set_error("cannot map %s: %s", path, strerror(errno)); return MXE_ERR_IO; }
    madvise((void*)base, size, MADV_SEQUENTIAL);
    struct Unmap { const char* p; size_t n; ~Unmap() { munmap((void*)p, n); } } unmap{base, size};
    if (size >= 4) {
        const unsigned char* m = (const unsigned char*)base;
        const char* tool = (m[0] == 0x1f && m[1] == 0x8b) ? "gzip"
                         : (m[0] == 'B' && m[1] == 'Z' && m[2] == 'h') ? "bzip2"
                         : (m[0] == 0xfd && m[1] == '7' && m[2] == 'z' && m[3] == 'X') ? "xz"
                         : (m[0] == 0x28 && m[1] == 0xb5 && m[2] == 0x2f && m[3] == 0xfd) ? "zstd" : nullptr;
        if (tool) return read_compressed(path, tool, seq, offsets, names);
    }
    return parse_text(path, base, size, seq, offsets, names);
}

// plain FASTA / FASTQ text in memory (a mapped file or the output of a decompressor)
static int parse_text(const char* path, const char* base, size_t size, HostText& seq, std::vector<uint64_t>& offsets, std::vector<std::string>& names)
{
    if (size == 0) { offsets.push_back(0); seq.resize(64); memset(seq.data(), 0, 64); return MXE_OK; }
    if (base[0] != '>' && base[0] != '@' && base[0] != '\n' && base[0] != '\r' && (base[0] < 0x20 || (unsigned char)base[0] > 0x7e)) {
        set_error("%s does not look like FASTA / FASTQ text (first byte 0x%02x)", path, (unsigned char)base[0]);
        return MXE_ERR_IO;
    }
    if (base[0] == '@') return parse_fastq(base, size, seq, offsets, names);

    const int threads = host_threads();
    const char* fend = base + size;
    // chunks of ~8 MB that start at line starts
    const size_t CH = (size_t)8 << 20;
    std::vector<size_t> bounds{0};
    for (size_t q = CH; q < size; q += CH) {
        const char* nl = (const char*)memchr(base + q - 1, '\n', size - (q - 1));
        const size_t b = nl ? (size_t)(nl + 1 - base) : size;
        if (b > bounds.back() && b < size) bounds.push_back(b);
    }
    bounds.push_back(size);
    std::vector<Chunk> chunks(bounds.size() - 1);
    for (size_t i = 0; i + 1 < bounds.size(); i++) { chunks[i].b = bounds[i]; chunks[i].e = bounds[i + 1]; }

    // pass A: headers and sequence byte counts per chunk
    parallel_for(chunks.size(), threads, [&](size_t ci) {
        Chunk& c = chunks[ci];
        const char* p = base + c.b;
        const char* end = base + c.e;
        while (p < end) {
            const char* nl = (const char*)memchr(p, '\n', fend - p);
            if (!nl) nl = fend;
            if (*p == '>') c.hdr.emplace_back((size_t)(p - base), c.nseq);
            else { size_t len = nl - p; if (len && p[len - 1] == '\r') len--; c.nseq += len; }
            p = nl + 1;
        }
    });
    // lines before the first header of the file are ignored
    size_t first = chunks.size();
    for (size_t ci = 0; ci < chunks.size(); ci++) if (!chunks[ci].hdr.empty()) { first = ci; break; }
    uint64_t total = 0, dropped = 0;
    for (size_t ci = 0; ci < chunks.size(); ci++) {
        Chunk& c = chunks[ci];
        c.out = total;
        if (ci < first) continue;
        if (ci == first) { dropped = c.hdr[0].second; total += c.nseq - dropped; }
        else total += c.nseq;
    }
    size_t n_hdr = 0;
    for (auto& c : chunks) n_hdr += c.hdr.size();
    names.reserve(n_hdr); offsets.reserve(n_hdr + 1);
    for (size_t ci = first; ci < chunks.size(); ci++)
        for (auto& h : chunks[ci].hdr) {
            const char* s = base + h.first + 1;
            const char* nl = (const char*)memchr(s, '\n', fend - s);
            names.push_back(header_id(s, nl ? nl : fend));
            offsets.push_back(chunks[ci].out + h.second - (ci == first ? dropped : 0));
        }
    offsets.push_back(total);
    seq.resize(total + 64);                              // uninitialised: every byte below `total` is written in pass B
    char* out_base = seq.data();

    // pass B: copy + upper-case
    parallel_for(chunks.size(), threads, [&](size_t ci) {
        if (ci < first) return;
        const Chunk& c = chunks[ci];
        const char* p = base + c.b;
        const char* end = base + c.e;
        char* dst = out_base + c.out;
        bool active = ci > first;
        while (p < end) {
            const char* nl = (const char*)memchr(p, '\n', fend - p);
            if (!nl) nl = fend;
            if (*p == '>') active = true;
            else if (active) { size_t len = nl - p; if (len && p[len - 1] == '\r') len--; copy_upper(dst, p, len); dst += len; }
            p = nl + 1;
        }
    });
    memset(out_base + total, 0, 64);
    return MXE_OK;
}

// ---------------------------------------------------------------------------------------------------- TSV text
struct TextBuf {
    std::vector<char, DefaultInit<char>> v;
    size_t n = 0;
    char* room(size_t need) { if (v.size() < n + need) v.resize(std::max(v.size() * 2, n + need + (1 << 16))); return v.data() + n; }
};

static void format_records(const mxe_sketch* S, const char* text, uint32_t c0, uint32_t c1, int with_pos, int with_strand,
                           int with_seq, TextBuf& out)
{
    size_t i = std::lower_bound(S->h_contig, S->h_contig + S->n, c0) - S->h_contig;
    const size_t per = 24 + (with_pos ? 12 : 0) + (with_strand ? 2 : 0) + (with_seq ? (size_t)S->k + 1 : 0);
    for (uint32_t c = c0; c < c1; c++) {
        const std::string& nm = S->names[c];
        char* p = out.room(nm.size() + 2);
        memcpy(p, nm.data(), nm.size()); p += nm.size();
        *p++ = '\t';
        out.n = p - out.v.data();
        bool first = true;
        while (i < S->n && S->h_contig[i] == c) {
            p = out.room(per + 2);
            if (!first) *p++ = ' ';
            first = false;
            p = put_u64(p, S->h_out_hash[i]);
            if (with_pos) { *p++ = ':'; p = put_u64(p, S->h_pos[i]); }
            if (with_strand) { *p++ = ':'; *p++ = S->h_forward[i] ? '+' : '-'; }
            if (with_seq) {
                *p++ = ':';
                copy_upper(p, text + S->offsets[c] + S->h_pos[i], (size_t)S->k);
                p += S->k;
            }
            out.n = p - out.v.data();
            i++;
        }
        p = out.room(1);
        *p++ = '\n';
        out.n = p - out.v.data();
    }
}

int write_tsv_text(const mxe_sketch* S, const char* text, FILE* f, int with_pos, int with_strand, int with_seq)
{
    const uint32_t nc = S->n_contigs;
    if (nc == 0) return MXE_OK;
    int threads = host_threads();
    if (S->n < (1u << 16)) threads = 1;
    // contiguous record ranges with about the same number of minimizers
    std::vector<uint32_t> cut{0};
    for (int t = 1; t < threads; t++) {
        const uint64_t it = S->n * (uint64_t)t / threads;
        const uint32_t c = it < S->n ? S->h_contig[it] : nc;
        if (c > cut.back() && c < nc) cut.push_back(c);
    }
    cut.push_back(nc);
    const size_t n_ranges = cut.size() - 1;
    std::vector<TextBuf> bufs(n_ranges);
    parallel_for(n_ranges, threads, [&](size_t r) { format_records(S, text, cut[r], cut[r + 1], with_pos, with_strand, with_seq, bufs[r]); });
    bool ok = true;
    for (auto& b : bufs) ok = ok && fwrite(b.v.data(), 1, b.n, f) == b.n;
    return ok ? MXE_OK : MXE_ERR_IO;
}

}  // namespace mxe
