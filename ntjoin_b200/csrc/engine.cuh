// engine.cuh -- object definitions behind the opaque C handles.
#pragma once
#include "common.cuh"

constexpr int MXE_N_CHUNK_EVENTS = 64;

struct mxe_engine : public mxe::Engine {
    // host-input path: two copy streams (two copy engines) + events
    cudaStream_t copy_stream[2] = {nullptr, nullptr};
    cudaEvent_t ev_ready = nullptr;
    cudaEvent_t ev_chunk[MXE_N_CHUNK_EVENTS] = {nullptr};
    int h2d_chunk_mb = 256;
    // pinned host block pool (grow-only, reused across steps)
    struct Pinned { void* p; size_t bytes; };
    std::vector<Pinned> pinned_free;
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
    // sequence text for --seq output: either an owned copy of the whole input or nothing
    std::vector<char> seq_text;       // upper-cased concatenated sequence (mxe_sketch_file)
    const uint8_t* seq_borrowed = nullptr;   // caller buffer (mxe_sketch_buffers), valid while caller keeps it
};

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
    // lazy host copy (one pinned block)
    void* h_block = nullptr; size_t h_bytes = 0;
    uint8_t* h_uniq = nullptr; uint8_t* h_keep = nullptr; uint64_t* h_vertices = nullptr;
    uint64_t* h_eu = nullptr; uint64_t* h_ev = nullptr; uint32_t* h_emask = nullptr; double* h_ew = nullptr;
};

namespace mxe {
int sketch_device_impl(mxe_engine* e, const uint8_t* d_seq, uint64_t n, const uint64_t* offsets, uint32_t n_contigs,
                       int k, int w, int flags, mxe_sketch* out, const uint8_t* h_seq = nullptr);
int filter_and_edges_impl(mxe_engine* e, const uint64_t* const* d_hash, const uint32_t* const* d_contig,
                          const uint64_t* n, int n_asm, const double* weights, mxe_result* out);
}
