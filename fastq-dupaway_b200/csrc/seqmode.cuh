// seqmode.cuh - whole-input modes on the device:
//   * sequence-based deduplication: SeqDupRemover<T>::filterSE/PE = ExternalSorter / PairedExternalSorter +
//     the comparator scan (src/seq_dup_remover.hpp:40-218, src/external_sort.hpp, src/paired_external_sort.hpp,
//     src/comparator.cpp:45-91)
//   * --fast --unordered: two ID-tag sorts + merge-join + pair set (src/hash_dup_remover.hpp:150-192,257-347)
// The raw input stays in HBM (segments of at most max_chunk_bytes), K1 (parse_pack.cuh) splits and packs it,
// sortlib.cuh sorts record indices by packed key rows (stable on the input index), and the scans below decide
// which records are written.  Nothing spills to disk; nothing is computed on the host.
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/fqd.h"
#include "common.cuh"
#include "hashset.cuh"
#include "parse_pack.cuh"
#include "sortlib.cuh"

namespace fqd {

#define SEQ_TRY(call)                                                                            \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            char b_[512];                                                                        \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            *err = b_;                                                                           \
            return FQD_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// device helpers on packed rows (3-bit codes, 20 per word, first base in the top bits; see common.cuh)
__device__ __forceinline__ bool rows_equal_range(const u64* a, const u64* b, u32 w0, u32 w1) {
    u64 diff = 0;
    for (u32 w = w0; w < w1; ++w) diff |= a[w] ^ b[w];
    return diff == 0;
}
// first nb bases equal?
__device__ __forceinline__ bool prefix_equal(const u64* a, const u64* b, u32 nb) {
    const u32 full = nb / BASES_PER_WORD, rem = nb % BASES_PER_WORD;
    u64 diff = 0;
    for (u32 w = 0; w < full; ++w) diff |= a[w] ^ b[w];
    if (rem) diff |= (a[full] ^ b[full]) & (~0ull << (3u * (BASES_PER_WORD - rem)));
    return diff == 0;
}
// number of differing bases (SeqUtils::hammingDistance, src/seq_utils.cpp:65-72, on equal-length sequences)
__device__ __forceinline__ u32 hamming_words(const u64* a, const u64* b, u32 W) {
    u32 d = 0;
    for (u32 w = 0; w < W; ++w) {
        u64 x = a[w] ^ b[w];
        x = (x | (x >> 1) | (x >> 2)) & 0x0249249249249249ull;
        d += (u32)__popcll(x);
    }
    return d;
}

struct ScanParams {
    const u64* rows; u32 stride; u32 W; u32 mates;
    const u32* len0; const u32* len1;     // sequence lengths (bases) per record and mate
    const u32* perm; u64 n; u32 dist;
    u32* keep;                            // [n] in sorted order: 1 = written
    u32* brk;                             // hamming: definite cluster breaks
};

// TightComparator (src/comparator.cpp:45-58) against the previous record of the sorted stream
__global__ void k_scan_tight(const ScanParams p) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        u32 k = 1;
        if (i > 0) {
            const u64* a = p.rows + (u64)p.perm[i] * p.stride;
            const u64* b = p.rows + (u64)p.perm[i - 1] * p.stride;
            k = rows_equal_range(a, b, 0, p.W * p.mates) ? 0u : 1u;
        }
        p.keep[i] = k;
    }
}
// LooseComparator (src/comparator.cpp:60-74) + the "keep the longest as reference" rule
// (src/seq_dup_remover.hpp:93-98,194-202).  On the sorted stream the reference head is always the previous
// record, so the test is: the previous record's mates are prefixes of mine and, paired, the
// overlap is same-sided.
__global__ void k_scan_loose(const ScanParams p) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        u32 k = 1;
        if (i > 0) {
            const u32 c = p.perm[i], h = p.perm[i - 1];
            const u64* a = p.rows + (u64)c * p.stride;
            const u64* b = p.rows + (u64)h * p.stride;
            const u32 lc1 = p.len0[c], lh1 = p.len0[h];
            bool dup = prefix_equal(a, b, min(lc1, lh1));
            if (dup && p.mates == 2) {
                const u32 lc2 = p.len1[c], lh2 = p.len1[h];
                dup = prefix_equal(a + p.W, b + p.W, min(lc2, lh2));
                if (dup) dup = ((lh1 <= lc1) && (lh2 <= lc2)) || ((lh1 > lc1) && (lh2 > lc2));
            }
            k = dup ? 0u : 1u;
        }
        p.keep[i] = k;
    }
}
// HammingComparator (src/comparator.cpp:76-91) compares against the cluster HEAD, a sequential greedy scan.
// Position j is a definite cluster break when its length differs from j-1 or hamming(j-1, j) > 2d on a mate
// (triangle inequality); the literal scan then runs independently inside each segment.
__global__ void k_ham_breaks(const ScanParams p) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        u32 b = 1;
        if (i > 0) {
            const u32 c = p.perm[i], h = p.perm[i - 1];
            const u64* a = p.rows + (u64)c * p.stride;
            const u64* q = p.rows + (u64)h * p.stride;
            bool same = p.len0[c] == p.len0[h] && hamming_words(a, q, p.W) <= 2u * p.dist;
            if (same && p.mates == 2) same = p.len1[c] == p.len1[h] && hamming_words(a + p.W, q + p.W, p.W) <= 2u * p.dist;
            b = same ? 0u : 1u;
        }
        p.brk[i] = b;
    }
}
__global__ void k_ham_segments(const ScanParams p) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        if (!p.brk[i]) continue;
        p.keep[i] = 1;
        u32 head = p.perm[i];
        for (u64 j = i + 1; j < p.n && !p.brk[j]; ++j) {
            const u32 c = p.perm[j];
            const u64* a = p.rows + (u64)c * p.stride;
            const u64* q = p.rows + (u64)head * p.stride;
            // lengths are equal inside a segment (a length change is a break)
            bool dup = hamming_words(a, q, p.W) <= p.dist;
            if (dup && p.mates == 2) dup = hamming_words(a + p.W, q + p.W, p.W) <= p.dist;
            p.keep[j] = dup ? 0u : 1u;
            if (!dup) head = c;
        }
    }
}

__global__ void k_finish_segment(const ChunkCtl* ctl, const u32* rec_start, u64 logical_base, u64* rec_off, u32* rec_len,
                                 RunState* run, ChunkCtl* ctl_out) {
    const u32 n = ctl->n_records;
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        rec_off[i] = logical_base + rec_start[i];
        rec_len[i] = rec_start[i + 1] - rec_start[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctl_out->consumed = rec_start[n];      // rec_start[0] is written by tile 0 even when n == 0
        run->chunk_pairs = n;
    }
}
__global__ void k_advance_run(RunState* run) { run->n_records += run->chunk_pairs; }

__global__ void k_emit(const u32* keep, const u32* excl, const u32* perm, u64 n, const u64* off0, const u32* len0,
                       const u64* off1, const u32* len1, u64* o_off0, u32* o_len0, u64* o_off1, u32* o_len1, u32* o_idx) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        if (!keep[i]) continue;
        const u32 o = excl[i], g = perm[i];
        o_idx[o] = g;
        o_off0[o] = off0[g]; o_len0[o] = len0[g];
        if (off1) { o_off1[o] = off1[g]; o_len1[o] = len1[g]; }
    }
}

// ---------------------------------------------------------------------------------------------------------
struct SeqSegment { u8* d = nullptr; size_t cap = 0, fill = 0; u64 logical_base = 0; };

struct SeqMate {
    std::vector<SeqSegment> segs;
    RunState* d_run = nullptr;
    u64 n_records = 0;
    u64* d_rec_off = nullptr;
    u32* d_rec_len = nullptr;
    u32* d_seq_len = nullptr;
    bool finished = false;
};

struct SeqState {
    fqd_config cfg;
    cudaStream_t stream = nullptr;
    int sm = 148;
    u32 W = 0, mates = 1, row_words = 0;
    u64 capacity = 0;
    size_t seg_bytes = 0;
    u32 chunk_cap = 0;
    SeqMate mate[2];
    u64* d_keys = nullptr;
    u64* d_tile_state = nullptr; u32 tiles_cap = 0;
    ChunkCtl* d_ctl = nullptr; ChunkCtl* h_ctl = nullptr;
    u32* d_rec_start = nullptr;
    u64* d_hash = nullptr;
    // results
    u64 n = 0, n_out = 0;
    fqd_stats_t stats;
    bool finished = false;
    std::vector<u64> h_off[2];
    std::vector<u32> h_len[2];
    std::vector<void*> scratch;      // freed at destroy / reset
    u64 launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double ms = 0.0;
};

static int seq_alloc_tables(SeqState* s, std::string* err) {
    SEQ_TRY(cudaMalloc(&s->d_keys, s->capacity * s->row_words * sizeof(u64)));
    for (u32 m = 0; m < s->mates; ++m) {
        SeqMate& mt = s->mate[m];
        SEQ_TRY(cudaMalloc(&mt.d_run, sizeof(RunState)));
        SEQ_TRY(cudaMemsetAsync(mt.d_run, 0, sizeof(RunState), s->stream));
        SEQ_TRY(cudaMalloc(&mt.d_rec_off, s->capacity * sizeof(u64)));
        SEQ_TRY(cudaMalloc(&mt.d_rec_len, s->capacity * sizeof(u32)));
        SEQ_TRY(cudaMalloc(&mt.d_seq_len, s->capacity * sizeof(u32)));
    }
    s->tiles_cap = (u32)((s->seg_bytes + PP_TILE - 1) / PP_TILE) + 1;
    SEQ_TRY(cudaMalloc(&s->d_tile_state, (size_t)s->tiles_cap * sizeof(u64)));
    SEQ_TRY(cudaMalloc(&s->d_ctl, sizeof(ChunkCtl)));
    SEQ_TRY(cudaHostAlloc(&s->h_ctl, sizeof(ChunkCtl), cudaHostAllocDefault));
    SEQ_TRY(cudaMalloc(&s->d_rec_start, ((size_t)s->chunk_cap + 1) * sizeof(u32)));
    SEQ_TRY(cudaMalloc(&s->d_hash, (size_t)s->chunk_cap * sizeof(u64)));
    return FQD_OK;
}

static int seq_create(SeqState** out, const fqd_config* cfg, cudaStream_t stream, int sm, u32 W, std::string* err) {
    SeqState* s = new SeqState();
    s->cfg = *cfg; s->stream = stream; s->sm = sm; s->W = W;
    s->mates = cfg->paired ? 2 : 1;
    s->row_words = W * s->mates;
    memset(&s->stats, 0, sizeof s->stats);
    if (cfg->max_records == 0) { *err = "max_records must be > 0"; delete s; return FQD_ERR_INVALID; }
    s->capacity = cfg->max_records;
    s->seg_bytes = (size_t)cfg->max_chunk_bytes;
    s->chunk_cap = (u32)std::min<u64>(cfg->max_chunk_records ? cfg->max_chunk_records : std::max<u64>(s->seg_bytes / 16, 1024), s->capacity);
    cudaEventCreate(&s->ev0); cudaEventCreate(&s->ev1);
    int rc = seq_alloc_tables(s, err);
    if (rc) { *out = s; return rc; }
    *out = s;
    return FQD_OK;
}

static void seq_free_results(SeqState* s) {
    for (void* p : s->scratch) cudaFree(p);
    s->scratch.clear();
    for (int m = 0; m < 2; ++m) { s->h_off[m].clear(); s->h_len[m].clear(); }
}

static void seq_destroy(SeqState* s) {
    if (!s) return;
    seq_free_results(s);
    for (int m = 0; m < 2; ++m) {
        for (auto& sg : s->mate[m].segs) cudaFree(sg.d);
        cudaFree(s->mate[m].d_run); cudaFree(s->mate[m].d_rec_off); cudaFree(s->mate[m].d_rec_len); cudaFree(s->mate[m].d_seq_len);
    }
    cudaFree(s->d_keys); cudaFree(s->d_tile_state); cudaFree(s->d_ctl); cudaFree(s->d_rec_start); cudaFree(s->d_hash);
    if (s->h_ctl) cudaFreeHost(s->h_ctl);
    if (s->ev0) { cudaEventDestroy(s->ev0); cudaEventDestroy(s->ev1); }
    delete s;
}

static int seq_reset(SeqState* s, std::string* err) {
    seq_free_results(s);
    for (u32 m = 0; m < s->mates; ++m) {
        for (auto& sg : s->mate[m].segs) cudaFree(sg.d);
        s->mate[m].segs.clear();
        s->mate[m].n_records = 0; s->mate[m].finished = false;
        SEQ_TRY(cudaMemsetAsync(s->mate[m].d_run, 0, sizeof(RunState), s->stream));
    }
    s->n = s->n_out = 0; s->finished = false; s->ms = 0;
    memset(&s->stats, 0, sizeof s->stats);
    return FQD_OK;
}

static void seq_set_error(SeqState* s, int code, int ch, u64 rec, int mate) {
    if (s->stats.err) return;
    s->stats.err = code; s->stats.err_char = ch; s->stats.err_record = rec; s->stats.err_mate = mate;
}

// Parse the current (last) segment of a mate; move its incomplete tail into a fresh segment unless `final`.
static int seq_parse_segment(SeqState* s, int m, bool final, std::string* err) {
    SeqMate& mt = s->mate[m];
    SeqSegment& sg = mt.segs.back();
    if (sg.fill == 0) return FQD_OK;
    const u32 n_tiles = (u32)((sg.fill + PP_TILE - 1) / PP_TILE);
    const u64 room = s->capacity - mt.n_records;
    if (room == 0) { seq_set_error(s, FQD_ERR_CAPACITY, 0, mt.n_records, m); return FQD_OK; }
    SEQ_TRY(cudaEventRecord(s->ev0, s->stream));
    k_init_chunk<<<std::max(1u, std::min(n_tiles / 256 + 1, 1024u)), 256, 0, s->stream>>>(s->d_ctl, s->d_tile_state, n_tiles);
    ParseParams p;
    p.raw = sg.d; p.n = (u32)sg.fill; p.n_tiles = n_tiles; p.tile_state = s->d_tile_state; p.ctl = s->d_ctl; p.run = mt.d_run;
    p.rec_start = s->d_rec_start; p.cap = (u32)std::min<u64>(s->chunk_cap, room); p.keys = s->d_keys; p.key_capacity = s->capacity;
    p.row_words = s->row_words; p.mate_off = m * s->W; p.W = s->W; p.hash = s->d_hash; p.seq_len = mt.d_seq_len + mt.n_records;
    p.word0 = nullptr; p.dup = nullptr; p.strict = 0; p.hash_salt = m * 4096u;
    if (s->cfg.format == FQD_FORMAT_FASTQ) k_parse_pack<4><<<n_tiles, PP_THREADS, 0, s->stream>>>(p);
    else k_parse_pack<2><<<n_tiles, PP_THREADS, 0, s->stream>>>(p);
    k_finish_segment<<<s->sm * 4, 256, 0, s->stream>>>(s->d_ctl, s->d_rec_start, sg.logical_base, mt.d_rec_off + mt.n_records,
                                                       mt.d_rec_len + mt.n_records, mt.d_run, s->d_ctl);
    k_advance_run<<<1, 1, 0, s->stream>>>(mt.d_run);
    s->launches += 4;
    SEQ_TRY(cudaEventRecord(s->ev1, s->stream));
    SEQ_TRY(cudaMemcpyAsync(s->h_ctl, s->d_ctl, sizeof(ChunkCtl), cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); s->ms += ms;
    const ChunkCtl& c = *s->h_ctl;
    const u64 first = mt.n_records;
    if (c.err_parse != NO_ERR) {
        int code = (int)((c.err_parse >> 8) & 0xFF);
        seq_set_error(s, code == PERR_BAD_START ? FQD_ERR_BAD_START : FQD_ERR_LEN_MISMATCH, (int)(c.err_parse & 0xFF), first + (c.err_parse >> 16), m);
    }
    if (c.too_long == 1) seq_set_error(s, FQD_ERR_SEQ_TOO_LONG, 0, first, m);
    if (c.too_long == 2) seq_set_error(s, FQD_ERR_CAPACITY, 0, first, m);
    if (c.pad) seq_set_error(s, FQD_ERR_UNSUPPORTED_BYTE, 0, first, m);
    mt.n_records += c.n_records;
    const size_t consumed = c.consumed, tail = sg.fill - consumed;
    if (c.n_records == 0 && !final && sg.fill >= sg.cap) {
        *err = "a single record does not fit into one device segment (raise max_chunk_bytes)";
        return FQD_ERR_CAPACITY;
    }
    if (!final) {
        SeqSegment nx;
        nx.cap = s->seg_bytes;
        SEQ_TRY(cudaMalloc(&nx.d, nx.cap + 4096));
        nx.logical_base = sg.logical_base + consumed;
        if (tail) SEQ_TRY(cudaMemcpyAsync(nx.d, sg.d + consumed, tail, cudaMemcpyDeviceToDevice, s->stream));
        nx.fill = tail;
        mt.segs.back().fill = consumed;
        mt.segs.push_back(nx);
    } else {
        if (tail) {
            // an incomplete last record is dropped silently (src/fastqview.cpp:114-115), but its first byte is still
            // checked when the record before it is fetched (src/fastqview.cpp:91-92)
            u8 b = 0;
            SEQ_TRY(cudaMemcpy(&b, sg.d + consumed, 1, cudaMemcpyDeviceToHost));
            const u8 lead = s->cfg.format == FQD_FORMAT_FASTQ ? '@' : '>';
            if (b != lead) seq_set_error(s, FQD_ERR_BAD_START, b, mt.n_records, m);
        }
        mt.segs.back().fill = consumed;
    }
    return FQD_OK;
}

static int seq_append(SeqState* s, int m, const void* buf, size_t n, bool is_device, std::string* err) {
    if (m < 0 || (u32)m >= s->mates) { *err = "bad mate index"; return FQD_ERR_INVALID; }
    if (s->finished) { *err = "fqd_append after fqd_finish"; return FQD_ERR_INVALID; }
    SeqMate& mt = s->mate[m];
    const u8* src = (const u8*)buf;
    while (n) {
        if (mt.segs.empty()) {
            SeqSegment sg; sg.cap = s->seg_bytes;
            SEQ_TRY(cudaMalloc(&sg.d, sg.cap + 4096));
            mt.segs.push_back(sg);
        }
        SeqSegment& sg = mt.segs.back();
        const size_t take = std::min(n, sg.cap - sg.fill);
        if (take) {
            SEQ_TRY(cudaMemcpyAsync(sg.d + sg.fill, src, take, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s->stream));
            sg.fill += take; src += take; n -= take;
        }
        if (sg.fill == sg.cap) {
            if (!is_device) SEQ_TRY(cudaStreamSynchronize(s->stream));     // the caller may reuse its buffer
            int rc = seq_parse_segment(s, m, false, err);
            if (rc) return rc;
        }
    }
    if (!is_device) SEQ_TRY(cudaStreamSynchronize(s->stream));
    return FQD_OK;
}

// ---------------------------------------------------------------------------------------------------------
// sort driver
struct SortScratch {
    u64 *keyA = nullptr, *keyB = nullptr;
    u32 *aA = nullptr, *aB = nullptr, *bA = nullptr, *bB = nullptr;
    u32 *hist = nullptr, *hist_scan = nullptr;
    u64* scan_state = nullptr; u32* ticket = nullptr; u64* d_total = nullptr;
    u64 n_cap = 0; u64 hist_cap = 0; u64 state_cap = 0;
};

template <class T>
static int seq_dalloc(SeqState* s, T** p, size_t count, std::string* err) {
    void* q = nullptr;
    SEQ_TRY(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    s->scratch.push_back(q);
    *p = (T*)q;
    return FQD_OK;
}

static int sort_scratch_alloc(SeqState* s, SortScratch& sc, u64 n, std::string* err) {
    sc.n_cap = n;
    const u64 nblocks = (n + RS_TILE - 1) / RS_TILE;
    sc.hist_cap = 256 * std::max<u64>(nblocks, 1);
    sc.state_cap = std::max<u64>((std::max(sc.hist_cap, n) + SCAN_TILE - 1) / SCAN_TILE + 1, 16);
    int rc;
    if ((rc = seq_dalloc(s, &sc.keyA, n, err))) return rc;
    if ((rc = seq_dalloc(s, &sc.keyB, n, err))) return rc;
    if ((rc = seq_dalloc(s, &sc.aA, n, err))) return rc;
    if ((rc = seq_dalloc(s, &sc.aB, n, err))) return rc;
    if ((rc = seq_dalloc(s, &sc.bA, n, err))) return rc;
    if ((rc = seq_dalloc(s, &sc.bB, n, err))) return rc;
    if ((rc = seq_dalloc(s, &sc.hist, sc.hist_cap, err))) return rc;
    if ((rc = seq_dalloc(s, &sc.hist_scan, sc.hist_cap, err))) return rc;
    if ((rc = seq_dalloc(s, &sc.scan_state, sc.state_cap, err))) return rc;
    if ((rc = seq_dalloc(s, &sc.ticket, 4, err))) return rc;
    if ((rc = seq_dalloc(s, &sc.d_total, 2, err))) return rc;
    return FQD_OK;
}

static inline unsigned seq_grid(SeqState* s, u64 n, int threads = 256) {
    u64 b = (n + threads - 1) / threads;
    return (unsigned)std::max<u64>(1, std::min<u64>(b, (u64)s->sm * 16));
}

// exclusive scan of n u32; when total != nullptr the grand total is copied to the host (synchronises)
static int device_scan(SeqState* s, SortScratch& sc, const u32* in, u32* out, u64 n, u64* total, std::string* err) {
    if (n == 0) { if (total) *total = 0; return FQD_OK; }
    const u64 tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    SEQ_TRY(cudaMemsetAsync(sc.scan_state, 0, tiles * sizeof(u64), s->stream));
    SEQ_TRY(cudaMemsetAsync(sc.ticket, 0, sizeof(u32), s->stream));
    k_scan_exclusive<<<(unsigned)tiles, SCAN_THREADS, 0, s->stream>>>(in, out, n, sc.scan_state, sc.ticket, sc.d_total);
    s->launches++;
    if (total) {
        SEQ_TRY(cudaMemcpyAsync(total, sc.d_total, sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
        SEQ_TRY(cudaStreamSynchronize(s->stream));
    }
    return FQD_OK;
}

// stable LSD radix sort of (key, a[, b]) on key bits [bit_lo, bit_hi); results end up in the *A buffers
static int radix_sort(SeqState* s, SortScratch& sc, u64 n, u32 bit_lo, u32 bit_hi, bool has_b, std::string* err) {
    if (n < 2) return FQD_OK;
    const u32 nblocks = (u32)((n + RS_TILE - 1) / RS_TILE);
    for (u32 shift = bit_lo; shift < bit_hi; shift += 8) {
        k_radix_hist<<<nblocks, RS_THREADS, 0, s->stream>>>(sc.keyA, n, shift, sc.hist, nblocks);
        int rc = device_scan(s, sc, sc.hist, sc.hist_scan, 256ull * nblocks, nullptr, err);
        if (rc) return rc;
        if (has_b) k_radix_scatter<true><<<nblocks, RS_THREADS, 0, s->stream>>>(sc.keyA, sc.aA, sc.bA, sc.keyB, sc.aB, sc.bB, n, shift, sc.hist_scan, nblocks);
        else k_radix_scatter<false><<<nblocks, RS_THREADS, 0, s->stream>>>(sc.keyA, sc.aA, nullptr, sc.keyB, sc.aB, nullptr, n, shift, sc.hist_scan, nblocks);
        s->launches += 2;
        std::swap(sc.keyA, sc.keyB); std::swap(sc.aA, sc.aB);
        if (has_b) std::swap(sc.bA, sc.bB);
    }
    SEQ_TRY(cudaGetLastError());
    return FQD_OK;
}

// Sort record indices [0, n) by rows (words [w_begin, w_begin + n_words) of each row, `word_bits` significant bits
// per word), ties by index.  perm receives the sorted indices.
static int sort_rows(SeqState* s, const u64* rows, u32 stride, u32 w_begin, u32 n_words, u32 word_bits, u64 n, u32* perm, std::string* err) {
    SortScratch sc;
    int rc = sort_scratch_alloc(s, sc, n, err);
    if (rc) return rc;
    u32 *head, *excl, *gid, *gsize, *gdiff, *flag, *posA, *posB;
    if ((rc = seq_dalloc(s, &head, n, err)) || (rc = seq_dalloc(s, &excl, n, err)) || (rc = seq_dalloc(s, &gid, n, err)) ||
        (rc = seq_dalloc(s, &gsize, n, err)) || (rc = seq_dalloc(s, &gdiff, n, err)) || (rc = seq_dalloc(s, &flag, n, err)) ||
        (rc = seq_dalloc(s, &posA, n, err)) || (rc = seq_dalloc(s, &posB, n, err))) return rc;
    const u32 hi_bit = (word_bits + 7) / 8 * 8;

    // round 0: all records by their first word
    k_iota_u32<<<seq_grid(s, n), 256, 0, s->stream>>>(sc.aA, n);
    k_gather_word<<<seq_grid(s, n), 256, 0, s->stream>>>(rows + w_begin, stride, 0, sc.aA, n, sc.keyA);
    s->launches += 2;
    if ((rc = radix_sort(s, sc, n, 0, hi_bit, false, err))) return rc;
    SEQ_TRY(cudaMemcpyAsync(perm, sc.aA, n * sizeof(u32), cudaMemcpyDeviceToDevice, s->stream));

    u64 n_act = n;                  // items in keyA/aA (sorted by the words used so far), bA = group of each
    const u32* pos_in = nullptr;    // their positions in perm (nullptr = identity)
    const u32* seg_in = nullptr;
    for (u32 w = 0; w < n_words && n_act > 1; ++w) {
        // groups of equal (previous group, word w)
        k_mark_heads<<<seq_grid(s, n_act), 256, 0, s->stream>>>(sc.keyA, seg_in, n_act, head);
        u64 n_groups = 0;
        if ((rc = device_scan(s, sc, head, excl, n_act, &n_groups, err))) return rc;
        SEQ_TRY(cudaMemsetAsync(gsize, 0, n_groups * sizeof(u32), s->stream));
        SEQ_TRY(cudaMemsetAsync(gdiff, 0, n_groups * sizeof(u32), s->stream));
        k_group_ids<<<seq_grid(s, n_act), 256, 0, s->stream>>>(head, excl, n_act, gid, gsize);
        s->launches += 2;
        if (w + 1 >= n_words) break;      // every word used: remaining ties are identical rows, already in index order
        k_mark_unresolved<<<seq_grid(s, n_act), 256, 0, s->stream>>>(rows + w_begin, stride, w + 1, n_words, sc.aA, gid, n_act, gdiff);
        k_active_flags<<<seq_grid(s, n_act), 256, 0, s->stream>>>(gid, gsize, gdiff, n_act, flag);
        u64 n_next = 0;
        if ((rc = device_scan(s, sc, flag, excl, n_act, &n_next, err))) return rc;
        s->launches += 2;
        if (n_next == 0) break;
        // compact the unresolved items (position in perm, record index, group id), keeping their order
        k_compact_active<<<seq_grid(s, n_act), 256, 0, s->stream>>>(flag, excl, pos_in, sc.aA, gid, n_act, posB, sc.aB, sc.bB);
        std::swap(sc.aA, sc.aB); std::swap(sc.bA, sc.bB); std::swap(posA, posB);
        pos_in = posA;
        n_act = n_next;
        // sort them by (group, word w+1): LSD = word first, then group
        k_gather_word<<<seq_grid(s, n_act), 256, 0, s->stream>>>(rows + w_begin, stride, w + 1, sc.aA, n_act, sc.keyA);
        s->launches += 2;
        if ((rc = radix_sort(s, sc, n_act, 0, hi_bit, true, err))) return rc;
        u32 gbits = 1; while ((1ull << gbits) < n_groups) ++gbits;
        k_u32_to_u64key<<<seq_grid(s, n_act), 256, 0, s->stream>>>(sc.bA, n_act, sc.keyA);
        if ((rc = radix_sort(s, sc, n_act, 0, (gbits + 7) / 8 * 8, true, err))) return rc;
        k_scatter_perm<<<seq_grid(s, n_act), 256, 0, s->stream>>>(posA, sc.aA, n_act, perm);
        // keys for the next grouping step: word w+1 of the newly ordered items, groups in bA
        k_gather_word<<<seq_grid(s, n_act), 256, 0, s->stream>>>(rows + w_begin, stride, w + 1, sc.aA, n_act, sc.keyA);
        s->launches += 3;
        seg_in = sc.bA;
    }
    SEQ_TRY(cudaGetLastError());
    return FQD_OK;
}

// ---------------------------------------------------------------------------------------------------------
static int seq_finish_sequence_mode(SeqState* s, std::string* err) {
    const u64 n = s->n;
    int rc;
    u32* perm;
    if ((rc = seq_dalloc(s, &perm, n, err))) return rc;
    if ((rc = sort_rows(s, s->d_keys, s->row_words, 0, s->row_words, 60, n, perm, err))) return rc;

    u32 *keep, *excl, *brk = nullptr;
    if ((rc = seq_dalloc(s, &keep, n, err)) || (rc = seq_dalloc(s, &excl, n, err))) return rc;
    ScanParams sp;
    sp.rows = s->d_keys; sp.stride = s->row_words; sp.W = s->W; sp.mates = s->mates;
    sp.len0 = s->mate[0].d_seq_len; sp.len1 = s->mates == 2 ? s->mate[1].d_seq_len : nullptr;
    sp.perm = perm; sp.n = n; sp.dist = s->cfg.hamming_dist; sp.keep = keep; sp.brk = nullptr;
    if (s->cfg.mode == FQD_MODE_SEQ_TIGHT) k_scan_tight<<<seq_grid(s, n), 256, 0, s->stream>>>(sp);
    else if (s->cfg.mode == FQD_MODE_SEQ_LOOSE) k_scan_loose<<<seq_grid(s, n), 256, 0, s->stream>>>(sp);
    else {
        if ((rc = seq_dalloc(s, &brk, n, err))) return rc;
        sp.brk = brk;
        k_ham_breaks<<<seq_grid(s, n), 256, 0, s->stream>>>(sp);
        k_ham_segments<<<seq_grid(s, n), 256, 0, s->stream>>>(sp);
        s->launches++;
    }
    s->launches++;
    SortScratch sc;      // only the scan scratch is used here
    if ((rc = seq_dalloc(s, &sc.scan_state, (n + SCAN_TILE - 1) / SCAN_TILE + 16, err)) || (rc = seq_dalloc(s, &sc.ticket, 4, err)) ||
        (rc = seq_dalloc(s, &sc.d_total, 2, err))) return rc;
    u64 n_out = 0;
    if ((rc = device_scan(s, sc, keep, excl, n, &n_out, err))) return rc;
    s->n_out = n_out;
    u64 *o_off[2] = {nullptr, nullptr}; u32 *o_len[2] = {nullptr, nullptr}; u32* o_idx;
    if ((rc = seq_dalloc(s, &o_idx, n_out, err))) return rc;
    for (u32 m = 0; m < s->mates; ++m)
        if ((rc = seq_dalloc(s, &o_off[m], n_out, err)) || (rc = seq_dalloc(s, &o_len[m], n_out, err))) return rc;
    k_emit<<<seq_grid(s, n), 256, 0, s->stream>>>(keep, excl, perm, n, s->mate[0].d_rec_off, s->mate[0].d_rec_len,
                                                  s->mates == 2 ? s->mate[1].d_rec_off : nullptr, s->mates == 2 ? s->mate[1].d_rec_len : nullptr,
                                                  o_off[0], o_len[0], o_off[1], o_len[1], o_idx);
    s->launches++;
    for (u32 m = 0; m < s->mates; ++m) {
        s->h_off[m].resize(n_out); s->h_len[m].resize(n_out);
        if (n_out) {
            SEQ_TRY(cudaMemcpyAsync(s->h_off[m].data(), o_off[m], n_out * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
            SEQ_TRY(cudaMemcpyAsync(s->h_len[m].data(), o_len[m], n_out * sizeof(u32), cudaMemcpyDeviceToHost, s->stream));
        }
    }
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    s->stats.total = n;
    s->stats.dups = n - n_out;
    return FQD_OK;
}

static int seq_finish_unordered(SeqState* s, std::string* err);

static int seq_finish(SeqState* s, std::string* err) {
    if (s->finished) { *err = "fqd_finish called twice"; return FQD_ERR_INVALID; }
    s->finished = true;
    for (u32 m = 0; m < s->mates; ++m) {
        SeqMate& mt = s->mate[m];
        if (mt.segs.empty()) { seq_set_error(s, FQD_ERR_EMPTY, 0, 0, m); continue; }
        // a segment may hold more records than one parse takes (chunk_cap): keep parsing its carried tail
        for (;;) {
            const u64 before = mt.n_records;
            int rc = seq_parse_segment(s, m, false, err);
            if (rc) return rc;
            if (mt.n_records == before || mt.segs.back().fill == 0 || s->stats.err) break;
        }
        int rc = seq_parse_segment(s, m, true, err);
        if (rc) return rc;
    }
    if (s->stats.err) return FQD_OK;            // data error: reported through fqd_stats, like the fast mode
    u64 n = s->mate[0].n_records;
    if (s->mates == 2 && !s->cfg.unordered) n = std::min(n, s->mate[1].n_records);      // stops at the shorter file
    if (n == 0 || (s->mates == 2 && s->mate[1].n_records == 0)) { seq_set_error(s, FQD_ERR_EMPTY, 0, 0, 0); return FQD_OK; }
    s->n = n;
    SEQ_TRY(cudaEventRecord(s->ev0, s->stream));
    int rc = s->cfg.unordered ? seq_finish_unordered(s, err) : seq_finish_sequence_mode(s, err);
    if (rc) return rc;
    SEQ_TRY(cudaEventRecord(s->ev1, s->stream));
    SEQ_TRY(cudaEventSynchronize(s->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); s->ms += ms;
    return FQD_OK;
}

static int seq_finish_unordered(SeqState* s, std::string* err) {
    (void)s;
    *err = "--unordered is not built yet";
    return FQD_ERR_INVALID;
}

static int seq_emission(SeqState* s, fqd_emission_t* out, std::string* err) {
    if (!s->finished) { *err = "fqd_emission before fqd_finish"; return FQD_ERR_INVALID; }
    memset(out, 0, sizeof *out);
    out->n_out = s->n_out;
    for (u32 m = 0; m < s->mates; ++m) { out->off[m] = (const uint64_t*)s->h_off[m].data(); out->len[m] = s->h_len[m].data(); }
    return FQD_OK;
}
static void seq_stats(SeqState* s, fqd_stats_t* st) { *st = s->stats; }
static double seq_device_ms(SeqState* s) { return s->ms; }
static u64 seq_launches(SeqState* s) { return s->launches; }

}  // namespace fqd
