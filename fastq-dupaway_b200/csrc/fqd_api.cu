// fqd_api.cu - C ABI of libfqd_cuda.so (declared in include/fqd.h).  Host side of the device pipeline:
// owns the CUDA stream, the key store, the hash table and the per-chunk tables; launches the kernels.
// There is no CPU fallback anywhere in this file: every data path runs the sm_100a kernels.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/fqd.h"
#include "common.cuh"
#include "hashset.cuh"
#include "parse_pack.cuh"
#include "seqmode.cuh"
#include "shard.cuh"
#include "shard2.cuh"
#include "synth.cuh"

using namespace fqd;

static thread_local std::string g_create_error;

struct MateChunk {
    u8* d_raw = nullptr;          // staging for host pushes (max_chunk_bytes + slack)
    u64* d_tile_state = nullptr;
    ChunkCtl* d_ctl = nullptr;
    u32* d_rec_start = nullptr;   // [cap+1]
    u64* d_hash = nullptr;        // [cap]
    u32* h_rec_start = nullptr;   // pinned
    ChunkCtl* h_ctl = nullptr;    // pinned
    u32 n_tiles_cap = 0;
};

struct Shard2State {
    u32 N = 0, me = 0, region_rows = 0, n_blocks_cap = 0;
    u32 *d_block_cnt = nullptr, *d_block_base = nullptr, *d_dest[2] = {nullptr, nullptr}, *d_totals = nullptr;
    u64 *d_final_hash = nullptr, *d_stage_keys = nullptr;
    RunState *d_stage_run = nullptr, *d_run2 = nullptr, *h_run2 = nullptr;
    u64* d_hash_regions = nullptr;      // [2][N * region_rows]   written by my peers (and me)
    u32* d_counts = nullptr;            // [2][S2_MAX]            rows per source, written by my peers
    u8* d_flags = nullptr;              // [N * region_rows]      what k_insert2 decides
    u8* d_flags_in = nullptr;           // [2][N * region_rows]   flags of MY records, written by their owners
    unsigned long long* d_ndups = nullptr;
    u32* h_totals = nullptr;
    u32* d_chunk_n = nullptr;           // [2] records of the chunk packed with this parity
    u64* d_stage_rows[2] = {nullptr, nullptr};   // rows of a chunk grouped by owner, laid out like the owners' regions
    u64* d_stage_hash[2] = {nullptr, nullptr};
    cudaStream_t s_pack = nullptr, s_ins = nullptr, s_copy = nullptr;
    cudaEvent_t ev_scatter[2] = {nullptr, nullptr}, ev_flags[2] = {nullptr, nullptr}, ev_staged[2] = {nullptr, nullptr};
    // the same two, as plain events, for peers of the SAME process (waiting on an interprocess-capable event of another
    // device from inside its own process crashes driver 580.159 in cuStreamWaitEvent)
    cudaEvent_t lv_scatter[2] = {nullptr, nullptr}, lv_flags[2] = {nullptr, nullptr};
    u64* peer_keys[S2_MAX] = {}; u64* peer_hash[S2_MAX] = {}; u32* peer_counts[S2_MAX] = {}; u8* peer_flags_in[S2_MAX] = {};
    cudaEvent_t peer_scatter[S2_MAX][2] = {}, peer_flags[S2_MAX][2] = {};
    bool imported[S2_MAX] = {}, linked[S2_MAX] = {};      // linked: a handle of this process (nothing to close)
    u64 chunks_packed = 0, chunks_inserted = 0, chunks_applied = 0;
    u32 last_n = 0;
    cudaEvent_t t_ins = nullptr;
};

struct fqd_handle {
    fqd_config cfg;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int sm_count = 148;
    u32 W = 0, row_words = 0;
    u64 key_capacity = 0;
    u64* d_keys = nullptr;
    u64* d_table = nullptr;
    u64 n_buckets = 0;
    u32 bucket_shift = 0;
    RunState* d_run = nullptr;
    RunState* h_run = nullptr;    // pinned
    MateChunk mate[2];
    u32 cap = 0;                  // records per chunk
    u8* d_dup = nullptr;
    u8* h_dup = nullptr;          // pinned
    fqd_stats_t stats;
    std::string err;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending_events;
    std::vector<cudaEvent_t> event_pool;
    double device_ms = 0.0;
    u64 launches = 0;
    bool pending_async = false;
    // fqd_push_prefetch / fqd_push_staged: the next chunk's host-to-device copy runs on its own stream into the
    // alternate raw buffers while the current chunk is being processed
    u8* d_raw_alt[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_done[2] = {nullptr, nullptr};
    size_t staged_n[2][2] = {{0, 0}, {0, 0}};      // [slot][mate]
    int pf_slot = 0, run_slot = 0, n_staged = 0;
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
    bool profile = false;
    fqd_profile_t prof;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_parse, prof_insert, prof_scatter;
    SeqState* seq = nullptr;      // whole-input modes (seqmode.cuh)
    // multi-GPU --fast mode (shard.cuh)
    SeqState* shard_ctx = nullptr;   // stream / scratch bookkeeping for the sort primitives
    SortScratch shard_sc;
    u64* d_stage_keys = nullptr;     // [cap * row_words] packed keys of the chunk being exchanged
    RunState* d_stage_run = nullptr; // zero: K1 writes the chunk's rows at slot 0
    u64* d_final_hash = nullptr;     // [cap]
    u32* d_counts = nullptr;         // [n_shards + 1]
    u64* d_recv_hash = nullptr;      // [cap_recv]
    u64 recv_cap = 0;
    u64 shard_last_n = 0;
    u32 grows = 0;                   // times the key store / table were grown in place
    struct Shard2State* s2 = nullptr; // multi-GPU --fast mode, round 2 (shard2.cuh)
    u64* d_surv = nullptr;           // fqd_keep_survivors: global indices of the written records, ascending
    u32* d_surv_cnt = nullptr;       // survivors per SV_BLOCK records of the current chunk
};

#define CUDA_TRY(h, call)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            char b_[512];                                                                      \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            if (h) (h)->err = b_; else g_create_error = b_;                                    \
            return FQD_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

static int fail(fqd_handle* h, int code, const char* msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

extern "C" int fqd_abi_version(void) { return FQD_ABI_VERSION; }

extern "C" const char* fqd_last_error(const fqd_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int fqd_host_alloc(void** p, size_t bytes) {
    return cudaHostAlloc(p, bytes, cudaHostAllocDefault) == cudaSuccess ? FQD_OK : FQD_ERR_CUDA;
}
extern "C" int fqd_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? FQD_OK : FQD_ERR_CUDA; }

extern "C" int fqd_device_alloc(int device, void** d_ptr, size_t bytes) {
    if (cudaSetDevice(device) != cudaSuccess) return FQD_ERR_CUDA;
    return cudaMalloc(d_ptr, bytes) == cudaSuccess ? FQD_OK : FQD_ERR_CUDA;
}
extern "C" int fqd_device_free(int device, void* d_ptr) {
    if (cudaSetDevice(device) != cudaSuccess) return FQD_ERR_CUDA;
    return cudaFree(d_ptr) == cudaSuccess ? FQD_OK : FQD_ERR_CUDA;
}
extern "C" int fqd_memcpy_d2h(int device, void* dst, const void* d_src, size_t bytes) {
    if (cudaSetDevice(device) != cudaSuccess) return FQD_ERR_CUDA;
    return cudaMemcpy(dst, d_src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? FQD_OK : FQD_ERR_CUDA;
}
extern "C" int fqd_memcpy_h2d(int device, void* d_dst, const void* src, size_t bytes) {
    if (cudaSetDevice(device) != cudaSuccess) return FQD_ERR_CUDA;
    // cudaMemcpy from pageable memory may return once the data sits in the staging buffer, before the DMA to the
    // device has finished; it is ordered only against the default stream, and the handles use non-blocking streams.
    if (cudaMemcpy(d_dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return FQD_ERR_CUDA;
    return cudaDeviceSynchronize() == cudaSuccess ? FQD_OK : FQD_ERR_CUDA;
}

// ---- peer-memory exchange (multi-GPU, one process per GPU): a rank exports a device buffer, its peers map it and
// write their part of an all-to-all straight into it over NVLink with the copy engines (cudaMemcpyAsync on mapped
// peer memory), instead of NCCL's SM-driven send/recv kernels.
extern "C" int fqd_ipc_export(int device, void* d_ptr, void* handle64) {
    if (cudaSetDevice(device) != cudaSuccess || !d_ptr || !handle64) return FQD_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    return cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, d_ptr) == cudaSuccess ? FQD_OK : FQD_ERR_CUDA;
}
extern "C" int fqd_ipc_open(int device, const void* handle64, void** d_ptr) {
    if (cudaSetDevice(device) != cudaSuccess || !d_ptr || !handle64) return FQD_ERR_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    return cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? FQD_OK : FQD_ERR_CUDA;
}
extern "C" int fqd_ipc_close(int device, void* d_ptr) {
    if (cudaSetDevice(device) != cudaSuccess) return FQD_ERR_CUDA;
    return cudaIpcCloseMemHandle(d_ptr) == cudaSuccess ? FQD_OK : FQD_ERR_CUDA;
}
extern "C" int fqd_peer_copy_async(int device, void* d_dst, const void* d_src, size_t bytes, void* cuda_stream) {
    if (cudaSetDevice(device) != cudaSuccess) return FQD_ERR_CUDA;
    if (bytes == 0) return FQD_OK;
    return cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, (cudaStream_t)cuda_stream) == cudaSuccess ? FQD_OK : FQD_ERR_CUDA;
}

static void table_geometry(u64 capacity, u64* n_buckets, u32* shift);
static u32 words_for(u32 max_seq_len, bool byte_keys = false) {
    u32 w = byte_keys ? (max_seq_len + 1 + 7) / 8 : (max_seq_len + BASES_PER_WORD - 1) / BASES_PER_WORD;
    if (w < 2) w = 2;
    return (w + 1u) & ~1u;      // even, so rows are 16-byte aligned
}

static void shard2_free(fqd_handle* h);

extern "C" void fqd_destroy(fqd_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    shard2_free(h);
    if (h->seq) seq_destroy(h->seq);
    if (h->shard_ctx) {
        seq_free_results(h->shard_ctx); delete h->shard_ctx;
        cudaFree(h->d_stage_keys); cudaFree(h->d_stage_run); cudaFree(h->d_final_hash); cudaFree(h->d_counts); cudaFree(h->d_recv_hash);
    }
    if (h->timer0) { cudaEventDestroy(h->timer0); cudaEventDestroy(h->timer1); }
    for (auto& pe : h->prof_parse) { cudaEventDestroy(pe.first); cudaEventDestroy(pe.second); }
    for (auto& pe : h->prof_insert) { cudaEventDestroy(pe.first); cudaEventDestroy(pe.second); }
    for (auto& pe : h->prof_scatter) { cudaEventDestroy(pe.first); cudaEventDestroy(pe.second); }
    for (auto& pe : h->pending_events) { cudaEventDestroy(pe.first); cudaEventDestroy(pe.second); }
    for (auto e : h->event_pool) cudaEventDestroy(e);
    for (int m = 0; m < 2; ++m) {
        MateChunk& c = h->mate[m];
        cudaFree(c.d_raw); cudaFree(c.d_tile_state); cudaFree(c.d_ctl); cudaFree(c.d_rec_start); cudaFree(c.d_hash);
        if (c.h_rec_start) cudaFreeHost(c.h_rec_start);
        if (c.h_ctl) cudaFreeHost(c.h_ctl);
    }
    cudaFree(h->d_keys); cudaFree(h->d_table); cudaFree(h->d_run); cudaFree(h->d_dup); cudaFree(h->d_surv); cudaFree(h->d_surv_cnt);
    if (h->h_run) cudaFreeHost(h->h_run);
    if (h->h_dup) cudaFreeHost(h->h_dup);
    for (int k = 0; k < 2; ++k) { cudaFree(h->d_raw_alt[k]); if (h->copy_done[k]) cudaEventDestroy(h->copy_done[k]); }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

static int create_impl(const fqd_config* cfg, fqd_handle* h) {
    h->cfg = *cfg;
    memset(&h->stats, 0, sizeof h->stats);
    memset(&h->prof, 0, sizeof h->prof);
    int ndev = 0;
    CUDA_TRY(h, cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(h, FQD_ERR_CUDA, "no such CUDA device");
    CUDA_TRY(h, cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_TRY(h, cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(h, FQD_ERR_CUDA, "libfqd_cuda is built for sm_100a (B200) only");
    h->sm_count = prop.multiProcessorCount;
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_TRY(h, pp_init_tables());
    CUDA_TRY(h, cudaFuncSetAttribute(k_parse_pack<4, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_TRY(h, cudaFuncSetAttribute(k_parse_pack<2, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_TRY(h, cudaFuncSetAttribute(k_parse_pack<4, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_TRY(h, cudaFuncSetAttribute(k_parse_pack<2, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));

    const int mates = cfg->paired ? 2 : 1;
    if (cfg->byte_keys && (cfg->mode == FQD_MODE_FAST || cfg->unordered))
        return fail(h, FQD_ERR_INVALID, "byte_keys is for sequence-based modes (--fast accepts {A,C,G,T,N} only, src/seq_utils.cpp:3-21)");
    h->W = words_for(cfg->max_seq_len ? cfg->max_seq_len : 150, cfg->byte_keys != 0);
    h->row_words = h->W * mates;
    if (cfg->max_chunk_bytes == 0 || cfg->max_chunk_bytes >= (1ull << 32) - (1ull << 20))
        return fail(h, FQD_ERR_INVALID, "max_chunk_bytes must be in (0, 4 GiB - 1 MiB)");
    h->cap = (u32)(cfg->max_chunk_records ? cfg->max_chunk_records : std::max<u64>(cfg->max_chunk_bytes / 64, 1024));
    const bool whole_input = cfg->mode != FQD_MODE_FAST || cfg->unordered;

    CUDA_TRY(h, cudaMalloc(&h->d_run, sizeof(RunState)));
    CUDA_TRY(h, cudaMemsetAsync(h->d_run, 0, sizeof(RunState), h->stream));
    CUDA_TRY(h, cudaHostAlloc(&h->h_run, sizeof(RunState), cudaHostAllocDefault));

    if (!whole_input) {
        if (cfg->max_records == 0) return fail(h, FQD_ERR_INVALID, "max_records must be > 0");
        h->key_capacity = cfg->max_records;
        CUDA_TRY(h, cudaMalloc(&h->d_keys, h->key_capacity * h->row_words * sizeof(u64)));
        u64 nb;
        table_geometry(h->key_capacity, &nb, &h->bucket_shift);
        h->n_buckets = nb;
        CUDA_TRY(h, cudaMalloc(&h->d_table, nb * 4 * sizeof(u64)));
        CUDA_TRY(h, cudaMemsetAsync(h->d_table, 0xFF, nb * 4 * sizeof(u64), h->stream));
        for (int m = 0; m < mates; ++m) {
            MateChunk& c = h->mate[m];
            c.n_tiles_cap = (u32)((cfg->max_chunk_bytes + PP_TILE - 1) / PP_TILE) + 1;
            CUDA_TRY(h, cudaMalloc(&c.d_raw, cfg->max_chunk_bytes + 4096));
            CUDA_TRY(h, cudaMalloc(&c.d_tile_state, (size_t)c.n_tiles_cap * sizeof(u64)));
            CUDA_TRY(h, cudaMalloc(&c.d_ctl, sizeof(ChunkCtl)));
            CUDA_TRY(h, cudaMalloc(&c.d_rec_start, ((size_t)h->cap + 1) * sizeof(u32)));
            CUDA_TRY(h, cudaMalloc(&c.d_hash, (size_t)h->cap * sizeof(u64)));
            CUDA_TRY(h, cudaHostAlloc(&c.h_rec_start, ((size_t)h->cap + 1) * sizeof(u32), cudaHostAllocDefault));
            CUDA_TRY(h, cudaHostAlloc(&c.h_ctl, sizeof(ChunkCtl), cudaHostAllocDefault));
        }
        CUDA_TRY(h, cudaMalloc(&h->d_dup, (size_t)h->cap + 64));
        CUDA_TRY(h, cudaHostAlloc(&h->h_dup, (size_t)h->cap + 64, cudaHostAllocDefault));
    } else {
        int rc = seq_create(&h->seq, cfg, h->stream, h->sm_count, h->W, &h->err);
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return FQD_OK;
}

extern "C" int fqd_create(const fqd_config* cfg, fqd_handle** out) {
    if (!cfg || !out) return fail(nullptr, FQD_ERR_INVALID, "null argument");
    if (cfg->abi_version != FQD_ABI_VERSION) return fail(nullptr, FQD_ERR_INVALID, "ABI version mismatch");
    if (cfg->unordered && (cfg->mode != FQD_MODE_FAST || !cfg->paired))
        return fail(nullptr, FQD_ERR_INVALID, "--unordered needs --fast and paired input");   // src/main.cpp:158-164
    fqd_handle* h = new fqd_handle();
    int rc = create_impl(cfg, h);
    if (rc) { g_create_error = h->err; fqd_destroy(h); *out = nullptr; return rc; }
    *out = h;
    return FQD_OK;
}

// -----------------------------------------------------------------------------------------------------------
static cudaEvent_t get_event(fqd_handle* h) {
    if (!h->event_pool.empty()) { cudaEvent_t e = h->event_pool.back(); h->event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}

// Enqueue parse+pack for one mate of a chunk.
static int launch_parse(fqd_handle* h, int m, const u8* d_raw, size_t n, u8* d_dup_unused) {
    (void)d_dup_unused;
    MateChunk& c = h->mate[m];
    const u32 n_tiles = (u32)((n + PP_TILE - 1) / PP_TILE);
    k_init_chunk<<<std::max(1u, std::min(n_tiles / 256 + 1, 1024u)), 256, 0, h->stream>>>(c.d_ctl, c.d_tile_state, n_tiles);
    h->launches++;
    if (n_tiles == 0) return FQD_OK;
    ParseParams p;
    p.raw = d_raw; p.n = (u32)n; p.n_tiles = n_tiles; p.tile_state = c.d_tile_state; p.ctl = c.d_ctl; p.run = h->d_run;
    p.rec_start = c.d_rec_start; p.cap = h->cap; p.keys = h->d_keys; p.key_capacity = h->key_capacity;
    p.row_words = h->row_words; p.mate_off = m * h->W; p.W = h->W; p.hash = c.d_hash; p.seq_len = nullptr; p.word0 = nullptr;
    p.strict = 1; p.hash_salt = m * 4096u; p.bad_rec = nullptr; p.dup = (m == 0) ? h->d_dup : nullptr; p.byte_keys = 0; p.skip = 0;
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (h->profile) { pe0 = get_event(h); pe1 = get_event(h); cudaEventRecord(pe0, h->stream); }
    pp_launch(h->cfg.format == FQD_FORMAT_FASTQ, p, h->stream);
    h->launches += 2;                 // + the two head kernels of pp_launch
    if (h->profile) { cudaEventRecord(pe1, h->stream); h->prof_parse.emplace_back(pe0, pe1); h->prof.parse_launches++; h->prof.parse_bytes += n; }
    h->launches++;
    return FQD_OK;
}

static int enqueue_fast_chunk(fqd_handle* h, const void* d_r1, size_t n1, const void* d_r2, size_t n2) {
    const bool paired = h->cfg.paired != 0;
    if (((uintptr_t)d_r1 & 15) || (paired && ((uintptr_t)d_r2 & 15))) return fail(h, FQD_ERR_INVALID, "device buffers must be 16-byte aligned");
    if (n1 > h->cfg.max_chunk_bytes || n2 > h->cfg.max_chunk_bytes) return fail(h, FQD_ERR_INVALID, "chunk larger than max_chunk_bytes");
    cudaEvent_t e0 = get_event(h), e1 = get_event(h);
    CUDA_TRY(h, cudaEventRecord(e0, h->stream));
    launch_parse(h, 0, (const u8*)d_r1, n1, nullptr);
    if (paired) launch_parse(h, 1, (const u8*)d_r2, n2, nullptr);
    InsertParams ip;
    ip.table = h->d_table; ip.bucket_shift = h->bucket_shift; ip.bucket_mask = h->n_buckets - 1; ip.keys = h->d_keys;
    ip.row_words = h->row_words; ip.key_capacity = h->key_capacity;
    ip.hash1 = h->mate[0].d_hash; ip.hash2 = paired ? h->mate[1].d_hash : nullptr;
    ip.ctl1 = h->mate[0].d_ctl; ip.ctl2 = paired ? h->mate[1].d_ctl : nullptr; ip.run = h->d_run; ip.dup = h->d_dup;
    ip.hash_mul = 1; ip.hash_final = 0;
    k_chunk_begin<<<1, 1, 0, h->stream>>>(ip);
    cudaEvent_t ie0 = nullptr, ie1 = nullptr;
    if (h->profile) { ie0 = get_event(h); ie1 = get_event(h); cudaEventRecord(ie0, h->stream); }
    insert_launch(ip, h->sm_count * 8, h->stream);
    if (h->profile) { cudaEventRecord(ie1, h->stream); h->prof_insert.emplace_back(ie0, ie1); h->prof.insert_launches++; }
    k_count_dups<<<h->sm_count * 2, HS_THREADS, 0, h->stream>>>(h->d_dup, h->d_run);
    if (h->d_surv) {
        const unsigned sb = (h->cap + SV_BLOCK - 1) / SV_BLOCK;
        k_surv_count<<<sb, HS_THREADS, 0, h->stream>>>(h->d_dup, h->d_run, h->d_surv_cnt);
        k_surv_write<<<sb, HS_THREADS, 0, h->stream>>>(h->d_dup, h->d_run, h->d_surv_cnt, h->d_surv, h->key_capacity);
        h->launches += 2;
    }
    k_chunk_end<<<1, 1, 0, h->stream>>>(h->d_run, h->mate[0].d_ctl, paired ? h->mate[1].d_ctl : nullptr);
    h->launches += 4;
    CUDA_TRY(h, cudaEventRecord(e1, h->stream));
    h->pending_events.emplace_back(e0, e1);
    CUDA_TRY(h, cudaGetLastError());
    return FQD_OK;
}

static int drain_events(fqd_handle* h) {
    for (auto& pe : h->prof_parse) {
        float ms = 0.f; CUDA_TRY(h, cudaEventElapsedTime(&ms, pe.first, pe.second)); h->prof.parse_ms += ms;
        h->event_pool.push_back(pe.first); h->event_pool.push_back(pe.second);
    }
    h->prof_parse.clear();
    for (auto& pe : h->prof_insert) {
        float ms = 0.f; CUDA_TRY(h, cudaEventElapsedTime(&ms, pe.first, pe.second)); h->prof.insert_ms += ms;
        h->event_pool.push_back(pe.first); h->event_pool.push_back(pe.second);
    }
    h->prof_insert.clear();
    for (auto& pe : h->prof_scatter) {
        float ms = 0.f; CUDA_TRY(h, cudaEventElapsedTime(&ms, pe.first, pe.second)); h->prof.scatter_ms += ms;
        h->event_pool.push_back(pe.first); h->event_pool.push_back(pe.second);
    }
    h->prof_scatter.clear();
    for (auto& pe : h->pending_events) {
        float ms = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, pe.first, pe.second));
        h->device_ms += ms;
        h->event_pool.push_back(pe.first);
        h->event_pool.push_back(pe.second);
    }
    h->pending_events.clear();
    return FQD_OK;
}

// Fold the device counters / error words of the chunk that just finished into the host statistics.
// Order of errors follows the reference's lazy record fetch (src/bufferedinput.hpp:90-103): fetching pair k
// pre-parses record k+1 of each mate (left first), so a malformed record e aborts before pair e-1 is
// processed; a bad base in pair j aborts while pair j is being keyed (left mate first).
static int fold_chunk(fqd_handle* h, u64 first_record, u64* n_ok) {
    const int mates = h->cfg.paired ? 2 : 1;
    u64 pairs = h->h_run->chunk_pairs;
    *n_ok = pairs;
    if (h->stats.err) { *n_ok = 0; return h->stats.err; }
    long long best_t = -1; int best_code = 0, best_char = 0, best_mate = 0; u64 best_rec = 0, best_ok = 0;
    bool have = false;
    u64 first_too_long = ~0ull;      // chunk-local index of the first record whose sequence does not fit the key rows
    for (int m = 0; m < mates; ++m) {
        const ChunkCtl& c = *h->mate[m].h_ctl;
        if (c.err_parse != NO_ERR) {
            u64 e = c.err_parse >> 16; int code = (int)((c.err_parse >> 8) & 0xFF); int ch = (int)(c.err_parse & 0xFF);
            if (e <= pairs) {     // a malformed record right after the last processed pair still aborts it
                long long t = e == 0 ? (first_record == 0 ? -8 + m : -4 + m) : (long long)(e - 1) * 4 + m;
                if (!have || t < best_t) {
                    have = true; best_t = t; best_char = ch; best_rec = first_record + e; best_mate = m;
                    best_code = code == PERR_BAD_START ? FQD_ERR_BAD_START : FQD_ERR_LEN_MISMATCH;
                    best_ok = e == 0 ? 0 : e - 1;
                }
            }
        }
        if (c.err_base != NO_ERR) {
            u64 j = c.err_base >> 32; int ch = (int)(c.err_base & 0xFF);
            if (j < pairs) {
                long long t = (long long)j * 4 + 2 + m;
                if (!have || t < best_t) {
                    have = true; best_t = t; best_char = ch; best_rec = first_record + j; best_code = FQD_ERR_BAD_BASE; best_ok = j; best_mate = m;
                }
            }
        }
        if ((c.too_long & TL_SEQ) && c.too_long_rec < pairs) first_too_long = std::min<u64>(first_too_long, c.too_long_rec);
    }
    // A record that is too long for the key rows was not packed: flags at and after it are not valid.  The job must be run
    // again with wider rows unless the reference would have stopped BEFORE that record anyway.
    if (first_too_long != ~0ull && (!have || first_too_long <= best_rec - first_record)) {
        have = true; best_code = FQD_ERR_SEQ_TOO_LONG; best_char = 0; best_rec = first_record + first_too_long; best_mate = 0; best_ok = 0;
    }
    if (h->h_run->capacity_exceeded && !have) { have = true; best_code = FQD_ERR_CAPACITY; best_ok = pairs; }
    if (have) {
        h->stats.err = best_code; h->stats.err_char = best_char; h->stats.err_record = best_rec; h->stats.err_mate = best_mate;
        *n_ok = best_ok;
    }
    return h->stats.err;
}

static int fold_fast_chunk(fqd_handle* h, size_t n1, size_t n2, fqd_chunk_result* res);

static void table_geometry(u64 capacity, u64* n_buckets, u32* shift) {
    u64 nb = 1024;
    while (nb * 2 < capacity) nb <<= 1;        // 4*nb entries >= 2*capacity  (load factor <= 0.5)
    u32 lg = 0; while ((1ull << lg) < nb) ++lg;
    *n_buckets = nb; *shift = 64 - lg;
}

// In-place growth of the key store and the table (the reference's set simply rehashes, src/hash_dup_remover.hpp:113-114):
// called when a chunk did not fit (k_chunk_begin refused it); the caller then runs the same chunk again.  Returns
// FQD_ERR_CAPACITY when the device has no room for the larger store next to the old one.
static int grow_fast(fqd_handle* h, u64 need) {
    const u64 used = h->h_run->n_records;
    u64 cap = std::max<u64>(h->key_capacity * 2, used + 2 * need);
    const size_t row_bytes = (size_t)h->row_words * sizeof(u64);
    u64 nb; u32 shift;
    u64* keys = nullptr; u64* table = nullptr;
    for (;; cap = used + need + (cap - used - need) / 2) {     // not enough memory for 2x: try less head room
        table_geometry(cap, &nb, &shift);
        const bool same_table = nb == h->n_buckets;
        if (cudaMalloc(&keys, cap * row_bytes) == cudaSuccess && (same_table || cudaMalloc(&table, nb * 4 * sizeof(u64)) == cudaSuccess)) break;
        cudaGetLastError();
        if (keys) { cudaFree(keys); keys = nullptr; }
        if (cap <= used + need + 1024) return fail(h, FQD_ERR_CAPACITY, "no device memory left to grow the key store");
    }
    CUDA_TRY(h, cudaMemcpyAsync(keys, h->d_keys, used * row_bytes, cudaMemcpyDeviceToDevice, h->stream));
    if (table) {
        CUDA_TRY(h, cudaMemsetAsync(table, 0xFF, nb * 4 * sizeof(u64), h->stream));
        RehashParams rp;
        rp.old_table = h->d_table; rp.old_entries = h->n_buckets * 4; rp.table = table; rp.bucket_shift = shift; rp.bucket_mask = nb - 1;
        rp.keys = keys; rp.row_words = h->row_words; rp.W = h->W; rp.mates = h->cfg.paired ? 2 : 1; rp.hash_mul = 1;
        k_rehash<<<h->sm_count * 8, HS_THREADS, 0, h->stream>>>(rp);
        h->launches++;
    }
    RunState* r = h->d_run;
    CUDA_TRY(h, cudaMemsetAsync(&r->capacity_exceeded, 0, sizeof(u32), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(&r->sticky_set, 0, sizeof(u32), h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->d_surv) {
        u64* sv = nullptr;
        if (cudaMalloc(&sv, cap * sizeof(u64)) != cudaSuccess) { cudaGetLastError(); cudaFree(keys); cudaFree(table); return fail(h, FQD_ERR_CAPACITY, "no device memory left to grow the survivor list"); }
        CUDA_TRY(h, cudaMemcpy(sv, h->d_surv, h->h_run->n_survivors * sizeof(u64), cudaMemcpyDeviceToDevice));
        cudaFree(h->d_surv); h->d_surv = sv;
    }
    cudaFree(h->d_keys); h->d_keys = keys; h->key_capacity = cap;
    if (table) { cudaFree(h->d_table); h->d_table = table; h->n_buckets = nb; h->bucket_shift = shift; }
    h->grows++;
    CUDA_TRY(h, cudaGetLastError());
    return FQD_OK;
}

static int read_back_chunk(fqd_handle* h) {
    const int mates = h->cfg.paired ? 2 : 1;
    for (int m = 0; m < mates; ++m)
        CUDA_TRY(h, cudaMemcpyAsync(h->mate[m].h_ctl, h->mate[m].d_ctl, sizeof(ChunkCtl), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_run, h->d_run, sizeof(RunState), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return drain_events(h);
}

static int enqueue_fast_chunk(fqd_handle* h, const void* d_r1, size_t n1, const void* d_r2, size_t n2);

// d_r1 / d_r2: where the chunk's bytes are on the device (needed to run it again after the set has grown)
static int finish_fast_chunk(fqd_handle* h, const void* d_r1, size_t n1, const void* d_r2, size_t n2, fqd_chunk_result* res) {
    int rc = read_back_chunk(h);
    if (rc) return rc;
    for (int attempt = 0; h->h_run->capacity_exceeded && attempt < 4; ++attempt) {
        if (grow_fast(h, h->h_run->chunk_wanted) != FQD_OK) break;       // reported as FQD_ERR_CAPACITY below
        rc = enqueue_fast_chunk(h, d_r1, n1, d_r2, n2);
        if (!rc) rc = read_back_chunk(h);
        if (rc) return rc;
    }
    return fold_fast_chunk(h, n1, n2, res);
}

static int fold_fast_chunk(fqd_handle* h, size_t n1, size_t n2, fqd_chunk_result* res) {
    const int mates = h->cfg.paired ? 2 : 1;
    const u64 pairs = h->h_run->chunk_pairs;
    const u64 first_record = h->h_run->n_records - pairs;
    for (int m = 0; m < mates; ++m)
        CUDA_TRY(h, cudaMemcpyAsync(h->mate[m].h_rec_start, h->mate[m].d_rec_start, (pairs + 1) * sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_dup, h->d_dup, pairs, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    u64 n_ok = pairs;
    fold_chunk(h, first_record, &n_ok);
    // statistics count only what the reference would have processed before an error
    u64 dups_ok = h->h_run->chunk_dups;
    if (n_ok < pairs) { dups_ok = 0; for (u64 i = 0; i < n_ok; ++i) dups_ok += h->h_dup[i]; }
    h->stats.total += n_ok;
    h->stats.dups += dups_ok;
    if (res) {
        memset(res, 0, sizeof *res);
        res->n_records = n_ok;
        res->first_record = first_record;
        res->n_survivors = n_ok - dups_ok;
        res->dup = h->h_dup;
        const size_t nn[2] = {n1, n2};
        for (int m = 0; m < mates; ++m) {
            res->rec_start[m] = h->mate[m].h_rec_start;
            res->consumed[m] = nn[m] ? h->mate[m].h_rec_start[pairs] : 0;
        }
    }
    return FQD_OK;
}

extern "C" int fqd_push_device(fqd_handle* h, const void* d_r1, size_t n1, const void* d_r2, size_t n2, fqd_chunk_result* res) {
    if (!h) return FQD_ERR_INVALID;
    if (h->cfg.mode != FQD_MODE_FAST || h->cfg.unordered) return fail(h, FQD_ERR_INVALID, "fqd_push* is for ordered --fast mode; use fqd_append/fqd_finish");
    if (h->pending_async) { int rc = fqd_sync(h); if (rc) return rc; }
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = enqueue_fast_chunk(h, d_r1, n1, h->cfg.paired ? d_r2 : nullptr, h->cfg.paired ? n2 : 0);
    if (rc) return rc;
    return finish_fast_chunk(h, d_r1, n1, h->cfg.paired ? d_r2 : nullptr, h->cfg.paired ? n2 : 0, res);
}

extern "C" int fqd_push(fqd_handle* h, const char* r1, size_t n1, const char* r2, size_t n2, fqd_chunk_result* res) {
    if (!h) return FQD_ERR_INVALID;
    if (h->cfg.mode != FQD_MODE_FAST || h->cfg.unordered) return fail(h, FQD_ERR_INVALID, "fqd_push* is for ordered --fast mode; use fqd_append/fqd_finish");
    if (n1 > h->cfg.max_chunk_bytes || n2 > h->cfg.max_chunk_bytes) return fail(h, FQD_ERR_INVALID, "chunk larger than max_chunk_bytes");
    if (h->pending_async) { int rc = fqd_sync(h); if (rc) return rc; }
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (n1) CUDA_TRY(h, cudaMemcpyAsync(h->mate[0].d_raw, r1, n1, cudaMemcpyHostToDevice, h->stream));
    if (h->cfg.paired && n2) CUDA_TRY(h, cudaMemcpyAsync(h->mate[1].d_raw, r2, n2, cudaMemcpyHostToDevice, h->stream));
    int rc = enqueue_fast_chunk(h, h->mate[0].d_raw, n1, h->cfg.paired ? h->mate[1].d_raw : nullptr, h->cfg.paired ? n2 : 0);
    if (rc) return rc;
    return finish_fast_chunk(h, h->mate[0].d_raw, n1, h->cfg.paired ? h->mate[1].d_raw : nullptr, h->cfg.paired ? n2 : 0, res);
}

// Same contract as fqd_push, split in two so that the copy of chunk c+1 overlaps the processing of chunk c:
//   fqd_push_prefetch(c+1) ... fqd_push_staged(c) -> res(c).  At most two chunks are staged at any time; the host
//   buffers must stay valid until the fqd_push_staged call that consumes them returns.
extern "C" int fqd_push_prefetch(fqd_handle* h, const char* r1, size_t n1, const char* r2, size_t n2) {
    if (!h) return FQD_ERR_INVALID;
    if (h->cfg.mode != FQD_MODE_FAST || h->cfg.unordered) return fail(h, FQD_ERR_INVALID, "fqd_push* is for ordered --fast mode; use fqd_append/fqd_finish");
    if (n1 > h->cfg.max_chunk_bytes || n2 > h->cfg.max_chunk_bytes) return fail(h, FQD_ERR_INVALID, "chunk larger than max_chunk_bytes");
    if (h->n_staged >= 2) return fail(h, FQD_ERR_INVALID, "fqd_push_prefetch: two chunks are staged already");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    const int mates = h->cfg.paired ? 2 : 1;
    if (!h->copy_stream) {
        CUDA_TRY(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) CUDA_TRY(h, cudaEventCreateWithFlags(&h->copy_done[k], cudaEventDisableTiming));
        for (int m = 0; m < mates; ++m) CUDA_TRY(h, cudaMalloc(&h->d_raw_alt[m], h->cfg.max_chunk_bytes + 4096));
    }
    const int s = h->pf_slot;
    const char* src[2] = {r1, r2};
    const size_t nn[2] = {n1, h->cfg.paired ? n2 : 0};
    for (int m = 0; m < mates; ++m) {
        u8* dst = s == 0 ? h->mate[m].d_raw : h->d_raw_alt[m];
        if (nn[m]) CUDA_TRY(h, cudaMemcpyAsync(dst, src[m], nn[m], cudaMemcpyHostToDevice, h->copy_stream));
        h->staged_n[s][m] = nn[m];
    }
    CUDA_TRY(h, cudaEventRecord(h->copy_done[s], h->copy_stream));
    h->pf_slot ^= 1; h->n_staged++;
    return FQD_OK;
}
extern "C" int fqd_push_staged(fqd_handle* h, fqd_chunk_result* res) {
    if (!h) return FQD_ERR_INVALID;
    if (h->n_staged == 0) return fail(h, FQD_ERR_INVALID, "fqd_push_staged without fqd_push_prefetch");
    if (h->pending_async) { int rc = fqd_sync(h); if (rc) return rc; }
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    const int s = h->run_slot;
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->copy_done[s], 0));
    const u8* d1 = s == 0 ? h->mate[0].d_raw : h->d_raw_alt[0];
    const u8* d2 = h->cfg.paired ? (s == 0 ? h->mate[1].d_raw : h->d_raw_alt[1]) : nullptr;
    const size_t n1 = h->staged_n[s][0], n2 = h->staged_n[s][1];
    h->run_slot ^= 1; h->n_staged--;
    int rc = enqueue_fast_chunk(h, d1, n1, d2, n2);
    if (rc) return rc;
    return finish_fast_chunk(h, d1, n1, d2, n2, res);
}

extern "C" int fqd_push_device_async(fqd_handle* h, const void* d_r1, size_t n1, const void* d_r2, size_t n2) {
    if (!h) return FQD_ERR_INVALID;
    if (h->cfg.mode != FQD_MODE_FAST || h->cfg.unordered) return fail(h, FQD_ERR_INVALID, "fqd_push* is for ordered --fast mode");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    h->pending_async = true;
    return enqueue_fast_chunk(h, d_r1, n1, h->cfg.paired ? d_r2 : nullptr, h->cfg.paired ? n2 : 0);
}

extern "C" int fqd_sync(fqd_handle* h) {
    if (!h) return FQD_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (h->seq) { CUDA_TRY(h, cudaStreamSynchronize(h->stream)); return drain_events(h); }
    const int mates = h->cfg.paired ? 2 : 1;
    for (int m = 0; m < mates; ++m)
        CUDA_TRY(h, cudaMemcpyAsync(h->mate[m].h_ctl, h->mate[m].d_ctl, sizeof(ChunkCtl), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_run, h->d_run, sizeof(RunState), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    drain_events(h);
    if (h->pending_async) {
        // async pushes: totals come from the device counters; the first chunk that raised a data error left its error
        // words in the run state (k_chunk_end), whichever chunk it was
        h->pending_async = false;
        u64 n_ok;
        h->stats.total = h->h_run->n_records;
        h->stats.dups = h->h_run->n_dups;
        if (h->h_run->sticky_set) {
            const RunState& r = *h->h_run;
            for (int m = 0; m < mates; ++m) {
                ChunkCtl& c = *h->mate[m].h_ctl;
                c.err_parse = r.sticky_parse[m]; c.err_base = r.sticky_base[m]; c.too_long = r.sticky_too_long[m]; c.too_long_rec = r.sticky_too_long_rec[m];
            }
            const u32 keep_pairs = h->h_run->chunk_pairs;
            h->h_run->chunk_pairs = r.sticky_pairs;
            fold_chunk(h, r.sticky_first, &n_ok);
            h->h_run->chunk_pairs = keep_pairs;
        } else {
            fold_chunk(h, h->h_run->n_records - h->h_run->chunk_pairs, &n_ok);
        }
    }
    return FQD_OK;
}

extern "C" int fqd_timer_start(fqd_handle* h) {
    if (!h) return FQD_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (!h->timer0) { CUDA_TRY(h, cudaEventCreate(&h->timer0)); CUDA_TRY(h, cudaEventCreate(&h->timer1)); }
    CUDA_TRY(h, cudaEventRecord(h->timer0, h->stream));
    return FQD_OK;
}
extern "C" int fqd_timer_stop(fqd_handle* h, double* ms) {
    if (!h || !h->timer0) return FQD_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaEventRecord(h->timer1, h->stream));
    CUDA_TRY(h, cudaEventSynchronize(h->timer1));
    float f = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&f, h->timer0, h->timer1));
    if (ms) *ms = f;
    return FQD_OK;
}
extern "C" int fqd_profile_enable(fqd_handle* h, int on) {
    if (!h) return FQD_ERR_INVALID;
    h->profile = on != 0;
    if (on) memset(&h->prof, 0, sizeof h->prof);
    return FQD_OK;
}
extern "C" int fqd_profile_get(fqd_handle* h, fqd_profile_t* out) {
    if (!h || !out) return FQD_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    int rc = drain_events(h);
    if (rc) return rc;
    h->prof.parse_records = h->seq ? 0 : h->prof.parse_records;
    *out = h->prof;
    return FQD_OK;
}

extern "C" int fqd_keep_survivors(fqd_handle* h, int on) {
    if (!h) return FQD_ERR_INVALID;
    if (h->cfg.mode != FQD_MODE_FAST || h->cfg.unordered) return fail(h, FQD_ERR_INVALID, "fqd_keep_survivors is for ordered --fast mode (fqd_emission lists the written records of the other modes)");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (on && !h->d_surv) {
        if (h->stats.total || h->pending_async) return fail(h, FQD_ERR_INVALID, "fqd_keep_survivors: enable it before the first chunk (or after fqd_reset)");
        CUDA_TRY(h, cudaMalloc(&h->d_surv, h->key_capacity * sizeof(u64)));
        CUDA_TRY(h, cudaMalloc(&h->d_surv_cnt, ((size_t)h->cap / SV_BLOCK + 2) * sizeof(u32)));
    } else if (!on && h->d_surv) {
        cudaFree(h->d_surv); cudaFree(h->d_surv_cnt); h->d_surv = nullptr; h->d_surv_cnt = nullptr;
    }
    return FQD_OK;
}
extern "C" int fqd_survivors(fqd_handle* h, uint64_t first, uint64_t* dst, uint64_t cap, uint64_t* n_total, const uint64_t** d_list) {
    if (!h || !h->d_surv) return fail(h, FQD_ERR_INVALID, "fqd_survivors: call fqd_keep_survivors(h, 1) before the first chunk");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (h->pending_async) { int rc = fqd_sync(h); if (rc) return rc; }
    CUDA_TRY(h, cudaMemcpyAsync(h->h_run, h->d_run, sizeof(RunState), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const u64 n = h->h_run->n_survivors;
    if (n_total) *n_total = n;
    if (d_list) *d_list = (const uint64_t*)h->d_surv;
    if (dst && first < n) {
        const u64 k = std::min<u64>(cap, n - first);
        CUDA_TRY(h, cudaMemcpy(dst, h->d_surv + first, k * sizeof(u64), cudaMemcpyDeviceToHost));
    }
    return FQD_OK;
}

extern "C" int fqd_reset(fqd_handle* h) {
    if (!h) return FQD_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (h->seq) return seq_reset(h->seq, &h->err);
    cudaEvent_t e0 = get_event(h), e1 = get_event(h);
    CUDA_TRY(h, cudaEventRecord(e0, h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->d_table, 0xFF, h->n_buckets * 4 * sizeof(u64), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->d_run, 0, sizeof(RunState), h->stream));
    CUDA_TRY(h, cudaEventRecord(e1, h->stream));
    h->pending_events.emplace_back(e0, e1);
    h->launches += 2;
    memset(&h->stats, 0, sizeof h->stats);
    return FQD_OK;
}

// -----------------------------------------------------------------------------------------------------------
// multi-GPU --fast mode
extern "C" int fqd_set_stream(fqd_handle* h, void* cuda_stream) {
    if (!h) return FQD_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
    if (h->seq) h->seq->stream = h->stream;
    if (h->shard_ctx) h->shard_ctx->stream = h->stream;
    return FQD_OK;
}
extern "C" size_t fqd_shard_row_bytes(fqd_handle* h) { return h ? ((size_t)h->row_words + 1) * sizeof(u64) : 0; }

static int shard_init(fqd_handle* h) {
    if (h->shard_ctx) return FQD_OK;
    if (h->cfg.mode != FQD_MODE_FAST || h->cfg.unordered) return fail(h, FQD_ERR_INVALID, "sharded path: ordered --fast only");
    h->shard_ctx = new SeqState();
    h->shard_ctx->stream = h->stream; h->shard_ctx->sm = h->sm_count;
    int rc = sort_scratch_alloc(h->shard_ctx, h->shard_sc, h->cap, &h->err);
    if (rc) return rc;
    CUDA_TRY(h, cudaMalloc(&h->d_stage_keys, (size_t)h->cap * h->row_words * sizeof(u64)));
    CUDA_TRY(h, cudaMalloc(&h->d_stage_run, sizeof(RunState)));
    CUDA_TRY(h, cudaMemsetAsync(h->d_stage_run, 0, sizeof(RunState), h->stream));
    CUDA_TRY(h, cudaMalloc(&h->d_final_hash, (size_t)h->cap * sizeof(u64)));
    CUDA_TRY(h, cudaMalloc(&h->d_counts, 64 * sizeof(u32)));
    h->recv_cap = (u64)h->cap * 2;
    CUDA_TRY(h, cudaMalloc(&h->d_recv_hash, h->recv_cap * sizeof(u64)));
    cudaFree(h->d_dup); cudaFreeHost(h->h_dup);
    CUDA_TRY(h, cudaMalloc(&h->d_dup, h->recv_cap + 64));
    CUDA_TRY(h, cudaHostAlloc(&h->h_dup, h->recv_cap + 64, cudaHostAllocDefault));
    return FQD_OK;
}

// single-end: d_raw2 == nullptr.  paired-end: both chunks must hold the same records (cut at the same record index).
static int shard_pack_impl(fqd_handle* h, const void* d_raw, size_t n, const void* d_raw2, size_t n2, uint32_t n_shards, void* d_send,
                           uint64_t* counts, uint64_t* n_records) {
    if (!h || !counts || n_shards == 0 || n_shards > 32) return FQD_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = shard_init(h);
    if (rc) return rc;
    const int mates = h->cfg.paired ? 2 : 1;
    if ((mates == 2) != (d_raw2 != nullptr)) return fail(h, FQD_ERR_INVALID, "paired handle needs both chunks, single-end handle one");
    if (n > h->cfg.max_chunk_bytes || n2 > h->cfg.max_chunk_bytes) return fail(h, FQD_ERR_INVALID, "chunk larger than max_chunk_bytes");
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (h->profile) { pe0 = get_event(h); pe1 = get_event(h); cudaEventRecord(pe0, h->stream); }
    for (int m = 0; m < mates; ++m) {
        MateChunk& c = h->mate[m];
        const u8* raw = (const u8*)(m ? d_raw2 : d_raw);
        const size_t nb = m ? n2 : n;
        const u32 n_tiles = (u32)((nb + PP_TILE - 1) / PP_TILE);
        k_init_chunk<<<std::max(1u, std::min(n_tiles / 256 + 1, 1024u)), 256, 0, h->stream>>>(c.d_ctl, c.d_tile_state, n_tiles);
        ParseParams p;
        p.raw = raw; p.n = (u32)nb; p.n_tiles = n_tiles; p.tile_state = c.d_tile_state; p.ctl = c.d_ctl; p.run = h->d_stage_run;
        p.rec_start = c.d_rec_start; p.cap = h->cap; p.keys = h->d_stage_keys; p.key_capacity = h->cap;
        p.row_words = h->row_words; p.mate_off = m * h->W; p.W = h->W; p.hash = c.d_hash; p.seq_len = nullptr; p.word0 = nullptr;
        p.strict = 1; p.hash_salt = m * 4096u; p.dup = nullptr; p.bad_rec = nullptr; p.byte_keys = 0; p.skip = 0;
        if (n_tiles) {
            pp_launch(h->cfg.format == FQD_FORMAT_FASTQ, p, h->stream);
            h->launches += 2;
        }
        if (h->profile) { h->prof.parse_launches++; h->prof.parse_bytes += nb; }
    }
    if (h->profile) { cudaEventRecord(pe1, h->stream); h->prof_parse.emplace_back(pe0, pe1); }
    // owner of every key, one stable radix pass groups the records by owner
    MateChunk& c = h->mate[0];
    SortScratch& sc = h->shard_sc;
    const unsigned g = (unsigned)h->sm_count * 8;
    k_shard_owner<<<g, 256, 0, h->stream>>>(c.d_hash, mates == 2 ? h->mate[1].d_hash : nullptr, c.d_ctl, n_shards, sc.keyA, h->d_final_hash, sc.aA);
    for (int m = 0; m < mates; ++m)
        CUDA_TRY(h, cudaMemcpyAsync(h->mate[m].h_ctl, h->mate[m].d_ctl, sizeof(ChunkCtl), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const u64 nrec = c.h_ctl->n_records;
    if (mates == 2 && h->mate[1].h_ctl->n_records != nrec)
        return fail(h, FQD_ERR_INVALID, "sharded paired input: the two chunks hold different numbers of records");
    h->shard_ctx->stream = h->stream;
    rc = radix_sort(h->shard_ctx, sc, nrec, 0, 8, false, &h->err);
    if (rc) return rc;
    k_shard_counts<<<1, 64, 0, h->stream>>>(sc.keyA, c.d_ctl, n_shards, h->d_counts);
    k_shard_gather<<<g, 256, 0, h->stream>>>(h->d_stage_keys, h->row_words, h->d_final_hash, sc.aA, c.d_ctl, (u64*)d_send);
    u32 hc[64];
    CUDA_TRY(h, cudaMemcpyAsync(hc, h->d_counts, (n_shards + 1) * sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (u32 k = 0; k < n_shards; ++k) counts[k] = hc[k + 1] - hc[k];
    if (n_records) *n_records = nrec;
    h->shard_last_n = nrec;
    h->launches += 3 + 2 * mates + h->shard_ctx->launches; h->shard_ctx->launches = 0;
    u64 n_ok;
    h->h_run->chunk_pairs = (u32)nrec; h->h_run->capacity_exceeded = 0;
    fold_chunk(h, h->stats.total, &n_ok);
    CUDA_TRY(h, cudaGetLastError());
    return FQD_OK;
}
extern "C" int fqd_shard_pack(fqd_handle* h, const void* d_raw, size_t n, uint32_t n_shards, void* d_send, uint64_t* counts, uint64_t* n_records) {
    return shard_pack_impl(h, d_raw, n, nullptr, 0, n_shards, d_send, counts, n_records);
}
extern "C" int fqd_shard_pack_pe(fqd_handle* h, const void* d_r1, size_t n1, const void* d_r2, size_t n2, uint32_t n_shards, void* d_send,
                                 uint64_t* counts, uint64_t* n_records) {
    if (!d_r2) return fail(h, FQD_ERR_INVALID, "fqd_shard_pack_pe needs both mates");
    return shard_pack_impl(h, d_r1, n1, d_r2, n2, n_shards, d_send, counts, n_records);
}

extern "C" int fqd_shard_insert(fqd_handle* h, const void* d_recv, uint64_t n_recv, uint32_t n_shards, void* d_flags) {
    if (!h || !h->shard_ctx) return FQD_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (n_recv > h->recv_cap) return fail(h, FQD_ERR_CAPACITY, "more rows received than the shard buffers hold");
    const unsigned g = (unsigned)h->sm_count * 8;
    CUDA_TRY(h, cudaMemsetAsync(d_flags, 0, n_recv, h->stream));
    k_shard_set_pairs<<<1, 1, 0, h->stream>>>(h->d_run, (u32)n_recv, h->key_capacity);
    k_shard_append<<<g, 256, 0, h->stream>>>((const u64*)d_recv, (u32)n_recv, h->row_words, h->d_keys, h->d_run, h->key_capacity, h->d_recv_hash);
    InsertParams ip;
    ip.table = h->d_table; ip.bucket_shift = h->bucket_shift; ip.bucket_mask = h->n_buckets - 1; ip.keys = h->d_keys;
    ip.row_words = h->row_words; ip.key_capacity = h->key_capacity; ip.hash1 = h->d_recv_hash; ip.hash2 = nullptr;
    ip.ctl1 = nullptr; ip.ctl2 = nullptr; ip.run = h->d_run; ip.dup = (u8*)d_flags; ip.hash_mul = n_shards; ip.hash_final = 1;
    cudaEvent_t ie0 = nullptr, ie1 = nullptr;
    if (h->profile) { ie0 = get_event(h); ie1 = get_event(h); cudaEventRecord(ie0, h->stream); }
    insert_launch(ip, g, h->stream);
    if (h->profile) { cudaEventRecord(ie1, h->stream); h->prof_insert.emplace_back(ie0, ie1); h->prof.insert_launches++; }
    k_chunk_end<<<1, 1, 0, h->stream>>>(h->d_run, nullptr, nullptr);
    h->launches += 4;
    CUDA_TRY(h, cudaGetLastError());
    return FQD_OK;
}

extern "C" int fqd_shard_apply(fqd_handle* h, const void* d_flags_back, uint64_t* chunk_dups) {
    if (!h || !h->shard_ctx) return FQD_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    MateChunk& c = h->mate[0];
    const unsigned g = (unsigned)h->sm_count * 4;
    k_shard_flags_back<<<g, 256, 0, h->stream>>>((const u8*)d_flags_back, h->shard_sc.aA, c.d_ctl, h->d_dup);
    // count this chunk's duplicates with the same reduction the single-GPU path uses
    k_shard_set_pairs<<<1, 1, 0, h->stream>>>(h->d_stage_run, (u32)h->shard_last_n, ~0ull);
    k_count_dups<<<h->sm_count * 2, HS_THREADS, 0, h->stream>>>(h->d_dup, h->d_stage_run);
    CUDA_TRY(h, cudaMemcpyAsync(h->h_run, h->d_stage_run, sizeof(RunState), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->launches += 3;
    const u64 d = h->h_run->chunk_dups;
    h->stats.total += h->shard_last_n;
    h->stats.dups += d;
    if (chunk_dups) *chunk_dups = d;
    // K1 of the next chunk must again see slot base 0
    CUDA_TRY(h, cudaMemsetAsync(h->d_stage_run, 0, sizeof(RunState), h->stream));
    return FQD_OK;
}

extern "C" int fqd_shard_read_flags(fqd_handle* h, void* dst, size_t n) {
    if (!h || !h->shard_ctx || n > h->shard_last_n) return FQD_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaMemcpyAsync(dst, h->d_dup, n, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return FQD_OK;
}


// -----------------------------------------------------------------------------------------------------------
// multi-GPU --fast mode, round 2 (shard2.cuh): regions in the owners' key stores, rows written over peer memory by the
// scatter kernel, ordering by interprocess events.  Call sequence per rank (host barriers of the caller marked |):
//   fqd_shard2_init; fqd_shard2_export -> all-gather -> fqd_shard2_import for every rank |
//   pack(0) | for every chunk c: [pack(c+1)] insert(c) | apply(c)        ... fqd_shard2_finish
// Every call only enqueues; a barrier makes sure that an event has been RECORDED (enqueued) by its owner before a peer
// enqueues the wait for it (cudaStreamWaitEvent waits for the most recent record at the time of the call).
struct Shard2Blob { cudaIpcMemHandle_t keys, hash, counts, flags_in; cudaIpcEventHandle_t scatter[2], flags[2]; };

static void shard2_free(fqd_handle* h) {
    Shard2State* s = h->s2;
    if (!s) return;
    cudaSetDevice(h->cfg.device);
    if (s->s_pack) cudaStreamSynchronize(s->s_pack);
    if (s->s_copy) cudaStreamSynchronize(s->s_copy);
    if (s->s_ins) cudaStreamSynchronize(s->s_ins);
    for (int k = 0; k < 2; ++k) { cudaFree(s->d_stage_rows[k]); cudaFree(s->d_stage_hash[k]); if (s->ev_staged[k]) cudaEventDestroy(s->ev_staged[k]); }
    if (s->s_copy) cudaStreamDestroy(s->s_copy);
    for (u32 r = 0; r < s->N; ++r) {
        if (r == s->me || !s->imported[r] || s->linked[r]) continue;
        cudaIpcCloseMemHandle(s->peer_keys[r]); cudaIpcCloseMemHandle(s->peer_hash[r]); cudaIpcCloseMemHandle(s->peer_counts[r]); cudaIpcCloseMemHandle(s->peer_flags_in[r]);
        for (int k = 0; k < 2; ++k) { cudaEventDestroy(s->peer_scatter[r][k]); cudaEventDestroy(s->peer_flags[r][k]); }
    }
    cudaFree(s->d_block_cnt); cudaFree(s->d_block_base); cudaFree(s->d_dest[0]); cudaFree(s->d_dest[1]); cudaFree(s->d_totals);
    cudaFree(s->d_final_hash); cudaFree(s->d_stage_keys); cudaFree(s->d_stage_run); cudaFree(s->d_run2); cudaFree(s->d_hash_regions);
    cudaFree(s->d_counts); cudaFree(s->d_flags); cudaFree(s->d_flags_in); cudaFree(s->d_ndups); cudaFree(s->d_chunk_n);
    if (s->h_run2) cudaFreeHost(s->h_run2);
    if (s->h_totals) cudaFreeHost(s->h_totals);
    for (int k = 0; k < 2; ++k) { if (s->ev_scatter[k]) cudaEventDestroy(s->ev_scatter[k]); if (s->ev_flags[k]) cudaEventDestroy(s->ev_flags[k]); if (s->lv_scatter[k]) cudaEventDestroy(s->lv_scatter[k]); if (s->lv_flags[k]) cudaEventDestroy(s->lv_flags[k]); }
    if (s->t_ins) cudaEventDestroy(s->t_ins);
    if (s->s_pack) cudaStreamDestroy(s->s_pack);
    if (s->s_ins) cudaStreamDestroy(s->s_ins);
    delete s; h->s2 = nullptr;
}

extern "C" size_t fqd_shard2_blob_bytes(void) { return sizeof(Shard2Blob); }

extern "C" int fqd_shard2_init(fqd_handle* h, uint32_t n_shards, uint32_t me, uint32_t region_rows) {
    if (!h || n_shards == 0 || n_shards > S2_MAX || me >= n_shards || region_rows == 0) return FQD_ERR_INVALID;
    if (h->cfg.mode != FQD_MODE_FAST || h->cfg.unordered) return fail(h, FQD_ERR_INVALID, "sharded path: ordered --fast only");
    if (h->s2) return fail(h, FQD_ERR_INVALID, "fqd_shard2_init: already initialised");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    region_rows = (region_rows + 15u) & ~15u;
    if (region_rows >= (1u << 27)) return fail(h, FQD_ERR_INVALID, "fqd_shard2_init: region_rows must be below 2^27");
    Shard2State* s = new Shard2State();
    h->s2 = s;
    s->N = n_shards; s->me = me; s->region_rows = region_rows;
    s->n_blocks_cap = (h->cap + S2_BLOCK - 1) / S2_BLOCK;
    const size_t reg = (size_t)n_shards * region_rows;
    CUDA_TRY(h, cudaStreamCreateWithFlags(&s->s_pack, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&s->s_ins, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&s->s_copy, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        CUDA_TRY(h, cudaMalloc(&s->d_stage_rows[k], reg * h->row_words * sizeof(u64)));
        CUDA_TRY(h, cudaMalloc(&s->d_stage_hash[k], reg * sizeof(u64)));
        CUDA_TRY(h, cudaEventCreateWithFlags(&s->ev_staged[k], cudaEventDisableTiming));
    }
    CUDA_TRY(h, cudaMalloc(&s->d_block_cnt, (size_t)s->n_blocks_cap * S2_MAX * sizeof(u32)));
    CUDA_TRY(h, cudaMalloc(&s->d_block_base, (size_t)s->n_blocks_cap * S2_MAX * sizeof(u32)));
    for (int k = 0; k < 2; ++k) CUDA_TRY(h, cudaMalloc(&s->d_dest[k], (size_t)h->cap * sizeof(u32)));
    CUDA_TRY(h, cudaMalloc(&s->d_totals, (S2_MAX + 1) * sizeof(u32)));
    CUDA_TRY(h, cudaMemset(s->d_totals, 0, (S2_MAX + 1) * sizeof(u32)));
    CUDA_TRY(h, cudaMalloc(&s->d_final_hash, (size_t)h->cap * sizeof(u64)));
    CUDA_TRY(h, cudaMalloc(&s->d_stage_keys, (size_t)h->cap * h->row_words * sizeof(u64)));
    CUDA_TRY(h, cudaMalloc(&s->d_stage_run, sizeof(RunState)));
    CUDA_TRY(h, cudaMemset(s->d_stage_run, 0, sizeof(RunState)));
    CUDA_TRY(h, cudaMalloc(&s->d_run2, sizeof(RunState)));
    CUDA_TRY(h, cudaMemset(s->d_run2, 0, sizeof(RunState)));
    CUDA_TRY(h, cudaHostAlloc(&s->h_run2, sizeof(RunState), cudaHostAllocDefault));
    CUDA_TRY(h, cudaHostAlloc(&s->h_totals, (S2_MAX + 1) * sizeof(u32), cudaHostAllocDefault));
    CUDA_TRY(h, cudaMalloc(&s->d_hash_regions, 2 * reg * sizeof(u64)));
    CUDA_TRY(h, cudaMalloc(&s->d_counts, 2 * S2_MAX * sizeof(u32)));
    CUDA_TRY(h, cudaMemset(s->d_counts, 0, 2 * S2_MAX * sizeof(u32)));
    CUDA_TRY(h, cudaMalloc(&s->d_flags, reg));
    CUDA_TRY(h, cudaMalloc(&s->d_flags_in, 2 * reg));
    CUDA_TRY(h, cudaMalloc(&s->d_chunk_n, 2 * sizeof(u32)));
    CUDA_TRY(h, cudaMalloc(&s->d_ndups, sizeof(unsigned long long)));
    CUDA_TRY(h, cudaMemset(s->d_ndups, 0, sizeof(unsigned long long)));
    for (int k = 0; k < 2; ++k) {
        CUDA_TRY(h, cudaEventCreateWithFlags(&s->ev_scatter[k], cudaEventDisableTiming | cudaEventInterprocess));
        CUDA_TRY(h, cudaEventCreateWithFlags(&s->ev_flags[k], cudaEventDisableTiming | cudaEventInterprocess));
        CUDA_TRY(h, cudaEventCreateWithFlags(&s->lv_scatter[k], cudaEventDisableTiming));
        CUDA_TRY(h, cudaEventCreateWithFlags(&s->lv_flags[k], cudaEventDisableTiming));
    }
    s->peer_keys[me] = h->d_keys; s->peer_hash[me] = s->d_hash_regions; s->peer_counts[me] = s->d_counts; s->peer_flags_in[me] = s->d_flags_in;
    for (int k = 0; k < 2; ++k) { s->peer_scatter[me][k] = s->ev_scatter[k]; s->peer_flags[me][k] = s->ev_flags[k]; }
    s->imported[me] = true;
    CUDA_TRY(h, cudaDeviceSynchronize());
    return FQD_OK;
}

extern "C" int fqd_shard2_export(fqd_handle* h, void* blob) {
    if (!h || !h->s2 || !blob) return FQD_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    Shard2State* s = h->s2;
    Shard2Blob b;
    memset(&b, 0, sizeof b);
    CUDA_TRY(h, cudaIpcGetMemHandle(&b.keys, h->d_keys));
    CUDA_TRY(h, cudaIpcGetMemHandle(&b.hash, s->d_hash_regions));
    CUDA_TRY(h, cudaIpcGetMemHandle(&b.counts, s->d_counts));
    CUDA_TRY(h, cudaIpcGetMemHandle(&b.flags_in, s->d_flags_in));
    for (int k = 0; k < 2; ++k) {
        CUDA_TRY(h, cudaIpcGetEventHandle(&b.scatter[k], s->ev_scatter[k]));
        CUDA_TRY(h, cudaIpcGetEventHandle(&b.flags[k], s->ev_flags[k]));
    }
    memcpy(blob, &b, sizeof b);
    return FQD_OK;
}

extern "C" int fqd_shard2_import(fqd_handle* h, uint32_t rank, const void* blob) {
    if (!h || !h->s2 || !blob || rank >= h->s2->N) return FQD_ERR_INVALID;
    Shard2State* s = h->s2;
    if (rank == s->me) return FQD_OK;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    Shard2Blob b;
    memcpy(&b, blob, sizeof b);
    CUDA_TRY(h, cudaIpcOpenMemHandle((void**)&s->peer_keys[rank], b.keys, cudaIpcMemLazyEnablePeerAccess));
    CUDA_TRY(h, cudaIpcOpenMemHandle((void**)&s->peer_hash[rank], b.hash, cudaIpcMemLazyEnablePeerAccess));
    CUDA_TRY(h, cudaIpcOpenMemHandle((void**)&s->peer_counts[rank], b.counts, cudaIpcMemLazyEnablePeerAccess));
    CUDA_TRY(h, cudaIpcOpenMemHandle((void**)&s->peer_flags_in[rank], b.flags_in, cudaIpcMemLazyEnablePeerAccess));
    for (int k = 0; k < 2; ++k) {
        CUDA_TRY(h, cudaIpcOpenEventHandle(&s->peer_scatter[rank][k], b.scatter[k]));
        CUDA_TRY(h, cudaIpcOpenEventHandle(&s->peer_flags[rank][k], b.flags[k]));
    }
    s->imported[rank] = true;
    return FQD_OK;
}

__global__ void k_shard2_chunk_end(RunState* run, const ChunkCtl* ctl1, const ChunkCtl* ctl2, u32* n_out) {
    u32 n = ctl1->n_records;
    if (ctl2) n = min(n, ctl2->n_records);
    *n_out = n;
    run->chunk_pairs = n; run->chunk_dups = 0;
    if (!run->sticky_set) {
        const ChunkCtl* c[2] = {ctl1, ctl2 ? ctl2 : ctl1};
        bool any = false;
        for (int m = 0; m < (ctl2 ? 2 : 1); ++m) any |= c[m]->err_parse != NO_ERR || c[m]->err_base != NO_ERR || (c[m]->too_long & TL_SEQ);
        if (any) {
            run->sticky_set = 1; run->sticky_pairs = n; run->sticky_first = run->n_records;
            for (int m = 0; m < 2; ++m) {
                run->sticky_parse[m] = c[m]->err_parse; run->sticky_base[m] = c[m]->err_base;
                run->sticky_too_long[m] = c[m]->too_long; run->sticky_too_long_rec[m] = c[m]->too_long_rec;
            }
        }
    }
    run->n_records += n;
}

// split + pack chunk number `chunk` of THIS rank's input and scatter its rows to their owners.  Chunks are numbered
// alike on every rank; chunk c of rank r holds the records that follow chunk c of rank r - 1 in the global input.
extern "C" int fqd_shard2_pack(fqd_handle* h, uint64_t chunk, const void* d_r1, size_t n1, const void* d_r2, size_t n2) {
    if (!h || !h->s2) return FQD_ERR_INVALID;
    Shard2State* s = h->s2;
    const int mates = h->cfg.paired ? 2 : 1;
    if ((mates == 2) != (d_r2 != nullptr)) return fail(h, FQD_ERR_INVALID, "paired handle needs both chunks, single-end handle one");
    if (n1 > h->cfg.max_chunk_bytes || n2 > h->cfg.max_chunk_bytes) return fail(h, FQD_ERR_INVALID, "chunk larger than max_chunk_bytes");
    if (chunk != s->chunks_packed) return fail(h, FQD_ERR_INVALID, "fqd_shard2_pack: chunks must be packed in order");
    for (u32 r = 0; r < s->N; ++r) if (!s->imported[r]) return fail(h, FQD_ERR_INVALID, "fqd_shard2_pack: a peer has not been imported");
    if ((chunk + 1) * s->N * (u64)s->region_rows > h->key_capacity) return fail(h, FQD_ERR_CAPACITY, "fqd_shard2_pack: the key store has no region left for this chunk");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    const int par = (int)(chunk & 1);
    cudaStream_t st = s->s_pack;
    // the regions of this parity are free once every owner has sent back the flags of chunk - 2
    if (chunk >= 2) {
        for (u32 r = 0; r < s->N; ++r) CUDA_TRY(h, cudaStreamWaitEvent(st, s->peer_flags[r][par], 0));
        CUDA_TRY(h, cudaStreamWaitEvent(st, s->ev_scatter[par], 0));     // my staging area of this parity has been copied out
    }
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (h->profile) { pe0 = get_event(h); pe1 = get_event(h); cudaEventRecord(pe0, st); }
    for (int m = 0; m < mates; ++m) {
        MateChunk& c = h->mate[m];
        const u8* raw = (const u8*)(m ? d_r2 : d_r1);
        const size_t nb = m ? n2 : n1;
        const u32 n_tiles = (u32)((nb + PP_TILE - 1) / PP_TILE);
        k_init_chunk<<<std::max(1u, std::min(n_tiles / 256 + 1, 1024u)), 256, 0, st>>>(c.d_ctl, c.d_tile_state, n_tiles);
        ParseParams p;
        p.raw = raw; p.n = (u32)nb; p.n_tiles = n_tiles; p.tile_state = c.d_tile_state; p.ctl = c.d_ctl; p.run = s->d_stage_run;
        p.rec_start = c.d_rec_start; p.cap = h->cap; p.keys = s->d_stage_keys; p.key_capacity = h->cap;
        p.row_words = h->row_words; p.mate_off = m * h->W; p.W = h->W; p.hash = c.d_hash; p.seq_len = nullptr; p.word0 = nullptr;
        p.strict = 1; p.hash_salt = m * 4096u; p.dup = nullptr; p.bad_rec = nullptr; p.byte_keys = 0; p.skip = 0;
        if (n_tiles) { pp_launch(h->cfg.format == FQD_FORMAT_FASTQ, p, st); h->launches += 2; }
        if (h->profile) { h->prof.parse_launches++; h->prof.parse_bytes += nb; }
        h->launches += 2;
    }
    if (h->profile) { cudaEventRecord(pe1, st); h->prof_parse.emplace_back(pe0, pe1); }
    Shard2Src sp;
    memset(&sp, 0, sizeof sp);
    sp.stage_keys = s->d_stage_keys; sp.row_words = h->row_words; sp.hash1 = h->mate[0].d_hash; sp.hash2 = mates == 2 ? h->mate[1].d_hash : nullptr;
    sp.ctl1 = h->mate[0].d_ctl; sp.ctl2 = mates == 2 ? h->mate[1].d_ctl : nullptr; sp.n_shards = s->N; sp.me = s->me;
    sp.region_rows = s->region_rows; sp.chunk = chunk; sp.block_cnt = s->d_block_cnt; sp.block_base = s->d_block_base;
    sp.dest = s->d_dest[par]; sp.final_hash = s->d_final_hash; sp.totals = s->d_totals;
    const size_t reg = (size_t)s->N * s->region_rows;
    sp.stage_rows = s->d_stage_rows[par]; sp.stage_hash = s->d_stage_hash[par];
    for (u32 r = 0; r < s->N; ++r) sp.peer_counts[r] = s->peer_counts[r] + (size_t)par * S2_MAX;
    cudaEvent_t se0 = nullptr, se1 = nullptr;
    if (h->profile) { se0 = get_event(h); se1 = get_event(h); cudaEventRecord(se0, st); }
    k_shard_count2<<<s->n_blocks_cap, 256, 0, st>>>(sp);
    k_shard_bases2<<<1, 256, 0, st>>>(sp, s->n_blocks_cap);
    k_shard_scatter2<<<s->n_blocks_cap, 256, 0, st>>>(sp);
    if (h->profile) { cudaEventRecord(se1, st); h->prof_scatter.emplace_back(se0, se1); h->prof.scatter_launches++; }
    k_shard2_chunk_end<<<1, 1, 0, st>>>(s->d_run2, sp.ctl1, sp.ctl2, s->d_chunk_n + par);
    h->launches += 4;
    // the copy engines take it from here: every owner's part of the staging area goes straight into that owner's key
    // store region (chunk, me) and hash region (me) - fixed sizes, nothing to ask the device; K1 of the next chunk
    // starts on the pack stream meanwhile
    CUDA_TRY(h, cudaEventRecord(s->ev_staged[par], st));
    CUDA_TRY(h, cudaStreamWaitEvent(s->s_copy, s->ev_staged[par], 0));
    const size_t row_bytes = (size_t)h->row_words * sizeof(u64);
    const u64 region0 = (chunk * s->N + s->me) * (u64)s->region_rows;
    for (u32 k = 0; k < s->N; ++k) {
        const u32 o = (s->me + 1 + k) % s->N;                 // peers first, round-robin (spreads the link load), myself last
        CUDA_TRY(h, cudaMemcpyAsync(s->peer_keys[o] + region0 * h->row_words, s->d_stage_rows[par] + (size_t)o * s->region_rows * h->row_words,
                                    (size_t)s->region_rows * row_bytes, cudaMemcpyDefault, s->s_copy));
        CUDA_TRY(h, cudaMemcpyAsync(s->peer_hash[o] + (size_t)par * reg + (size_t)s->me * s->region_rows, s->d_stage_hash[par] + (size_t)o * s->region_rows,
                                    (size_t)s->region_rows * sizeof(u64), cudaMemcpyDefault, s->s_copy));
    }
    CUDA_TRY(h, cudaEventRecord(s->ev_scatter[par], s->s_copy));
    CUDA_TRY(h, cudaEventRecord(s->lv_scatter[par], s->s_copy));
    s->chunks_packed++;
    CUDA_TRY(h, cudaGetLastError());
    return FQD_OK;
}

// owner side of chunk `chunk`: every source's rows are in my key store once their scatter events have fired
extern "C" int fqd_shard2_insert(fqd_handle* h, uint64_t chunk) {
    if (!h || !h->s2) return FQD_ERR_INVALID;
    Shard2State* s = h->s2;
    if (chunk != s->chunks_inserted) return fail(h, FQD_ERR_INVALID, "fqd_shard2_insert: chunks must be inserted in order");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    const int par = (int)(chunk & 1);
    cudaStream_t st = s->s_ins;
    for (u32 r = 0; r < s->N; ++r) CUDA_TRY(h, cudaStreamWaitEvent(st, s->peer_scatter[r][par], 0));
    const size_t reg = (size_t)s->N * s->region_rows;
    CUDA_TRY(h, cudaMemsetAsync(s->d_flags, 0, reg, st));
    Insert2Params ip;
    ip.table = h->d_table; ip.bucket_shift = h->bucket_shift; ip.bucket_mask = h->n_buckets - 1; ip.keys = h->d_keys; ip.row_words = h->row_words;
    ip.hash = s->d_hash_regions + (size_t)par * reg; ip.counts = s->d_counts + (size_t)par * S2_MAX; ip.n_shards = s->N; ip.region_rows = s->region_rows;
    ip.chunk = chunk; ip.flags = s->d_flags; ip.run = s->d_run2;
    cudaEvent_t ie0 = nullptr, ie1 = nullptr;
    if (h->profile) { ie0 = get_event(h); ie1 = get_event(h); cudaEventRecord(ie0, st); }
    const unsigned g = (unsigned)h->sm_count * 8;
    if (h->row_words == 8) k_insert2<8><<<g, HS_THREADS, 0, st>>>(ip);
    else if (h->row_words == 16) k_insert2<16><<<g, HS_THREADS, 0, st>>>(ip);
    else k_insert2<0><<<g, HS_THREADS, 0, st>>>(ip);
    if (h->profile) { cudaEventRecord(ie1, st); h->prof_insert.emplace_back(ie0, ie1); h->prof.insert_launches++; }
    FlagsBack2 fb;
    memset(&fb, 0, sizeof fb);
    fb.flags = s->d_flags; fb.counts = ip.counts; fb.n_shards = s->N; fb.region_rows = s->region_rows; fb.me = s->me;
    for (u32 r = 0; r < s->N; ++r) fb.peer_flags[r] = s->peer_flags_in[r] + (size_t)par * reg;
    k_shard_flags_send2<<<h->sm_count, 256, 0, st>>>(fb);
    h->launches += 3;
    CUDA_TRY(h, cudaEventRecord(s->ev_flags[par], st));
    CUDA_TRY(h, cudaEventRecord(s->lv_flags[par], st));
    s->chunks_inserted++;
    CUDA_TRY(h, cudaGetLastError());
    return FQD_OK;
}

// source side again: the owners' flags for my records of chunk `chunk` -> h->d_dup (record order) + duplicate count
extern "C" int fqd_shard2_apply(fqd_handle* h, uint64_t chunk) {
    if (!h || !h->s2) return FQD_ERR_INVALID;
    Shard2State* s = h->s2;
    if (chunk != s->chunks_applied || chunk >= s->chunks_packed) return fail(h, FQD_ERR_INVALID, "fqd_shard2_apply: chunks must be applied in order, after their pack");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    const int par = (int)(chunk & 1);
    cudaStream_t st = s->s_pack;             // after this chunk's pack, before the pack that reuses its parity
    for (u32 r = 0; r < s->N; ++r) CUDA_TRY(h, cudaStreamWaitEvent(st, s->peer_flags[r][par], 0));
    const size_t reg = (size_t)s->N * s->region_rows;
    // (the control blocks already belong to the next chunk's pack: the record count of this one was put aside)
    k_shard_flags2<<<h->sm_count * 4, 256, 0, st>>>(s->d_flags_in + (size_t)par * reg, s->d_dest[par], s->d_chunk_n + par, s->region_rows, h->d_dup, s->d_ndups);
    h->launches += 1;
    s->chunks_applied++;
    CUDA_TRY(h, cudaGetLastError());
    return FQD_OK;
}

// waits for everything enqueued on this rank; totals of THIS rank's records; data errors / region overflow -> fqd_stats
extern "C" int fqd_shard2_finish(fqd_handle* h, uint64_t* n_records, uint64_t* n_dups) {
    if (!h || !h->s2) return FQD_ERR_INVALID;
    Shard2State* s = h->s2;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_pack));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_copy));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_ins));
    unsigned long long nd = 0;
    CUDA_TRY(h, cudaMemcpy(&nd, s->d_ndups, sizeof nd, cudaMemcpyDeviceToHost));
    CUDA_TRY(h, cudaMemcpy(s->h_run2, s->d_run2, sizeof(RunState), cudaMemcpyDeviceToHost));
    CUDA_TRY(h, cudaMemcpy(s->h_totals, s->d_totals, (S2_MAX + 1) * sizeof(u32), cudaMemcpyDeviceToHost));
    drain_events(h);
    h->stats.total = s->h_run2->n_records;
    h->stats.dups = nd;
    if (s->h_totals[S2_MAX]) { h->stats.err = FQD_ERR_CAPACITY; h->err = "a (chunk, source) region of an owner's key store overflowed: raise region_rows"; }
    if (s->h_run2->sticky_set && !h->stats.err) {
        const RunState& r = *s->h_run2;
        const int mates = h->cfg.paired ? 2 : 1;
        for (int m = 0; m < mates; ++m) {
            ChunkCtl& c = *h->mate[m].h_ctl;
            memset(&c, 0, sizeof c);
            c.err_parse = r.sticky_parse[m]; c.err_base = r.sticky_base[m]; c.too_long = r.sticky_too_long[m]; c.too_long_rec = r.sticky_too_long_rec[m];
        }
        h->h_run->chunk_pairs = r.sticky_pairs; h->h_run->capacity_exceeded = 0;
        u64 n_ok;
        fold_chunk(h, r.sticky_first, &n_ok);
    }
    if (n_records) *n_records = h->stats.total;
    if (n_dups) *n_dups = nd;
    return FQD_OK;
}

// start a new job on the same handles (every rank, between two barriers): empty set, counters cleared
extern "C" int fqd_shard2_reset(fqd_handle* h) {
    if (!h || !h->s2) return FQD_ERR_INVALID;
    Shard2State* s = h->s2;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_pack));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_copy));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_ins));
    CUDA_TRY(h, cudaMemsetAsync(h->d_table, 0xFF, h->n_buckets * 4 * sizeof(u64), s->s_ins));
    CUDA_TRY(h, cudaMemsetAsync(s->d_run2, 0, sizeof(RunState), s->s_ins));
    CUDA_TRY(h, cudaMemsetAsync(s->d_ndups, 0, sizeof(unsigned long long), s->s_ins));
    CUDA_TRY(h, cudaMemsetAsync(s->d_totals, 0, (S2_MAX + 1) * sizeof(u32), s->s_ins));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_ins));
    s->chunks_packed = s->chunks_inserted = s->chunks_applied = 0;
    memset(&h->stats, 0, sizeof h->stats);
    return FQD_OK;
}


// ---- the same protocol inside ONE process that drives several GPUs (the drop-in binary, host/dup_remover.cpp): peers are
// linked by pointer instead of CUDA IPC, ordinary events work across the devices of one process, and because one host
// thread enqueues everything in program order no barrier is needed at all.
extern "C" int fqd_shard2_link(fqd_handle* h, uint32_t rank, fqd_handle* peer) {
    if (!h || !h->s2 || !peer || !peer->s2 || rank >= h->s2->N || peer->s2->me != rank || peer->s2->N != h->s2->N ||
        peer->s2->region_rows != h->s2->region_rows || peer->row_words != h->row_words) return FQD_ERR_INVALID;
    Shard2State* s = h->s2;
    if (rank == s->me) return FQD_OK;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (peer->cfg.device != h->cfg.device) {
        int can = 0;
        CUDA_TRY(h, cudaDeviceCanAccessPeer(&can, h->cfg.device, peer->cfg.device));
        if (!can) return fail(h, FQD_ERR_CUDA, "fqd_shard2_link: the two devices cannot access each other's memory");
        const cudaError_t e = cudaDeviceEnablePeerAccess(peer->cfg.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_TRY(h, e);
        cudaGetLastError();
    }
    Shard2State* q = peer->s2;
    s->peer_keys[rank] = peer->d_keys; s->peer_hash[rank] = q->d_hash_regions; s->peer_counts[rank] = q->d_counts; s->peer_flags_in[rank] = q->d_flags_in;
    for (int k = 0; k < 2; ++k) { s->peer_scatter[rank][k] = q->lv_scatter[k]; s->peer_flags[rank][k] = q->lv_flags[k]; }
    s->imported[rank] = true; s->linked[rank] = true;
    return FQD_OK;
}

// fqd_shard2_pack for a chunk that is still in (pinned) HOST memory: H2D copy inside; waits until the chunk has been split
// (not for the exchange) and tells how many records (pairs) it holds and where the incomplete tail of each mate begins -
// what the reader needs to cut the next chunk (BufferedInput::refresh, src/bufferedinput.hpp:66-74).
extern "C" int fqd_shard2_push_host(fqd_handle* h, uint64_t chunk, const char* r1, size_t n1, const char* r2, size_t n2,
                                    uint64_t* n_records, uint64_t* consumed) {
    if (!h || !h->s2 || !n_records || !consumed) return FQD_ERR_INVALID;
    Shard2State* s = h->s2;
    const int mates = h->cfg.paired ? 2 : 1;
    if (n1 > h->cfg.max_chunk_bytes || n2 > h->cfg.max_chunk_bytes) return fail(h, FQD_ERR_INVALID, "chunk larger than max_chunk_bytes");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (n1) CUDA_TRY(h, cudaMemcpyAsync(h->mate[0].d_raw, r1, n1, cudaMemcpyHostToDevice, s->s_pack));
    if (mates == 2 && n2) CUDA_TRY(h, cudaMemcpyAsync(h->mate[1].d_raw, r2, n2, cudaMemcpyHostToDevice, s->s_pack));
    int rc = fqd_shard2_pack(h, chunk, h->mate[0].d_raw, n1, mates == 2 ? h->mate[1].d_raw : nullptr, mates == 2 ? n2 : 0);
    if (rc) return rc;
    for (int m = 0; m < mates; ++m)
        CUDA_TRY(h, cudaMemcpyAsync(h->mate[m].h_ctl, h->mate[m].d_ctl, sizeof(ChunkCtl), cudaMemcpyDeviceToHost, s->s_pack));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_pack));
    u64 n = h->mate[0].h_ctl->n_records;
    if (mates == 2) n = std::min<u64>(n, h->mate[1].h_ctl->n_records);
    const size_t nn[2] = {n1, n2};
    for (int m = 0; m < mates; ++m) {
        u32 c = 0;
        if (nn[m]) CUDA_TRY(h, cudaMemcpy(&c, h->mate[m].d_rec_start + n, sizeof(u32), cudaMemcpyDeviceToHost));
        consumed[m] = c;
    }
    if (mates == 1) consumed[1] = 0;
    *n_records = n;
    s->last_n = (u32)n;
    return FQD_OK;
}

// After fqd_shard2_insert(chunk) on EVERY handle and fqd_shard2_apply(chunk) on this one: what fqd_push would have
// returned for this chunk (record offsets, duplicate flags, data errors in the reference's order of events).
// first_record = global index of the chunk's first record (pair) in the whole input.
extern "C" int fqd_shard2_result(fqd_handle* h, uint64_t first_record, size_t n1, size_t n2, fqd_chunk_result* res) {
    if (!h || !h->s2) return FQD_ERR_INVALID;
    Shard2State* s = h->s2;
    const int mates = h->cfg.paired ? 2 : 1;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_pack));
    const u64 pairs = s->last_n;
    for (int m = 0; m < mates; ++m)
        CUDA_TRY(h, cudaMemcpyAsync(h->mate[m].h_rec_start, h->mate[m].d_rec_start, (pairs + 1) * sizeof(u32), cudaMemcpyDeviceToHost, s->s_pack));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_dup, h->d_dup, pairs, cudaMemcpyDeviceToHost, s->s_pack));
    CUDA_TRY(h, cudaMemcpyAsync(s->h_totals, s->d_totals, (S2_MAX + 1) * sizeof(u32), cudaMemcpyDeviceToHost, s->s_pack));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_pack));
    drain_events(h);
    h->h_run->chunk_pairs = (u32)pairs; h->h_run->capacity_exceeded = s->h_totals[S2_MAX] ? 1u : 0u;
    h->stats.err = 0;                                    // per-chunk report: the driver stops at the first error
    u64 n_ok = pairs;
    fold_chunk(h, first_record, &n_ok);
    u64 dups_ok = 0;
    for (u64 i = 0; i < n_ok; ++i) dups_ok += h->h_dup[i];
    h->stats.total += n_ok;
    h->stats.dups += dups_ok;
    if (res) {
        memset(res, 0, sizeof *res);
        res->n_records = n_ok;
        res->first_record = first_record;
        res->n_survivors = n_ok - dups_ok;
        res->dup = h->h_dup;
        const size_t nn[2] = {n1, n2};
        for (int m = 0; m < mates; ++m) {
            res->rec_start[m] = h->mate[m].h_rec_start;
            res->consumed[m] = nn[m] ? h->mate[m].h_rec_start[pairs] : 0;
        }
    }
    return FQD_OK;
}

// Stopwatch over BOTH streams of this rank: start records an event on the (idle) pack stream, stop records one at the end
// of each stream, waits, and returns the longer interval - first kernel start to last kernel end on this GPU.
extern "C" int fqd_shard2_timer_start(fqd_handle* h) {
    if (!h || !h->s2) return FQD_ERR_INVALID;
    Shard2State* s = h->s2;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (!h->timer0) { CUDA_TRY(h, cudaEventCreate(&h->timer0)); CUDA_TRY(h, cudaEventCreate(&h->timer1)); }
    if (!s->t_ins) CUDA_TRY(h, cudaEventCreate(&s->t_ins));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_ins));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_copy));
    CUDA_TRY(h, cudaStreamSynchronize(s->s_pack));
    CUDA_TRY(h, cudaEventRecord(h->timer0, s->s_pack));
    return FQD_OK;
}
extern "C" int fqd_shard2_timer_stop(fqd_handle* h, double* ms) {
    if (!h || !h->s2 || !h->timer0 || !h->s2->t_ins) return FQD_ERR_INVALID;
    Shard2State* s = h->s2;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaEventRecord(h->timer1, s->s_pack));
    CUDA_TRY(h, cudaEventRecord(s->t_ins, s->s_ins));
    CUDA_TRY(h, cudaEventSynchronize(h->timer1));
    CUDA_TRY(h, cudaEventSynchronize(s->t_ins));
    float a = 0.f, b = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&a, h->timer0, h->timer1));
    CUDA_TRY(h, cudaEventElapsedTime(&b, h->timer0, s->t_ins));
    if (ms) *ms = std::max(a, b);
    return FQD_OK;
}

// duplicate flags (record order) of the chunk last applied on this rank
extern "C" int fqd_shard2_read_flags(fqd_handle* h, void* dst, size_t n) {
    if (!h || !h->s2) return FQD_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->s2->s_pack));
    CUDA_TRY(h, cudaMemcpy(dst, h->d_dup, n, cudaMemcpyDeviceToHost));
    return FQD_OK;
}

// -----------------------------------------------------------------------------------------------------------
// Host barrier between the rank processes of one box (POSIX shared memory + two atomics, a few microseconds): what the
// sharded loop needs between "everybody has enqueued" and "now enqueue the waits" - a gloo / NCCL barrier costs 0.2 - 1 ms
// per chunk with 8 ranks, which is a quarter of a chunk's GPU time.
struct HostBar { std::atomic<uint32_t> count; std::atomic<uint32_t> gen; };
struct HostBarHandle { HostBar* bar; uint32_t n; };
extern "C" int fqd_hostbar_open(const char* name, uint32_t n_ranks, void** out) {
    if (!name || !out || n_ranks == 0) return FQD_ERR_INVALID;
    const int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
    if (fd < 0) return FQD_ERR_INVALID;
    if (ftruncate(fd, sizeof(HostBar)) != 0) { close(fd); return FQD_ERR_INVALID; }       // new objects are zero-filled
    void* p = mmap(nullptr, sizeof(HostBar), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return FQD_ERR_INVALID;
    HostBarHandle* hb = new HostBarHandle{(HostBar*)p, n_ranks};
    *out = hb;
    return FQD_OK;
}
extern "C" int fqd_hostbar_wait(void* handle) {
    HostBarHandle* hb = (HostBarHandle*)handle;
    if (!hb) return FQD_ERR_INVALID;
    const uint32_t g = hb->bar->gen.load(std::memory_order_acquire);
    if (hb->bar->count.fetch_add(1, std::memory_order_acq_rel) + 1 == hb->n) {
        hb->bar->count.store(0, std::memory_order_relaxed);
        hb->bar->gen.store(g + 1, std::memory_order_release);
    } else {
        for (uint32_t spins = 0; hb->bar->gen.load(std::memory_order_acquire) == g; ++spins)
            if (spins > 2000) sched_yield();
    }
    return FQD_OK;
}
extern "C" int fqd_hostbar_close(void* handle, const char* unlink_name) {
    HostBarHandle* hb = (HostBarHandle*)handle;
    if (hb) { munmap(hb->bar, sizeof(HostBar)); delete hb; }
    if (unlink_name) shm_unlink(unlink_name);
    return FQD_OK;
}

extern "C" int fqd_stats(fqd_handle* h, fqd_stats_t* out) {
    if (!h || !out) return FQD_ERR_INVALID;
    if (h->seq) seq_stats(h->seq, &h->stats);
    *out = h->stats;
    return FQD_OK;
}

extern "C" int fqd_device_time_ms(fqd_handle* h, double* ms, uint64_t* launches) {
    if (!h) return FQD_ERR_INVALID;
    if (h->seq) { h->device_ms = seq_device_ms(h->seq); h->launches = seq_launches(h->seq); }
    if (ms) *ms = h->device_ms;
    if (launches) *launches = h->launches;
    return FQD_OK;
}

// -----------------------------------------------------------------------------------------------------------
// whole-input modes
extern "C" int fqd_append(fqd_handle* h, int mate, const char* buf, size_t n) {
    if (!h || !h->seq) return fail(h, FQD_ERR_INVALID, "fqd_append is for sequence / unordered modes");
    cudaSetDevice(h->cfg.device);
    return seq_append(h->seq, mate, buf, n, false, &h->err);
}
extern "C" int fqd_append_device(fqd_handle* h, int mate, const void* d_buf, size_t n) {
    if (!h || !h->seq) return fail(h, FQD_ERR_INVALID, "fqd_append_device is for sequence / unordered modes");
    cudaSetDevice(h->cfg.device);
    return seq_append(h->seq, mate, d_buf, n, true, &h->err);
}
extern "C" int fqd_adopt_device(fqd_handle* h, int mate, const void* d_buf, size_t n) {
    if (!h || !h->seq) return fail(h, FQD_ERR_INVALID, "fqd_adopt_device is for sequence / unordered modes");
    cudaSetDevice(h->cfg.device);
    int rc = seq_adopt(h->seq, mate, d_buf, n, &h->err);
    seq_stats(h->seq, &h->stats);
    return rc;
}
extern "C" int fqd_finish(fqd_handle* h) {
    if (!h || !h->seq) return fail(h, FQD_ERR_INVALID, "fqd_finish is for sequence / unordered modes");
    cudaSetDevice(h->cfg.device);
    int rc = seq_finish(h->seq, &h->err);
    seq_stats(h->seq, &h->stats);
    return rc;
}
extern "C" int fqd_finish_scan(fqd_handle* h) {
    if (!h || !h->seq) return fail(h, FQD_ERR_INVALID, "fqd_finish_scan is for sequence-based modes");
    cudaSetDevice(h->cfg.device);
    int rc = seq_finish_scan(h->seq, &h->err);
    seq_stats(h->seq, &h->stats);
    return rc;
}
extern "C" int fqd_finish_emit(fqd_handle* h) {
    if (!h || !h->seq) return fail(h, FQD_ERR_INVALID, "fqd_finish_emit is for sequence-based modes");
    cudaSetDevice(h->cfg.device);
    int rc = seq_finish_emit(h->seq, &h->err);
    seq_stats(h->seq, &h->stats);
    return rc;
}
extern "C" size_t fqd_boundary_bytes(fqd_handle* h) { return h && h->seq ? boundary_words(h->seq->row_words) * sizeof(u64) : 0; }
extern "C" int fqd_boundary_get(fqd_handle* h, void* state) {
    if (!h || !h->seq || !state) return fail(h, FQD_ERR_INVALID, "fqd_boundary_get is for sequence-based modes");
    cudaSetDevice(h->cfg.device);
    return seq_boundary_get(h->seq, (u64*)state, &h->err);
}
extern "C" int fqd_boundary_fix(fqd_handle* h, const void* prev_state) {
    if (!h || !h->seq || !prev_state) return fail(h, FQD_ERR_INVALID, "fqd_boundary_fix is for sequence-based modes");
    cudaSetDevice(h->cfg.device);
    return seq_boundary_fix(h->seq, (const u64*)prev_state, &h->err);
}
extern "C" int fqd_partition_sample(fqd_handle* h, uint32_t n_samples, uint64_t* samples, uint64_t* n_records) {
    if (!h || !h->seq || !samples || !n_records) return fail(h, FQD_ERR_INVALID, "fqd_partition_sample is for sequence-based modes");
    cudaSetDevice(h->cfg.device);
    int rc = seq_partition_sample(h->seq, n_samples, (u64*)samples, (u64*)n_records, &h->err);
    seq_stats(h->seq, &h->stats);
    return rc;
}
extern "C" int fqd_partition_plan(fqd_handle* h, const uint64_t* splitters, uint32_t n_ranges, uint64_t* counts, uint64_t* bytes) {
    if (!h || !h->seq || !counts || !bytes || (n_ranges > 1 && !splitters)) return fail(h, FQD_ERR_INVALID, "fqd_partition_plan: bad arguments");
    cudaSetDevice(h->cfg.device);
    return seq_partition_plan(h->seq, (const u64*)splitters, n_ranges, (u64*)counts, (u64*)bytes, &h->err);
}
extern "C" int fqd_partition_gather(fqd_handle* h, int mate, void* d_out) {
    if (!h || !h->seq) return fail(h, FQD_ERR_INVALID, "fqd_partition_gather is for sequence-based modes");
    cudaSetDevice(h->cfg.device);
    return seq_partition_gather(h->seq, mate, d_out, &h->err);
}
static bool un_handle(fqd_handle* h) { return h && h->seq && h->cfg.unordered; }
extern "C" int fqd_unordered_prepare(fqd_handle* h, uint64_t* n_left, uint64_t* n_right) {
    if (!un_handle(h) || !n_left || !n_right) return fail(h, FQD_ERR_INVALID, "fqd_unordered_prepare is for --unordered handles");
    cudaSetDevice(h->cfg.device);
    int rc = seq_unordered_prepare(h->seq, (u64*)n_left, (u64*)n_right, &h->err);
    seq_stats(h->seq, &h->stats);
    return rc;
}
extern "C" int fqd_unordered_enter(fqd_handle* h, int side, uint64_t i, uint64_t* pos) {
    if (!un_handle(h) || !pos || side < 0 || side > 1) return fail(h, FQD_ERR_INVALID, "fqd_unordered_enter: bad arguments");
    cudaSetDevice(h->cfg.device);
    return seq_unordered_enter(h->seq, side, i, (u64*)pos, &h->err);
}
extern "C" int fqd_unordered_join(fqd_handle* h, uint64_t limit_i, uint64_t limit_j, uint64_t final_i, uint64_t final_j, uint64_t out[4]) {
    if (!un_handle(h) || !out) return fail(h, FQD_ERR_INVALID, "fqd_unordered_join is for --unordered handles");
    cudaSetDevice(h->cfg.device);
    return seq_unordered_join(h->seq, limit_i, limit_j, final_i, final_j, (u64*)out, &h->err);
}
extern "C" size_t fqd_unordered_row_bytes(fqd_handle* h) { return un_handle(h) ? ((size_t)2 * h->seq->W + 1) * sizeof(u64) : 0; }
extern "C" int fqd_unordered_rows(fqd_handle* h, uint64_t limit, uint32_t n_shards, void* d_send, uint64_t* counts) {
    if (!un_handle(h) || !counts || (!d_send && limit)) return fail(h, FQD_ERR_INVALID, "fqd_unordered_rows: bad arguments");
    cudaSetDevice(h->cfg.device);
    return seq_unordered_rows(h->seq, limit, n_shards, d_send, (u64*)counts, &h->err);
}
extern "C" int fqd_unordered_insert(fqd_handle* h, const void* d_recv, uint64_t n_recv, uint32_t n_shards, void* d_flags) {
    if (!un_handle(h) || (n_recv && (!d_recv || !d_flags)) || n_shards == 0) return fail(h, FQD_ERR_INVALID, "fqd_unordered_insert: bad arguments");
    cudaSetDevice(h->cfg.device);
    return seq_unordered_insert(h->seq, d_recv, n_recv, n_shards, d_flags, &h->err);
}
extern "C" int fqd_unordered_apply(fqd_handle* h, const void* d_flags_back, uint64_t limit, int report_bad) {
    if (!un_handle(h)) return fail(h, FQD_ERR_INVALID, "fqd_unordered_apply is for --unordered handles");
    cudaSetDevice(h->cfg.device);
    int rc = seq_unordered_apply(h->seq, d_flags_back, limit, report_bad, &h->err);
    seq_stats(h->seq, &h->stats);
    return rc;
}
extern "C" int fqd_emit(fqd_handle* h, int mate, void* dst, size_t cap, size_t* n_bytes, int* done) {
    if (!h || !h->seq || !dst || !n_bytes || !done) return fail(h, FQD_ERR_INVALID, "fqd_emit is for sequence / unordered modes");
    cudaSetDevice(h->cfg.device);
    return seq_emit(h->seq, mate, dst, cap, n_bytes, done, &h->err);
}
extern "C" int fqd_emit_clusters(fqd_handle* h, int mate, void* dst, size_t cap, size_t* n_bytes, int* done) {
    if (!h || !h->seq || !dst || !n_bytes || !done) return fail(h, FQD_ERR_INVALID, "fqd_emit_clusters is for sequence-based modes");
    cudaSetDevice(h->cfg.device);
    return seq_emit_clusters(h->seq, mate, dst, cap, n_bytes, done, &h->err);
}
extern "C" int fqd_discard_input(fqd_handle* h, int on) {
    if (!h || !h->seq) return fail(h, FQD_ERR_INVALID, "fqd_discard_input is for sequence / unordered modes");
    for (u32 m = 0; m < h->seq->mates; ++m)
        if (!h->seq->mate[m].segs.empty()) return fail(h, FQD_ERR_INVALID, "fqd_discard_input after the first fqd_append");
    h->seq->discard = on != 0;
    return FQD_OK;
}
extern "C" int fqd_emission_count(fqd_handle* h, uint64_t* n_written, uint64_t* n_sorted) {
    if (!h || !h->seq) return fail(h, FQD_ERR_INVALID, "fqd_emission_count is for sequence / unordered modes");
    if (!h->seq->finished) return fail(h, FQD_ERR_INVALID, "fqd_emission_count before fqd_finish");
    const SeqState* s = h->seq;
    if (n_written) *n_written = s->n_out;
    if (n_sorted) *n_sorted = (s->cfg.unordered || s->cfg.mode == FQD_MODE_FAST || s->stats.err || !s->d_perm || !s->d_keep) ? 0 : s->n;
    return FQD_OK;
}
extern "C" int fqd_emission_read(fqd_handle* h, int mate, uint64_t first, uint64_t count, uint64_t* off, uint32_t* len) {
    if (!h || !h->seq || (count && (!off || !len))) return fail(h, FQD_ERR_INVALID, "fqd_emission_read: bad arguments");
    cudaSetDevice(h->cfg.device);
    return seq_emission_read(h->seq, mate, first, count, (u64*)off, len, &h->err);
}
extern "C" int fqd_cluster_read(fqd_handle* h, int mate, uint64_t first, uint64_t count, uint64_t* off, uint32_t* len, uint8_t* head) {
    if (!h || !h->seq || (count && (!off || !len || !head))) return fail(h, FQD_ERR_INVALID, "fqd_cluster_read: bad arguments");
    cudaSetDevice(h->cfg.device);
    return seq_cluster_read(h->seq, mate, first, count, (u64*)off, len, head, &h->err);
}
extern "C" int fqd_device_memory(int device, size_t* free_bytes, size_t* total_bytes) {
    size_t f = 0, t = 0;
    if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&f, &t) != cudaSuccess) { cudaGetLastError(); return FQD_ERR_CUDA; }
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return FQD_OK;
}
extern "C" int fqd_emission(fqd_handle* h, fqd_emission_t* out) {
    if (!h || !h->seq || !out) return fail(h, FQD_ERR_INVALID, "fqd_emission is for sequence / unordered modes");
    return seq_emission(h->seq, out, &h->err);
}

#ifdef FQD_K1_TIMELINE
// experiment builds only (csrc/Makefile `variants`): per-phase clock64 stamps of every TL_STRIDE-th tile of the last K1 launch
extern "C" int fqd_debug_k1_timeline(long long* dst, size_t n_values) {
    const size_t n = std::min(n_values, (size_t)TL_CAP * TL_SLOTS);
    if (cudaDeviceSynchronize() != cudaSuccess) return FQD_ERR_CUDA;
    return cudaMemcpyFromSymbol(dst, g_k1_timeline, n * sizeof(long long)) == cudaSuccess ? FQD_OK : FQD_ERR_CUDA;
}
#endif

// -----------------------------------------------------------------------------------------------------------
extern "C" size_t fqd_synth_record_bytes(uint32_t read_len) { return 22u + 2u * (size_t)read_len; }

extern "C" int fqd_synth_fastq(int device, void* d_out, uint64_t first, uint64_t count, uint32_t read_len, int mate,
                               uint64_t seed, uint32_t dup_permille, uint32_t n_permille, int variant) {
    if (cudaSetDevice(device) != cudaSuccess) return FQD_ERR_CUDA;
    if (count == 0) return FQD_OK;
    SynthParams p;
    p.out = (u8*)d_out; p.first = first; p.count = count; p.read_len = read_len; p.mate = mate; p.seed = seed;
    p.dup_permille = dup_permille; p.n_permille = n_permille; p.variant = variant;
    u64 blocks = std::min<u64>((count + 7) / 8, 148ull * 64);
    k_synth_fastq<<<(unsigned)blocks, 256>>>(p);
    if (cudaGetLastError() != cudaSuccess) return FQD_ERR_CUDA;
    return cudaDeviceSynchronize() == cudaSuccess ? FQD_OK : FQD_ERR_CUDA;
}
