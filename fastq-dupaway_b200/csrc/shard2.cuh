// shard2.cuh - multi-GPU --fast mode, round 2: hash-range ownership with NO host round trip and NO staging copy per chunk.
//
// Round 1 grouped a chunk's key rows by owner with a radix pass, gathered them into a send buffer, moved them with the
// copy engines after the ranks had all-gathered their sizes on the host, and appended what arrived to the key store:
// three extra kernels (+28 % kernel time) and five host synchronisations per chunk - 0.63 of linear at 8 GPUs.  Here
//   * the key store of an owner is a sequence of fixed REGIONS, one per (chunk, source rank), so the slot of a row
//         slot = ((chunk * N + source) * region_rows) + position of the row among the source's rows for this owner
//     is known to the SOURCE before anything is exchanged, and slot order is still global input order (chunks are dealt
//     round-robin in input order, sources in rank order, positions in record order) - K2's "smallest slot wins" rule
//     keeps exactly the record the reference keeps (src/hash_dup_remover.hpp:133-139);
//   * k_shard_count2 / k_shard_bases2 / k_shard_scatter2 partition a chunk's rows stably by owner in one pass over the
//     rows (per-block counts -> exclusive bases -> ranks by warp match) into a staging area that is laid out like the
//     owners' regions; the COPY ENGINES then move every owner's part STRAIGHT INTO THAT OWNER'S KEY STORE and hash
//     regions over mapped peer memory (fixed-size copies of region_rows rows: no size has to cross the host), while the
//     SMs already split the next chunk.  (The first version of this file scattered with 16-byte stores from the SMs:
//     6.3 GB per GPU and job over NVLink on the pack stream, K1 of the next chunk behind it - 0.69 of linear at 8 GPUs.)
//     The per-source counts go into the owner's header with plain peer stores;
//   * the owner inserts region by region (k_insert2: the same probe as K2), and writes one flag byte per row back into
//     the SOURCE's flag regions, again over peer memory; k_shard_flags2 puts them at the source's records.
// Ordering between ranks is carried by interprocess CUDA events (fqd_shard2_* in fqd_api.cu); the host only enqueues.
#pragma once
#include "common.cuh"
#include "hashset.cuh"

namespace fqd {

constexpr u32 S2_BLOCK = 4096;          // records per block of the partition kernels
constexpr u32 S2_MAX = 16;              // ranks

struct Shard2Src {
    const u64* stage_keys;              // [cap * row_words] rows of this chunk as K1 packed them
    u32 row_words;
    const u64* hash1; const u64* hash2; // raw per-mate hashes of K1
    const ChunkCtl* ctl1; const ChunkCtl* ctl2;
    u32 n_shards, me;
    u32 region_rows;                    // capacity of one (chunk, source) region
    u64 chunk;                          // chunk number (same on every rank)
    u32* block_cnt;                     // [n_blocks * S2_MAX]
    u32* block_base;                    // [n_blocks * S2_MAX]
    u32* dest;                          // [cap] owner << 27 | position, for the way back
    u64* final_hash;                    // [cap]
    u32* totals;                        // [S2_MAX + 1]: rows per owner of this chunk, [S2_MAX] = overflow flag
    u64* stage_rows;                    // [n_shards * region_rows * row_words] rows grouped by owner (this chunk parity)
    u64* stage_hash;                    // [n_shards * region_rows]
    // per owner (mapped peer memory; [me] = my own)
    u32* peer_counts[S2_MAX];           // header of chunk parity: [n_shards]
};

__device__ __forceinline__ u32 s2_records(const Shard2Src& p) {
    u32 n = p.ctl1->n_records;
    if (p.ctl2) n = min(n, p.ctl2->n_records);
    return n;
}

// owner + finalised hash of every record; rows per (block, owner)
__global__ void __launch_bounds__(256) k_shard_count2(const Shard2Src p) {
    __shared__ u32 s_cnt[S2_MAX];
    if (threadIdx.x < S2_MAX) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const u32 n = s2_records(p);
    const u32 lo = blockIdx.x * S2_BLOCK, hi = min(n, lo + S2_BLOCK);
    for (u32 i0 = lo + (threadIdx.x & ~31u); i0 < hi; i0 += 256) {           // whole warps: the match below is warp-wide
        const u32 i = i0 + (threadIdx.x & 31u);
        const bool in = i < hi;
        u32 o = 0xFFu;
        if (in) {
            const u64 h = p.hash2 ? pair_hash(p.hash1[i], p.hash2[i]) : mix64(p.hash1[i]);     // pair key = both mates
            p.final_hash[i] = h;
            o = (u32)__umul64hi(h, (u64)p.n_shards);
        }
        // warp-aggregated: one shared-memory atomic per distinct owner per warp
        const u32 peers = __match_any_sync(0xFFFFFFFFu, o);
        if (in && (threadIdx.x & 31u) == (u32)__ffs((int)peers) - 1u) atomicAdd(&s_cnt[o], (u32)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < S2_MAX) p.block_cnt[blockIdx.x * S2_MAX + threadIdx.x] = s_cnt[threadIdx.x];
}

// exclusive prefix over the blocks, per owner (one block; a few thousand values); totals -> the owners' headers
__global__ void __launch_bounds__(256) k_shard_bases2(const Shard2Src p, u32 n_blocks_cap) {
    __shared__ u32 s_part[256];
    const u32 n = s2_records(p);
    const u32 n_blocks = min(n_blocks_cap, (n + S2_BLOCK - 1) / S2_BLOCK);
    for (u32 o = 0; o < p.n_shards; ++o) {
        // thread t owns blocks [t * per, (t + 1) * per)
        const u32 per = (n_blocks + 255u) / 256u;
        u32 sum = 0;
        for (u32 b = threadIdx.x * per; b < min(n_blocks, (threadIdx.x + 1) * per); ++b) sum += p.block_cnt[b * S2_MAX + o];
        s_part[threadIdx.x] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 run = 0;
            for (u32 t = 0; t < 256; ++t) { const u32 v = s_part[t]; s_part[t] = run; run += v; }
            p.totals[o] = run;
            if (run > p.region_rows) p.totals[S2_MAX] = 1;                 // a region overflows: the caller fails the job
            // the owner's header for this chunk: how many rows source `me` sent (clamped: nothing is written past a region)
            p.peer_counts[o][p.me] = min(run, p.region_rows);
        }
        __syncthreads();
        u32 run = s_part[threadIdx.x];
        for (u32 b = threadIdx.x * per; b < min(n_blocks, (threadIdx.x + 1) * per); ++b) {
            p.block_base[b * S2_MAX + o] = run;
            run += p.block_cnt[b * S2_MAX + o];
        }
        __syncthreads();
    }
}

// stable scatter: row i -> owner o, position base(block, o) + rank of i among the block's records of owner o
__global__ void __launch_bounds__(256) k_shard_scatter2(const Shard2Src p) {
    __shared__ u32 s_warp[8][S2_MAX];           // rows per (warp, owner) of this block, then exclusive bases
    __shared__ u32 s_pos[S2_BLOCK];              // owner << 27 | position of the block's records
    const u32 n = s2_records(p);
    const u32 lo = blockIdx.x * S2_BLOCK, hi = min(n, lo + S2_BLOCK);
    if (lo >= n) return;
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (u32 k = threadIdx.x; k < 8 * S2_MAX; k += 256) (&s_warp[0][0])[k] = 0;
    __syncthreads();
    // warp w takes the block's records [w * 512, (w + 1) * 512), 32 at a time, in order
    constexpr u32 PERW = S2_BLOCK / 8;
    const u32 wlo = lo + warp * PERW, whi = min(hi, wlo + PERW);
    for (u32 i0 = wlo; i0 < whi; i0 += 32) {
        const u32 i = i0 + lane;
        const bool in = i < whi;
        const u32 o = in ? (u32)__umul64hi(p.final_hash[i], (u64)p.n_shards) : 0xFFu;
        const u32 peers = __match_any_sync(0xFFFFFFFFu, o);
        if (in && lane == (u32)__ffs((int)peers) - 1u) s_warp[warp][o] += (u32)__popc(peers);
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < p.n_shards) {
        u32 run = p.block_base[blockIdx.x * S2_MAX + threadIdx.x];
        for (u32 w = 0; w < 8; ++w) { const u32 v = s_warp[w][threadIdx.x]; s_warp[w][threadIdx.x] = run; run += v; }
    }
    __syncthreads();
    for (u32 i0 = wlo; i0 < whi; i0 += 32) {
        const u32 i = i0 + lane;
        const bool in = i < whi;
        const u32 o = in ? (u32)__umul64hi(p.final_hash[i], (u64)p.n_shards) : 0xFFu;
        const u32 peers = __match_any_sync(0xFFFFFFFFu, o);
        const u32 rank = (u32)__popc(peers & ((1u << lane) - 1u));
        u32 pos = 0;
        if (in) pos = s_warp[warp][o] + rank;
        __syncwarp();
        if (in && lane == (u32)__ffs((int)peers) - 1u) s_warp[warp][o] += (u32)__popc(peers);
        __syncwarp();
        if (in) {
            const u32 d = (o << 27) | pos;
            s_pos[i - lo] = d;
            p.dest[i] = d;
            if (pos < p.region_rows) p.stage_hash[(u64)o * p.region_rows + pos] = p.final_hash[i];
        }
    }
    __syncthreads();
    // rows: 16 bytes per thread, the lanes of a row next to each other
    const u32 q = p.row_words / 2;                                     // 16-byte pieces per row
    const u32 pieces = (hi - lo) * q;
    for (u32 t = threadIdx.x; t < pieces; t += 256) {
        const u32 r = t / q, w = t % q;
        const u32 d = s_pos[r];
        const u32 o = d >> 27, pos = d & 0x7FFFFFFu;
        if (pos >= p.region_rows) continue;
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p.stage_keys + (u64)(lo + r) * p.row_words + 2u * w);
        *reinterpret_cast<ulonglong2*>(p.stage_rows + ((u64)o * p.region_rows + pos) * p.row_words + 2u * w) = v;
    }
}

// ---- owner side: insert the rows of one chunk, region by region (source rank order = global input order)
struct Insert2Params {
    u64* table; u32 bucket_shift; u64 bucket_mask;
    const u64* keys; u32 row_words;
    const u64* hash;                    // hash regions of this chunk parity
    const u32* counts;                  // header: rows per source
    u32 n_shards, region_rows;
    u64 chunk;
    u8* flags;                          // [n_shards * region_rows] one byte per row, zeroed before
    RunState* run;                      // n_records counts rows, n_dups duplicates (statistics)
};

template <int RW>
__global__ void __launch_bounds__(HS_THREADS, 4) k_insert2(const Insert2Params p) {
    __shared__ u32 s_start[S2_MAX + 1];
    if (threadIdx.x == 0) {
        u32 run = 0;
        for (u32 s = 0; s < p.n_shards; ++s) { s_start[s] = run; run += min(p.counts[s], p.region_rows); }
        s_start[p.n_shards] = run;
    }
    __syncthreads();
    const u32 n = s_start[p.n_shards];
    const u64 chunk_base = p.chunk * p.n_shards * (u64)p.region_rows;
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u32 s = 0;
        while (i >= s_start[s + 1]) ++s;
        const u32 local = s * p.region_rows + (i - s_start[s]);
        const u64 h = p.hash[local] * (u64)p.n_shards;      // my share of the hash range, spread over my own table
        const u64 slot = chunk_base + local;
        const u64 tag = (h >> 8) & 0xFFFFFFull;
        const u64 mine = (tag << 40) | slot;
        const u64* myrow = p.keys + slot * p.row_words;
        u64 b = h >> p.bucket_shift;
        bool done = false;
        while (!done) {
            u64* bp = p.table + b * 4;
            ulonglong2 e01 = __ldcg(reinterpret_cast<const ulonglong2*>(bp));
            ulonglong2 e23 = __ldcg(reinterpret_cast<const ulonglong2*>(bp) + 1);
            u64 e[4] = {e01.x, e01.y, e23.x, e23.y};
#pragma unroll
            for (int k = 0; k < 4 && !done; ++k) {
                u64 cur = e[k];
                if (cur == HS_EMPTY) {
                    u64 old = atomicCAS(bp + k, HS_EMPTY, mine);
                    if (old == HS_EMPTY) { done = true; break; }
                    cur = old;
                }
                if ((cur >> 40) == tag) {
                    const u64 other = cur & HS_SLOT_MASK;
                    if (rows_equal<RW>(myrow, p.keys + other * p.row_words, p.row_words)) {
                        u64 old = atomicMin(bp + k, mine);
                        if (old < mine) p.flags[local] = 1;                                   // an earlier record holds this key
                        else p.flags[(u32)((old & HS_SLOT_MASK) - chunk_base)] = 1;         // I displaced a later row of this chunk
                        done = true;
                    }
                }
            }
            b = (b + 1) & p.bucket_mask;
        }
    }
}

// flags of one chunk -> the sources' flag regions (peer memory), contiguous per source
struct FlagsBack2 {
    const u8* flags; const u32* counts; u32 n_shards, region_rows, me;
    u8* peer_flags[S2_MAX];             // source s: its flag regions of this chunk parity, [n_shards * region_rows]
};
__global__ void __launch_bounds__(256) k_shard_flags_send2(const FlagsBack2 p) {
    for (u32 s = 0; s < p.n_shards; ++s) {
        const u32 n = min(p.counts[s], p.region_rows);
        const u8* src = p.flags + (u64)s * p.region_rows;
        u8* dst = p.peer_flags[s] + (u64)p.me * p.region_rows;
        const u32 n16 = n / 16;
        for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < n16; t += gridDim.x * blockDim.x)
            reinterpret_cast<uint4*>(dst)[t] = reinterpret_cast<const uint4*>(src)[t];
        if (blockIdx.x == 0 && threadIdx.x < n - n16 * 16) dst[n16 * 16 + threadIdx.x] = src[n16 * 16 + threadIdx.x];
    }
}

// source side: flags that came back -> my records, + this chunk's duplicate count
__global__ void __launch_bounds__(256) k_shard_flags2(const u8* __restrict__ flags_in, const u32* __restrict__ dest, const u32* n_ptr,
                                                      u32 region_rows, u8* __restrict__ dup, unsigned long long* n_dups) {
    const u32 n = *n_ptr;
    u32 c = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u32 d = dest[i];
        const u8 f = flags_in[(u64)(d >> 27) * region_rows + (d & 0x7FFFFFFu)];
        dup[i] = f;
        c += f;
    }
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31u) == 0 && c) atomicAdd(n_dups, (unsigned long long)c);
}

}  // namespace fqd
