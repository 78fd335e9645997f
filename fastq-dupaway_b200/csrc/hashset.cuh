// hashset.cuh - K2/K3: exact first-occurrence membership on the device.
//
// Replaces std::unordered_set<setRecord>::find + insert in HashDupRemover<T>::impl_filterSE / impl_filterPE
// (src/hash_dup_remover.hpp:113-114,133-139,237-245).  The reference's set compares full keys
// (operator==, src/hash_dup_remover.cpp:10-14,26-33) - the hash only picks a bucket - so this
// table, too, verifies the FULL packed key before calling two records equal; the 24-bit tag only filters.
//
// Table: open addressing, buckets of 4 x 64-bit entries (one 32-byte sector), linear probing over buckets.
//   entry = tag(24) << 40 | slot(40);  EMPTY = ~0.   slot = index of the key row in the key store = global
//   record index in input order.  Claim with atomicCAS; when an equal key is found, atomicMin on the entry
//   keeps the smallest slot (= first occurrence in input order, exactly the record the reference writes),
//   and whoever loses is flagged as duplicate.  Rows of earlier chunks always have smaller slots.
#pragma once
#include "common.cuh"

namespace fqd {

constexpr u64 HS_EMPTY = ~0ull;
constexpr u64 HS_SLOT_MASK = (1ull << 40) - 1ull;
constexpr int HS_THREADS = 256;

struct InsertParams {
    u64* table;             // n_buckets * 4 entries
    u32 bucket_shift;       // bucket = hash >> bucket_shift   (n_buckets = 2^(64 - shift))
    u64 bucket_mask;
    const u64* keys;        // key store
    u32 row_words;
    u64 key_capacity;
    const u64* hash1;       // [cap] chunk-local
    const u64* hash2;       // paired: second mate, else nullptr
    const ChunkCtl* ctl1;
    const ChunkCtl* ctl2;   // paired else nullptr
    RunState* run;
    u8* dup;                // [cap] chunk-local flags, zero-initialised
    u64 hash_mul;           // 1, or the number of shards: a shard owns the hash range [k/N,(k+1)/N) and spreads its
                            // keys over its own table with the low 64 bits of hash * N
    u32 hash_final;         // 1: hash1 already holds the finalised 64-bit hash (sharded path)
};

// RW = row width known at compile time (8: one 150 bp mate, 16: a 150 bp pair; 0: any, run-time loop).  With RW known
// every 16-byte load of both rows is issued before the first compare: one DRAM round trip instead of RW / 2.
template <int RW>
__device__ __forceinline__ bool rows_equal(const u64* a, const u64* b, u32 words) {
    // rows are 16-byte aligned (row_words is even)
    const ulonglong2* pa = reinterpret_cast<const ulonglong2*>(a);
    const ulonglong2* pb = reinterpret_cast<const ulonglong2*>(b);
    u64 diff = 0;
    if (RW > 0) {
        ulonglong2 x[RW / 2 > 0 ? RW / 2 : 1], y[RW / 2 > 0 ? RW / 2 : 1];
#pragma unroll
        for (int i = 0; i < RW / 2; ++i) { x[i] = __ldcg(pa + i); y[i] = __ldcg(pb + i); }
#pragma unroll
        for (int i = 0; i < RW / 2; ++i) diff |= (x[i].x ^ y[i].x) | (x[i].y ^ y[i].y);
    } else {
        for (u32 i = 0; i < words / 2; ++i) {
            ulonglong2 x = pa[i], y = __ldcg(pb + i);
            diff |= (x.x ^ y.x) | (x.y ^ y.y);
        }
    }
    return diff == 0;
}

__global__ void __launch_bounds__(HS_THREADS) k_chunk_begin(InsertParams p) {
    // number of records (pairs) this chunk contributes: both mates advance in lock-step and stop at the
    // shorter one (src/hash_dup_remover.hpp:228-230)
    u32 n = p.ctl1->n_records;
    if (p.ctl2) n = min(n, p.ctl2->n_records);
    // A chunk that does not fit is not processed at all (nothing inserted, counters untouched): the host grows the key
    // store and the table in place (grow_fast) and runs the same chunk again.
    u64 room = p.key_capacity - p.run->n_records;
    if ((u64)n > room) { p.run->chunk_wanted = n; n = 0; p.run->capacity_exceeded = 1; }
    p.run->chunk_pairs = n;
    p.run->chunk_dups = 0;
}

// Growth of the set (the reference's unordered_set rehashes as it fills, src/hash_dup_remover.hpp:113-114): every entry
// of the old table is placed in the new, larger one.  The hash is recomputed from the key row (what K1 computed when
// the record was packed); the entries are distinct keys, so there is nothing to compare.
struct RehashParams {
    const u64* old_table; u64 old_entries;
    u64* table; u32 bucket_shift; u64 bucket_mask;
    const u64* keys; u32 row_words; u32 W; u32 mates; u64 hash_mul;
};
__global__ void __launch_bounds__(HS_THREADS) k_rehash(const RehashParams p) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.old_entries; i += stride) {
        const u64 e = p.old_table[i];
        if (e == HS_EMPTY) continue;
        const u64 slot = e & HS_SLOT_MASK;
        const u64* row = p.keys + slot * p.row_words;
        u64 hm[2] = {0, 0};
        for (u32 m = 0; m < p.mates; ++m)
            for (u32 w = 0; w < p.W; ++w) hm[m] += word_hash(row[m * p.W + w], pos_keys(m * 4096u + w));
        u64 h = p.mates == 2 ? pair_hash(hm[0], hm[1]) : mix64(hm[0]);
        h *= p.hash_mul;
        const u64 mine = (((h >> 8) & 0xFFFFFFull) << 40) | slot;
        u64 b = h >> p.bucket_shift;
        for (bool done = false; !done; b = (b + 1) & p.bucket_mask)
            for (int k = 0; k < 4 && !done; ++k)
                if (p.table[b * 4 + k] == HS_EMPTY && atomicCAS(p.table + b * 4 + k, HS_EMPTY, mine) == HS_EMPTY) done = true;
    }
}

template <int RW>
__global__ void __launch_bounds__(HS_THREADS, 4) k_insert(const InsertParams p) {
    const u32 n = p.run->chunk_pairs;
    const u64 slot_base = p.run->n_records;
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u64 h = p.hash1[i];                       // raw multilinear sums from K1
        if (!p.hash_final) h = p.hash2 ? pair_hash(h, p.hash2[i]) : mix64(h);
        h *= p.hash_mul;
        const u64 slot = slot_base + i;
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
                    if (old == HS_EMPTY) { done = true; break; }      // first of its key so far
                    cur = old;
                }
                if ((cur >> 40) == tag) {
                    const u64 other = cur & HS_SLOT_MASK;
                    if (rows_equal<RW>(myrow, p.keys + other * p.row_words, p.row_words)) {
                        u64 old = atomicMin(bp + k, mine);
                        if (old < mine) p.dup[i] = 1;                          // an earlier record holds this key
                        else p.dup[(u32)((old & HS_SLOT_MASK) - slot_base)] = 1; // I displaced a later record
                        done = true;
                    }
                }
            }
            b = (b + 1) & p.bucket_mask;
        }
    }
}

// (A two-pass variant - claim empty slots first, run the full probe only over the ~30 % of records that met a matching
// tag, as dense warps - was measured in round 2 and dropped: 12.5 ms per 100 M-read job against 9.5 ms.  The records put
// aside read their hash and their bucket a second time, and this kernel is bound by the number of random DRAM accesses,
// not by the latency chain of a warp's slowest lane.)
static inline void insert_launch(const InsertParams& p, unsigned grid, cudaStream_t stream) {
    if (p.row_words == 8) k_insert<8><<<grid, HS_THREADS, 0, stream>>>(p);
    else if (p.row_words == 16) k_insert<16><<<grid, HS_THREADS, 0, stream>>>(p);
    else k_insert<0><<<grid, HS_THREADS, 0, stream>>>(p);
}

// K3: count this chunk's duplicates and advance the run counters (single host-sync-free hand-over to the
// next chunk: the next K1 reads run->n_records as its slot base).
__global__ void __launch_bounds__(HS_THREADS) k_count_dups(const u8* dup, RunState* run) {
    __shared__ u32 s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const u32 n = run->chunk_pairs;
    u32 c = 0;
    const u32 stride = gridDim.x * blockDim.x * 16u;
    for (u32 i = (blockIdx.x * blockDim.x + threadIdx.x) * 16u; i < n; i += stride) {
        if (i + 16u <= n) {
            uint4 v = *reinterpret_cast<const uint4*>(dup + i);
            c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);   // flags are 0/1 bytes
        } else {
            for (u32 j = i; j < n; ++j) c += dup[j];
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(&run->chunk_dups, s_cnt);
}

// Survivor list (fqd_keep_survivors): the global input indices of the records that are WRITTEN, ascending, appended chunk
// by chunk - what the reference's `output << record` loop amounts to (src/hash_dup_remover.hpp:136-138,240-244).  Two
// passes over the chunk's flag bytes: survivors per block of SV_BLOCK records, then every block sums the counts before it
// (a few hundred values) and writes its indices at their ranks.
constexpr u32 SV_BLOCK = 8192;
__global__ void __launch_bounds__(HS_THREADS) k_surv_count(const u8* __restrict__ dup, const RunState* run, u32* __restrict__ block_cnt) {
    __shared__ u32 s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const u32 n = run->chunk_pairs;
    const u32 lo = blockIdx.x * SV_BLOCK, hi = min(n, lo + SV_BLOCK);
    u32 c = 0;
    for (u32 i = lo + threadIdx.x * 16u; i < hi; i += HS_THREADS * 16u) {
        if (i + 16u <= hi) {
            const uint4 v = *reinterpret_cast<const uint4*>(dup + i);
            c += 16u - (__popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w));       // flags are 0/1 bytes
        } else {
            for (u32 j = i; j < hi; ++j) c += dup[j] ? 0u : 1u;
        }
    }
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = s_cnt;
}
__global__ void __launch_bounds__(HS_THREADS) k_surv_write(const u8* __restrict__ dup, const RunState* run, const u32* __restrict__ block_cnt,
                                                           u64* __restrict__ out, u64 out_cap) {
    __shared__ u32 s_warp[HS_THREADS / 32];
    __shared__ u32 s_before;
    const u32 n = run->chunk_pairs;
    const u32 lo = blockIdx.x * SV_BLOCK, hi = min(n, lo + SV_BLOCK);
    if (lo >= n) return;
    // survivors of this chunk before my block
    u32 b = 0;
    for (u32 j = threadIdx.x; j < blockIdx.x; j += HS_THREADS) b += block_cnt[j];
    b = __reduce_add_sync(0xFFFFFFFFu, b);
    if (threadIdx.x == 0) s_before = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&s_before, b);
    // my SV_BLOCK / HS_THREADS consecutive records
    constexpr u32 PER = SV_BLOCK / HS_THREADS;
    const u32 first = lo + threadIdx.x * PER;
    u32 keep = 0, c = 0;
#pragma unroll 8
    for (u32 k = 0; k < PER; ++k)
        if (first + k < hi && !dup[first + k]) { keep |= 1u << k; ++c; }
    u32 incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((threadIdx.x & 31) >= (u32)d) incl += x;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    u32 wbase = 0;
    for (u32 w = 0; w < (threadIdx.x >> 5); ++w) wbase += s_warp[w];
    u64 at = run->n_survivors + s_before + wbase + incl - c;
    const u64 slot0 = run->n_records + first;
    while (keep) {
        const u32 k = (u32)__ffs((int)keep) - 1u;
        keep &= keep - 1u;
        if (at < out_cap) out[at] = slot0 + k;
        ++at;
    }
}

__global__ void k_chunk_end(RunState* run, const ChunkCtl* ctl1, const ChunkCtl* ctl2) {
    if (ctl1 && !run->sticky_set) {
        const ChunkCtl* c[2] = {ctl1, ctl2 ? ctl2 : ctl1};
        bool any = false;
        for (int m = 0; m < (ctl2 ? 2 : 1); ++m) any |= c[m]->err_parse != NO_ERR || c[m]->err_base != NO_ERR || (c[m]->too_long & TL_SEQ);
        if (any || run->capacity_exceeded) {
            run->sticky_set = 1; run->sticky_pairs = run->capacity_exceeded ? run->chunk_wanted : run->chunk_pairs; run->sticky_first = run->n_records;
            for (int m = 0; m < 2; ++m) {
                run->sticky_parse[m] = c[m]->err_parse; run->sticky_base[m] = c[m]->err_base;
                run->sticky_too_long[m] = c[m]->too_long; run->sticky_too_long_rec[m] = c[m]->too_long_rec;
            }
        }
    }
    run->n_records += run->chunk_pairs;
    run->n_dups += run->chunk_dups;
    run->n_survivors += run->chunk_pairs - run->chunk_dups;
}

__global__ void k_fill_u64(u64* p, u64 n, u64 v) {
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

}  // namespace fqd
