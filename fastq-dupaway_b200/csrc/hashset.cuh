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

__device__ __forceinline__ bool rows_equal(const u64* a, const u64* b, u32 words) {
    // rows are 16-byte aligned (row_words is even)
    const ulonglong2* pa = reinterpret_cast<const ulonglong2*>(a);
    const ulonglong2* pb = reinterpret_cast<const ulonglong2*>(b);
    u64 diff = 0;
    for (u32 i = 0; i < words / 2; ++i) {
        ulonglong2 x = pa[i], y = __ldcg(pb + i);
        diff |= (x.x ^ y.x) | (x.y ^ y.y);
    }
    return diff == 0;
}

__global__ void __launch_bounds__(HS_THREADS) k_chunk_begin(InsertParams p) {
    // number of records (pairs) this chunk contributes: both mates advance in lock-step and stop at the
    // shorter one (src/hash_dup_remover.hpp:228-230)
    u32 n = p.ctl1->n_records;
    if (p.ctl2) n = min(n, p.ctl2->n_records);
    u64 room = p.key_capacity - p.run->n_records;
    if ((u64)n > room) { n = (u32)room; p.run->capacity_exceeded = 1; }
    p.run->chunk_pairs = n;
    p.run->chunk_dups = 0;
}

__global__ void __launch_bounds__(HS_THREADS) k_insert(const InsertParams p) {
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
                    if (rows_equal(myrow, p.keys + other * p.row_words, p.row_words)) {
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

__global__ void k_chunk_end(RunState* run) {
    run->n_records += run->chunk_pairs;
    run->n_dups += run->chunk_dups;
    run->n_survivors += run->chunk_pairs - run->chunk_dups;
}

__global__ void k_fill_u64(u64* p, u64 n, u64 v) {
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

}  // namespace fqd
