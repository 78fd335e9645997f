// shard.cuh - multi-GPU --fast mode: hash-range ownership of keys across the GPUs of one box.
// Every rank splits + packs its own slice of the input (K1), then partitions the packed keys by owner
//     owner(key) = floor(hash(key) * N / 2^64)
// with one stable radix pass, the ranks exchange (key row, hash) records with ONE all-to-all over NVLink
// (torch.distributed / NCCL, driven by the host binding), every owner inserts what it received into its own
// hash set (K2, first occurrence = smallest position in the global input order), and the duplicate flags travel
// back with a second, byte-sized all-to-all so that the origin rank - which still holds the raw bytes - writes
// its survivors in input order.  (the reference has no counterpart: it is single-threaded.)
#pragma once
#include "common.cuh"

namespace fqd {

// owner of every record of the chunk + finalised hash
__global__ void k_shard_owner(const u64* __restrict__ raw_hash, const u64* __restrict__ raw_hash2, const ChunkCtl* ctl, u32 n_shards,
                              u64* __restrict__ owner_key, u64* __restrict__ final_hash, u32* __restrict__ idx) {
    const u32 n = ctl->n_records;
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const u64 h = raw_hash2 ? pair_hash(raw_hash[i], raw_hash2[i]) : mix64(raw_hash[i]);     // pair key = both mates
        final_hash[i] = h;
        owner_key[i] = __umul64hi(h, (u64)n_shards);
        idx[i] = (u32)i;
    }
}
// per-owner counts from the owner-sorted key list (tiny: n_shards <= 8 values via binary search by one thread each)
__global__ void k_shard_counts(const u64* __restrict__ sorted_owner, const ChunkCtl* ctl, u32 n_shards, u32* counts) {
    const u32 n = ctl->n_records;
    const u32 o = threadIdx.x;
    if (o > n_shards) return;
    u32 lo = 0, hi = n;                      // first position with owner >= o
    while (lo < hi) { u32 mid = (lo + hi) >> 1; if (sorted_owner[mid] < o) lo = mid + 1; else hi = mid; }
    counts[o] = lo;                          // counts[o] = start of owner o; counts[n_shards] = n
}
// send rows: [W words of key][1 word hash], grouped by owner, input order kept inside each group
__global__ void k_shard_gather(const u64* __restrict__ keys, u32 row_words, const u64* __restrict__ final_hash, const u32* __restrict__ sorted_idx,
                               const ChunkCtl* ctl, u64* __restrict__ send) {
    const u32 n = ctl->n_records;
    const u32 rw = row_words + 1;
    u64 step = (u64)gridDim.x * blockDim.x;
    const u64 total = (u64)n * rw;
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += step) {
        const u64 r = t / rw; const u32 w = (u32)(t % rw);
        const u32 src = sorted_idx[r];
        send[t] = w < row_words ? keys[(u64)src * row_words + w] : final_hash[src];
    }
}
// append received rows to the key store and hand their hashes to the insert kernel
__global__ void k_shard_append(const u64* __restrict__ recv, u32 n_recv, u32 row_words, u64* __restrict__ keys, const RunState* run,
                               u64 key_capacity, u64* __restrict__ hash_out) {
    const u32 rw = row_words + 1;
    const u64 base = run->n_records;
    u64 step = (u64)gridDim.x * blockDim.x;
    const u64 total = (u64)n_recv * rw;
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += step) {
        const u64 r = t / rw; const u32 w = (u32)(t % rw);
        if (base + r >= key_capacity) continue;
        if (w < row_words) keys[(base + r) * row_words + w] = recv[t];
        else hash_out[r] = recv[t];
    }
}
__global__ void k_shard_set_pairs(RunState* run, u32 n, u64 key_capacity) {
    u64 room = key_capacity - run->n_records;
    if ((u64)n > room) { n = (u32)room; run->capacity_exceeded = 1; }
    run->chunk_pairs = n; run->chunk_dups = 0;
}
// flags come back grouped by owner in the order the rows were sent: put them at their records
__global__ void k_shard_flags_back(const u8* __restrict__ flags_sorted, const u32* __restrict__ sorted_idx, const ChunkCtl* ctl, u8* __restrict__ dup) {
    const u32 n = ctl->n_records;
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) dup[sorted_idx[i]] = flags_sorted[i];
}

}  // namespace fqd
