// common.cuh - shared device helpers for libfqd_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fqd {

typedef unsigned long long u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

// ---------------------------------------------------------------------------------------------------------
// Packed sequence keys.
//   One key "mate row" = W 64-bit words; each word holds 20 bases, 3 bits per base, first base in the most
//   significant position (bits 59..57), so that comparing rows word by word as unsigned integers is the
//   byte order of sequence+'\n' that FastqView::cmp defines (src/fastqview.cpp:56-67):
//        code 0 = end of sequence / padding ('\n' sorts below every base)   A=1  C=2  G=3  N=4  T=5
//   Equality of rows <=> equal length and equal bytes, which is setRecord::operator== on the reference's
//   base-5 packing (src/hash_dup_remover.cpp:10-14, src/seq_utils.cpp:23-49).
constexpr int BASES_PER_WORD = 20;

// Per-chunk control block written by the kernels, read back by the host (or by later kernels).
struct ChunkCtl {
    u32 ticket;            // unused since round 2 (tiles are taken in block-index order)
    u32 n_newlines;        // total '\n' in the chunk (written by the last tile)
    u32 n_records;         // complete records parsed (<= capacity)
    u32 consumed;          // bytes covered by those records
    u64 err_parse;         // min over records of (record << 16 | code << 8 | char); ~0 = none
    u64 err_base;          // min over records of (record << 32 | pos << 8 | char);  ~0 = none
    u32 too_long;          // bit mask (atomicOr): TL_SEQ a sequence exceeded the key row, TL_CAPACITY a record beyond the
                           // key store / record tables, TL_TAG an ID tag longer than the tag rows
    u32 pad;               // 1: byte outside {A,C,G,T,N} in a mode that accepts any byte; 2: byte below '\n' (byte keys)
    u32 too_long_rec;      // smallest chunk-local index of a record with TL_SEQ (~0 = none)
    u32 pad2;
};
enum { TL_SEQ = 1, TL_CAPACITY = 2, TL_TAG = 4 };

// Persistent per-handle device state.
struct RunState {
    u64 n_records;         // records (pairs) inserted so far == next key-store slot
    u64 n_dups;            // duplicates so far
    u64 n_survivors;
    u32 capacity_exceeded; // key store or table overflow
    u32 chunk_pairs;       // pairs in the chunk being processed (min over mates)
    u32 chunk_dups;
    u32 chunk_wanted;      // capacity_exceeded: pairs the refused chunk holds
    // first chunk that raised a data error, kept across chunks for callers that do not read every chunk's control
    // block back (fqd_push_device_async): the mates' error words as K1 left them, and where the chunk began
    u32 sticky_set, sticky_pairs;
    u64 sticky_first;
    u64 sticky_parse[2], sticky_base[2];
    u32 sticky_too_long[2], sticky_too_long_rec[2];
};

constexpr u64 NO_ERR = ~0ull;
enum { PERR_BAD_START = 1, PERR_LEN_MISMATCH = 2 };

// ---------------------------------------------------------------------------------------------------------
// hashing: multilinear over 32-bit limbs with per-position odd keys, finalised by a 64-bit mixer.
__device__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}
// Two odd 32-bit keys for word position w (salted per mate).  A key row hashes to
//     sum_w  lo32(word_w) * keyA(w) + hi32(word_w) * keyB(w)      (mod 2^64, two IMAD.WIDE per word)
// stored raw by K1 and finalised with mix64 by whoever consumes it.  The hash only places keys in the table and
// provides the 24-bit tag; equality is always decided on the full key rows.
__host__ __device__ __forceinline__ uint2 pos_keys(u32 w) {
    u64 z = (u64)(w + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return make_uint2((u32)z | 1u, (u32)(z >> 32) | 1u);
}
__device__ __forceinline__ u64 word_hash(u64 word, uint2 k) {
    return (u64)(u32)word * k.x + (u64)(u32)(word >> 32) * k.y;
}
__device__ __forceinline__ u64 pair_hash(u64 h1, u64 h2) {
    return mix64(h1 + 0x9E3779B97F4A7C15ull * ((h2 << 31) | (h2 >> 33)));
}

// ---------------------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA, 1-D): global -> shared, completion on an mbarrier.
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    u32 ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// HBM -> L2 only (no shared-memory destination): bytes a later CTA will stage
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

__device__ __forceinline__ u64 ld_volatile_u64(const u64* p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(u64* p, u64 v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

}  // namespace fqd
