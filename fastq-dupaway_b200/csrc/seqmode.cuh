// seqmode.cuh - whole-input modes on the device:
//   * sequence-based deduplication: SeqDupRemover<T>::filterSE/PE = ExternalSorter / PairedExternalSorter +
//     the comparator scan (src/seq_dup_remover.hpp:40-218, src/external_sort.hpp, src/paired_external_sort.hpp,
//     src/comparator.cpp:45-91)
//   * --fast --unordered: two ID-tag sorts + merge-join + pair set (src/hash_dup_remover.hpp:150-192,257-347)
// The raw input stays in HBM (segments of at most max_chunk_bytes), K1 (parse_pack.cuh) splits and packs it,
// sortlib.cuh sorts record indices by packed key rows (stable on the input index), and the scans below decide
// which records are written.  Nothing spills to disk; nothing is computed on the host.
#pragma once
#include "shard.cuh"
#include <chrono>
#include <cstdio>
#include <cstdlib>
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
// Rows hold `bits` bits per symbol: 3 (codes of {A,C,G,T,N}, 20 per word) or 8 (raw bytes, 8 per word).
// first nb symbols equal?
__device__ __forceinline__ bool prefix_equal(const u64* a, const u64* b, u32 nb, u32 bits) {
    const u32 per = bits == 8u ? 8u : (u32)BASES_PER_WORD;
    const u32 full = nb / per, rem = nb % per;
    u64 diff = 0;
    for (u32 w = 0; w < full; ++w) diff |= a[w] ^ b[w];
    if (rem) diff |= (a[full] ^ b[full]) & (~0ull << (bits * (per - rem)));
    return diff == 0;
}
// number of differing symbols (SeqUtils::hammingDistance, src/seq_utils.cpp:65-72, on equal-length sequences)
__device__ __forceinline__ u32 hamming_words(const u64* a, const u64* b, u32 W, u32 bits) {
    u32 d = 0;
    for (u32 w = 0; w < W; ++w) {
        u64 x = a[w] ^ b[w];
        if (bits == 8u) {
            x |= x >> 4; x |= x >> 2; x |= x >> 1;
            x &= 0x0101010101010101ull;
        } else {
            x = (x | (x >> 1) | (x >> 2)) & 0x0249249249249249ull;
        }
        d += (u32)__popcll(x);
    }
    return d;
}

struct ScanParams {
    const u64* rows; u32 stride; u32 W; u32 mates;
    const u32* len0; const u32* len1;     // sequence lengths (bases) per record and mate
    const u32* perm; u64 n; u32 dist;
    u32 bits;                             // bits per symbol of the key rows (3 or 8)
    u32* keep;                            // [n] in sorted order: 1 = written
    u32* brk;                             // hamming: definite cluster breaks
    u32* tail_head;                       // hamming: record that is the cluster head after the last sorted record
                                          // (0xFFFFFFFF: the head handed in by the previous key range)
};

// Boundary state between two key ranges of a sorted stream that is split across GPUs (multi-GPU sequence mode).
// Layout in 64-bit words, rw = row words:  [0,rw) last sorted row   [rw] its lengths (mate 1 | mate 2 << 32)
//   [rw+1, 2rw+1) row of the cluster head after the last record   [2rw+1] its lengths   [2rw+2] 1 = valid
__host__ __device__ inline u32 boundary_words(u32 rw) { return 2u * rw + 4u; }

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
            bool dup = prefix_equal(a, b, min(lc1, lh1), p.bits);
            if (dup && p.mates == 2) {
                const u32 lc2 = p.len1[c], lh2 = p.len1[h];
                dup = prefix_equal(a + p.W, b + p.W, min(lc2, lh2), p.bits);
                if (dup) dup = ((lh1 <= lc1) && (lh2 <= lc2)) || ((lh1 > lc1) && (lh2 > lc2));
            }
            k = dup ? 0u : 1u;
        }
        p.keep[i] = k;
    }
}
// ---- warp-cooperative adjacent scan (tight, loose, the Hamming breaks).  The thread-per-record kernels above read two
// rows per record at random (its own and its predecessor's: 2 x 128 bytes per pair, 12.8 GB at 50 M pairs).  Here a warp
// walks a run of consecutive sorted positions; a row is 16 bytes in each of Q = row_words / 2 neighbouring lanes, loaded
// ONCE, and the predecessor's 16 bytes come from the lane group before (shuffle) - or, for the first row of a step, from
// the registers that kept the last row of the step before.
constexpr u32 COOP_RUN = 512;           // sorted positions per warp
template <int MODE>                     // 1 tight, 2 loose, 3 hamming breaks
__global__ void __launch_bounds__(256) k_scan_coop(const ScanParams p, u32 Q) {
    const u32 lane = threadIdx.x & 31u;
    const u64 warp_id = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 lo = warp_id * COOP_RUN;
    if (lo >= p.n) return;
    const u64 hi = min(p.n, lo + COOP_RUN);
    const u32 rows_per_step = 32u / Q;
    const u32 piece = lane % Q, grp = lane / Q;
    const u32 Qm = p.W / 2u;                                   // pieces per mate
    const u32 mate = piece / Qm;                               // which mate this lane's 16 bytes belong to
    const u32 sym_per_word = p.bits == 8u ? 8u : (u32)BASES_PER_WORD;
    const u32 first_sym = (piece % Qm) * 2u * sym_per_word;    // first symbol of the mate covered by this piece
    u32* out = MODE == 3 ? p.brk : p.keep;
    // the row (and lengths) before position lo, held like "the last row of the previous step"
    ulonglong2 carry = make_ulonglong2(0ull, 0ull);
    u32 carry_len = 0;
    if (lo > 0) {
        const u32 h = p.perm[lo - 1];
        carry = *reinterpret_cast<const ulonglong2*>(p.rows + (u64)h * p.stride + 2u * piece);
        if (MODE != 1) carry_len = mate ? p.len1[h] : p.len0[h];
    }
    for (u64 base = lo; base < hi; base += rows_per_step) {
        const u64 i = base + grp;
        const bool in = i < hi;
        ulonglong2 v = make_ulonglong2(0ull, 0ull);
        u32 len = 0;
        if (in) {
            const u32 c = p.perm[i];
            v = *reinterpret_cast<const ulonglong2*>(p.rows + (u64)c * p.stride + 2u * piece);
            if (MODE != 1) len = mate ? p.len1[c] : p.len0[c];
        }
        // predecessor's piece: the lane group before, or the carry for the first row of the step
        ulonglong2 pv;
        pv.x = __shfl_up_sync(0xFFFFFFFFu, v.x, Q);
        pv.y = __shfl_up_sync(0xFFFFFFFFu, v.y, Q);
        u32 plen = __shfl_up_sync(0xFFFFFFFFu, len, Q);
        if (grp == 0) { pv = carry; plen = carry_len; }
        // next step's carry: the last row of this step (lanes 32 - Q + piece)
        carry.x = __shfl_sync(0xFFFFFFFFu, v.x, 32u - Q + piece);
        carry.y = __shfl_sync(0xFFFFFFFFu, v.y, 32u - Q + piece);
        carry_len = __shfl_sync(0xFFFFFFFFu, len, 32u - Q + piece);
        bool bad;                                               // this piece says "not a duplicate of the predecessor"
        if (MODE == 1) {
            bad = ((v.x ^ pv.x) | (v.y ^ pv.y)) != 0ull;
        } else if (MODE == 2) {
            // the predecessor's mate is a prefix of mine over min(len) symbols (src/comparator.cpp:60-63)
            const u32 nb = min(len, plen);
            u64 dx = v.x ^ pv.x, dy = v.y ^ pv.y;
            const u32 s0 = first_sym, s1 = first_sym + sym_per_word;
            if (nb <= s0) dx = 0; else if (nb < s0 + sym_per_word) dx &= ~0ull << (p.bits * (sym_per_word - (nb - s0)));
            if (nb <= s1) dy = 0; else if (nb < s1 + sym_per_word) dy &= ~0ull << (p.bits * (sym_per_word - (nb - s1)));
            bad = (dx | dy) != 0ull;
        } else {
            u64 x = v.x ^ pv.x, y = v.y ^ pv.y;
            if (p.bits == 8u) {
                x |= x >> 4; x |= x >> 2; x |= x >> 1; x &= 0x0101010101010101ull;
                y |= y >> 4; y |= y >> 2; y |= y >> 1; y &= 0x0101010101010101ull;
            } else {
                x = (x | (x >> 1) | (x >> 2)) & 0x0249249249249249ull;
                y = (y | (y >> 1) | (y >> 2)) & 0x0249249249249249ull;
            }
            u32 d = (u32)__popcll(x) + (u32)__popcll(y);
            for (u32 o = 1; o < Qm; o <<= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, o);      // distance of the whole mate
            bad = d > 2u * p.dist || len != plen;
        }
        const u32 votes = __ballot_sync(0xFFFFFFFFu, bad);
        // paired loose: lengths of the second mate, from the first lane that holds it (every lane takes part in the shuffle)
        const u32 src2 = p.mates == 2 ? grp * Q + Qm : lane;
        const u32 lc2 = __shfl_sync(0xFFFFFFFFu, len, src2), lh2 = __shfl_sync(0xFFFFFFFFu, plen, src2);
        if (in && piece == 0) {
            const u32 gmask = (Q == 32u ? 0xFFFFFFFFu : ((1u << Q) - 1u)) << (grp * Q);
            bool brk = (votes & gmask) != 0u;
            if (MODE == 2 && !brk && p.mates == 2) {
                // the overlap must be same-sided (src/comparator.cpp:65-74)
                const u32 lc1 = len, lh1 = plen;
                brk = !(((lh1 <= lc1) && (lh2 <= lc2)) || ((lh1 > lc1) && (lh2 > lc2)));
            }
            out[i] = (i == 0 || brk) ? 1u : 0u;
        }
    }
}

// The literal loop of SeqDupRemover::impl_filterSE/PE with the LooseComparator (src/seq_dup_remover.hpp:78-101,173-208,
// src/comparator.cpp:60-74), one thread.  Only used when a sequence holds a byte below the line feed: such a byte sorts
// an extension BEFORE its prefix, and the head of a cluster is then no longer the previous record.
__global__ void k_scan_loose_literal(const ScanParams p) {
    if (p.n == 0) return;
    u32 h = p.perm[0];
    p.keep[0] = 1;
    for (u64 i = 1; i < p.n; ++i) {
        const u32 c = p.perm[i];
        const u64* a = p.rows + (u64)c * p.stride;
        const u64* b = p.rows + (u64)h * p.stride;
        const u32 lc1 = p.len0[c], lh1 = p.len0[h];
        u32 lc2 = 0, lh2 = 0;
        bool dup = prefix_equal(a, b, min(lc1, lh1), p.bits);
        if (dup && p.mates == 2) {
            lc2 = p.len1[c]; lh2 = p.len1[h];
            dup = prefix_equal(a + p.W, b + p.W, min(lc2, lh2), p.bits);
            if (dup) dup = ((lh1 <= lc1) && (lh2 <= lc2)) || ((lh1 > lc1) && (lh2 > lc2));
        }
        p.keep[i] = dup ? 0u : 1u;
        if (!dup) h = c;
        else if (lh1 <= lc1 && (p.mates == 1 || lh2 <= lc2)) h = c;      // keep the longest as the reference
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
            bool same = p.len0[c] == p.len0[h] && hamming_words(a, q, p.W, p.bits) <= 2u * p.dist;
            if (same && p.mates == 2) same = p.len1[c] == p.len1[h] && hamming_words(a + p.W, q + p.W, p.W, p.bits) <= 2u * p.dist;
            b = same ? 0u : 1u;
        }
        p.brk[i] = b;
    }
}
// One thread per segment walks the first HAM_SHORT records of its segment (segments are cluster-sized on ordinary
// libraries); a segment that is longer - an amplicon, a low-complexity repeat, a heavily duplicated library: it can be
// most of the input - is handed, with the head reached so far, to k_ham_long, where a whole block tests 256 candidates
// against the current head at a time.  (Round 1 walked every segment to its end in one thread: a million-member
// Hamming chain was a million dependent DRAM round trips on one lane.)
constexpr u32 HAM_SHORT = 48;
struct HamLong { u64 next; u32 head; u32 pad; };
__device__ __forceinline__ bool ham_dup(const ScanParams& p, u32 c, const u64* q) {
    const u64* a = p.rows + (u64)c * p.stride;
    // lengths are equal inside a segment (a length change is a break)
    bool dup = hamming_words(a, q, p.W, p.bits) <= p.dist;
    if (dup && p.mates == 2) dup = hamming_words(a + p.W, q + p.W, p.W, p.bits) <= p.dist;
    return dup;
}
__global__ void k_ham_segments(const ScanParams p, HamLong* longs, u32* n_long) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        if (!p.brk[i]) continue;
        p.keep[i] = 1;
        u32 head = p.perm[i];
        u64 j = i + 1;
        const u64 stop = min(p.n, i + 1 + HAM_SHORT);
        for (; j < stop && !p.brk[j]; ++j) {
            const u32 c = p.perm[j];
            const bool dup = ham_dup(p, c, p.rows + (u64)head * p.stride);
            p.keep[j] = dup ? 0u : 1u;
            if (!dup) head = c;
        }
        if (j == p.n) { if (p.tail_head) *p.tail_head = head; }      // cluster head at the end of this key range
        else if (j == stop && !p.brk[j]) {                            // the segment goes on: a block takes over
            const u32 k = atomicAdd(n_long, 1u);
            longs[k].next = j; longs[k].head = head;
        }
    }
}
// The greedy scan of one long segment, 256 candidates per step: everything before the first candidate that is NOT a
// duplicate of the current head is one (keep = 0); that candidate becomes the head (keep = 1) and the scan goes on
// behind it.  Same result as the sequential loop of src/seq_dup_remover.hpp:78-101 with src/comparator.cpp:76-91.
__global__ void __launch_bounds__(256) k_ham_long(const ScanParams p, const HamLong* longs, const u32* n_long) {
    __shared__ u32 s_first;              // offset in the batch of the first non-duplicate / break (256 = none)
    __shared__ u32 s_is_break;
    const u32 t = threadIdx.x;
    for (u32 e = blockIdx.x; e < *n_long; e += gridDim.x) {
        u64 j = longs[e].next;
        u32 head = longs[e].head;
        for (;;) {
            if (t == 0) { s_first = 256u; s_is_break = 0; }
            __syncthreads();
            const u64 idx = j + t;
            u32 what = 0;                // 1 = break (or end of the stream), 2 = not a duplicate of the head
            u32 c = 0;
            if (idx >= p.n || p.brk[idx]) what = 1;
            else { c = p.perm[idx]; if (!ham_dup(p, c, p.rows + (u64)head * p.stride)) what = 2; }
            if (what) atomicMin(&s_first, t);
            __syncthreads();
            const u32 first = s_first;
            if (t < first) p.keep[idx] = 0;                      // duplicates of the head
            if (t == first && what == 1) s_is_break = 1;
            if (t == first && what == 2) p.keep[idx] = 1;
            __syncthreads();
            if (first == 256u) { j += 256; continue; }
            if (s_is_break) {
                if (j + first >= p.n && t == 0 && p.tail_head) *p.tail_head = head;
                break;
            }
            head = p.perm[j + first];                            // every thread reads the new head itself
            j += first + 1;
        }
        __syncthreads();
    }
}

// ---- the first sorted records of a key range, re-evaluated against the boundary state of the range before it
__global__ void k_fix_first_tight(const ScanParams p, const u64* prev) {
    if (p.n == 0 || !prev[2 * p.stride + 2]) return;
    p.keep[0] = rows_equal_range(p.rows + (u64)p.perm[0] * p.stride, prev, 0, p.W * p.mates) ? 0u : 1u;
}
__global__ void k_fix_first_loose(const ScanParams p, const u64* prev) {
    if (p.n == 0 || !prev[2 * p.stride + 2]) return;
    const u32 c = p.perm[0];
    const u64* a = p.rows + (u64)c * p.stride;
    const u64 pl = prev[p.stride];
    const u32 lc1 = p.len0[c], lh1 = (u32)pl;
    bool dup = prefix_equal(a, prev, min(lc1, lh1), p.bits);
    if (dup && p.mates == 2) {
        const u32 lc2 = p.len1[c], lh2 = (u32)(pl >> 32);
        dup = prefix_equal(a + p.W, prev + p.W, min(lc2, lh2), p.bits);
        if (dup) dup = ((lh1 <= lc1) && (lh2 <= lc2)) || ((lh1 > lc1) && (lh2 > lc2));
    }
    p.keep[0] = dup ? 0u : 1u;
}
// tail-hamming: is the first record a definite break against the last record of the previous range?  If not, the
// first segment belongs to the previous range's last cluster: scan it again, starting from that cluster's head.
__global__ void k_fix_first_ham(const ScanParams p, const u64* prev) {
    if (p.n == 0 || !prev[2 * p.stride + 2]) return;
    const u32 c0 = p.perm[0];
    const u64* a0 = p.rows + (u64)c0 * p.stride;
    const u64 pl = prev[p.stride];
    bool same = p.len0[c0] == (u32)pl && hamming_words(a0, prev, p.W, p.bits) <= 2u * p.dist;
    if (same && p.mates == 2) same = p.len1[c0] == (u32)(pl >> 32) && hamming_words(a0 + p.W, prev + p.W, p.W, p.bits) <= 2u * p.dist;
    if (!same) { p.brk[0] = 1; return; }
    p.brk[0] = 0;
    const u64* hq = prev + p.stride + 1;          // head of the previous range's last cluster
    u32 head = 0xFFFFFFFFu;
    u64 j = 0;
    for (; j < p.n && (j == 0 || !p.brk[j]); ++j) {
        const u32 c = p.perm[j];
        const u64* a = p.rows + (u64)c * p.stride;
        bool dup = hamming_words(a, hq, p.W, p.bits) <= p.dist;
        if (dup && p.mates == 2) dup = hamming_words(a + p.W, hq + p.W, p.W, p.bits) <= p.dist;
        p.keep[j] = dup ? 0u : 1u;
        if (!dup) { head = c; hq = a; }
    }
    if (j == p.n) *p.tail_head = head;
}
// boundary state after the last sorted record of this range (prev: the state this range started from, or nullptr)
__global__ void k_tail_state(const ScanParams p, const u64* prev, int hamming, u64* out) {
    const u32 rw = p.stride;
    const u32 t = threadIdx.x;
    const u32 last = p.perm[p.n - 1];
    u32 head = last;
    bool ext = false;
    if (hamming) { head = *p.tail_head; ext = head == 0xFFFFFFFFu; }
    for (u32 w = t; w < rw; w += blockDim.x) {
        out[w] = p.rows[(u64)last * rw + w];
        out[rw + 1 + w] = ext ? prev[rw + 1 + w] : p.rows[(u64)head * rw + w];
    }
    if (t == 0) {
        out[rw] = (u64)p.len0[last] | (p.len1 ? (u64)p.len1[last] << 32 : 0ull);
        out[2 * rw + 1] = ext ? prev[2 * rw + 1] : ((u64)p.len0[head] | (p.len1 ? (u64)p.len1[head] << 32 : 0ull));
        out[2 * rw + 2] = 1; out[2 * rw + 3] = 0;
    }
}

// ---- repartition of the input records by key range (multi-GPU sequence mode)
// owner of record i = number of splitters <= (word 0, word 1) of its row (word 1 = 0 for rows of a single word)
__global__ void k_range_owner(const u64* rows, u32 stride, u32 nw, u64 n, const u64* splitters, u32 n_split, u64* owner_key, u32* idx) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const u64 w0 = rows[i * stride], w1 = nw > 1 ? rows[i * stride + 1] : 0ull;
        u32 lo = 0, hi = n_split;            // first splitter > key
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            const u64 s0 = splitters[2 * mid], s1 = splitters[2 * mid + 1];
            if (s0 < w0 || (s0 == w0 && s1 <= w1)) lo = mid + 1; else hi = mid;
        }
        owner_key[i] = lo;
        idx[i] = (u32)i;
    }
}
__global__ void k_range_starts(const u64* sorted_owner, u64 n, u32 n_shards, u64* starts) {
    const u32 o = threadIdx.x;
    if (o > n_shards) return;
    u64 lo = 0, hi = n;                      // first position with owner >= o
    while (lo < hi) { u64 mid = (lo + hi) >> 1; if (sorted_owner[mid] < o) lo = mid + 1; else hi = mid; }
    starts[o] = lo;
}
__global__ void k_gather_u32(const u32* src, const u32* perm, u64 n, u32* out) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) out[i] = src[perm[i]];
}
__global__ void k_pick_u64(const u64* src, const u64* pos, u32 n, u64 total_pos, u64 total, u64* out) {
    const u32 t = threadIdx.x;
    if (t < n) out[t] = pos[t] >= total_pos ? total : src[pos[t]];
}
__global__ void k_sample_rows(const u64* rows, u32 stride, u32 nw, u64 n, u32 n_samples, u64* out) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_samples) return;
    const u64 i = (u64)(((unsigned __int128)j * n) / n_samples);
    out[2 * j] = rows[i * stride]; out[2 * j + 1] = nw > 1 ? rows[i * stride + 1] : 0ull;
}
// ---- record gathers.  A warp takes 32 records at a time: every lane fetches the geometry of ONE record (coalesced
// table reads, the segment search in parallel), then the warp copies the 32 records one after the other with the
// geometry broadcast by shuffles.  (One record per warp-iteration left the copy waiting ~2 us for four dependent
// loads per 322 bytes: 0.8 TB/s; source and destination are byte-aligned only, so the copy itself moves bytes.)
__device__ __forceinline__ const u8* seg_locate(u64 o, const u64* seg_base, u8* const* seg_ptr, u32 n_segs) {
    u32 lo = 0, hi = n_segs;             // last segment whose logical base is <= o
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (seg_base[mid] <= o) lo = mid; else hi = mid; }
    return seg_ptr[lo] + (o - seg_base[lo]);
}
// copy `n` bytes src -> dst (+ `pre` dashes in front) for the up-to-32 records held one per lane
__device__ __forceinline__ void warp_copy_records(const u8* my_src, u8* my_dst, u32 my_n, u32 my_pre, u32 n_valid, u32 lane) {
    for (u32 j = 0; j < n_valid; ++j) {
        const u8* src = reinterpret_cast<const u8*>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(my_src), j));
        u8* d = reinterpret_cast<u8*>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(my_dst), j));
        const u32 n = __shfl_sync(0xFFFFFFFFu, my_n, j);
        const u32 pre = __shfl_sync(0xFFFFFFFFu, my_pre, j);
        if (lane < pre) d[lane] = '-';
        d += pre;
        u32 i = lane;
        for (; i + 96u < n; i += 128u) {          // four independent loads in flight per lane
            const u8 a = src[i], b = src[i + 32u], c = src[i + 64u], e = src[i + 96u];
            d[i] = a; d[i + 32u] = b; d[i + 64u] = c; d[i + 96u] = e;
        }
        for (; i < n; i += 32u) d[i] = src[i];
    }
}

// records in `perm` order, packed back to back at dst[r] (segmented input)
__global__ void k_gather_records_perm(const u64* rec_off, const u32* rec_len, const u32* perm, const u64* dst, u64 count,
                                      const u64* seg_base, u8* const* seg_ptr, u32 n_segs, u8* out) {
    const u32 lane = threadIdx.x & 31u;
    const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 r0 = (((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; r0 < count; r0 += warps * 32u) {
        const u64 r = r0 + lane;
        const u8* src = nullptr; u8* d = nullptr; u32 n = 0;
        if (r < count) {
            const u32 g = perm[r];
            src = seg_locate(rec_off[g], seg_base, seg_ptr, n_segs);
            d = out + dst[r];
            n = rec_len[g];
        }
        warp_copy_records(src, d, n, 0u, (u32)min((u64)32, count - r0), lane);
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
// --fast --unordered
// ID tag of a record (FastqViewWithId::read_new, src/fastqview.cpp:190-204; Fasta twin src/fastaview.cpp:153-167):
// after the first '.' of the ID line (anywhere in it, description and '\n' included), else after the lead
// character; up to the first ' ' at/after the tag start, else to the end of the line INCLUDING the '\n'.
// Tag order = strncmp + shorter-first (src/fastqview.cpp:168-178): the bytes are packed big-endian, 8 per word,
// zero padded, so that unsigned word-by-word comparison is that order.
__global__ void k_extract_tags(const u8* raw, const u32* rec_start, const ChunkCtl* ctl, u32 TW, u64* tags, ChunkCtl* ctl_out) {
    const u32 n = ctl->n_records;
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += step) {
        const u8* p = raw + rec_start[r];
        const u32 rec_len = rec_start[r + 1] - rec_start[r];
        u32 idlen = 0, dot = 0xFFFFFFFFu;
        for (u32 i = 0; i < rec_len; ++i) {
            const u8 c = p[i];
            if (c == '.' && dot == 0xFFFFFFFFu) dot = i;
            if (c == '\n') { idlen = i + 1; break; }
        }
        const u32 t0 = dot != 0xFFFFFFFFu ? dot + 1 : 1u;
        u32 t1 = idlen;
        for (u32 i = t0; i < idlen; ++i) if (p[i] == ' ') { t1 = i; break; }
        const u32 tl = t1 > t0 ? t1 - t0 : 0u;
        if (tl > TW * 8u) atomicOr(&ctl_out->too_long, (u32)TL_TAG);
        u64* out = tags + r * TW;
        for (u32 w = 0; w < TW; ++w) {
            u64 v = 0;
            for (u32 b = 0; b < 8; ++b) {
                const u32 k = w * 8 + b;
                v = (v << 8) | (k < tl ? (u64)p[t0 + k] : 0ull);
            }
            out[w] = v;
        }
    }
}

__device__ __forceinline__ int cmp_rows(const u64* a, const u64* b, u32 nw) {
    for (u32 w = 0; w < nw; ++w) {
        if (a[w] < b[w]) return -1;
        if (a[w] > b[w]) return 1;
    }
    return 0;
}
// first position in the tag-sorted list (perm over rows `tags`) whose tag is >= key (upper = false) or > key
__device__ __forceinline__ u32 tag_bound(const u64* tags, const u32* perm, u32 n, u32 TW, const u64* key, bool upper) {
    u32 lo = 0, hi = n;
    while (lo < hi) {
        const u32 mid = lo + ((hi - lo) >> 1);
        const int c = cmp_rows(tags + (u64)perm[mid] * TW, key, TW);
        if (c < 0 || (upper && c == 0)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct JoinParams {
    const u64* tagsL; const u32* permL; u32 n;
    const u64* tagsR; const u32* permR; u32 m;
    u32 TW;
    u32* matchL;        // [n] partner position in R's sorted list or ~0
    u32* matchR;        // [m] partner position in L's sorted list or ~0 (pre-filled)
};
struct JoinStop { u32 is, js, limit_i, limit_j, final_equal, use_a; };

// The two-pointer walk of impl_filterPE_unordered (src/hash_dup_remover.hpp:279-315) pairs the k-th record of a
// tag in L with the k-th record of the same tag in R.
__global__ void k_join_match(const JoinParams p) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        const u64* key = p.tagsL + (u64)p.permL[i] * p.TW;
        const u32 g0 = tag_bound(p.tagsL, p.permL, p.n, p.TW, key, false);
        const u32 rank = (u32)i - g0;
        const u32 lb = tag_bound(p.tagsR, p.permR, p.m, p.TW, key, false);
        const u32 ub = tag_bound(p.tagsR, p.permR, p.m, p.TW, key, true);
        u32 partner = 0xFFFFFFFFu;
        if (rank < ub - lb) { partner = lb + rank; p.matchR[partner] = (u32)i; }
        p.matchL[i] = partner;
    }
}
// Where the walk stops: the loop ends as soon as either side has fetched its LAST record; exactly
// one more comparison is made there.  enter_j(i) = R position when the walk first stands on L[i].
__device__ __forceinline__ u32 join_enter(const u64* tagsA, const u32* permA, u32 na, const u64* tagsB, const u32* permB, u32 nb, u32 TW, u32 i) {
    if (i == 0) return 0;
    const u64* key = tagsA + (u64)permA[i - 1] * TW;
    const u32 g0 = tag_bound(tagsA, permA, na, TW, key, false);
    const u32 rank = (i - 1) - g0;
    const u32 lb = tag_bound(tagsB, permB, nb, TW, key, false);
    const u32 ub = tag_bound(tagsB, permB, nb, TW, key, true);
    return lb + min(rank + 1u, ub - lb);
}
__global__ void k_join_stop(const JoinParams p, JoinStop* out) {
    const u32 ja = join_enter(p.tagsL, p.permL, p.n, p.tagsR, p.permR, p.m, p.TW, p.n - 1);
    JoinStop s;
    if (ja < p.m - 1) { s.use_a = 1; s.is = p.n - 1; s.js = ja; }
    else { s.use_a = 0; s.js = p.m - 1; s.is = join_enter(p.tagsR, p.permR, p.m, p.tagsL, p.permL, p.n, p.TW, p.m - 1); }
    s.limit_i = s.is; s.limit_j = s.js;
    s.final_equal = 0;
    if (s.is < p.n && s.js < p.m)
        s.final_equal = cmp_rows(p.tagsL + (u64)p.permL[s.is] * p.TW, p.tagsR + (u64)p.permR[s.js] * p.TW, p.TW) == 0 ? 1u : 0u;
    *out = s;
}
// Tag ranges across GPUs: the driver knows where the job's walk stops and hands every range its part of that state
// (limits clamped to the range, is / js = ~0 when the stop state's record lives in another range); equal tags always
// share a range, so the last comparison can only succeed where both of its records are.
__global__ void k_join_final(const JoinParams p, JoinStop* st) {
    JoinStop s = *st;
    s.final_equal = 0;
    if (s.is < p.n && s.js < p.m)
        s.final_equal = cmp_rows(p.tagsL + (u64)p.permL[s.is] * p.TW, p.tagsR + (u64)p.permR[s.js] * p.TW, p.TW) == 0 ? 1u : 0u;
    *st = s;
}
// position in the other list when the walk first stands on element i of this one (side 0: this = L)
__global__ void k_join_enter(const JoinParams p, int side, u32 i, u32* out) {
    *out = side == 0 ? join_enter(p.tagsL, p.permL, p.n, p.tagsR, p.permR, p.m, p.TW, i)
                     : join_enter(p.tagsR, p.permR, p.m, p.tagsL, p.permL, p.n, p.TW, i);
}
__global__ void k_join_flags(const JoinParams p, const JoinStop* st, u32* emit, u32* unL, u32* unR) {
    const JoinStop s = *st;
    u64 step = (u64)gridDim.x * blockDim.x;
    const u64 tot = (u64)max(p.n, p.m);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += step) {
        if (i < p.n) {
            const u32 pr = p.matchL[i];
            // pairs visited before the stop state, plus the stop state itself when its tags are equal
            const bool before = pr != 0xFFFFFFFFu && i < s.limit_i && pr < s.limit_j;
            const bool last = i == s.is && s.final_equal;
            emit[i] = (before || last) ? 1u : 0u;
            unL[i] = (i < s.limit_i && pr == 0xFFFFFFFFu) ? 1u : 0u;
        }
        if (i < p.m) unR[i] = (i < s.limit_j && p.matchR[i] == 0xFFFFFFFFu) ? 1u : 0u;
    }
}
// pair key store for the emitted pairs, in emission order
__global__ void k_build_pairs(const JoinParams p, const JoinStop* st, const u32* emit, const u32* excl, const u64* rows, u32 stride, u32 W,
                              const u64* hashL, const u64* hashR, const u32* badL, const u32* badR,
                              u64* pair_rows, u64* h1, u64* h2, u32* idxL, u32* idxR, u64* first_bad) {
    const JoinStop s = *st;
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        if (!emit[i]) continue;
        const u32 e = excl[i];
        const u32 pr = (i == s.is && s.final_equal) ? s.js : p.matchL[i];
        const u32 gl = p.permL[i], gr = p.permR[pr];
        idxL[e] = gl; idxR[e] = gr;
        const u64* a = rows + (u64)gl * stride;
        const u64* b = rows + (u64)gr * stride + W;
        u64* o = pair_rows + (u64)e * 2 * W;
        for (u32 w = 0; w < W; ++w) { o[w] = a[w]; o[W + w] = b[w]; }
        h1[e] = hashL[gl]; h2[e] = hashR[gr];
        // setRecordPair keys the left mate first (src/hash_dup_remover.cpp:16-24): its bad byte is reported first
        const u32 bl = badL[gl], br = badR[gr];
        if (bl != 0xFFFFFFFFu) atomicMin(first_bad, ((u64)e << 8) | (bl & 0xFFu));
        else if (br != 0xFFFFFFFFu) atomicMin(first_bad, ((u64)e << 8) | (br & 0xFFu));
    }
}
// pairs of one tag range -> the GPUs that own their hash range (same rule and row layout as shard.cuh)
__global__ void k_un_owner(const u64* __restrict__ h1, const u64* __restrict__ h2, u64 n, u32 n_shards, u64* __restrict__ owner_key, u32* __restrict__ idx) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        owner_key[i] = __umul64hi(pair_hash(h1[i], h2[i]), (u64)n_shards);
        idx[i] = (u32)i;
    }
}
__global__ void k_un_gather(const u64* __restrict__ pair_rows, u32 row_words, const u64* __restrict__ h1, const u64* __restrict__ h2,
                            const u32* __restrict__ sorted_idx, u64 n, u64* __restrict__ send) {
    const u32 rw = row_words + 1;
    u64 step = (u64)gridDim.x * blockDim.x;
    const u64 total = n * rw;
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += step) {
        const u64 r = t / rw; const u32 w = (u32)(t % rw);
        const u32 src = sorted_idx[r];
        send[t] = w < row_words ? pair_rows[(u64)src * row_words + w] : pair_hash(h1[src], h2[src]);
    }
}
__global__ void k_un_flags_back(const u8* __restrict__ flags_sorted, const u32* __restrict__ sorted_idx, u64 n, u8* __restrict__ dup) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) dup[sorted_idx[i]] = flags_sorted[i];
}
__global__ void k_keep_from_dup(const u8* dup, u64 n, u64 limit, u32* keep) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) keep[i] = (i < limit && !dup[i]) ? 1u : 0u;
}
__global__ void k_emit_pairs(const u32* keep, const u32* excl, u64 n, const u32* idxL, const u32* idxR, const u64* off0, const u32* len0,
                             const u64* off1, const u32* len1, u64* o_off0, u32* o_len0, u64* o_off1, u32* o_len1) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        if (!keep[i]) continue;
        const u32 o = excl[i];
        o_off0[o] = off0[idxL[i]]; o_len0[o] = len0[idxL[i]];
        o_off1[o] = off1[idxR[i]]; o_len1[o] = len1[idxR[i]];
    }
}
__global__ void k_set_chunk_pairs(RunState* run, u32 n) { run->n_records = 0; run->chunk_pairs = n; run->chunk_dups = 0; }

// ---------------------------------------------------------------------------------------------------------
// owned: pool memory holding a copy of the input.  !owned: a view into a caller's buffer (fqd_adopt_device): d is the
// view's start aligned down to 16 bytes (bulk copies need it), the view proper begins `skip` bytes later.
struct SeqSegment { u8* d = nullptr; size_t cap = 0, fill = 0; u64 logical_base = 0; bool owned = true; u32 skip = 0; };

struct SeqMate {
    std::vector<SeqSegment> segs;
    bool adopted = false;            // the input is a caller's buffer, cut into views at record boundaries
    RunState* d_run = nullptr;
    u64 n_records = 0;
    u64* d_rec_off = nullptr;
    u32* d_rec_len = nullptr;
    u32* d_seq_len = nullptr;
    u64* d_hash = nullptr;        // unordered: raw key hash per record
    u32* d_bad = nullptr;         // unordered: first byte outside {A,C,G,T,N} per record (~0 = none)
    u64* d_tags = nullptr;        // unordered: TW words per record
    bool finished = false;
};

struct SortScratch {
    u64 *keyA = nullptr, *keyB = nullptr;
    u32 *aA = nullptr, *aB = nullptr, *bA = nullptr, *bB = nullptr;
    u32 *hist = nullptr, *hist_scan = nullptr;
    u64* scan_state = nullptr; u32* ticket = nullptr; u64* d_total = nullptr;
    u64 n_cap = 0; u64 hist_cap = 0; u64 state_cap = 0;
};

// --unordered by stages: fqd_finish runs them back to back on one GPU; across GPUs (tag ranges, sharded_unordered.py) the
// driver runs them one by one and supplies, between them, what only the whole job knows: where the walk stops, where
// this range's pairs stand in the job's emission order, and which of them another range has seen first.
struct UnJoin {
    int stage = 0;              // 1 tags sorted + partners known, 2 emission order + pair keys built, 3 finished
    u32 *permL = nullptr, *permR = nullptr, *matchL = nullptr, *matchR = nullptr, *emit = nullptr, *unL = nullptr, *unR = nullptr, *excl = nullptr;
    JoinStop* d_stop = nullptr;
    SortScratch sc;             // scan fields only
    u64 E = 0, unmatched = 0, hbad = ~0ull; u32 final_equal = 0;
    u64 *pair_rows = nullptr, *h1 = nullptr, *h2 = nullptr, *first_bad = nullptr;
    u32 *idxL = nullptr, *idxR = nullptr, *keep = nullptr;
    u8* dup = nullptr;
    u32* send_idx = nullptr; u64 n_send = 0;      // rows handed to the owners of their hash range, in sending order
};

struct SeqState {
    fqd_config cfg;
    cudaStream_t stream = nullptr;
    int sm = 148;
    u32 W = 0, mates = 1, row_words = 0, TW = 4;
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
    u64* d_o_off[2] = {nullptr, nullptr};      // device copies of the emission (off, len) lists
    u32* d_o_len[2] = {nullptr, nullptr};
    u64 emit_cursor[2] = {0, 0};
    u64* d_seg_base[2] = {nullptr, nullptr};   // per mate: logical base of every segment (+ sentinel)
    u8** d_seg_ptr[2] = {nullptr, nullptr};
    u32 n_segs[2] = {0, 0};
    u8* d_stage = nullptr; size_t stage_cap = 0;
    u32* d_dst = nullptr; size_t dst_cap = 0;
    u64* em_scan_state = nullptr; u32* em_ticket = nullptr; u64* em_total = nullptr;
    // stages of fqd_finish (the multi-GPU driver runs them one by one, with the boundary exchange in between)
    bool parsed = false, scanned = false;
    u32 *d_perm = nullptr, *d_keep = nullptr, *d_brk = nullptr, *d_tail_head = nullptr;
    u64 *d_bound_prev = nullptr, *d_bound_out = nullptr;
    bool has_prev = false;
    // repartition by key range
    u32* d_part_perm = nullptr; u64* d_part_starts = nullptr; u64* d_part_off[2] = {nullptr, nullptr};
    u32 part_G = 0; u64 part_n = 0;
    u32* d_part_perm1 = nullptr; u64 part_n1 = 0;      // --unordered: the second file is partitioned by its own tags
    UnJoin un;
    u32* d_cl_len[2] = {nullptr, nullptr};     // --write-clusters: line length per sorted position
    u64 cl_cursor[2] = {0, 0};
    bool low_bytes = false;          // byte keys: a sequence holds a byte below '\n' (see k_scan_loose_literal)
    bool discard = false;            // fqd_discard_input: a segment's raw bytes are freed as soon as it is split and packed
    u64* d_clr_off = nullptr; u32* d_clr_len = nullptr; u8* d_clr_head = nullptr; u64 clr_cap = 0;    // fqd_cluster_read window
    bool lists_on_host = false;      // h_off / h_len filled (only fqd_emission needs them)
    u32* h_lenwin = nullptr;         // pinned window of record lengths for fqd_emit's batch cuts
    std::vector<void*> scratch;      // freed at destroy / reset
    u64 launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double ms = 0.0;
};

// Transient device memory (input segments, sort / scan / emission scratch) comes from the device's stream-ordered
// memory pool: cudaMalloc costs ~0.1 ms per MiB on a busy B200 (170 ms of a 400 ms job at 20 M pairs), pool
// blocks are reused by the next job of the same handle without any driver call.
static int seq_pool_init(SeqState* s, std::string* err) {
    cudaMemPool_t pool;
    SEQ_TRY(cudaDeviceGetDefaultMemPool(&pool, s->cfg.device));
    uint64_t keep = ~0ull;
    SEQ_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    return FQD_OK;
}
static void seq_pool_trim(SeqState* s) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, s->cfg.device) == cudaSuccess) { cudaStreamSynchronize(s->stream); cudaMemPoolTrimTo(pool, 0); }
}

static int seq_alloc_tables(SeqState* s, std::string* err) {
    SEQ_TRY(cudaMalloc(&s->d_keys, s->capacity * s->row_words * sizeof(u64)));
    for (u32 m = 0; m < s->mates; ++m) {
        SeqMate& mt = s->mate[m];
        SEQ_TRY(cudaMalloc(&mt.d_run, sizeof(RunState)));
        SEQ_TRY(cudaMemsetAsync(mt.d_run, 0, sizeof(RunState), s->stream));
        SEQ_TRY(cudaMalloc(&mt.d_rec_off, s->capacity * sizeof(u64)));
        SEQ_TRY(cudaMalloc(&mt.d_rec_len, s->capacity * sizeof(u32)));
        SEQ_TRY(cudaMalloc(&mt.d_seq_len, s->capacity * sizeof(u32)));
        if (s->cfg.unordered) {
            SEQ_TRY(cudaMalloc(&mt.d_hash, s->capacity * sizeof(u64)));
            SEQ_TRY(cudaMalloc(&mt.d_bad, s->capacity * sizeof(u32)));
            SEQ_TRY(cudaMemsetAsync(mt.d_bad, 0xFF, s->capacity * sizeof(u32), s->stream));
            SEQ_TRY(cudaMalloc(&mt.d_tags, s->capacity * s->TW * sizeof(u64)));
        }
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
    s->TW = (cfg->max_tag_len ? cfg->max_tag_len : 32) / 8 + ((cfg->max_tag_len ? cfg->max_tag_len : 32) % 8 ? 1 : 0);
    memset(&s->stats, 0, sizeof s->stats);
    if (cfg->max_records == 0) { *err = "max_records must be > 0"; delete s; return FQD_ERR_INVALID; }
    s->capacity = cfg->max_records;
    s->seg_bytes = (size_t)cfg->max_chunk_bytes;
    s->chunk_cap = (u32)std::min<u64>(cfg->max_chunk_records ? cfg->max_chunk_records : std::max<u64>(s->seg_bytes / 16, 1024), s->capacity);
    cudaEventCreate(&s->ev0); cudaEventCreate(&s->ev1);
    int rc = seq_pool_init(s, err);
    if (!rc) rc = seq_alloc_tables(s, err);
    if (rc) { *out = s; return rc; }
    *out = s;
    return FQD_OK;
}

static void seq_free_results(SeqState* s) {
    for (void* p : s->scratch) cudaFreeAsync(p, s->stream);
    s->scratch.clear();
    for (int m = 0; m < 2; ++m) { s->h_off[m].clear(); s->h_len[m].clear(); }
    s->lists_on_host = false;
}

static void seq_destroy(SeqState* s) {
    if (!s) return;
    seq_free_results(s);
    for (int m = 0; m < 2; ++m) for (auto& sg : s->mate[m].segs) if (sg.owned && sg.d) cudaFreeAsync(sg.d, s->stream);
    seq_pool_trim(s);
    if (s->h_lenwin) cudaFreeHost(s->h_lenwin);
    for (int m = 0; m < 2; ++m) {
        cudaFree(s->mate[m].d_run); cudaFree(s->mate[m].d_rec_off); cudaFree(s->mate[m].d_rec_len); cudaFree(s->mate[m].d_seq_len);
        cudaFree(s->mate[m].d_hash); cudaFree(s->mate[m].d_bad); cudaFree(s->mate[m].d_tags);
    }
    cudaFree(s->d_keys); cudaFree(s->d_tile_state); cudaFree(s->d_ctl); cudaFree(s->d_rec_start); cudaFree(s->d_hash);
    if (s->h_ctl) cudaFreeHost(s->h_ctl);
    if (s->ev0) { cudaEventDestroy(s->ev0); cudaEventDestroy(s->ev1); }
    delete s;
}

static int seq_reset(SeqState* s, std::string* err) {
    seq_free_results(s);
    for (u32 m = 0; m < s->mates; ++m) {
        for (auto& sg : s->mate[m].segs) if (sg.owned && sg.d) cudaFreeAsync(sg.d, s->stream);
        s->mate[m].segs.clear(); s->mate[m].adopted = false;
        s->mate[m].n_records = 0; s->mate[m].finished = false;
        SEQ_TRY(cudaMemsetAsync(s->mate[m].d_run, 0, sizeof(RunState), s->stream));
        if (s->mate[m].d_bad) SEQ_TRY(cudaMemsetAsync(s->mate[m].d_bad, 0xFF, s->capacity * sizeof(u32), s->stream));
    }
    s->n = s->n_out = 0; s->finished = false; s->ms = 0;
    s->parsed = s->scanned = s->has_prev = false; s->low_bytes = false;
    s->d_cl_len[0] = s->d_cl_len[1] = nullptr; s->cl_cursor[0] = s->cl_cursor[1] = 0;
    s->d_perm = s->d_keep = s->d_brk = s->d_tail_head = nullptr; s->d_bound_prev = s->d_bound_out = nullptr;
    s->d_part_perm = nullptr; s->d_part_starts = nullptr; s->d_part_off[0] = s->d_part_off[1] = nullptr; s->part_G = 0; s->part_n = 0;
    s->d_part_perm1 = nullptr; s->part_n1 = 0;
    s->un = UnJoin();
    s->emit_cursor[0] = s->emit_cursor[1] = 0;
    for (int m = 0; m < 2; ++m) { s->d_o_off[m] = nullptr; s->d_o_len[m] = nullptr; s->d_seg_base[m] = nullptr; s->d_seg_ptr[m] = nullptr; }
    s->d_stage = nullptr; s->stage_cap = 0; s->d_dst = nullptr; s->dst_cap = 0; s->em_scan_state = nullptr;
    s->d_clr_off = nullptr; s->d_clr_len = nullptr; s->d_clr_head = nullptr; s->clr_cap = 0;
    memset(&s->stats, 0, sizeof s->stats);
    return FQD_OK;
}

// FQD_TRACE=1: wall-clock checkpoints on stderr (each one synchronises the stream; for finding host-side stalls)
struct SeqTrace {
    bool on; std::chrono::steady_clock::time_point t0;
    SeqTrace() : on(getenv("FQD_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(cudaStream_t st, const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[fqd trace] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

static void seq_set_error(SeqState* s, int code, int ch, u64 rec, int mate) {
    if (s->stats.err) return;
    s->stats.err = code; s->stats.err_char = ch; s->stats.err_record = rec; s->stats.err_mate = mate;
}

// The record tables and the key rows are full: make them larger in place (the host sized them from the file size, which
// a pipe does not have; the reference's vectors simply grow, src/external_sort.hpp:88-117).  A parse that stops at the
// tables' end leaves the rest of its segment as the carried tail, so parsing just continues after this.
template <class T>
static cudaError_t seq_regrow(T** p, u64 old_count, u64 new_count, cudaStream_t st, int fill = -1) {
    if (!*p) return cudaSuccess;
    T* q = nullptr;
    cudaError_t e = cudaMalloc(&q, new_count * sizeof(T));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(q, *p, old_count * sizeof(T), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess && fill >= 0) e = cudaMemsetAsync(q + old_count, fill, (new_count - old_count) * sizeof(T), st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaFree(q); return e; }
    cudaFree(*p);
    *p = q;
    return cudaSuccess;
}
static bool seq_grow(SeqState* s) {
    const u64 old = s->capacity;
    for (u64 add = std::max<u64>(old, 4096); add >= 1024; add /= 2) {      // double; with less head room when memory is short
        const u64 cap = old + add;
        // all or nothing per attempt: a failed allocation leaves the pointers that were not reached untouched, and the
        // ones that were already moved are simply larger than they need to be
        size_t free_b = 0, total_b = 0;
        const size_t per_rec = s->row_words * 8 + s->mates * (8 + 4 + 4 + (s->cfg.unordered ? 8 + 4 + s->TW * 8 : 0));
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || (double)cap * per_rec * 1.02 > (double)free_b) continue;
        bool ok = seq_regrow(&s->d_keys, old * s->row_words, cap * s->row_words, s->stream) == cudaSuccess;
        for (u32 m = 0; ok && m < s->mates; ++m) {
            SeqMate& mt = s->mate[m];
            ok = seq_regrow(&mt.d_rec_off, old, cap, s->stream) == cudaSuccess && seq_regrow(&mt.d_rec_len, old, cap, s->stream) == cudaSuccess &&
                 seq_regrow(&mt.d_seq_len, old, cap, s->stream) == cudaSuccess && seq_regrow(&mt.d_hash, old, cap, s->stream) == cudaSuccess &&
                 seq_regrow(&mt.d_bad, old, cap, s->stream, 0xFF) == cudaSuccess && seq_regrow(&mt.d_tags, old * s->TW, cap * s->TW, s->stream) == cudaSuccess;
        }
        if (!ok) { cudaGetLastError(); return false; }
        s->capacity = cap;
        return true;
    }
    return false;
}

// Parse the current (last) segment of a mate; move its incomplete tail into a fresh segment unless `final`.
static int seq_parse_segment(SeqState* s, int m, bool final, std::string* err) {
    SeqMate& mt = s->mate[m];
    SeqSegment& sg = mt.segs.back();
    if (sg.fill == 0 || !sg.d) return FQD_OK;
    const u32 n_tiles = (u32)((sg.fill + PP_TILE - 1) / PP_TILE);
    if (s->capacity == mt.n_records && !seq_grow(s)) { seq_set_error(s, FQD_ERR_CAPACITY, 0, mt.n_records, m); return FQD_OK; }
    const u64 room = s->capacity - mt.n_records;
    SEQ_TRY(cudaEventRecord(s->ev0, s->stream));
    k_init_chunk<<<std::max(1u, std::min(n_tiles / 256 + 1, 1024u)), 256, 0, s->stream>>>(s->d_ctl, s->d_tile_state, n_tiles);
    ParseParams p;
    p.raw = sg.d; p.n = (u32)sg.fill; p.n_tiles = n_tiles; p.tile_state = s->d_tile_state; p.ctl = s->d_ctl; p.run = mt.d_run;
    p.rec_start = s->d_rec_start; p.cap = (u32)std::min<u64>(s->chunk_cap, room); p.keys = s->d_keys; p.key_capacity = s->capacity;
    p.row_words = s->row_words; p.mate_off = m * s->W; p.W = s->W; p.hash = s->d_hash; p.seq_len = mt.d_seq_len + mt.n_records;
    p.word0 = nullptr; p.dup = nullptr; p.strict = 0; p.hash_salt = m * 4096u; p.bad_rec = nullptr;
    p.byte_keys = s->cfg.byte_keys ? 1u : 0u; p.skip = sg.skip;
    if (s->cfg.unordered) { p.hash = mt.d_hash + mt.n_records; p.bad_rec = mt.d_bad + mt.n_records; }
    pp_launch(s->cfg.format == FQD_FORMAT_FASTQ, p, s->stream);
    s->launches += 2;               // the two head kernels of pp_launch
    if (s->cfg.unordered) {
        k_extract_tags<<<s->sm * 8, 128, 0, s->stream>>>(sg.d, s->d_rec_start, s->d_ctl, s->TW, mt.d_tags + mt.n_records * s->TW, s->d_ctl);
        s->launches++;
    }
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
    if (c.too_long & TL_SEQ) seq_set_error(s, FQD_ERR_SEQ_TOO_LONG, 0, first, m);
    // TL_CAPACITY: the parse stopped at the tables' end (p.cap); what it left is the carried tail, the next call grows
    if (c.too_long & TL_TAG) seq_set_error(s, FQD_ERR_TAG_TOO_LONG, 0, first, m);
    if (c.pad == 2 && s->cfg.byte_keys) s->low_bytes = true;
    else if (c.pad && !s->cfg.unordered) seq_set_error(s, FQD_ERR_UNSUPPORTED_BYTE, 0, first, m);
    mt.n_records += c.n_records;
    const size_t consumed = c.consumed, tail = sg.fill - consumed;
    if (c.n_records == 0 && !final && sg.owned && sg.fill >= sg.cap) {
        *err = "a single record does not fit into one device segment (raise max_chunk_bytes)";
        return FQD_ERR_CAPACITY;
    }
    if (!final && !sg.owned) {
        mt.segs.back().fill = consumed;          // the next view (seq_adopt) starts where this one stopped
    } else if (!final) {
        SeqSegment nx;
        nx.cap = s->seg_bytes;
        SEQ_TRY(cudaMallocAsync(&nx.d, nx.cap + 4096, s->stream));
        nx.logical_base = sg.logical_base + consumed;
        if (tail) SEQ_TRY(cudaMemcpyAsync(nx.d, sg.d + consumed, tail, cudaMemcpyDeviceToDevice, s->stream));
        nx.fill = tail;
        mt.segs.back().fill = consumed;
        if (s->discard) {               // stream-ordered: after the parse kernels and the copy of the tail
            SEQ_TRY(cudaFreeAsync(mt.segs.back().d, s->stream));
            mt.segs.back().d = nullptr;
        }
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
        if (s->discard && sg.owned && sg.d) {
            SEQ_TRY(cudaFreeAsync(mt.segs.back().d, s->stream));
            mt.segs.back().d = nullptr;
        }
    }
    return FQD_OK;
}

// The whole input of one mate is a device buffer of the caller: no copy, the engine parses it in place through views of
// at most seg_bytes that are cut at record boundaries.  One call per mate, 16-byte aligned, instead of fqd_append*.
static int seq_adopt(SeqState* s, int m, const void* d_buf, size_t n, std::string* err) {
    if (m < 0 || (u32)m >= s->mates) { *err = "bad mate index"; return FQD_ERR_INVALID; }
    if (s->finished || s->parsed) { *err = "fqd_adopt_device after fqd_finish"; return FQD_ERR_INVALID; }
    SeqMate& mt = s->mate[m];
    if (!mt.segs.empty()) { *err = "fqd_adopt_device: the mate already has input"; return FQD_ERR_INVALID; }
    if ((uintptr_t)d_buf & 15u) { *err = "fqd_adopt_device: the buffer must be 16-byte aligned"; return FQD_ERR_INVALID; }
    mt.adopted = true;
    u8* base = (u8*)d_buf;
    size_t off = 0;
    while (off < n && !s->stats.err) {
        u8* addr = base + off;
        const u32 skip = (u32)((uintptr_t)addr & 15u);
        const size_t remaining = n - off;
        SeqSegment sg;
        sg.owned = false; sg.skip = skip; sg.d = addr - skip;
        sg.fill = sg.cap = std::min(s->seg_bytes, remaining) + skip;
        sg.logical_base = off - skip;
        mt.segs.push_back(sg);
        int rc = seq_parse_segment(s, m, false, err);
        if (rc) return rc;
        const size_t consumed = mt.segs.back().fill;          // relative to sg.d; the view now ends at a record boundary
        if (consumed <= skip) {                               // no complete record in this view
            mt.segs.pop_back();
            if (remaining > s->seg_bytes) { *err = "a single record does not fit into one device segment (raise max_chunk_bytes)"; return FQD_ERR_CAPACITY; }
            // an incomplete last record is dropped silently (src/fastqview.cpp:114-115), but its first byte is still
            // checked when the record before it is fetched (src/fastqview.cpp:91-92)
            u8 b = 0;
            SEQ_TRY(cudaMemcpy(&b, addr, 1, cudaMemcpyDeviceToHost));
            const u8 lead = s->cfg.format == FQD_FORMAT_FASTQ ? '@' : '>';
            if (b != lead && !s->stats.err) seq_set_error(s, FQD_ERR_BAD_START, b, mt.n_records, m);
            break;
        }
        off += consumed - skip;
    }
    if (mt.segs.empty()) {            // nothing but an incomplete record: keep an empty view so that the mate is not "absent"
        SeqSegment sg; sg.owned = false; sg.d = base; sg.fill = sg.cap = 0; mt.segs.push_back(sg);
    }
    return FQD_OK;
}

static int seq_append(SeqState* s, int m, const void* buf, size_t n, bool is_device, std::string* err) {
    if (m < 0 || (u32)m >= s->mates) { *err = "bad mate index"; return FQD_ERR_INVALID; }
    if (s->finished) { *err = "fqd_append after fqd_finish"; return FQD_ERR_INVALID; }
    SeqMate& mt = s->mate[m];
    if (mt.adopted) { *err = "fqd_append after fqd_adopt_device"; return FQD_ERR_INVALID; }
    const u8* src = (const u8*)buf;
    while (n) {
        if (s->stats.err) return FQD_OK;        // a data error is sticky: what follows it is never looked at
        if (mt.segs.empty()) {
            SeqSegment sg; sg.cap = s->seg_bytes;
            SEQ_TRY(cudaMallocAsync(&sg.d, sg.cap + 4096, s->stream));
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
template <class T>
static int seq_dalloc(SeqState* s, T** p, size_t count, std::string* err) {
    void* q = nullptr;
    SEQ_TRY(cudaMallocAsync(&q, std::max<size_t>(count, 1) * sizeof(T), s->stream));
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
    // the scatter reorders a tile in shared memory: more than the 48 KB a kernel gets without asking
    SEQ_TRY(cudaFuncSetAttribute(k_radix_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(true)));
    SEQ_TRY(cudaFuncSetAttribute(k_radix_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(false)));
    // digits that are the same in every key need no pass (sc.d_total doubles as the two-word result)
    u64 varying = ~0ull;
    if (n >= (1u << 16)) {
        u64 h_oa[2] = {0ull, ~0ull};
        u64* d_oa = nullptr;
        SEQ_TRY(cudaMallocAsync((void**)&d_oa, 2 * sizeof(u64), s->stream));
        SEQ_TRY(cudaMemcpyAsync(d_oa, h_oa, sizeof h_oa, cudaMemcpyHostToDevice, s->stream));
        k_key_or_and<<<s->sm * 4, 256, 0, s->stream>>>(sc.keyA, n, d_oa);
        SEQ_TRY(cudaMemcpyAsync(h_oa, d_oa, sizeof h_oa, cudaMemcpyDeviceToHost, s->stream));
        SEQ_TRY(cudaStreamSynchronize(s->stream));
        SEQ_TRY(cudaFreeAsync(d_oa, s->stream));
        varying = h_oa[0] ^ h_oa[1];
        s->launches++;
    }
    for (u32 shift = bit_lo; shift < bit_hi; shift += 8) {
        if (((varying >> shift) & 0xFFull) == 0) continue;
        k_radix_hist<<<nblocks, RS_THREADS, 0, s->stream>>>(sc.keyA, n, shift, sc.hist, nblocks);
        int rc = device_scan(s, sc, sc.hist, sc.hist_scan, 256ull * nblocks, nullptr, err);
        if (rc) return rc;
        if (has_b) k_radix_scatter<true><<<nblocks, RS_THREADS, rs_smem_bytes(true), s->stream>>>(sc.keyA, sc.aA, sc.bA, sc.keyB, sc.aB, sc.bB, n, shift, sc.hist_scan, nblocks);
        else k_radix_scatter<false><<<nblocks, RS_THREADS, rs_smem_bytes(false), s->stream>>>(sc.keyA, sc.aA, nullptr, sc.keyB, sc.aB, nullptr, n, shift, sc.hist_scan, nblocks);
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
    SeqTrace tr;
    int rc = sort_scratch_alloc(s, sc, n, err);
    if (rc) return rc;
    u32 *head, *excl, *gid, *gsize, *gdiff, *flag, *posA, *posB;
    if ((rc = seq_dalloc(s, &head, n, err)) || (rc = seq_dalloc(s, &excl, n, err)) || (rc = seq_dalloc(s, &gid, n, err)) ||
        (rc = seq_dalloc(s, &gsize, n, err)) || (rc = seq_dalloc(s, &gdiff, n, err)) || (rc = seq_dalloc(s, &flag, n, err)) ||
        (rc = seq_dalloc(s, &posA, n, err)) || (rc = seq_dalloc(s, &posB, n, err))) return rc;
    MidUnit* mid = nullptr; MidCtl* n_mid = nullptr;
    if ((rc = seq_dalloc(s, &mid, MID_LIST_CAP, err)) || (rc = seq_dalloc(s, &n_mid, 1, err))) return rc;
    const bool fine = getenv("FQD_TRACE_SORT") != nullptr;
    const u32 hi_bit = (word_bits + 7) / 8 * 8;
    tr.mark(s->stream, "  sort: scratch alloc");

    // round 0: all records by their first word
    k_iota_u32<<<seq_grid(s, n), 256, 0, s->stream>>>(sc.aA, n);
    k_gather_word<<<seq_grid(s, n), 256, 0, s->stream>>>(rows + w_begin, stride, 0, sc.aA, n, sc.keyA);
    s->launches += 2;
    if ((rc = radix_sort(s, sc, n, 0, hi_bit, false, err))) return rc;
    SEQ_TRY(cudaMemcpyAsync(perm, sc.aA, n * sizeof(u32), cudaMemcpyDeviceToDevice, s->stream));
    tr.mark(s->stream, "  sort: round 0");

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
        if (fine) tr.mark(s->stream, "    round: heads, groups");
        k_mark_unresolved<<<seq_grid(s, n_act), 256, 0, s->stream>>>(rows + w_begin, stride, w + 1, n_words, sc.aA, gid, n_act, gdiff);
        if (fine) tr.mark(s->stream, "    round: mark unresolved");
        if (!getenv("FQD_SORT_NO_SMALL_GROUPS")) {
            SEQ_TRY(cudaMemsetAsync(n_mid, 0, sizeof(MidCtl), s->stream));
            // `flag` is free until k_active_flags: it holds the list of small groups (at most n_act / 2 of them)
            k_collect_groups<<<seq_grid(s, n_act), 256, 0, s->stream>>>(head, gid, gsize, gdiff, n_act, flag, mid, n_mid);
            if (fine) tr.mark(s->stream, "    round: collect groups");
            k_sort_small_groups<<<seq_grid(s, n_act / 2 + 1), 256, 0, s->stream>>>(rows + w_begin, stride, w + 1, n_words, gid, gsize, gdiff, sc.aA, pos_in, perm, flag, n_mid);
            if (fine) tr.mark(s->stream, "    round: small groups");
            k_sort_mid_groups<<<s->sm * 8, 256, 0, s->stream>>>(rows + w_begin, stride, w + 1, n_words, gid, gdiff, sc.aA, pos_in, perm, mid, n_mid);
            s->launches += 3;
        }
        if (fine) tr.mark(s->stream, "    round: mid groups");
        k_active_flags<<<seq_grid(s, n_act), 256, 0, s->stream>>>(gid, gsize, gdiff, n_act, flag);
        u64 n_next = 0;
        if ((rc = device_scan(s, sc, flag, excl, n_act, &n_next, err))) return rc;
        s->launches += 2;
        if (fine) tr.mark(s->stream, "    round: active scan");
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
        if (fine) tr.mark(s->stream, "    round: sort of the unresolved");
    }
    tr.mark(s->stream, "  sort: refine rounds");
    SEQ_TRY(cudaGetLastError());
    return FQD_OK;
}

// ---------------------------------------------------------------------------------------------------------
static ScanParams seq_scan_params(SeqState* s) {
    ScanParams sp;
    sp.rows = s->d_keys; sp.stride = s->row_words; sp.W = s->W; sp.mates = s->mates;
    sp.len0 = s->mate[0].d_seq_len; sp.len1 = s->mates == 2 ? s->mate[1].d_seq_len : nullptr;
    sp.perm = s->d_perm; sp.n = s->n; sp.dist = s->cfg.hamming_dist; sp.keep = s->d_keep; sp.brk = s->d_brk;
    sp.bits = s->cfg.byte_keys ? 8u : 3u;
    sp.tail_head = s->d_tail_head;
    return sp;
}

// stage 1 of fqd_finish: sort + comparator scan -> keep flags in sorted order (as if nothing preceded this input)
static int seq_scan_stage(SeqState* s, std::string* err) {
    const u64 n = s->n;
    int rc;
    SeqTrace tr;
    if ((rc = seq_dalloc(s, &s->d_perm, n, err))) return rc;
    if ((rc = sort_rows(s, s->d_keys, s->row_words, 0, s->row_words, s->cfg.byte_keys ? 64 : 60, n, s->d_perm, err))) return rc;
    tr.mark(s->stream, "sort_rows");
    if ((rc = seq_dalloc(s, &s->d_keep, n, err)) || (rc = seq_dalloc(s, &s->d_tail_head, 1, err)) ||
        (rc = seq_dalloc(s, &s->d_bound_prev, boundary_words(s->row_words), err)) ||
        (rc = seq_dalloc(s, &s->d_bound_out, boundary_words(s->row_words), err))) return rc;
    SEQ_TRY(cudaMemsetAsync(s->d_bound_prev, 0, boundary_words(s->row_words) * sizeof(u64), s->stream));
    if (s->cfg.mode == FQD_MODE_SEQ_HAMMING && (rc = seq_dalloc(s, &s->d_brk, n, err))) return rc;
    const ScanParams sp = seq_scan_params(s);
    // rows of 2^k 16-byte pieces (150 bp: 4 per mate): the warp-cooperative scan, every row loaded once
    const u32 Q = s->row_words / 2u, Qm = s->W / 2u;
    const bool coop = Q >= 1 && Q <= 32 && (Q & (Q - 1)) == 0 && Qm >= 1 && (Qm & (Qm - 1)) == 0 && !getenv("FQD_SCAN_PER_THREAD");
    const unsigned coop_blocks = (unsigned)(((n + COOP_RUN - 1) / COOP_RUN + 7) / 8);
    if (s->cfg.mode == FQD_MODE_SEQ_TIGHT && coop) k_scan_coop<1><<<coop_blocks, 256, 0, s->stream>>>(sp, Q);
    else if (s->cfg.mode == FQD_MODE_SEQ_TIGHT) k_scan_tight<<<seq_grid(s, n), 256, 0, s->stream>>>(sp);
    else if (s->cfg.mode == FQD_MODE_SEQ_LOOSE && s->low_bytes) k_scan_loose_literal<<<1, 1, 0, s->stream>>>(sp);
    else if (s->cfg.mode == FQD_MODE_SEQ_LOOSE && coop) k_scan_coop<2><<<coop_blocks, 256, 0, s->stream>>>(sp, Q);
    else if (s->cfg.mode == FQD_MODE_SEQ_LOOSE) k_scan_loose<<<seq_grid(s, n), 256, 0, s->stream>>>(sp);
    else {
        HamLong* d_long = nullptr; u32* d_nlong = nullptr;
        if ((rc = seq_dalloc(s, &d_long, n / HAM_SHORT + 2, err)) || (rc = seq_dalloc(s, &d_nlong, 1, err))) return rc;
        SEQ_TRY(cudaMemsetAsync(d_nlong, 0, sizeof(u32), s->stream));
        if (coop) k_scan_coop<3><<<coop_blocks, 256, 0, s->stream>>>(sp, Q);
        else k_ham_breaks<<<seq_grid(s, n), 256, 0, s->stream>>>(sp);
        k_ham_segments<<<seq_grid(s, n), 256, 0, s->stream>>>(sp, d_long, d_nlong);
        k_ham_long<<<s->sm * 4, 256, 0, s->stream>>>(sp, d_long, d_nlong);
        s->launches += 2;
    }
    s->launches++;
    s->scanned = true;
    tr.mark(s->stream, "scan");
    return FQD_OK;
}

// The boundary state after this input's last sorted record (host buffer of boundary_words(row_words) words).
static int seq_boundary_get(SeqState* s, u64* out, std::string* err) {
    if (!s->scanned) { *err = "boundary state before the scan stage"; return FQD_ERR_INVALID; }
    if (s->low_bytes && s->cfg.mode == FQD_MODE_SEQ_LOOSE) {
        *err = "loose mode on sequences with bytes below the line feed cannot be split into key ranges";
        return FQD_ERR_INVALID;
    }
    const ScanParams sp = seq_scan_params(s);
    k_tail_state<<<1, 64, 0, s->stream>>>(sp, s->d_bound_prev, s->cfg.mode == FQD_MODE_SEQ_HAMMING ? 1 : 0, s->d_bound_out);
    s->launches++;
    SEQ_TRY(cudaMemcpyAsync(out, s->d_bound_out, boundary_words(s->row_words) * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    return FQD_OK;
}
// Re-evaluate the first sorted records against the boundary state of the key range that precedes this input.
static int seq_boundary_fix(SeqState* s, const u64* prev, std::string* err) {
    if (!s->scanned || s->finished) { *err = "boundary fix outside the scan stage"; return FQD_ERR_INVALID; }
    SEQ_TRY(cudaMemcpyAsync(s->d_bound_prev, prev, boundary_words(s->row_words) * sizeof(u64), cudaMemcpyHostToDevice, s->stream));
    const ScanParams sp = seq_scan_params(s);
    if (s->cfg.mode == FQD_MODE_SEQ_TIGHT) k_fix_first_tight<<<1, 1, 0, s->stream>>>(sp, s->d_bound_prev);
    else if (s->cfg.mode == FQD_MODE_SEQ_LOOSE) k_fix_first_loose<<<1, 1, 0, s->stream>>>(sp, s->d_bound_prev);
    else k_fix_first_ham<<<1, 1, 0, s->stream>>>(sp, s->d_bound_prev);
    s->launches++;
    SEQ_TRY(cudaStreamSynchronize(s->stream));       // prev is the caller's buffer
    s->has_prev = true;
    return FQD_OK;
}

// stage 2 of fqd_finish: survivors -> emission lists
static int seq_emit_stage(SeqState* s, std::string* err) {
    const u64 n = s->n;
    int rc;
    SeqTrace tr;
    u32* excl;
    if ((rc = seq_dalloc(s, &excl, n, err))) return rc;
    u32 *keep = s->d_keep, *perm = s->d_perm;
    SortScratch sc;      // only the scan scratch is used here
    if ((rc = seq_dalloc(s, &sc.scan_state, (n + SCAN_TILE - 1) / SCAN_TILE + 16, err)) || (rc = seq_dalloc(s, &sc.ticket, 4, err)) ||
        (rc = seq_dalloc(s, &sc.d_total, 2, err))) return rc;
    u64 n_out = 0;
    if ((rc = device_scan(s, sc, keep, excl, n, &n_out, err))) return rc;
    tr.mark(s->stream, "survivor count");
    s->n_out = n_out;
    u64 *o_off[2] = {nullptr, nullptr}; u32 *o_len[2] = {nullptr, nullptr}; u32* o_idx;
    if ((rc = seq_dalloc(s, &o_idx, n_out, err))) return rc;
    for (u32 m = 0; m < s->mates; ++m)
        if ((rc = seq_dalloc(s, &o_off[m], n_out, err)) || (rc = seq_dalloc(s, &o_len[m], n_out, err))) return rc;
    k_emit<<<seq_grid(s, n), 256, 0, s->stream>>>(keep, excl, perm, n, s->mate[0].d_rec_off, s->mate[0].d_rec_len,
                                                  s->mates == 2 ? s->mate[1].d_rec_off : nullptr, s->mates == 2 ? s->mate[1].d_rec_len : nullptr,
                                                  o_off[0], o_len[0], o_off[1], o_len[1], o_idx);
    s->launches++;
    for (u32 m = 0; m < s->mates; ++m) { s->d_o_off[m] = o_off[m]; s->d_o_len[m] = o_len[m]; }
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    tr.mark(s->stream, "emit lists");
    s->stats.total = n;
    s->stats.dups = n - n_out;
    return FQD_OK;
}

static int seq_finish_unordered(SeqState* s, std::string* err);

// everything appended so far is parsed; s->n = records (pairs) that take part
static int seq_parse_rest(SeqState* s, std::string* err, bool allow_empty = false) {
    if (s->parsed) return FQD_OK;
    s->parsed = true;
    SeqTrace tr;
    for (u32 m = 0; m < s->mates; ++m) {
        SeqMate& mt = s->mate[m];
        if (mt.segs.empty()) { if (!allow_empty) seq_set_error(s, FQD_ERR_EMPTY, 0, 0, m); continue; }
        if (mt.adopted) continue;                 // fqd_adopt_device parsed everything in place
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
    tr.mark(s->stream, "finish: last segments");
    if (s->stats.err) return FQD_OK;            // data error: reported through fqd_stats, like the fast mode
    u64 n = s->mate[0].n_records;
    if (s->mates == 2 && !s->cfg.unordered) n = std::min(n, s->mate[1].n_records);      // stops at the shorter file
    if (n == 0 || (s->mates == 2 && s->mate[1].n_records == 0)) {
        if (!allow_empty) seq_set_error(s, FQD_ERR_EMPTY, 0, 0, 0);
        if (!allow_empty || !s->cfg.unordered) return FQD_OK;
    }
    s->n = n;
    return FQD_OK;
}

// fqd_finish_scan / fqd_finish_emit: the two halves of fqd_finish for sequence-based modes
static int seq_finish_scan(SeqState* s, std::string* err) {
    if (s->finished || s->scanned) { *err = "fqd_finish_scan called twice"; return FQD_ERR_INVALID; }
    if (s->cfg.unordered) { *err = "staged finish is for sequence-based modes"; return FQD_ERR_INVALID; }
    int rc = seq_parse_rest(s, err);
    if (rc || s->stats.err) return rc;
    SEQ_TRY(cudaEventRecord(s->ev0, s->stream));
    if ((rc = seq_scan_stage(s, err))) return rc;
    SEQ_TRY(cudaEventRecord(s->ev1, s->stream));
    SEQ_TRY(cudaEventSynchronize(s->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); s->ms += ms;
    return FQD_OK;
}
static int seq_finish_emit(SeqState* s, std::string* err) {
    if (s->finished) { *err = "fqd_finish called twice"; return FQD_ERR_INVALID; }
    s->finished = true;
    if (s->stats.err) return FQD_OK;
    if (!s->scanned) { *err = "fqd_finish_emit before fqd_finish_scan"; return FQD_ERR_INVALID; }
    SEQ_TRY(cudaEventRecord(s->ev0, s->stream));
    int rc = seq_emit_stage(s, err);
    if (rc) return rc;
    SEQ_TRY(cudaEventRecord(s->ev1, s->stream));
    SEQ_TRY(cudaEventSynchronize(s->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); s->ms += ms;
    return FQD_OK;
}

static int seq_finish_unordered(SeqState* s, std::string* err);

static int seq_finish(SeqState* s, std::string* err) {
    if (s->finished) { *err = "fqd_finish called twice"; return FQD_ERR_INVALID; }
    if (!s->cfg.unordered) {
        int rc = s->scanned ? FQD_OK : seq_finish_scan(s, err);
        if (rc) return rc;
        return seq_finish_emit(s, err);
    }
    s->finished = true;
    int rc = seq_parse_rest(s, err);
    if (rc || s->stats.err) return rc;
    SEQ_TRY(cudaEventRecord(s->ev0, s->stream));
    if ((rc = seq_finish_unordered(s, err))) return rc;
    SEQ_TRY(cudaEventRecord(s->ev1, s->stream));
    SEQ_TRY(cudaEventSynchronize(s->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); s->ms += ms;
    return FQD_OK;
}

static JoinParams un_join_params(SeqState* s) {
    UnJoin& u = s->un;
    JoinParams jp;
    jp.tagsL = s->mate[0].d_tags; jp.permL = u.permL; jp.n = (u32)s->mate[0].n_records;
    jp.tagsR = s->mate[1].d_tags; jp.permR = u.permR; jp.m = (u32)s->mate[1].n_records;
    jp.TW = s->TW; jp.matchL = u.matchL; jp.matchR = u.matchR;
    return jp;
}

// stage 1: ExternalSorter<*ViewWithId> on each file (src/hash_dup_remover.hpp:163-173), stable on the input index, and
// the partner of every record (k-th record of a tag in L with the k-th of the same tag in R)
static int un_prepare(SeqState* s, std::string* err) {
    UnJoin& u = s->un;
    const u64 n = s->mate[0].n_records, m = s->mate[1].n_records;
    const u64 big = std::max(n, m);
    int rc;
    if ((rc = seq_dalloc(s, &u.permL, n, err)) || (rc = seq_dalloc(s, &u.permR, m, err))) return rc;
    if (n && (rc = sort_rows(s, s->mate[0].d_tags, s->TW, 0, s->TW, 64, n, u.permL, err))) return rc;
    if (m && (rc = sort_rows(s, s->mate[1].d_tags, s->TW, 0, s->TW, 64, m, u.permR, err))) return rc;
    if ((rc = seq_dalloc(s, &u.matchL, n, err)) || (rc = seq_dalloc(s, &u.matchR, m, err)) || (rc = seq_dalloc(s, &u.emit, n, err)) ||
        (rc = seq_dalloc(s, &u.unL, n, err)) || (rc = seq_dalloc(s, &u.unR, m, err)) || (rc = seq_dalloc(s, &u.excl, big, err)) ||
        (rc = seq_dalloc(s, &u.d_stop, 1, err))) return rc;
    if ((rc = seq_dalloc(s, &u.sc.scan_state, (big + SCAN_TILE - 1) / SCAN_TILE + 16, err)) || (rc = seq_dalloc(s, &u.sc.ticket, 4, err)) ||
        (rc = seq_dalloc(s, &u.sc.d_total, 2, err))) return rc;
    if (n) SEQ_TRY(cudaMemsetAsync(u.matchL, 0xFF, n * sizeof(u32), s->stream));
    if (m) SEQ_TRY(cudaMemsetAsync(u.matchR, 0xFF, m * sizeof(u32), s->stream));
    if (n && m) {
        k_join_match<<<seq_grid(s, n), 256, 0, s->stream>>>(un_join_params(s));
        s->launches++;
    }
    SEQ_TRY(cudaGetLastError());
    u.stage = 1;
    return FQD_OK;
}

// stage 2: which pairs the walk emits (given == nullptr: both whole files are here, the stop state is computed; else
// it is this range's share of the job's stop state), their order, their keys
static int un_join(SeqState* s, const JoinStop* given, std::string* err) {
    UnJoin& u = s->un;
    const u64 n = s->mate[0].n_records, m = s->mate[1].n_records;
    const u64 big = std::max(n, m);
    const JoinParams jp = un_join_params(s);
    int rc;
    if (!given) k_join_stop<<<1, 1, 0, s->stream>>>(jp, u.d_stop);
    else {
        SEQ_TRY(cudaMemcpyAsync(u.d_stop, given, sizeof(JoinStop), cudaMemcpyHostToDevice, s->stream));
        k_join_final<<<1, 1, 0, s->stream>>>(jp, u.d_stop);
    }
    if (big) k_join_flags<<<seq_grid(s, big), 256, 0, s->stream>>>(jp, u.d_stop, u.emit, u.unL, u.unR);
    s->launches += 2;
    u64 uL = 0, uR = 0, E = 0;
    if ((rc = device_scan(s, u.sc, u.unL, u.excl, n, &uL, err))) return rc;
    if ((rc = device_scan(s, u.sc, u.unR, u.excl, m, &uR, err))) return rc;
    if ((rc = device_scan(s, u.sc, u.emit, u.excl, n, &E, err))) return rc;
    JoinStop hs;
    SEQ_TRY(cudaMemcpyAsync(&hs, u.d_stop, sizeof hs, cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    u.final_equal = hs.final_equal;
    u.unmatched = uL + uR;
    u.E = E;
    u.hbad = ~0ull;
    u.stage = 2;
    if (E == 0) return FQD_OK;
    const u32 W = s->W;
    if ((rc = seq_dalloc(s, &u.pair_rows, E * 2 * W, err)) || (rc = seq_dalloc(s, &u.h1, E, err)) || (rc = seq_dalloc(s, &u.h2, E, err)) ||
        (rc = seq_dalloc(s, &u.first_bad, 1, err)) || (rc = seq_dalloc(s, &u.idxL, E, err)) || (rc = seq_dalloc(s, &u.idxR, E, err)) ||
        (rc = seq_dalloc(s, &u.keep, E, err)) || (rc = seq_dalloc(s, &u.dup, E + 64, err))) return rc;
    SEQ_TRY(cudaMemsetAsync(u.first_bad, 0xFF, sizeof(u64), s->stream));
    SEQ_TRY(cudaMemsetAsync(u.dup, 0, E + 64, s->stream));
    k_build_pairs<<<seq_grid(s, n), 256, 0, s->stream>>>(jp, u.d_stop, u.emit, u.excl, s->d_keys, s->row_words, W, s->mate[0].d_hash, s->mate[1].d_hash,
                                                         s->mate[0].d_bad, s->mate[1].d_bad, u.pair_rows, u.h1, u.h2, u.idxL, u.idxR, u.first_bad);
    s->launches++;
    SEQ_TRY(cudaMemcpyAsync(&u.hbad, u.first_bad, sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    return FQD_OK;
}

// exact first-occurrence set over `n_rows` pair keys (src/hash_dup_remover.hpp:291-306): the row with the smallest
// position among equal keys stays, the others get their flag set
static int un_insert_rows(SeqState* s, const u64* rows, const u64* h1, const u64* h2, u64 n_rows, u64 hash_mul, u8* flags, std::string* err) {
    u64* table; RunState* run;
    u64 nb = 1024;
    while (nb * 2 < n_rows) nb <<= 1;
    u32 lg = 0; while ((1ull << lg) < nb) ++lg;
    int rc;
    if ((rc = seq_dalloc(s, &table, nb * 4, err)) || (rc = seq_dalloc(s, &run, 1, err))) return rc;
    SEQ_TRY(cudaMemsetAsync(table, 0xFF, nb * 4 * sizeof(u64), s->stream));
    k_set_chunk_pairs<<<1, 1, 0, s->stream>>>(run, (u32)n_rows);
    InsertParams ip;
    ip.table = table; ip.bucket_shift = 64 - lg; ip.bucket_mask = nb - 1; ip.keys = rows; ip.row_words = 2 * s->W; ip.key_capacity = n_rows;
    ip.hash1 = h1; ip.hash2 = h2; ip.ctl1 = nullptr; ip.ctl2 = nullptr; ip.run = run; ip.dup = flags; ip.hash_mul = hash_mul; ip.hash_final = h2 ? 0 : 1;
    insert_launch(ip, s->sm * 8, s->stream);
    s->launches += 2;
    SEQ_TRY(cudaGetLastError());
    return FQD_OK;
}

// stage 3: survivors = emitted pairs before `limit` whose flag is clear; their records, in emission order
static int un_finish(SeqState* s, u64 limit, bool report_bad, std::string* err) {
    UnJoin& u = s->un;
    const u64 E = u.E;
    int rc;
    u.stage = 3;
    s->n = E; s->n_out = 0;
    s->stats.total = limit; s->stats.dups = 0;
    if (report_bad) seq_set_error(s, FQD_ERR_BAD_BASE, (int)(u.hbad & 0xFF), limit, 0);
    if (E == 0) return FQD_OK;
    k_keep_from_dup<<<seq_grid(s, E), 256, 0, s->stream>>>(u.dup, E, limit, u.keep);
    u64 n_out = 0;
    if ((rc = device_scan(s, u.sc, u.keep, u.excl, E, &n_out, err))) return rc;
    s->n_out = n_out;
    s->stats.dups = limit - n_out;
    u64* o_off[2]; u32* o_len[2];
    for (u32 k = 0; k < 2; ++k)
        if ((rc = seq_dalloc(s, &o_off[k], n_out, err)) || (rc = seq_dalloc(s, &o_len[k], n_out, err))) return rc;
    k_emit_pairs<<<seq_grid(s, E), 256, 0, s->stream>>>(u.keep, u.excl, E, u.idxL, u.idxR, s->mate[0].d_rec_off, s->mate[0].d_rec_len,
                                                        s->mate[1].d_rec_off, s->mate[1].d_rec_len, o_off[0], o_len[0], o_off[1], o_len[1]);
    s->launches += 2;
    for (u32 k = 0; k < 2; ++k) { s->d_o_off[k] = o_off[k]; s->d_o_len[k] = o_len[k]; }
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    return FQD_OK;
}

static int seq_finish_unordered(SeqState* s, std::string* err) {
    UnJoin& u = s->un;
    int rc;
    if ((rc = un_prepare(s, err)) || (rc = un_join(s, nullptr, err))) return rc;
    s->stats.unmatched = u.unmatched + (u.final_equal ? 0 : 1);
    s->stats.total = u.E;
    s->n = u.E;
    s->n_out = 0;
    if (u.E == 0) { s->stats.dups = 0; u.stage = 3; return FQD_OK; }
    if ((rc = un_insert_rows(s, u.pair_rows, u.h1, u.h2, u.E, 1, u.dup, err))) return rc;
    // a matched pair holds a byte outside {A,C,G,T,N}: the run aborts when it is keyed
    const bool bad = u.hbad != ~0ull;
    return un_finish(s, bad ? (u.hbad >> 8) : u.E, bad, err);
}

// ---- --unordered across GPUs: the stages one by one (fastq-dupaway_b200/sharded_unordered.py holds the protocol).
// This engine holds the records of both files whose tags fall into ONE tag range; ranges in rank order = tag order.
static int seq_unordered_prepare(SeqState* s, u64* n_left, u64* n_right, std::string* err) {
    if (!s->cfg.unordered) { *err = "fqd_unordered_prepare is for --unordered handles"; return FQD_ERR_INVALID; }
    if (s->finished || s->un.stage) { *err = "fqd_unordered_prepare called twice"; return FQD_ERR_INVALID; }
    int rc = seq_parse_rest(s, err, true);          // a range may well be empty on one side or on both
    if (rc) return rc;
    *n_left = s->mate[0].n_records; *n_right = s->mate[1].n_records;
    if (s->stats.err) return FQD_OK;
    SEQ_TRY(cudaEventRecord(s->ev0, s->stream));
    if ((rc = un_prepare(s, err))) return rc;
    SEQ_TRY(cudaEventRecord(s->ev1, s->stream));
    SEQ_TRY(cudaEventSynchronize(s->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); s->ms += ms;
    return FQD_OK;
}
static int seq_unordered_enter(SeqState* s, int side, u64 i, u64* pos, std::string* err) {
    if (s->un.stage < 1) { *err = "fqd_unordered_enter before fqd_unordered_prepare"; return FQD_ERR_INVALID; }
    const u64 na = s->mate[side ? 1 : 0].n_records;
    if (i > na) { *err = "fqd_unordered_enter: position outside the list"; return FQD_ERR_INVALID; }
    u32* d_out; u32 h = 0;
    int rc;
    if ((rc = seq_dalloc(s, &d_out, 1, err))) return rc;
    k_join_enter<<<1, 1, 0, s->stream>>>(un_join_params(s), side, (u32)i, d_out);
    s->launches++;
    SEQ_TRY(cudaMemcpyAsync(&h, d_out, sizeof h, cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    *pos = h;
    return FQD_OK;
}
// out: pairs emitted here, unmatched records here (the job adds one when its last comparison fails), emission index of
// the first pair holding a byte outside {A,C,G,T,N} (~0 = none), 1 when the job's last comparison happened here and matched
static int seq_unordered_join(SeqState* s, u64 limit_i, u64 limit_j, u64 final_i, u64 final_j, u64* out, std::string* err) {
    if (s->un.stage != 1) { *err = "fqd_unordered_join needs fqd_unordered_prepare first (once)"; return FQD_ERR_INVALID; }
    JoinStop g;
    g.limit_i = (u32)std::min<u64>(limit_i, s->mate[0].n_records); g.limit_j = (u32)std::min<u64>(limit_j, s->mate[1].n_records);
    g.is = final_i < s->mate[0].n_records ? (u32)final_i : 0xFFFFFFFFu;
    g.js = final_j < s->mate[1].n_records ? (u32)final_j : 0xFFFFFFFFu;
    g.final_equal = 0; g.use_a = 0;
    SEQ_TRY(cudaEventRecord(s->ev0, s->stream));
    int rc = un_join(s, &g, err);
    if (rc) return rc;
    SEQ_TRY(cudaEventRecord(s->ev1, s->stream));
    SEQ_TRY(cudaEventSynchronize(s->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); s->ms += ms;
    out[0] = s->un.E; out[1] = s->un.unmatched; out[2] = s->un.hbad == ~0ull ? ~0ull : (s->un.hbad >> 8); out[3] = s->un.final_equal;
    return FQD_OK;
}
// rows [pair key 2W words][hash] of the pairs before `limit`, grouped by the owner of their hash range, emission order
// kept inside a group; counts[n_shards]
static int seq_unordered_rows(SeqState* s, u64 limit, u32 n_shards, void* d_send, u64* counts, std::string* err) {
    UnJoin& u = s->un;
    if (u.stage != 2) { *err = "fqd_unordered_rows needs fqd_unordered_join first"; return FQD_ERR_INVALID; }
    if (n_shards == 0 || n_shards > 64) { *err = "bad number of hash ranges"; return FQD_ERR_INVALID; }
    limit = std::min(limit, u.E);
    for (u32 o = 0; o < n_shards; ++o) counts[o] = 0;
    u.n_send = limit;
    if (limit == 0) return FQD_OK;
    int rc;
    SortScratch sc;
    u64* d_starts;
    if ((rc = sort_scratch_alloc(s, sc, limit, err)) || (rc = seq_dalloc(s, &d_starts, n_shards + 1, err)) ||
        (rc = seq_dalloc(s, &u.send_idx, limit, err))) return rc;
    SEQ_TRY(cudaEventRecord(s->ev0, s->stream));
    k_un_owner<<<seq_grid(s, limit), 256, 0, s->stream>>>(u.h1, u.h2, limit, n_shards, sc.keyA, sc.aA);
    if ((rc = radix_sort(s, sc, limit, 0, 8, false, err))) return rc;
    k_range_starts<<<1, 128, 0, s->stream>>>(sc.keyA, limit, n_shards, d_starts);
    SEQ_TRY(cudaMemcpyAsync(u.send_idx, sc.aA, limit * sizeof(u32), cudaMemcpyDeviceToDevice, s->stream));
    k_un_gather<<<seq_grid(s, limit * (2 * s->W + 1)), 256, 0, s->stream>>>(u.pair_rows, 2 * s->W, u.h1, u.h2, u.send_idx, limit, (u64*)d_send);
    s->launches += 3;
    std::vector<u64> h_starts(n_shards + 1);
    SEQ_TRY(cudaMemcpyAsync(h_starts.data(), d_starts, (n_shards + 1) * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaEventRecord(s->ev1, s->stream));
    SEQ_TRY(cudaEventSynchronize(s->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); s->ms += ms;
    for (u32 o = 0; o < n_shards; ++o) counts[o] = h_starts[o + 1] - h_starts[o];
    SEQ_TRY(cudaGetLastError());
    return FQD_OK;
}
// owner side: the rows every range sent here, in rank order (= the job's emission order); d_flags[i] = 1 when an
// earlier row holds the same key
static int seq_unordered_insert(SeqState* s, const void* d_recv, u64 n_recv, u32 n_shards, void* d_flags, std::string* err) {
    if (!s->cfg.unordered) { *err = "fqd_unordered_insert is for --unordered handles"; return FQD_ERR_INVALID; }
    if (n_recv == 0) return FQD_OK;
    if (n_recv > 0xFFFFFFF0ull) { *err = "too many rows for one hash range"; return FQD_ERR_CAPACITY; }
    const u32 RW = 2 * s->W;
    u64 *keys, *hash; RunState* run;
    int rc;
    if ((rc = seq_dalloc(s, &keys, n_recv * RW, err)) || (rc = seq_dalloc(s, &hash, n_recv, err)) || (rc = seq_dalloc(s, &run, 1, err))) return rc;
    SEQ_TRY(cudaEventRecord(s->ev0, s->stream));
    SEQ_TRY(cudaMemsetAsync(d_flags, 0, n_recv, s->stream));
    SEQ_TRY(cudaMemsetAsync(run, 0, sizeof(RunState), s->stream));
    k_shard_append<<<seq_grid(s, n_recv * (RW + 1)), 256, 0, s->stream>>>((const u64*)d_recv, (u32)n_recv, RW, keys, run, n_recv, hash);
    s->launches++;
    if ((rc = un_insert_rows(s, keys, hash, nullptr, n_recv, n_shards, (u8*)d_flags, err))) return rc;
    SEQ_TRY(cudaEventRecord(s->ev1, s->stream));
    SEQ_TRY(cudaEventSynchronize(s->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); s->ms += ms;
    return FQD_OK;
}
// flags of the rows this range sent (in sending order) -> survivors and emission lists.  unmatched: the JOB's figure is
// kept by the driver; this handle reports its own share.
static int seq_unordered_apply(SeqState* s, const void* d_flags_back, u64 limit, int report_bad, std::string* err) {
    UnJoin& u = s->un;
    if (u.stage != 2) { *err = "fqd_unordered_apply needs fqd_unordered_join first"; return FQD_ERR_INVALID; }
    if (s->finished) { *err = "fqd_finish called twice"; return FQD_ERR_INVALID; }
    s->finished = true;
    limit = std::min(limit, u.E);
    SEQ_TRY(cudaEventRecord(s->ev0, s->stream));
    if (u.n_send) {
        k_un_flags_back<<<seq_grid(s, u.n_send), 256, 0, s->stream>>>((const u8*)d_flags_back, u.send_idx, u.n_send, u.dup);
        s->launches++;
    }
    s->stats.unmatched = u.unmatched;
    int rc = un_finish(s, limit, report_bad != 0, err);
    if (rc) return rc;
    SEQ_TRY(cudaEventRecord(s->ev1, s->stream));
    SEQ_TRY(cudaEventSynchronize(s->ev1));
    float ms = 0; cudaEventElapsedTime(&ms, s->ev0, s->ev1); s->ms += ms;
    return FQD_OK;
}

// The (offset, length) emission lists on the host: only fqd_emission() needs them (the CLI gathers on the device).
static int seq_lists_to_host(SeqState* s, std::string* err) {
    if (s->lists_on_host) return FQD_OK;
    for (u32 m = 0; m < s->mates; ++m) {
        s->h_off[m].resize(s->n_out); s->h_len[m].resize(s->n_out);
        if (s->n_out && s->d_o_off[m]) {
            SEQ_TRY(cudaMemcpyAsync(s->h_off[m].data(), s->d_o_off[m], s->n_out * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
            SEQ_TRY(cudaMemcpyAsync(s->h_len[m].data(), s->d_o_len[m], s->n_out * sizeof(u32), cudaMemcpyDeviceToHost, s->stream));
        }
    }
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    s->lists_on_host = true;
    return FQD_OK;
}

// ---------------------------------------------------------------------------------------------------------
// device-side output gather (one warp per record)
__global__ void k_gather_records(const u64* off, const u32* len, const u32* dst, u64 count, const u64* seg_base, u8* const* seg_ptr,
                                 u32 n_segs, u8* out) {
    const u32 lane = threadIdx.x & 31u;
    const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 r0 = (((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; r0 < count; r0 += warps * 32u) {
        const u64 r = r0 + lane;
        const u8* src = nullptr; u8* d = nullptr; u32 n = 0;
        if (r < count) {
            src = seg_locate(off[r], seg_base, seg_ptr, n_segs);
            d = out + dst[r];
            n = len[r];
        }
        warp_copy_records(src, d, n, 0u, (u32)min((u64)32, count - r0), lane);
    }
}

// ---- --write-clusters (src/seq_dup_remover.hpp:60-62,75-76,89-101 and the paired twin; src/file_utils.cpp:98-112):
// one line per record in sorted order - the ID line of a written record (cluster head), "--" + ID line of a removed one
__global__ void k_cluster_lens(const u32* perm, const u32* keep, const u64* rec_off, const u32* rec_len, u64 n,
                               const u64* seg_base, u8* const* seg_ptr, u32 n_segs, u32* out_len) {
    const u32 lane = threadIdx.x & 31u;
    const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 i = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const u32 g = perm[i];
        const u8* src = seg_locate(rec_off[g], seg_base, seg_ptr, n_segs);
        const u32 len = rec_len[g];
        u32 idlen = len;
        for (u32 j0 = 0; j0 < len; j0 += 32) {
            const u32 j = j0 + lane;
            const u32 hit = __ballot_sync(0xFFFFFFFFu, j < len && src[j] == '\n');
            if (hit) { idlen = j0 + (u32)__ffs((int)hit); break; }
        }
        if (lane == 0) out_len[i] = idlen + (keep[i] ? 0u : 2u);
    }
}
__global__ void k_gather_clusters(const u32* perm, const u32* keep, const u64* rec_off, const u32* cl_len, const u32* dst, u64 k0, u64 count,
                                  const u64* seg_base, u8* const* seg_ptr, u32 n_segs, u8* out) {
    const u32 lane = threadIdx.x & 31u;
    const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 r0 = (((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; r0 < count; r0 += warps * 32u) {
        const u64 r = r0 + lane;
        const u8* src = nullptr; u8* d = nullptr; u32 n = 0, pre = 0;
        if (r < count) {
            const u64 i = k0 + r;
            src = seg_locate(rec_off[perm[i]], seg_base, seg_ptr, n_segs);
            pre = keep[i] ? 0u : 2u;
            n = cl_len[i] - pre;
            d = out + dst[r];
        }
        warp_copy_records(src, d, n, pre, (u32)min((u64)32, count - r0), lane);
    }
}

// per mate: where every input segment lives (logical base -> device pointer), for the record gathers
static int seq_upload_segtab(SeqState* s, int m, std::string* err) {
    if (s->discard) { *err = "the raw input was discarded (fqd_discard_input): fetch the records with fqd_emission_read / fqd_cluster_read"; return FQD_ERR_INVALID; }
    if (s->d_seg_base[m]) return FQD_OK;
    SeqMate& mt = s->mate[m];
    std::vector<u64> base; std::vector<u8*> ptr;
    for (auto& sg : mt.segs) { base.push_back(sg.logical_base); ptr.push_back(sg.d); }
    s->n_segs[m] = (u32)base.size();
    int rc;
    if ((rc = seq_dalloc(s, &s->d_seg_base[m], base.size(), err)) || (rc = seq_dalloc(s, &s->d_seg_ptr[m], ptr.size(), err))) return rc;
    SEQ_TRY(cudaMemcpyAsync(s->d_seg_base[m], base.data(), base.size() * sizeof(u64), cudaMemcpyHostToDevice, s->stream));
    SEQ_TRY(cudaMemcpyAsync(s->d_seg_ptr[m], ptr.data(), ptr.size() * sizeof(u8*), cudaMemcpyHostToDevice, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    return FQD_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Repartition by key range (multi-GPU sequence mode): every rank parses its slice of the input,
// the ranks agree on G-1 splitters over (word 0, word 1) of the key rows, every record goes to the rank that owns
// its key range - as raw bytes, grouped by owner, input order kept - and that rank deduplicates what it received
// with the ordinary single-GPU engine.  Rows with equal leading words share an owner, so exact duplicates never
// straddle two ranks; prefix and Hamming neighbours can, which the boundary state above takes care of.
static int seq_partition_sample(SeqState* s, u32 n_samples, u64* out, u64* n_records, std::string* err) {
    const bool un = s->cfg.unordered != 0;
    int rc = seq_parse_rest(s, err, un);
    if (rc) return rc;
    for (u32 i = 0; i < 2 * n_samples; ++i) out[i] = ~0ull;
    if (un) {
        // --unordered: the two files are partitioned independently, each by its own ID tags; half of the samples each
        *n_records = s->stats.err ? 0 : s->mate[0].n_records + s->mate[1].n_records;
        if (s->stats.err || n_samples < 2) return FQD_OK;
        const u32 half = n_samples / 2;
        u64* d_out;
        if ((rc = seq_dalloc(s, &d_out, 2ull * n_samples, err))) return rc;
        SEQ_TRY(cudaMemsetAsync(d_out, 0xFF, 2ull * n_samples * sizeof(u64), s->stream));
        for (u32 m = 0; m < 2; ++m) {
            if (!s->mate[m].n_records) continue;
            k_sample_rows<<<(half + 255) / 256, 256, 0, s->stream>>>(s->mate[m].d_tags, s->TW, s->TW, s->mate[m].n_records, half, d_out + 2ull * m * half);
            s->launches++;
        }
        SEQ_TRY(cudaMemcpyAsync(out, d_out, 2ull * n_samples * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
        SEQ_TRY(cudaStreamSynchronize(s->stream));
        return FQD_OK;
    }
    *n_records = s->stats.err ? 0 : s->n;
    if (s->stats.err == FQD_ERR_EMPTY) { s->stats.err = 0; s->n = 0; }      // an empty slice is fine here
    if (s->stats.err || s->n == 0 || n_samples == 0) return FQD_OK;
    u64* d_out;
    if ((rc = seq_dalloc(s, &d_out, 2ull * n_samples, err))) return rc;
    k_sample_rows<<<(n_samples + 255) / 256, 256, 0, s->stream>>>(s->d_keys, s->row_words, s->row_words, s->n, n_samples, d_out);
    s->launches++;
    SEQ_TRY(cudaMemcpyAsync(out, d_out, 2ull * n_samples * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    return FQD_OK;
}

// One partition: records [0, n) with key rows `rows` go to the owner of their (word 0, word 1); perm = the records
// grouped by owner (input order kept), counts[G] += records per owner, and for the mates [m0, m1) that travel with this
// partition the byte offset of every record in the owner-grouped stream + bytes[m * G + o].
static int partition_one(SeqState* s, const u64* rows, u32 stride, u32 nw, u64 n, const u64* d_split, u32 G, u32 m0, u32 m1,
                         u32** perm_out, u64* counts, u64* bytes, std::string* err) {
    if (n == 0) return FQD_OK;
    int rc;
    SortScratch sc;
    u64* d_starts;
    if ((rc = sort_scratch_alloc(s, sc, n, err)) || (rc = seq_dalloc(s, &d_starts, G + 1, err)) || (rc = seq_dalloc(s, perm_out, n, err))) return rc;
    k_range_owner<<<seq_grid(s, n), 256, 0, s->stream>>>(rows, stride, nw, n, d_split, G - 1, sc.keyA, sc.aA);
    if ((rc = radix_sort(s, sc, n, 0, 8, false, err))) return rc;
    k_range_starts<<<1, 128, 0, s->stream>>>(sc.keyA, n, G, d_starts);
    SEQ_TRY(cudaMemcpyAsync(*perm_out, sc.aA, n * sizeof(u32), cudaMemcpyDeviceToDevice, s->stream));
    s->launches += 2;
    std::vector<u64> h_starts(G + 1), h_pick(G + 1);
    SEQ_TRY(cudaMemcpyAsync(h_starts.data(), d_starts, (G + 1) * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    u32* lens;
    u64* d_pick;
    if ((rc = seq_dalloc(s, &lens, n, err)) || (rc = seq_dalloc(s, &d_pick, G + 1, err))) return rc;
    for (u32 m = m0; m < m1; ++m) {
        if ((rc = seq_dalloc(s, &s->d_part_off[m], n + 1, err))) return rc;
        k_gather_u32<<<seq_grid(s, n), 256, 0, s->stream>>>(s->mate[m].d_rec_len, *perm_out, n, lens);
        const u64 tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
        SEQ_TRY(cudaMemsetAsync(sc.scan_state, 0, tiles * sizeof(u64), s->stream));
        SEQ_TRY(cudaMemsetAsync(sc.ticket, 0, sizeof(u32), s->stream));
        k_scan_exclusive_t<u64><<<(unsigned)tiles, SCAN_THREADS, 0, s->stream>>>(lens, s->d_part_off[m], n, sc.scan_state, sc.ticket, sc.d_total);
        SEQ_TRY(cudaMemcpyAsync(s->d_part_off[m] + n, sc.d_total, sizeof(u64), cudaMemcpyDeviceToDevice, s->stream));
        k_pick_u64<<<1, 128, 0, s->stream>>>(s->d_part_off[m], d_starts, G + 1, n + 1, 0, d_pick);
        s->launches += 3;
        SEQ_TRY(cudaMemcpyAsync(h_pick.data(), d_pick, (G + 1) * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
        SEQ_TRY(cudaStreamSynchronize(s->stream));
        for (u32 o = 0; o < G; ++o) bytes[m * G + o] = h_pick[o + 1] - h_pick[o];
    }
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    for (u32 o = 0; o < G; ++o) counts[o] += h_starts[o + 1] - h_starts[o];
    SEQ_TRY(cudaGetLastError());
    return FQD_OK;
}

// splitters: (G-1) x 2 words, ascending.  counts[G]: records per owner (--unordered: of both files together);
// bytes[mates * G]: raw bytes per mate and owner
static int seq_partition_plan(SeqState* s, const u64* splitters, u32 G, u64* counts, u64* bytes, std::string* err) {
    if (!s->parsed) { *err = "fqd_partition_plan before fqd_partition_sample"; return FQD_ERR_INVALID; }
    if (G == 0 || G > 64) { *err = "bad number of key ranges"; return FQD_ERR_INVALID; }
    const bool un = s->cfg.unordered != 0;
    s->part_G = G;
    s->part_n = un ? s->mate[0].n_records : s->n;
    s->part_n1 = un ? s->mate[1].n_records : 0;
    for (u32 o = 0; o < G; ++o) { counts[o] = 0; for (u32 m = 0; m < s->mates; ++m) bytes[m * G + o] = 0; }
    if (s->part_n + s->part_n1 == 0) return FQD_OK;
    int rc;
    u64* d_split;
    if ((rc = seq_dalloc(s, &d_split, 2ull * std::max(G, 2u), err))) return rc;
    if (G > 1) SEQ_TRY(cudaMemcpyAsync(d_split, splitters, 2ull * (G - 1) * sizeof(u64), cudaMemcpyHostToDevice, s->stream));
    if (!un) return partition_one(s, s->d_keys, s->row_words, s->row_words, s->part_n, d_split, G, 0, s->mates, &s->d_part_perm, counts, bytes, err);
    if ((rc = partition_one(s, s->mate[0].d_tags, s->TW, s->TW, s->part_n, d_split, G, 0, 1, &s->d_part_perm, counts, bytes, err))) return rc;
    return partition_one(s, s->mate[1].d_tags, s->TW, s->TW, s->part_n1, d_split, G, 1, 2, &s->d_part_perm1, counts, bytes, err);
}

// the records of one mate, grouped by owner (input order inside a group), written back to back into d_out
static int seq_partition_gather(SeqState* s, int m, void* d_out, std::string* err) {
    if (m < 0 || (u32)m >= s->mates) { *err = "bad mate index"; return FQD_ERR_INVALID; }
    const bool second = s->cfg.unordered && m == 1;
    const u32* perm = second ? s->d_part_perm1 : s->d_part_perm;
    const u64 n = second ? s->part_n1 : s->part_n;
    if (!perm && n) { *err = "fqd_partition_gather before fqd_partition_plan"; return FQD_ERR_INVALID; }
    if (n == 0) return FQD_OK;
    int rc = seq_upload_segtab(s, m, err);
    if (rc) return rc;
    k_gather_records_perm<<<seq_grid(s, n), 256, 0, s->stream>>>(s->mate[m].d_rec_off, s->mate[m].d_rec_len, perm, s->d_part_off[m], n,
                                                                 s->d_seg_base[m], s->d_seg_ptr[m], s->n_segs[m], (u8*)d_out);
    s->launches++;
    SEQ_TRY(cudaGetLastError());
    return FQD_OK;
}

static int seq_emit(SeqState* s, int m, void* dst, size_t cap, size_t* n_bytes, int* done, std::string* err) {
    if (!s->finished) { *err = "fqd_emit before fqd_finish"; return FQD_ERR_INVALID; }
    if (m < 0 || (u32)m >= s->mates) { *err = "bad mate index"; return FQD_ERR_INVALID; }
    *n_bytes = 0; *done = 0;
    if (s->stats.err && s->stats.err != FQD_ERR_BAD_BASE) { *done = 1; return FQD_OK; }
    { int rc0 = seq_upload_segtab(s, m, err); if (rc0) return rc0; }
    const u64 k0 = s->emit_cursor[m];
    if (k0 >= s->n_out) { *done = 1; return FQD_OK; }
    // as many whole records as fit into cap (and into 2^31 bytes, the u32 offset range of one batch)
    const size_t lim = std::min<size_t>(cap, (size_t)1 << 31);
    constexpr u64 LENWIN = 1u << 20;          // record lengths are fetched window by window into pinned memory
    if (!s->h_lenwin) SEQ_TRY(cudaHostAlloc(&s->h_lenwin, LENWIN * sizeof(u32), cudaHostAllocDefault));
    const u64 nwin = std::min<u64>(LENWIN, s->n_out - k0);
    SEQ_TRY(cudaMemcpyAsync(s->h_lenwin, s->d_o_len[m] + k0, nwin * sizeof(u32), cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    u64 k1 = k0; size_t bytes = 0;
    while (k1 < k0 + nwin && bytes + s->h_lenwin[k1 - k0] <= lim) { bytes += s->h_lenwin[k1 - k0]; ++k1; }
    if (k1 == k0) { *err = "fqd_emit: cap is smaller than one record"; return FQD_ERR_INVALID; }
    const u64 cnt = k1 - k0;
    int rc;
    if (s->stage_cap < bytes) { if ((rc = seq_dalloc(s, &s->d_stage, bytes + 256, err))) return rc; s->stage_cap = bytes; }
    if (s->dst_cap < cnt) {
        if ((rc = seq_dalloc(s, &s->d_dst, cnt, err))) return rc;
        s->dst_cap = cnt;
        if ((rc = seq_dalloc(s, &s->em_scan_state, (cnt + SCAN_TILE - 1) / SCAN_TILE + 16, err)) || (rc = seq_dalloc(s, &s->em_ticket, 4, err)) ||
            (rc = seq_dalloc(s, &s->em_total, 2, err))) return rc;
    }
    SortScratch sc; sc.scan_state = s->em_scan_state; sc.ticket = s->em_ticket; sc.d_total = s->em_total;
    if ((rc = device_scan(s, sc, s->d_o_len[m] + k0, s->d_dst, cnt, nullptr, err))) return rc;
    k_gather_records<<<seq_grid(s, cnt), 256, 0, s->stream>>>(s->d_o_off[m] + k0, s->d_o_len[m] + k0, s->d_dst, cnt, s->d_seg_base[m],
                                                                   s->d_seg_ptr[m], s->n_segs[m], s->d_stage);
    s->launches++;
    SEQ_TRY(cudaMemcpyAsync(dst, s->d_stage, bytes, cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    s->emit_cursor[m] = k1;
    *n_bytes = bytes;
    *done = k1 >= s->n_out ? 1 : 0;
    return FQD_OK;
}

// Streams the <out>.clusters text of one mate (sequence-based modes), like seq_emit streams the records.
static int seq_emit_clusters(SeqState* s, int m, void* dst, size_t cap, size_t* n_bytes, int* done, std::string* err) {
    if (!s->finished) { *err = "fqd_emit_clusters before fqd_finish"; return FQD_ERR_INVALID; }
    if (s->cfg.unordered || s->cfg.mode == FQD_MODE_FAST) { *err = "cluster files exist in sequence-based modes only"; return FQD_ERR_INVALID; }
    if (m < 0 || (u32)m >= s->mates) { *err = "bad mate index"; return FQD_ERR_INVALID; }
    *n_bytes = 0; *done = 0;
    if (s->stats.err || s->n == 0) { *done = 1; return FQD_OK; }
    int rc = seq_upload_segtab(s, m, err);
    if (rc) return rc;
    const u64 n = s->n;
    if (!s->d_cl_len[m]) {
        if ((rc = seq_dalloc(s, &s->d_cl_len[m], n, err))) return rc;
        k_cluster_lens<<<seq_grid(s, n * 32), 256, 0, s->stream>>>(s->d_perm, s->d_keep, s->mate[m].d_rec_off, s->mate[m].d_rec_len, n,
                                                                   s->d_seg_base[m], s->d_seg_ptr[m], s->n_segs[m], s->d_cl_len[m]);
        s->launches++;
    }
    const u64 k0 = s->cl_cursor[m];
    if (k0 >= n) { *done = 1; return FQD_OK; }
    const size_t lim = std::min<size_t>(cap, (size_t)1 << 31);
    constexpr u64 LENWIN = 1u << 20;
    if (!s->h_lenwin) SEQ_TRY(cudaHostAlloc(&s->h_lenwin, LENWIN * sizeof(u32), cudaHostAllocDefault));
    const u64 nwin = std::min<u64>(LENWIN, n - k0);
    SEQ_TRY(cudaMemcpyAsync(s->h_lenwin, s->d_cl_len[m] + k0, nwin * sizeof(u32), cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    u64 k1 = k0; size_t bytes = 0;
    while (k1 < k0 + nwin && bytes + s->h_lenwin[k1 - k0] <= lim) { bytes += s->h_lenwin[k1 - k0]; ++k1; }
    if (k1 == k0) { *err = "fqd_emit_clusters: cap is smaller than one line"; return FQD_ERR_INVALID; }
    const u64 cnt = k1 - k0;
    if (s->stage_cap < bytes) { if ((rc = seq_dalloc(s, &s->d_stage, bytes + 256, err))) return rc; s->stage_cap = bytes; }
    if (s->dst_cap < cnt) {
        if ((rc = seq_dalloc(s, &s->d_dst, cnt, err))) return rc;
        s->dst_cap = cnt;
        if ((rc = seq_dalloc(s, &s->em_scan_state, (cnt + SCAN_TILE - 1) / SCAN_TILE + 16, err)) || (rc = seq_dalloc(s, &s->em_ticket, 4, err)) ||
            (rc = seq_dalloc(s, &s->em_total, 2, err))) return rc;
    }
    SortScratch sc; sc.scan_state = s->em_scan_state; sc.ticket = s->em_ticket; sc.d_total = s->em_total;
    if ((rc = device_scan(s, sc, s->d_cl_len[m] + k0, s->d_dst, cnt, nullptr, err))) return rc;
    k_gather_clusters<<<seq_grid(s, cnt), 256, 0, s->stream>>>(s->d_perm, s->d_keep, s->mate[m].d_rec_off, s->d_cl_len[m], s->d_dst, k0, cnt,
                                                                    s->d_seg_base[m], s->d_seg_ptr[m], s->n_segs[m], s->d_stage);
    s->launches++;
    SEQ_TRY(cudaMemcpyAsync(dst, s->d_stage, bytes, cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    s->cl_cursor[m] = k1;
    *n_bytes = bytes;
    *done = k1 >= n ? 1 : 0;
    return FQD_OK;
}

static int seq_emission(SeqState* s, fqd_emission_t* out, std::string* err) {
    if (!s->finished) { *err = "fqd_emission before fqd_finish"; return FQD_ERR_INVALID; }
    memset(out, 0, sizeof *out);
    int rc = seq_lists_to_host(s, err);
    if (rc) return rc;
    out->n_out = s->n_out;
    for (u32 m = 0; m < s->mates; ++m) { out->off[m] = (const uint64_t*)s->h_off[m].data(); out->len[m] = s->h_len[m].data(); }
    return FQD_OK;
}
// A window of one mate's emission list, straight into the caller's memory (fqd_emission_read).
static int seq_emission_read(SeqState* s, int m, u64 first, u64 count, u64* off, u32* len, std::string* err) {
    if (!s->finished) { *err = "fqd_emission_read before fqd_finish"; return FQD_ERR_INVALID; }
    if (m < 0 || (u32)m >= s->mates) { *err = "bad mate index"; return FQD_ERR_INVALID; }
    if (first > s->n_out || count > s->n_out - first) { *err = "fqd_emission_read: window beyond the emission list"; return FQD_ERR_INVALID; }
    if (count == 0) return FQD_OK;
    if (!s->d_o_off[m] || !s->d_o_len[m]) { *err = "fqd_emission_read: no emission list"; return FQD_ERR_INVALID; }
    SEQ_TRY(cudaMemcpyAsync(off, s->d_o_off[m] + first, count * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaMemcpyAsync(len, s->d_o_len[m] + first, count * sizeof(u32), cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    return FQD_OK;
}

// --write-clusters without the raw input: who stands at the sorted positions [first, first + count) and is it written?
__global__ void k_cluster_index(const u32* __restrict__ perm, const u32* __restrict__ keep, const u64* __restrict__ rec_off,
                                const u32* __restrict__ rec_len, u64 first, u64 count, u64* __restrict__ out_off, u32* __restrict__ out_len,
                                u8* __restrict__ out_head) {
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < count; r += (u64)gridDim.x * blockDim.x) {
        const u32 g = perm[first + r];
        out_off[r] = rec_off[g];
        out_len[r] = rec_len[g];
        out_head[r] = keep[first + r] ? 1 : 0;
    }
}
static int seq_cluster_read(SeqState* s, int m, u64 first, u64 count, u64* off, u32* len, u8* head, std::string* err) {
    if (!s->finished) { *err = "fqd_cluster_read before fqd_finish"; return FQD_ERR_INVALID; }
    if (s->cfg.unordered || s->cfg.mode == FQD_MODE_FAST) { *err = "cluster files exist in sequence-based modes only"; return FQD_ERR_INVALID; }
    if (m < 0 || (u32)m >= s->mates) { *err = "bad mate index"; return FQD_ERR_INVALID; }
    const u64 n = s->stats.err ? 0 : s->n;
    if (first > n || count > n - first) { *err = "fqd_cluster_read: window beyond the records processed"; return FQD_ERR_INVALID; }
    if (count == 0) return FQD_OK;
    if (!s->d_perm || !s->d_keep) { *err = "fqd_cluster_read: no sorted order"; return FQD_ERR_INVALID; }
    int rc;
    if (s->clr_cap < count) {
        if ((rc = seq_dalloc(s, &s->d_clr_off, count, err)) || (rc = seq_dalloc(s, &s->d_clr_len, count, err)) ||
            (rc = seq_dalloc(s, &s->d_clr_head, count, err))) return rc;
        s->clr_cap = count;
    }
    k_cluster_index<<<seq_grid(s, count), 256, 0, s->stream>>>(s->d_perm, s->d_keep, s->mate[m].d_rec_off, s->mate[m].d_rec_len, first, count,
                                                              s->d_clr_off, s->d_clr_len, s->d_clr_head);
    s->launches++;
    SEQ_TRY(cudaMemcpyAsync(off, s->d_clr_off, count * sizeof(u64), cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaMemcpyAsync(len, s->d_clr_len, count * sizeof(u32), cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaMemcpyAsync(head, s->d_clr_head, count, cudaMemcpyDeviceToHost, s->stream));
    SEQ_TRY(cudaStreamSynchronize(s->stream));
    return FQD_OK;
}

static void seq_stats(SeqState* s, fqd_stats_t* st) { *st = s->stats; }
static double seq_device_ms(SeqState* s) { return s->ms; }
static u64 seq_launches(SeqState* s) { return s->launches; }

}  // namespace fqd
