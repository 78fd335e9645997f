// sortlib.cuh - device primitives for the whole-input modes: single-pass exclusive scan, stable LSD radix sort
// of (64-bit key, 32-bit payload[, 32-bit payload]) and the multi-word "sort by rounds" driver that replaces
// the reference's std::sort + k-way heap merge on disk (src/external_sort.hpp:88-207,
// src/paired_external_sort.hpp:112-257).  Order = unsigned lexicographic order of a key row (word 0 first),
// ties broken by the smaller record index (stable), which is the documented tie-break.
#pragma once
#include "common.cuh"

namespace fqd {

// ---------------------------------------------------------------------------------------------------------
// exclusive scan of u32 (decoupled look-back, one pass).  state[] must hold ceil(n / SCAN_TILE) zeroed words.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr u32 SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class OUT>      // u32 (the caller knows the total fits) or u64
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_exclusive_t(const u32* __restrict__ in, OUT* __restrict__ out, u64 n,
                                                                    u64* state, u32* ticket, u64* total_out) {
    __shared__ u32 warp_sum[SCAN_THREADS / 32];
    __shared__ u32 s_tile;
    __shared__ u64 s_prefix;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    const u64 base = (u64)tile * SCAN_TILE + (u64)tid * SCAN_ITEMS;
    u32 v[SCAN_ITEMS];
    u32 sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        sum += v[i];
    }
    u32 incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (u32)d) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 ws = lane < SCAN_THREADS / 32 ? warp_sum[lane] : 0u;
        u32 wi = ws;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            u32 t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= (u32)d) wi += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sum[lane] = wi - ws;
        const u64 total = __shfl_sync(0xFFFFFFFFu, wi, SCAN_THREADS / 32 - 1);
        // state word: flag (2 bits) << 62 | value (62 bits)
        u64 P = 0;
        if (tile == 0) {
            if (lane == 0) st_volatile_u64(state, (2ull << 62) | total);
        } else {
            if (lane == 0) st_volatile_u64(state + tile, (1ull << 62) | total);
            long long look = (long long)tile - 1;
            for (;;) {
                long long idx = look - lane;
                u64 s = (2ull << 62);
                if (idx >= 0) { do { s = ld_volatile_u64(state + idx); } while ((s >> 62) == 0); }
                u32 is_prefix = __ballot_sync(0xFFFFFFFFu, (s >> 62) == 2);
                u64 val = s & ((1ull << 62) - 1);
                if (is_prefix) {
                    u32 first = (u32)__ffs((int)is_prefix) - 1u;
                    if (lane > first) val = 0;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xFFFFFFFFu, val, d);
                P += val;
                if (is_prefix) break;
                look -= 32;
            }
            if (lane == 0) st_volatile_u64(state + tile, (2ull << 62) | (P + total));
        }
        if (lane == 0) {
            s_prefix = P;
            if (total_out && (u64)(tile + 1) * SCAN_TILE >= n) *total_out = P + total;
        }
    }
    __syncthreads();
    OUT run = (OUT)s_prefix + (OUT)(warp_sum[warp] + (incl - sum));
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}
#define k_scan_exclusive k_scan_exclusive_t<u32>

// ---------------------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits, stable.  Items = (key64, a32[, b32]); one pass = histogram, scan, scatter.
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr u32 RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const u64* __restrict__ keys, u64 n, u32 shift, u32* __restrict__ hist, u32 nblocks) {
    __shared__ u32 h[256];
    const u32 tid = threadIdx.x;
    h[tid] = 0;
    __syncthreads();
    const u64 base = (u64)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int i = 0; i < RS_ITEMS; ++i) {
        u64 g = base + (u64)i * RS_THREADS + tid;
        if (g < n) atomicAdd(&h[(keys[g] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    hist[(u64)tid * nblocks + blockIdx.x] = h[tid];      // bin-major: a plain exclusive scan yields global bases
}

// Which bits differ between any two keys?  out[0] |= keys, out[1] &= keys: a digit whose bits are the same in every key
// (the leading zeros of a counter in an ID tag, the padding of a short last word) needs no pass.
__global__ void __launch_bounds__(256) k_key_or_and(const u64* __restrict__ keys, u64 n, u64* out) {
    u64 o = 0, a = ~0ull;
    const u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) { const u64 k = keys[i]; o |= k; a &= k; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { o |= __shfl_xor_sync(0xFFFFFFFFu, o, d); a &= __shfl_xor_sync(0xFFFFFFFFu, a, d); }
    if ((threadIdx.x & 31u) == 0) { atomicOr(out, o); atomicAnd(out + 1, a); }
}

// Scatter of one pass.  Round 1 wrote every item straight to its global position: 32 lanes, up to 32 different bins,
// 8- and 4-byte writes all over the output - the sort ran at a quarter of the bandwidth its traffic needs (10.5 ms for
// 8 passes over 50 M (key, index) pairs).  Now a tile's 4096 items are first put in digit order in SHARED memory
// (tile-local position = digits before + earlier items of the same digit), and written out from there: consecutive
// threads write consecutive items of a digit's run, i.e. consecutive global positions (runs of 16 items on average).
constexpr size_t rs_smem_bytes(bool has_b) {
    return (size_t)(RS_THREADS / 32) * 256 * 4 + 3 * 256 * 4 + (size_t)RS_TILE * (8 + 4 + (has_b ? 4 : 0));
}
template <bool HAS_B>
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const u64* __restrict__ keys, const u32* __restrict__ a, const u32* __restrict__ b,
                                                               u64* __restrict__ keys_out, u32* __restrict__ a_out, u32* __restrict__ b_out,
                                                               u64 n, u32 shift, const u32* __restrict__ base_scanned, u32 nblocks) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    u64* s_key = reinterpret_cast<u64*>(rs_smem);                                   // [RS_TILE]
    u32* s_a = reinterpret_cast<u32*>(s_key + RS_TILE);                             // [RS_TILE]
    u32* s_b = s_a + RS_TILE;                                                       // [RS_TILE] (HAS_B)
    u32 (*whist)[256] = reinterpret_cast<u32 (*)[256]>(s_a + (HAS_B ? 2 : 1) * RS_TILE);   // [warps][256]
    u32* gbase = &whist[0][0] + (RS_THREADS / 32) * 256;                            // [256] global start of this tile's digit run
    u32* dstart = gbase + 256;                                                      // [256] tile-local start of the digit
    u32* dscan = dstart + 256;                                                      // [256] scratch of the block scan
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (u32 i = tid; i < (RS_THREADS / 32) * 256; i += RS_THREADS) (&whist[0][0])[i] = 0;
    gbase[tid] = base_scanned[(u64)tid * nblocks + blockIdx.x];
    __syncthreads();
    // blocked arrangement keeps the input order: warp w owns items [w*512, (w+1)*512) of the tile, step s lane l
    const u64 tbase = (u64)blockIdx.x * RS_TILE;
    const u64 wbase = tbase + (u64)warp * (32 * RS_ITEMS);
    u64 k[RS_ITEMS];
    u16 rank[RS_ITEMS];
#pragma unroll
    for (int s = 0; s < RS_ITEMS; ++s) {
        const u64 g = wbase + (u64)s * 32 + lane;
        const bool ok = g < n;
        k[s] = ok ? keys[g] : ~0ull;
        const u32 d = ok ? (u32)((k[s] >> shift) & 0xFFu) : 256u;      // 256 = out of range: no peer among real digits
        const u32 peers = __match_any_sync(0xFFFFFFFFu, d);
        const u32 before = __popc(peers & ((1u << lane) - 1u));
        u32 old = 0;
        if (ok && before == 0) {                     // leader of its digit in this step
            old = whist[warp][d];
            whist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xFFFFFFFFu, old, __ffs((int)peers) - 1);
        rank[s] = (u16)(old + before);
        __syncwarp();                                // the next step's leaders read this step's counter updates
    }
    __syncthreads();
    {   // thread = digit: exclusive scan of the digit's counts across the warps; the digit's total goes into the block scan
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) {
            u32 c = whist[w][tid];
            whist[w][tid] = run;
            run += c;
        }
        // exclusive scan of the 256 totals -> tile-local start of every digit
        u32 incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (u32)d) incl += t;
        }
        if (lane == 31) dscan[warp] = incl;
        __syncthreads();
        u32 wofs = 0;
        for (u32 w = 0; w < warp; ++w) wofs += dscan[w];
        dstart[tid] = wofs + incl - run;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < RS_ITEMS; ++s) {
        const u64 g = wbase + (u64)s * 32 + lane;
        if (g < n) {
            const u32 d = (u32)((k[s] >> shift) & 0xFFu);
            const u32 lp = dstart[d] + whist[warp][d] + rank[s];
            s_key[lp] = k[s];
            s_a[lp] = a[g];
            if (HAS_B) s_b[lp] = b[g];
        }
    }
    __syncthreads();
    const u32 n_tile = (u32)min((u64)RS_TILE, n - tbase);
    for (u32 j = tid; j < n_tile; j += RS_THREADS) {
        const u64 key = s_key[j];
        const u32 d = (u32)((key >> shift) & 0xFFu);
        const u64 pos = (u64)gbase[d] + (j - dstart[d]);
        keys_out[pos] = key;
        a_out[pos] = s_a[j];
        if (HAS_B) b_out[pos] = s_b[j];
    }
}

// ---------------------------------------------------------------------------------------------------------
// small helper kernels of the multi-word sort
__global__ void k_iota_u32(u32* p, u64 n) {
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = (u32)i;
}
__global__ void k_memset_u32(u32* p, u64 n, u32 v) {
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}
// key[i] = word `w` of the row of record idx[i]  (rows: base + idx * stride)
__global__ void k_gather_word(const u64* __restrict__ rows, u32 stride, u32 w, const u32* __restrict__ idx, u64 n, u64* __restrict__ key) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) key[i] = rows[(u64)idx[i] * stride + w];
}
// After sorting the active items by (segment, word): flags of the refined segmentation.
//   head[i] = 1 when item i starts a new (segment, word) group
__global__ void k_mark_heads(const u64* __restrict__ key, const u32* __restrict__ seg, u64 n, u32* __restrict__ head) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
        head[i] = (i == 0 || key[i] != key[i - 1] || (seg && seg[i] != seg[i - 1])) ? 1u : 0u;
}
// seg_id[i] = (inclusive scan of head)[i] - 1 is produced by the exclusive scan + head; group sizes via atomics.
__global__ void k_group_ids(const u32* __restrict__ head, const u32* __restrict__ head_excl, u64 n, u32* __restrict__ gid, u32* __restrict__ gsize) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        u32 g = head_excl[i] + head[i] - 1u;
        gid[i] = g;
        atomicAdd(&gsize[g], 1u);
    }
}
// A group still needs refinement when it has more than one member and its members' rows differ somewhere in the
// words that have not been used yet (identical rows are already in index order thanks to stability).
__global__ void k_mark_unresolved(const u64* __restrict__ rows, u32 stride, u32 w_next, u32 n_words, const u32* __restrict__ idx,
                                  const u32* __restrict__ gid, u64 n, u32* __restrict__ gdiff) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        if (i == 0 || gid[i] != gid[i - 1]) continue;
        const u64* a = rows + (u64)idx[i] * stride;
        const u64* b = rows + (u64)idx[i - 1] * stride;
        u64 diff = 0;
        for (u32 w = w_next; w < n_words; ++w) diff |= a[w] ^ b[w];
        if (diff) gdiff[gid[i]] = 1u;
    }
}
// Small unresolved groups (a duplicate with a substitution near its end, a truncated copy next to its original: members
// that share their first word and differ somewhere later) are put in order right here, one thread per group, by an
// insertion sort on the remaining words - instead of one more round of gather + two radix sorts PER WORD until the
// differing word is reached (a difference in the last 10 bases of a 150 bp read took 7 rounds: 19 - 23 ms at 50 M
// pairs).  Stable: members arrive in index order and only strictly greater predecessors are shifted.
constexpr u32 SMALL_GROUP = 16;
constexpr u32 MID_GROUP = 1024;         // up to here blocks rank the group's members against each other (k_sort_mid_groups)
constexpr u32 MID_CHUNK = 64;           // members one block ranks at a time: a large group is spread over several blocks
constexpr u32 MID_LIST_CAP = 1u << 16;
constexpr u64 MID_WORK_CAP = 1ull << 27;   // row comparisons; beyond it (or beyond the list) the groups go to the next round instead
struct MidUnit { u32 first; u32 size; u32 chunk; };
struct MidCtl { u32 n_units; u32 n_small; u64 work; };
// The unresolved groups of this round, by size: heads of the small ones go to a dense list (so that the threads sorting
// them sit side by side - one group per thread on scattered lanes was bound by the latency of a single lane per warp:
// 2.5 ms for 1.5 M groups), the middle-sized ones become (group, chunk of 64 members) units.
__global__ void __launch_bounds__(256) k_collect_groups(const u32* __restrict__ head, const u32* __restrict__ gid, const u32* __restrict__ gsize,
                                                        const u32* __restrict__ gdiff, u64 n, u32* __restrict__ small_list,
                                                        MidUnit* __restrict__ units, MidCtl* __restrict__ mc) {
    const u64 step = (u64)gridDim.x * blockDim.x;
    const u32 lane = threadIdx.x & 31u;
    for (u64 b = (u64)blockIdx.x * blockDim.x; b < n; b += step) {
        const u64 i = b + threadIdx.x;
        u32 m = 0;
        if (i < n && head[i]) {
            const u32 g = gid[i];
            if (gdiff[g]) m = gsize[g];
        }
        const bool small = m >= 2u && m <= SMALL_GROUP;
        const u32 ballot = __ballot_sync(0xFFFFFFFFu, small);
        if (ballot) {
            u32 base = 0;
            if (lane == 0) base = atomicAdd(&mc->n_small, (u32)__popc(ballot));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (small) small_list[base + __popc(ballot & ((1u << lane) - 1u))] = (u32)i;
        }
        if (m > SMALL_GROUP && m <= MID_GROUP) {
            const u32 chunks = (m + MID_CHUNK - 1) / MID_CHUNK;
            const u32 k = atomicAdd(&mc->n_units, chunks);
            atomicAdd((unsigned long long*)&mc->work, (unsigned long long)m * m);
            for (u32 c = 0; c < chunks && k + c < MID_LIST_CAP; ++c) { units[k + c].first = (u32)i; units[k + c].size = m; units[k + c].chunk = c; }
        }
    }
}
__global__ void __launch_bounds__(256) k_sort_small_groups(const u64* __restrict__ rows, u32 stride, u32 w_next, u32 n_words,
                                                           const u32* __restrict__ gid, const u32* __restrict__ gsize, u32* __restrict__ gdiff,
                                                           const u32* __restrict__ idx, const u32* __restrict__ pos_in, u32* __restrict__ perm,
                                                           const u32* __restrict__ small_list, const MidCtl* __restrict__ mc) {
    const u32 n_list = mc->n_small;
    const u32 step = gridDim.x * blockDim.x;
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < n_list; e += step) {
        const u32 i = small_list[e];
        const u32 g = gid[i];
        const u32 m = gsize[g];
        u32 v[SMALL_GROUP];
        for (u32 k = 0; k < m; ++k) v[k] = idx[i + k];
        for (u32 k = 1; k < m; ++k) {
            const u32 cur = v[k];
            const u64* a = rows + (u64)cur * stride;
            u32 j = k;
            while (j > 0) {
                const u64* b = rows + (u64)v[j - 1] * stride;
                bool greater = false;                       // is the predecessor strictly greater than cur?
                for (u32 w = w_next; w < n_words; ++w) {
                    const u64 x = b[w], y = a[w];
                    if (x != y) { greater = x > y; break; }
                }
                if (!greater) break;
                v[j] = v[j - 1];
                --j;
            }
            v[j] = cur;
        }
        for (u32 k = 0; k < m; ++k) perm[pos_in ? pos_in[i + k] : i + k] = v[k];
        gdiff[g] = 0;                                       // resolved: not part of the next round
    }
}
// Groups of 17 .. 1024 members (a popular read with a few variants): every member's rank = members that are smaller, or
// equal and earlier (stable); a block takes 64 members, four threads each (a quarter of the comparisons per thread).
// A handful of such groups used to drag the whole sort through one round per remaining word (15 for a pair of 150 bp
// reads) with ~1 ms of fixed cost each.  All or nothing: when the input has very many of them (read names that share
// their first 8 bytes in groups of a hundred) the next radix round is the cheaper tool and the kernel leaves them alone.
__global__ void __launch_bounds__(256) k_sort_mid_groups(const u64* __restrict__ rows, u32 stride, u32 w_next, u32 n_words,
                                                         const u32* __restrict__ gid, u32* __restrict__ gdiff, const u32* __restrict__ idx,
                                                         const u32* __restrict__ pos_in, u32* __restrict__ perm,
                                                         const MidUnit* __restrict__ units, const MidCtl* __restrict__ mc) {
    __shared__ u32 v[MID_GROUP];
    if (mc->n_units > MID_LIST_CAP || mc->work > MID_WORK_CAP) return;      // many such groups: a radix round is cheaper
    const u32 n_units = mc->n_units;
    const u32 part = threadIdx.x & 3u;
    for (u32 e = blockIdx.x; e < n_units; e += gridDim.x) {
        const u32 first = units[e].first, m = units[e].size;
        __syncthreads();
        for (u32 t = threadIdx.x; t < m; t += blockDim.x) v[t] = idx[first + t];
        __syncthreads();
        const u32 t = units[e].chunk * MID_CHUNK + (threadIdx.x >> 2);
        u32 rank = 0;
        if (t < m) {
            const u64* a = rows + (u64)v[t] * stride;
            for (u32 j = part; j < m; j += 4u) {
                if (j == t) continue;
                const u64* b = rows + (u64)v[j] * stride;
                int c = 0;                                   // sign of (b - a) on the remaining words
                for (u32 w = w_next; w < n_words; ++w) {
                    const u64 x = b[w], y = a[w];
                    if (x != y) { c = x < y ? -1 : 1; break; }
                }
                if (c < 0 || (c == 0 && j < t)) ++rank;
            }
        }
        rank += __shfl_xor_sync(0xFFFFFFFFu, rank, 1);
        rank += __shfl_xor_sync(0xFFFFFFFFu, rank, 2);
        if (t < m && part == 0) perm[pos_in ? pos_in[first + rank] : first + rank] = v[t];
        if (threadIdx.x == 0 && units[e].chunk == 0) gdiff[gid[first]] = 0;         // resolved: not part of the next round
    }
}
__global__ void k_active_flags(const u32* __restrict__ gid, const u32* __restrict__ gsize, const u32* __restrict__ gdiff, u64 n, u32* __restrict__ flag) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        u32 g = gid[i];
        flag[i] = (gsize[g] > 1u && gdiff[g]) ? 1u : 0u;
    }
}
// compaction of the active items: (position in perm, record index, group id)
__global__ void k_compact_active(const u32* __restrict__ flag, const u32* __restrict__ excl, const u32* __restrict__ pos_in, const u32* __restrict__ idx,
                                 const u32* __restrict__ gid, u64 n, u32* __restrict__ pos_out, u32* __restrict__ idx_out, u32* __restrict__ gid_out) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        if (flag[i]) {
            u32 o = excl[i];
            pos_out[o] = pos_in ? pos_in[i] : (u32)i;
            idx_out[o] = idx[i];
            gid_out[o] = gid[i];
        }
    }
}
__global__ void k_scatter_perm(const u32* __restrict__ pos, const u32* __restrict__ idx, u64 n, u32* __restrict__ perm) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) perm[pos[i]] = idx[i];
}
__global__ void k_u32_to_u64key(const u32* __restrict__ in, u64 n, u64* __restrict__ out) {
    u64 step = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) out[i] = in[i];
}

}  // namespace fqd
