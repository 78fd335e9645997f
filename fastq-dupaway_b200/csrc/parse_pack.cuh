// parse_pack.cuh - K1: fused record splitter + sequence packer + key hasher (one pass over the raw bytes).
//
// Replaces, per record, FastqView::read_new / FastaView::read_new (src/fastqview.cpp:89-119,
// src/fastaview.cpp:75-93: 4 or 2 newline searches, '@'/'>' check, len(seq)==len(qual)) and
// SeqUtils::seq2hash (src/seq_utils.cpp:35-49) for a whole chunk of raw bytes resident in HBM.
//
// One CTA handles one 16 KiB tile (+1 KiB halo) staged in shared memory by a 1-D bulk async copy (TMA);
// tiles are taken in ticket order so that every predecessor of a tile is already running.
//   1. newline bitmask of the window (SIMD-in-register byte compare, dp4a gathers the flags)
//   2. block scan of per-thread newline counts -> local rank of every '\n'; positions compacted by rank into
//      shared memory; decoupled look-back across tiles (128 predecessors per hop) -> global rank of the tile's
//      first newline.  The rank modulo lines-per-record says which newline ends a record (the reference, too,
//      simply counts 4 (2) newlines per record).
//   3. one thread per record owned by the tile (its first byte lies in the tile): line ends = consecutive
//      entries of the compacted positions; validation; queue (sequence offset, length)
//   4. groups of 8 lanes pack one queued sequence each into 3-bit codes (20 bases per 64-bit word, see
//      common.cuh), write the key row with coalesced 64-bit stores and reduce the multilinear key hash.
#pragma once
#include "common.cuh"

namespace fqd {

constexpr int PP_THREADS = 256;
constexpr u32 PP_TILE = 16384;
constexpr u32 PP_HALO = 1024;
constexpr u32 PP_WINDOW = PP_TILE + PP_HALO;
constexpr u32 PP_NW = PP_WINDOW / 64;          // 64-bit mask words per window
constexpr u32 PP_QCAP = 256;                   // records packed per round
constexpr u32 PP_NLCAP = 1024;                 // newline positions compacted per window; denser tiles take the slow path
constexpr u32 PP_HKEYS = 64;                   // hash keys cached in shared memory (words per mate)
constexpr u32 PP_NONE = 0xFFFFFFFFu;
constexpr int PP_MIN_CTAS = 5;

struct ParseParams {
    const u8* raw;          // chunk bytes (16-byte aligned)
    u32 n;                  // chunk length
    u32 n_tiles;
    u64* tile_state;        // [n_tiles], zero-initialised: flag<<32 | value
    ChunkCtl* ctl;
    const RunState* run;    // slot base = run->n_records
    u32* rec_start;         // [cap+1] chunk-local record start offsets
    u32 cap;                // records per chunk the tables hold
    u64* keys;              // key store
    u64 key_capacity;       // slots in the key store
    u32 row_words;          // 64-bit words per key row (all mates)
    u32 mate_off;           // word offset of this mate inside the row
    u32 W;                  // words of this mate
    u64* hash;              // [cap] per-mate raw key hash (finalise with mix64), chunk-local index
    u32* seq_len;           // optional [cap]: sequence length in bases (chunk-local index)
    u64* word0;             // optional [cap]: first key word (radix-sort key), chunk-local index
    u8* dup;                // optional [cap]: duplicate flags, cleared here for every record of the chunk
    u32* bad_rec;           // optional [cap], initialised to ~0: (position << 8 | byte) of the first byte outside
                            // {A,C,G,T,N} of each record
    u32 strict;             // 1: bytes outside {A,C,G,T,N} are an error (fast mode, src/seq_utils.cpp:17-19)
    u32 hash_salt;          // distinguishes mates in the position keys
};

// ---- newline detection: 16 input bytes -> 16-bit mask of '\n' positions (bit j <-> byte j)
__device__ __forceinline__ u32 nl_flags(u32 w, u32 c_nl, u32 c_7f) {
    // exact per-byte test of w == 0x0A: 0x80 in each matching byte
    u32 a = (w ^ c_nl) & c_7f;
    u32 s = a + c_7f;
    return ~s & ~w & 0x80808080u;
}
__device__ __forceinline__ u32 nl_mask16(uint4 v, u32 c_nl, u32 c_7f) {
    // dp4a with weights 1,2,4,8 (16..128) turns the four 0x80 flags of a word into 128 * nibble
    u32 lo = __dp4a(nl_flags(v.x, c_nl, c_7f), 0x08040201u, 0u);
    lo = __dp4a(nl_flags(v.y, c_nl, c_7f), 0x80402010u, lo);
    u32 hi = __dp4a(nl_flags(v.z, c_nl, c_7f), 0x08040201u, 0u);
    hi = __dp4a(nl_flags(v.w, c_nl, c_7f), 0x80402010u, hi);
    return (lo + hi * 256u) >> 7;
}

// ---- packing
// 4 bases, first base in the most significant byte -> 12 bits of codes; diff = nonzero bytes where invalid
__device__ __forceinline__ u32 pack4(u32 v, u32& diff) {
    u32 x = (v >> 1) & 0x07070707u;
    u32 y = (x | (x >> 4)) & 0x00FF00FFu;
    u32 z = (y | (y >> 8)) & 0xFFFFu;                          // 4 selector nibbles
    u32 codes = __byte_perm(0x03050201u, 0x04000000u, z);      // idx (c>>1)&7: A0 C1 T2 G3 N7
    u32 canon = __byte_perm(0x47544341u, 0x4E000000u, z);      // 'A','C','T','G',0,0,0,'N'
    diff = v ^ canon;
    u32 t = (codes | (codes >> 5)) & 0x003F003Fu;
    return (t | (t >> 10)) & 0xFFFu;
}

struct PackOut { u64 word; u32 bad; };   // bad: 0 or 0x10000 | pos_in_word << 8 | char

// Pack the 20 bases whose first byte is x[0]'s byte a (= off & 3); nvalid of them belong to the sequence.
__device__ __forceinline__ PackOut pack20(const u32 (&x)[6], u32 a, u32 nvalid) {
    PackOut o;
    const u32 sel = 0x0123u + a * 0x1111u;     // align + byte-reverse in one PRMT
    u32 d[5], g4[5], v[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        v[j] = __byte_perm(x[j], x[j + 1], sel);
        g4[j] = pack4(v[j], d[j]);
    }
    u32 hi = (g4[0] << 16) | (g4[1] << 4) | (g4[2] >> 8);
    u32 lo = (g4[2] << 24) | (g4[3] << 12) | g4[4];
    u64 word = ((u64)hi << 32) | lo;
    if (nvalid < 20u) {
        word &= ~0ull << (3u * (20u - nvalid));
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            // bytes of group j that belong to the sequence: the top clamp(nvalid - 4j, 0, 4)
            int vj = (int)nvalid - 4 * j;
            vj = max(0, min(4, vj));
            d[j] &= __funnelshift_lc(0u, 0xFFFFFFFFu, 32u - 8u * (u32)vj);
        }
    }
    o.word = word;
    o.bad = 0;
    if (d[0] | d[1] | d[2] | d[3] | d[4]) {
#pragma unroll
        for (int j = 4; j >= 0; --j) {
            if (d[j]) {
                u32 byte = __clz(d[j]) >> 3;                       // 0 = first base of the group
                u32 ch = (v[j] >> (24 - 8 * byte)) & 0xFFu;
                o.bad = 0x10000u | ((4 * j + byte) << 8) | ch;
            }
        }
    }
    return o;
}

// generic (slow) fetch of 24 bytes at window-local offset off & ~3: shared memory when staged, else global
__device__ __noinline__ void fetch24_slow(const u8* win, const u8* raw, u32 base, u32 n, u32 off, u32 (&x)[6]) {
    const u32 o4 = off & ~3u;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        u32 v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const u32 lp = o4 + 4 * i + b;
            u32 c = 0;
            if (lp < PP_WINDOW) c = win[lp];
            else { const u64 g = (u64)base + lp; if (g < n) c = raw[g]; }
            v |= c << (8 * b);
        }
        x[i] = v;
    }
}

__device__ __forceinline__ u32 find_nl(const u64* mask64, u32 pos, const u8* raw, u32 base, u32 n) {
    if (pos < PP_WINDOW) {
        u32 w = pos >> 6;
        u64 m = mask64[w] & (~0ull << (pos & 63u));
        for (;;) {
            if (m) return w * 64u + (u32)__ffsll((long long)m) - 1u;
            if (++w >= PP_NW) break;
            m = mask64[w];
        }
        pos = PP_WINDOW;
    }
    u64 g = (u64)base + pos;
    while (g < n) {
        if (raw[g] == '\n') return (u32)(g - base);
        ++g;
    }
    return PP_NONE;
}

// Validation of one record whose first byte is at window-local offset start_l and whose line ends are e0..e3
// (PP_NONE = not found before the end of the chunk).  Produces the queue entry:
//   qoff = window-local offset of the sequence (PP_NONE: nothing to pack), qlen = bases | fast-path flag << 31
template <int LPR>
__device__ __forceinline__ void finish_record(const ParseParams& p, const u8* win, u32 base, u64 slot, u32 R, u32 start_l,
                                              u32 e0, u32 e1, u32 e2, u32 e3, u32& qoff, u32& qlen) {
    qoff = PP_NONE; qlen = 0;
    const u32 elast = (LPR == 4) ? e3 : e1;
    if (elast == PP_NONE) return;                       // incomplete record: left to the next chunk
    const u8 lead = (LPR == 4) ? '@' : '>';
    const u32 c0 = start_l < PP_WINDOW ? win[start_l] : p.raw[(u64)base + start_l];
    if (c0 != lead) {
        atomicMin(&p.ctl->err_parse, ((u64)R << 16) | ((u64)PERR_BAD_START << 8) | c0);
    } else if (LPR == 4 && (e1 - e0) != (e3 - e2)) {
        atomicMin(&p.ctl->err_parse, ((u64)R << 16) | ((u64)PERR_LEN_MISMATCH << 8));
    } else {
        const u32 nb = e1 - e0 - 1u;
        if (slot >= p.key_capacity) { p.ctl->too_long = 2; return; }
        if (nb > p.W * BASES_PER_WORD) { p.ctl->too_long = 1; return; }
        qoff = e0 + 1u;
        // fast path: every word of the row can be fetched from the staged window
        const u32 fast = (qoff + p.W * BASES_PER_WORD + 4u <= PP_WINDOW) ? 0x80000000u : 0u;
        qlen = nb | fast;
    }
}

template <int LPR>   // lines per record: 4 = FASTQ, 2 = FASTA
__global__ void __launch_bounds__(PP_THREADS, PP_MIN_CTAS) k_parse_pack(const ParseParams p) {
    __shared__ __align__(128) u8 win[PP_WINDOW];
    __shared__ __align__(16) u64 mask64[PP_NW];
    __shared__ u16 nlpos[PP_NLCAP];
    __shared__ u32 q_off[PP_QCAP];
    __shared__ u32 q_len[PP_QCAP];
    __shared__ uint2 s_hkey[PP_HKEYS];
    __shared__ u32 warp_sum[PP_THREADS / 32];
    __shared__ u32 s_tile, s_P, s_total, s_halo;
    __shared__ __align__(8) u64 mbar;

    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31u, warp = tid >> 5;
    const u64 slot_base = p.run->n_records;

    if (tid == 0) {
        s_tile = atomicAdd(&p.ctl->ticket, 1u);      // tiles are processed in ticket order: every predecessor
        mbar_init(&mbar, 1);                         // of a tile is already resident (look-back cannot deadlock)
    }
    if (tid < PP_HKEYS) s_hkey[tid] = pos_keys(p.hash_salt + tid);
    __syncthreads();
    const u32 tile = s_tile;
    const u32 base = tile * PP_TILE;
    const u32 valid = min(PP_WINDOW, p.n - base);
    if (tid == 0) {
        const u32 bytes = (valid + 15u) & ~15u;
        mbar_expect_tx(&mbar, bytes);
        bulk_g2s(win, p.raw + base, bytes, &mbar);
    }
    // constants the compiler must keep in registers (one LOP3 per use instead of two with immediates)
    u32 c_nl, c_7f;
    asm volatile("mov.u32 %0, 0x0A0A0A0A;" : "=r"(c_nl));
    asm volatile("mov.u32 %0, 0x7F7F7F7F;" : "=r"(c_7f));
    mbar_wait(&mbar, 0);

    // ---- 1. newline bitmask of the window
    {
        u16* m16 = reinterpret_cast<u16*>(mask64);
        const uint4* w4 = reinterpret_cast<const uint4*>(win);
        if (valid == PP_WINDOW) {
#pragma unroll
            for (u32 it = 0; it < PP_TILE / 16 / PP_THREADS; ++it) {
                const u32 u = tid + it * PP_THREADS;
                m16[u] = (u16)nl_mask16(w4[u], c_nl, c_7f);
            }
            if (tid < PP_HALO / 16) {
                const u32 u = tid + PP_TILE / 16;
                m16[u] = (u16)nl_mask16(w4[u], c_nl, c_7f);
            }
        } else {
            const u32 n_units = (valid + 15u) >> 4;
            for (u32 u = tid; u < PP_WINDOW / 16; u += PP_THREADS) {
                u32 m = 0;
                if (u < n_units) {
                    m = nl_mask16(w4[u], c_nl, c_7f);
                    u32 rem = valid - u * 16u;
                    if (rem < 16u) m &= (1u << rem) - 1u;
                }
                m16[u] = (u16)m;
            }
        }
    }
    __syncthreads();

    // ---- 2. local ranks: block scan over the tile's mask words (+ the halo words, ranked after them)
    const u64 my_mask = mask64[tid];               // PP_TILE/64 == PP_THREADS words cover the tile proper
    const u32 cnt = (u32)__popcll(my_mask);
    u32 incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (u32)d) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 ws = lane < PP_THREADS / 32 ? warp_sum[lane] : 0u;
        u32 wi = ws;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            u32 t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= (u32)d) wi += t;
        }
        if (lane < PP_THREADS / 32) warp_sum[lane] = wi - ws;      // exclusive warp offsets
        const u32 total = __shfl_sync(0xFFFFFFFFu, wi, PP_THREADS / 32 - 1);
        // publish this tile's aggregate as early as possible
        if (lane == 0) {
            if (tile == 0) st_volatile_u64(p.tile_state, (2ull << 32) | total);
            else st_volatile_u64(p.tile_state + tile, (1ull << 32) | total);
            s_total = total;
        }
        // halo words (PP_HALO/64 <= 32): ranks continue after the tile's
        const u64 hm = lane < (PP_NW - PP_THREADS) ? mask64[PP_THREADS + lane] : 0ull;
        const u32 hc = (u32)__popcll(hm);
        u32 hi = hc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 t = __shfl_up_sync(0xFFFFFFFFu, hi, d);
            if (lane >= (u32)d) hi += t;
        }
        const u32 htot = __shfl_sync(0xFFFFFFFFu, hi, 31);
        if (lane == 0) s_halo = htot;
        u32 r = total + hi - hc;
        u64 m = hm;
        while (m) {
            u32 b = (u32)__ffsll((long long)m) - 1u;
            m &= m - 1;
            if (r < PP_NLCAP) nlpos[r] = (u16)((PP_THREADS + lane) * 64u + b);
            ++r;
        }
    }
    __syncthreads();
    const u32 T = s_total;
    const u32 lex = warp_sum[warp] + (incl - cnt);      // local rank of my first newline
    {
        u64 m = my_mask;
        u32 r = lex;
        while (m) {
            u32 b = (u32)__ffsll((long long)m) - 1u;
            m &= m - 1;
            if (r < PP_NLCAP) nlpos[r] = (u16)(tid * 64u + b);
            ++r;
        }
    }
    // ---- decoupled look-back (warp 0) for the global rank of the tile's first newline
    if (warp == 0) {
        u32 P = 0;
        if (tile != 0) {
            // window of 128 predecessors per hop (4 per lane): one L2 round trip must cover more tiles than
            // the chip starts in that time, otherwise the distance to the nearest published prefix grows
            int look = (int)tile - 1;
            for (;;) {
                u64 s4[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    int idx = look - (int)lane - 32 * r;
                    s4[r] = (2ull << 32);
                    if (idx >= 0) s4[r] = ld_volatile_u64(p.tile_state + idx);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    int idx = look - (int)lane - 32 * r;
                    while ((s4[r] >> 32) == 0) s4[r] = ld_volatile_u64(p.tile_state + idx);
                }
                bool found = false;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (!found) {
                        u32 is_prefix = __ballot_sync(0xFFFFFFFFu, (s4[r] >> 32) == 2);
                        u32 val = (u32)s4[r];
                        if (is_prefix) {
                            u32 first = (u32)__ffs((int)is_prefix) - 1u;     // closest predecessor holding a prefix
                            if (lane > first) val = 0;
                            found = true;
                        }
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xFFFFFFFFu, val, d);
                        P += val;
                    }
                }
                if (found) break;
                look -= 128;
            }
            if (lane == 0) st_volatile_u64(p.tile_state + tile, (2ull << 32) | (u64)(P + T));
        }
        if (lane == 0) {
            s_P = P;
            if (tile == p.n_tiles - 1) {
                u32 all = P + T;
                p.ctl->n_newlines = all;
                u32 nrec = all / LPR;
                p.ctl->n_records = nrec < p.cap ? nrec : p.cap;
            }
        }
    }
    __syncthreads();
    const u32 P = s_P;
    const u32 WN = T + s_halo;                          // newlines in the whole window
    const bool dense = WN <= PP_NLCAP;

    // records owned by this tile: those whose preceding newline (rank LPR*R-1) lies in the tile proper
    u32 R_first = (P + LPR) / LPR;
    const u32 R_last = (P + T) / LPR;
    if (tile == 0) R_first = 0;
    const u32 n_owned = R_last >= R_first ? R_last - R_first + 1u : 0u;

    for (u32 rbase = 0; rbase < n_owned; rbase += PP_QCAP) {
        const u32 n_round = min(PP_QCAP, n_owned - rbase);
        // ---- 3. owners: geometry + validation
        if (dense) {
            // one thread per record: the record's line ends are consecutive entries of nlpos
            for (u32 o = tid; o < n_round; o += PP_THREADS) {
                const u32 R = R_first + rbase + o;
                const int j0 = (int)(LPR * R) - 1 - (int)P;          // local rank of the newline before the record
                const u32 start_l = j0 < 0 ? 0u : (u32)nlpos[j0] + 1u;
                const u32 gstart = base + start_l;
                u32 qoff = PP_NONE, qlen = 0;
                if (R <= p.cap) {
                    p.rec_start[R] = gstart;
                    if (p.dup && R < p.cap) p.dup[R] = 0;
                }
                if (R < p.cap && gstart < p.n) {
                    u32 e[4];
#pragma unroll
                    for (int k = 0; k < LPR; ++k) {
                        const u32 j = (u32)(j0 + 1 + k);
                        if (j < WN) e[k] = nlpos[j];
                        else {
                            const u32 prev = e[k > 0 ? k - 1 : 0];
                            const u32 from = k == 0 ? start_l : (prev == PP_NONE ? PP_NONE : prev + 1u);
                            e[k] = from == PP_NONE ? PP_NONE : find_nl(mask64, max(from, PP_WINDOW), p.raw, base, p.n);
                        }
                    }
                    finish_record<LPR>(p, win, base, slot_base + R, R, start_l, e[0], e[1], LPR == 4 ? e[2] : e[1],
                                       LPR == 4 ? e[3] : e[1], qoff, qlen);
                }
                q_off[o] = qoff;
                q_len[o] = qlen;
            }
        } else {
            // very dense tile (tiny records): every thread walks its own newline bits
            u64 m = my_mask;
            u32 k = P + lex;
            bool virt = (tile == 0 && tid == 0);      // record 0 starts at offset 0 with no newline before it
            while (m || virt) {
                u32 R, start_l;
                if (virt) { virt = false; R = 0; start_l = 0; }
                else {
                    u32 b = (u32)__ffsll((long long)m) - 1u;
                    m &= m - 1;
                    u32 kk = k++;
                    if ((kk + 1u) % LPR) continue;
                    R = (kk + 1u) / LPR;
                    start_l = tid * 64u + b + 1u;
                }
                const u32 gstart = base + start_l;
                if (rbase == 0 && R <= p.cap) {
                    p.rec_start[R] = gstart;
                    if (p.dup && R < p.cap) p.dup[R] = 0;
                }
                const u32 o = R - R_first;
                if (o < rbase || o >= rbase + PP_QCAP) continue;
                u32 qoff = PP_NONE, qlen = 0;
                if (R < p.cap && gstart < p.n) {
                    u32 e0 = find_nl(mask64, start_l, p.raw, base, p.n);
                    u32 e1 = e0 == PP_NONE ? PP_NONE : find_nl(mask64, e0 + 1u, p.raw, base, p.n);
                    u32 e3 = e1, e2 = e1;
                    if (LPR == 4) {
                        e2 = e1 == PP_NONE ? PP_NONE : find_nl(mask64, e1 + 1u, p.raw, base, p.n);
                        e3 = e2 == PP_NONE ? PP_NONE : find_nl(mask64, e2 + 1u, p.raw, base, p.n);
                    }
                    finish_record<LPR>(p, win, base, slot_base + R, R, start_l, e0, e1, e2, e3, qoff, qlen);
                }
                q_off[o - rbase] = qoff;
                q_len[o - rbase] = qlen;
            }
        }
        __syncthreads();

        // ---- 4. pack: 8 lanes per record
        {
            const u32 g8 = tid >> 3, l8 = tid & 7u;
            const u32 gmask = 0xFFu << (lane & 24u);
            for (u32 q = g8; q < n_round; q += PP_THREADS / 8) {
                const u32 off = q_off[q];
                if (off == PP_NONE) continue;
                const u32 ql = q_len[q];
                const u32 nb = ql & 0x7FFFFFFFu;
                const u32 R = R_first + rbase + q;
                u64* row = p.keys + (slot_base + R) * p.row_words + p.mate_off;
                u64 hsum = 0, w0 = 0;
                u32 bad = 0, badw = 0;
                for (u32 w = l8; w < p.W; w += 8) {
                    const u32 done = w * BASES_PER_WORD;
                    const u32 pos = off + done;
                    const u32 nvalid = nb > done ? min(nb - done, (u32)BASES_PER_WORD) : 0u;
                    u32 x[6];
                    if (ql >> 31) {          // staged in shared memory
                        const u32* w32 = reinterpret_cast<const u32*>(win) + (pos >> 2);
#pragma unroll
                        for (int i = 0; i < 6; ++i) x[i] = w32[i];
                    } else {
                        fetch24_slow(win, p.raw, base, p.n, pos, x);
                    }
                    const PackOut po = pack20(x, pos & 3u, nvalid);
                    row[w] = po.word;
                    if (w == 0) w0 = po.word;
                    const uint2 hk = w < PP_HKEYS ? s_hkey[w] : pos_keys(p.hash_salt + w);
                    hsum += word_hash(po.word, hk);
                    if (po.bad && !bad) { bad = po.bad; badw = w; }
                }
                hsum += __shfl_xor_sync(gmask, hsum, 1);
                hsum += __shfl_xor_sync(gmask, hsum, 2);
                hsum += __shfl_xor_sync(gmask, hsum, 4);
                if (l8 == 0) {
                    p.hash[R] = hsum;
                    if (p.seq_len) p.seq_len[R] = nb;
                    if (p.word0) p.word0[R] = w0;
                }
                if (bad) {
                    if (p.bad_rec) atomicMin(&p.bad_rec[R], ((badw * BASES_PER_WORD + ((bad >> 8) & 0xFFu)) << 8) | (bad & 0xFFu));
                    if (p.strict) {
                        u32 pos = badw * BASES_PER_WORD + ((bad >> 8) & 0xFFu);
                        atomicMin(&p.ctl->err_base, ((u64)R << 32) | ((u64)pos << 8) | (bad & 0xFFu));
                    } else {
                        p.ctl->pad = 1;   // non-ACGTN byte seen in a mode that accepts any byte
                    }
                }
            }
        }
        __syncthreads();
    }
}

__global__ void k_init_chunk(ChunkCtl* ctl, u64* tile_state, u32 n_tiles) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        ctl->ticket = 0; ctl->n_newlines = 0; ctl->n_records = 0; ctl->consumed = 0;
        ctl->err_parse = NO_ERR; ctl->err_base = NO_ERR; ctl->too_long = 0; ctl->pad = 0;
    }
    for (; i < n_tiles; i += gridDim.x * blockDim.x) tile_state[i] = 0;
}


}  // namespace fqd
