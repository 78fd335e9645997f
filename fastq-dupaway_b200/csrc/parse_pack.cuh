// parse_pack.cuh - K1: fused record splitter + sequence packer + key hasher (one pass over the raw bytes).
//
// Replaces, per record, FastqView::read_new / FastaView::read_new (src/fastqview.cpp:89-119,
// src/fastaview.cpp:75-93: 4 or 2 newline searches, '@'/'>' check, len(seq)==len(qual)) and
// SeqUtils::seq2hash (src/seq_utils.cpp:35-49) for a whole chunk of raw bytes resident in HBM.
//
// One CTA handles one 16 KiB tile (+1 KiB halo) staged in shared memory by a 1-D bulk async copy (TMA).
//   1. newline bitmask of the window (SIMD-in-register compare, 16 B per step)
//   2. block scan of per-thread newline counts, decoupled look-back across tiles -> global rank of every '\n'
//      (the rank modulo lines-per-record tells which newline ends a record: the reference, too, simply counts
//      4 (2) newlines per record)
//   3. the thread that sees the newline preceding a record start owns that record: it finds the record's
//      line ends in the bitmask, validates it and queues (sequence offset, length)
//   4. groups of 8 lanes pack one queued sequence each into 3-bit codes (20 bases per 64-bit word, see
//      common.cuh), write the key row with coalesced 64-bit stores and reduce the key hash.
#pragma once
#include "common.cuh"

namespace fqd {

constexpr int PP_THREADS = 256;
constexpr u32 PP_TILE = 16384;
constexpr u32 PP_HALO = 1024;
constexpr u32 PP_WINDOW = PP_TILE + PP_HALO;
constexpr u32 PP_NW = PP_WINDOW / 64;          // 64-bit mask words per window
constexpr u32 PP_QCAP = 256;                   // records packed per round
constexpr u32 PP_NONE = 0xFFFFFFFFu;

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
    u64* hash;              // [cap] per-mate key hash, chunk-local index
    u32* seq_len;           // optional [cap]: sequence length in bases (chunk-local index)
    u64* word0;             // optional [cap]: first key word (radix-sort key), chunk-local index
    u8* dup;                // optional [cap]: duplicate flags, cleared here for every record of the chunk
    u32 strict;             // 1: bytes outside {A,C,G,T,N} are an error (fast mode, src/seq_utils.cpp:17-19)
    u32 hash_salt;          // distinguishes mates in the position keys
};

// 16 input bytes -> 16-bit mask of '\n' positions (bit j <-> byte j).
__device__ __forceinline__ u32 nl_flags(u32 w) {
    // exact per-byte zero test of (w ^ 0x0A0A0A0A): flag in bit 7 of each matching byte
    u32 a = (w ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu;
    u32 s = a + 0x7F7F7F7Fu;
    return ~s & ~w & 0x80808080u;
}
__device__ __forceinline__ u32 nl_mask16(uint4 v) {
    u32 p0 = nl_flags(v.x) * 0x00204081u;   // gathers the 4 flags into bits 28..31
    u32 p1 = nl_flags(v.y) * 0x00204081u;
    u32 p2 = nl_flags(v.z) * 0x00204081u;
    u32 p3 = nl_flags(v.w) * 0x00204081u;
    return (p0 >> 28) | ((p1 >> 24) & 0xF0u) | ((p2 >> 20) & 0xF00u) | ((p3 >> 16) & 0xF000u);
}

struct PackOut { u64 word; u32 bad; };   // bad: 0 or (pos_in_word << 8 | char) + 0x10000

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

// Pack up to 20 bases starting at window-local offset `off` (nvalid of them belong to the sequence).
__device__ __forceinline__ PackOut pack_word(const u8* win, const u8* raw, u32 base, u32 n, u32 off, u32 nvalid) {
    PackOut o; o.word = 0; o.bad = 0;
    if (nvalid == 0) return o;
    u32 x[6];
    u32 a = off & 3u;
    if (off + 24u <= PP_WINDOW) {
        const u32* w32 = reinterpret_cast<const u32*>(win) + (off >> 2);
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = w32[i];
    } else {   // sequence runs past the staged window: gather bytes from global memory
        u64 g = (u64)base + (off & ~3u);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            u32 v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                u64 p = g + 4 * i + b;
                u32 c = (p < n) ? raw[p] : 0u;
                v |= c << (8 * b);
            }
            x[i] = v;
        }
    }
    u32 sel = 0x0123u + a * 0x1111u;     // align + byte-reverse in one PRMT
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
            int vj = (int)nvalid - 4 * j;
            u32 m = vj <= 0 ? 0u : (vj >= 4 ? 0xFFFFFFFFu : (0xFFFFFFFFu << (8 * (4 - vj))));
            d[j] &= m;
        }
    }
    o.word = word;
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

// Geometry + validation of one record whose first byte is at window-local offset start_l and whose LPR
// line ends are e[0..LPR) (PP_NONE = not found before the end of the chunk).  Returns the queue entry.
template <int LPR>
__device__ __forceinline__ void finish_record(const ParseParams& p, const u8* win, u32 base, u32 R, u32 start_l,
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
        qoff = e0 + 1u;
        qlen = e1 - e0 - 1u;
    }
}

constexpr u32 PP_NLCAP = 1024;     // newline positions compacted per window; denser tiles take the slow path
constexpr int PP_MIN_CTAS = 5;

template <int LPR>   // lines per record: 4 = FASTQ, 2 = FASTA
__global__ void __launch_bounds__(PP_THREADS, PP_MIN_CTAS) k_parse_pack(const ParseParams p) {
    __shared__ __align__(128) u8 win[PP_WINDOW];
    __shared__ __align__(16) u64 mask64[PP_NW];
    __shared__ u16 nlpos[PP_NLCAP];
    __shared__ u32 q_off[PP_QCAP];
    __shared__ u32 q_len[PP_QCAP];
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
    __syncthreads();
    const u32 tile = s_tile;
    {
        const u32 base = tile * PP_TILE;
        const u32 valid = min(PP_WINDOW, p.n - base);
        if (tid == 0) {
            const u32 bytes = (valid + 15u) & ~15u;
            mbar_expect_tx(&mbar, bytes);
            bulk_g2s(win, p.raw + base, bytes, &mbar);
        }
        mbar_wait(&mbar, 0);

        // ---- 1. newline bitmask of the window
        {
            const u32 n_units = (valid + 15u) >> 4;
            u16* m16 = reinterpret_cast<u16*>(mask64);
#pragma unroll
            for (u32 it = 0; it < (PP_WINDOW / 16 + PP_THREADS - 1) / PP_THREADS; ++it) {
                const u32 u = tid + it * PP_THREADS;
                if (u < PP_WINDOW / 16) {
                    u32 m = 0;
                    if (u < n_units) {
                        uint4 v = reinterpret_cast<const uint4*>(win)[u];
                        m = nl_mask16(v);
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
                                const u32 from = k == 0 ? start_l : (e[k - 1] == PP_NONE ? PP_NONE : e[k - 1] + 1u);
                                e[k] = from == PP_NONE ? PP_NONE : find_nl(mask64, max(from, PP_WINDOW), p.raw, base, p.n);
                            }
                        }
                        finish_record<LPR>(p, win, base, R, start_l, e[0], e[1], LPR == 4 ? e[2] : e[1], LPR == 4 ? e[3] : e[1], qoff, qlen);
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
                        finish_record<LPR>(p, win, base, R, start_l, e0, e1, e2, e3, qoff, qlen);
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
                    const u32 nb = q_len[q];
                    const u32 R = R_first + rbase + q;
                    const u64 slot = slot_base + R;
                    if (slot >= p.key_capacity) { if (l8 == 0) p.ctl->too_long = 2; continue; }
                    if (nb > p.W * BASES_PER_WORD) { if (l8 == 0) p.ctl->too_long = 1; continue; }
                    u64* row = p.keys + slot * p.row_words + p.mate_off;
                    u64 hsum = 0, w0 = 0;
                    u32 bad = 0, badw = 0;
                    for (u32 w = l8; w < p.W; w += 8) {
                        u32 done = w * BASES_PER_WORD;
                        u32 nvalid = nb > done ? min(nb - done, (u32)BASES_PER_WORD) : 0u;
                        PackOut po = pack_word(win, p.raw, base, p.n, off + done, nvalid);
                        row[w] = po.word;
                        if (w == 0) w0 = po.word;
                        hsum += word_hash(po.word, pos_key(p.hash_salt + w));
                        if (po.bad && !bad) { bad = po.bad; badw = w; }
                    }
                    hsum += __shfl_xor_sync(gmask, hsum, 1);
                    hsum += __shfl_xor_sync(gmask, hsum, 2);
                    hsum += __shfl_xor_sync(gmask, hsum, 4);
                    if (l8 == 0) {
                        p.hash[R] = mix64(hsum);
                        if (p.seq_len) p.seq_len[R] = nb;
                        if (p.word0) p.word0[R] = w0;
                    }
                    if (bad && p.strict) {
                        u32 pos = badw * BASES_PER_WORD + ((bad >> 8) & 0xFFu);
                        atomicMin(&p.ctl->err_base, ((u64)R << 32) | ((u64)pos << 8) | (bad & 0xFFu));
                    } else if (bad) {
                        p.ctl->pad = 1;   // non-ACGTN byte seen in a mode that accepts any byte
                    }
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace fqd
