// parse_pack.cuh - K1: fused record splitter + sequence packer + key hasher (one pass over the raw bytes).
//
// Replaces, per record, FastqView::read_new / FastaView::read_new (src/fastqview.cpp:89-119,
// src/fastaview.cpp:75-93: 4 or 2 newline searches, '@'/'>' check, len(seq)==len(qual)) and
// SeqUtils::seq2hash (src/seq_utils.cpp:35-49) for a whole chunk of raw bytes resident in HBM.
//
// One CTA handles one 16 KiB tile (+1 KiB halo) staged in shared memory by a 1-D bulk async copy (TMA); tiles are
// taken in block-index order.
//   0. while the tile's bytes are in flight the CTA counts the newlines of the tile PP_LEAD ahead ("scout") and publishes
//      the count; later warp 0 resolves that tile's prefix (decoupled look-back, 128 predecessors per hop) - so P, the
//      number of newlines before the CTA's own tile, was resolved two CTA lifetimes before the CTA started
//   1. newline bitmask of the window (SIMD-in-register byte compare, dp4a gathers the flags)
//   2. block scan of per-thread newline counts -> local rank of every '\n'; positions compacted by rank into shared memory
//   3. P modulo lines-per-record says which newline ends a record (the reference, too, simply counts 4 (2) newlines per
//      record), P / lines-per-record is the index of the tile's first record.  Groups of 4 lanes take one record each
//      (its first byte lies in the tile): line ends = consecutive entries of the compacted positions, validation, then
//      the sequence is packed into 3-bit codes (20 bases per 64-bit word, see common.cuh; or raw bytes, 8 per word:
//      template parameter BYTES) and committed: record offset, errors, 128-bit row stores, the multilinear key hash.
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
    u32 skip;               // < 16: the chunk proper starts at raw + skip (raw is aligned down for the bulk copies; the bytes
                            //       before belong to the record before); record 0 starts there
    u32 byte_keys;          // 1: rows hold the raw bytes of sequence + '\n', 8 per word, first byte in the top bits
                            //    (sequence-based modes on inputs with bytes outside {A,C,G,T,N}); 0: 3-bit codes, 20 per word
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

// slow path of the packer: the 24 source bytes are not all inside the staged window
__device__ __noinline__ PackOut pack_word_slow(const u8* win, const u8* raw, u32 base, u32 n, u32 pos, u32 nvalid) {
    u32 x[6];
    fetch24_slow(win, raw, base, n, pos, x);
    return pack20(x, pos & 3u, nvalid);
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

template <int ID, int N> __device__ __forceinline__ void bar_sync_c() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
template <int ID, int N> __device__ __forceinline__ void bar_arrive_c() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
constexpr u32 PP_PACKERS = PP_THREADS - 32;        // threads of warps 1..7
constexpr u32 PP_GROUP = 4;                        // lanes that pack one record (two key words each per 8 words)
enum { RS_NONE = 0, RS_OK = 1, RS_BAD_START = 2, RS_LEN_MISMATCH = 3, RS_TOO_LONG = 4 };

// Geometry + validation of one record whose first byte is at window-local offset start_l and whose line ends are
// e0..e3 (PP_NONE = not found before the end of the chunk).  Nothing here depends on the record's index, so it can
// run before the look-back has delivered it.  Returns RS_* | first byte << 8;
//   qoff = window-local offset of the sequence, qlen = bases | fast-path flag << 31   (RS_OK only)
template <int LPR, bool BYTES>
__device__ __forceinline__ u32 classify_record(const ParseParams& p, const u8* win, u32 base, u32 start_l,
                                               u32 e0, u32 e1, u32 e2, u32 e3, u32& qoff, u32& qlen) {
    qoff = PP_NONE; qlen = 0;
    const u32 elast = (LPR == 4) ? e3 : e1;
    if (elast == PP_NONE) return RS_NONE;               // incomplete record: left to the next chunk
    const u8 lead = (LPR == 4) ? '@' : '>';
    const u32 c0 = start_l < PP_WINDOW ? win[start_l] : p.raw[(u64)base + start_l];
    if (c0 != lead) return RS_BAD_START | (c0 << 8);
    if (LPR == 4 && (e1 - e0) != (e3 - e2)) return RS_LEN_MISMATCH;
    const u32 nb = e1 - e0 - 1u;
    const u32 room = BYTES ? p.W * 8u - 1u : p.W * BASES_PER_WORD;            // byte keys include the '\n'
    if (nb > room) return RS_TOO_LONG;
    qoff = e0 + 1u;
    // fast path: every word of the row can be fetched from the staged window
    const u32 fast = (qoff + room + 8u <= PP_WINDOW) ? 0x80000000u : 0u;
    qlen = nb | fast;
    return RS_OK;
}

// What the owner thread of record R does once R is known (src/fastqview.cpp:121-138 error precedence).
__device__ __forceinline__ void commit_record(const ParseParams& p, u64 slot_base, u32 R, u32 gstart, u32 status) {
    if (R <= p.cap) {
        p.rec_start[R] = gstart;
        if (p.dup && R < p.cap) p.dup[R] = 0;
    }
    if (R >= p.cap) return;
    const u32 rs = status & 0xFFu;
    if (rs == RS_BAD_START) atomicMin(&p.ctl->err_parse, ((u64)R << 16) | ((u64)PERR_BAD_START << 8) | (status >> 8));
    else if (rs == RS_LEN_MISMATCH) atomicMin(&p.ctl->err_parse, ((u64)R << 16) | ((u64)PERR_LEN_MISMATCH << 8));
    else if (rs == RS_OK || rs == RS_TOO_LONG) {
        if (slot_base + R >= p.key_capacity) atomicOr(&p.ctl->too_long, (u32)TL_CAPACITY);
        else if (rs == RS_TOO_LONG) { atomicOr(&p.ctl->too_long, (u32)TL_SEQ); atomicMin(&p.ctl->too_long_rec, R); }
    }
}

// One 20-base key word of the sequence at window-local offset off (nb bases).  The last, partial word of a
// sequence is built from the LAST 20 bases of the sequence and shifted, so that no byte beyond the sequence is ever
// looked at (no masking of the validity check); only sequences shorter than 20 bases take the masked path.
//   bad: 0, or 0x80000000 | position in the sequence << 8 | offending byte
// byte keys: word w = bytes [8w, 8w + 8) of sequence + '\n' (the '\n' is the one in the file), zero padded.  Unsigned
// word order is then the byte order FastqView::cmp defines for ANY byte (src/fastqview.cpp:56-67).
__device__ __forceinline__ u64 pack_word_bytes(const u8* win, const ParseParams& p, u32 base, u32 off, u32 ql, u32 w) {
    const u32 nb = (ql & 0x7FFFFFFFu) + 1u;           // symbols incl. the terminator
    const u32 first = w * 8u;
    if (nb <= first) return 0ull;
    const u32 pos = off + first;
    u32 x[3];
    if (ql >> 31) {
        const u32* w32 = reinterpret_cast<const u32*>(win) + (pos >> 2);
        x[0] = w32[0]; x[1] = w32[1]; x[2] = w32[2];
    } else {
        u32 y[6];
        fetch24_slow(win, p.raw, base, p.n, pos, y);
        x[0] = y[0]; x[1] = y[1]; x[2] = y[2];
    }
    const u32 sel = 0x0123u + (pos & 3u) * 0x1111u;   // align + byte-reverse
    const u32 hi = __byte_perm(x[0], x[1], sel), lo = __byte_perm(x[1], x[2], sel);
    u64 word = ((u64)hi << 32) | lo;
    const u32 valid = nb - first;
    const u64 vmask = valid < 8u ? ~0ull << (8u * (8u - valid)) : ~0ull;
    word &= vmask;
    // a byte below the line feed (0x0A) breaks "a prefix sorts before its extensions": tell the host side, the loose
    // scan then runs the reference's literal loop (no byte >= 0x80 is flagged: ~word clears those)
    const u64 t = (word | 0x8080808080808080ull) - 0x0A0A0A0A0A0A0A0Aull;
    if (~t & ~word & 0x8080808080808080ull & vmask) p.ctl->pad = 2;
    return word;
}

template <bool BYTES>
__device__ __forceinline__ u64 pack_word(const u8* win, const ParseParams& p, u32 base, u32 off, u32 ql, u32 w, u32& bad) {
    const u32 nb = ql & 0x7FFFFFFFu;
    const u32 done = w * BASES_PER_WORD;
    bad = 0;
    if (BYTES) return pack_word_bytes(win, p, base, off, ql, w);
    if (nb <= done) return 0ull;
    u32 nvalid = min(nb - done, (u32)BASES_PER_WORD);
    u32 first = done, shift = 0;
    if (nvalid < BASES_PER_WORD && nb >= BASES_PER_WORD) {
        first = nb - BASES_PER_WORD;
        shift = 3u * (BASES_PER_WORD - nvalid);
        nvalid = BASES_PER_WORD;
    }
    const u32 pos = off + first;
    PackOut po;
    if (ql >> 31) {          // staged in shared memory
        const u32* w32 = reinterpret_cast<const u32*>(win) + (pos >> 2);
        u32 x[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = w32[i];
        po = pack20(x, pos & 3u, nvalid);
    } else {
        po = pack_word_slow(win, p.raw, base, p.n, pos, nvalid);
    }
    if (po.bad) bad = 0x80000000u | ((first + ((po.bad >> 8) & 0xFFu)) << 8) | (po.bad & 0xFFu);
    return (po.word << shift) & 0x0FFFFFFFFFFFFFFFull;
}

// Per-record results of a pack group: PP_GROUP lanes reduce the key hash; lane 0 writes the per-record tables.
__device__ __forceinline__ void commit_group(const ParseParams& p, u32 R, u32 nb, u32 l4, u32 gmask, u64 hsum, u64 w0, u32 bad) {
    hsum += __shfl_xor_sync(gmask, hsum, 1);
    hsum += __shfl_xor_sync(gmask, hsum, 2);
    if (l4 == 0) {
        p.hash[R] = hsum;
        if (p.seq_len) p.seq_len[R] = nb;
        if (p.word0) p.word0[R] = w0;
    }
    if (bad) {
        const u32 v = bad & 0x7FFFFFFFu;               // position in the sequence << 8 | byte
        if (p.bad_rec) atomicMin(&p.bad_rec[R], v);
        if (p.strict) atomicMin(&p.ctl->err_base, ((u64)R << 32) | v);
        else p.ctl->pad = 1;                           // non-ACGTN byte seen in a mode that accepts any byte
    }
}

// pos_keys(mate * 4096 + w) for w < PP_HKEYS, both mates (filled once by pp_init_tables)
__constant__ uint2 c_hkeys[2 * PP_HKEYS];
static inline cudaError_t pp_init_tables() {
    uint2 h[2 * PP_HKEYS];
    for (u32 m = 0; m < 2; ++m)
        for (u32 w = 0; w < PP_HKEYS; ++w) h[m * PP_HKEYS + w] = pos_keys(m * 4096u + w);
    return cudaMemcpyToSymbol(c_hkeys, h, sizeof(h));
}

// ---- build-time experiments (profiles/r02_k1_summary.md): FQD_K1_LEAD=<tiles> scout distance, FQD_K1_TIMELINE records
// per-phase clocks of every TL_STRIDE-th tile.
#ifndef FQD_K1_LEAD
#define FQD_K1_LEAD 1480     // = 2 x the CTAs resident on 148 SMs (PP_MIN_CTAS each): the tile that starts about two CTA lifetimes from now
#endif
constexpr u32 PP_LEAD = FQD_K1_LEAD;        // the tile whose newlines this CTA counts
constexpr u32 PP_PLEAD = PP_LEAD / 2;       // the tile whose prefix this CTA resolves: every count it needs is a CTA lifetime old
#ifndef FQD_K1_OWNERS
#define FQD_K1_OWNERS 0
#endif
#ifdef FQD_K1_TIMELINE
constexpr u32 TL_STRIDE = 61, TL_SLOTS = 12, TL_CAP = 4096;
__device__ long long g_k1_timeline[TL_CAP * TL_SLOTS];
#define TL_STAMP(cond, k) do { if ((cond) && tile % TL_STRIDE == 0 && tile / TL_STRIDE < TL_CAP) g_k1_timeline[(tile / TL_STRIDE) * TL_SLOTS + (k)] = clock64(); } while (0)
#define TL_ENTRY() const long long tl_t0 = clock64()
#define TL_STAMP0(cond) do { if ((cond) && tile % TL_STRIDE == 0 && tile / TL_STRIDE < TL_CAP) g_k1_timeline[(tile / TL_STRIDE) * TL_SLOTS] = tl_t0; } while (0)
#else
#define TL_STAMP(cond, k) do { } while (0)
#define TL_ENTRY() do { } while (0)
#define TL_STAMP0(cond) do { } while (0)
#endif

// Decoupled look-back (one warp): exclusive prefix of tile > 0 over the published per-tile newline counts.
// tile_state[t] = flag << 32 | value; flag 1: value = count of tile t, flag 2: value = count of tiles 0..t.
// 128 predecessors per hop (4 per lane): one L2 round trip must cover more tiles than the chip starts in that time.
__device__ __forceinline__ u32 pp_lookback(const u64* tile_state, u32 tile, u32 lane) {
    u32 P = 0;
    int look = (int)tile - 1;
    for (;;) {
        u64 s4[4];
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
            int idx = look - (int)lane - 32 * r4;
            s4[r4] = (2ull << 32);
            if (idx >= 0) s4[r4] = ld_volatile_u64(tile_state + idx);
        }
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
            int idx = look - (int)lane - 32 * r4;
            while ((s4[r4] >> 32) == 0) s4[r4] = ld_volatile_u64(tile_state + idx);
        }
        bool found = false;
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
            if (!found) {
                u32 is_prefix = __ballot_sync(0xFFFFFFFFu, (s4[r4] >> 32) == 2);
                u32 val = (u32)s4[r4];
                if (is_prefix) {
                    u32 first = (u32)__ffs((int)is_prefix) - 1u;     // closest predecessor holding a prefix
                    if (lane > first) val = 0;
                    found = true;
                }
                P += __reduce_add_sync(0xFFFFFFFFu, val);
            }
        }
        if (found) return P;
        look -= 128;
    }
}

// LPR = lines per record: 4 = FASTQ, 2 = FASTA.  BYTES = raw-byte key rows (ParseParams::byte_keys; sequence-based
// modes on arbitrary alphabets) - a template parameter so that the 3-bit instantiation every other path uses carries
// none of its code or registers.
//
// One CTA per tile, tiles in block-index order (the hardware starts the CTAs of a 1-D grid in index order, so every
// predecessor of a tile is resident or done when the tile starts - the assumption every single-pass look-back scan makes).
// The one thing a tile needs from its predecessors is P, the number of newlines before it.  Round 1 published a tile's
// count after its own scan and hid the look-back behind a guess of P mod LPR; the packers still waited 3 500 cycles (of
// a 14 600-cycle CTA life) for the slowest of ~400 neighbours, and the look-back itself never takes less than ~4 000
// cycles (three L2 round trips: the nearest resolved prefix is as far away as the look-back is long).  So nobody waits
// for it any more - every CTA works for a tile PP_LEAD ahead of its own (about two CTA lifetimes):
//   * while its own tile is on its way through the TMA, all 8 warps count the newlines of that later tile ("scout":
//     plain 16-byte loads, which also pull the tile into L2) and publish the count;
//   * after the barriers of its own tile, warp 0 resolves that later tile's prefix by decoupled look-back and
//     publishes it, while warps 1..7 split and pack the CTA's own tile.
// A tile therefore finds its own prefix resolved long before it starts (one load), splits with the exact P - no guess, no
// second round, no wait - and every spin in this kernel is on something published by a CTA that started earlier (the
// first PP_LEAD tiles are resolved by k_scout_head / k_scan_head before the kernel starts): deadlock-free whatever the
// number of resident CTAs.
template <int LPR, bool BYTES = false>
__global__ void __launch_bounds__(PP_THREADS, PP_MIN_CTAS) k_parse_pack(const ParseParams p) {
    __shared__ __align__(128) u8 win[PP_WINDOW];
    __shared__ __align__(16) u64 mask64[PP_NW];
    __shared__ u16 nlpos[PP_NLCAP];
    __shared__ u32 q_off[PP_QCAP];
    __shared__ u32 q_len[PP_QCAP];
    __shared__ uint2 s_hkey[PP_HKEYS];
    __shared__ u32 warp_sum[PP_THREADS / 32], s_scout[PP_THREADS / 32];
    __shared__ u32 s_P, s_halo;
    __shared__ __align__(8) u64 mbar;

    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31u, warp = tid >> 5;
    const u64 slot_base = p.run->n_records;
    TL_ENTRY();
    const u32 tile = blockIdx.x;
    const u32 base = tile * PP_TILE;
    const u32 valid = min(PP_WINDOW, p.n - base);
    const u32 stile = tile + PP_LEAD;                   // the tile I count for its future CTA
    const bool scouting = stile < p.n_tiles;
    // constants the compiler must keep in registers (one LOP3 per use instead of two with immediates)
    u32 c_nl, c_7f;
    asm volatile("mov.u32 %0, 0x0A0A0A0A;" : "=r"(c_nl));
    asm volatile("mov.u32 %0, 0x7F7F7F7F;" : "=r"(c_7f));

    // ---- 0. own tile: one bulk async copy (TMA); scout tile: 64 bytes per thread, loads in flight next to it
    if (tid == 0) {
        mbar_init(&mbar, 1);
        const u32 bytes = (valid + 15u) & ~15u;
        mbar_expect_tx(&mbar, bytes);
        bulk_g2s(win, p.raw + base, bytes, &mbar);
    }
    // scout loads: 16 bytes per thread and step, consecutive threads on consecutive units (any assignment will do, only
    // the COUNT matters).  They are not looked at before this CTA's own mask and scan are done: a cold-DRAM latency
    // that nobody waits for.
    uint4 sv[4];
    const u64 sbase = (u64)stile * PP_TILE + tid * 16u;          // 64-bit: a chunk may end within 4 GiB of 2^32
    const bool scout_full = (u64)(stile + 1u) * PP_TILE <= (u64)p.n;      // every tile but the chunk's last one
    if (scout_full) {
#pragma unroll
        for (int k = 0; k < 4; ++k) sv[k] = __ldcg(reinterpret_cast<const uint4*>(p.raw + sbase) + k * PP_THREADS);
    } else if (scouting) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u64 o = sbase + (u64)k * (PP_THREADS * 16u);
            sv[k] = make_uint4(0u, 0u, 0u, 0u);
            if (o < (u64)p.n) sv[k] = __ldcg(reinterpret_cast<const uint4*>(p.raw + o));      // like the bulk copy, the last unit may reach past n
        }
    }
    if (tid < PP_HKEYS) s_hkey[tid] = c_hkeys[((p.hash_salt >> 12) & 1u) * PP_HKEYS + tid];
    if (warp == 0) { __syncwarp(); mbar_wait(&mbar, 0); }
    __syncthreads();                                    // the window is staged; s_P (scouted tiles), s_hkey visible
    TL_STAMP0(tid == 0);
    TL_STAMP(tid == 0, 1);

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
    TL_STAMP(tid == 0, 2);

    // ---- 2. local ranks: block scan over the tile's mask words (+ the halo words, ranked after them)
    // PP_TILE/64 == PP_THREADS words cover the tile proper; newlines before the chunk's first byte are not ours
    const u64 my_mask = mask64[tid] & ((tile == 0 && tid == 0) ? (~0ull << p.skip) : ~0ull);
    const u32 cnt = (u32)__popcll(my_mask);
    u32 incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (u32)d) incl += x;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    u32 T, lex;                                     // newlines in the tile proper; local rank of my first newline
    {
        const u32 ws = lane < PP_THREADS / 32 ? warp_sum[lane] : 0u;
        u32 wi = ws;
#pragma unroll
        for (int d = 1; d < PP_THREADS / 32; d <<= 1) {
            u32 x = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= (u32)d) wi += x;
        }
        T = __shfl_sync(0xFFFFFFFFu, wi, PP_THREADS / 32 - 1);
        lex = __shfl_sync(0xFFFFFFFFu, wi - ws, warp) + (incl - cnt);
    }
    TL_STAMP(tid == 0, 3);
    if (tid == 0) {
        // resolved two CTA lifetimes ago by the CTA that scouted this tile (the first PP_LEAD tiles: by pp_scout_head,
        // before this kernel started): inclusive prefix = P + T
        u64 own = ld_volatile_u64(p.tile_state + tile);
        while ((own >> 32) != 2) own = ld_volatile_u64(p.tile_state + tile);
        const u32 all = (u32)own;
        s_P = all - T;
        if (tile == p.n_tiles - 1) {
            p.ctl->n_newlines = all;
            const u32 nrec = all / LPR;
            p.ctl->n_records = nrec < p.cap ? nrec : p.cap;
        }
    }
    if (cnt) {
        // straight-line for the first two newlines of my 64 bytes (a FASTQ line pair "...\n+\n" at most), loop for more
        u64 m = my_mask;
        const u32 b0 = (u32)__ffsll((long long)m) - 1u;
        if (lex < PP_NLCAP) nlpos[lex] = (u16)(tid * 64u + b0);
        m &= m - 1;
        if (m) {
            const u32 b1 = (u32)__ffsll((long long)m) - 1u;
            if (lex + 1u < PP_NLCAP) nlpos[lex + 1u] = (u16)(tid * 64u + b1);
            m &= m - 1;
            u32 r = lex + 2u;
            while (m) {
                const u32 b = (u32)__ffsll((long long)m) - 1u;
                m &= m - 1;
                if (r < PP_NLCAP) nlpos[r] = (u16)(tid * 64u + b);
                ++r;
            }
        }
    }
    if (warp == PP_THREADS / 32 - 1) {
        // the last warp also ranks the halo words (PP_HALO/64 <= 32): their ranks continue after the tile's
        const u64 hm = lane < (PP_NW - PP_THREADS) ? mask64[PP_THREADS + lane] : 0ull;
        const u32 hc = (u32)__popcll(hm);
        u32 hi = hc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 x = __shfl_up_sync(0xFFFFFFFFu, hi, d);
            if (lane >= (u32)d) hi += x;
        }
        const u32 htot = __shfl_sync(0xFFFFFFFFu, hi, 31);
        if (lane == 0) s_halo = htot;
        u32 r = T + hi - hc;
        u64 m = hm;
        while (m) {
            u32 b = (u32)__ffsll((long long)m) - 1u;
            m &= m - 1;
            if (r < PP_NLCAP) nlpos[r] = (u16)((PP_THREADS + lane) * 64u + b);
            ++r;
        }
    }
    if (scouting) {
        u32 c2 = 0;
        if (scout_full) {
            // only the COUNT matters: the 0x80 flags of a word add up to 128 x (newlines in the word) in one dp4a
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                c2 = __dp4a(nl_flags(sv[k].x, c_nl, c_7f), 0x01010101u, c2);
                c2 = __dp4a(nl_flags(sv[k].y, c_nl, c_7f), 0x01010101u, c2);
                c2 = __dp4a(nl_flags(sv[k].z, c_nl, c_7f), 0x01010101u, c2);
                c2 = __dp4a(nl_flags(sv[k].w, c_nl, c_7f), 0x01010101u, c2);
            }
            c2 >>= 7;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u32 m = nl_mask16(sv[k], c_nl, c_7f);
                const u64 o = sbase + (u64)k * (PP_THREADS * 16u);
                if (o + 16u > (u64)p.n) m = o < (u64)p.n ? (m & ((1u << (u32)((u64)p.n - o)) - 1u)) : 0u;
                c2 += __popc(m);
            }
        }
        c2 = __reduce_add_sync(0xFFFFFFFFu, c2);
        if (lane == 0) s_scout[warp] = c2;
    }
    __syncthreads();                                    // newline positions complete; s_P, s_halo, s_scout visible
    TL_STAMP(tid == 0, 4);
    const u32 WN = T + s_halo;                          // newlines in the whole window
    const bool compact = WN <= PP_NLCAP;                // the compacted positions hold every newline of the window
    const u32 P = s_P;
    if (warp == 0) {
        // warp 0 works for the future while warps 1..7 split and pack this tile: it publishes the count of the scout
        // tile (PP_LEAD ahead) and resolves the prefix of the tile PP_PLEAD ahead.  Not the same tile: the counts just
        // before the scout tile are being published right now by this CTA's neighbours, and a look-back over them would
        // wait for the slowest of 128 CTAs (11 500 cycles, measured); the counts before the nearer tile are a CTA
        // lifetime old, so this look-back never spins and finds a resolved prefix within one or two hops.
        if (scouting) {
            u32 T_scout = 0;
#pragma unroll
            for (int w = 0; w < PP_THREADS / 32; ++w) T_scout += s_scout[w];
            if (lane == 0) st_volatile_u64(p.tile_state + stile, (1ull << 32) | (u64)T_scout);
        }
        const u32 ptile = tile + PP_PLEAD;
        if (ptile >= PP_LEAD && ptile < p.n_tiles) {          // (the first PP_LEAD tiles were resolved by k_scan_head)
            TL_STAMP(lane == 0, 9);
            const u32 Pu = pp_lookback(p.tile_state, ptile, lane);
            u64 own = ld_volatile_u64(p.tile_state + ptile);
            while ((own >> 32) == 0) own = ld_volatile_u64(p.tile_state + ptile);
            if (lane == 0) st_volatile_u64(p.tile_state + ptile, (2ull << 32) | (u64)(Pu + (u32)own));
            TL_STAMP(lane == 0, 10);
        }
        if (compact) return;
    }
    const u32 l4 = tid & (PP_GROUP - 1u);
    const u32 gmask = 0xFu << (lane & 28u);

    if (compact) {
        // ================= normal tile: PP_GROUP lanes per record, 64 records per round.  A record's line ends are
        // consecutive entries of nlpos; c = P mod LPR says which newline ends a record (the reference, too, simply counts
        // 4 (2) newlines per record).  Every lane of a group derives the geometry of its record itself (identical work,
        // broadcast loads) - no hand-over through shared memory, no barrier.
        const u32 g = (tid - 32u) / PP_GROUP;
        const u32 c = P % LPR;
        const u32 Rrel_first = tile == 0 ? 0u : 1u;     // owned records, counted from floor(P / LPR)
        const u32 Rrel_last = (c + T) / LPR;
        const u32 n_owned = Rrel_last >= Rrel_first ? Rrel_last - Rrel_first + 1u : 0u;
        for (u32 rbase = 0; rbase < n_owned; rbase += PP_PACKERS / PP_GROUP) {
            const u32 n_round = min((u32)(PP_PACKERS / PP_GROUP), n_owned - rbase);
            // ---- geometry + validation
            u32 status = RS_NONE, gstart = 0, off = PP_NONE, ql = 0;
#if FQD_K1_OWNERS
            // one owner thread per record (the first n_round packer threads) hands the geometry over through shared memory
            const u32 wtid = tid - 32u;
            if (rbase) bar_sync_c<1, PP_PACKERS>();                       // the queue of the round before has been read
            if (wtid < n_round) {
                const u32 Rrel = Rrel_first + rbase + wtid;
#else
            if (g < n_round) {
                const u32 Rrel = Rrel_first + rbase + g;
#endif
                const int j0 = (int)(LPR * Rrel) - 1 - (int)c;            // local rank of the newline before the record
                const u32 start_l = j0 < 0 ? p.skip : (u32)nlpos[j0] + 1u;
                gstart = base + start_l;
                if (gstart < p.n) {
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
                    status = classify_record<LPR, BYTES>(p, win, base, start_l, e[0], e[1], LPR == 4 ? e[2] : e[1],
                                                  LPR == 4 ? e[3] : e[1], off, ql);
                }
#if FQD_K1_OWNERS
                commit_record(p, slot_base, P / LPR + Rrel, gstart, status);
                q_off[wtid] = off; q_len[wtid] = ql;
            }
            bar_sync_c<1, PP_PACKERS>();
            off = PP_NONE; ql = 0;
            if (g < n_round) { off = q_off[g]; ql = q_len[g]; }
            TL_STAMP(tid == 32 && rbase == 0, 5);
            const u32 R = P / LPR + Rrel_first + rbase + g;
#else
            }
            TL_STAMP(tid == 32 && rbase == 0, 5);
            const u32 R = P / LPR + Rrel_first + rbase + g;
            if (g < n_round && l4 == 0) commit_record(p, slot_base, R, gstart, status);
#endif
            // ---- pack + commit
            if (off != PP_NONE && R < p.cap && slot_base + R < p.key_capacity) {
                u64* row = p.keys + (slot_base + R) * p.row_words + p.mate_off;
                u32 bad = 0;
                u64 hsum = 0, w0 = 0;
                if (p.W <= 8u) {
                    const u32 w = 2u * l4;
                    if (w < p.W) {
                        u32 bad_b;
                        const u64 wa = pack_word<BYTES>(win, p, base, off, ql, w, bad);
                        const u64 wb = pack_word<BYTES>(win, p, base, off, ql, w + 1u, bad_b);
                        if (!bad) bad = bad_b;
                        hsum = word_hash(wa, s_hkey[w]) + word_hash(wb, s_hkey[w + 1u]);
                        *reinterpret_cast<ulonglong2*>(row + w) = make_ulonglong2(wa, wb);
                        w0 = wa;
                    }
                } else {
                    for (u32 w = 2u * l4; w < p.W; w += 2u * PP_GROUP) {
                        u32 bad_a, bad_b;
                        const u64 xa = pack_word<BYTES>(win, p, base, off, ql, w, bad_a);
                        const u64 xb = pack_word<BYTES>(win, p, base, off, ql, w + 1u, bad_b);
                        *reinterpret_cast<ulonglong2*>(row + w) = make_ulonglong2(xa, xb);
                        if (w == 0) w0 = xa;
                        if (!bad) bad = bad_a ? bad_a : bad_b;
                        const uint2 ka = w < PP_HKEYS ? s_hkey[w] : pos_keys(p.hash_salt + w);
                        const uint2 kb = w + 1u < PP_HKEYS ? s_hkey[w + 1u] : pos_keys(p.hash_salt + w + 1u);
                        hsum += word_hash(xa, ka) + word_hash(xb, kb);
                    }
                }
                TL_STAMP(tid == 32 && rbase == 0, 6);
                commit_group(p, R, ql & 0x7FFFFFFFu, l4, gmask, hsum, w0, bad);
            }
            TL_STAMP(tid == 32, 8);
        }
        return;
    }

    // ================= very dense tile (tiny records; all 8 warps, after the look-back): every thread walks its
    // own newline bits
    {
        u32 R_first = (P + LPR) / LPR;
        const u32 R_last = (P + T) / LPR;
        if (tile == 0) R_first = 0;
        const u32 n_owned = R_last >= R_first ? R_last - R_first + 1u : 0u;
        const u32 g = tid / PP_GROUP;
        for (u32 rbase = 0; rbase < n_owned; rbase += PP_QCAP) {
            const u32 n_round = min(PP_QCAP, n_owned - rbase);
            u64 m = my_mask;
            u32 k = P + lex;
            bool virt = (tile == 0 && tid == 0);      // record 0 starts at offset 0 with no newline before it
            while (m || virt) {
                u32 R, start_l;
                if (virt) { virt = false; R = 0; start_l = p.skip; }
                else {
                    u32 b = (u32)__ffsll((long long)m) - 1u;
                    m &= m - 1;
                    u32 kk = k++;
                    if ((kk + 1u) % LPR) continue;
                    R = (kk + 1u) / LPR;
                    start_l = tid * 64u + b + 1u;
                }
                const u32 gstart = base + start_l;
                const u32 o = R - R_first;
                if (o < rbase || o >= rbase + PP_QCAP) continue;
                u32 qoff = PP_NONE, qlen = 0, status = RS_NONE;
                if (gstart < p.n) {
                    u32 e0 = find_nl(mask64, start_l, p.raw, base, p.n);
                    u32 e1 = e0 == PP_NONE ? PP_NONE : find_nl(mask64, e0 + 1u, p.raw, base, p.n);
                    u32 e3 = e1, e2 = e1;
                    if (LPR == 4) {
                        e2 = e1 == PP_NONE ? PP_NONE : find_nl(mask64, e1 + 1u, p.raw, base, p.n);
                        e3 = e2 == PP_NONE ? PP_NONE : find_nl(mask64, e2 + 1u, p.raw, base, p.n);
                    }
                    status = classify_record<LPR, BYTES>(p, win, base, start_l, e0, e1, e2, e3, qoff, qlen);
                }
                commit_record(p, slot_base, R, gstart, status);
                q_off[o - rbase] = qoff;
                q_len[o - rbase] = qlen;
            }
            __syncthreads();
            for (u32 q = g; q < n_round; q += PP_THREADS / PP_GROUP) {
                const u32 off = q_off[q];
                if (off == PP_NONE) continue;
                const u32 ql = q_len[q];
                const u32 R = R_first + rbase + q;
                if (R >= p.cap || slot_base + R >= p.key_capacity) continue;
                u64* row = p.keys + (slot_base + R) * p.row_words + p.mate_off;
                u64 hsum = 0, w0 = 0;
                u32 bad = 0;
                for (u32 w = 2u * l4; w < p.W; w += 2u * PP_GROUP) {
                    u32 bad_a, bad_b;
                    const u64 xa = pack_word<BYTES>(win, p, base, off, ql, w, bad_a);
                    const u64 xb = pack_word<BYTES>(win, p, base, off, ql, w + 1u, bad_b);
                    *reinterpret_cast<ulonglong2*>(row + w) = make_ulonglong2(xa, xb);
                    if (w == 0) w0 = xa;
                    if (!bad) bad = bad_a ? bad_a : bad_b;
                    const uint2 ka = w < PP_HKEYS ? s_hkey[w] : pos_keys(p.hash_salt + w);
                    const uint2 kb = w + 1u < PP_HKEYS ? s_hkey[w + 1u] : pos_keys(p.hash_salt + w + 1u);
                    hsum += word_hash(xa, ka) + word_hash(xb, kb);
                }
                commit_group(p, R, ql & 0x7FFFFFFFu, l4, gmask, hsum, w0, bad);
            }
            __syncthreads();
        }
    }
}

// The first PP_LEAD tiles of a chunk have no CTA ahead of them to count their newlines: a small kernel does it before K1
// starts (PP_LEAD x 16 KiB = 24 MB, a few microseconds), a second one turns the counts into resolved prefixes.  After
// that every tile of the chunk finds its prefix published by somebody who started EARLIER - which is what makes the
// spin-waits of K1 deadlock-free whatever the number of resident CTAs.
__global__ void __launch_bounds__(PP_THREADS) k_scout_head(const ParseParams p) {
    __shared__ u32 warp_sum[PP_THREADS / 32];
    const u32 tile = blockIdx.x, tid = threadIdx.x;
    u32 c_nl, c_7f;
    asm volatile("mov.u32 %0, 0x0A0A0A0A;" : "=r"(c_nl));
    asm volatile("mov.u32 %0, 0x7F7F7F7F;" : "=r"(c_7f));
    const u64 sbase = (u64)tile * PP_TILE + tid * 64u;
    const uint4* sp = reinterpret_cast<const uint4*>(p.raw + sbase);
    u32 cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const u64 o = sbase + 16u * k;
        if (o >= (u64)p.n) continue;
        u32 m = nl_mask16(__ldcg(sp + k), c_nl, c_7f);
        if (o + 16u > (u64)p.n) m &= (1u << (u32)((u64)p.n - o)) - 1u;
        if (o < (u64)p.skip) m &= o + 16u <= (u64)p.skip ? 0u : ~0u << (u32)((u64)p.skip - o);      // newlines before the chunk proper are not ours
        cnt += __popc(m);
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if ((tid & 31u) == 0) warp_sum[tid >> 5] = cnt;
    __syncthreads();
    if (tid == 0) {
        u32 tot = 0;
#pragma unroll
        for (int w = 0; w < PP_THREADS / 32; ++w) tot += warp_sum[w];
        p.tile_state[tile] = (1ull << 32) | (u64)tot;
    }
}
__global__ void __launch_bounds__(1024) k_scan_head(u64* tile_state, u32 n_head) {
    __shared__ u32 s_part[1024];
    // thread t owns entries [t * per, (t + 1) * per)
    const u32 per = (n_head + 1023u) / 1024u;
    const u32 lo = min(n_head, threadIdx.x * per), hi = min(n_head, lo + per);
    u32 sum = 0;
    for (u32 i = lo; i < hi; ++i) sum += (u32)tile_state[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (u32 d = 1; d < 1024; d <<= 1) {          // inclusive scan of the partial sums
        const u32 v = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    u32 run = s_part[threadIdx.x] - sum;
    for (u32 i = lo; i < hi; ++i) { run += (u32)tile_state[i]; tile_state[i] = (2ull << 32) | (u64)run; }
}

// one CTA per tile
static inline void pp_launch(bool fastq, const ParseParams& p, cudaStream_t stream) {
    if (p.n_tiles == 0) return;
    const u32 n_head = p.n_tiles < PP_LEAD ? p.n_tiles : PP_LEAD;
    k_scout_head<<<n_head, PP_THREADS, 0, stream>>>(p);
    k_scan_head<<<1, 1024, 0, stream>>>(p.tile_state, n_head);
    if (p.byte_keys) {
        if (fastq) k_parse_pack<4, true><<<p.n_tiles, PP_THREADS, 0, stream>>>(p);
        else k_parse_pack<2, true><<<p.n_tiles, PP_THREADS, 0, stream>>>(p);
    } else {
        if (fastq) k_parse_pack<4, false><<<p.n_tiles, PP_THREADS, 0, stream>>>(p);
        else k_parse_pack<2, false><<<p.n_tiles, PP_THREADS, 0, stream>>>(p);
    }
}

__global__ void k_init_chunk(ChunkCtl* ctl, u64* tile_state, u32 n_tiles) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        ctl->ticket = 0; ctl->n_newlines = 0; ctl->n_records = 0; ctl->consumed = 0;
        ctl->err_parse = NO_ERR; ctl->err_base = NO_ERR; ctl->too_long = 0; ctl->pad = 0; ctl->too_long_rec = ~0u; ctl->pad2 = 0;
    }
    for (; i < n_tiles; i += gridDim.x * blockDim.x) tile_state[i] = 0;
}


}  // namespace fqd
