// pinflate.hpp - ONE gzip member inflated by many threads (the reference inflates on the main thread,
// src/file_utils.cpp:59-66).  `gzip file.fastq` and pigz write a single member, so member-level parallelism
// (pargz.hpp) finds nothing to split there; this does, inside the deflate stream:
//   * the compressed range is cut at nominal byte boundaries into chunks; the task of a chunk searches, from its
//     boundary on, the first bit position that parses as the header of a dynamic-Huffman block (complete code-length
//     / literal / distance codes - random bits almost never pass), and decodes from there up to the first block
//     boundary at or beyond the NEXT nominal boundary;
//   * the 32 KiB of history before a chunk are unknown while it is decoded, so the decoder emits 16-bit symbols:
//     a byte, or a MARKER "byte j of the window that precedes this chunk" (markers are copied by later matches like
//     any other symbol);
//   * the consumer walks the chunks in order.  A chunk is accepted only if it starts exactly at the bit where the
//     accepted stream ended - the member's first block is a true start, so by induction every accepted chunk began at a
//     true block boundary and its symbols, with the markers replaced from the now known window, are exactly the bytes
//     zlib would have produced.  A chunk whose start was a false positive (or skipped a block the search does not
//     recognise: stored, fixed, final) is dropped or bridged by a filler decode from the accepted position;
//   * marker replacement + narrowing to bytes + CRC-32 of accepted chunks run on the pool again; the CRCs are combined in
//     order and checked, with the length, against the member's trailer.
// Own deflate decoder (zlib cannot emit markers): canonical Huffman with a 12-bit direct table and a bit-serial slow
// path for longer codes.  Unit-tested on the CPU against Python's gzip (tests/test_host_io.py).
#pragma once
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <vector>

#include "workers.hpp"

namespace fqdhost {

namespace pinfl {

constexpr unsigned kFastBits = 12;
constexpr size_t kWindow = 32768;
constexpr uint16_t kMarker = 0x8000;
constexpr uint64_t kNone = ~0ull;

// malloc'd array without value-initialisation (std::vector::resize would write every byte once more)
template <class T> struct RawBuf {
    T* p = nullptr; size_t n = 0, cap = 0;
    RawBuf() = default;
    RawBuf(const RawBuf&) = delete;
    RawBuf& operator=(const RawBuf&) = delete;
    ~RawBuf() { std::free(p); }
    bool reserve(size_t want) {
        if (want <= cap) return true;
        T* q = (T*)std::realloc(p, want * sizeof(T));
        if (!q) return false;
        p = q; cap = want;
        return true;
    }
    void swap(RawBuf& o) { std::swap(p, o.p); std::swap(n, o.n); std::swap(cap, o.cap); }
    void release() { std::free(p); p = nullptr; n = cap = 0; }
    size_t size() const { return n; }
    T* data() { return p; }
    const T* data() const { return p; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
};

struct BitReader {
    const uint8_t* base; const uint8_t* p; const uint8_t* end;
    uint64_t buf = 0; unsigned cnt = 0;
    size_t past = 0;             // virtual zero bytes consumed beyond `end`
    BitReader(const uint8_t* b, size_t size, uint64_t bitpos) : base(b), p(b + (bitpos >> 3)), end(b + size) {
        if (p > end) p = end;
        refill();
        unsigned skip = (unsigned)(bitpos & 7);
        buf >>= skip; cnt -= skip;
    }
    inline void refill() {
        if (end - p >= 8) {
            uint64_t v; std::memcpy(&v, p, 8);
            buf |= v << cnt;
            p += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56) {
                if (p < end) buf |= (uint64_t)*p++ << cnt; else ++past;
                cnt += 8;
            }
        }
    }
    inline uint32_t peek(unsigned n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    inline void drop(unsigned n) { buf >>= n; cnt -= n; }
    inline uint32_t bits(unsigned n) { uint32_t v = peek(n); drop(n); return v; }
    uint64_t bitpos() const { return (uint64_t)(p - base + past) * 8 - cnt; }
    bool overrun() const { return bitpos() > (uint64_t)(end - base) * 8; }
    void align() { drop(cnt & 7); }
};

struct Huff {
    uint16_t fast[1u << kFastBits];     // (symbol << 4) | length, 0 = longer than kFastBits (or unused)
    uint16_t count[16];
    uint16_t symbol[288];
    // returns the unused code space: 0 complete, > 0 incomplete, < 0 over-subscribed
    int build(const uint8_t* lens, int n) {
        std::memset(count, 0, sizeof count);
        for (int i = 0; i < n; ++i) ++count[lens[i]];
        int left = 1;
        for (int l = 1; l <= 15; ++l) { left <<= 1; left -= count[l]; if (left < 0) return left; }
        uint16_t offs[16]; offs[1] = 0;
        for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
        for (int i = 0; i < n; ++i) if (lens[i]) symbol[offs[lens[i]]++] = (uint16_t)i;
        std::memset(fast, 0, sizeof fast);
        uint32_t code = 0; int idx = 0;
        for (unsigned l = 1; l <= kFastBits; ++l) {
            for (int k = 0; k < count[l]; ++k, ++idx, ++code) {
                uint32_t r = 0;
                for (unsigned b = 0; b < l; ++b) r |= ((code >> b) & 1u) << (l - 1 - b);
                const uint16_t e = (uint16_t)((symbol[idx] << 4) | l);
                for (uint32_t i = r; i < (1u << kFastBits); i += 1u << l) fast[i] = e;
            }
            code <<= 1;
        }
        return left;
    }
    // -1 on an invalid code
    inline int decode(BitReader& br) const {
        const uint16_t e = fast[br.peek(kFastBits)];
        if (e) { br.drop(e & 15u); return e >> 4; }
        int code = 0, first = 0, index = 0;
        uint64_t b = br.buf;
        for (int l = 1; l <= 15; ++l) {
            code |= (int)(b & 1); b >>= 1;
            const int c = count[l];
            if (code - c < first) { br.drop((unsigned)l); return symbol[index + (code - first)]; }
            index += c; first += c; first <<= 1; code <<= 1;
        }
        return -1;
    }
};

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// Pre-decoded table for the hot loop: one lookup gives the literal, or the base value and the number of extra bits
// of a length / distance symbol.  Entry: bits 0-7 code bits to drop, 8-11 extra bits (or the width of a subtable),
// 12-13 kind, 16-31 value (byte, base, subtable offset).  Codes longer than PB bits go through a subtable.
enum : uint32_t { kKindBase = 0u << 12, kKindLiteral = 1u << 12, kKindSub = 2u << 12, kKindOther = 3u << 12, kKindMask = 3u << 12 };
constexpr uint32_t kEntryEob = kKindOther | (0u << 16), kEntryInvalid = kKindOther | (1u << 16);

template <unsigned PB> struct FastTable {
    uint32_t t[(1u << PB) + 288 * 16];
    void build(const Huff& h, bool litlen) {
        for (uint32_t i = 0; i < (1u << PB); ++i) t[i] = kEntryInvalid;
        auto entry = [litlen](unsigned sym, unsigned nbits) -> uint32_t {
            if (litlen) {
                if (sym < 256) return kKindLiteral | (sym << 16) | nbits;
                if (sym == 256) return kEntryEob | nbits;
                if (sym > 285) return kEntryInvalid | nbits;
                return kKindBase | ((uint32_t)kLenBase[sym - 257] << 16) | ((uint32_t)kLenExtra[sym - 257] << 8) | nbits;
            }
            if (sym > 29) return kEntryInvalid | nbits;
            return kKindBase | ((uint32_t)kDistBase[sym] << 16) | ((uint32_t)kDistExtra[sym] << 8) | nbits;
        };
        auto reversed = [](uint32_t code, unsigned l) { uint32_t r = 0; for (unsigned b = 0; b < l; ++b) r |= ((code >> b) & 1u) << (l - 1 - b); return r; };
        // short codes straight into the primary table; the longest code behind every PB-bit prefix sizes its subtable
        uint8_t maxlen[1u << PB];
        bool any_long = false;
        uint32_t code = 0; int idx = 0;
        for (unsigned l = 1; l <= 15; ++l) {
            for (int k = 0; k < h.count[l]; ++k, ++idx, ++code) {
                const uint32_t r = reversed(code, l);
                if (l <= PB) {
                    const uint32_t e = entry(h.symbol[idx], l);
                    for (uint32_t i = r; i < (1u << PB); i += 1u << l) t[i] = e;
                } else {
                    if (!any_long) { std::memset(maxlen, 0, sizeof maxlen); any_long = true; }
                    uint8_t& m = maxlen[r & ((1u << PB) - 1)];
                    if (l > m) m = (uint8_t)l;
                }
            }
            code <<= 1;
        }
        if (litlen) {
            // two literals in one entry where both codes fit into the PB index bits (short codes dominate in text):
            // bit 8 = a second literal follows, value = first | second << 8
            uint32_t single[1u << PB];
            std::memcpy(single, t, sizeof single);
            for (uint32_t i = 0; i < (1u << PB); ++i) {
                const uint32_t e = single[i];
                if ((e & kKindMask) != kKindLiteral) continue;
                const unsigned l1 = e & 0xFFu;
                if (l1 >= PB) continue;
                const uint32_t e2 = single[i >> l1];
                if ((e2 & kKindMask) != kKindLiteral || (e2 & 0xFFu) > PB - l1) continue;
                t[i] = kKindLiteral | (1u << 8) | ((e >> 16) << 16) | ((e2 >> 16) << 24) | (l1 + (e2 & 0xFFu));
            }
        }
        if (!any_long) return;
        uint32_t next = 1u << PB;
        for (uint32_t pfx = 0; pfx < (1u << PB); ++pfx) {
            if (!maxlen[pfx]) continue;
            const unsigned sb = maxlen[pfx] - PB;
            t[pfx] = kKindSub | (next << 16) | (sb << 8) | PB;
            for (uint32_t i = 0; i < (1u << sb); ++i) t[next + i] = kEntryInvalid;
            next += 1u << sb;
        }
        code = 0; idx = 0;
        for (unsigned l = 1; l <= 15; ++l) {
            for (int k = 0; k < h.count[l]; ++k, ++idx, ++code) {
                if (l <= PB) continue;
                const uint32_t r = reversed(code, l);
                const uint32_t sub = t[r & ((1u << PB) - 1)];
                const uint32_t off = sub >> 16, sb = (sub >> 8) & 15u;
                const uint32_t e = entry(h.symbol[idx], l - PB);
                for (uint32_t i = r >> PB; i < (1u << sb); i += 1u << (l - PB)) t[off + i] = e;
            }
            code <<= 1;
        }
    }
    inline uint32_t lookup(BitReader& br) const {       // consumes the code's bits
        uint32_t e = t[br.buf & ((1u << PB) - 1)];
        if ((e & kKindMask) == kKindSub) {
            br.drop(PB);
            e = t[(e >> 16) + (uint32_t)(br.buf & ((1u << ((e >> 8) & 15u)) - 1))];
        }
        br.drop(e & 0xFFu);
        return e;
    }
};

// Header of a dynamic block (after the 3 block-type bits).  strict = the search's acceptance test.
inline bool read_dynamic_header(BitReader& br, Huff& lit, Huff& dist, bool strict) {
    br.refill();
    const unsigned hlit = br.bits(5) + 257, hdist = br.bits(5) + 1, hclen = br.bits(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    uint8_t cl[19] = {0};
    for (unsigned i = 0; i < hclen; ++i) { if (br.cnt < 3) br.refill(); cl[kClOrder[i]] = (uint8_t)br.bits(3); }
    unsigned kraft = 0;                                       // zlib, too, wants this code complete
    for (int i = 0; i < 19; ++i) if (cl[i]) kraft += 128u >> cl[i];
    if (kraft != 128u) return false;
    Huff clh;
    if (clh.build(cl, 19) != 0) return false;
    uint8_t lens[286 + 30];
    unsigned n = 0;
    while (n < hlit + hdist) {
        br.refill();
        int s = clh.decode(br);
        if (s < 0) return false;
        if (s < 16) { lens[n++] = (uint8_t)s; continue; }
        unsigned rep; uint8_t v = 0;
        if (s == 16) { if (n == 0) return false; v = lens[n - 1]; rep = 3 + br.bits(2); }
        else if (s == 17) rep = 3 + br.bits(3);
        else rep = 11 + br.bits(7);
        if (n + rep > hlit + hdist) return false;
        while (rep--) lens[n++] = v;
    }
    if (lens[256] == 0) return false;                         // no end-of-block code
    const int l_left = lit.build(lens, (int)hlit);
    if (l_left < 0 || (l_left > 0 && (strict || hlit - lit.count[0] != 1))) return false;
    const int d_left = dist.build(lens + hlit, (int)hdist);
    if (d_left < 0) return false;
    if (d_left > 0) {                                         // incomplete: only "at most one distance code" is legal
        const unsigned used = hdist - dist.count[0];
        if (used > 1) return false;
    }
    return !br.overrun();
}

inline void fixed_tables(Huff& lit, Huff& dist) {
    uint8_t lens[288];
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    lit.build(lens, 288);
    uint8_t d[30];
    for (int i = 0; i < 30; ++i) d[i] = 5;
    dist.build(d, 30);
}

// First bit position in [from, to) that passes as the start of a non-final dynamic block, or kNone.
inline uint64_t find_block(const uint8_t* data, size_t size, uint64_t from, uint64_t to) {
    auto tables = std::make_unique<std::pair<Huff, Huff>>();
    const uint64_t total = (uint64_t)size * 8;
    for (uint64_t b = from; b < to && b + 64 < total; ++b) {
        // cheap gate on 13 bits: BFINAL = 0, BTYPE = 2, HLIT <= 29, HDIST <= 29
        const size_t byte = (size_t)(b >> 3);
        uint32_t v = 0;
        std::memcpy(&v, data + byte, byte + 4 <= size ? 4 : size - byte);
        v >>= (b & 7);
        if ((v & 7u) != 4u) continue;
        if (((v >> 3) & 31u) > 29u || ((v >> 8) & 31u) > 29u) continue;
        BitReader br(data, size, b + 3);
        if (!read_dynamic_header(br, tables->first, tables->second, true)) continue;
        // a few symbols must decode to something legal
        bool ok = true;
        for (int i = 0; i < 64 && ok; ++i) {
            br.refill();
            int s = tables->first.decode(br);
            if (s < 0 || s > 285) { ok = false; break; }
            if (s == 256) break;
            if (s > 256) {
                br.drop(kLenExtra[s - 257]);
                int d = tables->second.decode(br);
                if (d < 0 || d > 29) { ok = false; break; }
                br.refill();
                br.drop(kDistExtra[d]);
            }
        }
        if (ok && !br.overrun()) return b;
    }
    return kNone;
}

struct Chunk {
    // plan
    uint64_t from_bit = 0;      // exact start (exact = true) or where the search begins
    uint64_t stop_bit = 0;      // decode up to the first block boundary at or beyond this bit
    uint64_t search_end = 0;    // the search gives up here
    bool exact = false;         // start is known to be a block start (first block of the member, fillers)
    bool known_window = false;  // nothing precedes (first block of the member): a reference before the start is an error
    // result
    uint64_t start_bit = kNone; // where decoding began
    uint64_t end_bit = 0;       // block boundary where it stopped
    bool final_block = false, failed = false, truncated = false;
    RawBuf<uint16_t> sym;       // symbols from the start of the chunk ...
    RawBuf<uint8_t> tail;       // ... and, once 32 KiB in a row were free of markers, plain bytes for the rest:
    size_t tail_skip = 0;       //     tail[0, tail_skip) is a copy of those 32 KiB (the decoder's window), not output
    bool ready = false;
};

struct DecodeResult {
    uint64_t end_bit = 0;       // block boundary where decoding stopped
    bool final_block = false, failed = false, truncated = false;
    bool too_big = false;       // failed because max_symbols (or memory) was exhausted, not because of the data
};

// Decode deflate blocks from start_bit up to the first block boundary at or beyond stop_bit (or the final block),
// APPENDING to out (out[base] is where this deflate stream, or chunk, began).  T = uint16_t: the window before the start may be unknown (known_window = false) and references
// into it become markers; T = uint8_t: plain bytes, nothing may reach before out[base] (the start of this stream).
struct DecodeScratch {          // the four tables of the block being decoded (~60 KB; one per task, not per call)
    Huff lit, dist;
    FastTable<11> flit;
    FastTable<8> fdist;
};

template <class T>
inline void decode_blocks(DecodeScratch& sc, const uint8_t* data, size_t size, uint64_t start_bit, uint64_t stop_bit,
                          bool known_window, size_t max_symbols, RawBuf<T>& out, DecodeResult& r, size_t base = ~(size_t)0) {
    static_assert(sizeof(T) <= 2, "byte or 16-bit symbols");
    constexpr bool kMarkers = sizeof(T) == 2;
    constexpr unsigned kStep = 8 / sizeof(T);          // symbols per 8-byte copy step
    Huff* lit = &sc.lit; Huff* dist = &sc.dist;
    FastTable<11>* flit = &sc.flit; FastTable<8>* fdist = &sc.fdist;
    BitReader br(data, size, start_bit);
    if (base == ~(size_t)0) base = out.n;               // out[base] is the first symbol of this deflate stream / chunk
    const size_t n0 = out.n;
    size_t n = out.n;
    const size_t guess = (size_t)((stop_bit > start_bit && stop_bit != kNone ? stop_bit - start_bit : 0) / 8) * 5 + (1u << 16);
    if (!out.reserve(n + std::min(guess, std::max<size_t>(max_symbols, 1u << 16)))) { r.failed = true; return; }
    const uint64_t total_bits = (uint64_t)size * 8;
    r.end_bit = start_bit;
    for (;;) {
        br.refill();
        const unsigned bfinal = br.bits(1), btype = br.bits(2);
        if (btype == 3) { r.failed = true; break; }
        if (btype == 0) {
            br.align(); br.refill();
            const uint32_t len = br.bits(16), nlen = br.bits(16);
            if (br.overrun()) { r.failed = r.truncated = true; break; }
            if ((len ^ nlen) != 0xFFFFu) { r.failed = true; break; }
            if (br.bitpos() + (uint64_t)len * 8 > total_bits) { r.failed = r.truncated = true; }
            uint64_t pos = br.bitpos() >> 3;
            size_t take = r.truncated ? (size_t)(size - std::min<uint64_t>(size, pos)) : len;
            if (n + take + 320 > out.cap) {
                if (n + take - base > max_symbols || !out.reserve(std::max(out.cap * 2, n + take + 320))) { r.failed = r.too_big = true; r.truncated = false; break; }
            }
            if (sizeof(T) == 1) std::memcpy(&out[n], data + pos, take);
            else for (size_t i = 0; i < take; ++i) out[n + i] = (T)data[pos + i];
            n += take;
            if (r.truncated) break;
            br = BitReader(data, size, (pos + len) * 8);
        } else {
            if (btype == 1) fixed_tables(*lit, *dist);
            else if (!read_dynamic_header(br, *lit, *dist, false)) { r.failed = true; r.truncated = br.overrun(); break; }
            bool bad = false, eob = false;
            flit->build(*lit, true); fdist->build(*dist, false);
            // hot loop: at least 16 input bytes ahead, so the refill is one 8-byte load and no bit beyond the end of
            // the input is ever consumed; one refill (>= 56 bits) covers a whole match (15 + 5 + 15 + 13) or 4 lookups
            while (br.end - br.p >= 16) {
                if (n + 320 > out.cap) {
                    if (n - base >= max_symbols || !out.reserve(out.cap * 2)) { bad = r.too_big = true; break; }
                }
                br.refill();
                uint32_t e = flit->lookup(br);
                if ((e & kKindMask) == kKindLiteral) {
                    // an entry holds one or two literals (the second slot is written either way: slack)
                    out[n] = (T)((e >> 16) & 0xFFu); out[n + 1] = (T)(e >> 24); n += 1 + ((e >> 8) & 1u);
                    for (int more = 0; more < 3; ++more) {          // 15 + 3 x 11 bits <= one refill
                        e = flit->t[br.buf & 0x7FFu];
                        if ((e & kKindMask) != kKindLiteral) break;
                        br.drop(e & 0xFFu);
                        out[n] = (T)((e >> 16) & 0xFFu); out[n + 1] = (T)(e >> 24); n += 1 + ((e >> 8) & 1u);
                    }
                    continue;
                }
                if ((e & kKindMask) != kKindBase) { if ((e & ~0xFFu) == kEntryEob) eob = true; else bad = true; break; }
                const unsigned len = (e >> 16) + br.bits((e >> 8) & 15u);
                const uint32_t d = fdist->lookup(br);
                if ((d & kKindMask) != kKindBase) { bad = true; break; }
                const size_t dd = (d >> 16) + br.bits((d >> 8) & 15u);
                if (dd > n - base) {
                    if (!kMarkers || known_window || dd > kWindow) { bad = true; break; }
                    // (part of) the source lies in the unknown window before this chunk
                    for (unsigned i = 0; i < len; ++i, ++n)
                        out[n] = dd > n - base ? (T)(kMarker | (kWindow - (dd - (n - base)))) : out[n - dd];
                } else if (dd >= kStep) {
                    // 8 bytes per step; may write up to kStep - 1 symbols of slack beyond the match
                    const T* src = &out[n - dd]; T* dst = &out[n];
                    for (unsigned i = 0; i < len; i += kStep) std::memcpy(dst + i, src + i, 8);
                    n += len;
                } else if (dd == 1) {
                    const T v = out[n - 1];
                    for (unsigned i = 0; i < len; ++i) out[n + i] = v;
                    n += len;
                } else {
                    for (unsigned i = 0; i < len; ++i, ++n) out[n] = out[n - dd];
                }
            }
            // the last bytes of the input (and nothing else): symbol by symbol, with the end in view
            for (; !bad && !eob;) {
                if (n + 320 > out.cap) {
                    if (n - base >= max_symbols || !out.reserve(out.cap * 2)) { bad = r.too_big = true; break; }
                }
                br.refill();
                int s = lit->decode(br);
                if (br.past && br.overrun()) { bad = true; break; }      // the symbol needs bits beyond the end of the input
                if (s < 256) {
                    if (s < 0) { bad = true; break; }
                    out[n++] = (T)s;
                    continue;
                }
                if (s == 256) break;
                if (s > 285) { bad = true; break; }
                const unsigned len = kLenBase[s - 257] + br.bits(kLenExtra[s - 257]);
                int d = dist->decode(br);
                if (d < 0 || d > 29) { bad = true; break; }
                br.refill();
                const size_t dd = kDistBase[d] + br.bits(kDistExtra[d]);
                if (br.past && br.overrun()) { bad = true; break; }
                if (dd > n - base) {
                    if (!kMarkers || known_window || dd > kWindow) { bad = true; break; }
                    for (unsigned i = 0; i < len; ++i, ++n)
                        out[n] = dd > n - base ? (T)(kMarker | (kWindow - (dd - (n - base)))) : out[n - dd];
                } else {
                    for (unsigned i = 0; i < len; ++i, ++n) out[n] = out[n - dd];
                }
            }
            if (bad || br.overrun()) { r.failed = true; r.truncated = br.overrun(); break; }
        }
        r.end_bit = br.bitpos();
        if (bfinal) { r.final_block = true; break; }
        if (r.end_bit >= stop_bit) break;
    }
    if (r.failed && !r.truncated) n = n0;
    out.n = n;
}

inline bool marker_free(const uint16_t* s, size_t n) {
    size_t i = 0;
#if defined(__SSE2__)
    __m128i acc = _mm_setzero_si128();
    for (; i + 8 <= n; i += 8) acc = _mm_or_si128(acc, _mm_loadu_si128((const __m128i*)(s + i)));
    if (_mm_movemask_epi8(acc) & 0xAAAA) return false;
#endif
    for (; i < n; ++i) if (s[i] & kMarker) return false;
    return true;
}

// One chunk of a member: find its start (unless exact); 16-bit symbols while references into the unknown window are
// still alive, block by block, and plain bytes (the faster decoder, no marker pass afterwards) from the first block
// boundary at which the last 32 KiB hold no marker any more - in text that happens early.
inline void decode_chunk(const uint8_t* data, size_t size, Chunk& c, size_t max_symbols) {
    if (c.exact) c.start_bit = c.from_bit;
    else c.start_bit = find_block(data, size, c.from_bit, c.search_end);
    if (c.start_bit == kNone) { c.failed = true; return; }
    DecodeResult r;
    auto sc = std::make_unique<DecodeScratch>();
    c.sym.n = 0; c.tail.n = 0; c.tail_skip = 0;
    uint64_t pos = c.start_bit;
    bool bytes_now = c.known_window;                   // nothing precedes the member's first block: bytes at once
    while (!bytes_now) {
        r = DecodeResult();
        decode_blocks<uint16_t>(*sc, data, size, pos, pos + 1, false, max_symbols, c.sym, r, 0);     // one block
        if (r.failed || r.final_block || r.end_bit >= c.stop_bit) break;
        pos = r.end_bit;
        if (c.sym.n >= kWindow && marker_free(c.sym.data() + c.sym.n - kWindow, kWindow)) {
            if (!c.tail.reserve(kWindow + (size_t)((c.stop_bit > pos ? c.stop_bit - pos : 0) / 8) * 5 + (1u << 16))) break;
            for (size_t i = 0; i < kWindow; ++i) c.tail[i] = (uint8_t)c.sym[c.sym.n - kWindow + i];
            c.tail.n = c.tail_skip = kWindow;
            bytes_now = true;
        }
    }
    if (bytes_now) {
        r = DecodeResult();
        decode_blocks<uint8_t>(*sc, data, size, pos, c.stop_bit, true, max_symbols, c.tail, r, 0);
    }
    c.end_bit = r.end_bit; c.final_block = r.final_block; c.failed = r.failed; c.truncated = r.truncated;
    if (c.failed && !c.truncated) { c.sym.n = 0; c.tail.n = c.tail_skip = 0; }
}

// Length of the gzip member header at p (RFC 1952); 0 if it is malformed, kCutOff if the file ends inside it.
constexpr size_t kCutOff = ~(size_t)0;
inline size_t gzip_header_len(const uint8_t* data, size_t size, size_t p) {
    const size_t p0 = p;
    if (p + 4 <= size && (data[p] != 0x1f || data[p + 1] != 0x8b || data[p + 2] != 8 || (data[p + 3] & 0xe0))) return 0;
    if (p + 10 > size) return kCutOff;
    const unsigned flg = data[p + 3];
    p += 10;
    if (flg & 4) { if (p + 2 > size) return kCutOff; size_t xlen = data[p] | (size_t)data[p + 1] << 8; p += 2 + xlen; if (p > size) return kCutOff; }
    for (unsigned bit : {8u, 16u})
        if (flg & bit) { while (p < size && data[p]) ++p; if (p >= size) return kCutOff; ++p; }
    if (flg & 2) { p += 2; if (p > size) return kCutOff; }
    return p - p0;
}

struct Piece {           // an accepted chunk on its way to bytes
    RawBuf<uint16_t> sym;
    std::vector<uint8_t> window;     // the 32 KiB before it (shorter at the start of the member)
    RawBuf<uint8_t> bytes;           // sym with the markers replaced ...
    RawBuf<uint8_t> tail;            // ... followed by tail[tail_skip, n): the part that was decoded as bytes
    size_t tail_skip = 0;
    size_t size() const { return bytes.n + (tail.n - tail_skip); }
    uint32_t crc = 0;
    bool ready = false, bad = false; // bad: a marker points before the start of the member's output
};

inline void resolve_piece(Piece& pc) {
    const size_t n = pc.sym.size();
    if (!pc.bytes.reserve(std::max<size_t>(n, 1))) { pc.bad = true; return; }
    pc.bytes.n = n;
    const uint8_t* w = pc.window.data();
    const size_t missing = kWindow - pc.window.size();   // markers index a full window; its first `missing` bytes do not exist
    const uint16_t* s = pc.sym.data();
    uint8_t* o = pc.bytes.data();
    size_t i = 0;
    while (i < n) {
        // stretches without markers (everything, some 100 KB into a chunk of text) are narrowed 16 symbols at a time
#if defined(__SSE2__)
        while (i + 16 <= n) {
            const __m128i a = _mm_loadu_si128((const __m128i*)(s + i)), b = _mm_loadu_si128((const __m128i*)(s + i + 8));
            if (_mm_movemask_epi8(_mm_or_si128(a, b)) & 0xAAAA) break;           // bit 15 of some symbol: a marker
            _mm_storeu_si128((__m128i*)(o + i), _mm_packus_epi16(a, b));
            i += 16;
        }
#else
        while (i + 16 <= n) {
            uint16_t any = 0;
            for (int k = 0; k < 16; ++k) any |= s[i + k];
            if (any & kMarker) break;
            for (int k = 0; k < 16; ++k) o[i + k] = (uint8_t)s[i + k];
            i += 16;
        }
#endif
        const size_t stop = std::min(n, i + 16);
        for (; i < stop; ++i) {
            const uint16_t v = s[i];
            if (v < kMarker) { o[i] = (uint8_t)v; continue; }
            const size_t j = v & 0x7FFFu;
            if (j < missing) { pc.bad = true; o[i] = 0; } else o[i] = w[j - missing];
        }
    }
    pc.crc = (uint32_t)crc32_z(0L, (const Bytef*)o, n);
    if (pc.tail.n > pc.tail_skip)        // (crc32_z with a null buffer would RESET the value)
        pc.crc = (uint32_t)crc32_z(pc.crc, (const Bytef*)pc.tail.data() + pc.tail_skip, pc.tail.n - pc.tail_skip);
}

}  // namespace pinfl

// size knobs for tests (small files must still be cut into many pieces)
inline size_t env_size(const char* name, size_t dflt) {
    const char* e = std::getenv(name);
    long long v = e ? std::atoll(e) : 0;
    return v > 0 ? (size_t)v : dflt;
}

class ParallelMemberInflater {
public:
    // data/size: the whole mapped file; member_start: offset of a gzip member header
    ParallelMemberInflater(const unsigned char* data, size_t size, size_t member_start, size_t chunk_bytes = 1u << 20, int window = 0);
    ~ParallelMemberInflater();
    ParallelMemberInflater(const ParallelMemberInflater&) = delete;
    ParallelMemberInflater& operator=(const ParallelMemberInflater&) = delete;
    // up to n bytes of the member; 0 once it is complete (then done() is true)
    size_t read(char* dst, size_t n);
    bool done() const { return m_done; }
    bool truncated() const { return m_truncated; }
    size_t end_offset() const { return m_end_offset; }    // first byte after the member's trailer
    size_t chunks_accepted() const { return m_accepted; }
    size_t chunks_dropped() const { return m_dropped; }
    size_t fillers() const { return m_fillers; }
    size_t accepted_compressed() const { return (size_t)(m_end_bit >> 3) - m_deflate_start; }   // ... and what it inflated to:
    size_t accepted_output() const { return m_sym_bytes + m_direct_bytes; }
    size_t symbol_bytes() const { return m_sym_bytes; }      // decoded as 16-bit symbols (markers alive)
    size_t direct_bytes() const { return m_direct_bytes; }   // decoded as bytes

private:
    struct Sync {
        std::mutex mu; std::condition_variable cv; int inflight = 0;
        // symbol / byte arrays go round (a fresh multi-megabyte malloc is an mmap + a page fault per 4 KiB)
        std::vector<std::unique_ptr<pinfl::RawBuf<uint16_t>>> spare_sym;
        std::vector<std::unique_ptr<pinfl::RawBuf<uint8_t>>> spare_bytes;
        template <class T> static void take(std::vector<std::unique_ptr<pinfl::RawBuf<T>>>& from, pinfl::RawBuf<T>& into) {
            if (!from.empty()) { into.swap(*from.back()); into.n = 0; from.pop_back(); }
        }
        template <class T> static void give(std::vector<std::unique_ptr<pinfl::RawBuf<T>>>& to, pinfl::RawBuf<T>& b, size_t keep) {
            if (b.p && to.size() < keep) { to.emplace_back(new pinfl::RawBuf<T>()); to.back()->swap(b); }
            b.release();
        }
    };
    using ChunkPtr = std::shared_ptr<pinfl::Chunk>;
    using PiecePtr = std::shared_ptr<pinfl::Piece>;
    void launch(ChunkPtr c, bool express);
    void top_up();
    enum Step { ADVANCED, WINDOW_FULL, CHAIN_DONE };
    Step advance_chain();        // look at the oldest planned chunk: accept / drop / bridge
    void accept(ChunkPtr c);

    const unsigned char* m_data; size_t m_size;
    size_t m_chunk; int m_window;
    RunAhead m_ahead;
    std::shared_ptr<Sync> m_sync;
    size_t m_deflate_start = 0;
    size_t m_next_boundary = 0;          // nominal byte boundary of the next chunk to plan
    std::deque<ChunkPtr> m_chunks;       // planned, in stream order
    std::deque<PiecePtr> m_pieces;       // accepted, being resolved / copied out
    size_t m_piece_off = 0;
    uint64_t m_end_bit = 0;              // accepted stream ends here
    std::vector<uint8_t> m_win;          // last <= 32 KiB of accepted output
    uint32_t m_crc = 0; uint64_t m_len = 0;
    bool m_chain_done = false, m_done = false, m_truncated = false, m_first = true;
    size_t m_end_offset = 0;
    size_t m_accepted = 0, m_dropped = 0, m_fillers = 0, m_sym_bytes = 0, m_direct_bytes = 0;
};

// ---- implementation -------------------------------------------------------------------------------------------
inline ParallelMemberInflater::ParallelMemberInflater(const unsigned char* data, size_t size, size_t member_start,
                                                      size_t chunk_bytes, int window)
    : m_data(data), m_size(size), m_chunk(std::max<size_t>(env_size("FQD_PINFLATE_CHUNK", chunk_bytes), 1u << 12)),
      m_window(window > 0 ? window : 2 * io_threads() + 2), m_ahead(m_window), m_sync(std::make_shared<Sync>()) {
    const size_t hdr = pinfl::gzip_header_len(data, size, member_start);
    if (!hdr || hdr == pinfl::kCutOff) throw std::runtime_error("gzip error");
    const size_t p = member_start + hdr;
    m_deflate_start = p;
    m_end_bit = (uint64_t)p * 8;
    m_next_boundary = p;
}

inline ParallelMemberInflater::~ParallelMemberInflater() {
    std::unique_lock<std::mutex> g(m_sync->mu);
    m_sync->cv.wait(g, [this] { return m_sync->inflight == 0; });
}

inline void ParallelMemberInflater::launch(ChunkPtr c, bool express) {
    std::shared_ptr<Sync> sy = m_sync;
    const unsigned char* data = m_data; const size_t size = m_size;
    const size_t max_symbols = std::max<size_t>(m_chunk * 1100, 64u << 20);     // deflate expands at most ~1032 x
    { std::lock_guard<std::mutex> g(sy->mu); ++sy->inflight; }
    const size_t keep = (size_t)m_window;
    WorkerPool::shared().submit([sy, c, data, size, max_symbols, keep] {
        { std::lock_guard<std::mutex> g(sy->mu); Sync::take(sy->spare_sym, c->sym); Sync::take(sy->spare_bytes, c->tail); }
        pinfl::decode_chunk(data, size, *c, max_symbols);
        {
            std::lock_guard<std::mutex> g(sy->mu);
            if (c->sym.n == 0) Sync::give(sy->spare_sym, c->sym, keep);
            if (c->tail.n == 0) Sync::give(sy->spare_bytes, c->tail, 2 * keep);
            c->ready = true; --sy->inflight;
        }
        sy->cv.notify_all();
    }, express);
}

inline void ParallelMemberInflater::top_up() {
    while ((int)(m_chunks.size() + m_pieces.size()) < m_ahead.units() && m_next_boundary < m_size) {
        auto c = std::make_shared<pinfl::Chunk>();
        const size_t next = std::min(m_size, m_next_boundary + m_chunk);
        c->from_bit = (uint64_t)m_next_boundary * 8;
        c->stop_bit = (uint64_t)next * 8;
        c->search_end = c->stop_bit;
        if (m_first) { c->exact = true; c->known_window = true; m_first = false; }
        m_next_boundary = next;
        m_chunks.push_back(c);
        launch(c, false);
    }
}

inline void ParallelMemberInflater::accept(ChunkPtr c) {
    using namespace pinfl;
    auto pc = std::make_shared<Piece>();
    pc->window = m_win;
    pc->sym.swap(c->sym);
    pc->tail.swap(c->tail); pc->tail_skip = c->tail_skip;
    // the window the next chunk will need: the last 32 KiB of (window + symbols, markers replaced + byte part)
    const size_t n = pc->sym.size(), nt = pc->tail.n - pc->tail_skip;
    m_sym_bytes += n; m_direct_bytes += nt;
    m_ahead.observed(pc->sym.cap * sizeof(uint16_t) + pc->tail.cap + n);      // symbols + byte part + the resolved copy
    std::vector<uint8_t> nw;
    nw.reserve(kWindow);
    if (nt >= kWindow) {
        nw.assign(pc->tail.data() + pc->tail.n - kWindow, pc->tail.data() + pc->tail.n);
    } else {
        const size_t from_sym = std::min(n, kWindow - nt);
        const size_t keep_old = std::min(m_win.size(), kWindow - nt - from_sym);
        nw.insert(nw.end(), m_win.end() - (ptrdiff_t)keep_old, m_win.end());
        const size_t missing = kWindow - m_win.size();
        for (size_t i = n - from_sym; i < n; ++i) {
            const uint16_t v = pc->sym[i];
            if (v < kMarker) nw.push_back((uint8_t)v);
            else {
                const size_t j = v & 0x7FFFu;
                if (j < missing) throw std::runtime_error("gzip error");       // distance too far back
                nw.push_back(m_win[j - missing]);
            }
        }
        nw.insert(nw.end(), pc->tail.data() + pc->tail_skip, pc->tail.data() + pc->tail.n);
    }
    m_win.swap(nw);
    m_end_bit = c->end_bit;
    ++m_accepted;
    m_pieces.push_back(pc);
    std::shared_ptr<Sync> sy = m_sync;
    { std::lock_guard<std::mutex> g(sy->mu); ++sy->inflight; }
    const size_t keep = (size_t)m_window;
    WorkerPool::shared().submit([sy, pc, keep] {
        { std::lock_guard<std::mutex> g(sy->mu); Sync::take(sy->spare_bytes, pc->bytes); }
        resolve_piece(*pc);
        { std::lock_guard<std::mutex> g(sy->mu); Sync::give(sy->spare_sym, pc->sym, keep); pc->ready = true; --sy->inflight; }
        sy->cv.notify_all();
    }, true);
}

// One step of the chain: look at the oldest planned chunk, accept / drop / bridge.
inline ParallelMemberInflater::Step ParallelMemberInflater::advance_chain() {
    using namespace pinfl;
    if (m_chain_done) return CHAIN_DONE;
    top_up();
    if (m_chunks.empty() && m_next_boundary < m_size) return WINDOW_FULL;   // accepted pieces wait to be read first
    if (m_chunks.empty()) {
        // the plan reached the end of the file without a final block: bridge from the accepted position
        if (m_end_bit >= (uint64_t)m_size * 8) { m_chain_done = true; m_truncated = true; m_end_offset = m_size; return CHAIN_DONE; }
        auto f = std::make_shared<Chunk>();
        f->from_bit = m_end_bit; f->exact = true; f->known_window = m_accepted == 0;
        f->stop_bit = (uint64_t)m_size * 8;
        m_chunks.push_back(f); ++m_fillers;
        launch(f, true);
    }
    ChunkPtr c = m_chunks.front();
    {
        std::unique_lock<std::mutex> g(m_sync->mu);
        m_sync->cv.wait(g, [&] { return c->ready; });
    }
    const bool usable = !c->failed || c->truncated;
    if (c->exact && c->start_bit == m_end_bit) {
        if (!usable) throw std::runtime_error("gzip error");                 // a decode from a true block start failed
    } else if (!usable || c->start_bit == kNone || c->start_bit < m_end_bit) {
        m_chunks.pop_front(); ++m_dropped;                                   // false start, or nothing found: overtaken
        return ADVANCED;
    } else if (c->start_bit > m_end_bit) {
        // a gap between the accepted stream and this chunk's start: decode it, up to that start
        auto f = std::make_shared<Chunk>();
        f->from_bit = m_end_bit; f->exact = true; f->known_window = m_accepted == 0;
        f->stop_bit = c->start_bit;
        m_chunks.push_front(f); ++m_fillers;
        launch(f, true);
        return ADVANCED;
    }
    // starts exactly where the accepted stream ends
    m_chunks.pop_front();
    accept(c);
    if (c->truncated) { m_chain_done = true; m_truncated = true; m_end_offset = m_size; return CHAIN_DONE; }
    if (c->final_block) {
        m_chain_done = true;
        const size_t trailer = (size_t)((m_end_bit + 7) >> 3);
        if (trailer + 8 > m_size) { m_truncated = true; m_end_offset = m_size; }
        else m_end_offset = trailer + 8;
        return CHAIN_DONE;
    }
    return ADVANCED;
}

inline size_t ParallelMemberInflater::read(char* dst, size_t n) {
    size_t got = 0;
    while (got < n && !m_done) {
        if (!m_pieces.empty()) {
            PiecePtr pc = m_pieces.front();
            // keep the chain moving while the head piece is being resolved
            for (;;) {
                { std::lock_guard<std::mutex> g(m_sync->mu); if (pc->ready) break; }
                if (advance_chain() != ADVANCED) {
                    std::unique_lock<std::mutex> g(m_sync->mu);
                    m_sync->cv.wait(g, [&] { return pc->ready; });
                    break;
                }
            }
            if (pc->bad) throw std::runtime_error("gzip error");
            const size_t total = pc->size();
            if (m_piece_off == 0) {
                m_crc = (uint32_t)crc32_combine(m_crc, pc->crc, (z_off_t)total);
                m_len += total;
            }
            while (got < n && m_piece_off < total) {
                // first the resolved symbols, then the part that was decoded as bytes
                const bool in_tail = m_piece_off >= pc->bytes.n;
                const uint8_t* src = in_tail ? pc->tail.data() + pc->tail_skip + (m_piece_off - pc->bytes.n) : pc->bytes.data() + m_piece_off;
                const size_t avail = in_tail ? total - m_piece_off : pc->bytes.n - m_piece_off;
                const size_t k = std::min(n - got, avail);
                std::memcpy(dst + got, src, k);
                got += k; m_piece_off += k;
            }
            if (m_piece_off == total) {
                {
                    std::lock_guard<std::mutex> g(m_sync->mu);
                    Sync::give(m_sync->spare_bytes, pc->bytes, 2 * (size_t)m_window);
                    Sync::give(m_sync->spare_bytes, pc->tail, 2 * (size_t)m_window);
                }
                m_pieces.pop_front(); m_piece_off = 0;
            }
            continue;
        }
        if (m_chain_done) {
            if (!m_truncated) {
                const unsigned char* t = m_data + m_end_offset - 8;
                const uint32_t crc = t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
                const uint32_t isz = t[4] | (uint32_t)t[5] << 8 | (uint32_t)t[6] << 16 | (uint32_t)t[7] << 24;
                if (crc != m_crc || isz != (uint32_t)m_len) throw std::runtime_error("gzip error");
            }
            m_done = true;
            break;
        }
        advance_chain();
    }
    return got;
}

}  // namespace fqdhost
