// synth.cuh - counter-based synthetic FASTQ generator for benchmarks and large tests.
// Every byte is a pure function of (seed, record index, mate), so any slice of the stream can be produced
// independently on any GPU.  Record i:
//     "@SYN.%010u <mate>\n" + bases + "\n+\n" + quals + "\n"          (22 + 2*read_len bytes, 322 at 150 bp)
// With probability dup_permille/1000 record i copies the sequence (both mates) of a uniformly chosen earlier
// record j < i; chains are followed to their root, so a prefix of the stream is self-contained.  Quality
// strings are i.i.d. over 4 symbols and never copied, so duplicates differ in ID and quality (the choice of
// representative is visible in the output).  n_permille/1000 of the root reads carry one 'N'.
// variant 1 (loose): 10% of duplicates are truncated by 1..10 bases (both mates), the record is kept at a
//   fixed size by padding the ID description with 'x'.  variant 2 (tail-hamming): 10% of duplicates get up to
//   2 substitutions in the last 10 bases of each mate.
#pragma once
#include "common.cuh"

namespace fqd {

__host__ __device__ __forceinline__ u64 splitmix(u64 x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ u64 synth_rng(u64 seed, u64 i, u64 stream) {
    return splitmix(splitmix(seed ^ (stream * 0xD1342543DE82EF95ull)) + i * 0x9E3779B97F4A7C15ull);
}
// source record of i: i itself if it is not a duplicate
__host__ __device__ __forceinline__ u64 synth_src(u64 seed, u64 i, u32 dup_permille) {
    if (i == 0) return 0;
    u64 r = synth_rng(seed, i, 1);
    if ((r % 1000ull) >= dup_permille) return i;
    return (r >> 20) % i;
}
__host__ __device__ __forceinline__ u64 synth_root(u64 seed, u64 i, u32 dup_permille) {
    for (;;) {
        u64 j = synth_src(seed, i, dup_permille);
        if (j == i) return i;
        i = j;
    }
}
__host__ __device__ __forceinline__ char synth_base(u64 seed, u64 root, int mate, u32 pos, u32 read_len, u32 n_permille) {
    u64 r = synth_rng(seed, root * 64ull + (pos >> 5), 16 + mate);
    char b = "ACGT"[(r >> (2 * (pos & 31u))) & 3u];
    u64 rn = synth_rng(seed, root, 8 + mate);
    if ((rn % 1000ull) < n_permille && ((rn >> 20) % read_len) == pos) b = 'N';
    return b;
}

struct SynthParams {
    u8* out; u64 first; u64 count; u32 read_len; int mate; u64 seed; u32 dup_permille; u32 n_permille; int variant;
};

__global__ void __launch_bounds__(256) k_synth_fastq(const SynthParams p) {
    const u32 rec_bytes = 22u + 2u * p.read_len;
    const u32 lane = threadIdx.x & 31u;
    const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 r = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < p.count; r += warps) {
        const u64 i = p.first + r;
        const u64 root = synth_root(p.seed, i, p.dup_permille);
        const bool is_dup = root != i;
        u32 trunc = 0;           // bases removed from the end (variant 1)
        u32 nsub = 0;            // substitutions in the tail (variant 2)
        const u64 rv = synth_rng(p.seed, i, 3);
        if (is_dup && (rv % 10ull) == 0) {
            if (p.variant == 1) trunc = 1u + (u32)((rv >> 8) % 10ull);
            if (p.variant == 2) nsub = 1u + (u32)((rv >> 8) % 2ull);
        }
        if (trunc >= p.read_len) trunc = 0;
        const u32 L = p.read_len - trunc;
        const u32 idlen = 18u + 2u * trunc;     // keeps the record size fixed
        u8* o = p.out + r * (u64)rec_bytes;
        for (u32 b = lane; b < rec_bytes; b += 32) {
            u8 c;
            if (b < idlen) {
                if (b < 5) c = (u8)"@SYN."[b];
                else if (b < 15) {
                    u64 v = i; u32 d = 14u - b;           // digit position from the right
                    for (u32 k = 0; k < d; ++k) v /= 10ull;
                    c = (u8)('0' + (v % 10ull));
                } else if (b == 15) c = ' ';
                else if (b == 16) c = (u8)('0' + p.mate);
                else if (b == idlen - 1) c = '\n';
                else if (b == 17) c = ' ';
                else c = 'x';
            } else if (b < idlen + L) {
                u32 pos = b - idlen;
                c = (u8)synth_base(p.seed, root, p.mate, pos, p.read_len, p.n_permille);
                if (nsub && pos + 10u >= p.read_len) {
                    // substitute at up to 2 tail positions chosen per (record, mate)
                    u64 rs = synth_rng(p.seed, i, 4 + p.mate);
                    u32 p0 = p.read_len - 1u - (u32)(rs % 10ull);
                    u32 p1 = p.read_len - 1u - (u32)((rs >> 8) % 10ull);
                    if (pos == p0 || (nsub > 1 && pos == p1)) c = (u8)"ACGT"[((rs >> 16) + pos) & 3u];
                }
            } else if (b == idlen + L) c = '\n';
            else if (b == idlen + L + 1) c = '+';
            else if (b == idlen + L + 2) c = '\n';
            else if (b < idlen + 2 * L + 3) {
                u32 pos = b - (idlen + L + 3);
                u64 rq = synth_rng(p.seed, i * 16ull + (pos >> 5), 32 + p.mate);
                c = (u8)"FGHI"[(rq >> (2 * (pos & 31u))) & 3u];
            } else c = '\n';
            o[b] = c;
        }
    }
}

}  // namespace fqd
