// io_selftest.cpp - CPU-only harness around io.hpp / pargz.hpp (no CUDA: the two pinned-memory entry points of
// the C ABI are replaced by malloc here).  tests/test_host_io.py drives it:
//   io_selftest cat  <file> [block_bytes]   file -> BlockReader ring -> stdout; "stats ..." line on stderr
//   io_selftest put  <file>                 stdin -> OutputFile (plain or .gz by extension)
//   io_selftest bench <file> [repeats]      InputFile::read into one buffer; prints uncompressed GB/s as JSON
//   io_selftest filter <in> <out> [block]   the drivers' pipeline (dup_remover.cpp: BlockReader ring -> MateStream ->
//                                           chunk -> survivor runs -> AsyncWriter -> OutputFile, blocks released
//                                           behind the writes) with a stand-in for the engine's verdict: 4-line
//                                           records, record i is dropped when i % 3 == 1; prints GB/s as JSON
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "io.hpp"

extern "C" int fqd_host_alloc(void** p, size_t bytes) { *p = std::malloc(bytes); return *p ? FQD_OK : FQD_ERR_CUDA; }
extern "C" int fqd_host_free(void* p) { std::free(p); return FQD_OK; }

using namespace fqdhost;

static int cmd_cat(const std::string& name, size_t block) {
    // the ring, exactly as the drivers use it (dup_remover.cpp)
    BlockReader reader(name, block);
    size_t total = 0, blocks = 0;
    while (Block* b = reader.next()) {
        if (b->len && std::fwrite(b->data(), 1, b->len, stdout) != b->len) return 2;
        total += b->len; ++blocks;
        reader.release(b);
    }
    std::fflush(stdout);
    std::fprintf(stderr, "stats bytes=%zu blocks=%zu\n", total, blocks);
    return 0;
}

static int cmd_stat(const std::string& name) {
    // InputFile alone: how the archive was decoded
    InputFile f(name);
    std::vector<char> buf(8u << 20);
    size_t total = 0;
    while (!f.eof()) total += f.read(buf.data(), buf.size());
    const ParallelGzSource* ps = f.parallel_source();
    std::printf("{\"bytes\": %zu, \"parallel\": %s, \"bgzf\": %s, \"tasks\": %zu, \"serial_members\": %zu, \"dropped\": %zu, \"member_chunks\": %zu, \"symbol_bytes\": %zu, \"direct_bytes\": %zu, \"expansion\": %.4f}\n",
                total, ps ? "true" : "false", ps && ps->bgzf() ? "true" : "false", ps ? ps->parallel_tasks() : (size_t)0,
                ps ? ps->serial_members() : (size_t)0, ps ? ps->dropped_tasks() : (size_t)0, ps ? ps->member_chunks() : (size_t)0, ps ? ps->symbol_bytes() : (size_t)0, ps ? ps->direct_bytes() : (size_t)0, f.expansion_hint());
    return 0;
}

static int cmd_put(const std::string& name) {
    OutputFile out(name);
    std::vector<char> buf(1u << 20);
    size_t n;
    // odd-sized writes on purpose: pieces must not depend on the caller's write sizes
    size_t want = 777;
    while ((n = std::fread(buf.data(), 1, std::min(want, buf.size()), stdin)) > 0) {
        out.write(buf.data(), n);
        want = want * 3 + 1;
        if (want > buf.size()) want = 1000;
    }
    out.close();
    return 0;
}

static int cmd_bench(const std::string& name, int repeats) {
    std::vector<char> buf(64u << 20);
    double best = 0; size_t bytes = 0;
    for (int r = 0; r < repeats; ++r) {
        auto t0 = std::chrono::steady_clock::now();
        InputFile f(name);
        size_t total = 0;
        while (!f.eof()) total += f.read(buf.data(), buf.size());
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        bytes = total;
        best = std::max(best, (double)total / s / 1e9);
    }
    std::printf("{\"file\": \"%s\", \"bytes\": %zu, \"threads\": %d, \"GBps\": %.3f}\n", name.c_str(), bytes, io_threads(), best);
    return 0;
}

static int cmd_filter(const std::string& in, const std::string& out_name, size_t block) {
    auto t0 = std::chrono::steady_clock::now();
    OutputFile out(out_name);
    MateStream ms;
    ms.reader.reset(new BlockReader(in, block));
    AsyncWriter writer(out);
    {
        AsyncWriter* w = &writer; BlockReader* r = ms.reader.get();
        ms.release = [w, r](Block* b) { w->then([r, b] { r->release(b); }); };
    }
    std::vector<uint32_t> rec_start;
    std::vector<uint8_t> dup;
    std::vector<Run> runs;
    size_t global = 0, in_bytes = 0;
    for (;;) {
        if (ms.len < block / 2) ms.refill();
        rec_start.clear(); dup.clear();
        rec_start.push_back(0);
        const char* p = ms.ptr; const char* end = ms.ptr + ms.len;
        // FQD_SELFTEST_FIXED=<bytes>: records of a fixed size (no newline search: measures the I/O pipeline alone)
        static const size_t fixed = std::getenv("FQD_SELFTEST_FIXED") ? (size_t)std::atoll(std::getenv("FQD_SELFTEST_FIXED")) : 0;
        for (; fixed;) {
            if ((size_t)(end - p) < fixed) break;
            p += fixed;
            rec_start.push_back((uint32_t)(p - ms.ptr));
            dup.push_back((global + dup.size()) % 3 == 1);
        }
        for (; !fixed;) {
            const char* q = p; int lines = 0;
            while (lines < 4 && q < end) { const char* nl = (const char*)memchr(q, '\n', end - q); if (!nl) break; q = nl + 1; ++lines; }
            if (lines < 4) break;
            p = q;
            rec_start.push_back((uint32_t)(p - ms.ptr));
            dup.push_back((global + dup.size()) % 3 == 1);
        }
        const size_t n = dup.size();
        const size_t bytes = survivor_runs(rec_start.data(), dup.data(), n, runs);
        writer.write_runs(ms.ptr, std::move(runs), bytes);
        runs = std::vector<Run>();
        global += n;
        in_bytes += rec_start[n];
        ms.ptr += rec_start[n]; ms.len -= rec_start[n];
        if (n == 0 && !ms.refill()) break;
    }
    writer.drain();
    out.close();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("{\"records\": %zu, \"in_bytes\": %zu, \"threads\": %d, \"GBps\": %.3f}\n", global, in_bytes, io_threads(), (double)in_bytes / s / 1e9);
    return 0;
}

// the block decoder of pinflate.hpp alone, one thread: the whole first member as ONE chunk, then the search
static int cmd_rawinflate(const std::string& name) {
    int fd = ::open(name.c_str(), O_RDONLY);
    MappedFile mf(fd);
    if (!mf.ok()) return 2;
    ParallelMemberInflater probe(mf.data(), mf.size(), 0);     // parses the header
    (void)probe;
    pinfl::Chunk c;
    size_t p = 10;                                             // plain header assumed (no name / extra)
    if (mf.data()[3] & 8) { while (mf.data()[p]) ++p; ++p; }
    c.from_bit = p * 8; c.stop_bit = mf.size() * 8; c.exact = true; c.known_window = true;
    {   // output array sized from ISIZE and touched once, so that the timing is the decoder's and not the page faults'
        const unsigned char* t = mf.data() + mf.size() - 4;
        size_t isize = t[0] | (size_t)t[1] << 8 | (size_t)t[2] << 16 | (size_t)t[3] << 24;
        if (c.sym.reserve(isize + (1u << 20))) std::memset(c.sym.data(), 0, c.sym.cap * sizeof(uint16_t));
    }
    auto t0 = std::chrono::steady_clock::now();
    { pinfl::DecodeResult r; auto sc = std::make_unique<pinfl::DecodeScratch>(); pinfl::decode_blocks<uint16_t>(*sc, mf.data(), mf.size(), c.from_bit, c.stop_bit, true, (size_t)1 << 34, c.sym, r); c.failed = r.failed; c.final_block = r.final_block; }
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("{\"symbols\": %zu, \"failed\": %d, \"final\": %d, \"decode_GBps\": %.3f", c.sym.size(), (int)c.failed, (int)c.final_block, c.sym.size() / s / 1e9);
    t0 = std::chrono::steady_clock::now();
    size_t found = 0, tries = 0;
    for (size_t b = mf.size() / 8; b < mf.size(); b += mf.size() / 8, ++tries)
        found += pinfl::find_block(mf.data(), mf.size(), b * 8, (b + (1u << 20)) * 8) != pinfl::kNone;
    s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf(", \"searches\": %zu, \"found\": %zu, \"ms_per_search\": %.3f", tries, found, 1e3 * s / std::max<size_t>(tries, 1));
    pinfl::Piece pc; pc.sym.swap(c.sym);
    t0 = std::chrono::steady_clock::now();
    pinfl::resolve_piece(pc);
    s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf(", \"resolve_GBps\": %.3f}\n", pc.bytes.size() / s / 1e9);
    ::close(fd);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: io_selftest cat|stat|put|bench <file> [arg]\n"); return 64; }
    const std::string cmd = argv[1], name = argv[2];
    try {
        if (cmd == "cat") return cmd_cat(name, argc > 3 ? (size_t)std::atoll(argv[3]) : (4u << 20));
        if (cmd == "stat") return cmd_stat(name);
        if (cmd == "rawinflate") return cmd_rawinflate(name);
        if (cmd == "put") return cmd_put(name);
        if (cmd == "filter" && argc > 3) return cmd_filter(name, argv[3], argc > 4 ? (size_t)std::atoll(argv[4]) : (4u << 20));
        if (cmd == "bench") return cmd_bench(name, argc > 3 ? std::atoi(argv[3]) : 3);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 64;
}
