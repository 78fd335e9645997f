// pargz.hpp - parallel gzip ingest / egress on host threads, the successor of the reference's
// single-threaded Boost.Iostreams gzip filters (src/file_utils.cpp:59-66 input, :83-92 output).
//
// Ingest (ParallelGzSource).  A gzip FILE is a sequence of members, each an independent deflate stream with its own
// CRC-32 / ISIZE trailer; bgzip (BGZF), `cat a.gz b.gz`, Illumina's bcl-convert and most sequencers' lane merges
// produce many.  Members are inflated concurrently:
//   * the compressed file is mapped read-only; candidate member starts are found without inflating anything -
//     exactly, by hopping over BGZF headers (extra subfield "BC" holds the member size), or, for ordinary gzip, by
//     scanning for the header magic 1f 8b 08 with sane flag bits (a superset of the true starts);
//   * the planner cuts the file at candidates into tasks of >= `span` compressed bytes; a worker inflates from its
//     task's start through whole members until a member ends at or beyond the task's limit, and reports where it ended;
//   * the consumer accepts a task only if it starts exactly where the previous accepted one ended (offset 0 is a
//     true start, so by induction every accepted task started at a true member boundary and zlib has verified every
//     member's CRC-32 and length).  Work that started on a false candidate (the magic inside compressed data) fails
//     or is dropped; a gap is closed by a filler task.  The result is exact, never probabilistic.
//   * a member too large for a task (single-member files, pigz output) is inflated serially, streaming, straight
//     into the caller's buffer - today's behaviour - and the parallel plan resumes at the next member boundary.
// Egress (ParallelGzSink): the output is cut into 1 MiB pieces, each deflated as a member of its own on a worker
// and written in order (a valid multi-member gzip file, RFC 1952 section 2.2: `zcat`, Python and this reader read it as one stream).
// The compressed bytes of the reference's .gz outputs are unpinned; the decompressed content is what
// parity is about.
//
// No CUDA in this file: it is unit-tested on the CPU (host/io_selftest.cpp, tests/test_host_io.py).
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "workers.hpp"
#include "pinflate.hpp"

namespace fqdhost {

// Read-only mapping of a regular file; ok() is false for pipes, empty files and mmap failures.
class MappedFile {
public:
    explicit MappedFile(int fd) {
        struct stat sb;
        if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode) || sb.st_size <= 0) return;
        void* p = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) return;
        m_p = (const unsigned char*)p; m_n = (size_t)sb.st_size;
        madvise(p, m_n, MADV_SEQUENTIAL);
    }
    ~MappedFile() { if (m_p) munmap((void*)m_p, m_n); }
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
    bool ok() const { return m_p != nullptr; }
    const unsigned char* data() const { return m_p; }
    size_t size() const { return m_n; }
private:
    const unsigned char* m_p = nullptr;
    size_t m_n = 0;
};

namespace gzdetail {

constexpr size_t kMinMember = 20;      // 10 header + 2 deflate + 8 trailer

inline bool looks_like_member(const unsigned char* d, size_t size, size_t p) {
    return p + kMinMember <= size && d[p] == 0x1f && d[p + 1] == 0x8b && d[p + 2] == 0x08 && (d[p + 3] & 0xe0) == 0;
}

// BGZF member size from the "BC" extra subfield, 0 if this is not a BGZF header.
inline size_t bgzf_member_size(const unsigned char* d, size_t size, size_t p) {
    if (!looks_like_member(d, size, p) || !(d[p + 3] & 4) || p + 12 > size) return 0;
    size_t xlen = d[p + 10] | (size_t)d[p + 11] << 8;
    size_t q = p + 12, xend = q + xlen;
    if (xend > size) return 0;
    while (q + 4 <= xend) {
        size_t slen = d[q + 2] | (size_t)d[q + 3] << 8;
        if (d[q] == 'B' && d[q + 1] == 'C' && slen == 2 && q + 6 <= xend) {
            size_t bsize = (d[q + 4] | (size_t)d[q + 5] << 8) + 1;
            return (bsize >= kMinMember && p + bsize <= size) ? bsize : 0;
        }
        q += 4 + slen;
    }
    return 0;
}

inline bool all_zero(const unsigned char* d, size_t from, size_t to) {
    for (size_t i = from; i < to; ++i) if (d[i]) return false;
    return true;
}

enum Status { PENDING, DONE, TRUNCATED, TOO_BIG, FAILED };

struct Task {
    size_t start = 0, limit = 0;   // planned compressed range [start, limit)
    size_t end = 0;                // where the last complete member ended (DONE / TRUNCATED)
    char* out = nullptr;           // inflated bytes (malloc'd, recycled by the source)
    size_t out_len = 0, out_cap = 0;
    Status status = PENDING;
    bool ready = false;
};

// State shared between a source and its in-flight tasks (tasks may outlive the consumer's interest in them).
struct Shared {
    const unsigned char* data = nullptr;
    size_t size = 0;
    size_t max_overrun = 0, max_out = 0;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<bool> cancel{false};
    int inflight = 0;
};

// Inflate whole members from t.start until one ends at or beyond t.limit (pinflate.hpp's decoder in byte mode: the
// window is known - nothing precedes a member -, CRC-32 and length are checked against each trailer).
inline void inflate_task(Shared& sh, Task& t) {
    using namespace pinfl;
    const unsigned char* d = sh.data;
    RawBuf<uint8_t> out;                                  // borrows the task's (recycled) array
    out.p = (uint8_t*)t.out; out.cap = t.out_cap; out.n = 0;
    const size_t in_stop = std::min(sh.size, t.limit + sh.max_overrun);   // give up beyond this without a member end
    size_t pos = t.start;
    t.end = t.start;
    Status st = PENDING;
    auto sc = std::make_unique<DecodeScratch>();
    if (!out.reserve(std::max<size_t>((t.limit - t.start) * 4 + (64u << 10), 1u << 20))) st = FAILED;
    while (st == PENDING) {
        if (sh.cancel.load(std::memory_order_relaxed)) { st = FAILED; break; }
        const size_t hdr = gzip_header_len(d, sh.size, pos);
        if (hdr == 0) { st = FAILED; break; }
        if (hdr == kCutOff) { st = TRUNCATED; t.end = sh.size; break; }
        const size_t member_out = out.n;
        DecodeResult r;
        decode_blocks<uint8_t>(*sc, d, in_stop, (uint64_t)(pos + hdr) * 8, kNone, true, sh.max_out, out, r);
        const bool out_of_input = r.failed && r.truncated;
        size_t trailer = (size_t)((r.end_bit + 7) >> 3);
        if (out_of_input || (!r.failed && trailer + 8 > in_stop)) {
            // the member runs beyond the task's allowance, or beyond the end of the file
            st = in_stop == sh.size ? TRUNCATED : TOO_BIG;
            if (st == TRUNCATED) t.end = sh.size;
            break;
        }
        if (r.failed) { st = r.too_big ? TOO_BIG : FAILED; break; }
        const uint32_t crc = d[trailer] | (uint32_t)d[trailer + 1] << 8 | (uint32_t)d[trailer + 2] << 16 | (uint32_t)d[trailer + 3] << 24;
        const uint32_t isz = d[trailer + 4] | (uint32_t)d[trailer + 5] << 8 | (uint32_t)d[trailer + 6] << 16 | (uint32_t)d[trailer + 7] << 24;
        const size_t mlen = out.n - member_out;
        if (isz != (uint32_t)mlen || crc != (uint32_t)crc32_z(0L, (const Bytef*)out.p + member_out, mlen)) { st = FAILED; break; }
        if (out.n > sh.max_out) { st = TOO_BIG; break; }
        pos = trailer + 8;
        t.end = pos;
        if (pos >= t.limit || pos >= sh.size) { st = DONE; break; }
        if (!looks_like_member(d, sh.size, pos)) { st = DONE; break; }   // padding or garbage: the consumer decides
    }
    t.out = (char*)out.p; t.out_len = st == FAILED || st == TOO_BIG ? 0 : out.n; t.out_cap = out.cap;
    out.p = nullptr; out.n = out.cap = 0;
    t.status = st;
}

}  // namespace gzdetail

class ParallelGzSource {
public:
    // `span`: compressed bytes per task; `window`: tasks planned ahead of the consumer (0 = 2 x threads)
    ParallelGzSource(const unsigned char* data, size_t size, size_t span = 2u << 20, int window = 0)
        : m_sh(std::make_shared<gzdetail::Shared>()), m_span(std::max<size_t>(env_size("FQD_GZ_SPAN", span), 64)),
          m_window(window > 0 ? window : 2 * io_threads()), m_ahead(m_window) {
        m_sh->data = data; m_sh->size = size;
        m_sh->max_overrun = std::max<size_t>(16u << 20, 4 * m_span);
        m_sh->max_out = 512u << 20;
        m_max_task = env_size("FQD_GZ_MAX_TASK", std::max<size_t>(8u << 20, 4 * m_span));
        const char* e = std::getenv("FQD_PINFLATE");           // 0: members too large for a task are inflated serially
        m_block_parallel = !(e && e[0] == '0');
        m_bgzf = gzdetail::bgzf_member_size(data, size, 0) != 0;
        m_hop = 0;
    }
    ~ParallelGzSource() {
        m_sh->cancel = true;
        std::unique_lock<std::mutex> g(m_sh->mu);
        m_sh->cv.wait(g, [this] { return m_sh->inflight == 0; });
        g.unlock();
        for (auto& t : m_tasks) std::free(t->out);
        for (auto& t : m_zombies) std::free(t->out);
        if (m_cur) std::free(m_cur->out);
        for (auto& b : m_spare) std::free(b.first);
        m_member.reset();
        if (m_serial) inflateEnd(&m_z);
    }
    ParallelGzSource(const ParallelGzSource&) = delete;
    ParallelGzSource& operator=(const ParallelGzSource&) = delete;

    // read up to n bytes; returns the number read (< n only at end of input); throws on a corrupt stream
    size_t read(char* dst, size_t n) {
        using namespace gzdetail;
        size_t got = 0;
        while (got < n && !m_eof) {
            if (m_serial) { got += serial_read(dst + got, n - got); continue; }
            if (m_member) {
                got += m_member->read(dst + got, n - got);
                m_hint_in.store(m_hint_in_base + m_member->accepted_compressed(), std::memory_order_relaxed);
                m_hint_out.store(m_hint_out_base + m_member->accepted_output(), std::memory_order_relaxed);
                if (m_member->done()) {
                    m_pos = m_member->end_offset();
                    m_member_chunks += m_member->chunks_accepted();
                    m_symbol_bytes += m_member->symbol_bytes(); m_direct_bytes += m_member->direct_bytes();
                    ++m_serial_members;
                    m_hint_in_base = m_hint_in.load(std::memory_order_relaxed); m_hint_out_base = m_hint_out.load(std::memory_order_relaxed);
                    m_member.reset();
                }
                continue;
            }
            if (m_cur) {
                size_t k = std::min(n - got, m_cur->out_len - m_cur_off);
                std::memcpy(dst + got, m_cur->out + m_cur_off, k);
                got += k; m_cur_off += k;
                if (m_cur_off == m_cur->out_len) {
                    m_pos = m_cur->end;
                    recycle(*m_cur);
                    m_cur.reset();
                }
                continue;
            }
            if (m_sh->size - std::min(m_pos, m_sh->size) < kMinMember) { m_eof = true; break; }   // end (or a cut-off header)
            if (!looks_like_member(m_sh->data, m_sh->size, m_pos)) {
                // zero padding after the last member is harmless (gzip(1) ignores it); anything else is an error
                if (m_pos > 0 && all_zero(m_sh->data, m_pos, m_sh->size)) { m_eof = true; break; }
                throw std::runtime_error("gzip error");
            }
            // tasks that start before the accepted position began on a false candidate (or were overtaken)
            while (!m_tasks.empty() && m_tasks.front()->start < m_pos) drop_front();
            if (m_plan_pos < m_pos) m_plan_pos = m_pos;
            if (m_tasks.empty() || m_tasks.front()->start != m_pos) {
                // nothing was planned to start here: close the gap up to the next planned start
                size_t limit = m_tasks.empty() ? plan_limit(m_pos) : m_tasks.front()->start;
                if (m_tasks.empty()) m_plan_pos = limit;
                launch(m_pos, limit, true);
            }
            top_up();
            std::shared_ptr<Task> t = m_tasks.front();
            {
                std::unique_lock<std::mutex> g(m_sh->mu);
                m_sh->cv.wait(g, [&] { return t->ready; });
            }
            m_tasks.pop_front();
            switch (t->status) {
            case DONE: case TRUNCATED:
                ++m_parallel_tasks;
                m_ahead.observed(t->out_cap);
                m_hint_in_base += t->end - t->start; m_hint_out_base += t->out_len;
                m_hint_in.store(m_hint_in_base, std::memory_order_relaxed); m_hint_out.store(m_hint_out_base, std::memory_order_relaxed);
                m_cur = t; m_cur_off = 0;
                if (t->out_len == 0) { m_pos = t->end; recycle(*t); m_cur.reset(); }
                break;
            case TOO_BIG:
                recycle(*t);
                if (m_block_parallel) m_member.reset(new ParallelMemberInflater(m_sh->data, m_sh->size, m_pos));
                else begin_serial();
                break;
            default:
                recycle(*t);
                throw std::runtime_error("gzip error");
            }
        }
        return got;
    }
    bool eof() const { return m_eof; }
    // uncompressed / compressed bytes of what has been inflated so far (0 until 1 MiB of input is accounted for);
    // may be read from another thread - the drivers size the device tables from it
    double expansion_hint() const {
        const size_t in = m_hint_in.load(std::memory_order_relaxed), out = m_hint_out.load(std::memory_order_relaxed);
        return in >= (1u << 20) ? (double)out / (double)in : 0.0;
    }
    // statistics for tests / the selftest's bench
    size_t parallel_tasks() const { return m_parallel_tasks; }
    size_t serial_members() const { return m_serial_members; }     // members too large for a task (serial or pinflate)
    size_t symbol_bytes() const { return m_symbol_bytes; }
    size_t direct_bytes() const { return m_direct_bytes; }
    size_t member_chunks() const { return m_member_chunks; }       // ... and the block-parallel chunks they were cut into
    size_t dropped_tasks() const { return m_dropped; }
    bool bgzf() const { return m_bgzf; }

private:
    using TaskPtr = std::shared_ptr<gzdetail::Task>;

    // smallest candidate member start >= from (or the file size)
    size_t next_candidate(size_t from) {
        using namespace gzdetail;
        const unsigned char* d = m_sh->data; const size_t size = m_sh->size;
        if (from >= size) return size;
        if (m_bgzf) {
            if (m_hop > from) return m_hop;      // the planner asks in increasing order; m_hop is a true start
            while (m_hop < from) {
                size_t b = bgzf_member_size(d, size, m_hop);
                if (!b) { m_bgzf = false; break; }           // an ordinary member (or the end): scan from here on
                m_hop += b;
            }
            if (m_bgzf) return std::min(m_hop, size);
        }
        size_t p = from;
        while (p + kMinMember <= size) {
            const void* q = memchr(d + p, 0x1f, size - kMinMember + 1 - p);
            if (!q) break;
            p = (size_t)((const unsigned char*)q - d);
            if (looks_like_member(d, size, p)) return p;
            ++p;
        }
        return size;
    }
    size_t plan_limit(size_t start) { return next_candidate(std::min(m_sh->size, start + m_span)); }

    void launch(size_t start, size_t limit, bool express) {
        using namespace gzdetail;
        TaskPtr t = std::make_shared<Task>();
        t->start = start; t->limit = limit;
        if (express) m_tasks.push_front(t); else m_tasks.push_back(t);
        if (limit - start > m_max_task) {      // one huge member: not worth a speculative attempt
            t->status = TOO_BIG; t->ready = true;
            return;
        }
        if (!m_spare.empty()) { t->out = m_spare.back().first; t->out_cap = m_spare.back().second; m_spare.pop_back(); }
        std::shared_ptr<Shared> sh = m_sh;
        { std::lock_guard<std::mutex> g(sh->mu); ++sh->inflight; }
        WorkerPool::shared().submit([sh, t] {
            inflate_task(*sh, *t);
            { std::lock_guard<std::mutex> g(sh->mu); t->ready = true; --sh->inflight; }
            sh->cv.notify_all();
        }, express);
    }
    void top_up() {
        while ((int)m_tasks.size() < m_ahead.units() && m_plan_pos < m_sh->size) {
            size_t limit = plan_limit(m_plan_pos);
            launch(m_plan_pos, limit, false);
            m_plan_pos = limit;
        }
    }
    void drop_front() {
        TaskPtr t = m_tasks.front();
        m_tasks.pop_front();
        ++m_dropped;
        bool ready;
        { std::lock_guard<std::mutex> g(m_sh->mu); ready = t->ready; }
        if (ready) recycle(*t);
        else m_zombies.push_back(t);          // still running: its buffer is reclaimed later
        reap();
    }
    void reap() {
        for (size_t i = 0; i < m_zombies.size();) {
            bool ready;
            { std::lock_guard<std::mutex> g(m_sh->mu); ready = m_zombies[i]->ready; }
            if (ready) { recycle(*m_zombies[i]); m_zombies[i] = m_zombies.back(); m_zombies.pop_back(); } else ++i;
        }
    }
    void recycle(gzdetail::Task& t) {
        if (t.out) {
            if (m_spare.size() < (size_t)m_window && t.out_cap <= (256u << 20)) m_spare.emplace_back(t.out, t.out_cap);
            else std::free(t.out);
            t.out = nullptr; t.out_cap = 0;
        }
    }

    // One member too large for a task: stream it into the caller's buffer, resume the plan where it ends.
    void begin_serial() {
        std::memset(&m_z, 0, sizeof m_z);
        if (inflateInit2(&m_z, 15 + 16) != Z_OK) throw std::runtime_error("zlib: inflateInit2 failed");
        m_serial = true;
        m_serial_pos = m_pos;
    }
    size_t serial_read(char* dst, size_t n) {
        using namespace gzdetail;
        const unsigned char* d = m_sh->data; const size_t size = m_sh->size;
        size_t got = 0;
        while (got < n) {
            if (m_serial_pos >= size) { end_serial(size); m_eof = true; break; }      // truncated inside the member
            m_z.next_in = (Bytef*)(d + m_serial_pos);
            m_z.avail_in = (uInt)std::min<size_t>(size - m_serial_pos, 1u << 30);
            m_z.next_out = (Bytef*)dst + got;
            m_z.avail_out = (uInt)std::min<size_t>(n - got, 1u << 30);
            const uInt before = m_z.avail_out;
            int rc = inflate(&m_z, Z_NO_FLUSH);
            got += before - m_z.avail_out;
            m_serial_pos = (size_t)(m_z.next_in - d);
            if (rc == Z_STREAM_END) {
                ++m_serial_members;
                end_serial(m_serial_pos);
                break;                                   // back to the parallel plan
            }
            if (rc != Z_OK && rc != Z_BUF_ERROR) { end_serial(m_serial_pos); throw std::runtime_error("gzip error"); }
        }
        return got;
    }
    void end_serial(size_t pos) {
        inflateEnd(&m_z);
        m_serial = false;
        m_pos = pos;
    }

    std::shared_ptr<gzdetail::Shared> m_sh;
    size_t m_span, m_max_task = 0;
    int m_window;
    RunAhead m_ahead;
    bool m_bgzf = false;
    size_t m_hop = 0;                 // BGZF: a true member start, advanced by header hops
    size_t m_pos = 0;                 // compressed offset up to which the output has been accepted
    size_t m_plan_pos = 0;            // limit of the last planned task
    std::deque<TaskPtr> m_tasks;      // planned, in file order
    std::vector<TaskPtr> m_zombies;
    TaskPtr m_cur;                    // accepted task being copied out
    size_t m_cur_off = 0;
    std::vector<std::pair<char*, size_t>> m_spare;
    bool m_eof = false, m_serial = false, m_block_parallel = true;
    std::unique_ptr<ParallelMemberInflater> m_member;
    size_t m_member_chunks = 0, m_symbol_bytes = 0, m_direct_bytes = 0;
    std::atomic<size_t> m_hint_in{0}, m_hint_out{0};
    size_t m_hint_in_base = 0, m_hint_out_base = 0;
    z_stream m_z;
    size_t m_serial_pos = 0;
    size_t m_parallel_tasks = 0, m_serial_members = 0, m_dropped = 0;
};

// Ordered multi-member gzip writer.
class ParallelGzSink {
public:
    explicit ParallelGzSink(FILE* f, size_t piece = 1u << 20, int level = Z_DEFAULT_COMPRESSION)
        : m_f(f), m_piece(std::max<size_t>(piece, 1)), m_level(level), m_window(2 * io_threads()),
          m_sh(std::make_shared<Sync>()) {
        m_buf.reserve(m_piece);
    }
    ~ParallelGzSink() { try { finish(); } catch (...) {} }
    ParallelGzSink(const ParallelGzSink&) = delete;
    ParallelGzSink& operator=(const ParallelGzSink&) = delete;

    void write(const char* p, size_t n) {
        while (n) {
            size_t k = std::min(n, m_piece - m_buf.size());
            m_buf.insert(m_buf.end(), p, p + k);
            p += k; n -= k;
            if (m_buf.size() == m_piece) flush_piece();
        }
    }
    // compress what is left, wait for every piece, write them in order (the FILE stays open)
    void finish() {
        if (m_finished) return;
        m_finished = true;
        if (!m_buf.empty() || m_members == 0) flush_piece();      // an empty output is one empty member
        drain(true);
    }
    bool failed() const { return m_failed; }

private:
    struct Piece { std::vector<char> in; std::vector<unsigned char> out; bool ready = false, ok = false; };
    struct Sync { std::mutex mu; std::condition_variable cv; };

    void flush_piece() {
        auto pc = std::make_shared<Piece>();
        pc->in.swap(m_buf);
        m_buf.reserve(m_piece);
        ++m_members;
        m_q.push_back(pc);
        std::shared_ptr<Sync> sh = m_sh;
        const int level = m_level;
        WorkerPool::shared().submit([pc, sh, level] {
            z_stream z;
            std::memset(&z, 0, sizeof z);
            bool ok = deflateInit2(&z, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) == Z_OK;
            if (ok) {
                pc->out.resize(deflateBound(&z, (uLong)pc->in.size()) + 32);
                z.next_in = (Bytef*)pc->in.data(); z.avail_in = (uInt)pc->in.size();
                z.next_out = pc->out.data(); z.avail_out = (uInt)pc->out.size();
                ok = deflate(&z, Z_FINISH) == Z_STREAM_END;
                pc->out.resize(pc->out.size() - z.avail_out);
                deflateEnd(&z);
            }
            std::vector<char>().swap(pc->in);
            { std::lock_guard<std::mutex> g(sh->mu); pc->ok = ok; pc->ready = true; }
            sh->cv.notify_all();
        });
        drain(false);
    }
    // write finished pieces at the head of the queue; block while more than `window` are pending (or for all)
    void drain(bool all) {
        while (!m_q.empty()) {
            auto& pc = m_q.front();
            {
                std::unique_lock<std::mutex> g(m_sh->mu);
                if (!pc->ready) {
                    if (!all && (int)m_q.size() <= m_window) return;
                    m_sh->cv.wait(g, [&] { return pc->ready; });
                }
            }
            if (!pc->ok) m_failed = true;
            else if (m_f && std::fwrite(pc->out.data(), 1, pc->out.size(), m_f) != pc->out.size()) m_failed = true;
            m_q.pop_front();
        }
    }

    FILE* m_f;
    size_t m_piece;
    int m_level, m_window;
    std::shared_ptr<Sync> m_sh;
    std::vector<char> m_buf;
    std::deque<std::shared_ptr<Piece>> m_q;
    size_t m_members = 0;
    bool m_finished = false, m_failed = false;
};

// Plain files: `parts` concurrent pread()s fill one block (page cache -> pinned memory at more than one core's
// memcpy rate).  Returns the bytes read (short only at end of file).
inline size_t parallel_pread(int fd, char* dst, size_t n, size_t offset) {
    auto pread_all = [fd](char* p, size_t len, size_t off) -> size_t {
        size_t got = 0;
        while (got < len) {
            ssize_t r = pread(fd, p + got, len - got, (off_t)(off + got));
            if (r <= 0) break;
            got += (size_t)r;
        }
        return got;
    };
    const size_t min_part = 4u << 20;
    int parts = (int)std::min<size_t>(std::min(io_threads(), 8), n / min_part);
    if (parts <= 1) return pread_all(dst, n, offset);
    struct Join { std::mutex mu; std::condition_variable cv; int left; };
    auto join = std::make_shared<Join>();
    join->left = parts - 1;
    const size_t per = (n / parts + 4095) & ~(size_t)4095;
    auto gots = std::make_shared<std::vector<size_t>>(parts, 0);
    for (int i = 1; i < parts; ++i) {
        size_t b = std::min(n, per * i), e = std::min(n, per * (i + 1));
        WorkerPool::shared().submit([=] {
            (*gots)[i] = pread_all(dst + b, e - b, offset + b);
            { std::lock_guard<std::mutex> g(join->mu); --join->left; }
            join->cv.notify_all();
        });
    }
    (*gots)[0] = pread_all(dst, std::min(n, per), offset);
    { std::unique_lock<std::mutex> g(join->mu); join->cv.wait(g, [&] { return join->left == 0; }); }
    size_t total = 0;
    for (int i = 0; i < parts; ++i) {
        size_t want = std::min(n, per * (i + 1)) - std::min(n, per * i);
        total += (*gots)[i];
        if ((*gots)[i] < want) break;      // end of file inside this part
    }
    return total;
}

}  // namespace fqdhost
