// io.hpp - host-side byte I/O of the drop-in binary.  Same observable behaviour as the reference's
// FileUtils (src/file_utils.{hpp,cpp}): ".gz" is decided by the file name alone, independently per file
// (:42-48,71-92); a missing input prints "Cannot open file <name>" and throws (:110-121 of the header).
// Different mechanics: zlib directly instead of Boost.Iostreams, and a reader THREAD that fills a small ring
// of PINNED blocks (fqd_host_alloc) so that inflate / read() overlap the H2D copies and the kernels - the
// successor of BufferedInput<T>'s single malloc'd block (src/bufferedinput.hpp:28-36,57-88).
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <cerrno>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <filesystem>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fqd.h"
#include "pargz.hpp"

namespace fqdhost {

// FQD_TRACE=1: wall-clock checkpoints of the host pipeline on stderr (milliseconds since the first call)
inline void trace(const char* what) {
    static const bool on = std::getenv("FQD_TRACE") != nullptr;
    if (!on) return;
    static const auto t0 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[host-trace] %9.1f ms  %s\n",
                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), what);
}

inline bool has_gz_ext(const std::string& name) { return std::filesystem::path(name).extension() == ".gz"; }

// Plain files are read with concurrent pread()s, ".gz" files are inflated member-parallel (pargz.hpp); pipes,
// single-member archives and FQD_IO_THREADS=1 take the serial zlib path.
class InputFile {
public:
    explicit InputFile(const std::string& name) : m_name(name), m_gz(has_gz_ext(name)) {
        m_fd = ::open(name.c_str(), O_RDONLY);
        if (m_fd < 0) {
            std::cerr << "Cannot open file " << name << std::endl;
            throw std::runtime_error("File does not exist or cannot be opened!");
        }
        const bool parallel = io_threads() > 1;
        if (m_gz) {
            if (parallel) {
                m_map.reset(new MappedFile(m_fd));
                if (m_map->ok()) { m_par.reset(new ParallelGzSource(m_map->data(), m_map->size())); return; }
                m_map.reset();
            }
            m_f = fdopen(m_fd, "rb");
            if (!m_f) throw std::runtime_error("File does not exist or cannot be opened!");
            std::memset(&m_z, 0, sizeof m_z);
            if (inflateInit2(&m_z, 15 + 16) != Z_OK) throw std::runtime_error("zlib: inflateInit2 failed");
            m_zinit = true;
            m_in.resize(1 << 20);
        } else {
            struct stat sb;
            m_regular = parallel && fstat(m_fd, &sb) == 0 && S_ISREG(sb.st_mode);
            if (!m_regular) {
                m_f = fdopen(m_fd, "rb");
                if (!m_f) throw std::runtime_error("File does not exist or cannot be opened!");
            }
        }
    }
    ~InputFile() {
        if (m_zinit) inflateEnd(&m_z);
        m_par.reset(); m_map.reset();
        if (m_f) std::fclose(m_f); else if (m_fd >= 0) ::close(m_fd);
    }
    InputFile(const InputFile&) = delete;
    InputFile& operator=(const InputFile&) = delete;
    // read up to n bytes; returns the number read (< n only at end of input)
    size_t read(char* dst, size_t n) {
        if (m_par) {
            size_t got = m_par->read(dst, n);
            if (m_par->eof()) m_eof = true;
            return got;
        }
        if (!m_gz) {
            size_t got;
            if (m_regular) { got = parallel_pread(m_fd, dst, n, m_off); m_off += got; }
            else got = std::fread(dst, 1, n, m_f);
            if (got < n) m_eof = true;
            return got;
        }
        size_t got = 0;
        while (got < n && !m_eof) {
            if (m_z.avail_in == 0) {
                m_z.avail_in = (uInt)std::fread(m_in.data(), 1, m_in.size(), m_f);
                m_z.next_in = (Bytef*)m_in.data();
                if (m_z.avail_in == 0) { m_eof = true; break; }
            }
            m_z.next_out = (Bytef*)dst + got;
            size_t want = n - got;
            m_z.avail_out = (uInt)std::min<size_t>(want, 1u << 30);
            uInt before = m_z.avail_out;
            int rc = inflate(&m_z, Z_NO_FLUSH);
            got += before - m_z.avail_out;
            if (rc == Z_STREAM_END) inflateReset(&m_z);           // multi-member gzip
            else if (rc != Z_OK && rc != Z_BUF_ERROR) throw std::runtime_error("gzip error");
        }
        return got;
    }
    bool eof() const { return m_eof; }
    const ParallelGzSource* parallel_source() const { return m_par.get(); }
    double expansion_hint() const { return m_par ? m_par->expansion_hint() : 0.0; }
private:
    std::string m_name;
    bool m_gz, m_eof = false, m_regular = false, m_zinit = false;
    int m_fd = -1;
    size_t m_off = 0;
    FILE* m_f = nullptr;
    z_stream m_z;
    std::vector<char> m_in;
    std::unique_ptr<MappedFile> m_map;
    std::unique_ptr<ParallelGzSource> m_par;
};

// One pinned block: [head room | data].  The consumer copies the (small) unconsumed tail of the previous
// block into the head room so that every chunk handed to fqd_push is contiguous.
struct Block {
    char* base = nullptr;
    size_t head = 0, cap = 0, len = 0;
    bool last = false;
    char* data() const { return base + head; }
};

class BlockReader {
public:
    BlockReader(const std::string& name, size_t block_bytes, int nbuf = 3) : m_file(name), m_block(block_bytes) {
        for (int i = 0; i < nbuf; ++i) {
            Block* b = new Block();
            b->head = block_bytes; b->cap = block_bytes;
            void* p = nullptr;
            if (fqd_host_alloc(&p, b->head + b->cap + 64) != FQD_OK) throw std::runtime_error("pinned host allocation failed (no usable CUDA device? this build has no CPU path)");
            b->base = (char*)p;
            m_all.push_back(b);
            m_free.push_back(b);
        }
        m_thread = std::thread([this] { this->run(); });
    }
    ~BlockReader() {
        { std::lock_guard<std::mutex> g(m_mu); m_stop = true; }
        m_cv.notify_all();
        if (m_thread.joinable()) m_thread.join();
        for (Block* b : m_all) { fqd_host_free(b->base); delete b; }
    }
    // next filled block, or nullptr after the last one; rethrows reader errors
    Block* next() {
        std::unique_lock<std::mutex> g(m_mu);
        m_cv.wait(g, [this] { return !m_full.empty() || m_done || m_err; });
        if (m_err) std::rethrow_exception(m_err);
        if (m_full.empty()) return nullptr;
        Block* b = m_full.front(); m_full.pop_front();
        return b;
    }
    void release(Block* b) {
        { std::lock_guard<std::mutex> g(m_mu); m_free.push_back(b); }
        m_cv.notify_all();
    }
    size_t block_bytes() const { return m_block; }
    // .gz input: measured uncompressed / compressed ratio of what has been inflated so far, 0 if unknown
    double expansion_hint() const { return m_file.expansion_hint(); }
private:
    void run() {
        try {
            for (;;) {
                Block* b;
                {
                    std::unique_lock<std::mutex> g(m_mu);
                    m_cv.wait(g, [this] { return !m_free.empty() || m_stop; });
                    if (m_stop) return;
                    b = m_free.front(); m_free.pop_front();
                }
                b->len = m_file.read(b->data(), b->cap);
                b->last = m_file.eof();
                bool last = b->last;
                {
                    std::lock_guard<std::mutex> g(m_mu);
                    if (b->len > 0 || !last) m_full.push_back(b); else m_free.push_back(b);
                    if (last) m_done = true;
                }
                m_cv.notify_all();
                if (last) return;
            }
        } catch (...) {
            std::lock_guard<std::mutex> g(m_mu);
            m_err = std::current_exception();
            m_done = true;
            m_cv.notify_all();
        }
    }
    InputFile m_file;
    size_t m_block;
    std::vector<Block*> m_all;
    std::deque<Block*> m_free, m_full;
    std::mutex m_mu;
    std::condition_variable m_cv;
    std::thread m_thread;
    bool m_stop = false, m_done = false;
    std::exception_ptr m_err;
};

// The consumer's view of one input file: [ptr, ptr + len) is contiguous, unconsumed input.  refill() appends the
// next block behind the unconsumed tail (the tail is copied into the block's head room - the successor of
// BufferedInput::refresh's memmove, src/bufferedinput.hpp:66-74).  A block that is left behind goes back to its
// reader through `release`, which the drivers route through the AsyncWriter that may still be reading from it.
struct MateStream {
    std::unique_ptr<BlockReader> reader;
    std::function<void(Block*)> release;   // default: straight back to the reader
    Block* cur = nullptr;      // block that holds [ptr, ptr+len)
    Block* peeked = nullptr;   // next block, fetched early to look at its first byte
    char* ptr = nullptr;
    size_t len = 0;
    bool no_more = false;      // reader exhausted
    Block* fetch() {
        if (peeked) { Block* b = peeked; peeked = nullptr; return b; }
        if (no_more) return nullptr;
        Block* b = reader->next();
        if (!b) no_more = true;
        return b;
    }
    bool refill() {
        Block* b = fetch();
        if (!b) return false;
        if (len > b->head) throw std::runtime_error("Not enough memory to read a single object!");
        char* dst = b->data() - len;
        if (len) memcpy(dst, ptr, len);
        if (cur) { if (release) release(cur); else reader->release(cur); }
        cur = b; ptr = dst; len += b->len;
        return true;
    }
    // first byte that follows [ptr, ptr+len) in the file, or -1 at end of input
    int peek_next_byte() {
        if (!peeked) {
            if (no_more) return -1;
            peeked = reader->next();
            if (!peeked) { no_more = true; return -1; }
        }
        return peeked->len ? (unsigned char)peeked->data()[0] : -1;
    }
};

// FQD_GZ_LEVEL = 1..9: deflate level of ".gz" outputs (default: zlib's default, as the reference's gzip filter)
inline int gz_level() {
    static const int level = [] {
        const char* e = std::getenv("FQD_GZ_LEVEL");
        const int v = e ? std::atoi(e) : 0;
        return v >= 1 && v <= 9 ? v : Z_DEFAULT_COMPRESSION;
    }();
    return level;
}

// A stretch of bytes to write: `len` bytes at base + off, landing `out_off` bytes after the start of the job.
struct Run { size_t off, len, out_off; };

// UniversalOutputFile (src/file_utils.cpp:83-92): plain or gzip by extension.  ".gz" output is deflated on the
// worker pool as a multi-member archive (pargz.hpp; with FQD_IO_THREADS=1 by zlib's gzwrite).  Plain output goes
// through pwrite(): small writes are buffered, a list of runs (the survivors of a chunk) is gathered and written by
// several workers at once, each at the file offset its bytes belong to.
class OutputFile {
public:
    explicit OutputFile(const std::string& name) : m_gz(has_gz_ext(name)) {
        if (m_gz && io_threads() <= 1) {
            const std::string mode = gz_level() == Z_DEFAULT_COMPRESSION ? "wb" : "wb" + std::to_string(gz_level());
            m_g = gzopen(name.c_str(), mode.c_str());
            if (m_g) gzbuffer(m_g, 1 << 20);
        } else if (m_gz) {
            m_f = std::fopen(name.c_str(), "wb");
            if (m_f) std::setvbuf(m_f, nullptr, _IOFBF, 4 << 20);
            if (m_f) m_sink.reset(new ParallelGzSink(m_f, 1u << 20, gz_level()));
        } else {
            m_fd = ::open(name.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
            m_wbuf.reserve(kBuf);
            // /dev/stdout, a FIFO, `-o >(pigz > x.gz)`: no offsets there - every byte goes through write() in order
            struct stat sb;
            m_seekable = m_fd >= 0 && fstat(m_fd, &sb) == 0 && S_ISREG(sb.st_mode) && ::lseek(m_fd, 0, SEEK_CUR) != (off_t)-1;
        }
    }
    ~OutputFile() { close(); }
    OutputFile(const OutputFile&) = delete;
    OutputFile& operator=(const OutputFile&) = delete;
    void write(const char* p, size_t n) {
        if (m_g) {
            while (n) { unsigned w = (unsigned)std::min<size_t>(n, 1u << 30); gzwrite(m_g, p, w); p += w; n -= w; }
        } else if (m_sink) {
            m_sink->write(p, n);
        } else if (m_fd >= 0) {
            if (n >= kBuf) { flush(); put_all(p, n, m_off); m_off += n; return; }
            if (m_wbuf.size() + n > kBuf) flush();
            m_wbuf.insert(m_wbuf.end(), p, p + n);
        }
    }
    // runs must be sorted by out_off and dense: out_off[i+1] = out_off[i] + len[i], out_off[0] = 0
    void write_runs(const char* base, const Run* runs, size_t n_runs, size_t total) {
        if (m_fd < 0 || !m_seekable || total < (8u << 20) || io_threads() <= 1) {
            for (size_t i = 0; i < n_runs; ++i) write(base + runs[i].off, runs[i].len);
            return;
        }
        flush();
        const int parts = (int)std::min<size_t>(std::min(io_threads(), 16), total / (2u << 20));
        const size_t per = (total + parts - 1) / parts;
        struct Join { std::mutex mu; std::condition_variable cv; int left; };
        auto join = std::make_shared<Join>();      // shared: a worker may still be inside notify when the waiter returns
        join->left = parts;
        const size_t file_off = m_off;
        auto pwrite_all = [this](const char* q, size_t k, size_t at) { put_all(q, k, at); };
        for (int k = 0; k < parts; ++k) {
            const size_t lo = std::min(total, per * k), hi = std::min(total, per * (k + 1));
            auto work = [=] {
                // first run that reaches beyond lo
                size_t a = 0, b = n_runs;
                while (a < b) { size_t m = (a + b) / 2; if (runs[m].out_off + runs[m].len <= lo) a = m + 1; else b = m; }
                std::vector<char> buf;
                size_t buf_at = lo, pos = lo;
                for (size_t i = a; i < n_runs && pos < hi; ++i) {
                    const size_t skip = pos - runs[i].out_off;
                    const size_t k2 = std::min(runs[i].len - skip, hi - pos);
                    const char* src = base + runs[i].off + skip;
                    if (k2 >= (1u << 20)) {                     // long stretch: straight from the source
                        if (!buf.empty()) { pwrite_all(buf.data(), buf.size(), file_off + buf_at); buf.clear(); }
                        pwrite_all(src, k2, file_off + pos);
                        buf_at = pos + k2;
                    } else {
                        if (buf.empty()) { buf.reserve(kBuf); buf_at = pos; }
                        if (buf.size() + k2 > kBuf) { pwrite_all(buf.data(), buf.size(), file_off + buf_at); buf_at += buf.size(); buf.clear(); }
                        buf.insert(buf.end(), src, src + k2);
                    }
                    pos += k2;
                }
                if (!buf.empty()) pwrite_all(buf.data(), buf.size(), file_off + buf_at);
                { std::lock_guard<std::mutex> g(join->mu); --join->left; }
                join->cv.notify_all();
            };
            if (k + 1 < parts) WorkerPool::shared().submit(work); else work();
        }
        { std::unique_lock<std::mutex> g(join->mu); join->cv.wait(g, [&] { return join->left == 0; }); }
        m_off += total;
    }
    // errno of the first failed write of a plain output (0 = none); the drivers turn it into a failed run after close()
    int error() const { return m_err.load(); }
    void close() {
        if (m_sink) { m_sink->finish(); m_sink.reset(); }
        if (m_g) { gzclose(m_g); m_g = nullptr; }
        if (m_f) { std::fclose(m_f); m_f = nullptr; }
        if (m_fd >= 0) { flush(); ::close(m_fd); m_fd = -1; }
    }
private:
    static constexpr size_t kBuf = 4u << 20;
    // Regular files: pwrite at the offset the bytes belong to (several workers at once).  Anything else: write() in
    // call order (write_runs takes its sequential path there).  The first failure is kept for error().
    void put_all(const char* p, size_t n, size_t off) {
        while (n) {
            const ssize_t w = m_seekable ? ::pwrite(m_fd, p, n, (off_t)off) : ::write(m_fd, p, n);
            if (w < 0 && errno == EINTR) continue;
            if (w <= 0) { int expected = 0; m_err.compare_exchange_strong(expected, w < 0 ? errno : EIO); return; }
            p += w; n -= (size_t)w; off += (size_t)w;
        }
    }
    void flush() {
        if (m_fd >= 0 && !m_wbuf.empty()) { put_all(m_wbuf.data(), m_wbuf.size(), m_off); m_off += m_wbuf.size(); m_wbuf.clear(); }
    }
    bool m_gz;
    bool m_seekable = false;
    std::atomic<int> m_err{0};
    FILE* m_f = nullptr;
    gzFile m_g = nullptr;
    int m_fd = -1;
    size_t m_off = 0;
    std::vector<char> m_wbuf;
    std::unique_ptr<ParallelGzSink> m_sink;
};

// Writes behind the caller's back, in order: one thread per output file executes the queued jobs (lists of runs
// out of a pinned input block, or callbacks such as "hand this block back to its reader") first in, first out, so
// the next chunk is pushed to the device while the survivors of the previous one are still being written.
class AsyncWriter {
public:
    explicit AsyncWriter(OutputFile& out) : m_out(out) { m_thread = std::thread([this] { run(); }); }
    ~AsyncWriter() {
        { std::lock_guard<std::mutex> g(m_mu); m_stop = true; }
        m_cv.notify_all();
        if (m_thread.joinable()) m_thread.join();      // finishes what is queued first
    }
    void write_runs(const char* base, std::vector<Run>&& runs, size_t total) {
        if (runs.empty()) return;
        Job j; j.base = base; j.runs = std::move(runs); j.total = total;
        push(std::move(j));
    }
    // bytes the job owns (a vector: its heap block does not move with the job)
    void write_owned(std::vector<char>&& bytes) {
        if (bytes.empty()) return;
        Job j; j.owned = std::move(bytes);
        push(std::move(j));
    }
    void then(std::function<void()> f) { Job j; j.fn = std::move(f); push(std::move(j)); }
    void drain() {
        std::unique_lock<std::mutex> g(m_mu);
        m_cv.wait(g, [this] { return m_q.empty() && !m_busy; });
    }
private:
    struct Job { const char* base = nullptr; std::vector<Run> runs; size_t total = 0; std::function<void()> fn; std::vector<char> owned; };
    void push(Job&& j) {
        { std::lock_guard<std::mutex> g(m_mu); m_q.push_back(std::move(j)); }
        m_cv.notify_all();
    }
    void run() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> g(m_mu);
                m_cv.wait(g, [this] { return m_stop || !m_q.empty(); });
                if (m_q.empty()) return;
                j = std::move(m_q.front()); m_q.pop_front();
                m_busy = true;
            }
            if (j.fn) j.fn();
            else if (!j.owned.empty()) m_out.write(j.owned.data(), j.owned.size());
            else m_out.write_runs(j.base, j.runs.data(), j.runs.size(), j.total);
            { std::lock_guard<std::mutex> g(m_mu); m_busy = false; }
            m_cv.notify_all();
        }
    }
    OutputFile& m_out;
    std::deque<Job> m_q;
    std::mutex m_mu;
    std::condition_variable m_cv;
    std::thread m_thread;
    bool m_stop = false, m_busy = false;
};

// Survivors of a chunk as runs of consecutive written records (rec_start has n + 1 entries).
inline size_t survivor_runs(const uint32_t* rec_start, const uint8_t* dup, size_t n, std::vector<Run>& runs) {
    runs.clear();
    size_t total = 0, i = 0;
    while (i < n) {
        if (dup[i]) { ++i; continue; }
        size_t j = i;
        while (j + 1 < n && !dup[j + 1]) ++j;
        const size_t len = rec_start[j + 1] - rec_start[i];
        runs.push_back(Run{rec_start[i], len, total});
        total += len;
        i = j + 1;
    }
    return total;
}

// ClusterFile (src/file_utils.cpp:98-112): "<out>.clusters", head = ID line, member = "--" + ID line.
class ClusterFile {
public:
    void open(const std::string& base) { m_f = std::fopen((base + ".clusters").c_str(), "wb"); }
    ~ClusterFile() { if (m_f) std::fclose(m_f); }
    void write_cluster_head(const char* p, size_t n) { if (m_f) std::fwrite(p, 1, n, m_f); }
    void write_cluster_item(const char* p, size_t n) { if (m_f) { std::fwrite("--", 1, 2, m_f); std::fwrite(p, 1, n, m_f); } }
private:
    FILE* m_f = nullptr;
};

}  // namespace fqdhost
