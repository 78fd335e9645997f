// replay.hpp - the host's second look at its input, for whole-input jobs whose raw bytes do not stay on the device
// (fqd_discard_input, include/fqd.h): the device keeps key rows + record tables, sorts / scans / joins, and answers with
// (offset, length) lists; the WRITTEN records are then fetched here.  The reference meets the same problem with its
// external sort - sorted chunks of 2/3 memlimit written to `<tmp>/chunks/<k>.tmp` and merged from disk
// (src/external_sort.hpp:88-117,120-207; src/file_utils.cpp:116-130 creates and removes the temporary directory).
// Here: a plain regular file is simply mapped (nothing is written); anything that can be read only once or is not the
// bytes themselves (a pipe, a FIFO, a ".gz") is SPOOLED - every inflated block is written to an unlinked temporary
// file (by an ordered asynchronous writer, several workers per block) while it streams to the device - and that file
// is mapped afterwards.  Host memory stays what -m allows: the
// mappings are page cache, which the kernel drops under pressure.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <stdexcept>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace fqdhost {

class InputReplay {
public:
    // `plain_regular`: the input is an uncompressed regular file - map it, spool nothing.
    // `spool_dir`: where the temporary file of the other kind goes (FQD_SPOOL_DIR, else next to the output file).
    InputReplay(const std::string& in_name, bool plain_regular, const std::string& spool_dir) {
        if (plain_regular) {
            m_fd = ::open(in_name.c_str(), O_RDONLY);
            if (m_fd < 0) throw std::runtime_error("cannot open " + in_name + " a second time: " + std::strerror(errno));
            m_spooling = false;
            return;
        }
        m_spooling = true;
        std::string dir = spool_dir.empty() ? std::string(".") : spool_dir;
        std::string tmpl = dir + "/fqd_spool_XXXXXX";
        m_fd = ::mkstemp(tmpl.data());
        if (m_fd < 0) throw std::runtime_error("cannot create the input spool in " + dir + ": " + std::strerror(errno) + " (set FQD_SPOOL_DIR)");
        ::unlink(tmpl.c_str());          // the file lives as long as this descriptor: nothing to clean up, even after a crash
    }
    ~InputReplay() {
        if (m_p) ::munmap((void*)m_p, m_n);
        if (m_fd >= 0) ::close(m_fd);
    }
    InputReplay(const InputReplay&) = delete;
    InputReplay& operator=(const InputReplay&) = delete;

    bool spooling() const { return m_spooling; }
    // the input is complete (and, when spooling, written through proc_path() by the driver's ordered writer): map it
    void seal() {
        if (m_p) return;
        struct stat sb;
        if (fstat(m_fd, &sb) != 0) throw std::runtime_error("fstat failed on the input");
        m_n = (size_t)sb.st_size;
        if (m_n == 0) return;
        void* p = ::mmap(nullptr, m_n, PROT_READ, MAP_SHARED, m_fd, 0);
        if (p == MAP_FAILED) throw std::runtime_error(std::string("cannot map the input for the output gather: ") + std::strerror(errno));
        m_p = (const char*)p;
    }
    // Map the pages in now, several threads at once (the gather touches every page of the mapping in random order; one
    // batched walk costs a fraction of a fault per record).  Skipped when the input is larger than half of the RAM.
    void prefault(int threads) {
#ifndef MADV_POPULATE_READ
#define MADV_POPULATE_READ 22
#endif
        if (!m_p || m_n < ((size_t)64 << 20)) return;
        const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGESIZE);
        if (pages > 0 && psz > 0 && m_n > (size_t)pages * (size_t)psz / 2) return;
        threads = std::max(1, std::min(threads, 16));
        const size_t per = ((m_n / (size_t)threads) + 4095) & ~(size_t)4095;
        std::vector<std::thread> ts;
        for (int t = 0; t < threads; ++t) {
            const size_t lo = std::min(m_n, per * (size_t)t), hi = std::min(m_n, per * (size_t)(t + 1));
            if (lo >= hi) break;
            ts.emplace_back([this, lo, hi] { (void)::madvise((void*)(m_p + lo), hi - lo, MADV_POPULATE_READ); });
        }
        for (auto& t : ts) t.join();
    }
    const char* data() const { return m_p; }
    size_t size() const { return m_n; }
    // A name under which the (unlinked) spool can be opened again while this object lives: a job that has to start over
    // - wider key rows, byte keys, longer tags - reads a pipe's bytes from here the second time.
    std::string proc_path() const { return "/proc/self/fd/" + std::to_string(m_fd); }

private:
    int m_fd = -1;
    bool m_spooling = false;
    const char* m_p = nullptr;
    size_t m_n = 0;
};

inline std::string spool_dir_for(const std::string& out_name) {
    if (const char* e = std::getenv("FQD_SPOOL_DIR")) if (*e) return e;
    const std::string parent = std::filesystem::path(out_name).parent_path().string();
    if (parent.empty()) return ".";
    if (parent.rfind("/dev", 0) == 0 || parent.rfind("/proc", 0) == 0) {      // -o /dev/stdout, -o >(...)
        const char* t = std::getenv("TMPDIR");
        return t && *t ? t : "/tmp";
    }
    return parent;
}

}  // namespace fqdhost
