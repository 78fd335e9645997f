// workers.hpp - the host's worker pool (inflate / deflate / pread / pwrite tasks of pargz.hpp, pinflate.hpp, io.hpp).
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace fqdhost {

// Threads for inflate / deflate / pread: FQD_IO_THREADS, else the hardware's (at most 64).
inline int io_threads() {
    static const int n = [] {
        const char* e = std::getenv("FQD_IO_THREADS");
        int v = e ? std::atoi(e) : 0;
        if (v <= 0) {
            unsigned hc = std::thread::hardware_concurrency();
            v = hc ? (int)std::min(hc, 64u) : 4;
        }
        return std::max(v, 1);
    }();
    return n;
}

// How far a reader may run ahead of its consumer: `full` units (tasks, chunks) as long as their decoded size is
// ordinary, fewer when the input expands enormously (a gigabyte of zeros is a megabyte of deflate), so that what is
// in flight stays within FQD_IO_AHEAD_MB (default 2048) of memory.  Starts small until the first sizes are known.
class RunAhead {
public:
    explicit RunAhead(int full) : m_full(std::max(full, 2)), m_now(std::min(m_full, 4)) {
        const char* e = std::getenv("FQD_IO_AHEAD_MB");
        const long long mb = e ? std::atoll(e) : 0;
        m_budget = (size_t)(mb > 0 ? mb : 2048) << 20;
    }
    int units() const { return m_now; }
    void observed(size_t bytes_in_memory) {           // one unit finished and occupied this much
        m_avg = m_seen ? (m_avg * 7 + bytes_in_memory) / 8 : bytes_in_memory;
        ++m_seen;
        const size_t fit = m_budget / std::max<size_t>(m_avg, 1);
        m_now = (int)std::min<size_t>((size_t)m_full, std::max<size_t>(fit, 2));
    }
private:
    int m_full, m_now;
    size_t m_budget = 0, m_avg = 0, m_seen = 0;
};

// Fixed pool, FIFO with an express lane.  Tasks never wait for other tasks, so sharing one pool between all
// readers and writers cannot deadlock.
class WorkerPool {
public:
    explicit WorkerPool(int n) {
        for (int i = 0; i < n; ++i) m_threads.emplace_back([this] { run(); });
    }
    ~WorkerPool() {
        { std::lock_guard<std::mutex> g(m_mu); m_stop = true; }
        m_cv.notify_all();
        for (auto& t : m_threads) t.join();
    }
    void submit(std::function<void()> f, bool express = false) {
        {
            std::lock_guard<std::mutex> g(m_mu);
            if (express) m_q.push_front(std::move(f)); else m_q.push_back(std::move(f));
        }
        m_cv.notify_one();
    }
    static WorkerPool& shared() {
        static WorkerPool pool(io_threads());
        return pool;
    }
private:
    void run() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> g(m_mu);
                m_cv.wait(g, [this] { return m_stop || !m_q.empty(); });
                if (m_q.empty()) return;          // stop requested and nothing left
                f = std::move(m_q.front()); m_q.pop_front();
            }
            f();
        }
    }
    std::vector<std::thread> m_threads;
    std::deque<std::function<void()>> m_q;
    std::mutex m_mu;
    std::condition_variable m_cv;
    bool m_stop = false;
};

}  // namespace fqdhost
