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
