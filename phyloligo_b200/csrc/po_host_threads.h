// Small host-thread helpers shared by the host-only translation units (po_host_sink.cu,
// po_host_mirror.cu).
#pragma once
#include <cstdint>
#include <thread>
#include <vector>

namespace po {

static inline int pick_threads(int threads, int64_t work_bytes) {
    if (threads <= 0) {
        threads = (int)std::thread::hardware_concurrency();
        if (threads <= 0) threads = 1;
    }
    if (threads > 256) threads = 256;
    // below ~1 MB per thread the start-up of a thread costs more than it moves
    const int64_t useful = work_bytes / (1 << 20) + 1;
    if ((int64_t)threads > useful) threads = (int)useful;
    return threads;
}

// run fn(t) for t in [0, threads) on `threads` host threads (the caller's thread is one of them)
template <typename F>
static inline void run_threads(int threads, F fn) {
    if (threads <= 1) {
        fn(0);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve((size_t)threads - 1);
    for (int t = 1; t < threads; ++t) pool.emplace_back(fn, t);
    fn(0);
    for (auto& th : pool) th.join();
}

}  // namespace po
