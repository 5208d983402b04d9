// Minimal fork-join helper for the host-side passes over the lensmap (merge, finish, tile plan).
#pragma once

#include <atomic>
#include <thread>
#include <vector>

namespace blinky {

// calls f(i) for every i in [0, n); dynamic scheduling over `threads` threads (<= 1: inline)
template <class F>
void parallel_for(int n, int threads, F f) {
    if (threads > n) threads = n;
    if (threads <= 1) {
        for (int i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    pool.reserve(static_cast<size_t>(threads));
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&]() {
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= n) break;
                f(i);
            }
        });
    for (auto &th : pool) th.join();
}

}  // namespace blinky
