// TEST INFRASTRUCTURE ONLY -- tiny spin-wait thread pool for the oracle's colour-parallel sweeps
// (this image's gcc has no libgomp).  Mirrors the reference's rayon workers + StageSync barrier
// (src/dynamics/solver/staged_island_solver/sync.rs:39-185): every stage is a barrier.
#pragma once
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

namespace orc {

class Pool {
public:
    static Pool& get() {
        static Pool p;
        return p;
    }
    void set_threads(int n) {
        if (n < 1) n = 1;
        if (n == nthreads_) return;
        stop();
        nthreads_ = n;
        start();
    }
    int threads() const { return nthreads_; }
    // Runs fn(i) for i in [begin, end), statically partitioned; returns when all are done.
    template <class F>
    void parallel_for(int begin, int end, int grain, const F& fn) {
        int n = end - begin;
        if (nthreads_ <= 1 || n <= grain) {
            for (int i = begin; i < end; ++i) fn(i);
            return;
        }
        std::function<void(int)> job = [&](int tid) {
            int chunk = (n + nthreads_ - 1) / nthreads_;
            int b = begin + tid * chunk, e = b + chunk < end ? b + chunk : end;
            for (int i = b; i < e; ++i) fn(i);
        };
        job_ = &job;
        pending_.store(nthreads_ - 1, std::memory_order_release);
        epoch_.fetch_add(1, std::memory_order_acq_rel);
        job(0);
        int spins = 0;
        while (pending_.load(std::memory_order_acquire) != 0) {
            if (++spins > 2000) { std::this_thread::yield(); spins = 1000; }
        }
    }
    ~Pool() { stop(); }

private:
    Pool() {}
    void start() {
        quit_.store(false);
        // The epoch a worker starts from is fixed HERE, not when its thread first runs: a worker that came up
        // after the first parallel_for had already bumped the epoch would otherwise skip that job for ever.
        const uint64_t seen0 = epoch_.load(std::memory_order_acquire);
        for (int t = 1; t < nthreads_; ++t)
            workers_.emplace_back([this, t, seen0] {
                uint64_t seen = seen0;
                for (;;) {
                    uint64_t e;
                    int spins = 0;
                    while ((e = epoch_.load(std::memory_order_acquire)) == seen) {
                        if (quit_.load(std::memory_order_acquire)) return;
                        if (++spins > 2000) { std::this_thread::yield(); spins = 1000; }
                    }
                    seen = e;
                    (*job_)(t);
                    pending_.fetch_sub(1, std::memory_order_acq_rel);
                }
            });
    }
    void stop() {
        quit_.store(true);
        for (auto& w : workers_) w.join();
        workers_.clear();
    }
    int nthreads_ = 1;
    std::vector<std::thread> workers_;
    std::atomic<uint64_t> epoch_{0};
    std::atomic<int> pending_{0};
    std::atomic<bool> quit_{false};
    std::function<void(int)>* job_ = nullptr;
};

}  // namespace orc
