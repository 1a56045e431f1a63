// host_worker.h — one persistent helper thread per map (product code).
// The reference spawns a std::thread per frame for its velocity estimation and joins it before the newborn step
// (g-ch/DSP-map include/dsp_dynamic.h:297, 311).  Here the helper is created once; update() hands it the frame's
// estimation job BEFORE it starts enqueueing the frame's kernels and collects the result when the newborn input is
// needed, so the ~30 kernel launches of the first half of the frame are no longer ahead of the estimation on the
// calling thread's critical path.  At frame rates of a few hundred Hz and more the helper is caught while it is still
// spinning after the previous job (no wake-up latency); at camera rates it sleeps on a condition variable in between.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

class HostWorker {
public:
    HostWorker() = default;
    HostWorker(const HostWorker &) = delete;
    HostWorker &operator=(const HostWorker &) = delete;
    ~HostWorker() { stop(); }

    bool running() const { return th_.joinable(); }
    void start() {
        if (running()) return;
        quit_ = false;
        seq_submitted_.store(0, std::memory_order_relaxed);
        seq_done_.store(0, std::memory_order_relaxed);
        th_ = std::thread([this] { loop(); });
    }
    void stop() {
        if (!running()) return;
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    // One job at a time: submit() must be followed by wait() before the next submit().
    void submit(std::function<void()> job) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = std::move(job);
            seq_submitted_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_one();
    }
    // Spins (the caller has nothing else to do: without the helper it would be running the job itself), yielding after
    // a while so an oversubscribed host still makes progress.
    void wait() {
        const unsigned long long want = seq_submitted_.load(std::memory_order_acquire);
        unsigned spins = 0;
        while (seq_done_.load(std::memory_order_acquire) != want) {
            if (++spins > 20000) std::this_thread::yield();
        }
    }

private:
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            // spin briefly for the next job (a steady high-rate loop submits again within a fraction of a millisecond) ...
            const auto t0 = std::chrono::steady_clock::now();
            bool got = false;
            for (unsigned it = 0;; ++it) {
                if (seq_submitted_.load(std::memory_order_acquire) != seen) { got = true; break; }
                if ((it & 255u) == 255u && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(2000)) break;
            }
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                if (!got) cv_.wait(lk, [&] { return quit_ || seq_submitted_.load(std::memory_order_acquire) != seen; });  // ... then sleep
                if (quit_ && seq_submitted_.load(std::memory_order_acquire) == seen) return;
                job = std::move(job_);
                job_ = nullptr;
                seen = seq_submitted_.load(std::memory_order_acquire);
            }
            if (job) {
                try { job(); } catch (...) {}  // the job's owner sees an unchanged output; never leave wait() spinning
            }
            seq_done_.store(seen, std::memory_order_release);
        }
    }

    std::thread th_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::function<void()> job_;
    std::atomic<unsigned long long> seq_submitted_{0}, seq_done_{0};
    bool quit_ = false;
};
