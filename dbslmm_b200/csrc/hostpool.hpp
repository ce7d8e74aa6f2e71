// hostpool.hpp -- a small persistent pool of host worker threads (one per handle).
// The streaming fit builds its plan while the first panel rows cross PCIe: ~25 short tasks (per-SNP arrays, tile lists, step
// lists per batch, tensor maps) that used to be 25 std::thread creations, issued one after the other by the calling thread
// (20-50 us each).  Workers sleep on a condition variable between bursts and poll for ~50 us after a task before they do.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace dbslmm {

struct TaskGroup {
    std::atomic<int> pending{0};
};

class HostPool {
public:
    HostPool() = default;
    HostPool(const HostPool&) = delete;
    HostPool& operator=(const HostPool&) = delete;
    ~HostPool() { shutdown(); }

    void start(int n) {
        std::lock_guard<std::mutex> lk(m_);
        while ((int)workers_.size() < n) workers_.emplace_back([this]() { loop(); });
    }
    int size() const { return (int)workers_.size(); }

    // Runs f on a worker (inline if the pool has no workers).  g.pending counts the group's unfinished tasks.
    void submit(TaskGroup& g, std::function<void()> f) {
        if (workers_.empty()) { f(); return; }
        g.pending.fetch_add(1, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(m_);
            q_.emplace_back(&g, std::move(f));
        }
        cv_.notify_one();
    }

    // Returns when every task of g has finished; the caller runs queued tasks (of any group) meanwhile.
    void wait(TaskGroup& g) {
        for (;;) {
            if (g.pending.load(std::memory_order_acquire) == 0) return;
            Task t;
            {
                std::unique_lock<std::mutex> lk(m_);
                if (q_.empty()) {
                    if (g.pending.load(std::memory_order_acquire) == 0) return;
                    cv_done_.wait_for(lk, std::chrono::microseconds(50));
                    continue;
                }
                t = std::move(q_.front());
                q_.pop_front();
            }
            run(t);
        }
    }

    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (std::thread& t : workers_) if (t.joinable()) t.join();
        workers_.clear();
        stop_ = false;
    }

private:
    typedef std::pair<TaskGroup*, std::function<void()>> Task;
    void run(Task& t) {
        t.second();
        if (t.first->pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            std::lock_guard<std::mutex> lk(m_);
            cv_done_.notify_all();
        }
    }
    void loop() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(m_);
                if (q_.empty() && !stop_) {
                    // poll briefly (a burst of tasks usually follows), then sleep
                    lk.unlock();
                    const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(50);
                    bool got = false;
                    while (std::chrono::steady_clock::now() < until) {
                        if (!q_empty_relaxed()) { got = true; break; }
                        std::this_thread::yield();
                    }
                    lk.lock();
                    if (!got) cv_.wait(lk, [this]() { return stop_ || !q_.empty(); });
                }
                if (q_.empty()) {
                    if (stop_) return;
                    continue;
                }
                t = std::move(q_.front());
                q_.pop_front();
            }
            run(t);
        }
    }
    bool q_empty_relaxed() {
        std::lock_guard<std::mutex> lk(m_);
        return q_.empty();
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, cv_done_;
    std::deque<Task> q_;
    bool stop_ = false;
};

}  // namespace dbslmm
