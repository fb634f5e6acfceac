// A small process-wide pool of host threads for the feeds' parallel loops (pread slices, the
// high-bit test, the device gzip feed's uploads).  The loops run every few milliseconds on
// buffers of tens of megabytes: starting and joining sixteen threads each time cost a quarter of
// such a loop.
#pragma once

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include <unistd.h>

namespace tdg {

class Pool {
public:
    static Pool &get()
    {
        static Pool p;
        return p;
    }

    // fn(0) .. fn(n - 1), fn(0) on the calling thread; returns when all are done.  One loop at a
    // time: a second caller waits for the first.
    void run(int n, const std::function<void(int)> &fn)
    {
        if (n <= 1) {
            if (n == 1) fn(0);
            return;
        }
        std::lock_guard<std::mutex> one(region_);
        if (getpid() != pid_) {
            // a fork()ed child has the bookkeeping of its parent's workers but not the threads: start afresh
            new (&workers_) std::vector<std::thread>();
            pid_ = getpid();
        }
        grow(n - 1);
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            n_ = n;
            next_ = 1;
            left_ = n - 1;
            gen_++;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return left_ == 0; });
        fn_ = nullptr;
    }

    ~Pool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        if (getpid() == pid_)
            for (auto &t : workers_) t.join();
        else
            new (&workers_) std::vector<std::thread>();
    }

private:
    void grow(int want)
    {
        while ((int)workers_.size() < want) workers_.emplace_back([this] { work(); });
    }

    void work()
    {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int)> *fn = nullptr;
            int idx = -1;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || (gen_ != seen && next_ < n_); });
                if (stop_) return;
                idx = next_++;
                fn = fn_;
                if (next_ >= n_) seen = gen_;
            }
            (*fn)(idx);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--left_ == 0) done_.notify_all();
            }
        }
    }

    std::mutex region_, mu_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    const std::function<void(int)> *fn_ = nullptr;
    int n_ = 0, next_ = 0, left_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
    pid_t pid_ = getpid();
};

}  // namespace tdg
