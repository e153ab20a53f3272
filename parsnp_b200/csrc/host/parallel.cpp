#include "parallel.h"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <exception>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace pb200 {
namespace {

struct Job {
    const std::function<void(long)>* fn = nullptr;
    long nchunks = 0;
    int helpers = 0;
    std::atomic<long> next{0};
    std::atomic<long> done{0};
    std::atomic<int> joined{0};
    std::mutex m;
    std::condition_variable cv;
    std::exception_ptr err;

    void work() {
        long finished = 0;
        for (long c; (c = next.fetch_add(1)) < nchunks;) {
            try {
                (*fn)(c);
            } catch (...) {
                std::lock_guard<std::mutex> lk(m);
                if (!err) err = std::current_exception();
            }
            ++finished;
        }
        if (finished && done.fetch_add(finished) + finished == nchunks) {
            std::lock_guard<std::mutex> lk(m);
            cv.notify_all();
        }
    }
};

class WorkerPool {
public:
    static WorkerPool& get() {
        static WorkerPool p;
        return p;
    }
    void run(int nthreads, long nchunks, const std::function<void(long)>& fn) {
        bool expected = false;
        if (!busy_.compare_exchange_strong(expected, true)) {        // pool in use by another caller, or a nested call
            for (long c = 0; c < nchunks; ++c) fn(c);
            return;
        }
        std::shared_ptr<Job> job = std::make_shared<Job>();
        job->fn = &fn;
        job->nchunks = nchunks;
        job->helpers = (int)std::min<long>(nthreads, nchunks) - 1;
        {
            std::lock_guard<std::mutex> lk(m_);
            while ((int)workers_.size() < job->helpers) workers_.emplace_back([this]() { loop(); });
            job_ = job;
            ++gen_;
        }
        cv_.notify_all();
        job->work();
        {
            std::unique_lock<std::mutex> lk(job->m);
            job->cv.wait(lk, [&]() { return job->done.load() >= job->nchunks; });
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            job_.reset();
        }
        busy_.store(false);
        if (job->err) std::rethrow_exception(job->err);
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }

private:
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&]() { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                job = job_;
            }
            if (job && job->joined.fetch_add(1) < job->helpers) job->work();
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::vector<std::thread> workers_;
    std::shared_ptr<Job> job_;
    uint64_t gen_ = 0;
    bool stop_ = false;
    std::atomic<bool> busy_{false};
};

}  // namespace

void parallel_run(int nthreads, long nchunks, const std::function<void(long)>& fn) { WorkerPool::get().run(nthreads, nchunks, fn); }

}  // namespace pb200
