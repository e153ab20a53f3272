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
    static WorkerPool& get(int which) {
        static WorkerPool p[2];
        return p[which ? 1 : 0];
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

namespace { thread_local int tl_pool = 0; }
void parallel_use_second_pool(bool on) { tl_pool = on ? 1 : 0; }
void parallel_run(int nthreads, long nchunks, const std::function<void(long)>& fn) { WorkerPool::get(tl_pool).run(nthreads, nchunks, fn); }

namespace {
// std::sort's steps (libstdc++ __sort: __introsort_loop with depth 2 lg n, then __final_insertion_sort) on [v, v + n), driven
// range by range; T = element type, less = the comparison the caller's std::sort call would have used
template <class T, class Less>
void literal_sort_impl(T* v, size_t n, int threads, Less less) {
#if defined(__GLIBCXX__)
    if (threads > 1 && n >= 32768) {
        auto cmp = __gnu_cxx::__ops::__iter_comp_iter(less);
        struct Range { T* first; T* last; long depth; };
        std::vector<Range> ranges(1, Range{v, v + n, (long)std::__lg((long)n) * 2}), next;
        const long leaf = (long)std::max<size_t>(4096, n / ((size_t)threads * 8));
        std::vector<size_t> big;
        std::vector<T*> cuts;
        for (;;) {
            // one step of __introsort_loop for every range that is still large: partition; both sides go on with depth - 1
            big.clear();
            for (size_t i = 0; i < ranges.size(); ++i)
                if (ranges[i].last - ranges[i].first > leaf && ranges[i].depth > 0) big.push_back(i);
            if (big.empty()) break;
            cuts.assign(big.size(), nullptr);
            parallel_chunks(threads, (long)big.size(), [&](long b) {
                const Range& r = ranges[big[(size_t)b]];
                cuts[(size_t)b] = std::__unguarded_partition_pivot(r.first, r.last, cmp);
            });
            next.clear();
            size_t bi = 0;
            for (size_t i = 0; i < ranges.size(); ++i) {
                const Range& r = ranges[i];
                if (bi < big.size() && big[bi] == i) {
                    next.push_back(Range{r.first, cuts[bi], r.depth - 1});
                    next.push_back(Range{cuts[bi], r.last, r.depth - 1});
                    ++bi;
                } else next.push_back(r);
            }
            ranges.swap(next);
        }
        // the rest of every range with the library's own loop (heap sort when the depth budget is spent), then its share of the
        // final insertion pass: elements left of a partition cut are never greater than elements right of it, so the global
        // (unguarded) insertion would stop at the cut where the guarded one stops at the range's first element
        parallel_chunks(threads, (long)ranges.size(), [&](long i) {
            const Range& r = ranges[(size_t)i];
            std::__introsort_loop(r.first, r.last, r.depth, cmp);
            std::__insertion_sort(r.first, r.last, cmp);
        });
        return;
    }
#endif
    std::sort(v, v + n, less);
}
}  // namespace

void literal_std_sort_by_first(std::pair<int64_t, int>* v, size_t n, int threads) {
    typedef std::pair<int64_t, int> P;
    // keys and ids that fit 32 bits each travel as one 64-bit word compared on its upper half: the algorithm sees the same
    // outcome of every comparison and makes the same moves on elements half the size
    bool narrow = n < ((size_t)1 << 31);
    if (narrow) {
        const long per = 16384, nch = ((long)n + per - 1) / per;
        std::atomic<int> wide(0);
        std::vector<uint64_t> w(n);
        parallel_chunks(n >= 65536 ? threads : 1, nch, [&](long c) {
            for (size_t i = (size_t)c * per; i < std::min(n, (size_t)(c + 1) * per); ++i) {
                if ((uint64_t)v[i].first >> 32 || v[i].second < 0) { wide.store(1, std::memory_order_relaxed); break; }
                w[i] = ((uint64_t)v[i].first << 32) | (uint32_t)v[i].second;
            }
        });
        if (!wide.load()) {
            literal_sort_impl(w.data(), n, threads, [](uint64_t x, uint64_t y) { return (x >> 32) < (y >> 32); });
            parallel_chunks(n >= 65536 ? threads : 1, nch, [&](long c) {
                for (size_t i = (size_t)c * per; i < std::min(n, (size_t)(c + 1) * per); ++i) v[i] = P((int64_t)(w[i] >> 32), (int)(uint32_t)w[i]);
            });
            return;
        }
    }
    literal_sort_impl(v, n, threads, [](const P& x, const P& y) { return x.first < y.first; });
}

}  // namespace pb200
