#include "parallel.h"
#include <algorithm>
#include <atomic>
#include <climits>
#include <condition_variable>
#include <cstdint>
#include <cstring>
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
// ---- std::sort(v, v + n, less), the same permutation, ties included (libstdc++ __sort: __introsort_loop with depth 2 lg n,
// then __final_insertion_sort), driven range by range:
//   * the two sides of a partition are independent: ranges are partitioned level by level, the rest of every range is left
//     to the library's own std::__introsort_loop / std::__insertion_sort (the final insertion pass never moves anything
//     across a partition cut: elements left of a cut are not greater than elements right of it);
//   * `tie_keys` given (the keys that occur more than once): only a range that still holds two elements with the same key
//     has to be followed literally - everywhere else the ascending order is unique, so any sort will do.  After a partition
//     around pivot key P a tied key < P is in the left part, > P in the right part (== P: both parts are looked at);
//   * the literal partition is restated here (std::__unguarded_partition_pivot = median of first+1 / mid / last-1 moved to
//     first, then the two-pointer loop) so that a LONG scan of one pointer - on the MUM list's shape the down scan crosses
//     the whole range for 34 rounds in a row - can skip the blocks that hold no candidate (BlockIndex below).
// Conservative block summaries of the array being sorted (blocks of 1024 records): lo[b] <= every key in block b <= hi[b].
// A long pointer scan of the partition ("the next record that is not less / not greater than the pivot") skips the blocks
// that cannot hold one; swaps only widen the summaries, so a skipped block never held a candidate - the scan stops at the
// same record as the plain loop.  (On the MUM list's shape the down scan crosses ~130 000 records that are all greater than
// the pivot, 34 rounds in a row.)
template <class T, class KeyOf>
struct BlockIndex {
    static constexpr int SHIFT = 10;
    const T* base = nullptr;
    KeyOf key_of;
    std::vector<int64_t> lo, hi;
    BlockIndex(const T* v, size_t n, KeyOf k, int threads) : base(v), key_of(k) {
        const long nb = (long)((n + (1u << SHIFT) - 1) >> SHIFT);
        lo.assign((size_t)nb, INT64_MAX);
        hi.assign((size_t)nb, INT64_MIN);
        parallel_chunks(nb >= 64 ? threads : 1, nb, [&](long b) {
            int64_t mn = INT64_MAX, mx = INT64_MIN;
            for (size_t i = (size_t)b << SHIFT; i < std::min(n, ((size_t)b + 1) << SHIFT); ++i) { const int64_t k2 = key_of(v[i]); mn = std::min(mn, k2); mx = std::max(mx, k2); }
            lo[(size_t)b] = mn; hi[(size_t)b] = mx;
        });
    }
    inline void wrote(const T* p) {                                          // *p has just been overwritten
        const size_t b = (size_t)(p - base) >> SHIFT;
        const int64_t k2 = key_of(*p);
        if (k2 < lo[b]) lo[b] = k2;
        if (k2 > hi[b]) hi[b] = k2;
    }
};
template <class T, class Less, class BI>
T* find_up(T* lo, const T& pivot, Less less, BI* bi) {                        // first p >= lo with !less(*p, pivot) (exists)
    for (int i = 0; i < 256; ++i, ++lo) if (!less(*lo, pivot)) return lo;
    if (!bi) { while (less(*lo, pivot)) ++lo; return lo; }
    const int64_t pk = bi->key_of(pivot);
    for (;;) {
        const size_t b = (size_t)(lo - bi->base) >> BI::SHIFT;
        T* bend = const_cast<T*>(bi->base) + ((b + 1) << BI::SHIFT);          // (the scan ends before the array does)
        if (bi->hi[b] >= pk) { for (; lo < bend; ++lo) if (!less(*lo, pivot)) return lo; }
        lo = bend;
    }
}
template <class T, class Less, class BI>
T* find_down(T* hi, const T& pivot, Less less, BI* bi) {                      // last p <= hi with !less(pivot, *p) (exists)
    for (int i = 0; i < 256; ++i, --hi) if (!less(pivot, *hi)) return hi;
    if (!bi) { while (less(pivot, *hi)) --hi; return hi; }
    const int64_t pk = bi->key_of(pivot);
    for (;;) {
        const size_t b = (size_t)(hi - bi->base) >> BI::SHIFT;
        T* bbeg = const_cast<T*>(bi->base) + (b << BI::SHIFT);
        if (bi->lo[b] <= pk) { for (; hi >= bbeg; --hi) if (!less(pivot, *hi)) return hi; }
        hi = bbeg - 1;
    }
}
template <class T, class Less, class BI>
T* partition_pivot_literal(T* first, T* last, Less less, BI* bi) {
    T* mid = first + (last - first) / 2;
    T *a = first + 1, *b = mid, *c = last - 1;                               // std::__move_median_to_first(first, a, b, c)
    T* m = nullptr;
    if (less(*a, *b)) {
        if (less(*b, *c)) m = b;
        else if (less(*a, *c)) m = c;
        else m = a;
    } else if (less(*a, *c)) m = a;
    else if (less(*b, *c)) m = c;
    else m = b;
    std::iter_swap(first, m);
    if (bi) { bi->wrote(first); bi->wrote(m); }
    T* lo = first + 1;                                                       // std::__unguarded_partition(first + 1, last, first)
    T* hi = last;
    const T& pivot = *first;
    for (;;) {
        lo = find_up(lo, pivot, less, bi);
        --hi;
        hi = find_down(hi, pivot, less, bi);
        if (!(lo < hi)) return lo;
        std::iter_swap(lo, hi);
        if (bi) { bi->wrote(lo); bi->wrote(hi); }
        ++lo;
    }
}
// any-order sort of a range without equal keys: 32-bit keys in the upper half of a word -> LSD byte radix; else std::sort
template <class T, class Less> void unique_order_sort(T* first, T* last, Less less, std::vector<T>&) { std::sort(first, last, less); }
template <class Less> void unique_order_sort(uint64_t* first, uint64_t* last, Less less, std::vector<uint64_t>& tmp) {
    const size_t n = (size_t)(last - first);
    if (n < 2048) { std::sort(first, last, less); return; }
    tmp.resize(n);
    uint64_t* src = first;
    uint64_t* dst = tmp.data();
    uint64_t all_or = 0, all_and = ~0ull;
    for (size_t i = 0; i < n; ++i) { all_or |= first[i]; all_and &= first[i]; }
    for (int shift = 32; shift < 64; shift += 8) {
        if ((((all_or ^ all_and) >> shift) & 0xff) == 0) continue;            // the byte is the same everywhere
        size_t cnt[257] = {0};
        for (size_t i = 0; i < n; ++i) cnt[((src[i] >> shift) & 0xff) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (size_t i = 0; i < n; ++i) dst[cnt[(src[i] >> shift) & 0xff]++] = src[i];
        std::swap(src, dst);
    }
    if (src != first) std::memcpy(first, src, n * sizeof(uint64_t));
}

template <class T, class Less, class KeyOf>
void literal_sort_impl(T* v, size_t n, int threads, Less less, KeyOf key_of, const int64_t* tie_keys, size_t ntie, const T* hint) {
#if defined(__GLIBCXX__)
    if (n >= 32768 && (threads > 1 || tie_keys)) {
        auto cmp = __gnu_cxx::__ops::__iter_comp_iter(less);
        struct Range { T* first; T* last; long depth; std::vector<int64_t> ties; bool literal; };
        std::vector<Range> ranges, next;
        ranges.push_back(Range{v, v + n, (long)std::__lg((long)n) * 2, std::vector<int64_t>(), true});
        if (tie_keys) ranges[0].ties.assign(tie_keys, tie_keys + ntie);
        const long leaf = (long)std::max<size_t>(4096, n / ((size_t)std::max(1, threads) * 8));
        std::unique_ptr<BlockIndex<T, KeyOf>> blocks;
        if (tie_keys) blocks.reset(new BlockIndex<T, KeyOf>(v, n, key_of, threads));
        std::vector<size_t> big;
        std::vector<T*> cuts;
        auto count_key = [&](const T* f, const T* l, int64_t k) { long c = 0; for (; f != l; ++f) c += key_of(*f) == k; return c; };
        for (;;) {
            // one step of __introsort_loop for every range that must be followed and is still large
            big.clear();
            for (size_t i = 0; i < ranges.size(); ++i) {
                const Range& r = ranges[i];
                if (r.literal && r.depth > 0 && (tie_keys ? r.last - r.first > 16 : r.last - r.first > leaf)) big.push_back(i);
            }
            if (big.empty()) break;
            cuts.assign(big.size(), nullptr);
            if (tie_keys) {
                // (few ranges, possibly long ones: one after the other, long pointer scans skip through the block summaries)
                for (size_t b = 0; b < big.size(); ++b) cuts[b] = partition_pivot_literal(ranges[big[b]].first, ranges[big[b]].last, less, blocks.get());
            } else {
                parallel_chunks(threads, (long)big.size(), [&](long b) {
                    const Range& r = ranges[big[(size_t)b]];
                    cuts[(size_t)b] = std::__unguarded_partition_pivot(r.first, r.last, cmp);
                });
            }
            next.clear();
            size_t bi = 0;
            for (size_t i = 0; i < ranges.size(); ++i) {
                Range& r = ranges[i];
                if (!(bi < big.size() && big[bi] == i)) { next.push_back(std::move(r)); continue; }
                T* cut = cuts[bi++];
                Range L{r.first, cut, r.depth - 1, std::vector<int64_t>(), true}, R{cut, r.last, r.depth - 1, std::vector<int64_t>(), true};
                if (tie_keys) {
                    const int64_t P = key_of(*r.first);                       // (the pivot stays at `first`)
                    for (int64_t k : r.ties) {
                        if (k < P) L.ties.push_back(k);
                        else if (k > P) R.ties.push_back(k);
                        else {
                            if (count_key(L.first, L.last, k) >= 2) L.ties.push_back(k);
                            if (count_key(R.first, R.last, k) >= 2) R.ties.push_back(k);
                        }
                    }
                    L.literal = !L.ties.empty();
                    R.literal = !R.ties.empty();
                }
                next.push_back(std::move(L));
                next.push_back(std::move(R));
            }
            ranges.swap(next);
        }
        // the rest: literal ranges with the library's own loop (heap sort when the depth budget is spent) and their share of the
        // final insertion pass; ranges without equal keys by any sort
        parallel_chunks(threads, (long)ranges.size(), [&](long i) {
            const Range& r = ranges[(size_t)i];
            if (r.literal) {
                std::__introsort_loop(r.first, r.last, r.depth, cmp);
                std::__insertion_sort(r.first, r.last, cmp);
            } else {
                // one possible order: the caller's ascending list says which (unless a tied pair straddles the range's border:
                // then the range holds ONE of the two, and only sorting tells which); else already sorted?; else a radix sort
                const size_t f = (size_t)(r.first - v), l = (size_t)(r.last - v);
                if (hint && l > f && !(f > 0 && key_of(hint[f - 1]) == key_of(hint[f])) && !(l < n && key_of(hint[l - 1]) == key_of(hint[l]))) {
                    std::memcpy((void*)r.first, (const void*)(hint + f), (l - f) * sizeof(T));
                    return;
                }
                if (std::is_sorted(r.first, r.last, less)) return;
                std::vector<T> tmp;
                unique_order_sort(r.first, r.last, less, tmp);
            }
        });
        return;
    }
#endif
    (void)key_of; (void)tie_keys; (void)ntie; (void)hint;
    std::sort(v, v + n, less);
}
}  // namespace

void literal_std_sort_by_first(std::pair<int64_t, int>* v, size_t n, int threads, const int64_t* tie_keys, size_t ntie,
                               const std::pair<int64_t, int>* hint) {
    if (!tie_keys) hint = nullptr;
    typedef std::pair<int64_t, int> P;
    // keys and ids that fit 32 bits each travel as one 64-bit word compared on its upper half: the algorithm sees the same
    // outcome of every comparison and makes the same moves on elements half the size
    bool narrow = n < ((size_t)1 << 31);
    if (narrow) {
        const long per = 16384, nch = ((long)n + per - 1) / per;
        std::atomic<int> wide(0);
        std::unique_ptr<uint64_t[]> w_(new uint64_t[n]), wh_(hint ? new uint64_t[n] : nullptr);      // (uninitialised: filled in parallel)
        uint64_t* const w = w_.get();
        uint64_t* const wh = wh_.get();
        parallel_chunks(n >= 65536 ? threads : 1, nch, [&](long c) {
            for (size_t i = (size_t)c * per; i < std::min(n, (size_t)(c + 1) * per); ++i) {
                if ((uint64_t)v[i].first >> 32 || v[i].second < 0) { wide.store(1, std::memory_order_relaxed); break; }
                w[i] = ((uint64_t)v[i].first << 32) | (uint32_t)v[i].second;
                if (hint) wh[i] = ((uint64_t)hint[i].first << 32) | (uint32_t)hint[i].second;       // (the same records in another order)
            }
        });
        if (!wide.load()) {
            literal_sort_impl(w, n, threads, [](uint64_t x, uint64_t y) { return (x >> 32) < (y >> 32); },
                              [](uint64_t x) { return (int64_t)(x >> 32); }, tie_keys, ntie, (const uint64_t*)wh);
            parallel_chunks(n >= 65536 ? threads : 1, nch, [&](long c) {
                for (size_t i = (size_t)c * per; i < std::min(n, (size_t)(c + 1) * per); ++i) v[i] = P((int64_t)(w[i] >> 32), (int)(uint32_t)w[i]);
            });
            return;
        }
    }
    literal_sort_impl(v, n, threads, [](const P& x, const P& y) { return x.first < y.first; }, [](const P& x) { return x.first; }, tie_keys, ntie, hint);
}

}  // namespace pb200
