// Aligner::doWork (src/parsnp.cpp:173-317) on several host threads - and still the reference's result, bit for bit.
//
// The reference pops the region with the smallest start[0], accepts its candidates against `mumlayout` (trim, src/parsnp.cpp:
// 1399-1477), pushes the sub-regions around every new MUM (determineRegion, 1199-1290) and repeats.  Sub-regions lie inside
// their parent on the reference, so the whole subtree of one initial region (a gap between two anchors) is finished before the
// next initial region is popped: the reference's order is "gap after gap, each gap depth first by start[0]".
//
// A TASK is a run of consecutive initial regions.  Its SPAN in genome g is the interval between the anchor bits that bound
// its first and its last region there.  If the anchors are collinear the spans of different tasks are disjoint in every
// genome, and everything a task reads or writes for a forward-strand candidate lies inside its own span: tasks commute.
// What does not commute is the reference's handling of reverse-strand genomes inside sub-regions: TMum mirrors the start on
// the WHOLE genome (src/TMum.cpp:35), so trimming reads - and an accepted MUM writes - `mumlayout` at an unrelated place,
// i.e. in some other task's span.  Those FOREIGN accesses are what this file is about:
//
//   * two bitmaps: `truth` (everything written so far) and S = anchors + foreign writes ("what a span looks like before its
//     own task has run");
//   * a foreign READ by task k of a position owned by task m:  m < k  -> wait until m is done, read truth (the reference has
//     finished m's gaps by then);  m > k or unowned -> read S (the reference has not started them);  logged per task;
//   * a foreign WRITE is only made by the LOWEST unfinished task (it waits for that), goes to truth and S, and then every
//     task above it that has started and either owns one of the written bits or has logged a read of them is stale: all
//     tasks above the writer are stopped, their spans are restored from S, and they run again (a "restart": ~0.3 per
//     alignment on configs[1], where 1 accepted MUM in 8000 is such a write);
//   * ties on start[0] between different regions make the reference's order depend on its unstable std::sort over the whole
//     queue (the sequential loop replays that call literally): the first task that meets one ends the parallel part; the
//     layout is rebuilt from the MUMs of the finished tasks below it and process_queue_exact continues from there.  Anchors
//     that are not collinear (rearranged genomes: spans out of order, or initial regions that tie) skip the parallel part.
//
// Tasks are handed out in ascending order, so the lowest running task never waits and the scheme cannot deadlock.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include "accept_impl.h"
#include "parallel.h"

namespace pb200 {

namespace {
struct TaskAborted {};
struct NeedFallback {};
inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}
enum : int { T_IDLE = 0, T_RUNNING = 1, T_DONE = 2, T_DIRTY = 3 };
}  // namespace

struct ReplayTask {
    int first = 0, last = 0;                    // order[first, last): the task's initial regions in ascending start[0] order
    std::atomic<int> state{T_IDLE};
    std::atomic<bool> abort{false};
    MumPool mp;                                 // accepted MUMs in pop order
    RegionPool rp;                              // the task's regions (initial ones copied in, then the children)
    struct FRead { int g; int64_t a, b; };      // foreign reads, positions [a, b]
    std::vector<FRead> freads;
    std::unique_ptr<Aligner::CandCache> local;  // regions searched on demand (not predicted by the speculation)
    int64_t n_fread = 0, n_fwrite = 0, misses = 0, regions = 0;
    double t_wait = 0, t_search = 0;
    void reset() {
        mp.mums.clear(); mp.start.clear(); mp.fwd.clear();
        rp.coord.clear(); rp.slen.clear();
        freads.clear(); local.reset();
        n_fread = n_fwrite = misses = regions = 0;
    }
};

struct ReplayCtx {
    Aligner& A;
    const int n;
    int ntasks = 0;
    std::unique_ptr<ReplayTask[]> tasks;
    std::vector<int64_t> hlo, hhi;              // span of task k in genome g: [hlo[g * ntasks + k], hhi[..]] (both ends are anchor bits)
    std::vector<BitRow> S;
    std::vector<BitRow>& truth;
    std::mutex mu;
    std::condition_variable cv;
    int next_task = 0, running = 0, done_count = 0, max_started = -1, fallback_from = INT_MAX;
    bool paused = false;
    std::atomic<int> done_prefix{0};
    std::exception_ptr error;
    int64_t restarts = 0;
    unsigned jitter = 0;
    std::vector<int> order;                     // indices into A.initial_regions_, ascending start[0]

    ReplayCtx(Aligner& a) : A(a), n(a.n_), truth(a.truth_.layout) {}

    // ---- ownership of position x in genome g: owner task (or -1) and the extent [slo, shi] of that ownership segment
    inline void seg(int g, int64_t x, int& owner, int64_t& slo, int64_t& shi) const {
        const int64_t* L = &hlo[(size_t)g * ntasks];
        const int64_t* H = &hhi[(size_t)g * ntasks];
        const int k = (int)(std::upper_bound(L, L + ntasks, x) - L) - 1;
        if (k >= 0 && x <= H[k]) { owner = k; slo = L[k]; shi = H[k]; return; }
        owner = -1;
        slo = k >= 0 ? H[k] + 1 : 0;
        shi = k + 1 < ntasks ? L[k + 1] - 1 : A.len_[g];
    }
    void wait_done(ReplayTask& T, int m) {
        const double t0 = now_s();
        while (tasks[m].state.load(std::memory_order_acquire) != T_DONE) {
            if (T.abort.load(std::memory_order_relaxed)) throw TaskAborted();
            cpu_relax();
            std::this_thread::yield();
        }
        T.t_wait += now_s() - t0;
    }
    // the bitmap task k sees for a segment owned by `owner`; need = a lower task that has to finish first
    inline const BitRow* source(int k, int g, int owner, int& need) const {
        if (owner >= 0 && owner < k) {
            if (tasks[owner].state.load(std::memory_order_acquire) != T_DONE) { need = owner; return nullptr; }
            return &truth[g];
        }
        return owner == k ? &truth[g] : &S[g];
    }
    // a foreign read: body(need) runs under the mutex; when it meets a segment of an unfinished lower task it sets need and
    // returns; the read is retried after that task is done
    template <class Body>
    int64_t foreign_read(ReplayTask& T, int g, int64_t lo, int64_t hi, Body&& body) {
        for (;;) {
            int need = -1;
            {
                std::lock_guard<std::mutex> lk(mu);
                if (T.abort.load(std::memory_order_relaxed)) throw TaskAborted();
                const int64_t r = body(need);
                if (need < 0) {
                    T.freads.push_back(ReplayTask::FRead{g, lo, hi});
                    ++T.n_fread;
                    return r;
                }
            }
            wait_done(T, need);
        }
    }
    int64_t f_run_up(ReplayTask& T, int k, int g, int64_t a, int64_t b) {
        return foreign_read(T, g, a, b - 1, [&](int& need) -> int64_t {
            int64_t x = a, total = 0;
            while (x < b) {
                int owner; int64_t slo, shi;
                seg(g, x, owner, slo, shi);
                const BitRow* src = source(k, g, owner, need);
                if (!src) return 0;
                const int64_t e = std::min(b, shi + 1);
                const int64_t r = src->run_up(x, e);
                total += r;
                if (r < e - x) break;
                x = e;
            }
            return total;
        });
    }
    int64_t f_run_down(ReplayTask& T, int k, int g, int64_t a, int64_t b) {
        return foreign_read(T, g, a, b - 1, [&](int& need) -> int64_t {
            int64_t x = b, total = 0;
            while (x > a) {
                int owner; int64_t slo, shi;
                seg(g, x - 1, owner, slo, shi);
                const BitRow* src = source(k, g, owner, need);
                if (!src) return 0;
                const int64_t s = std::max(a, slo);
                const int64_t r = src->run_down(s, x);
                total += r;
                if (r < x - s) break;
                x = s;
            }
            return total;
        });
    }
    bool f_get(ReplayTask& T, int k, int g, int64_t i) {
        return foreign_read(T, g, i, i, [&](int& need) -> int64_t {
            int owner; int64_t slo, shi;
            seg(g, i, owner, slo, shi);
            const BitRow* src = source(k, g, owner, need);
            return src ? (int64_t)src->get(i) : 0;
        }) != 0;
    }
    int64_t f_prev_set(ReplayTask& T, int k, int g, int64_t i) {
        if (i < 0) return -1;
        return foreign_read(T, g, 0, i, [&](int& need) -> int64_t {          // (logged extent: conservative)
            int64_t x = i;
            while (x >= 0) {
                int owner; int64_t slo, shi;
                seg(g, x, owner, slo, shi);
                const BitRow* src = source(k, g, owner, need);
                if (!src) return -1;
                const int64_t r = src->prev_set_from(x, slo);
                if (r >= 0) return r;
                x = slo - 1;
            }
            return -1;
        });
    }
    int64_t f_next_set(ReplayTask& T, int k, int g, int64_t i, int64_t limit) {
        if (i >= limit) return limit;
        return foreign_read(T, g, i, limit - 1, [&](int& need) -> int64_t {
            int64_t x = i;
            while (x < limit) {
                int owner; int64_t slo, shi;
                seg(g, x, owner, slo, shi);
                const BitRow* src = source(k, g, owner, need);
                if (!src) return limit;
                const int64_t e = std::min(limit, shi + 1);
                const int64_t r = src->next_set(x, e);
                if (r < e) return r;
                x = e;
            }
            return limit;
        });
    }
    void foreign_commit(ReplayTask& T, int k, const int64_t* st, int64_t length, const int64_t* lo, const int64_t* hi);
    void restart_above(int k);
    void run_task(int k, std::vector<int>& found, std::vector<int>& children);
    void worker();
};

// a task's view of the layout
struct TaskAccess {
    ReplayCtx& X;
    ReplayTask& T;
    const int k;
    const int64_t* lo;                          // the task's span per genome
    const int64_t* hi;
    inline bool home(int g, int64_t a, int64_t b) const { return a >= lo[g] && b <= hi[g] + 1; }     // [a, b) inside the span
    inline bool get(int g, int64_t i) { return home(g, i, i + 1) ? X.truth[g].get(i) : X.f_get(T, k, g, i); }
    inline int64_t run_up(int g, int64_t a, int64_t b) {
        if (a >= b) return 0;
        return home(g, a, b) ? X.truth[g].run_up(a, b) : X.f_run_up(T, k, g, a, b);
    }
    inline int64_t run_down(int g, int64_t a, int64_t b) {
        if (a >= b) return 0;
        return home(g, a, b) ? X.truth[g].run_down(a, b) : X.f_run_down(T, k, g, a, b);
    }
    inline int64_t prev_set(int g, int64_t i) {
        if (i < 0) return -1;
        if (i >= lo[g] && i <= hi[g]) {
            const int64_t r = X.truth[g].prev_set_from(i, lo[g]);     // lo is an anchor bit (or position 0): the scan ends inside the span
            if (r >= 0 || lo[g] == 0) return r;
        }
        return X.f_prev_set(T, k, g, i);
    }
    inline int64_t next_set(int g, int64_t i, int64_t limit) {
        if (i >= limit) return limit;
        if (i >= lo[g] && i <= hi[g]) {
            const int64_t e = std::min(limit, hi[g] + 1);               // hi is an anchor bit (or the sentinel)
            const int64_t r = X.truth[g].next_set(i, e);
            if (r < e || e == limit) return r;
        }
        return X.f_next_set(T, k, g, i, limit);
    }
    inline void commit(const int64_t* st, int64_t length, int n) {
        bool foreign = false;
        for (int g = 0; g < n; ++g) foreign |= !home(g, st[g], st[g] + length);
        if (!foreign) { for (int g = 0; g < n; ++g) X.truth[g].set_range_atomic(st[g], st[g] + length); return; }
        X.foreign_commit(T, k, st, length, lo, hi);
    }
};

void ReplayCtx::foreign_commit(ReplayTask& T, int k, const int64_t* st, int64_t length, const int64_t* lo, const int64_t* hi) {
    // only the lowest unfinished task writes outside its span: everything below it is final, so is its own state
    {
        const double t0 = now_s();
        while (done_prefix.load(std::memory_order_acquire) < k) {
            if (T.abort.load(std::memory_order_relaxed)) throw TaskAborted();
            cpu_relax();
            std::this_thread::yield();
        }
        T.t_wait += now_s() - t0;
    }
    bool conflict = false;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (T.abort.load(std::memory_order_relaxed)) throw TaskAborted();
        for (int g = 0; g < n; ++g) {
            const int64_t a = st[g], b = st[g] + length;
            truth[g].set_range_atomic(a, b);
            if (a >= lo[g] && b <= hi[g] + 1) continue;
            S[g].set_range(a, b);
            ++T.n_fwrite;
            for (int m = k + 1; m <= max_started && !conflict; ++m) {
                if (tasks[m].state.load(std::memory_order_relaxed) == T_IDLE) continue;
                if (a <= hhi[(size_t)g * ntasks + m] && b - 1 >= hlo[(size_t)g * ntasks + m]) { conflict = true; break; }
                for (const ReplayTask::FRead& fr : tasks[m].freads)
                    if (fr.g == g && fr.a <= b - 1 && fr.b >= a) { conflict = true; break; }
            }
        }
        if (conflict) {
            paused = true;
            for (int m = k + 1; m <= max_started; ++m) tasks[m].abort.store(true, std::memory_order_relaxed);
        }
    }
    if (conflict) restart_above(k);
}

// every task above k that has started is stale: wait until they have stopped, give their spans back their pre-task state and
// hand them out again
void ReplayCtx::restart_above(int k) {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return running == 1; });                  // (task k itself)
    for (int m = k + 1; m <= max_started; ++m) {
        ReplayTask& M = tasks[m];
        const int st = M.state.load(std::memory_order_relaxed);
        if (st != T_IDLE) {
            for (int g = 0; g < n; ++g) truth[g].copy_range_from(S[g], hlo[(size_t)g * ntasks + m], hhi[(size_t)g * ntasks + m] + 1);
            if (st == T_DONE) --done_count;
            M.reset();
            M.state.store(T_IDLE, std::memory_order_release);
        }
        M.abort.store(false, std::memory_order_relaxed);
    }
    next_task = k + 1;
    max_started = k;
    paused = false;
    ++restarts;
    cv.notify_all();
}

namespace {
struct QE { int64_t s0; int id; int slice; uint64_t hash; };
}

// the loop of Aligner::process_queue_exact restricted to one task (its queue is a contiguous piece of the reference's queue:
// every other region has a smaller key and is finished, or a larger key and is untouched)
void ReplayCtx::run_task(int k, std::vector<int>& found, std::vector<int>& children) {
    ReplayTask& T = tasks[k];
    T.reset();
    T.rp.n = n;
    std::vector<int64_t> lo((size_t)n), hi((size_t)n);
    for (int g = 0; g < n; ++g) { lo[g] = hlo[(size_t)g * ntasks + k]; hi[g] = hhi[(size_t)g * ntasks + k]; }
    TaskAccess acc{*this, T, k, lo.data(), hi.data()};
    const size_t cbytes = sizeof(int64_t) * 2 * (size_t)n;
    std::vector<QE> fast;                       // descending start[0]: the back is the front of the reference's vector
    fast.reserve((size_t)(T.last - T.first) + 16);
    for (int p = T.last - 1; p >= T.first; --p) {
        const int i = order[(size_t)p];
        const int r = A.initial_regions_[(size_t)i];
        const int id = T.rp.add(A.rp_.start(r), A.rp_.end(r));
        fast.push_back(QE{T.rp.start(id)[0], id, (size_t)i < A.slice_of_initial_.size() ? A.slice_of_initial_[(size_t)i] : -1,
                          Aligner::coords_hash(T.rp.start(id), 2 * n)});
    }
    if (k == 0) {
        // the reference takes regions.begin() BEFORE its first sort (src/parsnp.cpp:192-195): the very first region searched is
        // the first one pushed, whatever its key (the right side of the first anchor when its left side was too short, while
        // the second anchor's left side starts one base earlier)
        for (size_t q = 0; q < fast.size(); ++q)
            if (order[(size_t)T.first + (fast.size() - 1 - q)] == 0) { std::rotate(fast.begin() + (long)q, fast.begin() + (long)q + 1, fast.end()); break; }
    }
    auto fast_pos = [&](int64_t key) {          // index of the first element (from the front) with s0 <= key
        size_t pos = fast.size();
        while (pos > 0 && fast[pos - 1].s0 < key) --pos;
        if (pos > 0 && fast[pos - 1].s0 == key) return pos - 1;
        return pos;
    };
    auto req = [&](int a, int b) { return std::memcmp(T.rp.start(a), T.rp.start(b), cbytes) == 0; };
    std::vector<int64_t> lS((size_t)n), lE((size_t)n), rS((size_t)n), rE((size_t)n);
    int ready_upto = 0;
    unsigned rnd = 12345u + (unsigned)k * 2654435761u;
    while (!fast.empty()) {
        if (T.abort.load(std::memory_order_relaxed)) throw TaskAborted();
        if (jitter) {                           // tests: shake the interleavings
            rnd = rnd * 1664525u + 1013904223u;
            if ((rnd >> 16) % jitter == 0) std::this_thread::sleep_for(std::chrono::microseconds((rnd >> 8) % 200));
            else if ((rnd >> 20) % 3 == 0) std::this_thread::yield();
        }
        const QE cur = fast.back();
        fast.pop_back();
        ++T.regions;
        const Aligner::CandCache* C = nullptr;
        if (cur.slice >= 0 && cur.slice < (int)A.slice_cache_.size()) {
            if (cur.slice >= ready_upto) {
                const double t0 = now_s();
                A.wait_slice_quiet(cur.slice);
                T.t_wait += now_s() - t0;
                ready_upto = cur.slice + 1;
            }
            C = A.slice_cache_[(size_t)cur.slice].get();
        }
        int ci = C ? C->lookup(T.rp.start(cur.id), cur.hash) : -1;
        if (ci < 0) { C = &A.main_cache_; ci = A.main_cache_.lookup(T.rp.start(cur.id), cur.hash); }
        if (ci < 0 && T.local) { C = T.local.get(); ci = C->lookup(T.rp.start(cur.id), cur.hash); }
        if (ci < 0) {                           // a region the speculation did not predict: search it now
            const double t0 = now_s();
            if (!T.local) T.local.reset(new Aligner::CandCache);
            A.search_regions(*T.local, T.rp, std::vector<int>(1, cur.id), false);
            T.t_search += now_s() - t0;
            ++T.misses;
            C = T.local.get();
            ci = C->lookup(T.rp.start(cur.id), cur.hash);
        }
        found.clear();
        A.accept_candidates_t(T.rp.start(cur.id), T.rp.end(cur.id), T.rp.slen[(size_t)cur.id], *C, ci, acc, T.mp, found, false);
        children.clear();
        int64_t lsl = 0;
        for (size_t i = 0; i < found.size(); ++i) {
            const MumRec& m = T.mp.mums[(size_t)found[i]];
            const int64_t* ms = &T.mp.start[(size_t)m.off];
            if (i == 0) lsl = A.det_region_t(acc, ms, m.length, true, lS.data(), lE.data());
            const int64_t rsl = A.det_region_t(acc, ms, m.length, false, rS.data(), rE.data());
            if (lsl > A.prm_.q) children.push_back(T.rp.add(lS.data(), lE.data()));
            if (rsl > A.prm_.q) children.push_back(T.rp.add(rS.data(), rE.data()));
            if (i + 1 < found.size()) {
                const MumRec& m2 = T.mp.mums[(size_t)found[i + 1]];
                lsl = A.det_region_t(acc, &T.mp.start[(size_t)m2.off], m2.length, true, lS.data(), lE.data());
            }
        }
        // sort + drop adjacent duplicates (src/parsnp.cpp:291-306): with distinct keys an ordered insert; a tie between
        // DIFFERENT regions makes the order a property of std::sort over the whole queue -> sequential loop from this task on
        for (size_t a = 0; a < children.size(); ++a) {
            const int64_t key = T.rp.start(children[a])[0];
            const size_t pos = fast_pos(key);
            if (pos < fast.size() && fast[pos].s0 == key && !req(fast[pos].id, children[a])) {
                if (getenv("PB200_REPLAY_DEBUG")) fprintf(stderr, "[pb200 replay] task %d: child ties with a queued region at start[0] = %lld\n", k, (long long)key);
                throw NeedFallback();
            }
            for (size_t b = 0; b < a; ++b)
                if (key == T.rp.start(children[b])[0] && !req(children[a], children[b])) {
                    if (getenv("PB200_REPLAY_DEBUG")) fprintf(stderr, "[pb200 replay] task %d: two children tie at start[0] = %lld\n", k, (long long)key);
                    throw NeedFallback();
                }
        }
        for (int ch : children) {                               // identical duplicates collapse (the first one stays)
            const int64_t key = T.rp.start(ch)[0];
            const size_t pos = fast_pos(key);
            if (pos < fast.size() && fast[pos].s0 == key) continue;
            fast.insert(fast.begin() + (long)pos, QE{key, ch, cur.slice, Aligner::coords_hash(T.rp.start(ch), 2 * n)});
        }
    }
}

void ReplayCtx::worker() {
    std::vector<int> found, children;
    for (;;) {
        int k = -1;
        {
            std::unique_lock<std::mutex> lk(mu);
            for (;;) {
                if (error) return;
                const int limit = std::min(ntasks, fallback_from);
                if (!paused && next_task < limit) break;
                if (!paused && running == 0) return;            // nothing to hand out, nobody who could restart anything
                cv.wait(lk);
            }
            k = next_task++;
            tasks[k].state.store(T_RUNNING, std::memory_order_release);
            tasks[k].abort.store(false, std::memory_order_relaxed);
            ++running;
            max_started = std::max(max_started, k);
        }
        int outcome = T_DONE;
        std::exception_ptr err;
        try { run_task(k, found, children); }
        catch (const TaskAborted&) { outcome = T_DIRTY; }
        catch (const NeedFallback&) { outcome = -1; }
        catch (...) { err = std::current_exception(); outcome = T_DIRTY; }
        {
            std::lock_guard<std::mutex> lk(mu);
            --running;
            if (err && !error) {
                error = err;
                for (int m = 0; m < ntasks; ++m) tasks[m].abort.store(true, std::memory_order_relaxed);
            }
            if (outcome == -1) {
                // the parallel part ends below this task; tasks above it are pointless now
                fallback_from = std::min(fallback_from, k);
                for (int m = k + 1; m <= max_started; ++m) tasks[m].abort.store(true, std::memory_order_relaxed);
                outcome = T_DIRTY;
            }
            tasks[k].state.store(outcome, std::memory_order_release);
            if (outcome == T_DONE) {
                ++done_count;
                int p = done_prefix.load(std::memory_order_relaxed);
                while (p < ntasks && tasks[p].state.load(std::memory_order_relaxed) == T_DONE) ++p;
                done_prefix.store(p, std::memory_order_release);
            }
            cv.notify_all();
        }
    }
}

// (the replay tasks wait for speculation slices without touching the shared statistics)
void Aligner::wait_slice_quiet(int slice) {
    if (slice < 0 || slice >= (int)slice_cache_.size()) return;
    std::unique_lock<std::mutex> lk(slice_mu_);
    slice_cv_.wait(lk, [&] { return slices_ready_ > slice; });
}

bool Aligner::do_work_parallel() {
    const int W = replay_threads_ > 0 ? replay_threads_ : threads_;
    const size_t R = initial_regions_.size();
    const char* et = getenv("PB200_REPLAY_TASK");               // initial regions per task (tests: 1 = one gap per task)
    const char* em = getenv("PB200_REPLAY_MODE");               // "seq" = never, "par" = also for tiny inputs / one worker
    const bool force = em && std::strcmp(em, "par") == 0;
    if (em && std::strcmp(em, "seq") == 0) return false;
    const bool dbg = getenv("PB200_REPLAY_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[pb200 replay] W %d R %zu trace %d pipeline %d\n", W, R, (int)trace_on_, (int)pipeline_);
    if (trace_on_ || !pipeline_ || R == 0) return false;
    if (!force && (W < 2 || R < 512)) return false;
    ReplayCtx X(*this);
    if (const char* ej = getenv("PB200_REPLAY_JITTER")) X.jitter = (unsigned)std::max(0, atoi(ej));
    // P1: the keys of the initial regions are distinct, so the reference's first sort has one possible outcome.  (Push order is
    // NOT ascending: the right side of anchor i starts one base behind the left side of anchor i+1 - the same gap twice.)
    X.order.resize(R);
    for (size_t i = 0; i < R; ++i) X.order[i] = (int)i;
    std::stable_sort(X.order.begin(), X.order.end(), [&](int a, int b) { return rstart(initial_regions_[(size_t)a])[0] < rstart(initial_regions_[(size_t)b])[0]; });
    for (size_t i = 1; i < R; ++i)
        if (!(rstart(initial_regions_[(size_t)X.order[i - 1]])[0] < rstart(initial_regions_[(size_t)X.order[i]])[0])) {
            if (dbg) fprintf(stderr, "[pb200 replay] two initial regions tie on start[0]: sequential\n");
            return false;
        }
    size_t pos0 = 0;
    while (pos0 < R && X.order[pos0] != 0) ++pos0;
    // spans of the initial regions: between the anchor bits that bound them, per genome (the layout holds exactly the anchors)
    std::vector<int64_t> rlo(R * (size_t)n_), rhi(R * (size_t)n_);
    std::atomic<int> bad(0);
    {
        const long per_blk = 512;
        parallel_chunks(threads_, ((long)R + per_blk - 1) / per_blk, [&](long c) {
            for (size_t p = (size_t)c * per_blk; p < std::min(R, (size_t)(c + 1) * per_blk); ++p) {
                const int r = initial_regions_[(size_t)X.order[p]];
                for (int g = 0; g < n_; ++g) {
                    const int64_t s = rstart(r)[g], e = rend(r)[g];
                    if (s < 0 || e > len_[g] || e < s) { bad.store(1); continue; }
                    int64_t a = truth_.layout[g].prev_set(s > 0 ? s - 1 : 0);
                    if (a < 0) a = 0;
                    const int64_t b = truth_.layout[g].next_set(e, len_[g] + 1);   // (the sentinel bit at len ends every scan)
                    rlo[p * (size_t)n_ + g] = a;
                    rhi[p * (size_t)n_ + g] = std::min(b, len_[g]);
                }
            }
        });
    }
    if (bad.load()) return false;
    // tasks: runs of >= `per` regions of the sorted list, cut only where the next region starts behind everything before it in
    // every genome (the two regions of one gap, and whatever else overlaps, stay together)
    const size_t per = et ? (size_t)std::max(1, atoi(et)) : std::min<size_t>(128, std::max<size_t>(8, R / ((size_t)W * 24) + 1));
    std::vector<int> cuts(1, 0);
    {
        std::vector<int64_t> runmax((size_t)n_, -1);
        size_t count = 0;
        for (size_t p = 0; p < R; ++p) {
            if (count >= per && p > pos0) {
                bool ok = true;
                for (int g = 0; g < n_ && ok; ++g) ok = runmax[(size_t)g] <= rlo[p * (size_t)n_ + g];
                if (ok) { cuts.push_back((int)p); count = 0; }
            }
            for (int g = 0; g < n_; ++g) runmax[(size_t)g] = std::max(runmax[(size_t)g], rhi[p * (size_t)n_ + g]);
            ++count;
        }
        cuts.push_back((int)R);
    }
    X.ntasks = (int)cuts.size() - 1;
    if (!force && X.ntasks < 2 * W) {
        if (dbg) fprintf(stderr, "[pb200 replay] only %d independent tasks for %d workers: sequential\n", X.ntasks, W);
        return false;
    }
    X.tasks.reset(new ReplayTask[(size_t)X.ntasks]);
    X.hlo.assign((size_t)n_ * X.ntasks, 0);
    X.hhi.assign((size_t)n_ * X.ntasks, 0);
    for (int k = 0; k < X.ntasks; ++k) {
        ReplayTask& T = X.tasks[k];
        T.first = cuts[(size_t)k];
        T.last = cuts[(size_t)k + 1];
        for (int g = 0; g < n_; ++g) {
            int64_t lo = INT64_MAX, hi = -1;
            for (int p = T.first; p < T.last; ++p) {
                lo = std::min(lo, rlo[(size_t)p * n_ + g]);
                hi = std::max(hi, rhi[(size_t)p * n_ + g]);
            }
            X.hlo[(size_t)g * X.ntasks + k] = lo;
            X.hhi[(size_t)g * X.ntasks + k] = hi;
        }
    }
    // P2: collinear anchors - the spans ascend with the task in every genome (neighbours may share their boundary anchor bit)
    for (int g = 0; g < n_; ++g)
        for (int k = 0; k + 1 < X.ntasks; ++k)
            if (X.hhi[(size_t)g * X.ntasks + k] > X.hlo[(size_t)g * X.ntasks + k + 1] ||
                X.hlo[(size_t)g * X.ntasks + k] > X.hhi[(size_t)g * X.ntasks + k]) {
                if (dbg) fprintf(stderr, "[pb200 replay] spans of tasks %d and %d not in order in genome %d: sequential\n", k, k + 1, g);
                return false;
            }
    X.S.resize((size_t)n_);
    parallel_chunks(threads_, (long)n_, [&](long g) { X.S[(size_t)g] = truth_.layout[(size_t)g]; });
    const size_t anchors_in_pool = mp_.mums.size();

    const int nw = std::max(1, std::min(W, X.ntasks));
    std::vector<std::thread> th;
    for (int w = 1; w < nw; ++w) th.emplace_back([&X] { X.worker(); });
    X.worker();
    for (auto& t : th) t.join();
    if (X.error) std::rethrow_exception(X.error);

    // ---- the finished tasks' MUMs into the pools, task after task = the reference's push order
    const double tm0 = now_s();
    const int F = std::min(X.ntasks, X.fallback_from);
    std::vector<size_t> base((size_t)F + 1, 0);
    for (int k = 0; k < F; ++k) base[(size_t)k + 1] = base[(size_t)k] + X.tasks[k].mp.mums.size();
    const size_t M = base[(size_t)F], m0 = mp_.mums.size(), s0 = mp_.start.size(), a0 = all_mums_.size();
    mp_.mums.resize(m0 + M);
    mp_.start.resize(s0 + M * (size_t)n_);
    mp_.fwd.resize(s0 + M * (size_t)n_);
    all_mums_.resize(a0 + M);
    parallel_chunks(threads_, (long)F, [&](long k) {
        const ReplayTask& T = X.tasks[k];
        const size_t b = base[(size_t)k], cnt = T.mp.mums.size();
        if (!cnt) return;
        std::memcpy(&mp_.start[s0 + b * (size_t)n_], T.mp.start.data(), cnt * (size_t)n_ * sizeof(int64_t));
        std::memcpy(&mp_.fwd[s0 + b * (size_t)n_], T.mp.fwd.data(), cnt * (size_t)n_);
        for (size_t i = 0; i < cnt; ++i) {
            MumRec m = T.mp.mums[i];
            m.off = (int64_t)(s0 + (b + i) * (size_t)n_);
            mp_.mums[m0 + b + i] = m;
            all_mums_[a0 + b + i] = (int)(m0 + b + i);
        }
    });
    for (int k = 0; k < F; ++k) {
        const ReplayTask& T = X.tasks[k];
        stats_.replay_foreign_reads += T.n_fread;
        stats_.replay_foreign_writes += T.n_fwrite;
        stats_.replay_misses += T.misses;
        stats_.t_replay_wait += T.t_wait / nw;
        stats_.t_replay_search += T.t_search;
    }
    stats_.replay_tasks = F;
    stats_.replay_restarts = X.restarts;
    stats_.replay_workers = nw;
    stats_.t_replay_merge += now_s() - tm0;
    if (F < X.ntasks) {
        // a tie inside task F: rebuild the layout as the reference has it when it reaches that task (anchors + the MUMs of
        // everything before), then its own loop from there
        stats_.replay_fallback = 1;
        (void)anchors_in_pool;
        parallel_chunks(threads_, (long)n_, [&](long g) {
            BitRow& row = truth_.layout[(size_t)g];
            row.init(len_[(size_t)g] + 1);
            row.set_range(len_[(size_t)g], len_[(size_t)g] + 1);
            for (const MumRec& m : mp_.mums) {
                const int64_t s = mp_.start[(size_t)m.off + (size_t)g];
                row.set_range(s, s + m.length);
            }
        });
        // (a tie in the very first task: everything as pushed, the reference's first pop precedes its first sort)
        std::vector<int> rest, rest_slice;
        for (size_t p = F == 0 ? 0 : (size_t)X.tasks[F].first; p < R; ++p) {
            const size_t i = F == 0 ? p : (size_t)X.order[p];
            rest.push_back(initial_regions_[i]);
            rest_slice.push_back(i < slice_of_initial_.size() ? slice_of_initial_[i] : -1);
        }
        std::vector<int> out;
        process_queue_exact(rest, rp_, truth_.layout, mp_, out, &rest_slice);
        all_mums_.insert(all_mums_.end(), out.begin(), out.end());
    }
    return true;
}

}  // namespace pb200
