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
//
// FINAL GAPS.  When the engine followed the recursion itself (cuda/recursion.cuh) it has already run loop D + determineRegion
// for every region it discovered, on a scratch layout and level by level instead of in the reference's order.  Inside one gap
// that difference cannot matter when (a) no candidate that reached the trim loop had a reverse-strand genome (everything read
// and written lies inside the region) - or every such candidate dies on the anchors alone (foreign_harmless below) -, (b) the accepted MUMs of every region ascend in every genome (sibling sub-regions are
// disjoint: they commute), (c) the two overlapping regions of a gap pair were searched in the reference's order by one CTA and
// the second one accepted nothing (more bits only trim more: it accepts nothing in the reference's order either), and (d) no
// MUM of another gap was written into the gap's span - neither on the device (its list of writes outside their region) nor
// here (every accepted MUM with a reverse-strand genome marks the gaps it touches).  For such a gap the task appends the
// engine's accepted MUMs - regions in ascending start[0] order, candidates in order: the reference's pop order - and sets
// their bits, without lookup, trim, determineRegion or queue.  Everything else (the foreign access rules above, restarts) is
// unchanged: a final gap is just a gap whose replay is cheap.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
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
    int gfirst = 0, glast = 0;                  // its gaps (groups of initial regions with overlapping spans): [gfirst, glast)
    std::atomic<int> state{T_IDLE};
    std::atomic<bool> abort{false};
    int worker = -1;                            // the worker whose arenas hold the task's output ...
    size_t mbegin = 0, mend = 0;                // ... its accepted MUMs in pop order: wmp[worker].mums[mbegin, mend)
    // foreign reads, positions [a, b]: appended by the task's own thread (entry first, then the count with release order), read
    // by a lower task's thread that checks whether its foreign write came too late for this task
    struct FRead { int g; int64_t a, b; };
    static constexpr uint32_t LOG_CAP = 2048;
    FRead* flog = nullptr;                      // (LOG_CAP entries of the context's slab)
    std::atomic<uint32_t> nlog{0};
    std::atomic<bool> log_overflow{false};      // more reads than the log holds: "has read everything"
    std::unique_ptr<Aligner::CandCache> local;  // regions searched on demand (not predicted by the speculation)
    int64_t n_fread = 0, n_fwrite = 0, misses = 0, regions = 0, slow_iters = 0, final_gaps = 0, final_mums = 0;
    double t_wait = 0, t_search = 0;
    struct SegCache { int g = -1, owner = -1, gap = -1; int64_t slo = 0, shi = -1; } sc;   // last ownership lookup (run_up and run_down of one candidate hit the same segment)
    uint64_t pc[7] = {0, 0, 0, 0, 0, 0, 0};       // PB200_PROFILE_HOST: cycles in setup / pop / lookup / accept / det_region / queue / final gaps
    void reset() {
        worker = -1; mbegin = mend = 0;
        nlog.store(0, std::memory_order_relaxed); log_overflow.store(false, std::memory_order_relaxed);
        local.reset();
        n_fread = n_fwrite = misses = regions = slow_iters = final_gaps = final_mums = 0;
        for (auto& x : pc) x = 0;
        sc = SegCache();
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
    int nw = 1;                                 // workers
    double t_setup = 0;
    std::vector<int> order;                     // indices into A.initial_regions_, ascending start[0]
    std::vector<int64_t> keys;                  // their start[0], same order
    // GAPS: maximal runs of initial regions whose spans overlap (the right side of anchor i and the left side of anchor i+1 are
    // the same gap).  Gap j owns [glo[j * n + g], ghi[j * n + g]] in genome g; gaps ascend in every genome.  A gap is DONE when
    // its task has popped a region that starts behind it: nothing of its subtree is left in the queue.
    int ngaps = 0;
    std::vector<int32_t> glo, ghi;
    std::vector<int64_t> gend0;                 // largest end[0] of the gap's regions
    std::unique_ptr<std::atomic<uint8_t>[]> gdone;
    std::vector<int> gcut;                      // gap j = positions [gcut[j], gcut[j + 1]) of `order`
    // final gaps (see the head of the file): candidates after classify_gaps(), their regions [gr0, gr1) in the engine's sorted
    // list; gtouched[j] = a MUM with a reverse-strand genome was (or, by the engine's prediction, would be) written into gap j
    bool any_final = false;
    std::vector<uint8_t> gfinal;
    std::vector<int32_t> gr0, gr1;
    std::unique_ptr<std::atomic<uint8_t>[]> gtouched;
    void classify_gaps();
    bool foreign_harmless(size_t r) const;
    void mark_touched(int g, int64_t a, int64_t b);
    void apply_final_gap(ReplayTask& T, MumPool& MP, int j, const int64_t* lo, const int64_t* hi);
    std::unique_ptr<ReplayTask::FRead[]> log_slab;
    // per worker: an append-only arena for the MUMs of the tasks it ran (a restarted task leaves its first attempt behind as
    // unreferenced garbage) and a region pool that is cleared for every task
    std::vector<MumPool> wmp;
    std::vector<RegionPool> wrp;

    ReplayCtx(Aligner& a) : A(a), n(a.n_), truth(a.truth_.layout) {}

    // ---- ownership of position x in genome g: the task whose span holds it (or -1), the gap inside that task (or -1: between
    // gaps, i.e. anchor bits and stretches too short to be searched - never written by a task's own accepts) and the extent
    // [slo, shi] of that ownership segment
    inline void seg_lookup(int g, int64_t x, int& owner, int& gap, int64_t& slo, int64_t& shi) const {
        const int64_t* L = &hlo[(size_t)g * ntasks];
        const int64_t* H = &hhi[(size_t)g * ntasks];
        const int k = (int)(std::upper_bound(L, L + ntasks, x) - L) - 1;
        gap = -1;
        if (!(k >= 0 && x <= H[k])) {
            owner = -1;
            slo = k >= 0 ? H[k] + 1 : 0;
            shi = k + 1 < ntasks ? L[k + 1] - 1 : A.len_[g];
            return;
        }
        owner = k;
        // last gap j of the task with glo[j][g] <= x
        int a = tasks[k].gfirst, b = tasks[k].glast;             // invariant: glo[a-1] <= x < glo[b]
        const int g0 = a, g1 = b;
        while (a < b) {
            const int mid = (a + b) >> 1;
            if ((int64_t)glo[(size_t)mid * n + g] <= x) a = mid + 1; else b = mid;
        }
        const int j = a - 1;
        if (j >= g0 && x <= (int64_t)ghi[(size_t)j * n + g]) { gap = j; slo = glo[(size_t)j * n + g]; shi = ghi[(size_t)j * n + g]; return; }
        slo = j >= g0 ? (int64_t)ghi[(size_t)j * n + g] + 1 : L[k];
        shi = j + 1 < g1 ? (int64_t)glo[(size_t)(j + 1) * n + g] - 1 : H[k];
    }
    inline void seg(ReplayTask& T, int g, int64_t x, int& owner, int& gap, int64_t& slo, int64_t& shi) const {
        ReplayTask::SegCache& c = T.sc;
        if (c.g != g || x < c.slo || x > c.shi) { c.g = g; seg_lookup(g, x, c.owner, c.gap, c.slo, c.shi); }
        owner = c.owner; gap = c.gap; slo = c.slo; shi = c.shi;
    }
    void wait_gap(ReplayTask& T, int j) {
        const double t0 = now_s();
        while (!gdone[j].load(std::memory_order_acquire)) {
            if (T.abort.load(std::memory_order_relaxed)) throw TaskAborted();
            cpu_relax();
            std::this_thread::yield();
        }
        T.t_wait += now_s() - t0;
    }
    // the bitmap task k sees for a segment; need = a gap of a lower task that has to be finished first
    inline const BitRow* source(int k, int g, int owner, int gap, int& need) const {
        if (owner < 0 || gap < 0 || owner > k) return &S[g];      // untouched by the reference at this point (or never written)
        if (owner == k) return &truth[g];
        if (!gdone[gap].load(std::memory_order_acquire)) { need = gap; return nullptr; }
        return &truth[g];
    }
    // A foreign read takes no lock.  It is logged BEFORE the bits are read (entry, count, full fence, read); a foreign write sets
    // its bits, issues a full fence and THEN looks at the logs: whichever way the two interleave, either the reader sees the
    // new bits or the writer sees the log entry (and restarts the reader's task).  body(need) sets need when it meets an
    // unfinished gap of a lower task: the read is retried after that gap is done.
    template <class Body>
    int64_t foreign_read(ReplayTask& T, int g, int64_t lo, int64_t hi, Body&& body) {
        const uint32_t at = T.nlog.load(std::memory_order_relaxed);
        if (at < ReplayTask::LOG_CAP) {
            T.flog[at] = ReplayTask::FRead{g, lo, hi};
            T.nlog.store(at + 1, std::memory_order_release);
        } else {
            T.log_overflow.store(true, std::memory_order_release);
        }
        ++T.n_fread;
        for (;;) {
            if (T.abort.load(std::memory_order_relaxed)) throw TaskAborted();
            std::atomic_thread_fence(std::memory_order_seq_cst);
            int need = -1;
            const int64_t r = body(need);
            if (need < 0) return r;
            wait_gap(T, need);
        }
    }
    int64_t f_run_up(ReplayTask& T, int k, int g, int64_t a, int64_t b) {
        return foreign_read(T, g, a, b - 1, [&](int& need) -> int64_t {
            int64_t x = a, total = 0;
            while (x < b) {
                int owner, gap; int64_t slo, shi;
                seg(T, g, x, owner, gap, slo, shi);
                const BitRow* src = source(k, g, owner, gap, need);
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
                int owner, gap; int64_t slo, shi;
                seg(T, g, x - 1, owner, gap, slo, shi);
                const BitRow* src = source(k, g, owner, gap, need);
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
            int owner, gap; int64_t slo, shi;
            seg(T, g, i, owner, gap, slo, shi);
            const BitRow* src = source(k, g, owner, gap, need);
            return src ? (int64_t)src->get(i) : 0;
        }) != 0;
    }
    int64_t f_prev_set(ReplayTask& T, int k, int g, int64_t i) {
        if (i < 0) return -1;
        return foreign_read(T, g, 0, i, [&](int& need) -> int64_t {          // (logged extent: conservative)
            int64_t x = i;
            while (x >= 0) {
                int owner, gap; int64_t slo, shi;
                seg(T, g, x, owner, gap, slo, shi);
                const BitRow* src = source(k, g, owner, gap, need);
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
                int owner, gap; int64_t slo, shi;
                seg(T, g, x, owner, gap, slo, shi);
                const BitRow* src = source(k, g, owner, gap, need);
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
    void run_task(int k, int w, std::vector<int>& found, std::vector<int>& children);
    void worker(int w);
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
    inline void commit(const int64_t* st, int64_t length, int n, const uint8_t* fw) {
        if (X.any_final)                          // (before the bits: a gap that is about to be taken as final must see the mark)
            for (int g = 0; g < n; ++g) if (!fw[g]) X.mark_touched(g, st[g], st[g] + length);
        bool foreign = false;
        for (int g = 0; g < n; ++g) foreign |= !home(g, st[g], st[g] + length);
        if (!foreign) {
            // words strictly inside the span have no other writer (a lower task's foreign write there restarts this task and
            // restores the span): plain stores; the two boundary words may be shared with the neighbouring spans
            for (int g = 0; g < n; ++g) X.truth[g].set_range_owned(st[g], st[g] + length, lo[g] >> 6, hi[g] >> 6);
            return;
        }
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
            S[g].set_range_atomic(a, b);
            ++T.n_fwrite;
        }
        std::atomic_thread_fence(std::memory_order_seq_cst);       // bits before logs (see foreign_read)
        for (int g = 0; g < n && !conflict; ++g) {
            const int64_t a = st[g], b = st[g] + length;
            if (a >= lo[g] && b <= hi[g] + 1) continue;
            for (int m = k + 1; m <= max_started && !conflict; ++m) {
                ReplayTask& M = tasks[m];
                if (M.state.load(std::memory_order_relaxed) == T_IDLE) continue;
                if (a <= hhi[(size_t)g * ntasks + m] && b - 1 >= hlo[(size_t)g * ntasks + m]) { conflict = true; break; }
                if (M.log_overflow.load(std::memory_order_acquire)) { conflict = true; break; }
                const uint32_t cnt = M.nlog.load(std::memory_order_acquire);
                for (uint32_t i = 0; i < cnt; ++i)
                    if (M.flog[i].g == g && M.flog[i].a <= b - 1 && M.flog[i].b >= a) { conflict = true; break; }
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
            for (int j = M.gfirst; j < M.glast; ++j) gdone[j].store(0, std::memory_order_relaxed);
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

// the gaps that [a, b) of genome g touches (gaps ascend in every genome; the bounding anchor bits count: conservative)
void ReplayCtx::mark_touched(int g, int64_t a, int64_t b) {
    if (a >= b || ngaps == 0) return;
    int lo = 0, hi = ngaps;                     // first gap j with ghi[j][g] >= a
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)ghi[(size_t)mid * n + g] < a) lo = mid + 1; else hi = mid;
    }
    for (int j = lo; j < ngaps && (int64_t)glo[(size_t)j * n + g] <= b - 1; ++j) gtouched[j].store(1, std::memory_order_relaxed);
}

// A region the engine flagged because a candidate with a reverse-strand genome reached its trim loop: harmless if every such
// candidate trims to nothing already on the layout that holds ONLY the anchors (= `truth` before the first task runs).  Bits
// are only ever added during the recursion, and the interval that survives the trim loop on a layout with more bits lies
// inside the one that survives with fewer (start and end only move inwards, genome after genome): such a candidate is
// rejected on the engine's scratch layout and on the reference's layout alike, whatever they hold at that moment - it read
// foreign positions, but nothing depends on what it saw, and it wrote nothing.
bool ReplayCtx::foreign_harmless(size_t r) const {
    const Aligner::DeviceDecisions& D = A.dev_;
    if (!D.lon || !D.fwd) return false;
    const int nq = n - 1;
    const WindowRec& w = D.wins[r];
    const int64_t* rs = D.coords + r * 2 * (size_t)n;
    const int64_t* re = rs + n;
    int64_t st_buf[64];
    std::vector<int64_t> st_vec;
    int64_t* st = st_buf;
    if (n > 64) { st_vec.resize((size_t)n); st = st_vec.data(); }
    for (int32_t c = 0; c < w.ncand; ++c) {
        const size_t ci = (size_t)w.cand_off + (size_t)c;
        const uint8_t* fwj = D.fwd + ci * (size_t)nq;
        bool rev = false;
        for (int j = 0; j < nq; ++j) rev |= !fwj[j];
        if (!rev) continue;
        // the checks that precede the trim loop (accept_impl.h; src/parsnp.cpp:1723, src/TMum.cpp:13-72)
        const int64_t LON = D.lon[ci];
        const uint64_t dsp0 = (uint64_t)((int64_t)D.k[ci] + 1 + w.ref_start);
        bool bad = (uint64_t)(dsp0 - (uint64_t)rs[0]) > (uint64_t)(uint32_t)(re[0] - rs[0]);
        st[0] = (int64_t)dsp0 - 1;
        bool any_fail = st[0] + LON > A.len_[0] || st[0] < 0;
        const int32_t* spj = D.sp + ci * (size_t)nq;
        for (int j = 1; j < n; ++j) {
            const uint64_t dsp = (uint64_t)((int64_t)spj[j - 1] + 1 + rs[j]);
            bad |= (uint64_t)(dsp - (uint64_t)rs[j]) > (uint64_t)(uint32_t)(re[j] - rs[j]);
            int64_t s2 = (int64_t)dsp - 1;
            if (!fwj[j - 1]) s2 = A.len_[(size_t)j] - (s2 + LON);
            any_fail |= (s2 + LON > A.len_[(size_t)j]) | (s2 < 0);
            st[j] = s2;
        }
        if (bad || any_fail || LON < 5) continue;            // never reaches the trim loop
        if (D.acc_shift[ci] >= 0) return false;              // the engine accepted it: where it was placed depends on what it saw
        int64_t length = LON;
        for (int j = 0; j < n; ++j) {
            const int64_t t1 = truth[(size_t)j].run_up(st[j], st[j] + length);
            if (t1) { for (int i = 0; i < n; ++i) st[i] += t1; length -= t1; }
            const int64_t t2 = truth[(size_t)j].run_down(st[j], st[j] + length);
            length -= t2;
            if (length <= 0) break;
        }
        if (length >= 2) return false;                       // survives on the anchors alone: later bits decide
    }
    return true;
}

// which gaps may take the engine's accept decisions as final (conditions (a)-(d) at the head of the file; the dynamic part of
// (d) is checked again when the gap's turn comes)
void ReplayCtx::classify_gaps() {
    gfinal.assign((size_t)ngaps, 0);
    gr0.assign((size_t)ngaps, 0);
    gr1.assign((size_t)ngaps, 0);
    any_final = false;
    const Aligner::DeviceDecisions& D = A.dev_;
    if (!D.valid || ngaps < 2) return;
    const size_t NR = D.nregions, stride = 2 * (size_t)n;
    for (size_t i = 0; i < D.nfw; ++i) {
        const int g = D.fw[3 * i];
        if (g < 0 || g >= n) return;
        mark_touched(g, D.fw[3 * i + 1], (int64_t)D.fw[3 * i + 1] + D.fw[3 * i + 2]);
    }
    struct S0 {                                 // start[0] of the engine's regions as a random-access range
        const int64_t* c; size_t stride;
        int64_t operator()(size_t r) const { return c[r * stride]; }
    } start0{D.coords, stride};
    auto first_at_least = [&](int64_t v) {
        size_t a = 0, b = NR;
        while (a < b) { const size_t mid = (a + b) >> 1; if (start0(mid) < v) a = mid + 1; else b = mid; }
        return a;
    };
    std::atomic<int> some(0);
    static const bool why = getenv("PB200_REPLAY_DEBUG") != nullptr;
    std::atomic<long> reason[8];
    for (auto& x : reason) x.store(0);
    // (opt-in: on configs[1] it turns 778 of the 3 270 replayed gaps into final ones - the anchors cover less than half of the
    //  genomes there, so most mirrored candidates do survive on them - and the check costs the classification more (+0.9 ms on
    //  4 threads) than the tasks gain (0.1 ms))
    const bool harmless_ok = getenv("PB200_HARMLESS_FOREIGN") != nullptr;
    const long per = 512;
    parallel_chunks(ngaps > 2048 ? A.threads_ : 1, ((long)ngaps + per - 1) / per, [&](long c) {
        bool mine = false;
        // (gap 0 stays with the replay: the reference's first pop precedes its first sort, see run_task)
        for (int j = std::max<int>(1, (int)(c * per)); j < std::min<int>(ngaps, (int)((c + 1) * per)); ++j) {
            const int p0 = gcut[(size_t)j], p1 = gcut[(size_t)j + 1], ni = p1 - p0;
            if (ni > 2 || gtouched[j].load(std::memory_order_relaxed)) { if (why) reason[ni > 2 ? 0 : 1]++; continue; }
            const size_t r0 = first_at_least(glo[(size_t)j * n]), r1 = first_at_least(ghi[(size_t)j * n]);
            if (r1 - r0 < (size_t)ni || r1 > (size_t)INT32_MAX) continue;
            bool ok = true;
            int ninit = 0, nsecond = 0;
            int64_t prev = INT64_MIN;
            for (size_t r = r0; r < r1 && ok; ++r) {
                const uint32_t f = D.flags[r];
                const WindowRec& w = D.wins[r];
                const int64_t* rc = D.coords + r * stride;
                ok = w.ncand >= 0 && (uint64_t)w.cand_off + (uint64_t)w.ncand <= D.ncands && rc[0] > prev;
                if (ok && (f & REC_ORDER_MASK)) ok = (f & REC_ORDER_MASK) == 1u && harmless_ok && foreign_harmless(r);      // (1 = only "a reverse-strand candidate reached the trim loop")
                prev = rc[0];
                const int par = D.parent[r];
                if (par < 0) {
                    bool known = false;         // one of the gap's own initial regions
                    for (int p = p0; p < p1 && !known; ++p)
                        known = std::memcmp(rc, A.rstart(A.initial_regions_[(size_t)order[(size_t)p]]), sizeof(int64_t) * stride) == 0;
                    ok = ok && known;
                    ++ninit;
                    if (f & REC_SECOND) ++nsecond;
                } else if ((size_t)par < r0 || (size_t)par >= r1) ok = false;
            }
            if (!ok || ninit != ni || (ni == 2 && nsecond != 1)) {
                if (why) {
                    uint32_t fl = 0;
                    bool unsearched = false;
                    for (size_t r = r0; r < r1; ++r) { fl |= D.flags[r] & REC_ORDER_MASK; unsearched |= D.wins[r].ncand < 0; }
                    reason[unsearched ? 2 : (fl & 1) ? 3 : (fl & 2) ? 4 : (fl & 4) ? 5 : (fl & 8) ? 6 : 7]++;
                }
                continue;
            }
            gfinal[(size_t)j] = 1;
            gr0[(size_t)j] = (int32_t)r0;
            gr1[(size_t)j] = (int32_t)r1;
            mine = true;
        }
        if (mine) some.store(1, std::memory_order_relaxed);
    });
    any_final = some.load() != 0;
    if (why) fprintf(stderr, "[pb200 replay] gaps kept for the replay: >2 initial regions %ld, touched by a foreign write %ld, unsearched region %ld, "
                             "reverse-strand candidate %ld, second of a pair accepted %ld, accepts not collinear %ld, requeued %ld, other %ld\n",
                     reason[0].load(), reason[1].load(), reason[2].load(), reason[3].load(), reason[4].load(), reason[5].load(), reason[6].load(), reason[7].load());
}

// the engine's accepted MUMs of gap j into the task's output and the layout, in the reference's pop order
void ReplayCtx::apply_final_gap(ReplayTask& T, MumPool& MP, int j, const int64_t* lo, const int64_t* hi) {
    const Aligner::DeviceDecisions& D = A.dev_;
    const size_t stride = 2 * (size_t)n, N = (size_t)n;
    const int nq = n - 1;
    const size_t r0 = (size_t)gr0[(size_t)j], r1 = (size_t)gr1[(size_t)j];
    // room for the gap's MUMs first (their number is known), then filled through plain pointers
    size_t cnt = 0;
    for (size_t r = r0; r < r1; ++r) {
        const WindowRec& w = D.wins[r];
        const int32_t* sh = D.acc_shift + w.cand_off;
        for (int32_t c = 0; c < w.ncand; ++c) cnt += sh[c] >= 0;
    }
    ++T.final_gaps;
    if (!cnt) return;
    const size_t m0 = MP.mums.size(), s0 = MP.start.size();
    MP.mums.resize(m0 + cnt);
    MP.start.resize(s0 + cnt * N);
    MP.fwd.resize(s0 + cnt * N);
    std::memset(&MP.fwd[s0], 1, cnt * N);
    MumRec* mrec = &MP.mums[m0];
    int64_t* st = &MP.start[s0];
    int64_t off = (int64_t)s0;
    for (size_t r = r0; r < r1; ++r) {
        const WindowRec& w = D.wins[r];
        const int64_t* rs = D.coords + r * stride;
        const int64_t rsl = D.slen[r];
        for (int32_t c = 0; c < w.ncand; ++c) {
            const size_t ci = (size_t)w.cand_off + (size_t)c;
            const int64_t sh = D.acc_shift[ci];
            if (sh < 0) continue;
            const int64_t length = D.acc_len[ci];
            st[0] = w.ref_start + D.k[ci] + sh;
            const int32_t* sp = D.sp + ci * (size_t)nq;
            for (int g = 1; g < n; ++g) st[g] = rs[g] + sp[g - 1] + sh;
            for (int g = 0; g < n; ++g) {
                const int64_t a = st[g], b = a + length;
                if (a < lo[g] || b > hi[g] + 1 || length < 2) throw std::logic_error("parsnp_b200: a MUM of a final gap lies outside its task");
                const int64_t wa = a >> 6, wb = (b - 1) >> 6;
                if (wa == wb && wa > (lo[g] >> 6) && wa < (hi[g] >> 6)) {
                    // (one word strictly inside the task's span - the rule: this task is its only writer)
                    uint64_t* word = truth[(size_t)g].words_mut() + wa;
                    __atomic_store_n(word, __atomic_load_n(word, __ATOMIC_RELAXED) | ((~0ull << (a & 63)) & (~0ull >> (63 - ((b - 1) & 63)))), __ATOMIC_RELAXED);
                } else truth[(size_t)g].set_range_owned(a, b, lo[g] >> 6, hi[g] >> 6);
            }
            mrec->length = length;
            mrec->slength = rsl;
            mrec->off = off;
            mrec->alive = true;
            ++mrec;
            st += N;
            off += (int64_t)N;
        }
    }
    T.final_mums += (int64_t)cnt;
    T.mend = MP.mums.size();
}

namespace {
// gap, pos: of an initial region (its gap and its position in `order`), -1 for sub-regions; id < 0: an initial region of a final
// gap that has not been copied into the task's pool (it will be needed only if the gap has to be replayed after all)
struct QE { int64_t s0; int id; int slice; uint64_t hash; int gap; int pos; };
}

// the loop of Aligner::process_queue_exact restricted to one task (its queue is a contiguous piece of the reference's queue:
// every other region has a smaller key and is finished, or a larger key and is untouched)
#if defined(__x86_64__)
#define PB_TICKS() __builtin_ia32_rdtsc()
#else
#define PB_TICKS() ((uint64_t)std::chrono::steady_clock::now().time_since_epoch().count())
#endif
#define PROF_MARK(i) do { if (prof) { const uint64_t x_ = PB_TICKS(); T.pc[i] += x_ - pt; pt = x_; } } while (0)

void ReplayCtx::run_task(int k, int w, std::vector<int>& found, std::vector<int>& children) {
    ReplayTask& T = tasks[k];
    MumPool& MP = wmp[(size_t)w];
    RegionPool& RP = wrp[(size_t)w];
    static const bool prof = getenv("PB200_PROFILE_HOST") != nullptr;
    uint64_t pt = prof ? PB_TICKS() : 0;
    T.reset();
    T.worker = w;
    T.mbegin = T.mend = MP.mums.size();
    RP.n = n;
    RP.coord.clear();
    RP.slen.clear();
    std::vector<int64_t> lo((size_t)n), hi((size_t)n);
    for (int g = 0; g < n; ++g) { lo[g] = hlo[(size_t)g * ntasks + k]; hi[g] = hhi[(size_t)g * ntasks + k]; }
    TaskAccess acc{*this, T, k, lo.data(), hi.data()};
    const size_t cbytes = sizeof(int64_t) * 2 * (size_t)n;
    std::vector<QE> fast;                       // descending start[0]: the back is the front of the reference's vector
    fast.reserve((size_t)(T.last - T.first) + 16);
    for (int p = T.last - 1, j = T.glast - 1; p >= T.first; --p) {
        const int i = order[(size_t)p];
        while (p < gcut[(size_t)j]) --j;
        const int slice = (size_t)i < A.slice_of_initial_.size() ? A.slice_of_initial_[(size_t)i] : -1;
        if (any_final && gfinal[(size_t)j]) { fast.push_back(QE{keys[(size_t)p], -1, slice, 0, j, p}); continue; }
        const int r = A.initial_regions_[(size_t)i];
        const int id = RP.add(A.rp_.start(r), A.rp_.end(r));
        fast.push_back(QE{RP.start(id)[0], id, slice, Aligner::coords_hash(RP.start(id), 2 * n), j, p});
    }
    auto materialize = [&](QE& e) {
        if (e.id >= 0) return;
        const int r = A.initial_regions_[(size_t)order[(size_t)e.pos]];
        e.id = RP.add(A.rp_.start(r), A.rp_.end(r));
        e.hash = Aligner::coords_hash(RP.start(e.id), 2 * n);
    };
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
    auto req = [&](int a, int b) { return std::memcmp(RP.start(a), RP.start(b), cbytes) == 0; };
    std::vector<int64_t> lS((size_t)n), lE((size_t)n), rS((size_t)n), rE((size_t)n);
    int ready_upto = 0;
    int gnext = T.gfirst;                       // first gap of the task that is not done yet
    unsigned rnd = 12345u + (unsigned)k * 2654435761u;
    // slow mode = the reference's literal vector handling while two DIFFERENT regions share a start[0] (see below)
    bool slow = false, first_pop = true;
    int applied = -1;                           // the gap whose MUMs were just taken from the engine
    int pf_gap = T.gfirst - 1;                  // layout rows prefetched up to this gap
    std::vector<QE> lvec;                       // ascending; the front is the front of the reference's vector
    PROF_MARK(0);
    const size_t R = order.size();
    while (slow ? !lvec.empty() : !fast.empty()) {
        if (T.abort.load(std::memory_order_relaxed)) throw TaskAborted();
        if (jitter) {                           // tests: shake the interleavings
            rnd = rnd * 1664525u + 1013904223u;
            if ((rnd >> 16) % jitter == 0) std::this_thread::sleep_for(std::chrono::microseconds((rnd >> 8) % 200));
            else if ((rnd >> 20) % 3 == 0) std::this_thread::yield();
        }
        QE cur;
        if (slow) { cur = lvec.front(); lvec.erase(lvec.begin()); }
        else { cur = fast.back(); fast.pop_back(); }
        ++T.regions;
        // regions come in ascending start[0] order and a sub-region starts at most one base before its parent: every gap that
        // ends before cur.s0 - 1 has nothing left in the queue and never will - its part of the layout is final
        while (gnext < T.glast && gend0[(size_t)gnext] <= cur.s0 - 1) gdone[gnext++].store(1, std::memory_order_release);
        if (cur.gap > pf_gap) {
            // the layout rows of the gaps ahead: nine-odd cache lines per gap that nothing else brings in (a gap costs little
            // more than these misses once its decisions come from the engine)
            for (int j = std::max(pf_gap + 1, cur.gap) + 1; j <= cur.gap + 3 && j < T.glast; ++j)
                for (int g = 0; g < n; ++g) __builtin_prefetch(truth[(size_t)g].words() + (glo[(size_t)j * n + g] >> 6), 1);
            pf_gap = cur.gap + 2;
        }
        PROF_MARK(1);
        if (cur.gap >= 0 && any_final) {
            if (cur.gap == applied) continue;                   // (the gap's other initial region)
            if (!slow && gfinal[(size_t)cur.gap] && !gtouched[cur.gap].load(std::memory_order_relaxed)) {
                apply_final_gap(T, MP, cur.gap, lo.data(), hi.data());
                applied = cur.gap;
                first_pop = false;
                PROF_MARK(6);
                continue;
            }
        }
        materialize(cur);
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
        int ci = C ? C->lookup(RP.start(cur.id), cur.hash) : -1;
        if (ci < 0) { C = &A.main_cache_; ci = A.main_cache_.lookup(RP.start(cur.id), cur.hash); }
        if (ci < 0 && T.local) { C = T.local.get(); ci = C->lookup(RP.start(cur.id), cur.hash); }
        if (ci < 0) {                           // a region the speculation did not predict: search it now
            const double t0 = now_s();
            if (!T.local) T.local.reset(new Aligner::CandCache);
            A.search_regions(*T.local, RP, std::vector<int>(1, cur.id), false);
            T.t_search += now_s() - t0;
            ++T.misses;
            C = T.local.get();
            ci = C->lookup(RP.start(cur.id), cur.hash);
        }
        found.clear();
        PROF_MARK(2);
        A.accept_candidates_t(RP.start(cur.id), RP.end(cur.id), RP.slen[(size_t)cur.id], *C, ci, acc, MP, found, false);
        T.mend = MP.mums.size();
        PROF_MARK(3);
        children.clear();
        int64_t lsl = 0;
        for (size_t i = 0; i < found.size(); ++i) {
            const MumRec& m = MP.mums[(size_t)found[i]];
            const int64_t* ms = &MP.start[(size_t)m.off];
            if (i == 0) lsl = A.det_region_t(acc, ms, m.length, true, lS.data(), lE.data());
            const int64_t rsl = A.det_region_t(acc, ms, m.length, false, rS.data(), rE.data());
            if (lsl > A.prm_.q) children.push_back(RP.add(lS.data(), lE.data()));
            if (rsl > A.prm_.q) children.push_back(RP.add(rS.data(), rE.data()));
            if (i + 1 < found.size()) {
                const MumRec& m2 = MP.mums[(size_t)found[i + 1]];
                lsl = A.det_region_t(acc, &MP.start[(size_t)m2.off], m2.length, true, lS.data(), lE.data());
            }
        }
        PROF_MARK(4);
        // sort + drop adjacent duplicates (src/parsnp.cpp:291-306).  With distinct keys: an ordered insert.
        if (!slow) {
            bool distinct_tie = false;
            for (size_t a = 0; a < children.size() && !distinct_tie; ++a) {
                const int64_t key = RP.start(children[a])[0];
                const size_t pos = fast_pos(key);
                if (pos < fast.size() && fast[pos].s0 == key) {
                    materialize(fast[pos]);
                    if (!req(fast[pos].id, children[a])) distinct_tie = true;
                }
                for (size_t b = 0; b < a && !distinct_tie; ++b)
                    if (key == RP.start(children[b])[0] && !req(children[a], children[b])) distinct_tie = true;
            }
            if (!distinct_tie) {
                for (int ch : children) {                       // identical duplicates collapse (the first one stays)
                    const int64_t key = RP.start(ch)[0];
                    const size_t pos = fast_pos(key);
                    if (pos < fast.size() && fast[pos].s0 == key) continue;
                    fast.insert(fast.begin() + (long)pos, QE{key, ch, cur.slice, Aligner::coords_hash(RP.start(ch), 2 * n), -1, -1});
                }
                first_pop = false;
                PROF_MARK(5);
                continue;
            }
            // A tie between DIFFERENT regions: the order is what the reference's unstable std::sort over its WHOLE vector makes
            // of it.  That vector is known here: this task's pending regions (ascending), then every initial region behind the
            // task (untouched so far, ascending, all keys above the task's), then the new children - so the call is replayed
            // literally on (key, tag) records: same algorithm, same comparisons, same permutation.  Only before the very first
            // sort (the reference's vector is still in push order) the run goes back to the sequential loop.
            if (getenv("PB200_REPLAY_DEBUG")) fprintf(stderr, "[pb200 replay] task %d: tie between different regions, literal queue until it is gone\n", k);
            if (k == 0 && first_pop) throw NeedFallback();
            for (QE& e : fast) materialize(e);
            lvec.assign(fast.rbegin(), fast.rend());
            fast.clear();
            slow = true;
        }
        first_pop = false;
        ++T.slow_iters;
        const double tslow0 = prof ? now_s() : 0;
        const size_t before = lvec.size();
        for (int ch : children) lvec.push_back(QE{RP.start(ch)[0], ch, cur.slice, Aligner::coords_hash(RP.start(ch), 2 * n), -1, -1});
        if (!lvec.empty()) {
            std::vector<int64_t> lkeys(lvec.size());
            for (size_t i = 0; i < lvec.size(); ++i) lkeys[i] = lvec[i].s0;
            std::sort(lkeys.begin(), lkeys.end());
            bool tie = false;
            for (size_t i = 1; i < lkeys.size() && !tie; ++i) tie = lkeys[i] == lkeys[i - 1];
            if (!tie) {
                // (bounded insertion sort or std::sort in the reference's loop: with distinct keys both give THE ascending order)
                std::stable_sort(lvec.begin(), lvec.end(), [](const QE& a, const QE& b) { return a.s0 < b.s0; });
            } else {
                std::vector<std::pair<int64_t, int>> gv;
                gv.reserve(lvec.size() + (R - (size_t)T.last));
                for (size_t i = 0; i < before; ++i) gv.emplace_back(lvec[i].s0, (int)i);
                for (size_t p = (size_t)T.last; p < R; ++p) gv.emplace_back(keys[p], -1);
                for (size_t i = before; i < lvec.size(); ++i) gv.emplace_back(lvec[i].s0, (int)i);
                std::sort(gv.begin(), gv.end(), [](const std::pair<int64_t, int>& a, const std::pair<int64_t, int>& b) { return a.first < b.first; });
                std::vector<QE> nl;
                nl.reserve(lvec.size());
                for (const auto& e : gv) if (e.second >= 0) nl.push_back(lvec[(size_t)e.second]);
                lvec.swap(nl);
            }
        }
        for (size_t m = 0; m + 1 < lvec.size();) {
            if (req(lvec[m].id, lvec[m + 1].id)) lvec.erase(lvec.begin() + (long)m);
            else ++m;
        }
        bool strict = true;
        for (size_t m = 0; m + 1 < lvec.size() && strict; ++m) strict = lvec[m].s0 < lvec[m + 1].s0;
        if (strict) {
            fast.assign(lvec.rbegin(), lvec.rend());
            lvec.clear();
            slow = false;
        }
        if (prof) fprintf(stderr, "[pb200 replay] task %d: literal queue iteration %.2f ms\n", k, (now_s() - tslow0) * 1e3);
    }
    while (gnext < T.glast) gdone[gnext++].store(1, std::memory_order_release);
}

void ReplayCtx::worker(int w) {
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
        try { run_task(k, w, found, children); }
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

Aligner::~Aligner() {
    if (spec_thread_.joinable()) spec_thread_.join();
    if (replay_prep_thread_.joinable()) replay_prep_thread_.join();
    delete replay_ctx_;
}

// The task structure depends on the anchors only: with the engine following the recursion on the device, it is built on a
// second thread while the host would otherwise wait for the GPU.
void Aligner::replay_prepare_async() {
    replay_prep_thread_ = std::thread([this] {
        // (with few cores per rank a second set of helpers only competes with the main thread's own passes: at 4 threads per
        //  rank, 8 ranks on 32 cores, the replay then waited longer for this structure than the helpers saved)
        parallel_use_second_pool(threads_ >= 8);
        try { replay_ctx_ = replay_prepare(); } catch (...) { replay_prep_error_ = std::current_exception(); }
        replay_prepared_ = true;
    });
}

bool Aligner::do_work_parallel() {
    if (replay_prep_thread_.joinable()) replay_prep_thread_.join();
    if (replay_prep_error_) std::rethrow_exception(replay_prep_error_);
    if (!replay_prepared_) { replay_ctx_ = replay_prepare(); replay_prepared_ = true; }
    if (!replay_ctx_) return false;
    std::unique_ptr<ReplayCtx> holder(replay_ctx_);
    replay_ctx_ = nullptr;
    return replay_run(*holder);
}

ReplayCtx* Aligner::replay_prepare() {
    int W = replay_threads_ > 0 ? replay_threads_ : threads_;
    if (const char* ew = getenv("PB200_REPLAY_THREADS")) W = std::max(1, atoi(ew));
    const size_t R = initial_regions_.size();
    const char* et = getenv("PB200_REPLAY_TASK");               // initial regions per task (tests: 1 = one gap per task)
    const char* em = getenv("PB200_REPLAY_MODE");               // "seq" = never, "par" = also for tiny inputs / one worker
    const bool force = em && std::strcmp(em, "par") == 0;
    if (em && std::strcmp(em, "seq") == 0) return nullptr;
    const bool dbg = getenv("PB200_REPLAY_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[pb200 replay] W %d R %zu trace %d pipeline %d\n", W, R, (int)trace_on_, (int)pipeline_);
    if (trace_on_ || !pipeline_ || R == 0) return nullptr;
    if (!force && (W < 2 || R < 512)) return nullptr;
    const double tsetup0 = now_s();
    static const bool profs = getenv("PB200_PROFILE_HOST") != nullptr;
    double tsec[8] = {tsetup0, 0, 0, 0, 0, 0, 0, 0};
    std::unique_ptr<ReplayCtx> XP(new ReplayCtx(*this));
    ReplayCtx& X = *XP;
    if (const char* ej = getenv("PB200_REPLAY_JITTER")) X.jitter = (unsigned)std::max(0, atoi(ej));
    // P1: the keys of the initial regions are distinct, so the reference's first sort has one possible outcome.  (Push order is
    // NOT ascending: the right side of anchor i starts one base behind the left side of anchor i+1 - the same gap twice.)
    std::vector<int64_t> pushkey(R);
    {
        const long per_blk = 4096;
        parallel_chunks(R > 16384 ? threads_ : 1, ((long)R + per_blk - 1) / per_blk, [&](long c) {
            for (size_t i = (size_t)c * per_blk; i < std::min(R, (size_t)(c + 1) * per_blk); ++i) pushkey[i] = rstart(initial_regions_[i])[0];
        });
    }
    X.order.resize(R);
    for (size_t i = 0; i < R; ++i) X.order[i] = (int)i;
    {   // (push order is ascending up to neighbour swaps: one insertion pass; anything else falls back to a real sort)
        size_t moves = 0;
        for (size_t i = 1; i < R && moves <= 4 * R; ++i) {
            const int x = X.order[i];
            const int64_t kx = pushkey[(size_t)x];
            size_t j = i;
            while (j > 0 && pushkey[(size_t)X.order[j - 1]] > kx) { X.order[j] = X.order[j - 1]; --j; ++moves; }
            X.order[j] = x;
        }
        if (moves > 4 * R) {
            for (size_t i = 0; i < R; ++i) X.order[i] = (int)i;
            std::stable_sort(X.order.begin(), X.order.end(), [&](int a, int b) { return pushkey[(size_t)a] < pushkey[(size_t)b]; });
        }
    }
    tsec[1] = now_s();
    X.keys.resize(R);                            // start[0] of the initial regions in ascending order
    for (size_t i = 0; i < R; ++i) X.keys[i] = pushkey[(size_t)X.order[i]];
    for (size_t i = 1; i < R; ++i)
        if (!(X.keys[i - 1] < X.keys[i])) {
            if (dbg) fprintf(stderr, "[pb200 replay] two initial regions tie on start[0]: sequential\n");
            return nullptr;
        }
    size_t pos0 = 0;
    while (pos0 < R && X.order[pos0] != 0) ++pos0;
    // spans of the initial regions: between the anchor bits that bound them, per genome (the layout holds exactly the anchors);
    // cut[p] = region p starts behind everything before it in every genome.  Collinear anchors make the span ends ascend with p,
    // so "everything before" is the previous region; anything else declines.
    pod_vector<int64_t> rlo(R * (size_t)n_), rhi(R * (size_t)n_);       // (filled in parallel below: the workers touch the new pages)
    std::vector<uint8_t> cut(R, 0);
    std::atomic<int> bad(0);
    const bool have_lo = anchors_on_device_ && initial_lo_.size() == R * (size_t)n_;
    const bool check_spans = have_lo && getenv("PB200_CHECK_SPANS") != nullptr;
    {
        const long per_blk = 512;
        parallel_chunks(threads_, ((long)R + per_blk - 1) / per_blk, [&](long c) {
            for (size_t p = (size_t)c * per_blk; p < std::min(R, (size_t)(c + 1) * per_blk); ++p) {
                const int r = initial_regions_[(size_t)X.order[p]];
                const int32_t* dlo = have_lo ? &initial_lo_[(size_t)X.order[p] * (size_t)n_] : nullptr;
                for (int g = 0; g < n_; ++g) {
                    const int64_t s = rstart(r)[g], e = rend(r)[g];
                    if (s < 0 || e > len_[g] || e < s) { bad.store(1); continue; }
                    if (dlo) {
                        // the engine made these regions (determineRegion on the anchors' layout): it knows the set bit on their left;
                        // on the right a non-empty region [s, e] ends one base before one (an anchor's start, the next set bit, or
                        // the sentinel) - no walk over nine bitmap rows per region
                        rlo[p * (size_t)n_ + g] = dlo[g];
                        rhi[p * (size_t)n_ + g] = std::min(e + 1, len_[g]);
                        if (check_spans) {
                            int64_t a = truth_.layout[g].prev_set(s > 0 ? s - 1 : 0);
                            if (a < 0) a = 0;
                            const int64_t b = std::min(truth_.layout[g].next_set(e, len_[g] + 1), len_[g]);
                            if (a != dlo[g] || b != rhi[p * (size_t)n_ + g]) bad.store(2);
                        }
                        continue;
                    }
                    int64_t a = truth_.layout[g].prev_set(s > 0 ? s - 1 : 0);
                    if (a < 0) a = 0;
                    const int64_t b = truth_.layout[g].next_set(e, len_[g] + 1);   // (the sentinel bit at len ends every scan)
                    rlo[p * (size_t)n_ + g] = a;
                    rhi[p * (size_t)n_ + g] = std::min(b, len_[g]);
                }
            }
        });
        if (bad.load() == 2) throw std::logic_error("parsnp_b200: the engine's region bounds differ from the layout (PB200_CHECK_SPANS)");
        if (bad.load()) return nullptr;
        tsec[2] = now_s();
        parallel_chunks(threads_, ((long)R + per_blk - 1) / per_blk, [&](long c) {
            for (size_t p = std::max<size_t>(1, (size_t)c * per_blk); p < std::min(R, (size_t)(c + 1) * per_blk); ++p) {
                bool ok = true;
                for (int g = 0; g < n_; ++g) {
                    if (rhi[(p - 1) * (size_t)n_ + g] > rhi[p * (size_t)n_ + g]) bad.store(1);          // not ascending
                    ok &= rhi[(p - 1) * (size_t)n_ + g] <= rlo[p * (size_t)n_ + g];
                }
                cut[p] = ok ? 1 : 0;
            }
        });
        if (bad.load()) {
            if (dbg) fprintf(stderr, "[pb200 replay] the spans of the initial regions do not ascend in every genome: sequential\n");
            return nullptr;
        }
    }
    tsec[3] = now_s();
    // gaps: runs of regions between cuts (the two regions of one anchor gap, and whatever else overlaps, stay together);
    // tasks: runs of gaps with >= `per` regions
    const size_t per = et ? (size_t)std::max(1, atoi(et)) : std::min<size_t>(128, std::max<size_t>(8, R / ((size_t)W * 24) + 1));
    std::vector<int>& gcut = X.gcut;             // gap j = positions [gcut[j], gcut[j+1])
    gcut.assign(1, 0);
    for (size_t p = 1; p < R; ++p) if (cut[p]) gcut.push_back((int)p);
    gcut.push_back((int)R);
    X.ngaps = (int)gcut.size() - 1;
    X.glo.resize((size_t)X.ngaps * n_);
    X.ghi.resize((size_t)X.ngaps * n_);
    X.gend0.resize((size_t)X.ngaps);
    X.gdone.reset(new std::atomic<uint8_t>[(size_t)X.ngaps]);
    X.gtouched.reset(new std::atomic<uint8_t>[(size_t)X.ngaps]);
    parallel_chunks(X.ngaps > 4096 ? threads_ : 1, ((long)X.ngaps + 1023) / 1024, [&](long c) {
        for (int j = (int)c * 1024; j < std::min(X.ngaps, (int)(c + 1) * 1024); ++j) {
            X.gdone[j].store(0, std::memory_order_relaxed);
            X.gtouched[j].store(0, std::memory_order_relaxed);
            int64_t e0 = 0;
            for (int g = 0; g < n_; ++g) {
                int64_t lo = INT64_MAX, hi = -1;
                for (int p = gcut[(size_t)j]; p < gcut[(size_t)j + 1]; ++p) {
                    lo = std::min(lo, rlo[(size_t)p * n_ + g]);
                    hi = std::max(hi, rhi[(size_t)p * n_ + g]);
                }
                X.glo[(size_t)j * n_ + g] = (int32_t)lo;
                X.ghi[(size_t)j * n_ + g] = (int32_t)hi;
            }
            for (int p = gcut[(size_t)j]; p < gcut[(size_t)j + 1]; ++p) e0 = std::max(e0, rend(initial_regions_[(size_t)X.order[(size_t)p]])[0]);
            X.gend0[(size_t)j] = e0;
        }
    });
    // P2: collinear anchors - the gaps ascend in every genome (neighbours may share a boundary anchor bit)
    parallel_chunks(X.ngaps > 4096 ? threads_ : 1, ((long)X.ngaps + 1023) / 1024, [&](long c) {
        for (int j = (int)c * 1024; j < std::min(X.ngaps - 1, (int)(c + 1) * 1024); ++j)
            for (int g = 0; g < n_; ++g)
                if (X.ghi[(size_t)j * n_ + g] > X.glo[(size_t)(j + 1) * n_ + g] || X.glo[(size_t)j * n_ + g] > X.ghi[(size_t)j * n_ + g]) bad.store(1);
    });
    if (bad.load()) {
        if (dbg) fprintf(stderr, "[pb200 replay] gaps not in order in some genome: sequential\n");
        return nullptr;
    }
    tsec[4] = now_s();
    // the reference's first pop is the first region PUSHED (see run_task): it must belong to the first gap
    if ((int)pos0 >= gcut[1]) {
        if (dbg) fprintf(stderr, "[pb200 replay] the first region pushed is not in the first gap: sequential\n");
        return nullptr;
    }
    std::vector<int> tcut(1, 0);                 // task k = gaps [tcut[k], tcut[k+1])
    {
        size_t count = 0;
        for (int j = 0; j < X.ngaps; ++j) {
            if (count >= per && j > 0) { tcut.push_back(j); count = 0; }
            count += (size_t)(gcut[(size_t)j + 1] - gcut[(size_t)j]);
        }
        tcut.push_back(X.ngaps);
    }
    X.ntasks = (int)tcut.size() - 1;
    if (!force && X.ntasks < 2 * W) {
        if (dbg) fprintf(stderr, "[pb200 replay] only %d independent tasks for %d workers: sequential\n", X.ntasks, W);
        return nullptr;
    }
    X.tasks.reset(new ReplayTask[(size_t)X.ntasks]);
    X.hlo.assign((size_t)n_ * X.ntasks, 0);
    X.hhi.assign((size_t)n_ * X.ntasks, 0);
    for (int k = 0; k < X.ntasks; ++k) {
        ReplayTask& T = X.tasks[k];
        T.gfirst = tcut[(size_t)k];
        T.glast = tcut[(size_t)k + 1];
        T.first = gcut[(size_t)T.gfirst];
        T.last = gcut[(size_t)T.glast];
        for (int g = 0; g < n_; ++g) {
            X.hlo[(size_t)g * X.ntasks + k] = X.glo[(size_t)T.gfirst * n_ + g];
            X.hhi[(size_t)g * X.ntasks + k] = X.ghi[(size_t)(T.glast - 1) * n_ + g];
        }
    }
    tsec[5] = now_s();
    X.S.resize((size_t)n_);
    parallel_chunks(threads_, (long)n_, [&](long g) { X.S[(size_t)g] = truth_.layout[(size_t)g]; });
    tsec[6] = now_s();
    X.nw = std::max(1, std::min(W, X.ntasks));
    X.log_slab.reset(new ReplayTask::FRead[(size_t)X.ntasks * ReplayTask::LOG_CAP]);
    for (int k = 0; k < X.ntasks; ++k) X.tasks[k].flog = X.log_slab.get() + (size_t)k * ReplayTask::LOG_CAP;
    X.t_setup = now_s() - tsetup0;
    if (profs) fprintf(stderr, "[pb200 replay setup ms] order %.2f spans %.2f cuts %.2f gaps %.2f tasks %.2f S %.2f logs %.2f\n", (tsec[1] - tsec[0]) * 1e3, (tsec[2] - tsec[1]) * 1e3,
                       (tsec[3] - tsec[2]) * 1e3, (tsec[4] - tsec[3]) * 1e3, (tsec[5] - tsec[4]) * 1e3, (tsec[6] - tsec[5]) * 1e3, (now_s() - tsec[6]) * 1e3);
    return XP.release();
}

bool Aligner::replay_run(ReplayCtx& X) {
    const size_t R = initial_regions_.size();
    const int nw = X.nw;
    X.wmp.resize((size_t)nw);
    X.wrp.resize((size_t)nw);
    {
        const size_t est = (size_t)stats_.anchors * 4 / (size_t)nw + 1024;       // (about 3 recursion MUMs per anchor on divergent genomes)
        for (auto& m : X.wmp) { m.mums.reserve(est); m.start.reserve(est * (size_t)n_); m.fwd.reserve(est * (size_t)n_); }
    }
    const double tc0 = now_s();
    X.classify_gaps();
    if (disc_index_deferred_) {
        // the tasks look regions up by their coordinates - but not those of final gaps: the index is built without them (a final
        // gap that has to be replayed after all, because a foreign write touched it, searches its regions on demand)
        std::vector<uint8_t> skip(dev_.nregions, 0);
        if (X.any_final)
            parallel_chunks(X.ngaps > 4096 ? threads_ : 1, ((long)X.ngaps + 1023) / 1024, [&](long c) {
                for (int j = (int)c * 1024; j < std::min(X.ngaps, (int)(c + 1) * 1024); ++j)
                    if (X.gfinal[(size_t)j]) std::memset(&skip[(size_t)X.gr0[(size_t)j]], 1, (size_t)(X.gr1[(size_t)j] - X.gr0[(size_t)j]));
            });
        build_region_index(&skip);
    }
    const double tr0 = now_s();
    // (the process-wide pool of sleeping threads: no thread is created per alignment.  If the pool is busy - the host's own
    //  level-by-level speculation uses it when the engine does not follow the recursion itself - the workers run one after the
    //  other on this thread: the first takes every task)
    if (getenv("PB200_REPLAY_OWN_THREADS")) {                   // tests: real concurrency whatever the pool is doing
        std::vector<std::thread> th;
        for (int w = 1; w < nw; ++w) th.emplace_back([&X, w] { X.worker(w); });
        X.worker(0);
        for (auto& t : th) t.join();
    } else {
        parallel_chunks(nw, (long)nw, [&X](long w) { X.worker((int)w); });
    }
    if (X.error) std::rethrow_exception(X.error);
    if (getenv("PB200_PROFILE_HOST")) {
        int64_t fg = 0, fm = 0;
        for (int k = 0; k < X.ntasks; ++k) { fg += X.tasks[k].final_gaps; fm += X.tasks[k].final_mums; }
        fprintf(stderr, "[pb200 replay ms] setup %.2f classify %.2f tasks %.2f (%d workers, %d tasks, %d gaps, %lld final with %lld MUMs)\n", X.t_setup * 1e3,
                (tr0 - tc0) * 1e3, (now_s() - tr0) * 1e3, nw, X.ntasks, X.ngaps, (long long)fg, (long long)fm);
    }

    // ---- the finished tasks' MUMs into the pools, task after task = the reference's push order
    const double tm0 = now_s();
    const int F = std::min(X.ntasks, X.fallback_from);
    std::vector<size_t> base((size_t)F + 1, 0);
    for (int k = 0; k < F; ++k) base[(size_t)k + 1] = base[(size_t)k] + (X.tasks[k].mend - X.tasks[k].mbegin);
    const size_t M = base[(size_t)F], m0 = mp_.mums.size(), s0 = mp_.start.size(), a0 = all_mums_.size();
    mp_.mums.resize(m0 + M);
    mp_.start.resize(s0 + M * (size_t)n_);
    mp_.fwd.resize(s0 + M * (size_t)n_);
    all_mums_.resize(a0 + M);
    parallel_chunks(threads_, (long)F, [&](long k) {
        const ReplayTask& T = X.tasks[k];
        const size_t b = base[(size_t)k], cnt = T.mend - T.mbegin;
        if (!cnt) return;
        const MumPool& MP = X.wmp[(size_t)T.worker];
        const size_t src = (size_t)MP.mums[T.mbegin].off;             // (a task's rows are contiguous in its worker's arena)
        std::memcpy(&mp_.start[s0 + b * (size_t)n_], &MP.start[src], cnt * (size_t)n_ * sizeof(int64_t));
        std::memcpy(&mp_.fwd[s0 + b * (size_t)n_], &MP.fwd[src], cnt * (size_t)n_);
        for (size_t i = 0; i < cnt; ++i) {
            MumRec m = MP.mums[T.mbegin + i];
            m.off = (int64_t)(s0 + (b + i) * (size_t)n_);
            mp_.mums[m0 + b + i] = m;
            all_mums_[a0 + b + i] = (int)(m0 + b + i);
        }
    });
    // ---- the ascending start[0] order of everything found so far, for the LCB stage (Aligner::sort_final_mums): every task's
    // MUMs sorted on their own (tasks ascend on the reference), merged with the anchors (accepted in ascending order)
    sorted_hint_.clear();
    if (F == X.ntasks && a0 == m0) {             // (no sequential tail; all_mums_ ids == pool ids)
        std::vector<int> rec(M);
        parallel_chunks(threads_, (long)F, [&](long k) {
            const size_t b = base[(size_t)k], cnt = base[(size_t)k + 1] - b;
            for (size_t i = 0; i < cnt; ++i) rec[b + i] = (int)(m0 + b + i);
            std::stable_sort(rec.begin() + (long)b, rec.begin() + (long)(b + cnt), [&](int x, int y) { return mp_.start[(size_t)mp_.mums[(size_t)x].off] < mp_.start[(size_t)mp_.mums[(size_t)y].off]; });
        });
        auto key = [&](int id) { return mp_.start[(size_t)mp_.mums[(size_t)id].off]; };
        sorted_hint_.resize(a0 + M);
        const long P = std::max<long>(1, std::min<long>(threads_ * 4, (long)(a0 / 1024) + 1));
        std::vector<size_t> as((size_t)P + 1), rs_((size_t)P + 1);
        for (long c = 0; c <= P; ++c) {
            as[(size_t)c] = a0 * (size_t)c / (size_t)P;
            if (c == 0) rs_[0] = 0;
            else if (c == P) rs_[(size_t)P] = M;
            else {
                const int64_t kx = key(all_mums_[as[(size_t)c]]);       // recursion MUMs below this anchor go to the parts before
                rs_[(size_t)c] = (size_t)(std::lower_bound(rec.begin(), rec.end(), kx, [&](int id, int64_t v) { return key(id) < v; }) - rec.begin());
            }
        }
        parallel_chunks(threads_, P, [&](long c) {
            std::merge(all_mums_.begin() + (long)as[(size_t)c], all_mums_.begin() + (long)as[(size_t)c + 1], rec.begin() + (long)rs_[(size_t)c],
                       rec.begin() + (long)rs_[(size_t)c + 1], sorted_hint_.begin() + (long)(as[(size_t)c] + rs_[(size_t)c]),
                       [&](int x, int y) { return key(x) < key(y); });
        });
    }
    uint64_t pcs[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < F; ++k) {
        const ReplayTask& T = X.tasks[k];
        stats_.replay_foreign_reads += T.n_fread;
        stats_.replay_foreign_writes += T.n_fwrite;
        stats_.replay_misses += T.misses;
        stats_.slow_queue_iters += T.slow_iters;
        stats_.replay_final_gaps += T.final_gaps;
        stats_.replay_final_mums += T.final_mums;
        for (int i = 0; i < 7; ++i) pcs[i] += T.pc[i];
        stats_.t_replay_wait += T.t_wait / nw;
        stats_.t_replay_search += T.t_search;
    }
    if (getenv("PB200_PROFILE_HOST"))
        fprintf(stderr, "[pb200 replay tasks cycles] setup %llu pop %llu lookup %llu accept %llu det_region %llu queue %llu final gaps %llu\n", (unsigned long long)pcs[0],
                (unsigned long long)pcs[1], (unsigned long long)pcs[2], (unsigned long long)pcs[3], (unsigned long long)pcs[4], (unsigned long long)pcs[5], (unsigned long long)pcs[6]);
    stats_.replay_tasks = F;
    stats_.replay_gaps = X.ngaps;
    stats_.replay_restarts = X.restarts;
    stats_.replay_workers = nw;
    stats_.t_replay_merge += now_s() - tm0;
    if (F < X.ntasks) {
        // a tie inside task F: rebuild the layout as the reference has it when it reaches that task (anchors + the MUMs of
        // everything before), then its own loop from there
        stats_.replay_fallback = 1;
        build_region_index(nullptr);
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
