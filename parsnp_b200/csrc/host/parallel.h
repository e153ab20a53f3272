// parsnp_b200 - host-side work sharing for the orchestrator's data-parallel passes (gathers, pairwise tests, the
// speculative level).  The reference's host code is single-threaded apart from the `#pragma omp parallel for` over LCBs in
// Aligner::writeOutput (src/parsnp.cpp:648); these passes have no counterpart there.
//
// One process-wide pool of sleeping workers (condition variable, no spinning: a co-scheduled rank on the same host is
// never starved).  A pass hands out chunk indices through an atomic counter; the caller works too.  Calls from a second
// thread while the pool is busy, and nested calls, simply run their chunks inline.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <utility>

namespace pb200 {

void parallel_run(int nthreads, long nchunks, const std::function<void(long)>& fn);
// A thread that works BESIDE the orchestrator's main thread for a while (the replay's task structure is built while the main
// thread waits for the engine and indexes its answer) takes its helpers from a second pool: on the first one its passes would
// run inline whenever the main thread is inside one of its own.  Per calling thread.
void parallel_use_second_pool(bool on);

template <class F>
inline void parallel_chunks(int nthreads, long nchunks, F&& fn) {
    if (nthreads <= 1 || nchunks <= 1) { for (long c = 0; c < nchunks; ++c) fn(c); return; }
    const std::function<void(long)> f(std::ref(fn));
    parallel_run(nthreads, nchunks, f);
}

// std::sort(v, v + n, by .first) - the SAME permutation as that call, ties included - on several threads.
// The reference orders its MUM list with std::sort on start[0] (src/TMum.cpp:151, src/parsnp.cpp:331, 2566); when two MUMs
// share a start the outcome is whatever libstdc++'s introsort does with them, so the host replays that call literally on
// (start0, id) records (aligner.cpp: sort_final_mums) - 17 ms for the 195 405 MUMs of configs[1] on one thread, three times
// per alignment.  Here the library's own steps are driven in parallel: the two sides of every partition are independent
// (std::__unguarded_partition_pivot per range, level by level, same depth budget), small ranges finish with the library's
// std::__introsort_loop, and the final insertion pass runs per range (every range starts at a partition cut: nothing moves
// across it).  Same comparisons on the same data in every range, hence the same permutation.
// tie_keys (optional): the keys that occur more than once, ascending - then only the ranges that still hold two elements with
// the same key are followed literally; every other range has ONE ascending order and is sorted by the fastest means.
// hint (optional, with tie_keys): the same records in ascending key order (tied ones in any order) - such a range is then copied.
void literal_std_sort_by_first(std::pair<int64_t, int>* v, size_t n, int threads, const int64_t* tie_keys = nullptr, size_t ntie = 0,
                               const std::pair<int64_t, int>* hint = nullptr);

}  // namespace pb200
