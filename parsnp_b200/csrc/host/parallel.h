// parsnp_b200 - host-side work sharing for the orchestrator's data-parallel passes (gathers, pairwise tests, the
// speculative level).  The reference's host code is single-threaded apart from the `#pragma omp parallel for` over LCBs in
// Aligner::writeOutput (src/parsnp.cpp:648); these passes have no counterpart there.
//
// One process-wide pool of sleeping workers (condition variable, no spinning: a co-scheduled rank on the same host is
// never starved).  A pass hands out chunk indices through an atomic counter; the caller works too.  Calls from a second
// thread while the pool is busy, and nested calls, simply run their chunks inline.
#pragma once
#include <functional>

namespace pb200 {

void parallel_run(int nthreads, long nchunks, const std::function<void(long)>& fn);

template <class F>
inline void parallel_chunks(int nthreads, long nchunks, F&& fn) {
    if (nthreads <= 1 || nchunks <= 1) { for (long c = 0; c < nchunks; ++c) fn(c); return; }
    const std::function<void(long)> f(std::ref(fn));
    parallel_run(nthreads, nchunks, f);
}

}  // namespace pb200
