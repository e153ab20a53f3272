// oracle/hosttest_capi.cpp - TEST-ONLY library entry points: the product's host orchestrator
// (parsnp_b200/csrc/host/*.cpp) linked with the CPU search backends of oracle/ so that `-m "not gpu"` tests can
// exercise the host logic (queue order, trim, LCB chaining) without a GPU.  Never linked into libparsnp_b200.so.
#include "../parsnp_b200/csrc/host/result.h"
#include "../parsnp_b200/csrc/host/sharded.h"
#include <algorithm>
#include <memory>
#include <vector>
#include "../parsnp_b200/csrc/host/parallel.h"
#include <stdexcept>
#include <cstdlib>
#include <cstring>

namespace pb200_oracle {
pb200::SearchBackend* make_spec_backend();
pb200::SearchBackend* make_ref_backend();
pb200::SearchBackend* make_replay_backend();
pb200::SearchBackend* make_discover_backend();
pb200::StagedWindowEngine* as_staged(pb200::SearchBackend* b);
struct SpecCand { int32_t k, lon; std::vector<int32_t> sp; std::vector<uint8_t> fwd; };
void spec_window(const uint8_t* R, int64_t n, int nq, const uint8_t* const* Q, const int64_t* m, int minsize, std::vector<SpecCand>& out);
void spec_lrp(const uint8_t* R, int64_t n, std::vector<int32_t>& lrp);
}

extern "C" {
// backend: 0 = brute-force specification (oracle/mumspec.cpp), 1 = real csgmum (oracle/ref_backend.cpp),
//          2 = csgmum behind the record/replay table (host orchestrator timing, tools/host_bench.py)
//          3 = 2 + the CPU emulation of the engine's device-resident discovery (oracle/discover_emul.cpp): the host's
//              "final gaps" path without a GPU
int pbtest_align(int backend, int n, const uint8_t* const* seqs, const int64_t* lens, const pb200_params* prm, pb200_result** out) {
    try {
        pb200::SearchBackend* be = backend == 0 ? pb200_oracle::make_spec_backend() : backend == 2 ? pb200_oracle::make_replay_backend()
                                  : backend == 3 ? pb200_oracle::make_discover_backend() : pb200_oracle::make_ref_backend();
        pb200::Aligner a(n, seqs, lens, pb200::to_align_params(prm), be);
        a.enable_trace(prm->flags & PB200_FLAG_TRACE_WINDOWS);
        a.set_speculate(!(prm->flags & PB200_FLAG_NO_SPECULATION));
        a.set_threads(pb200::default_host_threads());
        bool ok = a.run();
        *out = pb200::make_result(a, (prm->flags & PB200_FLAG_UNALIGNED) != 0);
        delete be;
        return ok ? 0 : PB200_ERR_NO_MUMS;
    } catch (const std::exception& e) { pb200::g_last_error = e.what(); return PB200_ERR_INTERNAL; }
}
// the N>1 path on CPU: host orchestrator replicated on every rank, search sharded over ranks (queries for large windows,
// windows for the recursion batches, host/sharded.cpp) on the csgmum backend; collectives = caller callbacks (gloo)
namespace {
struct CbComm : public pb200::Comm {
    pb200_allgather_cb ag; pb200_allreduce_cb ar; pb200_bcast_cb bc; void* user;
    void allgather(const void* s, void* r, size_t b, bool d) override { if (ag(user, s, r, (int64_t)b, d) != 0) throw std::runtime_error("allgather failed"); }
    void allreduce_i32(int32_t* p, size_t c, bool mx, bool d) override { if (ar(user, p, (int64_t)c, mx, d) != 0) throw std::runtime_error("allreduce failed"); }
    void bcast(void* p, size_t b, int root, bool d) override { if (bc(user, p, (int64_t)b, root, d) != 0) throw std::runtime_error("bcast failed"); }
};
}
int pbtest_align_sharded(int rank, int world, pb200_allgather_cb ag, pb200_allreduce_cb ar, pb200_bcast_cb bc, int n,
                         const uint8_t* const* seqs, const int64_t* lens, const pb200_params* prm, pb200_result** out, int64_t* counters) {
    try {
        std::unique_ptr<pb200::SearchBackend> local(pb200_oracle::make_ref_backend());
        CbComm comm; comm.rank = rank; comm.world = world; comm.ag = ag; comm.ar = ar; comm.bc = bc; comm.user = nullptr;
        pb200::ShardedBackend sb(local.get(), pb200_oracle::as_staged(local.get()), &comm, true);
        pb200::Aligner a(n, seqs, lens, pb200::to_align_params(prm), &sb);
        a.set_threads(pb200::default_host_threads());
        a.set_pipeline(false);            // collectives inside the search: same call order on every rank
        bool ok = a.run();
        *out = pb200::make_result(a, (prm->flags & PB200_FLAG_UNALIGNED) != 0);
        if (counters) { counters[0] = sb.staged_windows; counters[1] = sb.sharded_small_windows; }
        return ok ? 0 : PB200_ERR_NO_MUMS;
    } catch (const std::exception& e) { pb200::g_last_error = e.what(); return PB200_ERR_INTERNAL; }
}
// one window through a CPU backend; outputs malloc'ed like pb200_search_windows
int pbtest_search_windows(int backend, int n, const uint8_t* const* seqs, const int64_t* lens, int ntasks, const pb200_window* tasks,
                          const int64_t* coords, int64_t** cand_off, int32_t** k, int32_t** lon, int32_t** sp, uint8_t** fwd) {
    pb200::SearchBackend* be = backend == 0 ? pb200_oracle::make_spec_backend() : pb200_oracle::make_ref_backend();
    be->set_genomes(n, seqs, lens);
    pb200::CandBatch cb;
    be->search((const pb200::WindowTask*)tasks, ntasks, coords, cb);
    cb.compact(ntasks);
    delete be;
    auto dup = [](const void* p, size_t bytes) { void* q = malloc(bytes ? bytes : 1); if (bytes) memcpy(q, p, bytes); return q; };
    *cand_off = (int64_t*)dup(cb.off.data(), cb.off.size() * 8);
    *k = (int32_t*)dup(cb.k.data(), cb.k.size() * 4);
    *lon = (int32_t*)dup(cb.lon.data(), cb.lon.size() * 4);
    *sp = (int32_t*)dup(cb.sp.data(), cb.sp.size() * 4);
    *fwd = (uint8_t*)dup(cb.fwd.data(), cb.fwd.size());
    return 0;
}
// parallel.h's literal_std_sort_by_first against the call it replaces: 0 = same permutation
int pbtest_literal_sort_check(const int64_t* keys, int64_t n, int threads) {
    std::vector<std::pair<int64_t, int>> a((size_t)n), b;
    for (int64_t i = 0; i < n; ++i) a[(size_t)i] = std::make_pair(keys[i], (int)i);
    b = a;
    std::sort(a.begin(), a.end(), [](const std::pair<int64_t, int>& x, const std::pair<int64_t, int>& y) { return x.first < y.first; });
    pb200::literal_std_sort_by_first(b.data(), (size_t)n, threads);
    for (int64_t i = 0; i < n; ++i) if (a[(size_t)i] != b[(size_t)i]) return 1;
    // the same with the tied keys given (ranges without equal keys are then sorted by other means)
    std::vector<int64_t> tk;
    for (int64_t i = 1; i < n; ++i) if (a[(size_t)i].first == a[(size_t)i - 1].first && (tk.empty() || tk.back() != a[(size_t)i].first)) tk.push_back(a[(size_t)i].first);
    for (int64_t i = 0; i < n; ++i) b[(size_t)i] = std::make_pair(keys[i], (int)i);
    pb200::literal_std_sort_by_first(b.data(), (size_t)n, threads, tk.data(), tk.size());
    for (int64_t i = 0; i < n; ++i) if (a[(size_t)i] != b[(size_t)i]) return 2;
    // ... and with the ascending list as a hint, its tied records in the OTHER order
    std::vector<std::pair<int64_t, int>> h(a);
    for (int64_t i = 1; i < n; ++i) if (h[(size_t)i].first == h[(size_t)i - 1].first) { std::swap(h[(size_t)i], h[(size_t)i - 1]); ++i; }
    for (int64_t i = 0; i < n; ++i) b[(size_t)i] = std::make_pair(keys[i], (int)i);
    pb200::literal_std_sort_by_first(b.data(), (size_t)n, threads, tk.data(), tk.size(), h.data());
    for (int64_t i = 0; i < n; ++i) if (a[(size_t)i] != b[(size_t)i]) return 3;
    return 0;
}
int pbtest_lrp(const uint8_t* R, int64_t n, int32_t* out) {
    std::vector<int32_t> v; pb200_oracle::spec_lrp(R, n, v); memcpy(out, v.data(), n * 4); return 0;
}
}
