#include "result.h"
#include "parallel.h"
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <thread>
#include <csignal>
#include <cstdio>
#include <execinfo.h>
#include <dlfcn.h>
#include <unistd.h>

namespace pb200 {
thread_local std::string g_last_error;

pb200_result* make_result(const Aligner& a, bool unaligned) {
    pb200_result* r = new pb200_result;
    const int n = a.n();
    r->n = n;
    const int64_t M = a.num_mums();
    r->m_length.resize(M); r->m_slength.resize(M);
    r->m_start.resize(M * n); r->m_fwd.resize(M * n);        // (m_end = start + length: filled when somebody asks for it)
    const long per = 4096;
    parallel_chunks(M > 32768 ? default_host_threads() : 1, ((long)M + per - 1) / per, [&](long c) {
        for (int64_t i = c * per; i < std::min<int64_t>(M, (c + 1) * per); ++i) {
            const MumRec& m = a.mum(i);
            r->m_length[i] = m.length; r->m_slength[i] = m.slength;
            const int64_t* s = a.mum_start(i); const uint8_t* f = a.mum_fwd(i);
            std::memcpy(&r->m_start[i * n], s, sizeof(int64_t) * (size_t)n);
            std::memcpy(&r->m_fwd[i * n], f, (size_t)n);
        }
    });
    r->c_mum_off.push_back(0);
    for (const ClusterRec& c : a.clusters()) {
        for (int mi : c.mums) r->c_mum_idx.push_back(mi);
        r->c_mum_off.push_back((int64_t)r->c_mum_idx.size());
        r->c_type.push_back(c.type); r->c_nmums.push_back(c.type == 1 ? (int64_t)c.mums.size() : 2); r->c_length.push_back(c.length);
        r->c_start.insert(r->c_start.end(), c.start.begin(), c.start.end());
        r->c_end.insert(r->c_end.end(), c.end.begin(), c.end.end());
    }
    for (auto& p : a.window_trace()) { r->trace.push_back(p.first); r->trace.push_back(p.second); }
    if (unaligned) a.unaligned_regions(r->u_genome, r->u_start, r->u_end);
    const AlignStats& s = a.stats();
    r->stats = { (double)s.anchors, (double)s.regions_searched, (double)s.spec_regions, (double)s.replay_misses,
                 (double)s.spec_levels, (double)s.windows_searched, (double)s.candidates, (double)s.slow_queue_iters,
                 s.t_anchor_search, s.t_anchor_host, s.t_spec_search, s.t_spec_host, s.t_replay, s.t_replay_search,
                 s.t_lcb, s.t_total, (double)s.host_threads, s.t_search_prep, s.t_search_backend, s.t_search_cache, s.t_replay_wait,
                 (double)s.spec_slices, (double)s.mums_filtered, (double)s.clusters_filtered,
                 (double)s.replay_tasks, (double)s.replay_foreign_reads, (double)s.replay_foreign_writes, (double)s.replay_restarts,
                 (double)s.replay_fallback, (double)s.replay_workers, s.t_replay_merge, (double)s.spec_deferred,
                 (double)s.replay_gaps, (double)s.replay_final_gaps, (double)s.replay_final_mums };
    return r;
}

// PB200_BACKTRACE=1: a fatal signal inside the library prints its frames as offsets into the shared object (addr2line -e
// libparsnp_b200.so <offset>) before the default action - a debugging aid for boxes without a debugger
namespace {
void fatal_signal_handler(int sig) {
    void* frames[64];
    const int nf = backtrace(frames, 64);
    Dl_info me;
    const char* base = dladdr((void*)&fatal_signal_handler, &me) ? (const char*)me.dli_fbase : nullptr;
    char buf[256];
    int len = snprintf(buf, sizeof buf, "[pb200] fatal signal %d, library %s loaded at %p; frames:\n", sig, me.dli_fname ? me.dli_fname : "?", (const void*)base);
    if (write(2, buf, (size_t)len) < 0) {}
    for (int i = 0; i < nf; ++i) {
        Dl_info di;
        const bool in = dladdr(frames[i], &di) && di.dli_fbase;
        len = snprintf(buf, sizeof buf, "  #%d %s +0x%lx\n", i, in && di.dli_fname ? di.dli_fname : "?", in ? (unsigned long)((const char*)frames[i] - (const char*)di.dli_fbase) : (unsigned long)(uintptr_t)frames[i]);
        if (write(2, buf, (size_t)len) < 0) {}
    }
    signal(sig, SIG_DFL);
    raise(sig);
}
}  // namespace
void install_backtrace_handler() {
    static bool done = false;
    if (done || !getenv("PB200_BACKTRACE")) return;
    done = true;
    signal(SIGSEGV, fatal_signal_handler);
    signal(SIGABRT, fatal_signal_handler);
    signal(SIGBUS, fatal_signal_handler);
}

int default_host_threads() {
    if (const char* e = getenv("PB200_HOST_THREADS")) { int v = atoi(e); return v < 1 ? 1 : v; }
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    int ranks = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) { int v = atoi(e); if (v > 0) ranks = v; }
    int t = hw / ranks;
    if (t > 32) t = 32;
    return t < 1 ? 1 : t;
}

AlignParams to_align_params(const pb200_params* p) {
    AlignParams a;
    a.c = p->c; a.d = p->d; a.q = p->q; a.p = p->p; a.diag_diff = p->diagdiff;
    if (a.diag_diff < 0.0 || a.diag_diff > 10000000) a.diag_diff = 1.0;      // src/parsnp.cpp:2873-2876
    a.random = p->filter; a.anchors_only = p->anchors_only != 0;
    if (p->anchors) a.anchors = p->anchors;
    if (p->mums) a.mums = p->mums;
    return a;
}
}  // namespace pb200

extern "C" {
void pb200_params_default(pb200_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->c = 21; p->d = 300; p->q = 30; p->p = 15000000; p->diagdiff = 0.12f; p->filter = 1; p->anchors_only = 0;
    p->anchors = "1.1*(Log(S))"; p->mums = "1.1*(Log(S))";
}
const char* pb200_last_error(void) { return pb200::g_last_error.c_str(); }
int pb200_result_n(const pb200_result* r) { return r->n; }
int64_t pb200_result_num_mums(const pb200_result* r) { return (int64_t)r->m_length.size(); }
namespace {
// end[i][k] = start[i][k] + length[i] (TMum::end, src/TMum.cpp:13-72) into `dst` (M * n entries)
void fill_mum_ends(const pb200_result* r, int64_t* dst) {
    const int64_t M = (int64_t)r->m_length.size();
    const int n = r->n;
    const long per = 4096;
    pb200::parallel_chunks(M > 32768 ? pb200::default_host_threads() : 1, ((long)M + per - 1) / per, [&](long c) {
        for (int64_t i = c * per; i < std::min<int64_t>(M, (c + 1) * per); ++i)
            for (int k = 0; k < n; ++k) dst[i * n + k] = r->m_start[i * n + k] + r->m_length[i];
    });
}
}  // namespace
int pb200_result_mums(const pb200_result* r, int64_t* length, int64_t* slength, int64_t* start, int64_t* end, uint8_t* fwd) {
    if (length) std::memcpy(length, r->m_length.data(), r->m_length.size() * 8);
    if (slength) std::memcpy(slength, r->m_slength.data(), r->m_slength.size() * 8);
    if (start) std::memcpy(start, r->m_start.data(), r->m_start.size() * 8);
    if (end) fill_mum_ends(r, end);
    if (fwd) std::memcpy(fwd, r->m_fwd.data(), r->m_fwd.size());
    return 0;
}
int pb200_result_mums_view(const pb200_result* r, const int64_t** length, const int64_t** slength, const int64_t** start, const int64_t** end,
                           const uint8_t** fwd) {
    if (length) *length = r->m_length.data();
    if (slength) *slength = r->m_slength.data();
    if (start) *start = r->m_start.data();
    if (end) {                                   // (materialised on the first request; callers that only need starts and lengths pass NULL)
        pb200_result* w = const_cast<pb200_result*>(r);
        if (w->m_end.size() != w->m_start.size()) { w->m_end.resize(w->m_start.size()); fill_mum_ends(r, w->m_end.data()); }
        *end = r->m_end.data();
    }
    if (fwd) *fwd = r->m_fwd.data();
    return 0;
}
int64_t pb200_result_num_clusters(const pb200_result* r) { return (int64_t)r->c_type.size(); }
int pb200_result_clusters(const pb200_result* r, int32_t* type, int64_t* nmums, int64_t* length, int64_t* start, int64_t* end) {
    if (type) std::memcpy(type, r->c_type.data(), r->c_type.size() * 4);
    if (nmums) std::memcpy(nmums, r->c_nmums.data(), r->c_nmums.size() * 8);
    if (length) std::memcpy(length, r->c_length.data(), r->c_length.size() * 8);
    if (start) std::memcpy(start, r->c_start.data(), r->c_start.size() * 8);
    if (end) std::memcpy(end, r->c_end.data(), r->c_end.size() * 8);
    return 0;
}
int pb200_result_cluster_mums(const pb200_result* r, int64_t* off, int64_t* idx) {
    if (off) std::memcpy(off, r->c_mum_off.data(), r->c_mum_off.size() * 8);
    if (idx) std::memcpy(idx, r->c_mum_idx.data(), r->c_mum_idx.size() * 8);
    return (int)r->c_mum_idx.size();
}
int64_t pb200_result_unaligned(const pb200_result* r, int32_t* genome, int64_t* start, int64_t* end) {
    if (genome) std::memcpy(genome, r->u_genome.data(), r->u_genome.size() * 4);
    if (start) std::memcpy(start, r->u_start.data(), r->u_start.size() * 8);
    if (end) std::memcpy(end, r->u_end.data(), r->u_end.size() * 8);
    return (int64_t)r->u_genome.size();
}
int64_t pb200_result_num_trace(const pb200_result* r) { return (int64_t)r->trace.size() / 2; }
int pb200_result_trace(const pb200_result* r, int64_t* pairs) { std::memcpy(pairs, r->trace.data(), r->trace.size() * 8); return 0; }
int pb200_result_stats(const pb200_result* r, double* values, int cap) {
    int k = 0;
    for (; k < cap && k < (int)r->stats.size(); ++k) values[k] = r->stats[k];
    return k;
}
const char* pb200_stats_names(void) {
    return "anchors,regions_searched,spec_regions,replay_misses,spec_levels,windows_searched,candidates,slow_queue_iters,"
           "t_anchor_search,t_anchor_host,t_spec_search,t_spec_host,t_replay,t_replay_search,t_lcb,t_total,host_threads,t_search_prep,t_search_backend,t_search_cache,t_replay_wait,spec_slices,mums_filtered,clusters_filtered,"
           "replay_tasks,replay_foreign_reads,replay_foreign_writes,replay_restarts,replay_fallback,replay_workers,t_replay_merge,spec_deferred,"
           "replay_gaps,replay_final_gaps,replay_final_mums";
}
void pb200_result_free(pb200_result* r) { delete r; }
int pb200_minsize(const char* expr, int64_t slength) {
    // (the reference's calculator exits the process on a division by zero, src/Converter.cpp:252; here: -1 + pb200_last_error)
    try { return pb200::MinSizeExpr(expr)(slength); }
    catch (const std::exception& e) { pb200::g_last_error = e.what(); return -1; }
}
void pb200_free_buffer(void* p) { free(p); }
}
