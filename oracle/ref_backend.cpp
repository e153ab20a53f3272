// oracle/ref_backend.cpp - SearchBackend that calls the REAL reference csgmum (oracle/_ref/libcsgmum_ref.so:
// unmodified src/csgmum/csg.c + mum.c behind the shim of oracle/build_ref.py) the way Aligner::setMums1 does
// (src/parsnp.cpp:1570-1695).  TEST INFRASTRUCTURE ONLY (see oracle/build_ref.py); lets the CPU-only tests run the
// product's host orchestrator (parsnp_b200/csrc/host) against the reference search without a GPU.
#include <atomic>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include "../parsnp_b200/csrc/common.h"
#include "../parsnp_b200/csrc/host/sharded.h"

extern "C" {
void* ref_index_build(const char* text, long n, double factor);
void ref_index_free(void* h);
void ref_find_um(void* h, const char* q, long m, unsigned long* SP, int* pairUPEP);
void ref_intersect_um(void* h, int* masterUPEP, int* pairUPEP, int size, unsigned long* SP);
void ref_merge_master(int* masterUPEP, int* masterRCUPEP, int size, unsigned long* fwdSP, char* fwdflag, unsigned long* rcSP);
}

namespace pb200_oracle {

static inline char comp(char c) {
    switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return 'N'; }
}

// Calls of Find_UM the checker had to skip because the real one would have read past the end of the query buffer (see the
// comment at the first use).  tools/fuzz_host.py reads it: on such inputs the reference binary's own answer depends on what
// the heap holds behind that buffer, so a difference there is not a parity failure of the product.
static std::atomic<long> g_runoff_skips{0};
extern "C" long pbtest_runoff_skips(int reset) { return reset ? g_runoff_skips.exchange(0) : g_runoff_skips.load(); }

static bool shares_symbol(const uint8_t* a, int64_t na, const uint8_t* b, int64_t nb) {
    bool ina[256] = {false};
    for (int64_t i = 0; i < na; ++i) ina[a[i]] = true;
    for (int64_t i = 0; i < nb; ++i) if (ina[b[i]]) return true;
    g_runoff_skips.fetch_add(1);
    return false;
}

class RefBackend : public pb200::SearchBackend, public pb200::StagedWindowEngine {
public:
    void set_genomes(int n, const uint8_t* const* seq, const int64_t* len) override {
        n_ = n; seq_.assign(seq, seq + n); len_.assign(len, len + n);
    }
    void search(const pb200::WindowTask* tasks, int ntasks, const int64_t* coords, pb200::CandBatch& out) override {
        const int nq = n_ - 1;
        out.clear(); out.nq = nq; out.off.push_back(0);
        for (int t = 0; t < ntasks; ++t) {
            const int64_t n = tasks[t].ref_len;
            const int64_t* qs = coords + tasks[t].coord_off; const int64_t* ql = qs + nq;
            // (the reference sizes its node pool int(factor)*n = 2n with no bounds check; on windows of a few
            //  dozen bases construction can run past it - undefined behaviour the real binary survives by luck.  The checker
            //  gives small windows a roomier pool: the result is the same whenever the pool suffices.)
            void* ix = ref_index_build((const char*)seq_[0] + tasks[t].ref_start, (long)n, n < 512 ? 8.0 : 2.0);
            std::vector<int> Master(2 * n), MasterRC(2 * n), Pair(2 * n, 0), PairRC(2 * n, 0);
            for (int64_t i = 0; i < n; ++i) { Master[2 * i] = 0; Master[2 * i + 1] = (int)n; MasterRC[2 * i] = 0; MasterRC[2 * i + 1] = (int)n; }
            std::vector<std::vector<unsigned long>> MSP(nq, std::vector<unsigned long>(n, 0));
            std::vector<std::vector<char>> FW(nq, std::vector<char>(n, 1));
            std::vector<unsigned long> tmp(n);
            std::vector<char> rc;
            for (int q = 0; q < nq; ++q) {
                std::fill(tmp.begin(), tmp.end(), 0ul);
                const char* Q = (const char*)seq_[q + 1] + qs[q];
                rc.resize(ql[q]);
                for (int64_t i = 0; i < ql[q]; ++i) rc[i] = comp(Q[ql[q] - 1 - i]);
                // Find_UM first skips query symbols that start no edge at the root (src/csgmum/mum.c:193-198) with no bound: a
                // strand sharing NO symbol with the window (a poly-N window of a few dozen bases) makes it run off the end of
                // the query buffer.  The real binary survives that by luck (and reports no match); the checker skips the call.
                if (shares_symbol(seq_[0] + tasks[t].ref_start, n, (const uint8_t*)Q, ql[q]))
                    ref_find_um(ix, Q, (long)ql[q], MSP[q].data(), Pair.data());
                if (shares_symbol(seq_[0] + tasks[t].ref_start, n, (const uint8_t*)rc.data(), ql[q]))
                    ref_find_um(ix, rc.data(), (long)ql[q], tmp.data(), PairRC.data());
                ref_intersect_um(ix, Master.data(), Pair.data(), (int)n, MSP[q].data());
                ref_intersect_um(ix, MasterRC.data(), PairRC.data(), (int)n, tmp.data());
                ref_merge_master(Master.data(), MasterRC.data(), (int)n, MSP[q].data(), FW[q].data(), tmp.data());
            }
            int M_EP = 0;
            for (int64_t k = 0; k < n; ++k) {
                int UP = Master[2 * k], EP = Master[2 * k + 1];
                if (EP > M_EP && UP < EP && EP - k >= tasks[t].minsize) {
                    out.k.push_back((int32_t)k); out.lon.push_back((int32_t)(EP - k));
                    for (int q = 0; q < nq; ++q) { out.sp.push_back((int32_t)MSP[q][k]); out.fwd.push_back((uint8_t)FW[q][k]); }
                }
                M_EP = EP;
            }
            out.off.push_back((int64_t)out.k.size());
            ref_index_free(ix);
        }
    }
    // ---- StagedWindowEngine on host buffers: the same csgmum calls, restricted to a block of queries and started from a
    //      given Master (the reference's Master/MasterRC arrays are in/out, src/csgmum/mum.c:125-175)
    int64_t staged_threshold = 5000;
    bool wants_staged(const pb200::WindowTask& t, const int64_t*) override { return t.ref_len > staged_threshold; }
    bool buffers_on_device() const override { return false; }
    void window_begin(const pb200::WindowTask& t, const int64_t* coords, bool) override {
        w_ = t; wcoords_ = coords;
        if (ix_) ref_index_free(ix_);
        ix_ = ref_index_build((const char*)seq_[0] + t.ref_start, (long)t.ref_len, t.ref_len < 512 ? 8.0 : 2.0);
        up_.assign(t.ref_len, 0); ep_.assign(t.ref_len, (int32_t)t.ref_len); init_.clear();
    }
    void window_index_buffers(std::vector<std::pair<void*, size_t>>&) override {}     // every rank builds its own CSG
    int window_n() const override { return (int)w_.ref_len; }
    void window_scan(int q0, int q1) override { q0_ = q0; q1_ = q1; }
    void window_fold(bool init) override {
        const int64_t n = w_.ref_len; const int nq = n_ - 1;
        const int64_t* qs = wcoords_ + w_.coord_off; const int64_t* ql = qs + nq;
        if (init) { up_.assign(n, 0); ep_.assign(n, (int32_t)n); }
        std::vector<int> Master(2 * n), MasterRC(2 * n), Pair(2 * n, 0), PairRC(2 * n, 0);
        for (int64_t i = 0; i < n; ++i) { Master[2 * i] = MasterRC[2 * i] = up_[i]; Master[2 * i + 1] = MasterRC[2 * i + 1] = ep_[i]; }
        msp_.assign(q1_ - q0_, std::vector<unsigned long>(n, 0)); fw_.assign(q1_ - q0_, std::vector<char>(n, 1));
        std::vector<unsigned long> tmp(n); std::vector<char> rc;
        for (int q = q0_; q < q1_; ++q) {
            std::fill(tmp.begin(), tmp.end(), 0ul);
            const char* Q = (const char*)seq_[q + 1] + qs[q];
            rc.resize(ql[q]);
            for (int64_t i = 0; i < ql[q]; ++i) rc[i] = comp(Q[ql[q] - 1 - i]);
            if (shares_symbol(seq_[0] + w_.ref_start, w_.ref_len, (const uint8_t*)Q, ql[q]))
                ref_find_um(ix_, Q, (long)ql[q], msp_[q - q0_].data(), Pair.data());
            if (shares_symbol(seq_[0] + w_.ref_start, w_.ref_len, (const uint8_t*)rc.data(), ql[q]))
                ref_find_um(ix_, rc.data(), (long)ql[q], tmp.data(), PairRC.data());
            ref_intersect_um(ix_, Master.data(), Pair.data(), (int)n, msp_[q - q0_].data());
            ref_intersect_um(ix_, MasterRC.data(), PairRC.data(), (int)n, tmp.data());
            ref_merge_master(Master.data(), MasterRC.data(), (int)n, msp_[q - q0_].data(), fw_[q - q0_].data(), tmp.data());
        }
        for (int64_t i = 0; i < n; ++i) { up_[i] = Master[2 * i]; ep_[i] = Master[2 * i + 1]; }
    }
    int32_t* window_master_up() override { return up_.data(); }
    int32_t* window_master_ep() override { return ep_.data(); }
    int32_t* window_gather_buffer(size_t ints) override { gather_.resize(ints); return gather_.data(); }
    void window_apply_prefix(const int32_t* g, int, int rank) override {
        const int64_t n = w_.ref_len;
        init_.assign(n, (int32_t)n);
        for (int r = 0; r < rank; ++r) for (int64_t k = 0; k < n; ++k) init_[k] = std::min(init_[k], g[(size_t)r * n + k]);
        up_.assign(n, 0); ep_ = init_;
    }
    uint32_t window_emit() override {
        cand_.clear();
        int prev = 0;
        for (int64_t k = 0; k < w_.ref_len; ++k) {
            if (ep_[k] > prev && up_[k] < ep_[k] && ep_[k] - k >= w_.minsize) cand_.push_back((int32_t)k);
            prev = ep_[k];
        }
        return (uint32_t)cand_.size();
    }
    void window_pass2(std::vector<int32_t>& k, std::vector<int32_t>& lon, std::vector<int32_t>& sp, std::vector<uint8_t>& fwd) override {
        for (int32_t c : cand_) {
            k.push_back(c); lon.push_back(ep_[c] - c);
            for (int q = q0_; q < q1_; ++q) { sp.push_back((int32_t)msp_[q - q0_][c]); fwd.push_back((uint8_t)fw_[q - q0_][c]); }
        }
    }
private:
    int n_ = 0; std::vector<const uint8_t*> seq_; std::vector<int64_t> len_;
    pb200::WindowTask w_{}; const int64_t* wcoords_ = nullptr; void* ix_ = nullptr; int q0_ = 0, q1_ = 0;
    std::vector<int32_t> up_, ep_, init_, gather_, cand_;
    std::vector<std::vector<unsigned long>> msp_; std::vector<std::vector<char>> fw_;
};

// Record/replay: answers every window it has seen before from a process-wide table and sends the others to csgmum, in parallel
// (the windows are independent).  After one warm-up alignment the same input runs with a search that costs next to nothing,
// i.e. the host orchestrator alone can be timed on the CPU at the sizes where the GPU makes the search disappear
// (tools/host_bench.py).  The key is the window itself (reference interval, minimum length, query intervals), so the way the
// orchestrator batches its windows does not matter.
class RecordReplayBackend : public pb200::SearchBackend {
public:
    struct Entry { std::vector<int32_t> k, lon, sp; std::vector<uint8_t> fwd; };
    void set_genomes(int n, const uint8_t* const* seq, const int64_t* len) override {
        n_ = n; seq_.assign(seq, seq + n); len_.assign(len, len + n);
        // the table's keys are window coordinates: it belongs to ONE genome set (another set: start afresh)
        uint64_t h = 1469598103934665603ull;
        for (int g = 0; g < n; ++g) {
            h = (h ^ (uint64_t)len[g]) * 1099511628211ull;
            for (int64_t i = 0; i + 8 <= len[g]; i += 8) { uint64_t w; std::memcpy(&w, seq[g] + i, 8); h = (h ^ w) * 1099511628211ull; }
            for (int64_t i = len[g] & ~(int64_t)7; i < len[g]; ++i) h = (h ^ seq[g][i]) * 1099511628211ull;
        }
        std::lock_guard<std::mutex> lk(mu());
        if (h != genome_hash()) { table().clear(); genome_hash() = h; }
    }
    static uint64_t& genome_hash() { static uint64_t h = 0; return h; }
    void search(const pb200::WindowTask* tasks, int ntasks, const int64_t* coords, pb200::CandBatch& out) override {
        const int nq = n_ - 1;
        std::vector<std::string> keys((size_t)ntasks);
        std::vector<const Entry*> hit((size_t)ntasks, nullptr);
        std::vector<int> miss;
        {
            std::lock_guard<std::mutex> lk(mu());
            for (int t = 0; t < ntasks; ++t) {
                std::string& key = keys[t];
                key.assign((const char*)&tasks[t].ref_start, 8); key.append((const char*)&tasks[t].ref_len, 8);
                key.append((const char*)&tasks[t].minsize, 4); key.append((const char*)(coords + tasks[t].coord_off), (size_t)16 * nq);
                auto it = table().find(key);
                if (it != table().end()) hit[t] = it->second.get(); else miss.push_back(t);
            }
        }
        if (!miss.empty()) {
            std::vector<std::unique_ptr<Entry>> fresh(miss.size());
            std::atomic<size_t> next{0};
            auto work = [&]() {
                RefBackend rb;
                rb.set_genomes(n_, seq_.data(), len_.data());
                pb200::CandBatch cb;
                for (size_t i; (i = next.fetch_add(1)) < miss.size();) {
                    rb.search(&tasks[miss[i]], 1, coords, cb);
                    std::unique_ptr<Entry> e(new Entry);
                    e->k.assign(cb.k.begin(), cb.k.end()); e->lon.assign(cb.lon.begin(), cb.lon.end());
                    e->sp.assign(cb.sp.begin(), cb.sp.end()); e->fwd.assign(cb.fwd.begin(), cb.fwd.end());
                    fresh[i] = std::move(e);
                }
            };
            const unsigned T = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), miss.size());
            std::vector<std::thread> th;
            for (unsigned i = 1; i < T; ++i) th.emplace_back(work);
            work();
            for (auto& x : th) x.join();
            std::lock_guard<std::mutex> lk(mu());
            for (size_t i = 0; i < miss.size(); ++i) {
                auto ins = table().emplace(keys[miss[i]], std::move(fresh[i]));
                hit[miss[i]] = ins.first->second.get();
            }
            misses() += (long)miss.size();
        }
        out.clear(); out.nq = nq; out.off.assign((size_t)ntasks + 1, 0);
        for (int t = 0; t < ntasks; ++t) out.off[t + 1] = out.off[t] + (int64_t)hit[t]->k.size();
        const size_t tot = (size_t)out.off[ntasks];
        out.k.resize(tot); out.lon.resize(tot); out.sp.resize(tot * nq); out.fwd.resize(tot * nq);
        for (int t = 0; t < ntasks; ++t) {
            const Entry& e = *hit[t];
            const size_t b = (size_t)out.off[t], c = e.k.size();
            if (!c) continue;
            std::memcpy(out.k.data() + b, e.k.data(), c * 4);
            std::memcpy(out.lon.data() + b, e.lon.data(), c * 4);
            if (nq) { std::memcpy(out.sp.data() + b * nq, e.sp.data(), c * nq * 4); std::memcpy(out.fwd.data() + b * nq, e.fwd.data(), c * nq); }
        }
    }
    static std::mutex& mu() { static std::mutex m; return m; }
    static std::unordered_map<std::string, std::unique_ptr<Entry>>& table() { static std::unordered_map<std::string, std::unique_ptr<Entry>> t; return t; }
    static long& misses() { static long m = 0; return m; }
private:
    int n_ = 0;
    std::vector<const uint8_t*> seq_;
    std::vector<int64_t> len_;
};
extern "C" long pbtest_replay_misses(int reset) { long v = RecordReplayBackend::misses(); if (reset) RecordReplayBackend::misses() = 0; return v; }
extern "C" void pbtest_replay_clear() { std::lock_guard<std::mutex> lk(RecordReplayBackend::mu()); RecordReplayBackend::table().clear(); }

pb200::SearchBackend* make_ref_backend() { return new RefBackend(); }
pb200::SearchBackend* make_replay_backend() { return new RecordReplayBackend(); }
pb200::StagedWindowEngine* as_staged(pb200::SearchBackend* b) { return dynamic_cast<RefBackend*>(b); }

}  // namespace pb200_oracle
