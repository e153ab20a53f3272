// oracle/mumspec.cpp - CPU restatement (specification) of the reference's multi-MUM window search.
//
// TEST INFRASTRUCTURE ONLY: linked into oracle/_ref/libpb200_hosttest.so, never into the product library.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
//
// It restates, as plain brute force over diagonals (O(n*m) per strand), what the reference computes for one
// reference window of Aligner::setMums1 (src/parsnp.cpp:1570-1695) through csgmum:
//   A1  u[l]  = l + longest prefix of R[l..) that occurs elsewhere in R          (leaf label, src/csgmum/mum.c:219-223)
//   A2  events = maximal exact matches (j,l,L) with l+L > u[l]                    (Find_UM, src/csgmum/mum.c:177-250)
//       per reference start l: EP = largest end, UP = max(u[l], 2nd largest end) (Test_UM, src/csgmum/mum.c:27-45)
//   A3  propagation along k = inclusive scan of (floor, top1, top2)               (Intersect_UM, src/csgmum/mum.c:125-175)
//   A4  fold over queries in ini order, strand merge with ties -> reverse         (Merge_Master, src/csgmum/mum.c:92-123)
//   A5  emission: EP[k] > EP[k-1] && UP[k] < EP[k] && EP[k]-k >= minsize           (src/parsnp.cpp:1633-1695)
// The restatement is pinned against the real csg.c/mum.c (oracle/_ref/libcsgmum_ref.so) and against candidate
// dumps of the real parsnp_core (hook H2 of oracle/build_ref.py) in tests/test_oracle_spec.py.
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../parsnp_b200/csrc/common.h"

namespace {

inline uint8_t comp(uint8_t c) {
    switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return 'N'; }
}

struct PerK { int32_t UP, EP; int64_t SP; };

// per-strand dense (UP', EP', SP') after Find_UM + Intersect_UM
void strand_scan(const uint8_t* R, int64_t n, const std::vector<int32_t>& lrp, const uint8_t* Q, int64_t m,
                 std::vector<PerK>& out) {
    std::vector<PerK> pair((size_t)n, PerK{0, 0, 0});
    // A2: maximal matches along every diagonal d = l - j ; events visited in increasing j (order only matters for ties)
    struct Ev { int64_t j, l, L; };
    std::vector<Ev> evs;
    for (int64_t d = -(m - 1); d <= n - 1; ++d) {
        int64_t j = d < 0 ? -d : 0, l = d < 0 ? 0 : d;
        int64_t run = 0;
        for (; j < m && l < n; ++j, ++l) {
            if (Q[j] == R[l]) { ++run; }
            else { if (run) evs.push_back(Ev{j - run, l - run, run}); run = 0; }
        }
        if (run) evs.push_back(Ev{j - run, l - run, run});
    }
    std::sort(evs.begin(), evs.end(), [](const Ev& a, const Ev& b) { return a.j < b.j || (a.j == b.j && a.l < b.l); });
    for (const Ev& e : evs) {
        if (e.L <= lrp[e.l]) continue;                      // not unique in R  <=> l+L <= u[l]
        PerK& p = pair[e.l];
        int32_t MMP = (int32_t)(e.l + e.L);
        if (p.EP == 0) { p.EP = MMP; p.UP = (int32_t)(e.l + lrp[e.l]); p.SP = e.j; continue; }
        if (MMP > p.UP) {
            if (MMP > p.EP) { p.UP = p.EP; p.EP = MMP; p.SP = e.j; }
            else p.UP = MMP;
        }
    }
    // A3 as the associative (fl, t1, t2) scan, evaluated sequentially
    out.assign((size_t)n, PerK{0, 0, 0});
    int32_t fl = 0, t1 = 0, t2 = 0; int64_t diag = 0;   // diag = SP - l of the event holding t1
    for (int64_t k = 0; k < n; ++k) {
        if (pair[k].EP != 0 || pair[k].UP != 0) {
            fl = std::max(fl, pair[k].UP);
            int32_t e = pair[k].EP;
            if (e >= t1) { t2 = t1; t1 = e; diag = pair[k].SP - k; }
            else if (e > t2) t2 = e;
        }
        out[k].UP = std::max(fl, t2);
        out[k].EP = std::max(t1, fl);
        out[k].SP = k + diag;
    }
}

}  // namespace

namespace pb200_oracle {

// brute force: longest repeated prefix per position (A1)
void spec_lrp(const uint8_t* R, int64_t n, std::vector<int32_t>& lrp) {
    lrp.assign((size_t)n, 0);
    for (int64_t d = 1; d < n; ++d) {
        int32_t run = 0;
        for (int64_t l = n - 1 - d; l >= 0; --l) {
            if (R[l] == R[l + d]) { ++run; if (run > lrp[l]) lrp[l] = run; if (run > lrp[l + d]) lrp[l + d] = run; }
            else run = 0;
        }
    }
}

struct SpecCand { int32_t k, lon; std::vector<int32_t> sp; std::vector<uint8_t> fwd; };

void spec_window(const uint8_t* R, int64_t n, int nq, const uint8_t* const* Q, const int64_t* m, int minsize,
                 std::vector<SpecCand>& out) {
    std::vector<int32_t> lrp;
    spec_lrp(R, n, lrp);
    std::vector<int32_t> MUP((size_t)n, 0), MEP((size_t)n, (int32_t)n);
    std::vector<std::vector<int64_t>> SPq((size_t)nq, std::vector<int64_t>((size_t)n, 0));
    std::vector<std::vector<uint8_t>> FWq((size_t)nq, std::vector<uint8_t>((size_t)n, 0));
    std::vector<PerK> F, C;
    std::vector<uint8_t> rc;
    for (int q = 0; q < nq; ++q) {
        strand_scan(R, n, lrp, Q[q], m[q], F);
        rc.resize((size_t)m[q]);
        for (int64_t t = 0; t < m[q]; ++t) rc[(size_t)t] = comp(Q[q][m[q] - 1 - t]);
        strand_scan(R, n, lrp, rc.data(), m[q], C);
        for (int64_t k = 0; k < n; ++k) {                      // A4
            int32_t fe = std::min(MEP[k], F[k].EP), ce = std::min(MEP[k], C[k].EP);
            if (fe > ce) { MUP[k] = std::max(MUP[k], F[k].UP); MEP[k] = fe; SPq[q][k] = F[k].SP; FWq[q][k] = 1; }
            else         { MUP[k] = std::max(MUP[k], C[k].UP); MEP[k] = ce; SPq[q][k] = C[k].SP; FWq[q][k] = 0; }
        }
    }
    out.clear();
    int32_t prevEP = 0;
    for (int64_t k = 0; k < n; ++k) {                           // A5
        if (MEP[k] > prevEP && MUP[k] < MEP[k] && MEP[k] - k >= minsize) {
            SpecCand c; c.k = (int32_t)k; c.lon = (int32_t)(MEP[k] - k);
            for (int q = 0; q < nq; ++q) { c.sp.push_back((int32_t)SPq[q][k]); c.fwd.push_back(FWq[q][k]); }
            out.push_back(c);
        }
        prevEP = MEP[k];
    }
}

// SearchBackend over the specification (slow; tiny inputs only)
class SpecBackend : public pb200::SearchBackend {
public:
    void set_genomes(int n, const uint8_t* const* seq, const int64_t* len) override {
        n_ = n; seq_.assign(seq, seq + n); len_.assign(len, len + n);
    }
    void search(const pb200::WindowTask* tasks, int ntasks, const int64_t* coords, pb200::CandBatch& out) override {
        const int nq = n_ - 1;
        out.clear(); out.nq = nq; out.off.push_back(0);
        std::vector<const uint8_t*> Q((size_t)nq); std::vector<int64_t> m((size_t)nq);
        std::vector<SpecCand> cands;
        for (int t = 0; t < ntasks; ++t) {
            const int64_t* qs = coords + tasks[t].coord_off; const int64_t* ql = qs + nq;
            for (int q = 0; q < nq; ++q) { Q[q] = seq_[q + 1] + qs[q]; m[q] = ql[q]; }
            spec_window(seq_[0] + tasks[t].ref_start, tasks[t].ref_len, nq, Q.data(), m.data(), tasks[t].minsize, cands);
            for (auto& c : cands) {
                out.k.push_back(c.k); out.lon.push_back(c.lon);
                out.sp.insert(out.sp.end(), c.sp.begin(), c.sp.end());
                out.fwd.insert(out.fwd.end(), c.fwd.begin(), c.fwd.end());
            }
            out.off.push_back((int64_t)out.k.size());
        }
    }
private:
    int n_ = 0; std::vector<const uint8_t*> seq_; std::vector<int64_t> len_;
};

pb200::SearchBackend* make_spec_backend() { return new SpecBackend(); }

}  // namespace pb200_oracle
