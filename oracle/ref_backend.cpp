// oracle/ref_backend.cpp - SearchBackend that calls the REAL reference csgmum (oracle/_ref/libcsgmum_ref.so:
// unmodified src/csgmum/csg.c + mum.c behind the shim of oracle/build_ref.py) the way Aligner::setMums1 does
// (src/parsnp.cpp:1570-1695).  TEST INFRASTRUCTURE ONLY (see oracle/build_ref.py); lets the CPU-only tests run the
// product's host orchestrator (parsnp_b200/csrc/host) against the reference search without a GPU.
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include "../parsnp_b200/csrc/common.h"

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

class RefBackend : public pb200::SearchBackend {
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
            void* ix = ref_index_build((const char*)seq_[0] + tasks[t].ref_start, (long)n, 2.0);
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
                ref_find_um(ix, Q, (long)ql[q], MSP[q].data(), Pair.data());
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
private:
    int n_ = 0; std::vector<const uint8_t*> seq_; std::vector<int64_t> len_;
};

pb200::SearchBackend* make_ref_backend() { return new RefBackend(); }

}  // namespace pb200_oracle
