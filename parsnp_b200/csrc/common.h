// parsnp_b200 - shared plain-data types between the C++ host orchestrator and the CUDA search engine.
//
// Vocabulary follows the reference (marbl/parsnp): a *region* is a TRegion (src/LCR.hh) = one half-open
// interval per genome; a *window* is one pass of the reference-window loop of Aligner::setMums1
// (src/parsnp.cpp:1519-1547); a *candidate* is one `Mum{DSP,LON,forward}` emitted by the loop at
// src/parsnp.cpp:1633-1695; a MUM is an accepted TMum (src/TMum.hh); an LCB is a Cluster (src/LCB.hh).
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

namespace pb200 {

// One reference window of one region: everything the search (index build + scan + fold + emission) depends on.
// The search is a pure function of these coordinates and of the genome texts (it never reads mumlayout).
struct WindowTask {
    int64_t ref_start;       // rs[0].ini_region (src/parsnp.cpp:1546)
    int64_t ref_len;         // rs[0].len_region (src/parsnp.cpp:1545)
    int64_t coord_off;       // offset into the task coordinate pool: q_start[n-1], q_len[n-1] (genomes 1..n-1)
    int32_t minsize;         // src/parsnp.cpp:1502-1514
    int32_t pad;
};

// Candidates of a batch of windows, SoA. Candidate c of window w lives at index off[w]+c (increasing k).
struct CandBatch {
    int nq = 0;                      // number of query genomes (n-1)
    std::vector<int64_t> off;        // [ntasks+1]
    std::vector<int32_t> k;          // reference start inside the window (0-based)
    std::vector<int32_t> lon;        // LON
    std::vector<int32_t> sp;         // [ncand * nq] start inside the query region, in the winning strand's coordinates
    std::vector<uint8_t> fwd;        // [ncand * nq] 1 = forward strand won
    void clear() { off.clear(); k.clear(); lon.clear(); sp.clear(); fwd.clear(); }
};

// Search engine interface. The product implementation is CUDA-only (cuda/engine.cu); tests may link the
// CPU specification from oracle/ behind the same interface to exercise the host logic without a GPU.
class SearchBackend {
public:
    virtual ~SearchBackend() {}
    // genomes are given once (ASCII A,C,G,T,N only; see ingest rules src/parsnp.cpp:2999-3133)
    virtual void set_genomes(int n, const uint8_t* const* seq, const int64_t* len) = 0;
    // coords: pool referenced by WindowTask::coord_off
    virtual void search(const WindowTask* tasks, int ntasks, const int64_t* coords, CandBatch& out) = 0;
};

}  // namespace pb200
