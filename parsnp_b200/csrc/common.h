// parsnp_b200 - shared plain-data types between the C++ host orchestrator and the CUDA search engine.
//
// Vocabulary follows the reference (marbl/parsnp): a *region* is a TRegion (src/LCR.hh) = one half-open
// interval per genome; a *window* is one pass of the reference-window loop of Aligner::setMums1
// (src/parsnp.cpp:1519-1547); a *candidate* is one `Mum{DSP,LON,forward}` emitted by the loop at
// src/parsnp.cpp:1633-1695; a MUM is an accepted TMum (src/TMum.hh); an LCB is a Cluster (src/LCB.hh).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <memory>
#include <utility>
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

// vector whose resize() leaves new elements uninitialised (the candidate arrays are filled by device-to-host copies)
template <class T>
struct default_init_allocator : std::allocator<T> {
    template <class U> struct rebind { using other = default_init_allocator<U>; };
    template <class U> void construct(U* p) noexcept { ::new ((void*)p) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};
template <class T> using pod_vector = std::vector<T, default_init_allocator<T>>;

// Candidates of a batch of windows, SoA. Candidate c of window w lives at index off[w]+c (increasing k), c < count(w).
// A backend may return the windows' blocks in any order (`cnt` filled); with `cnt` empty the blocks are contiguous in
// window order and off has ntasks+1 entries.
struct CandBatch {
    int nq = 0;                      // number of query genomes (n-1)
    std::vector<int64_t> off;        // [ntasks] (+1 when contiguous)
    std::vector<int32_t> cnt;        // [ntasks] or empty
    pod_vector<int32_t> k;           // reference start inside the window (0-based)
    pod_vector<int32_t> lon;         // LON
    pod_vector<int32_t> sp;          // [ncand * nq] start inside the query region, in the winning strand's coordinates
    pod_vector<uint8_t> fwd;         // [ncand * nq] 1 = forward strand won
    // readers go through these: the batch's own arrays, or (device discovery) views into the engine's pinned staging memory
    const int32_t* vk = nullptr; const int32_t* vlon = nullptr; const int32_t* vsp = nullptr; const uint8_t* vfwd = nullptr;
    size_t vcount = 0;
    const int32_t* K() const { return vk ? vk : k.data(); }
    const int32_t* LON() const { return vlon ? vlon : lon.data(); }
    const int32_t* SP() const { return vsp ? vsp : sp.data(); }
    const uint8_t* FWD() const { return vfwd ? vfwd : fwd.data(); }
    size_t ncands() const { return vk ? vcount : k.size(); }
    void clear() { off.clear(); cnt.clear(); k.clear(); lon.clear(); sp.clear(); fwd.clear(); vk = vlon = vsp = nullptr; vfwd = nullptr; vcount = 0; }
    int32_t count(int t) const { return cnt.empty() ? (int32_t)(off[t + 1] - off[t]) : cnt[t]; }
    // rewrite into window order: off[ntasks+1] increasing, cnt empty
    void compact(int ntasks) {
        if (cnt.empty()) return;
        std::vector<int64_t> noff((size_t)ntasks + 1, 0);
        for (int t = 0; t < ntasks; ++t) noff[t + 1] = noff[t] + cnt[t];
        const int64_t tot = noff[ntasks];
        pod_vector<int32_t> nk((size_t)tot), nl((size_t)tot), ns((size_t)tot * nq);
        pod_vector<uint8_t> nf((size_t)tot * nq);
        for (int t = 0; t < ntasks; ++t) {
            const size_t c = (size_t)cnt[t], a = (size_t)off[t], b = (size_t)noff[t];
            if (!c) continue;
            std::memcpy(nk.data() + b, k.data() + a, c * 4);
            std::memcpy(nl.data() + b, lon.data() + a, c * 4);
            if (nq) {
                std::memcpy(ns.data() + b * nq, sp.data() + a * nq, c * nq * 4);
                std::memcpy(nf.data() + b * nq, fwd.data() + a * nq, c * nq);
            }
        }
        k.swap(nk); lon.swap(nl); sp.swap(ns); fwd.swap(nf);
        off.swap(noff);
        cnt.clear();
    }
};

#if defined(__CUDACC__)
#define PB_HD __host__ __device__
#else
#define PB_HD
#endif
// hash of a region's 2n coordinates (start[n], end[n]): start and end of the first and the last genome - regions equal there and
// different elsewhere are rare, and every table compares all coordinates on a hit.  Shared by the host's candidate cache and the
// engine, which delivers the hashes of the regions it discovered.
PB_HD inline uint64_t region_coords_hash(const int64_t* p, int count) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    const int half = count / 2;
    const int idx[4] = {0, half - 1, half, count - 1};
    for (int t = 0; t < 4; ++t) {
        h ^= (uint64_t)p[idx[t]] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h *= 0xff51afd7ed558ccdull;
        h ^= h >> 29;
    }
    return h;
}
// one searched reference window of a region in a candidate cache
struct WindowRec { int64_t ref_start, ref_len; int64_t cand_off; int32_t ncand; int32_t chunk; };

// Device-resident discovery of the recursion (cuda/recursion.cuh): every region the engine searched while following the
// accept / trim / determineRegion rules on a scratch copy of mumlayout, with the candidates of its (single) window.
struct RecursionRequest {
    int n = 0;                           // genomes
    const int64_t* coords = nullptr;     // initial regions: start[n] then end[n] each
    int nregions = 0;
    bool upload_layout = true;           // false: keep the scratch layout of the previous call (next slice of the same alignment)
    bool resume = false;                 // true: collect the run that SearchBackend::anchor_stage started (nothing else is read)
    const uint64_t* const* layout = nullptr;   // mumlayout after the anchors: per genome its words ...
    const int64_t* layout_words = nullptr;     // ... and their number (len + 1 bits incl. the sentinel)
    int q = 30;                          // ini [LCB] q
    int64_t p = 15000000;                // ini [LCB] p
    const int32_t* minsize_tab = nullptr;      // minsize(slength) of the ini `mums` expression for slength < minsize_n
    int minsize_n = 0;
};
// All arrays are VIEWS into memory of the backend (pinned staging), valid until its next discover_recursion call; regions are in
// ascending start[0] order, candidates grouped accordingly.
constexpr uint32_t REC_ORDER_MASK = 15, REC_SECOND = 16;
struct RecursionResult {
    size_t nregions = 0, ncands = 0;
    const int64_t* coords = nullptr;     // [nregions * 2n]: start[n] then end[n]
    const int64_t* slen = nullptr;       // TRegion::slength
    const uint64_t* hashes = nullptr;    // region_coords_hash of every region
    const WindowRec* wins = nullptr;     // the region's single window; ncand < 0: not searched (left to the caller)
    const int32_t* k = nullptr;          // candidates as in CandBatch (sp / fwd: nq per candidate)
    const int32_t* lon = nullptr;
    const int32_t* sp = nullptr;
    const uint8_t* fwd = nullptr;
    // the engine's own accept decisions (it followed setMums1's loop D on a scratch layout), for callers that can tell where they
    // cannot depend on the order (host/replay.cpp takes them as final for such gaps): per region flags - a bit of
    // REC_ORDER_MASK set = the region's pass may depend on the order (a reverse-strand candidate reached the trim loop, the
    // second region of a gap pair accepted something, its accepted MUMs are not collinear, it was searched outside its pair's
    // order), REC_SECOND = searched as the second region of a pair - and the sorted position of its parent region (-1:
    // initial); per candidate the trim shift (-1: not accepted) and the accepted length; and the accepted MUMs that were
    // written OUTSIDE their region (genome, start, length) - nfw > fw_cap: the list is incomplete.  flags == nullptr: not provided
    const uint32_t* flags = nullptr;
    const int32_t* parent = nullptr;
    const int32_t* acc_shift = nullptr;
    const int32_t* acc_len = nullptr;
    const int32_t* fw = nullptr;
    size_t nfw = 0, fw_cap = 0;
    int64_t levels = 0, deferred = 0, dropped = 0, searched = 0;
};

// The anchor stage on the device (cuda/anchors.cuh): candidates of the whole-genome region -> accepted anchors, mumlayout and
// the regions between the anchors, with the recursion (above) started from them right away.
struct AnchorRequest {
    int n = 0;
    const WindowTask* tasks = nullptr;   // the reference windows of the whole-genome region (src/parsnp.cpp:1519-1547)
    int ntasks = 0;
    const int64_t* coords = nullptr;     // the task coordinate pool: q_start[n-1] = 0, q_len[n-1] = genome lengths
    const int64_t* layout_words = nullptr;     // words per genome row of mumlayout (len + 1 bits incl. the sentinel)
    int q = 30;
    int64_t p = 15000000;
    const int32_t* minsize_tab = nullptr;      // for the recursion that follows (RecursionRequest)
    int minsize_n = 0;
    bool follow_recursion = true;        // false: ini anchorsonly
};
struct AnchorResult {
    // status 1: everything below is valid (views into the backend's pinned memory until its next call), the recursion is
    //           running on the device and discover_recursion(resume) collects it
    // status 2: the candidates overlap or are not collinear: `cands` holds them in search() format, the caller accepts them
    int status = 0;
    size_t ncand = 0, nanchors = 0, nregions = 0;
    const int32_t* a_start = nullptr;    // [nanchors * n]
    const int32_t* a_lon = nullptr;      // [nanchors]
    const uint8_t* a_fwd = nullptr;      // [nanchors * n]
    const int32_t* r_coords = nullptr;   // [nregions * 2n]: start[n] then LENGTH[n], in push order
    const int32_t* r_lo = nullptr;       // [nregions * n] (optional): the set bit of mumlayout that bounds the region on the left, per genome (0: none)
    const uint64_t* layout = nullptr;    // all rows of mumlayout after the anchors, row g at layout_off[g] words
    std::vector<int64_t> layout_off;
    CandBatch cands;
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
    // optional: follow the recursion on the device; false = not supported for this input (the host then discovers the regions
    // level by level through search())
    virtual bool discover_recursion(const RecursionRequest&, RecursionResult&) { return false; }
    // optional: the anchor stage (search of the whole-genome region + accept + regions between the anchors) on the device.
    // Returns AnchorResult::status (0 = not supported: the caller goes through search())
    virtual int anchor_stage(const AnchorRequest&, AnchorResult&) { return 0; }
};

}  // namespace pb200
