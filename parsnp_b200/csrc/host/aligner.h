// Host orchestrator of the MUM + LCB path: anchors -> recursive inter-anchor search -> LCB chaining.
//
// This is a from-scratch C++ restatement of the *observable behaviour* of the reference's Aligner hot methods
// (src/parsnp.cpp: setInitialClusters 2121-2174, setMums1 1484-1862 [loop D: validation/trim/accept],
// trim 1399-1477, determineRegion 1199-1290, doWork 173-317, filterRandom1 327-425, setFinalClusters 2563-2719,
// filterRandomClustersSimple1 433-497, setInterClusterRegions 2389-2460) with different data structures:
//   * mumlayout is a packed 64-bit bitmap with word-level scans instead of vector<bool> bit loops;
//   * MUMs / regions live in flat SoA pools instead of vector<vector<long>> objects;
//   * the index build + scan + fold + emission of setMums1 (its calls into csgmum) is delegated to a
//     SearchBackend (the CUDA engine), batched over many regions at once;
//   * the recursion runs as  (1) a speculative, level-synchronous pass (slice after slice of the initial regions, on its
//     own thread) that only exists to discover which regions will be searched and to batch them onto the GPU, beside
//     (2) an exact sequential replay in the reference's own order (pop smallest start[0], push children, sort, drop
//     adjacent duplicates) that consumes the finished slices, looks the candidates up by region coordinates and asks
//     the GPU for any region the speculation did not predict.
//     The search is a pure function of the region coordinates, so pass (2) is exact by construction.
#pragma once
#include <cstdint>
#include <climits>
#include <string>
#include <vector>
#include <unordered_map>
#include <map>
#include <memory>
#include <mutex>
#include <condition_variable>
#include <thread>
#include "../common.h"
#include "minsize.h"

namespace pb200 {

struct AlignParams {                 // ini keys, src/parsnp.cpp:2866-2901
    int c = 21;                      // [LCB] c
    int d = 300;                     // [LCB] d
    int q = 30;                      // [LCB] q
    int64_t p = 15000000;            // [LCB] p
    float diag_diff = 0.12f;         // [LCB] diagdiff
    int random = 1;                  // [MUM] filter
    bool anchors_only = false;       // [MUM] anchorsonly
    std::string anchors = "1.1*(Log(S))";
    std::string mums = "1.1*(Log(S))";
};

// packed mumlayout row (src/parsnp.cpp:3181-3186: len+1 bits, sentinel bit at len)
class BitRow {
public:
    void init(int64_t nbits_with_sentinel);
    // (words are read with relaxed atomic loads: the replay tasks and the speculative passes read rows that other threads
    //  extend with atomic ORs at other bit positions of the same words)
    inline uint64_t word(int64_t wi) const { return __atomic_load_n(&w_[(size_t)wi], __ATOMIC_RELAXED); }
    inline bool get(int64_t i) const { return (word(i >> 6) >> (i & 63)) & 1ull; }
    inline void set_range(int64_t a, int64_t b) {   // [a,b)
        if (a >= b) return;
        const int64_t wa = a >> 6, wb = (b - 1) >> 6;
        if (wa == wb) w_[wa] |= (~0ull << (a & 63)) & (~0ull >> (63 - ((b - 1) & 63)));
        else set_range_slow(a, b);
    }
    void set_range_slow(int64_t a, int64_t b);
    void set_range_atomic(int64_t a, int64_t b);   // same, safe against concurrent writers of neighbouring bits
    // same with plain (relaxed atomic) loads and stores instead of locked read-modify-writes for the words strictly inside
    // (wlo, whi): for a writer that is the only one for those words; the words at and outside the bounds are OR-ed atomically
    void set_range_owned(int64_t a, int64_t b, int64_t wlo, int64_t whi);
    void clear_range(int64_t a, int64_t b);   // [a,b)
    // # consecutive set bits a, a+1, ... (< b); the common case (bit a clear) is answered inline
    inline int64_t run_up(int64_t a, int64_t b) const { return (a >= b || !get(a)) ? 0 : run_up_slow(a, b); }
    // # consecutive set bits b-1, b-2, ... (>= a)
    inline int64_t run_down(int64_t a, int64_t b) const { return (a >= b || !get(b - 1)) ? 0 : run_down_slow(a, b); }
    int64_t run_up_slow(int64_t a, int64_t b) const;
    int64_t run_down_slow(int64_t a, int64_t b) const;
    int64_t prev_set(int64_t i) const;        // largest set index <= i, or -1
    int64_t prev_set_from(int64_t i, int64_t lo) const;   // largest set index in [lo, i], or -1
    void copy_range_from(const BitRow& src, int64_t a, int64_t b);   // bits [a,b) := src's (atomic per word)
    int64_t next_set(int64_t i, int64_t limit) const;   // smallest set index in [i,limit), or limit
    int64_t nbits() const { return nbits_; }
    const uint64_t* words() const { return w_.data(); }
    uint64_t* words_mut() { return w_.data(); }
    int64_t nwords() const { return (int64_t)w_.size(); }
private:
    std::vector<uint64_t> w_;
    int64_t nbits_ = 0;
};

struct MumRec {
    int64_t length;
    int64_t slength;
    int64_t off;       // offset into mum_start_/mum_fwd_ pools (n entries)
    bool alive;
};

struct ClusterRec {
    int type;                       // 1 = LCB, 0 = inter-cluster record
    int64_t length;
    std::vector<int> mums;          // indices into the (sorted) final MUM list; empty for type 0
    std::vector<int64_t> start, end;
};

// flat pools (shared by the sequential path; thread-local in the chunk-parallel paths)
struct RegionPool {
    int n = 0;
    pod_vector<int64_t> coord;      // 2n per region: start[n], end[n] (resize leaves new rows uninitialised: filled by the caller)
    pod_vector<int64_t> slen;       // TRegion::slength
    int add(const int64_t* start, const int64_t* end);
    const int64_t* ext_coord = nullptr;   // a read-only pool over somebody else's array (the engine's discovery result)
    inline const int64_t* start(int r) const { return (ext_coord ? ext_coord : coord.data()) + (size_t)r * 2 * n; }
    inline const int64_t* end(int r) const { return start(r) + n; }
    inline int size() const { return (int)slen.size(); }
};
struct MumPool {
    std::vector<MumRec> mums;
    pod_vector<int64_t> start;      // (pod_vector: a parallel fill after resize() touches the new pages first)
    pod_vector<uint8_t> fwd;
};
struct ReplayCtx;
struct AlignStats {
    int64_t anchors = 0, regions_searched = 0, spec_regions = 0, replay_misses = 0, spec_levels = 0,
            windows_searched = 0, candidates = 0, slow_queue_iters = 0, host_threads = 1;
    double t_anchor_search = 0, t_anchor_host = 0, t_spec_search = 0, t_spec_host = 0, t_replay = 0,
           t_replay_search = 0, t_lcb = 0, t_total = 0;
    double t_search_prep = 0, t_search_backend = 0, t_search_cache = 0;   // split of the search_regions calls (all phases)
    double t_replay_wait = 0;        // replay blocked on a speculation slice still in flight
    int64_t spec_slices = 0;
    int64_t mums_filtered = 0, clusters_filtered = 0;      // Aligner::filtered / filtered_clusters (src/parsnp.cpp:406,451,460)
    // parallel exact replay (replay.cpp): tasks run, foreign (outside the task's own span) reads / writes, restarts after a
    // foreign write hit a running task, 1 = the run fell back to the sequential loop (ties or non-collinear anchors), workers
    int64_t replay_tasks = 0, replay_foreign_reads = 0, replay_foreign_writes = 0, replay_restarts = 0, replay_fallback = 0,
            replay_workers = 1, spec_deferred = 0;
    int64_t replay_gaps = 0, replay_final_gaps = 0, replay_final_mums = 0;      // gaps whose accept decisions were taken from the engine as final
    double t_replay_merge = 0;
};

class Aligner {
public:
    Aligner(int n, const uint8_t* const* seq, const int64_t* len, const AlignParams& prm, SearchBackend* be);
    ~Aligner();
    // returns false when no MUMs were found (reference: "NO MUMS FOUND", src/parsnp.cpp:3223-3229)
    bool run();

    // ---- results (valid after run()) ----
    int n() const { return n_; }
    // final MUM list in the order of this->mums at writeOutput time (sorted by start[0])
    int64_t num_mums() const { return (int64_t)final_mums_.size(); }
    const MumRec& mum(int64_t i) const { return mums_[final_mums_[i]]; }
    const int64_t* mum_start(int64_t i) const { return &mum_start_[mums_[final_mums_[i]].off]; }
    const uint8_t* mum_fwd(int64_t i) const { return &mum_fwd_[mums_[final_mums_[i]].off]; }
    const std::vector<ClusterRec>& clusters() const { return clusters_; }
    const AlignStats& stats() const { return stats_; }
    // Aligner::setUnalignableRegions (src/parsnp.cpp:2310-2382): (genome, startpos, endpos) records in the reference's order
    void unaligned_regions(std::vector<int32_t>& genome, std::vector<int64_t>& start, std::vector<int64_t>& end) const;
    // sequence of searched windows in exact reference order (ref_start, ref_len) - for order tests
    const std::vector<std::pair<int64_t, int64_t>>& window_trace() const { return trace_; }
    void enable_trace(bool on) { trace_on_ = on; }
    void set_speculate(bool on) { speculate_ = on; }
    void set_threads(int t) { threads_ = t < 1 ? 1 : t; }
    // speculation slices run on their own thread while the replay consumes the finished ones.  Off = lock step: a backend whose
    // search involves collectives must see the same call sequence on every rank, so there is no speculation thread AND the
    // speculative accept walks the frontier on one thread - chunks racing on the scratch layout can accept different candidates
    // where two regions overlap in a rearranged query genome, which is harmless for one rank (the replay searches what the
    // cache misses) but would make the ranks' search calls, i.e. their collectives, diverge.
    void set_pipeline(bool on) { pipeline_ = on; }

private:
    inline const int64_t* rstart(int r) const { return rp_.start(r); }
    inline const int64_t* rend(int r) const { return rp_.end(r); }
    static uint64_t coords_hash(const int64_t* p, int count);

    struct World {                     // one copy of the mutable alignment state
        std::vector<BitRow> layout;
    };

    // ---- candidate cache: the candidates of searched regions, keyed by the regions' coordinates.  One instance per producer
    // (anchors and on-demand searches: the main thread; every speculation slice: the speculation thread); a published
    // instance is read-only.
    struct CacheEntry { int region; int64_t first_win; int nwin; };
    typedef WindowRec WinRec;
    // open-addressing index hash(coords) -> cache entry
    struct CoordIndex {
        struct Slot { uint64_t h; int32_t v; int32_t pad; };      // hash and entry side by side: one cache line per probe
        std::vector<Slot> s;
        size_t count = 0;
        void insert(uint64_t hash, int value);
        void reserve(size_t entries);          // room for `entries` more without rehashing
        // an EMPTY index filled from hashes[i] -> i for every i with valid[i], on `threads` threads (slots are claimed with a CAS)
        void build_parallel(const std::vector<uint64_t>& hashes, const std::vector<uint8_t>& valid, int threads);
        template <class Pred> int find(uint64_t hash, Pred pred) const {
            if (s.empty()) return -1;
            const size_t mask = s.size() - 1;
            for (size_t i = (size_t)hash & mask;; i = (i + 1) & mask) {
                if (s[i].v < 0) return -1;
                if (s[i].h == hash && pred(s[i].v)) return s[i].v;
            }
        }
        void prefetch(uint64_t hash) const { if (!s.empty()) __builtin_prefetch(&s[(size_t)hash & (s.size() - 1)]); }
    };
    struct CandCache {
        RegionPool rp;                         // own copies of the searched regions' coordinates (the keys)
        std::vector<CacheEntry> entries;       // .region indexes rp
        CoordIndex map;
        std::vector<WinRec> wins;
        std::vector<CandBatch> chunks;         // one per search call; candidates stay where the backend delivered them
        std::unordered_map<int64_t, int> minsize[2];
        const std::vector<uint8_t>* sorted_valid = nullptr;   // set: entry r is region r of `rp`, ascending start[0]; [r] = it was searched
        int lookup(const int64_t* coords) const;    // -> index into entries or -1
        int lookup(const int64_t* coords, uint64_t hash) const;      // hash = coords_hash(coords) computed by the caller
    };
    int minsize_cached(CandCache& C, bool anchors, int64_t slength);
    // batched GPU search of regions `regs` of pool `src`, fills C
    void search_regions(CandCache& C, const RegionPool& src, const std::vector<int>& regs, bool anchors);
    // its three parts: the reference windows of the regions, (the backend call,) the answer into the cache
    void make_window_tasks(CandCache& C, const RegionPool& src, const std::vector<int>& regs, bool anchors, std::vector<WindowTask>& tasks,
                           std::vector<int64_t>& coords, std::vector<int>& first_task);
    void install_search_result(CandCache& C, const RegionPool& src, const std::vector<int>& regs, const std::vector<WindowTask>& tasks,
                               const std::vector<int>& first_task, CandBatch& cb);
    void install_device_anchors(const AnchorResult& res, int whole);

    // setMums1 loop D on cached candidates; appends accepted MUMs to `mp`, their indices to `found`
    void accept_candidates(const int64_t* rs, const int64_t* re, int64_t rsl, const CandCache& C, int cache_idx, std::vector<BitRow>& layout,
                           MumPool& mp, std::vector<int>& found, bool atomic, bool trace);
    // the same for a long candidate list on an empty layout (anchors): non-overlapping candidates in parallel
    void accept_candidates_parallel(const int64_t* rs, const int64_t* re, int64_t rsl, const CandCache& C, int cache_idx,
                                    std::vector<BitRow>& layout, MumPool& mp, std::vector<int>& found, bool trace);
    // the two loops above and determineRegion over a layout access policy (accept_impl.h): one bitmap (DirectAccess) or a
    // replay task's view of the shared bitmaps (replay.cpp)
    template <class Acc> void accept_candidates_t(const int64_t* rs, const int64_t* re, int64_t rsl, const CandCache& C, int cache_idx, Acc& acc,
                                                  MumPool& mp, std::vector<int>& found, bool trace);
    template <class Acc> int64_t det_region_t(Acc& acc, const int64_t* mstart, int64_t mlen, bool left, int64_t* S, int64_t* E) const;
    friend struct ReplayCtx;
    friend struct ReplayTask;
    friend struct TaskAccess;
    // doWork as independent tasks (runs of consecutive initial regions) on several threads, exact: replay.cpp.
    // Returns false when the preconditions do not hold (nothing touched): the caller runs process_queue_exact.
    bool do_work_parallel();
    ReplayCtx* replay_prepare();         // nullptr = the preconditions do not hold
    bool replay_run(ReplayCtx& X);
    void replay_prepare_async();
    void set_replay_threads(int t) { replay_threads_ = t; }
    // doWork's loop over a queue of regions living in `rp`, in the exact reference order
    void process_queue_exact(const std::vector<int>& initial, RegionPool& rp, std::vector<BitRow>& layout, MumPool& mp,
                             std::vector<int>& out_mums, const std::vector<int>* slice_ids = nullptr);
    // one speculative level over frontier[a,b): children coordinates appended to `out`
    void speculate_range(const CandCache& C, const RegionPool& F, const std::vector<int>& frontier, size_t a, size_t b,
                         std::vector<BitRow>& layout, MumPool& mp, RegionPool& out, bool atomic);

    void set_initial_clusters();     // anchors
    // level-synchronous discovery for one slice of the initial regions (ids into `src`) on the scratch layout `spec`
    void speculate_slice(CandCache& C, const RegionPool& src, const std::vector<int>& initial, World& spec);
    void speculation_thread_main();
    bool discover_on_device();           // the engine follows the recursion itself (SearchBackend::discover_recursion)
    bool discover_slice(int k);
    void build_region_index(const std::vector<uint8_t>* skip);
    const CandCache* wait_slice(int slice);
    void wait_slice_quiet(int slice);
    void do_work_exact();
    void filter_random1();
    void sort_final_mums();
    void set_final_clusters(std::vector<ClusterRec>& out);
    void filter_clusters_simple(std::vector<ClusterRec>& cl);
    void set_inter_cluster_regions(std::vector<ClusterRec>& cl);

    int n_;
    std::vector<const uint8_t*> seq_;
    std::vector<int64_t> len_;
    AlignParams prm_;
    SearchBackend* be_;
    MinSizeExpr anchor_expr_, mum_expr_;

    RegionPool rp_;
    MumPool mp_;
    std::vector<MumRec>& mums_ = mp_.mums;
    pod_vector<int64_t>& mum_start_ = mp_.start;
    pod_vector<uint8_t>& mum_fwd_ = mp_.fwd;
    std::vector<int> all_mums_;       // this->mums in push order (ids into mums_)
    std::vector<int> final_mums_;
    std::vector<int> sorted_hint_;           // all_mums_ in ascending start[0] order when the producer knows it (parallel replay), else empty
    bool final_sorted_ = false;              // final_mums_ is in ascending start[0] order
    int64_t final_min_length_ = 0;           // shortest MUM seen by the last sort_final_mums()

    World truth_;
    std::vector<int> initial_regions_;
    pod_vector<int32_t> initial_lo_;         // per initial region and genome: the bounding set bit on its left, when the engine delivered it (else empty)

    CandCache main_cache_;                                  // anchors + regions the speculation did not predict
    std::vector<std::unique_ptr<CandCache>> slice_cache_;   // one per speculation slice
    std::vector<std::vector<int>> slice_regions_;           // initial regions of every slice (ids into frozen_rp_)
    std::vector<int> slice_of_initial_;                     // slice of initial_regions_[i]
    RegionPool frozen_rp_;                                  // the initial regions' coordinates as the speculation thread sees them
    World spec_world_;                                      // scratch copy of mumlayout for the speculation
    std::shared_ptr<const std::vector<int32_t>> disc_tab_;  // device discovery: minsize table, initial coordinates, slice bounds
    pod_vector<int64_t> disc_coords_;
    std::vector<uint64_t> disc_hashes_;                     // the discovery's regions: coordinate hashes and "was searched", kept while the
    std::vector<uint8_t> disc_valid_;                       // index over them is deferred (build_region_index)
    bool disc_index_deferred_ = false;
    std::vector<size_t> disc_begin_;
    std::thread spec_thread_;
    std::mutex slice_mu_, backend_mu_;
    std::condition_variable slice_cv_;
    int slices_ready_ = 0;                                  // slices [0, slices_ready_) are published (guarded by slice_mu_)
    std::exception_ptr spec_error_;
    bool pipeline_ = true;
    bool anchors_on_device_ = false;                        // the engine placed the anchors and is following the recursion (anchor_stage)
    // the engine's own accept decisions of its discovery (RecursionResult: views into its pinned memory, valid during the replay),
    // which the parallel replay takes as final for the gaps where they cannot depend on the order (replay.cpp); valid = false:
    // every gap is replayed
    struct DeviceDecisions {
        bool valid = false;
        size_t nregions = 0, ncands = 0, nfw = 0;
        const int64_t* coords = nullptr; const int64_t* slen = nullptr; const WindowRec* wins = nullptr;
        const int32_t* k = nullptr; const int32_t* lon = nullptr; const int32_t* sp = nullptr; const uint8_t* fwd = nullptr;
        const uint32_t* flags = nullptr; const int32_t* parent = nullptr; const int32_t* acc_shift = nullptr; const int32_t* acc_len = nullptr;
        const int32_t* fw = nullptr;
    } dev_;
    int replay_threads_ = 0;                                // 0 = threads_
    ReplayCtx* replay_ctx_ = nullptr;
    std::thread replay_prep_thread_;
    std::exception_ptr replay_prep_error_;
    bool replay_prepared_ = false;

    std::vector<ClusterRec> clusters_;
    int threads_ = 1;
    AlignStats stats_;
    std::vector<std::pair<int64_t, int64_t>> trace_;
    bool trace_on_ = false;
    bool speculate_ = true;
};

}  // namespace pb200
