// Multi-GPU sharding of the search (one process per GPU, the host orchestrator replicated on every rank).
//
// The reference has no distributed backend; this is the B200-side design of section 8(e):
//   * large windows (the anchor search, src/parsnp.cpp:1570-1695): QUERIES are sharded in contiguous blocks that keep
//     the ini order; the window index is built on rank 0 and broadcast; every rank scans and folds its own queries;
//     one exchange step makes the fold exact -
//         Merge_Master picks the strand of query i from min(M, EPf) > min(M, EPc) with M = min over EARLIER queries of
//         max(EPf, EPc) (src/csgmum/mum.c:92-123, 170-171), so each rank needs the prefix-min of the earlier ranks'
//         block minima: all-gather of the per-rank block minima, local prefix, second local fold, then all-reduce
//         (min EP, max UP); emission is replicated, the per-candidate (strand, start) columns are all-gathered;
//   * small windows (the recursion, src/parsnp.cpp:173-317): WINDOWS are sharded in contiguous blocks, results
//     all-gathered and re-assembled in task order.
// The collectives go through a `Comm` supplied by the caller (torch.distributed: NCCL over NVLink on the GPUs, gloo in
// the CPU tests), so this file has no GPU or NCCL dependency.
#pragma once
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>
#include "../common.h"

namespace pb200 {

struct Comm {
    int rank = 0, world = 1;
    virtual ~Comm() {}
    // recv holds world * bytes_per_rank; `device` = the pointers are device memory of this rank's GPU
    virtual void allgather(const void* send, void* recv, size_t bytes_per_rank, bool device) = 0;
    virtual void allreduce_i32(int32_t* buf, size_t count, bool is_max, bool device) = 0;
    virtual void bcast(void* buf, size_t bytes, int root, bool device) = 0;
};

// What a search engine must expose so that one window can be searched with its queries sharded over ranks.
class StagedWindowEngine {
public:
    virtual ~StagedWindowEngine() {}
    virtual bool wants_staged(const WindowTask& t, const int64_t* coords) = 0;   // true: window goes through the staged path
    virtual bool buffers_on_device() const = 0;
    virtual void window_begin(const WindowTask& t, const int64_t* coords, bool build_index) = 0;
    virtual void window_index_buffers(std::vector<std::pair<void*, size_t>>& bufs) = 0;   // to broadcast from the builder
    virtual int window_n() const = 0;
    virtual void window_scan(int q0, int q1) = 0;                 // events of queries [q0, q1) (0-based query index)
    virtual void window_fold(bool init) = 0;                      // fold local queries into Master
    virtual int32_t* window_master_up() = 0;                      // n ints
    virtual int32_t* window_master_ep() = 0;                      // n ints
    virtual int32_t* window_gather_buffer(size_t ints) = 0;       // scratch (device or host like the master arrays)
    // gathered = world x n block minima; sets initEP = min(n, blocks of ranks < rank), Master = (UP 0, EP initEP)
    virtual void window_apply_prefix(const int32_t* gathered, int world, int rank) = 0;
    virtual uint32_t window_emit() = 0;                           // candidates from the current (global) Master
    // k, lon for every candidate; sp/fwd [ncand x (q1-q0)] for the local queries, replayed from initEP
    virtual void window_pass2(std::vector<int32_t>& k, std::vector<int32_t>& lon, std::vector<int32_t>& sp, std::vector<uint8_t>& fwd) = 0;
};

class ShardedBackend : public SearchBackend {
public:
    ShardedBackend(SearchBackend* local, StagedWindowEngine* staged, Comm* comm, bool bcast_index)
        : local_(local), staged_(staged), comm_(comm), bcast_index_(bcast_index) {}
    void set_genomes(int n, const uint8_t* const* seq, const int64_t* len) override { n_ = n; local_->set_genomes(n, seq, len); }
    void set_n(int n) { n_ = n; }
    void search(const WindowTask* tasks, int ntasks, const int64_t* coords, CandBatch& out) override;
    int64_t staged_windows = 0, sharded_small_windows = 0;

private:
    void search_staged(const WindowTask& t, const int64_t* coords, std::vector<int32_t>& k, std::vector<int32_t>& lon,
                       std::vector<int32_t>& sp, std::vector<uint8_t>& fwd);
    SearchBackend* local_;
    StagedWindowEngine* staged_;
    Comm* comm_;
    bool bcast_index_;
    int n_ = 0;
};

}  // namespace pb200
