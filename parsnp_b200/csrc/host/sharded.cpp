#include "sharded.h"
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>

namespace pb200 {

namespace {
inline size_t pad16(size_t x) { return (x + 15) & ~(size_t)15; }
}

void ShardedBackend::search_staged(const WindowTask& t, const int64_t* coords, std::vector<int32_t>& k, std::vector<int32_t>& lon,
                                   std::vector<int32_t>& sp, std::vector<uint8_t>& fwd) {
    const int nq = n_ - 1;
    const int W = comm_->world, r = comm_->rank;
    const bool dev = staged_->buffers_on_device();
    const int q0 = (int)((int64_t)nq * r / W), q1 = (int)((int64_t)nq * (r + 1) / W);
    // 1. window index: built once (rank 0) and broadcast, or rebuilt everywhere
    staged_->window_begin(t, coords, bcast_index_ ? (r == 0) : true);
    if (bcast_index_) {
        std::vector<std::pair<void*, size_t>> bufs;
        staged_->window_index_buffers(bufs);
        for (auto& b : bufs) comm_->bcast(b.first, b.second, 0, dev);
    }
    const int n = staged_->window_n();
    // 2. local scan + first fold: block minimum B_r[k] = min over local queries of max(EPf, EPc)
    staged_->window_scan(q0, q1);
    staged_->window_fold(true);
    // 3. exchange block minima, exact prefix for the order-dependent strand choice, second fold
    int32_t* gathered = staged_->window_gather_buffer((size_t)W * n);
    comm_->allgather(staged_->window_master_ep(), gathered, (size_t)n * 4, dev);
    staged_->window_apply_prefix(gathered, W, r);
    staged_->window_fold(false);
    // 4. global Master: EP = min over ranks, UP = max over ranks
    comm_->allreduce_i32(staged_->window_master_ep(), (size_t)n, false, dev);
    comm_->allreduce_i32(staged_->window_master_up(), (size_t)n, true, dev);
    // 5. emission (replicated) and the local columns of every candidate
    const uint32_t ncand = staged_->window_emit();
    std::vector<int32_t> lk, llon, lsp;
    std::vector<uint8_t> lfwd;
    staged_->window_pass2(lk, llon, lsp, lfwd);
    if (lk.size() != ncand) throw std::runtime_error("sharded search: candidate count mismatch");
    // 6. all-gather the columns (padded to the widest block) and interleave them back in query order
    const int wmax = (nq + W - 1) / W + 1;
    const size_t blk = pad16((size_t)ncand * wmax * 4) + pad16((size_t)ncand * wmax);
    std::vector<uint8_t> send(blk, 0), recv(blk * W, 0);
    const int nl = q1 - q0;
    for (uint32_t c = 0; c < ncand; ++c) {
        if (nl) std::memcpy(send.data() + ((size_t)c * wmax) * 4, &lsp[(size_t)c * nl], (size_t)nl * 4);
        if (nl) std::memcpy(send.data() + pad16((size_t)ncand * wmax * 4) + (size_t)c * wmax, &lfwd[(size_t)c * nl], (size_t)nl);
    }
    if (blk) comm_->allgather(send.data(), recv.data(), blk, false);
    const size_t base = k.size();
    k.insert(k.end(), lk.begin(), lk.end());
    lon.insert(lon.end(), llon.begin(), llon.end());
    const size_t bsp = sp.size();
    sp.resize(bsp + (size_t)ncand * nq);
    fwd.resize(bsp + (size_t)ncand * nq);
    (void)base;
    for (int p = 0; p < W; ++p) {
        const int p0 = (int)((int64_t)nq * p / W), p1 = (int)((int64_t)nq * (p + 1) / W);
        const uint8_t* b = recv.data() + blk * p;
        for (uint32_t c = 0; c < ncand; ++c) {
            if (p1 > p0) {
                std::memcpy(&sp[bsp + (size_t)c * nq + p0], b + ((size_t)c * wmax) * 4, (size_t)(p1 - p0) * 4);
                std::memcpy(&fwd[bsp + (size_t)c * nq + p0], b + pad16((size_t)ncand * wmax * 4) + (size_t)c * wmax, (size_t)(p1 - p0));
            }
        }
    }
    staged_windows++;
}

void ShardedBackend::search(const WindowTask* tasks, int ntasks, const int64_t* coords, CandBatch& out) {
    const int nq = n_ - 1;
    const int W = comm_->world, r = comm_->rank;
    out.clear();
    out.nq = nq;
    // The replicated orchestrators must arrive here in lock step with the same windows; a rank that is somewhere else would
    // have its blocks attributed to the wrong windows.  One 16-byte all-gather per call turns that into an error.
    {
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
        for (int t = 0; t < ntasks; ++t) {
            mix((uint64_t)tasks[t].ref_start); mix((uint64_t)tasks[t].ref_len); mix((uint64_t)(uint32_t)tasks[t].minsize);
            const int64_t* c = coords + tasks[t].coord_off;
            for (int j = 0; j < 2 * nq; ++j) mix((uint64_t)c[j]);
        }
        uint64_t mine[2] = {(uint64_t)ntasks, h};
        std::vector<uint64_t> all((size_t)2 * W);
        comm_->allgather(mine, all.data(), sizeof(mine), false);
        for (int p = 0; p < W; ++p)
            if (all[2 * p] != mine[0] || all[2 * p + 1] != mine[1])
                throw std::runtime_error("sharded search: rank " + std::to_string(r) + " and rank " + std::to_string(p) +
                                         " are not searching the same windows (" + std::to_string(ntasks) + " vs " +
                                         std::to_string(all[2 * p]) + " tasks)");
    }
    std::vector<int> small_ids, staged_ids;
    for (int t = 0; t < ntasks; ++t) (staged_->wants_staged(tasks[t], coords) ? staged_ids : small_ids).push_back(t);
    std::vector<int32_t> t_cnt(ntasks, 0);
    std::vector<int64_t> t_base(ntasks, 0);
    std::vector<int8_t> t_src(ntasks, 0);
    // ---- small windows: contiguous blocks of the task list per rank
    std::vector<int32_t> sm_k, sm_lon, sm_sp;
    std::vector<uint8_t> sm_fwd;
    const int ns = (int)small_ids.size();
    if (ns) {
        const int a = (int)((int64_t)ns * r / W), b = (int)((int64_t)ns * (r + 1) / W);
        std::vector<WindowTask> mine;
        for (int i = a; i < b; ++i) mine.push_back(tasks[small_ids[i]]);
        CandBatch cb;
        cb.nq = nq;
        if (!mine.empty()) { local_->search(mine.data(), (int)mine.size(), coords, cb); cb.compact((int)mine.size()); }
        else cb.off.assign(1, 0);
        int64_t hdr[2] = {(int64_t)mine.size(), (int64_t)cb.k.size()};
        std::vector<int64_t> hdrs((size_t)2 * W);
        comm_->allgather(hdr, hdrs.data(), sizeof(hdr), false);
        int64_t maxT = 0, maxC = 0;
        for (int p = 0; p < W; ++p) { maxT = std::max(maxT, hdrs[2 * p]); maxC = std::max(maxC, hdrs[2 * p + 1]); }
        const size_t o_cnt = 0, o_k = pad16((size_t)maxT * 4), o_lon = o_k + pad16((size_t)maxC * 4),
                     o_sp = o_lon + pad16((size_t)maxC * 4), o_fwd = o_sp + pad16((size_t)maxC * nq * 4),
                     blk = o_fwd + pad16((size_t)maxC * nq);
        std::vector<uint8_t> send(blk, 0), recv(blk * W, 0);
        for (size_t i = 0; i < mine.size(); ++i) {
            int32_t c = (int32_t)(cb.off[i + 1] - cb.off[i]);
            std::memcpy(send.data() + o_cnt + i * 4, &c, 4);
        }
        if (!cb.k.empty()) {
            std::memcpy(send.data() + o_k, cb.k.data(), cb.k.size() * 4);
            std::memcpy(send.data() + o_lon, cb.lon.data(), cb.lon.size() * 4);
            if (nq) {
                std::memcpy(send.data() + o_sp, cb.sp.data(), cb.sp.size() * 4);
                std::memcpy(send.data() + o_fwd, cb.fwd.data(), cb.fwd.size());
            }
        }
        if (blk) comm_->allgather(send.data(), recv.data(), blk, false);
        for (int p = 0; p < W; ++p) {
            const int pa = (int)((int64_t)ns * p / W), pb = (int)((int64_t)ns * (p + 1) / W);
            const uint8_t* bptr = recv.data() + blk * p;
            const int64_t pc = hdrs[2 * p + 1];
            const size_t hb = sm_k.size();
            sm_k.resize(hb + pc); sm_lon.resize(hb + pc); sm_sp.resize((hb + pc) * nq); sm_fwd.resize((hb + pc) * nq);
            if (pc) {
                std::memcpy(sm_k.data() + hb, bptr + o_k, (size_t)pc * 4);
                std::memcpy(sm_lon.data() + hb, bptr + o_lon, (size_t)pc * 4);
                if (nq) {
                    std::memcpy(sm_sp.data() + hb * nq, bptr + o_sp, (size_t)pc * nq * 4);
                    std::memcpy(sm_fwd.data() + hb * nq, bptr + o_fwd, (size_t)pc * nq);
                }
            }
            int64_t run = (int64_t)hb;
            for (int i = pa; i < pb; ++i) {
                int32_t c;
                std::memcpy(&c, bptr + o_cnt + (size_t)(i - pa) * 4, 4);
                t_cnt[small_ids[i]] = c;
                t_base[small_ids[i]] = run;
                run += c;
            }
        }
        sharded_small_windows += ns;
    }
    // ---- large windows: queries sharded, one exchange per window
    std::vector<int32_t> bg_k, bg_lon, bg_sp;
    std::vector<uint8_t> bg_fwd;
    for (int t : staged_ids) {
        t_src[t] = 1;
        t_base[t] = (int64_t)bg_k.size();
        search_staged(tasks[t], coords, bg_k, bg_lon, bg_sp, bg_fwd);
        t_cnt[t] = (int32_t)((int64_t)bg_k.size() - t_base[t]);
    }
    // ---- assemble in task order
    out.off.resize(ntasks + 1);
    int64_t tot = 0;
    for (int t = 0; t < ntasks; ++t) { out.off[t] = tot; tot += t_cnt[t]; }
    out.off[ntasks] = tot;
    out.k.resize(tot); out.lon.resize(tot); out.sp.resize((size_t)tot * nq); out.fwd.resize((size_t)tot * nq);
    for (int t = 0; t < ntasks; ++t) {
        const int32_t cnt = t_cnt[t];
        if (!cnt) continue;
        const std::vector<int32_t>& sk = t_src[t] ? bg_k : sm_k;
        const std::vector<int32_t>& sl = t_src[t] ? bg_lon : sm_lon;
        const std::vector<int32_t>& ss = t_src[t] ? bg_sp : sm_sp;
        const std::vector<uint8_t>& sf = t_src[t] ? bg_fwd : sm_fwd;
        std::memcpy(&out.k[out.off[t]], &sk[t_base[t]], (size_t)cnt * 4);
        std::memcpy(&out.lon[out.off[t]], &sl[t_base[t]], (size_t)cnt * 4);
        if (nq) {
            std::memcpy(&out.sp[(size_t)out.off[t] * nq], &ss[(size_t)t_base[t] * nq], (size_t)cnt * nq * 4);
            std::memcpy(&out.fwd[(size_t)out.off[t] * nq], &sf[(size_t)t_base[t] * nq], (size_t)cnt * nq);
        }
    }
}

}  // namespace pb200
