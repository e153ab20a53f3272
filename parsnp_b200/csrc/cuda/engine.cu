// CUDA search engine + C ABI of libparsnp_b200.so (see include/parsnp_b200.h).  sm_100a only, no CPU path.
#include <cuda_runtime.h>
#include <malloc.h>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <cstdlib>
#include <string>
#include <memory>
#include <mutex>
#include <numeric>
#include "../../../include/parsnp_b200.h"
#include "../common.h"
#include "../host/aligner.h"
#include "../host/result.h"
#include "../host/sharded.h"
#include "../host/parallel.h"
#include "util.cuh"
#include "bigpath.cuh"
#include "smallpath.cuh"
#include "recursion.cuh"
#include "anchors.cuh"

namespace pb200 {

int64_t g_kernel_launches = 0;

// ASCII (A,C,G,T,N) -> forward codes and reverse-complement codes (A0 C1 G2 T3 N4; complement = 3-c, N stays N)
__global__ void encode_kernel(const uint8_t* __restrict__ ascii, int64_t len, uint8_t* __restrict__ fwd, uint8_t* __restrict__ rc) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    uint8_t a = ascii[i], c;
    switch (a) { case 'A': c = 0; break; case 'C': c = 1; break; case 'G': c = 2; break; case 'T': c = 3; break; default: c = 4; }
    fwd[i] = c;
    if (rc) rc[len - 1 - i] = c < 4 ? (uint8_t)(3 - c) : (uint8_t)4;
}

// initEP[k] = min(n, block minima of the ranks before `rank`); Master = (UP 0, EP initEP)
__global__ void prefix_min_kernel(const int32_t* __restrict__ gathered, int world, int rank, int n, int32_t* __restrict__ initEP,
                                  int32_t* __restrict__ MUP, int32_t* __restrict__ MEP) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int m = n;
    for (int r = 0; r < rank; ++r) m = min(m, gathered[(size_t)r * n + k]);
    initEP[k] = m;
    MUP[k] = 0;
    MEP[k] = m;
}

// anchors.cuh kernels with the number of anchors read on the device (launched for the number of candidates)
__global__ void anchor_push_flags_bounded(int n, int q, const uint32_t* __restrict__ nanchors, unsigned int ncand, const int32_t* __restrict__ REG,
                                          const int32_t* __restrict__ SL, uint32_t* __restrict__ slot) {
    const unsigned int x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (x >= ncand) return;                                 // (slot holds 2 entries per candidate)
    const unsigned int na = *nanchors;
    if (x >= na) { if (lane == 0) { slot[2 * x] = 0; slot[2 * x + 1] = 0; } return; }
    const int32_t* me = REG + (size_t)x * 4 * n;
    int d_prev = 0, d_own = 0;
    for (int i = lane; i < 2 * n; i += 32) {
        if (x > 0 && me[i] != (me - 4 * n)[2 * n + i]) d_prev = 1;
        if (me[i] != me[2 * n + i]) d_own = 1;
    }
    d_prev = __any_sync(0xffffffffu, d_prev);
    d_own = __any_sync(0xffffffffu, d_own);
    if (lane == 0) {
        slot[2 * x] = (SL[2 * x] > q && (x == 0 || d_prev)) ? 1u : 0u;
        slot[2 * x + 1] = (SL[2 * x + 1] > q && d_own) ? 1u : 0u;
    }
}
__global__ void anchor_push_bounded(int n, const uint32_t* __restrict__ nanchors, const int32_t* __restrict__ REG, const int32_t* __restrict__ SL,
                                    const uint32_t* __restrict__ slot, const uint32_t* __restrict__ pos, rec::Store St, uint8_t* __restrict__ pair,
                                    unsigned int cap, const unsigned long long* __restrict__ bits, const int64_t* __restrict__ bit_off,
                                    int32_t* __restrict__ LO) {
    const unsigned int x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned int na = *nanchors;
    if (x >= na) return;
    const int32_t* me = REG + (size_t)x * 4 * n;
    for (int side = 0; side < 2; ++side) {
        if (!slot[2 * x + side]) continue;
        const unsigned int id = pos[2 * x + side];
        if (id >= cap) continue;
        const int32_t* S = me + 2 * n * side;
        int32_t* c = St.coords + (size_t)id * 2 * n;
        for (int g = lane; g < n; g += 32) {
            c[g] = S[g]; c[n + g] = S[n + g] - S[g];
            // the set bit that bounds the region on the left (the replay's ownership spans, host/replay.cpp): a left side starts
            // right behind one; a right side starts two bases behind the anchor's last base, or one if the skipped base is set
            const int32_t b = S[g] - 1;
            LO[(size_t)id * n + g] = b <= 0 ? 0 : (rec::bit_get(bits + bit_off[g], b) ? b : b - 1);
        }
        if (lane == 0) {
            int p = 0;
            if (side == 1 && x + 1 < na && slot[2 * x + 2]) {
                const int32_t* nx = me + 4 * n;                 // left side of the next anchor
                p = nx[0] == S[0] - 1 && nx[n] == S[n];
            }
            pair[id] = (uint8_t)p;
        }
    }
    (void)SL;
}

class CudaEngine : public SearchBackend, public StagedWindowEngine {
public:
    explicit CudaEngine(int device) : device_(device) {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) throw CudaError("parsnp_b200 requires a CUDA device (none found); there is no CPU fallback");
        if (device < 0 || device >= count) throw CudaError("parsnp_b200: bad device ordinal");
        PB_CUDA(cudaSetDevice(device));
        int cc_major = 0;
        PB_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
        if (cc_major < 10) throw CudaError("parsnp_b200 kernels are built for sm_100a (Blackwell) only");
        PB_CUDA(cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, device));
        PB_CUDA(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
        {   // keep freed blocks in the pool: later engines (one per pb200_align call) re-use them without cudaMalloc
            cudaMemPool_t pool;
            PB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
            uint64_t thr = UINT64_MAX;
            PB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        }
        big_.tm = &timers;
        const char* fp = getenv("PB200_FORCE_PATH");      // tests: "big" routes every window through the large-window path
        force_big_ = (fp && std::string(fp) == "big") ? 1 : 0;
        const char* fk = getenv("PB200_FORCE_3BIT_KEYS");  // tests: always use the 21-mer 3-bit key path
        force_3bit_ = fk && fk[0] == '1';
        // {n_cap (power of two), m_cap, event store capacity (all strands), candidate capacity, threads,
        //  group text bytes (>= 2 al(m_cap + 15)), group seed rows (>= m_cap), seed-hit queue, staged events}
        classes_[0] = small::ClassCfg{256, 512, 512, 64, 128, 3584, 640, 512, 192};
        classes_[1] = small::ClassCfg{1024, 2048, 1536, 256, 256, 8192, 2048, 1024, 384};
        classes_[2] = small::ClassCfg{4096, 4096, 3072, 512, 256, 8448, 4096, 2048, 1024};
        PB_CUDA(cudaFuncSetAttribute(small::small_region_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    ~CudaEngine() override {
        cudaSetDevice(device_);
        if (st_) { cudaStreamSynchronize(st_); cudaStreamDestroy(st_); }
        if (st2_) { cudaStreamSynchronize(st2_); cudaStreamDestroy(st2_); }
    }

    void set_genomes(int n, const uint8_t* const* seq, const int64_t* len) override {
        PB_CUDA(cudaSetDevice(device_));
        n_ = n;
        len_.assign(len, len + n);
        gfwd_.assign(n, 0);
        grc_.assign(n, 0);
        int64_t tot = 0, maxlen = 0;
        auto pad = [](int64_t x) { return (x + 63) & ~(int64_t)63; };
        for (int i = 0; i < n; ++i) {
            if (len[i] >= ((int64_t)1 << 31) - 64) throw CudaError("parsnp_b200: genome longer than 2^31 bases");
            gfwd_[i] = tot; tot += pad(len[i] + 16);
            if (i > 0) { grc_[i] = tot; tot += pad(len[i] + 16); }
            maxlen = std::max(maxlen, len[i]);
        }
        uint8_t* text = text_.ensure((size_t)tot + 256, false, st_);
        PB_CUDA(cudaMemsetAsync(text, 7, (size_t)tot + 256, st_));         // 7 never matches a base code
        // raw ASCII staged per genome (all copies and encode launches are queued back to back, one sync at the end)
        (void)maxlen;
        int64_t raw_tot = 0;
        std::vector<int64_t> raw_off(n, 0);
        for (int i = 0; i < n; ++i) { raw_off[i] = raw_tot; raw_tot += pad(len[i]); }
        uint8_t* stage = stage_.ensure((size_t)raw_tot + 64, false, st_);
        for (int i = 0; i < n; ++i) {
            if (len[i] == 0) continue;
            PB_CUDA(cudaMemcpyAsync(stage + raw_off[i], seq[i], (size_t)len[i], cudaMemcpyHostToDevice, st_));
            pb200::launch(encode_kernel, (unsigned)((len[i] + 255) / 256), 256, 0, st_, stage + raw_off[i], len[i], text + gfwd_[i],
                          i > 0 ? text + grc_[i] : nullptr);
        }
        h2d_bytes_ = 0;
        for (int i = 0; i < n; ++i) h2d_bytes_ += len[i];
        // positions of non-ACGT symbols in the reference (N is an ordinary symbol; windows without any take the 2-bit key path);
        // scanned in blocks so that the common all-ACGT block is one vectorisable reduction
        ref_n_pos_.clear();
        const int64_t BLK = 4096;
        for (int64_t b0 = 0; b0 < len[0]; b0 += BLK) {
            const int64_t b1 = std::min(len[0], b0 + BLK);
            const uint8_t* p = seq[0];
            int bad = 0;
            for (int64_t i = b0; i < b1; ++i) { const uint8_t c = p[i]; bad += (c != 'A') & (c != 'C') & (c != 'G') & (c != 'T'); }
            if (!bad) continue;
            for (int64_t i = b0; i < b1; ++i) {
                const uint8_t c = p[i];
                if (c != 'A' && c != 'C' && c != 'G' && c != 'T') ref_n_pos_.push_back(i);
            }
        }
        int64_t* d = gmeta_.ensure((size_t)3 * n, false, st_);
        std::vector<int64_t> meta(3 * (size_t)n);
        for (int i = 0; i < n; ++i) { meta[i] = gfwd_[i]; meta[n + i] = grc_[i]; meta[2 * n + i] = len_[i]; }
        PB_CUDA(cudaMemcpyAsync(d, meta.data(), meta.size() * 8, cudaMemcpyHostToDevice, st_));
        stage_.release(st_);                                                 // stream-ordered: freed after the encode kernels ran
        PB_CUDA(cudaStreamSynchronize(st_));
    }

    void search(const WindowTask* tasks, int ntasks, const int64_t* coords, CandBatch& out) override {
        PB_CUDA(cudaSetDevice(device_));
        const int nq = n_ - 1;
        out.clear();
        out.nq = nq;
        // find the extent of the coordinate pool
        int64_t ncoords = 0;
        for (int t = 0; t < ntasks; ++t) ncoords = std::max(ncoords, tasks[t].coord_off + 2 * (int64_t)nq);
        // classify
        const double th0 = wall_s();
        std::vector<int> cls(ntasks, 3);
        std::vector<std::vector<int>> by_class(4);
        for (int t = 0; t < ntasks; ++t) {
            const int c = classify(tasks[t], coords);
            cls[t] = c;
            by_class[c].push_back(t);
        }
        // per-task results
        std::vector<int64_t> t_base(ntasks, 0);
        std::vector<int32_t> t_cnt(ntasks, 0);
        std::vector<int8_t> t_src(ntasks, 0);              // 0 = small arrays, 1 = big arrays
        small_cands_ = 0;
        bg_k_.clear(); bg_lon_.clear(); bg_sp_.clear(); bg_fwd_.clear();
        const bool any_small = !by_class[0].empty() || !by_class[1].empty() || !by_class[2].empty();
        const double th1 = wall_s();
        host_classify_s += th1 - th0;
        if (any_small) upload_small_tasks(tasks, ntasks, coords, ncoords);
        host_upload_s += wall_s() - th1;
        for (int c = 0; c < 3; ++c) {
            if (by_class[c].empty()) continue;
            std::vector<int> retry;
            run_small(c, by_class[c], nq, t_base, t_cnt, retry);
            for (int t : retry) by_class[c + 1].push_back(t);
        }
        const double th2 = wall_s();
        for (int t : by_class[3]) {
            t_src[t] = 1;
            t_base[t] = (int64_t)bg_k_.size();
            run_big(tasks[t], coords, nq);
            t_cnt[t] = (int32_t)((int64_t)bg_k_.size() - t_base[t]);
        }
        host_big_s += wall_s() - th2;
        // gather the windows' candidate blocks (pinned staging: kernel completion order; large windows: bg_ arrays) into window
        // order, so that the accept passes (ascending reference order) stream through memory
        const double th3 = wall_s();
        out.off.assign((size_t)ntasks + 1, 0);
        for (int t = 0; t < ntasks; ++t) out.off[t + 1] = out.off[t] + t_cnt[t];
        const int64_t tot = out.off[ntasks];
        out.k.resize((size_t)tot); out.lon.resize((size_t)tot); out.sp.resize((size_t)tot * nq); out.fwd.resize((size_t)tot * nq);
        out.cnt.clear();
        {
            const int32_t* pk = pin_k_.get(); const int32_t* pl = pin_lon_.get(); const int32_t* ps = pin_sp_.get(); const uint8_t* pf = pin_fwd_.get();
            const long per = 1024;
            parallel_chunks(tot > 65536 ? default_host_threads() : 1, ((long)ntasks + per - 1) / per, [&](long c) {
                for (long t = c * per; t < std::min<long>(ntasks, (c + 1) * per); ++t) {
                    const size_t cn = (size_t)t_cnt[t], a = (size_t)t_base[t], b = (size_t)out.off[t];
                    if (!cn) continue;
                    const int32_t* sk = t_src[t] ? bg_k_.data() : pk;
                    const int32_t* sl = t_src[t] ? bg_lon_.data() : pl;
                    const int32_t* ss = t_src[t] ? bg_sp_.data() : ps;
                    const uint8_t* sf = t_src[t] ? bg_fwd_.data() : pf;
                    std::memcpy(out.k.data() + b, sk + a, cn * 4);
                    std::memcpy(out.lon.data() + b, sl + a, cn * 4);
                    if (nq) {
                        std::memcpy(out.sp.data() + b * nq, ss + a * nq, cn * nq * 4);
                        std::memcpy(out.fwd.data() + b * nq, sf + a * nq, cn * nq);
                    }
                }
            });
        }
        host_gather_s += wall_s() - th3;
    }

    // ---- the recursion followed on the device (cuda/recursion.cuh): one upload (or none: anchor_stage), a fixed number of
    // levels per round enqueued without synchronising, one download.  false = left to the host's level-by-level discovery.
    struct RecRun {
        bool active = false;
        size_t cap = 0, cand_cap = 0;
        rec::Params P; rec::Store St; rec::Queues Q;
        small::ClassCfg cfg[rec::NCLASS];
        int ctas[rec::NCLASS] = {0, 0, 0};
        int gpl = 1, level = 0;
        unsigned int* ctr = nullptr;
        unsigned long long* d_cnt = nullptr;
        int32_t* d_k = nullptr; int32_t* d_lon = nullptr; int32_t* d_sp = nullptr; uint8_t* d_fw = nullptr;
        double t0 = 0;
    } rr_;
    typedef void (*RecKernel)(const uint8_t*, const int64_t*, const int64_t*, const int64_t*, rec::Params, rec::Store, rec::Queues, int, int, small::ClassCfg,
                              unsigned long long*, unsigned long long, int32_t*, int32_t*, int32_t*, uint8_t*);
    RecKernel rec_kernel() const {
        return rr_.gpl == 1 ? rec::recursion_level_kernel<1> : (rr_.gpl == 2 ? rec::recursion_level_kernel<2> : (rr_.gpl == 4 ? rec::recursion_level_kernel<4> : rec::recursion_level_kernel<7>));
    }
    typedef void (*RecAcceptKernel)(const uint8_t*, const int64_t*, const int64_t*, rec::Params, rec::Store, rec::Queues, int, const int32_t*, const int32_t*,
                                    const int32_t*, const uint8_t*);
    RecAcceptKernel rec_accept_kernel() const {
        return rr_.gpl == 1 ? rec::recursion_accept_kernel<1> : (rr_.gpl == 2 ? rec::recursion_accept_kernel<2> : (rr_.gpl == 4 ? rec::recursion_accept_kernel<4> : rec::recursion_accept_kernel<7>));
    }
    bool rec_supported(int n, int q) const {
        if (n != n_ || n < 2 || n > 224 || q < 0 || force_big_ || getenv("PB200_NO_DEVICE_RECURSION")) return false;
        for (int g = 0; g < n; ++g) if (len_[g] >= ((int64_t)1 << 31) - 64) return false;
        return true;
    }
    // buffers and parameters for a run over at most R_est initial regions; the scratch layout is allocated, not filled
    void rec_setup(size_t R_est, const int64_t* layout_words, const int32_t* tab, int tabn, int q, int64_t p) {
        const int n = n_, nq = n - 1;
        rr_ = RecRun();
        rr_.t0 = wall_s();
        const size_t cap = R_est * 4 + 65536;
        rr_.cap = cap;
        std::vector<int64_t> bit_off((size_t)n + 1, 0);
        for (int g = 0; g < n; ++g) bit_off[(size_t)g + 1] = bit_off[(size_t)g] + layout_words[g];
        unsigned long long* d_bits = r_bits_.ensure((size_t)bit_off[(size_t)n] + 8, false, st_);
        r_bits_words_ = bit_off[(size_t)n];
        r_bit_off_host_ = bit_off;
        int64_t* d_bit_off = r_bitoff_.ensure((size_t)n + 1, false, st_);
        int64_t* h_bit_off = r_pin_bitoff_.ensure((size_t)n + 1);
        std::memcpy(h_bit_off, bit_off.data(), ((size_t)n + 1) * 8);
        PB_CUDA(cudaMemcpyAsync(d_bit_off, h_bit_off, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st_));
        int32_t* d_tab = r_tab_.ensure((size_t)std::max(tabn, 1), false, st_);
        int32_t* h_tab = r_pin_tab_.ensure((size_t)std::max(tabn, 1));
        std::memcpy(h_tab, tab, (size_t)tabn * 4);
        PB_CUDA(cudaMemcpyAsync(d_tab, h_tab, (size_t)tabn * 4, cudaMemcpyHostToDevice, st_));
        rec::Store& St = rr_.St;
        St.coords = r_coords_.ensure(cap * 2 * (size_t)n, false, st_);
        St.slen = r_slen_.ensure(cap, false, st_);
        St.minsize = r_minsize_.ensure(cap, false, st_);
        St.ncand = r_ncand_.ensure(cap, false, st_);
        St.cand_base = r_candbase_.ensure(cap, false, st_);
        St.flags = r_flags_.ensure(cap, false, st_);
        St.parent = r_parent_.ensure(cap, false, st_);
        int32_t* lists = r_lists_.ensure(cap * (2 * rec::NCLASS + 1), false, st_);
        unsigned int* ctr = r_ctr_.ensure(32, false, st_);
        PB_CUDA(cudaMemsetAsync(ctr, 0, 32 * sizeof(unsigned int), st_));
        rr_.ctr = ctr;
        rec::Queues& Q = rr_.Q;
        for (int h = 0; h < 2; ++h) for (int c = 0; c < rec::NCLASS; ++c) Q.list[h][c] = lists + cap * (size_t)(h * rec::NCLASS + c);
        Q.deferred = lists + cap * (size_t)(2 * rec::NCLASS);
        Q.count = ctr; Q.taken = ctr + 8; Q.nregions = ctr + 16; Q.ndeferred = ctr + 17; Q.dropped = ctr + 18; Q.nfw = ctr + 19; Q.taken2 = ctr + 26;
        Q.fw = r_fw_.ensure(3 * (size_t)rec::FW_CAP, false, st_);
        Q.cap = (unsigned int)std::min<size_t>(cap, 0x1fffffffu);
        r_pairflag_ = r_pair_.ensure(cap + 16, false, st_);
        // candidate arrays: sized from what earlier alignments of this process needed
        size_t cand_cap = std::max<size_t>(r_cand_hint_, R_est * s_cand_per_region_x16_.load() / 16 + 65536);
        const size_t cap_limit = std::max<size_t>((size_t)1 << 16, ((size_t)4 << 30) / (size_t)(8 + 5 * nq));       // <= 4 GiB of candidate arrays
        cand_cap = std::min(cand_cap, cap_limit);
        rr_.cand_cap = cand_cap;
        rr_.d_cnt = d_candcnt_.ensure(1, false, st_);
        PB_CUDA(cudaMemsetAsync(rr_.d_cnt, 0, 8, st_));
        rr_.d_k = d_ck_.ensure(cand_cap, false, st_);
        rr_.d_lon = d_clon_.ensure(cand_cap, false, st_);
        rr_.d_sp = d_csp_.ensure(cand_cap * (size_t)nq, false, st_);
        rr_.d_fw = d_cfwd_.ensure(cand_cap * (size_t)nq, false, st_);
        St.acc_shift = r_accshift_.ensure(cand_cap, false, st_);
        St.acc_len = r_acclen_.ensure(cand_cap, false, st_);
        rec::Params& P = rr_.P;
        P.n = n; P.q = q; P.p = p; P.minsize_tab = d_tab; P.minsize_n = tabn; P.bit_off = d_bit_off; P.bits = d_bits;
        for (int c = 0; c < rec::NCLASS; ++c) {
            rr_.cfg[c] = classes_[c];
            rr_.cfg[c].ev_cap += 4 * nq;
            if (rr_.cfg[c].ev_cap > 60000) rr_.cfg[c].ev_cap = 60000;
            P.n_cap[c] = rr_.cfg[c].n_cap; P.m_cap[c] = rr_.cfg[c].m_cap;
        }
        rr_.gpl = n <= 32 ? 1 : (n <= 64 ? 2 : (n <= 128 ? 4 : 7));
        RecKernel kern = rec_kernel();
        if (!r_attr_set_[rr_.gpl]) {
            PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            r_attr_set_[rr_.gpl] = true;
        }
        for (int c = 0; c < rec::NCLASS; ++c) {
            int per_sm = 1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, rr_.cfg[c].threads, rr_.cfg[c].smem_bytes(nq)) != cudaSuccess || per_sm < 1) per_sm = 1;
            rr_.ctas[c] = sm_count_ * per_sm;
        }
        rr_.level = 0;
        rr_.active = true;
    }
    // level 0 lists from the regions in the store (their number is read on the device), then one round of levels
    void rec_seed(size_t R_upper) {
        pb200::launch(rec::seed_lists_kernel, (unsigned)((R_upper + 255) / 256), 256, 0, st_, rr_.P, rr_.St, rr_.Q, (const unsigned int*)rr_.Q.nregions,
                      (const uint8_t*)r_pairflag_);
    }
    void rec_launch_round() {
        const int nq = n_ - 1;
        RecKernel kern = rec_kernel();
        RecAcceptKernel akern = rec_accept_kernel();
        for (int l = 0; l < 8; ++l, ++rr_.level) {
            for (int c = 0; c < rec::NCLASS; ++c) {
                timers.start(GpuTimers::T_SMALL + c, st_);
                pb200::launch(kern, rr_.ctas[c], rr_.cfg[c].threads, rr_.cfg[c].smem_bytes(nq), st_, text_.get(), gmeta_.get(), gmeta_.get() + n_, gmeta_.get() + 2 * n_,
                              rr_.P, rr_.St, rr_.Q, rr_.level, c, rr_.cfg[c], rr_.d_cnt, (unsigned long long)rr_.cand_cap, rr_.d_k, rr_.d_lon, rr_.d_sp, rr_.d_fw);
                timers.stop(GpuTimers::T_SMALL + c, st_);
            }
            // (one warp per work-list entry; the grid is sized for the first levels, the tail levels leave most of it idle for microseconds)
            timers.start(GpuTimers::T_SMALL_ACCEPT, st_);
            pb200::launch(akern, sm_count_ * 8, 128, 0, st_, text_.get(), gmeta_.get(), gmeta_.get() + 2 * n_, rr_.P, rr_.St, rr_.Q, rr_.level,
                          (const int32_t*)rr_.d_k, (const int32_t*)rr_.d_lon, (const int32_t*)rr_.d_sp, (const uint8_t*)rr_.d_fw);
            timers.stop(GpuTimers::T_SMALL_ACCEPT, st_);
            pb200::launch(rec::level_advance_kernel, 1, 32, 0, st_, rr_.Q, rr_.level);
        }
        PB_CUDA(cudaGetLastError());
    }
    bool discover_recursion(const RecursionRequest& rq, RecursionResult& out) override {
        PB_CUDA(cudaSetDevice(device_));
        const int n = n_, nq = n - 1;
        if (rq.resume) {
            if (!rr_.active) return false;                  // (anchor_stage did not start a run)
            return rec_finish(out);
        }
        if (!rec_supported(rq.n, rq.q) || rq.nregions <= 0) return false;
        const size_t R = (size_t)rq.nregions;
        if (!rq.upload_layout) {
            int64_t words = 0;
            for (int g = 0; g < n; ++g) words += rq.layout_words[g];
            if (r_bits_words_ != words || !r_bits_.get()) return false;
        }
        rec_setup(R, rq.layout_words, rq.minsize_tab, rq.minsize_n, rq.q, rq.p);
        if (rq.upload_layout)
            for (int g = 0; g < n; ++g)
                PB_CUDA(cudaMemcpyAsync(rr_.P.bits + r_bit_off_host_[(size_t)g], rq.layout[g], (size_t)rq.layout_words[g] * 8, cudaMemcpyHostToDevice, st_));
        {   // initial regions: start[n], len[n] as int32 (pinned staging)
            // + one flag per region: "this region and the next one are the two sides of one anchor gap" (the right side of anchor
            // i is pushed before the left side of anchor i+1, which starts one base earlier and is searched first)
            int32_t* h = r_pin_coords_.ensure(R * 2 * (size_t)n + (R + 3) / 4 + 16);
            uint8_t* hp = reinterpret_cast<uint8_t*>(h + R * 2 * (size_t)n);
            const long per = 4096;
            parallel_chunks(R > 16384 ? default_host_threads() : 1, ((long)R + per - 1) / per, [&](long c) {
                for (size_t r = (size_t)c * per; r < std::min(R, (size_t)(c + 1) * per); ++r) {
                    const int64_t* s = rq.coords + r * 2 * (size_t)n;
                    int32_t* d = h + r * 2 * (size_t)n;
                    for (int g = 0; g < n; ++g) { d[g] = (int32_t)s[g]; d[n + g] = (int32_t)(s[n + g] - s[g]); }
                    const int64_t* t = s + 2 * (size_t)n;
                    hp[r] = (r + 1 < R && t[0] == s[0] - 1 && t[n] == s[n]) ? 1 : 0;
                }
            });
            PB_CUDA(cudaMemcpyAsync(rr_.St.coords, h, R * 2 * (size_t)n * 4, cudaMemcpyHostToDevice, st_));
            PB_CUDA(cudaMemcpyAsync(r_pairflag_, hp, R, cudaMemcpyHostToDevice, st_));
            unsigned int* hr = r_pin_ctr_.ensure(40);
            hr[0] = (unsigned int)R;
            PB_CUDA(cudaMemcpyAsync(rr_.Q.nregions, hr, 4, cudaMemcpyHostToDevice, st_));
        }
        (void)nq;
        rec_seed(R);
        rec_launch_round();
        return rec_finish(out);
    }
    // waits for the levels in flight, runs further rounds while a deeper level exists, sorts and downloads
    bool rec_finish(RecursionResult& out) {
        const int n = n_, nq = n - 1;
        unsigned int* h_ctr = r_pin_ctr_.ensure(40);
        for (int round = 0; round < 64; ++round) {
            PB_CUDA(cudaMemcpyAsync(h_ctr, rr_.ctr, 32 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st_));
            PB_CUDA(cudaMemcpyAsync(h_ctr + 32, rr_.d_cnt, 8, cudaMemcpyDeviceToHost, st_));   // (candidates produced so far)
            PB_CUDA(cudaStreamSynchronize(st_));
            const int nx = rr_.level & 1;                   // the lists the next level would read
            if (h_ctr[nx * rec::NCLASS] + h_ctr[nx * rec::NCLASS + 1] + h_ctr[nx * rec::NCLASS + 2] == 0) break;
            rec_launch_round();
        }
        // ---- results: sorted by start[0] on the device, then one download
        const size_t NR = std::min<size_t>(h_ctr[16], rr_.cap);
        uint32_t* sk0 = r_sortk_.ensure(2 * NR + 64, false, st_);
        uint32_t* sv0 = r_sortv_.ensure(2 * NR + 64, false, st_);
        uint32_t* sk1 = sk0 + NR; uint32_t* sv1 = sv0 + NR;
        const unsigned int nr32 = (unsigned int)NR;
        pb200::launch(rec::region_keys_kernel, (unsigned)((NR + 255) / 256), 256, 0, st_, rr_.St, n, nr32, sk0, sv0);
        int key_bits = 1;
        while (((int64_t)1 << key_bits) <= len_[0] && key_bits < 32) ++key_bits;
        const int which = r_sorter_.sort<uint32_t, uint32_t>(sk0, sk1, sv0, sv1, (int64_t)NR, 0, key_bits, st_);
        const uint32_t* perm = which ? sv1 : sv0;
        uint32_t* cnt = which ? sk0 : sk1;                 // (the key buffer that does not hold the sorted keys is free)
        uint32_t* d_total = reinterpret_cast<uint32_t*>(rr_.ctr + 24);
        uint32_t* inv = r_inv_.ensure(NR + 64, false, st_);
        pb200::launch(rec::sorted_counts_kernel, (unsigned)((NR + 255) / 256), 256, 0, st_, rr_.St, perm, nr32, cnt, inv);
        r_scanner_.scan<prim::OpSum, true>(cnt, cnt, (int64_t)NR, d_total, st_);
        // (capacity of the regrouped candidate arrays = what was produced: read the counter first)
        unsigned long long used = 0;
        std::memcpy(&used, h_ctr + 32, 8);                  // (read back with the level counters: final after the last level)
        const size_t NCmax = (size_t)std::min<unsigned long long>(used, rr_.cand_cap);
        if (used > rr_.cand_cap) r_cand_hint_ = (size_t)used + (size_t)used / 4;          // (the windows that did not fit are searched on demand)
        // one device block laid out like the pinned block the host reads: coords | slen | hashes | wins | k | lon | sp | fwd
        auto al64 = [](size_t x) { return (x + 63) & ~(size_t)63; };
        size_t off[14];
        off[0] = 0;
        off[1] = off[0] + al64(NR * 2 * (size_t)n * 8);
        off[2] = off[1] + al64(NR * 8);
        off[3] = off[2] + al64(NR * 8);
        off[4] = off[3] + al64(NR * sizeof(WindowRec));
        off[5] = off[4] + al64(NCmax * 4);
        off[6] = off[5] + al64(NCmax * 4);
        off[7] = off[6] + al64(NCmax * (size_t)nq * 4);
        off[8] = off[7] + al64(NCmax * (size_t)nq);          // flags
        off[9] = off[8] + al64(NR * 4);                      // parent
        off[10] = off[9] + al64(NR * 4);                     // acc_shift
        off[11] = off[10] + al64(NCmax * 4);                 // acc_len
        off[12] = off[11] + al64(NCmax * 4);                 // foreign writes
        const size_t NFW = std::min<size_t>(h_ctr[19], rec::FW_CAP);
        off[13] = off[12] + al64(NFW * 12);
        uint8_t* d_out = r_out_.ensure(off[13] + 64, false, st_);
        pb200::launch(rec::gather_sorted_kernel, (unsigned)((NR * 32 + 255) / 256), 256, 0, st_, rr_.St, n, perm, cnt, nr32, rr_.d_k, rr_.d_lon, rr_.d_sp, rr_.d_fw,
                      (int64_t*)(d_out + off[0]), (int64_t*)(d_out + off[1]), (WindowRec*)(d_out + off[3]), (uint64_t*)(d_out + off[2]),
                      (int32_t*)(d_out + off[4]), (int32_t*)(d_out + off[5]), (int32_t*)(d_out + off[6]), d_out + off[7],
                      (const uint32_t*)inv, (uint32_t*)(d_out + off[8]), (int32_t*)(d_out + off[9]), (int32_t*)(d_out + off[10]), (int32_t*)(d_out + off[11]));
        if (NFW) PB_CUDA(cudaMemcpyAsync(d_out + off[12], rr_.Q.fw, NFW * 12, cudaMemcpyDeviceToDevice, st_));
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaMemcpyAsync(h_ctr, rr_.ctr, 32 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st_));
        uint8_t* stage = r_pin_stage_.ensure(off[13] + 64);
        PB_CUDA(cudaMemcpyAsync(stage, d_out, off[13], cudaMemcpyDeviceToHost, st_));
        PB_CUDA(cudaStreamSynchronize(st_));
        const double t1 = wall_s();
        const size_t NC = std::min<size_t>(h_ctr[24], NCmax);
        out.nregions = NR; out.ncands = NC;
        out.coords = (const int64_t*)(stage + off[0]); out.slen = (const int64_t*)(stage + off[1]); out.hashes = (const uint64_t*)(stage + off[2]);
        out.wins = (const WindowRec*)(stage + off[3]); out.k = (const int32_t*)(stage + off[4]); out.lon = (const int32_t*)(stage + off[5]);
        out.sp = (const int32_t*)(stage + off[6]); out.fwd = stage + off[7];
        out.flags = (const uint32_t*)(stage + off[8]); out.parent = (const int32_t*)(stage + off[9]);
        out.acc_shift = (const int32_t*)(stage + off[10]); out.acc_len = (const int32_t*)(stage + off[11]);
        out.fw = (const int32_t*)(stage + off[12]); out.nfw = h_ctr[19]; out.fw_cap = rec::FW_CAP;
        const double t2 = wall_s();
        out.levels = rr_.level; out.deferred = h_ctr[17]; out.dropped = h_ctr[18];
        // statistics: searched windows, their reference / query bases (bench.py's algorithmic-byte model)
        int64_t searched = 0, rb = 0, qb = 0, cands = 0;
        {
            const long per = 4096;
            const long nb = ((long)NR + per - 1) / per;
            std::vector<int64_t> acc((size_t)nb * 4 + 4, 0);
            parallel_chunks(NR > 16384 ? default_host_threads() : 1, nb, [&](long c) {
                int64_t s = 0, r_ = 0, q_ = 0, k_ = 0;
                for (size_t r = (size_t)c * per; r < std::min(NR, (size_t)(c + 1) * per); ++r) {
                    if (out.wins[r].ncand < 0) continue;
                    ++s; k_ += out.wins[r].ncand;
                    const int64_t* cc = out.coords + r * 2 * (size_t)n;
                    r_ += cc[n] - cc[0];
                    for (int g = 1; g < n; ++g) q_ += cc[n + g] - cc[g];
                }
                acc[(size_t)c * 4] = s; acc[(size_t)c * 4 + 1] = r_; acc[(size_t)c * 4 + 2] = q_; acc[(size_t)c * 4 + 3] = k_;
            });
            for (long c = 0; c < nb; ++c) { searched += acc[(size_t)c * 4]; rb += acc[(size_t)c * 4 + 1]; qb += acc[(size_t)c * 4 + 2]; cands += acc[(size_t)c * 4 + 3]; }
        }
        out.searched = searched;
        small_windows += searched; small_ref_bases += rb; small_query_bases += qb;
        if (searched >= 1024) {
            const uint32_t r16 = (uint32_t)std::min<int64_t>(1 << 20, cands * 20 / searched + 16);
            uint32_t cur = s_cand_per_region_x16_.load();
            while (r16 > cur && !s_cand_per_region_x16_.compare_exchange_weak(cur, r16)) {}
        }
        host_rec_device_s += t1 - rr_.t0; host_rec_d2h_s += t2 - t1; host_rec_copy_s += wall_s() - t2;
        rr_.active = false;
        return true;
    }

    // ---- the anchor stage on the device (cuda/anchors.cuh)
    int anchor_stage(const AnchorRequest& rq, AnchorResult& out) override {
        PB_CUDA(cudaSetDevice(device_));
        const int n = n_, nq = n - 1;
        if (!rec_supported(rq.n, rq.q) || rq.ntasks <= 0 || getenv("PB200_NO_DEVICE_ANCHORS")) return 0;
        for (int t = 0; t < rq.ntasks; ++t) if (classify(rq.tasks[t], rq.coords) != 3) return 0;      // (tiny genomes: the batch path)
        const double t0 = wall_s();
        // ---- candidates of every reference window, left on the device: k (genome-global), lon, sp, fwd
        std::vector<uint32_t> wcount((size_t)rq.ntasks, 0);
        size_t NCA = 0;
        for (int t = 0; t < rq.ntasks; ++t) {
            const WindowTask& wt = rq.tasks[t];
            const int64_t* qs = rq.coords + wt.coord_off;
            const int64_t* ql = qs + nq;
            std::vector<big::StrandDesc> sd((size_t)2 * nq);
            for (int q = 0; q < nq; ++q) {
                const int g = q + 1;
                sd[2 * q] = big::StrandDesc{text_.get() + gfwd_[g] + qs[q], (int32_t)ql[q], 0};
                sd[2 * q + 1] = big::StrandDesc{text_.get() + grc_[g] + (len_[g] - qs[q] - ql[q]), (int32_t)ql[q], 0};
            }
            const uint8_t* R = text_.get() + gfwd_[0] + wt.ref_start;
            big_.build_index(R, (int)wt.ref_len, wt.minsize, st_, window_is_n_free(wt.ref_start, wt.ref_len));
            big_.scan_events(R, (int)wt.ref_len, nq, sd, wt.minsize, st_);
            big_.fold(true, st_);
            big_.emit(st_);
            const big::BigPath::DeviceCands dc = big_.pass2_device(nullptr, st_);
            big_windows++;
            big_ref_bases += wt.ref_len;
            for (int q = 0; q < nq; ++q) big_query_bases += ql[q];
            big_events += big_.last_events;
            index_rounds += big_.last_index.rounds;
            wcount[(size_t)t] = dc.ncand;
            if (dc.ncand) {
                int32_t* ak = a_k_.ensure(NCA + dc.ncand, true, st_);
                int32_t* al = a_lon_.ensure(NCA + dc.ncand, true, st_);
                int32_t* as = a_sp_.ensure((NCA + dc.ncand) * (size_t)nq, true, st_);
                uint8_t* af = a_fwd_.ensure((NCA + dc.ncand) * (size_t)nq, true, st_);
                pb200::launch(anc::add_offset_kernel, (dc.ncand + 255) / 256, 256, 0, st_, ak + NCA, dc.k, dc.ncand, (int32_t)0);
                PB_CUDA(cudaMemcpyAsync(al + NCA, dc.lon, (size_t)dc.ncand * 4, cudaMemcpyDeviceToDevice, st_));
                PB_CUDA(cudaMemcpyAsync(as + NCA * (size_t)nq, dc.sp, (size_t)dc.ncand * nq * 4, cudaMemcpyDeviceToDevice, st_));
                PB_CUDA(cudaMemcpyAsync(af + NCA * (size_t)nq, dc.fwd, (size_t)dc.ncand * nq, cudaMemcpyDeviceToDevice, st_));
                NCA += dc.ncand;
            }
        }
        out.ncand = NCA;
        host_anchor_search_s += wall_s() - t0;
        if (NCA == 0) { out.status = 1; out.nanchors = 0; out.nregions = 0; rr_.active = false; return 1; }      // (NO MUMS FOUND)
        // ---- coordinates (k made genome-global per window), collinearity + overlap sweep
        const unsigned int nca = (unsigned int)NCA;
        int32_t* ST = a_st_.ensure(2 * NCA * (size_t)n + 64, false, st_);
        int32_t* STT = ST + NCA * (size_t)n;
        uint8_t* valid = a_valid_.ensure(NCA + 64, false, st_);
        anc::Flags* d_flags = reinterpret_cast<anc::Flags*>(a_flags_.ensure(8, false, st_));
        PB_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(anc::Flags), st_));
        {
            size_t base = 0;
            for (int t = 0; t < rq.ntasks; ++t) {       // window-relative k -> genome position
                if (wcount[(size_t)t] && rq.tasks[t].ref_start)
                    pb200::launch(anc::add_offset_kernel, (wcount[(size_t)t] + 255) / 256, 256, 0, st_, a_k_.get() + base, reinterpret_cast<const uint32_t*>(a_k_.get() + base),
                                  wcount[(size_t)t], (int32_t)rq.tasks[t].ref_start);
                base += wcount[(size_t)t];
            }
        }
        const int64_t* d_glen = gmeta_.get() + 2 * n_;
        pb200::launch(anc::anchor_coords_kernel, (unsigned)((NCA * 32 + 255) / 256), 256, 0, st_, n, d_glen, nca, a_k_.get(), a_lon_.get(), a_sp_.get(), a_fwd_.get(),
                      ST, STT, valid);
        pb200::launch(anc::anchor_overlap_kernel, n, 1024, 0, st_, nca, STT, a_lon_.get(), valid, d_flags);
        anc::Flags* h_flags = reinterpret_cast<anc::Flags*>(a_pin_flags_.ensure(8));
        PB_CUDA(cudaMemcpyAsync(h_flags, d_flags, sizeof(anc::Flags), cudaMemcpyDeviceToHost, st_));
        PB_CUDA(cudaStreamSynchronize(st_));
        if (h_flags->noncollinear || h_flags->overlaps) {
            // the premise of the parallel accept does not hold everywhere: the host gets the candidates (window-relative k, as
            // search() delivers them) and runs its own accept pass
            CandBatch& cb = out.cands;
            cb.clear();
            cb.nq = nq;
            cb.off.assign((size_t)rq.ntasks + 1, 0);
            for (int t = 0; t < rq.ntasks; ++t) cb.off[(size_t)t + 1] = cb.off[(size_t)t] + wcount[(size_t)t];
            cb.k.resize(NCA); cb.lon.resize(NCA); cb.sp.resize(NCA * (size_t)nq); cb.fwd.resize(NCA * (size_t)nq);
            PB_CUDA(cudaMemcpyAsync(cb.k.data(), a_k_.get(), NCA * 4, cudaMemcpyDeviceToHost, st_));
            PB_CUDA(cudaMemcpyAsync(cb.lon.data(), a_lon_.get(), NCA * 4, cudaMemcpyDeviceToHost, st_));
            PB_CUDA(cudaMemcpyAsync(cb.sp.data(), a_sp_.get(), NCA * (size_t)nq * 4, cudaMemcpyDeviceToHost, st_));
            PB_CUDA(cudaMemcpyAsync(cb.fwd.data(), a_fwd_.get(), NCA * (size_t)nq, cudaMemcpyDeviceToHost, st_));
            PB_CUDA(cudaStreamSynchronize(st_));
            size_t base = 0;
            for (int t = 0; t < rq.ntasks; ++t) {
                const int32_t o = (int32_t)rq.tasks[t].ref_start;
                if (o) for (size_t c = base; c < base + wcount[(size_t)t]; ++c) cb.k[c] -= o;
                base += wcount[(size_t)t];
            }
            out.status = 2;
            rr_.active = false;
            return 2;
        }
        // ---- accept on the empty layout, regions between the anchors, push rules -> the recursion's store
        rec_setup(2 * NCA, rq.layout_words, rq.minsize_tab, rq.minsize_n, rq.q, rq.p);
        unsigned long long* bits = rr_.P.bits;
        const int64_t* d_bit_off = rr_.P.bit_off;
        PB_CUDA(cudaMemsetAsync(bits, 0, (size_t)r_bits_words_ * 8, st_));
        pb200::launch(anc::layout_sentinels_kernel, (n + 127) / 128, 128, 0, st_, n, d_glen, bits, d_bit_off);
        uint32_t* accepted = a_u32_.ensure(4 * NCA + 4 * 2 * NCA + 256, false, st_);
        uint32_t* xidx = accepted + NCA;
        uint32_t* slot = xidx + NCA;                         // [2 * anchors] <= 2 NCA
        uint32_t* pos = slot + 2 * NCA;
        uint32_t* d_tot = pos + 2 * NCA;                     // [0] anchors, [1] regions
        pb200::launch(anc::anchor_accept_kernel, (unsigned)((NCA * 32 + 255) / 256), 256, 0, st_, n, text_.get(), gmeta_.get(), nca, ST, a_lon_.get(), a_fwd_.get(),
                      valid, bits, d_bit_off, accepted);
        r_scanner_.scan<prim::OpSum, true>(accepted, xidx, (int64_t)NCA, d_tot, st_);
        // (anchors <= candidates: everything per anchor is sized by NCA, no synchronisation needed to continue)
        auto al64 = [](size_t x) { return (x + 63) & ~(size_t)63; };
        int32_t* REG = a_reg_.ensure(NCA * 4 * (size_t)n + 2 * NCA + 64, false, st_);
        int32_t* SL = REG + NCA * 4 * (size_t)n;
        size_t hoff[6];
        hoff[0] = 0;                                          // anchors' starts
        hoff[1] = hoff[0] + al64(NCA * (size_t)n * 4);       // lon
        hoff[2] = hoff[1] + al64(NCA * 4);                   // fwd
        hoff[3] = hoff[2] + al64(NCA * (size_t)n);           // layout
        hoff[4] = hoff[3] + al64((size_t)r_bits_words_ * 8); // flags
        hoff[5] = hoff[4] + 64;
        uint8_t* d_host = a_out_.ensure(hoff[5] + 64, false, st_);
        pb200::launch(anc::anchor_regions_kernel, (unsigned)((NCA * 32 + 255) / 256), 256, 0, st_, n, d_glen, nca, accepted, xidx, ST, a_lon_.get(), a_fwd_.get(),
                      (const unsigned long long*)bits, d_bit_off, REG, SL, (int32_t*)(d_host + hoff[0]), (int32_t*)(d_host + hoff[1]), d_host + hoff[2]);
        int32_t* d_lo = a_lo_.ensure((size_t)rr_.cap * (size_t)n + 64, false, st_);
        // push flags need the number of anchors: it is on the device; the kernels are launched for NCA and bounded there
        pb200::launch(anchor_push_flags_bounded, (unsigned)((NCA * 32 + 255) / 256), 256, 0, st_, n, rq.q, (const uint32_t*)d_tot, nca, (const int32_t*)REG, (const int32_t*)SL, slot);
        r_scanner_.scan<prim::OpSum, true>(slot, pos, (int64_t)(2 * NCA), d_tot + 1, st_);
        pb200::launch(anchor_push_bounded, (unsigned)((NCA * 32 + 255) / 256), 256, 0, st_, n, (const uint32_t*)d_tot, (const int32_t*)REG, (const int32_t*)SL,
                      (const uint32_t*)slot, (const uint32_t*)pos, rr_.St, r_pairflag_, rr_.Q.cap, (const unsigned long long*)bits, d_bit_off, d_lo);
        pb200::launch(anc::anchor_finish_kernel, 1, 32, 0, st_, (const uint32_t*)d_tot, (const uint32_t*)(d_tot + 1), d_flags, rr_.Q.nregions, rr_.Q.cap);
        // ---- what the host needs: the layout BEFORE the recursion scribbles on it, anchors, flags; then the initial regions
        PB_CUDA(cudaMemcpyAsync(d_host + hoff[3], bits, (size_t)r_bits_words_ * 8, cudaMemcpyDeviceToDevice, st_));
        PB_CUDA(cudaMemcpyAsync(d_host + hoff[4], d_flags, sizeof(anc::Flags), cudaMemcpyDeviceToDevice, st_));
        uint8_t* h_host = a_pin_out_.ensure(hoff[5] + 64);
        PB_CUDA(cudaMemcpyAsync(h_host + hoff[4], d_host + hoff[4], 64, cudaMemcpyDeviceToHost, st_));
        PB_CUDA(cudaStreamSynchronize(st_));                 // (the counts: how much of the rest to copy)
        const anc::Flags fl = *reinterpret_cast<const anc::Flags*>(h_host + hoff[4]);
        const size_t NA = fl.accepted, NR = std::min<size_t>(fl.regions, rr_.cap);
        if (getenv("PB200_DEBUG_ANCHORS")) fprintf(stderr, "[pb200 anchors] windows %d candidates %zu accepted %u regions %u noncollinear %u overlaps %u\n", rq.ntasks, NCA, fl.accepted, fl.regions, fl.noncollinear, fl.overlaps);
        if (rq.follow_recursion && NR > 0) {
            rec_seed(2 * NCA);
            rec_launch_round();                              // the GPU follows the recursion while the host builds its pools
        } else rr_.active = false;
        // copies on a second stream: they overlap the recursion kernels (the store's initial regions are only read by them)
        if (!st2_) PB_CUDA(cudaStreamCreateWithFlags(&st2_, cudaStreamNonBlocking));
        int32_t* h_rc = a_pin_rc_.ensure(NR * 2 * (size_t)n + 16);
        PB_CUDA(cudaMemcpyAsync(h_host + hoff[0], d_host + hoff[0], NA * (size_t)n * 4, cudaMemcpyDeviceToHost, st2_));
        PB_CUDA(cudaMemcpyAsync(h_host + hoff[1], d_host + hoff[1], NA * 4, cudaMemcpyDeviceToHost, st2_));
        PB_CUDA(cudaMemcpyAsync(h_host + hoff[2], d_host + hoff[2], NA * (size_t)n, cudaMemcpyDeviceToHost, st2_));
        PB_CUDA(cudaMemcpyAsync(h_host + hoff[3], d_host + hoff[3], (size_t)r_bits_words_ * 8, cudaMemcpyDeviceToHost, st2_));
        int32_t* h_lo = a_pin_lo_.ensure(NR * (size_t)n + 16);
        if (NR) PB_CUDA(cudaMemcpyAsync(h_rc, rr_.St.coords, NR * 2 * (size_t)n * 4, cudaMemcpyDeviceToHost, st2_));
        if (NR) PB_CUDA(cudaMemcpyAsync(h_lo, d_lo, NR * (size_t)n * 4, cudaMemcpyDeviceToHost, st2_));
        PB_CUDA(cudaStreamSynchronize(st2_));
        out.status = 1;
        out.nanchors = NA; out.nregions = NR;
        out.a_start = (const int32_t*)(h_host + hoff[0]); out.a_lon = (const int32_t*)(h_host + hoff[1]); out.a_fwd = h_host + hoff[2];
        out.layout = (const uint64_t*)(h_host + hoff[3]);
        out.layout_off = r_bit_off_host_;
        out.r_coords = h_rc;
        out.r_lo = h_lo;
        host_anchor_accept_s += wall_s() - t0;
        return 1;
    }

    // ---- StagedWindowEngine (query-sharded large windows, see host/sharded.h) ----
    int classify(const WindowTask& t, const int64_t* coords) const {
        const int nq = n_ - 1;
        const int64_t* ql = coords + t.coord_off + nq;
        int64_t maxm = 0;
        for (int q = 0; q < nq; ++q) maxm = std::max(maxm, ql[q]);
        int c = 3;
        for (int k = 0; k < 3; ++k)
            if (t.ref_len <= classes_[k].n_cap && maxm <= classes_[k].m_cap) { c = k; break; }
        if (t.minsize < 4) c = 3;          // the shared-memory path seeds with 4-base matches
        if (force_big_) c = 3;
        return c;
    }
    bool wants_staged(const WindowTask& t, const int64_t* coords) override { return classify(t, coords) == 3; }
    bool buffers_on_device() const override { return true; }
    void window_begin(const WindowTask& t, const int64_t* coords, bool build_index) override {
        PB_CUDA(cudaSetDevice(device_));
        const int nq = n_ - 1;
        w_task_ = t;
        w_R_ = text_.get() + gfwd_[0] + t.ref_start;
        const int64_t* qs = coords + t.coord_off;
        const int64_t* ql = qs + nq;
        w_strands_.resize((size_t)2 * nq);
        for (int q = 0; q < nq; ++q) {
            const int g = q + 1;
            w_strands_[2 * q] = big::StrandDesc{text_.get() + gfwd_[g] + qs[q], (int32_t)ql[q], 0};
            w_strands_[2 * q + 1] = big::StrandDesc{text_.get() + grc_[g] + (len_[g] - qs[q] - ql[q]), (int32_t)ql[q], 0};
        }
        if (build_index) { big_.build_index(w_R_, (int)t.ref_len, t.minsize, st_, window_is_n_free(t.ref_start, t.ref_len)); big_.ensure_index(st_); }
        else big_.alloc_index((int)t.ref_len, t.minsize, st_);
        PB_CUDA(cudaStreamSynchronize(st_));
        big_windows++;
    }
    void window_index_buffers(std::vector<std::pair<void*, size_t>>& bufs) override {
        const size_t n = (size_t)w_task_.ref_len;
        bufs.emplace_back((void*)big_.index_sa(), n * 4);
        bufs.emplace_back((void*)big_.index_lrp(), n * 4);
        bufs.emplace_back((void*)big_.index_table(), big_.index_table_entries() * sizeof(uint2));
    }
    int window_n() const override { return (int)w_task_.ref_len; }
    void window_scan(int q0, int q1) override {
        w_q0_ = q0; w_q1_ = q1;
        std::vector<big::StrandDesc> sd(w_strands_.begin() + 2 * q0, w_strands_.begin() + 2 * q1);
        big_.scan_events(w_R_, (int)w_task_.ref_len, q1 - q0, sd, w_task_.minsize, st_);
    }
    void window_fold(bool init) override { big_.fold(init, st_); PB_CUDA(cudaStreamSynchronize(st_)); }
    int32_t* window_master_up() override { return big_.master_up(st_); }
    int32_t* window_master_ep() override { return big_.master_ep(st_); }
    int32_t* window_gather_buffer(size_t ints) override { int32_t* p = w_gather_.ensure(ints, false, st_); PB_CUDA(cudaStreamSynchronize(st_)); return p; }
    void window_apply_prefix(const int32_t* gathered, int world, int rank) override {
        const int n = (int)w_task_.ref_len;
        int32_t* init = w_init_.ensure((size_t)n, false, st_);
        pb200::launch(prefix_min_kernel, (unsigned)((n + 255) / 256), 256, 0, st_, gathered, world, rank, n, init, big_.master_up(st_), big_.master_ep(st_));
    }
    uint32_t window_emit() override { return big_.emit(st_); }
    void window_pass2(std::vector<int32_t>& k, std::vector<int32_t>& lon, std::vector<int32_t>& sp, std::vector<uint8_t>& fwd) override {
        big_.pass2(w_init_.get(), st_, k, lon, sp, fwd);
    }
    cudaStream_t stream() const { return st_; }

    // ---- MUMi mode (Aligner::setMumi, src/parsnp.cpp:1869-2115): per query 1 - covered/len over the FIRST reference window
    void mumi(int64_t P, std::vector<double>& out) {
        PB_CUDA(cudaSetDevice(device_));
        const int nq = n_ - 1;
        out.assign(nq, 0.0);
        const int64_t L0 = len_[0];
        const int64_t p = P > L0 ? L0 : P;                 // src/parsnp.cpp:1903-1906; only partpos = 0 is ever processed (1907)
        if (p <= 0 || nq <= 0) return;
        std::vector<big::StrandDesc> sd((size_t)2 * nq);
        for (int q = 0; q < nq; ++q) {
            const int g = q + 1;
            sd[2 * q] = big::StrandDesc{text_.get() + gfwd_[g], (int32_t)len_[g], 0};
            sd[2 * q + 1] = big::StrandDesc{text_.get() + grc_[g], (int32_t)len_[g], 0};
        }
        const int minlon = 15;                             // only MUMs of >= 15 bp count (src/parsnp.cpp:2050-2062)
        big_.build_index(text_.get() + gfwd_[0], (int)p, minlon, st_, window_is_n_free(0, p));
        big_.scan_events(text_.get() + gfwd_[0], (int)p, nq, sd, minlon, st_);
        for (int q = 0; q < nq; ++q) {
            int total = (int)big_.mumi_covered(q, st_);
            const int minlen = (int)p;
            const float ratio = float(L0) / float(len_[q + 1]);
            if (ratio > 1.3 || ratio < 0.7) total = 0;     // src/parsnp.cpp:2074-2075
            if (total > minlen) total = minlen;
            out[q] = 1.0 - (float(total) / float(minlen));
        }
    }

    // test hooks: suffix array + lrp of a window of genome 0
    void debug_index(int64_t ref_start, int n, int minsize, uint32_t* sa, int32_t* lrp) {
        PB_CUDA(cudaSetDevice(device_));
        big_.build_index(text_.get() + gfwd_[0] + ref_start, n, minsize, st_, window_is_n_free(ref_start, n));
        big_.ensure_index(st_);
        PB_CUDA(cudaMemcpyAsync(sa, big_.d_sa(), (size_t)n * 4, cudaMemcpyDeviceToHost, st_));
        PB_CUDA(cudaMemcpyAsync(lrp, big_.d_lrp(), (size_t)n * 4, cudaMemcpyDeviceToHost, st_));
        PB_CUDA(cudaStreamSynchronize(st_));
    }
    int debug_index_flags() const { return (big_.last_index.fallback ? 1 : 0) | (big_.last_index.two_bit ? 2 : 0); }
    void set_force_class(int c) { force_big_ = c; }
    void collect_timers() { cudaSetDevice(device_); timers.collect(st_); }
    bool window_is_n_free(int64_t start, int64_t len) const {
        auto it = std::lower_bound(ref_n_pos_.begin(), ref_n_pos_.end(), start);
        return (it == ref_n_pos_.end() || *it >= start + len) && !force_3bit_;
    }
    int n() const { return n_; }
    int device() const { return device_; }
    int64_t h2d_bytes() const { return h2d_bytes_; }
    GpuTimers timers;
    int64_t big_windows = 0, small_windows = 0, small_retries = 0, big_events = 0, index_rounds = 0;
    int64_t small_class_tasks[3] = {0, 0, 0};
    double host_rec_device_s = 0, host_rec_d2h_s = 0, host_rec_copy_s = 0, host_anchor_search_s = 0, host_anchor_accept_s = 0;
    double host_classify_s = 0, host_upload_s = 0, host_small_wait_s = 0, host_small_d2h_s = 0, host_big_s = 0, host_gather_s = 0;   // wall clock, host side
    int64_t small_ref_bases = 0, small_query_bases = 0, big_ref_bases = 0, big_query_bases = 0;

private:
    void upload_small_tasks(const WindowTask* tasks, int ntasks, const int64_t* coords, int64_t ncoords) {
        small::TaskDev* td = pin_tasks_.ensure((size_t)ntasks);        // pinned staging: the copies below run at link speed
        h_task_n_.assign(ntasks, 0);
        h_task_m_.assign(ntasks, 0);
        const int nq_ = n_ - 1;
        for (int t = 0; t < ntasks; ++t) {
            h_task_n_[t] = tasks[t].ref_len;
            const int64_t* ql = coords + tasks[t].coord_off + nq_;
            for (int q = 0; q < nq_; ++q) h_task_m_[t] += ql[q];
            td[t].ref_off = gfwd_[0] + tasks[t].ref_start;
            td[t].n = (int32_t)tasks[t].ref_len;
            td[t].minsize = tasks[t].minsize;
            td[t].qcoord_off = tasks[t].coord_off;
        }
        int32_t* qc = pin_qc_.ensure((size_t)std::max<int64_t>(ncoords, 1));
        for (int64_t i = 0; i < ncoords; ++i) qc[i] = (int32_t)coords[i];
        small::TaskDev* d_t = d_tasks_.ensure((size_t)ntasks, false, st_);
        int32_t* d_q = d_qcoords_.ensure((size_t)std::max<int64_t>(ncoords, 1), false, st_);
        d_outs_.ensure((size_t)ntasks, false, st_);
        PB_CUDA(cudaMemcpyAsync(d_t, td, sizeof(small::TaskDev) * ntasks, cudaMemcpyHostToDevice, st_));
        if (ncoords) PB_CUDA(cudaMemcpyAsync(d_q, qc, (size_t)ncoords * 4, cudaMemcpyHostToDevice, st_));
        PB_CUDA(cudaStreamSynchronize(st_));
    }

    void run_small(int c, const std::vector<int>& ids_in, int nq, std::vector<int64_t>& t_base, std::vector<int32_t>& t_cnt,
                   std::vector<int>& retry) {
        std::vector<int> ids = ids_in;
        small::ClassCfg cfg = classes_[c];
        cfg.ev_cap += 4 * nq;            // the event store holds all strands of all queries of a window
        if (cfg.ev_cap > 60000) cfg.ev_cap = 60000;
        for (int t : ids_in) { small_ref_bases += h_task_n_[t]; small_query_bases += h_task_m_[t]; }
        // candidates per window seen so far in this process (x1.25) sizes the shared output buffer; an overflow re-runs only the
        // windows that did not fit
        static std::atomic<uint32_t> s_cand_per_task_x16(8 * 16);
        size_t cand_cap = std::max<size_t>(cand_cap_hint_, (size_t)ids.size() * s_cand_per_task_x16.load() / 16 + 4096);
        const size_t cap_limit = std::max<size_t>((size_t)1 << 16, ((size_t)1 << 30) / (size_t)(8 + 5 * std::max(nq, 1)));   // <= 1 GiB of candidate arrays
        if (cand_cap > cap_limit) cand_cap = cap_limit;
        const size_t ntasks_in = ids.size();
        size_t cands_out = 0;
        while (!ids.empty()) {
            const int nt = (int)ids.size();
            int32_t* d_ids = d_ids_.ensure((size_t)nt, false, st_);
            PB_CUDA(cudaMemcpyAsync(d_ids, ids.data(), (size_t)nt * 4, cudaMemcpyHostToDevice, st_));
            unsigned long long* d_cnt = d_candcnt_.ensure(1, false, st_);
            PB_CUDA(cudaMemsetAsync(d_cnt, 0, 8, st_));
            int32_t* d_k = d_ck_.ensure(cand_cap, false, st_);
            int32_t* d_lon = d_clon_.ensure(cand_cap, false, st_);
            int32_t* d_sp = d_csp_.ensure(cand_cap * std::max(nq, 1), false, st_);
            uint8_t* d_fw = d_cfwd_.ensure(cand_cap * std::max(nq, 1), false, st_);
            timers.start(GpuTimers::T_SMALL + c, st_);
            pb200::launch(small::small_region_kernel, nt, cfg.threads, cfg.smem_bytes(nq), st_, 
                text_.get(), gmeta_.get(), gmeta_.get() + n_, gmeta_.get() + 2 * n_, nq, d_tasks_.get(), d_qcoords_.get(), d_ids, nt, cfg,
                d_outs_.get(), d_cnt, (unsigned long long)cand_cap, d_k, d_lon, d_sp, d_fw);
            PB_CUDA(cudaGetLastError());
            timers.stop(GpuTimers::T_SMALL + c, st_);
            small_class_tasks[c] += nt;
            unsigned long long used = 0;
            PB_CUDA(cudaMemcpyAsync(&used, d_cnt, 8, cudaMemcpyDeviceToHost, st_));
            // outs are indexed by task id; fetch the ones of this launch
            int maxid = *std::max_element(ids.begin(), ids.end());
            small::TaskOut* h_outs_all = pin_outs_.ensure((size_t)maxid + 1);
            PB_CUDA(cudaMemcpyAsync(h_outs_all, d_outs_.get(), sizeof(small::TaskOut) * ((size_t)maxid + 1), cudaMemcpyDeviceToHost, st_));
            const double tw0 = wall_s();
            PB_CUDA(cudaStreamSynchronize(st_));
            const double tw1 = wall_s();
            host_small_wait_s += tw1 - tw0;
            const size_t got = (size_t)std::min<unsigned long long>(used, cand_cap);
            cands_out += got;
            const size_t hb = small_cands_;
            small_cands_ = hb + got;
            int32_t* pk = pin_k_.ensure(hb + got, hb);
            int32_t* pl = pin_lon_.ensure(hb + got, hb);
            int32_t* ps = pin_sp_.ensure((hb + got) * std::max(nq, 1), hb * nq);
            uint8_t* pf = pin_fwd_.ensure((hb + got) * std::max(nq, 1), hb * nq);
            if (got) {
                PB_CUDA(cudaMemcpyAsync(pk + hb, d_k, got * 4, cudaMemcpyDeviceToHost, st_));
                PB_CUDA(cudaMemcpyAsync(pl + hb, d_lon, got * 4, cudaMemcpyDeviceToHost, st_));
                if (nq) {
                    PB_CUDA(cudaMemcpyAsync(ps + hb * nq, d_sp, got * nq * 4, cudaMemcpyDeviceToHost, st_));
                    PB_CUDA(cudaMemcpyAsync(pf + hb * nq, d_fw, got * nq, cudaMemcpyDeviceToHost, st_));
                }
                PB_CUDA(cudaStreamSynchronize(st_));
            }
            host_small_d2h_s += wall_s() - tw1;
            std::vector<int> again;
            for (int t : ids) {
                const small::TaskOut& o = h_outs_all[t];
                if (o.ncand >= 0) { t_cnt[t] = o.ncand; t_base[t] = (int64_t)hb + o.cand_base; small_windows++; }
                else if (o.ncand == -2) again.push_back(t);      // global candidate buffer full: same class, bigger buffer
                else { retry.push_back(t); small_retries++; }     // per-CTA event/candidate capacity: next class
            }
            ids.swap(again);
            if (!ids.empty()) { cand_cap = cand_cap * 2 + 4096; cand_cap_hint_ = cand_cap; }
        }
        if (ntasks_in >= 1024) {
            const uint32_t r = (uint32_t)std::min<size_t>(1u << 20, cands_out * 20 / ntasks_in + 16);     // x16 fixed point, +25 %
            uint32_t cur = s_cand_per_task_x16.load();
            while (r > cur && !s_cand_per_task_x16.compare_exchange_weak(cur, r)) {}
        }
    }

    void run_big(const WindowTask& t, const int64_t* coords, int nq) {
        if (t.ref_len >= ((int64_t)1 << 31) - 64) throw CudaError("window longer than 2^31");
        const int64_t* qs = coords + t.coord_off;
        const int64_t* ql = qs + nq;
        std::vector<big::StrandDesc> sd((size_t)2 * nq);
        for (int q = 0; q < nq; ++q) {
            const int g = q + 1;
            sd[2 * q].q = text_.get() + gfwd_[g] + qs[q];
            sd[2 * q].m = (int32_t)ql[q];
            sd[2 * q].pad = 0;
            sd[2 * q + 1].q = text_.get() + grc_[g] + (len_[g] - qs[q] - ql[q]);
            sd[2 * q + 1].m = (int32_t)ql[q];
            sd[2 * q + 1].pad = 0;
        }
        big_.search(text_.get() + gfwd_[0] + t.ref_start, (int)t.ref_len, nq, sd, t.minsize, st_, bg_k_, bg_lon_, bg_sp_, bg_fwd_,
                    window_is_n_free(t.ref_start, t.ref_len));
        big_windows++;
        big_ref_bases += t.ref_len;
        for (int q = 0; q < nq; ++q) big_query_bases += ql[q];
        big_events += big_.last_events;
        index_rounds += big_.last_index.rounds;
    }

    int device_ = 0, n_ = 0, sm_count_ = 148, force_big_ = 0;
    bool force_3bit_ = false;
    std::vector<int64_t> ref_n_pos_;
    cudaStream_t st_ = nullptr;
    std::vector<int64_t> len_, gfwd_, grc_;
    int64_t h2d_bytes_ = 0;
    DevBuf<uint8_t> text_, stage_, d_cfwd_;
    DevBuf<int64_t> gmeta_;
    DevBuf<small::TaskDev> d_tasks_;
    DevBuf<small::TaskOut> d_outs_;
    DevBuf<int32_t> d_qcoords_, d_ids_, d_ck_, d_clon_, d_csp_;
    DevBuf<unsigned long long> d_candcnt_;
    DevBuf<int32_t> w_gather_, w_init_;
    WindowTask w_task_{};
    const uint8_t* w_R_ = nullptr;
    std::vector<big::StrandDesc> w_strands_;
    int w_q0_ = 0, w_q1_ = 0;
    PinBuf<small::TaskOut> pin_outs_;
    PinBuf<small::TaskDev> pin_tasks_;
    PinBuf<int32_t> pin_qc_;
    std::vector<int64_t> h_task_n_, h_task_m_;
    small::ClassCfg classes_[3];
    size_t max_smem_ = 0, cand_cap_hint_ = 0;
    big::BigPath big_;
    // device recursion (discover_recursion)
    DevBuf<unsigned long long> r_bits_;
    DevBuf<int64_t> r_bitoff_, r_candbase_;
    DevBuf<int32_t> r_tab_, r_coords_, r_slen_, r_minsize_, r_ncand_, r_lists_;
    DevBuf<unsigned int> r_ctr_;
    PinBuf<int64_t> r_pin_bitoff_;
    PinBuf<int32_t> r_pin_tab_, r_pin_coords_;
    PinBuf<unsigned int> r_pin_ctr_;
    PinBuf<uint8_t> r_pin_stage_;
    DevBuf<uint8_t> r_pair_;
    uint8_t* r_pairflag_ = nullptr;
    DevBuf<uint32_t> r_sortk_, r_sortv_;
    DevBuf<uint8_t> r_out_;
    DevBuf<uint32_t> r_flags_, r_inv_;
    DevBuf<int32_t> r_parent_, r_fw_, r_accshift_, r_acclen_;
    rsort::RadixSorter r_sorter_;
    prim::Scanner r_scanner_;
    size_t r_cand_hint_ = 0;
    // anchor stage on the device (anchor_stage)
    DevBuf<int32_t> a_k_, a_lon_, a_sp_, a_st_, a_reg_, a_lo_;
    DevBuf<uint8_t> a_fwd_, a_valid_, a_out_;
    DevBuf<uint32_t> a_u32_;
    DevBuf<unsigned int> a_flags_;
    PinBuf<unsigned int> a_pin_flags_;
    PinBuf<uint8_t> a_pin_out_;
    PinBuf<int32_t> a_pin_rc_, a_pin_lo_;
    cudaStream_t st2_ = nullptr;
    std::vector<int64_t> r_bit_off_host_;
    static std::atomic<uint32_t> s_cand_per_region_x16_;
    int64_t r_bits_words_ = -1;
    bool r_attr_set_[8] = {false, false, false, false, false, false, false, false};
    PinBuf<int32_t> pin_k_, pin_lon_, pin_sp_;       // small-window candidates of the current search() call (pinned staging)
    PinBuf<uint8_t> pin_fwd_;
    size_t small_cands_ = 0;
    std::vector<int32_t> bg_k_, bg_lon_, bg_sp_;
    std::vector<uint8_t> bg_fwd_;
};

std::atomic<uint32_t> CudaEngine::s_cand_per_region_x16_(6 * 16);

}  // namespace pb200

// =====================================================================================================  C ABI
namespace {
// Comm over caller-supplied collectives; synchronises the engine stream before handing device pointers out
struct CallbackComm : public pb200::Comm {
    pb200_allgather_cb ag = nullptr; pb200_allreduce_cb ar = nullptr; pb200_bcast_cb bc = nullptr; void* user = nullptr;
    pb200::CudaEngine* eng = nullptr;
    void pre(bool device) { if (device && eng) { cudaSetDevice(eng->device()); cudaStreamSynchronize(eng->stream()); } }
    void allgather(const void* send, void* recv, size_t bytes, bool device) override {
        pre(device);
        if (ag(user, send, recv, (int64_t)bytes, device ? 1 : 0) != 0) throw std::runtime_error("allgather callback failed");
    }
    void allreduce_i32(int32_t* buf, size_t count, bool is_max, bool device) override {
        pre(device);
        if (ar(user, buf, (int64_t)count, is_max ? 1 : 0, device ? 1 : 0) != 0) throw std::runtime_error("allreduce callback failed");
    }
    void bcast(void* buf, size_t bytes, int root, bool device) override {
        pre(device);
        if (bc(user, buf, (int64_t)bytes, root, device ? 1 : 0) != 0) throw std::runtime_error("bcast callback failed");
    }
};
}  // namespace

struct pb200_genomes {
    std::unique_ptr<pb200::CudaEngine> eng;
    std::unique_ptr<CallbackComm> comm;
    bool bcast_index = true;
    int n = 0;
    std::vector<const uint8_t*> seq;
    std::vector<int64_t> len;
};

namespace {
// One alignment allocates and frees ~100 MB of host vectors (candidate blocks, MUM / region pools, result arrays).  With
// glibc's defaults every block above the (dynamic) mmap threshold is mapped fresh and unmapped again: tens of thousands of
// page faults per alignment, and with one process per GPU they contend in the kernel.  Keep such blocks on the heap instead
// (PB200_KEEP_MALLOC_DEFAULTS=1 leaves the allocator alone).
void tune_host_allocator() {
    static std::once_flag once;
    std::call_once(once, [] {
        if (getenv("PB200_KEEP_MALLOC_DEFAULTS")) return;
        mallopt(M_MMAP_THRESHOLD, 32 * 1024 * 1024);       // the largest value glibc accepts
        mallopt(M_TRIM_THRESHOLD, 0x7fffffff);             // do not give the top of the heap back between alignments
    });
}

template <class F>
int guarded(F&& f) {
    try { return f(); }
    catch (const pb200::CudaError& e) {
        pb200::g_last_error = e.what();
        std::string s = e.what();
        return (s.find("requires a CUDA device") != std::string::npos || s.find("sm_100a") != std::string::npos) ? PB200_ERR_NO_CUDA : PB200_ERR_CUDA;
    }
    catch (const std::invalid_argument& e) { pb200::g_last_error = e.what(); return PB200_ERR_ARG; }
    catch (const std::exception& e) { pb200::g_last_error = e.what(); return PB200_ERR_INTERNAL; }
}
}  // namespace

extern "C" {

const char* pb200_version(void) { return "parsnp_b200 0.1 (sm_100a)"; }

int pb200_cuda_available(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return 0; }
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) return 0;
    return p.major >= 10 ? 1 : 0;
}

int pb200_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

int pb200_genomes_create(int device, int n, const uint8_t* const* seqs, const int64_t* lens, pb200_genomes** out) {
    return guarded([&]() {
        if (n < 1 || !seqs || !lens || !out) { pb200::g_last_error = "bad arguments"; return (int)PB200_ERR_ARG; }
        const bool prof = getenv("PB200_PROFILE_HOST") != nullptr;
        tune_host_allocator();
        pb200::install_backtrace_handler();
        const double t0 = pb200::wall_s();
        std::unique_ptr<pb200_genomes> g(new pb200_genomes);
        g->eng.reset(new pb200::CudaEngine(device));
        const double t1 = pb200::wall_s();
        g->n = n;
        g->seq.assign(seqs, seqs + n);
        g->len.assign(lens, lens + n);
        g->eng->set_genomes(n, seqs, lens);
        if (prof) fprintf(stderr, "[pb200 create ms] engine %.2f set_genomes %.2f\n", (t1 - t0) * 1e3, (pb200::wall_s() - t1) * 1e3);
        *out = g.release();
        return (int)PB200_OK;
    });
}
void pb200_genomes_free(pb200_genomes* g) { delete g; }

int pb200_search_windows(pb200_genomes* g, int ntasks, const pb200_window* tasks, const int64_t* coords, int64_t ncoords,
                         int64_t** cand_off, int32_t** k, int32_t** lon, int32_t** sp, uint8_t** fwd) {
    return guarded([&]() {
        (void)ncoords;
        static_assert(sizeof(pb200_window) == sizeof(pb200::WindowTask), "window layout");
        pb200::CandBatch cb;
        g->eng->search(reinterpret_cast<const pb200::WindowTask*>(tasks), ntasks, coords, cb);
        cb.compact(ntasks);
        auto dup = [](const void* p, size_t bytes) { void* q = malloc(bytes ? bytes : 1); if (bytes) memcpy(q, p, bytes); return q; };
        *cand_off = (int64_t*)dup(cb.off.data(), cb.off.size() * 8);
        *k = (int32_t*)dup(cb.k.data(), cb.k.size() * 4);
        *lon = (int32_t*)dup(cb.lon.data(), cb.lon.size() * 4);
        *sp = (int32_t*)dup(cb.sp.data(), cb.sp.size() * 4);
        *fwd = (uint8_t*)dup(cb.fwd.data(), cb.fwd.size());
        return (int)PB200_OK;
    });
}

// the engine re-uses the resident genomes; the Aligner's set_genomes call is a no-op for it
namespace {
class ResidentBackend : public pb200::SearchBackend {
public:
    explicit ResidentBackend(pb200::CudaEngine* e) : e_(e) {}
    void set_genomes(int, const uint8_t* const*, const int64_t*) override {}
    void search(const pb200::WindowTask* t, int nt, const int64_t* c, pb200::CandBatch& out) override { e_->search(t, nt, c, out); }
    bool discover_recursion(const pb200::RecursionRequest& rq, pb200::RecursionResult& out) override { return e_->discover_recursion(rq, out); }
    int anchor_stage(const pb200::AnchorRequest& rq, pb200::AnchorResult& out) override { return e_->anchor_stage(rq, out); }
private:
    pb200::CudaEngine* e_;
};
}  // namespace

int pb200_align_resident(pb200_genomes* g, const pb200_params* prm, pb200_result** out) {
    return guarded([&]() {
        if (!g || !prm || !out) { pb200::g_last_error = "bad arguments"; return (int)PB200_ERR_ARG; }
        ResidentBackend local(g->eng.get());
        std::unique_ptr<pb200::ShardedBackend> sharded;
        pb200::SearchBackend* bep = &local;
        if (g->comm && g->comm->world > 1) {
            sharded.reset(new pb200::ShardedBackend(&local, g->eng.get(), g->comm.get(), g->bcast_index));
            sharded->set_n(g->n);
            bep = sharded.get();
        }
        pb200::SearchBackend& be = *bep;
        const bool prof = getenv("PB200_PROFILE_HOST") != nullptr;
        double tp[5] = {pb200::wall_s(), 0, 0, 0, 0};
        bool ok = false;
        {
            pb200::Aligner a(g->n, g->seq.data(), g->len.data(), pb200::to_align_params(prm), &be);
            a.enable_trace(prm->flags & PB200_FLAG_TRACE_WINDOWS);
            a.set_speculate(!(prm->flags & PB200_FLAG_NO_SPECULATION));
            a.set_threads(pb200::default_host_threads());
            a.set_pipeline(!sharded);           // collectives inside the search: every rank must issue them in the same order
            tp[1] = pb200::wall_s();
            ok = a.run();
            tp[2] = pb200::wall_s();
            *out = pb200::make_result(a, (prm->flags & PB200_FLAG_UNALIGNED) != 0);
            tp[3] = pb200::wall_s();
        }
        tp[4] = pb200::wall_s();
        if (prof) fprintf(stderr, "[pb200 align ms] ctor %.2f run %.2f make_result %.2f dtor %.2f\n", (tp[1] - tp[0]) * 1e3, (tp[2] - tp[1]) * 1e3,
                          (tp[3] - tp[2]) * 1e3, (tp[4] - tp[3]) * 1e3);
        return ok ? (int)PB200_OK : (int)PB200_ERR_NO_MUMS;
    });
}

int pb200_align(int device, int n, const uint8_t* const* seqs, const int64_t* lens, const pb200_params* prm, pb200_result** out) {
    pb200_genomes* g = nullptr;
    int rc = pb200_genomes_create(device, n, seqs, lens, &g);
    if (rc != 0) return rc;
    rc = pb200_align_resident(g, prm, out);
    pb200_genomes_free(g);
    return rc;
}

int pb200_mumi(pb200_genomes* g, const pb200_params* prm, double* out) {
    return guarded([&]() {
        if (!g || !prm || !out) { pb200::g_last_error = "bad arguments"; return (int)PB200_ERR_ARG; }
        std::vector<double> v;
        g->eng->mumi(prm->p, v);
        for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
        return (int)PB200_OK;
    });
}

int pb200_engine_timers(pb200_genomes* g, double* values, int cap) {
    int k = 0;
    g->eng->collect_timers();
    const int T = pb200::GpuTimers::T_COUNT;
    for (int i = 0; i < T && k < cap; ++i) values[k++] = g->eng->timers.ms[i];
    for (int i = 0; i < T && k < cap; ++i) values[k++] = g->eng->timers.cnt[i];
    const double extra[22] = {(double)g->eng->small_class_tasks[0], (double)g->eng->small_class_tasks[1], (double)g->eng->small_class_tasks[2],(double)g->eng->big_windows, (double)g->eng->small_windows, (double)g->eng->small_retries,
                              (double)g->eng->big_events, (double)g->eng->index_rounds, (double)pb200::g_kernel_launches,
                              (double)g->eng->small_ref_bases, (double)g->eng->small_query_bases, (double)g->eng->big_ref_bases,
                              (double)g->eng->big_query_bases, g->eng->host_classify_s, g->eng->host_upload_s, g->eng->host_small_wait_s,
                              g->eng->host_small_d2h_s, g->eng->host_big_s, g->eng->host_gather_s, g->eng->host_rec_device_s, g->eng->host_rec_d2h_s,
                              g->eng->host_rec_copy_s};
    for (int i = 0; i < 22 && k < cap; ++i) values[k++] = extra[i];
    return k;
}
const char* pb200_engine_timer_names(void) {
    static std::string s = std::string(pb200::GpuTimers::names()) + ",tasks_class_a,tasks_class_b,tasks_class_c,big_windows,small_windows,small_retries,big_events,index_rounds,kernel_launches,small_ref_bases,small_query_bases,big_ref_bases,big_query_bases,host_classify_s,host_upload_s,host_small_wait_s,host_small_d2h_s,host_big_s,host_gather_s,host_rec_device_s,host_rec_d2h_s,host_rec_copy_s";
    return s.c_str();
}
void pb200_engine_reset_timers(pb200_genomes* g) {
    g->eng->collect_timers();
    g->eng->timers.reset();
    g->eng->big_windows = g->eng->small_windows = g->eng->small_retries = g->eng->big_events = g->eng->index_rounds = 0;
    pb200::g_kernel_launches = 0;
    g->eng->small_ref_bases = g->eng->small_query_bases = g->eng->big_ref_bases = g->eng->big_query_bases = 0;
    g->eng->small_class_tasks[0] = g->eng->small_class_tasks[1] = g->eng->small_class_tasks[2] = 0;
    g->eng->host_classify_s = g->eng->host_upload_s = g->eng->host_small_wait_s = g->eng->host_small_d2h_s = g->eng->host_big_s = g->eng->host_gather_s = 0;
    g->eng->host_rec_device_s = g->eng->host_rec_d2h_s = g->eng->host_rec_copy_s = 0;
}

// test hook (not in the public header): suffix array + longest-repeated-prefix of a window of genome 0
int pb200_debug_index(pb200_genomes* g, int64_t ref_start, int32_t n, int32_t minsize, uint32_t* sa, int32_t* lrp) {
    return guarded([&]() { g->eng->debug_index(ref_start, n, minsize, sa, lrp); return (int)PB200_OK; });
}

// test hook: bit 0 = the last debug_index window went through the general prefix-doubling path, bit 1 = 2-bit keys
int pb200_debug_index_flags(pb200_genomes* g) { return g->eng->debug_index_flags(); }

int pb200_comm_set(pb200_genomes* g, int rank, int world, pb200_allgather_cb ag, pb200_allreduce_cb ar, pb200_bcast_cb bc, void* user,
                   int bcast_index) {
    if (!g || world < 1 || rank < 0 || rank >= world || !ag || !ar || !bc) { pb200::g_last_error = "bad arguments"; return PB200_ERR_ARG; }
    g->comm.reset(new CallbackComm);
    g->comm->rank = rank; g->comm->world = world; g->comm->ag = ag; g->comm->ar = ar; g->comm->bc = bc; g->comm->user = user;
    g->comm->eng = g->eng.get();
    g->bcast_index = bcast_index != 0;
    return PB200_OK;
}
void pb200_comm_clear(pb200_genomes* g) { if (g) g->comm.reset(); }

}  // extern "C"
