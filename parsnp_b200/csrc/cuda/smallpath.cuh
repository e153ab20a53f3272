// Batched small-window search: one CTA runs the complete window search of Aligner::setMums1 (src/parsnp.cpp:1570-1695:
// index + Find_UM/Intersect_UM/Merge_Master for every query and strand + candidate emission) for ONE region whose
// reference window and query regions fit in shared memory.  The recursion of Aligner::doWork (src/parsnp.cpp:173-317)
// issues 10^5..10^6 such windows of ~10^2 bp; the reference builds and frees a suffix graph for each.
//
// In shared memory: the reference window R, lrp[l] (longest repeated prefix, brute force), the running Master (UP, EP),
// one query strand at a time, the strand's MEM events.  Pass 1 folds all queries into Master and emits candidate
// positions; pass 2 (only if there are candidates) replays the fold at the candidate positions to recover every
// query's strand flag and start position.  MEMs are found by sampling every minsize-th cell of each diagonal and
// extending (a run of >= minsize matches must contain a sampled cell).
#pragma once
#include "util.cuh"

namespace pb200 {
namespace small {

struct TaskDev {
    int64_t ref_off;      // offset of the window start inside the device text (forward text of genome 0)
    int32_t n;            // window length
    int32_t minsize;
    int64_t qcoord_off;   // offset into qcoords: start[nq], len[nq] (int32 each)
};
struct TaskOut {
    int32_t ncand;        // -1 = overflow (events or candidates): the task must be re-run with larger capacity
    int32_t pad;
    int64_t cand_base;    // first candidate slot in the global candidate arrays
};
struct Ev { uint16_t l, e; int32_t d1; };

struct ClassCfg {
    int n_cap, m_cap, ev_cap, cand_cap;
    size_t smem_bytes() const {
        return (size_t)align4(n_cap) + align4(m_cap) + 3 * (size_t)align4(2 * n_cap) + 2 * (size_t)ev_cap * sizeof(Ev) + 2 * (size_t)align4(2 * cand_cap) + 64;
    }
    __host__ __device__ static size_t align4(size_t x) { return (x + 15) & ~(size_t)15; }
};

constexpr int SM_THREADS = 128;

// MEM events of one strand: Q (m bases in smem) against R (n bases in smem)
__device__ inline void find_events(const uint8_t* __restrict__ R, int n, const uint8_t* __restrict__ Q, int m,
                                   const uint16_t* __restrict__ lrp, int minsize, Ev* __restrict__ ev, int* __restrict__ ev_n, int ev_cap) {
    const int s = minsize;
    const int ndiag = n + m - 1;
    for (int dd = threadIdx.x; dd < ndiag; dd += SM_THREADS) {
        const int d = dd - (m - 1);                  // d = l - j
        const int j_start = d < 0 ? -d : 0, l_start = d < 0 ? 0 : d;
        const int len = min(m - j_start, n - l_start);
        for (int t = s - 1; t < len; t += s) {
            int j = j_start + t, l = l_start + t;
            if (Q[j] != R[l]) continue;
            int c = 0;
            while (c < s && t - 1 - c >= 0 && Q[j - 1 - c] == R[l - 1 - c]) ++c;
            if (c >= s) continue;                    // the previous sampled cell lies in the same run
            int e = 1;
            while (t + e < len && Q[j + e] == R[l + e]) ++e;
            const int L = c + e, l0 = l - c;
            if (L >= minsize && L > (int)lrp[l0]) {
                int slot = atomicAdd(ev_n, 1);
                if (slot < ev_cap) { ev[slot].l = (uint16_t)l0; ev[slot].e = (uint16_t)(l0 + L); ev[slot].d1 = (j - c) - l0; }
            }
        }
    }
}
// (UP', EP', d1) of a strand at reference position k  (Intersect_UM closed form over the strand's events)
__device__ __forceinline__ void eval_at(const Ev* __restrict__ ev, int ne, const uint16_t* __restrict__ lrp, int k, int& UP, int& EP, int& d1) {
    int fl = 0, t1 = 0, t2 = 0, dd = 0;
    for (int i = 0; i < ne; ++i) {
        int l = ev[i].l;
        if (l > k) continue;
        int e = ev[i].e;
        fl = max(fl, l + (int)lrp[l]);
        if (e > t1) { t2 = t1; t1 = e; dd = ev[i].d1; }
        else if (e == t1) { t2 = t1; }
        else if (e > t2) t2 = e;
    }
    UP = max(fl, t2);
    EP = max(t1, fl);
    d1 = dd;
}

__global__ void __launch_bounds__(SM_THREADS) small_region_kernel(
    const uint8_t* __restrict__ text, const int64_t* __restrict__ gbase_fwd, const int64_t* __restrict__ gbase_rc,
    const int64_t* __restrict__ glen, int nq, const TaskDev* __restrict__ tasks, const int32_t* __restrict__ qcoords,
    const int32_t* __restrict__ task_ids, int ntasks, ClassCfg cfg, TaskOut* __restrict__ outs,
    unsigned long long* __restrict__ cand_counter, unsigned long long cand_cap_global, int32_t* __restrict__ out_k,
    int32_t* __restrict__ out_lon, int32_t* __restrict__ out_sp, uint8_t* __restrict__ out_fwd) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int ti = blockIdx.x;
    if (ti >= ntasks) return;
    const int task_id = task_ids[ti];
    const TaskDev tk = tasks[task_id];
    const int n = tk.n, minsize = tk.minsize;
    size_t off = 0;
    uint8_t* R = smem + off; off += ClassCfg::align4(cfg.n_cap);
    uint8_t* Q = smem + off; off += ClassCfg::align4(cfg.m_cap);
    uint16_t* lrp = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::align4(2 * cfg.n_cap);
    uint16_t* MUP = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::align4(2 * cfg.n_cap);
    uint16_t* MEP = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::align4(2 * cfg.n_cap);
    Ev* evF = reinterpret_cast<Ev*>(smem + off); off += (size_t)cfg.ev_cap * sizeof(Ev);
    Ev* evC = reinterpret_cast<Ev*>(smem + off); off += (size_t)cfg.ev_cap * sizeof(Ev);
    uint16_t* candK = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::align4(2 * cfg.cand_cap);
    uint16_t* candM = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::align4(2 * cfg.cand_cap);
    int* s_int = reinterpret_cast<int*>(smem + off);     // [0]=nF [1]=nC [2]=ncand [3]=overflow ; [4..5] = cand base (int64)
    const int tid = threadIdx.x;
    const int32_t* qs = qcoords + tk.qcoord_off;
    const int32_t* ql = qs + nq;

    for (int i = tid; i < n; i += SM_THREADS) { R[i] = text[tk.ref_off + i]; MUP[i] = 0; MEP[i] = (uint16_t)n; }
    if (tid < 8) s_int[tid] = 0;
    __syncthreads();
    // A1: lrp[l] = longest prefix of R[l..) occurring at another position
    for (int l = tid; l < n; l += SM_THREADS) {
        int best = 0;
        const uint8_t c0 = R[l];
        for (int l2 = 0; l2 < n; ++l2) {
            if (R[l2] != c0 || l2 == l) continue;
            int t = 1;
            const int lim = n - max(l, l2);
            while (t < lim && R[l + t] == R[l2 + t]) ++t;
            best = max(best, t);
        }
        lrp[l] = (uint16_t)best;
    }
    __syncthreads();

    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            if (s_int[3] || s_int[2] == 0) break;
            for (int c = tid; c < s_int[2]; c += SM_THREADS) candM[c] = (uint16_t)n;
            __syncthreads();
        }
        for (int q = 0; q < nq; ++q) {
            const int m = ql[q];
            const int64_t g = q + 1;
            const int64_t f_off = gbase_fwd[g] + qs[q];
            const int64_t c_off = gbase_rc[g] + (glen[g] - qs[q] - m);
            // forward strand
            for (int i = tid; i < m; i += SM_THREADS) Q[i] = text[f_off + i];
            if (tid == 0) { s_int[0] = 0; s_int[1] = 0; }
            __syncthreads();
            find_events(R, n, Q, m, lrp, minsize, evF, &s_int[0], cfg.ev_cap);
            __syncthreads();
            // reverse strand
            for (int i = tid; i < m; i += SM_THREADS) Q[i] = text[c_off + i];
            __syncthreads();
            find_events(R, n, Q, m, lrp, minsize, evC, &s_int[1], cfg.ev_cap);
            __syncthreads();
            const int nF = s_int[0], nC = s_int[1];
            if (nF > cfg.ev_cap || nC > cfg.ev_cap) { if (tid == 0) s_int[3] = 1; __syncthreads(); break; }
            if (pass == 0) {
                for (int k = tid; k < n; k += SM_THREADS) {
                    int UPf, EPf, df, UPc, EPc, dc;
                    eval_at(evF, nF, lrp, k, UPf, EPf, df);
                    eval_at(evC, nC, lrp, k, UPc, EPc, dc);
                    int mep = MEP[k], mup = MUP[k];
                    int fe = min(mep, EPf), ce = min(mep, EPc);
                    if (fe > ce) { mup = max(mup, UPf); mep = fe; }
                    else { mup = max(mup, UPc); mep = ce; }
                    MUP[k] = (uint16_t)mup; MEP[k] = (uint16_t)mep;
                }
            } else {
                const int nc = s_int[2];
                const int64_t base = *reinterpret_cast<int64_t*>(&s_int[4]);
                for (int c = tid; c < nc; c += SM_THREADS) {
                    const int k = candK[c];
                    int UPf, EPf, df, UPc, EPc, dc;
                    eval_at(evF, nF, lrp, k, UPf, EPf, df);
                    eval_at(evC, nC, lrp, k, UPc, EPc, dc);
                    int M = candM[c];
                    int fe = min(M, EPf), ce = min(M, EPc);
                    const size_t o = (size_t)(base + c) * nq + q;
                    if (fe > ce) { out_sp[o] = k + df; out_fwd[o] = 1; candM[c] = (uint16_t)fe; }
                    else { out_sp[o] = k + dc; out_fwd[o] = 0; candM[c] = (uint16_t)ce; }
                }
            }
            __syncthreads();
        }
        if (pass == 0) {
            __syncthreads();
            if (s_int[3]) break;
            // A5: ordered emission by warp 0
            if (tid < 32) {
                int count = 0;
                for (int b = 0; b < n; b += 32) {
                    int k = b + tid;
                    bool f = false;
                    if (k < n) {
                        int prev = k ? (int)MEP[k - 1] : 0;
                        int ep = MEP[k];
                        f = ep > prev && (int)MUP[k] < ep && ep - k >= minsize;
                    }
                    unsigned bal = __ballot_sync(0xffffffffu, f);
                    if (f) {
                        int slot = count + __popc(bal & ((1u << tid) - 1));
                        if (slot < cfg.cand_cap) candK[slot] = (uint16_t)k;
                    }
                    count += __popc(bal);
                }
                if (tid == 0) {
                    if (count > cfg.cand_cap) s_int[3] = 1;
                    else {
                        s_int[2] = count;
                        if (count > 0) {
                            unsigned long long base = atomicAdd(cand_counter, (unsigned long long)count);
                            if (base + count > cand_cap_global) s_int[3] = 2;
                            *reinterpret_cast<int64_t*>(&s_int[4]) = (int64_t)base;
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    const int ovf = s_int[3];
    const int nc = ovf ? 0 : s_int[2];
    const int64_t base = *reinterpret_cast<int64_t*>(&s_int[4]);
    for (int c = tid; c < nc; c += SM_THREADS) {
        int k = candK[c];
        out_k[base + c] = k;
        out_lon[base + c] = (int)MEP[k] - k;
    }
    if (tid == 0) { outs[task_id].ncand = ovf ? -ovf : nc; outs[task_id].cand_base = base; outs[task_id].pad = 0; }
}

}  // namespace small
}  // namespace pb200
