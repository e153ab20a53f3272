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
struct Ev { uint16_t l, e, u, pad; int32_t d1; };      // ref start, ref end, uniqueness floor u = l + lrp[l], diagonal (query start - ref start)

struct ClassCfg {
    int n_cap, m_cap, ev_cap, cand_cap, threads;
    __host__ __device__ static size_t al(size_t x) { return (x + 15) & ~(size_t)15; }
    // ev_cap = capacity of the event store for ALL strands of ALL queries of the window
    __host__ __device__ size_t smem_bytes(int nq) const {
        return al(n_cap) + 2 * al(m_cap) + 3 * al(2 * (size_t)n_cap) + al((size_t)ev_cap * sizeof(Ev)) + al(2 * (size_t)(2 * nq + 2)) +
               2 * al(2 * (size_t)cand_cap) + 64;
    }
};

constexpr int SM_MAX_THREADS = 256;
constexpr uint16_t LRP_UNKNOWN = 0xFFFFu;

// longest prefix of R[l..) that occurs at another position of R (A1), computed on demand and cached
__device__ __forceinline__ int lazy_lrp(const uint8_t* __restrict__ R, int n, uint16_t* __restrict__ lrp, int l) {
    uint16_t v = lrp[l];
    if (v != LRP_UNKNOWN) return v;
    int best = 0;
    const uint8_t c0 = R[l];
    for (int l2 = 0; l2 < n; ++l2) {
        if (R[l2] != c0 || l2 == l) continue;
        int t = 1;
        const int lim = n - max(l, l2);
        while (t < lim && R[l + t] == R[l2 + t]) ++t;
        best = max(best, t);
    }
    lrp[l] = (uint16_t)best;         // racing writers store the same value
    return best;
}

// MEM events of one strand: Q (m bases in smem) against R (n bases in smem).  Seeds = 2-base matches at every
// (minsize-1)-th query position against every reference position (a dense rows x n grid, no divergence in the
// enumeration); a match of >= minsize bases contains exactly one seed whose left extension is shorter than the
// seed spacing, which reports it.
__device__ inline void find_events(const uint8_t* __restrict__ R, int n, const uint8_t* __restrict__ Q, int m,
                                   uint16_t* __restrict__ lrp, int minsize, Ev* __restrict__ ev, int* __restrict__ ev_n, int ev_cap,
                                   int nthreads) {
    const int step = max(1, minsize - 1);
    for (int j = 0; j + 1 < m; j += step) {
        const uint8_t q0 = Q[j], q1 = Q[j + 1];
        for (int l = threadIdx.x; l + 1 < n; l += nthreads) {
            if (R[l] != q0 || R[l + 1] != q1) continue;
            int c = 0;
            const int cmax = min(step, min(j, l));
            while (c < cmax && Q[j - 1 - c] == R[l - 1 - c]) ++c;
            if (c >= step) continue;                  // the previous seed row lies in the same match
            int e = 2;
            const int emax = min(m - j, n - l);
            while (e < emax && Q[j + e] == R[l + e]) ++e;
            const int L = c + e, l0 = l - c;
            if (L < minsize) continue;
            const int lr = lazy_lrp(R, n, lrp, l0);
            if (L > lr) {
                int slot = atomicAdd(ev_n, 1);
                if (slot < ev_cap) {
                    ev[slot].l = (uint16_t)l0; ev[slot].e = (uint16_t)(l0 + L); ev[slot].u = (uint16_t)(l0 + lr); ev[slot].pad = 0;
                    ev[slot].d1 = (j - c) - l0;
                }
            }
        }
    }
    // a window shorter than 2 bases cannot seed; minsize >= 2 always holds for the callers (q = 30 -> minsize >= 6)
}
// (UP', EP', d1) of a strand at reference position k  (Intersect_UM closed form over the strand's events)
__device__ __forceinline__ void eval_at(const Ev* __restrict__ ev, int ne, int k, int& UP, int& EP, int& d1) {
    int fl = 0, t1 = 0, t2 = 0, dd = 0;
    for (int i = 0; i < ne; ++i) {
        int l = ev[i].l;
        if (l > k) continue;
        int e = ev[i].e;
        fl = max(fl, (int)ev[i].u);
        if (e > t1) { t2 = t1; t1 = e; dd = ev[i].d1; }
        else if (e == t1) { t2 = t1; }
        else if (e > t2) t2 = e;
    }
    UP = max(fl, t2);
    EP = max(t1, fl);
    d1 = dd;
}

__global__ void __launch_bounds__(SM_MAX_THREADS) small_region_kernel(
    const uint8_t* __restrict__ text, const int64_t* __restrict__ gbase_fwd, const int64_t* __restrict__ gbase_rc,
    const int64_t* __restrict__ glen, int nq, const TaskDev* __restrict__ tasks, const int32_t* __restrict__ qcoords,
    const int32_t* __restrict__ task_ids, int ntasks, ClassCfg cfg, TaskOut* __restrict__ outs,
    unsigned long long* __restrict__ cand_counter, unsigned long long cand_cap_global, int32_t* __restrict__ out_k,
    int32_t* __restrict__ out_lon, int32_t* __restrict__ out_sp, uint8_t* __restrict__ out_fwd) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int ti = blockIdx.x;
    if (ti >= ntasks) return;
    const int T = blockDim.x;
    const int task_id = task_ids[ti];
    const TaskDev tk = tasks[task_id];
    const int n = tk.n, minsize = tk.minsize;
    size_t off = 0;
    uint8_t* R = smem + off; off += ClassCfg::al(cfg.n_cap);
    uint8_t* Qf = smem + off; off += ClassCfg::al(cfg.m_cap);
    uint8_t* Qc = smem + off; off += ClassCfg::al(cfg.m_cap);
    uint16_t* lrp = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.n_cap);
    uint16_t* MUP = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.n_cap);
    uint16_t* MEP = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.n_cap);
    Ev* evs = reinterpret_cast<Ev*>(smem + off); off += ClassCfg::al((size_t)cfg.ev_cap * sizeof(Ev));
    uint16_t* evoff = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)(2 * nq + 2));   // [2*nq+1] strand starts
    uint16_t* candK = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.cand_cap);
    uint16_t* candM = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.cand_cap);
    int* s_int = reinterpret_cast<int*>(smem + off);     // [0]=event count [2]=ncand [3]=overflow ; [4..5] = cand base (int64)
    const int tid = threadIdx.x;
    const int32_t* qs = qcoords + tk.qcoord_off;
    const int32_t* ql = qs + nq;

    for (int i = tid; i < n; i += T) { R[i] = text[tk.ref_off + i]; MUP[i] = 0; MEP[i] = (uint16_t)n; lrp[i] = LRP_UNKNOWN; }
    if (tid < 8) s_int[tid] = 0;
    __syncthreads();

    // pass 0: fold all queries (ini order) into Master, keeping every strand's events in shared memory
    for (int q = 0; q < nq; ++q) {
        const int m = ql[q];
        const int64_t g = q + 1;
        const int64_t f_off = gbase_fwd[g] + qs[q];
        const int64_t c_off = gbase_rc[g] + (glen[g] - qs[q] - m);
        for (int i = tid; i < m; i += T) { Qf[i] = text[f_off + i]; Qc[i] = text[c_off + i]; }
        __syncthreads();
        const int e0 = min(s_int[0], cfg.ev_cap);
        __syncthreads();
        find_events(R, n, Qf, m, lrp, minsize, evs, &s_int[0], cfg.ev_cap, T);
        __syncthreads();
        const int e1 = min(s_int[0], cfg.ev_cap);
        __syncthreads();
        find_events(R, n, Qc, m, lrp, minsize, evs, &s_int[0], cfg.ev_cap, T);
        __syncthreads();
        const int e2 = s_int[0];
        if (e2 > cfg.ev_cap) { if (tid == 0) s_int[3] = 1; __syncthreads(); break; }
        if (tid == 0) { evoff[2 * q] = (uint16_t)e0; evoff[2 * q + 1] = (uint16_t)e1; evoff[2 * q + 2] = (uint16_t)e2; }
        for (int k = tid; k < n; k += T) {
            int UPf, EPf, df, UPc, EPc, dc;
            eval_at(evs + e0, e1 - e0, k, UPf, EPf, df);
            eval_at(evs + e1, e2 - e1, k, UPc, EPc, dc);
            int mep = MEP[k], mup = MUP[k];
            int fe = min(mep, EPf), ce = min(mep, EPc);
            if (fe > ce) { mup = max(mup, UPf); mep = fe; }
            else { mup = max(mup, UPc); mep = ce; }
            MUP[k] = (uint16_t)mup; MEP[k] = (uint16_t)mep;
        }
        __syncthreads();
    }
    __syncthreads();
    // A5: ordered emission by warp 0
    if (!s_int[3] && tid < 32) {
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
                if (slot < cfg.cand_cap) { candK[slot] = (uint16_t)k; candM[slot] = (uint16_t)n; }
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
    const int ovf = s_int[3];
    const int nc = ovf ? 0 : s_int[2];
    const int64_t base = *reinterpret_cast<int64_t*>(&s_int[4]);
    // pass 1: replay the fold at the candidate positions from the stored events (strand flag + start per query)
    for (int c = tid; c < nc; c += T) {
        const int k = candK[c];
        int M = n;
        for (int q = 0; q < nq; ++q) {
            const int e0 = evoff[2 * q], e1 = evoff[2 * q + 1], e2 = evoff[2 * q + 2];
            int UPf, EPf, df, UPc, EPc, dc;
            eval_at(evs + e0, e1 - e0, k, UPf, EPf, df);
            eval_at(evs + e1, e2 - e1, k, UPc, EPc, dc);
            int fe = min(M, EPf), ce = min(M, EPc);
            const size_t o = (size_t)(base + c) * nq + q;
            if (fe > ce) { out_sp[o] = k + df; out_fwd[o] = 1; M = fe; }
            else { out_sp[o] = k + dc; out_fwd[o] = 0; M = ce; }
        }
        out_k[base + c] = k;
        out_lon[base + c] = (int)MEP[k] - k;
    }
    if (tid == 0) { outs[task_id].ncand = ovf ? -ovf : nc; outs[task_id].cand_base = base; outs[task_id].pad = 0; }
}

}  // namespace small
}  // namespace pb200
