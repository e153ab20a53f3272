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
        return al(n_cap) + 2 * al(m_cap) + 2 * al(2 * (size_t)m_cap) + 4 * al(2 * (size_t)n_cap) + al((size_t)ev_cap * sizeof(Ev)) + al(2 * (size_t)(2 * nq + 2)) +
               2 * al(2 * (size_t)cand_cap) + 64;
    }
};

constexpr int SM_MAX_THREADS = 256;
constexpr uint16_t LRP_UNKNOWN = 0xFFFFu;

// ---- shared-memory string compares, 8 bytes per step (arrays are 16-byte aligned and padded, over-reads are clamped by `limit`)
__device__ __forceinline__ uint64_t sld8u(const uint8_t* p) {
    const uint64_t* a = reinterpret_cast<const uint64_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)7);
    const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(p) & 7) * 8;
    uint64_t lo = a[0];
    if (sh == 0) return lo;
    return (lo >> sh) | (a[1] << (64 - sh));
}
// equal leading bytes of a[0..limit) and b[0..limit)
__device__ __forceinline__ int smatch_fwd(const uint8_t* a, const uint8_t* b, int limit) {
    int t = 0;
    while (t < limit) {
        uint64_t x = sld8u(a + t) ^ sld8u(b + t);
        if (x) { t += (__ffsll((long long)x) - 1) >> 3; break; }
        t += 8;
    }
    return t < limit ? t : limit;
}
// equal bytes going left: a[-1]==b[-1], a[-2]==b[-2], ... at most `limit` (the caller guarantees a-limit, b-limit are inside the arrays)
__device__ __forceinline__ int smatch_bwd(const uint8_t* a, const uint8_t* b, int limit) {
    int c = 0;
    while (c + 8 <= limit) {
        uint64_t x = sld8u(a - c - 8) ^ sld8u(b - c - 8);
        if (x) return c + (__clzll((long long)x) >> 3);
        c += 8;
    }
    while (c < limit && a[-1 - c] == b[-1 - c]) ++c;
    return c;
}

// MEM events of one query (both strands) against R.  Seeds = SEED_K-base matches (packed 3 bits per base in R4[l]) at every
// (minsize-SEED_K+1)-th query position against every reference position: a dense rows x n grid without divergence in
// the enumeration, ~1/256 random hits.  A match of >= minsize bases contains exactly one seed whose left extension is
// shorter than the seed spacing, and that seed reports it.  Events are stored unconditionally (strand in pad bit 1);
// uniqueness is decided afterwards by resolve_unique().
constexpr int SEED_K = 4;
__device__ __forceinline__ uint32_t pack4(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 3) | ((uint32_t)p[2] << 6) | ((uint32_t)p[3] << 9);
}
__device__ __forceinline__ void seed_hit(const uint8_t* __restrict__ R, int n, const uint8_t* __restrict__ Q, int m, int j, int l, int step,
                                         int minsize, int strand, Ev* __restrict__ ev, int* __restrict__ ev_n, int ev_cap) {
    const int cmax = min(step, min(j, l));
    int c = 0;
    while (c < cmax && Q[j - 1 - c] == R[l - 1 - c]) ++c;
    if (c >= step) return;                        // the previous seed row lies in the same match
    const int emax = min(m - j, n - l);
    const int e = SEED_K + smatch_fwd(Q + j + SEED_K, R + l + SEED_K, emax - SEED_K);
    const int L = c + e, l0 = l - c;
    if (L < minsize) return;
    int slot = atomicAdd(ev_n, 1);
    if (slot < ev_cap) {
        ev[slot].l = (uint16_t)l0; ev[slot].e = (uint16_t)(l0 + L); ev[slot].u = 0; ev[slot].pad = (uint16_t)(strand << 1);
        ev[slot].d1 = (j - c) - l0;
    }
}
// Q4f/Q4c: packed seed codes of the sampled query rows (filled by the caller, one row per thread)
__device__ inline void find_events(const uint8_t* __restrict__ R, const uint16_t* __restrict__ R4, int n,
                                   const uint8_t* __restrict__ Qf, const uint8_t* __restrict__ Qc, int m,
                                   const uint16_t* __restrict__ Q4f, const uint16_t* __restrict__ Q4c, int nrows, int step,
                                   int minsize, Ev* __restrict__ ev, int* __restrict__ ev_n, int ev_cap, int nthreads) {
    for (int l = threadIdx.x; l + SEED_K <= n; l += nthreads) {
        const uint32_t r4 = R4[l];
        for (int row = 0; row < nrows; ++row) {
            const uint32_t qf = Q4f[row], qc = Q4c[row];       // same address for the whole warp: broadcast
            if (r4 == qf) seed_hit(R, n, Qf, m, row * step, l, step, minsize, 0, ev, ev_n, ev_cap);
            if (r4 == qc) seed_hit(R, n, Qc, m, row * step, l, step, minsize, 1, ev, ev_n, ev_cap);
        }
    }
}
// A1 on demand: for every new event, lrp[l] = longest prefix of R[l..) occurring at another position of R (one warp per
// event, lanes stride over the other positions; cached per reference position), then u = l + lrp and validity L > lrp.
__device__ inline void resolve_unique(const uint8_t* __restrict__ R, int n, uint16_t* __restrict__ lrp, Ev* __restrict__ ev, int e_begin,
                                      int e_end, int nthreads) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = nthreads >> 5;
    for (int i = e_begin + warp; i < e_end; i += nwarps) {
        const int l = ev[i].l;
        int v = lrp[l];
        if (v == LRP_UNKNOWN) {
            int best = 0;
            const uint8_t c0 = R[l];
            for (int l2 = lane; l2 < n; l2 += 32) {
                if (l2 == l || R[l2] != c0) continue;
                best = max(best, 1 + smatch_fwd(R + l + 1, R + l2 + 1, n - max(l, l2) - 1));
            }
            for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
            v = best;
            if (lane == 0) lrp[l] = (uint16_t)v;      // racing warps store the same value
        }
        if (lane == 0) {
            const int L = (int)ev[i].e - l;
            ev[i].u = (uint16_t)(l + v);
            ev[i].pad = (uint16_t)((ev[i].pad & 2) | ((L > v) ? 1 : 0));   // bit0: unique in R (an event of Find_UM); bit1: reverse strand
        }
    }
}
// (UP', EP', d1) of both strands of one query at reference position k (Intersect_UM closed form over the query's events)
struct StrandVal { int UP, EP, d1; };
struct Acc { int fl, t1, t2, dd; };
__device__ __forceinline__ void acc_add(Acc& a, int u, int e, int d1) {
    a.fl = max(a.fl, u);
    if (e > a.t1) { a.t2 = a.t1; a.t1 = e; a.dd = d1; }
    else if (e == a.t1) { a.t2 = a.t1; }
    else if (e > a.t2) a.t2 = e;
}
__device__ __forceinline__ void eval_at(const Ev* __restrict__ ev, int ne, int k, StrandVal& F, StrandVal& C) {
    Acc af = {0, 0, 0, 0}, ac = {0, 0, 0, 0};
    for (int i = 0; i < ne; ++i) {
        const int l = ev[i].l;
        const int tag = ev[i].pad;
        if (l > k || !(tag & 1)) continue;
        if (tag & 2) acc_add(ac, (int)ev[i].u, (int)ev[i].e, ev[i].d1);
        else acc_add(af, (int)ev[i].u, (int)ev[i].e, ev[i].d1);
    }
    F.UP = max(af.fl, af.t2); F.EP = max(af.t1, af.fl); F.d1 = af.dd;
    C.UP = max(ac.fl, ac.t2); C.EP = max(ac.t1, ac.fl); C.d1 = ac.dd;
}

__global__ void __launch_bounds__(SM_MAX_THREADS, 6) small_region_kernel(
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
    uint16_t* R4 = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.n_cap);
    uint16_t* Q4f = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.m_cap);
    uint16_t* Q4c = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.m_cap);
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
    for (int i = tid; i + SEED_K <= n; i += T) R4[i] = (uint16_t)pack4(R + i);
    __syncthreads();

    // pass 0: fold all queries (ini order) into Master, keeping every query's events in shared memory
    int e0 = 0;
    const int step = max(1, minsize - SEED_K + 1);
    for (int q = 0; q < nq; ++q) {
        const int m = ql[q];
        const int64_t g = q + 1;
        const int64_t f_off = gbase_fwd[g] + qs[q];
        const int64_t c_off = gbase_rc[g] + (glen[g] - qs[q] - m);
        for (int i = tid; i < m; i += T) { Qf[i] = text[f_off + i]; Qc[i] = text[c_off + i]; }
        __syncthreads();
        const int nrows = m >= SEED_K ? (m - SEED_K) / step + 1 : 0;
        for (int row = tid; row < nrows; row += T) { Q4f[row] = (uint16_t)pack4(Qf + row * step); Q4c[row] = (uint16_t)pack4(Qc + row * step); }
        __syncthreads();
        find_events(R, R4, n, Qf, Qc, m, Q4f, Q4c, nrows, step, minsize, evs, &s_int[0], cfg.ev_cap, T);
        __syncthreads();
        const int e2 = s_int[0];
        if (e2 > cfg.ev_cap) { if (tid == 0) s_int[3] = 1; __syncthreads(); break; }
        resolve_unique(R, n, lrp, evs, e0, e2, T);
        if (tid == 0) { evoff[q] = (uint16_t)e0; evoff[q + 1] = (uint16_t)e2; }
        __syncthreads();
        for (int k = tid; k < n; k += T) {
            StrandVal F, C;
            eval_at(evs + e0, e2 - e0, k, F, C);
            int mep = MEP[k], mup = MUP[k];
            int fe = min(mep, F.EP), ce = min(mep, C.EP);
            if (fe > ce) { mup = max(mup, F.UP); mep = fe; }
            else { mup = max(mup, C.UP); mep = ce; }
            MUP[k] = (uint16_t)mup; MEP[k] = (uint16_t)mep;
        }
        e0 = e2;
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
            const int a0 = evoff[q], a1 = evoff[q + 1];
            StrandVal F, C;
            eval_at(evs + a0, a1 - a0, k, F, C);
            int fe = min(M, F.EP), ce = min(M, C.EP);
            const size_t o = (size_t)(base + c) * nq + q;
            if (fe > ce) { out_sp[o] = k + F.d1; out_fwd[o] = 1; M = fe; }
            else { out_sp[o] = k + C.d1; out_fwd[o] = 0; M = ce; }
        }
        out_k[base + c] = k;
        out_lon[base + c] = (int)MEP[k] - k;
    }
    if (tid == 0) { outs[task_id].ncand = ovf ? -ovf : nc; outs[task_id].cand_base = base; outs[task_id].pad = 0; }
}

}  // namespace small
}  // namespace pb200
