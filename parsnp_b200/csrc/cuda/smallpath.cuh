// Batched small-window search: one CTA runs the complete window search of Aligner::setMums1 (src/parsnp.cpp:1570-1695:
// index + Find_UM/Intersect_UM/Merge_Master for every query and strand + candidate emission) for ONE region whose
// reference window and query regions fit in shared memory.  The recursion of Aligner::doWork (src/parsnp.cpp:173-317)
// issues 10^5..10^6 such windows of ~10^2 bp; the reference builds and frees a suffix graph for each.
//
// The queries of a window are processed in GROUPS (as many consecutive queries as fit the shared-memory text buffer), each
// group in a few block-wide stages so that every stage is a flat, balanced loop over all queries of the group:
//   1. both strands of the group's query regions arrive by TMA bulk copies (cp.async.bulk of the 16-byte-aligned superset of
//      each region, one issuing thread per query, completion on an mbarrier that doubles as the block barrier of the stage)
//   2. pack the sampled seed rows: K bases at 3 bits each, every (minsize-K+1)-th query position, both strands.  K is chosen
//      per window so that consecutive seeds of a diagonal abut or overlap (K >= (minsize+1)/2, 4 <= K <= 10)
//   3. every (row, strand) looks its seed up in a hash table of the window's K-mers (open addressing in shared memory, built
//      once per window).  A match of >= minsize bases contains exactly one seed whose left extension is shorter than the seed
//      spacing; with abutting seeds that is the seed whose predecessor on the diagonal is NOT a hit, so only those are queued
//   4. one thread per queued hit: left extension, right extension -> MEM events (staging buffer)
//   5. one warp per event: uniqueness floor lrp[l] on demand (cached per reference position)
//   6. counting sort of the staged events by query into the window's event store
//   7. one thread per reference position: fold the group's queries (ini order) into the running Master (UP, EP)
// After the last group: ordered emission of candidate positions and, for those only, a replay of the fold from the stored
// events to recover every query's strand flag and start position.
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
    int32_t ncand;        // < 0 = overflow (-1: a per-CTA capacity, -2: the global candidate buffer): the task must be re-run
    int32_t pad;
    int64_t cand_base;    // first candidate slot in the global candidate arrays
};
// MEM event: le = ref start | ref end << 16 ; ut = uniqueness floor u = l + lrp[l] | tag << 16 ; d1 = query start - ref start
// tag: bit0 = unique in R (an event of Find_UM), bit1 = reverse strand, bits 2.. = query index inside its group (staging only)
struct Ev { uint32_t le, ut; int32_t d1; };

constexpr int GROUP_MAX = 64;           // queries per group

struct ClassCfg {
    int n_cap, m_cap, ev_cap, cand_cap, threads;
    int qbuf;             // bytes of query text per group (both strands, each with room for its misalignment); >= 2 * al(m_cap + 15)
    int rows_cap;         // seed rows per group (incl. one pad row per query); >= m_cap
    int hq_cap;           // seed-hit queue entries per group
    int stg_cap;          // staged events per group
    __host__ __device__ static size_t al(size_t x) { return (x + 15) & ~(size_t)15; }
    // ev_cap = capacity of the event store for ALL strands of ALL queries of the window
    __host__ __device__ size_t smem_bytes(int nq) const {
        return al(n_cap + 16) + 3 * al(2 * (size_t)n_cap) + al(4 * (size_t)n_cap) + al(8 * (size_t)n_cap) + al((size_t)qbuf) + al(8 * (size_t)rows_cap) +
               al((size_t)rows_cap) + al(4 * (size_t)hq_cap) +
               al((size_t)stg_cap * sizeof(Ev)) + al((size_t)ev_cap * sizeof(Ev)) + al(2 * (size_t)(nq + 2)) + 2 * al(2 * (size_t)cand_cap) +
               al(2 * 3 * (size_t)(GROUP_MAX + 1)) + al(4 * 2 * (size_t)GROUP_MAX) + 64;
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

constexpr int SEED_K_MIN = 4, SEED_K_MAX = 10;
constexpr uint32_t SEED_PAD = 0xffffffffu;      // never a seed code (codes use 30 bits), never a table entry
__device__ __forceinline__ uint32_t pack_seed(const uint8_t* p, int K) {
    uint32_t c = 0;
    for (int t = 0; t < K; ++t) c |= (uint32_t)p[t] << (3 * t);
    return c;
}
__device__ __forceinline__ uint32_t seed_hash(uint32_t code, int hbits) { return (code * 0x9E3779B1u) >> (32 - hbits); }

// A1 on demand: for every staged event, lrp[l] = longest prefix of R[l..) occurring at another position of R (one warp per
// event, lanes stride over the other positions; cached per reference position), then u = l + lrp and validity L > lrp.
// Repeats shorter than the K-base seed are reported as 0: every event has L >= minsize >= K, and a floor u <= l + K - 1 can
// neither invalidate the event nor reach an emitted candidate's end (emission needs EP - k >= minsize with k >= l).
__device__ inline void resolve_unique(const uint8_t* __restrict__ R, const uint32_t* __restrict__ R4, int n, int SEED_K, uint16_t* __restrict__ lrp,
                                      Ev* __restrict__ ev, int count, int nthreads) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = nthreads >> 5;
    for (int i = warp; i < count; i += nwarps) {
        const uint32_t le = ev[i].le;
        const int l = (int)(le & 0xffffu);
        int v = lrp[l];
        if (v == LRP_UNKNOWN) {
            int best = 0;
            const uint32_t r4 = R4[l];
            for (int l2 = lane; l2 + SEED_K <= n; l2 += 32) {
                if (l2 == l || R4[l2] != r4) continue;
                best = max(best, SEED_K + smatch_fwd(R + l + SEED_K, R + l2 + SEED_K, n - max(l, l2) - SEED_K));
            }
            for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
            v = best;
            if (lane == 0) lrp[l] = (uint16_t)v;      // racing warps store the same value
        }
        if (lane == 0) {
            const int L = (int)(le >> 16) - l;
            const uint32_t tag = (ev[i].ut >> 16) & ~1u;
            ev[i].ut = (uint32_t)(l + v) | ((tag | (L > v ? 1u : 0u)) << 16);
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
        const uint32_t le = ev[i].le, ut = ev[i].ut;
        const int l = (int)(le & 0xffffu);
        if (l > k || !(ut & 0x10000u)) continue;
        if (ut & 0x20000u) acc_add(ac, (int)(ut & 0xffffu), (int)(le >> 16), ev[i].d1);
        else acc_add(af, (int)(ut & 0xffffu), (int)(le >> 16), ev[i].d1);
    }
    F.UP = max(af.fl, af.t2); F.EP = max(af.t1, af.fl); F.d1 = af.dd;
    C.UP = max(ac.fl, ac.t2); C.EP = max(ac.t1, ac.fl); C.d1 = ac.dd;
}

// Shared-memory carve-up of one CTA (sizes from ClassCfg); the same layout serves every window the CTA processes.
struct SmemView {
    uint8_t* Rbuf; uint16_t* lrp; uint16_t* MUP; uint16_t* MEP; uint32_t* R4; uint32_t* tab; uint8_t* QB; uint32_t* Q4; uint8_t* rowq;
    uint32_t* HQ; Ev* stg; Ev* evs; uint16_t* evoff; uint16_t* candK; uint16_t* candM; uint16_t* qoff; uint16_t* rowoff; uint16_t* qm;
    int* qcnt; int* qfill; int* s_int;
    __device__ void carve(unsigned char* smem, const ClassCfg& cfg, int nq) {
        size_t off = 0;
        Rbuf = smem + off; off += ClassCfg::al(cfg.n_cap + 16);       // the window, at its global misalignment
        lrp = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.n_cap);
        MUP = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.n_cap);
        MEP = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.n_cap);
        R4 = reinterpret_cast<uint32_t*>(smem + off); off += ClassCfg::al(4 * (size_t)cfg.n_cap);          // seed code of every window position
        tab = reinterpret_cast<uint32_t*>(smem + off); off += ClassCfg::al(8 * (size_t)cfg.n_cap);         // hash table of positions (2 n_cap slots)
        QB = smem + off; off += ClassCfg::al((size_t)cfg.qbuf);                                              // group query text
        Q4 = reinterpret_cast<uint32_t*>(smem + off); off += ClassCfg::al(8 * (size_t)cfg.rows_cap);       // seed rows: [2 row] = fwd, [2 row + 1] = rc
        rowq = smem + off; off += ClassCfg::al((size_t)cfg.rows_cap);                                        // group-local query of each row
        HQ = reinterpret_cast<uint32_t*>(smem + off); off += ClassCfg::al(4 * (size_t)cfg.hq_cap);         // seed hits: l | row << 12 | strand << 31
        stg = reinterpret_cast<Ev*>(smem + off); off += ClassCfg::al((size_t)cfg.stg_cap * sizeof(Ev));
        evs = reinterpret_cast<Ev*>(smem + off); off += ClassCfg::al((size_t)cfg.ev_cap * sizeof(Ev));
        evoff = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)(nq + 2));        // [nq+1] first event of each query
        candK = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.cand_cap);
        candM = reinterpret_cast<uint16_t*>(smem + off); off += ClassCfg::al(2 * (size_t)cfg.cand_cap);
        qoff = reinterpret_cast<uint16_t*>(smem + off);                                                     // [GROUP_MAX+1] byte offset of the query's text in QB
        rowoff = qoff + (GROUP_MAX + 1);                                                                    // [GROUP_MAX+1] first seed row
        qm = rowoff + (GROUP_MAX + 1);                                                                      // [GROUP_MAX] region length
        off += ClassCfg::al(2 * 3 * (size_t)(GROUP_MAX + 1));
        qcnt = reinterpret_cast<int*>(smem + off);                                                          // [GROUP_MAX] staged events per query
        qfill = qcnt + GROUP_MAX;
        off += ClassCfg::al(4 * 2 * (size_t)GROUP_MAX);
        // [0]=hit count [1]=staged events [2]=ncand [3]=overflow [4..5]=cand base (int64) [6]=group size
        s_int = reinterpret_cast<int*>(smem + off);
    }
};

// The complete search of ONE window by the whole CTA.  `s_bar` = two mbarriers initialised by the caller ([0] count 1: the window
// text; [1] count blockDim.x: a group's query texts); wphase / qphase carry their parities from window to window, so a CTA can
// process any number of windows.  On return (after a __syncthreads) sv.s_int[3] = overflow code (0 = none, 1 = a per-CTA
// capacity, 2 = the global candidate buffer), sv.s_int[2] = candidates written, sv.s_int[4..5] = their first slot in the global
// arrays, sv.candK[] = their window positions, sv.MEP[] = the folded end positions.
__device__ __forceinline__ void small_window(
    const uint8_t* __restrict__ text, const int64_t* __restrict__ gbase_fwd, const int64_t* __restrict__ gbase_rc,
    const int64_t* __restrict__ glen, int nq, const TaskDev tk, const int32_t* __restrict__ qs, const int32_t* __restrict__ ql,
    const ClassCfg& cfg, SmemView& sv, uint64_t* s_bar, uint8_t* s_mis, uint32_t& wphase, uint32_t& qphase,
    unsigned long long* __restrict__ cand_counter, unsigned long long cand_cap_global, int32_t* __restrict__ out_k,
    int32_t* __restrict__ out_lon, int32_t* __restrict__ out_sp, uint8_t* __restrict__ out_fwd) {
    const int T = blockDim.x;
    const int n = tk.n, minsize = tk.minsize;
    uint8_t* Rbuf = sv.Rbuf; uint16_t* lrp = sv.lrp; uint16_t* MUP = sv.MUP; uint16_t* MEP = sv.MEP; uint32_t* R4 = sv.R4; uint32_t* tab = sv.tab;
    uint8_t* QB = sv.QB; uint32_t* Q4 = sv.Q4; uint8_t* rowq = sv.rowq; uint32_t* HQ = sv.HQ; Ev* stg = sv.stg; Ev* evs = sv.evs;
    uint16_t* evoff = sv.evoff; uint16_t* candK = sv.candK; uint16_t* candM = sv.candM; uint16_t* qoff = sv.qoff; uint16_t* rowoff = sv.rowoff;
    uint16_t* qm = sv.qm; int* qcnt = sv.qcnt; int* qfill = sv.qfill; int* s_int = sv.s_int;
    const int tid = threadIdx.x;
    // seed length: consecutive seeds of a diagonal abut (step <= K) whenever minsize <= 19
    const int SEED_K = min(SEED_K_MAX, max(SEED_K_MIN, (minsize + 2) >> 1));
    const int step = max(1, minsize - SEED_K + 1);
    const bool abut = step <= SEED_K;
    int hbits = 1;
    while ((1 << hbits) < 2 * cfg.n_cap) ++hbits;
    const uint32_t hmask = (1u << hbits) - 1u;
    const uint8_t* R = Rbuf + (int)(tk.ref_off & 15);
    for (int i = tid; i < n; i += T) { MUP[i] = 0; MEP[i] = (uint16_t)n; lrp[i] = LRP_UNKNOWN; }
    for (int i = tid; i < (1 << hbits); i += T) tab[i] = SEED_PAD;
    if (tid < 8) s_int[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)ClassCfg::al((size_t)(tk.ref_off & 15) + (size_t)n);
        fence_proxy_async();
        mbar_arrive_expect_tx(&s_bar[0], bytes);
        bulk_g2s(Rbuf, text + (tk.ref_off & ~(int64_t)15), bytes, &s_bar[0]);
    }
    const int nseed = n >= SEED_K ? n - SEED_K + 1 : 0;            // reference positions holding a seed
    bool window_ready = false;

    int e0 = 0;                                                     // events stored so far (all earlier groups)
    int q0 = 0;
    while (q0 < nq) {
        // ---- group [q0, q0 + G): as many queries as fit the text / row buffers
        if (tid == 0) {
            int bytes = 0, rows = 0, g = 0;
            while (q0 + g < nq && g < GROUP_MAX) {
                const int m = ql[q0 + g];
                const int nb = 2 * (int)ClassCfg::al((size_t)m + 15);
                const int nr = 1 + (m >= SEED_K ? (m - SEED_K) / step + 1 : 0);           // one pad row in front of every query
                if (g > 0 && (bytes + nb > cfg.qbuf || rows + nr > cfg.rows_cap)) break;
                qoff[g] = (uint16_t)bytes; rowoff[g] = (uint16_t)(rows + 1); qm[g] = (uint16_t)m;
                qcnt[g] = 0;
                bytes += nb; rows += nr; ++g;
            }
            qoff[g] = (uint16_t)bytes; rowoff[g] = (uint16_t)(rows + 1);
            s_int[6] = g; s_int[0] = 0; s_int[1] = 0;
            if (bytes > cfg.qbuf || rows > cfg.rows_cap) s_int[3] = 1;      // (cannot happen for a correctly classified task)
        }
        __syncthreads();
        const int G = s_int[6];
        if (s_int[3]) break;
        // ---- 1. both strands of every query region of the group: thread g issues the two bulk copies of query g
        if (tid < G) {
            const int g = tid;
            const int m = qm[g];
            const int64_t gi = q0 + g + 1;
            const int64_t f_off = gbase_fwd[gi] + qs[q0 + g];
            const int64_t c_off = gbase_rc[gi] + (glen[gi] - qs[q0 + g] - m);
            const uint32_t mf = (uint32_t)(f_off & 15), mc = (uint32_t)(c_off & 15);
            const uint32_t bf = m ? (uint32_t)ClassCfg::al(mf + (size_t)m) : 0u, bc = m ? (uint32_t)ClassCfg::al(mc + (size_t)m) : 0u;
            s_mis[2 * g] = (uint8_t)mf; s_mis[2 * g + 1] = (uint8_t)mc;
            uint8_t* slot = QB + qoff[g];
            fence_proxy_async();
            mbar_arrive_expect_tx(&s_bar[1], bf + bc);
            if (bf) bulk_g2s(slot, text + (f_off & ~(int64_t)15), bf, &s_bar[1]);
            if (bc) bulk_g2s(slot + ClassCfg::al((size_t)m + 15), text + (c_off & ~(int64_t)15), bc, &s_bar[1]);
        } else {
            mbar_arrive(&s_bar[1]);
        }
        if (!window_ready) {
            // the window's seed codes and their hash table (the group's copies are in flight meanwhile)
            mbar_wait(&s_bar[0], wphase);
            for (int i = tid; i < nseed; i += T) {
                const uint32_t code = pack_seed(R + i, SEED_K);
                R4[i] = code;
                uint32_t h = seed_hash(code, hbits);
                while (atomicCAS(&tab[h], SEED_PAD, (uint32_t)i) != SEED_PAD) h = (h + 1) & hmask;
            }
            window_ready = true;
        }
        mbar_wait(&s_bar[1], qphase);
        qphase ^= 1u;
        // ---- 2. seed rows (row r0-1 of every query is a pad row: "no predecessor on the diagonal")
        const int total_rows = rowoff[G] - 1;
        for (int g = 0; g < G; ++g) {
            const int m = qm[g];
            const int r0 = rowoff[g], nr = rowoff[g + 1] - 1 - r0;
            const uint8_t* Qf = QB + qoff[g] + s_mis[2 * g];
            const uint8_t* Qc = QB + qoff[g] + ClassCfg::al((size_t)m + 15) + s_mis[2 * g + 1];
            for (int row = tid; row < nr; row += T) {
                Q4[2 * (r0 + row)] = pack_seed(Qf + row * step, SEED_K);
                Q4[2 * (r0 + row) + 1] = pack_seed(Qc + row * step, SEED_K);
                rowq[r0 + row] = (uint8_t)g;
            }
            if (tid == 0) { Q4[2 * (r0 - 1)] = SEED_PAD; Q4[2 * (r0 - 1) + 1] = SEED_PAD; }
        }
        __syncthreads();
        // ---- 3. every (row, strand) looks its seed up in the window's table -> hit queue
        for (int it = tid; it < 2 * total_rows; it += T) {
            const uint32_t code = Q4[it];
            if (code == SEED_PAD) continue;
            const uint32_t prev = Q4[it - 2];                   // the seed one row up the diagonal (pad row: never matches)
            uint32_t h = seed_hash(code, hbits);
            for (;;) {
                const uint32_t l = tab[h];
                if (l == SEED_PAD) break;
                h = (h + 1) & hmask;
                if (R4[l] != code) continue;
                // abutting seeds: the left extension reaches the seed spacing iff the predecessor seed is a hit too
                if (abut && (int)l >= step && R4[l - step] == prev) continue;
                const int slot = atomicAdd(&s_int[0], 1);
                if (slot < cfg.hq_cap) HQ[slot] = l | ((uint32_t)(it >> 1) << 12) | ((uint32_t)(it & 1) << 31);
            }
        }
        __syncthreads();
        const int nh = s_int[0];
        if (nh > cfg.hq_cap) { if (tid == 0) s_int[3] = 1; __syncthreads(); break; }
        // ---- 4. one thread per hit: extend left (de-duplication) and right -> staged MEM events
        for (int h = tid; h < nh; h += T) {
            const uint32_t hw = HQ[h];
            const int l = (int)(hw & 0xfffu), row = (int)((hw >> 12) & 0x7ffffu), strand = (int)(hw >> 31);
            const int g = rowq[row];
            const int m = qm[g];
            const int j = (row - rowoff[g]) * step;
            const uint8_t* Q = QB + qoff[g] + (strand ? ClassCfg::al((size_t)m + 15) : 0) + s_mis[2 * g + strand];
            const int cmax = min(step, min(j, l));
            const int c = smatch_bwd(Q + j, R + l, cmax);
            if (c >= step) continue;                      // the previous seed row lies in the same match
            const int emax = min(m - j, n - l);
            const int e = SEED_K + smatch_fwd(Q + j + SEED_K, R + l + SEED_K, emax - SEED_K);
            const int L = c + e, l0 = l - c;
            if (L < minsize) continue;
            const int slot = atomicAdd(&s_int[1], 1);
            if (slot < cfg.stg_cap) {
                stg[slot].le = (uint32_t)l0 | ((uint32_t)(l0 + L) << 16);
                stg[slot].ut = ((uint32_t)(strand << 1) | ((uint32_t)g << 2)) << 16;
                stg[slot].d1 = (j - c) - l0;
                atomicAdd(&qcnt[g], 1);
            }
        }
        __syncthreads();
        const int ns = s_int[1];
        if (ns > cfg.stg_cap || e0 + ns > cfg.ev_cap) { if (tid == 0) s_int[3] = 1; __syncthreads(); break; }
        // ---- 5. uniqueness; thread 0 lays out the group's slice of the event store meanwhile
        if (tid == 0) {
            int run = e0;
            for (int g = 0; g < G; ++g) { evoff[q0 + g] = (uint16_t)run; qfill[g] = run; run += qcnt[g]; }
            evoff[q0 + G] = (uint16_t)run;
        }
        resolve_unique(R, R4, n, SEED_K, lrp, stg, ns, T);
        __syncthreads();
        // ---- 6. counting sort by query into the event store
        for (int i = tid; i < ns; i += T) {
            const Ev e = stg[i];
            const int g = (int)(e.ut >> 18);
            const int pos = atomicAdd(&qfill[g], 1);
            evs[pos].le = e.le;
            evs[pos].ut = e.ut & 0x3ffffu;
            evs[pos].d1 = e.d1;
        }
        __syncthreads();
        // ---- 7. fold the group's queries (ini order) into Master
        for (int k = tid; k < n; k += T) {
            int mep = MEP[k], mup = MUP[k];
            for (int g = 0; g < G; ++g) {
                const int a0 = evoff[q0 + g], a1 = evoff[q0 + g + 1];
                StrandVal F, C;
                eval_at(evs + a0, a1 - a0, k, F, C);
                const int fe = min(mep, F.EP), ce = min(mep, C.EP);
                if (fe > ce) { mup = max(mup, F.UP); mep = fe; }
                else { mup = max(mup, C.UP); mep = ce; }
            }
            MUP[k] = (uint16_t)mup; MEP[k] = (uint16_t)mep;
        }
        e0 += ns;
        q0 += G;
        __syncthreads();
    }
    if (!window_ready) mbar_wait(&s_bar[0], wphase);   // (no query, or an early exit: never leave with the window copy in flight)
    wphase ^= 1u;
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
    // replay of the fold at the candidate positions from the stored events (strand flag + start per query)
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
    __syncthreads();
}

__global__ void __launch_bounds__(SM_MAX_THREADS, 4) small_region_kernel(
    const uint8_t* __restrict__ text, const int64_t* __restrict__ gbase_fwd, const int64_t* __restrict__ gbase_rc,
    const int64_t* __restrict__ glen, int nq, const TaskDev* __restrict__ tasks, const int32_t* __restrict__ qcoords,
    const int32_t* __restrict__ task_ids, int ntasks, ClassCfg cfg, TaskOut* __restrict__ outs,
    unsigned long long* __restrict__ cand_counter, unsigned long long cand_cap_global, int32_t* __restrict__ out_k,
    int32_t* __restrict__ out_lon, int32_t* __restrict__ out_sp, uint8_t* __restrict__ out_fwd) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint64_t s_bar[2];                          // [0]: the window text has arrived  [1]: a group's query texts have arrived
    __shared__ uint8_t s_mis[2 * GROUP_MAX];               // misalignment (0..15) of every staged strand: the string starts there in its slot
    const int ti = blockIdx.x;
    if (ti >= ntasks) return;
    const int task_id = task_ids[ti];
    const TaskDev tk = tasks[task_id];
    SmemView sv;
    sv.carve(smem, cfg, nq);
    if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], (uint32_t)blockDim.x); fence_mbar_init(); }
    uint32_t wphase = 0, qphase = 0;
    const int32_t* qs = qcoords + tk.qcoord_off;
    small_window(text, gbase_fwd, gbase_rc, glen, nq, tk, qs, qs + nq, cfg, sv, s_bar, s_mis, wphase, qphase, cand_counter, cand_cap_global,
                 out_k, out_lon, out_sp, out_fwd);
    if (threadIdx.x == 0) {
        const int ovf = sv.s_int[3];
        outs[task_id].ncand = ovf ? -ovf : sv.s_int[2];
        outs[task_id].cand_base = *reinterpret_cast<int64_t*>(&sv.s_int[4]);
        outs[task_id].pad = 0;
    }
}

}  // namespace small
}  // namespace pb200
