// Device-resident discovery of the recursion of Aligner::doWork (src/parsnp.cpp:173-317).
//
// The host's exact replay (host/replay.cpp) needs, for every region it will pop, the candidates of that region's window search.
// Which regions exist is only known after the candidates of their parents have been validated, trimmed against `mumlayout` and
// placed (setMums1 loop D, src/parsnp.cpp:1713-1842; trim 1399-1477; determineRegion 1199-1290).  Round 1 ran that discovery on
// the host, level by level, with a GPU batch + upload + download + gather per level.  Here the whole discovery stays on the
// device: one CTA searches a region's window in shared memory (smallpath.cuh) and then, with its first warp, runs the accept /
// trim / determineRegion rules on a SCRATCH copy of mumlayout in HBM and appends the child regions to the next level's work
// list itself.  The host enqueues a fixed number of levels without synchronising and reads everything back once.
//
// It is a PREDICTOR, exactly like the host pass it replaces: regions of one level are processed in no particular order, so a
// reverse-strand candidate (whose mirrored coordinates point into some other region) may see another state of
// the scratch layout than the reference would.  The replay recomputes every accept decision in the reference's order on the
// true layout and merely LOOKS UP the candidate lists by region coordinates; a region this pass did not predict is searched on
// demand.  Results are therefore exact whatever this pass does; it only has to be right almost always to be fast.
#pragma once
#include "smallpath.cuh"

namespace pb200 {
namespace rec {

constexpr int NCLASS = 3;
constexpr unsigned int FW_CAP = 65536;
// a work-list entry: region id, optionally the PAIR (id, id + 1) = the two regions that cover one gap (the left side of a MUM
// starts one base before the right side of its predecessor): the reference searches the one with the smaller start first and
// the second one then sees its bits, so one CTA takes both, in that order
constexpr int32_t E_PAIR = 1 << 30, E_SECOND_FIRST = 1 << 29, E_ID = (1 << 29) - 1;

// work lists of one run: level L reads list[L & 1][c][0 .. count[L & 1][c]) and appends children to list[(L + 1) & 1]
struct Queues {
    int32_t* list[2][NCLASS];           // region ids
    unsigned int* count;                // [2][NCLASS] entries appended
    unsigned int* taken;                // [2][NCLASS] entries handed out (dynamic scheduling inside a level): search kernels ...
    unsigned int* taken2;               //                                                                      ... accept kernel
    unsigned int* nregions;             // regions in the store
    unsigned int* ndeferred;            // regions the device could not take (too large, minsize < 4, ...): left to the host
    unsigned int* dropped;              // children lost to a full store / list (the replay searches them on demand)
    int32_t* deferred;                  // their ids
    unsigned int* nfw;                  // accepted MUMs written outside their region (reverse-strand genomes): count ...
    int32_t* fw;                        // ... and (genome, start, length) records, FW_CAP at most (more: the host takes nothing as final)
    unsigned int cap;                   // capacity of the region store and of every list
};

// one region: start[n] then len[n] at coords + id * 2n; everything else per region below
struct Store {
    int32_t* coords;
    int32_t* slen;                      // TRegion::slength
    int32_t* minsize;
    int32_t* ncand;                     // -1 = not searched (deferred / dropped), else candidates in the global arrays
    int64_t* cand_base;
    // what the host needs to take the accept decisions made here as FINAL where they cannot depend on the order (host/replay.cpp):
    uint32_t* flags;                    // F_* below
    int32_t* parent;                    // the region whose accepted MUM this one lies beside (-1: an initial region)
    int32_t* acc_shift;                 // per candidate (parallel to the global candidate arrays): -1 = not accepted, else the trim shift
    int32_t* acc_len;                   //                                                          its accepted length
};
// a region's accept pass is order-independent unless ...
constexpr uint32_t F_FOREIGN = 1;       // ... a candidate that reached the trim loop had a reverse-strand genome (coordinates mirrored on
                                        //     the whole genome: it read - and, if accepted, wrote - the layout somewhere else)
constexpr uint32_t F_SECOND = 2;        // ... it is the second region of a gap pair and accepted something (its sub-regions overlap the
                                        //     first one's)
constexpr uint32_t F_NONCOLLINEAR = 4;  // ... its accepted MUMs do not ascend in every genome (sub-regions overlap each other)
constexpr uint32_t F_REQUEUED = 8;      // ... it was one region of a gap pair and overflowed a per-CTA capacity: it is searched again later,
                                        //     on its own, i.e. not in the pair's order
constexpr uint32_t F_ORDER_MASK = 15;   // any of the above: the host replays the region's gap itself
constexpr uint32_t I_SECOND = 16;       // (information) the region was searched as the SECOND one of a gap pair, right after its mate
static_assert(F_ORDER_MASK == REC_ORDER_MASK && I_SECOND == REC_SECOND, "common.h: RecursionResult::flags");

struct Params {
    int n;                              // genomes
    int q;                              // ini [LCB] q: sub-regions need slength > q
    int64_t p;                          // ini [LCB] p: a region longer than this has several windows (host only)
    const int32_t* minsize_tab;         // minsize(slength) of the `mums` expression for slength < minsize_n
    int minsize_n;
    const int64_t* bit_off;             // per genome: first word of its row in `bits`
    unsigned long long* bits;           // scratch mumlayout (len + 1 bits per genome, sentinel at len)
    int n_cap[NCLASS], m_cap[NCLASS];
};

// ---- scratch layout access (L2: rows are written with atomics by other CTAs and by this one)
__device__ __forceinline__ unsigned long long ld_word(const unsigned long long* p) { return __ldcg(p); }
__device__ __forceinline__ bool bit_get(const unsigned long long* row, int64_t i) { return (ld_word(row + (i >> 6)) >> (i & 63)) & 1ull; }
// # consecutive set bits a, a+1, ... (< b)
__device__ inline int64_t bits_run_up(const unsigned long long* row, int64_t a, int64_t b) {
    int64_t i = a;
    while (i < b) {
        const unsigned long long inv = ~(ld_word(row + (i >> 6)) >> (i & 63));
        const int avail = 64 - (int)(i & 63);
        const int z = inv ? __ffsll((long long)inv) - 1 : 64;
        if (z < avail) { i += z; break; }
        i += avail;
    }
    if (i > b) i = b;
    return i - a;
}
// # consecutive set bits b-1, b-2, ... (>= a)
__device__ inline int64_t bits_run_down(const unsigned long long* row, int64_t a, int64_t b) {
    int64_t i = b - 1;
    while (i >= a) {
        const int pos = (int)(i & 63);
        const unsigned long long inv = ~(ld_word(row + (i >> 6)) << (63 - pos));
        const int z = inv ? __clzll((long long)inv) : 64;
        const int avail = pos + 1;
        if (z < avail) { i -= z; break; }
        i -= avail;
    }
    if (i < a - 1) i = a - 1;
    return (b - 1) - i;
}
__device__ inline int64_t bits_prev_set(const unsigned long long* row, int64_t i) {          // largest set index <= i, or -1
    if (i < 0) return -1;
    int64_t wi = i >> 6;
    unsigned long long cur = ld_word(row + wi) & (~0ull >> (63 - (i & 63)));
    for (;;) {
        if (cur) return (wi << 6) + 63 - __clzll((long long)cur);
        if (wi == 0) return -1;
        cur = ld_word(row + --wi);
    }
}
__device__ inline int64_t bits_next_set(const unsigned long long* row, int64_t i, int64_t limit) {   // smallest set index in [i, limit), or limit
    if (i >= limit) return limit;
    int64_t wi = i >> 6;
    const int64_t wl = (limit - 1) >> 6;
    unsigned long long cur = ld_word(row + wi) & (~0ull << (i & 63));
    for (;;) {
        if (cur) { const int64_t r = (wi << 6) + __ffsll((long long)cur) - 1; return r < limit ? r : limit; }
        if (wi >= wl) return limit;
        cur = ld_word(row + ++wi);
    }
}
__device__ inline void bits_set_range(unsigned long long* row, int64_t a, int64_t b) {       // [a, b)
    if (a >= b) return;
    const int64_t wa = a >> 6, wb = (b - 1) >> 6;
    const unsigned long long ma = ~0ull << (a & 63), mb = ~0ull >> (63 - ((b - 1) & 63));
    if (wa == wb) { atomicOr(row + wa, ma & mb); return; }
    atomicOr(row + wa, ma);
    for (int64_t w = wa + 1; w < wb; ++w) atomicOr(row + w, ~0ull);
    atomicOr(row + wb, mb);
}

__device__ __forceinline__ int warp_or(int v) { return __any_sync(0xffffffffu, v); }
__device__ __forceinline__ int64_t warp_min64(int64_t v) {
    for (int o = 16; o > 0; o >>= 1) { const int64_t x = __shfl_xor_sync(0xffffffffu, v, o); v = x < v ? x : v; }
    return v;
}

// class of a region (same rule as CudaEngine::classify), NCLASS = the device cannot take it
__device__ __forceinline__ int class_of(const Params& P, int64_t ref_len, int64_t max_m, int minsize) {
    if (minsize < 4 || ref_len > P.p || ref_len <= 0) return NCLASS;
    for (int c = 0; c < NCLASS; ++c)
        if (ref_len <= P.n_cap[c] && max_m <= P.m_cap[c]) return c;
    return NCLASS;
}

// class + minsize of region (S, E) (lanes hold genomes lane, lane + 32, ...)
template <int GPL>
__device__ inline int region_class(const Params& P, const int64_t (&S)[GPL], const int64_t (&E)[GPL], int64_t slength, int lane, int& ms) {
    const int n = P.n;
    int64_t mm = 0;
    for (int t = 0; t < GPL; ++t) { const int g = lane + 32 * t; if (g >= 1 && g < n) mm = max(mm, E[t] - S[t]); }
    mm = -warp_min64(-mm);
    const int64_t L0 = __shfl_sync(0xffffffffu, E[0] - S[0], 0);
    ms = slength >= 0 && slength < P.minsize_n ? P.minsize_tab[slength] : 0;
    return class_of(P, L0, mm, ms);
}
template <int GPL>
__device__ inline void write_region(const Params& P, const Store& St, unsigned int id, const int64_t (&S)[GPL], const int64_t (&E)[GPL], int64_t slength,
                                    int ms, int lane, int parent) {
    const int n = P.n;
    int32_t* c = St.coords + (size_t)id * 2 * n;
    for (int t = 0; t < GPL; ++t) {
        const int g = lane + 32 * t;
        if (g < n) { c[g] = (int32_t)S[t]; c[n + g] = (int32_t)(E[t] - S[t]); }
    }
    if (lane == 0) { St.slen[id] = (int32_t)slength; St.minsize[id] = ms; St.ncand[id] = -1; St.cand_base[id] = 0; St.flags[id] = 0; St.parent[id] = parent; }
}
__device__ inline void enqueue(const Queues& Q, int next, int cls, int32_t entry) {        // one lane
    if (cls < NCLASS) {
        const unsigned int at = atomicAdd(&Q.count[next * NCLASS + cls], 1u);
        if (at < Q.cap) Q.list[next][cls][at] = entry;
        else { atomicSub(&Q.count[next * NCLASS + cls], 1u); atomicAdd(Q.dropped, 1u); }
    } else {
        const unsigned int at = atomicAdd(Q.ndeferred, 1u);
        if (at < Q.cap) Q.deferred[at] = entry & E_ID;
    }
}
// append region (S[], E[]) to the store and to the next level's list of its class
template <int GPL>
__device__ inline void push_region(const Params& P, const Store& St, const Queues& Q, int next, const int64_t (&S)[GPL], const int64_t (&E)[GPL],
                                   int64_t slength, int lane, int parent) {
    int ms;
    const int cls = region_class<GPL>(P, S, E, slength, lane, ms);
    unsigned int id = 0;
    if (lane == 0) id = atomicAdd(Q.nregions, 1u);
    id = __shfl_sync(0xffffffffu, id, 0);
    if (id >= Q.cap) { if (lane == 0) { atomicSub(Q.nregions, 1u); atomicAdd(Q.dropped, 1u); } return; }
    write_region<GPL>(P, St, id, S, E, slength, ms, lane, parent);
    if (lane == 0) enqueue(Q, next, cls, (int32_t)id);
}
// the two regions of one gap: A (smaller start[0], searched first) and B, as ONE entry when the device can take both
template <int GPL>
__device__ inline void push_pair(const Params& P, const Store& St, const Queues& Q, int next, const int64_t (&SA)[GPL], const int64_t (&EA)[GPL],
                                 int64_t slA, const int64_t (&SB)[GPL], const int64_t (&EB)[GPL], int64_t slB, int lane, int parent) {
    int msA, msB;
    const int clsA = region_class<GPL>(P, SA, EA, slA, lane, msA);
    const int clsB = region_class<GPL>(P, SB, EB, slB, lane, msB);
    if (clsA >= NCLASS || clsB >= NCLASS) {
        push_region<GPL>(P, St, Q, next, SA, EA, slA, lane, parent);
        push_region<GPL>(P, St, Q, next, SB, EB, slB, lane, parent);
        return;
    }
    unsigned int id = 0;
    if (lane == 0) id = atomicAdd(Q.nregions, 2u);
    id = __shfl_sync(0xffffffffu, id, 0);
    if (id + 1 >= Q.cap) { if (lane == 0) { atomicSub(Q.nregions, 2u); atomicAdd(Q.dropped, 2u); } return; }
    write_region<GPL>(P, St, id, SA, EA, slA, msA, lane, parent);
    write_region<GPL>(P, St, id + 1, SB, EB, slB, msB, lane, parent);
    if (lane == 0) enqueue(Q, next, max(clsA, clsB), (int32_t)id | E_PAIR);
}

// setMums1 loop D + determineRegion for the candidates of one searched region, by ONE WARP, on the scratch layout.
// GPL = genomes per lane (n <= 32 * GPL).  cand arrays: the region's candidates at [base, base + nc) of the global arrays.
template <int GPL>
__device__ inline void accept_region(const Params& P, const Store& St, const Queues& Q, int next, const uint8_t* __restrict__ text,
                                     const int64_t* __restrict__ gbase_fwd, const int64_t* __restrict__ glen, int region, int nc, int64_t base,
                                     const int32_t* __restrict__ out_k, const int32_t* __restrict__ out_lon, const int32_t* __restrict__ out_sp,
                                     const uint8_t* __restrict__ out_fwd, bool second) {
    const int lane = threadIdx.x & 31;
    const int n = P.n, nq = n - 1;
    uint32_t rflags = 0;
    for (int c = lane; c < nc; c += 32) St.acc_shift[base + c] = -1;
    __syncwarp();
    const int32_t* rc = St.coords + (size_t)region * 2 * n;
    int64_t rs[GPL], rl[GPL], gl[GPL];
    const unsigned long long* row[GPL];
    for (int t = 0; t < GPL; ++t) {
        const int g = lane + 32 * t;
        rs[t] = g < n ? rc[g] : 0;
        rl[t] = g < n ? rc[n + g] : 0;
        gl[t] = g < n ? glen[g] : 0;
        row[t] = P.bits + (g < n ? P.bit_off[g] : 0);
    }
    int nacc = 0;
    int pc = -1;                                                // the previous accepted candidate, its shift and length
    int64_t pshift = 0, plen = 0;
    for (int c = 0; c < nc; ++c) {
        const int64_t LON = out_lon[base + c];
        const int k = out_k[base + c];
        int64_t st[GPL];
        bool fw[GPL];
        int fail = 0;
        for (int t = 0; t < GPL; ++t) {
            const int g = lane + 32 * t;
            st[t] = 0; fw[t] = true;
            if (g >= n) continue;
            int64_t off = g == 0 ? k : out_sp[(size_t)(base + c) * nq + (g - 1)];
            fw[t] = g == 0 ? true : out_fwd[(size_t)(base + c) * nq + (g - 1)] != 0;
            // range pre-check in unsigned arithmetic (src/parsnp.cpp:1723): (dsp - rs) > length, dsp = off + 1 + rs
            if ((unsigned long long)(off + 1) > (unsigned long long)(unsigned int)rl[t]) fail = 1;
            int64_t s = rs[t] + off;
            if (!fw[t]) s = gl[t] - (s + LON);                  // TMum ctor: mirrored on the WHOLE genome (src/TMum.cpp:35)
            if (s + LON > gl[t] || s < 0) fail = 1;
            st[t] = s;
        }
        if (warp_or(fail) || LON < 5) continue;
        int rev = 0;
        for (int t = 0; t < GPL; ++t) rev |= (lane + 32 * t < n) && !fw[t];
        rev = warp_or(rev);
        if (rev) rflags |= F_FOREIGN;
        // trim (src/parsnp.cpp:1399-1477): genome after genome, every trim shifts ALL genomes.  Only the two ends of an interval
        // are looked at, so when no genome has a set bit at either end (the rule inside a gap) nothing moves.
        int64_t length = LON, shift = 0;
        int touch = 0;
        for (int t = 0; t < GPL; ++t) {
            const int g = lane + 32 * t;
            if (g < n) touch |= (int)bit_get(row[t], st[t]) | (int)bit_get(row[t], st[t] + LON - 1);
        }
        if (warp_or(touch)) {
            for (int g = 0; g < n && length > 0; ++g) {
                const int t = g >> 5, owner = g & 31;
                int64_t t1 = 0, t2 = 0;
                if (lane == owner) t1 = bits_run_up(row[t], st[t] + shift, st[t] + shift + length);
                t1 = __shfl_sync(0xffffffffu, t1, owner);
                shift += t1; length -= t1;
                if (lane == owner) t2 = bits_run_down(row[t], st[t] + shift, st[t] + shift + length);
                t2 = __shfl_sync(0xffffffffu, t2, owner);
                length -= t2;
            }
        }
        if (length < 2) continue;
        // reverse-strand genomes are verified against the reference substring (src/parsnp.cpp:1800-1825)
        const int64_t s0 = __shfl_sync(0xffffffffu, st[0] + shift, 0);
        int badmum = 0;
        for (int g = 1; g < n && rev; ++g) {
            const int t = g >> 5, owner = g & 31;
            const int isrev = __shfl_sync(0xffffffffu, (int)!fw[t], owner);
            if (!isrev) continue;
            const int64_t sg = __shfl_sync(0xffffffffu, st[t] + shift, owner);
            const uint8_t* g0 = text + gbase_fwd[0] + s0;
            const uint8_t* gk = text + gbase_fwd[g] + sg;
            int bad = 0;
            for (int64_t x = lane; x < length; x += 32) {
                const uint8_t a = gk[length - 1 - x];
                bad |= (a < 4 ? (uint8_t)(3 - a) : (uint8_t)4) != g0[x];
            }
            if (warp_or(bad)) { badmum = 1; break; }
        }
        if (badmum) continue;
        for (int t = 0; t < GPL; ++t) {
            const int g = lane + 32 * t;
            if (g < n) bits_set_range(const_cast<unsigned long long*>(row[t]), st[t] + shift, st[t] + shift + length);
        }
        if (rev) {                                              // a write outside the region: the host has to know where
            for (int t = 0; t < GPL; ++t) {
                const int g = lane + 32 * t;
                if (g < n && !fw[t]) {
                    const unsigned int at = atomicAdd(Q.nfw, 1u);
                    if (at < FW_CAP) { Q.fw[3 * at] = g; Q.fw[3 * at + 1] = (int32_t)(st[t] + shift); Q.fw[3 * at + 2] = (int32_t)length; }
                }
            }
        }
        // accepted MUMs must ascend, without overlap, in every genome - else the sub-regions between them overlap
        if (nacc > 0) {
            int bad = 0;
            for (int t = 0; t < GPL; ++t) {
                const int g = lane + 32 * t;
                if (g >= n) continue;
                const int64_t poff = g == 0 ? out_k[base + pc] : out_sp[(size_t)(base + pc) * nq + (g - 1)];
                const int64_t pend = rs[t] + poff + pshift + plen;     // (forward: reverse-strand ones are flagged anyway)
                if (st[t] + shift < pend) bad = 1;
            }
            if (warp_or(bad)) rflags |= F_NONCOLLINEAR;
        }
        if (lane == 0) { St.acc_shift[base + c] = (int32_t)shift; St.acc_len[base + c] = (int32_t)length; }
        pc = c; pshift = shift; plen = length;
        ++nacc;
    }
    if (second && nacc > 0) rflags |= F_SECOND;
    if (lane == 0 && rflags) St.flags[region] |= rflags;
    if (nacc == 0) return;
    __threadfence();                                            // (this warp's own atomics are ordered before its reads below)
    __syncwarp();
    // determineRegion around every accepted MUM, on the layout as it is after ALL accepts of the region (src/parsnp.cpp:251-289).
    // The right side of MUM a-1 and the left side of MUM a are the same gap: pushed as a pair, left side first.
    int64_t pS[GPL], pE[GPL], psl = -1;                         // pending right side of the previous MUM (psl <= q: none)
    for (int c = 0; c < nc; ++c) {
        const int64_t shift = __ldcg(St.acc_shift + base + c);  // (written by lane 0 above: L2)
        if (shift < 0) continue;
        const int64_t length = __ldcg(St.acc_len + base + c);
        const int64_t LON = out_lon[base + c];
        const int k = out_k[base + c];
        int64_t lS[GPL], lE[GPL], rS[GPL], rE[GPL];
        int64_t lsl = 500000000, rsl = 500000000;
        for (int t = 0; t < GPL; ++t) {
            const int g = lane + 32 * t;
            lS[t] = lE[t] = rS[t] = rE[t] = 0;
            if (g >= n) continue;
            const int64_t off = g == 0 ? k : out_sp[(size_t)(base + c) * nq + (g - 1)];
            const bool f = g == 0 ? true : out_fwd[(size_t)(base + c) * nq + (g - 1)] != 0;
            int64_t s = rs[t] + off;
            if (!f) s = gl[t] - (s + LON);
            s += shift;
            int64_t cp = bits_prev_set(row[t], s - 1);
            if (cp < 0) cp = 0;
            lS[t] = cp + 1; lE[t] = s - 1;
            const int64_t en = s + length;
            int64_t cq = en + 1;
            if (cq < gl[t]) cq = bits_next_set(row[t], cq, gl[t]);
            rS[t] = en + 1; rE[t] = cq - 1;
            lsl = min(lsl, lE[t] - lS[t]);
            rsl = min(rsl, rE[t] - rS[t]);
        }
        lsl = warp_min64(lsl);
        rsl = warp_min64(rsl);
        if (lsl > P.q && psl > P.q) push_pair<GPL>(P, St, Q, next, lS, lE, lsl, pS, pE, psl, lane, region);
        else if (lsl > P.q) push_region<GPL>(P, St, Q, next, lS, lE, lsl, lane, region);
        else if (psl > P.q) push_region<GPL>(P, St, Q, next, pS, pE, psl, lane, region);
        for (int t = 0; t < GPL; ++t) { pS[t] = rS[t]; pE[t] = rE[t]; }
        psl = rsl;
    }
    if (psl > P.q) push_region<GPL>(P, St, Q, next, pS, pE, psl, lane, region);
}

// One level of one size class: CTAs take regions from the level's list until it is empty (dynamic scheduling), search the
// region's window (small_window) and run accept_region on the result.  A window that overflows a per-CTA capacity moves to the
// next class (next level's list); one that does not fit the global candidate buffer is left to the host.
template <int GPL>
__global__ void __launch_bounds__(small::SM_MAX_THREADS, 4) recursion_level_kernel(
    const uint8_t* __restrict__ text, const int64_t* __restrict__ gbase_fwd, const int64_t* __restrict__ gbase_rc,
    const int64_t* __restrict__ glen, Params P, Store St, Queues Q, int level, int cls, small::ClassCfg cfg,
    unsigned long long* __restrict__ cand_counter, unsigned long long cand_cap_global, int32_t* __restrict__ out_k,
    int32_t* __restrict__ out_lon, int32_t* __restrict__ out_sp, uint8_t* __restrict__ out_fwd) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint64_t s_bar[2];
    __shared__ uint8_t s_mis[2 * small::GROUP_MAX];
    __shared__ int s_next;
    const int cur = level & 1, next = cur ^ 1;
    if (Q.count[cur * NCLASS + cls] == 0) return;               // (most launches of the deeper levels: an empty list costs a launch, nothing else)
    const int nq = P.n - 1;
    small::SmemView sv;
    sv.carve(smem, cfg, nq);
    if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], (uint32_t)blockDim.x); fence_mbar_init(); }
    uint32_t wphase = 0, qphase = 0;
    const unsigned int total = min(Q.count[cur * NCLASS + cls], Q.cap);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_next = (int)atomicAdd(&Q.taken[cur * NCLASS + cls], 1u);
        __syncthreads();
        const unsigned int ti = (unsigned int)s_next;
        if (ti >= total) break;
        const int32_t entry = Q.list[cur][cls][ti];
        const int first = (entry & E_ID) + ((entry & E_PAIR) && (entry & E_SECOND_FIRST) ? 1 : 0);
        const int second = (entry & E_PAIR) ? (entry & E_ID) + ((entry & E_SECOND_FIRST) ? 0 : 1) : -1;
        for (int half = 0; half < 2; ++half) {
            const int region = half == 0 ? first : second;
            if (region < 0) break;
            if (half == 1) __syncthreads();                     // (every thread has read the first window's counters)
            const int32_t* rc = St.coords + (size_t)region * 2 * P.n;
            small::TaskDev tk;
            tk.ref_off = gbase_fwd[0] + rc[0];
            tk.n = rc[P.n];
            tk.minsize = St.minsize[region];
            tk.qcoord_off = 0;
            small::small_window(text, gbase_fwd, gbase_rc, glen, nq, tk, rc + 1, rc + P.n + 1, cfg, sv, s_bar, s_mis, wphase, qphase, cand_counter,
                                cand_cap_global, out_k, out_lon, out_sp, out_fwd);
            const int ovf = sv.s_int[3];
            const int nc = sv.s_int[2];
            const int64_t base = *reinterpret_cast<int64_t*>(&sv.s_int[4]);
            if (ovf) {
                // a per-CTA capacity: this region alone goes to the next class (next level's list); anything else: the host
                if (threadIdx.x == 0) {
                    if (entry & E_PAIR) St.flags[region] |= F_REQUEUED;
                    enqueue(Q, next, ovf == 1 && cls + 1 < NCLASS ? cls + 1 : NCLASS, (int32_t)region);
                }
                continue;
            }
            if (threadIdx.x == 0) { St.ncand[region] = nc; St.cand_base[region] = base; }
        }
    }
}

// The accept half of a level: ONE WARP per work-list entry runs loop D + determineRegion (accept_region) on what the search
// kernels of this level left in the candidate arrays - the two regions of a pair in the reference's order.  Its own kernel
// because it is a chain of dependent L2 round trips (bit words of the scratch layout): as the tail of the search CTA it
// left three of four warps waiting at a barrier (ncu: 12 barrier stalls per issue); here every SM keeps 32+ such chains in flight.
template <int GPL>
__global__ void __launch_bounds__(128) recursion_accept_kernel(const uint8_t* __restrict__ text, const int64_t* __restrict__ gbase_fwd,
                                                               const int64_t* __restrict__ glen, Params P, Store St, Queues Q, int level,
                                                               const int32_t* __restrict__ out_k, const int32_t* __restrict__ out_lon,
                                                               const int32_t* __restrict__ out_sp, const uint8_t* __restrict__ out_fwd) {
    const int cur = level & 1, next = cur ^ 1;
    const int lane = threadIdx.x & 31;
    for (int cls = 0; cls < NCLASS; ++cls) {
        const unsigned int total = min(Q.count[cur * NCLASS + cls], Q.cap);
        // (a warp that can see the list is empty or handed out leaves without touching the counter: 4 736 warps hammering one
        //  address made an EMPTY level cost 19 us)
        for (;;) {
            unsigned int ti = 0xffffffffu;
            if (lane == 0 && __ldcg(&Q.taken2[cur * NCLASS + cls]) < total) ti = atomicAdd(&Q.taken2[cur * NCLASS + cls], 1u);
            ti = __shfl_sync(0xffffffffu, ti, 0);
            if (ti >= total) break;
            const int32_t entry = Q.list[cur][cls][ti];
            const int first = (entry & E_ID) + ((entry & E_PAIR) && (entry & E_SECOND_FIRST) ? 1 : 0);
            const int second = (entry & E_PAIR) ? (entry & E_ID) + ((entry & E_SECOND_FIRST) ? 0 : 1) : -1;
            for (int half = 0; half < 2; ++half) {
                const int region = half == 0 ? first : second;
                if (region < 0) break;
                if (half == 1 && lane == 0) St.flags[region] |= I_SECOND;
                const int nc = St.ncand[region];                // (-1: it overflowed a capacity and waits in the next level's list)
                if (nc > 0) accept_region<GPL>(P, St, Q, next, text, gbase_fwd, glen, region, nc, St.cand_base[region], out_k, out_lon, out_sp, out_fwd, half == 1);
                __syncwarp();
            }
        }
    }
}

// between two levels: the finished level's counters are cleared for re-use two levels later
__global__ void level_advance_kernel(Queues Q, int level) {
    const int cur = level & 1;
    if (threadIdx.x < NCLASS) { Q.count[cur * NCLASS + threadIdx.x] = 0; Q.taken[cur * NCLASS + threadIdx.x] = 0; Q.taken2[cur * NCLASS + threadIdx.x] = 0; }
}

// level 0: classify the initial regions (uploaded into the store by the host) into the first lists.  pair[id] = 1: regions id
// and id + 1 are the right side of an anchor and the left side of the next one - one entry, the second searched first
__global__ void seed_lists_kernel(Params P, Store St, Queues Q, const unsigned int* __restrict__ count_ptr, const uint8_t* __restrict__ pair) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    const int count = (int)*count_ptr;
    if (id >= count) return;
    const int n = P.n;
    const int32_t* c = St.coords + (size_t)id * 2 * n;
    int64_t mm = 0, sl = 500000000;
    for (int g = 0; g < n; ++g) { if (g) mm = max(mm, (int64_t)c[n + g]); sl = min(sl, (int64_t)c[n + g]); }
    const int ms = sl >= 0 && sl < P.minsize_n ? P.minsize_tab[sl] : 0;
    St.slen[id] = (int32_t)sl;
    St.minsize[id] = ms;
    St.ncand[id] = -1;
    St.cand_base[id] = 0;
    St.flags[id] = 0;
    St.parent[id] = -1;
    int cls = class_of(P, c[n], mm, ms);
    const bool second = id > 0 && pair[id - 1];
    bool head = pair[id] && id + 1 < count;
    if (second || head) {
        // class of the mate (recomputed: both threads must take the same decision)
        const int o = second ? id - 1 : id + 1;
        const int32_t* d = St.coords + (size_t)o * 2 * n;
        int64_t mo = 0, so = 500000000;
        for (int g = 0; g < n; ++g) { if (g) mo = max(mo, (int64_t)d[n + g]); so = min(so, (int64_t)d[n + g]); }
        const int mso = so >= 0 && so < P.minsize_n ? P.minsize_tab[so] : 0;
        const int co = class_of(P, d[n], mo, mso);
        if (cls < NCLASS && co < NCLASS) {
            if (second) return;                                 // the head enqueues the pair
            enqueue(Q, 0, max(cls, co), (int32_t)id | E_PAIR | E_SECOND_FIRST);
            return;
        }
    }
    enqueue(Q, 0, cls, (int32_t)id);
}

// ---- after the last level: regions in ascending start[0] order (the order in which the replay pops them, so its lookups and
// its candidate reads stream through memory), candidates regrouped to match
__global__ void region_keys_kernel(Store St, int n, unsigned int nr, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nr) return;
    keys[i] = (uint32_t)St.coords[(size_t)i * 2 * n];
    vals[i] = i;
}
// + the inverse permutation (store id -> sorted position) for the parent links
__global__ void sorted_counts_kernel(Store St, const uint32_t* __restrict__ perm, unsigned int nr, uint32_t* __restrict__ cnt, uint32_t* __restrict__ inv) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nr) return;
    const int c = St.ncand[perm[i]];
    cnt[i] = c > 0 ? (uint32_t)c : 0u;
    inv[perm[i]] = i;
}
// one warp per region: its record in the host's final form (int64 start / end coordinates, slength, window record, coordinate
// hash) and its candidates, into sorted position
__global__ void gather_sorted_kernel(Store St, int n, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ newbase, unsigned int nr,
                                     const int32_t* __restrict__ k, const int32_t* __restrict__ lon, const int32_t* __restrict__ sp,
                                     const uint8_t* __restrict__ fwd, int64_t* __restrict__ o_coords, int64_t* __restrict__ o_slen,
                                     WindowRec* __restrict__ o_wins, uint64_t* __restrict__ o_hash, int32_t* __restrict__ o_k,
                                     int32_t* __restrict__ o_lon, int32_t* __restrict__ o_sp, uint8_t* __restrict__ o_fwd,
                                     const uint32_t* __restrict__ inv, uint32_t* __restrict__ o_flags, int32_t* __restrict__ o_parent,
                                     int32_t* __restrict__ o_shift, int32_t* __restrict__ o_alen) {
    const unsigned int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= nr) return;
    const unsigned int r = perm[i];
    const int nq = n - 1;
    const int32_t* c = St.coords + (size_t)r * 2 * n;
    int64_t* oc = o_coords + (size_t)i * 2 * n;
    for (int g = lane; g < n; g += 32) { const int64_t s = c[g]; oc[g] = s; oc[n + g] = s + c[n + g]; }
    const int nc = St.ncand[r];
    const int64_t ob = St.cand_base[r], nb = newbase[i];
    __syncwarp();
    if (lane == 0) {
        o_slen[i] = St.slen[r];
        WindowRec w;
        w.ref_start = c[0]; w.ref_len = c[n]; w.cand_off = nb; w.ncand = nc; w.chunk = 0;
        o_wins[i] = w;
        const int64_t h4[4] = {(int64_t)c[0], (int64_t)c[n - 1], (int64_t)c[0] + c[n], (int64_t)c[n - 1] + c[2 * n - 1]};
        // region_coords_hash over (start[0], start[n-1], end[0], end[n-1]) - the same four values, the same arithmetic
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (int t = 0; t < 4; ++t) {
            h ^= (uint64_t)h4[t] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
            h *= 0xff51afd7ed558ccdull;
            h ^= h >> 29;
        }
        o_hash[i] = h;
        o_flags[i] = St.flags[r];
        const int par = St.parent[r];
        o_parent[i] = par < 0 ? -1 : (int32_t)inv[par];
    }
    if (nc <= 0) return;
    for (int x = lane; x < nc; x += 32) {
        o_k[nb + x] = k[ob + x]; o_lon[nb + x] = lon[ob + x];
        o_shift[nb + x] = St.acc_shift[ob + x]; o_alen[nb + x] = St.acc_len[ob + x];
    }
    const int64_t rows = (int64_t)nc * nq;
    for (int64_t x = lane; x < rows; x += 32) { o_sp[nb * nq + x] = sp[ob * nq + x]; o_fwd[nb * nq + x] = fwd[ob * nq + x]; }
}

}  // namespace rec
}  // namespace pb200
