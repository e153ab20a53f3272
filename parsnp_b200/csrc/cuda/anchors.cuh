// The anchor stage of Aligner::setInitialClusters (src/parsnp.cpp:2121-2174) on the device, EXACTLY:
// setMums1's loop D over the candidates of the whole-genome region (src/parsnp.cpp:1713-1842) and determineRegion + the push
// rules for the regions between the anchors (src/parsnp.cpp:2153-2172).
//
// Why this can be exact in parallel: the layout is empty when the anchors are placed, so a candidate whose interval overlaps no
// other candidate's interval in any genome is untouched by the trim loop whatever the order (the only bits in its intervals
// would be its own).  The kernels below verify that premise for ALL valid candidates - starts ascending with the candidate
// order and no overlap, in every genome - and report anything else to the host, which then takes the candidates and runs its
// own (partly sequential) accept pass, exactly as in round 1 (host/aligner.cpp: accept_candidates_parallel).  On collinear
// genome sets nothing overlaps: the candidates never leave the device, the recursion (recursion.cuh) starts from the regions
// written here, and the host only receives the accepted anchors, the layout and the initial regions - while the GPU is
// already following the recursion.
#pragma once
#include "recursion.cuh"

namespace pb200 {
namespace anc {

struct Flags {                      // device counters / flags of one anchor stage
    unsigned int noncollinear;      // some genome's starts do not ascend with the candidate order
    unsigned int overlaps;          // valid candidates overlapping another valid candidate in some genome
    unsigned int accepted;          // anchors
    unsigned int regions;           // initial regions pushed
};

// one warp per candidate: coordinates in every genome (TMum ctor, src/TMum.cpp:13-72, on the whole-genome region: a
// reverse-strand start is mirrored on the genome length - correct here), the range pre-check of src/parsnp.cpp:1723 and the
// LON >= 5 rule.  ST = candidate-major, STT = genome-major copy for the per-genome sweeps.
__global__ void anchor_coords_kernel(int n, const int64_t* __restrict__ glen, unsigned int ncand, const int32_t* __restrict__ k,
                                     const int32_t* __restrict__ lon, const int32_t* __restrict__ sp, const uint8_t* __restrict__ fwd,
                                     int32_t* __restrict__ ST, int32_t* __restrict__ STT, uint8_t* __restrict__ valid) {
    const unsigned int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= ncand) return;
    const int nq = n - 1;
    const int64_t LON = lon[c];
    int fail = 0;
    for (int g = lane; g < n; g += 32) {
        const int64_t off = g == 0 ? k[c] : sp[(size_t)c * nq + (g - 1)];
        const bool f = g == 0 ? true : fwd[(size_t)c * nq + (g - 1)] != 0;
        const int64_t L = glen[g];
        if ((unsigned long long)(off + 1) > (unsigned long long)L) fail = 1;       // (dsp - 0) > length of the whole-genome region
        int64_t s = off;
        if (!f) s = L - (s + LON);
        if (s + LON > L || s < 0) fail = 1;
        ST[(size_t)c * n + g] = (int32_t)s;
        STT[(size_t)g * ncand + c] = (int32_t)s;
    }
    fail = __any_sync(0xffffffffu, fail);
    if (lane == 0) valid[c] = (!fail && LON >= 5) ? 1 : 0;
}

// one CTA per genome: do the valid candidates' starts ascend with the candidate order, and does any interval start before the
// end of an earlier one?  Two passes over the genome's column with block-wide prefix maxima.
__global__ void __launch_bounds__(1024) anchor_overlap_kernel(unsigned int ncand, const int32_t* __restrict__ STT, const int32_t* __restrict__ lon,
                                                              const uint8_t* __restrict__ valid, Flags* __restrict__ flags) {
    __shared__ long long s_ms[1024], s_me[1024];
    const int g = blockIdx.x, t = threadIdx.x, T = blockDim.x;
    const int32_t* col = STT + (size_t)g * ncand;
    const unsigned int per = (ncand + T - 1) / T;
    const unsigned int c0 = min(ncand, t * per), c1 = min(ncand, c0 + per);
    long long ms = -1, me = -1;
    for (unsigned int c = c0; c < c1; ++c)
        if (valid[c]) { ms = max(ms, (long long)col[c]); me = max(me, (long long)col[c] + lon[c]); }
    s_ms[t] = ms; s_me[t] = me;
    __syncthreads();
    for (int o = 1; o < T; o <<= 1) {                       // inclusive prefix maxima
        long long a = t >= o ? s_ms[t - o] : -1, b = t >= o ? s_me[t - o] : -1;
        __syncthreads();
        s_ms[t] = max(s_ms[t], a); s_me[t] = max(s_me[t], b);
        __syncthreads();
    }
    long long ps = t ? s_ms[t - 1] : -1, pe = t ? s_me[t - 1] : -1;
    unsigned int bad_order = 0, ov = 0;
    for (unsigned int c = c0; c < c1; ++c) {
        if (!valid[c]) continue;
        const long long s = col[c], e = s + lon[c];
        if (s < ps) bad_order = 1;
        if (s < pe) ++ov;
        ps = max(ps, s); pe = max(pe, e);
    }
    if (bad_order) atomicOr(&flags->noncollinear, 1u);
    if (ov) atomicAdd(&flags->overlaps, ov);
}

// one warp per candidate (nothing overlaps, nothing to trim): the reverse-strand check of src/parsnp.cpp:1800-1825 and the
// mumlayout update
__global__ void anchor_accept_kernel(int n, const uint8_t* __restrict__ text, const int64_t* __restrict__ gbase_fwd, unsigned int ncand,
                                     const int32_t* __restrict__ ST, const int32_t* __restrict__ lon, const uint8_t* __restrict__ fwd,
                                     const uint8_t* __restrict__ valid, unsigned long long* __restrict__ bits, const int64_t* __restrict__ bit_off,
                                     uint32_t* __restrict__ accepted) {
    const unsigned int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= ncand) return;
    if (lane == 0) accepted[c] = 0;
    if (!valid[c]) return;
    const int nq = n - 1;
    const int64_t length = lon[c];
    if (length < 2 || n <= 1) return;
    const int32_t* st = ST + (size_t)c * n;
    for (int g = 1; g < n; ++g) {
        if (fwd[(size_t)c * nq + (g - 1)]) continue;        // (uniform over the warp)
        const uint8_t* g0 = text + gbase_fwd[0] + st[0];
        const uint8_t* gk = text + gbase_fwd[g] + st[g];
        int bad = 0;
        for (int64_t x = lane; x < length; x += 32) {
            const uint8_t a = gk[length - 1 - x];
            bad |= (a < 4 ? (uint8_t)(3 - a) : (uint8_t)4) != g0[x];
        }
        if (__any_sync(0xffffffffu, bad)) return;
    }
    for (int g = lane; g < n; g += 32) rec::bits_set_range(bits + bit_off[g], st[g], st[g] + length);
    if (lane == 0) accepted[c] = 1;
}

// one warp per accepted anchor, on the FINAL layout (all anchors placed): its record for the host and determineRegion on both
// sides (src/parsnp.cpp:1199-1290).  REG[x] = lS[n] lE[n] rS[n] rE[n], SL[2x], SL[2x+1] = the two slengths
__global__ void anchor_regions_kernel(int n, const int64_t* __restrict__ glen, unsigned int ncand, const uint32_t* __restrict__ accepted,
                                      const uint32_t* __restrict__ xidx, const int32_t* __restrict__ ST, const int32_t* __restrict__ lon,
                                      const uint8_t* __restrict__ fwd, const unsigned long long* __restrict__ bits,
                                      const int64_t* __restrict__ bit_off, int32_t* __restrict__ REG, int32_t* __restrict__ SL,
                                      int32_t* __restrict__ a_st, int32_t* __restrict__ a_lon, uint8_t* __restrict__ a_fwd) {
    const unsigned int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= ncand || !accepted[c]) return;
    const unsigned int x = xidx[c];
    const int nq = n - 1;
    const int64_t length = lon[c];
    int64_t lsl = 500000000, rsl = 500000000;
    int32_t* reg = REG + (size_t)x * 4 * n;
    for (int g = lane; g < n; g += 32) {
        const int64_t s = ST[(size_t)c * n + g], L = glen[g];
        const unsigned long long* row = bits + bit_off[g];
        int64_t cp = rec::bits_prev_set(row, s - 1);
        if (cp < 0) cp = 0;
        const int64_t en = s + length;
        int64_t cq = en + 1;
        if (cq < L) cq = rec::bits_next_set(row, cq, L);
        reg[g] = (int32_t)(cp + 1); reg[n + g] = (int32_t)(s - 1);
        reg[2 * n + g] = (int32_t)(en + 1); reg[3 * n + g] = (int32_t)(cq - 1);
        lsl = min(lsl, (s - 1) - (cp + 1));
        rsl = min(rsl, (cq - 1) - (en + 1));
        a_st[(size_t)x * n + g] = (int32_t)s;
        a_fwd[(size_t)x * n + g] = g == 0 ? (uint8_t)1 : fwd[(size_t)c * nq + (g - 1)];
    }
    lsl = rec::warp_min64(lsl);
    rsl = rec::warp_min64(rsl);
    if (lane == 0) { SL[2 * x] = (int32_t)lsl; SL[2 * x + 1] = (int32_t)rsl; a_lon[x] = (int32_t)length; }
}

// the push rules of src/parsnp.cpp:2153-2172, one warp per anchor: the left region unless it equals the previous anchor's
// right region; the right region unless it equals the anchor's own left region; both only with slength > q
__global__ void anchor_push_flags_kernel(int n, int q, unsigned int nanchors, const int32_t* __restrict__ REG, const int32_t* __restrict__ SL,
                                         uint32_t* __restrict__ slot) {
    const unsigned int x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (x >= nanchors) return;
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

// the pushed regions into the recursion's store, in push order (= their ids); pair[id] = 1: region id (a right side) and id + 1
// (the next anchor's left side, one base earlier) cover the same gap
__global__ void anchor_push_kernel(int n, unsigned int nanchors, const int32_t* __restrict__ REG, const int32_t* __restrict__ SL,
                                   const uint32_t* __restrict__ slot, const uint32_t* __restrict__ pos, rec::Store St, uint8_t* __restrict__ pair,
                                   unsigned int cap) {
    const unsigned int x = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (x >= nanchors) return;
    const int32_t* me = REG + (size_t)x * 4 * n;
    for (int side = 0; side < 2; ++side) {
        if (!slot[2 * x + side]) continue;
        const unsigned int id = pos[2 * x + side];
        if (id >= cap) continue;
        const int32_t* S = me + 2 * n * side;
        int32_t* c = St.coords + (size_t)id * 2 * n;
        for (int g = lane; g < n; g += 32) { c[g] = S[g]; c[n + g] = S[n + g] - S[g]; }
        if (lane == 0) {
            int p = 0;
            if (side == 1 && x + 1 < nanchors && slot[2 * x + 2]) {
                const int32_t* nx = me + 4 * n;                 // left side of the next anchor
                p = nx[0] == S[0] - 1 && nx[n] == S[n];
            }
            pair[id] = (uint8_t)p;
        }
    }
}

// device -> device: the counters the recursion starts from
__global__ void anchor_finish_kernel(const uint32_t* __restrict__ accepted_total, const uint32_t* __restrict__ regions_total, Flags* flags,
                                     unsigned int* __restrict__ nregions, unsigned int cap) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        flags->accepted = *accepted_total;
        flags->regions = *regions_total;
        *nregions = min(*regions_total, cap);
    }
}

__global__ void add_offset_kernel(int32_t* __restrict__ dst, const uint32_t* __restrict__ src, unsigned int n, int32_t offset) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (int32_t)src[i] + offset;
}

// sentinel bit at position len of every genome's row (src/parsnp.cpp:3181-3186)
__global__ void layout_sentinels_kernel(int n, const int64_t* __restrict__ glen, unsigned long long* __restrict__ bits, const int64_t* __restrict__ bit_off) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) atomicOr(bits + bit_off[g] + (glen[g] >> 6), 1ull << (glen[g] & 63));
}

}  // namespace anc
}  // namespace pb200
