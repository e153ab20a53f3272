// Large-window search: suffix index of the reference window + seed-and-extend MEM scan of every query strand +
// (floor, top1, top2) scan + ordered fold over queries + candidate emission.
//
// Replaces, for one reference window of Aligner::setMums1 (src/parsnp.cpp:1570-1695):
//   build_CSG + find_leaves (src/csgmum/csg.c:448-575)   -> suffix array by packed 21-mer radix sort + prefix doubling,
//                                                            adjacent LCP, lrp[l] = longest repeated prefix (u[l] = l + lrp[l])
//   Find_UM + Test_UM (src/csgmum/mum.c:27-45,177-250)    -> sampled k-mer seeds (SA bucket table) + left/right extension
//                                                            = maximal exact matches with L >= minsize and L > lrp[l]
//   Intersect_UM (src/csgmum/mum.c:125-175)               -> associative scan of (fl, t1, t2, diag) over events sorted by l
//   Master update + Merge_Master (mum.c:92-123)           -> per-position fold over queries in ini order (ties -> reverse)
//   emission loop (src/parsnp.cpp:1633-1695)              -> flag + ordered compaction, then per-candidate SP/strand pass
// Events shorter than minsize are pruned (exactness-preserving).
#pragma once
#include <algorithm>
#include "util.cuh"
#include "radix_sort.cuh"
#include "scan_prims.cuh"

namespace pb200 {
namespace big {

constexpr int KEY_BASES = 21;           // 3 bits per base (code+1; 0 = past the end of the window)
constexpr int MAX_SEED_K = 12;

struct StrandDesc { const uint8_t* q; int32_t m; int32_t pad; };

// number of equal leading bytes of a[0..limit) and b[0..limit)
__device__ __forceinline__ int match_len(const uint8_t* a, const uint8_t* b, int limit) {
    int t = 0;
    // long matches: 4 independent 8-byte compares per step, so a 100-base extension is ~4 dependent round trips, not 13
    while (t + 32 <= limit) {
        uint64_t x0 = load8u(a + t) ^ load8u(b + t), x1 = load8u(a + t + 8) ^ load8u(b + t + 8);
        uint64_t x2 = load8u(a + t + 16) ^ load8u(b + t + 16), x3 = load8u(a + t + 24) ^ load8u(b + t + 24);
        if (x0 | x1 | x2 | x3) {
            if (x0) return t + ((__ffsll((long long)x0) - 1) >> 3);
            if (x1) return t + 8 + ((__ffsll((long long)x1) - 1) >> 3);
            if (x2) return t + 16 + ((__ffsll((long long)x2) - 1) >> 3);
            return t + 24 + ((__ffsll((long long)x3) - 1) >> 3);
        }
        t += 32;
    }
    while (t < limit) {
        uint64_t x = load8u(a + t) ^ load8u(b + t);
        if (x) { t += (__ffsll((long long)x) - 1) >> 3; break; }
        t += 8;
    }
    return t < limit ? t : limit;
}

// ------------------------------------------------------------------ index build
__global__ void make_keys_kernel(const uint8_t* __restrict__ R, int n, uint64_t* __restrict__ keys, uint32_t* __restrict__ sa) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = 0;
#pragma unroll
    for (int t = 0; t < KEY_BASES; ++t) {
        uint64_t c = (i + t < n) ? (uint64_t)(R[i + t] + 1) : 0ull;
        key = (key << 3) | c;
    }
    keys[i] = key;
    sa[i] = (uint32_t)i;
}

// N-free windows: 16 bases at 2 bits per base in a 32-bit key (past the end = 0; the prefix doubling orders the few suffixes
// shorter than 16 correctly because their second key is 0 = "before everything")
constexpr int KEY2_BASES = 16;
__global__ void make_keys2_kernel(const uint8_t* __restrict__ R, int n, uint32_t* __restrict__ keys, uint32_t* __restrict__ sa) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t w0 = load8u(R + i), w1 = load8u(R + i + 8);
    uint32_t key = 0;
#pragma unroll
    for (int t = 0; t < KEY2_BASES; ++t) {
        uint32_t c = (uint32_t)((t < 8 ? (w0 >> (8 * t)) : (w1 >> (8 * (t - 8)))) & 3u);
        if (i + t >= n) c = 0;
        key = (key << 2) | c;
    }
    keys[i] = key;
    sa[i] = (uint32_t)i;
}
__device__ __forceinline__ bool kmer_of_key(uint64_t key, int k, uint32_t& code) {
    code = 0;
    for (int t = 0; t < k; ++t) {
        uint32_t v = (uint32_t)(key >> (3 * (KEY_BASES - 1 - t))) & 7u;
        if (v < 1u || v > 4u) return false;
        code = (code << 2) | (v - 1u);
    }
    return true;
}
// seed table entry (written by index_finish_kernel): k-mer absent = (0, 0).  Bucket of several suffixes: .x = first SA slot,
// .y = one past the last SA slot.  Bucket of ONE suffix (the common case): .x = the text position itself (saves the SA read
// in the scan) and .y = bit 31 + a 24-bit signature = the 4 bases before and the 4 bases after the k-mer (3 bits each,
// 7 = outside the window): a match of >= k+7 bases around the seed must agree with the reference on one of the two
// sides, so the scan can discard almost every chance k-mer hit without touching the reference text.
constexpr uint32_t SIG_FLAG = 0x80000000u;
__device__ __forceinline__ uint32_t side_sig(const uint8_t* __restrict__ T, int len, int pos) {   // 4 bases at pos..pos+3, 7 if outside
    uint32_t s = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int p = pos + t;
        uint32_t c = (p >= 0 && p < len) ? (uint32_t)T[p] : 7u;
        s = (s << 3) | c;
    }
    return s;
}
template <class K>
__global__ void head_flags_kernel(const K* __restrict__ keys, int n, uint32_t* __restrict__ flag, uint32_t* __restrict__ hv) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t f = (s == 0 || keys[s] != keys[s - 1]) ? 1u : 0u;
    flag[s] = f;
    hv[s] = f ? (uint32_t)s : 0u;
}
__global__ void rank_scatter_kernel(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ head, int n, uint32_t* __restrict__ rank) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) rank[sa[s]] = head[s];
}
// u[s] = 1 when s belongs to a group of >= 2 equal keys
__global__ void mark_unsorted_kernel(const uint32_t* __restrict__ flag, int n, uint32_t* __restrict__ u) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    bool single = flag[s] && (s == n - 1 || flag[s + 1]);
    u[s] = single ? 0u : 1u;
}
// compaction: out[pos[s]] = src ? src[s] : s   for u[s] != 0
__global__ void compact_kernel(const uint32_t* __restrict__ u, const uint32_t* __restrict__ pos, int n, const uint32_t* __restrict__ src,
                               uint32_t* __restrict__ out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n && u[s]) out[pos[s]] = src ? src[s] : (uint32_t)s;
}
__global__ void dbl_keys_kernel(const uint32_t* __restrict__ cs, int U, const uint32_t* __restrict__ sa, const uint32_t* __restrict__ rank,
                                int n, int h, int nbits, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= U) return;
    uint32_t i = sa[cs[c]];
    uint64_t head = rank[i];
    // second key: rank of the suffix h further on; suffixes that end before that sort first, the shorter one first
    uint64_t r2 = ((int64_t)i + h < n) ? (uint64_t)rank[i + h] + (uint64_t)n : (uint64_t)(n - 1 - (int64_t)i);
    keys[c] = (head << (nbits + 1)) | r2;
    vals[c] = i;
}
__global__ void dbl_writeback_kernel(const uint32_t* __restrict__ cs, int U, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                     uint32_t* __restrict__ sa, uint32_t* __restrict__ cflag, uint32_t* __restrict__ hv) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= U) return;
    uint32_t s = cs[c];
    sa[s] = vals[c];
    uint32_t f = (c == 0 || keys[c] != keys[c - 1]) ? 1u : 0u;
    cflag[c] = f;
    hv[c] = f ? s : 0u;
}
__global__ void dbl_rank_kernel(const uint32_t* __restrict__ vals, const uint32_t* __restrict__ head, int U, uint32_t* __restrict__ rank) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < U) rank[vals[c]] = head[c];
}

// ---- tied groups of the k-mer sort, fast path: one warp per group of <= 32 equal keys, every lane ranks its suffix against
// the others by direct comparison of the window text (depth-capped).  Larger or deeper groups raise *flag and the host
// re-runs the window through the general prefix-doubling path.
constexpr int TIE_MAX_GROUP = 32;
constexpr int TIE_MAX_DEPTH = 8192;
constexpr int TIE_ITEMS = 8;            // SA slots examined per lane
// suffix order inside the window: when one suffix is a prefix of the other, the shorter (= later start) sorts first
__device__ __forceinline__ bool suffix_less(const uint8_t* __restrict__ R, int n, int a, int b, bool& overflow) {
    const int lim = n - max(a, b);
    const int L = match_len(R + a, R + b, min(lim, TIE_MAX_DEPTH));
    if (L >= lim) return a > b;
    if (L >= TIE_MAX_DEPTH) { overflow = true; return a > b; }
    return R[a + L] < R[b + L];
}
template <class K>
__global__ void __launch_bounds__(256) tie_sort_kernel(const K* __restrict__ keys, int n, const uint8_t* __restrict__ R, uint32_t* __restrict__ sa,
                                                       uint32_t* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t base = warp_global * 32 * TIE_ITEMS;
    for (int r = 0; r < TIE_ITEMS; ++r) {
        const int64_t s = base + r * 32 + lane;
        bool start = false;
        if (s + 1 < n) {
            const K kx = keys[s];
            start = keys[s + 1] == kx && (s == 0 || keys[s - 1] != kx);
        }
        unsigned pend = __ballot_sync(0xffffffffu, start);
        while (pend) {
            const int src = __ffs(pend) - 1;
            pend &= pend - 1;
            const int64_t s0 = base + r * 32 + src;
            uint32_t stop = 0;
            if (lane == 0) stop = *reinterpret_cast<volatile uint32_t*>(flag);
            if (__shfl_sync(0xffffffffu, stop, 0)) return;          // another group already sent the window to the general path
            const K k0 = keys[s0];
            const int64_t t = s0 + lane;
            const unsigned m = __ballot_sync(0xffffffffu, t < n && keys[t] == k0);
            const int g = (m == 0xffffffffu) ? 32 : __ffs(~m) - 1;   // sorted keys: the group is a prefix of the lanes
            if (g == 32 && s0 + 32 < n && keys[s0 + 32] == k0) { if (lane == 0) atomicOr(flag, 1u); continue; }
            const int a = lane < g ? (int)sa[s0 + lane] : 0;
            int rank = 0;
            bool ovf = false;
            for (int j = 0; j < g; ++j) {
                const int b = __shfl_sync(0xffffffffu, a, j);
                if (lane < g && j != lane && suffix_less(R, n, b, a, ovf)) ++rank;
            }
            if (__any_sync(0xffffffffu, ovf)) { if (lane == 0) atomicOr(flag, 1u); continue; }
            if (lane < g) sa[s0 + rank] = (uint32_t)a;
        }
    }
}

// ---- one pass over the final suffix array: adjacent LCPs (from the sorted keys; the window text is read only for equal
// keys), lrp[l] = max(LCP left, LCP right) scattered to text order, and the seed table of the first k bases.
template <class K> struct KeyTraits;
template <> struct KeyTraits<uint32_t> {      // 16 bases, 2 bits each (N-free windows)
    __device__ static int common(uint32_t a, uint32_t b) { const uint32_t x = a ^ b; return x ? (__clz((int)x) >> 1) : KEY2_BASES; }
    __device__ static bool kmer(uint32_t key, int k, uint32_t& code) { code = key >> (2 * (KEY2_BASES - k)); return true; }
};
template <> struct KeyTraits<uint64_t> {      // 21 bases, 3 bits each (code + 1; 0 = past the window end)
    __device__ static int common(uint64_t a, uint64_t b) { const uint64_t x = a ^ b; return x ? ((__clzll((long long)x) - 1) / 3) : KEY_BASES; }
    __device__ static bool kmer(uint64_t key, int k, uint32_t& code) { return kmer_of_key(key, k, code); }
};
template <class K>
__device__ __forceinline__ int pair_lcp(const K* __restrict__ keys, const uint32_t* __restrict__ sa, const uint8_t* __restrict__ R, int n, int64_t s) {
    if (s <= 0 || s >= n) return 0;                    // lcp[0] = lcp[n] = 0
    const K ka = keys[s - 1], kb = keys[s];
    const int a = (int)sa[s - 1], b = (int)sa[s];
    const int lim = n - max(a, b);
    if (ka != kb) return min(KeyTraits<K>::common(ka, kb), lim);      // (key padding past the window end is cut off by lim)
    return match_len(R + a, R + b, lim);
}
template <class K>
__global__ void __launch_bounds__(256) index_finish_kernel(const K* __restrict__ keys, const uint32_t* __restrict__ sa, const uint8_t* __restrict__ R,
                                                           int n, int k, int32_t* __restrict__ lrp, uint2* __restrict__ table) {
    __shared__ int s_lcp[257];
    const int64_t b0 = (int64_t)blockIdx.x * 256;
    const int64_t s = b0 + threadIdx.x;
    s_lcp[threadIdx.x] = pair_lcp(keys, sa, R, n, s);
    if (threadIdx.x == 0) s_lcp[256] = pair_lcp(keys, sa, R, n, b0 + 256);
    __syncthreads();
    if (s >= n) return;
    const int l = (int)sa[s];
    lrp[l] = max(s_lcp[threadIdx.x], s_lcp[threadIdx.x + 1]);
    uint32_t c, cp = 0, cn = 0;
    if (!KeyTraits<K>::kmer(keys[s], k, c)) return;
    const bool same_prev = s > 0 && KeyTraits<K>::kmer(keys[s - 1], k, cp) && cp == c;
    const bool same_next = s + 1 < n && KeyTraits<K>::kmer(keys[s + 1], k, cn) && cn == c;
    if (!same_prev && !same_next) {
        table[c] = make_uint2((uint32_t)l, SIG_FLAG | (side_sig(R, n, l - 4) << 12) | side_sig(R, n, l + k));
    } else {
        if (!same_prev) table[c].x = (uint32_t)s;
        if (!same_next) table[c].y = (uint32_t)s + 1u;
    }
}

// ------------------------------------------------------------------ MEM scan (seed and extend)
// compare the k bases at q with the suffix of R starting at l (window end sorts first): <0, 0, >0
__device__ __forceinline__ int cmp_kmer(const uint8_t* __restrict__ R, int n, int l, const uint8_t* __restrict__ q, int k) {
    for (int t = 0; t < k; ++t) {
        int a = (l + t < n) ? (int)R[l + t] + 1 : 0;
        int b = (int)q[t] + 1;
        if (a != b) return a - b;
    }
    return 0;
}

// 4 bases packed like side_sig from bytes 4..7 of a little-endian 8-byte word (byte 4 = first base)
__device__ __forceinline__ uint32_t sig_hi4(uint64_t w) {
    return (uint32_t)((w >> 32) & 7u) << 9 | (uint32_t)((w >> 40) & 7u) << 6 | (uint32_t)((w >> 48) & 7u) << 3 | (uint32_t)((w >> 56) & 7u);
}

// one 256-byte step of a warp-cooperative comparison: equal leading bytes (0..8) of this lane's 8 bytes; 0 past the limit
__device__ __forceinline__ int coop_step(const uint8_t* __restrict__ Q, const uint8_t* __restrict__ R, int qp, int rp, int lm, int base, int lane) {
    const int off = base + lane * 8;
    if (off >= lm) return 0;
    const uint64_t x = load8u(Q + qp + off) ^ load8u(R + rp + off);
    return x ? ((__ffsll((long long)x) - 1) >> 3) : 8;
}
// stop = ballot(my < 8): if some lane saw the end of the match, e = its length (capped by the limit)
__device__ __forceinline__ bool coop_resolve(unsigned stop, int my, int lm, int base, int& e) {
    if (!stop) return false;
    const int fl = __ffs(stop) - 1;
    e = min(lm, base + fl * 8 + __shfl_sync(0xffffffffu, my, fl));
    return true;
}

// One thread per sampled query position (every step-th base): k-mer -> seed table -> flank-signature filter -> left extension
// (de-duplication: a match of >= minsize bases contains exactly one sampled seed whose left extension is shorter than the
// sample spacing).  The right extensions of a warp's surviving seeds are then done by the WHOLE warp, one seed after the other,
// 32 x 8 bytes per step with a ballot for the first mismatch: the extension of a 100-base match is one step for everybody
// instead of four divergent 32-byte steps for one lane.  Events are appended with one atomic per warp and round.
// (Measured and dropped: extending two seeds per iteration - registers cost more occupancy than the overlap gained, +15 %;
//  dropping seeds that continue the previous lane's diagonal from the flank signature alone - the saved left check is an L1
//  hit, the shuffle is not free, +4 %; loading lrp[] ahead of the extension - chance hits then pay a DRAM sector, +8 %.)
__global__ void __launch_bounds__(128, 16) seed_extend_kernel(const uint8_t* __restrict__ R, int n, const uint32_t* __restrict__ sa,
                                                          const int32_t* __restrict__ lrp, const uint2* __restrict__ table,
                                                          int k, int step, int minsize,
                                                          const StrandDesc* __restrict__ strands, uint64_t* __restrict__ ev_key,
                                                          uint64_t* __restrict__ ev_val, unsigned long long* __restrict__ ev_count,
                                                          unsigned long long ev_cap) {
    const unsigned FULL = 0xffffffffu;
    const int strand = blockIdx.y;
    const uint8_t* __restrict__ Q = strands[strand].q;
    const int m = strands[strand].m;
    const int lane = threadIdx.x & 31;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((idx - lane) * step + k > m) return;            // the whole warp lies past the end of the strand
    const long long jl = idx * step;
    const bool valid = jl + k <= m;
    const int j = valid ? (int)jl : 0;
    // ---- the lane's seed: bucket [lo, hi) of the suffix array, or one text position (`single`)
    int lo = 0, hi = 0;
    int single = -1;
    if (valid) {
        // k <= 12 bases = the low 2 bits of 12 consecutive bytes: two 8-byte loads; four bases are gathered by one multiply
        const uint64_t w0 = load8u(Q + j), w1 = load8u(Q + j + 8);
        const uint64_t m0 = k >= 8 ? ~0ull : ((1ull << (8 * k)) - 1ull);
        const uint64_t m1 = k > 8 ? ((1ull << (8 * (k - 8))) - 1ull) : 0ull;
        const bool plain = (((w0 & m0) | (w1 & m1)) & 0xfcfcfcfcfcfcfcfcull) == 0ull;      // no N (code 4) / padding among the k bases
        const uint32_t code = ((pack4x2((uint32_t)w0) << 16) | (pack4x2((uint32_t)(w0 >> 32)) << 8) | pack4x2((uint32_t)w1)) >> (2 * (MAX_SEED_K - k));
        if (plain) {
            const uint2 e = table[code];
            if (e.y & SIG_FLAG) {
                single = (int)e.x; hi = 1;
                if (minsize >= k + 7) {
                    // one of the two 4-base flanks must agree (see index_finish_kernel); flanks outside the query never agree
                    uint32_t ql, qr;
                    if (k == MAX_SEED_K && j >= 8 && j + 16 <= m) { ql = sig_hi4(load8u(Q + j - 8)); qr = sig_hi4(w1); }
                    else { ql = side_sig(Q, m, j - 4); qr = side_sig(Q, m, j + k); }
                    const uint32_t rl = (e.y >> 12) & 0xfffu, rr = e.y & 0xfffu;
                    if (ql != rl && qr != rr) hi = 0;
                }
            } else { lo = (int)e.x; hi = (int)e.y; }
        } else {
            // k-mer with N: binary search the suffix array (rare)
            int a = 0, b = n;
            while (a < b) { int mid = (a + b) >> 1; if (cmp_kmer(R, n, (int)sa[mid], Q + j, k) < 0) a = mid + 1; else b = mid; }
            lo = a;
            b = n;
            while (a < b) { int mid = (a + b) >> 1; if (cmp_kmer(R, n, (int)sa[mid], Q + j, k) <= 0) a = mid + 1; else b = mid; }
            hi = a;
        }
    }
    // ---- rounds: every lane takes the next suffix of its bucket; a round ends with the warp's cooperative right extensions
    int sidx = lo;
    while (__any_sync(FULL, sidx < hi)) {
        bool pend = false;
        int c = 0, l = 0, lim = 0;
        if (sidx < hi) {
            l = single >= 0 ? single : (int)sa[sidx];
            ++sidx;
            if (l + k <= n) {                             // (2-bit keys: suffixes shorter than k sit in the bucket of their padded key)
                const int room = min(j, l);               // bases to the left of the seed inside both strings
                const int cmax = min(step, room);
                // left extension, 8 bytes per compare (the byte just left of the seed is the top byte of the word); a compare may
                // look further left than cmax as long as it stays inside the strings - the result is capped
                while (c < cmax) {
                    if (c + 8 <= room) {
                        const uint64_t x = load8u(Q + j - c - 8) ^ load8u(R + l - c - 8);
                        if (x) { c += __clzll((long long)x) >> 3; break; }
                        c += 8;
                    } else {
                        while (c < cmax && Q[j - 1 - c] == R[l - 1 - c]) ++c;
                        break;
                    }
                }
                c = min(c, cmax);
                if (c < step) {                           // else: an earlier sampled seed lies inside the same match
                    lim = min(m - (j + k), n - (l + k));
                    // An event needs c + k + ext >= minsize.  Suffixes of a seed bucket that are not the match (every bucket of
                    // more than one suffix has them, and they carry no flank signature) end within a base or two: 8 bytes to the
                    // right of the seed settle that here, instead of a whole cooperative step of the warp per such pair (they
                    // were most of the pending seeds: ~40 % of the kernel's instructions went into their extensions)
                    const int need = minsize - k - c;     // bases still missing on the right
                    pend = true;
                    if (need > 0) {
                        if (lim < need) pend = false;
                        else {
                            const uint64_t x = load8u(Q + j + k) ^ load8u(R + l + k);
                            const int e8 = x ? ((__ffsll((long long)x) - 1) >> 3) : 8;
                            if (e8 < 8 && min(e8, lim) < need) pend = false;
                        }
                    }
                }
            }
        }
        // cooperative right extension of the pending seeds, two at a time: each half-warp compares 16 x 8 bytes of one seed per
        // step (most matches end within 128 bases: one step)
        int ext = 0;
        unsigned pm = __ballot_sync(FULL, pend);
        const int half = lane >> 4, hl = lane & 15;
        while (pm) {
            const int sA = __ffs(pm) - 1;
            pm &= pm - 1;
            int sB = -1;
            if (pm) { sB = __ffs(pm) - 1; pm &= pm - 1; }
            const int src = half ? sB : sA;                   // this half's seed (-1: none)
            const int ssrc = src < 0 ? 0 : src;
            const int qp = __shfl_sync(FULL, j + k, ssrc), rp = __shfl_sync(FULL, l + k, ssrc);
            const int lm_src = __shfl_sync(FULL, lim, ssrc);       // (every lane takes part in the shuffle)
            const int lm = src < 0 ? 0 : lm_src;
            int e = lm;
            bool done = lm == 0;
            for (int base = 0;; base += 128) {
                int my = 8;
                if (!done) {
                    const int off = base + hl * 8;
                    my = 0;                                     // past the limit: the match stops here
                    if (off < lm) {
                        const uint64_t x = load8u(Q + qp + off) ^ load8u(R + rp + off);
                        my = x ? ((__ffsll((long long)x) - 1) >> 3) : 8;
                    }
                }
                const unsigned hb = (__ballot_sync(FULL, !done && my < 8) >> (16 * half)) & 0xffffu;
                const int fl = hb ? __ffs(hb) - 1 : 0;
                const int mlen = __shfl_sync(FULL, my, 16 * half + fl);
                if (!done) {
                    if (hb) { e = min(lm, base + fl * 8 + mlen); done = true; }
                    else if (base + 128 >= lm) { e = lm; done = true; }
                }
                if (!__any_sync(FULL, !done)) break;
            }
            const int eA = __shfl_sync(FULL, e, 0), eB = __shfl_sync(FULL, e, 16);
            if (lane == sA) ext = eA;
            if (lane == sB) ext = eB;
        }
        // events of this round: one atomic per warp
        bool emit = false;
        int L = 0, l0 = 0;
        if (pend) {
            L = c + k + ext;
            l0 = l - c;
            emit = L >= minsize && L > lrp[l0];
        }
        const unsigned em = __ballot_sync(FULL, emit);
        if (em) {
            unsigned long long basev = 0;
            if (lane == __ffs(em) - 1) basev = atomicAdd(ev_count, (unsigned long long)__popc(em));
            basev = __shfl_sync(FULL, basev, __ffs(em) - 1);
            if (emit) {
                const unsigned long long slot = basev + (unsigned long long)__popc(em & ((1u << lane) - 1u));
                if (slot < ev_cap) {
                    ev_key[slot] = ((uint64_t)(uint32_t)strand << 32) | (uint32_t)l0;
                    ev_val[slot] = ((uint64_t)(uint32_t)(l0 + L) << 32) | (uint32_t)(j - c);
                }
            }
        }
    }
}

// segment bounds of each strand inside the sorted event array; also splits off the l column
__global__ void strand_segments_kernel(const uint64_t* __restrict__ ev_key, int E, uint32_t* __restrict__ seg_lo, uint32_t* __restrict__ seg_hi,
                                       uint32_t* __restrict__ evl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    uint64_t kx = ev_key[i];
    uint32_t st = (uint32_t)(kx >> 32);
    evl[i] = (uint32_t)kx;
    if (i == 0 || (uint32_t)(ev_key[i - 1] >> 32) != st) seg_lo[st] = (uint32_t)i;
    if (i == E - 1 || (uint32_t)(ev_key[i + 1] >> 32) != st) seg_hi[st] = (uint32_t)i + 1u;
}

// bounds[strand][t] = first event of the strand with l >= t * FOLD_TILE (t = 0..ntiles): the fold and the candidate pass look
// events up inside one tile instead of searching the whole strand
__global__ void tile_bounds_init_kernel(const uint32_t* __restrict__ seg_hi, int ns, int ntiles1, uint32_t* __restrict__ bounds) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)ns * ntiles1) return;
    bounds[i] = seg_hi[i / ntiles1];                 // (strands without events: seg_lo = seg_hi = 0)
}
__global__ void tile_bounds_fill_kernel(const uint64_t* __restrict__ ev_key, int E, int tile_shift, int ntiles1, uint32_t* __restrict__ bounds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    const uint64_t kx = ev_key[i];
    const uint32_t st = (uint32_t)(kx >> 32);
    const int t1 = (int)((uint32_t)kx >> tile_shift);          // tile of this event
    int t0 = 0;                                               // first tile whose lower bound this event is
    if (i > 0) {
        const uint64_t kp = ev_key[i - 1];
        if ((uint32_t)(kp >> 32) == st) {
            t0 = (int)((uint32_t)kp >> tile_shift) + 1;
        }
    }
    for (int t = t0; t <= t1; ++t) bounds[(size_t)st * ntiles1 + t] = (uint32_t)i;
}

// scan state: fl = max floor u[l], (t1, t2) = two largest ends, d1 = (query start - ref start) of the event holding t1
struct St { int fl, t1, t2, d1; };
__device__ __forceinline__ St st_combine(const St& a, const St& b) {     // a = earlier (smaller l)
    St r;
    r.fl = max(a.fl, b.fl);
    if (b.t1 >= a.t1) { r.t1 = b.t1; r.d1 = b.d1; r.t2 = max(a.t1, b.t2); }
    else { r.t1 = a.t1; r.d1 = a.d1; r.t2 = max(a.t2, b.t1); }
    return r;
}
__device__ __forceinline__ St st_shfl_up(const St& s, int o) {
    St r;
    r.fl = __shfl_up_sync(0xffffffffu, s.fl, o);
    r.t1 = __shfl_up_sync(0xffffffffu, s.t1, o);
    r.t2 = __shfl_up_sync(0xffffffffu, s.t2, o);
    r.d1 = __shfl_up_sync(0xffffffffu, s.d1, o);
    return r;
}
// one block per strand: inclusive scan of the strand's events (sorted by l); 8 consecutive events per thread
constexpr int ES_ITEMS = 8;
__global__ void __launch_bounds__(256) event_scan_kernel(const uint32_t* __restrict__ evl, const uint64_t* __restrict__ ev_val,
                                                         const int32_t* __restrict__ lrp, const uint32_t* __restrict__ seg_lo,
                                                         const uint32_t* __restrict__ seg_hi, int4* __restrict__ states) {
    __shared__ St s_w[8];
    const int strand = blockIdx.x;
    const int lo = (int)seg_lo[strand], hi = (int)seg_hi[strand];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    St carry = {0, 0, 0, 0};
    for (int base = lo; base < hi; base += 256 * ES_ITEMS) {
        const int i0 = base + threadIdx.x * ES_ITEMS;
        St v[ES_ITEMS];
        St x = {0, 0, 0, 0};
#pragma unroll
        for (int r = 0; r < ES_ITEMS; ++r) {
            const int i = i0 + r;
            St e = {0, 0, 0, 0};
            if (i < hi) {
                int l = (int)evl[i];
                uint64_t val = ev_val[i];
                e.fl = l + lrp[l];
                e.t1 = (int)(uint32_t)(val >> 32);
                e.t2 = 0;
                e.d1 = (int)(uint32_t)val - l;
            }
            x = st_combine(x, e);
            v[r] = x;                                   // thread-local inclusive prefix
        }
        St tot_thread = x;
        for (int o = 1; o < 32; o <<= 1) { St y = st_shfl_up(x, o); if (lane >= o) x = st_combine(y, x); }
        if (lane == 31) s_w[w] = x;
        __syncthreads();
        St pre = carry;
        for (int q = 0; q < w; ++q) pre = st_combine(pre, s_w[q]);
        St tot = carry;
        for (int q = 0; q < 8; ++q) tot = st_combine(tot, s_w[q]);
        // exclusive prefix of this thread = pre (+) inclusive warp scan of the previous lane
        St prev = st_shfl_up(x, 1);
        if (lane > 0) pre = st_combine(pre, prev);
        (void)tot_thread;
#pragma unroll
        for (int r = 0; r < ES_ITEMS; ++r) {
            const int i = i0 + r;
            if (i < hi) { St o = st_combine(pre, v[r]); states[i] = make_int4(o.fl, o.t1, o.t2, o.d1); }
        }
        carry = tot;
        __syncthreads();
    }
}

// last event index in [a,b) with l <= k, or a-1
__device__ __forceinline__ int last_le(const uint32_t* __restrict__ evl, int a, int b, uint32_t k) {
    int lo = a, hi = b;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (evl[mid] <= k) lo = mid + 1; else hi = mid; }
    return lo - 1;
}
__device__ __forceinline__ void strand_at(const uint32_t* __restrict__ evl, const int4* __restrict__ states, int seg_lo, int a, int b,
                                          uint32_t k, int& UP, int& EP, int& d1) {
    int t = last_le(evl, a, b, k);
    if (t < seg_lo) { UP = 0; EP = 0; d1 = 0; return; }
    int4 s = states[t];
    UP = max(s.x, s.z);
    EP = max(s.y, s.x);
    d1 = s.w;
}

constexpr int FOLD_THREADS = 256;
constexpr int FOLD_ITEMS = 4;
constexpr int FOLD_TILE = FOLD_THREADS * FOLD_ITEMS;
constexpr int FOLD_TILE_SHIFT = 10;
static_assert(FOLD_TILE == (1 << FOLD_TILE_SHIFT), "tile size");
constexpr int FOLD_QB = 4;              // queries per round (their tile events are staged in shared memory together)
constexpr int FOLD_EV = 96;             // staged events per strand and tile; denser tiles are searched in global memory
// Fold over `nq` queries starting at strand `s0` (ini order) for a tile of reference positions.  The scan state of a strand
// at position k is the state of its last event with l <= k: the events of the tile (bounds[]) plus the state carried in
// from the left are staged in shared memory, the lookup is a short walk there.
__global__ void __launch_bounds__(FOLD_THREADS) fold_kernel(const uint32_t* __restrict__ evl, const int4* __restrict__ states,
                                                            const uint32_t* __restrict__ seg_lo, const uint32_t* __restrict__ bounds,
                                                            int ntiles1, int s0, int nq, int n, int32_t* __restrict__ MUP,
                                                            int32_t* __restrict__ MEP, int init_master) {
    __shared__ uint32_t s_l[2 * FOLD_QB][FOLD_EV];
    __shared__ int4 s_st[2 * FOLD_QB][FOLD_EV + 1];         // [0] = state carried in from the left of the tile
    __shared__ int s_a[2 * FOLD_QB], s_b[2 * FOLD_QB], s_lo[2 * FOLD_QB];
    const int tile = blockIdx.x;
    const int tile0 = tile * FOLD_TILE;
    int up[FOLD_ITEMS], ep[FOLD_ITEMS];
#pragma unroll
    for (int r = 0; r < FOLD_ITEMS; ++r) {
        int k = tile0 + r * FOLD_THREADS + threadIdx.x;
        if (init_master || k >= n) { up[r] = 0; ep[r] = n; }
        else { up[r] = MUP[k]; ep[r] = MEP[k]; }
    }
    for (int qb = 0; qb < nq; qb += FOLD_QB) {
        const int ns = 2 * min(FOLD_QB, nq - qb);          // strands of this round
        if (threadIdx.x < ns) {
            const int st = s0 + 2 * qb + threadIdx.x;
            s_a[threadIdx.x] = (int)bounds[(size_t)st * ntiles1 + tile];
            s_b[threadIdx.x] = (int)bounds[(size_t)st * ntiles1 + tile + 1];
            s_lo[threadIdx.x] = (int)seg_lo[st];
        }
        __syncthreads();
        for (int x = threadIdx.x; x < ns * (FOLD_EV + 1); x += FOLD_THREADS) {
            const int s = x / (FOLD_EV + 1), i = x - s * (FOLD_EV + 1);
            const int a = s_a[s], cnt = s_b[s] - a;
            if (cnt > FOLD_EV) continue;
            if (i == 0) s_st[s][0] = (a - 1 >= s_lo[s]) ? states[a - 1] : make_int4(0, 0, 0, 0);
            else if (i <= cnt) { s_l[s][i - 1] = evl[a + i - 1]; s_st[s][i] = states[a + i - 1]; }
        }
        __syncthreads();
        for (int q = 0; q < ns / 2; ++q) {
#pragma unroll
            for (int r = 0; r < FOLD_ITEMS; ++r) {
                const int k = tile0 + r * FOLD_THREADS + threadIdx.x;
                if (k >= n) continue;
                int UP2[2], EP2[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int s = 2 * q + h;
                    const int a = s_a[s], cnt = s_b[s] - a;
                    if (cnt <= FOLD_EV) {
                        int p = 0;
                        while (p < cnt && s_l[s][p] <= (uint32_t)k) ++p;
                        const int4 v = s_st[s][p];
                        UP2[h] = max(v.x, v.z); EP2[h] = max(v.y, v.x);
                    } else {
                        int d;
                        strand_at(evl, states, s_lo[s], a, s_b[s], (uint32_t)k, UP2[h], EP2[h], d);
                    }
                }
                const int fe = min(ep[r], EP2[0]), ce = min(ep[r], EP2[1]);
                if (fe > ce) { up[r] = max(up[r], UP2[0]); ep[r] = fe; }
                else { up[r] = max(up[r], UP2[1]); ep[r] = ce; }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < FOLD_ITEMS; ++r) {
        int k = tile0 + r * FOLD_THREADS + threadIdx.x;
        if (k < n) { MUP[k] = up[r]; MEP[k] = ep[r]; }
    }
}

__global__ void emit_flags_kernel(const int32_t* __restrict__ MUP, const int32_t* __restrict__ MEP, int n, int minsize,
                                  uint32_t* __restrict__ flag) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int prev = k ? MEP[k - 1] : 0;
    int ep = MEP[k];
    flag[k] = (ep > prev && MUP[k] < ep && ep - k >= minsize) ? 1u : 0u;
}

// ---- MUMi mode (Aligner::setMumi, src/parsnp.cpp:1869-2115): per query, reference positions covered by MUMs >= 15 bp
// a[k] = EP[k] where the position is "good" (UP < EP and EP - k < n), else 0
__global__ void mumi_good_kernel(const int32_t* __restrict__ MUP, const int32_t* __restrict__ MEP, int n, uint32_t* __restrict__ a) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int up = MUP[k], ep = MEP[k];
    a[k] = (up < ep && ep - k < n) ? (uint32_t)ep : 0u;
}
// v[k] = EP[k] for emitted candidates of length >= 15 (emitted = good and EP greater than at the previous good position)
__global__ void mumi_val_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ prevmax, int n, uint32_t* __restrict__ v) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t ep = a[k];
    v[k] = (ep && ep > prevmax[k] && (int)ep - k >= 15) ? ep : 0u;
}
__global__ void mumi_cover_kernel(const uint32_t* __restrict__ runmax, int n, uint32_t* __restrict__ c) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) c[k] = runmax[k] > (uint32_t)k ? 1u : 0u;
}

// per (candidate, local query): the two strands' (EP, diagonal) at the candidate position
__global__ void pass2a_kernel(const uint32_t* __restrict__ ck, int ncand, const uint32_t* __restrict__ evl, const int4* __restrict__ states,
                              const uint32_t* __restrict__ seg_lo, const uint32_t* __restrict__ bounds, int ntiles1, int nq,
                              int4* __restrict__ tmp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)ncand * nq) return;
    const int c = (int)(idx / nq), q = (int)(idx % nq);
    const uint32_t k = ck[c];
    const int tile = (int)(k >> FOLD_TILE_SHIFT);
    int UPf, EPf, df, UPc, EPc, dc;
    const uint32_t* bf = bounds + (size_t)(2 * q) * ntiles1 + tile;
    strand_at(evl, states, (int)seg_lo[2 * q], (int)bf[0], (int)bf[1], k, UPf, EPf, df);
    const uint32_t* bc = bounds + (size_t)(2 * q + 1) * ntiles1 + tile;
    strand_at(evl, states, (int)seg_lo[2 * q + 1], (int)bc[0], (int)bc[1], k, UPc, EPc, dc);
    tmp[idx] = make_int4(EPf, EPc, df, dc);
}
// per candidate: replay the fold over the local queries (ini order) to recover every query's strand and start (SP)
__global__ void pass2b_kernel(const uint32_t* __restrict__ ck, int ncand, const int4* __restrict__ tmp, int nq, int n,
                              const int32_t* __restrict__ MEP, const int32_t* __restrict__ initEP, int32_t* __restrict__ out_lon,
                              int32_t* __restrict__ out_sp, uint8_t* __restrict__ out_fwd) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand) return;
    const uint32_t k = ck[c];
    int M = initEP ? initEP[k] : n;        // Master.EP[k] before this rank's first query (multi-GPU: prefix over earlier ranks)
    for (int q = 0; q < nq; ++q) {
        const int4 t = tmp[(size_t)c * nq + q];
        int fe = min(M, t.x), ce = min(M, t.y);
        if (fe > ce) { out_sp[(size_t)c * nq + q] = (int)k + t.z; out_fwd[(size_t)c * nq + q] = 1; M = fe; }
        else { out_sp[(size_t)c * nq + q] = (int)k + t.w; out_fwd[(size_t)c * nq + q] = 0; M = ce; }
    }
    out_lon[c] = MEP[k] - (int)k;
}

// ------------------------------------------------------------------ host driver
struct WindowIndexInfo { int rounds = 0; int64_t unsorted_after_sort = 0; bool two_bit = false; bool fallback = false; };

class BigPath {
public:
    GpuTimers* tm = nullptr;
    WindowIndexInfo last_index;

    // R: device text of the window (codes 0..4), n bases. strands: host array of 2*nq descriptors (device text pointers):
    // strand 2q = forward string of query q's region, 2q+1 = its reverse complement.
    void search(const uint8_t* R, int n, int nq, const std::vector<StrandDesc>& strands, int minsize, cudaStream_t st,
                std::vector<int32_t>& out_k, std::vector<int32_t>& out_lon, std::vector<int32_t>& out_sp, std::vector<uint8_t>& out_fwd,
                bool two_bit = false) {
        build_index(R, n, minsize, st, two_bit);
        scan(R, n, nq, strands, minsize, st, out_k, out_lon, out_sp, out_fwd);
    }

    // debug/test access (device pointers valid until the next build_index)
    const uint32_t* d_sa() const { return sa_ptr_; }
    const int32_t* d_lrp() const { return lrp_.get(); }

    // Index of one window.  Fast path: k-mer radix sort, tied groups ranked by direct text comparison (tie_sort_kernel), one
    // finishing pass (LCP -> lrp, seed table).  Windows with large or deep repeats raise a device flag; it is read at the next
    // natural synchronisation point (ensure_index) and the window is then re-indexed through the general prefix-doubling path.
    void build_index(const uint8_t* R, int n, int minsize, cudaStream_t st, bool two_bit = false) {
        idx_R_ = R; idx_n_ = n; idx_minsize_ = minsize; idx_two_bit_ = two_bit;
        const bool force_doubling = getenv("PB200_FORCE_DOUBLING") != nullptr;      // tests: always take the general path
        build_index_impl(st, force_doubling);
        if (!force_doubling) {
            uint32_t* h = tie_host_.ensure(1);
            PB_CUDA(cudaMemcpyAsync(h, tieflag_.get(), 4, cudaMemcpyDeviceToHost, st));
            idx_pending_ = true;
        }
    }
    // true if the window had to be re-indexed (anything computed from the index since build_index must be redone)
    bool ensure_index(cudaStream_t st) {
        if (!idx_pending_) return false;
        PB_CUDA(cudaStreamSynchronize(st));
        idx_pending_ = false;
        if (*tie_host_.get() == 0) return false;
        build_index_impl(st, true);
        last_index.fallback = true;
        return true;
    }


private:
    template <class K>
    void sort_and_finish(K* k0, K* k1, int end_bit, int64_t h0, cudaStream_t st, bool doubling, const uint8_t* key_text) {
        const int TB = 256;
        const int n = idx_n_;
        const uint8_t* R = idx_R_;
        const unsigned nb = (unsigned)((n + TB - 1) / TB);
        uint32_t* v0 = vals0_.get();
        uint32_t* v1 = vals1_.get();
        if (tm) tm->start(GpuTimers::T_INDEX_SORT, st);
        const int res = sorter_.sort<K, uint32_t>(k0, k1, v0, v1, n, 0, end_bit, st, key_text);
        const K* ks = res ? k1 : k0;
        uint32_t* sa = res ? v1 : v0;                          // the sorted values ARE the suffix array (refined in place below)
        sa_ptr_ = sa;
        if (tm) tm->stop(GpuTimers::T_INDEX_SORT, st);
        if (tm) tm->start(GpuTimers::T_INDEX_DOUBLING, st);
        uint32_t* flag = tieflag_.ensure(4, false, st);
        PB_CUDA(cudaMemsetAsync(flag, 0, 4, st));
        if (!doubling) {
            const int64_t per_block = (int64_t)TB * TIE_ITEMS;
            pb200::launch(tie_sort_kernel<K>, (unsigned)((n + per_block - 1) / per_block), TB, 0, st, ks, n, R, sa, flag);
        } else {
            refine_by_doubling(ks, sa, h0, st);
        }
        if (tm) tm->stop(GpuTimers::T_INDEX_DOUBLING, st);
        if (tm) tm->start(GpuTimers::T_INDEX_LCP, st);
        const size_t tsize = (size_t)1 << (2 * seed_k_);
        uint2* table = table_.ensure(tsize, false, st);
        PB_CUDA(cudaMemsetAsync(table, 0, tsize * sizeof(uint2), st));
        pb200::launch(index_finish_kernel<K>, nb, TB, 0, st, ks, sa, R, n, seed_k_, lrp_.get(), table);
        if (tm) tm->stop(GpuTimers::T_INDEX_LCP, st);
    }

    // general path: prefix doubling (Larsson-Sadakane with discarding) on the groups the k-mer sort left tied
    template <class K>
    void refine_by_doubling(const K* ks, uint32_t* sa, int64_t h0, cudaStream_t st) {
        const int TB = 256;
        const int n = idx_n_;
        const unsigned nb = (unsigned)((n + TB - 1) / TB);
        uint64_t* k0 = keys0_.get();
        uint64_t* k1 = keys1_.get();
        // the doubling rounds sort (key, value) pairs of their own: they must not clobber the sorted k-mer keys `ks` (needed by
        // the finishing pass) nor `sa`
        uint64_t* dk0 = dkeys0_.ensure((size_t)n, false, st);
        uint64_t* dk1 = dkeys1_.ensure((size_t)n, false, st);
        uint32_t* dv0 = dvals0_.ensure((size_t)n, false, st);
        uint32_t* dv1 = dvals1_.ensure((size_t)n, false, st);
        (void)k0; (void)k1;
        uint32_t* rank = rank_.ensure((size_t)n, false, st);
        uint32_t* tA = tmpA_.ensure((size_t)n + 1, false, st);
        uint32_t* tB = tmpB_.ensure((size_t)n + 1, false, st);
        uint32_t* tC = tmpC_.ensure((size_t)n + 1, false, st);
        uint32_t* d_tot = total_.ensure(4, false, st);
        pb200::launch(head_flags_kernel<K>, nb, TB, 0, st, ks, n, tA /*flag*/, tB /*hv*/);
        scanner_.scan<prim::OpMax, false>(tB, tB, n, nullptr, st);                 // tB = head index per SA slot
        pb200::launch(rank_scatter_kernel, nb, TB, 0, st, sa, tB, n, rank);
        pb200::launch(mark_unsorted_kernel, nb, TB, 0, st, tA, n, tC /*u*/);
        scanner_.scan<prim::OpSum, true>(tC, tB, n, d_tot, st);                    // tB = compact position
        uint32_t U = 0;
        PB_CUDA(cudaMemcpyAsync(&U, d_tot, 4, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        last_index.unsorted_after_sort = U;
        last_index.rounds = 0;
        if (U == 0) return;
        uint32_t* cs = cs0_.ensure((size_t)U, false, st);
        uint32_t* cs2 = cs1_.ensure((size_t)U, false, st);
        pb200::launch(compact_kernel, nb, TB, 0, st, tC, tB, n, nullptr, cs);
        int nbits = 1;
        while (((int64_t)1 << nbits) < (int64_t)n + 1) ++nbits;
        int64_t h = h0;
        while (U > 0) {
            const unsigned ub = (unsigned)((U + TB - 1) / TB);
            pb200::launch(dbl_keys_kernel, ub, TB, 0, st, cs, (int)U, sa, rank, n, (int)std::min<int64_t>(h, n), nbits, dk0, dv0);
            int r2 = sorter_.sort<uint64_t, uint32_t>(dk0, dk1, dv0, dv1, U, 0, 2 * nbits + 1, st);
            const uint64_t* k2 = r2 ? dk1 : dk0;
            const uint32_t* v2 = r2 ? dv1 : dv0;
            pb200::launch(dbl_writeback_kernel, ub, TB, 0, st, cs, (int)U, k2, v2, sa, tA /*cflag*/, tB /*hv*/);
            scanner_.scan<prim::OpMax, false>(tB, tB, U, nullptr, st);         // head per compact slot
            pb200::launch(dbl_rank_kernel, ub, TB, 0, st, v2, tB, (int)U, rank);
            pb200::launch(mark_unsorted_kernel, ub, TB, 0, st, tA, (int)U, tC);
            scanner_.scan<prim::OpSum, true>(tC, tB, U, d_tot, st);
            uint32_t U2 = 0;
            PB_CUDA(cudaMemcpyAsync(&U2, d_tot, 4, cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaStreamSynchronize(st));
            if (U2 > 0) pb200::launch(compact_kernel, ub, TB, 0, st, tC, tB, (int)U, cs, cs2);
            std::swap(cs, cs2);
            U = U2;
            h *= 2;
            last_index.rounds++;
            if (last_index.rounds > 40) throw CudaError("prefix doubling did not converge");
        }
    }

    void build_index_impl(cudaStream_t st, bool doubling) {
        const int TB = 256;
        const int n = idx_n_;
        const uint8_t* R = idx_R_;
        const unsigned nb = (unsigned)((n + TB - 1) / TB);
        uint64_t* k0 = keys0_.ensure((size_t)n, false, st);
        uint64_t* k1 = keys1_.ensure((size_t)n, false, st);
        uint32_t* v0 = vals0_.ensure((size_t)n, false, st);
        vals1_.ensure((size_t)n, false, st);
        lrp_.ensure((size_t)n, false, st);
        seed_k_ = std::min(MAX_SEED_K, std::max(1, idx_minsize_));
        last_index.two_bit = idx_two_bit_;
        last_index.fallback = false;
        last_index.rounds = 0;
        last_index.unsorted_after_sort = -1;
        if (idx_two_bit_) {
            // N-free window: 32-bit keys of 16 bases, 4 radix passes over (4 B key + 4 B value); the first pass packs the keys
            // straight from the window text (no key-generation kernel)
            uint32_t* q0 = reinterpret_cast<uint32_t*>(k0);
            uint32_t* q1 = reinterpret_cast<uint32_t*>(k1);
            if (n >= 2) sort_and_finish<uint32_t>(q0, q1, 2 * KEY2_BASES, KEY2_BASES, st, doubling, R);
            else {
                pb200::launch(make_keys2_kernel, nb, TB, 0, st, R, n, q0, v0);
                sort_and_finish<uint32_t>(q0, q1, 2 * KEY2_BASES, KEY2_BASES, st, doubling, nullptr);
            }
        } else {
            if (tm) tm->start(GpuTimers::T_INDEX_KEYS, st);
            pb200::launch(make_keys_kernel, nb, TB, 0, st, R, n, k0, v0);
            if (tm) tm->stop(GpuTimers::T_INDEX_KEYS, st);
            sort_and_finish<uint64_t>(k0, k1, 3 * KEY_BASES, KEY_BASES, st, doubling, nullptr);
        }
        PB_CUDA(cudaGetLastError());
    }
public:

    // ---- staged scan (the single-GPU path runs the stages back to back; the sharded path exchanges between them) ----
    // stage 1: MEM events of the given strands (2 per local query), ordered by (strand, l) and scanned
    void scan_events(const uint8_t* R, int n, int nq, const std::vector<StrandDesc>& strands, int minsize, cudaStream_t st) {
        const int TB = 256;
        const int ns = 2 * nq;
        const int k = seed_k_;
        const int step = std::max(1, minsize - k + 1);
        cur_n_ = n; cur_nq_ = nq; cur_minsize_ = minsize;
        StrandDesc* d_str = strands_.ensure((size_t)std::max(ns, 1), false, st);
        if (ns) PB_CUDA(cudaMemcpyAsync(d_str, strands.data(), sizeof(StrandDesc) * ns, cudaMemcpyHostToDevice, st));
        int64_t tot_m = 0;
        int max_m = 0;
        for (int s = 0; s < ns; ++s) { tot_m += strands[s].m; max_m = std::max(max_m, strands[s].m); }
        size_t cap = std::max<size_t>(ev_cap_hint_, (size_t)(tot_m / 16) + 65536);
        unsigned long long* d_cnt = evcount_.ensure(1, false, st);
        unsigned long long E = 0;
        if (tm) tm->start(GpuTimers::T_SCAN_SEED, st);
        for (;;) {
            uint64_t* ek = evk0_.ensure(cap, false, st);
            uint64_t* ev = evv0_.ensure(cap, false, st);
            PB_CUDA(cudaMemsetAsync(d_cnt, 0, 8, st));
            long long samples = ((long long)max_m + step - 1) / step;
            dim3 grid((unsigned)((samples + 127) / 128), (unsigned)std::max(ns, 1));
            if (samples > 0 && ns > 0)
                pb200::launch(seed_extend_kernel, grid, 128, 0, st, R, n, sa_ptr_, lrp_.get(), table_.get(), k, step, minsize, d_str, ek, ev,
                              d_cnt, (unsigned long long)cap);
            PB_CUDA(cudaMemcpyAsync(&E, d_cnt, 8, cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaStreamSynchronize(st));
            if (ensure_index(st)) continue;                 // the window needed the general index path: scan again
            if (E <= cap) break;
            cap = (size_t)E + (size_t)E / 8 + 1024;
            ev_cap_hint_ = cap;
        }
        if (tm) tm->stop(GpuTimers::T_SCAN_SEED, st);

        // order events by (strand, l)
        if (tm) tm->start(GpuTimers::T_SCAN_EVSORT, st);
        uint64_t* ek1 = evk1_.ensure(std::max<size_t>(cap, 1), false, st);
        uint64_t* ev1 = evv1_.ensure(std::max<size_t>(cap, 1), false, st);
        int sbits = 1;
        while ((1 << sbits) < ns) ++sbits;
        int lbits = 1;
        while (((int64_t)1 << lbits) < (int64_t)n) ++lbits;
        const uint64_t* eks = evk0_.get();
        const uint64_t* evs = evv0_.get();
        if (E > 1) {
            // two key fields: l in bits [0,lbits), strand in bits [32, 32+sbits): sort the low field, then the high one (LSD, stable)
            int r1 = sorter_.sort<uint64_t, uint64_t>(evk0_.get(), ek1, evv0_.get(), ev1, (int64_t)E, 0, lbits, st);
            uint64_t* a_k = r1 ? ek1 : evk0_.get(); uint64_t* b_k = r1 ? evk0_.get() : ek1;
            uint64_t* a_v = r1 ? ev1 : evv0_.get(); uint64_t* b_v = r1 ? evv0_.get() : ev1;
            int r2 = sorter_.sort<uint64_t, uint64_t>(a_k, b_k, a_v, b_v, (int64_t)E, 32, 32 + sbits, st);
            eks = r2 ? b_k : a_k;
            evs = r2 ? b_v : a_v;
        }
        uint32_t* seg_lo = seglo_.ensure((size_t)std::max(ns, 1), false, st);
        uint32_t* seg_hi = seghi_.ensure((size_t)std::max(ns, 1), false, st);
        PB_CUDA(cudaMemsetAsync(seg_lo, 0, (size_t)std::max(ns, 1) * 4, st));
        PB_CUDA(cudaMemsetAsync(seg_hi, 0, (size_t)std::max(ns, 1) * 4, st));
        uint32_t* evl = evl_.ensure(std::max<size_t>((size_t)E, 1), false, st);
        int4* states = states_.ensure(std::max<size_t>((size_t)E, 1), false, st);
        if (E > 0) pb200::launch(strand_segments_kernel, (unsigned)((E + TB - 1) / TB), TB, 0, st, eks, (int)E, seg_lo, seg_hi, evl);
        ntiles1_ = (n + FOLD_TILE - 1) / FOLD_TILE + 1;
        const size_t nb_ent = (size_t)std::max(ns, 1) * ntiles1_;
        uint32_t* bounds = bounds_.ensure(nb_ent, false, st);
        if (ns > 0) pb200::launch(tile_bounds_init_kernel, (unsigned)((nb_ent + TB - 1) / TB), TB, 0, st, seg_hi, ns, ntiles1_, bounds);
        if (E > 0) pb200::launch(tile_bounds_fill_kernel, (unsigned)((E + TB - 1) / TB), TB, 0, st, eks, (int)E, FOLD_TILE_SHIFT, ntiles1_, bounds);
        if (tm) tm->stop(GpuTimers::T_SCAN_EVSORT, st);

        if (tm) tm->start(GpuTimers::T_SCAN_EVSCAN, st);
        if (ns > 0) pb200::launch(event_scan_kernel, (unsigned)ns, 256, 0, st, evl, evs, lrp_.get(), seg_lo, seg_hi, states);
        if (tm) tm->stop(GpuTimers::T_SCAN_EVSCAN, st);
        last_events = (int64_t)E;
    }
    // stage 2: fold the local queries (ini order) into Master; init = true starts from (UP 0, EP n), else from the
    // current contents of master_up()/master_ep() (multi-GPU: EP preset to the prefix over earlier ranks, UP = 0)
    void fold(bool init, cudaStream_t st) {
        const int n = cur_n_;
        if (tm) tm->start(GpuTimers::T_SCAN_FOLD, st);
        int32_t* MUP = mup_.ensure((size_t)n, true, st);
        int32_t* MEP = mep_.ensure((size_t)n, true, st);
        pb200::launch(fold_kernel, (unsigned)((n + FOLD_TILE - 1) / FOLD_TILE), FOLD_THREADS, 0, st, evl_.get(), states_.get(), seglo_.get(),
                      bounds_.get(), ntiles1_, 0, cur_nq_, n, MUP, MEP, init ? 1 : 0);
        if (tm) tm->stop(GpuTimers::T_SCAN_FOLD, st);
    }
    int32_t* master_up(cudaStream_t st) { return mup_.ensure((size_t)cur_n_, true, st); }
    int32_t* master_ep(cudaStream_t st) { return mep_.ensure((size_t)cur_n_, true, st); }
    // stage 3: candidate positions from the (global) Master -> device list, returns the count
    uint32_t emit(cudaStream_t st) {
        const int TB = 256;
        const int n = cur_n_;
        if (tm) tm->start(GpuTimers::T_SCAN_EMIT, st);
        uint32_t* flag = tmpA_.ensure((size_t)n + 1, false, st);
        uint32_t* pos = tmpB_.ensure((size_t)n + 1, false, st);
        uint32_t* d_tot = total_.ensure(4, false, st);
        const unsigned nb = (unsigned)((n + TB - 1) / TB);
        pb200::launch(emit_flags_kernel, nb, TB, 0, st, mup_.get(), mep_.get(), n, cur_minsize_, flag);
        scanner_.scan<prim::OpSum, true>(flag, pos, n, d_tot, st);
        uint32_t ncand = 0;
        PB_CUDA(cudaMemcpyAsync(&ncand, d_tot, 4, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        uint32_t* ck = ck_.ensure(std::max<size_t>(ncand, 1), false, st);
        if (ncand) pb200::launch(compact_kernel, nb, TB, 0, st, flag, pos, n, nullptr, ck);
        if (tm) tm->stop(GpuTimers::T_SCAN_EMIT, st);
        cur_ncand_ = ncand;
        return ncand;
    }
    // stage 4: per candidate, strand flag and start of every LOCAL query; initEP (device, n ints) = Master.EP before the
    // first local query, or null.  Appends to the host vectors.
    // stage 4 on the device only: the candidates of the window stay where they are.  Pointers valid until the next window.
    struct DeviceCands { uint32_t ncand = 0; const uint32_t* k = nullptr; const int32_t* lon = nullptr; const int32_t* sp = nullptr; const uint8_t* fwd = nullptr; };
    DeviceCands pass2_device(const int32_t* initEP, cudaStream_t st) {
        const uint32_t ncand = cur_ncand_;
        const int nq = cur_nq_, n = cur_n_;
        DeviceCands dc;
        dc.ncand = ncand;
        if (tm) tm->start(GpuTimers::T_SCAN_PASS2, st);
        if (ncand) {
            int32_t* d_lon = olon_.ensure(ncand, false, st);
            int32_t* d_sp = osp_.ensure((size_t)ncand * std::max(nq, 1), false, st);
            uint8_t* d_fwd = ofwd_.ensure((size_t)ncand * std::max(nq, 1), false, st);
            if (nq) {
                int4* tmp = p2tmp_.ensure((size_t)ncand * nq, false, st);
                const long long tot = (long long)ncand * nq;
                pb200::launch(pass2a_kernel, (unsigned)((tot + 127) / 128), 128, 0, st, ck_.get(), (int)ncand, evl_.get(), states_.get(),
                              seglo_.get(), bounds_.get(), ntiles1_, nq, tmp);
                pb200::launch(pass2b_kernel, (ncand + 127) / 128, 128, 0, st, ck_.get(), (int)ncand, tmp, nq, n, mep_.get(), initEP, d_lon, d_sp,
                              d_fwd);
            } else {
                pb200::launch(pass2b_kernel, (ncand + 127) / 128, 128, 0, st, ck_.get(), (int)ncand, (const int4*)nullptr, 0, n, mep_.get(),
                              initEP, d_lon, d_sp, d_fwd);
            }
            dc.k = ck_.get(); dc.lon = d_lon; dc.sp = d_sp; dc.fwd = d_fwd;
        }
        if (tm) tm->stop(GpuTimers::T_SCAN_PASS2, st);
        PB_CUDA(cudaGetLastError());
        return dc;
    }
    void pass2(const int32_t* initEP, cudaStream_t st, std::vector<int32_t>& out_k, std::vector<int32_t>& out_lon,
               std::vector<int32_t>& out_sp, std::vector<uint8_t>& out_fwd) {
        const uint32_t ncand = cur_ncand_;
        const int nq = cur_nq_, n = cur_n_;
        if (tm) tm->start(GpuTimers::T_SCAN_PASS2, st);
        const size_t base = out_k.size();
        out_k.resize(base + ncand);
        out_lon.resize(base + ncand);
        const size_t bsp = out_sp.size();
        out_sp.resize(bsp + (size_t)ncand * nq);
        out_fwd.resize(bsp + (size_t)ncand * nq);
        if (ncand) {
            int32_t* d_lon = olon_.ensure(ncand, false, st);
            int32_t* d_sp = osp_.ensure((size_t)ncand * std::max(nq, 1), false, st);
            uint8_t* d_fwd = ofwd_.ensure((size_t)ncand * std::max(nq, 1), false, st);
            if (nq) {
                int4* tmp = p2tmp_.ensure((size_t)ncand * nq, false, st);
                const long long tot = (long long)ncand * nq;
                pb200::launch(pass2a_kernel, (unsigned)((tot + 127) / 128), 128, 0, st, ck_.get(), (int)ncand, evl_.get(), states_.get(),
                              seglo_.get(), bounds_.get(), ntiles1_, nq, tmp);
                pb200::launch(pass2b_kernel, (ncand + 127) / 128, 128, 0, st, ck_.get(), (int)ncand, tmp, nq, n, mep_.get(), initEP, d_lon, d_sp,
                              d_fwd);
            } else {
                pb200::launch(pass2b_kernel, (ncand + 127) / 128, 128, 0, st, ck_.get(), (int)ncand, (const int4*)nullptr, 0, n, mep_.get(),
                              initEP, d_lon, d_sp, d_fwd);
            }
            // device -> pinned staging at link speed, then a parallel copy into the (pageable) result vectors
            const size_t cq = (size_t)ncand * std::max(nq, 1);
            int32_t* hk = pin_k_.ensure(ncand);
            int32_t* hl = pin_lon_.ensure(ncand);
            int32_t* hs = pin_sp_.ensure(cq);
            uint8_t* hf = pin_fwd_.ensure(cq);
            PB_CUDA(cudaMemcpyAsync(hk, ck_.get(), (size_t)ncand * 4, cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaMemcpyAsync(hl, d_lon, (size_t)ncand * 4, cudaMemcpyDeviceToHost, st));
            if (nq) {
                PB_CUDA(cudaMemcpyAsync(hs, d_sp, (size_t)ncand * nq * 4, cudaMemcpyDeviceToHost, st));
                PB_CUDA(cudaMemcpyAsync(hf, d_fwd, (size_t)ncand * nq, cudaMemcpyDeviceToHost, st));
            }
            PB_CUDA(cudaStreamSynchronize(st));
            std::memcpy(out_k.data() + base, hk, (size_t)ncand * 4);
            std::memcpy(out_lon.data() + base, hl, (size_t)ncand * 4);
            if (nq) {
                const size_t bytes_sp = (size_t)ncand * nq * 4, bytes_fw = (size_t)ncand * nq;
                const size_t CH = (size_t)1 << 20;
                const long nch_sp = (long)((bytes_sp + CH - 1) / CH), nch_fw = (long)((bytes_fw + CH - 1) / CH);
                uint8_t* dsp = reinterpret_cast<uint8_t*>(out_sp.data() + bsp);
                uint8_t* dfw = out_fwd.data() + bsp;
                parallel_chunks(bytes_sp > 4 * CH ? default_host_threads() : 1, nch_sp + nch_fw, [&](long c) {
                    if (c < nch_sp) { const size_t o = (size_t)c * CH; std::memcpy(dsp + o, reinterpret_cast<const uint8_t*>(hs) + o, std::min(CH, bytes_sp - o)); }
                    else { const size_t o = (size_t)(c - nch_sp) * CH; std::memcpy(dfw + o, hf + o, std::min(CH, bytes_fw - o)); }
                });
            }
        }
        if (tm) tm->stop(GpuTimers::T_SCAN_PASS2, st);
        PB_CUDA(cudaGetLastError());
    }
    void scan(const uint8_t* R, int n, int nq, const std::vector<StrandDesc>& strands, int minsize, cudaStream_t st,
              std::vector<int32_t>& out_k, std::vector<int32_t>& out_lon, std::vector<int32_t>& out_sp, std::vector<uint8_t>& out_fwd) {
        scan_events(R, n, nq, strands, minsize, st);
        fold(true, st);
        emit(st);
        pass2(nullptr, st, out_k, out_lon, out_sp, out_fwd);
    }
    // MUMi of query q (index into the strands given to scan_events): number of reference positions covered by MUMs >= 15
    uint32_t mumi_covered(int q, cudaStream_t st) {
        const int TB = 256;
        const int n = cur_n_;
        const unsigned nb = (unsigned)((n + TB - 1) / TB);
        int32_t* MUP = mup_.ensure((size_t)n, false, st);
        int32_t* MEP = mep_.ensure((size_t)n, false, st);
        pb200::launch(fold_kernel, (unsigned)((n + FOLD_TILE - 1) / FOLD_TILE), FOLD_THREADS, 0, st, evl_.get(), states_.get(), seglo_.get(),
                      bounds_.get(), ntiles1_, 2 * q, 1, n, MUP, MEP, 1);
        uint32_t* a = tmpA_.ensure((size_t)n + 1, false, st);
        uint32_t* b = tmpB_.ensure((size_t)n + 1, false, st);
        uint32_t* c = tmpC_.ensure((size_t)n + 1, false, st);
        uint32_t* d_tot = total_.ensure(4, false, st);
        pb200::launch(mumi_good_kernel, nb, TB, 0, st, MUP, MEP, n, a);
        scanner_.scan<prim::OpMax, true>(a, b, n, nullptr, st);            // b = max EP over earlier good positions
        pb200::launch(mumi_val_kernel, nb, TB, 0, st, a, b, n, c);
        scanner_.scan<prim::OpMax, false>(c, c, n, nullptr, st);           // running max end of the counted MUMs
        pb200::launch(mumi_cover_kernel, nb, TB, 0, st, c, n, a);
        scanner_.scan<prim::OpSum, true>(a, b, n, d_tot, st);
        uint32_t tot = 0;
        PB_CUDA(cudaMemcpyAsync(&tot, d_tot, 4, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        return tot;
    }
    // multi-GPU: device pointers of the window index (for the broadcast from the building rank)
    uint32_t* index_sa() { return sa_ptr_; }
    int32_t* index_lrp() { return lrp_.get(); }
    uint2* index_table() { return table_.get(); }
    size_t index_table_entries() const { return (size_t)1 << (2 * seed_k_); }
    void alloc_index(int n, int minsize, cudaStream_t st) {      // receiving side of the broadcast
        seed_k_ = std::min(MAX_SEED_K, std::max(1, minsize));
        sa_ptr_ = vals0_.ensure((size_t)n, false, st);
        idx_pending_ = false;
        lrp_.ensure((size_t)n, false, st);
        table_.ensure(index_table_entries(), false, st);
    }
    int64_t last_events = 0;

private:
    rsort::RadixSorter sorter_;
    prim::Scanner scanner_;
    DevBuf<uint64_t> keys0_, keys1_, dkeys0_, dkeys1_, evk0_, evk1_, evv0_, evv1_;
    DevBuf<uint32_t> vals0_, vals1_, dvals0_, dvals1_, rank_, tmpA_, tmpB_, tmpC_, cs0_, cs1_, total_, seglo_, seghi_, evl_, ck_, tieflag_, bounds_;
    int ntiles1_ = 1;
    DevBuf<int32_t> lrp_, mup_, mep_, olon_, osp_;
    PinBuf<uint32_t> tie_host_;
    PinBuf<int32_t> pin_k_, pin_lon_, pin_sp_;
    PinBuf<uint8_t> pin_fwd_;
    uint32_t* sa_ptr_ = nullptr;             // suffix array of the current window (lives in one of the sort's value buffers)
    const uint8_t* idx_R_ = nullptr;
    int idx_n_ = 0, idx_minsize_ = 0;
    bool idx_two_bit_ = false, idx_pending_ = false;
    DevBuf<uint8_t> ofwd_;
    DevBuf<int4> states_, p2tmp_;
    DevBuf<uint2> table_;
    DevBuf<StrandDesc> strands_;
    DevBuf<unsigned long long> evcount_;
    size_t ev_cap_hint_ = 0;
    int seed_k_ = MAX_SEED_K;
    int cur_n_ = 0, cur_nq_ = 0, cur_minsize_ = 0;
    uint32_t cur_ncand_ = 0;
};

}  // namespace big
}  // namespace pb200
