// LSD radix sort (8-bit digits) of (key, value) pairs, one-sweep style: one global histogram kernel for all passes,
// then per pass ONE kernel that ranks a 2048-key tile in shared memory (warp match_any multisplit), resolves the
// tile's global digit offsets with a decoupled look-back over single-word (flag|count) tile states, stages the tile
// sorted by digit in shared memory and writes each digit run out contiguously (coalesced).
// Algorithmic HBM traffic per pass: read key+value, write key+value (+ one extra key read for the histograms).
//
// Used for: the suffix-array build (64-bit packed 21-mer keys, prefix-doubling keys) and the (strand, ref start)
// ordering of MEM events.  No reference counterpart (the reference builds a suffix graph online, src/csgmum/csg.c).
#pragma once
#include "util.cuh"

namespace pb200 {
namespace rsort {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 2048 keys per tile: ~64 registers, 32 KB smem -> 5-6 CTAs per SM
constexpr uint32_t ST_AGG = 1u << 30;
constexpr uint32_t ST_INCL = 2u << 30;
constexpr uint32_t ST_MASK = (1u << 30) - 1;

template <class K>
__global__ void __launch_bounds__(256) hist_kernel(const K* __restrict__ keys, int64_t n, int begin_bit, int end_bit, int npass,
                                                   uint32_t* __restrict__ ghist) {
    __shared__ uint32_t h[8 * 256];
    for (int i = threadIdx.x; i < npass * 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        K key = keys[i];
        for (int p = 0; p < npass; ++p) {
            int shift = begin_bit + 8 * p;
            int bits = min(8, end_bit - shift);
            uint32_t d = (uint32_t)(key >> shift) & ((1u << bits) - 1);
            atomicAdd(&h[p * 256 + d], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += blockDim.x)
        if (h[i]) atomicAdd(&ghist[i], h[i]);
}

// exclusive scan of each pass's 256 bins: one block of 256 threads per pass
__global__ void __launch_bounds__(256) scan_hist_kernel(const uint32_t* __restrict__ ghist, uint32_t* __restrict__ gofs) {
    __shared__ uint32_t wsum[8];
    const int p = blockIdx.x, d = threadIdx.x, lane = d & 31, w = d >> 5;
    uint32_t v = ghist[p * 256 + d];
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    uint32_t base = 0;
    for (int i = 0; i < w; ++i) base += wsum[i];
    gofs[p * 256 + d] = base + x - v;
}

template <class K, class V>
__global__ void __launch_bounds__(RS_THREADS, 4) onesweep_kernel(const K* __restrict__ kin, K* __restrict__ kout,
                                                              const V* __restrict__ vin, V* __restrict__ vout, int64_t n, int shift,
                                                              int bits, const uint32_t* __restrict__ gofs,
                                                              volatile uint32_t* status, uint32_t* tile_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K* s_keys = reinterpret_cast<K*>(smem_raw);                                  // RS_TILE
    V* s_vals = reinterpret_cast<V*>(smem_raw + sizeof(K) * RS_TILE);            // RS_TILE
    uint32_t* s_whist = reinterpret_cast<uint32_t*>(smem_raw + (sizeof(K) + sizeof(V)) * RS_TILE);   // [RS_WARPS][256]
    uint32_t* s_base = s_whist + RS_WARPS * 256;                                  // [256] global base - local start
    uint32_t* s_dstart = s_base + 256;                                            // [256] local start of digit run
    __shared__ int s_tile;
    __shared__ uint32_t s_wsum[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mask = (1u << bits) - 1;
    if (tid == 0) s_tile = (int)atomicAdd(tile_counter, 1u);
    for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS) s_whist[i] = 0;
    __syncthreads();
    const int tile = s_tile;
    const int64_t tile_base = (int64_t)tile * RS_TILE;
    const int64_t wbase = tile_base + (int64_t)warp * (RS_ITEMS * 32);
    K key[RS_ITEMS];
    V val[RS_ITEMS];
    uint16_t rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t i = wbase + r * 32 + lane;
        if (i < n) { key[r] = kin[i]; val[r] = vin[i]; }
        else { key[r] = (K)0; val[r] = (V)0; }
    }
    uint32_t* mywh = s_whist + warp * 256;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        uint32_t d = valid ? ((uint32_t)(key[r] >> shift) & mask) : 256u;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        uint32_t before = __popc(peers & ((1u << lane) - 1));
        uint32_t old = 0;
        if (lane == leader && valid) { old = mywh[d]; mywh[d] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = (uint16_t)(old + before);
        __syncwarp();
    }
    __syncthreads();
    // thread d owns digit d: offsets of each warp inside the digit run, tile count, look-back
    {
        const int d = tid;
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { uint32_t c = s_whist[w * 256 + d]; s_whist[w * 256 + d] = tot; tot += c; }
        // local exclusive scan over digits -> start of the digit run inside the staged tile
        uint32_t x = tot;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_wsum[warp] = x;
        // publish / look back while the other warps finish their scans
        uint32_t excl = 0;
        if (tile == 0) {
            status[(int64_t)tile * 256 + d] = tot | ST_INCL;
        } else {
            status[(int64_t)tile * 256 + d] = tot | ST_AGG;
            // look back over the predecessor tiles, 8 status words in flight per step (the first wave of resident tiles has
            // to walk back hundreds of tiles; one dependent L2 round trip per tile would serialise the whole pass)
            int t = tile - 1;
            bool done = false;
            while (!done) {
                uint32_t v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { v[u] = 2u << 30; if (t - u >= 0) v[u] = status[(int64_t)(t - u) * 256 + d]; }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (done) break;
                    const uint32_t f = v[u] & ~ST_MASK;
                    if (f == 0) break;                   // predecessor not published yet: re-read from here
                    excl += v[u] & ST_MASK;
                    --t;
                    if (f == ST_INCL) done = true;
                }
            }
            status[(int64_t)tile * 256 + d] = (excl + tot) | ST_INCL;
        }
        __syncthreads();
        uint32_t wb = 0;
        for (int i = 0; i < warp; ++i) wb += s_wsum[i];
        uint32_t dstart = wb + x - tot;
        s_dstart[d] = dstart;
        s_base[d] = gofs[d] + excl - dstart;      // global position = s_base[digit] + local position (mod 2^32)
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)(key[r] >> shift) & mask;
            uint32_t lp = s_dstart[d] + s_whist[warp * 256 + d] + rank[r];
            s_keys[lp] = key[r];
            s_vals[lp] = val[r];
        }
    }
    __syncthreads();
    const int cnt = (int)min((int64_t)RS_TILE, n - tile_base);
    for (int i = tid; i < cnt; i += RS_THREADS) {
        K k = s_keys[i];
        uint32_t d = (uint32_t)(k >> shift) & mask;
        uint32_t gp = s_base[d] + (uint32_t)i;
        kout[gp] = k;
        vout[gp] = s_vals[i];
    }
}

class RadixSorter {
public:
    // Sorts n pairs by key bits [begin_bit, end_bit).  Buffers 0 hold the input; returns the index (0/1) of the
    // buffer pair holding the sorted output.  Stable.
    template <class K, class V>
    int sort(K* k0, K* k1, V* v0, V* v1, int64_t n, int begin_bit, int end_bit, cudaStream_t st) {
        if (n <= 1 || end_bit <= begin_bit) return 0;
        if (n > (int64_t)ST_MASK) throw CudaError("radix sort: n too large");
        const int npass = (end_bit - begin_bit + 7) / 8;
        const int64_t tiles = (n + RS_TILE - 1) / RS_TILE;
        const size_t words = (size_t)npass * 256 * 2 + (size_t)npass + (size_t)npass * tiles * 256;
        uint32_t* tmp = tmp_.ensure(words, false, st);
        PB_CUDA(cudaMemsetAsync(tmp, 0, words * sizeof(uint32_t), st));
        uint32_t* ghist = tmp;
        uint32_t* gofs = tmp + (size_t)npass * 256;
        uint32_t* counters = gofs + (size_t)npass * 256;
        uint32_t* status = counters + npass;
        int hb = (int)std::min<int64_t>((n + 256 * 16 - 1) / (256 * 16), 148 * 8);
        pb200::launch(hist_kernel<K>, hb, 256, 0, st, k0, n, begin_bit, end_bit, npass, ghist);
        pb200::launch(scan_hist_kernel, npass, 256, 0, st, ghist, gofs);
        const size_t smem = (sizeof(K) + sizeof(V)) * RS_TILE + (RS_WARPS * 256 + 512) * sizeof(uint32_t);
        static bool attr_set[2][2] = {{false, false}, {false, false}};
        bool& a = attr_set[sizeof(K) == 8][sizeof(V) == 8];
        if (!a) {
            PB_CUDA(cudaFuncSetAttribute(onesweep_kernel<K, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            a = true;
        }
        K* kin = k0; K* kout = k1; V* vin = v0; V* vout = v1;
        int res = 0;
        for (int p = 0; p < npass; ++p) {
            int shift = begin_bit + 8 * p;
            int bits = std::min(8, end_bit - shift);
            pb200::launch(onesweep_kernel<K, V>, (unsigned)tiles, RS_THREADS, smem, st, kin, kout, vin, vout, n, shift, bits, gofs + (size_t)p * 256,
                                                                            status + (size_t)p * tiles * 256, counters + p);
            std::swap(kin, kout);
            std::swap(vin, vout);
            res ^= 1;
        }
        PB_CUDA(cudaGetLastError());
        return res;
    }
private:
    DevBuf<uint32_t> tmp_;
};

}  // namespace rsort
}  // namespace pb200
