// LSD radix sort (8-bit digits) of (key, value) pairs.  Per pass three kernels, none of which waits on another block:
//   tile_hist_kernel   per-tile digit histogram (reads the keys once)                      -> counts[tile][digit]
//   digit_scan_kernel  8 digits per block: exclusive scan over the tiles (in place) + the digits' totals
//   scatter_kernel     ranks a 2048-key tile in shared memory (warp match_any multisplit), stages the tile sorted by
//                      digit in shared memory and writes each digit run out contiguously (coalesced)
// The global offset of a digit is the sum of the totals of the smaller digits: digit_scan_kernel leaves the totals, every
// scatter block adds them up itself (256 values) - no separate histogram pass over the keys.
// The first pass can take its keys straight from the window text (N-free windows: the packed 16-mer of suffix i, value i),
// so the suffix-array build has no key-generation kernel and pass 1 reads 1 byte instead of 8 per suffix.
// (A one-sweep variant with decoupled look-back was measured first: with ~600 resident tiles every tile walks hundreds of
//  predecessor states and the pass time stayed at ~75 us whatever the record size - profiles/r01_*; the split form moves
//  8 (4) more bytes per key and pass but has no cross-block dependency.)
// HBM traffic per pass: keys read twice, values once, both written once.
//
// Used for: the suffix-array build (packed 16-mer / 21-mer keys, prefix-doubling keys) and the (strand, ref start)
// ordering of MEM events.  No reference counterpart (the reference builds a suffix graph online, src/csgmum/csg.c).
#pragma once
#include "util.cuh"

namespace pb200 {
namespace rsort {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 2048 keys per tile: ~64 registers, 32 KB smem -> 5-6 CTAs per SM
constexpr uint32_t ST_MASK = (1u << 30) - 1;

template <class K>
__global__ void __launch_bounds__(RS_THREADS) tile_hist_kernel(const K* __restrict__ kin, int64_t n, int shift, int bits,
                                                               uint32_t* __restrict__ counts, int64_t tiles, const uint8_t* __restrict__ text) {
    __shared__ uint32_t h[256];
    const int tid = threadIdx.x;
    h[tid] = 0;
    __syncthreads();
    const int64_t tile_base = (int64_t)blockIdx.x * RS_TILE;
    const uint32_t mask = (1u << bits) - 1;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t i = tile_base + r * RS_THREADS + tid;
        if (i < n) {
            uint32_t d;
            if (text) d = (shift == 0 ? text_key2_low8(text, n, i) : (uint32_t)(text_key2(text, n, i) >> shift)) & mask;
            else d = (uint32_t)(kin[i] >> shift) & mask;
            atomicAdd(&h[d], 1u);
        }
    }
    __syncthreads();
    // (tile-major: one coalesced 1 KB row per block, read back the same way by the scatter.  Measured and dropped: one wave of
    //  persistent blocks with the next tile's keys in flight - 19-25 us instead of 16 for 2 442 tiles, 40 % slower at 7 324)
    counts[(int64_t)blockIdx.x * 256 + tid] = h[tid];
}

// counts[tile][digit] -> per digit the exclusive prefix over the tiles (in place); totals[d] = number of keys with digit d.
// A block takes SCAN_DPB digits and cuts the tiles into SCAN_CHUNKS runs, one thread per (run, digit).  Runs of at most
// SCAN_REG tiles are held in registers: one round of independent loads, a scan over the run totals in shared memory, stores
// (the kernel is pure latency: 2.5 MB in L2 for a 5 Mbp window).  Longer runs are walked twice.
// (The first form - one block per digit over a digit-major matrix - made every histogram write and every offset read of the
// other two kernels a strided sector access and took 11.5 us of a 75 us pass for 2 442 tiles.)
constexpr int SCAN_DPB = 2, SCAN_CHUNKS = 512, SCAN_REG = 16;
__global__ void __launch_bounds__(SCAN_DPB * SCAN_CHUNKS) digit_scan_kernel(uint32_t* __restrict__ counts, int64_t tiles, uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_run[SCAN_DPB][SCAN_CHUNKS + 1];
    const int dl = threadIdx.x & (SCAN_DPB - 1), ch = threadIdx.x / SCAN_DPB;
    const int d = blockIdx.x * SCAN_DPB + dl;
    const int64_t per = (tiles + SCAN_CHUNKS - 1) / SCAN_CHUNKS;
    const int64_t t0 = min(tiles, (int64_t)ch * per), t1 = min(tiles, t0 + per);
    const bool in_regs = per <= SCAN_REG;
    uint32_t v[SCAN_REG];
    uint32_t sum = 0;
    if (in_regs) {
#pragma unroll
        for (int i = 0; i < SCAN_REG; ++i) { v[i] = t0 + i < t1 ? counts[(t0 + i) * 256 + d] : 0u; sum += v[i]; }
    } else {
#pragma unroll 4
        for (int64_t t = t0; t < t1; ++t) sum += counts[t * 256 + d];
    }
    s_run[dl][ch] = sum;
    __syncthreads();
    // warp w scans the run totals of digit w: SCAN_CHUNKS values, SCAN_CHUNKS / 32 per lane
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w < SCAN_DPB) {
        constexpr int PL = SCAN_CHUNKS / 32;
        uint32_t x = 0;
        for (int i = 0; i < PL; ++i) x += s_run[w][lane * PL + i];
        const uint32_t mine = x;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        uint32_t run = x - mine;
        for (int i = 0; i < PL; ++i) { const uint32_t c = s_run[w][lane * PL + i]; s_run[w][lane * PL + i] = run; run += c; }
        if (lane == 31) totals[blockIdx.x * SCAN_DPB + w] = x;
    }
    __syncthreads();
    uint32_t run = s_run[dl][ch];
    if (in_regs) {
#pragma unroll
        for (int i = 0; i < SCAN_REG; ++i) if (t0 + i < t1) { counts[(t0 + i) * 256 + d] = run; run += v[i]; }
    } else {
#pragma unroll 4
        for (int64_t t = t0; t < t1; ++t) { const uint32_t c = counts[t * 256 + d]; counts[t * 256 + d] = run; run += c; }
    }
}

template <class K, class V, bool FROM_TEXT>
__global__ void __launch_bounds__(RS_THREADS, (sizeof(K) == 4 ? (FROM_TEXT ? 5 : 6) : 4)) scatter_kernel(const K* __restrict__ kin, K* __restrict__ kout,
                                                                const V* __restrict__ vin, V* __restrict__ vout, int64_t n, int shift,
                                                                int bits, const uint32_t* __restrict__ offsets, int64_t tiles,
                                                                const uint32_t* __restrict__ totals, const uint8_t* __restrict__ text) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K* s_keys = reinterpret_cast<K*>(smem_raw);                                  // RS_TILE
    V* s_vals = reinterpret_cast<V*>(smem_raw + sizeof(K) * RS_TILE);            // RS_TILE
    uint32_t* s_whist = reinterpret_cast<uint32_t*>(smem_raw + (sizeof(K) + sizeof(V)) * RS_TILE);   // [RS_WARPS][256]
    uint32_t* s_base = s_whist + RS_WARPS * 256;                                  // [256] global base - local start
    uint32_t* s_dstart = s_base + 256;                                            // [256] local start of digit run
    __shared__ uint32_t s_wsum[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mask = (1u << bits) - 1;
    for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS) s_whist[i] = 0;
    const int64_t tile = blockIdx.x;
    const int64_t tile_base = tile * RS_TILE;
    const int64_t wbase = tile_base + (int64_t)warp * (RS_ITEMS * 32);
    // global start of digit `tid` for this tile = keys with a smaller digit (exclusive scan of the 256 totals) + the
    // digit's keys in earlier tiles
    {
        const uint32_t t = totals[tid];
        uint32_t x = t;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_wsum[warp] = x;
        __syncthreads();
        uint32_t wb = 0;
        for (int i = 0; i < warp; ++i) wb += s_wsum[i];
        s_base[tid] = offsets[tile * 256 + tid] + wb + x - t;                 // parked in shared memory (turned into base - local start below)
    }
    K key[RS_ITEMS];
    V val[RS_ITEMS];
    uint16_t rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            if (FROM_TEXT) { key[r] = (K)text_key2(text, n, i); val[r] = (V)i; }  // first pass of the suffix-array build
            else { key[r] = kin[i]; val[r] = vin[i]; }
        }
        else { key[r] = (K)0; val[r] = (V)0; }
    }
    __syncthreads();
    uint32_t* mywh = s_whist + warp * 256;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        uint32_t d = valid ? ((uint32_t)(key[r] >> shift) & mask) : 256u;
        // lanes holding the same digit: 9 ballots (8 digit bits + validity) - cheaper here than match.any, whose result
        // latency dominated the kernel (ncu: 39 % of the stall samples on the instruction consuming it)
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int bit = 0; bit < 9; ++bit) {
            const uint32_t vote = __ballot_sync(0xffffffffu, (d >> bit) & 1u);
            peers &= ((d >> bit) & 1u) ? vote : ~vote;
        }
        int leader = __ffs(peers) - 1;
        uint32_t before = __popc(peers & ((1u << lane) - 1));
        uint32_t old = 0;
        if (lane == leader && valid) { old = mywh[d]; mywh[d] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = (uint16_t)(old + before);
        __syncwarp();
    }
    __syncthreads();
    {
        const int d = tid;
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { uint32_t c = s_whist[w * 256 + d]; s_whist[w * 256 + d] = tot; tot += c; }
        uint32_t x = tot;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_wsum[warp] = x;
        __syncthreads();
        uint32_t wb = 0;
        for (int i = 0; i < warp; ++i) wb += s_wsum[i];
        uint32_t dstart = wb + x - tot;
        s_dstart[d] = dstart;
        s_base[d] = s_base[d] - dstart;           // global position = s_base[digit] + local position (mod 2^32)
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
    int sort(K* k0, K* k1, V* v0, V* v1, int64_t n, int begin_bit, int end_bit, cudaStream_t st, const uint8_t* first_pass_text = nullptr) {
        if (n <= 1 || end_bit <= begin_bit) {
            if (first_pass_text) throw CudaError("radix sort: a text-keyed sort needs at least one pass");
            return 0;
        }
        if (n > (int64_t)ST_MASK) throw CudaError("radix sort: n too large");
        const int npass = (end_bit - begin_bit + 7) / 8;
        const int64_t tiles = (n + RS_TILE - 1) / RS_TILE;
        const size_t words = (size_t)256 + (size_t)256 * tiles;
        uint32_t* tmp = tmp_.ensure(words, false, st);
        uint32_t* totals = tmp;
        uint32_t* counts = tmp + 256;
        const size_t smem = (sizeof(K) + sizeof(V)) * RS_TILE + (RS_WARPS * 256 + 512) * sizeof(uint32_t);
        static bool attr_set[2][2] = {{false, false}, {false, false}};
        bool& a = attr_set[sizeof(K) == 8][sizeof(V) == 8];
        if (!a) {
            PB_CUDA(cudaFuncSetAttribute(scatter_kernel<K, V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PB_CUDA(cudaFuncSetAttribute(scatter_kernel<K, V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            a = true;
        }
        K* kin = k0; K* kout = k1; V* vin = v0; V* vout = v1;
        int res = 0;
        for (int p = 0; p < npass; ++p) {
            int shift = begin_bit + 8 * p;
            int bits = std::min(8, end_bit - shift);
            const uint8_t* text = p == 0 ? first_pass_text : nullptr;
            pb200::launch(tile_hist_kernel<K>, (unsigned)tiles, RS_THREADS, 0, st, kin, n, shift, bits, counts, tiles, text);
            pb200::launch(digit_scan_kernel, 256 / SCAN_DPB, SCAN_DPB * SCAN_CHUNKS, 0, st, counts, tiles, totals);
            if (text) pb200::launch(scatter_kernel<K, V, true>, (unsigned)tiles, RS_THREADS, smem, st, kin, kout, vin, vout, n, shift, bits, counts, tiles, totals, text);
            else pb200::launch(scatter_kernel<K, V, false>, (unsigned)tiles, RS_THREADS, smem, st, kin, kout, vin, vout, n, shift, bits, counts, tiles, totals, text);
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
