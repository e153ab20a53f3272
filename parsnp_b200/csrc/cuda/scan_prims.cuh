// Device-wide scan primitives (three-kernel reduce / scan-of-partials / downsweep) used by the index build and the
// candidate compaction.  Small helper kernels, HBM-stream bound.
#pragma once
#include "util.cuh"

namespace pb200 {
namespace prim {

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

struct OpSum { __device__ static uint32_t id() { return 0u; } __device__ static uint32_t f(uint32_t a, uint32_t b) { return a + b; } };
struct OpMax { __device__ static uint32_t id() { return 0u; } __device__ static uint32_t f(uint32_t a, uint32_t b) { return a > b ? a : b; } };

template <class Op>
__device__ inline uint32_t block_inclusive_scan(uint32_t v, uint32_t* s_w /*[8]*/, uint32_t& block_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x = Op::f(y, x); }
    if (lane == 31) s_w[w] = x;
    __syncthreads();
    uint32_t base = Op::id();
    for (int i = 0; i < w; ++i) base = Op::f(base, s_w[i]);
    uint32_t tot = Op::id();
    for (int i = 0; i < SC_THREADS / 32; ++i) tot = Op::f(tot, s_w[i]);
    block_total = tot;
    __syncthreads();
    return Op::f(base, x);
}

// pass 1: per-tile totals
template <class Op>
__global__ void __launch_bounds__(SC_THREADS) tile_reduce_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ partial) {
    __shared__ uint32_t s_w[8];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    uint32_t acc = Op::id();
#pragma unroll
    for (int r = 0; r < SC_ITEMS; ++r) { int64_t i = base + r; if (i < n) acc = Op::f(acc, in[i]); }
    uint32_t tot;
    block_inclusive_scan<Op>(acc, s_w, tot);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
// pass 2: exclusive scan of the partials by one block (loops)
template <class Op>
__global__ void __launch_bounds__(SC_THREADS) partial_scan_kernel(uint32_t* __restrict__ partial, int64_t m, uint32_t* __restrict__ total) {
    __shared__ uint32_t s_w[8];
    uint32_t carry = Op::id();
    for (int64_t b = 0; b < m; b += SC_THREADS) {
        int64_t i = b + threadIdx.x;
        uint32_t v = i < m ? partial[i] : Op::id();
        uint32_t tot;
        uint32_t inc = block_inclusive_scan<Op>(v, s_w, tot);
        // exclusive value = carry (+) everything before i
        uint32_t prev = __shfl_up_sync(0xffffffffu, inc, 1);
        __shared__ uint32_t s_prev[SC_THREADS];
        s_prev[threadIdx.x] = inc;
        __syncthreads();
        uint32_t excl = threadIdx.x == 0 ? Op::id() : s_prev[threadIdx.x - 1];
        (void)prev;
        if (i < m) partial[i] = Op::f(carry, excl);
        carry = Op::f(carry, tot);
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}
// pass 3: downsweep; EXCLUSIVE selects exclusive/inclusive output
template <class Op, bool EXCLUSIVE>
__global__ void __launch_bounds__(SC_THREADS) tile_scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n,
                                                               const uint32_t* __restrict__ partial) {
    __shared__ uint32_t s_w[8];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    uint32_t v[SC_ITEMS];
    uint32_t acc = Op::id();
#pragma unroll
    for (int r = 0; r < SC_ITEMS; ++r) { int64_t i = base + r; v[r] = i < n ? in[i] : Op::id(); acc = Op::f(acc, v[r]); }
    uint32_t tot;
    uint32_t inc = block_inclusive_scan<Op>(acc, s_w, tot);
    // exclusive prefix of this thread = carry(block) (+) scan of previous threads
    __shared__ uint32_t s_inc[SC_THREADS];
    s_inc[threadIdx.x] = inc;
    __syncthreads();
    uint32_t run = Op::f(partial[blockIdx.x], threadIdx.x == 0 ? Op::id() : s_inc[threadIdx.x - 1]);
#pragma unroll
    for (int r = 0; r < SC_ITEMS; ++r) {
        int64_t i = base + r;
        uint32_t nxt = Op::f(run, v[r]);
        if (i < n) out[i] = EXCLUSIVE ? run : nxt;
        run = nxt;
    }
}

class Scanner {
public:
    // out may alias in. If total != nullptr the grand total is written there (device pointer).
    template <class Op, bool EXCLUSIVE>
    void scan(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* d_total, cudaStream_t st) {
        if (n <= 0) { if (d_total) PB_CUDA(cudaMemsetAsync(d_total, 0, 4, st)); return; }
        int64_t tiles = (n + SC_TILE - 1) / SC_TILE;
        uint32_t* partial = part_.ensure((size_t)tiles, false, st);
        pb200::launch(tile_reduce_kernel<Op>, (unsigned)tiles, SC_THREADS, 0, st, in, n, partial);
        pb200::launch(partial_scan_kernel<Op>, 1, SC_THREADS, 0, st, partial, tiles, d_total);
        pb200::launch(tile_scan_kernel<Op, EXCLUSIVE>, (unsigned)tiles, SC_THREADS, 0, st, in, out, n, partial);
        PB_CUDA(cudaGetLastError());
    }
private:
    DevBuf<uint32_t> part_;
};

}  // namespace prim
}  // namespace pb200
