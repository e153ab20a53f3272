// Small CUDA utilities: error checking, RAII device buffers, event timers.
#pragma once
#include <chrono>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <utility>
#include <stdexcept>
#include <string>
#include <vector>

namespace pb200 {

struct CudaError : public std::runtime_error {
    explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};

#define PB_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            throw pb200::CudaError(std::string(#call) + " failed: " + cudaGetErrorString(_e) + " at " + \
                                   __FILE__ + ":" + std::to_string(__LINE__));                          \
    } while (0)

// growable device buffer (never shrinks); contents are NOT preserved on growth unless keep=true
template <class T>
class DevBuf {
public:
    DevBuf() {}
    ~DevBuf() { if (p_) cudaFreeAsync(p_, 0); }     // stream-ordered free back into the pool: no device-wide synchronisation
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    T* get() const { return p_; }
    size_t capacity() const { return cap_; }
    T* ensure(size_t n, bool keep = false, cudaStream_t st = 0) {
        if (n <= cap_) return p_;
        size_t ncap = n + n / 4 + 64;
        T* q = nullptr;
        // stream-ordered allocation from the device's default pool (kept warm across engines: see CudaEngine ctor)
        PB_CUDA(cudaMallocAsync((void**)&q, ncap * sizeof(T), st));
        if (keep && p_ && cap_) PB_CUDA(cudaMemcpyAsync(q, p_, cap_ * sizeof(T), cudaMemcpyDeviceToDevice, st));
        if (p_) PB_CUDA(cudaFreeAsync(p_, st));
        p_ = q;
        cap_ = ncap;
        return p_;
    }
    void release(cudaStream_t st = 0) { if (p_) cudaFreeAsync(p_, st); p_ = nullptr; cap_ = 0; }
private:
    T* p_ = nullptr;
    size_t cap_ = 0;
};

// Process-wide cache of pinned host blocks: cudaHostAlloc costs milliseconds, and an engine (with its staging buffers) is
// created and destroyed by every pb200_align call.  Blocks are handed back on destruction and never returned to the OS.
class PinnedPool {
public:
    static void* take(size_t bytes, size_t& cap_bytes) {
        {
            std::lock_guard<std::mutex> lk(mu());
            auto& f = free_list();
            size_t best = f.size();
            for (size_t i = 0; i < f.size(); ++i)
                if (f[i].second >= bytes && (best == f.size() || f[i].second < f[best].second)) best = i;
            if (best != f.size()) {
                void* p = f[best].first;
                cap_bytes = f[best].second;
                f.erase(f.begin() + (long)best);
                return p;
            }
        }
        void* p = nullptr;
        cap_bytes = bytes + bytes / 4 + 4096;
        PB_CUDA(cudaHostAlloc(&p, cap_bytes, cudaHostAllocPortable));
        return p;
    }
    static void give(void* p, size_t cap_bytes) {
        if (!p) return;
        std::lock_guard<std::mutex> lk(mu());
        free_list().emplace_back(p, cap_bytes);
    }
private:
    static std::mutex& mu() { static std::mutex m; return m; }
    static std::vector<std::pair<void*, size_t>>& free_list() { static std::vector<std::pair<void*, size_t>> f; return f; }
};

// pinned host buffer (from the process-wide pool); ensure(n, keep) preserves the first `keep` elements on growth
template <class T>
class PinBuf {
public:
    PinBuf() {}
    ~PinBuf() { PinnedPool::give(p_, cap_bytes_); }
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
    T* ensure(size_t n, size_t keep = 0) {
        if (n * sizeof(T) <= cap_bytes_ && p_) return p_;
        size_t ncap = 0;
        T* q = (T*)PinnedPool::take(std::max<size_t>(n, 16) * sizeof(T), ncap);
        if (keep && p_) memcpy(q, p_, keep * sizeof(T));
        PinnedPool::give(p_, cap_bytes_);
        p_ = q;
        cap_bytes_ = ncap;
        return p_;
    }
    T* get() const { return p_; }
private:
    T* p_ = nullptr;
    size_t cap_bytes_ = 0;
};

// Named GPU timers: CUDA event pairs recorded on the engine stream WITHOUT synchronising; collect() (called once
// the stream is idle) folds the elapsed times into per-group accumulators.  Always on; ~2 event records per group.
class GpuTimers {
public:
    enum { T_INDEX_KEYS = 0, T_INDEX_SORT, T_INDEX_DOUBLING, T_INDEX_LCP, T_INDEX_TABLE, T_SCAN_SEED, T_SCAN_EVSORT, T_SCAN_EVSCAN,
           T_SCAN_FOLD, T_SCAN_EMIT, T_SCAN_PASS2, T_SMALL, T_SMALL_B, T_SMALL_C, T_SMALL_ACCEPT, T_COUNT };
    static const char* names() {
        return "index_keys_ms,index_sort_ms,index_doubling_ms,index_lcp_ms,index_table_ms,scan_seed_ms,scan_evsort_ms,scan_evscan_ms,"
               "scan_fold_ms,scan_emit_ms,scan_pass2_ms,small_regions_ms,small_b_ms,small_c_ms,small_accept_ms,"
               "n_index_keys,n_index_sort,n_index_doubling,n_index_lcp,n_index_table,n_scan_seed,n_scan_evsort,n_scan_evscan,"
               "n_scan_fold,n_scan_emit,n_scan_pass2,n_small_regions,n_small_b,n_small_c,n_small_accept";
    }
    GpuTimers() { reset(); }
    ~GpuTimers() { for (auto& e : pool_) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); } }
    bool enabled = true;
    void start(int id, cudaStream_t s) {
        if (!enabled) return;
        if (used_ == pool_.size()) {
            Pair p; p.id = 0;
            cudaEventCreate(&p.a); cudaEventCreate(&p.b);
            pool_.push_back(p);
        }
        cur_ = used_++;
        pool_[cur_].id = id;
        cudaEventRecord(pool_[cur_].a, s);
    }
    void stop(int id, cudaStream_t s) {
        if (!enabled) return;
        (void)id;
        cudaEventRecord(pool_[cur_].b, s);
        if (used_ >= 4096) collect(s);
    }
    // requires all recorded work to have completed (or completes it)
    void collect(cudaStream_t s) {
        if (!used_) return;
        cudaStreamSynchronize(s);
        for (size_t i = 0; i < used_; ++i) {
            float t = 0;
            if (cudaEventElapsedTime(&t, pool_[i].a, pool_[i].b) == cudaSuccess) { ms[pool_[i].id] += t; cnt[pool_[i].id] += 1; }
        }
        used_ = 0;
    }
    void reset() { for (int i = 0; i < T_COUNT; ++i) { ms[i] = 0; cnt[i] = 0; } used_ = 0; }
    double ms[T_COUNT];
    double cnt[T_COUNT];
private:
    struct Pair { cudaEvent_t a, b; int id; };
    std::vector<Pair> pool_;
    size_t used_ = 0, cur_ = 0;
};

// every kernel launch of the library goes through here so that the engine can report how many kernels it launched
inline double wall_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
extern int64_t g_kernel_launches;
template <class... KArgs, class... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    ++g_kernel_launches;
    kernel<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
}

// unaligned 8-byte load: three aligned 32-bit words + two funnel shifts (SHF); may read up to 11 bytes past p: every text buffer is
// padded.  (The first form - two aligned 64-bit loads and a 64-bit shift pair - was ~14 instructions of 32-bit arithmetic per
// call and a quarter of seed_extend_kernel's instruction stream.)
__device__ __forceinline__ uint64_t load8u(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3) * 8;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    return ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
}
// four 2-bit base codes (one per byte, first base in byte 0) -> 8 bits, first base most significant
__device__ __forceinline__ uint32_t pack4x2(uint32_t w) { return ((w & 0x03030303u) * 0x40100401u) >> 24; }
// key of suffix i of an N-free window: its first 16 bases at 2 bits each, first base most significant, past the end = 0
__device__ __forceinline__ uint32_t text_key2(const uint8_t* __restrict__ R, int64_t n, int64_t i) {
    uint64_t w0 = load8u(R + i), w1 = load8u(R + i + 8);
    const int64_t valid = n - i;
    if (valid < 16) {
        w0 = valid >= 8 ? w0 : (valid <= 0 ? 0ull : (w0 & ((1ull << (8 * valid)) - 1ull)));
        w1 = valid <= 8 ? 0ull : (w1 & ((1ull << (8 * (valid - 8))) - 1ull));
    }
    return (pack4x2((uint32_t)w0) << 24) | (pack4x2((uint32_t)(w0 >> 32)) << 16) | (pack4x2((uint32_t)w1) << 8) | pack4x2((uint32_t)(w1 >> 32));
}

// its low 8 bits (bases 12..15 of the suffix) without reading the other twelve
__device__ __forceinline__ uint32_t text_key2_low8(const uint8_t* __restrict__ R, int64_t n, int64_t i) {
    uint32_t w = (uint32_t)load8u(R + i + 12);
    const int64_t valid = n - (i + 12);
    if (valid < 4) w = valid <= 0 ? 0u : (w & ((1u << (8 * valid)) - 1u));
    return pack4x2(w);
}

// ---- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) global -> shared with mbarrier completion
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's earlier generic-proxy accesses to shared memory before its later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n }" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n MBAR_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra MBAR_DONE;\n bra MBAR_WAIT;\n MBAR_DONE:\n }" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// dst, src 16-byte aligned, bytes a multiple of 16; completion is signalled on `bar` (complete_tx)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace pb200
