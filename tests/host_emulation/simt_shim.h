// TEST INFRASTRUCTURE — a small SIMT emulator for kernels whose threads cooperate: every CUDA thread of a block runs as
// an OS thread; __syncthreads / __syncthreads_or are block barriers, the warp intrinsics (__shfl_*_sync, __ballot_sync,
// __any_sync, __syncwarp) are 32-thread rendezvous with a value exchange, shared memory is ordinary static storage (blocks of
// a launch run one after the other), atomics are real atomics.  A thread that returns from the kernel leaves its barriers, as
// an exited CUDA thread does.  Enough to run csrc/at_seirm.cu on the host (tests/host_emulation/seirm_emulation.cpp); slow
// (a rendezvous costs microseconds), so only for small populations.  Never linked into the product.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static thread_local dim3 threadIdx;
static dim3 blockIdx, blockDim, gridDim;

struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct double2 { double x, y; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

// ---------------------------------------------------------------------------------------------- rendezvous
struct emu_barrier {
    std::mutex m;
    std::condition_variable cv;
    int expected = 0, arrived = 0;
    unsigned long long generation = 0;
    int acc_or = 0, result_or = 0;
    unsigned long long slot[32];                         // warp exchange: one 64-bit value per lane
    unsigned pred_bits = 0, result_bits = 0;

    void reset(int n) { expected = n; arrived = 0; acc_or = 0; pred_bits = 0; }
    void complete_locked()
    {
        result_or = acc_or; acc_or = 0;
        result_bits = pred_bits; pred_bits = 0;
        arrived = 0;
        ++generation;
        cv.notify_all();
    }
    // returns after every participant of this generation has arrived
    void arrive_and_wait(int or_value = 0, int lane = -1, bool pred = false)
    {
        std::unique_lock<std::mutex> lk(m);
        acc_or |= or_value;
        if (lane >= 0 && pred) pred_bits |= 1u << lane;
        const unsigned long long gen = generation;
        if (++arrived >= expected) complete_locked();
        else cv.wait(lk, [&] { return generation != gen; });
    }
    void leave()                                          // the thread exits the kernel
    {
        std::unique_lock<std::mutex> lk(m);
        --expected;
        if (expected > 0 && arrived >= expected) complete_locked();
    }
};

struct emu_block_ctx {
    emu_barrier block;
    std::vector<emu_barrier> warps;
};
static emu_block_ctx* g_emu_ctx = nullptr;
static thread_local int t_emu_warp = 0, t_emu_lane = 0;

static inline void __syncthreads() { g_emu_ctx->block.arrive_and_wait(); }
static inline int __syncthreads_or(int pred)
{
    emu_barrier& b = g_emu_ctx->block;
    b.arrive_and_wait(pred != 0);
    const int r = b.result_or;
    b.arrive_and_wait();                                  // nobody starts the next reduction before everybody has read
    return r;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { g_emu_ctx->warps[t_emu_warp].arrive_and_wait(); }

template <class T> static inline T emu_exchange(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "warp exchange of up to 64-bit values");
    emu_barrier& w = g_emu_ctx->warps[t_emu_warp];
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.slot[t_emu_lane] = raw;
    w.arrive_and_wait();
    T out = v;
    if (src_lane >= 0 && src_lane < 32) { raw = w.slot[src_lane]; memcpy(&out, &raw, sizeof(T)); }
    w.arrive_and_wait();                                  // slots are reused by the next exchange
    return out;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) { (void)width; return emu_exchange(v, src & 31); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32)
{
    (void)width;
    const int src = t_emu_lane + (int)d;
    return emu_exchange(v, src < 32 ? src : t_emu_lane);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32)
{
    (void)width;
    const int src = t_emu_lane - (int)d;
    return emu_exchange(v, src >= 0 ? src : t_emu_lane);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) { (void)width; return emu_exchange(v, (t_emu_lane ^ mask) & 31); }
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    emu_barrier& w = g_emu_ctx->warps[t_emu_warp];
    w.arrive_and_wait(0, t_emu_lane, pred != 0);
    const unsigned r = w.result_bits & mask;
    w.arrive_and_wait();
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0u; }

// ---------------------------------------------------------------------------------------------- intrinsics
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcv(const T* p)
{
    T out;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    memcpy(&out, reinterpret_cast<const void*>(p), sizeof(T));
    return out;
}
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
#define __expf(x) expf(x)          /* glibc declares its own __expf */
static inline long long __double2ll_rn(double x) { return llrint(x); }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift)
{
    return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (shift & 31u));
}
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __vminu4(unsigned a, unsigned b)        // per-byte unsigned minimum
{
    unsigned r = 0;
    for (int k = 0; k < 32; k += 8) { unsigned x = (a >> k) & 255u, y = (b >> k) & 255u; r |= (x < y ? x : y) << k; }
    return r;
}
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __brev(unsigned v)
{
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
static inline long long clock64()
{
    return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <class T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline double atomicAdd(double* p, double v)
{
    unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = __atomic_load_n(q, __ATOMIC_SEQ_CST), want;
    double cur;
    do {
        memcpy(&cur, &old, 8);
        cur += v;
        memcpy(&want, &cur, 8);
    } while (!__atomic_compare_exchange_n(q, &old, want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    memcpy(&cur, &old, 8);
    return cur;
}
template <class T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicMax(T* p, T v)
{
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
using std::max;
using std::min;

// ---------------------------------------------------------------------------------------------- runtime
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97,
       cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static int g_emu_sm_count = 2;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int)
{
    *v = attr == cudaDevAttrMultiProcessorCount ? g_emu_sm_count : 232448;      // 227 KB of shared memory per block
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }

alignas(128) static unsigned char g_emu_dynamic_smem[232448];
#define AT_DYNAMIC_SMEM(type, name, alignment) type* const name = reinterpret_cast<type*>(g_emu_dynamic_smem)

static inline void emu_run_grid(dim3 grid, dim3 block, const std::function<void()>& body)
{
    gridDim = grid;
    blockDim = block;
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            blockIdx = dim3(bx, by, 0);
            emu_block_ctx ctx;
            ctx.block.reset(nthreads);
            ctx.warps = std::vector<emu_barrier>(nwarps);
            for (int w = 0; w < nwarps; ++w) ctx.warps[w].reset(std::min(32, nthreads - 32 * w));
            g_emu_ctx = &ctx;
            std::vector<std::thread> pool;
            pool.reserve(nthreads);
            for (int t = 0; t < nthreads; ++t)
                pool.emplace_back([&, t] {
                    threadIdx = dim3((unsigned)t, 0, 0);
                    t_emu_warp = t >> 5;
                    t_emu_lane = t & 31;
                    body();
                    ctx.warps[t_emu_warp].leave();
                    ctx.block.leave();
                });
            for (auto& th : pool) th.join();
            g_emu_ctx = nullptr;
        }
}

// kernel<<<grid, block, smem, stream>>>(args...)
#define AT_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...)                                     \
    do {                                                                                           \
        (void)(smem);                                                                              \
        (void)(stream);                                                                            \
        emu_run_grid(dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); });                       \
    } while (0)
#define AT_LAUNCH(kernel, grid, block, stream, ...) AT_LAUNCH_SMEM(kernel, grid, block, 0, stream, __VA_ARGS__)
