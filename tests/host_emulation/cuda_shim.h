// TEST INFRASTRUCTURE — the few CUDA names the kernels of this package use, defined for the host, so that a .cu file can
// be compiled as plain C++ and its kernels run thread after thread (tests/host_emulation/*.cpp).  One thread per block:
// every thread is lane 0 of its warp and warp reductions degenerate to the thread's own value.  Valid for kernels whose
// threads do not communicate inside a launch except through order-independent atomics.  Never linked into the product.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shfl_xor_sync(mask, value, offset) 0
#define __shfl_down_sync(mask, value, offset) 0

struct emu_dim { int x; };
static emu_dim blockIdx, blockDim, gridDim, threadIdx;

struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct double2 { double x, y; };
struct longlong2 { long long x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline longlong2 make_longlong2(long long x, long long y) { return longlong2{x, y}; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { const unsigned o = *p; *p |= v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }

// ---- runtime calls the C-ABI wrappers make
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 1; return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }

// kernel<<<grid, block, 0, stream>>>(args...)  ->  `grid` single-thread blocks, one after the other
#define AT_LAUNCH(kernel, grid, block, stream, ...)            \
    do {                                                       \
        (void)(block);                                         \
        (void)(stream);                                        \
        gridDim.x = (grid);                                    \
        blockDim.x = 1;                                        \
        threadIdx.x = 0;                                       \
        for (int b_ = 0; b_ < gridDim.x; ++b_) {               \
            blockIdx.x = b_;                                   \
            kernel(__VA_ARGS__);                               \
        }                                                      \
    } while (0)
