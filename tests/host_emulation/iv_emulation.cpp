// TEST INFRASTRUCTURE — compiles the DEVICE code of agenttorch_b200/csrc/at_interventions.cu for the host, so that the
// kernel source itself (not a restatement) can be checked against the oracle on a machine without a GPU
// (tests/test_interventions_emulated.py).  Threads run one after the other; one thread per block, so every thread is
// lane 0 of its warp and the warp reduction degenerates to the thread's own counts.  Never linked into the product.
#include <cmath>
#include <cstdint>
#include <cstring>

#define AT_IV_HOST_EMULATION 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shfl_xor_sync(mask, value, offset) 0

struct emu_dim { int x; };
static emu_dim blockIdx, blockDim, gridDim, threadIdx;
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float4 __ldcs(const float4* p) { return *p; }
static inline void __stcs(float4* p, float4 v) { *p = v; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }

#include "../../agenttorch_b200/csrc/at_interventions.cu"

template <class Kernel, class Args>
static void run_grid(Kernel kernel, const Args& a, bool vec, int blocks)
{
    gridDim.x = blocks;
    blockDim.x = 1;
    threadIdx.x = 0;
    for (int b = 0; b < blocks; ++b) {
        blockIdx.x = b;
        kernel(a, vec);
    }
}

extern "C" {
void emu_quarantine(const at_iv_quarantine_args* a, int vec, int blocks) { run_grid(iv_quarantine_kernel, *a, vec != 0, blocks); }
void emu_testing(const at_iv_testing_args* a, int vec, int blocks) { run_grid(iv_testing_kernel, *a, vec != 0, blocks); }
void emu_seirms(const at_iv_seirms_args* a, int vec, int blocks)
{
    run_grid(iv_seirms_kernel, *a, vec != 0, blocks);
    blockIdx.x = 0;
    threadIdx.x = 0;
    iv_seirms_epilogue_kernel(*a);
}
}
