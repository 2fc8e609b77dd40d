// at_seirm.cu — sm_100a kernels + C ABI for the SEIRM agent update (see include/at_seirm.h).
//
// One launch of `seirm_step_kernel` advances every owned agent through substep '0'
// (NewTransmission, reference transition.py:131-252) and/or substep '1' (SEIRMProgression,
// seirm_progression/transition.py:90-121) and encodes the infectiousness of the next step.
//
// Data flow per warp task (32 consecutive destination rows, lane <-> row):
//   1. stream the rows' contiguous col_idx range with 128-bit loads (lane <-> 4 consecutive edges),
//      test every source against the "is infectious" bitmap staged in SHARED memory (1 bit per
//      source; global/L1 fallback when it does not fit), and only for hits gather the source's
//      1-byte infectiousness code from the L2-resident code array; hit codes are parked in a
//      per-warp shared-memory byte lane, one byte per edge;
//   2. every lane walks the bytes of ITS row in edge order and accumulates lambda in float64 —
//      the same order as the reference's index_add, no atomics, no per-edge global intermediate;
//   3. epilogue: exposure sampling, stage/timer update, progression, next-step code + bitmap word,
//      block-reduced daily series; the last CTA to finish folds the per-CTA partials in fixed order.
// Memory bound by the col_idx stream (4 B/edge); no tensor cores.

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "at_seirm.h"

#define AT_THREADS 1024
#define AT_WARPS (AT_THREADS / 32)
#define AT_CHUNK 1024                       // edges parked per warp per pass (bytes of smem)
#define AT_MAX_GRID 1024
#define AT_PARTIAL_STRIDE 4                 // doubles per CTA
#define AT_STATIC_SMEM (AT_WARPS * AT_CHUNK + 2 * 256 * 4 + 64)

static thread_local char g_err[512] = "";

static int fail(const char* fmt, const char* detail) {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return 1;
}

// ------------------------------------------------------------------------------------------ noise
// SplitMix64 output number ((step << 40) | agent) of the stream seeded `seed`
// (numpy restatement: oracle/noise.py, torch restatement: agenttorch_b200/noise.py).
__host__ __device__ __forceinline__ uint64_t at_noise_u64(uint64_t seed, uint32_t step, uint64_t agent) {
    uint64_t z = (((uint64_t)step << 40) | agent) + 1ull;
    z = z * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ float at_u_exposure(uint64_t z) { return (float)(z >> 40) * (1.0f / 16777216.0f); }
__host__ __device__ __forceinline__ float at_u_isolation(uint64_t z) {
    return (float)((z >> 8) & 0xFFFFFFull) * (1.0f / 16777216.0f);
}

// --------------------------------------------------------------------------------- typed accessors
template <typename T> __device__ __forceinline__ int load_timer(const void* p, int64_t i);
template <> __device__ __forceinline__ int load_timer<float>(const void* p, int64_t i) { return (int)((const float*)p)[i]; }
template <> __device__ __forceinline__ int load_timer<int16_t>(const void* p, int64_t i) { return (int)((const int16_t*)p)[i]; }
template <typename T> __device__ __forceinline__ void store_timer(void* p, int64_t i, int v);
template <> __device__ __forceinline__ void store_timer<float>(void* p, int64_t i, int v) { ((float*)p)[i] = (float)v; }
template <> __device__ __forceinline__ void store_timer<int16_t>(void* p, int64_t i, int v) { ((int16_t*)p)[i] = (int16_t)v; }
template <typename T> __device__ __forceinline__ int load_age(const void* p, int64_t i);
template <> __device__ __forceinline__ int load_age<float>(const void* p, int64_t i) { return (int)((const float*)p)[i]; }
template <> __device__ __forceinline__ int load_age<uint8_t>(const void* p, int64_t i) { return (int)((const uint8_t*)p)[i]; }

__device__ __forceinline__ int4 ldg_stream_int4(const int32_t* p) {
    int4 v;
    asm volatile("ld.global.cs.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

struct KModel {   // device copy of at_seirm_model (kernel parameter, by value)
    int n_sfs, n_sfi, n_lg, e2i, i2r, inf_time;
    float ibar;
    float sfs[AT_MAX_SFS], sfi[AT_MAX_SFI], lg[AT_MAX_LAMGAMMA];
};

// ------------------------------------------------------------------------------------ step kernel
template <bool SMEM_BM, bool WEIGHTED, typename TimerT, typename AgeT>
__global__ void __launch_bounds__(AT_THREADS, 1)
seirm_step_kernel(const KModel m, const at_seirm_launch a, const int bitmap_words)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [hit lanes: AT_WARPS*AT_CHUNK bytes][A table 256 f32][G table 256 f32][red 64 B][bitmap...]
    uint32_t* s_hit_all = reinterpret_cast<uint32_t*>(smem_raw);
    float* s_A = reinterpret_cast<float*>(smem_raw + AT_WARPS * AT_CHUNK);
    float* s_G = s_A + 256;
    uint32_t* s_bm = reinterpret_cast<uint32_t*>(smem_raw + AT_STATIC_SMEM);
    __shared__ double s_red[AT_WARPS][3];
    __shared__ int s_is_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool do_trans = (a.flags & AT_DO_TRANSMISSION) != 0;
    const bool do_prog = (a.flags & AT_DO_PROGRESSION) != 0;
    const bool do_enc = (a.flags & AT_DO_ENCODE) != 0;
    const bool want_grad = a.dne_dr != nullptr;

    // code -> (SFInfector[int(stage)], lam_gamma[idx]) tables; code = 1 + sidx*K + idx
    if (tid < 256) {
        float av = 0.f, gv = 0.f;
        if (tid >= 1 && tid <= 3 * m.n_lg) {
            int sidx = (tid - 1) / m.n_lg, idx = (tid - 1) % m.n_lg;
            av = sidx < m.n_sfi ? m.sfi[sidx] : 0.f;
            gv = m.lg[idx];
        }
        s_A[tid] = av;
        s_G[tid] = gv;
    }
    if (SMEM_BM && do_trans) {
        const uint4* src = reinterpret_cast<const uint4*>(a.bitmap_in);
        uint4* dst = reinterpret_cast<uint4*>(s_bm);
        const int n4 = bitmap_words >> 2;                 // bitmap buffers are padded to 16 B
        for (int i = tid; i < n4; i += AT_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    uint32_t* s_hit = s_hit_all + warp * (AT_CHUNK / 4);
    const float R = do_trans ? a.r_value[a.r_index] : 0.f;
    const double ibar = (double)m.ibar;
    const int t = a.t;
    const uint64_t seed = a.seed_ptr ? *a.seed_ptr : a.seed;
    const int64_t ntasks = (a.n_rows + 31) >> 5;

    double acc_ne = 0.0, acc_dne = 0.0, acc_dead = 0.0;

    for (int64_t task = (int64_t)blockIdx.x * AT_WARPS + warp; task < ntasks; task += (int64_t)gridDim.x * AT_WARPS) {
        const int64_t row = task * 32 + lane;
        const bool valid = row < a.n_rows;
        float stage = valid ? a.stage[row] : 3.0f;
        const bool susceptible = valid && (stage == 0.0f);
        double lam = 0.0, lam1 = 0.0;

        if (do_trans && __any_sync(0xffffffffu, susceptible)) {
            const int64_t rclamp = valid ? row : a.n_rows;
            const int rp0 = a.row_ptr[rclamp];
            const int rp1 = a.row_ptr[valid ? row + 1 : a.n_rows];
            const int e_begin = __shfl_sync(0xffffffffu, rp0, 0);
            const int e_end = __shfl_sync(0xffffffffu, rp1, 31);
            float rs = 0.f, s_i = 0.f;
            if (susceptible) {
                int ag = load_age<AgeT>(a.age, row);
                ag = ag < 0 ? 0 : (ag >= m.n_sfs ? m.n_sfs - 1 : ag);
                s_i = m.sfs[ag];
                rs = __fmul_rn(R, s_i);
            }
            for (int cb = e_begin & ~3; cb < e_end; cb += AT_CHUNK) {
                const int ce = min(cb + AT_CHUNK, e_end);
                // ---- 1. stream + filter: lane <-> 4 consecutive edges, two 128-bit loads in flight
                for (int k0 = cb; k0 < ce; k0 += 256) {
                    const int ka = k0 + 4 * lane, kb = ka + 128;
                    int4 va = make_int4(0, 0, 0, 0), vb = make_int4(0, 0, 0, 0);
                    if (ka < ce) va = ldg_stream_int4(a.col_idx + ka);
                    if (kb < ce) vb = ldg_stream_int4(a.col_idx + kb);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int k = half ? kb : ka;
                        if (k < ce) {
                            const int4 v = half ? vb : va;
                            const int srcs[4] = {v.x, v.y, v.z, v.w};
                            uint32_t packed = 0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int e = k + j;
                                if (e >= e_begin && e < e_end) {
                                    const uint32_t s = (uint32_t)srcs[j];
                                    const uint32_t word = SMEM_BM ? s_bm[s >> 5] : __ldg(a.bitmap_in + (s >> 5));
                                    if ((word >> (s & 31)) & 1u) packed |= (uint32_t)__ldg(a.code_in + s) << (8 * j);
                                }
                            }
                            s_hit[(k - cb) >> 2] = packed;
                        }
                    }
                }
                __syncwarp();
                // ---- 2. each lane folds the hits of its own row, in edge order, in float64
                const int lo = max(rp0, cb), hi = min(rp1, ce);
                if (susceptible && lo < hi) {
                    const int w0 = (lo - cb) >> 2, w1 = (hi - cb + 3) >> 2;
                    for (int w = w0; w < w1; ++w) {
                        uint32_t v = s_hit[w];
                        if (v == 0u) continue;
                        const int base = cb + 4 * w;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t c = (v >> (8 * j)) & 255u;
                            const int e = base + j;
                            if (c != 0u && e >= lo && e < hi) {
                                // reference product order: (((R*S_i)*A_j)*B_e)*G_j in float32, then / I_bar in float64
                                float term = __fmul_rn(rs, s_A[c]);
                                if (WEIGHTED) term = __fmul_rn(term, a.edge_weight[e]);
                                term = __fmul_rn(term, s_G[c]);
                                lam += (double)term / ibar;
                                if (want_grad) {
                                    float t1 = __fmul_rn(s_i, s_A[c]);
                                    if (WEIGHTED) t1 = __fmul_rn(t1, a.edge_weight[e]);
                                    t1 = __fmul_rn(t1, s_G[c]);
                                    lam1 += (double)t1 / ibar;
                                }
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }

        // ---- 3. per-agent epilogue (lane <-> row)
        float dne = 0.f;
        bool st_changed = false;
        int nst = 0, itime = 0;
        bool nst_loaded = false, nst_dirty = false, itime_loaded = false, itime_dirty = false;
        if (valid) {
            if (do_trans && susceptible && (lam > 0.0 || lam1 > 0.0)) {
                float wiso, u;
                if (a.iso_mode == AT_ISO_TENSOR) wiso = a.isolation[row];
                if (a.u_mode == AT_U_TENSOR) u = a.uniform[row];
                if (a.iso_mode != AT_ISO_TENSOR || a.u_mode != AT_U_TENSOR) {
                    const uint64_t z = at_noise_u64(seed, (uint32_t)t, (uint64_t)(a.row_offset + row));
                    if (a.u_mode != AT_U_TENSOR) u = at_u_exposure(z);
                    if (a.iso_mode == AT_ISO_HASH_SOFT) wiso = at_u_isolation(z);
                    else if (a.iso_mode == AT_ISO_HASH_HARD) wiso = at_u_isolation(z) < a.p_isolate ? 1.0f : 0.0f;
                    else if (a.iso_mode == AT_ISO_NONE) wiso = 0.0f;
                }
                const double q = exp(-lam);                 // prob_not_infected (transition.py:220)
                const double p = 1.0 - q;
                const float omw = 1.0f - wiso;              // float32, as (1.0 - will_isolate)
                const float ne = ((double)u < p) ? omw : 0.0f;
                if (want_grad) dne = (float)((double)omw * q * lam1);
                if (ne != 0.0f) {
                    stage = ne;                              // 0 + ne
                    st_changed = true;
                    nst = t + 1 + m.e2i; nst_loaded = nst_dirty = true;
                    itime = t; itime_loaded = itime_dirty = true;
                    acc_ne += (double)ne;
                }
                acc_dne += (double)dne;
            }
            if (do_prog) {
                const bool is_e = (stage == 1.0f), is_i = (stage == 2.0f);
                if (is_e || is_i) {
                    if (!nst_loaded) { nst = load_timer<TimerT>(a.next_stage_time, row); nst_loaded = true; }
                    if (nst <= t) {
                        if (is_i) {
                            acc_dead += (double)stage / 2.0;
                            if (a.recovered_step) a.recovered_step[row] = (uint8_t)t;
                        }
                        if (nst == t) { nst = is_i ? m.inf_time : t + m.i2r; nst_dirty = true; }
                        stage += 1.0f;
                        st_changed = true;
                    }
                }
            }
            if (st_changed) a.stage[row] = stage;
            if (nst_dirty) store_timer<TimerT>(a.next_stage_time, row, nst);
            if (itime_dirty) store_timer<TimerT>(a.infected_time, row, itime);
            if (want_grad) a.dne_dr[row] = dne;
        }
        if (do_enc) {
            uint32_t code = 0;
            if (valid && stage > 0.0f && stage < 3.0f) {
                if (!itime_loaded) itime = load_timer<TimerT>(a.infected_time, row);
                const int sidx = (int)stage;
                int idx = a.t_encode - itime - 1;
                if (idx < 0 || idx >= m.n_lg) {
                    if (a.error_flag && sidx < m.n_sfi && m.sfi[sidx] != 0.f) *a.error_flag = 1;
                    idx = idx < 0 ? 0 : m.n_lg - 1;
                }
                code = 1u + (uint32_t)(sidx * m.n_lg + idx);
                if (s_A[code] == 0.f || s_G[code] == 0.f) code = 0;   // contributes exactly 0 to every lambda
            }
            const uint32_t bits = __ballot_sync(0xffffffffu, code != 0u);
            if (valid) a.code_out[a.row_offset + row] = (uint8_t)code;
            if (lane == 0) a.bitmap_out[(a.row_offset >> 5) + task] = bits;
        }
    }

    // ---- deterministic reduction: warp shuffle -> smem -> per-CTA partial -> last CTA folds all
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_ne += __shfl_down_sync(0xffffffffu, acc_ne, o);
        acc_dne += __shfl_down_sync(0xffffffffu, acc_dne, o);
        acc_dead += __shfl_down_sync(0xffffffffu, acc_dead, o);
    }
    if (lane == 0) { s_red[warp][0] = acc_ne; s_red[warp][1] = acc_dne; s_red[warp][2] = acc_dead; }
    __syncthreads();
    int* ticket = reinterpret_cast<int*>(a.partials + (size_t)AT_MAX_GRID * AT_PARTIAL_STRIDE);
    if (tid == 0) {
        double s0 = 0, s1 = 0, s2 = 0;
        for (int w = 0; w < AT_WARPS; ++w) { s0 += s_red[w][0]; s1 += s_red[w][1]; s2 += s_red[w][2]; }
        double* mine = a.partials + (size_t)blockIdx.x * AT_PARTIAL_STRIDE;
        mine[0] = s0; mine[1] = s1; mine[2] = s2;
        __threadfence();
        s_is_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_is_last && tid == 0) {
        __threadfence();
        double s0 = 0, s1 = 0, s2 = 0;
        for (unsigned b = 0; b < gridDim.x; ++b) {
            const volatile double* p = a.partials + (size_t)b * AT_PARTIAL_STRIDE;
            s0 += p[0]; s1 += p[1]; s2 += p[2];
        }
        if (do_trans) {
            if (a.daily_infected) a.daily_infected[t] += (float)s0;          // transition.py:235
            if (a.sum_dne_dr) a.sum_dne_dr[t] = (float)s1;
        }
        if (do_prog) {
            if (a.death_count) a.death_count[t] = (float)s2;
            if (a.daily_deaths) a.daily_deaths[t] += (float)s2 * a.m_value[0];   // progression :52-60
        }
        *ticket = 0;
    }
}

// -------------------------------------------------------------------------------- backward kernel
// grad_r[t] = gI[t]*sum_dne[t] + sum_i dne[t,i] * c_i,  c_i = gD[recovered_step_i] * M / 2
#define AT_BWD_THREADS 256
#define AT_BWD_MAXT 64
__global__ void __launch_bounds__(AT_BWD_THREADS)
seirm_episode_bwd_kernel(int T, int64_t n, int64_t ld, const float* __restrict__ dne, const uint8_t* __restrict__ rstep,
                         const float* __restrict__ sum_dne, const float* __restrict__ death_count,
                         const float* __restrict__ gI, const float* __restrict__ gD, const float* __restrict__ m_value,
                         float* grad_r, float* grad_m, double* partials)
{
    __shared__ float s_gd[256];
    __shared__ double s_red[AT_BWD_THREADS / 32];
    __shared__ double s_tot[AT_BWD_MAXT];
    __shared__ int s_is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float half_m = m_value[0] * 0.5f;
    if (tid < 256) s_gd[tid] = (tid < T) ? gD[tid] * half_m : 0.f;
    if (tid < AT_BWD_MAXT) s_tot[tid] = 0.0;
    __syncthreads();
    // each thread owns a strided set of agents; the per-step sums are kept per tile to bound registers
    const int64_t tile = (int64_t)AT_BWD_THREADS * 4;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n; base += (int64_t)gridDim.x * tile) {
        float c[4];
        bool any = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = base + j * AT_BWD_THREADS + tid;
            c[j] = (i < n) ? s_gd[rstep[i]] : 0.f;
            any |= (c[j] != 0.f);
        }
        const bool warp_any = __any_sync(0xffffffffu, any);
        for (int t = 0; t < T; ++t) {
            double s = 0.0;
            if (warp_any) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t i = base + j * AT_BWD_THREADS + tid;
                    if (i < n && c[j] != 0.f) s += (double)(dne[(int64_t)t * ld + i] * c[j]);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            }
            if (lane == 0) s_red[warp] = s;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
                for (int w = 0; w < AT_BWD_THREADS / 32; ++w) tot += s_red[w];
                s_tot[t] += tot;
            }
            __syncthreads();
        }
    }
    int* ticket = reinterpret_cast<int*>(partials + (size_t)AT_MAX_GRID * AT_BWD_MAXT);
    if (tid < T) partials[(size_t)blockIdx.x * AT_BWD_MAXT + tid] = s_tot[tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) s_is_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_is_last) {
        __threadfence();
        if (tid < T) {
            double tot = 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) tot += ((const volatile double*)partials)[(size_t)b * AT_BWD_MAXT + tid];
            grad_r[tid] = (float)((double)gI[tid] * (double)sum_dne[tid] + tot);
        }
        if (tid == 0) {
            double gm = 0.0;
            for (int t = 0; t < T; ++t) gm += (double)gD[t] * (double)death_count[t];
            grad_m[0] = (float)gm;
            *ticket = 0;
        }
    }
}

__global__ void noise_fill_kernel(float* out, int64_t n, int64_t id_offset, uint64_t seed, int step, int what, float p) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t z = at_noise_u64(seed, (uint32_t)step, (uint64_t)(id_offset + i));
        float v;
        if (what == 0) v = at_u_exposure(z);
        else if (what == 1) v = at_u_isolation(z);
        else v = at_u_isolation(z) < p ? 1.0f : 0.0f;
        out[i] = v;
    }
}

// ------------------------------------------------------------------------------------------ C ABI
static int g_sm_count = 0, g_smem_optin = 0;
static int device_props() {
    if (g_sm_count) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    if (cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
    if (cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 1;
    return 0;
}

template <bool SMEM_BM, bool WEIGHTED, typename TimerT, typename AgeT>
static int launch_step(const KModel& km, const at_seirm_launch& a, int grid, size_t smem, int bitmap_words, cudaStream_t s) {
    auto k = seirm_step_kernel<SMEM_BM, WEIGHTED, TimerT, AgeT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    k<<<grid, AT_THREADS, smem, s>>>(km, a, bitmap_words);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail("seirm_step_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

template <bool SMEM_BM, bool WEIGHTED>
static int dispatch_types(const KModel& km, const at_seirm_launch& a, int grid, size_t smem, int words, cudaStream_t s) {
    if (a.timer_dtype == AT_TIMER_F32 && a.age_dtype == AT_AGE_F32) return launch_step<SMEM_BM, WEIGHTED, float, float>(km, a, grid, smem, words, s);
    if (a.timer_dtype == AT_TIMER_F32 && a.age_dtype == AT_AGE_U8) return launch_step<SMEM_BM, WEIGHTED, float, uint8_t>(km, a, grid, smem, words, s);
    if (a.timer_dtype == AT_TIMER_I16 && a.age_dtype == AT_AGE_F32) return launch_step<SMEM_BM, WEIGHTED, int16_t, float>(km, a, grid, smem, words, s);
    if (a.timer_dtype == AT_TIMER_I16 && a.age_dtype == AT_AGE_U8) return launch_step<SMEM_BM, WEIGHTED, int16_t, uint8_t>(km, a, grid, smem, words, s);
    return fail("%s", "unsupported timer_dtype / age_dtype");
}

extern "C" {

int at_seirm_abi_version(void) { return AT_SEIRM_ABI_VERSION; }
int64_t at_seirm_sizeof_launch(void) { return (int64_t)sizeof(at_seirm_launch); }
int64_t at_seirm_sizeof_model(void) { return (int64_t)sizeof(at_seirm_model); }
const char* at_last_error(void) { return g_err; }
int64_t at_seirm_workspace_bytes(void) {
    return (int64_t)AT_MAX_GRID * AT_BWD_MAXT * sizeof(double) + 64;   // covers both kernels' layouts
}
int64_t at_seirm_smem_bitmap_capacity(void) {
    if (device_props()) return 0;
    int64_t bytes = (int64_t)g_smem_optin - AT_STATIC_SMEM - 1024;     // 1 KB for the static __shared__ arrays
    if (bytes < 0) return 0;
    return (bytes & ~15ll) * 8;
}

int at_seirm_step(const at_seirm_model* model, const at_seirm_launch* launch, void* stream) {
    if (!model || !launch) return fail("%s", "null argument");
    const at_seirm_launch& a = *launch;
    if (a.abi_version != AT_SEIRM_ABI_VERSION) return fail("%s", "at_seirm_launch.abi_version mismatch");
    if (model->n_lam_gamma < 1 || model->n_lam_gamma > AT_MAX_LAMGAMMA) return fail("%s", "n_lam_gamma out of range (1..84)");
    if (model->n_sfs < 1 || model->n_sfs > AT_MAX_SFS || model->n_sfi < 1 || model->n_sfi > AT_MAX_SFI) return fail("%s", "table size out of range");
    if (a.n_rows < 0 || (a.row_offset & 31) != 0) return fail("%s", "row_offset must be a multiple of 32");
    if (!a.stage || !a.partials) return fail("%s", "stage / partials must not be null");
    const bool do_trans = a.flags & AT_DO_TRANSMISSION;
    if (do_trans && (!a.row_ptr || !a.col_idx || !a.code_in || !a.bitmap_in || !a.r_value || !a.age)) return fail("%s", "transmission needs CSR, code_in, bitmap_in, age, r_value");
    if (do_trans && a.iso_mode == AT_ISO_TENSOR && !a.isolation) return fail("%s", "iso_mode TENSOR without isolation pointer");
    if (do_trans && a.u_mode == AT_U_TENSOR && !a.uniform) return fail("%s", "u_mode TENSOR without uniform pointer");
    if ((a.flags & AT_DO_PROGRESSION) && (!a.next_stage_time || (a.daily_deaths && !a.m_value))) return fail("%s", "progression needs next_stage_time and m_value");
    if ((a.flags & AT_DO_ENCODE) && (!a.code_out || !a.bitmap_out || !a.infected_time)) return fail("%s", "encode needs code_out, bitmap_out, infected_time");
    if (a.t < 0 || a.t > 255) return fail("%s", "t out of range (0..255)");
    if (device_props()) return fail("%s", "cannot query device");
    if (a.n_rows == 0) return 0;

    KModel km;
    km.n_sfs = model->n_sfs; km.n_sfi = model->n_sfi; km.n_lg = model->n_lam_gamma;
    km.e2i = model->exposed_to_infected_time; km.i2r = model->infected_to_recovered_time; km.inf_time = model->infinity_time;
    km.ibar = model->i_bar;
    memcpy(km.sfs, model->sfs, sizeof(km.sfs)); memcpy(km.sfi, model->sfi, sizeof(km.sfi)); memcpy(km.lg, model->lam_gamma, sizeof(km.lg));

    const int64_t words64 = ((a.n_sources + 31) / 32 + 3) & ~3ll;   // bitmap buffers are padded to 16 B by the caller
    const int64_t cap = at_seirm_smem_bitmap_capacity();
    bool smem_bm = do_trans && a.n_sources <= cap;
    if (a.use_smem_bitmap == 0) smem_bm = false;
    if (a.use_smem_bitmap == 1 && do_trans && a.n_sources > cap) return fail("%s", "n_sources exceeds the shared-memory bitmap capacity");
    const int bitmap_words = (int)words64;
    const size_t smem = AT_STATIC_SMEM + (smem_bm ? (size_t)bitmap_words * 4 : 0);
    int grid = a.grid > 0 ? a.grid : g_sm_count;
    const int64_t ntasks = (a.n_rows + 31) / 32;
    const int64_t max_useful = (ntasks + AT_WARPS - 1) / AT_WARPS;
    if (grid > max_useful) grid = (int)max_useful;
    if (grid > AT_MAX_GRID) grid = AT_MAX_GRID;
    if (grid < 1) grid = 1;
    cudaStream_t s = (cudaStream_t)stream;
    const bool weighted = a.edge_weight != nullptr;
    if (smem_bm) return weighted ? dispatch_types<true, true>(km, a, grid, smem, bitmap_words, s) : dispatch_types<true, false>(km, a, grid, smem, bitmap_words, s);
    return weighted ? dispatch_types<false, true>(km, a, grid, smem, bitmap_words, s) : dispatch_types<false, false>(km, a, grid, smem, bitmap_words, s);
}

int at_seirm_episode_backward(int32_t num_steps, int64_t n_rows, int64_t ld, const float* dne_dr, const uint8_t* recovered_step,
                              const float* sum_dne_dr, const float* death_count, const float* g_daily_infected,
                              const float* g_daily_deaths, const float* m_value, float* grad_r, float* grad_m,
                              double* partials, int32_t grid, void* stream) {
    if (num_steps < 1 || num_steps > AT_BWD_MAXT) return fail("%s", "num_steps out of range (1..64)");
    if (!dne_dr || !recovered_step || !sum_dne_dr || !death_count || !g_daily_infected || !g_daily_deaths || !m_value || !grad_r || !grad_m || !partials)
        return fail("%s", "null argument");
    if (device_props()) return fail("%s", "cannot query device");
    int g = grid > 0 ? grid : g_sm_count * 4;
    const int64_t tiles = (n_rows + AT_BWD_THREADS * 4 - 1) / (AT_BWD_THREADS * 4);
    if (g > tiles) g = (int)tiles;
    if (g > AT_MAX_GRID) g = AT_MAX_GRID;
    if (g < 1) g = 1;
    seirm_episode_bwd_kernel<<<g, AT_BWD_THREADS, 0, (cudaStream_t)stream>>>(num_steps, n_rows, ld, dne_dr, recovered_step, sum_dne_dr,
        death_count, g_daily_infected, g_daily_deaths, m_value, grad_r, grad_m, partials);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("seirm_episode_bwd_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

int at_noise_fill(float* out, int64_t n, int64_t id_offset, uint64_t seed, int32_t step, int32_t what, float p_isolate, void* stream) {
    if (!out && n > 0) return fail("%s", "null output");
    if (n <= 0) return 0;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    noise_fill_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, n, id_offset, seed, step, what, p_isolate);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("noise_fill_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
