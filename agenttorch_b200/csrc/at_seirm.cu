// at_seirm.cu — sm_100a kernels + C ABI for the SEIRM agent update (see include/at_seirm.h).
//
// seirm_kernel<MODE> advances the agents through substep '0' (NewTransmission, reference
// transition.py:131-252) and/or substep '1' (SEIRMProgression, seirm_progression/transition.py:90-121)
// and encodes the infectiousness of the next step.
//
// Work unit: a TASK of 32 consecutive destination rows (lane <-> row), scheduled dynamically in groups.
//   gather   - the task's edge segment (column-blocked CSR, 4 B/edge, the only large stream) is staged in
//              shared memory by TMA bulk copies (cp.async.bulk + mbarrier), double-buffered per warp, the next
//              chunk always in flight.  Every source id is tested against the "hot" bitmap (1 bit per source:
//              infectiousness code != 0) held in SHARED memory; only hits gather the source's 1-byte code from
//              the L2-resident code array and are appended, in edge order, to a per-warp hit list.
//   fold     - each lane binary-searches the hit list for ITS row and adds pre-divided float64 terms
//              table[age][code] = f64(f32(((R*S)*A)*G)) / I_bar in edge order: the reference's message, its
//              float32 product order and its index_add order; no atomics, no per-edge global intermediate.
//   update   - exposure test (float32 fast path, exact float64 decision near the threshold), stage/timer
//              writes only where something changes, progression, next-step code + bitmaps.
// Sources beyond the shared-memory bitmap capacity are handled by column blocks: MODE_GATHER runs all blocks
// concurrently (CTA c serves block c % n_blocks) and leaves sparse per-block partial lambdas, MODE_UPDATE folds
// them in block order.  Daily sums use 64-bit fixed-point integer accumulation, so they are bitwise
// reproducible under dynamic scheduling.  HBM-bound by the col stream; no tensor cores.

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "at_seirm.h"

#define NT 512                      // threads per CTA
#define NW (NT / 32)                // warps per CTA
#define CH 512                      // edges per staged chunk (2 KB)
#define LIST_CAP 256                // hit-list entries per warp
#define MAXG 8                      // max tasks per scheduling group
#define WARP_SMEM (2 * CH * 4 + LIST_CAP * 4 + 64)
#define FX_SCALE 4294967296.0       // 2^32 fixed point for the daily sums

#define MODE_FUSED 0                // n_blocks == 1: gather + update per task
#define MODE_GATHER 1               // n_blocks > 1: partial lambdas of every column block
#define MODE_UPDATE 2               // update only (lambda from partials, or no transmission)

// workspace layout (bytes)
#define WS_ACC 0                    // int64[3]: sum ne, sum dne, deaths (fixed point)
#define WS_TICKET 32                // int: CTAs finished (main kernel)
#define WS_COUNTER 36               // int: group counter (main kernel)
#define WS_GCOUNTERS 64             // int[AT_MAX_BLOCKS]: group counters of the gather kernel
#define WS_GTICKET 384              // int
#define WS_BTICKET 512              // int: backward kernel
#define WS_BPART 1024               // double[BWD_MAX_GRID][BWD_MAXT]
#define BWD_MAX_GRID 1024
#define BWD_MAXT 64
#define WS_BYTES (WS_BPART + BWD_MAX_GRID * BWD_MAXT * 8)

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

// ------------------------------------------------------------------------------- TMA / mbarrier PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

struct KModel {   // device copy of at_seirm_model (kernel parameter, by value)
    int n_sfs, n_sfi, n_lg, e2i, i2r, inf_time;
    float ibar;
    float sfs[AT_MAX_SFS], sfi[AT_MAX_SFI], lg[AT_MAX_LAMGAMMA];
};

// -------------------------------------------------------------------------------------- the kernel
template <int MODE, bool WEIGHTED, typename TimerT, typename AgeT>
__global__ void __launch_bounds__(NT, 1)
seirm_kernel(const KModel m, const at_seirm_launch a, const int bm_words, const int NC, const int G)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // [NW per-warp areas][Tdiv f64 n_sfs*NC][T1div f64 n_sfs*NC][s_A f32 256][s_G f32 256][hot bitmap]
    double* s_T = reinterpret_cast<double*>(smem_raw + NW * WARP_SMEM);
    double* s_T1 = s_T + m.n_sfs * NC;
    float* s_A = reinterpret_cast<float*>(s_T1 + m.n_sfs * NC);
    float* s_G = s_A + 256;
    uint32_t* s_bm = reinterpret_cast<uint32_t*>(s_G + 256);
    __shared__ long long s_red[NW][3];
    __shared__ int s_is_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool do_trans = (a.flags & AT_DO_TRANSMISSION) != 0;
    const bool do_prog = (a.flags & AT_DO_PROGRESSION) != 0;
    const bool do_enc = (a.flags & AT_DO_ENCODE) != 0;
    const bool do_init = (a.flags & AT_DO_INIT) != 0;
    const bool want_grad = a.dne_dr != nullptr;
    const bool has_gather = (MODE != MODE_UPDATE) && do_trans;
    const int t = a.t;
    const int block = (MODE == MODE_GATHER) ? (int)(blockIdx.x % a.n_blocks) : 0;
    const float R = do_trans ? a.r_value[a.r_index] : 0.f;
    const double ibar = (double)m.ibar;
    char* ws = reinterpret_cast<char*>(a.workspace);

    // ---- tables: code -> (SFInfector[int(stage)], lam_gamma[idx]); code = 1 + sidx*K + idx
    if (tid < 256) {
        float av = 0.f, gv = 0.f;
        if (tid >= 1 && tid <= 3 * m.n_lg) {
            const int sidx = (tid - 1) / m.n_lg, idx = (tid - 1) % m.n_lg;
            av = sidx < m.n_sfi ? m.sfi[sidx] : 0.f;
            gv = m.lg[idx];
        }
        s_A[tid] = av;
        s_G[tid] = gv;
    }
    __syncthreads();
    if (has_gather) {
        // message of an unweighted edge, exactly as the reference forms it (transition.py:78-83):
        // float32 product chain ((R*S_i)*A_j)*G_j (B_n == 1 is exact), then float64 division by I_bar
        for (int i = tid; i < m.n_sfs * NC; i += NT) {
            const int ag = i / NC, c = i % NC;
            const float s = m.sfs[ag];
            const float av = c < 256 ? s_A[c] : 0.f, gv = c < 256 ? s_G[c] : 0.f;
            s_T[i] = (double)__fmul_rn(__fmul_rn(__fmul_rn(R, s), av), gv) / ibar;
            s_T1[i] = (double)__fmul_rn(__fmul_rn(s, av), gv) / ibar;
        }
        const uint4* src = reinterpret_cast<const uint4*>(a.hot_bitmap_in + (size_t)block * (a.block_size >> 5));
        uint4* dst = reinterpret_cast<uint4*>(s_bm);
        for (int i = tid; i < (bm_words >> 2); i += NT) dst[i] = __ldg(src + i);
    }
    unsigned char* wsm = smem_raw + warp * WARP_SMEM;
    int32_t* s_col = reinterpret_cast<int32_t*>(wsm);
    uint32_t* s_list = reinterpret_cast<uint32_t*>(wsm + 2 * CH * 4);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(wsm + 2 * CH * 4 + LIST_CAP * 4);
    if (lane == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int64_t ntasks = a.n_rows_pad >> 5;
    const int64_t ngroups = (ntasks + G - 1) / G;
    int* counter = reinterpret_cast<int*>(ws + (MODE == MODE_GATHER ? WS_GCOUNTERS + 4 * block : WS_COUNTER));
    const uint16_t* deg_b = a.deg ? a.deg + (size_t)block * a.n_rows_pad : nullptr;
    const int32_t* segb = a.seg_base4 ? a.seg_base4 + (size_t)block * ntasks : nullptr;
    const uint8_t* code_b = a.code_in ? a.code_in + (size_t)block * a.block_size : nullptr;
    const uint64_t seed = a.seed_ptr ? *a.seed_ptr : a.seed;

    long long acc_ne = 0, acc_dne = 0, acc_dead = 0;
    uint32_t issued = 0, consumed = 0;      // chunks issued / consumed by this warp since kernel start

    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(counter, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= ngroups) break;
        const int64_t task0 = (int64_t)g * G;
        const int nt = (int)min((int64_t)G, ntasks - task0);

        // ---- group meta: lane i holds the susceptible word, segment base and edge total of task i
        uint32_t my_sus = 0;
        int my_base4 = 0, my_tot = 0;
        if (has_gather) {
            if (lane < nt) {
                my_sus = a.sus_bitmap[task0 + lane];
                my_base4 = segb[task0 + lane];
            }
#pragma unroll
            for (int i = 0; i < MAXG; ++i) {
                if (i < nt) {
                    const int d = (int)deg_b[(task0 + i) * 32 + lane];
                    const int tot = __reduce_add_sync(0xffffffffu, d);
                    if (lane == i) my_tot = tot;
                }
            }
            if (my_sus == 0) my_tot = 0;        // no susceptible row: nothing to gather for this task
        }
        // prefetch pointer (task index in group, chunk index)
        int pf_ti = 0, pf_ci = 0;
        auto pf_skip_empty = [&]() {
            while (pf_ti < nt && __shfl_sync(0xffffffffu, my_tot, pf_ti) == 0) ++pf_ti;
        };
        auto pf_issue = [&]() {               // issue chunk (pf_ti, pf_ci) into slot issued&1, then advance
            const int tot = __shfl_sync(0xffffffffu, my_tot, pf_ti);
            const int base4 = __shfl_sync(0xffffffffu, my_base4, pf_ti);
            const int off = pf_ci * CH;
            const int n = min(CH, tot - off);
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)((n + 3) & ~3) * 4u;
                uint64_t* bar = &s_bar[issued & 1];
                mbar_expect_tx(bar, bytes);
                bulk_g2s(s_col + (issued & 1) * CH, a.col + ((int64_t)base4 * 4 + off), bytes, bar);
            }
            ++issued;
            if (off + CH < tot) ++pf_ci; else { ++pf_ti; pf_ci = 0; pf_skip_empty(); }
        };
        if (has_gather) {
            pf_skip_empty();
            if (pf_ti < nt) pf_issue();
        }

        for (int ti = 0; ti < nt; ++ti) {
            const int64_t task = task0 + ti;
            const int64_t row = task * 32 + lane;
            const bool valid = row < a.n_rows;
            double lam = 0.0, lam1 = 0.0;
            uint32_t sus_w = 0;
            int ag = -1;

            // ================================================================== gather + fold
            const int tot = has_gather ? __shfl_sync(0xffffffffu, my_tot, ti) : 0;
            if (has_gather) sus_w = __shfl_sync(0xffffffffu, my_sus, ti);
            if (tot > 0) {
                const int base4 = __shfl_sync(0xffffffffu, my_base4, ti);
                const int d = (int)deg_b[task * 32 + lane];
                int incl = d;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const int start = incl - d;
                const bool my_sus_bit = (sus_w >> lane) & 1u;
                int cnt = 0;

                auto fold = [&]() {
                    // entries are (edge offset << 8 | code), sorted by offset; mine are [start, start+d)
                    if (my_sus_bit && d > 0) {
                        int pos = 0;
#pragma unroll
                        for (int step = LIST_CAP / 2; step >= 1; step >>= 1) {
                            const int p = pos + step;
                            if (p <= cnt && (int)(s_list[p - 1] >> 8) < start) pos = p;
                        }
                        if (pos < cnt && ag < 0) {
                            ag = load_age<AgeT>(a.age, row);
                            ag = ag < 0 ? 0 : (ag >= m.n_sfs ? m.n_sfs - 1 : ag);
                        }
                        const int end = start + d;
                        for (; pos < cnt; ++pos) {
                            const uint32_t e = s_list[pos];
                            const int off = (int)(e >> 8);
                            if (off >= end) break;
                            const int c = (int)(e & 255u);
                            if (!WEIGHTED) {
                                lam += s_T[ag * NC + c];
                                if (want_grad) lam1 += s_T1[ag * NC + c];
                            } else {
                                const float w = a.edge_weight[(int64_t)base4 * 4 + off];
                                const float s = m.sfs[ag];
                                lam += (double)__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(R, s), s_A[c]), w), s_G[c]) / ibar;
                                if (want_grad) lam1 += (double)__fmul_rn(__fmul_rn(__fmul_rn(s, s_A[c]), w), s_G[c]) / ibar;
                            }
                        }
                    }
                    __syncwarp();
                    cnt = 0;
                };

                const int nch = (tot + CH - 1) / CH;
                for (int ci = 0; ci < nch; ++ci) {
                    if (pf_ti < nt) pf_issue();                         // keep the next chunk in flight
                    const uint32_t slot = consumed & 1u;
                    mbar_wait(&s_bar[slot], (consumed >> 1) & 1u);
                    ++consumed;
                    const int n = min(CH, tot - ci * CH);
                    const int4* sc = reinterpret_cast<const int4*>(s_col + slot * CH);
                    for (int mth = 0; mth * 128 < n; ++mth) {
                        const int i4 = mth * 32 + lane;
                        uint32_t h = 0;
                        int4 c = make_int4(0, 0, 0, 0);
                        if (i4 * 4 < n) {
                            c = sc[i4];
                            const uint32_t w0 = s_bm[(uint32_t)c.x >> 5], w1 = s_bm[(uint32_t)c.y >> 5];
                            const uint32_t w2 = s_bm[(uint32_t)c.z >> 5], w3 = s_bm[(uint32_t)c.w >> 5];
                            h = ((w0 >> (c.x & 31)) & 1u) | (((w1 >> (c.y & 31)) & 1u) << 1) |
                                (((w2 >> (c.z & 31)) & 1u) << 2) | (((w3 >> (c.w & 31)) & 1u) << 3);
                        }
                        if (__any_sync(0xffffffffu, h != 0u)) {
                            if (cnt > LIST_CAP - 128) fold();
                            const int k = __popc(h);
                            const uint32_t b0 = __ballot_sync(0xffffffffu, k & 1), b1 = __ballot_sync(0xffffffffu, k & 2),
                                           b2 = __ballot_sync(0xffffffffu, k & 4);
                            const uint32_t lt = (1u << lane) - 1u;
                            int pos = cnt + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
                            const uint32_t offb = (uint32_t)(ci * CH + i4 * 4) << 8;
                            if (h & 1u) s_list[pos++] = offb | __ldg(code_b + c.x);
                            if (h & 2u) s_list[pos++] = (offb + (1u << 8)) | __ldg(code_b + c.y);
                            if (h & 4u) s_list[pos++] = (offb + (2u << 8)) | __ldg(code_b + c.z);
                            if (h & 8u) s_list[pos++] = (offb + (3u << 8)) | __ldg(code_b + c.w);
                            cnt += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
                            __syncwarp();
                        }
                    }
                    __syncwarp();                                        // slot may be refilled two chunks later
                    if (cnt > 0 && (ci == nch - 1 || cnt > LIST_CAP - 128)) fold();
                }
            }

            if (MODE == MODE_GATHER) {
                const bool nz = (lam != 0.0) || (lam1 != 0.0);
                const uint32_t dw = __ballot_sync(0xffffffffu, nz);
                if (nz) {
                    a.lam_part[(size_t)block * a.n_rows_pad + row] = lam;
                    if (want_grad) a.lam1_part[(size_t)block * a.n_rows_pad + row] = lam1;
                }
                if (lane == 0) a.dirty[(size_t)block * ntasks + task] = dw;
                continue;
            }

            // ================================================================== update (lane <-> row)
            uint32_t exp_w = 0, inf_w = 0;
            float stage = 0.f;
            bool stage_known = false;
            if (do_init) {
                stage = valid ? a.stage[row] : 3.0f;
                stage_known = true;
                sus_w = __ballot_sync(0xffffffffu, stage == 0.0f);
                exp_w = __ballot_sync(0xffffffffu, stage == 1.0f);
                inf_w = __ballot_sync(0xffffffffu, stage == 2.0f);
            } else {
                if (!has_gather) sus_w = a.sus_bitmap[task];
                exp_w = a.exp_bitmap[task];
                inf_w = a.inf_bitmap[task];
            }
            bool is_s = (sus_w >> lane) & 1u, is_e = (exp_w >> lane) & 1u, is_i = (inf_w >> lane) & 1u;
            if (MODE == MODE_UPDATE && do_trans && a.dirty != nullptr && sus_w != 0u) {
                // fold the per-block partial lambdas in block order
                for (int b0 = 0; b0 < a.n_blocks; b0 += 32) {
                    uint32_t dwl = 0;
                    if (b0 + lane < a.n_blocks) dwl = a.dirty[(size_t)(b0 + lane) * ntasks + task];
                    uint32_t any_b = __ballot_sync(0xffffffffu, dwl != 0u);
                    while (any_b) {
                        const int bl = __ffs(any_b) - 1;
                        any_b &= any_b - 1;
                        const uint32_t dw = __shfl_sync(0xffffffffu, dwl, bl);
                        if ((dw >> lane) & 1u) {
                            lam += a.lam_part[(size_t)(b0 + bl) * a.n_rows_pad + row];
                            if (want_grad) lam1 += a.lam1_part[(size_t)(b0 + bl) * a.n_rows_pad + row];
                        }
                    }
                }
            }

            float dne = 0.f;
            int nst = 0, itime = 0;
            bool nst_known = false, itime_known = false;
            if (do_trans && is_s && (lam > 0.0 || lam1 > 0.0)) {
                float wiso = 0.f, u = 0.f;
                if (a.iso_mode == AT_ISO_TENSOR) wiso = a.isolation[row];
                if (a.u_mode == AT_U_TENSOR) u = a.uniform[row];
                if (a.iso_mode != AT_ISO_TENSOR || a.u_mode != AT_U_TENSOR) {
                    const uint64_t z = at_noise_u64(seed, (uint32_t)t, (uint64_t)(a.row_offset + row));
                    if (a.u_mode != AT_U_TENSOR) u = at_u_exposure(z);
                    if (a.iso_mode == AT_ISO_HASH_SOFT) wiso = at_u_isolation(z);
                    else if (a.iso_mode == AT_ISO_HASH_HARD) wiso = at_u_isolation(z) < a.p_isolate ? 1.0f : 0.0f;
                }
                // p = 1 - exp(-lambda) (transition.py:220-221); u < p decided in float32 when it is far from the
                // threshold (|error of p32| < 4e-7), in float64 otherwise: same decisions as the float64 reference
                const float q32 = expf(-(float)lam);
                const float p32 = 1.0f - q32;
                bool hit;
                if (fabsf(u - p32) > 1e-6f) hit = u < p32;
                else hit = (double)u < 1.0 - exp(-lam);
                const float omw = 1.0f - wiso;                  // float32, as (1.0 - will_isolate)
                const float ne = hit ? omw : 0.0f;
                if (want_grad) { dne = omw * q32 * (float)lam1; acc_dne += __double2ll_rn((double)dne * FX_SCALE); }
                if (ne != 0.0f) {
                    stage = ne; stage_known = true;            // 0 + ne
                    a.stage[row] = stage;
                    nst = t + 1 + m.e2i; nst_known = true;
                    itime = t; itime_known = true;
                    store_timer<TimerT>(a.next_stage_time, row, nst);
                    store_timer<TimerT>(a.infected_time, row, itime);
                    acc_ne += __double2ll_rn((double)ne * FX_SCALE);
                    is_s = false;
                    is_e = (stage == 1.0f);
                }
            }
            if (do_prog && (is_e || is_i)) {
                if (!nst_known) { nst = load_timer<TimerT>(a.next_stage_time, row); nst_known = true; }
                if (nst <= t) {
                    if (is_i) {
                        acc_dead += 1;                         // stage * rd / INFECTED_VAR = 2/2
                        if (a.recovered_step) a.recovered_step[row] = (uint8_t)t;
                    }
                    if (nst == t) store_timer<TimerT>(a.next_stage_time, row, is_i ? m.inf_time : t + m.i2r);
                    stage = is_i ? 3.0f : 2.0f;
                    stage_known = true;
                    a.stage[row] = stage;
                    const bool was_e = is_e;
                    is_e = false;
                    is_i = was_e;
                }
            }
            if (want_grad && valid && do_trans) a.dne_dr[row] = dne;

            uint32_t code = 0;
            if (do_enc) {
                // infectious = 0 < stage < 3; fractional stages (soft isolation, Q4) index SFInfector[0]
                const bool frac = stage_known && stage > 0.0f && stage < 3.0f && !(is_e || is_i);
                if (is_e || is_i || frac) {
                    if (!itime_known) itime = load_timer<TimerT>(a.infected_time, row);
                    const int sidx = is_e ? 1 : (is_i ? 2 : (int)stage);
                    int idx = a.t_encode - itime - 1;
                    if (idx < 0 || idx >= m.n_lg) {
                        if (a.error_flag && sidx < m.n_sfi && m.sfi[sidx] != 0.f) *a.error_flag = 1;
                        idx = idx < 0 ? 0 : m.n_lg - 1;
                    }
                    code = 1u + (uint32_t)(sidx * m.n_lg + idx);
                    if (s_A[code] == 0.f || s_G[code] == 0.f) code = 0;   // contributes exactly 0 to every lambda
                }
                const uint32_t hot = __ballot_sync(0xffffffffu, code != 0u);
                if (valid) a.code_out[a.row_offset + row] = (uint8_t)code;
                if (lane == 0) a.hot_bitmap_out[(a.row_offset >> 5) + task] = hot;
            }
            const uint32_t sus_o = __ballot_sync(0xffffffffu, is_s);
            const uint32_t exp_o = __ballot_sync(0xffffffffu, is_e);
            const uint32_t inf_o = __ballot_sync(0xffffffffu, is_i);
            if (lane == 0) {
                if (do_init || sus_o != sus_w) a.sus_bitmap[task] = sus_o;
                if (do_init || exp_o != exp_w) a.exp_bitmap[task] = exp_o;
                if (do_init || inf_o != inf_w) a.inf_bitmap[task] = inf_o;
            }
        }
    }

    // ---- daily sums: fixed-point integers, so the result does not depend on the schedule
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_ne += __shfl_down_sync(0xffffffffu, acc_ne, o);
        acc_dne += __shfl_down_sync(0xffffffffu, acc_dne, o);
        acc_dead += __shfl_down_sync(0xffffffffu, acc_dead, o);
    }
    if (lane == 0) { s_red[warp][0] = acc_ne; s_red[warp][1] = acc_dne; s_red[warp][2] = acc_dead; }
    __syncthreads();
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(ws + WS_ACC);
    int* ticket = reinterpret_cast<int*>(ws + (MODE == MODE_GATHER ? WS_GTICKET : WS_TICKET));
    if (tid == 0) {
        if (MODE != MODE_GATHER) {
            long long s0 = 0, s1 = 0, s2 = 0;
            for (int w = 0; w < NW; ++w) { s0 += s_red[w][0]; s1 += s_red[w][1]; s2 += s_red[w][2]; }
            if (s0) atomicAdd(&acc[0], (unsigned long long)s0);
            if (s1) atomicAdd(&acc[1], (unsigned long long)s1);
            if (s2) atomicAdd(&acc[2], (unsigned long long)s2);
        }
        __threadfence();
        s_is_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_is_last && tid == 0) {
        __threadfence();
        if (MODE == MODE_GATHER) {
            for (int b = 0; b < a.n_blocks; ++b) reinterpret_cast<int*>(ws + WS_GCOUNTERS)[b] = 0;
        } else {
            const volatile long long* av = reinterpret_cast<const volatile long long*>(acc);
            const double s0 = (double)av[0] / FX_SCALE, s1 = (double)av[1] / FX_SCALE, s2 = (double)av[2];
            if (do_trans) {
                if (a.daily_infected) a.daily_infected[t] += (float)s0;          // transition.py:235
                if (a.sum_dne_dr) a.sum_dne_dr[t] = (float)s1;
            }
            if (do_prog) {
                if (a.death_count) a.death_count[t] = (float)s2;
                if (a.daily_deaths) a.daily_deaths[t] += (float)s2 * a.m_value[0];   // progression :52-60
            }
            acc[0] = 0; acc[1] = 0; acc[2] = 0;
            *reinterpret_cast<int*>(ws + WS_COUNTER) = 0;
        }
        *ticket = 0;
    }
}

// -------------------------------------------------------------------------------- backward kernel
// grad_r[t] = gI[t]*sum_dne[t] + sum_i dne[t,i] * c_i,  c_i = gD[recovered_step_i] * M / 2
#define AT_BWD_THREADS 256
__global__ void __launch_bounds__(AT_BWD_THREADS)
seirm_episode_bwd_kernel(int T, int64_t n, int64_t ld, const float* __restrict__ dne, const uint8_t* __restrict__ rstep,
                         const float* __restrict__ sum_dne, const float* __restrict__ death_count,
                         const float* __restrict__ gI, const float* __restrict__ gD, const float* __restrict__ m_value,
                         float* grad_r, float* grad_m, char* ws)
{
    __shared__ float s_gd[256];
    __shared__ double s_red[AT_BWD_THREADS / 32];
    __shared__ double s_tot[BWD_MAXT];
    __shared__ int s_is_last;
    double* partials = reinterpret_cast<double*>(ws + WS_BPART);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float half_m = m_value[0] * 0.5f;
    if (tid < 256) s_gd[tid] = (tid < T) ? gD[tid] * half_m : 0.f;
    if (tid < BWD_MAXT) s_tot[tid] = 0.0;
    __syncthreads();
    // static tile assignment + fixed-order folds: deterministic
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
        if (!__syncthreads_or(warp_any)) continue;
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
    int* ticket = reinterpret_cast<int*>(ws + WS_BTICKET);
    if (tid < T) partials[(size_t)blockIdx.x * BWD_MAXT + tid] = s_tot[tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) s_is_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_is_last) {
        __threadfence();
        if (tid < T) {
            double tot = 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) tot += ((const volatile double*)partials)[(size_t)b * BWD_MAXT + tid];
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

static int code_count(int n_lg) { return ((1 + 3 * n_lg) + 7) & ~7; }
static size_t fixed_smem(int n_sfs, int n_lg) {
    return (size_t)NW * WARP_SMEM + (size_t)2 * n_sfs * code_count(n_lg) * 8 + 2 * 256 * 4;
}

template <int MODE, bool WEIGHTED, typename TimerT, typename AgeT>
static int launch_kernel(const KModel& km, const at_seirm_launch& a, int grid, size_t smem, int bm_words, int NC, int G, cudaStream_t s) {
    auto k = seirm_kernel<MODE, WEIGHTED, TimerT, AgeT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    k<<<grid, NT, smem, s>>>(km, a, bm_words, NC, G);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail("seirm_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

template <int MODE, bool WEIGHTED>
static int dispatch_types(const KModel& km, const at_seirm_launch& a, int grid, size_t smem, int words, int NC, int G, cudaStream_t s) {
    if (a.timer_dtype == AT_TIMER_F32 && a.age_dtype == AT_AGE_F32) return launch_kernel<MODE, WEIGHTED, float, float>(km, a, grid, smem, words, NC, G, s);
    if (a.timer_dtype == AT_TIMER_F32 && a.age_dtype == AT_AGE_U8) return launch_kernel<MODE, WEIGHTED, float, uint8_t>(km, a, grid, smem, words, NC, G, s);
    if (a.timer_dtype == AT_TIMER_I16 && a.age_dtype == AT_AGE_F32) return launch_kernel<MODE, WEIGHTED, int16_t, float>(km, a, grid, smem, words, NC, G, s);
    if (a.timer_dtype == AT_TIMER_I16 && a.age_dtype == AT_AGE_U8) return launch_kernel<MODE, WEIGHTED, int16_t, uint8_t>(km, a, grid, smem, words, NC, G, s);
    return fail("%s", "unsupported timer_dtype / age_dtype");
}

template <int MODE>
static int dispatch(const KModel& km, const at_seirm_launch& a, int grid, size_t smem, int words, int NC, int G, cudaStream_t s) {
    if (a.edge_weight != nullptr) return dispatch_types<MODE, true>(km, a, grid, smem, words, NC, G, s);
    return dispatch_types<MODE, false>(km, a, grid, smem, words, NC, G, s);
}

extern "C" {

int at_seirm_abi_version(void) { return AT_SEIRM_ABI_VERSION; }
int64_t at_seirm_sizeof_launch(void) { return (int64_t)sizeof(at_seirm_launch); }
int64_t at_seirm_sizeof_model(void) { return (int64_t)sizeof(at_seirm_model); }
const char* at_last_error(void) { return g_err; }
int64_t at_seirm_workspace_bytes(void) { return (int64_t)WS_BYTES; }

int64_t at_seirm_max_block_size(int32_t n_sfs, int32_t n_lam_gamma, int32_t want_grad) {
    (void)want_grad;
    if (device_props()) return 0;
    int64_t bytes = (int64_t)g_smem_optin - (int64_t)fixed_smem(n_sfs, n_lam_gamma) - 2048;   // static __shared__ + slack
    if (bytes < 4096) return 0;
    return ((bytes * 8) / 1024) * 1024;
}

int at_seirm_step(const at_seirm_model* model, const at_seirm_launch* launch, void* stream) {
    if (!model || !launch) return fail("%s", "null argument");
    const at_seirm_launch& a = *launch;
    if (a.abi_version != AT_SEIRM_ABI_VERSION) return fail("%s", "at_seirm_launch.abi_version mismatch");
    if (model->n_lam_gamma < 1 || model->n_lam_gamma > AT_MAX_LAMGAMMA) return fail("%s", "n_lam_gamma out of range (1..84)");
    if (model->n_sfs < 1 || model->n_sfs > AT_MAX_SFS || model->n_sfi < 1 || model->n_sfi > AT_MAX_SFI) return fail("%s", "table size out of range");
    if (a.n_rows < 0 || (a.row_offset & 31) != 0 || (a.n_rows_pad & 31) != 0 || a.n_rows_pad < a.n_rows) return fail("%s", "bad n_rows / n_rows_pad / row_offset");
    if (!a.stage || !a.workspace || !a.sus_bitmap || !a.exp_bitmap || !a.inf_bitmap) return fail("%s", "stage / workspace / row bitmaps must not be null");
    const bool do_trans = a.flags & AT_DO_TRANSMISSION;
    if (do_trans) {
        if (!a.deg || !a.seg_base4 || !a.col || !a.code_in || !a.hot_bitmap_in || !a.r_value || !a.age) return fail("%s", "transmission needs the graph, code_in, hot_bitmap_in, age, r_value");
        if (a.n_blocks < 1 || a.n_blocks > AT_MAX_BLOCKS || a.block_size < 1024 || (a.block_size & 1023)) return fail("%s", "bad n_blocks / block_size");
        if ((int64_t)a.n_blocks * a.block_size < a.n_sources) return fail("%s", "n_blocks * block_size does not cover n_sources");
        if (a.iso_mode == AT_ISO_TENSOR && !a.isolation) return fail("%s", "iso_mode TENSOR without isolation pointer");
        if (a.u_mode == AT_U_TENSOR && !a.uniform) return fail("%s", "u_mode TENSOR without uniform pointer");
        if (a.n_blocks > 1 && (!a.lam_part || !a.dirty || (a.dne_dr && !a.lam1_part))) return fail("%s", "n_blocks > 1 needs lam_part / lam1_part / dirty");
        if (a.flags & AT_DO_INIT) return fail("%s", "AT_DO_INIT cannot be combined with AT_DO_TRANSMISSION");
    }
    if ((a.flags & AT_DO_PROGRESSION) && (!a.next_stage_time || (a.daily_deaths && !a.m_value))) return fail("%s", "progression needs next_stage_time and m_value");
    if ((a.flags & AT_DO_ENCODE) && (!a.code_out || !a.hot_bitmap_out || !a.infected_time)) return fail("%s", "encode needs code_out, hot_bitmap_out, infected_time");
    if (a.t < 0 || a.t > 255) return fail("%s", "t out of range (0..255)");
    if (device_props()) return fail("%s", "cannot query device");
    if (a.n_rows == 0) return 0;

    KModel km;
    km.n_sfs = model->n_sfs; km.n_sfi = model->n_sfi; km.n_lg = model->n_lam_gamma;
    km.e2i = model->exposed_to_infected_time; km.i2r = model->infected_to_recovered_time; km.inf_time = model->infinity_time;
    km.ibar = model->i_bar;
    memcpy(km.sfs, model->sfs, sizeof(km.sfs)); memcpy(km.sfi, model->sfi, sizeof(km.sfi)); memcpy(km.lg, model->lam_gamma, sizeof(km.lg));

    const int NC = code_count(model->n_lam_gamma);
    const size_t fixed = fixed_smem(model->n_sfs, model->n_lam_gamma);
    int bm_words = 0;
    if (do_trans) {
        if (a.block_size > at_seirm_max_block_size(model->n_sfs, model->n_lam_gamma, a.dne_dr != nullptr)) return fail("%s", "block_size exceeds the shared-memory bitmap capacity (at_seirm_max_block_size)");
        bm_words = a.block_size / 32;
    }
    const int64_t ntasks = a.n_rows_pad / 32;
    int G = a.tasks_per_group > 0 ? a.tasks_per_group : (ntasks >= 131072 ? 8 : 4);
    if (G > MAXG) G = MAXG;
    const int64_t ngroups = (ntasks + G - 1) / G;
    int grid = a.grid > 0 ? a.grid : g_sm_count;
    if (grid > 1024) grid = 1024;
    cudaStream_t s = (cudaStream_t)stream;
    if (do_trans && a.n_blocks > 1) {
        int ggrid = grid < a.n_blocks ? a.n_blocks : grid;
        int rc = dispatch<MODE_GATHER>(km, a, ggrid, fixed + (size_t)bm_words * 4, bm_words, NC, G, s);
        if (rc) return rc;
        int64_t useful = (ngroups + NW - 1) / NW;
        int ugrid = grid > useful ? (int)useful : grid;
        return dispatch<MODE_UPDATE>(km, a, ugrid < 1 ? 1 : ugrid, fixed, 0, NC, G, s);
    }
    int64_t useful = (ngroups + NW - 1) / NW;
    if (grid > useful) grid = (int)useful;
    if (grid < 1) grid = 1;
    if (do_trans) return dispatch<MODE_FUSED>(km, a, grid, fixed + (size_t)bm_words * 4, bm_words, NC, G, s);
    return dispatch<MODE_UPDATE>(km, a, grid, fixed, 0, NC, G, s);
}

int at_seirm_episode_backward(int32_t num_steps, int64_t n_rows, int64_t ld, const float* dne_dr, const uint8_t* recovered_step,
                              const float* sum_dne_dr, const float* death_count, const float* g_daily_infected,
                              const float* g_daily_deaths, const float* m_value, float* grad_r, float* grad_m,
                              void* workspace, int32_t grid, void* stream) {
    if (num_steps < 1 || num_steps > BWD_MAXT) return fail("%s", "num_steps out of range (1..64)");
    if (!dne_dr || !recovered_step || !sum_dne_dr || !death_count || !g_daily_infected || !g_daily_deaths || !m_value || !grad_r || !grad_m || !workspace)
        return fail("%s", "null argument");
    if (device_props()) return fail("%s", "cannot query device");
    int g = grid > 0 ? grid : g_sm_count * 4;
    const int64_t tiles = (n_rows + AT_BWD_THREADS * 4 - 1) / (AT_BWD_THREADS * 4);
    if (g > tiles) g = (int)tiles;
    if (g > BWD_MAX_GRID) g = BWD_MAX_GRID;
    if (g < 1) g = 1;
    seirm_episode_bwd_kernel<<<g, AT_BWD_THREADS, 0, (cudaStream_t)stream>>>(num_steps, n_rows, ld, dne_dr, recovered_step, sum_dne_dr,
        death_count, g_daily_infected, g_daily_deaths, m_value, grad_r, grad_m, (char*)workspace);
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
