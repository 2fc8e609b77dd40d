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

#ifndef AT_HOST_EMULATION           // tests/host_emulation/ compiles this file for the CPU: a SIMT shim stands in for CUDA
#include <cuda_runtime.h>
#define AT_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define AT_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define AT_DYNAMIC_SMEM(type, name, alignment) extern __shared__ __align__(alignment) type name[]
#endif
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "at_seirm.h"
#include "at_noise.cuh"

#define NT_MAX 640                  // max threads per CTA (20 warps); the host may launch fewer warps
#define MAXG 8                      // max gather tasks per scheduling group
#define FX_SCALE 4294967296.0       // 2^32 fixed point for the daily sums

#define MODE_FUSED 0                // n_blocks == 1: gather + update per 32-row task
#define MODE_GATHER 1               // n_blocks > 1: partial lambdas of every column block, 256-row tasks
#define MODE_UPDATE 2               // update only (lambda from partials, or no transmission)
#define MODE_HASHED 3               // like FUSED, but the shared-memory bitmap is a HASHED filter: bit (src mod M);
                                    // col holds raw source ids and false positives die on code == 0

// workspace layout (bytes)
#define WS_ACC 0                    // int64[3]: sum ne, sum dne, deaths (fixed point)
#define WS_TICKET 32                // int: CTAs finished (main kernel)
#define WS_COUNTER 36               // int: group counter (main kernel)
#define WS_ANYHOT 40                // uint[3]: some agent of this rank is infectious in the step just encoded, one word per code
                                    // buffer in flight (at_seirm_launch.xchg_slot; read + reset by the exchange)
#define WS_HASEI 52                 // uint[3]: some agent of this rank is in E or I after the step just done (same slots as WS_ANYHOT)
#define WS_GCOUNTERS 64             // int[AT_MAX_BLOCKS]: group counters of the gather kernel
#define WS_GTICKET 384              // int
#define WS_BTICKET 512              // int: backward kernel
#define WS_BPART 1024               // double[BWD_MAX_GRID][BWD_MAXT]
#define BWD_MAX_GRID 1024
#define BWD_MAXT 64
#define WS_FIXSHIFT (WS_BPART + BWD_MAX_GRID * BWD_MAXT * 8)      // int: binary scale of the fixed-point terms of this step (stream layout)
#define WS_FIXTAB (WS_FIXSHIFT + 16)                             // int64[AT_MAX_SFS * 256]: the table itself, gather kernel -> update kernel
#define WS_BYTES (WS_FIXTAB + AT_MAX_SFS * 256 * 8)

static thread_local char g_err[512] = "";

static int fail(const char* fmt, const char* detail) {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return 1;
}

// ------------------------------------------------------------------------------------------ noise
// at_noise_u64: see at_noise.cuh
__host__ __device__ __forceinline__ float at_u_exposure(uint64_t z) { return (float)(z >> 40) * (1.0f / 16777216.0f); }
__host__ __device__ __forceinline__ float at_u_isolation(uint64_t z) {
    return (float)((z >> 8) & 0xFFFFFFull) * (1.0f / 16777216.0f);
}

// --------------------------------------------------------------------------------- typed accessors
__device__ __forceinline__ int load_timer(const void* p, int64_t i, bool f32) {
    return f32 ? (int)((const float*)p)[i] : (int)((const int16_t*)p)[i];
}
__device__ __forceinline__ void store_timer(void* p, int64_t i, int v, bool f32) {
    if (f32) ((float*)p)[i] = (float)v; else ((int16_t*)p)[i] = (int16_t)v;
}
__device__ __forceinline__ int load_age(const void* p, int64_t i, bool f32) {
    return f32 ? (int)((const float*)p)[i] : (int)((const uint8_t*)p)[i];
}

// ------------------------------------------------------------------------------- TMA / mbarrier PTX
#ifndef AT_HOST_EMULATION
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
__device__ __forceinline__ void fence_mbarrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#else
// host emulation: the bulk copy is a memcpy by the issuing thread; the mbarrier word counts completed phases (one copy per
// phase here), and a wait on parity P returns once the current phase's parity differs from P — the hardware's rule
inline void mbar_init(uint64_t* bar, uint32_t) { __atomic_store_n(bar, (uint64_t)0, __ATOMIC_RELEASE); }
inline void mbar_expect_tx(uint64_t*, uint32_t) {}
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    memcpy(dst, src, bytes);
    __atomic_fetch_add(bar, (uint64_t)1, __ATOMIC_RELEASE);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    while ((__atomic_load_n(bar, __ATOMIC_ACQUIRE) & 1u) == parity) __nanosleep(0);
}
inline void fence_mbarrier_init() {}
inline void fence_proxy_async_smem() {}
#endif

struct KModel {   // device copy of at_seirm_model (kernel parameter, by value)
    int n_sfs, n_sfi, n_lg, e2i, i2r, inf_time;
    float ibar;
    float sfs[AT_MAX_SFS], sfi[AT_MAX_SFI], lg[AT_MAX_LAMGAMMA];
    int n_layers;                       // >= 1
    float ibar_layers[AT_MAX_LAYERS];   // [0] == ibar for single-layer graphs
};

// per-warp shared memory: two col chunks, the hit mask and hit codes of the chunk being folded, two mbarriers
__host__ __device__ constexpr int warp_smem_bytes(int cap) { return 2 * cap * 4 + cap + 128 + 64; }

// -------------------------------------------------------------------------------------- the kernel
// RL4 = 128-bit loads per lane per chunk: a chunk holds CAP = 128*RL4 edges and lane L owns the RL = 4*RL4
// consecutive edges [L*RL, (L+1)*RL) of it.
template <int MODE, int RL4>
__global__ void __launch_bounds__(NT_MAX, 1)
seirm_kernel(const KModel m, const at_seirm_launch a, const int bm_words, const int NC, const int G)
{
    constexpr int RL = 4 * RL4, CAP = 32 * RL, WSM = warp_smem_bytes(CAP);
    constexpr int ROWS = (MODE == MODE_GATHER) ? 8 : 1;        // rows per lane in a gather task
    constexpr bool HASHED = (MODE == MODE_HASHED);
    AT_DYNAMIC_SMEM(unsigned char, smem_raw, 128);
    const int nw = blockDim.x >> 5;
    // [per-warp areas][Tdiv f64 n_sfs*NC][T1div f64 n_sfs*NC][s_A f32 256][s_G f32 256][hot bitmap]
    double* s_T = reinterpret_cast<double*>(smem_raw + (MODE != MODE_UPDATE ? nw * WSM : 0));
    double* s_T1 = s_T + m.n_sfs * NC;
    float* s_A = reinterpret_cast<float*>(s_T1 + m.n_sfs * NC);
    float* s_G = s_A + 256;
    const unsigned char* s_bm = reinterpret_cast<const unsigned char*>(s_G + 256);
    __shared__ long long s_red[NT_MAX / 32][3];
    __shared__ int s_is_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool do_trans = (a.flags & AT_DO_TRANSMISSION) != 0;
    const bool do_prog = (a.flags & AT_DO_PROGRESSION) != 0;
    const bool do_enc = (a.flags & AT_DO_ENCODE) != 0;
    const bool do_init = (a.flags & AT_DO_INIT) != 0;
    const bool lookahead = (a.flags & AT_DO_LOOKAHEAD) != 0;
    const bool want_grad = a.dne_dr != nullptr;
    const bool weighted = a.edge_weight != nullptr;
    const bool timer_f32 = a.timer_dtype == AT_TIMER_F32, age_f32 = a.age_dtype == AT_AGE_F32;
    constexpr bool HAS_GATHER = (MODE != MODE_UPDATE);
    const uint32_t hmask = (uint32_t)(bm_words - 1) << 2;        // HASHED: byte-offset mask of the filter words
    const int t = a.t;
    const int block = (MODE == MODE_GATHER) ? (int)(blockIdx.x % a.n_blocks) : 0;
    const float R = do_trans ? a.r_value[a.r_index] : 0.f;
    const double ibar = (double)m.ibar;
    char* ws = reinterpret_cast<char*>(a.workspace);

    __shared__ int s_any_hot, s_any_out, s_any_ei;
    if (tid == 0) { s_any_hot = 0; s_any_out = 0; s_any_ei = 0; }
    // ---- tables: code -> (SFInfector[int(stage)], lam_gamma[idx]); code = 1 + sidx*K + idx
    // (strided: a CTA may have fewer than 256 threads — few warps next to a large bitmap, or warps_per_cta set by the caller)
    for (int c = tid; c < 256; c += blockDim.x) {
        float av = 0.f, gv = 0.f;
        if (c >= 1 && c <= 3 * m.n_lg) {
            const int sidx = (c - 1) / m.n_lg, idx = (c - 1) % m.n_lg;
            av = sidx < m.n_sfi ? m.sfi[sidx] : 0.f;
            gv = m.lg[idx];
        }
        s_A[c] = av;
        s_G[c] = gv;
    }
    __syncthreads();
    unsigned char* wsm = smem_raw + warp * WSM;
    int32_t* s_col = reinterpret_cast<int32_t*>(wsm);
    uint8_t* s_code = wsm + 2 * CAP * 4;                                              // [CAP] hit codes
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(wsm + 2 * CAP * 4 + CAP);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(wsm + 2 * CAP * 4 + CAP + 128);
    if (HAS_GATHER) {
        // message of an unweighted edge, exactly as the reference forms it (transition.py:78-83):
        // float32 product chain ((R*S_i)*A_j)*G_j (B_n == 1 is exact), then float64 division by I_bar
        for (int i = tid; i < m.n_sfs * NC; i += blockDim.x) {
            const int ag = i / NC, c = i % NC;
            const float s = m.sfs[ag];
            const float av = c < 256 ? s_A[c] : 0.f, gv = c < 256 ? s_G[c] : 0.f;
            s_T[i] = (double)__fmul_rn(__fmul_rn(__fmul_rn(R, s), av), gv) / ibar;
            s_T1[i] = (double)__fmul_rn(__fmul_rn(s, av), gv) / ibar;
        }
        const uint4* src = reinterpret_cast<const uint4*>(a.hot_bitmap_in + (size_t)block * (a.block_size >> 5));
        uint4* dst = reinterpret_cast<uint4*>(const_cast<unsigned char*>(s_bm));
        uint32_t seen = 0;
        if (!HASHED) {
            for (int i = tid; i < (bm_words >> 2); i += blockDim.x) {
                const uint4 v = __ldg(src + i);
                dst[i] = v;
                seen |= v.x | v.y | v.z | v.w;
            }
        } else {
            // hashed filter: word i = OR of the words i, i+M, i+2M, ... of the full bitmap (M = bm_words, power of 2)
            const int total4 = (int)(a.n_sources >> 7);
            if (a.hot_filter_in != nullptr) {
                // already folded by the last CTA of the previous launch
                const uint4* fsrc = reinterpret_cast<const uint4*>(a.hot_filter_in);
                for (int i = tid; i < (bm_words >> 2); i += blockDim.x) dst[i] = __ldg(fsrc + i);
                seen = __ldg(a.hot_filter_in + bm_words);
            } else
            for (int i = tid; i < (bm_words >> 2); i += blockDim.x) {
                uint4 acc4 = make_uint4(0, 0, 0, 0);
                for (int k = i; k < total4; k += (bm_words >> 2)) {
                    const uint4 v = __ldg(src + k);
                    acc4.x |= v.x; acc4.y |= v.y; acc4.z |= v.z; acc4.w |= v.w;
                }
                dst[i] = acc4;
                seen |= acc4.x | acc4.y | acc4.z | acc4.w;
            }
        }
        if (seen != 0u) s_any_hot = 1;                       // benign race: every writer stores 1
        if (lane == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
        fence_mbarrier_init();
    }
    __syncthreads();

    const bool any_hot = s_any_hot != 0;
    // hashed filter with many sources per bit: most positives are false, check the exact bitmap before the code array
    const bool exact_check = HASHED && a.n_sources > (int64_t)bm_words * 32 * 24;
    const int ntasks32 = (int)(a.n_rows_pad >> 5);                            // 32-row tasks (bitmap words)
    const int ntasks = (MODE == MODE_GATHER) ? (int)(a.n_rows_pad >> 8) : ntasks32;
    const uint16_t* rs_b = HAS_GATHER ? a.row_start + (size_t)block * a.n_rows_pad : nullptr;
    const int2* seg_b = HAS_GATHER ? reinterpret_cast<const int2*>(a.seg_info) + (size_t)block * ntasks : nullptr;
    const uint8_t* code_b = HAS_GATHER ? a.code_in + (size_t)block * a.block_size : nullptr;
    const uint64_t seed = a.seed_ptr ? *a.seed_ptr : a.seed;
    if (HAS_GATHER) {   // the staging buffers must always hold valid encodings (lanes read past a chunk's end)
        for (int i = lane; i < 2 * CAP; i += 32) s_col[i] = 0;
        fence_proxy_async_smem();
        __syncwarp();
    }

    // ---- static cyclic schedule: this warp owns tasks first, first+stride, ...  (deterministic, and the
    //      next tasks are known in advance, so their metadata and edges are prefetched)
    int first, stride;
    if (MODE == MODE_GATHER) {
        const int ctas_b = ((int)gridDim.x - block + a.n_blocks - 1) / a.n_blocks;      // CTAs serving this block
        first = ((int)blockIdx.x / a.n_blocks) * nw + warp;
        stride = ctas_b * nw;
    } else {
        first = (int)blockIdx.x * nw + warp;
        stride = (int)gridDim.x * nw;
    }

    // Task metadata is prefetched three tasks ahead: `raw` is in flight (loaded last iteration), m2/m1/m0 are
    // finalised (nothing below touches a freshly loaded register, so the loads never stall the pipeline).
    struct Meta { int base4, len; uint32_t sus, exp, inf; };
    struct MetaRaw { int2 seg; uint32_t sus, exp, inf; uint4 s0, s1; };
    auto load_raw = [&](int task) -> MetaRaw {
        MetaRaw r; r.seg = make_int2(0, 0); r.sus = 0; r.exp = 0; r.inf = 0;
        r.s0 = make_uint4(0, 0, 0, 0); r.s1 = make_uint4(0, 0, 0, 0);
        if (task < ntasks) {
            if (MODE == MODE_GATHER) {
                if (do_trans) {
                    const uint4* sw = reinterpret_cast<const uint4*>(a.sus_bitmap + (size_t)task * 8);
                    r.s0 = sw[0]; r.s1 = sw[1];
                }
            } else if (!do_init) {
                r.sus = a.sus_bitmap[task];
                r.exp = a.exp_bitmap[task];
                r.inf = a.inf_bitmap[task];
            }
            if (HAS_GATHER && do_trans) r.seg = seg_b[task];
        }
        return r;
    };
    auto finalize = [&](const MetaRaw& r) -> Meta {
        Meta mt;
        mt.sus = (MODE == MODE_GATHER) ? (r.s0.x | r.s0.y | r.s0.z | r.s0.w | r.s1.x | r.s1.y | r.s1.z | r.s1.w) : r.sus;
        mt.exp = r.exp; mt.inf = r.inf;
        mt.base4 = r.seg.x;
        // no susceptible row, or no infectious source at all (in this column block): nothing to gather
        mt.len = (mt.sus != 0u && any_hot) ? r.seg.y : 0;
        return mt;
    };
    // per-lane row metadata of the NEXT task (row offsets inside its segment, destination age)
    struct RowMeta { uint4 rs; int age; };
    auto load_rows = [&](int task) -> RowMeta {
        RowMeta rm; rm.rs = make_uint4(0, 0, 0, 0); rm.age = 0;
        if (HAS_GATHER && do_trans && task < ntasks) {
            const int64_t row0 = (int64_t)task * (32 * ROWS) + lane * ROWS;
            if (ROWS == 1) {
                rm.rs.x = rs_b[row0];
                if (row0 < a.n_rows) rm.age = load_age(a.age, row0, age_f32);
            } else {
                rm.rs = *reinterpret_cast<const uint4*>(rs_b + row0);
            }
        }
        return rm;
    };
    // first chunks of consecutive tasks alternate between the two slots; a task's later chunks reuse its slot
    uint32_t first_issued = 0, first_consumed = 0, waits0 = 0, waits1 = 0;
    auto issue_chunk = [&](int base4, int len, int ci, uint32_t slot) {
        if (lane == 0) {
            const int off = ci * CAP;
            const uint32_t bytes = (uint32_t)((min(CAP, len - off) + 3) & ~3) * 4u;
            mbar_expect_tx(&s_bar[slot], bytes);
            bulk_g2s(s_col + slot * CAP, a.col + ((int64_t)base4 * 4 + off), bytes, &s_bar[slot]);
        }
    };

    long long acc_ne = 0, acc_dne = 0, acc_dead = 0;

    // ------------------------------------------------------------------ update of one 32-row subtask (lane <-> row)
    auto update_subtask = [&](int task, uint32_t sus_w, uint32_t exp_w, uint32_t inf_w, double lam, double lam1) {
        const int64_t row = (int64_t)task * 32 + lane;
        const bool valid = row < a.n_rows;
        const bool task_valid = (int64_t)task * 32 < a.n_rows;     // (hot words of padding tasks belong to somebody else)
        float stage = 0.f;
        bool stage_known = false;
        if (do_init) {
            stage = valid ? a.stage[row] : 3.0f;
            stage_known = true;
            sus_w = __ballot_sync(0xffffffffu, stage == 0.0f);
            exp_w = __ballot_sync(0xffffffffu, stage == 1.0f);
            inf_w = __ballot_sync(0xffffffffu, stage == 2.0f);
        }
        bool is_s = (sus_w >> lane) & 1u, is_e = (exp_w >> lane) & 1u, is_i = (inf_w >> lane) & 1u;
        const bool exposed_candidate = do_trans && is_s && (lam > 0.0 || lam1 > 0.0);
        // nothing to do: no exposure candidate, nobody in E or I
        if (!do_init && (exp_w | inf_w) == 0u && !__any_sync(0xffffffffu, exposed_candidate)) {
            if (want_grad && do_trans && valid) a.dne_dr[row] = 0.f;
            if (do_enc) {
                if (valid) a.code_out[a.row_offset + row] = 0;
                if (lane == 0 && task_valid) a.hot_bitmap_out[(a.row_offset >> 5) + task] = 0u;
            }
            return;
        }
        float dne = 0.f;
        int nst = 0, itime = 0;
        bool nst_known = false, itime_known = false;
        if (exposed_candidate) {
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
            const float q32 = __expf(-(float)lam);
            const float p32 = 1.0f - q32;
            bool hit;
            if (fabsf(u - p32) > 2e-5f) hit = u < p32;       // __expf: |error of p32| < 4e-6 for lambda < 88
            else hit = (double)u < 1.0 - exp(-lam);
            const float omw = 1.0f - wiso;                  // float32, as (1.0 - will_isolate)
            const float ne = hit ? omw : 0.0f;
            if (want_grad) { dne = omw * q32 * (float)lam1; acc_dne += __double2ll_rn((double)dne * FX_SCALE); }
            if (ne != 0.0f) {
                stage = ne; stage_known = true;            // 0 + ne
                a.stage[row] = stage;
                nst = t + 1 + m.e2i; nst_known = true;
                itime = t; itime_known = true;
                store_timer(a.next_stage_time, row, nst, timer_f32);
                store_timer(a.infected_time, row, itime, timer_f32);
                acc_ne += __double2ll_rn((double)ne * FX_SCALE);
                is_s = false;
                is_e = (stage == 1.0f);
            }
        }
        // common case: candidates were tested, nobody was exposed, nobody is in E or I
        if (!do_init && (exp_w | inf_w) == 0u && !__any_sync(0xffffffffu, stage_known)) {
            if (want_grad && do_trans && valid) a.dne_dr[row] = dne;
            if (do_enc) {
                if (valid) a.code_out[a.row_offset + row] = 0;
                if (lane == 0 && task_valid) a.hot_bitmap_out[(a.row_offset >> 5) + task] = 0u;
            }
            return;
        }
        if (do_prog && (is_e || is_i)) {
            if (!nst_known) { nst = load_timer(a.next_stage_time, row, timer_f32); nst_known = true; }
            if (nst <= t) {
                if (is_i) {
                    acc_dead += 1;                         // stage * rd / INFECTED_VAR = 2/2
                    if (a.recovered_step) a.recovered_step[row] = (uint8_t)t;
                }
                if (nst == t) { nst = is_i ? m.inf_time : t + m.i2r; store_timer(a.next_stage_time, row, nst, timer_f32); }
                stage = is_i ? 3.0f : 2.0f;
                stage_known = true;
                a.stage[row] = stage;
                const bool was_e = is_e;
                is_e = false;
                is_i = was_e;
            }
        }
        if (want_grad && valid && do_trans) a.dne_dr[row] = dne;
        if (do_enc) {
            uint32_t code = 0;
            // infectious = 0 < stage < 3; fractional stages (soft isolation, Q4) index SFInfector[0]
            const bool frac = stage_known && stage > 0.0f && stage < 3.0f && !(is_e || is_i);
            bool ce = is_e, ci = is_i;                       // classes the code is computed for
            if (lookahead && (is_e || is_i)) {
                // the stage after the progression of step t_encode-1 (progression :41-88: E/I rows that are due move on)
                if (!nst_known) { nst = load_timer(a.next_stage_time, row, timer_f32); nst_known = true; }
                if (nst <= a.t_encode - 1) { ci = ce; ce = false; }
            }
            if (ce || ci || frac) {
                if (!itime_known) itime = load_timer(a.infected_time, row, timer_f32);
                const int sidx = ce ? 1 : (ci ? 2 : (int)stage);
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
            if (lane == 0 && task_valid) {
                a.hot_bitmap_out[(a.row_offset >> 5) + task] = hot;
                if (hot != 0u) s_any_out = 1;                    // benign race: every writer stores 1
            }
        }
        const uint32_t sus_o = __ballot_sync(0xffffffffu, is_s);
        const uint32_t exp_o = __ballot_sync(0xffffffffu, is_e);
        const uint32_t inf_o = __ballot_sync(0xffffffffu, is_i);
        if (lane == 0) {
            if ((exp_o | inf_o) != 0u) s_any_ei = 1;             // benign race: every writer stores 1
            if (do_init || sus_o != sus_w) a.sus_bitmap[task] = sus_o;
            if (do_init || exp_o != exp_w) a.exp_bitmap[task] = exp_o;
            if (do_init || inf_o != inf_w) a.inf_bitmap[task] = inf_o;
        }
    };

    const bool update_only = (MODE == MODE_UPDATE) || (MODE != MODE_GATHER && !(do_trans && any_hot));
    if (update_only) {
        // ---- update-only pass (also taken by the fused kernels when no source is infectious: lambda == 0 everywhere)
        // ---- update-only kernel (blocked mode second pass, or launches without transmission): supertasks of 8
        //      subtasks; lane i < 8 prefetches the bitmap words of subtask i one supertask ahead, the dirty words of
        //      the column blocks one subtask ahead; partial lambdas are folded in block order
        const int nsuper = (ntasks32 + 7) >> 3;
        const bool use_partials = do_trans && a.dirty != nullptr;
        auto load_words = [&](int st, uint32_t& sw, uint32_t& ew, uint32_t& iw) {
            sw = 0; ew = 0; iw = 0;
            const int task = st * 8 + lane;
            if (!do_init && lane < 8 && st < nsuper && task < ntasks32) {
                sw = a.sus_bitmap[task]; ew = a.exp_bitmap[task]; iw = a.inf_bitmap[task];
            }
        };
        uint32_t nsus, nexp, ninf;
        load_words(first, nsus, nexp, ninf);
        for (int st = first; st < nsuper; st += stride) {
            const uint32_t csus = nsus, cexp = nexp, cinf = ninf;
            load_words(st + stride, nsus, nexp, ninf);
            if (!do_init && !use_partials && st * 8 + 8 <= ntasks32 && (int64_t)st * 256 + 256 <= a.n_rows &&
                !__any_sync(0xffffffffu, (cexp | cinf) != 0u)) {
                // all 256 rows inert (lambda == 0, nobody in E or I): only the per-step outputs have to be cleared
                const int64_t row0 = (int64_t)st * 256 + lane * 8;
                if (want_grad && do_trans) {
                    float4* d4 = reinterpret_cast<float4*>(a.dne_dr + row0);
                    d4[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                    d4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (do_enc) {
                    *reinterpret_cast<unsigned long long*>(a.code_out + a.row_offset + row0) = 0ull;
                    if (lane < 8) a.hot_bitmap_out[(a.row_offset >> 5) + st * 8 + lane] = 0u;
                }
                continue;
            }
            uint32_t dwl_next = 0;
            if (use_partials && lane < a.n_blocks) dwl_next = a.dirty[(size_t)lane * ntasks32 + st * 8];
            for (int r = 0; r < 8; ++r) {
                const int task = st * 8 + r;
                if (task >= ntasks32) break;
                const uint32_t sus_w = __shfl_sync(0xffffffffu, csus, r), exp_w = __shfl_sync(0xffffffffu, cexp, r),
                               inf_w = __shfl_sync(0xffffffffu, cinf, r);
                const uint32_t dwl = dwl_next;
                if (use_partials && r < 7 && task + 1 < ntasks32 && lane < a.n_blocks) dwl_next = a.dirty[(size_t)lane * ntasks32 + task + 1];
                double lam = 0.0, lam1 = 0.0;
                if (use_partials && sus_w != 0u) {
                    uint32_t any_b = __ballot_sync(0xffffffffu, dwl != 0u);
                    while (any_b) {
                        const int bl = __ffs(any_b) - 1;
                        any_b &= any_b - 1;
                        const uint32_t dw = __shfl_sync(0xffffffffu, dwl, bl);
                        if ((dw >> lane) & 1u) {
                            const double2 v = reinterpret_cast<const double2*>(a.lam_part)[(size_t)bl * a.n_rows_pad + (size_t)task * 32 + lane];
                            lam += v.x; lam1 += v.y;
                        }
                    }
                    for (int b = 32; b < a.n_blocks; ++b) {                    // rare: more than 32 column blocks
                        const uint32_t dw = a.dirty[(size_t)b * ntasks32 + task];
                        if ((dw >> lane) & 1u) {
                            const double2 v = reinterpret_cast<const double2*>(a.lam_part)[(size_t)b * a.n_rows_pad + (size_t)task * 32 + lane];
                            lam += v.x; lam1 += v.y;
                        }
                    }
                }
                update_subtask(task, sus_w, exp_w, inf_w, lam, lam1);
            }
        }
    }

    Meta m0 = finalize(load_raw(first)), m1 = finalize(load_raw(first + stride)), m2 = finalize(load_raw(first + 2 * stride));
    MetaRaw raw = load_raw(first + 3 * stride);
    RowMeta rows_next = load_rows(first);
    if (HAS_GATHER && !update_only && m0.len > 0) { issue_chunk(m0.base4, m0.len, 0, first_issued & 1u); ++first_issued; }

    for (int task = first; !update_only && task < ntasks; task += stride) {
        const Meta mt = m0;
        const RowMeta rows_cur = rows_next;
        m0 = m1; m1 = m2;
        m2 = finalize(raw);
        raw = load_raw(task + 4 * stride);
        rows_next = load_rows(task + stride);
        // first chunk of the next task goes in flight before this one is processed
        if (HAS_GATHER && m0.len > 0) { issue_chunk(m0.base4, m0.len, 0, first_issued & 1u); ++first_issued; }

        double lam = 0.0, lam1 = 0.0;        // FUSED: lambda of my row; GATHER: of the row being folded
        // ================================================================== gather + fold
        if (HAS_GATHER && mt.len > 0) {
            const int len = mt.len;
            const int64_t row0 = (int64_t)task * (32 * ROWS) + lane * ROWS;     // my first row
            int rstart[ROWS + 1];
            uint32_t my_sus;
            if (ROWS == 1) {
                rstart[0] = (int)rows_cur.rs.x;
                my_sus = (mt.sus >> lane) & 1u;
            } else {
                const uint4 v = rows_cur.rs;
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int r = 0; r < ROWS; ++r) rstart[r] = (int)((w[r >> 1] >> ((r & 1) * 16)) & 0xFFFFu);
                my_sus = reinterpret_cast<const uint8_t*>(a.sus_bitmap)[(size_t)task * 32 + lane];
            }
            {
                const int nxt = __shfl_down_sync(0xffffffffu, rstart[0], 1);
                rstart[ROWS] = (lane == 31) ? len : nxt;
            }
            int cur_r = -1;                  // GATHER: row (0..7) whose partial lambda is being accumulated
            int cur_ag = rows_cur.age < 0 ? 0 : (rows_cur.age >= m.n_sfs ? m.n_sfs - 1 : rows_cur.age);
            uint32_t dbits = 0;
            auto flush_row = [&]() {          // GATHER: write the partial lambda of row cur_r
                if (MODE == MODE_GATHER && cur_r >= 0 && (lam != 0.0 || lam1 != 0.0)) {
                    double2* dst = reinterpret_cast<double2*>(a.lam_part) + (size_t)block * a.n_rows_pad + row0 + cur_r;
                    if (dbits & (1u << cur_r)) { const double2 old = *dst; lam += old.x; lam1 += old.y; }
                    *dst = make_double2(lam, lam1);
                    dbits |= 1u << cur_r;
                }
                lam = 0.0; lam1 = 0.0;
            };

            const int nch = (len + CAP - 1) / CAP;
            const uint32_t slot = first_consumed & 1u;
            ++first_consumed;
            for (int ci = 0; ci < nch; ++ci) {
                if (ci > 0) {
                    // rare: a task larger than one chunk; its later chunks are fetched synchronously into its own
                    // slot, which has just been drained (the other slot holds the next task's prefetch)
                    __syncwarp();
                    issue_chunk(mt.base4, len, ci, slot);
                }
                mbar_wait(&s_bar[slot], (slot ? waits1 : waits0) & 1u);
                if (slot) ++waits1; else ++waits0;
                const int n = min(CAP, len - ci * CAP);
                const int32_t* sc = s_col + slot * CAP;
                const int4* sc4 = reinterpret_cast<const int4*>(sc) + lane * RL4;
                uint32_t h = 0;
#pragma unroll
                for (int q = 0; q < RL4; ++q) {
                    const int4 c = sc4[q];
                    uint32_t w0, w1, w2, w3;
                    if (!HASHED) {
                        w0 = *reinterpret_cast<const uint32_t*>(s_bm + ((uint32_t)c.x >> 5));
                        w1 = *reinterpret_cast<const uint32_t*>(s_bm + ((uint32_t)c.y >> 5));
                        w2 = *reinterpret_cast<const uint32_t*>(s_bm + ((uint32_t)c.z >> 5));
                        w3 = *reinterpret_cast<const uint32_t*>(s_bm + ((uint32_t)c.w >> 5));
                    } else {
                        w0 = *reinterpret_cast<const uint32_t*>(s_bm + (((uint32_t)c.x >> 3) & hmask));
                        w1 = *reinterpret_cast<const uint32_t*>(s_bm + (((uint32_t)c.y >> 3) & hmask));
                        w2 = *reinterpret_cast<const uint32_t*>(s_bm + (((uint32_t)c.z >> 3) & hmask));
                        w3 = *reinterpret_cast<const uint32_t*>(s_bm + (((uint32_t)c.w >> 3) & hmask));
                    }
                    // funnel shift: each test result enters h at bit 31 and moves down, one instruction per edge
                    h = __funnelshift_r(h, w0 >> (c.x & 31), 1);
                    h = __funnelshift_r(h, w1 >> (c.y & 31), 1);
                    h = __funnelshift_r(h, w2 >> (c.z & 31), 1);
                    h = __funnelshift_r(h, w3 >> (c.w & 31), 1);
                }
                h >>= (32 - RL);                                        // edge j of this lane's run is bit j
                {   // drop the bits of positions past the end of the chunk
                    const int mine = n - lane * RL;
                    h &= mine >= RL ? 0xffffffffu >> (32 - RL) : (mine <= 0 ? 0u : ((1u << mine) - 1u));
                }
                if (__any_sync(0xffffffffu, h != 0u)) {
                    {   // the lane that found a hit fetches its code (all L2 gathers of the chunk in flight together)
                        uint32_t hh = h;
                        while (hh) {
                            int o[8]; uint32_t cd[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) { o[j] = -1; if (hh) { o[j] = lane * RL + (__ffs(hh) - 1); hh &= hh - 1u; } }
                            if (HASHED && exact_check) {
                                // large source sets: a filter positive is first checked against the exact hot bitmap
                                // (n/8 bytes, L2-resident) so that only true hits touch the code array (n bytes)
                                uint32_t src[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    src[j] = 0; cd[j] = 0;
                                    if (o[j] >= 0) { src[j] = (uint32_t)sc[o[j]]; cd[j] = __ldg(a.hot_bitmap_in + (src[j] >> 5)); }
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const bool hot = o[j] >= 0 && ((cd[j] >> (src[j] & 31u)) & 1u);
                                    cd[j] = hot ? (uint32_t)__ldg(code_b + src[j]) : 0u;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    cd[j] = 0;
                                    if (o[j] >= 0) {
                                        const uint32_t enc = (uint32_t)sc[o[j]];
                                        cd[j] = __ldg(code_b + (HASHED ? enc : (((enc >> 7) << 5) | (enc & 31u))));
                                    }
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                if (o[j] >= 0) {
                                    s_code[o[j]] = (uint8_t)cd[j];
                                    if (HASHED && cd[j] == 0u) h &= ~(1u << (o[j] - lane * RL));   // filter false positive
                                }
                            }
                        }
                    }
                    s_mask[lane] = h;
                    __syncwarp();
                    // fold: every lane walks the hit bits of ITS rows, in edge order
                    const int lo_all = max(rstart[0] - ci * CAP, 0), hi_all = min(rstart[ROWS] - ci * CAP, n);
                    if (my_sus != 0u && lo_all < hi_all) {
                        for (int w = lo_all / RL; w <= (hi_all - 1) / RL; ++w) {
                            uint32_t bits = s_mask[w];
                            const int b_lo = max(lo_all - w * RL, 0), b_hi = min(hi_all - w * RL, RL);
                            bits &= (b_hi >= 32 ? 0xffffffffu : ((1u << b_hi) - 1u)) & ~((1u << b_lo) - 1u);
                            while (bits) {
                                const int off = w * RL + (__ffs(bits) - 1);
                                bits &= bits - 1u;
                                const int code = (int)s_code[off];
                                if (ROWS > 1) {
                                    const int toff = off + ci * CAP;
                                    int r = 0;
#pragma unroll
                                    for (int q = 1; q < ROWS; ++q) r += (toff >= rstart[q]) ? 1 : 0;
                                    if (!((my_sus >> r) & 1u)) continue;
                                    if (r != cur_r) {
                                        flush_row();
                                        cur_r = r;
                                        const int v = load_age(a.age, row0 + r, age_f32);
                                        cur_ag = v < 0 ? 0 : (v >= m.n_sfs ? m.n_sfs - 1 : v);
                                    }
                                }
                                if (!weighted) {
                                    lam += s_T[cur_ag * NC + code];
                                    if (want_grad) lam1 += s_T1[cur_ag * NC + code];
                                } else {
                                    const float wgt = a.edge_weight[(int64_t)mt.base4 * 4 + ci * CAP + off];
                                    const float sa = m.sfs[cur_ag];
                                    lam += (double)__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(R, sa), s_A[code]), wgt), s_G[code]) / ibar;
                                    if (want_grad) lam1 += (double)__fmul_rn(__fmul_rn(__fmul_rn(sa, s_A[code]), wgt), s_G[code]) / ibar;
                                }
                            }
                        }
                    }
                }
                __syncwarp();                                        // slot and mask are reused by later chunks
            }
            if (MODE == MODE_GATHER) {
                flush_row();
                reinterpret_cast<uint8_t*>(a.dirty + (size_t)block * ntasks32)[(size_t)task * 32 + lane] = (uint8_t)dbits;
            }
        } else if (MODE == MODE_GATHER) {
            reinterpret_cast<uint8_t*>(a.dirty + (size_t)block * ntasks32)[(size_t)task * 32 + lane] = 0;
        }
        if (MODE == MODE_GATHER) continue;

        // ================================================================== update (lane <-> row)
        update_subtask(task, mt.sus, mt.exp, mt.inf, lam, lam1);
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
            for (int w = 0; w < nw; ++w) { s0 += s_red[w][0]; s1 += s_red[w][1]; s2 += s_red[w][2]; }
            if (s0) atomicAdd(&acc[0], (unsigned long long)s0);
            if (s1) atomicAdd(&acc[1], (unsigned long long)s1);
            if (s2) atomicAdd(&acc[2], (unsigned long long)s2);
            if (s_any_out) atomicOr(reinterpret_cast<unsigned int*>(ws + WS_ANYHOT) + a.xchg_slot, 1u);
            if (s_any_ei) atomicOr(reinterpret_cast<unsigned int*>(ws + WS_HASEI) + a.xchg_slot, 1u);
        }
        __threadfence();
        s_is_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_is_last && tid == 0) {
        __threadfence();
        if (MODE != MODE_GATHER) {
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
            if (a.over_state != nullptr && a.over_local) {
                // one GPU: this launch knows whether anybody is left in E or I ("epidemic over" lets the next gathers return at once)
                unsigned int* he = reinterpret_cast<unsigned int*>(ws + WS_HASEI) + a.xchg_slot;
                *a.over_state = (*reinterpret_cast<volatile unsigned int*>(he) != 0u) ? 0 : 1;
                *he = 0u;
            }
        }
        *ticket = 0;
    }
}

// ------------------------------------------------------------------------------- stream-layout gather
// AT_LAYOUT_STREAM: the contact graph as flat streams, one per source block (include/at_seirm.h).  A CTA owns a
// contiguous range of 256-entry units; for every source block the range touches it stages the block's EXACT hot
// bitmap in shared memory (one TMA bulk copy) and its warps stream the units through registers: 2 x 128-bit loads
// per lane and unit, three units in flight per warp, one shared-memory bitmap test (5 instructions) per edge.  The
// rare hits are parked (entry + unit) in a per-warp list; when the list fills up the warp resolves it in one sweep,
// every lane up to ST_RK hits and all loads of a phase in flight together: tile of the unit -> susceptible bit of the
// row -> infectiousness code of the source -> the pair (row in tile, code) is appended to the tile's RECORD list
// (slot from an integer atomic counter; capacity = the tile's edge count, so it cannot overflow).  Pure integer
// work: no table, no R, no floating point.  MODE_TUPDATE turns a tile's records into its rows' lambdas by summing
// 64-bit FIXED-POINT terms — integer addition commutes, so the result does not depend on the order in which warps,
// CTAs or steps of the schedule appended the records: deterministic without ordering edges, no float atomics.
#define ST_NT 1024                  // max threads per CTA
#define ST_PEND 128                 // parked hits per warp
#define ST_RK 4                     // hits per lane in one resolve sweep (ST_PEND = 32 * ST_RK)

struct StreamResolveCtx {
    const int32_t* unit_tile; const uint32_t* sus; const uint8_t* code_b; const int64_t* tile_rec_base;
    uint32_t* tile_count; uint32_t* rec;
    int tile_shift;
};
__device__ __noinline__ void stream_resolve(const StreamResolveCtx& c, const uint2* s_pend, int count, int lane)
{
    __syncwarp();
    uint32_t e[ST_RK], sw[ST_RK], cd[ST_RK];
    int tile[ST_RK];
    int64_t base[ST_RK];
#pragma unroll
    for (int k = 0; k < ST_RK; ++k) {
        const int i = k * 32 + lane;
        e[k] = 0u; tile[k] = -1;                         // (layers < 128, so a valid unit_tile word is never negative)
        if (i < count) {
            const uint2 p = s_pend[i];
            e[k] = p.x;
            tile[k] = __ldg(c.unit_tile + p.y);          // tile in the low 24 bits, network layer of the unit's edges above
        }
    }
#pragma unroll
    for (int k = 0; k < ST_RK; ++k) {
        sw[k] = 0u; base[k] = 0;
        if (tile[k] >= 0) {
            const int64_t row = ((int64_t)(tile[k] & 0xFFFFFF) << c.tile_shift) + (int64_t)((e[k] >> 5) & 0x7FFu);
            sw[k] = (__ldg(c.sus + (row >> 5)) >> (row & 31)) & 1u;
            base[k] = __ldg(c.tile_rec_base + (tile[k] & 0xFFFFFF));
        }
    }
#pragma unroll
    for (int k = 0; k < ST_RK; ++k) {
        cd[k] = 0u;
        if (sw[k]) cd[k] = __ldg(c.code_b + (((e[k] >> 16) << 5) | (e[k] & 31u)));    // only susceptible rows use their lambda
    }
#pragma unroll
    for (int k = 0; k < ST_RK; ++k) {
        if (cd[k] != 0u) {
            const uint32_t slot = atomicAdd(c.tile_count + (tile[k] & 0xFFFFFF), 1u);
            c.rec[base[k] + slot] = ((uint32_t)(tile[k] >> 24) << 24) | (((e[k] >> 5) & 0x7FFu) << 8) | cd[k];
        }
    }
    __syncwarp();
}

// Table of the step's fixed-point lambda terms, computed ONCE per step (by the last CTA of the gather launch, for the update
// launch that follows it in the stream): term[age][code] = f64(f32(((R*S)*A)*G)) / I_bar, the reference's message
// (transition.py:78-83), rounded once to the 2^-shift grid, shift chosen so that in-degree * largest term < 2^62.
// R == 0 (lambda == 0 everywhere): the terms of d lambda / dR instead, which is all the update then needs.
__device__ void stream_build_term_table(const KModel& m, const at_seirm_launch& a, int NC, double* s_scratch /* n_layers*n_sfs*NC */,
                                        unsigned long long* s_tmax)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const float R = a.r_value[a.r_index];
    char* ws = reinterpret_cast<char*>(a.workspace);
    long long* tab = reinterpret_cast<long long*>(ws + WS_FIXTAB);
    if (tid == 0) *s_tmax = 0ull;
    __syncthreads();
    double mymax = 0.0;
    const int per_layer = m.n_sfs * NC, n_tab = m.n_layers * per_layer;          // table[layer][age][code]
    for (int i = tid; i < n_tab; i += nt) {
        const int layer = i / per_layer, ag = (i % per_layer) / NC, c = i % NC;
        const double ibar = (double)m.ibar_layers[layer];
        float av = 0.f, gv = 0.f;
        if (c >= 1 && c <= 3 * m.n_lg) {
            const int sidx = (c - 1) / m.n_lg, idx = (c - 1) % m.n_lg;
            av = sidx < m.n_sfi ? m.sfi[sidx] : 0.f;
            gv = m.lg[idx];
        }
        const float s = m.sfs[ag];
        const double d = (R != 0.f) ? (double)__fmul_rn(__fmul_rn(__fmul_rn(R, s), av), gv) / ibar
                                    : (double)__fmul_rn(__fmul_rn(s, av), gv) / ibar;
        s_scratch[i] = d;
        const double ad = fabs(d);
        if (ad < 1e300 && ad > mymax) mymax = ad;
    }
    if (mymax > 0.0) atomicMax(s_tmax, (unsigned long long)__double_as_longlong(mymax));   // bit order == value order
    __syncthreads();
    int shift = 0;
    const double tmax = __longlong_as_double((long long)*s_tmax);
    if (tmax > 0.0) {
        int ex = 0;
        frexp(tmax, &ex);                                       // tmax < 2^ex
        shift = 62 - ex - a.lam_deg_bits;
        shift = shift > 1000 ? 1000 : (shift < -1000 ? -1000 : shift);
    }
    for (int i = tid; i < n_tab; i += nt) {
        const double d = s_scratch[i];                          // written by this thread
        tab[i] = (fabs(d) < 1e300) ? __double2ll_rn(ldexp(d, shift)) : 0ll;
    }
    if (tid == 0) *reinterpret_cast<int*>(ws + WS_FIXSHIFT) = shift;
    __syncthreads();
}

// UF: units in flight per warp (register sets).  The body is shared by the __global__ entry points below, which differ in their
// register budget.
template <int UF>
__device__ __forceinline__ void
stream_gather_body(const KModel& m, const at_seirm_launch& a, const int bm_words, const int NC, const int tile_shift)
{
    AT_DYNAMIC_SMEM(unsigned char, smem_raw, 128);
    const int nw = blockDim.x >> 5;
    // [hot bitmap of the current block: bm_words + 4 words, the last 4 always zero (padding entries point there)]
    // [mbarrier 16 B][parked hits uint2 nw*ST_PEND][scratch of the term table: 16 B + f64 n_sfs*NC]
    uint32_t* s_bm = reinterpret_cast<uint32_t*>(smem_raw);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)(bm_words + 4) * 4);
    uint2* s_pend_all = reinterpret_cast<uint2*>(s_bar + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint2* s_pend = s_pend_all + warp * ST_PEND;
    // the epidemic is over (nobody of any rank in E or I): no source is infectious, the update has no record to read
    if (a.over_state != nullptr && *reinterpret_cast<const volatile int*>(a.over_state) > 0) return;
    if (blockIdx.x == gridDim.x - 1) {                          // scratch: after the parked-hit lists
        unsigned char* scratch = reinterpret_cast<unsigned char*>(s_pend_all + (size_t)nw * ST_PEND);
        stream_build_term_table(m, a, NC, reinterpret_cast<double*>(scratch + 16), reinterpret_cast<unsigned long long*>(scratch));
    }
    if (tid == 0) mbar_init(s_bar, 1);
    if (tid < 4) s_bm[bm_words + tid] = 0u;
    fence_mbarrier_init();

    // ---- this CTA's units: an equal share of all units, whatever blocks it touches
    const int64_t n_units = a.block_units[a.n_blocks];
    const int64_t U0 = n_units * (int64_t)blockIdx.x / (int64_t)gridDim.x, U1 = n_units * ((int64_t)blockIdx.x + 1) / (int64_t)gridDim.x;
    const int4* colv = reinterpret_cast<const int4*>(a.col);
    const int block_words = a.block_size >> 5;
    uint32_t bar_phase = 0;
    int count = 0;                                                // parked hits of this warp (warp-uniform)
    const uint32_t lt_mask = (1u << lane) - 1u;

    StreamResolveCtx rc;
    rc.unit_tile = a.unit_tile; rc.sus = a.sus_bitmap; rc.code_b = a.code_in; rc.tile_rec_base = a.tile_rec_base;
    rc.tile_count = a.tile_count; rc.rec = a.rec; rc.tile_shift = tile_shift;
    int block = 0;
    while (block < a.n_blocks && (int64_t)a.block_units[block + 1] <= U0) ++block;
    for (; block < a.n_blocks && (int64_t)a.block_units[block] < U1; ++block) {
        const int64_t s0 = max(U0, (int64_t)a.block_units[block]), s1 = min(U1, (int64_t)a.block_units[block + 1]);
        if (s0 >= s1) continue;
        // ---- stage the block's hot bitmap (TMA bulk copy, one mbarrier phase per block)
        __syncthreads();                                         // barrier initialised; previous block's bitmap no longer read
        // (the last block may reach past the end of the source id space: those bitmap words are zero-filled here, not read)
        const int valid_words = (int)min((int64_t)block_words, (a.n_sources >> 5) - (int64_t)block * block_words);
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)valid_words * 4u;
            fence_proxy_async_smem();
            mbar_expect_tx(s_bar, bytes);
            bulk_g2s(s_bm, a.hot_bitmap_in + (size_t)block * block_words, bytes, s_bar);
        }
        for (int i = valid_words + tid; i < block_words; i += blockDim.x) s_bm[i] = 0u;
        mbar_wait(s_bar, bar_phase & 1u);
        ++bar_phase;
        uint32_t seen = 0;
        {
            const uint4* b4 = reinterpret_cast<const uint4*>(s_bm);
            for (int i = tid; i < (valid_words >> 2); i += blockDim.x) { const uint4 v = b4[i]; seen |= v.x | v.y | v.z | v.w; }
        }
        if (!__syncthreads_or(seen != 0u)) continue;             // no infectious source in this block: nothing can hit
                                                                 // (the barrier also orders the zero-fill above before the tests)
        rc.code_b = a.code_in + (size_t)block * a.block_size;
        auto resolve = [&]() { stream_resolve(rc, s_pend, count, lane); count = 0; };
        // edge test: byte offset of the bitmap word from the top 16 bits, bit index in the low 5 (the shift uses them alone)
#define ST_TEST(E) h = __funnelshift_r(h, s_bm[(uint32_t)(E) >> 16] >> ((uint32_t)(E) & 31u), 1)
        const int my_units = (s0 + warp < s1) ? (int)((s1 - s0 - warp + nw - 1) / nw) : 0;
        const int4* p = colv + (s0 + warp) * (AT_STREAM_UNIT / 4) + lane;
        const size_t pstep = (size_t)nw * (AT_STREAM_UNIT / 4);
        uint32_t ucur = (uint32_t)(s0 + warp);                  // unit being processed
        const int4* pn = p;                                     // next unit to request
        int to_load = my_units;
        auto load_next = [&](int4& x0, int4& x1) {              // streaming loads (evict-first), the warp's units in order
            if (to_load > 0) {
                x0 = __ldcs(pn);
                x1 = __ldcs(pn + 32);
                pn += pstep;
            }
            --to_load;
        };
        auto process = [&](const int4 x0, const int4 x1) {
            uint32_t h = 0;
            ST_TEST(x0.x); ST_TEST(x0.y); ST_TEST(x0.z); ST_TEST(x0.w);
            ST_TEST(x1.x); ST_TEST(x1.y); ST_TEST(x1.z); ST_TEST(x1.w);
            uint32_t any = __ballot_sync(0xffffffffu, h != 0u);
            if (any) {
                h >>= 24;                                        // bits 0..3: entries of x0, bits 4..7: entries of x1
                do {
                    if (count > ST_PEND - 32) resolve();
                    if (h != 0u) {
                        const int b = __ffs((int)h) - 1;
                        h &= h - 1u;
                        const int4 xs = (b & 4) ? x1 : x0;
                        const int e = (b & 2) ? ((b & 1) ? xs.w : xs.z) : ((b & 1) ? xs.y : xs.x);
                        s_pend[count + __popc(any & lt_mask)] = make_uint2((uint32_t)e, ucur);
                    }
                    count += __popc(any);
                    any = __ballot_sync(0xffffffffu, h != 0u);
                } while (any);
            }
            ucur += (uint32_t)nw;
        };
        // UF register sets: slot u is processed, then refilled with the unit UF places further on (indices are static
        // after unrolling, so the sets live in registers)
        int4 x0[UF], x1[UF];
#pragma unroll
        for (int u = 0; u < UF; ++u) { x0[u] = make_int4(0, 0, 0, 0); x1[u] = x0[u]; load_next(x0[u], x1[u]); }
        for (int k = my_units; k > 0; k -= UF) {
#pragma unroll
            for (int u = 0; u < UF; ++u) {
                if (u < k) {
                    process(x0[u], x1[u]);
                    load_next(x0[u], x1[u]);
                }
            }
        }
#undef ST_TEST
        if (count > 0) resolve();                                // the code array of the next block is another one
    }
}

// NTHREADS: CTA size the entry point is compiled for (its register budget)
template <int NTHREADS, int UF>
__global__ void __launch_bounds__(NTHREADS, 1)
stream_gather_kernel(const KModel m, const at_seirm_launch a, const int bm_words, const int NC, const int tile_shift)
{
    stream_gather_body<UF>(m, a, bm_words, NC, tile_shift);
}
// The "lean" entry point: 1024 threads with at most 56 registers each, which leaves 8 K registers of every SM to a 256-thread
// CTA of another stream — the overlapped halo exchange of the next-but-one step runs beside it instead of waiting for it.
#ifndef AT_HOST_EMULATION
__global__ void __maxnreg__(56)
#else
__global__ void
#endif
stream_gather_lean_kernel(const KModel m, const at_seirm_launch a, const int bm_words, const int NC, const int tile_shift)
{
    stream_gather_body<2>(m, a, bm_words, NC, tile_shift);
}

// ------------------------------------------------------------------------------- stream-layout update
// Second kernel of a stream-layout step: transmission decision + progression + encoding of the next step
// (reference: transition.py:219-252, seirm_progression/transition.py:41-121).  A CTA takes whole row tiles and works in
// phases that each touch only the rows that need them (most rows of a step need nothing but their zeroed outputs):
//   0  zero-fill the tile's dne_dr / code_out / hot words with 128-bit stores; stage the tile's row bitmaps and (128-bit loads)
//      susceptibility classes in shared memory.  The tile's record count and base were fetched one tile ahead and each
//      thread's first two records are already in flight, so the dependent chain count -> records -> table costs no
//      round trip of its own
//   1  sum the tile's hit records (row in tile, source code) into one 64-bit FIXED-POINT lambda per row in shared memory
//      (integer atomics: the order of the records does not matter, the sum is bitwise reproducible; scale 2^fx_shift chosen
//      so that in-degree * largest term < 2^62)
//   A  the exposure candidates (rows that received a record) are decided densely, one candidate per thread, whatever
//      their position in the tile: noise, p = 1 - exp(-lambda), stage / timer writes, d ne / dR
//   B  the tile's rows in E or I, compacted into a list (like the candidates): progression and the code of the next step,
//      one row per thread, all timer loads of the tile in flight together
// Measured on the 8.4 M-agent hot step (gather + update, float32 state): per-task phase B, scalar age loads, no prefetch
// 254 us -> 201 us (profiles/r02_step_kernels_8p4M.md).
// The per-row logic is the one of seirm_kernel's update_subtask (the generic kernel), restated for launches with
// transmission (row bitmaps valid, no AT_DO_INIT).
#ifndef TU_THREADS
#define TU_THREADS 256
#endif
#define TU_MAX_TASKS (AT_STREAM_MAX_TILE_ROWS / 32)
#ifndef TU_MIN_CTAS
#define TU_MIN_CTAS 6               // resident update CTAs per SM the register budget is cut for (40 registers; measured: 8 -> 6 CTAs: -5 % per hot step)
#endif
template <bool TF32, bool GRAD>
__global__ void __launch_bounds__(TU_THREADS, TU_MIN_CTAS)
stream_update_kernel(const KModel m, const at_seirm_launch a, const int NC)
{
    AT_DYNAMIC_SMEM(unsigned char, smem_raw, 128);
    // [s_fix i64 n_layers*n_sfs*NC][s_A f32 256][s_G f32 256][s_lam i64 tile_rows][s_lamE i64 tile_rows, if dne_e_dr][s_cand u16 tile_rows]
    // [s_age u8 tile_rows]
    const int n_tab = m.n_layers * m.n_sfs * NC;
    const bool want_e = GRAD && a.dne_e_dr != nullptr;          // split d ne / dR by the stage of the sources (for d / d SFInfector)
    long long* s_fix = reinterpret_cast<long long*>(smem_raw);
    float* s_A = reinterpret_cast<float*>(s_fix + n_tab);
    float* s_G = s_A + 256;
    long long* s_lam = reinterpret_cast<long long*>(s_G + 256);
    long long* s_lamE = s_lam + a.tile_rows;                     // lambda from sources in stage E
    uint16_t* s_cand = reinterpret_cast<uint16_t*>(s_lam + (want_e ? 2 : 1) * a.tile_rows);
    uint8_t* s_age = reinterpret_cast<uint8_t*>(s_cand + a.tile_rows);
    __shared__ long long s_red[TU_THREADS / 32][3];
    __shared__ uint32_t s_w[3][TU_MAX_TASKS], s_orig[3][TU_MAX_TASKS], s_hot[TU_MAX_TASKS];   // sus / exp / inf words of the tile
    __shared__ int s_is_last, s_any_out, s_any_ei, s_ncand, s_nei;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = TU_THREADS / 32;
    const bool age_f32 = a.age_dtype == AT_AGE_F32;
    const bool do_prog = (a.flags & AT_DO_PROGRESSION) != 0, do_enc = (a.flags & AT_DO_ENCODE) != 0;
    const bool lookahead = (a.flags & AT_DO_LOOKAHEAD) != 0;
    const int t = a.t;
    const float R = a.r_value[a.r_index];
    const uint64_t seed = a.seed_ptr ? *a.seed_ptr : a.seed;
    char* ws = reinterpret_cast<char*>(a.workspace);

    if (tid == 0) { s_any_out = 0; s_any_ei = 0; s_ncand = 0; s_nei = 0; }
    // epidemic over on entry (nobody of any rank in E or I after the previous step): no record, no candidate, no progression.
    // The (T, N) saved-for-backward rows of this step are then left untouched (step_skipped tells the backward pass) and,
    // from the second such step on, so are the code / hot buffers (both of them are zero by then).
    const int over_in = (a.over_state != nullptr) ? *reinterpret_cast<const volatile int*>(a.over_state) : 0;
    const bool skip_saved = over_in >= 1 && a.step_skipped != nullptr;
    const bool skip_codes = over_in >= 2;
    const bool idle = over_in >= 1 && skip_codes && (skip_saved || !GRAD);      // nothing to do for any tile: no tables either
    if (blockIdx.x == 0 && tid == 0 && a.step_skipped != nullptr) a.step_skipped[t] = skip_saved ? 1 : 0;
    if (idle) {
        // nothing to do and nothing to publish but zeros: no tables, no tiles, no ticket (888 CTAs queueing for one atomic cost more
        // than everything else an idle launch does).  over_state is >= 2 whichever CTA reads it when, so CTA 0 may advance it alone.
        if (blockIdx.x == 0 && tid == 0) {
            if (a.sum_dne_dr) a.sum_dne_dr[t] = 0.f;
            if (do_prog && a.death_count) a.death_count[t] = 0.f;
            if (a.over_state != nullptr && a.over_local) {
                *a.over_state = over_in + 1;
                reinterpret_cast<unsigned int*>(ws + WS_HASEI)[a.xchg_slot] = 0u;
            }
        }
        return;
    }
    int fx_shift = 0;
    {
        for (int c = tid; c < 256; c += TU_THREADS) {
            float av = 0.f, gv = 0.f;
            if (c >= 1 && c <= 3 * m.n_lg) {
                const int sidx = (c - 1) / m.n_lg, idx = (c - 1) % m.n_lg;
                av = sidx < m.n_sfi ? m.sfi[sidx] : 0.f;
                gv = m.lg[idx];
            }
            s_A[c] = av;
            s_G[c] = gv;
        }
        // fixed-point term table of this step, built by the gather launch (stream_build_term_table)
        const long long* tab = reinterpret_cast<const long long*>(ws + WS_FIXTAB);
        for (int i = tid; i < n_tab; i += TU_THREADS) s_fix[i] = tab[i];
        fx_shift = *reinterpret_cast<const int*>(ws + WS_FIXSHIFT);
    }
    __syncthreads();
    const double inv_scale = ldexp(1.0, -fx_shift);             // exact power of two
    const double inv_r = (R != 0.f) ? 1.0 / (double)R : 1.0;

    const int tile_rows = a.tile_rows, tpt = tile_rows >> 5;   // 32-row tasks per tile
    const int n_tiles = (int)(a.n_rows_pad / tile_rows);
    const int n_rows = (int)a.n_rows;
    uint8_t* code_out = a.code_out + a.row_offset;
    uint32_t* hot_out = a.hot_bitmap_out + (a.row_offset >> 5);
    const uint64_t agent0 = (uint64_t)a.row_offset;
    long long acc_ne = 0, acc_dne = 0, acc_dead = 0;

    auto ld_timer = [&](const void* p, int row) -> int {
        return TF32 ? (int)reinterpret_cast<const float*>(p)[row] : (int)reinterpret_cast<const int16_t*>(p)[row];
    };
    auto st_timer = [&](void* p, int row, int v) {
        if (TF32) reinterpret_cast<float*>(p)[row] = (float)v; else reinterpret_cast<int16_t*>(p)[row] = (int16_t)v;
    };
    // infectiousness code of step t_encode for a row in stage class sidx infected at itime (0: contributes nothing)
    auto encode = [&](int sidx, int itime) -> uint32_t {
        int idx = a.t_encode - itime - 1;
        if (idx < 0 || idx >= m.n_lg) {
            if (a.error_flag && sidx < m.n_sfi && m.sfi[sidx] != 0.f) *a.error_flag = 1;
            idx = idx < 0 ? 0 : m.n_lg - 1;
        }
        uint32_t code = 1u + (uint32_t)(sidx * m.n_lg + idx);
        if (s_A[code] == 0.f || s_G[code] == 0.f) code = 0;       // contributes exactly 0 to every lambda
        return code;
    };

    // a tile waits for its record count, then for its records: both are fetched early (count and base one tile ahead)
    uint32_t nrec_next = 0u;
    int64_t base_next = 0;
    if ((int)blockIdx.x < n_tiles && over_in < 1) { nrec_next = a.tile_count[blockIdx.x]; base_next = a.tile_rec_base[blockIdx.x]; }
    for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x) {
        const int row_base = tile * tile_rows;
        const int rows_here = min(tile_rows, n_rows - row_base);        // valid rows of the tile (<= 0: padding only)
        const uint32_t nrec = nrec_next;
        const uint32_t* recs = a.rec + base_next;
        uint32_t rec0 = 0u, rec1 = 0u;
        if ((uint32_t)tid < nrec) rec0 = recs[tid];
        if ((uint32_t)tid + TU_THREADS < nrec) rec1 = recs[tid + TU_THREADS];
        const int next_tile = tile + (int)gridDim.x;
        nrec_next = 0u;
        if (next_tile < n_tiles && over_in < 1) { nrec_next = a.tile_count[next_tile]; base_next = a.tile_rec_base[next_tile]; }
        // ---- phase 0: outputs of rows that need nothing, bitmaps of the tile
        if (GRAD && !skip_saved) {
            float4* d4 = reinterpret_cast<float4*>(a.dne_dr + row_base);
            float4* e4 = want_e ? reinterpret_cast<float4*>(a.dne_e_dr + row_base) : nullptr;
            for (int i = tid; i * 4 < rows_here; i += TU_THREADS) {
                if (i * 4 + 4 <= rows_here) {
                    d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (want_e) e4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                } else for (int r = i * 4; r < rows_here; ++r) {
                    a.dne_dr[row_base + r] = 0.f;
                    if (want_e) a.dne_e_dr[row_base + r] = 0.f;
                }
            }
        }
        if (do_enc && !skip_codes) {
            uint4* c4 = reinterpret_cast<uint4*>(code_out + row_base);
            for (int i = tid; i * 16 < rows_here; i += TU_THREADS) {
                if (i * 16 + 16 <= rows_here) c4[i] = make_uint4(0u, 0u, 0u, 0u);
                else for (int r = i * 16; r < rows_here; ++r) code_out[row_base + r] = 0;
            }
        }
        if (tid < tpt) {
            const int task = tile * tpt + tid;
            const uint32_t sw = a.sus_bitmap[task], ew = a.exp_bitmap[task], iw = a.inf_bitmap[task];
            s_w[0][tid] = sw; s_w[1][tid] = ew; s_w[2][tid] = iw;
            s_orig[0][tid] = sw; s_orig[1][tid] = ew; s_orig[2][tid] = iw;
            s_hot[tid] = 0u;
            if (do_enc && !skip_codes && task * 32 < n_rows) hot_out[task] = 0u;     // (words of padding rows belong to somebody else)
        }
        if (nrec != 0u) {
            const bool full = rows_here == tile_rows;
            if (full && age_f32 && ((reinterpret_cast<uintptr_t>(a.age) & 15) == 0)) {
                const float4* a4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.age) + row_base);
                uint32_t* s_age4 = reinterpret_cast<uint32_t*>(s_age);
                const int hi = m.n_sfs - 1;
                for (int i = tid; i * 4 < tile_rows; i += TU_THREADS) {
                    const float4 v = a4[i];
                    const int a0 = min(max((int)v.x, 0), hi), a1 = min(max((int)v.y, 0), hi);
                    const int a2 = min(max((int)v.z, 0), hi), a3 = min(max((int)v.w, 0), hi);
                    s_age4[i] = (uint32_t)a0 | ((uint32_t)a1 << 8) | ((uint32_t)a2 << 16) | ((uint32_t)a3 << 24);
                }
                for (int i = tid; i < tile_rows; i += TU_THREADS) { s_lam[i] = 0ll; if (want_e) s_lamE[i] = 0ll; }
            } else if (full && !age_f32 && ((reinterpret_cast<uintptr_t>(a.age) & 15) == 0)) {
                const uint4* a4 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(a.age) + row_base);
                uint4* s_age16 = reinterpret_cast<uint4*>(s_age);
                const uint32_t hi4 = (uint32_t)(m.n_sfs - 1) * 0x01010101u;
                for (int i = tid; i * 16 < tile_rows; i += TU_THREADS) {
                    uint4 v = a4[i];
                    v.x = __vminu4(v.x, hi4); v.y = __vminu4(v.y, hi4); v.z = __vminu4(v.z, hi4); v.w = __vminu4(v.w, hi4);
                    s_age16[i] = v;
                }
                for (int i = tid; i < tile_rows; i += TU_THREADS) { s_lam[i] = 0ll; if (want_e) s_lamE[i] = 0ll; }
            } else
            for (int i = tid; i < tile_rows; i += TU_THREADS) {
                s_lam[i] = 0ll;
                if (want_e) s_lamE[i] = 0ll;
                int ag = (i < rows_here) ? load_age(a.age, row_base + i, age_f32) : 0;     // susceptibility class of the row
                s_age[i] = (uint8_t)(ag < 0 ? 0 : (ag >= m.n_sfs ? m.n_sfs - 1 : ag));
            }
        }
        __syncthreads();
        if (nrec != 0u) {
            // ---- phase 1: records -> fixed-point lambda per row
            auto add_record = [&](const uint32_t rec) {       // (layer << 24) | (row in tile << 8) | code
                const int r = (int)((rec >> 8) & 0x7FFu), code = (int)(rec & 255u);
                const long long v = s_fix[((int)(rec >> 24) * m.n_sfs + (int)s_age[r]) * NC + code];
                if (v != 0ll) {
                    // code = 1 + stage * n_lg + day: stage 1 = exposed (SFInfector[1])
                    if (want_e && (code - 1) / m.n_lg == 1) atomicAdd(reinterpret_cast<unsigned long long*>(s_lamE) + r, (unsigned long long)v);
                    // ---- phase A, part 1: the row's first record makes it an exposure candidate (records exist for susceptible rows
                    //      only, and all terms of a step have the sign of R, so the sum was zero exactly before the first one)
                    const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(s_lam) + r, (unsigned long long)v);
                    if (old == 0ull) s_cand[atomicAdd(&s_ncand, 1)] = (uint16_t)r;
                }
            };
            if ((uint32_t)tid < nrec) add_record(rec0);
            if ((uint32_t)tid + TU_THREADS < nrec) add_record(rec1);
            for (uint32_t i = tid + 2 * TU_THREADS; i < nrec; i += TU_THREADS) add_record(recs[i]);
            __syncthreads();
            if (tid == 0) a.tile_count[tile] = 0u;
            const int ncand = s_ncand;
            for (int i = tid; i < ncand; i += TU_THREADS) {
                const int r = (int)s_cand[i];
                const int row = row_base + r;
                const double v = (double)s_lam[r] * inv_scale;
                const double lam = (R != 0.f) ? v : 0.0, lam1 = v * inv_r;
                float wiso = 0.f, u = 0.f;
                if (a.iso_mode == AT_ISO_TENSOR) wiso = a.isolation[row];
                if (a.u_mode == AT_U_TENSOR) u = a.uniform[row];
                if (a.iso_mode != AT_ISO_TENSOR || a.u_mode != AT_U_TENSOR) {
                    const uint64_t z = at_noise_u64(seed, (uint32_t)t, agent0 + (uint64_t)row);
                    if (a.u_mode != AT_U_TENSOR) u = at_u_exposure(z);
                    if (a.iso_mode == AT_ISO_HASH_SOFT) wiso = at_u_isolation(z);
                    else if (a.iso_mode == AT_ISO_HASH_HARD) wiso = at_u_isolation(z) < a.p_isolate ? 1.0f : 0.0f;
                }
                // p = 1 - exp(-lambda) (transition.py:220-221); u < p decided in float32 when it is far from the threshold
                // (|error of p32| < 4e-6 for lambda < 88), in float64 otherwise: the decisions of the float64 reference
                const float q32 = __expf(-(float)lam);
                const float p32 = 1.0f - q32;
                bool hit;
                if (fabsf(u - p32) > 2e-5f) hit = u < p32;
                else hit = (double)u < 1.0 - exp(-lam);
                const float omw = 1.0f - wiso;                  // float32, as (1.0 - will_isolate)
                const float ne = hit ? omw : 0.0f;
                if (GRAD) {
                    const float dne = omw * q32 * (float)lam1;
                    acc_dne += __double2ll_rn((double)dne * FX_SCALE);
                    a.dne_dr[row] = dne;
                    if (want_e) a.dne_e_dr[row] = (float)((double)dne * ((double)s_lamE[r] / (double)s_lam[r]));
                }
                if (ne != 0.0f) {
                    a.stage[row] = ne;                         // 0 + ne
                    st_timer(a.next_stage_time, row, t + 1 + m.e2i);
                    st_timer(a.infected_time, row, t);
                    acc_ne += __double2ll_rn((double)ne * FX_SCALE);
                    const uint32_t bit = 1u << (r & 31);
                    atomicAnd(&s_w[0][r >> 5], ~bit);
                    if (ne == 1.0f) atomicOr(&s_w[1][r >> 5], bit);
                    else if (do_enc && ne > 0.0f && ne < 3.0f) {
                        // fractional stage (soft isolation, Q4): indexes SFInfector[int(stage)], never progresses
                        const uint32_t code = encode((int)ne, t);
                        if (code != 0u) { code_out[row] = (uint8_t)code; atomicOr(&s_hot[r >> 5], bit); }
                    }
                }
            }
            __syncthreads();
            if (tid == 0) s_ncand = 0;
        }
        // ---- phase B: the tile's rows in E or I, compacted into a list so that every row has its own thread and all timer loads
        //      of the tile are in flight together (timers written in phase A are visible: same CTA, after the barrier)
        if (tid < tpt) {
            const uint32_t ew = s_w[1][tid];
            uint32_t ei = ew | s_w[2][tid];
            if (ei != 0u) {
                int at = atomicAdd(&s_nei, __popc(ei));
                while (ei != 0u) {
                    const int b = __ffs((int)ei) - 1;
                    ei &= ei - 1u;
                    s_cand[at++] = (uint16_t)((tid * 32 + b) | (((ew >> b) & 1u) << 15));     // row in tile (< 2048), bit 15: in E
                }
            }
        }
        __syncthreads();
        {
            const int nei = s_nei;
            for (int i = tid; i < nei; i += TU_THREADS) {
                const int entry = (int)s_cand[i];
                const int r = entry & 0x7FF;
                const int row = row_base + r;
                const uint32_t bit = 1u << (r & 31);
                bool is_e = (entry >> 15) != 0, is_i = !is_e;
                int nst = (do_prog || lookahead) ? ld_timer(a.next_stage_time, row) : 0;     // both timers in flight together
                const int itime = do_enc ? ld_timer(a.infected_time, row) : 0;
                if (do_prog && nst <= t) {                   // progression (seirm_progression/transition.py:41-88)
                    if (is_i) {
                        acc_dead += 1;                      // stage * rd / INFECTED_VAR = 2/2
                        if (a.recovered_step) a.recovered_step[row] = (uint8_t)t;
                        atomicAnd(&s_w[2][r >> 5], ~bit);                             // I -> R
                    } else {
                        atomicAnd(&s_w[1][r >> 5], ~bit);                             // E -> I
                        atomicOr(&s_w[2][r >> 5], bit);
                    }
                    if (nst == t) { nst = is_i ? m.inf_time : t + m.i2r; st_timer(a.next_stage_time, row, nst); }
                    a.stage[row] = is_i ? 3.0f : 2.0f;
                    const bool was_e = is_e;
                    is_e = false;
                    is_i = was_e;
                }
                bool ce = is_e, ci = is_i;                       // classes the code is computed for
                if (lookahead && (is_e || is_i) && nst <= a.t_encode - 1) { ci = ce; ce = false; }   // after the progression of t_encode-1
                if (do_enc && (ce || ci)) {
                    const uint32_t code = encode(ce ? 1 : 2, itime);
                    if (code != 0u) { code_out[row] = (uint8_t)code; atomicOr(&s_hot[r >> 5], bit); }
                }
            }
        }
        __syncthreads();
        if (tid == 0) s_nei = 0;
        if (tid < tpt && do_enc && s_hot[tid] != 0u) { hot_out[tile * tpt + tid] = s_hot[tid]; s_any_out = 1; }   // benign race
        // ---- the tile's bitmap words that changed; hot words of tasks without a hot row
        if (tid < tpt) {
            const int task = tile * tpt + tid;
            if (s_w[0][tid] != s_orig[0][tid]) a.sus_bitmap[task] = s_w[0][tid];
            if (s_w[1][tid] != s_orig[1][tid]) a.exp_bitmap[task] = s_w[1][tid];
            if (s_w[2][tid] != s_orig[2][tid]) a.inf_bitmap[task] = s_w[2][tid];
            if ((s_w[1][tid] | s_w[2][tid]) != 0u) s_any_ei = 1;  // benign race: every writer stores 1
        }
        __syncthreads();                                         // shared arrays are reused by the next tile
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
    int* ticket = reinterpret_cast<int*>(ws + WS_TICKET);
    if (tid == 0) {
        long long s0 = 0, s1 = 0, s2 = 0;
        for (int w = 0; w < nw; ++w) { s0 += s_red[w][0]; s1 += s_red[w][1]; s2 += s_red[w][2]; }
        if (s0) atomicAdd(&acc[0], (unsigned long long)s0);
        if (s1) atomicAdd(&acc[1], (unsigned long long)s1);
        if (s2) atomicAdd(&acc[2], (unsigned long long)s2);
        if (s_any_out) atomicOr(reinterpret_cast<unsigned int*>(ws + WS_ANYHOT) + a.xchg_slot, 1u);
        if (s_any_ei) atomicOr(reinterpret_cast<unsigned int*>(ws + WS_HASEI) + a.xchg_slot, 1u);
        __threadfence();
        s_is_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_is_last && tid == 0) {
        __threadfence();
        const volatile long long* av = reinterpret_cast<const volatile long long*>(acc);
        const double s0 = (double)av[0] / FX_SCALE, s1 = (double)av[1] / FX_SCALE, s2 = (double)av[2];
        if (a.daily_infected) a.daily_infected[t] += (float)s0;          // transition.py:235
        if (a.sum_dne_dr) a.sum_dne_dr[t] = (float)s1;
        if (do_prog) {
            if (a.death_count) a.death_count[t] = (float)s2;
            if (a.daily_deaths) a.daily_deaths[t] += (float)s2 * a.m_value[0];   // progression :52-60
        }
        acc[0] = 0; acc[1] = 0; acc[2] = 0;
        if (a.over_state != nullptr && a.over_local) {
            // one GPU: nobody left in E or I => nothing is or will become infectious; the next gathers return at once
            unsigned int* he = reinterpret_cast<unsigned int*>(ws + WS_HASEI) + a.xchg_slot;
            *a.over_state = (*reinterpret_cast<volatile unsigned int*>(he) != 0u) ? 0 : over_in + 1;
            *he = 0u;
        }
        *ticket = 0;
    }
}

// ------------------------------------------------------------------------------------ initial encode
// flags == AT_DO_ENCODE | AT_DO_INIT, the first launch of every episode: the S / E / I row bitmaps from the stage array, the
// infectiousness code and hot bitmap of step t_encode (same results as the generic kernel's update-only pass).  One streaming
// pass: a warp takes 128 rows per round, four independent 128-byte loads of the stage array in flight per warp; the
// timers are read only for rows that are infectious (E, I or a fractional stage, Q4).
template <bool TF32>
__global__ void __launch_bounds__(256)
init_encode_kernel(KModel m, at_seirm_launch a)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int s_any_out, s_any_ei;
    if (tid == 0) { s_any_out = 0; s_any_ei = 0; }
    __syncthreads();
    char* ws = reinterpret_cast<char*>(a.workspace);
    const int ntasks32 = (int)(a.n_rows_pad >> 5);
    const int nquad = (ntasks32 + 3) >> 2;
    for (int q = (int)blockIdx.x * 8 + warp; q < nquad; q += (int)gridDim.x * 8) {
        float st[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t row = ((int64_t)q * 4 + j) * 32 + lane;
            st[j] = row < a.n_rows ? a.stage[row] : 3.0f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int task = q * 4 + j;
            const int64_t row = (int64_t)task * 32 + lane;
            const bool valid = row < a.n_rows;
            const float stage = st[j];
            const bool is_s = stage == 0.0f, is_e = stage == 1.0f, is_i = stage == 2.0f;
            // infectious = 0 < stage < 3; fractional stages (soft isolation, Q4) index SFInfector[0]
            const bool frac = stage > 0.0f && stage < 3.0f && !(is_e || is_i);
            uint32_t code = 0;
            if (is_e || is_i || frac) {
                const int itime = load_timer(a.infected_time, row, TF32);
                const int sidx = is_e ? 1 : (is_i ? 2 : (int)stage);
                int idx = a.t_encode - itime - 1;
                if (idx < 0 || idx >= m.n_lg) {
                    if (a.error_flag && sidx < m.n_sfi && m.sfi[sidx] != 0.f) *a.error_flag = 1;
                    idx = idx < 0 ? 0 : m.n_lg - 1;
                }
                code = 1u + (uint32_t)(sidx * m.n_lg + idx);
                const float av = sidx < m.n_sfi ? m.sfi[sidx] : 0.f;
                if (av == 0.f || m.lg[idx] == 0.f) code = 0;            // contributes exactly 0 to every lambda
            }
            const uint32_t hot = __ballot_sync(0xffffffffu, code != 0u);
            const uint32_t sus_o = __ballot_sync(0xffffffffu, is_s);
            const uint32_t exp_o = __ballot_sync(0xffffffffu, is_e);
            const uint32_t inf_o = __ballot_sync(0xffffffffu, is_i);
            if (valid) a.code_out[a.row_offset + row] = (uint8_t)code;
            if (lane == 0 && task < ntasks32) {
                if ((int64_t)task * 32 < a.n_rows) {                    // (hot words of padding tasks belong to somebody else)
                    a.hot_bitmap_out[(a.row_offset >> 5) + task] = hot;
                    if (hot != 0u) s_any_out = 1;                       // benign race: every writer stores 1
                }
                if ((exp_o | inf_o) != 0u) s_any_ei = 1;
                a.sus_bitmap[task] = sus_o;
                a.exp_bitmap[task] = exp_o;
                a.inf_bitmap[task] = inf_o;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (s_any_out) atomicOr(reinterpret_cast<unsigned int*>(ws + WS_ANYHOT) + a.xchg_slot, 1u);
        if (s_any_ei) atomicOr(reinterpret_cast<unsigned int*>(ws + WS_HASEI) + a.xchg_slot, 1u);
        __threadfence();
        int* ticket = reinterpret_cast<int*>(ws + WS_TICKET);
        if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
            __threadfence();
            if (a.over_state != nullptr && a.over_local) {
                unsigned int* he = reinterpret_cast<unsigned int*>(ws + WS_HASEI) + a.xchg_slot;
                *a.over_state = (*reinterpret_cast<volatile unsigned int*>(he) != 0u) ? 0 : 1;
                *he = 0u;
            }
            *ticket = 0;
        }
    }
}

// Folds the hot bitmap of the next step into the hashed filter (word i = OR of words i, i+M, i+2M, ...) once, so that
// the next step's CTAs read M words instead of the whole bitmap each; the word after the filter says "any source hot".
__global__ void __launch_bounds__(256)
hot_fold_kernel(const uint32_t* __restrict__ hot, uint32_t* __restrict__ filter, int hash_words, int total4)
{
    // blockIdx.y picks every gridDim.y-th slice, so the (latency-bound) loads of the slices run in parallel; the
    // filter was zeroed by the caller and is assembled with atomicOr (bitwise OR: order-independent)
    const int M4 = hash_words >> 2;
    const uint4* src = reinterpret_cast<const uint4*>(hot);
    uint32_t seen = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M4; i += gridDim.x * blockDim.x) {
        uint4 acc4 = make_uint4(0, 0, 0, 0);
        for (int k = i + (int)blockIdx.y * M4; k < total4; k += (int)gridDim.y * M4) {
            const uint4 v = __ldg(src + k);
            acc4.x |= v.x; acc4.y |= v.y; acc4.z |= v.z; acc4.w |= v.w;
        }
        if (acc4.x) atomicOr(&filter[4 * i + 0], acc4.x);
        if (acc4.y) atomicOr(&filter[4 * i + 1], acc4.y);
        if (acc4.z) atomicOr(&filter[4 * i + 2], acc4.z);
        if (acc4.w) atomicOr(&filter[4 * i + 3], acc4.w);
        seen |= acc4.x | acc4.y | acc4.z | acc4.w;
    }
    if (__syncthreads_or(seen != 0u) && threadIdx.x == 0) atomicOr(&filter[hash_words], 1u);
}

// Multi-GPU: after the all-gather of the code bytes every rank rebuilds the hot bitmap (code != 0) and the hashed
// filter locally, so only ONE array has to cross NVLink per step.  A warp handles 1024 sources per iteration:
// lane L owns sources [32L, 32L+32) of the tile = one bitmap word.
struct PeerPtrs { const uint8_t* p[16]; };

// PULL = true: fused halo exchange — the slices of the other ranks are read straight from their (symmetric-memory)
// code buffers over NVLink (P2P loads), copied into the local code array and turned into bitmap/filter bits in the
// same pass.  peers.p[r] = base of rank r's buffer (the rank's own base for its own slice).
template <bool PULL>
__global__ void __launch_bounds__(256)
hot_from_code_kernel(uint8_t* __restrict__ code, int64_t n_sources, uint32_t* __restrict__ hot,
                     uint32_t* __restrict__ filter, int hash_words, PeerPtrs peers, int rank, int64_t per)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_words = n_sources >> 5;
    uint32_t seen = 0;
    for (int64_t w0 = warp_id * 32; w0 < n_words; w0 += n_warps * 32) {
        const int64_t w = w0 + lane;
        uint32_t word = 0;
        if (w < n_words) {
            const uint8_t* base = code;
            bool remote = false;
            if (PULL) {
                const int owner = (int)((w0 * 32) / per);            // a 1024-source tile never straddles two slices
                remote = owner != rank;
                base = peers.p[owner];
            }
            const uint4* p = reinterpret_cast<const uint4*>(base + w * 32);
            uint4 a, b;
            if (PULL && remote) {
                a = __ldcv(p); b = __ldcv(p + 1);               // peer memory: never served from a stale L1 line
                uint4* q = reinterpret_cast<uint4*>(code + w * 32);
                q[0] = a; q[1] = b;
            } else {
                a = __ldg(p); b = __ldg(p + 1);
            }
            const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int k = 0; k < 4; ++k) word |= (((v[j] >> (8 * k)) & 255u) != 0u ? 1u : 0u) << (4 * j + k);
            }
            hot[w] = word;
            if (word != 0u && filter != nullptr) atomicOr(&filter[w & (hash_words - 1)], word);
        }
        seen |= word;
    }
    if (filter != nullptr && __syncthreads_or(seen != 0u) && threadIdx.x == 0) atomicOr(&filter[hash_words], 1u);
}

// Sparse, barrier-free halo exchange.  Every rank keeps ONE symmetric (peer-mapped) buffer per parity:
//   [ code bytes : n ][ hot bitmap : n/8 (+64) ][ arrival flags : 64 words ]
// Its step kernel writes the code bytes and the hot-bitmap words of its OWN slice.  This kernel then
//   1. tells every peer "my slice of this buffer is written" (a release store of the buffer's epoch into the peer's
//      flag word for this rank) and waits until every peer has said the same — the per-step barrier, inside the
//      kernel that needs it;
//   2. reads the peers' bitmap slices (1 bit per agent) and pulls a 32-byte code sector only where a bitmap word is
//      non-zero (about a quarter of the sectors at 1 % infectious, none in a step without infectious agents),
//      zeroing local sectors whose word just became zero;
//   3. folds the whole bitmap (own slice included) into the hashed filter.
// Double-buffering by step parity makes one rendezvous per step sufficient: a rank overwrites a buffer two steps
// later, after a rendezvous on the other buffer that every peer only joins once it has finished reading this one.
#define AT_XCHG_FLAG_WORDS 64
#define AT_XCHG_BITMAP_PAD 64
#ifndef AT_HOST_EMULATION
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#else
inline uint32_t ld_acquire_sys(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release_sys(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#endif

struct PeerBufs { uint8_t* p[16]; };

__global__ void __launch_bounds__(256)
exchange_sparse_kernel(uint8_t* __restrict__ local, int64_t n_sources, uint32_t* __restrict__ filter, int hash_words,
                       PeerBufs peers, int world, int rank, int64_t per, uint32_t* sync_state, uint32_t* any_word,
                       int32_t* error_flag, int32_t* over_state, int over_after, const uint32_t* __restrict__ halo_mask)
{
    // sync_state: [0] epoch of this buffer, [1] CTA ticket, [2 + r] "my copy of rank r's slice holds non-zero data"
    __shared__ uint32_t s_epoch;
    __shared__ int s_any[16], s_skip[16], s_ei[16];
    const int tid = threadIdx.x, lane = tid & 31;
    // the epidemic has been over for as many exchanges as there are code buffers: every copy is clear, nothing can change
    // (all ranks saw the same signals, so all of them skip the same calls)
    if (over_state != nullptr && *reinterpret_cast<const volatile int32_t*>(over_state) >= over_after) return;
    if (tid == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(sync_state) + 1u;
    __syncthreads();
    const uint32_t epoch = s_epoch;
    uint32_t* hot_local = reinterpret_cast<uint32_t*>(local + n_sources);
    const size_t flag_off = (size_t)n_sources + (size_t)(n_sources >> 3) + AT_XCHG_BITMAP_PAD;
    const int64_t words_per = per >> 5;
    if (blockIdx.x == 0) {
        // does my own slice hold any infectious agent?  (travels with the arrival signal, so that peers skip it)
        // (the step kernel leaves the answer in its workspace; without it the slice's bitmap is scanned)
        uint32_t acc = 0, ei = 1u;
        if (any_word != nullptr) {
            if (tid == 0) { acc = atomicExch(any_word, 0u); ei = atomicExch(any_word + (WS_HASEI - WS_ANYHOT) / 4, 0u); }
        } else {
            const uint4* mine4 = reinterpret_cast<const uint4*>(hot_local + rank * words_per);
            for (int64_t i = tid; i < (words_per >> 2); i += blockDim.x) { const uint4 v = mine4[i]; acc |= v.x | v.y | v.z | v.w; }
        }
        const int any_own = __syncthreads_or(acc != 0u);
        const int ei_own = __syncthreads_or(tid == 0 && ei != 0u);
        // arrival signal: epoch of the buffer, "somebody of mine is in E or I", "somebody of mine is infectious in this step"
        const uint32_t sig = (epoch << 2) | ((uint32_t)(ei_own != 0) << 1) | (uint32_t)(any_own != 0);
        if (tid < world && tid != rank) {
            __threadfence_system();
            st_release_sys(reinterpret_cast<uint32_t*>(peers.p[tid] + flag_off) + rank, sig);
        }
        if (tid == 0) *(reinterpret_cast<volatile uint32_t*>(local + flag_off) + rank) = sig;
    }
    if (tid < world) {
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(local + flag_off) + tid;
        const long long t0 = clock64();
        uint32_t v = ld_acquire_sys(mine);
        while ((int32_t)(((v >> 2) - epoch) << 2) < 0) {
            if (clock64() - t0 > 40000000000ll) {            // ~20 s: a peer never arrived; report instead of hanging
                if (error_flag) *error_flag = 2;
                break;
            }
            __nanosleep(100);
            v = ld_acquire_sys(mine);
        }
        // (a peer cannot be more than one use of this buffer ahead, so the low bit belongs to this epoch)
        s_any[tid] = (int)(v & 1u);
        s_ei[tid] = (int)((v >> 1) & 1u);
        const int live = (int)reinterpret_cast<volatile uint32_t*>(sync_state)[2 + tid];
        s_skip[tid] = !(v & 1u) && (tid == rank || !live);          // nothing hot there now and nothing to clear here
    }
    __syncthreads();

    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + tid) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_words = n_sources >> 5;
    uint32_t seen = 0;
    for (int64_t w0 = warp_id * 32; w0 < n_words; w0 += n_warps * 32) {
        const int owner = (int)((w0 * 32) / per);             // a 1024-source tile never straddles two slices
        if (s_skip[owner]) continue;
        // halo mask: the 32-source groups of a peer's slice that some edge of this rank's rows refers to (one bit per group, one
        // mask word per warp round); the others are neither read nor written
        uint32_t wanted = 0xFFFFFFFFu;
        if (halo_mask != nullptr && owner != rank) {
            wanted = halo_mask[w0 >> 5];
            if (wanted == 0u) continue;
        }
        const int64_t w = w0 + lane;
        uint32_t word = 0;
        if (w < n_words && ((wanted >> lane) & 1u)) {
            if (owner == rank) {
                word = hot_local[w];
            } else {
                const uint8_t* peer = peers.p[owner];
                const uint32_t old = hot_local[w];
                if (s_any[owner]) word = __ldcv(reinterpret_cast<const uint32_t*>(peer + n_sources) + w);   // never from a stale L1 line
                uint4* dst = reinterpret_cast<uint4*>(local + w * 32);
                if (word != 0u) {
                    const uint4* src = reinterpret_cast<const uint4*>(peer + w * 32);
                    const uint4 a = __ldcv(src), b = __ldcv(src + 1);
                    dst[0] = a; dst[1] = b;
                } else if (old != 0u) {
                    dst[0] = make_uint4(0, 0, 0, 0); dst[1] = make_uint4(0, 0, 0, 0);
                }
                if (word != old) hot_local[w] = word;
            }
            if (word != 0u && filter != nullptr) atomicOr(&filter[w & (hash_words - 1)], word);
        }
        seen |= word;
    }
    if (filter != nullptr && __syncthreads_or(seen != 0u) && tid == 0) atomicOr(&filter[hash_words], 1u);
    // the last CTA to finish publishes the epoch and the slice states (every CTA has read the old values by then)
    __shared__ int s_last;
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(reinterpret_cast<int*>(sync_state) + 1, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        if (tid < world) reinterpret_cast<volatile uint32_t*>(sync_state)[2 + tid] = (uint32_t)s_any[tid];
        if (tid == 0 && over_state != nullptr) {
            int any_ei = 0;
            for (int r = 0; r < world; ++r) any_ei |= s_ei[r];
            *over_state = any_ei ? 0 : *over_state + 1;
        }
        if (tid == 0) {
            reinterpret_cast<int*>(sync_state)[1] = 0;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(sync_state) = epoch;
        }
    }
}

// -------------------------------------------------------------------------------- backward kernel
// grad_r[t] = gI[t]*sum_dne[t] + sum_i dne[t,i] * c_i,  c_i = gD[recovered_step_i] * M / 2
#define AT_BWD_THREADS 256
// Agents with c_i != 0 (they recovered inside the episode) are sparse: a warp scans 32 agents at a time and, for
// every such agent, lane t fetches dne[t, agent] — lane <-> step, so no cross-lane reduction is needed until the end.
// An agent is an exposure candidate only up to the step of its exposure (= its infected_time), so lanes t > infected_time
// fetch nothing and agents infected before the episode (the seeds: most of those that recover inside it) are skipped.
// Static assignment and fixed-order folds: deterministic.
__global__ void __launch_bounds__(AT_BWD_THREADS)
seirm_episode_bwd_kernel(int T, int64_t n, int64_t ld, const float* __restrict__ dne, const uint8_t* __restrict__ rstep,
                         const float* __restrict__ sum_dne, const float* __restrict__ death_count,
                         const float* __restrict__ gI, const float* __restrict__ gD, const float* __restrict__ m_value,
                         float* grad_r, float* grad_m, char* ws, const int32_t* __restrict__ skipped,
                         const void* __restrict__ itime, int itime_f32)
{
    __shared__ float s_gd[256];
    __shared__ double s_acc[AT_BWD_THREADS / 32][BWD_MAXT];
    __shared__ int s_is_last;
    double* partials = reinterpret_cast<double*>(ws + WS_BPART);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float half_m = m_value[0] * 0.5f;
    if (tid < 256) s_gd[tid] = (tid < T) ? gD[tid] * half_m : 0.f;
    __syncthreads();
    double acc0 = 0.0, acc1 = 0.0;                  // steps lane and lane + 32
    // steps the forward skipped (epidemic over: their dne rows were not written and mean zero)
    const bool skip0 = skipped != nullptr && lane < T && skipped[lane] != 0;
    const bool skip1 = skipped != nullptr && lane + 32 < T && skipped[lane + 32] != 0;
    // scan: a warp takes 512 agents per round, lane <-> 16 consecutive agents (one 128-bit load of recovered_step); `hit` = the
    // lane's agents with c != 0 that can have been exposure candidates inside the episode
    const int64_t ngroups = (n + 511) >> 9;
    const int64_t wstride = (int64_t)gridDim.x * (AT_BWD_THREADS / 32);
    for (int64_t g = (int64_t)blockIdx.x * (AT_BWD_THREADS / 32) + warp; g < ngroups; g += wstride) {
        const int64_t base = g * 512 + (int64_t)lane * 16;
        uint32_t rs[4] = {0u, 0u, 0u, 0u};
        if (base + 16 <= n && (reinterpret_cast<uintptr_t>(rstep) & 15) == 0) {
            const uint4 v = *reinterpret_cast<const uint4*>(rstep + base);
            rs[0] = v.x; rs[1] = v.y; rs[2] = v.z; rs[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k)                                              // 255 = never recovered
                rs[k >> 2] |= (base + k < n ? (uint32_t)rstep[base + k] : 255u) << ((k & 3) * 8);
        }
        // byte k of the lane's 16 (selects, so that the words stay in registers)
        auto rbyte = [&](int k) -> uint32_t {
            const uint32_t w = k < 8 ? (k < 4 ? rs[0] : rs[1]) : (k < 12 ? rs[2] : rs[3]);
            return (w >> ((k & 3) * 8)) & 255u;
        };
        uint32_t hit = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (s_gd[(rs[k >> 2] >> ((k & 3) * 8)) & 255u] != 0.f) hit |= 1u << k;
        if (itime != nullptr) {
            // agents infected before the episode (the seeds) were never candidates in it
            for (uint32_t h = hit; h != 0u; h &= h - 1u) {
                const int k = __ffs((int)h) - 1;
                if (load_timer(itime, base + k, itime_f32 != 0) < 0) hit &= ~(1u << k);
            }
        }
        while (__any_sync(0xffffffffu, hit != 0u)) {
            // four agents per round: their (independent) loads are in flight together, the adds keep agent order
            float ci[4], v0[4], v1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ci[j] = 0.f; v0[j] = 0.f; v1[j] = 0.f;
                const uint32_t any = __ballot_sync(0xffffffffu, hit != 0u);
                if (any) {
                    const int src = __ffs((int)any) - 1;
                    int k = 0, last = BWD_MAXT;
                    float c = 0.f;
                    if (lane == src) {
                        k = __ffs((int)hit) - 1;
                        hit &= hit - 1u;
                        c = s_gd[rbyte(k)];
                        if (itime != nullptr) last = load_timer(itime, base + k, itime_f32 != 0);
                    }
                    k = __shfl_sync(0xffffffffu, k, src);
                    ci[j] = __shfl_sync(0xffffffffu, c, src);
                    const int upto = __shfl_sync(0xffffffffu, last, src);
                    const int64_t agent = g * 512 + (int64_t)src * 16 + k;
                    if (lane < T && !skip0 && lane <= upto) v0[j] = dne[(int64_t)lane * ld + agent];
                    if (lane + 32 < T && !skip1 && lane + 32 <= upto) v1[j] = dne[(int64_t)(lane + 32) * ld + agent];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc0 += (double)(v0[j] * ci[j]);
                acc1 += (double)(v1[j] * ci[j]);
            }
        }
    }
    s_acc[warp][lane] = acc0;
    s_acc[warp][lane + 32] = acc1;
    __syncthreads();
    int* ticket = reinterpret_cast<int*>(ws + WS_BTICKET);
    if (tid < T) {
        double tot = 0.0;
        for (int w = 0; w < AT_BWD_THREADS / 32; ++w) tot += s_acc[w][tid];
        partials[(size_t)blockIdx.x * BWD_MAXT + tid] = tot;
    }
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
static thread_local int g_sm_count = 0, g_smem_optin = 0, g_props_dev = -1;     // properties of the CURRENT device
static int device_props() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    if (g_sm_count && dev == g_props_dev) return 0;
    g_props_dev = dev;
    if (cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
    if (cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 1;
    return 0;
}

static int code_count(int n_lg) { return ((1 + 3 * n_lg) + 7) & ~7; }
static size_t table_smem(int n_sfs, int n_lg) { return (size_t)2 * n_sfs * code_count(n_lg) * 8 + 2 * 256 * 4; }

template <int MODE, int RL4>
static int launch_kernel(const KModel& km, const at_seirm_launch& a, int grid, int threads, size_t smem, int bm_words, int NC, int G, cudaStream_t s) {
    auto k = seirm_kernel<MODE, RL4>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    AT_LAUNCH_SMEM(k, grid, threads, smem, s, km, a, bm_words, NC, G);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail("seirm_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

template <int MODE>
static int dispatch(const KModel& km, const at_seirm_launch& a, int grid, int threads, size_t smem, int words, int NC, int G, cudaStream_t s) {
    switch (a.edges_per_chunk) {
        case 512: return launch_kernel<MODE, 4>(km, a, grid, threads, smem, words, NC, G, s);
        case 768: return launch_kernel<MODE, 6>(km, a, grid, threads, smem, words, NC, G, s);
        case 1024: return launch_kernel<MODE, 8>(km, a, grid, threads, smem, words, NC, G, s);
    }
    return fail("%s", "edges_per_chunk must be 512, 768 or 1024");
}

template <int NTHREADS, int UF>
static int launch_stream(const KModel& km, const at_seirm_launch& a, int grid, int threads, size_t smem, int bm_words, int NC, int tile_shift, cudaStream_t s) {
    auto k = stream_gather_kernel<NTHREADS, UF>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    AT_LAUNCH_SMEM(k, grid, threads, smem, s, km, a, bm_words, NC, tile_shift);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail("stream_gather_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" {

int at_seirm_abi_version(void) { return AT_SEIRM_ABI_VERSION; }
int64_t at_seirm_sizeof_launch(void) { return (int64_t)sizeof(at_seirm_launch); }
int64_t at_seirm_sizeof_model(void) { return (int64_t)sizeof(at_seirm_model); }
const char* at_last_error(void) { return g_err; }
int64_t at_seirm_workspace_bytes(void) { return (int64_t)WS_BYTES; }

int64_t at_seirm_max_block_size(int32_t n_sfs, int32_t n_lam_gamma, int32_t edges_per_chunk, int32_t warps) {
    if (device_props()) return 0;
    if (warps < 1) warps = NT_MAX / 32;
    int64_t bytes = (int64_t)g_smem_optin - (int64_t)table_smem(n_sfs, n_lam_gamma) - (int64_t)warps * warp_smem_bytes(edges_per_chunk) - 2048;
    if (bytes < 4096) return 0;
    return ((bytes * 8) / 1024) * 1024;
}

int64_t at_seirm_stream_max_block_size(int32_t n_sfs, int32_t n_lam_gamma, int32_t warps) {
    if (device_props()) return 0;
    if (warps < 1) warps = ST_NT / 32;
    int64_t bytes = (int64_t)g_smem_optin - (int64_t)warps * ST_PEND * 8 - 16 - 16 - 16 - (int64_t)n_sfs * code_count(n_lam_gamma) * 8 - 1024;
    if (bytes < 4096) return 0;
    int64_t bits = ((bytes * 8) / 1024) * 1024;
    return bits > AT_STREAM_MAX_BLOCK_SIZE ? AT_STREAM_MAX_BLOCK_SIZE : bits;
}

int at_seirm_step(const at_seirm_model* model, const at_seirm_launch* launch, void* stream) {
    if (!model || !launch) return fail("%s", "null argument");
    const at_seirm_launch& a = *launch;
    if (a.abi_version != AT_SEIRM_ABI_VERSION) return fail("%s", "at_seirm_launch.abi_version mismatch");
    if (model->n_lam_gamma < 1 || model->n_lam_gamma > AT_MAX_LAMGAMMA) return fail("%s", "n_lam_gamma out of range (1..84)");
    if (model->n_sfs < 1 || model->n_sfs > AT_MAX_SFS || model->n_sfi < 1 || model->n_sfi > AT_MAX_SFI) return fail("%s", "table size out of range");
    if (a.n_rows < 0 || (a.row_offset & 31) != 0 || (a.n_rows_pad & 255) != 0 || a.n_rows_pad < a.n_rows) return fail("%s", "bad n_rows / n_rows_pad / row_offset");
    if (!a.stage || !a.workspace || !a.sus_bitmap || !a.exp_bitmap || !a.inf_bitmap) return fail("%s", "stage / workspace / row bitmaps must not be null");
    const bool do_trans = a.flags & AT_DO_TRANSMISSION;
    if (a.layout != AT_LAYOUT_BCSR && a.layout != AT_LAYOUT_STREAM) return fail("%s", "unknown layout");
    const bool is_stream = a.layout == AT_LAYOUT_STREAM;
    const int epc = a.edges_per_chunk > 0 ? a.edges_per_chunk : 768;
    if (epc != 512 && epc != 768 && epc != 1024) return fail("%s", "edges_per_chunk must be 512, 768 or 1024");
    if (do_trans) {
        if ((!is_stream && (!a.row_start || !a.seg_info)) || !a.col || !a.code_in || !a.hot_bitmap_in || !a.r_value || !a.age) return fail("%s", "transmission needs the graph, code_in, hot_bitmap_in, age, r_value");
        if (a.n_blocks < 1 || a.n_blocks > AT_MAX_BLOCKS || a.block_size < 1024 || (a.block_size & 1023)) return fail("%s", "bad n_blocks / block_size");
        if ((int64_t)a.n_blocks * a.block_size < a.n_sources) return fail("%s", "n_blocks * block_size does not cover n_sources");
        if (!is_stream && a.rows_per_task != (a.n_blocks > 1 ? 256 : 32)) return fail("%s", "rows_per_task must be 32 (one block) or 256 (several blocks)");
        if (is_stream) {
            if (a.tile_rows < 32 || a.tile_rows > AT_STREAM_MAX_TILE_ROWS || (a.tile_rows & (a.tile_rows - 1)) || (a.n_rows_pad % a.tile_rows))
                return fail("%s", "stream layout: tile_rows must be a power of two in 32..2048 dividing n_rows_pad");
            if (a.block_size > AT_STREAM_MAX_BLOCK_SIZE) return fail("%s", "stream layout: block_size exceeds 2^21 - 1024 sources");
            if (a.edge_weight) return fail("%s", "stream layout: edge weights are not supported");
            if (a.hash_words) return fail("%s", "stream layout: hash_words must be 0");
            if (!a.unit_tile || !a.block_units || !a.rec || !a.tile_rec_base || !a.tile_count) return fail("%s", "stream layout needs unit_tile, block_units, rec, tile_rec_base and tile_count");
            if (a.lam_deg_bits < 0 || a.lam_deg_bits > 40) return fail("%s", "stream layout: lam_deg_bits out of range (0..40)");
        }
        if (a.iso_mode == AT_ISO_TENSOR && !a.isolation) return fail("%s", "iso_mode TENSOR without isolation pointer");
        if (a.u_mode == AT_U_TENSOR && !a.uniform) return fail("%s", "u_mode TENSOR without uniform pointer");
        if (!is_stream && a.n_blocks > 1 && (!a.lam_part || !a.dirty)) return fail("%s", "n_blocks > 1 needs lam_part / dirty");
        if (a.flags & AT_DO_INIT) return fail("%s", "AT_DO_INIT cannot be combined with AT_DO_TRANSMISSION");
    }
    if ((a.flags & AT_DO_PROGRESSION) && (!a.next_stage_time || (a.daily_deaths && !a.m_value))) return fail("%s", "progression needs next_stage_time and m_value");
    if ((a.flags & AT_DO_ENCODE) && (!a.code_out || !a.hot_bitmap_out || !a.infected_time)) return fail("%s", "encode needs code_out, hot_bitmap_out, infected_time");
    if (a.t < 0 || a.t > 255) return fail("%s", "t out of range (0..255)");
    if (a.xchg_slot < 0 || a.xchg_slot > 2) return fail("%s", "xchg_slot out of range (0..2)");
    if ((a.flags & AT_DO_LOOKAHEAD) && (!(a.flags & AT_DO_ENCODE) || !a.next_stage_time)) return fail("%s", "AT_DO_LOOKAHEAD needs AT_DO_ENCODE and next_stage_time");
    if (device_props()) return fail("%s", "cannot query device");
    if (a.n_rows == 0) return 0;

    KModel km;
    km.n_sfs = model->n_sfs; km.n_sfi = model->n_sfi; km.n_lg = model->n_lam_gamma;
    km.e2i = model->exposed_to_infected_time; km.i2r = model->infected_to_recovered_time; km.inf_time = model->infinity_time;
    km.ibar = model->i_bar;
    km.n_layers = model->n_layers > 1 ? model->n_layers : 1;
    if (km.n_layers > AT_MAX_LAYERS) return fail("%s", "n_layers out of range (<= 8)");
    for (int l = 0; l < AT_MAX_LAYERS; ++l) km.ibar_layers[l] = (model->n_layers > 1 && l < model->n_layers) ? model->i_bar_layers[l] : model->i_bar;
    if (km.n_layers > 1 && a.layout != AT_LAYOUT_STREAM && (a.flags & AT_DO_TRANSMISSION)) return fail("%s", "multi-layer contact networks need the stream layout");
    if ((int64_t)km.n_layers * model->n_sfs * code_count(model->n_lam_gamma) > (int64_t)AT_MAX_SFS * 256) return fail("%s", "layers x susceptibility classes x codes exceed the term table (4096 entries)");
    memcpy(km.sfs, model->sfs, sizeof(km.sfs)); memcpy(km.sfi, model->sfi, sizeof(km.sfi)); memcpy(km.lg, model->lam_gamma, sizeof(km.lg));

    const int NC = code_count(model->n_lam_gamma);
    const size_t tables = table_smem(model->n_sfs, model->n_lam_gamma);
    at_seirm_launch al = a;
    al.edges_per_chunk = epc;
    int warps = a.warps_per_cta > 0 ? a.warps_per_cta : NT_MAX / 32;
    if (warps > NT_MAX / 32) warps = NT_MAX / 32;
    int bm_words = 0;
    const bool hashed = do_trans && a.hash_words > 0;
    if (do_trans && !is_stream) {
        const int64_t bits = hashed ? (int64_t)a.hash_words * 32 : (int64_t)a.block_size;
        if (hashed && (a.n_blocks != 1 || (a.hash_words & (a.hash_words - 1)) != 0 || a.hash_words < 32)) return fail("%s", "hash_words must be a power of two >= 32 and n_blocks 1");
        while (warps > 4 && bits > at_seirm_max_block_size(model->n_sfs, model->n_lam_gamma, epc, warps)) --warps;
        if (bits > at_seirm_max_block_size(model->n_sfs, model->n_lam_gamma, epc, warps)) return fail("%s", "bitmap exceeds the shared-memory capacity (at_seirm_max_block_size)");
        bm_words = (int)(bits / 32);
    }
    const size_t smem_gather = (size_t)warps * warp_smem_bytes(epc) + tables + (size_t)bm_words * 4;
    const size_t smem_update = tables;
    const int64_t ntasks32 = a.n_rows_pad / 32;
    int grid = a.grid > 0 ? a.grid : g_sm_count;
    if (grid > 1024) grid = 1024;
    cudaStream_t s = (cudaStream_t)stream;
    int ugrid = a.grid > 0 ? a.grid : g_sm_count * 3;          // update kernel: 8-warp CTAs, 3 per SM
    {
        const int64_t nsuper = (ntasks32 + 7) / 8;
        const int64_t uuseful = (nsuper + 7) / 8;
        if (ugrid > uuseful) ugrid = (int)uuseful;
        if (ugrid < 1) ugrid = 1;
    }
    if (do_trans && is_stream) {
        int pw = a.warps_per_cta > 0 ? a.warps_per_cta : ST_NT / 32;
        if (pw > ST_NT / 32) pw = ST_NT / 32;
        if (a.block_size > at_seirm_stream_max_block_size(model->n_sfs, model->n_lam_gamma, pw))
            return fail("%s", "stream layout: block bitmap exceeds the shared-memory capacity (at_seirm_stream_max_block_size)");
        const int words = a.block_size / 32;
        const size_t smem = (size_t)(words + 4) * 4 + 16 + (size_t)pw * ST_PEND * 8 + 16 + (size_t)km.n_layers * model->n_sfs * NC * 8;
        int tile_shift = 0;
        while ((1 << tile_shift) < a.tile_rows) ++tile_shift;
        // the more warps, the fewer registers each has for units in flight: 32 warps x 3 units, 24 x 5, <= 16 x 8 (KB per SM: 96 / 120 / 128)
        int rc;
        if (a.reserve_sms > 0) {
            // leave room for a concurrent halo exchange: the lean gather (56 registers) here, one update CTA per SM fewer below
            auto k = stream_gather_lean_kernel;
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fail("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            AT_LAUNCH_SMEM(k, grid, (pw > 32 ? 32 : pw) * 32, smem, s, km, al, words, NC, tile_shift);
            e = cudaGetLastError();
            rc = (e != cudaSuccess) ? fail("stream_gather_lean_kernel launch: %s", cudaGetErrorString(e)) : 0;
        } else if (pw > 24 && a.tasks_per_group == 2) rc = launch_stream<1024, 2>(km, al, grid, pw * 32, smem, words, NC, tile_shift, s);
        else if (pw > 24) rc = launch_stream<1024, 3>(km, al, grid, pw * 32, smem, words, NC, tile_shift, s);
        else if (pw > 16) rc = launch_stream<768, 4>(km, al, grid, pw * 32, smem, words, NC, tile_shift, s);
        else rc = launch_stream<512, 6>(km, al, grid, pw * 32, smem, words, NC, tile_shift, s);
        if (rc) return rc;
        al.edges_per_chunk = 512;
        int tgrid = a.grid > 0 ? a.grid : g_sm_count * (a.reserve_sms > 0 ? TU_MIN_CTAS - 1 : TU_MIN_CTAS);
        const int64_t n_tiles = a.n_rows_pad / a.tile_rows;
        if (tgrid > n_tiles) tgrid = (int)n_tiles;
        const size_t usmem = (size_t)km.n_layers * model->n_sfs * NC * 8 + 2 * 256 * 4 + (size_t)a.tile_rows * (a.dne_e_dr ? 19 : 11);
        const bool tf32 = a.timer_dtype == AT_TIMER_F32, grad = a.dne_dr != nullptr;
        if (tf32 && grad) AT_LAUNCH_SMEM((stream_update_kernel<true, true>), tgrid, TU_THREADS, usmem, s, km, al, NC);
        else if (tf32) AT_LAUNCH_SMEM((stream_update_kernel<true, false>), tgrid, TU_THREADS, usmem, s, km, al, NC);
        else if (grad) AT_LAUNCH_SMEM((stream_update_kernel<false, true>), tgrid, TU_THREADS, usmem, s, km, al, NC);
        else AT_LAUNCH_SMEM((stream_update_kernel<false, false>), tgrid, TU_THREADS, usmem, s, km, al, NC);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail("stream_update_kernel launch: %s", cudaGetErrorString(e));
        return 0;
    }
    if (do_trans && a.n_blocks > 1) {
        const int Gg = a.tasks_per_group > 0 ? (a.tasks_per_group > MAXG ? MAXG : a.tasks_per_group) : 1;
        int ggrid = grid < a.n_blocks ? a.n_blocks : grid;
        int rc = dispatch<MODE_GATHER>(km, al, ggrid, warps * 32, smem_gather, bm_words, NC, Gg, s);
        if (rc) return rc;
        al.edges_per_chunk = 512;
        return launch_kernel<MODE_UPDATE, 4>(km, al, ugrid, 256, smem_update, 0, NC, 8, s);
    }
    int G = a.tasks_per_group > 0 ? a.tasks_per_group : (do_trans ? (ntasks32 >= 131072 ? 8 : 4) : 8);
    if (G > MAXG) G = MAXG;
    const int64_t ngroups = (ntasks32 + G - 1) / G;
    int rc;
    if (do_trans) {
        int64_t useful = (ngroups + warps - 1) / warps;
        if (grid > useful) grid = (int)useful;
        if (grid < 1) grid = 1;
        if (hashed) rc = dispatch<MODE_HASHED>(km, al, grid, warps * 32, smem_gather, bm_words, NC, G, s);
        else rc = dispatch<MODE_FUSED>(km, al, grid, warps * 32, smem_gather, bm_words, NC, G, s);
    } else if (a.flags == (AT_DO_ENCODE | AT_DO_INIT)) {
        // the first launch of an episode: a dedicated streaming pass
        const int64_t nquad = (ntasks32 + 3) / 4;
        int igrid = a.grid > 0 ? a.grid : g_sm_count * 8;
        if (igrid > (nquad + 7) / 8) igrid = (int)((nquad + 7) / 8);
        if (igrid < 1) igrid = 1;
        if (a.timer_dtype == AT_TIMER_F32) AT_LAUNCH(init_encode_kernel<true>, igrid, 256, s, km, al);
        else AT_LAUNCH(init_encode_kernel<false>, igrid, 256, s, km, al);
        cudaError_t e = cudaGetLastError();
        rc = (e != cudaSuccess) ? fail("init_encode_kernel launch: %s", cudaGetErrorString(e)) : 0;
    } else {
        al.edges_per_chunk = 512;
        rc = launch_kernel<MODE_UPDATE, 4>(km, al, ugrid, 256, smem_update, 0, NC, 8, s);
    }
    if (rc) return rc;
    if ((a.flags & AT_DO_ENCODE) && a.hot_filter_out != nullptr && a.hash_words > 0) {
        cudaError_t e = cudaMemsetAsync(a.hot_filter_out, 0, ((size_t)a.hash_words + 1) * 4, s);
        if (e != cudaSuccess) return fail("cudaMemsetAsync: %s", cudaGetErrorString(e));
        const int M4 = a.hash_words >> 2, total4 = (int)(a.n_sources >> 7);
        int fgrid = (M4 + 255) / 256;
        if (fgrid > g_sm_count) fgrid = g_sm_count;
        int slices = (total4 + M4 - 1) / M4;
        if (slices > 16) slices = 16;
        if (slices < 1) slices = 1;
        AT_LAUNCH(hot_fold_kernel, dim3(fgrid, slices), 256, s, a.hot_bitmap_out, a.hot_filter_out, a.hash_words, total4);
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail("hot_fold_kernel launch: %s", cudaGetErrorString(e));
    }
    return 0;
}

int at_seirm_episode_backward(int32_t num_steps, int64_t n_rows, int64_t ld, const float* dne_dr, const uint8_t* recovered_step,
                              const float* sum_dne_dr, const float* death_count, const float* g_daily_infected,
                              const float* g_daily_deaths, const float* m_value, float* grad_r, float* grad_m,
                              void* workspace, int32_t grid, const int32_t* step_skipped,
                              const void* infected_time, int32_t timer_dtype, void* stream) {
    if (num_steps < 1 || num_steps > BWD_MAXT) return fail("%s", "num_steps out of range (1..64)");
    if (!dne_dr || !recovered_step || !sum_dne_dr || !death_count || !g_daily_infected || !g_daily_deaths || !m_value || !grad_r || !grad_m || !workspace)
        return fail("%s", "null argument");
    if (device_props()) return fail("%s", "cannot query device");
    int g = grid > 0 ? grid : g_sm_count * 6;
    const int64_t tiles = (n_rows + AT_BWD_THREADS * 16 - 1) / (AT_BWD_THREADS * 16);       // a CTA round covers 8 warps x 512 agents
    if (g > tiles) g = (int)tiles;
    if (g > BWD_MAX_GRID) g = BWD_MAX_GRID;
    if (g < 1) g = 1;
    AT_LAUNCH(seirm_episode_bwd_kernel, g, AT_BWD_THREADS, (cudaStream_t)stream, num_steps, n_rows, ld, dne_dr, recovered_step, sum_dne_dr,
        death_count, g_daily_infected, g_daily_deaths, m_value, grad_r, grad_m, (char*)workspace, step_skipped,
        infected_time, timer_dtype == AT_TIMER_F32 ? 1 : 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("seirm_episode_bwd_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

static int hot_from_code_impl(uint8_t* code, int64_t n_sources, uint32_t* hot_bitmap, uint32_t* hot_filter, int32_t hash_words,
                              const void* const* peer_code, int32_t world, int32_t rank, void* stream) {
    if (!code || !hot_bitmap || n_sources <= 0 || (n_sources & 1023)) return fail("%s", "bad code / hot_bitmap / n_sources (multiple of 1024)");
    if (hot_filter && (hash_words < 32 || (hash_words & (hash_words - 1)))) return fail("%s", "hash_words must be a power of two >= 32");
    if (device_props()) return fail("%s", "cannot query device");
    cudaStream_t s = (cudaStream_t)stream;
    if (hot_filter) {
        cudaError_t e = cudaMemsetAsync(hot_filter, 0, ((size_t)hash_words + 1) * 4, s);
        if (e != cudaSuccess) return fail("cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
    int64_t tiles = (n_sources / 32 + 255) / 256;
    int grid = (int)(tiles < g_sm_count * 4 ? tiles : g_sm_count * 4);
    if (grid < 1) grid = 1;
    PeerPtrs pp;
    memset(&pp, 0, sizeof(pp));
    if (peer_code) {
        if (world < 1 || world > 16 || rank < 0 || rank >= world || (n_sources % world) || ((n_sources / world) & 1023))
            return fail("%s", "pull: world (<= 16) equal slices of whole 1024-source tiles are required");
        for (int r = 0; r < world; ++r) pp.p[r] = (const uint8_t*)peer_code[r];
        AT_LAUNCH(hot_from_code_kernel<true>, grid, 256, s, code, n_sources, hot_bitmap, hot_filter, hash_words, pp, rank, n_sources / world);
    } else {
        AT_LAUNCH(hot_from_code_kernel<false>, grid, 256, s, code, n_sources, hot_bitmap, hot_filter, hash_words, pp, 0, n_sources);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("hot_from_code_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

int at_seirm_hot_from_code(const uint8_t* code, int64_t n_sources, uint32_t* hot_bitmap, uint32_t* hot_filter,
                           int32_t hash_words, void* stream) {
    return hot_from_code_impl(const_cast<uint8_t*>(code), n_sources, hot_bitmap, hot_filter, hash_words, nullptr, 1, 0, stream);
}

int64_t at_seirm_exchange_bytes(int64_t n_sources) {
    return n_sources + (n_sources >> 3) + AT_XCHG_BITMAP_PAD + AT_XCHG_FLAG_WORDS * 4;
}

int at_seirm_exchange_sparse(uint8_t* local_buf, int64_t n_sources, uint32_t* hot_filter, int32_t hash_words,
                             void* const* peer_buf, int32_t world, int32_t rank, uint32_t* sync_state, void* step_workspace,
                             int32_t* error_flag, int32_t max_ctas, int32_t* over_state, int32_t over_after,
                             const uint32_t* halo_mask, void* stream) {
    if (!local_buf || !peer_buf || !sync_state) return fail("%s", "exchange_sparse: null buffer");
    if (n_sources <= 0 || (n_sources & 1023)) return fail("%s", "exchange_sparse: n_sources must be a positive multiple of 1024");
    if (world < 2 || world > 16 || rank < 0 || rank >= world || (n_sources % world) || ((n_sources / world) & 1023))
        return fail("%s", "exchange_sparse: 2..16 ranks owning equal slices of whole 1024-source tiles are required");
    if (hot_filter && (hash_words < 32 || (hash_words & (hash_words - 1)))) return fail("%s", "hash_words must be a power of two >= 32");
    if (device_props()) return fail("%s", "cannot query device");
    cudaStream_t s = (cudaStream_t)stream;
    if (hot_filter) {
        cudaError_t e = cudaMemsetAsync(hot_filter, 0, ((size_t)hash_words + 1) * 4, s);
        if (e != cudaSuccess) return fail("cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
    PeerBufs pb;
    memset(&pb, 0, sizeof(pb));
    for (int r = 0; r < world; ++r) {
        if (!peer_buf[r]) return fail("%s", "exchange_sparse: null peer pointer");
        pb.p[r] = (uint8_t*)peer_buf[r];
    }
    if (pb.p[rank] != local_buf) return fail("%s", "exchange_sparse: peer_buf[rank] must be the local buffer");
    // every CTA spins on the arrival flags: keep all of them co-resident (one wave)
    int64_t tiles = (n_sources / 32 + 255) / 256;
    int grid = (int)(tiles < g_sm_count * 4 ? tiles : g_sm_count * 4);
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    if (grid < 1) grid = 1;
    uint32_t* any_word = step_workspace ? reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(step_workspace) + WS_ANYHOT) : nullptr;
    AT_LAUNCH(exchange_sparse_kernel, grid, 256, s, local_buf, n_sources, hot_filter, hash_words, pb, world, rank, n_sources / world,
                                                sync_state, any_word, error_flag, over_state, over_after < 1 ? 2 : over_after, halo_mask);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("exchange_sparse_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

int at_seirm_pull_codes(uint8_t* code, int64_t n_sources, uint32_t* hot_bitmap, uint32_t* hot_filter, int32_t hash_words,
                        const void* const* peer_code, int32_t world, int32_t rank, void* stream) {
    if (!peer_code) return fail("%s", "peer_code must not be null");
    return hot_from_code_impl(code, n_sources, hot_bitmap, hot_filter, hash_words, peer_code, world, rank, stream);
}

int at_noise_fill(float* out, int64_t n, int64_t id_offset, uint64_t seed, int32_t step, int32_t what, float p_isolate, void* stream) {
    if (!out && n > 0) return fail("%s", "null output");
    if (n <= 0) return 0;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    AT_LAUNCH(noise_fill_kernel, grid, 256, (cudaStream_t)stream, out, n, id_offset, seed, step, what, p_isolate);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("noise_fill_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
