// at_interventions.cu — quarantine, testing and SEIRMS-progression substeps of the covid model for sm_100a
// (C ABI: include/at_interventions.h; semantics: oracle/interventions_port.py, pinned against the reference).
//
// All three are streaming per-agent passes over float32 SoA state — HBM-bound, no tensor cores.  A thread owns
// 4 consecutive agents: one 16-byte load per property, the draws regenerated in registers from the counter hash,
// one 16-byte store per property and only where a value changed (when the output aliases the input).  Grids are
// sized to the SM count (persistent grid-stride loop).
#ifndef AT_HOST_EMULATION           // tests/host_emulation/ compiles this file for the CPU: a shim stands in for the runtime
#include <cuda_runtime.h>
#define AT_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "at_interventions.h"
#include "at_noise.cuh"

namespace {

constexpr int IV_THREADS = 256;
constexpr int STREAM_QUARANTINE = 1, STREAM_TESTING = 2, STREAM_MORTALITY = 3;

thread_local char g_err[512] = "";

int fail(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

int check(const char* what)
{
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("%s: %s", what, cudaGetErrorString(e));
    return 0;
}

int default_grid()
{
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        sms = n;
    }
    return sms * 8;          // 8 CTAs of 256 threads per SM: full occupancy for a register-light streaming kernel
}

int blocks_for(int64_t groups)
{
    const int64_t need = (groups + IV_THREADS - 1) / IV_THREADS;
    const int64_t cap = default_grid();
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- 4-agent group access: vector when the caller's arrays are 16-byte aligned and the group is full
struct F4 { float v[4]; };

__device__ __forceinline__ F4 load4(const float* __restrict__ p, int64_t i, int64_t n, bool vec, float fill = 0.f)
{
    F4 r;
    if (vec && i + 3 < n) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(p + i));
        r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) r.v[k] = (i + k < n) ? p[i + k] : fill;
    }
    return r;
}

__device__ __forceinline__ void store4(float* __restrict__ p, int64_t i, int64_t n, bool vec, const F4& r)
{
    if (vec && i + 3 < n) {
        __stcs(reinterpret_cast<float4*>(p + i), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i + k < n) p[i + k] = r.v[k];
    }
}

__device__ __forceinline__ void store4_u8(uint8_t* __restrict__ p, int64_t i, int64_t n, const bool (&b)[4])
{
    if (i + 3 < n && (reinterpret_cast<uintptr_t>(p + i) & 3u) == 0) {
        const uint32_t w = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
        *reinterpret_cast<uint32_t*>(p + i) = w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i + k < n) p[i + k] = (uint8_t)b[k];
    }
}

__device__ __forceinline__ void load4_u8(const uint8_t* __restrict__ p, int64_t i, int64_t n, bool (&b)[4])
{
    if (i + 3 < n && (reinterpret_cast<uintptr_t>(p + i) & 3u) == 0) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(p + i);
#pragma unroll
        for (int k = 0; k < 4; ++k) b[k] = ((w >> (8 * k)) & 0xffu) != 0;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) b[k] = (i + k < n) && p[i + k] != 0;
    }
}

__device__ __forceinline__ bool differs(const F4& a, const F4& b)
{
    return a.v[0] != b.v[0] || a.v[1] != b.v[1] || a.v[2] != b.v[2] || a.v[3] != b.v[3];
}

__device__ __forceinline__ float u_hi(uint64_t z) { return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f); }
__device__ __forceinline__ float u_lo(uint64_t z) { return (float)(uint32_t)((z >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f); }
__device__ __forceinline__ uint64_t agent_key(int64_t id, int stream) { return (uint64_t)id | ((uint64_t)stream << 36); }

// ------------------------------------------------------------------------------------------- quarantine
__global__ void __launch_bounds__(IV_THREADS, 5) iv_quarantine_kernel(const at_iv_quarantine_args a, const bool vec)
{
    const uint64_t seed = a.seed_ptr ? *a.seed_ptr : a.seed;
    const bool decide = a.flags & AT_IV_DECIDE, apply = a.flags & AT_IV_APPLY;
    const int64_t groups = (a.n + 3) >> 2;
    const float tf = (float)a.t;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = g << 2;
        // every load of the group is issued before anything is computed: one round of memory latency per group
        const F4 q = load4(a.is_q_in, i, a.n, vec);
        F4 d, sv, us, ub;
        if (apply) d = load4(a.start_date_in, i, a.n, vec);
        if (decide && !a.stage_f64) sv = load4(reinterpret_cast<const float*>(a.stage), i, a.n, vec);
        if (decide && a.u_start) us = load4(a.u_start, i, a.n, vec, 1.f);
        if (decide && a.u_break) ub = load4(a.u_break, i, a.n, vec, 1.f);
        bool start[4], brk[4];
        if (!decide) {
            load4_u8(a.start_action, i, a.n, start);
            load4_u8(a.break_action, i, a.n, brk);
        }
        if (decide) {
            // quarantine/action.py:27-46,59-66: start = draw & not quarantined (I2) & (E or I); break = quarantined & draw
            float s[4];
            if (a.stage_f64) {
#pragma unroll
                for (int k = 0; k < 4; ++k) s[k] = (i + k < a.n) ? (float)reinterpret_cast<const double*>(a.stage)[i + k] : 0.f;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) s[k] = sv.v[k];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool ei = (s[k] == 1.f) | (s[k] == 2.f);
                start[k] = false;
                brk[k] = false;
                if (ei | (q.v[k] != 0.f)) {          // everybody else decides nothing: no hash for the susceptible majority
                    float u0 = 1.f, u1 = 1.f;
                    if (!a.u_start || !a.u_break) {
                        const uint64_t z = at_noise_u64(seed, (uint32_t)a.t, agent_key(a.id_offset + i + k, STREAM_QUARANTINE));
                        u0 = u_hi(z);
                        u1 = u_lo(z);
                    }
                    if (a.u_start) u0 = us.v[k];
                    if (a.u_break) u1 = ub.v[k];
                    start[k] = ei & (truncf(q.v[k]) == 0.f) & (u0 < a.p_start);
                    brk[k] = (q.v[k] != 0.f) & (u1 < a.p_break);
                }
            }
            if (a.start_action) store4_u8(a.start_action, i, a.n, start);
            if (a.break_action) store4_u8(a.break_action, i, a.n, brk);
        }
        if (apply) {
            // quarantine/transition.py:31-90: end (I1: not conditioned on the status), start, break — in this order
            F4 nq, nd;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool ends = tf >= d.v[k] + a.quarantine_days;
                float x = q.v[k] + (ends ? -1.f : 0.f);
                x = x + (start[k] ? 1.f : 0.f);
                x = x + (brk[k] ? -1.f : 0.f);
                float y = ends ? a.infinity_time : d.v[k];
                y = start[k] ? tf : y;
                y = brk[k] ? a.infinity_time : y;
                nq.v[k] = x;
                nd.v[k] = y;
            }
            if (a.is_q_out != a.is_q_in || differs(nq, q)) store4(a.is_q_out, i, a.n, vec, nq);
            if (a.start_date_out != a.start_date_in || differs(nd, d)) store4(a.start_date_out, i, a.n, vec, nd);
        }
    }
}

// ---------------------------------------------------------------------------------------------- testing
__global__ void __launch_bounds__(IV_THREADS, 4) iv_testing_kernel(const at_iv_testing_args a, const bool vec)
{
    const uint64_t seed = a.seed_ptr ? *a.seed_ptr : a.seed;
    const bool decide = a.flags & AT_IV_DECIDE, apply = a.flags & AT_IV_APPLY;
    const int64_t groups = (a.n + 3) >> 2;
    const float tf = (float)a.t;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = g << 2;
        F4 w, q, uc, aw, rd, re;
        if (decide) q = load4(a.is_q, i, a.n, vec);
        if (decide && a.u_comply) uc = load4(a.u_comply, i, a.n, vec, 1.f);
        if (!decide) w = load4(a.accept_action, i, a.n, vec);
        if (apply) {
            aw = load4(a.awaiting_in, i, a.n, vec);
            rd = load4(a.result_date_in, i, a.n, vec);
            re = load4(a.re_eligible_in, i, a.n, vec);
        }
        if (decide) {
            // testing/action.py:29-53: every stage is "eligible" (I3); willing = not quarantined (I2) & compliance draw
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float u = a.u_comply ? uc.v[k]
                                           : u_hi(at_noise_u64(seed, (uint32_t)a.t, agent_key(a.id_offset + i + k, STREAM_TESTING)));
                w.v[k] = ((truncf(q.v[k]) == 0.f) & (u < a.p_comply)) ? 1.f : 0.f;
            }
            if (a.accept_action) store4(a.accept_action, i, a.n, vec, w);
        }
        if (apply) {
            // testing/transition.py:53-58,96-100 (results due today), :146-163 (enrolment)
            F4 naw, nrd, nre;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool today = rd.v[k] == tf;
                float x = aw.v[k] + (today ? -1.f : 0.f);
                const float e = today ? tf + a.ineligible_days : re.v[k];
                float r = today ? a.infinity_time : rd.v[k];
                const bool ineligible = (tf < e) | (truncf(x) != 0.f);
                const bool enrolled = (truncf(w.v[k]) != 0.f) & !ineligible;
                x = x + (enrolled ? 1.f : 0.f);
                r = enrolled ? tf + a.result_delay_days : r;
                naw.v[k] = x;
                nrd.v[k] = r;
                nre.v[k] = e;
            }
            if (a.awaiting_out != a.awaiting_in || differs(naw, aw)) store4(a.awaiting_out, i, a.n, vec, naw);
            if (a.result_date_out != a.result_date_in || differs(nrd, rd)) store4(a.result_date_out, i, a.n, vec, nrd);
            if (a.re_eligible_out != a.re_eligible_in || differs(nre, re)) store4(a.re_eligible_out, i, a.n, vec, nre);
        }
    }
}

// ----------------------------------------------------------------------------------- SEIRMS progression
// The death draw is needed by the few agents whose infection ends today: kept out of line so that its 64-bit arithmetic does
// not take registers away from the loads of the streaming loop.
__device__ __noinline__ float seirms_death_draw(uint64_t seed, uint32_t t, int64_t id)
{
    return u_hi(at_noise_u64(seed, t, agent_key(id, STREAM_MORTALITY)));
}

__global__ void __launch_bounds__(IV_THREADS, 4) iv_seirms_kernel(const at_iv_seirms_args a, const bool vec)
{
    const uint64_t seed = a.seed_ptr ? *a.seed_ptr : a.seed;
    const int64_t groups = (a.n + 3) >> 2;
    const float tf = (float)a.t;
    const float m = *a.m_value;
    int n_due = 0, n_died = 0, n_survived = 0;
    // Two groups (8 agents) per thread and iteration, all their loads issued before the arithmetic (64 bytes in flight per
    // thread).  Measured: 81 us per launch at 33.5 M agents against 76 us with one group — the pass is bound by instruction
    // issue, not by latency (profiles/r01_interventions_kernels.md).
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g0 < groups; g0 += 2 * stride) {
        F4 s2[2], nt2[2], ud2[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t i = (g0 + h * stride) << 2;            // past the end: load4 fills, nothing is stored
            s2[h] = load4(a.stage_in, i, a.n, vec);
            nt2[h] = load4(a.next_stage_time_in, i, a.n, vec, 3.0e38f);
            if (a.u_death) ud2[h] = load4(a.u_death, i, a.n, vec, 1.f);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t g = g0 + h * stride;
            if (g < groups) {
                const int64_t i = g << 2;
                const F4 s = s2[h], nt = nt2[h], ud = ud2[h];
                F4 ns, nn;
                uint8_t codes[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool due = nt.v[k] <= tf;
                    const float st = s.v[k];
                    float x = st, y = nt.v[k];
                    uint8_t c = 0;
                    if (due && st == 1.f) {                    // E -> I: stage + (stage/E)*1            (transition.py:87-91,103)
                        x = 2.f;
                        c = 1;
                    } else if (due && st == 3.f) {             // R -> S: stage + (stage/R)*(-R)         (:92-96,106)
                        x = 0.f;
                        c = 2;
                    } else if (due && st == 2.f) {             // I: Bernoulli(M) death draw             (:66-70)
                        const float u = a.u_death ? ud.v[k] : seirms_death_draw(seed, (uint32_t)a.t, a.id_offset + i + k);
                        const bool died = u < m;
                        ++n_due;
                        if (died) {                            // I5: the dead take the one-level path -> stage 3
                            x = 3.f;
                            c = 4;
                            ++n_died;
                        } else {                               //     the survivors the two-level path -> stage 4
                            x = 4.f;
                            c = 3;
                            ++n_survived;
                        }
                    }
                    if (st == 1.f && nt.v[k] == tf) y = tf + a.infected_to_recovered_time;      // :131-133; others: see the epilogue (I6)
                    ns.v[k] = x;
                    nn.v[k] = y;
                    codes[k] = c;
                }
                if (a.stage_out != a.stage_in || differs(ns, s)) store4(a.stage_out, i, a.n, vec, ns);
                if (a.next_stage_time_out != a.next_stage_time_in || differs(nn, nt)) store4(a.next_stage_time_out, i, a.n, vec, nn);
                if (a.move_code) {
                    if (i + 3 < a.n && (reinterpret_cast<uintptr_t>(a.move_code + i) & 3u) == 0) {
                        *reinterpret_cast<uint32_t*>(a.move_code + i) =
                            (uint32_t)codes[0] | ((uint32_t)codes[1] << 8) | ((uint32_t)codes[2] << 16) | ((uint32_t)codes[3] << 24);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (i + k < a.n) a.move_code[i + k] = codes[k];
                    }
                }
            }
        }
    }
    // integer counts: order-independent, so the result is bitwise reproducible for any grid
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_due += __shfl_xor_sync(0xffffffffu, n_due, o);
        n_died += __shfl_xor_sync(0xffffffffu, n_died, o);
        n_survived += __shfl_xor_sync(0xffffffffu, n_survived, o);
    }
    if ((threadIdx.x & 31) == 0 && n_due) {
        atomicAdd(reinterpret_cast<unsigned long long*>(a.counts + 0), (unsigned long long)n_due);
        if (n_died) atomicAdd(reinterpret_cast<unsigned long long*>(a.counts + 1), (unsigned long long)n_died);
        if (n_survived) atomicAdd(reinterpret_cast<unsigned long long*>(a.counts + 2), (unsigned long long)n_survived);
    }
}

__global__ void iv_seirms_epilogue_kernel(const at_iv_seirms_args a)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int64_t due = a.counts[0], died = a.counts[1], survived = a.counts[2];
    // I6 (transition.py:135-141): `new[newly_dead.long()] = INFINITY` then `new[newly_recovered.long()] = R2S`, where (I5)
    // newly_dead holds the survivors' 0/1 flags and newly_recovered the deads': the VALUES index rows 0 and 1.
    float* nst = a.next_stage_time_out;
    if (survived < a.n) nst[0] = a.infinity_time;
    if (survived > 0) nst[1] = a.infinity_time;
    if (died < a.n) nst[0] = a.recovered_to_susceptible_time;
    if (died > 0) nst[1] = a.recovered_to_susceptible_time;
    // transition.py:56-64: daily_deaths[t] += float(sum(stage * due_infected / 2)) * M
    if (a.daily_deaths && a.t < a.num_steps) a.daily_deaths[a.t] = a.daily_deaths[a.t] + (float)due * (*a.m_value);
}

}  // namespace

extern "C" {

int at_iv_abi_version(void) { return AT_IV_ABI_VERSION; }
const char* at_iv_last_error(void) { return g_err; }
int64_t at_iv_sizeof_quarantine_args(void) { return (int64_t)sizeof(at_iv_quarantine_args); }
int64_t at_iv_sizeof_testing_args(void) { return (int64_t)sizeof(at_iv_testing_args); }
int64_t at_iv_sizeof_seirms_args(void) { return (int64_t)sizeof(at_iv_seirms_args); }

int at_iv_quarantine(const at_iv_quarantine_args* args, void* stream)
{
    if (!args) return fail("at_iv_quarantine: %s", "null args");
    const at_iv_quarantine_args& a = *args;
    if (a.abi_version != AT_IV_ABI_VERSION) return fail("at_iv_quarantine: %s", "ABI version mismatch (rebuild libat_interventions.so or update the binding)");
    if (a.n < 0 || a.n >= (1ll << 36) || a.id_offset < 0 || a.id_offset + a.n > (1ll << 36)) return fail("at_iv_quarantine: %s", "agent ids must stay below 2^36");
    if (!(a.flags & (AT_IV_DECIDE | AT_IV_APPLY))) return fail("at_iv_quarantine: %s", "flags select nothing");
    if (a.n == 0) return 0;
    if (!a.is_q_in) return fail("at_iv_quarantine: %s", "is_q_in missing");
    if ((a.flags & AT_IV_DECIDE) && !a.stage) return fail("at_iv_quarantine: %s", "DECIDE needs the disease stage");
    if ((a.flags & AT_IV_DECIDE) && !(a.flags & AT_IV_APPLY) && !a.start_action && !a.break_action)
        return fail("at_iv_quarantine: %s", "DECIDE without APPLY needs an action tensor to write");
    if (!(a.flags & AT_IV_DECIDE) && (!a.start_action || !a.break_action)) return fail("at_iv_quarantine: %s", "APPLY without DECIDE reads both action tensors");
    if ((a.flags & AT_IV_APPLY) && (!a.start_date_in || !a.is_q_out || !a.start_date_out)) return fail("at_iv_quarantine: %s", "APPLY needs start_date_in, is_q_out, start_date_out");
    const bool vec = aligned16(a.is_q_in) && aligned16(a.start_date_in) && aligned16(a.is_q_out) && aligned16(a.start_date_out) &&
                     aligned16(a.u_start) && aligned16(a.u_break) && (a.stage_f64 || aligned16(a.stage));
    AT_LAUNCH(iv_quarantine_kernel, blocks_for((a.n + 3) >> 2), IV_THREADS, reinterpret_cast<cudaStream_t>(stream), a, vec);
    return check("iv_quarantine_kernel launch");
}

int at_iv_testing(const at_iv_testing_args* args, void* stream)
{
    if (!args) return fail("at_iv_testing: %s", "null args");
    const at_iv_testing_args& a = *args;
    if (a.abi_version != AT_IV_ABI_VERSION) return fail("at_iv_testing: %s", "ABI version mismatch (rebuild libat_interventions.so or update the binding)");
    if (a.n < 0 || a.n >= (1ll << 36) || a.id_offset < 0 || a.id_offset + a.n > (1ll << 36)) return fail("at_iv_testing: %s", "agent ids must stay below 2^36");
    if (!(a.flags & (AT_IV_DECIDE | AT_IV_APPLY))) return fail("at_iv_testing: %s", "flags select nothing");
    if (a.n == 0) return 0;
    if ((a.flags & AT_IV_DECIDE) && !a.is_q) return fail("at_iv_testing: %s", "DECIDE needs is_q");
    if ((a.flags & AT_IV_DECIDE) && !(a.flags & AT_IV_APPLY) && !a.accept_action) return fail("at_iv_testing: %s", "DECIDE without APPLY needs accept_action");
    if (!(a.flags & AT_IV_DECIDE) && !a.accept_action) return fail("at_iv_testing: %s", "APPLY without DECIDE reads accept_action");
    if ((a.flags & AT_IV_APPLY) && (!a.awaiting_in || !a.result_date_in || !a.re_eligible_in || !a.awaiting_out || !a.result_date_out || !a.re_eligible_out))
        return fail("at_iv_testing: %s", "APPLY needs the three state arrays, in and out");
    const bool vec = aligned16(a.is_q) && aligned16(a.u_comply) && aligned16(a.accept_action) && aligned16(a.awaiting_in) &&
                     aligned16(a.result_date_in) && aligned16(a.re_eligible_in) && aligned16(a.awaiting_out) &&
                     aligned16(a.result_date_out) && aligned16(a.re_eligible_out);
    AT_LAUNCH(iv_testing_kernel, blocks_for((a.n + 3) >> 2), IV_THREADS, reinterpret_cast<cudaStream_t>(stream), a, vec);
    return check("iv_testing_kernel launch");
}

int at_iv_seirms_progression(const at_iv_seirms_args* args, void* stream)
{
    if (!args) return fail("at_iv_seirms_progression: %s", "null args");
    const at_iv_seirms_args& a = *args;
    if (a.abi_version != AT_IV_ABI_VERSION) return fail("at_iv_seirms_progression: %s", "ABI version mismatch (rebuild libat_interventions.so or update the binding)");
    if (a.n < 2 || a.n >= (1ll << 36) || a.id_offset < 0 || a.id_offset + a.n > (1ll << 36))
        return fail("at_iv_seirms_progression: %s", "need 2 <= n and agent ids below 2^36 (the reference indexes rows 0 and 1 of next_stage_time)");
    if (!a.m_value || !a.stage_in || !a.next_stage_time_in || !a.stage_out || !a.next_stage_time_out || !a.counts)
        return fail("at_iv_seirms_progression: %s", "m_value, stage/next_stage_time (in and out) and counts are required");
    if (a.t < 0 || (a.daily_deaths && a.t >= a.num_steps)) return fail("at_iv_seirms_progression: %s", "t outside daily_deaths");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const bool vec = aligned16(a.stage_in) && aligned16(a.next_stage_time_in) && aligned16(a.stage_out) &&
                     aligned16(a.next_stage_time_out) && aligned16(a.u_death);
    AT_LAUNCH(iv_seirms_kernel, blocks_for((a.n + 3) >> 2), IV_THREADS, s, a, vec);
    if (check("iv_seirms_kernel launch")) return 1;
    AT_LAUNCH(iv_seirms_epilogue_kernel, 1, 32, s, a);
    return check("iv_seirms_epilogue_kernel launch");
}

}  // extern "C"
