// at_predprey.cu — sm_100a kernels + C ABI for the predator-prey grid model (see include/at_predprey.h).
//
// The reference walks Python lists: for every agent the row of a dense (cells x cells) adjacency matrix
// (substeps/move.py:24-46), for every eatable / target position a comparison against ALL agents
// (eat.py:80-89, hunt.py:70-80).  Here an agent is (cell, energy) and every "is anybody / is grown grass at this
// cell" question is one bit of a cell-indexed bitmap (2 MB for 16.8 M cells: L2-resident), so a step is two
// streaming passes — agents (16 B per agent) and cells (8 or 16 B per cell) — with random single-bit lookups
// served by L2.  Bits are set with atomicOr (bitwise OR: order-independent), so the results are deterministic.
// HBM-bound; no tensor cores.
//
//   pp_agent_kernel   thread <-> agent: EAT (prey on grown grass gain nutrition, mark `eaten`), HUNT (prey on a
//                     cell with a predator -> energy 0; predator on a cell with a prey gains nutrition), MOVE
//                     (counter-hash draw -> neighbour cell if energy > 0; energy -= work; mark occupancy of the
//                     new cell).  Fusing [EAT, HUNT](t) with MOVE(t+1) is legal: MOVE only needs the agent's own
//                     energy, and writes the OTHER pair of occupancy bitmaps.
//   pp_cell_kernel    lane <-> cell, warp <-> bitmap word: eaten -> countdown reset (float32-rounded increment, as
//                     the reference's float32 mask does, eat.py:72,99), GROW, and clearing of the finished step's
//                     occupancy words.

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "at_noise.cuh"
#include "at_predprey.h"

static thread_local char g_err[512] = "";

static int fail(const char* fmt, const char* detail) {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return 1;
}

#define PP_THREADS 256
#define PP_SPECIES_BIT 39          // species in bit 39 of the counter's agent field (oracle/noise.py: move_draw)

static int default_grid() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            sms = 148;
    }
    return sms * (2048 / PP_THREADS);          // full occupancy, one wave
}

__device__ __forceinline__ bool test_bit(const uint32_t* __restrict__ bm, int cell) {
    return (__ldg(bm + (cell >> 5)) >> (cell & 31)) & 1u;
}

// neighbour number (draw * count) >> 24 of the ascending list [k - width, k - 1, k + 1, k + width] (those that exist)
__device__ __forceinline__ int lattice_step(int cell, uint32_t draw24, int width, int64_t n_cells) {
    const int i = cell % width;
    const bool up = cell - width >= 0, left = i > 0, right = i < width - 1, down = (int64_t)cell + width < n_cells;
    const int cnt = (int)up + (int)left + (int)right + (int)down;
    if (cnt == 0) return cell;
    int k = (int)((draw24 * (uint32_t)cnt) >> 24);
    if (up) { if (k == 0) return cell - width; --k; }
    if (left) { if (k == 0) return cell - 1; --k; }
    if (right) { if (k == 0) return cell + 1; --k; }
    return cell + width;
}

__global__ void __launch_bounds__(PP_THREADS)
pp_occupancy_kernel(at_pp_launch a)
{
    const int64_t n = a.n_pred + a.n_prey;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const bool prey = idx >= a.n_pred;
        const int cell = prey ? a.prey_cell[idx - a.n_pred] : a.pred_cell[idx];
        atomicOr((prey ? a.occ_prey_in : a.occ_pred_in) + (cell >> 5), 1u << (cell & 31));
    }
}

__global__ void __launch_bounds__(PP_THREADS)
pp_agent_kernel(at_pp_model m, at_pp_launch a)
{
    const int flags = a.flags;
    const int64_t n = a.n_pred + a.n_prey;
    const uint64_t seed = a.seed_ptr ? *a.seed_ptr : a.seed;
    const uint32_t t = (uint32_t)(a.t_move + (a.t_ptr ? *a.t_ptr : 0));
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const bool prey = idx >= a.n_pred;
        const int64_t i = prey ? idx - a.n_pred : idx;
        int32_t* cell_p = (prey ? a.prey_cell : a.pred_cell) + i;
        float* energy_p = (prey ? a.prey_energy : a.pred_energy) + i;
        const int cell0 = *cell_p;
        const float e0 = *energy_p;
        int cell = cell0;
        float e = e0;
        if (prey) {
            if ((flags & AT_PP_EAT) && test_bit(a.grown, cell)) {                     // eat.py:41, :93-97
                e = __fadd_rn(e, m.grass_nutrition);
                atomicOr(a.eaten + (cell >> 5), 1u << (cell & 31));
            }
            if ((flags & AT_PP_HUNT) && test_bit(a.occ_pred_in, cell)) e = 0.0f;       // hunt.py:81
        } else {
            if ((flags & AT_PP_HUNT) && test_bit(a.occ_prey_in, cell)) e = __fadd_rn(e, m.prey_nutrition);   // hunt.py:83-86
        }
        if (flags & AT_PP_MOVE) {
            if (e > 0.0f) {                                                         // move.py:89
                const uint64_t z = at_noise_u64(seed, t, ((uint64_t)(prey ? 1 : 0) << PP_SPECIES_BIT) | (uint64_t)i);
                const uint32_t draw = (uint32_t)(z >> 40);
                if (a.nbr_ptr == nullptr) {
                    cell = lattice_step(cell, draw, m.width, m.n_cells);
                } else {
                    const int64_t lo = a.nbr_ptr[cell], cnt = a.nbr_ptr[cell + 1] - lo;
                    if (cnt > 0) cell = a.nbr_idx[lo + (int64_t)(((uint64_t)draw * (uint64_t)cnt) >> 24)];
                }
            }
            if (prey ? m.apply_prey_work : m.apply_pred_work)                       // move.py:112-116
                e = __fadd_rn(e, -(prey ? m.prey_work : m.pred_work));
            uint32_t* occ = prey ? a.occ_prey_out : a.occ_pred_out;
            if (occ != nullptr) atomicOr(occ + (cell >> 5), 1u << (cell & 31));
        }
        if (cell != cell0) *cell_p = cell;
        if (e != e0) *energy_p = e;
    }
}

template <typename T>
__global__ void __launch_bounds__(PP_THREADS)
pp_cell_kernel(at_pp_model m, at_pp_launch a)
{
    const int flags = a.flags;
    const int lane = threadIdx.x & 31;
    const int64_t n_words = (m.n_cells + 31) >> 5;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    T* cd = reinterpret_cast<T*>(a.countdown);
    for (int64_t w = warp0; w < n_words; w += n_warps) {
        const uint32_t ew = (flags & AT_PP_EAT_APPLY) ? a.eaten[w] : 0u;
        if ((flags & AT_PP_GROW) || ew != 0u) {
            const int64_t cell = w * 32 + lane;
            const bool in = cell < m.n_cells;
            bool g = (a.grown[w] >> lane) & 1u;
            const bool eat = (ew >> lane) & 1u;
            if (in && ((flags & AT_PP_GROW) || eat)) {
                T c = cd[cell];
                if (eat) {
                    // countdown_mask[node] = regrowth_time - countdown[node] lives in a float32 tensor (eat.py:72,99)
                    const float reset = (float)((T)m.regrowth_time - c);
                    c = c + (T)reset;
                    g = false;                                                      // stage 1 + (-1), eat.py:98,104
                }
                if (flags & AT_PP_GROW) {                                           // grow.py:34-40
                    c = c - (T)1;
                    g = c <= (T)0;
                }
                cd[cell] = c;
            }
            const uint32_t word = __ballot_sync(0xffffffffu, in && g);
            if (lane == 0) {
                a.grown[w] = word;
                if (ew != 0u) a.eaten[w] = 0u;
            }
        }
        if ((flags & AT_PP_CLEAR) && lane == 0) {
            a.occ_pred_in[w] = 0u;
            a.occ_prey_in[w] = 0u;
        }
    }
}

// ----------------------------------------------------------------------------------- layout conversion
__global__ void __launch_bounds__(PP_THREADS)
pp_pack_cells_kernel(const int64_t* __restrict__ coords, int64_t n, int max_y, int64_t n_cells, int32_t* cells, int32_t* error_flag)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const longlong2 xy = reinterpret_cast<const longlong2*>(coords)[i];
        const int64_t cell = (int64_t)max_y * xy.x + xy.y;                           // move.py:36
        if (cell < 0 || cell >= n_cells) { if (error_flag) *error_flag = 1; cells[i] = 0; }
        else cells[i] = (int32_t)cell;
    }
}

__global__ void __launch_bounds__(PP_THREADS)
pp_unpack_cells_kernel(const int32_t* __restrict__ cells, int64_t n, int max_y, int64_t* coords)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int cell = cells[i];
        const int y = cell % max_y;                                                  // move.py:43-44
        reinterpret_cast<longlong2*>(coords)[i] = make_longlong2((cell - y) / max_y, y);
    }
}

__global__ void __launch_bounds__(PP_THREADS)
pp_pack_stage_kernel(const void* stage, int dtype, int64_t n_cells, uint32_t* grown)
{
    const int lane = threadIdx.x & 31;
    const int64_t n_words = (n_cells + 31) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words; w += n_warps) {
        const int64_t cell = w * 32 + lane;
        bool g = false;
        if (cell < n_cells) g = dtype == 0 ? ((const int64_t*)stage)[cell] == 1 : ((const float*)stage)[cell] == 1.0f;
        const uint32_t word = __ballot_sync(0xffffffffu, g);
        if (lane == 0) grown[w] = word;
    }
}

__global__ void __launch_bounds__(PP_THREADS)
pp_unpack_stage_kernel(const uint32_t* __restrict__ grown, int64_t n_cells, void* stage, int dtype)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t g = (grown[c >> 5] >> (c & 31)) & 1u;
        if (dtype == 0) ((int64_t*)stage)[c] = g; else ((float*)stage)[c] = (float)g;
    }
}

__global__ void __launch_bounds__(PP_THREADS)
pp_census_kernel(at_pp_launch a, int64_t n_cells, unsigned long long* out)
{
    const int64_t n = a.n_pred + a.n_prey;
    const int64_t n_words = (n_cells + 31) >> 5;
    unsigned long long c0 = 0, c1 = 0, c2 = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
        if (idx < a.n_pred) c0 += a.pred_energy[idx] > 0.0f; else c1 += a.prey_energy[idx - a.n_pred] > 0.0f;
    }
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) c2 += __popc(a.grown[w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_down_sync(0xffffffffu, c0, o);
        c1 += __shfl_down_sync(0xffffffffu, c1, o);
        c2 += __shfl_down_sync(0xffffffffu, c2, o);
    }
    if ((threadIdx.x & 31) == 0) {                       // integer sums: order-independent
        if (c0) atomicAdd(out + 0, c0);
        if (c1) atomicAdd(out + 1, c1);
        if (c2) atomicAdd(out + 2, c2);
    }
}

// ------------------------------------------------------------------------------------------- C ABI
static int check(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e)); return 1; }
    return 0;
}

static int blocks_for(int64_t n, int grid) {
    const int64_t need = (n + PP_THREADS - 1) / PP_THREADS;
    const int64_t cap = grid > 0 ? grid : default_grid();
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

extern "C" {

int at_pp_abi_version(void) { return AT_PP_ABI_VERSION; }
const char* at_pp_last_error(void) { return g_err; }
int64_t at_pp_sizeof_model(void) { return (int64_t)sizeof(at_pp_model); }
int64_t at_pp_sizeof_launch(void) { return (int64_t)sizeof(at_pp_launch); }
int64_t at_pp_bitmap_words(int64_t n_cells) { return (((n_cells + 31) >> 5) + 3) & ~(int64_t)3; }

int at_pp_step(const at_pp_model* model, const at_pp_launch* launch, void* stream)
{
    if (model == nullptr || launch == nullptr) return fail("at_pp_step: %s", "null model/launch");
    if (model->abi_version != AT_PP_ABI_VERSION || launch->abi_version != AT_PP_ABI_VERSION)
        return fail("at_pp_step: %s", "ABI version mismatch (rebuild libat_predprey.so or update the binding)");
    const at_pp_model& m = *model;
    const at_pp_launch& a = *launch;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int flags = a.flags;
    if (m.n_cells <= 0 || m.n_cells > 0x7fffffffll) return fail("at_pp_step: %s", "n_cells must be in [1, 2^31)");
    if (m.width <= 0 || m.max_y <= 0) return fail("at_pp_step: %s", "width and max_y must be positive");
    if (a.n_pred < 0 || a.n_prey < 0) return fail("at_pp_step: %s", "negative agent count");
    const int64_t n = a.n_pred + a.n_prey;
    const int agent_flags = flags & (AT_PP_EAT | AT_PP_HUNT | AT_PP_MOVE);
    const int cell_flags = flags & (AT_PP_EAT_APPLY | AT_PP_GROW | AT_PP_CLEAR);
    if ((agent_flags || (flags & AT_PP_OCCUPANCY)) && n > 0 &&
        ((a.n_pred && (!a.pred_cell || !a.pred_energy)) || (a.n_prey && (!a.prey_cell || !a.prey_energy))))
        return fail("at_pp_step: %s", "agent arrays missing");
    if ((flags & (AT_PP_EAT | AT_PP_EAT_APPLY)) && (!a.grown || !a.eaten)) return fail("at_pp_step: %s", "EAT needs grown + eaten bitmaps");
    if ((flags & (AT_PP_HUNT | AT_PP_OCCUPANCY | AT_PP_CLEAR)) && (!a.occ_pred_in || !a.occ_prey_in))
        return fail("at_pp_step: %s", "HUNT / OCCUPANCY / CLEAR need the occ_*_in bitmaps");
    if ((flags & (AT_PP_EAT_APPLY | AT_PP_GROW)) && (!a.grown || !a.countdown)) return fail("at_pp_step: %s", "cell pass needs grown + countdown");
    if ((a.nbr_ptr == nullptr) != (a.nbr_idx == nullptr)) return fail("at_pp_step: %s", "nbr_ptr and nbr_idx go together");

    if (flags & AT_PP_OCCUPANCY) {
        const size_t bytes = (size_t)at_pp_bitmap_words(m.n_cells) * 4;
        cudaError_t e = cudaMemsetAsync(a.occ_pred_in, 0, bytes, s);
        if (e == cudaSuccess) e = cudaMemsetAsync(a.occ_prey_in, 0, bytes, s);
        if (e != cudaSuccess) return fail("cudaMemsetAsync: %s", cudaGetErrorString(e));
        if (n > 0) {
            pp_occupancy_kernel<<<blocks_for(n, a.grid), PP_THREADS, 0, s>>>(a);
            if (check("pp_occupancy_kernel launch")) return 1;
        }
    }
    if (agent_flags && n > 0) {
        pp_agent_kernel<<<blocks_for(n, a.grid), PP_THREADS, 0, s>>>(m, a);
        if (check("pp_agent_kernel launch")) return 1;
    }
    if (cell_flags) {
        const int64_t threads = ((m.n_cells + 31) >> 5) * 32;
        if (a.countdown_f64) pp_cell_kernel<double><<<blocks_for(threads, a.grid), PP_THREADS, 0, s>>>(m, a);
        else pp_cell_kernel<float><<<blocks_for(threads, a.grid), PP_THREADS, 0, s>>>(m, a);
        if (check("pp_cell_kernel launch")) return 1;
    }
    return 0;
}

int at_pp_pack_cells(const int64_t* coords, int64_t n, int32_t max_y, int64_t n_cells, int32_t* cells, int32_t* error_flag, void* stream)
{
    if (n <= 0) return 0;
    if (!coords || !cells || max_y <= 0) return fail("at_pp_pack_cells: %s", "bad arguments");
    pp_pack_cells_kernel<<<blocks_for(n, 0), PP_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(coords, n, max_y, n_cells, cells, error_flag);
    return check("pp_pack_cells_kernel launch");
}

int at_pp_unpack_cells(const int32_t* cells, int64_t n, int32_t max_y, int64_t* coords, void* stream)
{
    if (n <= 0) return 0;
    if (!coords || !cells || max_y <= 0) return fail("at_pp_unpack_cells: %s", "bad arguments");
    pp_unpack_cells_kernel<<<blocks_for(n, 0), PP_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(cells, n, max_y, coords);
    return check("pp_unpack_cells_kernel launch");
}

int at_pp_pack_stage(const void* stage, int32_t stage_dtype, int64_t n_cells, uint32_t* grown, void* stream)
{
    if (n_cells <= 0 || !stage || !grown || (stage_dtype != 0 && stage_dtype != 1)) return fail("at_pp_pack_stage: %s", "bad arguments");
    pp_pack_stage_kernel<<<blocks_for(((n_cells + 31) >> 5) * 32, 0), PP_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(stage, stage_dtype, n_cells, grown);
    return check("pp_pack_stage_kernel launch");
}

int at_pp_unpack_stage(const uint32_t* grown, int64_t n_cells, void* stage, int32_t stage_dtype, void* stream)
{
    if (n_cells <= 0 || !stage || !grown || (stage_dtype != 0 && stage_dtype != 1)) return fail("at_pp_unpack_stage: %s", "bad arguments");
    pp_unpack_stage_kernel<<<blocks_for(n_cells, 0), PP_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(grown, n_cells, stage, stage_dtype);
    return check("pp_unpack_stage_kernel launch");
}

int at_pp_census(const at_pp_launch* launch, int64_t n_cells, int64_t* out3, void* stream)
{
    if (!launch || !out3 || !launch->grown) return fail("at_pp_census: %s", "bad arguments");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const cudaError_t e = cudaMemsetAsync(out3, 0, 24, s);
    if (e != cudaSuccess) return fail("cudaMemsetAsync: %s", cudaGetErrorString(e));
    const int64_t n = launch->n_pred + launch->n_prey;
    const int64_t work = n > ((n_cells + 31) >> 5) ? n : ((n_cells + 31) >> 5);
    pp_census_kernel<<<blocks_for(work, 0), PP_THREADS, 0, s>>>(*launch, n_cells, reinterpret_cast<unsigned long long*>(out3));
    return check("pp_census_kernel launch");
}

}  // extern "C"
