// at_predprey.cu — sm_100a kernels + C ABI for the predator-prey grid model (see include/at_predprey.h).
//
// The reference walks Python lists: for every agent the row of a dense (cells x cells) adjacency matrix
// (substeps/move.py:24-46), for every eatable / target position a comparison against ALL agents
// (eat.py:80-89, hunt.py:70-80).  Here an agent is (cell, energy) and every "is anybody / is grown grass at this
// cell" question is a bit of the cell's INFO byte (16.8 MB for 4096 x 4096 cells: L2-resident), so a step is two
// streaming passes — agents (16 B per agent) and cells (1 + 4 or 8 B per cell, read and written) — with ONE random
// byte lookup and at most two random bit-sets per agent, served by L2.  Bits are set with atomicOr (bitwise OR:
// order-independent), so the results are deterministic.  No tensor cores.
//
//   pp_agent_kernel   thread <-> agent: EAT (prey on grown grass gain nutrition, mark `eaten`), HUNT (prey on a
//                     cell with a predator -> energy 0; predator on a cell with a prey gains nutrition), MOVE
//                     (counter-hash draw -> neighbour cell if energy > 0; energy -= work; mark occupancy of the
//                     new cell).  Fusing [EAT, HUNT](t) with MOVE(t+1) is legal: MOVE only needs the agent's own
//                     energy, and writes the OTHER pair of occupancy bitmaps.
//   pp_cell_kernel    thread <-> 4 cells (one info word + one countdown vector): eaten -> countdown reset
//                     (float32-rounded increment, as the reference's float32 mask does, eat.py:72,99), GROW, and
//                     clearing of the finished step's occupancy bits.

#ifndef AT_HOST_EMULATION           // tests/host_emulation/ compiles this file for the CPU: a shim stands in for the runtime
#include <cuda_runtime.h>
#define AT_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif
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

#define PP_GROWN 1u
#define PP_EATEN 2u
#define PP_OCC_PRED(parity) (4u << (2 * (parity)))
#define PP_OCC_PREY(parity) (8u << (2 * (parity)))

// The info byte of a cell, read from L2 (other threads of the same launch OR bits into neighbouring positions of the
// byte with atomics; the bits tested here are not written by this launch, so any snapshot of the byte is good).
__device__ __forceinline__ uint32_t cell_info(const uint8_t* info, int cell) {
    return (uint32_t)__ldcg(info + cell);          // benign race with mark_cell (see above); .cg / .nc / default loads time the same
}
// (measured: aggregating the marks of a warp per word with __match_any_sync costs 2.4x more than the atomics it saves)
__device__ __forceinline__ void mark_cell(uint8_t* info, int cell, uint32_t bits) {
    atomicOr(reinterpret_cast<uint32_t*>(info) + (cell >> 2), bits << (8 * (cell & 3)));     // OR: order-independent; benign race with cell_info
}

// cell / width without a hardware division: q = (cell * mul) >> (31 + log2ceil(width)), exact for 0 <= cell < 2^31
// (round-up method of Granlund & Montgomery; mul = floor(2^(31+l) / width) + 1 fits in 32 bits)
struct FastDiv { uint32_t mul; uint32_t shift; };
static FastDiv make_fastdiv(uint32_t d) {
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    FastDiv f;
    f.mul = (uint32_t)(((1ull << (31 + l)) / d) + 1ull);
    f.shift = 31 + l;
    return f;
}
__device__ __forceinline__ int fast_div(int n, FastDiv f) { return (int)(((uint64_t)(uint32_t)n * f.mul) >> f.shift); }

// neighbour number (draw * count) >> 24 of the ascending list [k - width, k - 1, k + 1, k + width] (those that exist)
__device__ __forceinline__ int lattice_step(int cell, int i /* cell % width */, uint32_t draw24, int width, int n_cells) {
    const bool up = cell >= width, left = i > 0, right = i < width - 1, down = cell < n_cells - width;
    const int cnt = (int)up + (int)left + (int)right + (int)down;
    if (cnt == 0) return cell;
    const int k = (int)((draw24 * (uint32_t)cnt) >> 24);
    // k-th existing neighbour, branch-free: position of each candidate in the ascending list
    const int p_left = (int)up, p_right = p_left + (int)left;
    int next = cell + width;                                                     // the last candidate
    next = (right && k == p_right) ? cell + 1 : next;
    next = (left && k == p_left) ? cell - 1 : next;
    next = (up && k == 0) ? cell - width : next;
    return next;
}

// Where a cell's info byte (and countdown) is stored.  Untiled (tiles_per_row == 0): at the cell number.  Tiled: the
// lattice is cut into tiles of 8 rows x 16 columns = one 128-byte line of info bytes, so that the four neighbours of a
// cell are almost always in the same line and a random walker stays near its line for many steps.
__device__ __forceinline__ int storage_of(int row, int col, int tiles_per_row) {
    return ((((row >> 3) * tiles_per_row + (col >> 4)) << 7) | ((row & 7) << 4) | (col & 15));
}

__global__ void __launch_bounds__(PP_THREADS)
pp_occupancy_kernel(at_pp_model m, at_pp_launch a, FastDiv fd)
{
    const int64_t n = a.n_pred + a.n_prey;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const bool prey = idx >= a.n_pred;
        const int cell = prey ? a.prey_cell[idx - a.n_pred] : a.pred_cell[idx];
        int addr = cell;
        if (m.tiles_per_row) { const int row = fast_div(cell, fd); addr = storage_of(row, cell - row * m.width, m.tiles_per_row); }
        mark_cell(a.info, addr, prey ? PP_OCC_PREY(a.occ_in) : PP_OCC_PRED(a.occ_in));
    }
}

// One species.  A thread owns 4 consecutive agent slots (one 16-byte load per array, four independent cell lookups
// in flight), the last n % 4 slots are handled one by one.  EAT (prey), HUNT, then MOVE.
template <bool PREY>
struct AgentPass {
    const at_pp_model& m;
    const at_pp_launch& a;
    FastDiv fd;
    uint64_t seed;
    uint32_t t;
    int flags;
    bool look, move, eat, hunt, mark_eaten, apply_work, tiled;
    uint32_t other_here, my_mark;
    float work;
    const int32_t* __restrict__ ids;

    __device__ __forceinline__ AgentPass(const at_pp_model& m_, const at_pp_launch& a_, FastDiv fd_, uint64_t seed_, uint32_t t_)
        : m(m_), a(a_), fd(fd_), seed(seed_), t(t_) {
        flags = a.flags;
        look = (flags & (AT_PP_EAT | AT_PP_HUNT)) != 0;
        move = (flags & AT_PP_MOVE) != 0;
        eat = PREY && (flags & AT_PP_EAT);
        hunt = (flags & AT_PP_HUNT) != 0;
        mark_eaten = !(flags & AT_PP_EAT_IMPLICIT);
        other_here = PREY ? PP_OCC_PRED(a.occ_in) : PP_OCC_PREY(a.occ_in);
        my_mark = a.occ_out >= 0 ? (PREY ? PP_OCC_PREY(a.occ_out) : PP_OCC_PRED(a.occ_out)) : 0u;
        apply_work = PREY ? m.apply_prey_work : m.apply_pred_work;
        work = -(PREY ? m.prey_work : m.pred_work);
        ids = PREY ? a.prey_id : a.pred_id;
        tiled = m.tiles_per_row != 0;
    }

    struct Where { int row, col, addr; };
    // row / column / storage address of a cell (one multiply-shift; skipped when neither is needed)
    __device__ __forceinline__ Where locate(int cell) const {
        Where w; w.row = 0; w.col = 0; w.addr = cell;
        if (tiled || (move && a.nbr_ptr == nullptr)) {
            w.row = fast_div(cell, fd);
            w.col = cell - w.row * m.width;
            if (tiled) w.addr = storage_of(w.row, w.col, m.tiles_per_row);
        }
        return w;
    }

    // `here` = info byte of the agent's cell (0 when nothing looks at it)
    __device__ __forceinline__ void one(int id, uint32_t here, Where w, int& cell, float& e) const {
        if (PREY) {
            if (eat && (here & PP_GROWN)) {                                           // eat.py:41, :93-97
                e = __fadd_rn(e, m.grass_nutrition);
                if (mark_eaten && !(here & PP_EATEN)) mark_cell(a.info, w.addr, PP_EATEN);
            }
            if (hunt && (here & other_here)) e = 0.0f;                                // hunt.py:81
        } else {
            if (hunt && (here & other_here)) e = __fadd_rn(e, m.prey_nutrition);      // hunt.py:83-86
        }
        if (move) {
            int addr = w.addr;
            if (e > 0.0f) {                                                         // move.py:89
                // draws follow the agent number, not the slot
                const uint64_t who = ((uint64_t)(PREY ? 1 : 0) << PP_SPECIES_BIT) | (uint64_t)(uint32_t)id;
                const uint32_t draw = (uint32_t)(at_noise_u64(seed, t, who) >> 40);
                if (a.nbr_ptr == nullptr) {
                    const int next = lattice_step(cell, w.col, draw, m.width, (int)m.n_cells);
                    if (tiled) {
                        const int d = next - cell;                                  // +-width: one row; +-1: one column
                        const int row = w.row + (d == m.width ? 1 : (d == -m.width ? -1 : 0));
                        const int col = w.col + ((d == m.width || d == -m.width) ? 0 : d);
                        addr = storage_of(row, col, m.tiles_per_row);
                    } else {
                        addr = next;
                    }
                    cell = next;
                } else {
                    const int64_t lo = a.nbr_ptr[cell], cnt = a.nbr_ptr[cell + 1] - lo;
                    if (cnt > 0) cell = a.nbr_idx[lo + (int64_t)(((uint64_t)draw * (uint64_t)cnt) >> 24)];
                    addr = cell;
                }
            }
            if (apply_work) e = __fadd_rn(e, work);                                 // move.py:112-116
            if (my_mark) mark_cell(a.info, addr, my_mark);
        }
    }

    __device__ __forceinline__ void run(bool vec) const {
        const int n = (int)(PREY ? a.n_prey : a.n_pred);
        int32_t* __restrict__ cells = PREY ? a.prey_cell : a.pred_cell;
        float* __restrict__ energies = PREY ? a.prey_energy : a.pred_energy;
        const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
        const int n4 = vec ? n >> 2 : 0;
        for (int q = tid; q < n4; q += nthr) {
            // the agent arrays stream through once per launch: evict-first, so that the info bytes stay in L2
            const int4 c0 = __ldcs(reinterpret_cast<const int4*>(cells) + q);
            const float4 e0 = __ldcs(reinterpret_cast<const float4*>(energies) + q);
            int4 id = make_int4(4 * q, 4 * q + 1, 4 * q + 2, 4 * q + 3);
            if (ids != nullptr && move) id = __ldcs(reinterpret_cast<const int4*>(ids) + q);
            const Where w0 = locate(c0.x), w1 = locate(c0.y), w2 = locate(c0.z), w3 = locate(c0.w);
            uint32_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;
            if (look) {                                              // four independent lookups in flight
                h0 = cell_info(a.info, w0.addr); h1 = cell_info(a.info, w1.addr);
                h2 = cell_info(a.info, w2.addr); h3 = cell_info(a.info, w3.addr);
            }
            int4 c = c0;
            float4 e = e0;
            one(id.x, h0, w0, c.x, e.x);
            one(id.y, h1, w1, c.y, e.y);
            one(id.z, h2, w2, c.z, e.z);
            one(id.w, h3, w3, c.w, e.w);
            if (c.x != c0.x || c.y != c0.y || c.z != c0.z || c.w != c0.w) __stcs(reinterpret_cast<int4*>(cells) + q, c);
            if (e.x != e0.x || e.y != e0.y || e.z != e0.z || e.w != e0.w) __stcs(reinterpret_cast<float4*>(energies) + q, e);
        }
        for (int i = 4 * n4 + tid; i < n; i += nthr) {
            const int cell0 = __ldcs(cells + i);
            const float e0 = __ldcs(energies + i);
            int cell = cell0;
            float e = e0;
            const Where w = locate(cell);
            one(ids ? __ldcs(ids + i) : i, look ? cell_info(a.info, w.addr) : 0u, w, cell, e);
            if (cell != cell0) __stcs(cells + i, cell);
            if (e != e0) __stcs(energies + i, e);
        }
    }
};

__global__ void __launch_bounds__(PP_THREADS)
pp_agent_kernel(at_pp_model m, at_pp_launch a, FastDiv fd, int vec_pred, int vec_prey)
{
    const uint64_t seed = a.seed_ptr ? *a.seed_ptr : a.seed;
    const uint32_t t = (uint32_t)(a.t_move + (a.t_ptr ? *a.t_ptr : 0));
    AgentPass<false>(m, a, fd, seed, t).run(vec_pred != 0);
    AgentPass<true>(m, a, fd, seed, t).run(vec_prey != 0);
}

// thread <-> 4 consecutive cells: one 32-bit info word + one 16/32-byte countdown vector
template <typename T>
__global__ void __launch_bounds__(PP_THREADS)
pp_cell_kernel(at_pp_model m, at_pp_launch a)
{
    const int flags = a.flags;
    const int64_t n_store = m.n_storage ? m.n_storage : m.n_cells;      // cells as stored (tiled: padded to whole tiles)
    const int64_t n_quads = (n_store + 3) >> 2;
    uint32_t* info4 = reinterpret_cast<uint32_t*>(a.info);
    T* cd = reinterpret_cast<T*>(a.countdown);
    const uint32_t clear_mask = (flags & AT_PP_CLEAR) ? (PP_OCC_PRED(a.occ_in) | PP_OCC_PREY(a.occ_in)) * 0x01010101u : 0u;
    const bool grow = (flags & AT_PP_GROW) != 0, apply = (flags & AT_PP_EAT_APPLY) != 0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t w0 = info4[q];
        uint32_t w = w0 & ~clear_mask;
        // eaten cells, as PP_EATEN bits: marked by EAT, or (EAT_IMPLICIT) grown cells on which a prey stands
        const uint32_t eaten = !apply ? 0u
            : (flags & AT_PP_EAT_IMPLICIT) ? ((w0 & (PP_GROWN * 0x01010101u)) << 1) & ((w0 & (PP_OCC_PREY(a.occ_in) * 0x01010101u)) >> (2 + 2 * a.occ_in))
                                           : (w0 & (PP_EATEN * 0x01010101u));
        if (grow || eaten) {
            const int64_t c0 = q * 4;
            const bool full = c0 + 4 <= n_store;
            T c[4];
            if (full) {
                if constexpr (sizeof(T) == 4) {
                    const float4 v = __ldcs(reinterpret_cast<const float4*>(cd + c0));
                    c[0] = (T)v.x; c[1] = (T)v.y; c[2] = (T)v.z; c[3] = (T)v.w;
                } else {
                    const double2 v0 = __ldcs(reinterpret_cast<const double2*>(cd + c0)), v1 = __ldcs(reinterpret_cast<const double2*>(cd + c0 + 2));
                    c[0] = (T)v0.x; c[1] = (T)v0.y; c[2] = (T)v1.x; c[3] = (T)v1.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) c[j] = c0 + j < n_store ? cd[c0 + j] : (T)1;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t sh = 8 * j;
                if ((eaten >> sh) & PP_EATEN) {
                    // countdown_mask[node] = regrowth_time - countdown[node] lives in a float32 tensor (eat.py:72,99)
                    const float reset = (float)((T)m.regrowth_time - c[j]);
                    c[j] = c[j] + (T)reset;
                    w &= ~((PP_GROWN | PP_EATEN) << sh);                            // stage 1 + (-1), eat.py:98,104
                }
                if (grow) {                                                         // grow.py:34-40
                    c[j] = c[j] - (T)1;
                    w = (w & ~(PP_GROWN << sh)) | ((c[j] <= (T)0 && c0 + j < n_store) ? (PP_GROWN << sh) : 0u);
                }
            }
            if (full) {
                if constexpr (sizeof(T) == 4) {
                    __stcs(reinterpret_cast<float4*>(cd + c0), make_float4((float)c[0], (float)c[1], (float)c[2], (float)c[3]));
                } else {
                    __stcs(reinterpret_cast<double2*>(cd + c0), make_double2((double)c[0], (double)c[1]));
                    __stcs(reinterpret_cast<double2*>(cd + c0 + 2), make_double2((double)c[2], (double)c[3]));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (c0 + j < n_store) cd[c0 + j] = c[j];
            }
        }
        if (w != w0) info4[q] = w;
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
pp_pack_stage_kernel(const void* stage, int dtype, int64_t n_cells, int64_t n_bytes, uint8_t* info)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_bytes; c += (int64_t)gridDim.x * blockDim.x) {
        bool g = false;
        if (c < n_cells) g = dtype == 0 ? ((const int64_t*)stage)[c] == 1 : ((const float*)stage)[c] == 1.0f;
        info[c] = g ? (uint8_t)PP_GROWN : (uint8_t)0;
    }
}

__global__ void __launch_bounds__(PP_THREADS)
pp_unpack_stage_kernel(const uint8_t* __restrict__ info, int64_t n_cells, void* stage, int dtype)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t g = info[c] & PP_GROWN;
        if (dtype == 0) ((int64_t*)stage)[c] = g; else ((float*)stage)[c] = (float)g;
    }
}

__global__ void __launch_bounds__(PP_THREADS)
pp_census_kernel(at_pp_launch a, int64_t n_cells, unsigned long long* out)
{
    const int64_t n = a.n_pred + a.n_prey;
    const int64_t n_words = (n_cells + 3) >> 2;                    // padding bytes are never grown
    unsigned long long c0 = 0, c1 = 0, c2 = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
        if (idx < a.n_pred) c0 += a.pred_energy[idx] > 0.0f; else c1 += a.prey_energy[idx - a.n_pred] > 0.0f;
    }
    const uint32_t* info4 = reinterpret_cast<const uint32_t*>(a.info);
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) c2 += __popc(info4[w] & (PP_GROWN * 0x01010101u));
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
int64_t at_pp_info_bytes(int64_t n_cells) { return (n_cells + 15) & ~(int64_t)15; }
int64_t at_pp_storage_cells(int32_t width, int64_t n_cells, int32_t tiled) {
    if (!tiled || width <= 0) return n_cells;
    const int64_t rows = (n_cells + width - 1) / width;
    return ((rows + 7) / 8) * ((width + 15) / 16) * 128;
}

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
    if (m.tiles_per_row < 0 || (m.tiles_per_row && (a.nbr_ptr || m.n_cells % m.width || m.tiles_per_row != (m.width + 15) / 16 ||
                                                    m.n_storage != at_pp_storage_cells(m.width, m.n_cells, 1))))
        return fail("at_pp_step: %s", "tiled storage needs the implicit lattice, n_cells = width * rows, tiles_per_row = ceil(width / 16) and n_storage = at_pp_storage_cells()");
    if (!m.tiles_per_row && m.n_storage && m.n_storage != m.n_cells) return fail("at_pp_step: %s", "untiled storage: n_storage must be 0 or n_cells");
    if (a.n_pred < 0 || a.n_prey < 0 || a.n_pred > 0x7fffffffll || a.n_prey > 0x7fffffffll)
        return fail("at_pp_step: %s", "agent counts must be in [0, 2^31)");
    const int64_t n = a.n_pred + a.n_prey;
    const int agent_flags = flags & (AT_PP_EAT | AT_PP_HUNT | AT_PP_MOVE);
    const int cell_flags = flags & (AT_PP_EAT_APPLY | AT_PP_GROW | AT_PP_CLEAR);
    if ((agent_flags || (flags & AT_PP_OCCUPANCY)) && n > 0 &&
        ((a.n_pred && (!a.pred_cell || !a.pred_energy)) || (a.n_prey && (!a.prey_cell || !a.prey_energy))))
        return fail("at_pp_step: %s", "agent arrays missing");
    if ((flags & (AT_PP_OCCUPANCY | AT_PP_EAT | AT_PP_HUNT | AT_PP_EAT_APPLY | AT_PP_GROW | AT_PP_CLEAR)) && !a.info)
        return fail("at_pp_step: %s", "info bytes missing");
    if ((flags & AT_PP_MOVE) && a.occ_out >= 0 && !a.info) return fail("at_pp_step: %s", "MOVE with occ_out >= 0 needs the info bytes");
    if ((flags & (AT_PP_HUNT | AT_PP_OCCUPANCY | AT_PP_CLEAR)) && (a.occ_in < 0 || a.occ_in > 1)) return fail("at_pp_step: %s", "occ_in must be 0 or 1");
    if (a.occ_out > 1) return fail("at_pp_step: %s", "occ_out must be < 0, 0 or 1");
    if ((flags & (AT_PP_EAT_APPLY | AT_PP_GROW)) && !a.countdown) return fail("at_pp_step: %s", "cell pass needs the countdown array");
    if ((flags & AT_PP_EAT_IMPLICIT) && ((flags & (AT_PP_EAT | AT_PP_EAT_APPLY)) != (AT_PP_EAT | AT_PP_EAT_APPLY) || a.occ_in < 0 || a.occ_in > 1))
        return fail("at_pp_step: %s", "EAT_IMPLICIT needs EAT and EAT_APPLY in the same call and a valid occ_in");
    if ((a.nbr_ptr == nullptr) != (a.nbr_idx == nullptr)) return fail("at_pp_step: %s", "nbr_ptr and nbr_idx go together");

    if ((flags & AT_PP_OCCUPANCY) && n > 0) {
        AT_LAUNCH(pp_occupancy_kernel, blocks_for(n, a.grid), PP_THREADS, s, m, a, make_fastdiv((uint32_t)m.width));
        if (check("pp_occupancy_kernel launch")) return 1;
    }
    if (agent_flags && n > 0) {
        // 16-byte vector access when the arrays allow it (torch allocations do; odd views fall back to scalar access)
        auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
        const int vec_pred = aligned(a.pred_cell) && aligned(a.pred_energy) && aligned(a.pred_id);
        const int vec_prey = aligned(a.prey_cell) && aligned(a.prey_energy) && aligned(a.prey_id);
        const int64_t longest = a.n_pred > a.n_prey ? a.n_pred : a.n_prey;
        AT_LAUNCH(pp_agent_kernel, blocks_for((longest + 3) / 4 + 3, a.grid), PP_THREADS, s, m, a, make_fastdiv((uint32_t)m.width), vec_pred, vec_prey);
        if (check("pp_agent_kernel launch")) return 1;
    }
    if (cell_flags) {
        const int64_t threads = ((m.n_storage ? m.n_storage : m.n_cells) + 3) >> 2;
        if (a.countdown_f64) AT_LAUNCH(pp_cell_kernel<double>, blocks_for(threads, a.grid), PP_THREADS, s, m, a);
        else AT_LAUNCH(pp_cell_kernel<float>, blocks_for(threads, a.grid), PP_THREADS, s, m, a);
        if (check("pp_cell_kernel launch")) return 1;
    }
    return 0;
}

int at_pp_pack_cells(const int64_t* coords, int64_t n, int32_t max_y, int64_t n_cells, int32_t* cells, int32_t* error_flag, void* stream)
{
    if (n <= 0) return 0;
    if (!coords || !cells || max_y <= 0) return fail("at_pp_pack_cells: %s", "bad arguments");
    AT_LAUNCH(pp_pack_cells_kernel, blocks_for(n, 0), PP_THREADS, reinterpret_cast<cudaStream_t>(stream), coords, n, max_y, n_cells, cells, error_flag);
    return check("pp_pack_cells_kernel launch");
}

int at_pp_unpack_cells(const int32_t* cells, int64_t n, int32_t max_y, int64_t* coords, void* stream)
{
    if (n <= 0) return 0;
    if (!coords || !cells || max_y <= 0) return fail("at_pp_unpack_cells: %s", "bad arguments");
    AT_LAUNCH(pp_unpack_cells_kernel, blocks_for(n, 0), PP_THREADS, reinterpret_cast<cudaStream_t>(stream), cells, n, max_y, coords);
    return check("pp_unpack_cells_kernel launch");
}

int at_pp_pack_stage(const void* stage, int32_t stage_dtype, int64_t n_cells, uint8_t* info, void* stream)
{
    if (n_cells <= 0 || !stage || !info || (stage_dtype != 0 && stage_dtype != 1)) return fail("at_pp_pack_stage: %s", "bad arguments");
    const int64_t n_bytes = at_pp_info_bytes(n_cells);
    AT_LAUNCH(pp_pack_stage_kernel, blocks_for(n_bytes, 0), PP_THREADS, reinterpret_cast<cudaStream_t>(stream), stage, stage_dtype, n_cells, n_bytes, info);
    return check("pp_pack_stage_kernel launch");
}

int at_pp_unpack_stage(const uint8_t* info, int64_t n_cells, void* stage, int32_t stage_dtype, void* stream)
{
    if (n_cells <= 0 || !stage || !info || (stage_dtype != 0 && stage_dtype != 1)) return fail("at_pp_unpack_stage: %s", "bad arguments");
    AT_LAUNCH(pp_unpack_stage_kernel, blocks_for(n_cells, 0), PP_THREADS, reinterpret_cast<cudaStream_t>(stream), info, n_cells, stage, stage_dtype);
    return check("pp_unpack_stage_kernel launch");
}

int at_pp_census(const at_pp_launch* launch, int64_t n_cells, int64_t* out3, void* stream)
{
    if (!launch || !out3 || !launch->info) return fail("at_pp_census: %s", "bad arguments");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const cudaError_t e = cudaMemsetAsync(out3, 0, 24, s);
    if (e != cudaSuccess) return fail("cudaMemsetAsync: %s", cudaGetErrorString(e));
    const int64_t n = launch->n_pred + launch->n_prey;
    const int64_t work = n > ((n_cells + 3) >> 2) ? n : ((n_cells + 3) >> 2);
    AT_LAUNCH(pp_census_kernel, blocks_for(work, 0), PP_THREADS, s, *launch, n_cells, reinterpret_cast<unsigned long long*>(out3));
    return check("pp_census_kernel launch");
}

}  // extern "C"
