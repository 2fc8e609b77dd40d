/* at_predprey.h — C ABI of libat_predprey.so: the predator-prey grid model's four substeps on one B200.
 *
 * Replaces, for `agent_torch/models/predator_prey` (BASELINE configs[3]):
 *   substeps/move.py:49-123   FindNeighbors + DecideMovement + UpdatePositions      -> AT_PP_MOVE
 *   substeps/eat.py:24-110    FindEatableGrass + EatGrass                            -> AT_PP_EAT (+ AT_PP_EAT_APPLY)
 *   substeps/hunt.py:24-92    FindTargets + HuntPrey                                 -> AT_PP_HUNT
 *   substeps/grow.py:24-46    GrowGrass                                              -> AT_PP_GROW
 *   core/helpers/environment.py:6-9  grid_network (dense adjacency)                  -> implicit lattice (width, n_cells)
 *
 * Conventions (same as at_seirm.h): every function returns 0 on success, non-zero on error (text from
 * at_pp_last_error()), never throws; the library allocates nothing — all buffers are caller-owned DEVICE memory;
 * all work is stream-ordered on `stream` (a cudaStream_t) and CUDA-graph capturable.  No CPU fallback.
 *
 * Data layout (SoA): an agent is (cell int32, energy float32) with cell = max_y * x + y (move.py:36).  A cell is one
 * INFO byte + one countdown (float32 or float64, as the population file gave it).  The info byte answers every
 * "what is at this cell" question of a step with ONE random access:
 *     bit 0  grown      growth_stage == 1
 *     bit 1  eaten      set by EAT, consumed (and cleared) by EAT_APPLY
 *     bit 2  a predator stands here   } occupancy of parity 0
 *     bit 3  a prey stands here       }
 *     bit 4, 5  the same for parity 1 (the occupancy of consecutive steps alternates between the two pairs)
 * The info array has at_pp_info_bytes(n_cells) bytes (n_cells rounded up to a multiple of 16), or
 * at_pp_storage_cells(...) bytes with tiled storage.
 */
#ifndef AT_PREDPREY_H
#define AT_PREDPREY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AT_PP_ABI_VERSION 4

/* at_pp_launch.flags — what one call does, always in this order:
 *   OCCUPANCY  mark the cells holding >= 1 predator / >= 1 prey in the occupancy bits of parity `occ_in` (needed by
 *              HUNT when the positions did not come from a MOVE of this library; the bits must be clear before)
 *   EAT, HUNT  per agent, on the positions as they are; HUNT reads the occupancy bits of parity `occ_in`
 *   MOVE       per agent: draw + step to a neighbouring cell if energy > 0, energy -= work; marks the occupancy
 *              bits of parity `occ_out` at the new cell (occ_out < 0: no marking)
 *   EAT_APPLY  per cell: eaten grass -> stage 0, countdown c + float32(regrowth_time - c)
 *   GROW       per cell: countdown -= 1, stage = (countdown <= 0)
 *   CLEAR      per cell: clear the occupancy bits of parity `occ_in` (their step is over)
 * Modifier EAT_IMPLICIT (with EAT and EAT_APPLY in the same call, occupancy of parity `occ_in` valid): a grown cell
 * holding a prey IS an eaten cell, so EAT sets no bit and EAT_APPLY derives it — one atomic less per eating prey. */
#define AT_PP_OCCUPANCY 1
#define AT_PP_EAT 2
#define AT_PP_HUNT 4
#define AT_PP_MOVE 8
#define AT_PP_EAT_APPLY 16
#define AT_PP_GROW 32
#define AT_PP_CLEAR 64
#define AT_PP_EAT_IMPLICIT 128

typedef struct at_pp_model {
    int32_t abi_version;
    int32_t width;            /* lattice: cell k touches k +- width and, inside a run of `width` cells, k +- 1 (grid_network) */
    int32_t max_y;            /* coordinates: x = cell / max_y, y = cell % max_y (move.py:36,43-44) */
    int32_t apply_pred_work;  /* 0 when the YAML output mapping drops the new predator energy (controller.py:73-81) */
    int32_t apply_prey_work;  /* 0 for the shipped config.yaml:281-285 (pred_energy listed twice) */
    int32_t tiles_per_row;    /* 0: info / countdown stored at the cell number.  ceil(width / 16): stored in tiles of 8 rows x 16
                                 columns (one 128-byte line of info bytes per tile), see at_pp_storage_cells */
    int64_t n_cells;
    float pred_work;          /* agents/predator/stride_work */
    float prey_work;          /* agents/prey/stride_work */
    float prey_nutrition;     /* agents/prey/nutritional_value: gained by a predator that hunts */
    float grass_nutrition;    /* objects/grass/nutritional_value: gained by a prey that eats */
    float regrowth_time;      /* objects/grass/regrowth_time */
    float reserved_f;
    int64_t n_storage;        /* cells as stored: n_cells (or 0) untiled, at_pp_storage_cells(width, n_cells, 1) tiled */
} at_pp_model;

typedef struct at_pp_launch {
    int32_t abi_version;
    int32_t flags;
    int32_t t_move;           /* step number that addresses the MOVE draws */
    int32_t countdown_f64;    /* 1: `countdown` is double, 0: float */
    uint64_t seed;            /* used when seed_ptr == NULL */
    const uint64_t* seed_ptr; /* device: lets a captured CUDA graph be replayed with another seed */
    const int32_t* t_ptr;     /* device, optional: t_move = *t_ptr + t_move (replayable graphs) */
    int64_t n_pred, n_prey;
    int32_t* pred_cell;
    float* pred_energy;
    int32_t* prey_cell;
    float* prey_energy;
    uint8_t* info;            /* [at_pp_info_bytes(n_cells)] cell info bytes, see above */
    void* countdown;          /* [n_cells] */
    int32_t occ_in;           /* parity (0/1) of the occupancy bits that describe the CURRENT positions */
    int32_t occ_out;          /* parity MOVE marks with the NEW positions; < 0: none */
    const int32_t* pred_id;   /* optional: the agent number that addresses the draws of slot i (the caller keeps the agents */
    const int32_t* prey_id;   /* sorted by cell for locality and records their original numbers here); NULL = slot number   */
    const int64_t* nbr_ptr;   /* optional CSR adjacency (ascending neighbours) for non-lattice networks; NULL = lattice */
    const int32_t* nbr_idx;
    int32_t grid;             /* CTAs per kernel; 0 = library default (multiple of the SM count) */
    int32_t reserved;
} at_pp_launch;

int at_pp_abi_version(void);
const char* at_pp_last_error(void);
int64_t at_pp_sizeof_model(void);
int64_t at_pp_sizeof_launch(void);
int64_t at_pp_info_bytes(int64_t n_cells);
/* Tiled storage: cell (row, col) of the lattice lives at ((row/8 * tiles_per_row + col/16) * 128 + (row%8) * 16 + col%16);
 * cells of the padding must hold info 0 and countdown +inf.  The pack/unpack helpers below work on cell-number order;
 * the caller permutes (agenttorch_b200/predprey.py does it with one gather per array at the boundary). */
int64_t at_pp_storage_cells(int32_t width, int64_t n_cells, int32_t tiled);

/* One call = up to three kernels (occupancy; agents; cells), see the flag list. */
int at_pp_step(const at_pp_model* model, const at_pp_launch* launch, void* stream);

/* (n,2) int64 [x, y] coordinates  <->  int32 cells;  (n_cells) stage values (int64 / float32, == 1 means grown)
 * <->  bit 0 of the info bytes (pack overwrites the whole byte).  `stage_dtype`: 0 = int64, 1 = float32. */
int at_pp_pack_cells(const int64_t* coords, int64_t n, int32_t max_y, int64_t n_cells, int32_t* cells, int32_t* error_flag, void* stream);
int at_pp_unpack_cells(const int32_t* cells, int64_t n, int32_t max_y, int64_t* coords, void* stream);
int at_pp_pack_stage(const void* stage, int32_t stage_dtype, int64_t n_cells, uint8_t* info, void* stream);
int at_pp_unpack_stage(const uint8_t* info, int64_t n_cells, void* stage, int32_t stage_dtype, void* stream);

/* Counts for reporting: out[0] = predators with energy > 0, out[1] = prey with energy > 0, out[2] = grown cells. */
int at_pp_census(const at_pp_launch* launch, int64_t n_cells, int64_t* out3, void* stream);

#ifdef __cplusplus
}
#endif
#endif
