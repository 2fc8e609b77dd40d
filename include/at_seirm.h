/*
 * at_seirm.h — C ABI of libat_seirm.so: the B200 (sm_100a) implementation of AgentTorch's
 * per-step SEIRM agent update.
 *
 * The reference has no native layer: its "FFI" for this path is the Python operator surface
 *   - NewTransmission.forward           agent_torch/models/covid/substeps/new_transmission/transition.py:131-252
 *       (PyG MessagePassing.propagate gather + scatter-add, _lam :52-85, StraightThroughBernoulli
 *        core/distributions/distributions.py:8-24, update_stages/.._times :102-129)
 *   - SEIRMProgression.forward          agent_torch/models/covid/substeps/seirm_progression/transition.py:90-121
 *   - torch autograd over both          (loss.backward() through an episode)
 * Each entry point below names the reference function(s) it replaces.  The library is called
 * from Python through ctypes by torch.library custom ops (agenttorch_b200/ops.py); see
 * INTEGRATION.md for the binding a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success, non-zero on error (message via
 * at_last_error()); nothing throws across the boundary.  The library is stateless: it allocates
 * nothing and owns nothing — every buffer is caller-allocated DEVICE memory (all pointers are
 * device pointers unless the comment says host) and every launch goes to the caller's stream,
 * so launches are legal inside CUDA-graph capture.
 */
#ifndef AT_SEIRM_H
#define AT_SEIRM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AT_SEIRM_ABI_VERSION 10
#define AT_MAX_SFS 16        /* SFSusceptibility entries (config.yaml has 9)                     */
#define AT_MAX_SFI 8         /* SFInfector entries (config.yaml has 5)                           */
#define AT_MAX_LAMGAMMA 84   /* lam_gamma_integrals entries; 1 + 3*K infectiousness codes fit a byte */
#define AT_MAX_BLOCKS 64     /* column blocks of the contact graph                               */
#define AT_MAX_LAYERS 8      /* network layers (distinct per-edge network types) of a multi-layer contact graph */

/* dtype tags for the timer / age arrays */
#define AT_TIMER_F32 0       /* reference layout: float32 holding integers (Q3)                  */
#define AT_TIMER_I16 1       /* packed engine layout                                             */
#define AT_AGE_F32 0
#define AT_AGE_U8 1

/* isolation decision W (MakeIsolationDecision.forward, new_transmission/action.py:51-67) */
#define AT_ISO_TENSOR 0      /* W read from `isolation` (float32, any value in [0,1])             */
#define AT_ISO_HASH_SOFT 1   /* W = uniform[0,1) from the counter hash (action.py:54)             */
#define AT_ISO_HASH_HARD 2   /* W = 1 with probability p_isolate else 0                           */
#define AT_ISO_NONE 3        /* W = 0                                                             */
/* exposure uniform U (stands in for torch.bernoulli, distributions.py:16) */
#define AT_U_TENSOR 0        /* U read from `uniform` (float32)                                   */
#define AT_U_HASH 1          /* U from the counter hash: top 24 bits of SplitMix64                */

/* contact-graph layouts (at_seirm_launch.layout) */
#define AT_LAYOUT_BCSR 0     /* row_start + seg_info + col: 32-row tasks (one column block) or 256-row tasks    */
#define AT_LAYOUT_STREAM 1   /* col + unit_tile only: flat per-source-block streams of self-describing entries    */
#define AT_STREAM_UNIT 256          /* entries per unit (1 KB): streams are padded to whole units per (block, tile)   */
#define AT_STREAM_MAX_TILE_ROWS 2048
#define AT_STREAM_MAX_BLOCK_SIZE ((1 << 21) - 1024)

/* what one launch does */
#define AT_DO_TRANSMISSION 1 /* substep '0' transition (NewTransmission.forward)                  */
#define AT_DO_PROGRESSION 2  /* substep '1' transition (SEIRMProgression.forward)                 */
#define AT_DO_ENCODE 4       /* write the infectiousness code/bitmap of step `t_encode`           */
#define AT_DO_INIT 8         /* derive the row bitmaps (susceptible/exposed/infected) from `stage`
                                instead of reading them (first launch on a state, or after the
                                caller edited the state)                                          */

#define AT_DO_LOOKAHEAD 16   /* with AT_DO_ENCODE: `t_encode` lies one step beyond the state this launch leaves, so the
                                code is computed for the stage every agent WILL have after the progression of step
                                t_encode-1 (E/I agents whose next_stage_time <= t_encode-1 move on; agents exposed in
                                between contribute lam_gamma[0]*... at t_encode, which must be 0 — the caller checks).
                                Lets the halo exchange of step t+2 run while step t+1 computes.           */

/* Model constants (simulation_metadata + environment tables of config.yaml). Host struct. */
typedef struct at_seirm_model {
    int32_t n_sfs, n_sfi, n_lam_gamma;
    int32_t exposed_to_infected_time;   /* EXPOSED_TO_INFECTED_TIME, transition.py:25-27      */
    int32_t infected_to_recovered_time; /* INFECTED_TO_RECOVERED_TIME, progression :31-33      */
    int32_t infinity_time;              /* num_steps_per_episode + 20, progression :28-30      */
    float i_bar;                        /* mean_interactions[dst[edge_type]] — one scalar when every edge has the same
                                           network type (Q1; network_from_file writes type 1 everywhere)            */
    float sfs[AT_MAX_SFS];
    float sfi[AT_MAX_SFI];
    float lam_gamma[AT_MAX_LAMGAMMA];
    /* multi-layer contact networks (transition.py:71-73: I_bar is gathered PER EDGE by the edge's network type k,
       I_bar_e = mean_interactions[dst[k]]): one I_bar per distinct type; n_layers <= 1 means "i_bar for every edge".
       Layers are carried by the stream layout only (unit_tile holds the layer of each unit in its top 8 bits). */
    int32_t n_layers;
    float i_bar_layers[AT_MAX_LAYERS];
} at_seirm_model;

/*
 * One launch of the step kernels.  Host struct, passed by pointer; pointers inside are device.
 *
 * Contact graph = column-blocked, destination-sorted CSR ("BCSR"), replacing the (2,E) COO edge_index that
 * the reference gathers through PyG (utils.py:129-148, transition.py:209-218):
 *   destinations (rows) are grouped in gather tasks of rows_per_task rows (32 when n_blocks == 1, 256 otherwise);
 *   sources are split in n_blocks ranges of block_size ids; segment s = block*n_gtasks + gtask holds the edges of
 *   the task's rows whose source lies in the block, row by row, each row in ORIGINAL edge order:
 *     seg_info[2*s]   = first entry / 4 (segments start 16-byte aligned),  seg_info[2*s+1] = number of edges,
 *     row_start[block][row] = offset of the row's first edge inside its segment (uint16; a row ends where the
 *     next one starts, the last one at the segment length),
 *     col[] = block-local source id c encoded as ((c >> 5) << 7) | (c & 31): byte offset of its bitmap word
 *     times 32, plus its bit — so that the shared-memory bitmap test needs no address arithmetic.
 *   Edges are used as written: directed, duplicates and self-loops kept (Q9).
 *
 * layout == AT_LAYOUT_STREAM (the layout for source sets larger than one shared-memory bitmap, and the default for
 *   unweighted graphs): rows are grouped in TILES of tile_rows rows (a power of two <= 2048; n_rows_pad is a multiple
 *   of it), sources in n_blocks ranges of block_size <= 2^21 - 1024 ids.  For every source block the edges whose
 *   source lies in it form ONE flat stream, tile after tile, each (block, tile) run padded to whole UNITS of
 *   AT_STREAM_UNIT entries with entries that can never hit.  A 32-bit entry describes itself:
 *     col[] = (bitmap word of the block-local source << 16) | (row_in_tile << 5) | (bit of the source in its word),
 *   unit_tile[u] = tile of unit u in its low 24 bits and the network layer of its edges in the top 8 (runs are
 *   per (block, layer, tile)), block_units[b] = first unit of block b (block_units[n_blocks] = all units).
 *   The gather kernel streams the units with 128-bit loads against the block's EXACT hot bitmap in shared memory and
 *   appends, for every hit on a susceptible row, the record (layer << 24) | (row_in_tile << 8) | source code to the row tile's record
 *   list (rec / tile_rec_base / tile_count).  The update kernel sums a tile's records into one 64-bit FIXED-POINT
 *   lambda per row — an integer sum, hence independent of the order in which the hits arrived (bitwise reproducible
 *   without any floating-point atomic) — and turns it back into float64.  Edge weights are not supported
 *   (edge_weight must be NULL).
 */
typedef struct at_seirm_launch {
    int32_t abi_version;        /* = AT_SEIRM_ABI_VERSION */
    int32_t flags;              /* AT_DO_* */
    int32_t t;                  /* current_step */
    int32_t t_encode;           /* step whose infectiousness is encoded (normally t+1) */
    /* graph */
    int64_t n_rows;             /* owned agents */
    int64_t n_rows_pad;         /* n_rows rounded up to 256 (leading dimension of row_start, dirty, lam_part) */
    int64_t n_sources;          /* length of code_in (multiple of 1024) */
    int64_t row_offset;         /* global id of local row 0 (multiple of 32): slice of this rank in
                                   code/hot_bitmap, and the agent id fed to the counter hash */
    int32_t n_blocks, block_size;   /* block_size multiple of 1024 */
    int32_t rows_per_task, edges_per_chunk;  /* 32 or 256 (BCSR; ignored for STREAM); TMA chunk capacity: 512, 768 or 1024 edges */
    const uint16_t* row_start;  /* [n_blocks][n_rows_pad] */
    const int32_t* seg_info;    /* [n_blocks*n_gtasks][2] = (first entry / 4, edges) */
    const int32_t* col;         /* encoded source ids, segment-padded, + >= 1024 ints of tail padding */
    const float* edge_weight;   /* parallel to col, or NULL when all weights are 1 (utils.py:144) */
    /* agent state, updated in place and only where it changes */
    float* stage;               /* n_rows */
    void* next_stage_time;      /* n_rows, dtype timer_dtype */
    void* infected_time;        /* n_rows */
    const void* age;            /* n_rows, dtype age_dtype */
    int32_t timer_dtype, age_dtype;
    /* row bitmaps (one bit per owned row, word = task): stage == 0 / == 1 / == 2.  Read unless AT_DO_INIT,
       always rewritten. */
    uint32_t* sus_bitmap;       /* n_rows_pad/32 words */
    uint32_t* exp_bitmap;
    uint32_t* inf_bitmap;
    /* infectiousness of step t (read) and of step t_encode (written for owned rows), per SOURCE id */
    const uint8_t* code_in;     /* n_sources */
    const uint32_t* hot_bitmap_in;  /* n_sources/32 (+4) words: code != 0 */
    uint8_t* code_out;          /* written at [row_offset, row_offset+n_rows) */
    uint32_t* hot_bitmap_out;
    /* optional, hashed mode on one GPU: the folded filter (hash_words words, word i = OR of the hot words i, i+M, ...,
       followed by one word that is non-zero iff any source is hot).  _in: read instead of folding hot_bitmap_in in
       every CTA; _out: written by the last CTA to finish, from hot_bitmap_out.  NULL = fold in the kernel prologue. */
    const uint32_t* hot_filter_in;
    uint32_t* hot_filter_out;
    /* noise */
    int32_t iso_mode, u_mode;
    const float* isolation;     /* n_rows or NULL */
    const float* uniform;       /* n_rows or NULL */
    uint64_t seed;
    const uint64_t* seed_ptr;   /* optional device word overriding `seed` (keeps CUDA graphs seed-agnostic) */
    float p_isolate;
    int32_t _pad0;
    /* parameters (device scalars so that no host sync is needed) */
    const float* r_value;       /* R of this step = r_value[r_index] */
    int64_t r_index;
    const float* m_value;       /* M = m_value[0] */
    /* outputs */
    float* daily_infected;      /* [t] += sum(ne)              (transition.py:235) */
    float* daily_deaths;        /* [t] += M * sum(stage*rd/2)  (progression :41-61) */
    float* dne_dr;              /* n_rows or NULL: d ne_i / d R, saved for backward */
    float* sum_dne_dr;          /* [t] = sum_i dne_dr, or NULL */
    float* death_count;         /* [t] = sum(stage*rd/2), or NULL (d daily_deaths / d M) */
    uint8_t* recovered_step;    /* n_rows or NULL: set to t where the I->R transition fires */
    /* scratch for n_blocks > 1 (per-block partial lambdas), else NULL */
    double* lam_part;           /* [n_blocks][n_rows_pad][2]: (lambda, d lambda / dR) of the rows flagged in dirty */
    int32_t layout;             /* AT_LAYOUT_* */
    int32_t tile_rows;          /* AT_LAYOUT_STREAM: rows per tile (power of two, 32..2048), else 0 */
    uint32_t* dirty;            /* [n_blocks][n_rows_pad/32]: which rows of lam_part hold a value this step */
    /* workspace: at_seirm_workspace_bytes() bytes, zeroed once by the caller, reset by every launch */
    void* workspace;
    int32_t* error_flag;        /* or NULL; set non-zero on a lam_gamma index out of range (Q10) */
    int32_t grid;               /* CTAs, 0 = one per SM */
    int32_t tasks_per_group;    /* scheduling granule (1..8 gather tasks), 0 = library default */
    int32_t warps_per_cta;      /* 0 = as many (<= 16) as shared memory allows */
    int32_t hash_words;         /* 0: exact bitmap, col holds encoded ids.  > 0 (power of two, n_blocks == 1): the
                                   shared-memory bitmap is a hashed filter of hash_words words (bit src mod 32*hash_words),
                                   col holds RAW source ids, false positives are rejected by code_in[src] == 0 */
    /* AT_LAYOUT_STREAM only (NULL / 0 otherwise) */
    const int32_t* unit_tile;   /* [units] tile of every unit */
    const int32_t* block_units; /* [n_blocks + 1] first unit of every source block */
    const int64_t* tile_rec_base; /* [tiles] first record slot of every tile (slots = edges of the tile's rows) */
    uint32_t* tile_count;       /* [tiles] records of the tile this step, all zero between launches */
    uint32_t* rec;              /* [edges] hit records (layer << 24) | (row_in_tile << 8) | source code, grouped by tile */
    int32_t lam_deg_bits;       /* ceil(log2(largest in-degree + 1)): head room of the fixed-point sums */
    int32_t xchg_slot;          /* 0..2: which "my slice holds an infectious agent" word of the workspace this launch sets
                                   (one per code buffer in flight; at_seirm_exchange_sparse reads and clears it) */
    float* dne_e_dr;            /* n_rows or NULL (stream layout): the part of dne_dr that comes from sources in stage E
                                   (SFInfector[1]); with dne_dr it gives d ne / d SFInfector[1], [2] — saved for backward */
    int32_t* over_state;        /* or NULL: device word, > 0 once NO agent of ANY rank is in E or I ("the epidemic is over": nothing
                                   is or will become infectious; agents with a fractional stage never are).  The stream gather then
                                   returns at once.  Zeroed by the caller before the first launch of an episode; maintained by the
                                   launches themselves when over_local != 0 (one GPU), else by at_seirm_exchange_sparse */
    int32_t over_local;         /* != 0: one GPU — the update launch itself maintains *over_state (nobody in E or I after this step) */
    int32_t reserve_sms;        /* stream layout, != 0: the step kernels leave part of every SM free (lean gather entry point with 56
                                   registers, one update CTA per SM fewer), so that a halo exchange issued on another stream runs
                                   beside them instead of waiting for them (they otherwise take every register of every SM) */
    int32_t* step_skipped;      /* or NULL: [t] = 1 if this (stream-layout) launch found the epidemic over on entry and therefore
                                   did NOT touch dne_dr / dne_e_dr of step t (they are all zero by meaning; the backward pass skips
                                   the step), else 0 */
} at_seirm_launch;

/* ---- queries (host only) ---- */
int at_seirm_abi_version(void);
const char* at_last_error(void);
/* sizeof() of the two host structs, so that a foreign-language mirror can be checked at load time */
int64_t at_seirm_sizeof_launch(void);
int64_t at_seirm_sizeof_model(void);
int64_t at_seirm_workspace_bytes(void);
/* largest block_size (sources per column block) whose "hot" bitmap the gather stages in shared memory next to
   `warps` warps' TMA buffers of `edges_per_chunk` edges, for a model with n_sfs susceptibility classes and
   n_lam_gamma integrals (the term tables share the same shared memory) */
int64_t at_seirm_max_block_size(int32_t n_sfs, int32_t n_lam_gamma, int32_t edges_per_chunk, int32_t warps);
/* the same for AT_LAYOUT_STREAM (no staging buffers: the stream goes through registers); never more than
   AT_STREAM_MAX_BLOCK_SIZE */
int64_t at_seirm_stream_max_block_size(int32_t n_sfs, int32_t n_lam_gamma, int32_t warps);

/*
 * at_seirm_step — one step (or part of one, per the flags) of the agent update on `stream`.
 * Replaces NewTransmission.forward (transition.py:131-252, including the PyG gather/scatter-add and the
 * straight-through Bernoulli), SEIRMProgression.forward (seirm_progression/transition.py:90-121), and
 * produces what their autograd graph would have saved (dne_dr, death_count, recovered_step) without any
 * per-edge intermediate.  BCSR with n_blocks == 1: one fused launch; n_blocks > 1 or AT_LAYOUT_STREAM: a gather
 * launch over all column blocks followed by the update launch.
 */
int at_seirm_step(const at_seirm_model* model, const at_seirm_launch* launch, void* stream);

/*
 * at_seirm_episode_backward — reverse pass of an episode, replacing torch autograd over the
 * T x (transmission, progression) graph (SURVEY.md 3.6, A.1 "reverse form"):
 *   grad_r[t] = g_daily_infected[t] * sum_dne_dr[t]
 *             + sum_i dne_dr[t,i] * g_daily_deaths[recovered_step[i]] * M / 2
 *   grad_m[0] = sum_t g_daily_deaths[t] * death_count[t]
 * dne_dr is (T, ld) row-major with the first n_rows entries of each row used.  step_skipped (or NULL): steps whose dne_dr row
 * was left untouched by the forward because the epidemic was over (at_seirm_launch.step_skipped); they count as all zero.
 * infected_time (or NULL; timer_dtype as in at_seirm_launch): the agents' infected_time as the episode left it.  An agent is an
 * exposure candidate only up to the step of its (last) exposure, so dne_dr[t,i] == 0 for t > infected_time[i] and those
 * entries are not fetched.  Pass NULL when something other than the transmission substep may have made an agent susceptible
 * again during the episode (an intervention hook): every entry is fetched then.
 */
int at_seirm_episode_backward(int32_t num_steps, int64_t n_rows, int64_t ld,
                              const float* dne_dr, const uint8_t* recovered_step,
                              const float* sum_dne_dr, const float* death_count,
                              const float* g_daily_infected, const float* g_daily_deaths,
                              const float* m_value, float* grad_r, float* grad_m,
                              void* workspace, int32_t grid, const int32_t* step_skipped,
                              const void* infected_time, int32_t timer_dtype, void* stream);

/*
 * at_seirm_hot_from_code — multi-GPU helper: after the ranks' code slices have been all-gathered (the per-step halo
 * exchange of boundary-agent infectiousness; no reference equivalent, core/distributed_runner.py:218-228 is a stub),
 * rebuild hot_bitmap (bit = code != 0) and, if hot_filter != NULL, the hashed filter of hash_words (+1 flag) words.
 */
int at_seirm_hot_from_code(const uint8_t* code, int64_t n_sources, uint32_t* hot_bitmap, uint32_t* hot_filter,
                           int32_t hash_words, void* stream);

/*
 * at_seirm_pull_codes — the halo exchange fused with the rebuild, for code arrays in symmetric (peer-mapped) memory:
 * peer_code is a HOST array of `world` device pointers, peer_code[r] = base of rank r's code buffer as mapped into
 * this process (own buffer for r == rank).  After a device-side barrier (all ranks have written their slice), one
 * kernel reads the other ranks' slices over NVLink (P2P loads), stores them into the local `code` array and sets the
 * hot bitmap / hashed filter bits in the same pass.
 */
int at_seirm_pull_codes(uint8_t* code, int64_t n_sources, uint32_t* hot_bitmap, uint32_t* hot_filter, int32_t hash_words,
                        const void* const* peer_code, int32_t world, int32_t rank, void* stream);

/*
 * at_seirm_exchange_sparse — the per-step halo exchange as ONE kernel with the rendezvous inside (no separate barrier,
 * no communicator).  Each rank owns one symmetric (peer-mapped) buffer per step parity of at_seirm_exchange_bytes(n)
 * bytes: [code : n][hot bitmap : n/8 (+64)][arrival flags]; point at_seirm_launch.code_* / hot_bitmap_* at its first
 * two parts.  peer_buf is a HOST array of `world` device pointers (peer_buf[rank] == local_buf).  The kernel signals
 * "my slice is written" to every peer, waits for theirs, pulls 32-byte code sectors only where a peer's bitmap word is
 * non-zero (a slice whose owner reports no infectious agent, and of which no stale copy is held, is skipped altogether),
 * and folds the whole bitmap into hot_filter (may be NULL).  sync_state: 32 zero-initialised device words per buffer,
 * owned by this call sequence.  step_workspace: the at_seirm_launch.workspace of the step kernel that wrote this rank's
 * slice (it records there whether the slice holds any infectious agent; NULL = scan the slice).  max_ctas: 0 = as many
 * 256-thread CTAs as fit in one wave; > 0 caps the grid (an exchange that runs BESIDE the step kernels of the next step
 * should only occupy the SMs those leave free, at_seirm_launch.reserve_sms).  over_state (or NULL): the arrival signal also says
 * whether the rank still holds an agent in E or I; an exchange in which no rank does increments *over_state, any other resets
 * it, and once it has reached over_after (the number of code buffers in rotation: every stale copy has been cleared by then) the
 * call returns without a rendezvous — identically on every rank, because every rank saw the same signals.  The caller zeroes
 * it when an episode starts.  A peer that never arrives sets *error_flag = 2 after ~20 s.
 * halo_mask (or NULL = everything): one bit per 32-source group (= per bitmap word and 32-byte code sector), n_sources / 1024
 * words: set iff some edge of THIS rank's rows has its source in the group.  Groups outside the mask are neither read from
 * the peers nor written locally (their local copy stays zero), so what crosses NVLink is the infectiousness of the rank's
 * boundary (halo) agents only — the whole population only when every group is referenced, as on a uniformly random graph.
 * Replaces nothing in the reference (core/distributed_runner.py:218-228 is `pass`).
 */
int64_t at_seirm_exchange_bytes(int64_t n_sources);
int at_seirm_exchange_sparse(uint8_t* local_buf, int64_t n_sources, uint32_t* hot_filter, int32_t hash_words,
                             void* const* peer_buf, int32_t world, int32_t rank, uint32_t* sync_state, void* step_workspace,
                             int32_t* error_flag, int32_t max_ctas, int32_t* over_state, int32_t over_after,
                             const uint32_t* halo_mask, void* stream);

/*
 * at_noise_fill — materialise the counter-hash noise (same values the step kernel generates in
 * registers) for the tensor-mode API: what = 0 exposure uniform U, 1 isolation uniform,
 * 2 hard isolation decision (uses p_isolate).  Stands in for torch.rand (action.py:54).
 */
int at_noise_fill(float* out, int64_t n, int64_t id_offset, uint64_t seed, int32_t step,
                  int32_t what, float p_isolate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AT_SEIRM_H */
