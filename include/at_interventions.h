/* at_interventions.h — C ABI of libat_interventions.so: the covid model's quarantine, testing and SEIRMS-progression
 * substeps on one B200.
 *
 * Replaces, for `agent_torch/models/covid/substeps` (classes the reference ships but does not register):
 *   quarantine/action.py:16-68        StartCompliance, BreakCompliance              -> at_iv_quarantine  AT_IV_DECIDE
 *   quarantine/transition.py:31-122   UpdateQuarantineStatus (end, start, break)    -> at_iv_quarantine  AT_IV_APPLY
 *   testing/action.py:29-53           AcceptTest                                    -> at_iv_testing     AT_IV_DECIDE
 *   testing/transition.py:41-175      UpdateTestStatus (results, enrolment)         -> at_iv_testing     AT_IV_APPLY
 *   seirms_progression/transition.py:48-181  SEIRMSProgression                      -> at_iv_seirms_progression
 *   core/helpers/soft.py:70-76        discrete_sample (Gumbel-softmax, tau=0.1, hard)  -> `u < p` on the counter hash
 *   core/helpers/distributions.py     StraightThroughBernoulli                         -> `u < p`, gradient by the caller
 *
 * Conventions (same as at_seirm.h): every function returns 0 on success, non-zero on error (text from
 * at_iv_last_error()), never throws; the library allocates nothing — all buffers are caller-owned DEVICE memory; all
 * work is stream-ordered on `stream` (a cudaStream_t) and CUDA-graph capturable.  No CPU fallback.
 *
 * Layout: one float32 per agent and property, exactly the reference's state tensors (`value * torch.ones(shape)`,
 * core/initializer.py:112): is_quarantined and awaiting_test_result are float32 COUNTERS (they drift below zero,
 * quirk I1 in oracle/interventions_port.py), the dates are integral float32.  `*_in` and `*_out` may alias.  A
 * thread owns 4 consecutive agents (16-byte loads), and writes a 16-byte group only if one of its values changed
 * (when in and out alias), so a step moves ~24 B per untouched agent instead of 44.
 *
 * Draws: uniform k * 2^-24 from SplitMix64 of (seed, step, agent | stream << 36) — quarantine stream 1 (start = top
 * 24 bits, break = bits 8..31), testing stream 2, mortality stream 3 (csrc/at_noise.cuh, oracle/noise.py) — or
 * caller tensors (`u_*` non-NULL).
 */
#ifndef AT_INTERVENTIONS_H
#define AT_INTERVENTIONS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AT_IV_ABI_VERSION 1

#define AT_IV_DECIDE 1 /* run the policies: draw, write the action tensors that are non-NULL */
#define AT_IV_APPLY 2  /* run the transition; without DECIDE the actions are READ from the action tensors */

typedef struct at_iv_quarantine_args {
    int32_t abi_version;
    int32_t flags;
    int32_t t;                  /* state["current_step"] */
    int32_t stage_f64;          /* 1: `stage` is double (the reference's CSV path), 0: float */
    int64_t n;
    int64_t id_offset;          /* global id of agent 0 (addresses the draws) */
    uint64_t seed;
    const uint64_t* seed_ptr;   /* device; overrides `seed` when non-NULL (graph replay with another seed) */
    float p_start;              /* environment/quarantine_start_prob */
    float p_break;              /* environment/quarantine_break_prob */
    float quarantine_days;      /* simulation_metadata.quarantine_days */
    float infinity_time;        /* simulation_metadata.INFINITY_TIME */
    const void* stage;          /* DECIDE */
    const float* u_start;       /* optional injected uniforms */
    const float* u_break;
    const float* is_q_in;
    const float* start_date_in; /* APPLY */
    uint8_t* start_action;      /* 0/1 bytes (torch.bool storage) */
    uint8_t* break_action;
    float* is_q_out;            /* APPLY */
    float* start_date_out;
} at_iv_quarantine_args;

typedef struct at_iv_testing_args {
    int32_t abi_version;
    int32_t flags;
    int32_t t;
    int32_t reserved;
    int64_t n;
    int64_t id_offset;
    uint64_t seed;
    const uint64_t* seed_ptr;
    float p_comply;             /* environment/test_compliance_prob */
    float ineligible_days;      /* simulation_metadata.test_ineligible_days */
    float result_delay_days;    /* simulation_metadata.test_result_delay_days */
    float infinity_time;
    const float* u_comply;      /* optional injected uniforms */
    const float* is_q;          /* DECIDE */
    float* accept_action;       /* float 0/1 like the reference's soft logical_and (testing/action.py:49-51) */
    const float* awaiting_in;   /* APPLY */
    const float* result_date_in;
    const float* re_eligible_in;
    float* awaiting_out;
    float* result_date_out;
    float* re_eligible_out;
} at_iv_testing_args;

typedef struct at_iv_seirms_args {
    int32_t abi_version;
    int32_t t;
    int64_t n;                  /* >= 2 (the reference indexes rows 0 and 1, quirk I6) */
    int64_t id_offset;
    uint64_t seed;
    const uint64_t* seed_ptr;
    float infected_to_recovered_time;
    float infinity_time;        /* num_steps_per_episode + 20 (transition.py:28-30) */
    float recovered_to_susceptible_time;
    int32_t num_steps;          /* length of daily_deaths */
    const float* m_value;       /* device, 1 float: mortality rate */
    const float* u_death;       /* optional injected uniforms */
    const float* stage_in;
    const float* next_stage_time_in;
    float* stage_out;
    float* next_stage_time_out;
    uint8_t* move_code;         /* optional, per agent: 0 none, 1 E->I, 2 R->S, 3 I->4 (survived), 4 I->3 (died): what the
                                   caller's backward needs (d new_stage / d stage = 1, 2, 0, 3, 2.5; quirk I7) */
    int64_t* counts;            /* device, 4 x int64, ZEROED by the caller: [due infected, died, survived, unused] */
    float* daily_deaths;        /* (num_steps,) in place: [t] += float(due infected) * M */
} at_iv_seirms_args;

int at_iv_abi_version(void);
const char* at_iv_last_error(void);
int64_t at_iv_sizeof_quarantine_args(void);
int64_t at_iv_sizeof_testing_args(void);
int64_t at_iv_sizeof_seirms_args(void);

/* One launch: policies and/or transition of the quarantine substep, per `flags`. */
int at_iv_quarantine(const at_iv_quarantine_args* args, void* stream);
/* One launch: policy and/or transition of the testing substep. */
int at_iv_testing(const at_iv_testing_args* args, void* stream);
/* Two launches: the per-agent pass, then a one-thread epilogue (rows 0/1 of next_stage_time, daily_deaths[t]). */
int at_iv_seirms_progression(const at_iv_seirms_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif
