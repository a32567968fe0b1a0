/*
 * swarm_b200.h -- C ABI of the B200-native swarm-environment hot path (libswarm_b200.so).
 *
 * The reference (allentran/golds-rl-gym) is pure Python and has NO FFI of its own; its
 * boundary for this path is three Python contracts.  Each entry point
 * below names the reference interface it stands in for.  A reference maintainer binds these
 * with ctypes; the package `golds-rl-gym_b200/` is exactly that binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name starts with `host_`;
 *   - the library allocates nothing, never synchronises (except the *_host call) and launches
 *     on the caller's stream; all entry points are thread-safe for distinct buffers AND distinct
 *     streams (the two-kernel step keeps one internal side stream + fork/join event pair per
 *     (device, caller stream); two host threads must not step on the SAME stream concurrently);
 *   - return value: 0 = ok, negative = SwarmStatus error (swarm_strerror);
 *   - layouts are C order.  x:(E,N,2) f64, xa:(E,A,2) f64, actions:(E,A,2),
 *     grid:(E,G,G,2) f32 indexed [env][x_bin][y_bin][channel], positions:(E,A,2) u8.
 *
 * Arithmetic: the O(N) integrator state is FP64 like the reference's numpy arrays
 * (fed_gym/envs/multiagent.py:30-86); the O(N^2) pair forces (multiagent.py:88-115) are FP32.
 */
#ifndef SWARM_B200_H
#define SWARM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWARM_ABI_VERSION 3

/* opaque: a cudaStream_t passed as a pointer-sized handle (0 = legacy default stream) */
typedef void* swarm_stream_t;

typedef enum SwarmStatus {
    SWARM_OK = 0,
    SWARM_ERR_NULL = -1,        /* a required pointer is NULL                          */
    SWARM_ERR_SIZE = -2,        /* unsupported E/N/A/G (see swarm_validate)            */
    SWARM_ERR_LAUNCH = -3,      /* CUDA launch / runtime error (swarm_last_cuda_error) */
    SWARM_ERR_FLAGS = -4        /* inconsistent flags / missing optional buffer        */
} SwarmStatus;

/* fed_gym/envs/multiagent.py:8-21 (class constants) + fed_gym/__init__.py:21-33 (TimeLimit 128)
 * + fed_gym/agents/state_processors.py:17-23 (box).  POD, passed by pointer, copied at call. */
typedef struct SwarmParams {
    int32_t n_envs;             /* E : envs in THIS shard                                   */
    int32_t n_locusts;          /* N : SwarmEnv.N_LOCUSTS (80)                              */
    int32_t n_agents;           /* A : SwarmEnv.N_AGENTS (10)                               */
    int32_t grid_size;          /* G : SwarmStateProcessor.grid_size (84 for PAAC)          */
    int32_t n_burn_in;          /* SwarmEnv.N_BURN_IN (10)                                  */
    int32_t max_episode_steps;  /* gym TimeLimit (128); 0 = raw SwarmEnv, no limit          */
    int32_t math_mode;          /* 0 = fast (MUFU rsqrt/ex2/rcp), 1 = precise (IEEE)        */
    int32_t tuning;             /* 0 = automatic launch shape.  Otherwise (tests / experiments; results are bitwise
                                 * the same for every value): bits 4-5 = rasteriser placement (1 follower kernel,
                                 * 2 raster warps in the step kernel, 3 the step's own threads after the step);
                                 * all other bits must be 0                                                   */
    double noise;               /* NOISE 1e-4   */
    double gravity;             /* GRAVITY -1   */
    double wind;                /* WIND_SPEED 1 */
    double F;                   /* 0.5          */
    double L;                   /* 10           */
    double dt;                  /* 0.05         */
    double box_width;           /* WIDTH 3  -> x range [mean-1.5, mean+1.5]                 */
    double box_height;          /* HEIGHT 3 -> y range [0, 2*HEIGHT]                        */
    uint64_t seed;              /* Philox key                                               */
    int64_t env_id_offset;      /* global id of local env 0 (shard-invariant RNG streams)   */
} SwarmParams;

/* The env batch's persistent state (SwarmEnv.states / .agent_noise[t] / .particle_noise[t] /
 * TimeLimit._elapsed_steps).  noise_* hold the FROZEN row N_BURN_IN that every post-reset step
 * re-uses because SwarmEnv._step never advances self.t (multiagent.py:38,40). */
typedef struct SwarmState {
    double* x;                  /* (E,N,2) */
    double* xa;                 /* (E,A,2) */
    double* noise_x;            /* (E,N,2) unscaled N(0,1) */
    double* noise_a;            /* (E,A,2) unscaled N(0,1) */
    int32_t* elapsed;           /* (E) steps since reset   */
    uint32_t* episode;          /* (E) resets so far (Philox counter word) */
    uint32_t* work;             /* nullable (work_words uint32): scratch words, ZERO when handed over and zero again
                                 * after every call.  With at least 2 + E of them swarm_step may produce the
                                 * observation with a second kernel that follows the step on an internal
                                 * higher-priority stream (per-env ready flags in work[2..2+E)); with fewer (>= 2) it
                                 * only uses the work queue work[0..1]; with none, static env assignment and raster
                                 * warps inside the step kernel.  Not shared between concurrently running calls. */
    uint64_t work_words;        /* number of uint32 words behind `work` (0 if work == NULL); checked, never trusted */
} SwarmState;

/* The reference's random draws of one reset (multiagent.py:51-56), injected for parity tests.
 * Only rows 0..n_burn_in of the 138-row noise tables are ever read by the reference. */
typedef struct SwarmInjectedDraws {
    const double* x0;             /* (E,N,2)              np.random.rand            */
    const double* xa0;            /* (E,A,2)                                        */
    const double* burn_actions;   /* (E,n_burn_in,A,2)    N(0,1), unclipped         */
    const double* agent_noise;    /* (E,n_burn_in+1,A,2)                            */
    const double* particle_noise; /* (E,n_burn_in+1,N,2)                            */
} SwarmInjectedDraws;

#define SWARM_STEP_AUTO_RESET   1u  /* SwarmRunner._run: on done, reset and observe the new episode */
#define SWARM_STEP_CLIP_ACTIONS 2u  /* SwarmRunner.transform_actions_for_env, in place on actions   */
#define SWARM_STEP_ACTIONS_F64  4u  /* actions_f64 is used instead of actions_f32                   */
#define SWARM_STEP_NO_ACTION_WIND 8u  /* SwarmEnv._step(add_wind=False), multiagent.py:30,35-36: the step's
                                       * actions are used as they are (burn-in steps of an auto-reset still
                                       * add the wind, like the reference's _reset -> step)                */
#define SWARM_STEP_INKERNEL_RASTER 16u /* never use the follower kernel (see swarm_step); same results     */

typedef struct SwarmStepIO {
    float* actions_f32;         /* (E,A,2) PAAC shared_actions dtype (paac.py:269); in/out if CLIP  */
    double* actions_f64;        /* (E,A,2) gym-facade dtype; used when SWARM_STEP_ACTIONS_F64        */
    const double* noise_a;      /* nullable (E,A,2): per-step override of the frozen row            */
    const double* noise_x;      /* nullable (E,N,2)                                                 */
    float* reward;              /* (E)  = -mean_j |v_j|^2                                           */
    uint8_t* done;              /* (E)  reward >= 0 || ++elapsed >= max_episode_steps               */
    float* grid;                /* nullable (E,G,G,2): fused rasterise of the post-step state       */
    uint8_t* positions;         /* nullable (E,A,2); required iff grid != NULL                      */
    float* v_out;               /* nullable (E,N,2): pre-cutoff locust velocities (diagnostics)     */
    uint32_t flags;
    uint32_t reserved;
} SwarmStepIO;

int swarm_abi_version(void);
const char* swarm_strerror(int status);
const char* swarm_last_cuda_error(void);

/* Size / resource check for a parameter set; SWARM_OK or SWARM_ERR_SIZE.  No GPU work. */
int swarm_validate(const SwarmParams* p);

/* SwarmEnv._reset (multiagent.py:46-63): draws + n_burn_in burn-in steps; elapsed=0; episode+=1.
 * mask: nullable (E) u8, only envs with mask!=0 are reset.  draws: nullable -> Philox4x32-10
 * keyed by (seed, env_id_offset+e, episode) as documented in csrc/swarm_philox.cuh. */
int swarm_reset(const SwarmParams* p, const SwarmState* st, const uint8_t* mask,
                const SwarmInjectedDraws* draws, swarm_stream_t stream);

/* SwarmEnv._step (multiagent.py:30-44) + TimeLimit + (optionally) the SwarmRunner._run
 * auto-reset (emulator_runner.py:126-135) and SwarmStateProcessor.process_state
 * (state_processors.py:29-42) of the resulting state: one kernel, or -- for large swarms when
 * st->work is given -- the step kernel plus a rasteriser kernel that follows it concurrently on
 * an internal stream (joined back into `stream` before the call returns its work to it).
 * The call can be captured into a CUDA graph; instantiate such a graph with
 * cudaGraphInstantiateFlagUseNodePriority (PyTorch does), or the follower loses its priority
 * and runs after the step instead of next to it.
 * The follower spins on per-env flags the step kernel raises, so the two kernels must be able to
 * run side by side or step-first.  Eager launches guarantee it (the step is launched first); a tool
 * that re-orders or isolates graph nodes (graph-node profiling, some debuggers) does not: pass
 * SWARM_STEP_INKERNEL_RASTER or set the environment variable SWARM_B200_NO_FOLLOWER=1 there.  The
 * spin is bounded (a follower that sees no progress for ~20 s traps: the call then fails with a
 * CUDA launch error at the next synchronisation instead of hanging).
 * reset_draws: nullable; injected draws used by auto-reset instead of Philox. */
int swarm_step(const SwarmParams* p, const SwarmState* st, const SwarmStepIO* io,
               const SwarmInjectedDraws* reset_draws, swarm_stream_t stream);

/* The launch shape swarm_step would use for these arguments (no launch; needs the device for occupancy queries):
 * out = { force mode, filler warp (0/1), rasteriser placement (0 none, 1 follower kernel, 2 raster warps, 3 the step's
 * own threads), threads per CTA, CTAs, dynamic shared memory bytes, follower CTAs, kernel launches }. */
int swarm_step_plan(const SwarmParams* p, const SwarmState* st, const SwarmStepIO* io, int32_t out[8]);

/* Same call with HOST buffers for the per-step inputs/results; synchronises the stream before
 * returning.  With PINNED host buffers (cudaHostAlloc / torch pin_memory: device-visible under
 * UVA) the kernel reads host_actions and writes host_reward / host_done directly over PCIe --
 * no staging copies (io->actions_f32 / io->reward / io->done are then left untouched).  With
 * pageable memory it copies host_actions to io->actions_f32, runs swarm_step and copies
 * io->reward / io->done back.
 * host_grid / host_positions: nullable (both or neither; need io->grid): the compact observation
 * (E,G,G,2) f32 + (E,A,2) u8 is also copied back to the host (copy engine, after the step) -- the
 * numpy-out case of the reference's process_state. */
int swarm_step_host(const SwarmParams* p, const SwarmState* st, const SwarmStepIO* io,
                    const float* host_actions, float* host_reward, uint8_t* host_done,
                    float* host_grid, uint8_t* host_positions, swarm_stream_t stream);

/* Debug only, and only in a library built with -DSWARM_TRACE (the production build compiles the hooks out: they cost
 * 0.5-2.5 %): from now on every step / rasteriser kernel of this process records phase timestamps of each CTA's first
 * env into device_words (n_words uint64, zeroed by the caller; 32 words per record: for phase ph < 16 the global
 * timer in ns at [2 ph] and the SM cycle counter at [2 ph + 1]; record = CTA index of the step kernel, E + env for
 * the follower kernel; phases: see scripts/trace_step.py).  NULL switches it off.  Not thread-safe. */
void swarm_debug_trace(uint64_t* device_words, int64_t n_words);

/* Destroys the CUDA graphs swarm_step_host caches for the CALLING host thread (one per argument
 * set, at most 32).  They are also destroyed when the thread exits. */
void swarm_step_host_clear(void);

/* SwarmStateProcessor.process_state (state_processors.py:25-42): grid + uint8 agent cells.
 * box: nullable (E,4) f64 = [lo_x, hi_x, lo_y, hi_y], i.e. _get_bounding_box (:25-27) of
 * vstack([x, xa]).  n_agents may be 0 here (then xa/positions may be NULL): the box/grid of a
 * bare point set, as tests/env_tests.py:39 uses _get_bounding_box(state[0]).
 * The observation is numpy's bit for bit either way; with a box the window's centre is always taken by numpy's
 * sequential FP64 sum (the box reports it), without one -- and inside swarm_step -- by a parallel sum that is verified
 * against a rigorous error bound and replaced by the sequential one only when a point sits within that bound of a bin edge. */
int swarm_rasterize(const SwarmParams* p, const double* x, const double* xa,
                    float* grid, uint8_t* positions, double* box, swarm_stream_t stream);

/* SwarmRunner.get_local_states (emulator_runner.py:98-111) for the whole batch:
 * expanded (E,A,G,G,3) f32 = grid channels + one-hot at positions[e][a]. */
int swarm_expand_obs(const SwarmParams* p, const float* grid, const uint8_t* positions,
                     float* expanded, swarm_stream_t stream);

/* SwarmEnv.v_calculate (multiagent.py:88-115): v (E,N,2) f32 and reward (E) f32 from x, xa. */
int swarm_forces(const SwarmParams* p, const double* x, const double* xa,
                 float* v, float* reward, swarm_stream_t stream);

/* SwarmEnv.x_update (multiagent.py:70-75) on n particles, in place on x AND v like the
 * reference (v gets the ground cutoff); noise is the already-scaled additive term. */
int swarm_x_update(double* x, double* v, const double* noise, int64_t n, double dt, swarm_stream_t stream);

/* SwarmEnv.xv_cutoff (multiagent.py:77-86) on n particles, in place. */
int swarm_xv_cutoff(double* x, double* v, int64_t n, swarm_stream_t stream);

/* SwarmEnv.s (multiagent.py:65-68): out[i] = F exp(-r[i]/L) - exp(-r[i]), FP64. */
int swarm_s_potential(const double* r, double* out, int64_t n, double F, double L, swarm_stream_t stream);

/* SwarmRunner.transform_actions_for_env (emulator_runner.py:113-118), in place, n rows of 2. */
int swarm_clip_actions(float* actions, int64_t n_rows, float max_norm, swarm_stream_t stream);

/* Materialise the Philox draws that swarm_reset would use for each env's CURRENT episode
 * counter (st->episode), in the SwarmInjectedDraws layout, so a test can inject the very
 * same numbers into the reference.  Output pointers are the (non-const) draws buffers. */
int swarm_philox_draws(const SwarmParams* p, const SwarmState* st, double* x0, double* xa0,
                       double* burn_actions, double* agent_noise, double* particle_noise,
                       swarm_stream_t stream);

/* Raw Philox4x32-10 blocks for known-answer tests: out[i] = philox(ctr[i], key[i]). */
int swarm_philox_raw(const uint32_t* ctr /*(n,4)*/, const uint32_t* key /*(n,2)*/,
                     uint32_t* out /*(n,4)*/, int64_t n, swarm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SWARM_B200_H */
