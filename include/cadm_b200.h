/*
 * cadm_b200.h -- C ABI of the B200-native CEM/MPC planning engine.
 *
 * This is the drop-in boundary for ONE hot path of younggyoseo/CaDM: the per-decision rollout of
 * n_candidates x n_particles action sequences through the probabilistic-ensemble dynamics MLP (and the
 * CaDM context encoder) over the planning horizon.  In the reference that path is a TensorFlow 1.15
 * graph behind `compile_function` (cadm/utils/tensor_utils.py:6-11: sess.run(outputs, feed_dict)); the
 * entry points below are what a binding for that boundary needs.  Each one cites what it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - All `const float*` / `float*` / `int32_t*` arguments are DEVICE pointers unless the name ends in
 *     `_host`.  The caller owns every buffer it passes; the engine copies what it keeps.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls are asynchronous with
 *     respect to the host unless stated otherwise.
 *   - Return value: 0 on success, negative CADM_ERR_* otherwise; cadm_last_error() gives the text.
 *   - A handle is bound to the CUDA device that was current at cadm_plan_create() and is not thread-safe.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with CADM_ERR_CUDA.
 *
 * Tensor layouts are row-major with the index order written in brackets.
 *   E ensemble members, p particles, n candidates (global), n_local = n / world candidates of this rank,
 *   m environments planned at once, h horizon, D obs dim, P processed-obs dim, A action dim, C context
 *   dim, K history length, k elites.
 */
#ifndef CADM_B200_H
#define CADM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CADM_ABI_VERSION 1

/* environment closures baked into the rollout epilogue (obs_preproc / obs_postproc / tf_reward_fn) */
#define CADM_ENV_HALFCHEETAH   0  /* cadm/envs/half_cheetah_env.py:46-56,82-88 (also half_cheetah_cripple_env.py:60-96) */
#define CADM_ENV_ANT           1  /* cadm/envs/ant_env.py:52-59,89-98 */
#define CADM_ENV_SLIM_HUMANOID 2  /* cadm/envs/slim_humanoid_env.py:39-46,95-111 */
#define CADM_ENV_CARTPOLE      3  /* cadm/envs/classic_control.py:94-101,154-166 (discrete actions) */
#define CADM_ENV_PENDULUM      4  /* cadm/envs/classic_control.py:209-218,284-291 */

/* arithmetic of the MLP contractions */
#define CADM_PREC_FP32   0  /* fp32 FFMA: the exact-fp32 path */
#define CADM_PREC_TC_3X  1  /* tcgen05 tensor cores, fp16 hi/lo split, 3 MMAs per product (fp32-class accuracy) */
#define CADM_PREC_TC_1X  2  /* tcgen05 tensor cores, single fp16 pass (fast; does NOT meet the 1e-4 bar) */

/* how a particle is paired with a context-encoder member */
#define CADM_CTX_REFERENCE 0  /* reproduce cadm/dynamics/core/utils.py:433-439 exactly (quirks Q2, Q3) */
#define CADM_CTX_MATCHED   1  /* dynamics member e sees encoder member e (what training does, utils.py:377) */

#define CADM_OK            0
#define CADM_ERR_ARG      -1
#define CADM_ERR_CUDA     -2
#define CADM_ERR_STATE    -3
#define CADM_ERR_UNSUPPORTED -4

/* Mirrors the constructor arguments of MLPEnsembleCEMDynamicsModel
 * (cadm/dynamics/mlp_cadm_ensemble_cem_dynamics.py:26-54) plus the CEM constants the reference hard-codes
 * (cadm/dynamics/core/utils.py:111-116). */
typedef struct CadmConfig {
    int32_t struct_size;     /* sizeof(CadmConfig), for ABI checking */
    int32_t env_id;          /* CADM_ENV_* */
    int32_t obs_dim;         /* D */
    int32_t proc_obs_dim;    /* P */
    int32_t act_dim;         /* A */
    int32_t ctx_dim;         /* C; 0 = no context encoder (PE-TS / vanilla) */
    int32_t hist_len;        /* K (ignored when ctx_dim == 0) */
    int32_t hidden;          /* width of the hidden layers (reference: 200) */
    int32_t n_hidden;        /* number of hidden layers (reference: 4) */
    int32_t enc_hidden[3];   /* context encoder hidden widths (reference: 256,128,64) */
    int32_t ensemble;        /* E */
    int32_t particles;       /* p, multiple of E */
    int32_t candidates;      /* n (global) */
    int32_t horizon;         /* h */
    int32_t m_max;           /* largest number of environments planned in one call */
    int32_t deterministic;   /* 1: delta = mean (vanilla DM), 0: PE-TS sampling */
    int32_t discrete;        /* 1: discrete actions (random shooting only; one-hot input) */
    int32_t num_elites;      /* reference: 50 */
    int32_t cem_iters;       /* reference: 5 */
    float   alpha;           /* reference: 0.1 */
    int32_t precision;       /* CADM_PREC_* */
    int32_t rank;            /* candidate shard of this process */
    int32_t world;           /* number of shards; candidates % world == 0 */
    int32_t context_layout;  /* CADM_CTX_* */
    float   max_torque;      /* pendulum only (classic_control.py:213); 0 -> 2.0 */
} CadmConfig;

int         cadm_abi_version(void);
/* text of the last error on this handle (handle == NULL: last error of cadm_plan_create) */
const char* cadm_last_error(const void* handle);

/* Builds the engine: replaces the graph construction in MLPEnsembleCEMDynamicsModel.__init__
 * (mlp_ensemble_cem_dynamics.py:86-189 / mlp_cadm_ensemble_cem_dynamics.py:107-342). */
int cadm_plan_create(const CadmConfig* cfg, void** handle);
int cadm_plan_destroy(void* handle);

/* Dynamics-MLP variables.  W[l] is [E, in_l, out_l], b[l] is [E, 1, out_l] (create_dense_layer,
 * core/utils.py:635-647); l = 0..n_hidden-1 hidden layers, then output_mu, then output_logvar
 * (n_layers = n_hidden + 2); max_logvar / min_logvar are [D] (core/utils.py:70-71).
 * Replaces the variable assignment of load() (mlp_ensemble_cem_dynamics.py:333-342); call again after fit(). */
int cadm_plan_set_weights(void* handle, const float* const* W, const float* const* b, int32_t n_layers,
                          const float* max_logvar, const float* min_logvar, void* stream);

/* Context-encoder variables cp_hidden_0..2, cp_output (core/utils.py:594-612): W[l] [E, in, out], b[l] [E, 1, out]. */
int cadm_plan_set_encoder(void* handle, const float* const* W, const float* const* b, int32_t n_layers, void* stream);

/* The normalisation placeholders of get_action (mlp_cadm_ensemble_cem_dynamics.py:345-356):
 * obs [P], act [A], delta [D], cp_obs [D*K], cp_act [A*K]; the cp_* pointers may be NULL when ctx_dim == 0. */
int cadm_plan_set_norm(void* handle,
                       const float* obs_mean, const float* obs_std,
                       const float* act_mean, const float* act_std,
                       const float* delta_mean, const float* delta_std,
                       const float* cp_obs_mean, const float* cp_obs_std,
                       const float* cp_act_mean, const float* cp_act_std, void* stream);

/* _get_context_pred (mlp_cadm_ensemble_cem_dynamics.py:337-342,369-380):
 * cp_obs [m, D*K], cp_act [m, A*K] -> ctx [E, m, C]. */
int cadm_encode_context(void* handle, int32_t m, const float* cp_obs, const float* cp_act, float* ctx, void* stream);

/* One model step in the training-graph layout -- the reference's compiled `_get_pred`
 * (mlp_ensemble_cem_dynamics.py:185-189) plus the sampled next state.
 * obs [E, B, D], act [E, B, A] (discrete: one-hot), ctx [E, B, C] or NULL, eps [E, B, D] or NULL (NULL and not
 * deterministic: Philox stream EPS of `seed`, row id e*B+b) -> next_obs, mu, logvar [E, B, D] (any may be NULL). */
int cadm_predict(void* handle, int32_t B, const float* obs, const float* act, const float* ctx,
                 const float* eps, uint64_t seed, float* next_obs, float* mu, float* logvar, void* stream);

/* The horizon loop alone (core/utils.py:137-168 / 431-472) for explicit action sequences.
 * obs [m, D], actions [m, n_local, h, A], ctx_raw [E, m, C] or NULL, eps [h, E, R, D] or NULL with
 * R = (p/E) m n and the reference row order (NULL: Philox), `it` selects the context pairing of CEM iteration it
 * -> particle_returns [m, n_local, p], states [h, m, n_local, p, D] or NULL (state after each step). */
int cadm_rollout(void* handle, int32_t m, int32_t it, const float* obs, const float* actions, const float* ctx_raw,
                 const float* eps, uint64_t seed, float* particle_returns, float* states, void* stream);

/* ---- one CEM decision, phase by phase (the multi-rank form; the host all-gathers between phases) ---- */

/* Start a decision: stores obs [m, D], init_mean / init_var [m, h, A] and runs the context encoder on
 * cp_obs [m, D*K], cp_act [m, A*K] (NULL when ctx_dim == 0).  core/utils.py:121-128, 400-407. */
int cadm_cem_begin(void* handle, int32_t m, const float* obs, const float* cp_obs, const float* cp_act,
                   const float* init_mean, const float* init_var, void* stream);

/* CEM iteration `it`: sample this rank's candidates (core/utils.py:131-135), roll them out (:137-168) and
 * average over particles (:170) into this rank's slice of the returns buffer.
 * z [iters, m, n, h, A] or NULL (NULL: Philox stream Z of `seed`); eps [iters, h, E, R, D] or NULL. */
int cadm_cem_rollout(void* handle, int32_t it, uint64_t seed, const float* z, const float* eps, void* stream);

/* Device buffer [world, m, n_local] of candidate returns; rank r writes slice r.  With world > 1 the host
 * all-gathers it in place between cadm_cem_rollout and cadm_cem_refit (ncclAllGather, one per iteration). */
float* cadm_cem_returns_buffer(void* handle);
int64_t cadm_cem_returns_slice_elems(void* handle);   /* m * n_local of the decision in flight */

/* ---- fused all-gather over peer memory (optional; replaces the host-side ncclAllGather above) ----
 * Every rank exports its exchange block (returns buffer + arrival flags, one CUDA allocation) as a CUDA IPC handle,
 * the host exchanges the `world` handles (any transport; cadm_b200/parallel.py uses torch.distributed), and each rank
 * attaches them.  From then on cadm_cem_refit does the exchange itself: the CTA of environment mi averages this rank's
 * particle returns (core/utils.py:170), stores the slice into every rank's returns buffer over NVLink, publishes the
 * iteration's epoch in flag (rank, mi) of every rank and waits on its own flags before the top-k -- no scatter kernel,
 * no collective call and no host synchronisation between the phases.  With the peers attached the single-call forms
 * (cadm_plan_cem, cadm_plan_cem_host, cadm_session_act) work at world > 1 as well: every rank calls them with the same
 * inputs and arrives at the same plan.  All ranks must be on one node and must call in lockstep (they do: the plan is
 * data-independent).  The caller makes sure every rank has attached before any rank starts a decision (a barrier after
 * cadm_peer_attach).  The device-side wait is bounded (option "peer_timeout_ms", default 30 s): if a peer never delivers,
 * cadm_cem_finish of that decision fails with CADM_ERR_CUDA; option "peer_clear_timeout" forgets the report once the
 * caller has re-synchronised the ranks. */
#define CADM_IPC_HANDLE_BYTES 64
int cadm_peer_export(void* handle, void* ipc_handle_out /* CADM_IPC_HANDLE_BYTES */);
int cadm_peer_attach(void* handle, const void* ipc_handles /* world x CADM_IPC_HANDLE_BYTES, rank order */, int32_t count);
int cadm_peer_enabled(void* handle);

/* top_k + gather + refit + EMA (core/utils.py:171-182) from the gathered returns buffer. */
int cadm_cem_refit(void* handle, int32_t it, void* stream);

/* Results of the decision: mean / var [m, h, A] (mean is the UNCLIPPED optimal_action_var, core/utils.py:184),
 * returns [iters, m, n], elites [iters, m, k] int32 (global candidate ids).  Any pointer may be NULL. */
int cadm_cem_finish(void* handle, float* mean, float* var, float* returns, int32_t* elites, void* stream);

/* The whole decision: begin + iters x (rollout, refit) + finish (world == 1, or world > 1 with the peers attached).
 * This is `_get_cem_action` (mlp_ensemble_cem_dynamics.py:173-178 / mlp_cadm_...:320-328). */
int cadm_plan_cem(void* handle, int32_t m, const float* obs, const float* cp_obs, const float* cp_act,
                  const float* init_mean, const float* init_var, uint64_t seed, const float* z, const float* eps,
                  float* mean, float* var, float* returns, int32_t* elites, void* stream);

/* Same with HOST buffers: copies the inputs to the device, runs the decision, copies `action_host`
 * [m, h, A] = clip(mean, -1, 1) back (mlp_ensemble_cem_dynamics.py:205-206) and synchronises the stream.
 * This is the call behind DynamicsModel.get_action(); pinned host memory makes the copies asynchronous. */
int cadm_plan_cem_host(void* handle, int32_t m, const float* obs_host, const float* cp_obs_host,
                       const float* cp_act_host, const float* init_mean_host, const float* init_var_host,
                       uint64_t seed, float* action_host, void* stream);

/* ---- sampler-side state on the device (the step either side of the decision) ----
 * What cadm/samplers/sampler.py keeps in NumPy between two get_actions() calls lives in the engine instead: the warm-start
 * plan `prev_sol` (sampler.py:52,118-119), the constant `init_var` = 0.25 (:53), the K-step history buffers that feed the
 * context encoder and their fill counters (:94-97,164-178), and the per-episode resets (:55-57,190-195).  A control step
 * is then  cadm_session_act (H2D of obs [m, D], one decision, D2H of the clipped first actions [m, A])  followed, after
 * the environment step, by  cadm_session_observe (H2D of next_obs [m, D] and done [m]; asynchronous).  world == 1, or
 * world > 1 with the peers attached (every rank keeps the same state and calls in lockstep). */
int cadm_session_reset(void* handle, int32_t m, const uint8_t* mask_host /* [m] or NULL = all */, void* stream);
int cadm_session_act(void* handle, int32_t m, const float* obs_host, uint64_t seed, float* action_host, void* stream);
/* history entry: state_diff == 0: obs (sampler.py:166-177); 1: next_obs - obs, subtracted in fp32 (run_cadm_pets.py:137
 * default); 2: `next_obs_host` holds the entry itself -- for a host whose environments hand out float64 observations and
 * that wants the reference's rounding, float32(next_obs - obs) with the subtraction in float64 */
int cadm_session_observe(void* handle, int32_t m, const float* next_obs_host, const uint8_t* done_host, int32_t state_diff,
                         void* stream);
/* copies of the state for inspection (any pointer may be NULL): prev_sol [m, h, A], history [m, D*K] / [m, A*K], counts [m] */
int cadm_session_state(void* handle, int32_t m, float* prev_sol_host, float* hist_obs_host, float* hist_act_host,
                       int32_t* counts_host, void* stream);

/* Random shooting, `_get_rs_action` (core/utils.py:186-246 / 490-561).  u: explicit draws, [m, n_local, h, A]
 * uniform(-1,1) actions (continuous) or NULL (Philox); for discrete models u_int [m, n_local, h] int32 or NULL.
 * world == 1 only.  -> action [m, A] (continuous, unclipped) or action_int [m], returns [m, n], best [m]. */
int cadm_plan_rs(void* handle, int32_t m, const float* obs, const float* cp_obs, const float* cp_act,
                 uint64_t seed, const float* u, const int32_t* u_int, const float* eps,
                 float* action, int32_t* action_int, float* returns, int32_t* best, void* stream);

/* Switch the arithmetic of the MLP contractions at run time (CADM_PREC_*); both weight images are kept packed. */
int cadm_set_precision(void* handle, int32_t precision);

/* Device self-test of the tensor-core path: out[128, N] = X[128, K] * W[K, N] computed with exactly the operand layouts,
 * descriptors and hi/lo split of the rollout kernel (terms = 3: CADM_PREC_TC_3X arithmetic, 1: CADM_PREC_TC_1X).
 * K <= 208, N a multiple of 16 <= 208.  Synchronises.  Diagnostic only. */
int cadm_selftest_tc_gemm(const float* X, const float* W, int32_t K, int32_t N, int32_t terms, float* out, void* stream);

/* Same for the swapped-operand kernel (rollout_tcs.cu): out[rows, N] = X[rows, K] * W[K, N] with the weights on the MMA's
 * M axis (two M = 128 tiles when N > 128), the rows on its N axis (MN-major B operand) and `kps` K16 blocks per weight
 * stage.  rows in {16, 32, 48, 64}, K <= 208, N <= 256.  Synchronises.  Diagnostic only. */
int cadm_selftest_tcs_gemm(const float* X, const float* W, int32_t rows, int32_t K, int32_t N, int32_t kps, int32_t terms,
                           float* out, void* stream);

/* Tuning knobs of the tensor-core path (the defaults are what bench.py measures):
 *   "tc_variant"  0 = pick by the measured wave-cost model (all rows on one kernel, or full waves of 128-row tiles + one wave of
 *                 small swapped-operand tiles for the rest: two launches), 1 = 128-row tiles (rollout_tc.cu), 2 = swapped
 *                 operands (rollout_tcs.cu), 3 = CTA pairs (rollout_tcp.cu: an experiment, slower; reference architecture only)
 *   "tcs_rows"    rows per tile of the swapped kernel: 0 = pick, else 16 / 32 / 48 / 64
 *   "tcs_kps"     K16 blocks per weight stage of the swapped kernel's image, 1..4 (before cadm_plan_set_weights)
 *   "tcs_skew"    start-delay step in cycles that de-phases the CTAs of the swapped kernel (0 = off)
 *   "trace"       1 = record the clock64 phase trace of CTA 0 (cadm_debug_trace); selects the kernel instantiation that has
 *                 the trace and the diagnostic switches compiled in (the production instantiation carries neither)
 *   "env_offset"  index of this engine's first environment in a larger, environment-sharded decision: added to the local
 *                 environment index in every Philox counter, so that a block of environments planned alone draws the same
 *                 numbers as inside the whole decision (cadm_b200/parallel.py EnvShardedPlanner)
 *   "pdl"         0 = launch the sample -> rollout -> refit chain without programmatic dependent launch (process-wide; also
 *                 CADM_PDL=0); default 1: the next kernel's prologue runs under its predecessor
 *   "peer_timeout_ms"     bound of the device-side wait for a peer's slice (fused all-gather), default 30000
 *   "peer_clear_timeout"  forget a reported peer timeout (after the caller has re-synchronised the ranks) */
int cadm_set_option(void* handle, const char* name, int32_t value);

/* Diagnostic micro-benchmark: n_mma back-to-back tcgen05.mma (M=128, N, K=16, fp16) on resident shared-memory operands,
 * round-robin over n_acc accumulators (1, 2, 4 or 8; 1 = each MMA depends on the previous one), in the operand layouts of
 * rollout_tc.cu (swapped = 0) or rollout_tcs.cu (swapped = 1).  a_lbo bits 16+ select the issue-loop style (0: elect per
 * MMA, 1: one elected thread); background bit 0 = a concurrent bulk-copy stream into shared memory, bit 1 = a commit every
 * 6 MMAs.  cycles_host[0] = issue time, [1] = time until the commit is observed (clock64), [2] = 8 KB copies completed. */
int cadm_selftest_tc_rate(int32_t N, int32_t n_mma, int32_t a_lbo, int32_t n_acc, int32_t swapped, int32_t background,
                          int64_t* cycles_host);

/* Diagnostic micro-benchmark of the swapped kernel's MMA stream: `iters` rounds of 4 weight stages (kps K16 blocks of R
 * weight rows each, 2 MMAs per block) issued by one elected thread; mode bit 0 = bulk copies stream into the ring slots
 * meanwhile, bit 1 = alternate accumulators, bit 2 = epilogue-like TMEM / shared-memory traffic from 16 warps.
 * cycles_host[0] = issue time, [1] = completion time (clock64), [2] = bytes streamed. */
int cadm_selftest_tcs_rate(int32_t rows, int32_t iters, int32_t R, int32_t kps, int32_t mode, int64_t* cycles_host);

/* Diagnostic: clock64 trace of CTA 0 of the last tensor-core rollout launched with timing enabled, [step][32] slots
 * (see rollout_tc.cu); copies `count` int64 values to the host.  Synchronises. */
int cadm_debug_trace(void* handle, int64_t* out_host, int32_t count);

/* Number of kernels this handle has launched so far (bench.py reports it as gpu_launches). */
int64_t cadm_launch_count(void* handle);
/* Name of the rollout kernel variant in use, e.g. "rollout_f32<32>" (for reports). */
const char* cadm_kernel_name(void* handle);
/* Duration in ms of the rollout kernels of the last cadm_plan_cem*, measured with CUDA events on `stream`
 * when timing is enabled (cadm_set_timing(handle, 1)); sum over the iterations.  Synchronises. */
int   cadm_set_timing(void* handle, int32_t on);
float cadm_last_rollout_ms(void* handle);

/* ---------------------------------------------------------------------------------------------------------------------
 * fit() on the device.  The training graph of the reference -- ensemble dynamics MLP, Gaussian NLL with
 * the soft-bounded log-variance, max / min logvar regulariser, per-layer L2 (mlp_ensemble_cem_dynamics.py:86-170,
 * core/utils.py:43-97, 635-647); for CaDM the context encoder trained end to end through the forward model and the
 * deterministic backward model weighted by back_coeff (mlp_cadm_ensemble_cem_dynamics.py:108-317, core/utils.py:251-372,
 * 569-624); tf.train.AdamOptimizer -- as hand-written forward / backward / Adam kernels (csrc/trainer.cu).  What the reference
 * feeds per minibatch through sess.run (mlp_ensemble_cem_dynamics.py:262-283, mlp_cadm_ensemble_cem_dynamics.py:478-527)
 * is here an index matrix into a dataset that stays on the device for the whole fit().
 *
 * Flat parameter vector (cadm_train_set_params / get_params / get_grads / adam_state), every tensor row-major:
 *   encoder (ctx_dim > 0):  for each layer  W [E, in, out], b [E, out]      ((D + A) K -> enc_hidden... -> C; relu, linear out)
 *   forward model:          for each hidden layer  W [E, in, H], b [E, H]   (in = P + A + C for the first; swish)
 *                           heads  W [E, H, 2 D] = [output_mu | output_logvar] column blocks, b [E, 2 D]
 *   max_logvar [D], min_logvar [D]
 *   backward model (has_back): hidden layers and heads like the forward model
 * ------------------------------------------------------------------------------------------------------------------- */
typedef struct CadmTrainConfig {
    int32_t struct_size;          /* sizeof(CadmTrainConfig) */
    int32_t env_id;               /* CADM_ENV_*: obs_preproc applied to the gathered observations */
    int32_t obs_dim;              /* D */
    int32_t proc_obs_dim;         /* P */
    int32_t act_dim;              /* A */
    int32_t ctx_dim;              /* C; 0 = PE-TS / vanilla model (no encoder) */
    int32_t hist_len;             /* K */
    int32_t hidden;               /* H */
    int32_t n_hidden;
    int32_t enc_hidden[3];        /* reference: 256, 128, 64 (0 ends the list) */
    int32_t ensemble;             /* E */
    int32_t deterministic;        /* 1: loss = mse (vanilla DM) */
    int32_t has_back;             /* backward model present (back_coeff > 0; CaDM only) */
    float   back_coeff;
    float   weight_decay_coeff;
    float   learning_rate;
    float   weight_decays[8];     /* hidden layer i: [i]; both heads: [n_hidden]  (core/utils.py:46-69) */
    float   context_weight_decays[4];   /* encoder hidden layer i: [i]; output layer: [number of hidden layers] */
    float   adam_beta1, adam_beta2, adam_eps;   /* tf.train.AdamOptimizer defaults: 0.9, 0.999, 1e-8 */
} CadmTrainConfig;

int         cadm_train_create(const CadmTrainConfig* cfg, void** handle);
int         cadm_train_destroy(void* handle);
const char* cadm_train_last_error(const void* handle);
int64_t     cadm_train_param_count(void* handle);
int64_t     cadm_train_launch_count(void* handle);
int cadm_train_set_params(void* handle, const float* flat_host, int64_t n);
int cadm_train_get_params(void* handle, float* flat_host, int64_t n);
/* gradient of the last training step, same layout (tests compare it with autograd and finite differences) */
int cadm_train_get_grads(void* handle, float* flat_host, int64_t n);
/* Adam slots m, v and step count t: set != 0 uploads, set == 0 downloads.  They live in the handle, so they persist across
 * fit() calls as long as the handle does (the reference creates the optimiser once, in the constructor). */
int cadm_train_adam_state(void* handle, int32_t set, float* m_host, float* v_host, int64_t n, int64_t* t_inout);
/* get_normalization_stats() order: obs mean/std [P], act mean/std [A], delta mean/std [D], cp_obs mean/std [D K],
 * cp_act mean/std [A K], back_delta mean/std [D]; count = 6 (PE-TS), 10 (CaDM) or 12 (CaDM with backward model). */
int cadm_train_set_norm(void* handle, const float* const* stats_host, int32_t count);
/* Uploads a dataset: which = 0 training rows, 1 validation rows.  obs / obs_next / delta / back_delta [rows, D], act [rows, A],
 * cp_obs [rows, D K], cp_act [rows, A K]; the arguments the model does not use may be NULL. */
int cadm_train_set_dataset(void* handle, int32_t which, int64_t rows, const float* obs_host, const float* act_host,
                           const float* delta_host, const float* obs_next_host, const float* back_delta_host,
                           const float* cp_obs_host, const float* cp_act_host);
/* One minibatch: idx_host [E, B] rows of dataset `which` (member e trains on its own bootstrap rows).  train != 0: forward,
 * backward, one Adam step; train == 0: losses only.  losses_host [4] = mse_loss, recon_loss, back_mse_loss, mu_loss
 * (mlp_ensemble_cem_dynamics.py:150-167, mlp_cadm_ensemble_cem_dynamics.py:266-314).  Synchronises. */
int cadm_train_step(void* handle, int32_t which, const int32_t* idx_host, int32_t B, int32_t train, float* losses_host);

#ifdef __cplusplus
}
#endif
#endif /* CADM_B200_H */
