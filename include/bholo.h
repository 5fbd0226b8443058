/* bholo.h -- C ABI of the B200 hologram reward / direct-binary-search engine.
 *
 * The reference (songyb111-gachon/binary-hologram-reinforcement-learning) has no
 * FFI of its own: its de-facto operator interface for this path is five calls
 * into the absent third-party package torchOptics plus the host loops around
 * them.  Every entry point below cites the reference lines it replaces.
 *
 * Conventions
 *   - return 0 on success, < 0 on error; bh_last_error() gives the message.
 *     Nothing throws or aborts.
 *   - pointers are HOST pointers unless the parameter is named d_* or an
 *     on_host flag says otherwise.  Outputs are caller-allocated.
 *   - a context owns all device memory of its E environments; it is not
 *     thread-safe, distinct contexts are independent.
 *   - work is enqueued on the stream given to bh_set_stream (default: the
 *     legacy default stream); functions with host outputs synchronise that
 *     stream only.
 *   - an action is the flat index  frame * N*N + row * N + col  of the
 *     reference (env.py:158-161).
 */
#ifndef BHOLO_H
#define BHOLO_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bh_ctx bh_ctx;

/* accept rules */
#define BH_RULE_ENV   0   /* keep iff psnr_after - previous >= 0   (env.py:191)   */
#define BH_RULE_DBS   1   /* keep iff psnr_after > previous        (DBS.py:273)   */
#define BH_RULE_NEVER 2   /* score and revert    (dbs-1024-1024-24-6464.py:371)   */
#define BH_RULE_ALWAYS 3  /* keep unconditionally (the "keep" branch itself, DBS.py:273-291) */

/* propagation method */
#define BH_METHOD_ASM     0
#define BH_METHOD_FRESNEL 1

typedef struct bh_result {      /* one evaluated flip; 40 bytes */
    double  psnr_after;         /* PSNR with the flip applied                     */
    double  d_sii;              /* change of sum(I^2)                             */
    double  d_sit;              /* change of sum(I*T)                             */
    int64_t action;             /* the evaluated action, -1 when the slot is idle */
    int32_t accept;             /* decision under the rule                        */
    int32_t sgn;                /* +1: pixel 0 -> 1, -1: 1 -> 0                   */
} bh_result;

int bh_abi_version(void);

/* Last error of ctx (or of the calling thread when ctx is NULL). */
const char* bh_last_error(const bh_ctx* ctx);

/* Create a context for n_env environments of F frames of N x N pixels in G
 * colour groups of F/G frames, wavelength wl[g] per group.
 * Replaces the constants baked into the reference envs (env.py:27-28,124;
 * env_1024_24.py:29-30,135-147).  pad: 1 = circular (FFT side P = N),
 * 2 = linear (P = 2N); relative: 1 = scale-invariant loss of tt.relativeLoss. */
int bh_create(bh_ctx** out, int device, int n_env, int N, int F, int G,
              const double* wl, double dx, double z, int pad, int relative, int method);
int bh_destroy(bh_ctx* ctx);
int bh_set_stream(bh_ctx* ctx, void* cuda_stream);

/* Target image of one env, float [G][N][N]  (env.py:106-107 target.cuda()). */
int bh_set_target(bh_ctx* ctx, int env, const float* target, int on_host);

/* Upload a binary stack int8 [F][N][N] and run the full propagation:
 * tt.Tensor + tt.simulate + abs()**2 + mean(dim=1) + relativeLoss
 * (env.py:123-132; env_1024_24.py:140-166).  Fills U, I and the loss sums. */
int bh_load_state(bh_ctx* ctx, int env, const int8_t* state, int on_host);

/* Copy the complete device state of env src (target, hologram, fields, reconstruction, loss
 * sums) to env dst: group rollouts that start from one reset state (env_group.py + GRPO-style
 * groups) pay one propagation per group, not per member. */
int bh_clone_env(bh_ctx* ctx, int src, int dst);

/* Re-propagate from the device-resident state (bounds fp32 drift of the
 * incremental updates); same arithmetic as bh_load_state. */
int bh_resync(bh_ctx* ctx, int env);

/* Current PSNR / MSE (env.py:131-132) and, optionally, the three sums
 * (sum I^2, sum I*T, sum T^2). */
int bh_get_metrics(bh_ctx* ctx, int env, double* psnr, double* mse, double* sums3);

/* Score n single-pixel flips against the CURRENT state without changing it
 * (env_group.py:96-119; dbs-1024-1024-24-6464.py:337-371; range.py:294-335).
 * env_ids may be NULL: every candidate then targets `env`. */
int bh_eval_flips(bh_ctx* ctx, int env, int64_t n, const int32_t* env_ids,
                  const int64_t* actions, double* psnr_after);

/* Score EVERY single-pixel flip of every frame against the current state in one
 * call (the exhaustive loops of dbs-1024-1024-24-6464.py:330-395, range.py:294-335,
 * DBS_1024_24-128.py:310-373): psnr_after[frame][row][col], F*N*N doubles, equal to
 * what bh_eval_flips returns for the same action.  Evaluated as cross-correlations
 * with the (even) impulse-response kernels through the FFT passes -- O(N^2 log N)
 * per frame instead of N^2 delta passes.  Device pointer when on_host = 0. */
int bh_sweep_all(bh_ctx* ctx, int env, double* psnr_after, int on_host);

/* bh_sweep_all followed by the decile statistics of dbs-1024-1024-24-6464.py:377-395 on the
 * device: pre_model float [F][N][N] (host) is the pre-binarisation output that defines the
 * bins, edges the 11 bin edges (half-open bins, last one closed).  attempted / improved:
 * int64[10]; gains: double[10] (sum of psnr_after - previous over the improving flips).
 * psnr_after may be NULL when only the statistics are wanted. */
int bh_sweep_stats(bh_ctx* ctx, int env, const float* pre_model, const double* edges,
                   int64_t* attempted, int64_t* improved, double* gains, double* psnr_after);

/* One environment step for n distinct environments: score the flip, keep or
 * revert it under `rule` (env.py:154-196 / DBS_1024_24.py:313-422 body).
 * results: n records. */
int bh_step_batch(bh_ctx* ctx, int n, const int32_t* env_ids, const int64_t* actions,
                  int rule, bh_result* results);

/* Host-side bookkeeping of a vectorised env step (env.py:158-167,184-196,214), done in the
 * same call so that an E-env step is one foreign call.  All arrays are HOST arrays of n
 * entries indexed like env_ids (state / state_record: [E][stride] int8 mirrors indexed by
 * env id).  After the flips are scored: steps += 1; state_record[action] += 1; psnr_change =
 * psnr_after - prev_psnr; rewards = psnr_change * reward_scale; kept flips toggle the state
 * mirror, count in flips and move prev_psnr; event[i] = 1 where the success / max_steps
 * branches of env.py:216-254 have to run (evaluated by the caller). */
typedef struct bh_vec_book {
    int8_t*        state;           /* may be NULL */
    int8_t*        state_record;    /* may be NULL */
    int64_t        stride;          /* elements per env in the two mirrors */
    double*        prev_psnr;       /* [E] in/out, indexed by env id */
    const double*  init_psnr;       /* [E] */
    int64_t*       steps;           /* [E] in/out */
    int64_t*       flips;           /* [E] in/out */
    const double*  t_psnr_diff;     /* [E] */
    const double*  t_psnr;          /* [E] */
    const int64_t* max_steps;       /* [E] */
    double         reward_scale;    /* RW = 800 (env.py:29) */
    double*        rewards;         /* [n] out */
    double*        psnr_change;     /* [n] out */
    double*        psnr_diff;       /* [n] out: psnr_after - init_psnr */
    int64_t*       last_candidate;  /* [E] out: -1 if kept, else the rejected action */
    uint8_t*       event;           /* [n] out */
} bh_vec_book;

int bh_vec_step(bh_ctx* ctx, int n, const int32_t* env_ids, const int64_t* actions, int rule,
                bh_result* results, const bh_vec_book* book);

/* The bookkeeping half of bh_vec_step on its own: applies env.py:163-167,184-196,214 to n scored
 * steps (results from bh_step_batch or read back from bh_step_batch_device).  Host only -- no
 * CUDA call, usable (and tested) without a GPU.  Returns 0, or -1 for an incomplete book
 * (message via bh_last_error(NULL)). */
int bh_vec_book_update(int n, const int32_t* env_ids, const int64_t* actions,
                       const bh_result* results, const bh_vec_book* book);

/* Same with DEVICE pointers and no synchronisation (inputs resident in HBM). */
int bh_step_batch_device(bh_ctx* ctx, int n, const int32_t* d_env_ids,
                         const int64_t* d_actions, int rule, bh_result* d_results);

/* Score only, DEVICE pointers, no synchronisation; n <= bh_max_tasks(). */
int bh_eval_flips_device(bh_ctx* ctx, int env, int n, const int32_t* d_env_ids,
                         const int64_t* d_actions, bh_result* d_results);
int bh_max_tasks(const bh_ctx* ctx);

/* Unconditionally apply one flip (DBS "keep" branch, DBS.py:273-291). */
int bh_commit_flip(bh_ctx* ctx, int env, int64_t action);

/* Greedy direct binary search over `order` (DBS.py:247-294,
 * DBS_1024_24.py:313-422): visit candidates in order, keep a flip iff the PSNR
 * strictly improves.  Runs on the device in speculative batches of up to
 * k_spec candidates (<= 0: adaptive); the decision sequence is that of the
 * sequential loop.  accepted: n bytes; psnr_trace (nullable): n doubles, PSNR
 * of every evaluated candidate; resync_every (<= 0: never): re-propagate after
 * that many accepted flips. */
int bh_dbs_run(bh_ctx* ctx, int env, const int64_t* order, int64_t n, int k_spec,
               int64_t resync_every, uint8_t* accepted, double* psnr_trace,
               int64_t* n_accepted, double* final_psnr);

/* Greedy DBS of n_env images at once: the dataset loop of DBS.py:208 / DBS_1024_24.py:211 is independent per
 * image, so iteration i scores candidate orders[e][i] of every environment in one launch and keeps the
 * improving flips in one launch -- no speculation, every evaluation counts, and each image's decisions are
 * those of its own sequential loop.  orders / accepted / psnr_trace: [n_env][n] (psnr_trace nullable);
 * env_ids NULL = 0..n_env-1; resync_every (<= 0: never): re-propagate all listed environments every that
 * many CANDIDATES; n_accepted / final_psnr: [n_env] (nullable). */
int bh_dbs_run_batch(bh_ctx* ctx, int n_env, const int32_t* env_ids, const int64_t* orders, int64_t n,
                     int64_t resync_every, uint8_t* accepted, double* psnr_trace,
                     int64_t* n_accepted, double* final_psnr);

/* Reconstruction float [G][N][N].  candidate_action >= 0 adds the intensity
 * change of that (uncommitted) flip -- obs["recon_image"] of a rejected step
 * (env.py:176-181). */
int bh_get_recon(bh_ctx* ctx, int env, float* out, int on_host, int64_t candidate_action);
int bh_get_state(bh_ctx* ctx, int env, int8_t* out, int on_host);

/* obs["recon_image"] of the n environments of a step in one call (env.py:176-181: the
 * reference returns result_after.cpu().numpy() on every step; appendix B-2: a rejected step
 * shows the reconstruction WITH the rejected flip).  out is an observation block float
 * [E][G][N][N] indexed by ENV ID that the caller keeps between calls:
 *   out_kind BH_OBS_PINNED_HOST  host block from bh_host_alloc; the kernel writes it over PCIe
 *            BH_OBS_DEVICE       device block of the caller (zero-copy observation on the GPU)
 *            BH_OBS_CONTEXT      device block `buffer` owned by the context (bh_recon_device_block)
 * The block is kept current plane by plane: a step changes one colour plane per environment, so
 * only that plane is rewritten, plus planes that went stale since this block was last written
 * (a rejected flip it still shows, flips kept while another block was current, a reset).  The
 * context tracks staleness for up to 4 blocks; `buffer` in [0,4) names the block, a caller that
 * alternates two blocks (double buffering) passes 0,1,0,1...
 * d_results: DEVICE records of the step that was just enqueued (NULL: the records of the last
 * bh_step_batch / bh_vec_step).  env_ids: host, NULL = 0..n-1.
 * flags: BH_OBS_COMMITTED_ONLY ignore the records (committed reconstruction, e.g. after a reset),
 *        BH_OBS_FULL rewrite every plane, BH_OBS_SYNC return when the block is written. */
#define BH_OBS_DEVICE       0
#define BH_OBS_PINNED_HOST  1
#define BH_OBS_CONTEXT      2
#define BH_OBS_COMMITTED_ONLY 1
#define BH_OBS_FULL           2
#define BH_OBS_SYNC           4
int bh_recon_batch(bh_ctx* ctx, int n, const int32_t* env_ids, const bh_result* d_results,
                   float* out, int out_kind, int buffer, int flags);
/* Device observation block `buffer` of the context, float [E][G][N][N] (allocated on first use). */
void* bh_recon_device_block(bh_ctx* ctx, int buffer);
/* Colour planes (N*N floats each) written by bh_recon_batch since the context was created:
 * the bytes of the observation path, for accounting.  Synchronises the stream. */
int64_t bh_recon_planes_written(bh_ctx* ctx);
/* Wait for everything enqueued on the context's stream. */
int bh_stream_sync(bh_ctx* ctx);
/* Field of one frame, complex64 [N][N] as interleaved floats (test hook). */
int bh_get_field(bh_ctx* ctx, int env, int frame, float* out, int on_host);

/* Page-locked host memory for observation buffers (bh_get_recon then runs at PCIe speed). */
void* bh_host_alloc(size_t bytes);
int bh_host_free(void* p);

/* Device pointers of the resident arrays (for zero-copy views): which =
 * 0 U, 1 I, 2 T, 3 state, 4 sums, 5 h, 6 H. */
void* bh_device_ptr(bh_ctx* ctx, int which);

/* Stand-alone operator: tt.simulate(field, z) for C frames of N x N
 * (env.py:127).  in: float [C][N][N] (is_complex = 0) or interleaved complex;
 * out: interleaved complex64 [C][N][N].  Pointers are device pointers when
 * on_host = 0. */
int bh_simulate(int device, void* cuda_stream, const float* in, int is_complex, int C, int N,
                double wl, double dx, double z, int pad, int method, float* out, int on_host);

/* Timing hooks for bench.py: launch the delta-eval kernel (or the propagation
 * passes) `reps` times on the context stream between two CUDA events and return
 * the average milliseconds per launch.  Launch i scores the n tasks
 * d_actions[(i % n_sets) * n ...] of environments d_env_ids[0..n), so that
 * consecutive launches stream different frames (working set >> L2). */
int bh_time_eval(bh_ctx* ctx, int n, const int32_t* d_env_ids, const int64_t* d_actions,
                 int n_sets, int reps, float* ms_per_launch);
/* Open-loop rollout: `steps` sequential flips per environment with actions known in advance (device
 * resident), in ONE persistent cooperative launch (k_rollout_t: per-environment barriers instead of two
 * launches per step).  Replaces `steps` calls of the reference's step()  (env.py:154-260 reward path) /
 * `steps` iterations of the DBS loop over several images (DBS_1024_24.py:313-422) when the actions do not
 * depend on the observations.  Action of (step t, slot e) = d_actions[t * act_step_stride + e * act_env_stride]
 * (< 0: idle), result likewise in d_results (may be NULL).  The n_env slots must name DISTINCT environments
 * (d_env_ids NULL: slot e = environment e).  Decisions, results and the final device state are
 * bit-identical to `steps` calls of bh_step_batch_device.  Image sizes that are not row regular (N does not
 * divide 1024) and RULE_NEVER run the two-kernel chain step by step (env strides must be 1 then).
 * Asynchronous on the context stream; bh_rollout_status synchronises and reports an aborted barrier. */
int bh_rollout_device(bh_ctx* ctx, int n_env, const int32_t* d_env_ids, const int64_t* d_actions,
                      int64_t act_step_stride, int64_t act_env_stride, int steps, int rule,
                      bh_result* d_results, int64_t res_step_stride, int64_t res_env_stride);
int bh_rollout_status(bh_ctx* ctx);
/* The step chain as bh_step_batch_device launches it (k_eval -> k_commit) for `reps` steps;
 * with_commit = 0 times the evaluations alone under the same rule. */
int bh_time_step(bh_ctx* ctx, int n, const int32_t* d_env_ids, const int64_t* d_actions,
                 int n_sets, int reps, int rule, int with_commit, float* ms_per_step);
/* k_commit alone with n kept flips per launch (24 N^2 algorithmic bytes each); re-propagates
 * every environment afterwards. */
int bh_time_commit(bh_ctx* ctx, int n, const int32_t* d_env_ids, const int64_t* d_actions,
                   int n_sets, int reps, float* ms_per_launch);
int bh_time_propagate(bh_ctx* ctx, int env, int reps, float* ms_per_launch);
/* Per-pass split of one propagation (summed over the colour groups), ms4 =
 * {row FFT, column FFT * H * inverse column FFT, inverse row FFT, intensity + loss sums}. */
int bh_time_propagate_passes(bh_ctx* ctx, int env, int reps, float* ms4);
/* Kernels launched by this context since creation (for "gpu_launches"). */
int64_t bh_launch_count(const bh_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* BHOLO_H */
