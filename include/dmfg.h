/*
 * dmfg.h -- C ABI of the B200 (sm_100a) population-dynamics hot path.
 *
 * The reference (011235813/discrete_mean_field_game) has no FFI / plugin
 * boundary: its only interface is the Python method surface of the solver
 * classes.  This header therefore DEFINES the boundary; each entry point cites
 * the reference method(s) it replaces.  A maintainer binds it with ctypes
 * (INTEGRATION.md shows the stub) -- no torch types cross this interface.
 *
 * Conventions
 *   - every function returns an int status: DMFG_OK (0) or a negative error;
 *     nothing throws across the ABI; dmfg_last_error() gives the message of
 *     the calling thread's last failure.
 *   - pointers are DEVICE pointers owned by the caller unless the name ends
 *     in _host; `stream` is a cudaStream_t passed as void* (NULL = default).
 *     All work is enqueued on that stream and nothing synchronises unless
 *     stated; there is no hidden global state, calls are re-entrant per stream.
 *   - "streams" (states, actions, noise, per-step scalars) have the element
 *     type selected by `dtype`; PARAMETERS and REDUCED SUMS (theta, w, the
 *     gradient accumulators) are always double.
 *   - bulk layouts are TIME-MAJOR so that one CTA's populations are contiguous
 *     at every step:  states [T+1][B][d], actions / noise [T][B][d][d],
 *     per-step scalars [T][B].
 *   - F = d(d+1)/2 + d + 1 critic features, ordered as the reference's code
 *     builds them (mfg_ac2.py:333-344): pi_i*pi_j for i<=j row-major, pi, 1.
 */
#ifndef DMFG_H
#define DMFG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMFG_VERSION 1

/* status codes */
#define DMFG_OK               0
#define DMFG_ERR_INVALID     -1   /* bad argument (NULL, size, enum)          */
#define DMFG_ERR_UNSUPPORTED -2   /* valid request this build cannot serve    */
#define DMFG_ERR_CUDA        -3   /* CUDA runtime error (see dmfg_last_error) */
#define DMFG_ERR_WORKSPACE   -4   /* workspace missing or too small           */

/* dtype of stream buffers */
#define DMFG_F32 0
#define DMFG_F64 1

/* reward functor */
#define DMFG_REWARD_NONE      0   /* rollout only (generate_trajectory)                     */
#define DMFG_REWARD_AC2       1   /* mfg_ac2.py:257-287  sum_i pi_i sum_j P_ij^2 (pi_j-pi_i) */
#define DMFG_REWARD_SYNTHETIC 2   /* mfg_synthetic.py:249-265  -1/2 sum_i pi_i |P_i|^2       */

/* factor on V(pi') in the TD error */
#define DMFG_DISCOUNT_STEP       0   /* gamma           (mfg_ac2.py:505) */
#define DMFG_DISCOUNT_CUMULATIVE 1   /* gamma^t running (ac_irl.py:691)  */

/* noise source for the Gamma variates of sample_action */
#define DMFG_NOISE_INJECTED 0   /* caller supplies y[T][B][d][d] (parity with the reference's draws) */
#define DMFG_NOISE_PHILOX   1   /* in-kernel Philox4x32-10 + Marsaglia-Tsang, keyed by (seed, population id) */
#define DMFG_NOISE_ACTIONS  2   /* noise_y holds the transition matrices P themselves (no normalisation):
                                   evaluates reward / gradient / pi' of GIVEN (state, action) pairs, e.g.
                                   calc_reward(P, pi) and calc_gradient_vectorized(P, pi) on caller data */

/* kernel selection (testing aid) */
#define DMFG_VARIANT_AUTO    0
#define DMFG_VARIANT_GENERIC 1   /* warp per population, any d <= DMFG_MAX_D         */
#define DMFG_VARIANT_FAST    2   /* half-warp per population, compile-time d (4/15/16), float or double streams */
#define DMFG_VARIANT_V2      3   /* throughput kernel: float streams, d = 15/16/20/21/32 (what AUTO picks when it applies) */

#define DMFG_MAX_D 256

/* ---- library ---------------------------------------------------------- */
int         dmfg_version(void);
const char* dmfg_last_error(void);
/* kernels launched through this library by the calling process so far (every launch site counts itself);
 * bench.py reports the difference over its timed region as `gpu_launches` */
uint64_t    dmfg_kernel_launches(void);
/* number of critic features for d topics (mfg_ac2.py:175) */
int64_t     dmfg_num_features(int32_t d);
/* length of the reduced accumulator written by dmfg_rollout / dmfg_td_accumulate: 2 + F doubles
 *   acc[0] = sum_{t,b} delta*g   acc[1..F] = sum_{t,b} delta*phi(pi_t)   acc[1+F] = sum_{t,b} r   */
int64_t     dmfg_acc_len(int32_t d);

/* ---- a1..a7, a9: batched rollout with frozen parameters ---------------- *
 * Replaces, for B populations at once and T transitions each:
 *   sample_action            mfg_ac2.py:211-254, ac_irl.py:509-548
 *   pi' = P^T pi             mfg_ac2.py:497
 *   calc_reward              mfg_ac2.py:257-287 / mfg_synthetic.py:249-265
 *   calc_features, TD error  mfg_ac2.py:325-344, 505
 *   calc_gradient_vectorized mfg_ac2.py:347-381, ac_irl.py:573-624
 *   generate_trajectory(ies) mfg_ac2.py:566-592, ac_irl.py:735-767
 * Every output pointer is optional (NULL = not produced).                  */
typedef struct dmfg_rollout_args {
    uint32_t struct_size;       /* sizeof(dmfg_rollout_args) */
    int32_t  dtype;             /* DMFG_F32 | DMFG_F64 */
    int32_t  d;                 /* topics */
    int32_t  T;                 /* transitions per episode (reference: 15) */
    int64_t  B;                 /* populations in this call */
    int64_t  pop_offset;        /* global id of population 0 (Philox subsequence = pop_offset + b) */

    double   theta;             /* policy parameter, used when theta_dev == NULL */
    const double* theta_dev;    /* optional device scalar overriding `theta` */
    double   shift;
    double   alpha_scale;
    double   gamma;
    int32_t  reward_kind;       /* DMFG_REWARD_* */
    int32_t  discount_kind;     /* DMFG_DISCOUNT_* */

    int32_t  noise_kind;        /* DMFG_NOISE_* */
    int32_t  variant;           /* DMFG_VARIANT_* */
    const void* noise_y;        /* [T][B][d][d] Gamma(alpha*alpha_scale,1) variates (INJECTED) */
    uint64_t seed;              /* PHILOX key */
    uint64_t step_offset;       /* PHILOX: added to t so that successive calls do not reuse draws */

    const void*   pi0;          /* [B][d] start states */
    const double* w;            /* [F] critic weights; NULL = no TD error / accumulators */
    const void*   rewards_in;   /* optional [T][B]: externally supplied rewards (overrides reward_kind in delta) */

    void* states;               /* [T+1][B][d] */
    void* actions;              /* [T][B][d][d] */
    void* alpha;                /* [T][B][d][d]  mat_alpha       (mfg_ac2.py:228) */
    void* alpha_deriv;          /* [T][B][d][d]  mat_alpha_deriv (mfg_ac2.py:232-234) */
    void* rewards;              /* [T][B] */
    void* deltas;               /* [T][B] TD errors (needs w) */
    void* grads;                /* [T][B] d log F / d theta */
    void* pi_final;             /* [B][d] */
    double* acc;                /* [2+F] reduced sums, OVERWRITTEN (needs w and workspace) */

    void*    workspace;         /* scratch: per-CTA partial sums (deterministic two-stage reduction) and,
                                   for the generic variant with w != NULL, unrequested intermediates */
    uint64_t workspace_bytes;   /* >= dmfg_rollout_workspace_bytes(args) */
    const uint64_t* step_offset_dev;  /* optional device scalar ADDED to step_offset: lets a captured CUDA graph of
                                         launches be replayed with a different Philox position every time */
} dmfg_rollout_args;

/* exact scratch need of this call (0 when none); looks at d, T, B, dtype, variant, w, acc and
 * which outputs are requested -- never at buffer contents */
uint64_t dmfg_rollout_workspace_bytes(const dmfg_rollout_args* args);
int dmfg_rollout(const dmfg_rollout_args* args, void* stream);

/* ---- a4, a5: TD errors + accumulators from recorded trajectories -------- *
 * delta_t = r_t + g_t V(pi_{t+1}) - V(pi_t) with external rewards (the IRL
 * reward net, ac_irl.py:683-691); sums delta*g and delta*phi like dmfg_rollout. */
typedef struct dmfg_td_args {
    uint32_t struct_size;
    int32_t  dtype;
    int32_t  d, T;
    int64_t  B;
    double   gamma;
    int32_t  discount_kind;
    int32_t  reserved;
    const void*   states;       /* [T+1][B][d] */
    const void*   rewards;      /* [T][B] */
    const void*   grads;        /* [T][B] */
    const double* w;            /* [F] */
    void*    deltas;            /* [T][B] optional */
    double*  acc;               /* [2+F] optional */
    void*    workspace;
    uint64_t workspace_bytes;   /* >= dmfg_td_workspace_bytes(args) */
} dmfg_td_args;
uint64_t dmfg_td_workspace_bytes(const dmfg_td_args* args);
int dmfg_td_accumulate(const dmfg_td_args* args, void* stream);

/* ---- a4: critic features and values -------------------------------------- *
 * calc_features / calc_value (mfg_ac2.py:290-344) for N states [N][d]:
 * features [N][F] and/or values [N] = phi(pi) . w (either output may be NULL). */
int dmfg_critic_eval(int32_t dtype, int32_t d, int64_t N, const void* states, const double* w,
                     void* features, void* values, void* stream);

/* ---- consumer of a9: evaluation metrics of generated trajectories ---------- *
 * actor_critic.evaluate / JSD (mfg_ac2.py:546-563, 595-670; ac_irl.py:1495-1570): per (trajectory b, hour h)
 * the L1 distance and the Jensen-Shannon divergence between the generated and the empirical distribution
 * (zeros -> 1e-100, M = (P+Q)/2 from the unnormalised inputs, each entropy() argument normalised).
 * Element (b,h,j) of X is X[b*stride_b + h*stride_h + j]: a time-major rollout record [H][B][d] has
 * (stride_b, stride_h) = (d, B*d).  l1 / jsd: [B][H] doubles, either may be NULL.                     */
int dmfg_traj_metrics(int32_t dtype, int32_t d, int64_t B, int32_t H, const void* generated, int64_t gen_stride_b,
                      int64_t gen_stride_h, const void* empirical, int64_t emp_stride_b, int64_t emp_stride_h,
                      double* l1, double* jsd, void* stream);

/* ---- consumer of a9 (next row f4): analytic check against the MFG backward equation ---------- *
 * mfg_synthetic.actor_critic.evaluate_synthetic / evaluate_synthetic_JSD (mfg_synthetic.py:726-899) for B
 * recorded trajectories at once.  actions: the time-major record [T][B][d][d] dmfg_rollout writes.  Per
 * trajectory: V^T = 0, V^n = r(P^n) + P^n V^{n+1} with r_i = -1/2 ||P_i||^2; A^n_ij = V^n_j - V^n_i (i != j),
 * A^n_ii = 1 - (sum_j V^n_j - d V^n_i);  l1[b][n] = sum_ij |P^n_ij - A^n_ij|,  jsd[b][n] = sum_i JSD(P^n_i, A^n_i)
 * (entries <= 0 -> 1e-100, mfg_synthetic.py:529-546).  l1 / jsd: [B][T] doubles, either may be NULL.            */
int dmfg_synthetic_check(int32_t dtype, int32_t d, int64_t B, int32_t T, const void* actions, double* l1, double* jsd,
                         void* stream);

/* ---- a5, a7: apply one actor-critic update on device --------------------- *
 * theta += lr_actor_eff * scale * acc[0];  w += lr_critic_eff * scale * acc[1..F]
 * (mfg_ac2.py:511-522).  lr_*_eff are the already-decayed step sizes; scale is
 * 1/B for the batch-mean update.  theta_dev and w are updated in place.     */
int dmfg_ac_apply_update(int32_t d, double* theta_dev, double* w, const double* acc,
                         double lr_critic_eff, double lr_actor_eff, double scale, void* stream);
/* the same with the two step sizes in device memory, lr_dev[0] = lr_critic_eff, lr_dev[1] = lr_actor_eff (a launch whose
 * arguments do not change between CUDA-graph replays) */
int dmfg_ac_apply_update_dev(int32_t d, double* theta_dev, double* w, const double* acc, const double* lr_dev,
                             double scale, void* stream);

/* ---- a1..a7 for update="per_step": ONE fused launch per transition ------------ *
 * The synchronous batch-mean step of B populations sharing (theta, w) -- what mfg_ac2.py:497-522 does for its one
 * population, reduced over the batch: sample P, pi' = P^T pi, reward, TD error, sum delta*g and sum delta*phi, THEN
 * theta += lr_actor_eff * scale * sum delta*g and w += lr_critic_eff * scale * sum delta*phi, all in one kernel (the
 * last CTA to finish reduces the per-CTA partial sums in CTA order and applies the update).  `args` as for dmfg_rollout
 * with T = 1, float streams, d in {15, 16, 21}; args->theta_dev / args->w are ignored in favour of the in/out pointers
 * below; outputs: args->pi_final (the next states) and args->acc (optional).  lr_dev (optional, device) = {lr_critic_eff,
 * lr_actor_eff} overrides the by-value step sizes (CUDA-graph replays).  Replaces rollout(T=1) -> reduce -> apply_update. */
uint64_t dmfg_ac_step_workspace_bytes(const dmfg_rollout_args* args);
int dmfg_ac_step(const dmfg_rollout_args* args, double* theta_dev, double* w, double lr_critic_eff, double lr_actor_eff,
                 double scale, const double* lr_dev, void* stream);

/* ---- a8: independent serial learners (exact reference semantics) -------- *
 * L learners, each with its OWN (theta, w) and per-step online updates, run E
 * episodes of T transitions: mfg_ac2.actor_critic.train (mfg_ac2.py:448-539),
 * AC_IRL.train with a closed-form reward (ac_irl.py:634-732) and the
 * (shift, theta0) sweep of mfg_synthetic.py:902-925.  L = 1 is config 1.    */
#define DMFG_LEARNERS_AUTO   0
#define DMFG_LEARNERS_GROUPS 1   /* 16 / 32 lanes per learner: throughput form (thousands of learners) */
#define DMFG_LEARNERS_CTA    2   /* a CTA per learner (128 threads; 352 at d = 21): latency form (ONE learner = the reference's own run) */
typedef struct dmfg_learners_args {
    uint32_t struct_size;
    int32_t  dtype;
    int32_t  d, T;
    int64_t  L;                 /* learners */
    int64_t  learner_offset;    /* global id of learner 0 (Philox subsequence) */
    int32_t  E;                 /* episodes in this call */
    int32_t  episode0;          /* index of the first episode: 0 (mfg_ac2.py:460) or 1 (ac_irl.py:649), or a resume point */

    double*  theta;             /* [L] in/out */
    double*  w;                 /* [L][F] in/out */
    const double* shift;        /* [L] or NULL -> shift_scalar */
    const double* alpha_scale;  /* [L] or NULL -> alpha_scale_scalar */
    double   shift_scalar;
    double   alpha_scale_scalar;
    double   gamma;
    double   lr_critic;
    double   lr_actor;
    int32_t  constant_lr;       /* 1 = constant step sizes (mfg_ac2.py:511,519) */
    int32_t  reward_kind;
    int32_t  discount_kind;
    int32_t  noise_kind;

    const void*    mat_pi0;     /* [S][d] start-state table (init_pi0, mfg_ac2.py:179-208) */
    int32_t        S;
    int32_t        layout;      /* DMFG_LEARNERS_*: 0 = auto (CTA per learner for a few learners, float streams, d = 15/16/21) */
    const int32_t* start_rows;  /* [L][E] injected start rows, or NULL -> Philox randint */
    const void*    noise_y;     /* [L][E][T][d][d] (INJECTED) */
    uint64_t       seed;
    int64_t        noise_episode_offset; /* PHILOX: stream position = episode + this, so that repeated train()
                                            calls (which restart the step-size schedule at episode0) never
                                            reuse draws */

    double*  theta_trace;       /* [L][E][T] theta after each step, optional */
    double*  delta_trace;       /* [L][E][T] optional */
    double*  total_reward;      /* [L][E] optional */
    void*    pi_final;          /* [L][d] optional: last state of the last episode */
} dmfg_learners_args;
int dmfg_ac_learners(const dmfg_learners_args* args, void* stream);

/* ---- a8 with the reward network in the loop: AC_IRL.train as ONE kernel ------ *
 * AC_IRL.train (ac_irl.py:634-732): the serial actor-critic learner whose reward is r_net(pi_t, P_t), queried through
 * sess.run for every transition upstream (ac_irl.py:683).  `args` as for dmfg_ac_learners (reward_kind is ignored, float
 * streams, d <= 16, one CTA per learner; use episode0 = 1 and DMFG_DISCOUNT_CUMULATIVE for the reference's semantics);
 * `net` gives the reward net shared by all learners.  Dropout: DMFG_DROPOUT_NONE or DMFG_DROPOUT_PHILOX -- transition
 * t of episode e of learner l draws the masks of sample id sample_offset + (l*E + e)*T + t, exactly as
 * dmfg_rnet_forward would for that id (the reference's dropout is active whenever the net is evaluated). */
typedef struct dmfg_irl_net_args {
    uint32_t struct_size;
    int32_t  n_fc3, n_fc4;
    int32_t  dropout;
    float    keep_prob;
    int32_t  reserved;
    const float* params;          /* [dmfg_rnet_param_count(d, n_fc3, n_fc4)] */
    uint64_t seed, sample_offset;
    float*   reward_trace;        /* optional [L][E][T] */
} dmfg_irl_net_args;
int dmfg_irl_learners(const dmfg_learners_args* args, const dmfg_irl_net_args* net, void* stream);

/* ---- host-buffer convenience wrapper (the end-to-end path) --------------- *
 * Same as dmfg_rollout but pi0 / noise_y / w and every output are HOST
 * pointers; the call allocates device buffers, copies in, runs, copies out
 * and synchronises `stream` before returning.  workspace is ignored.       */
int dmfg_rollout_host(const dmfg_rollout_args* args_with_host_pointers, void* stream);

/* ---- a10: reward network r_net(state, action) ----------------------------- *
 * networks.py:13-157 (r_net, r_net_dropout, r_net_l1l2, r_net_dropout_l1l2) as
 * instantiated by AC_IRL.create_network (ac_irl.py:232-267: f1=1, k1=5, f2=2,
 * k2=3): conv5x5+ReLU -> conv3x3(2ch)+ReLU -> NHWC flatten -> fc3+ReLU [dropout]
 * -> concat state -> fc4+ReLU [dropout] -> out+tanh.   float32 like the TF graph
 * (ac_irl.py:239-246).  Parameters are ONE flat vector in TF variable order:
 *   conv1/weights[5,5,1,1] conv1/biases[1] conv2/weights[3,3,1,2] conv2/biases[2]
 *   fc3/weights[2d^2,n3] fc3/biases[n3] fc4/weights[n3+d,n4] fc4/biases[n4]
 *   out/weights[n4,1] out/biases[1]
 * Transitions are rows of states [N][d] / actions [N][d][d]; a time-major rollout
 * record [T][B].. is such an array with N = T*B.  n3, n4 <= 8.  Forward: d <= 32 (16 lanes per transition up to d = 16,
 * 32 above); backward: d <= 16 and d = 20, 21 (the 20 x 20 action files of ac_irl.py:164-200 and mfg_ac2.py:25's
 * default); dmfg_rnet_backward_gen: d <= 16. */
#define DMFG_DROPOUT_NONE   0   /* reg = 'none' | 'l1l2'                                       */
#define DMFG_DROPOUT_MASKS  1   /* caller supplies 0/1 keep masks (parity)                     */
#define DMFG_DROPOUT_PHILOX 2   /* in-kernel Philox keep masks keyed by (seed, sample_offset+n) */

int64_t dmfg_rnet_param_count(int32_t d, int32_t n_fc3, int32_t n_fc4);
/* offsets[10] of the ten tensors above inside the flat vector */
int dmfg_rnet_param_offsets(int32_t d, int32_t n_fc3, int32_t n_fc4, int64_t* offsets10);

typedef struct dmfg_rnet_args {
    uint32_t struct_size;
    int32_t  d, n_fc3, n_fc4;
    int64_t  N;                   /* transitions */
    const float* params;          /* [dmfg_rnet_param_count] */
    const float* states;          /* [N][d] */
    const float* actions;         /* [N][d][d] */
    int32_t  dropout;             /* DMFG_DROPOUT_* */
    float    keep_prob;           /* 0.4 in the reference (networks.py:70,75) */
    const uint8_t* mask3;         /* [N][n_fc3] keep masks (MASKS) */
    const uint8_t* mask4;         /* [N][n_fc4] */
    uint64_t seed;                /* PHILOX */
    uint64_t sample_offset;       /* PHILOX: global id of transition 0 */
    float*   rewards;             /* [N] out (forward: required; backward: optional) */
    /* backward only */
    const float* drewards;        /* [N] dL/dr */
    float*   grad;                /* [param_count] dL/dparams */
    int32_t  accumulate;          /* 0: grad is overwritten, 1: added to */
    int32_t  reserved;
    void*    workspace;           /* backward: per-CTA partial gradients */
    uint64_t workspace_bytes;     /* >= dmfg_rnet_workspace_bytes(args) */
    /* Optional gather from a pool of trajectories resident in device memory (gather_T = 0: off).  states / actions
     * are then the POOL -- slot s holds the gather_T transitions of one trajectory in rows [s*gather_T, (s+1)*gather_T)
     * -- and transition n of the batch is read from row gather_slots[n / gather_T] * gather_T + n % gather_T.
     * Everything else (rewards, drewards, masks, the Philox dropout offset) stays indexed by n, so a gathered batch
     * gives bit for bit what the same trajectories stacked into [N][d] arrays give.  N <= DMFG_MAX_GATHER * gather_T;
     * dmfg_rnet_backward_gen / dmfg_irl_reward_step: trajectory-major, gather_T = T.  The slots travel by value with
     * the launch: update_reward (ac_irl.py:804-846) resamples 5 + 5 of the same few dozen trajectories per update and
     * needs no host-to-device copy this way. */
    int32_t  gather_T;
    int32_t  reserved2;
    int32_t  gather_slots[32];    /* DMFG_MAX_GATHER */
} dmfg_rnet_args;
#define DMFG_MAX_GATHER 32
uint64_t dmfg_rnet_workspace_bytes(const dmfg_rnet_args* args);
/* sess.run(reward_gen / reward_demo) (ac_irl.py:683,882) */
int dmfg_rnet_forward(const dmfg_rnet_args* args, void* stream);
/* d(sum_n drewards[n] * r[n]) / dparams: the reward-net part of optimizer.minimize (ac_irl.py:417-418) */
int dmfg_rnet_backward(const dmfg_rnet_args* args, void* stream);

/* ---- a11: MaxEnt IRL loss and its derivative w.r.t. the rewards ----------- *
 * create_training_method (ac_irl.py:382-413):
 *   first  = -(1/num_demo_traj) * sum(r_demo)
 *   second = ln((1/M) * sum_j z_j * exp(sum_t r_gen[j,t]))     (z_j = 1 when log_z == NULL, as upstream :406)
 * r_gen[j,t] lives at r_gen[t*gen_t_stride + j*gen_j_stride] (time-major record: (M,1); the reference's
 * trajectory-major feed: (1,T)); d_gen uses the same indexing.
 * loss_out[4] (double, device) = {first+second, first, second, ln sum_j z_j exp(R_j)},
 * evaluated as a log-sum-exp (ln z_j can be ~ -6000, which is why upstream needs its `c` normaliser).  */
typedef struct dmfg_irl_loss_args {
    uint32_t struct_size;
    int32_t  T;
    int64_t  n_demo;              /* demo transitions */
    int64_t  M;                   /* generated trajectories */
    int64_t  gen_t_stride, gen_j_stride;
    double   num_demo_traj;       /* quirk: the demo sum is divided by TRAJECTORIES (ac_irl.py:390) */
    const float* r_demo;          /* [n_demo] */
    const float* r_gen;           /* [M*T] */
    const float* log_z;           /* [M] or NULL */
    float*   d_demo;              /* [n_demo] out, optional */
    float*   d_gen;               /* [M*T] out, optional */
    double*  loss_out;            /* [4] out */
    void*    workspace;
    uint64_t workspace_bytes;     /* >= dmfg_irl_loss_workspace_bytes(M) */
} dmfg_irl_loss_args;
uint64_t dmfg_irl_loss_workspace_bytes(int64_t M);
int dmfg_irl_loss_grad(const dmfg_irl_loss_args* args, void* stream);

/* ---- a11/a12: the generated half of one reward update in ONE pass ---------- *
 * Replaces, for z_j = 1 (upstream's ac_irl.py:406) and T <= 16, the chain
 *   dmfg_rnet_forward(gen) -> dmfg_irl_loss_grad -> dmfg_rnet_backward(gen, d_gen):
 * one launch walks the generated batch trajectory by trajectory (net->states / actions, N = M*T, transition (j,t)
 * at t*gen_t_stride + j*gen_j_stride), backpropagates with the unnormalised weight exp(R_j), R_j = sum_t r[j,t],
 * and the reduced gradient is scaled by 1 / sum_j exp(R_j) before it is written / added to net->grad.
 * net->drewards is ignored; net->rewards (optional) receives r_gen.  r_demo [n_demo] are the rewards of the
 * demonstration batch (dmfg_rnet_backward(demo) with rewards != NULL hands them back).
 * loss_out[4] (double, device) = {first+second, first, second, ln sum_j exp(R_j)} as in dmfg_irl_loss_grad. */
typedef struct dmfg_irl_gen_args {
    uint32_t struct_size;
    int32_t  T;
    int64_t  M;                   /* generated trajectories */
    int64_t  gen_t_stride, gen_j_stride;
    int64_t  n_demo;
    double   num_demo_traj;
    const float* r_demo;          /* [n_demo] */
    double*  loss_out;            /* [4] out */
    int32_t  local_sums;          /* 0: as above.  1 (data-parallel form): net->grad receives the UNNORMALISED gradient
                                     sum_j exp(R_j) dR_j/dparams of THIS rank's trajectories and loss_out[0..1] its local
                                     sums {sum_j exp(R_j), sum r_demo}; all-reduce [grad_demo, grad, sums] over the ranks
                                     and hand the result to dmfg_irl_dp_finalize -- the step then does not depend on how
                                     the trajectories were sharded */
    int32_t  reserved;
} dmfg_irl_gen_args;
int dmfg_rnet_backward_gen(const dmfg_rnet_args* net, const dmfg_irl_gen_args* gen, void* stream);
/* reduced [2n+4] doubles = {demonstration gradient for dL/dr = -1 per transition (n), unnormalised generated gradient
 * (n), Z = sum_j exp(R_j), sum r_demo, demonstration trajectories, generated trajectories}, each summed over all
 * ranks; grad[n] (float) = grad_demo / N_demo + grad_gen / Z -- the gradient of ac_irl.py:390-406 on the WHOLE batch --
 * and loss_out[4] as in dmfg_irl_loss_grad (optional). */
int dmfg_irl_dp_finalize(int64_t n, const double* reduced, float* grad, double* loss_out, void* stream);

/* ---- a11: one TF-style Adam step on the flat parameter vector -------------- *
 * tf.train.AdamOptimizer(lr).minimize (ac_irl.py:417-418): step counts from 1,
 * lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t), p -= lr_t*m/(sqrt(v)+eps).  With l1l2 != 0
 * the gradient of l1_l2_regularizer() (sum|w| + sum w^2/2, networks.py:69,74) on the fc3
 * and fc4 weight ranges of an r_net(d,n3,n4) vector is added; reg_loss_out (optional,
 * device double) receives the regulariser value BEFORE the step.  grad_scale multiplies
 * grad first (1/world for a data-parallel mean).                                  */
int dmfg_adam_tf(int64_t n, float* params, float* m, float* v, const float* grad, double grad_scale,
                 int64_t step, double lr, double beta1, double beta2, double eps,
                 int32_t l1l2, int32_t d, int32_t n_fc3, int32_t n_fc4, double* reg_loss_out, void* stream);

/* ---- a12: ONE reward update in one call ------------------------------------ *
 * AC_IRL.update_reward (ac_irl.py:804-846 -> :382-418) on one rank with z_j = 1: the chain
 *   dmfg_rnet_backward(demo)  ->  dmfg_rnet_backward_gen(gen, g)  ->  dmfg_adam_tf(...)
 * behind one entry point: one trip through the host binding instead of three (a 5 + 5 trajectory update is host-bound:
 * its two reward-net launches take 16 us each).  With a workspace of dmfg_irl_reward_step_workspace_bytes(demo) the two
 * reductions, the loss kernel and the Adam kernel of the chain run as ONE launch behind the two backward launches
 * (six launches -> three); with the smaller dmfg_rnet_workspace_bytes it is the chain's own launches in order.  Either
 * way parameters, moments and gradient are bit-identical to the chain; the fast form's first loss term is summed per
 * CTA and agrees with the chain's to the last bits of a double.
 * demo: N demonstration transitions, drewards = the constant -1/num_demo_traj, rewards = r_demo out (required),
 * grad = the gradient buffer of the whole step, workspace as for dmfg_rnet_backward.  gen: the generated
 * transitions (its grad / accumulate / workspace fields are ignored: the step's buffers are used).  g: as for
 * dmfg_rnet_backward_gen (r_demo / n_demo are taken from `demo`).  step: the Adam state; params must be the vector
 * both `demo` and `gen` point at (it is updated in place AFTER the two reward-net launches). */
typedef struct dmfg_irl_step_args {
    uint32_t struct_size;
    int32_t  l1l2;
    float*   params;              /* in/out */
    float*   m;
    float*   v;
    int64_t  step;                /* counts from 1 */
    double   lr, beta1, beta2, eps;
    double*  reg_loss_out;        /* optional (device double) */
} dmfg_irl_step_args;
uint64_t dmfg_irl_reward_step_workspace_bytes(const dmfg_rnet_args* demo);
int dmfg_irl_reward_step(const dmfg_rnet_args* demo, const dmfg_rnet_args* gen, const dmfg_irl_gen_args* g,
                         const dmfg_irl_step_args* step, void* stream);

/* ---- a13: Dirichlet log-density of recorded actions under K policies ------ *
 * calc_pdf_action / calc_z (ac_irl.py:270-379), in log space and without the `c`
 * normaliser: logq[n][k] = sum_i ln Dir(a_n[i,:]; max(alpha_{theta_k}(s_n)[i,:], 1+1e-6)).  */
int dmfg_dirichlet_logq(int32_t d, int64_t N, int32_t K, const float* states, const float* actions,
                        const double* thetas, double shift, double* logq, void* stream);
/* ln z_j = ln K - ln num_start - logsumexp_k sum_t logq[t*t_stride + j*j_stride][k]   (ac_irl.py:377-379) */
int dmfg_irl_log_z(int64_t M, int32_t T, int32_t K, int64_t t_stride, int64_t j_stride, const double* logq,
                   double num_start_samples, float* log_z, void* stream);

/* ---- testing aids: direct access to the device math ---------------------- */
/* The tensor-core (tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators) path of the fc3 weight gradient of
 * dmfg_rnet_backward in isolation: out[512][8] = sum_p h_p^T . z_p for h [passes][16][512], z [passes][16][8]
 * (device pointers; one CTA; the operand tiles, descriptors and the commit / mbarrier hand-off are the kernel's own). */
int dmfg_umma_selftest(const float* h, const float* z, int32_t passes, float* out, void* stream);
/* ONE kind::tf32 MMA out[128][32] (columns 0..15 = the product, the rest keep the pattern 1000 + lane + column / 100 stored
 * into TMEM beforehand; a_mn_major < 0: no MMA at all) = A[128][8] . B[8][16] with each operand laid out in shared memory by the canonical
 * un-swizzled formula of the requested major-ness (0 = K-major, 1 = MN-major) and (LBO, SBO) in bytes: pins the matrix
 * descriptor semantics the reward-net kernels rely on. */
int dmfg_umma_probe(const float* A, const float* B, int32_t a_mn_major, uint32_t a_lbo, uint32_t a_sbo, int32_t b_mn_major,
                    uint32_t b_lbo, uint32_t b_sbo, float* out, void* stream, uint32_t a_desc_lbo, uint32_t a_desc_sbo,
                    uint32_t b_desc_lbo, uint32_t b_desc_sbo);   /* *_desc_*: the values written into the descriptors */
/* Philox4x32-10 on the host (same code the kernels run): out[4] = philox(ctr[4], key[2]) */
void dmfg_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out);
/* The Gamma sampler of the rollout kernels (the replacement of np.random.gamma, mfg_ac2.py:242) runs the
 * same generator with dmfg_gamma_philox_rounds() = 7 rounds (Philox4x32-7, the smallest Crush-resistant
 * member of the family in Salmon et al. SC'11); this is that block function on the host. */
int32_t dmfg_gamma_philox_rounds(void);
void dmfg_philox4x32_gamma(const uint32_t* ctr, const uint32_t* key, uint32_t* out);
/* n Gamma(shape[i],1) variates (float, device pointers) drawn exactly as the rollout kernels
 * draw them: element i uses pair slot i/2 of population `pop` */
int dmfg_gamma_sample(const float* shape, int64_t n, uint64_t seed, uint64_t pop, float* out, void* stream);
/* out[i] = digamma(x[i]) with the device implementation of the given dtype */
int dmfg_digamma(int32_t dtype, const void* x, int64_t n, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMFG_H */
