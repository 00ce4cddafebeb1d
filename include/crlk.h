/*
 * crlk.h -- C ABI of libcrlk.so: the B200 (sm_100a) implementation of
 * crl_kino's kinodynamic-RRT extend path.
 *
 * The reference is pure Python and has no FFI; its "plugin API" is duck-typed
 * Python.  Each entry point below names the reference
 * interface it replaces (file:line in the reference checkout); the host-side
 * mirror of those interfaces lives in crl_kino_b200/ and binds this header
 * with ctypes (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (CRLK_E_*);
 *     crlk_last_error() returns a thread-local message for the last failure.
 *   - pointers documented "dev" are device pointers owned by the caller
 *     (e.g. PyTorch tensors); the library never frees or retains them past the
 *     call.  Pointers documented "host" are host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *     Launches are asynchronous; no entry point synchronises unless stated.
 *   - state = [x, y, theta, v, omega] float64[5]; obstacle = [left, bottom,
 *     right, top] float32[4] (rigid.py:146 stores float32).
 *   - handles are opaque, bound to the CUDA device current at creation and
 *     are safe to use from one thread at a time.
 */
#ifndef CRLK_H
#define CRLK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRLK_OK 0
#define CRLK_E_INVALID (-1)   /* bad argument */
#define CRLK_E_CUDA (-2)      /* CUDA runtime error (message has the detail) */
#define CRLK_E_NODEVICE (-3)  /* no CUDA device: there is NO CPU fallback */
#define CRLK_E_CAPACITY (-4)  /* a caller-provided capacity is too small */

#define CRLK_ABI_VERSION 2

/* status codes written by the extend kernels (rrt.py:121-128) */
#define CRLK_EXT_TIMEOUT 0    /* n_max_extend steps executed */
#define CRLK_EXT_COLLISION 1  /* valid_state_check failed; open segment dropped */
#define CRLK_EXT_REACHED 2    /* ||xy - target|| <= reach_radius */
#define CRLK_EXT_INVALID 3    /* empty DWA grid: start state outside the velocity limits (the reference raises) */

typedef struct CrlkEnv CrlkEnv;          /* DifferentialDriveEnv + obstacle set */
typedef struct CrlkPolicy CrlkPolicy;    /* DDPG actor MLP */
typedef struct CrlkEstimator CrlkEstimator; /* estimator + classifier CNN pair */

/* DifferentialDriveEnv.__init__ (env/differential_env.py:28-53). */
typedef struct {
    double max_v, min_v, max_w, max_acc_v, max_acc_w, base_dt;
    double env_bounds[4];   /* xmin xmax ymin ymax */
    float robot_rect[4];    /* RectRobot rect, float32 values (differential_env.py:41) */
    double robot_radius;    /* > 0: CircleRobot(radius) footprint (utils/rigid.py:331-347) replaces the rectangle */
} CrlkEnvParams;

/* DWA.__init__ / set_dwa (policy/dwa.py:13-33).  use_clearance = 0 is the
 * shipped behaviour (dwa.py:64-69 is commented out: the obstacle term is the
 * constant 1/100); 1 evaluates those lines. */
typedef struct {
    double v_res, w_res, dt, predict_time;
    double to_goal_cost_gain, ob_gain, speed_cost_gain;
    int32_t skip;
    int32_t use_clearance;
} CrlkDwaParams;

/* RRT.steer constants (planner/rrt.py:99-112): n_max_extend = round(10/0.2),
 * n_tree = round(5/0.2), step_dt = 0.2, reach_radius = 1.0. */
typedef struct {
    int32_t n_max_extend;
    int32_t n_tree;
    double step_dt;
    double reach_radius;
} CrlkExtendParams;

const char *crlk_last_error(void);
int crlk_abi_version(void);
/* number of CUDA devices visible (0 => every compute entry point fails). */
int crlk_device_count(void);

/* ---- environment: DifferentialDriveEnv (env/differential_env.py:27-169) ---- */

/* set_obs (differential_env.py:120-123) + ctor.  rects_host: M x 4 float32. */
int crlk_env_create(const CrlkEnvParams *params, const float *rects_host, int32_t M, CrlkEnv **out);
int crlk_env_destroy(CrlkEnv *env);
int crlk_env_set_obs(CrlkEnv *env, const float *rects_host, int32_t M);
int crlk_env_num_obs(const CrlkEnv *env);

/* valid_state_check (differential_env.py:146-151) and get_clearance (:126-132)
 * for B states (dev, B x 5 f64, row stride 5).
 *   valid          dev u8[B]  (nullable) 1 = collision free
 *   clearance      dev f64[B] (nullable) min(100, min_obs dis2obs), -1 if colliding
 *   near_boundary  dev u8[B]  (nullable) 1 if some obstacle's distance is within
 *                  1e-6 of the 0.05 collision threshold (rigid.py:311): parity
 *                  with the reference is not claimed for those states. */
int crlk_valid_state(const CrlkEnv *env, const double *states, int32_t B, uint8_t *valid,
                     double *clearance, uint8_t *near_boundary, void *stream);

/* valid_point_check (differential_env.py:135-144): points dev K x 2 f64. */
int crlk_valid_points(const CrlkEnv *env, const double *points, int32_t K, uint8_t *valid,
                      void *stream);

/* motion_velocity (differential_env.py:55-73): states_in/out dev B x 5, cmds dev B x 2. */
int crlk_motion_velocity(const CrlkEnv *env, const double *states_in, const double *cmds,
                         double duration, int32_t B, double *states_out, void *stream);
/* motion (differential_env.py:76-83): acceleration command. */
int crlk_motion(const CrlkEnv *env, const double *states_in, const double *acc, double duration,
                int32_t B, double *states_out, void *stream);
/* dynamic_window (differential_env.py:107-114): out dev B x 4 [v_lo v_hi w_lo w_hi]. */
int crlk_dynamic_window(const CrlkEnv *env, const double *states, double dt, int32_t B,
                        double *out, void *stream);

/* DifferentialDriveGym._obs (env/differential_gym.py:180-190): goal in body
 * frame, v, omega, 27 x 27 local occupancy map (:62-91).
 *   goals   dev B x goal_stride f64 (only [0], [1] are read)
 *   obs64   dev f64[B x 733] (nullable)   the reference's float64 observation
 *   obs32   dev f32[B x ld32] (nullable)  same values cast to float32 (net.py:142),
 *           ld32 >= 733 (736 for the tensor-core MLP); padding is zero-filled
 *   near_boundary dev u8[B] (nullable): some sample point within 1e-9 of an
 *           obstacle face. */
int crlk_local_obs(const CrlkEnv *env, const double *states, const double *goals,
                   int32_t goal_stride, int32_t B, double *obs64, float *obs32, int32_t ld32,
                   uint8_t *near_boundary, void *stream);

/* ---- steering: DWA.control (policy/dwa.py:45-86) ---- */

/* One dynamic-window control decision per (state, goal) pair.
 *   opt_v     dev f64[B x 2]
 *   opt_idx   dev i32[B]   v-major index of the first minimum (np.argmin)
 *   opt_cost  dev f64[B]   (nullable)
 *   opt_traj  dev f64[B x traj_rows x 5] (nullable); traj_rows from crlk_dwa_traj_rows
 *   n_rollouts dev i32[B]  (nullable) realised grid size nv * nw
 *   near_tie  dev u8[B]    (nullable) the decision hangs on a last bit the C library may round differently (below).
 * Costs are formed in the reference's own arithmetic: heading costs from a correctly rounded atan2 on
 * positions integrated with correctly rounded sin / cos (the C library calls behind np.arctan2 /
 * math.sin / math.cos on Python scalars, dwa.py:58, differential_env.py:98-99), each column summed in
 * numpy's pairwise order (np.sum, dwa.py:76-78), every entry divided by its column sum (correctly
 * rounded), then (gain0 * c0 + gain1 * c1) + gain2 * c2 with each operation rounded once
 * (dwa.py:79-81), first minimum (np.argmin, :82).  The result is the reference's bit for bit, except
 * where the C library itself misrounds (7.5e-4 of its atan2 results, by one ulp).  near_tie marks the
 * decisions such a bit could change: a runner-up with a different heading cost within 8 ulp of the
 * winner, or one with the same heading cost that overtakes the winner when the heading column's sum
 * moves by one ulp.  Identical rollouts (grid values beyond the window clamp to the same command) are
 * not near ties: np.argmin and this kernel both keep the first.
 * An empty grid (v or omega outside the limits by more than acc * dt; the reference raises on the
 * empty cost table) gives opt_idx = -1, n_rollouts = 0 and NaN opt_v / opt_cost / opt_traj. */
int crlk_dwa_control(const CrlkEnv *env, const CrlkDwaParams *dwa, const double *states,
                     const double *goals, int32_t goal_stride, int32_t B, double *opt_v,
                     int32_t *opt_idx, double *opt_cost, double *opt_traj, int32_t *n_rollouts,
                     uint8_t *near_tie, void *stream);
/* rows of DWA.calc_trajectory (dwa.py:35-43) for these parameters (7 for dt 0.2). */
int crlk_dwa_traj_rows(const CrlkEnv *env, const CrlkDwaParams *dwa);
/* upper bound of the rollout grid size for these parameters. */
int crlk_dwa_max_rollouts(const CrlkEnv *env, const CrlkDwaParams *dwa);

/* ---- nearest node: RRT.choose_parent (planner/rrt.py:156-161) ---- */

/* Q independent trees in SoA form.  Tree q occupies elements [q*stride,
 * q*stride + n_nodes[q]) of each array.
 *   x64, y64   dev f64   exact coordinates (rescoring); node j of tree q is element
 *                        (q*stride + j) * xy_step: xy_step = 1 for two SoA arrays, 5 when x64 / y64
 *                        point at columns 0 / 1 of a [Q x stride x 5] state table (no second copy
 *                        of the coordinates: 48 instead of 64 bytes per node with the f32 copies)
 *   x32, y32   dev f32   the same coordinates rounded to float32 (the scanned
 *                        copy: 8 bytes per node); nullable => scan f64 directly
 *   coord_bound upper bound of |coordinate| over the trees and targets (e.g. the
 *              env bounds); sizes the float32 screening band.  Ignored if x32 is NULL.
 *   targets    dev f64[Q x target_stride]
 *   idx        dev i32[Q]; dist dev f64[Q]; near_tie dev u8[Q] (nullable)
 *   scratch    dev bytes, crlk_nearest_scratch_bytes(Q, max_nodes) of them. */
int crlk_nearest(const double *x64, const double *y64, int32_t xy_step, const float *x32, const float *y32,
                 int64_t stride, const int32_t *n_nodes, int32_t max_nodes, double coord_bound,
                 const double *targets, int32_t target_stride, int32_t Q, int32_t *idx, double *dist, uint8_t *near_tie,
                 void *scratch, void *stream);
int64_t crlk_nearest_scratch_bytes(int32_t Q, int32_t max_nodes);

/* ---- extension: RRT.steer (planner/rrt.py:99-135), DWA steering ---- */

/* One extension per (from, target) pair, all control steps on the device.
 *   from_states, targets  dev f64[Q x 5]
 *   path       dev f64[Q x n_max_extend x 5]  executed states, step 1..n_steps
 *   n_steps    dev i32[Q]
 *   status     dev i32[Q]  CRLK_EXT_*
 *   node_step  dev i32[Q x (n_max_extend/n_tree + 1)]  1-based step of each emitted node
 *   n_nodes    dev i32[Q]
 *   ctrl_idx   dev i32[Q x n_max_extend] (nullable) chosen rollout index per step
 *   flags      dev u8[Q] (nullable) bit0 near-boundary collision test, bit1 near-tie argmin (see
 *              crlk_dwa_control), bit3 empty DWA grid (status CRLK_EXT_INVALID, nothing executed)
 *   active     dev u8[Q] (nullable) 0 => skip this pair (n_steps = 0, n_nodes = 0)
 *   stats      dev u64[8] (nullable) ACCUMULATED (+=): [0] extensions run, [1] control steps executed,
 *              [2] rollouts scored, [3] exact FP64 footprint tests, [4] obstacle rectangles that went
 *              through the FP32 screens of the validity tests, [5] sincos chain tasks, [6] velocity
 *              chains, [7] reserved -- the executed-work counters behind bench.py's roofline. */
int crlk_extend_dwa(const CrlkEnv *env, const CrlkDwaParams *dwa, const CrlkExtendParams *ext,
                    const double *from_states, const double *targets, int32_t Q, double *path,
                    int32_t *n_steps, int32_t *status, int32_t *node_step, int32_t *n_nodes,
                    int32_t *ctrl_idx, uint8_t *flags, const uint8_t *active, uint64_t *stats,
                    void *stream);

/* dst[q, :] = src[q * q_stride + idx[q] * row_len : +row_len] -- picks the steer-from
 * node of every query out of the per-query node tables (rrt.py:70). */
int crlk_gather_rows(const double *src, int64_t q_stride, int32_t row_len, const int32_t *idx,
                     int32_t Q, double *dst, void *stream);

/* ---- learned steering: policy_forward (policy/rl_policy.py:83-109) ---- */

/* Actor MLP (policy/net.py:114-154).  weights_host[i]: [dims[i+1] x dims[i]]
 * float32 row-major (torch nn.Linear layout), biases_host[i]: [dims[i+1]].
 * low/high: action range, float32[dims[n_layers]] (gym Box float32). */
int crlk_policy_create(int32_t n_layers, const int32_t *dims, const float *const *weights_host,
                       const float *const *biases_host, const float *low, const float *high,
                       CrlkPolicy **out);
int crlk_policy_destroy(CrlkPolicy *p);
/* obs dev f32[B x ld_obs]; noise dev f32[B x 2] (nullable: no exploration
 * noise); act dev f32[B x 2] = clamp(tanh(mlp) + noise*eps) * scale + bias).
 * workspace dev bytes from crlk_policy_workspace_bytes(p, B).
 * Hidden layers whose widths are multiples of 128 run on the tensor cores (tcgen05, float32 operands as
 * bf16 planes, float32 accumulation: float32-level accuracy): below 49,152 rows ALL hidden layers run in one
 * dataflow launch (k_mlp_chain: tiles of all layers pulled from one list, row-tile completion flags
 * between layers), from 49,152 rows one launch per layer of the CTA-pair kernel (cta_group::2).  Other
 * widths use a float32 SIMT GEMM. */
int crlk_policy_forward(const CrlkPolicy *p, const float *obs, int32_t ld_obs, const float *noise,
                        float eps, int32_t B, float *act, void *workspace, void *stream);
int64_t crlk_policy_workspace_bytes(const CrlkPolicy *p, int32_t B);

/* RRT_RL.steer (planner/rrt_rl.py:22-63): like crlk_extend_dwa with the policy
 * as controller.  noise dev f32[Q x noise_rows x 2] (nullable): standard-normal draws,
 * one row per executed step (the reference draws torch.randn inside the actor,
 * net.py:147-150).  noise_cursor dev i32[Q] (nullable): first row to use per query;
 * advanced by the number of executed steps, so consecutive extensions of a query
 * walk its stream (the caller keeps noise_rows large enough).  Without a cursor
 * row k-1 is used at step k (noise_rows >= n_max_extend).
 * workspace: crlk_extend_rl_workspace_bytes, 256-byte aligned.
 * Every policy step runs on the compacted batch of queries that are still extending (row lists and
 * counts live on the device, no host synchronisation): one launch for the actor's hidden layers and one
 * for head + motion + validity + bookkeeping + next observation.  Batches of 49,152 queries or more (one
 * GEMM launch per layer) advance as two halves on `stream` and an internal second stream (joined before
 * the call's last kernel), so the GEMMs of one half overlap the other half's step kernel.  Results do
 * not depend on either.  The call is not re-entrant per device (serialised internally).
 * crlk_extend_rl_launches: kernels one call launches (for launch accounting; memsets not counted). */
int crlk_extend_rl(const CrlkEnv *env, const CrlkPolicy *p, const CrlkExtendParams *ext,
                   const double *from_states, const double *targets, int32_t Q, const float *noise,
                   int32_t noise_rows, int32_t *noise_cursor, float eps, double *path, int32_t *n_steps,
                   int32_t *status, int32_t *node_step, int32_t *n_nodes, float *actions, uint8_t *flags,
                   const uint8_t *active, void *workspace, void *stream);
int64_t crlk_extend_rl_workspace_bytes(const CrlkPolicy *p, int32_t Q);
int32_t crlk_extend_rl_launches(const CrlkPolicy *p, const CrlkExtendParams *ext, int32_t Q, int32_t with_noise_cursor);

/* ---- learned parent selection (estimator/network.py, rrt_rl_estimator.py:57-74) ---- */

/* params_host: the 10 tensors of one CNN state dict in the order
 * conv1.w conv1.b conv2.w conv2.b conv3.w conv3.b conv4.w conv4.b linear.w linear.b. */
int crlk_estimator_create(const float *const *est_params_host, const float *const *cls_params_host,
                          CrlkEstimator **out);
int crlk_estimator_destroy(CrlkEstimator *e);
/* obs dev f32[B x ld_obs] (the 733-vector of crlk_local_obs); est, prob dev f32[B]. */
int crlk_estimator_forward(const CrlkEstimator *e, const float *obs, int32_t ld_obs, int32_t B,
                           float *est, float *prob, void *stream);

/* RRT_RL_Estimator.choose_parent (planner/rrt_rl_estimator.py:26-79) for Q queries at once:
 * per node the Euclidean distance to the query's target; nodes with 1 <= d < 20 get their local
 * observation built and scored by both CNNs: cost = (est + 1) * 15 + 100 * (1 - round(prob));
 * other nodes keep the distance as cost; first minimum.  (0, 100) when no node passes the filter.
 *   states dev f64[Q x stride x 5] node tables, n_nodes dev i32[Q], targets dev f64[Q x 5]
 *   first_node dev i32[Q] (nullable): the candidate list of query q is nodes [first_node[q], n_nodes[q])
 *              (the `new_node_list` of rrt.py:78); idx stays an index into the whole tree
 *   idx dev i32[Q], cost dev f64[Q]; workspace: crlk_estimator_cp_workspace_bytes(Q, max_nodes). */
int crlk_estimator_choose_parent(const CrlkEnv *env, const CrlkEstimator *e, const double *states,
                                 int64_t stride, const int32_t *n_nodes, const int32_t *first_node,
                                 int32_t max_nodes, const double *targets, int32_t Q, int32_t *idx,
                                 double *cost, void *workspace, void *stream);
int64_t crlk_estimator_cp_workspace_bytes(int32_t Q, int32_t max_nodes);

/* ---- planner bookkeeping on the device: RRT.planning (planner/rrt.py:51-89) ---- */

/* Q trees in SoA form (device pointers; tree q owns elements [q*stride, (q+1)*stride)).
 * The optional path pool keeps, per node, the executed states between its parent
 * and itself (Node.path, rrt.py:120) for generate_final_course (rrt.py:137-144). */
typedef struct {
    double *x64, *y64;       /* [Q x stride] */
    float *x32, *y32;        /* [Q x stride] float32 copies scanned by crlk_nearest (nullable) */
    double *states;          /* [Q x stride x 5] */
    int32_t *parent;         /* [Q x stride] */
    int32_t *n_nodes;        /* [Q] */
    int64_t stride;          /* node capacity per tree */
    double *pool;            /* [Q x pool_cap x 5] (nullable: paths are not kept) */
    int32_t *seg_off, *seg_len;   /* [Q x stride] segment of each node inside its tree's pool */
    uint8_t *seg_flag;       /* [Q x stride] 1 = Node.path starts with the parent's state */
    int32_t *pool_used;      /* [Q] */
    int64_t pool_cap;
} CrlkForestView;

/* outputs of crlk_extend_dwa / crlk_extend_rl, as consumed by crlk_rrt_append */
typedef struct {
    const double *path;
    const int32_t *n_steps, *status, *node_step, *n_nodes;
} CrlkExtendView;

/* Sample acceptance of rrt.py:64-69 for one candidate per query:
 * active[q] = !done && iters < max_iter && valid && d_min <= dist < d_max;
 * accepted queries count one iteration (iters[q] += 1). */
int crlk_rrt_select(const uint8_t *valid, const double *dist, int32_t Q, int32_t max_iter, double d_min,
                    double d_max, const uint8_t *done, int32_t *iters, uint8_t *active, void *stream);

/* Append the nodes of one extension per active query (rrt.py:125-134), run the
 * goal test on the tree's last node (rrt.py:75, :81; radius goal_radius) and --
 * when primary != 0 -- choose the new node nearest to the goal for the retry
 * extension (rrt.py:78-80: cost < retry_radius).  first_has_start: 0 for the DWA
 * steer (the first segment's Node.path lacks the start state, rrt.py:102-104),
 * 1 for the RL steer (rrt_rl.py:27).  primary == 2: a primary pass whose retry choice is made by
 * the caller with its own choose_parent (RRT_RL_Estimator) and handed in through
 * crlk_rrt_retry_choice.  overflow (dev i32[1]) counts trees that ran out of node or pool capacity
 * (those queries are stopped and reported). */
int crlk_rrt_append(const CrlkForestView *forest, const CrlkExtendView *ext, const int32_t *parent_idx,
                    const uint8_t *active, const double *goals, int32_t Q, int32_t n_max_extend,
                    int32_t n_tree, int32_t first_has_start, int32_t primary, int32_t max_iter,
                    double goal_radius, double retry_radius, const int32_t *iters, uint8_t *done,
                    uint8_t *reached, uint8_t *retry_active, double *retry_from, int32_t *retry_parent,
                    int32_t *overflow, void *stream);

/* Retry decision after a crlk_rrt_append(primary = 2): n_before dev i32[Q] = n_nodes before that
 * append; (idx, cost) dev = choose_parent(new nodes, goal) (rrt.py:78); a query retries from node
 * idx iff it appended nodes, is not done and cost < retry_radius (rrt.py:79); otherwise it is
 * finished when its iterations are used up. */
int crlk_rrt_retry_choice(const CrlkForestView *forest, const int32_t *n_before, const int32_t *idx,
                          const double *cost, const uint8_t *active, int32_t Q, double retry_radius,
                          int32_t max_iter, const int32_t *iters, uint8_t *done, uint8_t *retry_active,
                          double *retry_from, int32_t *retry_parent, void *stream);

/* ---- gym side of the env: DifferentialDriveGym.step (env/differential_gym.py:136-178) ---- */

/* B environments over one obstacle set, one call per step:
 *   state <- motion_velocity(state, action, step_dt)   (:139; actions NULL: no motion, info/obs only)
 *   info  = _info()  (:147-166), columns in the dict order goal, goal_dis, heading, collision,
 *           clearance (= min(get_clearance, 1.0), -1 when colliding), v, |w|, step (= 1.0)
 *   reward = info @ reward_param  (:177-178; reward_param_host: 8 doubles, host memory)
 *   obs    = _obs()  (:180-190), float64 [B x 733] and/or float32 [B x ld32] as crlk_local_obs
 *   states dev f64[B x 5] IN/OUT; goals dev f64[B x goal_stride]; actions dev f64[B x 2] (v, w)
 *   info dev f64[B x 8], reward dev f64[B], near_boundary dev u8[B] (all nullable). */
int crlk_gym_step(const CrlkEnv *env, double *states, const double *goals, int32_t goal_stride,
                  const double *actions, double step_dt, const double *reward_param_host, int32_t B,
                  double *info, double *reward, double *obs64, float *obs32, int32_t ld32,
                  uint8_t *near_boundary, void *stream);

/* Test hook: the atan2 of the rollout scoring loop (dwa.py:58 `np.arctan2`), evaluated for n
 * argument pairs; y, x, out dev f64[n].  Correctly rounded (double-double evaluation with a rounding
 * test, csrc/crlk_crmath.h): the bits the C library returns, up to the library's own rare misroundings. */
int crlk_debug_atan2(const double *y, const double *x, int32_t n, double *out, void *stream);
/* Test hook: sin / cos as the heading chains and footprints evaluate them (math.sin / math.cos of
 * differential_env.py:98-99; correctly rounded for |th| <= 3.15); th, sn, cs dev f64[n]. */
int crlk_debug_sincos(const double *th, int32_t n, double *sn, double *cs, void *stream);
/* Test hook: fast[i] = the correctly rounded quotient a[i] / b[i] as the cost normalisation forms it
 * (one multiply + two fused corrections on y = 1 / b), exact[i] = a[i] / b[i] by the division operator. */
int crlk_debug_div(const double *a, const double *b, int32_t n, double *fast, double *exact, void *stream);

/* ---- whole-query planning on the device: RRT.planning (planner/rrt.py:51-89), DWA steering ---- */

typedef struct {
    int32_t max_iter;            /* RRT(max_iter), rrt.py:36: accepted samples per query */
    double d_min, d_max;         /* sample acceptance band 1.0 <= d < 20 (rrt.py:69) */
    double goal_radius;          /* 1.0 (rrt.py:75, :81) */
    double retry_radius;         /* 5.0 (rrt.py:79) */
    /* RRT.sample (rrt.py:146-154) on the device, used when no sample streams are passed:
     * draw i of query q is the goal with probability (goal_sample_rate + 1) / 101, else
     * state_lo + u * (state_hi - state_lo) per component, u from the counter-based generator
     * crlk_sample_u01(seed, q, i, component) below. */
    uint64_t seed;
    double state_lo[5], state_hi[5];   /* get_bounds()['state_bounds'] (differential_env.py:153-161) */
    int32_t goal_sample_rate;    /* 5 (rrt.py:36) */
} CrlkPlanParams;

/* The sample generator of crlk_plan_dwa, materialised: out dev f64[Q x S x 5] receives exactly the
 * draws the planner consumes when called with streams == NULL (same plan->seed / bounds / rate).
 * Counter based: u01(seed, q, i, c) = (mix(seed ^ q * 0x100000001B3 ^ i * 0x9E3779B97F4A7C15 ^
 * (c + 1) * 0xD6E8FEB86659FD93) >> 11) * 2^-53 with mix = the splitmix64 finaliser; c = 0..4 the state
 * components, c = 5 the goal-bias draw (goal iff floor(101 u) <= goal_sample_rate).
 * crl_kino_b200/workloads.py::device_sample_stream is the host mirror (bit-identical). */
int crlk_sample_stream(const CrlkPlanParams *plan, const double *goals, int32_t Q, int32_t S, double *out,
                       void *stream);

/* Q complete RRT.planning runs in ONE launch, one thread block per query from its
 * first sample to its last node: per consumed sample the validity check
 * (rrt.py:66), nearest node (:68), acceptance (:69), the extension (:72,
 * RRT.steer :99-135 with DWA.control as controller), node append with path
 * segments, goal test (:75) and the steer-to-goal retry (:78-83).  Replaces the
 * crlk_valid_state / crlk_nearest / crlk_rrt_select / crlk_extend_dwa /
 * crlk_rrt_append rounds for DWA steering; results are identical.
 *   forest   trees holding the start node (n_nodes[q] = 1) on entry, the final
 *            trees on return; with a path pool, pool_cap >= 2 * max_iter * n_max_extend
 *   goals    dev f64[Q x 5]
 *   streams  dev f64[Q x S x 5] pre-drawn samples per query (RRT.sample, rrt.py:146-154,
 *            draws from host RNG state; the stream is that sequence), consumed in order;
 *            NULL: up to S samples per query are drawn on the device (plan->seed, see CrlkPlanParams)
 *   iters    dev i32[Q] accepted samples (loop index of rrt.py:60 at exit)
 *   cursor   dev i32[Q] samples consumed (== S with iters < max_iter: stream exhausted)
 *   reached  dev u8[Q]  reach_exactly (rrt.py:76, :82)
 *   flags    dev u8[Q] (nullable) bit0 near-boundary collision test, bit1 near-tie argmin
 *            somewhere in the run, bit2 out of node / pool capacity
 *   overflow dev i32[1] += number of queries that ran out of capacity
 *   stats    dev u64[3] (nullable) += extensions, control steps, rollouts
 *   order    dev i32[Q] (nullable) a permutation of 0..Q-1: the order in which thread blocks pick
 *            queries up (scheduling only -- e.g. longest expected first; results do not depend on it) */
int crlk_plan_dwa(const CrlkEnv *env, const CrlkDwaParams *dwa, const CrlkExtendParams *ext,
                  const CrlkPlanParams *plan, const CrlkForestView *forest, const double *goals,
                  const double *streams, int32_t S, int32_t Q, int32_t *iters, int32_t *cursor,
                  uint8_t *reached, uint8_t *flags, int32_t *overflow, uint64_t *stats,
                  const int32_t *order, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CRLK_H */
