/*
 * crlk_oracle.c -- CPU restatement of crl_kino's kinodynamic-RRT extend path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (crl_kino_b200/)
 * may import, link or call this file.  Allowed callers: tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
 *
 * Parity status: PINNED against golden vectors produced by importing the
 * unmodified Python reference in the authoring container
 * (tests/golden/make_golden.py -> tests/golden/ *.npz; checked by
 * tests/test_oracle_golden.py).  The reference ships no tests of its own
 * (SURVEY.md section 4), so those generated vectors are the pin.
 *
 * Every function is a scalar, double-precision restatement in the reference's
 * own operation order, compiled with -ffp-contract=off so that no a*b+c is
 * fused unless the reference's numpy/BLAS call fuses it too (noted inline).
 * Citations are file:line into the reference checkout.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_TWO_PI (2.0 * M_PI)

typedef struct {
    double max_v, min_v, max_w, max_acc_v, max_acc_w, base_dt;
    double env_bounds[4]; /* xmin xmax ymin ymax */
    float robot_rect[4];  /* left bottom right top, float32 VALUES (rigid.py:263) */
    double robot_radius;  /* > 0: CircleRobot(radius) footprint (rigid.py:331-347) instead of the rectangle */
} OrcEnv;

typedef struct {
    double v_res, w_res, dt, predict_time;
    double to_goal_cost_gain, ob_gain, speed_cost_gain;
    int skip;
    int use_clearance; /* 0 = shipped behaviour (dwa.py:64-69 commented out) */
} OrcDwa;

/* Python builtin min/max on floats: min(a,b) -> b if b < a else a (first wins ties). */
static inline double pymin(double a, double b) { return (b < a) ? b : a; }
static inline double pymax(double a, double b) { return (b > a) ? b : a; }

/* np.linalg.norm of a 2-vector: sqrt(x.dot(x)); the BLAS ddot kernel numpy
 * links evaluates the n=2 tail as fma(b, b, a*a) (measured in the authoring
 * container; tests/golden/make_golden.py records the probe). */
static inline double norm2(double a, double b) { return sqrt(fma(b, b, a * a)); }

/* differential_env.py:20-24.  Python/numpy floor-mod for a positive divisor. */
double orc_normalize_angle(double a)
{
    double m = fmod(a, ORC_TWO_PI);
    if (m != 0.0) {
        if (m < 0.0) m += ORC_TWO_PI;
    } else {
        m = 0.0;
    }
    if (m > M_PI) m -= ORC_TWO_PI;
    return m;
}

/* differential_env.py:107-114.  out = [v_lo, v_hi, w_lo, w_hi]. */
void orc_dynamic_window(const OrcEnv *e, const double *state, double dt, double *out)
{
    out[1] = pymin(state[3] + e->max_acc_v * dt, e->max_v);
    out[0] = pymax(state[3] - e->max_acc_v * dt, e->min_v);
    out[3] = pymin(state[4] + e->max_acc_w * dt, e->max_w);
    out[2] = pymax(state[4] - e->max_acc_w * dt, -e->max_w);
}

/* differential_env.py:85-105 (mutates state in place, like the reference). */
static void motion_base(const OrcEnv *e, double *state, const double *input_u, double dt)
{
    double c[4];
    orc_dynamic_window(e, state, dt, c);
    double u0 = state[3] + input_u[0] * dt;
    double u1 = state[4] + input_u[1] * dt;
    u0 = pymax(c[0], u0);
    u0 = pymin(c[1], u0);
    u1 = pymax(c[2], u1);
    u1 = pymin(c[3], u1);
    state[2] = state[2] + u1 * dt;
    state[2] = orc_normalize_angle(state[2]);
    state[0] = state[0] + u0 * cos(state[2]) * dt;
    state[1] = state[1] + u0 * sin(state[2]) * dt;
    state[3] = u0;
    state[4] = u1;
}

/* differential_env.py:55-73. */
void orc_motion_velocity(const OrcEnv *e, const double *state_in, const double *cmd,
                         double duration, double *state_out)
{
    double s[5];
    memcpy(s, state_in, sizeof s);
    double t = 0.0;
    while (t < duration) {
        double iu[2];
        iu[0] = (cmd[0] - s[3]) / e->base_dt;
        iu[1] = (cmd[1] - s[4]) / e->base_dt;
        motion_base(e, s, iu, e->base_dt);
        t += e->base_dt;
    }
    memcpy(state_out, s, sizeof s);
}

/* differential_env.py:76-83 (acceleration command; used by SST only). */
void orc_motion(const OrcEnv *e, const double *state_in, const double *acc,
                double duration, double *state_out)
{
    double s[5];
    memcpy(s, state_in, sizeof s);
    double t = 0.0;
    while (t < duration) {
        motion_base(e, s, acc, e->base_dt);
        t += e->base_dt;
    }
    memcpy(state_out, s, sizeof s);
}

/* dwa.py:35-43.  Returns the number of rows written (1 + number of steps). */
int orc_calc_trajectory(const OrcEnv *e, const double *x_init, const double *cmd, double dt,
                        double predict_time, double *traj, int max_rows)
{
    double x[5];
    memcpy(x, x_init, sizeof x);
    memcpy(traj, x, sizeof x);
    int rows = 1;
    double t = 0.0;
    while (t <= predict_time) {
        orc_motion_velocity(e, x, cmd, dt, x);
        if (rows < max_rows) memcpy(traj + 5 * rows, x, sizeof x);
        rows++;
        t += dt;
    }
    return rows;
}

/* np.arange(start, stop, step) for float64 (numpy multiarray ctors.c
 * PyArray_ArangeObj + DOUBLE_fill): len = ceil((stop-start)/step);
 * a[0] = start, a[1] = start+step, a[i] = start + i*(a[1]-a[0]). */
int orc_arange(double start, double stop, double step, double *out, int cap)
{
    double v = (stop - start) / step;
    long len = (long)ceil(v);
    if (len <= 0) return 0;
    double next = start + step;
    double delta = next - start;
    for (long i = 0; i < len && i < cap; i++) {
        if (i == 0) out[i] = start;
        else if (i == 1) out[i] = next;
        else out[i] = start + (double)i * delta;
    }
    return (int)len;
}

/* numpy's pairwise summation (umath loops_utils DOUBLE_pairwise_sum), which is
 * what np.sum runs on a 1-D float64 view (dwa.py:76-78).  stride in elements. */
double orc_pairwise_sum(const double *a, long n, long stride)
{
    if (n < 8) {
        double res = 0.0;
        for (long i = 0; i < n; i++) res += a[i * stride];
        return res;
    } else if (n <= 128) {
        double r[8];
        long i;
        for (int k = 0; k < 8; k++) r[k] = a[k * stride];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; k++) r[k] += a[(i + k) * stride];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i * stride];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return orc_pairwise_sum(a, n2, stride) + orc_pairwise_sum(a + n2 * stride, n - n2, stride);
    }
}

/* ------------------------------------------------------------------ geometry */

/* rigid.py:14-42 with the float32 intermediates numpy produces there:
 * segment is a float32 array (rigid.py:169-174), so rl and squared_length are
 * float32; everything else promotes to float64 (NEP 50). */
static double point_segment_dist(double px, double py, float s0x, float s0y, float s1x, float s1y)
{
    float rlx = s1x - s0x, rly = s1y - s0y;
    float sq = rlx * rlx + rly * rly;
    float nrm = sqrtf(sq);
    float sql = nrm * nrm;
    double cx, cy;
    if (sql == 0.0f) {
        cx = s0x; cy = s0y;
    } else {
        double rpx = px - (double)s0x, rpy = py - (double)s0y;
        double dot = (rpx * (double)rlx + rpy * (double)rly) / (double)sql;
        if (dot < 0.0) { cx = s0x; cy = s0y; }
        else if (dot > 1.0) { cx = s1x; cy = s1y; }
        else { cx = (double)s0x + (double)rlx * dot; cy = (double)s0y + (double)rly * dot; }
    }
    return norm2(px - cx, py - cy);
}

/* rigid.py:168-180.  rect = left bottom right top. */
static double rect_dis2point(const float *r, double px, double py)
{
    float l = r[0], b = r[1], rr = r[2], t = r[3];
    double dis = INFINITY;
    dis = pymin(dis, point_segment_dist(px, py, l, b, l, t));
    dis = pymin(dis, point_segment_dist(px, py, l, t, rr, t));
    dis = pymin(dis, point_segment_dist(px, py, rr, t, rr, b));
    dis = pymin(dis, point_segment_dist(px, py, rr, b, l, b));
    return dis;
}

/* rigid.py:88-116 + 45-85: does segment s->g hit the AABB (slab test)? */
static int segment_hits_aabb(double sx, double sy, double gx, double gy, const float *r)
{
    double dx = gx - sx, dy = gy - sy;
    double len = norm2(dx, dy);
    double ux = dx / len, uy = dy / len;
    double fx = (ux == 0.0) ? 1e300 : 1.0 / ux;
    double fy = (uy == 0.0) ? 1e300 : 1.0 / uy;
    double t1 = ((double)r[0] - sx) * fx;
    double t2 = ((double)r[2] - sx) * fx;
    double t3 = ((double)r[1] - sy) * fy;
    double t4 = ((double)r[3] - sy) * fy;
    double tmin = pymax(pymin(t1, t2), pymin(t3, t4));
    double tmax = pymin(pymax(t1, t2), pymax(t3, t4));
    if (tmax < 0) return 0;
    if (tmin > tmax) return 0;
    double t = (tmin >= 0) ? pymin(tmin, tmax) : tmax;
    double px = sx + ux * t, py = sy + uy * t;
    return norm2(px - sx, py - sy) <= len;
}

/* rigid.py:265-313.  Returns -1 on collision else the rectangle distance. */
double orc_dis2obs(const float *robot_rect, const double *pose, const float *obs)
{
    double c = cos(pose[2]), s = sin(pose[2]);
    float l = robot_rect[0], b = robot_rect[1], r = robot_rect[2], t = robot_rect[3];
    float bx[4] = {l, l, r, r}, by[4] = {b, t, t, b};
    double wx[4], wy[4];
    for (int i = 0; i < 4; i++) {
        /* wTb @ [bx, by, 1] (rigid.py:223-227, 190-199) */
        wx[i] = (c * (double)bx[i] + (-s) * (double)by[i]) + pose[0];
        wy[i] = (s * (double)bx[i] + c * (double)by[i]) + pose[1];
    }
    for (int i = 0; i < 4; i++)
        if ((double)obs[0] <= wx[i] && wx[i] <= (double)obs[2] &&
            (double)obs[1] <= wy[i] && wy[i] <= (double)obs[3])
            return -1.0;
    for (int i = 0; i < 4; i++) {
        int j = (i + 1) % 4;
        if (segment_hits_aabb(wx[i], wy[i], wx[j], wy[j], obs)) return -1.0;
    }
    double dis = INFINITY;
    for (int i = 0; i < 4; i++) dis = pymin(dis, rect_dis2point(obs, wx[i], wy[i]));
    float ox[4] = {obs[0], obs[0], obs[2], obs[2]}, oy[4] = {obs[1], obs[3], obs[3], obs[1]};
    for (int i = 0; i < 4; i++) {
        /* inv(wTb) @ [ox, oy, 1] (rigid.py:231-237) in closed form R^T (p - t) */
        double qx = (double)ox[i] - pose[0], qy = (double)oy[i] - pose[1];
        double px = c * qx + s * qy;
        double py = (-s) * qx + c * qy;
        dis = pymin(dis, rect_dis2point(robot_rect, px, py));
    }
    return (dis > 0.05) ? dis : -1.0;
}

/* CircleRobot.dis2obs (rigid.py:336-347): the centre is wTb @ [0, 0, 1] = the pose's xy. */
double orc_dis2obs_circle(double radius, const double *pose, const float *obs)
{
    double cx = pose[0], cy = pose[1];
    if ((double)obs[0] <= cx && cx <= (double)obs[2] && (double)obs[1] <= cy && cy <= (double)obs[3])
        return -1.0;                                   /* point_in_obstacle, closed (rigid.py:148-156) */
    double dis = rect_dis2point(obs, cx, cy) - radius;
    return (dis > 0.05) ? dis : -1.0;
}

static double env_dis2obs(const OrcEnv *e, const double *state, const float *obs)
{
    return (e->robot_radius > 0.0) ? orc_dis2obs_circle(e->robot_radius, state, obs)
                                   : orc_dis2obs(e->robot_rect, state, obs);
}

/* differential_env.py:146-151 (sequential early exit). obs = M x 4 float32. */
int orc_valid_state(const OrcEnv *e, const float *obs, int M, const double *state)
{
    for (int k = 0; k < M; k++)
        if (env_dis2obs(e, state, obs + 4 * k) == -1.0) return 0;
    return 1;
}

/* differential_env.py:126-132. */
double orc_get_clearance(const OrcEnv *e, const float *obs, int M, const double *state)
{
    double dis = 100.0;
    for (int k = 0; k < M; k++) dis = pymin(dis, env_dis2obs(e, state, obs + 4 * k));
    return dis;
}

/* differential_env.py:135-144 + rigid.py:158-166 (closed intervals). */
void orc_valid_points(const float *obs, int M, const double *pts, int K, uint8_t *valid)
{
    for (int i = 0; i < K; i++) {
        double x = pts[2 * i], y = pts[2 * i + 1];
        int col = 0;
        for (int k = 0; k < M && !col; k++) {
            const float *r = obs + 4 * k;
            col = (x >= (double)r[0] && x <= (double)r[2] && y >= (double)r[1] && y <= (double)r[3]);
        }
        valid[i] = !col;
    }
}

/* differential_gym.py:180-190 with :62-91.  samples = K x 2 body-frame points
 * (first coordinate = forward axis), obs_out = 4 + K doubles. */
void orc_local_obs(const float *obs, int M, const double *state, const double *goal,
                   const double *samples, int K, double *obs_out)
{
    double c = cos(state[2]), s = sin(state[2]);
    /* goal in body frame: inv(wTb) @ [gx, gy, 1] */
    double qx = goal[0] - state[0], qy = goal[1] - state[1];
    obs_out[0] = c * qx + s * qy;
    obs_out[1] = (-s) * qx + c * qy;
    obs_out[2] = state[3];
    obs_out[3] = state[4];
    for (int i = 0; i < K; i++) {
        double bx = samples[2 * i], by = samples[2 * i + 1];
        double p[2];
        p[0] = (c * bx + (-s) * by) + state[0];
        p[1] = (s * bx + c * by) + state[1];
        uint8_t v;
        orc_valid_points(obs, M, p, 1, &v);
        obs_out[4 + i] = v ? 1.0 : 0.0;
    }
}

/* --------------------------------------------------------------------- DWA */

/* dwa.py:45-86.  costs_out (nullable) receives the normalised R x 3 table,
 * cost_sum_out (nullable) the R final costs.  Returns R (rollout count) or
 * -1 if cap is too small. */
int orc_dwa_control(const OrcEnv *e, const OrcDwa *d, const float *obs, int M,
                    const double *x_init, const double *goal, double *opt_v, int *opt_idx,
                    double *opt_cost, double *opt_traj, int traj_cap_rows, int *traj_rows,
                    double *cost_sum_out, int cap)
{
    double dw[4];
    orc_dynamic_window(e, x_init, d->dt, dw);
    int capg = 1 << 16;
    double *vg = (double *)malloc(sizeof(double) * capg);
    double *wg = (double *)malloc(sizeof(double) * capg);
    int nv = orc_arange(dw[0], dw[1] + d->v_res, d->v_res, vg, capg);
    int nw = orc_arange(dw[2], dw[3] + d->w_res, d->w_res, wg, capg);
    long R = (long)nv * nw;
    if (nv > capg || nw > capg || (cost_sum_out && R > cap)) { free(vg); free(wg); return -1; }
    double *cost = (double *)malloc(sizeof(double) * 3 * (R > 0 ? R : 1));
    double traj[64 * 5];
    int end_point = 1;
    long k = 0;
    for (int iv = 0; iv < nv; iv++)
        for (int iw = 0; iw < nw; iw++, k++) {
            double cmd[2] = {vg[iv], wg[iw]};
            int rows = orc_calc_trajectory(e, x_init, cmd, d->dt, d->predict_time, traj, 64);
            if (rows > 64) rows = 64;
            double ang = atan2(goal[1] - traj[5 * end_point + 1], goal[0] - traj[5 * end_point + 0]);
            double c0 = fabs(ang - traj[5 * end_point + 2]);
            c0 = pymin(c0, ORC_TWO_PI - c0);
            double min_dis = 100.0;
            if (d->use_clearance) {
                for (int i = 0; i < rows; i += d->skip) {
                    double dis = pymax(orc_get_clearance(e, obs, M, traj + 5 * i), 0.001);
                    min_dis = pymin(dis, min_dis);
                }
            }
            cost[3 * k + 0] = c0;
            cost[3 * k + 1] = 1.0 / min_dis;
            cost[3 * k + 2] = dw[1] - traj[5 * (rows - 1) + 3];
        }
    for (int i = 0; i < 3; i++) {
        double s = orc_pairwise_sum(cost + i, R, 3);
        if (s != 0.0)
            for (long j = 0; j < R; j++) cost[3 * j + i] /= s;
    }
    long best = 0;
    double bestc = INFINITY;
    for (long j = 0; j < R; j++) {
        double cs = (d->to_goal_cost_gain * cost[3 * j] + d->ob_gain * cost[3 * j + 1]) +
                    d->speed_cost_gain * cost[3 * j + 2];
        if (cost_sum_out) cost_sum_out[j] = cs;
        if (cs < bestc) { bestc = cs; best = j; }
    }
    opt_v[0] = vg[best / nw];
    opt_v[1] = wg[best % nw];
    if (opt_idx) *opt_idx = (int)best;
    if (opt_cost) *opt_cost = bestc;
    if (opt_traj) {
        int rows = orc_calc_trajectory(e, x_init, opt_v, d->dt, d->predict_time, opt_traj, traj_cap_rows);
        if (traj_rows) *traj_rows = rows;
    }
    free(vg); free(wg); free(cost);
    return (int)R;
}

/* ----------------------------------------------------------------- planner */

/* rrt.py:156-161: first argmin of np.linalg.norm over (x, y). */
void orc_nearest(const double *xs, const double *ys, long n, double tx, double ty,
                 int *idx, double *dist)
{
    long best = 0;
    double bd = INFINITY;
    for (long i = 0; i < n; i++) {
        double dd = norm2(xs[i] - tx, ys[i] - ty);
        if (dd < bd) { bd = dd; best = i; }
    }
    *idx = (int)best;
    *dist = bd;
}

/* rrt.py:99-135 (DWA steer), batch-friendly output form:
 *   path      [n_max_extend x 5]  every executed state (step 1..n_steps)
 *   n_steps   executed control steps
 *   status    0 = ran out of steps, 1 = collision, 2 = reached target
 *   node_step [<=n_nodes]  step index (1-based) at which each emitted node sits
 *   ctrl_idx  [n_max_extend] chosen rollout index per step (nullable)
 * Returns the number of nodes emitted (0..n_max_extend/n_tree+1). */
int orc_extend_dwa(const OrcEnv *e, const OrcDwa *d, const float *obs, int M,
                   const double *from_state, const double *target, int n_max_extend, int n_tree,
                   double *path, int *n_steps, int *status, int *node_step, int *ctrl_idx)
{
    double state[5], v[2];
    memcpy(state, from_state, sizeof state);
    int n_nodes = 0, st = 0, n = 0;
    for (int k = 1; k <= n_max_extend; k++) {
        int oi;
        orc_dwa_control(e, d, obs, M, state, target, v, &oi, NULL, NULL, 0, NULL, NULL, 0);
        if (ctrl_idx) ctrl_idx[k - 1] = oi;
        orc_motion_velocity(e, state, v, 0.2, state);
        memcpy(path + 5 * (k - 1), state, sizeof state);
        n = k;
        if (!orc_valid_state(e, obs, M, state)) { st = 1; break; }
        if (norm2(state[0] - target[0], state[1] - target[1]) <= 1.0) {
            node_step[n_nodes++] = k; st = 2; break;
        }
        if (k % n_tree == 0) node_step[n_nodes++] = k;
    }
    *n_steps = n;
    *status = st;
    return n_nodes;
}

/* One control step of the RL steer loop AFTER the policy produced `action`
 * (rrt_rl.py:43-57 / differential_gym.py:136-145): advance, then classify.
 * Returns 0 continue, 1 collision, 2 reached. */
int orc_rl_step(const OrcEnv *e, const float *obs, int M, double *state, const double *action,
                const double *target)
{
    orc_motion_velocity(e, state, action, 1.0 / 5.0, state);
    if (!orc_valid_state(e, obs, M, state)) return 1;
    if (norm2(state[0] - target[0], state[1] - target[1]) <= 1.0) return 2;
    return 0;
}

/* ------------------------------------------------ batch drivers (pthreads) */
/* Used by the CPU-baseline timing legs of bench.py and by tests that need many
 * oracle evaluations; a plain work-stealing parallel-for over independent items. */
#include <pthread.h>

typedef void (*orc_item_fn)(void *ctx, int i);
typedef struct { orc_item_fn fn; void *ctx; int n; int next; pthread_mutex_t mu; } OrcPool;

static void *orc_pool_worker(void *p)
{
    OrcPool *pool = (OrcPool *)p;
    for (;;) {
        pthread_mutex_lock(&pool->mu);
        int i = pool->next++;
        pthread_mutex_unlock(&pool->mu);
        if (i >= pool->n) break;
        pool->fn(pool->ctx, i);
    }
    return NULL;
}

static void orc_parallel_for(orc_item_fn fn, void *ctx, int n, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 512) threads = 512;
    OrcPool pool = {fn, ctx, n, 0, PTHREAD_MUTEX_INITIALIZER};
    if (threads == 1 || n <= 1) { for (int i = 0; i < n; i++) fn(ctx, i); return; }
    pthread_t th[512];
    int started = 0;
    for (int t = 0; t < threads; t++)
        if (pthread_create(&th[started], NULL, orc_pool_worker, &pool) == 0) started++;
    if (started == 0) orc_pool_worker(&pool);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

typedef struct {
    const OrcEnv *e; const OrcDwa *d; const float *obs; int M;
    const double *from_states, *targets; int n_max_extend, n_tree;
    double *paths; int *n_steps, *status, *node_steps, *n_nodes, *ctrl_idx;
} ExtendCtx;

static void extend_item(void *p, int q)
{
    ExtendCtx *c = (ExtendCtx *)p;
    int per = c->n_max_extend / c->n_tree + 1;
    c->n_nodes[q] = orc_extend_dwa(c->e, c->d, c->obs, c->M, c->from_states + 5 * q,
                                   c->targets + 5 * q, c->n_max_extend, c->n_tree,
                                   c->paths + (long)5 * c->n_max_extend * q, c->n_steps + q,
                                   c->status + q, c->node_steps + per * q,
                                   c->ctrl_idx ? c->ctrl_idx + (long)c->n_max_extend * q : NULL);
}

void orc_extend_dwa_batch(const OrcEnv *e, const OrcDwa *d, const float *obs, int M,
                          const double *from_states, const double *targets, int Q,
                          int n_max_extend, int n_tree, double *paths, int *n_steps,
                          int *status, int *node_steps, int *n_nodes, int *ctrl_idx, int threads)
{
    ExtendCtx c = {e, d, obs, M, from_states, targets, n_max_extend, n_tree,
                   paths, n_steps, status, node_steps, n_nodes, ctrl_idx};
    orc_parallel_for(extend_item, &c, Q, threads);
}

typedef struct { const OrcEnv *e; const float *obs; int M; const double *states; uint8_t *valid; double *clearance; } ValidCtx;

static void valid_item(void *p, int i)
{
    ValidCtx *c = (ValidCtx *)p;
    if (c->valid) c->valid[i] = (uint8_t)orc_valid_state(c->e, c->obs, c->M, c->states + 5 * i);
    if (c->clearance) c->clearance[i] = orc_get_clearance(c->e, c->obs, c->M, c->states + 5 * i);
}

void orc_valid_state_batch(const OrcEnv *e, const float *obs, int M, const double *states, int B,
                           uint8_t *valid, double *clearance, int threads)
{
    ValidCtx c = {e, obs, M, states, valid, clearance};
    orc_parallel_for(valid_item, &c, B, threads);
}

typedef struct { const double *xs, *ys; const int *n_nodes; long stride; const double *targets; int *idx; double *dist; } NearCtx;

static void near_item(void *p, int q)
{
    NearCtx *c = (NearCtx *)p;
    orc_nearest(c->xs + c->stride * q, c->ys + c->stride * q, c->n_nodes[q], c->targets[2 * q],
                c->targets[2 * q + 1], c->idx + q, c->dist + q);
}

void orc_nearest_batch(const double *xs, const double *ys, const int *n_nodes, long stride,
                       const double *targets, int Q, int *idx, double *dist, int threads)
{
    NearCtx c = {xs, ys, n_nodes, stride, targets, idx, dist};
    orc_parallel_for(near_item, &c, Q, threads);
}

typedef struct {
    const OrcEnv *e; const OrcDwa *d; const float *obs; int M; const double *states, *goals;
    double *opt_v; int *opt_idx; double *opt_cost; int *n_rollouts;
} DwaCtx;

static void dwa_item(void *p, int i)
{
    DwaCtx *c = (DwaCtx *)p;
    c->n_rollouts[i] = orc_dwa_control(c->e, c->d, c->obs, c->M, c->states + 5 * i, c->goals + 2 * i,
                                       c->opt_v + 2 * i, c->opt_idx + i, c->opt_cost + i, NULL, 0,
                                       NULL, NULL, 0);
}

void orc_dwa_control_batch(const OrcEnv *e, const OrcDwa *d, const float *obs, int M,
                           const double *states, const double *goals, int B, double *opt_v,
                           int *opt_idx, double *opt_cost, int *n_rollouts, int threads)
{
    DwaCtx c = {e, d, obs, M, states, goals, opt_v, opt_idx, opt_cost, n_rollouts};
    orc_parallel_for(dwa_item, &c, B, threads);
}

typedef struct { const float *obs; int M; const double *states, *goals, *samples; int K; double *out; } ObsCtx;

static void obs_item(void *p, int i)
{
    ObsCtx *c = (ObsCtx *)p;
    orc_local_obs(c->obs, c->M, c->states + 5 * i, c->goals + 2 * i, c->samples, c->K,
                  c->out + (long)(4 + c->K) * i);
}

void orc_local_obs_batch(const float *obs, int M, const double *states, const double *goals,
                         const double *samples, int K, int B, double *out, int threads)
{
    ObsCtx c = {obs, M, states, goals, samples, K, out};
    orc_parallel_for(obs_item, &c, B, threads);
}
