// libcrlk.so -- environment, DWA, nearest-node and fused-extend kernels (sm_100a).
// Compiled with --fmad=false: the float64 arithmetic must round exactly like
// the reference's scalar Python/numpy code; fused operations are explicit.
#include <math_constants.h>
#include <stdarg.h>
#include <stdlib.h>
#include <algorithm>

#include "crlk_device.cuh"

// ------------------------------------------------------------ error plumbing

static thread_local char g_err[512] = "";

void crlk_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char *crlk_last_error(void) { return g_err; }
extern "C" int crlk_abi_version(void) { return CRLK_ABI_VERSION; }

extern "C" int crlk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int crlk_require_device(void)
{
    if (crlk_device_count() <= 0) {
        crlk_set_error("no CUDA device visible: libcrlk has no CPU fallback");
        return CRLK_E_NODEVICE;
    }
    return CRLK_OK;
}

// Float loops of the reference whose trip count depends on rounding
// (differential_env.py:69 `while t < duration`, dwa.py:39 `while t <= predict_time`).
int crlk_count_loop_lt(double limit, double step)
{
    int n = 0;
    double t = 0.0;
    while (t < limit && n < 1000000) { t += step; n++; }
    return n;
}
int crlk_count_loop_le(double limit, double step)
{
    int n = 0;
    double t = 0.0;
    while (t <= limit && n < 1000000) { t += step; n++; }
    return n;
}

// ------------------------------------------------------------------- handles

#ifndef CRLK_GRID_CELL
#define CRLK_GRID_CELL 0.5f   // obstacle grid pitch in metres (tuned on the cfg4 map, see profiles/)
#endif
#ifndef CRLK_RASTER_SUB
#define CRLK_RASTER_SUB 4     // occupancy raster cells per grid cell and axis
#endif

static void fill_envp(const CrlkEnvParams *in, EnvP *p)
{
    p->max_v = in->max_v; p->min_v = in->min_v; p->max_w = in->max_w;
    p->max_acc_v = in->max_acc_v; p->max_acc_w = in->max_acc_w; p->base_dt = in->base_dt;
    p->rl = in->robot_rect[0]; p->rb = in->robot_rect[1];
    p->rr = in->robot_rect[2]; p->rt = in->robot_rect[3];
    float m = 0.f;
    const float xs[2] = {p->rl, p->rr}, ys[2] = {p->rb, p->rt};
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) m = fmaxf(m, sqrtf(xs[i] * xs[i] + ys[j] * ys[j]));
    p->rc = m * 1.0001f;
    p->radius = in->robot_radius > 0 ? in->robot_radius : 0.0;
    if (p->radius > 0) p->rc = (float)p->radius * 1.0001f;
}

// Uniform grid over the obstacles' bounding box: cell (cx, cy) lists every
// obstacle whose AABB touches it (CSR, row-major cells).
static int build_grid(CrlkEnv *env, const float *r, int M)
{
    float x0 = 0.f, y0 = 0.f, x1 = 1.f, y1 = 1.f;
    for (int i = 0; i < M; i++) {
        if (i == 0) { x0 = r[0]; y0 = r[1]; x1 = r[2]; y1 = r[3]; }
        x0 = fminf(x0, r[4 * i]); y0 = fminf(y0, r[4 * i + 1]);
        x1 = fmaxf(x1, r[4 * i + 2]); y1 = fmaxf(y1, r[4 * i + 3]);
    }
    float ext = fmaxf(x1 - x0, y1 - y0);
    float cell = fmaxf(CRLK_GRID_CELL, ext / 256.0f);
    int ncx = (int)floorf((x1 - x0) / cell) + 1, ncy = (int)floorf((y1 - y0) / cell) + 1;
    float ginv = 1.0f / cell;
    auto cx = [&](float v) { return (int)fminf(fmaxf(floorf((v - x0) * ginv), 0.f), (float)(ncx - 1)); };
    auto cy = [&](float v) { return (int)fminf(fmaxf(floorf((v - y0) * ginv), 0.f), (float)(ncy - 1)); };
    int ncell = ncx * ncy;
    int *start = (int *)calloc((size_t)ncell + 1, sizeof(int));
    for (int i = 0; i < M; i++)
        for (int yy = cy(r[4 * i + 1]); yy <= cy(r[4 * i + 3]); yy++)
            for (int xx = cx(r[4 * i]); xx <= cx(r[4 * i + 2]); xx++) start[yy * ncx + xx + 1]++;
    for (int c = 0; c < ncell; c++) start[c + 1] += start[c];
    int total = start[ncell];
    int *items = (int *)malloc(sizeof(int) * (size_t)(total + 1));
    int *fill = (int *)calloc((size_t)ncell, sizeof(int));
    for (int i = 0; i < M; i++)
        for (int yy = cy(r[4 * i + 1]); yy <= cy(r[4 * i + 3]); yy++)
            for (int xx = cx(r[4 * i]); xx <= cx(r[4 * i + 2]); xx++) {
                int c = yy * ncx + xx;
                items[start[c] + fill[c]++] = i;
            }
    // Occupancy raster for the point queries of the local observation (729 per observation): a
    // cell is 0 when no obstacle comes within `margin` of it, 1 when the obstacles together cover
    // it and `margin` around it, else 2.  The margin (>= 1e-4 m) is far larger than the float32 error of
    // the device's cell index, so a 0 / 1 answer is exactly what the scan over all obstacles gives
    // (closed intervals, rigid.py:158-166); only cells marked 2 still run that scan.
    const int RS = CRLK_RASTER_SUB;
    const int rnx = ncx * RS, rny = ncy * RS;
    const double h = (double)cell / RS;
    const double margin = fmax(1e-4, 1e-5 * (double)ext);
    uint8_t *raster = (uint8_t *)malloc((size_t)rnx * rny);
    enum { RASTER_CAND = 64 };
    int cand[RASTER_CAND];
    double xs[2 * RASTER_CAND + 2], ys[2 * RASTER_CAND + 2];
    for (int ry = 0; ry < rny; ry++)
        for (int rx = 0; rx < rnx; rx++) {
            const double cl = (double)x0 + rx * h - margin, cr = (double)x0 + (rx + 1) * h + margin;
            const double cb = (double)y0 + ry * h - margin, ct = (double)y0 + (ry + 1) * h + margin;
            // obstacles that reach into the (expanded) cell, clipped to it
            int nc = 0;
            for (int gy = cy((float)cb); gy <= cy((float)ct); gy++)
                for (int gx = cx((float)cl); gx <= cx((float)cr); gx++)
                    for (int k = start[gy * ncx + gx]; k < start[gy * ncx + gx + 1] && nc < RASTER_CAND; k++) {
                        const float *o = r + 4 * items[k];
                        const double l = o[0], b = o[1], rr = o[2], t = o[3];
                        if (rr < cl || l > cr || t < cb || b > ct) continue;
                        bool dup = false;
                        for (int q = 0; q < nc; q++) dup |= (cand[q] == items[k]);
                        if (!dup) cand[nc++] = items[k];
                    }
            uint8_t v = 0;
            if (nc >= RASTER_CAND) v = 2;
            else if (nc > 0) {
                // covered by the UNION of the (closed) rectangles?  Cut the cell at every rectangle
                // edge inside it; the cell is covered iff the centre of every elementary piece is.
                // (Abutting tiles cover a cell together; a one-ulp gap between two tiles leaves an
                // uncovered sliver, and the cell stays "mixed".)
                int nx = 0, ny = 0;
                xs[nx++] = cl; xs[nx++] = cr; ys[ny++] = cb; ys[ny++] = ct;
                for (int q = 0; q < nc; q++) {
                    const float *o = r + 4 * cand[q];
                    if (o[0] > cl && o[0] < cr) xs[nx++] = o[0];
                    if (o[2] > cl && o[2] < cr) xs[nx++] = o[2];
                    if (o[1] > cb && o[1] < ct) ys[ny++] = o[1];
                    if (o[3] > cb && o[3] < ct) ys[ny++] = o[3];
                }
                std::sort(xs, xs + nx); std::sort(ys, ys + ny);
                bool all = true;
                for (int a = 0; a + 1 < nx && all; a++)
                    for (int c2 = 0; c2 + 1 < ny && all; c2++) {
                        if (xs[a + 1] == xs[a] || ys[c2 + 1] == ys[c2]) continue;
                        const double mx = 0.5 * (xs[a] + xs[a + 1]), my = 0.5 * (ys[c2] + ys[c2 + 1]);
                        bool in = false;
                        for (int q = 0; q < nc && !in; q++) {
                            const float *o = r + 4 * cand[q];
                            in = (mx >= o[0] && mx <= o[2] && my >= o[1] && my <= o[3]);
                        }
                        all = in;
                    }
                v = all ? 1 : 2;
            }
            raster[(size_t)ry * rnx + rx] = v;
        }
    if (env->d_raster) cudaFree(env->d_raster);
    env->d_raster = nullptr;
    cudaError_t e3 = cudaMalloc(&env->d_raster, (size_t)rnx * rny);
    if (e3 == cudaSuccess) e3 = cudaMemcpy(env->d_raster, raster, (size_t)rnx * rny, cudaMemcpyHostToDevice);
    free(raster);
    if (e3 != cudaSuccess) {
        free(start); free(items); free(fill);
        crlk_set_error("occupancy raster upload failed: %s", cudaGetErrorString(e3));
        return CRLK_E_CUDA;
    }
    if (env->d_cell_start) cudaFree(env->d_cell_start);
    if (env->d_cell_items) cudaFree(env->d_cell_items);
    if (env->d_cell_rects) cudaFree(env->d_cell_rects);
    env->d_cell_start = env->d_cell_items = nullptr;
    env->d_cell_rects = nullptr;
    // the rectangles once more in cell order (16 B per list entry; ~1 MB for the 10,004-obstacle map)
    float4 *crects = (float4 *)malloc(sizeof(float4) * (size_t)(total + 1));
    for (int k = 0; k < total; k++)
        crects[k] = make_float4(r[4 * items[k]], r[4 * items[k] + 1], r[4 * items[k] + 2], r[4 * items[k] + 3]);
    cudaError_t e1 = cudaMalloc(&env->d_cell_start, sizeof(int) * ((size_t)ncell + 1));
    cudaError_t e2 = cudaMalloc(&env->d_cell_items, sizeof(int) * (size_t)(total + 1));
    if (e1 == cudaSuccess && e2 == cudaSuccess) e2 = cudaMalloc(&env->d_cell_rects, sizeof(float4) * (size_t)(total + 1));
    if (e1 == cudaSuccess && e2 == cudaSuccess) {
        e1 = cudaMemcpy(env->d_cell_start, start, sizeof(int) * ((size_t)ncell + 1), cudaMemcpyHostToDevice);
        e2 = cudaMemcpy(env->d_cell_items, items, sizeof(int) * (size_t)(total + 1), cudaMemcpyHostToDevice);
        if (e2 == cudaSuccess)
            e2 = cudaMemcpy(env->d_cell_rects, crects, sizeof(float4) * (size_t)total, cudaMemcpyHostToDevice);
    }
    free(start); free(items); free(fill); free(crects);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        crlk_set_error("obstacle grid upload failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        return CRLK_E_CUDA;
    }
    env->os.rects = env->d_rects; env->os.cell_start = env->d_cell_start;
    env->os.cell_items = env->d_cell_items;
    env->os.cell_rects = env->d_cell_rects;
    env->os.M = M; env->os.ncx = ncx; env->os.ncy = ncy;
    env->os.gx0 = x0; env->os.gy0 = y0; env->os.ginv = ginv;
    env->os.raster = env->d_raster; env->os.rnx = rnx; env->os.rny = rny;
    env->os.rinv = (float)(1.0 / h);
    return CRLK_OK;
}

extern "C" int crlk_env_set_obs(CrlkEnv *env, const float *rects_host, int32_t M)
{
    CRLK_REQUIRE(env != nullptr, "env is NULL");
    CRLK_REQUIRE(M >= 0 && (M == 0 || rects_host != nullptr), "rects");
    CRLK_CUDA(cudaSetDevice(env->device));
    memset(&env->os, 0, sizeof env->os);          // nothing published until every upload has succeeded
    if (env->d_rects) { CRLK_CUDA(cudaFree(env->d_rects)); env->d_rects = nullptr; }
    env->M = M;
    // one spare element so that a zero-obstacle env still has a valid pointer
    CRLK_CUDA(cudaMalloc(&env->d_rects, sizeof(float4) * (size_t)(M + 1)));
    if (M > 0)
        CRLK_CUDA(cudaMemcpy(env->d_rects, rects_host, sizeof(float4) * (size_t)M,
                             cudaMemcpyHostToDevice));
    return build_grid(env, rects_host, M);
}

extern "C" int crlk_env_destroy(CrlkEnv *env);

extern "C" int crlk_env_create(const CrlkEnvParams *params, const float *rects_host, int32_t M,
                               CrlkEnv **out)
{
    CRLK_REQUIRE(params != nullptr && out != nullptr, "params/out is NULL");
    CRLK_REQUIRE(params->base_dt > 0, "base_dt must be positive");
    int rc = crlk_require_device();
    if (rc) return rc;
    CrlkEnv *e = new CrlkEnv();
    memset(e, 0, sizeof *e);
    CRLK_CUDA(cudaGetDevice(&e->device));
    e->host = *params;
    fill_envp(params, &e->p);
    rc = crlk_env_set_obs(e, rects_host, M);
    if (rc) { crlk_env_destroy(e); return rc; }
    *out = e;
    return CRLK_OK;
}

extern "C" int crlk_env_destroy(CrlkEnv *env)
{
    if (!env) return CRLK_OK;
    if (env->d_rects) cudaFree(env->d_rects);
    if (env->d_cell_start) cudaFree(env->d_cell_start);
    if (env->d_cell_items) cudaFree(env->d_cell_items);
    if (env->d_cell_rects) cudaFree(env->d_cell_rects);
    if (env->d_raster) cudaFree(env->d_raster);
    delete env;
    return CRLK_OK;
}

extern "C" int crlk_env_num_obs(const CrlkEnv *env) { return env ? env->M : 0; }

// --------------------------------------------------- collision / clearance (K2)

#define WQ_CAP 64   // per-warp survivor queue

// Drain a warp's survivor queue: each lane evaluates one obstacle exactly.
__device__ __forceinline__ void warp_drain(const EnvP &e, const RobotPose &g, const float4 *rects,
                                           const int *queue, int count, double &best,
                                           bool &definite, bool &marginal)
{
    int lane = threadIdx.x & 31;
    for (int base = 0; base < count; base += 32) {
        int k = base + lane;
        if (k < count) {
            bool d, m;
            double dis = dis2obs_exact(e, g, __ldg(&rects[queue[k]]), d, m);
            definite |= d;
            marginal |= m;
            best = fmin(best, dis);   // -1 (collision) wins every min, as in get_clearance
        }
    }
}

// valid_state_check, flag only: one warp per pose screens the grid cells around the pose at the
// collision radius (FP32), queues the survivors and tests them exactly (FP64), one per lane.
__global__ void __launch_bounds__(256)
k_valid_state(EnvP e, ObsSet os, const double *__restrict__ states, int B, uint8_t *valid,
              uint8_t *near_boundary)
{
    __shared__ int s_queue[8][WQ_CAP];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int pose = blockIdx.x * 8 + warp;
    if (pose >= B) return;
    const float4 *__restrict__ rects = os.rects;
    const double *st = states + 5 * (size_t)pose;
    RobotPose g;
    robot_pose_setup(e, st[0], st[1], st[2], g);
    float px = (float)g.x, py = (float)g.y;
    int *queue = s_queue[warp];
    double best = CRLK_CLEARANCE_CAP;
    bool definite = false, marginal = false;
    const float thr = e.rc + (float)CRLK_COLLISION_DIS + CRLK_CULL_MARGIN;
    const float thr2 = thr * thr;
    int count = 0;   // warp-uniform
    grid_visit(os, px, py, thr, lane, 32, [&](int o, bool in) {
        bool keep = in && aabb_center_d2(__ldg(&rects[in ? o : 0]), px, py) <= thr2;
        unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (mask) {
            int pos = count + __popc(mask & ((1u << lane) - 1u));
            if (keep) queue[pos] = o;
            count += __popc(mask);
            __syncwarp();
            if (count > WQ_CAP - 32) {
                warp_drain(e, g, rects, queue, count, best, definite, marginal);
                count = 0;
                __syncwarp();
            }
        }
    });
    __syncwarp();
    warp_drain(e, g, rects, queue, count, best, definite, marginal);
    best = warp_min_d(best);
    bool any_def = __any_sync(0xffffffffu, definite);
    bool any_mar = __any_sync(0xffffffffu, marginal);
    if (lane == 0) {
        if (valid) valid[pose] = (best < 0.0) ? 0 : 1;
        if (near_boundary) near_boundary[pose] = (any_mar && !any_def) ? 1 : 0;
    }
}

// get_clearance (and the validity flag with it): one thread per pose, ring search over the grid
// (clearance_grid_thread).
__global__ void __launch_bounds__(128)
k_clearance(EnvP e, ObsSet os, const double *__restrict__ states, int B, uint8_t *valid, double *clearance,
            uint8_t *near_boundary)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const double *st = states + 5 * (size_t)i;
    bool d, m;
    double c = clearance_grid_thread(e, os, st[0], st[1], st[2], &d, &m);
    if (valid) valid[i] = (c < 0.0) ? 0 : 1;
    clearance[i] = c;
    if (near_boundary) near_boundary[i] = (m && !d) ? 1 : 0;
}

extern "C" int crlk_valid_state(const CrlkEnv *env, const double *states, int32_t B, uint8_t *valid,
                                double *clearance, uint8_t *near_boundary, void *stream)
{
    CRLK_ENV_OK(env);
    CRLK_REQUIRE(B >= 0 && (B == 0 || states != nullptr), "states");
    if (B == 0) return CRLK_OK;
    if (clearance)
        k_clearance<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->p, env->os, states, B, valid,
                                                                      clearance, near_boundary);
    else
        k_valid_state<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(env->p, env->os, states, B, valid,
                                                                    near_boundary);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

// valid_point_check: one thread per point, obstacles broadcast through L1.
__global__ void __launch_bounds__(256)
k_valid_points(const float4 *__restrict__ rects, int M, const double *__restrict__ pts, int K,
               uint8_t *valid)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    double x = pts[2 * (size_t)i], y = pts[2 * (size_t)i + 1];
    bool col = false;
    for (int o = 0; o < M; o++) {
        float4 r = __ldg(&rects[o]);
        col |= (x >= (double)r.x && x <= (double)r.z && y >= (double)r.y && y <= (double)r.w);
    }
    valid[i] = col ? 0 : 1;
}

extern "C" int crlk_valid_points(const CrlkEnv *env, const double *points, int32_t K, uint8_t *valid,
                                 void *stream)
{
    CRLK_REQUIRE(env != nullptr && valid != nullptr, "env/valid is NULL");
    CRLK_REQUIRE(K >= 0 && (K == 0 || points != nullptr), "points");
    if (K == 0) return CRLK_OK;
    k_valid_points<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(env->d_rects, env->M, points,
                                                                     K, valid);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

// ------------------------------------------------------------ dynamics kernels

__global__ void k_motion(EnvP e, const double *__restrict__ sin_, const double *__restrict__ cmd,
                         int n_sub, int B, int accel, double *sout)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double s[5];
#pragma unroll
    for (int k = 0; k < 5; k++) s[k] = sin_[5 * (size_t)i + k];
    double c0 = cmd[2 * (size_t)i], c1 = cmd[2 * (size_t)i + 1];
    if (!accel) {
        motion_velocity_dev(e, s, c0, c1, n_sub);
    } else {
        for (int k = 0; k < n_sub; k++) {
            double u0 = acc_substep(s[3], c0, e.max_acc_v, e.min_v, e.max_v, e.base_dt);
            double u1 = acc_substep(s[4], c1, e.max_acc_w, -e.max_w, e.max_w, e.base_dt);
            pose_substep(s, u0, u1, e.base_dt);
        }
    }
#pragma unroll
    for (int k = 0; k < 5; k++) sout[5 * (size_t)i + k] = s[k];
}

static int launch_motion(const CrlkEnv *env, const double *sin_, const double *cmd, double duration,
                         int B, int accel, double *sout, void *stream)
{
    CRLK_ENV_OK(env);
    CRLK_REQUIRE(B >= 0 && (B == 0 || (sin_ && cmd && sout)), "NULL array");
    if (B == 0) return CRLK_OK;
    int n_sub = crlk_count_loop_lt(duration, env->p.base_dt);
    k_motion<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->p, sin_, cmd, n_sub, B, accel,
                                                               sout);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

extern "C" int crlk_motion_velocity(const CrlkEnv *env, const double *states_in, const double *cmds,
                                    double duration, int32_t B, double *states_out, void *stream)
{
    return launch_motion(env, states_in, cmds, duration, B, 0, states_out, stream);
}

extern "C" int crlk_motion(const CrlkEnv *env, const double *states_in, const double *acc,
                           double duration, int32_t B, double *states_out, void *stream)
{
    return launch_motion(env, states_in, acc, duration, B, 1, states_out, stream);
}

__global__ void k_dynamic_window(EnvP e, const double *__restrict__ states, double dt, int B,
                                 double *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double v = states[5 * (size_t)i + 3], w = states[5 * (size_t)i + 4];
    out[4 * (size_t)i + 1] = pymin(v + e.max_acc_v * dt, e.max_v);
    out[4 * (size_t)i + 0] = pymax(v - e.max_acc_v * dt, e.min_v);
    out[4 * (size_t)i + 3] = pymin(w + e.max_acc_w * dt, e.max_w);
    out[4 * (size_t)i + 2] = pymax(w - e.max_acc_w * dt, -e.max_w);
}

extern "C" int crlk_dynamic_window(const CrlkEnv *env, const double *states, double dt, int32_t B,
                                   double *out, void *stream)
{
    CRLK_ENV_OK(env);
    CRLK_REQUIRE(B >= 0 && (B == 0 || (states && out)), "NULL array");
    if (B == 0) return CRLK_OK;
    k_dynamic_window<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->p, states, dt, B, out);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

// -------------------------------------------------------- local observation (R9)


// One CTA per observation, three sample points per thread.  Each point is a
// point query against the obstacle grid: only the lists of the cell(s) holding
// the point are tested (closed intervals, rigid.py:158-166).  A point inside an
// obstacle always falls in a cell of that obstacle's range because the cell
// function is monotone, so the result equals the scan over all obstacles.
// Outputs (each nullable): float64 rows (the reference dtype), zero-padded
// float32 rows, and/or the three bf16 operand planes of the actor's first layer
// ([3][Mpad][Kp]; the map entries are exactly 0/1 in bf16, so planes 2 and 3 are
// non-zero only in the first four columns and only their first 64 are written).
// With `list` the CTA index is a row of a compacted batch: row b reads state and
// goal of query list[b] and writes output row b; rows >= *count do nothing.
template <bool WANT_NEAR>
__global__ void __launch_bounds__(256)
k_local_obs(ObsGrid og, ObsSet os, const double *__restrict__ states,
            const double *__restrict__ goals, int goal_stride, int B, double *obs64, float *obs32,
            int ld32, uint8_t *near_boundary, const int *__restrict__ list, const int *__restrict__ count,
            __nv_bfloat16 *planes, int Mpad, int Kp)
{
    __shared__ int s_near;
    __shared__ double s_cs[2];
    const int b = blockIdx.x;
    if (b >= B) return;
    if (count && b >= *count) return;
    const int q = list ? list[b] : b;
    const double *st = states + 5 * (size_t)q;
    double x = st[0], y = st[1], th = st[2];
    if (threadIdx.x == 0) {           // one float64 sincos per observation, not per thread
        double sn0, cs0;
        crlk_sincos(th, &sn0, &cs0);
        s_cs[0] = cs0; s_cs[1] = sn0;
        s_near = 0;
    }
    __syncthreads();
    const double cs = s_cs[0], sn = s_cs[1];
    bool nearp = false;
    __nv_bfloat16 *p0 = planes ? planes + (size_t)b * Kp : nullptr;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int p = threadIdx.x + k * 256;
        if (p >= OBS_NPTS) break;
        int ib = p / OBS_SIDE, il = p - ib * OBS_SIDE;
        double bx = arange_at(og.ba0, og.ba1, og.bad, ib);
        double by = arange_at(og.lr0, og.lr1, og.lrd, il);
        double px, py;
        body_to_world(cs, sn, x, y, bx, by, px, py);     // differential_gym.py:83 (wTb @ [bx, by, 1])
        bool occ = local_map_occupied<WANT_NEAR>(os, px, py, nearp);
        double v = occ ? 0.0 : 1.0;
        if (obs64) obs64[(size_t)b * (4 + OBS_NPTS) + 4 + p] = v;
        if (obs32) obs32[(size_t)b * ld32 + 4 + p] = (float)v;
        if (p0) p0[4 + p] = __float2bfloat16_rn((float)v);
    }
    if (WANT_NEAR && nearp) s_near = 1;
    if (planes) {
        const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
        for (int c = 4 + OBS_NPTS + threadIdx.x; c < Kp; c += 256) p0[c] = z;
        if (threadIdx.x >= 4 && threadIdx.x < 64) {
            planes[((size_t)Mpad + b) * Kp + threadIdx.x] = z;
            planes[((size_t)2 * Mpad + b) * Kp + threadIdx.x] = z;
        }
    }
    if (threadIdx.x == 0) {
        const double *g = goals + (size_t)goal_stride * q;
        double qx = g[0] - x, qy = g[1] - y;     // differential_gym.py:186-187 as R^T (g - t)
        double gbx = cs * qx + sn * qy;
        double gby = (-sn) * qx + cs * qy;
        if (obs64) {
            double *o = obs64 + (size_t)b * (4 + OBS_NPTS);
            o[0] = gbx; o[1] = gby; o[2] = st[3]; o[3] = st[4];
        }
        if (obs32) {
            float *o = obs32 + (size_t)b * ld32;
            o[0] = (float)gbx; o[1] = (float)gby; o[2] = (float)st[3]; o[3] = (float)st[4];
            for (int k = 4 + OBS_NPTS; k < ld32; k++) o[k] = 0.f;
        }
        if (planes) {
            const float h[4] = {(float)gbx, (float)gby, (float)st[3], (float)st[4]};
            for (int k = 0; k < 4; k++) {
                __nv_bfloat16 a0, a1, a2;
                crlk_split3(h[k], a0, a1, a2);
                p0[k] = a0;
                planes[((size_t)Mpad + b) * Kp + k] = a1;
                planes[((size_t)2 * Mpad + b) * Kp + k] = a2;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && near_boundary) near_boundary[b] = (uint8_t)s_near;
}

// The raster path with one WARP per observation (eight per CTA, no block barrier): lane = column of the
// 27 x 27 sample grid, three grid rows per iteration, the raster screen on a float32 evaluation of the
// sample points (its margin covers the float32 error, crlk_device.cuh), the few points in partly covered
// raster cells queued and decided exactly in float64 with all lanes busy -- the observation of
// k_rl_advance as a kernel of its own.  The serial, correctly rounded sincos of the heading (one lane)
// is hidden by the other warps of the SM; with one CTA per observation it was a barrier-separated
// prologue of every CTA.  Same outputs, bit for bit, as k_local_obs<false>.
#define LOW_QUEUE 64
__global__ void __launch_bounds__(256)
k_local_obs_w(ObsGrid og, ObsSet os, const double *__restrict__ states, const double *__restrict__ goals,
              int goal_stride, int B, double *obs64, float *obs32, int ld32, const int *__restrict__ list,
              const int *__restrict__ count, __nv_bfloat16 *planes, int Mpad, int Kp)
{
    __shared__ int s_queue[8][LOW_QUEUE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * 8 + warp;
    if (b >= B) return;
    if (count && b >= *count) return;
    const int q = list ? list[b] : b;
    const double *st = states + 5 * (size_t)q;
    double sn = 0.0, cs = 0.0;
    if (lane == 0) crlk_sincos(st[2], &sn, &cs);
    const double px0 = st[0], py0 = st[1];
    const double ccs = __shfl_sync(0xffffffffu, cs, 0), ssn = __shfl_sync(0xffffffffu, sn, 0);
    __nv_bfloat16 *p0 = planes ? planes + (size_t)b * Kp : nullptr;
    auto put = [&](int p, bool occ) {
        const double v = occ ? 0.0 : 1.0;
        if (obs64) obs64[(size_t)b * (4 + OBS_NPTS) + 4 + p] = v;
        if (obs32) obs32[(size_t)b * ld32 + 4 + p] = (float)v;
        if (p0) p0[4 + p] = __float2bfloat16_rn((float)v);
    };
    const bool on = lane < OBS_SIDE;
    const double by = arange_at(og.lr0, og.lr1, og.lrd, on ? lane : 0);
    int *mq = s_queue[warp];
    int nq = 0;
    auto drain = [&]() {
        __syncwarp();
        for (int i = lane; i < nq; i += 32) {
            const int p = mq[i], ib = p / OBS_SIDE, il = p - ib * OBS_SIDE;
            double px, py;
            body_to_world(ccs, ssn, px0, py0, arange_at(og.ba0, og.ba1, og.bad, ib), arange_at(og.lr0, og.lr1, og.lrd, il),
                          px, py);                                    // differential_gym.py:83, exact
            bool nearp = false;
            put(p, local_map_occupied<true>(os, px, py, nearp));
        }
        nq = 0;
        __syncwarp();
    };
    const float ccs_f = (float)ccs, ssn_f = (float)ssn, by_f = (float)by;
    const float ox_f = fmaf(-ssn_f, by_f, (float)px0), oy_f = fmaf(ccs_f, by_f, (float)py0);
    const float my_bx_f = (float)arange_at(og.ba0, og.ba1, og.bad, on ? lane : 0);
    static_assert(OBS_SIDE % 3 == 0, "three grid rows per iteration");
    for (int ib = 0; ib < OBS_SIDE; ib += 3) {
        unsigned cls[3];
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const float bx_f = __shfl_sync(0xffffffffu, my_bx_f, ib + u);
            cls[u] = on ? local_map_raster_f(os, fmaf(ccs_f, bx_f, ox_f), fmaf(ssn_f, bx_f, oy_f)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const int p = (ib + u) * OBS_SIDE + lane;
            if (on && cls[u] < 2u) put(p, cls[u] != 0u);
            const bool mixed = on && cls[u] >= 2u;
            const unsigned mask = __ballot_sync(0xffffffffu, mixed);
            if (mask) {
                if (mixed) mq[nq + __popc(mask & ((1u << lane) - 1u))] = p;
                nq += __popc(mask);
                if (nq > LOW_QUEUE - 32) drain();
            }
        }
    }
    drain();
    if (planes) {
        const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
        for (int c = 4 + OBS_NPTS + lane; c < Kp; c += 32) p0[c] = z;
        for (int c = 4 + lane; c < 64; c += 32) {
            planes[((size_t)Mpad + b) * Kp + c] = z;
            planes[((size_t)2 * Mpad + b) * Kp + c] = z;
        }
    }
    if (lane == 0) {
        const double *g = goals + (size_t)goal_stride * q;
        const double qx = g[0] - px0, qy = g[1] - py0;          // differential_gym.py:186-187 as R^T (g - t)
        const double gbx = cs * qx + sn * qy, gby = (-sn) * qx + cs * qy;
        if (obs64) {
            double *o = obs64 + (size_t)b * (4 + OBS_NPTS);
            o[0] = gbx; o[1] = gby; o[2] = st[3]; o[3] = st[4];
        }
        if (obs32) {
            float *o = obs32 + (size_t)b * ld32;
            o[0] = (float)gbx; o[1] = (float)gby; o[2] = (float)st[3]; o[3] = (float)st[4];
            for (int k = 4 + OBS_NPTS; k < ld32; k++) o[k] = 0.f;
        }
        if (planes) {
            const float h[4] = {(float)gbx, (float)gby, (float)st[3], (float)st[4]};
            for (int k = 0; k < 4; k++) {
                __nv_bfloat16 a0, a1, a2;
                crlk_split3(h[k], a0, a1, a2);
                p0[k] = a0;
                planes[((size_t)Mpad + b) * Kp + k] = a1;
                planes[((size_t)2 * Mpad + b) * Kp + k] = a2;
            }
        }
    }
}

extern "C" int crlk_local_obs(const CrlkEnv *env, const double *states, const double *goals,
                              int32_t goal_stride, int32_t B, double *obs64, float *obs32,
                              int32_t ld32, uint8_t *near_boundary, void *stream)
{
    CRLK_ENV_OK(env);
    CRLK_REQUIRE(B >= 0 && (B == 0 || (states && goals)), "NULL array");
    CRLK_REQUIRE(goal_stride >= 2, "goal_stride < 2");
    CRLK_REQUIRE(obs32 == nullptr || ld32 >= 4 + OBS_NPTS, "ld32 < 733");
    if (B == 0) return CRLK_OK;
    if (near_boundary)
        k_local_obs<true><<<B, 256, 0, (cudaStream_t)stream>>>(make_obs_grid(), env->os, states, goals,
                                                              goal_stride, B, obs64, obs32, ld32,
                                                              near_boundary, nullptr, nullptr, nullptr, 0, 0);
    else
        k_local_obs_w<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(make_obs_grid(), env->os, states, goals,
                                                                    goal_stride, B, obs64, obs32, ld32, nullptr,
                                                                    nullptr, nullptr, 0, 0);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

// Compacted batch of the RL steer loop: row b < *count observes query list[b];
// output as float32 rows and/or bf16 operand planes (see k_local_obs).
int crlk_local_obs_compact(const CrlkEnv *env, const double *states, const double *goals,
                           int32_t goal_stride, int32_t Bmax, const int *list, const int *count,
                           float *obs32, int32_t ld32, void *planes, int32_t Mpad, int32_t Kp, void *stream)
{
    if (Bmax == 0) return CRLK_OK;
    k_local_obs_w<<<(Bmax + 7) / 8, 256, 0, (cudaStream_t)stream>>>(make_obs_grid(), env->os, states, goals,
                                                                   goal_stride, Bmax, nullptr, obs32, ld32,
                                                                   list, count, (__nv_bfloat16 *)planes, Mpad, Kp);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

// ------------------------------------------ gym step (differential_gym.py:136-178)

struct GymP { double rp[8]; double goal_radius, clearance_cap; };

// DifferentialDriveGym.step for one environment per thread: the state update
// (:139), _info (:147-166) and _reward = info_arr @ reward_param (:177-178).
// info columns follow the dict order of :148-157: goal, goal_dis, heading,
// collision, clearance, v, |w|, step.
__global__ void __launch_bounds__(128)
k_gym_step(EnvP e, ObsSet os, GymP gp, double *states, const double *__restrict__ goals, int goal_stride,
           const double *__restrict__ actions, int n_sub, int B, double *info, double *reward,
           uint8_t *near_boundary)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double s[5];
#pragma unroll
    for (int k = 0; k < 5; k++) s[k] = states[5 * (size_t)i + k];
    if (actions) motion_velocity_dev(e, s, actions[2 * (size_t)i], actions[2 * (size_t)i + 1], n_sub);
    const double gx = goals[(size_t)goal_stride * i], gy = goals[(size_t)goal_stride * i + 1];
    double dx = s[0] - gx, dy = s[1] - gy;
    double gd = sqrt(fma(dy, dy, dx * dx));                                   // np.linalg.norm (:158)
    double head = fabs(normalize_angle(atan2(gy - s[1], gx - s[0]) - s[2]));    // :161-162
    bool d, m;
    double clr = clearance_grid_thread(e, os, s[0], s[1], s[2], &d, &m);        // get_clearance: -1 on collision
    double v[8];
    v[0] = (gd <= gp.goal_radius) ? 1.0 : 0.0;
    v[1] = gd;
    v[2] = head;
    v[3] = (clr < 0.0) ? 1.0 : 0.0;                                             // not valid_state_check (:164)
    v[4] = pymin(clr, gp.clearance_cap);                                        // min(get_clearance, 1.0) (:163)
    v[5] = s[3];
    v[6] = fabs(s[4]);
    v[7] = 1.0;
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (info) info[8 * (size_t)i + k] = v[k];
        r += v[k] * gp.rp[k];
    }
    if (reward) reward[i] = r;
    if (near_boundary) near_boundary[i] = (m && !d) ? 1 : 0;
#pragma unroll
    for (int k = 0; k < 5; k++) states[5 * (size_t)i + k] = s[k];
}

extern "C" int crlk_gym_step(const CrlkEnv *env, double *states, const double *goals, int32_t goal_stride,
                             const double *actions, double step_dt, const double *reward_param_host,
                             int32_t B, double *info, double *reward, double *obs64, float *obs32,
                             int32_t ld32, uint8_t *near_boundary, void *stream)
{
    CRLK_ENV_OK(env);
    CRLK_REQUIRE(B >= 0 && (B == 0 || (states && goals)), "NULL array");
    CRLK_REQUIRE(goal_stride >= 2, "goal_stride < 2");
    CRLK_REQUIRE(!reward || reward_param_host, "reward_param required with reward");
    CRLK_REQUIRE(obs32 == nullptr || ld32 >= 4 + OBS_NPTS, "ld32 < 733");
    CRLK_REQUIRE(!actions || step_dt > 0, "step_dt must be positive");
    if (B == 0) return CRLK_OK;
    GymP gp;
    for (int k = 0; k < 8; k++) gp.rp[k] = reward_param_host ? reward_param_host[k] : 0.0;
    gp.goal_radius = 1.0;        // differential_gym.py:159
    gp.clearance_cap = 1.0;      // :163
    int n_sub = actions ? crlk_count_loop_lt(step_dt, env->p.base_dt) : 0;
    k_gym_step<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->p, env->os, gp, states, goals, goal_stride,
                                                                 actions, n_sub, B, info, reward, near_boundary);
    CRLK_LAUNCH_CHECK();
    if (obs64 || obs32) {
        k_local_obs_w<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(make_obs_grid(), env->os, states, goals,
                                                                    goal_stride, B, obs64, obs32, ld32, nullptr,
                                                                    nullptr, nullptr, 0, 0);
        CRLK_LAUNCH_CHECK();
    }
    return CRLK_OK;
}

// ------------------------------------------------------------------- DWA (K1)

// Two launch shapes of the DWA kernels (template parameter T = threads per CTA):
//   throughput: 256 threads, 3 CTAs per SM (85 registers; measured best of 2, 3, 4) -- many more
//               extensions than CTAs, the phases of co-resident CTAs interleave;
//   latency:    768 threads, 1 CTA per SM (same 85-register budget) -- few extensions per SM, the
//               launch time is the longest chain of dependent control steps, so each step gets
//               three times the threads.
#ifndef DWA_THREADS
#define DWA_THREADS 256
#endif
#ifndef DWA_MIN_BLOCKS
#define DWA_MIN_BLOCKS 2
#endif
#define DWA_WIDE_THREADS 768
// The whole-query planner kernel: the same shape as the extend kernel since the control step's cost
// and summation passes became exact (round 2: 256 x 2 measured 22.1 k queries/s against 20.3 k for
// the 128 x 4 shape that was best with the cheaper round-1 control step).
#ifndef PLAN_THREADS
#define PLAN_THREADS 256
#endif
#ifndef PLAN_MIN_BLOCKS
#define PLAN_MIN_BLOCKS 2
#endif

__global__ void k_debug_atan2(const double *__restrict__ y, const double *__restrict__ x, int n, double *out)
{
    __shared__ double tab[CRLK_ATAN_SMEM_DOUBLES];
    load_atan_table(tab, threadIdx.x, blockDim.x);
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = crlk_atan2(y[i], x[i], tab);
}

// Test hook: the rollout loop's atan2 over n argument pairs (dev f64[n] each).
extern "C" int crlk_debug_atan2(const double *y, const double *x, int32_t n, double *out, void *stream)
{
    CRLK_REQUIRE(n >= 0 && (n == 0 || (y && x && out)), "NULL array");
    if (n == 0) return CRLK_OK;
    k_debug_atan2<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(y, x, n, out);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

__global__ void k_debug_sincos(const double *__restrict__ th, int n, double *sn, double *cs)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) crlk_sincos(th[i], &sn[i], &cs[i]);
}

// Test hook: the sincos of the heading chains and footprints (differential_env.py:98-99), n angles.
extern "C" int crlk_debug_sincos(const double *th, int32_t n, double *sn, double *cs, void *stream)
{
    CRLK_REQUIRE(n >= 0 && (n == 0 || (th && sn && cs)), "NULL array");
    if (n == 0) return CRLK_OK;
    k_debug_sincos<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(th, n, sn, cs);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

__global__ void k_debug_div(const double *__restrict__ a, const double *__restrict__ b, int n, double *fast,
                            double *exact)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double y = 1.0 / b[i];
    fast[i] = div_rn_y(a[i], b[i], y);
    exact[i] = a[i] / b[i];
}

// Test hook: the cost normalisation's quotient (div_rn_y with y = 1 / b) next to the division
// operator, over n pairs (dev f64[n] each).
extern "C" int crlk_debug_div(const double *a, const double *b, int32_t n, double *fast, double *exact, void *stream)
{
    CRLK_REQUIRE(n >= 0 && (n == 0 || (a && b && fast && exact)), "NULL array");
    if (n == 0) return CRLK_OK;
    k_debug_div<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a, b, n, fast, exact);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

#define CL_CAP 1024      // capacity of the per-control-step obstacle list of the clearance term

struct DwaSmem {
    double *c0, *c1;
    double *vtab, *vlast, *c2;         // c2[i] = v_hi - v_last of v command i (speed column, dwa.py:72)
    double *wcos, *wsin, *wth, *wom;
    double *red_v, *red_s;            // [32] column-sum partials per warp (heading, speed); nearest_block scratch
    int *red_i;                       // [64] index of the per-warp minimum, then of the per-warp runner-up
    double *node;                     // [3][256] leaf sums of the three cost columns (pairwise tree)
    double *red2_v;                   // [32] third column-sum partial per warp
    double *red3_v, *red3_s;          // [32] minimum / runner-up per warp
    double *hdr;                      // [8]  v_hi v_lo w_hi w_lo of the current window
    int *ihdr;                        // [4]  nv nw
    RobotPose *pose;                  // footprint of the state the next validity test looks at
    double *atab;                     // atan coefficient table (crlk_atan2)
    float4 *clist;                    // [CL_CAP] obstacles that can matter to this control step's clearance term
    int *cl_n;                        // [2] list length (> CL_CAP: overflow, use the grid search), spare
    double *scal;                     // [16]: 4 winning v command, 5 winning w command
    int *iscal;                       // [8]: 1 winning rollout index, 2 near-tie flag
    int spos, rmax;
};


__host__ __device__ inline int dwa_spos(const DwaP &d)
{
    return d.use_clearance ? d.n_sub * d.n_outer : d.n_sub;
}
__host__ __device__ inline int dwa_rmax(const DwaP &d) { return d.nv_max * d.nw_max; }

__host__ __device__ inline size_t dwa_carve(const DwaP &d, unsigned char *base, DwaSmem *s)
{
    size_t off = 0;
    int spos = dwa_spos(d), rmax = dwa_rmax(d);
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
    size_t o_c0 = take(sizeof(double) * rmax);
    size_t o_c1 = take(d.use_clearance ? sizeof(double) * rmax : 0);
    size_t o_vtab = take(sizeof(double) * d.nv_max * spos);
    size_t o_vlast = take(sizeof(double) * d.nv_max);
    size_t o_c2n = take(sizeof(double) * d.nv_max);
    size_t o_wcos = take(sizeof(double) * d.nw_max * spos);
    size_t o_wsin = take(sizeof(double) * d.nw_max * spos);
    size_t o_wth = take(sizeof(double) * d.nw_max * d.n_outer);
    size_t o_wom = take(sizeof(double) * d.nw_max);
    size_t o_rv = take(sizeof(double) * 32);
    size_t o_rs = take(sizeof(double) * 32);
    size_t o_ri = take(sizeof(int) * 64);
    size_t o_nd = take(sizeof(double) * 768);
    size_t o_r2 = take(sizeof(double) * 32);
    size_t o_r3 = take(sizeof(double) * 64);
    size_t o_hd = take(sizeof(double) * 8);
    size_t o_ih = take(sizeof(int) * 4);
    size_t o_po = take(sizeof(RobotPose));
    size_t o_at = take(sizeof(double) * CRLK_ATAN_SMEM_DOUBLES);
    size_t o_cl = take(d.use_clearance ? sizeof(float4) * CL_CAP : 0);
    size_t o_cn = take(sizeof(int) * 4);
    size_t o_sc = take(sizeof(double) * 16);
    size_t o_is = take(sizeof(int) * 8);
    if (s) {
        s->c0 = (double *)(base + o_c0); s->c1 = (double *)(base + o_c1);
        s->vtab = (double *)(base + o_vtab); s->vlast = (double *)(base + o_vlast);
        s->c2 = (double *)(base + o_c2n);
        s->wcos = (double *)(base + o_wcos); s->wsin = (double *)(base + o_wsin);
        s->wth = (double *)(base + o_wth);
        s->wom = (double *)(base + o_wom);
        s->red_v = (double *)(base + o_rv); s->red_s = (double *)(base + o_rs);
        s->red_i = (int *)(base + o_ri);
        s->node = (double *)(base + o_nd);
        s->red2_v = (double *)(base + o_r2); s->hdr = (double *)(base + o_hd);
        s->red3_v = (double *)(base + o_r3); s->red3_s = s->red3_v + 32;
        s->ihdr = (int *)(base + o_ih); s->pose = (RobotPose *)(base + o_po);
        s->atab = (double *)(base + o_at);
        s->clist = (float4 *)(base + o_cl); s->cl_n = (int *)(base + o_cn);
        s->scal = (double *)(base + o_sc); s->iscal = (int *)(base + o_is);
        s->spos = spos; s->rmax = rmax;
    }
    return off;
}

size_t crlk_dwa_smem_bytes(const DwaP &d) { return dwa_carve(d, nullptr, nullptr); }

// np.arange(lo, hi + res, res): length and the (a0, a1, delta) triple.
__host__ __device__ inline int arange_len(double start, double stop, double step)
{
    double v = (stop - start) / step;
    double c = ceil(v);
    if (!(c > 0.0)) return 0;
    if (c > 1e6) return 1000000;
    return (int)c;
}

int crlk_make_dwa(const CrlkEnv *env, const CrlkDwaParams *in, DwaP *d)
{
    CRLK_REQUIRE(env && in && d, "NULL dwa params");
    CRLK_REQUIRE(in->v_res > 0 && in->w_res > 0 && in->dt > 0 && in->predict_time >= 0, "dwa params");
    CRLK_REQUIRE(in->skip >= 1, "skip < 1");
    d->v_res = in->v_res; d->w_res = in->w_res; d->dt = in->dt; d->predict_time = in->predict_time;
    d->g_goal = in->to_goal_cost_gain; d->g_ob = in->ob_gain; d->g_speed = in->speed_cost_gain;
    d->skip = in->skip; d->use_clearance = in->use_clearance ? 1 : 0;
    d->n_sub = crlk_count_loop_lt(in->dt, env->p.base_dt);
    d->n_outer = crlk_count_loop_le(in->predict_time, in->dt);
    // the window is at most 2*acc*dt wide (and never wider than the velocity range)
    double wv = fmin(2.0 * env->p.max_acc_v * in->dt, env->p.max_v - env->p.min_v);
    double ww = fmin(2.0 * env->p.max_acc_w * in->dt, 2.0 * env->p.max_w);
    d->nv_max = arange_len(0.0, wv + in->v_res, in->v_res) + 2;
    d->nw_max = arange_len(0.0, ww + in->w_res, in->w_res) + 2;
    CRLK_REQUIRE(d->n_sub >= 1 && d->n_outer >= 1 && d->n_sub * d->n_outer <= 4096, "dwa horizon");
    CRLK_REQUIRE((long)d->nv_max * d->nw_max <= 16384, "rollout grid too large (> 16384)");
    // the column sums reproduce numpy's pairwise tree with a butterfly over its complete top levels:
    // needs every leaf at depth Dmin or Dmin + 1 (true for all n <= 16384; checked once)
    static int tree_ok = -1;
    if (tree_ok < 0) {
        int ok = 1;
        for (int n = 1; n <= 16384 && ok; n++) {
            const int dmin = pw_min_depth(n), dmax = pw_max_depth(n);
            if (dmax - dmin > 1 || dmin > 7) ok = 0;
            // a level-Dmin node is a leaf or splits into two leaves: walk the extreme paths' sizes
            int lo = n, hi = n;
            for (int k = 0; k < dmin; k++) { lo = pw_split(lo); hi = hi - pw_split(hi); }
            if (hi > 128 && (hi - pw_split(hi) > 128)) ok = 0;
            (void)lo;
        }
        tree_ok = ok;
    }
    CRLK_REQUIRE(tree_ok == 1, "internal: pairwise-sum tree shape assumption violated");
    return CRLK_OK;
}

#ifdef CRLK_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_MARK(i) do { if (threadIdx.x == 0) { long long _t = clock64(); atomicAdd(&g_phase_cycles[i], (unsigned long long)(_t - _t0)); _t0 = _t; } } while (0)
#define PHASE_START() long long _t0 = clock64()
extern "C" int crlk_debug_phase_cycles(unsigned long long *out16, int reset)
{
    cudaMemcpyFromSymbol(out16, g_phase_cycles, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z); }
    return 0;
}
#else
#define PHASE_MARK(i) do { } while (0)
#define PHASE_START() do { } while (0)
#endif

struct DwaResult {
    double v, w;
    double cost;               // the winner's cost_sum entry (dwa.py:79-81)
    int idx, R;                // idx = -1, R = 0: empty rollout grid (state outside the velocity limits)
    int near_tie;
};

// get_clearance for one pose by one thread (used only when use_clearance = 1).
__device__ __forceinline__ double clearance_thread(const EnvP &e, const ObsSet &os, double x, double y,
                                                   double th, double bound = CRLK_CLEARANCE_CAP)
{
    return clearance_grid_thread(e, os, x, y, th, nullptr, nullptr, bound);
}

// The set of threads that runs one control step together.  GrpCta<T>: the whole CTA of T threads
// (plain __syncthreads; everything folds at compile time).  GrpPart: `n` threads of a CTA (a
// multiple of 32, warp aligned) that synchronise on their own named barrier, so that two parts of
// one CTA can run two control steps side by side, or join into one part of twice the size
// (k_extend_dwa_pair).
template <int T> struct GrpCta {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int size() const { return T; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ int sync_or(int p) const { return __syncthreads_or(p); }
};
struct GrpPart {
    int t, n, bar;
    __device__ __forceinline__ int tid() const { return t; }
    __device__ __forceinline__ int size() const { return n; }
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(n) : "memory"); }
    __device__ __forceinline__ int sync_or(int p) const
    {
        int r;
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\t"
                     "selp.s32 %0, 1, 0, q;\n\t}"
                     : "=r"(r) : "r"(p), "r"(bar), "r"(n) : "memory");
        return r;
    }
};

// DWA.control (dwa.py:45-86) by one CTA of DWA_THREADS threads.  `st` (5
// doubles) and the goal must be block-uniform.  Every thread returns the result.
// Executed control steps whose duration differs from the DWA's dt (never in the planners, which
// use 0.2 s for both): the general path, kept out of line so it does not bloat the hot kernels.
static __device__ __noinline__ void advance_state_general(const EnvP &e, const double *from, double cv, double cw,
                                                          int n_sub_step, double *s, RobotPose *pose)
{
    for (int c = 0; c < 5; c++) s[c] = from[c];
    motion_velocity_dev(e, s, cv, cw, n_sub_step);
    robot_pose_setup(e, s[0], s[1], s[2], *pose);
}

template <class G, bool CLR, class Epi>
__device__ DwaResult dwa_control_block(const G &g, const EnvP &e, const DwaP &d, const ObsSet &os,
                                       const double *st, double gx, double gy,
                                       const DwaSmem &sm, Epi epilogue)
{
    const int tid = g.tid(), T = g.size();
    const int lane = tid & 31, warp = tid >> 5;
    const double dt = e.base_dt;
    // dynamic window over d.dt (dwa.py:46, differential_env.py:107-114) and the grid sizes: one
    // thread (the two float64 divisions of arange_len are not worth 256+ copies)
    if (tid == 0) {
        double vh = pymin(st[3] + e.max_acc_v * d.dt, e.max_v);
        double vl = pymax(st[3] - e.max_acc_v * d.dt, e.min_v);
        double wh = pymin(st[4] + e.max_acc_w * d.dt, e.max_w);
        double wl = pymax(st[4] - e.max_acc_w * d.dt, -e.max_w);
        sm.hdr[0] = vh; sm.hdr[1] = vl; sm.hdr[2] = wh; sm.hdr[3] = wl;
        sm.ihdr[0] = min(arange_len(vl, vh + d.v_res, d.v_res), d.nv_max);
        sm.ihdr[1] = min(arange_len(wl, wh + d.w_res, d.w_res), d.nw_max);
        sm.cl_n[0] = 0;
    }
    const int S = d.n_sub * d.n_outer, spos = sm.spos;

    PHASE_START();
    g.sync();   // previous users of the shared tables are done; the header is visible
    PHASE_MARK(0);
    const double v_hi = sm.hdr[0], v_lo = sm.hdr[1], w_lo = sm.hdr[3];
    const int nv = sm.ihdr[0], nw = sm.ihdr[1];
    const double v1 = v_lo + d.v_res, vd = v1 - v_lo;
    const double w1 = w_lo + d.w_res, wd = w1 - w_lo;
    const int R = nv * nw;
    if (R == 0) {
        // np.arange came out empty: v or omega lies outside the vehicle limits by more than acc * dt
        // (the reference fails on the empty cost table, dwa.py:76).  Block-uniform exit.
        DwaResult res;
        res.v = res.w = res.cost = CUDART_NAN;
        res.idx = -1; res.R = 0; res.near_tie = 0;
        return res;
    }
    // velocity chains (one per v command) and heading chains (one per w command)
    for (int i = tid; i < nv; i += T) {
        double cv = arange_at(v_lo, v1, vd, i);
        double cur = st[3];
#pragma unroll 1
        for (int k = 0; k < S; k++) {
            double nxt = vel_substep(cur, cv, e.max_acc_v, e.min_v, e.max_v, dt);
            // f(cur) == cur is a fixed point of the (deterministic) sub-step: every later
            // sub-step returns the same value, so the chain can stop (bit-identical result)
            bool fixed = (nxt == cur);
            cur = nxt;
            if (k < spos) sm.vtab[i * spos + k] = cur;
            else if (fixed) break;
        }
        sm.vlast[i] = cur;
        sm.c2[i] = v_hi - cur;                     // speed cost of every rollout with this v command (dwa.py:72)
    }
    // heading chains: one task per (w command, sub-step) -- the omega / theta recurrences up to the
    // sub-step are cheap, the float64 sincos at its end is not, so the sub-steps run side by side
    for (int task = T - 1 - tid; task < nw * spos; task += T) {
        const int j = task / spos, ks = task - j * spos;
        double cw = arange_at(w_lo, w1, wd, j);
        double cur = st[4], th = st[2];
#pragma unroll 1
        for (int k = 0; k <= ks; k++) {
            cur = vel_substep(cur, cw, e.max_acc_w, -e.max_w, e.max_w, dt);
            th = normalize_angle(th + cur * dt);
        }
        double sn, cs;
        crlk_sincos(th, &sn, &cs);
        sm.wcos[ks * d.nw_max + j] = cs;          // [sub-step][w command]: neighbouring rollouts, neighbouring banks
        sm.wsin[ks * d.nw_max + j] = sn;
        if ((ks + 1) % d.n_sub == 0) sm.wth[j * d.n_outer + (ks + 1) / d.n_sub - 1] = th;
        if (ks + 1 == d.n_sub) sm.wom[j] = cur;      // omega after the first dt step
    }
    double clr0 = CRLK_CLEARANCE_CAP;
    if (CLR) {
        // Clearance term, block-cooperative part.  (1) The clearance of the current state (the shared
        // first row, dwa.py:65-67): the block gathers the obstacles within a radius guessed from the
        // previous control step's clearance and takes the minimum over them, one entry per thread --
        // valid when it is small enough that nothing outside that radius could beat it; otherwise one
        // thread runs the grid search.  (2) The obstacle list of the rollouts: a rollout pose lies
        // within `reach` of the current position and only clearances below min(clr0, 100) change a
        // rollout's minimum, so an obstacle matters only if its rectangle is closer than that + rc +
        // reach to the current position.  Each obstacle is taken from the one cell that holds its
        // closest point.
        const float reach = (float)(fmax(fabs(e.max_v), fabs(e.min_v)) * e.base_dt * S) * 1.001f + 1e-3f;
        const float p0x = (float)st[0], p0y = (float)st[1];
        const float hxm = (float)(os.ncx - 1), hym = (float)(os.ncy - 1);
        auto cellx = [&](float v) { return (int)fminf(fmaxf(floorf((v - os.gx0) * os.ginv), 0.f), hxm); };
        auto celly = [&](float v) { return (int)fminf(fmaxf(floorf((v - os.gy0) * os.ginv), 0.f), hym); };
        auto gather = [&](float rl) {
            const int bx0 = cellx(p0x - rl - 2e-3f), bx1 = cellx(p0x + rl + 2e-3f);
            const int by0 = celly(p0y - rl - 2e-3f), by1 = celly(p0y + rl + 2e-3f);
            const int bw = bx1 - bx0 + 1, ncell_box = bw * (by1 - by0 + 1);
            if ((long long)ncell_box > 64LL * T) {
                if (tid == 0) sm.cl_n[0] = CL_CAP + 1;      // open space: grid search per pose
                return;
            }
            for (int c = tid; c < ncell_box; c += T) {
                const int cy = by0 + c / bw, cx = bx0 + (c - (c / bw) * bw);
                const int s0 = __ldg(&os.cell_start[cy * os.ncx + cx]), s1 = __ldg(&os.cell_start[cy * os.ncx + cx + 1]);
                for (int k = s0; k < s1; k++) {
                    const float4 o = __ldg(&os.cell_rects[k]);
                    const float qx = fminf(fmaxf(p0x, o.x), o.z), qy = fminf(fmaxf(p0y, o.y), o.w);
                    const float dx = p0x - qx, dy = p0y - qy;
                    if (dx * dx + dy * dy > rl * rl) continue;
                    if (cellx(qx) != cx || celly(qy) != cy) continue;
                    const int pos = atomicAdd(&sm.cl_n[0], 1);
                    if (pos < CL_CAP) sm.clist[pos] = o;
                }
            }
        };
        const double prev = sm.hdr[5];                            // clearance of the previous control step's state
        const float guess = (float)((prev > 0.0 && prev < 20.0) ? prev + 0.25 : 1.0);
        const float ra = guess * 1.0001f + e.rc + 2.0f * CRLK_CULL_MARGIN;   // covers every obstacle closer than `guess` to the footprint
        gather(ra);
        g.sync();
        const int n0 = sm.cl_n[0];
        double mine = CRLK_CLEARANCE_CAP;
        if (n0 <= CL_CAP) {
            RobotPose g0;
            robot_pose_setup(e, st[0], st[1], st[2], g0);
            const float cf0 = (float)g0.c, sf0 = (float)g0.s;
            for (int k = tid; k < n0; k += T) {
                const float4 o = sm.clist[k];
                if (footprint_lb(e, cf0, sf0, p0x, p0y, o) > (float)fmax(mine, 0.0) * 1.0001f) continue;
                bool dd, mm;
                mine = fmin(mine, dis2obs_exact(e, g0, o, dd, mm));
            }
        }
        mine = warp_min_d(mine);
        if (lane == 0) sm.red_v[warp] = mine;
        g.sync();
        double c_l = (lane < T / 32) ? sm.red_v[lane] : CRLK_CLEARANCE_CAP;
        c_l = warp_min_d(c_l);
        // an obstacle outside the gathered radius is farther than ra - rc from the footprint
        const bool settled = n0 <= CL_CAP && (c_l < 0.0 || c_l <= (double)(ra - e.rc) - 4.0 * CRLK_CULL_MARGIN);
        if (!settled) {
            if (tid == T - 1) sm.hdr[4] = clearance_thread(e, os, st[0], st[1], st[2]);
            g.sync();
            c_l = sm.hdr[4];
        }
        clr0 = pymax(c_l, 0.001);
        const float rl = (float)fmin(clr0, CRLK_CLEARANCE_CAP) * 1.0001f + e.rc + reach + 2.0f * CRLK_CULL_MARGIN;
        g.sync();                                          // everyone has read cl_n / the first list
        if (tid == 0) { sm.cl_n[0] = 0; sm.hdr[5] = c_l; }
        g.sync();
        gather(rl);
    }
    g.sync();
    PHASE_MARK(1);
    const int cl_n = CLR ? sm.cl_n[0] : 0;
#ifdef CRLK_PHASE_TIMING
    if (CLR && tid == 0) {
        atomicAdd(&g_phase_cycles[10], (unsigned long long)min(cl_n, CL_CAP));
        atomicAdd(&g_phase_cycles[11], (unsigned long long)(cl_n > CL_CAP ? 1 : 0));
    }
#endif

    // rollouts: heading cost from trajectory row 1 (dwa.py:58-62); every thread also keeps the
    // running sums of its cost columns (dwa.py:76-78 divides each column by its sum)
    const int di = T / nw, dj = T - di * nw;     // r += T  <=>  (i, j) += (di, dj) with carry
    int ri = tid / nw, rj = tid - ri * nw;
    auto heading_cost = [&](int i, int j, double &x, double &y) {
        const double *vt = sm.vtab + i * spos;
        const double *wc = sm.wcos + j, *ws = sm.wsin + j;
        const int nwm = d.nw_max;
        x = st[0]; y = st[1];
        if (d.n_sub == 2) {
            // the planners' 0.2 s control step: two 0.1 s sub-steps, straight-line code (left to
            // itself the compiler unrolls the general loop 8/4/2/1-fold and spends more
            // instructions on the remainder cascade than on the two sub-steps)
            x = x + vt[0] * wc[0] * dt;
            y = y + vt[0] * ws[0] * dt;
            x = x + vt[1] * wc[nwm] * dt;
            y = y + vt[1] * ws[nwm] * dt;
        } else {
#pragma unroll 1
            for (int k = 0; k < d.n_sub; k++) {
                x = x + vt[k] * wc[k * nwm] * dt;
                y = y + vt[k] * ws[k * nwm] * dt;
            }
        }
        double ang = crlk_atan2(gy - y, gx - x, sm.atab);
        double c0 = fabs(ang - sm.wth[j * d.n_outer]);
        return pymin(c0, CRLK_TWO_PI - c0);
    };
    // Column sums are formed exactly as np.sum forms them (dwa.py:76-78; "numpy's pairwise
    // summation" in crlk_device.cuh): the tree's leaves (<= 128 elements) are added by eight strided
    // accumulators -- eight lanes here, lane k is numpy's r[k] -- and combined pairwise above.
    // Level Dmin of the tree is complete; its nodes are leaves or split once more into two leaves.
    // A "unit" is one leaf: unit u belongs to node u >> ush (ush = 1 when nodes split).
    const int Dmin = pw_min_depth(R), M = 1 << Dmin;
    int nmax = R;
    for (int b = 0; b < Dmin; b++) nmax -= pw_split(nmax);
    const int ush = nmax > 128 ? 1 : 0, U = M << ush;
    const int k8 = tid & 7, ngroups = T >> 3;
    const float inv_nw = 1.0f / (float)nw;
    // (offset, length) of unit u's leaf; length 0 past the end or for the absent half of an unsplit node
    auto unit_span = [&](int u, int &off, int &n) {
        const int gnode = u >> ush;
        off = 0; n = (u < U) ? R : 0;
        for (int b = Dmin - 1; b >= 0 && n > 0; b--) {
            const int n2 = pw_split(n);
            if ((gnode >> b) & 1) { off += n2; n -= n2; } else n = n2;
        }
        if (ush) {
            const bool split = n > 128;
            const int n2 = split ? pw_split(n) : n;
            if (u & 1) { off += n2; n = split ? n - n2 : 0; } else n = n2;
        }
    };
    if (!CLR) {
        // two independent rollouts per iteration: the FP64 dependency chains overlap
        int r = tid;
        for (; r + T < R; r += 2 * T) {
            const int i0 = ri, j0 = rj;
            ri += di; rj += dj;
            if (rj >= nw) { rj -= nw; ri++; }
            const int i1 = ri, j1 = rj;
            ri += di; rj += dj;
            if (rj >= nw) { rj -= nw; ri++; }
            double xa, ya, xb, yb;
            double ca = heading_cost(i0, j0, xa, ya);
            double cb = heading_cost(i1, j1, xb, yb);
            sm.c0[r] = ca;
            sm.c0[r + T] = cb;
        }
        if (r < R) {
            double xa, ya;
            sm.c0[r] = heading_cost(ri, rj, xa, ya);
        }
    } else {
        for (int r = tid; r < R; r += T) {
            const int i = ri, j = rj;
            ri += di; rj += dj;
            if (rj >= nw) { rj -= nw; ri++; }
            double x, y;
            sm.c0[r] = heading_cost(i, j, x, y);
            const double *vt = sm.vtab + i * spos;
            const double *wc = sm.wcos + j, *ws = sm.wsin + j;
            const int nwm = d.nw_max;
            double min_dis = pymin(clr0, 100.0);
            for (int row = 1; row <= d.n_outer; row++) {
                if (row > 1)
                    for (int k = (row - 1) * d.n_sub; k < row * d.n_sub; k++) {
                        x = x + vt[k] * wc[k * nwm] * dt;
                        y = y + vt[k] * ws[k * nwm] * dt;
                    }
                if (row % d.skip == 0) {
                    // only a clearance below the running minimum changes it (dwa.py:67-68), so the
                    // search is told to ignore everything farther than that
                    const int ks = row * d.n_sub - 1;
                    double cl = -2.0;
                    if (cl_n <= CL_CAP)
                        cl = clearance_from_list(e, sm.clist, cl_n, x, y, wc[ks * nwm], ws[ks * nwm], min_dis);
                    if (cl == -2.0) cl = clearance_thread(e, os, x, y, sm.wth[j * d.n_outer + row - 1], min_dis);
                    double dis = pymax(cl, 0.001);
                    min_dis = pymin(dis, min_dis);
                }
            }
            sm.c1[r] = 1.0 / min_dis;
        }
    }
    g.sync();
    PHASE_MARK(2);
    {
        // the summation pass over the stored columns; the speed column v_hi - v_last repeats nw times
        // per v command (dwa.py:72): element r reads c2[r / nw]
        for (int u0 = 0; u0 < U; u0 += ngroups) {             // warp-uniform trip count
            const int u = u0 + (tid >> 3);
            int off, n;
            unit_span(u, off, n);
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
            const int n8 = n - (n & 7);
            if (n >= 8) {
#pragma unroll 4
                for (int i = k8; i < n8; i += 8) {
                    const int r = off + i;
                    const int iv = __float2int_rd(((float)r + 0.5f) * inv_nw);   // r / nw, exact for r <= 16384
                    r0 += sm.c0[r];
                    r1 += CLR ? sm.c1[r] : (1.0 / 100.0);
                    r2 += sm.c2[iv];
                }
            }
            // one shuffle site for every lane of the warp (groups with n < 8 pass zeros through)
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                r0 += __shfl_xor_sync(0xffffffffu, r0, o);
                r1 += __shfl_xor_sync(0xffffffffu, r1, o);
                r2 += __shfl_xor_sync(0xffffffffu, r2, o);
            }
            for (int i = (n >= 8 ? n8 : 0); i < n; i++) {                          // the sequential tail
                const int r = off + i;
                const int iv = __float2int_rd(((float)r + 0.5f) * inv_nw);
                r0 += sm.c0[r];
                r1 += CLR ? sm.c1[r] : (1.0 / 100.0);
                r2 += sm.c2[iv];
            }
            if (k8 == 0 && u < U) { sm.node[u] = r0; sm.node[256 + u] = r1; sm.node[512 + u] = r2; }
        }
    }
    g.sync();
    double S0, S1, S2;
    {
        // the complete top of the tree: every warp reduces the M node values itself (no third barrier).
        // Node g = its leaf, or the sum of its two leaves (an absent second leaf contributed 0.0).
        auto nodeval = [&](const double *a, int gn) { return ush ? a[2 * gn] + a[2 * gn + 1] : a[gn]; };
        double t0, t1, t2;
        int steps;
        if (M >= 32) {
            const int per = M >> 5, base = lane * per;
            auto local = [&](const double *a) {
                if (per == 1) return nodeval(a, base);
                if (per == 2) return nodeval(a, base) + nodeval(a, base + 1);
                return (nodeval(a, base) + nodeval(a, base + 1)) + (nodeval(a, base + 2) + nodeval(a, base + 3));
            };
            t0 = local(sm.node); t1 = local(sm.node + 256); t2 = local(sm.node + 512);
            steps = 5;
        } else {
            const bool in = lane < M;
            t0 = in ? nodeval(sm.node, lane) : 0.0; t1 = in ? nodeval(sm.node + 256, lane) : 0.0;
            t2 = in ? nodeval(sm.node + 512, lane) : 0.0;
            steps = Dmin;
        }
        for (int sft = 0; sft < steps; sft++) {
            t0 += __shfl_xor_sync(0xffffffffu, t0, 1 << sft);
            t1 += __shfl_xor_sync(0xffffffffu, t1, 1 << sft);
            t2 += __shfl_xor_sync(0xffffffffu, t2, 1 << sft);
        }
        S0 = __shfl_sync(0xffffffffu, t0, 0);
        S1 = __shfl_sync(0xffffffffu, t1, 0);
        S2 = __shfl_sync(0xffffffffu, t2, 0);
    }
    PHASE_MARK(5);
    // weighted sum and first argmin (dwa.py:76-82) in the reference's own arithmetic: every column
    // entry divided by its column sum (correctly rounded quotient, div_rn_y), then
    // (g_goal * c0 + g_ob * c1) + g_speed * c2 with every product and sum rounded separately.
    const double y0 = (S0 != 0.0) ? 1.0 / S0 : 1.0;
    const double y1 = (S1 != 0.0) ? 1.0 / S1 : 1.0;
    const double y2 = (S2 != 0.0) ? 1.0 / S2 : 1.0;
    const bool z0 = S0 != 0.0, z1 = S1 != 0.0, z2 = S2 != 0.0;
    const double c1n_const = z1 ? (1.0 / 100.0) / S1 : (1.0 / 100.0);
    auto cost_of = [&](int r, int i) {
        const double c0 = sm.c0[r], c2 = sm.c2[i];
        const double c0n = z0 ? div_rn_y(c0, S0, y0) : c0;
        const double c1n = CLR ? (z1 ? div_rn_y(sm.c1[r], S1, y1) : sm.c1[r]) : c1n_const;
        const double c2n = z2 ? div_rn_y(c2, S2, y2) : c2;
        return (d.g_goal * c0n + d.g_ob * c1n) + d.g_speed * c2n;
    };
    // (minimum, its first index, runner-up, runner-up's index): the runner-up only feeds the near-tie report
    double best = CUDART_INF, second = CUDART_INF;
    int bi = 0x7fffffff, si = 0x7fffffff;   // bits 0-15 rollout index, bits 16-30 its v index (saves an integer division)
    {
        ri = tid / nw; rj = tid - ri * nw;
        for (int r = tid; r < R; r += T) {
            const int i = ri;
            ri += di; rj += dj;
            if (rj >= nw) { rj -= nw; ri++; }
            const double c = cost_of(r, i);
            if (c < best) { second = best; si = bi; best = c; bi = r | (i << 16); }       // r ascends: the first minimum stays
            else if (c < second) { second = c; si = r | (i << 16); }
        }
    }
    // The winner: (minimum, first index) across the warp by shuffles, across the block through shared
    // memory -- every thread then reads the per-warp results itself (no serial merge by one thread).
    double wb = best;
    int wi = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, wb, o);
        const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
        if (ob < wb || (ob == wb && (oi & 0xffff) < (wi & 0xffff))) { wb = ob; wi = oi; }
    }
    if (lane == 0) { sm.red3_v[warp] = wb; sm.red_i[warp] = wi; }
    g.sync();
    PHASE_MARK(3);
    wb = sm.red3_v[0]; wi = sm.red_i[0];
    for (int w = 1; w < (T >> 5); w++) {
        const double ob = sm.red3_v[w];
        const int oi = sm.red_i[w];
        if (ob < wb || (ob == wb && (oi & 0xffff) < (wi & 0xffff))) { wb = ob; wi = oi; }
    }
    // This thread's candidate for the runner-up (its best, or its second if its best is the winner), and
    // whether it lies within 8 ulp of the winner: only then can the decision be a near tie (below).
    const double cand = (bi == wi) ? second : best;
    const int candi = (bi == wi) ? si : bi;
    const int close = (candi != 0x7fffffff && cand - wb <= 8.0 * 2.220446049250313e-16 * fabs(wb)) ? 1 : 0;
    const int ix = wi & 0xffff;
    const int wrow = wi >> 16;
    if (tid == 0) {
        const int i = wrow, j = ix - i * nw;
        sm.scal[4] = arange_at(v_lo, v1, vd, i);
        sm.scal[5] = arange_at(w_lo, w1, wd, j);
        sm.scal[6] = wb;
        sm.iscal[1] = ix;
        PHASE_MARK(4);
        // trajectory row 1 of the winner (= motion_velocity(state, opt_v, d.dt), same operations)
        double row1[5];
        {
            const double *vt = sm.vtab + i * spos;
            const double *wc = sm.wcos + j, *ws = sm.wsin + j;
            double x = st[0], y = st[1];
#pragma unroll 1
            for (int k = 0; k < d.n_sub; k++) {
                x = x + vt[k] * wc[k * d.nw_max] * dt;
                y = y + vt[k] * ws[k * d.nw_max] * dt;
            }
            row1[0] = x; row1[1] = y; row1[2] = sm.wth[j * d.n_outer];
            row1[3] = vt[d.n_sub - 1]; row1[4] = sm.wom[j];
        }
        PHASE_MARK(6);
        // thread 0 only, before the closing barrier; also gets cos / sin of row 1's heading
        epilogue(sm.scal[4], sm.scal[5], ix, row1, sm.wcos[(d.n_sub - 1) * d.nw_max + j],
                 sm.wsin[(d.n_sub - 1) * d.nw_max + j]);
    }
    PHASE_MARK(9);
    // closing barrier, fused with the vote "is any other rollout within 8 ulp of the winner?"
    int tie = 0;
    if (g.sync_or(close)) {
        // Near-tie report.  The costs above are the reference's arithmetic applied to this build's
        // heading costs, which are the C library's bits except where the library itself misrounds
        // (7.5e-4 of its atan2 results, one ulp; crlk_crmath.h).  Such a last bit moves the heading
        // column's sum by an ulp in about one control step in a hundred, and with it every cost.  A
        // decision is reported when that could change it:
        //   * winner and runner-up have DIFFERENT heading costs and lie within 8 ulp of each other
        //     (their own last bits could decide); or
        //   * they share the heading cost bit for bit (the rollout grid's last velocity row repeats
        //     the one before it whenever the window ends at max_v: same first step, same atan2
        //     arguments, speed terms one ulp apart) and the winner changes when the column sum moves
        //     by one ulp either way (a rounding that merges the two costs, or stops merging them,
        //     hands the decision to the first index).
        // Identical rollouts (same heading AND speed cost: commands clamped to the same value) are
        // never near ties: np.argmin and this kernel both keep the first.
        // (rare path) the runner-up: lexicographic (cost, index) minimum of the candidates
        double rb = cand;
        int rix = candi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, rb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, rix, o);
            if (ob < rb || (ob == rb && (oi & 0xffff) < (rix & 0xffff))) { rb = ob; rix = oi; }
        }
        if (lane == 0) { sm.red3_s[warp] = rb; sm.red_i[32 + warp] = rix; }
        g.sync();
        if (tid == 0) {
            for (int w = 0; w < (T >> 5); w++) {
                const double ob = sm.red3_s[w];
                const int oi = sm.red_i[32 + w];
                if (w == 0 || ob < rb || (ob == rb && (oi & 0xffff) < (rix & 0xffff))) { rb = ob; rix = oi; }
            }
            const int i = wrow;
            const int sx = rix & 0xffff, s_i = rix >> 16;
            const bool same_c0 = sm.c0[sx] == sm.c0[ix] && (!CLR || sm.c1[sx] == sm.c1[ix]);
            if (!same_c0) tie = 1;
            else if (sm.c2[s_i] != sm.c2[i]) {
                // same heading (and obstacle) cost: only the common first part of the two costs moves
                const double c0 = sm.c0[ix];
                const double c1n = CLR ? (z1 ? div_rn_y(sm.c1[ix], S1, y1) : sm.c1[ix]) : c1n_const;
                const double k2w = d.g_speed * (z2 ? div_rn_y(sm.c2[i], S2, y2) : sm.c2[i]);
                const double k2r = d.g_speed * (z2 ? div_rn_y(sm.c2[s_i], S2, y2) : sm.c2[s_i]);
                for (int dir = -1; dir <= 1 && !tie; dir += 2) {
                    const double S0p = __longlong_as_double(__double_as_longlong(S0) + dir);
                    const double c0n = (S0p != 0.0) ? div_rn_y(c0, S0p, 1.0 / S0p) : c0;
                    const double head = d.g_goal * c0n + d.g_ob * c1n;
                    const double cw = head + k2w, cr = head + k2r;
                    const bool winner_stays = cw < cr || (cw == cr && ix < sx);
                    if (!winner_stays) tie = 1;
                }
            }
            sm.iscal[2] = tie;
        }
        g.sync();
        tie = sm.iscal[2];
    }
    PHASE_MARK(7);
    DwaResult res;
    res.v = sm.scal[4]; res.w = sm.scal[5]; res.cost = sm.scal[6];
    res.idx = sm.iscal[1]; res.near_tie = tie; res.R = R;
    return res;
}

template <int T, int MINB, bool CLR>
__global__ void __launch_bounds__(T, MINB)
k_dwa_control(EnvP e, DwaP d, ObsSet os,
              const double *__restrict__ states, const double *__restrict__ goals, int goal_stride,
              int B, double *opt_v, int *opt_idx, double *opt_cost, double *opt_traj,
              int *n_rollouts, uint8_t *near_tie)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DwaSmem sm;
    dwa_carve(d, smem_raw, &sm);
    __shared__ double s_state[5];
    load_atan_table(sm.atab, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) sm.hdr[5] = -1.0;                       // no previous control step yet
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < 5) s_state[threadIdx.x] = states[5 * (size_t)b + threadIdx.x];
        __syncthreads();
        double gx = goals[(size_t)goal_stride * b], gy = goals[(size_t)goal_stride * b + 1];
        DwaResult r = dwa_control_block<GrpCta<T>, CLR>(GrpCta<T>(), e, d, os, s_state, gx, gy, sm,
                                                        [](double, double, int, const double *, double, double) {});
        if (threadIdx.x == 0) {
            opt_v[2 * (size_t)b] = r.v;
            opt_v[2 * (size_t)b + 1] = r.w;
            if (opt_idx) opt_idx[b] = r.idx;
            if (opt_cost) opt_cost[b] = r.cost;
            if (n_rollouts) n_rollouts[b] = r.R;
            if (near_tie) near_tie[b] = (uint8_t)r.near_tie;
            if (opt_traj && r.R == 0) {
                double *tr = opt_traj + (size_t)b * (d.n_outer + 1) * 5;
                for (int k = 0; k < (d.n_outer + 1) * 5; k++) tr[k] = CUDART_NAN;
            } else if (opt_traj) {   // dwa.py:35-43 for the winner
                double s[5];
                for (int k = 0; k < 5; k++) s[k] = s_state[k];
                double *tr = opt_traj + (size_t)b * (d.n_outer + 1) * 5;
                for (int k = 0; k < 5; k++) tr[k] = s[k];
                for (int row = 1; row <= d.n_outer; row++) {
                    motion_velocity_dev(e, s, r.v, r.w, d.n_sub);
                    for (int k = 0; k < 5; k++) tr[row * 5 + k] = s[k];
                }
            }
        }
    }
}

// Launch-shape choice.  Measured on B200 (profiles/r1g_phases_wide.txt): the wide shape halves the
// rollout phase but every other phase of a control step gets slower, so it only wins when there
// are fewer work items than SMs; CRLK_DWA_WIDE=0/1 forces a shape.
static bool dwa_use_wide(int n_items)
{
    static const char *env = getenv("CRLK_DWA_WIDE");
    if (env) return atoi(env) != 0;
    return n_items <= 148;
}

static int set_smem_attr(const void *fn, size_t bytes)
{
    CRLK_REQUIRE(bytes <= 220 * 1024, "shared memory request exceeds 220 KB");
    if (bytes > 48 * 1024)
        CRLK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return CRLK_OK;
}

extern "C" int crlk_dwa_traj_rows(const CrlkEnv *env, const CrlkDwaParams *dwa)
{
    DwaP d;
    int rc = crlk_make_dwa(env, dwa, &d);
    return rc ? rc : d.n_outer + 1;
}

extern "C" int crlk_dwa_max_rollouts(const CrlkEnv *env, const CrlkDwaParams *dwa)
{
    DwaP d;
    int rc = crlk_make_dwa(env, dwa, &d);
    return rc ? rc : d.nv_max * d.nw_max;
}

extern "C" int crlk_dwa_control(const CrlkEnv *env, const CrlkDwaParams *dwa, const double *states,
                                const double *goals, int32_t goal_stride, int32_t B, double *opt_v,
                                int32_t *opt_idx, double *opt_cost, double *opt_traj,
                                int32_t *n_rollouts, uint8_t *near_tie, void *stream)
{
    CRLK_ENV_OK(env);
    DwaP d;
    int rc = crlk_make_dwa(env, dwa, &d);
    if (rc) return rc;
    CRLK_REQUIRE(B >= 0 && (B == 0 || (states && goals && opt_v)), "NULL array");
    CRLK_REQUIRE(goal_stride >= 2, "goal_stride < 2");
    if (B == 0) return CRLK_OK;
    size_t smem = crlk_dwa_smem_bytes(d);
    const bool wide = dwa_use_wide(B);
    auto kern = d.use_clearance ? (wide ? k_dwa_control<DWA_WIDE_THREADS, 1, true> : k_dwa_control<DWA_THREADS, DWA_MIN_BLOCKS, true>)
                                : (wide ? k_dwa_control<DWA_WIDE_THREADS, 1, false> : k_dwa_control<DWA_THREADS, DWA_MIN_BLOCKS, false>);
    rc = set_smem_attr((const void *)kern, smem);
    if (rc) return rc;
    kern<<<B, wide ? DWA_WIDE_THREADS : DWA_THREADS, smem, (cudaStream_t)stream>>>(
        env->p, d, env->os, states, goals, goal_stride, B, opt_v, opt_idx, opt_cost, opt_traj, n_rollouts, near_tie);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

// ------------------------------------------------- fused extend, DWA steering (R1)

// valid_state_check for a block-uniform pose by the whole CTA.  The footprint (*gp, shared
// memory, written by one thread before the preceding barrier) is read by every thread; each
// thread screens its share of the grid cells around the pose in FP32 and tests its survivors
// exactly in FP64.  Returns bit0 = collision, bit1 = the decision is within 1e-6 of the 0.05
// threshold (reported, not claimed).  Ends with block-wide barriers.
// The exact footprint test, out of line: it runs for the few obstacles that survive the FP32
// screens, and inlined it is a fifth of the extend kernels' code (eight float64 divisions).
static __device__ __noinline__ double dis2obs_exact_cold(const EnvP &e, const RobotPose &g, float4 o, bool &definite,
                                                         bool &marginal)
{
    return dis2obs_exact(e, g, o, definite, marginal);
}

template <class G>
__device__ int collision_block(const G &grp, const EnvP &e, const ObsSet &os, const RobotPose *gp, int *work = nullptr)
{
    const int tid = grp.tid(), T = grp.size();
    const RobotPose g = *gp;
    const double x = g.x, y = g.y;
    float px = (float)x, py = (float)y;
    float thr = e.rc + (float)CRLK_COLLISION_DIS + CRLK_CULL_MARGIN;
    float thr2 = thr * thr;
    const float cf = (float)g.c, sf = (float)g.s;
    const float cthr = (float)(CRLK_COLLISION_DIS + 2.0 * CRLK_NEAR_EPS);   // farther than this: not even marginal
    bool coll = false, definite = false, marginal = false;
    // each thread screens its share of the cell lists and tests its survivors at once
    // (a pose touches a few hundred listed obstacles at most: <= 1-2 per thread)
    int n_screened = 0;
    grid_visit_rows(os, px, py, thr, tid, T, [&](float4 r) {
        n_screened++;
        if (aabb_center_d2(r, px, py) <= thr2 && footprint_lb(e, cf, sf, px, py, r) <= cthr) {
            bool d, m;
            coll |= dis2obs_exact_cold(e, g, r, d, m) < 0.0;
            definite |= d; marginal |= m;
            if (work) atomicAdd(&work[0], 1);                 // exact FP64 footprint tests (rare)
        }
    });
    if (work && n_screened) atomicAdd(&work[1], n_screened);  // rectangles through the FP32 screens
    // barrier-fused reductions; most poses are not within the threshold band of any obstacle's
    // bounding circle, so no thread ran an exact test and one barrier settles it
    if (!grp.sync_or(coll | definite | marginal)) return 0;
    int any_coll = grp.sync_or(coll);
    int any_def = grp.sync_or(definite);
    int any_mar = grp.sync_or(marginal);
    return (any_coll ? 1 : 0) | ((any_mar && !any_def) ? 2 : 0);
}

struct ExtResult {
    int nn, status, steps, flags;
    unsigned long long rolls;
    unsigned tasks, chains;       // sincos chain tasks (nw x positions) and velocity chains (nv) over the steps
};

// RRT.steer (rrt.py:99-135) by one CTA: up to n_max_extend closed-loop control
// steps from s_state (block-uniform, shared memory; overwritten with the last
// executed state) toward (gx, gy).  path (nullable) receives one row of 5 per
// executed step; ctrl_idx (nullable) the chosen rollout per step.  on_node(k_node,
// step) is called by EVERY thread (uniformly) when a node is emitted at `step`
// with the node's state in s_state; it returns false to stop the extension
// (capacity exhausted).  Every thread returns the same result.
template <class G, bool CLR, class OnNode>
__device__ ExtResult extend_dwa_block(const G &g, const EnvP &e, const DwaP &d, const ExtP &x, const ObsSet &os,
                                      const DwaSmem &sm, double *s_state, double gx,
                                      double gy, double *path, int *ctrl_idx, OnNode on_node)
{
    ExtResult res;
    res.nn = 0; res.status = CRLK_EXT_TIMEOUT; res.steps = 0; res.flags = 0; res.rolls = 0; res.tasks = 0; res.chains = 0;
    for (int k = 1; k <= x.n_max_extend; k++) {
        // rrt.py:116-118: control, then thread 0 advances the state inside the control block's
        // closing section: s_state (nobody reads it any more) and the footprint of the new state
        // for the validity test, both visible after the block's closing barrier
        DwaResult r = dwa_control_block<G, CLR>(g, e, d, os, s_state, gx, gy, sm,
            [&](double cv, double cw, int ix, const double *row1, double cs1, double sn1) {
                double s[5];
                if (x.n_sub_step == d.n_sub) {
                    // executing the control for step_dt == dwa.dt IS the winner's trajectory row 1,
                    // and cos / sin of its heading are already in the chain tables
                    for (int c = 0; c < 5; c++) s[c] = row1[c];
                    robot_pose_from_cs(e, s[0], s[1], cs1, sn1, *sm.pose);
                } else {
                    advance_state_general(e, s_state, cv, cw, x.n_sub_step, s, sm.pose);
                }
                for (int c = 0; c < 5; c++) s_state[c] = s[c];
                if (path)
                    for (int c = 0; c < 5; c++) path[(size_t)(k - 1) * 5 + c] = s[c];
                if (ctrl_idx) ctrl_idx[k - 1] = ix;
            });
        if (r.R == 0) {
            // empty rollout grid (start state outside the velocity limits): nothing was executed
            res.status = CRLK_EXT_INVALID; res.flags |= 8;
            break;
        }
        res.flags |= r.near_tie ? 2 : 0;
        res.rolls += (unsigned long long)r.R;
        res.tasks += (unsigned)(sm.ihdr[1] * sm.spos);
        res.chains += (unsigned)sm.ihdr[0];
        res.steps = k;
        // rrt.py:121: valid_state_check on the new state
        PHASE_START();
        int col = collision_block(g, e, os, sm.pose, sm.cl_n + 2);
        PHASE_MARK(8);
        if (col & 2) res.flags |= 1;
        col &= 1;
        if (col == 1) { res.status = CRLK_EXT_COLLISION; break; }
        double ddx = s_state[0] - gx, ddy = s_state[1] - gy;
        double dist = sqrt(fma(ddy, ddy, ddx * ddx));                            // rrt.py:124
        if (dist <= x.reach_radius) {
            on_node(res.nn, k);
            res.nn++; res.status = CRLK_EXT_REACHED; break;
        }
        if (k % x.n_tree == 0) {                                                 // rrt.py:128
            bool go = on_node(res.nn, k);
            res.nn++;
            if (!go) break;
        }
    }
    return res;
}

template <int T, int MINB, bool CLR>
__global__ void __launch_bounds__(T, MINB)
k_extend_dwa(EnvP e, DwaP d, ExtP x, ObsSet os,
             const double *__restrict__ from_states, const double *__restrict__ targets, int Q,
             double *path, int *n_steps, int *status, int *node_step, int *n_nodes, int *ctrl_idx,
             uint8_t *flags, const uint8_t *__restrict__ active, int *work_counter,
             unsigned long long *stats, const int *__restrict__ order)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DwaSmem sm;
    dwa_carve(d, smem_raw, &sm);
    __shared__ double s_state[5];
    __shared__ int s_q;
    const int per = x.n_max_extend / x.n_tree + 1;
    const int tid = threadIdx.x;
    load_atan_table(sm.atab, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) sm.hdr[5] = -1.0;                       // no previous control step yet
    for (;;) {
        // dynamic work distribution: extensions finish after very different step counts
        __syncthreads();
        if (tid == 0) {
            const int w = atomicAdd(work_counter, 1);
            s_q = (w < Q && order) ? order[w] : w;
        }
        __syncthreads();
        const int q = s_q;
        if (q >= Q) break;
        if (active && !active[q]) {
            if (tid == 0) {
                n_steps[q] = 0; status[q] = CRLK_EXT_TIMEOUT; n_nodes[q] = 0;
                if (flags) flags[q] = 0;
            }
            continue;
        }
        if (tid < 5) s_state[tid] = from_states[5 * (size_t)q + tid];
        if (tid == 0) { sm.cl_n[2] = 0; sm.cl_n[3] = 0; }
        __syncthreads();
        const double gx = targets[5 * (size_t)q], gy = targets[5 * (size_t)q + 1];
        ExtResult r = extend_dwa_block<GrpCta<T>, CLR>(GrpCta<T>(), e, d, x, os, sm, s_state, gx, gy,
                                       path + (size_t)q * x.n_max_extend * 5,
                                       ctrl_idx ? ctrl_idx + (size_t)q * x.n_max_extend : nullptr,
                                       [&](int kn, int step) {
                                           if (tid == 0) node_step[(size_t)q * per + kn] = step;
                                           return true;
                                       });
        if (tid == 0) {
            n_steps[q] = r.steps; status[q] = r.status; n_nodes[q] = r.nn;
            if (flags) flags[q] = (uint8_t)r.flags;
            if (stats) {
                atomicAdd(&stats[0], 1ull);                            // extensions
                atomicAdd(&stats[1], (unsigned long long)r.steps);     // control steps
                atomicAdd(&stats[2], r.rolls);                         // rollouts scored
                atomicAdd(&stats[3], (unsigned long long)sm.cl_n[2]);  // exact FP64 footprint tests
                atomicAdd(&stats[4], (unsigned long long)sm.cl_n[3]);  // rectangles through the FP32 screens
                atomicAdd(&stats[5], (unsigned long long)r.tasks);     // sincos chain tasks
                atomicAdd(&stats[6], (unsigned long long)r.chains);    // velocity chains
            }
        }
    }
}

// ---- the same extensions by CTAs of two 256-thread parts that can join ------------------------
//
// A batch of extensions ends when its longest one does, and an extension is a chain of up to
// n_max_extend control steps that only one group of threads can walk.  One part (256 threads) per
// extension gives the best throughput while there is work for every part; the latency of a control
// step, though, is nearly halved when 512 threads share its rollouts, cost pass and column sums.
// So every SM runs ONE CTA of two parts:
//   * the CTA's first work item -- with the longest-first order, one of the batch's longest
//     extensions -- is run by both parts together from its first step;
//   * afterwards each part pulls its own items from the queue;
//   * a part that finds the queue empty offers itself to its neighbour (s_idle) and waits on the
//     join barrier; the neighbour sees the offer in the closing section of its current control step
//     and takes it at the top of the next one: the rest of that extension runs on 512 threads.
// Results do not depend on which threads computed them: the rollout grid is strided over whatever
// thread count the group has, and every reduction is order-exact (first minimum; numpy's pairwise
// sum tree).
#define PAIR_PART 256
#define PAIR_BAR_JOIN 3

struct PairXchg {
    double gx, gy;
    unsigned long long rolls;
    int k, nn, status, steps, flags, q, worker, done;
};

template <bool CLR>
__global__ void __launch_bounds__(2 * PAIR_PART, 1)
k_extend_dwa_pair(EnvP e, DwaP d, ExtP x, ObsSet os,
                  const double *__restrict__ from_states, const double *__restrict__ targets, int Q,
                  double *path, int *n_steps, int *status, int *node_step, int *n_nodes, int *ctrl_idx,
                  uint8_t *flags, const uint8_t *__restrict__ active, int *work_counter,
                  unsigned long long *stats, const int *__restrict__ order, int part_bytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_state[2][5];
    __shared__ int s_q[2], s_dec[2], s_tmp[2];
    __shared__ int s_idle;                 // 0, or 1 + the part that waits on the join barrier
    __shared__ PairXchg s_x;
    const int part = threadIdx.x / PAIR_PART, ltid = threadIdx.x - part * PAIR_PART;
    const int per = x.n_max_extend / x.n_tree + 1;
    DwaSmem sm;
    dwa_carve(d, smem_raw + (size_t)part * part_bytes, &sm);
    load_atan_table(sm.atab, ltid, PAIR_PART);
    if (ltid == 0) { sm.hdr[5] = -1.0; s_dec[part] = 0; }
    if (threadIdx.x == 0) { s_idle = 0; s_x.done = 0; }
    __syncthreads();

    GrpPart g{ltid, PAIR_PART, 1 + part};
    int mine = part;                       // whose tables / state this thread currently works on
    bool first = true, dry = false;
    for (;;) {
        int q, k0 = 1;
        double gx, gy;
        ExtResult res;
        res.nn = 0; res.status = CRLK_EXT_TIMEOUT; res.steps = 0; res.flags = 0; res.rolls = 0;
        bool skip = false;
        if (first) {
            // the CTA's first item: both parts, on part 0's tables
            first = false;
            if (threadIdx.x == 0) {
                const int w = atomicAdd(work_counter, 1);
                s_q[0] = (w < Q && order) ? order[w] : w;
            }
            __syncthreads();
            q = s_q[0];
            if (q >= Q) return;
            g = GrpPart{(int)threadIdx.x, 2 * PAIR_PART, 0};
            mine = 0;
            dwa_carve(d, smem_raw, &sm);
        } else {
            if (g.n != PAIR_PART) {        // back to two parts after a joined item
                g = GrpPart{ltid, PAIR_PART, 1 + part};
                mine = part;
                dwa_carve(d, smem_raw + (size_t)part * part_bytes, &sm);
                __syncthreads();
            }
            g.sync();
            if (ltid == 0) {
                const int w = atomicAdd(work_counter, 1);
                s_q[part] = (w < Q && order) ? order[w] : w;
            }
            g.sync();
            q = s_q[part];
            if (q >= Q) {
                // queue empty: offer this part to the neighbour, or -- if the neighbour already
                // waits -- release it; either way the CTA is finished once the barrier opens
                if (ltid == 0) s_tmp[part] = atomicCAS(&s_idle, 0, 1 + part);
                g.sync();
                const bool helper = s_tmp[part] == 0;
                if (!helper && ltid == 0) { s_x.done = 1; __threadfence_block(); }
                asm volatile("bar.sync %0, %1;" ::"r"(PAIR_BAR_JOIN), "r"(2 * PAIR_PART) : "memory");
                if (!helper || s_x.done) return;
                // joined: continue the neighbour's extension from its next control step
                mine = s_x.worker;
                dwa_carve(d, smem_raw + (size_t)mine * part_bytes, &sm);
                g = GrpPart{PAIR_PART + ltid, 2 * PAIR_PART, 0};
                q = s_x.q; k0 = s_x.k; gx = s_x.gx; gy = s_x.gy;
                res.nn = s_x.nn; res.status = s_x.status; res.steps = s_x.steps; res.flags = s_x.flags;
                res.rolls = s_x.rolls;
                dry = true;
                skip = true;               // state and targets are already in place
            }
        }
        double *st = s_state[mine];
        if (!skip) {
            if (active && !active[q]) {
                if (g.tid() == 0) {
                    n_steps[q] = 0; status[q] = CRLK_EXT_TIMEOUT; n_nodes[q] = 0;
                    if (flags) flags[q] = 0;
                }
                continue;
            }
            if (g.tid() < 5) st[g.tid()] = from_states[5 * (size_t)q + g.tid()];
            g.sync();
            gx = targets[5 * (size_t)q]; gy = targets[5 * (size_t)q + 1];
        }
        double *qpath = path + (size_t)q * x.n_max_extend * 5;
        int *qctrl = ctrl_idx ? ctrl_idx + (size_t)q * x.n_max_extend : nullptr;
        // RRT.steer (rrt.py:99-135), as extend_dwa_block, with the join test at the top of a step
        for (int k = k0; k <= x.n_max_extend; k++) {
            if (g.n == PAIR_PART && s_dec[mine]) {
                if (ltid == 0) {
                    s_x.gx = gx; s_x.gy = gy; s_x.rolls = res.rolls; s_x.k = k; s_x.nn = res.nn;
                    s_x.status = res.status; s_x.steps = res.steps; s_x.flags = res.flags; s_x.q = q;
                    s_x.worker = mine; s_x.done = 0;
                    __threadfence_block();
                }
                asm volatile("bar.sync %0, %1;" ::"r"(PAIR_BAR_JOIN), "r"(2 * PAIR_PART) : "memory");
                g = GrpPart{ltid, 2 * PAIR_PART, 0};
                dry = true;
            }
            const bool watch = g.n == PAIR_PART;
            DwaResult r = dwa_control_block<GrpPart, CLR>(g, e, d, os, st, gx, gy, sm,
                [&](double cv, double cw, int ix, const double *row1, double cs1, double sn1) {
                    double s[5];
                    if (x.n_sub_step == d.n_sub) {
                        for (int c = 0; c < 5; c++) s[c] = row1[c];
                        robot_pose_from_cs(e, s[0], s[1], cs1, sn1, *sm.pose);
                    } else {
                        advance_state_general(e, st, cv, cw, x.n_sub_step, s, sm.pose);
                    }
                    for (int c = 0; c < 5; c++) st[c] = s[c];
                    for (int c = 0; c < 5; c++) qpath[(size_t)(k - 1) * 5 + c] = s[c];
                    if (qctrl) qctrl[k - 1] = ix;
                    if (watch) s_dec[mine] = *(volatile int *)&s_idle;     // an offer from the other part?
                });
            if (r.R == 0) { res.status = CRLK_EXT_INVALID; res.flags |= 8; break; }
            res.flags |= r.near_tie ? 2 : 0;
            res.rolls += (unsigned long long)r.R;
            res.steps = k;
            int col = collision_block(g, e, os, sm.pose);                              // rrt.py:121
            if (col & 2) res.flags |= 1;
            if (col & 1) { res.status = CRLK_EXT_COLLISION; break; }
            double ddx = st[0] - gx, ddy = st[1] - gy;
            double dist = sqrt(fma(ddy, ddy, ddx * ddx));                            // rrt.py:124
            if (dist <= x.reach_radius) {
                if (g.tid() == 0) node_step[(size_t)q * per + res.nn] = k;
                res.nn++; res.status = CRLK_EXT_REACHED; break;
            }
            if (k % x.n_tree == 0) {                                                 // rrt.py:128
                if (g.tid() == 0) node_step[(size_t)q * per + res.nn] = k;
                res.nn++;
            }
        }
        if (g.tid() == 0) {
            n_steps[q] = res.steps; status[q] = res.status; n_nodes[q] = res.nn;
            if (flags) flags[q] = (uint8_t)res.flags;
            if (stats) {
                atomicAdd(&stats[0], 1ull);
                atomicAdd(&stats[1], (unsigned long long)res.steps);
                atomicAdd(&stats[2], res.rolls);
            }
        }
        if (dry) return;                   // joined because the queue was empty: nothing left to pull
    }
}

// ---------------------------------------------- one policy step of the RL extension, fused
//
// RRT_RL.steer (rrt_rl.py:40-58) per live row of a compacted batch, one CTA per row:
//   head      action = clamp(tanh(W h + b) (+ eps * noise) * scale + bias)   (net.py:145-153, ddpg.py:126-130)
//             from the last hidden activation h (float32 [rows][ld_h]) of the policy GEMMs;
//   motion    env.step's state update (differential_gym.py:139): one thread, float64;
//   validity  valid_state_check of the new state (rrt_rl.py:48) by the whole CTA (collision_block);
//   planner   path row, collision / reach / node-every-n_tree bookkeeping (rrt_rl.py:48-57); a query
//             that keeps extending takes the next free row of the NEXT step's batch;
//   observe   env._obs() of the new state (differential_gym.py:141,180-187) written straight into that
//             row of the first layer's bf16 operand planes (or float32 rows on the SIMT path), with
//             the cos / sin of the heading the motion step just computed.
// FIRST = true is the call before step 1: no head / motion / validity, only the observation of
// the start states (rrt_rl.py:34) into row b.
#define RLA_QUEUE 64
#ifndef RLA_MIN_BLOCKS
#define RLA_MIN_BLOCKS 4
#endif
template <bool FIRST>
__global__ void __launch_bounds__(256, RLA_MIN_BLOCKS)
k_rl_advance(EnvP e, ExtP x, ObsSet os, ObsGrid og, RlAdvP a)
{
    // one WARP per row, eight rows per CTA, no block-wide barrier: the float64 motion step of a row
    // runs on one lane while the other warps of the SM work on their rows
    __shared__ RobotPose s_pose[8];
    __shared__ int s_queue[8][RLA_QUEUE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x * 8 + warp;
    if (b >= a.n || b >= *a.count_in) return;
    const int q = a.list_in[b];
    const int per = x.n_max_extend / x.n_tree + 1;
    int slot = b;
    double st0, st1, st3, st4, cs, sn;       // state x, y, v, omega and the heading's cos / sin (lane 0)
    if (!FIRST) {
        // ---- head: two dot products over K
        const float *hr = a.h + (size_t)b * a.ld_h;
        float s0 = 0.f, s1 = 0.f;
        for (int k = lane; k < a.K; k += 32) {
            const float v = hr[k];
            s0 = fmaf(v, __ldg(&a.w[k]), s0);
            s1 = fmaf(v, __ldg(&a.w[a.K + k]), s1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        RobotPose &g = s_pose[warp];
        if (lane == 0) {
            float act[2];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                float z = tanhf((c == 0 ? s0 : s1) + __ldg(&a.bias[c]));
                if (a.noise) {
                    // draws past the end of a query's stream count as zero (the host sizes the stream)
                    const long long off = (long long)a.noise_off + (a.noise_cursor ? a.noise_cursor[q] : 0);
                    if (off < a.noise_rows) z = z + a.noise[(size_t)q * a.noise_stride + off * 2 + c] * a.eps;
                }
                z = z * a.scale[c] + a.hbias[c];
                act[c] = fminf(fmaxf(z, a.low[c]), a.high[c]);
            }
            // ---- motion (differential_gym.py:139): the float32 action promoted to float64
            double s[5];
            for (int c = 0; c < 5; c++) s[c] = a.state[5 * (size_t)q + c];
            cs = 1.0; sn = 0.0;
            for (int kk = 0; kk < x.n_sub_step; kk++) {
                const double u0 = vel_substep(s[3], (double)act[0], e.max_acc_v, e.min_v, e.max_v, e.base_dt);
                const double u1 = vel_substep(s[4], (double)act[1], e.max_acc_w, -e.max_w, e.max_w, e.base_dt);
                const double th = normalize_angle(s[2] + u1 * e.base_dt);
                crlk_sincos(th, &sn, &cs);
                s[2] = th;
                s[0] = s[0] + u0 * cs * e.base_dt;
                s[1] = s[1] + u0 * sn * e.base_dt;
                s[3] = u0; s[4] = u1;
            }
            if (x.n_sub_step == 0) crlk_sincos(s[2], &sn, &cs);
            for (int c = 0; c < 5; c++) {
                a.state[5 * (size_t)q + c] = s[c];
                a.path[((size_t)q * x.n_max_extend + (a.k - 1)) * 5 + c] = s[c];
            }
            st0 = s[0]; st1 = s[1]; st3 = s[3]; st4 = s[4];
            robot_pose_from_cs(e, s[0], s[1], cs, sn, g);
            if (a.actions) {
                a.actions[((size_t)q * x.n_max_extend + (a.k - 1)) * 2] = act[0];
                a.actions[((size_t)q * x.n_max_extend + (a.k - 1)) * 2 + 1] = act[1];
            }
            a.n_steps[q] = a.k;
        }
        __syncwarp();
        // ---- validity of the new state (rrt_rl.py:48): the warp screens the grid cells around the
        // footprint in FP32 and tests the survivors exactly in FP64
        const float px = (float)g.x, py = (float)g.y;
        const float thr = e.rc + (float)CRLK_COLLISION_DIS + CRLK_CULL_MARGIN, thr2 = thr * thr;
        const float cf = (float)g.c, sf = (float)g.s;
        const float lbthr = (float)(CRLK_COLLISION_DIS + 2.0 * CRLK_NEAR_EPS);
        int *queue = s_queue[warp];
        const float4 *__restrict__ rects = os.rects;
        bool coll = false, definite = false, marginal = false;
        int count = 0;
        auto drain = [&]() {
            for (int b0 = 0; b0 < count; b0 += 32)
                if (b0 + lane < count) {
                    bool dd, mm;
                    coll |= dis2obs_exact_cold(e, g, __ldg(&rects[queue[b0 + lane]]), dd, mm) < 0.0;
                    definite |= dd; marginal |= mm;
                }
            count = 0;
        };
        grid_visit(os, px, py, thr, lane, 32, [&](int o, bool in) {
            const float4 rr = __ldg(&rects[in ? o : 0]);
            const bool keep = in && aabb_center_d2(rr, px, py) <= thr2 && footprint_lb(e, cf, sf, px, py, rr) <= lbthr;
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (mask) {
                const int pos = count + __popc(mask & ((1u << lane) - 1u));
                if (keep) queue[pos] = o;
                count += __popc(mask);
                __syncwarp();
                if (count > RLA_QUEUE - 32) { drain(); __syncwarp(); }
            }
        });
        __syncwarp();
        drain();
        coll = __any_sync(0xffffffffu, coll);
        definite = __any_sync(0xffffffffu, definite);
        marginal = __any_sync(0xffffffffu, marginal);
        int sl = -1;
        if (lane == 0) {
            if (marginal && !definite && a.flags) a.flags[q] |= 1;
            const double gx = a.targets[5 * (size_t)q], gy = a.targets[5 * (size_t)q + 1];
            const double ddx = st0 - gx, ddy = st1 - gy;
            const double dist = sqrt(fma(ddy, ddy, ddx * ddx));
            bool go_on = true;
            int nn = a.nn[q];
            if (coll) {
                a.status[q] = CRLK_EXT_COLLISION; go_on = false;
            } else if (dist <= x.reach_radius) {
                a.node_step[(size_t)q * per + nn] = a.k;
                nn++; a.status[q] = CRLK_EXT_REACHED; go_on = false;
            } else if (a.k % x.n_tree == 0) {
                a.node_step[(size_t)q * per + nn] = a.k;
                nn++;
            }
            a.nn[q] = nn; a.n_nodes[q] = nn;
            if (go_on && a.k < x.n_max_extend) {
                sl = atomicAdd(a.count_out, 1);
                a.list_out[sl] = q;
            }
        }
        slot = __shfl_sync(0xffffffffu, sl, 0);
        if (slot < 0) return;
    } else if (lane == 0) {
        st0 = a.state[5 * (size_t)q]; st1 = a.state[5 * (size_t)q + 1];
        st3 = a.state[5 * (size_t)q + 3]; st4 = a.state[5 * (size_t)q + 4];
        crlk_sincos(a.state[5 * (size_t)q + 2], &sn, &cs);
    }
    // ---- observation of the (new) state into row `slot`: lane = column of the 27 x 27 sample grid,
    // one grid row per iteration (no index division, the lane's lateral offset is loop invariant).
    // Most points are settled by the occupancy raster; the few that fall into partly covered raster
    // cells are queued and decided afterwards with all lanes busy (decided in place they would stall
    // the whole warp on one or two lanes: 58 % of the iterations had such a point).
    const double px0 = __shfl_sync(0xffffffffu, st0, 0), py0 = __shfl_sync(0xffffffffu, st1, 0);
    const double ccs = __shfl_sync(0xffffffffu, cs, 0), ssn = __shfl_sync(0xffffffffu, sn, 0);
    __nv_bfloat16 *planes = (__nv_bfloat16 *)a.planes_v;
    __nv_bfloat16 *p0 = planes ? planes + (size_t)slot * a.Kp : nullptr;
    auto put = [&](int p, bool occ) {
        const float v = occ ? 0.f : 1.f;
        if (a.obs32) a.obs32[(size_t)slot * a.ld32 + 4 + p] = v;
        if (p0) p0[4 + p] = __float2bfloat16_rn(v);
    };
    {
        const bool on = lane < OBS_SIDE;
        const double by = arange_at(og.lr0, og.lr1, og.lrd, on ? lane : 0);
        int *mq = s_queue[warp];                     // the collision queue is idle by now
        int nq = 0;
        auto point = [&](int ib, double byv_ns, double byv_cs, double &px, double &py) {
            const double bx = arange_at(og.ba0, og.ba1, og.bad, ib);
            px = __dadd_rn(__dadd_rn(__dmul_rn(ccs, bx), byv_ns), px0);    // body_to_world (differential_gym.py:83)
            py = __dadd_rn(__dadd_rn(__dmul_rn(ssn, bx), byv_cs), py0);
        };
        auto drain = [&]() {
            __syncwarp();
            for (int i = lane; i < nq; i += 32) {
                const int p = mq[i], ib = p / OBS_SIDE, il = p - ib * OBS_SIDE;
                const double byq = arange_at(og.lr0, og.lr1, og.lrd, il);
                double px, py;
                point(ib, __dmul_rn(-ssn, byq), __dmul_rn(ccs, byq), px, py);
                bool nearp = false;
                put(p, local_map_occupied<true>(os, px, py, nearp));     // the exact cell walk
            }
            nq = 0;
            __syncwarp();
        };
        static_assert(OBS_SIDE % 3 == 0, "three grid rows per iteration");
        // The raster screen runs on a float32 evaluation of body_to_world (its margin covers the float32
        // error, crlk_device.cuh); only the queued points are evaluated exactly, in float64, by drain().
        // Lane i keeps the longitudinal offset of grid row i; a row's value comes by shuffle.
        const float ccs_f = (float)ccs, ssn_f = (float)ssn, by_f = (float)by;
        const float ox_f = fmaf(-ssn_f, by_f, (float)px0), oy_f = fmaf(ccs_f, by_f, (float)py0);
        const float my_bx_f = (float)arange_at(og.ba0, og.ba1, og.bad, on ? lane : 0);
        for (int ib = 0; ib < OBS_SIDE; ib += 3) {
            // three independent raster loads in flight before the first vote (a vote is a
            // convergence point: the compiler does not move loads across it)
            unsigned cls[3];
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const float bx_f = __shfl_sync(0xffffffffu, my_bx_f, ib + u);
                cls[u] = on ? local_map_raster_f(os, fmaf(ccs_f, bx_f, ox_f), fmaf(ssn_f, bx_f, oy_f)) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int p = (ib + u) * OBS_SIDE + lane;
                if (on && cls[u] < 2u) put(p, cls[u] != 0u);
                const bool mixed = on && cls[u] >= 2u;
                const unsigned mask = __ballot_sync(0xffffffffu, mixed);
                if (mask) {
                    if (mixed) mq[nq + __popc(mask & ((1u << lane) - 1u))] = p;
                    nq += __popc(mask);
                    if (nq > RLA_QUEUE - 32) drain();
                }
            }
        }
        drain();
    }
    if (planes) {
        const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
        for (int c = 4 + OBS_NPTS + lane; c < a.Kp; c += 32) p0[c] = z;
        for (int c = 4 + lane; c < 64; c += 32) {
            planes[((size_t)a.Mpad + slot) * a.Kp + c] = z;
            planes[((size_t)2 * a.Mpad + slot) * a.Kp + c] = z;
        }
    }
    if (lane == 0) {
        const double *gl = a.targets + 5 * (size_t)q;
        const double qx = gl[0] - st0, qy = gl[1] - st1;      // differential_gym.py:186-187 as R^T (g - t)
        const float hv[4] = {(float)(cs * qx + sn * qy), (float)((-sn) * qx + cs * qy), (float)st3, (float)st4};
        if (a.obs32) {
            float *o = a.obs32 + (size_t)slot * a.ld32;
            for (int c = 0; c < 4; c++) o[c] = hv[c];
            for (int c = 4 + OBS_NPTS; c < a.ld32; c++) o[c] = 0.f;
        }
        if (planes) {
            for (int c = 0; c < 4; c++) {
                __nv_bfloat16 a0, a1, a2;
                crlk_split3(hv[c], a0, a1, a2);
                p0[c] = a0;
                planes[((size_t)a.Mpad + slot) * a.Kp + c] = a1;
                planes[((size_t)2 * a.Mpad + slot) * a.Kp + c] = a2;
            }
        }
    }
}

// Launcher (crlk_nets.cu fills the parameter block; RlAdvP is repeated there as an opaque blob).
int crlk_rl_advance(const CrlkEnv *env, const ExtP *x, const void *params, int first, void *stream)
{
    const RlAdvP *a = (const RlAdvP *)params;
    if (a->n <= 0) return CRLK_OK;
    const int grid = (a->n + 7) / 8;
    if (first)
        k_rl_advance<true><<<grid, 256, 0, (cudaStream_t)stream>>>(env->p, *x, env->os, make_obs_grid(), *a);
    else
        k_rl_advance<false><<<grid, 256, 0, (cudaStream_t)stream>>>(env->p, *x, env->os, make_obs_grid(), *a);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

int crlk_make_ext(const CrlkEnv *env, const CrlkExtendParams *in, ExtP *x)
{
    CRLK_REQUIRE(in != nullptr, "extend params NULL");
    CRLK_REQUIRE(in->n_max_extend >= 1 && in->n_tree >= 1 && in->step_dt > 0, "extend params");
    x->n_max_extend = in->n_max_extend;
    x->n_tree = in->n_tree;
    x->n_sub_step = crlk_count_loop_lt(in->step_dt, env->p.base_dt);
    x->reach_radius = in->reach_radius;
    return CRLK_OK;
}

// Work counters of the persistent kernels: a ring of 4-byte slots per device, one fresh slot per
// launch (zeroed on the launch's stream), so launches on different streams never share a counter.
#define WORK_RING 4096
static int *g_work_ring[64] = {nullptr};
static unsigned g_work_next[64] = {0};

static int get_work_counter(int device, int **out, cudaStream_t stream)
{
    CRLK_REQUIRE(device >= 0 && device < 64, "device index");
    if (!g_work_ring[device]) {
        int *p = nullptr;
        CRLK_CUDA(cudaMalloc(&p, sizeof(int) * WORK_RING));
        int *expected = nullptr;
        if (!__atomic_compare_exchange_n(&g_work_ring[device], &expected, p, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE))
            cudaFree(p);                   // another thread installed the ring first
    }
    unsigned slot = __atomic_fetch_add(&g_work_next[device], 1u, __ATOMIC_RELAXED) % WORK_RING;
    int *c = g_work_ring[device] + slot;
    CRLK_CUDA(cudaMemsetAsync(c, 0, sizeof(int), stream));
    *out = c;
    return CRLK_OK;
}

// Longest-processing-time-first order of a batch of extensions.  An extension toward a target d
// metres away needs at least (d - reach) / (v_max * dt) control steps unless it collides first, and
// one CTA runs its steps one after the other: the launch cannot end before its longest extension
// does, so those must start first.  One CTA buckets the queries by distance (1/8 m bins, descending);
// the order inside a bin is whatever the atomics give (scheduling only: results do not depend on it).
#define LPT_BINS 256
__global__ void __launch_bounds__(1024)
k_lpt_order(const double *__restrict__ from_states, const double *__restrict__ targets,
            const uint8_t *__restrict__ active, int Q, int *order)
{
    __shared__ int s_bin[LPT_BINS];
    __shared__ int s_warp[32];
    for (int b = threadIdx.x; b < LPT_BINS; b += blockDim.x) s_bin[b] = 0;
    __syncthreads();
    auto key = [&](int q) {
        if (active && !active[q]) return LPT_BINS - 1;                       // inactive pairs go last
        const float dx = (float)(from_states[5 * (size_t)q] - targets[5 * (size_t)q]);
        const float dy = (float)(from_states[5 * (size_t)q + 1] - targets[5 * (size_t)q + 1]);
        const int k = (int)(sqrtf(dx * dx + dy * dy) * 8.0f);
        return LPT_BINS - 2 - min(max(k, 0), LPT_BINS - 2);                  // far targets -> low bins -> first
    };
    for (int q = threadIdx.x; q < Q; q += blockDim.x) atomicAdd(&s_bin[key(q)], 1);
    __syncthreads();
    // exclusive prefix over the 256 bins by the first 256 threads
    int v = (threadIdx.x < LPT_BINS) ? s_bin[threadIdx.x] : 0, inc = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (threadIdx.x < LPT_BINS) {
        int base = 0;
        for (int w = 0; w < warp; w++) base += s_warp[w];
        s_bin[threadIdx.x] = base + inc - v;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < Q; q += blockDim.x) order[atomicAdd(&s_bin[key(q)], 1)] = q;
}

// Scratch for the order arrays: a ring of ints per device, one slice per launch (like the work counters).
#define ORDER_RING (1 << 22)
static int *g_order_ring[64] = {nullptr};
static unsigned long long g_order_next[64] = {0};

static int get_order_slice(int device, int n, int **out)
{
    *out = nullptr;
    if (device < 0 || device >= 64 || n > ORDER_RING) return CRLK_OK;       // no ordering: plain queue order
    if (!g_order_ring[device]) {
        int *p = nullptr;
        CRLK_CUDA(cudaMalloc(&p, sizeof(int) * (size_t)ORDER_RING));
        int *expected = nullptr;
        if (!__atomic_compare_exchange_n(&g_order_ring[device], &expected, p, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE))
            cudaFree(p);
    }
    for (;;) {
        unsigned long long cur = __atomic_load_n(&g_order_next[device], __ATOMIC_RELAXED);
        unsigned long long off = cur % ORDER_RING;
        unsigned long long start = (off + (unsigned long long)n <= ORDER_RING) ? cur : cur + (ORDER_RING - off);
        if (__atomic_compare_exchange_n(&g_order_next[device], &cur, start + (unsigned long long)n, false,
                                        __ATOMIC_ACQ_REL, __ATOMIC_RELAXED)) {
            *out = g_order_ring[device] + (start % ORDER_RING);
            return CRLK_OK;
        }
    }
}

extern "C" int crlk_extend_dwa(const CrlkEnv *env, const CrlkDwaParams *dwa, const CrlkExtendParams *ext,
                               const double *from_states, const double *targets, int32_t Q,
                               double *path, int32_t *n_steps, int32_t *status, int32_t *node_step,
                               int32_t *n_nodes, int32_t *ctrl_idx, uint8_t *flags,
                               const uint8_t *active, uint64_t *stats, void *stream)
{
    CRLK_ENV_OK(env);
    DwaP d;
    ExtP x;
    int rc = crlk_make_dwa(env, dwa, &d);
    if (rc) return rc;
    rc = crlk_make_ext(env, ext, &x);
    if (rc) return rc;
    CRLK_REQUIRE(Q >= 0, "Q < 0");
    if (Q == 0) return CRLK_OK;
    CRLK_REQUIRE(from_states && targets && path && n_steps && status && node_step && n_nodes,
                 "NULL array");
    size_t smem = crlk_dwa_smem_bytes(d);
    const bool wide = dwa_use_wide(Q);
    int dev = env->device, sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const char *pair_env = getenv("CRLK_DWA_PAIR");
    const size_t part_bytes = (smem + 127) & ~(size_t)127;
    if (!wide && pair_env && atoi(pair_env) != 0 && 2 * part_bytes <= 200 * 1024) {
        // more work items than SMs: CTAs of two joinable parts, longest extensions first
        auto pk = d.use_clearance ? k_extend_dwa_pair<true> : k_extend_dwa_pair<false>;
        rc = set_smem_attr((const void *)pk, 2 * part_bytes);
        if (rc) return rc;
        int grid = sms < Q ? sms : Q;
        int *counter = nullptr, *order = nullptr;
        rc = get_work_counter(dev, &counter, (cudaStream_t)stream);
        if (rc) return rc;
        rc = get_order_slice(dev, Q, &order);
        if (rc) return rc;
        if (order) {
            k_lpt_order<<<1, 1024, 0, (cudaStream_t)stream>>>(from_states, targets, active, Q, order);
            CRLK_LAUNCH_CHECK();
        }
        pk<<<grid, 2 * PAIR_PART, 2 * part_bytes, (cudaStream_t)stream>>>(
            env->p, d, x, env->os, from_states, targets, Q, path, n_steps, status, node_step,
            n_nodes, ctrl_idx, flags, active, counter, (unsigned long long *)stats, order, (int)part_bytes);
        CRLK_LAUNCH_CHECK();
        return CRLK_OK;
    }
    auto kern = d.use_clearance ? (wide ? k_extend_dwa<DWA_WIDE_THREADS, 1, true> : k_extend_dwa<DWA_THREADS, DWA_MIN_BLOCKS, true>)
                                : (wide ? k_extend_dwa<DWA_WIDE_THREADS, 1, false> : k_extend_dwa<DWA_THREADS, DWA_MIN_BLOCKS, false>);
    const int threads = wide ? DWA_WIDE_THREADS : DWA_THREADS;
    rc = set_smem_attr((const void *)kern, smem);
    if (rc) return rc;
    CRLK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = sms * per_sm;
    if (grid > Q) grid = Q;
    int *counter = nullptr;
    rc = get_work_counter(dev, &counter, (cudaStream_t)stream);
    if (rc) return rc;
    int *order = nullptr;
    static const char *no_lpt = getenv("CRLK_NO_LPT");
    if (Q > grid && !no_lpt) {                       // with a CTA per pair there is nothing to order
        rc = get_order_slice(dev, Q, &order);
        if (rc) return rc;
        if (order) {
            k_lpt_order<<<1, 1024, 0, (cudaStream_t)stream>>>(from_states, targets, active, Q, order);
            CRLK_LAUNCH_CHECK();
        }
    }
    kern<<<grid, threads, smem, (cudaStream_t)stream>>>(
        env->p, d, x, env->os, from_states, targets, Q, path, n_steps, status, node_step,
        n_nodes, ctrl_idx, flags, active, counter, (unsigned long long *)stats, order);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

// --------------------------- whole-query planning on the device, DWA steering (R12)

struct PlanP {
    int max_iter, S;
    double d_min, d_max, goal_radius, retry_radius;
    unsigned long long seed;
    double lo[5], hi[5];
    int goal_rate;
};

// Counter-based uniform in [0, 1): splitmix64 finaliser over (seed, query, draw, component).
__host__ __device__ inline double crlk_sample_u01(unsigned long long seed, unsigned long long q, unsigned long long i,
                                                  unsigned long long c)
{
    unsigned long long z = seed ^ (q * 0x100000001B3ull) ^ (i * 0x9E3779B97F4A7C15ull) ^ ((c + 1ull) * 0xD6E8FEB86659FD93ull);
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * 1.1102230246251565e-16;      // 2^-53
}

// RRT.sample (rrt.py:146-154): draw i of query q -> component c of the sampled state.
__device__ __forceinline__ double plan_sample(const PlanP &pp, const double *goal, int q, int i, int c)
{
    const double ub = crlk_sample_u01(pp.seed, (unsigned long long)q, (unsigned long long)i, 5ull);
    if ((int)floor(ub * 101.0) <= pp.goal_rate) return goal[c];
    const double u = crlk_sample_u01(pp.seed, (unsigned long long)q, (unsigned long long)i, (unsigned long long)c);
    return pp.lo[c] + u * (pp.hi[c] - pp.lo[c]);
}

__global__ void k_sample_stream(PlanP pp, const double *__restrict__ goals, int Q, int S, double *out)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)Q * S * 5) return;
    const int c = (int)(t % 5);
    const long long r = t / 5;
    const int i = (int)(r % S), q = (int)(r / S);
    out[t] = plan_sample(pp, goals + 5 * (size_t)q, q, i, c);
}

static void fill_planp(const CrlkPlanParams *plan, int S, PlanP *pp)
{
    pp->max_iter = plan->max_iter; pp->S = S;
    pp->d_min = plan->d_min; pp->d_max = plan->d_max;
    pp->goal_radius = plan->goal_radius; pp->retry_radius = plan->retry_radius;
    pp->seed = plan->seed; pp->goal_rate = plan->goal_sample_rate;
    for (int c = 0; c < 5; c++) { pp->lo[c] = plan->state_lo[c]; pp->hi[c] = plan->state_hi[c]; }
}

extern "C" int crlk_sample_stream(const CrlkPlanParams *plan, const double *goals, int32_t Q, int32_t S,
                                  double *out, void *stream)
{
    CRLK_REQUIRE(plan != nullptr && Q >= 0 && S >= 0, "plan / shape");
    if (Q == 0 || S == 0) return CRLK_OK;
    CRLK_REQUIRE(goals && out, "NULL array");
    PlanP pp;
    fill_planp(plan, S, &pp);
    const long long n = (long long)Q * S * 5;
    k_sample_stream<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pp, goals, Q, S, out);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

// RRT.choose_parent (rrt.py:156-161) over one tree by the whole CTA: exact
// float64 distances, first minimum.  Every thread returns the same (idx, dist).
template <int T>
__device__ __forceinline__ void nearest_block(const double *xs, const double *ys, int lo, int n, double tx,
                                              double ty, double *red_v, int *red_i, int &idx, double &dist)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double bd = CUDART_INF;
    int bi = 0x7fffffff;
    for (int j = lo + tid; j < n; j += T) {
        double dx = xs[j] - tx, dy = ys[j] - ty;
        double dd = sqrt(fma(dy, dy, dx * dx));          // np.linalg.norm of a 2-vector
        if (dd < bd) { bd = dd; bi = j; }
    }
    warp_argmin(bd, bi);
    __syncthreads();                                     // earlier readers of red_v / red_i are done
    if (lane == 0) { red_v[warp] = bd; red_i[warp] = bi; }
    __syncthreads();
    bd = red_v[0]; bi = red_i[0];
#pragma unroll
    for (int w = 1; w < T / 32; w++) {
        double ob = red_v[w];
        int oi = red_i[w];
        if (ob < bd || (ob == bd && oi < bi)) { bd = ob; bi = oi; }
    }
    idx = bi; dist = bd;
}

// RRT.planning (rrt.py:51-89) with DWA steering, one CTA per query from the
// first sample to the last node: sample validity (rrt.py:66), nearest node
// (:68), acceptance band (:69), extension (:72), node append with path segments
// (:125-134), goal test (:75) and the steer-to-goal retry (:78-83).  Persistent
// CTAs pull queries from an atomic counter (queries differ widely in length).
// The tree is the same SoA forest crlk_rrt_append maintains; sampling stays a
// pre-drawn per-query stream (the reference's RNG order is host state).
template <int T, int MINB, bool CLR>
__global__ void __launch_bounds__(T, MINB)
k_plan_dwa(EnvP e, DwaP d, ExtP x, ObsSet os, CrlkForestView f, PlanP pp,
           const double *__restrict__ goals, const double *__restrict__ streams, int Q, int *iters_out,
           int *cursor_out, uint8_t *reached_out, uint8_t *flags_out, int *overflow, int *work_counter,
           unsigned long long *stats, const int *__restrict__ order)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DwaSmem sm;
    dwa_carve(d, smem_raw, &sm);
    __shared__ double s_state[5];
    __shared__ int s_q;
    const int tid = threadIdx.x;
    load_atan_table(sm.atab, threadIdx.x, blockDim.x);
    if (threadIdx.x == 0) sm.hdr[5] = -1.0;                       // no previous control step yet
    for (;;) {
        __syncthreads();
        if (tid == 0) s_q = atomicAdd(work_counter, 1);
        __syncthreads();
        if (s_q >= Q) break;
        const int q = order ? order[s_q] : s_q;       // work item -> query (longest-expected first)
        const size_t tb = (size_t)f.stride * q;
        double *tx64 = f.x64 + tb, *ty64 = f.y64 + tb;
        int n = f.n_nodes[q];
        int pool_used = f.pool ? f.pool_used[q] : 0;
        const double gx = goals[5 * (size_t)q], gy = goals[5 * (size_t)q + 1];
        int iters = 0, cur = 0, fl = 0;
        bool reached = false, over = false;
        unsigned long long n_ext = 0, n_ctl = 0, rolls = 0;

        // one extension from node `from` toward (tx, ty); emitted nodes go straight into the tree
        auto steer = [&](int from, double tx, double ty) {
            if (f.pool && (long long)pool_used + x.n_max_extend > f.pool_cap) { over = true; return; }
            __syncthreads();
            if (tid < 5) s_state[tid] = f.states[(tb + from) * 5 + tid];
            __syncthreads();
            double *prow = f.pool ? f.pool + ((size_t)f.pool_cap * q + pool_used) * 5 : nullptr;
            int prev_step = 0;
            ExtResult r = extend_dwa_block<GrpCta<T>, CLR>(GrpCta<T>(), e, d, x, os, sm, s_state, tx, ty, prow, nullptr,
                [&](int kn, int step) {
                    if (n >= f.stride) { over = true; return false; }
                    if (tid < 5) f.states[(tb + n) * 5 + tid] = s_state[tid];
                    if (tid == 0) {
                        tx64[n] = s_state[0]; ty64[n] = s_state[1];
                        if (f.x32) { f.x32[tb + n] = (float)s_state[0]; f.y32[tb + n] = (float)s_state[1]; }
                        f.parent[tb + n] = (kn == 0) ? from : n - 1;
                        if (f.pool) {
                            f.seg_off[tb + n] = pool_used + prev_step;
                            f.seg_len[tb + n] = step - prev_step;
                            // Node.path of a steer's first node lacks the parent state (rrt.py:102-104 vs :133)
                            f.seg_flag[tb + n] = (kn == 0) ? 0 : 1;
                        }
                    }
                    n++; prev_step = step;
                    return true;
                });
            pool_used += prev_step;       // rows past the last emitted node are dropped (rrt.py:121-123)
            fl |= r.flags;
            n_ext++; n_ctl += (unsigned long long)r.steps; rolls += r.rolls;
        };
        // np.linalg.norm(node_list[-1].state[:2] - goal.state[:2]) <= goal_radius
        auto last_at_goal = [&]() {
            __syncthreads();              // the node rows written by threads 0-4 are visible
            double dx = tx64[n - 1] - gx, dy = ty64[n - 1] - gy;
            return sqrt(fma(dy, dy, dx * dx)) <= pp.goal_radius;
        };

        while (iters < pp.max_iter && cur < pp.S && !over) {
            double sx, sy, sth;
            if (streams) {
                const double *sp = streams + ((size_t)q * pp.S + cur) * 5;
                sx = sp[0]; sy = sp[1]; sth = sp[2];
            } else {                            // drawn here (rrt.py:146-154): only x, y, theta are ever used
                const double *gl = goals + 5 * (size_t)q;
                sx = plan_sample(pp, gl, q, cur, 0); sy = plan_sample(pp, gl, q, cur, 1);
                sth = plan_sample(pp, gl, q, cur, 2);
            }
            cur++;
            if (tid == 0) robot_pose_setup(e, sx, sy, sth, *sm.pose);
            __syncthreads();
            int col = collision_block(GrpCta<T>(), e, os, sm.pose);                   // rrt.py:66
            if (col & 2) fl |= 1;
            if (col & 1) continue;
            int idx;
            double dist;
            nearest_block<T>(tx64, ty64, 0, n, sx, sy, sm.red_v, sm.red_i, idx, dist);   // rrt.py:68
            if (!(dist >= pp.d_min && dist < pp.d_max)) continue;                     // rrt.py:69
            iters++;
            // pass 0: the extension toward the sample (rrt.py:72); pass 1: the retry toward the goal
            // from the new node nearest to it (rrt.py:78-80).  One call site keeps one copy of the
            // extension code in the kernel.
            int from = idx;
            double tx = sx, ty = sy;
            for (int pass = 0; pass < 2; pass++) {
                const int n0 = n;
                steer(from, tx, ty);
                if (pass == 0 && n == n0) break;                                      // rrt.py:74
                if (last_at_goal()) { reached = true; break; }                        // rrt.py:75, :81
                if (pass == 1) break;
                int best;
                double bd;
                nearest_block<T>(tx64, ty64, n0, n, gx, gy, sm.red_v, sm.red_i, best, bd);   // rrt.py:78
                if (!(bd < pp.retry_radius) || over) break;                           // rrt.py:79
                from = best; tx = gx; ty = gy;
            }
            if (reached) break;
        }
        if (tid == 0) {
            f.n_nodes[q] = n;
            if (f.pool) f.pool_used[q] = pool_used;
            iters_out[q] = iters; cursor_out[q] = cur;
            reached_out[q] = reached ? 1 : 0;
            if (flags_out) flags_out[q] = (uint8_t)(fl | (over ? 4 : 0));
            if (over) atomicAdd(overflow, 1);
            if (stats) {
                atomicAdd(&stats[0], n_ext);
                atomicAdd(&stats[1], n_ctl);
                atomicAdd(&stats[2], rolls);
            }
        }
    }
}

extern "C" int crlk_plan_dwa(const CrlkEnv *env, const CrlkDwaParams *dwa, const CrlkExtendParams *ext,
                             const CrlkPlanParams *plan, const CrlkForestView *forest, const double *goals,
                             const double *streams, int32_t S, int32_t Q, int32_t *iters, int32_t *cursor,
                             uint8_t *reached, uint8_t *flags, int32_t *overflow, uint64_t *stats,
                             const int32_t *order, void *stream)
{
    CRLK_ENV_OK(env);
    CRLK_REQUIRE(plan != nullptr && forest != nullptr, "plan/forest is NULL");
    DwaP d;
    ExtP x;
    int rc = crlk_make_dwa(env, dwa, &d);
    if (rc) return rc;
    rc = crlk_make_ext(env, ext, &x);
    if (rc) return rc;
    CRLK_REQUIRE(Q >= 0 && S >= 0 && plan->max_iter >= 0, "shape");
    if (Q == 0) return CRLK_OK;
    CRLK_REQUIRE(goals && iters && cursor && reached && overflow, "NULL array");
    CRLK_REQUIRE(forest->x64 && forest->y64 && forest->states && forest->parent && forest->n_nodes &&
                 forest->stride >= 1, "forest");
    CRLK_REQUIRE(!forest->pool || (forest->seg_off && forest->seg_len && forest->seg_flag && forest->pool_used),
                 "forest path pool");
    CRLK_REQUIRE((forest->x32 == nullptr) == (forest->y32 == nullptr), "x32/y32 must both be given or both NULL");
    PlanP pp;
    fill_planp(plan, S, &pp);
    size_t smem = crlk_dwa_smem_bytes(d);
    const bool wide = dwa_use_wide(Q);
    auto kern = d.use_clearance ? (wide ? k_plan_dwa<DWA_WIDE_THREADS, 1, true> : k_plan_dwa<PLAN_THREADS, PLAN_MIN_BLOCKS, true>)
                                : (wide ? k_plan_dwa<DWA_WIDE_THREADS, 1, false> : k_plan_dwa<PLAN_THREADS, PLAN_MIN_BLOCKS, false>);
    const int threads = wide ? DWA_WIDE_THREADS : PLAN_THREADS;
    rc = set_smem_attr((const void *)kern, smem);
    if (rc) return rc;
    int dev = env->device, sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    CRLK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = sms * per_sm;
    if (grid > Q) grid = Q;
    int *counter = nullptr;
    rc = get_work_counter(dev, &counter, (cudaStream_t)stream);
    if (rc) return rc;
    kern<<<grid, threads, smem, (cudaStream_t)stream>>>(
        env->p, d, x, env->os, *forest, pp, goals, streams, Q, iters, cursor, reached, flags, overflow,
        counter, (unsigned long long *)stats, order);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

// ------------------------------------------------------------ nearest node (K3)

#define NN_THREADS 256
#define NN_CAND 512

struct NNPartial { double d, second; int idx; int pad; };

__device__ __forceinline__ void nn_merge(double &d, int &i, double &s, double od, int oi, double os)
{
    s = fmin(fmin(s, os), fmax(d, od));
    if (od < d || (od == d && oi < i)) { d = od; i = oi; }
}

// Block-wide exact (d, idx, second) reduction; result valid in thread 0.
__device__ void nn_block_reduce(double &d, int &i, double &s, double *sv, double *ss, int *si)
{
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double od = __shfl_xor_sync(0xffffffffu, d, o);
        double os = __shfl_xor_sync(0xffffffffu, s, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        nn_merge(d, i, s, od, oi, os);
    }
    if (lane == 0) { sv[warp] = d; ss[warp] = s; si[warp] = i; }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < NN_THREADS / 32; w++) nn_merge(d, i, s, sv[w], si[w], ss[w]);
}

// Scan one segment of one tree.  grid = Q * splits CTAs.
//   f32 path: squared distances in float32 over the SoA float copies (8 B per
//   node of HBM traffic); every node whose value falls inside a guard band of
//   the running block minimum is queued and re-scored exactly in float64.
//   f64 path (x32 == NULL): exact distances straight from the float64 arrays.
__global__ void __launch_bounds__(NN_THREADS)
k_nearest(const double *__restrict__ x64, const double *__restrict__ y64,
          const float *__restrict__ x32, const float *__restrict__ y32, long long stride, int xy_step,
          const int *__restrict__ n_nodes, const double *__restrict__ targets, int target_stride,
          float band_abs, int splits, int vec_ok, NNPartial *partial)
{
    __shared__ unsigned s_minbits;
    __shared__ int s_cnt;
    __shared__ int s_ci[NN_CAND];
    __shared__ float s_cv[NN_CAND];
    __shared__ double sv[NN_THREADS / 32], ss[NN_THREADS / 32];
    __shared__ int si[NN_THREADS / 32];
    const int q = blockIdx.x / splits, split = blockIdx.x - q * splits;
    const int n = n_nodes[q];
    const double tx = targets[(size_t)target_stride * q], ty = targets[(size_t)target_stride * q + 1];
    // segment bounds, multiples of 4 so that vector loads stay aligned
    int per = (((n + splits - 1) / splits) + 3) & ~3;
    int lo = min(n, split * per), hi = min(n, lo + per);
    const size_t base = (size_t)stride * q;
    double bd = CUDART_INF, bs = CUDART_INF;
    int bi = 0x7fffffff;
    bool exact_scan = (x32 == nullptr);

    if (!exact_scan) {
        if (threadIdx.x == 0) { s_minbits = 0x7f800000u; s_cnt = 0; }
        __syncthreads();
        const float ftx = (float)tx, fty = (float)ty;
        const float *xs = x32 + base, *ys = y32 + base;
        volatile unsigned *vmin = &s_minbits;
        auto visit = [&](float e2, int idx) {
            float m2 = __uint_as_float(*vmin);
            if (e2 <= m2 * 1.001f + band_abs) {
                if (e2 < m2) atomicMin(&s_minbits, __float_as_uint(e2));
                int pos = atomicAdd(&s_cnt, 1);
                if (pos < NN_CAND) { s_ci[pos] = idx; s_cv[pos] = e2; }
            }
        };
        int hi4 = lo + ((hi - lo) & ~3);
        if (vec_ok) {
            auto score4 = [&](float4 vx, float4 vy, int i) {
                float dx0 = vx.x - ftx, dy0 = vy.x - fty, dx1 = vx.y - ftx, dy1 = vy.y - fty;
                float dx2 = vx.z - ftx, dy2 = vy.z - fty, dx3 = vx.w - ftx, dy3 = vy.w - fty;
                float e0 = dx0 * dx0 + dy0 * dy0, e1 = dx1 * dx1 + dy1 * dy1;
                float e2 = dx2 * dx2 + dy2 * dy2, e3 = dx3 * dx3 + dy3 * dy3;
                float em = fminf(fminf(e0, e1), fminf(e2, e3));
                float m2 = __uint_as_float(*vmin);
                if (em <= m2 * 1.001f + band_abs) {
                    visit(e0, i); visit(e1, i + 1); visit(e2, i + 2); visit(e3, i + 3);
                }
            };
            const int stride4 = NN_THREADS * 4;
            int i = lo + threadIdx.x * 4;
            // four independent 16-byte loads per array in flight per thread
            for (; i + 3 * stride4 < hi4; i += 4 * stride4) {
                float4 x0 = __ldcs(reinterpret_cast<const float4 *>(xs + i));
                float4 y0 = __ldcs(reinterpret_cast<const float4 *>(ys + i));
                float4 x1 = __ldcs(reinterpret_cast<const float4 *>(xs + i + stride4));
                float4 y1 = __ldcs(reinterpret_cast<const float4 *>(ys + i + stride4));
                float4 x2 = __ldcs(reinterpret_cast<const float4 *>(xs + i + 2 * stride4));
                float4 y2 = __ldcs(reinterpret_cast<const float4 *>(ys + i + 2 * stride4));
                float4 x3 = __ldcs(reinterpret_cast<const float4 *>(xs + i + 3 * stride4));
                float4 y3 = __ldcs(reinterpret_cast<const float4 *>(ys + i + 3 * stride4));
                score4(x0, y0, i);
                score4(x1, y1, i + stride4);
                score4(x2, y2, i + 2 * stride4);
                score4(x3, y3, i + 3 * stride4);
            }
            for (; i < hi4; i += stride4) {
                float4 vx = __ldcs(reinterpret_cast<const float4 *>(xs + i));
                float4 vy = __ldcs(reinterpret_cast<const float4 *>(ys + i));
                score4(vx, vy, i);
            }
        } else {
            hi4 = lo;
        }
        for (int j = hi4 + threadIdx.x; j < hi; j += NN_THREADS) {
            float dx = xs[j] - ftx, dy = ys[j] - fty;
            visit(dx * dx + dy * dy, j);
        }
        __syncthreads();
        if (s_cnt > NN_CAND) exact_scan = true;      // (block-uniform) queue overflow: rescan exactly
        else {
            float m2 = __uint_as_float(s_minbits);
            float lim = m2 * 1.001f + band_abs;
            for (int c = threadIdx.x; c < s_cnt; c += NN_THREADS) {
                if (s_cv[c] <= lim) {
                    int j = s_ci[c];
                    double dx = x64[(base + j) * xy_step] - tx, dy = y64[(base + j) * xy_step] - ty;
                    double dd = sqrt(fma(dy, dy, dx * dx));            // np.linalg.norm (rrt.py:159)
                    nn_merge(bd, bi, bs, dd, j, CUDART_INF);
                }
            }
        }
    }
    if (exact_scan) {
        bd = CUDART_INF; bs = CUDART_INF; bi = 0x7fffffff;
        for (int j = lo + threadIdx.x; j < hi; j += NN_THREADS) {
            double dx = x64[(base + j) * xy_step] - tx, dy = y64[(base + j) * xy_step] - ty;
            double dd = sqrt(fma(dy, dy, dx * dx));
            nn_merge(bd, bi, bs, dd, j, CUDART_INF);
        }
    }
    nn_block_reduce(bd, bi, bs, sv, ss, si);
    if (threadIdx.x == 0) {
        NNPartial p;
        p.d = bd; p.second = bs; p.idx = bi; p.pad = 0;
        partial[(size_t)q * splits + split] = p;
    }
}

__global__ void k_nearest_final(const NNPartial *__restrict__ partial, int splits, int Q, int *idx,
                                double *dist, uint8_t *near_tie)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    double d = CUDART_INF, s = CUDART_INF;
    int i = 0x7fffffff;
    for (int k = 0; k < splits; k++) {
        NNPartial p = partial[(size_t)q * splits + k];
        nn_merge(d, i, s, p.d, p.idx, p.second);
    }
    idx[q] = (i == 0x7fffffff) ? 0 : i;
    dist[q] = d;
    if (near_tie) {
        double gap = s - d;
        near_tie[q] = (gap > 0.0 && gap <= 1e-12 * fabs(d)) ? 1 : 0;
    }
}

static int nn_splits(int Q, int max_nodes)
{
    static const int waves = getenv("CRLK_NN_WAVES") ? atoi(getenv("CRLK_NN_WAVES")) : 1;   // measured best: 1
    int want = (waves * 148 * 8 + Q - 1) / Q;          // several waves of 8 resident CTAs per SM
    int cap = (max_nodes + 8191) / 8192;               // at least 8k nodes per CTA
    int s = want < cap ? want : cap;
    return s < 1 ? 1 : s;
}

extern "C" int64_t crlk_nearest_scratch_bytes(int32_t Q, int32_t max_nodes)
{
    if (Q <= 0) return 0;
    return (int64_t)sizeof(NNPartial) * Q * nn_splits(Q, max_nodes);
}

extern "C" int crlk_nearest(const double *x64, const double *y64, int32_t xy_step, const float *x32, const float *y32,
                            int64_t stride, const int32_t *n_nodes, int32_t max_nodes,
                            double coord_bound, const double *targets, int32_t target_stride,
                            int32_t Q, int32_t *idx, double *dist, uint8_t *near_tie, void *scratch,
                            void *stream)
{
    CRLK_REQUIRE(Q >= 0, "Q < 0");
    if (Q == 0) return CRLK_OK;
    CRLK_REQUIRE(x64 && y64 && n_nodes && targets && idx && dist && scratch, "NULL array");
    CRLK_REQUIRE((x32 == nullptr) == (y32 == nullptr), "x32/y32 must both be given or both NULL");
    CRLK_REQUIRE(target_stride >= 2 && max_nodes >= 1 && stride >= max_nodes && xy_step >= 1, "shape");
    CRLK_REQUIRE(x32 == nullptr || coord_bound > 0, "coord_bound must be > 0 with the float32 scan");
    // guard band of the float32 screen: |d_f - d| <= delta = 8 * bound * 2^-24
    double delta = 8.0 * coord_bound * 5.9604644775390625e-08;
    float band_abs = (float)(1.7e7 * delta * delta);
    int splits = nn_splits(Q, max_nodes);
    int vec_ok = (x32 != nullptr) && (stride % 4 == 0) && (((uintptr_t)x32 & 15) == 0) &&
                 (((uintptr_t)y32 & 15) == 0);
    CRLK_REQUIRE((long long)Q * splits < 2147483647LL, "grid too large");
    k_nearest<<<Q * splits, NN_THREADS, 0, (cudaStream_t)stream>>>(x64, y64, x32, y32, (long long)stride, xy_step,
                                                            n_nodes, targets, target_stride,
                                                            band_abs, splits, vec_ok,
                                                            (NNPartial *)scratch);
    CRLK_LAUNCH_CHECK();
    k_nearest_final<<<(Q + 127) / 128, 128, 0, (cudaStream_t)stream>>>((NNPartial *)scratch, splits,
                                                                       Q, idx, dist, near_tie);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

// ------------------------------------------------------------------ plumbing

__global__ void k_gather_rows(const double *__restrict__ src, long long q_stride, int row_len,
                              const int *__restrict__ idx, int Q, double *dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q * row_len) return;
    int q = i / row_len, c = i - q * row_len;
    dst[i] = src[(size_t)q_stride * q + (size_t)idx[q] * row_len + c];
}

extern "C" int crlk_gather_rows(const double *src, int64_t q_stride, int32_t row_len,
                                const int32_t *idx, int32_t Q, double *dst, void *stream)
{
    CRLK_REQUIRE(Q >= 0 && row_len >= 1, "shape");
    if (Q == 0) return CRLK_OK;
    CRLK_REQUIRE(src && idx && dst, "NULL array");
    int n = Q * row_len;
    k_gather_rows<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src, (long long)q_stride, row_len,
                                                                   idx, Q, dst);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}
