// Device-side building blocks of the extend path (double precision, compiled
// with --fmad=false so that no multiply-add is fused unless written as fma()).
//
// Design notes (B200-first, not a translation of the Python reference):
//   * the velocity chain (v, omega) of a differential-drive rollout does not
//     depend on the pose, and the heading chain depends only on the omega
//     command, so a DWA grid of nv x nw rollouts needs only nw sincos chains
//     and nv velocity chains; positions are then pure multiply/add work;
//   * robot-vs-obstacle tests are an FP32 conservative screen over the whole
//     obstacle set followed by an exact FP64 evaluation of the few survivors.
#pragma once
#include "crlk_internal.h"

#define CRLK_PI 3.141592653589793
#define CRLK_TWO_PI 6.283185307179586
#define CRLK_COLLISION_DIS 0.05     // rigid.py:311
#define CRLK_NEAR_EPS 1e-6          // states this close to the threshold are reported, not compared
#define CRLK_CULL_MARGIN 1e-3f      // FP32 screen slack (coordinates are O(50): 1 ulp ~ 4e-6)
#define CRLK_CLEARANCE_CAP 100.0    // differential_env.py:127

// Python builtin min/max semantics on floats (first argument wins ties).
__device__ __forceinline__ double pymin(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double pymax(double a, double b) { return (b > a) ? b : a; }

// differential_env.py:20-24 (floor-mod by 2*pi, then wrap into (-pi, pi]).
__device__ __forceinline__ double normalize_angle(double a)
{
    double m = fmod(a, CRLK_TWO_PI);
    if (m != 0.0) {
        if (m < 0.0) m += CRLK_TWO_PI;
    } else {
        m = 0.0;
    }
    if (m > CRLK_PI) m -= CRLK_TWO_PI;
    return m;
}

// One base_dt sub-step of one velocity axis under a velocity command:
// differential_env.py:70 (input_u) + :87-95 (clamp into dynamic_window(state, base_dt)).
__device__ __forceinline__ double vel_substep(double cur, double cmd, double acc, double lo,
                                              double hi, double dt)
{
    double iu = (cmd - cur) / dt;
    double u = cur + iu * dt;
    double wlo = pymax(cur - acc * dt, lo);
    double whi = pymin(cur + acc * dt, hi);
    u = pymax(wlo, u);
    u = pymin(whi, u);
    return u;
}

// Same under an acceleration command (differential_env.py:76-95).
__device__ __forceinline__ double acc_substep(double cur, double a, double acc, double lo,
                                              double hi, double dt)
{
    double u = cur + a * dt;
    double wlo = pymax(cur - acc * dt, lo);
    double whi = pymin(cur + acc * dt, hi);
    u = pymax(wlo, u);
    u = pymin(whi, u);
    return u;
}

// ------------------------------------------------- correctly rounded atan2 / sincos
#include "crlk_cr_tables.h"

// full tables in global memory (read-only path): the sincos table and the slow path of atan2
__device__ const unsigned long long g_cr_atan_bits[CRLK_CR_ATAN_NK * CRLK_CR_ATAN_FIELDS] = CRLK_CR_ATAN_INIT;
__device__ const unsigned long long g_cr_sc_bits[CRLK_CR_SC_NK * 4] = CRLK_CR_SC_INIT;
#define CRM_ATAN_FULL(tab) ((const double *)g_cr_atan_bits)
#include "crlk_crmath.h"

#define CRLK_ATAN_SMEM_DOUBLES (CRLK_CR_ATAN_NK * CRLK_CR_ATAN_FAST_FIELDS)

// Stages the fast-path fields of the atan table in shared memory ([field][interval]: lanes that
// fall into different intervals read different banks); tid / n = the loading group's thread index / size.
__device__ __forceinline__ void load_atan_table(double *dst, int tid, int n)
{
    for (int i = tid; i < CRLK_ATAN_SMEM_DOUBLES; i += n) dst[i] = __longlong_as_double((long long)g_cr_atan_bits[i]);
}

// 1 / a for a normal, positive a: hardware seed (about 20 bits) + two Newton steps.
__device__ __forceinline__ double crlk_rcp_pos(double a) { return crm_rcp(a); }

static __device__ __noinline__ double crlk_atan2_special(double y, double x) { return atan2(y, x); }

// atan2 of the rollout scoring loop (dwa.py:58), correctly rounded (crlk_crmath.h): the same bits
// as the C library's result that np.arctan2 returns for Python scalars, up to the library's own
// rare misroundings (7.5e-4 of random arguments, tools/cr_check.cpp).  `tab`: load_atan_table's
// copy in shared memory.  Zero, tiny, huge and NaN arguments take the library path.
__device__ __forceinline__ double crlk_atan2(double y, double x, const double *tab)
{
    const double ay = fabs(y), ax = fabs(x);
    const double mx = ay > ax ? ay : ax, mn = ay > ax ? ax : ay;
    // the fast path wants 1e-280 < mx < 1e280 and a normal quotient: integer range tests on the
    // exponent fields (biased exponents 93 .. 1953); NaN, infinities, zero and subnormals fail them
    const unsigned ex = ((unsigned)__double2hiint(mx) >> 20) & 0x7ffu;
    const unsigned en = ((unsigned)__double2hiint(mn) >> 20) & 0x7ffu;
    if (ex - 93u > 1860u || en + 900u < ex || en == 0u) return crlk_atan2_special(y, x);
    return crm_atan2_core(y, x, tab);
}

// sin and cos of a heading, correctly rounded for |th| <= 3.15 (math.sin / math.cos of
// differential_env.py:98-99 on a normalised angle); other arguments take the library path.
__device__ __forceinline__ void crlk_sincos(double th, double *sn, double *cs)
{
    if (!crm_sincos_core(th, (const double *)g_cr_sc_bits, sn, cs)) sincos(th, sn, cs);
}

// Pose update of one sub-step (differential_env.py:97-103) given the new (u0, u1).
__device__ __forceinline__ void pose_substep(double *s, double u0, double u1, double dt)
{
    double th = normalize_angle(s[2] + u1 * dt);
    double sn, cs;
    crlk_sincos(th, &sn, &cs);
    s[2] = th;
    s[0] = s[0] + u0 * cs * dt;
    s[1] = s[1] + u0 * sn * dt;
    s[3] = u0;
    s[4] = u1;
}

// motion_velocity (differential_env.py:55-73) with n_sub pre-counted sub-steps.
__device__ __forceinline__ void motion_velocity_dev(const EnvP &e, double *s, double cv, double cw,
                                                    int n_sub)
{
    for (int k = 0; k < n_sub; k++) {
        double u0 = vel_substep(s[3], cv, e.max_acc_v, e.min_v, e.max_v, e.base_dt);
        double u1 = vel_substep(s[4], cw, e.max_acc_w, -e.max_w, e.max_w, e.base_dt);
        pose_substep(s, u0, u1, e.base_dt);
    }
}

// ------------------------------------------------------------------ geometry

struct RobotPose {
    double x, y, c, s;
    double wx[4], wy[4];   // robot corners in the world frame: (l,b) (l,t) (r,t) (r,b)
};

// Footprint corners from a pose whose heading is given as (cos, sin).
__device__ __forceinline__ void robot_pose_from_cs(const EnvP &e, double x, double y, double cs, double sn,
                                                   RobotPose &g)
{
    g.x = x; g.y = y; g.c = cs; g.s = sn;
    const double bx[4] = {(double)e.rl, (double)e.rl, (double)e.rr, (double)e.rr};
    const double by[4] = {(double)e.rb, (double)e.rt, (double)e.rt, (double)e.rb};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        g.wx[i] = (cs * bx[i] + (-sn) * by[i]) + x;   // rigid.py:223-227
        g.wy[i] = (sn * bx[i] + cs * by[i]) + y;
    }
}

__device__ __forceinline__ void robot_pose_setup(const EnvP &e, double x, double y, double th,
                                                 RobotPose &g)
{
    double sn, cs;
    crlk_sincos(th, &sn, &cs);
    g.x = x; g.y = y; g.c = cs; g.s = sn;
    const double bx[4] = {(double)e.rl, (double)e.rl, (double)e.rr, (double)e.rr};
    const double by[4] = {(double)e.rb, (double)e.rt, (double)e.rt, (double)e.rb};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        g.wx[i] = (cs * bx[i] + (-sn) * by[i]) + x;   // rigid.py:223-227
        g.wy[i] = (sn * bx[i] + cs * by[i]) + y;
    }
}

// Segment s->g against a closed AABB: slab test (rigid.py:45-116) in
// unnormalised form; t in [0, 1] spans the segment, zero direction components
// use the reference's 1e300 reciprocal.
__device__ __forceinline__ bool seg_hits_aabb(double sx, double sy, double gx, double gy, double l,
                                              double b, double r, double t)
{
    double dx = gx - sx, dy = gy - sy;
    double fx = (dx == 0.0) ? 1e300 : 1.0 / dx;
    double fy = (dy == 0.0) ? 1e300 : 1.0 / dy;
    double t1 = (l - sx) * fx, t2 = (r - sx) * fx;
    double t3 = (b - sy) * fy, t4 = (t - sy) * fy;
    double tmin = fmax(fmin(t1, t2), fmin(t3, t4));
    double tmax = fmin(fmax(t1, t2), fmax(t3, t4));
    return (tmax >= 0.0) && (tmin <= tmax) && (tmin <= 1.0);
}

// Squared distance from a point to the BOUNDARY of a rectangle (what the min
// over the four RectObs.dis2point edges is, rigid.py:168-180): outside it is
// the distance to the solid, inside the smallest margin.
__device__ __forceinline__ double point_rect_boundary_d2(double px, double py, double l, double b,
                                                         double r, double t)
{
    double ex = fmax(fmax(l - px, px - r), 0.0);
    double ey = fmax(fmax(b - py, py - t), 0.0);
    double in = fmin(fmin(px - l, r - px), fmin(py - b, t - py));
    return (in > 0.0) ? in * in : ex * ex + ey * ey;
}

// RectRobot.dis2obs (rigid.py:265-313): -1 on collision, else the distance.
// `definite` / `marginal` classify the decision against the 0.05 threshold.
__device__ __forceinline__ double dis2obs_exact(const EnvP &e, const RobotPose &g, float4 o,
                                                bool &definite, bool &marginal)
{
    double l = o.x, b = o.y, r = o.z, t = o.w;
    if (e.radius > 0.0) {
        // CircleRobot.dis2obs (rigid.py:336-347): centre inside the closed rectangle collides,
        // else boundary distance minus the radius against the same 0.05 threshold
        if (g.x >= l && g.x <= r && g.y >= b && g.y <= t) { definite = true; marginal = false; return -1.0; }
        double dc = sqrt(point_rect_boundary_d2(g.x, g.y, l, b, r, t)) - e.radius;
        marginal = fabs(dc - CRLK_COLLISION_DIS) < CRLK_NEAR_EPS;
        definite = dc <= CRLK_COLLISION_DIS - CRLK_NEAR_EPS;
        return (dc > CRLK_COLLISION_DIS) ? dc : -1.0;
    }
    bool hit = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int j = (i + 1) & 3;
        hit |= seg_hits_aabb(g.wx[i], g.wy[i], g.wx[j], g.wy[j], l, b, r, t);
    }
    if (hit) { definite = true; marginal = false; return -1.0; }
    double d2 = 1e300;
#pragma unroll
    for (int i = 0; i < 4; i++) d2 = fmin(d2, point_rect_boundary_d2(g.wx[i], g.wy[i], l, b, r, t));
    const double ox[4] = {l, l, r, r}, oy[4] = {b, t, t, b};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double qx = ox[i] - g.x, qy = oy[i] - g.y;       // rigid.py:231-237 as R^T (p - t)
        double px = g.c * qx + g.s * qy;
        double py = (-g.s) * qx + g.c * qy;
        d2 = fmin(d2, point_rect_boundary_d2(px, py, (double)e.rl, (double)e.rb, (double)e.rr,
                                             (double)e.rt));
    }
    double dis = sqrt(d2);
    marginal = fabs(dis - CRLK_COLLISION_DIS) < CRLK_NEAR_EPS;
    definite = dis <= CRLK_COLLISION_DIS - CRLK_NEAR_EPS;
    return (dis > CRLK_COLLISION_DIS) ? dis : -1.0;
}

// FP32 screen: squared distance from the robot's body origin to the solid AABB.
__device__ __forceinline__ float aabb_center_d2(float4 o, float px, float py)
{
    float ex = fmaxf(fmaxf(o.x - px, px - o.z), 0.0f);
    float ey = fmaxf(fmaxf(o.y - py, py - o.w), 0.0f);
    return ex * ex + ey * ey;
}

// FP32 lower bound of the footprint-to-obstacle distance: distance from the obstacle's CENTRE to
// the robot rectangle (body frame) minus the obstacle's circumradius (dist(A, .) is 1-Lipschitz),
// less a margin that covers the float32 rounding.  Negative when the centre lies inside the
// footprint (never used to skip then).  Tight for small obstacles (the 0.25 m cells of a dense
// map), useless for long walls -- those are caught by the centre-distance screen before it.
__device__ __forceinline__ float footprint_lb(const EnvP &e, float cf, float sf, float px, float py, float4 o)
{
    const float cx = 0.5f * (o.x + o.z) - px, cy = 0.5f * (o.y + o.w) - py;
    const float hx = 0.5f * (o.z - o.x), hy = 0.5f * (o.w - o.y);
    const float ro = sqrtf(hx * hx + hy * hy);
    float d;
    if (e.radius > 0.0) {
        d = sqrtf(cx * cx + cy * cy) - (float)e.radius;
    } else {
        const float qx = cf * cx + sf * cy, qy = cf * cy - sf * cx;
        const float ex = fmaxf(fmaxf(e.rl - qx, qx - e.rr), 0.0f);
        const float ey = fmaxf(fmaxf(e.rb - qy, qy - e.rt), 0.0f);
        d = sqrtf(ex * ex + ey * ey);
    }
    return d - ro * 1.0001f - CRLK_CULL_MARGIN;
}

// Visit every obstacle listed in the grid cells that the box [p - thr, p + thr]
// touches, `nthreads` cooperating threads striding each cell's list.  An
// obstacle spanning several visited cells is visited once per cell (harmless
// for min / or reductions).  The loops are uniform across the cooperating
// threads, so the visitor may use warp collectives when nthreads == 32.
template <class F>
__device__ __forceinline__ void grid_visit(const ObsSet &os, float px, float py, float thr, int t,
                                           int nthreads, F f)
{
    float r = thr + 2e-3f;
    float hx = (float)(os.ncx - 1), hy = (float)(os.ncy - 1);
    int cx0 = (int)fminf(fmaxf(floorf((px - r - os.gx0) * os.ginv), 0.f), hx);
    int cx1 = (int)fminf(fmaxf(floorf((px + r - os.gx0) * os.ginv), 0.f), hx);
    int cy0 = (int)fminf(fmaxf(floorf((py - r - os.gy0) * os.ginv), 0.f), hy);
    int cy1 = (int)fminf(fmaxf(floorf((py + r - os.gy0) * os.ginv), 0.f), hy);
    for (int cy = cy0; cy <= cy1; cy++) {
        // the cells of one grid row are adjacent in the CSR, so a row is one range
        int s = __ldg(&os.cell_start[cy * os.ncx + cx0]);
        int e = __ldg(&os.cell_start[cy * os.ncx + cx1 + 1]);
        for (int kb = s; kb < e; kb += nthreads) {
            int k = kb + t;
            bool in = k < e;
            int o = in ? __ldg(&os.cell_items[k]) : -1;
            f(o, in);
        }
    }
}

// The same visit for a whole CTA: the grid rows of the box are handled side by side (thread t takes
// row t / per, entries t % per, t % per + per, ...), so a thread's loads are one dependent chain
// cell_start -> cell_rects instead of a chain cell_start -> cell_items -> rects per row.  Not uniform across threads:
// the visitor must not use collectives.
template <class F>
__device__ __forceinline__ void grid_visit_rows(const ObsSet &os, float px, float py, float thr, int t,
                                                int nthreads, F f)
{
    float r = thr + 2e-3f;
    float hx = (float)(os.ncx - 1), hy = (float)(os.ncy - 1);
    int cx0 = (int)fminf(fmaxf(floorf((px - r - os.gx0) * os.ginv), 0.f), hx);
    int cx1 = (int)fminf(fmaxf(floorf((px + r - os.gx0) * os.ginv), 0.f), hx);
    int cy0 = (int)fminf(fmaxf(floorf((py - r - os.gy0) * os.ginv), 0.f), hy);
    int cy1 = (int)fminf(fmaxf(floorf((py + r - os.gy0) * os.ginv), 0.f), hy);
    const int nrows = cy1 - cy0 + 1;
    if (nrows > nthreads) {       // never with the planners' footprints; keep the result complete anyway
        grid_visit(os, px, py, thr, t, nthreads, [&](int o, bool in) { if (in) f(__ldg(&os.rects[o])); });
        return;
    }
    const int per = nthreads / nrows;
    const int row = t / per, idx = t - row * per;
    if (row >= nrows) return;
    const int cy = cy0 + row;
    const int s = __ldg(&os.cell_start[cy * os.ncx + cx0]);
    const int e = __ldg(&os.cell_start[cy * os.ncx + cx1 + 1]);
    for (int k = s + idx; k < e; k += per) f(__ldg(&os.cell_rects[k]));
}

// get_clearance (differential_env.py:126-132) for one pose by ONE thread.  The grid cells are
// visited ring by ring around the pose's own cell (Chebyshev rings 0, 1, 2, ...): everything listed
// only in ring k or beyond lies at least (k - 1) cells away, so the search stops as soon as that
// distance exceeds best + rc.  An obstacle listed in several cells is handled exactly once, in the
// cell that holds its point closest to the pose (which is also the first ring that reaches it).
// Each obstacle passes two FP32 screens (body origin to the solid rectangle, footprint to the
// obstacle's bounding circle) before the exact FP64 test.  Returns min(100, min distance), or -1 on
// collision -- the same value as the scan over all obstacles.
// `bound`: only distances below it matter to the caller (a running minimum); the result is then
// min(bound, clearance), found without looking at obstacles that cannot get under the bound.
// This version runs the exact FP64 test on every obstacle that passes the screens (the slow but
// unconditional path; clearance_grid_thread below falls back to it).
static __device__ __noinline__ double clearance_grid_exact(const EnvP &e, const ObsSet &os, double x, double y,
                                                     double th, bool *definite_out, bool *marginal_out,
                                                     double bound = CRLK_CLEARANCE_CAP)
{
    RobotPose g;
    robot_pose_setup(e, x, y, th, g);
    const float px = (float)x, py = (float)y;
    const float cf = (float)g.c, sf = (float)g.s;
    const float hx = (float)(os.ncx - 1), hy = (float)(os.ncy - 1);
    const float cell = 1.0f / os.ginv;
    auto cellx = [&](float v) { return (int)fminf(fmaxf(floorf((v - os.gx0) * os.ginv), 0.f), hx); };
    auto celly = [&](float v) { return (int)fminf(fmaxf(floorf((v - os.gy0) * os.ginv), 0.f), hy); };
    const int cx0 = cellx(px), cy0 = celly(py);
    double best = fmin(bound, CRLK_CLEARANCE_CAP);
    bool definite = false, marginal = false;
    float bf = (float)best * 1.0001f;                       // float upper bound of the best distance so far
    float lim = bf + e.rc + CRLK_CULL_MARGIN;               // body origin farther than this: cannot beat it
    const int kmax = max(max(cx0, os.ncx - 1 - cx0), max(cy0, os.ncy - 1 - cy0));
    auto visit = [&](int cx, int cy) {
        const int c = cy * os.ncx + cx;
        const int s = __ldg(&os.cell_start[c]), en = __ldg(&os.cell_start[c + 1]);
        for (int k = s; k < en; k++) {
            const float4 o = __ldg(&os.cell_rects[k]);
            const float qx = fminf(fmaxf(px, o.x), o.z), qy = fminf(fmaxf(py, o.y), o.w);   // closest point of the AABB
            const float dx = px - qx, dy = py - qy;
            if (dx * dx + dy * dy > lim * lim) continue;
            if (cellx(qx) != cx || celly(qy) != cy) continue;            // handled from another cell
            if (footprint_lb(e, cf, sf, px, py, o) > bf) continue;
            bool dd, mm;
            const double dis = dis2obs_exact(e, g, o, dd, mm);
            definite |= dd; marginal |= mm;
            if (dis < best) {
                best = dis;
                bf = (float)fmax(best, 0.0) * 1.0001f;
                lim = bf + e.rc + CRLK_CULL_MARGIN;
            }
        }
    };
    for (int k = 0; k <= kmax && best >= 0.0; k++) {
        if (k >= 1 && (float)(k - 1) * cell > lim + 2e-3f) break;
        if (k == 0) { visit(cx0, cy0); continue; }
        const int xl = cx0 - k, xr = cx0 + k, yb = cy0 - k, yt = cy0 + k;
        const int xa = max(xl, 0), xb = min(xr, os.ncx - 1);
        if (yb >= 0) for (int cx = xa; cx <= xb && best >= 0.0; cx++) visit(cx, yb);
        if (yt < os.ncy) for (int cx = xa; cx <= xb && best >= 0.0; cx++) visit(cx, yt);
        const int ya = max(yb + 1, 0), yc = min(yt - 1, os.ncy - 1);
        if (xl >= 0) for (int cy = ya; cy <= yc && best >= 0.0; cy++) visit(xl, cy);
        if (xr < os.ncx) for (int cy = ya; cy <= yc && best >= 0.0; cy++) visit(xr, cy);
    }
    if (definite_out) *definite_out = definite;
    if (marginal_out) *marginal_out = marginal;
    return (best < 0.0) ? -1.0 : best;
}


// ---- float32 mirror of dis2obs_exact: the same tests and formulas in single precision.  Used only
// to rank obstacles; every value that is reported comes from the float64 routine.
struct RobotPoseF {
    float x, y, c, s;
    float wx[4], wy[4];
};

__device__ __forceinline__ bool seg_hits_aabb_f(float sx, float sy, float gx, float gy, float l, float b, float r,
                                                float t)
{
    float dx = gx - sx, dy = gy - sy;
    float fx = (dx == 0.0f) ? 1e30f : __frcp_rn(dx);
    float fy = (dy == 0.0f) ? 1e30f : __frcp_rn(dy);
    float t1 = (l - sx) * fx, t2 = (r - sx) * fx;
    float t3 = (b - sy) * fy, t4 = (t - sy) * fy;
    float tmin = fmaxf(fminf(t1, t2), fminf(t3, t4));
    float tmax = fminf(fmaxf(t1, t2), fmaxf(t3, t4));
    return (tmax >= 0.0f) && (tmin <= tmax) && (tmin <= 1.0f);
}

__device__ __forceinline__ float point_rect_boundary_d2_f(float px, float py, float l, float b, float r, float t)
{
    float ex = fmaxf(fmaxf(l - px, px - r), 0.0f);
    float ey = fmaxf(fmaxf(b - py, py - t), 0.0f);
    float in = fminf(fminf(px - l, r - px), fminf(py - b, t - py));
    return (in > 0.0f) ? in * in : ex * ex + ey * ey;
}

__device__ __forceinline__ float dis2obs_f32(const EnvP &e, const RobotPoseF &g, float4 o)
{
    if (e.radius > 0.0) {
        if (g.x >= o.x && g.x <= o.z && g.y >= o.y && g.y <= o.w) return -1.0f;
        float dc = sqrtf(point_rect_boundary_d2_f(g.x, g.y, o.x, o.y, o.z, o.w)) - (float)e.radius;
        return (dc > (float)CRLK_COLLISION_DIS) ? dc : -1.0f;
    }
    bool hit = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int j = (i + 1) & 3;
        hit |= seg_hits_aabb_f(g.wx[i], g.wy[i], g.wx[j], g.wy[j], o.x, o.y, o.z, o.w);
    }
    if (hit) return -1.0f;
    float d2 = 1e30f;
#pragma unroll
    for (int i = 0; i < 4; i++) d2 = fminf(d2, point_rect_boundary_d2_f(g.wx[i], g.wy[i], o.x, o.y, o.z, o.w));
    const float ox[4] = {o.x, o.x, o.z, o.z}, oy[4] = {o.y, o.w, o.w, o.y};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float qx = ox[i] - g.x, qy = oy[i] - g.y;
        float px = g.c * qx + g.s * qy, py = g.c * qy - g.s * qx;
        d2 = fminf(d2, point_rect_boundary_d2_f(px, py, e.rl, e.rb, e.rr, e.rt));
    }
    float dis = sqrtf(d2);
    return (dis > (float)CRLK_COLLISION_DIS) ? dis : -1.0f;
}

// Separating-axis lower bound of the rectangle-to-rectangle distance in float32: the gap between the
// two boxes along the world axes (robot's world AABB vs the obstacle) and along the robot's axes
// (robot rectangle vs the obstacle's bounding box in the body frame); the distance of two convex
// sets is at least their gap along any axis.  ~40 instructions, exact whenever a face of one
// rectangle looks at the other.
__device__ __forceinline__ float rect_sat_lb(const EnvP &e, const RobotPoseF &g, float wl, float wr, float wb,
                                             float wt, float4 o)
{
    const float gw = fmaxf(fmaxf(o.x - wr, wl - o.z), fmaxf(o.y - wt, wb - o.w));        // world axes
    const float cx = 0.5f * (o.x + o.z) - g.x, cy = 0.5f * (o.y + o.w) - g.y;
    const float hx = 0.5f * (o.z - o.x), hy = 0.5f * (o.w - o.y);
    const float qx = g.c * cx + g.s * cy, qy = g.c * cy - g.s * cx;                       // obstacle centre, body frame
    const float rx = fabsf(g.c) * hx + fabsf(g.s) * hy, ry = fabsf(g.s) * hx + fabsf(g.c) * hy;   // its half extents there
    const float gb = fmaxf(fmaxf((qx - rx) - e.rr, e.rl - (qx + rx)), fmaxf((qy - ry) - e.rt, e.rb - (qy + ry)));
    return fmaxf(gw, gb) - 1e-4f;
}

#define CLR_EPS 2e-4f      // bound of |float32 distance - float64 distance| used below (coordinates are O(50))
#define CLR_CAND 12

// get_clearance for one pose by ONE thread, float32 ranking + float64 verdict.  The ring search of
// clearance_grid_exact, but an obstacle that passes the two screens is first measured in float32;
// the thread only remembers the obstacles whose float32 distance is within 2 eps of the smallest
// one seen (the exact minimum is necessarily among them) and runs the float64 test on those few at
// the very end -- where all threads of a warp are at the same instruction again, instead of each
// lane dragging the whole warp through a float64 test at a different moment of its scan.
// Falls back to clearance_grid_exact when the candidate list overflows (many obstacles at the same
// distance) or when float32 saw a collision that float64 does not confirm.
static __device__ __noinline__ double clearance_grid_thread(const EnvP &e, const ObsSet &os, double x, double y,
                                                     double th, bool *definite_out, bool *marginal_out,
                                                     double bound = CRLK_CLEARANCE_CAP)
{
    RobotPose g;
    robot_pose_setup(e, x, y, th, g);
    RobotPoseF gf;
    gf.x = (float)g.x; gf.y = (float)g.y; gf.c = (float)g.c; gf.s = (float)g.s;
#pragma unroll
    for (int i = 0; i < 4; i++) { gf.wx[i] = (float)g.wx[i]; gf.wy[i] = (float)g.wy[i]; }
    const float px = gf.x, py = gf.y, cf = gf.c, sf = gf.s;
    const float hx = (float)(os.ncx - 1), hy = (float)(os.ncy - 1);
    const float cell = 1.0f / os.ginv;
    auto cellx = [&](float v) { return (int)fminf(fmaxf(floorf((v - os.gx0) * os.ginv), 0.f), hx); };
    auto celly = [&](float v) { return (int)fminf(fmaxf(floorf((v - os.gy0) * os.ginv), 0.f), hy); };
    const int cx0 = cellx(px), cy0 = celly(py);
    const double b0 = fmin(bound, CRLK_CLEARANCE_CAP);
    float min32 = (float)b0 * 1.000001f + CLR_EPS;          // float32 view of the bound, rounded up
    int cid[CLR_CAND];
    float cd[CLR_CAND];
    int nc = 0;
    bool overflow = false, stop = false;
    float bf = fmaxf(min32, 0.0f) + 2.0f * CLR_EPS;         // candidates have a float32 distance <= bf
    float lim = bf + e.rc + CRLK_CULL_MARGIN;
    const int kmax = max(max(cx0, os.ncx - 1 - cx0), max(cy0, os.ncy - 1 - cy0));
    auto visit = [&](int cx, int cy) {
        const int c = cy * os.ncx + cx;
        const int s = __ldg(&os.cell_start[c]), en = __ldg(&os.cell_start[c + 1]);
        for (int k = s; k < en; k++) {
            const int id = __ldg(&os.cell_items[k]);
            const float4 o = __ldg(&os.cell_rects[k]);
            const float qx = fminf(fmaxf(px, o.x), o.z), qy = fminf(fmaxf(py, o.y), o.w);   // closest point of the AABB
            const float dx = px - qx, dy = py - qy;
            if (dx * dx + dy * dy > lim * lim) continue;
            if (cellx(qx) != cx || celly(qy) != cy) continue;            // handled from another cell
            if (footprint_lb(e, cf, sf, px, py, o) > bf) continue;
            const float d = dis2obs_f32(e, gf, o);
            if (d > min32 + 2.0f * CLR_EPS) continue;
            if (d < 0.0f) {
                // float32 sees a collision: nothing can rank below it; let float64 confirm this one
                cid[0] = id; cd[0] = d; nc = 1; min32 = d; stop = true;
                return;
            }
            if (d < min32) {
                // new smallest float32 distance: drop the remembered obstacles that are now out of range
                min32 = d;
                int m = 0;
                for (int q = 0; q < nc; q++)
                    if (cd[q] <= d + 2.0f * CLR_EPS) { cid[m] = cid[q]; cd[m] = cd[q]; m++; }
                nc = m;
                bf = fmaxf(min32, 0.0f) + 2.0f * CLR_EPS;
                lim = bf + e.rc + CRLK_CULL_MARGIN;
            }
            if (nc < CLR_CAND) { cid[nc] = id; cd[nc] = d; nc++; }
            else overflow = true;
        }
    };
    // one loop over the cells of ring 0, 1, 2, ... in spiral order with a single visit site, so that
    // the lanes of a warp (each on its own pose) meet again at every cell instead of drifting apart
    // through four separate edge loops
    for (int k = 0; k <= kmax && !overflow && !stop; k++) {
        if (k >= 1 && (float)(k - 1) * cell > lim + 2e-3f) break;
        const int ncell = (k == 0) ? 1 : 8 * k;
        int dx = -k, dy = -k, sx = 1, sy = 0, run = 0;
        for (int t = 0; t < ncell && !stop; t++) {
            const int cx = cx0 + dx, cy = cy0 + dy;
            if (cx >= 0 && cx < os.ncx && cy >= 0 && cy < os.ncy) visit(cx, cy);
            // walk the ring: bottom edge left to right, right edge upwards, top edge right to left, left edge down
            dx += sx; dy += sy;
            if (++run == 2 * k) { run = 0; const int tx = sx; sx = -sy; sy = tx; }
        }
    }
    // float64 verdict on the remembered obstacles (all lanes of the warp arrive here together)
    double best = b0;
    bool definite = false, marginal = false, confirmed = false;
    for (int q = 0; q < nc; q++) {
        bool dd, mm;
        const double dis = dis2obs_exact(e, g, __ldg(&os.rects[cid[q]]), dd, mm);
        definite |= dd; marginal |= mm;
        confirmed |= (dis < 0.0);
        best = fmin(best, dis);
    }
    if (overflow || (min32 < 0.0f && !confirmed))
        return clearance_grid_exact(e, os, x, y, th, definite_out, marginal_out, bound);
    if (definite_out) *definite_out = definite;
    if (marginal_out) *marginal_out = marginal;
    return (best < 0.0) ? -1.0 : best;
}

// get_clearance against a short list of obstacles in shared memory (the per-rollout clearance term:
// every pose of a control step lies within the prediction horizon of the current state, so the block
// gathers the obstacles that can matter once and each thread then loops over that list -- the same
// trip count in every lane, no grid walk, no global loads).  `list` must hold every obstacle whose
// rectangle is closer than bound + rc to the pose's body origin.  Same float32 ranking / float64
// verdict as clearance_grid_thread; returns a negative NaN-free sentinel (-2) when the caller has to
// fall back to the grid search (candidate overflow, unconfirmed float32 collision).
static __device__ __noinline__ double clearance_from_list(const EnvP &e, const float4 *list, int n, double x,
                                                          double y, double cs, double sn, double bound)
{
    RobotPose g;
    robot_pose_from_cs(e, x, y, cs, sn, g);
    RobotPoseF gf;
    gf.x = (float)g.x; gf.y = (float)g.y; gf.c = (float)g.c; gf.s = (float)g.s;
#pragma unroll
    for (int i = 0; i < 4; i++) { gf.wx[i] = (float)g.wx[i]; gf.wy[i] = (float)g.wy[i]; }
    const float px = gf.x, py = gf.y;
    const float wl = fminf(fminf(gf.wx[0], gf.wx[1]), fminf(gf.wx[2], gf.wx[3]));
    const float wr = fmaxf(fmaxf(gf.wx[0], gf.wx[1]), fmaxf(gf.wx[2], gf.wx[3]));
    const float wb = fminf(fminf(gf.wy[0], gf.wy[1]), fminf(gf.wy[2], gf.wy[3]));
    const float wt = fmaxf(fmaxf(gf.wy[0], gf.wy[1]), fmaxf(gf.wy[2], gf.wy[3]));
    const double b0 = fmin(bound, CRLK_CLEARANCE_CAP);
    float min32 = (float)b0 * 1.000001f + CLR_EPS;
    int cid[CLR_CAND];
    float cd[CLR_CAND];
    int nc = 0;
    bool overflow = false;
    float bf = fmaxf(min32, 0.0f) + 2.0f * CLR_EPS;
    float lim = bf + e.rc + CRLK_CULL_MARGIN;
    for (int k = 0; k < n; k++) {
        const float4 o = list[k];
        const float qx = fminf(fmaxf(px, o.x), o.z), qy = fminf(fmaxf(py, o.y), o.w);
        const float dx = px - qx, dy = py - qy;
        if (dx * dx + dy * dy > lim * lim) continue;
        if (e.radius > 0.0 ? footprint_lb(e, gf.c, gf.s, px, py, o) > bf
                           : rect_sat_lb(e, gf, wl, wr, wb, wt, o) > bf) continue;
        const float d = dis2obs_f32(e, gf, o);
        if (d > min32 + 2.0f * CLR_EPS) continue;
        if (d < 0.0f) { cid[0] = k; cd[0] = d; nc = 1; min32 = d; break; }
        if (d < min32) {
            min32 = d;
            int m = 0;
            for (int q = 0; q < nc; q++)
                if (cd[q] <= d + 2.0f * CLR_EPS) { cid[m] = cid[q]; cd[m] = cd[q]; m++; }
            nc = m;
            bf = fmaxf(min32, 0.0f) + 2.0f * CLR_EPS;
            lim = bf + e.rc + CRLK_CULL_MARGIN;
        }
        if (nc < CLR_CAND) { cid[nc] = k; cd[nc] = d; nc++; }
        else { overflow = true; break; }
    }
    double best = b0;
    bool confirmed = false;
    for (int q = 0; q < nc; q++) {
        bool dd, mm;
        const double dis = dis2obs_exact(e, g, list[cid[q]], dd, mm);
        confirmed |= (dis < 0.0);
        best = fmin(best, dis);
    }
    if (overflow || (min32 < 0.0f && !confirmed)) return -2.0;
    return (best < 0.0) ? -1.0 : best;
}

// ------------------------------------------------ local occupancy map (R9)

#define OBS_NPTS 729
#define OBS_SIDE 27
#define OBS_NEAR_EPS 1e-9

struct ObsGrid {      // np.arange(-2, 6, 0.3) / np.arange(-4, 4, 0.3) (differential_gym.py:67-71)
    double ba0, ba1, bad;   // start, start+step, delta
    double lr0, lr1, lrd;
};

static inline ObsGrid make_obs_grid()
{
    ObsGrid g;
    g.ba0 = -2.0; g.ba1 = -2.0 + 0.3; g.bad = g.ba1 - g.ba0;
    g.lr0 = -4.0; g.lr1 = -4.0 + 0.3; g.lrd = g.lr1 - g.lr0;
    return g;
}

__device__ __forceinline__ double arange_at(double a0, double a1, double d, int i)
{
    return (i == 0) ? a0 : (i == 1) ? a1 : a0 + (double)i * d;
}

// wTb @ [bx, by, 1] with every operation rounded separately (no FMA contraction,
// whatever the translation unit's -fmad setting): identical bits in every kernel.
__device__ __forceinline__ void body_to_world(double cs, double sn, double x, double y, double bx, double by,
                                              double &px, double &py)
{
    px = __dadd_rn(__dadd_rn(__dmul_rn(cs, bx), __dmul_rn(-sn, by)), x);
    py = __dadd_rn(__dadd_rn(__dmul_rn(sn, bx), __dmul_rn(cs, by)), y);
}

// valid_point_check for one point as a query against the obstacle grid (closed
// intervals, rigid.py:158-166).  A point inside an obstacle always falls in a
// cell of that obstacle's range (the cell function is monotone), so this equals
// the scan over all obstacles.  `near` is OR-ed with "within 1e-9 of a face".
// The raster screen of a sample point alone: 0 = free, 1 = inside an obstacle, 2 = the raster cell is
// only partly covered (or the point lies outside the raster): local_map_occupied decides.
__device__ __forceinline__ unsigned local_map_raster(const ObsSet &os, double px, double py)
{
    const float rx = ((float)px - os.gx0) * os.rinv, ry = ((float)py - os.gy0) * os.rinv;
    if (rx >= 0.f && ry >= 0.f && rx < (float)os.rnx && ry < (float)os.rny)
        return __ldg(&os.raster[(int)ry * os.rnx + (int)rx]);
    return 2u;
}

// The same screen for a point evaluated in float32: the raster's margin (>= 1e-4 m, build_grid) is an order
// of magnitude above the error of a float32 evaluation of body_to_world (< 2e-5 m for coordinates
// below 100 m), so a 0 / 1 answer for the float32 point holds for the exact float64 point as well.
__device__ __forceinline__ unsigned local_map_raster_f(const ObsSet &os, float pxf, float pyf)
{
    const float rx = (pxf - os.gx0) * os.rinv, ry = (pyf - os.gy0) * os.rinv;
    if (rx >= 0.f && ry >= 0.f && rx < (float)os.rnx && ry < (float)os.rny)
        return __ldg(&os.raster[(int)ry * os.rnx + (int)rx]);
    return 2u;
}

template <bool WANT_NEAR = true>
__device__ __forceinline__ bool local_map_occupied(const ObsSet &os, double px, double py, bool &near)
{
    if (!WANT_NEAR) {
        // raster first: most sample points fall into cells that no obstacle touches or that one covers
        const float rx = ((float)px - os.gx0) * os.rinv, ry = ((float)py - os.gy0) * os.rinv;
        if (rx >= 0.f && ry >= 0.f && rx < (float)os.rnx && ry < (float)os.rny) {
            const unsigned c = __ldg(&os.raster[(int)ry * os.rnx + (int)rx]);
            if (c < 2u) return c != 0u;
        }
    }
    const float hx = (float)(os.ncx - 1), hy = (float)(os.ncy - 1);
    float fx0 = floorf(((float)px - 1e-6f - os.gx0) * os.ginv), fx1 = floorf(((float)px + 1e-6f - os.gx0) * os.ginv);
    float fy0 = floorf(((float)py - 1e-6f - os.gy0) * os.ginv), fy1 = floorf(((float)py + 1e-6f - os.gy0) * os.ginv);
    bool occ = false;
    if (fx1 >= 0.f && fx0 <= hx && fy1 >= 0.f && fy0 <= hy) {     // else outside every obstacle
        int cx0 = (int)fmaxf(fx0, 0.f), cx1 = (int)fminf(fx1, hx);
        int cy0 = (int)fmaxf(fy0, 0.f), cy1 = (int)fminf(fy1, hy);
        for (int cy = cy0; cy <= cy1; cy++) {
            int s0 = __ldg(&os.cell_start[cy * os.ncx + cx0]);
            int e0 = __ldg(&os.cell_start[cy * os.ncx + cx1 + 1]);
            for (int i = s0; i < e0; i++) {
                float4 r = __ldg(&os.cell_rects[i]);
                double l = r.x, bo = r.y, rr = r.z, t = r.w;
                occ |= (px >= l && px <= rr && py >= bo && py <= t);
                if (!WANT_NEAR) continue;
                bool in_big = (px >= l - OBS_NEAR_EPS && px <= rr + OBS_NEAR_EPS &&
                               py >= bo - OBS_NEAR_EPS && py <= t + OBS_NEAR_EPS);
                bool in_small = (px >= l + OBS_NEAR_EPS && px <= rr - OBS_NEAR_EPS &&
                                 py >= bo + OBS_NEAR_EPS && py <= t - OBS_NEAR_EPS);
                near |= in_big && !in_small;
            }
        }
    }
    return occ;
}

// ------------------------------------------------- numpy's pairwise summation
//
// np.sum over a float64 column (dwa.py:76-78) runs DOUBLE_pairwise_sum: blocks of <= 128 elements
// are added with eight strided accumulators, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a
// sequential tail; longer inputs are split at n2 = n/2 - (n/2) % 8 and the two halves added.  The
// tree only depends on n, the halves differ by < 16 elements, so all leaves sit at depth Dmin or
// Dmin + 1 (checked on the host for every grid size a DwaP can produce, crlk_make_dwa) and the top
// Dmin levels form a complete binary tree -- which a warp-shuffle butterfly reproduces exactly
// (IEEE addition is commutative; the butterfly pairs exactly the tree's siblings).

__host__ __device__ inline int pw_split(int n) { const int h = n >> 1; return h - (h & 7); }
__host__ __device__ inline int pw_min_depth(int n) { int d = 0; while (n > 128) { n = pw_split(n); d++; } return d; }
__host__ __device__ inline int pw_max_depth(int n) { int d = 0; while (n > 128) { n -= pw_split(n); d++; } return d; }

// a / b rounded to nearest, given y = RN(1 / b): one multiply and two fused corrections instead
// of the division sequence (Markstein: q' = RN(q + (a - b q) y) is the correctly rounded quotient
// when y is the correctly rounded reciprocal and q is within an ulp).  crlk_debug_div compares it
// with the division operator on the device (tests/test_gpu_dwa.py).
__device__ __forceinline__ double div_rn_y(double a, double b, double y)
{
    const double q = a * y;
    const double r = fma(-b, q, a);
    return fma(r, y, q);
}

// ---------------------------------------------------------------- reductions

__device__ __forceinline__ double warp_min_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// lexicographic (value, index) minimum: the FIRST minimum wins (np.argmin).
__device__ __forceinline__ void warp_argmin(double &v, int &i)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}
