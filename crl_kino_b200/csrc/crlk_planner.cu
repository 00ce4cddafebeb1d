// libcrlk.so -- device-side bookkeeping of the planner loop (RRT.planning,
// planner/rrt.py:51-89) for Q queries advanced in lock step: sample acceptance,
// node append with path segments, goal test and the steer-to-goal retry choice.
// Compiled with --fmad=false (float64 distance tests must match np.linalg.norm).
#include <math_constants.h>

#include "crlk_device.cuh"

// rrt.py:64-69: a query takes part in this round iff it is not finished, its
// sample is a valid state and the nearest node lies in [d_min, d_max).
__global__ void k_rrt_select(const uint8_t *__restrict__ valid, const double *__restrict__ dist, int Q,
                             int max_iter, double d_min, double d_max, const uint8_t *__restrict__ done,
                             int *iters, uint8_t *active)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    bool a = !done[q] && iters[q] < max_iter && valid[q] && dist[q] >= d_min && dist[q] < d_max;
    active[q] = a ? 1 : 0;
    if (a) iters[q] += 1;
}

extern "C" int crlk_rrt_select(const uint8_t *valid, const double *dist, int32_t Q, int32_t max_iter,
                               double d_min, double d_max, const uint8_t *done, int32_t *iters,
                               uint8_t *active, void *stream)
{
    CRLK_REQUIRE(Q >= 0, "Q < 0");
    if (Q == 0) return CRLK_OK;
    CRLK_REQUIRE(valid && dist && done && iters && active, "NULL array");
    k_rrt_select<<<(Q + 127) / 128, 128, 0, (cudaStream_t)stream>>>(valid, dist, Q, max_iter, d_min, d_max,
                                                                   done, iters, active);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

// One warp per query: append the nodes an extension emitted (rrt.py:125-134),
// copy their path segments into the query's pool, test the goal on the last
// node of the tree (rrt.py:75 / :81) and, in the primary pass, pick the new
// node nearest to the goal for the retry (rrt.py:78-80).
__global__ void __launch_bounds__(256)
k_rrt_append(CrlkForestView f, CrlkExtendView x, const int *__restrict__ parent_idx,
             const uint8_t *__restrict__ active, const double *__restrict__ goals, int Q, int n_max_extend,
             int per, int first_has_start, int primary, int max_iter, double goal_radius, double retry_radius,
             const int *__restrict__ iters, uint8_t *done, uint8_t *reached, uint8_t *retry_active,
             double *retry_from, int *retry_parent, int *overflow)
{
    int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= Q) return;
    if (primary && lane == 0) retry_active[q] = 0;
    if (!active[q] || done[q]) return;
    const int nn = x.n_nodes[q];
    const size_t tb = (size_t)f.stride * q;
    int n0 = f.n_nodes[q];
    if (nn > 0) {
        if (n0 + nn > f.stride) {                   // capacity: report and stop this query
            if (lane == 0) { atomicAdd(overflow, 1); done[q] = 1; }
            return;
        }
        int prev_step = 0;
        int pool_used = f.pool ? f.pool_used[q] : 0;
        for (int k = 0; k < nn; k++) {
            const int step = x.node_step[(size_t)q * per + k];
            const double *src = x.path + ((size_t)q * n_max_extend + (step - 1)) * 5;
            const int n = n0 + k;
            if (lane < 5) f.states[(tb + n) * 5 + lane] = src[lane];
            if (lane == 0) {
                f.x64[tb + n] = src[0]; f.y64[tb + n] = src[1];
                if (f.x32) { f.x32[tb + n] = (float)src[0]; f.y32[tb + n] = (float)src[1]; }
                f.parent[tb + n] = (k == 0) ? parent_idx[q] : n - 1;
            }
            if (f.pool) {
                const int len = step - prev_step;
                if ((long long)pool_used + len > f.pool_cap) {
                    if (lane == 0) { atomicAdd(overflow, 1); done[q] = 1; }
                    return;
                }
                const double *ps = x.path + ((size_t)q * n_max_extend + prev_step) * 5;
                double *pd = f.pool + ((size_t)f.pool_cap * q + pool_used) * 5;
                for (int i = lane; i < len * 5; i += 32) pd[i] = ps[i];
                if (lane == 0) {
                    f.seg_off[tb + n] = pool_used;
                    f.seg_len[tb + n] = len;
                    // does node.path start with the parent's state? (rrt.py:102-104 vs :133)
                    f.seg_flag[tb + n] = (k == 0 && !first_has_start) ? 0 : 1;
                }
                pool_used += len;
            }
            prev_step = step;
        }
        __syncwarp();
        if (lane == 0) {
            f.n_nodes[q] = n0 + nn;
            if (f.pool) f.pool_used[q] = pool_used;
        }
    }
    if (lane != 0) return;
    const double gx = goals[5 * (size_t)q], gy = goals[5 * (size_t)q + 1];
    bool fin = false;
    if (nn > 0 || !primary) {
        // node_list[-1] against the goal (np.linalg.norm: sqrt(fma(dy,dy,dx*dx)))
        int last = n0 + nn - 1;
        double dx = f.x64[tb + last] - gx, dy = f.y64[tb + last] - gy;
        if (nn > 0) {       // the node was written by this warp: read back the source values
            const int step = x.node_step[(size_t)q * per + nn - 1];
            const double *src = x.path + ((size_t)q * n_max_extend + (step - 1)) * 5;
            dx = src[0] - gx; dy = src[1] - gy;
        }
        if (sqrt(fma(dy, dy, dx * dx)) <= goal_radius) { reached[q] = 1; done[q] = 1; fin = true; }
    }
    if (primary == 1 && !fin && nn > 0) {
        // choose_parent(new_node_list, goal): first minimum
        int best = 0;
        double bd = CUDART_INF;
        for (int k = 0; k < nn; k++) {
            const int step = x.node_step[(size_t)q * per + k];
            const double *src = x.path + ((size_t)q * n_max_extend + (step - 1)) * 5;
            double dx = src[0] - gx, dy = src[1] - gy;
            double dd = sqrt(fma(dy, dy, dx * dx));
            if (dd < bd) { bd = dd; best = k; }
        }
        if (bd < retry_radius) {
            const int step = x.node_step[(size_t)q * per + best];
            const double *src = x.path + ((size_t)q * n_max_extend + (step - 1)) * 5;
            for (int c = 0; c < 5; c++) retry_from[5 * (size_t)q + c] = src[c];
            retry_parent[q] = n0 + best;
            retry_active[q] = 1;
        }
    }
    // out of iterations: finished after this round's last pass (primary == 2: crlk_rrt_retry_choice decides)
    if (primary != 2 && !fin && iters[q] >= max_iter && (!primary || !retry_active[q])) done[q] = 1;
}

// Retry decision of a primary pass run with primary == 2: (idx, cost) come from the caller's own
// choose_parent over the nodes the pass appended (rrt.py:78-79 with RRT_RL_Estimator.choose_parent).
__global__ void k_rrt_retry_choice(CrlkForestView f, const int *__restrict__ n_before, const int *__restrict__ idx,
                                   const double *__restrict__ cost, const uint8_t *__restrict__ active, int Q,
                                   double retry_radius, int max_iter, const int *__restrict__ iters, uint8_t *done,
                                   uint8_t *retry_active, double *retry_from, int *retry_parent)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    if (!active[q] || done[q]) return;
    bool retry = f.n_nodes[q] > n_before[q] && cost[q] < retry_radius;
    if (retry) {
        const double *src = f.states + ((size_t)f.stride * q + idx[q]) * 5;
        for (int c = 0; c < 5; c++) retry_from[5 * (size_t)q + c] = src[c];
        retry_parent[q] = idx[q];
        retry_active[q] = 1;
    } else if (iters[q] >= max_iter) {
        done[q] = 1;
    }
}

extern "C" int crlk_rrt_retry_choice(const CrlkForestView *forest, const int32_t *n_before, const int32_t *idx,
                                     const double *cost, const uint8_t *active, int32_t Q, double retry_radius,
                                     int32_t max_iter, const int32_t *iters, uint8_t *done, uint8_t *retry_active,
                                     double *retry_from, int32_t *retry_parent, void *stream)
{
    CRLK_REQUIRE(forest && forest->states && forest->n_nodes, "forest");
    CRLK_REQUIRE(Q >= 0, "Q < 0");
    if (Q == 0) return CRLK_OK;
    CRLK_REQUIRE(n_before && idx && cost && active && iters && done && retry_active && retry_from && retry_parent,
                 "NULL array");
    k_rrt_retry_choice<<<(Q + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*forest, n_before, idx, cost, active, Q,
                                                                         retry_radius, max_iter, iters, done,
                                                                         retry_active, retry_from, retry_parent);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

extern "C" int crlk_rrt_append(const CrlkForestView *forest, const CrlkExtendView *ext,
                               const int32_t *parent_idx, const uint8_t *active, const double *goals,
                               int32_t Q, int32_t n_max_extend, int32_t n_tree, int32_t first_has_start,
                               int32_t primary, int32_t max_iter, double goal_radius, double retry_radius,
                               const int32_t *iters, uint8_t *done, uint8_t *reached, uint8_t *retry_active,
                               double *retry_from, int32_t *retry_parent, int32_t *overflow, void *stream)
{
    CRLK_REQUIRE(forest && ext, "NULL view");
    CRLK_REQUIRE(Q >= 0 && n_tree >= 1 && n_max_extend >= 1, "shape");
    if (Q == 0) return CRLK_OK;
    CRLK_REQUIRE(parent_idx && active && goals && iters && done && reached && overflow, "NULL array");
    CRLK_REQUIRE(!primary || (retry_active && retry_from && retry_parent), "retry outputs required");
    CRLK_REQUIRE(forest->x64 && forest->y64 && forest->states && forest->parent && forest->n_nodes, "forest");
    CRLK_REQUIRE(!forest->pool || (forest->seg_off && forest->seg_len && forest->seg_flag && forest->pool_used),
                 "forest path pool");
    int per = n_max_extend / n_tree + 1;
    k_rrt_append<<<(Q + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        *forest, *ext, parent_idx, active, goals, Q, n_max_extend, per, first_has_start, primary, max_iter,
        goal_radius, retry_radius, iters, done, reached, retry_active, retry_from, retry_parent, overflow);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}
