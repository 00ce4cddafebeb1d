// Internal declarations shared by the translation units of libcrlk.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/crlk.h"

// Plain-old-data copies of the parameter structs as the kernels see them.
struct EnvP {
    double max_v, min_v, max_w, max_acc_v, max_acc_w, base_dt;
    float rl, rb, rr, rt;   // robot rectangle (float32 values)
    float rc;               // circumradius of the footprint about the body origin
    double radius;          // > 0: circular footprint (CircleRobot, rigid.py:331-347)
};

// Obstacle set on the device: rectangles plus a uniform grid of index lists
// (CSR) used to screen only the cells around a pose.
struct ObsSet {
    const float4 *rects;      // [M] left, bottom, right, top
    const int *cell_start;    // [ncx * ncy + 1]
    const int *cell_items;    // obstacle ids, cell by cell
    const float4 *cell_rects; // rects[cell_items[k]] stored again in cell order: one dependent load less
    int M, ncx, ncy;
    float gx0, gy0, ginv;     // grid origin and 1 / cell size
    // occupancy raster for point queries (same origin, CRLK_RASTER_SUB cells per grid cell and axis):
    // 0 = no obstacle touches the cell, 1 = the obstacles together cover the whole cell, 2 = mixed (test exactly)
    const uint8_t *raster;
    int rnx, rny;
    float rinv;
};

struct DwaP {
    double v_res, w_res, dt, predict_time;
    double g_goal, g_ob, g_speed;
    int skip, use_clearance;
    int n_sub;     // base_dt sub-steps per dt      (differential_env.py:69 loop count)
    int n_outer;   // dt steps per rollout           (dwa.py:39 loop count)
    int nv_max, nw_max;
};

struct ExtP {
    int n_max_extend, n_tree;
    int n_sub_step;        // base_dt sub-steps per executed control step
    double reach_radius;
};

struct CrlkEnv {
    int device;
    CrlkEnvParams host;
    EnvP p;
    float4 *d_rects;  // M obstacles [l, b, r, t]
    int *d_cell_start, *d_cell_items;
    float4 *d_cell_rects;
    uint8_t *d_raster;
    ObsSet os;
    int M;
};

struct CrlkPolicy {
    int device;
    int n_layers;
    int dims[9];
    int kpad[8];          // dims[i] rounded up to 8 (row stride of activations / weights)
    float *d_w[8];        // [dims[i+1]][kpad[i]] row-major, zero padded
    float *d_b[8];
    float low[2], high[2], scale[2], bias[2];
    int max_dim;
    struct CrlkPolicyTc *tc;   // tensor-core path of the hidden layers (NULL if the widths do not tile)
};

struct CrlkEstimator {
    int device;
    float *d_params;      // est params then cls params, each packed (see crlk_nets.cu)
};

void crlk_set_error(const char *fmt, ...);
int crlk_count_loop_lt(double limit, double step);   // t = 0; while (t <  limit) t += step
int crlk_count_loop_le(double limit, double step);   // t = 0; while (t <= limit) t += step
int crlk_make_dwa(const CrlkEnv *env, const CrlkDwaParams *in, DwaP *out);
size_t crlk_dwa_smem_bytes(const DwaP &d);
int crlk_make_ext(const CrlkEnv *env, const CrlkExtendParams *in, ExtP *x);

#define CRLK_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            crlk_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return CRLK_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)

#define CRLK_REQUIRE(cond, msg)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            crlk_set_error("invalid argument: %s", msg);         \
            return CRLK_E_INVALID;                                \
        }                                                         \
    } while (0)

#define CRLK_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            crlk_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return CRLK_E_CUDA;                                                           \
        }                                                                                 \
    } while (0)

// an environment whose last crlk_env_set_obs failed holds no obstacle set: every entry point refuses it
#define CRLK_ENV_OK(env) CRLK_REQUIRE((env) != nullptr && (env)->os.rects != nullptr, \
                                      "env is NULL or holds no obstacle set (crlk_env_set_obs failed)")

int crlk_require_device(void);
int crlk_local_obs_compact(const CrlkEnv *env, const double *states, const double *goals,
                           int32_t goal_stride, int32_t Bmax, const int *list, const int *count,
                           float *obs32, int32_t ld32, void *planes, int32_t Mpad, int32_t Kp, void *stream);

#ifdef __CUDACC__
#include <cuda_bf16.h>
// x = a + b + c exactly (three bf16 planes cover the 24-bit float32 significand)
__device__ __forceinline__ void crlk_split3(float x, __nv_bfloat16 &a, __nv_bfloat16 &b, __nv_bfloat16 &c)
{
    a = __float2bfloat16_rn(x);
    float r1 = x - __bfloat162float(a);
    b = __float2bfloat16_rn(r1);
    float r2 = r1 - __bfloat162float(b);
    c = __float2bfloat16_rn(r2);
}
#endif

// parameter block of the fused RL policy step (k_rl_advance, crlk_kernels.cu)
struct RlAdvP {
    const float *h; int ld_h;                 // last hidden activation, compacted rows
    const float *w, *bias; int K;             // head weights [2][K], bias [2]
    const float *noise; long long noise_stride; const int *noise_cursor; int noise_off, noise_rows; float eps;
    float low[2], high[2], scale[2], hbias[2];
    const double *targets;                    // [Q][5]
    const int *list_in, *count_in;            // rows of this step
    int *list_out, *count_out;                // rows of the next step
    double *state; int *nn;                   // per query: current state, nodes emitted
    double *path; int *n_steps, *status, *node_step, *n_nodes; float *actions; uint8_t *flags;
    float *obs32; int ld32;                   // SIMT path (nullable)
    void *planes_v; int Mpad, Kp;      // tensor-core path (nullable)
    int k, n;                                 // 1-based step, rows capacity
};

int crlk_rl_advance(const CrlkEnv *env, const ExtP *x, const void *params, int first, void *stream);

// tensor-core MLP (crlk_mlp_tc.cu)
struct CrlkPolicyTc;
int crlk_tc_supported(const int *dims, int n_layers);
int crlk_tc_create(const int *dims, int n_layers, const float *const *weights_host, CrlkPolicyTc **out);
void crlk_tc_destroy(CrlkPolicyTc *t);
int crlk_tc_chain_rows(int B);
size_t crlk_tc_workspace_bytes(const CrlkPolicyTc *t, int B);
void crlk_tc_input_planes(const CrlkPolicyTc *t, void *workspace, int B, void **planes, int *Mpad, int *Kp);
int crlk_tc_layers_forward(const CrlkPolicyTc *t, int B, void *workspace, const float *const *d_bias,
                           const int *n_rows_dev, int a_lo_kb, const float **out, int *ld_out, cudaStream_t st);
int crlk_tc_hidden_forward(const CrlkPolicyTc *t, const float *obs, int ld_obs, int B, void *workspace,
                           const float *const *d_bias, const float **out, int *ld_out, cudaStream_t st);
