// libcrlk.so -- learned steering: actor MLP, estimator/classifier CNN and the
// RL extend loop.  Compiled with --fmad=false like the other float64 units: k_rl_step shares the
// float64 dynamics / footprint helpers of crlk_device.cuh with crlk_motion_velocity, crlk_gym_step
// and crlk_valid_state, and must round exactly like them.  The float32 network code fuses
// multiply-adds explicitly (fmaf); its parity is tolerance-based.
#include <math_constants.h>
#include <stdlib.h>

#include "crlk_device.cuh"

// ---------------------------------------------------------------- policy MLP

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

extern "C" int crlk_policy_destroy(CrlkPolicy *p);

extern "C" int crlk_policy_create(int32_t n_layers, const int32_t *dims,
                                  const float *const *weights_host, const float *const *biases_host,
                                  const float *low, const float *high, CrlkPolicy **out)
{
    CRLK_REQUIRE(out && dims && weights_host && biases_host && low && high, "NULL argument");
    CRLK_REQUIRE(n_layers >= 2 && n_layers <= 8, "n_layers must be in [2, 8]");
    CRLK_REQUIRE(dims[n_layers] == 2, "the actor must output 2 actions (v, omega)");
    int rc = crlk_require_device();
    if (rc) return rc;
    CrlkPolicy *p = new CrlkPolicy();
    memset(p, 0, sizeof *p);
    // built inside a lambda: on any failure the partly built handle is destroyed, not leaked
    auto build = [&]() -> int {
        CRLK_CUDA(cudaGetDevice(&p->device));
        p->n_layers = n_layers;
        p->max_dim = 0;
        for (int i = 0; i <= n_layers; i++) {
            CRLK_REQUIRE(dims[i] >= 1, "layer width < 1");
            p->dims[i] = dims[i];
        }
        for (int i = 0; i < n_layers; i++) {
            int K = dims[i], N = dims[i + 1];
            int kp = round_up(K, 16);
            p->kpad[i] = kp;
            if (i > 0 && kp > p->max_dim) p->max_dim = kp;
            float *hw = (float *)calloc((size_t)N * kp, sizeof(float));
            CRLK_REQUIRE(hw != nullptr, "out of host memory");
            for (int n = 0; n < N; n++) memcpy(hw + (size_t)n * kp, weights_host[i] + (size_t)n * K, sizeof(float) * K);
            cudaError_t e1 = cudaMalloc(&p->d_w[i], sizeof(float) * (size_t)N * kp);
            cudaError_t e2 = e1 == cudaSuccess ? cudaMemcpy(p->d_w[i], hw, sizeof(float) * (size_t)N * kp, cudaMemcpyHostToDevice) : e1;
            free(hw);
            CRLK_CUDA(e2);
            CRLK_CUDA(cudaMalloc(&p->d_b[i], sizeof(float) * N));
            CRLK_CUDA(cudaMemcpy(p->d_b[i], biases_host[i], sizeof(float) * N, cudaMemcpyHostToDevice));
        }
        for (int k = 0; k < 2; k++) {
            // net.py:134-138 in float32
            p->low[k] = low[k]; p->high[k] = high[k];
            p->bias[k] = (low[k] + high[k]) / 2.0f;
            p->scale[k] = (high[k] - low[k]) / 2.0f;
        }
        p->tc = nullptr;
        if (crlk_tc_supported(p->dims, n_layers)) {
            int rc2 = crlk_tc_create(p->dims, n_layers, weights_host, &p->tc);
            if (rc2) return rc2;
        }
        return CRLK_OK;
    };
    rc = build();
    if (rc) {
        p->n_layers = n_layers;          // destroy walks every layer slot (unset ones are NULL)
        crlk_policy_destroy(p);
        return rc;
    }
    *out = p;
    return CRLK_OK;
}

extern "C" int crlk_policy_destroy(CrlkPolicy *p)
{
    if (!p) return CRLK_OK;
    for (int i = 0; i < p->n_layers; i++) {
        if (p->d_w[i]) cudaFree(p->d_w[i]);
        if (p->d_b[i]) cudaFree(p->d_b[i]);
    }
    crlk_tc_destroy(p->tc);
    delete p;
    return CRLK_OK;
}

extern "C" int64_t crlk_policy_workspace_bytes(const CrlkPolicy *p, int32_t B)
{
    if (!p || B <= 0) return 0;
    int64_t simt = (int64_t)sizeof(float) * 2 * (int64_t)round_up(B, 64) * p->max_dim + 256;
    int64_t tc = p->tc ? (int64_t)crlk_tc_workspace_bytes(p->tc, B) : 0;
    return simt > tc ? simt : tc;
}

// C[B, N] = relu(A[B, lda] . W[N, K]^T + b), K % 16 == 0, rows of A/W 16-byte aligned.
// 64 x 64 output tile, 16-deep k slices, 4 x 4 register block per thread.
#define GM_BM 64
#define GM_BN 64
#define GM_BK 16
__global__ void __launch_bounds__(256)
k_linear_relu(const float *__restrict__ A, int lda, const float *__restrict__ W, int K,
              const float *__restrict__ bias, float *__restrict__ Cout, int ldc, int B, int N,
              const int *__restrict__ n_rows_dev)
{
    __shared__ float As[GM_BK][GM_BM + 4];
    __shared__ float Ws[GM_BK][GM_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GM_BM, n0 = blockIdx.x * GM_BN;
    if (n_rows_dev) B = min(B, *n_rows_dev);       // compacted batch: live rows counted on the device
    if (m0 >= B) return;
    const int tr = tid / 16, tc = tid % 16;        // 16 x 16 threads, each 4 x 4 outputs
    const int lr = tid / 4, lk = (tid % 4) * 4;    // loader: row lr, k offset lk
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += GM_BK) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = a;
        if (m0 + lr < B) a = *reinterpret_cast<const float4 *>(A + (size_t)(m0 + lr) * lda + k0 + lk);
        if (n0 + lr < N) w = __ldg(reinterpret_cast<const float4 *>(W + (size_t)(n0 + lr) * K + k0 + lk));
        As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
        Ws[lk + 0][lr] = w.x; Ws[lk + 1][lr] = w.y; Ws[lk + 2][lr] = w.z; Ws[lk + 3][lr] = w.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GM_BK; k++) {
            float4 av = *reinterpret_cast<const float4 *>(&As[k][tr * 4]);
            float4 wv = *reinterpret_cast<const float4 *>(&Ws[k][tc * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int m = m0 + tr * 4 + i;
        if (m >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int n = n0 + tc * 4 + j;
            if (n < N) Cout[(size_t)m * ldc + n] = fmaxf(acc[i][j] + __ldg(&bias[n]), 0.f);
        }
    }
}

struct HeadP { float low[2], high[2], scale[2], bias[2]; };

static bool policy_uses_tc(const CrlkPolicy *p)
{
    static const bool force_simt = getenv("CRLK_MLP_SIMT") != nullptr;   // A/B switch for the tests
    return p->tc && !force_simt;
}

// Last layer (K -> 2) + tanh + exploration noise + scale/bias + clamp
// (net.py:145-153, ddpg.py:126-130).  One warp per row.
__global__ void __launch_bounds__(256)
k_policy_head(const float *__restrict__ A, int lda, const float *__restrict__ W, int K,
              const float *__restrict__ bias, const float *__restrict__ noise, long long noise_stride,
              const int *__restrict__ noise_cursor, int noise_off, int noise_rows, float eps, HeadP hp, int B,
              float *__restrict__ act, const int *__restrict__ n_rows_dev, const int *__restrict__ list)
{
    int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    if (n_rows_dev && row >= *n_rows_dev) return;
    const int q = list ? list[row] : row;          // the query this (compacted) row belongs to
    const float *a = A + (size_t)row * lda;
    float s0 = 0.f, s1 = 0.f;
    for (int k = lane; k < K; k += 32) {
        float x = a[k];
        s0 = fmaf(x, __ldg(&W[k]), s0);
        s1 = fmaf(x, __ldg(&W[K + k]), s1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane < 2) {
        float z = (lane == 0 ? s0 : s1) + __ldg(&bias[lane]);
        z = tanhf(z);
        if (noise) {
            // draws past the end of a query's stream count as zero (the host sizes the stream and
            // checks the cursors, planner/batched.py; this only keeps the read inside the buffer)
            long long off = (long long)noise_off + (noise_cursor ? noise_cursor[q] : 0);
            if (off < noise_rows) z = z + noise[(size_t)q * noise_stride + off * 2 + lane] * eps;
        }
        z = z * hp.scale[lane] + hp.bias[lane];
        z = fminf(fmaxf(z, hp.low[lane]), hp.high[lane]);
        act[2 * (size_t)row + lane] = z;
    }
}

// Hidden layers + head.  Operand source: float32 `obs` rows, or (obs == NULL) bf16
// operand planes already placed in the workspace (crlk_tc_input_planes; a_lo_kb as
// in crlk_tc_layers_forward).  n_rows_dev / list (nullable): compacted batch of the
// RL steer loop -- live row count on the device and the query of each row.
// Hidden layers: the last hidden activation (float32 rows) through *out / *ld_out.
static int policy_hidden_impl(const CrlkPolicy *p, const float *obs, int ld_obs, int B, void *workspace,
                              cudaStream_t st, const int *n_rows_dev, int a_lo_kb, const float **out, int *ld_out)
{
    const int L = p->n_layers;
    const float *in = obs;
    int ld_in = ld_obs;
    if (policy_uses_tc(p)) {
        // hidden layers on the tensor cores (tcgen05, bf16 split operands, float32 accumulate)
        int rc = obs ? crlk_tc_hidden_forward(p->tc, obs, ld_obs, B, workspace, p->d_b, &in, &ld_in, st)
                     : crlk_tc_layers_forward(p->tc, B, workspace, p->d_b, n_rows_dev, a_lo_kb, &in, &ld_in, st);
        if (rc) return rc;
    } else {
        CRLK_REQUIRE(obs != nullptr, "float32 observations required on the SIMT path");
        float *buf[2];
        size_t half = (size_t)round_up(B, 64) * p->max_dim;
        buf[0] = (float *)workspace;
        buf[1] = buf[0] + half;
        for (int i = 0; i < L - 1; i++) {
            int N = p->dims[i + 1];
            int ldc = p->kpad[i + 1];
            float *outb = buf[i & 1];
            dim3 grid((N + GM_BN - 1) / GM_BN, (B + GM_BM - 1) / GM_BM);
            k_linear_relu<<<grid, 256, 0, st>>>(in, ld_in, p->d_w[i], p->kpad[i], p->d_b[i], outb, ldc, B, N,
                                               n_rows_dev);
            CRLK_LAUNCH_CHECK();
            in = outb;
            ld_in = ldc;
        }
    }
    *out = in;
    *ld_out = ld_in;
    return CRLK_OK;
}

static int policy_forward_impl(const CrlkPolicy *p, const float *obs, int ld_obs, const float *noise,
                               long long noise_stride, const int *noise_cursor, int noise_off, int noise_rows,
                               float eps, int B,
                               float *act, void *workspace, cudaStream_t st, const int *n_rows_dev = nullptr,
                               const int *list = nullptr, int a_lo_kb = 0)
{
    const int L = p->n_layers;
    const float *in = nullptr;
    int ld_in = 0;
    int rc = policy_hidden_impl(p, obs, ld_obs, B, workspace, st, n_rows_dev, a_lo_kb, &in, &ld_in);
    if (rc) return rc;
    HeadP hp;
    for (int k = 0; k < 2; k++) {
        hp.low[k] = p->low[k]; hp.high[k] = p->high[k]; hp.scale[k] = p->scale[k]; hp.bias[k] = p->bias[k];
    }
    k_policy_head<<<(B + 7) / 8, 256, 0, st>>>(in, ld_in, p->d_w[L - 1], p->kpad[L - 1], p->d_b[L - 1],
                                              noise, noise_stride, noise_cursor, noise_off, noise_rows, eps, hp, B,
                                              act, n_rows_dev, list);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

extern "C" int crlk_policy_forward(const CrlkPolicy *p, const float *obs, int32_t ld_obs,
                                   const float *noise, float eps, int32_t B, float *act,
                                   void *workspace, void *stream)
{
    CRLK_REQUIRE(p != nullptr, "policy is NULL");
    CRLK_REQUIRE(B >= 0, "B < 0");
    if (B == 0) return CRLK_OK;
    CRLK_REQUIRE(obs && act && workspace, "NULL array");
    CRLK_REQUIRE(ld_obs >= p->kpad[0] && ld_obs % 4 == 0,
                 "ld_obs must be >= the input width rounded up to 16, zero padded");
    CRLK_REQUIRE(((uintptr_t)obs & 15) == 0 && ((uintptr_t)workspace & 15) == 0, "16-byte alignment");
    return policy_forward_impl(p, obs, ld_obs, noise, 2, nullptr, 0, 1, eps, B, act, workspace, (cudaStream_t)stream);
}

// ------------------------------------------------------- estimator CNNs (R11)

#define CNN_PARAMS 24452
#define CNN_THREADS 128
// packed parameter offsets of one net
#define O_W1 0
#define O_B1 72
#define O_W2 80
#define O_B2 1232
#define O_W3 1248
#define O_B3 5856
#define O_W4 5888
#define O_B4 24320
#define O_WL 24384
#define O_BL 24451

extern "C" int crlk_estimator_create(const float *const *est_params_host,
                                     const float *const *cls_params_host, CrlkEstimator **out)
{
    CRLK_REQUIRE(est_params_host && cls_params_host && out, "NULL argument");
    int rc = crlk_require_device();
    if (rc) return rc;
    static const int sizes[10] = {72, 8, 1152, 16, 4608, 32, 18432, 64, 67, 1};
    float *h = (float *)malloc(sizeof(float) * 2 * CNN_PARAMS);
    for (int net = 0; net < 2; net++) {
        const float *const *src = net ? cls_params_host : est_params_host;
        size_t off = (size_t)net * CNN_PARAMS;
        for (int i = 0; i < 10; i++) {
            if (!src[i]) { free(h); crlk_set_error("invalid argument: NULL parameter tensor"); return CRLK_E_INVALID; }
            memcpy(h + off, src[i], sizeof(float) * sizes[i]);
            off += sizes[i];
        }
    }
    // conv weights [co][ci][3x3] -> [ci][3x3][co]: the threads of a warp work on neighbouring output
    // channels, so each weight load becomes one cache line instead of one line per lane
    {
        static const int w_off[4] = {O_W1, O_W2, O_W3, O_W4}, w_ci[4] = {1, 8, 16, 32}, w_co[4] = {8, 16, 32, 64};
        float *tmp = (float *)malloc(sizeof(float) * 64 * 32 * 9);
        for (int net = 0; net < 2; net++)
            for (int l = 0; l < 4; l++) {
                float *W = h + (size_t)net * CNN_PARAMS + w_off[l];
                const int CI = w_ci[l], CO = w_co[l];
                memcpy(tmp, W, sizeof(float) * CO * CI * 9);
                for (int co = 0; co < CO; co++)
                    for (int ci = 0; ci < CI; ci++)
                        for (int k = 0; k < 9; k++) W[(ci * 9 + k) * CO + co] = tmp[(co * CI + ci) * 9 + k];
            }
        free(tmp);
    }
    CrlkEstimator *e = new CrlkEstimator();
    CRLK_CUDA(cudaGetDevice(&e->device));
    CRLK_CUDA(cudaMalloc(&e->d_params, sizeof(float) * 2 * CNN_PARAMS));
    CRLK_CUDA(cudaMemcpy(e->d_params, h, sizeof(float) * 2 * CNN_PARAMS, cudaMemcpyHostToDevice));
    free(h);
    *out = e;
    return CRLK_OK;
}

extern "C" int crlk_estimator_destroy(CrlkEstimator *e)
{
    if (!e) return CRLK_OK;
    if (e->d_params) cudaFree(e->d_params);
    delete e;
    return CRLK_OK;
}

// Activations of one sample in shared memory, every map stored with a one-pixel zero border so
// that the 3x3 taps need no bounds tests (a zero tap adds exactly nothing: fma(0, w, acc) == acc).
#define CNN_IN_W 29      // 27 + 2
#define CNN_A_W 15       // 13 + 2
#define CNN_B_W 8        //  6 + 2
#define CNN_C_W 5        //  3 + 2
struct CnnSmem {
    float in[CNN_IN_W * CNN_IN_W];
    float a[8 * CNN_A_W * CNN_A_W];
    float b[16 * CNN_B_W * CNN_B_W];
    float c[32 * CNN_C_W * CNN_C_W];
    float d[64];
};

__device__ __forceinline__ void cnn_zero(CnnSmem &s)
{
    float *p = reinterpret_cast<float *>(&s);
    for (int i = threadIdx.x; i < (int)(sizeof(CnnSmem) / sizeof(float)); i += CNN_THREADS) p[i] = 0.f;
}

// conv3x3(pad 1) + ReLU + maxpool2 (floor) from the bordered map in[CI][HP][HP] to the bordered map
// out[CO][OP][OP].  One task = (output channel, pooled row): the thread keeps the 2 x XT
// pre-pool outputs of that row pair in registers and walks the input channels, loading each
// channel's 4 x (XT + 2) input window and 9 weights once (216 FMA per 65 loads for conv2) --
// instead of one load pair per FMA and a single dependent accumulator.  Per accumulator the
// additions run in the same order as before (bias, then channel, ky, kx ascending).
template <int CI, int CO, int HP, int XT, int OP>
__device__ __forceinline__ void conv_pool_rows(const float *__restrict__ w, const float *__restrict__ bias,
                                               const float *in, float *out)
{
    constexpr int NPR = XT / 2;                    // pooled rows = pooled columns
    for (int task = threadIdx.x; task < CO * NPR; task += CNN_THREADS) {
        const int co = task / NPR, pr = task - co * NPR;
        float acc[2][XT];
        const float b0 = __ldg(&bias[co]);
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int x = 0; x < XT; x++) acc[r][x] = b0;
        for (int ci = 0; ci < CI; ci++) {
            const float *wp = w + (size_t)ci * 9 * CO + co;          // weights stored [ci][3x3][co]
            float wv[9];
#pragma unroll
            for (int k = 0; k < 9; k++) wv[k] = __ldg(&wp[k * CO]);
            const float *ip = in + ci * HP * HP + (2 * pr) * HP;     // bordered rows 2pr .. 2pr + 3
#pragma unroll
            for (int r = 0; r < 2; r++) {
#pragma unroll
                for (int ky = 0; ky < 3; ky++) {
                    float row[XT + 2];
#pragma unroll
                    for (int x = 0; x < XT + 2; x++) row[x] = ip[(r + ky) * HP + x];
#pragma unroll
                    for (int x = 0; x < XT; x++)
#pragma unroll
                        for (int kx = 0; kx < 3; kx++) acc[r][x] = fmaf(row[x + kx], wv[ky * 3 + kx], acc[r][x]);
                }
            }
        }
        float *op = out + co * OP * OP + (pr + 1) * OP + 1;
#pragma unroll
        for (int px = 0; px < NPR; px++) {
            float m = fmaxf(fmaxf(acc[0][2 * px], acc[0][2 * px + 1]), fmaxf(acc[1][2 * px], acc[1][2 * px + 1]));
            op[px] = fmaxf(m, 0.f);
        }
    }
}

// conv4 (32 -> 64, 3x3, pad 1) + ReLU + global average pool over the 3 x 3 map: one output channel
// per thread, nine accumulators.
__device__ __forceinline__ void conv4_avgpool(const float *__restrict__ P, const float *s_c, float *s_d)
{
    for (int co = threadIdx.x; co < 64; co += CNN_THREADS) {
        float acc[9];
        const float b0 = __ldg(&P[O_B4 + co]);
#pragma unroll
        for (int k = 0; k < 9; k++) acc[k] = b0;
        for (int ci = 0; ci < 32; ci++) {
            const float *wp = P + O_W4 + (size_t)ci * 9 * 64 + co;   // weights stored [ci][3x3][co]
            float wv[9], iv[CNN_C_W * CNN_C_W];
#pragma unroll
            for (int k = 0; k < 9; k++) wv[k] = __ldg(&wp[k * 64]);
#pragma unroll
            for (int k = 0; k < CNN_C_W * CNN_C_W; k++) iv[k] = s_c[ci * CNN_C_W * CNN_C_W + k];
#pragma unroll
            for (int y = 0; y < 3; y++)
#pragma unroll
                for (int x = 0; x < 3; x++)
#pragma unroll
                    for (int ky = 0; ky < 3; ky++)
#pragma unroll
                        for (int kx = 0; kx < 3; kx++)
                            acc[y * 3 + x] = fmaf(iv[(y + ky) * CNN_C_W + x + kx], wv[ky * 3 + kx], acc[y * 3 + x]);
        }
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 9; k++) sum += fmaxf(acc[k], 0.f);
        s_d[co] = sum / 9.0f;
    }
}

// All four convolution stages of one net over the sample in s.in; features land in s.d.
__device__ __forceinline__ void cnn_features(const float *__restrict__ P, CnnSmem &s)
{
    conv_pool_rows<1, 8, CNN_IN_W, 26, CNN_A_W>(P + O_W1, P + O_B1, s.in, s.a);     // 27 -> 13
    __syncthreads();
    conv_pool_rows<8, 16, CNN_A_W, 12, CNN_B_W>(P + O_W2, P + O_B2, s.a, s.b);      // 13 -> 6
    __syncthreads();
    conv_pool_rows<16, 32, CNN_B_W, 6, CNN_C_W>(P + O_W3, P + O_B3, s.b, s.c);      //  6 -> 3
    __syncthreads();
    conv4_avgpool(P, s.c, s.d);
    __syncthreads();
}

// Linear(67 -> 1) over [features(64), ext(3)] by the first warp; result in lane 0.
__device__ __forceinline__ float cnn_head(const float *__restrict__ P, const float *s_d, float e0, float e1,
                                          float e2)
{
    int lane = threadIdx.x;
    float acc = 0.f;
    for (int k = lane; k < 64; k += 32) acc = fmaf(s_d[k], __ldg(&P[O_WL + k]), acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    acc = fmaf(e0, __ldg(&P[O_WL + 64]), acc);
    acc = fmaf(e1, __ldg(&P[O_WL + 65]), acc);
    acc = fmaf(e2, __ldg(&P[O_WL + 66]), acc);
    return acc + __ldg(&P[O_BL]);
}

// One CTA per (sample, net): estimator/network.py:39-69 and :97-127.
__global__ void __launch_bounds__(CNN_THREADS)
k_estimator(const float *__restrict__ params, const float *__restrict__ obs, int ld_obs, int B,
            float *__restrict__ est, float *__restrict__ prob)
{
    __shared__ CnnSmem sm;
    const int b = blockIdx.x, net = blockIdx.y;
    if (b >= B) return;
    const float *P = params + (size_t)net * CNN_PARAMS;
    const float *o = obs + (size_t)b * ld_obs;
    cnn_zero(sm);
    __syncthreads();
    for (int i = threadIdx.x; i < 729; i += CNN_THREADS) {
        int y = i / 27, x = i - y * 27;
        sm.in[(y + 1) * CNN_IN_W + x + 1] = o[4 + i];
    }
    __syncthreads();
    cnn_features(P, sm);
    const float *s_d = sm.d;
    if (threadIdx.x < 32) {
        // ext = [||goal_body||, v, omega] (rrt_rl_estimator.py:59-62), float32
        float gx = o[0], gy = o[1];
        float e0 = sqrtf(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
        float acc = cnn_head(P, s_d, e0, o[2], o[3]);
        if (threadIdx.x == 0) {
            if (net == 0) est[b] = acc;
            else prob[b] = 1.0f / (1.0f + expf(-acc));
        }
    }
}

extern "C" int crlk_estimator_forward(const CrlkEstimator *e, const float *obs, int32_t ld_obs,
                                      int32_t B, float *est, float *prob, void *stream)
{
    CRLK_REQUIRE(e != nullptr, "estimator is NULL");
    CRLK_REQUIRE(B >= 0, "B < 0");
    if (B == 0) return CRLK_OK;
    CRLK_REQUIRE(obs && est && prob, "NULL array");
    CRLK_REQUIRE(ld_obs >= 733, "ld_obs < 733");
    dim3 grid(B, 2);
    k_estimator<<<grid, CNN_THREADS, 0, (cudaStream_t)stream>>>(e->d_params, obs, ld_obs, B, est, prob);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}


// ---- RRT_RL_Estimator.choose_parent on the device (rrt_rl_estimator.py:26-79) ----

// per node: Euclidean distance to the query's target and the [1, 20) filter (:29-31)
__global__ void k_cp_dist(const double *__restrict__ states, long long stride, const int *__restrict__ n_nodes,
                          const int *__restrict__ first_node, int max_nodes, const double *__restrict__ targets,
                          int Q, double *dist, int *skip)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)Q * max_nodes) return;
    int q = (int)(i / max_nodes), n = (int)(i - (long long)q * max_nodes);
    if (n >= n_nodes[q] || (first_node && n < first_node[q])) { skip[i] = 1; dist[i] = CUDART_INF; return; }
    const double *s = states + ((size_t)stride * q + n) * 5;
    double dx = s[0] - targets[5 * (size_t)q], dy = s[1] - targets[5 * (size_t)q + 1];
    double dd = sqrt(fma(dy, dy, __dmul_rn(dx, dx)));            // np.linalg.norm
    dist[i] = dd;
    skip[i] = (dd >= 1.0 && dd < 20.0) ? 0 : 1;
}

// One CTA per candidate node: local observation (differential_gym.py:180-190)
// straight into shared memory, then both CNNs (estimator/network.py).
__global__ void __launch_bounds__(CNN_THREADS)
k_estimator_nodes(ObsGrid og, ObsSet os, const float *__restrict__ params, const double *__restrict__ states,
                  long long stride, int max_nodes, const double *__restrict__ targets,
                  const int *__restrict__ skip, long long rows, float *__restrict__ est, float *__restrict__ prob)
{
    __shared__ CnnSmem sm;
    __shared__ float s_ext[4];
    const long long row = blockIdx.x;
    if (row >= rows || skip[row]) return;
    cnn_zero(sm);
    __syncthreads();
    const int q = (int)(row / max_nodes), nidx = (int)(row - (long long)q * max_nodes);
    const double *st = states + ((size_t)stride * q + nidx) * 5;
    const double x = st[0], y = st[1];
    double sn, cs;
    crlk_sincos(st[2], &sn, &cs);
    bool nearp = false;
    for (int p = threadIdx.x; p < OBS_NPTS; p += CNN_THREADS) {
        int ib = p / OBS_SIDE, il = p - ib * OBS_SIDE;
        double px, py;
        body_to_world(cs, sn, x, y, arange_at(og.ba0, og.ba1, og.bad, ib), arange_at(og.lr0, og.lr1, og.lrd, il),
                      px, py);
        sm.in[(ib + 1) * CNN_IN_W + il + 1] = local_map_occupied<false>(os, px, py, nearp) ? 0.f : 1.f;
    }
    if (threadIdx.x == 0) {
        double qx = __dadd_rn(targets[5 * (size_t)q], -x), qy = __dadd_rn(targets[5 * (size_t)q + 1], -y);
        float gbx = (float)__dadd_rn(__dmul_rn(cs, qx), __dmul_rn(sn, qy));
        float gby = (float)__dadd_rn(__dmul_rn(-sn, qx), __dmul_rn(cs, qy));
        s_ext[0] = sqrtf(__fadd_rn(__fmul_rn(gbx, gbx), __fmul_rn(gby, gby)));   // torch.norm in float32 (:59)
        s_ext[1] = (float)st[3];
        s_ext[2] = (float)st[4];
    }
    __syncthreads();
    for (int net = 0; net < 2; net++) {
        const float *P = params + (size_t)net * CNN_PARAMS;
        cnn_features(P, sm);
        if (threadIdx.x < 32) {
            float acc = cnn_head(P, sm.d, s_ext[0], s_ext[1], s_ext[2]);
            if (threadIdx.x == 0) {
                if (net == 0) est[row] = acc;
                else prob[row] = 1.0f / (1.0f + expf(-acc));
            }
        }
        __syncthreads();
    }
}

// per query: cost = (est + 1) * 15 + 100 * (1 - round(prob)) on the filtered nodes (float32
// arithmetic, :70-74), the plain distance elsewhere; first minimum (:76); (0, 100) if no
// node passed the filter (:33-34).
__global__ void __launch_bounds__(256)
k_cp_final(const double *__restrict__ dist, const int *__restrict__ skip, const float *__restrict__ est,
           const float *__restrict__ prob, const int *__restrict__ n_nodes, const int *__restrict__ first_node,
           int max_nodes, int Q, int *idx, double *cost)
{
    __shared__ double sv[8];
    __shared__ int si[8], sany[8];
    const int q = blockIdx.x;
    if (q >= Q) return;
    const int n = n_nodes[q];
    double best = CUDART_INF;
    int bi = 0x7fffffff, any = 0;
    for (int k = threadIdx.x; k < n; k += 256) {
        size_t i = (size_t)q * max_nodes + k;
        double c = dist[i];
        if (!skip[i]) {
            float fail = 1.0f - rintf(prob[i]);
            c = (double)__fmul_rn(__fadd_rn(est[i], 1.0f), 15.0f);
            c = c + (double)__fmul_rn(100.0f, fail);
            any = 1;
        }
        if (c < best || (c == best && k < bi)) { best = c; bi = k; }
    }
    warp_argmin(best, bi);
    any = __any_sync(0xffffffffu, any);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sv[warp] = best; si[warp] = bi; sany[warp] = any; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) {
            if (sv[w] < best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
            any |= sany[w];
        }
        // no candidate in the band: element 0 of the candidate list, cost 100 (:33-34)
        const int f0 = first_node ? first_node[q] : 0;
        if (!any) { idx[q] = (f0 < n) ? f0 : 0; cost[q] = 100.0; }
        else { idx[q] = bi; cost[q] = best; }
    }
}

extern "C" int64_t crlk_estimator_cp_workspace_bytes(int32_t Q, int32_t max_nodes)
{
    if (Q <= 0 || max_nodes <= 0) return 0;
    int64_t rows = (int64_t)Q * max_nodes;
    return rows * (8 + 4 + 4 + 4) + 1024;
}

extern "C" int crlk_estimator_choose_parent(const CrlkEnv *env, const CrlkEstimator *e, const double *states,
                                            int64_t stride, const int32_t *n_nodes, const int32_t *first_node,
                                            int32_t max_nodes, const double *targets, int32_t Q, int32_t *idx,
                                            double *cost, void *workspace, void *stream)
{
    CRLK_ENV_OK(env);
    CRLK_REQUIRE(e != nullptr, "estimator is NULL");
    CRLK_REQUIRE(Q >= 0 && max_nodes >= 1 && stride >= max_nodes, "shape");
    if (Q == 0) return CRLK_OK;
    CRLK_REQUIRE(states && n_nodes && targets && idx && cost && workspace, "NULL array");
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)Q * max_nodes;
    CRLK_REQUIRE(rows < 2147483647LL, "too many candidate nodes for one call");
    unsigned char *ws = (unsigned char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    double *dist = (double *)ws;
    int *skip = (int *)(ws + rows * 8);
    float *est = (float *)(ws + rows * 12);
    float *prob = (float *)(ws + rows * 16);
    k_cp_dist<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(states, (long long)stride, n_nodes, first_node,
                                                            max_nodes, targets, Q, dist, skip);
    CRLK_LAUNCH_CHECK();
    k_estimator_nodes<<<(unsigned)rows, CNN_THREADS, 0, st>>>(make_obs_grid(), env->os, e->d_params, states,
                                                             (long long)stride, max_nodes, targets, skip, rows,
                                                             est, prob);
    CRLK_LAUNCH_CHECK();
    k_cp_final<<<Q, 256, 0, st>>>(dist, skip, est, prob, n_nodes, first_node, max_nodes, Q, idx, cost);
    CRLK_LAUNCH_CHECK();
    return CRLK_OK;
}

// ------------------------------------------------ RL extend (rrt_rl.py:22-63)


struct RlWs {
    double *state;     // [Q][5]
    int *nn;           // [Q]
    int *list[2];      // [Q] queries still extending (ping-pong between steps)
    int *count;        // [n_max_extend + 2] live rows per step
    float *obs;        // [Q][736] (SIMT path only)
    float *act;        // [Q][2], compacted rows
    void *mlp;         // policy workspace
};

static size_t rl_carve(const CrlkPolicy *p, int Q, unsigned char *base, RlWs *w)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    size_t o_state = take(sizeof(double) * 5 * Q);
    size_t o_nn = take(sizeof(int) * Q);
    size_t o_l0 = take(sizeof(int) * Q);
    size_t o_l1 = take(sizeof(int) * Q);
    size_t o_cnt = take(sizeof(int) * 4096);
    size_t o_obs = take(policy_uses_tc(p) ? 0 : sizeof(float) * 736 * (size_t)Q);
    size_t o_act = take(sizeof(float) * 2 * Q);
    size_t o_mlp = take((size_t)crlk_policy_workspace_bytes(p, Q));
    if (w) {
        w->state = (double *)(base + o_state); w->nn = (int *)(base + o_nn);
        w->list[0] = (int *)(base + o_l0); w->list[1] = (int *)(base + o_l1);
        w->count = (int *)(base + o_cnt);
        w->obs = (float *)(base + o_obs);
        w->act = (float *)(base + o_act); w->mlp = base + o_mlp;
    }
    return off;
}

// Large batches are run as 2 or 4 slices on as many streams (see crlk_extend_rl).
#define RL_MAX_SPLIT 4
static int rl_splits(int Q, bool tc)
{
    static const char *env = getenv("CRLK_RL_SPLIT");
    if (env) { int v = atoi(env); return v >= 4 ? 4 : v >= 2 ? 2 : 1; }
    // With the hidden layers in one dataflow launch (k_mlp_chain) a single slice is fastest: two chains
    // cannot share an SM, and half-size batches fill it worse (7.6 ms in one slice, 8.9 ms in two at 8192
    // queries).  With one launch per layer two slices overlap one slice's GEMMs with the other's k_rl_advance.
    if (tc && crlk_tc_chain_rows(Q)) return 1;
    return Q >= 2048 ? 2 : 1;
}

static void rl_slice(int Q, int ns, int i, int *q0, int *n)
{
    const int base = (Q + ns - 1) / ns;
    *q0 = i * base < Q ? i * base : Q;
    *n = (*q0 + base <= Q) ? base : Q - *q0;
}

extern "C" int64_t crlk_extend_rl_workspace_bytes(const CrlkPolicy *p, int32_t Q)
{
    if (!p || Q <= 0) return 0;
    int64_t best = (int64_t)rl_carve(p, Q, nullptr, nullptr);
    for (int ns = 2; ns <= RL_MAX_SPLIT; ns *= 2) {
        int64_t v = (int64_t)ns * ((int64_t)rl_carve(p, (Q + ns - 1) / ns, nullptr, nullptr) + 256) + 512;
        if (v > best) best = v;
    }
    return best;
}

// Kernels one crlk_extend_rl call launches: per slice k_rl_init + the first observation, then per policy
// step and slice the hidden layers (one dataflow launch, or one per layer) + k_rl_advance.
extern "C" int32_t crlk_extend_rl_launches(const CrlkPolicy *p, const CrlkExtendParams *ext, int32_t Q,
                                           int32_t with_noise_cursor)
{
    if (!p || !ext || Q <= 0) return 0;
    const bool tc = policy_uses_tc(p);
    const int ns = rl_splits(Q, tc);
    int slices = 0, per_step = 0;
    for (int i = 0; i < ns; i++) {
        int q0, n;
        rl_slice(Q, ns, i, &q0, &n);
        if (n <= 0) continue;
        slices++;
        per_step += ((tc && crlk_tc_chain_rows(n)) ? 1 : p->n_layers - 1) + 1;
    }
    return 2 * slices + ext->n_max_extend * per_step + (with_noise_cursor ? 1 : 0);
}

// count[] is zeroed before this launch; the active queries become step 1's row list.
__global__ void k_rl_init(const double *__restrict__ from_states, const uint8_t *__restrict__ active,
                          int Q, double *state, int *nn, int *list0, int *count, int *n_steps, int *status,
                          int *n_nodes, uint8_t *flags)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    for (int c = 0; c < 5; c++) state[5 * (size_t)q + c] = from_states[5 * (size_t)q + c];
    nn[q] = 0;
    n_steps[q] = 0; status[q] = CRLK_EXT_TIMEOUT; n_nodes[q] = 0;
    if (flags) flags[q] = 0;
    if (!active) {
        list0[q] = q;
        if (q == 0) count[0] = Q;
    } else {
        // warp-aggregated append (order inside the list does not affect any result: rows are independent)
        bool a = active[q] != 0;
        unsigned m = __ballot_sync(__activemask(), a);
        if (a) {
            int lane = threadIdx.x & 31, leader = __ffs(m) - 1, base = 0;
            if (lane == leader) base = atomicAdd(&count[0], __popc(m));
            base = __shfl_sync(m, base, leader);
            list0[base + __popc(m & ((1u << lane) - 1u))] = q;
        }
    }
}

// every executed step consumed one noise draw (net.py:147-150)
__global__ void k_advance_cursor(int *cursor, const int *__restrict__ n_steps, const uint8_t *__restrict__ active,
                                 int Q)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < Q && (!active || active[q])) cursor[q] += n_steps[q];
}

#include <mutex>
struct RlSide { cudaStream_t stream[RL_MAX_SPLIT - 1]; cudaEvent_t fork, join[RL_MAX_SPLIT - 1]; };
static RlSide g_rl_side[64];
static std::mutex g_rl_mutex[64];      // the fork / launch / join sequence of one device is not re-entrant

// The side streams of the split RL extend and their fork / join events, per device.
static int rl_side(int device, RlSide **out)
{
    CRLK_REQUIRE(device >= 0 && device < 64, "device index");
    RlSide *s = &g_rl_side[device];
    if (!s->fork) {
        for (int i = 0; i < RL_MAX_SPLIT - 1; i++) {
            CRLK_CUDA(cudaStreamCreateWithFlags(&s->stream[i], cudaStreamNonBlocking));
            CRLK_CUDA(cudaEventCreateWithFlags(&s->join[i], cudaEventDisableTiming));
        }
        CRLK_CUDA(cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming));
    }
    *out = s;
    return CRLK_OK;
}

extern "C" int crlk_extend_rl(const CrlkEnv *env, const CrlkPolicy *p, const CrlkExtendParams *ext,
                              const double *from_states, const double *targets, int32_t Q,
                              const float *noise, int32_t noise_rows, int32_t *noise_cursor, float eps,
                              double *path, int32_t *n_steps, int32_t *status, int32_t *node_step,
                              int32_t *n_nodes, float *actions, uint8_t *flags, const uint8_t *active,
                              void *workspace, void *stream)
{
    CRLK_ENV_OK(env);
    CRLK_REQUIRE(p != nullptr, "policy is NULL");
    CRLK_REQUIRE(p->dims[0] == 733, "the actor must take the 733-vector observation");
    ExtP x;
    int rc = crlk_make_ext(env, ext, &x);
    if (rc) return rc;
    CRLK_REQUIRE(Q >= 0, "Q < 0");
    if (Q == 0) return CRLK_OK;
    CRLK_REQUIRE(from_states && targets && path && n_steps && status && node_step && n_nodes && workspace,
                 "NULL array");
    CRLK_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    CRLK_REQUIRE(!noise || noise_rows >= (noise_cursor ? 1 : x.n_max_extend), "noise_rows too small");
    CRLK_REQUIRE(x.n_max_extend + 2 <= 4096, "n_max_extend too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int per = x.n_max_extend / x.n_tree + 1;
    // Every step runs on the COMPACTED batch of queries that are still extending:
    // k_rl_step appends the survivors to the next step's row list and counts them on the
    // device; the observation, GEMM and head kernels read that count (no host sync) and
    // skip the rows / row tiles past it.
    // Large batches are cut into two halves that advance on two streams, launches interleaved
    // step by step: while one half's GEMMs hold the tensor pipe, the other half's observation /
    // motion / collision kernels use the rest of the SM.  Queries are independent, so the split
    // changes no result.
    const int ns = rl_splits(Q, policy_uses_tc(p));
    RlSide *side = nullptr;
    std::unique_lock<std::mutex> guard;
    if (ns > 1) {
        CRLK_REQUIRE(env->device >= 0 && env->device < 64, "device index");
        guard = std::unique_lock<std::mutex>(g_rl_mutex[env->device]);
        rc = rl_side(env->device, &side);
        if (rc) return rc;
        CRLK_CUDA(cudaEventRecord(side->fork, st));
        for (int i = 1; i < ns; i++) CRLK_CUDA(cudaStreamWaitEvent(side->stream[i - 1], side->fork, 0));
    }
    // Everything between the fork above and the join below runs inside `body`, so that an error on
    // any launch still joins the side streams back into the caller's stream before returning.
    auto body = [&]() -> int {
    struct Half { int q0, n; cudaStream_t s; RlWs w; void *planes; int Mpad, Kp; } h[RL_MAX_SPLIT];
    const bool tc = policy_uses_tc(p);
    unsigned char *wsb = (unsigned char *)workspace;
    for (int i = 0; i < ns; i++) {
        rl_slice(Q, ns, i, &h[i].q0, &h[i].n);
        h[i].s = (i == 0) ? st : side->stream[i - 1];
        size_t used = rl_carve(p, h[i].n > 0 ? h[i].n : 1, wsb, &h[i].w);
        wsb += (used + 255) & ~(size_t)255;
        h[i].planes = nullptr; h[i].Mpad = 0; h[i].Kp = 0;
        if (h[i].n <= 0) continue;
        if (tc) crlk_tc_input_planes(p->tc, h[i].w.mlp, h[i].n, &h[i].planes, &h[i].Mpad, &h[i].Kp);
        const int q0 = h[i].q0, n = h[i].n;
        CRLK_CUDA(cudaMemsetAsync(h[i].w.count, 0, sizeof(int) * (x.n_max_extend + 2), h[i].s));
        k_rl_init<<<(n + 127) / 128, 128, 0, h[i].s>>>(from_states + 5 * (size_t)q0, active ? active + q0 : nullptr, n,
                                                      h[i].w.state, h[i].w.nn, h[i].w.list[0], h[i].w.count,
                                                      n_steps + q0, status + q0, n_nodes + q0,
                                                      flags ? flags + q0 : nullptr);
        CRLK_LAUNCH_CHECK();
    }
    // Per step and slice: the policy's hidden layers on the compacted batch (tensor cores), then ONE
    // kernel for head + motion + validity + bookkeeping + the next observation (k_rl_advance).  The
    // survivors' row list and count stay on the device; kernels read the count and skip rows past it.
    RlAdvP a[RL_MAX_SPLIT];
    const int L = p->n_layers;
    for (int i = 0; i < ns; i++) {
        if (h[i].n <= 0) continue;
        const int q0 = h[i].q0;
        RlWs &w = h[i].w;
        RlAdvP &r = a[i];
        r.h = nullptr; r.ld_h = 0;
        r.w = p->d_w[L - 1]; r.bias = p->d_b[L - 1]; r.K = p->kpad[L - 1];
        r.noise = noise ? noise + (size_t)q0 * noise_rows * 2 : nullptr;
        r.noise_stride = (long long)noise_rows * 2;
        r.noise_cursor = noise_cursor ? noise_cursor + q0 : nullptr;
        r.noise_off = 0; r.noise_rows = noise_rows; r.eps = eps;
        for (int c = 0; c < 2; c++) { r.low[c] = p->low[c]; r.high[c] = p->high[c]; r.scale[c] = p->scale[c]; r.hbias[c] = p->bias[c]; }
        r.targets = targets + 5 * (size_t)q0;
        r.list_in = w.list[0]; r.count_in = w.count; r.list_out = nullptr; r.count_out = nullptr;
        r.state = w.state; r.nn = w.nn;
        r.path = path + (size_t)q0 * x.n_max_extend * 5;
        r.n_steps = n_steps + q0; r.status = status + q0; r.node_step = node_step + (size_t)q0 * per;
        r.n_nodes = n_nodes + q0;
        r.actions = actions ? actions + (size_t)q0 * x.n_max_extend * 2 : nullptr;
        r.flags = flags ? flags + q0 : nullptr;
        r.obs32 = tc ? nullptr : w.obs; r.ld32 = 736;
        r.planes_v = h[i].planes; r.Mpad = h[i].Mpad; r.Kp = h[i].Kp;
        r.k = 0; r.n = h[i].n;
        rc = crlk_rl_advance(env, &x, &r, 1, (void *)h[i].s);            // observation of the start states (rrt_rl.py:34)
        if (rc) return rc;
    }
    for (int k = 1; k <= x.n_max_extend; k++) {
        for (int i = 0; i < ns; i++) {
            if (h[i].n <= 0) continue;
            RlWs &w = h[i].w;
            RlAdvP &r = a[i];
            const int *cnt = w.count + (k - 1);
            // action = policy(obs) (rrt_rl.py:41); only the four goal / velocity inputs have low planes
            rc = policy_hidden_impl(p, tc ? nullptr : w.obs, 736, h[i].n, w.mlp, h[i].s, cnt, 1, &r.h, &r.ld_h);
            if (rc) return rc;
            r.k = k; r.noise_off = k - 1;
            r.list_in = w.list[(k - 1) & 1]; r.count_in = cnt;
            r.list_out = w.list[k & 1]; r.count_out = w.count + k;
            rc = crlk_rl_advance(env, &x, &r, 0, (void *)h[i].s);
            if (rc) return rc;
        }
    }
    return CRLK_OK;
    };
    rc = body();
    for (int i = 1; i < ns; i++) {
        if (cudaEventRecord(side->join[i - 1], side->stream[i - 1]) != cudaSuccess ||
            cudaStreamWaitEvent(st, side->join[i - 1], 0) != cudaSuccess) {
            if (!rc) { crlk_set_error("joining the side streams of crlk_extend_rl failed"); rc = CRLK_E_CUDA; }
        }
    }
    if (rc) return rc;
    if (noise && noise_cursor) {
        k_advance_cursor<<<(Q + 127) / 128, 128, 0, st>>>(noise_cursor, n_steps, active, Q);
        CRLK_LAUNCH_CHECK();
    }
    return CRLK_OK;
}
