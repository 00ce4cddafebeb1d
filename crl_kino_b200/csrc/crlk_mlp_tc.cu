// libcrlk.so -- actor MLP hidden layers on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM, operands staged by TMA; sm_100a only).
//
// The reference MLP is float32 (torch sgemm).  A single bf16 or tf32 pass is
// ~1e-3 off, too coarse for the 1e-4 trajectory contract after 50 closed-loop
// steps, so every float32 operand is split into three bf16 planes
//     x = x1 + x2 + x3,  x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2)
// and  A.W^T  is accumulated in float32 (TMEM) over the six plane products
//     A1W1 + A1W2 + A2W1 + A2W2 + A1W3 + A3W1      (dropped terms <= 2^-24 |a||w|),
// i.e. one bf16 GEMM with a 6x longer K loop whose TMA loads walk the planes.
// The epilogue adds the bias, applies ReLU and writes the next layer's three
// bf16 planes directly (and/or float32 for the head kernel).
//
// Tile: 128 (batch rows) x 128 (features) x 64 (k), UMMA 128x128x16, 4 smem
// stages of 32 KB, 128 TMEM columns.  Warp roles: 0 = TMA producer, 1 = MMA
// issuer (one thread), 2 = TMEM allocator, 4-7 = epilogue (one TMEM lane
// quadrant each).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "crlk_internal.h"

#define TC_BM 128
#define TC_BN 128
#define TC_BK 64
#ifndef TC_STAGES
#define TC_STAGES 3
#endif
#ifndef TC_MIN_BLOCKS
#define TC_MIN_BLOCKS 2
#endif
#define TC_STAGE_BYTES (TC_BM * TC_BK * 2)     // 16 KB per operand tile
#define TC_THREADS 256
#define TC_SMEM_BYTES (2 * TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/)
#define TC_NPAIR 6

__device__ __constant__ int c_pair_a[TC_NPAIR] = {0, 0, 1, 1, 0, 2};
__device__ __constant__ int c_pair_w[TC_NPAIR] = {0, 1, 0, 1, 2, 0};

// ---------------------------------------------------------------- PTX helpers

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// Bounded wait: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// K-major, 128-byte swizzled operand tile (rows of 64 bf16 = 128 B; 8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused with swizzle) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = 128.
__device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n)
{
    uint32_t d = 0;
    d |= 1u << 4;                    // c_format = F32
    d |= 1u << 7;                    // a_format = BF16
    d |= 1u << 10;                   // b_format = BF16
    d |= (uint32_t)(n >> 3) << 17;   // n_dim
    d |= (uint32_t)(m >> 4) << 24;   // m_dim
    return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void split3(float x, __nv_bfloat16 &a, __nv_bfloat16 &b, __nv_bfloat16 &c)
{
    a = __float2bfloat16_rn(x);
    float r1 = x - __bfloat162float(a);
    b = __float2bfloat16_rn(r1);
    float r2 = r1 - __bfloat162float(b);
    c = __float2bfloat16_rn(r2);
}

// bias + ReLU on 32 accumulator columns of one row; writes float32 and/or the
// three bf16 planes of the next layer's operand.
__device__ __forceinline__ void epilogue_store(const uint32_t (&r)[32], int m, int n, const float *__restrict__ bias,
                                               int Mpad, __nv_bfloat16 *__restrict__ planes_out, int ld_out,
                                               float *__restrict__ c32, int ldc)
{
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = fmaxf(__uint_as_float(r[j]) + __ldg(&bias[n + j]), 0.f);
    if (c32) {
        float4 *dst = reinterpret_cast<float4 *>(c32 + (size_t)m * ldc + n);
#pragma unroll
        for (int j = 0; j < 8; j++) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    if (planes_out) {
        uint32_t p0[16], p1[16], p2[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            __nv_bfloat16 a0, b0, c0, a1, b1, c1;
            split3(v[2 * j], a0, b0, c0);
            split3(v[2 * j + 1], a1, b1, c1);
            p0[j] = (uint32_t)__bfloat16_as_ushort(a0) | ((uint32_t)__bfloat16_as_ushort(a1) << 16);
            p1[j] = (uint32_t)__bfloat16_as_ushort(b0) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
            p2[j] = (uint32_t)__bfloat16_as_ushort(c0) | ((uint32_t)__bfloat16_as_ushort(c1) << 16);
        }
        const size_t plane = (size_t)Mpad * ld_out;
        uint4 *d0 = reinterpret_cast<uint4 *>(planes_out + (size_t)m * ld_out + n);
        uint4 *d1 = reinterpret_cast<uint4 *>(planes_out + plane + (size_t)m * ld_out + n);
        uint4 *d2 = reinterpret_cast<uint4 *>(planes_out + 2 * plane + (size_t)m * ld_out + n);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            d0[j] = make_uint4(p0[4 * j], p0[4 * j + 1], p0[4 * j + 2], p0[4 * j + 3]);
            d1[j] = make_uint4(p1[4 * j], p1[4 * j + 1], p1[4 * j + 2], p1[4 * j + 3]);
            d2[j] = make_uint4(p2[4 * j], p2[4 * j + 1], p2[4 * j + 2], p2[4 * j + 3]);
        }
    }
}

// The same epilogue with coalesced global stores (persistent and CTA-pair kernels).  A thread owns one
// accumulator ROW (TMEM lane), so a direct store instruction touches 32 different cache lines with 16 B
// each -- 192 such instructions per 128 x 128 tile made the epilogue, not the MMAs or the operand
// traffic, the longest phase of a tile.  Here the warp's 32 rows x 32 columns pass through a per-warp
// staging buffer in shared memory (XOR-swizzled 16 B pieces: conflict free both ways) and every
// store instruction then writes whole row segments: 8 rows x 64 B of a bf16 plane or 4 rows x 128 B
// of the float32 output.  nplanes = operand planes the next layer reads (2 in the three-product mode).
#define TC_STG_WARP_BYTES 6144                      // 3 planes x 32 rows x 64 B (float32: 32 rows x 128 B)
#define TC_STG_BYTES (4 * TC_STG_WARP_BYTES)
__device__ __forceinline__ void epilogue_store_staged(const uint32_t (&r)[32], int m_base, int lane, int n,
                                                      const float *__restrict__ bias, int B, int Mpad,
                                                      __nv_bfloat16 *__restrict__ planes_out, int ld_out, int nplanes,
                                                      float *__restrict__ c32, int ldc, uint4 *stg)
{
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = fmaxf(__uint_as_float(r[j]) + __ldg(&bias[n + j]), 0.f);
    if (c32) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            stg[lane * 8 + (j ^ (lane & 7))] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                                          __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int row = 4 * i + (lane >> 3), j = lane & 7, m = m_base + row;
            const uint4 x = stg[row * 8 + (j ^ (row & 7))];
            if (m < B) *reinterpret_cast<uint4 *>(c32 + (size_t)m * ldc + n + 4 * j) = x;
        }
        __syncwarp();
    }
    if (planes_out) {
        const int sw = (lane >> 1) & 3;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t a[4], b[4], c[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                __nv_bfloat16 a0, b0, c0, a1, b1, c1;
                split3(v[8 * j + 2 * u], a0, b0, c0);
                split3(v[8 * j + 2 * u + 1], a1, b1, c1);
                a[u] = (uint32_t)__bfloat16_as_ushort(a0) | ((uint32_t)__bfloat16_as_ushort(a1) << 16);
                b[u] = (uint32_t)__bfloat16_as_ushort(b0) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
                c[u] = (uint32_t)__bfloat16_as_ushort(c0) | ((uint32_t)__bfloat16_as_ushort(c1) << 16);
            }
            stg[lane * 4 + (j ^ sw)] = make_uint4(a[0], a[1], a[2], a[3]);
            stg[128 + lane * 4 + (j ^ sw)] = make_uint4(b[0], b[1], b[2], b[3]);
            if (nplanes > 2) stg[256 + lane * 4 + (j ^ sw)] = make_uint4(c[0], c[1], c[2], c[3]);
        }
        __syncwarp();
        const size_t plane = (size_t)Mpad * ld_out;
#pragma unroll
        for (int pl = 0; pl < 3; pl++) {
            if (pl >= nplanes) continue;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int row = 8 * i + (lane >> 2), j = lane & 3, m = m_base + row;
                const uint4 x = stg[pl * 128 + row * 4 + (j ^ ((row >> 1) & 3))];
                if (m < B) *reinterpret_cast<uint4 *>(planes_out + pl * plane + (size_t)m * ld_out + n + 8 * j) = x;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ the GEMM

// planes_out: [3][Mpad][ld_out] bf16 (nullable), c32: [B][ldc] float (nullable).
__global__ void __launch_bounds__(TC_THREADS, TC_MIN_BLOCKS)
k_gemm_bf16x3(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
              int Mpad, int N, int kblocks, const float *__restrict__ bias, int B,
              __nv_bfloat16 *__restrict__ planes_out, int ld_out, float *__restrict__ c32, int ldc)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;    // SWIZZLE_128B needs 1024 B alignment
    const uint32_t sA = base, sB = base + TC_STAGES * TC_STAGE_BYTES;
    const uint32_t bars = sB + TC_STAGES * TC_STAGE_BYTES;           // full[S], empty[S], tmem_full, tmem slot
    const uint32_t full0 = bars, empty0 = bars + 8 * TC_STAGES, tfull = bars + 16 * TC_STAGES;
    const uint32_t tslot = tfull + 8;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tslot_ptr = (volatile uint32_t *)(gen + 2 * TC_STAGES * TC_STAGE_BYTES + 16 * TC_STAGES + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;
    const int iters = TC_NPAIR * kblocks;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "n"(TC_BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot_ptr;

    if (warp == 0 && lane == 0) {
        // ---- TMA producer
        for (int it = 0; it < iters; it++) {
            int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
            int pair = it / kblocks, kb = it - pair * kblocks;
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            mbar_expect_tx(full0 + 8 * s, 2 * TC_STAGE_BYTES);
            tma_load_2d(sA + s * TC_STAGE_BYTES, &tmA, kb * TC_BK, c_pair_a[pair] * Mpad + m0, full0 + 8 * s);
            tma_load_2d(sB + s * TC_STAGE_BYTES, &tmW, kb * TC_BK, c_pair_w[pair] * N + n0, full0 + 8 * s);
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer (single thread)
        const uint32_t idesc = umma_idesc_bf16(TC_BM, TC_BN);
        for (int it = 0; it < iters; it++) {
            int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
            mbar_wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint64_t da = umma_desc_sw128(sA + s * TC_STAGE_BYTES);
            uint64_t db = umma_desc_sw128(sB + s * TC_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; k++)      // UMMA_K = 16 bf16 = 32 bytes along the swizzled row
                umma_bf16(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it | k) != 0);
            umma_commit(empty0 + 8 * s);              // frees the smem stage when these MMAs retire
        }
        umma_commit(tfull);                           // accumulator complete
    } else if (warp >= 4) {
        // ---- epilogue: warp q owns TMEM lanes [32q, 32q + 32) = rows m0 + 32q + lane
        const int q = warp & 3;
        const int m = m0 + 32 * q + lane;
        mbar_wait(tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c = 0; c < TC_BN / 32; c++) {
            uint32_t r[32];
            tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * c), r);
            if (m < B) epilogue_store(r, m, n0 + 32 * c, bias, Mpad, planes_out, ld_out, c32, ldc);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_BN));
    }
}

// Variant with plane-complete stages: one stage holds the three A planes and the
// three W planes of a k block (6 tiles, 96 KB), and all six plane products are
// issued from it.  Each operand tile is then fetched from L2 once per k block
// instead of up to three times (the 128x128x64 tile at 64 FLOP per staged byte
// is L2-bandwidth bound otherwise).  2 stages, 1 CTA per SM.
#define TCF_STAGES 2
#define TCF_STAGE_BYTES (6 * TC_STAGE_BYTES)
#define TCF_SMEM_BYTES (TCF_STAGES * TCF_STAGE_BYTES + 1024 + 256)

__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_bf16x3_fused(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    int Mpad, int N, int kblocks, const float *__restrict__ bias, int B,
                    __nv_bfloat16 *__restrict__ planes_out, int ld_out, float *__restrict__ c32, int ldc)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + TCF_STAGES * TCF_STAGE_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * TCF_STAGES, tfull = bars + 16 * TCF_STAGES;
    const uint32_t tslot = tfull + 8;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tslot_ptr = (volatile uint32_t *)(gen + TCF_STAGES * TCF_STAGE_BYTES + 16 * TCF_STAGES + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TCF_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "n"(TC_BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot_ptr;

    if (warp == 0 && lane == 0) {
        for (int kb = 0; kb < kblocks; kb++) {
            int s = kb % TCF_STAGES, ph = (kb / TCF_STAGES) & 1;
            uint32_t st = base + s * TCF_STAGE_BYTES;
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            mbar_expect_tx(full0 + 8 * s, TCF_STAGE_BYTES);
#pragma unroll
            for (int pl = 0; pl < 3; pl++) {
                tma_load_2d(st + pl * TC_STAGE_BYTES, &tmA, kb * TC_BK, pl * Mpad + m0, full0 + 8 * s);
                tma_load_2d(st + (3 + pl) * TC_STAGE_BYTES, &tmW, kb * TC_BK, pl * N + n0, full0 + 8 * s);
            }
        }
    } else if (warp == 1 && lane == 0) {
        const uint32_t idesc = umma_idesc_bf16(TC_BM, TC_BN);
        for (int kb = 0; kb < kblocks; kb++) {
            int s = kb % TCF_STAGES, ph = (kb / TCF_STAGES) & 1;
            uint32_t st = base + s * TCF_STAGE_BYTES;
            mbar_wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int pair = 0; pair < TC_NPAIR; pair++) {
                uint64_t da = umma_desc_sw128(st + c_pair_a[pair] * TC_STAGE_BYTES);
                uint64_t db = umma_desc_sw128(st + (3 + c_pair_w[pair]) * TC_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; k++)
                    umma_bf16(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | pair | k) != 0);
            }
            umma_commit(empty0 + 8 * s);
        }
        umma_commit(tfull);
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int m = m0 + 32 * q + lane;
        mbar_wait(tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c = 0; c < TC_BN / 32; c++) {
            uint32_t r[32];
            tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * c), r);
            if (m < B) epilogue_store(r, m, n0 + 32 * c, bias, Mpad, planes_out, ld_out, c32, ldc);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_BN));
    }
}

// Persistent variant of the plane-complete kernel: one CTA per SM loops over
// output tiles; the accumulator is double buffered in TMEM (2 x 128 columns) so
// the epilogue of tile i (TMEM -> registers -> bias/ReLU/split -> global) runs
// under the MMAs of tile i + 1.
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

#define TCP_SMEM_BYTES (TCF_SMEM_BYTES + TC_STG_BYTES)
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_bf16x3_persist(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                      int Mpad, int N, int kblocks, const float *__restrict__ bias, int B,
                      __nv_bfloat16 *__restrict__ planes_out, int ld_out, float *__restrict__ c32, int ldc,
                      const int *__restrict__ n_rows_dev, int a_lo_kb, int nprod)
{
    // nprod = 6: all plane products down to 2^-24 of the float32 product (a1w1 a1w2 a2w1 a2w2 a1w3 a3w1);
    // nprod = 3: the products down to 2^-16 only (a1w1 a1w2 a2w1): planes 3 are neither loaded nor used.
    // n_rows_dev (nullable): the live row count is read on the device (compacted batches of the
    // RL steer loop), row tiles past it are skipped.  a_lo_kb: A planes 2 and 3 are zero beyond
    // the first a_lo_kb k blocks (the local map of the observation is exactly 0/1 in bf16), so
    // those blocks load 4 tiles and issue the 3 products of A plane 1 only.
    if (n_rows_dev) B = min(B, *n_rows_dev);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // stage = the A planes then the W planes of one k block: 6 tiles x 2 stages, or 4 tiles x 3 stages
    // in the three-product mode (same 192 KB; the third stage hides the TMA latency that the shorter
    // MMA burst of a stage no longer covers)
    const int npl = nprod == 6 ? 3 : 2, nst = nprod == 6 ? 2 : 3;
    const uint32_t stage_bytes = (uint32_t)(2 * npl) * TC_STAGE_BYTES;
    const uint32_t bars = base + TCF_STAGES * TCF_STAGE_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 8 * 3;
    const uint32_t tfull0 = bars + 16 * 3, tempty0 = tfull0 + 16;
    const uint32_t tslot = tempty0 + 16;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tslot_ptr = (volatile uint32_t *)(gen + TCF_STAGES * TCF_STAGE_BYTES + 16 * 3 + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = N / TC_BN, tiles = tiles_n * ((B + TC_BM - 1) / TC_BM);     // Mpad is the plane pitch
    // per-warp staging buffers of the epilogue, behind the 256 bytes of barriers
    uint4 *stg = reinterpret_cast<uint4 *>(gen + TCF_STAGES * TCF_STAGE_BYTES + 256 + (warp & 3) * TC_STG_WARP_BYTES);

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < nst; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "n"(2 * TC_BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot_ptr;

    if (warp == 0 && lane == 0) {
        int git = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int m0 = (t / tiles_n) * TC_BM, n0 = (t % tiles_n) * TC_BN;
            for (int kb = 0; kb < kblocks; kb++, git++) {
                int s = git % nst, ph = (git / nst) & 1;
                uint32_t st = base + s * stage_bytes;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                const bool lo = kb < a_lo_kb;
                mbar_expect_tx(full0 + 8 * s, ((lo ? npl : 1) + npl) * TC_STAGE_BYTES);
#pragma unroll
                for (int pl = 0; pl < 3; pl++) {
                    if (pl >= npl) continue;
                    if (pl == 0 || lo)
                        tma_load_2d(st + pl * TC_STAGE_BYTES, &tmA, kb * TC_BK, pl * Mpad + m0, full0 + 8 * s);
                    tma_load_2d(st + (npl + pl) * TC_STAGE_BYTES, &tmW, kb * TC_BK, pl * N + n0, full0 + 8 * s);
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        const uint32_t idesc = umma_idesc_bf16(TC_BM, TC_BN);
        int git = 0, ti = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ti++) {
            const int as = ti & 1, aph = (ti >> 1) & 1;
            mbar_wait(tempty0 + 8 * as, aph ^ 1);            // epilogue drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem + (uint32_t)(as * TC_BN);
            for (int kb = 0; kb < kblocks; kb++, git++) {
                int s = git % nst, ph = (git / nst) & 1;
                uint32_t st = base + s * stage_bytes;
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool lo = kb < a_lo_kb;
#pragma unroll
                for (int pair = 0; pair < TC_NPAIR; pair++) {
                    if (!lo && c_pair_a[pair] != 0) continue;
                    if (pair >= nprod) continue;             // the pairs are ordered by magnitude
                    uint64_t da = umma_desc_sw128(st + c_pair_a[pair] * TC_STAGE_BYTES);
                    uint64_t db = umma_desc_sw128(st + (npl + c_pair_w[pair]) * TC_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; k++)
                        umma_bf16(acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | pair | k) != 0);
                }
                umma_commit(empty0 + 8 * s);
            }
            umma_commit(tfull0 + 8 * as);
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        int ti = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ti++) {
            const int as = ti & 1, aph = (ti >> 1) & 1;
            const int m0 = (t / tiles_n) * TC_BM, n0 = (t % tiles_n) * TC_BN;
            mbar_wait(tfull0 + 8 * as, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < TC_BN / 32; c++) {
                uint32_t r[32];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(as * TC_BN + 32 * c), r);
                if (c == TC_BN / 32 - 1) {
                    // all of this warp's TMEM reads are done: hand the accumulator back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty0 + 8 * as);
                }
                if (m0 + 32 * q < B)       // warp-uniform: some row of this warp is live
                    epilogue_store_staged(r, m0 + 32 * q, lane, n0 + 32 * c, bias, B, Mpad, planes_out, ld_out, npl,
                                          c32, ldc, stg);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * TC_BN));
    }
}


// ---------------------------------------------------------------- all hidden layers in ONE launch
// Between the steps of the RL extension the batch is a few thousand rows: a layer is 90-360 tiles,
// i.e. 0.6-2.4 waves of the 148 SMs, and each of the four per-layer launches pays its own prologue,
// pipeline fill, partial last wave and drain (measured: 21-42 us per layer at 8192 rows against
// 5-11 us of MMA time).  This kernel runs the whole chain of layers as one persistent grid with a
// DATAFLOW schedule: the tiles of all layers form one global list (layer-major, row-tile-major inside a
// layer) that the CTAs pull from with an atomic counter; a tile of layer l needs row tile m of layer
// l - 1 complete (all its column tiles), which the epilogue warps publish with a release-add on
// done[l - 1][m] and the TMA producer acquires before it loads the A operand (the W tiles of the first k
// block are requested before the wait).  So the SMs that a layer's last wave leaves idle already run the
// next layer's first row tiles, and nothing is launched or drained in between.
// Deadlock freedom does not depend on all CTAs being resident: a tile only waits for tiles with smaller
// indices, and those were pulled by CTAs that are running.
// Activations live in two ping-pong plane buffers with ONE row pitch (ld_planes), so row tile m of layer
// l + 1's output may overwrite row tile m of layer l's input: every reader of those rows (the tiles
// (l, m, *)) has completed by then.  The first layer's input planes are a third buffer.
#define TCC_MAX_LAYERS 4
// From this many rows on, one launch per layer of the CTA-pair kernel beats the chain (measured: 0.799 vs 0.815 ms
// at 65,536 rows; at 24,576 rows the chain wins, 0.312 vs 0.359 ms)
#define TC_PAIR_MIN_ROWS 49152
#define TCC_RING 8
struct ChainP {
    CUtensorMap tmA[TCC_MAX_LAYERS];
    CUtensorMap tmW[TCC_MAX_LAYERS];
    const float *bias[TCC_MAX_LAYERS];
    __nv_bfloat16 *out_planes[TCC_MAX_LAYERS];      // null for the last layer (float32 output `last`)
    float *last;
    const int *n_rows_dev;
    int *ctrl;                                       // [0] tile counter; [64 + l * row_tiles_max + m] epilogue warps done
    int N[TCC_MAX_LAYERS], kblocks[TCC_MAX_LAYERS];
    int n_layers, Mpad, B, ld_planes, ld_last, a_lo_kb, nprod, row_tiles_max;
};

__device__ __forceinline__ int ld_acquire_gpu(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void red_release_gpu_add(int *p, int v)
{
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// BN = feature columns of a tile: 128, or 256 (three-product mode only: the A planes of a k block are then
// fetched once per 256 output columns -- 48 KB instead of 64 KB of operand traffic per 128 x 128 x 64 block of
// products, which is what bounds these GEMMs; two 96 KB stages, the whole TMEM as a double-buffered accumulator).
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_mlp_chain(const __grid_constant__ ChainP p)
{
    int B = p.B;
    if (p.n_rows_dev) B = min(B, *p.n_rows_dev);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t WPL = BN * TC_BK * 2;                           // bytes of one W plane tile
    const int npl = p.nprod == 6 ? 3 : 2, nst = (BN == 256 || p.nprod == 6) ? 2 : 3;
    const uint32_t stage_bytes = (uint32_t)npl * (TC_STAGE_BYTES + WPL);
    const uint32_t bars = base + TCF_STAGES * TCF_STAGE_BYTES;
    const uint32_t full0 = bars, empty0 = bars + 24, tfull0 = bars + 48, tempty0 = bars + 64, tslot = bars + 80;
    const uint32_t tq0 = bars + 96;                                    // TCC_RING tile-queue barriers, then the tile ids
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tslot_ptr = (volatile uint32_t *)(gen + TCF_STAGES * TCF_STAGE_BYTES + 80);
    volatile int *tile_ids = (volatile int *)(gen + TCF_STAGES * TCF_STAGE_BYTES + 96 + 8 * TCC_RING);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4 *stg = reinterpret_cast<uint4 *>(gen + TCF_STAGES * TCF_STAGE_BYTES + 256 + (warp & 3) * TC_STG_WARP_BYTES);

    const int row_tiles = (B + TC_BM - 1) / TC_BM;
    int total = 0;
    for (int l = 0; l < p.n_layers; l++) total += row_tiles * (p.N[l] / BN);
    auto decode = [&](int t, int &l, int &m0, int &n0) {
        l = 0;
        while (t >= row_tiles * (p.N[l] / BN)) { t -= row_tiles * (p.N[l] / BN); l++; }
        const int tn = p.N[l] / BN;
        m0 = (t / tn) * TC_BM; n0 = (t % tn) * BN;
    };

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 3; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
        for (int i = 0; i < TCC_RING; i++) mbar_init(tq0 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "n"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot_ptr;

    if (warp == 0 && lane == 0) {
        // ---- tile scheduler + TMA producer
        int git = 0;
        for (int ti = 0;; ti++) {
            int t = atomicAdd(&p.ctrl[0], 1);
            if (t >= total) t = -1;
            // the ring slot's previous tile (ti - TCC_RING) was read by every consumer long ago: the MMA thread
            // is at most one tile behind (3-stage ring, >= 8 k blocks per tile) and it cannot start tile j
            // before all epilogue warps have taken up tile j - 2
            tile_ids[ti & (TCC_RING - 1)] = t;
            mbar_arrive(tq0 + 8 * (ti & (TCC_RING - 1)));
            if (t < 0) break;
            int l, m0, n0;
            decode(t, l, m0, n0);
            const int kblocks = p.kblocks[l], a_lo = (l == 0 && p.a_lo_kb > 0) ? p.a_lo_kb : kblocks;
            const CUtensorMap *mA = &p.tmA[l], *mW = &p.tmW[l];
            const int N = p.N[l];
            for (int kb = 0; kb < kblocks; kb++, git++) {
                int s = git % nst, ph = (git / nst) & 1;
                uint32_t st = base + s * stage_bytes;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                const bool lo = kb < a_lo;
                mbar_expect_tx(full0 + 8 * s, (lo ? npl : 1) * TC_STAGE_BYTES + npl * WPL);
#pragma unroll
                for (int pl = 0; pl < 3; pl++)
                    if (pl < npl)
                        for (int hf = 0; hf < BN / 128; hf++)      // 128-row boxes of the W map
                            tma_load_2d(st + npl * TC_STAGE_BYTES + pl * WPL + hf * TC_STAGE_BYTES, mW, kb * TC_BK,
                                        pl * N + n0 + hf * 128, full0 + 8 * s);
                if (kb == 0 && l > 0) {
                    // row tile m of the previous layer: all column tiles stored by all four epilogue warps
                    const int *flag = p.ctrl + 64 + (l - 1) * p.row_tiles_max + m0 / TC_BM;
                    const int need = 4 * (p.N[l - 1] / BN);
                    for (uint32_t spin = 0; ld_acquire_gpu(flag) < need; spin++)
                        if (spin > (1u << 24)) __trap();
                    asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy stores -> this thread's TMA reads
                }
#pragma unroll
                for (int pl = 0; pl < 3; pl++)
                    if (pl < npl && (pl == 0 || lo))
                        tma_load_2d(st + pl * TC_STAGE_BYTES, mA, kb * TC_BK, pl * p.Mpad + m0, full0 + 8 * s);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer
        const uint32_t idesc = umma_idesc_bf16(TC_BM, BN);
        int git = 0;
        for (int ti = 0;; ti++) {
            mbar_wait(tq0 + 8 * (ti & (TCC_RING - 1)), (ti / TCC_RING) & 1);
            const int t = tile_ids[ti & (TCC_RING - 1)];
            if (t < 0) break;
            int l, m0, n0;
            decode(t, l, m0, n0);
            const int kblocks = p.kblocks[l], a_lo = (l == 0 && p.a_lo_kb > 0) ? p.a_lo_kb : kblocks;
            const int as = ti & 1, aph = (ti >> 1) & 1;
            mbar_wait(tempty0 + 8 * as, aph ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem + (uint32_t)(as * BN);
            for (int kb = 0; kb < kblocks; kb++, git++) {
                int s = git % nst, ph = (git / nst) & 1;
                uint32_t st = base + s * stage_bytes;
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool lo = kb < a_lo;
#pragma unroll
                for (int pair = 0; pair < TC_NPAIR; pair++) {
                    if (!lo && c_pair_a[pair] != 0) continue;
                    if (pair >= p.nprod) continue;
                    uint64_t da = umma_desc_sw128(st + c_pair_a[pair] * TC_STAGE_BYTES);
                    uint64_t db = umma_desc_sw128(st + npl * TC_STAGE_BYTES + c_pair_w[pair] * WPL);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; k++)
                        umma_bf16(acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | pair | k) != 0);
                }
                umma_commit(empty0 + 8 * s);
            }
            umma_commit(tfull0 + 8 * as);
        }
    } else if (warp >= 4) {
        // ---- epilogue
        const int q = warp & 3;
        for (int ti = 0;; ti++) {
            mbar_wait(tq0 + 8 * (ti & (TCC_RING - 1)), (ti / TCC_RING) & 1);
            const int t = tile_ids[ti & (TCC_RING - 1)];
            if (t < 0) break;
            int l, m0, n0;
            decode(t, l, m0, n0);
            const bool is_last = l == p.n_layers - 1;
            const int as = ti & 1, aph = (ti >> 1) & 1;
            mbar_wait(tfull0 + 8 * as, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++) {
                uint32_t r[32];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(as * BN + 32 * c), r);
                if (c == BN / 32 - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty0 + 8 * as);
                }
                if (m0 + 32 * q < B)
                    epilogue_store_staged(r, m0 + 32 * q, lane, n0 + 32 * c, p.bias[l], B, p.Mpad,
                                          is_last ? nullptr : p.out_planes[l], p.ld_planes, npl,
                                          is_last ? p.last : nullptr, p.ld_last, stg);
            }
            if (!is_last) {
                // publish: this warp's rows of tile (l, m, n) are in global memory
                asm volatile("fence.proxy.async;" ::: "memory");
                __threadfence();
                __syncwarp();
                if (lane == 0) red_release_gpu_add(p.ctrl + 64 + l * p.row_tiles_max + m0 / TC_BM, 1);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN));
    }
}


// ---------------------------------------------------------------- CTA-pair variant
// Two CTAs of a cluster (two SMs of a TPC) work on one 256 x 128 output tile with
// tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of A and HALF of the W tile
// (64 of the 128 feature rows); the tensor core reads both halves.  Per k block a CTA then
// fetches 48 KB of A + 24 KB of W from L2 instead of 48 + 48 -- the single-CTA kernel is bound by
// exactly that operand traffic -- and the smaller stage allows a 3-deep ring.
// Protocol: both producers load into their own shared memory but signal the LEADER's (rank 0)
// full barrier (cp.async.bulk.tensor ... cta_group::2 with the peer bit of the barrier address
// cleared); the leader's MMA thread issues all MMAs and releases a stage / publishes an
// accumulator with tcgen05.commit ... multicast::cluster to both CTAs; each CTA's epilogue warps
// drain their own 128 TMEM lanes and hand the accumulator back to the leader (remote arrive).
// PBN = feature columns of the pair tile (128 or 256): a CTA stages PBN / 2 rows of W per plane.
#define TCP_A_BYTES (3 * TC_STAGE_BYTES)                 // three A planes, 128 rows each
template <int PBN> struct PairCfg {
    static constexpr int b_plane = PBN / 2 * TC_BK * 2;           // bytes of one half W plane tile
    static constexpr int b_bytes = 3 * b_plane;
    static constexpr int stage = TCP_A_BYTES + b_bytes;           // 72 KB (PBN 128) / 96 KB (PBN 256)
    static constexpr int stages = (PBN == 128) ? 3 : 2;
    static constexpr int stg = (PBN == 256) ? TC_STG_BYTES : 0;    // epilogue staging (the 128-column ring leaves no room)
    static constexpr int smem = stages * stage + 1024 + 256 + stg;
};

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// TMA load of one box into this CTA's shared memory, completion bytes reported to the pair leader's barrier.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}

// arrive on the barrier at the same shared-memory offset in both CTAs of the pair when the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// arrive on a barrier of the pair leader (rank 0) from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar)
{
    asm volatile(
        "{\n\t.reg .b32 r;\n\t"
        "mapa.shared::cluster.u32 r, %0, 0;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [r];\n\t}"
        ::"r"(bar) : "memory");
}

template <int PBN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
k_gemm_bf16x3_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWh,
                   int Mpad, int N, int kblocks, const float *__restrict__ bias, int B,
                   __nv_bfloat16 *__restrict__ planes_out, int ld_out, float *__restrict__ c32, int ldc,
                   const int *__restrict__ n_rows_dev, int a_lo_kb, int nprod)
{
    if (n_rows_dev) B = min(B, *n_rows_dev);
    const int npl = nprod == 6 ? 3 : 2;              // planes in use (k_gemm_bf16x3_persist)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + PairCfg<PBN>::stages * PairCfg<PBN>::stage;
    const uint32_t full0 = bars, empty0 = bars + 8 * PairCfg<PBN>::stages;
    const uint32_t tfull0 = bars + 16 * PairCfg<PBN>::stages, tempty0 = tfull0 + 16;
    const uint32_t tslot = tempty0 + 16;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tslot_ptr = (volatile uint32_t *)(gen + PairCfg<PBN>::stages * PairCfg<PBN>::stage + 16 * PairCfg<PBN>::stages + 32);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4 *stg = reinterpret_cast<uint4 *>(gen + PairCfg<PBN>::stages * PairCfg<PBN>::stage + 256 + (warp & 3) * TC_STG_WARP_BYTES);
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int tiles_n = N / PBN, tiles = tiles_n * ((B + 2 * TC_BM - 1) / (2 * TC_BM));   // 256-row pair tiles

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < PairCfg<PBN>::stages; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; a++) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "n"(2 * PBN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                        // barriers of both CTAs are initialised before anyone signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot_ptr;

    if (warp == 0 && lane == 0) {
        // ---- producer (both CTAs): own A rows, own half of the W rows
        int git = 0;
        for (int t = pair; t < tiles; t += npairs) {
            const int m0 = (t / tiles_n) * 2 * TC_BM + (int)rank * TC_BM;
            const int n0 = (t % tiles_n) * PBN + (int)rank * (PBN / 2);
            for (int kb = 0; kb < kblocks; kb++, git++) {
                const int s = git % PairCfg<PBN>::stages, ph = (git / PairCfg<PBN>::stages) & 1;
                const uint32_t st = base + s * PairCfg<PBN>::stage;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                const bool lo = kb < a_lo_kb;
                if (rank == 0)                  // the leader's barrier counts the bytes of both CTAs
                    mbar_expect_tx(full0 + 8 * s, 2 * ((lo ? npl : 1) * TC_STAGE_BYTES + npl * PairCfg<PBN>::b_plane));
#pragma unroll
                for (int pl = 0; pl < 3; pl++) {
                    if (pl >= npl) continue;
                    if (pl == 0 || lo)
                        tma_load_2d_pair(st + pl * TC_STAGE_BYTES, &tmA, kb * TC_BK, pl * Mpad + m0, full0 + 8 * s);
                    tma_load_2d_pair(st + TCP_A_BYTES + pl * PairCfg<PBN>::b_plane, &tmWh, kb * TC_BK, pl * N + n0,
                                     full0 + 8 * s);
                }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ---- MMA issuer: one thread of the leader CTA, M = 256 across the pair
        const uint32_t idesc = umma_idesc_bf16(2 * TC_BM, PBN);
        int git = 0, ti = 0;
        for (int t = pair; t < tiles; t += npairs, ti++) {
            const int as = ti & 1, aph = (ti >> 1) & 1;
            mbar_wait(tempty0 + 8 * as, aph ^ 1);            // both CTAs' epilogues drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem + (uint32_t)(as * PBN);
            for (int kb = 0; kb < kblocks; kb++, git++) {
                const int s = git % PairCfg<PBN>::stages, ph = (git / PairCfg<PBN>::stages) & 1;
                const uint32_t st = base + s * PairCfg<PBN>::stage;
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool lo = kb < a_lo_kb;
#pragma unroll
                for (int pr = 0; pr < TC_NPAIR; pr++) {
                    if (!lo && c_pair_a[pr] != 0) continue;
                    if (pr >= nprod) continue;
                    uint64_t da = umma_desc_sw128(st + c_pair_a[pr] * TC_STAGE_BYTES);
                    uint64_t db = umma_desc_sw128(st + TCP_A_BYTES + c_pair_w[pr] * PairCfg<PBN>::b_plane);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; k++)
                        umma_bf16_pair(acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | pr | k) != 0);
                }
                umma_commit_pair(empty0 + 8 * s);            // frees the stage in both CTAs
            }
            umma_commit_pair(tfull0 + 8 * as);               // accumulator complete, both CTAs
        }
    } else if (warp >= 4) {
        // ---- epilogue (both CTAs): this CTA's 128 rows of the pair tile
        const int q = warp & 3;
        int ti = 0;
        for (int t = pair; t < tiles; t += npairs, ti++) {
            const int as = ti & 1, aph = (ti >> 1) & 1;
            const int m0 = (t / tiles_n) * 2 * TC_BM + (int)rank * TC_BM, n0 = (t % tiles_n) * PBN;
            const int m = m0 + 32 * q + lane;
            mbar_wait(tfull0 + 8 * as, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < PBN / 32; c++) {
                uint32_t r[32];
                tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(as * PBN + 32 * c), r);
                if (c == PBN / 32 - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(tempty0 + 8 * as);
                }
                if (PairCfg<PBN>::stg) {
                    if (m0 + 32 * q < B)
                        epilogue_store_staged(r, m0 + 32 * q, lane, n0 + 32 * c, bias, B, Mpad, planes_out, ld_out, npl,
                                              c32, ldc, stg);
                } else if (m < B)
                    epilogue_store(r, m, n0 + 32 * c, bias, Mpad, planes_out, ld_out, c32, ldc);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                        // the peer may still be reading this CTA's operands / signalling its barriers
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * PBN));
    }
}

// float32 [B, ld_in] -> three bf16 planes [3][Mpad][ld_out] (columns >= K zero filled)
__global__ void k_split_planes(const float *__restrict__ in, int ld_in, int K, int B, int Mpad,
                               __nv_bfloat16 *__restrict__ out, int ld_out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)Mpad * ld_out;
    if (i >= total) return;
    int m = (int)(i / ld_out), k = (int)(i - (size_t)m * ld_out);
    float x = (m < B && k < K && k < ld_in) ? in[(size_t)m * ld_in + k] : 0.f;
    __nv_bfloat16 a, b, c;
    split3(x, a, b, c);
    out[i] = a;
    out[total + i] = b;
    out[2 * total + i] = c;
}

// ------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D bf16 tensor [rows][cols] (row pitch ld elements), box = 64 columns x 128 rows, 128 B swizzle.
static int make_map(CUtensorMap *map, const void *ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                    uint32_t box_rows = TC_BM)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) { crlk_set_error("cuTensorMapEncodeTiled is not available from the driver"); return CRLK_E_CUDA; }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {ld * 2};
    cuuint32_t box[2] = {TC_BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { crlk_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return CRLK_E_CUDA; }
    return CRLK_OK;
}

static inline int rup(int a, int b) { return (a + b - 1) / b * b; }

struct TcLayer {
    __nv_bfloat16 *d_w;   // [3][N][Kp]
    CUtensorMap map_w;
    CUtensorMap map_wh;   // the same tensor with 64-row boxes (half W tiles of the CTA-pair kernel)
    int N, K, Kp;
};

struct CrlkPolicyTc {
    int nprod;            // 6 or 3 plane products per float32 product (k_gemm_bf16x3_persist)
    int n_hidden;
    TcLayer layer[8];
    int max_width;        // widest activation (elements), multiple of 64
};

static void host_split3(float x, uint16_t &a, uint16_t &b, uint16_t &c)
{
    auto rn = [](float f) -> uint16_t {
        uint32_t u;
        memcpy(&u, &f, 4);
        if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
        u += 0x7fffu + ((u >> 16) & 1u);
        return (uint16_t)(u >> 16);
    };
    auto up = [](uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; };
    a = rn(x);
    float r1 = x - up(a);
    b = rn(r1);
    float r2 = r1 - up(b);
    c = rn(r2);
}

// Can the hidden layers of this actor run on the tensor-core path?
int crlk_tc_supported(const int *dims, int n_layers)
{
    for (int i = 0; i < n_layers - 1; i++)
        if (dims[i + 1] % TC_BN != 0) return 0;
    return 1;
}

int crlk_tc_create(const int *dims, int n_layers, const float *const *weights_host, CrlkPolicyTc **out)
{
    CrlkPolicyTc *t = new CrlkPolicyTc();
    memset(t, 0, sizeof *t);
    t->n_hidden = n_layers - 1;
    {
        // three plane products (a1w1 + a1w2 + a2w1, terms down to 2^-16 of the product) keep the actor
        // within the 1e-4 contract with margin (max |error| 1.9e-6 against float64, the same as with all
        // six products: the float32 accumulation dominates); CRLK_TC_NPROD=6 restores the full split
        const char *e = getenv("CRLK_TC_NPROD");
        t->nprod = (e && atoi(e) == 6) ? 6 : 3;
    }
    t->max_width = 0;
    for (int i = 0; i < t->n_hidden; i++) {
        TcLayer &L = t->layer[i];
        L.N = dims[i + 1]; L.K = dims[i]; L.Kp = rup(dims[i], TC_BK);
        if (L.Kp > t->max_width) t->max_width = L.Kp;
        if (L.N > t->max_width) t->max_width = L.N;
        size_t plane = (size_t)L.N * L.Kp;
        uint16_t *h = (uint16_t *)calloc(3 * plane, sizeof(uint16_t));
        for (int n = 0; n < L.N; n++)
            for (int k = 0; k < L.K; k++) {
                uint16_t a, b, c;
                host_split3(weights_host[i][(size_t)n * L.K + k], a, b, c);
                h[(size_t)n * L.Kp + k] = a;
                h[plane + (size_t)n * L.Kp + k] = b;
                h[2 * plane + (size_t)n * L.Kp + k] = c;
            }
        CRLK_CUDA(cudaMalloc(&L.d_w, 3 * plane * sizeof(uint16_t)));
        CRLK_CUDA(cudaMemcpy(L.d_w, h, 3 * plane * sizeof(uint16_t), cudaMemcpyHostToDevice));
        free(h);
        int rc = make_map(&L.map_w, L.d_w, 3ull * L.N, L.Kp, L.Kp);
        if (rc) return rc;
        rc = make_map(&L.map_wh, L.d_w, 3ull * L.N, L.Kp, L.Kp, TC_BN / 2);
        if (rc) return rc;
    }
    CRLK_CUDA(cudaFuncSetAttribute(k_gemm_bf16x3, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    CRLK_CUDA(cudaFuncSetAttribute(k_gemm_bf16x3_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, TCF_SMEM_BYTES));
    CRLK_CUDA(cudaFuncSetAttribute(k_gemm_bf16x3_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_BYTES));
    CRLK_CUDA(cudaFuncSetAttribute(k_mlp_chain<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_BYTES));
    CRLK_CUDA(cudaFuncSetAttribute(k_mlp_chain<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM_BYTES));
    CRLK_CUDA(cudaFuncSetAttribute(k_gemm_bf16x3_pair<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<128>::smem));
    CRLK_CUDA(cudaFuncSetAttribute(k_gemm_bf16x3_pair<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<256>::smem));
    *out = t;
    return CRLK_OK;
}

void crlk_tc_destroy(CrlkPolicyTc *t)
{
    if (!t) return;
    for (int i = 0; i < t->n_hidden; i++)
        if (t->layer[i].d_w) cudaFree(t->layer[i].d_w);
    delete t;
}

size_t crlk_tc_workspace_bytes(const CrlkPolicyTc *t, int B)
{
    size_t Mpad = (size_t)rup(B, TC_BM);
    size_t planes = 3 * Mpad * t->max_width * sizeof(uint16_t);
    size_t last = Mpad * t->layer[t->n_hidden - 1].N * sizeof(float);
    // + the third plane buffer and the control words of the one-launch chain (k_mlp_chain)
    size_t ctrl = (64 + (size_t)TCC_MAX_LAYERS * (Mpad / TC_BM)) * sizeof(int);
    return 3 * planes + last + ((ctrl + 255) & ~(size_t)255) + 1024;
}

// The layer-0 operand planes inside a workspace: [3][Mpad][Kp0] bf16.
void crlk_tc_input_planes(const CrlkPolicyTc *t, void *workspace, int B, void **planes, int *Mpad, int *Kp)
{
    uint8_t *ws = (uint8_t *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    *planes = ws;
    *Mpad = rup(B, TC_BM);
    *Kp = t->layer[0].Kp;
}

// Hidden layers over operand planes that are already in the workspace
// (crlk_tc_input_planes).  n_rows_dev (nullable): live row count on the device.
// a_lo_kb: leading k blocks of layer 0 whose A planes 2/3 are non-zero (<= 0: all).
int crlk_tc_layers_forward(const CrlkPolicyTc *t, int B, void *workspace, const float *const *d_bias,
                           const int *n_rows_dev, int a_lo_kb, const float **out, int *ld_out, cudaStream_t st)
{
    const int Mpad = rup(B, TC_BM);
    uint8_t *ws = (uint8_t *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    size_t planes_bytes = 3 * (size_t)Mpad * t->max_width * sizeof(uint16_t);
    __nv_bfloat16 *pl[2] = {(__nv_bfloat16 *)ws, (__nv_bfloat16 *)(ws + planes_bytes)};
    float *last = (float *)(ws + 2 * planes_bytes);
    static const bool unfused = getenv("CRLK_MLP_UNFUSED") != nullptr;    // A/B switches
    static const bool persist = getenv("CRLK_MLP_NOPERSIST") == nullptr;
    static int n_sms = 0;
    if (!n_sms) { int dv = 0; cudaGetDevice(&dv); cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dv); }
    const bool need_persist = n_rows_dev != nullptr || a_lo_kb > 0 || t->nprod != 6;
    // Small and mid-size batches: all hidden layers in one dataflow launch (k_mlp_chain).  CRLK_MLP_CHAIN=0
    // restores one launch per layer (the A/B switch; results are identical -- same tiles, same k order).
    static const int chain_env = getenv("CRLK_MLP_CHAIN") ? atoi(getenv("CRLK_MLP_CHAIN")) : 1;
    static const int pair_forced = getenv("CRLK_MLP_PAIR") ? atoi(getenv("CRLK_MLP_PAIR")) : -1;
    if (chain_env && !unfused && persist && t->n_hidden <= TCC_MAX_LAYERS && pair_forced <= 0 &&
        (B < TC_PAIR_MIN_ROWS || pair_forced == 0)) {
        float *lastbuf = last;
        const size_t last_bytes = ((size_t)Mpad * t->layer[t->n_hidden - 1].N * sizeof(float) + 255) & ~(size_t)255;
        __nv_bfloat16 *X = pl[1], *Y = (__nv_bfloat16 *)((uint8_t *)last + last_bytes);
        int *ctrl = (int *)((uint8_t *)Y + planes_bytes);
        ChainP cp;
        memset(&cp, 0, sizeof cp);
        int total = 0;
        static const int bn_env = getenv("CRLK_MLP_BN") ? atoi(getenv("CRLK_MLP_BN")) : 256;
        // 256-column tiles pay once a layer is many waves of tiles (measured: 0.264 vs 0.298 ms at 20,000 rows,
        // but 0.158 vs 0.148 ms at 8,192 and 8.6 vs 7.6 ms per RL extension pass); CRLK_MLP_BN=128 / 256 forces
        static const bool bn_forced = getenv("CRLK_MLP_BN") != nullptr;
        bool wide = bn_env == 256 && t->nprod == 3 && (bn_forced || B >= 12288);
        for (int i = 0; i < t->n_hidden; i++) wide = wide && t->layer[i].N % 256 == 0;
        const int BNsel = wide ? 256 : TC_BN;
        for (int i = 0; i < t->n_hidden; i++) {
            const TcLayer &L = t->layer[i];
            const __nv_bfloat16 *in = i == 0 ? pl[0] : ((i - 1) % 2 == 0 ? X : Y);
            int rc = make_map(&cp.tmA[i], in, 3ull * Mpad, L.Kp, i == 0 ? L.Kp : t->max_width);
            if (rc) return rc;
            cp.tmW[i] = L.map_w;
            cp.bias[i] = d_bias[i];
            cp.out_planes[i] = (i == t->n_hidden - 1) ? nullptr : (i % 2 == 0 ? X : Y);
            cp.N[i] = L.N;
            cp.kblocks[i] = L.Kp / TC_BK;
            total += (L.N / BNsel) * (Mpad / TC_BM);
        }
        cp.last = lastbuf; cp.ld_last = t->layer[t->n_hidden - 1].N;
        cp.n_rows_dev = n_rows_dev; cp.ctrl = ctrl;
        cp.n_layers = t->n_hidden; cp.Mpad = Mpad; cp.B = B; cp.ld_planes = t->max_width;
        cp.a_lo_kb = a_lo_kb; cp.nprod = t->nprod; cp.row_tiles_max = Mpad / TC_BM;
        CRLK_CUDA(cudaMemsetAsync(ctrl, 0, (64 + (size_t)t->n_hidden * cp.row_tiles_max) * sizeof(int), st));
        if (wide) k_mlp_chain<256><<<total < n_sms ? total : n_sms, TC_THREADS, TCP_SMEM_BYTES, st>>>(cp);
        else k_mlp_chain<128><<<total < n_sms ? total : n_sms, TC_THREADS, TCP_SMEM_BYTES, st>>>(cp);
        CRLK_LAUNCH_CHECK();
        *out = last;
        *ld_out = t->layer[t->n_hidden - 1].N;
        return CRLK_OK;
    }
    for (int i = 0; i < t->n_hidden; i++) {
        const TcLayer &L = t->layer[i];
        CUtensorMap mapA;
        int rc = make_map(&mapA, pl[i & 1], 3ull * Mpad, L.Kp, L.Kp);
        if (rc) return rc;
        bool is_last = (i == t->n_hidden - 1);
        dim3 grid(L.N / TC_BN, Mpad / TC_BM);
        const int kblocks = L.Kp / TC_BK;
        // CTA-pair (cta_group::2) kernel with 256 x 256 tiles: 15 % faster than the single-CTA kernel once
        // every pair has several tiles to work through (measured: 203 vs 177 algorithmic TFLOP/s at 65,536
        // rows), slower below that (fewer, larger tiles per wave).  CRLK_MLP_PAIR=0 / 128 / 256 forces.
        static const int pair_env = getenv("CRLK_MLP_PAIR") ? atoi(getenv("CRLK_MLP_PAIR")) : -1;
        const int pair = pair_env >= 0 ? pair_env : (B >= TC_PAIR_MIN_ROWS ? 256 : 0);
        if (pair == 256 && Mpad % (2 * TC_BM) == 0 && L.N % 256 == 0) {
            int tiles = (L.N / 256) * (Mpad / (2 * TC_BM));
            int pairs = tiles < n_sms / 2 ? tiles : n_sms / 2;
            k_gemm_bf16x3_pair<256><<<2 * pairs, TC_THREADS, PairCfg<256>::smem, st>>>(
                mapA, L.map_w, Mpad, L.N, kblocks, d_bias[i], B,
                is_last ? nullptr : pl[(i + 1) & 1], L.N, is_last ? last : nullptr, L.N, n_rows_dev,
                (i == 0 && a_lo_kb > 0) ? a_lo_kb : kblocks, t->nprod);
        } else if (pair && Mpad % (2 * TC_BM) == 0) {
            int tiles = (L.N / TC_BN) * (Mpad / (2 * TC_BM));
            int pairs = tiles < n_sms / 2 ? tiles : n_sms / 2;
            k_gemm_bf16x3_pair<128><<<2 * pairs, TC_THREADS, PairCfg<128>::smem, st>>>(
                mapA, L.map_wh, Mpad, L.N, kblocks, d_bias[i], B,
                is_last ? nullptr : pl[(i + 1) & 1], L.N, is_last ? last : nullptr, L.N, n_rows_dev,
                (i == 0 && a_lo_kb > 0) ? a_lo_kb : kblocks, t->nprod);
        } else if (unfused && !need_persist)
            k_gemm_bf16x3<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(
                mapA, L.map_w, Mpad, L.N, kblocks, d_bias[i], B,
                is_last ? nullptr : pl[(i + 1) & 1], L.N, is_last ? last : nullptr, L.N);
        else if (persist || need_persist) {
            int tiles = (L.N / TC_BN) * (Mpad / TC_BM);
            int ctas = tiles < n_sms ? tiles : n_sms;
            k_gemm_bf16x3_persist<<<ctas, TC_THREADS, TCP_SMEM_BYTES, st>>>(
                mapA, L.map_w, Mpad, L.N, kblocks, d_bias[i], B,
                is_last ? nullptr : pl[(i + 1) & 1], L.N, is_last ? last : nullptr, L.N, n_rows_dev,
                (i == 0 && a_lo_kb > 0) ? a_lo_kb : kblocks, t->nprod);
        } else
            k_gemm_bf16x3_fused<<<grid, TC_THREADS, TCF_SMEM_BYTES, st>>>(
                mapA, L.map_w, Mpad, L.N, kblocks, d_bias[i], B,
                is_last ? nullptr : pl[(i + 1) & 1], L.N, is_last ? last : nullptr, L.N);
        CRLK_LAUNCH_CHECK();
    }
    *out = last;
    *ld_out = t->layer[t->n_hidden - 1].N;
    return CRLK_OK;
}

// 1 when crlk_tc_layers_forward runs a batch of B rows as one dataflow launch (k_mlp_chain).
int crlk_tc_chain_rows(int B)
{
    static const int chain_env = getenv("CRLK_MLP_CHAIN") ? atoi(getenv("CRLK_MLP_CHAIN")) : 1;
    static const int pair_forced = getenv("CRLK_MLP_PAIR") ? atoi(getenv("CRLK_MLP_PAIR")) : -1;
    return chain_env && pair_forced <= 0 && (B < TC_PAIR_MIN_ROWS || pair_forced == 0);
}

// Hidden layers: obs (float32) -> float32 activations of the last hidden layer.
// Returns the pointer / leading dimension of that activation through out/ld_out.
int crlk_tc_hidden_forward(const CrlkPolicyTc *t, const float *obs, int ld_obs, int B, void *workspace,
                           const float *const *d_bias, const float **out, int *ld_out, cudaStream_t st)
{
    const int Mpad = rup(B, TC_BM);
    uint8_t *ws = (uint8_t *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    // layer-0 operand planes from the float32 observation
    const TcLayer &L = t->layer[0];
    size_t total = (size_t)Mpad * L.Kp;
    k_split_planes<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(obs, ld_obs, L.K, B, Mpad, (__nv_bfloat16 *)ws,
                                                                   L.Kp);
    CRLK_LAUNCH_CHECK();
    return crlk_tc_layers_forward(t, B, workspace, d_bias, nullptr, 0, out, ld_out, st);
}
