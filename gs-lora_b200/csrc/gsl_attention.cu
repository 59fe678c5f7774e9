// gslora-b200: fused multi-head self-attention forward / backward for the ViT token grid (N <= 208 tokens,
// head dim 64) -- Attention.forward of the reference (vit_pytorch_face/vit_face.py:358-379):
//     dots = einsum(q, k) * scale ; attn = softmax(dots) ; out = einsum(attn, v)
// The reference materialises [B, h, N, N] scores in HBM (636 MB per layer at bs 512); here a persistent CTA walks over
// (image, head) pairs: Q, K, V (and dO) slabs arrive by TMA (3-D tensor map, 128B swizzle, rows >= N zero-filled) into
// a double-buffered shared-memory ring so the next pair's loads overlap this pair's MMAs, the score row block stays in
// registers, and only O (+ the row log-sum-exp for the backward) goes back to HBM.
// Tensor-core path: mma.sync m16n8k16 (fp16 operands, fp32 accumulate) fed by ldmatrix.  N = 197 is one
// tile set, so there is no online-softmax loop.
// Backward is two passes over the same smem-resident tiles (no atomics, deterministic):
//   pass A  per 16-query tile : S, P, dP = dO V^T, dS = P (dP - delta)      -> dQ = scale * dS K
//   pass B  per 16-key tile   : S^T, P^T, dP^T                              -> dV = P^T dO, dK = scale * dS^T Q
#include "gsl_common.cuh"
#include <stdlib.h>
#include <cuda.h>
#include "gsl_kernels.h"

namespace gsl {

static constexpr int ATT_WARPS = 13;                    // one 16-row tile per warp (13 * 16 = 208 >= 197 tokens)
static constexpr int ATT_MAX_TOKENS = 208;
static constexpr int ATT_MAX_KT = ATT_MAX_TOKENS / 8;   // 26 key n-tiles of 8

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// tile layout: row-major [rows][64 halves] = 128-byte rows, 16-byte chunks XOR-swizzled with (row & 7)
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
    return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}
// A fragment (16 rows x 16 k) at (row0, k-chunk pair kc) -- also the address pattern of a transposed B fragment pair
__device__ __forceinline__ uint32_t frag_a_addr(uint32_t base, int row0, int kc, int lane) {
    return tile_addr(base, row0 + (lane & 7) + ((lane >> 3) & 1) * 8, kc * 2 + (lane >> 4));
}
// B fragments (non transposed) of one 8-row n-tile for two consecutive k-steps (k chunks kc4*4 .. kc4*4+3)
__device__ __forceinline__ uint32_t frag_b_addr(uint32_t base, int n0, int kc4, int lane) {
    return tile_addr(base, n0 + (lane & 7), kc4 * 4 + (lane >> 3));
}

// one [npad x 64] fp16 slab (rows >= N zero-filled by TMA) of column block `col` of image b
__device__ __forceinline__ void tma_load_slab(const void* desc, uint32_t bar, uint32_t dst, int col, int b) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(col), "r"(0), "r"(b) : "memory");
}

// ------------------------------------------------------------------------------------------------ forward
// Each warp owns 16 query rows.  The key axis is processed in two halves (<= 112 + 96 keys) with an online-softmax
// rescale between them so the score fragment is 56 registers instead of 104: two CTAs (14 warps) fit per SM and one
// CTA's global->shared fill overlaps the other's MMAs.
static constexpr int ATT_HALF0 = 14;    // n-tiles (of 8 keys) in the first half

template <int NT>
__device__ __forceinline__ void attn_fwd_half(const uint32_t (&qf)[4][4], uint32_t sK, uint32_t sV, int j0, int nt, int N, int lane, float sl2,
                                              float& mx0, float& mx1, float& sum0, float& sum1, float (&o)[8][4]) {
    const int t = lane & 3;
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kc4 = 0; kc4 < 2; ++kc4) {
#pragma unroll
        for (int j = 0; j < NT; j += 2) {
            if (j < nt) {
                uint32_t ba[4], bb[4];
                ldsm_x4(ba, frag_b_addr(sK, (j0 + j) * 8, kc4, lane));
                ldsm_x4(bb, frag_b_addr(sK, (j0 + j + 1) * 8, kc4, lane));     // nt is even (npad % 16 == 0)
                mma16816(s[j], qf[2 * kc4], ba[0], ba[1]);
                mma16816(s[j + 1], qf[2 * kc4], bb[0], bb[1]);
                mma16816(s[j], qf[2 * kc4 + 1], ba[2], ba[3]);
                mma16816(s[j + 1], qf[2 * kc4 + 1], bb[2], bb[3]);
            }
        }
    }
    float nm0 = mx0, nm1 = mx1;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        if (j < nt) {
            const int c = (j0 + j) * 8 + 2 * t;
            if (c >= N) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
            if (c + 1 >= N) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
            nm0 = fmaxf(nm0, fmaxf(s[j][0], s[j][1]));
            nm1 = fmaxf(nm1, fmaxf(s[j][2], s[j][3]));
        }
    }
    nm0 = fmaxf(nm0, __shfl_xor_sync(0xffffffffu, nm0, 1)); nm0 = fmaxf(nm0, __shfl_xor_sync(0xffffffffu, nm0, 2));
    nm1 = fmaxf(nm1, __shfl_xor_sync(0xffffffffu, nm1, 1)); nm1 = fmaxf(nm1, __shfl_xor_sync(0xffffffffu, nm1, 2));
    // rescale what has been accumulated so far (first half: mx = -inf, everything is still zero)
    const float a0 = mx0 == -INFINITY ? 0.f : exp2f((mx0 - nm0) * sl2);
    const float a1 = mx1 == -INFINITY ? 0.f : exp2f((mx1 - nm1) * sl2);
    mx0 = nm0; mx1 = nm1;
    sum0 *= a0; sum1 *= a1;
#pragma unroll
    for (int n = 0; n < 8; ++n) { o[n][0] *= a0; o[n][1] *= a0; o[n][2] *= a1; o[n][3] *= a1; }
    const float off0 = nm0 * sl2, off1 = nm1 * sl2;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        if (j < nt) {
            s[j][0] = exp2f(fmaf(s[j][0], sl2, -off0)); s[j][1] = exp2f(fmaf(s[j][1], sl2, -off0));
            s[j][2] = exp2f(fmaf(s[j][2], sl2, -off1)); s[j][3] = exp2f(fmaf(s[j][3], sl2, -off1));
            sum0 += s[j][0] + s[j][1];
            sum1 += s[j][2] + s[j][3];
        }
    }
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
        if (2 * kk < nt) {
            uint32_t pa[4];
            pa[0] = pack_half2(s[2 * kk][0], s[2 * kk][1]);
            pa[1] = pack_half2(s[2 * kk][2], s[2 * kk][3]);
            pa[2] = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            pa[3] = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int dn = 0; dn < 4; ++dn) {
                uint32_t vf[4];
                ldsm_x4_t(vf, frag_a_addr(sV, (j0 + 2 * kk) * 8, dn, lane));
                mma16816(o[2 * dn], pa, vf[0], vf[1]);
                mma16816(o[2 * dn + 1], pa, vf[2], vf[3]);
            }
        }
    }
}

__global__ void __maxnreg__(128)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, __half* __restrict__ out, int64_t ldo, float* __restrict__ lse,
                     int B, int N, int heads, float scale) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int npad = (N + 15) & ~15;
    const int nkt = npad / 8;
    const int D = heads * 64;
    const uint32_t tile_bytes = npad * 128;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + 2 * 3 * tile_bytes;
    const int nwork = B * heads;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(bars, 1);
        mbar_init(bars + 8, 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int w, int buf) {
        const int wb = w / heads, wh = w % heads;
        const uint32_t dst = sbase + buf * 3 * tile_bytes;
        mbar_arrive_expect_tx(bars + 8 * buf, 3 * tile_bytes);
        tma_load_slab(&tmQKV, bars + 8 * buf, dst, wh * 64, wb);
        tma_load_slab(&tmQKV, bars + 8 * buf, dst + tile_bytes, D + wh * 64, wb);
        tma_load_slab(&tmQKV, bars + 8 * buf, dst + 2 * tile_bytes, 2 * D + wh * 64, wb);
    };
    if (threadIdx.x == 0 && (int)blockIdx.x < nwork) issue(blockIdx.x, 0);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const float sl2 = scale * 1.4426950408889634f;
    const int nt0 = nkt < ATT_HALF0 ? nkt : ATT_HALF0;
    const int nt1 = nkt - nt0;

    int it = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
        const int b = w / heads, h = w % heads;
        if (threadIdx.x == 0 && w + (int)gridDim.x < nwork) issue(w + gridDim.x, (it + 1) & 1);
        mbar_wait(bars + 8 * (it & 1), (it >> 1) & 1);
        const uint32_t sQ = sbase + (it & 1) * 3 * tile_bytes, sK = sQ + tile_bytes, sV = sK + tile_bytes;
        const int qt = warp;
        if (qt * 16 < npad) {
            uint32_t qf[4][4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) ldsm_x4(qf[kk], frag_a_addr(sQ, qt * 16, kk, lane));
            float o[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
            float mx0 = -INFINITY, mx1 = -INFINITY, sum0 = 0.f, sum1 = 0.f;
            attn_fwd_half<ATT_HALF0>(qf, sK, sV, 0, nt0, N, lane, sl2, mx0, mx1, sum0, sum1, o);
            if (nt1 > 0) attn_fwd_half<ATT_MAX_KT - ATT_HALF0>(qf, sK, sV, ATT_HALF0, nt1, N, lane, sl2, mx0, mx1, sum0, sum1, o);
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
            const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
            const int r0 = qt * 16 + g, r1 = r0 + 8;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                if (r0 < N) *reinterpret_cast<uint32_t*>(out + ((int64_t)b * N + r0) * ldo + h * 64 + n * 8 + 2 * t) = pack_half2(o[n][0] * inv0, o[n][1] * inv0);
                if (r1 < N) *reinterpret_cast<uint32_t*>(out + ((int64_t)b * N + r1) * ldo + h * 64 + n * 8 + 2 * t) = pack_half2(o[n][2] * inv1, o[n][3] * inv1);
            }
            if (t == 0 && lse != nullptr) {
                if (r0 < N) lse[((int64_t)b * heads + h) * N + r0] = mx0 * scale + logf(sum0);
                if (r1 < N) lse[((int64_t)b * heads + h) * N + r1] = mx1 * scale + logf(sum1);
            }
        }
        __syncthreads();      // everyone is done with this buffer before it is refilled two iterations later
    }
}

int make_tmap_qkv(CUtensorMap* map, const void* ptr, int64_t ld, int B, int N, int cols, int npad);

static int attention_grid(int nwork) {
    const int sms = device_sm_count();
    return nwork < sms ? nwork : sms;
}

int attention_fwd_tc(const __half* qkv, int64_t ld, __half* out, int64_t ldo, float* lse, int B, int N, int heads, float scale, cudaStream_t s);

// GSLORA_ATTN=mma selects the mma.sync kernels below; the default is the tcgen05 forward (gsl_attention_tc.cu)
static bool attention_use_tc() {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("GSLORA_ATTN"); mode = (e && e[0] == 'm') ? 0 : 1; }
    return mode == 1;
}

int attention_fwd(const __half* qkv, int64_t ld, __half* out, int64_t ldo, float* lse, int B, int N, int heads, float scale, cudaStream_t s) {
    if (attention_use_tc()) return attention_fwd_tc(qkv, ld, out, ldo, lse, B, N, heads, scale, s);
    GSL_REQUIRE(N >= 1 && N <= ATT_MAX_TOKENS, "attention: tokens=%d outside [1, %d]", N, ATT_MAX_TOKENS);
    GSL_REQUIRE(ld % 8 == 0 && ldo % 2 == 0, "attention: qkv pitch must be a multiple of 8 halves");
    const int npad = (N + 15) & ~15;
    CUtensorMap tm;
    int rc = make_tmap_qkv(&tm, qkv, ld, B, N, 3 * heads * 64, npad);
    if (rc) return rc;
    const int smem = 1024 + 2 * 3 * npad * 128 + 64;
    static bool attr = false;
    if (!attr) {
        GSL_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 2 * 3 * ATT_MAX_TOKENS * 128 + 64));
        attr = true;
    }
    attention_fwd_kernel<<<attention_grid(B * heads), ATT_WARPS * 32, smem, s>>>(tm, out, ldo, lse, B, N, heads, scale);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ backward
__global__ void __maxnreg__(128)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const __half* __restrict__ out, int64_t ldo,
                     const float* __restrict__ lse, __half* __restrict__ dqkv, int64_t lddqkv, int B, int N, int heads, float scale) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int npad = (N + 15) & ~15;
    const int nkt = npad / 8;
    const int D = heads * 64;
    const uint32_t tile_bytes = npad * 128;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + 2 * 4 * tile_bytes;
    float* s_lse = reinterpret_cast<float*>(smem + 2 * 4 * tile_bytes + 64);      // LSE * log2(e)
    float* s_delta = s_lse + npad;
    const int nwork = B * heads;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
        mbar_init(bars, 1);
        mbar_init(bars + 8, 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int w, int buf) {
        const int wb = w / heads, wh = w % heads;
        const uint32_t dst = sbase + buf * 4 * tile_bytes;
        mbar_arrive_expect_tx(bars + 8 * buf, 4 * tile_bytes);
        tma_load_slab(&tmQKV, bars + 8 * buf, dst, wh * 64, wb);
        tma_load_slab(&tmQKV, bars + 8 * buf, dst + tile_bytes, D + wh * 64, wb);
        tma_load_slab(&tmQKV, bars + 8 * buf, dst + 2 * tile_bytes, 2 * D + wh * 64, wb);
        tma_load_slab(&tmDO, bars + 8 * buf, dst + 3 * tile_bytes, wh * 64, wb);
    };
    if (threadIdx.x == 0 && (int)blockIdx.x < nwork) issue(blockIdx.x, 0);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const float sl2 = scale * 1.4426950408889634f;

  int it = 0;
  for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
    const int b = w / heads, h = w % heads;
    if (threadIdx.x == 0 && w + (int)gridDim.x < nwork) issue(w + gridDim.x, (it + 1) & 1);
    mbar_wait(bars + 8 * (it & 1), (it >> 1) & 1);
    const uint32_t sQ = sbase + (it & 1) * 4 * tile_bytes, sK = sQ + tile_bytes, sV = sK + tile_bytes, sdO = sV + tile_bytes;

    // delta[row] = sum_d dO[row, d] * O[row, d] ; one warp per row, 2 halves per lane (dO from the smem slab, O from HBM)
    for (int row = warp; row < npad; row += ATT_WARPS) {
        float d = 0.f;
        if (row < N) {
            uint32_t dov;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(dov) : "r"(tile_addr(sdO, row, lane >> 2) + (lane & 3) * 4));
            const float2 a = unpack_half2(dov);
            const float2 o = unpack_half2(*reinterpret_cast<const uint32_t*>(out + ((int64_t)b * N + row) * ldo + h * 64 + 2 * lane));
            d = a.x * o.x + a.y * o.y;
        }
        d = warp_sum(d);
        if (lane == 0) {
            s_delta[row] = d;
            s_lse[row] = row < N ? lse[((int64_t)b * heads + h) * N + row] * 1.4426950408889634f : 0.f;
        }
    }
    __syncthreads();

    // ---------------- pass A: dQ for 16-query tiles, 64 keys per chunk
    for (int qt = warp; qt * 16 < npad; qt += ATT_WARPS) {
        const int r0 = qt * 16 + g, r1 = r0 + 8;
        const float l0 = s_lse[r0], l1 = s_lse[r1], d0 = s_delta[r0], d1 = s_delta[r1];
        float dq[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
        for (int j0 = 0; j0 < nkt; j0 += 8) {
            float sc[8][4], dp[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
                dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
            }
#pragma unroll
            for (int kc4 = 0; kc4 < 2; ++kc4) {
                uint32_t qf0[4], qf1[4], df0[4], df1[4];
                ldsm_x4(qf0, frag_a_addr(sQ, qt * 16, 2 * kc4, lane));
                ldsm_x4(qf1, frag_a_addr(sQ, qt * 16, 2 * kc4 + 1, lane));
                ldsm_x4(df0, frag_a_addr(sdO, qt * 16, 2 * kc4, lane));
                ldsm_x4(df1, frag_a_addr(sdO, qt * 16, 2 * kc4 + 1, lane));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j0 + j < nkt) {
                        uint32_t bk[4], bv[4];
                        ldsm_x4(bk, frag_b_addr(sK, (j0 + j) * 8, kc4, lane));
                        ldsm_x4(bv, frag_b_addr(sV, (j0 + j) * 8, kc4, lane));
                        mma16816(sc[j], qf0, bk[0], bk[1]);
                        mma16816(dp[j], df0, bv[0], bv[1]);
                        mma16816(sc[j], qf1, bk[2], bk[3]);
                        mma16816(dp[j], df1, bv[2], bv[3]);
                    }
                }
            }
            // dS = P * (dP - delta), P = exp(scale * S - LSE); masked key columns give 0
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = (j0 + j) * 8 + 2 * t;
                const bool v0 = (j0 + j < nkt) && c < N, v1 = (j0 + j < nkt) && (c + 1) < N;
                sc[j][0] = v0 ? exp2f(fmaf(sc[j][0], sl2, -l0)) * (dp[j][0] - d0) : 0.f;
                sc[j][1] = v1 ? exp2f(fmaf(sc[j][1], sl2, -l0)) * (dp[j][1] - d0) : 0.f;
                sc[j][2] = v0 ? exp2f(fmaf(sc[j][2], sl2, -l1)) * (dp[j][2] - d1) : 0.f;
                sc[j][3] = v1 ? exp2f(fmaf(sc[j][3], sl2, -l1)) * (dp[j][3] - d1) : 0.f;
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if ((j0 + 2 * kk) < nkt) {
                    uint32_t pa[4];
                    pa[0] = pack_half2(sc[2 * kk][0], sc[2 * kk][1]);
                    pa[1] = pack_half2(sc[2 * kk][2], sc[2 * kk][3]);
                    pa[2] = pack_half2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
                    pa[3] = pack_half2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
                    for (int dn = 0; dn < 4; ++dn) {
                        uint32_t kf[4];
                        ldsm_x4_t(kf, frag_a_addr(sK, (j0 + 2 * kk) * 8, dn, lane));
                        mma16816(dq[2 * dn], pa, kf[0], kf[1]);
                        mma16816(dq[2 * dn + 1], pa, kf[2], kf[3]);
                    }
                }
            }
        }
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            if (r0 < N) *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)b * N + r0) * lddqkv + h * 64 + n * 8 + 2 * t) = pack_half2(dq[n][0] * scale, dq[n][1] * scale);
            if (r1 < N) *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)b * N + r1) * lddqkv + h * 64 + n * 8 + 2 * t) = pack_half2(dq[n][2] * scale, dq[n][3] * scale);
        }
    }

    // ---------------- pass B: dK, dV for 16-key tiles (rows of the transposed problem are keys), 32 queries per chunk
    for (int kt = warp; kt * 16 < npad; kt += ATT_WARPS) {
        const int k0 = kt * 16 + g, k1 = k0 + 8;
        float dk[8][4], dv[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
            dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
        }
        for (int j0 = 0; j0 < nkt; j0 += 4) {
            float st[4][4], dpt[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
                dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
            }
#pragma unroll
            for (int kc4 = 0; kc4 < 2; ++kc4) {
                uint32_t kf0[4], kf1[4], vf0[4], vf1[4];
                ldsm_x4(kf0, frag_a_addr(sK, kt * 16, 2 * kc4, lane));
                ldsm_x4(kf1, frag_a_addr(sK, kt * 16, 2 * kc4 + 1, lane));
                ldsm_x4(vf0, frag_a_addr(sV, kt * 16, 2 * kc4, lane));
                ldsm_x4(vf1, frag_a_addr(sV, kt * 16, 2 * kc4 + 1, lane));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j0 + j < nkt) {
                        uint32_t bq[4], bd[4];
                        ldsm_x4(bq, frag_b_addr(sQ, (j0 + j) * 8, kc4, lane));
                        ldsm_x4(bd, frag_b_addr(sdO, (j0 + j) * 8, kc4, lane));
                        mma16816(st[j], kf0, bq[0], bq[1]);
                        mma16816(dpt[j], vf0, bd[0], bd[1]);
                        mma16816(st[j], kf1, bq[2], bq[3]);
                        mma16816(dpt[j], vf1, bd[2], bd[3]);
                    }
                }
            }
            // columns are queries: P^T[k, q] = exp(scale * S^T - LSE[q]); dS^T = P^T * (dP^T - delta[q])
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int q = (j0 + j) * 8 + 2 * t;
                const bool live = (j0 + j) < nkt;
                const bool v0 = live && q < N, v1 = live && (q + 1) < N;
                const float lq0 = live ? s_lse[q] : 0.f, lq1 = live ? s_lse[q + 1] : 0.f;
                const float dq0 = live ? s_delta[q] : 0.f, dq1 = live ? s_delta[q + 1] : 0.f;
                const float p00 = (v0 && k0 < N) ? exp2f(fmaf(st[j][0], sl2, -lq0)) : 0.f;
                const float p01 = (v1 && k0 < N) ? exp2f(fmaf(st[j][1], sl2, -lq1)) : 0.f;
                const float p10 = (v0 && k1 < N) ? exp2f(fmaf(st[j][2], sl2, -lq0)) : 0.f;
                const float p11 = (v1 && k1 < N) ? exp2f(fmaf(st[j][3], sl2, -lq1)) : 0.f;
                st[j][0] = p00; st[j][1] = p01; st[j][2] = p10; st[j][3] = p11;
                dpt[j][0] = p00 * (dpt[j][0] - dq0); dpt[j][1] = p01 * (dpt[j][1] - dq1);
                dpt[j][2] = p10 * (dpt[j][2] - dq0); dpt[j][3] = p11 * (dpt[j][3] - dq1);
            }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                if ((j0 + 2 * kk) < nkt) {
                    uint32_t pa[4], da[4];
                    pa[0] = pack_half2(st[2 * kk][0], st[2 * kk][1]);
                    pa[1] = pack_half2(st[2 * kk][2], st[2 * kk][3]);
                    pa[2] = pack_half2(st[2 * kk + 1][0], st[2 * kk + 1][1]);
                    pa[3] = pack_half2(st[2 * kk + 1][2], st[2 * kk + 1][3]);
                    da[0] = pack_half2(dpt[2 * kk][0], dpt[2 * kk][1]);
                    da[1] = pack_half2(dpt[2 * kk][2], dpt[2 * kk][3]);
                    da[2] = pack_half2(dpt[2 * kk + 1][0], dpt[2 * kk + 1][1]);
                    da[3] = pack_half2(dpt[2 * kk + 1][2], dpt[2 * kk + 1][3]);
#pragma unroll
                    for (int dn = 0; dn < 4; ++dn) {
                        uint32_t bo[4], bq[4];
                        ldsm_x4_t(bo, frag_a_addr(sdO, (j0 + 2 * kk) * 8, dn, lane));
                        ldsm_x4_t(bq, frag_a_addr(sQ, (j0 + 2 * kk) * 8, dn, lane));
                        mma16816(dv[2 * dn], pa, bo[0], bo[1]);
                        mma16816(dk[2 * dn], da, bq[0], bq[1]);
                        mma16816(dv[2 * dn + 1], pa, bo[2], bo[3]);
                        mma16816(dk[2 * dn + 1], da, bq[2], bq[3]);
                    }
                }
            }
        }
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int col = h * 64 + n * 8 + 2 * t;
            if (k0 < N) {
                *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)b * N + k0) * lddqkv + D + col) = pack_half2(dk[n][0] * scale, dk[n][1] * scale);
                *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)b * N + k0) * lddqkv + 2 * D + col) = pack_half2(dv[n][0], dv[n][1]);
            }
            if (k1 < N) {
                *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)b * N + k1) * lddqkv + D + col) = pack_half2(dk[n][2] * scale, dk[n][3] * scale);
                *reinterpret_cast<uint32_t*>(dqkv + ((int64_t)b * N + k1) * lddqkv + 2 * D + col) = pack_half2(dv[n][2], dv[n][3]);
            }
        }
    }
    __syncthreads();      // buffer + s_lse / s_delta are free for the next (image, head)
  }
}

int attention_bwd_tc(const __half* qkv, int64_t ld, const __half* out, int64_t ldo, const __half* dout, int64_t lddo, const float* lse,
                     __half* dqkv, int64_t lddqkv, int B, int N, int heads, float scale, cudaStream_t s);

int attention_bwd(const __half* qkv, int64_t ld, const __half* out, int64_t ldo, const __half* dout, int64_t lddo, const float* lse,
                  __half* dqkv, int64_t lddqkv, int B, int N, int heads, float scale, cudaStream_t s) {
    if (attention_use_tc()) return attention_bwd_tc(qkv, ld, out, ldo, dout, lddo, lse, dqkv, lddqkv, B, N, heads, scale, s);
    GSL_REQUIRE(N >= 1 && N <= ATT_MAX_TOKENS, "attention_bwd: tokens=%d outside [1, %d]", N, ATT_MAX_TOKENS);
    GSL_REQUIRE(ld % 8 == 0 && lddo % 8 == 0 && ldo % 2 == 0 && lddqkv % 2 == 0, "attention_bwd: pitches must be multiples of 8 halves");
    const int npad = (N + 15) & ~15;
    CUtensorMap tq, td;
    int rc;
    if ((rc = make_tmap_qkv(&tq, qkv, ld, B, N, 3 * heads * 64, npad))) return rc;
    if ((rc = make_tmap_qkv(&td, dout, lddo, B, N, heads * 64, npad))) return rc;
    const int smem = 1024 + 2 * 4 * npad * 128 + 64 + 2 * npad * 4;
    static bool attr = false;
    if (!attr) {
        GSL_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            1024 + 2 * 4 * ATT_MAX_TOKENS * 128 + 64 + 2 * ATT_MAX_TOKENS * 4));
        attr = true;
    }
    attention_bwd_kernel<<<attention_grid(B * heads), ATT_WARPS * 32, smem, s>>>(tq, td, out, ldo, lse, dqkv, lddqkv, B, N, heads, scale);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gsl
