// gslora-b200: fused multi-head self-attention forward / backward for the ViT token grid (N <= 208 tokens,
// head dim 64) -- Attention.forward of the reference (vit_pytorch_face/vit_face.py:358-379):
//     dots = einsum(q, k) * scale ; attn = softmax(dots) ; out = einsum(attn, v)
// The reference materialises [B, h, N, N] scores in HBM (636 MB per layer at bs 512); here one CTA owns one
// (image, head): Q, K, V (and dO) live in swizzled shared memory, the whole score row block stays in
// registers, and only O (+ the row log-sum-exp for the backward) goes back to HBM.
// Tensor-core path: mma.sync m16n8k16 (fp16 operands, fp32 accumulate) fed by ldmatrix.  N = 197 is one
// tile set, so there is no online-softmax loop.
// Backward is two passes over the same smem-resident tiles (no atomics, deterministic):
//   pass A  per 16-query tile : S, P, dP = dO V^T, dS = P (dP - delta)      -> dQ = scale * dS K
//   pass B  per 16-key tile   : S^T, P^T, dP^T                              -> dV = P^T dO, dK = scale * dS^T Q
#include "gsl_common.cuh"
#include "gsl_kernels.h"

namespace gsl {

static constexpr int ATT_WARPS = 7;
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

// cooperative load of a [N x 64] fp16 slab (row pitch ld) into a swizzled tile, zero padding rows >= N
__device__ __forceinline__ void load_tile(uint32_t sbase, const __half* __restrict__ g, int64_t ld, int N, int npad) {
    for (int i = threadIdx.x; i < npad * 8; i += blockDim.x) {
        const int row = i >> 3, chunk = i & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row < N) v = *reinterpret_cast<const uint4*>(g + (int64_t)row * ld + chunk * 8);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_addr(sbase, row, chunk)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(ATT_WARPS * 32, 1)
attention_fwd_kernel(const __half* __restrict__ qkv, int64_t ld, __half* __restrict__ out, int64_t ldo, float* __restrict__ lse,
                     int N, int heads, float scale) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int npad = (N + 15) & ~15;
    const int nkt = npad / 8;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int D = heads * 64;
    const uint32_t sQ = smem_u32(smem), sK = sQ + npad * 128, sV = sK + npad * 128;
    const __half* base = qkv + (int64_t)b * N * ld + h * 64;
    load_tile(sQ, base, ld, N, npad);
    load_tile(sK, base + D, ld, N, npad);
    load_tile(sV, base + 2 * D, ld, N, npad);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const float sl2 = scale * 1.4426950408889634f;

    for (int qt = warp; qt * 16 < npad; qt += ATT_WARPS) {
        uint32_t qf[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ldsm_x4(qf[kk], frag_a_addr(sQ, qt * 16, kk, lane));
        float s[ATT_MAX_KT][4];
#pragma unroll
        for (int j = 0; j < ATT_MAX_KT; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
            if (j < nkt) {
#pragma unroll
                for (int kc4 = 0; kc4 < 2; ++kc4) {
                    uint32_t bf[4];
                    ldsm_x4(bf, frag_b_addr(sK, j * 8, kc4, lane));
                    mma16816(s[j], qf[2 * kc4], bf[0], bf[1]);
                    mma16816(s[j], qf[2 * kc4 + 1], bf[2], bf[3]);
                }
            }
        }
        // softmax over the key axis: rows g (regs 0,1) and g+8 (regs 2,3)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < ATT_MAX_KT; ++j) {
            if (j < nkt) {
                const int c = j * 8 + 2 * t;
                if (c >= N) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
                if (c + 1 >= N) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
                mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
                mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int j = 0; j < ATT_MAX_KT; ++j) {
            if (j < nkt) {
                s[j][0] = exp2f((s[j][0] - mx0) * sl2); s[j][1] = exp2f((s[j][1] - mx0) * sl2);
                s[j][2] = exp2f((s[j][2] - mx1) * sl2); s[j][3] = exp2f((s[j][3] - mx1) * sl2);
                sum0 += s[j][0] + s[j][1];
                sum1 += s[j][2] + s[j][3];
            }
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

        float o[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < ATT_MAX_KT / 2; ++kk) {
            if (kk * 16 < npad) {
                uint32_t pa[4];
                pa[0] = pack_half2(s[2 * kk][0], s[2 * kk][1]);
                pa[1] = pack_half2(s[2 * kk][2], s[2 * kk][3]);
                pa[2] = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
                pa[3] = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
                for (int dn = 0; dn < 4; ++dn) {
                    uint32_t vf[4];
                    ldsm_x4_t(vf, frag_a_addr(sV, kk * 16, dn, lane));
                    mma16816(o[2 * dn], pa, vf[0], vf[1]);
                    mma16816(o[2 * dn + 1], pa, vf[2], vf[3]);
                }
            }
        }
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
}

int attention_fwd(const __half* qkv, int64_t ld, __half* out, int64_t ldo, float* lse, int B, int N, int heads, float scale, cudaStream_t s) {
    GSL_REQUIRE(N >= 1 && N <= ATT_MAX_TOKENS, "attention: tokens=%d outside [1, %d]", N, ATT_MAX_TOKENS);
    GSL_REQUIRE(ld % 8 == 0 && ldo % 2 == 0, "attention: qkv pitch must be a multiple of 8 halves");
    const int npad = (N + 15) & ~15;
    const int smem = 3 * npad * 128;
    static bool attr = false;
    if (!attr) {
        GSL_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * ATT_MAX_TOKENS * 128));
        attr = true;
    }
    attention_fwd_kernel<<<B * heads, ATT_WARPS * 32, smem, s>>>(qkv, ld, out, ldo, lse, N, heads, scale);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ backward
__global__ void __launch_bounds__(ATT_WARPS * 32, 1)
attention_bwd_kernel(const __half* __restrict__ qkv, int64_t ld, const __half* __restrict__ out, int64_t ldo, const __half* __restrict__ dout,
                     int64_t lddo, const float* __restrict__ lse, __half* __restrict__ dqkv, int64_t lddqkv, int N, int heads, float scale) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int npad = (N + 15) & ~15;
    const int nkt = npad / 8;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int D = heads * 64;
    const uint32_t sQ = smem_u32(smem), sK = sQ + npad * 128, sV = sK + npad * 128, sdO = sV + npad * 128;
    float* s_lse = reinterpret_cast<float*>(smem + 4 * npad * 128);      // LSE * log2(e)
    float* s_delta = s_lse + npad;
    const __half* base = qkv + (int64_t)b * N * ld + h * 64;
    load_tile(sQ, base, ld, N, npad);
    load_tile(sK, base + D, ld, N, npad);
    load_tile(sV, base + 2 * D, ld, N, npad);
    load_tile(sdO, dout + (int64_t)b * N * lddo + h * 64, lddo, N, npad);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const float sl2 = scale * 1.4426950408889634f;

    // delta[row] = sum_d dO[row, d] * O[row, d] ; one warp per row, 2 halves per lane
    for (int row = warp; row < npad; row += ATT_WARPS) {
        float d = 0.f;
        if (row < N) {
            const float2 a = unpack_half2(*reinterpret_cast<const uint32_t*>(dout + ((int64_t)b * N + row) * lddo + h * 64 + 2 * lane));
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

    // ---------------- pass A: dQ for 16-query tiles
    for (int qt = warp; qt * 16 < npad; qt += ATT_WARPS) {
        const int r0 = qt * 16 + g, r1 = r0 + 8;
        const float l0 = s_lse[r0], l1 = s_lse[r1], d0 = s_delta[r0], d1 = s_delta[r1];
        float dq[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
        for (int j0 = 0; j0 < nkt; j0 += 8) {          // 64 keys per chunk
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
                        uint32_t bf[4];
                        ldsm_x4(bf, frag_b_addr(sK, (j0 + j) * 8, kc4, lane));
                        mma16816(sc[j], qf0, bf[0], bf[1]);
                        mma16816(sc[j], qf1, bf[2], bf[3]);
                        ldsm_x4(bf, frag_b_addr(sV, (j0 + j) * 8, kc4, lane));
                        mma16816(dp[j], df0, bf[0], bf[1]);
                        mma16816(dp[j], df1, bf[2], bf[3]);
                    }
                }
            }
            // dS = P * (dP - delta), P = exp(scale * S - LSE); masked key columns give 0
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = (j0 + j) * 8 + 2 * t;
                const bool v0 = (j0 + j < nkt) && c < N, v1 = (j0 + j < nkt) && (c + 1) < N;
                sc[j][0] = v0 ? exp2f(sc[j][0] * sl2 - l0) * (dp[j][0] - d0) : 0.f;
                sc[j][1] = v1 ? exp2f(sc[j][1] * sl2 - l0) * (dp[j][1] - d0) : 0.f;
                sc[j][2] = v0 ? exp2f(sc[j][2] * sl2 - l1) * (dp[j][2] - d1) : 0.f;
                sc[j][3] = v1 ? exp2f(sc[j][3] * sl2 - l1) * (dp[j][3] - d1) : 0.f;
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

    // ---------------- pass B: dK, dV for 16-key tiles (rows of the transposed problem are keys)
    for (int kt = warp; kt * 16 < npad; kt += ATT_WARPS) {
        const int k0 = kt * 16 + g, k1 = k0 + 8;
        float dk[8][4], dv[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
            dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
        }
        for (int j0 = 0; j0 < nkt; j0 += 8) {          // 64 queries per chunk
            float st[8][4], dpt[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
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
                for (int j = 0; j < 8; ++j) {
                    if (j0 + j < nkt) {
                        uint32_t bf[4];
                        ldsm_x4(bf, frag_b_addr(sQ, (j0 + j) * 8, kc4, lane));
                        mma16816(st[j], kf0, bf[0], bf[1]);
                        mma16816(st[j], kf1, bf[2], bf[3]);
                        ldsm_x4(bf, frag_b_addr(sdO, (j0 + j) * 8, kc4, lane));
                        mma16816(dpt[j], vf0, bf[0], bf[1]);
                        mma16816(dpt[j], vf1, bf[2], bf[3]);
                    }
                }
            }
            // columns are queries: P^T[k, q] = exp(scale * S^T - LSE[q]); dS^T = P^T * (dP^T - delta[q])
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int q = (j0 + j) * 8 + 2 * t;
                const bool live = (j0 + j) < nkt;
                const bool v0 = live && q < N, v1 = live && (q + 1) < N;
                const float lq0 = live ? s_lse[q] : 0.f, lq1 = live ? s_lse[q + 1] : 0.f;
                const float dq0 = live ? s_delta[q] : 0.f, dq1 = live ? s_delta[q + 1] : 0.f;
                const float p00 = (v0 && k0 < N) ? exp2f(st[j][0] * sl2 - lq0) : 0.f;
                const float p01 = (v1 && k0 < N) ? exp2f(st[j][1] * sl2 - lq1) : 0.f;
                const float p10 = (v0 && k1 < N) ? exp2f(st[j][2] * sl2 - lq0) : 0.f;
                const float p11 = (v1 && k1 < N) ? exp2f(st[j][3] * sl2 - lq1) : 0.f;
                st[j][0] = p00; st[j][1] = p01; st[j][2] = p10; st[j][3] = p11;
                dpt[j][0] = p00 * (dpt[j][0] - dq0); dpt[j][1] = p01 * (dpt[j][1] - dq1);
                dpt[j][2] = p10 * (dpt[j][2] - dq0); dpt[j][3] = p11 * (dpt[j][3] - dq1);
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
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
                        uint32_t bf[4];
                        ldsm_x4_t(bf, frag_a_addr(sdO, (j0 + 2 * kk) * 8, dn, lane));
                        mma16816(dv[2 * dn], pa, bf[0], bf[1]);
                        mma16816(dv[2 * dn + 1], pa, bf[2], bf[3]);
                        ldsm_x4_t(bf, frag_a_addr(sQ, (j0 + 2 * kk) * 8, dn, lane));
                        mma16816(dk[2 * dn], da, bf[0], bf[1]);
                        mma16816(dk[2 * dn + 1], da, bf[2], bf[3]);
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
}

int attention_bwd(const __half* qkv, int64_t ld, const __half* out, int64_t ldo, const __half* dout, int64_t lddo, const float* lse,
                  __half* dqkv, int64_t lddqkv, int B, int N, int heads, float scale, cudaStream_t s) {
    GSL_REQUIRE(N >= 1 && N <= ATT_MAX_TOKENS, "attention_bwd: tokens=%d outside [1, %d]", N, ATT_MAX_TOKENS);
    GSL_REQUIRE(ld % 8 == 0 && lddo % 8 == 0 && ldo % 2 == 0 && lddqkv % 2 == 0, "attention_bwd: pitches must be multiples of 8 halves");
    const int npad = (N + 15) & ~15;
    const int smem = 4 * npad * 128 + 2 * npad * 4;
    static bool attr = false;
    if (!attr) {
        GSL_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            4 * ATT_MAX_TOKENS * 128 + 2 * ATT_MAX_TOKENS * 4));
        attr = true;
    }
    attention_bwd_kernel<<<B * heads, ATT_WARPS * 32, smem, s>>>(qkv, ld, out, ldo, dout, lddo, lse, dqkv, lddqkv, N, heads, scale);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gsl
