// gslora-b200: HBM-bound row kernels around the GEMMs -- patchify, LayerNorm forward/backward,
// the skinny LoRA contractions (T = X A^T, U = dY B, dA/dB reductions over the token dimension),
// and the fp32->fp16 weight-cache casts.  All are sized by bytes moved, not FLOPs: 16-byte vector
// loads, one warp per row (or per 16 rows for the mma-based skinny products), fp32 arithmetic.
#include "gsl_common.cuh"
#include "gsl_kernels.h"
#include <cuda_fp8.h>

namespace gsl {

// ------------------------------------------------------------------------------------------------ patchify
// Reference: einops 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' (vit_pytorch_face/vit_face.py:530) and, for the
// torchvision family, conv_proj's implicit (c p1 p2) patch vector.  Token 0 (cls slot) is a zero row so that the
// patch-embedding GEMM's rows line up 1:1 with the [B, tokens, D] residual stream.
// T = float: the reference's loader output (transforms.ToTensor(): fp32 NCHW in [0, 1]).
// T = uint8_t: raw pixels, NCHW (layout 0, transforms.PILToTensor()) or NHWC (layout 1, decoded image rows); ToTensor's `/ 255` and the optional
// transforms.Normalize(mean, std) of the ImageNet configs (train_own_forget_cl.py:138-139) are applied on the fly with IEEE division, so
// the fp32 value that gets rounded to fp16 is bit-identical to what the reference's transform pipeline would have produced on the host.
struct PixelNorm { float mean[4]; float std[4]; int enabled; };

// One CTA per (image, patch row): the C x patch x S strip of the image is read ONCE with coalesced loads into shared memory (converted to
// fp32 there: the u8 arithmetic runs once per pixel), then the w patch vectors of that row leave as coalesced half2 stores; a per-CTA
// table maps the patch-vector element e to its strip offset, so the scatter costs one shared-memory lookup per element instead of the
// integer divisions.  HBM-bound: reads the image once, writes the patch matrix once.
template <typename T>
__global__ void __launch_bounds__(256) patchify_kernel(const T* __restrict__ img, __half* __restrict__ out, int64_t ld, int B, int C, int S,
                                                       int patch, int order, int layout, PixelNorm nrm) {
    pdl_prologue();
    extern __shared__ __align__(16) uint8_t pf_smem[];
    const int w = S / patch, P = w * w, pd = C * patch * patch, half_pd = pd >> 1;
    const int plane = patch * S;                          // one channel of the strip: [patch rows][S pixels]
    float* strip = reinterpret_cast<float*>(pf_smem);     // [C][patch][S]
    int* lut = reinterpret_cast<int*>(strip + C * plane); // e -> c * plane + p1 * S + p2
    const int b = blockIdx.x / w, ph = blockIdx.x % w;
    const int tid = threadIdx.x;
    for (int e = tid; e < pd; e += blockDim.x) {
        int c, p1, p2;
        if (order == 0) { c = e % C; p2 = (e / C) % patch; p1 = e / (C * patch); }
        else { p2 = e % patch; p1 = (e / patch) % patch; c = e / (patch * patch); }
        lut[e] = c * plane + p1 * S + p2;
    }
    auto conv = [&](T v, int c) -> float {
        if constexpr (sizeof(T) == 1) {
            float f = __fdiv_rn((float)v, 255.f);
            if (nrm.enabled) f = __fdiv_rn(f - nrm.mean[c & 3], nrm.std[c & 3]);
            return f;
        } else {
            return v;
        }
    };
    if (layout == 0) {          // NCHW: channel c contributes `patch` contiguous image rows = plane contiguous elements
        for (int c = 0; c < C; ++c) {
            const T* src = img + (((int64_t)b * C + c) * S + (int64_t)ph * patch) * S;
            if constexpr (sizeof(T) == 4) {
                if ((plane & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
                    for (int i = tid; i < (plane >> 2); i += blockDim.x)
                        reinterpret_cast<float4*>(strip + c * plane)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
                    continue;
                }
            }
            for (int i = tid; i < plane; i += blockDim.x) strip[c * plane + i] = conv(__ldg(src + i), c);
        }
    } else {                    // NHWC: `patch` contiguous image rows of S * C interleaved elements
        const T* src = img + ((int64_t)b * S + (int64_t)ph * patch) * S * C;
        const int n = plane * C;
        for (int i = tid; i < n; i += blockDim.x) {
            const int c = i % C, px = i / C;               // px = p1 * S + x
            strip[c * plane + px] = conv(__ldg(src + i), c);
        }
    }
    __syncthreads();
    const int64_t row0 = (int64_t)b * (P + 1) + 1 + (int64_t)ph * w;
    const int n_out = w * half_pd;
    for (int i = tid; i < n_out; i += blockDim.x) {
        const int pw = i / half_pd, e0 = (i - pw * half_pd) * 2;
        const int off = pw * patch;
        *reinterpret_cast<__half2*>(out + (row0 + pw) * ld + e0) = __floats2half2_rn(strip[lut[e0] + off], strip[lut[e0 + 1] + off]);
    }
    if (ph == 0)                // token 0 (cls slot) is a zero row
        for (int i = tid; i < half_pd; i += blockDim.x) *reinterpret_cast<__half2*>(out + (int64_t)b * (P + 1) * ld + 2 * i) = __floats2half2_rn(0.f, 0.f);
}

template <typename T>
static int patchify_launch(const T* img, __half* out, int64_t ld, int B, int C, int S, int patch, int order, int layout, const PixelNorm& nrm,
                           cudaStream_t s) {
    GSL_REQUIRE(S % patch == 0, "image size %d not divisible by patch %d", S, patch);
    GSL_REQUIRE((C * patch * patch) % 2 == 0 && ld % 2 == 0, "patchify: patch_dim and ld must be even");
    if (B == 0) return 0;
    const int w = S / patch;
    const size_t smem = (size_t)C * patch * S * 4 + (size_t)C * patch * patch * 4;
    GSL_REQUIRE(smem <= 200 * 1024, "patchify: a %d x %d x %d strip does not fit in shared memory", C, patch, S);
    static size_t smem_set[2] = {0, 0};           // per instantiation (float / uint8)
    size_t& cur = smem_set[sizeof(T) == 1 ? 1 : 0];
    if (smem > cur) {
        GSL_CHECK_CUDA(cudaFuncSetAttribute(patchify_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    GSL_CHECK_CUDA(launch_pdl(patchify_kernel<T>, dim3((unsigned)(B * w)), dim3(256), smem, s, img, out, ld, B, C, S, patch, order, layout, nrm));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int patchify_f16(const float* img, __half* out, int64_t ld, int B, int C, int S, int patch, int order, cudaStream_t s) {
    PixelNorm nrm{};
    return patchify_launch<float>(img, out, ld, B, C, S, patch, order, 0, nrm, s);
}

int patchify_u8_f16(const uint8_t* img, int layout, const float* mean, const float* std, __half* out, int64_t ld, int B, int C, int S, int patch,
                    int order, cudaStream_t s) {
    GSL_REQUIRE(layout == 0 || layout == 1, "patchify_u8: layout must be 0 (NCHW) or 1 (NHWC)");
    GSL_REQUIRE(C >= 1 && C <= 4, "patchify_u8: 1..4 channels");
    GSL_REQUIRE((mean == nullptr) == (std == nullptr), "patchify_u8: pass both mean and std (host pointers, C floats each) or neither");
    PixelNorm nrm{};
    if (mean) {
        nrm.enabled = 1;
        for (int c = 0; c < C; ++c) {
            GSL_REQUIRE(std[c] != 0.f, "patchify_u8: std[%d] is zero", c);
            nrm.mean[c] = mean[c]; nrm.std[c] = std[c];
        }
    }
    return patchify_launch<uint8_t>(img, out, ld, B, C, S, patch, order, layout, nrm, s);
}

// ------------------------------------------------------------------------------------------------ LayerNorm forward
// nn.LayerNorm(dim) of PreNorm (vit_face.py:316-323): y = (x - mean) * rstd * gamma + beta, biased variance.
// One warp per row, VEC float4 per lane (D = 128 * VEC); writes fp16 (the next GEMM's A operand) + row stats.
template <int VEC>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, __half* __restrict__ y,
                                                            int64_t ldy, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int64_t M) {
    pdl_prologue();
    constexpr int D = VEC * 128;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
    float4 v[VEC];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        v[i] = xr[lane + 32 * i];
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(sum) * (1.0f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
    uint2* yr = reinterpret_cast<uint2*>(y + row * ldy);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
        uint2 o;
        o.x = pack_half2((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
        o.y = pack_half2((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
        yr[lane + 32 * i] = o;
    }
}

int layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, __half* y, int64_t ldy,
                  float* mean, float* rstd, int64_t M, int D, cudaStream_t s) {
    GSL_REQUIRE(D % 128 == 0 && D <= 1024, "layernorm: D=%d must be a multiple of 128 and <= 1024", D);
    GSL_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "layernorm: leading dimensions must be multiples of 4");
    const int warps = 8;
    const int blocks = (int)((M + warps - 1) / warps);
#define GSL_LN_CASE(V) case V: GSL_CHECK_CUDA(launch_pdl(layernorm_fwd_kernel<V>, dim3(blocks), dim3(warps * 32), 0, s, x, ldx, gamma, beta, eps, y, ldy, mean, rstd, M)); break;
    switch (D / 128) {
        GSL_LN_CASE(1) GSL_LN_CASE(2) GSL_LN_CASE(3) GSL_LN_CASE(4) GSL_LN_CASE(5) GSL_LN_CASE(6) GSL_LN_CASE(7) GSL_LN_CASE(8)
    }
#undef GSL_LN_CASE
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// dx = dres + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma,  xhat = (x - mean) * rstd.
// (autograd's native_layer_norm_backward for the frozen-affine case: gamma/beta get no gradient because
//  lora.mark_only_lora_as_trainable froze them, train_own_forget_cl.py:316.)
// Emits the fp32 gradient stream and its fp16 copy (A operand of the next dX GEMM).
template <int VEC, bool DY16>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const void* __restrict__ dy_v, int64_t lddy, const float* __restrict__ x, int64_t ldx,
                                                            const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                            const float* __restrict__ gamma, const float* __restrict__ dres, int64_t lddres,
                                                            float* __restrict__ dx, int64_t lddx, __half* __restrict__ dx16, int64_t lddx16,
                                                            int64_t M, uint32_t drop_thresh, DropSeed drop_seed_in, float drop_scale, int dres_period) {
    pdl_prologue();
    const uint32_t drop_seed = drop_thresh ? drop_seed_resolve(drop_seed_in) : 0u;
    constexpr int D = VEC * 128;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const float mean = mean_in[row], rstd = rstd_in[row];
    // dres_period > 0: the residual gradient is non-zero only at rows that are multiples of the period (the cls tokens under the last
    // block, whose other tokens are dead) and `dres` holds those rows compacted, one per image
    const float* dres_row = nullptr;
    if (dres) {
        if (dres_period == 0) dres_row = dres + row * lddres;
        else if (row % dres_period == 0) dres_row = dres + (row / dres_period) * lddres;
    }
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
    const float4* dyr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_v) + (DY16 ? 0 : row * lddy));
    const uint2* dyh = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(dy_v) + (DY16 ? row * lddy : 0));
    float4 xh[VEC], g[VEC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float4 xv = xr[lane + 32 * i];
        float4 dv;
        if (DY16) {
            const uint2 h = dyh[lane + 32 * i];
            const float2 a = unpack_half2(h.x), b = unpack_half2(h.y);
            dv = make_float4(a.x, a.y, b.x, b.y);
        } else {
            dv = dyr[lane + 32 * i];
        }
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
        xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
        g[i] = make_float4(dv.x * gm.x, dv.y * gm.y, dv.z * gm.z, dv.w * gm.w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
    const float mg = warp_sum(s1) * (1.0f / D);
    const float mgx = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        float4 o;
        o.x = rstd * (g[i].x - mg - xh[i].x * mgx);
        o.y = rstd * (g[i].y - mg - xh[i].y * mgx);
        o.z = rstd * (g[i].z - mg - xh[i].z * mgx);
        o.w = rstd * (g[i].w - mg - xh[i].w * mgx);
        if (dres_row) {
            const float4 r = reinterpret_cast<const float4*>(dres_row)[lane + 32 * i];
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (dx) reinterpret_cast<float4*>(dx + row * lddx)[lane + 32 * i] = o;
        if (dx16) {
            if (drop_thresh) {      // the fp16 copy feeds a branch that sits behind a Dropout: apply that site's mask
                const uint32_t e0 = (uint32_t)row * (uint32_t)D + (uint32_t)(lane + 32 * i) * 4u;
                float s0, s1, s2, s3;
                drop_pair(e0, drop_seed, drop_thresh, drop_scale, s0, s1);
                drop_pair(e0 + 2, drop_seed, drop_thresh, drop_scale, s2, s3);
                o.x *= s0; o.y *= s1; o.z *= s2; o.w *= s3;
            }
            uint2 h;
            h.x = pack_half2(o.x, o.y);
            h.y = pack_half2(o.z, o.w);
            reinterpret_cast<uint2*>(dx16 + row * lddx16)[lane + 32 * i] = h;
        }
    }
}

int layernorm_bwd(const void* dy, int dy_is_fp16, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd, const float* gamma,
                  const float* dres, int64_t lddres, float* dx, int64_t lddx, __half* dx16, int64_t lddx16, int64_t M, int D,
                  float drop_p, DropSeed drop_seed, cudaStream_t s, int dres_period) {
    const uint32_t dth = drop_thresh15(drop_p);
    const float dsc = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
    GSL_REQUIRE(D % 128 == 0 && D <= 1024, "layernorm_bwd: D=%d must be a multiple of 128 and <= 1024", D);
    const int warps = 8;
    const int blocks = (int)((M + warps - 1) / warps);
#define GSL_LN_CASE(V) case V: \
        if (dy_is_fp16) GSL_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel<V, true>, dim3(blocks), dim3(warps * 32), 0, s, dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx, dx16, lddx16, M, dth, drop_seed, dsc, dres_period)); \
        else GSL_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel<V, false>, dim3(blocks), dim3(warps * 32), 0, s, dy, lddy, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx, dx16, lddx16, M, dth, drop_seed, dsc, dres_period)); \
        break;
    switch (D / 128) {
        GSL_LN_CASE(1) GSL_LN_CASE(2) GSL_LN_CASE(3) GSL_LN_CASE(4) GSL_LN_CASE(5) GSL_LN_CASE(6) GSL_LN_CASE(7) GSL_LN_CASE(8)
    }
#undef GSL_LN_CASE
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ skinny X * A^T
// T[m, j] = sum_k X[m, k] * A[j, k],  j < 16  -- the rank-r projections of loralib.Linear:
//   forward  T = x A^T  (loralib Linear.forward),   backward  U = dY B  (A := B^T).
// mma.sync m16n8k16 with a K permutation chosen so that every lane's operand loads are contiguous
// 16-byte vectors straight from global memory (no smem): within a 64-wide k chunk lane (g = lane/4,
// t = lane%4) owns k in [16t, 16t+16) of rows g and g+8; mma step s consumes its halves [4s, 4s+4).
// One warp = 16 rows; HBM-bound (reads X once).
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// FOLD: A16 holds 32 rows -- rows [0, 16) = fp16(A) ("hi"), rows [16, 32) = fp16(A - hi) ("lo") -- and both feed the same accumulator,
// so the LoRA factor enters with ~22 significand bits (precision mode "split": its rounding is systematic, like the frozen weights').
template <int NT, bool FOLD>   // NT n-tiles of 8 (r <= 8 -> 1, r <= 16 -> 2)
__global__ void __launch_bounds__(128) lora_down_kernel(const __half* __restrict__ X, int64_t ldx, const __half* __restrict__ A, int64_t lda,
                                                        __half* __restrict__ out, int64_t ldo, int64_t M, int K) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t row_base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 16;
    if (row_base >= M) return;
    const int64_t r0 = row_base + g, r1 = row_base + g + 8;
    const bool ok0 = r0 < M, ok1 = r1 < M;
    const __half* x0 = X + (ok0 ? r0 : row_base) * ldx + 16 * t;
    const __half* x1 = X + (ok1 ? r1 : row_base) * ldx + 16 * t;
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    const int nchunks = K / 64;
#pragma unroll 2
    for (int c = 0; c < nchunks; ++c) {
        const uint4 p0 = *reinterpret_cast<const uint4*>(x0 + c * 64);
        const uint4 p1 = *reinterpret_cast<const uint4*>(x0 + c * 64 + 8);
        const uint4 q0 = *reinterpret_cast<const uint4*>(x1 + c * 64);
        const uint4 q1 = *reinterpret_cast<const uint4*>(x1 + c * 64 + 8);
        const uint32_t xa[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        const uint32_t xb[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const __half* ar = A + (int64_t)(8 * n + g) * lda + c * 64 + 16 * t;
            const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(ar));
            const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(ar + 8));
            const uint32_t wb[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int s = 0; s < 4; ++s) mma_16816(acc[n], xa[2 * s], xb[2 * s], xa[2 * s + 1], xb[2 * s + 1], wb[2 * s], wb[2 * s + 1]);
            if (FOLD) {
                const uint4 l0 = __ldg(reinterpret_cast<const uint4*>(ar + 16 * lda));
                const uint4 l1 = __ldg(reinterpret_cast<const uint4*>(ar + 16 * lda + 8));
                const uint32_t lb[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
                for (int s = 0; s < 4; ++s) mma_16816(acc[n], xa[2 * s], xb[2 * s], xa[2 * s + 1], xb[2 * s + 1], lb[2 * s], lb[2 * s + 1]);
            }
        }
    }
    // K tail (K % 64 in {16, 32, 48}): same scheme on 16-wide pieces, lanes t >= pieces contribute zeros
    const int tail = K - nchunks * 64;
    if (tail > 0) {
        const int pieces = tail / 16;
        uint32_t xa[8] = {0, 0, 0, 0, 0, 0, 0, 0}, xb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (t < pieces) {
            const uint4 p0 = *reinterpret_cast<const uint4*>(x0 + nchunks * 64);
            const uint4 p1 = *reinterpret_cast<const uint4*>(x0 + nchunks * 64 + 8);
            const uint4 q0 = *reinterpret_cast<const uint4*>(x1 + nchunks * 64);
            const uint4 q1 = *reinterpret_cast<const uint4*>(x1 + nchunks * 64 + 8);
            xa[0] = p0.x; xa[1] = p0.y; xa[2] = p0.z; xa[3] = p0.w; xa[4] = p1.x; xa[5] = p1.y; xa[6] = p1.z; xa[7] = p1.w;
            xb[0] = q0.x; xb[1] = q0.y; xb[2] = q0.z; xb[3] = q0.w; xb[4] = q1.x; xb[5] = q1.y; xb[6] = q1.z; xb[7] = q1.w;
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
#pragma unroll
            for (int part = 0; part < (FOLD ? 2 : 1); ++part) {
                uint32_t wb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (t < pieces) {
                    const __half* ar = A + (int64_t)(16 * part + 8 * n + g) * lda + nchunks * 64 + 16 * t;
                    const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(ar));
                    const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(ar + 8));
                    wb[0] = w0.x; wb[1] = w0.y; wb[2] = w0.z; wb[3] = w0.w; wb[4] = w1.x; wb[5] = w1.y; wb[6] = w1.z; wb[7] = w1.w;
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) mma_16816(acc[n], xa[2 * s], xb[2 * s], xa[2 * s + 1], xb[2 * s + 1], wb[2 * s], wb[2 * s + 1]);
            }
        }
    }
    // c0,c1 -> (row g, cols 2t, 2t+1) ; c2,c3 -> (row g+8, ...) ; columns [8*NT, 16) are zero padding
#pragma unroll
    for (int n = 0; n < 2; ++n) {
        const uint32_t v0 = n < NT ? pack_half2(acc[n < NT ? n : 0][0], acc[n < NT ? n : 0][1]) : 0u;
        const uint32_t v1 = n < NT ? pack_half2(acc[n < NT ? n : 0][2], acc[n < NT ? n : 0][3]) : 0u;
        if (ok0) *reinterpret_cast<uint32_t*>(out + r0 * ldo + 8 * n + 2 * t) = v0;
        if (ok1) *reinterpret_cast<uint32_t*>(out + r1 * ldo + 8 * n + 2 * t) = v1;
    }
}

int lora_down(const __half* X, int64_t ldx, const __half* A16, int64_t lda, __half* out, int64_t ldo, int64_t M, int K, int r,
              cudaStream_t s, int fold) {
    GSL_REQUIRE(K % 16 == 0 && ldx % 8 == 0 && lda % 8 == 0 && ldo % 2 == 0, "lora_down: K %% 16, ldx %% 8, lda %% 8 required (K=%d)", K);
    GSL_REQUIRE(r >= 1 && r <= 16, "lora_down: rank must be in [1, 16] (got %d)", r);
    const int warps = 4;
    const int64_t groups = (M + 15) / 16;
    const int blocks = (int)((groups + warps - 1) / warps);
    if (r <= 8) {
        if (fold) GSL_CHECK_CUDA(launch_pdl(lora_down_kernel<1, true>, dim3(blocks), dim3(warps * 32), 0, s, X, ldx, A16, lda, out, ldo, M, K));
        else GSL_CHECK_CUDA(launch_pdl(lora_down_kernel<1, false>, dim3(blocks), dim3(warps * 32), 0, s, X, ldx, A16, lda, out, ldo, M, K));
    } else {
        if (fold) GSL_CHECK_CUDA(launch_pdl(lora_down_kernel<2, true>, dim3(blocks), dim3(warps * 32), 0, s, X, ldx, A16, lda, out, ldo, M, K));
        else GSL_CHECK_CUDA(launch_pdl(lora_down_kernel<2, false>, dim3(blocks), dim3(warps * 32), 0, s, X, ldx, A16, lda, out, ldo, M, K));
    }
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ skinny L^T * R over tokens
// P[n, j] = scale * sum_m L[m, n] * R[m, j]   (n < N, j < r):  dB = s * dY^T T,  dA^T = s * X^T U  (SURVEY Appendix C).
// HBM-bound (reads L once).  Split-M: CTA (col-block of 256, m-range) streams [64 m x 256 n] tiles of L and [64 m x 16] tiles
// of R through a 3-stage cp.async ring; each warp owns 32 columns and contracts over m with mma.sync m16n8k16 -- both
// operands are "transposed" views of row-major tiles, fetched with ldmatrix.trans (no explicit transpose anywhere).
// Partials go to a workspace and a second tiny kernel reduces them in a fixed order (deterministic, no atomics).
static constexpr int SK_COLS = 256, SK_ROWS = 64, SK_STAGES = 3, SK_THREADS = 256;
static constexpr int SK_L_BYTES = SK_ROWS * SK_COLS * 2;   // 32 KB
static constexpr int SK_R_BYTES = SK_ROWS * 16 * 2;        //  2 KB
static constexpr int SK_STAGE_BYTES = SK_L_BYTES + SK_R_BYTES;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

template <int R>
__global__ void __launch_bounds__(SK_THREADS, 2) skinny_tn_partial_kernel(const __half* __restrict__ L, int64_t ldl, const __half* __restrict__ Rm, int64_t ldr,
                                                                         float* __restrict__ partial, int64_t M, int N, int rows_per_split) {
    pdl_prologue();
    extern __shared__ __align__(128) uint8_t sk_smem[];
    const uint32_t sbase = smem_u32(sk_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int col0 = blockIdx.x * SK_COLS;
    const int split = blockIdx.y;
    const int64_t m0 = (int64_t)split * rows_per_split;
    const int64_t m1 = (m0 + rows_per_split < M) ? m0 + rows_per_split : M;
    const int nchunks = m1 > m0 ? (int)((m1 - m0 + SK_ROWS - 1) / SK_ROWS) : 0;

    auto load_chunk = [&](int c, int stage) {
        const int64_t mb = m0 + (int64_t)c * SK_ROWS;
        const uint32_t sL = sbase + stage * SK_STAGE_BYTES, sR = sL + SK_L_BYTES;
#pragma unroll
        for (int i = 0; i < (SK_ROWS * SK_COLS / 8) / SK_THREADS; ++i) {
            const int idx = tid + i * SK_THREADS;
            const int row = idx >> 5, chunk = idx & 31;
            const bool ok = (mb + row < m1) && (col0 + chunk * 8 < N);
            const __half* src = L + (ok ? (mb + row) * ldl + col0 + chunk * 8 : 0);
            cp_async16(sL + row * 512 + (((chunk & ~7) | ((chunk ^ row) & 7)) << 4), src, ok);
        }
        if (tid < SK_ROWS * 2) {
            const int row = tid >> 1, chunk = tid & 1;
            const bool ok = (mb + row < m1);
            const __half* src = Rm + (ok ? (mb + row) * ldr + chunk * 8 : 0);
            cp_async16(sR + row * 32 + chunk * 16, src, ok);
        }
    };

    constexpr int NJ = R / 8;
    float acc[2][NJ][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[a][j][0] = acc[a][j][1] = acc[a][j][2] = acc[a][j][3] = 0.f;

    for (int c = 0; c < SK_STAGES - 1; ++c) {
        if (c < nchunks) load_chunk(c, c);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int c = 0; c < nchunks; ++c) {
        asm volatile("cp.async.wait_group %0;" ::"n"(SK_STAGES - 2) : "memory");
        __syncthreads();
        if (c + SK_STAGES - 1 < nchunks) load_chunk(c + SK_STAGES - 1, (c + SK_STAGES - 1) % SK_STAGES);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const uint32_t sL = sbase + (c % SK_STAGES) * SK_STAGE_BYTES, sR = sL + SK_L_BYTES;
#pragma unroll
        for (int ks = 0; ks < SK_ROWS / 16; ++ks) {
            // B fragments: R tile rows m (k), columns j (n), transposed load
            uint32_t bf[NJ][2];
            {
                const int row = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                if (NJ == 1) {
                    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(bf[0][0]), "=r"(bf[0][1]) : "r"(sR + row * 32));
                } else {
                    uint32_t r4[4];
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(r4[0]), "=r"(r4[1]), "=r"(r4[2]), "=r"(r4[3]) : "r"(sR + row * 32 + (lane >> 4) * 16));
                    bf[0][0] = r4[0]; bf[0][1] = r4[1]; bf[NJ - 1][0] = r4[2]; bf[NJ - 1][1] = r4[3];
                }
            }
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                // A fragment (rows = n, k = m) from the [m][n] tile: transposed load
                const int row = ks * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
                const int chunk = warp * 4 + a * 2 + ((lane >> 3) & 1);
                uint32_t af[4];
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(af[0]), "=r"(af[1]), "=r"(af[2]), "=r"(af[3])
                             : "r"(sL + row * 512 + (((chunk & ~7) | ((chunk ^ row) & 7)) << 4)));
#pragma unroll
                for (int j = 0; j < NJ; ++j) mma_16816(acc[a][j], af[0], af[1], af[2], af[3], bf[j][0], bf[j][1]);
            }
        }
    }
    // c0,c1 -> (n = g, j = 2t, 2t+1) ; c2,c3 -> (n = g + 8)
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int n = col0 + warp * 32 + a * 16 + g;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (n < N) *reinterpret_cast<float2*>(partial + ((int64_t)split * N + n) * R + j * 8 + 2 * t) = make_float2(acc[a][j][0], acc[a][j][1]);
            if (n + 8 < N) *reinterpret_cast<float2*>(partial + ((int64_t)split * N + n + 8) * R + j * 8 + 2 * t) = make_float2(acc[a][j][2], acc[a][j][3]);
        }
    }
}

// Reduction of the split partials in a FIXED order (deterministic, no atomics).  The partials were just written and sit in L2, so the
// kernel is pure latency: RED_PARTS threads share one output (part p sums splits p, p + RED_PARTS, ...) with RED_DEPTH independent loads
// in flight each, and the parts are combined through shared memory in part order.
static constexpr int RED_PARTS = 4, RED_OUTS = 64, RED_DEPTH = 8;
template <int R>
__global__ void __launch_bounds__(RED_PARTS * RED_OUTS) skinny_tn_reduce_kernel(const float* __restrict__ partial, int splits, int N, float scale,
                                                                                float* __restrict__ out, int64_t ldo, int transpose_out, int r_out,
                                                                                int accumulate) {
    pdl_prologue();
    __shared__ float s_part[RED_PARTS][RED_OUTS];
    const int lo = threadIdx.x % RED_OUTS, part = threadIdx.x / RED_OUTS;
    const int i = blockIdx.x * RED_OUTS + lo;
    const int64_t stride = (int64_t)N * R;
    float s = 0.f;
    if (i < N * R) {
        int k = part;
        for (; k + (RED_DEPTH - 1) * RED_PARTS < splits; k += RED_DEPTH * RED_PARTS) {
            float v[RED_DEPTH];
#pragma unroll
            for (int u = 0; u < RED_DEPTH; ++u) v[u] = partial[(int64_t)(k + u * RED_PARTS) * stride + i];
#pragma unroll
            for (int u = 0; u < RED_DEPTH; ++u) s += v[u];
        }
        for (; k < splits; k += RED_PARTS) s += partial[(int64_t)k * stride + i];
    }
    s_part[part][lo] = s;
    __syncthreads();
    if (part == 0 && i < N * R) {
        float t = s_part[0][lo];
#pragma unroll
        for (int p = 1; p < RED_PARTS; ++p) t += s_part[p][lo];
        const int n = i / R, j = i % R;
        if (j < r_out) {
            float* o = transpose_out ? out + (int64_t)j * ldo + n : out + (int64_t)n * ldo + j;
            *o = (accumulate ? *o : 0.f) + t * scale;
        }
    }
}

static int skinny_splits(int64_t M, int N) {
    const int colblocks = (N + SK_COLS - 1) / SK_COLS;
    int splits = (device_sm_count() * 2 + colblocks - 1) / colblocks;
    const int64_t max_splits = (M + SK_ROWS - 1) / SK_ROWS;
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    return splits;
}

size_t skinny_tn_workspace(int64_t M, int N, int r) {
    const int R = r <= 8 ? 8 : 16;
    return (size_t)skinny_splits(M, N) * N * R * sizeof(float);
}

int skinny_tn(const __half* L, int64_t ldl, const __half* Rm, int64_t ldr, float* out, int64_t ldo, int transpose_out, float scale,
              int accumulate, int64_t M, int N, int r, float* workspace, size_t workspace_bytes, cudaStream_t s) {
    GSL_REQUIRE(r >= 1 && r <= 16, "skinny_tn: rank must be in [1, 16] (got %d)", r);
    GSL_REQUIRE(N % 8 == 0 && ldl % 8 == 0 && ldr % 8 == 0, "skinny_tn: N, ldl, ldr must be multiples of 8");
    GSL_REQUIRE(workspace_bytes >= skinny_tn_workspace(M, N, r), "skinny_tn: workspace too small");
    const int splits = skinny_splits(M, N);
    int rows_per_split = (int)((M + splits - 1) / splits);
    rows_per_split = (rows_per_split + 63) / 64 * 64;
    dim3 grid((N + SK_COLS - 1) / SK_COLS, splits);
    const int smem = SK_STAGES * SK_STAGE_BYTES;
    static bool attr = false;
    if (!attr) {
        GSL_CHECK_CUDA(cudaFuncSetAttribute(skinny_tn_partial_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GSL_CHECK_CUDA(cudaFuncSetAttribute(skinny_tn_partial_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    if (r <= 8) {
        GSL_CHECK_CUDA(launch_pdl(skinny_tn_partial_kernel<8>, dim3(grid), dim3(SK_THREADS), smem, s, L, ldl, Rm, ldr, workspace, M, N, rows_per_split));
        GSL_COUNT_LAUNCH(1);
        GSL_CHECK_CUDA(launch_pdl(skinny_tn_reduce_kernel<8>, dim3((N * 8 + RED_OUTS - 1) / RED_OUTS), dim3(RED_PARTS * RED_OUTS), 0, s, workspace, splits, N, scale, out, ldo, transpose_out, r, accumulate));
        GSL_COUNT_LAUNCH(1);
    } else {
        GSL_CHECK_CUDA(launch_pdl(skinny_tn_partial_kernel<16>, dim3(grid), dim3(SK_THREADS), smem, s, L, ldl, Rm, ldr, workspace, M, N, rows_per_split));
        GSL_COUNT_LAUNCH(1);
        GSL_CHECK_CUDA(launch_pdl(skinny_tn_reduce_kernel<16>, dim3((N * 16 + RED_OUTS - 1) / RED_OUTS), dim3(RED_PARTS * RED_OUTS), 0, s, workspace, splits, N, scale, out, ldo, transpose_out, r, accumulate));
        GSL_COUNT_LAUNCH(1);
    }
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ fused LoRA side pass
// One read of a wide activation L [M, N] (G = Dropout(gelu(h)) or dH, N = mlp_dim) yields BOTH LoRA by-products that need it
// (SURVEY Appendix C):
//   T[m, j] = sum_n L[m, n] * P[j, n]             down projection   T2 = G A2^T,   U1 = dH B1          -> fp16 [M, 16]
//   Q[n, j] = scale * sum_m L[m, n] * R[m, j]      token contraction dA2^T = s G^T U2,  dB1 = s dH^T T1 -> fp32
// so the rank-r intermediates never cost an extra pass over L.  CTA = one m-range and ALL N columns: 16 warps x (N / 16)
// columns; 16-row chunks of L stream through a cp.async ring; every warp reads each 16 x 16 block of its columns twice with
// ldmatrix (plain = A operand of the down projection, .trans = A operand of the token contraction) and feeds mma.sync
// m16n8k16; P stays in registers as B fragments; the per-warp partial rows of T are summed across the 16 warps through
// shared memory; the Q partials of every CTA go to the workspace and skinny_tn_reduce_kernel adds them in a fixed order.
// HBM-bound: algorithmic bytes = 2 M N.
static constexpr int SP_WARPS = 16, SP_THREADS = SP_WARPS * 32, SP_ROWS = 16;
static constexpr int SP_RED_BYTES = 2 * SP_WARPS * SP_ROWS * 8 * 4;

template <int NB, int STAGES, bool FOLD>   // NB = 16-column blocks per warp (N = 256 * NB); FOLD: P has 32 rows, hi [0, 16) + lo [16, 32) (see lora_down_kernel)
__global__ void __launch_bounds__(SP_THREADS, 1) lora_side_kernel(const __half* __restrict__ L, int64_t ldl, const __half* __restrict__ P, int64_t ldp,
                                                                 __half* __restrict__ T, int64_t ldt, const __half* __restrict__ Rm, int64_t ldr,
                                                                 float* __restrict__ partial, int64_t M, int rows_per_cta) {
    pdl_prologue();
    constexpr int N = 256 * NB, ROW_BYTES = 2 * N, CHUNKS = N / 8;          // 16-byte chunks per row
    constexpr int L_BYTES = SP_ROWS * ROW_BYTES, STAGE_BYTES = L_BYTES + SP_ROWS * 32;
    extern __shared__ __align__(128) uint8_t sp_smem[];
    const uint32_t sbase = smem_u32(sp_smem);
    float* red = reinterpret_cast<float*>(sp_smem + STAGES * STAGE_BYTES);  // [2][16 warps][16 rows x 8]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int64_t m0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t m1 = (m0 + rows_per_cta < M) ? m0 + rows_per_cta : M;
    const int nchunks = m1 > m0 ? (int)((m1 - m0 + SP_ROWS - 1) / SP_ROWS) : 0;

    auto load_chunk = [&](int c, int stage) {
        const int64_t mb = m0 + (int64_t)c * SP_ROWS;
        const uint32_t sL = sbase + stage * STAGE_BYTES, sR = sL + L_BYTES;
#pragma unroll
        for (int i = 0; i < (SP_ROWS * CHUNKS) / SP_THREADS; ++i) {
            const int idx = tid + i * SP_THREADS;
            const int row = idx / CHUNKS, chunk = idx % CHUNKS;
            const bool ok = mb + row < m1;
            const __half* src = L + (ok ? (mb + row) * ldl + chunk * 8 : 0);
            cp_async16(sL + row * ROW_BYTES + (((chunk & ~7) | ((chunk ^ row) & 7)) << 4), src, ok);
        }
        if (tid < SP_ROWS * 2) {
            const int row = tid >> 1, chunk = tid & 1;
            const bool ok = mb + row < m1;
            const __half* src = Rm + (ok ? (mb + row) * ldr + chunk * 8 : 0);
            cp_async16(sR + row * 32 + chunk * 16, src, ok);
        }
    };

    // B fragments of the down projection (k = n, 8 output columns j): lane (g, t) holds P[g][c + 2t, 2t+1] and P[g][c + 8 + 2t, ...]
    uint32_t pf[NB][2], pl[FOLD ? NB : 1][2];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        const __half* pr = P + (int64_t)g * ldp + (warp * NB + nb) * 16 + 2 * t;
        pf[nb][0] = __ldg(reinterpret_cast<const uint32_t*>(pr));
        pf[nb][1] = __ldg(reinterpret_cast<const uint32_t*>(pr + 8));
        if (FOLD) {
            pl[nb][0] = __ldg(reinterpret_cast<const uint32_t*>(pr + 16 * ldp));
            pl[nb][1] = __ldg(reinterpret_cast<const uint32_t*>(pr + 16 * ldp + 8));
        }
    }
    float accQ[NB][4];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) accQ[nb][0] = accQ[nb][1] = accQ[nb][2] = accQ[nb][3] = 0.f;

    for (int c = 0; c < STAGES - 1; ++c) {
        if (c < nchunks) load_chunk(c, c);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int c = 0; c < nchunks; ++c) {
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
        __syncthreads();
        if (c + STAGES - 1 < nchunks) load_chunk(c + STAGES - 1, (c + STAGES - 1) % STAGES);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const uint32_t sL = sbase + (c % STAGES) * STAGE_BYTES, sR = sL + L_BYTES;
        uint32_t bf[2];     // token contraction B fragment: R tile rows m (k), columns j (n), transposed load
        {
            const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
            asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(bf[0]), "=r"(bf[1]) : "r"(sR + row * 32));
        }
        float accT[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            const int cb = (warp * NB + nb) * 2;        // first 16-byte chunk of this 16-column block
            uint32_t af[4];
            {   // rows = n, k = m: transposed 8x8 blocks of the [m][n] tile
                const int row = (lane & 7) + ((lane >> 4) & 1) * 8;
                const int chunk = cb + ((lane >> 3) & 1);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(af[0]), "=r"(af[1]), "=r"(af[2]), "=r"(af[3])
                             : "r"(sL + row * ROW_BYTES + (((chunk & ~7) | ((chunk ^ row) & 7)) << 4)));
            }
            mma_16816(accQ[nb], af[0], af[1], af[2], af[3], bf[0], bf[1]);
            {   // rows = m, k = n: plain 8x8 blocks
                const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
                const int chunk = cb + (lane >> 4);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(af[0]), "=r"(af[1]), "=r"(af[2]), "=r"(af[3])
                             : "r"(sL + row * ROW_BYTES + (((chunk & ~7) | ((chunk ^ row) & 7)) << 4)));
            }
            mma_16816(accT, af[0], af[1], af[2], af[3], pf[nb][0], pf[nb][1]);
            if (FOLD) mma_16816(accT, af[0], af[1], af[2], af[3], pl[nb][0], pl[nb][1]);
        }
        // per-warp partial T rows -> shared, summed by the first 128 threads: c0,c1 -> (row g, j = 2t, 2t+1); c2,c3 -> (row g + 8)
        float* rbuf = red + (c & 1) * (SP_WARPS * SP_ROWS * 8);
        *reinterpret_cast<float2*>(rbuf + warp * (SP_ROWS * 8) + g * 8 + 2 * t) = make_float2(accT[0], accT[1]);
        *reinterpret_cast<float2*>(rbuf + warp * (SP_ROWS * 8) + (g + 8) * 8 + 2 * t) = make_float2(accT[2], accT[3]);
        __syncthreads();
        if (tid < 2 * SP_ROWS * 8) {
            const int e = tid & (SP_ROWS * 8 - 1), row = e >> 3, j = e & 7;
            const int64_t m = m0 + (int64_t)c * SP_ROWS + row;
            if (m < m1) {
                float sum = 0.f;
                if (tid < SP_ROWS * 8) {
#pragma unroll
                    for (int w = 0; w < SP_WARPS; ++w) sum += rbuf[w * (SP_ROWS * 8) + e];
                }
                T[m * ldt + (tid < SP_ROWS * 8 ? j : 8 + j)] = __float2half_rn(sum);      // columns 8..15 are zero padding
            }
        }
    }
    // Q partials: c0,c1 -> (n = g, j = 2t, 2t+1) ; c2,c3 -> (n = g + 8)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        const int n = (warp * NB + nb) * 16 + g;
        *reinterpret_cast<float2*>(partial + ((int64_t)blockIdx.x * N + n) * 8 + 2 * t) = make_float2(accQ[nb][0], accQ[nb][1]);
        *reinterpret_cast<float2*>(partial + ((int64_t)blockIdx.x * N + n + 8) * 8 + 2 * t) = make_float2(accQ[nb][2], accQ[nb][3]);
    }
}

static bool lora_side_fused_ok(int N, int r) { return r <= 8 && N % 256 == 0 && N / 256 >= 1 && N / 256 <= 12; }
static int lora_side_ctas(int64_t M) {
    const int64_t chunks = (M + SP_ROWS - 1) / SP_ROWS;
    const int sms = device_sm_count();
    return (int)(chunks < sms ? chunks : sms);
}
size_t lora_side_workspace(int64_t M, int N, int r) {
    const size_t fallback = skinny_tn_workspace(M, N, r);
    if (!lora_side_fused_ok(N, r)) return fallback;
    const size_t fused = (size_t)lora_side_ctas(M) * N * 8 * sizeof(float);
    return fused > fallback ? fused : fallback;
}

template <int NB, int STAGES, bool FOLD>
static int launch_lora_side(const __half* L, int64_t ldl, const __half* P16, int64_t ldp, __half* T, int64_t ldt, const __half* Rm, int64_t ldr,
                            float* workspace, int64_t M, int ctas, int rows_per_cta, cudaStream_t s) {
    constexpr int smem = STAGES * (SP_ROWS * 512 * NB + SP_ROWS * 32) + SP_RED_BYTES;
    static_assert(smem <= 232448, "lora_side: stage ring does not fit in shared memory");
    static bool attr = false;
    if (!attr) {
        GSL_CHECK_CUDA(cudaFuncSetAttribute(lora_side_kernel<NB, STAGES, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    GSL_CHECK_CUDA(launch_pdl(lora_side_kernel<NB, STAGES, FOLD>, dim3(ctas), dim3(SP_THREADS), smem, s, L, ldl, P16, ldp, T, ldt, Rm, ldr, workspace, M, rows_per_cta));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int lora_side(const __half* L, int64_t ldl, const __half* P16, int64_t ldp, __half* T, int64_t ldt, const __half* Rm, int64_t ldr,
              float* out, int64_t ldo, int transpose_out, float scale, int accumulate, int64_t M, int N, int r,
              float* workspace, size_t workspace_bytes, cudaStream_t s, int fold) {
    GSL_REQUIRE(r >= 1 && r <= 16, "lora_side: rank must be in [1, 16] (got %d)", r);
    GSL_REQUIRE(N % 16 == 0 && ldl % 8 == 0 && ldp % 8 == 0 && ldr % 8 == 0 && ldt % 8 == 0, "lora_side: N %% 16 and pitches %% 8 required");
    GSL_REQUIRE(workspace_bytes >= lora_side_workspace(M, N, r), "lora_side: workspace too small");
    if (!lora_side_fused_ok(N, r)) {        // rank 16 / odd widths: the two separate passes
        int rc = lora_down(L, ldl, P16, ldp, T, ldt, M, N, r, s, fold);
        if (rc) return rc;
        return skinny_tn(L, ldl, Rm, ldr, out, ldo, transpose_out, scale, accumulate, M, N, r, workspace, workspace_bytes, s);
    }
    const int ctas = lora_side_ctas(M);
    int rows_per_cta = (int)((M + ctas - 1) / ctas);
    rows_per_cta = (rows_per_cta + SP_ROWS - 1) / SP_ROWS * SP_ROWS;
    int rc = -1;
#define GSL_SP_CASE(NBV, ST) case NBV: \
        rc = fold ? launch_lora_side<NBV, ST, true>(L, ldl, P16, ldp, T, ldt, Rm, ldr, workspace, M, ctas, rows_per_cta, s) \
                  : launch_lora_side<NBV, ST, false>(L, ldl, P16, ldp, T, ldt, Rm, ldr, workspace, M, ctas, rows_per_cta, s); break;
    switch (N / 256) {
        GSL_SP_CASE(1, 4) GSL_SP_CASE(2, 4) GSL_SP_CASE(3, 4) GSL_SP_CASE(4, 4) GSL_SP_CASE(5, 4) GSL_SP_CASE(6, 4)
        GSL_SP_CASE(7, 3) GSL_SP_CASE(8, 3) GSL_SP_CASE(9, 2) GSL_SP_CASE(10, 2) GSL_SP_CASE(11, 2) GSL_SP_CASE(12, 2)
    }
#undef GSL_SP_CASE
    if (rc) return rc;
    GSL_CHECK_CUDA(launch_pdl(skinny_tn_reduce_kernel<8>, dim3((N * 8 + RED_OUTS - 1) / RED_OUTS), dim3(RED_PARTS * RED_OUTS), 0, s, workspace, ctas, N, scale, out, ldo, transpose_out, r, accumulate));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ casts
// dst = fp16(v), v = src * scale; dst_lo (optional, same layout) = fp16(v - dst): the second term of a split operand (hi + lo = v to ~2^-22);
// dst_lo8 (optional, same layout, one byte per element) = e4m3(v - dst): the FP8 residual of precision mode "split8" (scale = 2^shift there)
__global__ void cast_kernel(const float* __restrict__ src, int64_t lds, __half* __restrict__ dst, __half* __restrict__ dst_lo, uint8_t* __restrict__ dst_lo8,
                            int64_t ldd, int64_t rows, int64_t cols, float scale, int transpose) {
    const int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i % cols;
        const float v = src[r * lds + c] * scale;
        const __half h = __float2half_rn(v);
        const int64_t o = transpose ? c * ldd + r : r * ldd + c;
        dst[o] = h;
        if (dst_lo) dst_lo[o] = __float2half_rn(v - __half2float(h));
        if (dst_lo8) dst_lo8[o] = (uint8_t)__nv_cvt_float_to_fp8(v - __half2float(h), __NV_SATFINITE, __NV_E4M3);
    }
}

int cast_f32_to_f16(const float* src, int64_t lds, __half* dst, int64_t ldd, int64_t rows, int64_t cols, float scale, int transpose,
                    cudaStream_t s, __half* dst_lo, uint8_t* dst_lo8) {
    const int64_t total = rows * cols;
    if (total == 0) return 0;
    int blocks = (int)((total + 255) / 256);
    const int cap = device_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cast_kernel<<<blocks, 256, 0, s>>>(src, lds, dst, dst_lo, dst_lo8, ldd, rows, cols, scale, transpose);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int fill_zero(void* ptr, size_t bytes, cudaStream_t s) {
    GSL_CHECK_CUDA(cudaMemsetAsync(ptr, 0, bytes, s));
    return 0;
}

}  // namespace gsl
