// gslora-b200: last-block shortcut.  ViT_face pools the cls token only (vit_face.py:540), so in the LAST Transformer block
// every non-cls token is dead: its attention output, FFN output and all gradients through them never reach the loss.
// The engine therefore runs the last block's attention for the single cls query per (image, head) and the rest of that
// block on the B compacted cls rows.  These are the two kernels for that single-query attention (forward / backward):
// one warp per (image, head), K and V rows streamed from HBM exactly once (forward) / twice (backward), fp32 math.
// Results are identical to the dense computation restricted to what reaches the loss.
#include "gsl_common.cuh"
#include "gsl_kernels.h"

namespace gsl {

static constexpr int CLS_MAX_KEYS = 256;     // keys per lane: 8

__device__ __forceinline__ float dot64_h(const __half* __restrict__ row, const float (&q)[64]) {
    float acc = 0.f;
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 u = __ldg(r4 + c);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = unpack_half2(w[k]);
            acc = fmaf(f.x, q[c * 8 + 2 * k], acc);
            acc = fmaf(f.y, q[c * 8 + 2 * k + 1], acc);
        }
    }
    return acc;
}

// qkv fp16 [B*N, ld] (q | k | v column blocks of heads*64).  o_cls fp16 [B, ldo] (heads*64 columns), lse_cls [B, heads].
__global__ void __launch_bounds__(128) cls_attention_fwd_kernel(const __half* __restrict__ qkv, int64_t ld, __half* __restrict__ o_cls, int64_t ldo,
                                                                float* __restrict__ lse_cls, int B, int N, int heads, float scale) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= B * heads) return;
    const int b = w / heads, h = w % heads, D = heads * 64;
    const __half* base = qkv + (int64_t)b * N * ld + h * 64;
    float q[64];
    {
        const uint4* q4 = reinterpret_cast<const uint4*>(base);      // token 0 = cls
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 u = __ldg(q4 + c);
            const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 f = unpack_half2(ww[k]); q[c * 8 + 2 * k] = f.x; q[c * 8 + 2 * k + 1] = f.y; }
        }
    }
    float s[CLS_MAX_KEYS / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < CLS_MAX_KEYS / 32; ++i) {
        const int j = lane + 32 * i;
        s[i] = j < N ? scale * dot64_h(base + (int64_t)j * ld + D, q) : -INFINITY;
        mx = fmaxf(mx, s[i]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < CLS_MAX_KEYS / 32; ++i) { s[i] = (lane + 32 * i) < N ? __expf(s[i] - mx) : 0.f; sum += s[i]; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    // o[d] = sum_j p_j V[j, d]; lane owns d = 2*lane, 2*lane+1; p_j broadcast by shuffle, V rows read coalesced
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < N; ++j) {
        float p = 0.f;
#pragma unroll
        for (int i = 0; i < CLS_MAX_KEYS / 32; ++i) if ((j >> 5) == i) p = __shfl_sync(0xffffffffu, s[i], j & 31);
        const float2 v = unpack_half2(__ldg(reinterpret_cast<const uint32_t*>(base + (int64_t)j * ld + 2 * D) + lane));
        o0 = fmaf(p, v.x, o0);
        o1 = fmaf(p, v.y, o1);
    }
    *reinterpret_cast<uint32_t*>(o_cls + (int64_t)b * ldo + h * 64 + 2 * lane) = pack_half2(o0 * inv, o1 * inv);
    if (lane == 0) lse_cls[(int64_t)b * heads + h] = mx + __logf(sum);
}

int cls_attention_fwd(const __half* qkv, int64_t ld, __half* o_cls, int64_t ldo, float* lse_cls, int B, int N, int heads, float scale, cudaStream_t s) {
    GSL_REQUIRE(N <= CLS_MAX_KEYS && ld % 8 == 0 && ldo % 2 == 0, "cls_attention: tokens=%d > %d or unaligned pitch", N, CLS_MAX_KEYS);
    const int warps = 4;
    cls_attention_fwd_kernel<<<(B * heads + warps - 1) / warps, warps * 32, 0, s>>>(qkv, ld, o_cls, ldo, lse_cls, B, N, heads, scale);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Backward for the single cls query.  do_cls fp16 [B, lddo]; writes dqkv fp16 [B*N, lddqkv]: dQ row of token 0, dK / dV rows of every token.
// (dQ rows of tokens > 0 are zero: the caller clears that column block.)
__global__ void __launch_bounds__(128) cls_attention_bwd_kernel(const __half* __restrict__ qkv, int64_t ld, const __half* __restrict__ o_cls, int64_t ldo,
                                                                const __half* __restrict__ do_cls, int64_t lddo, const float* __restrict__ lse_cls,
                                                                __half* __restrict__ dqkv, int64_t lddqkv, int B, int N, int heads, float scale) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= B * heads) return;
    const int b = w / heads, h = w % heads, D = heads * 64;
    const __half* base = qkv + (int64_t)b * N * ld + h * 64;
    __half* dbase = dqkv + (int64_t)b * N * lddqkv + h * 64;
    float q[64], g[64];
    {
        const uint4* q4 = reinterpret_cast<const uint4*>(base);
        const uint4* g4 = reinterpret_cast<const uint4*>(do_cls + (int64_t)b * lddo + h * 64);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 u = __ldg(q4 + c), v = __ldg(g4 + c);
            const uint32_t uw[4] = {u.x, u.y, u.z, u.w}, vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 a = unpack_half2(uw[k]), c2 = unpack_half2(vw[k]);
                q[c * 8 + 2 * k] = a.x; q[c * 8 + 2 * k + 1] = a.y;
                g[c * 8 + 2 * k] = c2.x; g[c * 8 + 2 * k + 1] = c2.y;
            }
        }
    }
    // delta = dO . O  (lane owns 2 elements)
    const float2 ov = unpack_half2(*reinterpret_cast<const uint32_t*>(o_cls + (int64_t)b * ldo + h * 64 + 2 * lane));
    const float2 gv = unpack_half2(*reinterpret_cast<const uint32_t*>(do_cls + (int64_t)b * lddo + h * 64 + 2 * lane));
    const float delta = warp_sum(ov.x * gv.x + ov.y * gv.y);
    const float lse = lse_cls[(int64_t)b * heads + h];
    float ds[CLS_MAX_KEYS / 32];
#pragma unroll
    for (int i = 0; i < CLS_MAX_KEYS / 32; ++i) {
        const int j = lane + 32 * i;
        ds[i] = 0.f;
        if (j < N) {
            const float p = __expf(scale * dot64_h(base + (int64_t)j * ld + D, q) - lse);
            const float dp = dot64_h(base + (int64_t)j * ld + 2 * D, g);
            ds[i] = p * (dp - delta);
            // dV[j, :] = p * dO ; dK[j, :] = scale * dS_j * q   (this lane owns the whole 128-byte row of key j)
            uint4* dv4 = reinterpret_cast<uint4*>(dbase + (int64_t)j * lddqkv + 2 * D);
            uint4* dk4 = reinterpret_cast<uint4*>(dbase + (int64_t)j * lddqkv + D);
            const float kd = scale * ds[i];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint4 a, k4;
                a.x = pack_half2(p * g[8 * c], p * g[8 * c + 1]); a.y = pack_half2(p * g[8 * c + 2], p * g[8 * c + 3]);
                a.z = pack_half2(p * g[8 * c + 4], p * g[8 * c + 5]); a.w = pack_half2(p * g[8 * c + 6], p * g[8 * c + 7]);
                k4.x = pack_half2(kd * q[8 * c], kd * q[8 * c + 1]); k4.y = pack_half2(kd * q[8 * c + 2], kd * q[8 * c + 3]);
                k4.z = pack_half2(kd * q[8 * c + 4], kd * q[8 * c + 5]); k4.w = pack_half2(kd * q[8 * c + 6], kd * q[8 * c + 7]);
                dv4[c] = a;
                dk4[c] = k4;
            }
        }
    }
    // dq[d] = scale * sum_j dS_j K[j, d]
    float d0 = 0.f, d1 = 0.f;
    for (int j = 0; j < N; ++j) {
        float dsj = 0.f;
#pragma unroll
        for (int i = 0; i < CLS_MAX_KEYS / 32; ++i) if ((j >> 5) == i) dsj = __shfl_sync(0xffffffffu, ds[i], j & 31);
        const float2 kv = unpack_half2(__ldg(reinterpret_cast<const uint32_t*>(base + (int64_t)j * ld + D) + lane));
        d0 = fmaf(dsj, kv.x, d0);
        d1 = fmaf(dsj, kv.y, d1);
    }
    *reinterpret_cast<uint32_t*>(dbase + 2 * lane) = pack_half2(d0 * scale, d1 * scale);
}

int cls_attention_bwd(const __half* qkv, int64_t ld, const __half* o_cls, int64_t ldo, const __half* do_cls, int64_t lddo, const float* lse_cls,
                      __half* dqkv, int64_t lddqkv, int B, int N, int heads, float scale, cudaStream_t s) {
    GSL_REQUIRE(N <= CLS_MAX_KEYS && ld % 8 == 0 && lddqkv % 8 == 0 && lddo % 8 == 0, "cls_attention_bwd: tokens=%d > %d or unaligned pitch", N, CLS_MAX_KEYS);
    // dQ of every non-cls token is zero
    GSL_CHECK_CUDA(cudaMemset2DAsync(dqkv, (size_t)lddqkv * 2, 0, (size_t)heads * 64 * 2, (size_t)B * N, s));
    const int warps = 4;
    cls_attention_bwd_kernel<<<(B * heads + warps - 1) / warps, warps * 32, 0, s>>>(qkv, ld, o_cls, ldo, do_cls, lddo, lse_cls, dqkv, lddqkv, B, N, heads, scale);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// rows[b] = src[b * stride_rows] for b < B : gathers the cls rows of a [B*tokens, cols] matrix into [B, cols] (16-byte granules)
__global__ void gather_rows_kernel(const uint4* __restrict__ src, int64_t src_pitch16, uint4* __restrict__ dst, int64_t dst_pitch16, int B, int granules) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * granules) return;
    const int b = (int)(i / granules), g = (int)(i % granules);
    dst[(int64_t)b * dst_pitch16 + g] = src[(int64_t)b * src_pitch16 + g];
}

// copies row_bytes from every `tokens`-th row of src (pitch src_pitch_bytes) into consecutive rows of dst, or the reverse (scatter = 1)
int copy_cls_rows(const void* src, int64_t src_pitch_bytes, void* dst, int64_t dst_pitch_bytes, int B, int64_t row_bytes, cudaStream_t s) {
    GSL_REQUIRE(row_bytes % 16 == 0 && src_pitch_bytes % 16 == 0 && dst_pitch_bytes % 16 == 0, "copy_cls_rows: 16-byte granularity required");
    const int granules = (int)(row_bytes / 16);
    const int64_t total = (int64_t)B * granules;
    gather_rows_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>((const uint4*)src, src_pitch_bytes / 16, (uint4*)dst, dst_pitch_bytes / 16, B, granules);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gsl
