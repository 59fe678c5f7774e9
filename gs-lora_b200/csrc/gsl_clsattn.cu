// gslora-b200: last-block shortcut.  ViT_face pools the cls token only (vit_face.py:540), so in the LAST Transformer block
// every non-cls token is dead: its attention output, FFN output and all gradients through them never reach the loss.
// The engine therefore runs the last block's attention for the single cls query per (image, head) and the rest of that
// block on the B compacted cls rows.  These are the two kernels for that single-query attention (forward / backward):
// one warp per (image, head), K and V rows streamed from HBM exactly once (forward) / twice (backward), fp32 math.
// Results are identical to the dense computation restricted to what reaches the loss.
#include "gsl_common.cuh"
#include "gsl_kernels.h"

namespace gsl {

static constexpr int CLS_MAX_KEYS = 256;
static constexpr int CLS_ITERS = CLS_MAX_KEYS / 4;      // 4 keys per warp iteration: 8 lanes x 16 bytes cover one 128-byte K / V row

// lane layout: group g = lane / 8 owns keys j = 4 i + g, chunk c = lane % 8 owns head dims [8c, 8c + 8)
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = unpack_half2(w[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
}
__device__ __forceinline__ float group_sum8(float v) {      // sum over the 8 lanes of a key group
    v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}
__device__ __forceinline__ float across_groups(float v) {   // sum over the 4 key groups (same chunk)
    v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}

// qkv fp16 [B*N, ld] (q | k | v column blocks of heads*64).  o_cls fp16 [B, ldo] (heads*64 columns), lse_cls [B, heads].
__global__ void __launch_bounds__(128) cls_attention_fwd_kernel(const __half* __restrict__ qkv, int64_t ld, __half* __restrict__ o_cls, int64_t ldo,
                                                                float* __restrict__ lse_cls, int B, int N, int heads, float scale) {
    pdl_prologue();
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31, g = lane >> 3, c = lane & 7;
    if (w >= B * heads) return;
    const int b = w / heads, h = w % heads, D = heads * 64;
    const __half* base = qkv + (int64_t)b * N * ld + h * 64 + c * 8;
    float q[8];
    load8(base, q);                                           // token 0 = cls
    const int iters = (N + 3) >> 2;
    float s[CLS_ITERS];
    float mx = -INFINITY;
    // Branch-free in chunks of 8 iterations: rows past the end are CLAMPED (a valid address, the value is masked afterwards), so the eight
    // 16-byte loads of a chunk are independent straight-line code and go out together.  (With a per-lane `if (j < N)` around each load the
    // compiler emitted load -> dot -> shuffles strictly one key group after the other: 0.20 ms for 413 MB, pure latency.)
#pragma unroll
    for (int c8 = 0; c8 < CLS_ITERS; c8 += 8) {
        if (c8 < iters) {                                       // warp-uniform
            float kk[8][8];
#pragma unroll
            for (int u = 0; u < 8; ++u) load8(base + (int64_t)min(4 * (c8 + u) + g, N - 1) * ld + D, kk[u]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = 4 * (c8 + u) + g;
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < 8; ++d) acc = fmaf(kk[u][d], q[d], acc);
                acc = group_sum8(acc);
                s[c8 + u] = j < N ? acc * scale : -INFINITY;
                mx = fmaxf(mx, s[c8 + u]);
            }
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) s[c8 + u] = -INFINITY;
        }
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    float sum = 0.f, o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c8 = 0; c8 < CLS_ITERS; c8 += 8) {
        if (c8 < iters) {
            float vv[8][8];
#pragma unroll
            for (int u = 0; u < 8; ++u) load8(base + (int64_t)min(4 * (c8 + u) + g, N - 1) * ld + 2 * D, vv[u]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float p = __expf(s[c8 + u] - mx);         // masked keys: s = -inf -> p = 0
                sum += p;
#pragma unroll
                for (int d = 0; d < 8; ++d) o[d] = fmaf(p, vv[u][d], o[d]);
            }
        }
    }
    sum = across_groups(sum);
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d] = across_groups(o[d]);
    if (g == 0) {
        const float inv = 1.0f / sum;
        uint4 u;
        u.x = pack_half2(o[0] * inv, o[1] * inv); u.y = pack_half2(o[2] * inv, o[3] * inv);
        u.z = pack_half2(o[4] * inv, o[5] * inv); u.w = pack_half2(o[6] * inv, o[7] * inv);
        *reinterpret_cast<uint4*>(o_cls + (int64_t)b * ldo + h * 64 + c * 8) = u;
        if (c == 0) lse_cls[(int64_t)b * heads + h] = mx + __logf(sum);
    }
}

int cls_attention_fwd(const __half* qkv, int64_t ld, __half* o_cls, int64_t ldo, float* lse_cls, int B, int N, int heads, float scale, cudaStream_t s) {
    GSL_REQUIRE(N <= CLS_MAX_KEYS && ld % 8 == 0 && ldo % 8 == 0, "cls_attention: tokens=%d > %d or unaligned pitch", N, CLS_MAX_KEYS);
    const int warps = 4;
    GSL_CHECK_CUDA(launch_pdl(cls_attention_fwd_kernel, dim3((B * heads + warps - 1) / warps), dim3(warps * 32), 0, s, qkv, ld, o_cls, ldo, lse_cls, B, N, heads, scale));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Backward for the single cls query.  do_cls fp16 [B, lddo]; writes dqkv fp16 [B*N, lddqkv]: dQ row of token 0, dK / dV rows of every token.
// (dQ rows of tokens > 0 are zero: the launcher clears that column block.)
__global__ void __launch_bounds__(128) cls_attention_bwd_kernel(const __half* __restrict__ qkv, int64_t ld, const __half* __restrict__ o_cls, int64_t ldo,
                                                                const __half* __restrict__ do_cls, int64_t lddo, const float* __restrict__ lse_cls,
                                                                __half* __restrict__ dqkv, int64_t lddqkv, int B, int N, int heads, float scale) {
    pdl_prologue();
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31, g = lane >> 3, c = lane & 7;
    if (w >= B * heads) return;
    const int b = w / heads, h = w % heads, D = heads * 64;
    const __half* base = qkv + (int64_t)b * N * ld + h * 64 + c * 8;
    __half* dbase = dqkv + (int64_t)b * N * lddqkv + h * 64 + c * 8;
    float q[8], gr[8], ov[8];
    load8(base, q);
    load8(do_cls + (int64_t)b * lddo + h * 64 + c * 8, gr);
    load8(o_cls + (int64_t)b * ldo + h * 64 + c * 8, ov);
    float dl = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) dl = fmaf(gr[d], ov[d], dl);
    const float delta = group_sum8(dl);                       // dO . O
    const float lse = lse_cls[(int64_t)b * heads + h];
    const int iters = (N + 3) >> 2;
    float dq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        const int j = 4 * i + g;
        float sa = 0.f, da = 0.f;
        float k[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, v[8];
        if (j < N) {
            load8(base + (int64_t)j * ld + D, k);
            load8(base + (int64_t)j * ld + 2 * D, v);
#pragma unroll
            for (int d = 0; d < 8; ++d) { sa = fmaf(k[d], q[d], sa); da = fmaf(v[d], gr[d], da); }
        }
        sa = group_sum8(sa);
        da = group_sum8(da);
        if (j < N) {
            const float p = __expf(sa * scale - lse);
            const float ds = p * (da - delta);
            const float kd = scale * ds;
            uint4 a, k4;
            a.x = pack_half2(p * gr[0], p * gr[1]); a.y = pack_half2(p * gr[2], p * gr[3]);
            a.z = pack_half2(p * gr[4], p * gr[5]); a.w = pack_half2(p * gr[6], p * gr[7]);
            k4.x = pack_half2(kd * q[0], kd * q[1]); k4.y = pack_half2(kd * q[2], kd * q[3]);
            k4.z = pack_half2(kd * q[4], kd * q[5]); k4.w = pack_half2(kd * q[6], kd * q[7]);
            *reinterpret_cast<uint4*>(dbase + (int64_t)j * lddqkv + 2 * D) = a;          // dV[j] = p dO
            *reinterpret_cast<uint4*>(dbase + (int64_t)j * lddqkv + D) = k4;             // dK[j] = scale dS_j q
#pragma unroll
            for (int d = 0; d < 8; ++d) dq[d] = fmaf(ds, k[d], dq[d]);                    // dq += dS_j K[j]
        }
    }
#pragma unroll
    for (int d = 0; d < 8; ++d) dq[d] = across_groups(dq[d]);
    if (g == 0) {
        uint4 u;
        u.x = pack_half2(dq[0] * scale, dq[1] * scale); u.y = pack_half2(dq[2] * scale, dq[3] * scale);
        u.z = pack_half2(dq[4] * scale, dq[5] * scale); u.w = pack_half2(dq[6] * scale, dq[7] * scale);
        *reinterpret_cast<uint4*>(dbase) = u;                                              // dQ of the cls token
    }
}

int cls_attention_bwd(const __half* qkv, int64_t ld, const __half* o_cls, int64_t ldo, const __half* do_cls, int64_t lddo, const float* lse_cls,
                      __half* dqkv, int64_t lddqkv, int B, int N, int heads, float scale, cudaStream_t s) {
    GSL_REQUIRE(N <= CLS_MAX_KEYS && ld % 8 == 0 && lddqkv % 8 == 0 && lddo % 8 == 0 && ldo % 8 == 0, "cls_attention_bwd: tokens=%d > %d or unaligned pitch", N, CLS_MAX_KEYS);
    // dQ of every non-cls token is zero
    GSL_CHECK_CUDA(cudaMemset2DAsync(dqkv, (size_t)lddqkv * 2, 0, (size_t)heads * 64 * 2, (size_t)B * N, s));
    const int warps = 4;
    GSL_CHECK_CUDA(launch_pdl(cls_attention_bwd_kernel, dim3((B * heads + warps - 1) / warps), dim3(warps * 32), 0, s, qkv, ld, o_cls, ldo, do_cls, lddo, lse_cls, dqkv, lddqkv, B, N, heads, scale));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// rows[b] = src[b * stride_rows] for b < B : gathers the cls rows of a [B*tokens, cols] matrix into [B, cols] (16-byte granules)
__global__ void gather_rows_kernel(const uint4* __restrict__ src, int64_t src_pitch16, uint4* __restrict__ dst, int64_t dst_pitch16, int B, int granules) {
    pdl_prologue();
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
    GSL_CHECK_CUDA(launch_pdl(gather_rows_kernel, dim3((int)((total + 255) / 256)), dim3(256), 0, s, (const uint4*)src, src_pitch_bytes / 16, (uint4*)dst, dst_pitch_bytes / 16, B, granules));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gsl
