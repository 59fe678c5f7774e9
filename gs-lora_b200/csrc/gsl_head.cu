// gslora-b200: classification head of ViT_face and its losses.
//   forward : cls row -> mlp_head LayerNorm (vit_face.py:498-500,540-543) -> CosFace (vit_face.py:171-208)
//             logits = s * (cos(emb, W_c) - m * onehot) -> per-sample cross entropy + top-1 hit
//             (nn.CrossEntropyLoss / train_accuracy of engine_cl.py:65-66,73-74)
//   backward: d logits (+ optional d emb, e.g. the GS-LoRA++ prototype term) -> gradient of the cls rows
//             of the final residual stream (normalize Jacobian + LayerNorm backward).
// Tiny (0.1 MFLOP / image): one CTA per image, fp32 SIMT, no tensor cores.
#include "gsl_common.cuh"
#include "gsl_kernels.h"

namespace gsl {

static constexpr int HEAD_THREADS = 128;
static constexpr int HEAD_MAX_D = 1024;
static constexpr int HEAD_MAX_C = 1024;

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    return t;
}

__global__ void __launch_bounds__(HEAD_THREADS) head_fwd_kernel(HeadArgs a) {
    __shared__ float s_e[HEAD_MAX_D];
    __shared__ float s_logit[HEAD_MAX_C];
    __shared__ float red[HEAD_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x, D = a.D, C = a.C;
    const float* xr = a.x + (int64_t)b * a.tokens * a.ldx;
    float sum = 0.f;
    for (int d = tid; d < D; d += HEAD_THREADS) { s_e[d] = xr[d]; sum += s_e[d]; }
    const float mean = block_sum(sum, red) / D;
    float sq = 0.f;
    for (int d = tid; d < D; d += HEAD_THREADS) { const float t = s_e[d] - mean; sq += t * t; }
    const float rstd = rsqrtf(block_sum(sq, red) / D + a.eps);
    float en = 0.f;
    for (int d = tid; d < D; d += HEAD_THREADS) {
        const float xh = (s_e[d] - mean) * rstd;
        const float e = xh * a.gamma[d] + a.beta[d];
        if (a.xhat) a.xhat[(int64_t)b * D + d] = xh;
        a.emb[(int64_t)b * D + d] = e;
        s_e[d] = e;
        en += e * e;
    }
    const float enorm = fmaxf(sqrtf(block_sum(en, red)), 1e-12f);     // F.normalize eps
    if (tid == 0 && a.rstd) a.rstd[b] = rstd;
    if (a.W == nullptr) return;
    const int label = a.labels ? (int)a.labels[b] : -1;
    const int warp = tid >> 5, lane = tid & 31;
    for (int c = warp; c < C; c += HEAD_THREADS / 32) {
        const float* w = a.W + (int64_t)c * D;
        float dot = 0.f, wn = 0.f;
        for (int d = lane; d < D; d += 32) { const float wv = __ldg(w + d); dot += wv * s_e[d]; wn += wv * wv; }
        dot = warp_sum(dot); wn = warp_sum(wn);
        if (lane == 0) {
            const float cosv = dot / (enorm * fmaxf(sqrtf(wn), 1e-12f));
            const float lg = a.head_type == 1 ? dot + (a.head_b ? a.head_b[c] : 0.f)          // heads.head Linear (modified_VIT.py:35-37)
                                              : a.cos_s * (c == label ? cosv - a.cos_m : cosv);
            s_logit[c] = lg;
            a.logits[(int64_t)b * C + c] = lg;
        }
    }
    __syncthreads();
    if (warp == 0) {
        float mx = -INFINITY; int arg = 0;
        for (int c = lane; c < C; c += 32) if (s_logit[c] > mx) { mx = s_logit[c]; arg = c; }
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
        }
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += __expf(s_logit[c] - mx);
        se = warp_sum(se);
        if (lane == 0) {
            if (a.ce) a.ce[b] = (label >= 0) ? (mx + logf(se) - s_logit[label]) : 0.f;
            if (a.correct) a.correct[b] = (arg == label) ? 1 : 0;
        }
    }
}

int head_fwd(const HeadArgs& a, cudaStream_t s) {
    GSL_REQUIRE(a.D <= HEAD_MAX_D && a.C <= HEAD_MAX_C, "head: D=%d C=%d exceed limits", a.D, a.C);
    head_fwd_kernel<<<a.B, HEAD_THREADS, 0, s>>>(a);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// dlogits[b, c] = scale * coef * (softmax(logits[b])[c] - onehot[c]);  coef lives on the device so the
// bounded-forget gate relu(BND - CE_f) (engine_cl.py:78) never needs a host round trip.
__global__ void ce_grad_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ coef_dev,
                               float scale, float* __restrict__ dlogits, int B, int C) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const float coef = (coef_dev ? coef_dev[0] : 1.0f) * scale;
    const float* lr = logits + (int64_t)b * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lr[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += __expf(lr[c] - mx);
    se = warp_sum(se);
    const float inv = 1.0f / se;
    const int label = (int)labels[b];
    for (int c = lane; c < C; c += 32) dlogits[(int64_t)b * C + c] = coef * (__expf(lr[c] - mx) * inv - (c == label ? 1.f : 0.f));
}

int ce_grad(const float* logits, const int64_t* labels, const float* coef_dev, float scale, float* dlogits, int B, int C, cudaStream_t s) {
    const int warps = 4;
    ce_grad_kernel<<<(B + warps - 1) / warps, warps * 32, 0, s>>>(logits, labels, coef_dev, scale, dlogits, B, C);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void __launch_bounds__(HEAD_THREADS) head_bwd_kernel(HeadBwdArgs a) {
    __shared__ float s_de[HEAD_MAX_D];     // d ehat, then d e
    __shared__ float s_dc[HEAD_MAX_C];     // d cos[c] / ||W_c||
    __shared__ float red[HEAD_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x, D = a.D, C = a.C;
    const int warp = tid >> 5, lane = tid & 31;
    const float* e = a.emb + (int64_t)b * D;
    float en = 0.f;
    for (int d = tid; d < D; d += HEAD_THREADS) { en += e[d] * e[d]; s_de[d] = 0.f; }
    const float enorm = fmaxf(sqrtf(block_sum(en, red)), 1e-12f);
    if (a.dlogits && a.W && a.head_type == 1) {
        // Linear head: d emb = d logits . W
        for (int c = tid; c < C; c += HEAD_THREADS) s_dc[c] = a.dlogits[(int64_t)b * C + c];
        __syncthreads();
        for (int d = tid; d < D; d += HEAD_THREADS) {
            float acc = 0.f;
            for (int c = 0; c < C; ++c) acc += s_dc[c] * __ldg(a.W + (int64_t)c * D + d);
            s_de[d] = acc;
        }
    } else if (a.dlogits && a.W) {
        for (int c = warp; c < C; c += HEAD_THREADS / 32) {
            const float* w = a.W + (int64_t)c * D;
            float wn = 0.f;
            for (int d = lane; d < D; d += 32) { const float wv = __ldg(w + d); wn += wv * wv; }
            wn = warp_sum(wn);
            if (lane == 0) s_dc[c] = a.cos_s * a.dlogits[(int64_t)b * C + c] / fmaxf(sqrtf(wn), 1e-12f);
        }
        __syncthreads();
        // d ehat[d] = sum_c dcos[c] * what[c, d]
        for (int d = tid; d < D; d += HEAD_THREADS) {
            float acc = 0.f;
            for (int c = 0; c < C; ++c) acc += s_dc[c] * __ldg(a.W + (int64_t)c * D + d);
            s_de[d] = acc;
        }
        __syncthreads();
        // d e = (d ehat - ehat * (ehat . d ehat)) / ||e||
        float dot = 0.f;
        for (int d = tid; d < D; d += HEAD_THREADS) dot += s_de[d] * e[d];
        dot = block_sum(dot, red) / enorm;
        for (int d = tid; d < D; d += HEAD_THREADS) s_de[d] = (s_de[d] - (e[d] / enorm) * dot) / enorm;
    }
    __syncthreads();
    // LayerNorm backward of mlp_head (frozen affine)
    float s1 = 0.f, s2 = 0.f;
    for (int d = tid; d < D; d += HEAD_THREADS) {
        float de = s_de[d];
        if (a.demb) de += a.demb[(int64_t)b * D + d];
        const float g = de * a.gamma[d];
        s_de[d] = g;
        s1 += g;
        s2 += g * a.xhat[(int64_t)b * D + d];
    }
    const float mg = block_sum(s1, red) / D;
    const float mgx = block_sum(s2, red) / D;
    const float rstd = a.rstd[b];
    for (int d = tid; d < D; d += HEAD_THREADS) {
        const float v = a.gscale * rstd * (s_de[d] - mg - a.xhat[(int64_t)b * D + d] * mgx);
        if (a.dx) a.dx[(int64_t)b * a.tokens * a.lddx + d] = v;
        if (a.dx16) {
            float m = 1.0f;
            if (a.drop_p > 0.f) {
                const uint32_t e = (uint32_t)(b * a.tokens) * (uint32_t)D + (uint32_t)d;
                const uint32_t h = drop_bits(e >> 1, a.drop_seed);
                const uint32_t bits = ((e & 1u) ? (h >> 16) : h) & 0x7FFFu;
                m = bits >= drop_thresh15(a.drop_p) ? 1.0f / (1.0f - a.drop_p) : 0.f;
            }
            a.dx16[(int64_t)b * a.tokens * a.lddx16 + d] = __float2half_rn(v * m);
        }
    }
}

int head_bwd(const HeadBwdArgs& a, cudaStream_t s) {
    GSL_REQUIRE(a.D <= HEAD_MAX_D && a.C <= HEAD_MAX_C, "head_bwd: D=%d C=%d exceed limits", a.D, a.C);
    head_bwd_kernel<<<a.B, HEAD_THREADS, 0, s>>>(a);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gsl
