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
static constexpr int HEAD_IMGS = 4;          // images per CTA (head_fwd / head_bwd): W rows are shared by the CTA's images

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

// HEAD_IMGS images per CTA: every row of W is fetched once per CTA and dotted against all of its images (the per-image version re-read the
// whole [C, D] matrix from L2 for each image and was pure load latency); per (image, class) the summation order is unchanged.
__global__ void __launch_bounds__(HEAD_THREADS) head_fwd_kernel(HeadArgs a) {
    pdl_prologue();
    __shared__ float s_e[HEAD_IMGS][HEAD_MAX_D];
    __shared__ float s_logit[HEAD_IMGS][HEAD_MAX_C];
    __shared__ float red[HEAD_THREADS / 32];
    __shared__ float s_enorm[HEAD_IMGS];
    const int b0 = blockIdx.x * HEAD_IMGS, tid = threadIdx.x, D = a.D, C = a.C;
    const int nimg = min(HEAD_IMGS, a.B - b0);
    for (int i = 0; i < nimg; ++i) {
        const int b = b0 + i;
        const float* xr = a.x + (int64_t)b * a.tokens * a.ldx;
        float sum = 0.f;
        for (int d = tid; d < D; d += HEAD_THREADS) { s_e[i][d] = xr[d]; sum += s_e[i][d]; }
        const float mean = block_sum(sum, red) / D;
        float sq = 0.f;
        for (int d = tid; d < D; d += HEAD_THREADS) { const float t = s_e[i][d] - mean; sq += t * t; }
        const float rstd = rsqrtf(block_sum(sq, red) / D + a.eps);
        float en = 0.f;
        for (int d = tid; d < D; d += HEAD_THREADS) {
            const float xh = (s_e[i][d] - mean) * rstd;
            const float e = xh * a.gamma[d] + a.beta[d];
            if (a.xhat) a.xhat[(int64_t)b * D + d] = xh;
            a.emb[(int64_t)b * D + d] = e;
            s_e[i][d] = e;
            en += e * e;
        }
        const float enorm = fmaxf(sqrtf(block_sum(en, red)), 1e-12f);     // F.normalize eps
        if (tid == 0) { s_enorm[i] = enorm; if (a.rstd) a.rstd[b] = rstd; }
    }
    if (a.W == nullptr) return;
    __syncthreads();
    // a label outside [0, C) never indexes anything: its sample reports CE = NaN (the reference's F.cross_entropy would raise) and is never "correct"
    const int warp = tid >> 5, lane = tid & 31;
    long long label_raw = -1;
    if (lane < nimg && a.labels) label_raw = (long long)a.labels[b0 + lane];       // lane i keeps image i's label
    const bool label_bad = a.labels && lane < nimg && (label_raw < 0 || label_raw >= C);
    const int label = (a.labels && lane < nimg && !label_bad) ? (int)label_raw : -1;
    for (int c = warp; c < C; c += HEAD_THREADS / 32) {
        const float* w = a.W + (int64_t)c * D;
        float dot[HEAD_IMGS], wn = 0.f;
#pragma unroll
        for (int i = 0; i < HEAD_IMGS; ++i) dot[i] = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float wv = __ldg(w + d);
            wn += wv * wv;
#pragma unroll
            for (int i = 0; i < HEAD_IMGS; ++i) dot[i] += wv * s_e[i][d];
        }
        wn = warp_sum(wn);
        float mine = 0.f;
#pragma unroll
        for (int i = 0; i < HEAD_IMGS; ++i) { const float t = warp_sum(dot[i]); if (lane == i) mine = t; }
        if (lane < nimg) {
            const float cosv = mine / (s_enorm[lane] * fmaxf(sqrtf(wn), 1e-12f));
            const float lg = a.head_type == 1 ? mine + (a.head_b ? a.head_b[c] : 0.f)          // heads.head Linear (modified_VIT.py:35-37)
                                              : a.cos_s * (c == label ? cosv - a.cos_m : cosv);
            s_logit[lane][c] = lg;
            a.logits[(int64_t)(b0 + lane) * C + c] = lg;
        }
    }
    __syncthreads();
    for (int i = warp; i < nimg; i += HEAD_THREADS / 32) {      // one warp per image: arg max, log-sum-exp, CE
        const int lab_bad = __shfl_sync(0xffffffffu, (int)label_bad, i);
        const int lab = __shfl_sync(0xffffffffu, label, i);
        float mx = -INFINITY; int arg = 0;
        for (int c = lane; c < C; c += 32) if (s_logit[i][c] > mx) { mx = s_logit[i][c]; arg = c; }
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
        }
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += __expf(s_logit[i][c] - mx);
        se = warp_sum(se);
        if (lane == 0) {
            const int b = b0 + i;
            if (a.ce) a.ce[b] = lab_bad ? __int_as_float(0x7fc00000) : (lab >= 0) ? (mx + logf(se) - s_logit[i][lab]) : 0.f;
            if (a.correct) a.correct[b] = (arg == lab) ? 1 : 0;
        }
    }
}

int head_fwd(const HeadArgs& a, cudaStream_t s) {
    GSL_REQUIRE(a.D <= HEAD_MAX_D && a.C <= HEAD_MAX_C, "head: D=%d C=%d exceed limits", a.D, a.C);
    if (a.B == 0) return 0;
    GSL_CHECK_CUDA(launch_pdl(head_fwd_kernel, dim3((a.B + HEAD_IMGS - 1) / HEAD_IMGS), dim3(HEAD_THREADS), 0, s, a));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// dlogits[b, c] = scale * coef * (softmax(logits[b])[c] - onehot[c]);  coef lives on the device so the
// bounded-forget gate relu(BND - CE_f) (engine_cl.py:78) never needs a host round trip.
__global__ void ce_grad_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ coef_dev,
                               float scale, float* __restrict__ dlogits, int B, int C) {
    pdl_prologue();
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
    GSL_CHECK_CUDA(launch_pdl(ce_grad_kernel, dim3((B + warps - 1) / warps), dim3(warps * 32), 0, s, logits, labels, coef_dev, scale, dlogits, B, C));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void __launch_bounds__(HEAD_THREADS) head_bwd_kernel(HeadBwdArgs a) {
    pdl_prologue();
    __shared__ float s_de[HEAD_IMGS][HEAD_MAX_D];     // d ehat, then d e
    __shared__ float s_dc[HEAD_IMGS][HEAD_MAX_C];     // d cos[c] / ||W_c||
    __shared__ float red[HEAD_THREADS / 32];
    __shared__ float s_enorm[HEAD_IMGS];
    const int b0 = blockIdx.x * HEAD_IMGS, tid = threadIdx.x, D = a.D, C = a.C;
    const int nimg = min(HEAD_IMGS, a.B - b0);
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = 0; i < HEAD_IMGS; ++i) {
        const float* e = a.emb + (int64_t)(b0 + (i < nimg ? i : 0)) * D;
        float en = 0.f;
        for (int d = tid; d < D; d += HEAD_THREADS) { en += e[d] * e[d]; s_de[i][d] = 0.f; }
        const float enorm = fmaxf(sqrtf(block_sum(en, red)), 1e-12f);
        if (tid == 0) s_enorm[i] = enorm;
    }
    if (a.dlogits && a.W) {
        if (a.head_type == 1) {
            // Linear head: d emb = d logits . W
            for (int i = 0; i < HEAD_IMGS; ++i)
                for (int c = tid; c < C; c += HEAD_THREADS) s_dc[i][c] = i < nimg ? a.dlogits[(int64_t)(b0 + i) * C + c] : 0.f;
        } else {
            for (int c = warp; c < C; c += HEAD_THREADS / 32) {
                const float* w = a.W + (int64_t)c * D;
                float wn = 0.f;
                for (int d = lane; d < D; d += 32) { const float wv = __ldg(w + d); wn += wv * wv; }
                wn = warp_sum(wn);
                if (lane < HEAD_IMGS) s_dc[lane][c] = lane < nimg ? a.cos_s * a.dlogits[(int64_t)(b0 + lane) * C + c] / fmaxf(sqrtf(wn), 1e-12f) : 0.f;
            }
        }
        __syncthreads();
        // d ehat[d] = sum_c dcos[c] * what[c, d]  (Linear head: d emb directly); every W element is loaded once for the CTA's images
        for (int d = tid; d < D; d += HEAD_THREADS) {
            float acc[HEAD_IMGS];
#pragma unroll
            for (int i = 0; i < HEAD_IMGS; ++i) acc[i] = 0.f;
            for (int c = 0; c < C; ++c) {
                const float wv = __ldg(a.W + (int64_t)c * D + d);
#pragma unroll
                for (int i = 0; i < HEAD_IMGS; ++i) acc[i] += s_dc[i][c] * wv;
            }
#pragma unroll
            for (int i = 0; i < HEAD_IMGS; ++i) s_de[i][d] = acc[i];
        }
        if (a.head_type != 1) {
            __syncthreads();
            // d e = (d ehat - ehat * (ehat . d ehat)) / ||e||
            for (int i = 0; i < nimg; ++i) {
                const float* e = a.emb + (int64_t)(b0 + i) * D;
                const float enorm = s_enorm[i];
                float dot = 0.f;
                for (int d = tid; d < D; d += HEAD_THREADS) dot += s_de[i][d] * e[d];
                dot = block_sum(dot, red) / enorm;
                for (int d = tid; d < D; d += HEAD_THREADS) s_de[i][d] = (s_de[i][d] - (e[d] / enorm) * dot) / enorm;
            }
        }
    }
    __syncthreads();
    // LayerNorm backward of mlp_head (frozen affine)
    const uint32_t dseed = a.drop_p > 0.f ? drop_seed_resolve(a.drop_seed) : 0u;
    for (int i = 0; i < nimg; ++i) {
        const int b = b0 + i;
        float s1 = 0.f, s2 = 0.f;
        for (int d = tid; d < D; d += HEAD_THREADS) {
            float de = s_de[i][d];
            if (a.demb) de += a.demb[(int64_t)b * D + d];
            const float g = de * a.gamma[d];
            s_de[i][d] = g;
            s1 += g;
            s2 += g * a.xhat[(int64_t)b * D + d];
        }
        const float mg = block_sum(s1, red) / D;
        const float mgx = block_sum(s2, red) / D;
        const float rstd = a.rstd[b];
        for (int d = tid; d < D; d += HEAD_THREADS) {
            const float v = a.gscale * rstd * (s_de[i][d] - mg - a.xhat[(int64_t)b * D + d] * mgx);
            if (a.dx) a.dx[(int64_t)b * a.tokens * a.lddx + d] = v;
            if (a.dx16) {
                float m = 1.0f;
                if (a.drop_p > 0.f) {
                    const uint32_t e = (uint32_t)(b * a.tokens) * (uint32_t)D + (uint32_t)d;
                    const uint32_t h = drop_bits(e >> 1, dseed);
                    const uint32_t bits = ((e & 1u) ? (h >> 16) : h) & 0x7FFFu;
                    m = bits >= drop_thresh15(a.drop_p) ? 1.0f / (1.0f - a.drop_p) : 0.f;
                }
                a.dx16[(int64_t)b * a.tokens * a.lddx16 + d] = __float2half_rn(v * m);
            }
        }
    }
}

int head_bwd(const HeadBwdArgs& a, cudaStream_t s) {
    GSL_REQUIRE(a.D <= HEAD_MAX_D && a.C <= HEAD_MAX_C, "head_bwd: D=%d C=%d exceed limits", a.D, a.C);
    if (a.B == 0) return 0;
    GSL_CHECK_CUDA(launch_pdl(head_bwd_kernel, dim3((a.B + HEAD_IMGS - 1) / HEAD_IMGS), dim3(HEAD_THREADS), 0, s, a));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ GS-LoRA++ prototype term
// engine_cl.get_prototype_loss (engine_cl.py:571-603, distance "kl"):
//   F.kl_div(log_softmax(emb), log_softmax(proto[label]), reduction="batchmean", log_target=True) = mean_b sum_d pt_bd (lt_bd - lo_bd)
// used as  w_f * relu(BND_pro - KL_forget) + w_r * KL_remain  (engine_cl.py:97-101).  One warp per sample; the per-sample KL goes to
// loss_sums (-> sums[6], sums[7]); the gradient kernel recomputes the two softmaxes and scales them with the gate read from the
// (all-reduced) sums:  d/d emb_b = coef_b (softmax(emb_b) - softmax(proto_b)),  coef = w_r / n_r  or  -w_f [KL_f < BND_pro] / n_f.
__device__ __forceinline__ void proto_row_stats(const float* __restrict__ e, const float* __restrict__ p, int D, int lane, float& me, float& se,
                                                float& mp, float& sp) {
    me = -INFINITY; mp = -INFINITY;
    for (int d = lane; d < D; d += 32) { me = fmaxf(me, e[d]); mp = fmaxf(mp, p[d]); }
    me = warp_max(me); mp = warp_max(mp);
    se = 0.f; sp = 0.f;
    for (int d = lane; d < D; d += 32) { se += __expf(e[d] - me); sp += __expf(p[d] - mp); }
    se = warp_sum(se); sp = warp_sum(sp);
}
__global__ void prototype_kl_fwd_kernel(const float* __restrict__ emb, const int64_t* __restrict__ labels, const float* __restrict__ proto, int B, int D,
                                        float* __restrict__ kl) {
    pdl_prologue();
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const float* e = emb + (int64_t)b * D;
    const float* p = proto + labels[b] * D;
    float me, se, mp, sp;
    proto_row_stats(e, p, D, lane, me, se, mp, sp);
    const float le = me + __logf(se), lp = mp + __logf(sp);     // log-sum-exp of both rows
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float lt = p[d] - lp, lo = e[d] - le;
        acc += __expf(lt) * (lt - lo);
    }
    acc = warp_sum(acc);
    if (lane == 0) kl[b] = acc;
}
__global__ void prototype_kl_grad_kernel(const float* __restrict__ emb, const int64_t* __restrict__ labels, const float* __restrict__ proto,
                                         const float* __restrict__ sums, int n_remain_local, int B, int D, float w_f, float w_r, float BND_pro,
                                         float* __restrict__ demb) {
    pdl_prologue();
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    float coef;
    if (b < n_remain_local) coef = sums[1] > 0.f ? w_r / sums[1] : 0.f;
    else {
        const float mean_f = sums[3] > 0.f ? sums[7] / sums[3] : 0.f;
        coef = (sums[3] > 0.f && mean_f < BND_pro) ? -w_f / sums[3] : 0.f;      // relu'(BND_pro - KL_f) = [KL_f < BND_pro]
    }
    const float* e = emb + (int64_t)b * D;
    const float* p = proto + labels[b] * D;
    float me, se, mp, sp;
    proto_row_stats(e, p, D, lane, me, se, mp, sp);
    const float ie = 1.0f / se, ip = 1.0f / sp;
    for (int d = lane; d < D; d += 32) demb[(int64_t)b * D + d] = coef * (__expf(e[d] - me) * ie - __expf(p[d] - mp) * ip);
}
int prototype_kl_fwd(const float* emb, const int64_t* labels, const float* proto, int B, int D, float* kl, cudaStream_t s) {
    const int warps = 4;
    GSL_CHECK_CUDA(launch_pdl(prototype_kl_fwd_kernel, dim3((B + warps - 1) / warps), dim3(warps * 32), 0, s, emb, labels, proto, B, D, kl));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int prototype_kl_grad(const float* emb, const int64_t* labels, const float* proto, const float* sums, int n_remain_local, int B, int D, float w_f,
                      float w_r, float BND_pro, float* demb, cudaStream_t s) {
    const int warps = 4;
    GSL_CHECK_CUDA(launch_pdl(prototype_kl_grad_kernel, dim3((B + warps - 1) / warps), dim3(warps * 32), 0, s, emb, labels, proto, sums, n_remain_local, B, D, w_f, w_r, BND_pro, demb));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ class prototypes
// util.utils.calculate_prototypes (util/utils.py:502-549): the reference walks every sample on the host (`embeds_sum[label.item()] += embed`,
// one device sync per image) and divides by the count at the end.  Here one CTA per class scans the batch's labels and adds the matching
// embedding rows IN BATCH ORDER into the class row (each thread owns its columns, so the fp32 summation order is the reference's:
// dataset order), counts go to a float per class; class_means scales by the correctly rounded reciprocal of the count, which is how ATen
// evaluates `cuda_tensor / python_int`.  Deterministic, no atomics, no host sync.
__global__ void class_sums_kernel(const float* __restrict__ emb, const int64_t* __restrict__ labels, int B, int D, float* __restrict__ sums,
                                  float* __restrict__ counts) {
    const int c = blockIdx.x;
    __shared__ int s_hit[256];
    constexpr int MAXV = 8;                       // D <= 8 * blockDim.x
    float acc[MAXV];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int d = threadIdx.x + i * blockDim.x;
        acc[i] = d < D ? sums[(int64_t)c * D + d] : 0.f;
    }
    int n = 0;
    for (int b0 = 0; b0 < B; b0 += blockDim.x) {
        const int b = b0 + threadIdx.x;
        s_hit[threadIdx.x] = (b < B && labels[b] == (int64_t)c) ? 1 : 0;
        __syncthreads();
        const int lim = min((int)blockDim.x, B - b0);
        for (int j = 0; j < lim; ++j) {
            if (!s_hit[j]) continue;
            const float* e = emb + (int64_t)(b0 + j) * D;
#pragma unroll
            for (int i = 0; i < MAXV; ++i) {
                const int d = threadIdx.x + i * blockDim.x;
                if (d < D) acc[i] += e[d];
            }
            ++n;
        }
        __syncthreads();
    }
    if (n) {
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int d = threadIdx.x + i * blockDim.x;
            if (d < D) sums[(int64_t)c * D + d] = acc[i];
        }
        if (threadIdx.x == 0) counts[c] += (float)n;
    }
}
__global__ void class_means_kernel(const float* __restrict__ sums, const float* __restrict__ counts, int C, int D, float* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)C * D) return;
    const float n = counts[i / D];
    // `feature_sum / count` runs on the device in the reference (util/utils.py:547), where ATen divides a CUDA tensor by a host scalar as
    // a * (1.0f / b): reproduce exactly that rounding
    out[i] = n > 0.f ? __fmul_rn(sums[i], __frcp_rn(n)) : 0.f;
}
int class_sums(const float* emb, const int64_t* labels, int B, int D, int C, float* sums, float* counts, cudaStream_t s) {
    GSL_REQUIRE(D >= 1 && D <= 8 * 256, "class_sums: embedding width %d outside [1, 2048]", D);
    GSL_REQUIRE(C >= 1 && B >= 0, "class_sums: bad sizes");
    if (B == 0) return 0;
    class_sums_kernel<<<C, 256, 0, s>>>(emb, labels, B, D, sums, counts);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int class_means(const float* sums, const float* counts, int C, int D, float* out, cudaStream_t s) {
    const int64_t n = (int64_t)C * D;
    class_means_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(sums, counts, C, D, out);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gsl
