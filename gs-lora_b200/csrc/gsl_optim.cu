// gslora-b200: group-Lasso structure penalty fused into the AdamW update of the LoRA tensors.
//   engine_cl.get_structure_loss (engine_cl.py:349-432): loss_s = sum_g sqrt(sum_{P in g} sum P^2)
//   torch.optim.AdamW via timm create_optimizer (train_own_forget_cl.py:811-813), engine_cl.py:123-125
// One CTA per group (= the 4 LoRA matrices of one Transformer block, contiguous in the flat parameter
// buffer): warp-shuffle reduction of sum p^2 -> n_g, then in the same kernel
//   grad = g_data * grad_scale + alpha * p / n_g          (autograd of alpha * loss_s; 0 when n_g == 0,
//                                                           where the reference produces NaN)
//   m, v EMA ; p *= 1 - lr*wd ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// n_g (pre-update) is written out: sum_g n_g is the structure-loss scalar the reference logs.
// Overflow guard of the loss-scaled fp16 gradient stream: a group whose gradient holds an inf / NaN is NOT updated (parameters and
// moments keep their values) and reports n_g = NaN, which the host turns into a FloatingPointError when it reads the step's scalars.
#include "gsl_common.cuh"
#include "gsl_kernels.h"

namespace gsl {

__global__ void __launch_bounds__(1024) grouplasso_adamw_kernel(OptimArgs a, float bc1, float bc2_sqrt) {
    pdl_prologue();
    if (a.state != nullptr) {       // graph replay: the step count and the learning rate of THIS replay come from the device step state
        const int t = a.state->adam_step;
        a.lr = a.state->lr;
        bc1 = (float)(1.0 - pow((double)a.beta1, (double)t));
        bc2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, (double)t));
    }
    __shared__ float red[32];
    __shared__ float s_norm;
    const int gidx = blockIdx.x;
    const int64_t lo = a.group_offsets[gidx], hi = a.group_offsets[gidx + 1];
    float ss = 0.f;
    int bad = 0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float p = a.params[i];
        ss += p * p;
        bad |= !isfinite(a.grads[i]);
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    bad = __syncthreads_or(bad);
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            s_norm = sqrtf(t);
            if (a.group_norms) a.group_norms[gidx] = bad ? __int_as_float(0x7fc00000) : s_norm;
        }
    }
    __syncthreads();
    if (bad) return;
    const float norm = s_norm;
    const float lasso = (a.alpha != 0.f && norm > 0.f) ? a.alpha / norm : 0.f;
    const float decay = 1.0f - a.lr * a.wd;
    const float step = a.lr / bc1;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        float p = a.params[i];
        const float g = a.grads[i] * a.grad_scale + lasso * p;
        const float m = a.beta1 * a.m[i] + (1.0f - a.beta1) * g;
        const float v = a.beta2 * a.v[i] + (1.0f - a.beta2) * g * g;
        a.m[i] = m;
        a.v[i] = v;
        p *= decay;
        p -= step * m / (sqrtf(v) / bc2_sqrt + a.eps);
        a.params[i] = p;
    }
}

int grouplasso_adamw_step(const OptimArgs& a, cudaStream_t s) {
    GSL_REQUIRE(a.num_groups > 0 && (a.step >= 1 || a.state != nullptr), "optimizer: bad arguments (groups=%d step=%d)", a.num_groups, a.step);
    const int hstep = a.step >= 1 ? a.step : 1;
    const double bc1 = 1.0 - pow((double)a.beta1, (double)hstep);
    const double bc2 = 1.0 - pow((double)a.beta2, (double)hstep);
    GSL_CHECK_CUDA(launch_pdl(grouplasso_adamw_kernel, dim3(a.num_groups), dim3(1024), 0, s, a, (float)bc1, (float)sqrt(bc2)));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// util.cal_norm.get_norm_of_lora (util/cal_norm.py:121-143): per tensor ||P||_F or ||P||_1; the host sums
// the four tensors of a group (note: a different quantity from the group-lasso norm above).
__global__ void __launch_bounds__(256) tensor_norms_kernel(const float* __restrict__ params, const int* __restrict__ offs, int type, float* __restrict__ out) {
    __shared__ float red[8];
    const int tI = blockIdx.x;
    const int64_t lo = offs[tI], hi = offs[tI + 1];
    float acc = 0.f;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) { const float p = params[i]; acc += type == 0 ? p * p : fabsf(p); }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        out[tI] = type == 0 ? sqrtf(t) : t;
    }
}

int tensor_norms(const float* params, const int* tensor_offsets, int num_tensors, int type, float* out, cudaStream_t s) {
    tensor_norms_kernel<<<num_tensors, 256, 0, s>>>(params, tensor_offsets, type, out);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gsl
