// gslora-b200: internal C++ launcher interface shared by the engine and the C-ABI (include/gslora.h).
// Every launcher is asynchronous on `stream`, returns 0 on success (else a cudaError_t / -1 with text in
// gsl_last_error()), allocates nothing and never synchronises.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gsl_common.cuh"

namespace gsl {

enum GemmEpi {
    EPI_F16 = 0,           // out0(fp16) = acc + bias
    EPI_F32 = 1,           // out0(fp32) = acc + bias                    [+ out1(fp16) copy]
    EPI_GELU = 2,          // h = acc + bias ; out1(fp16) = Dropout(gelu(h)) ; out0(fp16) = d out1 / d h   (FFN fc1; h itself is not stored)
    EPI_GELU_BWD = 3,      // out0(fp16) = acc * aux(fp16)                                  (dH = dG * [mask * gelu'(h)] saved by EPI_GELU)
    EPI_RES_F32 = 4,       // out0(fp32) = acc + bias + aux(fp32)                            (residual adds)
    EPI_PERIODIC_F32 = 5,  // out0(fp32) = acc + aux_table(fp32)[row % period]             (patch embed + pos/cls)
    EPI_F16_ROWDOT = 6,    // out0(fp16) = acc ; rowdot = per-row, per-64-column dot products of acc with aux(fp16), as two 32-column partial
                           // sums (dO = dY Wo with delta = rowsum(dO * O) per head for the attention backward)
};

struct GemmArgs {
    const __half* A = nullptr; int64_t lda = 0;     // [M, K] row-major
    const __half* B = nullptr; int64_t ldb = 0;     // [N, K] row-major (a torch Linear weight)
    const __half* B_lo = nullptr;                   // optional second term of a split weight: C = epi(A (B + B_lo)^T), same shape / ldb as B
    const uint8_t* B_lo8 = nullptr; int lo8_shift = 0;   // split8: B = fp16(W * 2^shift), B_lo8 = e4m3(W * 2^shift - B) [N, K] bytes (pitch ldb):
                                                    // C = epi(2^-shift (A B^T + e5m2(A) B_lo8^T)), the residual term on the FP8 tensor path
    int64_t M = 0, N = 0, K = 0;
    int epi = EPI_F16;
    const float* bias = nullptr;
    void* out0 = nullptr; int64_t ld0 = 0;
    void* out1 = nullptr; int64_t ld1 = 0;
    const void* aux = nullptr; int64_t ldaux = 0; int64_t aux_period = 0;
    float* rowdot = nullptr;   // EPI_F16_ROWDOT: [2][M / period][N / 64][period] fp32 (period = aux_period, or M when 0): part h holds the dot products over
                               // columns [64 c + 32 h, 64 c + 32 h + 32); the consumer adds the two parts
    int cta_group = 0;   // 0 = library default, 1 or 2
    int block_n = 0;     // 0 = auto, 128 or 256
    float drop_p = 0.f;  // dropout on the produced value (before the residual add; on both outputs of EPI_GELU)
    DropSeed drop_seed;  // immediate seed, or (dev, key): derived on the device from the step state (graph replay)
};
int gemm_f16(const GemmArgs& a, cudaStream_t stream);
void gemm_set_default_cta_group(int cg);
int device_sm_count();

// ---- elementwise / normalisation (gsl_rowops.cu)
// img [B,C,S,S] fp32 NCHW -> patches fp16 [B*(P+1), ld] with a zero row at token 0 of every image;
// order 0: (p1 p2 c) channel fastest (vit_face.py:530); order 1: (c p1 p2) (torchvision conv_proj)
int patchify_f16(const float* img, __half* out, int64_t ld, int B, int C, int S, int patch, int order, cudaStream_t s);
// the same from raw uint8 pixels (layout 0 NCHW, 1 NHWC): ToTensor's /255 and an optional Normalize(mean, std) (HOST pointers, C floats each, or
// both null) are applied in flight -- the input pipeline moves 1 byte per pixel over PCIe instead of 4
int patchify_u8_f16(const uint8_t* img, int layout, const float* mean, const float* std, __half* out, int64_t ld, int B, int C, int S, int patch,
                    int order, cudaStream_t s);
// y = LN(x) * gamma + beta -> fp16 [M, ldy] ; saves mean/rstd
int layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, __half* y, int64_t ldy,
                  float* mean, float* rstd, int64_t M, int D, cudaStream_t s);
// dx = dres + LNbwd(dy) ; writes fp32 dx and an fp16 copy (GEMM operand for the next dX GEMM)
int layernorm_bwd(const void* dy, int dy_is_fp16, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                  const float* gamma, const float* dres, int64_t lddres, float* dx, int64_t lddx, __half* dx16, int64_t lddx16,
                  int64_t M, int D, float drop_p, DropSeed drop_seed, cudaStream_t s,    // dropout mask applies to the fp16 copy only
                  int dres_period = 0);   // > 0: dres holds one compacted row per `dres_period` rows (added at rows r % period == 0, zero elsewhere)
// T[M, 0:16] = X[M, K] * A16[16, K]^T  (fp16 in, fp32 accumulate, fp16 out at out[:, 0:16], row pitch ldo)
// fold != 0: A16 has 32 rows, [0, 16) = fp16(A) and [16, 32) = fp16(A - fp16(A)); both halves feed the same accumulator (split-precision LoRA factor)
int lora_down(const __half* X, int64_t ldx, const __half* A16, int64_t lda, __half* out, int64_t ldo, int64_t M, int K, int r, cudaStream_t s,
              int fold = 0);
// dW[R, 16-ish] style skinny reductions over M (split-M partials + deterministic second pass):
//   out[n, j] = scale * sum_m  L[m, n] * Rm[m, j]   n < N, j < r    (L fp16 [M, ldl], Rm fp16 [M, ldr])
//   accumulate != 0 adds into `out` (second data stream / gradient accumulation)
int skinny_tn(const __half* L, int64_t ldl, const __half* Rm, int64_t ldr, float* out, int64_t ldo, int transpose_out,
              float scale, int accumulate, int64_t M, int N, int r, float* workspace, size_t workspace_bytes, cudaStream_t s);
size_t skinny_tn_workspace(int64_t M, int N, int r);
// Fused side pass over a wide activation L [M, N] (one read): T[M, 0:16] = L * P16[16, N]^T (fp16) AND
// out[n, j] = scale * sum_m L[m, n] * Rm[m, j] (same conventions as skinny_tn).  Rank 16 / N not a multiple of 256 run as the
// two separate kernels.  Workspace: lora_side_workspace bytes.
int lora_side(const __half* L, int64_t ldl, const __half* P16, int64_t ldp, __half* T, int64_t ldt, const __half* Rm, int64_t ldr,
              float* out, int64_t ldo, int transpose_out, float scale, int accumulate, int64_t M, int N, int r,
              float* workspace, size_t workspace_bytes, cudaStream_t s, int fold = 0);      // fold: P16 has 32 rows (hi | lo), as in lora_down
size_t lora_side_workspace(int64_t M, int N, int r);
// fp32 -> fp16 casts with optional scale / transpose / column placement (weight cache building, LoRA operand packing)
int cast_f32_to_f16(const float* src, int64_t lds, __half* dst, int64_t ldd, int64_t rows, int64_t cols, float scale,
                    int transpose, cudaStream_t s, __half* dst_lo = nullptr,     // dst_lo: fp16(v - dst), the second term of a split operand
                    uint8_t* dst_lo8 = nullptr);                                  // dst_lo8: e4m3(v - dst) (split8; pass scale = 2^shift)     // dst_lo: fp16(v - fp16(v)), same layout as dst
int fill_zero(void* ptr, size_t bytes, cudaStream_t s);

// ---- attention (gsl_attention_fwd.cu / gsl_attention_bwd.cu, tcgen05); qkv fp16 [B*N, ld] with q|k|v column blocks of heads*64
int attention_fwd(const __half* qkv, int64_t ld, __half* out, int64_t ldo, float* lse, int B, int N, int heads, float scale, cudaStream_t s);
int attention_bwd(const __half* qkv, int64_t ld, const __half* out, int64_t ldo, const __half* dout, int64_t lddo, const float* lse,
                  __half* dqkv, int64_t lddqkv, int B, int N, int heads, float scale, cudaStream_t s,
                  const float* delta_parts = nullptr);   // [2][B, heads, N]: delta = rowsum(dO * O) per head as the two partial sums of EPI_F16_ROWDOT (then `out` is unused)

// ---- last-block shortcut: single (cls) query attention + cls-row gather / scatter (gsl_clsattn.cu)
int cls_attention_fwd(const __half* qkv, int64_t ld, __half* o_cls, int64_t ldo, float* lse_cls, int B, int N, int heads, float scale, cudaStream_t s);
int cls_attention_bwd(const __half* qkv, int64_t ld, const __half* o_cls, int64_t ldo, const __half* do_cls, int64_t lddo, const float* lse_cls,
                      __half* dqkv, int64_t lddqkv, int B, int N, int heads, float scale, cudaStream_t s);
// dst row b <- src row b (row pitches in bytes): with src_pitch = tokens * pitch this gathers the cls rows, with dst_pitch = tokens * pitch it scatters
int copy_cls_rows(const void* src, int64_t src_pitch_bytes, void* dst, int64_t dst_pitch_bytes, int B, int64_t row_bytes, cudaStream_t s);

// ---- head / losses (gsl_head.cu)
struct HeadArgs {
    const float* x; int64_t ldx;      // final residual stream [B*N, D]; cls row = b * tokens
    int tokens;
    const float* gamma; const float* beta; float eps;    // mlp_head LayerNorm
    const float* W;                   // [C, D] CosFace weight (loss.weight) or Linear head weight
    const float* head_b = nullptr;    // [C] Linear head bias (head_type 1)
    int head_type = 0;                // 0 = CosFace (vit_face.py:171-208), 1 = Linear (torchvision heads.head)
    const int64_t* labels;            // [B]
    float cos_s, cos_m;
    int B, D, C;
    float* emb;                       // [B, D]
    float* logits;                    // [B, C]
    float* ce;                        // [B] per-sample cross entropy
    int* correct;                     // [B] argmax == label
    float* xhat;                      // [B, D] saved normalised cls rows
    float* rstd;                      // [B]
};
int head_fwd(const HeadArgs& a, cudaStream_t s);
// d logits / d emb (either may be null) -> gradient wrt the cls rows of the final residual stream, scaled by gscale;
// writes dx (fp32 [B*N, D], only cls rows touched) and dx16
struct HeadBwdArgs {
    const float* dlogits;             // [B, C] or null
    const float* demb;                // [B, D] or null
    const float* emb; const float* W; const int64_t* labels; const float* xhat; const float* rstd; const float* gamma;
    float cos_s; int B, D, C, tokens;
    int head_type = 0;
    float gscale;
    float* dx; int64_t lddx; __half* dx16; int64_t lddx16;
    float drop_p = 0.f; DropSeed drop_seed;          // mask of the last block's fc2-output dropout, applied to dx16 only
};
int head_bwd(const HeadBwdArgs& a, cudaStream_t s);
// dlogits[b, :] = coef * (softmax(logits[b]) - onehot) / B_total ; coef read from device (gate applied by caller)
int ce_grad(const float* logits, const int64_t* labels, const float* coef_dev, float scale, float* dlogits, int B, int C, cudaStream_t s);
int loss_sums(const float* ce, const int* correct, const float* kl, int n_remain, int B, float* sums, cudaStream_t s);
// GS-LoRA++ prototype term (engine_cl.py:571-603, 97-101): per-sample KL(log_softmax(emb) || log_softmax(proto[label])) and its gated gradient
int prototype_kl_fwd(const float* emb, const int64_t* labels, const float* proto, int B, int D, float* kl, cudaStream_t s);
int prototype_kl_grad(const float* emb, const int64_t* labels, const float* proto, const float* sums, int n_remain_local, int B, int D, float w_f,
                      float w_r, float BND_pro, float* demb, cudaStream_t s);
// calculate_prototypes (util/utils.py:502-549): sums[label_b, :] += emb[b, :] in batch order, counts[label_b] += 1; means = sums / counts (0 rows
// for classes never seen).  sums [C, D] / counts [C] are caller-zeroed accumulators carried across batches.
int class_sums(const float* emb, const int64_t* labels, int B, int D, int C, float* sums, float* counts, cudaStream_t s);
int class_means(const float* sums, const float* counts, int C, int D, float* out, cudaStream_t s);
int unlearn_ce_grad(const float* logits, const int64_t* labels, const float* sums, int n_remain_local, int B, int C, float beta, float BND,
                    float* dlogits, cudaStream_t s);

// ---- optimizer (gsl_optim.cu)
struct OptimArgs {
    float* params; const float* grads; float* m; float* v;   // flat fp32 buffers, `n` elements
    const int* group_offsets;   // device [G + 1] element offsets of the group-lasso groups in the flat buffer
    int num_groups; int64_t n;
    float lr, wd, beta1, beta2, eps, alpha, grad_scale;      // grad_scale multiplies grads (1 / loss scale / world)
    int step;                                                // 1-based
    const StepState* state = nullptr;                        // when set, `step` and `lr` are read from this device block at run time (graph replay)
    float* group_norms;         // device [G] : sqrt(sum p^2) per group, pre-update (structure loss terms)
};
int grouplasso_adamw_step(const OptimArgs& a, cudaStream_t s);
// per tensor Frobenius / L1 norms for util.cal_norm.get_norm_of_lora: out[t] = ||P_t||_F (type 0) or ||P_t||_1 (type 1)
int tensor_norms(const float* params, const int* tensor_offsets, int num_tensors, int type, float* out, cudaStream_t s);

}  // namespace gsl
