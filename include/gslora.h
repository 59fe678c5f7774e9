/* gslora-b200 -- C ABI of the Blackwell-native GS-LoRA unlearning hot path.
 *
 * The reference (bjzhb666/GS-LoRA) is pure Python/PyTorch and has NO FFI: its seam is the Python module
 * surface.  These entry points are what a ctypes binding on the reference side calls
 * (see INTEGRATION.md); each cites the reference code it replaces.  Conventions:
 *   - every pointer is a raw DEVICE pointer owned by the caller (PyTorch's allocator); the library borrows
 *     it for the call, allocates nothing persistent and never synchronises; `stream` is a cudaStream_t.
 *   - int64 sizes / leading dimensions in ELEMENTS; return 0 = ok, otherwise a cudaError_t or -1 with a
 *     message available from gsl_last_error().
 *   - no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef GSLORA_H
#define GSLORA_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* gsl_last_error(void);
int gsl_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
long long gsl_launch_count(void);
/* default tcgen05 cta_group (1 or 2) for the GEMM family */
void gsl_set_gemm_cta_group(int cta_group);

/* epilogues of gsl_gemm_f16 */
enum {
  GSL_EPI_F16 = 0,          /* out0 fp16 = acc + bias                                              */
  GSL_EPI_F32 = 1,          /* out0 fp32 = acc + bias                     [+ out1 fp16 copy]        */
  GSL_EPI_GELU = 2,         /* h = acc + bias: out1 fp16 = Dropout(gelu_erf(h)), out0 fp16 = d out1/d h */
  GSL_EPI_GELU_BWD = 3,     /* out0 fp16 = acc * aux fp16  (aux = out0 of GSL_EPI_GELU)             */
  GSL_EPI_RES_F32 = 4,      /* out0 fp32 = acc + bias + aux fp32                                    */
  GSL_EPI_PERIODIC_F32 = 5, /* out0 fp32 = acc + aux_table fp32[row % aux_period]                   */
  GSL_EPI_F16_ROWDOT = 6    /* out0 fp16 = acc; out1 = fp32 [2][M / period][N / 64][period] (period = aux_period, M when 0): per row and per 64-column
                               block the dot product of acc with aux fp16, as the partial sums over the block's two 32-column halves */
};

/* C[M,N] = epi(A[M,K] * B[N,K]^T), fp16 operands, fp32 accumulation on tcgen05 tensor cores.
 * Replaces torch.nn.functional.linear / loralib.Linear.forward on the hot path
 * (vit_pytorch_face/vit_face.py:330-334,360,377,531; for loralib Linear.forward the caller passes the merged
 * operand B = fp16(W + s*lora_B*lora_A), which is what the engine caches) and the dX GEMMs
 * autograd builds for engine_cl.py:124.  drop_p > 0 applies a counter-based nn.Dropout mask (seed drop_seed, element index
 * row * N + col) to the produced value: before the residual add (RES), after the table add (PERIODIC), on out1 (GELU), and as a
 * factor in GELU_BWD. */
int gsl_gemm_f16(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                 int epi, const float* bias, void* out0, int64_t ld0, void* out1, int64_t ld1,
                 const void* aux, int64_t ldaux, int64_t aux_period, int cta_group, int block_n, float drop_p, uint32_t drop_seed,
                 void* stream);
/* The same with a split B operand: C = epi(A * (B + B_lo)^T), B_lo = fp16(W - B) the rounding residual of the fp32 weight W
 * (same shape / ldb as B).  Both terms are contracted against the SAME shared-memory A tile into one TMEM accumulator, so the
 * frozen weights enter with ~22 significand bits (GslConfig.precision = 1). */
int gsl_gemm_f16_split(const void* A, int64_t lda, const void* B, const void* B_lo, int64_t ldb, int64_t M, int64_t N, int64_t K,
                       int epi, const float* bias, void* out0, int64_t ld0, void* out1, int64_t ld1,
                       const void* aux, int64_t ldaux, int64_t aux_period, int cta_group, int block_n, float drop_p, uint32_t drop_seed,
                       void* stream);
/* Precision mode "split8": B = fp16(W * 2^shift), B_lo8 = e4m3(W * 2^shift - B) [N, K] bytes with the same pitch ldb (gsl_cast_f32_to_f16_split8):
   C = epi(2^-shift (A B^T + e5m2(A) B_lo8^T)).  The residual term runs on the FP8 tensor path (kind::f8f6f4) at twice the fp16 rate, the e5m2 copy
   of every A tile is made inside the kernel: 1.5x the tensor work of gsl_gemm_f16 instead of 2x.  Needs K % 64 == 0, ldb % 16 == 0; cta_group 2. */
int gsl_gemm_f16_split8(const void* A, int64_t lda, const void* B, const void* B_lo8, int shift, int64_t ldb, int64_t M, int64_t N, int64_t K, int epi,
                        const float* bias, void* out0, int64_t ld0, void* out1, int64_t ld1, const void* aux, int64_t ldaux,
                        int64_t aux_period, int block_n, float drop_p, uint32_t drop_seed, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Op-level entry points (each one kernel family; used by the engine and by the parity tests)
 * ---------------------------------------------------------------------------------------------------- */

/* einops 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' (vit_face.py:530) -> fp16 [B*(P+1), ld], zero row at token 0.
 * order 0 = (p1 p2 c) ViT_face, 1 = (c p1 p2) torchvision conv_proj. */
int gsl_patchify_f16(const float* img, void* out, int64_t ld, int B, int C, int S, int patch, int order, void* stream);
/* The same from raw uint8 pixels (layout 0 = NCHW as transforms.PILToTensor() stacks them, 1 = NHWC decoded rows): transforms.ToTensor()'s
 * `/ 255` and the optional transforms.Normalize(mean, std) of the ImageNet runs (train/train_own_forget_cl.py:138-139) happen in flight with
 * IEEE division, so the value rounded to fp16 is the one the reference's host transform produces.  mean / std: HOST pointers to C floats, or
 * both NULL.  Replaces the fp32 H2D copy of util/data_prefetcher.py:4-7 by one byte per pixel. */
int gsl_patchify_u8_f16(const uint8_t* img, int layout, const float* mean, const float* std, void* out, int64_t ld, int B, int C, int S,
                        int patch, int order, void* stream);
/* nn.LayerNorm forward (vit_face.py:316-323): fp32 in, fp16 out, row mean / rstd saved. */
int gsl_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y16, int64_t ldy,
                      float* mean, float* rstd, int64_t M, int D, void* stream);
/* LayerNorm backward w.r.t. the input (frozen affine) fused with the residual-gradient add. */
int gsl_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                      const float* gamma, const float* dres, int64_t lddres, float* dx, int64_t lddx, void* dx16, int64_t lddx16,
                      int64_t M, int D, void* stream);
/* T[M, 0:16] = X[M,K] * A16[16,K]^T   (loralib Linear.forward's  x @ A^T ; backward U = dY @ B with A16 = B^T).  1 <= r <= 16; rows >= r of
 * A16 are zero padding.  gsl_lora_down_split: A16 has 32 rows, [0,16) = fp16(A), [16,32) = fp16(A - fp16(A)); both feed one accumulator. */
int gsl_lora_down(const void* X16, int64_t ldx, const void* A16, int64_t lda, void* out16, int64_t ldo, int64_t M, int K, int r, void* stream);
int gsl_lora_down_split(const void* X16, int64_t ldx, const void* A32, int64_t lda, void* out16, int64_t ldo, int64_t M, int K, int r, void* stream);
/* out[n, j] (or out[j, n] if transpose_out) (+)= scale * sum_m L[m, n] * R[m, j]  -- dB = s dY^T T, dA = s U^T X. */
size_t gsl_skinny_tn_workspace(int64_t M, int N, int r);
int gsl_skinny_tn(const void* L16, int64_t ldl, const void* R16, int64_t ldr, float* out, int64_t ldo, int transpose_out, float scale,
                  int accumulate, int64_t M, int N, int r, float* workspace, size_t workspace_bytes, void* stream);
/* Fused side pass over a wide activation L16 [M, N] (one HBM read):  T16[M, 0:16] = L * P16[16, N]^T  (gsl_lora_down)  AND
 * out[n, j] (+)= scale * sum_m L[m, n] * R16[m, j]  (gsl_skinny_tn).  Backward of loralib Linear:
 * (T2 = G A2^T, dA2 = s U2^T G) and (U1 = dH B1, dB1 = s dH^T T1).  Workspace: gsl_lora_side_workspace bytes. */
size_t gsl_lora_side_workspace(int64_t M, int N, int r);
int gsl_lora_side(const void* L16, int64_t ldl, const void* P16, int64_t ldp, void* T16, int64_t ldt, const void* R16, int64_t ldr,
                  float* out, int64_t ldo, int transpose_out, float scale, int accumulate, int64_t M, int N, int r,
                  float* workspace, size_t workspace_bytes, void* stream);
/* gsl_lora_side with a split P operand (32 rows: hi | lo, as gsl_lora_down_split) */
int gsl_lora_side_split(const void* L16, int64_t ldl, const void* P32, int64_t ldp, void* T16, int64_t ldt, const void* R16, int64_t ldr,
                        float* out, int64_t ldo, int transpose_out, float scale, int accumulate, int64_t M, int N, int r,
                        float* workspace, size_t workspace_bytes, void* stream);
/* Attention.forward (vit_face.py:358-379) on qkv fp16 [B*N, ld] (q | k | v blocks of heads*64 columns). */
int gsl_attention_fwd(const void* qkv16, int64_t ld, void* out16, int64_t ldo, float* lse, int B, int N, int heads, float scale, void* stream);
int gsl_attention_bwd(const void* qkv16, int64_t ld, const void* out16, int64_t ldo, const void* dout16, int64_t lddo, const float* lse,
                      void* dqkv16, int64_t lddqkv, int B, int N, int heads, float scale, void* stream);
/* The same with delta = rowsum(dO * O) per (image, head, token) precomputed: rowdot = out1 of the GSL_EPI_F16_ROWDOT GEMM that produced dO
   (M = B*N rows, aux = the forward's attention output, aux_period = N), fp32 [2][B, heads, N]; the kernel adds the two parts. */
int gsl_attention_bwd_rowdot(const void* qkv16, int64_t ld, const void* dout16, int64_t lddo, const float* lse, const float* rowdot,
                             void* dqkv16, int64_t lddqkv, int B, int N, int heads, float scale, void* stream);
/* fp32 -> fp16 cast (optional scale / transpose) used to build the frozen-weight operand caches. */
int gsl_cast_f32_to_f16(const float* src, int64_t lds, void* dst16, int64_t ldd, int64_t rows, int64_t cols, float scale, int transpose, void* stream);
/* split form: dst16 = fp16(v), dst_lo16 = fp16(v - dst16), v = src * scale (the operand pair of gsl_gemm_f16_split) */
int gsl_cast_f32_to_f16_split(const float* src, int64_t lds, void* dst16, void* dst_lo16, int64_t ldd, int64_t rows, int64_t cols, float scale,
                              int transpose, void* stream);
/* dst16 = fp16(src * 2^shift), dst_lo8 = e4m3(src * 2^shift - dst16): the operand pair of gsl_gemm_f16_split8 */
int gsl_cast_f32_to_f16_split8(const float* src, int64_t lds, void* dst16, void* dst_lo8, int64_t ldd, int64_t rows, int64_t cols, int shift,
                               int transpose, void* stream);

/* Fused group-Lasso + AdamW (engine_cl.get_structure_loss engine_cl.py:349-432 + torch.optim.AdamW as built by
 * timm create_optimizer, train_own_forget_cl.py:811-813).  group_offsets: device int32 [G+1] element offsets.
 * group_norms (device [G], optional) receives the pre-update sqrt(sum p^2) of each group. */
int gsl_grouplasso_adamw_step(float* params, const float* grads, float* m, float* v, const int32_t* group_offsets, int num_groups,
                              int64_t n, float lr, float wd, float beta1, float beta2, float eps, float alpha, float grad_scale,
                              int step, float* group_norms, void* stream);
/* gsl_grouplasso_adamw_step for CUDA-graph capture: the 1-based step count (bias corrections) and the learning rate are read at run time from
 * the device step state { uint64 seed; int32 adam_step; float lr } at `state_dev`. */
int gsl_grouplasso_adamw_step_dev(float* params, const float* grads, float* m, float* v, const int32_t* group_offsets, int num_groups,
                                  int64_t n, float wd, float beta1, float beta2, float eps, float alpha, float grad_scale,
                                  const void* state_dev, float* group_norms, void* stream);
/* adds n to the launch counter (a replayed CUDA graph launches the kernels it captured without passing through this library) */
void gsl_count_launches(long long n);
/* util.cal_norm.get_norm_of_lora (util/cal_norm.py:121-143): out[t] = ||P_t||_F (type 0) or ||P_t||_1 (type 1). */
int gsl_tensor_norms(const float* params, const int32_t* tensor_offsets, int num_tensors, int type, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Engine: ViT_face forward / selective backward on caller-owned memory (vit_pytorch_face/vit_face.py:449-548,
 * engine_cl.py:59-125).  One engine per model instance and device.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct GslConfig {
  int32_t image_size, patch_size, channels, dim, depth, heads, mlp_dim, num_class, lora_rank;
  int32_t max_batch;     /* images per forward call */
  int32_t num_slots;     /* activation sets kept alive at once (autograd: one per un-backwarded forward) */
  int32_t patch_order;   /* 0 = (p1 p2 c) ViT_face, 1 = (c p1 p2) torchvision conv_proj */
  float attn_scale;      /* dim ** -0.5 for ViT_face (vit_face.py:346) */
  float ln_eps;          /* 1e-5 */
  float cos_s, cos_m;    /* CosFace s = 64, m = 0.35 (vit_face.py:158) */
  float lora_scaling;    /* lora_alpha / r = 1 / r */
  float grad_scale;      /* power-of-two loss scale of the fp16 gradient stream (unscaled again in dA/dB) */
  float dropout;         /* nn.Dropout p of to_out / after GELU / after fc2 (vit_face.py:332,334,356), applied when dropout_seed != 0 */
  float emb_dropout;     /* nn.Dropout p after the pos-embedding add (vit_face.py:489,537) */
  int32_t head_type;     /* 0 = CosFace on LN(cls) (ViT_face), 1 = Linear + bias on LN(cls) (torchvision heads.head, modified_VIT.py:23-39) */
  int32_t precision;     /* 0 = "fast": every frozen weight rounded to fp16 once (LoRA gradients ~1-2e-3 of the FP32 reference);
                            1 = "split": weights and LoRA factors enter the GEMMs as fp16 hi + lo pairs (22 significand bits; gradients <= 1e-3,
                            the north-star parity bar) at 2x the tensor-pipe work;
                            2 = "split8" (the Python surface's default): frozen weights as fp16(W 2^12) + e4m3 residual, the residual term on the FP8 tensor
                            path against an in-kernel e5m2 copy of the activation tile (~15 significand bits; gradients <= 1e-3) at 1.5x the tensor-pipe
                            work; LoRA factors as fp16 hi + lo pairs; needs mlp_dim % 64 == 0.  Activations are fp16 with fp32 accumulation in all three. */
  int32_t lora_pos;      /* 0 = "FFN": lora.Linear on net.0 / net.3 (every GS-LoRA script); 1 = "Attention": lora.MergedLinear(r, enable_lora = [T, T, T]) on
                            to_qkv and plain FFN Linears (vit_face.py:349-355, 405-425; engine.py:650-656) */
} GslConfig;

/* Frozen parameter pointer table order for gsl_engine_bind_params (fp32 device pointers, reference state_dict names):
 *   [0] pos_embedding  [1] cls_token  [2] patch_to_embedding.weight  [3] patch_to_embedding.bias
 *   [4] mlp_head.0.weight  [5] mlp_head.0.bias  [6] loss.weight  [7] NULL
 *   (torchvision family: encoder.pos_embedding, class_token, conv_proj.weight [D, C*p*p], conv_proj.bias, encoder.ln.{weight,bias},
 *    heads.head.weight, heads.head.bias; per block ln_1, self_attention.in_proj_{weight,bias}, out_proj, ln_2, mlp.0, mlp.3)
 *   then per block i (12 entries): 0.fn.norm.{weight,bias}, 0.fn.fn.to_qkv.{weight, bias(NULL for ViT_face)},
 *   0.fn.fn.to_out.0.{weight,bias}, 1.fn.norm.{weight,bias}, 1.fn.fn.net.0.{weight,bias}, 1.fn.fn.net.3.{weight,bias}
 * lora_flat / grad_flat: fp32 [depth][ lora_A(net.0) r*D | lora_B(net.0) H*r | lora_A(net.3) r*H | lora_B(net.3) D*r ]
 *   lora_pos = 1:       fp32 [depth][ to_qkv.lora_A 3r*D (A_q | A_k | A_v) | to_qkv.lora_B 3*inner*r (B_q | B_k | B_v) ]  */
#define GSL_NUM_GLOBAL_PARAMS 8
#define GSL_NUM_BLOCK_PARAMS 12

size_t gsl_engine_workspace_bytes(const GslConfig* cfg);
int gsl_engine_create(const GslConfig* cfg, void* workspace, size_t workspace_bytes, void** handle_out);
void gsl_engine_destroy(void* handle);
int gsl_engine_bind_params(void* handle, const void* const* frozen_ptrs, int num_ptrs, float* lora_flat, float* grad_flat);
/* rebuild the fp16 frozen-weight caches (after load_state_dict / loralib merge or unmerge mutated `weight`) */
int gsl_engine_refresh_frozen(void* handle, void* stream);
/* repack the fp16 LoRA operands (after any update of lora_A / lora_B) */
int gsl_engine_refresh_lora(void* handle, void* stream);
/* ViT_face.forward(img, label): img fp32 [B,C,S,S], labels int64 [B] (may be NULL: emb only).  use_lora = 0 runs the
 * merged / r == 0 form (F.linear only).  dropout_seed != 0 = train mode: the four nn.Dropout sites draw counter-based masks from
 * it (the backward of the same slot regenerates them).  Results stay in the slot: see gsl_engine_slot_ptr. */
int gsl_engine_forward(void* handle, int slot, const float* img, const int64_t* labels, int B, int use_lora, uint64_t dropout_seed,
                       void* stream);
/* gsl_engine_forward on raw uint8 pixels (see gsl_patchify_u8_f16 for layout / mean / std). */
int gsl_engine_forward_u8(void* handle, int slot, const uint8_t* img, int layout, const float* mean, const float* std, const int64_t* labels,
                          int B, int use_lora, uint64_t dropout_seed, void* stream);
/* gsl_engine_forward for CUDA-graph capture: train-mode dropout masks are derived ON THE DEVICE from the 64-bit base seed stored at `seed_dev`
 * (first field of the 16-byte step state { uint64 seed; int32 adam_step; float lr }), so a captured step can be replayed with a fresh mask per
 * replay by rewriting that block; dropout_on = 0 disables dropout.  The slot's backward reads the same block. */
int gsl_engine_forward_dev(void* handle, int slot, const void* img, int img_is_u8_layout, const int64_t* labels, int B, int use_lora,
                           int dropout_on, const void* seed_dev, void* stream);
/* selective backward of engine_cl.py:124: upstream d logits [B,C] and/or d emb [B,D] (fp32, may be NULL) ->
 * LoRA gradients written (accumulate = 0) or added (accumulate = 1) into grad_flat. */
int gsl_engine_backward(void* handle, int slot, const float* dlogits, const float* demb, int accumulate, void* stream);
/* XFINAL: fp32 [B, dim] -- the cls rows of the last block's output (the only rows of that block that are computed) */
enum { GSL_SLOT_EMB = 0, GSL_SLOT_LOGITS = 1, GSL_SLOT_CE = 2, GSL_SLOT_CORRECT = 3, GSL_SLOT_XFINAL = 4 };
void* gsl_engine_slot_ptr(void* handle, int slot, int what);
int64_t gsl_engine_lora_offset(void* handle, int block, int which);   /* which: 0 A(net.0) 1 B(net.0) 2 A(net.3) 3 B(net.3); lora_pos 1: 0 A(to_qkv) 1 B(to_qkv) */
int64_t gsl_engine_lora_numel(void* handle);

/* Step losses of engine_cl.train_one_epoch (engine_cl.py:65-80) on device, no host sync:
 *   sums[0..7] = { sum CE_remain, n_remain, sum CE_forget, n_forget, hits_remain, hits_forget, sum KL_remain, sum KL_forget } over ce/correct/kl[0:B]
 *   (kl may be null -> zeros),
 *   samples [0, n_remain) are the remain batch, [n_remain, B) the forget batch.  (Allreduce `sums` for data parallel.) */
int gsl_loss_sums(const float* ce, const int32_t* correct, const float* kl, int n_remain, int B, float* sums, void* stream);
/* GS-LoRA++ prototype term, engine_cl.get_prototype_loss (engine_cl.py:571-603, distance "kl") and its use at engine_cl.py:97-101:
 *   kl[b] = sum_d softmax(proto[label_b])_d * (log_softmax(proto[label_b])_d - log_softmax(emb_b)_d)      (batchmean is taken by gsl_loss_sums:
 *   sums[6] / sums[1] = KL_remain, sums[7] / sums[3] = KL_forget);  gsl_prototype_kl_grad writes d[w_f relu(BND_pro - KL_f) + w_r KL_r] / d emb. */
int gsl_prototype_kl_fwd(const float* emb, const int64_t* labels, const float* proto, int B, int D, float* kl, void* stream);
int gsl_prototype_kl_grad(const float* emb, const int64_t* labels, const float* proto, const float* sums, int n_remain_local, int B, int D,
                          float w_f, float w_r, float BND_pro, float* demb, void* stream);
/* Class prototypes, util.utils.calculate_prototypes (util/utils.py:502-549): per batch sums[label_b, :] += emb[b, :] (in batch order: the
 * reference's fp32 summation order) and counts[label_b] += 1 on caller-zeroed accumulators sums [C, D] / counts [C]; gsl_class_means writes
 * sums / counts (zero rows for classes never seen).  No host sync per sample (the reference does one `.item()` per image). */
int gsl_class_sums(const float* emb, const int64_t* labels, int B, int D, int C, float* sums, float* counts, void* stream);
int gsl_class_means(const float* sums, const float* counts, int C, int D, float* out, void* stream);
/* dlogits[b] = w_b * (softmax(logits[b]) - onehot(label_b)) with w_b = 1/n_remain (remain) or
 *   -beta * [CE_forget_mean < BND] / n_forget (forget), counts and means taken from `sums` (device). */
int gsl_unlearn_ce_grad(const float* logits, const int64_t* labels, const float* sums, int n_remain_local, int B, int C,
                        float beta, float BND, float* dlogits, void* stream);

#ifdef __cplusplus
}
#endif
#endif
