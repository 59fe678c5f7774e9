// gslora-b200: native engine that runs the GS-LoRA unlearning inner loop (forward / selective backward /
// fused optimizer) as a sequence of sm_100a kernels on caller-owned memory.  C++ side of include/gslora.h.
#pragma once
#include "../../include/gslora.h"
#include "gsl_kernels.h"

#include <vector>

namespace gsl {

struct BlockFrozen {            // fp32 parameters of one Transformer block (caller memory, reference names in gslora.h)
    const float *ln1_w, *ln1_b, *qkv_w, *qkv_b, *out_w, *out_b, *ln2_w, *ln2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
};

// A cached GEMM B operand: `hi` = fp16(W); `lo` = fp16(W - hi) in precision mode "split" (cfg.precision = 1), else null.
struct WOp {                    // one cached B operand of the GEMM family
    __half* hi = nullptr;       // fp16(W)  (precision "split8": fp16(W * 2^GSL_LO8_SHIFT))
    __half* lo = nullptr;       // precision "split": fp16(W - hi)
    uint8_t* lo8 = nullptr;     // precision "split8": e4m3(W * 2^GSL_LO8_SHIFT - hi)
};
// split8: power-of-two pre-scale of the (hi, lo8) pair.  The residual of a weight in the binade 2^e is <= 2^(e - 11): with the shift it lands
// in e4m3's normal range [2^-6, 448] for |W| from 2^-7 up to 16, and hi stays inside fp16 for |W| < 16 (ViT weights are O(0.01 .. 1)).
static constexpr int GSL_LO8_SHIFT = 12;

struct BlockCache {             // fp16 operand caches (engine workspace)
    WOp qkv_w16, qkv_wT16, out_w16, out_wT16;
    // FFN weights with the LoRA branch folded in while LoRA is live (un-merged train mode):  W' = W + s * B A, rounded to fp16 once per
    // optimizer step.  x W'^T = x W^T + s (x A^T) B^T and dY W' = dY W + s (dY B) A, so neither GEMM needs the rank-r intermediates.
    WOp fc1_w16;        // [H, D]  W1'
    WOp fc1T_w16;       // [D, H]  W1'^T   (B operand of dLN2 = dH W1')
    WOp fc2_w16;        // [D, H]  W2'
    WOp fc2T_w16;       // [H, D]  W2'^T   (B operand of dG = dY2 W2')
    // rank-r operands, 32 rows each: rows [0, 16) = fp16 value (rows >= r zero), rows [16, 32) = fp16 of the rounding residual (split mode, else zero)
    __half *A1h, *A2h;  // [32, D], [32, H]   lora_A    -> T = x A^T for dB
    __half *B1T, *B2T;  // [32, H], [32, D]   lora_B^T  -> U = dY B for dA
    // lora_pos = 1 (LoRA on to_qkv, loralib MergedLinear): one operand set per slice g = q, k, v
    __half *Aqkv;       // [3][32, D]      A_g
    __half *BqkvT;      // [3][32, inner]  B_g^T
};

struct BlockActs {              // saved activations of one block for one slot
    float *ln1_mean, *ln1_rstd, *ln2_mean, *ln2_rstd, *lse;
    __half *qkv16, *o16;
    __half *xn1_16;     // [M, D]  LN1(x), input of to_qkv: saved only with lora_pos = 1 (dA_g = s U_g^T LN1(x), T_g = LN1(x) A_g^T)
    __half *xn2_16;     // [M, D]  LN2(x), input of fc1
    __half *gp16;       // [M, H]  d Dropout(gelu(h)) / d h  (EPI_GELU out0)
    __half *g16;        // [M, H]  Dropout(gelu(h)), input of fc2
};

struct ClsActs {                // last block, compacted to the B cls rows (see gsl_clsattn.cu)
    float *xin32, *xmid32, *xout32, *ln2_mean, *ln2_rstd, *lse;
    __half *o16, *xn2_16, *gp16, *g16;
};

struct Slot {
    std::vector<float*> x;      // 2L+1 residual-stream snapshots, fp32 [M, D]
    std::vector<BlockActs> blk;
    ClsActs cls;
    float *emb, *logits, *ce, *xhat, *head_rstd;
    int* correct;
    int batch = 0;
    int used_lora = 0;
    uint64_t drop_seed = 0;         // base seed of the forward's dropout masks (0 = no dropout); with seed_dev set only its non-zero-ness matters
    const unsigned long long* seed_dev = nullptr;   // graph replay: the base seed lives in the device step state and is read by the kernels
};

class Engine {
public:
    GslConfig cfg;
    int tokens, patch_dim, M_max;
    // caller-owned parameter memory
    const float *pos_embedding, *cls_token, *patch_w, *patch_b, *head_ln_w, *head_ln_b, *loss_w, *head_b;
    std::vector<BlockFrozen> frozen;
    float* lora_flat = nullptr;     // [depth * (rD + Hr + rH + Dr)] fp32, block-major: A1, B1, A2, B2
    float* grad_flat = nullptr;
    // workspace
    uint8_t* ws = nullptr; size_t ws_bytes = 0;
    WOp patch_w16; float* posb = nullptr;
    std::vector<BlockCache> cache;
    std::vector<Slot> slots;
    // transients (shared by all slots)
    __half *patches16, *xn16, *dy16, *dh16, *do16, *dqkv16;
    float* attn_delta;      // [2][B, heads, tokens] partial sums of delta = rowsum(dO * O) (EPI_F16_ROWDOT -> attention_bwd)
    __half *t1_16, *t2_16, *u1_16, *u2_16;      // rank-r by-products [M, 16] of the block being back-propagated
    float *dx32, *dxn32, *skinny_ws;
    float *cls_dx32, *cls_dxn32; __half *cls_dy16, *cls_dh16, *cls_do16;
    size_t skinny_ws_bytes = 0;
    void* pack_ptrs_dev = nullptr;
    void* merge_jobs_dev = nullptr;
    int ffn_cache_mode = -1;        // -1 stale, 0 plain W, 1 W + s B A
    int* group_offsets_dev = nullptr; int* tensor_offsets_dev = nullptr; float* group_norms_dev = nullptr; float* tensor_norms_dev = nullptr;
    bool params_bound = false;

    static size_t workspace_bytes(const GslConfig& c);
    int init(const GslConfig& c, void* workspace, size_t bytes);
    int bind_params(const void* const* ptrs, int n, float* lora, float* grads);
    int refresh_frozen(cudaStream_t s);
    int refresh_lora(cudaStream_t s);
    // img_kind 0: fp32 NCHW (ToTensor output); 1 / 2: uint8 NCHW / NHWC with /255 and optional Normalize(mean, std) (host pointers) in flight
    int forward(int slot, const void* img, int img_kind, const float* mean, const float* std, const int64_t* labels, int B, int use_lora,
                uint64_t dropout_seed, cudaStream_t s, const unsigned long long* seed_dev = nullptr);
    int backward(int slot, const float* dlogits, const float* demb, int accumulate, cudaStream_t s);
    int64_t lora_block_elems() const;
    int64_t lora_offset(int block, int which) const;   // which: 0 A1, 1 B1, 2 A2, 3 B2  (lora_pos 1: 0 A_qkv, 1 B_qkv)
    int lora_tensors_per_block() const { return cfg.lora_pos == 1 ? 2 : 4; }
private:
    size_t carve(bool assign);
    int ensure_ffn_weights(int use_lora, cudaStream_t s);
    int ffn_forward(int l, int64_t M, __half* xn2, float* ln_mean, float* ln_rstd, const float* x_mid, __half* gp16, __half* g16, float* x_out,
                    float pdrop, uint64_t dseed, cudaStream_t s);
    DropSeed site(uint64_t base, int block, int site_id) const;     // immediate per-site seed, or its device-derived form while seed_dev_cur is set
    const unsigned long long* seed_dev_cur = nullptr;
    int attn_lora_grads(int l, int64_t M, const __half* dqkv, const __half* xn1, int accumulate, cudaStream_t s);
    int ffn_backward(int l, int64_t M, __half* dy, float* dx, __half* dh, float* dxn, const __half* xn2, const __half* gp16, const __half* g16,
                     const float* x_mid, const float* ln_mean, const float* ln_rstd, int accumulate, float pdrop, uint64_t dseed, cudaStream_t s);
};

}  // namespace gsl
