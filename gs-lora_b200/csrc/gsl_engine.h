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

struct BlockCache {             // fp16 operand caches (engine workspace)
    __half *qkv_w16, *qkv_wT16, *out_w16, *out_wT16;
    __half *fc1_cat;    // [H, D+16]  = [W1 | s*B1 | 0]
    __half *fc1T_cat;   // [D, H+16]  = [W1^T | s*A1^T | 0]
    __half *fc2_cat;    // [D, H+16]  = [W2 | s*B2 | 0]
    __half *fc2T_cat;   // [H, D+16]  = [W2^T | s*A2^T | 0]
    __half *A1h, *A2h;  // [16, D], [16, H]   lora_A (rows >= r zero)
    __half *B1T, *B2T;  // [16, H], [16, D]   lora_B^T
};

struct BlockActs {              // saved activations of one block for one slot
    float *ln1_mean, *ln1_rstd, *ln2_mean, *ln2_rstd, *lse;
    __half *qkv16, *o16, *xn2cat16, *h16, *gcat16;
};

struct ClsActs {                // last block, compacted to the B cls rows (see gsl_clsattn.cu)
    float *xin32, *xmid32, *xout32, *ln2_mean, *ln2_rstd, *lse;
    __half *o16, *xn2cat16, *h16, *gcat16;
};

struct Slot {
    std::vector<float*> x;      // 2L+1 residual-stream snapshots, fp32 [M, D]
    std::vector<BlockActs> blk;
    ClsActs cls;
    float *emb, *logits, *ce, *xhat, *head_rstd;
    int* correct;
    int batch = 0;
    int used_lora = 0;
    uint64_t drop_seed = 0;
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
    __half* patch_w16 = nullptr; float* posb = nullptr;
    std::vector<BlockCache> cache;
    std::vector<Slot> slots;
    // transients (shared by all slots)
    __half *patches16, *xn16, *dxcat16, *dhcat16, *do16, *dqkv16;
    float *dx32, *dxn32, *skinny_ws;
    float *cls_dx32, *cls_dxn32; __half *cls_dxcat16, *cls_dhcat16, *cls_do16;
    size_t skinny_ws_bytes = 0;
    void* pack_ptrs_dev = nullptr;
    int* group_offsets_dev = nullptr; int* tensor_offsets_dev = nullptr; float* group_norms_dev = nullptr; float* tensor_norms_dev = nullptr;
    bool params_bound = false;

    static size_t workspace_bytes(const GslConfig& c);
    int init(const GslConfig& c, void* workspace, size_t bytes);
    int bind_params(const void* const* ptrs, int n, float* lora, float* grads);
    int refresh_frozen(cudaStream_t s);
    int refresh_lora(cudaStream_t s);
    int forward(int slot, const float* img, const int64_t* labels, int B, int use_lora, uint64_t dropout_seed, cudaStream_t s);
    int backward(int slot, const float* dlogits, const float* demb, int accumulate, cudaStream_t s);
    int64_t lora_block_elems() const;
    int64_t lora_offset(int block, int which) const;   // which: 0 A1, 1 B1, 2 A2, 3 B2
private:
    size_t carve(bool assign);
    int ffn_forward(int l, int64_t M, __half* xn2cat, float* ln_mean, float* ln_rstd, const float* x_mid, __half* h16, __half* gcat, float* x_out,
                    int use_lora, float pdrop, uint64_t dseed, cudaStream_t s);
    int ffn_backward(int l, int64_t M, __half* dxcat, float* dx, __half* dhcat, float* dxn, const __half* xn2cat, const __half* h16, const __half* gcat,
                     const float* x_mid, const float* ln_mean, const float* ln_rstd, int accumulate, float pdrop, uint64_t dseed, cudaStream_t s);
};

}  // namespace gsl
