// gslora-b200: extern "C" surface declared in include/gslora.h.
#include "../../include/gslora.h"
#include "gsl_engine.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include <atomic>

namespace gsl {
std::atomic<long long> g_launches{0};
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
bool pdl_enabled() {
    // opt-in: measured on B200 (profiles/r01r_pdl_ab.md) the early-trigger form is 0.8 ms / step SLOWER than plain stream order
    static const bool on = [] { const char* e = getenv("GSLORA_PDL"); return e && e[0] == '1'; }();
    return on;
}
}  // namespace gsl

using namespace gsl;

extern "C" {

const char* gsl_last_error(void) { return g_err; }
int gsl_version(void) { return 100; }
long long gsl_launch_count(void) { return g_launches.load(); }
void gsl_set_gemm_cta_group(int cta_group) { gemm_set_default_cta_group(cta_group); }

int gsl_gemm_f16(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K, int epi,
                 const float* bias, void* out0, int64_t ld0, void* out1, int64_t ld1, const void* aux, int64_t ldaux,
                 int64_t aux_period, int cta_group, int block_n, float drop_p, uint32_t drop_seed, void* stream) {
    GemmArgs a;
    a.A = (const __half*)A; a.lda = lda; a.B = (const __half*)B; a.ldb = ldb;
    a.M = M; a.N = N; a.K = K; a.epi = epi; a.bias = bias;
    a.out0 = out0; a.ld0 = ld0; a.out1 = out1; a.ld1 = ld1;
    a.aux = aux; a.ldaux = ldaux; a.aux_period = aux_period;
    if (epi == EPI_F16_ROWDOT) { a.rowdot = (float*)out1; a.out1 = nullptr; }
    a.cta_group = cta_group; a.block_n = block_n; a.drop_p = drop_p; a.drop_seed = drop_seed;
    return gemm_f16(a, (cudaStream_t)stream);
}
int gsl_gemm_f16_split(const void* A, int64_t lda, const void* B, const void* B_lo, int64_t ldb, int64_t M, int64_t N, int64_t K, int epi,
                       const float* bias, void* out0, int64_t ld0, void* out1, int64_t ld1, const void* aux, int64_t ldaux,
                       int64_t aux_period, int cta_group, int block_n, float drop_p, uint32_t drop_seed, void* stream) {
    if (B_lo == nullptr) { set_last_error("gsl_gemm_f16_split: B_lo is null"); return -1; }
    GemmArgs a;
    a.A = (const __half*)A; a.lda = lda; a.B = (const __half*)B; a.B_lo = (const __half*)B_lo; a.ldb = ldb;
    a.M = M; a.N = N; a.K = K; a.epi = epi; a.bias = bias;
    a.out0 = out0; a.ld0 = ld0; a.out1 = out1; a.ld1 = ld1;
    a.aux = aux; a.ldaux = ldaux; a.aux_period = aux_period;
    if (epi == EPI_F16_ROWDOT) { a.rowdot = (float*)out1; a.out1 = nullptr; }
    a.cta_group = cta_group; a.block_n = block_n; a.drop_p = drop_p; a.drop_seed = drop_seed;
    return gemm_f16(a, (cudaStream_t)stream);
}

int gsl_gemm_f16_split8(const void* A, int64_t lda, const void* B, const void* B_lo8, int shift, int64_t ldb, int64_t M, int64_t N, int64_t K, int epi,
                        const float* bias, void* out0, int64_t ld0, void* out1, int64_t ld1, const void* aux, int64_t ldaux,
                        int64_t aux_period, int block_n, float drop_p, uint32_t drop_seed, void* stream) {
    if (B_lo8 == nullptr) { set_last_error("gsl_gemm_f16_split8: B_lo8 is null"); return -1; }
    GemmArgs a;
    a.A = (const __half*)A; a.lda = lda; a.B = (const __half*)B; a.B_lo8 = (const uint8_t*)B_lo8; a.lo8_shift = shift; a.ldb = ldb;
    a.M = M; a.N = N; a.K = K; a.epi = epi; a.bias = bias;
    a.out0 = out0; a.ld0 = ld0; a.out1 = out1; a.ld1 = ld1;
    a.aux = aux; a.ldaux = ldaux; a.aux_period = aux_period;
    if (epi == EPI_F16_ROWDOT) { a.rowdot = (float*)out1; a.out1 = nullptr; }
    a.cta_group = 2; a.block_n = block_n; a.drop_p = drop_p; a.drop_seed = drop_seed;
    return gemm_f16(a, (cudaStream_t)stream);
}

#define ST(x) ((cudaStream_t)(x))

int gsl_patchify_f16(const float* img, void* out, int64_t ld, int B, int C, int S, int patch, int order, void* stream) {
    return patchify_f16(img, (__half*)out, ld, B, C, S, patch, order, ST(stream));
}
int gsl_patchify_u8_f16(const uint8_t* img, int layout, const float* mean, const float* std, void* out, int64_t ld, int B, int C, int S, int patch,
                        int order, void* stream) {
    return patchify_u8_f16(img, layout, mean, std, (__half*)out, ld, B, C, S, patch, order, ST(stream));
}
int gsl_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y16, int64_t ldy, float* mean,
                      float* rstd, int64_t M, int D, void* stream) {
    return layernorm_fwd(x, ldx, gamma, beta, eps, (__half*)y16, ldy, mean, rstd, M, D, ST(stream));
}
int gsl_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd, const float* gamma,
                      const float* dres, int64_t lddres, float* dx, int64_t lddx, void* dx16, int64_t lddx16, int64_t M, int D, void* stream) {
    return layernorm_bwd(dy, 0, lddy, x, ldx, mean, rstd, gamma, dres, lddres, dx, lddx, (__half*)dx16, lddx16, M, D, 0.f, 0u, ST(stream));
}
int gsl_lora_down(const void* X16, int64_t ldx, const void* A16, int64_t lda, void* out16, int64_t ldo, int64_t M, int K, int r, void* stream) {
    return lora_down((const __half*)X16, ldx, (const __half*)A16, lda, (__half*)out16, ldo, M, K, r, ST(stream));
}
int gsl_lora_down_split(const void* X16, int64_t ldx, const void* A32, int64_t lda, void* out16, int64_t ldo, int64_t M, int K, int r, void* stream) {
    return lora_down((const __half*)X16, ldx, (const __half*)A32, lda, (__half*)out16, ldo, M, K, r, ST(stream), 1);
}
size_t gsl_skinny_tn_workspace(int64_t M, int N, int r) { return skinny_tn_workspace(M, N, r); }
int gsl_skinny_tn(const void* L16, int64_t ldl, const void* R16, int64_t ldr, float* out, int64_t ldo, int transpose_out, float scale,
                  int accumulate, int64_t M, int N, int r, float* workspace, size_t workspace_bytes, void* stream) {
    return skinny_tn((const __half*)L16, ldl, (const __half*)R16, ldr, out, ldo, transpose_out, scale, accumulate, M, N, r, workspace,
                     workspace_bytes, ST(stream));
}
size_t gsl_lora_side_workspace(int64_t M, int N, int r) { return lora_side_workspace(M, N, r); }
int gsl_lora_side(const void* L16, int64_t ldl, const void* P16, int64_t ldp, void* T16, int64_t ldt, const void* R16, int64_t ldr,
                  float* out, int64_t ldo, int transpose_out, float scale, int accumulate, int64_t M, int N, int r,
                  float* workspace, size_t workspace_bytes, void* stream) {
    return lora_side((const __half*)L16, ldl, (const __half*)P16, ldp, (__half*)T16, ldt, (const __half*)R16, ldr, out, ldo, transpose_out, scale,
                     accumulate, M, N, r, workspace, workspace_bytes, ST(stream));
}
int gsl_lora_side_split(const void* L16, int64_t ldl, const void* P32, int64_t ldp, void* T16, int64_t ldt, const void* R16, int64_t ldr,
                        float* out, int64_t ldo, int transpose_out, float scale, int accumulate, int64_t M, int N, int r,
                        float* workspace, size_t workspace_bytes, void* stream) {
    return lora_side((const __half*)L16, ldl, (const __half*)P32, ldp, (__half*)T16, ldt, (const __half*)R16, ldr, out, ldo, transpose_out, scale,
                     accumulate, M, N, r, workspace, workspace_bytes, ST(stream), 1);
}
int gsl_attention_fwd(const void* qkv16, int64_t ld, void* out16, int64_t ldo, float* lse, int B, int N, int heads, float scale, void* stream) {
    return attention_fwd((const __half*)qkv16, ld, (__half*)out16, ldo, lse, B, N, heads, scale, ST(stream));
}
int gsl_attention_bwd(const void* qkv16, int64_t ld, const void* out16, int64_t ldo, const void* dout16, int64_t lddo, const float* lse,
                      void* dqkv16, int64_t lddqkv, int B, int N, int heads, float scale, void* stream) {
    return attention_bwd((const __half*)qkv16, ld, (const __half*)out16, ldo, (const __half*)dout16, lddo, lse, (__half*)dqkv16, lddqkv, B, N,
                         heads, scale, ST(stream), nullptr);
}
int gsl_attention_bwd_rowdot(const void* qkv16, int64_t ld, const void* dout16, int64_t lddo, const float* lse, const float* rowdot,
                             void* dqkv16, int64_t lddqkv, int B, int N, int heads, float scale, void* stream) {
    if (rowdot == nullptr) { set_last_error("gsl_attention_bwd_rowdot: rowdot is null"); return -1; }
    return attention_bwd((const __half*)qkv16, ld, nullptr, 0, (const __half*)dout16, lddo, lse, (__half*)dqkv16, lddqkv, B, N, heads, scale,
                         ST(stream), rowdot);
}
int gsl_cast_f32_to_f16(const float* src, int64_t lds, void* dst16, int64_t ldd, int64_t rows, int64_t cols, float scale, int transpose,
                        void* stream) {
    return cast_f32_to_f16(src, lds, (__half*)dst16, ldd, rows, cols, scale, transpose, ST(stream));
}
int gsl_cast_f32_to_f16_split8(const float* src, int64_t lds, void* dst16, void* dst_lo8, int64_t ldd, int64_t rows, int64_t cols, int shift,
                               int transpose, void* stream) {
    if (shift < 0 || shift > 24) { set_last_error("gsl_cast_f32_to_f16_split8: shift %d outside [0, 24]", shift); return -1; }
    return cast_f32_to_f16(src, lds, (__half*)dst16, ldd, rows, cols, ldexpf(1.0f, shift), transpose, ST(stream), nullptr, (uint8_t*)dst_lo8);
}
int gsl_cast_f32_to_f16_split(const float* src, int64_t lds, void* dst16, void* dst_lo16, int64_t ldd, int64_t rows, int64_t cols, float scale,
                              int transpose, void* stream) {
    return cast_f32_to_f16(src, lds, (__half*)dst16, ldd, rows, cols, scale, transpose, ST(stream), (__half*)dst_lo16);
}
int gsl_grouplasso_adamw_step(float* params, const float* grads, float* m, float* v, const int32_t* group_offsets, int num_groups, int64_t n,
                              float lr, float wd, float beta1, float beta2, float eps, float alpha, float grad_scale, int step,
                              float* group_norms, void* stream) {
    OptimArgs a;
    a.params = params; a.grads = grads; a.m = m; a.v = v; a.group_offsets = group_offsets; a.num_groups = num_groups; a.n = n;
    a.lr = lr; a.wd = wd; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.alpha = alpha; a.grad_scale = grad_scale; a.step = step;
    a.group_norms = group_norms;
    return grouplasso_adamw_step(a, ST(stream));
}
int gsl_grouplasso_adamw_step_dev(float* params, const float* grads, float* m, float* v, const int32_t* group_offsets, int num_groups, int64_t n,
                                  float wd, float beta1, float beta2, float eps, float alpha, float grad_scale, const void* state_dev,
                                  float* group_norms, void* stream) {
    if (state_dev == nullptr) { set_last_error("gsl_grouplasso_adamw_step_dev: state_dev is null"); return -1; }
    OptimArgs a;
    a.params = params; a.grads = grads; a.m = m; a.v = v; a.group_offsets = group_offsets; a.num_groups = num_groups; a.n = n;
    a.lr = 0.f; a.wd = wd; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.alpha = alpha; a.grad_scale = grad_scale; a.step = 0;
    a.state = (const StepState*)state_dev;
    a.group_norms = group_norms;
    return grouplasso_adamw_step(a, ST(stream));
}
void gsl_count_launches(long long n) { g_launches.fetch_add(n); }
int gsl_tensor_norms(const float* params, const int32_t* tensor_offsets, int num_tensors, int type, float* out, void* stream) {
    return tensor_norms(params, tensor_offsets, num_tensors, type, out, ST(stream));
}

size_t gsl_engine_workspace_bytes(const GslConfig* cfg) { return Engine::workspace_bytes(*cfg); }
int gsl_engine_create(const GslConfig* cfg, void* workspace, size_t workspace_bytes, void** handle_out) {
    Engine* e = new Engine();
    int rc = e->init(*cfg, workspace, workspace_bytes);
    if (rc) { delete e; *handle_out = nullptr; return rc; }
    *handle_out = e;
    return 0;
}
void gsl_engine_destroy(void* handle) { delete (Engine*)handle; }
int gsl_engine_bind_params(void* handle, const void* const* frozen_ptrs, int num_ptrs, float* lora_flat, float* grad_flat) {
    return ((Engine*)handle)->bind_params(frozen_ptrs, num_ptrs, lora_flat, grad_flat);
}
int gsl_engine_refresh_frozen(void* handle, void* stream) { return ((Engine*)handle)->refresh_frozen(ST(stream)); }
int gsl_engine_refresh_lora(void* handle, void* stream) { return ((Engine*)handle)->refresh_lora(ST(stream)); }
int gsl_engine_forward(void* handle, int slot, const float* img, const int64_t* labels, int B, int use_lora, uint64_t dropout_seed,
                       void* stream) {
    return ((Engine*)handle)->forward(slot, img, 0, nullptr, nullptr, labels, B, use_lora, dropout_seed, ST(stream));
}
int gsl_engine_forward_u8(void* handle, int slot, const uint8_t* img, int layout, const float* mean, const float* std, const int64_t* labels,
                          int B, int use_lora, uint64_t dropout_seed, void* stream) {
    return ((Engine*)handle)->forward(slot, img, 1 + (layout != 0), mean, std, labels, B, use_lora, dropout_seed, ST(stream));
}
int gsl_engine_forward_dev(void* handle, int slot, const void* img, int img_is_u8_layout, const int64_t* labels, int B, int use_lora, int dropout_on,
                           const void* seed_dev, void* stream) {
    // img_is_u8_layout: 0 = fp32 NCHW, 1 = uint8 NCHW, 2 = uint8 NHWC (no Normalize on this entry point)
    return ((Engine*)handle)->forward(slot, img, img_is_u8_layout, nullptr, nullptr, labels, B, use_lora, dropout_on ? 1ull : 0ull, ST(stream),
                                      (const unsigned long long*)seed_dev);
}
int gsl_engine_backward(void* handle, int slot, const float* dlogits, const float* demb, int accumulate, void* stream) {
    return ((Engine*)handle)->backward(slot, dlogits, demb, accumulate, ST(stream));
}
void* gsl_engine_slot_ptr(void* handle, int slot, int what) {
    Engine* e = (Engine*)handle;
    if (slot < 0 || slot >= (int)e->slots.size()) return nullptr;
    Slot& s = e->slots[slot];
    switch (what) {
        case GSL_SLOT_EMB: return s.emb;
        case GSL_SLOT_LOGITS: return s.logits;
        case GSL_SLOT_CE: return s.ce;
        case GSL_SLOT_CORRECT: return s.correct;
        case GSL_SLOT_XFINAL: return s.cls.xout32;
        default: return nullptr;
    }
}
int64_t gsl_engine_lora_offset(void* handle, int block, int which) { return ((Engine*)handle)->lora_offset(block, which); }
int64_t gsl_engine_lora_numel(void* handle) { Engine* e = (Engine*)handle; return e->cfg.depth * e->lora_block_elems(); }
int gsl_loss_sums(const float* ce, const int32_t* correct, const float* kl, int n_remain, int B, float* sums, void* stream) {
    return loss_sums(ce, correct, kl, n_remain, B, sums, ST(stream));
}
int gsl_prototype_kl_fwd(const float* emb, const int64_t* labels, const float* proto, int B, int D, float* kl, void* stream) {
    return prototype_kl_fwd(emb, labels, proto, B, D, kl, ST(stream));
}
int gsl_prototype_kl_grad(const float* emb, const int64_t* labels, const float* proto, const float* sums, int n_remain_local, int B, int D,
                          float w_f, float w_r, float BND_pro, float* demb, void* stream) {
    return prototype_kl_grad(emb, labels, proto, sums, n_remain_local, B, D, w_f, w_r, BND_pro, demb, ST(stream));
}
int gsl_class_sums(const float* emb, const int64_t* labels, int B, int D, int C, float* sums, float* counts, void* stream) {
    return class_sums(emb, labels, B, D, C, sums, counts, ST(stream));
}
int gsl_class_means(const float* sums, const float* counts, int C, int D, float* out, void* stream) {
    return class_means(sums, counts, C, D, out, ST(stream));
}
int gsl_unlearn_ce_grad(const float* logits, const int64_t* labels, const float* sums, int n_remain_local, int B, int C, float beta, float BND,
                        float* dlogits, void* stream) {
    return unlearn_ce_grad(logits, labels, sums, n_remain_local, B, C, beta, BND, dlogits, ST(stream));
}

}  // extern "C"
