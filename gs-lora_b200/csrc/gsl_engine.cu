// gslora-b200: engine -- orchestrates the sm_100a kernels into ViT_face.forward (vit_face.py:523-548) and the
// selective backward autograd builds for engine_cl.py:124 (gradients for lora_A / lora_B only; dX flows through
// the frozen layers from the head down to block 0's FFN, nothing below that has a trainable ancestor).
//
// Memory: the caller hands over ONE workspace (a torch uint8 tensor); carve() lays out
//   fp16 operand caches of the frozen weights and their transposes; the FFN caches hold W + s*B*A while LoRA is live
//   `num_slots` activation sets (what the backward needs: residual snapshots, LN stats, q/k/v, O, LSE,
//    LN2(x), mask*gelu'(h), G = Dropout(gelu(h)))
//   transient gradient buffers shared by all slots.
#include "gsl_engine.h"
#include "gsl_common.cuh"
#include <cuda_fp8.h>

#include <cstdlib>
#include <cstring>

namespace gsl {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline void set_b(GemmArgs& g, const WOp& w) { g.B = w.hi; g.B_lo = w.lo; g.B_lo8 = w.lo8; g.lo8_shift = w.lo8 ? GSL_LO8_SHIFT : 0; }
// frozen weight -> cached operand in the engine's precision mode (optionally transposed)
static int cast_w(const float* src, int64_t lds, const WOp& w, int64_t ldd, int64_t rows, int64_t cols, int transpose, cudaStream_t s) {
    return cast_f32_to_f16(src, lds, w.hi, ldd, rows, cols, w.lo8 ? ldexpf(1.0f, GSL_LO8_SHIFT) : 1.f, transpose, s, w.lo, w.lo8);
}

int64_t Engine::lora_block_elems() const {
    const int64_t r = cfg.lora_rank, D = cfg.dim, H = cfg.mlp_dim, inner = (int64_t)cfg.heads * 64;
    if (cfg.lora_pos == 1) return 3 * r * D + 3 * inner * r;
    return r * D + H * r + r * H + D * r;
}
int64_t Engine::lora_offset(int block, int which) const {
    const int64_t r = cfg.lora_rank, D = cfg.dim, H = cfg.mlp_dim;
    int64_t off = block * lora_block_elems();
    if (cfg.lora_pos == 1) return off + (which >= 1 ? 3 * r * D : 0);
    if (which >= 1) off += r * D;
    if (which >= 2) off += H * r;
    if (which >= 3) off += r * H;
    return off;
}

// One pass computes sizes (assign = false) or assigns pointers (assign = true); both walk the same sequence.
size_t Engine::carve(bool assign) {
    size_t off = 0;
    auto take = [&](size_t bytes) -> uint8_t* {
        off = align_up(off, 1024);
        uint8_t* p = assign ? ws + off : nullptr;
        off += bytes;
        return p;
    };
    const int64_t D = cfg.dim, H = cfg.mlp_dim, L = cfg.depth, C = cfg.num_class, Bm = cfg.max_batch;
    const int64_t inner = (int64_t)cfg.heads * 64;
    const int64_t M = Bm * tokens;
    auto take_w = [&](size_t elems, int precision) -> WOp {        // one cached B operand: hi [, lo | lo8]
        WOp w;
        w.hi = (__half*)take(elems * 2);
        w.lo = precision == 1 ? (__half*)take(elems * 2) : nullptr;
        w.lo8 = precision == 2 ? (uint8_t*)take(elems) : nullptr;
        return w;
    };
    const int wprec = cfg.precision;
    const WOp pw = take_w((size_t)D * patch_dim, wprec == 2 ? 1 : wprec);   // patch embedding: K = patch_dim may not be a multiple of 64 -> fp16 residual
    auto* pb = (float*)take((size_t)tokens * D * 4);
    if (assign) { patch_w16 = pw; posb = pb; cache.assign(L, BlockCache()); }
    for (int l = 0; l < L; ++l) {
        BlockCache c;
        c.qkv_w16 = take_w((size_t)3 * inner * D, wprec);
        c.qkv_wT16 = take_w((size_t)D * 3 * inner, wprec);
        c.out_w16 = take_w((size_t)D * inner, wprec);
        c.out_wT16 = take_w((size_t)inner * D, wprec);
        c.fc1_w16 = take_w((size_t)H * D, wprec == 2 ? 1 : wprec);     // split8: the fc1 GEMM is epilogue-bound (GELU, two outputs) and measured 4 % faster
                                                                       // with the fp16 residual (no converter warps, 112 instead of 96 registers per thread)
        c.fc1T_w16 = take_w((size_t)D * H, wprec);
        c.fc2_w16 = take_w((size_t)D * H, wprec);
        c.fc2T_w16 = take_w((size_t)H * D, wprec);
        c.A1h = (__half*)take((size_t)32 * D * 2);
        c.A2h = (__half*)take((size_t)32 * H * 2);
        c.B1T = (__half*)take((size_t)32 * H * 2);
        c.B2T = (__half*)take((size_t)32 * D * 2);
        c.Aqkv = cfg.lora_pos == 1 ? (__half*)take((size_t)3 * 32 * D * 2) : nullptr;
        c.BqkvT = cfg.lora_pos == 1 ? (__half*)take((size_t)3 * 32 * inner * 2) : nullptr;
        if (assign) cache[l] = c;
    }
    if (assign) slots.assign(cfg.num_slots, Slot());
    for (int s = 0; s < cfg.num_slots; ++s) {
        Slot sl;
        for (int i = 0; i < 2 * L + 1; ++i) sl.x.push_back((float*)take((size_t)M * D * 4));
        for (int l = 0; l < L; ++l) {
            BlockActs a;
            a.ln1_mean = (float*)take((size_t)M * 4); a.ln1_rstd = (float*)take((size_t)M * 4);
            a.ln2_mean = (float*)take((size_t)M * 4); a.ln2_rstd = (float*)take((size_t)M * 4);
            a.lse = (float*)take((size_t)Bm * cfg.heads * tokens * 4);
            a.qkv16 = (__half*)take((size_t)M * 3 * inner * 2);
            a.o16 = (__half*)take((size_t)M * inner * 2);
            a.xn1_16 = cfg.lora_pos == 1 ? (__half*)take((size_t)M * D * 2) : nullptr;
            a.xn2_16 = (__half*)take((size_t)M * D * 2);
            a.gp16 = (__half*)take((size_t)M * H * 2);
            a.g16 = (__half*)take((size_t)M * H * 2);
            sl.blk.push_back(a);
        }
        sl.cls.xin32 = (float*)take((size_t)Bm * D * 4); sl.cls.xmid32 = (float*)take((size_t)Bm * D * 4); sl.cls.xout32 = (float*)take((size_t)Bm * D * 4);
        sl.cls.ln2_mean = (float*)take((size_t)Bm * 4); sl.cls.ln2_rstd = (float*)take((size_t)Bm * 4); sl.cls.lse = (float*)take((size_t)Bm * cfg.heads * 4);
        sl.cls.o16 = (__half*)take((size_t)Bm * inner * 2); sl.cls.xn2_16 = (__half*)take((size_t)Bm * D * 2);
        sl.cls.gp16 = (__half*)take((size_t)Bm * H * 2); sl.cls.g16 = (__half*)take((size_t)Bm * H * 2);
        sl.emb = (float*)take((size_t)Bm * D * 4);
        sl.logits = (float*)take((size_t)Bm * C * 4);
        sl.ce = (float*)take((size_t)Bm * 4);
        sl.xhat = (float*)take((size_t)Bm * D * 4);
        sl.head_rstd = (float*)take((size_t)Bm * 4);
        sl.correct = (int*)take((size_t)Bm * 4);
        if (assign) slots[s] = sl;
    }
    auto* t_patches = (__half*)take((size_t)M * patch_dim * 2);
    auto* t_xn = (__half*)take((size_t)M * D * 2);
    auto* t_dy = (__half*)take((size_t)M * D * 2);
    auto* t_dh = (__half*)take((size_t)M * H * 2);
    __half* t_tu[4];
    for (int i = 0; i < 4; ++i) t_tu[i] = (__half*)take((size_t)M * 16 * 2);
    auto* t_do = (__half*)take((size_t)M * inner * 2);
    auto* t_dqkv = (__half*)take((size_t)M * 3 * inner * 2);
    auto* t_delta = (float*)take((size_t)2 * Bm * cfg.heads * tokens * 4);
    auto* t_dx32 = (float*)take((size_t)M * D * 4);
    auto* t_dxn32 = (float*)take((size_t)M * D * 4);
    auto* t_cdx = (float*)take((size_t)Bm * D * 4);
    auto* t_cdxn = (float*)take((size_t)Bm * D * 4);
    auto* t_cdy = (__half*)take((size_t)Bm * D * 2);
    auto* t_cdh = (__half*)take((size_t)Bm * H * 2);
    auto* t_cdo = (__half*)take((size_t)Bm * inner * 2);
    size_t sk = skinny_tn_workspace(M, (int)(H > D ? H : D), cfg.lora_rank);
    const size_t sk_side = lora_side_workspace(M, (int)H, cfg.lora_rank);
    if (sk_side > sk) sk = sk_side;
    auto* t_sk = (float*)take(sk);
    auto* t_go = (int*)take((size_t)(L + 1) * 4);
    auto* t_to = (int*)take((size_t)(4 * L + 1) * 4);      // (4 tensors per block with lora_pos 0, 2 with lora_pos 1)
    auto* t_gn = (float*)take((size_t)L * 4);
    auto* t_tn = (float*)take((size_t)4 * L * 4);
    auto* t_pp = (void*)take((size_t)L * 8 * sizeof(void*));
    auto* t_mj = (void*)take((size_t)L * 3 * 128);
    if (assign) {
        patches16 = t_patches; xn16 = t_xn; dy16 = t_dy; dh16 = t_dh; do16 = t_do; dqkv16 = t_dqkv; attn_delta = t_delta;
        t1_16 = t_tu[0]; t2_16 = t_tu[1]; u1_16 = t_tu[2]; u2_16 = t_tu[3];
        dx32 = t_dx32; dxn32 = t_dxn32; skinny_ws = t_sk; skinny_ws_bytes = sk;
        cls_dx32 = t_cdx; cls_dxn32 = t_cdxn; cls_dy16 = t_cdy; cls_dh16 = t_cdh; cls_do16 = t_cdo;
        group_offsets_dev = t_go; tensor_offsets_dev = t_to; group_norms_dev = t_gn; tensor_norms_dev = t_tn; pack_ptrs_dev = t_pp;
        merge_jobs_dev = t_mj;
    }
    return align_up(off, 1024);
}

static int validate(const GslConfig& c) {
    GSL_REQUIRE(c.image_size % c.patch_size == 0, "image_size %% patch_size != 0");
    GSL_REQUIRE(c.dim % 128 == 0 && c.dim <= 1024, "dim=%d must be a multiple of 128 and <= 1024", c.dim);
    GSL_REQUIRE(c.mlp_dim % 64 == 0, "mlp_dim=%d must be a multiple of 64", c.mlp_dim);
    GSL_REQUIRE(c.heads * 64 == c.dim || c.heads > 0, "bad heads");
    GSL_REQUIRE(c.lora_rank >= 1 && c.lora_rank <= 16, "lora_rank=%d: the rank-r side kernels hold r <= 16 (args.py --lora_rank)", c.lora_rank);
    GSL_REQUIRE(c.lora_pos == 0 || c.lora_pos == 1, "lora_pos=%d: 0 (FFN) or 1 (Attention)", c.lora_pos);
    GSL_REQUIRE(c.precision >= 0 && c.precision <= 2, "precision=%d: 0 (fast: fp16 weights), 1 (split: fp16 hi + lo weights) or 2 (split8: fp16 hi + e4m3 lo)", c.precision);
    GSL_REQUIRE(c.precision != 2 || c.mlp_dim % 64 == 0, "precision split8 needs every GEMM K (dim, mlp_dim, 3 * heads * 64) to be a multiple of 64");
    GSL_REQUIRE((c.channels * c.patch_size * c.patch_size) % 16 == 0, "patch_dim must be a multiple of 16");
    GSL_REQUIRE(c.max_batch >= 1 && c.num_slots >= 1 && c.depth >= 1, "bad max_batch / num_slots / depth");
    const int tokens = (c.image_size / c.patch_size) * (c.image_size / c.patch_size) + 1;
    GSL_REQUIRE(tokens <= 208, "tokens=%d > 208 not supported by the attention kernel", tokens);
    GSL_REQUIRE(c.num_class <= 1024, "num_class=%d > 1024", c.num_class);
    return 0;
}

size_t Engine::workspace_bytes(const GslConfig& c) {
    if (validate(c) != 0) return 0;
    Engine e;
    e.cfg = c;
    e.tokens = (c.image_size / c.patch_size) * (c.image_size / c.patch_size) + 1;
    e.patch_dim = c.channels * c.patch_size * c.patch_size;
    return e.carve(false);
}

int Engine::init(const GslConfig& c, void* workspace, size_t bytes) {
    int rc = validate(c);
    if (rc) return rc;
    cfg = c;
    tokens = (c.image_size / c.patch_size) * (c.image_size / c.patch_size) + 1;
    patch_dim = c.channels * c.patch_size * c.patch_size;
    M_max = c.max_batch * tokens;
    const size_t need = carve(false);
    GSL_REQUIRE(bytes >= need, "workspace too small: %zu < %zu bytes", bytes, need);
    GSL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "workspace must be 1024-byte aligned");
    ws = (uint8_t*)workspace; ws_bytes = bytes;
    carve(true);
    // offsets of groups / tensors in the flat LoRA buffer
    const int tpb = lora_tensors_per_block();
    std::vector<int> go(c.depth + 1), to(tpb * c.depth + 1);
    for (int l = 0; l <= c.depth; ++l) go[l] = (int)(l * lora_block_elems());
    for (int l = 0; l < c.depth; ++l)
        for (int w = 0; w < tpb; ++w) to[tpb * l + w] = (int)lora_offset(l, w);
    to[tpb * c.depth] = (int)(c.depth * lora_block_elems());
    GSL_CHECK_CUDA(cudaMemcpy(group_offsets_dev, go.data(), go.size() * 4, cudaMemcpyHostToDevice));
    GSL_CHECK_CUDA(cudaMemcpy(tensor_offsets_dev, to.data(), to.size() * 4, cudaMemcpyHostToDevice));
    std::vector<void*> pp;
    for (int l = 0; l < c.depth; ++l) {
        const BlockCache& bc = cache[l];
        void* row[8] = {bc.A1h, bc.A2h, bc.B1T, bc.B2T, bc.Aqkv, bc.BqkvT, nullptr, nullptr};
        pp.insert(pp.end(), row, row + 8);
    }
    GSL_CHECK_CUDA(cudaMemcpy(pack_ptrs_dev, pp.data(), pp.size() * sizeof(void*), cudaMemcpyHostToDevice));
    return 0;
}

// out = fp16(W + sc * B A) and its transpose, W [R, C] fp32, A [r, C], B [R, r] (loralib.Linear's merged weight, layers.py train/eval);
// sc = 0 gives the plain fp16 cast.  One 32 x 32 tile per CTA, the transpose goes through shared memory.  Split mode: out_lo / outT_lo
// receive fp16(v - fp16(v)), the second term of the split operand (gsl_gemm.cu SPLIT).
// A / B null = no LoRA on this weight (plain cast).  groups > 1: loralib.MergedLinear -- the R rows are `groups` equal slices, slice g with its own
// pair A_g = A[g r : (g + 1) r, :], B_g = B[rows of slice g, :]  (merge_AB's grouped 1x1 conv, loralib layers.py).
struct MergeJob {
    const float *W, *A, *B;
    __half *out, *outT, *out_lo, *outT_lo;
    uint8_t *out_lo8, *outT_lo8;        // split8: e4m3 residual of the 2^shift-scaled value
    int R, C, groups;
    float s_out, s_outT;                // power-of-two pre-scale of each output (2^GSL_LO8_SHIFT where that output carries an e4m3 residual, else 1)
};
static constexpr int MERGE_JOB_SLOT = 128;
__global__ void __launch_bounds__(256) merge_weights_kernel(const uint8_t* __restrict__ jobs_raw, int r, float sc) {
    pdl_prologue();
    __shared__ float tile[32][33];
    const MergeJob j = *reinterpret_cast<const MergeJob*>(jobs_raw + (size_t)blockIdx.z * MERGE_JOB_SLOT);
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    if (c0 >= j.C || r0 >= j.R) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = r0 + ty + 8 * i, col = c0 + tx;
        float v = j.W[(int64_t)row * j.C + col];
        if (sc != 0.f && j.A != nullptr) {
            const int a0 = (row / (j.R / j.groups)) * r;        // first row of this slice's A_g
            float d = 0.f;
            for (int k = 0; k < r; ++k) d = fmaf(j.B[(int64_t)row * r + k], j.A[(int64_t)(a0 + k) * j.C + col], d);
            v = fmaf(sc, d, v);
        }
        tile[ty + 8 * i][tx] = v;
        v *= j.s_out;               // 1, or 2^shift where this output carries an e4m3 residual (exact)
        const __half hv = __float2half_rn(v);
        j.out[(int64_t)row * j.C + col] = hv;
        if (j.out_lo) j.out_lo[(int64_t)row * j.C + col] = __float2half_rn(v - __half2float(hv));
        if (j.out_lo8) j.out_lo8[(int64_t)row * j.C + col] = (uint8_t)__nv_cvt_float_to_fp8(v - __half2float(hv), __NV_SATFINITE, __NV_E4M3);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int col = c0 + ty + 8 * i, row = r0 + tx;
        const float v = tile[tx][ty + 8 * i] * j.s_outT;
        const __half hv = __float2half_rn(v);
        j.outT[(int64_t)col * j.R + row] = hv;
        if (j.outT_lo) j.outT_lo[(int64_t)col * j.R + row] = __float2half_rn(v - __half2float(hv));
        if (j.outT_lo8) j.outT_lo8[(int64_t)col * j.R + row] = (uint8_t)__nv_cvt_float_to_fp8(v - __half2float(hv), __NV_SATFINITE, __NV_E4M3);
    }
}

int Engine::ensure_ffn_weights(int use_lora, cudaStream_t s) {
    const int mode = use_lora ? 1 : 0;
    if (ffn_cache_mode == mode) return 0;
    const int D = cfg.dim, H = cfg.mlp_dim, inner3 = 3 * cfg.heads * 64;
    int big = H > D ? H : D;
    if (cfg.lora_pos == 1 && inner3 > big) big = inner3;
    dim3 grid(big / 32, big / 32, (cfg.lora_pos == 1 ? 3 : 2) * cfg.depth);
    GSL_CHECK_CUDA(launch_pdl(merge_weights_kernel, dim3(grid), dim3(256), 0, s, (const uint8_t*)merge_jobs_dev, cfg.lora_rank, use_lora ? cfg.lora_scaling : 0.f));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    ffn_cache_mode = mode;
    return 0;
}

int Engine::bind_params(const void* const* p, int n, float* lora, float* grads) {
    GSL_REQUIRE(n == GSL_NUM_GLOBAL_PARAMS + GSL_NUM_BLOCK_PARAMS * cfg.depth, "expected %d parameter pointers, got %d",
                GSL_NUM_GLOBAL_PARAMS + GSL_NUM_BLOCK_PARAMS * cfg.depth, n);
    pos_embedding = (const float*)p[0]; cls_token = (const float*)p[1]; patch_w = (const float*)p[2]; patch_b = (const float*)p[3];
    head_ln_w = (const float*)p[4]; head_ln_b = (const float*)p[5]; loss_w = (const float*)p[6]; head_b = (const float*)p[7];
    frozen.assign(cfg.depth, BlockFrozen());
    for (int l = 0; l < cfg.depth; ++l) {
        const void* const* q = p + GSL_NUM_GLOBAL_PARAMS + GSL_NUM_BLOCK_PARAMS * l;
        BlockFrozen& f = frozen[l];
        f.ln1_w = (const float*)q[0]; f.ln1_b = (const float*)q[1]; f.qkv_w = (const float*)q[2]; f.qkv_b = (const float*)q[3];
        f.out_w = (const float*)q[4]; f.out_b = (const float*)q[5]; f.ln2_w = (const float*)q[6]; f.ln2_b = (const float*)q[7];
        f.fc1_w = (const float*)q[8]; f.fc1_b = (const float*)q[9]; f.fc2_w = (const float*)q[10]; f.fc2_b = (const float*)q[11];
        GSL_REQUIRE(f.ln1_w && f.ln1_b && f.qkv_w && f.out_w && f.out_b && f.ln2_w && f.ln2_b && f.fc1_w && f.fc1_b && f.fc2_w && f.fc2_b,
                    "null frozen parameter in block %d", l);
    }
    GSL_REQUIRE(pos_embedding && cls_token && patch_w && patch_b && head_ln_w && head_ln_b, "null global parameter");
    GSL_REQUIRE(lora != nullptr, "lora_flat is null");
    lora_flat = lora; grad_flat = grads;
    params_bound = true;
    ffn_cache_mode = -1;
    // merge jobs: per block  fc1 (W1 [H, D], A1, B1)  and  fc2 (W2 [D, H], A2, B2); lora_pos 1: to_qkv (3 slices) with LoRA, fc1 / fc2 plain
    std::vector<MergeJob> jobs;
    const bool attn_lora = cfg.lora_pos == 1;
    for (int l = 0; l < cfg.depth; ++l) {
        MergeJob j1, j2;
        j1.groups = j2.groups = 1;
        if (attn_lora) {
            MergeJob jq;
            jq.W = frozen[l].qkv_w; jq.A = lora_flat + lora_offset(l, 0); jq.B = lora_flat + lora_offset(l, 1);
            jq.out = cache[l].qkv_w16.hi; jq.outT = cache[l].qkv_wT16.hi; jq.out_lo = cache[l].qkv_w16.lo; jq.outT_lo = cache[l].qkv_wT16.lo;
            jq.out_lo8 = cache[l].qkv_w16.lo8; jq.outT_lo8 = cache[l].qkv_wT16.lo8;
            jq.R = 3 * cfg.heads * 64; jq.C = cfg.dim; jq.groups = 3;
            jobs.push_back(jq);
        }
        j1.W = frozen[l].fc1_w; j1.A = attn_lora ? nullptr : lora_flat + lora_offset(l, 0); j1.B = attn_lora ? nullptr : lora_flat + lora_offset(l, 1);
        j1.out = cache[l].fc1_w16.hi; j1.outT = cache[l].fc1T_w16.hi; j1.out_lo = cache[l].fc1_w16.lo; j1.outT_lo = cache[l].fc1T_w16.lo;
        j1.out_lo8 = cache[l].fc1_w16.lo8; j1.outT_lo8 = cache[l].fc1T_w16.lo8;
        j1.R = cfg.mlp_dim; j1.C = cfg.dim;
        j2.W = frozen[l].fc2_w; j2.A = attn_lora ? nullptr : lora_flat + lora_offset(l, 2); j2.B = attn_lora ? nullptr : lora_flat + lora_offset(l, 3);
        j2.out = cache[l].fc2_w16.hi; j2.outT = cache[l].fc2T_w16.hi; j2.out_lo = cache[l].fc2_w16.lo; j2.outT_lo = cache[l].fc2T_w16.lo;
        j2.out_lo8 = cache[l].fc2_w16.lo8; j2.outT_lo8 = cache[l].fc2T_w16.lo8;
        j2.R = cfg.dim; j2.C = cfg.mlp_dim;
        jobs.push_back(j1); jobs.push_back(j2);
    }
    for (MergeJob& j : jobs) { j.s_out = j.out_lo8 ? ldexpf(1.0f, GSL_LO8_SHIFT) : 1.0f; j.s_outT = j.outT_lo8 ? ldexpf(1.0f, GSL_LO8_SHIFT) : 1.0f; }
    static_assert(sizeof(MergeJob) <= MERGE_JOB_SLOT, "MergeJob slot");
    std::vector<uint8_t> raw(jobs.size() * MERGE_JOB_SLOT, 0);
    for (size_t i = 0; i < jobs.size(); ++i) memcpy(raw.data() + i * MERGE_JOB_SLOT, &jobs[i], sizeof(MergeJob));
    GSL_CHECK_CUDA(cudaMemcpy(merge_jobs_dev, raw.data(), raw.size(), cudaMemcpyHostToDevice));
    return 0;
}

// posb[n, :] = pos[n] + (n == 0 ? cls : patch_bias)   (vit_face.py:531-536 folded into the patch-embed epilogue)
__global__ void posb_kernel(const float* __restrict__ pos, const float* __restrict__ cls, const float* __restrict__ pbias, float* __restrict__ out,
                            int tokens, int D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= tokens * D) return;
    const int n = i / D, d = i % D;
    out[i] = pos[i] + (n == 0 ? cls[d] : pbias[d]);
}

int Engine::refresh_frozen(cudaStream_t s) {
    GSL_REQUIRE(params_bound, "bind_params first");
    const int D = cfg.dim, H = cfg.mlp_dim, inner = cfg.heads * 64;
    int rc;
    if ((rc = cast_w(patch_w, patch_dim, patch_w16, patch_dim, D, patch_dim, 0, s))) return rc;
    posb_kernel<<<(tokens * D + 255) / 256, 256, 0, s>>>(pos_embedding, cls_token, patch_b, posb, tokens, D);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    for (int l = 0; l < cfg.depth; ++l) {
        const BlockFrozen& f = frozen[l];
        BlockCache& c = cache[l];
        if ((rc = cast_w(f.qkv_w, D, c.qkv_w16, D, 3 * inner, D, 0, s))) return rc;
        if ((rc = cast_w(f.qkv_w, D, c.qkv_wT16, 3 * inner, 3 * inner, D, 1, s))) return rc;
        if ((rc = cast_w(f.out_w, inner, c.out_w16, inner, D, inner, 0, s))) return rc;
        if ((rc = cast_w(f.out_w, inner, c.out_wT16, D, D, inner, 1, s))) return rc;
        if ((rc = fill_zero(c.A1h, (size_t)32 * D * 2, s))) return rc;
        if ((rc = fill_zero(c.A2h, (size_t)32 * H * 2, s))) return rc;
        if ((rc = fill_zero(c.B1T, (size_t)32 * H * 2, s))) return rc;
        if ((rc = fill_zero(c.B2T, (size_t)32 * D * 2, s))) return rc;
        if (c.Aqkv && (rc = fill_zero(c.Aqkv, (size_t)3 * 32 * D * 2, s))) return rc;
        if (c.BqkvT && (rc = fill_zero(c.BqkvT, (size_t)3 * 32 * inner * 2, s))) return rc;
    }
    ffn_cache_mode = -1;        // the FFN caches are rebuilt (with or without the LoRA delta) by the next forward
    return refresh_lora(s);
}

// One launch repacks the fp16 LoRA operands of every block from the flat fp32 parameter buffer: A1h/A2h = lora_A, B1T/B2T = lora_B^T
// (the rank-r by-products T = x A^T, U = dY B of the backward).  The FFN weight caches are re-merged lazily by the next forward.
struct LoraPackPtrs {
    __half *A1h, *A2h, *B1T, *B2T, *Aqkv, *BqkvT;
    void* unused[2];
};
// lora_pos 1: flat block = [ to_qkv.lora_A (3r x D) | to_qkv.lora_B (3 inner x r) ] -> Aqkv[g] = A_g, BqkvT[g] = B_g^T
__global__ void lora_pack_attn_kernel(const float* __restrict__ flat, const LoraPackPtrs* __restrict__ ptrs, int D, int inner, int r, int per_block, int split) {
    const int l = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_block) return;
    const LoraPackPtrs p = ptrs[l];
    const float v = flat[(int64_t)l * per_block + i];
    const __half hv = __float2half_rn(v);
    const __half lv = __float2half_rn(v - __half2float(hv));
    __half* dst; int64_t off, lo_off;
    if (i < 3 * r * D) {
        const int g = i / (r * D), k = i % (r * D);
        dst = p.Aqkv + (int64_t)g * 32 * D; off = k; lo_off = (int64_t)16 * D;
    } else {
        const int k = i - 3 * r * D, row = k / r, j = k % r, g = row / inner, o = row % inner;
        dst = p.BqkvT + (int64_t)g * 32 * inner; off = (int64_t)j * inner + o; lo_off = (int64_t)16 * inner;
    }
    dst[off] = hv;
    if (split) dst[off + lo_off] = lv;
}
__global__ void lora_pack_kernel(const float* __restrict__ flat, const LoraPackPtrs* __restrict__ ptrs, int D, int H, int r, int per_block, int split) {
    pdl_prologue();
    const int l = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_block) return;
    const LoraPackPtrs p = ptrs[l];
    const float v = flat[(int64_t)l * per_block + i];
    const __half hv = __float2half_rn(v);
    const __half lv = __float2half_rn(v - __half2float(hv));      // split mode: residual rows [16, 32) of each operand
    int k = i;
    __half* dst; int64_t off, lo_off;
    if (k < r * D) { dst = p.A1h; off = k; lo_off = (int64_t)16 * D; }                                              // lora_A(net.0) [r, D]
    else if ((k -= r * D) < H * r) { dst = p.B1T; off = (int64_t)(k % r) * H + k / r; lo_off = (int64_t)16 * H; }    // lora_B(net.0) [H, r]
    else if ((k -= H * r) < r * H) { dst = p.A2h; off = k; lo_off = (int64_t)16 * H; }                              // lora_A(net.3) [r, H]
    else { k -= r * H; dst = p.B2T; off = (int64_t)(k % r) * D + k / r; lo_off = (int64_t)16 * D; }                  // lora_B(net.3) [D, r]
    dst[off] = hv;
    if (split) dst[off + lo_off] = lv;
}

int Engine::refresh_lora(cudaStream_t s) {
    GSL_REQUIRE(params_bound, "bind_params first");
    const int per_block = (int)lora_block_elems();
    dim3 grid((per_block + 255) / 256, cfg.depth);
    if (cfg.lora_pos == 1)
        lora_pack_attn_kernel<<<grid, 256, 0, s>>>(lora_flat, (const LoraPackPtrs*)pack_ptrs_dev, cfg.dim, cfg.heads * 64, cfg.lora_rank, per_block, cfg.precision >= 1 ? 1 : 0);
    else
        GSL_CHECK_CUDA(launch_pdl(lora_pack_kernel, dim3(grid), dim3(256), 0, s, lora_flat, (const LoraPackPtrs*)pack_ptrs_dev, cfg.dim, cfg.mlp_dim, cfg.lora_rank, per_block, cfg.precision >= 1 ? 1 : 0));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    if (ffn_cache_mode == 1) ffn_cache_mode = -1;       // W + s B A is stale
    return 0;
}

// experiment switch (dev): GSLORA_DXN32=1 keeps the dLN GEMM outputs in fp32 instead of fp16
static bool dxn_fp32() {
    static const bool on = [] { const char* e = getenv("GSLORA_DXN32"); return e && e[0] == '1'; }();
    return on;
}

// per-site dropout seeds: site 0 emb (block index = depth), 1 attention to_out, 2 after GELU, 3 after fc2
static inline uint32_t site_seed(uint64_t base, int block, int site) {
    return drop_hash((uint32_t)(block * 4 + site + 1), (uint32_t)base ^ (uint32_t)(base >> 32));
}

DropSeed Engine::site(uint64_t base, int block, int site_id) const {
    DropSeed d(site_seed(base, block, site_id));
    if (seed_dev_cur) { d.dev = seed_dev_cur; d.key = (uint32_t)(block * 4 + site_id + 1); }      // same derivation, evaluated on the device at run time
    return d;
}

// x_out = FeedForward(LN2(x_mid)) + x_mid on M rows (dense tokens, or the compacted cls rows of the last block).  The LoRA branches
// of both lora.Linear layers are inside the cached weights (ensure_ffn_weights), so this is LN -> GEMM(+GELU) -> GEMM(+residual).
int Engine::ffn_forward(int l, int64_t M, __half* xn2, float* ln_mean, float* ln_rstd, const float* x_mid, __half* gp16, __half* g16, float* x_out,
                        float pdrop, uint64_t dseed, cudaStream_t s) {
    const int D = cfg.dim, H = cfg.mlp_dim;
    const BlockFrozen& f = frozen[l];
    const BlockCache& c = cache[l];
    int rc;
    if ((rc = layernorm_fwd(x_mid, D, f.ln2_w, f.ln2_b, cfg.ln_eps, xn2, D, ln_mean, ln_rstd, M, D, s))) return rc;
    {
        GemmArgs g;
        g.A = xn2; g.lda = D; set_b(g, c.fc1_w16); g.ldb = D; g.M = M; g.N = H; g.K = D;
        g.epi = EPI_GELU; g.bias = f.fc1_b; g.out0 = gp16; g.ld0 = H; g.out1 = g16; g.ld1 = H;
        g.drop_p = pdrop; g.drop_seed = site(dseed, l, 2);
        if ((rc = gemm_f16(g, s))) return rc;
    }
    {
        GemmArgs g;
        g.A = g16; g.lda = H; set_b(g, c.fc2_w16); g.ldb = H; g.M = M; g.N = D; g.K = H;
        g.epi = EPI_RES_F32; g.bias = f.fc2_b; g.out0 = x_out; g.ld0 = D; g.aux = x_mid; g.ldaux = D;
        g.drop_p = pdrop; g.drop_seed = site(dseed, l, 3);
        if ((rc = gemm_f16(g, s))) return rc;
    }
    return 0;
}

int Engine::forward(int slot, const void* img, int img_kind, const float* mean, const float* std, const int64_t* labels, int B, int use_lora,
                    uint64_t dropout_seed, cudaStream_t s, const unsigned long long* seed_dev) {
    GSL_REQUIRE(img_kind >= 0 && img_kind <= 2, "image kind %d (0 fp32 NCHW, 1 uint8 NCHW, 2 uint8 NHWC)", img_kind);
    GSL_REQUIRE(params_bound, "bind_params first");
    GSL_REQUIRE(slot >= 0 && slot < cfg.num_slots, "slot %d out of range", slot);
    GSL_REQUIRE(B >= 1 && B <= cfg.max_batch, "batch %d outside [1, %d]", B, cfg.max_batch);
    const int D = cfg.dim, L = cfg.depth, inner = cfg.heads * 64;
    const int64_t M = (int64_t)B * tokens;
    Slot& S = slots[slot];
    S.batch = B; S.used_lora = use_lora; S.drop_seed = dropout_seed; S.seed_dev = dropout_seed ? seed_dev : nullptr;
    seed_dev_cur = S.seed_dev;
    const float pdrop = dropout_seed ? cfg.dropout : 0.f, pemb = dropout_seed ? cfg.emb_dropout : 0.f;
    int rc;
    if ((rc = ensure_ffn_weights(use_lora, s))) return rc;
    if (img_kind == 0) rc = patchify_f16((const float*)img, patches16, patch_dim, B, cfg.channels, cfg.image_size, cfg.patch_size, cfg.patch_order, s);
    else rc = patchify_u8_f16((const uint8_t*)img, img_kind - 1, mean, std, patches16, patch_dim, B, cfg.channels, cfg.image_size, cfg.patch_size,
                              cfg.patch_order, s);
    if (rc) return rc;
    {   // patch_to_embedding + cls token + pos_embedding (vit_face.py:531-536)
        GemmArgs g;
        g.A = patches16; g.lda = patch_dim; set_b(g, patch_w16); g.ldb = patch_dim; g.M = M; g.N = D; g.K = patch_dim;
        g.epi = EPI_PERIODIC_F32; g.out0 = S.x[0]; g.ld0 = D; g.aux = posb; g.ldaux = D; g.aux_period = tokens;
        g.drop_p = pemb; g.drop_seed = site(dropout_seed, L, 0);
        if ((rc = gemm_f16(g, s))) return rc;
    }
    for (int l = 0; l < L; ++l) {
        const BlockFrozen& f = frozen[l];
        const BlockCache& c = cache[l];
        BlockActs& a = S.blk[l];
        float* x_in = S.x[2 * l];
        // ---- x = Attention(LN(x)) + x
        __half* xn1 = a.xn1_16 ? a.xn1_16 : xn16;        // kept for the backward only when to_qkv carries LoRA
        if ((rc = layernorm_fwd(x_in, D, f.ln1_w, f.ln1_b, cfg.ln_eps, xn1, D, a.ln1_mean, a.ln1_rstd, M, D, s))) return rc;
        {
            GemmArgs g;
            g.A = xn1; g.lda = D; set_b(g, c.qkv_w16); g.ldb = D; g.M = M; g.N = 3 * inner; g.K = D;
            g.epi = EPI_F16; g.bias = f.qkv_b; g.out0 = a.qkv16; g.ld0 = 3 * inner;
            if ((rc = gemm_f16(g, s))) return rc;
        }
        if (l < L - 1) {
            float* x_mid = S.x[2 * l + 1];
            float* x_out = S.x[2 * l + 2];
            if ((rc = attention_fwd(a.qkv16, 3 * inner, a.o16, inner, a.lse, B, tokens, cfg.heads, cfg.attn_scale, s))) return rc;
            GemmArgs g;
            g.A = a.o16; g.lda = inner; set_b(g, c.out_w16); g.ldb = inner; g.M = M; g.N = D; g.K = inner;
            g.epi = EPI_RES_F32; g.bias = f.out_b; g.out0 = x_mid; g.ld0 = D; g.aux = x_in; g.ldaux = D;
            g.drop_p = pdrop; g.drop_seed = site(dropout_seed, l, 1);
            if ((rc = gemm_f16(g, s))) return rc;
            // ---- x = FeedForward(LN(x)) + x, loralib.Linear on both projections
            if ((rc = ffn_forward(l, M, a.xn2_16, a.ln2_mean, a.ln2_rstd, x_mid, a.gp16, a.g16, x_out, pdrop, dropout_seed, s))) return rc;
        } else {
            // ---- last block: only the cls token is pooled (vit_face.py:540), every other token of this block is dead.
            //      Single-query attention per (image, head), then out-proj / FFN on the B compacted cls rows.
            ClsActs& k = S.cls;
            if ((rc = cls_attention_fwd(a.qkv16, 3 * inner, k.o16, inner, k.lse, B, tokens, cfg.heads, cfg.attn_scale, s))) return rc;
            if ((rc = copy_cls_rows(x_in, (int64_t)tokens * D * 4, k.xin32, (int64_t)D * 4, B, (int64_t)D * 4, s))) return rc;
            GemmArgs g;
            g.A = k.o16; g.lda = inner; set_b(g, c.out_w16); g.ldb = inner; g.M = B; g.N = D; g.K = inner;
            g.epi = EPI_RES_F32; g.bias = f.out_b; g.out0 = k.xmid32; g.ld0 = D; g.aux = k.xin32; g.ldaux = D;
            g.drop_p = pdrop; g.drop_seed = site(dropout_seed, l, 1);
            if ((rc = gemm_f16(g, s))) return rc;
            if ((rc = ffn_forward(l, B, k.xn2_16, k.ln2_mean, k.ln2_rstd, k.xmid32, k.gp16, k.g16, k.xout32, pdrop, dropout_seed, s))) return rc;
        }
    }
    HeadArgs h;
    h.x = S.cls.xout32; h.ldx = D; h.tokens = 1; h.gamma = head_ln_w; h.beta = head_ln_b; h.eps = cfg.ln_eps;
    h.head_type = cfg.head_type; h.head_b = head_b;
    h.W = (labels || cfg.head_type == 1) ? loss_w : nullptr; h.labels = labels; h.cos_s = cfg.cos_s; h.cos_m = cfg.cos_m; h.B = B; h.D = D; h.C = cfg.num_class;
    h.emb = S.emb; h.logits = S.logits; h.ce = S.ce; h.correct = S.correct; h.xhat = S.xhat; h.rstd = S.head_rstd;
    if (labels && cfg.head_type == 0) GSL_REQUIRE(loss_w != nullptr, "labelled forward needs loss.weight");
    return head_fwd(h, s);
}

// Backward of x_out = FeedForward(LN2(x_mid)) + x_mid on M rows.  In: dx (fp32) and dy (fp16, already carrying the fc2-output dropout
// mask) = gradient w.r.t. x_out.  Out: dA / dB of both LoRA layers; unless l == 0, dx / dy = gradient w.r.t. x_mid (the fp16 copy masked
// for the attention to_out dropout).  Closed forms: SURVEY Appendix C; the dX GEMMs use the merged weights (dY W' = dY W + s (dY B) A),
// the rank-r by-products T = x A^T, U = dY B exist only here and the two wide ones ride on the single pass that reads G / dH anyway.
int Engine::ffn_backward(int l, int64_t M, __half* dy, float* dx, __half* dh, float* dxn, const __half* xn2, const __half* gp16, const __half* g16,
                         const float* x_mid, const float* ln_mean, const float* ln_rstd, int accumulate, float pdrop, uint64_t dseed, cudaStream_t s) {
    const int D = cfg.dim, H = cfg.mlp_dim, r = cfg.lora_rank;
    const BlockFrozen& f = frozen[l];
    const BlockCache& c = cache[l];
    const float wscale = cfg.lora_scaling / cfg.grad_scale;     // dA, dB carry the LoRA scaling and undo the loss scale
    float* gA1 = grad_flat + lora_offset(l, 0);
    float* gB1 = grad_flat + lora_offset(l, 1);
    float* gA2 = grad_flat + lora_offset(l, 2);
    float* gB2 = grad_flat + lora_offset(l, 3);
    int rc;
    const int fold = cfg.precision >= 1 ? 1 : 0;                // split mode: the rank-r operands carry their rounding residual too
    const bool ffn_lora = cfg.lora_pos == 0;
    if (ffn_lora) {
    if ((rc = lora_down(dy, D, c.B2T, D, u2_16, 16, M, D, r, s, fold))) return rc;                                                        // U2 = dY2 B2
    if ((rc = lora_side(g16, H, c.A2h, H, t2_16, 16, u2_16, 16, gA2, H, 1, wscale, accumulate, M, H, r, skinny_ws, skinny_ws_bytes, s, fold))) return rc;   // T2 = G A2^T, dA2 = s U2^T G
    if ((rc = skinny_tn(dy, D, t2_16, 16, gB2, r, 0, wscale, accumulate, M, D, r, skinny_ws, skinny_ws_bytes, s))) return rc;       // dB2 = s dY2^T T2
    }
    {   // dH = (dY2 W2') * d[Dropout(gelu(h))] / dh
        GemmArgs g;
        g.A = dy; g.lda = D; set_b(g, c.fc2T_w16); g.ldb = D; g.M = M; g.N = H; g.K = D;
        g.epi = EPI_GELU_BWD; g.out0 = dh; g.ld0 = H; g.aux = gp16; g.ldaux = H;
        if ((rc = gemm_f16(g, s))) return rc;
    }
    if (ffn_lora) {
    if ((rc = lora_down(xn2, D, c.A1h, D, t1_16, 16, M, D, r, s, fold))) return rc;                                                       // T1 = LN2(x) A1^T
    if ((rc = lora_side(dh, H, c.B1T, H, u1_16, 16, t1_16, 16, gB1, r, 0, wscale, accumulate, M, H, r, skinny_ws, skinny_ws_bytes, s, fold))) return rc;    // U1 = dH B1, dB1 = s dH^T T1
    if ((rc = skinny_tn(xn2, D, u1_16, 16, gA1, D, 1, wscale, accumulate, M, D, r, skinny_ws, skinny_ws_bytes, s))) return rc;      // dA1 = s U1^T LN2(x)
    }
    if (l == 0 && ffn_lora) return 0;       // nothing trainable below block 0's FFN (with LoRA on to_qkv block 0's attention still is)
    {   // dLN2 = dH W1'
        GemmArgs g;
        g.A = dh; g.lda = H; set_b(g, c.fc1T_w16); g.ldb = H; g.M = M; g.N = D; g.K = H;
        g.epi = dxn_fp32() ? EPI_F32 : EPI_F16; g.out0 = dxn; g.ld0 = D;            // fp16: halves the traffic of the LayerNorm-backward pass that consumes it
        if ((rc = gemm_f16(g, s))) return rc;
    }
    return layernorm_bwd(dxn, dxn_fp32() ? 0 : 1, D, x_mid, D, ln_mean, ln_rstd, f.ln2_w, dx, D, dx, D, dy, D, M, D, pdrop, site(dseed, l, 1), s);
}

// LoRA on to_qkv (loralib.MergedLinear, one (A_g, B_g) pair per slice g = q, k, v; vit_face.py:349-355): with dQKV [M, 3 inner] from the attention
// backward and LN1(x) saved by the forward, per slice   U_g = dQKV_g B_g,  dA_g = s U_g^T LN1(x),  T_g = LN1(x) A_g^T,  dB_g = s dQKV_g^T T_g
// (SURVEY Appendix C applied to each slice; dLN1 = dQKV W'qkv already carries the LoRA term through the merged weight).
int Engine::attn_lora_grads(int l, int64_t M, const __half* dqkv, const __half* xn1, int accumulate, cudaStream_t s) {
    const int D = cfg.dim, inner = cfg.heads * 64, r = cfg.lora_rank;
    const BlockCache& c = cache[l];
    const float wscale = cfg.lora_scaling / cfg.grad_scale;
    const int fold = cfg.precision >= 1 ? 1 : 0;
    float* gA = grad_flat + lora_offset(l, 0);      // [3r, D]
    float* gB = grad_flat + lora_offset(l, 1);      // [3 inner, r]
    int rc;
    for (int g = 0; g < 3; ++g) {
        const __half* dq = dqkv + (int64_t)g * inner;
        if ((rc = lora_down(dq, 3 * inner, c.BqkvT + (int64_t)g * 32 * inner, inner, u1_16, 16, M, inner, r, s, fold))) return rc;
        if ((rc = skinny_tn(xn1, D, u1_16, 16, gA + (int64_t)g * r * D, D, 1, wscale, accumulate, M, D, r, skinny_ws, skinny_ws_bytes, s))) return rc;
        if ((rc = lora_down(xn1, D, c.Aqkv + (int64_t)g * 32 * D, D, t1_16, 16, M, D, r, s, fold))) return rc;
        if ((rc = skinny_tn(dq, 3 * inner, t1_16, 16, gB + (int64_t)g * inner * r, r, 0, wscale, accumulate, M, inner, r, skinny_ws, skinny_ws_bytes, s))) return rc;
    }
    return 0;
}

int Engine::backward(int slot, const float* dlogits, const float* demb, int accumulate, cudaStream_t s) {
    GSL_REQUIRE(params_bound && grad_flat != nullptr, "bind_params (with a gradient buffer) first");
    GSL_REQUIRE(slot >= 0 && slot < cfg.num_slots, "slot %d out of range", slot);
    Slot& S = slots[slot];
    GSL_REQUIRE(S.batch > 0, "slot %d holds no forward", slot);
    GSL_REQUIRE(S.used_lora, "backward needs a forward run with use_lora = 1 (train mode, unmerged)");
    GSL_REQUIRE(dlogits || demb, "backward needs d logits and/or d emb");
    const int B = S.batch, D = cfg.dim, L = cfg.depth, inner = cfg.heads * 64;
    const int64_t M = (int64_t)B * tokens;
    const uint64_t dseed = S.drop_seed;
    seed_dev_cur = S.seed_dev;
    const float pdrop = dseed ? cfg.dropout : 0.f;
    const bool attn_lora = cfg.lora_pos == 1;
    int rc;
    // ---------------- head: gradient of the B cls rows of the last block's output (scaled by the loss scale)
    HeadBwdArgs hb;
    hb.dlogits = dlogits; hb.demb = demb; hb.emb = S.emb; hb.W = loss_w; hb.labels = nullptr; hb.xhat = S.xhat; hb.rstd = S.head_rstd;
    hb.head_type = cfg.head_type;
    hb.gamma = head_ln_w; hb.cos_s = cfg.cos_s; hb.B = B; hb.D = D; hb.C = cfg.num_class; hb.tokens = 1; hb.gscale = cfg.grad_scale;
    hb.dx = cls_dx32; hb.lddx = D; hb.dx16 = cls_dy16; hb.lddx16 = D;
    hb.drop_p = pdrop; hb.drop_seed = site(dseed, L - 1, 3);
    if ((rc = head_bwd(hb, s))) return rc;
    {   // ---------------- last block on the compacted cls rows
        const int l = L - 1;
        const BlockFrozen& f = frozen[l];
        const BlockCache& c = cache[l];
        BlockActs& a = S.blk[l];
        ClsActs& k = S.cls;
        if ((rc = ffn_backward(l, B, cls_dy16, cls_dx32, cls_dh16, cls_dxn32, k.xn2_16, k.gp16, k.g16, k.xmid32, k.ln2_mean, k.ln2_rstd,
                               accumulate, pdrop, dseed, s))) return rc;
        if (l == 0 && !attn_lora) return 0;
        {   // dO (cls rows) = dY Wo
            GemmArgs g;
            g.A = cls_dy16; g.lda = D; set_b(g, c.out_wT16); g.ldb = D; g.M = B; g.N = inner; g.K = D;
            g.epi = EPI_F16; g.out0 = cls_do16; g.ld0 = inner;
            if ((rc = gemm_f16(g, s))) return rc;
        }
        if ((rc = cls_attention_bwd(a.qkv16, 3 * inner, k.o16, inner, cls_do16, inner, k.lse, dqkv16, 3 * inner, B, tokens, cfg.heads, cfg.attn_scale, s))) return rc;
        if (attn_lora && (rc = attn_lora_grads(l, M, dqkv16, a.xn1_16, accumulate, s))) return rc;
        if (l == 0) return 0;
        {   // dLN1 = dQKV Wqkv  (dense: dK / dV reach every token)
            GemmArgs g;
            g.A = dqkv16; g.lda = 3 * inner; set_b(g, c.qkv_wT16); g.ldb = 3 * inner; g.M = M; g.N = D; g.K = 3 * inner;
            g.epi = dxn_fp32() ? EPI_F32 : EPI_F16; g.out0 = dxn32; g.ld0 = D;
            if ((rc = gemm_f16(g, s))) return rc;
        }
        // residual gradient of this block's input: zero except the cls rows, which LayerNorm backward picks from the compacted cls_dx32
        if ((rc = layernorm_bwd(dxn32, dxn_fp32() ? 0 : 1, D, S.x[2 * l], D, a.ln1_mean, a.ln1_rstd, f.ln1_w, cls_dx32, D, dx32, D, dy16, D, M, D, pdrop,
                                site(dseed, l - 1, 3), s, tokens))) return rc;
    }
    for (int l = L - 2; l >= 0; --l) {
        const BlockFrozen& f = frozen[l];
        const BlockCache& c = cache[l];
        BlockActs& a = S.blk[l];
        // ---------------- FFN: y = fc2(gelu(fc1(LN2(x)))) + x
        if ((rc = ffn_backward(l, M, dy16, dx32, dh16, dxn32, a.xn2_16, a.gp16, a.g16, S.x[2 * l + 1], a.ln2_mean, a.ln2_rstd, accumulate,
                               pdrop, dseed, s))) return rc;
        if (l == 0 && !attn_lora) break;      // nothing trainable below block 0's FFN
        // ---------------- attention: y = to_out(attn(to_qkv(LN1(x)))) + x
        {   // dO = dY Wo; the epilogue also emits delta = rowsum(dO * O) per (image, head, token) for the attention backward
            GemmArgs g;
            g.A = dy16; g.lda = D; set_b(g, c.out_wT16); g.ldb = D; g.M = M; g.N = inner; g.K = D;
            g.epi = EPI_F16_ROWDOT; g.out0 = do16; g.ld0 = inner; g.aux = a.o16; g.ldaux = inner; g.aux_period = tokens; g.rowdot = attn_delta;
            if ((rc = gemm_f16(g, s))) return rc;
        }
        if ((rc = attention_bwd(a.qkv16, 3 * inner, a.o16, inner, do16, inner, a.lse, dqkv16, 3 * inner, B, tokens, cfg.heads, cfg.attn_scale, s,
                                attn_delta))) return rc;
        if (attn_lora && (rc = attn_lora_grads(l, M, dqkv16, a.xn1_16, accumulate, s))) return rc;
        if (l == 0) break;      // block 0: to_qkv's LoRA is the last trainable thing on the way down
        {   // dLN1 = dQKV Wqkv
            GemmArgs g;
            g.A = dqkv16; g.lda = 3 * inner; set_b(g, c.qkv_wT16); g.ldb = 3 * inner; g.M = M; g.N = D; g.K = 3 * inner;
            g.epi = dxn_fp32() ? EPI_F32 : EPI_F16; g.out0 = dxn32; g.ld0 = D;
            if ((rc = gemm_f16(g, s))) return rc;
        }
        if ((rc = layernorm_bwd(dxn32, dxn_fp32() ? 0 : 1, D, S.x[2 * l], D, a.ln1_mean, a.ln1_rstd, f.ln1_w, dx32, D, dx32, D, dy16, D, M, D, pdrop,
                                site(dseed, l - 1, 3), s))) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ step-loss helpers
__global__ void loss_sums_kernel(const float* __restrict__ ce, const int* __restrict__ correct, const float* __restrict__ kl, int n_remain, int B,
                                 float* __restrict__ sums) {
    pdl_prologue();
    __shared__ float red[8][32];
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        if (b < n_remain) { v[0] += ce[b]; v[1] += 1.f; v[4] += correct ? (float)correct[b] : 0.f; v[6] += kl ? kl[b] : 0.f; }
        else { v[2] += ce[b]; v[3] += 1.f; v[5] += correct ? (float)correct[b] : 0.f; v[7] += kl ? kl[b] : 0.f; }
    }
    for (int k = 0; k < 8; ++k) {
        v[k] = warp_sum(v[k]);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[threadIdx.x][i];
        sums[threadIdx.x] = t;
    }
}

// sums[0..7] = sum CE remain, n remain, sum CE forget, n forget, hits remain, hits forget, sum KL remain, sum KL forget (kl may be null)
int loss_sums(const float* ce, const int* correct, const float* kl, int n_remain, int B, float* sums, cudaStream_t s) {
    GSL_CHECK_CUDA(launch_pdl(loss_sums_kernel, dim3(1), dim3(1024), 0, s, ce, correct, kl, n_remain, B, sums));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// loss = CE_r + beta * relu(BND - CE_f)  (engine_cl.py:65-80,118-120): d/dlogits per sample
__global__ void unlearn_ce_grad_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ sums,
                                       int n_remain_local, int B, int C, float beta, float BND, float* __restrict__ dlogits) {
    pdl_prologue();
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    float w;
    if (b < n_remain_local) w = sums[1] > 0.f ? 1.0f / sums[1] : 0.f;
    else {
        const float mean_f = sums[3] > 0.f ? sums[2] / sums[3] : 0.f;
        w = (sums[3] > 0.f && mean_f < BND) ? -beta / sums[3] : 0.f;      // relu'(BND - CE_f) = [CE_f < BND]
    }
    const float* lr = logits + (int64_t)b * C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lr[c]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += __expf(lr[c] - mx);
    se = warp_sum(se);
    const float inv = 1.0f / se;
    const int label = (int)labels[b];
    for (int c = lane; c < C; c += 32) dlogits[(int64_t)b * C + c] = w * (__expf(lr[c] - mx) * inv - (c == label ? 1.f : 0.f));
}

int unlearn_ce_grad(const float* logits, const int64_t* labels, const float* sums, int n_remain_local, int B, int C, float beta, float BND,
                    float* dlogits, cudaStream_t s) {
    const int warps = 4;
    GSL_CHECK_CUDA(launch_pdl(unlearn_ce_grad_kernel, dim3((B + warps - 1) / warps), dim3(warps * 32), 0, s, logits, labels, sums, n_remain_local, B, C, beta, BND, dlogits));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gsl
