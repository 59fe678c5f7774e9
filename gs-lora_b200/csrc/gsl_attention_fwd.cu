// gslora-b200: attention forward on tcgen05 tensor cores (Attention.forward, vit_pytorch_face/vit_face.py:358-379).
//
// One persistent CTA per SM walks over work items = (image, head, 128-row query tile); N <= 208 tokens gives one or two items per
// (image, head) pair, which share the K / V slabs.  Everything an item needs after its Q tile lives in TENSOR MEMORY; item j uses
// buffer b = j & 1 (256 columns each):
//   S_j = Q_tile K^T      tcgen05.mma  M=128, N=npad, K=64  (A, B from shared memory)   -> TMEM columns [0, npad) of buffer b
//   softmax               thread = query row (TMEM lane), two sweeps over the row with double-buffered 16-column tcgen05.ld:
//                         max sweep, then exp2 sweep; the fp16 probabilities go straight BACK INTO TMEM with tcgen05.st, over the
//                         low half of the score columns they came from (P[:, 2c : 2c+2] packs into S column c, which the sweep
//                         has already read) -- no shared-memory P tile, no proxy fence, no cross-thread max exchange
//   O_j = P_j V           tcgen05.mma with the A operand read from TMEM (P, K-major), B = V slab rows as MN-major operand
//                         -> TMEM columns [128, 192) of buffer b (score columns the exp sweep is done with)
//   epilogue              1 / rowsum, fp16 O tile staged in shared memory (two staging tiles), one TMA store (rows >= N clipped)
// Two worker groups of four warps work on the two buffers; the tensor core runs S of item j + 1 and P V of item j - 1 and the writers
// drain O of item j - 1 while item j is exponentiated.  K and V slabs are double-buffered across (image, head) pairs: the next pair's
// slabs stream in during the current pair's softmax.
// Measured (B200, config-2 shape, clock64 phase trace of one CTA, scripts/dev_attn_trace.py): 0.26 ms per launch (round 2's earlier
// single-group kernel with P in shared memory: 0.354 ms), 4450 clk per item.  The bound is the TMEM READ path, not the MUFU pipe: one
// sweep over the [128 x 208] fp32 scores (106 KB) takes 1700 clk whatever the prefetch depth = 64 B / clk / SM, the exp sweep the same
// plus the O drain (32 KB), i.e. ~3850 clk of tcgen05.ld per item against 1664 clk of MUFU.EX2.  A single-sweep softmax would need the
// row in registers (2 threads x 104 values), which leaves no register file for a second group.  Forcing the groups to alternate their
// exp sweeps (a turn barrier) was measured 5 % slower: one warp per scheduler cannot keep the MUFU pipe busy on its own.
// Warp roles: warps 0-3 worker group 0, 4-7 worker group 1, 8-11 output writers, 12 TMA producer, 13 MMA issuer (+ TMEM allocation).
#include "gsl_common.cuh"
#include <cuda.h>
#include <cstdlib>
#include "gsl_kernels.h"

namespace gsl {

int make_tmap_qkv(CUtensorMap* map, const void* ptr, int64_t ld, int B, int N, int cols, int npad);

static constexpr int AF_GROUP_WARPS = 4;
static constexpr int AF_WORKER_WARPS = 2 * AF_GROUP_WARPS;
static constexpr int AF_WRITER_WARPS = 4;
static constexpr int AF_WRITERS = AF_WRITER_WARPS * 32;
static constexpr int AF_THREADS = (AF_WORKER_WARPS + AF_WRITER_WARPS + 2) * 32;
static constexpr uint32_t AF_W_WRITE = AF_WORKER_WARPS, AF_W_PROD = AF_W_WRITE + AF_WRITER_WARPS, AF_W_MMA = AF_W_PROD + 1;
static constexpr int AF_MAX_TOKENS = 208;

static constexpr uint32_t AF_QT = 128 * 128;                    // one query tile [128 x 64] fp16
static constexpr uint32_t AF_SLAB = AF_MAX_TOKENS * 128;        // K / V slab [208 x 64] fp16
static constexpr uint32_t AF_Q = 0;                             // two parities (per item)
static constexpr uint32_t AF_K = AF_Q + 2 * AF_QT;              // two parities (per pair)
static constexpr uint32_t AF_V = AF_K + 2 * AF_SLAB;            // two parities (per pair)
static constexpr uint32_t AF_STAGE = AF_V + 2 * AF_SLAB;        // two [128 rows x 64] fp16 staging tiles of the O TMA stores
static constexpr uint32_t AF_STATS = AF_STAGE + 2 * AF_QT;      // [2 buffers][sum, max][128] row statistics for the writers
static constexpr uint32_t AF_BARS = AF_STATS + 2 * 2 * 128 * 4;
static constexpr uint32_t AF_SMEM = AF_BARS + 256;
static_assert(AF_K % 1024 == 0 && AF_V % 1024 == 0 && AF_SLAB % 1024 == 0 && AF_STAGE % 1024 == 0, "tiles must stay 1024-byte aligned");
static_assert(AF_SMEM + 1024 <= 232448, "attention forward: shared memory budget");

// tensor memory: buffer b at column 256 b: scores [0, 208), P (fp16 pairs) over [0, 104), O accumulator [128, 192)
static constexpr uint32_t AF_T_BUF = 256, AF_T_O = 128;

#ifdef GSL_ATTN_TRACE      // dev build only (scripts/dev_attn_trace.py): per-phase clock stamps of CTA 0, [role 0..15][item 0..63][slot 0..3]
__device__ long long g_af_trace[16 * 64 * 4];
#define AF_TRACE(role, item, slot) do { if (blockIdx.x == 0 && lane == 0 && (item) < 64) g_af_trace[((role) * 64 + (item)) * 4 + (slot)] = clock64(); } while (0)
#else
#define AF_TRACE(role, item, slot) do { } while (0)
#endif

enum : uint32_t {   // mbarrier indices (+ parity / buffer)
    F_FULL_Q = 0, F_FREE_Q = 2, F_FULL_K = 4, F_FREE_K = 6, F_FULL_V = 8, F_FREE_V = 10,
    F_S_FULL = 12, F_P_READY = 14, F_O_FULL = 16, F_O_FREE = 18, F_STATS_READY = 20, F_STATS_FREE = 22, F_COUNT = 24
};

__device__ __forceinline__ uint64_t af_desc(uint32_t smem_addr) {      // 128-byte rows, 128B swizzle, 8-row atoms of 1024 bytes
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __host__ constexpr uint32_t af_idesc(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void af_tma(const void* desc, uint32_t bar, uint32_t dst, int col, int row, int b) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(col), "r"(row), "r"(b) : "memory");
}
__device__ __forceinline__ void af_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
// 8 consecutive 32-bit columns of this thread's TMEM lane <- 16 packed fp16
__device__ __forceinline__ void af_st8(uint32_t taddr, const uint4& a, const uint4& b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void af_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D (TMEM) += A (TMEM, K-major: lane = row, column c holds k = 2c, 2c + 1) * B (shared memory descriptor)
__device__ __forceinline__ void af_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float af_max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// running max over 16 score columns starting at column c0 (columns >= N are padding: S = 0 there, they must not win); a tree of
// 3-input maxima (depth 3) instead of a 8-deep chain on the running value
__device__ __forceinline__ float af_max16(const uint32_t (&v)[16], float mx, int c0, int N) {
    if (c0 + 16 <= N) {
        auto f = [&](int i) { return __uint_as_float(v[i]); };
        const float a = af_max3(f(0), f(1), f(2)), b = af_max3(f(3), f(4), f(5)), c = af_max3(f(6), f(7), f(8));
        const float d = af_max3(f(9), f(10), f(11)), e = af_max3(f(12), f(13), f(14));
        mx = af_max3(af_max3(mx, a, b), af_max3(c, d, e), f(15));
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) if (c0 + j < N) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    return mx;
}
// The exp sweep is software-pipelined by hand across the 16-column pieces: phase 1 (16 FFMA + 16 MUFU.EX2) of piece k + 1 is issued BEFORE
// phase 2 (row sum, fp16 packing, tcgen05.st) of piece k, so the MUFU pipe always has the next piece's exponentials queued while the
// dependent tail of the previous piece drains (one warp per scheduler runs this sweep at a time: without the overlap every piece paid its
// MUFU -> FADD -> F2FP -> STTM latency chain serially, 14 clk per exponential instead of the pipe's 8).
// The registers a tcgen05.ld has filled are "pinned" after tcgen05.wait::ld so that no consumer can be scheduled above the wait.
__device__ __forceinline__ void af_pin16(uint32_t (&v)[16]) {
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                      "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory");
}
// phase 1: p = exp2(s * sl2 - off) for 16 columns (columns >= N: p = 0)
__device__ __forceinline__ void af_exp16_issue(const uint32_t (&v)[16], float (&p)[16], int c0, int N, float sl2, float off) {
    if (c0 + 16 <= N) {
#pragma unroll
        for (int j = 0; j < 16; ++j) p[j] = ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -off));
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) p[j] = (c0 + j < N) ? ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -off)) : 0.f;
    }
}
// phase 2: row-sum contribution and the 16 fp16 values as two 16-byte chunks
__device__ __forceinline__ float af_exp16_finish(const float (&p)[16], uint4& c0_out, uint4& c1_out) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j += 4) { s0 += p[j]; s1 += p[j + 1]; s2 += p[j + 2]; s3 += p[j + 3]; }
    c0_out = make_uint4(pack_half2(p[0], p[1]), pack_half2(p[2], p[3]), pack_half2(p[4], p[5]), pack_half2(p[6], p[7]));
    c1_out = make_uint4(pack_half2(p[8], p[9]), pack_half2(p[10], p[11]), pack_half2(p[12], p[13]), pack_half2(p[14], p[15]));
    return (s0 + s1) + (s2 + s3);
}

__global__ void __launch_bounds__(AF_THREADS, 1)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ0, const __grid_constant__ CUtensorMap tmQ1, const __grid_constant__ CUtensorMap tmKV,
                        const __grid_constant__ CUtensorMap tmO, float* __restrict__ lse, int B, int N, int heads, float scale) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sb = smem_u32(smem);
    const int npad = (N + 15) & ~15;
    const int nt = npad > 128 ? 2 : 1;
    const int n1r = nt == 2 ? npad - 128 : 0;
    const int D = heads * 64;
    auto bar = [&](uint32_t i) { return sb + AF_BARS + 8u * i; };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + AF_BARS + 8 * F_COUNT);
    float* s_sum = reinterpret_cast<float*>(smem + AF_STATS);       // [2 buffers][128]
    float* s_rowmax = s_sum + 2 * 128;                              // [2 buffers][128]

    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const int npairs_total = B * heads;
    // items of this CTA: pairs blockIdx.x, +gridDim.x, ...; item j = (pair it = j / nt, tile t = j % nt)
    const int my_pairs = npairs_total > (int)blockIdx.x ? (npairs_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int nitems = my_pairs * nt;

    if (warp == AF_W_MMA) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ0); tma_prefetch_desc(&tmQ1); tma_prefetch_desc(&tmKV); tma_prefetch_desc(&tmO);
            for (uint32_t i = 0; i < F_COUNT; ++i) mbar_init(bar(i), 1);
            // worker / writer barriers count WARPS (one elected lane arrives after __syncwarp)
            for (uint32_t i = 0; i < 2; ++i) {
                mbar_init(bar(F_P_READY + i), AF_GROUP_WARPS); mbar_init(bar(F_O_FREE + i), AF_WRITER_WARPS);
                mbar_init(bar(F_STATS_READY + i), AF_GROUP_WARPS); mbar_init(bar(F_STATS_FREE + i), AF_WRITER_WARPS);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<1>(smem_u32(tmem_ptr_smem), 512);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_prologue();     // set-up above overlaps the previous kernel's tail; global memory (TMA, LSE) only from here on

    if (warp == AF_W_PROD) {
        // ===================================================== TMA producer (runs up to a whole pair ahead: K / V have two slabs each)
        if (lane == 0) {
            for (int j = 0; j < nitems; ++j) {
                const int it = j / nt, t = j % nt;
                const int w = blockIdx.x + it * gridDim.x;
                const int wb = w / heads, wh = w % heads;
                const uint32_t kb = it & 1, uk = it >> 1;
                if (t == 0) {
                    if (uk >= 1) mbar_wait(bar(F_FREE_K + kb), (uk - 1) & 1);
                    mbar_arrive_expect_tx(bar(F_FULL_K + kb), (uint32_t)npad * 128u);
                    af_tma(&tmKV, bar(F_FULL_K + kb), sb + AF_K + kb * AF_SLAB, D + wh * 64, 0, wb);
                }
                const uint32_t qb = j & 1, u = j >> 1;
                if (u >= 1) mbar_wait(bar(F_FREE_Q + qb), (u - 1) & 1);
                if (t == 0) {
                    mbar_arrive_expect_tx(bar(F_FULL_Q + qb), AF_QT);
                    af_tma(&tmQ0, bar(F_FULL_Q + qb), sb + AF_Q + qb * AF_QT, wh * 64, 0, wb);
                } else {
                    mbar_arrive_expect_tx(bar(F_FULL_Q + qb), (uint32_t)n1r * 128u);
                    af_tma(&tmQ1, bar(F_FULL_Q + qb), sb + AF_Q + qb * AF_QT, wh * 64, 128, wb);
                }
                if (t == 0) {
                    if (uk >= 1) mbar_wait(bar(F_FREE_V + kb), (uk - 1) & 1);
                    mbar_arrive_expect_tx(bar(F_FULL_V + kb), (uint32_t)npad * 128u);
                    af_tma(&tmKV, bar(F_FULL_V + kb), sb + AF_V + kb * AF_SLAB, 2 * D + wh * 64, 0, wb);
                }
            }
        }
    } else if (warp == AF_W_MMA) {
        // ===================================================== MMA issuer
        const uint32_t idesc_s = af_idesc(128, (uint32_t)npad, 0, 0);
        constexpr uint32_t idesc_o = af_idesc(128, 64, 0, 1);       // A = P from TMEM (K-major), B = V slab rows, MN-major
        auto issue_pv = [&](int jj) {
            const int it = jj / nt, t = jj % nt;
            const uint32_t pb = jj & 1, kb = it & 1;
            mbar_wait(bar(F_P_READY + pb), (jj >> 1) & 1);                  // P of item jj sits in TMEM
            if (t == 0) mbar_wait(bar(F_FULL_V + kb), (it >> 1) & 1);
            tcgen05_fence_after();
            AF_TRACE(13, jj, 2);
            if (lane == 0) {
                const uint32_t tb = tmem_base + pb * AF_T_BUF;
                const int nk = npad / 16;
                for (int k = 0; k < nk; ++k) af_umma_ts(tb + AF_T_O, tb + 8 * k, af_desc(sb + AF_V + kb * AF_SLAB + k * 2048), idesc_o, k != 0);
                umma_commit<1>(bar(F_O_FULL + pb));
                if (t == nt - 1) umma_commit<1>(bar(F_FREE_V + kb));
            }
            __syncwarp();
        };
        for (int j = 0; j < nitems; ++j) {
            const int it = j / nt, t = j % nt;
            const uint32_t qb = j & 1, kb = it & 1;
            if (t == 0) mbar_wait(bar(F_FULL_K + kb), (it >> 1) & 1);
            mbar_wait(bar(F_FULL_Q + qb), (j >> 1) & 1);
            AF_TRACE(13, j, 0);
            if (j >= 2) mbar_wait(bar(F_O_FREE + qb), ((j >> 1) - 1) & 1);  // item j - 2 has left this TMEM buffer (P consumed, O read out)
            tcgen05_fence_after();
            AF_TRACE(13, j, 1);
            if (lane == 0) {
                const uint64_t da = af_desc(sb + AF_Q + qb * AF_QT), db = af_desc(sb + AF_K + kb * AF_SLAB);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16<1>(tmem_base + qb * AF_T_BUF, da + 2 * k, db + 2 * k, idesc_s, k != 0);
                umma_commit<1>(bar(F_S_FULL + qb));
                umma_commit<1>(bar(F_FREE_Q + qb));
                if (t == nt - 1) umma_commit<1>(bar(F_FREE_K + kb));
            }
            __syncwarp();
            if (j >= 1) issue_pv(j - 1);
        }
        if (nitems >= 1) issue_pv(nitems - 1);
    } else if (warp < AF_WORKER_WARPS) {
        // ===================================================== workers: group g owns TMEM buffer g, thread = one query row
        const uint32_t g = warp >> 2, quarter = warp & 3;
        const int rl = quarter * 32 + lane;
        const float sl2 = scale * 1.4426950408889634f;
        const uint32_t ts = tmem_base + ((quarter * 32u) << 16) + g * AF_T_BUF;
        const int npieces = npad / 16;                      // 16-column pieces of a score row
        for (int j = (int)g; j < nitems; j += 2) {
            const int t = j % nt;
            const uint32_t u = (uint32_t)j >> 1;
            const bool rows_on = (t * 128 + (int)quarter * 32) < N;        // warp-uniform
            AF_TRACE(warp, j, 0);
            mbar_wait(bar(F_S_FULL + g), u & 1);
            tcgen05_fence_after();
            AF_TRACE(warp, j, 1);
            if (rows_on) {
                uint32_t va[16], vb[16];
                // ---- max sweep: piece k + 1 is in flight while piece k is reduced (deeper prefetch does not help: the sweep runs at the
                //      TMEM read bandwidth, measured 1700 clk for the 106 KB of a [128 x 208] fp32 score tile = 64 B / clk / SM)
                float mx = -INFINITY;
                af_ld16(ts, va);
                for (int k = 0; k < npieces; k += 2) {
                    tmem_ld_wait();
                    af_pin16(va);
                    if (k + 1 < npieces) af_ld16(ts + (k + 1) * 16, vb);
                    mx = af_max16(va, mx, k * 16, N);
                    if (k + 1 < npieces) {
                        tmem_ld_wait();
                        af_pin16(vb);
                        if (k + 2 < npieces) af_ld16(ts + (k + 2) * 16, va);
                        mx = af_max16(vb, mx, (k + 1) * 16, N);
                    }
                }
                // ---- exp sweep: p = exp2(scale * log2e * (s - max)) -> fp16 pairs back into TMEM columns [8 k, 8 k + 8) (already read)
                AF_TRACE(warp, j, 2);
                const float off = mx * sl2;
                float sum = 0.f;
                float pa[16], pb[16];
                uint4 c0, c1;
                af_ld16(ts, va);
                tmem_ld_wait();
                af_pin16(va);
                if (1 < npieces) af_ld16(ts + 16, vb);
                af_exp16_issue(va, pa, 0, N, sl2, off);
                for (int k = 0; k < npieces; k += 2) {
                    // here: pa = exponentials of piece k (in flight), vb = scores of piece k + 1 (load in flight)
                    if (k + 1 < npieces) {
                        tmem_ld_wait();
                        af_pin16(vb);
                        if (k + 2 < npieces) af_ld16(ts + (k + 2) * 16, va);
                        af_exp16_issue(vb, pb, (k + 1) * 16, N, sl2, off);
                    }
                    sum += af_exp16_finish(pa, c0, c1);
                    af_st8(ts + k * 8, c0, c1);
                    if (k + 1 < npieces) {
                        if (k + 2 < npieces) {
                            tmem_ld_wait();
                            af_pin16(va);
                            if (k + 3 < npieces) af_ld16(ts + (k + 3) * 16, vb);
                            af_exp16_issue(va, pa, (k + 2) * 16, N, sl2, off);
                        }
                        sum += af_exp16_finish(pb, c0, c1);
                        af_st8(ts + (k + 1) * 8, c0, c1);
                    }
                }
                af_st_wait();
                AF_TRACE(warp, j, 3);
                // row statistics for the writers, handed over through their own barrier pair: the P_READY -> MMA -> tcgen05.commit -> O_FULL chain
                // orders them too, but only through the tensor core's asynchronous arrival (compute-sanitizer racecheck cannot follow that)
                if (u >= 1) mbar_wait(bar(F_STATS_FREE + g), (u - 1) & 1);
                s_sum[g * 128 + rl] = sum;
                s_rowmax[g * 128 + rl] = mx;
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar(F_STATS_READY + g)); mbar_arrive(bar(F_P_READY + g)); }
        }
    } else if (warp < AF_W_PROD) {
        // ===================================================== output writers: thread = one row of the [128 x 64] O tile
        const uint32_t quarter = warp & 3;
        const int rl = quarter * 32 + lane;
        const uint32_t tlane = tmem_base + ((quarter * 32u) << 16);
        const bool elected = warp == AF_W_WRITE && lane == 0;
        for (int jj = 0; jj < nitems; ++jj) {
            const int it = jj / nt, t = jj % nt;
            const int w = blockIdx.x + it * gridDim.x;
            const int b = w / heads, h = w % heads;
            const uint32_t pb = jj & 1;
            const uint32_t stage = sb + AF_STAGE + pb * AF_QT;
            const int row = t * 128 + rl;
            const bool rows_on = (t * 128 + (int)quarter * 32) < N;
            uint32_t o[32];
            mbar_wait(bar(F_O_FULL + pb), (jj >> 1) & 1);
            tcgen05_fence_after();
            AF_TRACE(warp, jj, 0);
            mbar_wait(bar(F_STATS_READY + pb), (jj >> 1) & 1);
            const float sum = rows_on ? s_sum[pb * 128 + rl] : 1.f;
            const float mx = rows_on ? s_rowmax[pb * 128 + rl] : 0.f;
            const float inv = 1.0f / sum;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(F_STATS_FREE + pb));
            if (elected) tma_store_wait_read<1>();      // the store of item jj - 2 has finished reading this staging tile
            asm volatile("bar.sync 2, %0;" ::"n"(AF_WRITERS) : "memory");
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                if (rows_on) {
                    tmem_ld_32x32(tlane + pb * AF_T_BUF + AF_T_O + hh * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + sw128_off(rl, hh * 4 + q)),
                                     "r"(pack_half2(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv)),
                                     "r"(pack_half2(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv)),
                                     "r"(pack_half2(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv)),
                                     "r"(pack_half2(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv)) : "memory");
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            AF_TRACE(warp, jj, 1);
            if (lane == 0) mbar_arrive(bar(F_O_FREE + pb));         // O and this buffer's statistics have been read: the buffer may take item jj + 2
            if (rows_on && row < N && lse != nullptr) lse[((int64_t)b * heads + h) * N + row] = mx * scale + __logf(sum);
            fence_proxy_async_smem();
            asm volatile("bar.sync 2, %0;" ::"n"(AF_WRITERS) : "memory");
            if (elected) {
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(stage), "r"(h * 64), "r"(t * 128), "r"(b) : "memory");
                tma_store_commit();
            }
        }
        if (elected) tma_store_wait_all();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == AF_W_MMA) {
        tcgen05_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

int attention_fwd(const __half* qkv, int64_t ld, __half* out, int64_t ldo, float* lse, int B, int N, int heads, float scale, cudaStream_t s) {
    GSL_REQUIRE(N >= 1 && N <= AF_MAX_TOKENS, "attention: tokens=%d outside [1, %d]", N, AF_MAX_TOKENS);
    GSL_REQUIRE(ld % 8 == 0 && ldo % 8 == 0, "attention: pitches must be multiples of 8 halves");
    GSL_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "attention: output must be 16-byte aligned");
    const int npad = (N + 15) & ~15;
    const int n1 = npad > 128 ? npad - 128 : 16;        // (the tile-1 map is unused when N <= 128)
    CUtensorMap tq0, tq1, tkv, to;
    int rc;
    if ((rc = make_tmap_qkv(&tq0, qkv, ld, B, N, 3 * heads * 64, 128))) return rc;
    if ((rc = make_tmap_qkv(&tq1, qkv, ld, B, N, 3 * heads * 64, n1))) return rc;
    if ((rc = make_tmap_qkv(&tkv, qkv, ld, B, N, 3 * heads * 64, npad))) return rc;
    if ((rc = make_tmap_qkv(&to, out, ldo, B, N, heads * 64, 128))) return rc;
    const int smem = AF_SMEM + 1024;
    static bool attr = false;
    if (!attr) {
        GSL_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    const int sms = device_sm_count();
    const int nwork = B * heads;
    GSL_CHECK_CUDA(launch_pdl(attention_fwd_tc_kernel, dim3(nwork < sms ? nwork : sms), dim3(AF_THREADS), smem, s, tq0, tq1, tkv, to, lse, B, N, heads, scale));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

#ifdef GSL_ATTN_TRACE
extern "C" int gsl_debug_attn_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_af_trace, sizeof(g_af_trace));
}
#endif

}  // namespace gsl
