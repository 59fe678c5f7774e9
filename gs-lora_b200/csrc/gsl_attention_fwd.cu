// gslora-b200: attention forward on tcgen05 tensor cores (Attention.forward, vit_pytorch_face/vit_face.py:358-379).
//
// One persistent CTA per SM walks over work items = (image, head, 128-row query tile); N <= 208 tokens gives one or two items per
// (image, head) pair, which share the K / V slabs.  Per item j:
//   S_j = Q_tile K^T      tcgen05.mma  M=128, N=npad, K=64        -> TMEM S[j & 1]   (double-buffered: 2 x 208 columns)
//   softmax               8 worker warps, thread = (query row, half of the key columns): ONE TMEM read of the own half row into registers,
//                         half-row max (FMNMX3) exchanged with the partner half through shared memory, exp2 on the own half,
//                         fp16 P into a 128B-swizzled K-major tile P[j & 1] (= the A operand of the next MMA), partial row sums to smem
//   O_j = P_j V           tcgen05.mma  M=128, N=64, K=npad; V is consumed straight from its TMA slab as an MN-major B operand
//                         -> TMEM O (its own 64 columns, so S buffers are recycled as soon as the workers have read them)
//   epilogue              scale by 1 / (sum_half0 + sum_half1), fp16 O tile staged in shared memory, one TMA store (rows >= N clipped)
// Software pipeline: the workers run  softmax(j), softmax(j+1), ...; the writers trail them with the epilogues; the MMA issuer runs
// S(j+1) (as soon as S[(j+1) & 1] has been read) and P V of item j (as soon as P_j is written): the tensor core works on the
// neighbouring items while the MUFU-bound softmax of item j runs, and TMEM loads are double-buffered inside the sweeps.
// Warp roles: warps 0-7 softmax workers (setmaxnreg 176), warps 8-11 output writers (O out of TMEM, 1/sum, staging, TMA store, LSE - off
// the workers' critical path), warp 12 TMA producer, warp 13 MMA issuer (+ TMEM allocation), warps 14-15 idle.
#include "gsl_common.cuh"
#include <cuda.h>
#include <cstdlib>
#include "gsl_kernels.h"

namespace gsl {

int make_tmap_qkv(CUtensorMap* map, const void* ptr, int64_t ld, int B, int N, int cols, int npad);

static constexpr int AF_WORKER_WARPS = 8;
static constexpr int AF_WORKERS = AF_WORKER_WARPS * 32;
static constexpr int AF_WRITER_WARPS = 4;
static constexpr int AF_WRITERS = AF_WRITER_WARPS * 32;
static constexpr int AF_THREADS = AF_WORKERS + AF_WRITERS + 128;
static constexpr uint32_t AF_W_WRITE = AF_WORKER_WARPS, AF_W_PROD = AF_W_WRITE + AF_WRITER_WARPS, AF_W_MMA = AF_W_PROD + 1;
static constexpr int AF_MAX_TOKENS = 208;

static constexpr uint32_t AF_QT = 128 * 128;                    // one query tile [128 x 64] fp16
static constexpr uint32_t AF_SLAB = AF_MAX_TOKENS * 128;        // K / V slab [208 x 64] fp16
static constexpr uint32_t AF_PT = 4 * 16384;                    // full P tile [128 q x 256 keys] fp16 as four 64-key K-major blocks
static constexpr uint32_t AF_Q = 0;                             // two parities (per item)
static constexpr uint32_t AF_K = AF_Q + 2 * AF_QT;
static constexpr uint32_t AF_V = AF_K + AF_SLAB;
static constexpr uint32_t AF_PA = AF_V + AF_SLAB;               // P tile of even items: four 64-key blocks of [128 rows x 128 B]
static constexpr uint32_t AF_PB = AF_PA + AF_PT;                // P tile of odd items: with two query tiles these are the tile-1 items, whose
static constexpr uint32_t AF_PB_BYTES = 3 * 80 * 128 + 16384;   //   blocks hold only 80 rows (the M = 128 MMA still reads 128 rows of the last block:
                                                                //   keep that in bounds); with one tile npad <= 128 needs two full blocks
static constexpr uint32_t AF_STAGE = AF_PB + AF_PB_BYTES;       // [128 rows x 64] fp16 staging tile of the O TMA store
static constexpr uint32_t AF_SUMS = AF_STAGE + AF_QT;           // [2 parities][2 halves][128] partial row sums, [2 parities][128] row maxima,
static constexpr uint32_t AF_BARS = AF_SUMS + (2 * 2 * 128 + 2 * 128 + 2 * 2 * 128) * 4; //   [2 parities][2 halves][128] half-row maxima (exchange)
static constexpr uint32_t AF_SMEM = AF_BARS + 256;
static_assert(AF_K % 1024 == 0 && AF_V % 1024 == 0 && AF_PA % 1024 == 0 && AF_PB % 1024 == 0 && AF_STAGE % 1024 == 0, "tiles must stay 1024-byte aligned");
static_assert(AF_SMEM + 1024 <= 232448, "attention forward: shared memory budget");

enum : uint32_t {
    F_FULL_Q = 0 /* +parity */, F_FREE_Q = 2 /* +parity */, F_FULL_K = 4, F_FREE_K = 5, F_FULL_V = 8, F_FREE_V = 9,
    F_S_FULL = 10 /* +buf */, F_S_FREE = 12 /* +buf */, F_P_READY = 14 /* +buf */, F_O_FULL = 16, F_O_FREE = 17, F_STATS_FREE = 18 /* +buf */,
    F_COUNT = 20
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
__device__ __forceinline__ float af_max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ void af_bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(AF_WORKERS) : "memory"); }

// running max over 16 score columns starting at column c0 (columns >= N are padding: S = 0 there, they must not win)
__device__ __forceinline__ float af_max16(const uint32_t (&v)[16], float mx, int c0, int N) {
    if (c0 + 16 <= N) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) mx = af_max3(mx, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) if (c0 + j < N) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    return mx;
}
// p = exp2(s * sl2 - off) for 16 columns -> two 16-byte chunks of fp16; returns the sum of the 16 p
__device__ __forceinline__ float af_exp16(const uint32_t (&v)[16], uint4& c0_out, uint4& c1_out, int c0, int N, float sl2, float off) {
    float p[16];
    if (c0 + 16 <= N) {
#pragma unroll
        for (int j = 0; j < 16; ++j) p[j] = ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -off));
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) p[j] = (c0 + j < N) ? ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -off)) : 0.f;
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j += 2) { s0 += p[j]; s1 += p[j + 1]; }
    c0_out = make_uint4(pack_half2(p[0], p[1]), pack_half2(p[2], p[3]), pack_half2(p[4], p[5]), pack_half2(p[6], p[7]));
    c1_out = make_uint4(pack_half2(p[8], p[9]), pack_half2(p[10], p[11]), pack_half2(p[12], p[13]), pack_half2(p[14], p[15]));
    return s0 + s1;
}

__global__ void __launch_bounds__(AF_THREADS, 1)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ0, const __grid_constant__ CUtensorMap tmQ1, const __grid_constant__ CUtensorMap tmKV,
                        const __grid_constant__ CUtensorMap tmO, float* __restrict__ lse, int B, int N, int heads, float scale, int dbg) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sb = smem_u32(smem);
    const int npad = (N + 15) & ~15;
    const int nt = npad > 128 ? 2 : 1;
    const int n1r = nt == 2 ? npad - 128 : 0;
    const int D = heads * 64;
    auto bar = [&](uint32_t i) { return sb + AF_BARS + 8u * i; };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + AF_BARS + 8 * F_COUNT);
    float* s_sums = reinterpret_cast<float*>(smem + AF_SUMS);
    float* s_rowmax = s_sums + 2 * 2 * 128;        // [2 parities][128]: final row max of the item (for its LSE)
    float* s_max = s_rowmax + 2 * 128;             // [2 parities][2 halves][128]: exchange of the half-row maxima.  Per-parity slots: a warp of
                                                   // one half may run a whole item ahead of its partner half (only the barrier below couples
                                                   // them), so item j + 1's write must not land on the slot the partner still reads for item j
                                                   // (compute-sanitizer racecheck, profiles/r02e_sanitizer_racecheck.log)
    const uint32_t pb_stride = nt == 2 ? 80u * 128u : 16384u;      // block stride of the odd-item P tile
    auto p_base = [&](uint32_t pb) { return pb == 0 ? sb + AF_PA : sb + AF_PB; };
    auto p_stride = [&](uint32_t pb) { return pb == 0 ? 16384u : pb_stride; };

    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const int npairs_total = B * heads;
    // items of this CTA: pairs blockIdx.x, +gridDim.x, ...; item j = (pair it = j / nt, tile t = j % nt)
    const int my_pairs = npairs_total > (int)blockIdx.x ? (npairs_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int nitems = my_pairs * nt;

    if (warp == AF_W_MMA) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ0); tma_prefetch_desc(&tmQ1); tma_prefetch_desc(&tmKV); tma_prefetch_desc(&tmO);
            for (uint32_t i = 0; i < F_COUNT; ++i) mbar_init(bar(i), 1);
            for (uint32_t i = 0; i < 2; ++i) { mbar_init(bar(F_S_FREE + i), AF_WORKER_WARPS); mbar_init(bar(F_P_READY + i), AF_WORKER_WARPS); }
            mbar_init(bar(F_O_FREE), AF_WRITER_WARPS);
            for (uint32_t i = 0; i < 2; ++i) mbar_init(bar(F_STATS_FREE + i), AF_WRITER_WARPS);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<1>(smem_u32(tmem_ptr_smem), 512);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_prologue();     // set-up above overlaps the previous kernel's tail; global memory (TMA, LSE, stats) only from here on
    constexpr uint32_t T_S = 0, T_S_STRIDE = 288, T_O = 224;      // S[0] = [0, 208), O = [224, 288), S[1] = [288, 496): all 32-column aligned

    // 512 threads x 128 registers at launch; producer / MMA drop to 56, the writers to 96, the workers take the 13312 freed
    if (warp >= AF_W_PROD) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    } else if (warp >= AF_W_WRITE) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    }
    if (warp == AF_W_PROD) {
        // ===================================================== TMA producer
        if (lane == 0) {
            for (int j = 0; j < nitems; ++j) {
                const int it = j / nt, t = j % nt;
                const int w = blockIdx.x + it * gridDim.x;
                const int wb = w / heads, wh = w % heads;
                if (t == 0) {
                    if (it >= 1) mbar_wait(bar(F_FREE_K), (it - 1) & 1);
                    mbar_arrive_expect_tx(bar(F_FULL_K), (uint32_t)npad * 128u);
                    af_tma(&tmKV, bar(F_FULL_K), sb + AF_K, D + wh * 64, 0, wb);
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
                // V last: the previous pair's V is released by its final P V MMAs, which the issuer launches only after this item's
                // S MMAs - K and Q of this item must already be on their way or the two warps wait on each other
                if (t == 0) {
                    if (it >= 1) mbar_wait(bar(F_FREE_V), (it - 1) & 1);
                    mbar_arrive_expect_tx(bar(F_FULL_V), (uint32_t)npad * 128u);
                    af_tma(&tmKV, bar(F_FULL_V), sb + AF_V, 2 * D + wh * 64, 0, wb);
                }
            }
        }
    } else if (warp == AF_W_MMA) {
        // ===================================================== MMA issuer
        const uint32_t idesc_s = af_idesc(128, (uint32_t)npad, 0, 0);
        constexpr uint32_t idesc_o = af_idesc(128, 64, 0, 1);       // B = V slab rows, MN-major
        auto issue_pv = [&](int jj) {
            const int it = jj / nt, t = jj % nt;
            const uint32_t pb = jj & 1;
            mbar_wait(bar(F_P_READY + pb), (jj >> 1) & 1);                  // P tile written
            if (jj >= 1) mbar_wait(bar(F_O_FREE), (jj - 1) & 1);            // previous item's O has been read out of TMEM
            if (t == 0) mbar_wait(bar(F_FULL_V), it & 1);
            tcgen05_fence_after();
            if (lane == 0) {
                const int nk = (dbg & 8) ? 1 : npad / 16;
                for (int k = 0; k < nk; ++k)
                    umma_f16<1>(tmem_base + T_O, af_desc(p_base(pb) + (k >> 2) * p_stride(pb) + (k & 3) * 32), af_desc(sb + AF_V + k * 2048), idesc_o, k != 0);
                umma_commit<1>(bar(F_O_FULL));
                if (t == nt - 1) umma_commit<1>(bar(F_FREE_V));
            }
            __syncwarp();
        };
        for (int j = 0; j < nitems; ++j) {
            const int it = j / nt, t = j % nt;
            const uint32_t qb = j & 1;
            if (t == 0) mbar_wait(bar(F_FULL_K), it & 1);
            mbar_wait(bar(F_FULL_Q + qb), (j >> 1) & 1);
            if (j >= 2) mbar_wait(bar(F_S_FREE + qb), ((j >> 1) - 1) & 1);  // the workers have read S of item j - 2
            tcgen05_fence_after();
            if (lane == 0) {
                const uint64_t da = af_desc(sb + AF_Q + qb * AF_QT), db = af_desc(sb + AF_K);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16<1>(tmem_base + T_S + qb * T_S_STRIDE, da + 2 * k, db + 2 * k, idesc_s, k != 0);
                umma_commit<1>(bar(F_S_FULL + qb));
                umma_commit<1>(bar(F_FREE_Q + qb));
                if (t == nt - 1) umma_commit<1>(bar(F_FREE_K));
            }
            __syncwarp();
            if (j >= 1) issue_pv(j - 1);
        }
        if (nitems >= 1) issue_pv(nitems - 1);
    } else if (warp < AF_WORKER_WARPS) {
        // ===================================================== workers: thread = (query row, half of the key columns)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
        const uint32_t quarter = warp & 3, hf = warp >> 2;
        const int rl = quarter * 32 + lane;
        const float sl2 = scale * 1.4426950408889634f;
        const uint32_t tlane = tmem_base + ((quarter * 32u) << 16);
        const int npieces = npad / 16;                      // 16-column pieces of a score row
        const int h0p = (npieces + 1) / 2;                  // half 0 takes pieces [0, h0p), half 1 the rest
        const int p_lo = hf == 0 ? 0 : h0p, p_hi = hf == 0 ? h0p : npieces;
        for (int j = 0; j < nitems; ++j) {
            const int t = j % nt;
            const uint32_t sbuf = j & 1;
            const bool rows_on = (t * 128 + (int)quarter * 32) < N;        // warp-uniform
            const uint32_t ts = tlane + T_S + sbuf * T_S_STRIDE;
            mbar_wait(bar(F_S_FULL + sbuf), (j >> 1) & 1);
            tcgen05_fence_after();
            // ---- ONE TMEM read: the own half of the score row (up to 7 pieces of 16 columns) goes to registers with all loads in flight
            //      at once (a tcgen05.ld round trip is ~200 cycles: reading piece by piece made both sweeps latency-bound); the max sweep
            //      and the exp sweep then run from registers, and S[sbuf] goes back to the MMA issuer before any math
            float mx = -INFINITY;
            float sum = 0.f;
            constexpr int MAXP = 7;                                         // pieces per half: ceil(208 / 16 / 2)
            uint32_t sv[MAXP][16];
            const int cnt = p_hi - p_lo;                                    // <= MAXP
            if (rows_on) {
#pragma unroll
                for (int k = 0; k < MAXP; ++k) if (k < cnt) af_ld16(ts + (p_lo + k) * 16, sv[k]);
                tmem_ld_wait();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(F_S_FREE + sbuf));               // S[sbuf] may be overwritten by item j + 2
            if (rows_on && !(dbg & 1)) {
#pragma unroll
                for (int k = 0; k < MAXP; ++k) if (k < cnt) mx = af_max16(sv[k], mx, (p_lo + k) * 16, N);
            }
            s_max[sbuf * 256 + hf * 128 + rl] = mx;
            af_bar_workers();                   // exchange of the half-row maxima
            mx = fmaxf(mx, s_max[sbuf * 256 + (hf ^ 1) * 128 + rl]);
            if (dbg & 1) mx = 0.f;
            if (rows_on && !(dbg & 2)) {
                // ---- exp sweep: p = exp2(scale * log2e * (s - max)), fp16 -> swizzled K-major P tile
                const float off = mx * sl2;
                const uint32_t ptile = p_base(sbuf), pstr = p_stride(sbuf);
                const bool row_fits = nt == 1 || t == 0 || rl < n1r;       // tile-1 blocks hold n1r rows only (the rest is >= N anyway)
#pragma unroll
                for (int k = 0; k < MAXP; ++k) {
                    if (k < cnt) {
                        const int pc = p_lo + k;
                        uint4 c0, c1;
                        sum += af_exp16(sv[k], c0, c1, pc * 16, N, sl2, off);
                        if (!(dbg & 4) && row_fits) {
                            const uint32_t blk = ptile + (pc >> 2) * pstr;      // 64-key block
                            const int ch = (pc & 3) * 2;                        // first 16-byte chunk of this piece inside the block
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + sw128_off(rl, ch)), "r"(c0.x), "r"(c0.y), "r"(c0.z), "r"(c0.w) : "memory");
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + sw128_off(rl, ch + 1)), "r"(c1.x), "r"(c1.y), "r"(c1.z), "r"(c1.w) : "memory");
                        }
                    }
                }
            }
            // row statistics for the writers (the writers of item j - 2 must be done with this parity's slots)
            if (j >= 2) mbar_wait(bar(F_STATS_FREE + sbuf), ((j >> 1) - 1) & 1);
            s_sums[sbuf * 256 + hf * 128 + rl] = sum;
            if (hf == 0) s_rowmax[sbuf * 128 + rl] = mx;
            fence_proxy_async_smem();          // P tile -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(F_P_READY + sbuf));      // (release: the statistics above are ordered before the P V MMAs and O_FULL)
        }
    } else if (warp < AF_W_PROD) {
        // ===================================================== output writers: thread = one row of the [128 x 64] O tile
        const uint32_t quarter = warp & 3;
        const int rl = quarter * 32 + lane;
        const uint32_t tlane = tmem_base + ((quarter * 32u) << 16);
        const bool elected = warp == AF_W_WRITE && lane == 0;
        const uint32_t stage = sb + AF_STAGE;
        for (int jj = 0; jj < nitems; ++jj) {
            const int it = jj / nt, t = jj % nt;
            const int w = blockIdx.x + it * gridDim.x;
            const int b = w / heads, h = w % heads;
            const uint32_t pb = jj & 1;
            const int row = t * 128 + rl;
            const bool rows_on = (t * 128 + (int)quarter * 32) < N;
            uint32_t o[32];
            mbar_wait(bar(F_O_FULL), jj & 1);
            tcgen05_fence_after();
            const float sum = s_sums[pb * 256 + rl] + s_sums[pb * 256 + 128 + rl];
            const float mx = s_rowmax[pb * 128 + rl];
            const float inv = 1.0f / sum;
            if (elected) tma_store_wait_read<0>();      // the previous store has finished reading the staging tile
            asm volatile("bar.sync 2, %0;" ::"n"(AF_WRITERS) : "memory");
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                if (rows_on) {
                    tmem_ld_32x32(tlane + T_O + hh * 32, o);
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
            if (lane == 0) { mbar_arrive(bar(F_O_FREE)); mbar_arrive(bar(F_STATS_FREE + pb)); }     // O and this parity's statistics have been read
            if (rows_on && row < N && lse != nullptr) lse[((int64_t)b * heads + h) * N + row] = mx * scale + __logf(sum);
            fence_proxy_async_smem();
            asm volatile("bar.sync 2, %0;" ::"n"(AF_WRITERS) : "memory");
            if (elected && !(dbg & 16)) {
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
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("GSL_ATTN_DBG"); dbg = e ? atoi(e) : 0; }       // dev switch: knock out stages to time the rest
    GSL_CHECK_CUDA(launch_pdl(attention_fwd_tc_kernel, dim3(nwork < sms ? nwork : sms), dim3(AF_THREADS), smem, s, tq0, tq1, tkv, to, lse, B, N, heads, scale, dbg));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gsl
