// gslora-b200: the frozen-weight GEMM family of the GS-LoRA hot path on tcgen05 tensor cores.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )      A, B fp16 K-major, fp32 accumulation in TMEM
//
// This one kernel template serves every dense contraction of the reference's step
// (vit_pytorch_face/vit_face.py:326-379 forward, autograd's dX GEMMs in engine_cl.py:124):
//   QKV / attention-out / patch-embed projections, FFN fc1 (+bias +GELU) and fc2 (+bias +residual),
//   and the backward dX GEMMs through the frozen weights (with the GELU' epilogue).
// The LoRA branch  s*(x A^T) B^T  of loralib.Linear.forward is folded into the cached operand by the engine
// (W' = W + s B A, gsl_engine.cu merge_weights_kernel), so the FFN GEMMs are plain dense contractions here.
//
// SPLIT (precision mode "split", the default of the engine): the B operand arrives as TWO fp16 matrices B_hi + B_lo
// (B_hi = fp16(W), B_lo = fp16(W - B_hi): 22 significand bits of the fp32 weight survive) and every k-step issues two MMAs
// on the SAME A tile into the same accumulator.  The frozen weights are the only operands whose rounding is systematic
// (identical for every token of every step); with them exact to 2^-22 the LoRA gradients meet the 1e-3 parity bar
// (DESIGN.md section 2), for +1 B tile of shared memory per stage and 2x the tensor-pipe work.
//
// SPLIT = 2 (precision mode "split8"): the residual term runs on the FP8 tensor path at twice the rate.  B_hi = fp16(W * 2^s) and
// B_lo8 = e4m3(W * 2^s - B_hi) (the power-of-two pre-scale puts the residual, ~2^-12 |W|, into e4m3's range; the epilogue multiplies the
// accumulator by 2^-s); two converter warps per CTA turn every fp16 A tile that lands in shared memory into an e5m2 copy (64-byte rows,
// 64B swizzle), and each 64-wide k-block issues 4 x kind::f16 (A * B_hi^T, K = 16) + 2 x kind::f8f6f4 (A8 * B_lo8^T, K = 32) into the same
// fp32 accumulator: 1.5x the tensor work of the plain GEMM instead of 2x.  The weight keeps ~15 significant bits (11 + e4m3's 4), the
// error of the e5m2 activation copy only touches the 2^-12-sized residual term (2^-15 relative, random per element).
// Measured (B200, M = 201 728): the K = 1536 / 2048 GEMMs with the light fp16 epilogue (4-stage operand ring) run 10-16 % faster than with
// the fp16 residual and are tensor-bound again (90-92 % active); the GEMMs whose epilogue staging leaves 3 stages (fc2, dH) gain 1-3 %
// (the TMA -> convert -> MMA chain is one hop longer), the epilogue-bound fc1 loses 4 % (the engine keeps the fp16 residual there).
// Shared-memory bandwidth (TMA writes + converter + MMA operand reads = 112 KB per k-block) sits at ~85 % of its peak in both modes.
//
// Structure (persistent, warp specialised, 576 threads):
//   warp 0    TMA producer      cp.async.bulk.tensor 128B-swizzled A/B tiles -> smem ring (mbarrier full/empty)
//   warp 1    MMA issuer        one lane issues tcgen05.mma (cta_group::1 or ::2), accumulators in TMEM,
//                               tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2-17 epilogue         two groups of 8 warps on alternate 128 x 64 (fp16) / 128 x 32 (fp32) stripes of the accumulator:
//                               tcgen05.ld TMEM -> registers -> fused math -> swizzled smem -> one TMA store per stripe;
//                               auxiliary input stripes (residual / pre-GELU H) arrive by TMA
// Two TMEM accumulator buffers (2 x BLOCK_N columns) overlap the epilogue of tile i with the MMAs of tile i+1.
#include "gsl_common.cuh"
#include "gsl_kernels.h"

#include <cudaTypedefs.h>
#include <mutex>

namespace gsl {

static constexpr int BLOCK_M = 128;   // rows per CTA (UMMA M = 128 * CTA_GROUP)
static constexpr int BLOCK_K = 64;    // 64 fp16 = one 128-byte swizzle row
static constexpr int SMEM_LIMIT = 232448;

struct GemmParams {
    int M, N, K;
    int num_m_tiles;      // super-tiles of 128 * CTA_GROUP rows
    int num_n_tiles;
    const float* bias;    // [N] or nullptr
    const float* table;   // EPI_PERIODIC_F32: fp32 [period, ld_table]
    int period, ld_table;
    int has_out1;         // EPI_F32: also emit an fp16 copy through tmO1
    float* rowdot;        // EPI_F16_ROWDOT: [2][M / period][N / 64][period]
    float acc_scale;      // split8: 2^-s, undoes the pre-scale of the (B_hi, B_lo8) pair on the accumulator
    uint32_t drop_thresh; // round(p * 32768) (0 = no dropout)
    DropSeed drop_seed;
    float drop_scale;     // 1 / (1 - p)
    GeluConsts gelu;      // EPI_GELU: constants of gelu_pair, scaled by the keep scale
};

template <int EPI> struct EpiTraits;
//                                                          out0 bytes/elem, out1 bytes/elem (0 = none), aux bytes/elem (0 = none)
template <> struct EpiTraits<EPI_F16>          { static constexpr int O0 = 2, O1 = 0, AUX = 0; };
template <> struct EpiTraits<EPI_F32>          { static constexpr int O0 = 4, O1 = 2, AUX = 0; };
template <> struct EpiTraits<EPI_GELU>         { static constexpr int O0 = 2, O1 = 2, AUX = 0; };
template <> struct EpiTraits<EPI_GELU_BWD>     { static constexpr int O0 = 2, O1 = 0, AUX = 2; };
template <> struct EpiTraits<EPI_RES_F32>      { static constexpr int O0 = 4, O1 = 0, AUX = 4; };
template <> struct EpiTraits<EPI_PERIODIC_F32> { static constexpr int O0 = 4, O1 = 0, AUX = 0; };
template <> struct EpiTraits<EPI_F16_ROWDOT>   { static constexpr int O0 = 2, O1 = 0, AUX = 2; };

static constexpr int EPI_GROUPS = 2;                      // two independent epilogue groups work on alternate stripes
static constexpr int EPI_GROUP_WARPS = 8;                 // per group: two warps per TMEM lane quarter, each takes half of a stripe's columns
static constexpr int EPI_WARPS = EPI_GROUPS * EPI_GROUP_WARPS;

template <int CG, int BN, int EPI, int SPLIT = 0, int EG = EPI_GROUPS>
struct GemmCfg {
    using T = EpiTraits<EPI>;
    static constexpr int A_STAGE = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_STAGE = (BN / CG) * BLOCK_K * 2;
    static constexpr int A8_STAGE = SPLIT == 2 ? BLOCK_M * BLOCK_K : 0;                            // e5m2 copy of the A tile (split8)
    static constexpr int B2_STAGE = SPLIT == 1 ? B_STAGE : SPLIT == 2 ? (BN / CG) * BLOCK_K : 0;   // B_lo tile: fp16, or e4m3 (split8)
    static constexpr int A_STRIDE = A_STAGE + A8_STAGE;  // per stage: A tile [, A8 tile]
    static constexpr int B_STRIDE = B_STAGE + B2_STAGE;  // per stage: B_hi tile [, B_lo tile]
    static constexpr int STAGE = A_STRIDE + B_STRIDE;
    static constexpr int CONV_WARPS = SPLIT == 2 ? 2 : 0;                                          // fp16 -> e5m2 converters of the A tiles
    static constexpr int THREADS = 64 + EPI_WARPS * 32 + CONV_WARPS * 32;
    // the epilogue works on stripes of the 128 x BN accumulator: 64 columns when every output is fp16, 32 columns when
    // the main output is fp32 -- either way one 128-byte swizzle row per accumulator row, one TMA store per stripe.
    static constexpr int STRIPE = T::O0 == 2 ? 64 : 32;
    static constexpr int CW = STRIPE / 2;                 // columns per epilogue warp
    static constexpr int O0_BUF = BLOCK_M * STRIPE * T::O0;
    static constexpr int O1_BUF = BLOCK_M * STRIPE * T::O1;
    static constexpr int AUX_BUF = BLOCK_M * STRIPE * T::AUX;
    // aux stripes in flight per group: a TMA round trip is ~3x a stripe's epilogue time, so two when the operand ring keeps >= 4 stages
    // EG = epilogue groups that WORK (1 or 2; the warps of an idle group only hand the accumulator back).  With a long K loop the epilogue
    // of a tile is a fraction of its MMA time, and one group's staging instead of two buys the split8 operand ring its 4th stage -- which
    // the TMA -> convert -> MMA chain of that mode needs (3 stages: tensor pipe 67 % active on fc2 / dH, 4 stages: 90 %).
    static constexpr int AUX_DEPTH = (SMEM_LIMIT - 1024 - EG * (O0_BUF + O1_BUF + 2 * AUX_BUF) - 1024) / STAGE >= 4 ? 2 : 1;
    static constexpr int EPI_TOTAL = EG * (O0_BUF + O1_BUF + AUX_DEPTH * AUX_BUF);   // one staging set per working group (the groups alternate)
    static constexpr int BAR_BYTES = 1024;
    static constexpr int STAGES_RAW = (SMEM_LIMIT - 1024 - EPI_TOTAL - BAR_BYTES) / STAGE;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE + EPI_TOTAL + BAR_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    static_assert(STAGES >= (SPLIT ? 2 : 3), "not enough shared memory for the operand ring");
    static_assert(SPLIT != 2 || CG == 2, "split8 is built for cta_group::2 only");
    static_assert(A8_STAGE % 1024 == 0 && B2_STAGE % 1024 == 0, "operand tiles must stay 1024-byte aligned");
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
    static_assert(O0_BUF % 1024 == 0 && (O1_BUF % 1024 == 0) && (AUX_BUF % 1024 == 0), "staging must keep 1024-byte alignment");
};

__device__ __forceinline__ void epi_bar_sync(uint32_t group) { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(EPI_GROUP_WARPS * 32) : "memory"); }

// load CW fp32 accumulator columns of this warp's 32 TMEM lanes, as CW / 2 pairs
template <int CW>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float2 (&f)[CW / 2]) {
    if constexpr (CW == 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
    } else {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr) : "memory");
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
    }
}

// fp16 -> e5m2 pairs (split8 converter warps)
__device__ __forceinline__ uint32_t cvt_e5m2x4(uint32_t h01, uint32_t h23) {
    uint16_t a, b;
    asm("cvt.rn.satfinite.e5m2x2.f16x2 %0, %1;" : "=h"(a) : "r"(h01));
    asm("cvt.rn.satfinite.e5m2x2.f16x2 %0, %1;" : "=h"(b) : "r"(h23));
    return (uint32_t)a | ((uint32_t)b << 16);
}
// K-major operand tile with 64-byte rows (64 one-byte elements), 64B swizzle, 8-row atoms of 512 bytes
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
// instruction descriptor, kind::f8f6f4: D = F32, A = E5M2, B = E4M3, both K-major
__device__ __host__ constexpr uint32_t umma_idesc_f8(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (0u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
template <int CTA_GROUP>
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CTA_GROUP == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    }
}

template <int CG, int BN, int EPI, int SPLIT, int EG>
__global__ void __launch_bounds__((GemmCfg<CG, BN, EPI, SPLIT, EG>::THREADS), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB2,
                    const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                    const __grid_constant__ CUtensorMap tmAux, const GemmParams p) {
    using Cfg = GemmCfg<CG, BN, EPI, SPLIT, EG>;
    using T = EpiTraits<EPI>;
    constexpr int S = Cfg::STAGES;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t sA = smem_base;
    const uint32_t sB = sA + S * Cfg::A_STRIDE;             // sA: per stage A tile [, e5m2 copy]; sB: per stage B_hi tile [, B_lo tile]
    const uint32_t sEpi = sB + S * Cfg::B_STRIDE;
    const uint32_t sBar = sEpi + Cfg::EPI_TOTAL;
    auto full_bar = [&](int i) { return sBar + 8u * i; };
    auto empty_bar = [&](int i) { return sBar + 8u * (S + i); };
    auto tfull_bar = [&](int i) { return sBar + 8u * (2 * S + i); };
    auto tempty_bar = [&](int i) { return sBar + 8u * (2 * S + 2 + i); };
    auto aux_bar = [&](int i) { return sBar + 8u * (2 * S + 4 + i); };      // [group][slot]
    auto afull_bar = [&](int i) { return sBar + 8u * (2 * S + 8 + i); };    // split8: this CTA's A tile of stage i has landed (local)
    auto conv_bar = [&](int i) { return sBar + 8u * (3 * S + 8 + i); };     // split8: both CTAs' e5m2 copies of stage i are written (leader)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + (sBar - smem_base) + 8 * (4 * S + 8));
    static_assert(8 * (4 * S + 8) + 4 <= Cfg::BAR_BYTES, "barrier block");

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;

    if (CG == 2) cluster_sync_all();   // both CTAs of the pair are resident before the paired TMEM allocation

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (SPLIT) tma_prefetch_desc(&tmB2);
        tma_prefetch_desc(&tmO0);
        if (T::O1) tma_prefetch_desc(&tmO1);
        if (T::AUX) tma_prefetch_desc(&tmAux);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < S; ++i) {
                mbar_init(full_bar(i), CG);     // producer arrivals (leader + peer), tx bytes from both CTAs
                mbar_init(empty_bar(i), 1);     // one tcgen05.commit
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(tfull_bar(i), 1);             // one tcgen05.commit
                mbar_init(tempty_bar(i), CG * EPI_WARPS);          // one elected lane per epilogue warp of the pair
            }
            for (int i = 0; i < 4; ++i) mbar_init(aux_bar(i), 1);
            if (SPLIT == 2)
                for (int i = 0; i < S; ++i) { mbar_init(afull_bar(i), 1); mbar_init(conv_bar(i), CG * Cfg::CONV_WARPS); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<CG>(smem_u32(tmem_ptr_smem), Cfg::TMEM_COLS);
    }
    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_prologue();     // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail; global memory from here on

    const int num_clusters = gridDim.x / CG;
    const int cluster_id = blockIdx.x / CG;
    const int total_tiles = p.num_m_tiles * p.num_n_tiles;
    const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < total_tiles; t += num_clusters) {
                const int mt = t / p.num_n_tiles, nt = t % p.num_n_tiles;
                const int m_base = (mt * CG + (int)cta_rank) * BLOCK_M;
                const int n_base = nt * BN + (int)cta_rank * (BN / CG);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    if (SPLIT == 2) {
                        // the A tile reports to THIS CTA's barrier (its converter warps wait for it); the B tiles to the leader's
                        mbar_arrive_expect_tx(afull_bar(stage), Cfg::A_STAGE);
                        tma_load_2d<1>(&tmA, afull_bar(stage), sA + stage * Cfg::A_STRIDE, kb * BLOCK_K, m_base);
                        if (leader) mbar_arrive_expect_tx(full_bar(stage), Cfg::B_STRIDE * CG);
                    } else {
                        if (leader) mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE * CG);
                        tma_load_2d<CG>(&tmA, full_bar(stage), sA + stage * Cfg::A_STRIDE, kb * BLOCK_K, m_base);
                    }
                    tma_load_2d<CG>(&tmB, full_bar(stage), sB + stage * Cfg::B_STRIDE, kb * BLOCK_K, n_base);
                    if (SPLIT) tma_load_2d<CG>(&tmB2, full_bar(stage), sB + stage * Cfg::B_STRIDE + Cfg::B_STAGE, kb * BLOCK_K, n_base);
                    if (!leader) mbar_arrive_cluster(full_bar(stage), 0);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (leader CTA of the pair only)
        if (leader) {
            constexpr uint32_t idesc = umma_idesc_f16(BLOCK_M * CG, BN);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int t = cluster_id; t < total_tiles; t += num_clusters, ++iter) {
                const int acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    if (SPLIT == 2) mbar_wait_cluster(conv_bar(stage), phase);   // A tiles landed in both CTAs and their e5m2 copies are written
                    tcgen05_fence_after();
                    if (lane == 0) {
                        const uint64_t da = umma_desc_sw128(sA + stage * Cfg::A_STRIDE);
                        const uint64_t db = umma_desc_sw128(sB + stage * Cfg::B_STRIDE);
                        const int krem = p.K - kb * BLOCK_K;
                        const int nk = krem >= BLOCK_K ? BLOCK_K / 16 : (krem + 15) / 16;   // ragged last k-block (K % 64 != 0)
                        if (SPLIT == 2) {
                            constexpr uint32_t idesc8 = umma_idesc_f8(BLOCK_M * CG, BN);
                            const uint64_t da8 = umma_desc_sw64(sA + stage * Cfg::A_STRIDE + Cfg::A_STAGE);
                            const uint64_t db8 = umma_desc_sw64(sB + stage * Cfg::B_STRIDE + Cfg::B_STAGE);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k) umma_f16<CG>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 32; ++k) umma_f8<CG>(d_tmem, da8 + 2 * k, db8 + 2 * k, idesc8, 1u);    // + A8 * B_lo8^T (K = 32 per MMA)
                        } else {
                            const uint64_t db2 = umma_desc_sw128(sB + stage * Cfg::B_STRIDE + Cfg::B_STAGE);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k) {
                                if (k < nk) {
                                    umma_f16<CG>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                                    if (SPLIT == 1) umma_f16<CG>(d_tmem, da + 2 * k, db2 + 2 * k, idesc, 1u);     // + A * B_lo^T on the same A tile
                                }
                            }
                        }
                        umma_commit<CG>(empty_bar(stage));                 // smem stage free once these MMAs retire
                        if (kb == num_kb - 1) umma_commit<CG>(tfull_bar(acc));   // accumulator complete
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 2 + EPI_WARPS) {
        // ===================================================== split8 converter warps: fp16 A tile -> e5m2 copy, stage by stage
        // task = (row, 16-byte output chunk): 32 bytes of fp16 (two 16-byte chunks of the 128B-swizzled row) -> 16 e5m2 bytes of the
        // 64B-swizzled row; lanes 4 r .. 4 r + 3 cover one row, so a warp reads 1 KB of contiguous rows per pass
        if constexpr (SPLIT == 2) {
            const int ct = (int)(warp - (2 + EPI_WARPS)) * 32 + (int)lane;         // 0 .. 63
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < total_tiles; t += num_clusters) {
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(afull_bar(stage), phase);
                    const uint32_t src = sA + stage * Cfg::A_STRIDE, dst = src + Cfg::A_STAGE;
                    constexpr int NTASK = (BLOCK_M * 4) / (Cfg::CONV_WARPS * 32);
                    uint32_t q[NTASK][8];
#pragma unroll
                    for (int i = 0; i < NTASK; ++i) {       // all loads first: issued one task at a time the ld -> cvt -> st chains run serially
                        const uint32_t task = (uint32_t)(i * Cfg::CONV_WARPS * 32 + ct), row = task >> 2, oc = task & 3;    // (~500 clk per tile: +40 % GEMM time)
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q[i][0]), "=r"(q[i][1]), "=r"(q[i][2]), "=r"(q[i][3]) : "r"(src + sw128_off(row, 2 * oc)));
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q[i][4]), "=r"(q[i][5]), "=r"(q[i][6]), "=r"(q[i][7]) : "r"(src + sw128_off(row, 2 * oc + 1)));
                    }
#pragma unroll
                    for (int i = 0; i < NTASK; ++i) {
                        const uint32_t task = (uint32_t)(i * Cfg::CONV_WARPS * 32 + ct), row = task >> 2, oc = task & 3;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + sw64_off(row, oc)),
                                     "r"(cvt_e5m2x4(q[i][0], q[i][1])), "r"(cvt_e5m2x4(q[i][2], q[i][3])), "r"(cvt_e5m2x4(q[i][4], q[i][5])),
                                     "r"(cvt_e5m2x4(q[i][6], q[i][7])) : "memory");
                    }
                    fence_proxy_async_smem();       // generic-proxy writes -> visible to the tensor core's reads
                    __syncwarp();
                    if (lane == 0) { if (leader) mbar_arrive(conv_bar(stage)); else mbar_arrive_cluster(conv_bar(stage), 0); }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================================================== epilogue: 2 groups x 8 warps, groups take alternate stripes
        constexpr int STRIPE = Cfg::STRIPE, CW = Cfg::CW;
        const uint32_t quarter = warp & 3;              // TMEM lane quarter this warp may access
        const uint32_t ew = warp - 2;
        const uint32_t grp = ew >> 3;                   // epilogue group
        const uint32_t half = (ew & 7) >> 2;            // which half of the stripe's columns
        const uint32_t row = quarter * 32 + lane;       // row of the 128-row tile owned by this thread
        const bool elected = ((ew & 7) == 0 && lane == 0);
        const uint32_t gbase = sEpi + grp * (Cfg::O0_BUF + Cfg::O1_BUF + Cfg::AUX_DEPTH * Cfg::AUX_BUF);
        const uint32_t o0_buf = gbase, aux_base = gbase + Cfg::O0_BUF, o1_buf = gbase + Cfg::O0_BUF + Cfg::AUX_DEPTH * Cfg::AUX_BUF;
        const bool write_o1 = (EPI == EPI_GELU) || (EPI == EPI_F32 && p.has_out1);
        const uint32_t dseed = p.drop_thresh ? drop_seed_resolve(p.drop_seed) : 0u;

        int iter = 0;
        uint32_t aux_it = 0;     // aux stripes consumed by this group so far: slot = aux_it % AUX_DEPTH, parity = (aux_it / AUX_DEPTH) & 1
        for (int t = cluster_id; t < total_tiles; t += num_clusters, ++iter) {
            const int mt = t / p.num_n_tiles, nt = t % p.num_n_tiles;
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int m_base = (mt * CG + (int)cta_rank) * BLOCK_M;
            const int n_base = nt * BN;
            const int n_rem = p.N - n_base;
            const int nstripes = ((n_rem >= BN ? BN : n_rem) + STRIPE - 1) / STRIPE;
            const bool tile_live = m_base < p.M;     // CTA-uniform (the second CTA of a pair can be past the M tail)

            if (T::AUX && tile_live && elected && grp < (uint32_t)EG) {       // this group's first two aux stripes of the tile
#pragma unroll
                for (int q = 0; q < Cfg::AUX_DEPTH; ++q) {
                    const int sq = (int)grp + q * EG;
                    if (sq < nstripes) {
                        const uint32_t slot = (aux_it + q) % Cfg::AUX_DEPTH;
                        mbar_arrive_expect_tx(aux_bar(grp * 2 + slot), Cfg::AUX_BUF);
                        tma_load_2d<1>(&tmAux, aux_bar(grp * 2 + slot), aux_base + slot * Cfg::AUX_BUF, n_base + sq * STRIPE, m_base);
                    }
                }
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tcgen05_fence_after();
            if ((int)grp >= nstripes || grp >= (uint32_t)EG) {      // nothing to do for this group (narrow tile, or an idle group): just release the accumulator
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) { if (CG == 2 && !leader) mbar_arrive_cluster(tempty_bar(acc), 0); else mbar_arrive(tempty_bar(acc)); }
            }

            for (int sidx = grp; sidx < (grp < (uint32_t)EG ? nstripes : 0); sidx += EG) {
                const int n0 = n_base + sidx * STRIPE + (int)half * CW;      // first column of this warp
                float2 v[CW / 2];                                            // this thread's CW accumulator columns, as pairs
                tmem_ld_cols<CW>(tmem_base + ((quarter * 32u) << 16) + acc * BN + sidx * STRIPE + half * CW, v);
                if (sidx + EG >= nstripes) {
                    // this warp's last TMEM read of the accumulator (tcgen05.wait::ld is warp-wide): hand the buffer back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (CG == 2 && !leader) mbar_arrive_cluster(tempty_bar(acc), 0); else mbar_arrive(tempty_bar(acc)); }
                }
                if (!tile_live) continue;
                const uint32_t e0 = (uint32_t)(m_base + (int)row) * (uint32_t)p.N + (uint32_t)n0;     // dropout counter of column n0
                if (SPLIT == 2) {               // undo the 2^s pre-scale of the weight pair
                    const float2 sc = make_float2(p.acc_scale, p.acc_scale);
#pragma unroll
                    for (int j = 0; j < CW / 2; ++j) v[j] = mul2(v[j], sc);
                }

                if (p.bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < CW; j += 4) {
                        if (n0 + j < p.N) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                            v[j / 2] = add2(v[j / 2], make_float2(b4.x, b4.y));
                            v[j / 2 + 1] = add2(v[j / 2 + 1], make_float2(b4.z, b4.w));
                        }
                    }
                }
                if (EPI == EPI_RES_F32) {      // out = Dropout(acc + bias) + residual   (to_out / fc2 output, then the Residual add)
                    const uint32_t aux_buf = aux_base + (aux_it % Cfg::AUX_DEPTH) * Cfg::AUX_BUF;
                    mbar_wait(aux_bar(grp * 2 + (aux_it % Cfg::AUX_DEPTH)), (aux_it / Cfg::AUX_DEPTH) & 1);
                    ++aux_it;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {       // fp32 residual stripe [128 x 32]: this warp's 16 columns = chunks half*4 .. +3
                        float4 r;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(aux_buf + sw128_off(row, half * 4 + j)));
                        if (p.drop_thresh) {
                            v[2 * j] = fma2(v[2 * j], drop_pair2(e0 + 4 * j, dseed, p.drop_thresh, p.drop_scale), make_float2(r.x, r.y));
                            v[2 * j + 1] = fma2(v[2 * j + 1], drop_pair2(e0 + 4 * j + 2, dseed, p.drop_thresh, p.drop_scale), make_float2(r.z, r.w));
                        } else {
                            v[2 * j] = add2(v[2 * j], make_float2(r.x, r.y));
                            v[2 * j + 1] = add2(v[2 * j + 1], make_float2(r.z, r.w));
                        }
                    }
                }
                if (EPI == EPI_GELU_BWD) {     // out = acc * aux, aux = keep-scale * mask * gelu'(h) as written by EPI_GELU (fp16 stripe [128 x 64])
                    const uint32_t aux_buf = aux_base + (aux_it % Cfg::AUX_DEPTH) * Cfg::AUX_BUF;
                    mbar_wait(aux_bar(grp * 2 + (aux_it % Cfg::AUX_DEPTH)), (aux_it / Cfg::AUX_DEPTH) & 1);
                    ++aux_it;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t h0, h1, h2, h3;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(h0), "=r"(h1), "=r"(h2), "=r"(h3) : "r"(aux_buf + sw128_off(row, half * 4 + j)));
                        v[4 * j] = mul2(v[4 * j], unpack_half2(h0));
                        v[4 * j + 1] = mul2(v[4 * j + 1], unpack_half2(h1));
                        v[4 * j + 2] = mul2(v[4 * j + 2], unpack_half2(h2));
                        v[4 * j + 3] = mul2(v[4 * j + 3], unpack_half2(h3));
                    }
                }
                if (EPI == EPI_F16_ROWDOT) {   // dot product of this thread's 32 accumulator columns with the aux stripe (fp16 [128 x 64]) -> rowdot part `half`
                    const uint32_t aux_buf = aux_base + (aux_it % Cfg::AUX_DEPTH) * Cfg::AUX_BUF;
                    mbar_wait(aux_bar(grp * 2 + (aux_it % Cfg::AUX_DEPTH)), (aux_it / Cfg::AUX_DEPTH) & 1);
                    ++aux_it;
                    float2 d = make_float2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t h0, h1, h2, h3;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(h0), "=r"(h1), "=r"(h2), "=r"(h3) : "r"(aux_buf + sw128_off(row, half * 4 + j)));
                        d = fma2(v[4 * j], unpack_half2(h0), d);
                        d = fma2(v[4 * j + 1], unpack_half2(h1), d);
                        d = fma2(v[4 * j + 2], unpack_half2(h2), d);
                        d = fma2(v[4 * j + 3], unpack_half2(h3), d);
                    }
                    const int r = m_base + (int)row;
                    if (r < p.M && n0 < p.N) {
                        const int period = p.period > 0 ? p.period : p.M, nseg = p.N >> 6;
                        p.rowdot[(size_t)half * p.M * nseg + ((size_t)(r / period) * nseg + (n0 >> 6)) * period + (r % period)] = d.x + d.y;
                    }
                }
                if (EPI == EPI_PERIODIC_F32) {
                    const int r = m_base + (int)row;
                    if (r < p.M) {
                        const float* trow = p.table + (size_t)(r % p.period) * p.ld_table + n0;
#pragma unroll
                        for (int j = 0; j < CW; j += 4) {
                            if (n0 + j < p.N) {
                                const float4 t4 = __ldg(reinterpret_cast<const float4*>(trow + j));
                                v[j / 2] = add2(v[j / 2], make_float2(t4.x, t4.y));
                                v[j / 2 + 1] = add2(v[j / 2 + 1], make_float2(t4.z, t4.w));
                            }
                        }
                    }
                    if (p.drop_thresh) {                            // emb_dropout after the pos-embedding add
#pragma unroll
                        for (int j = 0; j < CW; j += 2) v[j / 2] = mul2(v[j / 2], drop_pair2(e0 + j, dseed, p.drop_thresh, p.drop_scale));
                    }
                }
                uint32_t pk0[CW / 2], pk1[CW / 2];      // packed half2 outputs (fp16 epilogues)
                if (EPI == EPI_GELU) {         // out1 = Dropout(gelu(h)), out0 = d out1 / d h   (h = acc + bias never leaves the SM)
                    // Two branch-free loops: a per-pair `if (drop_thresh)` puts every pair in its own basic block and the rcp -> polynomial
                    // -> ex2 -> combine chains (11 dependent steps) of the 16 pairs run strictly one after the other (SASS of round 1's
                    // visit n: stall_wait dominated the epilogue).  In one block ptxas interleaves independent pairs.
#pragma unroll
                    for (int j = 0; j < CW / 2; ++j) {
                        float2 g, gp;
                        gelu_pair(v[j], p.gelu, g, gp);
                        pk0[j] = pack_half2(gp.x, gp.y);
                        pk1[j] = pack_half2(g.x, g.y);
                    }
                    if (p.drop_thresh) {
                        const uint32_t thr2 = p.drop_thresh | (p.drop_thresh << 16);
#pragma unroll
                        for (int j = 0; j < CW / 2; ++j) {
                            const uint32_t keep = drop_keep_mask2(e0 + 2 * j, dseed, thr2);
                            pk0[j] &= keep;
                            pk1[j] &= keep;
                        }
                    }
                } else if (T::O0 == 2) {
#pragma unroll
                    for (int j = 0; j < CW / 2; ++j) pk0[j] = pack_half2(v[j].x, v[j].y);
                } else if (T::O1) {            // fp16 copy of an fp32 stripe
#pragma unroll
                    for (int j = 0; j < CW / 2; ++j) pk1[j] = pack_half2(v[j].x, v[j].y);
                }

                // staging reuse: this group's previous TMA store must have finished reading smem; every thread of the group
                // has consumed the aux stripe by the time it reaches the barrier
                if (elected) tma_store_wait_read<0>();
                epi_bar_sync(grp);
                if (T::AUX && elected && sidx + Cfg::AUX_DEPTH * EG < nstripes) {     // refill the slot just consumed: two stripes ahead
                    const uint32_t slot = (aux_it - 1) % Cfg::AUX_DEPTH;      // (aux_it was advanced when this stripe's aux was consumed)
                    mbar_arrive_expect_tx(aux_bar(grp * 2 + slot), Cfg::AUX_BUF);
                    tma_load_2d<1>(&tmAux, aux_bar(grp * 2 + slot), aux_base + slot * Cfg::AUX_BUF, n_base + (sidx + Cfg::AUX_DEPTH * EG) * STRIPE, m_base);
                }

                if (T::O0 == 4) {   // fp32 [128 x 32] stripe: this warp's 16 columns = 16-byte chunks half*4 .. +3
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
                                     ::"r"(o0_buf + sw128_off(row, half * 4 + j)), "f"(v[2 * j].x), "f"(v[2 * j].y), "f"(v[2 * j + 1].x), "f"(v[2 * j + 1].y) : "memory");
                } else {            // fp16 [128 x 64] stripe: this warp's 32 columns = chunks half*4 .. +3
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                     ::"r"(o0_buf + sw128_off(row, half * 4 + j)), "r"(pk0[4 * j]), "r"(pk0[4 * j + 1]), "r"(pk0[4 * j + 2]), "r"(pk0[4 * j + 3]) : "memory");
                }
                if (T::O1 && write_o1) {
                    if (EPI == EPI_GELU) {      // fp16 [128 x 64]
#pragma unroll
                        for (int j = 0; j < CW / 8; ++j)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                         ::"r"(o1_buf + sw128_off(row, half * 4 + j)), "r"(pk1[4 * j]), "r"(pk1[4 * j + 1]), "r"(pk1[4 * j + 2]), "r"(pk1[4 * j + 3]) : "memory");
                    } else {                    // fp16 copy of an fp32 stripe, [128 x 32] = 64-byte rows: this warp's 16 columns = chunks half*2 .. +1
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                         ::"r"(o1_buf + sw64_off(row, half * 2 + j)), "r"(pk1[4 * j]), "r"(pk1[4 * j + 1]), "r"(pk1[4 * j + 2]), "r"(pk1[4 * j + 3]) : "memory");
                    }
                }
                fence_proxy_async_smem();
                epi_bar_sync(grp);
                if (elected) {
                    tma_store_2d(&tmO0, o0_buf, n_base + sidx * STRIPE, m_base);
                    if (T::O1 && write_o1) tma_store_2d(&tmO1, o1_buf, n_base + sidx * STRIPE, m_base);
                    tma_store_commit();
                }
            }
        }
        if (elected) tma_store_wait_all();
    }

    // ===================================================== teardown
    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------------------------- host side

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_encode_once;

static int load_encode() {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    });
    return g_encode ? 0 : -1;
}

// 2-D row-major tensor [rows, cols] with leading dimension ld (elements); box = [box_rows, box_cols]
int make_tmap_2d(CUtensorMap* map, const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int64_t ld,
                 int box_rows, int box_cols) {
    GSL_REQUIRE(load_encode() == 0, "cuTensorMapEncodeTiled entry point not available");
    GSL_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base pointer must be 16-byte aligned (%p)", ptr);
    GSL_REQUIRE((ld * elem_bytes) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (ld=%lld)", (long long)ld);
    const int inner_bytes = box_cols * elem_bytes;
    CUtensorMapSwizzle sw;
    if (inner_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
    else if (inner_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
    else if (inner_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
    else { set_last_error("unsupported TMA box width %d bytes", inner_bytes); return -1; }
    const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld * elem_bytes)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GSL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r,
                (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols);
    return 0;
}

// 3-D view [B, N, cols] of a row-major [B*N, ld] fp16 matrix; box = [1, npad, 64 cols] with 128B swizzle.  Rows in
// [N, npad) are out of bounds in dimension 1 and arrive as zeros (attention slabs).
int make_tmap_qkv(CUtensorMap* map, const void* ptr, int64_t ld, int B, int N, int cols, int npad) {
    GSL_REQUIRE(load_encode() == 0, "cuTensorMapEncodeTiled entry point not available");
    GSL_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0, "attention TMA: pointer / pitch must be 16-byte aligned");
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)(ld * 2), (cuuint64_t)((int64_t)N * ld * 2)};
    cuuint32_t box[3] = {64, (cuuint32_t)npad, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GSL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (attention) failed (%d)", (int)r);
    return 0;
}

int device_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

template <int CG, int BN, int EPI, int SPLIT, int EG = EPI_GROUPS>
static int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
    using Cfg = GemmCfg<CG, BN, EPI, SPLIT, EG>;
    using T = EpiTraits<EPI>;
    CUtensorMap tmA, tmB, tmB2, tmO0, tmO1, tmAux;
    int rc;
    if ((rc = make_tmap_2d(&tmA, a.A, 2, a.M, a.K, a.lda, BLOCK_M, BLOCK_K))) return rc;
    if ((rc = make_tmap_2d(&tmB, a.B, 2, a.N, a.K, a.ldb, BN / CG, BLOCK_K))) return rc;
    if (SPLIT == 2) { if ((rc = make_tmap_2d(&tmB2, a.B_lo8, 1, a.N, a.K, a.ldb, BN / CG, BLOCK_K))) return rc; }
    else if (SPLIT == 1) { if ((rc = make_tmap_2d(&tmB2, a.B_lo, 2, a.N, a.K, a.ldb, BN / CG, BLOCK_K))) return rc; } else tmB2 = tmB;
    if ((rc = make_tmap_2d(&tmO0, a.out0, T::O0, a.M, a.N, a.ld0, BLOCK_M, Cfg::STRIPE))) return rc;
    const bool has_o1 = (EPI == EPI_GELU) || (T::O1 && a.out1 != nullptr);
    if (has_o1) { if ((rc = make_tmap_2d(&tmO1, a.out1, 2, a.M, a.N, a.ld1, BLOCK_M, Cfg::STRIPE))) return rc; } else tmO1 = tmO0;
    if (T::AUX) { if ((rc = make_tmap_2d(&tmAux, a.aux, T::AUX, a.M, a.N, a.ldaux, BLOCK_M, Cfg::STRIPE))) return rc; } else tmAux = tmO0;

    GemmParams p;
    p.M = (int)a.M; p.N = (int)a.N; p.K = (int)a.K;
    p.num_m_tiles = (int)((a.M + BLOCK_M * CG - 1) / (BLOCK_M * CG));
    p.num_n_tiles = (int)((a.N + BN - 1) / BN);
    p.bias = a.bias;
    p.table = (EPI == EPI_PERIODIC_F32) ? reinterpret_cast<const float*>(a.aux) : nullptr;
    p.period = (int)a.aux_period; p.ld_table = (int)a.ldaux;
    p.has_out1 = has_o1 ? 1 : 0;
    p.rowdot = a.rowdot;
    p.acc_scale = SPLIT == 2 ? ldexpf(1.0f, -a.lo8_shift) : 1.0f;
    p.drop_thresh = drop_thresh15(a.drop_p);
    p.drop_seed = a.drop_seed;
    p.drop_scale = a.drop_p > 0.f ? 1.0f / (1.0f - a.drop_p) : 1.0f;
    p.gelu = make_gelu_consts(p.drop_thresh ? p.drop_scale : 1.0f);

    auto kern = gemm_tcgen05_kernel<CG, BN, EPI, SPLIT, EG>;
    static bool attr_set = false;
    if (!attr_set) {
        GSL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int sms = device_sm_count();
    const int total = p.num_m_tiles * p.num_n_tiles;
    int clusters = sms / CG;
    if (clusters > total) clusters = total;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CG);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    GSL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmB2, tmO0, tmO1, tmAux, p));
    GSL_COUNT_LAUNCH(1);
    return 0;
}

template <int CG, int BN, int SPLIT>
static int dispatch_epi(const GemmArgs& a, cudaStream_t s) {
    switch (a.epi) {
        case EPI_F16: return launch_gemm<CG, BN, EPI_F16, SPLIT>(a, s);
        case EPI_F32: return launch_gemm<CG, BN, EPI_F32, SPLIT>(a, s);
        case EPI_GELU: return launch_gemm<CG, BN, EPI_GELU, SPLIT>(a, s);
        // split8 with a long K loop: one working epilogue group (its staging alone leaves room for the ring's 4th stage, see GemmCfg)
        case EPI_GELU_BWD: if (SPLIT == 2 && a.K >= 512) return launch_gemm<CG, BN, EPI_GELU_BWD, SPLIT, 1>(a, s);
                           return launch_gemm<CG, BN, EPI_GELU_BWD, SPLIT>(a, s);
        case EPI_RES_F32: if (SPLIT == 2 && a.K >= 1024) return launch_gemm<CG, BN, EPI_RES_F32, SPLIT, 1>(a, s);
                          return launch_gemm<CG, BN, EPI_RES_F32, SPLIT>(a, s);
        case EPI_PERIODIC_F32: return launch_gemm<CG, BN, EPI_PERIODIC_F32, SPLIT>(a, s);
        case EPI_F16_ROWDOT: return launch_gemm<CG, BN, EPI_F16_ROWDOT, SPLIT>(a, s);
        default: set_last_error("unknown GEMM epilogue %d", a.epi); return -1;
    }
}

static int g_default_cta_group = 2;
void gemm_set_default_cta_group(int cg) { g_default_cta_group = (cg == 1) ? 1 : 2; }

int gemm_f16(const GemmArgs& a, cudaStream_t stream) {
    GSL_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "empty GEMM %lldx%lldx%lld", (long long)a.M, (long long)a.N, (long long)a.K);
    GSL_REQUIRE(a.K % 16 == 0 && a.N % 8 == 0, "GEMM needs K %% 16 == 0 and N %% 8 == 0 (N=%lld K=%lld)", (long long)a.N, (long long)a.K);
    GSL_REQUIRE(a.A && a.B && a.out0, "null GEMM operand");
    GSL_REQUIRE(a.drop_p >= 0.f && a.drop_p < 1.f && (a.drop_p == 0.f || a.M * a.N < (int64_t)4294967296LL), "bad dropout p / tensor too large for the 32-bit mask counter");
    if (a.epi == EPI_GELU) GSL_REQUIRE(a.out1 != nullptr, "EPI_GELU needs out1");
    if (a.epi == EPI_GELU_BWD || a.epi == EPI_RES_F32 || a.epi == EPI_PERIODIC_F32 || a.epi == EPI_F16_ROWDOT) GSL_REQUIRE(a.aux != nullptr, "epilogue %d needs aux", a.epi);
    if (a.epi == EPI_F16_ROWDOT) GSL_REQUIRE(a.rowdot != nullptr && a.N % 64 == 0 && (a.aux_period == 0 || a.M % a.aux_period == 0), "EPI_F16_ROWDOT needs rowdot, N %% 64 == 0 and M %% aux_period == 0");
    if (a.epi == EPI_PERIODIC_F32) GSL_REQUIRE(a.aux_period > 0, "EPI_PERIODIC_F32 needs aux_period > 0");
    const int cg = a.B_lo8 ? 2 : a.cta_group ? a.cta_group : g_default_cta_group;      // (split8 exists for cta_group::2 only)
    const int bn = a.block_n ? a.block_n : ((a.N % 256 == 0 || a.N > 1024) ? 256 : 128);
    if (a.B_lo8 != nullptr) {
        GSL_REQUIRE(a.B_lo == nullptr, "pass either B_lo (fp16 residual) or B_lo8 (e4m3 residual of the 2^shift-scaled weight), not both");
        GSL_REQUIRE(a.K % 64 == 0 && a.ldb % 16 == 0 && a.lo8_shift >= 0 && a.lo8_shift <= 24,
                    "split8 GEMM needs K %% 64 == 0, ldb %% 16 == 0, shift in [0, 24] (K=%lld ldb=%lld shift=%d)", (long long)a.K, (long long)a.ldb, a.lo8_shift);
        return bn == 256 ? dispatch_epi<2, 256, 2>(a, stream) : dispatch_epi<2, 128, 2>(a, stream);
    }
    if (a.B_lo != nullptr) {
        if (cg == 2) return bn == 256 ? dispatch_epi<2, 256, 1>(a, stream) : dispatch_epi<2, 128, 1>(a, stream);
        return bn == 256 ? dispatch_epi<1, 256, 1>(a, stream) : dispatch_epi<1, 128, 1>(a, stream);
    }
    if (cg == 2) return bn == 256 ? dispatch_epi<2, 256, 0>(a, stream) : dispatch_epi<2, 128, 0>(a, stream);
    return bn == 256 ? dispatch_epi<1, 256, 0>(a, stream) : dispatch_epi<1, 128, 0>(a, stream);
}

}  // namespace gsl
