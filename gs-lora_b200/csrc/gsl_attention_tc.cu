// gslora-b200: attention forward on tcgen05 tensor cores (Attention.forward, vit_pytorch_face/vit_face.py:358-379).
//
// One persistent CTA per SM walks over (image, head) pairs.  Per pair the whole problem is two MMA row tiles:
//   S_i = Q_i K^T      tcgen05.mma  M=128, N=npad (<= 208), K=64      accumulators in TMEM (208 fp32 columns per tile)
//   softmax            128 threads per tile, ONE THREAD PER QUERY ROW (TMEM lane == row): two sweeps over the row with
//                      tcgen05.ld (max, then exp2 / sum), no shuffles; P is written as fp16 into a 128B-swizzled
//                      K-major shared-memory tile = the A operand of the next MMA
//   O_i = P_i V        tcgen05.mma  M=128, N=64, K=npad; V is consumed straight from its TMA slab as an MN-major B operand
//                      (no transpose anywhere); O_i overwrites the first 64 columns of S_i's TMEM region
//   epilogue           the row's thread scales by 1/sum and stores 128 contiguous bytes of O, plus the row LSE
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocation), warps 2-5 / 6-9 the two softmax groups.
// mbarriers hand TMEM regions and smem slabs back and forth; while group 0 exponentiates tile 0 the tensor core
// computes S_1, and the next pair's Q/K/V loads are issued as soon as the MMAs that read them retire.
#include "gsl_common.cuh"
#include <cuda.h>
#include "gsl_kernels.h"

namespace gsl {

int make_tmap_qkv(CUtensorMap* map, const void* ptr, int64_t ld, int B, int N, int cols, int npad);

static constexpr int ATC_THREADS = 320;
static constexpr int ATC_MAX_TOKENS = 208;

// K-major / MN-major smem descriptors for 128-byte rows with 128B swizzle (8-row atoms of 1024 bytes)
__device__ __forceinline__ uint64_t atc_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    // LBO (bits 16-29) stays 0: unused, every operand here is a single 64-element atom wide along its leading dimension
    d |= (uint64_t)(1024 >> 4) << 32;         // SBO: 8 rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
    return d;
}
__device__ __host__ constexpr uint32_t atc_idesc(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void atc_tma_slab(const void* desc, uint32_t bar, uint32_t dst, int col, int b) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(col), "r"(0), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, __half* __restrict__ out, int64_t ldo, float* __restrict__ lse,
                        int B, int N, int heads, float scale) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int npad = (N + 15) & ~15;
    const int ntiles = npad > 128 ? 2 : 1;
    const int D = heads * 64;
    const uint32_t slab = npad * 128;
    const uint32_t sQ = smem_u32(smem), sK = sQ + slab, sV = sK + slab;
    const uint32_t sP = (sV + slab + 1023) & ~1023u;            // two P tiles of 4 x [128 x 64] fp16 blocks (64 KB each)
    const uint32_t bars = sP + 2 * 65536;
    const uint32_t b_qk_full = bars, b_v_full = bars + 8, b_qk_free = bars + 16, b_v_free = bars + 24;
    auto b_s_full = [&](int i) { return bars + 32 + 8 * i; };
    auto b_p_ready = [&](int i) { return bars + 48 + 8 * i; };
    auto b_o_full = [&](int i) { return bars + 64 + 8 * i; };
    auto b_s_free = [&](int i) { return bars + 80 + 8 * i; };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + (bars - sQ) + 96);

    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const int nwork = B * heads;

    if (warp == 1) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQKV);
            mbar_init(b_qk_full, 1); mbar_init(b_v_full, 1); mbar_init(b_qk_free, 1); mbar_init(b_v_free, 1);
            for (int i = 0; i < 2; ++i) {
                mbar_init(b_s_full(i), 1); mbar_init(b_p_ready(i), 128); mbar_init(b_o_full(i), 1); mbar_init(b_s_free(i), 128);
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

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int it = 0;
            for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
                const int wb = w / heads, wh = w % heads;
                if (it > 0) mbar_wait(b_qk_free, (it - 1) & 1);
                mbar_arrive_expect_tx(b_qk_full, 2 * slab);
                atc_tma_slab(&tmQKV, b_qk_full, sQ, wh * 64, wb);
                atc_tma_slab(&tmQKV, b_qk_full, sK, D + wh * 64, wb);
                if (it > 0) mbar_wait(b_v_free, (it - 1) & 1);
                mbar_arrive_expect_tx(b_v_full, slab);
                atc_tma_slab(&tmQKV, b_v_full, sV, 2 * D + wh * 64, wb);
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        const uint32_t idesc_s = atc_idesc(128, (uint32_t)npad, 0, 0);
        constexpr uint32_t idesc_o = atc_idesc(128, 64, 0, 1);      // B = V slab, MN-major
        int it = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
            const uint32_t ph = it & 1;
            mbar_wait(b_qk_full, ph);
            for (int i = 0; i < ntiles; ++i) {
                mbar_wait(b_s_free(i), ph ^ 1);
                tcgen05_fence_after();
                if (lane == 0) {
                    const uint64_t da = atc_desc(sQ + i * 16384), db = atc_desc(sK);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16<1>(tmem_base + i * 256, da + 2 * k, db + 2 * k, idesc_s, k != 0);
                    umma_commit<1>(b_s_full(i));
                }
                __syncwarp();
            }
            if (lane == 0) umma_commit<1>(b_qk_free);
            __syncwarp();
            mbar_wait(b_v_full, ph);
            for (int i = 0; i < ntiles; ++i) {
                mbar_wait(b_p_ready(i), ph);
                tcgen05_fence_after();
                if (lane == 0) {
                    const int nk = npad / 16;
                    for (int j = 0; j < nk; ++j) {
                        const uint64_t da = atc_desc(sP + i * 65536 + (j >> 2) * 16384 + (j & 3) * 32);
                        const uint64_t db = atc_desc(sV + j * 2048);
                        umma_f16<1>(tmem_base + i * 256, da, db, idesc_o, j != 0);
                    }
                    umma_commit<1>(b_o_full(i));
                }
                __syncwarp();
            }
            if (lane == 0) umma_commit<1>(b_v_free);
            __syncwarp();
        }
    } else {
        // ===================================================== softmax / epilogue groups: one thread per query row
        const int wg = (warp - 2) >> 2;
        const uint32_t quarter = warp & 3;
        const int row_local = quarter * 32 + lane;
        const int row = wg * 128 + row_local;
        const float sl2 = scale * 1.4426950408889634f;
        const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + wg * 256;
        const uint32_t sPt = sP + wg * 65536;
        if (wg < ntiles) {
            int it = 0;
            for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                const int b = w / heads, h = w % heads;
                mbar_wait(b_s_full(wg), ph);
                tcgen05_fence_after();
                // sweep 1: row max over the valid key columns
                float mx = -INFINITY;
                for (int c0 = 0; c0 < npad; c0 += 32) {
                    if (npad - c0 >= 32) {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (c0 + j < N) mx = fmaxf(mx, __uint_as_float(v[j]));
                    } else {
                        uint32_t v[16];
                        tmem_ld16(taddr + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (c0 + j < N) mx = fmaxf(mx, __uint_as_float(v[j]));
                    }
                }
                // sweep 2: p = exp(scale * (s - max)), fp16 -> swizzled K-major tile (64-key blocks of [128 x 128 B])
                const float off = mx * sl2;
                float sum = 0.f;
                for (int c0 = 0; c0 < npad; c0 += 32) {
                    float pv[32];
                    const int cnt = npad - c0 >= 32 ? 32 : 16;
                    if (cnt == 32) {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) pv[j] = (c0 + j < N) ? ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -off)) : 0.f;
                    } else {
                        uint32_t v[16];
                        tmem_ld16(taddr + c0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) pv[j] = (c0 + j < N) ? ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -off)) : 0.f;
#pragma unroll
                        for (int j = 16; j < 32; ++j) pv[j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum += pv[j];
                    const uint32_t blk = sPt + (c0 >> 6) * 16384;
                    const int chunk0 = (c0 & 63) >> 3;       // first 16-byte chunk inside the 64-key block
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (q * 8 < cnt)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                         ::"r"(blk + sw128_off(row_local, chunk0 + q)), "r"(pack_half2(pv[8 * q], pv[8 * q + 1])),
                                           "r"(pack_half2(pv[8 * q + 2], pv[8 * q + 3])), "r"(pack_half2(pv[8 * q + 4], pv[8 * q + 5])),
                                           "r"(pack_half2(pv[8 * q + 6], pv[8 * q + 7])) : "memory");
                    }
                }
                fence_proxy_async_smem();          // P tile -> visible to the tensor core's async proxy
                tcgen05_fence_before();
                mbar_arrive(b_p_ready(wg));
                // epilogue: O row = (P V) / sum
                mbar_wait(b_o_full(wg), ph);
                tcgen05_fence_after();
                uint32_t o0[32], o1[32];
                tmem_ld_32x32(taddr, o0);
                tmem_ld_32x32(taddr + 32, o1);
                tmem_ld_wait();
                tcgen05_fence_before();
                mbar_arrive(b_s_free(wg));         // TMEM region may be overwritten by the next pair's S
                if (row < N) {
                    const float inv = 1.0f / sum;
                    uint4* dst = reinterpret_cast<uint4*>(out + ((int64_t)b * N + row) * ldo + h * 64);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 u;
                        u.x = pack_half2(__uint_as_float(o0[8 * q]) * inv, __uint_as_float(o0[8 * q + 1]) * inv);
                        u.y = pack_half2(__uint_as_float(o0[8 * q + 2]) * inv, __uint_as_float(o0[8 * q + 3]) * inv);
                        u.z = pack_half2(__uint_as_float(o0[8 * q + 4]) * inv, __uint_as_float(o0[8 * q + 5]) * inv);
                        u.w = pack_half2(__uint_as_float(o0[8 * q + 6]) * inv, __uint_as_float(o0[8 * q + 7]) * inv);
                        dst[q] = u;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 u;
                        u.x = pack_half2(__uint_as_float(o1[8 * q]) * inv, __uint_as_float(o1[8 * q + 1]) * inv);
                        u.y = pack_half2(__uint_as_float(o1[8 * q + 2]) * inv, __uint_as_float(o1[8 * q + 3]) * inv);
                        u.z = pack_half2(__uint_as_float(o1[8 * q + 4]) * inv, __uint_as_float(o1[8 * q + 5]) * inv);
                        u.w = pack_half2(__uint_as_float(o1[8 * q + 6]) * inv, __uint_as_float(o1[8 * q + 7]) * inv);
                        dst[4 + q] = u;
                    }
                    if (lse != nullptr) lse[((int64_t)b * heads + h) * N + row] = mx * scale + __logf(sum);
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

int attention_fwd_tc(const __half* qkv, int64_t ld, __half* out, int64_t ldo, float* lse, int B, int N, int heads, float scale, cudaStream_t s) {
    GSL_REQUIRE(N >= 1 && N <= ATC_MAX_TOKENS, "attention: tokens=%d outside [1, %d]", N, ATC_MAX_TOKENS);
    GSL_REQUIRE(ld % 8 == 0 && ldo % 8 == 0, "attention: pitches must be multiples of 8 halves");
    GSL_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "attention: output must be 16-byte aligned");
    const int npad = (N + 15) & ~15;
    CUtensorMap tm;
    int rc = make_tmap_qkv(&tm, qkv, ld, B, N, 3 * heads * 64, npad);
    if (rc) return rc;
    const int smem = 1024 + 3 * ATC_MAX_TOKENS * 128 + 1024 + 2 * 65536 + 256;
    static bool attr = false;
    if (!attr) {
        GSL_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    const int sms = device_sm_count();
    const int nwork = B * heads;
    attention_fwd_tc_kernel<<<nwork < sms ? nwork : sms, ATC_THREADS, smem, s>>>(tm, out, ldo, lse, B, N, heads, scale);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}


// ------------------------------------------------------------------------------------------------ backward (tcgen05)
// Same roles, 2 x 2 blocking of the (query, key) plane into 128 x 128 blocks (N <= 256 -> at most 4 blocks per pair):
//   for key tile kt:  for query tile qt:
//       S  = Q_qt K_kt^T ,  dP = dO_qt V_kt^T                      (TMEM columns [0,128) and [128,256))
//       workers (one thread per query row and 32-column group): P = exp(scale*S - LSE), dS = P * (dP - delta)
//                -> fp16 tiles [128 q x 128 keys] in shared memory (two 64-key K-major blocks each)
//       dQ_qt += dS K_kt        (A = dS tile K-major,        B = K slab rows as MN-major)      TMEM [256,320) / [320,384)
//       dV_kt += P^T  dO_qt     (A = P  tile read MN-major,   B = dO slab rows as MN-major)     TMEM [448,512)
//       dK_kt += dS^T Q_qt      (A = dS tile read MN-major,   B = Q  slab rows as MN-major)     TMEM [384,448)
//   dK_kt / dV_kt are read out after the inner loop, dQ after the outer loop.  Five matmuls per block, no recompute, no
//   atomics; all transposes are descriptor modes (MN-major operands), nothing is transposed in memory.
static constexpr int ATB_WORKER_WARPS = 16;
static constexpr int ATB_THREADS = 64 + ATB_WORKER_WARPS * 32;

__device__ __forceinline__ uint64_t atc_desc_mn2(uint32_t smem_addr, uint32_t lbo_bytes) {    // MN-major, two 64-element atoms along MN
    return atc_desc(smem_addr) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__device__ __forceinline__ void atb_bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(ATB_WORKER_WARPS * 32) : "memory"); }

__global__ void __launch_bounds__(ATB_THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const __half* __restrict__ out, int64_t ldo,
                        const float* __restrict__ lse, __half* __restrict__ dqkv, int64_t lddqkv, int B, int N, int heads, float scale) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nt = N > 128 ? 2 : 1;                 // 128-row tiles along queries and along keys
    const int D = heads * 64;
    constexpr uint32_t SLAB = 256 * 128;            // TMA box of 256 rows: rows >= N arrive as zeros
    const uint32_t sQ = smem_u32(smem), sK = sQ + SLAB, sV = sK + SLAB, sdO = sV + SLAB;
    const uint32_t sPt = sdO + SLAB, sSt = sPt + 32768;     // P and dS tiles: [128 q x 128 keys] fp16 = 2 blocks of 16 KB
    const uint32_t bars = sSt + 32768;
    const uint32_t b_slabs_full = bars, b_slabs_free = bars + 8, b_sdp_full = bars + 16, b_pds_ready = bars + 24, b_tiles_free = bars + 32,
                   b_dkv_full = bars + 40, b_dkv_free = bars + 48, b_dq_full = bars + 56, b_dq_free = bars + 64;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + (bars - sQ) + 80);
    float* s_stat = reinterpret_cast<float*>(smem + (bars - sQ) + 128);      // [2 parity][2: lse*log2e, delta][256]

    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const int nwork = B * heads;
    constexpr uint32_t NWORK_THREADS = ATB_WORKER_WARPS * 32;

    if (warp == 1) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQKV);
            tma_prefetch_desc(&tmDO);
            mbar_init(b_slabs_full, 1); mbar_init(b_slabs_free, 1); mbar_init(b_sdp_full, 1); mbar_init(b_pds_ready, NWORK_THREADS);
            mbar_init(b_tiles_free, 1); mbar_init(b_dkv_full, 1); mbar_init(b_dkv_free, NWORK_THREADS); mbar_init(b_dq_full, 1);
            mbar_init(b_dq_free, NWORK_THREADS);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc<1>(smem_u32(tmem_ptr_smem), 512);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    constexpr uint32_t T_S = 0, T_DP = 128, T_DQ = 256, T_DK = 384, T_DV = 448;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int it = 0;
            for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
                const int wb = w / heads, wh = w % heads;
                if (it > 0) mbar_wait(b_slabs_free, (it - 1) & 1);
                mbar_arrive_expect_tx(b_slabs_full, 4 * SLAB);
                atc_tma_slab(&tmQKV, b_slabs_full, sQ, wh * 64, wb);
                atc_tma_slab(&tmQKV, b_slabs_full, sK, D + wh * 64, wb);
                atc_tma_slab(&tmQKV, b_slabs_full, sV, 2 * D + wh * 64, wb);
                atc_tma_slab(&tmDO, b_slabs_full, sdO, wh * 64, wb);
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        constexpr uint32_t idesc_s = atc_idesc(128, 128, 0, 0);     // S, dP : A, B K-major
        constexpr uint32_t idesc_dq = atc_idesc(128, 64, 0, 1);     // dQ    : A = dS K-major, B = K rows MN-major
        constexpr uint32_t idesc_dkv = atc_idesc(128, 64, 1, 1);    // dK, dV: A = tile^T (MN-major), B = slab rows MN-major
        int it = 0;
        uint32_t n = 0;      // global block counter
        uint32_t m = 0;      // global key-tile counter
        auto issue_sdp = [&](int qt, int kt) {      // S = Q_qt K_kt^T, dP = dO_qt V_kt^T
            tcgen05_fence_after();
            if (lane == 0) {
                const uint64_t aq = atc_desc(sQ + qt * 16384), bk = atc_desc(sK + kt * 16384);
                const uint64_t ad = atc_desc(sdO + qt * 16384), bv = atc_desc(sV + kt * 16384);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16<1>(tmem_base + T_S, aq + 2 * k, bk + 2 * k, idesc_s, k != 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16<1>(tmem_base + T_DP, ad + 2 * k, bv + 2 * k, idesc_s, k != 0);
                umma_commit<1>(b_sdp_full);
            }
            __syncwarp();
        };
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
            mbar_wait(b_slabs_full, it & 1);
            issue_sdp(0, 0);
            for (int kt = 0; kt < nt; ++kt, ++m) {
                for (int qt = 0; qt < nt; ++qt, ++n) {
                    mbar_wait(b_pds_ready, n & 1);                  // P / dS tiles written, S / dP TMEM consumed
                    // software pipeline: the next block's S / dP go to the tensor core first, so the workers can start on them
                    // while this block's dQ / dV / dK MMAs run
                    if (qt + 1 < nt) issue_sdp(qt + 1, kt);
                    else if (kt + 1 < nt) issue_sdp(0, kt + 1);
                    if (kt == 0 && qt == 0 && it > 0) mbar_wait(b_dq_free, (it - 1) & 1);   // previous pair's dQ has been read out of TMEM
                    if (qt == 0 && m > 0) mbar_wait(b_dkv_free, (m - 1) & 1);               // previous key tile's dK / dV have been read out
                    tcgen05_fence_after();
                    if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {               // dQ_qt += dS K_kt, contraction over 128 keys
                            const uint64_t da = atc_desc(sSt + (j >> 2) * 16384 + (j & 3) * 32);
                            const uint64_t db = atc_desc(sK + kt * 16384 + j * 2048);
                            umma_f16<1>(tmem_base + T_DQ + qt * 64, da, db, idesc_dq, (kt | j) != 0);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {               // dV_kt += P^T dO_qt, contraction over 128 queries
                            const uint64_t da = atc_desc_mn2(sPt + j * 2048, 16384);
                            const uint64_t db = atc_desc(sdO + qt * 16384 + j * 2048);
                            umma_f16<1>(tmem_base + T_DV, da, db, idesc_dkv, (qt | j) != 0);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {               // dK_kt += dS^T Q_qt
                            const uint64_t da = atc_desc_mn2(sSt + j * 2048, 16384);
                            const uint64_t db = atc_desc(sQ + qt * 16384 + j * 2048);
                            umma_f16<1>(tmem_base + T_DK, da, db, idesc_dkv, (qt | j) != 0);
                        }
                        umma_commit<1>(b_tiles_free);
                        if (qt == nt - 1) umma_commit<1>(b_dkv_full);
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) {
                umma_commit<1>(b_dq_full);
                umma_commit<1>(b_slabs_free);
            }
            __syncwarp();
        }
    } else {
        // ===================================================== workers: thread = (row, 32-column group)
        const uint32_t quarter = warp & 3;
        const uint32_t cg = (warp - 2) >> 2;                        // 0..3
        const int rl = quarter * 32 + lane;                         // row inside a 128-row tile
        const float sl2 = scale * 1.4426950408889634f;
        const uint32_t tlane = tmem_base + ((quarter * 32u) << 16);
        const int wt = threadIdx.x - 64;                            // worker thread id 0..511
        int it = 0;
        uint32_t n = 0, m = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
            const int b = w / heads, h = w % heads;
            float* st_lse = s_stat + (it & 1) * 512;
            float* st_del = st_lse + 256;
            mbar_wait(b_slabs_full, it & 1);
            // delta[q] = dO[q,:] . O[q,:]   (dO from the smem slab, O from HBM), LSE pre-scaled by log2(e)
            if (wt < 256) {
                float d = 0.f, l = 0.f;
                if (wt < N) {
                    const uint4* orow = reinterpret_cast<const uint4*>(out + ((int64_t)b * N + wt) * ldo + h * 64);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 ov = __ldg(orow + c);
                        uint32_t d0, d1, d2, d3;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3) : "r"(sdO + sw128_off(wt, c)));
                        const uint32_t oo[4] = {ov.x, ov.y, ov.z, ov.w}, dd[4] = {d0, d1, d2, d3};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 a = unpack_half2(oo[q]), g = unpack_half2(dd[q]);
                            d = fmaf(a.x, g.x, d); d = fmaf(a.y, g.y, d);
                        }
                    }
                    l = lse[((int64_t)b * heads + h) * N + wt] * 1.4426950408889634f;
                }
                st_del[wt] = d;
                st_lse[wt] = l;
            }
            atb_bar_workers();
            for (int kt = 0; kt < nt; ++kt, ++m) {
                for (int qt = 0; qt < nt; ++qt, ++n) {
                    const int q = qt * 128 + rl;
                    const float lq = st_lse[q], dq = st_del[q];
                    mbar_wait(b_sdp_full, n & 1);
                    tcgen05_fence_after();
                    uint32_t sv[32], pv[32];
                    tmem_ld_32x32(tlane + T_S + cg * 32, sv);
                    tmem_ld_32x32(tlane + T_DP + cg * 32, pv);
                    tmem_ld_wait();
                    uint32_t pp[16], ds[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const int key = kt * 128 + cg * 32 + j;
                        const bool v0 = q < N && key < N, v1 = q < N && key + 1 < N;
                        const float p0 = v0 ? ex2_approx(fmaf(__uint_as_float(sv[j]), sl2, -lq)) : 0.f;
                        const float p1 = v1 ? ex2_approx(fmaf(__uint_as_float(sv[j + 1]), sl2, -lq)) : 0.f;
                        pp[j >> 1] = pack_half2(p0, p1);
                        ds[j >> 1] = pack_half2(p0 * (__uint_as_float(pv[j]) - dq), p1 * (__uint_as_float(pv[j + 1]) - dq));
                    }
                    if (n > 0) mbar_wait(b_tiles_free, (n - 1) & 1);            // previous block's MMAs have read the tiles
                    const uint32_t boff = (cg >> 1) * 16384;                        // 64-key block inside the tile
                    const uint32_t c0 = (cg & 1) * 4;                               // first 16-byte chunk of this thread's 32 columns
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sPt + boff + sw128_off(rl, c0 + c)),
                                     "r"(pp[4 * c]), "r"(pp[4 * c + 1]), "r"(pp[4 * c + 2]), "r"(pp[4 * c + 3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sSt + boff + sw128_off(rl, c0 + c)),
                                     "r"(ds[4 * c]), "r"(ds[4 * c + 1]), "r"(ds[4 * c + 2]), "r"(ds[4 * c + 3]) : "memory");
                    }
                    fence_proxy_async_smem();
                    tcgen05_fence_before();
                    mbar_arrive(b_pds_ready);
                }
                // dK_kt, dV_kt complete: thread = (key row, 16-column group)
                mbar_wait(b_dkv_full, m & 1);
                tcgen05_fence_after();
                {
                    uint32_t kv[16], vv[16];
                    tmem_ld16(tlane + T_DK + cg * 16, kv);
                    tmem_ld16(tlane + T_DV + cg * 16, vv);
                    tmem_ld_wait();
                    tcgen05_fence_before();
                    mbar_arrive(b_dkv_free);
                    const int key = kt * 128 + rl;
                    if (key < N) {
                        __half* rowp = dqkv + ((int64_t)b * N + key) * lddqkv + h * 64 + cg * 16;
                        uint4 u0, u1;
                        u0.x = pack_half2(__uint_as_float(kv[0]) * scale, __uint_as_float(kv[1]) * scale);
                        u0.y = pack_half2(__uint_as_float(kv[2]) * scale, __uint_as_float(kv[3]) * scale);
                        u0.z = pack_half2(__uint_as_float(kv[4]) * scale, __uint_as_float(kv[5]) * scale);
                        u0.w = pack_half2(__uint_as_float(kv[6]) * scale, __uint_as_float(kv[7]) * scale);
                        u1.x = pack_half2(__uint_as_float(kv[8]) * scale, __uint_as_float(kv[9]) * scale);
                        u1.y = pack_half2(__uint_as_float(kv[10]) * scale, __uint_as_float(kv[11]) * scale);
                        u1.z = pack_half2(__uint_as_float(kv[12]) * scale, __uint_as_float(kv[13]) * scale);
                        u1.w = pack_half2(__uint_as_float(kv[14]) * scale, __uint_as_float(kv[15]) * scale);
                        reinterpret_cast<uint4*>(rowp + D)[0] = u0;
                        reinterpret_cast<uint4*>(rowp + D)[1] = u1;
                        u0.x = pack_half2(__uint_as_float(vv[0]), __uint_as_float(vv[1]));
                        u0.y = pack_half2(__uint_as_float(vv[2]), __uint_as_float(vv[3]));
                        u0.z = pack_half2(__uint_as_float(vv[4]), __uint_as_float(vv[5]));
                        u0.w = pack_half2(__uint_as_float(vv[6]), __uint_as_float(vv[7]));
                        u1.x = pack_half2(__uint_as_float(vv[8]), __uint_as_float(vv[9]));
                        u1.y = pack_half2(__uint_as_float(vv[10]), __uint_as_float(vv[11]));
                        u1.z = pack_half2(__uint_as_float(vv[12]), __uint_as_float(vv[13]));
                        u1.w = pack_half2(__uint_as_float(vv[14]), __uint_as_float(vv[15]));
                        reinterpret_cast<uint4*>(rowp + 2 * D)[0] = u0;
                        reinterpret_cast<uint4*>(rowp + 2 * D)[1] = u1;
                    }
                }
            }
            // dQ of both query tiles
            mbar_wait(b_dq_full, it & 1);
            tcgen05_fence_after();
            {
                uint32_t q0[16], q1[16];
                tmem_ld16(tlane + T_DQ + cg * 16, q0);
                if (nt > 1) tmem_ld16(tlane + T_DQ + 64 + cg * 16, q1);
                tmem_ld_wait();
                tcgen05_fence_before();
                mbar_arrive(b_dq_free);
                auto store_dq = [&](const uint32_t (&qq)[16], int q) {
                    uint4 u0, u1;
                    u0.x = pack_half2(__uint_as_float(qq[0]) * scale, __uint_as_float(qq[1]) * scale);
                    u0.y = pack_half2(__uint_as_float(qq[2]) * scale, __uint_as_float(qq[3]) * scale);
                    u0.z = pack_half2(__uint_as_float(qq[4]) * scale, __uint_as_float(qq[5]) * scale);
                    u0.w = pack_half2(__uint_as_float(qq[6]) * scale, __uint_as_float(qq[7]) * scale);
                    u1.x = pack_half2(__uint_as_float(qq[8]) * scale, __uint_as_float(qq[9]) * scale);
                    u1.y = pack_half2(__uint_as_float(qq[10]) * scale, __uint_as_float(qq[11]) * scale);
                    u1.z = pack_half2(__uint_as_float(qq[12]) * scale, __uint_as_float(qq[13]) * scale);
                    u1.w = pack_half2(__uint_as_float(qq[14]) * scale, __uint_as_float(qq[15]) * scale);
                    uint4* dst = reinterpret_cast<uint4*>(dqkv + ((int64_t)b * N + q) * lddqkv + h * 64 + cg * 16);
                    dst[0] = u0;
                    dst[1] = u1;
                };
                if (rl < N) store_dq(q0, rl);
                if (nt > 1 && 128 + rl < N) store_dq(q1, 128 + rl);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

int attention_bwd_tc(const __half* qkv, int64_t ld, const __half* out, int64_t ldo, const __half* dout, int64_t lddo, const float* lse,
                     __half* dqkv, int64_t lddqkv, int B, int N, int heads, float scale, cudaStream_t s) {
    GSL_REQUIRE(N >= 1 && N <= 256, "attention_bwd: tokens=%d outside [1, 256]", N);
    GSL_REQUIRE(ld % 8 == 0 && lddo % 8 == 0 && ldo % 8 == 0 && lddqkv % 8 == 0, "attention_bwd: pitches must be multiples of 8 halves");
    GSL_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dqkv)) & 15) == 0, "attention_bwd: 16-byte alignment required");
    CUtensorMap tq, td;
    int rc;
    if ((rc = make_tmap_qkv(&tq, qkv, ld, B, N, 3 * heads * 64, 256))) return rc;
    if ((rc = make_tmap_qkv(&td, dout, lddo, B, N, heads * 64, 256))) return rc;
    const int smem = 1024 + 4 * 32768 + 2 * 32768 + 128 + 4096 + 64;
    static bool attr = false;
    if (!attr) {
        GSL_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    const int sms = device_sm_count();
    const int nwork = B * heads;
    attention_bwd_tc_kernel<<<nwork < sms ? nwork : sms, ATB_THREADS, smem, s>>>(tq, td, out, ldo, lse, dqkv, lddqkv, B, N, heads, scale);
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gsl
