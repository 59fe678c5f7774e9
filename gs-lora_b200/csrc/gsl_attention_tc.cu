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
                mbar_init(b_s_full(i), 1); mbar_init(b_p_ready(i), 4); mbar_init(b_o_full(i), 1); mbar_init(b_s_free(i), 4);   // per-warp arrivals
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
                __syncwarp();
                if (lane == 0) mbar_arrive(b_p_ready(wg));
                // epilogue: O row = (P V) / sum
                mbar_wait(b_o_full(wg), ph);
                tcgen05_fence_after();
                uint32_t o0[32], o1[32];
                tmem_ld_32x32(taddr, o0);
                tmem_ld_32x32(taddr + 32, o1);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_s_free(wg));      // TMEM region may be overwritten by the next pair's S
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


}  // namespace gsl
