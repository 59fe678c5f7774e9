// gslora-b200: attention backward on tcgen05 tensor cores (autograd of Attention.forward, vit_pytorch_face/vit_face.py:358-379,
// reached from loss_total.backward(), engine_cl.py:124).
//
// One persistent CTA per SM walks over (image, head) pairs.  The (query, key) plane of a pair is cut into 128-row tiles
// (N <= 208 tokens: tile 0 = rows 0..127, tile 1 = the npad - 128 <= 80 remaining rows) and processed as up to 2 x 2 blocks,
// key tile outer, query tile inner.  Per block (qt, kt), five matmuls and one elementwise stage:
//   S  = Q_qt K_kt^T ,  dP = dO_qt V_kt^T                        TMEM columns [0,128) and [128,256)      (N = valid keys of kt)
//   workers (8 warps, thread = query row x 64-key half):  P = exp(scale S - LSE),  dS = P (dP - delta)
//                -> fp16 tiles [128 q x 128 keys] in shared memory (two 64-key K-major blocks each)
//   dQ_qt += dS K_kt       (A = dS tile K-major,       B = K rows as MN-major)      TMEM [256,320) / [320,384)
//   dV_kt += P^T  dO_qt    (A = P  tile read MN-major,  B = dO rows as MN-major)     TMEM [448,512)
//   dK_kt += dS^T Q_qt     (A = dS tile read MN-major,  B = Q  rows as MN-major)     TMEM [384,448)
// All transposes are descriptor modes; nothing is transposed in memory, nothing is recomputed, no atomics.
//
// Pipeline (what keeps the tensor pipe busy):
//   * the workers hand S / dP back (sdp_free) as soon as they sit in registers, so the issuer launches S / dP of block n+1
//     BEFORE the dQ / dV / dK MMAs of block n: the tensor core computes the next scores while the workers exponentiate;
//   * operand tiles have their own buffers and barriers: K0 / V0 are released after the last kt = 0 block and the next pair's
//     K0 / V0 stream in during the kt = 1 blocks; Q0 / dO0 are double-buffered across pairs; the small tile-1 buffers are
//     refilled at the pair boundary, one block ahead of their first use;
//   * delta = rowsum(dO * O) and LSE of the NEXT pair are prepared by a dedicated warp straight from global memory.
// Warp roles: warps 0-7 workers (two warpgroups, raised to 176 registers with setmaxnreg: few fat threads with 16-wide independent
// chains hide the MUFU / TMEM latencies far better than many thin ones), warps 8-11 gradient writers (read dK / dV / dQ out of TMEM,
// stage them and issue the TMA stores - off the workers' critical path, where they cost 5.8 of 15.5 us per pair), warp 12 TMA producer,
// warp 13 MMA issuer (+ TMEM allocation), warps 14-15 delta.
#include "gsl_common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <type_traits>
#include "gsl_kernels.h"

namespace gsl {

int make_tmap_qkv(CUtensorMap* map, const void* ptr, int64_t ld, int B, int N, int cols, int npad);

static constexpr int AB_WORKER_WARPS = 8;
static constexpr int AB_WORKERS = AB_WORKER_WARPS * 32;
static constexpr int AB_DELTA_THREADS = 64;
static constexpr int AB_STORE_WARPS = 4;                  // one per TMEM lane quarter
static constexpr int AB_STORERS = AB_STORE_WARPS * 32;
static constexpr int AB_THREADS = AB_WORKERS + AB_STORERS + 128;       // + one warpgroup: producer, MMA issuer, two delta warps
static constexpr uint32_t AB_W_STORE = AB_WORKER_WARPS, AB_W_PROD = AB_W_STORE + AB_STORE_WARPS, AB_W_MMA = AB_W_PROD + 1, AB_W_DELTA = AB_W_PROD + 2;
static constexpr int AB_MAX_TOKENS = 208;

// shared memory map (bytes from the 1024-aligned base)
static constexpr uint32_t AB_T0 = 128 * 128;            // a 128-row tile of a [rows x 64] fp16 slab
static constexpr uint32_t AB_T1 = 80 * 128;             // tile 1: at most 80 rows
static constexpr uint32_t AB_Q0 = 0;                    // two parities
static constexpr uint32_t AB_DO0 = AB_Q0 + 2 * AB_T0;   // two parities
static constexpr uint32_t AB_K0 = AB_DO0 + 2 * AB_T0;
static constexpr uint32_t AB_V0 = AB_K0 + AB_T0;
static constexpr uint32_t AB_Q1 = AB_V0 + AB_T0;
static constexpr uint32_t AB_DO1 = AB_Q1 + AB_T1;
static constexpr uint32_t AB_K1 = AB_DO1 + AB_T1;
static constexpr uint32_t AB_V1 = AB_K1 + AB_T1;
static constexpr uint32_t AB_P = AB_V1 + AB_T1;         // P tile  [128 q x 128 keys] fp16 = 2 blocks of 16 KB
static constexpr uint32_t AB_DS = AB_P + 32768;         // dS tile
static constexpr uint32_t AB_STAGE = AB_DS + 32768;     // [128 rows x 64] fp16 staging tile of the dQ / dK / dV TMA stores
static constexpr uint32_t AB_BARS = AB_STAGE + AB_T0;
static constexpr uint32_t AB_STATS = AB_BARS + 256;     // [2 parities][lse * log2e, delta][256] floats
static constexpr uint32_t AB_SMEM = AB_STATS + 2 * 2 * 256 * 4;
static_assert(AB_STAGE % 1024 == 0, "staging tile must be 1024-byte aligned");
static_assert(AB_Q1 % 1024 == 0 && AB_T1 % 1024 == 0 && AB_P % 1024 == 0, "operand tiles must stay 1024-byte aligned (128B swizzle atoms)");
static_assert(AB_SMEM + 1024 <= 232448, "attention backward: shared memory budget");

enum : uint32_t {   // mbarrier indices
    B_FULL_Q0 = 0 /* +parity */, B_FULL_K0 = 2, B_FULL_1A = 3, B_FULL_1B = 4, B_FREE_Q0 = 5 /* +parity */, B_FREE_K0 = 7, B_FREE_1A = 8,
    B_FREE_1B = 9, B_SDP_FULL = 10, B_SDP_FREE = 11, B_PDS_READY = 12, B_TILES_FREE = 13, B_DKV_FULL = 14, B_DKV_FREE = 15,
    B_DQ_FULL = 16, B_DQ_FREE = 17, B_STATS_READY = 18 /* +parity */, B_STATS_FREE = 20 /* +parity */, B_COUNT = 22
};

// K-major / MN-major smem descriptors for 128-byte rows with 128B swizzle (8-row atoms of 1024 bytes)
__device__ __forceinline__ uint64_t ab_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;         // SBO: 8 rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint64_t ab_desc_mn2(uint32_t smem_addr, uint32_t lbo_bytes) {    // MN-major, two 64-element atoms along MN
    return ab_desc(smem_addr) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__device__ __host__ constexpr uint32_t ab_idesc(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void ab_tma(const void* desc, uint32_t bar, uint32_t dst, int col, int row, int b) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(col), "r"(row), "r"(b) : "memory");
}
__device__ __forceinline__ void ab_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

// P = exp2(S * sl2 - lq), dS = P * (dP - dq) for 16 columns, packed to fp16 pairs (two 16-byte chunks each)
__device__ __forceinline__ void ab_math16(const uint32_t (&sv)[16], const uint32_t (&pv)[16], uint4& pp0, uint4& pp1, uint4& ds0, uint4& ds1, int q, int kb,
                                          int N, float sl2, float lq, float dq) {
    uint32_t po[8], so[8];
    if (q < N && kb + 16 <= N) {            // interior: no masking
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(sv[j]), sl2, -lq));
            const float p1 = ex2_approx(fmaf(__uint_as_float(sv[j + 1]), sl2, -lq));
            po[j >> 1] = pack_half2(p0, p1);
            so[j >> 1] = pack_half2(p0 * (__uint_as_float(pv[j]) - dq), p1 * (__uint_as_float(pv[j + 1]) - dq));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const bool v0 = q < N && kb + j < N, v1 = q < N && kb + j + 1 < N;
            const float p0 = v0 ? ex2_approx(fmaf(__uint_as_float(sv[j]), sl2, -lq)) : 0.f;
            const float p1 = v1 ? ex2_approx(fmaf(__uint_as_float(sv[j + 1]), sl2, -lq)) : 0.f;
            po[j >> 1] = pack_half2(p0, p1);
            so[j >> 1] = pack_half2(v0 ? p0 * (__uint_as_float(pv[j]) - dq) : 0.f, v1 ? p1 * (__uint_as_float(pv[j + 1]) - dq) : 0.f);
        }
    }
    pp0 = make_uint4(po[0], po[1], po[2], po[3]); pp1 = make_uint4(po[4], po[5], po[6], po[7]);
    ds0 = make_uint4(so[0], so[1], so[2], so[3]); ds1 = make_uint4(so[4], so[5], so[6], so[7]);
}
__device__ __forceinline__ void ab_bar_storers() { asm volatile("bar.sync 2, %0;" ::"n"(AB_STORERS) : "memory"); }
// 32 fp32 accumulator values -> 32 fp16 (four 16-byte chunks), scaled
struct AbRow32 { uint4 c0, c1, c2, c3; };
__device__ __forceinline__ AbRow32 ab_pack32(const uint32_t (&v)[32], float scale) {
    auto pk = [&](int i) { return pack_half2(__uint_as_float(v[i]) * scale, __uint_as_float(v[i + 1]) * scale); };
    AbRow32 r;
    r.c0 = make_uint4(pk(0), pk(2), pk(4), pk(6));
    r.c1 = make_uint4(pk(8), pk(10), pk(12), pk(14));
    r.c2 = make_uint4(pk(16), pk(18), pk(20), pk(22));
    r.c3 = make_uint4(pk(24), pk(26), pk(28), pk(30));
    return r;
}
// Write-out of one [128 rows x 64] gradient tile (dQ / dK / dV) by the four writer warps: every thread drops the 64 values of row `rl`
// as fp16 into the swizzled staging tile, one elected thread issues a single TMA store (rows >= N are clipped by the tensor map).
// (Per-thread 32-byte global stores at a 6 KB row pitch cost ~0.7 ms per launch.)
__device__ __forceinline__ void ab_stage_store(const void* tmap, uint32_t stage, const AbRow32& lo, const AbRow32& hi, bool write, int rl, bool elected,
                                               int col, int row0, int b) {
    if (elected) tma_store_wait_read<0>();      // the previous store has finished reading the staging tile
    ab_bar_storers();
    if (write) {
        auto st = [&](int chunk, const uint4& v) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + sw128_off(rl, chunk)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        };
        st(0, lo.c0); st(1, lo.c1); st(2, lo.c2); st(3, lo.c3); st(4, hi.c0); st(5, hi.c1); st(6, hi.c2); st(7, hi.c3);
    }
    fence_proxy_async_smem();
    ab_bar_storers();
    if (elected) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(stage), "r"(col), "r"(row0), "r"(b) : "memory");
        tma_store_commit();
    }
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV0, const __grid_constant__ CUtensorMap tmQKV1,
                        const __grid_constant__ CUtensorMap tmDO0, const __grid_constant__ CUtensorMap tmDO1,
                        const __grid_constant__ CUtensorMap tmDQKV, const __half* __restrict__ out, int64_t ldo, const __half* __restrict__ dout, int64_t lddo,
                        const float* __restrict__ lse, const float* __restrict__ delta_parts, __half* __restrict__ dqkv, int64_t lddqkv, int B, int N, int heads,
                        float scale, int dbg) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sb = smem_u32(smem);
    const int npad = (N + 15) & ~15;
    const int nt = npad > 128 ? 2 : 1;              // 128-row tiles along queries and along keys
    const int n0r = nt == 2 ? 128 : npad;           // rows (queries / keys) of tile 0 that the MMAs contract over
    const int n1r = nt == 2 ? npad - 128 : 0;       // rows of tile 1
    const int D = heads * 64;
    auto bar = [&](uint32_t i) { return sb + AB_BARS + 8u * i; };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + AB_BARS + 8 * B_COUNT);
    float* s_stat = reinterpret_cast<float*>(smem + AB_STATS);

    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const int nwork = B * heads;
    const int nblk = nt * nt;

    if (warp == AB_W_MMA) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQKV0); tma_prefetch_desc(&tmQKV1); tma_prefetch_desc(&tmDO0); tma_prefetch_desc(&tmDO1); tma_prefetch_desc(&tmDQKV);
            for (uint32_t i = 0; i < B_COUNT; ++i) mbar_init(bar(i), 1);
            // worker-side barriers count WARPS: one elected lane arrives after __syncwarp() (512 per-thread arrivals on one mbarrier
            // serialise in the shared-memory atomics unit and cost microseconds per block)
            mbar_init(bar(B_SDP_FREE), AB_WORKER_WARPS); mbar_init(bar(B_PDS_READY), AB_WORKER_WARPS); mbar_init(bar(B_DKV_FREE), AB_STORE_WARPS);
            mbar_init(bar(B_DQ_FREE), AB_STORE_WARPS);
            for (uint32_t p = 0; p < 2; ++p) { mbar_init(bar(B_STATS_READY + p), AB_DELTA_THREADS / 32); mbar_init(bar(B_STATS_FREE + p), AB_WORKER_WARPS); }
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
    constexpr uint32_t T_S = 0, T_DP = 128, T_DQ = 256, T_DK = 384, T_DV = 448;

    // 512 threads x 128 registers at launch; producer / MMA / delta drop to 56, the writers to 96, the workers take the 13312 freed
    if (warp >= AB_W_PROD) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    } else if (warp >= AB_W_STORE) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    }
    if (warp == AB_W_PROD) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int it = 0;
            for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
                const int wb = w / heads, wh = w % heads;
                const uint32_t p = it & 1, k = it >> 1;
                if (k >= 1) mbar_wait(bar(B_FREE_Q0 + p), (k - 1) & 1);
                mbar_arrive_expect_tx(bar(B_FULL_Q0 + p), 2 * AB_T0);
                ab_tma(&tmQKV0, bar(B_FULL_Q0 + p), sb + AB_Q0 + p * AB_T0, wh * 64, 0, wb);
                ab_tma(&tmDO0, bar(B_FULL_Q0 + p), sb + AB_DO0 + p * AB_T0, wh * 64, 0, wb);
                if (it >= 1) mbar_wait(bar(B_FREE_K0), (it - 1) & 1);
                mbar_arrive_expect_tx(bar(B_FULL_K0), 2 * AB_T0);
                ab_tma(&tmQKV0, bar(B_FULL_K0), sb + AB_K0, D + wh * 64, 0, wb);
                ab_tma(&tmQKV0, bar(B_FULL_K0), sb + AB_V0, 2 * D + wh * 64, 0, wb);
                if (nt == 2) {
                    const uint32_t bytes1 = (uint32_t)n1r * 128u;
                    if (it >= 1) mbar_wait(bar(B_FREE_1A), (it - 1) & 1);
                    mbar_arrive_expect_tx(bar(B_FULL_1A), 2 * bytes1);
                    ab_tma(&tmQKV1, bar(B_FULL_1A), sb + AB_Q1, wh * 64, 128, wb);
                    ab_tma(&tmDO1, bar(B_FULL_1A), sb + AB_DO1, wh * 64, 128, wb);
                    if (it >= 1) mbar_wait(bar(B_FREE_1B), (it - 1) & 1);
                    mbar_arrive_expect_tx(bar(B_FULL_1B), 2 * bytes1);
                    ab_tma(&tmQKV1, bar(B_FULL_1B), sb + AB_K1, D + wh * 64, 128, wb);
                    ab_tma(&tmQKV1, bar(B_FULL_1B), sb + AB_V1, 2 * D + wh * 64, 128, wb);
                }
            }
        }
    } else if (warp == AB_W_MMA) {
        // ===================================================== MMA issuer
        // block descriptor of the previous block, whose dQ / dV / dK MMAs are issued after the next block's S / dP
        int pv_valid = 0, pv_qt = 0, pv_kt = 0, pv_it = 0;
        uint32_t pv_n = 0, pv_m = 0;
        auto q_addr = [&](int qt, int it_) { return qt == 0 ? sb + AB_Q0 + (it_ & 1) * AB_T0 : sb + AB_Q1; };
        auto do_addr = [&](int qt, int it_) { return qt == 0 ? sb + AB_DO0 + (it_ & 1) * AB_T0 : sb + AB_DO1; };
        auto k_addr = [&](int kt) { return kt == 0 ? sb + AB_K0 : sb + AB_K1; };
        auto v_addr = [&](int kt) { return kt == 0 ? sb + AB_V0 : sb + AB_V1; };
        auto issue_back = [&]() {
            const int qt = pv_qt, kt = pv_kt;
            mbar_wait(bar(B_PDS_READY), pv_n & 1);                                              // P / dS tiles written
            if (kt == 0 && qt == 0 && pv_it > 0) mbar_wait(bar(B_DQ_FREE), (pv_it - 1) & 1);    // previous pair's dQ has been read out
            if (qt == 0 && pv_m > 0) mbar_wait(bar(B_DKV_FREE), (pv_m - 1) & 1);                // previous key tile's dK / dV have been read out
            tcgen05_fence_after();
            if (lane == 0) {
                if (!(dbg & 4)) {
                constexpr uint32_t idesc_dq = ab_idesc(128, 64, 0, 1);      // A = dS K-major, B = K rows MN-major
                constexpr uint32_t idesc_dkv = ab_idesc(128, 64, 1, 1);     // A = tile^T (MN-major), B = slab rows MN-major
                const int ksteps = (kt == 0 ? n0r : n1r) / 16;              // contraction over this key tile's keys
                const int qsteps = (qt == 0 ? n0r : n1r) / 16;              // contraction over this query tile's rows
                const uint32_t sK = k_addr(kt), sQ = q_addr(qt, pv_it), sdO = do_addr(qt, pv_it);
                for (int j = 0; j < ksteps; ++j)                            // dQ_qt += dS K_kt
                    umma_f16<1>(tmem_base + T_DQ + qt * 64, ab_desc(sb + AB_DS + (j >> 2) * 16384 + (j & 3) * 32), ab_desc(sK + j * 2048), idesc_dq,
                                (kt | j) != 0);
                for (int j = 0; j < qsteps; ++j)                            // dV_kt += P^T dO_qt
                    umma_f16<1>(tmem_base + T_DV, ab_desc_mn2(sb + AB_P + j * 2048, 16384), ab_desc(sdO + j * 2048), idesc_dkv, (qt | j) != 0);
                for (int j = 0; j < qsteps; ++j)                            // dK_kt += dS^T Q_qt
                    umma_f16<1>(tmem_base + T_DK, ab_desc_mn2(sb + AB_DS + j * 2048, 16384), ab_desc(sQ + j * 2048), idesc_dkv, (qt | j) != 0);
                }
                umma_commit<1>(bar(B_TILES_FREE));
                if (qt == nt - 1) umma_commit<1>(bar(B_DKV_FULL));
                if (kt == 0 && qt == nt - 1) umma_commit<1>(bar(B_FREE_K0));        // every MMA that reads K0 / V0 has been issued
                if (kt == nt - 1 && qt == nt - 1) {                                 // last block of the pair
                    umma_commit<1>(bar(B_DQ_FULL));
                    umma_commit<1>(bar(B_FREE_Q0 + (pv_it & 1)));
                    if (nt == 2) { umma_commit<1>(bar(B_FREE_1A)); umma_commit<1>(bar(B_FREE_1B)); }
                }
            }
            __syncwarp();
        };
        int it = 0;
        uint32_t n = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
            const uint32_t p = it & 1;
            for (int kt = 0; kt < nt; ++kt) {
                for (int qt = 0; qt < nt; ++qt, ++n) {
                    if (kt == 0 && qt == 0) { mbar_wait(bar(B_FULL_Q0 + p), (it >> 1) & 1); mbar_wait(bar(B_FULL_K0), it & 1); }
                    if (kt == 0 && qt == 1) mbar_wait(bar(B_FULL_1A), it & 1);
                    if (kt == 1 && qt == 0) mbar_wait(bar(B_FULL_1B), it & 1);
                    if (n > 0) mbar_wait(bar(B_SDP_FREE), (n - 1) & 1);             // the workers hold S / dP of the previous block in registers
                    tcgen05_fence_after();
                    if (lane == 0) {
                        const uint32_t idesc_s = ab_idesc(128, (uint32_t)(kt == 0 ? n0r : n1r), 0, 0);
                        const uint64_t aq = ab_desc(q_addr(qt, it)), bk = ab_desc(k_addr(kt));
                        const uint64_t ad = ab_desc(do_addr(qt, it)), bv = ab_desc(v_addr(kt));
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16<1>(tmem_base + T_S, aq + 2 * k, bk + 2 * k, idesc_s, k != 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16<1>(tmem_base + T_DP, ad + 2 * k, bv + 2 * k, idesc_s, k != 0);
                        umma_commit<1>(bar(B_SDP_FULL));
                    }
                    __syncwarp();
                    if (pv_valid) issue_back();
                    pv_valid = 1; pv_qt = qt; pv_kt = kt; pv_it = it; pv_n = n; pv_m = (uint32_t)(it * nt + kt);
                }
            }
        }
        if (pv_valid) issue_back();
    } else if (warp < AB_WORKER_WARPS) {
        // ===================================================== workers: thread = (query row, 64-key half of the block)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
        const uint32_t quarter = warp & 3;                          // TMEM lane quarter this warp may access
        const uint32_t cg = warp >> 2;                              // 0..1: which 64-key block of the [128 x 128] tile
        const int rl = quarter * 32 + lane;                         // row inside a 128-row tile
        const float sl2 = scale * 1.4426950408889634f;
        const uint32_t tlane = tmem_base + ((quarter * 32u) << 16);
        int it = 0;
        uint32_t n = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
            const uint32_t p = it & 1;
            const float* st_lse = s_stat + p * 512;
            const float* st_del = st_lse + 256;
            mbar_wait(bar(B_STATS_READY + p), (it >> 1) & 1);
            for (int blk = 0; blk < nblk; ++blk, ++n) {
                const int kt = blk / nt, qt = blk % nt;
                const int nk = kt == 0 ? n0r : n1r;
                const int q = qt * 128 + rl;
                const bool rows_on = (qt * 128 + (int)quarter * 32) < N && !(dbg & 1);          // warp-uniform
                const int npieces = rows_on ? min(4, max(0, (nk - (int)cg * 64 + 15) / 16)) : 0;    // 16-column pieces this warp computes
                const int key0 = kt * 128 + cg * 64;
                // packed fp16 P / dS of this thread's 64 columns: piece k, 16-byte chunk c (scalars: arrays of uint4 end up in local memory)
                uint4 pp00, pp01, pp10, pp11, pp20, pp21, pp30, pp31, ds00, ds01, ds10, ds11, ds20, ds21, ds30, ds31;
                pp00 = pp01 = pp10 = pp11 = pp20 = pp21 = pp30 = pp31 = ds00 = ds01 = ds10 = ds11 = ds20 = ds21 = ds30 = ds31 = make_uint4(0u, 0u, 0u, 0u);
                mbar_wait(bar(B_SDP_FULL), n & 1);
                tcgen05_fence_after();
                const float lq = rows_on ? st_lse[q] : 0.f, dq = rows_on ? st_del[q] : 0.f;
                // software pipeline over the four 16-column pieces: the TMEM loads of piece k+1 are in flight while piece k is computed;
                // S / dP go back to the MMA issuer as soon as the last piece sits in registers
                uint32_t sa[16], pa[16], sb_[16], pb[16];
                if (npieces > 0) { ab_ld16(tlane + T_S + cg * 64, sa); ab_ld16(tlane + T_DP + cg * 64, pa); }
                tmem_ld_wait();
                if (npieces > 1) { ab_ld16(tlane + T_S + cg * 64 + 16, sb_); ab_ld16(tlane + T_DP + cg * 64 + 16, pb); }
                if (npieces > 0) ab_math16(sa, pa, pp00, pp01, ds00, ds01, q, key0, N, sl2, lq, dq);
                tmem_ld_wait();
                if (npieces > 2) { ab_ld16(tlane + T_S + cg * 64 + 32, sa); ab_ld16(tlane + T_DP + cg * 64 + 32, pa); }
                if (npieces > 1) ab_math16(sb_, pb, pp10, pp11, ds10, ds11, q, key0 + 16, N, sl2, lq, dq);
                tmem_ld_wait();
                if (npieces > 3) { ab_ld16(tlane + T_S + cg * 64 + 48, sb_); ab_ld16(tlane + T_DP + cg * 64 + 48, pb); }
                if (npieces > 2) ab_math16(sa, pa, pp20, pp21, ds20, ds21, q, key0 + 32, N, sl2, lq, dq);
                tmem_ld_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_SDP_FREE));        // S / dP may be overwritten by the next block's scores
                if (npieces > 3) ab_math16(sb_, pb, pp30, pp31, ds30, ds31, q, key0 + 48, N, sl2, lq, dq);

                if (n > 0) mbar_wait(bar(B_TILES_FREE), (n - 1) & 1);               // previous block's MMAs have read the tiles
                if (!(dbg & 2)) {
                    const uint32_t boff = cg * 16384;                               // this thread's 64-key block of the tile
                    auto sts2 = [&](int chunk, const uint4& pv4, const uint4& dv4) {
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + AB_P + boff + sw128_off(rl, chunk)),
                                     "r"(pv4.x), "r"(pv4.y), "r"(pv4.z), "r"(pv4.w) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + AB_DS + boff + sw128_off(rl, chunk)),
                                     "r"(dv4.x), "r"(dv4.y), "r"(dv4.z), "r"(dv4.w) : "memory");
                    };
                    if (npieces > 0) { sts2(0, pp00, ds00); sts2(1, pp01, ds01); }
                    if (npieces > 1) { sts2(2, pp10, ds10); sts2(3, pp11, ds11); }
                    if (npieces > 2) { sts2(4, pp20, ds20); sts2(5, pp21, ds21); }
                    if (npieces > 3) { sts2(6, pp30, ds30); sts2(7, pp31, ds31); }
                }
                fence_proxy_async_smem();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_PDS_READY));

            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_STATS_FREE + p));      // this pair's LSE / delta are no longer needed
        }
    } else if (warp < AB_W_PROD) {
        // ===================================================== gradient writers: thread = one row of a [128 x 64] tile
        const uint32_t quarter = warp & 3;
        const int rl = quarter * 32 + lane;
        const uint32_t tlane = tmem_base + ((quarter * 32u) << 16);
        const bool elected = warp == AB_W_STORE && lane == 0;
        const uint32_t stage = sb + AB_STAGE;
        auto read64 = [&](uint32_t tcol, float sc, AbRow32& lo, AbRow32& hi) {
            uint32_t t32[32];
            tmem_ld_32x32(tlane + tcol, t32);
            tmem_ld_wait();
            lo = ab_pack32(t32, sc);
            tmem_ld_32x32(tlane + tcol + 32, t32);
            tmem_ld_wait();
            hi = ab_pack32(t32, sc);
        };
        int it = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
            const int b = w / heads, h = w % heads;
            for (int kt = 0; kt < nt; ++kt) {
                const uint32_t m = (uint32_t)(it * nt + kt);
                const bool rows_live = (kt * 128 + (int)quarter * 32) < N;
                AbRow32 lo = {}, hi = {};
                mbar_wait(bar(B_DKV_FULL), m & 1);
                tcgen05_fence_after();
                if (rows_live) read64(T_DK, scale, lo, hi);
                if (!(dbg & 16)) ab_stage_store(&tmDQKV, stage, lo, hi, rows_live, rl, elected, D + h * 64, kt * 128, b);
                if (rows_live) read64(T_DV, 1.0f, lo, hi);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_DKV_FREE));       // dK / dV accumulators may be overwritten by the next key tile
                if (!(dbg & 16)) ab_stage_store(&tmDQKV, stage, lo, hi, rows_live, rl, elected, 2 * D + h * 64, kt * 128, b);
            }
            {   // dQ of both query tiles
                const bool live0 = (int)quarter * 32 < N, live1 = nt == 2 && (128 + (int)quarter * 32) < N;
                AbRow32 lo = {}, hi = {};
                mbar_wait(bar(B_DQ_FULL), it & 1);
                tcgen05_fence_after();
                if (live0) read64(T_DQ, scale, lo, hi);
                if (nt == 2) {
                    if (!(dbg & 16)) ab_stage_store(&tmDQKV, stage, lo, hi, live0, rl, elected, h * 64, 0, b);
                    if (live1) read64(T_DQ + 64, scale, lo, hi);
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_DQ_FREE));
                if (!(dbg & 16)) ab_stage_store(&tmDQKV, stage, lo, hi, nt == 2 ? live1 : live0, rl, elected, h * 64, nt == 2 ? 128 : 0, b);
            }
        }
        if (elected) tma_store_wait_all();
    } else if (warp >= AB_W_DELTA) {
        // ===================================================== delta warps: stats of the pair, one pair ahead of the workers
        const int dt = threadIdx.x - AB_W_DELTA * 32;
        int it = 0;
        for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++it) {
            const int b = w / heads, h = w % heads;
            const uint32_t p = it & 1, k = it >> 1;
            float* st_lse = s_stat + p * 512;
            float* st_del = st_lse + 256;
            if (k >= 1) mbar_wait(bar(B_STATS_FREE + p), (k - 1) & 1);
            if (delta_parts != nullptr) {
                // delta arrives precomputed (EPI_F16_ROWDOT epilogue of the GEMM that produced dO) as two partial sums per row: coalesced
                // fp32 reads.  (Computing it here from the O and dO rows made this warp pair the critical path of the whole kernel: 2 x 197
                // strided 128-byte rows per pair with a few loads in flight = 0.13 of 0.63 ms per launch.)
                const int64_t base = ((int64_t)b * heads + h) * N, part = (int64_t)B * heads * N;
                for (int r = dt; r < 256; r += AB_DELTA_THREADS) {
                    float d = 0.f, l = 0.f;
                    if (r < N && !(dbg & 8)) {
                        d = __ldg(delta_parts + base + r) + __ldg(delta_parts + part + base + r);
                        l = lse[base + r] * 1.4426950408889634f;
                    }
                    st_del[r] = d;
                    st_lse[r] = l;
                }
            } else
            for (int r = dt; r < 256; r += AB_DELTA_THREADS) {
                float d = 0.f, l = 0.f;
                if (r < N && !(dbg & 8)) {
                    // delta[q] = dO[q,:] . O[q,:], LSE pre-scaled by log2(e)
                    const uint4* orow = reinterpret_cast<const uint4*>(out + ((int64_t)b * N + r) * ldo + h * 64);
                    const uint4* grow = reinterpret_cast<const uint4*>(dout + ((int64_t)b * N + r) * lddo + h * 64);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint4 ov = __ldg(orow + c), gv = __ldg(grow + c);
                        const uint32_t oo[4] = {ov.x, ov.y, ov.z, ov.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 a = unpack_half2(oo[e]), g = unpack_half2(gg[e]);
                            d = fmaf(a.x, g.x, d); d = fmaf(a.y, g.y, d);
                        }
                    }
                    l = lse[((int64_t)b * heads + h) * N + r] * 1.4426950408889634f;
                }
                st_del[r] = d;
                st_lse[r] = l;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_STATS_READY + p));     // release (cumulative over the warp's stores) -> the workers' acquire
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == AB_W_MMA) {
        tcgen05_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

int attention_bwd(const __half* qkv, int64_t ld, const __half* out, int64_t ldo, const __half* dout, int64_t lddo, const float* lse,
                     __half* dqkv, int64_t lddqkv, int B, int N, int heads, float scale, cudaStream_t s, const float* delta_parts) {
    GSL_REQUIRE(N >= 1 && N <= AB_MAX_TOKENS, "attention_bwd: tokens=%d outside [1, %d]", N, AB_MAX_TOKENS);
    GSL_REQUIRE(ld % 8 == 0 && lddo % 8 == 0 && ldo % 8 == 0 && lddqkv % 8 == 0, "attention_bwd: pitches must be multiples of 8 halves");
    GSL_REQUIRE(out != nullptr || delta_parts != nullptr, "attention_bwd: needs the forward output or precomputed delta parts");
    GSL_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dqkv)) & 15) == 0,
                "attention_bwd: 16-byte alignment required");
    const int npad = (N + 15) & ~15;
    const int n0 = npad > 128 ? 128 : npad, n1 = npad > 128 ? npad - 128 : 16;    // (the tile-1 maps are unused when N <= 128)
    CUtensorMap tq0, tq1, td0, td1;
    int rc;
    if ((rc = make_tmap_qkv(&tq0, qkv, ld, B, N, 3 * heads * 64, 128))) return rc;
    if ((rc = make_tmap_qkv(&tq1, qkv, ld, B, N, 3 * heads * 64, n1))) return rc;
    if ((rc = make_tmap_qkv(&td0, dout, lddo, B, N, heads * 64, 128))) return rc;
    if ((rc = make_tmap_qkv(&td1, dout, lddo, B, N, heads * 64, n1))) return rc;
    (void)n0;
    CUtensorMap tout;
    if ((rc = make_tmap_qkv(&tout, dqkv, lddqkv, B, N, 3 * heads * 64, 128))) return rc;
    const int smem = AB_SMEM + 1024;
    static bool attr = false;
    if (!attr) {
        GSL_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    const int sms = device_sm_count();
    const int nwork = B * heads;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("GSL_ATTN_DBG"); dbg = e ? atoi(e) : 0; }       // dev switch: knock out stages to time the rest
    GSL_CHECK_CUDA(launch_pdl(attention_bwd_tc_kernel, dim3(nwork < sms ? nwork : sms), dim3(AB_THREADS), smem, s, tq0, tq1, td0, td1, tout, out, ldo, dout, lddo, lse, delta_parts, dqkv, lddqkv, B, N,
                                                                                heads, scale, dbg));
    GSL_COUNT_LAUNCH(1);
    GSL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gsl
