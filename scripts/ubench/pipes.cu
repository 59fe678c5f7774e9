// Dev microbenchmark (B200): issue rates of FFMA / FFMA2 / MUFU and their overlap, per SM per clock.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/pipes.bin scripts/ubench/pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITERS 4096
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, float a, float b, long long* cyc) {
    float x[8];
    unsigned long long y[8];
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 0.001f + i; y[i] = pk(x[i], x[i] + 0.5f); }
    const unsigned long long ab = pk(a, a), bb = pk(b, b);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {           // FFMA x8
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
        } else if (MODE == 1) {    // FFMA2 x8
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(ab), "l"(bb));
        } else if (MODE == 2) {    // MUFU.EX2 x8
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        } else if (MODE == 3) {    // MUFU.RCP x8
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        } else if (MODE == 4) {    // 8 FFMA2 + 2 MUFU
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(ab), "l"(bb));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[0]));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[1]));
        } else if (MODE == 5) {    // 8 FFMA + 2 MUFU
#pragma unroll
            for (int i = 0; i < 6; ++i) x[2 + i] = fmaf(x[2 + i], a, b);
            x[2] = fmaf(x[2], a, b); x[3] = fmaf(x[3], a, b);
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[0]));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[1]));
        } else if (MODE == 6) {    // 8 LOP3 (alu)
#pragma unroll
            for (int i = 0; i < 8; ++i) { uint32_t u = __float_as_uint(x[i]); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); x[i] = __uint_as_float(u); }
        } else if (MODE == 7) {    // 4 FFMA2 + 4 LOP3
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[i]) : "l"(ab), "l"(bb));
#pragma unroll
            for (int i = 4; i < 8; ++i) { uint32_t u = __float_as_uint(x[i]); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); x[i] = __uint_as_float(u); }
        } else if (MODE == 8) {    // F2FP pack x8
#pragma unroll
            for (int i = 0; i < 8; ++i) { uint32_t u; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(x[i]), "f"(x[(i + 1) & 7])); x[i] = __uint_as_float(u); }
        } else if (MODE == 9) {    // IMAD x8
#pragma unroll
            for (int i = 0; i < 8; ++i) { uint32_t u = __float_as_uint(x[i]); u = u * 0x9E3779B1u + 12345u; x[i] = __uint_as_float(u); }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i] + (float)(y[i] & 0xff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double ops_per_iter_per_thread) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, sms * 512 * 4); cudaMalloc(&cyc, sms * 8);
    k<MODE><<<sms, 512>>>(out, 1.0001f, 0.5f, cyc);
    cudaDeviceSynchronize();
    k<MODE><<<sms, 512>>>(out, 1.0001f, 0.5f, cyc);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
    // 512 threads = 16 warps/SM = 4 warps per scheduler
    double warp_instr_per_clk_per_smsp = ops_per_iter_per_thread * ITERS * 4 / c;
    printf("%-28s %8.0f cycles  -> %.3f warp-instr/clk/SMSP  (%.1f lane-ops/clk/SM)\n", name, c, warp_instr_per_clk_per_smsp, warp_instr_per_clk_per_smsp * 128);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FFMA x8", 8);
    run<1>("FFMA2 x8", 8);
    run<2>("MUFU.EX2 x8", 8);
    run<3>("MUFU.RCP x8", 8);
    run<4>("8 FFMA2 + 2 MUFU", 10);
    run<5>("8 FFMA + 2 MUFU", 10);
    run<6>("LOP3 x8", 8);
    run<7>("4 FFMA2 + 4 LOP3", 8);
    run<8>("F2FP x8", 8);
    run<9>("IMAD x8", 8);
    return 0;
}
