// Pipe-throughput microbenchmark for the fused talking-heads design (tools only; not part of libspe_b200.so).
// Measures per-SM per-clock rates of FFMA, FFMA2 (fma.rn.f32x2), HFMA2, EX2, mma.sync, and co-issue mixes on sm_100a.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, float seed, long long* clk) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed * (float)(threadIdx.x + i);
    float w0 = seed * 1.0001f, w1 = seed * 0.9999f;
    unsigned long long w01;
    asm("mov.b64 %0, {%1, %2};" : "=l"(w01) : "f"(w0), "f"(w1));
    long long t0 = clock64();
    if (MODE == 0) {            // FFMA 3-reg
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], w0, w1);
        }
    } else if (MODE == 1) {     // FFMA2
        unsigned long long p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(w01));
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(w01));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p[i]));
    } else if (MODE == 2) {     // HFMA2 f16
        __half2 h[16];
        __half2 hw = __floats2half2_rn(w0, w1);
#pragma unroll
        for (int i = 0; i < 16; ++i) h[i] = __floats2half2_rn(a[i], a[i] * 0.5f);
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) h[i] = __hfma2(h[i], hw, hw);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = __low2float(h[i]) + __high2float(h[i]);
    } else if (MODE == 3) {     // EX2
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
    } else if (MODE == 4) {     // 4 FFMA : 1 EX2 co-issue
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], w0, w1);
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
    } else if (MODE == 5) {     // mma.sync m16n8k16 f16 -> f32
        uint32_t A0 = __float_as_uint(a[0]), A1 = __float_as_uint(a[1]), A2 = __float_as_uint(a[2]), A3 = __float_as_uint(a[3]);
        uint32_t B0 = __float_as_uint(a[4]), B1 = __float_as_uint(a[5]);
        float c[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = a[i];
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[4 * i]), "+f"(c[4 * i + 1]), "+f"(c[4 * i + 2]), "+f"(c[4 * i + 3]) : "r"(A0), "r"(A1), "r"(A2), "r"(A3), "r"(B0), "r"(B1));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = c[i];
    } else if (MODE == 6) {     // 8 FFMA2 + 4 EX2 + 4 FFMA (scale) co-issue: the forward inner-loop mix
        unsigned long long p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
        float e[4] = {a[0], a[1], a[2], a[3]};
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(w01));
#pragma unroll
            for (int i = 0; i < 2; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[i]));
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(w01));
#pragma unroll
            for (int i = 2; i < 4; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[i]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p[i]));
        a[0] += e[0] + e[1] + e[2] + e[3];
    } else if (MODE == 7) {     // FFMA with different weight regs (3 distinct source regs, like a real mix)
        float w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = seed * (1.f + 0.001f * i);
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[(i + 1) & 15], w[i & 7], a[i]);
        }
    } else if (MODE == 8) {     // FFMA2 real-mix form: acc += x_pair * w_pair with distinct registers
        unsigned long long p[8], x[4], w[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
#pragma unroll
        for (int i = 0; i < 4; ++i) { asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(a[i] * 0.5f), "f"(a[i + 4] * 0.25f)); asm("mov.b64 %0, {%1, %2};" : "=l"(w[i]) : "f"(seed * i), "f"(seed * (i + 0.5f))); }
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(x[i & 3]), "l"(w[(i + r) & 3]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p[i]));
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[MODE] = t1 - t0;
}

int main() {
    float* out; long long* clk;
    cudaMalloc(&out, 148 * 512 * 4); cudaMallocManaged(&clk, 16 * 8);
    const char* names[] = {"FFMA(16 indep chains)", "FFMA2", "HFMA2.f16", "EX2", "16 FFMA + 4 EX2", "mma.sync m16n8k16 f16", "16 FFMA2 + 4 EX2", "FFMA 3 distinct src", "FFMA2 3 distinct src"};
    const double ops_per_iter[] = {16, 32, 32, 16, 16, 4 * 2048.0 / 32, 32, 16, 32};   // per-thread MACs (or EX2) per loop iteration
#define RUN(M) k<M><<<148, 512>>>(out, 1.0f, clk); cudaDeviceSynchronize(); k<M><<<148, 512>>>(out, 1.0f, clk); cudaDeviceSynchronize();
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    for (int m = 0; m < 9; ++m) {
        double per_clk_sm = ops_per_iter[m] * ITERS * 512.0 / (double)clk[m];
        printf("%-28s clocks %10lld  -> %.1f ops/clk/SM (thread-level MAC or EX2%s)\n", names[m], clk[m], per_clk_sm, m == 4 || m == 6 ? "; FMA ops only" : "");
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
