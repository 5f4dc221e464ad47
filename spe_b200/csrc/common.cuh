// Shared helpers for libspe_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/spe_b200.h"

extern thread_local char g_spe_err[512];
extern std::atomic<int64_t> g_spe_launches;

#define SPE_FAIL(...)                                          \
    do {                                                       \
        snprintf(g_spe_err, sizeof(g_spe_err), __VA_ARGS__);   \
        return -1;                                             \
    } while (0)

#define SPE_CHECK(cond, ...)          \
    do {                              \
        if (!(cond)) SPE_FAIL(__VA_ARGS__); \
    } while (0)

#define SPE_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) SPE_FAIL("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// ---- optional per-family device timing (bench.py roofline): CUDA events around launches on their stream ----
enum { SPE_FAM_GEMM = 0, SPE_FAM_TALKING_FWD = 1, SPE_FAM_TALKING_BWD = 2, SPE_FAM_SOFTMAX = 3, SPE_FAM_LAYERNORM = 4,
       SPE_FAM_MATCHER = 5, SPE_FAM_OTHER = 6, SPE_FAM_GEMM_ATTN = 7, SPE_FAM_ATTN_FUSED = 8, SPE_FAM_COUNT = 9 };
extern bool g_spe_prof_on;
int spe_prof_begin_(int fam, double work, cudaStream_t st, const char* tag = nullptr);
void spe_prof_end_(int idx, cudaStream_t st);
struct SpeProfScope {
    int idx; cudaStream_t st;
    SpeProfScope(int fam, double work, cudaStream_t s, const char* tag = nullptr) : idx(g_spe_prof_on ? spe_prof_begin_(fam, work, s, tag) : -1), st(s) {}
    ~SpeProfScope() { if (idx >= 0) spe_prof_end_(idx, st); }
};

// after a <<<>>> launch
#define SPE_LAUNCHED()                                                                     \
    do {                                                                                   \
        g_spe_launches.fetch_add(1, std::memory_order_relaxed);                            \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) SPE_FAIL("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

static inline int spe_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// talking-heads kernels for head counts outside {2, 4, 8} (talking_generic.cu)
// talking_h8.cu: H = 8 streamed from global memory on m16n8k8 fragments (s16: S holds f16 logits)
int spe_talking_h8_grid(int B, int Nq);
bool spe_talking_h8_fits(long long ldS, long long ldA);
int spe_talking_h8_fwd(const void* S, int s16, void* A, const float* Wl, const float* bl, const float* Ww, const float* bw, float* stats, int B, int Nq, int Nk,
                       long long ldS, long long ldA, cudaStream_t st);
int spe_talking_h8_bwd(const void* S, int s16, const void* dA, void* dS, const float* Wl, const float* bl, const float* Ww, const float* stats, int B, int Nq,
                       int Nk, long long ldS, long long ldA, float* part, cudaStream_t st);
// talking_h16.cu: H = 16 on mma.sync (s16: S holds f16 logits)
int spe_talking_h16_grid(int B, int Nq);
int spe_talking_h16_fwd(const void* S, int s16, void* A, const float* Wl, const float* bl, const float* Ww, const float* bw, float* stats, int B, int Nq, int Nk,
                        long long ldS, long long ldA, cudaStream_t st);
int spe_talking_h16_bwd(const void* S, int s16, const void* dA, void* dS, const float* Wl, const float* bl, const float* Ww, const float* stats, int B, int Nq,
                        int Nk, long long ldS, long long ldA, float* part, cudaStream_t st);
int spe_talking_generic_grid(int B, int Nq);
int spe_talking_generic_fwd(const float* S, void* A, const float* Wl, const float* bl, const float* Ww, const float* bw, float* stats, int B, int H, int Nq,
                            int Nk, long long ldS, long long ldA, cudaStream_t st);
int spe_talking_generic_bwd(const float* S, const void* dA, void* dS, const float* Wl, const float* bl, const float* Ww, const float* stats, int B, int H,
                            int Nq, int Nk, long long ldS, long long ldA, float* part, cudaStream_t st);

__device__ __forceinline__ float bf16_to_f(uint16_t v) { return __uint_as_float(((uint32_t)v) << 16); }
__device__ __forceinline__ uint16_t f_to_bf16(float f) {
    __nv_bfloat16 b = __float2bfloat16_rn(f);
    return *reinterpret_cast<uint16_t*>(&b);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2_rn(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint16_t f_to_f16(float f) {
    uint16_t r;
    asm("cvt.rn.f16.f32 %0, %1;" : "=h"(r) : "f"(f));
    return r;
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum for blockDim.x <= 1024; `red` is >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}

// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7): one exp + a degree-5 polynomial -- the libm erff costs ~4x more
// instructions and made the GELU epilogues ALU-bound.  e = exp(-x*x) is returned for reuse by the GELU derivative.
__device__ __forceinline__ float fast_erf(float x, float* e_out) {
    const float ax = fabsf(x);
    const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
    const float e = __expf(-ax * ax);
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float r = 1.f - p * t * e;
    if (e_out) *e_out = e;
    return copysignf(r, x);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + fast_erf(x * 0.70710678118654752f, nullptr)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    float e;                                             // e = exp(-x^2/2)
    const float cdf = 0.5f * (1.f + fast_erf(x * 0.70710678118654752f, &e));
    return fmaf(x * 0.3989422804014327f, e, cdf);
}
