// Talking-heads mix -> softmax -> mix (cait.py:381-386) for head counts the mma.sync kernels of rowwise.cu do not cover
// (they are built around 8-row fragments: H in {2, 4, 8}).  This file serves every other H <= 16, in particular H = 16 of
// CaiT-M36 (BASELINE configs[3]) and H = 6 / 12 of the XS / larger CaiT variants.
//
// One CTA (256 threads) per (b, q) row, one key per thread and pass; the H x H mixes are FP32 FMAs against weights held in
// shared memory (broadcast reads).  Statistics convention is the one of rowwise.cu:  L2 = log2(e) (Wl S + bl),
// stats[row][g] = c2 = max_j L2 + log2(sum_j 2^(L2 - max)),  P = 2^(L2 - c2).
//   fwd : sweep A online (max, sum) per mixed head -> c2;  sweep B: P -> second mix (+ bw) -> bf16 A.
//   bwd : sweep B: P, dP = Ww^T dA, rho = sum_j P dP, dWw += dA (x) P;   sweep C: dL = P (dP - rho), dS = Wl^T dL,
//         dWl += dL (x) S.  The outer products are accumulated by thread (g, h) = (tid / H, tid % H) from chunk slabs in
//         shared memory; per-CTA partials go to `part` and are reduced by talking_bwd_finalize_kernel (rowwise.cu).
//   dbl = 0 (softmax is shift invariant) and dbw is computed exactly by the caller (ops.py) -- both slots stay 0 here.
#include "common.cuh"

namespace {

constexpr float LOG2E_G = 1.4426950408889634f;
constexpr int TG_THREADS = 256;

template <int H>
__device__ __forceinline__ void tg_load_s(const float* __restrict__ Sb, long long hS, int j, float (&s)[H]) {
#pragma unroll
    for (int h = 0; h < H; ++h) s[h] = Sb[h * hS + j];
}

template <int H>
__device__ __forceinline__ void tg_logits(const float (&s)[H], const float* __restrict__ sWl, const float* __restrict__ sbl, float (&L)[H]) {
#pragma unroll
    for (int g = 0; g < H; ++g) {
        float a = sbl[g];
#pragma unroll
        for (int h = 0; h < H; ++h) a = fmaf(sWl[g * H + h], s[h], a);
        L[g] = a;
    }
}

template <int H>
__global__ void __launch_bounds__(TG_THREADS) talking_fwd_generic_kernel(const float* __restrict__ S, uint16_t* __restrict__ A, const float* __restrict__ Wl,
                                                                         const float* __restrict__ bl, const float* __restrict__ Ww,
                                                                         const float* __restrict__ bw, float* __restrict__ stats, int rows_total, int Nq,
                                                                         int Nk, long long ldS, long long ldA) {
    __shared__ float sWl[H * H], sWw[H * H], sbl[H], sbw[H];
    __shared__ float redm[TG_THREADS / 32][H], redz[TG_THREADS / 32][H], sc2[H];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < H * H; i += TG_THREADS) { sWl[i] = Wl[i] * LOG2E_G; sWw[i] = Ww[i]; }
    if (tid < H) { sbl[tid] = bl[tid] * LOG2E_G; sbw[tid] = bw[tid]; }
    __syncthreads();
    const long long hS = (long long)Nq * ldS, hA = (long long)Nq * ldA;
    for (int row = blockIdx.x; row < rows_total; row += gridDim.x) {
        const int b = row / Nq, q = row % Nq;
        const float* Sb = S + ((long long)b * H * Nq + q) * ldS;
        uint16_t* Ab = A + ((long long)b * H * Nq + q) * ldA;
        float m[H], z[H];
#pragma unroll
        for (int g = 0; g < H; ++g) { m[g] = -INFINITY; z[g] = 0.f; }
        for (int j = tid; j < Nk; j += TG_THREADS) {
            float s[H], L[H];
            tg_load_s<H>(Sb, hS, j, s);
            tg_logits<H>(s, sWl, sbl, L);
#pragma unroll
            for (int g = 0; g < H; ++g) {
                const float mn = fmaxf(m[g], L[g]);
                z[g] = z[g] * exp2f(m[g] - mn) + exp2f(L[g] - mn);
                m[g] = mn;
            }
        }
#pragma unroll
        for (int g = 0; g < H; ++g) {
            const float M = warp_max(m[g]);
            const float zz = warp_sum(m[g] == -INFINITY ? 0.f : z[g] * exp2f(m[g] - M));
            if (lane == 0) { redm[warp][g] = M; redz[warp][g] = zz; }
        }
        __syncthreads();
        if (tid < H) {
            float M = -INFINITY, Z = 0.f;
            for (int w = 0; w < TG_THREADS / 32; ++w) M = fmaxf(M, redm[w][tid]);
            for (int w = 0; w < TG_THREADS / 32; ++w) Z += redm[w][tid] == -INFINITY ? 0.f : redz[w][tid] * exp2f(redm[w][tid] - M);
            const float c2 = M + log2f(Z);
            sc2[tid] = c2;
            if (stats) stats[(long long)row * H + tid] = c2;
        }
        __syncthreads();
        for (int j = tid; j < (int)ldA; j += TG_THREADS) {
            if (j < Nk) {
                float s[H], L[H];
                tg_load_s<H>(Sb, hS, j, s);
                tg_logits<H>(s, sWl, sbl, L);
#pragma unroll
                for (int g = 0; g < H; ++g) L[g] = exp2f(L[g] - sc2[g]);                // P
#pragma unroll
                for (int o = 0; o < H; ++o) {
                    float a = sbw[o];
#pragma unroll
                    for (int g = 0; g < H; ++g) a = fmaf(sWw[o * H + g], L[g], a);
                    Ab[o * hA + j] = f_to_bf16(a);
                }
            } else {
#pragma unroll
                for (int o = 0; o < H; ++o) Ab[o * hA + j] = 0;                          // keep the padding columns clean
            }
        }
        __syncthreads();                                                                 // sc2 / red* are reused by the next row
    }
}

template <int H>
__global__ void __launch_bounds__(TG_THREADS) talking_bwd_generic_kernel(const float* __restrict__ S, const uint16_t* dA, uint16_t* dS, const float* __restrict__ Wl,
                                                                         const float* __restrict__ bl, const float* __restrict__ Ww,
                                                                         const float* __restrict__ stats, int rows_total, int Nq, int Nk, long long ldS,
                                                                         long long ldA, float* __restrict__ part) {
    constexpr int NP = 2 * H * H + 2 * H;
    constexpr int PITCH = TG_THREADS + 1;
    __shared__ float sWl2[H * H], sWl[H * H], sWw[H * H], sbl[H];
    __shared__ float slabX[H][PITCH], slabY[H][PITCH];            // chunk slabs of the two outer-product operands
    __shared__ float redr[TG_THREADS / 32][H], srho[H], sc2[H];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < H * H; i += TG_THREADS) { sWl[i] = Wl[i]; sWl2[i] = Wl[i] * LOG2E_G; sWw[i] = Ww[i]; }
    if (tid < H) sbl[tid] = bl[tid] * LOG2E_G;
    __syncthreads();
    const long long hS = (long long)Nq * ldS, hA = (long long)Nq * ldA;
    const int og = tid / H, oh = tid % H;                          // this thread's (row, col) of the H x H outer products
    const bool outer = tid < H * H;
    float accWw = 0.f, accWl = 0.f;
    const int nchunk = ((int)ldA + TG_THREADS - 1) / TG_THREADS;
    for (int row = blockIdx.x; row < rows_total; row += gridDim.x) {
        const int b = row / Nq, q = row % Nq;
        const float* Sb = S + ((long long)b * H * Nq + q) * ldS;
        const uint16_t* dAb = dA + ((long long)b * H * Nq + q) * ldA;
        uint16_t* dSb = dS + ((long long)b * H * Nq + q) * ldA;
        if (tid < H) sc2[tid] = stats[(long long)row * H + tid];
        __syncthreads();
        // ---- sweep B
        float rho[H];
#pragma unroll
        for (int g = 0; g < H; ++g) rho[g] = 0.f;
        for (int c = 0; c < nchunk; ++c) {
            const int j = c * TG_THREADS + tid;
            float p[H], d[H];
            if (j < Nk) {
                float s[H];
                tg_load_s<H>(Sb, hS, j, s);
                tg_logits<H>(s, sWl2, sbl, p);
#pragma unroll
                for (int g = 0; g < H; ++g) p[g] = exp2f(p[g] - sc2[g]);
#pragma unroll
                for (int o = 0; o < H; ++o) d[o] = bf16_to_f(dAb[o * hA + j]);
#pragma unroll
                for (int g = 0; g < H; ++g) {
                    float dp = 0.f;
#pragma unroll
                    for (int o = 0; o < H; ++o) dp = fmaf(sWw[o * H + g], d[o], dp);
                    rho[g] = fmaf(p[g], dp, rho[g]);
                }
            } else {
#pragma unroll
                for (int g = 0; g < H; ++g) { p[g] = 0.f; d[g] = 0.f; }
            }
#pragma unroll
            for (int g = 0; g < H; ++g) { slabX[g][tid] = d[g]; slabY[g][tid] = p[g]; }
            __syncthreads();
            if (outer) {
                float a = 0.f;
#pragma unroll 8
                for (int t = 0; t < TG_THREADS; ++t) a = fmaf(slabX[og][t], slabY[oh][t], a);     // dWw[o][g] += dA_o P_g
                accWw += a;
            }
            __syncthreads();
        }
#pragma unroll
        for (int g = 0; g < H; ++g) {
            const float r = warp_sum(rho[g]);
            if (lane == 0) redr[warp][g] = r;
        }
        __syncthreads();
        if (tid < H) {
            float r = 0.f;
            for (int w = 0; w < TG_THREADS / 32; ++w) r += redr[w][tid];
            srho[tid] = r;
        }
        __syncthreads();
        // ---- sweep C
        for (int c = 0; c < nchunk; ++c) {
            const int j = c * TG_THREADS + tid;
            float s[H], l[H];
            if (j < Nk) {
                float p[H], d[H];
                tg_load_s<H>(Sb, hS, j, s);
                tg_logits<H>(s, sWl2, sbl, p);
#pragma unroll
                for (int o = 0; o < H; ++o) d[o] = bf16_to_f(dAb[o * hA + j]);
#pragma unroll
                for (int g = 0; g < H; ++g) {
                    float dp = 0.f;
#pragma unroll
                    for (int o = 0; o < H; ++o) dp = fmaf(sWw[o * H + g], d[o], dp);
                    l[g] = exp2f(p[g] - sc2[g]) * (dp - srho[g]);
                }
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    float a = 0.f;
#pragma unroll
                    for (int g = 0; g < H; ++g) a = fmaf(sWl[g * H + h], l[g], a);
                    dSb[h * hA + j] = f_to_bf16(a);                  // in place over dA: this thread has read all of column j above
                }
            } else {
#pragma unroll
                for (int g = 0; g < H; ++g) { s[g] = 0.f; l[g] = 0.f; }
                if (j < (int)ldA) {
#pragma unroll
                    for (int h = 0; h < H; ++h) dSb[h * hA + j] = 0;
                }
            }
#pragma unroll
            for (int g = 0; g < H; ++g) { slabX[g][tid] = l[g]; slabY[g][tid] = s[g]; }
            __syncthreads();
            if (outer) {
                float a = 0.f;
#pragma unroll 8
                for (int t = 0; t < TG_THREADS; ++t) a = fmaf(slabX[og][t], slabY[oh][t], a);     // dWl[g][h] += dL_g S_h
                accWl += a;
            }
            __syncthreads();
        }
    }
    float* pr = part + (long long)blockIdx.x * NP;
    for (int i = tid; i < NP; i += TG_THREADS) pr[i] = 0.f;
    __syncthreads();
    if (outer) { pr[tid] = accWl; pr[H * H + H + tid] = accWw; }
}

int tg_grid(int B, int Nq) {
    const long long rows = (long long)B * Nq;
    return (int)(rows < 2LL * spe_num_sms() ? rows : 2LL * spe_num_sms());
}

template <int H>
int tg_fwd(const float* S, void* A, const float* Wl, const float* bl, const float* Ww, const float* bw, float* stats, int B, int Nq, int Nk, long long ldS,
           long long ldA, cudaStream_t st) {
    talking_fwd_generic_kernel<H><<<tg_grid(B, Nq), TG_THREADS, 0, st>>>(S, reinterpret_cast<uint16_t*>(A), Wl, bl, Ww, bw, stats, B * Nq, Nq, Nk, ldS, ldA);
    SPE_LAUNCHED();
    return 0;
}
template <int H>
int tg_bwd(const float* S, const void* dA, void* dS, const float* Wl, const float* bl, const float* Ww, const float* stats, int B, int Nq, int Nk, long long ldS,
           long long ldA, float* part, cudaStream_t st) {
    talking_bwd_generic_kernel<H><<<tg_grid(B, Nq), TG_THREADS, 0, st>>>(S, reinterpret_cast<const uint16_t*>(dA), reinterpret_cast<uint16_t*>(dS), Wl, bl, Ww, stats,
                                                                         B * Nq, Nq, Nk, ldS, ldA, part);
    SPE_LAUNCHED();
    return 0;
}

}  // namespace

// entry points used by rowwise.cu's dispatchers.  Return the grid size (= number of `part` rows) through *grid_out.
bool spe_talking_generic_supported(int H) { return H == 1 || H == 3 || H == 6 || H == 12 || H == 16; }
int spe_talking_generic_grid(int B, int Nq) { return tg_grid(B, Nq); }

int spe_talking_generic_fwd(const float* S, void* A, const float* Wl, const float* bl, const float* Ww, const float* bw, float* stats, int B, int H, int Nq,
                            int Nk, long long ldS, long long ldA, cudaStream_t st) {
    switch (H) {
        case 1: return tg_fwd<1>(S, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, st);
        case 3: return tg_fwd<3>(S, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, st);
        case 6: return tg_fwd<6>(S, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, st);
        case 12: return tg_fwd<12>(S, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, st);
        case 16: return tg_fwd<16>(S, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, st);
        default: SPE_FAIL("talking-heads kernels: unsupported head count %d (1/2/3/4/6/8/12/16)", H);
    }
}

int spe_talking_generic_bwd(const float* S, const void* dA, void* dS, const float* Wl, const float* bl, const float* Ww, const float* stats, int B, int H,
                            int Nq, int Nk, long long ldS, long long ldA, float* part, cudaStream_t st) {
    SPE_CHECK(stats, "talking-heads backward (generic head count) needs the forward statistics");
    switch (H) {
        case 1: return tg_bwd<1>(S, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, part, st);
        case 3: return tg_bwd<3>(S, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, part, st);
        case 6: return tg_bwd<6>(S, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, part, st);
        case 12: return tg_bwd<12>(S, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, part, st);
        case 16: return tg_bwd<16>(S, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, part, st);
        default: SPE_FAIL("talking-heads kernels: unsupported head count %d (1/2/3/4/6/8/12/16)", H);
    }
}

// ------------------------------------------------------------------------------------------------
// TSCAM_cait_two_branch.std_reweighting (cait.py:801-806, :827): class activation maps from the PER-HEAD class-attention
// probabilities of block 0:  cam[b,c,n] = sum_h P[b,h,q0+c,k0+n] * w[b,h,c],   w = minmax_h( std_n P[b,h,q0+c,k0+.] ).
// One CTA per (b, c); P bf16 [B,H,Lq,ldP] (what the attention forward keeps for its backward).  No gradient (the maps only
// feed the CPU pseudo-label code, engine.py:356-398).
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) cam_std_reweight_kernel(const uint16_t* __restrict__ P, int H, int Lq, long long ldP, int q0, int C, int k0, int N,
                                                               float* __restrict__ out) {
    __shared__ float red[32];
    __shared__ float sw[64];
    const int b = blockIdx.x / C, c = blockIdx.x % C;
    const uint16_t* Pb = P + (((long long)b * H) * Lq + (q0 + c)) * ldP + k0;
    const long long hP = (long long)Lq * ldP;
    for (int h = 0; h < H; ++h) {
        float s1 = 0.f;
        for (int n = threadIdx.x; n < N; n += blockDim.x) s1 += bf16_to_f(Pb[h * hP + n]);
        const float mean = block_sum(s1, red) / (float)N;
        float s2 = 0.f;
        for (int n = threadIdx.x; n < N; n += blockDim.x) { const float d = bf16_to_f(Pb[h * hP + n]) - mean; s2 += d * d; }
        const float var = block_sum(s2, red) / (float)(N - 1);                 // torch.std: unbiased
        if (threadIdx.x == 0) sw[h] = sqrtf(var);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float mn = INFINITY, mx = -INFINITY;
        for (int h = 0; h < H; ++h) mn = fminf(mn, sw[h]);
        for (int h = 0; h < H; ++h) mx = fmaxf(mx, sw[h] - mn);
        for (int h = 0; h < H; ++h) sw[h] = (sw[h] - mn) / mx;
    }
    __syncthreads();
    float* o = out + ((long long)b * C + c) * N;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float a = 0.f;
        for (int h = 0; h < H; ++h) a = fmaf(bf16_to_f(Pb[h * hP + n]), sw[h], a);
        o[n] = a;
    }
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int spe_cam_std_reweight(const void* P_bf16, int B, int H, int Lq, int64_t ldP, int q0, int C, int k0, int N,
                                                                          float* out, void* stream) {
    SPE_CHECK(P_bf16 && out && B > 0 && H > 0 && H <= 64 && C > 0 && N > 1 && q0 >= 0 && q0 + C <= Lq && k0 >= 0 && k0 + N <= ldP,
              "spe_cam_std_reweight: bad argument");
    cam_std_reweight_kernel<<<B * C, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint16_t*>(P_bf16), H, Lq, ldP, q0, C, k0, N, out);
    SPE_LAUNCHED();
    return 0;
}
