// Talking-heads mix -> softmax -> mix (cait.py:381-386) and its backward for H = 16 (CaiT-M36, BASELINE configs[3]) on mma.sync.
//
// With 16 heads the two H x H mixes ARE m16n8k16 shapes, so every column of a (b, q) row costs a handful of tensor instructions
// instead of 768 FP32 FMAs fed from shared memory (talking_generic.cu, which this file replaces for H = 16; the other odd head
// counts stay there).  The kernels are HBM bound by design: S is streamed twice, A written once.
//
// Orientation: the products are computed TRANSPOSED,  L^T[j, g] = sum_h S^T[j, h] Wl^T[h, g]:  keys j are the 16 MMA rows, heads the
// k / n dimensions.  Consequences:
//   * the accumulator fragment of one mix (rows j, head columns) is register-for-register the A fragment of the next mix (the
//     flash-attention P -> PV chaining): S -> L -> P -> A and dA -> dP -> dL -> dS need no shuffles and no shared memory;
//   * MMA row r of tile t belongs to thread group gid = r % 8 and maps to key  j = 64 c + 8 gid + 2 t + (r / 8): every thread owns 8
//     CONSECUTIVE keys of 4 heads {2 tig, 2 tig + 1, 2 tig + 8, 2 tig + 9}  ->  32-byte loads of S and 16-byte stores of A / dS, each
//     warp instruction covering whole 128-byte lines;
//   * the parameter gradients  dWw[o, g] = sum_j dA[j, o] P[j, g],  dWl[g, h] = sum_j dL[j, g] S[j, h]  contract over the MMA ROW index:
//     the same fragments are transposed in registers with movmatrix (8 x 8 b16 blocks) and fed to two more MMAs per tile; the key
//     permutation above is irrelevant for a sum over keys.
// Precision (as rowwise.cu for H <= 8): logits through one f16 pass (S rounded to f16, Wl log2e in f16, f32 accumulate); the forward
// starts the accumulator at bias - c2 + 4, so ex2.f16x2(pack(L')) = 2^4 P IS the A fragment of the output mix (2^-4 Ww); gradient mixes in
// bf16 with the weights split hi + lo (two MMAs).  Full 64-key chunks take a mask-free path; addresses advance by pointer increments.
// Statistics convention of rowwise.cu / talking_generic.cu:  stats[row][g] = c2 = max_j L2 + log2 sum_j 2^(L2 - max),  L2 = log2e (Wl S + bl).
#include "common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int T16_THREADS = 256;
constexpr int T16_WARPS = T16_THREADS / 32;
constexpr float T16_LOG2E = 1.4426950408889634f;
constexpr float T16_NEG = -1e30f;

__device__ __forceinline__ uint32_t pk_f16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pk_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 8 x 8 b16 transpose inside the warp: in = M[gid][2 tig, 2 tig + 1]  ->  out = M[2 tig, 2 tig + 1][gid]
__device__ __forceinline__ uint32_t movm(uint32_t x) {
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

constexpr float T16_SHIFT = 4.f;        // the forward carries P' = 2^4 P in f16 (output mix weights 2^-4 Ww)
constexpr float T16_UNSHIFT = 0.0625f;

__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) {
    uint32_t y;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t x) { return __half22float2(*reinterpret_cast<const __half2*>(&x)); }
__device__ __forceinline__ uint32_t h2_to_bf2(uint32_t x) {
    const float2 f = h2_to_f2(x);
    return pk_bf16(f.x, f.y);
}

// 8 consecutive logits of the 4 head slots {2 tig, 2 tig + 1, 2 tig + 8, 2 tig + 9} as f16 pairs along the keys: r[hs][t] = {key 2 t, key 2 t + 1}
// (fp32 logits are rounded here: the mix rounds them to f16 anyway).  MASK: keys >= Nk come back as 0 (the pitch padding may hold anything).
template <bool S16, bool MASK>
__device__ __forceinline__ void load_s(const void* __restrict__ S, long long off0, long long hS, int jb, int Nk, uint32_t (&r)[4][4]) {
    if (!MASK || jb < Nk) {
#pragma unroll
        for (int hs = 0; hs < 4; ++hs) {
            const long long off = off0 + (long long)((hs & 1) + 8 * (hs >> 1)) * hS;
            if (S16) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(S) + off));
                r[hs][0] = v.x; r[hs][1] = v.y; r[hs][2] = v.z; r[hs][3] = v.w;
            } else {
                const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(S) + off);
                const float4 x = __ldg(p), y = __ldg(p + 1);
                r[hs][0] = pk_f16(x.x, x.y); r[hs][1] = pk_f16(x.z, x.w); r[hs][2] = pk_f16(y.x, y.y); r[hs][3] = pk_f16(y.z, y.w);
            }
        }
        if (MASK) {
#pragma unroll
            for (int hs = 0; hs < 4; ++hs)
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (jb + 2 * t >= Nk) r[hs][t] = 0u;
                    else if (jb + 2 * t + 1 >= Nk) r[hs][t] &= 0xffffu;
                }
        }
    } else {
#pragma unroll
        for (int hs = 0; hs < 4; ++hs)
#pragma unroll
            for (int t = 0; t < 4; ++t) r[hs][t] = 0u;
    }
}
template <bool MASK>
__device__ __forceinline__ void load_da(const uint16_t* dp, long long hA, int jb, int Nk, uint32_t (&r)[4][4]) {
    if (!MASK || jb < Nk) {
#pragma unroll
        for (int hs = 0; hs < 4; ++hs) {
            const uint4 v = *reinterpret_cast<const uint4*>(dp + (long long)((hs & 1) + 8 * (hs >> 1)) * hA);
            r[hs][0] = v.x; r[hs][1] = v.y; r[hs][2] = v.z; r[hs][3] = v.w;
        }
        if (MASK) {
#pragma unroll
            for (int hs = 0; hs < 4; ++hs)
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (jb + 2 * t >= Nk) r[hs][t] = 0u;
                    else if (jb + 2 * t + 1 >= Nk) r[hs][t] &= 0xffffu;
                }
        }
    } else {
#pragma unroll
        for (int hs = 0; hs < 4; ++hs)
#pragma unroll
            for (int t = 0; t < 4; ++t) r[hs][t] = 0u;
    }
}
// tile t of a chunk as an m16k16 A fragment [rows = keys (slot 2 t | slot 2 t + 1), k = the 16 heads]
__device__ __forceinline__ void frag(const uint32_t (&r)[4][4], int t, uint32_t (&a)[4]) {
    a[0] = prmt(r[0][t], r[1][t], 0x5410u); a[1] = prmt(r[0][t], r[1][t], 0x7632u);
    a[2] = prmt(r[2][t], r[3][t], 0x5410u); a[3] = prmt(r[2][t], r[3][t], 0x7632u);
}

struct Mix1W {            // Wl log2e as the B operand [k = h, n = g] of the transposed logit mix, f16
    uint32_t b[2][2];
};
__device__ __forceinline__ void load_mix1(const float* __restrict__ Wl, int gid, int tig, Mix1W& w) {
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
        const float* r = Wl + (8 * nb + gid) * 16;
        w.b[nb][0] = pk_f16(r[2 * tig] * T16_LOG2E, r[2 * tig + 1] * T16_LOG2E);
        w.b[nb][1] = pk_f16(r[2 * tig + 8] * T16_LOG2E, r[2 * tig + 9] * T16_LOG2E);
    }
}
// L[r][gs] = init[gs] + sum_h S[key r][h] Wl'[g(gs)][h];  r = 0 -> key slot 2 t, r = 1 -> slot 2 t + 1;  gs = head slot
template <bool MASK>
__device__ __forceinline__ void mix1(const uint32_t (&a)[4], const Mix1W& w, const float (&init)[4], int jb, int t, int Nk, float (&L)[2][4]) {
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
        float d[4] = {init[2 * nb], init[2 * nb + 1], init[2 * nb], init[2 * nb + 1]};
        mma_f16(d, a[0], a[1], a[2], a[3], w.b[nb][0], w.b[nb][1]);
        L[0][2 * nb] = d[0]; L[0][2 * nb + 1] = d[1];
        L[1][2 * nb] = d[2]; L[1][2 * nb + 1] = d[3];
    }
    if (MASK) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
            if (jb + 2 * t + r >= Nk) {
#pragma unroll
                for (int gs = 0; gs < 4; ++gs) L[r][gs] = T16_NEG;
            }
    }
}

// ---- forward chunk steps (64 keys per warp: 4 tiles) ----
template <bool S16, bool MASK>
__device__ __forceinline__ void fwd_chunk_a(const void* __restrict__ S, long long off, long long hS, int jb, int Nk, const Mix1W& w1, const float (&b1)[4],
                                            float (&m)[4], float (&z)[4]) {
    uint32_t r[4][4];
    load_s<S16, MASK>(S, off, hS, jb, Nk, r);
    float L[4][2][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        uint32_t a[4];
        frag(r, t, a);
        mix1<MASK>(a, w1, b1, jb, t, Nk, L[t]);
    }
#pragma unroll
    for (int gs = 0; gs < 4; ++gs) {
        float mx = m[gs];
#pragma unroll
        for (int t = 0; t < 4; ++t) mx = fmaxf(mx, fmaxf(L[t][0][gs], L[t][1][gs]));
        float acc = z[gs] * ex2f(m[gs] - mx);
#pragma unroll
        for (int t = 0; t < 4; ++t) acc += ex2f(L[t][0][gs] - mx) + ex2f(L[t][1][gs] - mx);
        z[gs] = acc;
        m[gs] = mx;
    }
}
template <bool S16, bool MASK>
__device__ __forceinline__ void fwd_chunk_b(const void* __restrict__ S, long long off, long long hS, uint16_t* __restrict__ ap, long long hA, int jb, int Nk,
                                            int ldA, const Mix1W& w1, const uint32_t (&w2)[2][2], const float (&ci)[4], const float (&b2)[4]) {
    uint32_t r[4][4], pkt[4][4];
    load_s<S16, MASK>(S, off, hS, jb, Nk, r);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        uint32_t a[4];
        frag(r, t, a);
        float L[2][4];
        mix1<MASK>(a, w1, ci, jb, t, Nk, L);
        // P' = 2^4 P = 2^L packed = the A fragment of the output mix (rows = keys, k = mixed heads)
        const uint32_t p0 = ex2_h2(pk_f16(L[0][0], L[0][1])), p1 = ex2_h2(pk_f16(L[1][0], L[1][1]));
        const uint32_t p2 = ex2_h2(pk_f16(L[0][2], L[0][3])), p3 = ex2_h2(pk_f16(L[1][2], L[1][3]));
        const bool v0 = !MASK || jb + 2 * t < Nk, v1 = !MASK || jb + 2 * t + 1 < Nk;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            float e[4] = {b2[2 * nb], b2[2 * nb + 1], b2[2 * nb], b2[2 * nb + 1]};
            mma_f16(e, p0, p1, p2, p3, w2[nb][0], w2[nb][1]);
            pkt[2 * nb][t] = pk_bf16(v0 ? e[0] : 0.f, v1 ? e[2] : 0.f);                    // padding keys stay clean (the PV GEMM reads the pitch)
            pkt[2 * nb + 1][t] = pk_bf16(v0 ? e[1] : 0.f, v1 ? e[3] : 0.f);
        }
    }
    if (!MASK || jb < ldA) {
#pragma unroll
        for (int os = 0; os < 4; ++os)
            *reinterpret_cast<uint4*>(ap + (long long)((os & 1) + 8 * (os >> 1)) * hA) = make_uint4(pkt[os][0], pkt[os][1], pkt[os][2], pkt[os][3]);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------------
template <bool S16>
__global__ void __launch_bounds__(T16_THREADS, 2) th16_fwd_kernel(const void* __restrict__ S, uint16_t* __restrict__ A, const float* __restrict__ Wl,
                                                                  const float* __restrict__ bl, const float* __restrict__ Ww, const float* __restrict__ bw,
                                                                  float* __restrict__ stats, int rows_total, int Nq, int Nk, int ldS, int ldA) {
    __shared__ float redm[2][T16_WARPS][16], redz[2][T16_WARPS][16], sc2[2][16];      // double buffered over rows: no end-of-row barrier
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    Mix1W w1;
    load_mix1(Wl, gid, tig, w1);
    uint32_t w2[2][2];                      // 2^-4 Ww as the B operand [k = g, n = o] of the output mix, f16
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
        const float* r = Ww + (8 * nb + gid) * 16;
        w2[nb][0] = pk_f16(r[2 * tig] * T16_UNSHIFT, r[2 * tig + 1] * T16_UNSHIFT);
        w2[nb][1] = pk_f16(r[2 * tig + 8] * T16_UNSHIFT, r[2 * tig + 9] * T16_UNSHIFT);
    }
    const float b1[4] = {bl[2 * tig] * T16_LOG2E, bl[2 * tig + 1] * T16_LOG2E, bl[2 * tig + 8] * T16_LOG2E, bl[2 * tig + 9] * T16_LOG2E};
    const float b2[4] = {bw[2 * tig], bw[2 * tig + 1], bw[2 * tig + 8], bw[2 * tig + 9]};
    const long long hS = (long long)Nq * ldS, hA = (long long)Nq * ldA;
    const int nch = (ldA + 63) / 64, nfull = Nk / 64;                       // chunks [0, nfull) hold valid keys only: no masks
    int rb = blockIdx.x / Nq, rq = blockIdx.x % Nq;
    int rowit = 0;
    for (int row = blockIdx.x; row < rows_total; row += gridDim.x, ++rowit) {
        const int par = rowit & 1;
        const long long s_off = ((long long)rb * 16 * Nq + rq) * ldS + (long long)(2 * tig) * hS + gid * 8;
        uint16_t* Ab = A + ((long long)rb * 16 * Nq + rq) * ldA + (long long)(2 * tig) * hA + gid * 8;
        // ---- sweep A: online (max, sum) of the mixed logits, 4 head slots per thread
        float m[4], z[4];
#pragma unroll
        for (int gs = 0; gs < 4; ++gs) { m[gs] = T16_NEG; z[gs] = 0.f; }
        {
            int c = warp;
            for (; c < nfull; c += T16_WARPS) fwd_chunk_a<S16, false>(S, s_off + c * 64, hS, 0, Nk, w1, b1, m, z);
            for (; c < nch; c += T16_WARPS) fwd_chunk_a<S16, true>(S, s_off + c * 64, hS, c * 64 + gid * 8, Nk, w1, b1, m, z);
        }
#pragma unroll
        for (int gs = 0; gs < 4; ++gs) {
            if (m[gs] == T16_NEG) z[gs] = 0.f;           // no valid key seen: (NEG, junk) must vanish in the merge
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) {
                const float mo = __shfl_xor_sync(0xffffffffu, m[gs], off), zo = __shfl_xor_sync(0xffffffffu, z[gs], off);
                const float mx = fmaxf(m[gs], mo);
                z[gs] = z[gs] * ex2f(m[gs] - mx) + zo * ex2f(mo - mx);
                m[gs] = mx;
            }
            if (gid == 0) {
                const int g = 2 * tig + (gs & 1) + 8 * (gs >> 1);
                redm[par][warp][g] = m[gs];
                redz[par][warp][g] = z[gs];
            }
        }
        __syncthreads();
        if (tid < 16) {
            float M = T16_NEG, Z = 0.f;
#pragma unroll
            for (int w = 0; w < T16_WARPS; ++w) M = fmaxf(M, redm[par][w][tid]);
#pragma unroll
            for (int w = 0; w < T16_WARPS; ++w) Z += redz[par][w][tid] * ex2f(redm[par][w][tid] - M);
            const float c2 = M + __log2f(Z);
            sc2[par][tid] = c2;
            if (stats) stats[(long long)row * 16 + tid] = c2;
        }
        __syncthreads();
        // ---- sweep B: the accumulator starts at bias - c2 + 4, so 2^L is 2^4 P
        float ci[4];
#pragma unroll
        for (int gs = 0; gs < 4; ++gs) ci[gs] = b1[gs] - sc2[par][2 * tig + (gs & 1) + 8 * (gs >> 1)] + T16_SHIFT;
        {
            int c = warp;
            for (; c < nfull; c += T16_WARPS) fwd_chunk_b<S16, false>(S, s_off + c * 64, hS, Ab + c * 64, hA, 0, Nk, ldA, w1, w2, ci, b2);
            for (; c < nch; c += T16_WARPS) fwd_chunk_b<S16, true>(S, s_off + c * 64, hS, Ab + c * 64, hA, c * 64 + gid * 8, Nk, ldA, w1, w2, ci, b2);
        }
        rq += gridDim.x;
        while (rq >= Nq) { rq -= Nq; ++rb; }
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// backward: dS (may be written in place over dA), per-CTA partial dWl / dWw
// ------------------------------------------------------------------------------------------------------------------------
struct GradW {            // a weight matrix as the B operand of a gradient mix, bf16 hi + lo
    uint32_t hi[2][2], lo[2][2];
};
// B[k][n] = W[k * 16 + n] (k = the contracted head, n = the produced head), rows k = 2 tig.. of column n = 8 nb + gid
__device__ __forceinline__ void load_gradw(const float* __restrict__ W, int gid, int tig, GradW& w) {
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            const float x0 = W[(2 * tig + 8 * kk) * 16 + 8 * nb + gid], x1 = W[(2 * tig + 1 + 8 * kk) * 16 + 8 * nb + gid];
            const uint32_t h = pk_bf16(x0, x1);
            const float r0 = x0 - __uint_as_float(h << 16), r1 = x1 - __uint_as_float(h & 0xffff0000u);
            w.hi[nb][kk] = h;
            w.lo[nb][kk] = pk_bf16(r0, r1);
        }
}
// X[r][slot] (fp32, rows = the two key slots of the tile) -> bf16 A fragment
__device__ __forceinline__ void pack_a_bf16(const float (&X)[2][4], uint32_t (&a)[4]) {
    a[0] = pk_bf16(X[0][0], X[0][1]); a[1] = pk_bf16(X[1][0], X[1][1]);
    a[2] = pk_bf16(X[0][2], X[0][3]); a[3] = pk_bf16(X[1][2], X[1][3]);
}
// Y[r][slot] = sum_k X[r][k] W[k][slot]  (bf16 X fragment, hi + lo weights)
__device__ __forceinline__ void grad_mix(const uint32_t (&a)[4], const GradW& w, float (&Y)[2][4]) {
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        mma_bf16(d, a[0], a[1], a[2], a[3], w.hi[nb][0], w.hi[nb][1]);
        mma_bf16(d, a[0], a[1], a[2], a[3], w.lo[nb][0], w.lo[nb][1]);
        Y[0][2 * nb] = d[0]; Y[0][2 * nb + 1] = d[1];
        Y[1][2 * nb] = d[2]; Y[1][2 * nb + 1] = d[3];
    }
}
// acc[nb] (rows m = head of X, columns n = head of Y) += sum over the tile's 16 keys of X[key][m] Y[key][n];  x, y = A-style fragments
__device__ __forceinline__ void outer_acc(const uint32_t (&x)[4], const uint32_t (&y)[4], float (&acc)[2][4]) {
    const uint32_t xa0 = movm(x[0]), xa1 = movm(x[2]), xa2 = movm(x[1]), xa3 = movm(x[3]);
    const uint32_t y0 = movm(y[0]), y1 = movm(y[1]), y2 = movm(y[2]), y3 = movm(y[3]);
    mma_bf16(acc[0], xa0, xa1, xa2, xa3, y0, y1);
    mma_bf16(acc[1], xa0, xa1, xa2, xa3, y2, y3);
}

template <bool S16, bool MASK>
__device__ __forceinline__ void bwd_chunk_b(const void* __restrict__ S, long long off, long long hS, const uint16_t* dp, long long hA, int jb, int Nk,
                                            const Mix1W& w1, const float (&ci)[4], const GradW& gww, float (&rho)[4], float (&accWw)[2][4]) {
    uint32_t r[4][4], rd[4][4];
    load_s<S16, MASK>(S, off, hS, jb, Nk, r);
    load_da<MASK>(dp, hA, jb, Nk, rd);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        uint32_t a[4], da[4], pa[4];
        frag(r, t, a);
        frag(rd, t, da);
        float P[2][4], dP[2][4];
        mix1<MASK>(a, w1, ci, jb, t, Nk, P);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int gs = 0; gs < 4; ++gs) P[rr][gs] = ex2f(P[rr][gs]);
        grad_mix(da, gww, dP);
#pragma unroll
        for (int gs = 0; gs < 4; ++gs) rho[gs] += P[0][gs] * dP[0][gs] + P[1][gs] * dP[1][gs];
        pack_a_bf16(P, pa);
        outer_acc(da, pa, accWw);                                    // [o, g]
    }
}
template <bool S16, bool MASK>
__device__ __forceinline__ void bwd_chunk_c(const void* __restrict__ S, long long off, long long hS, const uint16_t* dp, uint16_t* gp, long long hA, int jb,
                                            int Nk, int ldA, const Mix1W& w1, const float (&ci)[4], const GradW& gww, const GradW& gwl, const float (&rho)[4],
                                            float (&accWl)[2][4]) {
    uint32_t r[4][4], rd[4][4], pkt[4][4];
    load_s<S16, MASK>(S, off, hS, jb, Nk, r);
    load_da<MASK>(dp, hA, jb, Nk, rd);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        uint32_t a[4], da[4], la[4], sa[4];
        frag(r, t, a);
        frag(rd, t, da);
        float L[2][4], dP[2][4], dL[2][4], dSv[2][4];
        mix1<MASK>(a, w1, ci, jb, t, Nk, L);
        grad_mix(da, gww, dP);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int gs = 0; gs < 4; ++gs) dL[rr][gs] = ex2f(L[rr][gs]) * (dP[rr][gs] - rho[gs]);
        pack_a_bf16(dL, la);
        grad_mix(la, gwl, dSv);
#pragma unroll
        for (int os = 0; os < 4; ++os) pkt[os][t] = pk_bf16(dSv[0][os], dSv[1][os]);
#pragma unroll
        for (int i = 0; i < 4; ++i) sa[i] = h2_to_bf2(a[i]);
        outer_acc(la, sa, accWl);                                    // [g, h]
    }
    if (!MASK || jb < ldA) {
#pragma unroll
        for (int os = 0; os < 4; ++os)
            *reinterpret_cast<uint4*>(gp + (long long)((os & 1) + 8 * (os >> 1)) * hA) = make_uint4(pkt[os][0], pkt[os][1], pkt[os][2], pkt[os][3]);
    }
}

template <bool S16, int NT>
__global__ void __launch_bounds__(NT, 1) th16_bwd_kernel(const void* __restrict__ S, const uint16_t* dA, uint16_t* dS, const float* __restrict__ Wl,
                                                         const float* __restrict__ bl, const float* __restrict__ Ww, const float* __restrict__ stats,
                                                         int rows_total, int Nq, int Nk, int ldS, int ldA, float* __restrict__ part) {
    constexpr int NW = NT / 32;
    constexpr int NP = 2 * 16 * 16 + 2 * 16;
    __shared__ float redr[2][NW][16];               // double buffered over rows
    __shared__ float spart[NP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    Mix1W w1;
    load_mix1(Wl, gid, tig, w1);
    const float b1[4] = {bl[2 * tig] * T16_LOG2E, bl[2 * tig + 1] * T16_LOG2E, bl[2 * tig + 8] * T16_LOG2E, bl[2 * tig + 9] * T16_LOG2E};
    GradW gww, gwl;
    load_gradw(Ww, gid, tig, gww);          // dP[j, g] = sum_o dA[j, o] Ww[o, g]
    load_gradw(Wl, gid, tig, gwl);          // dS[j, h] = sum_g dL[j, g] Wl[g, h]
    float accWw[2][4], accWl[2][4];
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int i = 0; i < 4; ++i) { accWw[nb][i] = 0.f; accWl[nb][i] = 0.f; }
    for (int i = tid; i < NP; i += NT) spart[i] = 0.f;
    const long long hS = (long long)Nq * ldS, hA = (long long)Nq * ldA;
    const int nch = (ldA + 63) / 64, nfull = Nk / 64;
    int rb = blockIdx.x / Nq, rq = blockIdx.x % Nq;
    int rowit = 0;
    for (int row = blockIdx.x; row < rows_total; row += gridDim.x, ++rowit) {
        const long long s_off = ((long long)rb * 16 * Nq + rq) * ldS + (long long)(2 * tig) * hS + gid * 8;
        const long long a_off = ((long long)rb * 16 * Nq + rq) * ldA + (long long)(2 * tig) * hA + gid * 8;
        const uint16_t* dAb = dA + a_off;
        uint16_t* dSb = dS + a_off;                  // may alias dAb: every thread reads its own (head, key group) cells first
        float ci[4], rho[4];
#pragma unroll
        for (int gs = 0; gs < 4; ++gs) { ci[gs] = b1[gs] - __ldg(stats + (long long)row * 16 + 2 * tig + (gs & 1) + 8 * (gs >> 1)); rho[gs] = 0.f; }
        const int par = rowit & 1;
        // ---- sweep B: rho[g] = sum_j P dP,  dWw += dA^T P
        {
            int c = warp;
            for (; c < nfull; c += NW) bwd_chunk_b<S16, false>(S, s_off + c * 64, hS, dAb + c * 64, hA, 0, Nk, w1, ci, gww, rho, accWw);
            for (; c < nch; c += NW) bwd_chunk_b<S16, true>(S, s_off + c * 64, hS, dAb + c * 64, hA, c * 64 + gid * 8, Nk, w1, ci, gww, rho, accWw);
        }
#pragma unroll
        for (int gs = 0; gs < 4; ++gs) {
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) rho[gs] += __shfl_xor_sync(0xffffffffu, rho[gs], off);
            if (gid == 0) redr[par][warp][2 * tig + (gs & 1) + 8 * (gs >> 1)] = rho[gs];
        }
        __syncthreads();
#pragma unroll
        for (int gs = 0; gs < 4; ++gs) {             // every thread sums the per-warp partials of its own four heads (two barriers per row, not four)
            float r = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) r += redr[par][w][2 * tig + (gs & 1) + 8 * (gs >> 1)];
            rho[gs] = r;
        }
        // ---- sweep C: dL = P (dP - rho),  dS = dL Wl (in place over dA),  dWl += dL^T S
        {
            int c = warp;
            for (; c < nfull; c += NW) bwd_chunk_c<S16, false>(S, s_off + c * 64, hS, dAb + c * 64, dSb + c * 64, hA, 0, Nk, ldA, w1, ci, gww, gwl, rho, accWl);
            for (; c < nch; c += NW)
                bwd_chunk_c<S16, true>(S, s_off + c * 64, hS, dAb + c * 64, dSb + c * 64, hA, c * 64 + gid * 8, Nk, ldA, w1, ci, gww, gwl, rho, accWl);
        }
        // no end-of-row barrier: redr is double buffered and the mid-row barrier keeps the warps within one row of each other
        rq += gridDim.x;
        while (rq >= Nq) { rq -= Nq; ++rb; }
    }
    // per-CTA partials in the layout talking_bwd_finalize_kernel reduces: [dWl (g, h) | dbl = 0 | dWw (o, g) | dbw = 0]
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int mrow = gid + 8 * (i >> 1), ncol = 8 * nb + 2 * tig + (i & 1);
            atomicAdd(&spart[mrow * 16 + ncol], accWl[nb][i]);
            atomicAdd(&spart[16 * 16 + 16 + mrow * 16 + ncol], accWw[nb][i]);
        }
    __syncthreads();
    float* pr = part + (long long)blockIdx.x * NP;
    for (int i = tid; i < NP; i += NT) pr[i] = spart[i];
}

int t16_grid(int B, int Nq) {
    const long long rows = (long long)B * Nq;
    return (int)(rows < 2LL * spe_num_sms() ? rows : 2LL * spe_num_sms());
}

}  // namespace

// entry points used by rowwise.cu's dispatchers (H = 16 only; s16: S holds f16 logits)
int spe_talking_h16_grid(int B, int Nq) { return t16_grid(B, Nq); }

int spe_talking_h16_fwd(const void* S, int s16, void* A, const float* Wl, const float* bl, const float* Ww, const float* bw, float* stats, int B, int Nq, int Nk,
                        long long ldS, long long ldA, cudaStream_t st) {
    SPE_CHECK(ldS % 8 == 0 && ldA % 8 == 0 && ldS >= Nk && ldA >= Nk && ldS < (1LL << 31) && ldA < (1LL << 31), "talking-heads H=16: bad leading dimensions");
    const int grid = t16_grid(B, Nq);
    if (s16)
        th16_fwd_kernel<true><<<grid, T16_THREADS, 0, st>>>(S, reinterpret_cast<uint16_t*>(A), Wl, bl, Ww, bw, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA);
    else
        th16_fwd_kernel<false><<<grid, T16_THREADS, 0, st>>>(S, reinterpret_cast<uint16_t*>(A), Wl, bl, Ww, bw, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA);
    SPE_LAUNCHED();
    return 0;
}

int spe_talking_h16_bwd(const void* S, int s16, const void* dA, void* dS, const float* Wl, const float* bl, const float* Ww, const float* stats, int B, int Nq,
                        int Nk, long long ldS, long long ldA, float* part, cudaStream_t st) {
    SPE_CHECK(stats, "talking-heads backward (H = 16) needs the forward statistics");
    SPE_CHECK(ldS % 8 == 0 && ldA % 8 == 0 && ldS >= Nk && ldA >= Nk && ldS < (1LL << 31) && ldA < (1LL << 31), "talking-heads H=16: bad leading dimensions");
    const int grid = t16_grid(B, Nq);
    // warps per CTA (one CTA per SM): 16 warps at <= 128 registers hide more of the load latency than 8 at 209 (measured at cfg4, fp16 logits: 0.92 / 0.79 / 0.73 ms for 256 / 384 / 512 threads) (SPE_TH16_BWD_THREADS: A/B)
    static const int nt = getenv("SPE_TH16_BWD_THREADS") ? atoi(getenv("SPE_TH16_BWD_THREADS")) : 512;
    const uint16_t* dA16 = reinterpret_cast<const uint16_t*>(dA);
    uint16_t* dS16 = reinterpret_cast<uint16_t*>(dS);
#define T16_BWD(S16_, NT_) th16_bwd_kernel<S16_, NT_><<<grid, NT_, 0, st>>>(S, dA16, dS16, Wl, bl, Ww, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA, part)
    if (s16) {
        if (nt == 256) T16_BWD(true, 256); else if (nt == 512) T16_BWD(true, 512); else T16_BWD(true, 384);
    } else {
        if (nt == 256) T16_BWD(false, 256); else if (nt == 512) T16_BWD(false, 512); else T16_BWD(false, 384);
    }
#undef T16_BWD
    SPE_LAUNCHED();
    return 0;
}
