// Talking-heads mix -> softmax -> mix (cait.py:381-386) and its backward for H = 8 (CaiT-S24 / the benchmarked cfg2), fp16 logits:
// the row staging of rowwise.cu (bulk async copies of the (b, q) row into shared memory, double buffered) with the transposed, register-
// chained mixes of talking_h16.cu on m16n8k8 fragments (8 heads = the k / n extent of one MMA).
//
// Why: the rowwise.cu kernels spend ~31 (forward) / ~41 (backward) instructions per logit (ncu: issue bound at 2.8-3.0 TB/s); a first
// version of this file that streamed from global memory needed ~10 but was latency bound at N = 1600 (3 steps per warp and sweep).
// Here every thread owns 4 consecutive keys of 2 heads (one 8-byte shared-memory read per head), builds MMA fragments with byte
// permutes, and chains the mixes in registers (keys are the MMA rows, so an accumulator fragment IS the next A fragment):
//   forward  sweep A: L = S Wl'^T (+ bias)        -> online (max, sum); the exponentials of the statistics run as ex2.f16x2
//            sweep B: accumulator initialised with bias - c2 + 4, so P' = 2^4 P = ex2.f16x2(pack(L')) is ALREADY the A fragment of the
//                     output mix (Ww 2^-4): 2 packs + 2 MUFU per 16 x 8 tile, no per-element subtract / convert
//   backward sweep B: P (f32), dP = dA Ww, rho = sum P dP, dWw += dA^T P       sweep C: dL = P (dP - rho), dS = dL Wl, dWl += dL^T S
//            parameter-gradient outer products: movmatrix transposes of the same fragments, one m16n8k16 MMA per tile (rows 8..15 idle).
// Precision: as rowwise.cu -- logits through one f16 pass; probabilities in f16 (2^4 P: relative 2^-11, tails down to 4e-9); the
// exponent argument of sweep B is rounded to f16 (|L'| <= 4 for the dominant probabilities: <= 7e-4 relative); gradient mixes bf16 with
// hi + lo weights.  Statistics convention of rowwise.cu: stats[row][g] = c2 = max_j L2 + log2 sum_j 2^(L2 - max), L2 = log2e (Wl S + bl).
#include "common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr float T8_LOG2E = 1.4426950408889634f;
constexpr float T8_NEG = -1e30f;
constexpr float T8_SHIFT = 4.f;          // P' = 2^4 P in f16
constexpr float T8_UNSHIFT = 0.0625f;

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
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) {
    uint32_t y;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t x) { return __half22float2(*reinterpret_cast<const __half2*>(&x)); }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ uint32_t movm(uint32_t x) {
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
// D (16 x 8, f32) += A (16 x 8) B (8 x 8)
__device__ __forceinline__ void mma8_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void mma8_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void mma16_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- row staging (as rowwise.cu): the (b, q) row of all 8 heads comes into shared memory by bulk async copies, double buffered ----
__device__ __forceinline__ uint32_t t8_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t8_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void t8_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t8_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (clock64() - t0 > 8000000000LL) __trap();          // a protocol bug traps instead of hanging the GPU
    }
}
__device__ __forceinline__ void t8_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared-memory pitch of a 16-bit head row: = 8 (mod 32) elements, so the 8-byte reads of a half warp (4 key groups x 4 head pairs,
// head pairs 2 pitches apart) fall into 16 distinct 8-byte bank pairs
__host__ __device__ __forceinline__ int t8_pitch(int ld) { return ld + ((40 - ld % 32) % 32); }

#ifndef T8_UNROLL
#define T8_UNROLL 2
#endif
constexpr int T8_UNROLL_N = T8_UNROLL;      // full-step loops: independent chains of consecutive steps interleave
constexpr int T8_CH = 32;       // keys per warp step: 2 MMA tiles; a thread owns 4 consecutive keys of heads 2 tig, 2 tig + 1

// the 4 keys [jb, jb + 4) of heads 2 tig and 2 tig + 1 from a staged row (sp = this thread's position in the row of head 2 tig), as
// 16-bit pairs along the keys: r[hs][t] = {key 2 t, key 2 t + 1}.  MASK: keys >= Nk come back as 0 (the pitch padding may hold anything).
template <bool MASK>
__device__ __forceinline__ void lds_row(const uint16_t* __restrict__ sp, int pitch, int jb, int Nk, uint32_t (&r)[2][2]) {
    const uint2 v0 = *reinterpret_cast<const uint2*>(sp), v1 = *reinterpret_cast<const uint2*>(sp + pitch);
    r[0][0] = v0.x; r[0][1] = v0.y; r[1][0] = v1.x; r[1][1] = v1.y;
    if (MASK) {
#pragma unroll
        for (int hs = 0; hs < 2; ++hs)
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                if (jb + 2 * t >= Nk) r[hs][t] = 0u;
                else if (jb + 2 * t + 1 >= Nk) r[hs][t] &= 0xffffu;
            }
    }
}
// tile t of a step as an A fragment [rows = keys (slot 2 t | slot 2 t + 1), k = heads 2 tig, 2 tig + 1]
__device__ __forceinline__ void frag(const uint32_t (&r)[2][2], int t, uint32_t& a0, uint32_t& a1) {
    a0 = prmt(r[0][t], r[1][t], 0x5410u);
    a1 = prmt(r[0][t], r[1][t], 0x7632u);
}
// key validity of a partial group: d0, d1 belong to key jb + 2 t, d2, d3 to key jb + 2 t + 1
__device__ __forceinline__ void mask_tile(float (&d)[4], int jb, int t, int Nk, float v) {
    if (jb + 2 * t >= Nk) { d[0] = v; d[1] = v; }
    if (jb + 2 * t + 1 >= Nk) { d[2] = v; d[3] = v; }
}

// ---- forward steps (32 keys per warp: 2 tiles) ----
template <bool MASK>
__device__ __forceinline__ void fwd_step_a(const uint16_t* __restrict__ sp, int pS, int jb, int Nk, uint32_t w1, const float (&b1)[2], float (&m)[2], float (&z)[2]) {
    uint32_t r[2][2];
    lds_row<MASK>(sp, pS, jb, Nk, r);
    float L[2][4];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        uint32_t a0, a1;
        frag(r, t, a0, a1);
        L[t][0] = b1[0]; L[t][1] = b1[1]; L[t][2] = b1[0]; L[t][3] = b1[1];
        mma8_f16(L[t], a0, a1, w1);
        if (MASK) mask_tile(L[t], jb, t, Nk, T8_NEG);
    }
#pragma unroll
    for (int gs = 0; gs < 2; ++gs) {
        const float mx = fmaxf(fmaxf(m[gs], fmaxf(L[0][gs], L[0][2 + gs])), fmaxf(L[1][gs], L[1][2 + gs]));
        const float acc = (ex2f(L[0][gs] - mx) + ex2f(L[0][2 + gs] - mx)) + (ex2f(L[1][gs] - mx) + ex2f(L[1][2 + gs] - mx));
        z[gs] = z[gs] * ex2f(m[gs] - mx) + acc;
        m[gs] = mx;
    }
}
template <bool MASK>
__device__ __forceinline__ void fwd_step_b(const uint16_t* __restrict__ sp, int pS, uint16_t* __restrict__ ap, long long hA, int jb, int Nk, int ldA, uint32_t w1,
                                           uint32_t w2, const float (&ci)[2], const float (&b2)[2]) {
    uint32_t r[2][2], pkt[2][2];
    lds_row<MASK>(sp, pS, jb, Nk, r);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        uint32_t a0, a1;
        frag(r, t, a0, a1);
        float d[4] = {ci[0], ci[1], ci[0], ci[1]};
        mma8_f16(d, a0, a1, w1);
        if (MASK) mask_tile(d, jb, t, Nk, T8_NEG);
        const uint32_t p0 = ex2_h2(pk_f16(d[0], d[1])), p1 = ex2_h2(pk_f16(d[2], d[3]));     // = the A fragment of the output mix
        float e[4] = {b2[0], b2[1], b2[0], b2[1]};
        mma8_f16(e, p0, p1, w2);
        if (MASK) mask_tile(e, jb, t, Nk, 0.f);                                              // padding keys stay clean (the PV GEMM reads the pitch)
        pkt[0][t] = pk_bf16(e[0], e[2]);
        pkt[1][t] = pk_bf16(e[1], e[3]);
    }
    if (!MASK || jb < ldA) {
        *reinterpret_cast<uint2*>(ap) = make_uint2(pkt[0][0], pkt[0][1]);
        *reinterpret_cast<uint2*>(ap + hA) = make_uint2(pkt[1][0], pkt[1][1]);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// forward.  dynamic shared memory: [NBUF][8][pS] f16 row buffers
template <int NW, int NBUF>
__global__ void __launch_bounds__(NW * 32, NW <= 5 ? 4 : NW <= 10 ? 3 : 2) th8_fwd_kernel(const uint16_t* __restrict__ S, uint16_t* __restrict__ A, const float* __restrict__ Wl,
                                                            const float* __restrict__ bl, const float* __restrict__ Ww, const float* __restrict__ bw,
                                                            float* __restrict__ stats, int rows_total, int Nq, int Nk, int ldS, int ldA) {
    extern __shared__ __align__(128) uint8_t t8sm[];
    uint16_t* Sbuf = reinterpret_cast<uint16_t*>(t8sm);
    __shared__ __align__(8) uint64_t bars[NBUF];
    __shared__ float redm[2][NW][8], redz[2][NW][8], sc2[2][8];      // double buffered over rows
    const int pS = t8_pitch(ldS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    // B operands: logit mix [k = h, n = g] = Wl[g][h] log2e;  output mix [k = g, n = o] = Ww[o][g] 2^-4
    const uint32_t w1 = pk_f16(Wl[gid * 8 + 2 * tig] * T8_LOG2E, Wl[gid * 8 + 2 * tig + 1] * T8_LOG2E);
    const uint32_t w2 = pk_f16(Ww[gid * 8 + 2 * tig] * T8_UNSHIFT, Ww[gid * 8 + 2 * tig + 1] * T8_UNSHIFT);
    const float b1[2] = {bl[2 * tig] * T8_LOG2E, bl[2 * tig + 1] * T8_LOG2E};
    const float b2[2] = {bw[2 * tig], bw[2 * tig + 1]};
    const long long hS = (long long)Nq * ldS, hA = (long long)Nq * ldA;
    const uint32_t bar0 = t8_smem(bars);
    if (tid == 0) {
        for (int k = 0; k < NBUF; ++k) t8_mbar_init(bar0 + 8 * k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int row, int buf) {                                 // thread 0 only
        const int b = row / Nq, q = row % Nq;
        const uint32_t bar = bar0 + 8 * buf;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        t8_mbar_expect_tx(bar, (uint32_t)(8 * ldS * 2));
        for (int h = 0; h < 8; ++h)
            t8_bulk_g2s(t8_smem(Sbuf + ((size_t)buf * 8 + h) * pS), S + ((long long)b * 8 * Nq + q) * ldS + h * hS, (uint32_t)(ldS * 2), bar);
    };
    const int nst = (ldA + T8_CH - 1) / T8_CH;
    const int s0 = warp * nst / NW, s1 = (warp + 1) * nst / NW;         // contiguous steps per warp
    const int sf = min(s1, max(s0, Nk / T8_CH));                        // steps [s0, sf) hold valid keys only: no masks
    const int toff = (2 * tig) * pS + s0 * T8_CH + gid * 4;             // this thread's position inside a staged row
    // ring of NBUF row buffers: NBUF - 1 rows are in flight while one is processed
    if (tid == 0)
        for (int k = 0; k < NBUF - 1; ++k)
            if (blockIdx.x + (long long)k * gridDim.x < rows_total) issue(blockIdx.x + k * gridDim.x, k);
    int it = 0;
    int rb = blockIdx.x / Nq, rq = blockIdx.x % Nq;                     // (image, query) of the row, advanced without divisions
    for (int row = blockIdx.x; row < rows_total; row += gridDim.x, ++it) {
        const int buf = it % NBUF;
        if (tid == 0 && row + (long long)(NBUF - 1) * gridDim.x < rows_total)
            issue(row + (NBUF - 1) * gridDim.x, (it + NBUF - 1) % NBUF);       // that buffer was released by the previous iteration's last barrier
        t8_mbar_wait(bar0 + 8 * buf, (uint32_t)(it / NBUF) & 1u);
        const int par = it & 1;
        const uint16_t* sp0 = Sbuf + (size_t)buf * 8 * pS + toff;
        uint16_t* ap0 = A + ((long long)rb * 8 * Nq + rq) * ldA + (long long)(2 * tig) * hA + s0 * T8_CH + gid * 4;
        // ---- sweep A
        float m[2] = {T8_NEG, T8_NEG}, z[2] = {0.f, 0.f};
        {
            const uint16_t* sp = sp0;
            int st = s0;
#pragma unroll T8_UNROLL_N
            for (; st < sf; ++st, sp += T8_CH) fwd_step_a<false>(sp, pS, 0, Nk, w1, b1, m, z);
            for (; st < s1; ++st, sp += T8_CH) fwd_step_a<true>(sp, pS, st * T8_CH + gid * 4, Nk, w1, b1, m, z);
        }
#pragma unroll
        for (int gs = 0; gs < 2; ++gs) {
            if (m[gs] == T8_NEG) z[gs] = 0.f;             // no valid key seen: (NEG, junk) must vanish in the merge
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) {
                const float mo = __shfl_xor_sync(0xffffffffu, m[gs], off), zo = __shfl_xor_sync(0xffffffffu, z[gs], off);
                const float mx = fmaxf(m[gs], mo);
                z[gs] = z[gs] * ex2f(m[gs] - mx) + zo * ex2f(mo - mx);
                m[gs] = mx;
            }
            if (gid == 0) { redm[par][warp][2 * tig + gs] = m[gs]; redz[par][warp][2 * tig + gs] = z[gs]; }
        }
        __syncthreads();
        if (tid < 8) {                                   // (merging in every thread instead saves this barrier pair in the backward, but costs
            float M = T8_NEG, Z = 0.f;                   //  16 more MUFU per thread here and the forward is MUFU bound: 0.152 -> 0.162 ms)
#pragma unroll
            for (int w = 0; w < NW; ++w) M = fmaxf(M, redm[par][w][tid]);
#pragma unroll
            for (int w = 0; w < NW; ++w) Z += redz[par][w][tid] * ex2f(redm[par][w][tid] - M);
            const float c2 = M + __log2f(Z);
            sc2[par][tid] = c2;
            if (stats) stats[(long long)row * 8 + tid] = c2;
        }
        __syncthreads();
        const float ci[2] = {b1[0] - sc2[par][2 * tig] + T8_SHIFT, b1[1] - sc2[par][2 * tig + 1] + T8_SHIFT};
        // ---- sweep B
        {
            const uint16_t* sp = sp0;
            uint16_t* ap = ap0;
            int st = s0;
#pragma unroll T8_UNROLL_N
            for (; st < sf; ++st, sp += T8_CH, ap += T8_CH) fwd_step_b<false>(sp, pS, ap, hA, 0, Nk, ldA, w1, w2, ci, b2);
            for (; st < s1; ++st, sp += T8_CH, ap += T8_CH) fwd_step_b<true>(sp, pS, ap, hA, st * T8_CH + gid * 4, Nk, ldA, w1, w2, ci, b2);
        }
        __syncthreads();                                 // the row buffer is free again (red* are double buffered)
        rq += gridDim.x;
        while (rq >= Nq) { rq -= Nq; ++rb; }
    }
}

// ------------------------------------------------------------------------------------------------------------------------
struct GradW8 { uint32_t hi, lo; };
// B[k][n] = W[k * 8 + n]: rows k = 2 tig, 2 tig + 1 of column n = gid, bf16 hi + lo
__device__ __forceinline__ GradW8 load_gradw8(const float* __restrict__ W, int gid, int tig) {
    const float x0 = W[(2 * tig) * 8 + gid], x1 = W[(2 * tig + 1) * 8 + gid];
    GradW8 w;
    w.hi = pk_bf16(x0, x1);
    w.lo = pk_bf16(x0 - __uint_as_float(w.hi << 16), x1 - __uint_as_float(w.hi & 0xffff0000u));
    return w;
}
// f16x2 -> bf16x2 (same two values)
__device__ __forceinline__ uint32_t h2_to_bf2(uint32_t x) {
    const float2 f = h2_to_f2(x);
    return pk_bf16(f.x, f.y);
}
// acc (rows 0..7 = heads of X, columns = heads of Y) += sum over the tile's 16 keys X[key][.] Y[key][.];  (x0, x1), (y0, y1): m16k8 A-style
__device__ __forceinline__ void outer8(uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1, float (&acc)[4]) {
    mma16_bf16(acc, movm(x0), 0u, movm(x1), 0u, movm(y0), movm(y1));
}

// ---- backward steps ----
template <bool MASK>
__device__ __forceinline__ void bwd_step_b(const uint16_t* __restrict__ sp, int pS, const uint16_t* __restrict__ dp, int pA, int jb, int Nk, uint32_t w1,
                                           const float (&ci)[2], const GradW8& gww, float (&rho)[2], float (&accWw)[4]) {
    uint32_t r[2][2], rd[2][2];
    lds_row<MASK>(sp, pS, jb, Nk, r);
    lds_row<MASK>(dp, pA, jb, Nk, rd);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        uint32_t a0, a1, d0, d1;
        frag(r, t, a0, a1);
        frag(rd, t, d0, d1);
        float P[4] = {ci[0], ci[1], ci[0], ci[1]};
        mma8_f16(P, a0, a1, w1);
        if (MASK) mask_tile(P, jb, t, Nk, T8_NEG);
#pragma unroll
        for (int i = 0; i < 4; ++i) P[i] = ex2f(P[i]);
        float dP[4] = {0.f, 0.f, 0.f, 0.f};
        mma8_bf16(dP, d0, d1, gww.hi);
        mma8_bf16(dP, d0, d1, gww.lo);
        rho[0] += P[0] * dP[0] + P[2] * dP[2];
        rho[1] += P[1] * dP[1] + P[3] * dP[3];
        outer8(d0, d1, pk_bf16(P[0], P[1]), pk_bf16(P[2], P[3]), accWw);            // [o, g]
    }
}
template <bool MASK>
__device__ __forceinline__ void bwd_step_c(const uint16_t* __restrict__ sp, int pS, const uint16_t* __restrict__ dp, int pA, uint16_t* gp, long long hA, int jb,
                                           int Nk, int ldA, uint32_t w1, const float (&ci)[2], const GradW8& gww, const GradW8& gwl, const float (&rho)[2],
                                           float (&accWl)[4]) {
    uint32_t r[2][2], rd[2][2], pkt[2][2];
    lds_row<MASK>(sp, pS, jb, Nk, r);
    lds_row<MASK>(dp, pA, jb, Nk, rd);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        uint32_t a0, a1, d0, d1;
        frag(r, t, a0, a1);
        frag(rd, t, d0, d1);
        float P[4] = {ci[0], ci[1], ci[0], ci[1]};
        mma8_f16(P, a0, a1, w1);
        if (MASK) mask_tile(P, jb, t, Nk, T8_NEG);
        float dP[4] = {0.f, 0.f, 0.f, 0.f};
        mma8_bf16(dP, d0, d1, gww.hi);
        mma8_bf16(dP, d0, d1, gww.lo);
        float dL[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) dL[i] = ex2f(P[i]) * (dP[i] - rho[i & 1]);
        const uint32_t l0 = pk_bf16(dL[0], dL[1]), l1 = pk_bf16(dL[2], dL[3]);
        float dSv[4] = {0.f, 0.f, 0.f, 0.f};
        mma8_bf16(dSv, l0, l1, gwl.hi);
        mma8_bf16(dSv, l0, l1, gwl.lo);
        pkt[0][t] = pk_bf16(dSv[0], dSv[2]);
        pkt[1][t] = pk_bf16(dSv[1], dSv[3]);
        outer8(l0, l1, h2_to_bf2(a0), h2_to_bf2(a1), accWl);                         // [g, h]
    }
    if (!MASK || jb < ldA) {
        *reinterpret_cast<uint2*>(gp) = make_uint2(pkt[0][0], pkt[0][1]);
        *reinterpret_cast<uint2*>(gp + hA) = make_uint2(pkt[1][0], pkt[1][1]);
    }
}

// backward.  dynamic shared memory: [2][8][pS] f16 logit rows, then [2][8][pA] bf16 dA rows
template <int NW>
__global__ void __launch_bounds__(NW * 32, 2) th8_bwd_kernel(const uint16_t* __restrict__ S, const uint16_t* dA, uint16_t* dS, const float* __restrict__ Wl,
                                                            const float* __restrict__ bl, const float* __restrict__ Ww, const float* __restrict__ stats,
                                                            int rows_total, int Nq, int Nk, int ldS, int ldA, float* __restrict__ part) {
    constexpr int NT = NW * 32;
    constexpr int NP = 2 * 8 * 8 + 2 * 8;
    extern __shared__ __align__(128) uint8_t t8sm[];
    const int pS = t8_pitch(ldS), pA = t8_pitch(ldA);
    uint16_t* Sbuf = reinterpret_cast<uint16_t*>(t8sm);
    uint16_t* Dbuf = Sbuf + (size_t)2 * 8 * pS;
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ float redr[2][NW][8], spart[NP];           // redr double buffered over rows
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
    const uint32_t w1 = pk_f16(Wl[gid * 8 + 2 * tig] * T8_LOG2E, Wl[gid * 8 + 2 * tig + 1] * T8_LOG2E);
    const float b1[2] = {bl[2 * tig] * T8_LOG2E, bl[2 * tig + 1] * T8_LOG2E};
    const GradW8 gww = load_gradw8(Ww, gid, tig);       // dP[j, g] = sum_o dA[j, o] Ww[o, g]
    const GradW8 gwl = load_gradw8(Wl, gid, tig);       // dS[j, h] = sum_g dL[j, g] Wl[g, h]
    float accWw[4] = {0.f, 0.f, 0.f, 0.f}, accWl[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < NP; i += NT) spart[i] = 0.f;
    const long long hS = (long long)Nq * ldS, hA = (long long)Nq * ldA;
    const uint32_t bar0 = t8_smem(bars);
    if (tid == 0) {
        t8_mbar_init(bar0, 1); t8_mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int row, int buf) {                                 // thread 0 only
        const int b = row / Nq, q = row % Nq;
        const uint32_t bar = bar0 + 8 * buf;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        t8_mbar_expect_tx(bar, (uint32_t)(8 * ldS * 2 + 8 * ldA * 2));
        for (int h = 0; h < 8; ++h) {
            t8_bulk_g2s(t8_smem(Sbuf + ((size_t)buf * 8 + h) * pS), S + ((long long)b * 8 * Nq + q) * ldS + h * hS, (uint32_t)(ldS * 2), bar);
            t8_bulk_g2s(t8_smem(Dbuf + ((size_t)buf * 8 + h) * pA), dA + ((long long)b * 8 * Nq + q) * ldA + h * hA, (uint32_t)(ldA * 2), bar);
        }
    };
    const int nst = (ldA + T8_CH - 1) / T8_CH;
    const int s0 = warp * nst / NW, s1 = (warp + 1) * nst / NW;
    const int sf = min(s1, max(s0, Nk / T8_CH));
    const int toffS = (2 * tig) * pS + s0 * T8_CH + gid * 4, toffA = (2 * tig) * pA + s0 * T8_CH + gid * 4;
    if (tid == 0 && blockIdx.x < rows_total) issue(blockIdx.x, 0);
    int it = 0;
    int rb = blockIdx.x / Nq, rq = blockIdx.x % Nq;
    for (int row = blockIdx.x; row < rows_total; row += gridDim.x, ++it) {
        const int buf = it & 1;
        if (tid == 0 && row + (int)gridDim.x < rows_total) issue(row + gridDim.x, buf ^ 1);
        const float c2a = __ldg(stats + (long long)row * 8 + 2 * tig), c2b = __ldg(stats + (long long)row * 8 + 2 * tig + 1);
        t8_mbar_wait(bar0 + 8 * buf, (uint32_t)(it >> 1) & 1u);
        const int par = it & 1;
        const uint16_t* sp0 = Sbuf + (size_t)buf * 8 * pS + toffS;
        const uint16_t* dp0 = Dbuf + (size_t)buf * 8 * pA + toffA;
        // dS may alias this row of dA: it is staged in shared memory by now
        uint16_t* gp0 = dS + ((long long)rb * 8 * Nq + rq) * ldA + (long long)(2 * tig) * hA + s0 * T8_CH + gid * 4;
        const float ci[2] = {b1[0] - c2a, b1[1] - c2b};
        float rho[2] = {0.f, 0.f};
        // ---- sweep B
        {
            const uint16_t *sp = sp0, *dp = dp0;
            int st = s0;
#pragma unroll T8_UNROLL_N
            for (; st < sf; ++st, sp += T8_CH, dp += T8_CH) bwd_step_b<false>(sp, pS, dp, pA, 0, Nk, w1, ci, gww, rho, accWw);
            for (; st < s1; ++st, sp += T8_CH, dp += T8_CH) bwd_step_b<true>(sp, pS, dp, pA, st * T8_CH + gid * 4, Nk, w1, ci, gww, rho, accWw);
        }
#pragma unroll
        for (int gs = 0; gs < 2; ++gs) {
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) rho[gs] += __shfl_xor_sync(0xffffffffu, rho[gs], off);
            if (gid == 0) redr[par][warp][2 * tig + gs] = rho[gs];
        }
        __syncthreads();
#pragma unroll
        for (int gs = 0; gs < 2; ++gs) {                 // every thread sums the per-warp partials of its own two heads
            float x = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) x += redr[par][w][2 * tig + gs];
            rho[gs] = x;
        }
        // ---- sweep C
        {
            const uint16_t *sp = sp0, *dp = dp0;
            uint16_t* gp = gp0;
            int st = s0;
#pragma unroll T8_UNROLL_N
            for (; st < sf; ++st, sp += T8_CH, dp += T8_CH, gp += T8_CH) bwd_step_c<false>(sp, pS, dp, pA, gp, hA, 0, Nk, ldA, w1, ci, gww, gwl, rho, accWl);
            for (; st < s1; ++st, sp += T8_CH, dp += T8_CH, gp += T8_CH)
                bwd_step_c<true>(sp, pS, dp, pA, gp, hA, st * T8_CH + gid * 4, Nk, ldA, w1, ci, gww, gwl, rho, accWl);
        }
        __syncthreads();                                 // the row buffers are free again
        rq += gridDim.x;
        while (rq >= Nq) { rq -= Nq; ++rb; }
    }
    // per-CTA partials: [dWl (g, h) | dbl = 0 | dWw (o, g) | dbw = 0]  (talking_bwd_finalize_kernel's layout)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        atomicAdd(&spart[gid * 8 + 2 * tig + i], accWl[i]);
        atomicAdd(&spart[64 + 8 + gid * 8 + 2 * tig + i], accWw[i]);
    }
    __syncthreads();
    float* pr = part + (long long)blockIdx.x * NP;
    for (int i = tid; i < NP; i += NT) pr[i] = spart[i];
}

int t8_grid(int B, int Nq) {
    const long long rows = (long long)B * Nq;
    return (int)(rows < 2LL * spe_num_sms() ? rows : 2LL * spe_num_sms());
}
int t8_threads(const char* env, int dflt) {
    const char* e = getenv(env);
    const int v = e ? atoi(e) : dflt;
    return (v == 5 || v == 8 || v == 10 || v == 16) ? v : dflt;
}

}  // namespace

int spe_talking_h8_grid(int B, int Nq) { return t8_grid(B, Nq); }

// the staged kernels take fp16 logits only (the default storage); fp32 logits stay on rowwise.cu
bool spe_talking_h8_fits(long long ldS, long long ldA) {
    return (size_t)2 * 8 * (t8_pitch((int)ldS) + t8_pitch((int)ldA)) * 2 + 128 <= 110 * 1024;      // two CTAs per SM
}

int spe_talking_h8_fwd(const void* S, int s16, void* A, const float* Wl, const float* bl, const float* Ww, const float* bw, float* stats, int B, int Nq, int Nk,
                       long long ldS, long long ldA, cudaStream_t st) {
    SPE_CHECK(s16, "talking-heads H=8 (staged): fp16 logits only");
    SPE_CHECK(ldS % 8 == 0 && ldA % 8 == 0 && ldS >= Nk && ldA >= Nk && spe_talking_h8_fits(ldS, ldA), "talking-heads H=8: bad leading dimensions");
    static const int ctas = getenv("SPE_TH8_FWD_CTAS") ? atoi(getenv("SPE_TH8_FWD_CTAS")) : 3;      // resident CTAs per SM the grid is sized for (measured at cfg2: 2 x 10 warps 0.209 ms, 3 x 10 0.203, 3 x 8 0.189)
    const long long rows = (long long)B * Nq;
    const int grid = (int)(rows < (long long)ctas * spe_num_sms() ? rows : (long long)ctas * spe_num_sms());
    static const int nw = t8_threads("SPE_TH8_FWD_WARPS", 8);
    static const int nbuf_env = getenv("SPE_TH8_FWD_NBUF") ? atoi(getenv("SPE_TH8_FWD_NBUF")) : 2;     // deeper rings (3, 4 rows in flight) change nothing: not bound by bytes in flight
    const size_t row_bytes = (size_t)8 * t8_pitch((int)ldS) * 2;
    const int nbuf = (nbuf_env >= 4 && 4 * row_bytes + 128 <= 110 * 1024) ? 4 : (nbuf_env >= 3 && 3 * row_bytes + 128 <= 110 * 1024) ? 3 : 2;
    const size_t smem = nbuf * row_bytes + 128;      // + slack: the last step reads up to 24 keys past the pitch (masked)
    const uint16_t* S16p = reinterpret_cast<const uint16_t*>(S);
    uint16_t* A16 = reinterpret_cast<uint16_t*>(A);
#define T8_FWD(NW_, NB_)                                                                                                         \
    {                                                                                                                            \
        static bool attr = false;                                                                                                \
        if (!attr) { SPE_CUDA(cudaFuncSetAttribute(th8_fwd_kernel<NW_, NB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)); attr = true; } \
        th8_fwd_kernel<NW_, NB_><<<grid, NW_ * 32, smem, st>>>(S16p, A16, Wl, bl, Ww, bw, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA);      \
    }
#define T8_FWD_NB(NW_) { if (nbuf == 4) T8_FWD(NW_, 4) else if (nbuf == 3) T8_FWD(NW_, 3) else T8_FWD(NW_, 2) }
    if (nw == 5) T8_FWD_NB(5) else if (nw == 8) T8_FWD_NB(8) else if (nw == 16) T8_FWD_NB(16) else T8_FWD_NB(10)
#undef T8_FWD_NB
#undef T8_FWD
    SPE_LAUNCHED();
    return 0;
}

int spe_talking_h8_bwd(const void* S, int s16, const void* dA, void* dS, const float* Wl, const float* bl, const float* Ww, const float* stats, int B, int Nq,
                       int Nk, long long ldS, long long ldA, float* part, cudaStream_t st) {
    SPE_CHECK(s16, "talking-heads H=8 (staged): fp16 logits only");
    SPE_CHECK(stats, "talking-heads backward (H = 8) needs the forward statistics");
    SPE_CHECK(ldS % 8 == 0 && ldA % 8 == 0 && ldS >= Nk && ldA >= Nk && ldS < (1LL << 31) && ldA < (1LL << 31), "talking-heads H=8: bad leading dimensions");
    const int grid = t8_grid(B, Nq);
    const size_t smem = (size_t)2 * 8 * (t8_pitch((int)ldS) + t8_pitch((int)ldA)) * 2 + 128;
    static const int nw = t8_threads("SPE_TH8_BWD_WARPS", 8);
    const uint16_t* S16p = reinterpret_cast<const uint16_t*>(S);
    const uint16_t* dA16 = reinterpret_cast<const uint16_t*>(dA);
    uint16_t* dS16 = reinterpret_cast<uint16_t*>(dS);
#define T8_BWD(NW_)                                                                                                              \
    {                                                                                                                            \
        static bool attr = false;                                                                                                \
        if (!attr) { SPE_CUDA(cudaFuncSetAttribute(th8_bwd_kernel<NW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024)); attr = true; } \
        th8_bwd_kernel<NW_><<<grid, NW_ * 32, smem, st>>>(S16p, dA16, dS16, Wl, bl, Ww, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA, part); \
    }
    if (nw == 5) T8_BWD(5) else if (nw == 8) T8_BWD(8) else if (nw == 16) T8_BWD(16) else T8_BWD(10)
#undef T8_BWD
    SPE_LAUNCHED();
    return 0;
}
