// Row-wise / elementwise kernels of the backbone and transformer (HBM-bound: coalesced, vectorised,
// grid sized in multiples of the SM count).  Reference call sites: include/spe_b200.h.
#include "common.cuh"

namespace {

constexpr int LN_MAXCH = 8;   // D <= 1024 (float4 per lane per 128-column chunk)

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                            float eps, long long rows, int D, uint16_t* __restrict__ y16, float* __restrict__ y32,
                                                            float* __restrict__ mean_o, float* __restrict__ rstd_o) {
    const int lane = threadIdx.x & 31;
    const long long row0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long stride = (long long)gridDim.x * (blockDim.x >> 5);
    const int nch = (D + 127) / 128;
    for (long long row = row0; row < rows; row += stride) {
        const float* xr = x + row * D;
        float4 v[LN_MAXCH];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < LN_MAXCH; ++k) {
            const int c = k * 128 + lane * 4;
            if (k < nch && c < D) {
                v[k] = *reinterpret_cast<const float4*>(xr + c);
                s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
            } else v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float mean = warp_sum(s) / (float)D;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < LN_MAXCH; ++k) {
            const int c = k * 128 + lane * 4;
            if (k < nch && c < D) {
                const float a0 = v[k].x - mean, a1 = v[k].y - mean, a2 = v[k].z - mean, a3 = v[k].w - mean;
                q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
        if (lane == 0) { if (mean_o) mean_o[row] = mean; if (rstd_o) rstd_o[row] = rstd; }
#pragma unroll
        for (int k = 0; k < LN_MAXCH; ++k) {
            const int c = k * 128 + lane * 4;
            if (k < nch && c < D) {
                const float4 ww = __ldg(reinterpret_cast<const float4*>(w + c));
                const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
                float4 o;
                o.x = (v[k].x - mean) * rstd * ww.x + bb.x;
                o.y = (v[k].y - mean) * rstd * ww.y + bb.y;
                o.z = (v[k].z - mean) * rstd * ww.z + bb.z;
                o.w = (v[k].w - mean) * rstd * ww.w + bb.w;
                if (y32) *reinterpret_cast<float4*>(y32 + row * D + c) = o;
                if (y16) *reinterpret_cast<uint2*>(y16 + row * D + c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward: warp per row, per-lane column accumulators for dw/db, one atomic per column per CTA.
// NCH = number of 128-column chunks, compile time: the run-time-bounded version kept 4 x LN_MAXCH float4 arrays live (164 registers,
// ONE 8-warp CTA per SM -> ~5 MB of loads in flight chip-wide, a third of what HBM needs).  D = 384 (NCH 3) now runs 3 CTAs per SM.
// ------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(256, (NCH <= 2 ? 4 : (NCH == 3 ? 3 : (NCH <= 6 ? 2 : 1)))) layernorm_bwd_kernel(const uint16_t* __restrict__ dy16, const float* __restrict__ dy32,
                                                            const float* __restrict__ dres, const float* __restrict__ x,
                                                            const float* __restrict__ w, const float* __restrict__ mean_i,
                                                            const float* __restrict__ rstd_i, long long rows, int D, float* __restrict__ dx,
                                                            float* __restrict__ dw, float* __restrict__ db) {
    extern __shared__ float sm[];     // [2][D] per-CTA column sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) sm[c] = 0.f;
    __syncthreads();
    float4 aw[NCH], ab[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) { aw[k] = make_float4(0.f, 0.f, 0.f, 0.f); ab[k] = aw[k]; }
    for (long long row = (long long)blockIdx.x * nw + warp; row < rows; row += (long long)gridDim.x * nw) {
        const float mean = mean_i[row], rstd = rstd_i[row];
        float4 g[NCH], xh[NCH];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c = k * 128 + lane * 4;
            if (c < D) {
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                if (dy32) d = *reinterpret_cast<const float4*>(dy32 + row * D + c);
                if (dy16) {
                    const uint2 p = *reinterpret_cast<const uint2*>(dy16 + row * D + c);
                    const float2 a = unpack_bf16x2(p.x), bq = unpack_bf16x2(p.y);
                    d.x += a.x; d.y += a.y; d.z += bq.x; d.w += bq.y;
                }
                const float4 xv = *reinterpret_cast<const float4*>(x + row * D + c);
                const float4 ww = __ldg(reinterpret_cast<const float4*>(w + c));
                xh[k] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
                g[k] = make_float4(d.x * ww.x, d.y * ww.y, d.z * ww.z, d.w * ww.w);
                s1 += (g[k].x + g[k].y) + (g[k].z + g[k].w);
                s2 += (g[k].x * xh[k].x + g[k].y * xh[k].y) + (g[k].z * xh[k].z + g[k].w * xh[k].w);
                aw[k].x += d.x * xh[k].x; aw[k].y += d.y * xh[k].y; aw[k].z += d.z * xh[k].z; aw[k].w += d.w * xh[k].w;
                ab[k].x += d.x; ab[k].y += d.y; ab[k].z += d.z; ab[k].w += d.w;
            }
        }
        const float c1 = warp_sum(s1) / (float)D, c2 = warp_sum(s2) / (float)D;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c = k * 128 + lane * 4;
            if (c < D) {
                float4 o;
                o.x = rstd * (g[k].x - c1 - xh[k].x * c2);
                o.y = rstd * (g[k].y - c1 - xh[k].y * c2);
                o.z = rstd * (g[k].z - c1 - xh[k].z * c2);
                o.w = rstd * (g[k].w - c1 - xh[k].w * c2);
                if (dres) {
                    const float4 r = *reinterpret_cast<const float4*>(dres + row * D + c);
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                *reinterpret_cast<float4*>(dx + row * D + c) = o;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = k * 128 + lane * 4;
        if (c < D) {
            atomicAdd(&sm[c + 0], aw[k].x); atomicAdd(&sm[c + 1], aw[k].y); atomicAdd(&sm[c + 2], aw[k].z); atomicAdd(&sm[c + 3], aw[k].w);
            atomicAdd(&sm[D + c + 0], ab[k].x); atomicAdd(&sm[D + c + 1], ab[k].y); atomicAdd(&sm[D + c + 2], ab[k].z); atomicAdd(&sm[D + c + 3], ab[k].w);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        if (dw) atomicAdd(dw + c, sm[c]);
        if (db) atomicAdd(db + c, sm[D + c]);
    }
}

// ------------------------------------------------------------------------------------------------
// softmax over keys (optionally masked), fp32 in -> bf16 out; one CTA per (b,h,q) row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ S, uint16_t* __restrict__ P, const uint8_t* __restrict__ mask,
                                                          int H, int Nq, int Nk, long long ldS, long long ldP, float* __restrict__ pmean) {
    extern __shared__ float row[];
    __shared__ float red[32];
    const long long r = blockIdx.x;                  // (b*H + h)*Nq + q
    const int q = (int)(r % Nq);
    const int b = (int)(r / ((long long)H * Nq));
    const float* s = S + r * ldS;
    const uint8_t* mk = mask ? mask + (long long)b * Nk : nullptr;
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < Nk; j += blockDim.x) {
        float v = s[j];
        if (mk && mk[j]) v = -INFINITY;
        row[j] = v;
        mx = fmaxf(mx, v);
    }
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int j = threadIdx.x; j < Nk; j += blockDim.x) {
        const float e = __expf(row[j] - mx);
        row[j] = e;
        sum += e;
    }
    sum = block_sum(sum, red);
    const float inv = 1.f / sum;
    uint16_t* p = P + r * ldP;
    for (int j = threadIdx.x; j < Nk; j += blockDim.x) {
        const float v = row[j] * inv;
        p[j] = f_to_bf16(v);
        if (pmean) atomicAdd(pmean + ((long long)b * Nq + q) * Nk + j, v / (float)H);
    }
    for (int j = Nk + threadIdx.x; j < ldP; j += blockDim.x) p[j] = 0;      // keep the padding columns clean
}

__global__ void __launch_bounds__(256) softmax_bwd_kernel(const uint16_t* __restrict__ P, const uint16_t* __restrict__ dP, uint16_t* __restrict__ dS,
                                                          int Nk, long long ldP) {
    extern __shared__ float row[];    // [2][Nk]
    __shared__ float red[32];
    const long long r = blockIdx.x;
    const uint16_t* p = P + r * ldP;
    const uint16_t* dp = dP + r * ldP;
    float* pr = row;
    float* dr = row + Nk;
    float dot = 0.f;
    for (int j = threadIdx.x; j < Nk; j += blockDim.x) {
        const float a = bf16_to_f(p[j]), d = bf16_to_f(dp[j]);
        pr[j] = a; dr[j] = d;
        dot += a * d;
    }
    dot = block_sum(dot, red);
    uint16_t* o = dS + r * ldP;
    for (int j = threadIdx.x; j < Nk; j += blockDim.x) o[j] = f_to_bf16(pr[j] * (dr[j] - dot));
    for (int j = Nk + threadIdx.x; j < ldP; j += blockDim.x) o[j] = 0;
}

// ------------------------------------------------------------------------------------------------
// softmax, warp-per-row variants (Nk <= NV*128): the whole row lives in registers, all loads are issued up front, no
// shared memory and no block barriers.  The CTA-per-row kernels above remain for longer rows.
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256, 2) softmax_fwd_warp_kernel(const float* __restrict__ S, uint16_t* __restrict__ P, const uint8_t* __restrict__ mask,
                                                               long long rows_total, int H, int Nq, int Nk, long long ldS, long long ldP,
                                                               float* __restrict__ pmean) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows_total; r += warps) {
        const int q = (int)(r % Nq);
        const int b = (int)(r / ((long long)H * Nq));
        const float* s = S + r * ldS;
        const uint8_t* mk = mask ? mask + (long long)b * Nk : nullptr;
        float4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = 4 * lane + 128 * i;
            v[i] = (j + 4 <= ldS) ? *reinterpret_cast<const float4*>(s + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = 4 * lane + 128 * i;
            float* e = reinterpret_cast<float*>(&v[i]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (j + c >= Nk || (mk && mk[j + c])) e[c] = -INFINITY;
                mx = fmaxf(mx, e[c]);
            }
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float* e = reinterpret_cast<float*>(&v[i]);
#pragma unroll
            for (int c = 0; c < 4; ++c) { e[c] = __expf(e[c] - mx); sum += e[c]; }
        }
        const float inv = 1.f / warp_sum(sum);
        uint16_t* p = P + r * ldP;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = 4 * lane + 128 * i;
            if (j + 4 <= ldP) {
                const float a0 = v[i].x * inv, a1 = v[i].y * inv, a2 = v[i].z * inv, a3 = v[i].w * inv;     // exactly 0 beyond Nk
                *reinterpret_cast<uint2*>(p + j) = make_uint2(pack_bf16x2(a0, a1), pack_bf16x2(a2, a3));
                if (pmean) {
                    const float av[4] = {a0, a1, a2, a3};
#pragma unroll
                    for (int c = 0; c < 4; ++c) if (j + c < Nk) atomicAdd(pmean + ((long long)b * Nq + q) * Nk + j + c, av[c] / (float)H);
                }
            }
        }
    }
}

template <int NV>
__global__ void __launch_bounds__(256, 2) softmax_bwd_warp_kernel(const uint16_t* __restrict__ P, const uint16_t* dP, uint16_t* dS, long long rows_total, int Nk,
                                                               long long ldP) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows_total; r += warps) {
        const uint16_t* p = P + r * ldP;
        const uint16_t* dp = dP + r * ldP;
        uint2 pv[NV], dv[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = 4 * lane + 128 * i;
            const bool in = j + 4 <= ldP;
            pv[i] = in ? *reinterpret_cast<const uint2*>(p + j) : make_uint2(0u, 0u);
            dv[i] = in ? *reinterpret_cast<const uint2*>(dp + j) : make_uint2(0u, 0u);
        }
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = 4 * lane + 128 * i;
            const float2 p0 = unpack_bf16x2(pv[i].x), p1 = unpack_bf16x2(pv[i].y), d0 = unpack_bf16x2(dv[i].x), d1 = unpack_bf16x2(dv[i].y);
            // P is exactly 0 in the padding; dP's padding was never written -> guard the products
            dot += (j < Nk ? p0.x * d0.x : 0.f) + (j + 1 < Nk ? p0.y * d0.y : 0.f) + (j + 2 < Nk ? p1.x * d1.x : 0.f) + (j + 3 < Nk ? p1.y * d1.y : 0.f);
        }
        dot = warp_sum(dot);
        uint16_t* o = dS + r * ldP;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = 4 * lane + 128 * i;
            if (j + 4 <= ldP) {
                const float2 p0 = unpack_bf16x2(pv[i].x), p1 = unpack_bf16x2(pv[i].y), d0 = unpack_bf16x2(dv[i].x), d1 = unpack_bf16x2(dv[i].y);
                const float a0 = j < Nk ? p0.x * (d0.x - dot) : 0.f, a1 = j + 1 < Nk ? p0.y * (d0.y - dot) : 0.f;
                const float a2 = j + 2 < Nk ? p1.x * (d1.x - dot) : 0.f, a3 = j + 3 < Nk ? p1.y * (d1.y - dot) : 0.f;
                *reinterpret_cast<uint2*>(o + j) = make_uint2(pack_bf16x2(a0, a1), pack_bf16x2(a2, a3));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// talking-heads mix -> softmax -> mix   (cait.py:381-386)
//   ONE WARP PER (b, q) ROW, streaming: no block-level barriers, no staging of the H x Nk slab.
//   fwd : sweep A = logits L (head mix) with an online (max, sum) per mixed head; sweep B re-reads S (L2-hot),
//         recomputes L, P = exp(L-m)/z, second head mix, bf16 store.
//   bwd : sweep A as above; sweep B: P, dP = Ww^T dA, rho = sum_j P dP (+ dWw, dbw partials);
//         sweep C: dL = P (dP - rho), dS = Wl^T dL (+ dWl, dbl partials).  Parameter-gradient partials stay in
//         registers across all rows of a CTA and are reduced once at the end (deterministic two-stage reduction).
//   The H x H mixes are FP32 FMAs against weights broadcast from shared memory (FP32/MUFU bound).
// ------------------------------------------------------------------------------------------------
// =================================================================================================
// Talking-heads kernels, tensor-path formulation.
// One warp per (b, q) row, 32 keys per step.  Every H x H head mix runs as mma.sync.m16n8k16 (bf16 inputs, fp32
// accumulate) with split-precision operands so the result keeps ~16 mantissa bits:
//     A (16 x 16) = [ W_hi | W_hi ]   rows 0-7       B (16 x 8) = [ x_hi ]  k 0-7   (x = logits / probabilities, fp32)
//                   [ W_lo | W_lo ]   rows 8-15                   [ x_lo ]  k 8-15
//     D rows 0-7 + D rows 8-15  =  (W_hi + W_lo)(x_hi + x_lo)
// Lane (q4 = lane/4, r4 = lane%4) owns head q4 and 8 contiguous keys [base + 8 r4, +8) in the accumulator layout, and heads
// (2 r4, 2 r4 + 1) at 4 keys [base + 4 q4, +4) in the B-operand layout; tile t of a step maps mma column n to key
// base + 4 n + t, so both layouts are contiguous in memory.  movmatrix.trans converts accumulator -> B-operand layout.
// The H x H parameter-gradient outer products are mma.sync accumulations over the keys as well.
// =================================================================================================
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movm_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
// split (x, y) into packed bf16 hi parts and packed bf16 lo (residual) parts
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(x, y);
    const float2 h = unpack_bf16x2(hi);
    lo = pack_bf16x2(x - h.x, y - h.y);
}

struct MixFrag { uint32_t hi, lo; };     // A-operand registers of one H x H matrix M[row][k]: a0 = a2 = hi, a1 = a3 = lo

template <int H>
__device__ __forceinline__ MixFrag load_mix_frag(const float* __restrict__ M, bool transpose, float scale, int q4, int r4) {
    float w0 = 0.f, w1 = 0.f;
    if (q4 < H) {
        const int k0 = 2 * r4, k1 = 2 * r4 + 1;
        if (k0 < H) w0 = scale * (transpose ? M[k0 * H + q4] : M[q4 * H + k0]);
        if (k1 < H) w1 = scale * (transpose ? M[k1 * H + q4] : M[q4 * H + k1]);
    }
    MixFrag f;
    split2(w0, w1, f.hi, f.lo);
    return f;
}
// x (B-operand layout, split precision) -> W x for the lane's head q4 at mma columns 2 r4, 2 r4 + 1
__device__ __forceinline__ void mix_tile(const MixFrag& W, uint32_t bhi, uint32_t blo, float bias, float& o0, float& o1) {
    float d[4] = {bias, bias, 0.f, 0.f};
    mma16816(d, W.hi, W.lo, W.hi, W.lo, bhi, blo);
    o0 = d[0] + d[2]; o1 = d[1] + d[3];
}

// ---- single-pass fp16 variant of the two FORWARD mixes (logits S and probabilities P) ----
// The split-precision scheme above keeps ~16 mantissa bits; the forward does not need them: S carries ~2^-9 relative noise from its bf16
// Q / K inputs already, fp16 rounding adds 2^-11 (tools/tf32_mix_numerics.py: +2e-3 on the attention output against a 3e-3..1e-2
// floor), and A = Ww P is rounded to bf16 on output anyway.  One cvt.rn.f16x2 per operand pair replaces pack / unpack / subtract / pack
// and the lo-part movmatrix.  (The backward keeps the bf16 split: its dA / dL operands need the bf16 exponent range.)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void mma16816_f16(float (&d)[4], uint32_t a0, uint32_t b0) {       // rows 8-15 of A and k 8-15 are zero
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(0u), "r"(0u), "r"(0u), "r"(b0), "r"(0u));
}
template <int H>
__device__ __forceinline__ uint32_t load_mix_frag16(const float* __restrict__ M, float scale, int q4, int r4) {
    float w0 = 0.f, w1 = 0.f;
    if (q4 < H) {
        const int k0 = 2 * r4, k1 = 2 * r4 + 1;
        if (k0 < H) w0 = scale * M[q4 * H + k0];
        if (k1 < H) w1 = scale * M[q4 * H + k1];
    }
    return pack_f16x2(w0, w1);
}
__device__ __forceinline__ void mix_tile16(uint32_t W16, uint32_t b16, float bias, float& o0, float& o1) {
    float d[4] = {bias, bias, 0.f, 0.f};
    mma16816_f16(d, W16, b16);
    o0 = d[0]; o1 = d[1];
}

constexpr float LOG2E = 1.4426950408889634f;
__device__ __forceinline__ float fast_ex2(float x) {           // single MUFU.EX2 (flushes denormals; 2^-inf = 0)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// raw S of one 32-key step in the B-operand layout: 4 keys [base + 4 q4, +4) of heads 2 r4 and 2 r4 + 1
struct SRaw { float4 s0, s1; };
template <int H>
__device__ __forceinline__ SRaw load_sraw(const float* __restrict__ Sb, long long hS, int base, long long ldS, int q4, int r4) {
    SRaw r;
    r.s0 = make_float4(0.f, 0.f, 0.f, 0.f); r.s1 = r.s0;
    const int cb = base + 4 * q4;
    if (cb + 4 <= ldS) {
        if (2 * r4 < H) r.s0 = *reinterpret_cast<const float4*>(Sb + (2 * r4) * hS + cb);
        if (2 * r4 + 1 < H) r.s1 = *reinterpret_cast<const float4*>(Sb + (2 * r4 + 1) * hS + cb);
    }
    return r;
}
// logits of one 32-key step in the accumulator layout: L2[i] = log2(e) * L[q4][base + 8 r4 + i]   (-inf beyond Nk)
// also returns the hi parts of the split B-operand registers of S (reused by the dWl outer product)
template <int H, bool TAIL = true>
__device__ __forceinline__ void step_logits(const SRaw& raw, int base, int Nk, const MixFrag& Wl, float bl2, int q4, int r4, float (&L2)[8],
                                            uint32_t (&shi)[4]) {
    const int cb = base + 4 * q4;
    const float a0[4] = {raw.s0.x, raw.s0.y, raw.s0.z, raw.s0.w}, a1[4] = {raw.s1.x, raw.s1.y, raw.s1.z, raw.s1.w};
    const bool tail = TAIL && base + 32 > Nk;       // warp-uniform: only the last step of a row has invalid keys
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const bool ok = !tail || cb + t < Nk;       // padding columns of S are never written: sanitise
        uint32_t lo;
        split2(ok ? a0[t] : 0.f, ok ? a1[t] : 0.f, shi[t], lo);
        mix_tile(Wl, shi[t], lo, bl2, L2[t], L2[4 + t]);
    }
    if (tail) {
        const int c8 = base + 8 * r4;
#pragma unroll
        for (int i = 0; i < 8; ++i) if (c8 + i >= Nk) L2[i] = -INFINITY;
    }
}

// fp16 single-pass version of step_logits (forward only)
template <int H, bool TAIL = true>
__device__ __forceinline__ void step_logits16(const SRaw& raw, int base, int Nk, uint32_t Wl16, float bl2, int q4, int r4, float (&L2)[8]) {
    const int cb = base + 4 * q4;
    const float a0[4] = {raw.s0.x, raw.s0.y, raw.s0.z, raw.s0.w}, a1[4] = {raw.s1.x, raw.s1.y, raw.s1.z, raw.s1.w};
    const bool tail = TAIL && base + 32 > Nk;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const bool ok = !tail || cb + t < Nk;       // padding columns of S are never written: sanitise
        mix_tile16(Wl16, pack_f16x2(ok ? a0[t] : 0.f, ok ? a1[t] : 0.f), bl2, L2[t], L2[4 + t]);
    }
    if (tail) {
        const int c8 = base + 8 * r4;
#pragma unroll
        for (int i = 0; i < 8; ++i) if (c8 + i >= Nk) L2[i] = -INFINITY;
    }
}

// backward flavour: logits through the fp16 single pass (identical to what the forward computed its statistics from), plus the
// bf16-packed S of the step in the B-operand layout for the dWl outer product
template <int H, bool TAIL = true>
__device__ __forceinline__ void step_logits16_shi(const SRaw& raw, int base, int Nk, uint32_t Wl16, float bl2, int q4, int r4, float (&L2)[8], uint32_t (&shi)[4]) {
    const int cb = base + 4 * q4;
    const float a0[4] = {raw.s0.x, raw.s0.y, raw.s0.z, raw.s0.w}, a1[4] = {raw.s1.x, raw.s1.y, raw.s1.z, raw.s1.w};
    const bool tail = TAIL && base + 32 > Nk;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const bool ok = !tail || cb + t < Nk;
        const float x0 = ok ? a0[t] : 0.f, x1 = ok ? a1[t] : 0.f;
        shi[t] = pack_bf16x2(x0, x1);
        mix_tile16(Wl16, pack_f16x2(x0, x1), bl2, L2[t], L2[4 + t]);
    }
    if (tail) {
        const int c8 = base + 8 * r4;
#pragma unroll
        for (int i = 0; i < 8; ++i) if (c8 + i >= Nk) L2[i] = -INFINITY;
    }
}

// sweep A: m2[q4] = max_j L2, iz = 1 / sum_j 2^(L2 - m2)     (values for the lane's head q4, identical in its 4 lanes)
template <int H>
__device__ __forceinline__ void talking_stats2(const float* __restrict__ Sb, long long hS, int Nk, long long ldS, const MixFrag& Wl, float bl2, int q4,
                                               int r4, float& m2, float& iz) {
    float m = -INFINITY, z = 0.f;
    SRaw nxt = load_sraw<H>(Sb, hS, 0, ldS, q4, r4);
    for (int base = 0; base < Nk; base += 32) {
        const SRaw cur = nxt;
        if (base + 32 < Nk) nxt = load_sraw<H>(Sb, hS, base + 32, ldS, q4, r4);      // prefetch one step ahead
        float L2[8];
        uint32_t shi[4];
        step_logits<H>(cur, base, Nk, Wl, bl2, q4, r4, L2, shi);
        float mx = L2[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) mx = fmaxf(mx, L2[i]);
        const float mn = fmaxf(m, mx);
        if (mn > -INFINITY) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += exp2f(L2[i] - mn);
            z = z * exp2f(m - mn) + acc;
            m = mn;
        }
    }
    float M = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 2));
    float zz = (m == -INFINITY) ? 0.f : z * exp2f(m - M);
    zz += __shfl_xor_sync(0xffffffffu, zz, 1);
    zz += __shfl_xor_sync(0xffffffffu, zz, 2);
    m2 = M;
    iz = 1.f / zz;
}

template <int H>
__global__ void __launch_bounds__(256, 3) talking_fwd_kernel(const float* __restrict__ S, uint16_t* __restrict__ A, const float* __restrict__ Wl,
                                                             const float* __restrict__ bl, const float* __restrict__ Ww, const float* __restrict__ bw,
                                                             float* __restrict__ stats, int rows_total, int Nq, int Nk, long long ldS, long long ldA) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q4 = lane >> 2, r4 = lane & 3;
    const MixFrag fWl = load_mix_frag<H>(Wl, false, LOG2E, q4, r4), fWw = load_mix_frag<H>(Ww, false, 1.f, q4, r4);
    const float bl2 = q4 < H ? bl[q4] * LOG2E : 0.f, bwv = q4 < H ? bw[q4] : 0.f;
    const long long hS = (long long)Nq * ldS, hA = (long long)Nq * ldA;
    for (int rowi = blockIdx.x * 8 + warp; rowi < rows_total; rowi += gridDim.x * 8) {
        const int b = rowi / Nq, q = rowi % Nq;
        const float* Sb = S + ((long long)b * H * Nq + q) * ldS;
        uint16_t* Ab = A + ((long long)b * H * Nq + q) * ldA;
        float m2, iz;
        talking_stats2<H>(Sb, hS, Nk, ldS, fWl, bl2, q4, r4, m2, iz);
        const float c2 = m2 - log2f(iz);            // p = 2^(L2 - m2) * iz = 2^(L2 - c2)
        if (stats && r4 == 0 && q4 < H) stats[(long long)rowi * H + q4] = c2;
        SRaw nxt = load_sraw<H>(Sb, hS, 0, ldS, q4, r4);
        for (int base = 0; base < ldA; base += 32) {
            const SRaw cur = nxt;
            if (base + 32 < ldA) nxt = load_sraw<H>(Sb, hS, base + 32, ldS, q4, r4);
            float L2[8], out[8];
            uint32_t shi[4];
            step_logits<H>(cur, base, Nk, fWl, bl2, q4, r4, L2, shi);
            float p[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = exp2f(L2[i] - c2);          // 0 beyond Nk (L2 = -inf)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                uint32_t hi, lo;
                split2(p[t], p[4 + t], hi, lo);
                mix_tile(fWw, movm_trans(hi), movm_trans(lo), bwv, out[t], out[4 + t]);
            }
            const int c8 = base + 8 * r4;
            if (q4 < H && c8 + 8 <= ldA) {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (c8 + i >= Nk) out[i] = 0.f;    // keep the padding columns [Nk, ldA) zero
                *reinterpret_cast<uint4*>(Ab + q4 * hA + c8) =
                    make_uint4(pack_bf16x2(out[0], out[1]), pack_bf16x2(out[2], out[3]), pack_bf16x2(out[4], out[5]), pack_bf16x2(out[6], out[7]));
            }
        }
    }
}

// backward: dA (bf16) -> dS (bf16, may alias dA), parameter-gradient partials per CTA in `part` [gridDim.x][2*H*H + 2*H]
template <int H>
__global__ void __launch_bounds__(256, 2) talking_bwd_kernel(const float* __restrict__ S, const uint16_t* dA, uint16_t* dS,
                                                             const float* __restrict__ Wl, const float* __restrict__ bl, const float* __restrict__ Ww,
                                                             int rows_total, int Nq, int Nk, long long ldS, long long ldA, float* __restrict__ part) {
    constexpr int NP = 2 * H * H + 2 * H;
    __shared__ float redbuf[8][NP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q4 = lane >> 2, r4 = lane & 3;
    const MixFrag fWl = load_mix_frag<H>(Wl, false, LOG2E, q4, r4);      // L2 = log2e (Wl S + bl)
    const MixFrag fWwT = load_mix_frag<H>(Ww, true, 1.f, q4, r4);        // dP = Ww^T dA
    const MixFrag fWlT = load_mix_frag<H>(Wl, true, 1.f, q4, r4);        // dS = Wl^T dL
    const float bl2 = q4 < H ? bl[q4] * LOG2E : 0.f;
    const long long hS = (long long)Nq * ldS, hA = (long long)Nq * ldA;

    float accWw[4] = {0.f, 0.f, 0.f, 0.f}, accWl[4] = {0.f, 0.f, 0.f, 0.f};     // [row q4][col 2 r4 + {0,1}] of dWw[o][g], dWl[g][h]
    float abw = 0.f, abl = 0.f;                                                  // dbw[q4], dbl[q4] lane partials

    // dA of one step in the accumulator layout (head q4, 8 contiguous keys), sanitised beyond Nk
    auto load_dA_raw = [&](const uint16_t* dAb, int base) {
        const int c8 = base + 8 * r4;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (q4 < H && c8 + 8 <= ldA) v = *reinterpret_cast<const uint4*>(dAb + q4 * hA + c8);
        return v;
    };
    auto unpack_dA = [&](const uint4& v, int base, float (&d)[8]) {
        const int c8 = base + 8 * r4;
        float2 t;
        t = unpack_bf16x2(v.x); d[0] = t.x; d[1] = t.y;
        t = unpack_bf16x2(v.y); d[2] = t.x; d[3] = t.y;
        t = unpack_bf16x2(v.z); d[4] = t.x; d[5] = t.y;
        t = unpack_bf16x2(v.w); d[6] = t.x; d[7] = t.y;
#pragma unroll
        for (int i = 0; i < 8; ++i) if (c8 + i >= Nk) d[i] = 0.f;
    };

    for (int rowi = blockIdx.x * 8 + warp; rowi < rows_total; rowi += gridDim.x * 8) {
        const int b = rowi / Nq, q = rowi % Nq;
        const float* Sb = S + ((long long)b * H * Nq + q) * ldS;
        const uint16_t* dAb = dA + ((long long)b * H * Nq + q) * ldA;
        uint16_t* dSb = dS + ((long long)b * H * Nq + q) * ldA;
        float m2, iz;
        talking_stats2<H>(Sb, hS, Nk, ldS, fWl, bl2, q4, r4, m2, iz);
        const float c2 = m2 - log2f(iz);
        // ---- sweep B: rho = sum_j P dP;  dWw += dA (x) P;  dbw += dA
        float rho = 0.f;
        SRaw nxt = load_sraw<H>(Sb, hS, 0, ldS, q4, r4);
        uint4 dnxt = load_dA_raw(dAb, 0);
        for (int base = 0; base < Nk; base += 32) {
            const SRaw cur = nxt;
            const uint4 dcur = dnxt;
            if (base + 32 < Nk) { nxt = load_sraw<H>(Sb, hS, base + 32, ldS, q4, r4); dnxt = load_dA_raw(dAb, base + 32); }
            float L2[8], d[8], dP[8], p[8];
            uint32_t shi[4];
            step_logits<H>(cur, base, Nk, fWl, bl2, q4, r4, L2, shi);
            unpack_dA(dcur, base, d);
            uint32_t dpk[4], ppk[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                p[t] = exp2f(L2[t] - c2); p[4 + t] = exp2f(L2[4 + t] - c2);
                dpk[t] = pack_bf16x2(d[t], d[4 + t]);
                ppk[t] = pack_bf16x2(p[t], p[4 + t]);
                mix_tile(fWwT, movm_trans(dpk[t]), 0u, 0.f, dP[t], dP[4 + t]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { rho += p[i] * dP[i]; abw += d[i]; }
            // outer products over this step's 32 keys: rows = dA heads, cols = P heads (the key order inside the k dimension
            // is the same arbitrary permutation for both operands)
            mma16816(accWw, dpk[0], 0u, dpk[1], 0u, ppk[0], ppk[1]);
            mma16816(accWw, dpk[2], 0u, dpk[3], 0u, ppk[2], ppk[3]);
        }
        rho += __shfl_xor_sync(0xffffffffu, rho, 1);
        rho += __shfl_xor_sync(0xffffffffu, rho, 2);
        // ---- sweep C: dL = P (dP - rho);  dS = Wl^T dL;  dWl += dL (x) S;  dbl += dL
        nxt = load_sraw<H>(Sb, hS, 0, ldS, q4, r4);
        dnxt = load_dA_raw(dAb, 0);
        for (int base = 0; base < ldA; base += 32) {
            const SRaw cur = nxt;
            const uint4 dcur = dnxt;
            // (in-place dS: the prefetched dA of the next step is read before this step's store; different keys anyway)
            if (base + 32 < ldA) { nxt = load_sraw<H>(Sb, hS, base + 32, ldS, q4, r4); dnxt = load_dA_raw(dAb, base + 32); }
            float L2[8], d[8], dP[8], l[8], o[8];
            uint32_t shi[4];
            step_logits<H>(cur, base, Nk, fWl, bl2, q4, r4, L2, shi);
            unpack_dA(dcur, base, d);
            uint32_t lpk[4], spk[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) mix_tile(fWwT, movm_trans(pack_bf16x2(d[t], d[4 + t])), 0u, 0.f, dP[t], dP[4 + t]);
#pragma unroll
            for (int i = 0; i < 8; ++i) { l[i] = exp2f(L2[i] - c2) * (dP[i] - rho); abl += l[i]; }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                lpk[t] = pack_bf16x2(l[t], l[4 + t]);
                spk[t] = movm_trans(shi[t]);                               // S (hi part) back in the accumulator layout: head q4
                mix_tile(fWlT, movm_trans(lpk[t]), 0u, 0.f, o[t], o[4 + t]);
            }
            mma16816(accWl, lpk[0], 0u, lpk[1], 0u, spk[0], spk[1]);
            mma16816(accWl, lpk[2], 0u, lpk[3], 0u, spk[2], spk[3]);
            const int c8 = base + 8 * r4;
            if (q4 < H && c8 + 8 <= ldA) {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (c8 + i >= Nk) o[i] = 0.f;
                // all dA reads of this (row, keys) happened above (same lanes) -> in-place overwrite is safe when dS aliases dA
                *reinterpret_cast<uint4*>(dSb + q4 * hA + c8) =
                    make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
            }
        }
    }
    // CTA reduction of the partials -> part[blockIdx.x][...]: layout dWl[H*H], dbl[H], dWw[H*H], dbw[H]
    for (int i = lane; i < NP; i += 32) redbuf[warp][i] = 0.f;
    __syncwarp();
    abw += __shfl_xor_sync(0xffffffffu, abw, 1); abw += __shfl_xor_sync(0xffffffffu, abw, 2);
    abl += __shfl_xor_sync(0xffffffffu, abl, 1); abl += __shfl_xor_sync(0xffffffffu, abl, 2);
    if (q4 < H) {
        const int col = 2 * r4;
        if (col < H) { redbuf[warp][q4 * H + col] = accWl[0]; redbuf[warp][H * H + H + q4 * H + col] = accWw[0]; }
        if (col + 1 < H) { redbuf[warp][q4 * H + col + 1] = accWl[1]; redbuf[warp][H * H + H + q4 * H + col + 1] = accWw[1]; }
        if (r4 == 0) { redbuf[warp][H * H + q4] = abl; redbuf[warp][2 * H * H + H + q4] = abw; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NP; i += 256) {
        float sum = 0.f;
        for (int w2 = 0; w2 < 8; ++w2) sum += redbuf[w2][i];
        part[(long long)blockIdx.x * NP + i] = sum;
    }
}

// =================================================================================================
// Row-staged variants (the default): ONE persistent CTA per SM slot, 8 warps share one (b, q) row.  The row's operands
// (S: H x ld fp32, and dA: H x ld bf16 in the backward) are brought into shared memory ONCE by bulk async copies
// (cp.async.bulk + mbarrier), double buffered so row r+1 streams in while row r is processed; the 2-3 sweeps then read
// shared memory only.  SM<->L2 traffic per row drops from 3 x (S + dA) to 1 x, which is what bounds these kernels.
// The keys of a row are split into 8 contiguous segments (one per warp); softmax statistics / rho are combined through
// shared memory.  The forward also saves the statistics so the backward needs no statistics sweep.
// =================================================================================================
__device__ __forceinline__ uint32_t rw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rw_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void rw_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rw_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}
__device__ __forceinline__ void rw_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// shared-memory row pitches of the row-staged kernels.  fp32 rows: pitch = 4 (mod 16) words, so the four heads a quarter-warp touches
// in the B-operand layout (rows 2 r4) and the two heads it touches in the accumulator layout fall into disjoint bank groups
// (N = 1600 unpadded put every head on the same banks: 56% of the shared wavefronts were conflicts).  bf16 rows: pitch = 32 (mod 64).
__host__ __device__ __forceinline__ int talking_pitch_f32(int ld) { return ld + ((20 - ld % 16) % 16); }
__host__ __device__ __forceinline__ int talking_pitch_bf16(int ld) { return ld + ((96 - ld % 64) % 64); }
// fp16 logit rows (S kept in fp16 in HBM: half the N^2 traffic of the fp32 logits; the head mix rounds S to fp16 anyway, so the
// forward is bit-identical): pitch = 16 (mod 32) halves, i.e. consecutive head PAIRS (the B-operand layout reads heads 2 r4, 2 r4 + 1)
// start 64 bytes apart modulo 128 -- the 8-byte lane reads of r4 = 0 / 1 and of r4 = 2 / 3 share a wavefront without conflicts
__host__ __device__ __forceinline__ int talking_pitch_f16(int ld) { return ld + ((48 - ld % 32) % 32); }

// raw fp16 S of one 32-key step in the B-operand layout: 4 keys [base + 4 q4, +4) of heads 2 r4 (h0) and 2 r4 + 1 (h1)
struct SRaw16 { uint2 h0, h1; };
template <int H>
__device__ __forceinline__ SRaw16 load_sraw16(const uint16_t* __restrict__ Sb, int pH, int base, int ldS, int q4, int r4) {
    SRaw16 r;
    r.h0 = make_uint2(0u, 0u); r.h1 = r.h0;
    const int cb = base + 4 * q4;
    if (cb + 4 <= ldS) {
        if (2 * r4 < H) r.h0 = *reinterpret_cast<const uint2*>(Sb + (2 * r4) * pH + cb);
        if (2 * r4 + 1 < H) r.h1 = *reinterpret_cast<const uint2*>(Sb + (2 * r4 + 1) * pH + cb);
    }
    return r;
}
// (head 2 r4, head 2 r4 + 1) of key t as one f16x2 register -- the mix operand, straight from the stored halves
__device__ __forceinline__ uint32_t sraw16_pair(const SRaw16& r, int t) {
    const uint32_t a = t < 2 ? r.h0.x : r.h0.y, b = t < 2 ? r.h1.x : r.h1.y;
    return __byte_perm(a, b, (t & 1) ? 0x7632 : 0x5410);
}
__device__ __forceinline__ float f16_bits_to_f(uint32_t h) {
    float f;
    asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"((uint16_t)h));
    return f;
}

// ---- per-step bodies (TAIL = false: all 32 keys of the step are valid, no masking code at all) ----
// forward sweep A: mixed logits of one step -> online (max, sum); the logits are written back over S in shared memory
// (accumulator layout: head q4, keys base + 8 r4 .. +7) so sweep B needs neither the split nor the mix again.
template <int H, bool TAIL>
__device__ __forceinline__ void tfwd_step_a(float* Sb, int pS, int ldS, int base, int Nk, uint32_t fWl, float bl2, int q4, int r4, float& m, float& z) {
    float L2[8];
    step_logits16<H, TAIL>(load_sraw<H>(Sb, pS, base, ldS, q4, r4), base, Nk, fWl, bl2, q4, r4, L2);
    __syncwarp();                                                    // every lane's reads of this step's S precede the overwrite
    const int c8 = base + 8 * r4;
    if (q4 < H && c8 + 8 <= ldS) {
        float4* d = reinterpret_cast<float4*>(Sb + q4 * pS + c8);
        d[0] = make_float4(L2[0], L2[1], L2[2], L2[3]);
        d[1] = make_float4(L2[4], L2[5], L2[6], L2[7]);
    }
    float mx = fmaxf(fmaxf(fmaxf(L2[0], L2[1]), fmaxf(L2[2], L2[3])), fmaxf(fmaxf(L2[4], L2[5]), fmaxf(L2[6], L2[7])));
    const float mn = fmaxf(m, mx);
    if (!TAIL || mn > -INFINITY) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += fast_ex2(L2[i] - mn);
        z = z * fast_ex2(m - mn) + acc;
        m = mn;
    }
}
// forward sweep B: P = 2^(L2 - c2) from the cached logits -> second mix -> bf16
template <int H, bool TAIL>
__device__ __forceinline__ void tfwd_step_b(const float* Sb, int pS, int ldS, int base, int Nk, uint32_t fWw, float bwv, float c2, uint16_t* Ab, long long hA,
                                            int ldA, int q4, int r4) {
    const int c8 = base + 8 * r4;
    float p[8], out[8];
    if (q4 < H && c8 + 8 <= ldS) {
        const float4 u = *reinterpret_cast<const float4*>(Sb + q4 * pS + c8), w = *reinterpret_cast<const float4*>(Sb + q4 * pS + c8 + 4);
        p[0] = fast_ex2(u.x - c2); p[1] = fast_ex2(u.y - c2); p[2] = fast_ex2(u.z - c2); p[3] = fast_ex2(u.w - c2);
        p[4] = fast_ex2(w.x - c2); p[5] = fast_ex2(w.y - c2); p[6] = fast_ex2(w.z - c2); p[7] = fast_ex2(w.w - c2);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) mix_tile16(fWw, movm_trans(pack_f16x2(p[t], p[4 + t])), bwv, out[t], out[4 + t]);
    if (q4 < H && c8 + 8 <= ldA) {
        if (TAIL) {
#pragma unroll
            for (int i = 0; i < 8; ++i) if (c8 + i >= Nk) out[i] = 0.f;
        }
        *reinterpret_cast<uint4*>(Ab + q4 * hA + c8) =
            make_uint4(pack_bf16x2(out[0], out[1]), pack_bf16x2(out[2], out[3]), pack_bf16x2(out[4], out[5]), pack_bf16x2(out[6], out[7]));
    }
}

// fp16-S flavour of sweep A: the raw logits come from the fp16 row buffer, the mixed logits go to a separate fp32 cache
template <int H, bool TAIL>
__device__ __forceinline__ void tfwd_step_a16(const uint16_t* S16b, int pH, float* Lb, int pS, int ldS, int base, int Nk, uint32_t fWl, float bl2, int q4, int r4,
                                              float& m, float& z) {
    float L2[8];
    const SRaw16 raw = load_sraw16<H>(S16b, pH, base, ldS, q4, r4);
    const int cb = base + 4 * q4;
    const bool tail = TAIL && base + 32 > Nk;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const bool ok = !tail || cb + t < Nk;       // padding columns of S are never written: sanitise
        mix_tile16(fWl, ok ? sraw16_pair(raw, t) : 0u, bl2, L2[t], L2[4 + t]);
    }
    const int c8 = base + 8 * r4;
    if (tail) {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (c8 + i >= Nk) L2[i] = -INFINITY;
    }
    if (q4 < H && c8 + 8 <= ldS) {
        float4* d = reinterpret_cast<float4*>(Lb + q4 * pS + c8);
        d[0] = make_float4(L2[0], L2[1], L2[2], L2[3]);
        d[1] = make_float4(L2[4], L2[5], L2[6], L2[7]);
    }
    float mx = fmaxf(fmaxf(fmaxf(L2[0], L2[1]), fmaxf(L2[2], L2[3])), fmaxf(fmaxf(L2[4], L2[5]), fmaxf(L2[6], L2[7])));
    const float mn = fmaxf(m, mx);
    if (!TAIL || mn > -INFINITY) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += fast_ex2(L2[i] - mn);
        z = z * fast_ex2(m - mn) + acc;
        m = mn;
    }
}

// NW warps share one row; the steps of a row are dealt out evenly (NW = 10 fits the 50 steps of N = 1600 exactly).
// S16: S is fp16 in HBM (row staged by the same bulk copies at half the bytes); the fp32 cache of the mixed logits is then a
// separate single buffer (the in-place overwrite of the fp32 flavour needs equal element sizes).
template <int H, int NW, bool S16>
__global__ void __launch_bounds__(NW * 32, 2) talking_fwd_rows_kernel(const void* __restrict__ Sv, uint16_t* __restrict__ A, const float* __restrict__ Wl,
                                                                      const float* __restrict__ bl, const float* __restrict__ Ww, const float* __restrict__ bw,
                                                                      float* __restrict__ stats, int rows_total, int Nq, int Nk, int ldS, int ldA) {
    extern __shared__ __align__(128) uint8_t rsm[];
    const int pS = talking_pitch_f32(ldS);                             // padded row pitch: conflict-free in both fragment layouts
    const int pH = talking_pitch_f16(ldS);
    float* Sbuf = reinterpret_cast<float*>(rsm);                       // fp32 S: [2][H][pS] (logits cached in place);  fp16 S: [H][pS] logit cache
    uint16_t* S16buf = reinterpret_cast<uint16_t*>(rsm + (size_t)H * pS * 4);     // fp16 S: [2][H][pH]
    const float* S = reinterpret_cast<const float*>(Sv);
    const uint16_t* Sh = reinterpret_cast<const uint16_t*>(Sv);
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ float red[NW][8][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q4 = lane >> 2, r4 = lane & 3;
    const uint32_t fWl = load_mix_frag16<H>(Wl, LOG2E, q4, r4), fWw = load_mix_frag16<H>(Ww, 1.f, q4, r4);     // fp16 single-pass mixes (forward)
    const float bl2 = q4 < H ? bl[q4] * LOG2E : 0.f, bwv = q4 < H ? bw[q4] : 0.f;
    const long long hS = (long long)Nq * ldS, hA = (long long)Nq * ldA;
    const uint32_t bar0 = rw_smem_u32(bars);
    if (tid == 0) {
        rw_mbar_init(bar0, 1); rw_mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int row, int buf) {                                 // thread 0 only
        const int b = row / Nq, q = row % Nq;
        const uint32_t bar = bar0 + 8 * buf;
        // the generic-proxy writes (cached logits) of the row that used this buffer before are ordered before the async-proxy refill
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (S16) {
            rw_mbar_expect_tx(bar, (uint32_t)(H * ldS * 2));
            for (int h = 0; h < H; ++h)
                rw_bulk_g2s(rw_smem_u32(S16buf + ((size_t)buf * H + h) * pH), Sh + ((long long)b * H * Nq + q) * ldS + h * hS, (uint32_t)(ldS * 2), bar);
        } else {
            rw_mbar_expect_tx(bar, (uint32_t)(H * ldS * 4));
            for (int h = 0; h < H; ++h)
                rw_bulk_g2s(rw_smem_u32(Sbuf + ((size_t)buf * H + h) * pS), S + ((long long)b * H * Nq + q) * ldS + h * hS, (uint32_t)(ldS * 4), bar);
        }
    };
    const int nst = (ldA + 31) / 32;                                   // steps of 32 keys; both sweeps use the same deal
    const int s0 = warp * nst / NW, s1 = (warp + 1) * nst / NW;
    if (tid == 0 && blockIdx.x < rows_total) issue(blockIdx.x, 0);
    int it = 0;
    for (int row = blockIdx.x; row < rows_total; row += gridDim.x, ++it) {
        const int buf = it & 1;
        // buffer buf^1 was last touched in the previous iteration, which ended with __syncthreads
        if (tid == 0 && row + (int)gridDim.x < rows_total) issue(row + gridDim.x, buf ^ 1);
        rw_mbar_wait(bar0 + 8 * buf, (uint32_t)(it >> 1) & 1u);
        float* Sb = S16 ? Sbuf : Sbuf + (size_t)buf * H * pS;
        const uint16_t* S16b = S16buf + (size_t)buf * H * pH;
        const int b = row / Nq, q = row % Nq;
        uint16_t* Ab = A + ((long long)b * H * Nq + q) * ldA;
        // ---- sweep A (this warp's steps): mixed logits -> smem, online (max, sum)
        float m = -INFINITY, z = 0.f;
        for (int st = s0; st < s1; ++st) {
            const int base = st * 32;
            if (S16) {
                if (base + 32 <= Nk) tfwd_step_a16<H, false>(S16b, pH, Sb, pS, ldS, base, Nk, fWl, bl2, q4, r4, m, z);
                else tfwd_step_a16<H, true>(S16b, pH, Sb, pS, ldS, base, Nk, fWl, bl2, q4, r4, m, z);
            } else {
                if (base + 32 <= Nk) tfwd_step_a<H, false>(Sb, pS, ldS, base, Nk, fWl, bl2, q4, r4, m, z);
                else tfwd_step_a<H, true>(Sb, pS, ldS, base, Nk, fWl, bl2, q4, r4, m, z);
            }
        }
        {
            float M = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 2));
            float zz = (m == -INFINITY) ? 0.f : z * fast_ex2(m - M);
            zz += __shfl_xor_sync(0xffffffffu, zz, 1);
            zz += __shfl_xor_sync(0xffffffffu, zz, 2);
            if (r4 == 0) { red[warp][q4][0] = M; red[warp][q4][1] = zz; }
        }
        __syncthreads();
        float m2 = -INFINITY, zt = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < NW; ++w2) m2 = fmaxf(m2, red[w2][q4][0]);
#pragma unroll
        for (int w2 = 0; w2 < NW; ++w2) { const float mw = red[w2][q4][0]; zt += (mw == -INFINITY) ? 0.f : red[w2][q4][1] * fast_ex2(mw - m2); }
        const float c2 = m2 + log2f(zt);                                // p = 2^(L2 - m2) / zt = 2^(L2 - c2)
        if (stats && warp == 0 && r4 == 0 && q4 < H) stats[(long long)row * H + q4] = c2;
        // ---- sweep B: P -> second mix -> bf16
        for (int st = s0; st < s1; ++st) {
            const int base = st * 32;
            if (base + 32 <= Nk) tfwd_step_b<H, false>(Sb, pS, ldS, base, Nk, fWw, bwv, c2, Ab, hA, ldA, q4, r4);
            else tfwd_step_b<H, true>(Sb, pS, ldS, base, Nk, fWw, bwv, c2, Ab, hA, ldA, q4, r4);
        }
        __syncthreads();                                               // row buffer and `red` are free again
    }
}

// ---- backward step bodies.  Sweep B leaves everything sweep C needs in shared memory, in place:
//   dP (fp32, accumulator layout)       over the step's S values            (same H x 32 footprint),
//   P  (bf16, the lane's 8 keys)        over the lane's own 16-byte dA slot,
//   S_hi (bf16, B-operand layout)       in the lane-private side buffer X   (16 bytes per lane and step).
template <int H, bool TAIL>
__device__ __forceinline__ void tbwd_step_b(float* Sb, uint16_t* Db, uint4* Xst, int pS, int pA, int ldS, int ldA, int base, int Nk, uint32_t fWl, const MixFrag& fWwT,
                                            float bl2, float c2, int q4, int r4, int lane, float& rho, float (&accWw)[4]) {
    float L2[8], d[8], dP[8], p[8];
    uint32_t shi[4], dpk[4], ppk[4];
    step_logits16_shi<H, TAIL>(load_sraw<H>(Sb, pS, base, ldS, q4, r4), base, Nk, fWl, bl2, q4, r4, L2, shi);
    const int c8 = base + 8 * r4;
    const bool slot = q4 < H && c8 + 8 <= ldA;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (slot) v = *reinterpret_cast<const uint4*>(Db + q4 * pA + c8);
    {
        float2 t;
        t = unpack_bf16x2(v.x); d[0] = t.x; d[1] = t.y;
        t = unpack_bf16x2(v.y); d[2] = t.x; d[3] = t.y;
        t = unpack_bf16x2(v.z); d[4] = t.x; d[5] = t.y;
        t = unpack_bf16x2(v.w); d[6] = t.x; d[7] = t.y;
    }
    if (TAIL) {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (c8 + i >= Nk) d[i] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        p[t] = fast_ex2(L2[t] - c2); p[4 + t] = fast_ex2(L2[4 + t] - c2);
        dpk[t] = pack_bf16x2(d[t], d[4 + t]);
        ppk[t] = pack_bf16x2(p[t], p[4 + t]);
        mix_tile(fWwT, movm_trans(dpk[t]), 0u, 0.f, dP[t], dP[4 + t]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) rho += p[i] * dP[i];
    mma16816(accWw, dpk[0], 0u, dpk[1], 0u, ppk[0], ppk[1]);
    mma16816(accWw, dpk[2], 0u, dpk[3], 0u, ppk[2], ppk[3]);
    __syncwarp();                                                    // all lanes have consumed this step's S before it is overwritten
    if (slot) {
        *reinterpret_cast<uint4*>(Db + q4 * pA + c8) = make_uint4(ppk[0], ppk[1], ppk[2], ppk[3]);
        if (c8 + 8 <= ldS) {
            float4* dst = reinterpret_cast<float4*>(Sb + q4 * pS + c8);
            dst[0] = make_float4(dP[0], dP[1], dP[2], dP[3]);
            dst[1] = make_float4(dP[4], dP[5], dP[6], dP[7]);
        }
    }
    Xst[lane] = make_uint4(shi[0], shi[1], shi[2], shi[3]);
}
template <int H, bool TAIL>
__device__ __forceinline__ void tbwd_step_c(const float* Sb, const uint16_t* Db, const uint4* Xst, int pS, int pA, int ldS, int ldA, int base, int Nk, const MixFrag& fWlT,
                                            float rho, uint16_t* dSb, long long hA, int q4, int r4, int lane, float (&accWl)[4]) {
    const int c8 = base + 8 * r4;
    const bool slot = q4 < H && c8 + 8 <= ldA && c8 + 8 <= ldS;
    float l[8], o[8];
    uint32_t lpk[4];
    if (slot) {
        const uint4 pp = *reinterpret_cast<const uint4*>(Db + q4 * pA + c8);
        const float4 u = *reinterpret_cast<const float4*>(Sb + q4 * pS + c8), w = *reinterpret_cast<const float4*>(Sb + q4 * pS + c8 + 4);
        const float2 p0 = unpack_bf16x2(pp.x), p1 = unpack_bf16x2(pp.y), p2 = unpack_bf16x2(pp.z), p3 = unpack_bf16x2(pp.w);   // (p[t], p[4+t])
        l[0] = p0.x * (u.x - rho); l[4] = p0.y * (w.x - rho);
        l[1] = p1.x * (u.y - rho); l[5] = p1.y * (w.y - rho);
        l[2] = p2.x * (u.z - rho); l[6] = p2.y * (w.z - rho);
        l[3] = p3.x * (u.w - rho); l[7] = p3.y * (w.w - rho);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) l[i] = 0.f;
    }
    const uint4 sx = Xst[lane];
    const uint32_t shi[4] = {sx.x, sx.y, sx.z, sx.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        lpk[t] = pack_bf16x2(l[t], l[4 + t]);
        mix_tile(fWlT, movm_trans(lpk[t]), 0u, 0.f, o[t], o[4 + t]);
    }
    if (!TAIL || base < Nk) {
        mma16816(accWl, lpk[0], 0u, lpk[1], 0u, movm_trans(shi[0]), movm_trans(shi[1]));
        mma16816(accWl, lpk[2], 0u, lpk[3], 0u, movm_trans(shi[2]), movm_trans(shi[3]));
    }
    if (q4 < H && c8 + 8 <= ldA) {
        if (TAIL) {
#pragma unroll
            for (int i = 0; i < 8; ++i) if (c8 + i >= Nk) o[i] = 0.f;
        }
        *reinterpret_cast<uint4*>(dSb + q4 * hA + c8) =
            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
}

// fp16-S flavours: the raw logits stay in their own fp16 row buffer (never overwritten), so sweep C re-reads them for the dWl outer
// product and the lane-private S_hi side buffer is not needed; dP goes to the fp32 buffer as before.
template <int H, bool TAIL>
__device__ __forceinline__ void tbwd_step_b16(const uint16_t* S16b, int pH, float* Sb, uint16_t* Db, int pS, int pA, int ldS, int ldA, int base, int Nk, uint32_t fWl,
                                              const MixFrag& fWwT, float bl2, float c2, int q4, int r4, float& rho, float (&accWw)[4]) {
    float L2[8], d[8], dP[8], p[8];
    uint32_t dpk[4], ppk[4];
    {
        const SRaw16 raw = load_sraw16<H>(S16b, pH, base, ldS, q4, r4);
        const int cb = base + 4 * q4;
        const bool tail = TAIL && base + 32 > Nk;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const bool ok = !tail || cb + t < Nk;
            mix_tile16(fWl, ok ? sraw16_pair(raw, t) : 0u, bl2, L2[t], L2[4 + t]);
        }
        if (tail) {
            const int c8t = base + 8 * r4;
#pragma unroll
            for (int i = 0; i < 8; ++i) if (c8t + i >= Nk) L2[i] = -INFINITY;
        }
    }
    const int c8 = base + 8 * r4;
    const bool slot = q4 < H && c8 + 8 <= ldA;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (slot) v = *reinterpret_cast<const uint4*>(Db + q4 * pA + c8);
    {
        float2 t;
        t = unpack_bf16x2(v.x); d[0] = t.x; d[1] = t.y;
        t = unpack_bf16x2(v.y); d[2] = t.x; d[3] = t.y;
        t = unpack_bf16x2(v.z); d[4] = t.x; d[5] = t.y;
        t = unpack_bf16x2(v.w); d[6] = t.x; d[7] = t.y;
    }
    if (TAIL) {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (c8 + i >= Nk) d[i] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        p[t] = fast_ex2(L2[t] - c2); p[4 + t] = fast_ex2(L2[4 + t] - c2);
        dpk[t] = pack_bf16x2(d[t], d[4 + t]);
        ppk[t] = pack_bf16x2(p[t], p[4 + t]);
        mix_tile(fWwT, movm_trans(dpk[t]), 0u, 0.f, dP[t], dP[4 + t]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) rho += p[i] * dP[i];
    mma16816(accWw, dpk[0], 0u, dpk[1], 0u, ppk[0], ppk[1]);
    mma16816(accWw, dpk[2], 0u, dpk[3], 0u, ppk[2], ppk[3]);
    if (slot) {
        *reinterpret_cast<uint4*>(Db + q4 * pA + c8) = make_uint4(ppk[0], ppk[1], ppk[2], ppk[3]);
        if (c8 + 8 <= ldS) {
            float4* dst = reinterpret_cast<float4*>(Sb + q4 * pS + c8);
            dst[0] = make_float4(dP[0], dP[1], dP[2], dP[3]);
            dst[1] = make_float4(dP[4], dP[5], dP[6], dP[7]);
        }
    }
}
template <int H, bool TAIL>
__device__ __forceinline__ void tbwd_step_c16(const uint16_t* S16b, int pH, const float* Sb, const uint16_t* Db, int pS, int pA, int ldS, int ldA, int base, int Nk,
                                              const MixFrag& fWlT, float rho, uint16_t* dSb, long long hA, int q4, int r4, float (&accWl)[4]) {
    const int c8 = base + 8 * r4;
    const bool slot = q4 < H && c8 + 8 <= ldA && c8 + 8 <= ldS;
    float l[8], o[8];
    uint32_t lpk[4], shi[4];
    if (slot) {
        const uint4 pp = *reinterpret_cast<const uint4*>(Db + q4 * pA + c8);
        const float4 u = *reinterpret_cast<const float4*>(Sb + q4 * pS + c8), w = *reinterpret_cast<const float4*>(Sb + q4 * pS + c8 + 4);
        const float2 p0 = unpack_bf16x2(pp.x), p1 = unpack_bf16x2(pp.y), p2 = unpack_bf16x2(pp.z), p3 = unpack_bf16x2(pp.w);   // (p[t], p[4+t])
        l[0] = p0.x * (u.x - rho); l[4] = p0.y * (w.x - rho);
        l[1] = p1.x * (u.y - rho); l[5] = p1.y * (w.y - rho);
        l[2] = p2.x * (u.z - rho); l[6] = p2.y * (w.z - rho);
        l[3] = p3.x * (u.w - rho); l[7] = p3.y * (w.w - rho);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) l[i] = 0.f;
    }
    {
        // S of the step in the B-operand layout, bf16 (heads 2 r4, 2 r4 + 1 of keys base + 4 q4 + t)
        const SRaw16 raw = load_sraw16<H>(S16b, pH, base, ldS, q4, r4);
        const int cb = base + 4 * q4;
        const bool tail = TAIL && base + 32 > Nk;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const bool ok = !tail || cb + t < Nk;
            const uint32_t pr = ok ? sraw16_pair(raw, t) : 0u;
            shi[t] = pack_bf16x2(f16_bits_to_f(pr & 0xffffu), f16_bits_to_f(pr >> 16));
        }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        lpk[t] = pack_bf16x2(l[t], l[4 + t]);
        mix_tile(fWlT, movm_trans(lpk[t]), 0u, 0.f, o[t], o[4 + t]);
    }
    if (!TAIL || base < Nk) {
        mma16816(accWl, lpk[0], 0u, lpk[1], 0u, movm_trans(shi[0]), movm_trans(shi[1]));
        mma16816(accWl, lpk[2], 0u, lpk[3], 0u, movm_trans(shi[2]), movm_trans(shi[3]));
    }
    if (q4 < H && c8 + 8 <= ldA) {
        if (TAIL) {
#pragma unroll
            for (int i = 0; i < 8; ++i) if (c8 + i >= Nk) o[i] = 0.f;
        }
        *reinterpret_cast<uint4*>(dSb + q4 * hA + c8) =
            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
}

template <int H, int NW, bool S16>
__global__ void __launch_bounds__(NW * 32, 2) talking_bwd_rows_kernel(const void* __restrict__ Sv, const uint16_t* dA, uint16_t* dS, const float* __restrict__ Wl,
                                                                      const float* __restrict__ bl, const float* __restrict__ Ww,
                                                                      const float* __restrict__ stats, int rows_total, int Nq, int Nk, int ldS, int ldA,
                                                                      float* __restrict__ part) {
    extern __shared__ __align__(128) uint8_t rsm[];
    // single row buffer per CTA; TWO CTAs share an SM, so one CTA's row load overlaps the other's math
    const int pS = talking_pitch_f32(ldS), pA = talking_pitch_bf16(ldA), pH = talking_pitch_f16(ldS);   // padded pitches: no bank conflicts
    float* Sbuf = reinterpret_cast<float*>(rsm);                                            // [H][pS]   S, then dP   (fp16 S: dP only)
    uint16_t* Dbuf = reinterpret_cast<uint16_t*>(rsm + (size_t)H * pS * 4);                 // [H][pA]   dA, then P
    uint4* Xbuf = reinterpret_cast<uint4*>(rsm + (size_t)H * pS * 4 + (size_t)H * pA * 2);  // [steps][32]  S_hi fragments (fp32 S)
    uint16_t* S16buf = reinterpret_cast<uint16_t*>(Xbuf);                                   // [H][pH]   fp16 S (instead of Xbuf)
    const float* S = reinterpret_cast<const float*>(Sv);
    const uint16_t* Sh = reinterpret_cast<const uint16_t*>(Sv);
    constexpr int NP = 2 * H * H + 2 * H;
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ float red[NW][8];
    __shared__ float redbuf[NW][NP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q4 = lane >> 2, r4 = lane & 3;
    const uint32_t fWl = load_mix_frag16<H>(Wl, LOG2E, q4, r4);          // logits: the forward's fp16 single pass
    const MixFrag fWwT = load_mix_frag<H>(Ww, true, 1.f, q4, r4);
    const MixFrag fWlT = load_mix_frag<H>(Wl, true, 1.f, q4, r4);
    const float bl2 = q4 < H ? bl[q4] * LOG2E : 0.f;
    const long long hS = (long long)Nq * ldS, hA = (long long)Nq * ldA;
    const uint32_t bar0 = rw_smem_u32(bars);
    if (tid == 0) {
        rw_mbar_init(bar0, 1); rw_mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int row) {
        const int b = row / Nq, q = row % Nq;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes of the previous row before the async refill
        rw_mbar_expect_tx(bar0, (uint32_t)(H * ldS * (S16 ? 2 : 4) + H * ldA * 2));
        for (int h = 0; h < H; ++h) {
            if (S16) rw_bulk_g2s(rw_smem_u32(S16buf + (size_t)h * pH), Sh + ((long long)b * H * Nq + q) * ldS + h * hS, (uint32_t)(ldS * 2), bar0);
            else rw_bulk_g2s(rw_smem_u32(Sbuf + (size_t)h * pS), S + ((long long)b * H * Nq + q) * ldS + h * hS, (uint32_t)(ldS * 4), bar0);
            rw_bulk_g2s(rw_smem_u32(Dbuf + (size_t)h * pA), dA + ((long long)b * H * Nq + q) * ldA + h * hA, (uint32_t)(ldA * 2), bar0);
        }
    };
    const int nst = (ldA + 31) / 32;
    const int s0 = warp * nst / NW, s1 = (warp + 1) * nst / NW;

    float accWw[4] = {0.f, 0.f, 0.f, 0.f}, accWl[4] = {0.f, 0.f, 0.f, 0.f};
    if (tid == 0 && blockIdx.x < rows_total) issue(blockIdx.x);
    int it = 0;
    float c2n = (q4 < H && blockIdx.x < rows_total) ? stats[(long long)blockIdx.x * H + q4] : 0.f;
    for (int row = blockIdx.x; row < rows_total; row += gridDim.x, ++it) {
        const float c2 = c2n;
        if (q4 < H && row + (int)gridDim.x < rows_total) c2n = stats[(long long)(row + gridDim.x) * H + q4];     // consumed one row later
        rw_mbar_wait(bar0, (uint32_t)it & 1u);
        const int b = row / Nq, q = row % Nq;
        uint16_t* dSb = dS + ((long long)b * H * Nq + q) * ldA;
        // ---- sweep B: rho = sum_j P dP;  dWw += dA (x) P;  leaves dP / P / S_hi in shared memory
        float rho = 0.f;
        for (int st = s0; st < s1; ++st) {
            const int base = st * 32;
            if (S16) {
                if (base + 32 <= Nk) tbwd_step_b16<H, false>(S16buf, pH, Sbuf, Dbuf, pS, pA, ldS, ldA, base, Nk, fWl, fWwT, bl2, c2, q4, r4, rho, accWw);
                else tbwd_step_b16<H, true>(S16buf, pH, Sbuf, Dbuf, pS, pA, ldS, ldA, base, Nk, fWl, fWwT, bl2, c2, q4, r4, rho, accWw);
            } else if (base + 32 <= Nk) tbwd_step_b<H, false>(Sbuf, Dbuf, Xbuf + (size_t)st * 32, pS, pA, ldS, ldA, base, Nk, fWl, fWwT, bl2, c2, q4, r4, lane, rho, accWw);
            else tbwd_step_b<H, true>(Sbuf, Dbuf, Xbuf + (size_t)st * 32, pS, pA, ldS, ldA, base, Nk, fWl, fWwT, bl2, c2, q4, r4, lane, rho, accWw);
        }
        rho += __shfl_xor_sync(0xffffffffu, rho, 1);
        rho += __shfl_xor_sync(0xffffffffu, rho, 2);
        if (r4 == 0) red[warp][q4] = rho;
        __syncthreads();
        rho = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < NW; ++w2) rho += red[w2][q4];
        // ---- sweep C: dL = P (dP - rho);  dS = Wl^T dL;  dWl += dL (x) S     (dbl = sum dL is identically 0: not accumulated)
        for (int st = s0; st < s1; ++st) {
            const int base = st * 32;
            if (S16) {
                if (base + 32 <= Nk) tbwd_step_c16<H, false>(S16buf, pH, Sbuf, Dbuf, pS, pA, ldS, ldA, base, Nk, fWlT, rho, dSb, hA, q4, r4, accWl);
                else tbwd_step_c16<H, true>(S16buf, pH, Sbuf, Dbuf, pS, pA, ldS, ldA, base, Nk, fWlT, rho, dSb, hA, q4, r4, accWl);
            } else if (base + 32 <= Nk) tbwd_step_c<H, false>(Sbuf, Dbuf, Xbuf + (size_t)st * 32, pS, pA, ldS, ldA, base, Nk, fWlT, rho, dSb, hA, q4, r4, lane, accWl);
            else tbwd_step_c<H, true>(Sbuf, Dbuf, Xbuf + (size_t)st * 32, pS, pA, ldS, ldA, base, Nk, fWlT, rho, dSb, hA, q4, r4, lane, accWl);
        }
        __syncthreads();                                               // every warp is done with the row buffer
        if (tid == 0 && row + (int)gridDim.x < rows_total) issue(row + gridDim.x);
    }
    // CTA reduction of the partials -> part[blockIdx.x][...]: layout dWl[H*H], dbl[H] (= 0), dWw[H*H], dbw[H] (= 0: see ops.py, exact formula)
    for (int i = lane; i < NP; i += 32) redbuf[warp][i] = 0.f;
    __syncwarp();
    if (q4 < H) {
        const int col = 2 * r4;
        if (col < H) { redbuf[warp][q4 * H + col] = accWl[0]; redbuf[warp][H * H + H + q4 * H + col] = accWw[0]; }
        if (col + 1 < H) { redbuf[warp][q4 * H + col + 1] = accWl[1]; redbuf[warp][H * H + H + q4 * H + col + 1] = accWw[1]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NP; i += NW * 32) {
        float sum = 0.f;
        for (int w2 = 0; w2 < NW; ++w2) sum += redbuf[w2][i];
        part[(long long)blockIdx.x * NP + i] = sum;
    }
}

// one warp per output element: lanes stride over the per-CTA partials
__global__ void __launch_bounds__(256) talking_bwd_finalize_kernel(const float* __restrict__ part, int nblocks, int H, float* __restrict__ dWl,
                                                                   float* __restrict__ dbl, float* __restrict__ dWw, float* __restrict__ dbw) {
    const int NP = 2 * H * H + 2 * H;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= NP) return;
    float s = 0.f;
    for (int k = lane; k < nblocks; k += 32) s += part[(long long)k * NP + i];
    s = warp_sum(s);
    if (lane != 0) return;
    if (i < H * H) dWl[i] += s;
    else if (i < H * H + H) dbl[i - H * H] += s;
    else if (i < 2 * H * H + H) dWw[i - H * H - H] += s;
    else dbw[i - 2 * H * H - H] += s;
}

// ------------------------------------------------------------------------------------------------
// LayerScale branch backward + column sums
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layerscale_bwd_kernel(const float* __restrict__ dout, const uint16_t* __restrict__ y, const float* __restrict__ gamma,
                                                             long long rows, int D, uint16_t* __restrict__ dy, float* __restrict__ dgamma,
                                                             float* __restrict__ dbias) {
    // block (32 x 8): x -> 4-column groups, y -> rows
    __shared__ float4 rg[8][32], rb[8][32];
    const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
    const bool ok = c < D;
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;
    float4 gm = ag;
    if (ok) gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
    for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < rows && ok; r += (long long)gridDim.y * 8) {
        const float4 d = *reinterpret_cast<const float4*>(dout + r * D + c);
        const uint2 yp = *reinterpret_cast<const uint2*>(y + r * D + c);
        const float2 y0 = unpack_bf16x2(yp.x), y1 = unpack_bf16x2(yp.y);
        const float4 o = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        *reinterpret_cast<uint2*>(dy + r * D + c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        ag.x += d.x * y0.x; ag.y += d.y * y0.y; ag.z += d.z * y1.x; ag.w += d.w * y1.y;
        ab.x += o.x; ab.y += o.y; ab.z += o.z; ab.w += o.w;
    }
    rg[threadIdx.y][threadIdx.x] = ag; rb[threadIdx.y][threadIdx.x] = ab;
    __syncthreads();
    if (threadIdx.y == 0 && ok) {
        for (int k = 1; k < 8; ++k) {
            const float4 a = rg[k][threadIdx.x], b2 = rb[k][threadIdx.x];
            ag.x += a.x; ag.y += a.y; ag.z += a.z; ag.w += a.w;
            ab.x += b2.x; ab.y += b2.y; ab.z += b2.z; ab.w += b2.w;
        }
        if (dgamma) { atomicAdd(dgamma + c, ag.x); atomicAdd(dgamma + c + 1, ag.y); atomicAdd(dgamma + c + 2, ag.z); atomicAdd(dgamma + c + 3, ag.w); }
        if (dbias) { atomicAdd(dbias + c, ab.x); atomicAdd(dbias + c + 1, ab.y); atomicAdd(dbias + c + 2, ab.z); atomicAdd(dbias + c + 3, ab.w); }
    }
}

__global__ void __launch_bounds__(256) colsum_bf16_kernel(const uint16_t* __restrict__ x, long long rows, int N, long long ld, float* __restrict__ out,
                                                          long long x_bstride, long long out_bstride) {
    __shared__ float2 red[8][32];
    x += (long long)blockIdx.z * x_bstride;
    out += (long long)blockIdx.z * out_bstride;
    const int c = (blockIdx.x * 32 + threadIdx.x) * 2;
    float2 acc = make_float2(0.f, 0.f);
    if (c + 1 < N) {
        for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8) {
            const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + r * ld + c));
            acc.x += v.x; acc.y += v.y;
        }
    } else if (c < N) {
        for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8) acc.x += bf16_to_f(x[r * ld + c]);
    }
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        for (int k = 1; k < 8; ++k) { acc.x += red[k][threadIdx.x].x; acc.y += red[k][threadIdx.x].y; }
        atomicAdd(out + c, acc.x);
        if (c + 1 < N) atomicAdd(out + c + 1, acc.y);
    }
}

__global__ void axpby_cast_kernel(const float* __restrict__ x, const float* __restrict__ y, float a, float b, long long n, uint16_t* __restrict__ o16,
                                  float* __restrict__ o32) {
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        if (i + 3 < n) {
            float4 v = *reinterpret_cast<const float4*>(x + i);
            v.x *= a; v.y *= a; v.z *= a; v.w *= a;
            if (y) { const float4 u = *reinterpret_cast<const float4*>(y + i); v.x += b * u.x; v.y += b * u.y; v.z += b * u.z; v.w += b * u.w; }
            if (o32) *reinterpret_cast<float4*>(o32 + i) = v;
            if (o16) *reinterpret_cast<uint2*>(o16 + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        } else {
            for (long long k = i; k < n; ++k) {
                float v = a * x[k] + (y ? b * y[k] : 0.f);
                if (o32) o32[k] = v;
                if (o16) o16[k] = f_to_bf16(v);
            }
        }
    }
}

__global__ void cast_bf16_f32_kernel(const uint16_t* __restrict__ x, float* __restrict__ y, long long n, int accumulate) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = bf16_to_f(x[i]);
        y[i] = accumulate ? y[i] + v : v;
    }
}

// ------------------------------------------------------------------------------------------------
// patch im2col: img f32 [B,3,H,W] -> bf16 [B*h*w, 3*p*p], 4 pixels per thread
// ------------------------------------------------------------------------------------------------
__global__ void im2col_patch_kernel(const float* __restrict__ img, int B, int Himg, int Wimg, int p, uint16_t* __restrict__ out) {
    const int h = Himg / p, w = Wimg / p, Kc = 3 * p * p;
    const long long total = (long long)B * h * w * Kc / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 4;
        const int k = (int)(e % Kc);
        const long long tok = e / Kc;
        const int px = (int)(tok % w), py = (int)((tok / w) % h), b = (int)(tok / ((long long)w * h));
        const int c = k / (p * p), ky = (k / p) % p, kx = k % p;
        const float* src = img + (((long long)b * 3 + c) * Himg + (py * p + ky)) * Wimg + px * p + kx;
        float4 v;
        if ((Wimg & 3) == 0) v = *reinterpret_cast<const float4*>(src);
        else v = make_float4(src[0], src[1], src[2], src[3]);          // image rows not 16-byte aligned (e.g. W = 1333, COCO shape): scalar loads
        *reinterpret_cast<uint2*>(out + e) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}

// ------------------------------------------------------------------------------------------------
// bicubic (A = -0.75, align_corners = False), token-major
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_coeffs(float t, float* c) {
    const float A = -0.75f;
    float x = t + 1.f;  c[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
    x = t;              c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    x = 1.f - t;        c[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    x = 2.f - t;        c[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

template <bool BWD>
__global__ void bicubic_tokens_kernel(const float* __restrict__ a, float* __restrict__ o, int sh, int sw, int dh, int dw, int D) {
    // FWD: a = src [sh*sw,D], o = dst [dh*dw,D].   BWD: a = ddst, o = dsrc (atomic scatter)
    const long long total = (long long)dh * dw * D;
    const float scy = (float)sh / (float)dh, scx = (float)sw / (float)dw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % D);
        const int ox = (int)((i / D) % dw), oy = (int)(i / ((long long)D * dw));
        const float ry = scy * (oy + 0.5f) - 0.5f, rx = scx * (ox + 0.5f) - 0.5f;
        const int iy = (int)floorf(ry), ix = (int)floorf(rx);
        float cy[4], cx[4];
        cubic_coeffs(ry - iy, cy);
        cubic_coeffs(rx - ix, cx);
        float acc = 0.f;
        const float gin = BWD ? a[i] : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int yy = min(max(iy - 1 + u, 0), sh - 1);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int xx = min(max(ix - 1 + v, 0), sw - 1);
                if (BWD) atomicAdd(o + ((long long)yy * sw + xx) * D + c, gin * cy[u] * cx[v]);
                else acc += a[((long long)yy * sw + xx) * D + c] * cy[u] * cx[v];
            }
        }
        if (!BWD) o[i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// sine position encodings
// ------------------------------------------------------------------------------------------------
__global__ void sine_pos_2d_kernel(const uint8_t* __restrict__ mask, int B, int h, int w, int D, float* __restrict__ pos, uint16_t* __restrict__ pos16) {
    const int npf = D / 2;
    const long long total = (long long)B * h * w * npf;
    const float two_pi = 6.283185307179586f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % npf);
        const int x = (int)((i / npf) % w), y = (int)((i / ((long long)npf * w)) % h), b = (int)(i / ((long long)npf * w * h));
        const uint8_t* m = mask + (long long)b * h * w;
        float ye = 0.f, yt = 0.f, xe = 0.f, xt = 0.f;
        for (int yy = 0; yy < h; ++yy) { const float nm = m[yy * w + x] ? 0.f : 1.f; yt += nm; if (yy <= y) ye += nm; }
        for (int xx = 0; xx < w; ++xx) { const float nm = m[y * w + xx] ? 0.f : 1.f; xt += nm; if (xx <= x) xe += nm; }
        ye = ye / (yt + 1e-6f) * two_pi;
        xe = xe / (xt + 1e-6f) * two_pi;
        const float dim_t = powf(10000.f, (float)(2 * (f / 2)) / (float)npf);
        const float ay = ye / dim_t, ax = xe / dim_t;
        const float vy = (f & 1) ? cosf(ay) : sinf(ay);
        const float vx = (f & 1) ? cosf(ax) : sinf(ax);
        const long long tok = ((long long)b * h + y) * w + x;
        if (pos) { pos[tok * D + f] = vy; pos[tok * D + npf + f] = vx; }
        if (pos16) { pos16[tok * D + f] = f_to_bf16(vy); pos16[tok * D + npf + f] = f_to_bf16(vx); }
    }
}

template <bool BWD>
__global__ void query_sine_kernel(const float* __restrict__ ref, const float* __restrict__ demb, long long n, int D, float* __restrict__ out) {
    // FWD: out = emb [n,D].  BWD: out = dref [n,2], one warp per row
    const int half = D / 2;
    const float two_pi = 6.283185307179586f;
    if (!BWD) {
        const long long total = n * half;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int f = (int)(i % half);
            const long long r = i / half;
            const float dim_t = powf(10000.f, (float)(2 * (f / 2)) / 128.f);
            const float ax = ref[r * 2 + 0] * two_pi / dim_t, ay = ref[r * 2 + 1] * two_pi / dim_t;
            out[r * D + f] = (f & 1) ? cosf(ay) : sinf(ay);
            out[r * D + half + f] = (f & 1) ? cosf(ax) : sinf(ax);
        }
    } else {
        const int lane = threadIdx.x & 31;
        const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        if (r >= n) return;
        const float rx = ref[r * 2 + 0], ry = ref[r * 2 + 1];
        float gx = 0.f, gy = 0.f;
        for (int f = lane; f < half; f += 32) {
            const float k = two_pi / powf(10000.f, (float)(2 * (f / 2)) / 128.f);
            const float ay = ry * k, ax = rx * k;
            gy += demb[r * D + f] * ((f & 1) ? -sinf(ay) : cosf(ay)) * k;
            gx += demb[r * D + half + f] * ((f & 1) ? -sinf(ax) : cosf(ax)) * k;
        }
        gx = warp_sum(gx); gy = warp_sum(gy);
        if (lane == 0) { out[r * 2 + 0] = gx; out[r * 2 + 1] = gy; }
    }
}

// dpre = dout * (h > 0)   (bf16 x bf16 -> bf16), for a ReLU fused into the producing GEMM's epilogue
__global__ void relu_bwd_kernel(const uint16_t* __restrict__ dout, const uint16_t* __restrict__ h, uint16_t* __restrict__ out, long long n) {
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (long long)gridDim.x * blockDim.x * 2) {
        if (i + 1 < n) {
            const uint32_t d = *reinterpret_cast<const uint32_t*>(dout + i), hv = *reinterpret_cast<const uint32_t*>(h + i);
            const float2 hf = unpack_bf16x2(hv);
            uint32_t o = d;
            if (!(hf.x > 0.f)) o &= 0xffff0000u;
            if (!(hf.y > 0.f)) o &= 0x0000ffffu;
            *reinterpret_cast<uint32_t*>(out + i) = o;
        } else {
            out[i] = bf16_to_f(h[i]) > 0.f ? dout[i] : (uint16_t)0;
        }
    }
}

inline int grid_for(long long work_items, int threads) {
    long long g = (work_items + threads - 1) / threads;
    const long long cap = (long long)spe_num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" __attribute__((visibility("default"))) int spe_layernorm_fwd(const float* x, const float* w, const float* b, float eps, int64_t rows, int D, void* y_bf16, float* y_f32,
                                 float* mean, float* rstd, void* stream) {
    SPE_CHECK(x && w && b && rows > 0, "spe_layernorm_fwd: bad argument");
    SPE_CHECK(D % 4 == 0 && D <= 128 * LN_MAXCH, "spe_layernorm_fwd: D=%d must be a multiple of 4 and <= %d", D, 128 * LN_MAXCH);
    const int g = (int)((rows + 7) / 8 < (long long)spe_num_sms() * 8 ? (rows + 7) / 8 : (long long)spe_num_sms() * 8);
    SpeProfScope prof(SPE_FAM_LAYERNORM, (double)rows * D * (4.0 + (y_bf16 ? 2.0 : 0.0) + (y_f32 ? 4.0 : 0.0)), ST(stream));
    layernorm_fwd_kernel<<<g, 256, 0, ST(stream)>>>(x, w, b, eps, rows, D, reinterpret_cast<uint16_t*>(y_bf16), y_f32, mean, rstd);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* dres, const float* x, const float* w, const float* mean,
                                 const float* rstd, int64_t rows, int D, float* dx, float* dw, float* db, void* stream) {
    SPE_CHECK((dy_bf16 || dy_f32) && x && w && mean && rstd && dx && rows > 0, "spe_layernorm_bwd: bad argument");
    SPE_CHECK(D % 4 == 0 && D <= 128 * LN_MAXCH, "spe_layernorm_bwd: unsupported D=%d", D);
    const int nch = (D + 127) / 128;
    const int per_sm = nch <= 2 ? 4 : (nch == 3 ? 3 : (nch <= 6 ? 2 : 1));
    long long g = (rows + 7) / 8;
    if (g > (long long)spe_num_sms() * per_sm) g = (long long)spe_num_sms() * per_sm;
    SpeProfScope prof(SPE_FAM_LAYERNORM, (double)rows * D * (8.0 + (dy_bf16 ? 2.0 : 0.0) + (dy_f32 ? 4.0 : 0.0)), ST(stream));
#define SPE_LN_BWD(NCH_) layernorm_bwd_kernel<NCH_><<<(int)g, 256, 2 * D * sizeof(float), ST(stream)>>>(reinterpret_cast<const uint16_t*>(dy_bf16), dy_f32, dres, x, w, \
                                                                                                      mean, rstd, rows, D, dx, dw, db)
    switch (nch) {
        case 1: SPE_LN_BWD(1); break;
        case 2: SPE_LN_BWD(2); break;
        case 3: SPE_LN_BWD(3); break;
        case 4: SPE_LN_BWD(4); break;
        case 5: case 6: SPE_LN_BWD(6); break;
        default: SPE_LN_BWD(8); break;
    }
#undef SPE_LN_BWD
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_softmax_fwd(const float* S, void* P, const uint8_t* mask, int B, int H, int Nq, int Nk, int64_t ldS, int64_t ldP, float* pmean,
                               void* stream) {
    SPE_CHECK(S && P && B > 0 && H > 0 && Nq > 0 && Nk > 0 && ldS >= Nk && ldP >= Nk, "spe_softmax_fwd: bad argument");
    SPE_CHECK((size_t)Nk * 4 <= 200 * 1024, "spe_softmax_fwd: Nk too large");
    static bool done = false;
    if (!done) { SPE_CUDA(cudaFuncSetAttribute(softmax_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); done = true; }
    if (pmean) SPE_CUDA(cudaMemsetAsync(pmean, 0, (size_t)B * Nq * Nk * 4, ST(stream)));
    SpeProfScope prof(SPE_FAM_SOFTMAX, (double)B * H * Nq * Nk * 6.0, ST(stream));
    if (ldS % 4 == 0 && ldP % 4 == 0 && Nk <= 2048 && getenv("SPE_SOFTMAX_CTA_ROWS") == nullptr) {
        const long long rows = (long long)B * H * Nq;
        long long g = (rows + 7) / 8;
        if (g > (long long)spe_num_sms() * 8) g = (long long)spe_num_sms() * 8;
        if (Nk <= 512) softmax_fwd_warp_kernel<4><<<(unsigned)g, 256, 0, ST(stream)>>>(S, reinterpret_cast<uint16_t*>(P), mask, rows, H, Nq, Nk, ldS, ldP, pmean);
        else softmax_fwd_warp_kernel<16><<<(unsigned)g, 256, 0, ST(stream)>>>(S, reinterpret_cast<uint16_t*>(P), mask, rows, H, Nq, Nk, ldS, ldP, pmean);
        SPE_LAUNCHED();
        return 0;
    }
    softmax_fwd_kernel<<<(unsigned)((long long)B * H * Nq), 256, (size_t)Nk * 4, ST(stream)>>>(S, reinterpret_cast<uint16_t*>(P), mask, H, Nq, Nk, ldS,
                                                                                                ldP, pmean);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_softmax_bwd(const void* P, const void* dP, void* dS, int B, int H, int Nq, int Nk, int64_t ldP, void* stream) {
    SPE_CHECK(P && dP && dS && B > 0 && H > 0 && Nq > 0 && Nk > 0 && ldP >= Nk, "spe_softmax_bwd: bad argument");
    SPE_CHECK((size_t)Nk * 8 <= 200 * 1024, "spe_softmax_bwd: Nk too large");
    static bool done = false;
    if (!done) { SPE_CUDA(cudaFuncSetAttribute(softmax_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); done = true; }
    SpeProfScope prof(SPE_FAM_SOFTMAX, (double)B * H * Nq * Nk * 6.0, ST(stream));
    if (ldP % 4 == 0 && Nk <= 2048 && getenv("SPE_SOFTMAX_CTA_ROWS") == nullptr) {
        const long long rows = (long long)B * H * Nq;
        long long g = (rows + 7) / 8;
        if (g > (long long)spe_num_sms() * 8) g = (long long)spe_num_sms() * 8;
        if (Nk <= 512) softmax_bwd_warp_kernel<4><<<(unsigned)g, 256, 0, ST(stream)>>>(reinterpret_cast<const uint16_t*>(P), reinterpret_cast<const uint16_t*>(dP), reinterpret_cast<uint16_t*>(dS), rows, Nk, ldP);
        else softmax_bwd_warp_kernel<16><<<(unsigned)g, 256, 0, ST(stream)>>>(reinterpret_cast<const uint16_t*>(P), reinterpret_cast<const uint16_t*>(dP), reinterpret_cast<uint16_t*>(dS), rows, Nk, ldP);
        SPE_LAUNCHED();
        return 0;
    }
    softmax_bwd_kernel<<<(unsigned)((long long)B * H * Nq), 256, (size_t)Nk * 8, ST(stream)>>>(reinterpret_cast<const uint16_t*>(P),
                                                                                                reinterpret_cast<const uint16_t*>(dP),
                                                                                                reinterpret_cast<uint16_t*>(dS), Nk, ldP);
    SPE_LAUNCHED();
    return 0;
}

// warps per row-staged CTA: 8 or 10, whichever deals the 32-key steps of a row out more evenly (ties: 8)
static int talking_warps(int ld) {
    const int nst = (ld + 31) / 32;
    const int w8 = (nst + 7) / 8 * 8 - nst, w10 = (nst + 9) / 10 * 10 - nst;
    const char* e = getenv("SPE_TALKING_WARPS");
    if (e) return atoi(e) == 10 ? 10 : 8;
    return (w10 * 8 < w8 * 10 && nst >= 10) ? 10 : 8;
}

static int talking_grid(int B, int Nq, int per_sm) {
    const long long blocks = ((long long)B * Nq + 7) / 8;
    const long long cap = (long long)spe_num_sms() * per_sm;
    return (int)(blocks < cap ? blocks : cap);
}

// shared memory of the row-staged kernels with fp16 logits
static size_t talking_fwd_smem16(int H, int ldS) { return (size_t)H * talking_pitch_f32(ldS) * 4 + (size_t)2 * H * talking_pitch_f16(ldS) * 2; }
static size_t talking_bwd_smem16(int H, int ldS, int ldA) {
    return (size_t)H * (talking_pitch_f32(ldS) * 4 + talking_pitch_bf16(ldA) * 2 + talking_pitch_f16(ldS) * 2);
}

template <int H>
static int talking_fwd_launch(const void* Sv, void* A, const float* Wl, const float* bl, const float* Ww, const float* bw, float* stats, int B, int Nq,
                              int Nk, int64_t ldS, int64_t ldA, cudaStream_t st, bool s16 = false) {
    const float* S = reinterpret_cast<const float*>(Sv);
    SpeProfScope prof(SPE_FAM_TALKING_FWD, (double)B * H * Nq * Nk * (s16 ? 4.0 : 6.0), st);   // algorithmic bytes: S read + A bf16 write
    const size_t smem = s16 ? talking_fwd_smem16(H, (int)ldS) : (size_t)2 * H * talking_pitch_f32((int)ldS) * 4;
    if (s16) SPE_CHECK(stats && smem <= 100 * 1024 + 4096, "spe_talking_softmax_fwd_s16: row does not fit the staged kernel (check spe_talking_s16_supported)");
    if (stats && smem <= 100 * 1024 + 4096 && (s16 || getenv("SPE_TALKING_WARP_ROWS") == nullptr)) {
        static bool done = false;
        if (!done) {
            SPE_CUDA(cudaFuncSetAttribute(talking_fwd_rows_kernel<H, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
            SPE_CUDA(cudaFuncSetAttribute(talking_fwd_rows_kernel<H, 10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
            SPE_CUDA(cudaFuncSetAttribute(talking_fwd_rows_kernel<H, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
            SPE_CUDA(cudaFuncSetAttribute(talking_fwd_rows_kernel<H, 10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
            done = true;
        }
        const long long rows = (long long)B * Nq;
        const int grid = (int)(rows < 2LL * spe_num_sms() ? rows : 2LL * spe_num_sms());
        uint16_t* A16 = reinterpret_cast<uint16_t*>(A);
        const bool w10 = talking_warps((int)ldA) == 10;
        if (s16 && w10) talking_fwd_rows_kernel<H, 10, true><<<grid, 320, smem, st>>>(Sv, A16, Wl, bl, Ww, bw, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA);
        else if (s16) talking_fwd_rows_kernel<H, 8, true><<<grid, 256, smem, st>>>(Sv, A16, Wl, bl, Ww, bw, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA);
        else if (w10) talking_fwd_rows_kernel<H, 10, false><<<grid, 320, smem, st>>>(Sv, A16, Wl, bl, Ww, bw, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA);
        else talking_fwd_rows_kernel<H, 8, false><<<grid, 256, smem, st>>>(Sv, A16, Wl, bl, Ww, bw, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA);
    } else {
        talking_fwd_kernel<H><<<talking_grid(B, Nq, 4), 256, 0, st>>>(S, reinterpret_cast<uint16_t*>(A), Wl, bl, Ww, bw, stats, B * Nq, Nq, Nk, ldS, ldA);
    }
    SPE_LAUNCHED();
    return 0;
}

// SPE_TH8_STREAM=0: keep H = 8 (fp16 logits) on the kernels of this file (A/B switch); default: the register-chained kernels of talking_h8.cu
static bool talking_h8_on(int H) {
    static const bool off = getenv("SPE_TH8_STREAM") != nullptr && getenv("SPE_TH8_STREAM")[0] == '0';
    return H == 8 && !off;
}
static int talking_h8_bwd_run(const void* S, int s16, const void* dA, void* dS, const float* Wl, const float* bl, const float* Ww, const float* stats, int B, int Nq,
                              int Nk, long long ldS, long long ldA, float* dWl, float* dbl, float* dWw, float* dbw, float* workspace, cudaStream_t st);

// SPE_TH16_GENERIC=1: keep H = 16 on the CUDA-core kernels (A/B switch)
static bool talking_h16_on(int H) {
    static const bool off = getenv("SPE_TH16_GENERIC") != nullptr && getenv("SPE_TH16_GENERIC")[0] == '1';
    return H == 16 && !off;
}

extern "C" __attribute__((visibility("default"))) int spe_talking_softmax_fwd(const float* S, void* A, const float* Wl, const float* bl, const float* Ww, const float* bw, float* stats,
                                       int B, int H, int Nq, int Nk, int64_t ldS, int64_t ldA, void* stream) {
    SPE_CHECK(S && A && Wl && bl && Ww && bw && B > 0 && Nq > 0 && Nk > 0, "spe_talking_softmax_fwd: bad argument");
    SPE_CHECK(ldS % 8 == 0 && ldA % 8 == 0 && ldS >= Nk && ldA >= Nk, "spe_talking_softmax_fwd: leading dims must be multiples of 8 and >= Nk");
    switch (H) {
        case 2: return talking_fwd_launch<2>(S, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, ST(stream));
        case 4: return talking_fwd_launch<4>(S, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, ST(stream));
        case 8: return talking_fwd_launch<8>(S, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, ST(stream));
        default: {
            // H = 16 (CaiT-M36): mma.sync formulation in talking_h16.cu; 6 / 12 of the XS variants: CUDA cores, talking_generic.cu
            SpeProfScope prof(SPE_FAM_TALKING_FWD, (double)B * H * Nq * Nk * 6.0, ST(stream));
            if (talking_h16_on(H)) return spe_talking_h16_fwd(S, 0, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, ST(stream));
            return spe_talking_generic_fwd(S, A, Wl, bl, Ww, bw, stats, B, H, Nq, Nk, ldS, ldA, ST(stream));
        }
    }
}

// fp16 logits (S from spe_gemm with c_dtype = SPE_DT_F16): row-staged kernels only, H in {2, 4, 8}
extern "C" __attribute__((visibility("default"))) int spe_talking_s16_supported(int H, int Nk, int64_t ldS, int64_t ldA) {
    if (ldS % 8 != 0 || ldA % 8 != 0 || ldS < Nk || ldA < Nk) return 0;
    if (talking_h16_on(H)) return 1;                     // streamed from global memory: no row-size limit
    if (talking_h8_on(H) && spe_talking_h8_fits(ldS, ldA)) return 1;
    if (H != 2 && H != 4 && H != 8) return 0;
    return talking_fwd_smem16(H, (int)ldS) <= 100 * 1024 + 4096 && talking_bwd_smem16(H, (int)ldS, (int)ldA) <= 104 * 1024 ? 1 : 0;
}
extern "C" __attribute__((visibility("default"))) int spe_talking_softmax_fwd_s16(const void* S16, void* A, const float* Wl, const float* bl, const float* Ww,
                                                                                  const float* bw, float* stats, int B, int H, int Nq, int Nk, int64_t ldS,
                                                                                  int64_t ldA, void* stream) {
    SPE_CHECK(S16 && A && Wl && bl && Ww && bw && stats && B > 0 && Nq > 0 && Nk > 0, "spe_talking_softmax_fwd_s16: bad argument");
    SPE_CHECK(spe_talking_s16_supported(H, Nk, ldS, ldA), "spe_talking_softmax_fwd_s16: unsupported shape H=%d Nk=%d", H, Nk);
    if (H == 16) {
        SpeProfScope prof(SPE_FAM_TALKING_FWD, (double)B * H * Nq * Nk * 4.0, ST(stream));
        return spe_talking_h16_fwd(S16, 1, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, ST(stream));
    }
    if (talking_h8_on(H) && spe_talking_h8_fits(ldS, ldA)) {
        SpeProfScope prof(SPE_FAM_TALKING_FWD, (double)B * H * Nq * Nk * 4.0, ST(stream));
        return spe_talking_h8_fwd(S16, 1, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, ST(stream));
    }
    switch (H) {
        case 2: return talking_fwd_launch<2>(S16, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, ST(stream), true);
        case 4: return talking_fwd_launch<4>(S16, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, ST(stream), true);
        default: return talking_fwd_launch<8>(S16, A, Wl, bl, Ww, bw, stats, B, Nq, Nk, ldS, ldA, ST(stream), true);
    }
}

static int talking_bwd_grid(int B, int Nq) { return talking_grid(B, Nq, 4); }

static int talking_bwd_rows_grid(int B, int Nq);
extern "C" __attribute__((visibility("default"))) int64_t spe_talking_softmax_bwd_workspace(int B, int H, int Nq, int Nk) {
    (void)Nk;
    const int g1 = talking_bwd_grid(B, Nq), g2 = talking_bwd_rows_grid(B, Nq);
    return (int64_t)(g1 > g2 ? g1 : g2) * (2 * H * H + 2 * H);
}

static int talking_bwd_rows_grid(int B, int Nq) {
    const long long rows = (long long)B * Nq;
    return (int)(rows < 2LL * spe_num_sms() ? rows : 2LL * spe_num_sms());
}

template <int H>
static int talking_bwd_launch(const void* Sv, const void* dA, void* dS, const float* Wl, const float* bl, const float* Ww, const float* stats, int B,
                              int Nq, int Nk, int64_t ldS, int64_t ldA, float* dWl, float* dbl, float* dWw, float* dbw, float* ws, cudaStream_t st,
                              bool s16 = false) {
    const float* S = reinterpret_cast<const float*>(Sv);
    const size_t smem = s16 ? talking_bwd_smem16(H, (int)ldS, (int)ldA)
                            : (size_t)H * (talking_pitch_f32((int)ldS) * 4 + talking_pitch_bf16((int)ldA) * 2) + (size_t)((ldA + 31) / 32) * 512;
    const bool rows_ok = stats && smem <= 104 * 1024 && (s16 || getenv("SPE_TALKING_WARP_ROWS") == nullptr);
    if (s16) SPE_CHECK(rows_ok, "spe_talking_softmax_bwd_s16: row does not fit the staged kernel (check spe_talking_s16_supported)");
    const int grid = rows_ok ? talking_bwd_rows_grid(B, Nq) : talking_bwd_grid(B, Nq);
    {
        SpeProfScope prof(SPE_FAM_TALKING_BWD, (double)B * H * Nq * Nk * (s16 ? 6.0 : 8.0), st);   // S read + dA bf16 read + dS bf16 write
        if (rows_ok) {
            static bool done = false;
            if (!done) {
                SPE_CUDA(cudaFuncSetAttribute(talking_bwd_rows_kernel<H, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024));
                SPE_CUDA(cudaFuncSetAttribute(talking_bwd_rows_kernel<H, 10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024));
                SPE_CUDA(cudaFuncSetAttribute(talking_bwd_rows_kernel<H, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024));
                SPE_CUDA(cudaFuncSetAttribute(talking_bwd_rows_kernel<H, 10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024));
                done = true;
            }
            const uint16_t* dA16 = reinterpret_cast<const uint16_t*>(dA);
            uint16_t* dS16 = reinterpret_cast<uint16_t*>(dS);
            const bool w10 = talking_warps((int)ldA) == 10;
            if (s16 && w10) talking_bwd_rows_kernel<H, 10, true><<<grid, 320, smem, st>>>(Sv, dA16, dS16, Wl, bl, Ww, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA, ws);
            else if (s16) talking_bwd_rows_kernel<H, 8, true><<<grid, 256, smem, st>>>(Sv, dA16, dS16, Wl, bl, Ww, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA, ws);
            else if (w10) talking_bwd_rows_kernel<H, 10, false><<<grid, 320, smem, st>>>(Sv, dA16, dS16, Wl, bl, Ww, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA, ws);
            else talking_bwd_rows_kernel<H, 8, false><<<grid, 256, smem, st>>>(Sv, dA16, dS16, Wl, bl, Ww, stats, B * Nq, Nq, Nk, (int)ldS, (int)ldA, ws);
        } else {
            talking_bwd_kernel<H><<<grid, 256, 0, st>>>(S, reinterpret_cast<const uint16_t*>(dA), reinterpret_cast<uint16_t*>(dS), Wl, bl, Ww, B * Nq, Nq, Nk,
                                                        ldS, ldA, ws);
        }
        SPE_LAUNCHED();
    }
    const int NP = 2 * H * H + 2 * H;
    talking_bwd_finalize_kernel<<<(NP + 7) / 8, 256, 0, st>>>(ws, grid, H, dWl, dbl, dWw, dbw);
    SPE_LAUNCHED();
    return 0;
}

static int talking_h8_bwd_run(const void* S, int s16, const void* dA, void* dS, const float* Wl, const float* bl, const float* Ww, const float* stats, int B, int Nq,
                              int Nk, long long ldS, long long ldA, float* dWl, float* dbl, float* dWw, float* dbw, float* workspace, cudaStream_t st) {
    const int H = 8, grid = spe_talking_h8_grid(B, Nq);
    {
        SpeProfScope prof(SPE_FAM_TALKING_BWD, (double)B * H * Nq * Nk * (s16 ? 6.0 : 8.0), st);
        if (spe_talking_h8_bwd(S, s16, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, workspace, st)) return -1;
    }
    const int NP = 2 * H * H + 2 * H;
    talking_bwd_finalize_kernel<<<(NP + 7) / 8, 256, 0, st>>>(workspace, grid, H, dWl, dbl, dWw, dbw);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_talking_softmax_bwd_s16(const void* S16, const void* dA, void* dS, const float* Wl, const float* bl,
                                                                                  const float* Ww, const float* bw, const float* stats, int B, int H, int Nq, int Nk,
                                                                                  int64_t ldS, int64_t ldA, float* dWl, float* dbl, float* dWw, float* dbw,
                                                                                  float* workspace, int64_t workspace_floats, void* stream) {
    (void)bw;
    SPE_CHECK(S16 && dA && dS && Wl && bl && Ww && stats && dWl && dbl && dWw && dbw && workspace, "spe_talking_softmax_bwd_s16: null argument");
    SPE_CHECK(spe_talking_s16_supported(H, Nk, ldS, ldA), "spe_talking_softmax_bwd_s16: unsupported shape H=%d Nk=%d", H, Nk);
    SPE_CHECK(workspace_floats >= spe_talking_softmax_bwd_workspace(B, H, Nq, Nk), "spe_talking_softmax_bwd_s16: workspace too small");
    if (talking_h8_on(H) && spe_talking_h8_fits(ldS, ldA)) return talking_h8_bwd_run(S16, 1, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, dWl, dbl, dWw, dbw, workspace, ST(stream));
    if (H == 16) {
        const int grid = spe_talking_h16_grid(B, Nq);
        {
            SpeProfScope prof(SPE_FAM_TALKING_BWD, (double)B * H * Nq * Nk * 6.0, ST(stream));
            if (spe_talking_h16_bwd(S16, 1, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, workspace, ST(stream))) return -1;
        }
        const int NP = 2 * H * H + 2 * H;
        talking_bwd_finalize_kernel<<<(NP + 7) / 8, 256, 0, ST(stream)>>>(workspace, grid, H, dWl, dbl, dWw, dbw);
        SPE_LAUNCHED();
        return 0;
    }
    switch (H) {
        case 2: return talking_bwd_launch<2>(S16, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, dWl, dbl, dWw, dbw, workspace, ST(stream), true);
        case 4: return talking_bwd_launch<4>(S16, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, dWl, dbl, dWw, dbw, workspace, ST(stream), true);
        default: return talking_bwd_launch<8>(S16, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, dWl, dbl, dWw, dbw, workspace, ST(stream), true);
    }
}

extern "C" __attribute__((visibility("default"))) int spe_talking_softmax_bwd(const float* S, const void* dA, void* dS, const float* Wl, const float* bl, const float* Ww, const float* bw,
                                       const float* stats, int B, int H, int Nq, int Nk, int64_t ldS, int64_t ldA, float* dWl, float* dbl, float* dWw,
                                       float* dbw, float* workspace, int64_t workspace_floats, void* stream) {
    (void)bw;
    SPE_CHECK(S && dA && dS && Wl && bl && Ww && dWl && dbl && dWw && dbw && workspace, "spe_talking_softmax_bwd: null argument");
    SPE_CHECK(ldS % 8 == 0 && ldA % 8 == 0 && ldS >= Nk && ldA >= Nk, "spe_talking_softmax_bwd: leading dims must be multiples of 8 and >= Nk");
    SPE_CHECK(workspace_floats >= spe_talking_softmax_bwd_workspace(B, H, Nq, Nk), "spe_talking_softmax_bwd: workspace too small");
    switch (H) {
        case 2: return talking_bwd_launch<2>(S, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, dWl, dbl, dWw, dbw, workspace, ST(stream));
        case 4: return talking_bwd_launch<4>(S, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, dWl, dbl, dWw, dbw, workspace, ST(stream));
        case 8: return talking_bwd_launch<8>(S, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, dWl, dbl, dWw, dbw, workspace, ST(stream));
        default: {
            const int grid = talking_h16_on(H) ? spe_talking_h16_grid(B, Nq) : spe_talking_generic_grid(B, Nq);
            {
                SpeProfScope prof(SPE_FAM_TALKING_BWD, (double)B * H * Nq * Nk * 8.0, ST(stream));
                if (talking_h16_on(H)) {
                    if (spe_talking_h16_bwd(S, 0, dA, dS, Wl, bl, Ww, stats, B, Nq, Nk, ldS, ldA, workspace, ST(stream))) return -1;
                } else if (spe_talking_generic_bwd(S, dA, dS, Wl, bl, Ww, stats, B, H, Nq, Nk, ldS, ldA, workspace, ST(stream))) return -1;
            }
            const int NP = 2 * H * H + 2 * H;
            talking_bwd_finalize_kernel<<<(NP + 7) / 8, 256, 0, ST(stream)>>>(workspace, grid, H, dWl, dbl, dWw, dbw);
            SPE_LAUNCHED();
            return 0;
        }
    }
}

extern "C" __attribute__((visibility("default"))) int spe_layerscale_bwd(const float* dout, const void* y_bf16, const float* gamma, int64_t rows, int D, void* dy_bf16, float* dgamma,
                                  float* dbias, void* stream) {
    SPE_CHECK(dout && y_bf16 && gamma && dy_bf16 && rows > 0 && D % 4 == 0, "spe_layerscale_bwd: bad argument");
    const int gx = (D + 127) / 128;
    long long gy = (rows + 7) / 8;
    const long long cap = ((long long)spe_num_sms() * 4 + gx - 1) / gx;
    if (gy > cap) gy = cap;
    layerscale_bwd_kernel<<<dim3(gx, (unsigned)gy), dim3(32, 8), 0, ST(stream)>>>(dout, reinterpret_cast<const uint16_t*>(y_bf16), gamma, rows, D,
                                                                                    reinterpret_cast<uint16_t*>(dy_bf16), dgamma, dbias);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_colsum_bf16(const void* x, int64_t rows, int N, int64_t ld, float* out, void* stream) {
    SPE_CHECK(x && out && rows > 0 && N > 0 && ld % 2 == 0, "spe_colsum_bf16: bad argument (ld must be even)");
    const int gx = (N + 63) / 64;
    long long gy = (rows + 7) / 8;
    const long long cap = ((long long)spe_num_sms() * 4 + gx - 1) / gx;
    if (gy > cap) gy = cap;
    colsum_bf16_kernel<<<dim3(gx, (unsigned)gy), dim3(32, 8), 0, ST(stream)>>>(reinterpret_cast<const uint16_t*>(x), rows, N, ld, out, 0, 0);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_colsum_bf16_batched(const void* x, int batch, int64_t rows, int N, int64_t ld, int64_t batch_stride,
                                                                             float* out, void* stream) {
    SPE_CHECK(x && out && batch > 0 && rows > 0 && N > 0 && ld % 2 == 0 && batch_stride % 2 == 0, "spe_colsum_bf16_batched: bad argument");
    const int gx = (N + 63) / 64;
    long long gy = (rows + 7) / 8;
    const long long cap = ((long long)spe_num_sms() * 4 + (long long)gx * batch - 1) / ((long long)gx * batch);
    if (gy > cap) gy = cap < 1 ? 1 : cap;
    colsum_bf16_kernel<<<dim3(gx, (unsigned)gy, batch), dim3(32, 8), 0, ST(stream)>>>(reinterpret_cast<const uint16_t*>(x), rows, N, ld, out, batch_stride, N);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_axpby_cast(const float* x, const float* y, float a, float b, int64_t n, void* out_bf16, float* out_f32, void* stream) {
    SPE_CHECK(x && n > 0 && (out_bf16 || out_f32), "spe_axpby_cast: bad argument");
    axpby_cast_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, ST(stream)>>>(x, y, a, b, n, reinterpret_cast<uint16_t*>(out_bf16), out_f32);
    SPE_LAUNCHED();
    return 0;
}

namespace {
__global__ void __launch_bounds__(256) cast_multi_kernel(const spe_cast_seg* __restrict__ segs) {
    const spe_cast_seg sg = segs[blockIdx.y];
    const float* __restrict__ src = sg.src;
    uint16_t* __restrict__ dst = reinterpret_cast<uint16_t*>(sg.dst);
    const long long n = sg.n;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (vec) {
        const long long n4 = n >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
            reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        }
        for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = f_to_bf16(src[i]);
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = f_to_bf16(src[i]);
    }
}
}  // namespace

extern "C" __attribute__((visibility("default"))) int spe_cast_f32_to_bf16_multi(const spe_cast_seg* segs, int count, int64_t max_n, void* stream) {
    SPE_CHECK(segs && count > 0 && max_n > 0, "spe_cast_f32_to_bf16_multi: bad argument");
    SPE_CHECK(count <= 65535, "spe_cast_f32_to_bf16_multi: too many segments");
    long long gx = (max_n / 4 + 255) / 256;
    if (gx > 32) gx = 32;
    if (gx < 1) gx = 1;
    cast_multi_kernel<<<dim3((unsigned)gx, (unsigned)count), 256, 0, ST(stream)>>>(segs);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream) {
    SPE_CHECK(x && y && n > 0, "spe_cast_bf16_to_f32: bad argument");
    cast_bf16_f32_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const uint16_t*>(x), y, n, 0);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_add_bf16_into_f32(const void* x, float* out, int64_t n, void* stream) {
    SPE_CHECK(x && out && n > 0, "spe_add_bf16_into_f32: bad argument");
    cast_bf16_f32_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const uint16_t*>(x), out, n, 1);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_im2col_patch(const float* img, int B, int H, int W, int p, void* out_bf16, void* stream) {
    SPE_CHECK(img && out_bf16 && B > 0 && p % 4 == 0 && H >= p && W >= p, "spe_im2col_patch: bad argument");      // H, W need not be multiples of p (conv floors)
    const long long total = (long long)B * (H / p) * (W / p) * 3 * p * p / 4;
    im2col_patch_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(img, B, H, W, p, reinterpret_cast<uint16_t*>(out_bf16));
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_bicubic_tokens_fwd(const float* src, int sh, int sw, int D, float* dst, int dh, int dw, void* stream) {
    SPE_CHECK(src && dst && sh > 0 && sw > 0 && dh > 0 && dw > 0 && D > 0, "spe_bicubic_tokens_fwd: bad argument");
    bicubic_tokens_kernel<false><<<grid_for((long long)dh * dw * D, 256), 256, 0, ST(stream)>>>(src, dst, sh, sw, dh, dw, D);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_bicubic_tokens_bwd(const float* ddst, int dh, int dw, int D, float* dsrc, int sh, int sw, void* stream) {
    SPE_CHECK(ddst && dsrc && sh > 0 && sw > 0 && dh > 0 && dw > 0 && D > 0, "spe_bicubic_tokens_bwd: bad argument");
    bicubic_tokens_kernel<true><<<grid_for((long long)dh * dw * D, 256), 256, 0, ST(stream)>>>(ddst, dsrc, sh, sw, dh, dw, D);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_sine_pos_2d(const uint8_t* mask, int B, int h, int w, int D, float* pos, void* pos_bf16, void* stream) {
    SPE_CHECK(mask && (pos || pos_bf16) && B > 0 && h > 0 && w > 0 && D % 4 == 0, "spe_sine_pos_2d: bad argument");
    sine_pos_2d_kernel<<<grid_for((long long)B * h * w * (D / 2), 256), 256, 0, ST(stream)>>>(mask, B, h, w, D, pos, reinterpret_cast<uint16_t*>(pos_bf16));
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_query_sine_fwd(const float* ref, int64_t n, int D, float* emb, void* stream) {
    SPE_CHECK(ref && emb && n > 0 && D % 4 == 0, "spe_query_sine_fwd: bad argument");
    query_sine_kernel<false><<<grid_for(n * (D / 2), 256), 256, 0, ST(stream)>>>(ref, nullptr, n, D, emb);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_query_sine_bwd(const float* ref, const float* demb, int64_t n, int D, float* dref, void* stream) {
    SPE_CHECK(ref && demb && dref && n > 0 && D % 4 == 0, "spe_query_sine_bwd: bad argument");
    query_sine_kernel<true><<<(unsigned)((n + 7) / 8), 256, 0, ST(stream)>>>(ref, demb, n, D, dref);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_relu_bwd_bf16(const void* dout, const void* h, void* out, int64_t n, void* stream) {
    SPE_CHECK(dout && h && out && n > 0, "spe_relu_bwd_bf16: bad argument");
    relu_bwd_kernel<<<grid_for((n + 1) / 2, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const uint16_t*>(dout), reinterpret_cast<const uint16_t*>(h),
                                                                         reinterpret_cast<uint16_t*>(out), n);
    SPE_LAUNCHED();
    return 0;
}
