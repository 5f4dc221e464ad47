// tcgen05 / TMEM / TMA GEMM for sm_100a:  C = epilogue(alpha * A * B^T)        (persistent, warp-specialised)
//
//   * one persistent CTA per SM walks the output tiles (128 x BN); 12 warps:
//       warp 0  : TMA producer  -- bf16 operand tiles global->shared (cp.async.bulk.tensor, SWIZZLE_128B) into a
//                 STAGES-deep mbarrier ring that runs continuously across tiles;
//       warp 1  : MMA issuer    -- one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16),
//                 fp32 accumulator in TMEM, DOUBLE BUFFERED (2 x BN columns) so tile i+1's main loop overlaps tile i's
//                 epilogue; tcgen05.commit releases smem stages and publishes the accumulator;
//       warp 2  : TMEM allocator;   warps 4..11 : epilogue (warp%4 = TMEM lane quarter, (warp-4)/4 = 64-column half);
//   * both K-major and MN-major operand layouts are consumed directly (UMMA descriptor major bits), so dgrad
//     (B = the weight, MN-major) and wgrad (both MN-major) need no transposed copies;
//   * epilogue: tcgen05.ld (32x32b) -> bias / activation / LayerScale / residual in registers; every tile-shaped
//     operand moves through 32-row x 128-byte SWIZZLE_128B shared-memory slabs with TMA (residual / aux_in: bulk tensor
//     loads; C / aux_out: bulk tensor stores) so epilogue HBM traffic is full-line and asynchronous; operands that are
//     not 16-byte regular (e.g. the 81-wide logits) fall back to direct per-thread global accesses;
//   * split-K for few-tile / long-K shapes (wgrad) with TMA reduce-add (cp.reduce.async.bulk.tensor .add.f32).
//
// Reference call sites replaced: see include/spe_b200.h (spe_gemm).
#include "common.cuh"
#include <cuda.h>
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 8;
constexpr uint32_t CHUNK_BYTES = 64 * BK * 2;   // one 64(mn) x 64(k) MN-major TMA box
constexpr uint32_t SLAB = 4096;                 // 32 rows x 128 B
constexpr int EPI_SLABS = 3;                   // per epilogue warp: {RC0, RC1, X}
constexpr uint32_t EPI_SMEM = NUM_EPI_WARPS * EPI_SLABS * SLAB;

struct EpiParams {
    void* C; int c_dtype; long long ldc, c_sb1, c_sb2;
    float alpha; const float* bias; int act;
    const uint16_t* aux_in; uint16_t* aux_out; long long ld_aux;
    const float* gamma; const float* residual; long long ldr, r_sb1, r_sb2;
    int split, split_stride;
    int M, N, K, batch2;
    int a_m1, a_m2, b_m1, b_m2;   // 0 = operand broadcast over that batch dim (stride 0), else 1
    int r_m1, r_m2;
    int tma_io;                   // epilogue tiles through TMA (else direct per-thread global access)
    int splits, kb_per_split;     // split-K
    uint32_t fd_n[2], fd_m[2], fd_s[2], fd_b[2];   // {mul, shift} of the divisions by tiles_n / tiles_m / splits / batch2 (tile decode)
    int accum;                    // C += ... (residual aliased C on entry): every tile is stored with a reduce-add
    int tiles_m, tiles_n, num_tiles;
    int mode, mode_nl;            // EM_* epilogue specialisation for lead / non-lead (split-K) tiles
    int dbg;                      // timing experiments only (SPE_GEMM_DBG bitmask): 1 no TMA store, 2 no smem writes, 4 no TMEM load, 8 no proxy fence
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// 16-byte chunk j of row r inside a 32-row x 128-byte slab laid out with the TMA SWIZZLE_128B pattern
__device__ __forceinline__ uint32_t slab_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

// UMMA shared-memory descriptor, SWIZZLE_128B, version 1 (sm_100).  lbo/sbo in bytes.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;       // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA-pair (cta_group::2) variants: the leader CTA (cluster rank 0) issues the MMAs for both CTAs of the cluster
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {          // shared::cta address -> shared::cluster address in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are signalled on an mbarrier of the LEADER CTA (cluster address), data into this CTA's smem
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {       // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// one lane of a CONVERGED warp (the tcgen05.mma operands must be warp-uniform: issued from inside an `if (lane == 0)` region the
// compiler cannot prove that and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST waterfall loop, ~140 cycles per MMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

#define TMEM_LD_32x32b_X32(taddr, r)                                                                         \
    asm volatile(                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                            \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"            \
        : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]), "=r"((r)[7]),     \
          "=r"((r)[8]), "=r"((r)[9]), "=r"((r)[10]), "=r"((r)[11]), "=r"((r)[12]), "=r"((r)[13]), "=r"((r)[14]), "=r"((r)[15]), \
          "=r"((r)[16]), "=r"((r)[17]), "=r"((r)[18]), "=r"((r)[19]), "=r"((r)[20]), "=r"((r)[21]), "=r"((r)[22]), "=r"((r)[23]), \
          "=r"((r)[24]), "=r"((r)[25]), "=r"((r)[26]), "=r"((r)[27]), "=r"((r)[28]), "=r"((r)[29]), "=r"((r)[30]), "=r"((r)[31]) \
        : "r"(taddr))

#define TMEM_LD_32x32b_X16(taddr, r)                                                                         \
    asm volatile(                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                            \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                     \
        : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]), "=r"((r)[7]),     \
          "=r"((r)[8]), "=r"((r)[9]), "=r"((r)[10]), "=r"((r)[11]), "=r"((r)[12]), "=r"((r)[13]), "=r"((r)[14]), "=r"((r)[15]) \
        : "r"(taddr))

__device__ __forceinline__ float apply_act(float v, int act, float aux) {
    switch (act) {
        case SPE_ACT_RELU: return fmaxf(v, 0.f);
        case SPE_ACT_GELU: return gelu_erf(v);
        case SPE_ACT_RELU_GRAD: return aux > 0.f ? v : 0.f;
        case SPE_ACT_GELU_GRAD: return v * gelu_erf_grad(aux);
        default: return v;
    }
}

// direct (non-TMA) epilogue of one row x 64 accumulator columns: operands that are not 16-byte regular
__device__ __noinline__ void epilogue_direct(const EpiParams& ep, const uint32_t* r, int m, int nb, int b1, int b2, bool lead) {
    if (m >= ep.M) return;
    const long long boff_c = (long long)b1 * ep.c_sb1 + (long long)b2 * ep.c_sb2;
    const long long boff_r = (long long)(b1 * ep.r_m1) * ep.r_sb1 + (long long)(b2 * ep.r_m2) * ep.r_sb2;
    float* Cf = reinterpret_cast<float*>(ep.C) + boff_c + (long long)m * ep.ldc;
    uint16_t* Cb = reinterpret_cast<uint16_t*>(ep.C) + boff_c + (long long)m * ep.ldc;
    const float* Rr = (ep.residual && lead) ? ep.residual + boff_r + (long long)m * ep.ldr : nullptr;
    const uint16_t* auxi = ep.aux_in ? ep.aux_in + (long long)m * ep.ld_aux : nullptr;
    uint16_t* auxo = ep.aux_out ? ep.aux_out + (long long)m * ep.ld_aux : nullptr;
    for (int j = 0; j < 64; ++j) {
        const int n = nb + j;
        if (n >= ep.N) break;
        float x = __uint_as_float(r[j]) * ep.alpha;
        if (ep.bias && lead) x += __ldg(ep.bias + n);
        if (auxo) auxo[n] = f_to_bf16(x);
        if (ep.act != SPE_ACT_NONE) x = apply_act(x, ep.act, auxi ? bf16_to_f(auxi[n]) : 0.f);
        if (ep.gamma) x *= __ldg(ep.gamma + n);
        if (Rr) x += Rr[n];
        const int dn = ep.split > 0 ? (n / ep.split) * ep.split_stride + (n % ep.split) : n;
        if (ep.splits > 1 || ep.accum) atomicAdd(Cf + dn, x);
        else if (ep.c_dtype == SPE_DT_F32) Cf[dn] = x;
        else Cb[dn] = ep.c_dtype == SPE_DT_F16 ? f_to_f16(x) : f_to_bf16(x);
    }
}


// ------------------------------------------------------------------------------------------------
// Epilogue math for one 64-column super-chunk of one row (thread = row), accumulator already in registers.
// Fully specialised at compile time: no branches, all 64 columns unrolled -> ~150 independent instructions per warp.
// (The previous run-time-flag version executed ~500 dependent, branchy instructions per super-chunk; with 2 epilogue
//  warps per scheduler that made EVERY GEMM epilogue-latency bound: the MMA warp spent its time waiting on `tempty`.)
//   ACT: 0 none, 1 relu, 2 gelu, 3 relu_grad (needs Xi), 4 gelu_grad (needs Xi)
// ------------------------------------------------------------------------------------------------
template <bool C32, bool BIAS, int ACT, bool GAMMA, bool RES, bool XO, bool F16 = false>
__device__ __forceinline__ void epi_fast(const EpiParams& ep, const uint32_t (&r)[64], int nb, int lane, uint32_t sRC, uint32_t sX) {
    const uint32_t rowoff = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const int n = nb + g * 8;
        float v[8];
        if constexpr (BIAS) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + n)), b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + n + 4));
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(__uint_as_float(r[g * 8 + j]), ep.alpha, bb[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]) * ep.alpha;
        }
        if constexpr (XO) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sX + rowoff + (((uint32_t)g ^ sw) << 4)), "r"(pack_bf16x2(v[0], v[1])),
                         "r"(pack_bf16x2(v[2], v[3])), "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7])) : "memory");
        }
        if constexpr (ACT == SPE_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if constexpr (ACT == SPE_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
        } else if constexpr (ACT == SPE_ACT_RELU_GRAD || ACT == SPE_ACT_GELU_GRAD) {
            uint4 a;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(sX + rowoff + (((uint32_t)g ^ sw) << 4)));
            const float2 t0 = unpack_bf16x2(a.x), t1 = unpack_bf16x2(a.y), t2 = unpack_bf16x2(a.z), t3 = unpack_bf16x2(a.w);
            const float ax[8] = {t0.x, t0.y, t1.x, t1.y, t2.x, t2.y, t3.x, t3.y};
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (ACT == SPE_ACT_RELU_GRAD) ? (ax[j] > 0.f ? v[j] : 0.f) : v[j] * gelu_erf_grad(ax[j]);
        }
        if constexpr (GAMMA) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(ep.gamma + n)), g1 = __ldg(reinterpret_cast<const float4*>(ep.gamma + n + 4));
            v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w; v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
        }
        if constexpr (RES) {
            // fp32 residual slabs; the C tile overlays them (each thread reads its chunk for group g before writing group g)
            const uint32_t sl = sRC + (g >> 2) * SLAB + rowoff;
            float4 r0, r1;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r0.x), "=f"(r0.y), "=f"(r0.z), "=f"(r0.w) : "r"(sl + (((uint32_t)((g & 3) * 2) ^ sw) << 4)));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r1.x), "=f"(r1.y), "=f"(r1.z), "=f"(r1.w) : "r"(sl + (((uint32_t)((g & 3) * 2 + 1) ^ sw) << 4)));
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        }
        if constexpr (C32) {
            const uint32_t sl = sRC + (g >> 2) * SLAB + rowoff;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sl + (((uint32_t)((g & 3) * 2) ^ sw) << 4)), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sl + (((uint32_t)((g & 3) * 2 + 1) ^ sw) << 4)), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
        } else if constexpr (F16) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sRC + rowoff + (((uint32_t)g ^ sw) << 4)), "r"(pack_f16x2_rn(v[0], v[1])),
                         "r"(pack_f16x2_rn(v[2], v[3])), "r"(pack_f16x2_rn(v[4], v[5])), "r"(pack_f16x2_rn(v[6], v[7])) : "memory");
        } else {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sRC + rowoff + (((uint32_t)g ^ sw) << 4)), "r"(pack_bf16x2(v[0], v[1])),
                         "r"(pack_bf16x2(v[2], v[3])), "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7])) : "memory");
        }
    }
}

// the specialisations that exist (host picks one; anything else takes the generic run-time path)
enum { EM_GENERIC = 0, EM_F32, EM_BF16, EM_BF16_BIAS, EM_F32_BIAS, EM_F32_BIAS_GAMMA_RES_XO, EM_F32_BIAS_RES, EM_F32_RES, EM_BF16_BIAS_RES,
       EM_BF16_BIAS_GELU_XO, EM_BF16_BIAS_RELU, EM_BF16_RELUGRAD, EM_BF16_GELUGRAD, EM_F16 };

// CG2 = CTA-pair mode: a cluster of 2 CTAs computes a 256 x BN tile with tcgen05.mma.cta_group::2 -- each CTA stages its own 128 rows
// of A and HALF of the B tile (BN/2 rows), the tensor core reads both halves, so the SM<->L2 operand traffic per flop drops by
// 25% (BN = 128) to 50% (BN = 256) against the single-CTA 128 x 128 tile.  Only the leader CTA issues MMAs; smem stages are
// released and accumulators published in both CTAs by multicast commits; every other role runs unchanged in both CTAs.
template <int BN, int STAGES, bool A_MN, bool B_MN, bool CG2>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                      const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                                                                      const __grid_constant__ CUtensorMap tmXi, const __grid_constant__ CUtensorMap tmXo,
                                                                      const EpiParams ep) {
    constexpr int BNL = CG2 ? BN / 2 : BN;             // B rows staged by THIS CTA
    constexpr uint32_t A_BYTES = BM * BK * 2;
    constexpr uint32_t B_BYTES = BNL * BK * 2;
    constexpr uint32_t TMEM_COLS = 2 * BN;               // double-buffered accumulator (BN in {64,128,256} -> 128 / 256 / 512 columns)
    static_assert(TMEM_COLS <= 512, "TMEM has 512 columns");
    // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((CG2 ? 2 * BM : BM) >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint8_t* sEpi = sB + STAGES * B_BYTES;                                  // 8 warps x 4 slabs
    uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + EPI_SMEM);          // full[S], empty[S], tfull[2], tempty[2], ebar[8]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + NUM_EPI_WARPS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // let the next kernel of the stream (if it was launched with programmatic stream serialization) start its prologue now; it
    // blocks at its own griddepcontrol.wait until this grid has completed
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t cta_rank = CG2 ? cluster_ctarank() : 0u;
    const int tile0 = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;            // first tile / tile stride of this CTA (pair)
    const int tstride = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int total_kb = (ep.K + BK - 1) / BK;
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, tfull0 = empty0 + 8 * STAGES, tempty0 = tfull0 + 16, ebar0 = tempty0 + 16;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull0 + 8 * b, 1);
            // one epilogue group (4 warps = 128 TMEM lanes) owns a buffer; pair mode: the groups of BOTH CTAs report to the leader
            mbar_init(tempty0 + 8 * b, CG2 ? NUM_EPI_WARPS : NUM_EPI_WARPS / 2);
        }
        for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(ebar0 + 8 * w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CG2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CG2) cluster_sync_all();             // both CTAs' barriers are initialised before any cross-CTA signal
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch -- and the launch latency
    // itself) overlaps the tail of the preceding GEMM in the stream; nothing before this point touches global memory.  The wait
    // returns when the preceding grid has completed and flushed (no-op for an ordinary launch).
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // tile -> coordinates (n fastest: consecutive CTAs share the A rows in L2)
    // (every role decodes every tile: the divisions are multiply-shift with host-made magics -- real divisions cost each epilogue
    //  warp ~5 dependent I2F/RCP/F2I chains per tile, which was a third of the per-tile epilogue latency on the K = 48 GEMMs)
    auto fdiv = [](uint32_t n, const uint32_t (&fd)[2]) -> uint32_t { return (__umulhi(n, fd[0]) + n) >> fd[1]; };
    auto decode = [&](int t, int& m0, int& n0, int& b1, int& b2, int& kb0, int& nkb, int& split) {
        const uint32_t q1 = fdiv((uint32_t)t, ep.fd_n);
        const int tn = t - (int)q1 * ep.tiles_n;
        const uint32_t z = fdiv(q1, ep.fd_m);
        const int tm = (int)q1 - (int)z * ep.tiles_m;
        const uint32_t zb = fdiv(z, ep.fd_s);
        split = (int)z - (int)zb * ep.splits;
        const uint32_t q4 = fdiv(zb, ep.fd_b);
        b1 = (int)q4; b2 = (int)zb - (int)q4 * ep.batch2;
        m0 = tm * (CG2 ? 2 * BM : BM) + (int)cta_rank * BM; n0 = tn * BN;
        kb0 = split * ep.kb_per_split;
        nkb = min(ep.kb_per_split, total_kb - kb0);
    };

    if (warp == 0) {
        {
            // ---------------- TMA producer: the whole (converged) warp walks the loop, one elected lane issues ----------------
            uint32_t it = 0;
            for (int t = tile0; t < ep.num_tiles; t += tstride) {
                int m0, n0, b1, b2, kb0, nkb, split;
                decode(t, m0, n0, b1, b2, kb0, nkb, split);
                const int nl0 = n0 + (int)cta_rank * BNL;                  // pair mode: this CTA stages B rows [nl0, nl0 + BN/2)
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);
                    const uint32_t a_dst = smem_u32(sA + s * A_BYTES), b_dst = smem_u32(sB + s * B_BYTES);
                    const int kc = (kb0 + kb) * BK;
                    if (elect_one()) {
                        if constexpr (CG2) {
                            // both CTAs' loads complete on the LEADER's full barrier; the leader alone arms it with the bytes of the pair
                            const uint32_t fbc = mapa_rank(full0 + 8 * s, 0u);
                            if (cta_rank == 0) mbar_expect_tx(full0 + 8 * s, 2u * (A_BYTES + B_BYTES));
                            if constexpr (!A_MN) {
                                tma_load_4d_pair(a_dst, &tmA, fbc, kc, m0, b2 * ep.a_m2, b1 * ep.a_m1);
                            } else {
#pragma unroll
                                for (int c = 0; c < BM / 64; ++c) tma_load_4d_pair(a_dst + c * CHUNK_BYTES, &tmA, fbc, m0 + c * 64, kc, b2 * ep.a_m2, b1 * ep.a_m1);
                            }
                            if constexpr (!B_MN) {
                                tma_load_4d_pair(b_dst, &tmB, fbc, kc, nl0, b2 * ep.b_m2, b1 * ep.b_m1);
                            } else {
#pragma unroll
                                for (int c = 0; c < BNL / 64; ++c) tma_load_4d_pair(b_dst + c * CHUNK_BYTES, &tmB, fbc, nl0 + c * 64, kc, b2 * ep.b_m2, b1 * ep.b_m1);
                            }
                        } else {
                            const uint32_t fb = full0 + 8 * s;
                            const bool skipA = (ep.dbg & 16) != 0, skipB = (ep.dbg & 32) != 0;      // timing experiments only
                            mbar_expect_tx(fb, (skipA ? 0u : A_BYTES) + (skipB ? 0u : B_BYTES));
                            if (skipA) {
                            } else if constexpr (!A_MN) {
                                tma_load_4d(a_dst, &tmA, fb, kc, m0, b2 * ep.a_m2, b1 * ep.a_m1);
                            } else {
#pragma unroll
                                for (int c = 0; c < BM / 64; ++c) tma_load_4d(a_dst + c * CHUNK_BYTES, &tmA, fb, m0 + c * 64, kc, b2 * ep.a_m2, b1 * ep.a_m1);
                            }
                            if (skipB) {
                            } else if constexpr (!B_MN) {
                                tma_load_4d(b_dst, &tmB, fb, kc, n0, b2 * ep.b_m2, b1 * ep.b_m1);
                            } else {
#pragma unroll
                                for (int c = 0; c < BN / 64; ++c) tma_load_4d(b_dst + c * CHUNK_BYTES, &tmB, fb, n0 + c * 64, kc, b2 * ep.b_m2, b1 * ep.b_m1);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (cta_rank == 0) {
            // ---------------- MMA issuer (pair mode: leader CTA only): the whole warp walks the loop, one elected lane issues ----------------
            uint32_t it = 0, ti = 0;
            for (int t = tile0; t < ep.num_tiles; t += tstride, ++ti) {
                int m0, n0, b1, b2, kb0, nkb, split;
                decode(t, m0, n0, b1, b2, kb0, nkb, split);
                const uint32_t buf = ti & 1u, use = ti >> 1;
                mbar_wait(tempty0 + 8 * buf, (use & 1u) ^ 1u);            // epilogue has drained this accumulator buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + buf * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(full0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_base = smem_u32(sA + s * A_BYTES), b_base = smem_u32(sB + s * B_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // K-major: 16 k-elements = 32 B inside the 128B swizzle row; SBO = 8 rows * 128 B.
                            // MN-major: 16 k-rows = 2 groups of 8 rows (SBO = 1024 B); LBO = next 64-wide MN chunk.
                            const uint64_t ad = A_MN ? umma_desc(a_base + k * 2048, CHUNK_BYTES, 1024) : umma_desc(a_base + k * 32, 0, 1024);
                            const uint64_t bd = B_MN ? umma_desc(b_base + k * 2048, CHUNK_BYTES, 1024) : umma_desc(b_base + k * 32, 0, 1024);
                            if (ep.dbg & 128) continue;      // timing experiment: no MMA issue
                            if constexpr (CG2) umma_f16_pair(tacc, ad, bd, IDESC, (kb | k) != 0 ? 1u : 0u);
                            else umma_f16(tacc, ad, bd, IDESC, (kb | k) != 0 ? 1u : 0u);
                        }
                        if constexpr (CG2) umma_commit_pair(empty0 + 8 * s);
                        else umma_commit(empty0 + 8 * s);       // frees the smem stage when these MMAs retire
                    }
                    __syncwarp();
                }
                if (elect_one()) {
                    if constexpr (CG2) umma_commit_pair(tfull0 + 8 * buf);
                    else umma_commit(tfull0 + 8 * buf);         // accumulator of this tile complete
                }
                __syncwarp();
            }
        }
        __syncwarp();
    } else if (warp >= EPI_WARP0) {
        // ---------------- epilogue warps ----------------
        const int e = warp - EPI_WARP0;
        const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32)   (hardware: warp%4)
        const int grp = e >> 2;                       // epilogue group: group g drains accumulator buffer g = every second tile of this CTA,
                                                      // so two tiles' epilogue latency chains (TMEM load, math, fence, bulk store) overlap
        const uint32_t sBase = smem_u32(sEpi + e * EPI_SLABS * SLAB);
        // slab assignment: C (+ residual, which it overlays) needs 2 slabs when fp32 is involved, aux needs 1.
        //   footprint <= 2 slabs -> two alternating sets {0,1} / {2,3}: the bulk stores of one super-chunk drain while the next is
        //   produced (wait_group.read 1);  footprint 3 -> slabs {0,1} + {2}, single-buffered (wait_group.read 0).
        const bool wide = ep.c_dtype == SPE_DT_F32 || ep.residual != nullptr;
        const bool alternate = false;   // (two alternating slab sets were measured: no gain -- the GEMMs are SM<->L2 fabric bound)
        uint32_t set = 0;
        const uint32_t ebar = ebar0 + 8 * e;
        const bool has_r = ep.residual != nullptr, has_xi = ep.aux_in != nullptr, has_xo = ep.aux_out != nullptr;
        const bool c32 = ep.c_dtype == SPE_DT_F32;
        uint32_t eph = 0, use = 0;
        bool stores_pending = false;
        const uint32_t tempty_remote0 = CG2 ? mapa_rank(tempty0, 0u) : 0u;        // the leader's tempty barriers (cluster address)
        auto release_acc = [&](uint32_t buf) {                                    // lane 0: hand the accumulator buffer back to the MMA warp
            if constexpr (CG2) mbar_arrive_cluster(tempty_remote0 + 8 * buf);
            else mbar_arrive(tempty0 + 8 * buf);
        };
        for (int t = tile0 + grp * tstride; t < ep.num_tiles; t += 2 * tstride, ++use) {
            int m0, n0, b1, b2, kb0, nkb, split;
            decode(t, m0, n0, b1, b2, kb0, nkb, split);
            const uint32_t buf = (uint32_t)grp;
            const bool lead = split == 0;             // split-K: bias / residual contributed once
            const int mrow = m0 + quarter * 32;
            mbar_wait(tfull0 + 8 * buf, use & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (ep.dbg & 64) {                        // timing experiment: barrier protocol only, no epilogue work
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) release_acc(buf);
                continue;
            }
            bool arrived = false;
#pragma unroll 1
            for (int sc = 0; sc < BN / 64; ++sc) {
                const int nb = n0 + sc * 64;
                const bool live = nb < ep.N && mrow < ep.M;
                const bool loads = live && ep.tma_io && ((has_r && lead) || has_xi);
                const uint32_t sRC = sBase + (alternate ? set * 2 * SLAB : 0u);
                const uint32_t sX = alternate ? sRC + SLAB : sBase + 2 * SLAB;
                // earlier bulk stores must have finished READING the slabs before they are refilled.  With tile operands to load
                // (residual / aux) that is now; otherwise only right before the first st.shared, after the TMEM load latency.
                const bool need_drain = live && ep.tma_io && stores_pending;
                auto drain = [&]() {
                    if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                };
                if (need_drain && loads) drain();
                if (loads && elect_one()) {
                    const bool two = nb + 32 < ep.N;
                    uint32_t bytes = 0;
                    if (has_r && lead) bytes += two ? 2 * SLAB : SLAB;
                    if (has_xi) bytes += SLAB;
                    mbar_expect_tx(ebar, bytes);
                    if (has_r && lead) {
                        tma_load_4d(sRC, &tmR, ebar, nb, mrow, b2 * ep.r_m2, b1 * ep.r_m1);
                        if (two) tma_load_4d(sRC + SLAB, &tmR, ebar, nb + 32, mrow, b2 * ep.r_m2, b1 * ep.r_m1);
                    }
                    if (has_xi) tma_load_4d(sX, &tmXi, ebar, nb, mrow, 0, 0);
                }
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * BN + (uint32_t)(sc * 64);
                const bool last_sc = sc + 1 >= BN / 64;
                if (!live || !ep.tma_io) {
                    uint32_t r[64];
                    TMEM_LD_32x32b_X32(taddr, r);
                    TMEM_LD_32x32b_X32(taddr + 32, r + 32);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (last_sc) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) release_acc(buf);
                        arrived = true;
                    }
                    if (live) epilogue_direct(ep, r, mrow + lane, nb, b1, b2, lead);
                    continue;
                }
                if (loads) { mbar_wait(ebar, eph); eph ^= 1u; }
                const int emode = (nb + 64 <= ep.N && (ep.dbg & 6) == 0) ? (lead ? ep.mode : ep.mode_nl) : EM_GENERIC;
                if (emode != EM_GENERIC) {
                    uint32_t r[64];
                    TMEM_LD_32x32b_X32(taddr, r);
                    TMEM_LD_32x32b_X32(taddr + 32, r + 32);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (last_sc) {
                        // last TMEM read of this warp for this tile: hand the accumulator buffer back to the MMA warp
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) release_acc(buf);
                        arrived = true;
                    }
                    if (need_drain && !loads) drain();
                    switch (emode) {
                        case EM_F32: epi_fast<true, false, 0, false, false, false>(ep, r, nb, lane, sRC, sX); break;
                        case EM_BF16: epi_fast<false, false, 0, false, false, false>(ep, r, nb, lane, sRC, sX); break;
                        case EM_F16: epi_fast<false, false, 0, false, false, false, true>(ep, r, nb, lane, sRC, sX); break;
                        case EM_BF16_BIAS: epi_fast<false, true, 0, false, false, false>(ep, r, nb, lane, sRC, sX); break;
                        case EM_F32_BIAS: epi_fast<true, true, 0, false, false, false>(ep, r, nb, lane, sRC, sX); break;
                        case EM_F32_BIAS_GAMMA_RES_XO: epi_fast<true, true, 0, true, true, true>(ep, r, nb, lane, sRC, sX); break;
                        case EM_F32_BIAS_RES: epi_fast<true, true, 0, false, true, false>(ep, r, nb, lane, sRC, sX); break;
                        case EM_F32_RES: epi_fast<true, false, 0, false, true, false>(ep, r, nb, lane, sRC, sX); break;
                        case EM_BF16_BIAS_RES: epi_fast<false, true, 0, false, true, false>(ep, r, nb, lane, sRC, sX); break;
                        case EM_BF16_BIAS_GELU_XO: epi_fast<false, true, SPE_ACT_GELU, false, false, true>(ep, r, nb, lane, sRC, sX); break;
                        case EM_BF16_BIAS_RELU: epi_fast<false, true, SPE_ACT_RELU, false, false, false>(ep, r, nb, lane, sRC, sX); break;
                        case EM_BF16_RELUGRAD: epi_fast<false, false, SPE_ACT_RELU_GRAD, false, false, false>(ep, r, nb, lane, sRC, sX); break;
                        default: epi_fast<false, false, SPE_ACT_GELU_GRAD, false, false, false>(ep, r, nb, lane, sRC, sX); break;
                    }
                } else {
                // NOTE on code size: every condition below is warp-uniform and is tested once per 8-column group (never per
                // element); the N-tail takes the same code with clamped vector loads.  (An earlier per-element-branch version
                // compiled to ~5000 SASS instructions per super-chunk and made every GEMM issue-bound in its epilogue.)
                const bool use_bias = ep.bias != nullptr && lead, use_res = has_r && lead;
                const int act = ep.act;
                if (need_drain && !loads) drain();
#pragma unroll 1
                for (int gg = 0; gg < 4; ++gg) {                        // rolled: 4 x 16 accumulator columns (keeps the code in the I-cache)
                uint32_t r[16];
                if (!(ep.dbg & 4)) {
                    TMEM_LD_32x32b_X16(taddr + gg * 16, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = 0u;
                }
                if (gg == 3 && last_sc) {
                    // last TMEM read of this warp for this tile: hand the accumulator buffer back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) release_acc(buf);
                    arrived = true;
                }
#pragma unroll
                for (int g2 = 0; g2 < 2; ++g2) {                        // 2 groups of 8 columns
                    const int g = gg * 2 + g2;
                    const int n = nb + g * 8;
                    // columns >= N are clipped by the TMA stores.  Per-column vectors (bias / gamma) of a group that straddles
                    // or lies beyond N are fetched element-wise (warp-uniform rare path).
                    const bool gfull = n + 8 <= ep.N;
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g2 * 8 + j]) * ep.alpha;
                    if (use_bias) {
                        if (gfull) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + n)), b1v = __ldg(reinterpret_cast<const float4*>(ep.bias + n + 4));
                            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1v.x; v[5] += b1v.y; v[6] += b1v.z; v[7] += b1v.w;
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) if (n + j < ep.N) v[j] += __ldg(ep.bias + n + j);
                        }
                    }
                    if (has_xo) {
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sX + slab_off(lane, g)), "r"(pack_bf16x2(v[0], v[1])),
                                     "r"(pack_bf16x2(v[2], v[3])), "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7])) : "memory");
                    }
                    if (act != SPE_ACT_NONE) {
                        if (act == SPE_ACT_RELU) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                        } else if (act == SPE_ACT_GELU) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
                        } else {
                            uint4 a;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(sX + slab_off(lane, g)));
                            float ax[8];
                            float2 tt;
                            tt = unpack_bf16x2(a.x); ax[0] = tt.x; ax[1] = tt.y;
                            tt = unpack_bf16x2(a.y); ax[2] = tt.x; ax[3] = tt.y;
                            tt = unpack_bf16x2(a.z); ax[4] = tt.x; ax[5] = tt.y;
                            tt = unpack_bf16x2(a.w); ax[6] = tt.x; ax[7] = tt.y;
                            if (act == SPE_ACT_RELU_GRAD) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = ax[j] > 0.f ? v[j] : 0.f;
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] *= gelu_erf_grad(ax[j]);
                            }
                        }
                    }
                    if (ep.gamma) {
                        if (gfull) {
                            const float4 g0 = __ldg(reinterpret_cast<const float4*>(ep.gamma + n)), g1 = __ldg(reinterpret_cast<const float4*>(ep.gamma + n + 4));
                            v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w; v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) if (n + j < ep.N) v[j] *= __ldg(ep.gamma + n + j);
                        }
                    }
                    if (use_res) {
                        // fp32 residual: columns g*8..g*8+7 = slab (g/4), 16B chunks (g%4)*2, +1.  The C tile aliases these slabs:
                        // every thread reads its residual chunk for group g before it writes any C chunk of group g (see below).
                        const uint32_t sl = sRC + (g >> 2) * SLAB;
                        float4 r0, r1;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r0.x), "=f"(r0.y), "=f"(r0.z), "=f"(r0.w) : "r"(sl + slab_off(lane, (g & 3) * 2)));
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r1.x), "=f"(r1.y), "=f"(r1.z), "=f"(r1.w) : "r"(sl + slab_off(lane, (g & 3) * 2 + 1)));
                        v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
                    }
                    if (ep.dbg & 2) {
                        if (v[0] == 123.456f) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sRC), "f"(v[1] + v[2] + v[3] + v[4] + v[5] + v[6] + v[7]) : "memory");
                    } else if (c32) {
                        const uint32_t sl = sRC + (g >> 2) * SLAB;
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sl + slab_off(lane, (g & 3) * 2)), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sl + slab_off(lane, (g & 3) * 2 + 1)), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
                    } else {
                        // bf16 C chunk g overlays residual columns 4g..4g+3, which this thread consumed at group g/2 <= g
                        if (ep.c_dtype == SPE_DT_F16)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sRC + slab_off(lane, g)), "r"(pack_f16x2_rn(v[0], v[1])),
                                         "r"(pack_f16x2_rn(v[2], v[3])), "r"(pack_f16x2_rn(v[4], v[5])), "r"(pack_f16x2_rn(v[6], v[7])) : "memory");
                        else
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sRC + slab_off(lane, g)), "r"(pack_bf16x2(v[0], v[1])),
                                     "r"(pack_bf16x2(v[2], v[3])), "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7])) : "memory");
                    }
                }
                }
                }   // generic path
                if (!(ep.dbg & 8)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (!(ep.dbg & 1) && elect_one()) {
                    if (ep.splits > 1 || ep.accum) {
                        tma_reduce_add_4d(&tmC, sRC, nb, mrow, b2, b1);
                        if (nb + 32 < ep.N) tma_reduce_add_4d(&tmC, sRC + SLAB, nb + 32, mrow, b2, b1);
                    } else if (c32) {
                        tma_store_4d(&tmC, sRC, nb, mrow, b2, b1);
                        if (nb + 32 < ep.N) tma_store_4d(&tmC, sRC + SLAB, nb + 32, mrow, b2, b1);
                    } else {
                        tma_store_4d(&tmC, sRC, nb, mrow, b2, b1);
                    }
                    if (has_xo) tma_store_4d(&tmXo, sX, nb, mrow, 0, 0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                stores_pending = true;
                set ^= 1u;
            }
            if (!arrived) {
                // this warp owns no super-chunk of the tile (BN = 64, upper half): still release the buffer
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) release_acc(buf);
            }
        }
        __syncwarp();
        if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CG2) cluster_sync_all();             // the peer's smem / TMEM / barriers stay alive until the leader's last MMA retired
    if (warp == 2) {
        if constexpr (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

int make_tmap(CUtensorMap* tm, const void* ptr, int major, int rows, int K, int64_t ld, int64_t sb1, int64_t sb2, int batch1,
              int batch2, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    SPE_CHECK(enc, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    SPE_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "gemm operand pointer must be 16-byte aligned");
    SPE_CHECK(ld % 8 == 0, "gemm operand leading dimension (%lld) must be a multiple of 8 elements", (long long)ld);
    SPE_CHECK((batch1 == 1 || sb1 % 8 == 0) && (batch2 == 1 || sb2 % 8 == 0), "gemm batch strides must be multiples of 8 elements");
    cuuint64_t gdim[4];
    cuuint64_t gstr[3];
    cuuint32_t box[4];
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint64_t inner = (major == SPE_MAJOR_K) ? (cuuint64_t)K : (cuuint64_t)rows;
    const cuuint64_t outer = (major == SPE_MAJOR_K) ? (cuuint64_t)rows : (cuuint64_t)K;
    // a batch dim with stride 0 is a broadcast: describe it with extent 1 (the kernel passes coordinate 0)
    const bool bc2 = batch2 > 1 && sb2 == 0, bc1 = batch1 > 1 && sb1 == 0;
    gdim[0] = inner; gdim[1] = outer; gdim[2] = bc2 ? 1 : (cuuint64_t)batch2; gdim[3] = bc1 ? 1 : (cuuint64_t)batch1;
    gstr[0] = (cuuint64_t)ld * 2;
    gstr[1] = (batch2 > 1 && !bc2 ? (cuuint64_t)sb2 : (cuuint64_t)ld * outer) * 2;
    gstr[2] = (batch1 > 1 && !bc1 ? (cuuint64_t)sb1 : (cuuint64_t)ld * outer) * 2;
    box[0] = 64;
    box[1] = (major == SPE_MAJOR_K) ? (cuuint32_t)box_rows : 64;
    box[2] = 1; box[3] = 1;
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SPE_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu ld %lld", (int)r, (unsigned long long)inner,
              (unsigned long long)outer, (long long)ld);
    return 0;
}

// epilogue tile operand [M, N] (row pitch ld elements), box = 32 rows x 128 bytes, SWIZZLE_128B. Returns 1 if not TMA-able.
int make_tmap_io(CUtensorMap* tm, const void* ptr, bool f32, int M, int N, int64_t ld, int64_t sb1, int64_t sb2, int batch1, int batch2) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return 1;
    const int es = f32 ? 4 : 2;
    const int per16 = 16 / es;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || ld % per16 != 0) return 1;
    const bool bc2 = batch2 > 1 && sb2 == 0, bc1 = batch1 > 1 && sb1 == 0;
    if ((batch1 > 1 && !bc1 && sb1 % per16 != 0) || (batch2 > 1 && !bc2 && sb2 % per16 != 0)) return 1;
    cuuint64_t gdim[4] = {(cuuint64_t)N, (cuuint64_t)M, bc2 ? 1 : (cuuint64_t)batch2, bc1 ? 1 : (cuuint64_t)batch1};
    cuuint64_t gstr[3];
    gstr[0] = (cuuint64_t)ld * es;
    gstr[1] = (batch2 > 1 && !bc2 ? (cuuint64_t)sb2 : (cuuint64_t)ld * M) * es;
    gstr[2] = (batch1 > 1 && !bc1 ? (cuuint64_t)sb1 : (cuuint64_t)ld * M) * es;
    cuuint32_t box[4] = {(cuuint32_t)(128 / es), 32, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}

struct IoMaps { CUtensorMap C, R, Xi, Xo; };

// SPE_GEMM_PDL (default 1): launch GEMMs with programmatic stream serialization -- a GEMM that follows another GEMM in the stream
// starts its prologue (and absorbs its launch latency) under the predecessor's tail.  ~11 us of a 13-25 us launch were fixed costs.
bool gemm_pdl() {
    static const bool on = [] { const char* e = getenv("SPE_GEMM_PDL"); return e == nullptr || atoi(e) != 0; }();
    return on;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, bool CG2>
int launch(const CUtensorMap& tA, const CUtensorMap& tB, const IoMaps& io, const EpiParams& ep, cudaStream_t st) {
    constexpr size_t SMEM = (size_t)STAGES * (BM * BK * 2 + (CG2 ? BN / 2 : BN) * BK * 2) + EPI_SMEM + (2 * STAGES + 4 + NUM_EPI_WARPS) * 8 + 16 + 1024;
    static_assert(SMEM <= 232448, "shared memory budget (227 KB) exceeded");
    static bool attr_done = false;
    auto kfn = gemm_tcgen05_kernel<BN, STAGES, A_MN, B_MN, CG2>;
    if (!attr_done) {
        SPE_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        attr_done = true;
    }
    if constexpr (CG2) {
        // one cluster of 2 CTAs (= one TPC's SM pair) per 256 x BN tile stream
        const int pairs = spe_num_sms() / 2;
        const int clusters = ep.num_tiles < pairs ? ep.num_tiles : pairs;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = gemm_pdl() ? 2 : 1;
        SPE_CUDA(cudaLaunchKernelEx(&cfg, kfn, tA, tB, io.C, io.R, io.Xi, io.Xo, ep));
    } else {
        const int grid = ep.num_tiles < spe_num_sms() ? ep.num_tiles : spe_num_sms();
        if (gemm_pdl()) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(NUM_THREADS);
            cfg.dynamicSmemBytes = SMEM;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            SPE_CUDA(cudaLaunchKernelEx(&cfg, kfn, tA, tB, io.C, io.R, io.Xi, io.Xo, ep));
        } else {
            kfn<<<grid, NUM_THREADS, SMEM, st>>>(tA, tB, io.C, io.R, io.Xi, io.Xo, ep);
        }
    }
    SPE_LAUNCHED();
    return 0;
}

template <int BN, int STAGES, bool CG2 = false>
int dispatch_major(int am, int bm, const CUtensorMap& tA, const CUtensorMap& tB, const IoMaps& io, const EpiParams& ep, cudaStream_t st) {
    if (am == SPE_MAJOR_K && bm == SPE_MAJOR_K) return launch<BN, STAGES, false, false, CG2>(tA, tB, io, ep, st);
    if (am == SPE_MAJOR_K && bm == SPE_MAJOR_MN) return launch<BN, STAGES, false, true, CG2>(tA, tB, io, ep, st);
    if (am == SPE_MAJOR_MN && bm == SPE_MAJOR_K) return launch<BN, STAGES, true, false, CG2>(tA, tB, io, ep, st);
    return launch<BN, STAGES, true, true, CG2>(tA, tB, io, ep, st);
}

}  // namespace

// tensor-map helpers shared with the fused attention kernels (attn_fused.cu)
int spe_make_tmap_bf16(CUtensorMap* tm, const void* ptr, int major, int rows, int K, int64_t ld, int64_t sb1, int64_t sb2, int batch1, int batch2, int box_rows) {
    return make_tmap(tm, ptr, major, rows, K, ld, sb1, sb2, batch1, batch2, box_rows);
}
void* spe_tmap_encode_fn() { return reinterpret_cast<void*>(get_encode()); }

// n / d == (umulhi(n, mul) + n) >> shr for every n < 2^31 (d >= 1):  shr = ceil(log2 d), mul = floor(2^32 (2^shr - d) / d) + 1
static void fastdiv_magic(uint32_t d, uint32_t (&out)[2]) {
    uint32_t shr = 0;
    while ((1ull << shr) < d) ++shr;
    out[0] = (uint32_t)((((1ull << shr) - d) << 32) / d + 1);
    out[1] = shr;
}

// debugging / experiment switches, read once (getenv scans the whole environment: ~0.5 us each, five per call adds up at 1300 GEMMs a step)
struct GemmEnv {
    bool bn256, direct_epilogue, generic_epilogue, no_splitk;
    int dbg, cg2;
    GemmEnv() {
        const char* c2 = getenv("SPE_GEMM_CG2");
        // CTA-pair tiles: 0 (default) never, 1 heuristic (M >= 1024, full 256-row tiles), 2 wherever legal.  Measured on cfg2 (round 1):
        // parity-green, but no gain yet (52.0 ms/step off vs 53.6 on): at K = 384 a launch is dominated by its ramp / barrier skeleton
        // and the epilogue, not by operand traffic (skipping ALL operand loads changes no GEMM time by more than 5%) -- see DESIGN.md.
        cg2 = c2 ? atoi(c2) : 0;
        bn256 = getenv("SPE_GEMM_BN256") != nullptr;
        direct_epilogue = getenv("SPE_GEMM_DIRECT_EPILOGUE") != nullptr;
        generic_epilogue = getenv("SPE_GEMM_GENERIC_EPILOGUE") != nullptr;
        no_splitk = getenv("SPE_GEMM_NO_SPLITK") != nullptr;
        const char* d = getenv("SPE_GEMM_DBG");
        dbg = d ? atoi(d) : 0;
    }
};

extern "C" __attribute__((visibility("default"))) int spe_gemm(const spe_gemm_args* a, void* stream) {
    static const GemmEnv env;
    SPE_CHECK(a && a->A && a->B && a->C, "spe_gemm: null argument");
    SPE_CHECK(a->M > 0 && a->N > 0 && a->K > 0 && a->batch1 > 0 && a->batch2 > 0, "spe_gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
    SPE_CHECK(a->act == SPE_ACT_NONE || a->act == SPE_ACT_RELU || a->act == SPE_ACT_GELU || a->aux_in, "spe_gemm: *_GRAD activation needs aux_in");
    // 128x256 tiles (fewer A re-reads, 2 pipeline stages) were measured: S-type GEMMs -6%, fc1 +15% -> opt-in only (SPE_GEMM_BN256=1)
    const bool wideN = a->N >= 512 && (((a->N + 255) / 256) * 256 - a->N) * 8 <= a->N && env.bn256;
    const int batch = a->batch1 * a->batch2;
    // CTA-pair (cta_group::2) 256 x BN tiles: dense GEMMs with enough rows that the 256-row tiles fill (see the kernel comment)
    const bool pair_ok = batch == 1 && a->N > 64 && a->M >= 512;
    const bool pair = pair_ok && (env.cg2 >= 2 || (env.cg2 == 1 && a->M >= 1024 && (a->M % 256 == 0 || a->M >= 4096)));
    const int BN = a->N <= 64 ? 64 : (pair ? (a->N >= 1024 ? 256 : 128) : (wideN ? 256 : 128));
    SPE_CHECK(batch == 1 || !(a->aux_in || a->aux_out), "spe_gemm: aux_in / aux_out are not batched");
    SPE_CHECK(a->c_dtype != SPE_DT_F16 || (!a->bias && !a->gamma && !a->residual && !a->aux_in && !a->aux_out && a->act == SPE_ACT_NONE && a->split == 0),
              "spe_gemm: fp16 output supports the plain alpha-scaled product only");
    CUtensorMap tA, tB;
    if (make_tmap(&tA, a->A, a->a_major, a->M, a->K, a->lda, a->a_sb1, a->a_sb2, a->batch1, a->batch2, BM)) return -1;
    if (make_tmap(&tB, a->B, a->b_major, a->N, a->K, a->ldb, a->b_sb1, a->b_sb2, a->batch1, a->batch2, pair ? BN / 2 : BN)) return -1;

    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.C = a->C; ep.c_dtype = a->c_dtype; ep.ldc = a->ldc; ep.c_sb1 = a->c_sb1; ep.c_sb2 = a->c_sb2;
    ep.alpha = a->alpha; ep.bias = a->bias; ep.act = a->act;
    ep.aux_in = reinterpret_cast<const uint16_t*>(a->aux_in); ep.aux_out = reinterpret_cast<uint16_t*>(a->aux_out); ep.ld_aux = a->ld_aux;
    ep.gamma = a->gamma; ep.residual = a->residual; ep.ldr = a->ldr; ep.r_sb1 = a->r_sb1; ep.r_sb2 = a->r_sb2;
    // residual aliasing C (same pointer and pitch) = in-place accumulation C += alpha A B [+ bias]: no residual tile is loaded,
    // every tile (and every K split) is reduce-added into C by the TMA store engine.  This is how wgrad accumulates into .grad.
    const bool accum = a->residual && a->residual == (const float*)a->C && a->c_dtype == SPE_DT_F32 && a->ldr == a->ldc &&
                       (batch == 1 || (a->r_sb1 == a->c_sb1 && a->r_sb2 == a->c_sb2)) && !a->gamma && a->act == SPE_ACT_NONE;
    if (accum) { ep.residual = nullptr; ep.accum = 1; }
    const float* residual = accum ? nullptr : a->residual;
    ep.split = a->split; ep.split_stride = a->split_stride;
    ep.M = a->M; ep.N = a->N; ep.K = a->K; ep.batch2 = a->batch2;
    ep.a_m1 = (a->batch1 > 1 && a->a_sb1 == 0) ? 0 : 1; ep.a_m2 = (a->batch2 > 1 && a->a_sb2 == 0) ? 0 : 1;
    ep.b_m1 = (a->batch1 > 1 && a->b_sb1 == 0) ? 0 : 1; ep.b_m2 = (a->batch2 > 1 && a->b_sb2 == 0) ? 0 : 1;
    ep.r_m1 = (a->batch1 > 1 && a->r_sb1 == 0) ? 0 : 1; ep.r_m2 = (a->batch2 > 1 && a->r_sb2 == 0) ? 0 : 1;

    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // ---- epilogue tiles through TMA when every tile operand is 16-byte regular (else direct global access)
    IoMaps io;
    memset(&io, 0, sizeof(io));
    const bool cf32 = a->c_dtype == SPE_DT_F32;
    const bool bias_ok = !a->bias || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0;
    const bool gamma_ok = !a->gamma || (reinterpret_cast<uintptr_t>(a->gamma) & 15) == 0;
    bool tma_io = a->split == 0 && !env.direct_epilogue && !(a->aux_in && a->aux_out) && bias_ok && gamma_ok;
    if (tma_io && make_tmap_io(&io.C, a->C, cf32, a->M, a->N, a->ldc, a->c_sb1, a->c_sb2, a->batch1, a->batch2)) tma_io = false;
    if (tma_io && residual && make_tmap_io(&io.R, residual, true, a->M, a->N, a->ldr, a->r_sb1, a->r_sb2, a->batch1, a->batch2)) tma_io = false;
    if (tma_io && a->aux_in && make_tmap_io(&io.Xi, a->aux_in, false, a->M, a->N, a->ld_aux, 0, 0, 1, 1)) tma_io = false;
    if (tma_io && a->aux_out && make_tmap_io(&io.Xo, a->aux_out, false, a->M, a->N, a->ld_aux, 0, 0, 1, 1)) tma_io = false;
    ep.tma_io = tma_io ? 1 : 0;
    ep.dbg = env.dbg;
    {
        const bool bi = a->bias != nullptr, ga = a->gamma != nullptr, re = residual != nullptr, xo = a->aux_out != nullptr, xi = a->aux_in != nullptr;
        int m = EM_GENERIC;
        if (tma_io && !env.generic_epilogue) {
            if (a->act == SPE_ACT_NONE && !xi) {
                if (cf32 && !bi && !ga && !re && !xo) m = EM_F32;
                else if (a->c_dtype == SPE_DT_F16 && !bi && !ga && !re && !xo) m = EM_F16;
                else if (!cf32 && !bi && !ga && !re && !xo) m = EM_BF16;
                else if (!cf32 && bi && !ga && !re && !xo) m = EM_BF16_BIAS;
                else if (cf32 && bi && !ga && !re && !xo) m = EM_F32_BIAS;
                else if (cf32 && bi && ga && re && xo) m = EM_F32_BIAS_GAMMA_RES_XO;
                else if (cf32 && bi && !ga && re && !xo) m = EM_F32_BIAS_RES;
                else if (cf32 && !bi && !ga && re && !xo) m = EM_F32_RES;
                else if (!cf32 && bi && !ga && re && !xo) m = EM_BF16_BIAS_RES;
            } else if (a->act == SPE_ACT_GELU && !cf32 && bi && !ga && !re && xo && !xi) m = EM_BF16_BIAS_GELU_XO;
            else if (a->act == SPE_ACT_RELU && !cf32 && bi && !ga && !re && !xo && !xi) m = EM_BF16_BIAS_RELU;
            else if (a->act == SPE_ACT_RELU_GRAD && !cf32 && !bi && !ga && !re && !xo && xi) m = EM_BF16_RELUGRAD;
            else if (a->act == SPE_ACT_GELU_GRAD && !cf32 && !bi && !ga && !re && !xo && xi) m = EM_BF16_GELUGRAD;
        }
        ep.mode = m;
        ep.mode_nl = (m != EM_GENERIC && cf32) ? EM_F32 : EM_GENERIC;      // split-K partial tiles: plain accumulate
    }
    // ---- split-K: few output tiles but a long reduction (wgrad).  fp32 contiguous C, no activation / aux / gamma.
    const int total_kb = (a->K + BK - 1) / BK;
    ep.tiles_m = pair ? (a->M + 2 * BM - 1) / (2 * BM) : (a->M + BM - 1) / BM;
    ep.tiles_n = (a->N + BN - 1) / BN;
    int splits = 1;
    const long long tiles = (long long)ep.tiles_m * ep.tiles_n * batch;
    SPE_CHECK(tiles < (1LL << 30), "spe_gemm: too many tiles");
    if (cf32 && batch == 1 && a->act == SPE_ACT_NONE && !a->aux_in && !a->aux_out && !a->gamma && (accum || a->residual != (const float*)a->C) &&
        a->ldc == a->N && a->split == 0 && tiles * 2 <= spe_num_sms() && total_kb >= 8 && !env.no_splitk) {
        splits = (int)((spe_num_sms() + tiles - 1) / tiles);
        if (splits > total_kb / 4) splits = total_kb / 4;
        if (splits < 1) splits = 1;
    }
    ep.kb_per_split = (total_kb + splits - 1) / splits;
    splits = (total_kb + ep.kb_per_split - 1) / ep.kb_per_split;      // no empty split
    ep.splits = splits;
    fastdiv_magic((uint32_t)ep.tiles_n, ep.fd_n); fastdiv_magic((uint32_t)ep.tiles_m, ep.fd_m);
    fastdiv_magic((uint32_t)splits, ep.fd_s); fastdiv_magic((uint32_t)a->batch2, ep.fd_b);
    ep.num_tiles = (int)(tiles * splits);
    if (splits > 1 && !accum) SPE_CUDA(cudaMemsetAsync(a->C, 0, (size_t)a->M * a->N * 4, st));
    char tag[64];
    if (g_spe_prof_on) snprintf(tag, sizeof(tag), "M%d N%d K%d b%d a%d b%d c%d", a->M, a->N, a->K, batch, a->a_major, a->b_major, a->c_dtype);
    // batched attention GEMMs (QK^T / PV and their gradients: K or N = head dim) are bound by their N^2 operand in HBM, not by the
    // tensor pipe: they are accounted as their own family with algorithmic BYTES; all other GEMMs with algorithmic FLOPs.
    const bool attn = batch > 1 && (a->K <= 128 || a->N <= 128);
    const double elt_c = cf32 ? 4.0 : 2.0;
    const double attn_bytes = ((double)a->M * a->K + (double)a->N * a->K) * 2.0 * batch + (double)a->M * a->N * elt_c * batch;
    SpeProfScope prof(attn ? SPE_FAM_GEMM_ATTN : SPE_FAM_GEMM, attn ? attn_bytes : 2.0 * a->M * a->N * (double)a->K * batch, st, tag);
    if (pair) {
        if (BN == 256) return dispatch_major<256, 4, true>(a->a_major, a->b_major, tA, tB, io, ep, st);
        return dispatch_major<128, 4, true>(a->a_major, a->b_major, tA, tB, io, ep, st);
    }
    if (BN == 64) return dispatch_major<64, 4>(a->a_major, a->b_major, tA, tB, io, ep, st);
    if (BN == 256) return dispatch_major<256, 2>(a->a_major, a->b_major, tA, tB, io, ep, st);
    return dispatch_major<128, 4>(a->a_major, a->b_major, tA, tB, io, ep, st);
}
