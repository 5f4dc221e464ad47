// Fused multi-head attention forward for sm_100a:  O = softmax(scale (Q K^T [+ Q2 K2^T]) + key_padding_mask) V
//   (nn.MultiheadAttention core, transformer.py:280;  models/attention.py:345-378 incl. the conditional cross-attention's
//    [content | position] concat expressed as two accumulated QK^T products, transformer.py:408-414)
//
// One CTA per (128-query block, head, image); the logits never touch HBM:
//   warp 0      : TMA producer -- Q (once), then K / V 128-key blocks through mbarrier rings (SWIZZLE_128B tiles);
//   warp 1      : one thread issues tcgen05.mma.cta_group::1.kind::f16:  S = Q K^T (+ Q2 K2^T) into a double-buffered TMEM
//                 accumulator (2 x 128 columns), and O += P V (A = P from shared memory, B = V consumed MN-major straight from
//                 the packed [B, Lk, H*dv] activation) into 64 further TMEM columns;
//   warp 2      : TMEM allocator;
//   warps 4..11 : softmax, two groups of 128 threads (thread = query row: no cross-lane reductions) that alternate key blocks:
//                 tcgen05.ld of the S row, exp2 with the statistics, bf16 P into the swizzled K-major shared tile that the PV
//                 MMA (and the TMA store of P) read.  The groups run out of phase so the MUFU pipe always has work.
// Two sweeps over the keys: sweep 1 = row statistics (online max / sum on the S tiles), sweep 2 = exact normalised P -> PV.
// The second QK^T costs 1/8 of the MUFU-bound softmax time and buys: no accumulator rescaling, and P that is already final when
// it is produced -- it is written once (bf16, TMA bulk stores) for the backward GEMMs, which is the only N^2 HBM traffic left
// in the forward (was: S f32 write + read, P write + read).  lse2 = max + log2(sum) per row is saved for a fused backward.
#include "common.cuh"
#include <cuda.h>
#include <string.h>
#include <stdlib.h>

int spe_make_tmap_bf16(CUtensorMap* tm, const void* ptr, int major, int rows, int K, int64_t ld, int64_t sb1, int64_t sb2, int batch1, int batch2, int box_rows);
void* spe_tmap_encode_fn();

namespace {

constexpr int BQ = 128;                 // query rows per CTA (= TMEM lanes)
constexpr int BKV = 128;                // keys per block
constexpr int KS = 2;                   // K ring stages with two QK segments (K and K2 tiles per stage); 2 * KS with one segment
constexpr int KSMAX = 2 * KS;
constexpr int VS = 3;                   // V ring stages
constexpr uint32_t TILE_B = 128 * 64 * 2;      // one [128 rows x 64 cols] bf16 SWIZZLE_128B tile (Q, K, half of P)
constexpr uint32_t VBOX_B = 64 * 64 * 2;       // one MN-major V box: 64 keys x 64 (dv, zero filled beyond dv)
constexpr int AT_THREADS = 384;
constexpr float LOG2E_F = 1.4426950408889634f;

struct AttnParams {
    int Lq, Lk, H;
    int nkb;                 // key blocks
    int two;                 // second QK segment present
    int ksteps1, ksteps2;    // UMMA k-steps (16 elements) of the two segments
    int dv;                  // value head dim (multiple of 16, <= 64)
    float scale2;            // scale * log2(e)
    const uint8_t* mask;     // [B, Lk] or null
    uint16_t* out; long long out_ld, out_sb;     // [B, Lq, H*dv] bf16
    float* lse;              // [B, H, Lq] or null
    int store_p, ldP;
    int dbg;                 // timing experiments only (SPE_ATTN_DBG): 1 = no MUFU (exp2 replaced by a multiply)
};

__device__ __forceinline__ uint32_t a_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void a_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void a_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void a_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool a_mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void a_mbar_wait(uint32_t bar, uint32_t parity) {
    if (a_mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!a_mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void a_tma_load(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void a_tma_store(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ uint64_t a_umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;       // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void a_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void a_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// one lane of a CONVERGED warp issues (tcgen05.mma / TMA operands live in uniform registers: issued from inside an `if (lane == 0)`
// region the compiler cannot prove uniformity and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST waterfall, ~140 cycles per MMA)
__device__ __forceinline__ bool a_elect() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float a_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#define A_TMEM_LD32(taddr, r)                                                                                 \
    asm volatile(                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                            \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"            \
        : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]), "=r"((r)[7]),     \
          "=r"((r)[8]), "=r"((r)[9]), "=r"((r)[10]), "=r"((r)[11]), "=r"((r)[12]), "=r"((r)[13]), "=r"((r)[14]), "=r"((r)[15]), \
          "=r"((r)[16]), "=r"((r)[17]), "=r"((r)[18]), "=r"((r)[19]), "=r"((r)[20]), "=r"((r)[21]), "=r"((r)[22]), "=r"((r)[23]), \
          "=r"((r)[24]), "=r"((r)[25]), "=r"((r)[26]), "=r"((r)[27]), "=r"((r)[28]), "=r"((r)[29]), "=r"((r)[30]), "=r"((r)[31]) \
        : "r"(taddr))
#define A_TMEM_LD16(taddr, r)                                                                                 \
    asm volatile(                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                            \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                     \
        : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]), "=r"((r)[7]),     \
          "=r"((r)[8]), "=r"((r)[9]), "=r"((r)[10]), "=r"((r)[11]), "=r"((r)[12]), "=r"((r)[13]), "=r"((r)[14]), "=r"((r)[15]) \
        : "r"(taddr))

// smem layout (all tiles 1024-byte aligned):  Q | Q2 | K[KS] | K2[KS] | V[VS] (2 boxes each) | P[2] (2 tiles each) | barriers | mask bits
constexpr uint32_t OFF_Q = 0, OFF_Q2 = OFF_Q + TILE_B, OFF_K = OFF_Q2 + TILE_B, OFF_K2 = OFF_K + KS * TILE_B, OFF_V = OFF_K2 + KS * TILE_B,
                   OFF_P = OFF_V + VS * 2 * VBOX_B, OFF_BAR = OFF_P + 2 * 2 * TILE_B;
constexpr int NBAR = 1 + 2 * KSMAX + 2 * VS + 4 + 4 + 1;      // qfull, kfull/kempty, vfull/vempty, sfull/sempty[2], pfull/pempty[2], ofull
constexpr uint32_t OFF_TSLOT = OFF_BAR + NBAR * 8, OFF_XCHG = OFF_TSLOT + 16, OFF_MBITS = OFF_XCHG + 2 * 128 * 8;

__global__ void __launch_bounds__(AT_THREADS, 1) attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmQ2,
                                                                 const __grid_constant__ CUtensorMap tmK2, const __grid_constant__ CUtensorMap tmP,
                                                                 const AttnParams ap) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = a_smem_u32(smem);
    const uint32_t bar0 = sbase + OFF_BAR;
    const uint32_t qfull = bar0, kfull0 = qfull + 8, kempty0 = kfull0 + 8 * KSMAX, vfull0 = kempty0 + 8 * KSMAX, vempty0 = vfull0 + 8 * VS,
                   sfull0 = vempty0 + 8 * VS, sempty0 = sfull0 + 16, pfull0 = sempty0 + 16, pempty0 = pfull0 + 16, ofull = pempty0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TSLOT);
    uint32_t* mbits = reinterpret_cast<uint32_t*>(smem + OFF_MBITS);          // [nkb][4]: bit set = key masked (padding or >= Lk)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    const int nkb = ap.nkb;
    // K ring: with one QK segment the K2 tiles are extra K stages (the two regions are contiguous)
    const uint32_t nks = ap.two ? (uint32_t)KS : (uint32_t)KSMAX;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
        a_mbar_init(qfull, 1);
        for (int s = 0; s < KSMAX; ++s) { a_mbar_init(kfull0 + 8 * s, 1); a_mbar_init(kempty0 + 8 * s, 1); }
        for (int s = 0; s < VS; ++s) { a_mbar_init(vfull0 + 8 * s, 1); a_mbar_init(vempty0 + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) {
            a_mbar_init(sfull0 + 8 * s, 1); a_mbar_init(sempty0 + 8 * s, 4);
            a_mbar_init(pfull0 + 8 * s, 1); a_mbar_init(pempty0 + 8 * s, 1);
        }
        a_mbar_init(ofull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // key mask bits: one ballot per 32 keys
    for (int w = warp; w < nkb * 4; w += AT_THREADS / 32) {
        const int j = w * 32 + lane;
        const bool masked = j >= ap.Lk || (ap.mask != nullptr && ap.mask[(long long)b * ap.Lk + j] != 0);
        const uint32_t bits = __ballot_sync(0xffffffffu, masked);
        if (lane == 0) mbits[w] = bits;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tO = tmem_base + 256;

    if (warp == 0) {
        // ---------------- TMA producer: the whole warp walks the loop, one elected lane issues ----------------
        if (a_elect()) {
            a_mbar_expect_tx(qfull, ap.two ? 2 * TILE_B : TILE_B);
            a_tma_load(sbase + OFF_Q, &tmQ, qfull, 0, q0, h, b);
            if (ap.two) a_tma_load(sbase + OFF_Q2, &tmQ2, qfull, 0, q0, h, b);
        }
        __syncwarp();
        uint32_t kit = 0;
        for (int sweep = 0; sweep < 2; ++sweep) {
            for (int j = 0; j < nkb; ++j, ++kit) {
                const uint32_t ks = kit % nks;
                a_mbar_wait(kempty0 + 8 * ks, ((kit / nks) & 1u) ^ 1u);
                if (a_elect()) {
                    a_mbar_expect_tx(kfull0 + 8 * ks, ap.two ? 2 * TILE_B : TILE_B);
                    a_tma_load(sbase + OFF_K + ks * TILE_B, &tmK, kfull0 + 8 * ks, 0, j * BKV, h, b);
                    if (ap.two) a_tma_load(sbase + OFF_K2 + ks * TILE_B, &tmK2, kfull0 + 8 * ks, 0, j * BKV, h, b);
                }
                __syncwarp();
                if (sweep == 1) {
                    const int vs = j % VS;
                    a_mbar_wait(vempty0 + 8 * vs, ((j / VS) & 1u) ^ 1u);
                    if (a_elect()) {
                        a_mbar_expect_tx(vfull0 + 8 * vs, 2 * VBOX_B);
                        // V tile MN-major: two boxes of [64 keys x 64 (dv, zero filled beyond dv)]
                        a_tma_load(sbase + OFF_V + vs * 2 * VBOX_B, &tmV, vfull0 + 8 * vs, 0, j * BKV, h, b);
                        a_tma_load(sbase + OFF_V + vs * 2 * VBOX_B + VBOX_B, &tmV, vfull0 + 8 * vs, 0, j * BKV + 64, h, b);
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        {
            // ---------------- MMA issuer: the whole warp walks the loop, one elected lane issues ----------------
            // instruction descriptors: D f32, A/B bf16; S: both K-major, N = 128;  PV: A K-major, B MN-major, N = dv
            const uint32_t IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BKV >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
            const uint32_t IDESC_PV = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(ap.dv >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
            a_mbar_wait(qfull, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t kit = 0, sit = 0;
            auto issue_s = [&]() {
                const uint32_t ks = kit % nks;
                const uint32_t sb = sit & 1u;
                a_mbar_wait(kfull0 + 8 * ks, (kit / nks) & 1u);
                a_mbar_wait(sempty0 + 8 * sb, ((sit >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (a_elect()) {
                    const uint32_t qa = sbase + OFF_Q, ka = sbase + OFF_K + ks * TILE_B;
                    for (int k = 0; k < ap.ksteps1; ++k) a_umma(tS + sb * 128, a_umma_desc(qa + k * 32, 0, 1024), a_umma_desc(ka + k * 32, 0, 1024), IDESC_S, k != 0);
                    if (ap.two) {
                        const uint32_t qa2 = sbase + OFF_Q2, ka2 = sbase + OFF_K2 + ks * TILE_B;
                        for (int k = 0; k < ap.ksteps2; ++k) a_umma(tS + sb * 128, a_umma_desc(qa2 + k * 32, 0, 1024), a_umma_desc(ka2 + k * 32, 0, 1024), IDESC_S, 1u);
                    }
                    a_commit(kempty0 + 8 * ks);
                    a_commit(sfull0 + 8 * sb);
                }
                __syncwarp();
                ++kit; ++sit;
            };
            for (int j = 0; j < nkb; ++j) issue_s();                  // sweep 1
            // sweep 2: the S tiles run TWO blocks ahead of the PV products (a softmax group finds its next tile ready when it
            // finishes a block; S_{j+2} reuses the TMEM buffer of S_j, which is free as soon as its group has loaded it)
            issue_s();
            if (nkb > 1) issue_s();
            for (int j = 0; j < nkb; ++j) {
                if (j + 2 < nkb) issue_s();
                const uint32_t pb = (uint32_t)j & 1u;
                const int vs = j % VS;
                a_mbar_wait(pfull0 + 8 * pb, ((uint32_t)j >> 1) & 1u);
                a_mbar_wait(vfull0 + 8 * vs, ((uint32_t)j / VS) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (a_elect()) {
                    const uint32_t pa = sbase + OFF_P + pb * 2 * TILE_B, va = sbase + OFF_V + vs * 2 * VBOX_B;
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k) {
                        // P: K-major, 64-key halves (k / 4), 32 bytes per k-step inside the 128-byte swizzle row
                        // V: MN-major, 64-key halves, 16 key rows = 2 groups of 8 rows (SBO 1024 B) per k-step
                        const uint64_t ad = a_umma_desc(pa + (k >> 2) * TILE_B + (k & 3) * 32, 0, 1024);
                        const uint64_t bd = a_umma_desc(va + (k >> 2) * VBOX_B + (k & 3) * 2048, VBOX_B, 1024);
                        a_umma(tO, ad, bd, IDESC_PV, (j | k) != 0 ? 1u : 0u);
                    }
                    a_commit(vempty0 + 8 * vs);
                    a_commit(pempty0 + 8 * pb);
                }
                __syncwarp();
            }
            if (a_elect()) a_commit(ofull);
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ---------------- softmax warps: two groups of 128 threads, thread = query row, groups alternate key blocks ----------------
        // Group g (warps 4+4g .. 7+4g; warp % 4 = TMEM lane quarter, a hardware rule) owns the key blocks j with j % 2 == g and P
        // buffer g.  The groups run out of phase: while one sits in its TMEM-load / barrier / store phase the other keeps the MUFU
        // pipe busy with exp2, which is what bounds this kernel (128 x 128 exponentials per block at 16 per clock per SM).
        const bool nomufu = (ap.dbg & 1) != 0;
#define EX(x) (nomufu ? (x) * 0.001f : a_ex2(x))
        const int quarter = warp & 3, grp = (warp - 4) >> 2;
        const int row = quarter * 32 + lane;                   // row inside the block = TMEM lane
        const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
        const bool gtid0 = (threadIdx.x & 127) == 0;           // first thread of the group: issues the bulk stores of P
        const uint32_t bar_id = 1u + (uint32_t)grp;
        float2* xchg = reinterpret_cast<float2*>(smem + OFF_XCHG);      // [2][128] (max, sum) of the two groups' blocks of a row
        float m = -INFINITY, l = 0.f;
        // ---- sweep 1: statistics over this group's key blocks
        for (int j = grp; j < nkb; j += 2) {
            const uint32_t sit = (uint32_t)j, sb = sit & 1u;
            a_mbar_wait(sfull0 + 8 * sb, (sit >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t r[64];
                if (!(ap.dbg & 2)) {
                    A_TMEM_LD32(tS + tlane + sb * 128 + hh * 64, r);
                    A_TMEM_LD32(tS + tlane + sb * 128 + hh * 64 + 32, r + 32);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int i = 0; i < 64; ++i) r[i] = (uint32_t)(i + lane);
                }
                if (hh == 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) a_mbar_arrive(sempty0 + 8 * sb);
                }
                const uint32_t bits0 = mbits[j * 4 + hh * 2], bits1 = mbits[j * 4 + hh * 2 + 1];
                if ((bits0 | bits1) == 0u) {                  // warp-uniform fast path: no masked key in these 64 columns
                    float cmx = __uint_as_float(r[0]);
#pragma unroll
                    for (int i = 1; i < 64; ++i) cmx = fmaxf(cmx, __uint_as_float(r[i]));
                    const float mn = fmaxf(m, cmx * ap.scale2);
                    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                    for (int i = 0; i < 64; i += 2) {
                        acc0 += EX(fmaf(__uint_as_float(r[i]), ap.scale2, -mn));
                        acc1 += EX(fmaf(__uint_as_float(r[i + 1]), ap.scale2, -mn));
                    }
                    l = l * a_ex2(m - mn) + (acc0 + acc1);
                    m = mn;
                } else {
                    float cmx = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 64; ++i) {
                        const uint32_t bits = i < 32 ? bits0 : bits1;
                        if (!((bits >> (i & 31)) & 1u)) cmx = fmaxf(cmx, __uint_as_float(r[i]) * ap.scale2);
                    }
                    const float mn = fmaxf(m, cmx);
                    if (mn > -INFINITY) {
                        float acc = 0.f;
#pragma unroll
                        for (int i = 0; i < 64; ++i) {
                            const uint32_t bits = i < 32 ? bits0 : bits1;
                            acc += ((bits >> (i & 31)) & 1u) ? 0.f : a_ex2(fmaf(__uint_as_float(r[i]), ap.scale2, -mn));
                        }
                        l = l * a_ex2(m - mn) + acc;          // m = -inf: l = 0 and 2^-inf = 0
                        m = mn;
                    }
                }
            }
        }
        // combine the two groups' partial statistics of the row
        xchg[grp * 128 + row] = make_float2(m, l);
        asm volatile("bar.sync 3, 256;" ::: "memory");
        {
            const float2 o = xchg[(grp ^ 1) * 128 + row];
            const float M = fmaxf(m, o.x);
            l = (m == -INFINITY ? 0.f : l * a_ex2(m - M)) + (o.x == -INFINITY ? 0.f : o.y * a_ex2(o.x - M));
            m = M;
        }
        const float lse2 = m + log2f(l);                       // p = 2^(x - lse2)
        if (grp == 0 && ap.lse != nullptr && q0 + row < ap.Lq) ap.lse[((long long)b * ap.H + h) * ap.Lq + q0 + row] = lse2;
        // ---- sweep 2: P -> shared (-> HBM), PV by the MMA warp
        const uint32_t sw = (uint32_t)(row & 7);
        const uint32_t pbase = sbase + OFF_P + (uint32_t)grp * 2 * TILE_B;
        for (int j = grp; j < nkb; j += 2) {
            const uint32_t sit = (uint32_t)(nkb + j), sb = sit & 1u, use = (uint32_t)j >> 1;
            a_mbar_wait(sfull0 + 8 * sb, (sit >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // P buffer `grp` is free once the PV MMA of this group's previous block has retired and its bulk store has read it
            a_mbar_wait(pempty0 + 8 * grp, (use & 1u) ^ 1u);
            if (gtid0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t r[64];
                if (!(ap.dbg & 4)) {
                    A_TMEM_LD32(tS + tlane + sb * 128 + hh * 64, r);
                    A_TMEM_LD32(tS + tlane + sb * 128 + hh * 64 + 32, r + 32);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int i = 0; i < 64; ++i) r[i] = (uint32_t)(i + lane);
                }
                if (hh == 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) a_mbar_arrive(sempty0 + 8 * sb);
                }
                const uint32_t bits0 = mbits[j * 4 + hh * 2], bits1 = mbits[j * 4 + hh * 2 + 1];
                // 64 keys = one [128 x 64] half tile of P: row `row`, 8 chunks of 16 bytes (SWIZZLE_128B pattern)
                const uint32_t pa = pbase + (uint32_t)hh * TILE_B + (uint32_t)row * 128u;
                if ((bits0 | bits1) == 0u) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        uint32_t w4[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            w4[i] = pack_bf16x2(EX(fmaf(__uint_as_float(r[g * 8 + 2 * i]), ap.scale2, -lse2)), EX(fmaf(__uint_as_float(r[g * 8 + 2 * i + 1]), ap.scale2, -lse2)));
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pa + (((uint32_t)g ^ sw) << 4)), "r"(w4[0]), "r"(w4[1]), "r"(w4[2]), "r"(w4[3]) : "memory");
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        uint32_t w4[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int c0 = g * 8 + 2 * i, c1 = c0 + 1;
                            const uint32_t bits = c0 < 32 ? bits0 : bits1;
                            const float p0 = ((bits >> (c0 & 31)) & 1u) ? 0.f : a_ex2(fmaf(__uint_as_float(r[c0]), ap.scale2, -lse2));
                            const float p1 = ((bits >> (c1 & 31)) & 1u) ? 0.f : a_ex2(fmaf(__uint_as_float(r[c1]), ap.scale2, -lse2));
                            w4[i] = pack_bf16x2(p0, p1);
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pa + (((uint32_t)g ^ sw) << 4)), "r"(w4[0]), "r"(w4[1]), "r"(w4[2]), "r"(w4[3]) : "memory");
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (gtid0) {
                if (ap.store_p) {
                    a_tma_store(&tmP, pbase, j * BKV, q0, h, b);
                    if (j * BKV + 64 < ap.ldP) a_tma_store(&tmP, pbase + TILE_B, j * BKV + 64, q0, h, b);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                a_mbar_arrive(pfull0 + 8 * grp);
            }
        }
        // ---- output: group 0 stores columns [0, 32), group 1 [32, 64) of the row's dv values
        a_mbar_wait(ofull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            uint32_t o[32];
            A_TMEM_LD32(tO + tlane + (uint32_t)grp * 32u, o);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (q0 + row < ap.Lq) {
                uint16_t* dst = ap.out + (long long)b * ap.out_sb + (long long)(q0 + row) * ap.out_ld + (long long)h * ap.dv + grp * 32;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (grp * 32 + g * 8 < ap.dv) {
                        const uint4 v = make_uint4(pack_bf16x2(__uint_as_float(o[g * 8]), __uint_as_float(o[g * 8 + 1])),
                                                   pack_bf16x2(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3])),
                                                   pack_bf16x2(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5])),
                                                   pack_bf16x2(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7])));
                        *reinterpret_cast<uint4*>(dst + g * 8) = v;
                    }
                }
            }
        }
        if (gtid0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}


// ================================================================================================
// Attention backward GEMMs, fused:  dV = P^T dO,  dK = alpha dS^T Q,  dQ = alpha dS K   in ONE pass over the N^2 tensors dS and P
// (attention.py:345-372 backward; cait.py:379-389 backward with P := A, the post-mix probabilities).  The three separate batched
// GEMMs read dS twice and P once; here every [128 q x 128 k] tile of dS and P is staged once by TMA and consumed by three
// tcgen05 products straight from the same shared-memory tiles -- a K-major tile read as operand A of dS K is, byte for byte, the
// MN-major operand A of dS^T Q (only the UMMA descriptor differs), and likewise Q / dO / K serve as MN-major B operands.
//   CTA = (128-key block j, head, image); dK_j and dV_j accumulate in TMEM over the query blocks; the dQ_i partial of each (i, j)
//   goes TMEM -> registers -> red.global.add.f32 into an fp32 accumulation buffer (L2 resident), cast to bf16 afterwards.
//   warp 0: TMA producer (2 stages of dS, P, Q, dO tiles; K_j once), warp 1: MMA issuer, warp 2: TMEM allocator,
//   warps 4..7: dQ drain per query block + final dK / dV stores.
// Padding: P and dS are exactly zero in the columns >= Lk (softmax / talking kernels) and TMA zero-fills rows >= Lq: no masks.
// ================================================================================================
struct BwdParams {
    int Lq, Lk, H, nqb;
    int d, dv, do_v;
    int from_dp;                                    // the "dS" operand holds dP: dS = P o (dP - delta) is formed in shared memory first
    const float* delta;                             // f32 [B, H, Lq] = rowsum(dO o O), from_dp only
    float alpha;
    float* dq_acc; long long dq_ld, dq_sb;          // fp32 [B, Lq, H*d]
    uint16_t* dk; long long dk_ld, dk_sb;          // bf16 [B, Lk, H*d]
    uint16_t* dvp; long long dv_ld, dv_sb;         // bf16 [B, Lk, H*dv]
    int dbg;                                        // timing experiments (SPE_ATTN_DBG): 8 = skip the dQ red.add
};

constexpr int BW_STAGES = 2;
constexpr uint32_t BW_STAGE_B = 2 * 2 * TILE_B + 2 * TILE_B;                 // dS (2 slabs), P (2 slabs), Q, dO
constexpr uint32_t BW_OFF_K = 0, BW_OFF_ST = TILE_B, BW_OFF_BAR = BW_OFF_ST + BW_STAGES * BW_STAGE_B;
constexpr int BW_NBAR = 1 + 3 * BW_STAGES + 4 + 1;                           // kfull, full/empty/ready[S], dqfull/dqempty[2], accfull
constexpr int BW_THREADS = 384;

__global__ void __launch_bounds__(BW_THREADS, 1) attn_bwd_gemms_kernel(const __grid_constant__ CUtensorMap tmDS, const __grid_constant__ CUtensorMap tmP,
                                                                       const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                                       const __grid_constant__ CUtensorMap tmDO, const BwdParams bp) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = a_smem_u32(smem);
    const uint32_t bar0 = sbase + BW_OFF_BAR;
    const uint32_t kfull = bar0, full0 = kfull + 8, empty0 = full0 + 8 * BW_STAGES, ready0 = empty0 + 8 * BW_STAGES, dqfull0 = ready0 + 8 * BW_STAGES,
                   dqempty0 = dqfull0 + 16, accfull = dqempty0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BW_OFF_BAR + BW_NBAR * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * BKV, h = blockIdx.y, b = blockIdx.z;
    const int nqb = bp.nqb;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmDS)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmP)) : "memory");
        a_mbar_init(kfull, 1);
        for (int s = 0; s < BW_STAGES; ++s) { a_mbar_init(full0 + 8 * s, 1); a_mbar_init(empty0 + 8 * s, 1); a_mbar_init(ready0 + 8 * s, 4); }
        for (int s = 0; s < 2; ++s) { a_mbar_init(dqfull0 + 8 * s, 1); a_mbar_init(dqempty0 + 8 * s, 4); }
        a_mbar_init(accfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tDV = tmem_base, tDK = tmem_base + 64, tDQ = tmem_base + 128;        // dQ: 2 x 64 columns

    if (warp == 0) {
        if (a_elect()) {
            a_mbar_expect_tx(kfull, TILE_B);
            a_tma_load(sbase + BW_OFF_K, &tmK, kfull, 0, k0, h, b);
        }
        __syncwarp();
        for (int i = 0; i < nqb; ++i) {
            const int s = i % BW_STAGES;
            a_mbar_wait(empty0 + 8 * s, (((uint32_t)i / BW_STAGES) & 1u) ^ 1u);
            if (a_elect()) {
                const uint32_t st = sbase + BW_OFF_ST + s * BW_STAGE_B, fb = full0 + 8 * s;
                const bool need_p = bp.do_v || bp.from_dp;
                a_mbar_expect_tx(fb, 3 * TILE_B + (need_p ? 2 * TILE_B : 0u) + (bp.do_v ? TILE_B : 0u));
                const int q0 = i * BQ;
                a_tma_load(st, &tmDS, fb, k0, q0, h, b);
                a_tma_load(st + TILE_B, &tmDS, fb, k0 + 64, q0, h, b);
                a_tma_load(st + 4 * TILE_B, &tmQ, fb, 0, q0, h, b);
                if (need_p) {
                    a_tma_load(st + 2 * TILE_B, &tmP, fb, k0, q0, h, b);
                    a_tma_load(st + 3 * TILE_B, &tmP, fb, k0 + 64, q0, h, b);
                }
                if (bp.do_v) a_tma_load(st + 5 * TILE_B, &tmDO, fb, 0, q0, h, b);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        {
            // D f32, A/B bf16.  dV / dK: A MN-major (bit 15), B MN-major (bit 16);  dQ: A K-major, B MN-major.  M = 128, N = dv | d.
            const uint32_t ID_T = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BQ >> 4) << 24);
            const uint32_t ID_DV = ID_T | ((uint32_t)(bp.dv >> 3) << 17), ID_DK = ID_T | ((uint32_t)(bp.d >> 3) << 17);
            const uint32_t ID_DQ = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(bp.d >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
            a_mbar_wait(kfull, 0);
            const uint32_t ka = sbase + BW_OFF_K;
            for (int i = 0; i < nqb; ++i) {
                const int s = i % BW_STAGES;
                const uint32_t qb = (uint32_t)i & 1u;
                a_mbar_wait((bp.from_dp ? ready0 : full0) + 8 * s, ((uint32_t)i / BW_STAGES) & 1u);
                a_mbar_wait(dqempty0 + 8 * qb, (((uint32_t)i >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = sbase + BW_OFF_ST + s * BW_STAGE_B;
                const uint32_t dsa = st, pa = st + 2 * TILE_B, qa = st + 4 * TILE_B, doa = st + 5 * TILE_B;
                if (a_elect()) {
#pragma unroll
                for (int k = 0; k < BQ / 16; ++k) {                    // contraction over the 128 query rows of the block
                    // A = (dS | P)^T: MN-major, M = keys in two 64-key slabs (LBO = slab stride), 16 query rows per step (2 x SBO)
                    // B = Q | dO    : MN-major, N = head dim inside the 128-byte rows, 16 query rows per step
                    if (bp.do_v) a_umma(tDV, a_umma_desc(pa + k * 2048, TILE_B, 1024), a_umma_desc(doa + k * 2048, TILE_B, 1024), ID_DV, (i | k) != 0 ? 1u : 0u);
                    a_umma(tDK, a_umma_desc(dsa + k * 2048, TILE_B, 1024), a_umma_desc(qa + k * 2048, TILE_B, 1024), ID_DK, (i | k) != 0 ? 1u : 0u);
                }
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k) {                   // contraction over the 128 keys
                    // A = dS: K-major, 64-key slabs, 32 bytes per step;  B = K_j: MN-major, 16 key rows per step
                    a_umma(tDQ + qb * 64, a_umma_desc(dsa + (k >> 2) * TILE_B + (k & 3) * 32, 0, 1024), a_umma_desc(ka + k * 2048, TILE_B, 1024), ID_DQ,
                           k != 0 ? 1u : 0u);
                }
                a_commit(empty0 + 8 * s);
                a_commit(dqfull0 + 8 * qb);
                }
                __syncwarp();
            }
            if (a_elect()) a_commit(accfull);
        }
        __syncwarp();
    } else if (warp >= 8) {
        // ---------------- softmax backward in shared memory (from_dp): dS = P o (dP - delta), thread = query row ----------------
        if (bp.from_dp) {
            const int row = (warp - 8) * 32 + lane;
            const uint32_t sw = (uint32_t)(row & 7);
            for (int i = 0; i < nqb; ++i) {
                const int s = i % BW_STAGES;
                const int q = i * BQ + row;
                const float dl = q < bp.Lq ? bp.delta[((long long)b * bp.H + h) * bp.Lq + q] : 0.f;
                a_mbar_wait(full0 + 8 * s, ((uint32_t)i / BW_STAGES) & 1u);
                const uint32_t st = sbase + BW_OFF_ST + s * BW_STAGE_B + (uint32_t)row * 128u;
#pragma unroll
                for (int c = 0; c < 16; ++c) {                          // 16-byte chunks: slab c / 8, chunk-in-row c % 8 (same swizzle in both tiles)
                    const uint32_t off = (uint32_t)(c >> 3) * TILE_B + ((((uint32_t)c & 7u) ^ sw) << 4);
                    uint4 dp, pp;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(dp.x), "=r"(dp.y), "=r"(dp.z), "=r"(dp.w) : "r"(st + off));
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(pp.x), "=r"(pp.y), "=r"(pp.z), "=r"(pp.w) : "r"(st + 2 * TILE_B + off));
                    const uint32_t dpw[4] = {dp.x, dp.y, dp.z, dp.w}, ppw[4] = {pp.x, pp.y, pp.z, pp.w};
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 a = unpack_bf16x2(dpw[e]), p = unpack_bf16x2(ppw[e]);
                        o[e] = pack_bf16x2(p.x * (a.x - dl), p.y * (a.y - dl));         // P = 0 in the padding -> dS = 0 there
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + off), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) a_mbar_arrive(ready0 + 8 * s);
            }
        }
    } else if (warp >= 4) {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
        for (int i = 0; i < nqb; ++i) {
            const uint32_t qb = (uint32_t)i & 1u;
            a_mbar_wait(dqfull0 + 8 * qb, ((uint32_t)i >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[64];
            A_TMEM_LD32(tDQ + tlane + qb * 64, r);
            A_TMEM_LD32(tDQ + tlane + qb * 64 + 32, r + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) a_mbar_arrive(dqempty0 + 8 * qb);
            const int q = i * BQ + row;
            if (q < bp.Lq && !(bp.dbg & 8)) {
                float* dst = bp.dq_acc + (long long)b * bp.dq_sb + (long long)q * bp.dq_ld + (long long)h * bp.d;
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                    if (g * 4 < bp.d)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + g * 4), "f"(__uint_as_float(r[g * 4]) * bp.alpha),
                                     "f"(__uint_as_float(r[g * 4 + 1]) * bp.alpha), "f"(__uint_as_float(r[g * 4 + 2]) * bp.alpha),
                                     "f"(__uint_as_float(r[g * 4 + 3]) * bp.alpha) : "memory");
                }
            }
        }
        a_mbar_wait(accfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int key = k0 + row;
        {
            uint32_t r[64];
            A_TMEM_LD32(tDK + tlane, r);
            A_TMEM_LD32(tDK + tlane + 32, r + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (key < bp.Lk) {
                uint16_t* dst = bp.dk + (long long)b * bp.dk_sb + (long long)key * bp.dk_ld + (long long)h * bp.d;
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    if (g * 8 < bp.d)
                        *reinterpret_cast<uint4*>(dst + g * 8) =
                            make_uint4(pack_bf16x2(__uint_as_float(r[g * 8]) * bp.alpha, __uint_as_float(r[g * 8 + 1]) * bp.alpha),
                                       pack_bf16x2(__uint_as_float(r[g * 8 + 2]) * bp.alpha, __uint_as_float(r[g * 8 + 3]) * bp.alpha),
                                       pack_bf16x2(__uint_as_float(r[g * 8 + 4]) * bp.alpha, __uint_as_float(r[g * 8 + 5]) * bp.alpha),
                                       pack_bf16x2(__uint_as_float(r[g * 8 + 6]) * bp.alpha, __uint_as_float(r[g * 8 + 7]) * bp.alpha));
            }
        }
        if (bp.do_v) {
            uint32_t r[64];
            A_TMEM_LD32(tDV + tlane, r);
            A_TMEM_LD32(tDV + tlane + 32, r + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (key < bp.Lk) {
                uint16_t* dst = bp.dvp + (long long)b * bp.dv_sb + (long long)key * bp.dv_ld + (long long)h * bp.dv;
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    if (g * 8 < bp.dv)
                        *reinterpret_cast<uint4*>(dst + g * 8) =
                            make_uint4(pack_bf16x2(__uint_as_float(r[g * 8]), __uint_as_float(r[g * 8 + 1])), pack_bf16x2(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3])),
                                       pack_bf16x2(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5])), pack_bf16x2(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7])));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

// delta[b,h,q] = sum_c dO[b,q,h*dv+c] * O[b,q,h*dv+c]   (= sum_j P dP of the softmax backward); one warp per (b, q) token
__global__ void __launch_bounds__(256) attn_delta_kernel(const uint16_t* __restrict__ dO, const uint16_t* __restrict__ O, int B, int H, int Lq, int dv, long long do_ld,
                                                         long long do_sb, long long o_ld, long long o_sb, float* __restrict__ delta) {
    const int lane = threadIdx.x & 31;
    const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tok >= (long long)B * Lq) return;
    const int b = (int)(tok / Lq), q = (int)(tok % Lq);
    const uint16_t* a = dO + (long long)b * do_sb + (long long)q * do_ld;
    const uint16_t* o = O + (long long)b * o_sb + (long long)q * o_ld;
    for (int h = 0; h < H; ++h) {
        float s = 0.f;
        for (int c = lane; c < dv; c += 32) s += bf16_to_f(a[h * dv + c]) * bf16_to_f(o[h * dv + c]);
        s = warp_sum(s);
        if (lane == 0) delta[((long long)b * H + h) * Lq + q] = s;
    }
}

// fp32 accumulation buffer [rows, E] -> bf16 destination with its own row pitch (the q third of a packed dqkv, for example)
__global__ void __launch_bounds__(256) dq_cast_kernel(const float* __restrict__ src, long long rows, int E, uint16_t* __restrict__ dst, long long dst_ld, long long dst_sb,
                                                      long long rows_per_batch) {
    const long long n4 = rows * (E / 4);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / (E / 4);
        const int c = (int)(t % (E / 4)) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + r * E + c);
        const long long bb = r / rows_per_batch, rr = r % rows_per_batch;
        *reinterpret_cast<uint2*>(dst + bb * dst_sb + rr * dst_ld + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}


// ================================================================================================
// Fused attention BACKWARD (standard attention; recomputing flavour): nothing N^2 is read from or written to HBM.
//   per (128-key block j, head, image):  for every 128-query block i
//     S   = Qa_i Ka_j^T [+ Qb_i Kb_j^T]        (tcgen05, TMEM)        P  = 2^(scale2 S - lse_i), masked keys / rows -> 0
//     dP  = dO_i V_j^T                          (tcgen05, TMEM)        dS = P o (dP - delta_i)
//     P, dS -> bf16 swizzled shared tiles (8 transform warps, two threads per query row)
//     dV_j += P^T dO_i,  dK_j += dS^T Qa_i   (TMEM accumulators over i),   dQ_i partial = dS Ka_j -> red.global.add.f32
//   With two QK segments (conditional cross-attention) the kernel is launched twice with the segments swapped: each launch produces
//   the gradients of ITS segment `a` (the logits always use both); dV only in the first.
//   warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4..11 transform, 12..15 dQ drain + dK / dV stores.
// ================================================================================================
struct Bwd2Params {
    int Lq, Lk, H, nqb;
    int ksa, ksb, ksv;          // UMMA k-steps: segment a, segment b (0: none), V
    int d, dv, do_v;
    float scale2, alpha;
    const uint8_t* mask;
    const float* lse; const float* delta;           // f32 [B,H,Lq]
    float* dq_acc; long long dq_ld, dq_sb;
    uint16_t* dk; long long dk_ld, dk_sb;
    uint16_t* dvp; long long dv_ld, dv_sb;
    int dbg;
};
constexpr uint32_t B2_OFF_KA = 0, B2_OFF_KB = TILE_B, B2_OFF_V = 2 * TILE_B, B2_OFF_ST = 3 * TILE_B;      // stage: Qa, Qb, dO
constexpr uint32_t B2_STAGE_B = 3 * TILE_B;
constexpr uint32_t B2_OFF_P = B2_OFF_ST + 2 * B2_STAGE_B, B2_OFF_DS = B2_OFF_P + 2 * TILE_B, B2_OFF_BAR = B2_OFF_DS + 2 * TILE_B;
constexpr int B2_NBAR = 1 + 4 + 4 + 4 + 1;           // kfull, full/empty[2], spfull/spempty/pdready/pdempty, dqfull/dqempty[2], accfull
constexpr int B2_THREADS = 512;

__global__ void __launch_bounds__(B2_THREADS, 1) attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmKa,
                                                                 const __grid_constant__ CUtensorMap tmQb, const __grid_constant__ CUtensorMap tmKb,
                                                                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                                                                 const Bwd2Params bp) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = a_smem_u32(smem);
    const uint32_t bar0 = sbase + B2_OFF_BAR;
    const uint32_t kfull = bar0, full0 = kfull + 8, empty0 = full0 + 16, spfull = empty0 + 16, spempty = spfull + 8, pdready = spempty + 8,
                   pdempty = pdready + 8, dqfull0 = pdempty + 8, dqempty0 = dqfull0 + 16, accfull = dqempty0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + B2_OFF_BAR + B2_NBAR * 8);
    uint32_t* mbits = tmem_slot + 4;                  // 4 words: masked keys of this block
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * BKV, h = blockIdx.y, b = blockIdx.z;
    const int nqb = bp.nqb;
    const bool two = bp.ksb > 0;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQa)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmDO)) : "memory");
        a_mbar_init(kfull, 1);
        for (int s = 0; s < 2; ++s) { a_mbar_init(full0 + 8 * s, 1); a_mbar_init(empty0 + 8 * s, 1); a_mbar_init(dqfull0 + 8 * s, 1); a_mbar_init(dqempty0 + 8 * s, 4); }
        a_mbar_init(spfull, 1); a_mbar_init(spempty, 8); a_mbar_init(pdready, 8); a_mbar_init(pdempty, 1);
        a_mbar_init(accfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a_smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 3) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int j = k0 + w * 32 + lane;
            const bool masked = j >= bp.Lk || (bp.mask != nullptr && bp.mask[(long long)b * bp.Lk + j] != 0);
            const uint32_t bits = __ballot_sync(0xffffffffu, masked);
            if (lane == 0) mbits[w] = bits;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 320, tDQ = tmem_base + 384;     // dQ: 2 x 64

    if (warp == 0) {
        if (a_elect()) {
            a_mbar_expect_tx(kfull, (two ? 3u : 2u) * TILE_B);
            a_tma_load(sbase + B2_OFF_KA, &tmKa, kfull, 0, k0, h, b);
            if (two) a_tma_load(sbase + B2_OFF_KB, &tmKb, kfull, 0, k0, h, b);
            a_tma_load(sbase + B2_OFF_V, &tmV, kfull, 0, k0, h, b);
        }
        __syncwarp();
        for (int i = 0; i < nqb; ++i) {
            const int s = i & 1;
            a_mbar_wait(empty0 + 8 * s, (((uint32_t)i >> 1) & 1u) ^ 1u);
            if (a_elect()) {
                const uint32_t st = sbase + B2_OFF_ST + s * B2_STAGE_B, fb = full0 + 8 * s;
                a_mbar_expect_tx(fb, (two ? 3u : 2u) * TILE_B);
                a_tma_load(st, &tmQa, fb, 0, i * BQ, h, b);
                if (two) a_tma_load(st + TILE_B, &tmQb, fb, 0, i * BQ, h, b);
                a_tma_load(st + 2 * TILE_B, &tmDO, fb, 0, i * BQ, h, b);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        {
            const uint32_t ID_S = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BKV >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);            // K-major x K-major, N = 128
            const uint32_t ID_T = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BQ >> 4) << 24);                 // MN-major A and B
            const uint32_t ID_DV = ID_T | ((uint32_t)(bp.dv >> 3) << 17), ID_DK = ID_T | ((uint32_t)(bp.d >> 3) << 17);
            const uint32_t ID_DQ = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(bp.d >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
            const uint32_t ka = sbase + B2_OFF_KA, kb = sbase + B2_OFF_KB, va = sbase + B2_OFF_V, pa = sbase + B2_OFF_P, dsa = sbase + B2_OFF_DS;
            auto issue_sdp = [&](int i) {
                const uint32_t st = sbase + B2_OFF_ST + (uint32_t)(i & 1) * B2_STAGE_B;
                a_mbar_wait(full0 + 8 * (i & 1), ((uint32_t)i >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (a_elect()) {
                    for (int k = 0; k < bp.ksa; ++k) a_umma(tS, a_umma_desc(st + k * 32, 0, 1024), a_umma_desc(ka + k * 32, 0, 1024), ID_S, k != 0);
                    for (int k = 0; k < bp.ksb; ++k) a_umma(tS, a_umma_desc(st + TILE_B + k * 32, 0, 1024), a_umma_desc(kb + k * 32, 0, 1024), ID_S, 1u);
                    for (int k = 0; k < bp.ksv; ++k) a_umma(tDP, a_umma_desc(st + 2 * TILE_B + k * 32, 0, 1024), a_umma_desc(va + k * 32, 0, 1024), ID_S, k != 0);
                    a_commit(spfull);
                }
                __syncwarp();
            };
            a_mbar_wait(kfull, 0);
            issue_sdp(0);
            for (int i = 0; i < nqb; ++i) {
                const uint32_t qb = (uint32_t)i & 1u;
                if (i + 1 < nqb) {
                    a_mbar_wait(spempty, (uint32_t)i & 1u);                 // the transform warps have loaded S_i / dP_i out of TMEM
                    issue_sdp(i + 1);
                }
                a_mbar_wait(pdready, (uint32_t)i & 1u);                     // P_i / dS_i tiles are in shared memory
                a_mbar_wait(dqempty0 + 8 * qb, (((uint32_t)i >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = sbase + B2_OFF_ST + qb * B2_STAGE_B;
                const uint32_t qa = st, doa = st + 2 * TILE_B;
                if (a_elect()) {
#pragma unroll
                    for (int k = 0; k < BQ / 16; ++k) {
                        if (bp.do_v) a_umma(tDV, a_umma_desc(pa + k * 2048, TILE_B, 1024), a_umma_desc(doa + k * 2048, TILE_B, 1024), ID_DV, (i | k) != 0 ? 1u : 0u);
                        a_umma(tDK, a_umma_desc(dsa + k * 2048, TILE_B, 1024), a_umma_desc(qa + k * 2048, TILE_B, 1024), ID_DK, (i | k) != 0 ? 1u : 0u);
                    }
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k)
                        a_umma(tDQ + qb * 64, a_umma_desc(dsa + (k >> 2) * TILE_B + (k & 3) * 32, 0, 1024), a_umma_desc(ka + k * 2048, TILE_B, 1024), ID_DQ, k != 0 ? 1u : 0u);
                    a_commit(empty0 + 8 * qb);
                    a_commit(dqfull0 + 8 * qb);
                    a_commit(pdempty);
                }
                __syncwarp();
            }
            if (a_elect()) a_commit(accfull);
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 12) {
        // ---------------- transform warps: two threads per query row (64 keys each) ----------------
        const int quarter = warp & 3, half = (warp - 4) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t tl = ((uint32_t)(quarter * 32) << 16) + (uint32_t)half * 64u;
        const uint32_t sw = (uint32_t)(row & 7);
        const uint32_t bits0 = mbits[half * 2], bits1 = mbits[half * 2 + 1];
        const uint32_t prow = sbase + B2_OFF_P + (uint32_t)half * TILE_B + (uint32_t)row * 128u;
        const uint32_t drow = sbase + B2_OFF_DS + (uint32_t)half * TILE_B + (uint32_t)row * 128u;
        const long long stat0 = ((long long)b * bp.H + h) * bp.Lq;
        for (int i = 0; i < nqb; ++i) {
            const int q = i * BQ + row;
            // rows beyond Lq: lse = +inf -> P = 0 (their Q / dO rows are zero-filled by TMA)
            const float lse = q < bp.Lq ? bp.lse[stat0 + q] : INFINITY;
            const float dl = q < bp.Lq ? bp.delta[stat0 + q] : 0.f;
            a_mbar_wait(spfull, (uint32_t)i & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t pk[32], dk[32];                      // packed bf16 pairs of this thread's 64 P and dS values
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t rs[32], rd[32];
                A_TMEM_LD32(tS + tl + hh * 32, rs);
                A_TMEM_LD32(tDP + tl + hh * 32, rd);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const uint32_t bits = hh == 0 ? bits0 : bits1;
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const float p0 = ((bits >> c) & 1u) ? 0.f : a_ex2(fmaf(__uint_as_float(rs[c]), bp.scale2, -lse));
                    const float p1 = ((bits >> (c + 1)) & 1u) ? 0.f : a_ex2(fmaf(__uint_as_float(rs[c + 1]), bp.scale2, -lse));
                    pk[hh * 16 + (c >> 1)] = pack_bf16x2(p0, p1);
                    dk[hh * 16 + (c >> 1)] = pack_bf16x2(p0 * (__uint_as_float(rd[c]) - dl), p1 * (__uint_as_float(rd[c + 1]) - dl));
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) a_mbar_arrive(spempty);
            a_mbar_wait(pdempty, ((uint32_t)i & 1u) ^ 1u);            // the output MMAs of block i-1 have read the tiles
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const uint32_t off = (((uint32_t)g ^ sw) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + off), "r"(pk[g * 4]), "r"(pk[g * 4 + 1]), "r"(pk[g * 4 + 2]), "r"(pk[g * 4 + 3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(drow + off), "r"(dk[g * 4]), "r"(dk[g * 4 + 1]), "r"(dk[g * 4 + 2]), "r"(dk[g * 4 + 3]) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) a_mbar_arrive(pdready);
        }
    } else if (warp >= 12) {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
        for (int i = 0; i < nqb; ++i) {
            const uint32_t qb = (uint32_t)i & 1u;
            a_mbar_wait(dqfull0 + 8 * qb, ((uint32_t)i >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[64];
            A_TMEM_LD32(tDQ + tlane + qb * 64, r);
            A_TMEM_LD32(tDQ + tlane + qb * 64 + 32, r + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) a_mbar_arrive(dqempty0 + 8 * qb);
            const int q = i * BQ + row;
            if (q < bp.Lq && !(bp.dbg & 8)) {
                float* dst = bp.dq_acc + (long long)b * bp.dq_sb + (long long)q * bp.dq_ld + (long long)h * bp.d;
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                    if (g * 4 < bp.d)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + g * 4), "f"(__uint_as_float(r[g * 4]) * bp.alpha),
                                     "f"(__uint_as_float(r[g * 4 + 1]) * bp.alpha), "f"(__uint_as_float(r[g * 4 + 2]) * bp.alpha),
                                     "f"(__uint_as_float(r[g * 4 + 3]) * bp.alpha) : "memory");
                }
            }
        }
        a_mbar_wait(accfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int key = k0 + row;
        {
            uint32_t r[64];
            A_TMEM_LD32(tDK + tlane, r);
            A_TMEM_LD32(tDK + tlane + 32, r + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (key < bp.Lk) {
                uint16_t* dst = bp.dk + (long long)b * bp.dk_sb + (long long)key * bp.dk_ld + (long long)h * bp.d;
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    if (g * 8 < bp.d)
                        *reinterpret_cast<uint4*>(dst + g * 8) =
                            make_uint4(pack_bf16x2(__uint_as_float(r[g * 8]) * bp.alpha, __uint_as_float(r[g * 8 + 1]) * bp.alpha),
                                       pack_bf16x2(__uint_as_float(r[g * 8 + 2]) * bp.alpha, __uint_as_float(r[g * 8 + 3]) * bp.alpha),
                                       pack_bf16x2(__uint_as_float(r[g * 8 + 4]) * bp.alpha, __uint_as_float(r[g * 8 + 5]) * bp.alpha),
                                       pack_bf16x2(__uint_as_float(r[g * 8 + 6]) * bp.alpha, __uint_as_float(r[g * 8 + 7]) * bp.alpha));
            }
        }
        if (bp.do_v) {
            uint32_t r[64];
            A_TMEM_LD32(tDV + tlane, r);
            A_TMEM_LD32(tDV + tlane + 32, r + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (key < bp.Lk) {
                uint16_t* dst = bp.dvp + (long long)b * bp.dv_sb + (long long)key * bp.dv_ld + (long long)h * bp.dv;
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    if (g * 8 < bp.dv)
                        *reinterpret_cast<uint4*>(dst + g * 8) =
                            make_uint4(pack_bf16x2(__uint_as_float(r[g * 8]), __uint_as_float(r[g * 8 + 1])), pack_bf16x2(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3])),
                                       pack_bf16x2(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5])), pack_bf16x2(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7])));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

extern "C" __attribute__((visibility("default"))) int spe_attention_fwd(const spe_attention_args* a, void* stream) {
    SPE_CHECK(a && a->q && a->k && a->v && a->out, "spe_attention_fwd: null argument");
    SPE_CHECK(a->B > 0 && a->H > 0 && a->Lq > 0 && a->Lk > 0, "spe_attention_fwd: bad shape");
    SPE_CHECK(a->d > 0 && a->d <= 64 && a->d % 16 == 0, "spe_attention_fwd: head dim %d must be a multiple of 16 and <= 64", a->d);
    SPE_CHECK(a->dv > 0 && a->dv <= 64 && a->dv % 16 == 0, "spe_attention_fwd: value head dim %d must be a multiple of 16 and <= 64", a->dv);
    const bool two = a->q2 != nullptr;
    SPE_CHECK(!two || (a->k2 && a->d2 > 0 && a->d2 <= 64 && a->d2 % 16 == 0), "spe_attention_fwd: bad second QK segment");
    SPE_CHECK((a->out_ld % 8) == 0 && (a->out_sb % 8) == 0 && ((a->H * a->dv) % 8) == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
              "spe_attention_fwd: output must be 16-byte regular");
    CUtensorMap tQ, tK, tV, tQ2, tK2, tP;
    memset(&tQ2, 0, sizeof(tQ2)); memset(&tK2, 0, sizeof(tK2)); memset(&tP, 0, sizeof(tP));
    // 4-D maps (d, token, head, image); K-major boxes [64 x 128 tokens], the V map is MN-major: boxes [64 (dv) x 64 keys]
    if (spe_make_tmap_bf16(&tQ, a->q, SPE_MAJOR_K, a->Lq, a->d, a->q_ld, a->q_sb, a->d, a->B, a->H, BQ)) return -1;
    if (spe_make_tmap_bf16(&tK, a->k, SPE_MAJOR_K, a->Lk, a->d, a->k_ld, a->k_sb, a->d, a->B, a->H, BKV)) return -1;
    if (spe_make_tmap_bf16(&tV, a->v, SPE_MAJOR_MN, a->dv, a->Lk, a->v_ld, a->v_sb, a->dv, a->B, a->H, 64)) return -1;
    if (two) {
        if (spe_make_tmap_bf16(&tQ2, a->q2, SPE_MAJOR_K, a->Lq, a->d2, a->q2_ld, a->q2_sb, a->d2, a->B, a->H, BQ)) return -1;
        if (spe_make_tmap_bf16(&tK2, a->k2, SPE_MAJOR_K, a->Lk, a->d2, a->k2_ld, a->k2_sb, a->d2, a->B, a->H, BKV)) return -1;
    }
    if (a->P) {
        PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(spe_tmap_encode_fn());
        SPE_CHECK(enc, "cuTensorMapEncodeTiled not available");
        SPE_CHECK(a->ldP % 8 == 0 && a->ldP >= a->Lk && (reinterpret_cast<uintptr_t>(a->P) & 15) == 0, "spe_attention_fwd: P must be 16-byte regular, ldP >= Lk");
        cuuint64_t gdim[4] = {(cuuint64_t)a->ldP, (cuuint64_t)a->Lq, (cuuint64_t)a->H, (cuuint64_t)a->B};
        cuuint64_t gstr[3] = {(cuuint64_t)a->ldP * 2, (cuuint64_t)a->Lq * a->ldP * 2, (cuuint64_t)a->H * a->Lq * a->ldP * 2};
        cuuint32_t box[4] = {64, 128, 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tP, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->P, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SPE_CHECK(r == CUDA_SUCCESS, "spe_attention_fwd: tensor map for P failed (%d)", (int)r);
    }
    AttnParams ap;
    memset(&ap, 0, sizeof(ap));
    ap.Lq = a->Lq; ap.Lk = a->Lk; ap.H = a->H;
    ap.nkb = (a->Lk + BKV - 1) / BKV;
    ap.two = two ? 1 : 0;
    ap.ksteps1 = a->d / 16; ap.ksteps2 = two ? a->d2 / 16 : 0;
    ap.dv = a->dv;
    ap.scale2 = a->scale * LOG2E_F;
    ap.mask = a->mask;
    ap.out = reinterpret_cast<uint16_t*>(a->out); ap.out_ld = a->out_ld; ap.out_sb = a->out_sb;
    ap.lse = a->lse;
    ap.store_p = a->P ? 1 : 0;
    ap.ldP = (int)a->ldP;
    {
        static const int dbg = getenv("SPE_ATTN_DBG") ? atoi(getenv("SPE_ATTN_DBG")) : 0;
        ap.dbg = dbg;
    }
    const size_t smem = OFF_MBITS + (size_t)ap.nkb * 16 + 1024;
    SPE_CHECK(smem <= 232448, "spe_attention_fwd: Lk = %d too long for the mask table", a->Lk);
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        SPE_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // algorithmic bytes: Q, K, V (+Q2, K2) read, O written, P written (bf16) -- the logits never reach HBM
    const double bytes = 2.0 * a->B * a->H * ((double)a->Lq * (a->d + (two ? a->d2 : 0) + a->dv) + (double)a->Lk * (a->d + (two ? a->d2 : 0) + a->dv)) +
                         (a->P ? 2.0 * a->B * a->H * (double)a->Lq * a->Lk : 0.0);
    SpeProfScope prof(SPE_FAM_ATTN_FUSED, bytes, st);
    dim3 grid((a->Lq + BQ - 1) / BQ, a->H, a->B);
    attn_fwd_kernel<<<grid, AT_THREADS, smem, st>>>(tQ, tK, tV, tQ2, tK2, tP, ap);
    SPE_LAUNCHED();
    return 0;
}

static int make_tmap_n2(CUtensorMap* tm, const void* ptr, int B, int H, int Lq, int64_t ld) {
    PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(spe_tmap_encode_fn());
    SPE_CHECK(enc, "cuTensorMapEncodeTiled not available");
    SPE_CHECK(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "attention N^2 operand must be 16-byte regular");
    cuuint64_t gdim[4] = {(cuuint64_t)ld, (cuuint64_t)Lq, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)Lq * ld * 2, (cuuint64_t)H * Lq * ld * 2};
    cuuint32_t box[4] = {64, 128, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SPE_CHECK(r == CUDA_SUCCESS, "tensor map for an attention N^2 operand failed (%d)", (int)r);
    return 0;
}

extern "C" __attribute__((visibility("default"))) int64_t spe_attention_bwd_gemms_workspace(int B, int H, int Lq, int d) { return (int64_t)B * Lq * H * d; }

extern "C" __attribute__((visibility("default"))) int spe_attention_bwd_gemms(const spe_attention_bwd_args* a, void* stream) {
    SPE_CHECK(a && a->dS && a->q && a->k && a->dq && a->dk && a->workspace, "spe_attention_bwd_gemms: null argument");
    SPE_CHECK(a->B > 0 && a->H > 0 && a->Lq > 0 && a->Lk > 0, "spe_attention_bwd_gemms: bad shape");
    SPE_CHECK(a->d > 0 && a->d <= 64 && a->d % 16 == 0, "spe_attention_bwd_gemms: head dim %d must be a multiple of 16 and <= 64", a->d);
    const bool do_v = a->dv_out != nullptr;
    SPE_CHECK(!do_v || (a->P && a->dO && a->dv > 0 && a->dv <= 64 && a->dv % 16 == 0), "spe_attention_bwd_gemms: dV needs P, dO, dv_out and dv in {16..64}");
    SPE_CHECK(a->ld >= a->Lk, "spe_attention_bwd_gemms: ld < Lk");
    const bool from_dp = a->delta != nullptr;
    SPE_CHECK(!from_dp || a->P, "spe_attention_bwd_gemms: delta given (dS operand holds dP) needs P");
    CUtensorMap tDS, tP, tQ, tK, tDO;
    memset(&tP, 0, sizeof(tP)); memset(&tDO, 0, sizeof(tDO));
    if (make_tmap_n2(&tDS, a->dS, a->B, a->H, a->Lq, a->ld)) return -1;
    if (a->P && make_tmap_n2(&tP, a->P, a->B, a->H, a->Lq, a->ld)) return -1;
    if (spe_make_tmap_bf16(&tQ, a->q, SPE_MAJOR_K, a->Lq, a->d, a->q_ld, a->q_sb, a->d, a->B, a->H, BQ)) return -1;
    if (spe_make_tmap_bf16(&tK, a->k, SPE_MAJOR_K, a->Lk, a->d, a->k_ld, a->k_sb, a->d, a->B, a->H, BKV)) return -1;
    if (do_v && spe_make_tmap_bf16(&tDO, a->dO, SPE_MAJOR_K, a->Lq, a->dv, a->do_ld, a->do_sb, a->dv, a->B, a->H, BQ)) return -1;
    SPE_CHECK(a->dq_ld % 4 == 0 && a->dq_sb % 4 == 0 && a->dk_ld % 8 == 0 && a->dk_sb % 8 == 0 && (!do_v || (a->dv_ld % 8 == 0 && a->dv_sb % 8 == 0)),
              "spe_attention_bwd_gemms: output pitches must be 16-byte regular");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int E = a->H * a->d;
    SPE_CUDA(cudaMemsetAsync(a->workspace, 0, (size_t)a->B * a->Lq * E * 4, st));
    BwdParams bp;
    memset(&bp, 0, sizeof(bp));
    bp.Lq = a->Lq; bp.Lk = a->Lk; bp.H = a->H; bp.nqb = (a->Lq + BQ - 1) / BQ;
    bp.d = a->d; bp.dv = do_v ? a->dv : 16; bp.do_v = do_v ? 1 : 0;
    bp.alpha = a->alpha;
    bp.from_dp = from_dp ? 1 : 0;
    bp.delta = a->delta;
    { static const int dbg = getenv("SPE_ATTN_DBG") ? atoi(getenv("SPE_ATTN_DBG")) : 0; bp.dbg = dbg; }
    bp.dq_acc = a->workspace; bp.dq_ld = E; bp.dq_sb = (long long)a->Lq * E;
    bp.dk = reinterpret_cast<uint16_t*>(a->dk); bp.dk_ld = a->dk_ld; bp.dk_sb = a->dk_sb;
    bp.dvp = reinterpret_cast<uint16_t*>(a->dv_out); bp.dv_ld = a->dv_ld; bp.dv_sb = a->dv_sb;
    constexpr size_t SMEM = BW_OFF_BAR + BW_NBAR * 8 + 16 + 1024;
    static_assert(SMEM <= 232448, "shared memory budget exceeded");
    static bool attr_done = false;
    if (!attr_done) {
        SPE_CUDA(cudaFuncSetAttribute(attn_bwd_gemms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        attr_done = true;
    }
    {
        // algorithmic bytes: dS (+ P) read once, Q / K / dO read, dQ / dK / dV written
        const double n2 = 2.0 * a->B * a->H * (double)a->Lq * a->Lk;
        SpeProfScope prof(SPE_FAM_GEMM_ATTN, n2 * (do_v ? 2.0 : 1.0) + 2.0 * a->B * a->H * ((double)a->Lq * (2 * a->d + a->dv) + (double)a->Lk * (2 * a->d + a->dv)), st);
        dim3 grid((a->Lk + BKV - 1) / BKV, a->H, a->B);
        attn_bwd_gemms_kernel<<<grid, BW_THREADS, SMEM, st>>>(tDS, tP, tQ, tK, tDO, bp);
        SPE_LAUNCHED();
    }
    const long long rows = (long long)a->B * a->Lq;
    long long blocks = (rows * (E / 4) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    dq_cast_kernel<<<(int)blocks, 256, 0, st>>>(a->workspace, rows, E, reinterpret_cast<uint16_t*>(a->dq), a->dq_ld, a->dq_sb, a->Lq);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_attention_delta(const void* dO, const void* O, int B, int H, int Lq, int dv, int64_t do_ld, int64_t do_sb,
                                                                         int64_t o_ld, int64_t o_sb, float* delta, void* stream) {
    SPE_CHECK(dO && O && delta && B > 0 && H > 0 && Lq > 0 && dv > 0, "spe_attention_delta: bad argument");
    const long long toks = (long long)B * Lq;
    attn_delta_kernel<<<(unsigned)((toks + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint16_t*>(dO), reinterpret_cast<const uint16_t*>(O), B, H, Lq,
                                                                                                    dv, do_ld, do_sb, o_ld, o_sb, delta);
    SPE_LAUNCHED();
    return 0;
}

// one launch of attn_bwd_kernel: gradients of segment (qa, ka) [+ dV]; the logits use both segments
static int attn_bwd_launch(const spe_attention_bwd2_args* a, const void* qa, int64_t qa_ld, int64_t qa_sb, const void* ka, int64_t ka_ld, int64_t ka_sb, int da,
                           const void* qb, int64_t qb_ld, int64_t qb_sb, const void* kb, int64_t kb_ld, int64_t kb_sb, int db, bool do_v,
                           void* dq, int64_t dq_ld, int64_t dq_sb, void* dk, int64_t dk_ld, int64_t dk_sb, cudaStream_t st) {
    CUtensorMap tQa, tKa, tQb, tKb, tV, tDO;
    memset(&tQb, 0, sizeof(tQb)); memset(&tKb, 0, sizeof(tKb));
    if (spe_make_tmap_bf16(&tQa, qa, SPE_MAJOR_K, a->Lq, da, qa_ld, qa_sb, da, a->B, a->H, BQ)) return -1;
    if (spe_make_tmap_bf16(&tKa, ka, SPE_MAJOR_K, a->Lk, da, ka_ld, ka_sb, da, a->B, a->H, BKV)) return -1;
    if (qb) {
        if (spe_make_tmap_bf16(&tQb, qb, SPE_MAJOR_K, a->Lq, db, qb_ld, qb_sb, db, a->B, a->H, BQ)) return -1;
        if (spe_make_tmap_bf16(&tKb, kb, SPE_MAJOR_K, a->Lk, db, kb_ld, kb_sb, db, a->B, a->H, BKV)) return -1;
    }
    if (spe_make_tmap_bf16(&tV, a->v, SPE_MAJOR_K, a->Lk, a->dv, a->v_ld, a->v_sb, a->dv, a->B, a->H, BKV)) return -1;
    if (spe_make_tmap_bf16(&tDO, a->dO, SPE_MAJOR_K, a->Lq, a->dv, a->do_ld, a->do_sb, a->dv, a->B, a->H, BQ)) return -1;
    const int E = a->H * da;
    SPE_CUDA(cudaMemsetAsync(a->workspace, 0, (size_t)a->B * a->Lq * E * 4, st));
    Bwd2Params bp;
    memset(&bp, 0, sizeof(bp));
    bp.Lq = a->Lq; bp.Lk = a->Lk; bp.H = a->H; bp.nqb = (a->Lq + BQ - 1) / BQ;
    bp.ksa = da / 16; bp.ksb = qb ? db / 16 : 0; bp.ksv = a->dv / 16;
    bp.d = da; bp.dv = a->dv; bp.do_v = do_v ? 1 : 0;
    bp.scale2 = a->scale * LOG2E_F; bp.alpha = a->scale;
    bp.mask = a->mask; bp.lse = a->lse; bp.delta = a->delta;
    { static const int dbg = getenv("SPE_ATTN_DBG") ? atoi(getenv("SPE_ATTN_DBG")) : 0; bp.dbg = dbg; }
    bp.dq_acc = a->workspace; bp.dq_ld = E; bp.dq_sb = (long long)a->Lq * E;
    bp.dk = reinterpret_cast<uint16_t*>(dk); bp.dk_ld = dk_ld; bp.dk_sb = dk_sb;
    bp.dvp = reinterpret_cast<uint16_t*>(a->dv_out); bp.dv_ld = a->dv_ld; bp.dv_sb = a->dv_sb;
    constexpr size_t SMEM = B2_OFF_BAR + B2_NBAR * 8 + 16 + 16 + 1024;
    static_assert(SMEM <= 232448, "shared memory budget exceeded");
    static bool attr_done = false;
    if (!attr_done) {
        SPE_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        attr_done = true;
    }
    {
        // algorithmic bytes: Q, K, V, dO read, dQ, dK, dV written -- no N^2 tensor touches HBM
        const double bytes = 2.0 * a->B * a->H * ((double)a->Lq * (2 * da + (qb ? db : 0) + a->dv) + (double)a->Lk * (2 * da + (qb ? db : 0) + 2 * a->dv));
        SpeProfScope prof(SPE_FAM_ATTN_FUSED, bytes, st);
        dim3 grid((a->Lk + BKV - 1) / BKV, a->H, a->B);
        attn_bwd_kernel<<<grid, B2_THREADS, SMEM, st>>>(tQa, tKa, tQb, tKb, tV, tDO, bp);
        SPE_LAUNCHED();
    }
    const long long rows = (long long)a->B * a->Lq;
    long long blocks = (rows * (E / 4) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    dq_cast_kernel<<<(int)blocks, 256, 0, st>>>(a->workspace, rows, E, reinterpret_cast<uint16_t*>(dq), dq_ld, dq_sb, a->Lq);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_attention_bwd(const spe_attention_bwd2_args* a, void* stream) {
    SPE_CHECK(a && a->q && a->k && a->v && a->dO && a->lse && a->delta && a->dq && a->dk && a->dv_out && a->workspace, "spe_attention_bwd: null argument");
    SPE_CHECK(a->B > 0 && a->H > 0 && a->Lq > 0 && a->Lk > 0, "spe_attention_bwd: bad shape");
    SPE_CHECK(a->d > 0 && a->d <= 64 && a->d % 16 == 0 && a->dv > 0 && a->dv <= 64 && a->dv % 16 == 0, "spe_attention_bwd: head dims must be multiples of 16 and <= 64");
    const bool two = a->q2 != nullptr;
    SPE_CHECK(!two || (a->k2 && a->dq2 && a->dk2 && a->d2 > 0 && a->d2 <= 64 && a->d2 % 16 == 0), "spe_attention_bwd: bad second QK segment");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (attn_bwd_launch(a, a->q, a->q_ld, a->q_sb, a->k, a->k_ld, a->k_sb, a->d, a->q2, a->q2_ld, a->q2_sb, a->k2, a->k2_ld, a->k2_sb, a->d2, true,
                        a->dq, a->dq_ld, a->dq_sb, a->dk, a->dk_ld, a->dk_sb, st)) return -1;
    if (two && attn_bwd_launch(a, a->q2, a->q2_ld, a->q2_sb, a->k2, a->k2_ld, a->k2_sb, a->d2, a->q, a->q_ld, a->q_sb, a->k, a->k_ld, a->k_sb, a->d, false,
                               a->dq2, a->dq2_ld, a->dq2_sb, a->dk2, a->dk2_ld, a->dk2_sb, st)) return -1;
    return 0;
}
