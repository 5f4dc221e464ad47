// Fused talking-heads attention for sm_100a (Attention_talking_head.forward, /root/reference/models/cait.py:374-393):
//   S_h = scale Q_h K_h^T,  L_g = sum_h Wl[g,h] S_h + bl[g],  P_g = softmax_keys(L_g),  A_g = sum_h Ww[g,h] P_h + bw[g],  O_g = A_g V_g
// No [B,H,N,N] tensor ever reaches HBM (the unfused pipeline moved ~4.4 GB of them per layer, forward + backward).
//
// Tile geometry (all kernels of this file): a CTA owns RB = 64 "stationary" rows (queries; keys in the key-stationary backward
// kernel) of one image for ALL heads and streams the other side in blocks of CB = 16 columns:
//   * tcgen05.mma M = 64, N = 16, K = 16 (bf16, f32 accumulate) per head: the H logit tiles of a block land side by side in
//     H x 16 TMEM columns.  An M = 64 accumulator occupies lanes [32 i, 32 i + 16) of the four lane quarters; a second tile
//     (the other buffer of the double-buffered S, or dA = dO V^T in the backward) interleaves into lanes [32 i + 16, 32 i + 32)
//     of the SAME columns, so one tcgen05.ld.16x256b per head and tile hands a thread S and dA of the same four positions.
//   * position warps (8: two per TMEM lane quarter, one per 8-column half): tcgen05.ld.16x256b gives thread (gid, tig) the rows
//     {gid, gid + 8} x columns {2 tig, 2 tig + 1} of every head = the mma.sync m16n8 accumulator layout.  The H x H head mixes run
//     on mma.sync.m16n8k16 straight from those registers: the thread's heads fill its own k-slots of the A fragment and the
//     weight operand is block diagonal over the quad (B[(t, i), (t', gs)] = [t == t'] W[2 gp + gs, head(i)]), so every output
//     lands in the thread that owns the position -- no shuffles, no shared-memory transposes (1024 MAC / clk / SM on the legacy
//     tensor path against 128 on the FMA pipe; measured with tools/micro/pipes.cu).  exp2 on the MUFU pipe with the softmax
//     statistics folded into the accumulator initialiser (C operand) of the first mix.
//   * the mixed probabilities (bf16) go to a [64 x 16] K-major no-swizzle tile per head in shared memory and feed
//     tcgen05.mma O_g += A_g V_g (V consumed MN-major from the streamed block); O for all heads: 64 x 384 f32 in TMEM.
// Forward = two launches: `stats` (row max / sum of the mixed logits, per column chunk) and `main` (exact P -> A -> A V).  The
// streamed range is split into `nchunk` chunks per row block so that the grid fills the 148 SMs; partial statistics are merged in
// the main kernel's prologue, partial outputs are reduce-added in fp32 when nchunk > 1.
#include "common.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <stdlib.h>

void* spe_tmap_encode_fn();

namespace {
using namespace umma;

constexpr int RB = 64;                  // stationary rows per CTA
constexpr int DHD = 48;                 // head dim (every CaiT variant: 192/4, 288/6, 384/8, 768/16)
constexpr int TF_THREADS = 320;         // warp 0 TMA, warp 1 MMA + TMEM, warps 2..9 positions (quarter = warp % 4)
constexpr int TF_THREADS_STATS = 576;   // statistics kernel: 16 position warps (80 registers per thread)
#ifndef TF_NPW_MAIN
#define TF_NPW_MAIN 8
#endif
constexpr float LOG2E = 1.4426950408889634f;
constexpr float P_SHIFT = 8.f;          // probabilities are carried as 2^8 P (fp16 operand of the second mix stays normal)
constexpr uint32_t XT_B = RB * 128;     // one [64 rows x 64 cols] bf16 SWIZZLE_128B tile of the stationary operand

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [B, N, D] bf16 view (token stride ld, image stride sb, elements): box = 64 columns x `rows` tokens, SWIZZLE_128B, OOB rows read as zero
int make_map(CUtensorMap* tm, const void* ptr, int D, int N, int B, int64_t ld, int64_t sb, int rows) {
    PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(spe_tmap_encode_fn());
    SPE_CHECK(enc, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    SPE_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0 && (B == 1 || sb % 8 == 0), "talking_fused: operand not 16-byte aligned");
    cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(B > 1 ? sb : ld * N) * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SPE_CHECK(r == CUDA_SUCCESS, "talking_fused: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// head mix on mma.sync:  y[g][i] = c(g, i) + sum_h W[g][h] x[h][i]   for the thread's four positions i = rowsel * 2 + j
// ---------------------------------------------------------------------------------------------------------------------------
// B fragments of a mix matrix W (out g x in h, row-major f32 in global memory, times `mul`): [hq][gp][2] packed 16-bit pairs.
// Thread (gid, tig) holds column n = gid of the block-diagonal operand: non-zero only for the k-slots of thread gid >> 1.
template <int H, bool BF16, bool TRANSPOSE>
__device__ __forceinline__ void load_wfrag(const float* __restrict__ W, float mul, int lane, uint32_t (&wf)[(H / 4) * (H / 2)][2]) {
    const int gid = lane >> 2, tig = lane & 3;
    const bool active = tig == (gid >> 1);
#pragma unroll
    for (int hq = 0; hq < H / 4; ++hq)
#pragma unroll
        for (int gp = 0; gp < H / 2; ++gp) {
            const int g = 2 * gp + (gid & 1);
            float w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int h = 4 * hq + i;
                w[i] = active ? mul * (TRANSPOSE ? W[h * H + g] : W[g * H + h]) : 0.f;
            }
            wf[hq * (H / 2) + gp][0] = BF16 ? pack_bf16(w[0], w[1]) : pack_f16(w[0], w[1]);
            wf[hq * (H / 2) + gp][1] = BF16 ? pack_bf16(w[2], w[3]) : pack_f16(w[2], w[3]);
        }
}

template <int H, bool BF16, class CInit>
__device__ __forceinline__ void head_mix(const float (&x)[H][4], const uint32_t (&wf)[(H / 4) * (H / 2)][2], CInit cinit, float (&y)[H][4]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        uint32_t a[H / 4][4];
#pragma unroll
        for (int hq = 0; hq < H / 4; ++hq) {
            if (BF16) {
                a[hq][0] = pack_bf16(x[4 * hq][j], x[4 * hq + 1][j]);
                a[hq][1] = pack_bf16(x[4 * hq][2 + j], x[4 * hq + 1][2 + j]);
                a[hq][2] = pack_bf16(x[4 * hq + 2][j], x[4 * hq + 3][j]);
                a[hq][3] = pack_bf16(x[4 * hq + 2][2 + j], x[4 * hq + 3][2 + j]);
            } else {
                a[hq][0] = pack_f16(x[4 * hq][j], x[4 * hq + 1][j]);
                a[hq][1] = pack_f16(x[4 * hq][2 + j], x[4 * hq + 1][2 + j]);
                a[hq][2] = pack_f16(x[4 * hq + 2][j], x[4 * hq + 3][j]);
                a[hq][3] = pack_f16(x[4 * hq + 2][2 + j], x[4 * hq + 3][2 + j]);
            }
        }
#pragma unroll
        for (int gp = 0; gp < H / 2; ++gp) {
            float d[4] = {cinit(2 * gp, j), cinit(2 * gp + 1, j), cinit(2 * gp, 2 + j), cinit(2 * gp + 1, 2 + j)};
#pragma unroll
            for (int hq = 0; hq < H / 4; ++hq) {
                if (BF16) hmma_bf16(d, a[hq], wf[hq * (H / 2) + gp][0], wf[hq * (H / 2) + gp][1], d);
                else hmma_f16(d, a[hq], wf[hq * (H / 2) + gp][0], wf[hq * (H / 2) + gp][1], d);
            }
            y[2 * gp][j] = d[0]; y[2 * gp + 1][j] = d[1]; y[2 * gp][2 + j] = d[2]; y[2 * gp + 1][2 + j] = d[3];
        }
    }
}

// A fragments of the thread's four positions: [j (column parity)][hq (head quad)][4]
template <int H, bool BF16>
__device__ __forceinline__ void pack_afrag(const float (&x)[H][4], uint32_t (&a)[2][H / 4][4]) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int hq = 0; hq < H / 4; ++hq) {
            if (BF16) {
                a[j][hq][0] = pack_bf16(x[4 * hq][j], x[4 * hq + 1][j]);
                a[j][hq][1] = pack_bf16(x[4 * hq][2 + j], x[4 * hq + 1][2 + j]);
                a[j][hq][2] = pack_bf16(x[4 * hq + 2][j], x[4 * hq + 3][j]);
                a[j][hq][3] = pack_bf16(x[4 * hq + 2][2 + j], x[4 * hq + 3][2 + j]);
            } else {
                a[j][hq][0] = pack_f16(x[4 * hq][j], x[4 * hq + 1][j]);
                a[j][hq][1] = pack_f16(x[4 * hq][2 + j], x[4 * hq + 1][2 + j]);
                a[j][hq][2] = pack_f16(x[4 * hq + 2][j], x[4 * hq + 3][j]);
                a[j][hq][3] = pack_f16(x[4 * hq + 2][2 + j], x[4 * hq + 3][2 + j]);
            }
        }
}
// head mix with a hand-interleaved side job: the H*H/4 mma.sync are issued in (j, hq, gp) order -- the accumulation chains of the
// H/2 output pairs stay H/2 instructions apart -- and `side(k)` runs after the k-th one (exp2 of another sub-tile: the MUFU pipe
// works under the tensor pipe from ONE warp; left to the compiler the phases serialise, which is what bounded the first version)
template <int H, bool BF16, class CInit, class Side>
__device__ __forceinline__ void head_mix_il(const uint32_t (&a)[2][H / 4][4], const uint32_t (&wf)[(H / 4) * (H / 2)][2], CInit cinit, float (&y)[H][4], Side side) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float d[H / 2][4];
#pragma unroll
        for (int gp = 0; gp < H / 2; ++gp) { d[gp][0] = cinit(2 * gp, j); d[gp][1] = cinit(2 * gp + 1, j); d[gp][2] = cinit(2 * gp, 2 + j); d[gp][3] = cinit(2 * gp + 1, 2 + j); }
#pragma unroll
        for (int hq = 0; hq < H / 4; ++hq)
#pragma unroll
            for (int gp = 0; gp < H / 2; ++gp) {
                if (BF16) hmma_bf16_v(d[gp], a[j][hq], wf[hq * (H / 2) + gp][0], wf[hq * (H / 2) + gp][1]);
                else hmma_f16_v(d[gp], a[j][hq], wf[hq * (H / 2) + gp][0], wf[hq * (H / 2) + gp][1]);
                side((j * (H / 4) + hq) * (H / 2) + gp);
            }
#pragma unroll
        for (int gp = 0; gp < H / 2; ++gp) { y[2 * gp][j] = d[gp][0]; y[2 * gp + 1][j] = d[gp][1]; y[2 * gp][2 + j] = d[gp][2]; y[2 * gp + 1][2 + j] = d[gp][3]; }
    }
}

// the H tiles of one block: head h at column h * CB (+ this warp's 8-column half), lanes [lane_base, lane_base + 16)
template <int H, int CB>
__device__ __forceinline__ void ld_tiles(uint32_t taddr, float (&x)[H][4]) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
        uint32_t r[4];
        UMMA_LD_16x256(taddr + (uint32_t)(h * CB), r);
#pragma unroll
        for (int i = 0; i < 4; ++i) x[h][i] = __uint_as_float(r[i]);
    }
}

// bf16 operand tiles for the accumulating tcgen05.mma: [8 row groups][H][CB / 8 column chunks][8 rows][16 B]  (K-major, no swizzle:
// LBO = 128 between the 8-column chunks, SBO = H * (CB / 8) * 128 between the 8-row groups; head g starts at g * (CB / 8) * 128)
template <int H, int CB>
__device__ __forceinline__ void st_tiles(uint32_t tile, int quarter, int chunk8, int lane, const float (&y)[H][4]) {
    const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int g = 0; g < H; ++g)
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
            const uint32_t addr = tile + (uint32_t)((quarter * 2 + rs) * (H * (CB / 8) * 128) + g * ((CB / 8) * 128) + chunk8 * 128 + gid * 16 + tig * 4);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(pack_bf16(y[g][rs * 2], y[g][rs * 2 + 1])) : "memory");
        }
}

struct TfFwdParams {
    int N, nblk, nchunk, bpc;            // tokens; column blocks in total / chunks / blocks per chunk (this launch's block size)
    int Npad;                            // row blocks * 64
    const float *Wl, *bl, *Ww, *bw;
    float scale;
    float2* part;                        // stats out: [B][nchunk][H][Npad] (max, sum) of 2^(log2e L) per row, chunk-local
    float* lse2;                         // main in: [B][H][N] log2-domain logsumexp (written by tf_merge_kernel; saved for the backward)
    uint16_t* out; long long out_ld, out_sb;      // bf16 [B, N, D]   (nchunk == 1)
    float* out32;                        // f32 [B, N, D] reduce-add target (nchunk > 1)
    int dbg;                             // timing experiments only (SPE_TF_DBG)
    long long* trace;                    // dbg & 512: clock64 stamps of CTA (0,0,0): [role][block][event]
};

#ifndef TF_CB_STATS
#define TF_CB_STATS 64
#endif
#ifndef TF_NSTG_STATS
#define TF_NSTG_STATS 2
#endif
#ifndef TF_NSTG_MAIN
#define TF_NSTG_MAIN 2
#endif
#ifndef TF_NABUF
#define TF_NABUF 2
#endif
constexpr int CB_STATS = TF_CB_STATS, CB_MAIN = 32;
#define TF_TRACE(role, blk, ev) do { if (p.trace && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && (blk) < 32 && lane == 0) p.trace[((role) * 32 + (blk)) * 8 + (ev)] = clock64(); } while (0)

template <int H, int CB, bool STATS>
struct FwdSmem {
    static constexpr int D = H * DHD, NT = D / 64;
    static constexpr int NSTG = STATS ? TF_NSTG_STATS : TF_NSTG_MAIN;  // stages of the streamed-block rings
    static constexpr int NABUF = TF_NABUF;
    static constexpr uint32_t YT = CB * 128;                           // one [CB rows x 64 cols] tile of a streamed block
    static constexpr uint32_t X = 0;                                   // Q: NT tiles [64 x 64]
    static constexpr uint32_t Y = X + NT * XT_B;                       // K ring
    static constexpr uint32_t Z = Y + NSTG * NT * YT;                  // V ring (main only)
    static constexpr uint32_t A = Z + (STATS ? 0 : NSTG * NT * YT);    // probability tiles [8 row groups][H][CB/8][8 rows][16 B]
    static constexpr uint32_t A_B = 8 * H * (CB / 8) * 128;
    static constexpr uint32_t BAR = A + (STATS ? 0 : NABUF * A_B);
    static constexpr int NBAR = 1 + 4 * NSTG + 8 + 1;
    static constexpr uint32_t TSLOT = BAR + NBAR * 8;
    static constexpr uint32_t WSM = TSLOT + 16;                        // Wl, Ww [H*H], bl, bw [H] (f32)
    static constexpr uint32_t LSE = WSM + (2 * H * H + 2 * H) * 4;     // main: lse2 of the CTA's rows [H][64];  stats: exchange [64][H] float2
    static constexpr uint32_t TOTAL = LSE + 3 * RB * H * 8 + 1024;     // + alignment slack
    static constexpr uint32_t S_COLS = H * CB;                         // TMEM: S tile, the two buffers in the two lane halves
    static constexpr uint32_t O_CAP = 512 - S_COLS;                    // O columns that fit lane half 0 beside S; the rest goes to lane half 1
    static_assert(STATS || (D <= 2 * (int)O_CAP && O_CAP % 64 == 0), "O does not fit TMEM");
    static constexpr uint32_t TMEM_COLS = STATS ? (S_COLS <= 32 ? 32 : S_COLS <= 64 ? 64 : S_COLS <= 128 ? 128 : S_COLS <= 256 ? 256 : 512) : 512;
};

// merge (m, l) pairs of the log2-domain running statistics; m = -1e30 (finite) means "nothing seen"
__device__ __forceinline__ void stat_merge(float& m, float& l, float m2, float l2) {
    const float M = fmaxf(m, m2);
    const float s1 = (m == M) ? 1.f : ex2(m - M), s2 = (m2 == M) ? 1.f : ex2(m2 - M);
    l = l * s1 + l2 * s2;
    m = M;
}

template <int H, int CB, bool STATS>
__global__ void __launch_bounds__(TF_THREADS, 1) tf_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                               const __grid_constant__ CUtensorMap tmV, const TfFwdParams p) {
    using SM = FwdSmem<H, CB, STATS>;
    constexpr int D = SM::D, NT = SM::NT, NSTG = SM::NSTG;
    constexpr int NPW = 8;                                    // position warps: two per lane quarter = per SM sub-partition
    constexpr int NCG = NPW / 4;                              // column groups (warps per lane quarter)
    constexpr int NSUB = CB / (8 * NCG);                      // 8-column sub-tiles per warp and block
    constexpr uint32_t YT = SM::YT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + SM::BAR;
    const uint32_t xfull = bar0, yfull0 = xfull + 8, yempty0 = yfull0 + 8 * NSTG, zfull0 = yempty0 + 8 * NSTG, zempty0 = zfull0 + 8 * NSTG,
                   sfull0 = zempty0 + 8 * NSTG, sempty0 = sfull0 + 16, afull0 = sempty0 + 16, aempty0 = afull0 + 16, ofull = aempty0 + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::TSLOT);
    float* wsm = reinterpret_cast<float*>(smem + SM::WSM);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * RB, chunk = blockIdx.y, b = blockIdx.z;
    const int blk0 = chunk * p.bpc;
    const int nb = min(p.bpc, p.nblk - blk0);                 // blocks of this CTA (>= 1 by construction)

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
        if (!STATS) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
        mbar_init(xfull, 1);
        for (int s = 0; s < NSTG; ++s) { mbar_init(yfull0 + 8 * s, 1); mbar_init(yempty0 + 8 * s, 1); mbar_init(zfull0 + 8 * s, 1); mbar_init(zempty0 + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(sfull0 + 8 * s, 1); mbar_init(sempty0 + 8 * s, STATS ? NPW : NPW / 2); mbar_init(afull0 + 8 * s, NPW / 2); mbar_init(aempty0 + 8 * s, 1); }
        mbar_init(ofull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(SM::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {
        // mix weights and (main) the rows' statistics into shared memory: a handful of coalesced loads instead of ~150 dependent ones per thread
        const int t = threadIdx.x - 64;
        if (t < H * H) { wsm[t] = p.Wl[t] * (LOG2E * p.scale); wsm[H * H + t] = p.Ww[t]; }
        if (t < H) { wsm[2 * H * H + t] = p.bl[t] * LOG2E; wsm[2 * H * H + H + t] = p.bw[t] * 256.f; }       // 2^P_SHIFT bw
        if (!STATS) {
            float* ls = reinterpret_cast<float*>(smem + SM::LSE);
            for (int i = t; i < H * RB; i += NPW * 32) {
                const int g = i / RB, r = i % RB;
                ls[i] = r0 + r < p.N ? p.lse2[((long long)b * H + g) * p.N + r0 + r] : 0.f;
            }
        }
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base;                             // S: H * CB columns, buffer sb in lanes [32 i + 16 sb, +16)
    const uint32_t tO = tmem_base + SM::S_COLS;                // O (main): columns [0, O_CAP) in lane half 0, the rest in lane half 1

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (elect()) {
            mbar_expect_tx(xfull, NT * XT_B);
            for (int t = 0; t < NT; ++t) tma_load_3d(sbase + SM::X + t * XT_B, &tmQ, xfull, t * 64, r0, b);
        }
        __syncwarp();
        for (int jl = 0; jl < nb; ++jl) {
            const int s = jl % NSTG;
            const uint32_t ph = ((uint32_t)(jl / NSTG) & 1u) ^ 1u;
            const int c0 = (blk0 + jl) * CB;
            mbar_wait(yempty0 + 8 * s, ph);
            TF_TRACE(0, jl, 0);
            if (elect()) {
                mbar_expect_tx(yfull0 + 8 * s, NT * YT);
                for (int t = 0; t < NT; ++t) tma_load_3d(sbase + SM::Y + (s * NT + t) * YT, &tmK, yfull0 + 8 * s, t * 64, c0, b);
            }
            __syncwarp();
            if (!STATS) {
                mbar_wait(zempty0 + 8 * s, ph);
                TF_TRACE(0, jl, 1);
                if (elect()) {
                    mbar_expect_tx(zfull0 + 8 * s, NT * YT);
                    for (int t = 0; t < NT; ++t) tma_load_3d(sbase + SM::Z + (s * NT + t) * YT, &tmV, zfull0 + 8 * s, t * 64, c0, b);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t ID_S = idesc_bf16(RB, CB, false, false);
        const uint64_t dX = desc(sbase + SM::X, 0, 1024, 2);
        mbar_wait(xfull, 0);
        fence_after();
        for (int jl = 0; jl <= nb; ++jl) {
            if (jl < nb) {
                const int s = jl % NSTG;
                const uint32_t sb = (uint32_t)jl & 1u;
                TF_TRACE(1, jl, 0);
                mbar_wait(yfull0 + 8 * s, (uint32_t)(jl / NSTG) & 1u);
                TF_TRACE(1, jl, 1);
                mbar_wait(sempty0 + 8 * sb, (((uint32_t)jl >> 1) & 1u) ^ 1u);
                TF_TRACE(1, jl, 2);
                fence_after();
                if (elect()) {
                    const uint64_t dY = desc(sbase + SM::Y + s * NT * YT, 0, 1024, 2);
                    const uint32_t td = tS + ((sb * 16u) << 16);
                    if (!(p.dbg & 1) || jl == 0)
#pragma unroll
                    for (int h = 0; h < H; ++h)
#pragma unroll
                        for (int ks = 0; ks < DHD / 16; ++ks) {
                            constexpr int dummy = 0; (void)dummy;
                            const int e = h * DHD + ks * 16;
                            mma_f16(td + (uint32_t)(h * CB), dX + (uint64_t)(((e >> 6) * XT_B + (e & 63) * 2) >> 4),
                                    dY + (uint64_t)(((e >> 6) * YT + (e & 63) * 2) >> 4), ID_S, ks != 0);
                        }
                    commit(yempty0 + 8 * s);
                    commit(sfull0 + 8 * sb);
                }
                __syncwarp();
                TF_TRACE(1, jl, 3);
            }
            if (!STATS && jl >= 1) {
                const int jp = jl - 1;
                const int s = jp % NSTG;
                const uint32_t ab = (uint32_t)jp % SM::NABUF;
                mbar_wait(afull0 + 8 * ab, ((uint32_t)jp / SM::NABUF) & 1u);
                TF_TRACE(1, jp, 4);
                mbar_wait(zfull0 + 8 * s, (uint32_t)(jp / NSTG) & 1u);
                TF_TRACE(1, jp, 5);
                fence_after();
                if (elect()) {
                    // A_g tile: K-major no-swizzle (LBO 128 between the 8-column chunks, SBO between the 8-row groups), one k-step per
                    // 16 columns; V_g: MN-major inside the 64-column tiles of the streamed block -- a head that straddles a tile
                    // boundary takes two MMAs
                    const uint64_t dA = desc(sbase + SM::A + ab * SM::A_B, 128, H * (CB / 8) * 128, 0);
                    const uint64_t dZ = desc(sbase + SM::Z + s * NT * YT, YT, 1024, 2);
                    if (!(p.dbg & 2) || jp == 0)
#pragma unroll
                    for (int g = 0; g < H; ++g)
#pragma unroll
                        for (int seg = 0; seg < 2; ++seg) {
                            const int e0 = g * DHD;
                            const int n0 = (64 - (e0 & 63)) < DHD ? (64 - (e0 & 63)) : DHD;
                            const int e = seg == 0 ? e0 : e0 + n0, n = seg == 0 ? n0 : DHD - n0;
                            if (n > 0) {
                                const uint32_t td = (uint32_t)e < SM::O_CAP ? tO + (uint32_t)e : tO + (uint32_t)(e - SM::O_CAP) + (16u << 16);
#pragma unroll
                                for (int ks = 0; ks < CB / 16; ++ks)
                                    mma_f16(td, dA + (uint64_t)((g * (CB / 8) * 128 + ks * 256) >> 4),
                                            dZ + (uint64_t)(((e >> 6) * YT + ks * 2048 + (e & 63) * 2) >> 4), idesc_bf16(RB, n, false, true), (jp | ks) != 0);
                            }
                        }
                    commit(zempty0 + 8 * s);
                    commit(aempty0 + 8 * ab);
                }
                __syncwarp();
                TF_TRACE(1, jp, 6);
            }
        }
        if (!STATS) {
            if (elect()) commit(ofull);
            __syncwarp();
        }
    } else {
        // ---------------- position warps ----------------
        const int quarter = warp & 3, cg = (warp - 2) >> 2;     // lane quarter; column group of the block
        const int gid = lane >> 2, tig = lane & 3;
        const uint32_t tl = (uint32_t)(quarter * 32) << 16;
        uint32_t wl[(H / 4) * (H / 2)][2];
        load_wfrag<H, false, false>(wsm, 1.f, lane, wl);
        const int row_a = r0 + quarter * 16 + gid;             // + 8 for rowsel 1
        // The two warps of a lane quarter share one SM sub-partition, i.e. one mma.sync pipe and one MUFU pipe.  Left alone they run
        // in lockstep (every block barrier re-aligns them) and the pipes take turns idling: measured 3100 clocks per block against
        // ~1000 of work on either pipe.  A token (named barrier pair, FA3-style ping-pong) serialises their tensor phases
        // [store / load / head mixes], so that one warp's exponentials always run under the other warp's mixes.
        const uint32_t tok_mine = 2u + 2u * (uint32_t)quarter + (uint32_t)cg, tok_other = 2u + 2u * (uint32_t)quarter + (uint32_t)(cg ^ 1);
        auto tok_acquire = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(tok_mine) : "memory"); };
        auto tok_release = [&]() { asm volatile("bar.arrive %0, 64;" ::"r"(tok_other) : "memory"); };
        const int T = nb * NSUB;                                // sub-tiles of this warp
        if (cg == 1) tok_release();                             // the cg = 0 warp starts
        if (STATS) {
            float c1[H];
#pragma unroll
            for (int g = 0; g < H; ++g) c1[g] = wsm[2 * H * H + g];
            // running statistics of the thread's own columns: reference m (log2 domain), l = sum 2^(L - m).  The reference rides in
            // the accumulator initialiser of the mix (C = log2e bl - m), so a value costs one EX2 and one FADD; it is re-based only
            // when a logit exceeds it by more than 2^8 (warp-uniform branch, rare after the first blocks)
            float m[H][2], l[H][2];
#pragma unroll
            for (int g = 0; g < H; ++g) { m[g][0] = m[g][1] = -1e30f; l[g][0] = l[g][1] = 0.f; }
            float y[H][4];
            auto load_mix = [&](int t, bool first) {            // tensor phase: S tile of sub-tile t -> y = log2e L - m
                const int jl = t / NSUB, sub = t % NSUB;
                const uint32_t sb = (uint32_t)jl & 1u;
                if (sub == 0) { mbar_wait(sfull0 + 8 * sb, ((uint32_t)jl >> 1) & 1u); fence_after(); }
                float x[H][4];
                ld_tiles<H, CB>(tS + tl + ((sb * 16u) << 16) + (uint32_t)((cg * NSUB + sub) * 8), x);
                ld_wait();
                if (sub == NSUB - 1) {
                    fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sempty0 + 8 * sb);
                }
                if (first) head_mix<H, false>(x, wl, [&](int g, int) { return c1[g]; }, y);
                else head_mix<H, false>(x, wl, [&](int g, int i) { return c1[g] - m[g][i >> 1]; }, y);
            };
            tok_acquire();
            load_mix(0, true);
            tok_release();
            for (int t = 0; t < T; ++t) {
                const int jl = t / NSUB, sub = t % NSUB;
                const int col = (blk0 + jl) * CB + (cg * NSUB + sub) * 8 + 2 * tig;
                if (col + 1 >= p.N) {                               // tail block: columns beyond N do not exist
#pragma unroll
                    for (int g = 0; g < H; ++g) {
                        if (col >= p.N) { y[g][0] = -INFINITY; y[g][2] = -INFINITY; }
                        y[g][1] = -INFINITY; y[g][3] = -INFINITY;
                    }
                }
                if (t == 0) {
#pragma unroll
                    for (int g = 0; g < H; ++g)
#pragma unroll
                        for (int rs = 0; rs < 2; ++rs) {
                            m[g][rs] = fmaxf(fmaxf(y[g][rs * 2], y[g][rs * 2 + 1]), -1e30f);
                            l[g][rs] = ex2(y[g][rs * 2] - m[g][rs]) + ex2(y[g][rs * 2 + 1] - m[g][rs]);
                        }
                } else {
                    bool up = false;
#pragma unroll
                    for (int g = 0; g < H; ++g)
#pragma unroll
                        for (int rs = 0; rs < 2; ++rs) up = up || fmaxf(y[g][rs * 2], y[g][rs * 2 + 1]) > 8.f;
                    if (__any_sync(0xffffffffu, up)) {
#pragma unroll
                        for (int g = 0; g < H; ++g)
#pragma unroll
                            for (int rs = 0; rs < 2; ++rs) {
                                const float bm = fmaxf(y[g][rs * 2], y[g][rs * 2 + 1]);
                                if (bm > 8.f) {
                                    l[g][rs] *= ex2(-bm);
                                    m[g][rs] += bm;
                                    y[g][rs * 2] -= bm; y[g][rs * 2 + 1] -= bm;
                                }
                            }
                    }
#pragma unroll
                    for (int g = 0; g < H; ++g)
#pragma unroll
                        for (int rs = 0; rs < 2; ++rs) l[g][rs] += ex2(y[g][rs * 2]) + ex2(y[g][rs * 2 + 1]);
                }
                tok_acquire();
                if (t + 1 < T) load_mix(t + 1, false);
                tok_release();
            }
            // combine: the four threads of a quad, then the two column-group warps of the quarter
#pragma unroll
            for (int g = 0; g < H; ++g)
#pragma unroll
                for (int rs = 0; rs < 2; ++rs) {
#pragma unroll
                    for (int o = 1; o <= 2; o <<= 1) {
                        const float m2 = __shfl_xor_sync(0xffffffffu, m[g][rs], o), l2 = __shfl_xor_sync(0xffffffffu, l[g][rs], o);
                        stat_merge(m[g][rs], l[g][rs], m2, l2);
                    }
                }
            float2* xc = reinterpret_cast<float2*>(smem + SM::LSE);                 // [NCG - 1][64][H]
            if (cg > 0 && tig == 0) {
#pragma unroll
                for (int g = 0; g < H; ++g)
#pragma unroll
                    for (int rs = 0; rs < 2; ++rs) xc[((cg - 1) * RB + quarter * 16 + rs * 8 + gid) * H + g] = make_float2(m[g][rs], l[g][rs]);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NPW * 32) : "memory");
            if (cg == 0 && tig == 0) {
#pragma unroll
                for (int g = 0; g < H; ++g)
#pragma unroll
                    for (int rs = 0; rs < 2; ++rs) {
                        for (int c = 0; c < NCG - 1; ++c) {
                            const float2 o = xc[(c * RB + quarter * 16 + rs * 8 + gid) * H + g];
                            stat_merge(m[g][rs], l[g][rs], o.x, o.y);
                        }
                        p.part[(((long long)b * p.nchunk + chunk) * H + g) * p.Npad + row_a + rs * 8] = make_float2(m[g][rs], l[g][rs]);
                    }
            }
        } else {
            uint32_t ww[(H / 4) * (H / 2)][2];
            load_wfrag<H, false, false>(wsm + H * H, 1.f, lane, ww);
            // c1[g][rs] = log2e bl[g] - lse2 + P_SHIFT (the first mix then yields the exponent of 2^8 P directly);  c2[g] = 2^8 bw[g]
            float c1[H][2], c2[H];
            const float* ls = reinterpret_cast<const float*>(smem + SM::LSE);
#pragma unroll
            for (int g = 0; g < H; ++g) {
                c2[g] = wsm[2 * H * H + H + g];
#pragma unroll
                for (int rs = 0; rs < 2; ++rs) c1[g][rs] = wsm[2 * H * H + g] - ls[g * RB + quarter * 16 + rs * 8 + gid] + P_SHIFT;
            }
            // The two warps of a lane quarter (= one SM sub-partition: one mma.sync pipe, one MUFU pipe) take ALTERNATE blocks: each
            // waits on its own barriers, so the pair drifts out of phase and one warp's exponentials run under the other's mixes.
            // (Sharing every block, the pair re-aligns at each block barrier and the pipes take turns idling: 3100 clocks of a 3750
            //  clock block period were the serialised sum of the pipe times; a token ping-pong and a hand-interleaved software
            //  pipeline were both measured slower: the accumulating tcgen05.mma stream competes for the tensor cores.)
            tok_acquire();
            tok_release();
            constexpr int NSB = CB / 8;                           // 8-column sub-tiles per block
            for (int jl = cg; jl < nb; jl += 2) {
                const uint32_t sb = (uint32_t)jl & 1u, ab = (uint32_t)jl % SM::NABUF;
                if (warp == 2) TF_TRACE(2, jl, 0);
                mbar_wait(sfull0 + 8 * sb, ((uint32_t)jl >> 1) & 1u);
                mbar_wait(aempty0 + 8 * ab, (((uint32_t)jl / SM::NABUF) & 1u) ^ 1u);
                fence_after();
                if (warp == 2) TF_TRACE(2, jl, 1);
#pragma unroll 1
                for (int sub = 0; sub < NSB; sub += 2) {
                    float xs[2][H][4];
#pragma unroll
                    for (int u = 0; u < 2; ++u) ld_tiles<H, CB>(tS + tl + ((sb * 16u) << 16) + (uint32_t)((sub + u) * 8), xs[u]);
                    ld_wait();
                    if (sub + 2 == NSB) {
                        fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(sempty0 + 8 * sb);
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        float y[H][4];
                        head_mix<H, false>(xs[u], wl, [&](int g, int i) { return c1[g][i >> 1]; }, y);
#pragma unroll
                        for (int g = 0; g < H; ++g)
#pragma unroll
                            for (int i = 0; i < 4; ++i) y[g][i] = ex2(y[g][i]);
                        head_mix<H, false>(y, ww, [&](int g, int) { return c2[g]; }, xs[u]);
                        st_tiles<H, CB>(sbase + SM::A + ab * SM::A_B, quarter, sub + u, lane, xs[u]);
                    }
                }
                if (warp == 2) TF_TRACE(2, jl, 2);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(afull0 + 8 * ab);
            }
            // ---- output: lanes [0, 16) of the quarter hold O columns [0, O_CAP), lanes [16, 32) the columns [O_CAP, D) of the same
            // 16 rows; each column-group warp takes half of the TMEM columns
            mbar_wait(ofull, 0);
            fence_after();
            const int row = r0 + quarter * 16 + (lane & 15);
            const float inv = 1.f / 256.f;
            constexpr int OC = D < (int)SM::O_CAP ? D : (int)SM::O_CAP;     // TMEM columns in use
#pragma unroll 1
            for (int c = cg * (OC / NCG); c < (cg + 1) * (OC / NCG); c += 16) {
                uint32_t o[16];
                UMMA_LD_32x32_X16(tO + tl + (uint32_t)c, o);
                ld_wait();
                const int col = lane < 16 ? c : (int)SM::O_CAP + c;
                if (row < p.N && col < D) {
                    if (p.out32 != nullptr) {
                        float* dst = p.out32 + ((long long)b * p.N + row) * D + col;
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + q4 * 4), "f"(__uint_as_float(o[q4 * 4]) * inv),
                                         "f"(__uint_as_float(o[q4 * 4 + 1]) * inv), "f"(__uint_as_float(o[q4 * 4 + 2]) * inv),
                                         "f"(__uint_as_float(o[q4 * 4 + 3]) * inv) : "memory");
                    } else {
                        uint16_t* dst = p.out + (long long)b * p.out_sb + (long long)row * p.out_ld + col;
                        uint4 v0, v1;
                        v0.x = pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                        v0.y = pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                        v0.z = pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
                        v0.w = pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
                        v1.x = pack_bf16(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
                        v1.y = pack_bf16(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
                        v1.z = pack_bf16(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
                        v1.w = pack_bf16(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
                        *reinterpret_cast<uint4*>(dst) = v0;
                        *reinterpret_cast<uint4*>(dst + 8) = v1;
                    }
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(SM::TMEM_COLS) : "memory");
}

// lse2[b][g][row] = log2-domain logsumexp merged over the column chunks of the stats kernel
__global__ void tf_merge_kernel(const float2* __restrict__ part, float* __restrict__ lse2, int B, int H, int N, int Npad, int nchunk) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * H * N) return;
    const int row = (int)(i % N);
    const int g = (int)((i / N) % H);
    const int b = (int)(i / ((long long)N * H));
    float m = -1e30f, l = 0.f;
    for (int c = 0; c < nchunk; ++c) {
        const float2 o = part[(((long long)b * nchunk + c) * H + g) * Npad + row];
        stat_merge(m, l, o.x, o.y);
    }
    lse2[i] = m + log2f(l);
}

__global__ void tf_cast_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long rows, int D, long long dst_ld, int N, long long dst_sb) {
    // src [B*N, D] f32 contiguous -> dst bf16 view; 4 elements per thread
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= rows * D) return;
    const long long r = i / D;
    const int c = (int)(i - r * D);
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    const long long bimg = r / N, n = r - bimg * N;
    uint2 o = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    *reinterpret_cast<uint2*>(dst + bimg * dst_sb + n * dst_ld + c) = o;
}

// chunks of the streamed range per row block: fill the SMs (waves x per-CTA time), ~`fixed` blocks' worth of per-CTA prologue / epilogue
int pick_nchunk(int ctas0, int nblk, double fixed) {
    static const char* env = getenv("SPE_TF_NCHUNK");
    if (env) { int v = atoi(env); if (v >= 1) return v > nblk ? nblk : v; }
    const int sms = spe_num_sms();
    int best = 1;
    double best_cost = 1e30;
    for (int nc = 1; nc <= 8 && nc <= nblk; ++nc) {
        const int bpc = (nblk + nc - 1) / nc;
        const int real_nc = (nblk + bpc - 1) / bpc;
        const double waves = (double)((ctas0 * real_nc + sms - 1) / sms);
        const double cost = waves * (bpc + fixed);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = real_nc; }
    }
    return best;
}

long long* g_tf_trace = nullptr;

template <int H>
int launch_fwd(const spe_talking_fused_args* a, cudaStream_t st) {
    using SMS = FwdSmem<H, CB_STATS, true>;
    using SMM = FwdSmem<H, CB_MAIN, false>;
    const int D = H * DHD, N = a->N, B = a->B;
    const int nrb = (N + RB - 1) / RB;
    const int Npad = nrb * RB;
    const int nblk_s = (N + CB_STATS - 1) / CB_STATS, nblk_m = (N + CB_MAIN - 1) / CB_MAIN;
    int nc_s = pick_nchunk(nrb * B, nblk_s, 3.0), nc_m = pick_nchunk(nrb * B, nblk_m, 4.0);
    const int bpc_s = (nblk_s + nc_s - 1) / nc_s, bpc_m = (nblk_m + nc_m - 1) / nc_m;
    nc_s = (nblk_s + bpc_s - 1) / bpc_s;
    nc_m = (nblk_m + bpc_m - 1) / bpc_m;
    const size_t part_b = (((size_t)B * nc_s * H * Npad * sizeof(float2)) + 255) / 256 * 256;
    const size_t o32_b = nc_m > 1 ? (size_t)B * N * D * sizeof(float) : 0;
    SPE_CHECK(a->workspace && (size_t)a->workspace_bytes >= part_b + o32_b, "spe_talking_fused_fwd: workspace too small");
    CUtensorMap tq, tks, tkm, tv;
    if (make_map(&tq, a->q, D, N, B, a->q_ld, a->q_sb, RB)) return -1;
    if (make_map(&tks, a->k, D, N, B, a->k_ld, a->k_sb, CB_STATS)) return -1;
    if (make_map(&tkm, a->k, D, N, B, a->k_ld, a->k_sb, CB_MAIN)) return -1;
    if (make_map(&tv, a->v, D, N, B, a->v_ld, a->v_sb, CB_MAIN)) return -1;
    TfFwdParams p;
    p.N = N; p.Npad = Npad;
    p.Wl = a->Wl; p.bl = a->bl; p.Ww = a->Ww; p.bw = a->bw; p.scale = a->scale;
    p.part = reinterpret_cast<float2*>(a->workspace);
    p.lse2 = a->lse2;
    p.out = reinterpret_cast<uint16_t*>(a->out); p.out_ld = a->out_ld; p.out_sb = a->out_sb;
    p.out32 = nc_m > 1 ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a->workspace) + part_b) : nullptr;
    static const char* dbg_env = getenv("SPE_TF_DBG");
    p.dbg = dbg_env ? atoi(dbg_env) : 0;
    if ((p.dbg & 512) && !g_tf_trace) { cudaMalloc(&g_tf_trace, 4 * 32 * 8 * 8); cudaMemset(g_tf_trace, 0, 4 * 32 * 8 * 8); }
    p.trace = nullptr;
    static bool attr_done = false;
    if (!attr_done) {
        SPE_CUDA(cudaFuncSetAttribute(tf_fwd_kernel<H, CB_STATS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMS::TOTAL));
        SPE_CUDA(cudaFuncSetAttribute(tf_fwd_kernel<H, CB_MAIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMM::TOTAL));
        attr_done = true;
    }
    const double pos = (double)B * N * N;
    {
        SpeProfScope ps(SPE_FAM_TALKING_FWD, pos * H * 4.0, st, "tf_stats");
        p.nblk = nblk_s; p.nchunk = nc_s; p.bpc = bpc_s;
        tf_fwd_kernel<H, CB_STATS, true><<<dim3(nrb, nc_s, B), TF_THREADS, SMS::TOTAL, st>>>(tq, tks, tks, p);
        SPE_LAUNCHED();
        const long long tot = (long long)B * H * N;
        tf_merge_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(p.part, p.lse2, B, H, N, Npad, nc_s);
        SPE_LAUNCHED();
    }
    if (p.out32) SPE_CUDA(cudaMemsetAsync(p.out32, 0, o32_b, st));
    {
        SpeProfScope ps(SPE_FAM_TALKING_FWD, pos * H * 4.0, st, "tf_main");
        p.nblk = nblk_m; p.nchunk = nc_m; p.bpc = bpc_m;
        p.trace = (p.dbg & 512) ? g_tf_trace : nullptr;
        tf_fwd_kernel<H, CB_MAIN, false><<<dim3(nrb, nc_m, B), TF_THREADS, SMM::TOTAL, st>>>(tq, tkm, tv, p);
        SPE_LAUNCHED();
    }
    if (p.out32) {
        const long long tot = (long long)B * N * D / 4;
        tf_cast_bf16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(p.out32, p.out, (long long)B * N, D, a->out_ld, N, a->out_sb);
        SPE_LAUNCHED();
    }
    return 0;
}

// parameter gradients: accumulate (+=) the launch's sums into the caller's gradient buffers
__global__ void tf_addw_kernel(const float* __restrict__ dw, float* __restrict__ dWw, float* __restrict__ dWl, int n) {
    const int i = threadIdx.x;
    if (i < n) dWw[i] += dw[i];
    else if (i < 2 * n) dWl[i - n] += dw[i];
}

// =============================================================================================================================
// Backward.  With p = softmax(L) (per mixed head g), A = Ww p + bw, O = A V:
//   dA_g = dO_g V_g^T,  dP_h = sum_g Ww[g,h] dA_g,  delta_h = rowsum(p_h dP_h),  dL_h = p_h (dP_h - delta_h),
//   dS_h = sum_g Wl[g,h] dL_g,  dQ_h = scale dS_h K_h,  dK_h = scale dS_h^T Q_h,  dV_g = A_g^T dO_g,
//   dWw[g,h] = sum dA_g p_h,  dWl[g,h] = sum dL_g S_h      (dbl = 0: softmax is shift invariant; dbw is exact outside)
// Three launches, all recomputing S and p from q, k and the saved lse2 (no N^2 tensor in HBM):
//   DELTA (query-stationary): delta, dWw          DQ (query-stationary): dQ, dWl          DKV (key-stationary): dK, dV
// Same tile geometry as the forward with 16-column blocks: the S tile sits in lane half 0 and the dA tile in lane half 1 of the
// same H x 16 TMEM columns (one tcgen05.ld.16x256b per head and tile gives a thread S and dA of the same four positions); the
// 64 x 384 accumulators (dQ | dK and dV) use the two lane halves of the remaining 384 columns.  The parameter gradients are
// tcgen05 products as well: D[(g, c), (h, c')] = sum_rows dL_g[row, c] S_h[row, c'] from the bf16 tiles read MN-major; only the
// c = c' diagonal is kept (8x redundant tensor work, but no 64-accumulator FMA loop in the position threads).
// Probabilities are carried as p' = 2^8 p (see the forward); every result that is linear in p' is scaled by 2^-8 when it leaves.
// =============================================================================================================================
enum { BW_DELTA = 0, BW_DQ = 1, BW_DKV = 2 };
constexpr int CBB = 16;

struct TfBwdParams {
    int N, nblk, nchunk, bpc, Npad;
    const float *Wl, *bl, *Ww, *bw;
    float scale;
    const float* lse2;                   // [B][H][N]
    float* delta;                        // [B][H][N]   DELTA: += (zeroed by the caller)   DQ / DKV: read
    float* dwacc;                        // [H*H]       DELTA: dWw +=    DQ: dWl +=
    float* out0;                         // f32 [B][N][D]  DQ: dq +=    DKV: dk +=
    float* out1;                         //                              DKV: dv +=
    long long* trace;
};

template <int H, int MODE>
struct BwdSmem {
    static constexpr int D = H * DHD, NT = D / 64;
    static constexpr int NTILE = MODE == BW_DQ ? 3 : 2;
    static constexpr int NSTG = MODE == BW_DQ ? 2 : 3;                 // streamed-block ring (TMA latency ~1700 clocks against a ~2500 clock block)
    static constexpr uint32_t YT = CBB * 128;
    static constexpr uint32_t X1 = 0, X2 = X1 + NT * XT_B;
    static constexpr uint32_t Y1 = X2 + NT * XT_B, Y2 = Y1 + NSTG * NT * YT;
    static constexpr uint32_t T_B = 8 * H * (CBB / 8) * 128;           // one bf16 tile [8 row groups][H][2][8 rows][16 B]
    static constexpr uint32_t T0 = Y2 + NSTG * NT * YT;
    static constexpr uint32_t BAR = T0 + NTILE * T_B;
    static constexpr int NBAR = 1 + 2 * NSTG + 5;
    static constexpr uint32_t TSLOT = BAR + NBAR * 8;
    static constexpr uint32_t WSM = TSLOT + 16;                        // Wl' [H*H], Ww [H*H], bl' [H], bw' [H]
    static constexpr uint32_t RC = WSM + (2 * H * H + 2 * H) * 4;      // row constants: lse2 [H][64], delta [H][64]
    static constexpr uint32_t FR = RC + 2 * H * RB * 4;                // packed B fragments of Ww (f16) and scale Wl^T (bf16): [2][H*H/4][32 lanes] words
    static constexpr uint32_t TOTAL = FR + 2 * (H * H / 4) * 32 * 4 + 1024;
    static constexpr uint32_t S_COLS = H * CBB;
};

// D[(g, c), (h, c')] accumulator (64 x 64, f32) -> dst[g * H + h] += mul * sum_c D[(g, c), (h, c)].  `half`: TMEM lane half.
template <int H>
__device__ __forceinline__ void trick_epilogue(uint32_t taddr_q, int quarter, int lane, int half, float* dst, float mul) {
    constexpr int CW = 64 / H;                                          // columns per head in the trick tile (8 or 16)
    const int m = quarter * 16 + (lane & 15), gr = m / CW, c = m % CW;
    float v[H];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t r[16];
        UMMA_LD_32x32_X16(taddr_q + (uint32_t)(k * 16), r);
        ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int n = k * 16 + i;                                   // column (h, c')
            if (n % CW == 0) v[n / CW] = 0.f;
            if ((n % CW) == c) v[n / CW] = __uint_as_float(r[i]);
        }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
#pragma unroll
        for (int o = 1; o < CW; o <<= 1) v[h] += __shfl_xor_sync(0xffffffffu, v[h], o);
        if (c == 0 && (lane >> 4) == half) atomicAdd(dst + gr * H + h, v[h] * mul);
    }
}

template <int H, int MODE>
__global__ void __launch_bounds__(TF_THREADS, 1) tf_bwd_kernel(const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
                                                               const __grid_constant__ CUtensorMap tmY1, const __grid_constant__ CUtensorMap tmY2,
                                                               const TfBwdParams p) {
    using SM = BwdSmem<H, MODE>;
    constexpr int D = SM::D, NT = SM::NT, NSTG = SM::NSTG;
    constexpr uint32_t YT = SM::YT;
    constexpr bool KEYST = MODE == BW_DKV;                            // key-stationary: rows = keys, columns = queries
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + SM::BAR;
    const uint32_t xfull = bar0, yfull0 = xfull + 8, yempty0 = yfull0 + 8 * NSTG, sfull = yempty0 + 8 * NSTG, sempty = sfull + 8, tfull = sempty + 8,
                   tempty = tfull + 8, accfull = tempty + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::TSLOT);
    float* wsm = reinterpret_cast<float*>(smem + SM::WSM);
    float* rc = reinterpret_cast<float*>(smem + SM::RC);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * RB, chunk = blockIdx.y, b = blockIdx.z;
    const int blk0 = chunk * p.bpc;
    const int nb = min(p.bpc, p.nblk - blk0);

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY1)) : "memory");
        mbar_init(xfull, 1);
        for (int s = 0; s < NSTG; ++s) { mbar_init(yfull0 + 8 * s, 1); mbar_init(yempty0 + 8 * s, 1); }
        mbar_init(sfull, 1); mbar_init(sempty, 8); mbar_init(tfull, 8); mbar_init(tempty, 1); mbar_init(accfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {
        const int t = threadIdx.x - 64;
        if (t < H * H) { wsm[t] = p.Wl[t] * (LOG2E * p.scale); wsm[H * H + t] = p.Ww[t]; }
        if (t < H) { wsm[2 * H * H + t] = p.bl[t] * LOG2E; wsm[2 * H * H + H + t] = p.bw[t] * 256.f; }
        if (!KEYST) {
            for (int i = t; i < H * RB; i += 256) {
                const int g = i / RB, r = i % RB;
                const bool ok = r0 + r < p.N;
                rc[i] = ok ? p.lse2[((long long)b * H + g) * p.N + r0 + r] : 0.f;
                rc[H * RB + i] = (ok && MODE == BW_DQ) ? p.delta[((long long)b * H + g) * p.N + r0 + r] : 0.f;
            }
        }
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tT0 = tmem_base, tT1 = tmem_base + (16u << 16);               // S tile | dA tile
    const uint32_t tA0 = tmem_base + SM::S_COLS, tA1 = tA0 + (16u << 16);         // accumulators: lane half 0 | 1

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (elect()) {
            mbar_expect_tx(xfull, 2 * NT * XT_B);
            for (int t = 0; t < NT; ++t) tma_load_3d(sbase + SM::X1 + t * XT_B, &tmX1, xfull, t * 64, r0, b);
            for (int t = 0; t < NT; ++t) tma_load_3d(sbase + SM::X2 + t * XT_B, &tmX2, xfull, t * 64, r0, b);
        }
        __syncwarp();
        for (int jl = 0; jl < nb; ++jl) {
            const int s = jl % NSTG;
            const int c0 = (blk0 + jl) * CBB;
            mbar_wait(yempty0 + 8 * s, ((uint32_t)(jl / NSTG) & 1u) ^ 1u);
            if (elect()) {
                mbar_expect_tx(yfull0 + 8 * s, 2 * NT * YT);
                for (int t = 0; t < NT; ++t) tma_load_3d(sbase + SM::Y1 + (s * NT + t) * YT, &tmY1, yfull0 + 8 * s, t * 64, c0, b);
                for (int t = 0; t < NT; ++t) tma_load_3d(sbase + SM::Y2 + (s * NT + t) * YT, &tmY2, yfull0 + 8 * s, t * 64, c0, b);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t ID_S = idesc_bf16(RB, CBB, false, false);
        constexpr uint32_t ID_TRICK = idesc_bf16(64, 64, true, true);
        constexpr int CW = 64 / H;                                      // columns per head in a parameter-gradient product
        const uint64_t dX1 = desc(sbase + SM::X1, 0, 1024, 2), dX2 = desc(sbase + SM::X2, 0, 1024, 2);
        auto issue_tiles = [&](int jl) {                                 // S -> lane half 0, dA -> lane half 1
            const int s = jl % NSTG;
            const uint64_t dY1 = desc(sbase + SM::Y1 + s * NT * YT, 0, 1024, 2), dY2 = desc(sbase + SM::Y2 + s * NT * YT, 0, 1024, 2);
#pragma unroll
            for (int h = 0; h < H; ++h)
#pragma unroll
                for (int ks = 0; ks < DHD / 16; ++ks) {
                    const int e = h * DHD + ks * 16;
                    const uint64_t oa = (uint64_t)(((e >> 6) * XT_B + (e & 63) * 2) >> 4), ob = (uint64_t)(((e >> 6) * YT + (e & 63) * 2) >> 4);
                    mma_f16(tT0 + (uint32_t)(h * CBB), dX1 + oa, dY1 + ob, ID_S, ks != 0);
                    mma_f16(tT1 + (uint32_t)(h * CBB), dX2 + oa, dY2 + ob, ID_S, ks != 0);
                }
        };
        // acc[:, head g] += tile_g (K-major, one k-step) x Y_g (MN-major inside the 64-column tiles of the streamed block)
        auto issue_acc = [&](uint32_t tile, uint32_t ybase, uint32_t tacc, bool accum) {
            const uint64_t dT = desc(tile, 128, H * (CBB / 8) * 128, 0);
            const uint64_t dY = desc(ybase, YT, 1024, 2);
#pragma unroll
            for (int g = 0; g < H; ++g)
#pragma unroll
                for (int seg = 0; seg < 2; ++seg) {
                    const int e0 = g * DHD;
                    const int n0 = (64 - (e0 & 63)) < DHD ? (64 - (e0 & 63)) : DHD;
                    const int e = seg == 0 ? e0 : e0 + n0, n = seg == 0 ? n0 : DHD - n0;
                    if (n > 0)
                        mma_f16(tacc + (uint32_t)e, dT + (uint64_t)((g * (CBB / 8) * 128) >> 4), dY + (uint64_t)(((e >> 6) * YT + (e & 63) * 2) >> 4),
                                idesc_bf16(RB, n, false, true), accum);
                }
        };
        // parameter-gradient product: D[(g, c), (h, c')] += sum_rows TA_g[row, c] TB_h[row, c'] (both tiles read MN-major, K = rows)
        auto issue_trick = [&](uint32_t ta, uint32_t tb, uint32_t tacc, bool accum) {
#pragma unroll
            for (int grp = 0; grp < CBB / CW; ++grp) {
                const uint64_t dA = desc(ta + grp * 128, 8 * H * (CBB / 8) * 16, (CW == 8 ? (CBB / 8) * 128 : 128), 0);
                const uint64_t dB = desc(tb + grp * 128, 8 * H * (CBB / 8) * 16, (CW == 8 ? (CBB / 8) * 128 : 128), 0);
#pragma unroll
                for (int ks = 0; ks < RB / 16; ++ks) {
                    const uint64_t o = (uint64_t)((ks * 2 * H * (CBB / 8) * 128) >> 4);
                    mma_f16(tacc, dA + o, dB + o, ID_TRICK, accum || grp != 0 || ks != 0);
                }
            }
        };
        mbar_wait(xfull, 0);
        mbar_wait(yfull0, 0);
        fence_after();
        if (elect()) { issue_tiles(0); commit(sfull); }
        __syncwarp();
        for (int jl = 0; jl < nb; ++jl) {
            const int s = jl % NSTG;
            TF_TRACE(1, jl, 0);
            if (jl + 1 < nb) {
                mbar_wait(yfull0 + 8 * ((jl + 1) % NSTG), (uint32_t)((jl + 1) / NSTG) & 1u);
                TF_TRACE(1, jl, 1);
                mbar_wait(sempty, (uint32_t)jl & 1u);                    // both tiles of block jl are in registers
                TF_TRACE(1, jl, 2);
                fence_after();
                if (elect()) { issue_tiles(jl + 1); commit(sfull); }
                __syncwarp();
                TF_TRACE(1, jl, 3);
            }
            mbar_wait(tfull, (uint32_t)jl & 1u);
            TF_TRACE(1, jl, 4);
            fence_after();
            if (elect()) {
                const uint32_t t0 = sbase + SM::T0, t1 = t0 + SM::T_B, t2 = t1 + SM::T_B;
                const uint32_t y1 = sbase + SM::Y1 + s * NT * YT, y2 = sbase + SM::Y2 + s * NT * YT;
                if (MODE == BW_DELTA) issue_trick(t1, t0, tA0, jl != 0);                       // dWw: dA (x) p
                if (MODE == BW_DQ) { issue_acc(t0, y1, tA0, jl != 0); issue_trick(t1, t2, tA1, jl != 0); }     // dQ += dS K ; dWl: dL (x) S
                if (MODE == BW_DKV) { issue_acc(t0, y1, tA0, jl != 0); issue_acc(t1, y2, tA1, jl != 0); }      // dK += dS^T Q ; dV += A^T dO
                commit(tempty);
                commit(yempty0 + 8 * s);
            }
            __syncwarp();
            TF_TRACE(1, jl, 5);
        }
        if (elect()) commit(accfull);
        __syncwarp();
    } else {
        // ---------------- position warps: quarter = lane quarter, cg = 8-column half of the block ----------------
        const int quarter = warp & 3, cg = (warp - 2) >> 2;
        const int gid = lane >> 2, tig = lane & 3;
        const uint32_t tl = (uint32_t)(quarter * 32) << 16;
        uint32_t wl[(H / 4) * (H / 2)][2], wwT[(H / 4) * (H / 2)][2];
        load_wfrag<H, false, false>(wsm, 1.f, lane, wl);                               // log2e scale Wl          (f16)
        load_wfrag<H, true, true>(wsm + H * H, 1.f, lane, wwT);                        // Ww^T: dP = Ww^T dA      (bf16)
        // the two remaining fragment sets are kept packed in shared memory (register budget) and re-read per block: 16 LDS each
        uint32_t* frs = reinterpret_cast<uint32_t*>(smem + SM::FR);
        constexpr int NFR = H * H / 4;                                                  // words per set and lane
        if (warp == 2) {
            uint32_t f0[(H / 4) * (H / 2)][2], f1[(H / 4) * (H / 2)][2];
            load_wfrag<H, false, false>(wsm + H * H, 1.f, lane, f0);                   // Ww                      (f16)
            load_wfrag<H, true, true>(wsm, 1.f / LOG2E, lane, f1);                     // scale Wl^T: dS = Wl^T dL (bf16)
#pragma unroll
            for (int i = 0; i < NFR; ++i) { frs[i * 32 + lane] = f0[i >> 1][i & 1]; frs[(NFR + i) * 32 + lane] = f1[i >> 1][i & 1]; }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        auto load_frs = [&](int set, uint32_t (&f)[(H / 4) * (H / 2)][2]) {
#pragma unroll
            for (int i = 0; i < NFR; ++i) f[i >> 1][i & 1] = frs[(set * NFR + i) * 32 + lane];
        };
        float c1[H][2], dl_[H][2];                                                     // query-stationary: per row
        float dsum[H][2];
#pragma unroll
        for (int g = 0; g < H; ++g)
#pragma unroll
            for (int rs = 0; rs < 2; ++rs) {
                c1[g][rs] = wsm[2 * H * H + g] - (KEYST ? 0.f : rc[g * RB + quarter * 16 + rs * 8 + gid]) + P_SHIFT;
                dl_[g][rs] = KEYST ? 0.f : rc[H * RB + g * RB + quarter * 16 + rs * 8 + gid];
                dsum[g][rs] = 0.f;
            }
        for (int jl = 0; jl < nb; ++jl) {
            if (warp == 2) TF_TRACE(2, jl, 0);
            mbar_wait(sfull, (uint32_t)jl & 1u);
            if (warp == 2) TF_TRACE(2, jl, 1);
            fence_after();
            float x[H][4], da[H][4];
            ld_tiles<H, CBB>(tT0 + tl + (uint32_t)(cg * 8), x);
            ld_tiles<H, CBB>(tT1 + tl + (uint32_t)(cg * 8), da);
            ld_wait();
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sempty);
            if (warp == 2) TF_TRACE(2, jl, 2);
            // per-column constants (key-stationary): lse2 and delta of the thread's two query columns
            float cl[H][2], cd[H][2];
            if (KEYST) {
                const int col = (blk0 + jl) * CBB + cg * 8 + 2 * tig;
#pragma unroll
                for (int g = 0; g < H; ++g)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const bool ok = col + j < p.N;
                        const long long idx = ((long long)b * H + g) * p.N + col + j;
                        cl[g][j] = c1[g][0] - (ok ? p.lse2[idx] : 0.f);
                        cd[g][j] = ok ? p.delta[idx] : 0.f;
                    }
            }
            mbar_wait(tempty, ((uint32_t)jl & 1u) ^ 1u);                 // the previous block's tiles have been consumed
            if (warp == 2) TF_TRACE(2, jl, 3);
            const uint32_t t0 = sbase + SM::T0, t1 = t0 + SM::T_B, t2 = t1 + SM::T_B;
            if (MODE == BW_DQ) st_tiles<H, CBB>(t2, quarter, cg, lane, x);                   // S (bf16) for dWl
            if (MODE == BW_DELTA) st_tiles<H, CBB>(t1, quarter, cg, lane, da);               // dA (bf16) for dWw
            float y[H][4];
            if (KEYST) head_mix<H, false>(x, wl, [&](int g, int i) { return cl[g][i & 1]; }, y);
            else head_mix<H, false>(x, wl, [&](int g, int i) { return c1[g][i >> 1]; }, y);
#pragma unroll
            for (int g = 0; g < H; ++g)
#pragma unroll
                for (int i = 0; i < 4; ++i) y[g][i] = ex2(y[g][i]);                      // p' = 2^8 p
            if (MODE == BW_DKV) {
                uint32_t ww[(H / 4) * (H / 2)][2];
                load_frs(0, ww);
                head_mix<H, false>(y, ww, [&](int g, int) { return wsm[2 * H * H + H + g]; }, x);      // A' = Ww p' + 2^8 bw
                st_tiles<H, CBB>(t1, quarter, cg, lane, x);
            }
            head_mix<H, true>(da, wwT, [](int, int) { return 0.f; }, x);                // dP
            if (MODE == BW_DELTA) {
                st_tiles<H, CBB>(t0, quarter, cg, lane, y);                               // p' (bf16) for dWw
#pragma unroll
                for (int g = 0; g < H; ++g)
#pragma unroll
                    for (int rs = 0; rs < 2; ++rs) dsum[g][rs] += y[g][rs * 2] * x[g][rs * 2] + y[g][rs * 2 + 1] * x[g][rs * 2 + 1];
            } else {
#pragma unroll
                for (int g = 0; g < H; ++g)
#pragma unroll
                    for (int i = 0; i < 4; ++i) y[g][i] *= x[g][i] - (KEYST ? cd[g][i & 1] : dl_[g][i >> 1]);      // dL' = p' (dP - delta)
                if (MODE == BW_DQ) st_tiles<H, CBB>(t1, quarter, cg, lane, y);           // dL' (bf16) for dWl
                uint32_t wlT[(H / 4) * (H / 2)][2];
                if (MODE == BW_DQ) load_wfrag<H, true, true>(wsm, 1.f / LOG2E, lane, wlT);
                else load_frs(1, wlT);                                                   // scale Wl^T: dS = Wl^T dL      (bf16)
                head_mix<H, true>(y, wlT, [](int, int) { return 0.f; }, x);
                st_tiles<H, CBB>(t0, quarter, cg, lane, x);                               // scale dS' (bf16)
            }
            if (warp == 2) TF_TRACE(2, jl, 4);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(tfull);
            if (warp == 2) TF_TRACE(2, jl, 5);
        }
        // ---------------- epilogue ----------------
        const float inv = 1.f / 256.f;
        if (MODE == BW_DELTA) {
#pragma unroll
            for (int g = 0; g < H; ++g)
#pragma unroll
                for (int rs = 0; rs < 2; ++rs) {
                    float v = dsum[g][rs];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    const int row = r0 + quarter * 16 + rs * 8 + gid;
                    if (tig == 0 && row < p.N) atomicAdd(p.delta + ((long long)b * H + g) * p.N + row, v * inv);
                }
        }
        mbar_wait(accfull, 0);
        fence_after();
        if (MODE == BW_DELTA && cg == 0) trick_epilogue<H>(tA0 + tl, quarter, lane, 0, p.dwacc, inv);
        if (MODE == BW_DQ && cg == 0) trick_epilogue<H>(tA0 + tl, quarter, lane, 1, p.dwacc, inv * p.scale);
        if (MODE != BW_DELTA) {
            // lanes [0, 16): accumulator 0 (dQ | dK), lanes [16, 32): accumulator 1 (dV) of the same 16 rows; each warp half of the columns
            const int row = r0 + quarter * 16 + (lane & 15);
            float* dst0 = (lane < 16 ? p.out0 : p.out1);
            const bool on = row < p.N && (lane < 16 || MODE == BW_DKV);
#pragma unroll 1
            for (int c = cg * (D / 2); c < (cg + 1) * (D / 2); c += 16) {
                uint32_t o[16];
                UMMA_LD_32x32_X16(tA0 + tl + (uint32_t)c, o);
                ld_wait();
                if (on) {
                    float* dst = dst0 + ((long long)b * p.N + row) * D + c;
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + q4 * 4), "f"(__uint_as_float(o[q4 * 4]) * inv),
                                     "f"(__uint_as_float(o[q4 * 4 + 1]) * inv), "f"(__uint_as_float(o[q4 * 4 + 2]) * inv),
                                     "f"(__uint_as_float(o[q4 * 4 + 3]) * inv) : "memory");
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// dq | dk | dv (f32 [3][B*N][D]) -> packed bf16 dqkv [B, N, 3 D]
__global__ void tf_pack_dqkv_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long rows, int D, long long dst_ld) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= 3 * rows * D) return;
    const long long which = i / (rows * D), rem = i - which * rows * D;
    const long long r = rem / D;
    const int c = (int)(rem - r * D);
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    *reinterpret_cast<uint2*>(dst + r * dst_ld + which * D + c) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
}

int g_bwd_trace_mode = -1;

template <int H, int MODE>
int launch_bwd_one(const CUtensorMap& x1, const CUtensorMap& x2, const CUtensorMap& y1, const CUtensorMap& y2, TfBwdParams p, int nrb, int B, cudaStream_t st,
                   const char* tag, double work) {
    using SM = BwdSmem<H, MODE>;
    static bool attr_done = false;
    if (!attr_done) {
        SPE_CUDA(cudaFuncSetAttribute(tf_bwd_kernel<H, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::TOTAL));
        attr_done = true;
    }
    p.trace = (g_bwd_trace_mode == MODE) ? g_tf_trace : nullptr;
    SpeProfScope ps(SPE_FAM_TALKING_BWD, work, st, tag);
    tf_bwd_kernel<H, MODE><<<dim3(nrb, p.nchunk, B), TF_THREADS, SM::TOTAL, st>>>(x1, x2, y1, y2, p);
    SPE_LAUNCHED();
    return 0;
}

template <int H>
int launch_bwd(const spe_talking_fused_bwd_args* a, cudaStream_t st) {
    const int D = H * DHD, N = a->N, B = a->B;
    const int nrb = (N + RB - 1) / RB, nblk = (N + CBB - 1) / CBB;
    int nc = pick_nchunk(nrb * B, nblk, 6.0);
    const int bpc = (nblk + nc - 1) / nc;
    nc = (nblk + bpc - 1) / bpc;
    const size_t delta_b = (((size_t)B * H * N * 4) + 255) / 256 * 256, dw_b = 256 * 4, acc_b = (size_t)3 * B * N * D * 4;
    SPE_CHECK(a->workspace && (size_t)a->workspace_bytes >= delta_b + dw_b + acc_b, "spe_talking_fused_bwd: workspace too small");
    uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
    float* delta = reinterpret_cast<float*>(ws);
    float* dw = reinterpret_cast<float*>(ws + delta_b);                 // [2][H*H]: dWw, dWl
    float* acc = reinterpret_cast<float*>(ws + delta_b + dw_b);        // [3][B*N][D]: dq, dk, dv
    SPE_CUDA(cudaMemsetAsync(ws, 0, delta_b + dw_b + acc_b, st));
    CUtensorMap q64, k64, v64, do64, q16, k16, v16, do16;
    if (make_map(&q64, a->q, D, N, B, a->q_ld, a->q_sb, RB) || make_map(&k64, a->k, D, N, B, a->k_ld, a->k_sb, RB) ||
        make_map(&v64, a->v, D, N, B, a->v_ld, a->v_sb, RB) || make_map(&do64, a->dO, D, N, B, a->do_ld, a->do_sb, RB) ||
        make_map(&q16, a->q, D, N, B, a->q_ld, a->q_sb, CBB) || make_map(&k16, a->k, D, N, B, a->k_ld, a->k_sb, CBB) ||
        make_map(&v16, a->v, D, N, B, a->v_ld, a->v_sb, CBB) || make_map(&do16, a->dO, D, N, B, a->do_ld, a->do_sb, CBB))
        return -1;
    TfBwdParams p;
    p.N = N; p.nblk = nblk; p.nchunk = nc; p.bpc = bpc; p.Npad = nrb * RB;
    p.Wl = a->Wl; p.bl = a->bl; p.Ww = a->Ww; p.bw = a->bw; p.scale = a->scale;
    p.lse2 = a->lse2; p.delta = delta;
    {
        static const char* dbg_env = getenv("SPE_TF_DBG");
        const int dbg = dbg_env ? atoi(dbg_env) : 0;
        if ((dbg & 1024) && !g_tf_trace) { cudaMalloc(&g_tf_trace, 4 * 32 * 8 * 8); cudaMemset(g_tf_trace, 0, 4 * 32 * 8 * 8); }
        p.trace = nullptr;
        g_bwd_trace_mode = (dbg & 1024) ? (dbg >> 12) : -1;
    }
    const double pos = (double)B * N * N * H;
    p.dwacc = dw; p.out0 = nullptr; p.out1 = nullptr;
    if (launch_bwd_one<H, BW_DELTA>(q64, do64, k16, v16, p, nrb, B, st, "tf_delta", pos * 4.0)) return -1;
    p.dwacc = dw + H * H; p.out0 = acc;
    if (launch_bwd_one<H, BW_DQ>(q64, do64, k16, v16, p, nrb, B, st, "tf_dq", pos * 4.0)) return -1;
    p.dwacc = nullptr; p.out0 = acc + (size_t)B * N * D; p.out1 = acc + (size_t)2 * B * N * D;
    if (launch_bwd_one<H, BW_DKV>(k64, v64, q16, do16, p, nrb, B, st, "tf_dkv", pos * 4.0)) return -1;
    {
        const long long tot = (long long)3 * B * N * D / 4;
        tf_pack_dqkv_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(acc, reinterpret_cast<uint16_t*>(a->dqkv), (long long)B * N, D, a->dqkv_ld);
        SPE_LAUNCHED();
        tf_addw_kernel<<<1, 2 * H * H, 0, st>>>(dw, a->dWw, a->dWl, H * H);
        SPE_LAUNCHED();
    }
    return 0;
}

}  // namespace

// timing experiments: copies the clock64 trace of the last main-kernel launch (SPE_TF_DBG & 512) to host memory [4][32][8]
extern "C" __attribute__((visibility("default"))) int spe_talking_fused_trace(long long* host_out) {
    if (!g_tf_trace) return -1;
    cudaDeviceSynchronize();
    return cudaMemcpy(host_out, g_tf_trace, 4 * 32 * 8 * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}
extern "C" __attribute__((visibility("default"))) int spe_talking_fused_supported(int H, int dh) { return (H == 4 || H == 8) && dh == DHD ? 1 : 0; }

extern "C" __attribute__((visibility("default"))) int64_t spe_talking_fused_fwd_workspace(int B, int H, int N, int dh) {
    const int nrb = (N + RB - 1) / RB;
    const int64_t part = (((int64_t)B * 8 * H * nrb * RB * 8 + 255) / 256) * 256;       // up to 8 chunks
    return part + (int64_t)B * N * H * dh * 4;
}

extern "C" __attribute__((visibility("default"))) int64_t spe_talking_fused_bwd_workspace(int B, int H, int N, int dh) {
    return (((int64_t)B * H * N * 4 + 255) / 256) * 256 + 1024 + (int64_t)3 * B * N * H * dh * 4;
}

extern "C" __attribute__((visibility("default"))) int spe_talking_fused_bwd(const spe_talking_fused_bwd_args* a, void* stream) {
    SPE_CHECK(a && a->q && a->k && a->v && a->dO && a->lse2 && a->dqkv && a->dWl && a->dWw && a->Wl && a->bl && a->Ww && a->bw, "spe_talking_fused_bwd: null argument");
    SPE_CHECK(spe_talking_fused_supported(a->H, a->dh), "spe_talking_fused_bwd: unsupported head geometry H=%d dh=%d", a->H, a->dh);
    SPE_CHECK(a->B > 0 && a->N > 0 && a->B <= 65535, "spe_talking_fused_bwd: bad shape");
    SPE_CHECK(a->dqkv_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(a->dqkv) & 7) == 0, "spe_talking_fused_bwd: dqkv not 8-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return a->H == 8 ? launch_bwd<8>(a, st) : launch_bwd<4>(a, st);
}

extern "C" __attribute__((visibility("default"))) int spe_talking_fused_fwd(const spe_talking_fused_args* a, void* stream) {
    SPE_CHECK(a && a->q && a->k && a->v && a->out && a->lse2 && a->Wl && a->bl && a->Ww && a->bw, "spe_talking_fused_fwd: null argument");
    SPE_CHECK(spe_talking_fused_supported(a->H, a->dh), "spe_talking_fused_fwd: unsupported head geometry H=%d dh=%d", a->H, a->dh);
    SPE_CHECK(a->B > 0 && a->N > 0 && a->B <= 65535, "spe_talking_fused_fwd: bad shape");
    SPE_CHECK(a->out_ld % 8 == 0 && a->out_sb % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "spe_talking_fused_fwd: output not 16-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return a->H == 8 ? launch_fwd<8>(a, st) : launch_fwd<4>(a, st);
}
