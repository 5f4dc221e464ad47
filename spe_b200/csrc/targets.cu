// GT jitter + repeat of SetCriterion.forward in training mode (models/conditional_detr.py:410-431), on the device.
//
// Reference, per GT box: draw 1000 candidates box * U(1 - j, 1 + j)^4 (cx, cy, w, h scaled independently), keep those whose IoU with the
// original exceeds 0.7, and emit r = hung_match_ratio rows: the first min(r - 1, #kept) kept candidates IN CANDIDATE ORDER, then the
// original box; labels / scores are repeated r times.  It is a Python double loop with ~15 tiny kernels and 4000 RNG draws per box.
// Here: one warp per GT box, lane l tests candidates l, l + 32, ... with a counter-based generator (Philox4x32-10: the four outputs
// of counter (candidate, box) are the four scales), a ballot keeps the candidate order, and the warp stops as soon as r - 1 are
// found.  The random stream is not torch's (that one is not reproducible across devices either): parity is the acceptance rule, the
// ordering rule and the distribution (tests/test_targets_gpu.py), not the bits.
// rng: device u64[2] = {seed, launch counter}; the counter is advanced by the last block, so a captured CUDA graph draws new
// candidates on every replay.
#include "common.cuh"

namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float iou_cxcywh(const float (&a)[4], const float (&b)[4]) {
    // util/box_ops.py:18-22, 33-46: cxcywh -> xyxy, intersection clamped at 0, union = area_a + area_b - inter
    const float ax0 = a[0] - 0.5f * a[2], ay0 = a[1] - 0.5f * a[3], ax1 = a[0] + 0.5f * a[2], ay1 = a[1] + 0.5f * a[3];
    const float bx0 = b[0] - 0.5f * b[2], by0 = b[1] - 0.5f * b[3], bx1 = b[0] + 0.5f * b[2], by1 = b[1] + 0.5f * b[3];
    const float w = fmaxf(fminf(ax1, bx1) - fmaxf(ax0, bx0), 0.f), h = fmaxf(fminf(ay1, by1) - fmaxf(ay0, by0), 0.f);
    const float inter = w * h;
    const float uni = (ax1 - ax0) * (ay1 - ay0) + (bx1 - bx0) * (by1 - by0) - inter;
    return inter / uni;
}

__global__ void __launch_bounds__(256) gt_jitter_repeat_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ labels, const float* __restrict__ scores,
                                                                const int32_t* __restrict__ offsets, int B, int cap_total, int ratio, float jitter, int n_try,
                                                                float iou_thr, unsigned long long* __restrict__ rng, float* __restrict__ boxes_out,
                                                                int32_t* __restrict__ labels_out, float* __restrict__ scores_out, int32_t* __restrict__ offsets_out,
                                                                int32_t* __restrict__ counts_out, unsigned int* __restrict__ ticket) {
    const int lane = threadIdx.x & 31;
    const int j = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);            // GT box of this warp
    const unsigned long long seed = rng[0], launch = rng[1];
    const int total = offsets[B];
    if (blockIdx.x == 0) {
        for (int b = threadIdx.x; b <= B; b += blockDim.x) {
            offsets_out[b] = offsets[b] * ratio;
            if (b < B && counts_out) counts_out[b] = (offsets[b + 1] - offsets[b]) * ratio;
        }
    }
    if (j < total && j < cap_total) {
        float box[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) box[c] = boxes[4 * j + c];
        const int want = ratio - 1;
        int found = 0;
        for (int base = 0; base < n_try && found < want; base += 32) {
            const int cand = base + lane;
            uint32_t r[4];
            philox4x32_10((uint32_t)cand, (uint32_t)j, (uint32_t)launch, (uint32_t)(launch >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
            float sb[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float u = (float)(r[c] >> 8) * (1.0f / 16777216.0f);              // [0, 1)
                sb[c] = box[c] * ((1.f - jitter) + u * (2.f * jitter));                  // uniform_(1 - j, 1 + j)
            }
            const bool keep = cand < n_try && iou_cxcywh(sb, box) > iou_thr;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            const int rank = found + __popc(m & ((1u << lane) - 1u));                   // position among the kept candidates, candidate order
            if (keep && rank < want) {
#pragma unroll
                for (int c = 0; c < 4; ++c) boxes_out[4 * ((long long)j * ratio + rank) + c] = sb[c];
            }
            found += __popc(m);
        }
        if (found > want) found = want;
        // remaining rows: the original box (box_j.repeat(r, 1) with the first `found` rows overwritten)
        for (int t = found + lane; t < ratio; t += 32) {
#pragma unroll
            for (int c = 0; c < 4; ++c) boxes_out[4 * ((long long)j * ratio + t) + c] = box[c];
        }
        for (int t = lane; t < ratio; t += 32) {
            labels_out[(long long)j * ratio + t] = labels[j];
            if (scores_out) scores_out[(long long)j * ratio + t] = scores ? scores[j] : 1.f;
        }
    }
    // the last block to finish advances the launch counter (every block has read it by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            rng[1] = launch + 1ull;
            *ticket = 0u;
        }
    }
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int spe_gt_jitter_repeat(const float* boxes, const int32_t* labels, const float* scores, const int32_t* offsets, int B,
                                                                           int cap_total, int ratio, float jitter, int n_try, float iou_thr, uint64_t* rng_state,
                                                                           float* boxes_out, int32_t* labels_out, float* scores_out, int32_t* offsets_out,
                                                                           int32_t* counts_out, uint32_t* ticket, void* stream) {
    SPE_CHECK(boxes && labels && offsets && rng_state && boxes_out && labels_out && offsets_out && ticket, "spe_gt_jitter_repeat: null argument");
    SPE_CHECK(B > 0 && cap_total > 0 && ratio >= 1 && ratio <= 64 && n_try >= 0 && jitter >= 0.f && jitter < 1.f, "spe_gt_jitter_repeat: bad argument");
    const int warps_per_block = 8;
    const int grid = (cap_total + warps_per_block - 1) / warps_per_block;
    gt_jitter_repeat_kernel<<<grid, 32 * warps_per_block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        boxes, labels, scores, offsets, B, cap_total, ratio, jitter, n_try, iou_thr, reinterpret_cast<unsigned long long*>(rng_state), boxes_out, labels_out,
        scores_out, offsets_out, counts_out, ticket);
    SPE_LAUNCHED();
    return 0;
}
