// Hungarian matcher on the GPU (models/matcher.py:41-87).  Compiled with -fmad=false so the fp32 cost
// follows the reference's unfused op order (SURVEY.md H2 iii).
//
//   spe_match_cost   : warp-per-query cost kernel, block-diagonal only
//   spe_lsap_batched : one CTA per image; scipy's rectangular shortest-augmenting-path LSAP
//                      (Crouse 2016) with fp64 duals and scipy's exact scan order /
//                      tie-breaking, the inner column scan parallelised with a lexicographic arg-min.
#include "common.cuh"
#include <math.h>

namespace {

// ------------------------------------------------------------------------------------------------
// cost matrix
// ------------------------------------------------------------------------------------------------
__global__ void match_cost_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                  const int32_t* __restrict__ gt_labels, const float* __restrict__ gt_boxes,
                                  const int32_t* __restrict__ gt_off, int Q, int C, float w_class, float w_bbox, float w_giou,
                                  float* __restrict__ cost, long long ldc, int gt_images) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= Q) return;
    const int bt = gt_images > 0 ? b % gt_images : b;      // decoder levels stacked along the batch share the targets
    const int g0 = gt_off[bt], G = gt_off[bt + 1] - g0;
    const float4 bx = *reinterpret_cast<const float4*>(boxes + ((long long)b * Q + q) * 4);
    // box_cxcywh_to_xyxy (util/box_ops.py:18-22)
    const float ax0 = bx.x - 0.5f * bx.z, ay0 = bx.y - 0.5f * bx.w, ax1 = bx.x + 0.5f * bx.z, ay1 = bx.y + 0.5f * bx.w;
    const float area_a = (ax1 - ax0) * (ay1 - ay0);
    const float* lrow = logits + ((long long)b * Q + q) * C;
    float* crow = cost + ((long long)b * Q + q) * ldc;
    for (int g = lane; g < G; g += 32) {
        const int lab = gt_labels[g0 + g];
        // matcher.py:62,70-74 (alpha .25, gamma 2 hard-coded)
        const float x = lrow[lab];
        const float p = 1.f / (1.f + expf(-x));
        const float neg = (0.75f * (p * p)) * (-logf((1.f - p) + 1e-8f));
        const float om = 1.f - p;
        const float pos = (0.25f * (om * om)) * (-logf(p + 1e-8f));
        const float c_class = pos - neg;
        const float4 t = *reinterpret_cast<const float4*>(gt_boxes + (long long)(g0 + g) * 4);
        // cdist p=1 (matcher.py:77)
        const float c_bbox = ((fabsf(bx.x - t.x) + fabsf(bx.y - t.y)) + fabsf(bx.z - t.z)) + fabsf(bx.w - t.w);
        // generalized_box_iou (util/box_ops.py:49-74)
        const float bx0 = t.x - 0.5f * t.z, by0 = t.y - 0.5f * t.w, bx1 = t.x + 0.5f * t.z, by1 = t.y + 0.5f * t.w;
        const float area_b = (bx1 - bx0) * (by1 - by0);
        const float iw = fmaxf(fminf(ax1, bx1) - fmaxf(ax0, bx0), 0.f), ih = fmaxf(fminf(ay1, by1) - fmaxf(ay0, by0), 0.f);
        const float inter = iw * ih;
        const float uni = (area_a + area_b) - inter;
        const float iou = inter / uni;
        const float ew = fmaxf(fmaxf(ax1, bx1) - fminf(ax0, bx0), 0.f), eh = fmaxf(fmaxf(ay1, by1) - fminf(ay0, by0), 0.f);
        const float earea = ew * eh;
        const float giou = iou - (earea - uni) / earea;
        // matcher.py:82
        crow[g] = (w_bbox * c_bbox + w_class * c_class) + w_giou * (-giou);
    }
}

// ------------------------------------------------------------------------------------------------
// LSAP
// ------------------------------------------------------------------------------------------------
struct Cand {
    double val;
    int it;        // scan position
    int unassigned;
};

// scipy's sequential rule: strictly smaller wins; on equality an unassigned column overrides.
// => among minima: the LAST unassigned in scan order if any, else the FIRST.
__device__ __forceinline__ bool cand_better(const Cand& a, const Cand& b) {
    if (a.it < 0) return false;
    if (b.it < 0) return true;
    if (a.val < b.val) return true;
    if (a.val > b.val) return false;
    if (a.unassigned != b.unassigned) return a.unassigned != 0;
    return a.unassigned ? (a.it > b.it) : (a.it < b.it);
}

__device__ __forceinline__ Cand cand_shfl_xor(const Cand& c, int o) {
    Cand r;
    r.val = __shfl_xor_sync(0xffffffffu, c.val, o);
    r.it = __shfl_xor_sync(0xffffffffu, c.it, o);
    r.unassigned = __shfl_xor_sync(0xffffffffu, c.unassigned, o);
    return r;
}

template <int T>
__device__ __forceinline__ void bar() {
    if (T == 32) __syncwarp(); else __syncthreads();
}

struct LsapShared {
    double min_val;
    int i, sink, num_remaining, index;
};

template <int T>
__global__ void __launch_bounds__(T) lsap_kernel(const float* __restrict__ cost_all, int nr, long long ldc, const int32_t* __restrict__ nc_arr,
                                                 int max_nc, int32_t* __restrict__ row_to_col, int max_small, int max_big, int stage_cost) {
    extern __shared__ double smem_d[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nc = nc_arr ? nc_arr[b] : max_nc;
    const float* cost = cost_all + (long long)b * nr * ldc;
    int32_t* out = row_to_col + (long long)b * nr;
    if (nc <= 0) {
        for (int q = tid; q < nr; q += T) out[q] = -1;
        return;
    }
    const bool transposed = nc < nr;              // scipy works on the transpose when nc < nr
    const int R = transposed ? nc : nr;           // internal rows (<= cols)
    const int Cn = transposed ? nr : nc;          // internal cols
    // shared carve (max_small >= R, max_big >= Cn)
    double* u = smem_d;                           // [max_small]
    double* v = u + max_small;                    // [max_big]
    double* spc = v + max_big;                    // [max_big]
    int* path = reinterpret_cast<int*>(spc + max_big);   // [max_big]
    int* row4col = path + max_big;                // [max_big]
    int* remaining = row4col + max_big;           // [max_big]
    int* col4row = remaining + max_big;           // [max_small]
    unsigned char* SR = reinterpret_cast<unsigned char*>(col4row + max_small);   // [max_small]
    unsigned char* SC = SR + max_small;           // [max_big]
    // optional: the problem's cost matrix staged in shared memory in INTERNAL orientation [R][Cn] -- every scan step of the transposed
    // case (the usual one: fewer GT than queries) reads one cost per lane at stride ldc from global memory, and the algorithm is a
    // chain of such scans: with the matrix on chip the kernel is no longer bound by that load latency.  Same values, same order.
    float* costS = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(SC + max_big) + 15) & ~uintptr_t(15));
    if (stage_cost) {
        if (transposed) {
            for (int e = tid; e < R * Cn; e += T) { const int j = e / R, i = e - j * R; costS[i * Cn + j] = __ldg(cost + (long long)j * ldc + i); }
        } else {
            for (int e = tid; e < R * Cn; e += T) { const int i = e / Cn, j = e - i * Cn; costS[e] = __ldg(cost + (long long)i * ldc + j); }
        }
    }
    __shared__ LsapShared sh;
    __shared__ Cand wbest[T / 32 > 0 ? T / 32 : 1];

    for (int i = tid; i < R; i += T) { u[i] = 0.0; col4row[i] = -1; }
    for (int j = tid; j < Cn; j += T) { v[j] = 0.0; row4col[j] = -1; }
    bar<T>();

    for (int cur = 0; cur < R; ++cur) {
        for (int j = tid; j < Cn; j += T) { spc[j] = INFINITY; SC[j] = 0; remaining[j] = Cn - 1 - j; }
        for (int i = tid; i < R; i += T) SR[i] = 0;
        if (tid == 0) { sh.min_val = 0.0; sh.i = cur; sh.sink = -1; sh.num_remaining = Cn; }
        bar<T>();
        while (true) {
            const int i = sh.i;
            const int num_remaining = sh.num_remaining;
            const double min_val = sh.min_val;
            const double ui = u[i];
            Cand best; best.val = INFINITY; best.it = -1; best.unassigned = 0;
            for (int it = tid; it < num_remaining; it += T) {
                const int j = remaining[it];
                const float cf = stage_cost ? costS[i * Cn + j] : (transposed ? __ldg(cost + (long long)j * ldc + i) : __ldg(cost + (long long)i * ldc + j));
                const double r = ((min_val + (double)cf) - ui) - v[j];
                double s = spc[j];
                if (r < s) { path[j] = i; spc[j] = r; s = r; }
                Cand c; c.val = s; c.it = it; c.unassigned = (row4col[j] == -1);
                if (cand_better(c, best)) best = c;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const Cand other = cand_shfl_xor(best, o);
                if (cand_better(other, best)) best = other;
            }
            if (T > 32) {
                if ((tid & 31) == 0) wbest[tid >> 5] = best;
                __syncthreads();
                if (tid < 32) {
                    Cand c2; c2.val = INFINITY; c2.it = -1; c2.unassigned = 0;
                    if (tid < T / 32) c2 = wbest[tid];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const Cand other = cand_shfl_xor(c2, o);
                        if (cand_better(other, c2)) c2 = other;
                    }
                    best = c2;
                }
            }
            if (tid == 0 && best.it < 0) sh.sink = -2;          // infeasible (non-finite costs): give up
            if (tid == 0 && best.it >= 0) {
                SR[i] = 1;
                sh.min_val = best.val;
                const int j = remaining[best.it];
                if (row4col[j] == -1) sh.sink = j; else sh.i = row4col[j];
                SC[j] = 1;
                remaining[best.it] = remaining[num_remaining - 1];
                sh.num_remaining = num_remaining - 1;
            }
            bar<T>();
            if (sh.sink != -1) break;
        }
        if (sh.sink == -2) {
            for (int q = tid; q < nr; q += T) out[q] = -1;
            return;
        }
        const double min_val = sh.min_val;
        // dual updates (App. C): u[cur] += minVal; u[i] += minVal - spc[col4row[i]] for i in SR\{cur}; v[j] -= minVal - spc[j] for j in SC
        for (int i = tid; i < R; i += T) {
            if (i == cur) u[i] += min_val;
            else if (SR[i]) u[i] += min_val - spc[col4row[i]];
        }
        for (int j = tid; j < Cn; j += T)
            if (SC[j]) v[j] -= min_val - spc[j];
        bar<T>();
        if (tid == 0) {
            int j = sh.sink;
            while (true) {
                const int i = path[j];
                row4col[j] = i;
                const int t = col4row[i];
                col4row[i] = j;
                j = t;
                if (i == cur) break;
            }
        }
        bar<T>();
    }
    if (!transposed) {
        for (int q = tid; q < nr; q += T) out[q] = col4row[q];
    } else {
        for (int q = tid; q < nr; q += T) out[q] = row4col[q];   // internal col = original row; row4col = original column (or -1)
    }
}

size_t lsap_smem_bytes(int small, int big) {
    return (size_t)(small + 2 * (size_t)big) * 8 + (size_t)(3 * (size_t)big + small) * 4 + (size_t)small + (size_t)big + 16;
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int spe_match_cost(const float* logits, const float* boxes, const int32_t* gt_labels, const float* gt_boxes,
                              const int32_t* gt_off, int B, int Q, int C, float w_class, float w_bbox, float w_giou, float* cost,
                              int64_t ldc, int gt_images, void* stream) {
    SPE_CHECK(logits && boxes && gt_labels && gt_boxes && gt_off && cost, "spe_match_cost: null argument");
    SPE_CHECK(B > 0 && Q > 0 && C > 0 && gt_images >= 0, "spe_match_cost: bad shape");
    const int warps = 4;
    dim3 grid((Q + warps - 1) / warps, B);
    match_cost_kernel<<<grid, warps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, boxes, gt_labels, gt_boxes, gt_off, Q, C, w_class,
                                                                                         w_bbox, w_giou, cost, ldc, gt_images);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int64_t spe_lsap_workspace_bytes(int B, int nr, int max_nc) {
    (void)B; (void)nr; (void)max_nc;
    return 16;   // all state lives in shared memory; kept for ABI stability
}

extern "C" __attribute__((visibility("default"))) int spe_lsap_batched(const float* cost, int B, int nr, int64_t ldc, const int32_t* nc, int max_nc, int32_t* row_to_col,
                                void* workspace, void* stream) {
    (void)workspace;
    SPE_CHECK(cost && row_to_col, "spe_lsap_batched: null argument");
    SPE_CHECK(B > 0 && nr > 0 && max_nc >= 0 && ldc >= max_nc, "spe_lsap_batched: bad shape");
    if (max_nc == 0) {
        SPE_CUDA(cudaMemsetAsync(row_to_col, 0xff, (size_t)B * nr * 4, reinterpret_cast<cudaStream_t>(stream)));
        return 0;
    }
    const int small = nr < max_nc ? nr : max_nc, big = nr < max_nc ? max_nc : nr;
    // per problem R = min(nr, nc[b]) <= small and Cn = max(nr, nc[b]) <= big
    const int mb = big;
    const int use_small = small;
    size_t smem = lsap_smem_bytes(use_small, mb);
    SPE_CHECK(smem <= 200 * 1024, "spe_lsap_batched: problem too large for shared memory (%zu B)", smem);
    // stage the cost matrix on chip when it fits next to the solver state (300 x 50 training problems: 60 KB; 300 x 1000 does not)
    const size_t cost_bytes = (size_t)use_small * mb * 4 + 16;
    const int stage_cost = smem + cost_bytes <= 160 * 1024 ? 1 : 0;
    if (stage_cost) smem += cost_bytes;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SpeProfScope prof(SPE_FAM_MATCHER, (double)B, st);     // work unit = images
    if (big <= 512) {
        static bool done = false;
        if (!done) { SPE_CUDA(cudaFuncSetAttribute(lsap_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); done = true; }
        lsap_kernel<32><<<B, 32, smem, st>>>(cost, nr, ldc, nc, max_nc, row_to_col, use_small, mb, stage_cost);
    } else {
        static bool done = false;
        if (!done) { SPE_CUDA(cudaFuncSetAttribute(lsap_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); done = true; }
        lsap_kernel<256><<<B, 256, smem, st>>>(cost, nr, ldc, nc, max_nc, row_to_col, use_small, mb, stage_cost);
    }
    SPE_LAUNCHED();
    return 0;
}
