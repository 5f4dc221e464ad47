// Set-criterion kernels (models/conditional_detr.py:225-319, 468-494, 504-561; util/box_ops.py:33-74).
// Every loss kernel also produces the gradient wrt its prediction input (fused fwd+bwd): the Python
// autograd shell only multiplies it by the incoming scalar gradient.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// focal classification loss + class_error + cardinality_error  (single CTA: B*Q*C is ~2e5 elements)
// ------------------------------------------------------------------------------------------------
// stage 1: one warp per (b,q) row -> per-CTA partials {loss, correct, matched} + per-image "object" counts (atomics on ints)
__global__ void __launch_bounds__(256) focal_loss_kernel(const float* __restrict__ logits, const int32_t* __restrict__ row_to_gt,
                                                         const int32_t* __restrict__ gt_labels, const int32_t* __restrict__ gt_off,
                                                         const float* __restrict__ gt_scores, const float* __restrict__ inv_num_boxes,
                                                         int B, int Q, int C, float alpha, float gamma, float* __restrict__ part,
                                                         int* __restrict__ card, float* __restrict__ dlogits) {
    __shared__ float red[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const float inb = *inv_num_boxes;
    float loss_sum = 0.f, correct = 0.f, matched = 0.f;
    for (int row = blockIdx.x * nwarps + warp; row < B * Q; row += gridDim.x * nwarps) {
        const int b = row / Q;
        const int g = row_to_gt[row];
        const int g0 = gt_off[b];
        const int label = g >= 0 ? gt_labels[g0 + g] : C;          // C = "no object": all-zero one-hot row (:246-253)
        float w = 1.f;
        if (gt_scores) {
            if (g >= 0) w = fminf(gt_scores[g0 + g] * 3.f, 1.f);  // :527-529
            else {                                                // :523-524 avg score of the image
                const int G = gt_off[b + 1] - g0;
                float s = 0.f;
                for (int t = 0; t < G; ++t) s += gt_scores[g0 + t];
                w = s / (float)G;
            }
        }
        const float* x = logits + (long long)row * C;
        float best = -INFINITY; int besti = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            const float xv = x[c];
            const float t = (c == label) ? 1.f : 0.f;
            const float p = 1.f / (1.f + __expf(-xv));
            // binary_cross_entropy_with_logits
            const float ce = fmaxf(xv, 0.f) - xv * t + log1pf(__expf(-fabsf(xv)));
            float pt = p * t + (1.f - p) * (1.f - t);
            const bool clamped = (pt < 1e-5f) || (pt > 1.f - 1e-5f);
            pt = fminf(fmaxf(pt, 1e-5f), 1.f - 1e-5f);
            const float om = 1.f - pt;
            const float mod = powf(om, gamma);
            const float at = alpha >= 0.f ? (alpha * t + (1.f - alpha) * (1.f - t)) : 1.f;
            loss_sum += w * at * ce * mod;
            if (dlogits) {
                const float dce = p - t;                                   // d ce / dx
                const float dpt = (t > 0.5f ? 1.f : -1.f) * p * (1.f - p);  // d p_t / dx
                const float dmod = clamped ? 0.f : -gamma * (mod / om) * dpt;
                dlogits[(long long)row * C + c] = w * at * (dce * mod + ce * dmod) * inb;
            }
            if (xv > best || (xv == best && c < besti)) { best = xv; besti = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0) {
            if (besti != C - 1) atomicAdd(&card[b], 1);            // :293-295 (integer atomics: deterministic)
            if (g >= 0) { matched += 1.f; if (besti == label) correct += 1.f; }
        }
    }
    loss_sum = block_sum(loss_sum, red);
    correct = block_sum(correct, red);
    matched = block_sum(matched, red);
    if (tid == 0) { part[blockIdx.x * 3 + 0] = loss_sum; part[blockIdx.x * 3 + 1] = correct; part[blockIdx.x * 3 + 2] = matched; }
}

// stage 2: deterministic reduction of the per-CTA partials
__global__ void __launch_bounds__(256) focal_loss_finalize_kernel(const float* __restrict__ part, int nblocks, const int* __restrict__ card,
                                                                  const int32_t* __restrict__ gt_off, const float* __restrict__ inv_num_boxes, int B,
                                                                  float* __restrict__ out) {
    __shared__ float red[32];
    float l = 0.f, c = 0.f, m = 0.f, e = 0.f;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { l += part[i * 3]; c += part[i * 3 + 1]; m += part[i * 3 + 2]; }
    for (int b = threadIdx.x; b < B; b += blockDim.x) e += fabsf((float)card[b] - (float)(gt_off[b + 1] - gt_off[b]));
    l = block_sum(l, red); c = block_sum(c, red); m = block_sum(m, red); e = block_sum(e, red);
    if (threadIdx.x == 0) {
        out[0] = l * (*inv_num_boxes);                                             // mean(1).sum()/num_boxes*Q
        out[1] = m > 0.f ? 100.f - c * (100.f / m) : 100.f;                        // util/misc.py:440-455
        out[2] = e / (float)B;                                                     // F.l1_loss mean
    }
}

// ------------------------------------------------------------------------------------------------
// L1 + GIoU on matched pairs, with gradients (forward-mode duals over x0,y0,x1,y1 of the prediction)
// ------------------------------------------------------------------------------------------------
struct D4 { float v, d[4]; };   // value + partials wrt (x0,y0,x1,y1)
__device__ __forceinline__ D4 d_const(float v) { D4 r; r.v = v; r.d[0] = r.d[1] = r.d[2] = r.d[3] = 0.f; return r; }
__device__ __forceinline__ D4 d_var(float v, int i) { D4 r = d_const(v); r.d[i] = 1.f; return r; }
__device__ __forceinline__ D4 d_sub(const D4& a, const D4& b) { D4 r; r.v = a.v - b.v; for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ D4 d_add(const D4& a, const D4& b) { D4 r; r.v = a.v + b.v; for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ D4 d_mul(const D4& a, const D4& b) { D4 r; r.v = a.v * b.v; for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ __forceinline__ D4 d_div(const D4& a, const D4& b) { D4 r; r.v = a.v / b.v; for (int i = 0; i < 4; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
__device__ __forceinline__ D4 d_min(const D4& a, const D4& b) { return a.v <= b.v ? a : b; }
__device__ __forceinline__ D4 d_max(const D4& a, const D4& b) { return a.v >= b.v ? a : b; }
__device__ __forceinline__ D4 d_relu(const D4& a) { return a.v >= 0.f ? a : d_const(0.f); }

__global__ void __launch_bounds__(256) box_loss_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ row_to_gt,
                                                       const float* __restrict__ gt_boxes, const int32_t* __restrict__ gt_off,
                                                       const float* __restrict__ gt_scores, const float* __restrict__ inv_num_boxes, int B, int Q,
                                                       float* __restrict__ out, float* __restrict__ d_l1, float* __restrict__ d_giou) {
    __shared__ float red[32];
    const float inb = *inv_num_boxes;
    float s_l1 = 0.f, s_g = 0.f;
    for (int row = threadIdx.x; row < B * Q; row += blockDim.x) {
        const int g = row_to_gt[row];
        float4 gl = make_float4(0.f, 0.f, 0.f, 0.f), gg = gl;
        if (g >= 0) {
            const int b = row / Q;
            const float4 s = *reinterpret_cast<const float4*>(boxes + (long long)row * 4);
            const float4 t = *reinterpret_cast<const float4*>(gt_boxes + (long long)(gt_off[b] + g) * 4);
            const float w = gt_scores ? gt_scores[gt_off[b] + g] : 1.f;
            s_l1 += w * (fabsf(s.x - t.x) + fabsf(s.y - t.y) + fabsf(s.z - t.z) + fabsf(s.w - t.w));
            auto sgn = [](float a) { return a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f); };
            gl = make_float4(w * sgn(s.x - t.x) * inb, w * sgn(s.y - t.y) * inb, w * sgn(s.z - t.z) * inb, w * sgn(s.w - t.w) * inb);
            // GIoU with duals
            const D4 x0 = d_var(s.x - 0.5f * s.z, 0), y0 = d_var(s.y - 0.5f * s.w, 1), x1 = d_var(s.x + 0.5f * s.z, 2), y1 = d_var(s.y + 0.5f * s.w, 3);
            const D4 X0 = d_const(t.x - 0.5f * t.z), Y0 = d_const(t.y - 0.5f * t.w), X1 = d_const(t.x + 0.5f * t.z), Y1 = d_const(t.y + 0.5f * t.w);
            const D4 area_a = d_mul(d_sub(x1, x0), d_sub(y1, y0));
            const D4 area_b = d_mul(d_sub(X1, X0), d_sub(Y1, Y0));
            const D4 iw = d_relu(d_sub(d_min(x1, X1), d_max(x0, X0))), ih = d_relu(d_sub(d_min(y1, Y1), d_max(y0, Y0)));
            const D4 inter = d_mul(iw, ih);
            const D4 uni = d_sub(d_add(area_a, area_b), inter);
            const D4 iou = d_div(inter, uni);
            const D4 ew = d_relu(d_sub(d_max(x1, X1), d_min(x0, X0))), eh = d_relu(d_sub(d_max(y1, Y1), d_min(y0, Y0)));
            const D4 earea = d_mul(ew, eh);
            const D4 giou = d_sub(iou, d_div(d_sub(earea, uni), earea));
            s_g += w * (1.f - giou.v);
            // chain to cxcywh: x0 = cx - w/2, x1 = cx + w/2
            const float k = -w * inb;
            gg = make_float4(k * (giou.d[0] + giou.d[2]), k * (giou.d[1] + giou.d[3]), k * 0.5f * (giou.d[2] - giou.d[0]),
                             k * 0.5f * (giou.d[3] - giou.d[1]));
        }
        if (d_l1) *reinterpret_cast<float4*>(d_l1 + (long long)row * 4) = gl;
        if (d_giou) *reinterpret_cast<float4*>(d_giou + (long long)row * 4) = gg;
    }
    s_l1 = block_sum(s_l1, red);
    s_g = block_sum(s_g, red);
    if (threadIdx.x == 0) { out[0] = s_l1 * inb; out[1] = s_g * inb; }
}

__global__ void __launch_bounds__(256) bce_logits_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n,
                                                         float* __restrict__ out, float* __restrict__ dx) {
    __shared__ float red[32];
    float s = 0.f;
    const float inv = 1.f / (float)n;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float xv = x[i], t = y[i];
        s += fmaxf(xv, 0.f) - xv * t + log1pf(__expf(-fabsf(xv)));
        if (dx) dx[i] = (1.f / (1.f + __expf(-xv)) - t) * inv;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s * inv;
}

__global__ void box_iou_pairwise_kernel(const float* __restrict__ a, int N, const float* __restrict__ b, int M, float* __restrict__ iou,
                                        float* __restrict__ uni_o, float* __restrict__ giou_o) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)N * M) return;
    const int i = (int)(idx / M), j = (int)(idx % M);
    const float4 p = *reinterpret_cast<const float4*>(a + (long long)i * 4);
    const float4 t = *reinterpret_cast<const float4*>(b + (long long)j * 4);
    const float area_a = (p.z - p.x) * (p.w - p.y), area_b = (t.z - t.x) * (t.w - t.y);
    const float iw = fmaxf(fminf(p.z, t.z) - fmaxf(p.x, t.x), 0.f), ih = fmaxf(fminf(p.w, t.w) - fmaxf(p.y, t.y), 0.f);
    const float inter = iw * ih;
    const float uni = area_a + area_b - inter;
    const float v = inter / uni;
    if (iou) iou[idx] = v;
    if (uni_o) uni_o[idx] = uni;
    if (giou_o) {
        const float ew = fmaxf(fmaxf(p.z, t.z) - fminf(p.x, t.x), 0.f), eh = fmaxf(fmaxf(p.w, t.w) - fminf(p.y, t.y), 0.f);
        const float ea = ew * eh;
        giou_o[idx] = v - (ea - uni) / ea;
    }
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int64_t spe_focal_loss_workspace_bytes(int B, int Q) {
    const int blocks = (B * Q + 7) / 8 < 4 * spe_num_sms() ? (B * Q + 7) / 8 : 4 * spe_num_sms();
    return (int64_t)blocks * 3 * 4 + (int64_t)B * 4;
}

extern "C" __attribute__((visibility("default"))) int spe_focal_loss(const float* logits, const int32_t* row_to_gt, const int32_t* gt_labels, const int32_t* gt_off,
                              const float* gt_scores, const float* inv_num_boxes, int B, int Q, int C, float alpha, float gamma, float* out,
                              float* dlogits, void* workspace, void* stream) {
    SPE_CHECK(logits && row_to_gt && gt_labels && gt_off && inv_num_boxes && out && workspace, "spe_focal_loss: null argument");
    SPE_CHECK(B > 0 && Q > 0 && C > 0, "spe_focal_loss: bad shape");
    const int blocks = (B * Q + 7) / 8 < 4 * spe_num_sms() ? (B * Q + 7) / 8 : 4 * spe_num_sms();
    float* part = reinterpret_cast<float*>(workspace);
    int* card = reinterpret_cast<int*>(part + (size_t)blocks * 3);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SPE_CUDA(cudaMemsetAsync(card, 0, (size_t)B * 4, st));
    focal_loss_kernel<<<blocks, 256, 0, st>>>(logits, row_to_gt, gt_labels, gt_off, gt_scores, inv_num_boxes, B, Q, C, alpha, gamma, part, card, dlogits);
    SPE_LAUNCHED();
    focal_loss_finalize_kernel<<<1, 256, 0, st>>>(part, blocks, card, gt_off, inv_num_boxes, B, out);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_box_loss(const float* boxes, const int32_t* row_to_gt, const float* gt_boxes, const int32_t* gt_off, const float* gt_scores,
                            const float* inv_num_boxes, int B, int Q, float* out, float* dboxes_l1, float* dboxes_giou, void* stream) {
    SPE_CHECK(boxes && row_to_gt && gt_boxes && gt_off && inv_num_boxes && out, "spe_box_loss: null argument");
    box_loss_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(boxes, row_to_gt, gt_boxes, gt_off, gt_scores, inv_num_boxes, B, Q, out,
                                                                           dboxes_l1, dboxes_giou);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_bce_logits(const float* x, const float* y, int64_t n, float* out, float* dx, void* stream) {
    SPE_CHECK(x && y && out && n > 0, "spe_bce_logits: bad argument");
    bce_logits_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, n, out, dx);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_box_iou_pairwise(const float* a, int N, const float* b, int M, float* iou, float* uni, float* giou, void* stream) {
    SPE_CHECK(a && b && N > 0 && M > 0, "spe_box_iou_pairwise: bad argument");
    const long long tot = (long long)N * M;
    box_iou_pairwise_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, N, b, M, iou, uni, giou);
    SPE_LAUNCHED();
    return 0;
}
