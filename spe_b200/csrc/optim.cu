// Optimizer step of the reference training loop on the flat buffers of spe_b200/dp.py:
//   torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)   (engine.py:163-164)
//   torch.optim.AdamW(3 parameter groups: lr / lr_backbone / lr_cls_head, weight_decay)   (main.py:177-190)
// The reference launches ~4 small kernels per parameter tensor (x ~600 tensors) for the norm and again for the update; here the
// gradients already live in ONE fp32 buffer, the parameters and both moments are laid out the same way, so a step is
//   spe_sumsq_f32 (two launches, deterministic)  ->  spe_adamw_tick (1 thread: beta^t, clip coefficient)  ->  spe_adamw_flat (one
//   HBM-bound pass: reads p, g, m, v, writes p, m, v [+ clipped g, + bf16 weight shadow]),
// all on the caller's stream, no host synchronisation (CUDA-graph capturable: step count, learning rates and the clip coefficient
// are device scalars).
#include "common.cuh"

namespace {

constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS) sumsq_partial_kernel(const float* __restrict__ x, long long n, float* __restrict__ partial) {
    __shared__ float red[32];
    float acc = 0.f;
    const long long n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x * OPT_THREADS) {
        const float4 v = x4[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0)
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) acc += x[i] * x[i];
    const float s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// fixed-order final reduction (double): the same bits every run
__global__ void __launch_bounds__(OPT_THREADS) sumsq_final_kernel(const float* __restrict__ partial, int np, float* __restrict__ out) {
    __shared__ double red[OPT_THREADS];
    double acc = 0.0;
    for (int i = threadIdx.x; i < np; i += OPT_THREADS) acc += (double)partial[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = OPT_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)red[0];
}

// state f32[4]: {beta1^t, beta2^t, clip coefficient, step count}
__global__ void adamw_tick_kernel(float* __restrict__ state, float beta1, float beta2, const float* __restrict__ gsumsq, float max_norm) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        state[0] *= beta1;
        state[1] *= beta2;
        state[3] += 1.f;
        float coef = 1.f;
        if (gsumsq != nullptr && max_norm > 0.f) {
            coef = max_norm / (sqrtf(gsumsq[0]) + 1e-6f);            // clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
            coef = coef > 1.f ? 1.f : coef;
        }
        state[2] = coef;
    }
}

struct AdamwSegs {
    int nseg;
    long long end[SPE_ADAMW_MAX_SEGMENTS];     // exclusive end offset of segment s (multiples of 4 elements)
    int group[SPE_ADAMW_MAX_SEGMENTS];
};

__global__ void __launch_bounds__(OPT_THREADS) adamw_flat_kernel(float* __restrict__ p, const float* g, float* __restrict__ m, float* __restrict__ v,
                                                                  long long n4, AdamwSegs segs, const float* __restrict__ lr, const float* __restrict__ wd,
                                                                  float beta1, float beta2, float eps, const float* __restrict__ state,
                                                                  float* g_out /* may alias g */, uint16_t* __restrict__ shadow) {
    const float b1t = state[0], b2t = state[1], coef = state[2];
    const float bc1 = 1.f - b1t, sbc2 = sqrtf(1.f - b2t);         // torch: step_size = lr / bias_correction1, denom = sqrt(v) / sqrt(bias_correction2) + eps
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x * OPT_THREADS) {
        const long long e = i << 2;
        int grp = segs.group[segs.nseg - 1];
#pragma unroll 1
        for (int s = 0; s < segs.nseg; ++s)
            if (e < segs.end[s]) { grp = segs.group[s]; break; }
        const float lr_g = lr[grp], wd_g = wd[grp];
        float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
        float* pa = reinterpret_cast<float*>(&pp);
        float* ga = reinterpret_cast<float*>(&gg);
        float* ma = reinterpret_cast<float*>(&mm);
        float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = ga[k] * coef;
            ga[k] = gk;
            float pk = pa[k] * (1.f - lr_g * wd_g);                  // decoupled weight decay first (torch.optim.AdamW)
            const float mk = ma[k] + (1.f - beta1) * (gk - ma[k]);   // lerp form, as torch's exp_avg.lerp_(grad, 1 - beta1)
            const float vk = beta2 * va[k] + (1.f - beta2) * gk * gk;
            const float denom = sqrtf(vk) / sbc2 + eps;
            pk -= (lr_g / bc1) * (mk / denom);
            pa[k] = pk; ma[k] = mk; va[k] = vk;
        }
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
        if (g_out) reinterpret_cast<float4*>(g_out)[i] = gg;
        if (shadow) {
            uint2 s2;
            s2.x = pack_bf16x2(pa[0], pa[1]);
            s2.y = pack_bf16x2(pa[2], pa[3]);
            reinterpret_cast<uint2*>(shadow)[i] = s2;
        }
    }
}

__global__ void __launch_bounds__(OPT_THREADS) scale_by_coef_kernel(float* __restrict__ x, long long n, const float* __restrict__ state) {
    const float coef = state[2];
    if (coef == 1.f) return;
    for (long long i = (long long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * OPT_THREADS) x[i] *= coef;
}

int opt_grid(long long work_items) {
    long long g = (work_items + OPT_THREADS - 1) / OPT_THREADS;
    const long long cap = (long long)spe_num_sms() * 8;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int64_t spe_sumsq_workspace_floats(void) { return (int64_t)spe_num_sms() * 8; }

extern "C" __attribute__((visibility("default"))) int spe_sumsq_f32(const float* x, int64_t n, float* out, float* workspace, void* stream) {
    SPE_CHECK(x && out && workspace && n > 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "spe_sumsq_f32: bad argument (x must be 16-byte aligned)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = opt_grid(n >> 2);
    sumsq_partial_kernel<<<grid, OPT_THREADS, 0, st>>>(x, n, workspace);
    SPE_LAUNCHED();
    sumsq_final_kernel<<<1, OPT_THREADS, 0, st>>>(workspace, grid, out);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_adamw_tick(float* state, float beta1, float beta2, const float* grad_sumsq, float max_norm, void* stream) {
    SPE_CHECK(state, "spe_adamw_tick: null state");
    adamw_tick_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(state, beta1, beta2, grad_sumsq, max_norm);
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_adamw_flat(const spe_adamw_args* a, void* stream) {
    SPE_CHECK(a && a->p && a->g && a->m && a->v && a->lr && a->wd && a->state && a->n > 0 && a->n % 4 == 0, "spe_adamw_flat: bad argument (n must be a multiple of 4)");
    SPE_CHECK(a->nseg >= 1 && a->nseg <= SPE_ADAMW_MAX_SEGMENTS, "spe_adamw_flat: 1..%d segments", SPE_ADAMW_MAX_SEGMENTS);
    AdamwSegs segs;
    segs.nseg = a->nseg;
    long long prev = 0;
    for (int s = 0; s < a->nseg; ++s) {
        SPE_CHECK(a->seg_end[s] > prev && a->seg_end[s] % 4 == 0 && a->seg_group[s] >= 0, "spe_adamw_flat: segment ends must increase in multiples of 4");
        segs.end[s] = a->seg_end[s];
        segs.group[s] = a->seg_group[s];
        prev = a->seg_end[s];
    }
    SPE_CHECK(prev == a->n, "spe_adamw_flat: the segments must cover the buffer");
    adamw_flat_kernel<<<opt_grid(a->n >> 2), OPT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        a->p, a->g, a->m, a->v, a->n >> 2, segs, a->lr, a->wd, a->beta1, a->beta2, a->eps, a->state, a->g_out, reinterpret_cast<uint16_t*>(a->shadow_bf16));
    SPE_LAUNCHED();
    return 0;
}

extern "C" __attribute__((visibility("default"))) int spe_scale_by_clip_coef(float* x, int64_t n, const float* state, void* stream) {
    SPE_CHECK(x && state && n > 0, "spe_scale_by_clip_coef: bad argument");
    scale_by_coef_kernel<<<opt_grid(n), OPT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, n, state);
    SPE_LAUNCHED();
    return 0;
}
