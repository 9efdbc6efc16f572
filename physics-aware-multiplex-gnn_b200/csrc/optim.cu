// Fused optimizer step on the flat parameter / gradient buffers: the reference's per-step
// sequence  clip_grad_norm_(max_norm) -> Adam.step() -> EMA(model)  (main_qm9.py:111-112,117; utils/ema.py:13-20)
// walks 390 tensors three times from Python.  Here it is two launches: a sum-of-squares reduction, then ONE
// streaming pass that reads (p, g, m, v, shadow) and writes (p, m, v, shadow) -- pure HBM/L2 traffic, 36 B per
// parameter, no host synchronisation (the clip coefficient is read from device memory).
#include "optim.cuh"

namespace pamnet {

constexpr int kOptThreads = 256;

__global__ void __launch_bounds__(kOptThreads) sumsq_kernel(const float* __restrict__ g, int64_t n4, double* __restrict__ out) {
    float s = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = ld4(g + 4 * i);
        s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    __shared__ double part[kOptThreads / 32];
    double d = (double)warp_sum(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kOptThreads / 32; ++w) t += part[w];
        atomicAdd(out, t);
    }
}

// torch.optim.Adam (amsgrad=False, L2 weight decay added to the gradient) with the formulas of its fp32 kernels:
//   g' = clip * g + wd * p;  m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// then utils/ema.py:16-20:  shadow = (1 - decay) p + decay shadow.
__global__ void __launch_bounds__(kOptThreads) adam_ema_kernel(const OptimArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n4) return;
    const int64_t e = 4 * i;
#pragma unroll
    for (int s = 0; s < kOptMaxSkip; ++s)
        if (s < a.n_skip && e >= a.skip_begin[s] && e < a.skip_end[s]) return;   // tensors without a gradient: untouched
    float clip = 1.f;
    if (a.max_norm > 0.f) {
        // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
        const float total = (float)sqrt(*a.sumsq);
        clip = fminf(a.max_norm / (total + 1e-6f), 1.f);
    }
    float4 p = ld4(a.p + e), g = ld4(a.g + e), m = ld4(a.m + e), v = ld4(a.v + e);
    float* pp = &p.x; float* gg = &g.x; float* mm = &m.x; float* vv = &v.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float gj = gg[j] * clip;
        if (a.weight_decay != 0.f) gj = fmaf(a.weight_decay, pp[j], gj);
        mm[j] = a.beta1 * mm[j] + (1.f - a.beta1) * gj;                 // exp_avg.lerp_(grad, 1 - beta1)
        vv[j] = a.beta2 * vv[j] + (1.f - a.beta2) * gj * gj;           // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
        const float denom = sqrtf(vv[j]) / a.bc2_sqrt + a.eps;
        pp[j] = pp[j] - a.step_size * (mm[j] / denom);
        gg[j] = gj;
    }
    st4(a.p + e, p); st4(a.m + e, m); st4(a.v + e, v);
    if (a.write_clipped_grad) st4(a.g + e, g);
    if (a.shadow) {
        float4 sh = ld4(a.shadow + e);
        sh.x = (1.f - a.ema_decay) * p.x + a.ema_decay * sh.x;
        sh.y = (1.f - a.ema_decay) * p.y + a.ema_decay * sh.y;
        sh.z = (1.f - a.ema_decay) * p.z + a.ema_decay * sh.z;
        sh.w = (1.f - a.ema_decay) * p.w + a.ema_decay * sh.w;
        st4(a.shadow + e, sh);
    }
}

int optimizer_step(const OptimArgs& a, cudaStream_t st) {
    PAMNET_CHECK_ARG(a.n4 >= 0 && a.n_skip >= 0 && a.n_skip <= kOptMaxSkip, "optimizer_step: bad sizes");
    if (a.n4 == 0) return 0;
    if (a.max_norm > 0.f) {
        PAMNET_CHECK_ARG(a.sumsq != nullptr, "optimizer_step: clipping needs the device scratch double");
        PAMNET_CUDA(cudaMemsetAsync(a.sumsq, 0, sizeof(double), st));
        const int grid = (int)(a.n4 / kOptThreads < 4 * kNumSM ? (a.n4 + kOptThreads - 1) / kOptThreads : 4 * kNumSM);
        prof_begin(KC_MISC, 16.0 * a.n4, st);
        sumsq_kernel<<<grid, kOptThreads, 0, st>>>(a.g, a.n4, a.sumsq);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    prof_begin(KC_MISC, 16.0 * a.n4 * (a.shadow ? 9.0 : 7.0), st);
    adam_ema_kernel<<<ceil_div(a.n4, kOptThreads), kOptThreads, 0, st>>>(a);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

}  // namespace pamnet
