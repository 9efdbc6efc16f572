// Fused clip + Adam + EMA step on flat buffers (optim.cu).
#pragma once
#include "common.cuh"

namespace pamnet {

constexpr int kOptMaxSkip = 4;

struct OptimArgs {
    float* p;            // parameters (flat, reference state_dict order, 128 B aligned tensors)
    float* g;            // gradients, same layout
    float* m;            // Adam exp_avg
    float* v;            // Adam exp_avg_sq
    float* shadow;       // EMA shadow parameters, nullptr = no EMA
    double* sumsq;       // device scratch: sum of squared gradients of this step (clip_grad_norm_'s total_norm^2)
    int64_t n4;          // number of float4 groups (buffer length / 4)
    int n_skip;          // element ranges [begin, end) of tensors that have no gradient (never touched)
    int write_clipped_grad;
    int64_t skip_begin[kOptMaxSkip], skip_end[kOptMaxSkip];
    float step_size;     // lr / (1 - beta1^t)
    float bc2_sqrt;      // sqrt(1 - beta2^t)
    float beta1, beta2, eps, weight_decay;
    float max_norm;      // <= 0: no clipping
    float ema_decay;
};

int optimizer_step(const OptimArgs& a, cudaStream_t st);

}  // namespace pamnet
