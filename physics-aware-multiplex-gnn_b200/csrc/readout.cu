// Fusion + pooling + loss (see readout.cuh).
#include "readout.cuh"

namespace pamnet {

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.2f * x; }

// per node: sum over layers of softmax_{global,local}(leaky_relu(att)) . out   (models.py:207-213)
__global__ void readout_node_kernel(const ReadoutArgs a) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.n_nodes) return;
    float acc = 0.f;
    for (int l = 0; l < a.n_layer; ++l) {
        const float sg = leaky(a.att[(size_t)(2 * l) * a.n_nodes + n]);
        const float sl = leaky(a.att[(size_t)(2 * l + 1) * a.n_nodes + n]);
        const float m = fmaxf(sg, sl);
        const float eg = expf(sg - m), el = expf(sl - m);
        const float inv = 1.0f / (eg + el);
        acc += a.out[(size_t)(2 * l) * a.n_nodes + n] * (eg * inv) + a.out[(size_t)(2 * l + 1) * a.n_nodes + n] * (el * inv);
    }
    if (a.sign) acc *= a.sign[n];
    a.node_val[n] = acc;
}

// one warp per graph, fixed-order lane-strided sum then shuffle tree: deterministic pooling
__global__ void readout_pool_kernel(const ReadoutArgs a) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= a.n_graphs) return;
    const int s = a.gptr[g], e = a.gptr[g + 1];
    float v = 0.f;
    for (int n = s + lane; n < e; n += 32) v += a.node_val[n];
    v = warp_sum(v);
    if (lane == 0) a.pooled[g] = a.pool_mean ? v / (float)max(e - s, 1) : v;
}

int readout_forward(const ReadoutArgs& a, cudaStream_t st) {
    if (a.n_nodes > 0) {
        prof_begin(KC_READOUT, 0.0, st);
        readout_node_kernel<<<ceil_div(a.n_nodes, 128), 128, 0, st>>>(a);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    if (a.n_graphs > 0) {
        prof_begin(KC_READOUT, 0.0, st);
        readout_pool_kernel<<<ceil_div(a.n_graphs, 4), 128, 0, st>>>(a);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    return 0;
}

__global__ void readout_bwd_kernel(const ReadoutArgs a) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.n_nodes) return;
    const int g = a.n2g[n];
    float gn = a.g_pooled[g];
    if (a.pool_mean) gn /= (float)max(a.gptr[g + 1] - a.gptr[g], 1);
    if (a.sign) gn *= a.sign[n];
    for (int l = 0; l < a.n_layer; ++l) {
        const size_t ig = (size_t)(2 * l) * a.n_nodes + n, il = (size_t)(2 * l + 1) * a.n_nodes + n;
        const float ag = a.att[ig], al = a.att[il];
        const float sg = leaky(ag), sl = leaky(al);
        const float m = fmaxf(sg, sl);
        const float eg = expf(sg - m), el = expf(sl - m);
        const float inv = 1.0f / (eg + el);
        const float wg = eg * inv, wl = el * inv;
        const float og = a.out[ig], ol = a.out[il];
        a.g_out[ig] = gn * wg;
        a.g_out[il] = gn * wl;
        // softmax backward: g_s_i = w_i * (g_w_i - sum_j g_w_j w_j), g_w_i = gn * out_i
        const float dotp = gn * og * wg + gn * ol * wl;
        const float gsg = wg * (gn * og - dotp), gsl = wl * (gn * ol - dotp);
        a.g_att[ig] = gsg * (ag > 0.f ? 1.f : 0.2f);
        a.g_att[il] = gsl * (al > 0.f ? 1.f : 0.2f);
    }
}

int readout_backward(const ReadoutArgs& a, cudaStream_t st) {
    if (a.n_nodes == 0) return 0;
    prof_begin(KC_READOUT, 0.0, st);
    readout_bwd_kernel<<<ceil_div(a.n_nodes, 128), 128, 0, st>>>(a);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// single block: loss = mean(|out - y|) (kind 0) or mean((out - y)^2) (kind 1), grad_out = d loss / d out
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ out, const float* __restrict__ y, int64_t n,
                                                   int kind, float* __restrict__ loss, float* __restrict__ grad) {
    __shared__ float red[8];
    float s = 0.f;
    const float inv = 1.0f / (float)n;
    for (int64_t i = threadIdx.x; i < n; i += 256) {
        const float d = out[i] - y[i];
        if (kind == 0) {
            s += fabsf(d);
            grad[i] = (d > 0.f ? inv : (d < 0.f ? -inv : 0.f));
        } else {
            s += d * d;
            grad[i] = 2.0f * d * inv;
        }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        loss[0] = t * inv;
    }
}

int loss_forward_backward(const float* out, const float* y, int64_t n, int kind, float* loss, float* grad_out,
                          cudaStream_t st) {
    PAMNET_CHECK_ARG(n > 0, "loss: empty batch");
    prof_begin(KC_READOUT, 0.0, st);
    loss_kernel<<<1, 256, 0, st>>>(out, y, n, kind, loss, grad_out);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

}  // namespace pamnet
