// Message-passing kernels (see message.cuh).  Warp-per-segment; a lane owns D/32 columns as float4/float2/float
// so every row access is one fully coalesced request; per-edge intermediates are recomputed in backward
// instead of being saved (the reference's autograd saves ~770 MB per QM9 step).
#include "message.cuh"

namespace pamnet {

constexpr int kMsgThreads = 128;
constexpr int kUnroll = 4;       // independent row loads kept in flight per warp (L2 latency ~1 us per dependent trip)   // 4 warps per CTA -> N/4 CTAs; QM9 bs=32 (N~600) gives ~150 CTAs on 148 SMs

#define ROW_FOR(i) _Pragma("unroll") for (int i = 0; i < RowVec<D>::C * RowVec<D>::V; ++i)

// row (edge slot / node / work item) of this thread: a warp per row, or 32 / LPR rows per warp for dim <= 32 (common.cuh)
template <int D>
__device__ __forceinline__ int msg_row() {
    return (blockIdx.x * (kMsgThreads / 32) + (threadIdx.x >> 5)) * RowVec<D>::RPW + (threadIdx.x & 31) / RowVec<D>::LPR;
}

// ---------------------------------------------------------------------------------------------
// global layer
// ---------------------------------------------------------------------------------------------
// One CTA (4 warps) per destination node: warp w takes incoming edges e0+w, e0+w+4, ... (kUnroll of them in
// flight), partial sums meet in shared memory and are added in warp order -> deterministic, and 4x the
// memory-level parallelism of a warp-per-node walk (620 nodes alone cannot hide L2 latency on 148 SMs).
template <int D>
__global__ void __launch_bounds__(kMsgThreads) global_msg_fwd_kernel(const GlobalMsgArgs a) {
    pdl_wait();
    pdl_trigger();
    __shared__ __align__(16) float part[kMsgThreads / 32][D];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = blockIdx.x;
    RowVec<D> pi, acc;
    pi.load(a.P + (size_t)n * 2 * D, lane);
    acc.zero();
    const int e0 = a.ptr[n], e1 = a.ptr[n + 1];
    constexpr int W = kMsgThreads / 32;
    for (int k = e0 + w; k < e1; k += W * kUnroll) {
        RowVec<D> pj[kUnroll], q[kUnroll], tt[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int kk = k + u * W;
            if (kk < e1) {
                const int s = a.src[kk];
                pj[u].load(a.P + (size_t)s * 2 * D + D, lane);
                q[u].load(a.QT + (size_t)kk * a.ldq, lane);
                tt[u].load(a.QT + (size_t)kk * a.ldq + D, lane);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (k + u * W < e1) {
                ROW_FOR(i) acc.v[i] += silu(pi.v[i] + pj[u].v[i] + q[u].v[i]) * tt[u].v[i];
            }
        }
    }
    acc.store(&part[w][0], lane);
    __syncthreads();
    if (w == 0) {
        RowVec<D> x;
        x.load(a.x1 + (size_t)n * D, lane);
#pragma unroll
        for (int ww = 0; ww < W; ++ww) {
            RowVec<D> p;
            p.load(&part[ww][0], lane);
            ROW_FOR(i) x.v[i] += p.v[i];
        }
        x.store(a.h + (size_t)n * D, lane);
    }
}

// dim <= 32: one lane group (D / 4 lanes) per destination node walks its incoming edges, kUnroll in flight; large graphs
// (the shape these dims are used on) have enough nodes to fill the GPU that way
template <int D>
__global__ void __launch_bounds__(kMsgThreads) global_msg_fwd_rows_kernel(const GlobalMsgArgs a) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int n = msg_row<D>();
    if (n >= a.n_nodes) return;
    RowVec<D> pi, acc;
    pi.load(a.P + (size_t)n * 2 * D, lane);
    acc.load(a.x1 + (size_t)n * D, lane);
    const int e0 = a.ptr[n], e1 = a.ptr[n + 1];
    for (int k = e0; k < e1; k += kUnroll) {
        RowVec<D> pj[kUnroll], q[kUnroll], tt[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (k + u < e1) {
                const int s = a.src[k + u];
                pj[u].load(a.P + (size_t)s * 2 * D + D, lane);
                q[u].load(a.QT + (size_t)(k + u) * a.ldq, lane);
                tt[u].load(a.QT + (size_t)(k + u) * a.ldq + D, lane);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (k + u < e1) {
                ROW_FOR(i) acc.v[i] += silu(pi.v[i] + pj[u].v[i] + q[u].v[i]) * tt[u].v[i];
            }
        }
    }
    acc.store(a.h + (size_t)n * D, lane);
}

// one warp per edge slot: every output row is per-edge, so the backward is embarrassingly edge-parallel
template <int D>
__global__ void __launch_bounds__(kMsgThreads) global_msg_bwd_kernel(const GlobalMsgArgs a) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int k = msg_row<D>();
    if (k >= a.n_edges) return;
    const int n = a.dst[k], s = a.src[k];
    RowVec<D> pi, g, pj, q, tt, gz, gt;
    pi.load(a.P + (size_t)n * 2 * D, lane);
    g.load(a.g_h + (size_t)n * D, lane);
    pj.load(a.P + (size_t)s * 2 * D + D, lane);
    q.load(a.QT + (size_t)k * a.ldq, lane);
    tt.load(a.QT + (size_t)k * a.ldq + D, lane);
    ROW_FOR(i) {
        const float z = pi.v[i] + pj.v[i] + q.v[i];
        const float sg = sigmoidf_(z);
        gt.v[i] = g.v[i] * (z * sg);                                     // grad Tt = g * SiLU(z)
        gz.v[i] = g.v[i] * tt.v[i] * (sg * (1.0f + z * (1.0f - sg)));   // grad z  = g * Tt * SiLU'(z)
    }
    gz.store(a.gQT + (size_t)k * a.ldq, lane);
    gt.store(a.gQT + (size_t)k * a.ldq + D, lane);
    if (a.g_P) {                     // z = Q_e + P_i[dst] + P_j[src]: grad z flows to both node rows
        gz.atomic_add(a.g_P + (size_t)n * 2 * D, lane);
        gz.atomic_add(a.g_P + (size_t)s * 2 * D + D, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// local layer
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kMsgThreads) local_edge_fwd_kernel(const LocalMsgArgs a) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int k = msg_row<D>();
    if (k >= a.n_edges) return;
    const int i = a.dst[k], j = a.src[k];
    RowVec<D> pi, pj, q, r;
    pi.load(a.P + (size_t)i * 4 * D + 2 * D, lane);
    pj.load(a.P + (size_t)j * 4 * D + 3 * D, lane);
    q.load(a.QR + (size_t)k * a.ldq + D, lane);
    r.load(a.QR + (size_t)k * a.ldq + 2 * D, lane);
    ROW_FOR(c) r.v[c] *= silu(pi.v[c] + pj.v[c] + q.v[c]);
    r.store(a.m_nb + (size_t)k * D, lane);
}

// one warp per edge slot: msum[k] = m_ji + sum_t m_nb[g_t] * SiLU(zq_t)   (local_message_passing.py:47-51)
template <int D>
__global__ void __launch_bounds__(kMsgThreads) local_trip_fwd_kernel(const LocalMsgArgs a) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int k = msg_row<D>();
    if (k >= a.n_edges) return;
    const int n = a.dst[k], j = a.src[k];
    RowVec<D> pi, pj, q, ms;
    pi.load(a.P + (size_t)n * 4 * D, lane);
    pj.load(a.P + (size_t)j * 4 * D + D, lane);
    q.load(a.QR + (size_t)k * a.ldq, lane);
    ROW_FOR(c) ms.v[c] = silu(pi.v[c] + pj.v[c] + q.v[c]);              // m_ji
    const int t0 = a.t_ptr[k], t1 = a.t_ptr[k + 1];
    for (int t = t0; t < t1; t += kUnroll) {                               // + m_other, in the reference's order
        RowVec<D> mn[kUnroll], zq[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (t + u < t1) {
                mn[u].load(a.m_nb + (size_t)a.t_gather[t + u] * D, lane);
                zq[u].load(a.zq + (size_t)(t + u) * a.ldt, lane);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (t + u < t1) {
                ROW_FOR(c) ms.v[c] += mn[u].v[c] * silu(zq[u].v[c]);
            }
        }
    }
    ms.store(a.msum + (size_t)k * D, lane);
}

// one warp per destination node: h = x1 + sum_k msum[k] * Rout[k]   (local_message_passing.py:53-54)
template <int D>
__global__ void __launch_bounds__(kMsgThreads) local_msg_fwd_kernel(const LocalMsgArgs a) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int n = msg_row<D>();
    if (n >= a.n_nodes) return;
    RowVec<D> acc;
    acc.load(a.x1 + (size_t)n * D, lane);
    const int e0 = a.ptr[n], e1 = a.ptr[n + 1];
    for (int k = e0; k < e1; k += kUnroll) {
        RowVec<D> ms[kUnroll], ro[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (k + u < e1) {
                ms[u].load(a.msum + (size_t)(k + u) * D, lane);
                ro[u].load(a.QR + (size_t)(k + u) * a.ldq + 3 * D, lane);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (k + u < e1) {
                ROW_FOR(c) acc.v[c] += ms[u].v[c] * ro[u].v[c];
            }
        }
    }
    acc.store(a.h + (size_t)n * D, lane);
}

// one warp per edge slot: grad Rout, grad (m_ji + m_other) = g_s, grad z_ji, and grad of the triplet gate zq
template <int D>
__global__ void __launch_bounds__(kMsgThreads) local_msg_bwd_kernel(const LocalMsgArgs a) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int k = msg_row<D>();
    if (k >= a.n_edges) return;
    const int n = a.dst[k], j = a.src[k];
    RowVec<D> pi, g, ro, ms, gs, gro, pj, q, gz;
    pi.load(a.P + (size_t)n * 4 * D, lane);
    g.load(a.g_h + (size_t)n * D, lane);
    ro.load(a.QR + (size_t)k * a.ldq + 3 * D, lane);
    ms.load(a.msum + (size_t)k * D, lane);
    pj.load(a.P + (size_t)j * 4 * D + D, lane);
    q.load(a.QR + (size_t)k * a.ldq, lane);
    ROW_FOR(c) {
        gro.v[c] = g.v[c] * ms.v[c];
        gs.v[c] = g.v[c] * ro.v[c];
        gz.v[c] = gs.v[c] * dsilu(pi.v[c] + pj.v[c] + q.v[c]);
    }
    gro.store(a.gQR + (size_t)k * a.ldq + 3 * D, lane);
    gs.store(a.g_s + (size_t)k * D, lane);
    gz.store(a.gQR + (size_t)k * a.ldq, lane);
    if (a.g_P) {
        gz.atomic_add(a.g_P + (size_t)n * 4 * D, lane);
        gz.atomic_add(a.g_P + (size_t)j * 4 * D + D, lane);
    }
    const int t0 = a.t_ptr[k], t1 = a.t_ptr[k + 1];
    for (int t = t0; t < t1; t += kUnroll) {
        RowVec<D> mn[kUnroll], zq[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (t + u < t1) {
                mn[u].load(a.m_nb + (size_t)a.t_gather[t + u] * D, lane);
                zq[u].load(a.zq + (size_t)(t + u) * a.ldt, lane);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (t + u < t1) {
                ROW_FOR(c) zq[u].v[c] = gs.v[c] * mn[u].v[c] * dsilu(zq[u].v[c]);   // grad zq = g_s m_nb SiLU'(zq)
                zq[u].store(a.gzq + (size_t)(t + u) * a.ldt, lane);
            }
        }
    }
}

// per gathered slot k': grad m_nb = sum over the triplets that gathered k' (ascending triplet id), then
// back through m_nb = SiLU(z_kj) * R
template <int D>
__global__ void __launch_bounds__(kMsgThreads) local_trip_bwd_kernel(const LocalMsgArgs a) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int k = msg_row<D>();
    if (k >= a.n_edges) return;
    RowVec<D> gm;
    gm.zero();
    const int u0 = a.tt_ptr[k], u1 = a.tt_ptr[k + 1];
    for (int u = u0; u < u1; u += kUnroll) {
        RowVec<D> gs[kUnroll], zq[kUnroll];
#pragma unroll
        for (int w = 0; w < kUnroll; ++w) {
            if (u + w < u1) {
                const int t = a.tt_t[u + w];
                gs[w].load(a.g_s + (size_t)a.t_owner[t] * D, lane);
                zq[w].load(a.zq + (size_t)t * a.ldt, lane);
            }
        }
#pragma unroll
        for (int w = 0; w < kUnroll; ++w) {
            if (u + w < u1) {
                ROW_FOR(c) gm.v[c] += gs[w].v[c] * silu(zq[w].v[c]);
            }
        }
    }
    const int i = a.dst[k], j = a.src[k];
    RowVec<D> pi, pj, q, r, gz, gr;
    pi.load(a.P + (size_t)i * 4 * D + 2 * D, lane);
    pj.load(a.P + (size_t)j * 4 * D + 3 * D, lane);
    q.load(a.QR + (size_t)k * a.ldq + D, lane);
    r.load(a.QR + (size_t)k * a.ldq + 2 * D, lane);
    ROW_FOR(c) {
        const float z = pi.v[c] + pj.v[c] + q.v[c];
        const float sg = sigmoidf_(z);
        gr.v[c] = gm.v[c] * (z * sg);
        gz.v[c] = gm.v[c] * r.v[c] * (sg * (1.0f + z * (1.0f - sg)));
    }
    gz.store(a.gQR + (size_t)k * a.ldq + D, lane);
    gr.store(a.gQR + (size_t)k * a.ldq + 2 * D, lane);
    if (a.g_P) {
        gz.atomic_add(a.g_P + (size_t)i * 4 * D + 2 * D, lane);
        gz.atomic_add(a.g_P + (size_t)j * 4 * D + 3 * D, lane);
    }
}

// one warp per (node, task): task 2b = sum of gz_b over incoming slots, 2b+1 = over outgoing slots
template <int D>
__global__ void __launch_bounds__(kMsgThreads) node_grad_gather_kernel(const NodeGatherArgs a) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int item = msg_row<D>();
    const int ntask = 2 * a.n_blocks;
    if (item >= a.n_nodes * ntask) return;
    const int n = item / ntask, task = item % ntask, b = task >> 1;
    const bool outgoing = task & 1;
    const int32_t* ptr = outgoing ? a.optr : a.ptr;
    RowVec<D> s;
    s.zero();
    for (int k = ptr[n], k1 = ptr[n + 1]; k < k1; k += kUnroll) {
        RowVec<D> v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (k + u < k1) {
                const int slot = outgoing ? a.opos[k + u] : k + u;
                v[u].load(a.gz + (size_t)slot * a.ldq + b * D, lane);
            }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (k + u < k1) { ROW_FOR(c) s.v[c] += v[u].v[c]; }
    }
    s.store(a.g_P + (size_t)n * ntask * D + task * D, lane);
}

#define DISPATCH_DIM(dim, KERNEL, rows, args, CLS, BYTES)                                             \
    do {                                                                                              \
        if ((rows) <= 0) return 0;                                                                    \
        const int grid = ceil_div((rows), kMsgThreads / 32);            /* a warp per row */          \
        prof_begin(CLS, BYTES, st);                                                                   \
        switch (dim) {                                                                                \
            case 128: launch_pdl(KERNEL<128>, dim3(grid), dim3(kMsgThreads), 0, st, args); break;     \
            case 64:  launch_pdl(KERNEL<64>, dim3(grid), dim3(kMsgThreads), 0, st, args); break;      \
            case 32:  launch_pdl(KERNEL<32>, dim3(ceil_div(grid, RowVec<32>::RPW)), dim3(kMsgThreads), 0, st, args); break; \
            case 16:  launch_pdl(KERNEL<16>, dim3(ceil_div(grid, RowVec<16>::RPW)), dim3(kMsgThreads), 0, st, args); break; \
            default: set_error("unsupported dim %d (16, 32, 64, 128)", dim); return -1;               \
        }                                                                                             \
        prof_end(st);                                                                                 \
        PAMNET_LAUNCH_CHECK();                                                                        \
        return 0;                                                                                     \
    } while (0)

// algorithmic bytes: every operand row touched once (row = 4*D bytes), index arrays as int32
static double nbytes(int dim, double node_rows, double edge_rows, double trip_rows, double idx) {
    return 4.0 * dim * (node_rows + edge_rows + trip_rows) + 4.0 * idx;
}

int global_msg_fwd(int dim, const GlobalMsgArgs& a, int n_edges, cudaStream_t st) {
    if (dim <= 32) {
        if (a.n_nodes <= 0) return 0;
        const double bytes = nbytes(dim, 3.0 * a.n_nodes, 3.0 * n_edges, 0, a.n_nodes + n_edges);
        prof_begin(KC_GLOBAL_MSG_FWD, bytes, st);
        if (dim == 32)
            launch_pdl(global_msg_fwd_rows_kernel<32>, dim3(ceil_div(a.n_nodes, (kMsgThreads / 32) * RowVec<32>::RPW)), dim3(kMsgThreads), 0, st, a);
        else if (dim == 16)
            launch_pdl(global_msg_fwd_rows_kernel<16>, dim3(ceil_div(a.n_nodes, (kMsgThreads / 32) * RowVec<16>::RPW)), dim3(kMsgThreads), 0, st, a);
        else { set_error("unsupported dim %d (16, 32, 64, 128)", dim); return -1; }
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
        return 0;
    }
    // one CTA per node: rows = n_nodes * warps-per-CTA so the dispatch macro's grid = n_nodes
    DISPATCH_DIM(dim, global_msg_fwd_kernel, a.n_nodes * (kMsgThreads / 32), a, KC_GLOBAL_MSG_FWD,
                 nbytes(dim, 3.0 * a.n_nodes, 3.0 * n_edges, 0, a.n_nodes + n_edges));
}
int global_msg_bwd(int dim, const GlobalMsgArgs& a, int n_edges, cudaStream_t st) {
    DISPATCH_DIM(dim, global_msg_bwd_kernel, a.n_edges, a, KC_GLOBAL_MSG_BWD,
                 nbytes(dim, 2.0 * a.n_nodes, 5.0 * n_edges, 0, a.n_nodes + n_edges));
}
int local_edge_fwd(int dim, const LocalMsgArgs& a, cudaStream_t st) {
    DISPATCH_DIM(dim, local_edge_fwd_kernel, a.n_edges, a, KC_LOCAL_EDGE_FWD,
                 nbytes(dim, 0, 5.0 * a.n_edges, 0, 2.0 * a.n_edges));
}
int local_trip_fwd(int dim, const LocalMsgArgs& a, int n_trip, cudaStream_t st) {
    DISPATCH_DIM(dim, local_trip_fwd_kernel, a.n_edges, a, KC_LOCAL_MSG_FWD,
                 nbytes(dim, 0, 4.0 * a.n_edges, 2.0 * n_trip, 3.0 * a.n_edges + n_trip));
}
int local_msg_fwd(int dim, const LocalMsgArgs& a, int n_trip, cudaStream_t st) {
    (void)n_trip;
    DISPATCH_DIM(dim, local_msg_fwd_kernel, a.n_nodes, a, KC_LOCAL_MSG_FWD,
                 nbytes(dim, 2.0 * a.n_nodes, 2.0 * a.n_edges, 0, a.n_nodes + a.n_edges));
}
int local_msg_bwd(int dim, const LocalMsgArgs& a, int n_trip, cudaStream_t st) {
    DISPATCH_DIM(dim, local_msg_bwd_kernel, a.n_edges, a, KC_LOCAL_MSG_BWD,
                 nbytes(dim, 2.0 * a.n_nodes, 7.0 * a.n_edges, 3.0 * n_trip, a.n_nodes + 2.0 * a.n_edges + n_trip));
}
int local_trip_bwd(int dim, const LocalMsgArgs& a, int n_trip, cudaStream_t st) {
    DISPATCH_DIM(dim, local_trip_bwd_kernel, a.n_edges, a, KC_LOCAL_TRIP_BWD,
                 nbytes(dim, 0, 6.0 * a.n_edges, 2.0 * n_trip, 3.0 * a.n_edges + 2.0 * n_trip));
}
int node_grad_gather(int dim, const NodeGatherArgs& a, int n_edges, cudaStream_t st) {
    DISPATCH_DIM(dim, node_grad_gather_kernel, a.n_nodes * 2 * a.n_blocks, a, KC_NODE_GATHER,
                 nbytes(dim, 2.0 * a.n_blocks * a.n_nodes, 2.0 * a.n_blocks * n_edges, 0, 2.0 * a.n_nodes + n_edges));
}

// ---------------------------------------------------------------------------------------------
// generic scatter-add (operator surface; index arbitrary -> atomics)
// ---------------------------------------------------------------------------------------------
__global__ void scatter_add_kernel(const float* __restrict__ src, const int64_t* __restrict__ index, int64_t n_rows,
                                   int64_t width, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * width) return;
    const int64_t r = i / width, c = i % width;
    atomicAdd(&out[index[r] * width + c], src[i]);
}

int scatter_add_rows(const float* src, const int64_t* index, int64_t n_rows, int64_t width, int64_t dim_size,
                     float* out, cudaStream_t st) {
    PAMNET_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * dim_size * width, st));
    if (n_rows * width == 0) return 0;
    scatter_add_kernel<<<ceil_div(n_rows * width, 256), 256, 0, st>>>(src, index, n_rows, width, out);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

}  // namespace pamnet
