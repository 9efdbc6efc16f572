// Message-passing kernels: gather endpoint rows + edge/triplet attributes, SiLU gate, segment-sum into the
// destination.  One warp owns one destination segment of the destination-sorted (CSR) edge list, so the
// sums are sequential, deterministic and atomic-free (the reference's torch_scatter uses atomicAdd).
//
// With W [x_i ; x_j ; e] = W_i x_i + W_j x_j + W_e e the per-edge work is elementwise:
//   global (global_message_passing.py:52-56):  m_e = SiLU(Pi[i] + Pj[j] + Q[e]) * Tt[e],   h_i = x1_i + sum_e m_e
//   local  (local_message_passing.py:46-54) :  m_nb[e] = SiLU(Pi'[i] + Pj'[j] + Qkj[e]) * R[e]
//                                              m_e = (SiLU(Pi[i] + Pj[j] + Qji[e]) + sum_t m_nb[g_t] * q_t) * Rout[e]
// where P* are per-node projections (written by the node chain) and Q/Tt/R/Rout/q are x-independent
// per-edge / per-triplet projections computed for all layers at once (model.cu).
#pragma once
#include "common.cuh"

namespace pamnet {

struct GlobalMsgArgs {
    int n_nodes, n_edges;
    const int32_t *ptr, *src, *dst;  // incoming CSR (+ destination per slot)
    const float* P;                  // [N, 2D]  Pi | Pj
    const float* QT;                 // [E, ldq] at this layer's column block: Q | Tt
    int ldq;
    const float* x1;                 // [N, D]
    float* h;                        // [N, D]   x1 + aggregated messages
    // backward
    const float* g_h;                // [N, D]
    float* gQT;                      // [E, ldq] at this layer's block: grad Q (= grad z) | grad Tt
    float* g_P;                      // [N, 2D] zero-initialised, or null: grad of P accumulated here with fp32 reductions
                                     // (replaces the node_grad_gather pass: one launch less on the critical path)
};
int global_msg_fwd(int dim, const GlobalMsgArgs& a, int n_edges, cudaStream_t st);
int global_msg_bwd(int dim, const GlobalMsgArgs& a, int n_edges, cudaStream_t st);

struct LocalMsgArgs {
    int n_nodes, n_edges;
    const int32_t *ptr, *src, *dst;  // incoming CSR of the local graph (+ destination per slot)
    const int32_t *t_ptr, *t_gather; // triplet segments per slot, gathered slot per triplet
    const int32_t *tt_ptr, *tt_t, *t_owner;   // triplets grouped by gathered slot
    const float* P;                  // [N, 4D]  Pi_ji | Pj_ji | Pi_kj | Pj_kj
    const float* QR;                 // [E, ldq] at this layer's block: Qji | Qkj | R | Rout
    int ldq;
    const float* zq;                 // [T, ldt] at this layer's block: pre-activation of mlp_sbf's 2nd linear
    int ldt;
    const float* x1;
    float* m_nb;                     // [E, D]
    float* msum;                     // [E, D]   m_ji + m_other (kept for backward)
    float* h;
    // backward
    const float* g_h;
    float* g_P;                      // [N, 4D] zero-initialised, or null (see GlobalMsgArgs::g_P)
    float* g_s;                      // [E, D]   grad of (m_ji + m_other)
    float* gQR;                      // [E, ldq] at this layer's block: grad z_ji | grad z_kj | grad R | grad Rout
    float* gzq;                      // [T, ldt] at this layer's block: grad of zq
};
int local_edge_fwd(int dim, const LocalMsgArgs& a, cudaStream_t st);
int local_trip_fwd(int dim, const LocalMsgArgs& a, int n_trip, cudaStream_t st);
int local_msg_fwd(int dim, const LocalMsgArgs& a, int n_trip, cudaStream_t st);
int local_msg_bwd(int dim, const LocalMsgArgs& a, int n_trip, cudaStream_t st);
int local_trip_bwd(int dim, const LocalMsgArgs& a, int n_trip, cudaStream_t st);

// g_P[n, (2b)D..] = sum over incoming slots of gz_b ; g_P[n, (2b+1)D..] = sum over outgoing slots of gz_b
struct NodeGatherArgs {
    int n_nodes, n_blocks;           // n_blocks = 1 (global) or 2 (local: z_ji, z_kj)
    const int32_t *ptr, *optr, *opos;
    const float* gz;                 // [E, ldq] at this layer's block; block b at column b*D
    int ldq;
    float* g_P;                      // [N, 2*n_blocks*D]
};
int node_grad_gather(int dim, const NodeGatherArgs& a, int n_edges, cudaStream_t st);

// generic torch_scatter.scatter(src, index, dim=0, reduce='add') for the operator surface
int scatter_add_rows(const float* src, const int64_t* index, int64_t n_rows, int64_t width, int64_t dim_size,
                     float* out, cudaStream_t st);

}  // namespace pamnet
