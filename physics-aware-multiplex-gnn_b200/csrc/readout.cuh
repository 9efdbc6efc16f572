// Fusion module + pooling (models.py:206-224) and the loss head (main_qm9.py:108), forward and backward.
#pragma once
#include "common.cuh"

namespace pamnet {

struct ReadoutArgs {
    int n_nodes, n_graphs, n_layer;
    int pool_mean;                 // rna: global_mean_pool (models.py:221), else global_add_pool
    const float* sign;             // PDBbind +-1 per node (models.py:125,218) or null
    const int32_t* gptr;           // node range per graph
    const int32_t* n2g;
    // heads: row 2*l = global layer l, row 2*l+1 = local layer l; each [N]
    const float* att;              // [2L, N]
    const float* out;              // [2L, N]
    float* node_val;               // [N] scratch (fused per-node value)
    float* pooled;                 // [G]
    // backward
    const float* g_pooled;         // [G]
    float* g_att;                  // [2L, N]
    float* g_out;                  // [2L, N]
};
int readout_forward(const ReadoutArgs& a, cudaStream_t st);
int readout_backward(const ReadoutArgs& a, cudaStream_t st);

int loss_forward_backward(const float* out, const float* y, int64_t n, int kind, float* loss, float* grad_out,
                          cudaStream_t st);

}  // namespace pamnet
