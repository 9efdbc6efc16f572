// Parameter / workspace layouts and the forward / backward drivers of the PAMNet hot path.
#pragma once
#include <vector>

#include "common.cuh"

namespace pamnet {

constexpr int kParamAlign = 32;   // floats; every tensor of the flat parameter buffer starts 128 B aligned
constexpr int kMaxLayers = 8;

struct Lin { int64_t w = -1, b = -1; };

// one message-passing layer ("half" of a layer pair): global or local
struct HalfP {
    int64_t W = -1;                       // attention vector [D,1]      (global_message_passing.py:26)
    Lin x1, x2, res[3][2], out[3], W_out;
    Lin m, We;                            // global: mlp_m [D,3D], W_edge_attr [D,D]
    Lin m_ji, m_kj, sbf[2], lin_rbf, lin_rbf_out;   // local (m_kj doubles as mlp_m_jj of PAMNet_s)
};

struct ModelP {
    int64_t emb = -1, init_linear = -1, freq_g = -1, freq_l = -1;
    int n_embed = 0;
    Lin rbf_g, rbf_l, sbf1, sbf2;         // PAMNet_s: its single mlp_sbf is stored in sbf1
    HalfP g[kMaxLayers], l[kMaxLayers];
    int64_t total = 0;
    std::vector<int64_t> offsets, numel;  // state_dict order
};

int build_param_layout(const pamnet_config_t& cfg, ModelP* mp);

size_t workspace_bytes(const pamnet_config_t& cfg, const pamnet_sizes_t& sz);
int64_t debug_ws_offset(const pamnet_config_t& cfg, const pamnet_sizes_t& sz, const char* name, int half);

int model_forward(const pamnet_config_t& cfg, const pamnet_sizes_t& sz, const pamnet_sbf_consts_t& sbf,
                  const float* params, const float* node_in, const float* sign, const float* pos, void* plan_base,
                  void* plan_trip, void* workspace, size_t workspace_bytes, float* out, cudaStream_t st, cudaStream_t aux,
                  void* prepared = nullptr);

int model_backward(const pamnet_config_t& cfg, const pamnet_sizes_t& sz, const pamnet_sbf_consts_t& sbf,
                   const float* params, const float* node_in, const float* sign, const float* pos, void* plan_base,
                   void* plan_trip, void* workspace, size_t workspace_bytes, const float* grad_out, float* grad_params,
                   cudaStream_t st, cudaStream_t aux, void* prepared = nullptr);

// gradient buckets for an overlapped data-parallel all-reduce (model.cu)
void set_grad_buckets(int on);
int wait_grad_bucket(int half, cudaStream_t stream);
int grad_bucket_range(const pamnet_config_t& cfg, int half, int64_t* lo, int64_t* hi);

// k-major chain weights + contiguous projection blocks, produced ahead of model_forward (optional)
size_t prepared_weights_bytes(const pamnet_config_t& cfg);
int prepare_weights(const pamnet_config_t& cfg, const float* params, void* prepared, cudaStream_t st);

}  // namespace pamnet
