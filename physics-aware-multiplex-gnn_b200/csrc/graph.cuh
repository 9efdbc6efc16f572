// Graph construction entry points (host launchers) and the execution-plan view.
#pragma once
#include "common.cuh"

namespace pamnet {

// Destination-sorted form of both edge lists plus the merged triplet list.  "slot" = position in the
// incoming CSR of a graph; all per-edge tensors of the layer kernels are stored in slot order.
struct Plan {
    int32_t *n2g, *gptr;                                  // graph id per node; node range per graph
    int32_t *g_ptr, *g_src, *g_dst, *g_eid, *g_optr, *g_opos;   // global: in-CSR (ptr, source, dest, API edge id), out-CSR
    int32_t *l_ptr, *l_src, *l_dst, *l_eid, *l_optr, *l_opos;   // local, same
    int32_t *t_split, *t_cnt, *t_ptr;                     // per local slot: #two-hop, #total, segment start
    int32_t *t_gather, *t_owner;                          // per triplet: gathered slot, owning slot
    int32_t *tt_ptr, *tt_t;                               // triplets grouped by gathered slot (ascending id)
    float *t_angle, *dist_g, *dist_l;
    int32_t *tmp_a, *tmp_b, *tmp_c, *tmp_d, *tmp_e, *tmp_f, *cnt4, *tmp4, *t_tmp, *cnt;   // build scratch
};

void plan_layout(const pamnet_sizes_t& sz, void* base, void* trip, Plan* out, size_t* base_bytes, size_t* trip_bytes);

int scan_exclusive(const int32_t* in, int32_t* out, int64_t n, int64_t* total64, cudaStream_t st);
int radius_count(const float* pos, const int64_t* batch, int64_t n_nodes, float r, int max_nb, int drop_self,
                 int32_t* deg, int32_t* ptr, int64_t* total_dev, cudaStream_t st);
int radius_fill(const float* pos, const int64_t* batch, int64_t n_nodes, float r, int max_nb, int drop_self,
                const int32_t* ptr, int64_t total, int64_t* edge_index, cudaStream_t st);
// cell-list variant for large graphs (graph_grid.cu): same edge list bit for bit
size_t radius_grid_scratch_bytes(int64_t n_nodes, int64_t n_graphs);
int radius_grid_count(const float* pos, const int64_t* batch, int64_t n_nodes, int64_t n_graphs, float r, int max_nb,
                      int drop_self, void* scratch, size_t scratch_bytes, int32_t* deg, int32_t* ptr, int64_t* total_dev,
                      cudaStream_t st);
int radius_grid_fill(const float* pos, const int64_t* batch, int64_t n_nodes, int64_t n_graphs, float r, int max_nb,
                     int drop_self, const void* scratch, const int32_t* ptr, int64_t total, int64_t* edge_index,
                     cudaStream_t st);
bool radius_grid_preferred(int64_t n_nodes, int64_t n_graphs);
int knn(const float* pos, const int64_t* batch, int64_t n_nodes, int k, int32_t* nbr, float* d2, cudaStream_t st);
int knn_edges_count(const int32_t* nbr, const float* pos, int64_t n_nodes, int k, float cutoff, int32_t* deg,
                    int32_t* ptr, int64_t* total_dev, cudaStream_t st);
int knn_edges_fill(const int32_t* nbr, const float* pos, int64_t n_nodes, int k, float cutoff, const int32_t* ptr,
                   int64_t total, int64_t* edge_index, cudaStream_t st);
int edge_filter_count(const int64_t* ei, int64_t n_edges, const float* pos, float cutoff, int32_t* keep, int32_t* ptr,
                      int64_t* total_dev, cudaStream_t st);
int edge_filter_fill(const int64_t* ei, int64_t n_edges, const int32_t* keep, const int32_t* ptr, int64_t total,
                     int64_t* out, cudaStream_t st);
size_t triplet_scratch_bytes(int64_t n_nodes, int64_t n_edges);
int triplet_count(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, void* scratch, int64_t* counts_dev,
                  cudaStream_t st);
int triplet_fill(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, const void* scratch, int64_t* const* out,
                 cudaStream_t st);
int plan_count(const pamnet_config_t& cfg, const pamnet_sizes_t& sz, const int64_t* edge_index_g,
               const int64_t* edge_index_l, const int64_t* batch, void* plan_base, int64_t* counts_dev,
               cudaStream_t st);
int plan_fill(const pamnet_config_t& cfg, const pamnet_sizes_t& sz, const float* pos, void* plan_base, void* plan_trip,
              cudaStream_t st);

// per-molecule front end (front_mol.cuh / front_mol.cu): count pass, then -- after the host has read the totals -- fill pass
struct MolArgs;
int mol_count(const MolArgs& a, cudaStream_t st);
int mol_fill(const MolArgs& a, cudaStream_t st);

}  // namespace pamnet
