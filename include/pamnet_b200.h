/*
 * pamnet_b200.h -- C ABI of libpamnet_sm100.so: the PAMNet message-passing hot path on B200 (sm_100a).
 *
 * The reference (XieResearchGroup/Physics-aware-Multiplex-GNN) is pure Python and has no FFI; the operator
 * boundary this library sits under is the set of torch / third-party calls its nn.Modules make.  Each entry
 * point below names the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - plain pointers and sizes only; every buffer (inputs, outputs, plan, workspace) is owned by the caller
 *    and lives in device memory unless a parameter says "host"; the library never allocates, frees or
 *    retains a pointer, and keeps no global state besides the thread-local error string;
 *  - every function enqueues on `stream` (a cudaStream_t passed as void*) and returns without synchronising;
 *  - return 0 = OK, < 0 = argument / shape error, > 0 = cudaError_t; text via pamnet_last_error();
 *  - API-visible index tensors are int64 (as torch LongTensor); floating point is fp32;
 *  - variable-length results use count -> (caller reads the count, allocates) -> fill.
 */
#ifndef PAMNET_B200_H
#define PAMNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAMNET_ABI_VERSION 1

enum { PAMNET_QM9 = 0, PAMNET_PDBBIND = 1, PAMNET_RNA = 2 };          /* models.py:104,117,138 */
enum { PAMNET_SOURCE_TO_TARGET = 0, PAMNET_TARGET_TO_SOURCE = 1 };    /* models.py:13 `flow`   */

/* Config(dataset, dim, n_layer, cutoff_l, cutoff_g, flow) -- models.py:12-19; PAMNet vs PAMNet_s -- :21,:227 */
typedef struct pamnet_config {
    int32_t dataset;      /* PAMNET_QM9 | PAMNET_PDBBIND | PAMNET_RNA */
    int32_t dim;          /* 16, 32, 64 or 128 */
    int32_t n_layer;
    int32_t flow;
    int32_t simple;       /* 1 = PAMNet_s (one-hop only, models.py:227-353) */
    float cutoff_l;
    float cutoff_g;
} pamnet_config_t;

/* Spherical-basis constants, utils/sbf.py:13-26 (zeros, stored fp32 there) and :41-49 (normalisers);
 * index l*6+m, l < 7, m < 6.  Passed by the caller (host memory, copied into kernel arguments). */
typedef struct pamnet_sbf_consts {
    double zeros[42];
    double norm[42];
} pamnet_sbf_consts_t;

int pamnet_abi_version(void);
const char* pamnet_last_error(void);

/* ---- parameters -------------------------------------------------------------------------------------
 * One flat fp32 buffer holding every tensor of the reference module's state_dict, in state_dict order
 * (models.py:22-56, global_message_passing.py:9-31, local_message_passing.py:9-34).
 * pamnet_param_count returns the number of tensors, pamnet_param_total the number of floats;
 * pamnet_param_offsets fills offsets[count] (in floats) and numel[count].  Gradients use the same layout. */
int pamnet_param_count(const pamnet_config_t* cfg);
int64_t pamnet_param_total(const pamnet_config_t* cfg);
int pamnet_param_offsets(const pamnet_config_t* cfg, int64_t* offsets, int64_t* numel);

/* ---- graph construction -----------------------------------------------------------------------------
 * torch_cluster.radius(x, x, r, batch, batch, max_num_neighbors) as called at models.py:110,128,301,
 * canonical semantics of oracle/graph_ops.py:radius_pairs; `batch` non-decreasing.
 * count: deg[n] (int32 [n_nodes], neighbours kept for query n after the optional self-loop drop of
 *        models.py:63), ptr[n_nodes+1] exclusive scan, *total_dev = ptr[n_nodes] (device int64).
 * fill : edge_index[2*total] int64 (row = query, col = neighbour; ordered by (query, neighbour)). */
int pamnet_radius_count(const float* pos, const int64_t* batch, int64_t n_nodes, float r,
                        int32_t max_num_neighbors, int32_t drop_self, int32_t* deg, int32_t* ptr,
                        int64_t* total_dev, void* stream);
int pamnet_radius_fill(const float* pos, const int64_t* batch, int64_t n_nodes, float r,
                       int32_t max_num_neighbors, int32_t drop_self, const int32_t* ptr, int64_t total,
                       int64_t* edge_index, void* stream);
/* Cell-list ("grid-hash") variant for graphs beyond molecule size (PDBbind complexes; models.py:128): per-graph uniform
 * grid with cells >= r sized on the device to the caller's scratch, 27-cell queries, neighbours sorted by index -- the
 * SAME edge list as pamnet_radius_count / _fill, bit for bit.  `scratch` (pamnet_radius_grid_scratch_bytes) is written
 * by _count and read by _fill. */
size_t pamnet_radius_grid_scratch_bytes(int64_t n_nodes, int64_t n_graphs);
int pamnet_radius_grid_count(const float* pos, const int64_t* batch, int64_t n_nodes, int64_t n_graphs, float r,
                             int32_t max_nb, int32_t drop_self, void* scratch, size_t scratch_bytes, int32_t* deg,
                             int32_t* ptr, int64_t* total_dev, void* stream);
int pamnet_radius_grid_fill(const float* pos, const int64_t* batch, int64_t n_nodes, int64_t n_graphs, float r,
                            int32_t max_nb, int32_t drop_self, const void* scratch, const int32_t* ptr, int64_t total,
                            int64_t* edge_index, void* stream);

/* torch_cluster.knn(x, x, k, batch, batch) at models.py:143 (oracle/graph_ops.py:knn_pairs): per query the k
 * nearest points of its own graph (itself included), ascending distance, ties to the lower index.
 * nbr[n_nodes*k] int32 (-1 padded when the graph has fewer than k points), d2[n_nodes*k] fp32. */
int pamnet_knn(const float* pos, const int64_t* batch, int64_t n_nodes, int32_t k, int32_t* nbr, float* d2,
               void* stream);
/* models.py:144-157: drop self, keep dist <= cutoff; count then fill like the radius pair. */
int pamnet_knn_edges_count(const int32_t* nbr, const float* pos, int64_t n_nodes, int32_t k, float cutoff,
                           int32_t* deg, int32_t* ptr, int64_t* total_dev, void* stream);
int pamnet_knn_edges_fill(const int32_t* nbr, const float* pos, int64_t n_nodes, int32_t k, float cutoff,
                          const int32_t* ptr, int64_t total, int64_t* edge_index, void* stream);

/* torch_geometric.utils.remove_self_loops (models.py:63) / boolean edge masks (models.py:131-136):
 * order-preserving compaction of edge_index[2*n_edges] by (row != col) and, when pos != NULL,
 * |pos[col]-pos[row]| <= cutoff.  count writes keep[n_edges] (int32 0/1), ptr[n_edges+1], *total_dev. */
int pamnet_edge_filter_count(const int64_t* edge_index, int64_t n_edges, const float* pos, float cutoff,
                             int32_t* keep, int32_t* ptr, int64_t* total_dev, void* stream);
int pamnet_edge_filter_fill(const int64_t* edge_index, int64_t n_edges, const int32_t* keep, const int32_t* ptr,
                            int64_t total, int64_t* edge_index_out, void* stream);

/* PAMNet.indices (models.py:68-98): triplet (two-hop) and pair (one-hop) index vectors of a local graph.
 * count: needs scratch of pamnet_triplet_scratch_bytes(); writes counts_dev[2] = {T2, T1} (device int64).
 * fill : the ten int64 vectors in the reference's order; two-hop ones have T2 entries, pair ones T1. */
size_t pamnet_triplet_scratch_bytes(int64_t n_nodes, int64_t n_edges);
int pamnet_triplet_count(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, void* scratch,
                         int64_t* counts_dev, void* stream);
int pamnet_triplet_fill(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, const void* scratch,
                        int64_t t2, int64_t t1,
                        int64_t* idx_i, int64_t* idx_j, int64_t* idx_k, int64_t* idx_kj, int64_t* idx_ji,
                        int64_t* idx_i_pair, int64_t* idx_j1_pair, int64_t* idx_j2_pair,
                        int64_t* idx_jj_pair, int64_t* idx_ji_pair, void* stream);

/* ---- execution plan ---------------------------------------------------------------------------------
 * Internal, destination-sorted (CSR) form of the two edge lists and of the triplet lists, built once per
 * batch and reused by every layer (everything graph-structural is layer-invariant: models.py:104-188).
 * The plan is two opaque caller-owned device blobs: `plan_base` (sized from N, G, E_g, E_l; needed by
 * plan_count) and `plan_trip` (sized from T2 + T1, allocated once the counts are known).
 * plan_count : builds the CSR part, counts triplets -> counts_dev[2] = {T2, T1} (for PAMNet_s T2 = 0);
 * plan_fill  : fills the triplet part, edge lengths and the reverse (gather-keyed) triplet lists. */
typedef struct pamnet_sizes {
    int64_t n_nodes, n_graphs, n_edges_g, n_edges_l, n_t2, n_t1;
} pamnet_sizes_t;

int pamnet_plan_bytes(const pamnet_config_t* cfg, const pamnet_sizes_t* sz, size_t* base_bytes, size_t* trip_bytes);
int pamnet_plan_count(const pamnet_config_t* cfg, const pamnet_sizes_t* sz, const int64_t* edge_index_g,
                      const int64_t* edge_index_l, const int64_t* batch, void* plan_base, int64_t* counts_dev,
                      void* stream);
int pamnet_plan_fill(const pamnet_config_t* cfg, const pamnet_sizes_t* sz, const float* pos, void* plan_base,
                     void* plan_trip, void* stream);

/* Device-side collation -- replaces PyG's DataLoader collate + `data.to(device)` of
 * main_qm9.py:59-60,103-104 for a dataset that is resident in device memory: concatenates x / pos / y of the chosen
 * molecules, offsets each molecule's edge_index by the atoms before it and writes the graph id per atom, in one launch.
 *   table [3, n_ids] int64 (device): molecule id | first atom of the molecule inside the batch | first bond inside the batch
 *   dataset: node_ptr / edge_ptr [M + 1], x_all [sum n], pos_all [sum n, 3], ei_all [2, e_all] with atom ids relative to
 *   the molecule (as every Data object stores them), y_all [M];  outputs: x [N], pos [N, 3], edge_index [2, n_edges],
 *   batch [N], y [n_ids] with N / n_edges = the sums the host used for the offsets in `table`. */
int pamnet_collate(const int64_t* table, int64_t n_ids, const int64_t* node_ptr, const int64_t* edge_ptr,
                   const float* x_all, const float* pos_all, const int64_t* ei_all, int64_t e_all, const float* y_all,
                   int64_t n_edges, float* x, float* pos, int64_t* edge_index, int64_t* batch, float* y, void* stream);

/* The whole front end of PAMNet.forward (models.py:104-177: radius / kNN graph, self-loop and cutoff filters, the
 * destination-sorted plan, triplet lists, distances, angles) in ONE call.  The host must learn E_g, E_l and the triplet
 * counts to size the caller-owned buffers, so the call synchronises the stream two or three times internally (pinned
 * 64-byte read-backs) instead of returning to the host language in between.
 *   edge_index_in [2, n_edges_in]: data.edge_index (QM9 only, else NULL / 0).
 *   eg_buf / el_buf: int64 buffers of 2 * cap_eg / 2 * cap_el elements; on success they hold edge_index_g [2, E_g] and
 *   edge_index_l [2, E_l] packed with row stride E (the local list is NOT copied when nothing was filtered: then the
 *   plan refers to edge_index_in (QM9) or to eg_buf (PDBbind) and el_buf is untouched).
 *   plan_base / plan_trip: blobs of cap_base / cap_trip bytes (layout: pamnet_plan_bytes of the final sizes).
 *   scratch: pamnet_plan_build_scratch_bytes(cfg, n_nodes, n_edges_in, cap_eg) bytes.
 * Returns 0 and fills *sizes_out; returns 1 when a capacity was too small -- need[0..3] = {E_g, E_l, base bytes, trip
 * bytes} as far as they are known (0 = not reached): grow and call again; < 0 / > 1 are errors as everywhere. */
size_t pamnet_plan_build_scratch_bytes(const pamnet_config_t* cfg, int64_t n_nodes, int64_t n_edges_in, int64_t cap_eg);
int pamnet_plan_build(const pamnet_config_t* cfg, const float* pos, const int64_t* batch, int64_t n_nodes,
                      int64_t n_graphs, const int64_t* edge_index_in, int64_t n_edges_in, int32_t max_nb,
                      int64_t* eg_buf, int64_t cap_eg, int64_t* el_buf, int64_t cap_el, void* plan_base,
                      size_t cap_base, void* plan_trip, size_t cap_trip, void* scratch, size_t scratch_bytes,
                      pamnet_sizes_t* sizes_out, int64_t* need, void* stream);

/* ---- the hot path: PAMNet.forward / autograd backward (models.py:100-224, main_qm9.py:107-110) -----
 * node_in : QM9 / RNA: atom-type id per node as fp32 [n_nodes] (models.py:107,140);
 *           PDBbind  : 18 features per node [n_nodes,18] (models.py:119).
 * sign    : PDBbind only, +-1 per node (models.py:122-125), else NULL.
 * out     : [n_graphs].  workspace keeps what backward needs; it must stay untouched between the calls.
 * backward: grad_out [n_graphs] -> grad_params (flat, same layout as params; fully overwritten).
 * aux_stream: optional second stream (NULL = none): the x-independent GEMMs run on it, overlapped with the
 *             sequential layer loop on `stream`; both calls return with all their work ordered before later
 *             work on `stream`.
 * prepared_weights: optional (NULL = the forward call makes them itself, inside the workspace): a blob of
 *             pamnet_prepared_weights_bytes(cfg) bytes filled by pamnet_prepare_weights(cfg, params, blob, aux_stream)
 *             -- the k-major copies of the node-chain linears and the contiguous projection blocks, which depend on the
 *             parameters only.  Issued on aux_stream BEFORE the graph is built, they run while the host waits for the
 *             edge counts.  Pass the same blob to forward and backward of a step. */
size_t pamnet_workspace_bytes(const pamnet_config_t* cfg, const pamnet_sizes_t* sz);
size_t pamnet_prepared_weights_bytes(const pamnet_config_t* cfg);
int pamnet_prepare_weights(const pamnet_config_t* cfg, const float* params, void* prepared_weights, void* stream);
int pamnet_model_forward(const pamnet_config_t* cfg, const pamnet_sizes_t* sz, const pamnet_sbf_consts_t* sbf,
                         const float* params, const float* node_in, const float* sign, const float* pos,
                         void* plan_base, void* plan_trip, void* workspace, size_t workspace_bytes,
                         int32_t save_for_backward, float* out, void* stream, void* aux_stream,
                         void* prepared_weights);
int pamnet_model_backward(const pamnet_config_t* cfg, const pamnet_sizes_t* sz, const pamnet_sbf_consts_t* sbf,
                          const float* params, const float* node_in, const float* sign, const float* pos,
                          void* plan_base, void* plan_trip, void* workspace, size_t workspace_bytes,
                          const float* grad_out, float* grad_params, void* stream, void* aux_stream,
                          void* prepared_weights);

/* ---- gradient buckets: overlapped data-parallel all-reduce (the reference is single-process) ----
 * The flat gradient buffer is laid out per layer.  After pamnet_grad_buckets(1), every pamnet_model_backward records,
 * for each layer-half h (global layer l = 2l, local layer l = 2l + 1), events on its internal streams at the point where
 * the slice [lo, hi) of pamnet_grad_bucket_range is final; pamnet_wait_grad_bucket(h, comm_stream) makes comm_stream
 * wait for them, so a per-bucket ncclAllReduce can run while the remaining halves are still being differentiated.
 * Completion order: 2L-1, 2L-2, ..., 0; everything outside the layer slices (embeddings, basis MLPs, frequencies) is
 * final when the backward call's own stream is (wait for that stream). */
int pamnet_grad_buckets(int32_t enable);
int pamnet_wait_grad_bucket(int32_t half, void* stream);
int pamnet_grad_bucket_range(const pamnet_config_t* cfg, int32_t half, int64_t* lo, int64_t* hi);

/* ---- library-owned gradient all-reduce (data parallelism without Python between the buckets) ----
 * pamnet_comm_unique_id fills a 128-byte ncclUniqueId on one rank; the host side broadcasts it (torch.distributed) and
 * every rank calls pamnet_comm_init(id, rank, world) with its device current.  From then on pamnet_model_backward
 * averages the gradients over the ranks itself: one ncclAllReduce per two-layer bucket on a communication stream as
 * soon as the bucket's weight gradients have been issued, the head of the buffer last; the caller's stream waits for
 * the last one, so `grad_params` holds the averaged gradient when the call's work completes.  pamnet_comm_enable(0)
 * bypasses the collective (single-rank passes), pamnet_comm_destroy releases the communicator.  NCCL is resolved with
 * dlopen("libnccl.so.2") at run time; without it these calls fail and everything else works. */
int pamnet_comm_unique_id(void* out128);
int pamnet_comm_init(const void* id128, int32_t rank, int32_t world);
int pamnet_comm_enable(int32_t on);
int pamnet_comm_destroy(void);

/* L1 / MSE loss + its gradient w.r.t. the prediction in one launch (main_qm9.py:108 F.l1_loss,
 * main_pdbbind.py MSE): loss_dev[0] = mean(|out-y|) or mean((out-y)^2); grad_out[g] = d loss / d out[g]. */
int pamnet_loss(const float* out, const float* y, int64_t n, int32_t kind /*0 = L1, 1 = MSE*/, float* loss_dev,
                float* grad_out, void* stream);

/* Fused optimizer step on the flat buffers (replaces, per step, the reference's
 * clip_grad_norm_(model.parameters(), max_norm) -> optim.Adam.step() -> EMA.__call__(model)
 * (main_qm9.py:111-112,117, utils/ema.py:13-20) with two launches and no host synchronisation.
 *   params / grads / exp_avg / exp_avg_sq / ema_shadow: device fp32 buffers of n elements in the layout of
 *   pamnet_param_offsets (n = pamnet_param_total, a multiple of 4); ema_shadow may be NULL (no EMA).
 *   skip_ranges: n_skip (<= 4) pairs [begin, end) of element offsets of tensors that carry no gradient on this
 *   dataset (torch.optim skips parameters whose .grad is None); they are left untouched.
 *   step: 1-based Adam step count (bias corrections are evaluated in double on the host side of this call).
 *   max_norm <= 0 disables clipping; otherwise sumsq_dev (device double) receives the squared total gradient norm
 *   of this step, as clip_grad_norm_ would return it, and gradients are scaled by min(1, max_norm / (norm + 1e-6)).
 *   write_clipped_grad != 0 also stores the scaled gradients back, as clip_grad_norm_ does in place. */
int pamnet_optimizer_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, float* ema_shadow,
                          int64_t n, const int64_t* skip_ranges, int32_t n_skip, int64_t step, float lr, float beta1,
                          float beta2, float eps, float weight_decay, float max_norm, float ema_decay,
                          int32_t write_clipped_grad, double* sumsq_dev, void* stream);

/* ---- operator surface, also the unit-test hooks ----------------------------------
 * torch_scatter.scatter(src, index, dim=0, dim_size, 'add') (local_message_passing.py:50,54) for a
 * non-decreasing or arbitrary index: out[dim_size, width] zero-initialised then summed in index order per row
 * group when sorted, by atomics otherwise. */
int pamnet_scatter_add(const float* src, const int64_t* index, int64_t n_rows, int64_t width, int64_t dim_size,
                       float* out, void* stream);
/* BesselBasisLayer.forward (layers/basic.py:74-76): rbf[n_edges,16]. */
int pamnet_bessel_rbf(const float* dist, int64_t n_edges, const float* freq, float cutoff, float* rbf,
                      void* stream);
/* SphericalBasisLayer.forward (layers/basic.py:107-116) in two steps: the per-edge radial part
 * radial[n_edges,42] = u(x) N_lm j_l(z_lm x) (evaluated in double, rounded once), then
 * out[n_trip,42] = radial[gather[t]] * Y_l0(angle[t]) with gather = idx_kj / idx_jj_pair. */
int pamnet_sbf_radial(const pamnet_sbf_consts_t* sbf, const float* dist, int64_t n_edges, float cutoff,
                      float* radial, void* stream);
int pamnet_spherical_basis(const pamnet_sbf_consts_t* sbf, const float* radial, const float* angle,
                           const int64_t* gather, int64_t n_trip, float* out, void* stream);
/* y = act(x W^T + b): nn.Linear (+ SiLU, layers/basic.py:11-22); W [n_out, n_in] row-major. */
int pamnet_linear(const float* x, int64_t n_rows, int32_t n_in, int32_t n_out, const float* w, const float* b,
                  int32_t silu, float* y, void* stream);

/* Plain fp32 GEMM hook used by the unit tests: mode 0: C = A B^T, 1: C = A B, 2: C = A^T B (+ column sums of A
 * into dbias when non-null); ksplit > 1 accumulates into a zero-initialised C with atomics. */
int pamnet_gemm(int32_t mode, const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
                int32_t M, int32_t N, int32_t K, int32_t ksplit, float* dbias, void* stream);

/* Test / debugging aids: byte offsets of named buffers inside the caller-owned workspace and plan blobs. */
int64_t pamnet_debug_ws_offset(const pamnet_config_t* cfg, const pamnet_sizes_t* sz, const char* name, int32_t half);
int64_t pamnet_debug_plan_offset(const pamnet_sizes_t* sz, int32_t which, int32_t* in_trip);
/* Kernel launches issued by this library in this process so far (bench.py "gpu_launches"). */
int64_t pamnet_debug_launch_count(void);
/* Optional per-kernel-class CUDA-event timing on the launching stream (bench.py roofline): begin arms it,
 * end synchronises and fills ms / launches / algorithmic bytes per class (13 classes, order of KernelClass in
 * csrc/common.cuh); returns the class count.  Not thread-safe; off by default. */
void pamnet_debug_profile_begin(void);
int pamnet_debug_profile_end(double* ms, int64_t* launches, double* bytes, double* flops /* fp32-equivalent, GEMM classes; may be NULL */);
/* Same records as a timeline (call instead of profile_end): class, stream tag, start/end ms; returns the count. */
int pamnet_debug_profile_timeline(int32_t* cls, int32_t* stream_tag, float* t0_ms, float* t1_ms, int32_t cap);
/* clock64 timeline of CTA 0 of the last tensor-core GEMM launch (HOST buffer of n <= 256 slots).  Only libraries
 * built with -DPAMNET_TC_TRACE record it; otherwise returns -1. */
int pamnet_debug_tc_trace(long long* out, int32_t n);
/* Same for the node chain kernel: 8 stamps per stage of CTA 0 of the last launch (stage start, prologue done, weights
 * landed, barrier, multiply loop done, partials exchanged, epilogue done, stage end). */
int pamnet_debug_chain_trace(long long* out, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* PAMNET_B200_H */
