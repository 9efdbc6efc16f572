// Kernels of the per-molecule front end (front_mol.cuh): a range pass, then one thread block per molecule for the count
// and the fill pass -- three launches per batch.
#include "front_mol.cuh"
#include "graph.cuh"

namespace pamnet {

__global__ void mol_ranges_kernel(const MolArgs a) {
    pdl_wait();
    pdl_trigger();
    mol_ranges_body(a, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(kMolThreads) mol_count_kernel(const MolArgs a) {
    pdl_wait();
    pdl_trigger();
    __shared__ MolSmem s;
    mol_count_body(a, s, (int)blockIdx.x);
}

__global__ void __launch_bounds__(kMolThreads) mol_fill_kernel(const MolArgs a) {
    pdl_wait();
    pdl_trigger();
    __shared__ MolSmem s;
    mol_fill_body(a, s, (int)blockIdx.x);
}

int mol_count(const MolArgs& a, cudaStream_t st) {
    PAMNET_CHECK_ARG(a.n_graphs > 0 && a.n_graphs <= kMolGraphs, "mol_count: n_graphs=%lld", (long long)a.n_graphs);
    {
        const int64_t n = (a.n_nodes > a.n_edges_in ? a.n_nodes : a.n_edges_in) + 1;
        const int blocks = ceil_div(n, 256) < 4 * kNumSM ? ceil_div(n, 256) : 4 * kNumSM;
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(mol_ranges_kernel, dim3(blocks), dim3(256), 0, st, a);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    prof_begin(KC_GRAPH, 0.0, st);
    launch_pdl(mol_count_kernel, dim3((unsigned)a.n_graphs), dim3(kMolThreads), 0, st, a);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

int mol_fill(const MolArgs& a, cudaStream_t st) {
    PAMNET_CHECK_ARG(a.n_graphs > 0 && a.n_graphs <= kMolGraphs, "mol_fill: n_graphs=%lld", (long long)a.n_graphs);
    prof_begin(KC_GRAPH, 0.0, st);
    launch_pdl(mol_fill_kernel, dim3((unsigned)a.n_graphs), dim3(kMolThreads), 0, st, a);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

}  // namespace pamnet
