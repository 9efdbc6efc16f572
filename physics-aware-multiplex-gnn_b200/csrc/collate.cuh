// Device-side collation of molecules that live in HBM: what PyG's DataLoader collate does on the
// host for every batch (main_qm9.py:59-60 DataLoader(...), :103-104 `for data in loader: data.to(device)`) -- concatenate
// x / pos / y, offset every molecule's edge_index by the number of atoms before it, write the graph id per atom -- as ONE
// launch over a dataset resident on the GPU (all of QM9 is ~60 MB).  The host sends only the molecule ids and their
// offsets inside the batch ([3, G] int64).  One thread block per molecule; block-strided loops without barriers, so the
// body also compiles for the host with one "thread" per block (tests/host_emul).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PC_HD __device__ __forceinline__
#define PC_TID ((int)threadIdx.x)
#define PC_NT ((int)blockDim.x)
#else
#define PC_HD static inline
#define PC_TID 0
#define PC_NT 1
#endif

namespace pamnet {

struct CollateArgs {
    const int64_t* table;          // [3, n_ids]: molecule id | first atom of the molecule inside the batch | first bond
    int64_t n_ids;
    const int64_t* node_ptr;       // dataset: [M + 1] first atom of every molecule
    const int64_t* edge_ptr;       // dataset: [M + 1] first bond-list entry of every molecule
    const float* x_all;            // dataset: [sum n]
    const float* pos_all;          // dataset: [sum n, 3]
    const int64_t* ei_all;         // dataset: [2, e_all], atom ids INSIDE the molecule (as each Data object stores them)
    int64_t e_all;
    const float* y_all;            // dataset: [M]
    int64_t n_edges;               // bonds in the batch (row stride of edge_index)
    float* x;                      // [N]
    float* pos;                    // [N, 3]
    int64_t* edge_index;           // [2, n_edges]
    int64_t* batch;                // [N]
    float* y;                      // [n_ids]
};

PC_HD void collate_body(const CollateArgs& A, int g) {
    const int64_t id = A.table[g], n0 = A.table[A.n_ids + g], e0 = A.table[2 * A.n_ids + g];
    const int64_t s = A.node_ptr[id], n = A.node_ptr[id + 1] - s;
    const int64_t es = A.edge_ptr[id], ne = A.edge_ptr[id + 1] - es;
    for (int64_t i = PC_TID; i < n; i += PC_NT) {
        A.x[n0 + i] = A.x_all[s + i];
        A.batch[n0 + i] = g;
    }
    for (int64_t i = PC_TID; i < 3 * n; i += PC_NT) A.pos[3 * n0 + i] = A.pos_all[3 * s + i];
    for (int64_t e = PC_TID; e < ne; e += PC_NT) {
        A.edge_index[e0 + e] = A.ei_all[es + e] + n0;
        A.edge_index[A.n_edges + e0 + e] = A.ei_all[A.e_all + es + e] + n0;
    }
    if (PC_TID == 0) A.y[g] = A.y_all[id];
}

#ifdef __CUDACC__
int collate(const CollateArgs& a, cudaStream_t st);
#endif

}  // namespace pamnet
