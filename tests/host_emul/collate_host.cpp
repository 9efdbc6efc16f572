// Serial host emulation of the device-side collation kernel (csrc/collate.cuh), one "thread" per block.
// Test infrastructure only.  Build: g++ -O1 -shared -fPIC.
#include "../../physics-aware-multiplex-gnn_b200/csrc/collate.cuh"

using namespace pamnet;

extern "C" int collate_host(const int64_t* table, int64_t n_ids, const int64_t* node_ptr, const int64_t* edge_ptr,
                            const float* x_all, const float* pos_all, const int64_t* ei_all, int64_t e_all,
                            const float* y_all, int64_t n_edges, float* x, float* pos, int64_t* edge_index,
                            int64_t* batch, float* y) {
    CollateArgs a;
    a.table = table; a.n_ids = n_ids; a.node_ptr = node_ptr; a.edge_ptr = edge_ptr; a.x_all = x_all; a.pos_all = pos_all;
    a.ei_all = ei_all; a.e_all = e_all; a.y_all = y_all; a.n_edges = n_edges; a.x = x; a.pos = pos;
    a.edge_index = edge_index; a.batch = batch; a.y = y;
    for (int g = (int)n_ids - 1; g >= 0; --g) collate_body(a, g);
    return 0;
}
