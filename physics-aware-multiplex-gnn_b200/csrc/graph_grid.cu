// Cell-list ("grid-hash") radius graph for graphs beyond molecule size (PDBbind complexes, RNA-size structures).
//
// radius_kernel (graph.cu) lets every query scan its whole graph: right for 12-29 atom molecules (the scan is
// warp-uniform and shorter than a cell lookup), O(n^2) for a 3 000-atom structure.  Here every graph gets its own
// uniform grid over its bounding box with cells at least r wide, sized on the DEVICE to fit a fixed caller-owned scratch
// budget (no host read-back): atoms are counted per cell, offsets come from the exclusive scan, a cursor pass fills the
// per-cell atom lists, and a query only visits its 27 neighbouring cells.  The result is the SAME edge list, bit for bit,
// as the brute-force kernel: membership is decided by the identical canon_d2(q, n) <= r^2 test (the grid only prunes
// candidates that cannot pass it), and every query's neighbours are sorted by index before the max_num_neighbors cut
// ("the first max_nb found in ascending index order", torch_cluster semantics of oracle/graph_ops.py), so the atomic
// cursor order inside a cell never shows.
#include "graph.cuh"

#include <stdlib.h>

#include "geom.cuh"

namespace pamnet {
namespace {

struct GridBox {            // per graph; written by grid_dims_kernel
    float ox, oy, oz;       // origin (bounding-box minimum)
    float ix, iy, iz;       // 1 / cell size per axis
    int gx, gy, gz;         // cells per axis
    int base;               // first cell of this graph in the global cell arrays
};

// order-preserving float <-> uint mapping for atomicMin / atomicMax
__device__ __forceinline__ unsigned f2o(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

__global__ void grid_bbox_init_kernel(unsigned* __restrict__ bbox, int64_t n_graphs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_graphs * 6) bbox[i] = (i % 6 < 3) ? 0xFFFFFFFFu : 0u;
}
__global__ void grid_bbox_kernel(const float* __restrict__ pos, const int64_t* __restrict__ batch, int64_t n,
                                 unsigned* __restrict__ bbox) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned* b = bbox + batch[i] * 6;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const unsigned o = f2o(pos[3 * i + a]);
        atomicMin(b + a, o);
        atomicMax(b + 3 + a, o);
    }
}
// one thread per graph: cells per axis (cell >= 1.0001 r, total <= cells_per_graph) and the graph's cell base
__global__ void grid_dims_kernel(const unsigned* __restrict__ bbox, int64_t n_graphs, float r, int cells_per_graph,
                                 GridBox* __restrict__ box) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_graphs) return;
    GridBox b;
    float ext[3];
    int dims[3];
    const float cell_min = r * 1.0001f;
    for (int a = 0; a < 3; ++a) {
        const unsigned lo = bbox[g * 6 + a], hi = bbox[g * 6 + 3 + a];
        const float mn = (lo == 0xFFFFFFFFu) ? 0.f : o2f(lo), mx = (lo == 0xFFFFFFFFu) ? 0.f : o2f(hi);   // empty graph
        (&b.ox)[a] = mn;
        ext[a] = fmaxf(mx - mn, 0.f);
        dims[a] = max(1, (int)floorf(ext[a] / cell_min));
    }
    // shrink the grid (larger cells are still correct) until it fits the per-graph budget
    while ((long long)dims[0] * dims[1] * dims[2] > cells_per_graph) {
        int a = dims[0] >= dims[1] ? (dims[0] >= dims[2] ? 0 : 2) : (dims[1] >= dims[2] ? 1 : 2);
        dims[a] = max(1, dims[a] - max(1, dims[a] / 8));
    }
    for (int a = 0; a < 3; ++a) {
        const float cs = dims[a] > 1 ? ext[a] / (float)dims[a] : 0.f;
        (&b.ix)[a] = cs > 0.f ? 1.0f / cs : 0.f;
    }
    b.gx = dims[0]; b.gy = dims[1]; b.gz = dims[2];
    b.base = (int)(g * cells_per_graph);
    box[g] = b;
}
__device__ __forceinline__ void cell_of(const GridBox& b, float x, float y, float z, int& cx, int& cy, int& cz) {
    cx = min(b.gx - 1, max(0, (int)((x - b.ox) * b.ix)));
    cy = min(b.gy - 1, max(0, (int)((y - b.oy) * b.iy)));
    cz = min(b.gz - 1, max(0, (int)((z - b.oz) * b.iz)));
}
__global__ void grid_cell_count_kernel(const float* __restrict__ pos, const int64_t* __restrict__ batch, int64_t n,
                                       const GridBox* __restrict__ box, int32_t* __restrict__ cell_cnt,
                                       int32_t* __restrict__ atom_cell) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GridBox b = box[batch[i]];
    int cx, cy, cz;
    cell_of(b, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], cx, cy, cz);
    const int c = b.base + (cz * b.gy + cy) * b.gx + cx;
    atom_cell[i] = c;
    atomicAdd(cell_cnt + c, 1);
}
__global__ void grid_cell_fill_kernel(const int32_t* __restrict__ atom_cell, int64_t n, const int32_t* __restrict__ cell_ptr,
                                      int32_t* __restrict__ cursor, int32_t* __restrict__ cell_atoms) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = atom_cell[i];
    cell_atoms[cell_ptr[c] + atomicAdd(cursor + c, 1)] = (int32_t)i;
}

// One thread per query.  COUNT: deg[q] = kept neighbours.  FILL: the query's segment, ascending neighbour index.
template <bool FILL>
__global__ void radius_grid_kernel(const float* __restrict__ pos, const int64_t* __restrict__ batch, int64_t n_nodes,
                                   float r2, int max_nb, int drop_self, const GridBox* __restrict__ box,
                                   const int32_t* __restrict__ cell_ptr, const int32_t* __restrict__ cell_atoms,
                                   int32_t* __restrict__ deg, const int32_t* __restrict__ ptr, int64_t total,
                                   int64_t* __restrict__ edge_index) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_nodes) return;
    const int64_t g = batch[q];
    const GridBox b = box[g];
    const float qx = pos[3 * q], qy = pos[3 * q + 1], qz = pos[3 * q + 2];
    int cx, cy, cz;
    cell_of(b, qx, qy, qz, cx, cy, cz);
    int tot = 0, lt = 0;                   // in range (self included), and of those with a smaller index than q
    const int64_t w0 = FILL ? ptr[q] : 0;
    const int kept_q = FILL ? (int)(ptr[q + 1] - ptr[q]) : 0;
    int w = 0;
    bool overflow = false;                 // more than max_nb in range: the cut needs the brute-force order (rare)
    for (int dz = -1; dz <= 1 && !overflow; ++dz) {
        const int z = cz + dz;
        if (z < 0 || z >= b.gz) continue;
        for (int dy = -1; dy <= 1 && !overflow; ++dy) {
            const int y = cy + dy;
            if (y < 0 || y >= b.gy) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int x = cx + dx;
                if (x < 0 || x >= b.gx) continue;
                const int c = b.base + (z * b.gy + y) * b.gx + x;
                for (int k = cell_ptr[c]; k < cell_ptr[c + 1]; ++k) {
                    const int64_t n = cell_atoms[k];
                    const float d2 = canon_d2(qx, qy, qz, pos[3 * n], pos[3 * n + 1], pos[3 * n + 2]);
                    if (d2 <= r2) {
                        ++tot;
                        lt += n < q;
                        if (FILL && !(drop_self && n == q)) {
                            if (w < kept_q) edge_index[total + w0 + w] = n;
                            ++w;
                        }
                    }
                }
            }
        }
        if (FILL && tot > max_nb) overflow = true;
    }
    if (!FILL) {
        // "found < max_nb" counts self; self is among the first max_nb found iff fewer than max_nb smaller indices are in range
        deg[q] = min(tot, max_nb) - ((drop_self && lt < max_nb) ? 1 : 0);
        return;
    }
    if (tot > max_nb) {
        // exact brute-force order for this query (ascending scan of its graph, stop after max_nb found)
        const int64_t s = lower_bound_i64(batch, n_nodes, g);
        int found = 0;
        int64_t ww = w0;
        for (int64_t n = s; n < n_nodes && batch[n] == g && found < max_nb; ++n) {
            if (canon_d2(qx, qy, qz, pos[3 * n], pos[3 * n + 1], pos[3 * n + 2]) <= r2) {
                ++found;
                if (!(drop_self && n == q)) { edge_index[ww] = q; edge_index[total + ww] = n; ++ww; }
            }
        }
        return;
    }
    // insertion sort of the segment by neighbour index (tens of entries), then the row ids
    int64_t* col = edge_index + total + w0;
    for (int i = 1; i < kept_q; ++i) {
        const int64_t v = col[i];
        int j = i - 1;
        while (j >= 0 && col[j] > v) { col[j + 1] = col[j]; --j; }
        col[j + 1] = v;
    }
    for (int i = 0; i < kept_q; ++i) edge_index[w0 + i] = q;
}

struct GridScratch {
    unsigned* bbox;
    GridBox* box;
    int32_t *cell_cnt, *cell_ptr, *cursor, *atom_cell, *cell_atoms;
    int cells_per_graph;
    int64_t n_cells;
};
size_t grid_layout(int64_t n_nodes, int64_t n_graphs, void* base, GridScratch* out) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off += align_up(bytes ? bytes : 1);
        return p;
    };
    GridScratch s;
    int64_t budget = 4 * n_nodes > 4096 ? 4 * n_nodes : 4096;
    s.cells_per_graph = (int)(budget / (n_graphs > 0 ? n_graphs : 1));
    if (s.cells_per_graph < 1) s.cells_per_graph = 1;
    s.n_cells = (int64_t)s.cells_per_graph * (n_graphs > 0 ? n_graphs : 1);
    s.bbox = static_cast<unsigned*>(take(sizeof(unsigned) * 6 * n_graphs));
    s.box = static_cast<GridBox*>(take(sizeof(GridBox) * n_graphs));
    s.cell_cnt = static_cast<int32_t*>(take(sizeof(int32_t) * (s.n_cells + 1)));
    s.cell_ptr = static_cast<int32_t*>(take(sizeof(int32_t) * (s.n_cells + 1)));
    s.cursor = static_cast<int32_t*>(take(sizeof(int32_t) * (s.n_cells + 1)));
    s.atom_cell = static_cast<int32_t*>(take(sizeof(int32_t) * n_nodes));
    s.cell_atoms = static_cast<int32_t*>(take(sizeof(int32_t) * n_nodes));
    if (out) *out = s;
    return off;
}

}  // namespace

size_t radius_grid_scratch_bytes(int64_t n_nodes, int64_t n_graphs) { return grid_layout(n_nodes, n_graphs, nullptr, nullptr); }

// builds the cell lists into `scratch` (kept for radius_grid_fill), then counts like radius_count
int radius_grid_count(const float* pos, const int64_t* batch, int64_t n_nodes, int64_t n_graphs, float r, int max_nb,
                      int drop_self, void* scratch, size_t scratch_bytes, int32_t* deg, int32_t* ptr, int64_t* total_dev,
                      cudaStream_t st) {
    GridScratch s;
    const size_t want = grid_layout(n_nodes, n_graphs, scratch, &s);
    PAMNET_CHECK_ARG(scratch && scratch_bytes >= want, "radius_grid: scratch too small: %zu < %zu", scratch_bytes, want);
    if (n_nodes > 0) {
        const int T = 128;
        prof_begin(KC_GRAPH, 0.0, st);
        grid_bbox_init_kernel<<<ceil_div(6 * n_graphs, T), T, 0, st>>>(s.bbox, n_graphs);
        grid_bbox_kernel<<<ceil_div(n_nodes, T), T, 0, st>>>(pos, batch, n_nodes, s.bbox);
        grid_dims_kernel<<<ceil_div(n_graphs, T), T, 0, st>>>(s.bbox, n_graphs, r, s.cells_per_graph, s.box);
        prof_end(st);
        count_launch(); count_launch(); count_launch();
        PAMNET_CUDA(cudaMemsetAsync(s.cell_cnt, 0, sizeof(int32_t) * (s.n_cells + 1), st));
        PAMNET_CUDA(cudaMemsetAsync(s.cursor, 0, sizeof(int32_t) * (s.n_cells + 1), st));
        prof_begin(KC_GRAPH, 0.0, st);
        grid_cell_count_kernel<<<ceil_div(n_nodes, T), T, 0, st>>>(pos, batch, n_nodes, s.box, s.cell_cnt, s.atom_cell);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
        PAMNET_TRY(scan_exclusive(s.cell_cnt, s.cell_ptr, s.n_cells, nullptr, st));
        prof_begin(KC_GRAPH, 0.0, st);
        grid_cell_fill_kernel<<<ceil_div(n_nodes, T), T, 0, st>>>(s.atom_cell, n_nodes, s.cell_ptr, s.cursor, s.cell_atoms);
        radius_grid_kernel<false><<<ceil_div(n_nodes, T), T, 0, st>>>(pos, batch, n_nodes, r * r, max_nb, drop_self, s.box,
                                                                        s.cell_ptr, s.cell_atoms, deg, nullptr, 0, nullptr);
        prof_end(st);
        count_launch();
        PAMNET_LAUNCH_CHECK();
    }
    return scan_exclusive(deg, ptr, n_nodes, total_dev, st);
}

int radius_grid_fill(const float* pos, const int64_t* batch, int64_t n_nodes, int64_t n_graphs, float r, int max_nb,
                     int drop_self, const void* scratch, const int32_t* ptr, int64_t total, int64_t* edge_index,
                     cudaStream_t st) {
    if (n_nodes == 0 || total == 0) return 0;
    GridScratch s;
    grid_layout(n_nodes, n_graphs, const_cast<void*>(scratch), &s);
    prof_begin(KC_GRAPH, 0.0, st);
    radius_grid_kernel<true><<<ceil_div(n_nodes, 128), 128, 0, st>>>(pos, batch, n_nodes, r * r, max_nb, drop_self, s.box,
                                                                       s.cell_ptr, s.cell_atoms, nullptr, ptr, total, edge_index);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// atoms per graph from which the cell list wins over the per-graph scan (PAMNET_RADIUS=grid / brute forces either)
bool radius_grid_preferred(int64_t n_nodes, int64_t n_graphs) {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("PAMNET_RADIUS");
        mode = (e && strcmp(e, "grid") == 0) ? 1 : (e && strcmp(e, "brute") == 0) ? 2 : 0;
    }
    if (mode == 1) return true;
    if (mode == 2) return false;
    return n_graphs > 0 && n_nodes / n_graphs >= 192;
}

}  // namespace pamnet
