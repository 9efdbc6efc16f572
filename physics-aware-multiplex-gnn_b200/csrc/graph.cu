// Graph construction on the device: radius / kNN neighbour search, edge filtering, destination-sorted CSR,
// triplet (two-hop) and pair (one-hop) enumeration, and the execution plan the layer kernels consume.
//
// Reference behaviour restated (no reference code exists for these -- they are third-party calls):
//   torch_cluster.radius / knn        models.py:110,128,143,301     -> radius_*, knn_*
//   remove_self_loops + dist masks    models.py:62-66,131-136,148-157 -> edge_filter_*, knn_edges_*
//   SparseTensor row gathers          models.py:68-98               -> incoming CSR walks (triplet_*, plan_*)
// Canonical ordering and tie rules: oracle/graph_ops.py.
#include "graph.cuh"
#include "geom.cuh"

namespace pamnet {

// ---------------------------------------------------------------------------------------------
// single-block exclusive scan (n up to a few million; graph sizes here are << that)
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;

__global__ void __launch_bounds__(kScanThreads) scan_exclusive_kernel(const int32_t* __restrict__ in,
                                                                      int32_t* __restrict__ out, int64_t n,
                                                                      int64_t* __restrict__ total64) {
    pdl_wait();
    pdl_trigger();
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += (int64_t)kScanThreads * kScanItems) {
        int32_t v[kScanItems];
        int32_t local = 0;
        const int64_t i0 = base + (int64_t)tid * kScanItems;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            v[k] = (i0 + k < n) ? in[i0 + k] : 0;
            local += v[k];
        }
        int32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int32_t w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;   // inclusive over warps
        }
        __syncthreads();
        const int32_t carry = carry_s;
        int32_t excl = carry + incl - local + (wid > 0 ? warp_tot[wid - 1] : 0);
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (i0 + k < n) out[i0 + k] = excl;
            excl += v[k];
        }
        __syncthreads();
        if (tid == kScanThreads - 1) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (tid == 0) {
        out[n] = carry_s;
        if (total64) *total64 = carry_s;
    }
}

int scan_exclusive(const int32_t* in, int32_t* out, int64_t n, int64_t* total64, cudaStream_t st) {
    prof_begin(KC_GRAPH, 0.0, st);
    launch_pdl(scan_exclusive_kernel, dim3(1), dim3(kScanThreads), 0, st, in, out, n, total64);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// graph segments from a non-decreasing batch vector
// ---------------------------------------------------------------------------------------------
// One thread per query; the scan over the query's own graph reads pos[n] warp-uniformly for small molecules.
template <bool FILL>
__global__ void radius_kernel(const float* __restrict__ pos, const int64_t* __restrict__ batch, int64_t n_nodes,
                              float r2, int max_nb, int drop_self, int32_t* __restrict__ deg,
                              const int32_t* __restrict__ ptr, int64_t total, int64_t* __restrict__ edge_index) {
    pdl_wait();
    pdl_trigger();
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_nodes) return;
    const int64_t g = batch[q];
    const int64_t s = lower_bound_i64(batch, n_nodes, g);
    const float qx = pos[3 * q], qy = pos[3 * q + 1], qz = pos[3 * q + 2];
    int found = 0, kept = 0;
    int64_t w = FILL ? ptr[q] : 0;
    for (int64_t n = s; n < n_nodes && batch[n] == g && found < max_nb; ++n) {
        float d2 = canon_d2(qx, qy, qz, pos[3 * n], pos[3 * n + 1], pos[3 * n + 2]);
        if (d2 <= r2) {
            ++found;
            if (!(drop_self && n == q)) {
                if (FILL) {
                    edge_index[w] = q;
                    edge_index[total + w] = n;
                    ++w;
                }
                ++kept;
            }
        }
    }
    if (!FILL) deg[q] = kept;
}

int radius_count(const float* pos, const int64_t* batch, int64_t n_nodes, float r, int max_nb, int drop_self,
                 int32_t* deg, int32_t* ptr, int64_t* total_dev, cudaStream_t st) {
    if (n_nodes > 0) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(radius_kernel<false>, dim3(ceil_div(n_nodes, 128)), dim3(128), 0, st, pos, batch, n_nodes, r * r, max_nb, drop_self,
                                                                     deg, nullptr, 0, nullptr);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    return scan_exclusive(deg, ptr, n_nodes, total_dev, st);
}

int radius_fill(const float* pos, const int64_t* batch, int64_t n_nodes, float r, int max_nb, int drop_self,
                const int32_t* ptr, int64_t total, int64_t* edge_index, cudaStream_t st) {
    if (n_nodes == 0 || total == 0) return 0;
    prof_begin(KC_GRAPH, 0.0, st);
    launch_pdl(radius_kernel<true>, dim3(ceil_div(n_nodes, 128)), dim3(128), 0, st, pos, batch, n_nodes, r * r, max_nb, drop_self,
                                                                nullptr, ptr, total, edge_index);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// kNN: one thread per query, sorted candidate list in shared memory ([slot][thread] layout).
// ---------------------------------------------------------------------------------------------
constexpr int kKnnThreads = 64;

__global__ void __launch_bounds__(kKnnThreads) knn_kernel(const float* __restrict__ pos,
                                                          const int64_t* __restrict__ batch, int64_t n_nodes, int k,
                                                          int32_t* __restrict__ nbr, float* __restrict__ d2out) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ unsigned char smem_raw[];
    float* sd = reinterpret_cast<float*>(smem_raw);                 // [k][kKnnThreads]
    int32_t* si = reinterpret_cast<int32_t*>(sd + (size_t)k * kKnnThreads);
    const int t = threadIdx.x;
    const int64_t q = (int64_t)blockIdx.x * kKnnThreads + t;
    if (q >= n_nodes) return;
    const int64_t g = batch[q];
    const int64_t s = lower_bound_i64(batch, n_nodes, g);
    const float qx = pos[3 * q], qy = pos[3 * q + 1], qz = pos[3 * q + 2];
    int cnt = 0;
    for (int64_t n = s; n < n_nodes && batch[n] == g; ++n) {
        float d2 = canon_d2(qx, qy, qz, pos[3 * n], pos[3 * n + 1], pos[3 * n + 2]);
        if (cnt == k && !(d2 < sd[(k - 1) * kKnnThreads + t])) continue;   // ties keep the earlier (lower) index
        int p = (cnt < k) ? cnt : k - 1;
        while (p > 0 && sd[(p - 1) * kKnnThreads + t] > d2) {
            sd[p * kKnnThreads + t] = sd[(p - 1) * kKnnThreads + t];
            si[p * kKnnThreads + t] = si[(p - 1) * kKnnThreads + t];
            --p;
        }
        sd[p * kKnnThreads + t] = d2;
        si[p * kKnnThreads + t] = (int32_t)n;
        if (cnt < k) ++cnt;
    }
    for (int p = 0; p < k; ++p) {
        nbr[q * k + p] = (p < cnt) ? si[p * kKnnThreads + t] : -1;
        d2out[q * k + p] = (p < cnt) ? sd[p * kKnnThreads + t] : 0.f;
    }
}

// kNN, one WARP per query (k <= 64): the sorted candidate list lives in registers, position p at lane p % 32,
// register p / 32.  The query's graph is scanned 32 candidates at a time (coalesced position loads); a candidate
// qualifies while the list is not full or its distance is strictly below the current k-th (ties keep the earlier =
// lower index, as in knn_kernel and oracle/graph_ops.py); qualifying candidates are inserted in index order with two
// ballots (insertion rank) and a warp shuffle (shift).  The thread-per-query kernel above spent 3.7 ms on the RNA
// batch (16 k atoms, ~2 k candidates each) shifting its shared-memory list; this one is bound by the scan.
constexpr int kKnnWarps = 4;
__global__ void __launch_bounds__(kKnnWarps * 32) knn_warp_kernel(const float* __restrict__ pos,
                                                                  const int64_t* __restrict__ batch, int64_t n_nodes,
                                                                  int k, int32_t* __restrict__ nbr,
                                                                  float* __restrict__ d2out) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * kKnnWarps + (threadIdx.x >> 5);
    if (q >= n_nodes) return;
    const int64_t g = batch[q];
    const int64_t s = lower_bound_i64(batch, n_nodes, g), e = lower_bound_i64(batch, n_nodes, g + 1);
    const float qx = pos[3 * q], qy = pos[3 * q + 1], qz = pos[3 * q + 2];
    const float inf = __int_as_float(0x7f800000);
    float ed[2] = {inf, inf};            // list entries at positions lane, 32 + lane
    int32_t ei[2] = {-1, -1};
    float tau = inf;                     // (distance, index) at position k - 1
    int32_t tau_i = 0x7fffffff;
    int cnt = 0;
    const int tl = (k - 1) & 31, tr = (k - 1) >> 5;
    // Scan order: the 32-candidate block that holds the query first, then outwards (+1, -1, +2, -2, ...).  Atoms are
    // stored along the chain, so index distance predicts spatial distance: scanned from the start of the graph every
    // candidate before the query was a new best (~1000 list insertions per RNA query, 586 us for the batch); scanned
    // outwards the k-th distance collapses in the first blocks.  The list is ordered by (distance, index) -- the same
    // total order as "ties keep the lower index" of an ascending scan -- so the result is identical.
    const int64_t nb = (e - s + 31) / 32, qb = (q - s) / 32;
    const int64_t reach = (qb > nb - 1 - qb) ? qb : nb - 1 - qb;
    for (int64_t stp = 0; stp <= 2 * reach; ++stp) {
        const int64_t off = (stp + 1) / 2, blk = (stp & 1) ? qb + off : qb - off;
        if (blk < 0 || blk >= nb) continue;
        const int64_t base = s + blk * 32;
        const int64_t n = base + lane;
        float d2 = inf;
        if (n < e) d2 = canon_d2(qx, qy, qz, pos[3 * n], pos[3 * n + 1], pos[3 * n + 2]);
        unsigned m = __ballot_sync(0xffffffffu, n < e && (d2 < tau || (d2 == tau && (int32_t)n < tau_i)));
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const float kd = __shfl_sync(0xffffffffu, d2, b);
            const int32_t ki = (int32_t)(base + b);
            if (!(kd < tau || (kd == tau && ki < tau_i))) continue;      // the k-th entry may have moved since the ballot
            // insertion rank = number of entries before (kd, ki) in (distance, index) order (empty slots: inf)
            const int pos_ins = __popc(__ballot_sync(0xffffffffu, ed[0] < kd || (ed[0] == kd && ei[0] < ki))) +
                                __popc(__ballot_sync(0xffffffffu, ed[1] < kd || (ed[1] == kd && ei[1] < ki)));
            // shift positions >= pos_ins up by one and drop the key in
            const float up0 = __shfl_up_sync(0xffffffffu, ed[0], 1), up1 = __shfl_up_sync(0xffffffffu, ed[1], 1);
            const int32_t ui0 = __shfl_up_sync(0xffffffffu, ei[0], 1), ui1 = __shfl_up_sync(0xffffffffu, ei[1], 1);
            const float last0 = __shfl_sync(0xffffffffu, ed[0], 31);
            const int32_t lasti0 = __shfl_sync(0xffffffffu, ei[0], 31);
            const int p0 = lane, p1 = 32 + lane;
            if (p1 > pos_ins) { ed[1] = lane ? up1 : last0; ei[1] = lane ? ui1 : lasti0; }
            else if (p1 == pos_ins) { ed[1] = kd; ei[1] = ki; }
            if (p0 > pos_ins) { ed[0] = up0; ei[0] = ui0; }
            else if (p0 == pos_ins) { ed[0] = kd; ei[0] = ki; }
            if (cnt < k) ++cnt;
            if (cnt == k) {
                tau = __shfl_sync(0xffffffffu, tr ? ed[1] : ed[0], tl);
                tau_i = __shfl_sync(0xffffffffu, tr ? ei[1] : ei[0], tl);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int p = r * 32 + lane;
        if (p < k) {
            nbr[q * k + p] = (p < cnt) ? ei[r] : -1;
            d2out[q * k + p] = (p < cnt) ? ed[r] : 0.f;
        }
    }
}

int knn(const float* pos, const int64_t* batch, int64_t n_nodes, int k, int32_t* nbr, float* d2, cudaStream_t st) {
    if (n_nodes == 0) return 0;
    if (k <= 64) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(knn_warp_kernel, dim3(ceil_div(n_nodes, kKnnWarps)), dim3(kKnnWarps * 32), 0, st, pos, batch, n_nodes, k, nbr, d2);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
        return 0;
    }
    size_t smem = (size_t)k * kKnnThreads * 8;
    PAMNET_CHECK_ARG(smem <= 200 * 1024, "knn: k=%d too large", k);
    PAMNET_CUDA(cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin(KC_GRAPH, 0.0, st);
    launch_pdl(knn_kernel, dim3(ceil_div(n_nodes, kKnnThreads)), dim3(kKnnThreads), smem, st, pos, batch, n_nodes, k, nbr, d2);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

template <bool FILL>
__global__ void knn_edges_kernel(const int32_t* __restrict__ nbr, const float* __restrict__ pos, int64_t n_nodes,
                                 int k, float cutoff, int32_t* __restrict__ deg, const int32_t* __restrict__ ptr,
                                 int64_t total, int64_t* __restrict__ edge_index) {
    pdl_wait();
    pdl_trigger();
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_nodes) return;
    int kept = 0;
    int64_t w = FILL ? ptr[q] : 0;
    for (int p = 0; p < k; ++p) {
        const int n = nbr[q * k + p];
        if (n < 0 || n == q) continue;
        if (edge_len(pos, n, q) <= cutoff) {
            if (FILL) {
                edge_index[w] = q;
                edge_index[total + w] = n;
                ++w;
            }
            ++kept;
        }
    }
    if (!FILL) deg[q] = kept;
}

int knn_edges_count(const int32_t* nbr, const float* pos, int64_t n_nodes, int k, float cutoff, int32_t* deg,
                    int32_t* ptr, int64_t* total_dev, cudaStream_t st) {
    if (n_nodes > 0) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(knn_edges_kernel<false>, dim3(ceil_div(n_nodes, 128)), dim3(128), 0, st, nbr, pos, n_nodes, k, cutoff, deg, nullptr, 0,
                                                                        nullptr);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    return scan_exclusive(deg, ptr, n_nodes, total_dev, st);
}

int knn_edges_fill(const int32_t* nbr, const float* pos, int64_t n_nodes, int k, float cutoff, const int32_t* ptr,
                   int64_t total, int64_t* edge_index, cudaStream_t st) {
    if (n_nodes == 0 || total == 0) return 0;
    prof_begin(KC_GRAPH, 0.0, st);
    launch_pdl(knn_edges_kernel<true>, dim3(ceil_div(n_nodes, 128)), dim3(128), 0, st, nbr, pos, n_nodes, k, cutoff, nullptr, ptr, total,
                                                                   edge_index);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// order-preserving edge compaction
// ---------------------------------------------------------------------------------------------
__global__ void edge_keep_kernel(const int64_t* __restrict__ ei, int64_t n_edges, const float* __restrict__ pos,
                                 float cutoff, int32_t* __restrict__ keep) {
    pdl_wait();
    pdl_trigger();
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int64_t a = ei[e], b = ei[n_edges + e];
    int k = (a != b);
    if (k && pos) k = edge_len(pos, b, a) <= cutoff;
    keep[e] = k;
}

__global__ void edge_compact_kernel(const int64_t* __restrict__ ei, int64_t n_edges, const int32_t* __restrict__ keep,
                                    const int32_t* __restrict__ ptr, int64_t total, int64_t* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges || !keep[e]) return;
    out[ptr[e]] = ei[e];
    out[total + ptr[e]] = ei[n_edges + e];
}

int edge_filter_count(const int64_t* ei, int64_t n_edges, const float* pos, float cutoff, int32_t* keep, int32_t* ptr,
                      int64_t* total_dev, cudaStream_t st) {
    if (n_edges > 0) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(edge_keep_kernel, dim3(ceil_div(n_edges, 256)), dim3(256), 0, st, ei, n_edges, pos, cutoff, keep);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    return scan_exclusive(keep, ptr, n_edges, total_dev, st);
}

int edge_filter_fill(const int64_t* ei, int64_t n_edges, const int32_t* keep, const int32_t* ptr, int64_t total,
                     int64_t* out, cudaStream_t st) {
    if (n_edges == 0 || total == 0) return 0;
    prof_begin(KC_GRAPH, 0.0, st);
    launch_pdl(edge_compact_kernel, dim3(ceil_div(n_edges, 256)), dim3(256), 0, st, ei, n_edges, keep, ptr, total, out);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// bucketed CSR: histogram -> scan -> unordered fill -> per-bucket insertion sort (buckets are node degrees)
// ---------------------------------------------------------------------------------------------
__global__ void split_edges_kernel(const int64_t* __restrict__ ei, int64_t n_edges, int dst_row,
                                   int32_t* __restrict__ dst, int32_t* __restrict__ src) {
    pdl_wait();
    pdl_trigger();
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    dst[e] = (int32_t)ei[(int64_t)dst_row * n_edges + e];
    src[e] = (int32_t)ei[(int64_t)(1 - dst_row) * n_edges + e];
}

__global__ void hist_kernel(const int32_t* __restrict__ keys, int64_t n, int32_t* __restrict__ cnt) {
    pdl_wait();
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&cnt[keys[i]], 1);
}

__global__ void bucket_fill_kernel(const int32_t* __restrict__ keys, int64_t n, const int32_t* __restrict__ ptr,
                                   int32_t* __restrict__ cursor, int32_t* __restrict__ items) {
    pdl_wait();
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = keys[i];
    items[ptr[k] + atomicAdd(&cursor[k], 1)] = (int32_t)i;
}

// order each bucket by (key2[item], item); key2 may be null (order by item only).  One warp per bucket: every
// lane ranks its items against the whole bucket (buckets are node degrees, ~20 for QM9, 49 for RNA kNN).
__global__ void __launch_bounds__(128) bucket_sort_kernel(const int32_t* __restrict__ ptr, int64_t n_buckets,
                                                          const int32_t* __restrict__ items_in,
                                                          int32_t* __restrict__ items_out,
                                                          const int32_t* __restrict__ key2) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= n_buckets) return;
    const int s = ptr[b], e = ptr[b + 1];
    for (int i = s + lane; i < e; i += 32) {
        const int it = items_in[i];
        const int k2 = key2 ? key2[it] : 0;
        int rank = 0;
        for (int j = s; j < e; ++j) {
            const int jt = items_in[j];
            const int j2 = key2 ? key2[jt] : 0;
            rank += (j2 < k2) || (j2 == k2 && jt < it);
        }
        items_out[s + rank] = it;
    }
}

// ptr[n_buckets+1], items[n]: items of bucket b sorted by (key2, item)
int build_buckets(const int32_t* keys, int64_t n, int64_t n_buckets, const int32_t* key2, int32_t* cnt_scratch,
                  int32_t* ptr, int32_t* items, int32_t* items_tmp, cudaStream_t st) {
    PAMNET_CUDA(cudaMemsetAsync(cnt_scratch, 0, sizeof(int32_t) * (n_buckets + 1), st));
    if (n > 0) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(hist_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, st, keys, n, cnt_scratch);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    PAMNET_TRY(scan_exclusive(cnt_scratch, ptr, n_buckets, nullptr, st));
    if (n > 0) {
        PAMNET_CUDA(cudaMemsetAsync(cnt_scratch, 0, sizeof(int32_t) * (n_buckets + 1), st));
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(bucket_fill_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, st, keys, n, ptr, cnt_scratch, items_tmp);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(bucket_sort_kernel, dim3(ceil_div(n_buckets, 4)), dim3(128), 0, st, ptr, n_buckets, items_tmp, items, key2);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    return 0;
}

__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int64_t n,
                                  int32_t* __restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

__global__ void invert_perm_kernel(const int32_t* __restrict__ perm, int64_t n, int32_t* __restrict__ inv) {
    pdl_wait();
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[perm[i]] = (int32_t)i;
}

// incoming CSR keyed by destination: ptr, eid (API edge id per slot), src/dst per slot.
// Slot order inside a destination = (source id, edge id): SparseTensor's sort at models.py:72.
int build_in_csr(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, int dst_row, int32_t* dst_api,
                 int32_t* src_api, int32_t* cnt_scratch, int32_t* tmp, int32_t* ptr, int32_t* eid, int32_t* src_csr,
                 int32_t* dst_csr, cudaStream_t st) {
    if (n_edges > 0) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(split_edges_kernel, dim3(ceil_div(n_edges, 256)), dim3(256), 0, st, edge_index, n_edges, dst_row, dst_api, src_api);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    PAMNET_TRY(build_buckets(dst_api, n_edges, n_nodes, src_api, cnt_scratch, ptr, eid, tmp, st));
    if (n_edges > 0) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(gather_i32_kernel, dim3(ceil_div(n_edges, 256)), dim3(256), 0, st, src_api, eid, n_edges, src_csr);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
        if (dst_csr) {
            prof_begin(KC_GRAPH, 0.0, st);
            launch_pdl(gather_i32_kernel, dim3(ceil_div(n_edges, 256)), dim3(256), 0, st, dst_api, eid, n_edges, dst_csr);
            prof_end(st);
            PAMNET_LAUNCH_CHECK();
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// triplets in the reference's (API) order -- PAMNet.indices, models.py:68-98
// ---------------------------------------------------------------------------------------------
// per API edge e = (j -> i):  two-hop count = |{p in in(j): src[p] != i}|, one-hop = |{p in in(i): src[p] != i}|
__global__ void triplet_count_kernel(const int32_t* __restrict__ dst_api, const int32_t* __restrict__ src_api,
                                     int64_t n_edges, const int32_t* __restrict__ ptr,
                                     const int32_t* __restrict__ src_csr, int32_t* __restrict__ c2,
                                     int32_t* __restrict__ c1) {
    pdl_wait();
    pdl_trigger();
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int j = src_api[e], i = dst_api[e];
    int n2 = 0, n1 = 0;
    for (int p = ptr[j]; p < ptr[j + 1]; ++p) n2 += (src_csr[p] != i);
    for (int p = ptr[i]; p < ptr[i + 1]; ++p) n1 += (src_csr[p] != i);
    c2[e] = n2;
    c1[e] = n1;
}

__global__ void triplet_fill_kernel(const int32_t* __restrict__ dst_api, const int32_t* __restrict__ src_api,
                                    int64_t n_edges, const int32_t* __restrict__ ptr,
                                    const int32_t* __restrict__ src_csr, const int32_t* __restrict__ eid,
                                    const int32_t* __restrict__ off2, const int32_t* __restrict__ off1,
                                    int64_t* __restrict__ idx_i, int64_t* __restrict__ idx_j,
                                    int64_t* __restrict__ idx_k, int64_t* __restrict__ idx_kj,
                                    int64_t* __restrict__ idx_ji, int64_t* __restrict__ idx_i_pair,
                                    int64_t* __restrict__ idx_j1_pair, int64_t* __restrict__ idx_j2_pair,
                                    int64_t* __restrict__ idx_jj_pair, int64_t* __restrict__ idx_ji_pair) {
    pdl_wait();
    pdl_trigger();
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int j = src_api[e], i = dst_api[e];
    int w = off2[e];
    for (int p = ptr[j]; p < ptr[j + 1]; ++p) {
        const int k = src_csr[p];
        if (k == i) continue;
        idx_i[w] = i; idx_j[w] = j; idx_k[w] = k; idx_kj[w] = eid[p]; idx_ji[w] = e;
        ++w;
    }
    w = off1[e];
    for (int p = ptr[i]; p < ptr[i + 1]; ++p) {
        const int j2 = src_csr[p];
        if (j2 == i) continue;
        idx_i_pair[w] = j; idx_j1_pair[w] = i; idx_j2_pair[w] = j2; idx_jj_pair[w] = eid[p]; idx_ji_pair[w] = e;
        ++w;
    }
}

struct TripletScratch {
    int32_t *dst_api, *src_api, *cnt, *ptr, *eid, *src_csr, *c2, *c1, *off2, *off1, *tmp;
};

static size_t triplet_scratch_layout(int64_t n_nodes, int64_t n_edges, void* base, TripletScratch* s) {
    size_t off = 0;
    auto take = [&](int64_t n) {
        int32_t* p = base ? reinterpret_cast<int32_t*>(static_cast<char*>(base) + off) : nullptr;
        off += align_up(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        return p;
    };
    TripletScratch t;
    t.dst_api = take(n_edges); t.src_api = take(n_edges); t.cnt = take(n_nodes + 1); t.ptr = take(n_nodes + 1);
    t.eid = take(n_edges); t.src_csr = take(n_edges); t.c2 = take(n_edges); t.c1 = take(n_edges);
    t.off2 = take(n_edges + 1); t.off1 = take(n_edges + 1); t.tmp = take(n_edges);
    if (s) *s = t;
    return off;
}

size_t triplet_scratch_bytes(int64_t n_nodes, int64_t n_edges) {
    return triplet_scratch_layout(n_nodes, n_edges, nullptr, nullptr);
}

int triplet_count(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, void* scratch, int64_t* counts_dev,
                  cudaStream_t st) {
    TripletScratch s;
    triplet_scratch_layout(n_nodes, n_edges, scratch, &s);
    PAMNET_TRY(build_in_csr(edge_index, n_edges, n_nodes, 1, s.dst_api, s.src_api, s.cnt, s.tmp, s.ptr, s.eid,
                            s.src_csr, nullptr, st));
    if (n_edges > 0) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(triplet_count_kernel, dim3(ceil_div(n_edges, 128)), dim3(128), 0, st, s.dst_api, s.src_api, n_edges, s.ptr, s.src_csr,
                                                                     s.c2, s.c1);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    PAMNET_TRY(scan_exclusive(s.c2, s.off2, n_edges, counts_dev, st));
    PAMNET_TRY(scan_exclusive(s.c1, s.off1, n_edges, counts_dev + 1, st));
    return 0;
}

int triplet_fill(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, const void* scratch, int64_t* const* out,
                 cudaStream_t st) {
    (void)edge_index;
    TripletScratch s;
    triplet_scratch_layout(n_nodes, n_edges, const_cast<void*>(scratch), &s);
    if (n_edges == 0) return 0;
    prof_begin(KC_GRAPH, 0.0, st);
    launch_pdl(triplet_fill_kernel, dim3(ceil_div(n_edges, 128)), dim3(128), 0, st, s.dst_api, s.src_api, n_edges, s.ptr, s.src_csr, s.eid,
                                                                s.off2, s.off1, out[0], out[1], out[2], out[3], out[4],
                                                                out[5], out[6], out[7], out[8], out[9]);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// execution plan
// ---------------------------------------------------------------------------------------------
// Two blobs: `base` holds everything whose size is known before the triplet count (so plan_count can run),
// `trip` holds the T-sized arrays and is allocated by the caller once {T2, T1} are known.
void plan_layout(const pamnet_sizes_t& sz, void* base, void* trip, Plan* out, size_t* base_bytes,
                 size_t* trip_bytes) {
    size_t off = 0;
    char* cur = static_cast<char*>(base);
    auto take_i = [&](int64_t n) {
        int32_t* p = cur ? reinterpret_cast<int32_t*>(cur + off) : nullptr;
        off += align_up(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        return p;
    };
    auto take_f = [&](int64_t n) { return reinterpret_cast<float*>(take_i(n)); };
    const int64_t N = sz.n_nodes, G = sz.n_graphs, Eg = sz.n_edges_g, El = sz.n_edges_l, T = sz.n_t2 + sz.n_t1;
    const int64_t Emax = Eg > El ? Eg : El;
    Plan p;
    p.n2g = take_i(N); p.gptr = take_i(G + 1);
    p.g_ptr = take_i(N + 1); p.g_src = take_i(Eg); p.g_dst = take_i(Eg); p.g_eid = take_i(Eg); p.g_optr = take_i(N + 1); p.g_opos = take_i(Eg);
    p.l_ptr = take_i(N + 1); p.l_src = take_i(El); p.l_dst = take_i(El); p.l_eid = take_i(El);
    p.l_optr = take_i(N + 1); p.l_opos = take_i(El);
    p.t_split = take_i(El); p.t_cnt = take_i(El); p.t_ptr = take_i(El + 1); p.tt_ptr = take_i(El + 1);
    p.dist_g = take_f(Eg); p.dist_l = take_f(El);
    p.tmp_a = take_i(Emax); p.tmp_b = take_i(Emax); p.tmp_c = take_i(Emax);
    p.tmp_d = take_i(Emax); p.tmp_e = take_i(Emax); p.tmp_f = take_i(Emax);
    p.cnt4 = take_i(4 * (N + 1)); p.tmp4 = take_i(2 * Eg + 2 * El);
    p.cnt = take_i((N > El ? N : El) + 2);
    if (base_bytes) *base_bytes = off;
    off = 0;
    cur = static_cast<char*>(trip);
    p.t_gather = take_i(T); p.t_owner = take_i(T); p.tt_t = take_i(T); p.t_tmp = take_i(T); p.t_angle = take_f(T);
    if (trip_bytes) *trip_bytes = off;
    if (out) *out = p;
}

__global__ void n2g_kernel(const int64_t* __restrict__ batch, int64_t n, int64_t n_graphs, int32_t* __restrict__ n2g,
                           int32_t* __restrict__ gptr) {
    pdl_wait();
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) n2g[i] = (int32_t)batch[i];
    if (i <= n_graphs) gptr[i] = (int32_t)lower_bound_i64(batch, n, i);
}

// per local CSR slot k (edge j -> i): two-hop and one-hop counts (same rule as triplet_count_kernel)
__global__ void plan_tcount_kernel(const int32_t* __restrict__ l_ptr, const int32_t* __restrict__ l_src,
                                   const int32_t* __restrict__ l_dst, int64_t n_edges, int two_hop,
                                   int32_t* __restrict__ t_split, int32_t* __restrict__ t_cnt,
                                   unsigned long long* __restrict__ totals) {
    pdl_wait();
    pdl_trigger();
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int n2 = 0, n1 = 0;
    if (k < n_edges) {
        const int j = l_src[k], i = l_dst[k];
        if (two_hop)
            for (int p = l_ptr[j]; p < l_ptr[j + 1]; ++p) n2 += (l_src[p] != i);
        for (int p = l_ptr[i]; p < l_ptr[i + 1]; ++p) n1 += (l_src[p] != i);
        t_split[k] = n2;
        t_cnt[k] = n2 + n1;
    }
    // block totals -> {T2, T1}
    __shared__ int s2[32], s1[32];
    int w2 = n2, w1 = n1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        w2 += __shfl_xor_sync(0xffffffffu, w2, o);
        w1 += __shfl_xor_sync(0xffffffffu, w1, o);
    }
    if ((threadIdx.x & 31) == 0) { s2[threadIdx.x >> 5] = w2; s1[threadIdx.x >> 5] = w1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = 0, b = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) { a += s2[w]; b += s1[w]; }
        atomicAdd(&totals[0], (unsigned long long)a);
        atomicAdd(&totals[1], (unsigned long long)b);
    }
}

// segment of slot k: first its two-hop entries (edges into the source j, minus the back edge), then its
// one-hop entries (edges into the target i) -- the order the reference's cat() gives (local_message_passing.py:38-40)
__global__ void plan_tfill_kernel(const int32_t* __restrict__ l_ptr, const int32_t* __restrict__ l_src,
                                  const int32_t* __restrict__ l_dst, int64_t n_edges, int two_hop,
                                  const int32_t* __restrict__ t_ptr, const float* __restrict__ pos,
                                  int32_t* __restrict__ t_gather, int32_t* __restrict__ t_owner,
                                  float* __restrict__ t_angle) {
    pdl_wait();
    pdl_trigger();
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_edges) return;
    const int j = l_src[k], i = l_dst[k];
    int w = t_ptr[k];
    if (two_hop) {
        for (int p = l_ptr[j]; p < l_ptr[j + 1]; ++p) {
            const int kk = l_src[p];
            if (kk == i) continue;
            t_gather[w] = p; t_owner[w] = (int32_t)k;
            t_angle[w] = bond_angle(pos, i, j, kk);       // (idx_i, idx_j, idx_k)
            ++w;
        }
    }
    for (int p = l_ptr[i]; p < l_ptr[i + 1]; ++p) {
        const int j2 = l_src[p];
        if (j2 == i) continue;
        t_gather[w] = p; t_owner[w] = (int32_t)k;
        t_angle[w] = bond_angle(pos, j, i, j2);           // (idx_i_pair, idx_j1_pair, idx_j2_pair)
        ++w;
    }
}

__global__ void csr_dist_kernel(const int32_t* __restrict__ ptr_g, const int32_t* __restrict__ src_g,
                                const int32_t* __restrict__ ptr_l, const int32_t* __restrict__ src_l, int64_t n_nodes,
                                const float* __restrict__ pos, float* __restrict__ dist_g, float* __restrict__ dist_l) {
    pdl_wait();
    pdl_trigger();
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_nodes) return;
    const int32_t* ptr = blockIdx.y ? ptr_l : ptr_g;
    const int32_t* src = blockIdx.y ? src_l : src_g;
    float* dist = blockIdx.y ? dist_l : dist_g;
    for (int p = ptr[n]; p < ptr[n + 1]; ++p) dist[p] = edge_len(pos, n, src[p]);
}

// ---- four CSRs (global in / out, local in / out) built by the SAME six launches -------------------------------
// job j: bucket key per API edge, bucket order key (key2, item): in-CSR = (source id, edge id) which is
// SparseTensor's order (models.py:72); out-CSR = edge id.
struct CsrJobs {
    const int32_t* keys[4];
    const int32_t* key2[4];
    int64_t n[4];
    int32_t *cnt[4], *ptr[4], *tmp[4], *items[4];
    int64_t n_buckets;
};

__global__ void split_edges2_kernel(const int64_t* __restrict__ eg, int64_t n_g, int g_dst_row,
                                    const int64_t* __restrict__ el, int64_t n_l, int32_t* __restrict__ g_dst,
                                    int32_t* __restrict__ g_src, int32_t* __restrict__ l_dst, int32_t* __restrict__ l_src,
                                    const int64_t* __restrict__ batch, int64_t n_nodes, int64_t n_graphs,
                                    int32_t* __restrict__ n2g, int32_t* __restrict__ gptr) {
    pdl_wait();
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_g) {
        g_dst[i] = (int32_t)eg[(int64_t)g_dst_row * n_g + i];
        g_src[i] = (int32_t)eg[(int64_t)(1 - g_dst_row) * n_g + i];
    }
    if (i < n_l) {            // local graph: i = edge_index[1] always (local_message_passing.py:37)
        l_dst[i] = (int32_t)el[n_l + i];
        l_src[i] = (int32_t)el[i];
    }
    if (i < n_nodes) n2g[i] = (int32_t)batch[i];
    if (i <= n_graphs) gptr[i] = (int32_t)lower_bound_i64(batch, n_nodes, i);
}

__global__ void hist4_kernel(const CsrJobs j) {
    pdl_wait();
    pdl_trigger();
    const int q = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < j.n[q]) atomicAdd(&j.cnt[q][j.keys[q][i]], 1);
}

// block q scans job q (same algorithm as scan_exclusive_kernel)
__global__ void __launch_bounds__(kScanThreads) scan4_kernel(const CsrJobs j) {
    pdl_wait();
    pdl_trigger();
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const int q = blockIdx.x;
    const int32_t* in = j.cnt[q];
    int32_t* out = j.ptr[q];
    const int64_t n = j.n_buckets;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += (int64_t)kScanThreads * kScanItems) {
        int32_t v[kScanItems];
        int32_t local = 0;
        const int64_t i0 = base + (int64_t)tid * kScanItems;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            v[k] = (i0 + k < n) ? in[i0 + k] : 0;
            local += v[k];
        }
        int32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int32_t w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int32_t carry = carry_s;
        int32_t excl = carry + incl - local + (wid > 0 ? warp_tot[wid - 1] : 0);
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (i0 + k < n) out[i0 + k] = excl;
            excl += v[k];
        }
        __syncthreads();
        if (tid == kScanThreads - 1) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (tid == 0) out[n] = carry_s;
}

__global__ void fill4_kernel(const CsrJobs j) {
    pdl_wait();
    pdl_trigger();
    const int q = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= j.n[q]) return;
    const int k = j.keys[q][i];
    j.tmp[q][j.ptr[q][k] + atomicAdd(&j.cnt[q][k], 1)] = (int32_t)i;
}

__global__ void __launch_bounds__(128) sort4_kernel(const CsrJobs j) {
    pdl_wait();
    pdl_trigger();
    const int q = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= j.n_buckets) return;
    const int32_t* ptr = j.ptr[q];
    const int32_t* in = j.tmp[q];
    const int32_t* key2 = j.key2[q];
    const int s = ptr[b], e = ptr[b + 1];
    for (int i = s + lane; i < e; i += 32) {
        const int it = in[i];
        const int k2 = key2 ? key2[it] : 0;
        int rank = 0;
        for (int u = s; u < e; ++u) {
            const int jt = in[u];
            const int j2 = key2 ? key2[jt] : 0;
            rank += (j2 < k2) || (j2 == k2 && jt < it);
        }
        j.items[q][s + rank] = it;
    }
}

// in-CSR: per-slot source / destination and the inverse permutation edge id -> slot
__global__ void csr_finish1_kernel(const int32_t* __restrict__ g_eid, const int32_t* __restrict__ g_src_api,
                                   const int32_t* __restrict__ g_dst_api, int64_t n_g, int32_t* __restrict__ g_src,
                                   int32_t* __restrict__ g_dst, int32_t* __restrict__ g_pos_of,
                                   const int32_t* __restrict__ l_eid, const int32_t* __restrict__ l_src_api,
                                   const int32_t* __restrict__ l_dst_api, int64_t n_l, int32_t* __restrict__ l_src,
                                   int32_t* __restrict__ l_dst, int32_t* __restrict__ l_pos_of) {
    pdl_wait();
    pdl_trigger();
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_g) {
        const int e = g_eid[k];
        g_src[k] = g_src_api[e]; g_dst[k] = g_dst_api[e]; g_pos_of[e] = (int32_t)k;
    }
    if (k < n_l) {
        const int e = l_eid[k];
        l_src[k] = l_src_api[e]; l_dst[k] = l_dst_api[e]; l_pos_of[e] = (int32_t)k;
    }
}

// out-CSR entries (edge ids, ascending per source) -> slots; triplet counts per local slot
__global__ void csr_finish2_kernel(int32_t* __restrict__ g_opos, const int32_t* __restrict__ g_pos_of, int64_t n_g,
                                   int32_t* __restrict__ l_opos, const int32_t* __restrict__ l_pos_of, int64_t n_l) {
    pdl_wait();
    pdl_trigger();
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n_g) g_opos[u] = g_pos_of[g_opos[u]];
    if (u < n_l) l_opos[u] = l_pos_of[l_opos[u]];
}

int plan_count(const pamnet_config_t& cfg, const pamnet_sizes_t& sz, const int64_t* edge_index_g,
               const int64_t* edge_index_l, const int64_t* batch, void* plan_base, int64_t* counts_dev,
               cudaStream_t st) {
    Plan p;
    plan_layout(sz, plan_base, nullptr, &p, nullptr, nullptr);
    const int64_t N = sz.n_nodes, Eg = sz.n_edges_g, El = sz.n_edges_l;
    const int64_t Emax = Eg > El ? Eg : El;
    // global graph: x_i = x[edge_index[dst_row]] is also the aggregation target (PyG propagate)
    const int g_dst_row = (cfg.flow == PAMNET_TARGET_TO_SOURCE) ? 0 : 1;
    {
        int64_t n = Emax > N ? Emax : N;
        n = n > sz.n_graphs + 1 ? n : sz.n_graphs + 1;
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(split_edges2_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, st, edge_index_g, Eg, g_dst_row, edge_index_l, El, p.tmp_a,
                                                              p.tmp_b, p.tmp_d, p.tmp_e, batch, N, sz.n_graphs, p.n2g,
                                                              p.gptr);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    CsrJobs j;
    memset(&j, 0, sizeof(j));
    j.n_buckets = N;
    const int32_t* keys[4] = {p.tmp_a, p.tmp_b, p.tmp_d, p.tmp_e};          // g dst, g src, l dst, l src
    const int32_t* key2[4] = {p.tmp_b, nullptr, p.tmp_e, nullptr};
    int32_t* ptrs[4] = {p.g_ptr, p.g_optr, p.l_ptr, p.l_optr};
    int32_t* items[4] = {p.g_eid, p.g_opos, p.l_eid, p.l_opos};
    for (int q = 0; q < 4; ++q) {
        j.keys[q] = keys[q]; j.key2[q] = key2[q]; j.n[q] = q < 2 ? Eg : El;
        j.cnt[q] = p.cnt4 + q * (N + 1); j.ptr[q] = ptrs[q]; j.items[q] = items[q];
        j.tmp[q] = p.tmp4 + (q < 2 ? q * Eg : 2 * Eg + (q - 2) * El);
    }
    PAMNET_CUDA(cudaMemsetAsync(p.cnt4, 0, sizeof(int32_t) * 4 * (N + 1), st));
    if (Emax > 0) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(hist4_kernel, dim3(dim3(ceil_div(Emax, 256), 4)), dim3(256), 0, st, j);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    prof_begin(KC_GRAPH, 0.0, st);
    launch_pdl(scan4_kernel, dim3(4), dim3(kScanThreads), 0, st, j);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    PAMNET_CUDA(cudaMemsetAsync(p.cnt4, 0, sizeof(int32_t) * 4 * (N + 1), st));
    if (Emax > 0) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(fill4_kernel, dim3(dim3(ceil_div(Emax, 256), 4)), dim3(256), 0, st, j);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(sort4_kernel, dim3(dim3(ceil_div(N, 4), 4)), dim3(128), 0, st, j);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(csr_finish1_kernel, dim3(ceil_div(Emax, 256)), dim3(256), 0, st, p.g_eid, p.tmp_b, p.tmp_a, Eg, p.g_src, p.g_dst,
                                                                p.tmp_c, p.l_eid, p.tmp_e, p.tmp_d, El, p.l_src,
                                                                p.l_dst, p.tmp_f);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(csr_finish2_kernel, dim3(ceil_div(Emax, 256)), dim3(256), 0, st, p.g_opos, p.tmp_c, Eg, p.l_opos, p.tmp_f, El);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    PAMNET_CUDA(cudaMemsetAsync(counts_dev, 0, 2 * sizeof(int64_t), st));
    if (El > 0) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(plan_tcount_kernel, dim3(ceil_div(El, 128)), dim3(128), 0, st, p.l_ptr, p.l_src, p.l_dst, El, cfg.simple ? 0 : 1,
                                                              p.t_split, p.t_cnt,
                                                              reinterpret_cast<unsigned long long*>(counts_dev));
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    PAMNET_TRY(scan_exclusive(p.t_cnt, p.t_ptr, El, nullptr, st));
    return 0;
}

int plan_fill(const pamnet_config_t& cfg, const pamnet_sizes_t& sz, const float* pos, void* plan_base,
              void* plan_trip, cudaStream_t st) {
    Plan p;
    plan_layout(sz, plan_base, plan_trip, &p, nullptr, nullptr);
    const int64_t N = sz.n_nodes, El = sz.n_edges_l, T = sz.n_t2 + sz.n_t1;
    if (El > 0 && T > 0) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(plan_tfill_kernel, dim3(ceil_div(El, 128)), dim3(128), 0, st, p.l_ptr, p.l_src, p.l_dst, El, cfg.simple ? 0 : 1, p.t_ptr,
                                                             pos, p.t_gather, p.t_owner, p.t_angle);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    PAMNET_TRY(build_buckets(p.t_gather, T, El, nullptr, p.cnt, p.tt_ptr, p.tt_t, p.t_tmp, st));
    if (N > 0) {
        prof_begin(KC_GRAPH, 0.0, st);
        launch_pdl(csr_dist_kernel, dim3(dim3(ceil_div(N, 128), 2)), dim3(128), 0, st, p.g_ptr, p.g_src, p.l_ptr, p.l_src, N, pos, p.dist_g,
                                                                   p.dist_l);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace pamnet
