// extern "C" surface of libpamnet_sm100.so (declared in include/pamnet_b200.h).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#include "basis.cuh"
#include "chain.cuh"
#include "collate.cuh"
#include "comm.cuh"
#include "front_mol.cuh"
#include "gemm.cuh"
#include "graph.cuh"
#include "message.cuh"
#include "model.cuh"
#include "optim.cuh"
#include "readout.cuh"

namespace pamnet {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

int pdl_level() {
    static int lvl = -1;
    if (lvl < 0) {
        const char* e = getenv("PAMNET_PDL");
        lvl = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
    }
    return lvl;
}
bool pdl_enabled() { return pdl_level() != 0; }

// ---- launch counter + event profiler (single-threaded use: bench / tests) --------------------------
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// The attribute is per (function, device): remember which devices a kernel was configured on.  A mutex-protected table
// instead of a function-local `static bool` (not thread-safe, and wrong for a second device in the same process).
int func_smem_once(const void* func, size_t bytes) {
    static std::mutex mu;
    static std::vector<std::pair<const void*, unsigned>> done;     // (kernel, device bit mask)
    int dev = 0;
    PAMNET_CUDA(cudaGetDevice(&dev));
    const unsigned bit = 1u << (dev & 31);
    std::lock_guard<std::mutex> lk(mu);
    for (auto& e : done)
        if (e.first == func) {
            if (e.second & bit) return 0;
            PAMNET_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            e.second |= bit;
            return 0;
        }
    PAMNET_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    done.emplace_back(func, bit);
    return 0;
}

struct ProfRec { int cls; double bytes; cudaEvent_t e0, e1; cudaStream_t st; double flops; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static size_t g_prof_used = 0;
static cudaEvent_t prof_event() {
    if (g_prof_used == g_prof_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        g_prof_pool.push_back(e);
    }
    return g_prof_pool[g_prof_used++];
}
void prof_begin(int cls, double bytes, cudaStream_t st) {
    if (!g_prof_on) return;
    ProfRec r{cls, bytes, prof_event(), prof_event(), st, 0.0};
    cudaEventRecord(r.e0, st);
    g_prof_recs.push_back(r);
}
void prof_flops(double flops) {     // attach fp32-equivalent flops to the record opened by the last prof_begin
    if (g_prof_on && !g_prof_recs.empty()) g_prof_recs.back().flops = flops;
}
void prof_end(cudaStream_t st) {
    if (!g_prof_on || g_prof_recs.empty()) return;
    cudaEventRecord(g_prof_recs.back().e1, st);
}
}  // namespace pamnet

using namespace pamnet;

#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define REQUIRE(p) PAMNET_CHECK_ARG((p) != nullptr, "null pointer: %s", #p)

extern "C" {

int pamnet_abi_version(void) { return PAMNET_ABI_VERSION; }
const char* pamnet_last_error(void) { return get_error(); }

int pamnet_param_count(const pamnet_config_t* cfg) {
    ModelP mp;
    if (!cfg || build_param_layout(*cfg, &mp) != 0) return -1;
    return (int)mp.offsets.size();
}
int64_t pamnet_param_total(const pamnet_config_t* cfg) {
    ModelP mp;
    if (!cfg || build_param_layout(*cfg, &mp) != 0) return -1;
    return mp.total;
}
int pamnet_param_offsets(const pamnet_config_t* cfg, int64_t* offsets, int64_t* numel) {
    REQUIRE(cfg); REQUIRE(offsets); REQUIRE(numel);
    ModelP mp;
    PAMNET_TRY(build_param_layout(*cfg, &mp));
    for (size_t i = 0; i < mp.offsets.size(); ++i) { offsets[i] = mp.offsets[i]; numel[i] = mp.numel[i]; }
    return 0;
}

int pamnet_radius_count(const float* pos, const int64_t* batch, int64_t n_nodes, float r, int32_t max_nb,
                        int32_t drop_self, int32_t* deg, int32_t* ptr, int64_t* total_dev, void* stream) {
    REQUIRE(pos); REQUIRE(batch); REQUIRE(deg); REQUIRE(ptr); REQUIRE(total_dev);
    PAMNET_CHECK_ARG(n_nodes >= 0 && max_nb > 0, "radius: n_nodes=%lld max=%d", (long long)n_nodes, max_nb);
    return radius_count(pos, batch, n_nodes, r, max_nb, drop_self, deg, ptr, total_dev, ST(stream));
}
int pamnet_radius_fill(const float* pos, const int64_t* batch, int64_t n_nodes, float r, int32_t max_nb,
                       int32_t drop_self, const int32_t* ptr, int64_t total, int64_t* edge_index, void* stream) {
    REQUIRE(pos); REQUIRE(batch); REQUIRE(ptr);
    PAMNET_CHECK_ARG(total == 0 || edge_index, "radius_fill: null edge_index");
    return radius_fill(pos, batch, n_nodes, r, max_nb, drop_self, ptr, total, edge_index, ST(stream));
}
size_t pamnet_radius_grid_scratch_bytes(int64_t n_nodes, int64_t n_graphs) {
    return radius_grid_scratch_bytes(n_nodes, n_graphs);
}
int pamnet_radius_grid_count(const float* pos, const int64_t* batch, int64_t n_nodes, int64_t n_graphs, float r,
                             int32_t max_nb, int32_t drop_self, void* scratch, size_t scratch_bytes, int32_t* deg,
                             int32_t* ptr, int64_t* total_dev, void* stream) {
    REQUIRE(pos); REQUIRE(batch); REQUIRE(deg); REQUIRE(ptr); REQUIRE(total_dev); REQUIRE(scratch);
    PAMNET_CHECK_ARG(n_nodes >= 0 && n_graphs > 0 && max_nb > 0, "radius_grid: n_nodes=%lld n_graphs=%lld max=%d",
                     (long long)n_nodes, (long long)n_graphs, max_nb);
    return radius_grid_count(pos, batch, n_nodes, n_graphs, r, max_nb, drop_self, scratch, scratch_bytes, deg, ptr,
                             total_dev, ST(stream));
}
int pamnet_radius_grid_fill(const float* pos, const int64_t* batch, int64_t n_nodes, int64_t n_graphs, float r,
                            int32_t max_nb, int32_t drop_self, const void* scratch, const int32_t* ptr, int64_t total,
                            int64_t* edge_index, void* stream) {
    REQUIRE(pos); REQUIRE(batch); REQUIRE(ptr); REQUIRE(scratch);
    PAMNET_CHECK_ARG(total == 0 || edge_index, "radius_grid_fill: null edge_index");
    return radius_grid_fill(pos, batch, n_nodes, n_graphs, r, max_nb, drop_self, scratch, ptr, total, edge_index, ST(stream));
}
int pamnet_knn(const float* pos, const int64_t* batch, int64_t n_nodes, int32_t k, int32_t* nbr, float* d2,
               void* stream) {
    REQUIRE(pos); REQUIRE(batch); REQUIRE(nbr); REQUIRE(d2);
    PAMNET_CHECK_ARG(k > 0, "knn: k=%d", k);
    return knn(pos, batch, n_nodes, k, nbr, d2, ST(stream));
}
int pamnet_knn_edges_count(const int32_t* nbr, const float* pos, int64_t n_nodes, int32_t k, float cutoff,
                           int32_t* deg, int32_t* ptr, int64_t* total_dev, void* stream) {
    REQUIRE(nbr); REQUIRE(pos); REQUIRE(deg); REQUIRE(ptr); REQUIRE(total_dev);
    return knn_edges_count(nbr, pos, n_nodes, k, cutoff, deg, ptr, total_dev, ST(stream));
}
int pamnet_knn_edges_fill(const int32_t* nbr, const float* pos, int64_t n_nodes, int32_t k, float cutoff,
                          const int32_t* ptr, int64_t total, int64_t* edge_index, void* stream) {
    REQUIRE(nbr); REQUIRE(pos); REQUIRE(ptr);
    return knn_edges_fill(nbr, pos, n_nodes, k, cutoff, ptr, total, edge_index, ST(stream));
}
int pamnet_edge_filter_count(const int64_t* edge_index, int64_t n_edges, const float* pos, float cutoff,
                             int32_t* keep, int32_t* ptr, int64_t* total_dev, void* stream) {
    REQUIRE(keep); REQUIRE(ptr); REQUIRE(total_dev);
    PAMNET_CHECK_ARG(n_edges == 0 || edge_index, "edge_filter: null edge_index");
    return edge_filter_count(edge_index, n_edges, pos, cutoff, keep, ptr, total_dev, ST(stream));
}
int pamnet_edge_filter_fill(const int64_t* edge_index, int64_t n_edges, const int32_t* keep, const int32_t* ptr,
                            int64_t total, int64_t* edge_index_out, void* stream) {
    REQUIRE(keep); REQUIRE(ptr);
    return edge_filter_fill(edge_index, n_edges, keep, ptr, total, edge_index_out, ST(stream));
}

size_t pamnet_triplet_scratch_bytes(int64_t n_nodes, int64_t n_edges) { return triplet_scratch_bytes(n_nodes, n_edges); }
int pamnet_triplet_count(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, void* scratch,
                         int64_t* counts_dev, void* stream) {
    REQUIRE(scratch); REQUIRE(counts_dev);
    return triplet_count(edge_index, n_edges, n_nodes, scratch, counts_dev, ST(stream));
}
int pamnet_triplet_fill(const int64_t* edge_index, int64_t n_edges, int64_t n_nodes, const void* scratch,
                        int64_t t2, int64_t t1, int64_t* idx_i, int64_t* idx_j, int64_t* idx_k, int64_t* idx_kj,
                        int64_t* idx_ji, int64_t* idx_i_pair, int64_t* idx_j1_pair, int64_t* idx_j2_pair,
                        int64_t* idx_jj_pair, int64_t* idx_ji_pair, void* stream) {
    REQUIRE(scratch);
    (void)t2; (void)t1;
    int64_t* out[10] = {idx_i, idx_j, idx_k, idx_kj, idx_ji, idx_i_pair, idx_j1_pair, idx_j2_pair, idx_jj_pair, idx_ji_pair};
    return triplet_fill(edge_index, n_edges, n_nodes, scratch, out, ST(stream));
}

int pamnet_plan_bytes(const pamnet_config_t* cfg, const pamnet_sizes_t* sz, size_t* base_bytes, size_t* trip_bytes) {
    REQUIRE(cfg); REQUIRE(sz);
    plan_layout(*sz, nullptr, nullptr, nullptr, base_bytes, trip_bytes);
    return 0;
}
int pamnet_plan_count(const pamnet_config_t* cfg, const pamnet_sizes_t* sz, const int64_t* edge_index_g,
                      const int64_t* edge_index_l, const int64_t* batch, void* plan_base, int64_t* counts_dev,
                      void* stream) {
    REQUIRE(cfg); REQUIRE(sz); REQUIRE(batch); REQUIRE(plan_base); REQUIRE(counts_dev);
    return plan_count(*cfg, *sz, edge_index_g, edge_index_l, batch, plan_base, counts_dev, ST(stream));
}
int pamnet_plan_fill(const pamnet_config_t* cfg, const pamnet_sizes_t* sz, const float* pos, void* plan_base,
                     void* plan_trip, void* stream) {
    REQUIRE(cfg); REQUIRE(sz); REQUIRE(pos); REQUIRE(plan_base); REQUIRE(plan_trip);
    return plan_fill(*cfg, *sz, pos, plan_base, plan_trip, ST(stream));
}

// ---- whole front end in one call ---------------------------------------------------------------------------------
namespace {
struct BuildScratch {
    int64_t* counts;
    int32_t *deg_a, *ptr_a, *deg_b, *ptr_b, *keep, *pk, *nbr, *mol_rng;
    float* d2;
    void* grid;             // cell lists of the large-graph radius search (graph_grid.cu)
    size_t grid_bytes;
};
constexpr int kKnnK = 50;            // models.py:143
size_t build_scratch_layout(int kind, int64_t n, int64_t n_edges_in, int64_t cap_eg, void* base, BuildScratch* out) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off += align_up(bytes ? bytes : 1);
        return p;
    };
    BuildScratch b;
    const int64_t ne = kind == PAMNET_PDBBIND ? cap_eg : n_edges_in;
    b.counts = static_cast<int64_t*>(take(8 * sizeof(int64_t)));
    b.deg_a = static_cast<int32_t*>(take(sizeof(int32_t) * n));
    b.ptr_a = static_cast<int32_t*>(take(sizeof(int32_t) * (n + 1)));
    b.deg_b = static_cast<int32_t*>(take(sizeof(int32_t) * n));
    b.ptr_b = static_cast<int32_t*>(take(sizeof(int32_t) * (n + 1)));
    b.keep = static_cast<int32_t*>(take(sizeof(int32_t) * ne));
    b.pk = static_cast<int32_t*>(take(sizeof(int32_t) * (ne + 1)));
    b.mol_rng = static_cast<int32_t*>(take(kind == PAMNET_QM9 ? sizeof(int32_t) * 2 * (n + 2) : 0));   // front_mol.cuh range tables
    b.grid = take(kind != PAMNET_RNA ? radius_grid_scratch_bytes(n, n) : 0);    // sized for the worst case n_graphs = n
    b.grid_bytes = kind != PAMNET_RNA ? radius_grid_scratch_bytes(n, n) : 0;
    b.nbr = static_cast<int32_t*>(take(kind == PAMNET_RNA ? sizeof(int32_t) * n * kKnnK : 0));
    b.d2 = static_cast<float*>(take(kind == PAMNET_RNA ? sizeof(float) * n * kKnnK : 0));
    if (out) *out = b;
    return off;
}
// PAMNET_FRONT=mol: per-molecule front end for QM9-shaped batches (read per call so that tests can switch it)
bool front_mol_enabled() {
    // default since round 2 (bit-identical to the generic kernels on a B200, tests/test_gpu_front_mol.py);
    // PAMNET_FRONT=generic selects the generic graph kernels
    const char* e = getenv("PAMNET_FRONT");
    return !(e && strcmp(e, "generic") == 0);
}
thread_local int64_t* g_host_counts = nullptr;      // pinned: the counters are read back between the build phases
int read_counts(const int64_t* dev, cudaStream_t st) {
    if (!g_host_counts) PAMNET_CUDA(cudaMallocHost(reinterpret_cast<void**>(&g_host_counts), 8 * sizeof(int64_t)));
    PAMNET_CUDA(cudaMemcpyAsync(g_host_counts, dev, 8 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PAMNET_CUDA(cudaStreamSynchronize(st));
    return 0;
}
}  // namespace

size_t pamnet_plan_build_scratch_bytes(const pamnet_config_t* cfg, int64_t n_nodes, int64_t n_edges_in, int64_t cap_eg) {
    if (!cfg) return 0;
    return build_scratch_layout(cfg->dataset, n_nodes, n_edges_in, cap_eg, nullptr, nullptr);
}

int pamnet_plan_build(const pamnet_config_t* cfg, const float* pos, const int64_t* batch, int64_t n_nodes,
                      int64_t n_graphs, const int64_t* edge_index_in, int64_t n_edges_in, int32_t max_nb,
                      int64_t* eg_buf, int64_t cap_eg, int64_t* el_buf, int64_t cap_el, void* plan_base,
                      size_t cap_base, void* plan_trip, size_t cap_trip, void* scratch, size_t scratch_bytes,
                      pamnet_sizes_t* sizes_out, int64_t* need, void* stream) {
    REQUIRE(cfg); REQUIRE(pos); REQUIRE(batch); REQUIRE(eg_buf); REQUIRE(el_buf); REQUIRE(plan_base); REQUIRE(plan_trip);
    REQUIRE(scratch); REQUIRE(sizes_out); REQUIRE(need);
    const int kind = cfg->dataset;
    PAMNET_CHECK_ARG(kind >= PAMNET_QM9 && kind <= PAMNET_RNA, "bad dataset id %d", kind);
    PAMNET_CHECK_ARG(n_nodes > 0 && n_graphs > 0 && max_nb > 0, "plan_build: n_nodes=%lld n_graphs=%lld",
                     (long long)n_nodes, (long long)n_graphs);
    PAMNET_CHECK_ARG(kind != PAMNET_QM9 || n_edges_in == 0 || edge_index_in, "plan_build: QM9 needs data.edge_index");
    BuildScratch b;
    const size_t want = build_scratch_layout(kind, n_nodes, n_edges_in, cap_eg, scratch, &b);
    PAMNET_CHECK_ARG(scratch_bytes >= want, "plan_build: scratch too small: %zu < %zu", scratch_bytes, want);
    cudaStream_t st = ST(stream);
    for (int i = 0; i < 4; ++i) need[i] = 0;
    PAMNET_CUDA(cudaMemsetAsync(b.counts, 0, 8 * sizeof(int64_t), st));

    // ---- small-molecule batches: the whole front end as three launches and ONE read-back (front_mol.cuh) ----------
    // Opt-in (PAMNET_FRONT=mol) until it has been confirmed on the GPU against the generic kernels below; batches it
    // cannot take (molecules above the per-block capacities, bond lists not grouped by molecule) fall through.
    if (kind == PAMNET_QM9 && front_mol_enabled() && n_graphs <= kMolGraphs && n_graphs <= n_nodes) {
        MolArgs a;
        memset(&a, 0, sizeof(a));
        a.pos = pos; a.batch = batch; a.n_nodes = n_nodes; a.n_graphs = n_graphs;
        a.ei_in = edge_index_in; a.n_edges_in = n_edges_in;
        a.r2 = cfg->cutoff_g * cfg->cutoff_g; a.max_nb = max_nb;
        a.g_dst_row = (cfg->flow == PAMNET_TARGET_TO_SOURCE) ? 0 : 1;
        a.two_hop = cfg->simple ? 0 : 1;
        a.gstart = b.mol_rng; a.estart = b.mol_rng + (n_nodes + 2);                      // n_graphs + 1 <= n_nodes + 1 entries each
        a.mc_eg = b.deg_a; a.mc_el = b.ptr_a; a.mc_t2 = b.deg_b; a.mc_t1 = b.ptr_b;      // n_graphs <= n_nodes entries each
        a.counts = reinterpret_cast<unsigned long long*>(b.counts);
        PAMNET_TRY(mol_count(a, st));
        PAMNET_TRY(read_counts(b.counts, st));
        if (g_host_counts[5] == 0 && g_host_counts[4] == n_edges_in) {
            pamnet_sizes_t sz;
            memset(&sz, 0, sizeof(sz));
            sz.n_nodes = n_nodes; sz.n_graphs = n_graphs;
            sz.n_edges_g = g_host_counts[0]; sz.n_edges_l = g_host_counts[1];
            sz.n_t2 = g_host_counts[2]; sz.n_t1 = g_host_counts[3];
            size_t bb = 0, tb = 0;
            plan_layout(sz, nullptr, nullptr, nullptr, &bb, &tb);
            need[0] = sz.n_edges_g; need[1] = sz.n_edges_l; need[2] = (int64_t)bb; need[3] = (int64_t)tb;
            if (sz.n_edges_g > cap_eg || sz.n_edges_l > cap_el || bb > cap_base || tb > cap_trip) return 1;
            Plan p;
            plan_layout(sz, plan_base, plan_trip, &p, nullptr, nullptr);
            a.Eg = sz.n_edges_g; a.El = sz.n_edges_l;
            a.eg_out = eg_buf;
            a.el_out = (sz.n_edges_l == n_edges_in) ? nullptr : el_buf;      // nothing dropped: the caller's list is used in place
            a.n2g = p.n2g; a.gptr = p.gptr;
            a.g_ptr = p.g_ptr; a.g_src = p.g_src; a.g_dst = p.g_dst; a.g_eid = p.g_eid; a.g_optr = p.g_optr; a.g_opos = p.g_opos;
            a.l_ptr = p.l_ptr; a.l_src = p.l_src; a.l_dst = p.l_dst; a.l_eid = p.l_eid; a.l_optr = p.l_optr; a.l_opos = p.l_opos;
            a.t_split = p.t_split; a.t_cnt = p.t_cnt; a.t_ptr = p.t_ptr; a.tt_ptr = p.tt_ptr;
            a.t_gather = p.t_gather; a.t_owner = p.t_owner; a.tt_t = p.tt_t;
            a.t_angle = p.t_angle; a.dist_g = p.dist_g; a.dist_l = p.dist_l;
            PAMNET_TRY(mol_fill(a, st));
            *sizes_out = sz;
            return 0;
        }
        for (int i = 0; i < 4; ++i) need[i] = 0;
        PAMNET_CUDA(cudaMemsetAsync(b.counts, 0, 8 * sizeof(int64_t), st));
    }

    // ---- phase 1: edge counts (models.py:110,115 / 128,131-136 / 143-157) ----------------------------------------
    // large graphs (PDBbind complexes): cell-list radius search instead of the per-graph scan, same edge list
    const bool use_grid = kind != PAMNET_RNA && radius_grid_preferred(n_nodes, n_graphs);
    if (kind == PAMNET_RNA) {
        PAMNET_TRY(knn(pos, batch, n_nodes, kKnnK, b.nbr, b.d2, st));
        PAMNET_TRY(knn_edges_count(b.nbr, pos, n_nodes, kKnnK, cfg->cutoff_g, b.deg_a, b.ptr_a, b.counts + 0, st));
        PAMNET_TRY(knn_edges_count(b.nbr, pos, n_nodes, kKnnK, cfg->cutoff_l, b.deg_b, b.ptr_b, b.counts + 1, st));
    } else {
        if (use_grid)
            PAMNET_TRY(radius_grid_count(pos, batch, n_nodes, n_graphs, cfg->cutoff_g, max_nb, 1, b.grid, b.grid_bytes, b.deg_a,
                                         b.ptr_a, b.counts + 0, st));
        else
            PAMNET_TRY(radius_count(pos, batch, n_nodes, cfg->cutoff_g, max_nb, 1, b.deg_a, b.ptr_a, b.counts + 0, st));
        if (kind == PAMNET_QM9)
            PAMNET_TRY(edge_filter_count(edge_index_in, n_edges_in, nullptr, 0.f, b.keep, b.pk, b.counts + 1, st));
    }
    PAMNET_TRY(read_counts(b.counts, st));
    int64_t Eg = g_host_counts[0], El = kind == PAMNET_PDBBIND ? 0 : g_host_counts[1];
    need[0] = Eg; need[1] = El;
    if (Eg > cap_eg || El > cap_el) return 1;

    // ---- phase 2: edge lists, CSRs, triplet counts ------------------------------------------------------------------
    const int64_t* el = el_buf;
    if (kind == PAMNET_RNA) {
        PAMNET_TRY(knn_edges_fill(b.nbr, pos, n_nodes, kKnnK, cfg->cutoff_g, b.ptr_a, Eg, eg_buf, st));
        PAMNET_TRY(knn_edges_fill(b.nbr, pos, n_nodes, kKnnK, cfg->cutoff_l, b.ptr_b, El, el_buf, st));
    } else {
        if (use_grid)
            PAMNET_TRY(radius_grid_fill(pos, batch, n_nodes, n_graphs, cfg->cutoff_g, max_nb, 1, b.grid, b.ptr_a, Eg, eg_buf, st));
        else
            PAMNET_TRY(radius_fill(pos, batch, n_nodes, cfg->cutoff_g, max_nb, 1, b.ptr_a, Eg, eg_buf, st));
        if (kind == PAMNET_QM9) {
            if (El == n_edges_in) el = edge_index_in;            // nothing dropped: use the caller's list in place
            else PAMNET_TRY(edge_filter_fill(edge_index_in, n_edges_in, b.keep, b.pk, El, el_buf, st));
        } else {
            PAMNET_TRY(edge_filter_count(eg_buf, Eg, pos, cfg->cutoff_l, b.keep, b.pk, b.counts + 1, st));
            PAMNET_TRY(read_counts(b.counts, st));
            El = g_host_counts[1];
            need[1] = El;
            if (El > cap_el) return 1;
            if (El == Eg) el = eg_buf;
            else PAMNET_TRY(edge_filter_fill(eg_buf, Eg, b.keep, b.pk, El, el_buf, st));
        }
    }
    pamnet_sizes_t sz;
    memset(&sz, 0, sizeof(sz));
    sz.n_nodes = n_nodes; sz.n_graphs = n_graphs; sz.n_edges_g = Eg; sz.n_edges_l = El;
    size_t bb = 0, tb = 0;
    plan_layout(sz, nullptr, nullptr, nullptr, &bb, &tb);
    need[2] = (int64_t)bb;
    if (bb > cap_base) return 1;
    PAMNET_TRY(plan_count(*cfg, sz, eg_buf, el, batch, plan_base, b.counts + 2, st));
    PAMNET_TRY(read_counts(b.counts, st));
    sz.n_t2 = g_host_counts[2]; sz.n_t1 = g_host_counts[3];
    plan_layout(sz, nullptr, nullptr, nullptr, &bb, &tb);
    need[3] = (int64_t)tb;
    if (tb > cap_trip) return 1;

    // ---- phase 3: triplet lists, angles, distances -------------------------------------------------------------------
    PAMNET_TRY(plan_fill(*cfg, sz, pos, plan_base, plan_trip, st));
    *sizes_out = sz;
    return 0;
}

size_t pamnet_workspace_bytes(const pamnet_config_t* cfg, const pamnet_sizes_t* sz) {
    if (!cfg || !sz) return 0;
    return workspace_bytes(*cfg, *sz);
}
int pamnet_model_forward(const pamnet_config_t* cfg, const pamnet_sizes_t* sz, const pamnet_sbf_consts_t* sbf,
                         const float* params, const float* node_in, const float* sign, const float* pos,
                         void* plan_base, void* plan_trip, void* workspace, size_t workspace_bytes,
                         int32_t save_for_backward, float* out, void* stream, void* aux_stream,
                         void* prepared_weights) {
    REQUIRE(cfg); REQUIRE(sz); REQUIRE(sbf); REQUIRE(params); REQUIRE(node_in); REQUIRE(pos); REQUIRE(plan_base);
    REQUIRE(plan_trip); REQUIRE(workspace); REQUIRE(out);
    (void)save_for_backward;
    return model_forward(*cfg, *sz, *sbf, params, node_in, sign, pos, plan_base, plan_trip, workspace,
                         workspace_bytes, out, ST(stream), ST(aux_stream), prepared_weights);
}
size_t pamnet_prepared_weights_bytes(const pamnet_config_t* cfg) { return cfg ? prepared_weights_bytes(*cfg) : 0; }
int pamnet_prepare_weights(const pamnet_config_t* cfg, const float* params, void* prepared_weights, void* stream) {
    REQUIRE(cfg); REQUIRE(params); REQUIRE(prepared_weights);
    return prepare_weights(*cfg, params, prepared_weights, ST(stream));
}
int pamnet_model_backward(const pamnet_config_t* cfg, const pamnet_sizes_t* sz, const pamnet_sbf_consts_t* sbf,
                          const float* params, const float* node_in, const float* sign, const float* pos,
                          void* plan_base, void* plan_trip, void* workspace, size_t workspace_bytes,
                          const float* grad_out, float* grad_params, void* stream, void* aux_stream,
                          void* prepared_weights) {
    REQUIRE(cfg); REQUIRE(sz); REQUIRE(sbf); REQUIRE(params); REQUIRE(node_in); REQUIRE(pos); REQUIRE(plan_base);
    REQUIRE(plan_trip); REQUIRE(workspace); REQUIRE(grad_out); REQUIRE(grad_params);
    return model_backward(*cfg, *sz, *sbf, params, node_in, sign, pos, plan_base, plan_trip, workspace,
                          workspace_bytes, grad_out, grad_params, ST(stream), ST(aux_stream), prepared_weights);
}

int pamnet_comm_unique_id(void* out128) { REQUIRE(out128); return comm_unique_id(out128); }
int pamnet_comm_init(const void* id128, int32_t rank, int32_t world) { REQUIRE(id128); return comm_init(id128, rank, world); }
int pamnet_comm_enable(int32_t on) { return comm_enable(on); }
int pamnet_comm_destroy(void) { return comm_destroy(); }

int pamnet_grad_buckets(int32_t enable) { set_grad_buckets(enable); return 0; }
int pamnet_wait_grad_bucket(int32_t half, void* stream) { return wait_grad_bucket(half, ST(stream)); }
int pamnet_grad_bucket_range(const pamnet_config_t* cfg, int32_t half, int64_t* lo, int64_t* hi) {
    REQUIRE(cfg); REQUIRE(lo); REQUIRE(hi);
    return grad_bucket_range(*cfg, half, lo, hi);
}

int64_t pamnet_debug_ws_offset(const pamnet_config_t* cfg, const pamnet_sizes_t* sz, const char* name, int32_t half) {
    if (!cfg || !sz || !name) return -1;
    return debug_ws_offset(*cfg, *sz, name, half);
}
// plan arrays for tests: which = 0 g_ptr, 1 g_src, 2 g_eid, 3 l_ptr, 4 l_src, 5 l_dst, 6 l_eid, 7 t_ptr, 8 t_gather,
// 9 t_owner, 10 t_split, 11 dist_g, 12 dist_l, 13 t_angle, 14 g_dst, 15 g_optr, 16 g_opos, 17 l_optr, 18 l_opos, 19 t_cnt,
// 20 tt_ptr, 21 tt_t, 22 n2g, 23 gptr; returns the byte offset inside base (0..12) or trip blob
int64_t pamnet_debug_plan_offset(const pamnet_sizes_t* sz, int32_t which, int32_t* in_trip) {
    if (!sz) return -1;
    Plan p;
    char* b = reinterpret_cast<char*>(0x1000);
    char* t = reinterpret_cast<char*>(0x1000);
    plan_layout(*sz, b, t, &p, nullptr, nullptr);
    const void* tab[] = {p.g_ptr, p.g_src, p.g_eid, p.l_ptr, p.l_src, p.l_dst, p.l_eid, p.t_ptr, p.t_gather,
                         p.t_owner, p.t_split, p.dist_g, p.dist_l, p.t_angle,
                         p.g_dst, p.g_optr, p.g_opos, p.l_optr, p.l_opos, p.t_cnt, p.tt_ptr, p.tt_t, p.n2g, p.gptr};
    if (which < 0 || which > 23) return -1;
    const bool trip = (which == 8 || which == 9 || which == 13 || which == 21);
    if (in_trip) *in_trip = trip;
    return (int64_t)(reinterpret_cast<const char*>(tab[which]) - (trip ? t : b));
}

int64_t pamnet_debug_launch_count(void) { return g_launches.load(); }
void pamnet_debug_profile_begin(void) {
    g_prof_recs.clear();
    g_prof_used = 0;
    g_prof_on = true;
}
// ms[KC_COUNT], launches[KC_COUNT], bytes[KC_COUNT]; synchronises the device
int pamnet_debug_profile_end(double* ms, int64_t* launches, double* bytes, double* flops) {
    g_prof_on = false;
    PAMNET_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < KC_COUNT; ++i) { ms[i] = 0; launches[i] = 0; bytes[i] = 0; if (flops) flops[i] = 0; }
    for (auto& r : g_prof_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) continue;
        ms[r.cls] += t; launches[r.cls] += 1; bytes[r.cls] += r.bytes;
        if (flops) flops[r.cls] += r.flops;
    }
    g_prof_recs.clear();
    g_prof_used = 0;
    return KC_COUNT;
}

// Timeline of the records since profile_begin (call INSTEAD of profile_end): per launch the kernel class, a stream
// tag (0 = first stream seen, 1 = second, ...) and start / end in ms relative to the first record.  Returns the
// number of records written (<= cap).
int pamnet_debug_profile_timeline(int32_t* cls, int32_t* stream_tag, float* t0_ms, float* t1_ms, int32_t cap) {
    g_prof_on = false;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    std::vector<cudaStream_t> streams;
    int n = 0;
    for (auto& r : g_prof_recs) {
        if (n >= cap) break;
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, g_prof_recs[0].e0, r.e0) != cudaSuccess) continue;
        if (cudaEventElapsedTime(&b, g_prof_recs[0].e0, r.e1) != cudaSuccess) continue;
        int tag = -1;
        for (size_t i = 0; i < streams.size(); ++i) if (streams[i] == r.st) tag = (int)i;
        if (tag < 0) { streams.push_back(r.st); tag = (int)streams.size() - 1; }
        cls[n] = r.cls; stream_tag[n] = tag; t0_ms[n] = a; t1_ms[n] = b;
        ++n;
    }
    g_prof_recs.clear();
    g_prof_used = 0;
    return n;
}

int pamnet_optimizer_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, float* ema_shadow,
                          int64_t n, const int64_t* skip_ranges, int32_t n_skip, int64_t step, float lr, float beta1,
                          float beta2, float eps, float weight_decay, float max_norm, float ema_decay,
                          int32_t write_clipped_grad, double* sumsq_dev, void* stream) {
    REQUIRE(params); REQUIRE(grads); REQUIRE(exp_avg); REQUIRE(exp_avg_sq);
    if (n % 4 != 0 || n < 0 || step < 1 || n_skip < 0 || n_skip > kOptMaxSkip || (n_skip > 0 && !skip_ranges)) {
        set_error("optimizer_step: n=%lld must be a non-negative multiple of 4, step >= 1, n_skip in [0, %d]",
                  (long long)n, kOptMaxSkip);
        return -1;
    }
    OptimArgs a;
    memset(&a, 0, sizeof(a));
    a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq; a.shadow = ema_shadow; a.sumsq = sumsq_dev;
    a.n4 = n / 4; a.n_skip = n_skip; a.write_clipped_grad = write_clipped_grad;
    for (int i = 0; i < n_skip; ++i) { a.skip_begin[i] = skip_ranges[2 * i]; a.skip_end[i] = skip_ranges[2 * i + 1]; }
    // bias corrections as torch.optim.Adam computes them (python floats = double), rounded once to fp32
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    a.step_size = (float)((double)lr / bc1);
    a.bc2_sqrt = (float)sqrt(bc2);
    a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm;
    a.ema_decay = ema_decay;
    return optimizer_step(a, ST(stream));
}

int pamnet_loss(const float* out, const float* y, int64_t n, int32_t kind, float* loss_dev, float* grad_out,
                void* stream) {
    REQUIRE(out); REQUIRE(y); REQUIRE(loss_dev); REQUIRE(grad_out);
    return loss_forward_backward(out, y, n, kind, loss_dev, grad_out, ST(stream));
}

int pamnet_collate(const int64_t* table, int64_t n_ids, const int64_t* node_ptr, const int64_t* edge_ptr,
                   const float* x_all, const float* pos_all, const int64_t* ei_all, int64_t e_all, const float* y_all,
                   int64_t n_edges, float* x, float* pos, int64_t* edge_index, int64_t* batch, float* y, void* stream) {
    REQUIRE(table); REQUIRE(node_ptr); REQUIRE(edge_ptr); REQUIRE(x_all); REQUIRE(pos_all); REQUIRE(y_all);
    REQUIRE(x); REQUIRE(pos); REQUIRE(batch); REQUIRE(y);
    PAMNET_CHECK_ARG(n_edges == 0 || (ei_all && edge_index), "collate: %lld bonds but no edge lists", (long long)n_edges);
    CollateArgs a;
    a.table = table; a.n_ids = n_ids; a.node_ptr = node_ptr; a.edge_ptr = edge_ptr; a.x_all = x_all; a.pos_all = pos_all;
    a.ei_all = ei_all; a.e_all = e_all; a.y_all = y_all; a.n_edges = n_edges; a.x = x; a.pos = pos;
    a.edge_index = edge_index; a.batch = batch; a.y = y;
    return collate(a, ST(stream));
}

int pamnet_scatter_add(const float* src, const int64_t* index, int64_t n_rows, int64_t width, int64_t dim_size,
                       float* out, void* stream) {
    REQUIRE(out);
    return scatter_add_rows(src, index, n_rows, width, dim_size, out, ST(stream));
}
int pamnet_bessel_rbf(const float* dist, int64_t n_edges, const float* freq, float cutoff, float* rbf,
                      void* stream) {
    REQUIRE(freq);
    return rbf_forward(dist, n_edges, freq, cutoff, rbf, ST(stream));
}
int pamnet_sbf_radial(const pamnet_sbf_consts_t* sbf, const float* dist, int64_t n_edges, float cutoff,
                      float* radial, void* stream) {
    REQUIRE(sbf);
    SbfTables tab;
    make_sbf_tables(*sbf, &tab);
    return sbf_radial(tab, dist, n_edges, cutoff, radial, ST(stream));
}
int pamnet_spherical_basis(const pamnet_sbf_consts_t* sbf, const float* radial, const float* angle,
                           const int64_t* gather, int64_t n_trip, float* out, void* stream) {
    REQUIRE(sbf);
    SbfTables tab;
    make_sbf_tables(*sbf, &tab);
    return sbf_combine(tab, radial, angle, gather, n_trip, out, ST(stream));
}
int pamnet_linear(const float* x, int64_t n_rows, int32_t n_in, int32_t n_out, const float* w, const float* b,
                  int32_t act, float* y, void* stream) {
    REQUIRE(w); REQUIRE(y);
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = GEMM_NT; a.epi = act ? EPI_BIAS_SILU : EPI_BIAS; a.M = (int)n_rows; a.N = n_out; a.K = n_in;
    a.ksplit = 1; a.nslots = 1;
    a.slot[0].A = x; a.slot[0].lda = n_in; a.slot[0].B = w; a.slot[0].ldb = n_in; a.slot[0].bias = b;
    a.slot[0].C = y; a.slot[0].ldc = n_out;
    return gemm_launch(a, ST(stream));
}
// Generic fp32 GEMM hook (unit tests of the NT / NN / TN paths): C[M,N] (+)= op(A) op(B)
int pamnet_gemm(int32_t mode, const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
                int32_t M, int32_t N, int32_t K, int32_t ksplit, float* dbias, void* stream) {
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = mode; a.epi = EPI_NONE; a.M = M; a.N = N; a.K = K; a.ksplit = ksplit; a.nslots = 1;
    a.slot[0].A = A; a.slot[0].lda = lda; a.slot[0].B = B; a.slot[0].ldb = ldb; a.slot[0].C = C; a.slot[0].ldc = ldc;
    a.slot[0].C2 = dbias;
    return gemm_launch(a, ST(stream));
}

// debugging aid: clock64 timeline of the tensor-core GEMM's CTA 0 (library built with -DPAMNET_TC_TRACE only)
int pamnet_debug_tc_trace(long long* out, int32_t n) { return tc_trace_read(out, n); }
int pamnet_debug_chain_trace(long long* out, int32_t n) { return chain_trace_read(out, n); }

}  // extern "C"
