// Host emulation of the per-molecule front end: compiles csrc/front_mol.cuh with one "thread" per block (PM_NT = 1,
// barriers are no-ops) so the CPU test suite can check its integer logic without a GPU -- or, with -DPM_HOST_THREADS=k
// -fsanitize=thread -pthread, with k real threads per block and a pthread barrier, so that ThreadSanitizer reports any
// pair of conflicting shared-memory accesses that no barrier separates.
// Test infrastructure only -- never loaded by the package.  Build: g++ -O1 -ffp-contract=off -shared -fPIC.
#include "../../physics-aware-multiplex-gnn_b200/csrc/front_mol.cuh"

#include <vector>

#ifdef PM_HOST_THREADS
#include <pthread.h>

#include <thread>
thread_local int pm_host_tid = 0;
static pthread_barrier_t g_barrier;
void pm_host_barrier() { pthread_barrier_wait(&g_barrier); }
// one block = PM_HOST_THREADS real threads running the same body
template <class F>
static void run_block(F body) {
    pthread_barrier_init(&g_barrier, nullptr, PM_HOST_THREADS);
    std::vector<std::thread> th;
    for (int t = 0; t < PM_HOST_THREADS; ++t) th.emplace_back([=] { pm_host_tid = t; body(); });
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&g_barrier);
}
#else
template <class F>
static void run_block(F body) { body(); }
#endif

using namespace pamnet;

// iarr: n2g gptr g_ptr g_src g_dst g_eid g_optr g_opos l_ptr l_src l_dst l_eid l_optr l_opos t_split t_cnt t_ptr tt_ptr
//       t_gather t_owner tt_t (21 int32 arrays); farr: t_angle dist_g dist_l
extern "C" int front_mol_host(int pass, const float* pos, const int64_t* batch, int64_t n_nodes, int64_t n_graphs,
                              const int64_t* ei_in, int64_t n_edges_in, float r2, int max_nb, int g_dst_row,
                              int two_hop, int32_t* mc, unsigned long long* counts, int64_t Eg, int64_t El,
                              int64_t* eg_out, int64_t* el_out, int32_t** iarr, float** farr) {
    if (n_graphs > kMolGraphs) return -1;
    MolArgs a{};
    a.pos = pos; a.batch = batch; a.n_nodes = n_nodes; a.n_graphs = n_graphs; a.ei_in = ei_in;
    a.n_edges_in = n_edges_in; a.r2 = r2; a.max_nb = max_nb; a.g_dst_row = g_dst_row; a.two_hop = two_hop;
    a.mc_eg = mc; a.mc_el = mc + n_graphs; a.mc_t2 = mc + 2 * n_graphs; a.mc_t1 = mc + 3 * n_graphs;
    a.counts = counts;
    static MolSmem s;
    static std::vector<int32_t> ranges;
    if (pass == 0) {
        ranges.assign(2 * (n_graphs + 1), -12345);          // poison: an unwritten entry must never be used
        a.gstart = ranges.data(); a.estart = ranges.data() + n_graphs + 1;
        for (int64_t t = 0; t < 3; ++t) mol_ranges_body(a, 2 - t, 3);      // three "threads", any order
        for (int m = 0; m < n_graphs; ++m) run_block([&a, m] { mol_count_body(a, s, m); });
        return 0;
    }
    a.gstart = ranges.data(); a.estart = ranges.data() + n_graphs + 1;      // tables of the preceding count pass
    a.Eg = Eg; a.El = El; a.eg_out = eg_out; a.el_out = el_out;
    int32_t** p = iarr;
    a.n2g = p[0]; a.gptr = p[1]; a.g_ptr = p[2]; a.g_src = p[3]; a.g_dst = p[4]; a.g_eid = p[5]; a.g_optr = p[6];
    a.g_opos = p[7]; a.l_ptr = p[8]; a.l_src = p[9]; a.l_dst = p[10]; a.l_eid = p[11]; a.l_optr = p[12];
    a.l_opos = p[13]; a.t_split = p[14]; a.t_cnt = p[15]; a.t_ptr = p[16]; a.tt_ptr = p[17]; a.t_gather = p[18];
    a.t_owner = p[19]; a.tt_t = p[20];
    a.t_angle = farr[0]; a.dist_g = farr[1]; a.dist_l = farr[2];
    // blocks in reverse order: the result must not depend on the order the molecules are processed in
    for (int m = (int)n_graphs - 1; m >= 0; --m) run_block([&a, m] { mol_fill_body(a, s, m); });
    return 0;
}

extern "C" void front_mol_caps(int* out) {
    out[0] = kMolAtoms; out[1] = kMolEdges; out[2] = kMolTrip; out[3] = kMolGraphs;
}
