// Per-molecule front end for QM9-shaped batches (models.py:104-115,159-177 + the plan of graph.cuh): ONE thread block
// per molecule builds, entirely in shared memory, the radius graph, both destination-sorted CSRs, the out-CSRs, the
// merged triplet lists with their angles, the reverse (gather-keyed) triplet lists and the edge lengths -- the same
// arrays, bit for bit, that the 20 generic launches of pamnet_plan_build (graph.cu) produce.
//
// Why it is legal: every edge and triplet stays inside one molecule, nodes are grouped by molecule
// (non-decreasing batch vector), and PyG's collate groups data.edge_index by molecule as well.  Then the slots of all
// per-edge arrays of a molecule are contiguous and start at the prefix sum of the earlier molecules' counts.  Three
// launches: pass 0 finds every molecule's atom and bond-list range and checks that both key sequences are non-decreasing;
// pass 1 (one block per molecule) checks that every bond of the range joins two of its atoms and the per-block
// capacities, and counts; the host reads the totals once; pass 2 fills.  Anything the checks reject falls back to the
// generic kernels.
//
// The body is written as block-strided loops separated by barriers, with no warp intrinsics, so that the very same
// source also compiles for the host with one "thread" per block (tests/host_emul): the CPU test suite checks the integer
// logic against an independent numpy restatement without a GPU.
#pragma once
#include "geom.cuh"

#ifdef __CUDACC__
#define PM_TID ((int)threadIdx.x)
#define PM_NT ((int)blockDim.x)
#define PM_SYNC() __syncthreads()
#define PM_POPC(x) __popcll(x)
PAMNET_HD int pm_atomic_add(int* p, int v) { return atomicAdd(p, v); }
PAMNET_HD void pm_atomic_add64(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
PAMNET_HD void pm_atomic_or64(unsigned long long* p, unsigned long long v) { atomicOr(p, v); }
#elif defined(PM_HOST_THREADS)
// host build with PM_HOST_THREADS real threads per block (tests/host_emul, run under ThreadSanitizer): the harness
// provides the thread index and a barrier, the atomics are the compiler's
extern thread_local int pm_host_tid;
void pm_host_barrier();
#define PM_TID pm_host_tid
#define PM_NT PM_HOST_THREADS
#define PM_SYNC() pm_host_barrier()
#define PM_POPC(x) __builtin_popcountll(x)
static inline int pm_atomic_add(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline void pm_atomic_add64(unsigned long long* p, unsigned long long v) { __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline void pm_atomic_or64(unsigned long long* p, unsigned long long v) { __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
#else
#define PM_TID 0
#define PM_NT 1
#define PM_SYNC() ((void)0)
#define PM_POPC(x) __builtin_popcountll(x)
static inline int pm_atomic_add(int* p, int v) { int o = *p; *p += v; return o; }
static inline void pm_atomic_add64(unsigned long long* p, unsigned long long v) { *p += v; }
static inline void pm_atomic_or64(unsigned long long* p, unsigned long long v) { *p |= v; }
#endif

namespace pamnet {

constexpr int kMolAtoms = 64;        // atoms per molecule (adjacency rows are 64-bit masks); QM9 has <= 29
constexpr int kMolEdges = 512;       // bond-list entries per molecule, both directions; QM9 has <= ~60
constexpr int kMolTrip = 6144;       // two-hop + one-hop entries per molecule
constexpr int kMolGraphs = 4096;     // molecules per batch (every block sums the counts of the earlier ones)
constexpr int kMolThreads = 256;

// flag bits in counts[5]
constexpr unsigned long long kMolFlagCaps = 1;      // a molecule exceeds a per-block capacity
constexpr unsigned long long kMolFlagGroup = 2;     // the bond list is not grouped by molecule / crosses molecules

struct MolArgs {
    const float* pos;              // [n_nodes, 3]
    const int64_t* batch;          // [n_nodes] non-decreasing
    int64_t n_nodes, n_graphs;
    const int64_t* ei_in;          // data.edge_index [2, n_edges_in]
    int64_t n_edges_in;
    float r2;                      // cutoff_g squared
    int max_nb;                    // torch_cluster max_num_neighbors (self included, like radius_kernel)
    int g_dst_row;                 // row of edge_index_g that is the aggregation target (graph.cu:plan_count)
    int two_hop;                   // 0 for PAMNet_s
    int32_t *gstart, *estart;      // [n_graphs + 1]: first atom / first bond-list entry of every molecule (mol_ranges_body)
    int32_t *mc_eg, *mc_el, *mc_t2, *mc_t1;     // per-molecule counts [n_graphs]
    unsigned long long* counts;    // [8]: E_g, E_l, T2, T1, covered bond-list entries, flags
    // ---- fill pass only ----
    int64_t Eg, El;                // totals (row strides of the int64 edge lists)
    int64_t* eg_out;               // [2, Eg]
    int64_t* el_out;               // [2, El] or null when nothing was filtered (the plan then refers to ei_in)
    int32_t *n2g, *gptr;
    int32_t *g_ptr, *g_src, *g_dst, *g_eid, *g_optr, *g_opos;
    int32_t *l_ptr, *l_src, *l_dst, *l_eid, *l_optr, *l_opos;
    int32_t *t_split, *t_cnt, *t_ptr, *tt_ptr;
    int32_t *t_gather, *t_owner, *tt_t;
    float *t_angle, *dist_g, *dist_l;
};

struct MolSmem {
    unsigned long long adj[kMolAtoms];     // adj[q] bit n: API edge (row0 = q, row1 = n)
    unsigned long long adjT[kMolAtoms];
    float px[kMolAtoms], py[kMolAtoms], pz[kMolAtoms];
    int pref_q[kMolAtoms + 1], pref_t[kMolAtoms + 1];      // exclusive prefixes of popc(adj), popc(adjT)
    int l_in[kMolAtoms + 1], l_out[kMolAtoms + 1];         // local in / out degrees, then exclusive prefixes
    int va[kMolEdges + 1], vb[kMolEdges + 1], vc[kMolEdges + 1], vx[kMolEdges + 1];   // scan operands / results
    short le_src[kMolEdges], le_dst[kMolEdges];            // kept bonds in API order (atom ids inside the molecule)
    short ls_src[kMolEdges], ls_dst[kMolEdges];            // the same per CSR slot
    short t_n2[kMolEdges];
    short tg[kMolTrip];                                    // gathered slot per triplet
    int a0, a1, lo, hi, n_kept, flag, bail;
    int eg_m, t2_m, t1_m;
    int og, ol, ot;
};

// ---- pass 0: atom and bond-list ranges of every molecule (grid-stride over atoms and bonds, any grid size) ---------------
// gstart[g] = first atom with batch >= g, estart[g] = first bond whose source atom belongs to a graph >= g; both need
// non-decreasing keys (else kMolFlagGroup).  i runs to n inclusive so that the entries up to [n_graphs] get written.
PAMNET_HD void mol_ranges_body(const MolArgs& A, int64_t gtid, int64_t gnt) {
    const int64_t G = A.n_graphs;
    for (int64_t i = gtid; i <= A.n_nodes; i += gnt) {
        const int64_t prev = i == 0 ? -1 : A.batch[i - 1];
        const int64_t cur = i == A.n_nodes ? G : A.batch[i];
        if (cur < prev || cur < 0 || cur > G || (i < A.n_nodes && cur == G)) { pm_atomic_or64(&A.counts[5], kMolFlagGroup); continue; }
        for (int64_t g = prev + 1; g <= cur; ++g) A.gstart[g] = (int32_t)i;
    }
    for (int64_t e = gtid; e <= A.n_edges_in; e += gnt) {
        int64_t prev = -1, cur = G;
        if (e > 0) {
            const int64_t v = A.ei_in[e - 1];
            prev = (v >= 0 && v < A.n_nodes) ? A.batch[v] : -2;
        }
        if (e < A.n_edges_in) {
            const int64_t v = A.ei_in[e];
            cur = (v >= 0 && v < A.n_nodes) ? A.batch[v] : -2;
        }
        if (prev == -2 || cur == -2 || cur < prev || cur > G || (e < A.n_edges_in && cur == G)) {
            pm_atomic_or64(&A.counts[5], kMolFlagGroup);
            continue;
        }
        for (int64_t g = prev + 1; g <= cur; ++g) A.estart[g] = (int32_t)e;
    }
}

// exclusive scan of v[0..n) into out[0..n] (out[n] = total); a, b: scratch of n ints; v, a, b, out distinct
PAMNET_HD void mol_scan(const int* v, int n, int* a, int* b, int* out) {
    for (int i = PM_TID; i < n; i += PM_NT) a[i] = v[i];
    PM_SYNC();
    int* src = a;
    int* dst = b;
    for (int off = 1; off < n; off <<= 1) {
        for (int i = PM_TID; i < n; i += PM_NT) dst[i] = src[i] + (i >= off ? src[i - off] : 0);
        PM_SYNC();
        int* t = src; src = dst; dst = t;
    }
    for (int i = PM_TID; i < n; i += PM_NT) out[i + 1] = src[i];
    if (PM_TID == 0) out[0] = 0;
    PM_SYNC();
}

// Shared first part of both passes: node and bond ranges, radius adjacency, validated + self-loop-free bond list in
// API order.  Returns false (block-uniform) when the molecule cannot be handled; the flag is then already recorded.
PAMNET_HD bool mol_setup(const MolArgs& A, MolSmem& s, int m) {
    if (PM_TID == 0) {
        // pass 0 found the batch vector or the bond list not grouped: the range tables are incomplete, touch nothing
        s.bail = A.counts[5] != 0;
        s.a0 = s.bail ? 0 : A.gstart[m]; s.a1 = s.bail ? 0 : A.gstart[m + 1];
        s.lo = s.bail ? 0 : A.estart[m]; s.hi = s.bail ? 0 : A.estart[m + 1];
        s.flag = 0; s.n_kept = 0; s.eg_m = 0; s.t2_m = 0; s.t1_m = 0; s.og = 0; s.ol = 0; s.ot = 0;
    }
    for (int i = PM_TID; i <= kMolAtoms; i += PM_NT) { s.l_in[i] = 0; s.l_out[i] = 0; }
    PM_SYNC();
    if (s.bail) return false;
    const int a0 = s.a0, nA = s.a1 - s.a0, lo = s.lo, nE = s.hi - s.lo;
    if (nA > kMolAtoms || nE > kMolEdges) {
        if (PM_TID == 0) {
            pm_atomic_or64(&A.counts[5], kMolFlagCaps);
            pm_atomic_add64(&A.counts[4], (unsigned long long)nE);
        }
        return false;
    }
    for (int q = PM_TID; q < nA; q += PM_NT) {
        s.px[q] = A.pos[3 * (int64_t)(a0 + q)];
        s.py[q] = A.pos[3 * (int64_t)(a0 + q) + 1];
        s.pz[q] = A.pos[3 * (int64_t)(a0 + q) + 2];
    }
    // bond list: both ends must be atoms of this molecule; self loops are dropped (models.py:62-63)
    for (int e = PM_TID; e < nE; e += PM_NT) {
        const int64_t u = A.ei_in[lo + e] - a0, v = A.ei_in[A.n_edges_in + lo + e] - a0;
        const bool ok = u >= 0 && u < nA && v >= 0 && v < nA;
        if (!ok) s.flag = 1;
        s.va[e] = (ok && u != v) ? 1 : 0;
        s.le_src[e] = (short)(ok ? u : 0);      // staged unfiltered, compacted below
        s.le_dst[e] = (short)(ok ? v : 0);
    }
    PM_SYNC();
    if (s.flag) {
        if (PM_TID == 0) {
            pm_atomic_or64(&A.counts[5], kMolFlagGroup);
            pm_atomic_add64(&A.counts[4], (unsigned long long)nE);
        }
        return false;
    }
    // radius neighbours in ascending index order, at most max_nb hits with the atom itself counted (radius_kernel)
    for (int q = PM_TID; q < nA; q += PM_NT) {
        unsigned long long bits = 0;
        int found = 0;
        for (int n = 0; n < nA && found < A.max_nb; ++n) {
            const float d2 = canon_d2(s.px[q], s.py[q], s.pz[q], s.px[n], s.py[n], s.pz[n]);
            if (d2 <= A.r2) {
                ++found;
                if (n != q) bits |= 1ull << n;
            }
        }
        s.adj[q] = bits;
    }
    // order-preserving compaction of the kept bonds
    mol_scan(s.va, nE, s.vb, s.vc, s.vx);       // vx[e] = position of entry e among the kept ones
    short ks = 0, kd = 0;
    // (two-step: read the staged entry, barrier, write it to its compacted place -- positions only move down)
    for (int base = 0; base < nE; base += PM_NT) {
        const int e = base + PM_TID;
        const bool keep = e < nE && s.va[e];
        if (keep) { ks = s.le_src[e]; kd = s.le_dst[e]; }
        PM_SYNC();
        if (keep) { s.le_src[s.vx[e]] = ks; s.le_dst[s.vx[e]] = kd; }
        PM_SYNC();
    }
    if (PM_TID == 0) s.n_kept = s.vx[nE];
    PM_SYNC();
    return true;
}

// ---- pass 1: counts ---------------------------------------------------------------------------------------------
PAMNET_HD void mol_count_body(const MolArgs& A, MolSmem& s, int m) {
    if (!mol_setup(A, s, m)) {
        if (PM_TID == 0) { A.mc_eg[m] = 0; A.mc_el[m] = 0; A.mc_t2[m] = 0; A.mc_t1[m] = 0; }
        return;
    }
    const int nA = s.a1 - s.a0, nL = s.n_kept;
    for (int q = PM_TID; q < nA; q += PM_NT) pm_atomic_add(&s.eg_m, PM_POPC(s.adj[q]));
    for (int e = PM_TID; e < nL; e += PM_NT) pm_atomic_add(&s.l_in[s.le_dst[e]], 1);
    PM_SYNC();
    // bond j -> i: two-hop entries = bonds into j that do not come from i; one-hop entries = bonds into i
    // (plan_tcount_kernel; self loops are gone, so no bond into i comes from i)
    for (int e = PM_TID; e < nL; e += PM_NT) {
        const int j = s.le_src[e], i = s.le_dst[e];
        int n2 = 0;
        if (A.two_hop) {
            n2 = s.l_in[j];
            for (int f = 0; f < nL; ++f) n2 -= (s.le_src[f] == i && s.le_dst[f] == j);
        }
        pm_atomic_add(&s.t2_m, n2);
        pm_atomic_add(&s.t1_m, s.l_in[i]);
    }
    PM_SYNC();
    if (PM_TID == 0) {
        const bool big = s.t2_m + s.t1_m > kMolTrip;
        if (big) pm_atomic_or64(&A.counts[5], kMolFlagCaps);
        A.mc_eg[m] = s.eg_m; A.mc_el[m] = nL; A.mc_t2[m] = s.t2_m; A.mc_t1[m] = s.t1_m;
        pm_atomic_add64(&A.counts[0], (unsigned long long)s.eg_m);
        pm_atomic_add64(&A.counts[1], (unsigned long long)nL);
        pm_atomic_add64(&A.counts[2], (unsigned long long)s.t2_m);
        pm_atomic_add64(&A.counts[3], (unsigned long long)s.t1_m);
        pm_atomic_add64(&A.counts[4], (unsigned long long)(s.hi - s.lo));
    }
}

// ---- pass 2: everything else ------------------------------------------------------------------------------------
PAMNET_HD void mol_fill_body(const MolArgs& A, MolSmem& s, int m) {
    if (!mol_setup(A, s, m)) return;        // cannot happen: the host only launches this pass when pass 1 raised no flag
    const int a0 = s.a0, nA = s.a1 - s.a0, nL = s.n_kept;
    // slot / edge-id / triplet offsets of this molecule = totals of the earlier ones
    for (int g = PM_TID; g < m; g += PM_NT) {
        pm_atomic_add(&s.og, A.mc_eg[g]);
        pm_atomic_add(&s.ol, A.mc_el[g]);
        pm_atomic_add(&s.ot, A.mc_t2[g] + A.mc_t1[g]);
    }
    for (int d = PM_TID; d < nA; d += PM_NT) {
        unsigned long long t = 0;
        for (int q = 0; q < nA; ++q) t |= ((s.adj[q] >> d) & 1ull) << q;
        s.adjT[d] = t;
    }
    for (int e = PM_TID; e < nL; e += PM_NT) {
        pm_atomic_add(&s.l_in[s.le_dst[e]], 1);
        pm_atomic_add(&s.l_out[s.le_src[e]], 1);
    }
    PM_SYNC();
    for (int arr = PM_TID; arr < 4; arr += PM_NT) {        // four short serial prefixes, one thread each
        int run = 0;
        for (int q = 0; q <= nA; ++q) {
            if (arr == 0) { s.pref_q[q] = run; if (q < nA) run += PM_POPC(s.adj[q]); }
            else if (arr == 1) { s.pref_t[q] = run; if (q < nA) run += PM_POPC(s.adjT[q]); }
            else if (arr == 2) { const int c = s.l_in[q]; s.l_in[q] = run; run += c; }
            else { const int c = s.l_out[q]; s.l_out[q] = run; run += c; }
        }
    }
    PM_SYNC();
    const int og = s.og, ol = s.ol, ot = s.ot;
    const bool dst1 = A.g_dst_row != 0;
    const int* g_in = dst1 ? s.pref_t : s.pref_q;       // in-CSR keyed by the aggregation target
    const int* g_out = dst1 ? s.pref_q : s.pref_t;      // out-CSR keyed by the other end

    // ---- global graph: API list (query ascending, neighbour ascending), in-CSR by (target, source), out-CSR by id ----
    for (int idx = PM_TID; idx < nA * nA; idx += PM_NT) {
        const int q = idx / nA, n = idx - q * nA;
        if (!((s.adj[q] >> n) & 1ull)) continue;
        const int rq = PM_POPC(s.adj[q] & ((1ull << n) - 1ull));
        const int rt = PM_POPC(s.adjT[n] & ((1ull << q) - 1ull));
        const int eid = s.pref_q[q] + rq;
        const int dst = dst1 ? n : q, src = dst1 ? q : n;
        const int slot = dst1 ? s.pref_t[n] + rt : eid;
        const int u = dst1 ? eid : s.pref_t[n] + rt;
        A.eg_out[og + eid] = a0 + q;
        A.eg_out[A.Eg + og + eid] = a0 + n;
        A.g_eid[og + slot] = og + eid;
        A.g_src[og + slot] = a0 + src;
        A.g_dst[og + slot] = a0 + dst;
        A.dist_g[og + slot] = edge_len(A.pos, a0 + dst, a0 + src);
        A.g_opos[og + u] = og + slot;
    }
    for (int d = PM_TID; d <= nA; d += PM_NT) {
        A.g_ptr[a0 + d] = og + g_in[d];
        A.g_optr[a0 + d] = og + g_out[d];
        A.l_ptr[a0 + d] = ol + s.l_in[d];
        A.l_optr[a0 + d] = ol + s.l_out[d];
        if (d < nA) A.n2g[a0 + d] = m;
    }
    if (PM_TID == 0) { A.gptr[m] = a0; A.gptr[m + 1] = s.a1; }

    // ---- local graph: slot = (target, source, id) rank; out-CSR position = (source, id) rank ------------------------
    for (int e = PM_TID; e < nL; e += PM_NT) {
        const int src = s.le_src[e], dst = s.le_dst[e];
        int r_in = 0, r_out = 0;
        for (int f = 0; f < nL; ++f) {
            const int fs = s.le_src[f], fd = s.le_dst[f];
            r_in += (fd == dst) && (fs < src || (fs == src && f < e));
            r_out += (fs == src) && (f < e);
        }
        const int slot = s.l_in[dst] + r_in, u = s.l_out[src] + r_out;
        s.ls_src[slot] = (short)src;
        s.ls_dst[slot] = (short)dst;
        A.l_eid[ol + slot] = ol + e;
        A.l_src[ol + slot] = a0 + src;
        A.l_dst[ol + slot] = a0 + dst;
        A.dist_l[ol + slot] = edge_len(A.pos, a0 + dst, a0 + src);
        A.l_opos[ol + u] = ol + slot;
        if (A.el_out) {
            A.el_out[ol + e] = a0 + src;
            A.el_out[A.El + ol + e] = a0 + dst;
        }
    }
    PM_SYNC();

    // ---- triplets per local slot k = (j -> i): two-hop entries first, then one-hop (plan_tcount / plan_tfill) --------
    for (int k = PM_TID; k < nL; k += PM_NT) {
        const int j = s.ls_src[k], i = s.ls_dst[k];
        int n2 = 0, n1 = 0;
        if (A.two_hop)
            for (int p = s.l_in[j]; p < s.l_in[j + 1]; ++p) n2 += (s.ls_src[p] != i);
        for (int p = s.l_in[i]; p < s.l_in[i + 1]; ++p) n1 += (s.ls_src[p] != i);
        s.t_n2[k] = (short)n2;
        s.va[k] = n2 + n1;
        A.t_split[ol + k] = n2;
        A.t_cnt[ol + k] = n2 + n1;
    }
    PM_SYNC();
    mol_scan(s.va, nL, s.vb, s.vc, s.vx);       // vx[k] = first triplet of slot k inside the molecule
    const int nT = s.vx[nL];
    for (int k = PM_TID; k <= nL; k += PM_NT) A.t_ptr[ol + k] = ot + s.vx[k];
    for (int k = PM_TID; k < nL; k += PM_NT) {
        const int j = s.ls_src[k], i = s.ls_dst[k];
        int w = s.vx[k];
        if (A.two_hop) {
            for (int p = s.l_in[j]; p < s.l_in[j + 1]; ++p) {
                const int kk = s.ls_src[p];
                if (kk == i) continue;
                A.t_gather[ot + w] = ol + p;
                A.t_owner[ot + w] = ol + k;
                A.t_angle[ot + w] = bond_angle(A.pos, a0 + i, a0 + j, a0 + kk);
                if (w < kMolTrip) s.tg[w] = (short)p;
                ++w;
            }
        }
        for (int p = s.l_in[i]; p < s.l_in[i + 1]; ++p) {
            const int j2 = s.ls_src[p];
            if (j2 == i) continue;
            A.t_gather[ot + w] = ol + p;
            A.t_owner[ot + w] = ol + k;
            A.t_angle[ot + w] = bond_angle(A.pos, a0 + j, a0 + i, a0 + j2);
            if (w < kMolTrip) s.tg[w] = (short)p;
            ++w;
        }
    }
    PM_SYNC();

    // ---- reverse lists: triplets grouped by the slot they gather, ascending triplet id (build_buckets of plan_fill) ---
    for (int p = PM_TID; p < nL; p += PM_NT) {
        int c = 0;
        for (int w = 0; w < nT; ++w) c += (s.tg[w] == p);
        s.va[p] = c;
    }
    PM_SYNC();
    mol_scan(s.va, nL, s.vb, s.vc, s.vx);
    for (int p = PM_TID; p <= nL; p += PM_NT) A.tt_ptr[ol + p] = ot + s.vx[p];
    for (int p = PM_TID; p < nL; p += PM_NT) {
        int o = ot + s.vx[p];
        for (int w = 0; w < nT; ++w)
            if (s.tg[w] == p) A.tt_t[o++] = ot + w;
    }
}

}  // namespace pamnet
