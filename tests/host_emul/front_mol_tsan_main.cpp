// Stand-alone runner of the multi-threaded host build of the per-molecule front end, for ThreadSanitizer:
//   g++ -O1 -g -ffp-contract=off -std=c++17 -DPM_HOST_THREADS=8 -fsanitize=thread -pthread front_mol_tsan_main.cpp -o runner
//   runner in.bin out.bin
// in.bin : int64 {n_nodes, n_graphs, n_edges_in, max_nb, g_dst_row, two_hop}, float r2, pos, batch, edge_index
// out.bin: int64 counts[8], then (when no flag was raised) eg [2,Eg] int64, el [2,El] int64, the 21 int32 plan arrays in
//          the order of front_mol_host.cpp, t_angle, dist_g, dist_l
#include <cstdio>
#include <cstdlib>

#include "front_mol_host.cpp"

template <class T>
static std::vector<T> rd(FILE* f, size_t n) {
    std::vector<T> v(n ? n : 1);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    auto h = rd<int64_t>(f, 6);
    auto r2 = rd<float>(f, 1);
    const int64_t n = h[0], g = h[1], e = h[2];
    auto pos = rd<float>(f, 3 * n);
    auto batch = rd<int64_t>(f, n);
    auto ei = rd<int64_t>(f, 2 * e);
    fclose(f);
    std::vector<int32_t> mc(4 * g + 1);
    std::vector<unsigned long long> counts(8, 0);
    front_mol_host(0, pos.data(), batch.data(), n, g, ei.data(), e, r2[0], (int)h[3], (int)h[4], (int)h[5], mc.data(),
                   counts.data(), 0, 0, nullptr, nullptr, nullptr, nullptr);
    FILE* o = fopen(argv[2], "wb");
    fwrite(counts.data(), 8, 8, o);
    if (counts[5] == 0 && (int64_t)counts[4] == e) {
        const int64_t Eg = counts[0], El = counts[1], T = counts[2] + counts[3];
        const int64_t sz[21] = {n, g + 1, n + 1, Eg, Eg, Eg, n + 1, Eg, n + 1, El, El, El, n + 1, El, El, El, El + 1, El + 1, T, T, T};
        const int64_t fsz[3] = {T, Eg, El};
        std::vector<std::vector<int32_t>> ia;
        std::vector<std::vector<float>> fa;
        std::vector<int32_t*> ip;
        std::vector<float*> fp;
        for (int i = 0; i < 21; ++i) { ia.emplace_back(sz[i] + 1, -7); }
        for (int i = 0; i < 3; ++i) { fa.emplace_back(fsz[i] + 1, 0.f); }
        for (auto& v : ia) ip.push_back(v.data());
        for (auto& v : fa) fp.push_back(v.data());
        std::vector<int64_t> eg(2 * Eg + 1, -7), el(2 * El + 1, -7);
        front_mol_host(1, pos.data(), batch.data(), n, g, ei.data(), e, r2[0], (int)h[3], (int)h[4], (int)h[5], mc.data(),
                       counts.data(), Eg, El, eg.data(), El == e ? nullptr : el.data(), ip.data(), fp.data());
        if (El == e) for (int64_t i = 0; i < 2 * e; ++i) el[i] = ei[i];
        fwrite(eg.data(), 8, 2 * Eg, o);
        fwrite(el.data(), 8, 2 * El, o);
        for (int i = 0; i < 21; ++i) fwrite(ia[i].data(), 4, sz[i], o);
        for (int i = 0; i < 3; ++i) fwrite(fa[i].data(), 4, fsz[i], o);
    }
    fclose(o);
    return 0;
}
