// Skinny GEMMs (gemm.cuh) for the dim = 16 / 32 models and the 16-wide basis layers: millions of rows against a
// weight matrix of a few hundred floats.  These are HBM-bound streams -- a 64 x 64 FFMA tile wastes 3/4 of its lanes
// on N = 16 and the tensor-core tile is 128 wide -- so:
//  * NT / NN ("row kernel"): one thread per output row; the weight matrix sits in shared memory and is read by
//    broadcast, the row is read once with 128-bit loads, bias / SiLU / SiLU' / z-output epilogues as in the other paths;
//  * TN ("column kernel", weight gradients): a warp walks rows k, lane n owns output columns n, n + 32, ... for all
//    M <= 32 output rows; A[k][:] and B[k][:] are one coalesced load each, A's values reach the lanes by shuffle;
//    fp64 register accumulators, CTAs combine with fp32 atomics.
#include "gemm.cuh"

namespace pamnet {

constexpr int kSmallMaxK = 128, kSmallMaxN = 32, kSmallThreads = 128;
constexpr int kSmallTnMaxM = 32, kSmallTnMaxN = 96;

template <int NT_, int EPI>       // NT_ = register columns per thread (16 or 32)
__global__ void __launch_bounds__(kSmallThreads) gemm_rows_kernel(const GemmArgs args) {
    pdl_wait();
    pdl_trigger();
    __shared__ __align__(16) float Ws[kSmallMaxK * kSmallMaxN];      // [k][NT_]
    const GemmSlot& sl = args.slot[blockIdx.z];
    const int M = sl.m > 0 ? sl.m : args.M, N = args.N, K = args.K, mode = args.mode;
    // weights -> shared memory as [k][n]
    for (int i = threadIdx.x; i < K * NT_; i += kSmallThreads) {
        const int k = i / NT_, n = i % NT_;
        float v = 0.f;
        if (n < N) {
            if (mode == GEMM_NT) v = sl.B[(size_t)n * sl.ldb + k];
            else if (args.nseg > 0) {
                const int s = k / args.seg_len;
                v = args.seg_B[s][(size_t)(k - s * args.seg_len) * args.seg_ldb[s] + n];
            } else v = sl.B[(size_t)k * sl.ldb + n];
        }
        Ws[i] = v;
    }
    __syncthreads();
    const int m = blockIdx.x * kSmallThreads + threadIdx.x;
    if (m >= M) return;
    float acc[NT_];
#pragma unroll
    for (int n = 0; n < NT_; ++n) acc[n] = 0.f;
    const float* a = sl.A + (size_t)m * sl.lda;
    const bool vec = ((reinterpret_cast<uintptr_t>(a) & 15) == 0) && (K % 4 == 0);
    if (vec) {
        for (int k = 0; k < K; k += 4) {
            const float4 av = ld4(a + k);
            const float ak[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float* w = Ws + (k + j) * NT_;
#pragma unroll
                for (int n = 0; n < NT_; ++n) acc[n] = fmaf(ak[j], w[n], acc[n]);
            }
        }
    } else {
        for (int k = 0; k < K; ++k) {
            const float ak = a[k];
            const float* w = Ws + k * NT_;
#pragma unroll
            for (int n = 0; n < NT_; ++n) acc[n] = fmaf(ak, w[n], acc[n]);
        }
    }
    float* c = sl.C ? sl.C + (size_t)m * sl.ldc : nullptr;
    if (EPI == EPI_NONE && args.ksplit > 1) {            // accumulate into a zero-initialised / shared output
#pragma unroll
        for (int n = 0; n < NT_; ++n) if (n < N) atomicAdd(c + n, acc[n]);
        return;
    }
    float* c2 = (EPI == EPI_BIAS_SILU && sl.C2) ? sl.C2 + (size_t)m * sl.ldc : nullptr;
    const float* z = (EPI == EPI_MUL_DSILU) ? sl.Z + (size_t)m * sl.ldz : nullptr;
    // rows are 64 / 128 B: 128-bit accesses when everything is aligned (a scalar store per column touches a different
    // 32 B sector in every lane -- the first version of this kernel wrote at a quarter of the rate it read)
    const bool v4 = (N % 4 == 0) && (!c || (reinterpret_cast<uintptr_t>(c) & 15) == 0) &&
                    (!c2 || (reinterpret_cast<uintptr_t>(c2) & 15) == 0) && (!z || (reinterpret_cast<uintptr_t>(z) & 15) == 0);
#pragma unroll
    for (int n0 = 0; n0 < NT_; n0 += 4) {
        if (n0 >= N) break;
        float v[4] = {acc[n0], acc[n0 + 1], acc[n0 + 2], acc[n0 + 3]};
        float zz[4] = {0.f, 0.f, 0.f, 0.f}, old[4] = {0.f, 0.f, 0.f, 0.f};
        if (z) {
            if (v4) { const float4 t = ld4(z + n0); zz[0] = t.x; zz[1] = t.y; zz[2] = t.z; zz[3] = t.w; }
            else for (int j = 0; j < 4; ++j) if (n0 + j < N) zz[j] = z[n0 + j];
        }
        if (c && args.accumulate) {
            if (v4) { const float4 t = ld4(c + n0); old[0] = t.x; old[1] = t.y; old[2] = t.z; old[3] = t.w; }
            else for (int j = 0; j < 4; ++j) if (n0 + j < N) old[j] = c[n0 + j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (EPI == EPI_BIAS || EPI == EPI_BIAS_SILU) {
                if (sl.bias && n0 + j < N) v[j] += sl.bias[n0 + j];
            } else if (EPI == EPI_MUL_DSILU) {
                v[j] *= dsilu(zz[j]);
            }
        }
        if (c2) {
            if (v4) st4(c2 + n0, make_float4(v[0], v[1], v[2], v[3]));
            else for (int j = 0; j < 4; ++j) if (n0 + j < N) c2[n0 + j] = v[j];
        }
        if (EPI == EPI_BIAS_SILU) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = silu(v[j]);
        }
        if (c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] += old[j];
            if (v4) st4(c + n0, make_float4(v[0], v[1], v[2], v[3]));
            else for (int j = 0; j < 4; ++j) if (n0 + j < N) c[n0 + j] = v[j];
        }
    }
}

// weight gradient: C[m][n] (+)= sum_k A[k][m] B[k][n]; C2[m] += sum_k A[k][m].  Always combines with atomics: the
// caller zero-initialises C (gemm_launch does it for plain-store launches).
//   MT_ output rows (16 or 32); NJ columns per lane (n = sub-lane + W j, W = 32 / RP lanes per row);
//   RP rows walked side by side by the lane groups of a warp (2 when N <= 16, so that no lane idles).
// A lane group reads its row of A and of B once, coalesced (lane sub holds A[k][sub + W r] and B[k][sub + W j]); the MT_
// values of A[k][:] every lane multiplies with reach it by warp shuffles from the lane that loaded them.  kColsUnroll
// rows per lane group are loaded before any is multiplied: the loop is a stream of dependent-free loads.
// Accuracy: these reductions run over up to millions of rows with heavy cancellation (a 1185-way fp32 atomic
// combine missed the 1e-5 parity bar on mlp_rbf_g.weight of the RNA checkpoint).  fp32 partial sums therefore cover
// at most kColsFlush x kColsUnroll rows and are folded into fp64 accumulators; only <= 4 x 148 per-CTA results meet in
// the fp32 output.
// The fp64 accumulators live in REGISTERS and meet once per CTA (lane groups by shuffle, warps one after the other
// through shared memory).  The first version folded every fp32 window into a shared fp64 tile with atomics -- a
// compare-and-swap loop with 16 lanes per address: on the 775 k-edge RNA batch each weight-gradient launch took
// 250-400 us for 99 MB of operands, ~6 % of the HBM rate (ncu launch list profiles/r02_launches_c4.csv); and every
// lane loaded all MT_ values of A[k][:] itself (16 broadcast loads per row).
constexpr int kColsUnroll = 4, kColsFlush = 8, kColsThreads = 256;
template <int MT_, int NJ, int RP>
__global__ void __launch_bounds__(kColsThreads) gemm_cols_kernel(const GemmArgs args, int rows_per_cta) {
    pdl_wait();
    pdl_trigger();
    constexpr int W = 32 / RP, LDN = 32 * NJ;
    constexpr int AR = (MT_ + W - 1) / W;              // registers of A per lane: rows m = sub + W r
    constexpr bool DREG = MT_ * NJ <= 48;              // fp64 accumulators fit the register file (else: shared fp64 tile + atomics)
    constexpr int DR_M = DREG ? MT_ : 1, DR_N = DREG ? NJ : 1;
    __shared__ double dacc[MT_ * LDN];                 // [m][n], n < 32 NJ
    __shared__ double dsum[W * AR];
    const GemmSlot& sl = args.slot[blockIdx.z];
    const int M = sl.m > 0 ? sl.m : args.M, N = args.N, K = args.K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % W, rp = lane / W;
    if (!DREG) {
        for (int i = threadIdx.x; i < MT_ * LDN; i += kColsThreads) dacc[i] = 0.0;
        __syncthreads();
    }
    float acc[MT_][NJ];
    double dreg[DR_M][DR_N];
#pragma unroll
    for (int i = 0; i < MT_; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int i = 0; i < DR_M; ++i)
#pragma unroll
        for (int j = 0; j < DR_N; ++j) dreg[i][j] = 0.0;
    float asum[AR];                                    // column sums of A: rows m = sub + W r of this lane group's rows
    double dasum[AR];
#pragma unroll
    for (int r = 0; r < AR; ++r) { asum[r] = 0.f; dasum[r] = 0.0; }
    auto flush = [&]() {
#pragma unroll
        for (int i = 0; i < MT_; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (DREG) dreg[DREG ? i : 0][DREG ? j : 0] += (double)acc[i][j];
                else if (acc[i][j] != 0.f) atomicAdd(&dacc[i * LDN + sub + W * j], (double)acc[i][j]);
                acc[i][j] = 0.f;
            }
#pragma unroll
        for (int r = 0; r < AR; ++r) { dasum[r] += (double)asum[r]; asum[r] = 0.f; }
    };
    const int k0 = blockIdx.x * rows_per_cta, k1 = min(K, k0 + rows_per_cta);
    constexpr int kStep = (kColsThreads / 32) * RP;    // rows between two consecutive rows of one lane group
    int it = 0;
    // (warp-uniform trip count: the shuffles below need every lane of the warp in every iteration)
    for (int kw = k0 + warp * RP; kw < k1; kw += kStep * kColsUnroll) {
        const int kb = kw + rp;
        float bv[kColsUnroll][NJ], am[kColsUnroll][AR];
#pragma unroll
        for (int u = 0; u < kColsUnroll; ++u) {
            const int k = kb + u * kStep;
            const bool live = k < k1;
            const float* a = sl.A + (size_t)(live ? k : k0) * sl.lda;
            const float* b = sl.B + (size_t)(live ? k : k0) * sl.ldb;
#pragma unroll
            for (int j = 0; j < NJ; ++j) bv[u][j] = (live && sub + W * j < N) ? b[sub + W * j] : 0.f;
#pragma unroll
            for (int r = 0; r < AR; ++r) am[u][r] = (live && sub + W * r < M) ? a[sub + W * r] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kColsUnroll; ++u) {
#pragma unroll
            for (int r = 0; r < AR; ++r) asum[r] += am[u][r];
#pragma unroll
            for (int i = 0; i < MT_; ++i) {
                // A[k][i] of this lane group's row: loaded by its lane i % W (register i / W)
                const float ai = __shfl_sync(0xffffffffu, am[u][i / W], i % W, W);
#pragma unroll
                for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(ai, bv[u][j], acc[i][j]);
            }
        }
        if (++it == kColsFlush) { flush(); it = 0; }
    }
    flush();
    // lane groups of a warp -> lane group 0
#pragma unroll
    for (int o = W; o < 32; o <<= 1) {
#pragma unroll
        for (int i = 0; i < DR_M; ++i)
#pragma unroll
            for (int j = 0; j < DR_N; ++j) dreg[i][j] += __shfl_xor_sync(0xffffffffu, dreg[i][j], o);
#pragma unroll
        for (int r = 0; r < AR; ++r) dasum[r] += __shfl_xor_sync(0xffffffffu, dasum[r], o);
    }
    // warps one after the other: plain read-modify-write, every lane of group 0 owns its addresses
    for (int w = 0; w < kColsThreads / 32; ++w) {
        if (warp == w && rp == 0) {
            if (DREG) {
#pragma unroll
                for (int i = 0; i < DR_M; ++i)
#pragma unroll
                    for (int j = 0; j < DR_N; ++j) {
                        double* d = &dacc[i * LDN + sub + W * j];
                        *d = (w == 0 ? 0.0 : *d) + dreg[i][j];
                    }
            }
#pragma unroll
            for (int r = 0; r < AR; ++r) {
                double* d = &dsum[sub + W * r];
                *d = (w == 0 ? 0.0 : *d) + dasum[r];
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < M * N; i += kColsThreads) {
        const int m = i / N, n = i % N;
        const float v = (float)dacc[m * LDN + n];
        if (v != 0.f) atomicAdd(&sl.C[(size_t)m * sl.ldc + n], v);
    }
    if (sl.C2 && threadIdx.x < M) {
        const float v = (float)dsum[threadIdx.x];
        if (v != 0.f) atomicAdd(&sl.C2[threadIdx.x], v);
    }
}

bool gemm_small_eligible(const GemmArgs& a) {
    if (a.mode == GEMM_TN) return a.M <= kSmallTnMaxM && a.N <= kSmallTnMaxN && !a.accumulate && a.epi == EPI_NONE;
    if (a.N > kSmallMaxN || a.K > kSmallMaxK) return false;
    if (a.ksplit > 1 && a.epi != EPI_NONE) return false;
    return true;
}

int gemm_small_launch(const GemmArgs& a, cudaStream_t st) {
    if (a.mode == GEMM_TN) {
        if (a.ksplit <= 1) {         // plain-store semantics: zero the outputs, then accumulate
            for (int i = 0; i < a.nslots; ++i) {
                const int M = a.slot[i].m > 0 ? a.slot[i].m : a.M;
                PAMNET_CUDA(cudaMemset2DAsync(a.slot[i].C, sizeof(float) * (size_t)a.slot[i].ldc, 0, sizeof(float) * a.N,
                                              (size_t)M, st));
            }
        }
        // at most ~4 CTAs per SM over all slots (few fp32 partials meet in the output), at least 256 rows each
        int ctas = ceil_div(a.K, 256);
        const int cap = 4 * kNumSM / (a.nslots > 0 ? a.nslots : 1) + 1;
        if (ctas > cap) ctas = cap;
        const int rows = ceil_div(a.K, ctas);
        dim3 grid(ceil_div(a.K, rows), 1, a.nslots);
#define COLS_CASE(MT_) \
        do { \
            if (a.N <= 16) PAMNET_CUDA(launch_pdl(gemm_cols_kernel<MT_, 1, 2>, grid, dim3(kColsThreads), 0, st, a, rows)); \
            else if (a.N <= 32) PAMNET_CUDA(launch_pdl(gemm_cols_kernel<MT_, 1, 1>, grid, dim3(kColsThreads), 0, st, a, rows)); \
            else PAMNET_CUDA(launch_pdl(gemm_cols_kernel<MT_, kSmallTnMaxN / 32, 1>, grid, dim3(kColsThreads), 0, st, a, rows)); \
        } while (0)
        if (a.M <= 16) COLS_CASE(16); else COLS_CASE(32);
#undef COLS_CASE
        return 0;
    }
    dim3 grid(ceil_div(a.M, kSmallThreads), 1, a.nslots);
#define ROWS_CASE(NT_, EPI_) PAMNET_CUDA(launch_pdl(gemm_rows_kernel<NT_, EPI_>, grid, dim3(kSmallThreads), 0, st, a)); break
#define ROWS_EPI(NT_)                                                   \
    switch (a.epi) {                                                    \
        case EPI_NONE: ROWS_CASE(NT_, EPI_NONE);                        \
        case EPI_BIAS: ROWS_CASE(NT_, EPI_BIAS);                        \
        case EPI_BIAS_SILU: ROWS_CASE(NT_, EPI_BIAS_SILU);              \
        default: ROWS_CASE(NT_, EPI_MUL_DSILU);                         \
    }
    if (a.N <= 16) { ROWS_EPI(16) } else { ROWS_EPI(32) }
#undef ROWS_EPI
#undef ROWS_CASE
    return 0;
}

}  // namespace pamnet
