// Skinny GEMMs (gemm.cuh) for the dim = 16 / 32 models and the 16-wide basis layers: millions of rows against a
// weight matrix of a few hundred floats.  These are HBM-bound streams -- a 64 x 64 FFMA tile wastes 3/4 of its lanes
// on N = 16 and the tensor-core tile is 128 wide -- so:
//  * NT / NN ("row kernel"): one thread per output row; the weight matrix sits in shared memory and is read by
//    broadcast, the row is read once with 128-bit loads, bias / SiLU / SiLU' / z-output epilogues as in the other paths;
//  * TN ("column kernel", weight gradients): a warp walks rows k, lane n owns output columns n, n + 32, ... for all
//    M <= 32 output rows; A[k][:] and B[k][:] are one coalesced load each, A's values reach the lanes by shuffle;
//    fp64 register accumulators, CTAs combine with fp32 atomics.
#include "gemm.cuh"

#include <stdio.h>
#include <stdlib.h>

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

// ---- weight gradient with M, N <= 16 (the dim = 16 models) on the tensor cores ------------------------------------------
// C[16 x 16] (+)= A^T B over K rows, K up to millions: per 8 rows a warp issues 12 mma.sync.m16n8k8 (a three-part tf32 split of both
// operands, six products, for two 8-column tiles) instead of 16 shuffles + 16 FMAs per ROW PAIR in the kernel above (which is issue-bound:
// ncu 50 % issue slots at 1.9 TB/s).  Fragments are loaded straight from global memory: lane (g, tq) reads
// A[k0 + tq (+4)][g (+8)] and B[k0 + tq (+4)][8 j + g] -- every request covers whole 32-byte sectors of 8 consecutive rows.
// Accuracy: fp32 over the 32 rows a warp has in flight, then fp64 registers; warps meet once per CTA in shared
// memory, CTAs with one fp32 atomic per output.  Column sums of A (bias gradient) ride along on the A fragments.
__device__ __forceinline__ void cols_mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// x = hi + mid + lo EXACTLY, each part a tf32 number (11 + 11 + <= 3 significant bits): hi = rn_tf32(x), mid =
// rn_tf32(x - hi), lo = the rest.  With all six products above 2^-33 |a b| the result is a true fp32 product sum.  (The
// two-part split of the big GEMMs leaves 2^-21 |a b| per term -- the tensor core truncates the low part -- which on
// cancellation-heavy node-level gradients of the RNA model came out at 1.03e-5 of the tensor's maximum, 3x the error
// of an fp32 FFMA sum and over the 1e-5 parity bar.)
__device__ __forceinline__ uint32_t rn_tf32_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void cols_split(float x, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
    hi = rn_tf32_bits(x);
    const float r1 = x - __uint_as_float(hi);
    mid = rn_tf32_bits(r1);
    lo = __float_as_uint(r1 - __uint_as_float(mid));
}
constexpr int kColsMmaSteps = 4;                                   // k-steps (8 rows each) in flight per warp
constexpr int kColsMmaMinK = 32768;                                // below: the SIMT kernel (time does not matter there)
__global__ void __launch_bounds__(kColsThreads) gemm_cols_mma_kernel(const GemmArgs args, int rows_per_cta) {
    pdl_wait();
    pdl_trigger();
    constexpr int NW = kColsThreads / 32;
    __shared__ double dacc[NW][16 * 16];
    __shared__ double dsum[NW][16];
    const GemmSlot& sl = args.slot[blockIdx.z];
    const int M = sl.m > 0 ? sl.m : args.M, N = args.N, K = args.K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int k0 = blockIdx.x * rows_per_cta, k1 = min(K, k0 + rows_per_cta);
    const bool m_lo = g < M, m_hi = g + 8 < M, n_0 = g < N, n_1 = g + 8 < N;
    const float* A = sl.A; const float* B = sl.B;
    const size_t lda = (size_t)sl.lda, ldb = (size_t)sl.ldb;
    double dc[2][4], ds[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) dc[j][i] = 0.0;
    for (int kb = k0 + warp * 8 * kColsMmaSteps; kb < k1; kb += NW * 8 * kColsMmaSteps) {
        float av[kColsMmaSteps][4], bv[kColsMmaSteps][4];
#pragma unroll
        for (int s = 0; s < kColsMmaSteps; ++s) {
            const int r0 = kb + 8 * s + tq, r1 = r0 + 4;
            const bool l0 = r0 < k1, l1 = r1 < k1;
            const float* a0 = A + (size_t)(l0 ? r0 : k0) * lda;
            const float* a1 = A + (size_t)(l1 ? r1 : k0) * lda;
            const float* b0 = B + (size_t)(l0 ? r0 : k0) * ldb;
            const float* b1 = B + (size_t)(l1 ? r1 : k0) * ldb;
            av[s][0] = (l0 && m_lo) ? a0[g] : 0.f;
            av[s][1] = (l0 && m_hi) ? a0[g + 8] : 0.f;
            av[s][2] = (l1 && m_lo) ? a1[g] : 0.f;
            av[s][3] = (l1 && m_hi) ? a1[g + 8] : 0.f;
            bv[s][0] = (l0 && n_0) ? b0[g] : 0.f;          // tile 0: b0 (k = tq), b1 (k = tq + 4)
            bv[s][1] = (l1 && n_0) ? b1[g] : 0.f;
            bv[s][2] = (l0 && n_1) ? b0[g + 8] : 0.f;      // tile 1
            bv[s][3] = (l1 && n_1) ? b1[g + 8] : 0.f;
        }
        // fp32 over the 32 rows in flight (the window of the SIMT kernel), then fp64
        float c[2][4], cx[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) { c[j][i] = 0.f; cx[j][i] = 0.f; }
        float as0 = 0.f, as1 = 0.f;
#pragma unroll
        for (int s = 0; s < kColsMmaSteps; ++s) {
            uint32_t ah[4], am[4], al[4], bh[4], bm[4], bl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { cols_split(av[s][i], ah[i], am[i], al[i]); cols_split(bv[s][i], bh[i], bm[i], bl[i]); }
            as0 += av[s][0] + av[s][2];
            as1 += av[s][1] + av[s][3];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t bhj[2] = {bh[2 * j], bh[2 * j + 1]}, bmj[2] = {bm[2 * j], bm[2 * j + 1]},
                               blj[2] = {bl[2 * j], bl[2 * j + 1]};
                cols_mma_tf32(cx[j], al, bhj);             // small terms in their own accumulator
                cols_mma_tf32(cx[j], ah, blj);
                cols_mma_tf32(cx[j], am, bmj);
                cols_mma_tf32(cx[j], am, bhj);
                cols_mma_tf32(cx[j], ah, bmj);
                cols_mma_tf32(c[j], ah, bhj);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) dc[j][i] += (double)c[j][i] + (double)cx[j][i];
        ds[0] += (double)as0;
        ds[1] += (double)as1;
    }
    // column sums: fold the four k-lanes of a row group
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        ds[0] += __shfl_xor_sync(0xffffffffu, ds[0], o);
        ds[1] += __shfl_xor_sync(0xffffffffu, ds[1], o);
    }
    // C fragment: c0 (g, 2 tq), c1 (g, 2 tq + 1), c2 (g + 8, 2 tq), c3 (g + 8, 2 tq + 1) of tile j (columns 8 j + ...)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        dacc[warp][g * 16 + 8 * j + 2 * tq] = dc[j][0];
        dacc[warp][g * 16 + 8 * j + 2 * tq + 1] = dc[j][1];
        dacc[warp][(g + 8) * 16 + 8 * j + 2 * tq] = dc[j][2];
        dacc[warp][(g + 8) * 16 + 8 * j + 2 * tq + 1] = dc[j][3];
    }
    if (tq == 0) { dsum[warp][g] = ds[0]; dsum[warp][g + 8] = ds[1]; }
    __syncthreads();
    {
        const int i = threadIdx.x;                         // 256 threads = 16 x 16 outputs
        const int m = i / 16, n = i % 16;
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) tot += dacc[w][i];
        const float v = (float)tot;
        if (m < M && n < N && v != 0.f) atomicAdd(&sl.C[(size_t)m * sl.ldc + n], v);
        if (sl.C2 && i < M) {
            double ts = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) ts += dsum[w][i];
            const float vs = (float)ts;
            if (vs != 0.f) atomicAdd(&sl.C2[i], vs);
        }
    }
}

// ---- rows kernel with N = 16, K = 16 or 32 (the dim = 16 models) on the tensor cores -----------------------------------
// C[m][0..15] = epi(sum_k A[m][k] W(k, n)): a warp owns 16-row tiles; per tile and 16 k it issues 24 mma.sync.m16n8k8 (the
// exact three-part tf32 split above, six products, 2 k-steps x 2 column tiles) -- the SIMT kernel spends 256 FMAs and 64
// shared-memory loads per ROW and stalls on the shared-memory pipe (ncu: mio_throttle / short_scoreboard, 20-47 % issue
// slots, 1.5-2.7 TB/s).  The sum over k and the order of the output columns are free, so both are PERMUTED to make every
// global access a 128-bit one: lane (g, tq) reads A[m0 + g (+8)][16 c + 4 tq .. + 3] and feeds k-step s with elements
// 2 s (slot tq) and 2 s + 1 (slot tq + 4); MMA column (tile j, c) stands for output column 4 (c / 2) + 2 j + c % 2, which
// puts C[m][4 tq .. 4 tq + 3] of rows g and g + 8 into the lane's accumulators.  The weight fragments (all three parts)
// follow the same maps and live in registers for the whole kernel.
constexpr int kRowsMmaThreads = 128;
template <int EPI, int KC>                                         // KC = K / 16
__global__ void __launch_bounds__(kRowsMmaThreads) gemm_rows_mma_kernel(const GemmArgs args) {
    constexpr int kRowsMmaTiles = 4 / KC;                          // row tiles in flight per warp (same registers for every KC)
    pdl_wait();
    pdl_trigger();
    const GemmSlot& sl = args.slot[blockIdx.z];
    const int M = sl.m > 0 ? sl.m : args.M;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, tq = lane & 3;
    // weight element (k, n)
    auto wld = [&](int k, int n) -> float {
        if (args.mode == GEMM_NT) return sl.B[(size_t)n * sl.ldb + k];
        if (args.nseg > 0) {
            const int sg = k / args.seg_len;
            return args.seg_B[sg][(size_t)(k - sg * args.seg_len) * args.seg_ldb[sg] + n];
        }
        return sl.B[(size_t)k * sl.ldb + n];
    };
    // B fragments: [chunk][k-step][tile][part][2]
    uint32_t wf[KC][2][2][3][2];
#pragma unroll
    for (int c = 0; c < KC; ++c)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = 4 * (g / 2) + 2 * j + (g & 1);
#pragma unroll
                for (int e = 0; e < 2; ++e)          // b0: slot tq -> element 2 s; b1: slot tq + 4 -> element 2 s + 1
                    cols_split(wld(16 * c + 4 * tq + 2 * s + e, n), wf[c][s][j][0][e], wf[c][s][j][1][e], wf[c][s][j][2][e]);
            }
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((EPI == EPI_BIAS || EPI == EPI_BIAS_SILU) && sl.bias) bias4 = ld4(sl.bias + 4 * tq);
    const int tile0 = (blockIdx.x * (kRowsMmaThreads / 32) + warp) * kRowsMmaTiles;
    float4 av[kRowsMmaTiles][KC][2];
#pragma unroll
    for (int t = 0; t < kRowsMmaTiles; ++t) {
        const int m0 = (tile0 + t) * 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = m0 + g + 8 * h;
#pragma unroll
            for (int c = 0; c < KC; ++c)
                av[t][c][h] = m < M ? ld4(sl.A + (size_t)m * sl.lda + 16 * c + 4 * tq) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int t = 0; t < kRowsMmaTiles; ++t) {
        const int m0 = (tile0 + t) * 16;
        if (m0 >= M) break;
        float acc[2][4], accx[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[j][i] = 0.f; accx[j][i] = 0.f; }
#pragma unroll
        for (int c = 0; c < KC; ++c) {
            const float lo_[4] = {av[t][c][0].x, av[t][c][0].y, av[t][c][0].z, av[t][c][0].w};      // row g
            const float hi_[4] = {av[t][c][1].x, av[t][c][1].y, av[t][c][1].z, av[t][c][1].w};      // row g + 8
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                // a0 (g, slot tq), a1 (g + 8, slot tq), a2 (g, slot tq + 4), a3 (g + 8, slot tq + 4)
                const float a_[4] = {lo_[2 * s], hi_[2 * s], lo_[2 * s + 1], hi_[2 * s + 1]};
                uint32_t ah[4], am[4], al[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) cols_split(a_[i], ah[i], am[i], al[i]);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    cols_mma_tf32(accx[j], al, wf[c][s][j][0]);
                    cols_mma_tf32(accx[j], ah, wf[c][s][j][2]);
                    cols_mma_tf32(accx[j], am, wf[c][s][j][1]);
                    cols_mma_tf32(accx[j], am, wf[c][s][j][0]);
                    cols_mma_tf32(accx[j], ah, wf[c][s][j][1]);
                    cols_mma_tf32(acc[j], ah, wf[c][s][j][0]);
                }
            }
        }
        // lane's outputs: rows g (h = 0) and g + 8 (h = 1), columns 4 tq + {0, 1} (tile 0) and + {2, 3} (tile 1)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = m0 + g + 8 * h;
            if (m >= M) continue;
            float v[4] = {acc[0][2 * h] + accx[0][2 * h], acc[0][2 * h + 1] + accx[0][2 * h + 1],
                          acc[1][2 * h] + accx[1][2 * h], acc[1][2 * h + 1] + accx[1][2 * h + 1]};
            float* c = sl.C ? sl.C + (size_t)m * sl.ldc + 4 * tq : nullptr;
            if (EPI == EPI_BIAS || EPI == EPI_BIAS_SILU) {
                v[0] += bias4.x; v[1] += bias4.y; v[2] += bias4.z; v[3] += bias4.w;
                if (EPI == EPI_BIAS_SILU) {
                    if (sl.C2) st4(sl.C2 + (size_t)m * sl.ldc + 4 * tq, make_float4(v[0], v[1], v[2], v[3]));
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = silu(v[i]);
                }
            } else if (EPI == EPI_MUL_DSILU) {
                const float4 z = ld4(sl.Z + (size_t)m * sl.ldz + 4 * tq);
                v[0] *= dsilu(z.x); v[1] *= dsilu(z.y); v[2] *= dsilu(z.z); v[3] *= dsilu(z.w);
            }
            if (!c) continue;
            if (EPI == EPI_NONE && args.ksplit > 1) {      // several slots / launches add into one output (gemm.cuh)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(c), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
                continue;
            }
            if (args.accumulate) {
                const float4 o = ld4(c);
                v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
            }
            st4(c, make_float4(v[0], v[1], v[2], v[3]));
        }
    }
}

static bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// N = 16 exactly, K = 16 / 32 / 64, everything 16-byte aligned, enough rows to matter
static bool rows_mma_ok(const GemmArgs& a) {
    static int dbg = -1;
    if (dbg < 0) dbg = getenv("PAMNET_DEBUG_ROWS") ? 1 : 0;
#define RM_NO(why) do { if (dbg && a.M >= 4096) fprintf(stderr, "rows_mma: no (%s) mode %d M %d N %d K %d epi %d ksplit %d nseg %d seg_len %d acc %d\n", why, a.mode, a.M, a.N, a.K, a.epi, a.ksplit, a.nseg, a.seg_len, a.accumulate); return false; } while (0)
    if (a.mode == GEMM_TN) return false;
    if (a.N != 16 || (a.K != 16 && a.K != 32 && a.K != 64) || (a.ksplit > 1 && a.epi != EPI_NONE) || a.M < 4096) RM_NO("shape");
    if (a.nseg > 0 && (a.seg_len % 16 != 0 || a.mode != GEMM_NN)) RM_NO("segments");
    for (int i = 0; i < a.nslots; ++i) {
        const GemmSlot& s = a.slot[i];
        if (!al16p(s.A) || s.lda % 4 != 0) RM_NO("A alignment");
        if (s.C && (!al16p(s.C) || s.ldc % 4 != 0)) RM_NO("C alignment");
        if (a.epi == EPI_BIAS_SILU && s.C2 && !al16p(s.C2)) RM_NO("C2 alignment");
        if (a.epi == EPI_MUL_DSILU && (!s.Z || !al16p(s.Z) || s.ldz % 4 != 0)) RM_NO("Z");
        if ((a.epi == EPI_BIAS || a.epi == EPI_BIAS_SILU) && s.bias && !al16p(s.bias)) RM_NO("bias alignment");
    }
#undef RM_NO
    return true;
}

// PAMNET_COLS=ffma: the SIMT column kernel for every shape (A/B switch)
static bool cols_mma_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("PAMNET_COLS"); on = (e && e[0] == 'f') ? 0 : 1; }
    return on == 1;
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
        if (a.M <= 16 && a.N <= 16 && a.K >= kColsMmaMinK && cols_mma_enabled()) {
            PAMNET_CUDA(launch_pdl(gemm_cols_mma_kernel, grid, dim3(kColsThreads), 0, st, a, rows));
            return 0;
        }
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
    if (cols_mma_enabled() && rows_mma_ok(a)) {
        const int kc = a.K / 16;
        dim3 grid(ceil_div(a.M, 16 * (kRowsMmaThreads / 32) * (4 / kc)), 1, a.nslots);
#define RM_CASE(EPI_) \
        do { \
            if (kc == 1) PAMNET_CUDA(launch_pdl(gemm_rows_mma_kernel<EPI_, 1>, grid, dim3(kRowsMmaThreads), 0, st, a)); \
            else if (kc == 2) PAMNET_CUDA(launch_pdl(gemm_rows_mma_kernel<EPI_, 2>, grid, dim3(kRowsMmaThreads), 0, st, a)); \
            else PAMNET_CUDA(launch_pdl(gemm_rows_mma_kernel<EPI_, 4>, grid, dim3(kRowsMmaThreads), 0, st, a)); \
        } while (0)
        switch (a.epi) {
            case EPI_NONE: RM_CASE(EPI_NONE); break;
            case EPI_BIAS: RM_CASE(EPI_BIAS); break;
            case EPI_BIAS_SILU: RM_CASE(EPI_BIAS_SILU); break;
            default: RM_CASE(EPI_MUL_DSILU); break;
        }
#undef RM_CASE
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
