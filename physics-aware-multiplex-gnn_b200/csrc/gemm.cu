// fp32 FFMA GEMM: BM x BN x 16 tiles, 256 threads, TM x TN register micro-tile (128x128 / 8x8 for the large
// edge/triplet contractions, 64x64 / 4x4 for small or skinny ones), register-prefetch double buffering.
// Results are plain IEEE fp32 sums (no tensor-core rounding), which is what holds the 1e-5 parity bar of the
// reference's nn.Linear layers (layers/basic.py:19-22).
#include "gemm.cuh"

#include <stdlib.h>

namespace pamnet {

constexpr int BK = 16, GT = 256;
constexpr int PAD = 4;

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Load 4 consecutive elements along the contiguous axis with range guards.
__device__ __forceinline__ float4 load4_guard(const float* __restrict__ p, int valid, bool vec_ok) {
    if (valid >= 4 && vec_ok) return ld4(p);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid > 0) r.x = p[0];
    if (valid > 1) r.y = p[1];
    if (valid > 2) r.z = p[2];
    if (valid > 3) r.w = p[3];
    return r;
}

// Micro-tile rows: TM/4 groups of 4 consecutive rows, group g at offset g * (BM / (TM/4)) + ty*4 (same for
// columns): keeps every shared-memory read a conflict-free / broadcast 128-bit access.
template <int BM, int BN, int TM, int TN, int EPI>
__global__ void __launch_bounds__(GT, (TM >= 8 ? 2 : 3)) gemm_kernel(const GemmArgs args) {
    static_assert((BM / TM) * (BN / TN) == GT, "256 threads");
    constexpr int GM = TM / 4, GN = TN / 4;          // 4-wide groups per thread
    constexpr int SM_ = BM / GM, SN_ = BN / GN;      // group stride
    constexpr int LA = BM / 64, LB = BN / 64;        // float4 loads per thread per k-tile
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const GemmSlot& sl = args.slot[blockIdx.z];
    const int M = sl.m > 0 ? sl.m : args.M, N = args.N, K = args.K, mode = args.mode;
    const int tiles_n = (N + BN - 1) / BN;
    const int m0 = (blockIdx.x / tiles_n) * BM, n0 = (blockIdx.x % tiles_n) * BN;
    const int t = threadIdx.x;
    if (m0 >= M) return;

    int k_begin = 0, k_end = K;
    if (args.ksplit > 1) {
        const int chunk = ((K + args.ksplit - 1) / args.ksplit + BK - 1) / BK * BK;
        k_begin = blockIdx.y * chunk;
        k_end = min(K, k_begin + chunk);
        if (k_begin >= k_end) return;
    }

    const bool a_kcontig = (mode != GEMM_TN);
    const bool b_kcontig = (mode == GEMM_NT);
    const bool a_vec = aligned16(sl.A) && (sl.lda % 4 == 0);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[LA], rb[LB];
    // k-contiguous operand: 64 rows x (4 float4 along k) per pass; m/n-contiguous: 16 k x (16 float4) per pass
    const int lk_row = t >> 2, lk_k = (t & 3) * 4;
    const int lm_k = t >> 4, lm_m = (t & 15) * 4;

    auto fetch = [&](int k0) {
#pragma unroll
        for (int p = 0; p < LA; ++p) {
            if (a_kcontig) {
                const int m = m0 + lk_row + 64 * p, k = k0 + lk_k;
                ra[p] = (m < M) ? load4_guard(sl.A + (size_t)m * sl.lda + k, k_end - k, a_vec) : make_float4(0, 0, 0, 0);
            } else {
                const int k = k0 + lm_k, m = m0 + lm_m + 64 * p;
                ra[p] = (k < k_end) ? load4_guard(sl.A + (size_t)k * sl.lda + m, M - m, a_vec) : make_float4(0, 0, 0, 0);
            }
        }
        const float* Bp = sl.B;
        int ldb = sl.ldb, kb = k0;
        if (args.nseg > 0) {
            const int s = k0 / args.seg_len;
            Bp = args.seg_B[s];
            ldb = args.seg_ldb[s];
            kb = k0 - s * args.seg_len;
        }
        const bool b_vec = aligned16(Bp) && (ldb % 4 == 0);
#pragma unroll
        for (int p = 0; p < LB; ++p) {
            if (b_kcontig) {
                const int n = n0 + lk_row + 64 * p, k = kb + lk_k;
                rb[p] = (n < N) ? load4_guard(Bp + (size_t)n * ldb + k, k_end - (k0 + lk_k), b_vec) : make_float4(0, 0, 0, 0);
            } else {
                const int k = kb + lm_k, n = n0 + lm_m + 64 * p;
                rb[p] = (k0 + lm_k < k_end) ? load4_guard(Bp + (size_t)k * ldb + n, N - n, b_vec) : make_float4(0, 0, 0, 0);
            }
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int p = 0; p < LA; ++p) {
            if (a_kcontig) {
                const int r = lk_row + 64 * p;
                As[buf][lk_k + 0][r] = ra[p].x; As[buf][lk_k + 1][r] = ra[p].y;
                As[buf][lk_k + 2][r] = ra[p].z; As[buf][lk_k + 3][r] = ra[p].w;
            } else {
                *reinterpret_cast<float4*>(&As[buf][lm_k][lm_m + 64 * p]) = ra[p];
            }
        }
#pragma unroll
        for (int p = 0; p < LB; ++p) {
            if (b_kcontig) {
                const int r = lk_row + 64 * p;
                Bs[buf][lk_k + 0][r] = rb[p].x; Bs[buf][lk_k + 1][r] = rb[p].y;
                Bs[buf][lk_k + 2][r] = rb[p].z; Bs[buf][lk_k + 3][r] = rb[p].w;
            } else {
                *reinterpret_cast<float4*>(&Bs[buf][lm_k][lm_m + 64 * p]) = rb[p];
            }
        }
    };

    const int ty = t >> 4, tx = t & 15;
    // weight-gradient mode: column sums of A (= bias gradient) ride along on the first column tile
    const bool do_bias = (mode == GEMM_TN) && sl.C2 != nullptr && n0 == 0 && tx == 0;
    float bsum[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) bsum[i] = 0.f;

    fetch(k_begin);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = (k0 + BK < k_end);
        if (more) fetch(k0 + BK);
        // the 8x8 body is 68 instructions per k: unrolling all 16 would be ~17 KB of code and thrash the
        // instruction cache (ncu: "no_instructions" was the top stall), so the big tile unrolls 4
#pragma unroll(TM >= 8 ? 4 : 16)
        for (int kk = 0; kk < BK; ++kk) {
            float av[TM], bv[TN];
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][g * SM_ + ty * 4]);
                av[g * 4 + 0] = a.x; av[g * 4 + 1] = a.y; av[g * 4 + 2] = a.z; av[g * 4 + 3] = a.w;
            }
#pragma unroll
            for (int g = 0; g < GN; ++g) {
                const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][g * SN_ + tx * 4]);
                bv[g * 4 + 0] = b.x; bv[g * 4 + 1] = b.y; bv[g * 4 + 2] = b.z; bv[g * 4 + 3] = b.w;
            }
            if (do_bias) {
#pragma unroll
                for (int i = 0; i < TM; ++i) bsum[i] += av[i];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    // epilogue: each thread owns groups of 4 consecutive columns -> 128-bit stores when the row is aligned
    const bool c_vec = aligned16(sl.C) && (sl.ldc % 4 == 0) && (sl.C2 == nullptr || mode == GEMM_TN || aligned16(sl.C2)) &&
                       (EPI != EPI_MUL_DSILU || (aligned16(sl.Z) && sl.ldz % 4 == 0));
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / 4) * SM_ + ty * 4 + (i & 3);
        if (m >= M) continue;
        if (do_bias) {
            if (args.ksplit > 1) atomicAdd(&sl.C2[m], bsum[i]); else sl.C2[m] = bsum[i];
        }
#pragma unroll
        for (int gj = 0; gj < GN; ++gj) {
            const int n = n0 + gj * SN_ + tx * 4;
            if (n >= N) continue;
            float v[4] = {acc[i][gj * 4 + 0], acc[i][gj * 4 + 1], acc[i][gj * 4 + 2], acc[i][gj * 4 + 3]};
            const size_t ci = (size_t)m * sl.ldc + n;
            const int nv = min(4, N - n);
            if (EPI == EPI_NONE && args.ksplit > 1) {
                for (int j = 0; j < nv; ++j) atomicAdd(&sl.C[ci + j], v[j]);
                continue;
            }
            const bool vec = c_vec && nv == 4;
            if (EPI == EPI_BIAS || EPI == EPI_BIAS_SILU) {
                if (sl.bias)
                    for (int j = 0; j < nv; ++j) v[j] += sl.bias[n + j];
                if (EPI == EPI_BIAS_SILU) {
                    if (sl.C2) {   // pre-activation (NT mode only)
                        if (vec) st4(sl.C2 + ci, make_float4(v[0], v[1], v[2], v[3]));
                        else for (int j = 0; j < nv; ++j) sl.C2[ci + j] = v[j];
                    }
                    for (int j = 0; j < nv; ++j) v[j] = silu(v[j]);
                }
            } else if (EPI == EPI_MUL_DSILU) {
                const size_t zi = (size_t)m * sl.ldz + n;
                if (vec) {
                    const float4 z = ld4(sl.Z + zi);
                    v[0] *= dsilu(z.x); v[1] *= dsilu(z.y); v[2] *= dsilu(z.z); v[3] *= dsilu(z.w);
                } else {
                    for (int j = 0; j < nv; ++j) v[j] *= dsilu(sl.Z[zi + j]);
                }
            }
            if (!sl.C) continue;
            if (args.accumulate)
                for (int j = 0; j < nv; ++j) v[j] += sl.C[ci + j];
            if (vec) st4(sl.C + ci, make_float4(v[0], v[1], v[2], v[3]));
            else for (int j = 0; j < nv; ++j) sl.C[ci + j] = v[j];
        }
    }
}

// bias gradients of a batch of weight-gradient slots: C2[m] (+)= sum_k A[k*lda + m]   (tensor-core path)
__global__ void __launch_bounds__(128) colsum_slots_kernel(const GemmArgs args, int rows_per_block) {
    const GemmSlot& sl = args.slot[blockIdx.z];
    if (!sl.C2) return;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (sl.m > 0 ? sl.m : args.M)) return;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(args.K, r0 + rows_per_block);
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += sl.A[(size_t)r * sl.lda + m];
    atomicAdd(&sl.C2[m], s);
}

// C[m, n] *= silu'(Z[m, n]) for a batch of slots (second half of a split-K data-gradient GEMM)
__global__ void __launch_bounds__(256) mul_dsilu_slots_kernel(const GemmArgs args) {
    const GemmSlot& sl = args.slot[blockIdx.y];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)args.M * args.N) return;
    const int m = (int)(i / args.N), n = (int)(i % args.N);
    sl.C[(size_t)m * sl.ldc + n] *= dsilu(sl.Z[(size_t)m * sl.ldz + n]);
}

__global__ void mul_dsilu_kernel(float* __restrict__ c, const float* __restrict__ z, int64_t n4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = ld4(c + 4 * i) * dsilu4(ld4(z + 4 * i));
    st4(c + 4 * i, v);
}
// c[i] *= silu'(z[i]), n a multiple of 4 (rows of D floats)
int mul_dsilu_launch(float* c, const float* z, int64_t n, cudaStream_t st) {
    if (n <= 0) return 0;
    prof_begin(KC_GEMM, 12.0 * n, st);
    mul_dsilu_kernel<<<ceil_div(n / 4, 256), 256, 0, st>>>(c, z, n / 4);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

static int gemm_backend() {      // 0 = FFMA only, 1 = round-1 tcgen05 kernel (gemm_tc.cu), 2 = persistent TMA kernel (gemm_tc2.cu)
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("PAMNET_GEMM");
        mode = (e && strcmp(e, "ffma") == 0) ? 0 : (e && strcmp(e, "tc1") == 0) ? 1 : 2;
    }
    return mode;
}

static int gemm_small_backend() {      // PAMNET_GEMM_SMALL=0 sends skinny problems to the FFMA tile kernel again
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("PAMNET_GEMM_SMALL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on;
}

static int gemm_launch_impl(const GemmArgs& a, cudaStream_t st, bool allow_tc2);

int gemm_launch(const GemmArgs& a, cudaStream_t st) {
    // Slots the TMA kernel cannot take (one-row slots of the merged node-level weight-gradient launch, unaligned
    // operands) go through the other kernels in a second launch; everything else stays together.
    if (gemm_backend() == 2 && a.nslots > 1 && a.nslots <= kGemmMaxSlots) {
        int n_ok = 0;
        for (int i = 0; i < a.nslots; ++i) n_ok += gemm_tc2_slot_ok(a, i) ? 1 : 0;
        if (n_ok > 0 && n_ok < a.nslots) {
            GemmArgs ok = a, rest = a;
            ok.nslots = rest.nslots = 0;
            for (int i = 0; i < a.nslots; ++i) {
                if (gemm_tc2_slot_ok(a, i)) ok.slot[ok.nslots++] = a.slot[i];
                else rest.slot[rest.nslots++] = a.slot[i];
            }
            if (gemm_tc2_eligible(ok)) {
                PAMNET_TRY(gemm_launch_impl(ok, st, true));
                return gemm_launch_impl(rest, st, false);
            }
        }
    }
    return gemm_launch_impl(a, st, true);
}

static int gemm_launch_impl(const GemmArgs& a, cudaStream_t st, bool allow_tc2) {
    PAMNET_CHECK_ARG(a.nslots >= 1 && a.nslots <= kGemmMaxSlots, "gemm: nslots=%d", a.nslots);
    PAMNET_CHECK_ARG(a.nseg <= kGemmMaxSeg, "gemm: nseg=%d", a.nseg);
    PAMNET_CHECK_ARG(a.nseg == 0 || (a.mode == GEMM_NN && a.seg_len % BK == 0 && a.nseg * a.seg_len == a.K),
                     "gemm: bad K segmentation");
    PAMNET_CHECK_ARG(a.ksplit <= 1 || (a.epi == EPI_NONE && !a.accumulate), "gemm: split-K needs EPI_NONE");
    if (a.M <= 0 || a.N <= 0) return 0;
    if (a.K <= 0) return 0;   // callers zero-fill outputs themselves when K == 0 matters
    double bytes = 4.0 * a.nslots * ((double)a.M * a.K + (double)a.K * a.N + (double)a.M * a.N);
    if (a.epi == EPI_MUL_DSILU || (a.epi == EPI_BIAS_SILU && a.slot[0].C2)) bytes += 4.0 * a.nslots * (double)a.M * a.N;
    const int ks = a.ksplit > 1 ? a.ksplit : 1;
    double flops = 0.0;
    for (int i = 0; i < a.nslots; ++i) flops += 2.0 * (a.slot[i].m > 0 ? a.slot[i].m : a.M) * (double)a.N * a.K;
    const bool use_tc2 = gemm_backend() == 2 && allow_tc2 && gemm_tc2_eligible(a);
    if (use_tc2 || (gemm_backend() >= 1 && gemm_tc_eligible(a))) {
        // The tensor core adds each MMA into the fp32 accumulator with truncation, so the error of one
        // accumulation chain grows linearly with its length (measured: 1e-5 relative at K = 1536).  Long
        // reductions are therefore split into <= 128-deep chains per CTA and combined with fp32 atomics
        // (round-to-nearest); a SiLU' epilogue then runs as a separate pass over the summed result.
        GemmArgs b = a;
        const bool splittable = a.mode != GEMM_NT && (a.epi == EPI_NONE || a.epi == EPI_MUL_DSILU);
        const int want = ceil_div(a.K, 128);
        const bool long_nn = splittable && a.ksplit <= 1 && a.K > 256;     // plain-store semantics: zero C first
        if (long_nn) {
            for (int i = 0; i < a.nslots; ++i) {
                PAMNET_CHECK_ARG(a.slot[i].ldc == a.N, "gemm: split reduction needs a contiguous output");
                PAMNET_CUDA(cudaMemsetAsync(a.slot[i].C, 0, sizeof(float) * (size_t)a.M * a.N, st));
            }
            b.ksplit = want;
            b.epi = EPI_NONE;
        } else if (a.ksplit > 1 && want > a.ksplit) {
            b.ksplit = want;                                                // already accumulating with atomics
        }
        prof_begin(KC_GEMM, bytes, st);
        prof_flops(flops);
        int rc2 = use_tc2 ? gemm_tc2_launch(b, st) : 1;
        if (rc2 < 0) return rc2;
        if (rc2 == 1) {       // not the TMA kernel (backend choice, or its tensor-map table is full)
            PAMNET_CHECK_ARG(gemm_tc_eligible(b), "gemm: problem fits neither tensor-core kernel");
            PAMNET_TRY(gemm_tc_launch(b, st));
        }
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
        if (long_nn && a.epi == EPI_MUL_DSILU) {
            dim3 grid(ceil_div((int64_t)a.M * a.N, 256), a.nslots);
            prof_begin(KC_GEMM, 0.0, st);
            mul_dsilu_slots_kernel<<<grid, 256, 0, st>>>(a);
            prof_end(st);
            PAMNET_LAUNCH_CHECK();
        }
        return 0;
    }
    if (gemm_small_backend() && gemm_small_eligible(a)) {
        prof_begin(KC_GEMM, bytes, st);
        prof_flops(flops);
        PAMNET_TRY(gemm_small_launch(a, st));
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
        return 0;
    }
    // big tiles only when they still give >= 2 waves of CTAs on 148 SMs
    const long big_ctas = (long)ceil_div(a.M, 128) * ceil_div(a.N, 128) * ks * a.nslots;
    const bool big = a.M >= 128 && a.N >= 128 && big_ctas >= 2 * kNumSM;
    prof_begin(KC_GEMM, bytes, st);
    prof_flops(flops);
#define GEMM_LAUNCH(EPI_)                                                              \
    do {                                                                               \
        if (big) {                                                                     \
            dim3 grid(ceil_div(a.M, 128) * ceil_div(a.N, 128), ks, a.nslots);          \
            gemm_kernel<128, 128, 8, 8, EPI_><<<grid, GT, 0, st>>>(a);                 \
        } else {                                                                       \
            dim3 grid(ceil_div(a.M, 64) * ceil_div(a.N, 64), ks, a.nslots);            \
            gemm_kernel<64, 64, 4, 4, EPI_><<<grid, GT, 0, st>>>(a);                   \
        }                                                                              \
    } while (0)
    switch (a.epi) {
        case EPI_NONE: GEMM_LAUNCH(EPI_NONE); break;
        case EPI_BIAS: GEMM_LAUNCH(EPI_BIAS); break;
        case EPI_BIAS_SILU: GEMM_LAUNCH(EPI_BIAS_SILU); break;
        default: GEMM_LAUNCH(EPI_MUL_DSILU); break;
    }
#undef GEMM_LAUNCH
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

__global__ void colsum_kernel(const float* __restrict__ X, int64_t rows, int cols, int ld, int64_t rows_per_block,
                              float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = min(rows, r0 + rows_per_block);
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += X[r * ld + c];
    atomicAdd(&out[c], s);
}

int colsum_launch(const float* X, int64_t rows, int cols, int ld, float* out, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return 0;
    const int64_t rpb = 256;
    dim3 grid(ceil_div(cols, 128), ceil_div(rows, rpb));
    colsum_kernel<<<grid, 128, 0, st>>>(X, rows, cols, ld, rpb, out);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

}  // namespace pamnet
