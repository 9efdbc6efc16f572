// Tensor-core path of the batched GEMM (gemm.cuh): tcgen05.mma kind::tf32 with fp32 accumulators in TMEM,
// 3xTF32 error compensation (A_hi*B_hi + A_hi*B_lo + A_lo*B_hi) so that the result stays within ~1e-6 of an
// fp32 FFMA GEMM -- the accuracy the reference's fp32 nn.Linear layers need (BASELINE.json: 1e-5 rel), which
// plain TF32 (10-bit mantissa, ~1e-3) cannot give and tcgen05 has no fp32 input kind for.
//
// One CTA = one 128 x 128 output tile of one slot.  All 256 threads stage the operands: coalesced fp32 global
// loads -> split into tf32 hi / lo parts in registers -> st.shared in the UMMA canonical no-swizzle layout
// (always K-major: a thread that reads an m/n-contiguous operand walks four k rows and so owns a k-quad),
// double-buffered 16-wide k chunks.  One elected thread issues the MMAs (6 per chunk) and commits them to an
// mbarrier that releases the stage; the epilogue reads the accumulator tile back with tcgen05.ld (each warp
// its own 32-lane quarter) and applies bias / SiLU / SiLU' exactly like the FFMA kernel.
#include "gemm.cuh"

#include <stdlib.h>

namespace pamnet {
namespace {

constexpr int TM = 128, TN = 128, TK = 16;
constexpr int TC_THREADS = 256;
constexpr int kTileBytes = 8192;                 // one operand part of one stage: 4 k-quads x 128 rows x 16 B
constexpr int kStageBytes = 4 * kTileBytes;      // A_hi, A_lo, B_hi, B_lo
constexpr int kStages = 2;
constexpr int kTmemCols = 128;
constexpr size_t kTcSmem = (size_t)kStages * kStageBytes + 128;

// K-major canonical (no swizzle): 16-byte unit (m, kchunk) at kchunk * KM_LBO + m * 16  -> core matrix = 8 rows x 16 B
constexpr uint32_t KM_LBO = TM * 16, KM_SBO = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    // cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48) | layout NONE [61,64)
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// cute::UMMA::InstrDescriptor: c_format F32 [4,6)=1, a/b format TF32 [7,10),[10,13)=2, a_major bit 15, b_major bit 16,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must become a launch failure, never a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) asm volatile("trap;");
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
        "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// tf32 split with round-to-nearest: x = hi + lo + O(2^-22 |x|)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    const float rest = x - hi;
    uint32_t l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rest));
    lo = __uint_as_float(l);
}
__device__ __forceinline__ void split4(const float4& x, float4& hi, float4& lo) {
    split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y);
    split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
}

__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ float4 ld4g(const float* __restrict__ p, int valid, bool vec_ok) {
    if (valid >= 4 && vec_ok) return ld4(p);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid > 0) r.x = p[0];
    if (valid > 1) r.y = p[1];
    if (valid > 2) r.z = p[2];
    if (valid > 3) r.w = p[3];
    return r;
}

// MODE: GEMM_NT / GEMM_NN / GEMM_TN (gemm.cuh); EPI: GemmEpi
template <int MODE, int EPI>
__global__ void __launch_bounds__(TC_THREADS) gemm_tc_kernel(const GemmArgs args) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mma_done[kStages];
    __shared__ uint32_t tmem_base_s;

    constexpr bool A_KM = (MODE != GEMM_TN);      // A k-contiguous in global memory -> K-major smem
    constexpr bool B_KM = (MODE == GEMM_NT);
    constexpr uint32_t IDESC = make_idesc(0, 0);   // both operands are staged K-major

    const GemmSlot& sl = args.slot[blockIdx.z];
    const int M = args.M, N = args.N, K = args.K;
    const int tiles_n = (N + TN - 1) / TN;
    const int m0 = (blockIdx.x / tiles_n) * TM, n0 = (blockIdx.x % tiles_n) * TN;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    int k_begin = 0, k_end = K;
    if (args.ksplit > 1) {
        const int chunk = ((K + args.ksplit - 1) / args.ksplit + TK - 1) / TK * TK;
        k_begin = blockIdx.y * chunk;
        k_end = min(K, k_begin + chunk);
        if (k_begin >= k_end) return;           // uniform per CTA: taken before any barrier / TMEM allocation
    }

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (t == 32) {
        for (int s = 0; s < kStages; ++s) mbar_init(&mma_done[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    const bool a_vec = al16(sl.A) && (sl.lda % 4 == 0);
    const int nchunks = (k_end - k_begin + TK - 1) / TK;
    // weight-gradient mode: thread t always loads column m0 + (t & 127) of A, so the column sums of A
    // (= the bias gradient) ride along for free on the first column tile
    const bool do_bias = (MODE == GEMM_TN) && sl.C2 != nullptr && n0 == 0;
    float asum = 0.f;

    // global -> registers for chunk kc (both operands, 2 x float4 each).  Called one chunk AHEAD of its use, so
    // the L2 latency of chunk kc+1 overlaps the split / stage / MMA-issue of chunk kc (ncu: long_scoreboard).
    auto fetch = [&](int kc, float4 (&ra)[2], float4 (&rb)[2]) {
        const int k0 = k_begin + kc * TK;
        // ---- global -> registers (both operands, 2 x float4 each) ---------------------------------------------
        const float* Bp = sl.B;
        int ldb = sl.ldb, kb = k0;
        if (args.nseg > 0) {
            const int s = k0 / args.seg_len;
            Bp = args.seg_B[s];
            ldb = args.seg_ldb[s];
            kb = k0 - s * args.seg_len;
        }
        const bool b_vec = al16(Bp) && (ldb % 4 == 0);
        // k-contiguous operand: f = 2t + j -> (row f / 4, k-quad f % 4): a thread reads 32 contiguous bytes of a row.
        // m/n-contiguous operand: (row t % 128, k-quads 2 (t / 128) + j): four scalar loads down k, each one a
        // fully coalesced 128 B request per warp -- the thread then owns a k-quad and stores it as ONE 16-byte
        // K-major unit, so no transposing shared-memory traffic (and no MN-major descriptors) is needed.
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int fk = t * 2 + j;
            const int r_t = t & 127, q_t = (t >> 7) * 2 + j;
            if (A_KM) {
                const int r = fk >> 2, k4 = fk & 3, m = m0 + r, k = k0 + k4 * 4;
                ra[j] = (m < M) ? ld4g(sl.A + (size_t)m * sl.lda + k, k_end - k, a_vec) : make_float4(0, 0, 0, 0);
            } else {
                const int m = m0 + r_t, k = k0 + q_t * 4;
                float x[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = (m < M && k + i < k_end) ? sl.A[(size_t)(k + i) * sl.lda + m] : 0.f;
                ra[j] = make_float4(x[0], x[1], x[2], x[3]);
            }
            if (B_KM) {
                const int r = fk >> 2, k4 = fk & 3, n = n0 + r, k = kb + k4 * 4;
                rb[j] = (n < N) ? ld4g(Bp + (size_t)n * ldb + k, k_end - (k0 + k4 * 4), b_vec) : make_float4(0, 0, 0, 0);
            } else {
                const int n = n0 + r_t, k = kb + q_t * 4, kg = k0 + q_t * 4;
                float x[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = (n < N && kg + i < k_end) ? Bp[(size_t)(k + i) * ldb + n] : 0.f;
                rb[j] = make_float4(x[0], x[1], x[2], x[3]);
            }
        }
    };

    float4 ra[2], rb[2], na[2], nb[2];
    fetch(0, ra, rb);
    for (int kc = 0; kc < nchunks; ++kc) {
        const int st = kc & 1;
        unsigned char* stage = smem_raw + st * kStageBytes;
        if (kc + 1 < nchunks) fetch(kc + 1, na, nb);
        // ---- wait until the MMAs that read this stage two chunks ago are done ------------------------------
        if (kc >= kStages) mbar_wait(&mma_done[st], ((kc / kStages) - 1) & 1);
        // ---- split + stage ------------------------------------------------------------------------------------
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int fk = t * 2 + j;
            const uint32_t off_row = (uint32_t)(fk & 3) * KM_LBO + (uint32_t)(fk >> 2) * 16;
            const uint32_t off_col = (uint32_t)((t >> 7) * 2 + j) * KM_LBO + (uint32_t)(t & 127) * 16;
            float4 hi, lo;
            if (!A_KM) asum += (ra[j].x + ra[j].y) + (ra[j].z + ra[j].w);
            split4(ra[j], hi, lo);
            uint32_t off = A_KM ? off_row : off_col;
            *reinterpret_cast<float4*>(stage + 0 * kTileBytes + off) = hi;
            *reinterpret_cast<float4*>(stage + 1 * kTileBytes + off) = lo;
            split4(rb[j], hi, lo);
            off = B_KM ? off_row : off_col;
            *reinterpret_cast<float4*>(stage + 2 * kTileBytes + off) = hi;
            *reinterpret_cast<float4*>(stage + 3 * kTileBytes + off) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
        __syncthreads();
        // ---- one thread issues this chunk's MMAs -------------------------------------------------------------
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = smem_u32(stage);
#pragma unroll
            for (int ks = 0; ks < TK / 8; ++ks) {
                // a k-step of 8 tf32 = two 16-byte k-quads (stride LBO)
                const uint32_t k_off = ks * 2 * KM_LBO;
                const uint64_t a_hi = make_desc(base + 0 * kTileBytes + k_off, KM_LBO, KM_SBO);
                const uint64_t a_lo = make_desc(base + 1 * kTileBytes + k_off, KM_LBO, KM_SBO);
                const uint64_t b_hi = make_desc(base + 2 * kTileBytes + k_off, KM_LBO, KM_SBO);
                const uint64_t b_lo = make_desc(base + 3 * kTileBytes + k_off, KM_LBO, KM_SBO);
                umma_tf32(tmem, a_lo, b_hi, IDESC, (kc | ks) ? 1u : 0u);   // small terms first
                umma_tf32(tmem, a_hi, b_lo, IDESC, 1u);
                umma_tf32(tmem, a_hi, b_hi, IDESC, 1u);
            }
            umma_commit(&mma_done[st]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) { ra[j] = na[j]; rb[j] = nb[j]; }
    }

    if (do_bias && m0 + (t & 127) < M) atomicAdd(&sl.C2[m0 + (t & 127)], asum);   // C2 is zero-initialised by the caller

    // ---- all MMAs done? (commits complete in issue order, so the last one covers everything) ----------------
    {
        const int last = nchunks - 1;
        mbar_wait(&mma_done[last & 1], (last / kStages) & 1);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: warp w reads lanes [32 (w % 4), +32) x columns [64 (w / 4), +64) ----------------------------
    const int q = warp & 3, half = warp >> 2;
    const int m = m0 + q * 32 + lane;
    const bool c_vec = al16(sl.C) && (sl.ldc % 4 == 0) && (sl.C2 == nullptr || MODE == GEMM_TN || al16(sl.C2)) &&
                       (EPI != EPI_MUL_DSILU || (al16(sl.Z) && sl.ldz % 4 == 0));
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
        const int col0 = half * 64 + cb * 32;
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);   // warp-collective: no early exit above
        if (m >= M) continue;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int n = n0 + col0 + g * 4;
            if (n >= N) continue;
            float* x = v + g * 4;
            const size_t ci = (size_t)m * sl.ldc + n;
            const int nv = min(4, N - n);
            if (EPI == EPI_NONE && args.ksplit > 1) {
                for (int j = 0; j < nv; ++j) atomicAdd(&sl.C[ci + j], x[j]);
                continue;
            }
            const bool vec = c_vec && nv == 4;
            if (EPI == EPI_BIAS || EPI == EPI_BIAS_SILU) {
                if (sl.bias)
                    for (int j = 0; j < nv; ++j) x[j] += sl.bias[n + j];
                if (EPI == EPI_BIAS_SILU) {
                    if (sl.C2) {
                        if (vec) st4(sl.C2 + ci, make_float4(x[0], x[1], x[2], x[3]));
                        else for (int j = 0; j < nv; ++j) sl.C2[ci + j] = x[j];
                    }
                    for (int j = 0; j < nv; ++j) x[j] = silu(x[j]);
                }
            } else if (EPI == EPI_MUL_DSILU) {
                const size_t zi = (size_t)m * sl.ldz + n;
                for (int j = 0; j < nv; ++j) x[j] *= dsilu(sl.Z[zi + j]);
            }
            if (!sl.C) continue;
            if (vec) st4(sl.C + ci, make_float4(x[0], x[1], x[2], x[3]));
            else for (int j = 0; j < nv; ++j) sl.C[ci + j] = x[j];
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

template <int MODE, int EPI>
int launch_tc(const GemmArgs& a, dim3 grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        PAMNET_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MODE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
        configured = true;
    }
    gemm_tc_kernel<MODE, EPI><<<grid, TC_THREADS, kTcSmem, st>>>(a);
    return 0;
}

}  // namespace

bool gemm_tc_eligible(const GemmArgs& a) {
    if (a.M < 96 || a.N < 96 || a.K < 16 || a.accumulate) return false;
    // a 128 x 128 tile costs ~10 us of latency regardless of its depth: small problems stay on the FFMA kernel
    if (a.mode == GEMM_TN) return a.K >= 2048;
    return a.M >= 512;
}

// Launch the tensor-core kernel for an eligible problem; returns 0 on success (the caller counts the launch).
int gemm_tc_launch(const GemmArgs& a, cudaStream_t st) {

    const int ks = a.ksplit > 1 ? a.ksplit : 1;
    dim3 grid(ceil_div(a.M, TM) * ceil_div(a.N, TN), ks, a.nslots);
#define TC_CASE(MODE_, EPI_) return launch_tc<MODE_, EPI_>(a, grid, st)
    switch (a.mode) {
        case GEMM_NT:
            switch (a.epi) {
                case EPI_NONE: TC_CASE(GEMM_NT, EPI_NONE);
                case EPI_BIAS: TC_CASE(GEMM_NT, EPI_BIAS);
                case EPI_BIAS_SILU: TC_CASE(GEMM_NT, EPI_BIAS_SILU);
                default: TC_CASE(GEMM_NT, EPI_MUL_DSILU);
            }
        case GEMM_NN:
            switch (a.epi) {
                case EPI_NONE: TC_CASE(GEMM_NN, EPI_NONE);
                case EPI_BIAS: TC_CASE(GEMM_NN, EPI_BIAS);
                case EPI_BIAS_SILU: TC_CASE(GEMM_NN, EPI_BIAS_SILU);
                default: TC_CASE(GEMM_NN, EPI_MUL_DSILU);
            }
        default:
            switch (a.epi) {
                case EPI_NONE: TC_CASE(GEMM_TN, EPI_NONE);
                case EPI_BIAS: TC_CASE(GEMM_TN, EPI_BIAS);
                case EPI_BIAS_SILU: TC_CASE(GEMM_TN, EPI_BIAS_SILU);
                default: TC_CASE(GEMM_TN, EPI_MUL_DSILU);
            }
    }
#undef TC_CASE
}

}  // namespace pamnet
