// Tensor-core path of the batched GEMM (gemm.cuh): tcgen05.mma kind::tf32 with fp32 accumulators in TMEM,
// 3xTF32 error compensation (A_hi*B_hi + A_hi*B_lo + A_lo*B_hi) so that the result stays within ~1e-6 of an
// fp32 FFMA GEMM -- the accuracy the reference's fp32 nn.Linear layers need (BASELINE.json: 1e-5 rel), which
// plain TF32 (10-bit mantissa, ~1e-3) cannot give and tcgen05 has no fp32 input kind for.
//
// One CTA = one 128 x 128 output tile of one slot.  cp.async streams the raw fp32 operands, several 16-wide k
// chunks ahead, straight into the UMMA canonical no-swizzle layout (always K-major: a thread that reads an
// m/n-contiguous operand copies four k rows and so owns a k-quad); the threads then split their own units into
// tf32 hi (in place) / lo (second buffer) parts shared -> shared.  One elected thread issues the MMAs (6 per
// chunk) and commits them to an mbarrier that releases the stage; the epilogue reads the accumulator tile back
// with tcgen05.ld (each warp its own 32-lane quarter) and applies bias / SiLU / SiLU' exactly like the FFMA kernel.
#include "gemm.cuh"

#include <stdlib.h>

namespace pamnet {
// optional in-kernel timeline of CTA (0,0,0) (clock64 stamps), compiled in with -DPAMNET_TC_TRACE; read back with
// pamnet_debug_tc_trace (abi.cu).  Slots: 0 start, 1 after setup, 2 epilogue start, 3 end, 16+4kc.. converter
// (landed, slot free, converted, next issued), 128+2kc.. MMA warp (operands ready, issued)
#ifdef PAMNET_TC_TRACE
__device__ long long g_tc_trace[256];
#define TC_STAMP(i) do { if (trace_on) g_tc_trace[(i)] = clock64(); } while (0)
#else
#define TC_STAMP(i) do { } while (0)
#endif
int tc2_trace_read(long long* out, int n);
int tc_trace_read(long long* out, int n) {
    { const char* e = getenv("PAMNET_GEMM"); if (!(e && (strcmp(e, "tc1") == 0 || strcmp(e, "ffma") == 0))) return tc2_trace_read(out, n); }
#ifdef PAMNET_TC_TRACE
    PAMNET_CUDA(cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(long long) * (n < 256 ? n : 256)));
    return 0;
#else
    (void)out; (void)n;
    set_error("built without PAMNET_TC_TRACE");
    return -1;
#endif
}
namespace {

constexpr int TM = 128, TN = 128, TK = 16;
constexpr int kConv = 256;                       // 8 warps: copy, tf32 split, epilogue
constexpr int TC_THREADS = kConv + 32;           // + 1 warp whose lane 0 issues the MMAs
constexpr int kTileBytes = 8192;                 // one operand part of one stage: 4 k-quads x 128 rows x 16 B
constexpr int kRawStage = 2 * kTileBytes;        // A | B: raw fp32 from cp.async, rounded in place to the tf32 hi parts
constexpr int kLoStage = 2 * kTileBytes;         // A_lo | B_lo
constexpr int kRing = 5, kAhead = 3, kLoRing = 2;
constexpr int kTmemCols = 128;
constexpr int kEpLd = TN + 4;                    // epilogue staging row stride (floats): conflict-free 128-bit rows
constexpr size_t kTcSmem = (size_t)kRing * kRawStage + (size_t)kLoRing * kLoStage + 128;   // 112 KB: two CTAs per SM

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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {     // release.cta: orders the thread's earlier st.shared
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// tf32 split with round-to-nearest: x = hi + lo + O(2^-22 |x|).  The rounding is done with integer ALU ops (add half
// an ulp of the 10-bit mantissa, clear the 13 low bits = cvt.rna.tf32.f32 for finite normal inputs): the cvt
// instruction runs on the 16-lane conversion pipe and alone cost ~500 cycles per chunk (in-kernel trace).
__device__ __forceinline__ float rna_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = rna_tf32(x);
    lo = rna_tf32(x - hi);
}
__device__ __forceinline__ void split4(const float4& x, float4& hi, float4& lo) {
    split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y);
    split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
}

__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// cp.async with zero-fill: copies `bytes` (0..16 resp. 0..4) from global memory and zero-fills the rest of the unit
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp4(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MODE: GEMM_NT / GEMM_NN / GEMM_TN (gemm.cuh); EPI: GemmEpi
//
// Pipeline per 16-deep k chunk kc (ring of kRing raw stages, kLoRing lo stages, loads kAhead chunks ahead):
//   cp.async (issued kAhead iterations earlier) has put the raw fp32 operands of chunk kc into ring slot kc % kRing,
//   already in UMMA K-major order -> every thread rounds ITS OWN units to tf32 in place (hi) and writes the
//   residuals to lo slot kc % kLoRing -> mbarrier.arrive on full[kc % kRing] -> cp.async for chunk kc + kAhead goes
//   into the slot chunk kc - 2 used.  A ninth warp waits on full[], executes the generic->async proxy fence and
//   issues the chunk's 6 MMAs, committing them to done[kc % kRing].
// The proxy fence compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC: it belongs in the MMA warp, which has no copies
// in flight.  Nothing in the converters' loop waits on global memory or on a CTA-wide barrier: loads have kAhead
// iterations to land, and the MMAs of chunk kc run while chunk kc + 1 is converted.  An in-kernel clock64 trace
// (tools/gemm_trace.py) showed the loop to be bound by the converters' own instruction latency, hence: all source
// addressing hoisted out of the loop, the four units of a thread loaded before any is stored (the in-place store
// otherwise serialises them), and an epilogue staged through shared memory so that global stores are 512 B rows.
template <int MODE, int EPI>
__global__ void __launch_bounds__(TC_THREADS) gemm_tc_kernel(const GemmArgs args) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mma_done[kRing], full_bar[kRing];
    __shared__ uint32_t tmem_base_s;

    constexpr bool A_KM = (MODE != GEMM_TN);      // A k-contiguous in global memory
    constexpr bool B_KM = (MODE == GEMM_NT);
    constexpr uint32_t IDESC = make_idesc(0, 0);   // both operands are staged K-major

    const GemmSlot& sl = args.slot[blockIdx.z];
    const int M = sl.m > 0 ? sl.m : args.M, N = args.N, K = args.K;
    const int tiles_n = (N + TN - 1) / TN;
    const int m0 = (blockIdx.x / tiles_n) * TM, n0 = (blockIdx.x % tiles_n) * TN;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (m0 >= M) return;                        // uniform per CTA (slots with fewer rows than the launch)
#ifdef PAMNET_TC_TRACE
    const bool trace_on = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (t == 0 || t == kConv);
#endif
    if (t == 0) TC_STAMP(0);

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
        for (int s = 0; s < kRing; ++s) { mbar_init(&mma_done[s], 1); mbar_init(&full_bar[s], kConv); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    const int nchunks = (k_end - k_begin + TK - 1) / TK;
    const uint32_t smem_base = smem_u32(smem_raw);
    // Placement inside an 8 KB operand tile (unit (row, k-quad) lives at quad * LBO + row * 16).  Shared memory serves
    // 128-bit accesses a quarter-warp at a time, so 8 consecutive lanes must touch 8 consecutive rows of ONE k-quad
    // (128 contiguous bytes) -- both for the cp.async writes and for the conversion pass (ncu: the naive
    // "4 quads of a row per 4 lanes" order cost 52 wavefronts per LDGSTS.128 and 16 per LDS/STS.128).
    //  * k-contiguous source: unit f = t + 256 j -> row (f / 32) * 8 + f % 8, quad (f / 8) % 4: a warp copies 8 rows x
    //    64 contiguous bytes per instruction (full sectors) and every thread later converts the units it copied.
    //  * m/n-contiguous source: 4-byte element e = t + 256 c -> row ((e / 32) % 16) * 8 + e % 8, k = 4 (e / 512) +
    //    (e / 8) % 4: a warp copies 4 k-rows x 8 consecutive rows (four full sectors) into 128 contiguous bytes.  The
    //    four elements of a unit come from four lanes of the SAME warp, so a __syncwarp() after the wait makes them
    //    visible to the lane that converts the unit: lane L owns units u = L, L + 32 of its warp's 64
    //    (u -> copy instruction c = u / 8: row (warp + 8 (c & 1)) * 8 + u % 8, quad c / 2).
    const int qd = (lane >> 3) & 3, l8 = lane & 7;
    uint32_t off_km[2], off_mn[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        off_km[j] = (uint32_t)qd * KM_LBO + (uint32_t)((warp + 8 * j) * 8 + l8) * 16;
        const int c = qd + 4 * j;
        off_mn[j] = (uint32_t)(c >> 1) * KM_LBO + (uint32_t)((warp + 8 * (c & 1)) * 8 + l8) * 16;
    }
    const int row_mn = (warp + 8 * (qd & 1)) * 8 + l8;     // the single row this thread converts (m/n-contiguous case)

    // ---- copy addressing, hoisted out of the chunk loop ------------------------------------------------------------
    // k-contiguous: row pointers (already offset by the thread's k-quad); m/n-contiguous: the two rows (parity of
    // the copy instruction) and the k-row lane / 8 inside each quad
    const float* a_row[2] = {nullptr, nullptr};
    const float* b_row[2] = {nullptr, nullptr};
    bool a_ok[2] = {false, false}, b_ok[2] = {false, false};
    uint32_t dst_km[2], dst_mn[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int r = (warp + 8 * j) * 8 + l8;
        a_ok[j] = m0 + r < M;
        b_ok[j] = n0 + r < N;
        if (A_KM) a_row[j] = sl.A + (size_t)(a_ok[j] ? m0 + r : 0) * sl.lda + qd * 4;
        if (B_KM) b_row[j] = sl.B + (size_t)(b_ok[j] ? n0 + r : 0) * sl.ldb + qd * 4;
        dst_km[j] = off_km[j];
        dst_mn[j] = (uint32_t)r * 16 + (uint32_t)qd * 4;          // + (c / 2) * LBO per copy instruction c, parity j
    }

    auto issue = [&](int kc) {
        if (kc < nchunks) {
            const int k0 = k_begin + kc * TK;
            const uint32_t sa = smem_base + (uint32_t)(kc % kRing) * kRawStage, sb = sa + kTileBytes;
            const int krem = k_end - k0;                       // > 0
            if (A_KM) {
                const int nb = max(0, min(16, (krem - qd * 4) * 4));
#pragma unroll
                for (int j = 0; j < 2; ++j) cp16(sa + dst_km[j], a_row[j] + k0, a_ok[j] ? nb : 0);
            } else {
                const float* base = sl.A + (size_t)k0 * sl.lda + m0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int kk = (c >> 1) * 4 + qd;
                    const bool ok = a_ok[c & 1] && kk < krem;
                    cp4(sa + dst_mn[c & 1] + (uint32_t)(c >> 1) * KM_LBO,
                        ok ? base + kk * sl.lda + (warp + 8 * (c & 1)) * 8 + l8 : sl.A, ok ? 4 : 0);
                }
            }
            if (B_KM) {
                const int nb = max(0, min(16, (krem - qd * 4) * 4));
#pragma unroll
                for (int j = 0; j < 2; ++j) cp16(sb + dst_km[j], b_row[j] + k0, b_ok[j] ? nb : 0);
            } else {
                const float* Bp = sl.B;
                int ldb = sl.ldb, kb = k0;
                if (args.nseg > 0) {
                    const int s = k0 / args.seg_len;
                    Bp = args.seg_B[s];
                    ldb = args.seg_ldb[s];
                    kb = k0 - s * args.seg_len;
                }
                const float* base = Bp + (size_t)kb * ldb + n0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int kk = (c >> 1) * 4 + qd;
                    const bool ok = b_ok[c & 1] && kk < krem;
                    cp4(sb + dst_mn[c & 1] + (uint32_t)(c >> 1) * KM_LBO,
                        ok ? base + kk * ldb + (warp + 8 * (c & 1)) * 8 + l8 : Bp, ok ? 4 : 0);
                }
            }
        }
        cp_commit();     // always: keeps the group count per iteration uniform
    };

    pdl_wait();          // operands (and the zeroed split-K output) come from earlier kernels of the stream
    if (warp < kConv / 32) {
#pragma unroll
        for (int p = 0; p < kAhead; ++p) issue(p);
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (t == 0) TC_STAMP(1);

    if (warp == kConv / 32) {
        // ---- MMA warp ---------------------------------------------------------------------------------------------
        if (lane == 0) {
            const uint64_t dconst = make_desc(0, KM_LBO, KM_SBO);
            for (int kc = 0; kc < nchunks; ++kc) {
                mbar_wait(&full_bar[kc % kRing], (kc / kRing) & 1);          // acquire: the converters' st.shared
                if (kc < 16) TC_STAMP(128 + 2 * kc);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
                const uint32_t rb = (smem_base + (uint32_t)(kc % kRing) * kRawStage) >> 4;
                const uint32_t lb = (smem_base + (uint32_t)kRing * kRawStage + (uint32_t)(kc % kLoRing) * kLoStage) >> 4;
#pragma unroll
                for (int ks = 0; ks < TK / 8; ++ks) {
                    // a k-step of 8 tf32 = two 16-byte k-quads (stride LBO)
                    const uint32_t k_off = (ks * 2 * KM_LBO) >> 4, tb = kTileBytes >> 4;
                    const uint64_t a_hi = dconst | ((rb + k_off) & 0x3FFF), b_hi = dconst | ((rb + tb + k_off) & 0x3FFF);
                    const uint64_t a_lo = dconst | ((lb + k_off) & 0x3FFF), b_lo = dconst | ((lb + tb + k_off) & 0x3FFF);
                    umma_tf32(tmem, a_lo, b_hi, IDESC, (kc | ks) ? 1u : 0u);   // small terms first
                    umma_tf32(tmem, a_hi, b_lo, IDESC, 1u);
                    umma_tf32(tmem, a_hi, b_hi, IDESC, 1u);
                }
                umma_commit(&mma_done[kc % kRing]);
                if (kc < 16) TC_STAMP(129 + 2 * kc);
            }
        }
    } else {
        // ---- converter warps --------------------------------------------------------------------------------------
        // weight-gradient mode: a thread always converts units of ONE column m0 + row_mn of A, so the column sums
        // of A (= the bias gradient) ride along for free on the first column tile
        const bool do_bias = (MODE == GEMM_TN) && sl.C2 != nullptr && n0 == 0;
        float asum = 0.f;

        for (int kc = 0; kc < nchunks; ++kc) {
            cp_wait<kAhead - 1>();                               // this thread's copies of chunk kc have landed
            if (!A_KM || !B_KM) __syncwarp();                    // ... and its warp's (units assembled from 4 lanes)
            if (t == 0 && kc < 16) TC_STAMP(16 + 4 * kc);
            // the lo slot of this chunk and the raw slot of chunk kc + kAhead were last read by the MMAs of chunk kc - 2
            if (kc >= 2) mbar_wait(&mma_done[(kc - 2) % kRing], ((kc - 2) / kRing) & 1);
            if (t == 0 && kc < 16) TC_STAMP(17 + 4 * kc);
            unsigned char* raw = smem_raw + (kc % kRing) * kRawStage;
            unsigned char* los = smem_raw + kRing * kRawStage + (kc % kLoRing) * kLoStage;
            float4 va[2], vb[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                va[j] = *reinterpret_cast<const float4*>(raw + (A_KM ? off_km[j] : off_mn[j]));
                vb[j] = *reinterpret_cast<const float4*>(raw + kTileBytes + (B_KM ? off_km[j] : off_mn[j]));
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t oa = A_KM ? off_km[j] : off_mn[j], ob = B_KM ? off_km[j] : off_mn[j];
                float4 hi, lo;
                if (!A_KM) asum += (va[j].x + va[j].y) + (va[j].z + va[j].w);
                split4(va[j], hi, lo);
                *reinterpret_cast<float4*>(raw + oa) = hi;
                *reinterpret_cast<float4*>(los + oa) = lo;
                split4(vb[j], hi, lo);
                *reinterpret_cast<float4*>(raw + kTileBytes + ob) = hi;
                *reinterpret_cast<float4*>(los + kTileBytes + ob) = lo;
            }
            mbar_arrive(&full_bar[kc % kRing]);
            if (t == 0 && kc < 16) TC_STAMP(18 + 4 * kc);
            issue(kc + kAhead);
            if (t == 0 && kc < 16) TC_STAMP(19 + 4 * kc);
        }
        cp_wait<0>();

        if (do_bias && m0 + row_mn < M) atomicAdd(&sl.C2[m0 + row_mn], asum);   // C2 is zero-initialised by the caller

        // ---- all MMAs done? (commits complete in issue order, so the last one covers everything) ----------------
        {
            const int last = nchunks - 1;
            mbar_wait(&mma_done[last % kRing], (last / kRing) & 1);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        pdl_trigger();   // the next kernel of the stream may set itself up while this tile is written back
        if (t == 0) TC_STAMP(2);

        // ---- epilogue ---------------------------------------------------------------------------------------------
        // phase 1: warp w reads TMEM lanes [32 (w % 4), +32) x columns [64 (w / 4), +64) and parks them in the (now
        // idle) operand ring as a [128][kEpLd] fp32 tile; phase 2: warp w streams rows 16 w .. 16 w + 15, one 512 B
        // row per instruction, through bias / SiLU / SiLU' to global memory.
        static_assert((size_t)TM * kEpLd * sizeof(float) <= (size_t)kRing * kRawStage, "epilogue tile must fit in the ring");
        float* ep = reinterpret_cast<float*>(smem_raw);
        {
            const int q = warp & 3, half = warp >> 2;
            float* dst = ep + (q * 32 + lane) * kEpLd + half * 64;
#pragma unroll 1
            for (int cb = 0; cb < 2; ++cb) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64 + cb * 32), v);
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    *reinterpret_cast<float4*>(dst + cb * 32 + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kConv) : "memory");

        const int n = n0 + lane * 4;
        const int nv = min(4, N - n);                         // <= 0: this lane's columns are outside the matrix
        const bool c_vec = al16(sl.C) && (sl.ldc % 4 == 0) && (sl.C2 == nullptr || MODE == GEMM_TN || al16(sl.C2)) &&
                           (EPI != EPI_MUL_DSILU || (al16(sl.Z) && sl.ldz % 4 == 0));
        const bool vec = c_vec && nv == 4;
        float bias_v[4] = {0.f, 0.f, 0.f, 0.f};
        if ((EPI == EPI_BIAS || EPI == EPI_BIAS_SILU) && sl.bias)
            for (int j = 0; j < 4; ++j) if (j < nv) bias_v[j] = sl.bias[n + j];
#pragma unroll 4
        for (int rr = 0; rr < 16; ++rr) {
            const int r = warp * 16 + rr, m = m0 + r;
            if (m >= M || nv <= 0) continue;
            const float4 acc = *reinterpret_cast<const float4*>(ep + r * kEpLd + lane * 4);
            float x[4] = {acc.x, acc.y, acc.z, acc.w};
            const size_t ci = (size_t)m * sl.ldc + n;
            if (EPI == EPI_NONE && args.ksplit > 1) {
                if (vec) red4(sl.C + ci, x[0], x[1], x[2], x[3]);
                else for (int j = 0; j < nv; ++j) atomicAdd(&sl.C[ci + j], x[j]);
                continue;
            }
            if (EPI == EPI_BIAS || EPI == EPI_BIAS_SILU) {
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] += bias_v[j];
                if (EPI == EPI_BIAS_SILU) {
                    if (sl.C2) {
                        if (vec) st4(sl.C2 + ci, make_float4(x[0], x[1], x[2], x[3]));
                        else for (int j = 0; j < nv; ++j) sl.C2[ci + j] = x[j];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = silu(x[j]);
                }
            } else if (EPI == EPI_MUL_DSILU) {
                const size_t zi = (size_t)m * sl.ldz + n;
                if (vec) {
                    const float4 z = ld4(sl.Z + zi);
                    x[0] *= dsilu(z.x); x[1] *= dsilu(z.y); x[2] *= dsilu(z.z); x[3] *= dsilu(z.w);
                } else {
                    for (int j = 0; j < nv; ++j) x[j] *= dsilu(sl.Z[zi + j]);
                }
            }
            if (!sl.C) continue;
            if (vec) st4(sl.C + ci, make_float4(x[0], x[1], x[2], x[3]));
            else for (int j = 0; j < nv; ++j) sl.C[ci + j] = x[j];
        }
        if (t == 0) TC_STAMP(3);
    }   // converter warps

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

template <int MODE, int EPI>
int launch_tc(const GemmArgs& a, dim3 grid, cudaStream_t st) {
    PAMNET_TRY(func_smem_once(reinterpret_cast<const void*>(gemm_tc_kernel<MODE, EPI>), kTcSmem));
    if (pdl_level() == 1) PAMNET_CUDA(launch_pdl(gemm_tc_kernel<MODE, EPI>, grid, dim3(TC_THREADS), kTcSmem, st, a));
    else gemm_tc_kernel<MODE, EPI><<<grid, TC_THREADS, kTcSmem, st>>>(a);
    return 0;
}

}  // namespace

bool gemm_tc_eligible(const GemmArgs& a) {
    if (a.M < 64 || a.N < 64 || a.K < 16 || a.accumulate) return false;
    // forward GEMMs cannot split their reduction (bias / SiLU epilogue) and one tensor-core accumulation chain loses
    // accuracy linearly with its length (gemm.cu): long ones stay on the FFMA kernel
    if (a.mode == GEMM_NT && a.K > 256) return false;
    // cp.async moves 16-byte units of the k-contiguous operands: they must be 16 B aligned row by row
    auto ok16 = [](const float* p, int ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0; };
    for (int i = 0; i < a.nslots; ++i) {
        if (a.mode != GEMM_TN && !ok16(a.slot[i].A, a.slot[i].lda)) return false;
        if (a.mode == GEMM_NT && !ok16(a.slot[i].B, a.slot[i].ldb)) return false;
    }
    return true;
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
