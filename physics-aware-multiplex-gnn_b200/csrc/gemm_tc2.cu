// Persistent, TMA-fed tcgen05 GEMM (round 2): the tensor-core path of the batched GEMM of gemm.cuh.
//
// Same arithmetic as gemm_tc.cu (tcgen05.mma kind::tf32, fp32 accumulators in TMEM, 3xTF32 error compensation:
// A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, layers/basic.py:19-22 at fp32 accuracy), different data path.  Round 1's kernel was
// bound by the SM's load/store unit: 8 converter warps copied raw operands with cp.async (4-byte copies for the
// m/n-contiguous operands), rewrote them in place as tf32 hi parts and wrote the lo parts -- 18.4 k cycles per
// 128^3 tile of which the MMAs need 3.2 k.  Here:
//   * operands arrive by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B, mbarrier transaction counts) straight from the
//     row-major global tensors through tensor maps: K-major tiles as one {32 k, 128 rows} box, m/n-contiguous
//     operands (weight-gradient and data-gradient modes) as four {32 mn, 32 k} boxes consumed with MN-MAJOR shared-memory
//     descriptors -- no element-wise staging of either layout; out-of-range rows / k are zero-filled by the TMA unit;
//   * the raw fp32 tile IS the hi operand: kind::tf32 ignores the 13 low mantissa bits (hi = trunc(x)).  Four converter
//     warps only produce lo = rna_tf32(x - trunc(x)) into a 2-slot ring (one 128-bit load + store per 4 elements;
//     round 1: LDGSTS + LDS + 2 STS);
//   * the kernel is persistent: a CTA walks (slot, k-split, tile) work items; TMEM holds TWO 128-column accumulators
//     so four epilogue warps drain tile i (tcgen05.ld -> per-warp staging -> 128-bit coalesced stores, bias / SiLU /
//     SiLU' / split-K reductions as before) while the MMA warp already works on tile i + 1 and the TMA warp prefetches
//     up to four 32-deep chunks ahead across tile boundaries;
//   * precision 1 (GemmArgs::precision == 1, opt-in): single-pass TF32 straight from the TMA tiles, no converters
//     (the reduced-precision node-MLP path of BASELINE.json configs[2]; >= bf16's 8-bit mantissa).
// Warp roles (448 threads): 0 TMA producer, 1 MMA issuer + TMEM owner, 2-5 converters, 6-13 epilogue (two warps per
// TMEM lane quarter, 64 columns each).
#include "gemm.cuh"

#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

namespace pamnet {
// optional clock64 timeline of CTA 0 (-DPAMNET_TC_TRACE; tools/gemm_trace.py).  Slots: 0 start, 1 setup done, 2 pdl_wait
// passed; per chunk c < 12: 8+4c TMA issued, 9+4c converters saw the data, 10+4c converted, 11+4c MMAs issued;
// per work item i < 4: 64+4i accumulator full (epilogue), 65+4i TMEM drained, 66+4i stores issued
#ifdef PAMNET_TC_TRACE
__device__ long long g_tc2_trace[96];
#define T2_STAMP(i) do { if (blockIdx.x == 0 && (i) < 96) g_tc2_trace[(i)] = clock64(); } while (0)
#else
#define T2_STAMP(i) do { } while (0)
#endif
int tc2_trace_read(long long* out, int n) {
#ifdef PAMNET_TC_TRACE
    PAMNET_CUDA(cudaMemcpyFromSymbol(out, g_tc2_trace, sizeof(long long) * (n < 96 ? n : 96)));
    return 0;
#else
    (void)out; (void)n;
    set_error("built without PAMNET_TC_TRACE");
    return -1;
#endif
}
namespace {

constexpr int TM = 128, TN = 128, TK = 32;
constexpr int kRaw = 3, kLo = 2;                  // ring depths (32-deep chunks)
constexpr int kTile = TM * TK * 4;                // 16 KB: one operand tile of one chunk
constexpr int kStage = 2 * kTile;                 // A | B
constexpr int kThreads = 448;
constexpr int kConvWarp0 = 2, kConvThreads = 128, kEpiWarp0 = 6, kEpiWarps = 8, kEpiThreads = kEpiWarps * 32;
constexpr int kStageLd = 36;                      // floats: row stride of the per-warp epilogue staging (conflict-free 128-bit)
constexpr int kEpiStage = 32 * kStageLd * 4;      // bytes per epilogue warp
constexpr size_t kSmem = 1024 /* alignment slack */ + (size_t)kRaw * kStage + (size_t)kLo * kStage + kEpiWarps * kEpiStage;
constexpr int kInline = 34;                       // by-value tensor maps per launch (4.3 KB)
constexpr int kTmemCols = 512;                    // two buffers x (main | cross-term) accumulators of 128 columns

struct Slot2 {
    const float* bias;
    const float* Z;
    float* C;
    float* C2;
    int ldc, ldz, m;
    int map_a, map_b;
};

struct Args2 {
    int M, N, K;
    int mode, epi, ksplit, nslots, precision;
    int nseg, seg_len;
    int tiles_m, tiles_n;
    int total;                                    // work items: nslots * ksplit * tiles_m * tiles_n
    int seg_map[kGemmMaxSeg];
    Slot2 slot[kGemmMaxSlots];
    const CUtensorMap* maps;                      // device table of tensor maps for STABLE operands (parameters); index >= 0
    CUtensorMap inl[kInline];                     // maps of this launch's other operands, by value; index -(i + 1)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
        if (spin > (1u << 26)) asm volatile("trap;");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
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
// 32 lanes x 32 columns -> v[32] (thread = lane / row); the caller issues tcgen05.wait::ld before reading v
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
          "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
          "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr) : "memory");
}
// orders later uses of v[] after the tcgen05.wait::ld that precedes this (volatile asms keep their relative order; without
// it the compiler may hoist register-only arithmetic on v[] above the wait)
__device__ __forceinline__ void reg_fence32(float* v) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                      "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]),
                      "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]),
                      "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31]));
}
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// lo part of the 3xTF32 split against the HARDWARE's hi = trunc_tf32(x): x - trunc(x) is exact in fp32 (<= 13
// significant bits); rounding it to tf32 here (instead of letting the tensor core truncate it too) keeps the split
// unbiased: measured on the configs[1] parity ladder, worst gradient 0.28 of the limit with, 0.52 without
__device__ __forceinline__ float lo_of(float x) {
    const float d = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    return __uint_as_float((__float_as_uint(d) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float4 lo4(const float4 v) { return make_float4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w)); }

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor), version 1:
//  K-major tile  [128 rows][128 B], SWIZZLE_128B (layout type 2; 16-byte chunk ^= row % 8): 8-row groups 1024 B apart
//      (SBO); a k-step of 8 advances the start address by 32 B.
//  MN-major tile [4 blocks][32 k-rows][128 B]: for 32-bit operands the only MN-major layout the tensor core accepts is
//      SWIZZLE_128B_BASE32B (layout type 1; 32-byte chunk ^= k-row % 4; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) with
//      atoms of 4 k-rows x 32 elements: atoms 512 B apart along k (SBO), 32-element blocks 4096 B apart (LBO); a k-step
//      of 8 (two atoms) advances the start address by 1024 B.
__device__ __forceinline__ uint64_t desc_base(bool mn_major) {
    const uint64_t lbo = mn_major ? (4096u >> 4) : 1u, sbo = mn_major ? (512u >> 4) : (1024u >> 4);
    return (lbo << 16) | (sbo << 32) | (1ull << 46) | ((mn_major ? 1ull : 2ull) << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 [4,6)=1, a/b format TF32 [7,10),[10,13)=2, a_major bit 15, b_major bit 16,
// N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

struct Work { int slot, ky, m0, n0; };
__device__ __forceinline__ Work decode(const Args2& a, int id) {
    // tile fastest, then k-split, then slot: CTAs working side by side share the slot's operands in L2
    const int tiles = a.tiles_m * a.tiles_n;
    Work w;
    const int tile = id % tiles;
    const int rest = id / tiles;
    w.ky = rest % a.ksplit;
    w.slot = rest / a.ksplit;
    w.m0 = (tile / a.tiles_n) * TM;
    w.n0 = (tile % a.tiles_n) * TN;
    return w;
}
// k range of a work item, in 32-deep chunks
__device__ __forceinline__ void k_range(const Args2& a, int ky, int& k_begin, int& k_end) {
    k_begin = 0; k_end = a.K;
    if (a.ksplit > 1) {
        const int chunk = ((a.K + a.ksplit - 1) / a.ksplit + TK - 1) / TK * TK;
        k_begin = ky * chunk;
        k_end = min(a.K, k_begin + chunk);
    }
}

__global__ void __launch_bounds__(kThreads, 1) gemm_tc2_kernel(const __grid_constant__ Args2 args) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[kRaw], raw_free[kRaw], conv_bar[kLo], lo_free[kLo], acc_full[2], acc_free[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_colsum[4][TM];        // per converter warp: column sums of A (bias gradient)

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (t == 0) T2_STAMP(0);
    // SWIZZLE_128B tiles need 1024 B alignment
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* raw_ring = smem;                                   // [kRaw][A 16 KB | B 16 KB]
    unsigned char* lo_ring = smem + (size_t)kRaw * kStage;            // [kLo][A_lo | B_lo]
    unsigned char* epi_stage = lo_ring + (size_t)kLo * kStage;        // [4 warps][32][kStageLd] floats

    const bool three = args.precision != 1;
    const bool four = args.precision == 4;        // + A_lo * B_lo (hi = trunc leaves |lo| < 2^-10 |x|: the term is ~2^-22 |ab| rms)
    const bool a_mn = args.mode == GEMM_TN, b_mn = args.mode != GEMM_NT;

    if (t == 0) {
        for (int s = 0; s < kRaw; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&raw_free[s], 1); }
        for (int s = 0; s < kLo; ++s) { mbar_init(&conv_bar[s], kConvThreads); mbar_init(&lo_free[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_free[s], kEpiThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (t == 0) T2_STAMP(1);

    pdl_wait();
    if (t == 0) T2_STAMP(2);          // operands (and zeroed split-K outputs) come from earlier kernels of the stream

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int c = 0;                                        // chunk counter across work items
            for (int id = blockIdx.x; id < args.total; id += gridDim.x) {
                const Work w = decode(args, id);
                const Slot2& sl = args.slot[w.slot];
                int k_begin, k_end;
                k_range(args, w.ky, k_begin, k_end);
                for (int k0 = k_begin; k0 < k_end; k0 += TK, ++c) {
                    const int s = c % kRaw;
                    mbar_wait(&raw_free[s], ((c / kRaw) & 1) ^ 1);
                    if (c < 12) T2_STAMP(8 + 4 * c);
                    mbar_expect_tx(&full_bar[s], kStage);
                    const uint32_t sa = smem_u32(raw_ring + (size_t)s * kStage), sb = sa + kTile;
                    const CUtensorMap* ma = sl.map_a >= 0 ? &args.maps[sl.map_a] : &args.inl[-1 - sl.map_a];
                    if (!a_mn) {
                        tma_load_2d(sa, ma, k0, w.m0, &full_bar[s]);
                    } else {
#pragma unroll
                        for (int b = 0; b < 4; ++b) tma_load_2d(sa + b * 4096, ma, w.m0 + 32 * b, k0, &full_bar[s]);
                    }
                    const CUtensorMap* mb = sl.map_b >= 0 ? &args.maps[sl.map_b] : &args.inl[-1 - sl.map_b];
                    int kb = k0;
                    if (args.nseg > 0) {
                        const int sg = k0 / args.seg_len;
                        const int mi = args.seg_map[sg];
                        mb = mi >= 0 ? &args.maps[mi] : &args.inl[-1 - mi];
                        kb = k0 - sg * args.seg_len;
                    }
                    if (!b_mn) {
                        tma_load_2d(sb, mb, kb, w.n0, &full_bar[s]);
                    } else {
#pragma unroll
                        for (int b = 0; b < 4; ++b) tma_load_2d(sb + b * 4096, mb, w.n0 + 32 * b, kb, &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(a_mn, b_mn);
            const uint64_t da = desc_base(a_mn), db = desc_base(b_mn);
            const uint32_t a_step = a_mn ? (1024u >> 4) : (32u >> 4), b_step = b_mn ? (1024u >> 4) : (32u >> 4);
            int c = 0, cc = 0, it = 0;                         // chunks, chunks that go through the converters, work items
            for (int id = blockIdx.x; id < args.total; id += gridDim.x) {
                const Work w = decode(args, id);
                int k_begin, k_end;
                k_range(args, w.ky, k_begin, k_end);
                if (k_begin >= k_end) continue;              // empty k-split slice: every role skips it
                // the converters also hold a chunk while they sum A's columns (bias gradient) in single-pass mode
                const bool use_conv = three || (a_mn && args.slot[w.slot].C2 != nullptr && w.n0 == 0);
                const int acc = it & 1;
                mbar_wait(&acc_free[acc], ((it >> 1) & 1) ^ 1);          // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // The tensor core adds every MMA into the fp32 accumulator with truncation: the error of a chain grows with
                // the number of accumulations and is biased (measured: the 48 accumulations of a 128-deep 3xTF32 chain put
                // the weight gradients at 0.8 of the parity limit, 20 x the error of an FFMA GEMM).  The hi * hi products
                // therefore get an accumulator of their own (16 accumulations per 128-deep chain); the cross terms, 2^-11
                // of its magnitude, go to a second one and are added in the epilogue with round-to-nearest.
                const uint32_t d = tmem + (uint32_t)(acc * 2 * TN), dx = d + TN;
                bool first = true;
                for (int k0 = k_begin; k0 < k_end; k0 += TK, ++c) {
                    const int s = c % kRaw, l = cc % kLo;
                    mbar_wait(&full_bar[s], (c / kRaw) & 1);             // TMA data landed
                    if (use_conv) mbar_wait(&conv_bar[l], (cc / kLo) & 1);   // acquire: the converters' (proxy-fenced) st.shared
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t ra = smem_u32(raw_ring + (size_t)s * kStage) >> 4, rb = ra + (kTile >> 4);
                    const uint32_t la = smem_u32(lo_ring + (size_t)l * kStage) >> 4, lb = la + (kTile >> 4);
#pragma unroll
                    for (int ks = 0; ks < TK / 8; ++ks) {
                        const uint64_t a_hi = da | ((ra + ks * a_step) & 0x3FFF), b_hi = db | ((rb + ks * b_step) & 0x3FFF);
                        if (three) {
                            const uint64_t a_lo = da | ((la + ks * a_step) & 0x3FFF), b_lo = db | ((lb + ks * b_step) & 0x3FFF);
                            if (four) umma_tf32(dx, a_lo, b_lo, idesc, first ? 0u : 1u);
                            umma_tf32(dx, a_lo, b_hi, idesc, (first && !four) ? 0u : 1u);
                            umma_tf32(dx, a_hi, b_lo, idesc, 1u);
                            umma_tf32(d, a_hi, b_hi, idesc, first ? 0u : 1u);
                        } else {
                            umma_tf32(d, a_hi, b_hi, idesc, first ? 0u : 1u);
                        }
                        first = false;
                    }
                    umma_commit(&raw_free[s]);
                    if (c < 12) T2_STAMP(11 + 4 * c);
                    if (use_conv) { umma_commit(&lo_free[l]); ++cc; }
                }
                umma_commit(&acc_full[acc]);
                ++it;
            }
        }
    } else if (warp < kEpiWarp0) {
        // ================= converters: lo = rna(x - trunc(x)), element-wise on the swizzled tiles =================
        const int ct = t - kConvWarp0 * 32;                  // 0..127
        // weight-gradient mode: column sums of A (= the bias gradient) ride along on the first column tile.  Thread ct
        // always sees the same logical 16-byte chunk of MN block j / 2 (j = its 8 copies): physical chunk ct % 8 of k-row
        // ct / 8 (+ 16), un-swizzled (32-byte chunk ^= k-row % 4).
        float asum[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) asum[i] = 0.f;
        int c = 0, cc = 0;
        for (int id = blockIdx.x; id < args.total; id += gridDim.x) {
            const Work w = decode(args, id);
            const Slot2& sl = args.slot[w.slot];
            int k_begin, k_end;
            k_range(args, w.ky, k_begin, k_end);
            if (k_begin >= k_end) continue;
            const bool do_bias = a_mn && sl.C2 != nullptr && w.n0 == 0;
            if (!three && !do_bias) { c += (k_end - k_begin + TK - 1) / TK; continue; }
            for (int k0 = k_begin; k0 < k_end; k0 += TK, ++c, ++cc) {
                const int s = c % kRaw, l = cc % kLo;
                mbar_wait(&full_bar[s], (c / kRaw) & 1);                 // TMA data landed
                if (ct == 0 && c < 12) T2_STAMP(9 + 4 * c);
                const float4* ra = reinterpret_cast<const float4*>(raw_ring + (size_t)s * kStage);
                mbar_wait(&lo_free[l], ((cc / kLo) & 1) ^ 1);            // the MMAs of the chunk that last used this slot are done
                if (three) {
                    float4* la = reinterpret_cast<float4*>(lo_ring + (size_t)l * kStage);
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {                    // A then B: 2048 units, 16 per thread
                        float4 v[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) v[q] = ra[ct + 128 * (j + q)];
                        if (do_bias && j < 8) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int b = (j + q) >> 1;
                                asum[4 * b + 0] += v[q].x; asum[4 * b + 1] += v[q].y; asum[4 * b + 2] += v[q].z; asum[4 * b + 3] += v[q].w;
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) la[ct + 128 * (j + q)] = lo4(v[q]);
                    }
                    // writer-side proxy fence: these generic-proxy stores are read by the tensor core (async proxy).  The
                    // converters have nothing else in flight, so it only waits for the 16 stores above; executed by the
                    // MMA thread instead it sat on the MMA issue path of every chunk (~400 cycles, clock64 trace).
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(&conv_bar[l]);
                    if (ct == 0 && c < 12) T2_STAMP(10 + 4 * c);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 v = ra[ct + 128 * j];
                        const int b = j >> 1;
                        asum[4 * b + 0] += v.x; asum[4 * b + 1] += v.y; asum[4 * b + 2] += v.z; asum[4 * b + 3] += v.w;
                    }
                    mbar_arrive(&conv_bar[l]);                           // the raw stage may be recycled once the MMAs are done too
                }
            }
            if (do_bias) {
                // Thread ct holds 16 partial column sums (4 blocks x its logical chunk lc).  Within a warp the four lanes
                // that share lc differ in k-row % 4 (lane bits 3-4) and, through the swizzle, in their physical 32-byte
                // chunk (lane bits 1-2): two xor-butterfly steps (masks 8|2 and 16|4) fold them; lanes 0-7 then hold the
                // warp's sums for lc = lane.  (The first version did 16 shared-memory atomics per thread onto 128
                // addresses: 20 % of the kernel's stall samples in the ncu capture of a weight-gradient launch.)
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    asum[i] += __shfl_xor_sync(0xffffffffu, asum[i], 10);
                    asum[i] += __shfl_xor_sync(0xffffffffu, asum[i], 20);
                }
                const int cw = ct >> 5, cl = ct & 31;
                if (cl < 8) {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
#pragma unroll
                        for (int e = 0; e < 4; ++e) s_colsum[cw][32 * b + 4 * cl + e] = asum[4 * b + e];
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) asum[i] = 0.f;
                asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory");
                const int M = sl.m > 0 ? sl.m : args.M;
                if (w.m0 + ct < M)      // C2 is zero-initialised by the caller
                    atomicAdd(&sl.C2[w.m0 + ct], (s_colsum[0][ct] + s_colsum[1][ct]) + (s_colsum[2][ct] + s_colsum[3][ct]));
                asm volatile("bar.sync 1, %0;" ::"n"(kConvThreads) : "memory");
            }
        }
    } else {
        // ================= epilogue =================
        // warp (q, half): TMEM lanes [32 q, +32) x columns [64 half, +64) of the tile
        const int q = warp & 3, half = (warp - kEpiWarp0) >> 2;
        float* stg = reinterpret_cast<float*>(epi_stage + (size_t)(warp - kEpiWarp0) * kEpiStage);
        const int N = args.N, epi = args.epi;
        int it = 0;
        int last_id = -1;
        for (int id = blockIdx.x; id < args.total; id += gridDim.x) {
            int kb, ke;
            k_range(args, decode(args, id).ky, kb, ke);
            if (kb < ke) last_id = id;
        }
        if (last_id < 0) pdl_trigger();
        for (int id = blockIdx.x; id < args.total; id += gridDim.x) {
            const Work w = decode(args, id);
            {
                int kb, ke;
                k_range(args, w.ky, kb, ke);
                if (kb >= ke) continue;
            }
            const Slot2& sl = args.slot[w.slot];
            const int M = sl.m > 0 ? sl.m : args.M;
            const int acc = it & 1;
            mbar_wait(&acc_full[acc], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (t == kEpiWarp0 * 32 && it < 4) T2_STAMP(64 + 4 * it);
            if (id == last_id) pdl_trigger();     // the next kernel of the stream may set itself up while the last tile is written back
            float* const C = sl.C;
            float* const C2 = sl.C2;
            const float* const Z = sl.Z;
            const float* const bias = sl.bias;
            const int ldc = sl.ldc, ldz = sl.ldz;
            const bool c_vec = al16(C) && (ldc % 4 == 0) && (C2 == nullptr || a_mn || al16(C2)) &&
                               (epi != EPI_MUL_DSILU || (al16(Z) && ldz % 4 == 0));
            const bool splitk = epi == EPI_NONE && args.ksplit > 1;
#pragma unroll 1
            for (int cb = 2 * half; cb < 2 * half + 2; ++cb) {
                float v[32];
                const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * TN + cb * 32);
                if (three) {                      // main + cross-term accumulators: both loads in flight, one wait
                    float vx[32];
                    tmem_ld32_nowait(ta, v);
                    tmem_ld32_nowait(ta + TN, vx);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    reg_fence32(v);
                    reg_fence32(vx);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += vx[i];
                } else {
                    tmem_ld32_nowait(ta, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    reg_fence32(v);
                }
                if (cb == 2 * half + 1) {         // everything of this accumulator is in registers: hand it back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&acc_free[acc]);
                    if (t == kEpiWarp0 * 32 && it < 4) T2_STAMP(65 + 4 * it);
                }
                __syncwarp();                      // previous block's readers are done with the staging buffer
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4)
                    *reinterpret_cast<float4*>(stg + lane * kStageLd + g4 * 4) = make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
                __syncwarp();
                const int n = w.n0 + cb * 32 + (lane & 7) * 4;
                const int nv = min(4, N - n);
                if (nv <= 0) continue;
                const bool vec = c_vec && nv == 4;
                float bias_v[4] = {0.f, 0.f, 0.f, 0.f};
                if ((epi == EPI_BIAS || epi == EPI_BIAS_SILU) && bias)
                    for (int j = 0; j < 4; ++j) if (j < nv) bias_v[j] = bias[n + j];
#pragma unroll 2
                for (int i = 0; i < 8; ++i) {
                    const int r = (lane >> 3) + 4 * i, m = w.m0 + q * 32 + r;
                    if (m >= M) continue;
                    const float4 a4 = *reinterpret_cast<const float4*>(stg + r * kStageLd + (lane & 7) * 4);
                    float x[4] = {a4.x, a4.y, a4.z, a4.w};
                    const size_t ci = (size_t)m * ldc + n;
                    if (splitk) {
                        if (vec) red4(C + ci, x[0], x[1], x[2], x[3]);
                        else for (int j = 0; j < nv; ++j) atomicAdd(&C[ci + j], x[j]);
                        continue;
                    }
                    if (epi == EPI_BIAS || epi == EPI_BIAS_SILU) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) x[j] += bias_v[j];
                        if (epi == EPI_BIAS_SILU) {
                            if (C2) {
                                if (vec) st4(C2 + ci, make_float4(x[0], x[1], x[2], x[3]));
                                else for (int j = 0; j < nv; ++j) C2[ci + j] = x[j];
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) x[j] = silu(x[j]);
                        }
                    } else if (epi == EPI_MUL_DSILU) {
                        const size_t zi = (size_t)m * ldz + n;
                        if (vec) {
                            const float4 z = ld4(Z + zi);
                            x[0] *= dsilu(z.x); x[1] *= dsilu(z.y); x[2] *= dsilu(z.z); x[3] *= dsilu(z.w);
                        } else {
                            for (int j = 0; j < nv; ++j) x[j] *= dsilu(Z[zi + j]);
                        }
                    }
                    if (!C) continue;
                    if (vec) st4(C + ci, make_float4(x[0], x[1], x[2], x[3]));
                    else for (int j = 0; j < nv; ++j) C[ci + j] = x[j];
                }
            }
            if (t == kEpiWarp0 * 32 && it < 4) T2_STAMP(66 + 4 * it);
            ++it;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// ---- tensor maps ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// Tensor maps.  A map depends on (address, shape, box, swizzle) only.
//  * STABLE operands -- anything inside the flat parameter buffer registered with gemm_set_stable_range (the weights: the
//    same addresses every step) -- live in a device-resident table, one per CUDA device: written once (host encode + copy
//    on a private stream, completed before the call returns), never modified; kernels receive the table pointer + index.
//  * every other operand (activations, gradients: workspace addresses and shapes change with every batch) travels BY VALUE
//    in the launch parameters (up to kInline per launch; the encoded maps are cached on the host by key, so a fixed-shape
//    loop does not even re-encode).  A table entry for them would cost an upload + stream synchronisation per new
//    (address, shape) -- hundreds per step as soon as batch shapes vary.
// When a launch needs more by-value maps than kInline, or the table is full, the caller falls back to the round-1 kernel.
constexpr int kTableCap = 1 << 15;
constexpr int kMaxDev = 16;
struct MapKey {
    const void* ptr; int64_t inner, outer, ld; int box_outer, swz32;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_outer == o.box_outer && swz32 == o.swz32;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
        h ^= (size_t)k.inner * 0xC2B2AE3D27D4EB4Full + (size_t)k.outer * 0x165667B19E3779F9ull + (size_t)k.ld * 31 + k.box_outer * 7 + k.swz32;
        return h;
    }
};
struct MapTable {
    CUtensorMap* dev = nullptr;
    CUtensorMap* pinned = nullptr;
    cudaStream_t copy_stream = nullptr;
    int used = 0;
    std::unordered_map<MapKey, int, MapKeyHash> index;
};
std::mutex g_map_mu;
MapTable g_tables[kMaxDev];
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_host_maps;      // encoded maps of non-stable operands (host only)
thread_local const float* g_stable_lo = nullptr;
thread_local const float* g_stable_hi = nullptr;

int encode_map(const float* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer, int swz32, CUtensorMap* out) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return -1; }
    const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swz32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for [%lld x %lld] ld %lld", (int)r, (long long)outer, (long long)inner, (long long)ld);
        return -1;
    }
    return 0;
}

// encoded map of a non-stable operand (host cache, bounded)
int host_map(const float* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer, int swz32, CUtensorMap* out) {
    const MapKey key{ptr, inner, outer, ld, box_outer, swz32};
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_host_maps.find(key);
    if (it != g_host_maps.end()) { *out = it->second; return 0; }
    if (encode_map(ptr, inner, outer, ld, box_outer, swz32, out) != 0) return -1;
    if (g_host_maps.size() > (1u << 14)) g_host_maps.clear();
    g_host_maps.emplace(key, *out);
    return 0;
}

// row-major fp32 [outer][inner] with leading dimension ld; box {32 inner, box_outer rows}, zero OOB fill; swz32: the
// 32-byte-atom 128 B swizzle of the MN-major operand layout, else the plain 128 B swizzle.  Returns the table index
// (>= 0), -1 on error (message set), -2 when the table is full.
int map_index(const float* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer, int swz32, const CUtensorMap** table) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) { set_error("gemm_tc2: bad device"); return -1; }
    std::lock_guard<std::mutex> lk(g_map_mu);
    MapTable& t = g_tables[dev];
    if (!t.dev) {
        if (cudaMalloc(reinterpret_cast<void**>(&t.dev), sizeof(CUtensorMap) * kTableCap) != cudaSuccess ||
            cudaMallocHost(reinterpret_cast<void**>(&t.pinned), sizeof(CUtensorMap)) != cudaSuccess ||
            cudaStreamCreateWithFlags(&t.copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
            set_error("gemm_tc2: cannot allocate the tensor-map table");
            return -1;
        }
    }
    *table = t.dev;
    const MapKey key{ptr, inner, outer, ld, box_outer, swz32};
    auto it = t.index.find(key);
    if (it != t.index.end()) return it->second;
    if (t.used >= kTableCap) return -2;
    if (encode_map(ptr, inner, outer, ld, box_outer, swz32, t.pinned) != 0) return -1;
    const int idx = t.used;
    if (cudaMemcpyAsync(t.dev + idx, t.pinned, sizeof(CUtensorMap), cudaMemcpyHostToDevice, t.copy_stream) != cudaSuccess ||
        cudaStreamSynchronize(t.copy_stream) != cudaSuccess) {
        set_error("gemm_tc2: tensor-map upload failed");
        return -1;
    }
    t.used = idx + 1;
    t.index.emplace(key, idx);
    return idx;
}

bool tma_ok(const float* p, int ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 4 == 0 && p != nullptr; }

}  // namespace

// [lo, lo + n) = the flat parameter buffer of the model call in progress on this host thread (model_forward / _backward)
void gemm_set_stable_range(const float* lo, size_t n_floats) {
    g_stable_lo = lo;
    g_stable_hi = lo ? lo + n_floats : nullptr;
}

int gemm_tc2_max_ctas() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("PAMNET_GEMM_CTAS"); v = e ? atoi(e) : kNumSM; if (v < 1) v = 1; }
    return v;
}

// slot i can go through the TMA kernel (16-byte aligned bases and row strides, full-size slot)
bool gemm_tc2_slot_ok(const GemmArgs& a, int i) {
    const GemmSlot& s = a.slot[i];
    if (s.m > 0 && s.m < 64) return false;
    if (!tma_ok(s.A, s.lda)) return false;
    if (a.nseg == 0 && !tma_ok(s.B, s.ldb)) return false;
    return true;
}

bool gemm_tc2_eligible(const GemmArgs& a) {
    if (a.M < 64 || a.N < 64 || a.K < 16 || a.accumulate) return false;
    if (a.mode == GEMM_NT && a.K > 256) return false;      // one accumulation chain per tile: see gemm.cu
    if (a.nseg > 0) {
        if (a.seg_len % TK != 0) return false;
        for (int s = 0; s < a.nseg; ++s) if (!tma_ok(a.seg_B[s], a.seg_ldb[s])) return false;
    }
    if (a.ksplit > 1) {
        const int chunk = ((a.K + a.ksplit - 1) / a.ksplit + TK - 1) / TK * TK;
        if (a.nseg > 0 && chunk % a.seg_len != 0 && a.seg_len % chunk != 0) return false;
    }
    for (int i = 0; i < a.nslots; ++i) if (!gemm_tc2_slot_ok(a, i)) return false;
    return true;
}

// returns 0 on success, -1 on error, 1 when the tensor-map table is full (caller falls back to the round-1 kernel)
int gemm_tc2_launch(const GemmArgs& a, cudaStream_t st) {
    Args2 b;
    memset(&b, 0, sizeof(b));
    b.M = a.M; b.N = a.N; b.K = a.K; b.mode = a.mode; b.epi = a.epi; b.ksplit = a.ksplit > 1 ? a.ksplit : 1;
    static int n_prod = -1;       // PAMNET_TC2_PROD=3: drop the lo * lo product (see kernel)
    if (n_prod < 0) { const char* e = getenv("PAMNET_TC2_PROD"); n_prod = (e && e[0] == '3') ? 3 : 4; }
    b.nslots = a.nslots; b.precision = a.precision == 1 ? 1 : n_prod; b.nseg = a.nseg; b.seg_len = a.seg_len;
    b.tiles_m = ceil_div(a.M, TM); b.tiles_n = ceil_div(a.N, TN);
    b.total = b.nslots * b.ksplit * b.tiles_m * b.tiles_n;
    const bool a_mn = a.mode == GEMM_TN, b_mn = a.mode != GEMM_NT;
    // map index of an operand: >= 0 table (stable), < 0 by value; -1000 = error, -2000 = by-value room exhausted (split
    // the launch), -3000 = table full (fall back to the round-1 kernel)
    struct Seen { const float* p; int64_t inner, outer, ld; int idx; };
    Seen seen[2 * kGemmMaxSlots + kGemmMaxSeg];
    int n_seen = 0, n_inl = 0;
    auto map_of = [&](const float* p, int64_t inner, int64_t outer, int64_t ld, int box, int swz32) -> int {
        for (int i = n_seen - 1; i >= 0; --i)          // slots of a batched launch mostly share one of their operands
            if (seen[i].p == p && seen[i].inner == inner && seen[i].outer == outer && seen[i].ld == ld) return seen[i].idx;
        int idx;
        if (p >= g_stable_lo && p < g_stable_hi) {
            idx = map_index(p, inner, outer, ld, box, swz32, &b.maps);
            if (idx < 0) return idx == -2 ? -3000 : -1000;
        } else {
            if (n_inl >= kInline) return -2000;
            if (host_map(p, inner, outer, ld, box, swz32, &b.inl[n_inl]) != 0) return -1000;
            idx = -1 - n_inl++;
        }
        seen[n_seen++] = Seen{p, inner, outer, ld, idx};
        return idx;
    };
    for (int s = 0; s < a.nseg; ++s) {
        // segment weight W[out = k][in = n]: MN-major B, rows = seg_len k, cols = N
        const int idx = map_of(a.seg_B[s], a.N, a.seg_len, a.seg_ldb[s], 32, 1);
        if (idx <= -1000) return idx == -1000 ? -1 : 1;
        b.seg_map[s] = idx;
    }
    for (int i = 0; i < a.nslots; ++i) {
        const GemmSlot& s = a.slot[i];
        const int M = s.m > 0 ? s.m : a.M;
        Slot2& d = b.slot[i];
        d.bias = s.bias; d.Z = s.Z; d.C = s.C; d.C2 = s.C2; d.ldc = s.ldc; d.ldz = s.ldz; d.m = s.m;
        // A: K-major [M rows][K] (NT / NN) or MN-major [K rows][M] (TN)
        const int ia = a_mn ? map_of(s.A, M, a.K, s.lda, 32, 1) : map_of(s.A, a.K, M, s.lda, TM, 0);
        int ib = 0;
        if (a.nseg == 0) ib = b_mn ? map_of(s.B, a.N, a.K, s.ldb, 32, 1) : map_of(s.B, a.K, a.N, s.ldb, TN, 0);
        if (ia == -1000 || ib == -1000) return -1;
        if (ia == -3000 || ib == -3000) return 1;
        if (ia == -2000 || ib == -2000) {
            // more distinct activation operands than one launch carries by value: two launches of half the slots each
            if (a.nslots < 2) return 1;
            const int h = a.nslots / 2;
            GemmArgs lo = a, hi = a;
            lo.nslots = h;
            hi.nslots = a.nslots - h;
            for (int j = 0; j < hi.nslots; ++j) hi.slot[j] = a.slot[h + j];
            const int r0 = gemm_tc2_launch(lo, st);
            if (r0 != 0) return r0 < 0 ? r0 : -1;      // (the table cannot fill up half way: by-value operands only)
            const int r1 = gemm_tc2_launch(hi, st);
            if (r1 == 1) set_error("gemm_tc2: tensor-map table filled up in the middle of a split launch");
            return r1 == 0 ? 0 : -1;
        }
        d.map_a = ia; d.map_b = ib;
    }
    PAMNET_TRY(func_smem_once(reinterpret_cast<const void*>(gemm_tc2_kernel), kSmem));
    // Work items per CTA: a persistent CTA holds its SM (216 KB of shared memory) for its whole list, and the layer loop's
    // chain kernels need 78 free SMs every ~25 us.  PAMNET_GEMM_TPC=n caps the list (several waves of short-lived CTAs);
    // measured: no difference at batch 32 (1.53-1.58 ms/step for 1, 2, 4, unlimited), 7.37 vs 6.41 ms/step at batch 256 for
    // 2 vs unlimited -> unlimited by default.
    static int tpc = -1;
    if (tpc < 0) { const char* e = getenv("PAMNET_GEMM_TPC"); tpc = e ? atoi(e) : 0; if (tpc < 1) tpc = 1 << 20; }
    int grid = b.total < gemm_tc2_max_ctas() ? b.total : gemm_tc2_max_ctas();
    const int by_tpc = ceil_div(b.total, tpc);
    if (by_tpc > grid) grid = by_tpc;
    if (pdl_level() == 1) PAMNET_CUDA(launch_pdl(gemm_tc2_kernel, dim3(grid), dim3(kThreads), kSmem, st, b));
    else gemm_tc2_kernel<<<grid, kThreads, kSmem, st>>>(b);
    return 0;
}

}  // namespace pamnet
