// Tensor-core stage interpreter for the row-local node chains (chain.cuh), D >= 64.
//
// Same stage semantics as chain.cu (the FFMA interpreter, still used for D < 64 and with PAMNET_CHAIN=ffma); what
// changes is how a D x D stage is multiplied.  A CTA still owns 8 rows and keeps the activations transposed in shared
// memory ([k][8]), which is exactly the B operand of mma.m16n8k8 (N = the 8 rows).  The problem is computed as
//     Y^T [D features x 8 rows] = A [D x D] * X^T [D x 8],        A[m][k] = the stage's weight matrix
// Warp w owns features 16 w .. 16 w + 15: its A fragments for all D / 8 k-steps were laid out contiguously by the
// weight-preparation kernel (frag_batch below), so a lane reads them with ONE conflict-free 128-bit shared-memory load per
// k-step and every weight element is read from shared memory exactly once per stage (the FFMA kernel read the 64 KB
// matrix twice and then exchanged 8 partial sums through shared memory: three barriers and ~1000 cycles of
// exchange + reduce per stage).  fp32 accuracy comes from the 3xTF32 split done in REGISTERS: hi = rna_tf32(x),
// lo = x - hi; products lo*hi + hi*lo + hi*hi accumulate in fp32 (same scheme as gemm_tc.cu, ~1e-6 relative).  The
// accumulator fragment (features g, g + 8 x rows 2t, 2t + 1) goes straight through the stage epilogue in registers.
//
// Why not tcgen05 here: with 8 .. 16 rows per CTA the UMMA tile would be M = 128 features x N = rows, operands in
// shared memory already split into tf32 planes -- 3 x 64 KB of operand reads and twice the L2 -> SM weight stream per
// stage; the stage is bound by streaming 64 KB of weights into every SM, not by tensor throughput.
#include "chain.cuh"

#include <stdlib.h>

namespace pamnet {
// optional clock64 timeline of CTA 0 / thread 0 (-DPAMNET_TC_TRACE builds; tools/chain_trace.py): 8 stamps per stage
#ifdef PAMNET_TC_TRACE
__device__ long long g_chain_mma_trace[26 * 8];
#define CM_STAMP(si, i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (si) < 26) g_chain_mma_trace[(si) * 8 + (i)] = clock64(); } while (0)
#else
#define CM_STAMP(si, i) do { } while (0)
#endif
int chain_mma_trace_read(long long* out, int n) {
#ifdef PAMNET_TC_TRACE
    PAMNET_CUDA(cudaMemcpyFromSymbol(out, g_chain_mma_trace, sizeof(long long) * (n < 208 ? n : 208)));
    return 0;
#else
    (void)out; (void)n;
    set_error("built without PAMNET_TC_TRACE");
    return -1;
#endif
}
namespace {

constexpr int kR = 8;     // rows per CTA = N of the MMA

__device__ __forceinline__ unsigned smem_addr_(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr_(dst)), "l"(src), "r"(bytes), "r"(smem_addr_(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t* bar, unsigned parity) {
    unsigned ok = 0;
    for (unsigned spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(smem_addr_(bar)), "r"(parity) : "memory");
        if (spin > (1u << 24)) asm volatile("trap;");      // a protocol bug must not hang the GPU
    }
}

// D (16x8, fp32) += A (16x8, tf32, row) * B (8x8, tf32, col)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// x = hi + lo + O(2^-21 |x|): hi rounded to nearest tf32 with integer ops, lo exact in fp32 (the tensor core drops its
// low mantissa bits: 2^-10 * 2^-11)
__device__ __forceinline__ void split_tf32_(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

template <int D, bool TP>
struct MmaChainCfg {
    static constexpr int WM = D / 16;                      // 16-feature M tiles
    static constexpr int CW = TP ? WM : 2 * WM;            // consumer warps: M tile (x k half in the latency variant)
    static constexpr int NC = CW * 32;                     // consumer threads
    static constexpr int NT = NC + 32;                     // + the weight-producer warp
    static constexpr int KS = D / 8;                       // k-steps per stage
    static constexpr int NBUF = TP ? 1 : 2;                // weight buffers
    static constexpr size_t smem_floats = (size_t)NBUF * D * D + (size_t)(kChainSlots + 4) * D * kR;
};

// ---- gather stages (kept out of line: their row registers must not constrain the GEMM stage's allocation) ----------
template <int D, int CW>
__device__ __noinline__ void gmsg_fwd_stage(const ChainStage& st, float* slots, int row0, int n_rows, int warp, int lane,
                                            int er, int ec) {
    constexpr int R = kR, EC = CW * 32 / R;
    auto slot_ptr = [&](int s) -> float* { return slots + s * D * R; };
            // Two warps per node when there are 16 consumer warps (each takes every other incoming edge), one otherwise;
            // a lane owns 4 (D = 128) / 2 (D = 64) columns, every row access is one coalesced request, kU rows in flight.
            constexpr int kU = 2, WPN = CW / R;                    // WPN = warps per node (2 or 1); kU x 3 rows in flight per warp
            float* part = slot_ptr(kChainWide);                  // [WPN][R][D] partial sums (the wide slot is free in forward chains)
            const int r = warp % R, wsub = warp / R;
            const int n = row0 + r;
            RowVec<D> acc;
            acc.zero();
            if (n < n_rows) {
                const float* P = st.g0; const float* QT = st.W; const int ldq = st.ldw;
                const int32_t* srcs = st.i_src;
                RowVec<D> pi;
                pi.load(P + (size_t)n * 2 * D, lane);
                const int e0 = st.i_ptr[n], e1 = st.i_ptr[n + 1];
                for (int k = e0 + wsub; k < e1; k += WPN * kU) {
                    RowVec<D> pj[kU], q[kU], tt[kU];
#pragma unroll
                    for (int u = 0; u < kU; ++u) {
                        const int kk = k + u * WPN;
                        if (kk < e1) {
                            pj[u].load(P + (size_t)srcs[kk] * 2 * D + D, lane);
                            q[u].load(QT + (size_t)kk * ldq, lane);
                            tt[u].load(QT + (size_t)kk * ldq + D, lane);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kU; ++u)
                        if (k + u * WPN < e1) {
#pragma unroll
                            for (int i = 0; i < RowVec<D>::C * RowVec<D>::V; ++i) acc.v[i] += silu(pi.v[i] + pj[u].v[i] + q[u].v[i]) * tt[u].v[i];
                        }
                }
            }
            acc.store(part + ((size_t)wsub * R + r) * D, lane);
            __syncthreads();
            {   // combine (fixed order), + x1, -> slot (transposed) and global h
                float* d = slot_ptr(st.dst);
                const bool live = row0 + er < n_rows;
                const float* x1 = st.g1; float* out_a = st.out_a;
                for (int c4 = ec; c4 < D / 4; c4 += EC) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live) {
                        v = ld4(x1 + (size_t)(row0 + er) * D + c4 * 4);
#pragma unroll
                        for (int ws = 0; ws < WPN; ++ws) v = v + ld4(part + ((size_t)ws * R + er) * D + c4 * 4);
                        if (out_a) st4(out_a + (size_t)(row0 + er) * D + c4 * 4, v);
                    }
                    float* qd = d + (c4 * 4) * R + er;
                    qd[0] = v.x; qd[R] = v.y; qd[2 * R] = v.z; qd[3 * R] = v.w;
                }
            }
            __syncthreads();
}

template <int D, int CW>
__device__ __noinline__ void gather_stage(const ChainStage& st, float* slots, int row0, int n_rows, int warp, int lane) {
    constexpr int R = kR;
    auto slot_ptr = [&](int s) -> float* { return slots + s * D * R; };
            // one warp per (node, task): LMSG_FWD has one task per node (h = x1 + sum msum * Rout), GATHER_BWD 2 * n_blocks
            // (block b of the per-edge gradient rows over incoming / outgoing slots); results go straight to the slot
            constexpr int kU = 2;
            const bool fwd = st.op == CH_LMSG_FWD;
            const int ntask = fwd ? 1 : st.width / D;
            float* d = slot_ptr(st.dst);
            const float* rows = st.W; const int ldr = st.ldw;
            float* out_a = st.out_a; const int ld_out = fwd ? D : st.width;
            for (int item = warp; item < R * ntask; item += CW) {
                const int r = item % R, task = item / R;
                const int n = row0 + r;
                RowVec<D> s;
                s.zero();
                if (n < n_rows) {
                    if (fwd) {
                        const float* msum = st.g0;
                        s.load(st.g1 + (size_t)n * D, lane);
                        for (int k = st.i_ptr[n], k1 = st.i_ptr[n + 1]; k < k1; k += kU) {
                            RowVec<D> ms[kU], ro[kU];
#pragma unroll
                            for (int u = 0; u < kU; ++u)
                                if (k + u < k1) { ms[u].load(msum + (size_t)(k + u) * D, lane); ro[u].load(rows + (size_t)(k + u) * ldr, lane); }
#pragma unroll
                            for (int u = 0; u < kU; ++u)
                                if (k + u < k1) {
#pragma unroll
                                    for (int i = 0; i < RowVec<D>::C * RowVec<D>::V; ++i) s.v[i] += ms[u].v[i] * ro[u].v[i];
                                }
                        }
                    } else {
                        const int b = task >> 1;
                        const bool outgoing = task & 1;
                        const int32_t* ptr = outgoing ? st.o_ptr : st.i_ptr;
                        const int32_t* opos = st.o_pos;
                        for (int k = ptr[n], k1 = ptr[n + 1]; k < k1; k += kU) {
                            RowVec<D> v[kU];
#pragma unroll
                            for (int u = 0; u < kU; ++u)
                                if (k + u < k1) v[u].load(rows + (size_t)(outgoing ? opos[k + u] : k + u) * ldr + b * D, lane);
#pragma unroll
                            for (int u = 0; u < kU; ++u)
                                if (k + u < k1) {
#pragma unroll
                                    for (int i = 0; i < RowVec<D>::C * RowVec<D>::V; ++i) s.v[i] += v[u].v[i];
                                }
                        }
                    }
                    if (out_a) s.store(out_a + (size_t)n * ld_out + task * D, lane);
                }
                // transposed slot: column c of this task -> d[(task * D + c) * R + r]
#pragma unroll
                for (int cc = 0; cc < RowVec<D>::C; ++cc)
#pragma unroll
                    for (int vv = 0; vv < RowVec<D>::V; ++vv)
                        d[(size_t)(task * D + (cc * 32 + lane) * RowVec<D>::V + vv) * R + r] = s.v[cc * RowVec<D>::V + vv];
            }
            __syncthreads();
}

// Two variants of one kernel.
//  * latency (TP = false; up to ~2 waves of 8-row CTAs, the batch-32 case): consumer warp w = (M tile w % WM, k half
//    w / WM) multiplies its 16 features x 8 rows over half of K; the two warps of an M tile swap the accumulator halves
//    they do NOT finish through shared memory (one named barrier per pair) so that every thread ends up with 2 of the
//    tile's elements.  Weights are double-buffered: the next stage's image streams in behind the current multiply.
//  * throughput (TP = true; thousands of rows, configs[2]): one warp per M tile over all of K, 4 elements per thread,
//    ONE weight buffer and <= 112 registers so that TWO CTAs share an SM: a stage is ~1500 cycles of tensor pipe and
//    ~1600 cycles of everything else (clock64 trace), and the co-resident CTA's multiply fills the other half.
// The producer warp only walks the stage table in lockstep (same CTA barriers) and asks the bulk-copy engine for the
// NEXT GEMM stage's fragment image: the request (~230 cycles of one thread) used to sit on thread 0's path to the
// multiply loop.
template <int D, bool TP>
__global__ void __launch_bounds__(MmaChainCfg<D, TP>::NT, TP ? 2 : 1) chain_mma_kernel(const ChainArgs args) {
    using C = MmaChainCfg<D, TP>;
    constexpr int R = kR, NC = C::NC, NT = C::NT, KS = C::KS, WM = C::WM, CW = C::CW, NBUF = C::NBUF;
    constexpr int KSW = TP ? KS : KS / 2;                 // k-steps per consumer warp
    constexpr int NE = TP ? 4 : 2;                        // finished elements per thread
    extern __shared__ __align__(16) float smem[];
    float* wbuf = smem;                                   // [NBUF][D * D]: fragment images of this (and the next) GEMM stage
    float* slots = wbuf + NBUF * D * D;                   // 3 x [D][R] transposed activations, then wide [4D][R]
    auto slot_ptr = [&](int s) -> float* { return slots + s * D * R; };

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int n_rows = args.n_rows;
    const int row0 = blockIdx.x * R;

    __shared__ __align__(16) ChainStage s_stage[kChainMaxStages];
    __shared__ __align__(8) uint64_t wbar[2];
    __shared__ float s_red[2 * R * CW];
    __shared__ __align__(8) float2 s_xch[TP ? 1 : CW * 32];
    {
        static_assert(sizeof(ChainStage) % 4 == 0, "word copy");
        const int nwords = args.n_stages * (int)(sizeof(ChainStage) / 4);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(args.st);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_stage);
        for (int i = t; i < nwords; i += NT) dst[i] = src[i];
    }
    if (t == 0) {
        mbar_init_(&wbar[0], 1);
        mbar_init_(&wbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_stages = args.n_stages;

    if (warp == CW) {
        // ---- producer warp: same CTA-barrier sequence as the consumers, nothing else ------------------------------
        auto issue_weights = [&](int si, int buf) {           // one bulk copy: the image is contiguous
            if (lane == 0) {
                mbar_expect_tx_(&wbar[buf], (unsigned)(D * D * sizeof(float)));
                bulk_g2s_(wbuf + buf * D * D, s_stage[si].W, (unsigned)(D * D * sizeof(float)), &wbar[buf]);
            }
        };
        int wcur = 0;
        {
            int first = -1;
            for (int i = 0; i < n_stages && first < 0; ++i)
                if (s_stage[i].op == CH_GEMM) first = i;
            if (first >= 0) issue_weights(first, 0);          // prepared weights: written before the predecessor started
        }
        for (int si = 0; si < n_stages; ++si) {
            const int op = s_stage[si].op;
            if (op == CH_GEMM) {
                const int nxt = s_stage[si].next_gemm;
                // two buffers: buffer wcur ^ 1 was last read by the previous GEMM stage, whose trailing barrier this warp
                // has passed; one buffer: it is free only after THIS stage's trailing barrier
                if (NBUF == 2 && nxt >= 0) issue_weights(nxt, wcur ^ 1);
                if (s_stage[si].psrc >= 0) __syncthreads();
                if (NBUF == 2) wcur ^= 1;
                __syncthreads();
                if (NBUF == 1 && nxt >= 0) issue_weights(nxt, 0);
            } else {
                __syncthreads();
                if (op == CH_DOT2 || op == CH_GMSG_FWD) __syncthreads();
            }
        }
        return;
    }

    // ---- consumer warps ---------------------------------------------------------------------------------------------
    const int g = lane >> 2, tq = lane & 3;
    const int mt = warp % WM, kh = TP ? 0 : warp / WM;
    const int f0 = 16 * mt + g;                           // fragment features f0, f0 + 8; rows 2 tq, 2 tq + 1
    const int fe = f0 + 8 * kh;                           // latency variant: the feature this thread finishes
    const int r0 = 2 * tq;
    // finished element e: feature fe_(e), row r0 + (e & 1)
    auto fe_ = [&](int e) { return TP ? f0 + 8 * (e >> 1) : fe; };
    const int er = t & (R - 1), ec = t / R;               // element-wise stages: (row, 4-column group), row fastest
    constexpr int EC = NC / R;
    const bool live_r[2] = {row0 + r0 < n_rows, row0 + r0 + 1 < n_rows};

    unsigned wphase[2] = {0u, 0u};
    float4 zpre = make_float4(0.f, 0.f, 0.f, 0.f);
    int zpre_stage = -1;
    int wcur = 0;
    // Everything above touched only kernel parameters and prepared weights; from here on the predecessor's outputs are read.
    pdl_wait();

    // A gather stage is always stage 0: it runs before the stage loop, out of line, so that its row registers and the call
    // do not touch the loop's register allocation (inside the loop the call sites cost 144 B of spills and 60 % of the
    // kernel's speed).
    int si0 = 0;
    if (s_stage[0].op == CH_GMSG_FWD) { gmsg_fwd_stage<D, CW>(s_stage[0], slots, row0, n_rows, warp, lane, er, ec); si0 = 1; }
    else if (s_stage[0].op == CH_LMSG_FWD || s_stage[0].op == CH_GATHER_BWD) { gather_stage<D, CW>(s_stage[0], slots, row0, n_rows, warp, lane); si0 = 1; }
    for (int si = si0; si < n_stages; ++si) {
        const ChainStage& st = s_stage[si];
        if (si == n_stages - 1) pdl_trigger();
        CM_STAMP(si, 0);
        if (st.op == CH_LOAD) {
            float* d = slot_ptr(st.dst);
            const int w4 = st.width / 4;
            const bool live = row0 + er < n_rows;
            const float* g0 = st.g0; const float* g1 = st.g1; float* out_a = st.out_a;
            const int ld_g = st.ld_g, ld_out = st.ld_out, add_slot = st.add_slot;
            for (int c4 = ec; c4 < w4; c4 += EC) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) {
                    v = ld4(g0 + (size_t)(row0 + er) * ld_g + c4 * 4);
                    if (g1) v = v + ld4(g1 + (size_t)(row0 + er) * ld_g + c4 * 4);
                    if (add_slot >= 0) {
                        const float* a = slot_ptr(add_slot) + (c4 * 4) * R + er;
                        v = v + make_float4(a[0], a[R], a[2 * R], a[3 * R]);
                    }
                    if (out_a) st4(out_a + (size_t)(row0 + er) * ld_out + c4 * 4, v);
                }
                float* q = d + (c4 * 4) * R + er;
                q[0] = v.x; q[R] = v.y; q[2 * R] = v.z; q[3 * R] = v.w;
            }
            __syncthreads();
        } else if (st.op == CH_HEADS_BWD) {
            // grad of o3 from the two heads: g_att * W + g_out * W_out.weight.  Optionally (zmul = the saved o3 activations)
            // also the heads' own weight gradients dW = o3^T g_att, dW_out = o3^T g_out, db_out = sum g_out: the CTA folds
            // its 8 rows with shuffles and adds one partial per column to global memory (they used to be two one-row
            // slots of the node-level weight-gradient GEMM launch).
            float* d = slot_ptr(st.dst);
            const bool live = row0 + er < n_rows;
            const float ga = live ? st.g0[row0 + er] : 0.f, go = live ? st.g1[row0 + er] : 0.f;
            const float* W = st.W; const float* Wo = st.bias; const float* o3 = st.zmul;
            float* gW = st.out_z; float* gWo = st.out_a; float* gbo = st.save_src;
            for (int c4 = ec; c4 < D / 4; c4 += EC) {
                const float4 w = ld4(W + c4 * 4), wo = ld4(Wo + c4 * 4);
                float* q = d + (c4 * 4) * R + er;
                q[0] = ga * w.x + go * wo.x; q[R] = ga * w.y + go * wo.y;
                q[2 * R] = ga * w.z + go * wo.z; q[3 * R] = ga * w.w + go * wo.w;
                if (o3) {
                    const float4 a = live ? ld4(o3 + (size_t)(row0 + er) * D + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float pa[4] = {ga * a.x, ga * a.y, ga * a.z, ga * a.w}, po[4] = {go * a.x, go * a.y, go * a.z, go * a.w};
#pragma unroll
                    for (int o = 1; o < R; o <<= 1)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            pa[j] += __shfl_xor_sync(0xffffffffu, pa[j], o);
                            po[j] += __shfl_xor_sync(0xffffffffu, po[j], o);
                        }
                    if (er == 0) {
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gW + c4 * 4), "f"(pa[0]), "f"(pa[1]), "f"(pa[2]), "f"(pa[3]) : "memory");
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gWo + c4 * 4), "f"(po[0]), "f"(po[1]), "f"(po[2]), "f"(po[3]) : "memory");
                    }
                }
            }
            if (o3 && gbo && ec == 0) {          // ec == 0: the 8 lanes t = 0..7
                float sgo = go;
#pragma unroll
                for (int o = 1; o < R; o <<= 1) sgo += __shfl_xor_sync(0x000000ffu, sgo, o);
                if (er == 0) atomicAdd(gbo, sgo);
            }
            __syncthreads();
        } else if (st.op == CH_DOT2) {
            const float* s = slot_ptr(st.src);
            const float* W = st.W; const float* Wo = st.bias;
            float a = 0.f, o = 0.f;
            for (int cc = ec; cc < D; cc += EC) {
                const float x = s[cc * R + er];
                a = fmaf(x, W[cc], a);
                o = fmaf(x, Wo[cc], o);
            }
            a += __shfl_xor_sync(0xffffffffu, a, 8);  o += __shfl_xor_sync(0xffffffffu, o, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 16); o += __shfl_xor_sync(0xffffffffu, o, 16);
            if (lane < R) { s_red[(warp * R + lane) * 2] = a; s_red[(warp * R + lane) * 2 + 1] = o; }
            __syncthreads();
            if (t < R && row0 + t < n_rows) {
                float sa = 0.f, so = 0.f;
                for (int w = 0; w < CW; ++w) { sa += s_red[(w * R + t) * 2]; so += s_red[(w * R + t) * 2 + 1]; }
                st.out_z[row0 + t] = sa;
                st.out_a[row0 + t] = so + st.g0[0];
            }
            __syncthreads();
        } else {  // CH_GEMM
            // Stage fields copied to registers up front: `st` lives in shared memory and every global store may alias it
            // as far as the compiler knows, so reading fields between the stores reloads them and serialises the
            // independent element chains (ncu: 42 % of the first version's samples sat in the epilogue).
            float* const out_z = st.out_z;
            float* const out_a = st.out_a;
            float* const post_save = st.post_save;
            const float* const add_g = st.add_g;
            const float* const bias = st.bias;
            const int ld_out = st.ld_out, act = st.act, add_slot = st.add_slot, dst = st.dst, post_dst = st.post_dst;
            const int psrc = st.psrc, nxt = st.next_gemm;
            const float* in = slot_ptr(st.src) + st.src_off * R;
            // epilogue operands requested now, consumed after the k-loop
            float bias_v[NE / 2], addg_v[NE], zpost_v[NE];
#pragma unroll
            for (int e = 0; e < NE; ++e) { addg_v[e] = 0.f; zpost_v[e] = 0.f; }
#pragma unroll
            for (int h = 0; h < NE / 2; ++h) bias_v[h] = bias ? bias[fe_(2 * h)] : 0.f;
            if (add_g) {
                const int ld_add = st.ld_add;
#pragma unroll
                for (int e = 0; e < NE; ++e)
                    if (live_r[e & 1]) addg_v[e] = add_g[(size_t)(row0 + r0 + (e & 1)) * ld_add + fe_(e)];
            }
            if (post_dst >= 0) {
                const float* pz = st.post_zmul;
#pragma unroll
                for (int e = 0; e < NE; ++e)
                    if (live_r[e & 1]) zpost_v[e] = pz[(size_t)(row0 + r0 + (e & 1)) * D + fe_(e)];
            }
            if (psrc >= 0) {          // prologue: src * SiLU'(zmul) -> psrc (and to global for the weight gradients)
                float* p = slot_ptr(psrc);
                const bool live = row0 + er < n_rows;
                const float* zmul = st.zmul;
                float* save_src = st.save_src;
                for (int c4 = ec; c4 < D / 4; c4 += EC) {
                    const float* gsrc = in + (c4 * 4) * R + er;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live) {
                        const float4 zz = (zpre_stage == si && c4 == ec) ? zpre : ld4(zmul + (size_t)(row0 + er) * D + c4 * 4);
                        const float4 dz = dsilu4(zz);
                        v = make_float4(gsrc[0] * dz.x, gsrc[R] * dz.y, gsrc[2 * R] * dz.z, gsrc[3 * R] * dz.w);
                        if (save_src) st4(save_src + (size_t)(row0 + er) * D + c4 * 4, v);
                    }
                    float* q = p + (c4 * 4) * R + er;
                    q[0] = v.x; q[R] = v.y; q[2 * R] = v.z; q[3 * R] = v.w;
                }
                in = p;
            }
            if (nxt >= 0 && s_stage[nxt].psrc >= 0 && ec < D / 4 && row0 + er < n_rows) {
                zpre = ld4(s_stage[nxt].zmul + (size_t)(row0 + er) * D + ec * 4);
                zpre_stage = nxt;
            }
            float addv[NE];                            // the residual slot was complete before this stage began
#pragma unroll
            for (int e = 0; e < NE; ++e) addv[e] = 0.f;
            if (add_slot >= 0) {
#pragma unroll
                for (int h = 0; h < NE / 2; ++h) {
                    const float2 a2 = *reinterpret_cast<const float2*>(slot_ptr(add_slot) + fe_(2 * h) * R + r0);
                    addv[2 * h] = a2.x; addv[2 * h + 1] = a2.y;
                }
            }
            CM_STAMP(si, 1);
            mbar_wait_(&wbar[wcur], wphase[wcur]);        // this stage's weights have landed
            wphase[wcur] ^= 1u;
            CM_STAMP(si, 2);
            if (psrc >= 0) __syncthreads();               // prologue visible (otherwise the previous stage's trailing barrier covers the inputs)
            CM_STAMP(si, 3);

            // ---- multiply over this warp's half of K: 3 products per k-step, six independent accumulator chains ----
            float acc_m[2][4], acc_x[2][4], acc_y[2][4];
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc_m[p][j] = acc_x[p][j] = acc_y[p][j] = 0.f;
            const float4* wf = reinterpret_cast<const float4*>(wbuf + wcur * D * D) + ((size_t)mt * KS + kh * KSW) * 32 + lane;
            const float* bp = in + (kh * KSW * 8 + tq) * R + g;
            if (args.precision == 1) {
                // single-pass TF32: the tensor core reads the fp32 bit patterns and ignores the 13 low mantissa bits
#pragma unroll
                for (int s = 0; s < KSW; ++s) {
                    const float4 av = wf[s * 32];
                    const uint32_t a4[4] = {__float_as_uint(av.x), __float_as_uint(av.y), __float_as_uint(av.z), __float_as_uint(av.w)};
                    const uint32_t b2[2] = {__float_as_uint(bp[(8 * s) * R]), __float_as_uint(bp[(8 * s + 4) * R])};
                    mma_tf32(acc_m[s & 1], a4, b2);
                }
            } else
#pragma unroll (TP ? 8 : KSW)
            for (int s = 0; s < KSW; ++s) {
                const float4 av = wf[s * 32];
                const float b0 = bp[(8 * s) * R], b1 = bp[(8 * s + 4) * R];
                uint32_t ah[4], al[4], bh[2], bl[2];
                split_tf32_(av.x, ah[0], al[0]); split_tf32_(av.y, ah[1], al[1]);
                split_tf32_(av.z, ah[2], al[2]); split_tf32_(av.w, ah[3], al[3]);
                split_tf32_(b0, bh[0], bl[0]); split_tf32_(b1, bh[1], bl[1]);
                mma_tf32(acc_x[s & 1], al, bh);
                mma_tf32(acc_y[s & 1], ah, bl);
                mma_tf32(acc_m[s & 1], ah, bh);
            }
            float c[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)     // small terms first
                c[j] = ((acc_x[0][j] + acc_x[1][j]) + (acc_y[0][j] + acc_y[1][j])) + (acc_m[0][j] + acc_m[1][j]);
            CM_STAMP(si, 4);
            // ---- latency variant: swap halves with the partner warp (other k half of the same M tile), keep feature fe ----
            if (!TP) {
                const float2 give = kh ? make_float2(c[0], c[1]) : make_float2(c[2], c[3]);
                s_xch[warp * 32 + lane] = give;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + mt) : "memory");
                const float2 got = s_xch[(warp ^ WM) * 32 + lane];     // WM is a power of two: partner = other k half
                // fixed order (k half 0 + k half 1) -> deterministic
                if (kh) { c[0] = got.x + c[2]; c[1] = got.y + c[3]; }
                else    { c[0] = c[0] + got.x; c[1] = c[1] + got.y; }
            }
            // ---- epilogue: NE elements (feature fe_(e), row r0 + (e & 1)), every operand already in registers ----
            float x[NE], z[NE];
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                z[e] = c[e] + bias_v[e >> 1];
                x[e] = act ? silu(z[e]) : z[e];
                x[e] = live_r[e & 1] ? (x[e] + addv[e]) + addg_v[e] : 0.f;
            }
            if (x[0] == 12345.678f) CM_STAMP(si, 7);   // (forces the math to complete before the next stamp)
            CM_STAMP(si, 5);
            if (out_z) {
#pragma unroll
                for (int e = 0; e < NE; ++e)
                    if (live_r[e & 1]) out_z[(size_t)(row0 + r0 + (e & 1)) * ld_out + fe_(e)] = z[e];
            }
            if (out_a) {
#pragma unroll
                for (int e = 0; e < NE; ++e)
                    if (live_r[e & 1]) out_a[(size_t)(row0 + r0 + (e & 1)) * ld_out + fe_(e)] = x[e];
            }
            if (dst >= 0 && dst != post_dst) {
#pragma unroll
                for (int h = 0; h < NE / 2; ++h)
                    *reinterpret_cast<float2*>(slot_ptr(dst) + fe_(2 * h) * R + r0) = make_float2(x[2 * h], x[2 * h + 1]);
            }
            if (post_dst >= 0) {          // the next stage's prologue, on register values
                float y[NE];
#pragma unroll
                for (int e = 0; e < NE; ++e) y[e] = live_r[e & 1] ? x[e] * dsilu(zpost_v[e]) : 0.f;
                if (post_save) {
#pragma unroll
                    for (int e = 0; e < NE; ++e)
                        if (live_r[e & 1]) post_save[(size_t)(row0 + r0 + (e & 1)) * D + fe_(e)] = y[e];
                }
#pragma unroll
                for (int h = 0; h < NE / 2; ++h)
                    *reinterpret_cast<float2*>(slot_ptr(post_dst) + fe_(2 * h) * R + r0) = make_float2(y[2 * h], y[2 * h + 1]);
            }
            CM_STAMP(si, 6);
            if (NBUF == 2) wcur ^= 1;
            __syncthreads();
        }
        CM_STAMP(si, 7);
    }
}

// A[m][k] of a D x D stage -> fragment image: float4 index (w * (D / 8) + s) * 32 + lane holds
// { A[16w+g][8s+t], A[16w+g+8][8s+t], A[16w+g][8s+t+4], A[16w+g+8][8s+t+4] },  g = lane / 4, t = lane % 4
constexpr int kFragJobs = 128;
struct FragArgs {
    int dim;
    int src_off[kFragJobs], dst_off[kFragJobs], ld[kFragJobs];
    char trans[kFragJobs];          // 0: A[m][k] = src[m * ld + k];  1: A[m][k] = src[k * ld + m]
};
__global__ void __launch_bounds__(256) frag_kernel(const float* __restrict__ src_base, float* __restrict__ dst_base,
                                                   const FragArgs a) {
    const int job = blockIdx.y, D = a.dim, ks = D / 8;
    const float* src = src_base + a.src_off[job];
    float4* dst = reinterpret_cast<float4*>(dst_base + a.dst_off[job]);
    const int ld = a.ld[job];
    const bool tr = a.trans[job] != 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D * D / 4; i += gridDim.x * blockDim.x) {
        const int lane = i & 31, s = (i >> 5) % ks, w = (i >> 5) / ks;
        const int m = 16 * w + (lane >> 2), k = 8 * s + (lane & 3);
        auto A = [&](int mm, int kk) { return tr ? src[(size_t)kk * ld + mm] : src[(size_t)mm * ld + kk]; };
        dst[i] = make_float4(A(m, k), A(m + 8, k), A(m, k + 4), A(m + 8, k + 4));
    }
}

template <int D, bool TP>
int chain_mma_launch_t(const ChainArgs& a, double bytes, cudaStream_t st) {
    using C = MmaChainCfg<D, TP>;
    const size_t smem = C::smem_floats * sizeof(float);
    PAMNET_TRY(func_smem_once(reinterpret_cast<const void*>(chain_mma_kernel<D, TP>), smem));
    prof_begin(KC_CHAIN, bytes, st);
    launch_pdl(chain_mma_kernel<D, TP>, dim3(ceil_div(a.n_rows, kR)), dim3(C::NT), smem, st, a);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// throughput variant (two CTAs per SM) from ~2 waves of 8-row CTAs on; PAMNET_CHAIN_TP=0 / 1 forces either
bool chain_throughput(int n_rows) {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("PAMNET_CHAIN_TP"); mode = e ? (e[0] == '1' ? 1 : 0) : 2; }
    if (mode != 2) return mode == 1;
    return n_rows > 2 * kNumSM * kR;
}

}  // namespace

bool chain_mma_enabled(int dim) {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("PAMNET_CHAIN"); mode = (e && strcmp(e, "ffma") == 0) ? 0 : 1; }
    return mode == 1 && (dim == 128 || dim == 64);
}

int chain_mma_launch(int dim, const ChainArgs& a, double bytes, cudaStream_t st) {
    switch (dim) {
        case 128: return (a.small_footprint || chain_throughput(a.n_rows)) ? chain_mma_launch_t<128, true>(a, bytes, st) : chain_mma_launch_t<128, false>(a, bytes, st);
        case 64:  return (a.small_footprint || chain_throughput(a.n_rows)) ? chain_mma_launch_t<64, true>(a, bytes, st) : chain_mma_launch_t<64, false>(a, bytes, st);
        default:
            set_error("chain_mma: unsupported dim %d (64, 128)", dim);
            return -1;
    }
}

int frag_batch(const float* src_base, float* dst_base, int dim, const FragJob* jobs, int n_jobs, cudaStream_t st) {
    for (int j0 = 0; j0 < n_jobs; j0 += kFragJobs) {
        FragArgs a;
        a.dim = dim;
        const int n = (n_jobs - j0 < kFragJobs) ? n_jobs - j0 : kFragJobs;
        for (int j = 0; j < n; ++j) {
            const FragJob& jb = jobs[j0 + j];
            a.src_off[j] = (int)jb.src_off; a.dst_off[j] = (int)jb.dst_off; a.ld[j] = jb.ld; a.trans[j] = (char)(jb.trans != 0);
        }
        dim3 grid(ceil_div(dim * dim / 4, 256), n);
        prof_begin(KC_BASIS, 0.0, st);
        frag_kernel<<<grid, 256, 0, st>>>(src_base, dst_base, a);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace pamnet
