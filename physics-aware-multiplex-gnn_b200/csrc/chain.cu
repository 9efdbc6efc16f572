// Stage interpreter for row-local node chains (see chain.cuh).  One CTA = 8 rows; activations live (transposed)
// in shared memory across all stages; whole D x D weight matrices are streamed k-major through a double buffer
// filled by the bulk-copy engine, the next stage's matrix in flight while the current one multiplies.
#include "chain.cuh"

#include <stdlib.h>

namespace pamnet {

// optional clock64 timeline of CTA 0 / thread 0 (-DPAMNET_TC_TRACE builds; tools/chain_trace.py): 8 stamps per stage
#ifdef PAMNET_TC_TRACE
__device__ long long g_chain_trace[16 * 16];
#define CH_STAMP(si, i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (si) < 16) g_chain_trace[(si) * 16 + (i)] = clock64(); } while (0)
#else
#define CH_STAMP(si, i) do { } while (0)
#endif
int chain_mma_trace_read(long long* out, int n);
int chain_trace_read(long long* out, int n) {
    if (chain_mma_enabled(128)) return chain_mma_trace_read(out, n);
#ifdef PAMNET_TC_TRACE
    PAMNET_CUDA(cudaMemcpyFromSymbol(out, g_chain_trace, sizeof(long long) * (n < 256 ? n : 256)));
    return 0;
#else
    (void)out; (void)n;
    set_error("built without PAMNET_TC_TRACE");
    return -1;
#endif
}

constexpr int kChainRows = 8;     // rows per CTA
constexpr int kChainKS = 8;       // k-slices: a CTA has kChainKS * 2 * D / 4 threads (4 x 4 tiles, two row groups)

// ---- weight streaming with the bulk-copy (TMA) engine --------------------------------------------------------------
// A D x D weight matrix per stage is 64 KB at D = 128.  Copying it with per-thread cp.async cost 8 LDGSTS.128 per
// thread and stage -- more load/store-unit time than the multiply loop's own shared-memory reads.  One warp now asks
// the bulk-copy engine for the rows (a single request when the matrix is contiguous) and every thread waits on the
// mbarrier whose transaction count the copies complete.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok = 0;
    for (unsigned spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
        if (spin > (1u << 24)) asm volatile("trap;");      // a protocol bug must not hang the GPU
    }
}

// Multiply loop: thread (cg, rg, ks) owns a 4 x 4 tile -- output columns 4 cg .. 4 cg + 3 (cg = tid % (D/4)), rows
// 4 rg .. 4 rg + 3 (rg = 0, 1) -- over k-slice ks of 8.  Per k one conflict-free 128-bit weight load (a warp reads a
// whole 512 B weight row at D = 128) and one broadcast 128-bit load of four row values (activations are kept
// TRANSPOSED in shared memory, [k][8]) feed 16 independent FMAs.
// Reduction: the eight partial sums meet in shared memory as [ks][column][row half] 16-byte chunks (XOR-swizzled so
// that both the writers' and the readers' accesses are conflict-free); thread t then finishes two rows of one column
// with eight 64-bit loads and stores its result as one 64-bit word of the transposed activation slot.
// 4*D threads per CTA (16 warps at D = 128): everything outside the multiply loop is a chain of dependent latencies
// (ncu: 78 % of the stall samples), which four warps per scheduler hide better than two at the same FMA issue load.
template <int D>
struct ChainCfg {
    static constexpr int R = kChainRows;
    static constexpr int T = kChainKS * D / 2;           // threads
    static constexpr int KL = D / kChainKS;              // k per slice
    static constexpr size_t smem_floats = 2 * (size_t)D * D + (size_t)kChainSlots * D * R + 4 * (size_t)D * R +
                                          (size_t)kChainKS * R * D;
};

template <int D>
__global__ void __launch_bounds__(kChainKS * D / 2, 1) chain_kernel(const ChainArgs args) {
    using C = ChainCfg<D>;
    constexpr int R = C::R, NT = C::T, KL = C::KL;
    extern __shared__ __align__(16) float smem[];
    float* wbuf = smem;                                   // [2][D][D]: this stage's weights + the next GEMM stage's
    float* slots = wbuf + 2 * D * D;                      // 3 x [D][R] transposed activations, then wide [4D][R]
    float* red = slots + (kChainSlots + 4) * D * R;       // [KS][R][D] partial sums
    auto slot_ptr = [&](int s) -> float* { return slots + s * D * R; };   // wide slot = index 3 (4x larger)

    const int t = threadIdx.x;
    static_assert(R == 8 && kChainKS == 8, "float4 pairs over 8 rows; 3-bit chunk swizzle over 8 k-slices");
    constexpr int CG = D / 4;
    const int cg = t % CG, rg = (t / CG) & 1, ks = t / (2 * CG);      // multiply role
    // finishing role: 8-byte piece t of the slice = rows frow, frow + 1 of column fc
    const int fc = t >> 2, frow = ((t >> 1) & 1) * 4 + (t & 1) * 2;
    const int row0 = blockIdx.x * R;
    const int n_rows = args.n_rows;
    // element-wise stages walk (r, c4) with r fastest so that a warp touches 8 rows x 64 contiguous bytes
    const int er = t & (R - 1), ec = t / R;

    // The stage table is a 3 KB kernel parameter; indexing it dynamically turns every field access into a dependent
    // constant-bank load, and cold constant-cache lines cost ~150 cycles each (clock64 trace: ~1000 cycles of stage
    // prologue and ~1500 of epilogue were nothing but these).  Stage descriptors are therefore staged in shared
    // memory once and each stage copies its descriptor into registers with independent loads.
    __shared__ __align__(16) ChainStage s_stage[kChainMaxStages];
    {
        static_assert(sizeof(ChainStage) % 4 == 0, "word copy");
        const int nwords = args.n_stages * (int)(sizeof(ChainStage) / 4);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(args.st);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_stage);
        for (int i = t; i < nwords; i += NT) dst[i] = src[i];
    }
    __shared__ __align__(8) uint64_t wbar[2];
    if (t == 0) {
        mbar_init(&wbar[0], 1);
        mbar_init(&wbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned wphase[2] = {0u, 0u};
    // warp 0 requests stage si's weights into buffer buf (all earlier reads of that buffer are behind a barrier)
    auto issue_weights = [&](int si, int buf) {
        if (t >= 32) return;
        const ChainStage& st = s_stage[si];
        float* dstw = wbuf + buf * D * D;
        if (t == 0) mbar_expect_tx(&wbar[buf], (unsigned)(D * D * sizeof(float)));
        if (st.ldw == D) {                                // contiguous matrix: one request
            if (t == 0) bulk_g2s(dstw, st.W, (unsigned)(D * D * sizeof(float)), &wbar[buf]);
        } else {
            __syncwarp();
            for (int r = t; r < D; r += 32)
                bulk_g2s(dstw + r * D, st.W + (size_t)r * st.ldw, (unsigned)(D * sizeof(float)), &wbar[buf]);
        }
    };
    auto next_gemm = [&](int from) {
        for (int i = from; i < args.n_stages; ++i)
            if (s_stage[i].op == CH_GEMM) return i;
        return -1;
    };
    float4 zpre = make_float4(0.f, 0.f, 0.f, 0.f);   // prefetched SiLU' operand of stage zpre_stage (thread's first c4)
    int zpre_stage = -1;
    int wcur = 0;
    {
        const int first = next_gemm(0);
        if (first >= 0) issue_weights(first, 0);
    }
    // Everything above touched only kernel parameters and weights (written by kernels that finished before the
    // predecessor of this launch started); from here on the predecessor's outputs are read.
    pdl_wait();

    const int n_stages = args.n_stages;
    for (int si = 0; si < n_stages; ++si) {
        const ChainStage& st = s_stage[si];               // fields are read from shared memory where they are used (a
                                                          // full register copy pushed the kernel over 128 registers)
        if (si == n_stages - 1) pdl_trigger();            // the next kernel's CTAs may start arriving
        CH_STAMP(si, 0);
        if (st.op == CH_LOAD) {
            float* d = slot_ptr(st.dst);
            const int w4 = st.width / 4;
            const bool live = row0 + er < n_rows;
            for (int c4 = ec; c4 < w4; c4 += NT / R) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) {
                    v = ld4(st.g0 + (size_t)(row0 + er) * st.ld_g + c4 * 4);
                    if (st.g1) v = v + ld4(st.g1 + (size_t)(row0 + er) * st.ld_g + c4 * 4);
                    if (st.add_slot >= 0) {     // + a slot already in shared memory (same thread wrote / reads it)
                        const float* a = slot_ptr(st.add_slot) + (c4 * 4) * R + er;
                        v = v + make_float4(a[0], a[R], a[2 * R], a[3 * R]);
                    }
                    if (st.out_a) st4(st.out_a + (size_t)(row0 + er) * st.ld_out + c4 * 4, v);
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
            for (int c4 = ec; c4 < D / 4; c4 += NT / R) {
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
            // att = o . W (global_message_passing.py:47), out = o . W_out.weight + b (:48)
            const float* s = slot_ptr(st.src);
            float a = 0.f, o = 0.f;
            for (int cc = ec; cc < D; cc += NT / R) {
                const float x = s[cc * R + er];
                a = fmaf(x, st.W[cc], a);
                o = fmaf(x, st.bias[cc], o);
            }
            // lanes with equal (lane & 7) hold the same row: fold lane bits 3 and 4, then the warps via smem
            a += __shfl_xor_sync(0xffffffffu, a, 8);  o += __shfl_xor_sync(0xffffffffu, o, 8);
            a += __shfl_xor_sync(0xffffffffu, a, 16); o += __shfl_xor_sync(0xffffffffu, o, 16);
            const int lane = t & 31, warp = t >> 5;
            if (lane < R) { red[(warp * R + lane) * 2] = a; red[(warp * R + lane) * 2 + 1] = o; }
            __syncthreads();
            if (t < R && row0 + t < n_rows) {
                float sa = 0.f, so = 0.f;
                for (int w = 0; w < NT / 32; ++w) { sa += red[(w * R + t) * 2]; so += red[(w * R + t) * 2 + 1]; }
                st.out_z[row0 + t] = sa;
                st.out_a[row0 + t] = so + st.g0[0];
            }
            __syncthreads();
        } else {  // CH_GEMM
            // prefetch the NEXT GEMM stage's weights into the other buffer (free since the previous stage's
            // trailing barrier), then wait only for this stage's group
            const int nxt = st.next_gemm;
            CH_STAMP(si, 8);
            if (nxt >= 0) issue_weights(nxt, wcur ^ 1);
            CH_STAMP(si, 9);

            const float* in = slot_ptr(st.src) + st.src_off * R;
            // operands of this stage's epilogue: requested now, consumed after the k-loop (an exposed L2 round
            // trip per stage was the largest remaining cost of the chain)
            float bias_v = st.bias ? st.bias[fc] : 0.f;
            float addg_v[2] = {0.f, 0.f};
            if (st.add_g) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (row0 + frow + i < n_rows) addg_v[i] = st.add_g[(size_t)(row0 + frow + i) * st.ld_add + fc];
            }
            float zpost_v[2] = {0.f, 0.f};       // SiLU' operand of the NEXT stage's (fused) prologue
            if (st.post_dst >= 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (row0 + frow + i < n_rows) zpost_v[i] = st.post_zmul[(size_t)(row0 + frow + i) * D + fc];
            }
            CH_STAMP(si, 10);
            if (st.psrc >= 0) {
                float* p = slot_ptr(st.psrc);
                const bool live = row0 + er < n_rows;
                for (int c4 = ec; c4 < D / 4; c4 += NT / R) {
                    const float* g = in + (c4 * 4) * R + er;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live) {
                        // zmul of this stage was prefetched during the previous GEMM stage when possible
                        const float4 zz = (zpre_stage == si && c4 == ec) ? zpre : ld4(st.zmul + (size_t)(row0 + er) * D + c4 * 4);
                        const float4 dz = dsilu4(zz);
                        v = make_float4(g[0] * dz.x, g[R] * dz.y, g[2 * R] * dz.z, g[3 * R] * dz.w);
                        if (st.save_src) st4(st.save_src + (size_t)(row0 + er) * D + c4 * 4, v);
                    }
                    float* q = p + (c4 * 4) * R + er;
                    q[0] = v.x; q[R] = v.y; q[2 * R] = v.z; q[3 * R] = v.w;
                }
                in = p;
            }
            // prefetch the NEXT GEMM stage's SiLU' operand (its prologue runs right after this stage's barrier)
            if (nxt >= 0 && s_stage[nxt].psrc >= 0 && ec < D / 4 && row0 + er < n_rows) {
                zpre = ld4(s_stage[nxt].zmul + (size_t)(row0 + er) * D + ec * 4);
                zpre_stage = nxt;
            }
            CH_STAMP(si, 1);
            mbar_wait(&wbar[wcur], wphase[wcur]);         // this stage's weights have landed
            wphase[wcur] ^= 1u;
            CH_STAMP(si, 2);
            __syncthreads();                              // prologue visible
            CH_STAMP(si, 3);

            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            const float* wp = wbuf + wcur * D * D + (ks * KL) * D + 4 * cg;
            const float* ap = in + (ks * KL) * R + 4 * rg;
            // explicit software pipeline: operands of step kk+1 are requested before the 16 FMAs of step kk issue.
            // (Packed FFMA2 was tried: same FMA throughput, plus operand packing -- 25 % slower in the clock64 trace.)
            float4 wv = ld4(wp), a0 = ld4(ap);
#pragma unroll
            for (int kk = 0; kk < KL; ++kk) {
                const int kn = (kk + 1 < KL) ? kk + 1 : kk;
                const float4 wn = ld4(wp + kn * D);
                const float4 a0n = ld4(ap + kn * R);
                const float av[4] = {a0.x, a0.y, a0.z, a0.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] = fmaf(av[i], wv.x, acc[i][0]);
                    acc[i][1] = fmaf(av[i], wv.y, acc[i][1]);
                    acc[i][2] = fmaf(av[i], wv.z, acc[i][2]);
                    acc[i][3] = fmaf(av[i], wv.w, acc[i][3]);
                }
                wv = wn; a0 = a0n;
            }
            CH_STAMP(si, 4);
            {   // partial sums -> red[ks][chunk], chunk (column, row half) = 2 * column + half, low 3 bits swizzled
                float* rk = red + ks * (D * R);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int chunk = 8 * cg + ((2 * j + rg) ^ (cg & 7));
                    st4(rk + chunk * 4, make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]));
                }
            }
            __syncthreads();
            CH_STAMP(si, 5);

            // thread t finishes rows frow, frow + 1 of column fc: slices summed in fixed order -> deterministic
            {
                const int lc = t >> 1, chunk = lc ^ ((lc >> 3) & 7);
                const float* rp = red + chunk * 4 + (t & 1) * 2;
                float2 v = make_float2(bias_v, bias_v);
#pragma unroll
                for (int s2 = 0; s2 < kChainKS; ++s2) {
                    const float2 p = *reinterpret_cast<const float2*>(rp + s2 * (D * R));
                    v.x += p.x; v.y += p.y;
                }
                float x[2] = {v.x, v.y};
                if (x[0] == 12345.678f) CH_STAMP(si, 15);   // (forces the loads to complete before the next stamp)
                CH_STAMP(si, 11);
                float addv[2] = {0.f, 0.f};
                if (st.add_slot >= 0) {
                    const float2 a = *reinterpret_cast<const float2*>(slot_ptr(st.add_slot) + fc * R + frow);
                    addv[0] = a.x; addv[1] = a.y;
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int r = frow + i;
                    const bool live = row0 + r < n_rows;
                    if (live && st.out_z) st.out_z[(size_t)(row0 + r) * st.ld_out + fc] = x[i];
                    if (st.act) x[i] = silu(x[i]);
                    x[i] += addv[i];
                    if (live && st.add_g) x[i] += addg_v[i];
                    if (!live) x[i] = 0.f;
                    if (live && st.out_a) st.out_a[(size_t)(row0 + r) * st.ld_out + fc] = x[i];
                }
                CH_STAMP(si, 12);
                if (st.dst >= 0 && st.dst != st.post_dst)
                    *reinterpret_cast<float2*>(slot_ptr(st.dst) + fc * R + frow) = make_float2(x[0], x[1]);
                if (st.post_dst >= 0) {          // the next stage's prologue, on register values
                    float y[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const bool live = row0 + frow + i < n_rows;
                        y[i] = live ? x[i] * dsilu(zpost_v[i]) : 0.f;
                        if (live && st.post_save) st.post_save[(size_t)(row0 + frow + i) * D + fc] = y[i];
                    }
                    *reinterpret_cast<float2*>(slot_ptr(st.post_dst) + fc * R + frow) = make_float2(y[0], y[1]);
                }
            }
            CH_STAMP(si, 6);
            wcur ^= 1;
            __syncthreads();
            CH_STAMP(si, 7);
        }
    }
}

template <int D>
static int chain_launch_t(const ChainArgs& a, double bytes, cudaStream_t st) {
    using C = ChainCfg<D>;
    const size_t smem = C::smem_floats * sizeof(float);
    PAMNET_TRY(func_smem_once(reinterpret_cast<const void*>(chain_kernel<D>), smem));
    prof_begin(KC_CHAIN, bytes, st);
    launch_pdl(chain_kernel<D>, dim3(ceil_div(a.n_rows, C::R)), dim3(C::T), smem, st, a);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// ---- row-group interpreter for dim <= 32 -----------------------------------------------------------------------------
// At dim 16 / 32 a stage is a 1 - 4 KB weight matrix: the CTA-per-8-rows kernel above (k split over 8 thread groups, partial
// sums through shared memory, three CTA barriers and a weight mbarrier per stage) spends its time on fixed per-stage
// latencies -- 19-30 us per launch on the 15 816-node RNA batch whatever the arithmetic.  Here FOUR LANES own a row (lane
// q of the group: columns q D/4 .. (q+1) D/4 - 1 of every stage's output): the row's slots live in shared memory as
// [element][row] (conflict-free), only __syncwarp separates a stage's writes from the next stage's reads -- no CTA
// barrier, no mbarrier anywhere in the stage loop -- and the weights are read straight from L1 / L2 (a warp reads one
// 64- or 128-byte weight row per k).  Same stage table, same semantics as chain_kernel.  (One thread per row was tried
// first: 3 warps per SM on that batch, every L1 / shared-memory latency exposed -- no faster than the CTA kernel.)
constexpr int kRowChainThreads = 128, kRowChainTPR = 4;
constexpr int kRowChainRows = kRowChainThreads / kRowChainTPR;      // rows per CTA
template <int D>
__global__ void __launch_bounds__(kRowChainThreads) chain_rows_kernel(const ChainArgs args) {
    constexpr int NT = kRowChainThreads, NR = kRowChainRows, TPR = kRowChainTPR, CQ = D / TPR;     // CQ columns per lane
    static_assert(CQ % 4 == 0, "128-bit column groups");
    extern __shared__ __align__(16) float smem[];         // [(kChainSlots + 4) * D][NR]
    __shared__ __align__(16) ChainStage s_stage[kChainMaxStages];
    const int t = threadIdx.x, lane = t & 31;
    {
        const int nwords = args.n_stages * (int)(sizeof(ChainStage) / 4);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(args.st);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_stage);
        for (int i = t; i < nwords; i += NT) dst[i] = src[i];
    }
    __syncthreads();
    pdl_wait();
    const int n_rows = args.n_rows, n_stages = args.n_stages;
    const int r = t / TPR, q = t % TPR;
    const int m = blockIdx.x * NR + r;
    const bool live = m < n_rows;
    const size_t mr = (size_t)(live ? m : 0);
    float* mine = smem + r;                               // element e of slot s of this row: mine[(s * D + e) * NR]
    const int c0 = q * CQ;                                // this lane's columns of a D-wide row
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };

    for (int si = 0; si < n_stages; ++si) {
        const ChainStage& st = s_stage[si];
        if (si == n_stages - 1) pdl_trigger();
        if (st.op == CH_LOAD) {
            float* d = mine + (size_t)st.dst * D * NR;
            const float* g0 = st.g0 + mr * st.ld_g;
            const float* g1 = st.g1 ? st.g1 + mr * st.ld_g : nullptr;
            const float* as = st.add_slot >= 0 ? mine + (size_t)st.add_slot * D * NR : nullptr;
            float* oa = st.out_a ? st.out_a + mr * st.ld_out : nullptr;
            const bool v4 = al16(g0) && (!g1 || al16(g1)) && (!oa || al16(oa));
            for (int c = 4 * q; c < st.width; c += 4 * TPR) {        // lanes interleave 16-byte groups: coalesced rows
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (live) {
                    if (v4) {
                        const float4 a = ld4(g0 + c);
                        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                        if (g1) { const float4 b = ld4(g1 + c); v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w; }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] = g0[c + j] + (g1 ? g1[c + j] : 0.f);
                    }
                    if (as) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) v[j] += as[(size_t)(c + j) * NR];
                    }
                    if (oa) {
                        if (v4) st4(oa + c, make_float4(v[0], v[1], v[2], v[3]));
                        else { for (int j = 0; j < 4; ++j) oa[c + j] = v[j]; }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) d[(size_t)(c + j) * NR] = v[j];
            }
            __syncwarp();
        } else if (st.op == CH_HEADS_BWD) {
            float* d = mine + (size_t)st.dst * D * NR;
            const float ga = live ? st.g0[m] : 0.f, go = live ? st.g1[m] : 0.f;
            const float* W = st.W; const float* Wo = st.bias; const float* o3 = st.zmul;
            float* gW = st.out_z; float* gWo = st.out_a; float* gbo = st.save_src;
#pragma unroll
            for (int cc = 0; cc < CQ; ++cc) {
                const int c = c0 + cc;
                d[(size_t)c * NR] = ga * W[c] + go * Wo[c];
                if (o3) {
                    const float a = live ? o3[mr * D + c] : 0.f;
                    float pa = ga * a, po = go * a;
#pragma unroll
                    for (int o = 16; o >= TPR; o >>= 1) {            // the 8 rows of the warp (lanes with the same q)
                        pa += __shfl_xor_sync(0xffffffffu, pa, o);
                        po += __shfl_xor_sync(0xffffffffu, po, o);
                    }
                    if (lane < TPR) { atomicAdd(gW + c, pa); atomicAdd(gWo + c, po); }
                }
            }
            if (o3 && gbo) {
                float sgo = go;
#pragma unroll
                for (int o = 16; o >= TPR; o >>= 1) sgo += __shfl_xor_sync(0xffffffffu, sgo, o);
                if (lane == 0) atomicAdd(gbo, sgo);
            }
            __syncwarp();
        } else if (st.op == CH_DOT2) {
            const float* x = mine + (size_t)st.src * D * NR;
            float a = 0.f, o = 0.f;
#pragma unroll
            for (int cc = 0; cc < CQ; ++cc) {
                const float xv = x[(size_t)(c0 + cc) * NR];
                a = fmaf(xv, st.W[c0 + cc], a);
                o = fmaf(xv, st.bias[c0 + cc], o);
            }
#pragma unroll
            for (int s2 = 1; s2 < TPR; s2 <<= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, s2);
                o += __shfl_xor_sync(0xffffffffu, o, s2);
            }
            if (live && q == 0) { st.out_z[m] = a; st.out_a[m] = o + st.g0[0]; }
        } else {  // CH_GEMM
            const float* in = mine + ((size_t)st.src * D + st.src_off) * NR;
            if (st.psrc >= 0) {       // prologue: src * SiLU'(zmul) -> psrc (and to global, for the weight gradients)
                float* p = mine + (size_t)st.psrc * D * NR;
                const float* zr = st.zmul + mr * D;
                float* sv = st.save_src ? st.save_src + mr * D : nullptr;
#pragma unroll
                for (int cc = 0; cc < CQ; cc += 4) {
                    const int c = c0 + cc;
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    if (live) {
                        const float4 dz = dsilu4(ld4(zr + c));
                        v[0] = in[(size_t)c * NR] * dz.x; v[1] = in[(size_t)(c + 1) * NR] * dz.y;
                        v[2] = in[(size_t)(c + 2) * NR] * dz.z; v[3] = in[(size_t)(c + 3) * NR] * dz.w;
                        if (sv) st4(sv + c, make_float4(v[0], v[1], v[2], v[3]));
                    }
                    // (psrc == src is allowed: every lane rewrites exactly the elements it read)
#pragma unroll
                    for (int j = 0; j < 4; ++j) p[(size_t)(c + j) * NR] = v[j];
                }
                in = p;
                __syncwarp();
            }
            float acc[CQ];
#pragma unroll
            for (int n = 0; n < CQ; ++n) acc[n] = 0.f;
            const float* W = st.W + c0;
            const int ldw = st.ldw;
#pragma unroll 8
            for (int k = 0; k < D; ++k) {
                const float xk = in[(size_t)k * NR];
                const float* wr = W + (size_t)k * ldw;
#pragma unroll
                for (int n = 0; n < CQ; n += 4) {
                    const float4 w = ld4(wr + n);
                    acc[n] = fmaf(xk, w.x, acc[n]); acc[n + 1] = fmaf(xk, w.y, acc[n + 1]);
                    acc[n + 2] = fmaf(xk, w.z, acc[n + 2]); acc[n + 3] = fmaf(xk, w.w, acc[n + 3]);
                }
            }
            __syncwarp();             // the row's four lanes are done reading `in` (post_dst may be that very slot)
            const float* bias = st.bias;
            float* oz = st.out_z ? st.out_z + mr * st.ld_out : nullptr;
            float* oa = st.out_a ? st.out_a + mr * st.ld_out : nullptr;
            const float* ag = st.add_g ? st.add_g + mr * st.ld_add : nullptr;
            const float* as = st.add_slot >= 0 ? mine + (size_t)st.add_slot * D * NR : nullptr;
            float* dd = (st.dst >= 0 && st.dst != st.post_dst) ? mine + (size_t)st.dst * D * NR : nullptr;
            float* pd = st.post_dst >= 0 ? mine + (size_t)st.post_dst * D * NR : nullptr;
            const float* pz = st.post_dst >= 0 ? st.post_zmul + mr * D : nullptr;
            float* ps = (st.post_dst >= 0 && st.post_save) ? st.post_save + mr * D : nullptr;
            const bool v4 = (!oz || al16(oz)) && (!oa || al16(oa)) && (!ag || al16(ag));
            const int act = st.act;
#pragma unroll
            for (int nn = 0; nn < CQ; nn += 4) {
                const int n = c0 + nn;
                float x[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = acc[nn + j] + (bias ? bias[n + j] : 0.f);
                if (live && oz) {
                    if (v4) st4(oz + n, make_float4(x[0], x[1], x[2], x[3]));
                    else { for (int j = 0; j < 4; ++j) oz[n + j] = x[j]; }
                }
                if (act) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = silu(x[j]);
                }
                if (as) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] += as[(size_t)(n + j) * NR];
                }
                if (live && ag) {
                    if (v4) { const float4 a = ld4(ag + n); x[0] += a.x; x[1] += a.y; x[2] += a.z; x[3] += a.w; }
                    else { for (int j = 0; j < 4; ++j) x[j] += ag[n + j]; }
                }
                if (!live) { x[0] = 0.f; x[1] = 0.f; x[2] = 0.f; x[3] = 0.f; }
                if (live && oa) {
                    if (v4) st4(oa + n, make_float4(x[0], x[1], x[2], x[3]));
                    else { for (int j = 0; j < 4; ++j) oa[n + j] = x[j]; }
                }
                if (dd) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) dd[(size_t)(n + j) * NR] = x[j];
                }
                if (pd) {            // the next stage's prologue, on register values
                    float y[4] = {0.f, 0.f, 0.f, 0.f};
                    if (live) {
                        const float4 dz = dsilu4(ld4(pz + n));
                        y[0] = x[0] * dz.x; y[1] = x[1] * dz.y; y[2] = x[2] * dz.z; y[3] = x[3] * dz.w;
                        if (ps) st4(ps + n, make_float4(y[0], y[1], y[2], y[3]));
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) pd[(size_t)(n + j) * NR] = y[j];
                }
            }
            __syncwarp();
        }
    }
}

// PAMNET_CHAIN_ROWS=0: the CTA-per-8-rows interpreter also at dim <= 32
static bool chain_rows_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("PAMNET_CHAIN_ROWS"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

template <int D>
static int chain_rows_launch_t(const ChainArgs& a, double bytes, cudaStream_t st) {
    const size_t smem = sizeof(float) * (kChainSlots + 4) * D * kRowChainRows;
    PAMNET_TRY(func_smem_once(reinterpret_cast<const void*>(chain_rows_kernel<D>), smem));
    prof_begin(KC_CHAIN, bytes, st);
    launch_pdl(chain_rows_kernel<D>, dim3(ceil_div(a.n_rows, kRowChainRows)), dim3(kRowChainThreads), smem, st, a);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// stage-table preprocessing shared by both interpreters: algorithmic bytes, prologue fusion, next-GEMM links
static double chain_prepare(int D, const ChainArgs& args, ChainArgs* out) {
    double bytes = 0.0;
    for (int i = 0; i < args.n_stages; ++i) {
        const ChainStage& s = args.st[i];
        const double row = 4.0 * args.n_rows * D;
        if (s.op == CH_LOAD) bytes += 4.0 * args.n_rows * s.width * ((s.g1 ? 2 : 1) + (s.out_a ? 1 : 0));
        if (s.op == CH_GEMM)
            bytes += 4.0 * D * D + row * ((s.zmul ? 1 : 0) + (s.save_src ? 1 : 0) + (s.out_z ? 1 : 0) +
                                          (s.out_a ? 1 : 0) + (s.add_g ? 1 : 0));
        if (s.op == CH_DOT2 || s.op == CH_HEADS_BWD) bytes += 8.0 * args.n_rows + 8.0 * D;
    }
    ChainArgs& a = *out;
    a = args;
    for (int i = 0; i < a.n_stages; ++i) {
        a.st[i].post_dst = -1;
        if (a.st[i].op < CH_GMSG_FWD) { a.st[i].post_zmul = nullptr; a.st[i].post_save = nullptr; }   // (gather stages keep their CSR pointers there)
    }
    // prologue fusion (see ChainStage::post_dst): stage i + 1 = GEMM with a SiLU' prologue on exactly what stage i wrote
    static int fuse = -1;
    if (fuse < 0) { const char* e = getenv("PAMNET_CHAIN_FUSE"); fuse = (e && e[0] == '0') ? 0 : 1; }
    for (int i = 0; fuse && i + 1 < a.n_stages; ++i) {
        ChainStage& p = a.st[i];
        ChainStage& c = a.st[i + 1];
        if (p.op != CH_GEMM || c.op != CH_GEMM || c.psrc < 0 || p.dst < 0 || c.src != p.dst || c.src_off != 0) continue;
        p.post_dst = c.psrc; p.post_zmul = c.zmul; p.post_save = c.save_src;
        c.src = c.psrc; c.psrc = -1; c.zmul = nullptr; c.save_src = nullptr;
    }
    for (int i = a.n_stages - 1, nxt = -1; i >= 0; --i) {
        a.st[i].next_gemm = nxt;
        if (a.st[i].op == CH_GEMM) nxt = i;
    }
    return bytes;
}

int chain_launch(int dim, const ChainArgs& args, cudaStream_t st) {
    PAMNET_CHECK_ARG(args.n_stages <= kChainMaxStages, "chain: %d stages", args.n_stages);
    if (args.n_rows <= 0) return 0;
    for (int i = 0; i < args.n_stages; ++i) {
        const ChainStage& s = args.st[i];
        if (s.op == CH_GEMM)
            PAMNET_CHECK_ARG(s.dst != s.src && (s.psrc < 0 || s.dst != s.psrc) && s.dst != kChainWide,
                             "chain stage %d: output slot aliases its input", i);
    }
    ChainArgs a;
    const double bytes = chain_prepare(dim, args, &a);
    static int node_mlp = -1;      // PAMNET_NODE_MLP=tf32: single-pass TF32 node MLPs (opt-in reduced precision, configs[2])
    if (node_mlp < 0) { const char* e = getenv("PAMNET_NODE_MLP"); node_mlp = (e && strcmp(e, "tf32") == 0) ? 1 : 0; }
    a.precision = node_mlp;
    if (chain_mma_enabled(dim)) return chain_mma_launch(dim, a, bytes, st);
    if (dim <= 32 && chain_rows_enabled()) {
        if (dim == 32) return chain_rows_launch_t<32>(a, bytes, st);
        if (dim == 16) return chain_rows_launch_t<16>(a, bytes, st);
    }
    switch (dim) {
        case 128: return chain_launch_t<128>(a, bytes, st);
        case 64:  return chain_launch_t<64>(a, bytes, st);
        case 32:  return chain_launch_t<32>(a, bytes, st);
        case 16:  return chain_launch_t<16>(a, bytes, st);
        default:
            set_error("chain: unsupported dim %d (16, 32, 64, 128)", dim);
            return -1;
    }
}

}  // namespace pamnet
