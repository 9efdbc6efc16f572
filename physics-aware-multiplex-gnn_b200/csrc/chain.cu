// Stage interpreter for row-local node chains (see chain.cuh).  One CTA = 8 rows; activations live (transposed)
// in shared memory across all stages; whole D x D weight matrices are streamed k-major through a double buffer
// filled by the bulk-copy engine, the next stage's matrix in flight while the current one multiplies.
#include "chain.cuh"

namespace pamnet {

constexpr int kChainRows = 8;     // rows per CTA
constexpr int kChainKS = 8;       // k-slices: a CTA has kChainKS * D / 2 threads (each owns two columns)

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

// Thread (c, ks): output columns c and c + D/2 (c = tid % (D/2)), k-slice ks = tid / (D/2) of 8.  It accumulates
// all 8 rows of its two columns over its eighth of K: per k two conflict-free 128 B weight requests per warp plus
// two broadcast 128-bit loads of the 8 row values (activations are kept TRANSPOSED in shared memory, [k][8]) feed
// 16 FMAs.  The eight partial sums meet in shared memory and thread (c, ks) finishes row ks of its two columns.
// 4*D threads per CTA = 16 warps at D = 128: enough warps to hide the shared-memory latency that a 2-warp-per-
// scheduler version could not (ncu: short_scoreboard), at 4 shared-memory wavefronts per 16 FMAs.
template <int D>
struct ChainCfg {
    static constexpr int R = kChainRows;
    static constexpr int T = kChainKS * D / 2;           // threads
    static constexpr int KL = D / kChainKS;              // k per slice
    static constexpr size_t smem_floats = 2 * (size_t)D * D + (size_t)kChainSlots * D * R + 4 * (size_t)D * R +
                                          (size_t)kChainKS * R * D;
};

template <int D>
__global__ void __launch_bounds__(kChainKS * D / 2) chain_kernel(const ChainArgs args) {
    using C = ChainCfg<D>;
    constexpr int R = C::R, NT = C::T, KL = C::KL;
    extern __shared__ __align__(16) float smem[];
    float* wbuf = smem;                                   // [2][D][D]: this stage's weights + the next GEMM stage's
    float* slots = wbuf + 2 * D * D;                      // 3 x [D][R] transposed activations, then wide [4D][R]
    float* red = slots + (kChainSlots + 4) * D * R;       // [KS][R][D] partial sums
    auto slot_ptr = [&](int s) -> float* { return slots + s * D * R; };   // wide slot = index 3 (4x larger)

    const int t = threadIdx.x;
    constexpr int H = D / 2;
    const int c = t % H, ks = t / H;
    const int row0 = blockIdx.x * R;
    const int n_rows = args.n_rows;
    // element-wise stages walk (r, c4) with r fastest so that a warp touches 8 rows x 64 contiguous bytes
    const int er = t & (R - 1), ec = t / R;

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
        const ChainStage& st = args.st[si];
        float* dstw = wbuf + buf * D * D;
        if (t == 0) mbar_expect_tx(&wbar[buf], (unsigned)(D * D * sizeof(float)));
        __syncwarp();
        if (st.ldw == D) {
            constexpr int kParts = D >= 32 ? 4 : 1;
            if (t < kParts) bulk_g2s(dstw + t * (D * D / kParts), st.W + t * (D * D / kParts),
                                     (unsigned)(D * D / kParts * sizeof(float)), &wbar[buf]);
        } else {
            for (int r = t; r < D; r += 32)
                bulk_g2s(dstw + r * D, st.W + (size_t)r * st.ldw, (unsigned)(D * sizeof(float)), &wbar[buf]);
        }
    };
    auto next_gemm = [&](int from) {
        for (int i = from; i < args.n_stages; ++i)
            if (args.st[i].op == CH_GEMM) return i;
        return -1;
    };
    float4 zpre = make_float4(0.f, 0.f, 0.f, 0.f);   // prefetched SiLU' operand of stage zpre_stage (thread's first c4)
    int zpre_stage = -1;
    int wcur = 0;
    {
        const int first = next_gemm(0);
        if (first >= 0) issue_weights(first, 0);
    }

    for (int si = 0; si < args.n_stages; ++si) {
        const ChainStage& st = args.st[si];
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
            // grad of o3 from the two heads: g_att * W + g_out * W_out.weight
            float* d = slot_ptr(st.dst);
            const bool live = row0 + er < n_rows;
            const float ga = live ? st.g0[row0 + er] : 0.f, go = live ? st.g1[row0 + er] : 0.f;
            for (int c4 = ec; c4 < D / 4; c4 += NT / R) {
                const float4 w = ld4(st.W + c4 * 4), wo = ld4(st.bias + c4 * 4);
                float* q = d + (c4 * 4) * R + er;
                q[0] = ga * w.x + go * wo.x; q[R] = ga * w.y + go * wo.y;
                q[2 * R] = ga * w.z + go * wo.z; q[3 * R] = ga * w.w + go * wo.w;
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
            const int nxt = next_gemm(si + 1);
            if (nxt >= 0) issue_weights(nxt, wcur ^ 1);

            const float* in = slot_ptr(st.src) + st.src_off * R;
            // operands of this stage's epilogue: requested now, consumed after the k-loop (an exposed L2 round
            // trip per stage was the largest remaining cost of the chain)
            const bool live_r = row0 + ks < n_rows;
            float bias_v[2] = {0.f, 0.f}, addg_v[2] = {0.f, 0.f};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (st.bias) bias_v[h] = st.bias[c + h * H];
                if (st.add_g && live_r) addg_v[h] = st.add_g[(size_t)(row0 + ks) * st.ld_add + c + h * H];
            }
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
            if (nxt >= 0 && args.st[nxt].psrc >= 0 && ec < D / 4 && row0 + er < n_rows) {
                zpre = ld4(args.st[nxt].zmul + (size_t)(row0 + er) * D + ec * 4);
                zpre_stage = nxt;
            }
            mbar_wait(&wbar[wcur], wphase[wcur]);         // this stage's weights have landed
            wphase[wcur] ^= 1u;
            __syncthreads();                              // prologue visible

            float acc0[R], acc1[R];
#pragma unroll
            for (int i = 0; i < R; ++i) acc0[i] = acc1[i] = 0.f;
            const float* wp = wbuf + wcur * D * D + (ks * KL) * D + c;
            const float* ap = in + (ks * KL) * R;
            // explicit software pipeline: operands of step kk+1 are requested before the 16 FMAs of step kk issue
            float w0 = wp[0], w1 = wp[H];
            float4 a0 = ld4(ap), a1 = ld4(ap + 4);
#pragma unroll
            for (int kk = 0; kk < KL; ++kk) {
                const int kn = (kk + 1 < KL) ? kk + 1 : kk;
                const float w0n = wp[kn * D], w1n = wp[kn * D + H];
                const float4 a0n = ld4(ap + kn * R), a1n = ld4(ap + kn * R + 4);
                acc0[0] = fmaf(a0.x, w0, acc0[0]); acc0[1] = fmaf(a0.y, w0, acc0[1]);
                acc0[2] = fmaf(a0.z, w0, acc0[2]); acc0[3] = fmaf(a0.w, w0, acc0[3]);
                acc0[4] = fmaf(a1.x, w0, acc0[4]); acc0[5] = fmaf(a1.y, w0, acc0[5]);
                acc0[6] = fmaf(a1.z, w0, acc0[6]); acc0[7] = fmaf(a1.w, w0, acc0[7]);
                acc1[0] = fmaf(a0.x, w1, acc1[0]); acc1[1] = fmaf(a0.y, w1, acc1[1]);
                acc1[2] = fmaf(a0.z, w1, acc1[2]); acc1[3] = fmaf(a0.w, w1, acc1[3]);
                acc1[4] = fmaf(a1.x, w1, acc1[4]); acc1[5] = fmaf(a1.y, w1, acc1[5]);
                acc1[6] = fmaf(a1.z, w1, acc1[6]); acc1[7] = fmaf(a1.w, w1, acc1[7]);
                w0 = w0n; w1 = w1n; a0 = a0n; a1 = a1n;
            }
#pragma unroll
            for (int i = 0; i < R; ++i) {
                red[(ks * R + i) * D + c] = acc0[i];
                red[(ks * R + i) * D + c + H] = acc1[i];
            }
            __syncthreads();

            // thread (c, ks) finishes row ks of its two columns: slices summed in fixed order -> deterministic
            static_assert(kChainKS == kChainRows, "one row per k-slice thread group");
            const int r = ks;
            const bool live = row0 + r < n_rows;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = c + h * H;
                float v = bias_v[h];
#pragma unroll
                for (int s2 = 0; s2 < kChainKS; ++s2) v += red[(s2 * R + r) * D + col];
                if (live && st.out_z) st.out_z[(size_t)(row0 + r) * st.ld_out + col] = v;
                if (st.act) v = silu(v);
                if (st.add_slot >= 0) v += slot_ptr(st.add_slot)[col * R + r];
                if (live && st.add_g) v += addg_v[h];
                if (!live) v = 0.f;
                if (st.dst >= 0) slot_ptr(st.dst)[col * R + r] = v;
                if (live && st.out_a) st.out_a[(size_t)(row0 + r) * st.ld_out + col] = v;
            }
            wcur ^= 1;
            __syncthreads();
        }
    }
}

template <int D>
static int chain_launch_t(const ChainArgs& args, cudaStream_t st) {
    using C = ChainCfg<D>;
    const size_t smem = C::smem_floats * sizeof(float);
    static bool configured = false;   // per-D instantiation; the attribute is per-function and idempotent
    if (!configured) {
        PAMNET_CUDA(cudaFuncSetAttribute(chain_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
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
    prof_begin(KC_CHAIN, bytes, st);
    chain_kernel<D><<<ceil_div(args.n_rows, C::R), C::T, smem, st>>>(args);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
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
    switch (dim) {
        case 128: return chain_launch_t<128>(args, st);
        case 64:  return chain_launch_t<64>(args, st);
        case 32:  return chain_launch_t<32>(args, st);
        case 16:  return chain_launch_t<16>(args, st);
        default:
            set_error("chain: unsupported dim %d (16, 32, 64, 128)", dim);
            return -1;
    }
}

}  // namespace pamnet
