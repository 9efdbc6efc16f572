// Stage interpreter for row-local node chains (see chain.cuh).  One CTA = R rows; activations live in
// shared memory across all stages; D x D weights are streamed k-major through a cp.async double buffer.
#include "chain.cuh"

namespace pamnet {

constexpr int kChainThreads = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Thread (c, g): output column c = tid % D, row group g = tid / D owning RT consecutive rows.  Per k the warp
// reads 32 consecutive weights (one conflict-free 128 B request) and RT activations by broadcast, so a weight
// element leaves shared memory once per row group instead of once per row.
template <int D, int RT>
struct ChainCfg {
    static constexpr int G = kChainThreads / D;          // row groups
    static constexpr int R = G * RT;                     // rows per CTA
    static constexpr int LD = D + 4;                     // slot row stride (floats)
    static constexpr int LDW = 4 * D + 4;                // wide slot row stride
    static constexpr size_t smem_floats = (size_t)kChainSlots * R * LD + (size_t)R * LDW + 2 * (size_t)D * D;
};

template <int D, int RT>
__global__ void __launch_bounds__(kChainThreads) chain_kernel(const ChainArgs args) {
    using C = ChainCfg<D, RT>;
    extern __shared__ __align__(16) float smem[];
    float* wbuf = smem;                                   // [2][D][D]: this stage's weights + the next GEMM stage's
    float* slots = smem + 2 * D * D;                      // [3][R][LD] then wide [R][LDW]
    auto slot_ptr = [&](int s) -> float* {
        return s == kChainWide ? slots + kChainSlots * C::R * C::LD : slots + s * C::R * C::LD;
    };
    auto slot_ld = [&](int s) -> int { return s == kChainWide ? C::LDW : C::LD; };

    const int t = threadIdx.x;
    const int c = t % D, r0 = (t / D) * RT;
    const int row0 = blockIdx.x * C::R;
    const int n_rows = args.n_rows;

    // whole D x D weight matrix of GEMM stage `si` -> wbuf[buf]; one commit group per matrix
    auto issue_weights = [&](int si, int buf) {
        const ChainStage& st = args.st[si];
        float* dstw = wbuf + buf * D * D;
        for (int f = t; f < D * (D / 4); f += kChainThreads) {
            const int r = f / (D / 4), cc = (f % (D / 4)) * 4;
            cp_async16(dstw + r * D + cc, st.W + (size_t)r * st.ldw + cc);
        }
        cp_async_commit();
    };
    auto next_gemm = [&](int from) {
        for (int i = from; i < args.n_stages; ++i)
            if (args.st[i].op == CH_GEMM) return i;
        return -1;
    };
    int wcur = 0;                                         // buffer holding the upcoming GEMM stage's weights
    {
        const int first = next_gemm(0);
        if (first >= 0) issue_weights(first, 0);
    }

    for (int si = 0; si < args.n_stages; ++si) {
        const ChainStage& st = args.st[si];
        if (st.op == CH_LOAD) {
            float* d = slot_ptr(st.dst);
            const int ld = slot_ld(st.dst), w4 = st.width / 4;
            for (int f = t; f < C::R * w4; f += kChainThreads) {
                const int r = f / w4, cc = (f % w4) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row0 + r < n_rows) {
                    v = ld4(st.g0 + (size_t)(row0 + r) * st.ld_g + cc);
                    if (st.g1) v = v + ld4(st.g1 + (size_t)(row0 + r) * st.ld_g + cc);
                    if (st.out_a) st4(st.out_a + (size_t)(row0 + r) * st.ld_out + cc, v);
                }
                st4(d + r * ld + cc, v);
            }
            __syncthreads();
        } else if (st.op == CH_HEADS_BWD) {
            // grad of o3 from the two heads: g_att * W + g_out * W_out.weight
            float* d = slot_ptr(st.dst);
            for (int f = t; f < C::R * (D / 4); f += kChainThreads) {
                const int r = f / (D / 4), cc = (f % (D / 4)) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row0 + r < n_rows) {
                    const float ga = st.g0[row0 + r], go = st.g1[row0 + r];
                    const float4 w = ld4(st.W + cc), wo = ld4(st.bias + cc);
                    v = make_float4(ga * w.x + go * wo.x, ga * w.y + go * wo.y, ga * w.z + go * wo.z,
                                    ga * w.w + go * wo.w);
                }
                st4(d + r * C::LD + cc, v);
            }
            __syncthreads();
        } else if (st.op == CH_DOT2) {
            // att = o . W (global_message_passing.py:47), out = o . W_out.weight + b (:48); one warp per row
            const float* s = slot_ptr(st.src);
            const int lane = t & 31, warp = t >> 5;
            for (int r = warp; r < C::R; r += kChainThreads / 32) {
                float a = 0.f, o = 0.f;
                for (int cc = lane; cc < D; cc += 32) {
                    const float x = s[r * C::LD + cc];
                    a = fmaf(x, st.W[cc], a);
                    o = fmaf(x, st.bias[cc], o);
                }
                a = warp_sum(a);
                o = warp_sum(o);
                if (lane == 0 && row0 + r < n_rows) {
                    st.out_z[row0 + r] = a;
                    st.out_a[row0 + r] = o + st.g0[0];
                }
            }
            __syncthreads();
        } else {  // CH_GEMM
            // prefetch the NEXT GEMM stage's weights into the other buffer (free since the previous stage's
            // trailing barrier), then wait only for this stage's group
            const int nxt = next_gemm(si + 1);
            if (nxt >= 0) issue_weights(nxt, wcur ^ 1);

            const float* in = slot_ptr(st.src) + st.src_off;
            int in_ld = slot_ld(st.src);
            if (st.psrc >= 0) {
                float* p = slot_ptr(st.psrc);
                for (int f = t; f < C::R * (D / 4); f += kChainThreads) {
                    const int r = f / (D / 4), cc = (f % (D / 4)) * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row0 + r < n_rows) {
                        const float4 g = ld4(in + r * in_ld + cc);
                        v = g * dsilu4(ld4(st.zmul + (size_t)(row0 + r) * D + cc));
                        if (st.save_src) st4(st.save_src + (size_t)(row0 + r) * D + cc, v);
                    }
                    st4(p + r * C::LD + cc, v);
                }
                in = p;
                in_ld = C::LD;
            }
            if (nxt >= 0) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncthreads();                              // weights landed for everyone; prologue visible

            float acc[RT];
#pragma unroll
            for (int i = 0; i < RT; ++i) acc[i] = 0.f;
            const float* wk = wbuf + wcur * D * D + c;
            const float* a0 = in + r0 * in_ld;
#pragma unroll 4
            for (int k = 0; k < D; k += 4) {
                float4 a[RT];
#pragma unroll
                for (int i = 0; i < RT; ++i) a[i] = ld4(a0 + i * in_ld + k);
                const float w0 = wk[(k + 0) * D], w1 = wk[(k + 1) * D], w2 = wk[(k + 2) * D], w3 = wk[(k + 3) * D];
#pragma unroll
                for (int i = 0; i < RT; ++i) {
                    acc[i] = fmaf(a[i].x, w0, acc[i]);
                    acc[i] = fmaf(a[i].y, w1, acc[i]);
                    acc[i] = fmaf(a[i].z, w2, acc[i]);
                    acc[i] = fmaf(a[i].w, w3, acc[i]);
                }
            }
            // epilogue: lanes own consecutive columns -> coalesced row segments
            const float b = st.bias ? st.bias[c] : 0.f;
#pragma unroll
            for (int i = 0; i < RT; ++i) {
                const int r = r0 + i;
                const bool live = row0 + r < n_rows;
                float v = acc[i] + b;
                if (live && st.out_z) st.out_z[(size_t)(row0 + r) * st.ld_out + c] = v;
                if (st.act) v = silu(v);
                if (st.add_slot >= 0) v += slot_ptr(st.add_slot)[r * C::LD + c];
                if (live && st.add_g) v += st.add_g[(size_t)(row0 + r) * st.ld_add + c];
                if (!live) v = 0.f;
                if (st.dst >= 0) slot_ptr(st.dst)[r * C::LD + c] = v;
                if (live && st.out_a) st.out_a[(size_t)(row0 + r) * st.ld_out + c] = v;
            }
            wcur ^= 1;
            __syncthreads();
        }
    }
}

template <int D, int RT>
static int chain_launch_t(const ChainArgs& args, cudaStream_t st) {
    using C = ChainCfg<D, RT>;
    const size_t smem = C::smem_floats * sizeof(float);
    static bool configured = false;   // per (D, RT) instantiation; attribute is per-function and idempotent
    if (!configured) {
        PAMNET_CUDA(cudaFuncSetAttribute(chain_kernel<D, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    chain_kernel<D, RT><<<ceil_div(args.n_rows, C::R), kChainThreads, smem, st>>>(args);
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
    // rows per CTA = (256 / D) * RT: small CTAs until the row count fills the 148 SMs twice over
    const int n = args.n_rows;
    switch (dim) {
        case 128: return (n >= 148 * 32) ? chain_launch_t<128, 8>(args, st) : chain_launch_t<128, 4>(args, st);
        case 64:  return (n >= 148 * 64) ? chain_launch_t<64, 8>(args, st) : chain_launch_t<64, 4>(args, st);
        case 32:  return (n >= 148 * 128) ? chain_launch_t<32, 8>(args, st) : chain_launch_t<32, 4>(args, st);
        case 16:  return (n >= 148 * 256) ? chain_launch_t<16, 8>(args, st) : chain_launch_t<16, 4>(args, st);
        default:
            set_error("chain: unsupported dim %d (16, 32, 64, 128)", dim);
            return -1;
    }
}

}  // namespace pamnet
