// Stage interpreter for row-local node chains (see chain.cuh).  One CTA = R rows; activations live in
// shared memory across all stages; D x D weights are streamed k-major through a cp.async double buffer.
#include "chain.cuh"

namespace pamnet {

constexpr int kChainThreads = 256;
constexpr int KC = 32;   // k rows of W per smem chunk

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int D, int RPT>
struct ChainCfg {
    static constexpr int CX = D / 4;                     // column groups of 4
    static constexpr int RY = kChainThreads / CX;        // row groups
    static constexpr int R = RY * RPT;                   // rows per CTA
    static constexpr int LD = D + 4;                     // slot row stride
    static constexpr int LDW = 4 * D + 4;                // wide slot row stride
    static constexpr size_t smem_floats = (size_t)kChainSlots * R * LD + (size_t)R * LDW + 2 * KC * D;
};

template <int D, int RPT>
__global__ void __launch_bounds__(kChainThreads) chain_kernel(const ChainArgs args) {
    using C = ChainCfg<D, RPT>;
    extern __shared__ __align__(16) float smem[];
    float* wbuf = smem;                                   // [2][KC][D]
    float* slots = smem + 2 * KC * D;                     // [3][R][LD] then wide [R][LDW]
    auto slot_ptr = [&](int s) -> float* {
        return s == kChainWide ? slots + kChainSlots * C::R * C::LD : slots + s * C::R * C::LD;
    };
    auto slot_ld = [&](int s) -> int { return s == kChainWide ? C::LDW : C::LD; };

    const int t = threadIdx.x;
    const int cx = t % C::CX, ry = t / C::CX;
    const int row0 = blockIdx.x * C::R;
    const int n_rows = args.n_rows;

    for (int si = 0; si < args.n_stages; ++si) {
        const ChainStage& st = args.st[si];
        if (st.op == CH_LOAD) {
            float* d = slot_ptr(st.dst);
            const int ld = slot_ld(st.dst), w4 = st.width / 4;
            for (int f = t; f < C::R * w4; f += kChainThreads) {
                const int r = f / w4, c = (f % w4) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row0 + r < n_rows) {
                    v = ld4(st.g0 + (size_t)(row0 + r) * st.ld_g + c);
                    if (st.g1) v = v + ld4(st.g1 + (size_t)(row0 + r) * st.ld_g + c);
                    if (st.out_a) st4(st.out_a + (size_t)(row0 + r) * st.ld_out + c, v);
                }
                st4(d + r * ld + c, v);
            }
            __syncthreads();
        } else if (st.op == CH_HEADS_BWD) {
            // grad of o3 from the two heads: g_att * W + g_out * W_out.weight
            float* d = slot_ptr(st.dst);
            for (int f = t; f < C::R * (D / 4); f += kChainThreads) {
                const int r = f / (D / 4), c = (f % (D / 4)) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row0 + r < n_rows) {
                    const float ga = st.g0[row0 + r], go = st.g1[row0 + r];
                    const float4 w = ld4(st.W + c), wo = ld4(st.bias + c);
                    v = make_float4(ga * w.x + go * wo.x, ga * w.y + go * wo.y, ga * w.z + go * wo.z,
                                    ga * w.w + go * wo.w);
                }
                st4(d + r * C::LD + c, v);
            }
            __syncthreads();
        } else if (st.op == CH_DOT2) {
            // att = o . W (global_message_passing.py:47), out = o . W_out.weight + b (:48); one warp per row pair
            const float* s = slot_ptr(st.src);
            const int lane = t & 31, warp = t >> 5;
            for (int r = warp; r < C::R; r += kChainThreads / 32) {
                float a = 0.f, o = 0.f;
                for (int c = lane; c < D; c += 32) {
                    const float x = s[r * C::LD + c];
                    a = fmaf(x, st.W[c], a);
                    o = fmaf(x, st.bias[c], o);
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
            // stream chunk 0 of W while the prologue runs
            constexpr int NCHUNK = D / KC > 0 ? D / KC : 1;
            constexpr int KCE = D < KC ? D : KC;     // k rows per chunk (D = 16 -> 16)
            auto issue_chunk = [&](int kc, int buf) {
                const float* Wk = st.W + (size_t)kc * KCE * st.ldw;
                float* dstw = wbuf + buf * KC * D;
                for (int f = t; f < KCE * (D / 4); f += kChainThreads) {
                    const int r = f / (D / 4), c = (f % (D / 4)) * 4;
                    cp_async16(dstw + r * D + c, Wk + (size_t)r * st.ldw + c);
                }
                cp_async_commit();
            };
            issue_chunk(0, 0);

            const float* in = slot_ptr(st.src) + st.src_off;
            int in_ld = slot_ld(st.src);
            if (st.psrc >= 0) {
                float* p = slot_ptr(st.psrc);
                for (int f = t; f < C::R * (D / 4); f += kChainThreads) {
                    const int r = f / (D / 4), c = (f % (D / 4)) * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row0 + r < n_rows) {
                        const float4 g = ld4(in + r * in_ld + c);
                        v = g * dsilu4(ld4(st.zmul + (size_t)(row0 + r) * D + c));
                        if (st.save_src) st4(st.save_src + (size_t)(row0 + r) * D + c, v);
                    }
                    st4(p + r * C::LD + c, v);
                }
                in = p;
                in_ld = C::LD;
                // visibility of p is covered by the __syncthreads in the first chunk iteration
            }

            float acc[RPT][4];
#pragma unroll
            for (int i = 0; i < RPT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

            for (int kc = 0; kc < NCHUNK; ++kc) {
                cp_async_wait<0>();
                __syncthreads();
                if (kc + 1 < NCHUNK) issue_chunk(kc + 1, (kc + 1) & 1);
                const float* wk = wbuf + (kc & 1) * KC * D;
#pragma unroll 2
                for (int k4 = 0; k4 < KCE / 4; ++k4) {
                    float4 a[RPT];
#pragma unroll
                    for (int i = 0; i < RPT; ++i) a[i] = ld4(in + (ry * RPT + i) * in_ld + kc * KCE + k4 * 4);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const float4 w = ld4(wk + (k4 * 4 + kk) * D + cx * 4);
#pragma unroll
                        for (int i = 0; i < RPT; ++i) {
                            const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
                            acc[i][0] = fmaf(av, w.x, acc[i][0]);
                            acc[i][1] = fmaf(av, w.y, acc[i][1]);
                            acc[i][2] = fmaf(av, w.z, acc[i][2]);
                            acc[i][3] = fmaf(av, w.w, acc[i][3]);
                        }
                    }
                }
            }
            // epilogue
            const int c = cx * 4;
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (st.bias) b4 = ld4(st.bias + c);
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int r = ry * RPT + i;
                const bool live = row0 + r < n_rows;
                float4 v = make_float4(acc[i][0] + b4.x, acc[i][1] + b4.y, acc[i][2] + b4.z, acc[i][3] + b4.w);
                if (live && st.out_z) st4(st.out_z + (size_t)(row0 + r) * st.ld_out + c, v);
                if (st.act) v = silu4(v);
                if (st.add_slot >= 0) v = v + ld4(slot_ptr(st.add_slot) + r * C::LD + c);
                if (live && st.add_g) v = v + ld4(st.add_g + (size_t)(row0 + r) * st.ld_add + c);
                if (!live) v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (st.dst >= 0) st4(slot_ptr(st.dst) + r * C::LD + c, v);
                if (live && st.out_a) st4(st.out_a + (size_t)(row0 + r) * st.ld_out + c, v);
            }
            __syncthreads();
        }
    }
}

template <int D, int RPT>
static int chain_launch_t(const ChainArgs& args, cudaStream_t st) {
    using C = ChainCfg<D, RPT>;
    const size_t smem = C::smem_floats * sizeof(float);
    static bool configured = false;   // per (D, RPT) instantiation; attribute is per-function and idempotent
    if (!configured) {
        PAMNET_CUDA(cudaFuncSetAttribute(chain_kernel<D, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    chain_kernel<D, RPT><<<ceil_div(args.n_rows, C::R), kChainThreads, smem, st>>>(args);
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
    // more, smaller CTAs while the row count cannot fill the 148 SMs with 2 rows per thread
    switch (dim) {
        case 128: return (args.n_rows >= 148 * 16) ? chain_launch_t<128, 2>(args, st) : chain_launch_t<128, 1>(args, st);
        case 64:  return (args.n_rows >= 148 * 32) ? chain_launch_t<64, 2>(args, st) : chain_launch_t<64, 1>(args, st);
        case 32:  return (args.n_rows >= 148 * 64) ? chain_launch_t<32, 2>(args, st) : chain_launch_t<32, 1>(args, st);
        case 16:  return (args.n_rows >= 148 * 128) ? chain_launch_t<16, 2>(args, st) : chain_launch_t<16, 1>(args, st);
        default:
            set_error("chain: unsupported dim %d (16, 32, 64, 128)", dim);
            return -1;
    }
}

}  // namespace pamnet
