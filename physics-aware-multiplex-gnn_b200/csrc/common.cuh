// Shared device/host helpers for libpamnet_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pamnet_b200.h"

namespace pamnet {

constexpr int kNumRbf = 16;        // models.py:37-38
constexpr int kNumSph = 7;         // models.py:22 num_spherical
constexpr int kNumRad = 6;         // models.py:22 num_radial
constexpr int kNumSbf = kNumSph * kNumRad;
constexpr int kFeatPdb = 18;       // models.py:35 init_linear in-features
constexpr int kNumSM = 148;        // B200

void set_error(const char* fmt, ...);
const char* get_error();

#define PAMNET_CHECK_ARG(cond, ...)                     \
    do {                                                \
        if (!(cond)) {                                  \
            ::pamnet::set_error(__VA_ARGS__);           \
            return -1;                                  \
        }                                               \
    } while (0)

#define PAMNET_CUDA(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ::pamnet::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                                __FILE__, __LINE__);                                        \
            return (int)_e;                                                                 \
        }                                                                                   \
    } while (0)

// ---- launch accounting / optional per-kernel-class event timing (bench.py roofline; off by default) ----
enum KernelClass : int {
    KC_GEMM = 0, KC_CHAIN, KC_GLOBAL_MSG_FWD, KC_GLOBAL_MSG_BWD, KC_LOCAL_EDGE_FWD, KC_LOCAL_MSG_FWD,
    KC_LOCAL_MSG_BWD, KC_LOCAL_TRIP_BWD, KC_NODE_GATHER, KC_BASIS, KC_GRAPH, KC_READOUT, KC_MISC, KC_COUNT
};
void prof_begin(int cls, double alg_bytes, cudaStream_t st);   // call right before a kernel launch
void prof_end(cudaStream_t st);                                // call right after it
void prof_flops(double flops);                                 // optional, right after prof_begin: fp32-equivalent flops
void count_launch();

#define PAMNET_LAUNCH_CHECK()                 \
    do {                                      \
        ::pamnet::count_launch();             \
        PAMNET_CUDA(cudaGetLastError());      \
    } while (0)

#define PAMNET_TRY(expr)            \
    do {                            \
        int _r = (expr);            \
        if (_r != 0) return _r;     \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device); thread-safe (abi.cu)
int func_smem_once(const void* func, size_t bytes);

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------------
// The layer loop is ~90 short dependent kernels on one stream; between two of them the GPU idles for the launch
// latency (~3 us, 13 % of the step in the CUDA-event timeline).  Kernels of that loop are launched with the
// programmatic-stream-serialization attribute: the next kernel's CTAs may become resident while the previous one
// drains, run their prologue (shared-memory carve-up, barrier init, weight prefetch) and then block in
// griddepcontrol.wait until the predecessor has completed and its writes are visible.  PAMNET_PDL=0 disables it,
// PAMNET_PDL=2 keeps it for the layer loop only (not for the GEMMs of the auxiliary streams).
bool pdl_enabled();
int pdl_level();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#ifdef __CUDACC__
// block until the preceding kernel of the stream has completed (no-op for ordinary launches)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// allow the next kernel of the stream to start launching (it still waits for this grid in its own pdl_wait)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// device math
// ---------------------------------------------------------------------------------------------
// sigmoid for SiLU / SiLU'.  PAMNET_SILU_MODE (build time, build.py):
//   0  exact expf + IEEE division (round 1): ncu attributed 27 % of the global message kernel's stall samples and 22 %
//      of the node chain's to this line -- the division alone is ~10 dependent instructions with a slow-path branch;
//   1  (default) exact expf, reciprocal by MUFU.RCP (__fdividef, <= 2 ulp): one more ulp per evaluation;
//   2  ex2.approx + MUFU.RCP (relative error of __expf grows with |x|).
// The parity ladder of tests/helpers.py is the gate for the default.
#ifndef PAMNET_SILU_MODE
#define PAMNET_SILU_MODE 1
#endif
#if PAMNET_SILU_MODE == 2
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
#elif PAMNET_SILU_MODE == 1
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + expf(-x)); }
#else
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
#endif
// SiLU and its derivative (layers/basic.py:11-16): s(z)*(1 + z*(1-s(z)))
__device__ __forceinline__ float silu(float z) { return z * sigmoidf_(z); }
__device__ __forceinline__ float dsilu(float z) {
    float s = sigmoidf_(z);
    return s * (1.0f + z * (1.0f - s));
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 operator*(float4 a, float4 b) {
    return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 silu4(float4 z) { return make_float4(silu(z.x), silu(z.y), silu(z.z), silu(z.w)); }
__device__ __forceinline__ float4 dsilu4(float4 z) {
    return make_float4(dsilu(z.x), dsilu(z.y), dsilu(z.z), dsilu(z.w));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Row vector held by one warp: D floats, lane owns VPL = D/32 consecutive-by-stride chunks.
// D = 128 -> one float4 per lane (a 512 B row is one fully coalesced request); D = 64 -> float2;
// D = 32 -> float; D = 16 -> float on the lower half-warp.
// ---------------------------------------------------------------------------------------------
template <int D>
struct RowVec {
    // D >= 64: a warp per row (float4 / float2 per lane).  D <= 32: D / 4 lanes per row, one float4 each, so a warp
    // carries 32 / LPR rows side by side (dim 16: 8 rows) instead of idling half its lanes on 4-byte accesses.
    static constexpr int V = (D >= 128 || D <= 32) ? 4 : 2;        // floats per lane per chunk
    static constexpr int LPR = (D <= 32) ? D / 4 : 32;              // lanes per row
    static constexpr int RPW = 32 / LPR;                            // rows per warp
    static constexpr int C = (D + LPR * V - 1) / (LPR * V);         // chunks per lane
    float v[C * V];

    __device__ __forceinline__ static bool active(int) { return true; }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < C * V; ++i) v[i] = 0.f;
    }
    __device__ __forceinline__ void load(const float* __restrict__ row, int lane) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float* p = row + (c * LPR + lane % LPR) * V;
            if (V == 4) {
                float4 t = ld4(p);
                v[c * 4 + 0] = t.x; v[c * 4 + 1] = t.y; v[c * 4 + 2] = t.z; v[c * 4 + 3] = t.w;
            } else if (V == 2) {
                float2 t = *reinterpret_cast<const float2*>(p);
                v[c * 2 + 0] = t.x; v[c * 2 + 1] = t.y;
            } else {
                v[c] = active(lane) ? *p : 0.f;
            }
        }
    }
    // row[...] += v with fp32 reductions in L2 (order-free: callers accept run-to-run rounding differences)
    __device__ __forceinline__ void atomic_add(float* __restrict__ row, int lane) const {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float* p = row + (c * LPR + lane % LPR) * V;
            if (V == 4) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[c * 4 + 0]), "f"(v[c * 4 + 1]),
                             "f"(v[c * 4 + 2]), "f"(v[c * 4 + 3]) : "memory");
            } else if (V == 2) {
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[c * 2 + 0]), "f"(v[c * 2 + 1]) : "memory");
            } else {
                if (active(lane)) atomicAdd(p, v[c]);
            }
        }
    }
    __device__ __forceinline__ void store(float* __restrict__ row, int lane) const {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float* p = row + (c * LPR + lane % LPR) * V;
            if (V == 4) {
                st4(p, make_float4(v[c * 4 + 0], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]));
            } else if (V == 2) {
                *reinterpret_cast<float2*>(p) = make_float2(v[c * 2 + 0], v[c * 2 + 1]);
            } else {
                if (active(lane)) *p = v[c];
            }
        }
    }
};

}  // namespace pamnet
