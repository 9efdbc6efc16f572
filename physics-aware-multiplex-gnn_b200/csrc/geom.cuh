// Geometry and search helpers shared by the generic graph kernels (graph.cu) and the per-molecule front end
// (front_mol.cuh).  One definition, so both paths evaluate distances and angles with the same instruction sequence.
// The header also compiles with a plain host compiler (tests/host_emul: serial emulation of front_mol.cuh).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define PAMNET_HD __device__ __forceinline__
#else
#define PAMNET_HD static inline
// host emulation only (compiled with -ffp-contract=off): the explicitly rounded device intrinsics are plain operations
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
#endif

namespace pamnet {

PAMNET_HD int64_t lower_bound_i64(const int64_t* __restrict__ a, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// canonical squared distance: ((dx*dx)+(dy*dy))+(dz*dz), no FMA contraction (oracle/graph_ops.py:_d2_block)
PAMNET_HD float canon_d2(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// edge length exactly as models.py:64-65 evaluates it on the API tensors: sqrt(sum((pos[i]-pos[j])^2))
PAMNET_HD float edge_len(const float* __restrict__ pos, int64_t a, int64_t b) {
    float dx = pos[3 * a] - pos[3 * b], dy = pos[3 * a + 1] - pos[3 * b + 1], dz = pos[3 * a + 2] - pos[3 * b + 2];
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

PAMNET_HD float bond_angle(const float* __restrict__ pos, int a, int b, int c) {
    // models.py:165-177: u = pos[b]-pos[a], v = pos[c]-pos[b]; atan2(|u x v|, u.v)
    const float ux = pos[3 * b] - pos[3 * a], uy = pos[3 * b + 1] - pos[3 * a + 1], uz = pos[3 * b + 2] - pos[3 * a + 2];
    const float vx = pos[3 * c] - pos[3 * b], vy = pos[3 * c + 1] - pos[3 * b + 1], vz = pos[3 * c + 2] - pos[3 * b + 2];
    const float dot = ux * vx + uy * vy + uz * vz;
    const float cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
    return atan2f(sqrtf(cx * cx + cy * cy + cz * cz), dot);
}

}  // namespace pamnet
