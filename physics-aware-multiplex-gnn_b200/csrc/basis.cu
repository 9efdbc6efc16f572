// Basis / embedding kernels (see basis.cuh).
#include "basis.cuh"

#include <stdlib.h>

namespace pamnet {

void make_sbf_tables(const pamnet_sbf_consts_t& c, SbfTables* t) {
    for (int i = 0; i < kNumSbf; ++i) {
        t->zeros[i] = c.zeros[i];
        t->norm[i] = c.norm[i];
    }
    // Legendre recurrence (utils/sbf.py:69-79) scaled by sqrt((2l+1)/4pi) (utils/sbf.py:62-66)
    double leg[kNumSph][kNumSph] = {};
    leg[0][0] = 1.0;
    leg[1][1] = 1.0;
    for (int j = 2; j < kNumSph; ++j)
        for (int p = 0; p <= j; ++p) {
            double v = 0.0;
            if (p >= 1) v += (2 * j - 1) * leg[j - 1][p - 1];
            v -= (j - 1) * leg[j - 2][p];
            leg[j][p] = v / j;
        }
    const double four_pi = 12.566370614359172;
    for (int l = 0; l < kNumSph; ++l)
        for (int p = 0; p < kNumSph; ++p) t->ycoef[l][p] = sqrt((2 * l + 1) / four_pi) * leg[l][p];
}

// Envelope u(x), exponent p = 5 (layers/basic.py:36-51; note x^p, not x^(p-1))
__device__ __forceinline__ float envelope_f(float x) {
    if (!(x < 1.0f)) return 0.0f;
    const float x5 = x * x * x * x * x;
    return 1.0f / x + (-21.0f) * x5 + 35.0f * x5 * x + (-15.0f) * x5 * x * x;
}
__device__ __forceinline__ double envelope_d(double x) {
    if (!(x < 1.0)) return 0.0;
    const double x5 = x * x * x * x * x;
    return 1.0 / x - 21.0 * x5 + 35.0 * x5 * x - 15.0 * x5 * x * x;
}

__global__ void rbf_kernel(const float* __restrict__ dist, int64_t n_edges, const float* __restrict__ freq,
                           float cutoff, float* __restrict__ rbf) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges * kNumRbf) return;
    const int64_t e = i / kNumRbf;
    const int n = (int)(i % kNumRbf);
    const float x = dist[e] / cutoff;
    rbf[i] = envelope_f(x) * sinf(freq[n] * x);
}

int rbf_forward(const float* dist, int64_t n_edges, const float* freq, float cutoff, float* rbf, cudaStream_t st) {
    if (n_edges == 0) return 0;
    prof_begin(KC_BASIS, 0.0, st);
    rbf_kernel<<<ceil_div(n_edges * kNumRbf, 256), 256, 0, st>>>(dist, n_edges, freq, cutoff, rbf);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// 16 columns x 16 row-lanes per block; fixed-order tree inside the block, one atomicAdd per block and column
__global__ void __launch_bounds__(256) rbf_freq_bwd_kernel(const float* __restrict__ dist, int64_t n_edges,
                                                           const float* __restrict__ freq, float cutoff,
                                                           const float* __restrict__ g_rbf, int64_t rows_per_block,
                                                           float* __restrict__ g_freq) {
    __shared__ float red[16][17];
    const int n = threadIdx.x & 15, rl = threadIdx.x >> 4;
    const int64_t e0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t e1 = min(n_edges, e0 + rows_per_block);
    const float f = freq[n];
    float s = 0.f;
    for (int64_t e = e0 + rl; e < e1; e += 16) {
        const float x = dist[e] / cutoff;
        s += g_rbf[e * kNumRbf + n] * envelope_f(x) * x * cosf(f * x);
    }
    red[rl][n] = s;
    __syncthreads();
    if (rl == 0) {
        float tot = 0.f;
        for (int r = 0; r < 16; ++r) tot += red[r][n];
        atomicAdd(&g_freq[n], tot);
    }
}

int rbf_freq_backward(const float* dist, int64_t n_edges, const float* freq, float cutoff, const float* g_rbf,
                      float* g_freq, cudaStream_t st) {
    if (n_edges == 0) return 0;
    const int64_t rpb = 128;
    prof_begin(KC_BASIS, 0.0, st);
    rbf_freq_bwd_kernel<<<ceil_div(n_edges, rpb), 256, 0, st>>>(dist, n_edges, freq, cutoff, g_rbf, rpb, g_freq);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// j_0..j_6 by upward recurrence in double (equivalent to the closed forms utils/sbf.py:29-38 generates; the
// reference evaluates those in fp32, where they lose up to 4 digits to cancellation -- SURVEY.md fact 5)
__global__ void sbf_radial_kernel(const SbfTables tab, const float* __restrict__ dist, int64_t n_edges, float cutoff,
                                  float* __restrict__ radial) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges * kNumRad) return;
    const int64_t e = i / kNumRad;
    const int m = (int)(i % kNumRad);
    const double x = (double)dist[e] / (double)cutoff;
    const double env = envelope_d(x);
#pragma unroll
    for (int l = 0; l < kNumSph; ++l) {
        const double a = tab.zeros[l * kNumRad + m] * x;
        double s, c;
        sincos(a, &s, &c);
        double jm = s / a, j = (s / a - c) / a;   // j_0, j_1
        double val = jm;
        if (l >= 1) val = j;
        for (int q = 1; q < l; ++q) {
            const double jn = (2 * q + 1) / a * j - jm;
            jm = j;
            j = jn;
            val = j;
        }
        radial[e * kNumSbf + l * kNumRad + m] = (float)(env * tab.norm[l * kNumRad + m] * val);
    }
}

int sbf_radial(const SbfTables& tab, const float* dist, int64_t n_edges, float cutoff, float* radial,
               cudaStream_t st) {
    if (n_edges == 0) return 0;
    prof_begin(KC_BASIS, 0.0, st);
    sbf_radial_kernel<<<ceil_div(n_edges * kNumRad, 128), 128, 0, st>>>(tab, dist, n_edges, cutoff, radial);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

__device__ __forceinline__ void zonal(const SbfTables& tab, double ct, float* out7) {
#pragma unroll
    for (int l = 0; l < kNumSph; ++l) {
        double v = 0.0;
#pragma unroll
        for (int p = kNumSph - 1; p >= 0; --p) v = v * ct + tab.ycoef[l][p];
        out7[l] = (float)v;
    }
}

__global__ void sbf_combine_kernel(const SbfTables tab, const float* __restrict__ radial,
                                   const float* __restrict__ angle, const int64_t* __restrict__ gather, int64_t n_trip,
                                   float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trip) return;
    float y[kNumSph];
    zonal(tab, cos((double)angle[t]), y);
    const float* r = radial + gather[t] * kNumSbf;
    for (int c = 0; c < kNumSbf; ++c) out[t * kNumSbf + c] = r[c] * y[c / kNumRad];
}

int sbf_combine(const SbfTables& tab, const float* radial, const float* angle, const int64_t* gather, int64_t n_trip,
                float* out, cudaStream_t st) {
    if (n_trip == 0) return 0;
    prof_begin(KC_BASIS, 0.0, st);
    sbf_combine_kernel<<<ceil_div(n_trip, 128), 128, 0, st>>>(tab, radial, angle, gather, n_trip, out);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// cos of the angle between u = p_b - p_a and v = p_c - p_b (models.py:165-177 take atan2(|u x v|, u.v) first)
__device__ __forceinline__ double cos_angle(const float* __restrict__ pos, int a, int b, int c) {
    const double ux = (double)pos[3 * b] - pos[3 * a], uy = (double)pos[3 * b + 1] - pos[3 * a + 1],
                 uz = (double)pos[3 * b + 2] - pos[3 * a + 2];
    const double vx = (double)pos[3 * c] - pos[3 * b], vy = (double)pos[3 * c + 1] - pos[3 * b + 1],
                 vz = (double)pos[3 * c + 2] - pos[3 * b + 2];
    const double dot = ux * vx + uy * vy + uz * vz;
    const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
    const double cr = sqrt(cx * cx + cy * cy + cz * cz);
    const double h = sqrt(dot * dot + cr * cr);
    return h > 0.0 ? dot / h : 1.0;   // atan2(0, 0) = 0
}

// one warp per triplet row: lanes cover the 88 columns
__global__ void __launch_bounds__(128) sbf_ext_kernel(const SbfTables tab, const int32_t* __restrict__ l_src,
                                                      const int32_t* __restrict__ l_dst,
                                                      const int32_t* __restrict__ t_ptr,
                                                      const int32_t* __restrict__ t_split,
                                                      const int32_t* __restrict__ t_gather,
                                                      const int32_t* __restrict__ t_owner, int64_t n_trip,
                                                      const float* __restrict__ pos, const float* __restrict__ radial,
                                                      float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (t >= n_trip) return;
    const int k = t_owner[t], p = t_gather[t];
    const int i = l_dst[k], j = l_src[k], o = l_src[p];
    const bool two_hop = (t - t_ptr[k]) < t_split[k];
    // two-hop: (idx_i, idx_j, idx_k) = (i, j, o); one-hop: (idx_i_pair, idx_j1_pair, idx_j2_pair) = (j, i, o)
    const double ct = two_hop ? cos_angle(pos, i, j, o) : cos_angle(pos, j, i, o);
    float y[kNumSph];
    zonal(tab, ct, y);
    const float* r = radial + (size_t)p * kNumSbf;
    float* row = out + t * kSbfExt;
    for (int c = lane; c < kSbfExt; c += 32) {
        float v = 0.f;
        if (c < 2 * kNumSbf) {
            const int cc = c < kNumSbf ? c : c - kNumSbf;
            if ((c < kNumSbf) == two_hop) v = r[cc] * y[cc / kNumRad];
        } else if (c == 2 * kNumSbf) {
            v = two_hop ? 1.f : 0.f;
        } else if (c == 2 * kNumSbf + 1) {
            v = two_hop ? 0.f : 1.f;
        }
        row[c] = v;
    }
}

int sbf_ext_forward(const SbfTables& tab, const Plan& plan, int64_t n_edges, int64_t n_trip, const float* pos,
                    const float* radial, float* sbf_ext, cudaStream_t st) {
    (void)n_edges;
    if (n_trip == 0) return 0;
    prof_begin(KC_BASIS, 0.0, st);
    sbf_ext_kernel<<<ceil_div(n_trip, 4), 128, 0, st>>>(tab, plan.l_src, plan.l_dst, plan.t_ptr, plan.t_split,
                                                        plan.t_gather, plan.t_owner, n_trip, pos, radial, sbf_ext);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Fused spherical-basis embedding for the small-dim models (dim <= 32; the RNA config has 4 M triplet rows):
// instead of materialising the [T, 88] extended operand (1.46 GB there, written once and read twice), the 42
// features radial[gathered edge] * Y_l0(angle) are formed in registers and multiplied by the weight block of the
// row's type (two-hop: mlp_sbf2, one-hop: mlp_sbf1; models.py:187-188) that sits in shared memory.  The seven
// zonal values per row are kept ([T, 8]) for the weight-gradient pass, which recomputes the features the same way.
// ---------------------------------------------------------------------------------------------
constexpr int kYsphLd = 8;

template <int DT>
__global__ void __launch_bounds__(128) sbf_embed_fwd_kernel(const SbfTables tab, const int32_t* __restrict__ l_src,
                                                            const int32_t* __restrict__ l_dst,
                                                            const int32_t* __restrict__ t_ptr,
                                                            const int32_t* __restrict__ t_split,
                                                            const int32_t* __restrict__ t_gather,
                                                            const int32_t* __restrict__ t_owner, int64_t n_trip,
                                                            const float* __restrict__ pos, const float* __restrict__ radial,
                                                            const float* __restrict__ w2, const float* __restrict__ b2,
                                                            const float* __restrict__ w1, const float* __restrict__ b1,
                                                            int dim, float* __restrict__ z_s, float* __restrict__ s_out,
                                                            float* __restrict__ ysph) {
    __shared__ float W[2][kNumSbf][DT];
    __shared__ float Bv[2][DT];
    for (int i = threadIdx.x; i < 2 * kNumSbf * DT; i += blockDim.x) {
        const int ty = i / (kNumSbf * DT), c = (i / DT) % kNumSbf, d = i % DT;
        const float* src = ty == 0 ? w2 : w1;
        (&W[0][0][0])[i] = (src && d < dim) ? src[d * kNumSbf + c] : 0.f;
    }
    for (int i = threadIdx.x; i < 2 * DT; i += blockDim.x) {
        const int ty = i / DT, d = i % DT;
        const float* src = ty == 0 ? b2 : b1;
        (&Bv[0][0])[i] = (src && d < dim) ? src[d] : 0.f;
    }
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trip) return;
    const int k = t_owner[t], p = t_gather[t];
    const int i = l_dst[k], j = l_src[k], o = l_src[p];
    const bool two_hop = (t - t_ptr[k]) < t_split[k];
    const double ct = two_hop ? cos_angle(pos, i, j, o) : cos_angle(pos, j, i, o);
    float y[kNumSph];
    zonal(tab, ct, y);
    st4(ysph + t * kYsphLd, make_float4(y[0], y[1], y[2], y[3]));
    const int ty = two_hop ? 0 : 1;
    st4(ysph + t * kYsphLd + 4, make_float4(y[4], y[5], y[6], (float)ty));      // slot 7: which of the two MLPs the triplet feeds
    const float* r = radial + (size_t)p * kNumSbf;
    float acc[DT];
#pragma unroll
    for (int d = 0; d < DT; ++d) acc[d] = 0.f;
#pragma unroll
    for (int l = 0; l < kNumSph; ++l) {
#pragma unroll
        for (int m = 0; m < kNumRad; ++m) {
            const int c = l * kNumRad + m;
            const float f = r[c] * y[l];
            const float* w = &W[ty][c][0];
#pragma unroll
            for (int d = 0; d < DT; ++d) acc[d] = fmaf(f, w[d], acc[d]);
        }
    }
    float* zr = z_s + t * dim;
    float* sr = s_out + t * dim;
#pragma unroll
    for (int d0 = 0; d0 < DT; d0 += 4) {
        if (d0 >= dim) break;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = acc[d0 + q] + Bv[ty][d0 + q];
        if (dim % 4 == 0) {
            st4(zr + d0, make_float4(v[0], v[1], v[2], v[3]));
            st4(sr + d0, make_float4(silu(v[0]), silu(v[1]), silu(v[2]), silu(v[3])));
        } else {
            for (int q = 0; q < 4; ++q)
                if (d0 + q < dim) { zr[d0 + q] = v[q]; sr[d0 + q] = silu(v[q]); }
        }
    }
}

bool sbf_fused_enabled(int dim) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("PAMNET_SBF_FUSED"); on = (e && e[0] == '0') ? 0 : 1; }
    return on && dim <= 32;
}

int sbf_embed_forward(const SbfTables& tab, const Plan& plan, int64_t n_trip, const float* pos, const float* radial,
                      const float* w2, const float* b2, const float* w1, const float* b1, int dim, float* z_s, float* s,
                      float* ysph, cudaStream_t st) {
    if (n_trip == 0) return 0;
    PAMNET_CHECK_ARG(dim <= 32, "sbf_embed_forward: dim=%d (<= 32)", dim);
    prof_begin(KC_BASIS, 0.0, st);
    const dim3 grid(ceil_div(n_trip, 128));
    if (dim <= 16)
        sbf_embed_fwd_kernel<16><<<grid, 128, 0, st>>>(tab, plan.l_src, plan.l_dst, plan.t_ptr, plan.t_split, plan.t_gather,
                                                       plan.t_owner, n_trip, pos, radial, w2, b2, w1, b1, dim, z_s, s, ysph);
    else
        sbf_embed_fwd_kernel<32><<<grid, 128, 0, st>>>(tab, plan.l_src, plan.l_dst, plan.t_ptr, plan.t_split, plan.t_gather,
                                                       plan.t_owner, n_trip, pos, radial, w2, b2, w1, b1, dim, z_s, s, ysph);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// weight gradients of mlp_sbf2 / mlp_sbf1: g_ext[d][col] = sum_t gz[t][d] * ext_t[col] over the 86 extended columns
// (ext_t recomputed per row).  A warp walks rows t; lane n owns columns n, n + 32, n + 64 for all dim output rows.
// fp32 partial sums cover at most 32 rows, then go into a PER-WARP fp64 tile in shared memory (plain read-modify-write:
// no atomics, every lane owns its addresses); the warps' tiles meet once per CTA and one fp32 atomic per output and CTA
// goes straight into the parameter-gradient buffers.  gz[t][:] is one coalesced load per row, its values reach the
// lanes by shuffle; the triplet's type comes from slot 7 of its zonal row (written by the forward kernel), so the only
// dependent load of a row is radial[t_gather[t]].
// (History, 951 k triplets of the RNA batch: shared fp64 atomics after every window, dim broadcast loads per row and
// lane, type looked up through t_owner -> t_ptr / t_split: 540 us.  fp64 accumulators in registers: 230 registers, one
// CTA per SM, 900 us -- the loop is latency-bound and occupancy is what hides it.)
constexpr int kSbfWgThreads = 128, kSbfWgUnroll = 4, kSbfWgFlush = 8;
template <int DT>
__global__ void __launch_bounds__(kSbfWgThreads, (DT <= 16 ? 3 : 1)) sbf_embed_wgrad_kernel(const int32_t* __restrict__ t_gather, int64_t n_trip,
                                                                     const float* __restrict__ radial,
                                                                     const float* __restrict__ ysph,
                                                                     const float* __restrict__ gz, int dim,
                                                                     float* __restrict__ gw2, float* __restrict__ gb2,
                                                                     float* __restrict__ gw1, float* __restrict__ gb1,
                                                                     int rows_per_cta) {
    // A row only touches the 42 weight columns + the bias of ITS type: lane n owns, per type, local columns n and n + 32
    // (local column 42 = the bias; lanes 11-31 have no second column).  Tile column = 42 type + local (bias: 84 + type).
    constexpr int NJ = 2, LDN = 96, NW = kSbfWgThreads / 32;
    extern __shared__ double dtile[];                  // [NW][DT][LDN]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* mine = dtile + (size_t)warp * DT * LDN;
    for (int i = lane; i < DT * LDN; i += 32) mine[i] = 0.0;
    __syncwarp();
    float acc[2][DT][NJ];
#pragma unroll
    for (int ty = 0; ty < 2; ++ty)
#pragma unroll
        for (int i = 0; i < DT; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[ty][i][j] = 0.f;
    const bool has2 = lane + 32 <= kNumSbf;            // second local column exists (43 local columns: 0..42)
    auto flush = [&]() {
#pragma unroll
        for (int ty = 0; ty < 2; ++ty)
#pragma unroll
            for (int i = 0; i < DT; ++i) {
                mine[i * LDN + ty * kNumSbf + lane] += (double)acc[ty][i][0];
                acc[ty][i][0] = 0.f;
                if (has2) {
                    const int lc = lane + 32;
                    mine[i * LDN + (lc < kNumSbf ? ty * kNumSbf + lc : 2 * kNumSbf + ty)] += (double)acc[ty][i][1];
                }
                acc[ty][i][1] = 0.f;
            }
    };
    const int64_t k0 = (int64_t)blockIdx.x * rows_per_cta, k1 = min(n_trip, k0 + (int64_t)rows_per_cta);
    // Two-deep software pipeline: the loads of the next kSbfWgUnroll rows are issued before the current rows are
    // multiplied (the loop body is one dependent load chain -- t_gather -> radial row -- followed by ~50 instructions per
    // row; at 12 warps per SM nothing else hides that latency: ncu 17 % warps active, 23 % issue slots).
    struct Rows { float bv[kSbfWgUnroll][NJ], ao[kSbfWgUnroll]; int ty[kSbfWgUnroll]; };
    auto fetch = [&](Rows& R, int64_t kb) {
#pragma unroll
        for (int u = 0; u < kSbfWgUnroll; ++u) {
            const int64_t t = kb + u * NW;
            const bool live = t < k1;
            const int64_t tt = live ? t : k0;
            const int p = t_gather[tt];
            const float* y = ysph + tt * kYsphLd;
            R.ty[u] = (int)y[7];
            const float* r = radial + (size_t)p * kNumSbf;
            R.bv[u][0] = live ? r[lane] * y[lane / kNumRad] : 0.f;
            const int lc = lane + 32;
            R.bv[u][1] = (live && has2) ? (lc < kNumSbf ? r[lc] * y[lc / kNumRad] : 1.f) : 0.f;
            R.ao[u] = (live && lane < dim) ? gz[tt * dim + lane] : 0.f;      // DT <= 32: one lane per gradient column
        }
    };
    int it = 0;
    Rows cur, nxt;
    if (k0 + warp < k1) fetch(cur, k0 + warp);
    for (int64_t kb = k0 + warp; kb < k1; kb += NW * kSbfWgUnroll) {
        const int64_t kn = kb + NW * kSbfWgUnroll;
        if (kn < k1) fetch(nxt, kn);
#pragma unroll
        for (int u = 0; u < kSbfWgUnroll; ++u) {
            if (cur.ty[u] == 0) {                       // warp-uniform: a row has one type
#pragma unroll
                for (int i = 0; i < DT; ++i) {
                    const float ai = __shfl_sync(0xffffffffu, cur.ao[u], i);
                    acc[0][i][0] = fmaf(ai, cur.bv[u][0], acc[0][i][0]);
                    acc[0][i][1] = fmaf(ai, cur.bv[u][1], acc[0][i][1]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < DT; ++i) {
                    const float ai = __shfl_sync(0xffffffffu, cur.ao[u], i);
                    acc[1][i][0] = fmaf(ai, cur.bv[u][0], acc[1][i][0]);
                    acc[1][i][1] = fmaf(ai, cur.bv[u][1], acc[1][i][1]);
                }
            }
        }
        if (++it == kSbfWgFlush) { flush(); it = 0; }
        if (kn < k1) cur = nxt;
    }
    flush();
    __syncthreads();
    for (int i = threadIdx.x; i < dim * (2 * kNumSbf + 2); i += kSbfWgThreads) {
        const int d = i / (2 * kNumSbf + 2), col = i % (2 * kNumSbf + 2);
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) tot += dtile[((size_t)w * DT + d) * LDN + col];
        const float v = (float)tot;
        if (v == 0.f) continue;
        if (col < kNumSbf) { if (gw2) atomicAdd(&gw2[d * kNumSbf + col], v); }
        else if (col < 2 * kNumSbf) atomicAdd(&gw1[d * kNumSbf + col - kNumSbf], v);
        else if (col == 2 * kNumSbf) { if (gb2) atomicAdd(&gb2[d], v); }
        else atomicAdd(&gb1[d], v);
    }
}

int sbf_embed_wgrad(const Plan& plan, int64_t n_trip, const float* radial, const float* ysph, const float* gz, int dim,
                    float* gw2, float* gb2, float* gw1, float* gb1, cudaStream_t st) {
    if (n_trip == 0) return 0;
    PAMNET_CHECK_ARG(dim <= 32, "sbf_embed_wgrad: dim=%d (<= 32)", dim);
    int ctas = ceil_div(n_trip, 256);
    if (ctas > 8 * kNumSM) ctas = 8 * kNumSM;
    const int rows = ceil_div(n_trip, ctas);
    const dim3 grid(ceil_div(n_trip, rows));
    const size_t smem16 = sizeof(double) * (kSbfWgThreads / 32) * 16 * 96, smem32 = 2 * smem16;
    prof_begin(KC_BASIS, 0.0, st);
    if (dim <= 16) {
        PAMNET_TRY(func_smem_once(reinterpret_cast<const void*>(sbf_embed_wgrad_kernel<16>), smem16));
        sbf_embed_wgrad_kernel<16><<<grid, kSbfWgThreads, smem16, st>>>(plan.t_gather, n_trip, radial, ysph, gz, dim, gw2, gb2,
                                                                         gw1, gb1, rows);
    } else {
        PAMNET_TRY(func_smem_once(reinterpret_cast<const void*>(sbf_embed_wgrad_kernel<32>), smem32));
        sbf_embed_wgrad_kernel<32><<<grid, kSbfWgThreads, smem32, st>>>(plan.t_gather, n_trip, radial, ysph, gz, dim, gw2, gb2,
                                                                         gw1, gb1, rows);
    }
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

__global__ void sbf_pack_kernel(int dim, const float* __restrict__ w2, const float* __restrict__ b2,
                                const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ w_ext) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim * kSbfExt) return;
    const int o = i / kSbfExt, c = i % kSbfExt;
    float v = 0.f;
    if (c < kNumSbf) v = w2 ? w2[o * kNumSbf + c] : 0.f;
    else if (c < 2 * kNumSbf) v = w1[o * kNumSbf + c - kNumSbf];
    else if (c == 2 * kNumSbf) v = b2 ? b2[o] : 0.f;
    else if (c == 2 * kNumSbf + 1) v = b1[o];
    w_ext[i] = v;
}

int sbf_weight_pack(int dim, const float* w2, const float* b2, const float* w1, const float* b1, float* w_ext,
                    cudaStream_t st) {
    prof_begin(KC_BASIS, 0.0, st);
    sbf_pack_kernel<<<ceil_div(dim * kSbfExt, 256), 256, 0, st>>>(dim, w2, b2, w1, b1, w_ext);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

__global__ void sbf_unpack_kernel(int dim, const float* __restrict__ g_ext, float* __restrict__ gw2,
                                  float* __restrict__ gb2, float* __restrict__ gw1, float* __restrict__ gb1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim * kSbfExt) return;
    const int o = i / kSbfExt, c = i % kSbfExt;
    const float v = g_ext[i];
    if (c < kNumSbf) { if (gw2) gw2[o * kNumSbf + c] = v; }
    else if (c < 2 * kNumSbf) gw1[o * kNumSbf + c - kNumSbf] = v;
    else if (c == 2 * kNumSbf) { if (gb2) gb2[o] = v; }
    else if (c == 2 * kNumSbf + 1) gb1[o] = v;
}

int sbf_weight_unpack_grad(int dim, const float* g_ext, float* gw2, float* gb2, float* gw1, float* gb1,
                           cudaStream_t st) {
    prof_begin(KC_BASIS, 0.0, st);
    sbf_unpack_kernel<<<ceil_div(dim * kSbfExt, 256), 256, 0, st>>>(dim, g_ext, gw2, gb2, gw1, gb1);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

__global__ void embed_fwd_kernel(const float* __restrict__ type_f, int64_t n_nodes, const float* __restrict__ emb,
                                 int n_embed, int dim, float* __restrict__ x) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes * dim) return;
    const int64_t n = i / dim;
    int ty = (int)type_f[n];                       // x_raw.long() (models.py:107,140)
    ty = ty < 0 ? 0 : (ty >= n_embed ? n_embed - 1 : ty);
    x[i] = emb[(size_t)ty * dim + i % dim];
}

int embed_forward(const float* type_f, int64_t n_nodes, const float* emb, int n_embed, int dim, float* x,
                  cudaStream_t st) {
    if (n_nodes == 0) return 0;
    prof_begin(KC_BASIS, 0.0, st);
    embed_fwd_kernel<<<ceil_div(n_nodes * dim, 256), 256, 0, st>>>(type_f, n_nodes, emb, n_embed, dim, x);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

// one block per (type, 32-column tile, node chunk): 8 node-strided partial sums per column, combined in fixed order; with
// more than one chunk (large graphs: the RNA batch has 15 816 nodes and 3 types x 16 columns -- three blocks scanning
// every node took 238 us) the chunks meet in fp32 atomics on the zero-initialised gradient
constexpr int kEmbChunk = 2048;
__global__ void __launch_bounds__(256) embed_bwd_kernel(const float* __restrict__ type_f, int64_t n_nodes,
                                                        const float* __restrict__ g_x, int dim,
                                                        float* __restrict__ g_emb) {
    __shared__ float red[8][33];
    const int ty = blockIdx.y;
    const int cl = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int64_t n0 = (int64_t)blockIdx.z * kEmbChunk, n1 = min(n_nodes, n0 + kEmbChunk);
    float s = 0.f;
    if (c < dim)
        for (int64_t n = n0 + part; n < n1; n += 8)
            if ((int)type_f[n] == ty) s += g_x[n * dim + c];
    red[part][cl] = s;
    __syncthreads();
    if (part == 0 && c < dim) {
        float tot = 0.f;
        for (int p = 0; p < 8; ++p) tot += red[p][cl];
        if (gridDim.z == 1) g_emb[(size_t)ty * dim + c] = tot;
        else atomicAdd(&g_emb[(size_t)ty * dim + c], tot);
    }
}

// g_emb must be zero on entry when n_nodes > kEmbChunk (model_backward zeroes the whole gradient buffer first)
int embed_backward(const float* type_f, int64_t n_nodes, const float* g_x, int n_embed, int dim, float* g_emb,
                   cudaStream_t st) {
    dim3 grid(ceil_div(dim, 32), n_embed, ceil_div(n_nodes > 0 ? n_nodes : 1, kEmbChunk));
    prof_begin(KC_BASIS, 0.0, st);
    embed_bwd_kernel<<<grid, 256, 0, st>>>(type_f, n_nodes, g_x, dim, g_emb);
    prof_end(st);
    PAMNET_LAUNCH_CHECK();
    return 0;
}

constexpr int kTransJobs = 128;
struct TransposeArgs {
    int n_jobs;
    int src_off[kTransJobs], dst_off[kTransJobs];
    short rows[kTransJobs], cols[kTransJobs];
    int ld[kTransJobs];
    char copy[kTransJobs];
};

// 32x32 tiles through shared memory; grid.y = job
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src_base, float* __restrict__ dst_base,
                                                        const TransposeArgs a) {
    __shared__ float tile[32][33];
    const int job = blockIdx.y;
    const int rows = a.rows[job], cols = a.cols[job], ld = a.ld[job];
    const float* src = src_base + a.src_off[job];
    float* dst = dst_base + a.dst_off[job];
    const int tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
    for (int tile_id = blockIdx.x; tile_id < tiles_r * tiles_c; tile_id += gridDim.x) {
        const int r0 = (tile_id / tiles_c) * 32, c0 = (tile_id % tiles_c) * 32;
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        for (int r = ty; r < 32; r += 8)
            if (r0 + r < rows && c0 + tx < cols) tile[r][tx] = src[(size_t)(r0 + r) * ld + c0 + tx];
        __syncthreads();
        if (a.copy[job]) {          // gather a strided block into a contiguous [rows][cols] matrix
            for (int r = ty; r < 32; r += 8)
                if (r0 + r < rows && c0 + tx < cols) dst[(size_t)(r0 + r) * cols + c0 + tx] = tile[r][tx];
        } else {
            for (int c = ty; c < 32; c += 8)
                if (c0 + c < cols && r0 + tx < rows) dst[(size_t)(c0 + c) * rows + r0 + tx] = tile[tx][c];
        }
        __syncthreads();
    }
}

int transpose_batch(const float* src_base, float* dst_base, const TransposeJob* jobs, int n_jobs, cudaStream_t st) {
    for (int j0 = 0; j0 < n_jobs; j0 += kTransJobs) {
        TransposeArgs a;
        a.n_jobs = (n_jobs - j0 < kTransJobs) ? n_jobs - j0 : kTransJobs;
        int max_tiles = 1;
        for (int j = 0; j < a.n_jobs; ++j) {
            const TransposeJob& jb = jobs[j0 + j];
            a.src_off[j] = (int)jb.src_off; a.dst_off[j] = (int)jb.dst_off;
            a.rows[j] = (short)jb.rows; a.cols[j] = (short)jb.cols; a.ld[j] = jb.ld; a.copy[j] = (char)(jb.copy != 0);
            const int tiles = ceil_div(jb.rows, 32) * ceil_div(jb.cols, 32);
            if (tiles > max_tiles) max_tiles = tiles;
        }
        dim3 grid(max_tiles > 16 ? 16 : max_tiles, a.n_jobs);
        prof_begin(KC_BASIS, 0.0, st);
        transpose_kernel<<<grid, 256, 0, st>>>(src_base, dst_base, a);
        prof_end(st);
        PAMNET_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace pamnet
