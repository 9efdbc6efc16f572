// Geometry -> basis kernels: Bessel RBF, spherical-Bessel x zonal-harmonic SBF (layers/basic.py:36-116,
// utils/sbf.py), atom-type embedding (models.py:107), plus their (few) parameter gradients.
#pragma once
#include "common.cuh"
#include "graph.cuh"

namespace pamnet {

constexpr int kSbfExt = 88;   // [sbf of two-hop rows (42) | sbf of one-hop rows (42) | 1{two-hop} | 1{one-hop} | 0 0]

struct SbfTables {
    double zeros[kNumSbf];
    double norm[kNumSbf];
    double ycoef[kNumSph][kNumSph];   // Y_l0 = sum_p ycoef[l][p] cos^p
};
void make_sbf_tables(const pamnet_sbf_consts_t& c, SbfTables* t);

// rbf[e, n] = u(d/c) * sin(freq_n * d/c)          (layers/basic.py:74-76)
int rbf_forward(const float* dist, int64_t n_edges, const float* freq, float cutoff, float* rbf, cudaStream_t st);
// g_freq[n] += sum_e g_rbf[e, n] * u(x) * x * cos(freq_n x)
int rbf_freq_backward(const float* dist, int64_t n_edges, const float* freq, float cutoff, const float* g_rbf,
                      float* g_freq, cudaStream_t st);
// radial[e, l*6+m] = u(x) N_lm j_l(z_lm x), evaluated in double (layers/basic.py:108-110)
int sbf_radial(const SbfTables& tab, const float* dist, int64_t n_edges, float cutoff, float* radial, cudaStream_t st);
// API form: out[t, 42] = radial[gather[t]] * Y_l0(angle[t])    (layers/basic.py:112-115)
int sbf_combine(const SbfTables& tab, const float* radial, const float* angle, const int64_t* gather, int64_t n_trip,
                float* out, cudaStream_t st);
// plan form: extended rows for the merged triplet list (angle evaluated in double from the positions)
int sbf_ext_forward(const SbfTables& tab, const Plan& plan, int64_t n_edges, int64_t n_trip, const float* pos,
                    const float* radial, float* sbf_ext, cudaStream_t st);

// fused small-dim path (dim <= 32): embedding straight from (radial, angle) without the [T, 88] operand, and its
// weight gradient; ysph [T, 8] carries the zonal values from forward to backward
bool sbf_fused_enabled(int dim);
int sbf_embed_forward(const SbfTables& tab, const Plan& plan, int64_t n_trip, const float* pos, const float* radial,
                      const float* w2, const float* b2, const float* w1, const float* b1, int dim, float* z_s, float* s,
                      float* ysph, cudaStream_t st);
int sbf_embed_wgrad(const Plan& plan, int64_t n_trip, const float* radial, const float* ysph, const float* gz, int dim,
                    float* gw2, float* gb2, float* gw1, float* gb1, cudaStream_t st);

// W_ext [D, 88] = [W_sbf2 | W_sbf1 | b_sbf2 | b_sbf1 | 0 0]; and the reverse scatter of its gradient
int sbf_weight_pack(int dim, const float* w2, const float* b2, const float* w1, const float* b1, float* w_ext,
                    cudaStream_t st);
int sbf_weight_unpack_grad(int dim, const float* g_ext, float* gw2, float* gb2, float* gw1, float* gb1,
                           cudaStream_t st);

int embed_forward(const float* type_f, int64_t n_nodes, const float* emb, int n_embed, int dim, float* x,
                  cudaStream_t st);
int embed_backward(const float* type_f, int64_t n_nodes, const float* g_x, int n_embed, int dim, float* g_emb,
                   cudaStream_t st);

// dst[c*rows + r] = src[r*ld + c] for a batch of matrices (transposed weights for the forward node chain)
struct TransposeJob { int64_t src_off, dst_off; int rows, cols, ld; int copy; };   // copy != 0: strided -> contiguous, no transpose
int transpose_batch(const float* src_base, float* dst_base, const TransposeJob* jobs_host, int n_jobs, cudaStream_t st);

}  // namespace pamnet
