// fp32 (FFMA) batched GEMM used for the x-independent dense contractions of the hot path:
// edge / triplet embeddings, per-layer edge projections for all layers at once, and their backward
// (data-grad with K segmented over layers, weight-grad with split-K over edges).
#pragma once
#include "common.cuh"

namespace pamnet {

enum GemmMode : int {
    GEMM_NT = 0,   // C[M,N] = A[M,K] * B[N,K]^T      forward of nn.Linear (B = weight [out,in])
    GEMM_NN = 1,   // C[M,N] = A[M,K] * B[K,N]        data gradient      (B = weight [out,in], K = out)
    GEMM_TN = 2,   // C[M,N] = A[K,M]^T * B[K,N]      weight gradient    (A = grad_z [rows,out], B = input [rows,in])
};

enum GemmEpi : int {
    EPI_NONE = 0,
    EPI_BIAS = 1,        // + bias[n]
    EPI_BIAS_SILU = 2,   // z = acc + bias[n] (bias may be null); C2 = z (if C2); C = silu(z) (if C)
    EPI_MUL_DSILU = 3,   // C = acc * silu'(Z[m,n])
};

constexpr int kGemmMaxSlots = 32;
constexpr int kGemmMaxSeg = 32;

struct GemmSlot {
    const float* A;
    const float* B;
    const float* bias;
    const float* Z;
    float* C;
    float* C2;
    int lda, ldb, ldc, ldz;
    int m;                // > 0: this slot has only m (< GemmArgs::M) rows; 0: GemmArgs::M
    int pad_;
};

struct GemmArgs {
    int M, N, K;
    int mode, epi;
    int accumulate;       // C += (plain read-modify-write; slots/tiles never overlap)
    int ksplit;           // > 1: split K across blockIdx.y, results combined with atomicAdd into C (EPI_NONE only)
    int nslots;           // blockIdx.z
    int precision;        // 0 / 3: fp32-accurate (3xTF32 on the tensor-core paths); 1: single-pass TF32 (gemm_tc2 only, opt-in)
    // K segmentation (GEMM_NN only): K = nseg*seg_len, segment s multiplies B = seg_B[s] (leading dim seg_ldb[s])
    int nseg, seg_len;
    GemmSlot slot[kGemmMaxSlots];
    const float* seg_B[kGemmMaxSeg];
    int seg_ldb[kGemmMaxSeg];
};

int gemm_launch(const GemmArgs& args, cudaStream_t st);

// tensor-core (tcgen05, 3xTF32) path, gemm_tc.cu; chosen by gemm_launch unless PAMNET_GEMM=ffma
bool gemm_tc_eligible(const GemmArgs& a);
int gemm_tc_launch(const GemmArgs& a, cudaStream_t st);
int tc_trace_read(long long* out, int n);

// persistent TMA-fed tcgen05 kernel (gemm_tc2.cu): the default tensor-core path; PAMNET_GEMM=tc1 selects gemm_tc.cu
bool gemm_tc2_eligible(const GemmArgs& a);
bool gemm_tc2_slot_ok(const GemmArgs& a, int slot);
int gemm_tc2_launch(const GemmArgs& a, cudaStream_t st);
void gemm_set_stable_range(const float* lo, size_t n_floats);    // parameter buffer of the model call in progress (this thread)

// skinny problems (N <= 32 rows-kernel, M <= 32 weight gradients), gemm_small.cu
bool gemm_small_eligible(const GemmArgs& a);
int gemm_small_launch(const GemmArgs& a, cudaStream_t st);

int mul_dsilu_launch(float* c, const float* z, int64_t n, cudaStream_t st);

// out[c] += sum_r X[r*ld + c]  (bias gradients); out must be zero-initialised by the caller
int colsum_launch(const float* X, int64_t rows, int cols, int ld, float* out, cudaStream_t st);

}  // namespace pamnet
