// Row-local chains of D x D Linear(+SiLU)(+residual) stages fused into one kernel.
//
// Everything a PAMNet layer does to node features between two message-passing steps is row-local
// (global_message_passing.py:35,39-48; local_message_passing.py:43,55-64): mlp_x2, three Res blocks, mlp_out,
// the two readout heads, then the next layer's mlp_x1 and the per-node halves of its edge-MLP
// (W [x_i ; x_j ; e] = W_i x_i + W_j x_j + W_e e).  A CTA owns R rows, keeps the activations in shared
// memory and streams the weights; the reference issues one cuBLAS + two elementwise launches per stage.
// The same interpreter runs the data-gradient chain in reverse (stage prologue multiplies by SiLU').
#pragma once
#include "common.cuh"

namespace pamnet {

enum ChainOp : int {
    CH_LOAD = 0, CH_GEMM = 1, CH_DOT2 = 2, CH_HEADS_BWD = 3,
    // Gather stages (tensor-core interpreter only): the node-level segment sums that used to be separate launches in front
    // of a chain, run as the chain's loading stage -- one launch and one [N, D] round trip less per half.
    CH_GMSG_FWD = 4,     // h[n] = x1[n] + sum_{k in in(n)} SiLU(P_i[n] + P_j[src k] + Q[k]) * Tt[k]   (global_message_passing.py:52-56,38)
    CH_LMSG_FWD = 5,     // h[n] = x1[n] + sum_{k in in(n)} msum[k] * Rout[k]                          (local_message_passing.py:53-54)
    CH_GATHER_BWD = 6,   // g_P[n][task] = sum of the per-edge gradient blocks over n's incoming / outgoing slots (wide slot)
};

constexpr int kChainMaxStages = 26;
constexpr int kChainSlots = 3;      // D-wide slots 0..2; slot 3 is the wide (4D) staging slot
constexpr int kChainWide = 3;

struct ChainStage {
    int op;
    int src, src_off;   // GEMM / DOT2 input slot and column offset inside it
    int psrc;           // GEMM: slot that receives src * silu'(zmul) (== src: in place; -1: no prologue)
    int dst;            // output slot, -1 = none
    int add_slot;       // slot added to the output after the activation, -1 = none
    int act;            // 1 = SiLU
    int width;          // LOAD: columns
    int ldw, ld_out, ld_add, ld_g;
    int next_gemm;      // index of the next CH_GEMM stage (-1: none); filled in by chain_launch
    // filled in by chain_launch when the NEXT stage is a GEMM whose prologue (src * silu'(zmul) -> psrc, save_src) reads
    // this stage's output: the multiplication then happens here, in the epilogue, on values that are still in registers
    int post_dst;       // slot that receives out * silu'(post_zmul), -1 = none
    union { const float* post_zmul; const int32_t* o_ptr; };   // (gather stages: outgoing CSR, see below)
    union { float* post_save; const int32_t* o_pos; };
    const float* W;     // GEMM: [D(k)][D(n)] k-major (transposed weight in forward, weight itself in backward)
    const float* bias;
    union { const float* zmul; const int32_t* i_ptr; };         // (gather stages: incoming CSR)
    union { float* save_src; const int32_t* i_src; };           // prologue result written to global (grad wrt pre-activation, for weight gradients)
    float* out_z;       // pre-activation written to global
    float* out_a;       // final value written to global
    const float* add_g; // global tensor added to the output after the activation
    const float* g0;    // LOAD: source; HEADS_BWD: grad_att [N]; DOT2: pointer to W_out.bias
    const float* g1;    // LOAD: optional second addend; HEADS_BWD: grad_out [N]
    // gather stages (always stage 0 of a chain): i_ptr / i_src = CSR of the nodes' incoming slots and the source node per
    // slot, o_ptr / o_pos = outgoing CSR and the slot of the k-th outgoing edge (unions above); GMSG_FWD: g0 = P [N, 2D],
    // g1 = x1, W = Q|Tt rows (ldw); LMSG_FWD: g0 = msum [E, D], g1 = x1, W = Rout rows (ldw); GATHER_BWD: W = per-edge
    // gradient rows (ldw), width = n_blocks * 2 * D, out_a = g_P
};

struct ChainArgs {
    int n_rows;
    int n_stages;
    int small_footprint; // 1: run the two-CTAs-per-SM variant whatever the row count (chains off the critical path: the readout
                         // heads run beside the layer loop, and 78 + 78 one-per-SM CTAs do not fit on 148 SMs)
    int precision;      // 0: fp32-accurate (3xTF32 / FFMA); 1: single-pass TF32 (tensor-core interpreter only; chain_launch sets it
                        // from PAMNET_NODE_MLP=tf32 -- the reduced-precision node-MLP path of BASELINE.json configs[2])
    ChainStage st[kChainMaxStages];
};

int chain_launch(int dim, const ChainArgs& args, cudaStream_t st);
int chain_trace_read(long long* out, int n);

// Tensor-core interpreter (chain_mma.cu; D = 64, 128 unless PAMNET_CHAIN=ffma).  Its GEMM stages read `W` as a
// FRAGMENT IMAGE of the stage's A[m][k] matrix (m = output feature, k = input feature) produced by frag_batch; callers
// (model.cu) pick the weight format with chain_mma_enabled(dim).
bool chain_mma_enabled(int dim);
int chain_mma_launch(int dim, const ChainArgs& prepared_args, double alg_bytes, cudaStream_t st);
struct FragJob { int64_t src_off, dst_off; int ld; int trans; };   // trans 0: A[m][k] = src[m*ld+k]; 1: A[m][k] = src[k*ld+m]
int frag_batch(const float* src_base, float* dst_base, int dim, const FragJob* jobs_host, int n_jobs, cudaStream_t st);

}  // namespace pamnet
