// Forward / backward drivers for PAMNet.forward (models.py:100-224) and its autograd backward
// (main_qm9.py:110), composed from the kernels in graph/basis/gemm/chain/message/readout.
//
// Structure (DESIGN.md "Execution schedule"):
//   phase A  x-independent dense work, once per step and for ALL layers at once: basis -> edge/triplet
//            embeddings (models.py:180-188) -> per-layer edge projections Q|Tt (global), Qji|Qkj|R|Rout (local)
//            and the triplet gate MLP (local_message_passing.py:17,49);
//   phase B  the sequential layer loop (models.py:196-204): per layer two message kernels + two node chains;
//   readout  models.py:206-224.
// Backward mirrors it: phase B reversed (data gradients, per-edge gradients written once), then phase A'
// turns the accumulated per-edge / per-triplet gradients into weight gradients with a few large GEMMs.
#include "model.cuh"

#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "basis.cuh"
#include "chain.cuh"
#include "comm.cuh"
#include "gemm.cuh"
#include "graph.cuh"
#include "message.cuh"
#include "readout.cuh"

namespace pamnet {

// ---------------------------------------------------------------------------------------------
// parameter layout == reference state_dict order, every tensor 128 B aligned
// ---------------------------------------------------------------------------------------------
namespace {
struct LayoutBuilder {
    ModelP* mp;
    int64_t cur = 0;
    int64_t take(int64_t n) {
        const int64_t off = cur;
        mp->offsets.push_back(off);
        mp->numel.push_back(n);
        cur = (cur + n + kParamAlign - 1) / kParamAlign * kParamAlign;
        return off;
    }
    Lin lin(int64_t out, int64_t in, bool bias = true) {
        Lin l;
        l.w = take(out * in);
        if (bias) l.b = take(out);
        return l;
    }
};
}  // namespace

int build_param_layout(const pamnet_config_t& cfg, ModelP* mp) {
    PAMNET_CHECK_ARG(cfg.dim == 16 || cfg.dim == 32 || cfg.dim == 64 || cfg.dim == 128,
                     "dim=%d unsupported (16, 32, 64, 128)", cfg.dim);
    PAMNET_CHECK_ARG(cfg.n_layer >= 1 && cfg.n_layer <= kMaxLayers, "n_layer=%d unsupported (1..%d)", cfg.n_layer,
                     kMaxLayers);
    PAMNET_CHECK_ARG(cfg.dataset >= PAMNET_QM9 && cfg.dataset <= PAMNET_RNA, "bad dataset id %d", cfg.dataset);
    PAMNET_CHECK_ARG(!cfg.simple || cfg.dataset == PAMNET_QM9, "PAMNet_s is QM9-only (models.py:286-287)");
    const int64_t D = cfg.dim;
    *mp = ModelP();
    LayoutBuilder b{mp};
    const bool rna = cfg.dataset == PAMNET_RNA;
    mp->n_embed = rna ? 3 : 5;                                   // models.py:31-34
    mp->emb = b.take(mp->n_embed * D);
    if (!rna && !cfg.simple) mp->init_linear = b.take(D * kFeatPdb);   // models.py:35
    mp->freq_g = b.take(kNumRbf);
    mp->freq_l = b.take(kNumRbf);
    mp->rbf_g = b.lin(D, kNumRbf);
    mp->rbf_l = b.lin(D, kNumRbf);
    mp->sbf1 = b.lin(D, kNumSbf);
    if (!cfg.simple) mp->sbf2 = b.lin(D, kNumSbf);
    auto res_and = [&](HalfP& h) {
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 2; ++s) h.res[r][s] = b.lin(D, D);
    };
    auto heads = [&](HalfP& h) {
        for (int s = 0; s < 3; ++s) h.out[s] = b.lin(D, D);
        h.W_out = b.lin(1, D);
    };
    for (int l = 0; l < cfg.n_layer; ++l) {                      // global_message_passing.py:13-26
        HalfP& h = mp->g[l];
        h.W = b.take(D);
        h.x1 = b.lin(D, D);
        h.x2 = b.lin(D, D);
        res_and(h);
        h.m = b.lin(D, 3 * D);
        h.We = b.lin(D, D, false);
        heads(h);
    }
    for (int l = 0; l < cfg.n_layer; ++l) {                      // local_message_passing.py:13-29
        HalfP& h = mp->l[l];
        h.W = b.take(D);
        h.x1 = b.lin(D, D);
        h.m_ji = b.lin(D, 3 * D);
        h.m_kj = b.lin(D, 3 * D);
        h.sbf[0] = b.lin(D, D);
        h.sbf[1] = b.lin(D, D);
        h.lin_rbf = b.lin(D, D, false);
        res_and(h);
        h.lin_rbf_out = b.lin(D, D, false);
        h.x2 = b.lin(D, D);
        heads(h);
    }
    mp->total = b.cur;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
namespace {

struct HalfWs {
    // prepared weights of the node chains.  FFMA interpreter: k-major (transposed) copies for the forward chain, projB =
    // the per-node blocks of the edge-MLP weights gathered contiguous for the backward chain, every other backward
    // stage reads the parameter itself (*B == nullptr).  Tensor-core interpreter (chain_mma.cu): fragment images of
    // A = W (forward: *T) and A = W^T (backward: *B) for every stage.
    float *x1T, *x2T, *resT[3][2], *outT[3], *projT;
    float *x1B, *x2B, *resB[3][2], *outB[3];
    float *projB;                           // [nP][D][D]
    // saved activations, each [N, D] unless noted
    float *P;                               // [N, nP*D]
    float *z_x1, *x1, *h, *z_x2, *a_x2;
    float *z_A[3], *a_A[3], *z_B[3], *r[3];
    float *z_o[3], *a_o[3];
    float *m_nb, *msum;                     // local only, [El, D]
    // backward
    float *gz_x1, *gz_x2, *gz_A[3], *gz_B[3], *gz_o[3];
    float *g_P;                             // [N, nP*D] grad of P (kept per half: consumed on the auxiliary stream)
    float *g_heads;                         // [N, D] grad of the half's x output through its two readout heads
};

struct Ws {
    float* wt;
    float *x0, *rbf_g, *rbf_l, *radial, *sbf_ext, *w_ext, *gw_ext;
    float* ysph;        // [T, 8] zonal values per triplet (fused small-dim spherical-basis path only)
    float *z_eg, *e_g, *z_el, *e_l, *z_s, *s;
    float *QT, *QR, *zq1, *aq1, *zq2;
    float *att, *out, *node_val;
    HalfWs half[2 * kMaxLayers];
    // backward
    float *g_att, *g_out, *g_h, *g_resx, *g_P, *g_s, *g_x0;
    float *gQT, *gQR, *gzq2, *gzq1, *gz_s, *gz_eg, *gz_el, *g_rbf_g, *g_rbf_l;
};

// PAMNET_GATHER_FUSE=0: node-level segment sums as their own launches again (the round-1 schedule)
inline bool gather_fusion_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("PAMNET_GATHER_FUSE"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

// PAMNET_HEADS_SMALL=0: head chains as one-CTA-per-SM launches again
inline int heads_small_footprint() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("PAMNET_HEADS_SMALL"); on = (e && e[0] == '0') ? 0 : 1; }
    return on;
}

// half index hh = 2*l (global layer l) or 2*l+1 (local layer l)
inline bool is_local(int hh) { return hh & 1; }
inline int nP_of(int hh) { return is_local(hh) ? 4 : 2; }

// `wbase` (optional): an external blob that holds the prepared-weights region (the first section of the layout,
// which depends on the configuration only); `weights_bytes` (optional) receives that region's size.
size_t ws_layout(const pamnet_config_t& cfg, const pamnet_sizes_t& sz, void* base, Ws* out, void* wbase = nullptr,
                 size_t* weights_bytes = nullptr) {
    size_t off = 0;
    void* cur_base = wbase ? wbase : base;
    auto take = [&](int64_t n) {
        float* p = cur_base ? reinterpret_cast<float*>(static_cast<char*>(cur_base) + off) : nullptr;
        off += align_up(sizeof(float) * (size_t)(n > 0 ? n : 1));
        return p;
    };
    const int64_t N = sz.n_nodes, Eg = sz.n_edges_g, El = sz.n_edges_l, T = sz.n_t2 + sz.n_t1;
    const int64_t D = cfg.dim, L = cfg.n_layer, H = 2 * L;
    Ws w;
    memset(&w, 0, sizeof(w));
    w.wt = nullptr;
    for (int hh = 0; hh < H; ++hh) {      // transposed weights first: their offsets stay small
        HalfWs& h = w.half[hh];
        h.x1T = take(D * D); h.x2T = take(D * D);
        for (int r = 0; r < 3; ++r) for (int s = 0; s < 2; ++s) h.resT[r][s] = take(D * D);
        for (int s = 0; s < 3; ++s) h.outT[s] = take(D * D);
        h.projT = take(nP_of(hh) * D * D);
        h.projB = take(nP_of(hh) * D * D);
        if (chain_mma_enabled((int)D)) {
            h.x1B = take(D * D); h.x2B = take(D * D);
            for (int r = 0; r < 3; ++r) for (int s = 0; s < 2; ++s) h.resB[r][s] = take(D * D);
            for (int s = 0; s < 3; ++s) h.outB[s] = take(D * D);
        }
    }
    if (weights_bytes) *weights_bytes = off;
    cur_base = base;
    w.x0 = take(N * D);
    w.rbf_g = take(Eg * kNumRbf); w.rbf_l = take(El * kNumRbf); w.radial = take(El * kNumSbf);
    const bool sbf_fused = sbf_fused_enabled((int)D);
    w.sbf_ext = take(sbf_fused ? 0 : T * kSbfExt); w.w_ext = take(D * kSbfExt); w.gw_ext = take(D * kSbfExt);
    w.ysph = take(sbf_fused ? T * 8 : 0);
    w.z_eg = take(Eg * D); w.e_g = take(Eg * D); w.z_el = take(El * D); w.e_l = take(El * D);
    w.z_s = take(T * D); w.s = take(T * D);
    w.QT = take(Eg * L * 2 * D); w.QR = take(El * L * 4 * D);
    w.zq1 = take(T * L * D); w.aq1 = take(T * L * D); w.zq2 = take(T * L * D);
    w.att = take(H * N); w.out = take(H * N); w.node_val = take(N);
    for (int hh = 0; hh < H; ++hh) {
        HalfWs& h = w.half[hh];
        h.P = take(N * nP_of(hh) * D);
        h.z_x1 = take(N * D); h.x1 = take(N * D); h.h = take(N * D); h.z_x2 = take(N * D); h.a_x2 = take(N * D);
        for (int r = 0; r < 3; ++r) {
            h.z_A[r] = take(N * D); h.a_A[r] = take(N * D); h.z_B[r] = take(N * D); h.r[r] = take(N * D);
        }
        for (int s = 0; s < 3; ++s) { h.z_o[s] = take(N * D); h.a_o[s] = take(N * D); }
        if (is_local(hh)) { h.m_nb = take(El * D); h.msum = take(El * D); }
        h.gz_x1 = take(N * D); h.gz_x2 = take(N * D);
        for (int r = 0; r < 3; ++r) { h.gz_A[r] = take(N * D); h.gz_B[r] = take(N * D); }
        for (int s = 0; s < 3; ++s) h.gz_o[s] = take(N * D);
        h.g_P = take(N * nP_of(hh) * D);
        h.g_heads = take(N * D);
    }
    w.g_att = take(H * N); w.g_out = take(H * N);
    w.g_h = take(N * D); w.g_resx = take(N * D); w.g_P = take(N * 4 * D); w.g_s = take(El * D); w.g_x0 = take(N * D);
    w.gQT = take(Eg * L * 2 * D); w.gQR = take(El * L * 4 * D);
    w.gzq2 = take(T * L * D); w.gzq1 = take(T * L * D); w.gz_s = take(T * D);
    w.gz_eg = take(Eg * D); w.gz_el = take(El * D); w.g_rbf_g = take(Eg * kNumRbf); w.g_rbf_l = take(El * kNumRbf);
    if (out) *out = w;
    return off;
}

const HalfP& half_params(const ModelP& mp, int hh) { return is_local(hh) ? mp.l[hh >> 1] : mp.g[hh >> 1]; }

// ---- small builders ---------------------------------------------------------------------------
ChainStage stage_zero() {
    ChainStage s;
    memset(&s, 0, sizeof(s));
    s.psrc = -1; s.dst = -1; s.add_slot = -1;
    return s;
}
ChainStage st_load(int dst, const float* g0, int ld_g, int width, const float* g1 = nullptr) {
    ChainStage s = stage_zero();
    s.op = CH_LOAD; s.dst = dst; s.g0 = g0; s.g1 = g1; s.ld_g = ld_g; s.width = width;
    return s;
}
ChainStage st_gemm(int src, int dst, const float* W, int ldw, const float* bias, int act) {
    ChainStage s = stage_zero();
    s.op = CH_GEMM; s.src = src; s.dst = dst; s.W = W; s.ldw = ldw; s.bias = bias; s.act = act;
    return s;
}

struct Prog {
    ChainArgs a;
    Prog(int n_rows) { memset(&a, 0, sizeof(a)); a.n_rows = n_rows; }
    ChainStage& add(const ChainStage& s) { a.st[a.n_stages] = s; return a.st[a.n_stages++]; }
};

// forward: x1 = SiLU(mlp_x1(x)) (global_message_passing.py:35 / local_message_passing.py:43) and the per-node
// halves of the edge MLPs.  `src` holds x; uses slot (src+2)%3 for x1.
void add_pre_fwd(Prog& p, const float* params, const HalfP& hp, const HalfWs& hw, int hh, int D, int src) {
    const int x1s = (src + 2) % 3;
    ChainStage& a = p.add(st_gemm(src, x1s, hw.x1T, D, params + hp.x1.b, 1));
    a.out_z = hw.z_x1; a.out_a = hw.x1; a.ld_out = D;
    const int nP = nP_of(hh);
    for (int c = 0; c < nP; ++c) {
        ChainStage& g = p.add(st_gemm(x1s, -1, hw.projT + (size_t)c * D * D, D, nullptr, 0));
        g.out_a = hw.P + c * D; g.ld_out = nP * D;
    }
}

// forward update block (global_message_passing.py:39-44); leaves x_out in slot 1 (and in hw.r[2])
void add_post_fwd(Prog& p, const float* params, const HalfP& hp, const HalfWs& hw, int D, const float* res_x,
                  const ChainStage* gather = nullptr) {
    if (gather) p.add(*gather);            // the half's node-level segment sum as the chain's loading stage (writes hw.h too)
    else p.add(st_load(0, hw.h, D, D));
    { ChainStage& s = p.add(st_gemm(0, 1, hw.x2T, D, params + hp.x2.b, 1)); s.out_z = hw.z_x2; s.out_a = hw.a_x2; s.ld_out = D; }
    int cur = 1;                                    // slot holding the Res input
    for (int r = 0; r < 3; ++r) {
        const int tmp = (cur + 2) % 3, nxt = (cur + 1) % 3;
        { ChainStage& s = p.add(st_gemm(cur, tmp, hw.resT[r][0], D, params + hp.res[r][0].b, 1));
          s.out_z = hw.z_A[r]; s.out_a = hw.a_A[r]; s.ld_out = D; }
        { ChainStage& s = p.add(st_gemm(tmp, nxt, hw.resT[r][1], D, params + hp.res[r][1].b, 1));
          s.add_slot = cur; s.out_z = hw.z_B[r]; s.out_a = hw.r[r]; s.ld_out = D;
          if (r == 0) { s.add_g = res_x; s.ld_add = D; } }
        cur = nxt;
    }
    // cur == 1 after three rotations (1 -> 2 -> 0 -> 1)
}

// the two readout heads of a half (global_message_passing.py:46-48): o = mlp_out(x_out), att = o . W, out = W_out o + b.
// Nothing in the layer loop depends on them, so they run as their own small chain off the critical path.
void add_heads_fwd(Prog& p, const float* params, const HalfP& hp, const HalfWs& hw, int D, float* att, float* out) {
    p.add(st_load(0, hw.r[2], D, D));
    int a = 0;
    for (int s3 = 0; s3 < 3; ++s3) {
        const int d = a ^ 1;
        ChainStage& s = p.add(st_gemm(a, d, hw.outT[s3], D, params + hp.out[s3].b, 1));
        s.out_z = hw.z_o[s3]; s.out_a = hw.a_o[s3]; s.ld_out = D;
        a = d;
    }
    ChainStage d2 = stage_zero();
    d2.op = CH_DOT2; d2.src = a; d2.W = params + hp.W; d2.bias = params + hp.W_out.w; d2.g0 = params + hp.W_out.b;
    d2.out_z = att; d2.out_a = out;
    p.add(d2);
}

// backward of add_pre_fwd for half hh: g_x1 = g_h + g_P . W_proj ; g_x = (g_x1 * SiLU'(z_x1)) . W_x1 + g_resx
// result in slot 2 (and in `g_x_out` if non-null)
void add_pre_bwd(Prog& p, const float* params, const HalfP& hp, const HalfWs& hw, int hh, int D, const Ws& w,
                 float* g_x_out, const ChainStage* gather = nullptr) {
    const int nP = nP_of(hh);
    if (gather) p.add(*gather);            // grad of P gathered from the per-edge gradients right here (writes hw.g_P too)
    else p.add(st_load(kChainWide, hw.g_P, nP * D, nP * D));
    p.add(st_load(0, w.g_h, D, D));
    int cur = 0;
    for (int c = 0; c < nP; ++c) {
        // W_c = block c of the edge-MLP weight ([out][in block]); the forward pass gathered it contiguous (projB)
        ChainStage& s = p.add(st_gemm(kChainWide, cur ^ 1, hw.projB + (size_t)c * D * D, D, nullptr, 0));
        s.src_off = c * D; s.add_slot = cur;
        cur ^= 1;
    }
    // nP is even -> cur == 0 holds g_x1
    ChainStage& s = p.add(st_gemm(cur, 2, hw.x1B ? hw.x1B : params + hp.x1.w, D, nullptr, 0));
    s.psrc = cur; s.zmul = hw.z_x1; s.save_src = hw.gz_x1; s.add_g = w.g_resx; s.ld_add = D;
    s.out_a = g_x_out; s.ld_out = D;
}

// backward of add_heads_fwd: grad of the half's x output through the heads -> hw.g_heads (and the grad_z of mlp_out).
// Depends only on the readout gradient, so all halves run up front, off the critical path.
void add_heads_bwd(Prog& p, const float* params, const HalfP& hp, const HalfWs& hw, int D, const float* g_att,
                   const float* g_out, float* gp) {
    ChainStage hb = stage_zero();
    hb.op = CH_HEADS_BWD; hb.dst = 0; hb.g0 = g_att; hb.g1 = g_out; hb.W = params + hp.W; hb.bias = params + hp.W_out.w;
    // the heads' own weight gradients (global_message_passing.py:47-48): dW = o3^T g_att, dW_out = o3^T g_out, db_out = sum g_out
    hb.zmul = hw.a_o[2]; hb.out_z = gp + hp.W; hb.out_a = gp + hp.W_out.w; hb.save_src = gp + hp.W_out.b;
    p.add(hb);
    auto bwd = [&](int src, int dst, const float* z, float* save, const float* W) -> ChainStage& {
        ChainStage& s = p.add(st_gemm(src, dst, W, D, nullptr, 0));
        s.psrc = src; s.zmul = z; s.save_src = save;
        return s;
    };
    auto wb = [&](const float* image, int64_t off) { return image ? image : params + off; };
    bwd(0, 1, hw.z_o[2], hw.gz_o[2], wb(hw.outB[2], hp.out[2].w));
    bwd(1, 0, hw.z_o[1], hw.gz_o[1], wb(hw.outB[1], hp.out[1].w));
    { ChainStage& s = bwd(0, 1, hw.z_o[0], hw.gz_o[0], wb(hw.outB[0], hp.out[0].w)); s.out_a = hw.g_heads; s.ld_out = D; }
}

// backward of add_post_fwd; expects grad wrt x_out (from the next half) in slot 2 when has_gx; writes g_h and g_resx
void add_post_bwd(Prog& p, const float* params, const HalfP& hp, const HalfWs& hw, int D, const Ws& w, bool has_gx) {
    auto bwd = [&](int src, int psrc, int dst, const float* z, float* save, const float* W, int add_slot) -> ChainStage& {
        ChainStage& s = p.add(st_gemm(src, dst, W, D, nullptr, 0));
        s.psrc = psrc; s.zmul = z; s.save_src = save; s.add_slot = add_slot;
        return s;
    };
    auto wb = [&](const float* image, int64_t off) { return image ? image : params + off; };
    { ChainStage& s = p.add(st_load(1, hw.g_heads, D, D)); s.add_slot = has_gx ? 2 : -1; }   // slot1 = g_r3
    bwd(1, 0, 2, hw.z_B[2], hw.gz_B[2], wb(hw.resB[2][1], hp.res[2][1].w), -1);
    bwd(2, 2, 0, hw.z_A[2], hw.gz_A[2], wb(hw.resB[2][0], hp.res[2][0].w), 1);                 // slot0 = g_r2
    bwd(0, 1, 2, hw.z_B[1], hw.gz_B[1], wb(hw.resB[1][1], hp.res[1][1].w), -1);
    { ChainStage& s = bwd(2, 2, 1, hw.z_A[1], hw.gz_A[1], wb(hw.resB[1][0], hp.res[1][0].w), 0);    // slot1 = g_r1
      s.out_a = w.g_resx; s.ld_out = D; }
    bwd(1, 0, 2, hw.z_B[0], hw.gz_B[0], wb(hw.resB[0][1], hp.res[0][1].w), -1);
    bwd(2, 2, 0, hw.z_A[0], hw.gz_A[0], wb(hw.resB[0][0], hp.res[0][0].w), 1);                 // slot0 = g_x2
    { ChainStage& s = bwd(0, 0, 1, hw.z_x2, hw.gz_x2, wb(hw.x2B, hp.x2.w), -1);
      s.out_a = w.g_h; s.ld_out = D; }
}

GemmArgs gemm_zero(int mode, int epi, int M, int N, int K) {
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = mode; a.epi = epi; a.M = M; a.N = N; a.K = K; a.ksplit = 1;
    return a;
}
GemmSlot slot(const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias = nullptr,
              float* C2 = nullptr, const float* Z = nullptr, int ldz = 0) {
    GemmSlot s;
    s.A = A; s.B = B; s.bias = bias; s.Z = Z; s.C = C; s.C2 = C2; s.lda = lda; s.ldb = ldb; s.ldc = ldc; s.ldz = ldz;
    s.m = 0; s.pad_ = 0;
    return s;
}
// split-K factor of the weight-gradient GEMMs: ~128 reduction rows per CTA keeps the dependent k-loop short
// (one tile fetch is ~1 us of L2 latency) and gives the small [D, D] outputs enough CTAs to fill the GPU
int pick_ksplit(int64_t K) {
    int64_t s = (K + 127) / 128;
    return (int)(s < 1 ? 1 : (s > 64 ? 64 : s));
}

// launch a list of slots sharing one GemmArgs header, 32 at a time
int gemm_multi(GemmArgs a, const std::vector<GemmSlot>& slots, cudaStream_t st) {
    for (size_t i = 0; i < slots.size(); i += kGemmMaxSlots) {
        const size_t n = slots.size() - i < (size_t)kGemmMaxSlots ? slots.size() - i : kGemmMaxSlots;
        a.nslots = (int)n;
        for (size_t j = 0; j < n; ++j) a.slot[j] = slots[i + j];
        PAMNET_TRY(gemm_launch(a, st));
    }
    return 0;
}

struct Ctx {
    pamnet_config_t cfg;
    pamnet_sizes_t sz;
    ModelP mp;
    Plan plan;
    Ws w;
    SbfTables tab;
    int D, L, H;
    int64_t N, G, Eg, El, T;
};

int make_ctx(const pamnet_config_t& cfg, const pamnet_sizes_t& sz, const pamnet_sbf_consts_t& sbf, void* plan_base,
             void* plan_trip, void* workspace, size_t ws_bytes, Ctx* c, void* prepared = nullptr) {
    PAMNET_TRY(build_param_layout(cfg, &c->mp));
    c->cfg = cfg; c->sz = sz;
    c->D = cfg.dim; c->L = cfg.n_layer; c->H = 2 * cfg.n_layer;
    c->N = sz.n_nodes; c->G = sz.n_graphs; c->Eg = sz.n_edges_g; c->El = sz.n_edges_l; c->T = sz.n_t2 + sz.n_t1;
    PAMNET_CHECK_ARG(c->N > 0 && c->G > 0, "empty batch (n_nodes=%lld, n_graphs=%lld)", (long long)c->N, (long long)c->G);
    PAMNET_CHECK_ARG(c->N < (1 << 30) && c->Eg * (int64_t)c->L * 2 * c->D < (1ll << 40), "batch too large");
    const size_t need = ws_layout(cfg, sz, workspace, &c->w, prepared);
    PAMNET_CHECK_ARG(ws_bytes >= need, "workspace too small: %zu < %zu", ws_bytes, need);
    plan_layout(sz, plan_base, plan_trip, &c->plan, nullptr, nullptr);
    make_sbf_tables(sbf, &c->tab);
    return 0;
}

}  // namespace

size_t workspace_bytes(const pamnet_config_t& cfg, const pamnet_sizes_t& sz) { return ws_layout(cfg, sz, nullptr, nullptr); }

// ---------------------------------------------------------------------------------------------
// prepared weights: k-major copies of every chain linear and the contiguous projection blocks of the backward chain.
// They depend on the parameters only, so a caller may produce them on another stream while the graph is being built
// (prepare_weights) and hand the blob to model_forward / model_backward; without it model_forward makes them itself.
// ---------------------------------------------------------------------------------------------
size_t prepared_weights_bytes(const pamnet_config_t& cfg) {
    pamnet_sizes_t sz;
    memset(&sz, 0, sizeof(sz));
    size_t wb = 0;
    ws_layout(cfg, sz, nullptr, nullptr, nullptr, &wb);
    return wb;
}

namespace {
int launch_weight_prep(const pamnet_config_t& cfg, const ModelP& mp, const Ws& w, const float* params, float* dst_base,
                       cudaStream_t st) {
    const int D = cfg.dim, H = 2 * cfg.n_layer;
    if (chain_mma_enabled(D)) {
        // fragment images (chain_mma.cu): forward stages multiply by A = W ([out][in]), backward stages by A = W^T
        std::vector<FragJob> fj;
        auto img = [&](int64_t src_off, int ld, float* fwd, float* bwd) {
            fj.push_back(FragJob{src_off, (int64_t)(fwd - dst_base), ld, 0});
            fj.push_back(FragJob{src_off, (int64_t)(bwd - dst_base), ld, 1});
        };
        for (int hh = 0; hh < H; ++hh) {
            const HalfP& hp = half_params(mp, hh);
            const HalfWs& hw = w.half[hh];
            img(hp.x1.w, D, hw.x1T, hw.x1B);
            img(hp.x2.w, D, hw.x2T, hw.x2B);
            for (int r = 0; r < 3; ++r) for (int s = 0; s < 2; ++s) img(hp.res[r][s].w, D, hw.resT[r][s], hw.resB[r][s]);
            for (int s = 0; s < 3; ++s) img(hp.out[s].w, D, hw.outT[s], hw.outB[s]);
            for (int cblk = 0; cblk < nP_of(hh); ++cblk) {      // per-node blocks of the edge-MLP weights ([D][3D])
                int64_t woff;
                if (!is_local(hh)) woff = hp.m.w + cblk * D;
                else woff = (cblk < 2 ? hp.m_ji.w : hp.m_kj.w) + (cblk & 1) * D;
                img(woff, 3 * D, hw.projT + (size_t)cblk * D * D, hw.projB + (size_t)cblk * D * D);
            }
        }
        return frag_batch(params, dst_base, D, fj.data(), (int)fj.size(), st);
    }
    std::vector<TransposeJob> jobs;
    auto job = [&](int64_t src_off, int rows, int cols, int ld, float* dst) {
        jobs.push_back(TransposeJob{src_off, (int64_t)(dst - dst_base), rows, cols, ld, 0});
    };
    auto gather = [&](int64_t src_off, int rows, int cols, int ld, float* dst) {
        jobs.push_back(TransposeJob{src_off, (int64_t)(dst - dst_base), rows, cols, ld, 1});
    };
    for (int hh = 0; hh < H; ++hh) {
        const HalfP& hp = half_params(mp, hh);
        const HalfWs& hw = w.half[hh];
        job(hp.x1.w, D, D, D, hw.x1T);
        job(hp.x2.w, D, D, D, hw.x2T);
        for (int r = 0; r < 3; ++r) for (int s = 0; s < 2; ++s) job(hp.res[r][s].w, D, D, D, hw.resT[r][s]);
        for (int s = 0; s < 3; ++s) job(hp.out[s].w, D, D, D, hw.outT[s]);
        if (!is_local(hh)) {
            job(hp.m.w, D, 2 * D, 3 * D, hw.projT);                       // -> [2D, D]
        } else {
            job(hp.m_ji.w, D, 2 * D, 3 * D, hw.projT);
            job(hp.m_kj.w, D, 2 * D, 3 * D, hw.projT + (size_t)2 * D * D);
        }
        // blocks for the backward chain (the weights of a step do not change before its backward)
        for (int cblk = 0; cblk < nP_of(hh); ++cblk) {
            int64_t woff;
            if (!is_local(hh)) woff = hp.m.w + cblk * D;
            else woff = (cblk < 2 ? hp.m_ji.w : hp.m_kj.w) + (cblk & 1) * D;
            gather(woff, D, D, 3 * D, hw.projB + (size_t)cblk * D * D);
        }
    }
    return transpose_batch(params, dst_base, jobs.data(), (int)jobs.size(), st);
}
}  // namespace

int prepare_weights(const pamnet_config_t& cfg, const float* params, void* prepared, cudaStream_t st) {
    ModelP mp;
    PAMNET_TRY(build_param_layout(cfg, &mp));
    pamnet_sizes_t sz;
    memset(&sz, 0, sizeof(sz));
    Ws w;
    ws_layout(cfg, sz, prepared, &w, prepared);      // only the weight pointers of `w` are meaningful
    return launch_weight_prep(cfg, mp, w, params, reinterpret_cast<float*>(prepared), st);
}

// byte offset of a named workspace buffer (test / debugging aid; -1 = unknown name)
int64_t debug_ws_offset(const pamnet_config_t& cfg, const pamnet_sizes_t& sz, const char* name, int half) {
    Ws w;
    char* base = reinterpret_cast<char*>(0x1000);   // fake non-null base; only differences are used
    ws_layout(cfg, sz, base, &w);
    const HalfWs& h = w.half[half < 0 ? 0 : half];
    struct { const char* n; const float* p; } tab[] = {
        {"x0", w.x0}, {"rbf_g", w.rbf_g}, {"rbf_l", w.rbf_l}, {"radial", w.radial}, {"sbf_ext", w.sbf_ext},
        {"e_g", w.e_g}, {"e_l", w.e_l}, {"s", w.s}, {"QT", w.QT}, {"QR", w.QR}, {"zq1", w.zq1}, {"aq1", w.aq1},
        {"zq2", w.zq2}, {"att", w.att}, {"out", w.out}, {"node_val", w.node_val},
        {"P", h.P}, {"x1", h.x1}, {"h", h.h}, {"a_x2", h.a_x2}, {"r0", h.r[0]}, {"r1", h.r[1]}, {"r2", h.r[2]},
        {"a_o2", h.a_o[2]}, {"m_nb", h.m_nb}, {"msum", h.msum}, {"x1T", h.x1T},
        {"g_att", w.g_att}, {"g_out", w.g_out}, {"g_h", w.g_h}, {"g_P", h.g_P}, {"g_x0", w.g_x0},
        {"gQT", w.gQT}, {"gQR", w.gQR}, {"gzq2", w.gzq2}, {"gz_s", w.gz_s}, {"gz_eg", w.gz_eg}, {"gz_el", w.gz_el},
        {"gz_x1", h.gz_x1}, {"gz_x2", h.gz_x2}, {"gz_o2", h.gz_o[2]},
    };
    for (auto& e : tab)
        if (strcmp(e.n, name) == 0) return e.p ? (int64_t)(reinterpret_cast<const char*>(e.p) - base) : -1;
    return -1;
}

// ---------------------------------------------------------------------------------------------
// two-stream schedule
// ---------------------------------------------------------------------------------------------
// The x-independent dense work (phase A / A') only meets the sequential layer loop (phase B) at one point per
// layer, so it runs on an auxiliary stream supplied by the caller and overlaps the latency-bound node chains,
// which occupy about half of the SMs.  Events come from a small per-thread pool (created once, re-recorded).
namespace {
struct Sched {
    cudaStream_t st, s2, s3;     // main, auxiliary (large x-independent GEMMs), internal (small weight-gradient GEMMs)
    cudaStream_t s4, s5;         // internal: independent per-edge / per-triplet GEMM chains of backward run side by side
    bool dual;
    std::vector<cudaEvent_t>* pool;
    int next = 0;
    cudaEvent_t fresh() {
        if ((size_t)next == pool->size()) {
            cudaEvent_t e;
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            pool->push_back(e);
        }
        return (*pool)[next++];
    }
    // everything issued so far on `from` happens before anything issued later on `to`
    int order(cudaStream_t from, cudaStream_t to) {
        if (!dual) return 0;
        cudaEvent_t e = fresh();
        PAMNET_CUDA(cudaEventRecord(e, from));
        PAMNET_CUDA(cudaStreamWaitEvent(to, e, 0));
        return 0;
    }
    int record(cudaStream_t from, cudaEvent_t* e) {
        if (!dual) return 0;
        *e = fresh();
        PAMNET_CUDA(cudaEventRecord(*e, from));
        return 0;
    }
    int wait(cudaStream_t to, cudaEvent_t e) {
        if (!dual) return 0;
        PAMNET_CUDA(cudaStreamWaitEvent(to, e, 0));
        return 0;
    }
};
// Library-owned streams and events, per host thread AND per device (streams / events belong to the device that was
// current when they were created; forward and backward run on different host threads under autograd).
//  * small: the node-level weight gradients are 128 x 128 outputs over ~600 rows, a dozen CTAs per launch; on their own
//    stream they fill SMs the other streams leave idle instead of queueing behind the large per-edge GEMMs;
//  * extra[2]: independent per-edge / per-triplet GEMM families of backward side by side;
//  * main: the sequential layer loop is the critical path; its kernels need whole SMs (the node chain keeps ~190 KB of
//    shared memory per CTA) and otherwise queue behind GEMM CTAs whenever SMs free up -> highest priority.
// All of them are forked from / joined into the caller's stream with events before every return.
constexpr int kMaxDevices = 16;
struct DeviceStreams {
    cudaStream_t main = nullptr, small = nullptr, extra[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> events;
    bool tried = false;
};
thread_local DeviceStreams g_dev_streams[kMaxDevices];

DeviceStreams* device_streams() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    DeviceStreams& d = g_dev_streams[dev];
    if (!d.tried) {
        d.tried = true;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);     // hi = numerically lowest = greatest priority
        if (cudaStreamCreateWithPriority(&d.main, cudaStreamNonBlocking, hi) != cudaSuccess) d.main = nullptr;
        if (cudaStreamCreateWithFlags(&d.small, cudaStreamNonBlocking) != cudaSuccess) d.small = nullptr;
        for (int i = 0; i < 2; ++i)
            if (cudaStreamCreateWithFlags(&d.extra[i], cudaStreamNonBlocking) != cudaSuccess) d.extra[i] = nullptr;
    }
    return &d;
}
thread_local std::vector<cudaEvent_t> g_fallback_events;

// ---- gradient buckets for an overlapped data-parallel all-reduce ----------------------------------
// The flat gradient buffer is laid out per layer; the slice of layer-half h is final once the backward iteration of half
// h - 1 has issued its weight-gradient launches (it contributes mlp_x1 of half h).  When enabled, model_backward records
// one event per contributing stream at that point; wait_grad_bucket makes a caller's (communication) stream wait for
// them, so a per-bucket NCCL all-reduce can run while the remaining halves are still being differentiated.  The table
// is global per device (backward runs on autograd's worker thread, the collective is issued from the main thread).
constexpr int kBucketStreams = 4;
struct BucketEvents {
    std::vector<cudaEvent_t> ev;       // [(2 * kMaxLayers) * kBucketStreams]
    int halves = 0;                    // halves recorded by the last backward on this device
    bool recorded[2 * kMaxLayers] = {};
};
std::mutex g_bucket_mu;
BucketEvents g_buckets[kMaxDevices];
std::atomic<int> g_buckets_on{0};

int record_bucket(int half, int n_halves, cudaStream_t const (&streams)[kBucketStreams]) {
    int dev = 0;
    PAMNET_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return 0;
    std::lock_guard<std::mutex> lk(g_bucket_mu);
    BucketEvents& b = g_buckets[dev];
    if (b.ev.empty()) {
        b.ev.resize(2 * kMaxLayers * kBucketStreams);
        for (auto& e : b.ev) PAMNET_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    for (int i = 0; i < kBucketStreams; ++i) PAMNET_CUDA(cudaEventRecord(b.ev[half * kBucketStreams + i], streams[i]));
    b.recorded[half] = true;
    b.halves = n_halves;
    return 0;
}
}  // namespace

void set_grad_buckets(int on) { g_buckets_on.store(on ? 1 : 0); }

// make `stream` wait until the gradient slice of layer-half `half` (0 .. 2L-1: global l = 2l, local l = 2l+1) of the last
// model_backward on the current device is complete
int wait_grad_bucket(int half, cudaStream_t stream) {
    int dev = 0;
    PAMNET_CUDA(cudaGetDevice(&dev));
    PAMNET_CHECK_ARG(dev >= 0 && dev < kMaxDevices && half >= 0 && half < 2 * kMaxLayers, "wait_grad_bucket: bad half %d", half);
    std::lock_guard<std::mutex> lk(g_bucket_mu);
    BucketEvents& b = g_buckets[dev];
    PAMNET_CHECK_ARG(!b.ev.empty() && b.recorded[half], "wait_grad_bucket: no backward recorded bucket %d (enable with pamnet_grad_buckets(1))", half);
    for (int i = 0; i < kBucketStreams; ++i) PAMNET_CUDA(cudaStreamWaitEvent(stream, b.ev[half * kBucketStreams + i], 0));
    return 0;
}

// [lo, hi) of layer-half `half` in the flat parameter / gradient layout
int grad_bucket_range(const pamnet_config_t& cfg, int half, int64_t* lo, int64_t* hi) {
    ModelP mp;
    PAMNET_TRY(build_param_layout(cfg, &mp));
    const int L = cfg.n_layer;
    PAMNET_CHECK_ARG(half >= 0 && half < 2 * L, "grad_bucket_range: bad half %d", half);
    const int l = half >> 1;
    if (half & 1) {
        *lo = mp.l[l].W;
        *hi = l + 1 < L ? mp.l[l + 1].W : mp.total;
    } else {
        *lo = mp.g[l].W;
        *hi = l + 1 < L ? mp.g[l + 1].W : mp.l[0].W;
    }
    return 0;
}

namespace {
// group boundaries over the layers: a short first group (its results are needed first in forward, last in backward),
// then growing ones
std::vector<int> layer_groups(int L) {
    std::vector<int> g{0};
    int step = 1;
    while (g.back() < L) {
        g.push_back(g.back() + step > L ? L : g.back() + step);
        step = step < 3 ? step + 1 : 3;
    }
    return g;
}

Sched make_sched(cudaStream_t st, cudaStream_t aux) {
    Sched s;
    s.st = st;
    s.dual = aux != nullptr && aux != st;
    s.s2 = s.dual ? aux : st;
    s.s3 = s.s4 = s.s5 = st;
    DeviceStreams* d = device_streams();
    if (s.dual && d) {
        if (d->main) s.st = d->main;
        s.s3 = d->small ? d->small : s.s2;
        s.s4 = d->extra[0] ? d->extra[0] : s.s2;
        s.s5 = d->extra[1] ? d->extra[1] : s.s2;
    } else if (s.dual) {
        s.s3 = s.s4 = s.s5 = s.s2;
    }
    s.pool = d ? &d->events : &g_fallback_events;
    return s;
}
}  // namespace

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
int model_forward(const pamnet_config_t& cfg, const pamnet_sizes_t& sz, const pamnet_sbf_consts_t& sbf,
                  const float* params, const float* node_in, const float* sign, const float* pos, void* plan_base,
                  void* plan_trip, void* workspace, size_t ws_bytes, float* out, cudaStream_t st_user, cudaStream_t aux,
                  void* prepared) {
    Ctx c;
    PAMNET_TRY(make_ctx(cfg, sz, sbf, plan_base, plan_trip, workspace, ws_bytes, &c, prepared));
    const int D = c.D, L = c.L, H = c.H;
    const int64_t N = c.N, Eg = c.Eg, El = c.El, T = c.T;
    const ModelP& mp = c.mp;
    Ws& w = c.w;
    const Plan& pl = c.plan;
    Sched sc = make_sched(st_user, aux);
    cudaStream_t st = sc.st, s2 = sc.s2;
    gemm_set_stable_range(params, (size_t)mp.total);          // weight operands keep their addresses: tensor maps in the device table

    if (st != st_user) PAMNET_TRY(sc.order(st_user, st));     // the plan and the inputs were produced on the caller's stream
    PAMNET_TRY(sc.order(st, s2));

    // ================= auxiliary stream(s): phase A (models.py:180-188 and the x-independent halves of the layers).
    // Two independent chains: the global-edge one (s2) and the local-edge / triplet one (sB, = s2 by default).
    // Default: ONE auxiliary stream and everything of group 0 issued before the main stream's first kernels.
    // PAMNET_FWD_SPLIT=1 runs the local / triplet chain on a second stream and issues it after the main stream's first
    // kernels -- measured slower (1.96-2.03 vs 1.89-1.90 ms/step): the two chains then compete for the SMs the first
    // node chains need.
    // Large graphs (the 775 k-edge / 951 k-triplet RNA batch): the kernels of either chain fill the GPU for 50-200 us each
    // and the two chains are independent, so they do run side by side (auto: E_g + T >= 300 k).
    static int fwd_split_env = -2;
    if (fwd_split_env == -2) { const char* e = getenv("PAMNET_FWD_SPLIT"); fwd_split_env = e ? (e[0] == '1' ? 1 : 0) : -1; }
    const int fwd_split = fwd_split_env >= 0 ? fwd_split_env : ((Eg + T >= 300000) ? 1 : 0);
    cudaStream_t sB = fwd_split ? sc.s4 : s2;
    if (sB != s2) PAMNET_TRY(sc.order(st, sB));
    auto embed_global = [&]() -> int {
        PAMNET_TRY(rbf_forward(pl.dist_g, Eg, params + mp.freq_g, cfg.cutoff_g, w.rbf_g, s2));
        GemmArgs a = gemm_zero(GEMM_NT, EPI_BIAS_SILU, (int)Eg, D, kNumRbf);
        a.nslots = 1;
        a.slot[0] = slot(w.rbf_g, kNumRbf, params + mp.rbf_g.w, kNumRbf, w.e_g, D, params + mp.rbf_g.b, w.z_eg);
        return gemm_launch(a, s2);
    };
    auto embed_local = [&]() -> int {
        PAMNET_TRY(rbf_forward(pl.dist_l, El, params + mp.freq_l, cfg.cutoff_l, w.rbf_l, sB));
        PAMNET_TRY(sbf_radial(c.tab, pl.dist_l, El, cfg.cutoff_l, w.radial, sB));
        const bool fused = sbf_fused_enabled(D);
        const float* w2 = cfg.simple ? nullptr : params + mp.sbf2.w;
        const float* b2 = cfg.simple ? nullptr : params + mp.sbf2.b;
        if (!fused) {
            PAMNET_TRY(sbf_ext_forward(c.tab, pl, El, T, pos, w.radial, w.sbf_ext, sB));
            PAMNET_TRY(sbf_weight_pack(D, w2, b2, params + mp.sbf1.w, params + mp.sbf1.b, w.w_ext, sB));
        }
        GemmArgs a = gemm_zero(GEMM_NT, EPI_BIAS_SILU, (int)El, D, kNumRbf);
        a.nslots = 1;
        a.slot[0] = slot(w.rbf_l, kNumRbf, params + mp.rbf_l.w, kNumRbf, w.e_l, D, params + mp.rbf_l.b, w.z_el);
        PAMNET_TRY(gemm_launch(a, sB));
        if (fused)      // small dims: features formed in registers, no [T, 88] operand (basis.cu)
            return sbf_embed_forward(c.tab, pl, T, pos, w.radial, w2, b2, params + mp.sbf1.w, params + mp.sbf1.b, D, w.z_s,
                                     w.s, w.ysph, sB);
        a.M = (int)T; a.K = kSbfExt;
        a.slot[0] = slot(w.sbf_ext, kSbfExt, w.w_ext, kSbfExt, w.s, D, nullptr, w.z_s);
        return gemm_launch(a, sB);
    };
    // layers are grouped ([0,1), [1,3), [3,L) for L = 6): one batched GEMM launch per group and operand keeps the
    // tiles-per-launch high (a tensor-core tile has a ~15 us latency floor) while later groups overlap phase B
    std::vector<int> grp = layer_groups(L);
    const int ngrp = (int)grp.size() - 1;
    std::vector<cudaEvent_t> ev_g(ngrp), ev_l(ngrp);
    std::vector<int> group_of(L);
    for (int gi = 0; gi < ngrp; ++gi)
        for (int l = grp[gi]; l < grp[gi + 1]; ++l) group_of[l] = gi;
    auto issue_group_g = [&](int gi) -> int {
        const int l0 = grp[gi], l1 = grp[gi + 1];
        {   // global: Q | Tt = e_g [W_m,e ; W_e]^T (+ b_m)      (global_message_passing.py:52-56)
            const int ldq = L * 2 * D;
            std::vector<GemmSlot> sl;
            for (int l = l0; l < l1; ++l) {
                const HalfP& hp = mp.g[l];
                sl.push_back(slot(w.e_g, D, params + hp.m.w + 2 * D, 3 * D, w.QT + l * 2 * D, ldq, params + hp.m.b));
                sl.push_back(slot(w.e_g, D, params + hp.We.w, D, w.QT + l * 2 * D + D, ldq));
            }
            PAMNET_TRY(gemm_multi(gemm_zero(GEMM_NT, EPI_BIAS, (int)Eg, D, D), sl, s2));
            PAMNET_TRY(sc.record(s2, &ev_g[gi]));
        }
        return 0;
    };
    auto issue_group_l = [&](int gi) -> int {
        const int l0 = grp[gi], l1 = grp[gi + 1];
        {   // local: Qji | Qkj | R | Rout and the triplet gate MLP   (local_message_passing.py:46-53)
            const int ldq = L * 4 * D, ldt = L * D;
            std::vector<GemmSlot> sl, s1, s2v;
            for (int l = l0; l < l1; ++l) {
                const HalfP& hp = mp.l[l];
                float* base = w.QR + l * 4 * D;
                sl.push_back(slot(w.e_l, D, params + hp.m_ji.w + 2 * D, 3 * D, base, ldq, params + hp.m_ji.b));
                sl.push_back(slot(w.e_l, D, params + hp.m_kj.w + 2 * D, 3 * D, base + D, ldq, params + hp.m_kj.b));
                sl.push_back(slot(w.e_l, D, params + hp.lin_rbf.w, D, base + 2 * D, ldq));
                sl.push_back(slot(w.e_l, D, params + hp.lin_rbf_out.w, D, base + 3 * D, ldq));
                s1.push_back(slot(w.s, D, params + hp.sbf[0].w, D, w.aq1 + l * D, ldt, params + hp.sbf[0].b, w.zq1 + l * D));
                s2v.push_back(slot(w.aq1 + l * D, ldt, params + hp.sbf[1].w, D, w.zq2 + l * D, ldt, params + hp.sbf[1].b));
            }
            PAMNET_TRY(gemm_multi(gemm_zero(GEMM_NT, EPI_BIAS, (int)El, D, D), sl, sB));
            PAMNET_TRY(gemm_multi(gemm_zero(GEMM_NT, EPI_BIAS_SILU, (int)T, D, D), s1, sB));
            PAMNET_TRY(gemm_multi(gemm_zero(GEMM_NT, EPI_BIAS, (int)T, D, D), s2v, sB));
            PAMNET_TRY(sc.record(sB, &ev_l[gi]));
        }
        return 0;
    };
    // host issue order right after the plan sync: the global chain up to its first group, the main stream's first
    // kernels, then the local chain; group gi+1 is issued just before the main stream works through group gi
    PAMNET_TRY(embed_global());
    if (!fwd_split) PAMNET_TRY(embed_local());
    PAMNET_TRY(issue_group_g(0));
    if (!fwd_split) PAMNET_TRY(issue_group_l(0));

    // ================= main stream: node input, transposed chain weights, phase B
    if (prepared) PAMNET_TRY(sc.order(s2, st));      // the caller produced them on the auxiliary stream
    else PAMNET_TRY(launch_weight_prep(cfg, mp, w, params, reinterpret_cast<float*>(workspace), st));
    if (cfg.dataset == PAMNET_PDBBIND) {          // models.py:119
        GemmArgs a = gemm_zero(GEMM_NT, EPI_NONE, (int)N, D, kFeatPdb);
        a.nslots = 1;
        a.slot[0] = slot(node_in, kFeatPdb, params + mp.init_linear, kFeatPdb, w.x0, D);
        PAMNET_TRY(gemm_launch(a, st));
    } else {                                      // models.py:107,140
        PAMNET_TRY(embed_forward(node_in, N, params + mp.emb, mp.n_embed, D, w.x0, st));
    }
    {
        Prog p((int)N);
        p.add(st_load(2, w.x0, D, D));
        add_pre_fwd(p, params, half_params(mp, 0), w.half[0], 0, D, 2);
        PAMNET_TRY(chain_launch(D, p.a, st));
    }
    if (fwd_split) {
        PAMNET_TRY(embed_local());
        PAMNET_TRY(issue_group_l(0));
    }
    // node-level segment sums run as the loading stage of the chain that consumes them (tensor-core interpreter only)
    const bool fuse_gather = chain_mma_enabled(D) && gather_fusion_enabled();
    ChainStage gstage;
    for (int hh = 0; hh < H; ++hh) {              // models.py:196-204
        const HalfWs& hw = w.half[hh];
        const int l = hh >> 1;
        if (!is_local(hh) && l == grp[group_of[l]] && group_of[l] + 1 < ngrp) {
            PAMNET_TRY(issue_group_g(group_of[l] + 1));
            PAMNET_TRY(issue_group_l(group_of[l] + 1));
        }
        if (!is_local(hh)) {
            if (l == grp[group_of[l]]) PAMNET_TRY(sc.wait(st, ev_g[group_of[l]]));
            GlobalMsgArgs a;
            memset(&a, 0, sizeof(a));
            a.n_nodes = (int)N; a.n_edges = (int)Eg; a.ptr = pl.g_ptr; a.src = pl.g_src; a.dst = pl.g_dst; a.P = hw.P;
            a.QT = w.QT + l * 2 * D; a.ldq = L * 2 * D; a.x1 = hw.x1; a.h = hw.h;
            if (fuse_gather) {
                gstage = stage_zero();
                gstage.op = CH_GMSG_FWD; gstage.dst = 0; gstage.g0 = hw.P; gstage.g1 = hw.x1; gstage.W = a.QT; gstage.ldw = a.ldq;
                gstage.i_ptr = pl.g_ptr; gstage.i_src = pl.g_src; gstage.out_a = hw.h;
            } else {
                PAMNET_TRY(global_msg_fwd(D, a, (int)Eg, st));
            }
        } else {
            if (l == grp[group_of[l]]) PAMNET_TRY(sc.wait(st, ev_l[group_of[l]]));
            LocalMsgArgs a;
            memset(&a, 0, sizeof(a));
            a.n_nodes = (int)N; a.n_edges = (int)El; a.ptr = pl.l_ptr; a.src = pl.l_src; a.dst = pl.l_dst;
            a.t_ptr = pl.t_ptr; a.t_gather = pl.t_gather; a.P = hw.P;
            a.QR = w.QR + l * 4 * D; a.ldq = L * 4 * D; a.zq = w.zq2 + l * D; a.ldt = L * D;
            a.x1 = hw.x1; a.m_nb = hw.m_nb; a.msum = hw.msum; a.h = hw.h;
            PAMNET_TRY(local_edge_fwd(D, a, st));
            PAMNET_TRY(local_trip_fwd(D, a, (int)T, st));
            if (fuse_gather) {
                gstage = stage_zero();
                gstage.op = CH_LMSG_FWD; gstage.dst = 0; gstage.g0 = hw.msum; gstage.g1 = hw.x1; gstage.W = a.QR + 3 * D; gstage.ldw = a.ldq;
                gstage.i_ptr = pl.l_ptr; gstage.out_a = hw.h;
            } else {
                PAMNET_TRY(local_msg_fwd(D, a, (int)T, st));
            }
        }
        Prog p((int)N);
        const float* res_x = hh == 0 ? w.x0 : w.half[hh - 1].r[2];
        add_post_fwd(p, params, half_params(mp, hh), hw, D, res_x, fuse_gather ? &gstage : nullptr);
        if (hh + 1 < H) add_pre_fwd(p, params, half_params(mp, hh + 1), w.half[hh + 1], hh + 1, D, 1);
        PAMNET_TRY(chain_launch(D, p.a, st));
        {   // readout heads of this half, concurrently with the next half
            PAMNET_TRY(sc.order(st, sc.s3));
            Prog ph((int)N);
            add_heads_fwd(ph, params, half_params(mp, hh), hw, D, w.att + (size_t)hh * N, w.out + (size_t)hh * N);
            ph.a.small_footprint = heads_small_footprint();
            PAMNET_TRY(chain_launch(D, ph.a, sc.s3));
        }
    }
    if (sc.s3 != st) PAMNET_TRY(sc.order(sc.s3, st));

    // ---- readout (models.py:206-224) ------------------------------------------------------------------
    ReadoutArgs r;
    memset(&r, 0, sizeof(r));
    r.n_nodes = (int)N; r.n_graphs = (int)c.G; r.n_layer = L; r.pool_mean = cfg.dataset == PAMNET_RNA;
    r.sign = cfg.dataset == PAMNET_PDBBIND ? sign : nullptr;
    r.gptr = pl.gptr; r.n2g = pl.n2g; r.att = w.att; r.out = w.out; r.node_val = w.node_val; r.pooled = out;
    PAMNET_TRY(readout_forward(r, st));
    PAMNET_TRY(sc.order(s2, st));      // nothing of this call is still running on the auxiliary streams afterwards
    if (sB != s2) PAMNET_TRY(sc.order(sB, st));
    if (st != st_user) PAMNET_TRY(sc.order(st, st_user));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
int model_backward(const pamnet_config_t& cfg, const pamnet_sizes_t& sz, const pamnet_sbf_consts_t& sbf,
                   const float* params, const float* node_in, const float* sign, const float* pos, void* plan_base,
                   void* plan_trip, void* workspace, size_t ws_bytes, const float* grad_out, float* gp,
                   cudaStream_t st_user, cudaStream_t aux, void* prepared) {
    (void)pos;
    Ctx c;
    PAMNET_TRY(make_ctx(cfg, sz, sbf, plan_base, plan_trip, workspace, ws_bytes, &c, prepared));
    const int D = c.D, L = c.L, H = c.H;
    const int64_t N = c.N, Eg = c.Eg, El = c.El, T = c.T;
    const ModelP& mp = c.mp;
    Ws& w = c.w;
    const Plan& pl = c.plan;
    Sched sc = make_sched(st_user, aux);
    cudaStream_t st = sc.st, s2 = sc.s2, s3 = sc.s3;
    gemm_set_stable_range(params, (size_t)mp.total);

    const bool native_comm = comm_active() && sc.dual;     // gradient all-reduce issued from here, bucket by bucket
    if (st != st_user) PAMNET_TRY(sc.order(st_user, st));
    PAMNET_CUDA(cudaMemsetAsync(gp, 0, sizeof(float) * mp.total, st));
    PAMNET_TRY(sc.order(st, s2));      // gradients zeroed (split-K GEMMs accumulate into them) before s2 / s3 start
    if (s3 != s2) PAMNET_TRY(sc.order(st, s3));
    // accumulators of the per-edge / per-triplet embedding gradients, summed over layers with fp32 atomics
    if (sc.s4 != s2) PAMNET_TRY(sc.order(st, sc.s4));
    if (sc.s5 != s2 && sc.s5 != sc.s4) PAMNET_TRY(sc.order(st, sc.s5));
    // grad of the per-node projections: deterministic CSR gather pass (default), or PAMNET_GATHER=atomic lets the message
    // kernels accumulate it themselves with fp32 reductions in L2 (12 launches fewer; measured no faster: the
    // reductions of one node's ~18 edges serialise in L2 for about as long as the gather kernel takes)
    static int atomic_gp = -1;
    if (atomic_gp < 0) { const char* e = getenv("PAMNET_GATHER"); atomic_gp = (e && strcmp(e, "atomic") == 0) ? 1 : 0; }
    if (atomic_gp) {
        for (int hh = 0; hh < H; ++hh)
            PAMNET_CUDA(cudaMemsetAsync(w.half[hh].g_P, 0, sizeof(float) * N * nP_of(hh) * D, s2));
        PAMNET_TRY(sc.order(s2, st));
    }
    PAMNET_CUDA(cudaMemsetAsync(w.gz_eg, 0, sizeof(float) * Eg * D, s2));
    PAMNET_CUDA(cudaMemsetAsync(w.gz_el, 0, sizeof(float) * El * D, sc.s4));
    PAMNET_CUDA(cudaMemsetAsync(w.gz_s, 0, sizeof(float) * T * D, sc.s5));

    ReadoutArgs r;
    memset(&r, 0, sizeof(r));
    r.n_nodes = (int)N; r.n_graphs = (int)c.G; r.n_layer = L; r.pool_mean = cfg.dataset == PAMNET_RNA;
    r.sign = cfg.dataset == PAMNET_PDBBIND ? sign : nullptr;
    r.gptr = pl.gptr; r.n2g = pl.n2g; r.att = w.att; r.out = w.out; r.g_pooled = grad_out; r.g_att = w.g_att; r.g_out = w.g_out;
    PAMNET_TRY(readout_backward(r, st));

    // ---- readout heads of every half, in the order the main stream will need them --------------------------------
    std::vector<cudaEvent_t> ev_heads(H);
    PAMNET_TRY(sc.order(st, s3));
    for (int hh = H - 1; hh >= 0; --hh) {
        Prog ph((int)N);
        add_heads_bwd(ph, params, half_params(mp, hh), w.half[hh], D, w.g_att + (size_t)hh * N, w.g_out + (size_t)hh * N, gp);
        ph.a.small_footprint = heads_small_footprint();
        PAMNET_TRY(chain_launch(D, ph.a, s3));
        PAMNET_TRY(sc.record(s3, &ev_heads[hh]));
    }

    const int ks_n = pick_ksplit(N);
    // ONE launch per half for every weight gradient that is a [D, D] (or [1, D]) reduction over the nodes:
    //  * the post-chain of half `hh` (hh >= 0) and mlp_x1 of half `x1_half` (x1_half >= 0), whose grad_z a finished
    //    chain kernel has written;
    //  (the two heads' own gradients dW = o3^T g_att, dW_out = o3^T g_out are produced by the heads chain: add_heads_bwd)
    //  * the per-node halves of the edge MLPs of half `proj_half`: dW[:, cD:(c+1)D] = g_P_c^T x1.
    auto node_wgrads = [&](int hh, int x1_half, int proj_half, cudaStream_t s) -> int {
        std::vector<GemmSlot> sl;
        auto lin = [&](const float* gz, const float* a_in, const Lin& p) {
            sl.push_back(slot(gz, D, a_in, D, gp + p.w, D, nullptr, gp + p.b));
        };
        if (x1_half >= 0) {
            const float* x_in = x1_half == 0 ? w.x0 : w.half[x1_half - 1].r[2];
            lin(w.half[x1_half].gz_x1, x_in, half_params(mp, x1_half).x1);
        }
        if (hh >= 0) {
            const HalfWs& hw = w.half[hh];
            const HalfP& hp = half_params(mp, hh);
            lin(hw.gz_x2, hw.h, hp.x2);
            for (int rr = 0; rr < 3; ++rr) {
                lin(hw.gz_A[rr], rr == 0 ? hw.a_x2 : hw.r[rr - 1], hp.res[rr][0]);
                lin(hw.gz_B[rr], hw.a_A[rr], hp.res[rr][1]);
            }
            lin(hw.gz_o[0], hw.r[2], hp.out[0]);
            lin(hw.gz_o[1], hw.a_o[0], hp.out[1]);
            lin(hw.gz_o[2], hw.a_o[1], hp.out[2]);
        }
        if (proj_half >= 0) {
            const HalfWs& hw = w.half[proj_half];
            const HalfP& hp = half_params(mp, proj_half);
            const int nP = nP_of(proj_half);
            for (int cblk = 0; cblk < nP; ++cblk) {
                int64_t woff;
                if (!is_local(proj_half)) woff = hp.m.w + cblk * D;
                else woff = (cblk < 2 ? hp.m_ji.w : hp.m_kj.w) + (cblk & 1) * D;
                sl.push_back(slot(hw.g_P + cblk * D, nP * D, hw.x1, D, gp + woff, 3 * D));
            }
        }
        GemmArgs a = gemm_zero(GEMM_TN, EPI_NONE, D, D, (int)N);
        a.ksplit = ks_n;
        return gemm_multi(a, sl, s);
    };
    // phase A' of layers [l0, l1): per-edge / per-triplet gradients -> weight gradients (batched over the layers of
    // the group), and their contribution to the gradient of the layer-invariant embeddings, accumulated over
    // groups in gz_eg / gz_el / gz_s with fp32 atomics (split-K)
    // The three families are independent of each other (they only meet in the atomically accumulated gz_* buffers),
    // and each launch is at most one wave of ~10 us tiles: on separate streams they pack the SMs side by side.
    auto global_wgrads = [&](int l0, int l1, cudaStream_t s) -> int {
        const int nl = l1 - l0;
        {   // global edges
            const int ldq = L * 2 * D;
            std::vector<GemmSlot> sl;
            GemmArgs b = gemm_zero(GEMM_NN, EPI_NONE, (int)Eg, D, nl * 2 * D);
            b.nslots = 1; b.nseg = 2 * nl; b.seg_len = D; b.ksplit = 2 * nl;
            for (int l = l0; l < l1; ++l) {
                const HalfP& hp = mp.g[l];
                const float* g = w.gQT + l * 2 * D;
                sl.push_back(slot(g, ldq, w.e_g, D, gp + hp.m.w + 2 * D, 3 * D, nullptr, gp + hp.m.b));
                sl.push_back(slot(g + D, ldq, w.e_g, D, gp + hp.We.w, D));
                b.seg_B[2 * (l - l0)] = params + hp.m.w + 2 * D; b.seg_ldb[2 * (l - l0)] = 3 * D;
                b.seg_B[2 * (l - l0) + 1] = params + hp.We.w; b.seg_ldb[2 * (l - l0) + 1] = D;
            }
            GemmArgs a = gemm_zero(GEMM_TN, EPI_NONE, D, D, (int)Eg);
            a.ksplit = pick_ksplit(Eg);
            PAMNET_TRY(gemm_multi(a, sl, s));
            b.slot[0] = slot(w.gQT + l0 * 2 * D, ldq, nullptr, 0, w.gz_eg, D);   // gz_eg += [gQ | gTt] [W_m,e ; W_e]
            PAMNET_TRY(gemm_launch(b, s));
        }
        return 0;
    };
    auto local_wgrads = [&](int l0, int l1, cudaStream_t s, cudaStream_t sdg) -> int {
        const int nl = l1 - l0;
        {   // local edges
            const int ldq = L * 4 * D;
            std::vector<GemmSlot> sl;
            GemmArgs b = gemm_zero(GEMM_NN, EPI_NONE, (int)El, D, nl * 4 * D);
            b.nslots = 1; b.nseg = 4 * nl; b.seg_len = D; b.ksplit = 4 * nl;
            for (int l = l0; l < l1; ++l) {
                const HalfP& hp = mp.l[l];
                const float* g = w.gQR + l * 4 * D;
                sl.push_back(slot(g, ldq, w.e_l, D, gp + hp.m_ji.w + 2 * D, 3 * D, nullptr, gp + hp.m_ji.b));
                sl.push_back(slot(g + D, ldq, w.e_l, D, gp + hp.m_kj.w + 2 * D, 3 * D, nullptr, gp + hp.m_kj.b));
                sl.push_back(slot(g + 2 * D, ldq, w.e_l, D, gp + hp.lin_rbf.w, D));
                sl.push_back(slot(g + 3 * D, ldq, w.e_l, D, gp + hp.lin_rbf_out.w, D));
                const int o = 4 * (l - l0);
                b.seg_B[o] = params + hp.m_ji.w + 2 * D; b.seg_ldb[o] = 3 * D;
                b.seg_B[o + 1] = params + hp.m_kj.w + 2 * D; b.seg_ldb[o + 1] = 3 * D;
                b.seg_B[o + 2] = params + hp.lin_rbf.w; b.seg_ldb[o + 2] = D;
                b.seg_B[o + 3] = params + hp.lin_rbf_out.w; b.seg_ldb[o + 3] = D;
            }
            GemmArgs a = gemm_zero(GEMM_TN, EPI_NONE, D, D, (int)El);
            a.ksplit = pick_ksplit(El);
            PAMNET_TRY(gemm_multi(a, sl, s));
            b.slot[0] = slot(w.gQR + l0 * 4 * D, ldq, nullptr, 0, w.gz_el, D);
            PAMNET_TRY(gemm_launch(b, s));
        }
        {   // triplet gate MLP (local_message_passing.py:17,49)
            const int ldt = L * D;
            std::vector<GemmSlot> w2, dg, w1;
            GemmArgs bs = gemm_zero(GEMM_NN, EPI_NONE, (int)T, D, nl * D);       // gz_s += gzq1 W_sbf0
            bs.nslots = 1; bs.nseg = nl; bs.seg_len = D; bs.ksplit = nl > 1 ? nl : 2;
            for (int l = l0; l < l1; ++l) {
                const HalfP& hp = mp.l[l];
                w2.push_back(slot(w.gzq2 + l * D, ldt, w.aq1 + l * D, ldt, gp + hp.sbf[1].w, D, nullptr, gp + hp.sbf[1].b));
                dg.push_back(slot(w.gzq2 + l * D, ldt, params + hp.sbf[1].w, D, w.gzq1 + l * D, ldt, nullptr, nullptr,
                                  w.zq1 + l * D, ldt));
                w1.push_back(slot(w.gzq1 + l * D, ldt, w.s, D, gp + hp.sbf[0].w, D, nullptr, gp + hp.sbf[0].b));
                bs.seg_B[l - l0] = params + hp.sbf[0].w; bs.seg_ldb[l - l0] = D;
            }
            GemmArgs t = gemm_zero(GEMM_TN, EPI_NONE, D, D, (int)T);
            t.ksplit = pick_ksplit(T);
            PAMNET_TRY(gemm_multi(t, w2, s));
            PAMNET_TRY(gemm_multi(gemm_zero(GEMM_NN, EPI_MUL_DSILU, (int)T, D, D), dg, sdg));
            PAMNET_TRY(gemm_multi(t, w1, sdg));
            bs.slot[0] = slot(w.gzq1 + l0 * D, ldt, nullptr, 0, w.gz_s, D);
            PAMNET_TRY(gemm_launch(bs, sdg));
        }
        return 0;
    };
    // ---- through the SiLU of the layer-invariant embeddings into their weights and the RBF frequencies.  Each family
    // finishes on the stream that accumulated its gz_* buffer, as soon as its last layer is in (only the global one
    // is left for after the main loop).
    // (this one is the tail of the whole step: its weight-gradient GEMM runs beside the data-gradient -> frequency
    // chain on the small-GEMM stream instead of in front of it)
    auto embed_tail_global = [&](cudaStream_t s) -> int {
        PAMNET_TRY(mul_dsilu_launch(w.gz_eg, w.z_eg, Eg * D, s));
        GemmArgs cw = gemm_zero(GEMM_TN, EPI_NONE, D, kNumRbf, (int)Eg);
        cw.nslots = 1; cw.ksplit = pick_ksplit(Eg);
        cw.slot[0] = slot(w.gz_eg, D, w.rbf_g, kNumRbf, gp + mp.rbf_g.w, kNumRbf, nullptr, gp + mp.rbf_g.b);
        if (s3 != s) PAMNET_TRY(sc.order(s, s3));
        PAMNET_TRY(gemm_launch(cw, s3 != s ? s3 : s));
        GemmArgs d = gemm_zero(GEMM_NN, EPI_NONE, (int)Eg, kNumRbf, D);
        d.nslots = 1;
        d.slot[0] = slot(w.gz_eg, D, params + mp.rbf_g.w, kNumRbf, w.g_rbf_g, kNumRbf);
        PAMNET_TRY(gemm_launch(d, s));
        return rbf_freq_backward(pl.dist_g, Eg, params + mp.freq_g, cfg.cutoff_g, w.g_rbf_g, gp + mp.freq_g, s);
    };
    auto embed_tail_local = [&](cudaStream_t s) -> int {
        PAMNET_TRY(mul_dsilu_launch(w.gz_el, w.z_el, El * D, s));
        GemmArgs cw = gemm_zero(GEMM_TN, EPI_NONE, D, kNumRbf, (int)El);
        cw.nslots = 1; cw.ksplit = pick_ksplit(El);
        cw.slot[0] = slot(w.gz_el, D, w.rbf_l, kNumRbf, gp + mp.rbf_l.w, kNumRbf, nullptr, gp + mp.rbf_l.b);
        PAMNET_TRY(gemm_launch(cw, s));
        GemmArgs d = gemm_zero(GEMM_NN, EPI_NONE, (int)El, kNumRbf, D);
        d.nslots = 1;
        d.slot[0] = slot(w.gz_el, D, params + mp.rbf_l.w, kNumRbf, w.g_rbf_l, kNumRbf);
        PAMNET_TRY(gemm_launch(d, s));
        return rbf_freq_backward(pl.dist_l, El, params + mp.freq_l, cfg.cutoff_l, w.g_rbf_l, gp + mp.freq_l, s);
    };
    auto embed_tail_sbf = [&](cudaStream_t s) -> int {      // SBF embeddings (models.py:187-188)
        PAMNET_TRY(mul_dsilu_launch(w.gz_s, w.z_s, T * D, s));
        if (sbf_fused_enabled(D))
            return sbf_embed_wgrad(pl, T, w.radial, w.ysph, w.gz_s, D, cfg.simple ? nullptr : gp + mp.sbf2.w,
                                   cfg.simple ? nullptr : gp + mp.sbf2.b, gp + mp.sbf1.w, gp + mp.sbf1.b, s);
        PAMNET_CUDA(cudaMemsetAsync(w.gw_ext, 0, sizeof(float) * D * kSbfExt, s));
        GemmArgs cw = gemm_zero(GEMM_TN, EPI_NONE, D, kSbfExt, (int)T);
        cw.nslots = 1; cw.ksplit = pick_ksplit(T);
        cw.slot[0] = slot(w.gz_s, D, w.sbf_ext, kSbfExt, w.gw_ext, kSbfExt);
        PAMNET_TRY(gemm_launch(cw, s));
        return sbf_weight_unpack_grad(D, w.gw_ext, cfg.simple ? nullptr : gp + mp.sbf2.w,
                                      cfg.simple ? nullptr : gp + mp.sbf2.b, gp + mp.sbf1.w, gp + mp.sbf1.b, s);
    };

    // ---- main stream: phase B reversed; auxiliary stream: weight gradients as soon as their inputs exist --------
    const bool fuse_gather = chain_mma_enabled(D) && gather_fusion_enabled() && !atomic_gp;
    ChainStage gstage = stage_zero();
    for (int hh = H - 1; hh >= 0; --hh) {
        const HalfWs& hw = w.half[hh];
        const HalfP& hp = half_params(mp, hh);
        const int l = hh >> 1;
        {
            Prog p((int)N);
            const bool has_gx = hh + 1 < H;   // the last layer's x output is unused (models.py:201-204)
            if (has_gx) add_pre_bwd(p, params, half_params(mp, hh + 1), w.half[hh + 1], hh + 1, D, w, nullptr,
                                    fuse_gather ? &gstage : nullptr);      // gstage: half hh + 1's gather, set up below last iteration
            add_post_bwd(p, params, hp, hw, D, w, has_gx);
            PAMNET_TRY(sc.wait(st, ev_heads[hh]));
            PAMNET_TRY(chain_launch(D, p.a, st));
        }
        NodeGatherArgs ng;
        memset(&ng, 0, sizeof(ng));
        ng.n_nodes = (int)N; ng.g_P = hw.g_P;
        if (!is_local(hh)) {
            GlobalMsgArgs a;
            memset(&a, 0, sizeof(a));
            a.n_nodes = (int)N; a.n_edges = (int)Eg; a.ptr = pl.g_ptr; a.src = pl.g_src; a.dst = pl.g_dst; a.P = hw.P;
            a.QT = w.QT + l * 2 * D; a.ldq = L * 2 * D; a.g_h = w.g_h; a.gQT = w.gQT + l * 2 * D;
            a.g_P = atomic_gp ? hw.g_P : nullptr;
            PAMNET_TRY(global_msg_bwd(D, a, (int)Eg, st));
            ng.n_blocks = 1; ng.ptr = pl.g_ptr; ng.optr = pl.g_optr; ng.opos = pl.g_opos;
            ng.gz = w.gQT + l * 2 * D; ng.ldq = L * 2 * D;
        } else {
            LocalMsgArgs a;
            memset(&a, 0, sizeof(a));
            a.n_nodes = (int)N; a.n_edges = (int)El; a.ptr = pl.l_ptr; a.src = pl.l_src; a.dst = pl.l_dst;
            a.t_ptr = pl.t_ptr; a.t_gather = pl.t_gather; a.tt_ptr = pl.tt_ptr; a.tt_t = pl.tt_t; a.t_owner = pl.t_owner;
            a.P = hw.P; a.QR = w.QR + l * 4 * D; a.ldq = L * 4 * D; a.zq = w.zq2 + l * D; a.ldt = L * D;
            a.m_nb = hw.m_nb; a.msum = hw.msum; a.g_h = w.g_h; a.g_s = w.g_s;
            a.gQR = w.gQR + l * 4 * D; a.gzq = w.gzq2 + l * D;
            a.g_P = atomic_gp ? hw.g_P : nullptr;
            PAMNET_TRY(local_msg_bwd(D, a, (int)T, st));
            PAMNET_TRY(local_trip_bwd(D, a, (int)T, st));
            ng.n_blocks = 2; ng.ptr = pl.l_ptr; ng.optr = pl.l_optr; ng.opos = pl.l_opos;
            ng.gz = w.gQR + l * 4 * D; ng.ldq = L * 4 * D;
        }
        if (fuse_gather) {
            // the gather of this half's per-edge gradients into grad P runs as the loading stage of the NEXT chain launch
            // (which consumes it); the projection weight gradients that read grad P follow that launch
            gstage = stage_zero();
            gstage.op = CH_GATHER_BWD; gstage.dst = kChainWide; gstage.width = 2 * ng.n_blocks * D;
            gstage.W = ng.gz; gstage.ldw = ng.ldq; gstage.i_ptr = ng.ptr; gstage.o_ptr = ng.optr; gstage.o_pos = ng.opos;
            gstage.out_a = hw.g_P;
        } else if (!atomic_gp) {
            PAMNET_TRY(node_grad_gather(D, ng, is_local(hh) ? (int)El : (int)Eg, st));
        }
        PAMNET_TRY(sc.order(st, s3));
        if (fuse_gather) PAMNET_TRY(node_wgrads(hh, hh + 1 < H ? hh + 1 : -1, hh + 1 < H ? hh + 1 : -1, s3));
        else PAMNET_TRY(node_wgrads(hh, hh + 1 < H ? hh + 1 : -1, hh, s3));
        // per-edge / per-triplet weight gradients of this half: everything they read is complete now
        if (is_local(hh)) {
            if (sc.s4 != s3) PAMNET_TRY(sc.order(st, sc.s4));
            if (sc.s5 != sc.s4) PAMNET_TRY(sc.order(st, sc.s5));
            // two layers per batched launch (the per-edge gradient buffers hold all layers side by side): half the launches,
            // twice the tiles per launch; layer l + 1 simply waits for layer l.  Per layer when a caller consumes the
            // per-half bucket events (their contract is "final after the next half's iteration").
            if (g_buckets_on.load()) PAMNET_TRY(local_wgrads(l, l + 1, sc.s4, sc.s5));
            else if (l % 2 == 0) PAMNET_TRY(local_wgrads(l, l + 2 < L ? l + 2 : L, sc.s4, sc.s5));
            if (l == 0) {                        // last local half: both local families are complete
                PAMNET_TRY(embed_tail_local(sc.s4));
                PAMNET_TRY(embed_tail_sbf(sc.s5));
            }
        } else {
            if (s3 != s2) PAMNET_TRY(sc.order(st, s2));
            if (g_buckets_on.load()) PAMNET_TRY(global_wgrads(l, l + 1, s2));
            else if (l % 2 == 0) PAMNET_TRY(global_wgrads(l, l + 2 < L ? l + 2 : L, s2));
        }
        if (g_buckets_on.load() && hh + 1 < H) {
            const cudaStream_t bs[kBucketStreams] = {s2, s3, sc.s4, sc.s5};
            PAMNET_TRY(record_bucket(hh + 1, H, bs));
        }
        // library-owned collective: layers [l, l + 1] of the kind that just completed (half hh + 1), once both are final
        if (native_comm && hh + 1 < H) {
            const int hc = hh + 1, lc = hc >> 1;
            const bool pair_done = (lc % 2 == 0);      // buckets of two layers [lc, lc + 1]: lc + 1 completed earlier
            if (pair_done) {
                const int l_hi = (lc + 1 < L) ? lc + 1 : lc;
                const HalfP* lo_p = (hc & 1) ? &mp.l[lc] : &mp.g[lc];
                int64_t lo = lo_p->W, hi;
                if (hc & 1) hi = l_hi + 1 < L ? mp.l[l_hi + 1].W : mp.total;
                else hi = l_hi + 1 < L ? mp.g[l_hi + 1].W : mp.l[0].W;
                cudaStream_t cs = comm_stream();
                for (cudaStream_t s : {s2, s3, sc.s4, sc.s5}) PAMNET_TRY(sc.order(s, cs));
                PAMNET_TRY(comm_allreduce_avg(gp + lo, hi - lo));
            }
        }
    }
    {   // into the node input
        Prog p((int)N);
        add_pre_bwd(p, params, half_params(mp, 0), w.half[0], 0, D, w, w.g_x0, fuse_gather ? &gstage : nullptr);
        PAMNET_TRY(chain_launch(D, p.a, st));
    }
    if (cfg.dataset == PAMNET_PDBBIND) {
        GemmArgs a = gemm_zero(GEMM_TN, EPI_NONE, D, kFeatPdb, (int)N);
        a.nslots = 1; a.ksplit = ks_n;
        a.slot[0] = slot(w.g_x0, D, node_in, kFeatPdb, gp + mp.init_linear, kFeatPdb);
        PAMNET_TRY(gemm_launch(a, st));
    } else {
        PAMNET_TRY(embed_backward(node_in, N, w.g_x0, mp.n_embed, D, gp + mp.emb, st));
    }
    PAMNET_TRY(sc.order(st, s3));
    PAMNET_TRY(node_wgrads(-1, 0, fuse_gather ? 0 : -1, s3));
    if (g_buckets_on.load()) {
        const cudaStream_t bs[kBucketStreams] = {s2, s3, sc.s4, sc.s5};
        PAMNET_TRY(record_bucket(0, H, bs));
    }
    if (s3 != s2) PAMNET_TRY(sc.order(st, s2));

    // ---- auxiliary stream: through the SiLU of the embeddings into their weights and the RBF frequencies -------
    PAMNET_TRY(embed_tail_global(s2));
    PAMNET_TRY(sc.order(s2, st));
    if (s3 != s2) PAMNET_TRY(sc.order(s3, st));
    if (sc.s4 != s2) PAMNET_TRY(sc.order(sc.s4, st));
    if (sc.s5 != s2 && sc.s5 != sc.s4) PAMNET_TRY(sc.order(sc.s5, st));
    if (native_comm) {
        // head of the buffer (embeddings, frequencies, basis MLPs: final with everything else) together with global
        // layers [0, 1], which complete with the last node-level weight gradients and are contiguous with it: ONE
        // exposed collective instead of two
        const int l_hi = 1 < L ? 1 : 0;
        const int64_t hi = l_hi + 1 < L ? mp.g[l_hi + 1].W : mp.l[0].W;
        cudaStream_t cs = comm_stream();
        PAMNET_TRY(sc.order(st, cs));
        PAMNET_TRY(comm_allreduce_avg(gp, hi));
        PAMNET_TRY(sc.order(cs, st));
    }
    if (st != st_user) PAMNET_TRY(sc.order(st, st_user));
    return 0;
}

}  // namespace pamnet
