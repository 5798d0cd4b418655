// fdpt_api.cu — context, parameters, workspace, forward orchestration, sampling loop and the C ABI of libfdpt.so.
// Reference interfaces replaced by each entry point are listed in include/fdpt.h.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <tuple>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/fdpt.h"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "kernels_diffusion.cuh"
#include "kernels_ipa.cuh"
#include "kernels_misc.cuh"
#include "et_fused.cuh"
#include "edge_embed_fused.cuh"
#include "tc_linear.cuh"
#include "gemm_tc.cuh"
#include "lin_tc.cuh"
#include "gemm_img.cuh"
#include "lin_tcw.cuh"
#include "tmem_a_test.cuh"
#include "backbone_tables.inc"

using namespace fdpt;

namespace {

struct ParamSpec {
  std::vector<int64_t> shape;
  float* dev = nullptr;
};

struct BlockParams {
  const float *head_w, *Wq, *bq, *Wkv, *bkv, *Wqp, *bqp, *Wkvp, *bkvp, *Wb, *bb, *Wd, *bd, *Wout, *bout;
  const float *ln_g, *ln_b, *Wskip, *bskip;
  struct TfLayer {
    const float *Win, *bin, *Wo, *bo, *W1, *b1, *W2, *b2, *n1g, *n1b, *n2g, *n2b;
    float *Win_s = nullptr, *bin_s = nullptr;  // in_proj with the q rows pre-multiplied by 1/sqrt(d_head) (operand-image attention path)
  } tf[TF_LAYERS];
  const float *Wpost, *bpost;
  const float *Wt1, *bt1, *Wt2, *bt2, *Wt3, *bt3, *tln_g, *tln_b;
  const float *Wbb, *bbb;
  // edge transition (blocks 0..2)
  const float *Wie, *bie, *We1, *be1, *We2, *be2, *Wef, *bef, *eln_g, *eln_b;
  // fp16 128B-swizzled weight images for the fused tcgen05 EdgeTransition kernel (et_fused.cuh)
  __half *imgW1cat = nullptr, *imgW2 = nullptr, *imgW3cat = nullptr;
  // IPA (kernels_ipa.cuh): fused projection weight [6816,256] / bias with per-head contiguous rows, linear_out.weight with columns
  // in cat' order, fp16 hi|lo operand image of linear_b.weight
  float *Wcat = nullptr, *bcat = nullptr, *Wout_perm = nullptr;
  // projection in operand-image column order (lin_tc.cuh IpaProjEpi): 24 blocks of 320 rows = [operand q|k|v][head][256 scalar | points | 0],
  // q's scalar rows pre-multiplied by sqrt(1/(3 C))
  float *Wimgproj = nullptr, *bimgproj = nullptr;
  __half* imgWout = nullptr;  // linear_out.weight (cat' column order) as split operand image [2 n-tiles][42 k-blocks][hi|lo][128][128 B]
  __half* imgWb = nullptr;
};

struct Workspace {
  char* base = nullptr;
  size_t bytes = 0;
  int capB = 0, capN = 0;
  // node side
  float *node_feat, *feat1d, *node0, *node, *tmpA, *tmpB, *tmpC;  // tmp: [M,320]-capable
  __half *imgF = nullptr, *imgT1 = nullptr, *imgT2 = nullptr, *imgN = nullptr;  // operand images passed between consecutive Linear layers (lin_tc Ximg / Yimg)
  size_t img_bytes = 0;
  float *PA;  // edge embedder layer-1 per-residue partial A f_i + b0
  float *proj, *kn, *cat;  // IPA: fused projections [M,6816], -gamma/2 |k_pts|^2 [M,8], concat [M,2688]
  int ldS;
  float *tf_x, *qkv, *att_o;
  float *upd, *quats, *trans, *dmask;
  float *n_emb, *U, *V, *Pf, *Qf;
  float *tors_u;
  // pair side
  float* S;  // [B,H,N,ldS] attention logits / probabilities (IPA and sequence transformer)
  uint8_t *qimg = nullptr, *kimg = nullptr, *vimg = nullptr, *pimg = nullptr;  // IPA operand images (gemm_img.cuh): Q', K', V' [B*H][JB][5][32 KB]; P [B*H][JB][2JB][32 KB]
  size_t qkv_img_bytes = 0, p_img_bytes = 0;
  uint8_t* cat_img = nullptr;  // concat row of the IPA as operand image [m-tile][42 k-blocks][32 KB] (A operand of linear_out on gemm_img)
  size_t cat_img_bytes = 0;
  uint8_t *tf_qimg = nullptr, *tf_kimg = nullptr, *tf_vimg = nullptr, *tf_pimg = nullptr;  // same for the sequence transformer: [B*4][JB][2][32 KB], P [B*4][JB][2JB][32 KB]
  size_t tf_qkv_img_bytes = 0, tf_p_img_bytes = 0;
  __half *z, *n_img;               // z: fp16 tile images [B][N][JB][32 KB]; n_img: [B][JB][32 KB]
  __half *f_img, *rel_tab;         // edge embedder: f_j k-block images [B][JB][16 KB]; fp16 relative-offset embedding table
  int JB;
  // outputs / sampling state
  float *pred_rigids, *trans_score, *psi, *rig_cur, *rig_next, *sc_ca, *t_emb_b, *t32_b, *bb_tmp;
  double *rot_score, *sigma_b, *sched_dev;
  int* sigma_idx_b;  // [B] row of the cached score table at the current step
  int* step_dev;  // [0] device step counter of the sampling loop, [1] number of steps T (the captured step graph reads both)
  void** call_ptrs;  // per-call buffer bases read by the captured step: [0] noise, [1] prot_traj, [2] rigid_0_traj, [3] trans_traj, [4] rigid_traj
  float* temb_tab;   // [4096, 32] this call's timestep-embedding table
};

}  // namespace

struct fdpt_ctx {
  fdpt_config cfg;
  int device = 0;
  std::string err;
  std::map<std::string, ParamSpec> params;
  bool finalized = false;
  BlockParams blk[NBLK];
  struct {
    const float *nW0, *nb0, *nW2, *nb2, *nW4, *nb4, *nln_g, *nln_b;
    const float *eW0, *eb0, *eW2, *eb2, *eW4, *eb4, *eln_g, *eln_b;
    const float *tW1, *tb1, *tW2, *tb2, *tWf, *tbf;
    __half *imgE0 = nullptr, *imgE2 = nullptr, *imgE4 = nullptr;  // edge embedder weights as fp16 operand images (edge_embed_fused.cuh)
  } top;
  float *bin_lower = nullptr, *ideal = nullptr, *psi_frame = nullptr, *atom_mask = nullptr;
  Workspace ws;
  int64_t launches = 0;
  int max_smem_optin = 0, max_smem_sm = 0, num_sms = 148;
  int gemm_tc = 1;   // node-side GEMMs on tcgen05 (3-term split TF32); 0 = SIMT fp32 kernel (bring-up / A-B switch)
  int mn_swap = 0;   // bring-up knob of the MN-major descriptor
  int dbg_flags = 0; // bring-up knob of lin_tc: bit 0 skip the epilogue stores, bit 1 skip the MMAs
  // Linear weights pre-split into fp16 hi|lo operand images (lin_tc.cuh), keyed by (weight pointer, row stride, N, K)
  struct PackedW { __half* img; int nkb, n_tiles; };
  std::map<std::tuple<const float*, int, int, int>, PackedW> packed;
  // CUDA graph of one (non-final) timestep of fdpt_sample, replayed with a device-side step counter; rebuilt when any baked
  // pointer or size changes
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    std::vector<unsigned char> key;
    int64_t launches = 0;
  } step_graph;
  int use_graph = 1;
  int64_t stat_captures = 0;      // per-timestep graphs captured so far
  int64_t stat_sample_host_us = 0; // host time the last fdpt_sample call spent enqueueing
  int lin_wres = 1;   // Linear layers whose CTAs own one n-tile keep the whole weight panel resident and stream the activation (lin_tcw.cuh); 0 = lin_tc
  int tf_img = 1;     // sequence-transformer attention GEMMs from operand images (in_proj epilogue -> gemm_img); 0 = gemm_tc path (A/B switch)
  int ipa_img = 1;    // IPA attention GEMMs from operand images (gemm_img.cuh); 0 = fp32 operands split on the fly (gemm_tc.cuh; A/B switch)
  cudaStream_t own_stream = nullptr;   // the legacy default stream cannot be captured: fdpt_sample then runs on this stream,
  cudaEvent_t fence_in = nullptr, fence_out = nullptr;  // fenced against the caller's stream with these events
  long long* et_dbg = nullptr;  // optional clock64 timeline buffer of the EdgeTransition kernel (FDPT_OPT_ET_TIMELINE)
  // so3.use_cached_score=True: [num_sigma, num_omega] score-norm table + omega boundaries (fdpt_set_score_table); null = series
  double *score_table = nullptr, *omega_bounds = nullptr;
  int tab_sigma = 0, tab_omega = 0;
  // streaming read-back: an event after every progress_chunk timesteps of the most recent fdpt_sample call
  int progress_chunk = 0, progress_steps = 0;
  std::vector<cudaEvent_t> progress_ev;   // pool, grown on demand
  int progress_used = 0;
  // live profiling (event pairs per slot)
  bool prof_on = false;
  struct ProfRec { cudaEvent_t a, b; int slot; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
};

namespace {

int fail(fdpt_ctx* c, int code, const char* fmt, ...);

// Every C-ABI entry point runs on the context's device and leaves the caller's current device (torch.cuda.current_device()) as it found it.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define GUARD(ctx)                                                                                          \
  DeviceGuard dg__((ctx)->device);                                                                          \
  if (!dg__.ok) return fail((ctx), FDPT_ERR_CUDA, "cudaSetDevice(%d) failed", (ctx)->device)

int fail(fdpt_ctx* c, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

cudaError_t gemm_dispatch(fdpt_ctx* c, const GemmArgs& g, bool b_kmajor, int batch, cudaStream_t st) {
  if (c->gemm_tc) return tc::launch_gemm_tc(g, b_kmajor, batch, st, c->num_sms);
  return launch_gemm(g, b_kmajor, batch, st);
}

#define CK(call)                                                                                                   \
  do {                                                                                                             \
    cudaError_t e__ = (call);                                                                                      \
    if (e__ != cudaSuccess) return fail(ctx, FDPT_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)

#define LAUNCH_CHECK()   \
  do {                   \
    ctx->launches++;     \
    CK(cudaGetLastError()); \
  } while (0)

cudaEvent_t prof_event(fdpt_ctx* c) {
  if (!c->ev_pool.empty()) {
    cudaEvent_t e = c->ev_pool.back();
    c->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
// RAII scope: records an event pair around everything enqueued on `st` during its lifetime
struct ProfScope {
  fdpt_ctx* c; cudaStream_t st; int idx = -1;
  ProfScope(fdpt_ctx* c_, int slot, cudaStream_t st_) : c(c_), st(st_) {
    if (!c->prof_on) return;
    fdpt_ctx::ProfRec r{prof_event(c), prof_event(c), slot};
    cudaEventRecord(r.a, st);
    c->prof.push_back(r);
    idx = (int)c->prof.size() - 1;
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(c->prof[idx].b, st);
  }
};

int f1_dim(const fdpt_ctx* c) { return c->cfg.with_aatype ? 54 : 33; }
// edge embedder input width: [f_i | f_j | relpos 32 | distogram 22 (only with embed_self_conditioning, score_network.py:95-96)]
int ein_dim(const fdpt_ctx* c) { return 2 * f1_dim(c) + EMB + (c->cfg.embed_self_conditioning ? NBINS : 0); }

std::vector<std::pair<std::string, std::vector<int64_t>>> expected_params(const fdpt_ctx* c) {
  std::vector<std::pair<std::string, std::vector<int64_t>>> v;
  auto lin = [&](const std::string& n, int64_t o, int64_t i) {
    v.push_back({n + ".weight", {o, i}});
    v.push_back({n + ".bias", {o}});
  };
  auto ln = [&](const std::string& n, int64_t cdim) {
    v.push_back({n + ".weight", {cdim}});
    v.push_back({n + ".bias", {cdim}});
  };
  const int f1 = f1_dim(c);
  const std::string ne = "embedding_layer.node_embedder", ee = "embedding_layer.edge_embedder";
  lin(ne + ".0", C_S, f1 + EMB); lin(ne + ".2", C_S, C_S); lin(ne + ".4", C_S, C_S); ln(ne + ".5", C_S);
  lin(ee + ".0", C_Z, ein_dim(c)); lin(ee + ".2", C_Z, C_Z); lin(ee + ".4", C_Z, C_Z); ln(ee + ".5", C_Z);
  const std::string t = "score_model.trunk.";
  for (int b = 0; b < NBLK; ++b) {
    const std::string bs = std::to_string(b), p = t + "ipa_" + bs;
    v.push_back({p + ".head_weights", {NH}});
    lin(p + ".linear_q", NH * C_HID, C_S); lin(p + ".linear_kv", 2 * NH * C_HID, C_S);
    lin(p + ".linear_q_points", NH * PQ * 3, C_S); lin(p + ".linear_kv_points", NH * (PQ + PV) * 3, C_S);
    lin(p + ".linear_b", NH, C_Z); lin(p + ".down_z", C_Z / 4, C_Z); lin(p + ".linear_out", C_S, CAT);
    ln(t + "ipa_ln_" + bs, C_S);
    lin(t + "skip_embed_" + bs, C_SKIP, C_S);
    for (int l = 0; l < TF_LAYERS; ++l) {
      const std::string q = t + "seq_tfmr_" + bs + ".layers." + std::to_string(l);
      v.push_back({q + ".self_attn.in_proj_weight", {3 * TF_D, TF_D}});
      v.push_back({q + ".self_attn.in_proj_bias", {3 * TF_D}});
      lin(q + ".self_attn.out_proj", TF_D, TF_D); lin(q + ".linear1", TF_D, TF_D); lin(q + ".linear2", TF_D, TF_D);
      ln(q + ".norm1", TF_D); ln(q + ".norm2", TF_D);
    }
    lin(t + "post_tfmr_" + bs, C_S, TF_D);
    const std::string nt = t + "node_transition_" + bs;
    lin(nt + ".linear_1", C_S, C_S); lin(nt + ".linear_2", C_S, C_S); lin(nt + ".linear_3", C_S, C_S); ln(nt + ".ln", C_S);
    lin(t + "bb_update_" + bs + ".linear", 6, C_S);
    if (b < NBLK - 1) {
      const std::string et = t + "edge_transition_" + bs;
      lin(et + ".initial_embed", C_S / 2, C_S); lin(et + ".trunk.0", ET_HID, ET_HID); lin(et + ".trunk.2", ET_HID, ET_HID);
      lin(et + ".final_layer", C_Z, ET_HID); ln(et + ".layer_norm", C_Z);
    }
  }
  const std::string tp = "score_model.torsion_pred";
  lin(tp + ".linear_1", C_S, C_S); lin(tp + ".linear_2", C_S, C_S); lin(tp + ".linear_final", 2, C_S);
  return v;
}

// keys that exist in the reference state_dict but are never used by the forward pass
bool is_unused_key(const std::string& k) {
  return k.find(".linear_rbf.") != std::string::npos || k.find("torsion_pred.linear_3.") != std::string::npos;
}

// Split a Linear weight (K <= 320) into its fp16 hi|lo operand images once; Lin finds it again by (pointer, stride, N, K).
int pack_linear(fdpt_ctx* ctx, const float* W, int ldw, int N, int K) {
  if (K > tc::LT_MAX_KB * 64) return FDPT_OK;
  const auto key = std::make_tuple(W, ldw, N, K);
  fdpt_ctx::PackedW& pw = ctx->packed[key];
  pw.nkb = (K + 63) / 64;
  pw.n_tiles = (N + 127) / 128;
  const size_t bytes = (size_t)pw.n_tiles * pw.nkb * tc::LT_STAGE_BYTES;
  if (!pw.img) CK(cudaMalloc(&pw.img, bytes));
  const long long chunks = (long long)pw.n_tiles * pw.nkb * 128 * 8;
  tc::pack_weight_split_kernel<<<(unsigned)((chunks + 255) / 256), 256>>>(W, ldw, N, K, pw.nkb, pw.n_tiles, pw.img);
  CK(cudaGetLastError());
  return FDPT_OK;
}

// One-time repack of the IPA parameters of a block (host side; kernels_ipa.cuh header describes the layouts).
int pack_ipa_params(fdpt_ctx* ctx, BlockParams& p) {
  std::vector<float> Wq((size_t)NH * C_HID * C_S), bq(NH * C_HID), Wkv((size_t)2 * NH * C_HID * C_S), bkv(2 * NH * C_HID);
  std::vector<float> Wqp((size_t)NH * PQ * 3 * C_S), bqp(NH * PQ * 3), Wkvp((size_t)NH * (PQ + PV) * 3 * C_S), bkvp(NH * (PQ + PV) * 3);
  std::vector<float> Wout((size_t)C_S * CAT), Wb((size_t)NH * C_Z);
  auto d2h = [&](std::vector<float>& h, const float* d) { return cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost); };
  CK(d2h(Wq, p.Wq)); CK(d2h(bq, p.bq)); CK(d2h(Wkv, p.Wkv)); CK(d2h(bkv, p.bkv)); CK(d2h(Wqp, p.Wqp)); CK(d2h(bqp, p.bqp));
  CK(d2h(Wkvp, p.Wkvp)); CK(d2h(bkvp, p.bkvp)); CK(d2h(Wout, p.Wout)); CK(d2h(Wb, p.Wb));
  std::vector<float> Wcat((size_t)PROJ_W * C_S), bcat(PROJ_W), Wop((size_t)C_S * CAT);
  auto put = [&](int dst_row, const std::vector<float>& W, const std::vector<float>& b, int src_row) {
    memcpy(&Wcat[(size_t)dst_row * C_S], &W[(size_t)src_row * C_S], C_S * sizeof(float));
    bcat[dst_row] = b[src_row];
  };
  const int NKV = PQ + PV;
  for (int h = 0; h < NH; ++h) {
    for (int c = 0; c < C_HID; ++c) {
      put(PROJ_Q + h * QK_W + c, Wq, bq, h * C_HID + c);                 // linear_q: [H, C] (ipa_pytorch.py:201-204)
      put(PROJ_K + h * QK_W + c, Wkv, bkv, h * 2 * C_HID + c);           // linear_kv: [H, 2C], k = first C (208-211)
      put(PROJ_V + h * V_W + c, Wkv, bkv, h * 2 * C_HID + C_HID + c);
    }
    for (int ax = 0; ax < 3; ++ax) {  // points: output split in 3 chunks x | y | z, each [H, P] (214-239)
      for (int q = 0; q < PQ; ++q) {
        put(PROJ_Q + h * QK_W + C_HID + ax * PQ + q, Wqp, bqp, ax * NH * PQ + h * PQ + q);
        put(PROJ_K + h * QK_W + C_HID + ax * PQ + q, Wkvp, bkvp, ax * NH * NKV + h * NKV + q);
      }
      for (int q = 0; q < PV; ++q) put(PROJ_V + h * V_W + C_HID + ax * PV + q, Wkvp, bkvp, ax * NH * NKV + h * NKV + PQ + q);
    }
  }
  // linear_out columns: reference concat [o (H*C) | o_pt x | y | z (H*PV each) | norms | o_pair]  ->  cat' order
  std::vector<int> src_col(CAT);
  for (int h = 0; h < NH; ++h) {
    for (int c = 0; c < C_HID; ++c) src_col[h * V_W + c] = h * C_HID + c;
    for (int ax = 0; ax < 3; ++ax)
      for (int q = 0; q < PV; ++q) src_col[h * V_W + C_HID + ax * PV + q] = CAT_OPT + ax * NH * PV + h * PV + q;
  }
  for (int k = CATP_NRM; k < CAT; ++k) src_col[k] = k;  // norms and o_pair keep their places
  for (int n = 0; n < C_S; ++n)
    for (int k = 0; k < CAT; ++k) Wop[(size_t)n * CAT + k] = Wout[(size_t)n * CAT + src_col[k]];
  // linear_b operand image: [2 k-blocks][16 rows][128 B swizzled]; rows 0-7 = fp16(W_b), rows 8-15 = fp16(W_b - fp16(W_b))
  std::vector<__half> img(2 * 16 * 64);
  for (int kb = 0; kb < 2; ++kb)
    for (int n = 0; n < 16; ++n)
      for (int c = 0; c < 8; ++c)
        for (int e = 0; e < 8; ++e) {
          const float wv = Wb[(size_t)(n & 7) * C_Z + kb * 64 + c * 8 + e];
          const __half hi = __float2half_rn(wv);
          const __half v = n < 8 ? hi : __float2half_rn(wv - __half2float(hi));
          img[(size_t)kb * 1024 + n * 64 + ((c ^ (n & 7)) << 3) + e] = v;
        }
  {
    constexpr int BLK = 320, NPROJ = 3 * NH * BLK;
    std::vector<float> W2((size_t)NPROJ * C_S, 0.f), b2(NPROJ, 0.f);
    const float s_qk = sqrtf(1.0f / (3.f * C_HID));
    for (int kind = 0; kind < 3; ++kind)
      for (int h = 0; h < NH; ++h) {
        const int src0 = kind == 0 ? PROJ_Q + h * QK_W : (kind == 1 ? PROJ_K + h * QK_W : PROJ_V + h * V_W);
        const int width = kind == 2 ? V_W : QK_W;
        for (int c = 0; c < width; ++c) {
          const float sc = (kind == 0 && c < C_HID) ? s_qk : 1.f;
          const size_t dst = (size_t)((kind * NH + h) * BLK + c);
          for (int k = 0; k < C_S; ++k) W2[dst * C_S + k] = Wcat[(size_t)(src0 + c) * C_S + k] * sc;
          b2[dst] = bcat[src0 + c] * sc;
        }
      }
    if (!p.Wimgproj) CK(cudaMalloc(&p.Wimgproj, W2.size() * sizeof(float)));
    if (!p.bimgproj) CK(cudaMalloc(&p.bimgproj, b2.size() * sizeof(float)));
    CK(cudaMemcpy(p.Wimgproj, W2.data(), W2.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p.bimgproj, b2.data(), b2.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  {
    constexpr int NKB = CAT / 64, NT = (C_S + 127) / 128;
    if (!p.Wout_perm) CK(cudaMalloc(&p.Wout_perm, Wop.size() * sizeof(float)));
    CK(cudaMemcpy(p.Wout_perm, Wop.data(), Wop.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (!p.imgWout) CK(cudaMalloc(&p.imgWout, (size_t)NT * NKB * tc::LT_STAGE_BYTES));
    const long long chunks = (long long)NT * NKB * 128 * 8;
    tc::pack_weight_split_kernel<<<(unsigned)((chunks + 255) / 256), 256>>>(p.Wout_perm, CAT, C_S, CAT, NKB, NT, p.imgWout);
    CK(cudaGetLastError());
  }
  if (!p.Wcat) CK(cudaMalloc(&p.Wcat, Wcat.size() * sizeof(float)));
  if (!p.bcat) CK(cudaMalloc(&p.bcat, bcat.size() * sizeof(float)));
  if (!p.Wout_perm) CK(cudaMalloc(&p.Wout_perm, Wop.size() * sizeof(float)));
  if (!p.imgWb) CK(cudaMalloc(&p.imgWb, img.size() * sizeof(__half)));
  CK(cudaMemcpy(p.Wcat, Wcat.data(), Wcat.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.bcat, bcat.data(), bcat.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.Wout_perm, Wop.data(), Wop.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p.imgWb, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
  return FDPT_OK;
}

// ---- workspace ---------------------------------------------------------------------------------------
template <typename T>
T* carve(char*& p, size_t n) {
  T* r = reinterpret_cast<T*>(p);
  p += ((n * sizeof(T) + 255) / 256) * 256;
  return r;
}

int reserve_ws(fdpt_ctx* ctx, int B, int N) {
  Workspace& w = ctx->ws;
  if (w.base && B <= w.capB && N == w.capN) return FDPT_OK;  // tile-image layout (and its zero padding) is per N
  CK(cudaDeviceSynchronize());
  if (w.base) CK(cudaFree(w.base));
  w = Workspace();
  const size_t M = (size_t)B * N;
  const int R = 4 * 4096;  // capacity of the relative-offset table
  for (int pass = 0; pass < 2; ++pass) {
    char* p = pass ? w.base : nullptr;
    w.node_feat = carve<float>(p, M * 96); w.feat1d = carve<float>(p, M * 64);
    w.node0 = carve<float>(p, M * C_S); w.node = carve<float>(p, M * C_S);
    w.tmpA = carve<float>(p, M * ET_HID); w.tmpB = carve<float>(p, M * ET_HID); w.tmpC = carve<float>(p, M * ET_HID);
    w.PA = carve<float>(p, M * C_Z);
    w.proj = carve<float>(p, M * PROJ_W); w.kn = carve<float>(p, M * NH);
    w.cat = carve<float>(p, M * CAT);
    w.tf_x = carve<float>(p, M * TF_D); w.qkv = carve<float>(p, M * 3 * TF_D); w.att_o = carve<float>(p, M * TF_D);
    w.upd = carve<float>(p, M * 8); w.quats = carve<float>(p, M * 4); w.trans = carve<float>(p, M * 4); w.dmask = carve<float>(p, M);
    w.n_emb = carve<float>(p, M * C_Z); w.U = carve<float>(p, M * ET_HID); w.V = carve<float>(p, M * ET_HID);
    w.Pf = carve<float>(p, M * C_Z); w.Qf = carve<float>(p, M * C_Z); w.tors_u = carve<float>(p, M * 2);
    const size_t JB = (size_t)(N + 127) / 128;
    w.z = carve<__half>(p, M * JB * 16384); w.n_img = carve<__half>(p, (size_t)B * JB * 16384);
    w.f_img = carve<__half>(p, (size_t)B * JB * 8192); w.rel_tab = carve<__half>(p, (size_t)R * EMB);
    w.S = carve<float>(p, M * NH * (size_t)((N + 3) & ~3));
    w.qkv_img_bytes = (size_t)B * NH * JB * IPA_IMG_KB * tc::LT_STAGE_BYTES;
    w.p_img_bytes = (size_t)B * NH * JB * 2 * JB * tc::LT_STAGE_BYTES;
    w.qimg = carve<uint8_t>(p, w.qkv_img_bytes); w.kimg = carve<uint8_t>(p, w.qkv_img_bytes); w.vimg = carve<uint8_t>(p, w.qkv_img_bytes);
    w.pimg = carve<uint8_t>(p, w.p_img_bytes);
    w.cat_img_bytes = ((M + 127) / 128) * (size_t)(CAT / 64) * tc::LT_STAGE_BYTES;
    w.cat_img = carve<uint8_t>(p, w.cat_img_bytes);
    w.tf_qkv_img_bytes = (size_t)B * TF_H * JB * 2 * tc::LT_STAGE_BYTES;
    w.tf_p_img_bytes = (size_t)B * TF_H * JB * 2 * JB * tc::LT_STAGE_BYTES;
    w.tf_qimg = carve<uint8_t>(p, w.tf_qkv_img_bytes); w.tf_kimg = carve<uint8_t>(p, w.tf_qkv_img_bytes); w.tf_vimg = carve<uint8_t>(p, w.tf_qkv_img_bytes);
    w.tf_pimg = carve<uint8_t>(p, w.tf_p_img_bytes);
    w.pred_rigids = carve<float>(p, M * 7); w.trans_score = carve<float>(p, M * 3); w.psi = carve<float>(p, M * 2);
    w.rig_cur = carve<float>(p, M * 7); w.rig_next = carve<float>(p, M * 7); w.sc_ca = carve<float>(p, M * 3);
    w.t_emb_b = carve<float>(p, (size_t)B * EMB); w.t32_b = carve<float>(p, B); w.bb_tmp = carve<float>(p, M * 15);
    w.img_bytes = ((M + 127) / 128) * (size_t)tc::LT_MAX_KB * tc::LT_STAGE_BYTES;
    w.imgF = carve<__half>(p, w.img_bytes / 2); w.imgT1 = carve<__half>(p, w.img_bytes / 2); w.imgT2 = carve<__half>(p, w.img_bytes / 2); w.imgN = carve<__half>(p, w.img_bytes / 2);
    w.rot_score = carve<double>(p, M * 3); w.sigma_b = carve<double>(p, B); w.sigma_idx_b = carve<int>(p, B); w.sched_dev = carve<double>(p, 4096 * FDPT_SCHED_COLS); w.step_dev = carve<int>(p, 64); w.call_ptrs = carve<void*>(p, 16); w.temb_tab = carve<float>(p, 4096 * EMB);
    if (!pass) {
      w.bytes = (size_t)(p - (char*)nullptr);
      CK(cudaMalloc(&w.base, w.bytes));
    }
  }
  w.capB = B; w.capN = N; w.JB = (N + 127) / 128; w.ldS = (N + 3) & ~3;
  CK(cudaMemset(w.z, 0, sizeof(__half) * M * w.JB * 16384));  // padded rows (j >= N) of the tile images stay zero
  CK(cudaMemset(w.imgF, 0, 4 * w.img_bytes));
  CK(cudaMemset(w.cat_img, 0, w.cat_img_bytes));  // rows >= M of the last tile stay zero
  CK(cudaMemset(w.tf_qimg, 0, 3 * w.tf_qkv_img_bytes + w.tf_p_img_bytes));
  CK(cudaMemset(w.qimg, 0, 3 * w.qkv_img_bytes + w.p_img_bytes));  // padding rows / columns of the IPA operand images stay zero                  // rows >= M of the last m-tile of the chained operand images stay zero
  return FDPT_OK;
}

// ---- small launch helpers ------------------------------------------------------------------------------
struct Lin {
  fdpt_ctx* ctx;
  cudaStream_t st;
  // y[M,N] (ldc) = epi(x[M,K] (lda) @ W[N,K]^T (ldb))
  int operator()(const float* x, int lda, const float* W, int ldb, const float* bias, float* y, int ldc, long long M, int N, int K,
                 int relu = 0, const float* residual = nullptr, int ldr = 0, const float* rowmask = nullptr, int accumulate = 0,
                 const __half* x_img = nullptr, __half* y_img = nullptr, const tc::IpaProjEpi* ipa = nullptr, int epi = 1) const {
    // x_img / y_img: operand-image chaining between consecutive Linear layers (lin_tc.cuh); only valid on the packed lin_tc path
    if (ctx->gemm_tc && !accumulate && M > 0) {
      auto it = ctx->packed.find(std::make_tuple(W, ldb, N, K));
      if (it != ctx->packed.end()) {
        const auto& pw = it->second;
        tc::LinTcArgs a;
        a.X = x; a.ldx = lda; a.M = (int)M; a.K = K; a.N = N; a.Wimg = pw.img; a.nkb = pw.nkb; a.n_tiles = pw.n_tiles;
        const int m_tiles = (int)((M + 127) / 128);
        {
          // n-tiles per CTA: minimise waves x (tiles per CTA + fixed per-CTA cost in tile units).  The plain "spread over the SMs" rule
          // left the 60-tile IPA projection with 154 CTAs on 148 SMs: a second wave of six CTAs doubled its time
          int best = 1;
          long long best_cost = -1;
          for (int tpc = 1; tpc <= pw.n_tiles; ++tpc) {
            const long long ctas = (long long)m_tiles * ((pw.n_tiles + tpc - 1) / tpc);
            const long long cost = ((ctas + ctx->num_sms - 1) / ctx->num_sms) * (tpc + 2);
            if (best_cost < 0 || cost < best_cost) {
              best_cost = cost;
              best = tpc;
            }
          }
          a.tiles_per_cta = best;
        }
        a.stg_cols = (ctx->dbg_flags & 1024) ? 32 : 16;  // 16-column staging patches: measured faster than 32 for every layer shape (profiles/r01_bench_lin_ablation.txt)
        if (ipa) a.stg_cols = -1;  // the image-writing epilogues (IPA projection, transformer in_proj) need no transposition patches: 17 KB more for the weight ring
        a.units = std::min(8, (int)((ctx->max_smem_optin - tc::lin_tc_fixed_bytes(pw.nkb, a.stg_cols)) / tc::LT_UNIT_BYTES));
        a.bias = bias; a.relu = relu; a.rowmask = rowmask; a.residual = residual; a.ldr = ldr; a.Y = y; a.ldy = ldc;
        a.dbg_flags = ctx->dbg_flags;
        a.dbg = (ctx->dbg_flags & (512 | (1 << 19))) ? ctx->et_dbg : nullptr;  // bit 19: stamps inside the weight-resident variant
        if (y_img) a.dbg = reinterpret_cast<long long*>(y_img);  // lin_tc_kernel<*, YIMG = true> writes the output image through this field
        if (x_img) a.X = reinterpret_cast<const float*>(x_img);  // lin_tc_kernel<XIMG = true> reads X as the operand image
        a.x_vec = tc::aligned16(x, lda, 0, 0);
        a.y_vec = (y && tc::aligned16(y, ldc, 0, 0) && (!residual || tc::aligned16(residual, ldr, 0, 0)) && N % 4 == 0) ? 1 : 0;
        dim3 grid(m_tiles, (pw.n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(tc::LT_THREADS);
        cfg.dynamicSmemBytes = tc::lin_tc_smem_bytes(pw.nkb, a.units, a.stg_cols);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = tc::g_use_pdl;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e;
        if (!ipa && ctx->lin_wres && a.tiles_per_cta == 1 && !(ctx->dbg_flags & (1 | 2 | 512))) {
          // one n-tile per CTA: weight-resident variant (the whole panel is prefetched under the predecessor kernel's tail)
          cfg.dynamicSmemBytes = tc::lin_tcw_smem_bytes(pw.nkb);
          if (x_img && y_img) e = cudaLaunchKernelEx(&cfg, tc::lin_tcw_kernel<true, true>, a);
          else if (x_img) e = cudaLaunchKernelEx(&cfg, tc::lin_tcw_kernel<true, false>, a);
          else if (y_img) e = cudaLaunchKernelEx(&cfg, tc::lin_tcw_kernel<false, true>, a);
          else e = cudaLaunchKernelEx(&cfg, tc::lin_tcw_kernel<false, false>, a);
        } else if (ipa) {
          a.ipa = *ipa;
          a.dbg = nullptr;
          if (epi == 2) e = x_img ? cudaLaunchKernelEx(&cfg, tc::lin_tc_kernel<true, false, 2>, a) : cudaLaunchKernelEx(&cfg, tc::lin_tc_kernel<false, false, 2>, a);
          else e = x_img ? cudaLaunchKernelEx(&cfg, tc::lin_tc_kernel<true, false, 1>, a) : cudaLaunchKernelEx(&cfg, tc::lin_tc_kernel<false, false, 1>, a);
        } else if (x_img && y_img) e = cudaLaunchKernelEx(&cfg, tc::lin_tc_kernel<true, true>, a);
        else if (x_img) e = cudaLaunchKernelEx(&cfg, tc::lin_tc_kernel<true, false>, a);
        else if (y_img) e = cudaLaunchKernelEx(&cfg, tc::lin_tc_kernel<false, true>, a);
        else e = cudaLaunchKernelEx(&cfg, tc::lin_tc_kernel<false, false>, a);
        ctx->launches++;
        if (e != cudaSuccess) return fail(ctx, FDPT_ERR_CUDA, "lin_tc launch: %s", cudaGetErrorString(e));
        return FDPT_OK;
      }
    }
    if (x_img || y_img || ipa) return fail(ctx, FDPT_ERR_STATE, "operand-image chaining needs the packed lin_tc path");
    GemmArgs g;
    g.A = x; g.lda = lda; g.B = W; g.ldb = ldb; g.C = y; g.ldc = ldc; g.M = (int)M; g.N = N; g.K = K;
    g.bias = bias; g.relu = relu; g.residual = residual; g.ldr = ldr; g.rowmask = rowmask; g.accumulate = accumulate;
    cudaError_t e = gemm_dispatch(ctx, g, true, 1, st);
    ctx->launches++;
    if (e != cudaSuccess) return fail(ctx, FDPT_ERR_CUDA, "gemm launch: %s", cudaGetErrorString(e));
    return FDPT_OK;
  }
};

#define RET(x)                 \
  do {                         \
    int r__ = (x);             \
    if (r__ != FDPT_OK) return r__; \
  } while (0)

template <int C>
int layernorm(fdpt_ctx* ctx, cudaStream_t st, const float* x, float* y, const float* g, const float* b, long long rows,
              const float* rowmask, const float* pairmask = nullptr, int nres = 0, long long row0 = 0) {
  if (rows <= 0) return FDPT_OK;
  layernorm_kernel<C><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, y, g, b, rows, rowmask, pairmask, nres, row0);
  LAUNCH_CHECK();
  return FDPT_OK;
}

// ---- embedder -------------------------------------------------------------------------------------------
int run_embed(fdpt_ctx* ctx, int B, int N, const fdpt_feats* in, float* node_out, __half* z_out, cudaStream_t st, float* node_copy = nullptr) {
  ProfScope ps(ctx, FDPT_PROF_EDGE_EMBED, st);
  Workspace& w = ctx->ws;
  const long long M = (long long)B * N, P = M * N;
  const int F1 = f1_dim(ctx), FN = F1 + EMB, EIN = ein_dim(ctx);
  Lin lin{ctx, st};
  if (in->rel_count <= 0 || in->rel_count > 4 * 4096) return fail(ctx, FDPT_ERR_INVALID, "rel_count %d out of range", in->rel_count);
  {
    dim3 blk(32, 8);
    node_feats_kernel<<<(unsigned)((M + 7) / 8), blk, 0, st>>>((int)M, N, ctx->cfg.with_aatype, in->aatype, in->fixed_mask, in->t_emb,
                                                                 in->t_emb_eps, in->idx_emb, w.feat1d, w.node_feat);
    LAUNCH_CHECK();
  }
  auto& T = ctx->top;
  if (ctx->gemm_tc && !(ctx->dbg_flags & 32768)) {  // hidden activations as operand images (lin_tc Ximg / Yimg)
    RET(lin(w.node_feat, FN, T.nW0, FN, T.nb0, nullptr, C_S, M, C_S, FN, 1, nullptr, 0, nullptr, 0, nullptr, w.imgT1));
    RET(lin(nullptr, C_S, T.nW2, C_S, T.nb2, nullptr, C_S, M, C_S, C_S, 1, nullptr, 0, nullptr, 0, w.imgT1, w.imgT2));
    RET(lin(nullptr, C_S, T.nW4, C_S, T.nb4, w.tmpA, C_S, M, C_S, C_S, 0, nullptr, 0, nullptr, 0, w.imgT2, nullptr));
  } else {
    RET(lin(w.node_feat, FN, T.nW0, FN, T.nb0, w.tmpA, C_S, M, C_S, FN, 1));
    RET(lin(w.tmpA, C_S, T.nW2, C_S, T.nb2, w.tmpB, C_S, M, C_S, C_S, 1));
    RET(lin(w.tmpB, C_S, T.nW4, C_S, T.nb4, w.tmpA, C_S, M, C_S, C_S, 0));
  }
  RET(layernorm<C_S>(ctx, st, w.tmpA, node_out, T.nln_g, T.nln_b, M, in->res_mask));
  if (node_copy) CK(cudaMemcpyAsync(node_copy, node_out, sizeof(float) * M * C_S, cudaMemcpyDeviceToDevice, st));
  // edge embedder (edge_embed_fused.cuh): W0 = [A (F1) | B (F1) | C (32) | D (22)];  PA_i = A f_i + b0 per residue (fp32 class),
  // everything pair-sized inside one fused tcgen05 kernel
  RET(lin(w.feat1d, F1, T.eW0, EIN, T.eb0, w.PA, C_Z, M, C_Z, F1, 0));
  {
    const long long chunks = (long long)B * w.JB * 128 * 8;
    tc::f_to_image_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(B, N, w.JB, F1, w.feat1d, w.f_img);
    LAUNCH_CHECK();
    const long long nrel = (long long)in->rel_count * EMB;
    tc::f32_to_f16_kernel<<<(unsigned)((nrel + 255) / 256), 256, 0, st>>>(nrel, in->rel_emb, w.rel_tab);
    LAUNCH_CHECK();
  }
  tc::EeArgs a;
  a.B = B; a.N = N; a.JB = w.JB; a.z_out = z_out; a.f_img = w.f_img; a.PA = w.PA; a.rel_tab = w.rel_tab; a.seq_idx = in->seq_idx;
  a.rel_min = in->rel_min; a.rel_count = in->rel_count; a.sc_ca = in->sc_ca_t; a.bin_lower = ctx->bin_lower; a.b2 = T.eb2; a.b4 = T.eb4;
  a.ln_g = T.eln_g; a.ln_b = T.eln_b; a.mask = in->res_mask; a.W0img = T.imgE0; a.W2img = T.imgE2; a.W4img = T.imgE4;
  a.tiles = M * w.JB;
  (void)P;
  a.dbg = (ctx->dbg_flags & 8192) ? ctx->et_dbg : nullptr;
  if (a.tiles >= (1LL << 31)) return fail(ctx, FDPT_ERR_INVALID, "B*N*ceil(N/128) = %lld tiles: the pair kernels index tiles with 32 bits", a.tiles);
  tc::ee_fused_kernel<<<(unsigned)std::min<long long>(ctx->num_sms, a.tiles), tc::EE_THREADS, tc::ee_smem_bytes(), st>>>(a);
  LAUNCH_CHECK();
  return FDPT_OK;
}

int launch_img_gemm(fdpt_ctx* ctx, const tc::GemmImgArgs& g, cudaStream_t st) {
  const long long tiles = (long long)g.m_tiles * g.n_tiles * g.batch;
  if (tiles <= 0) return FDPT_OK;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)std::min<long long>(ctx->num_sms, tiles));
  cfg.blockDim = dim3(tc::GI_THREADS);
  cfg.dynamicSmemBytes = tc::gemm_img_smem_bytes();
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = tc::g_use_pdl;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, tc::gemm_img_kernel, g);
  ctx->launches++;
  if (e != cudaSuccess) return fail(ctx, FDPT_ERR_CUDA, "gemm_img launch: %s", cudaGetErrorString(e));
  return FDPT_OK;
}

// ---- IPA (kernels_ipa.cuh) ------------------------------------------------------------------------------------
int run_ipa(fdpt_ctx* ctx, int blk, int B, int N, const float* s, const __half* z, const float* quats, const float* trans,
            const float* mask, float* out, int ldo, const float* residual, const float* outmask, cudaStream_t st,
            const float* ln_g = nullptr, const float* ln_b = nullptr, float* ln_out = nullptr, const __half* s_img = nullptr) {
  // s_img: optional operand image of s: the projection GEMM then skips its staging phase
  ProfScope ps(ctx, FDPT_PROF_IPA_TOTAL, st);
  Workspace& w = ctx->ws;
  const BlockParams& p = ctx->blk[blk];
  const long long M = (long long)B * N;
  Lin lin{ctx, st};
  if (w.JB > IPA_MAX_JB) return fail(ctx, FDPT_ERR_INVALID, "N=%d: the IPA kernel supports N <= %d", N, IPA_MAX_JB * 128);
  // q | q_pts, k | k_pts, v | v_pts of every head in one GEMM, then the frames applied in place
  const bool img = ctx->ipa_img && ctx->gemm_tc;
  // ipa_img mode 1: the projection GEMM's epilogue applies the frames and writes the Q' / K' / V' operand images itself (IpaProjEpi);
  //          mode 2: fp32 projection + separate ipa_prep_img kernel (A/B switch)
  const bool img_fused = img && ctx->ipa_img == 1;
  if (img_fused) {
    tc::IpaProjEpi e;
    e.quats = quats; e.trans = trans; e.head_w = p.head_w; e.mask = mask; e.kbias = w.kn; e.Qimg = w.qimg; e.Kimg = w.kimg; e.Vimg = w.vimg;
    e.n_res = N; e.JB = w.JB;
    RET(lin(s, C_S, p.Wimgproj, C_S, p.bimgproj, nullptr, 0, M, 3 * NH * 320, C_S, 0, nullptr, 0, nullptr, 0, s_img, nullptr, &e));
  } else {
    RET(lin(s, C_S, p.Wcat, C_S, p.bcat, w.proj, PROJ_W, M, PROJ_W, C_S, 0, nullptr, 0, nullptr, 0, s_img, nullptr));
  }
  std::unique_ptr<ProfScope> pattn(new ProfScope(ctx, FDPT_PROF_IPA_ATTN, st));  // every kernel that implements the attention itself: prep .. opt
  if (img) {
    // Q', K', V' as operand images (frames applied, s_qk and gamma folded into Q'), then S[b,h] = Q'_h K'_h^T + kbias on gemm_img
    IpaImgArgs ia;
    ia.M = (int)M; ia.N = N; ia.JB = w.JB; ia.proj = w.proj; ia.quats = quats; ia.trans = trans; ia.head_w = p.head_w; ia.mask = mask;
    ia.kbias = w.kn; ia.Qimg = w.qimg; ia.Kimg = w.kimg; ia.Vimg = w.vimg;
    if (!img_fused) {
      ipa_prep_img_kernel<<<(unsigned)M, 256, 0, st>>>(ia);
      LAUNCH_CHECK();
    }
    tc::GemmImgArgs g;
    memset(&g, 0, sizeof(g));
    const long long per_bh = (long long)w.JB * IPA_IMG_KB * tc::LT_STAGE_BYTES;
    g.A = w.qimg; g.sA = per_bh; g.nkbA = IPA_IMG_KB; g.B = w.kimg; g.sB = per_bh; g.nkbB = IPA_IMG_KB; g.b_mn = 0; g.nkb = IPA_IMG_KB;
    g.M = N; g.N = N; g.m_tiles = w.JB; g.n_tiles = w.JB; g.batch = B * NH; g.batch2 = 1;
    g.bias = w.kn; g.sBias = N; g.C = w.S; g.ldc = w.ldS; g.sC1 = (long long)N * w.ldS; g.sC2 = 0;
    RET(launch_img_gemm(ctx, g, st));
  } else {
  ipa_prep_kernel<<<(unsigned)M, 256, 0, st>>>((int)M, w.proj, quats, trans, p.head_w, mask, N, w.kn);
  LAUNCH_CHECK();
  {  // S[b,h] = s_qk * [q | g q_pts] . [k | k_pts]^T
    GemmArgs g;
    g.A = w.proj + PROJ_Q; g.lda = PROJ_W; g.sA1 = (long long)N * PROJ_W; g.sA2 = QK_W;
    g.B = w.proj + PROJ_K; g.ldb = PROJ_W; g.sB1 = (long long)N * PROJ_W; g.sB2 = QK_W;
    g.C = w.S; g.ldc = w.ldS; g.sC1 = (long long)NH * N * w.ldS; g.sC2 = (long long)N * w.ldS;
    g.M = N; g.N = N; g.K = QK_W; g.batch2 = NH; g.alpha = sqrtf(1.0f / (3.f * C_HID));
    g.bias = w.kn; g.sBias1 = (long long)NH * N; g.sBias2 = N;  // kbias[b][h][j]
    CK(gemm_dispatch(ctx, g, true, B * NH, st));
    ctx->launches++;
  }
  }
  {
    const IpaSmemPlan plan = ipa_core_plan(N, ctx->max_smem_optin, ctx->max_smem_sm, (ctx->dbg_flags & 65536) != 0);  // debug flag 65536: two-pass kernel also for N <= 384
    if (plan.rz < 1) return fail(ctx, FDPT_ERR_INVALID, "N=%d needs more shared memory than available in ipa_core", N);
    IpaCoreArgs a;
    a.B = B; a.N = N; a.JB = w.JB; a.ldS = w.ldS; a.S = w.S; a.z = z; a.Wb_img = p.imgWb; a.bb = p.bb;
    a.Wd = p.Wd; a.single_pass = plan.single_pass; a.bd = p.bd; a.cat = w.cat; a.rz = plan.rz; a.tmem_cols = plan.tmem_cols; a.rows = (int)M; a.Pg = img ? w.pimg : nullptr; a.mn_swap = ctx->mn_swap; a.dbg = (ctx->dbg_flags & 4) ? ctx->et_dbg : nullptr;
    ProfScope pc(ctx, FDPT_PROF_IPA_CORE, st);
    if (plan.single_pass)
      ipa_core_kernel<true><<<(unsigned)std::min<long long>((long long)ctx->num_sms, M), IPA_THREADS, plan.bytes, st>>>(a);
    else
      ipa_core_kernel<false><<<(unsigned)std::min<long long>((long long)ctx->num_sms * plan.ctas_per_sm, M), IPA_THREADS, plan.bytes, st>>>(a);
    LAUNCH_CHECK();
  }
  if (img) {  // [o | o_pt (global frame)][b,:,h,:] = P_h V'_h  -> cat'[:, h*292 : (h+1)*292]   (V' rows = j: MN-major B)
    tc::GemmImgArgs g;
    memset(&g, 0, sizeof(g));
    g.A = w.pimg; g.sA = (long long)w.JB * 2 * w.JB * tc::LT_STAGE_BYTES; g.nkbA = 2 * w.JB;
    g.B = w.vimg; g.sB = (long long)w.JB * IPA_IMG_KB * tc::LT_STAGE_BYTES; g.nkbB = IPA_IMG_KB; g.b_mn = 1; g.nkb = (N + 63) / 64;
    g.M = N; g.N = V_W; g.m_tiles = w.JB; g.n_tiles = (V_W + 127) / 128; g.batch = B * NH; g.batch2 = NH;
    g.bias = nullptr; g.C = w.cat; g.ldc = CAT; g.sC1 = (long long)N * CAT; g.sC2 = V_W;
    RET(launch_img_gemm(ctx, g, st));
  } else {  // [o | o_pt (global frame)][b,:,h,:] = A_h [V_h | v_pts_h]  -> cat'[:, h*292 : (h+1)*292]
    GemmArgs g;
    g.A = w.S; g.lda = w.ldS; g.sA1 = (long long)NH * N * w.ldS; g.sA2 = (long long)N * w.ldS;
    g.B = w.proj + PROJ_V; g.ldb = PROJ_W; g.sB1 = (long long)N * PROJ_W; g.sB2 = V_W;
    g.C = w.cat; g.ldc = CAT; g.sC1 = (long long)N * CAT; g.sC2 = V_W;
    g.M = N; g.N = V_W; g.K = N; g.batch2 = NH;
    CK(gemm_dispatch(ctx, g, false, B * NH, st));
    ctx->launches++;
  }
  const int OUT_SPLIT = (ctx->dbg_flags & 16384) ? 2 : 3;  // K = 2688 = 42 k-blocks = 3 x 14 (debug flag 16384: 2 x 21)
  const bool split_ok = ln_out && ctx->gemm_tc && !(ctx->dbg_flags & 256) && (CAT / 64) % OUT_SPLIT == 0 && w.tmpB - w.tmpA == w.tmpC - w.tmpB;
  if (img && split_ok && ctx->ipa_img != 3) {
    // inverse frames + norms AND the operand image of the concat row in one pass; linear_out then multiplies two ready images
    // (3-way split-K over the 42 k-blocks), the reduce is folded into the LayerNorm kernel as before
    ipa_opt_img_kernel<<<(unsigned)M, 256, 0, st>>>((int)M, w.cat, quats, trans, w.cat_img);
    LAUNCH_CHECK();
    pattn.reset();
    tc::GemmImgArgs g;
    memset(&g, 0, sizeof(g));
    const int nkb_all = CAT / 64, nkb = nkb_all / OUT_SPLIT;
    g.A = w.cat_img; g.sA = (long long)nkb * tc::LT_STAGE_BYTES; g.nkbA = nkb_all;
    g.B = reinterpret_cast<const uint8_t*>(p.imgWout); g.sB = (long long)nkb * tc::LT_STAGE_BYTES; g.nkbB = nkb_all; g.b_mn = 0; g.nkb = nkb;
    g.M = (int)M; g.N = C_S; g.m_tiles = (int)((M + 127) / 128); g.n_tiles = (C_S + 127) / 128; g.batch = OUT_SPLIT; g.batch2 = 1;
    g.bias = nullptr; g.C = w.tmpA; g.ldc = C_S; g.sC1 = w.tmpB - w.tmpA; g.sC2 = 0;
    RET(launch_img_gemm(ctx, g, st));
    sumk_layernorm_kernel<C_S><<<(unsigned)((M + 7) / 8), 256, 0, st>>>(w.tmpA, w.tmpB - w.tmpA, OUT_SPLIT, p.bout, residual, outmask, ln_out, ln_g,
                                                                        ln_b, M);
    LAUNCH_CHECK();
    return FDPT_OK;
  }
  ipa_opt_kernel<<<(unsigned)M, NH * PV, 0, st>>>((int)M, w.cat, quats, trans);
  LAUNCH_CHECK();
  pattn.reset();
  if (ln_out && ctx->gemm_tc && !(ctx->dbg_flags & 256) && (CAT / 64) % OUT_SPLIT == 0 && w.tmpB - w.tmpA == w.tmpC - w.tmpB) {
    // linear_out as a 3-way split-K GEMM (one long serial k-loop per CTA otherwise: 88 CTAs x 42 k-blocks; split: 264 CTAs x 14); the
    // partial products are summed, biased, masked and added to the residual inside the LayerNorm kernel that follows
    GemmArgs g;
    g.A = w.cat; g.lda = CAT; g.sA1 = CAT / OUT_SPLIT; g.B = p.Wout_perm; g.ldb = CAT; g.sB1 = CAT / OUT_SPLIT;
    g.C = w.tmpA; g.ldc = C_S; g.sC1 = w.tmpB - w.tmpA; g.M = (int)M; g.N = C_S; g.K = CAT / OUT_SPLIT;
    CK(gemm_dispatch(ctx, g, true, OUT_SPLIT, st));
    ctx->launches++;
    sumk_layernorm_kernel<C_S><<<(unsigned)((M + 7) / 8), 256, 0, st>>>(w.tmpA, w.tmpB - w.tmpA, OUT_SPLIT, p.bout, residual, outmask, ln_out, ln_g,
                                                                        ln_b, M);
    LAUNCH_CHECK();
    return FDPT_OK;
  }
  // linear_out (+ mask, + residual)
  RET(lin(w.cat, CAT, p.Wout_perm, CAT, p.bout, out, ldo, M, C_S, CAT, 0, residual, C_S, outmask));
  if (ln_out) RET(layernorm<C_S>(ctx, st, out, ln_out, ln_g, ln_b, M, nullptr));
  return FDPT_OK;
}

// ---- edge transition (fused tcgen05 kernel, et_fused.cuh) ------------------------------------------------------
int run_edge_transition(fdpt_ctx* ctx, int blk, int B, int N, const float* node, const __half* z_in, const float* mask, __half* z_out,
                        cudaStream_t st) {
  ProfScope ps(ctx, FDPT_PROF_EDGE_TRANSITION, st);
  Workspace& w = ctx->ws;
  const BlockParams& p = ctx->blk[blk];
  const long long M = (long long)B * N;
  Lin lin{ctx, st};
  // per-residue parts: n = initial_embed(node); U_i = W1[:,128:256] n_i + b1; Pf_i = Wf[:,128:256] n_i + bf
  if (ctx->gemm_tc && !(ctx->dbg_flags & 32768)) {  // n is written as fp32 (for the n_j images) AND as the operand image of the two GEMMs below
    RET(lin(node, C_S, p.Wie, C_S, p.bie, w.n_emb, C_Z, M, C_Z, C_S, 0, nullptr, 0, nullptr, 0, nullptr, w.imgN));
    RET(lin(nullptr, C_Z, p.We1 + C_Z, ET_HID, p.be1, w.U, ET_HID, M, ET_HID, C_Z, 0, nullptr, 0, nullptr, 0, w.imgN, nullptr));
    RET(lin(nullptr, C_Z, p.Wef + C_Z, ET_HID, p.bef, w.Pf, C_Z, M, C_Z, C_Z, 0, nullptr, 0, nullptr, 0, w.imgN, nullptr));
  } else {
    RET(lin(node, C_S, p.Wie, C_S, p.bie, w.n_emb, C_Z, M, C_Z, C_S));
    RET(lin(w.n_emb, C_Z, p.We1 + C_Z, ET_HID, p.be1, w.U, ET_HID, M, ET_HID, C_Z));
    RET(lin(w.n_emb, C_Z, p.Wef + C_Z, ET_HID, p.bef, w.Pf, C_Z, M, C_Z, C_Z));
  }
  {
    const long long chunks = (long long)B * w.JB * 128 * 16;
    tc::n_to_image_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(B, N, w.JB, w.n_emb, w.n_img);
    LAUNCH_CHECK();
  }
  tc::EtArgs a;
  a.B = B; a.N = N; a.JB = w.JB; a.z_in = z_in; a.z_out = z_out; a.n_img = w.n_img; a.Ui = w.U; a.Pf = w.Pf; a.b2 = p.be2;
  a.ln_g = p.eln_g; a.ln_b = p.eln_b; a.mask = mask; a.W1cat = p.imgW1cat; a.W2 = p.imgW2; a.W3cat = p.imgW3cat;
  a.tiles = M * w.JB;
  a.exp = (ctx->dbg_flags >> 20) & 15;
  if (a.tiles >= (1LL << 31)) return fail(ctx, FDPT_ERR_INVALID, "B*N*ceil(N/128) = %lld tiles: the pair kernels index tiles with 32 bits", a.tiles);
  a.dbg = (ctx->dbg_flags & 4) ? nullptr : ctx->et_dbg;
  {
    const int grid = (int)std::min<long long>(ctx->num_sms, a.tiles);
    tc::et_fused_kernel<<<grid, tc::ET_THREADS, tc::et_smem_bytes(), st>>>(a);
  }
  LAUNCH_CHECK();
  return FDPT_OK;
}

// 4-head attention of one encoder layer on qkv [M, 960] -> att_o [M, 320] (torch.nn.MultiheadAttention inside TransformerEncoderLayer,
// ipa_pytorch.py:433-443; boolean key-padding semantics)
int seq_attention(fdpt_ctx* ctx, int B, int N, const float* mask, cudaStream_t st, bool img = false) {
  Workspace& w = ctx->ws;
  if (img) {
    // the in_proj epilogue wrote Q (pre-scaled by 1/sqrt(d_head)), K, V as operand images: S = Q K^T, softmax -> P image, O = P V on gemm_img
    const long long per_bh = (long long)w.JB * 2 * tc::LT_STAGE_BYTES;
    tc::GemmImgArgs g;
    memset(&g, 0, sizeof(g));
    g.A = w.tf_qimg; g.sA = per_bh; g.nkbA = 2; g.B = w.tf_kimg; g.sB = per_bh; g.nkbB = 2; g.b_mn = 0; g.nkb = 2;
    g.M = N; g.N = N; g.m_tiles = w.JB; g.n_tiles = w.JB; g.batch = B * TF_H; g.batch2 = 1;
    g.bias = nullptr; g.C = w.S; g.ldc = w.ldS; g.sC1 = (long long)N * w.ldS; g.sC2 = 0;
    RET(launch_img_gemm(ctx, g, st));
    const long long rows = (long long)B * TF_H * N;
    softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(w.S, rows, N, w.ldS, TF_H * N, 1.0f, mask, w.tf_pimg, w.JB);
    LAUNCH_CHECK();
    memset(&g, 0, sizeof(g));
    g.A = w.tf_pimg; g.sA = (long long)w.JB * 2 * w.JB * tc::LT_STAGE_BYTES; g.nkbA = 2 * w.JB;
    g.B = w.tf_vimg; g.sB = per_bh; g.nkbB = 2; g.b_mn = 1; g.nkb = (N + 63) / 64;
    g.M = N; g.N = TF_DH; g.m_tiles = w.JB; g.n_tiles = 1; g.batch = B * TF_H; g.batch2 = TF_H;
    g.bias = nullptr; g.C = w.att_o; g.ldc = TF_D; g.sC1 = (long long)N * TF_D; g.sC2 = TF_DH;
    RET(launch_img_gemm(ctx, g, st));
    return FDPT_OK;
  }
    {
    GemmArgs g;  // S[b,h] = Q K^T  (rows padded to ldS floats: 16-byte aligned rows for the softmax and the P V operand loads)
    g.A = w.qkv; g.lda = 3 * TF_D; g.sA1 = (long long)N * 3 * TF_D; g.sA2 = TF_DH;
    g.B = w.qkv + TF_D; g.ldb = 3 * TF_D; g.sB1 = (long long)N * 3 * TF_D; g.sB2 = TF_DH;
    g.C = w.S; g.ldc = w.ldS; g.sC1 = (long long)TF_H * N * w.ldS; g.sC2 = (long long)N * w.ldS;
    g.M = N; g.N = N; g.K = TF_DH; g.batch2 = TF_H;
    CK(gemm_dispatch(ctx, g, true, B * TF_H, st));
    ctx->launches++;
  }
  {
    const long long rows = (long long)B * TF_H * N;
    softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(w.S, rows, N, w.ldS, TF_H * N, 1.0f / sqrtf((float)TF_DH), mask);
    LAUNCH_CHECK();
  }
  {
    GemmArgs g;  // O = P V
    g.A = w.S; g.lda = w.ldS; g.sA1 = (long long)TF_H * N * w.ldS; g.sA2 = (long long)N * w.ldS;
    g.B = w.qkv + 2 * TF_D; g.ldb = 3 * TF_D; g.sB1 = (long long)N * 3 * TF_D; g.sB2 = TF_DH;
    g.C = w.att_o; g.ldc = TF_D; g.sC1 = (long long)N * TF_D; g.sC2 = TF_DH;
    g.M = N; g.N = TF_DH; g.K = N; g.batch2 = TF_H;
    CK(gemm_dispatch(ctx, g, false, B * TF_H, st));
    ctx->launches++;
  }
  return FDPT_OK;
}

// ---- sequence transformer (post-norm encoder, ipa_pytorch.py:433-443, 533-539) ------------------------------------
int run_seq_tfmr(fdpt_ctx* ctx, int blk, int B, int N, const float* mask, cudaStream_t st) {
  ProfScope ps(ctx, FDPT_PROF_SEQ_TFMR, st);
  Workspace& w = ctx->ws;
  const BlockParams& p = ctx->blk[blk];
  const long long M = (long long)B * N;
  Lin lin{ctx, st};
  // x = [node | skip_embed(init_node)]
  copy_cols_kernel<<<(unsigned)((M * C_S + 255) / 256), 256, 0, st>>>(M, C_S, w.node, C_S, w.tf_x, TF_D, 0, nullptr);
  LAUNCH_CHECK();
  const bool chain = ctx->gemm_tc && !(ctx->dbg_flags & 32768);  // debug flag 32768: no operand-image chaining
  RET(lin(w.node0, C_S, p.Wskip, C_S, p.bskip, w.tf_x + C_S, TF_D, M, C_SKIP, C_S));
  for (int l = 0; l < TF_LAYERS; ++l) {
    const auto& L = p.tf[l];
    const bool img = ctx->gemm_tc && ctx->tf_img && L.Win_s;
    if (img) {
      tc::IpaProjEpi e;
      memset(&e, 0, sizeof(e));
      e.Qimg = w.tf_qimg; e.Kimg = w.tf_kimg; e.Vimg = w.tf_vimg; e.n_res = N; e.JB = w.JB;
      RET(lin(w.tf_x, TF_D, L.Win_s, TF_D, L.bin_s, nullptr, 0, M, 3 * TF_D, TF_D, 0, nullptr, 0, nullptr, 0, nullptr, nullptr, &e, 2));
    } else {
      RET(lin(w.tf_x, TF_D, L.Win, TF_D, L.bin, w.qkv, 3 * TF_D, M, 3 * TF_D, TF_D));
    }
    RET(seq_attention(ctx, B, N, mask, st, img));
    RET(lin(w.att_o, TF_D, L.Wo, TF_D, L.bo, w.tmpA, TF_D, M, TF_D, TF_D, 0, w.tf_x, TF_D));
    RET(layernorm<TF_D>(ctx, st, w.tmpA, w.tf_x, L.n1g, L.n1b, M, nullptr));
    if (chain) {  // linear1 hands its ReLU output to linear2 as a ready operand image (no fp32 round trip, no re-split)
      RET(lin(w.tf_x, TF_D, L.W1, TF_D, L.b1, nullptr, TF_D, M, TF_D, TF_D, 1, nullptr, 0, nullptr, 0, nullptr, w.imgF));
      RET(lin(nullptr, TF_D, L.W2, TF_D, L.b2, w.tmpB, TF_D, M, TF_D, TF_D, 0, w.tf_x, TF_D, nullptr, 0, w.imgF, nullptr));
    } else {
      RET(lin(w.tf_x, TF_D, L.W1, TF_D, L.b1, w.tmpA, TF_D, M, TF_D, TF_D, 1));
      RET(lin(w.tmpA, TF_D, L.W2, TF_D, L.b2, w.tmpB, TF_D, M, TF_D, TF_D, 0, w.tf_x, TF_D));
    }
    RET(layernorm<TF_D>(ctx, st, w.tmpB, w.tf_x, L.n2g, L.n2b, M, nullptr));
  }
  // node = node + post_tfmr(x)
  RET(lin(w.tf_x, TF_D, p.Wpost, TF_D, p.bpost, w.node, C_S, M, C_S, TF_D, 0, w.node, C_S));
  return FDPT_OK;
}

// ---- full forward -------------------------------------------------------------------------------------------------
int forward_impl(fdpt_ctx* ctx, int B, int N, const fdpt_feats* in, const fdpt_out* out, cudaStream_t st) {
  if (!ctx->finalized) return fail(ctx, FDPT_ERR_STATE, "parameters not finalised");
  if (B <= 0 || N <= 0) return fail(ctx, FDPT_ERR_INVALID, "bad B=%d N=%d", B, N);
  RET(reserve_ws(ctx, B, N));
  ProfScope pfwd(ctx, FDPT_PROF_FORWARD, st);
  Workspace& w = ctx->ws;
  const long long M = (long long)B * N;
  Lin lin{ctx, st};
  const float cs = ctx->cfg.coordinate_scaling;
  RET(run_embed(ctx, B, N, in, w.node0, w.z, st, w.node));
  init_frames_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>((int)M, in->rigids_t, cs, in->res_mask, in->fixed_mask, w.quats, w.trans,
                                                                  w.dmask);
  LAUNCH_CHECK();
  for (int b = 0; b < NBLK; ++b) {
    const BlockParams& p = ctx->blk[b];
    // node = LN(node + ipa(node) * mask)
    RET(run_ipa(ctx, b, B, N, w.node, w.z, w.quats, w.trans, in->res_mask, w.tmpC, C_S, w.node, in->res_mask, st, p.ln_g, p.ln_b, w.node));
    RET(run_seq_tfmr(ctx, b, B, N, in->res_mask, st));
    // node transition
    if (ctx->gemm_tc && !(ctx->dbg_flags & 32768)) {  // the hidden activations travel as operand images (lin_tc Ximg / Yimg)
      RET(lin(w.node, C_S, p.Wt1, C_S, p.bt1, nullptr, C_S, M, C_S, C_S, 1, nullptr, 0, nullptr, 0, nullptr, w.imgT1));
      RET(lin(nullptr, C_S, p.Wt2, C_S, p.bt2, nullptr, C_S, M, C_S, C_S, 1, nullptr, 0, nullptr, 0, w.imgT1, w.imgT2));
      RET(lin(nullptr, C_S, p.Wt3, C_S, p.bt3, w.tmpA, C_S, M, C_S, C_S, 0, w.node, C_S, nullptr, 0, w.imgT2, nullptr));
    } else {
      RET(lin(w.node, C_S, p.Wt1, C_S, p.bt1, w.tmpA, C_S, M, C_S, C_S, 1));
      RET(lin(w.tmpA, C_S, p.Wt2, C_S, p.bt2, w.tmpB, C_S, M, C_S, C_S, 1));
      RET(lin(w.tmpB, C_S, p.Wt3, C_S, p.bt3, w.tmpA, C_S, M, C_S, C_S, 0, w.node, C_S));
    }
    RET(layernorm<C_S>(ctx, st, w.tmpA, w.node, p.tln_g, p.tln_b, M, in->res_mask));
    // backbone update
    RET(lin(w.node, C_S, p.Wbb, C_S, p.bbb, w.upd, 6, M, 6, C_S));
    compose_update_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>((int)M, w.upd, w.dmask, w.quats, w.trans);
    LAUNCH_CHECK();
    if (b < NBLK - 1) RET(run_edge_transition(ctx, b, B, N, w.node, w.z, in->res_mask, w.z, st));
  }
  // heads
  float* rig = (out && out->rigids) ? out->rigids : w.pred_rigids;
  finish_frames_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>((int)M, w.quats, w.trans, cs, rig);
  LAUNCH_CHECK();
  double* rs = (out && out->rot_score) ? out->rot_score : w.rot_score;
  if (ctx->score_table && !in->sigma_idx) return fail(ctx, FDPT_ERR_INVALID, "a cached score table is installed: feats.sigma_idx is required");
  if (!ctx->score_table && !in->sigma) return fail(ctx, FDPT_ERR_INVALID, "feats.sigma is required");
  rot_score_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>((int)M, N, in->rigids_t, 7, w.quats, 4, in->sigma, in->res_mask, rs, ctx->score_table,
                                                            ctx->omega_bounds, ctx->tab_omega, in->sigma_idx);
  LAUNCH_CHECK();
  float* ts = (out && out->trans_score) ? out->trans_score : w.trans_score;
  trans_score_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>((int)M, N, in->rigids_t + 4, 7, rig + 4, 7, 1.0f, in->t32,
                                                                  (float)ctx->cfg.r3_min_b, (float)ctx->cfg.r3_max_b, ctx->cfg.r3_coordinate_scaling, 1,
                                                                  in->res_mask, ts);
  LAUNCH_CHECK();
  // torsion head (ipa_pytorch.py:347-363)
  auto& T = ctx->top;
  if (ctx->gemm_tc && !(ctx->dbg_flags & 32768)) {
    RET(lin(w.node, C_S, T.tW1, C_S, T.tb1, nullptr, C_S, M, C_S, C_S, 1, nullptr, 0, nullptr, 0, nullptr, w.imgT1));
    RET(lin(nullptr, C_S, T.tW2, C_S, T.tb2, nullptr, C_S, M, C_S, C_S, 0, w.node, C_S, nullptr, 0, w.imgT1, w.imgT2));
    RET(lin(nullptr, C_S, T.tWf, C_S, T.tbf, w.tors_u, 2, M, 2, C_S, 0, nullptr, 0, nullptr, 0, w.imgT2, nullptr));
  } else {
    RET(lin(w.node, C_S, T.tW1, C_S, T.tb1, w.tmpA, C_S, M, C_S, C_S, 1));
    RET(lin(w.tmpA, C_S, T.tW2, C_S, T.tb2, w.tmpB, C_S, M, C_S, C_S, 0, w.node, C_S));
    RET(lin(w.tmpB, C_S, T.tWf, C_S, T.tbf, w.tors_u, 2, M, 2, C_S));
  }
  float* psi = (out && out->psi) ? out->psi : w.psi;
  psi_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>((int)M, w.tors_u, in->fixed_mask, in->gt_psi, psi);
  LAUNCH_CHECK();
  if (out && out->atom37_bb) {
    backbone_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>((int)M, rig, psi, ctx->cfg.with_aatype ? in->aatype : nullptr, ctx->ideal,
                                                                 ctx->psi_frame, ctx->atom_mask, out->atom37_bb);
    LAUNCH_CHECK();
  }
  return FDPT_OK;
}

// fp32 [B,N,N,128] <-> fp16 tile images (unit entry points only; the hot path keeps z in image form)
int z_fp32_to_image(fdpt_ctx* ctx, int B, int N, const float* z, cudaStream_t st) {
  const long long chunks = (long long)B * N * ctx->ws.JB * 128 * 16;
  tc::z_to_image_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(B, N, ctx->ws.JB, z, ctx->ws.z);
  LAUNCH_CHECK();
  return FDPT_OK;
}
int z_image_to_fp32(fdpt_ctx* ctx, int B, int N, float* z, cudaStream_t st) {
  const long long chunks = (long long)B * N * ctx->ws.JB * 128 * 16;
  tc::image_to_z_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(B, N, ctx->ws.JB, ctx->ws.z, z);
  LAUNCH_CHECK();
  return FDPT_OK;
}

__global__ void set_step_kernel(int B, const int* __restrict__ step_ptr, const float* __restrict__ t_emb_tab, const double* __restrict__ sched,
                                float* __restrict__ t_emb_b, float* __restrict__ t32_b, double* __restrict__ sigma_b,
                                int* __restrict__ sigma_idx_b) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int step = *step_ptr;
  if (idx < B * EMB) t_emb_b[idx] = t_emb_tab[step * EMB + (idx % EMB)];
  if (idx < B) {
    t32_b[idx] = (float)sched[step * FDPT_SCHED_COLS + FDPT_SCHED_T32];
    sigma_b[idx] = sched[step * FDPT_SCHED_COLS + FDPT_SCHED_SIGMA];
    sigma_idx_b[idx] = (int)sched[step * FDPT_SCHED_COLS + FDPT_SCHED_SIGMA_IDX];
  }
}

// dst[step slice] = src (n floats)
__global__ void copy_slot_kernel(long long n, const float* __restrict__ src, float* dst, StepRef ref) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) step_resolve(dst, ref)[idx] = src[idx];
}

// end of a timestep: x_t <- x_{t-1}; the last thread bumps the step counter (every other kernel of the step has read it already)
__global__ void advance_state_kernel(long long n, const float* __restrict__ rig_next, float* __restrict__ rig_cur, int* __restrict__ step) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) rig_cur[idx] = rig_next[idx];
  if (idx == 0) *step += 1;
}

__global__ void copy_trans_kernel(int M, const float* __restrict__ rig, float* __restrict__ ca) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < M * 3) ca[idx] = rig[(idx / 3) * 7 + 4 + idx % 3];
}

}  // namespace

// =======================================================================================================
// C ABI
// =======================================================================================================
extern "C" {

const char* fdpt_version(void) { return "framedipt_b200 libfdpt 0.1 (sm_100a)"; }

const char* fdpt_last_error(const fdpt_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int fdpt_create(const fdpt_config* cfg, int device, fdpt_ctx** out) {
  if (!cfg || !out) return FDPT_ERR_INVALID;
  *out = nullptr;
  if (cfg->c_s != C_S || cfg->c_z != C_Z || cfg->c_hidden != C_HID || cfg->c_skip != C_SKIP || cfg->no_heads != NH ||
      cfg->no_qk_points != PQ || cfg->no_v_points != PV || cfg->num_blocks != NBLK || cfg->index_embed_size != EMB ||
      cfg->num_bins != NBINS || cfg->seq_tfmr_num_heads != TF_H || cfg->seq_tfmr_num_layers != TF_LAYERS)
    return FDPT_ERR_INVALID;  // kernels are specialised for the reference's default dims (config/base.yaml:55-79)
  if (!(cfg->coordinate_scaling > 0.f) || !(cfg->r3_coordinate_scaling > 0.f)) return FDPT_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return FDPT_ERR_CUDA;
  DeviceGuard dg(device);
  if (!dg.ok) return FDPT_ERR_CUDA;
  fdpt_ctx* ctx = new fdpt_ctx();
  ctx->cfg = *cfg;
  ctx->device = device;
  cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&ctx->max_smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
  cudaFuncSetAttribute(ipa_core_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(ipa_core_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::et_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::et_smem_bytes());
  cudaFuncSetAttribute(tc::ee_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::ee_smem_bytes());
  cudaFuncSetAttribute(tc::tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::tc_linear_smem_bytes(512));
  cudaFuncSetAttribute(tc::gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::gemm_tc_smem_bytes(128));
  cudaFuncSetAttribute(tc::lin_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::lin_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::lin_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::lin_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::lin_tcw_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::lin_tcw_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::lin_tcw_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::lin_tcw_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::lin_tc_kernel<false, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::lin_tc_kernel<true, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::lin_tc_kernel<false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::lin_tc_kernel<true, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  cudaFuncSetAttribute(tc::gemm_img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::gemm_img_smem_bytes());
  // distogram bin edges: torch.linspace(min_bin, max_bin, num_bins) in float32 (framedipt/data/utils.py:546)
  float lower[NBINS];
  {
    // torch.linspace (CPU, float32): step = (end - start) / (steps - 1); first half start + i*step, second half end - (steps-1-i)*step
    const float start = cfg->min_bin, end = cfg->max_bin;
    const float step = (end - start) / (float)(NBINS - 1);
    for (int i = 0; i < NBINS; ++i) lower[i] = (i < NBINS / 2) ? start + step * (float)i : end - step * (float)(NBINS - 1 - i);
  }
  bool ok = cudaMalloc(&ctx->bin_lower, sizeof(lower)) == cudaSuccess && cudaMalloc(&ctx->ideal, sizeof(kIdealBB)) == cudaSuccess &&
            cudaMalloc(&ctx->psi_frame, sizeof(kPsiFrame)) == cudaSuccess && cudaMalloc(&ctx->atom_mask, sizeof(kAtomMask)) == cudaSuccess;
  ok = ok && cudaMemcpy(ctx->bin_lower, lower, sizeof(lower), cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(ctx->ideal, kIdealBB, sizeof(kIdealBB), cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(ctx->psi_frame, kPsiFrame, sizeof(kPsiFrame), cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(ctx->atom_mask, kAtomMask, sizeof(kAtomMask), cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) {
    delete ctx;
    return FDPT_ERR_CUDA;
  }
  *out = ctx;
  return FDPT_OK;
}

int fdpt_destroy(fdpt_ctx* ctx) {
  if (!ctx) return FDPT_OK;
  GUARD(ctx);
  cudaDeviceSynchronize();
  for (auto& kv : ctx->params) cudaFree(kv.second.dev);
  for (auto& b : ctx->blk) {
    cudaFree(b.imgW1cat);
    cudaFree(b.imgW2);
    cudaFree(b.imgW3cat);
    cudaFree(b.Wcat);
    cudaFree(b.bcat);
    cudaFree(b.Wout_perm);
    for (auto& L : b.tf) {
      cudaFree(L.Win_s);
      cudaFree(L.bin_s);
    }
    cudaFree(b.imgWout);
    cudaFree(b.Wimgproj);
    cudaFree(b.bimgproj);
    cudaFree(b.imgWb);
  }
  if (ctx->step_graph.exec) cudaGraphExecDestroy(ctx->step_graph.exec);
  if (ctx->own_stream) {
    cudaStreamDestroy(ctx->own_stream);
    cudaEventDestroy(ctx->fence_in);
    cudaEventDestroy(ctx->fence_out);
  }
  cudaFreeHost(ctx->et_dbg);
  cudaFree(ctx->score_table);
  cudaFree(ctx->omega_bounds);
  for (auto e : ctx->progress_ev) cudaEventDestroy(e);
  for (auto& kv : ctx->packed) cudaFree(kv.second.img);
  cudaFree(ctx->top.imgE0);
  cudaFree(ctx->top.imgE2);
  cudaFree(ctx->top.imgE4);
  cudaFree(ctx->ws.base);
  cudaFree(ctx->bin_lower);
  cudaFree(ctx->ideal);
  cudaFree(ctx->psi_frame);
  cudaFree(ctx->atom_mask);
  delete ctx;
  return FDPT_OK;
}

int fdpt_num_params_expected(const fdpt_ctx* ctx) { return ctx ? (int)expected_params(ctx).size() : 0; }

int fdpt_load_param(fdpt_ctx* ctx, const char* key, const float* data, const int64_t* shape, int ndim) {
  if (!ctx || !key || !data || !shape) return FDPT_ERR_INVALID;
  GUARD(ctx);
  const std::string k(key);
  if (is_unused_key(k)) return FDPT_OK;  // accepted and ignored, like dead parameters in the reference
  bool found = false;
  for (auto& e : expected_params(ctx)) {
    if (e.first != k) continue;
    found = true;
    if ((int)e.second.size() != ndim) return fail(ctx, FDPT_ERR_PARAM, "size mismatch for %s: expected %zu dims, got %d", key, e.second.size(), ndim);
    for (int i = 0; i < ndim; ++i)
      if (e.second[i] != shape[i]) return fail(ctx, FDPT_ERR_PARAM, "size mismatch for %s: dim %d expected %lld, got %lld", key, i, (long long)e.second[i], (long long)shape[i]);
  }
  if (!found) return fail(ctx, FDPT_ERR_PARAM, "Unexpected key(s) in state_dict: %s", key);
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= (size_t)shape[i];
  ParamSpec& ps = ctx->params[k];
  if (!ps.dev) CK(cudaMalloc(&ps.dev, n * sizeof(float) + 64));
  ps.shape.assign(shape, shape + ndim);
  CK(cudaMemcpy(ps.dev, data, n * sizeof(float), cudaMemcpyDefault));
  ctx->finalized = false;
  ctx->step_graph.key.clear();  // weights are baked into the captured step
  return FDPT_OK;
}

int fdpt_finalize_params(fdpt_ctx* ctx) {
  if (!ctx) return FDPT_ERR_INVALID;
  GUARD(ctx);
  std::string missing;
  for (auto& e : expected_params(ctx))
    if (!ctx->params.count(e.first)) missing += (missing.empty() ? "" : ", ") + e.first;
  if (!missing.empty()) return fail(ctx, FDPT_ERR_PARAM, "Missing key(s) in state_dict: %s", missing.substr(0, 900).c_str());
  auto P = [&](const std::string& k) -> const float* { return ctx->params.at(k).dev; };
  const std::string ne = "embedding_layer.node_embedder", ee = "embedding_layer.edge_embedder", t = "score_model.trunk.";
  auto& T = ctx->top;
  T.nW0 = P(ne + ".0.weight"); T.nb0 = P(ne + ".0.bias"); T.nW2 = P(ne + ".2.weight"); T.nb2 = P(ne + ".2.bias");
  T.nW4 = P(ne + ".4.weight"); T.nb4 = P(ne + ".4.bias"); T.nln_g = P(ne + ".5.weight"); T.nln_b = P(ne + ".5.bias");
  T.eW0 = P(ee + ".0.weight"); T.eb0 = P(ee + ".0.bias"); T.eW2 = P(ee + ".2.weight"); T.eb2 = P(ee + ".2.bias");
  T.eW4 = P(ee + ".4.weight"); T.eb4 = P(ee + ".4.bias"); T.eln_g = P(ee + ".5.weight"); T.eln_b = P(ee + ".5.bias");
  const std::string tp = "score_model.torsion_pred";
  T.tW1 = P(tp + ".linear_1.weight"); T.tb1 = P(tp + ".linear_1.bias"); T.tW2 = P(tp + ".linear_2.weight"); T.tb2 = P(tp + ".linear_2.bias");
  T.tWf = P(tp + ".linear_final.weight"); T.tbf = P(tp + ".linear_final.bias");
  {
    const int F1 = f1_dim(ctx), EIN = ein_dim(ctx);
    if (!T.imgE0) CK(cudaMalloc(&T.imgE0, 32768));
    if (!T.imgE2) CK(cudaMalloc(&T.imgE2, 32768));
    if (!T.imgE4) CK(cudaMalloc(&T.imgE4, 32768));
    auto pack = [&](const float* W, int ldw, int K, int Kvalid, __half* img) {
      const long long chunks = (long long)C_Z * (K / 8);
      tc::pack_weight_image_kernel<<<(unsigned)((chunks + 255) / 256), 256>>>(W, ldw, C_Z, K, Kvalid, img);
    };
    // k-block 0: [C | D | 0]; without self-conditioning features the D columns do not exist: their image columns stay zero, so the
    // one-hot distogram bits the kernel builds multiply zeros
    pack(T.eW0 + 2 * F1, EIN, 64, EIN - 2 * F1, T.imgE0);
    pack(T.eW0 + F1, EIN, 64, F1, T.imgE0 + 8192);          // k-block 1: [B | 0]
    pack(T.eW2, C_Z, C_Z, C_Z, T.imgE2);
    pack(T.eW4, C_Z, C_Z, C_Z, T.imgE4);
    CK(cudaGetLastError());
  }
  for (int b = 0; b < NBLK; ++b) {
    BlockParams& p = ctx->blk[b];
    const std::string bs = std::to_string(b), ip = t + "ipa_" + bs;
    p.head_w = P(ip + ".head_weights");
    p.Wq = P(ip + ".linear_q.weight"); p.bq = P(ip + ".linear_q.bias"); p.Wkv = P(ip + ".linear_kv.weight"); p.bkv = P(ip + ".linear_kv.bias");
    p.Wqp = P(ip + ".linear_q_points.weight"); p.bqp = P(ip + ".linear_q_points.bias");
    p.Wkvp = P(ip + ".linear_kv_points.weight"); p.bkvp = P(ip + ".linear_kv_points.bias");
    p.Wb = P(ip + ".linear_b.weight"); p.bb = P(ip + ".linear_b.bias"); p.Wd = P(ip + ".down_z.weight"); p.bd = P(ip + ".down_z.bias");
    p.Wout = P(ip + ".linear_out.weight"); p.bout = P(ip + ".linear_out.bias");
    RET(pack_ipa_params(ctx, p));
    p.ln_g = P(t + "ipa_ln_" + bs + ".weight"); p.ln_b = P(t + "ipa_ln_" + bs + ".bias");
    p.Wskip = P(t + "skip_embed_" + bs + ".weight"); p.bskip = P(t + "skip_embed_" + bs + ".bias");
    for (int l = 0; l < TF_LAYERS; ++l) {
      const std::string q = t + "seq_tfmr_" + bs + ".layers." + std::to_string(l);
      auto& L = p.tf[l];
      L.Win = P(q + ".self_attn.in_proj_weight"); L.bin = P(q + ".self_attn.in_proj_bias");
      L.Wo = P(q + ".self_attn.out_proj.weight"); L.bo = P(q + ".self_attn.out_proj.bias");
      L.W1 = P(q + ".linear1.weight"); L.b1 = P(q + ".linear1.bias"); L.W2 = P(q + ".linear2.weight"); L.b2 = P(q + ".linear2.bias");
      L.n1g = P(q + ".norm1.weight"); L.n1b = P(q + ".norm1.bias"); L.n2g = P(q + ".norm2.weight"); L.n2b = P(q + ".norm2.bias");
      {  // in_proj with the 1/sqrt(d_head) of the attention logits folded into the q rows (rows 0 .. d_model-1)
        std::vector<float> Wh((size_t)3 * TF_D * TF_D), bh(3 * TF_D);
        CK(cudaMemcpy(Wh.data(), L.Win, Wh.size() * sizeof(float), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(bh.data(), L.bin, bh.size() * sizeof(float), cudaMemcpyDeviceToHost));
        const float sc = 1.0f / sqrtf((float)TF_DH);
        for (int r = 0; r < TF_D; ++r) {
          for (int k = 0; k < TF_D; ++k) Wh[(size_t)r * TF_D + k] *= sc;
          bh[r] *= sc;
        }
        if (!L.Win_s) CK(cudaMalloc(&L.Win_s, Wh.size() * sizeof(float)));
        if (!L.bin_s) CK(cudaMalloc(&L.bin_s, bh.size() * sizeof(float)));
        CK(cudaMemcpy(L.Win_s, Wh.data(), Wh.size() * sizeof(float), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(L.bin_s, bh.data(), bh.size() * sizeof(float), cudaMemcpyHostToDevice));
      }
    }
    p.Wpost = P(t + "post_tfmr_" + bs + ".weight"); p.bpost = P(t + "post_tfmr_" + bs + ".bias");
    const std::string nt = t + "node_transition_" + bs;
    p.Wt1 = P(nt + ".linear_1.weight"); p.bt1 = P(nt + ".linear_1.bias"); p.Wt2 = P(nt + ".linear_2.weight"); p.bt2 = P(nt + ".linear_2.bias");
    p.Wt3 = P(nt + ".linear_3.weight"); p.bt3 = P(nt + ".linear_3.bias"); p.tln_g = P(nt + ".ln.weight"); p.tln_b = P(nt + ".ln.bias");
    p.Wbb = P(t + "bb_update_" + bs + ".linear.weight"); p.bbb = P(t + "bb_update_" + bs + ".linear.bias");
    if (b < NBLK - 1) {
      const std::string et = t + "edge_transition_" + bs;
      p.Wie = P(et + ".initial_embed.weight"); p.bie = P(et + ".initial_embed.bias");
      p.We1 = P(et + ".trunk.0.weight"); p.be1 = P(et + ".trunk.0.bias"); p.We2 = P(et + ".trunk.2.weight"); p.be2 = P(et + ".trunk.2.bias");
      p.Wef = P(et + ".final_layer.weight"); p.bef = P(et + ".final_layer.bias");
      p.eln_g = P(et + ".layer_norm.weight"); p.eln_b = P(et + ".layer_norm.bias");
      // fp16 swizzled images: W1cat = [W1[:,0:128] | W1[:,256:384]] (384 x 256), W2 (384 x 384),
      // W3cat = [Wf | Wf[:,0:128] | Wf[:,256:384]] (128 x 640); image layout [k-block][rows][128 B]
      if (!p.imgW1cat) CK(cudaMalloc(&p.imgW1cat, sizeof(__half) * ET_HID * 256));
      if (!p.imgW2) CK(cudaMalloc(&p.imgW2, sizeof(__half) * ET_HID * ET_HID));
      if (!p.imgW3cat) CK(cudaMalloc(&p.imgW3cat, sizeof(__half) * C_Z * 640));
      auto pack = [&](const float* W, int ldw, int Nrows, int K, __half* img_kb0) {
        const long long chunks = (long long)Nrows * (K / 8);
        tc::pack_weight_image_kernel<<<(unsigned)((chunks + 255) / 256), 256>>>(W, ldw, Nrows, K, K, img_kb0);
      };
      pack(p.We1, ET_HID, ET_HID, 128, p.imgW1cat);
      pack(p.We1 + 2 * C_Z, ET_HID, ET_HID, 128, p.imgW1cat + (size_t)2 * ET_HID * 64);
      pack(p.We2, ET_HID, ET_HID, ET_HID, p.imgW2);
      pack(p.Wef, ET_HID, C_Z, ET_HID, p.imgW3cat);
      pack(p.Wef, ET_HID, C_Z, 128, p.imgW3cat + (size_t)6 * C_Z * 64);
      pack(p.Wef + 2 * C_Z, ET_HID, C_Z, 128, p.imgW3cat + (size_t)8 * C_Z * 64);
      CK(cudaGetLastError());
    }
  }
  {  // pre-split every Linear weight the node side multiplies with (lin_tc.cuh)
    const int F1 = f1_dim(ctx), FN = F1 + EMB, EIN = ein_dim(ctx);
    RET(pack_linear(ctx, T.nW0, FN, C_S, FN)); RET(pack_linear(ctx, T.nW2, C_S, C_S, C_S)); RET(pack_linear(ctx, T.nW4, C_S, C_S, C_S));
    RET(pack_linear(ctx, T.eW0, EIN, C_Z, F1));
    RET(pack_linear(ctx, T.tW1, C_S, C_S, C_S)); RET(pack_linear(ctx, T.tW2, C_S, C_S, C_S)); RET(pack_linear(ctx, T.tWf, C_S, 2, C_S));
    for (int b = 0; b < NBLK; ++b) {
      BlockParams& p = ctx->blk[b];
      RET(pack_linear(ctx, p.Wcat, C_S, PROJ_W, C_S)); RET(pack_linear(ctx, p.Wskip, C_S, C_SKIP, C_S));
      RET(pack_linear(ctx, p.Wimgproj, C_S, 3 * NH * 320, C_S));
      for (int l = 0; l < TF_LAYERS; ++l) {
        auto& L = p.tf[l];
        RET(pack_linear(ctx, L.Win, TF_D, 3 * TF_D, TF_D)); RET(pack_linear(ctx, L.Wo, TF_D, TF_D, TF_D));
        RET(pack_linear(ctx, L.Win_s, TF_D, 3 * TF_D, TF_D));
        RET(pack_linear(ctx, L.W1, TF_D, TF_D, TF_D)); RET(pack_linear(ctx, L.W2, TF_D, TF_D, TF_D));
      }
      RET(pack_linear(ctx, p.Wpost, TF_D, C_S, TF_D));
      RET(pack_linear(ctx, p.Wt1, C_S, C_S, C_S)); RET(pack_linear(ctx, p.Wt2, C_S, C_S, C_S)); RET(pack_linear(ctx, p.Wt3, C_S, C_S, C_S));
      RET(pack_linear(ctx, p.Wbb, C_S, 6, C_S));
      if (b < NBLK - 1) {
        RET(pack_linear(ctx, p.Wie, C_S, C_Z, C_S)); RET(pack_linear(ctx, p.We1 + C_Z, ET_HID, ET_HID, C_Z));
        RET(pack_linear(ctx, p.Wef + C_Z, ET_HID, C_Z, C_Z));
      }
    }
  }
  CK(cudaDeviceSynchronize());
  ctx->finalized = true;
  return FDPT_OK;
}

int fdpt_reserve(fdpt_ctx* ctx, int B, int N) {
  if (!ctx || B <= 0 || N <= 0) return FDPT_ERR_INVALID;
  GUARD(ctx);
  return reserve_ws(ctx, B, N);
}

int fdpt_profile_enable(fdpt_ctx* ctx, int on) {
  if (!ctx) return FDPT_ERR_INVALID;
  ctx->prof_on = on != 0;
  return FDPT_OK;
}

int fdpt_profile_read(fdpt_ctx* ctx, int slot, int* count, double* total_ms) {
  if (!ctx || !count || !total_ms || slot < 0 || slot >= FDPT_PROF_SLOTS) return FDPT_ERR_INVALID;
  GUARD(ctx);
  CK(cudaDeviceSynchronize());
  *count = 0;
  *total_ms = 0.0;
  std::vector<fdpt_ctx::ProfRec> keep;
  for (auto& r : ctx->prof) {
    if (r.slot != slot) {
      keep.push_back(r);
      continue;
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      *total_ms += ms;
      (*count)++;
    }
    ctx->ev_pool.push_back(r.a);
    ctx->ev_pool.push_back(r.b);
  }
  ctx->prof.swap(keep);
  return FDPT_OK;
}

int64_t fdpt_workspace_bytes(const fdpt_ctx* ctx) { return ctx ? (int64_t)ctx->ws.bytes : 0; }
int64_t fdpt_launch_count(const fdpt_ctx* ctx) { return ctx ? ctx->launches : 0; }
int64_t fdpt_stat(const fdpt_ctx* ctx, int which) {
  if (!ctx) return 0;
  switch (which) {
    case FDPT_STAT_GRAPH_CAPTURES: return ctx->stat_captures;
    case FDPT_STAT_SAMPLE_HOST_US: return ctx->stat_sample_host_us;
    default: return 0;
  }
}

int fdpt_forward(fdpt_ctx* ctx, int B, int N, const fdpt_feats* in, const fdpt_out* out, void* stream) {
  if (!ctx || !in) return FDPT_ERR_INVALID;
  GUARD(ctx);
  return forward_impl(ctx, B, N, in, out, (cudaStream_t)stream);
}

int fdpt_reverse(fdpt_ctx* ctx, int B, int N, const float* rigids_t, const double* rot_score, const float* trans_score,
                 const float* diffuse_mask, const double* z_rot, const double* z_trans, const double* sched_row, int center,
                 int diffuse_rot, int diffuse_trans, float* rigids_out, void* stream) {
  if (!ctx || !rigids_t || !rot_score || !trans_score || !diffuse_mask || !z_rot || !z_trans || !sched_row || !rigids_out) return FDPT_ERR_INVALID;
  GUARD(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  RET(reserve_ws(ctx, std::max(B, ctx->ws.capB), std::max(N, ctx->ws.capN)));
  CK(cudaMemcpyAsync(ctx->ws.sched_dev, sched_row, sizeof(double) * FDPT_SCHED_COLS, cudaMemcpyHostToDevice, st));
  ReverseArgs a;
  a.N = N; a.rigids_t = rigids_t; a.rot_score = rot_score; a.trans_score = trans_score; a.dmask = diffuse_mask; a.z_rot = z_rot;
  a.z_trans = z_trans; a.sched = ctx->ws.sched_dev; a.center = center; a.diffuse_rot = diffuse_rot; a.diffuse_trans = diffuse_trans;
  a.cs = ctx->cfg.r3_coordinate_scaling; a.rigids_out = rigids_out;
  reverse_kernel<<<B, 256, 0, st>>>(a);
  LAUNCH_CHECK();
  return FDPT_OK;
}

int fdpt_backbone(fdpt_ctx* ctx, int B, int N, const float* rigids, const float* psi, const int32_t* aatype, float* atom37_bb,
                  void* stream) {
  if (!ctx || !rigids || !psi || !atom37_bb) return FDPT_ERR_INVALID;
  GUARD(ctx);
  const int M = B * N;
  backbone_kernel<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(M, rigids, psi, aatype, ctx->ideal, ctx->psi_frame, ctx->atom_mask, atom37_bb);
  LAUNCH_CHECK();
  return FDPT_OK;
}

int fdpt_rot_score(fdpt_ctx* ctx, int B, int N, const float* quats_t, const float* quats_0, const double* sigma, const float* mask,
                   double* out, void* stream) {
  if (!ctx || !quats_t || !quats_0 || !sigma || !out) return FDPT_ERR_INVALID;
  GUARD(ctx);
  const int M = B * N;
  rot_score_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(M, N, quats_t, 4, quats_0, 4, sigma, mask, out);
  LAUNCH_CHECK();
  return FDPT_OK;
}

int fdpt_trans_score(fdpt_ctx* ctx, int B, int N, const float* trans_t, const float* trans_0, const float* t32, const float* mask,
                     int scale, float* out, void* stream) {
  if (!ctx || !trans_t || !trans_0 || !t32 || !out) return FDPT_ERR_INVALID;
  GUARD(ctx);
  const int M = B * N;
  trans_score_kernel<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(M, N, trans_t, 3, trans_0, 3, 1.0f, t32, (float)ctx->cfg.r3_min_b,
                                                                        (float)ctx->cfg.r3_max_b, ctx->cfg.r3_coordinate_scaling, scale, mask, out);
  LAUNCH_CHECK();
  return FDPT_OK;
}

int fdpt_sample(fdpt_ctx* ctx, int B, int N, const fdpt_feats* feats, int num_t, const double* sched, const float* t_emb_tab,
                const double* noise, uint64_t philox_seed, int self_condition, int center, int diffuse_rot, int diffuse_trans,
                const fdpt_traj* out, void* stream) {
  if (!ctx || !feats || !sched || !t_emb_tab || !out || num_t <= 0 || num_t > 4096) return FDPT_ERR_INVALID;
  const int use_philox = noise == nullptr;  /* throughput mode: normals drawn on the device; otherwise noise rows are indexed by step */
  GUARD(ctx);
  const auto host_t0 = std::chrono::steady_clock::now();
  struct HostTimer {
    fdpt_ctx* c; std::chrono::steady_clock::time_point t0;
    ~HostTimer() { c->stat_sample_host_us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count(); }
  } host_timer{ctx, host_t0};
  cudaStream_t user_st = (cudaStream_t)stream, st = user_st;
  RET(reserve_ws(ctx, B, N));
  const bool fenced = ctx->use_graph && (user_st == nullptr || user_st == cudaStreamLegacy || user_st == cudaStreamPerThread);
  if (fenced) {
    if (!ctx->own_stream) {
      CK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&ctx->fence_in, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&ctx->fence_out, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(ctx->fence_in, user_st));
    CK(cudaStreamWaitEvent(ctx->own_stream, ctx->fence_in, 0));
    st = ctx->own_stream;
  }
  Workspace& w = ctx->ws;
  const long long M = (long long)B * N;
  const int T = num_t;
  CK(cudaMemcpyAsync(w.sched_dev, sched, sizeof(double) * FDPT_SCHED_COLS * T, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(w.rig_cur, feats->rigids_t, sizeof(float) * M * 7, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(w.sc_ca, feats->sc_ca_t, sizeof(float) * M * 3, cudaMemcpyDeviceToDevice, st));
  const int fo = out->final_only;
  if (out->rigid_traj && !fo) CK(cudaMemcpyAsync(out->rigid_traj + (long long)T * M * 7, w.rig_cur, sizeof(float) * M * 7, cudaMemcpyDeviceToDevice, st));
  fdpt_feats f = *feats;
  f.rigids_t = w.rig_cur; f.sc_ca_t = w.sc_ca; f.t_emb = w.t_emb_b; f.t32 = w.t32_b; f.sigma = w.sigma_b; f.sigma_idx = w.sigma_idx_b;
  fdpt_out o;
  memset(&o, 0, sizeof(o));
  o.rigids = w.pred_rigids; o.rot_score = w.rot_score; o.trans_score = w.trans_score; o.psi = w.psi;
  const unsigned gM3 = (unsigned)((M * 3 + 255) / 256), gM = (unsigned)((M + 127) / 128), gM7 = (unsigned)((M * 7 + 255) / 256);
  {  // per-call state read by the (possibly replayed) step: counter, T, buffer bases, timestep-embedding table
    const int hdr[2] = {0, T};
    void* ptrs[5] = {(void*)noise, (void*)out->prot_traj, (void*)out->rigid_0_traj, (void*)out->trans_traj, (void*)out->rigid_traj};
    CK(cudaMemcpyAsync(w.step_dev, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(w.call_ptrs, ptrs, sizeof(ptrs), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(w.temb_tab, t_emb_tab, sizeof(float) * T * EMB, cudaMemcpyDeviceToDevice, st));
    // (pageable host sources: the runtime stages them before returning, so the stack arrays may go out of scope)
  }
  auto set_step = [&]() -> int {
    set_step_kernel<<<(B * EMB + 255) / 256, 256, 0, st>>>(B, w.step_dev, w.temb_tab, w.sched_dev, w.t_emb_b, w.t32_b, w.sigma_b, w.sigma_idx_b);
    LAUNCH_CHECK();
    return FDPT_OK;
  };
  const bool update_sc = !(self_condition & 2);  // inference_fn(embed_self_conditioning=False): sc_ca_t keeps its initial value
  if (self_condition & 1) {  // experiments/utils.py:571-578 (step counter = 0: t = 1.0)
    RET(set_step());
    RET(forward_impl(ctx, B, N, &f, &o, st));
    copy_trans_kernel<<<gM3, 256, 0, st>>>((int)M, w.pred_rigids, w.sc_ca);
    LAUNCH_CHECK();
  }
  // residue types of the trajectory's backbone atoms: the inference_fn call's own inpainting / input_aatype flags decide
  // (experiments/utils.py:549-555), independently of the model's
  const int32_t* aat = feats->aatype_bb_given ? feats->aatype_bb : (ctx->cfg.with_aatype ? feats->aatype : nullptr);
  // One timestep.  Per-step slices (noise, schedule row, trajectory slots) are resolved on the device from the step counter, so the
  // same enqueue sequence -- and therefore one captured CUDA graph -- serves every step that does a reverse update.
  auto enqueue_step = [&](bool last) -> int {
    RET(set_step());
    RET(forward_impl(ctx, B, N, &f, &o, st));
    if (!last) {
      if (update_sc) {  // experiments/utils.py:356-358
        copy_trans_kernel<<<gM3, 256, 0, st>>>((int)M, w.pred_rigids, w.sc_ca);
        LAUNCH_CHECK();
      }
      ReverseArgs a;
      a.N = N; a.rigids_t = w.rig_cur; a.rot_score = w.rot_score; a.trans_score = w.trans_score; a.dmask = w.dmask;
      a.z_rot = nullptr; a.z_trans = nullptr; a.noise_half = M * 3;
      a.use_philox = use_philox; a.philox_seed = philox_seed;
      a.noise_ref.step = w.step_dev; a.noise_ref.stride = 2 * M * 3; a.noise_ref.base = w.call_ptrs + 0;
      a.sched = w.sched_dev; a.sched_ref.step = w.step_dev; a.sched_ref.stride = FDPT_SCHED_COLS;
      a.center = center; a.diffuse_rot = diffuse_rot; a.diffuse_trans = diffuse_trans;
      a.cs = ctx->cfg.r3_coordinate_scaling; a.rigids_out = w.rig_next;
      reverse_kernel<<<B, 256, 0, st>>>(a);
      LAUNCH_CHECK();
    } else {
      CK(cudaMemcpyAsync(w.rig_next, w.pred_rigids, sizeof(float) * M * 7, cudaMemcpyDeviceToDevice, st));
    }
    // trajectory slot T-1-step (index 0 = final sample); final_only: only the last step writes, into slot 0
    const bool write = !fo || last;
    StepRef slot;
    if (!fo) {
      slot.step = w.step_dev; slot.T = w.step_dev + 1; slot.reversed = 1;
    }
    auto ref = [&](int ptr_slot, long long stride) {
      StepRef r = slot;
      r.stride = stride;
      r.base = w.call_ptrs + ptr_slot;
      return r;
    };
    // backbone of x_{t-1} and of the x0 prediction (experiments/utils.py:396-410); always computed, like the reference
    const bool wp = write && out->prot_traj, w0 = write && out->rigid_0_traj;
    backbone_kernel<<<gM, 128, 0, st>>>((int)M, w.rig_next, w.psi, aat, ctx->ideal, ctx->psi_frame, ctx->atom_mask, w.bb_tmp,
                                        wp ? ref(1, M * 15) : StepRef());
    LAUNCH_CHECK();
    backbone_kernel<<<gM, 128, 0, st>>>((int)M, w.pred_rigids, w.psi, aat, ctx->ideal, ctx->psi_frame, ctx->atom_mask, w.bb_tmp,
                                        w0 ? ref(2, M * 15) : StepRef());
    LAUNCH_CHECK();
    if (write && out->trans_traj) {
      trans0_kernel<<<gM, 128, 0, st>>>((int)M, w.pred_rigids, w.rig_next, feats->res_mask, feats->fixed_mask, nullptr, ref(3, M * 3));
      LAUNCH_CHECK();
    }
    if (write && out->rigid_traj) {
      copy_slot_kernel<<<gM7, 256, 0, st>>>(M * 7, w.rig_next, nullptr, ref(4, M * 7));
      LAUNCH_CHECK();
    }
    advance_state_kernel<<<gM7, 256, 0, st>>>(M * 7, w.rig_next, w.rig_cur, w.step_dev);
    LAUNCH_CHECK();
    return FDPT_OK;
  };
  // graph key: everything the captured step bakes in
  std::vector<unsigned char> key;
  auto put = [&](const void* p, size_t n) { key.insert(key.end(), (const unsigned char*)p, (const unsigned char*)p + n); };
  {
    const int ints[12] = {B, N, fo, center, diffuse_rot, diffuse_trans, ctx->gemm_tc, out->prot_traj != nullptr, out->rigid_0_traj != nullptr,
                          out->trans_traj != nullptr, out->rigid_traj != nullptr, use_philox};
    put(ints, sizeof(ints));
    put(&philox_seed, sizeof(philox_seed));
    put(&self_condition, sizeof(self_condition));
    put(&aat, sizeof(aat));
    put(&ctx->score_table, sizeof(ctx->score_table));
    put(&f, sizeof(f));
    put(&st, sizeof(st));
    put(&w.base, sizeof(w.base));
  }
  int n_graph_steps = 0;
  for (int s = 0; s < T; ++s) n_graph_steps += sched[s * FDPT_SCHED_COLS + FDPT_SCHED_IS_LAST] == 0.0;
  const bool use_graph = ctx->use_graph && !ctx->prof_on && n_graph_steps >= 2;
  ctx->progress_used = 0;
  ctx->progress_steps = T;
  auto mark_progress = [&](int s) -> int {  // event after step s when it closes a chunk (or the call)
    if (ctx->progress_chunk <= 0 || ((s + 1) % ctx->progress_chunk != 0 && s != T - 1)) return FDPT_OK;
    if (ctx->progress_used == (int)ctx->progress_ev.size()) {
      cudaEvent_t e;
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->progress_ev.push_back(e);
    }
    CK(cudaEventRecord(ctx->progress_ev[ctx->progress_used++], st));
    return FDPT_OK;
  };
  for (int s = 0; s < T; ++s) {
    const bool last = sched[s * FDPT_SCHED_COLS + FDPT_SCHED_IS_LAST] != 0.0;  // !(t > min_t)
    if (last || !use_graph) {
      RET(enqueue_step(last));
      RET(mark_progress(s));
      continue;
    }
    auto& sg = ctx->step_graph;
    if (!sg.exec || sg.key != key) {
      if (sg.exec) {
        cudaGraphExecDestroy(sg.exec);
        sg.exec = nullptr;
      }
      const int64_t l0 = ctx->launches;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      const int rc = enqueue_step(false);
      cudaGraph_t g = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(st, &g);
      if (rc != FDPT_OK) {
        if (g) cudaGraphDestroy(g);
        return rc;
      }
      if (ce != cudaSuccess) return fail(ctx, FDPT_ERR_CUDA, "graph capture of a timestep failed: %s", cudaGetErrorString(ce));
      sg.launches = ctx->launches - l0;
      ctx->launches = l0;
      const cudaError_t ie = cudaGraphInstantiate(&sg.exec, g, 0);
      cudaGraphDestroy(g);
      if (ie != cudaSuccess) return fail(ctx, FDPT_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(ie));
      sg.key = key;
      ctx->stat_captures++;
    }
    CK(cudaGraphLaunch(sg.exec, st));
    ctx->launches += sg.launches;
    RET(mark_progress(s));
  }
  if (out->psi_pred) CK(cudaMemcpyAsync(out->psi_pred, w.psi, sizeof(float) * M * 2, cudaMemcpyDeviceToDevice, st));
  if (fenced) {
    CK(cudaEventRecord(ctx->fence_out, st));
    CK(cudaStreamWaitEvent(user_st, ctx->fence_out, 0));
  }
  return FDPT_OK;
}

int fdpt_set_progress_chunk(fdpt_ctx* ctx, int chunk_steps) {
  if (!ctx || chunk_steps < 0) return FDPT_ERR_INVALID;
  ctx->progress_chunk = chunk_steps;
  return FDPT_OK;
}

int fdpt_wait_step(fdpt_ctx* ctx, int step) {
  if (!ctx) return FDPT_ERR_INVALID;
  GUARD(ctx);
  if (ctx->progress_chunk <= 0 || step < 0 || step >= ctx->progress_steps) return fail(ctx, FDPT_ERR_STATE, "no progress event covers step %d", step);
  int idx = step / ctx->progress_chunk;  // the chunk's closing event (the last chunk may be short: its event is the last one recorded)
  if (idx >= ctx->progress_used) idx = ctx->progress_used - 1;
  if (idx < 0) return fail(ctx, FDPT_ERR_STATE, "no fdpt_sample call recorded progress events");
  CK(cudaEventSynchronize(ctx->progress_ev[idx]));
  return FDPT_OK;
}

int fdpt_set_score_table(fdpt_ctx* ctx, const double* score_norms, int num_sigma, int num_omega, const double* omega_bounds) {
  if (!ctx) return FDPT_ERR_INVALID;
  GUARD(ctx);
  CK(cudaDeviceSynchronize());
  cudaFree(ctx->score_table);
  cudaFree(ctx->omega_bounds);
  ctx->score_table = ctx->omega_bounds = nullptr;
  ctx->tab_sigma = ctx->tab_omega = 0;
  ctx->step_graph.key.clear();
  if (!score_norms) return FDPT_OK;
  if (!omega_bounds || num_sigma <= 0 || num_omega <= 1) return FDPT_ERR_INVALID;
  CK(cudaMalloc(&ctx->score_table, sizeof(double) * (size_t)num_sigma * num_omega));
  CK(cudaMalloc(&ctx->omega_bounds, sizeof(double) * (size_t)(num_omega - 1)));
  CK(cudaMemcpy(ctx->score_table, score_norms, sizeof(double) * (size_t)num_sigma * num_omega, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->omega_bounds, omega_bounds, sizeof(double) * (size_t)(num_omega - 1), cudaMemcpyHostToDevice));
  ctx->tab_sigma = num_sigma;
  ctx->tab_omega = num_omega;
  return FDPT_OK;
}

int fdpt_rot_score_idx(fdpt_ctx* ctx, int B, int N, const float* quats_t, const float* quats_0, const int32_t* sigma_idx, const float* mask,
                       double* out, void* stream) {
  if (!ctx || !quats_t || !quats_0 || !sigma_idx || !out) return FDPT_ERR_INVALID;
  if (!ctx->score_table) return fail(ctx, FDPT_ERR_STATE, "no cached score table installed (fdpt_set_score_table)");
  GUARD(ctx);
  const int M = B * N;
  rot_score_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(M, N, quats_t, 4, quats_0, 4, nullptr, mask, out, ctx->score_table, ctx->omega_bounds,
                                                                  ctx->tab_omega, sigma_idx);
  LAUNCH_CHECK();
  return FDPT_OK;
}

int fdpt_sample_ref(fdpt_ctx* ctx, int B, int N, const float* impute, const float* diffuse_mask, const double* cdf, const double* omega_grid,
                    int num_omega, const double* draws, uint64_t philox_seed, int diffuse_rot, int diffuse_trans, float* rigids_out,
                    void* stream) {
  if (!ctx || B <= 0 || N <= 0 || !cdf || !omega_grid || num_omega < 2 || !rigids_out) return FDPT_ERR_INVALID;
  // se3_diffuser.py:484-507: without imputation values everything must be diffused
  if (!impute && (!diffuse_rot || !diffuse_trans || diffuse_mask)) return fail(ctx, FDPT_ERR_INVALID, "Must provide imputation values");
  GUARD(ctx);
  SampleRefArgs a;
  a.N = N; a.impute = impute; a.dmask = diffuse_mask; a.cdf = cdf; a.omega_grid = omega_grid; a.num_omega = num_omega; a.draws = draws;
  a.philox_seed = philox_seed; a.diffuse_rot = diffuse_rot; a.diffuse_trans = diffuse_trans; a.cs = ctx->cfg.r3_coordinate_scaling; a.out = rigids_out;
  sample_ref_kernel<<<dim3((N + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(a);
  LAUNCH_CHECK();
  return FDPT_OK;
}

int fdpt_seq_tfmr(fdpt_ctx* ctx, int blk, int B, int N, const float* node, const float* node0, const float* mask, float* tfmr_out,
                  float* node_out, void* stream) {
  if (!ctx || blk < 0 || blk >= NBLK || !node || !node0 || !mask) return FDPT_ERR_INVALID;
  if (!ctx->finalized) return fail(ctx, FDPT_ERR_STATE, "parameters not finalised");
  GUARD(ctx);
  RET(reserve_ws(ctx, B, N));
  cudaStream_t st = (cudaStream_t)stream;
  Workspace& w = ctx->ws;
  const size_t M = (size_t)B * N;
  CK(cudaMemcpyAsync(w.node, node, sizeof(float) * M * C_S, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(w.node0, node0, sizeof(float) * M * C_S, cudaMemcpyDeviceToDevice, st));
  RET(run_seq_tfmr(ctx, blk, B, N, mask, st));
  if (tfmr_out) CK(cudaMemcpyAsync(tfmr_out, w.tf_x, sizeof(float) * M * TF_D, cudaMemcpyDeviceToDevice, st));
  if (node_out) CK(cudaMemcpyAsync(node_out, w.node, sizeof(float) * M * C_S, cudaMemcpyDeviceToDevice, st));
  return FDPT_OK;
}

// ---- PDB text of one backbone model (host-side; "next" row (f)(2) of SURVEY §8) ------------------------------------------
// Restates framedipt/protein/protein.py:165-279 (to_pdb) for the 5 backbone slots the sampler produces (atom37 slots 0..4 =
// N, CA, C, CB, O), with the atom mask of framedipt/analysis/utils.py:128-129 (sum |xyz| > 1e-7).  Byte-for-byte the reference text.
int64_t fdpt_to_pdb(const float* atom37_bb, const int32_t* aatype, const int32_t* residue_index, const int32_t* chain_index,
                    const float* b_factors, int n_res, int model, int add_end, char* out, int64_t cap) {
  static const char* kRes3[21] = {"ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU",
                                  "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL", "UNK"};
  static const char* kAtom[5] = {"N", "CA", "C", "CB", "O"};
  static const char kChains[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789";
  if (!atom37_bb || !out || n_res <= 0 || cap <= 0) return FDPT_ERR_INVALID;
  for (int i = 0; i < n_res; ++i) {
    if (aatype && (aatype[i] < 0 || aatype[i] > 20)) return FDPT_ERR_INVALID;        // "Invalid aatypes." / res_index range
    if (chain_index && (chain_index[i] < 0 || chain_index[i] >= 62)) return FDPT_ERR_INVALID;  // PDB_MAX_CHAINS
  }
  int64_t pos = 0;
  char line[128];
  auto emit = [&](int len) -> bool {  // line.ljust(80) + newline: pads short lines, never truncates (coordinates >= 1e4 widen a line)
    if (len < 0 || len >= (int)sizeof(line)) return false;
    const int w = len < 80 ? 80 : len;
    if (pos + w + 1 > cap) return false;
    memcpy(out + pos, line, len);
    if (len < 80) memset(out + pos + len, ' ', 80 - len);
    out[pos + w] = '\n';
    pos += w + 1;
    return true;
  };
  auto aa = [&](int i) { return aatype ? aatype[i] : 0; };
  auto ch = [&](int i) { return chain_index ? chain_index[i] : 0; };
  auto ri = [&](int i) { return residue_index ? residue_index[i] : i; };
  if (!emit(snprintf(line, sizeof(line), "MODEL     %d", model))) return FDPT_ERR_INVALID;
  int atom_index = 1, last_chain = ch(0);
  for (int i = 0; i < n_res; ++i) {
    if (last_chain != ch(i)) {
      if (!emit(snprintf(line, sizeof(line), "%-6s%5d      %3s %c%4d", "TER", atom_index, kRes3[aa(i - 1)], kChains[ch(i - 1)], ri(i - 1))))
        return FDPT_ERR_INVALID;
      last_chain = ch(i);
      ++atom_index;
    }
    for (int k = 0; k < 5; ++k) {
      const float* p = atom37_bb + ((long long)i * 5 + k) * 3;
      if (!(fabsf(p[0]) + fabsf(p[1]) + fabsf(p[2]) > 1e-7f)) continue;
      const double bf = b_factors ? (double)b_factors[(long long)i * 5 + k] : 0.0;
      const int len = snprintf(line, sizeof(line), "%-6s%5d  %-3s%1s%3s %c%4d%1s   %8.3f%8.3f%8.3f%6.2f%6.2f          %2c%2s", "ATOM", atom_index,
                               kAtom[k], "", kRes3[aa(i)], kChains[ch(i)], ri(i), "", (double)p[0], (double)p[1], (double)p[2], 1.0, bf,
                               kAtom[k][0], "");
      if (!emit(len)) return FDPT_ERR_INVALID;
      ++atom_index;
    }
  }
  if (!emit(snprintf(line, sizeof(line), "%-6s%5d      %3s %c%4d", "TER", atom_index, kRes3[aa(n_res - 1)], kChains[ch(n_res - 1)], ri(n_res - 1))))
    return FDPT_ERR_INVALID;
  if (!emit(snprintf(line, sizeof(line), "ENDMDL"))) return FDPT_ERR_INVALID;
  if (add_end && !emit(snprintf(line, sizeof(line), "END"))) return FDPT_ERR_INVALID;
  return pos;
}

// ---- unit entry points ------------------------------------------------------------------------------------
int fdpt_linear(fdpt_ctx* ctx, int M, int N, int K, const float* x, const float* w, const float* bias, int act, float* y, void* stream) {
  if (!ctx || !x || !w || !y) return FDPT_ERR_INVALID;
  GUARD(ctx);
  // unit entry: Linear layers run on lin_tc with pre-split weights, so split this weight too (re-done on every call: the caller's
  // buffer is not a registered parameter and may have changed)
  if (ctx->gemm_tc) {
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    RET(pack_linear(ctx, w, K, N, K));
    CK(cudaDeviceSynchronize());
  }
  Lin lin{ctx, (cudaStream_t)stream};
  return lin(x, K, w, K, bias, y, N, M, N, K, act);
}

int fdpt_set_option(fdpt_ctx* ctx, int option, int value) {
  if (!ctx) return FDPT_ERR_INVALID;
  GUARD(ctx);
  switch (option) {
    case FDPT_OPT_GEMM_TC: ctx->gemm_tc = value != 0; return FDPT_OK;
    case FDPT_OPT_MN_SWAP: ctx->mn_swap = value != 0; return FDPT_OK;
    case FDPT_OPT_DEBUG_FLAGS: ctx->dbg_flags = value; ctx->step_graph.key.clear(); tc::g_force_bn = (value & 8) ? 128 : 0; tc::g_use_pdl = (value & 16) ? 0 : 1; return FDPT_OK;
    case FDPT_OPT_GRAPH: ctx->use_graph = value != 0; return FDPT_OK;
    case FDPT_OPT_ET_PAIR: return fail(ctx, FDPT_ERR_INVALID, "the CTA-pair EdgeTransition variant was removed (slower than the single-CTA kernel, DESIGN.md)");
    case FDPT_OPT_ET_R2_TMEM: return value ? FDPT_OK : fail(ctx, FDPT_ERR_INVALID, "the shared-memory hand-over of r2 was removed: the EdgeTransition kernel always feeds GEMM3 from tensor memory");
    case FDPT_OPT_LIN_WRES: ctx->lin_wres = value != 0; ctx->step_graph.key.clear(); return FDPT_OK;
    case FDPT_OPT_TF_IMG: ctx->tf_img = value != 0; ctx->step_graph.key.clear(); return FDPT_OK;
    case FDPT_OPT_IPA_IMG: ctx->ipa_img = value; ctx->step_graph.key.clear(); return FDPT_OK;
    case FDPT_OPT_ET_TIMELINE:
      if (value && !ctx->et_dbg) {
        // host-mapped so that the stamps (and the barrier-timeout records of tc::mbar_wait) survive a device-side trap
        CK(cudaHostAlloc(&ctx->et_dbg, (8 * 48 + 32) * sizeof(long long), cudaHostAllocMapped));
        memset(ctx->et_dbg, 0, (8 * 48 + 32) * sizeof(long long));
        unsigned long long* fail_buf = reinterpret_cast<unsigned long long*>(ctx->et_dbg) + 8 * 48;
        CK(cudaMemcpyToSymbol(tc::g_mbar_fail_buf, &fail_buf, sizeof(fail_buf)));
      }
      return FDPT_OK;
    default: return fail(ctx, FDPT_ERR_INVALID, "unknown option %d", option);
  }
}

int fdpt_debug_read(fdpt_ctx* ctx, int64_t* out, int n) {
  if (!ctx || !out || !ctx->et_dbg || n > 8 * 48 + 32) return FDPT_ERR_INVALID;
  GUARD(ctx);
  cudaDeviceSynchronize();  // may report a sticky error after a trap: the host-mapped buffer is still readable
  memcpy(out, ctx->et_dbg, n * sizeof(long long));
  return FDPT_OK;
}

int fdpt_matmul(fdpt_ctx* ctx, int batch, int M, int N, int K, const float* a, int lda, long long sa, const float* b, int ldb, long long sb,
                int b_kmajor, float alpha, float* c, int ldc, long long sc, void* stream) {
  if (!ctx || !a || !b || !c || batch <= 0) return FDPT_ERR_INVALID;
  GUARD(ctx);
  GemmArgs g;
  g.A = a; g.lda = lda; g.sA1 = sa; g.B = b; g.ldb = ldb; g.sB1 = sb; g.C = c; g.ldc = ldc; g.sC1 = sc;
  g.M = M; g.N = N; g.K = K; g.alpha = alpha;
  CK(gemm_dispatch(ctx, g, b_kmajor != 0, batch, (cudaStream_t)stream));
  ctx->launches++;
  return FDPT_OK;
}

int fdpt_bench_linear(fdpt_ctx* ctx, int M, int N, int K, const float* x, const float* w, const float* bias, float* y, int reps,
                      float* ms_per_call) {
  if (!ctx || !x || !w || !y || !ms_per_call || reps <= 0) return FDPT_ERR_INVALID;
  GUARD(ctx);
  CK(cudaDeviceSynchronize());
  RET(pack_linear(ctx, w, K, N, K));
  Lin lin{ctx, nullptr};
  for (int i = 0; i < 3; ++i) RET(lin(x, K, w, K, bias, y, N, M, N, K, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, nullptr));
  for (int i = 0; i < reps; ++i) RET(lin(x, K, w, K, bias, y, N, M, N, K, 0));
  CK(cudaEventRecord(e1, nullptr));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  *ms_per_call = ms / reps;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return FDPT_OK;
}

int fdpt_tmem_a_selftest(fdpt_ctx* ctx, const float* a, const float* b, float* d, void* stream) {
  if (!ctx || !a || !b || !d) return FDPT_ERR_INVALID;
  GUARD(ctx);
  tc::tmem_a_selftest_kernel<<<1, 128, 1024 + 16384 + 64, (cudaStream_t)stream>>>(a, b, d);
  LAUNCH_CHECK();
  return FDPT_OK;
}

int fdpt_tc_linear(fdpt_ctx* ctx, int M, int N, int K, const float* x, const float* w, const float* bias, int act, float* y,
                   void* stream) {
  if (!ctx || !x || !w || !y || M <= 0 || N % 128 || K % 64 || K > 512 || K <= 0 || N <= 0) return FDPT_ERR_INVALID;
  GUARD(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  __half* img = nullptr;
  CK(cudaMalloc(&img, (size_t)N * K * sizeof(__half)));
  const long long chunks = (long long)N * (K / 8);
  tc::pack_weight_image_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(w, K, N, K, K, img);
  LAUNCH_CHECK();
  tc::TcLinearArgs a{x, K, M, K, img, N, bias, act, y, N};
  const size_t smem = tc::tc_linear_smem_bytes(K);
  CK(cudaFuncSetAttribute(tc::tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc::tc_linear_kernel<<<(M + 127) / 128, 192, smem, st>>>(a);
  LAUNCH_CHECK();
  CK(cudaStreamSynchronize(st));
  CK(cudaFree(img));
  return FDPT_OK;
}

int fdpt_ipa(fdpt_ctx* ctx, int blk, int B, int N, const float* s, const float* z, const float* quats, const float* trans,
             const float* mask, float* out, void* stream) {
  if (!ctx || blk < 0 || blk >= NBLK || !s || !z || !quats || !trans || !mask || !out) return FDPT_ERR_INVALID;
  if (!ctx->finalized) return fail(ctx, FDPT_ERR_STATE, "parameters not finalised");
  GUARD(ctx);
  RET(reserve_ws(ctx, B, N));
  RET(z_fp32_to_image(ctx, B, N, z, (cudaStream_t)stream));
  return run_ipa(ctx, blk, B, N, s, ctx->ws.z, quats, trans, mask, out, C_S, nullptr, nullptr, (cudaStream_t)stream);
}

int fdpt_edge_transition(fdpt_ctx* ctx, int blk, int B, int N, const float* node, const float* z_in, const float* mask, float* z_out,
                         void* stream) {
  if (!ctx || blk < 0 || blk >= NBLK - 1 || !node || !z_in || !mask || !z_out) return FDPT_ERR_INVALID;
  if (!ctx->finalized) return fail(ctx, FDPT_ERR_STATE, "parameters not finalised");
  GUARD(ctx);
  RET(reserve_ws(ctx, B, N));
  RET(z_fp32_to_image(ctx, B, N, z_in, (cudaStream_t)stream));
  RET(run_edge_transition(ctx, blk, B, N, node, ctx->ws.z, mask, ctx->ws.z, (cudaStream_t)stream));
  return z_image_to_fp32(ctx, B, N, z_out, (cudaStream_t)stream);
}

int fdpt_embed(fdpt_ctx* ctx, int B, int N, const fdpt_feats* in, float* node_out, float* edge_out, void* stream) {
  if (!ctx || !in || !node_out || !edge_out) return FDPT_ERR_INVALID;
  if (!ctx->finalized) return fail(ctx, FDPT_ERR_STATE, "parameters not finalised");
  GUARD(ctx);
  RET(reserve_ws(ctx, B, N));
  RET(run_embed(ctx, B, N, in, node_out, ctx->ws.z, (cudaStream_t)stream));
  return z_image_to_fp32(ctx, B, N, edge_out, (cudaStream_t)stream);
}

}  // extern "C"
