/*
 * fdpt.h — C ABI of libfdpt.so, the B200 (sm_100a) sampler hot path of framedipt_b200.
 *
 * The reference (instadeepai/FrameDiPT) has no FFI: its boundary for this path is a Python call
 * surface.  Each entry point below names the reference interface it replaces (paths relative to
 * the reference repository root).  The Python shims in framedipt_b200/ bind these with ctypes
 * (see INTEGRATION.md); no torch types cross this boundary — only raw pointers and sizes.
 *
 * Conventions
 *   - one context per GPU, not thread-safe (one host thread per context);
 *   - all tensor pointers are DEVICE pointers to contiguous row-major arrays unless stated;
 *   - every call enqueues on the caller's `stream` (a cudaStream_t passed as void*) and returns
 *     0 on success or a negative fdpt_status; fdpt_last_error(ctx) gives the message;
 *   - no exception crosses the ABI;  dtypes: f32 unless stated, i32 indices, f64 where the
 *     reference computes in float64 (IGSO(3) score, reverse SDE step).
 */
#ifndef FDPT_H
#define FDPT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fdpt_ctx fdpt_ctx;

typedef enum {
  FDPT_OK = 0,
  FDPT_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  FDPT_ERR_CUDA = -2,      /* CUDA runtime error */
  FDPT_ERR_PARAM = -3,     /* unknown / missing / mis-shaped parameter */
  FDPT_ERR_STATE = -4      /* call order (params not finalised, no schedule ...) */
} fdpt_status;

/* Model + diffuser hyper-parameters: config/base.yaml:33-79 of the reference
 * (model_conf.*, model_conf.embed.*, model_conf.ipa.*, diffuser.r3.*, diffuser.so3.*). */
typedef struct {
  int32_t c_s, c_z, c_hidden, c_skip;
  int32_t no_heads, no_qk_points, no_v_points, num_blocks;
  int32_t index_embed_size, num_bins;
  float min_bin, max_bin;
  int32_t seq_tfmr_num_heads, seq_tfmr_num_layers;
  float coordinate_scaling;
  int32_t with_aatype;      /* 1: node features carry the 21-way aatype one-hot (inpainting / input_aatype) */
  double r3_min_b, r3_max_b;
  float r3_coordinate_scaling;     /* diffuser.r3.coordinate_scaling: used by the translation score and the reverse step
                                      (r3_diffuser.py:30, 344-440); `coordinate_scaling` above is model.ipa's (ipa_pytorch.py:472) */
  int32_t embed_self_conditioning; /* model.embed.embed_self_conditioning (score_network.py:95-96, 185): 0 = no distogram features */
} fdpt_config;

/* Per-step schedule row (doubles), filled on the host with the reference's numpy expressions
 * (so3_diffuser.py:288-323, r3_diffuser.py:48-96, experiments/utils.py:166-190). */
enum {
  FDPT_SCHED_T32 = 0,        /* t rounded through float32 (feats["t"]) */
  FDPT_SCHED_SIGMA = 1,      /* discrete_sigma[t_to_idx(t)] */
  FDPT_SCHED_SO3_G2DT = 2,   /* g(t)^2 * dt */
  FDPT_SCHED_SO3_NOISE = 3,  /* g(t) * sqrt(dt) * noise_scale */
  FDPT_SCHED_R3_BT = 4,      /* b_t */
  FDPT_SCHED_DT = 5,         /* dt */
  FDPT_SCHED_R3_NOISE = 6,   /* sqrt(b_t) * sqrt(dt) * noise_scale */
  FDPT_SCHED_IS_LAST = 7,    /* 1.0 when !(t > min_t): take the x0 prediction instead of a reverse step */
  FDPT_SCHED_SIGMA_IDX = 8,  /* t_to_idx(t): row of the cached score table (so3.use_cached_score=True, so3_diffuser.py:389-396) */
  FDPT_SCHED_COLS = 10
};

/* Input features of one forward = the feature dict of experiments/sampler.py:69-111, 267-354. */
typedef struct {
  const float* rigids_t;     /* [B,N,7] quat (w,x,y,z) + translation in Angstrom */
  const float* sc_ca_t;      /* [B,N,3] self-conditioning CA (Angstrom) */
  const float* res_mask;     /* [B,N] */
  const float* fixed_mask;   /* [B,N] */
  const int32_t* seq_idx;    /* [B,N] */
  const int32_t* aatype;     /* [B,N] after preprocess_aatype (framedipt/data/utils.py:565-610); NULL if with_aatype=0 */
  const float* gt_psi;       /* [B,N,2] torsion_angles_sin_cos[...,2,:] */
  const float* idx_emb;      /* [B,N,E] get_index_embedding(seq_idx), host-evaluated (score_network.py:17-38) */
  const float* rel_emb;      /* [R,E]   get_index_embedding(r) for r = rel_min .. rel_min+R-1 */
  int32_t rel_min, rel_count;
  const float* t_emb;        /* [B,E]   get_timestep_embedding(t)    (score_network.py:41-64), host-evaluated */
  const float* t_emb_eps;    /* [E]     get_timestep_embedding(1e-5) */
  const float* t32;          /* [B]     feats["t"] */
  const double* sigma;       /* [B]     discrete_sigma[t_to_idx(t)] */
  const int32_t* sigma_idx;  /* [B]     t_to_idx(t); only read when a cached score table is installed (fdpt_set_score_table) */
  const int32_t* aatype_bb;  /* [B,N]   residue types used for the backbone atoms of fdpt_sample's trajectories =
                                        preprocess_aatype(aatype, fixed_mask, inpainting, input_aatype) of the inference_fn CALL
                                        (experiments/utils.py:549-555), which may differ from the model's flags; NULL = ALA frames */
  int32_t aatype_bb_given;   /* 0: fall back to `aatype` (when the model takes aatype) */
} fdpt_feats;

/* Outputs of one forward = ScoreNetwork.forward's dict (score_network.py:262-275). Any pointer may be NULL. */
typedef struct {
  float* rigids;       /* [B,N,7] predicted x0 frames (quat + trans, Angstrom) */
  double* rot_score;   /* [B,N,3] float64 like the reference */
  float* trans_score;  /* [B,N,3] */
  float* psi;          /* [B,N,2] */
  float* atom37_bb;    /* [B,N,5,3] atom37 slots 0..4 (N,CA,C,CB,O) of the predicted frames; other 32 slots are zero */
} fdpt_out;

/* Trajectory buffers of fdpt_sample = the dict returned by inference_fn (experiments/utils.py:610-626).
 * Index 0 of each trajectory is the FINAL sample (the reference flips).  Any pointer may be NULL. */
typedef struct {
  float* prot_traj;     /* [T,B,N,5,3]  backbone of x_{t-1} (atom37 slots 0..4) */
  float* rigid_traj;    /* [T+1,B,N,7]  */
  float* trans_traj;    /* [T,B,N,3]    */
  float* rigid_0_traj;  /* [T,B,N,5,3]  backbone of the x0 prediction */
  float* psi_pred;      /* [B,N,2]      last step */
  int32_t final_only;   /* 1: only slot 0 (final sample) of prot_traj / rigid_traj / ... is written ([1,...] buffers) */
} fdpt_traj;

/* ---- lifetime -------------------------------------------------------------------------------- */
/* replaces: ScoreNetwork.__init__ / SE3Diffuser.__init__ (score_network.py:200-216, se3_diffuser.py:39-49) */
int fdpt_create(const fdpt_config* cfg, int device, fdpt_ctx** out);
int fdpt_destroy(fdpt_ctx* ctx);
const char* fdpt_last_error(const fdpt_ctx* ctx);
const char* fdpt_version(void);

/* replaces: nn.Module.load_state_dict (experiments/inference.py:149-161).  `data` may be a host or device
 * pointer to float32; key/shape are those of the reference state_dict. */
int fdpt_load_param(fdpt_ctx* ctx, const char* key, const float* data, const int64_t* shape, int ndim);
int fdpt_finalize_params(fdpt_ctx* ctx);
int fdpt_num_params_expected(const fdpt_ctx* ctx);

/* Reserve workspace for problems up to (B,N); optional (forward grows it lazily, which synchronises). */
int fdpt_reserve(fdpt_ctx* ctx, int B, int N);
int64_t fdpt_workspace_bytes(const fdpt_ctx* ctx);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* replaces: ScoreNetwork.forward (score_network.py:218-275) incl. Embedder (129-197), IpaScore.forward
 * (ipa_pytorch.py:509-572), calc_rot_score / calc_trans_score (se3_diffuser.py:269-292). */
int fdpt_forward(fdpt_ctx* ctx, int B, int N, const fdpt_feats* in, const fdpt_out* out, void* stream);

/* replaces: SE3Diffuser.reverse (se3_diffuser.py:346-401) + Rigid.to_tensor_7 (rigid_utils.py:1200-1212).
 * sched_row: host pointer to FDPT_SCHED_COLS doubles.  z_rot/z_trans: N(0,1) draws [B,N,3] float64 (device).
 * rigids_out may alias rigids_t. */
int fdpt_reverse(fdpt_ctx* ctx, int B, int N, const float* rigids_t, const double* rot_score, const float* trans_score,
                 const float* diffuse_mask, const double* z_rot, const double* z_trans, const double* sched_row,
                 int center, int diffuse_rot, int diffuse_trans, float* rigids_out, void* stream);

/* replaces: all_atom.compute_backbone (framedipt/protein/all_atom.py:147-176) as used by
 * get_atom_positions_from_rigids (experiments/utils.py:415-438).  aatype may be NULL (ALA). out [B,N,5,3]. */
int fdpt_backbone(fdpt_ctx* ctx, int B, int N, const float* rigids, const float* psi, const int32_t* aatype,
                  float* atom37_bb, void* stream);

/* replaces: SE3Diffuser.calc_rot_score / calc_trans_score as standalone calls */
int fdpt_rot_score(fdpt_ctx* ctx, int B, int N, const float* quats_t, const float* quats_0, const double* sigma,
                   const float* mask, double* out, void* stream);
int fdpt_trans_score(fdpt_ctx* ctx, int B, int N, const float* trans_t, const float* trans_0, const float* t32,
                     const float* mask, int scale, float* out, void* stream);

/* replaces: experiments.utils.inference_fn's loop (experiments/utils.py:557-626): one self-conditioning forward
 * at the first t, then num_t x (forward -> reverse | take x0 at the last step -> backbone).
 *   sched      host [num_t, FDPT_SCHED_COLS] doubles, step 0 = t=1.0 ... last = min_t
 *   t_emb_tab  device [num_t, E] timestep embeddings per step; feats->t_emb / t32 / sigma are ignored
 *   noise      device float64 [num_t-1, 2, B, N, 3] standard normals (rot then trans), reference draw order (parity mode: the
 *              reference's legacy numpy stream, drawn on the host); NULL = throughput mode: the reverse step draws its normals on
 *              the device from a counter-based Philox4x32-10 generator keyed by `philox_seed` (counter = step, residue, component)
 *   self_condition  bit 0: run the self-conditioning pre-pass at the first t (experiments/utils.py:571-578);
 *                   bit 1: never update sc_ca_t inside the loop (inference_fn(embed_self_conditioning=False), utils.py:356-358)
 *   feats->rigids_t is read as x_T and not modified; feats->sc_ca_t is the initial self-conditioning (zeros). */
int fdpt_sample(fdpt_ctx* ctx, int B, int N, const fdpt_feats* feats, int num_t, const double* sched,
                const float* t_emb_tab, const double* noise, uint64_t philox_seed, int self_condition, int center,
                int diffuse_rot, int diffuse_trans, const fdpt_traj* out, void* stream);

/* Streaming read-back of fdpt_sample's trajectories: with a chunk size c > 0 the next fdpt_sample calls record an event after every
 * c timesteps (and after the last one); fdpt_wait_step(ctx, s) blocks the calling host thread until timestep s (0-based) of the most
 * recent fdpt_sample call has completed on the device, so the caller can copy finished trajectory slots to the host while later steps
 * are still running (inference_fn's [T,B,N,37,3] arrays, experiments/utils.py:610-626).  c = 0 turns the events off. */
int fdpt_set_progress_chunk(fdpt_ctx* ctx, int chunk_steps);
int fdpt_wait_step(fdpt_ctx* ctx, int step);

/* replaces: SO3Diffuser.torch_score's table look-up when so3.use_cached_score=True (so3_diffuser.py:389-396): installs the
 * [num_sigma, num_omega] float64 score-norm table (host pointer, copied) and the omega grid boundaries discrete_omega[:-1]
 * ([num_omega-1] float64, host); afterwards fdpt_forward / fdpt_sample / fdpt_rot_score_idx look the score norm up
 * (row = sigma_idx, column = torch.bucketize(omega, boundaries)) instead of evaluating the series.  NULL table = back to the series. */
int fdpt_set_score_table(fdpt_ctx* ctx, const double* score_norms, int num_sigma, int num_omega, const double* omega_bounds);
int fdpt_rot_score_idx(fdpt_ctx* ctx, int B, int N, const float* quats_t, const float* quats_0, const int32_t* sigma_idx,
                       const float* mask, double* out, void* stream);

/* replaces: SE3Diffuser.sample_ref for B samples of ONE structure (se3_diffuser.py:455-529; SO3Diffuser.sample so3_diffuser.py:325-357,
 * R3Diffuser.sample_stationary_distribution r3_diffuser.py:294-331), SURVEY §8(f3): x_T built on the device instead of B host calls.
 *   impute [N,7] ground-truth frames (quat + trans, Angstrom) or NULL (de novo: identity / zeros), diffuse_mask [N] float or NULL (all diffused)
 *   cdf / omega_grid: device float64 [num_omega]: the row _cdf[t_to_idx(1.0)] and discrete_omega (inverse-CDF by np.interp's rule)
 *   draws: device float64 [B][7N] = per sample randn [N,3], rand [N], normal [N,3] in the reference's draw order (parity mode; the
 *          k-th diffused residue takes row k of the last block), or NULL = Philox keyed by philox_seed (throughput mode)
 *   out [B,N,7]. */
int fdpt_sample_ref(fdpt_ctx* ctx, int B, int N, const float* impute, const float* diffuse_mask, const double* cdf,
                    const double* omega_grid, int num_omega, const double* draws, uint64_t philox_seed, int diffuse_rot,
                    int diffuse_trans, float* rigids_out, void* stream);

/* replaces: framedipt.protein.protein.to_pdb (protein.py:165-279) as called by analysis.utils.write_prot_to_pdb (utils.py:78-156)
 * for one model of backbone atoms.  HOST function (no GPU work): atom37_bb [n_res,5,3] host floats (atom37 slots 0..4 = N,CA,C,CB,O;
 * all-zero atoms are masked out like the reference's atom37_mask), aatype / residue_index / chain_index [n_res] (NULL: ALA / 0..n-1 /
 * chain 0), b_factors [n_res,5] or NULL.  Writes the PDB text (80-column lines) into out[cap]; returns the byte count or < 0. */
int64_t fdpt_to_pdb(const float* atom37_bb, const int32_t* aatype, const int32_t* residue_index, const int32_t* chain_index,
                    const float* b_factors, int n_res, int model, int add_end, char* out, int64_t cap);

/* number of kernel launches enqueued by this context since creation (bench.py's gpu_launches) */
int64_t fdpt_launch_count(const fdpt_ctx* ctx);
/* Diagnostics counters (bench.py reports them next to the timed region). */
enum { FDPT_STAT_GRAPH_CAPTURES = 0 /* per-timestep CUDA graphs captured so far */,
       FDPT_STAT_SAMPLE_HOST_US = 1 /* host microseconds the last fdpt_sample call spent before returning */ };
int64_t fdpt_stat(const fdpt_ctx* ctx, int which);

/* Live per-kernel timing with CUDA events on the launching stream (bench.py's roofline numbers).
 * When enabled, every launch of the listed kernels / kernel groups inside fdpt_forward / fdpt_sample is bracketed
 * by an event pair; fdpt_profile_read synchronises the device, sums the elapsed times per slot and clears them. */
enum {
  FDPT_PROF_IPA_CORE = 0,         /* ipa_core kernel alone (the HBM-bound point-attention kernel) */
  FDPT_PROF_EDGE_TRANSITION = 1,  /* whole EdgeTransition (all its kernels) */
  FDPT_PROF_EDGE_EMBED = 2,       /* whole Embedder */
  FDPT_PROF_IPA_TOTAL = 3,        /* whole IPA incl. projections and linear_out */
  FDPT_PROF_SEQ_TFMR = 4,
  FDPT_PROF_FORWARD = 5,          /* whole forward */
  FDPT_PROF_IPA_ATTN = 6,         /* every kernel that implements the IPA attention: frames on points, Q.K^T, bias + softmax + o_pair, A.V,
                                     inverse frames + norms (everything between the projection GEMM and linear_out) */
  FDPT_PROF_SLOTS = 8
};
int fdpt_profile_enable(fdpt_ctx* ctx, int on);
int fdpt_profile_read(fdpt_ctx* ctx, int slot, int* count, double* total_ms);

/* ---- unit entry points (parity tests per kernel; same kernels the hot path launches) ---------- */
/* y[M,N] = act(x[M,K] @ w[N,K]^T + bias) ; act: 0 none, 1 relu */
int fdpt_linear(fdpt_ctx* ctx, int M, int N, int K, const float* x, const float* w, const float* bias, int act,
                float* y, void* stream);
/* c[b] = alpha * a[b] @ op(b[b]) for b < batch (strides sa/sb/sc in elements); b_kmajor=1: b is [N,K] (like a weight),
 * 0: b is [K,N] row-major.  Runs the node-side GEMM kernel the hot path uses: tcgen05 3-term split TF32 (fp32-class),
 * or the SIMT fp32 kernel for N < 16 / when FDPT_OPT_GEMM_TC is 0. */
int fdpt_matmul(fdpt_ctx* ctx, int batch, int M, int N, int K, const float* a, int lda, long long sa, const float* b,
                int ldb, long long sb, int b_kmajor, float alpha, float* c, int ldc, long long sc, void* stream);
/* times `reps` back-to-back launches of the Linear kernel (weights split once) with CUDA events: micro-benchmark aid */
int fdpt_bench_linear(fdpt_ctx* ctx, int M, int N, int K, const float* x, const float* w, const float* bias, float* y, int reps,
                      float* ms_per_call);
/* bring-up / A-B switches (not needed by integrators).
 * FDPT_OPT_DEBUG_FLAGS is a bit set used by the profiling tools under tools/:
 *      1  lin_tc: skip the epilogue stores            2  lin_tc: skip the MMAs
 *      4  route the clock64 timeline buffer to the IPA core kernel instead of EdgeTransition
 *      8  gemm_tc: 128-column tiles whenever N > 128  16  launch the GEMM kernels without programmatic dependent launch
 *    256  IPA linear_out as one GEMM instead of split-K + fused reduce
 *    512  lin_tc: write the clock64 timeline of CTA (0,0)   1024  lin_tc: 32-column epilogue staging passes
 *   8192  edge embedder: write the clock64 timeline of worker thread 0 of CTA 0
 *  16384  IPA linear_out: 2-way instead of 3-way split-K
 *  32768  no operand-image chaining between consecutive Linear layers (transformer FFN, node transition)
 *  65536  IPA core: the two-pass kernel (z rows fetched twice, two CTAs per SM) also for N <= 384 (A/B switch and test hook)
 * 1 << 19  clock64 stamps inside the weight-resident Linear kernel (tools/lin_timeline.py with LT_FLAG=524288)
 * 15 << 20 TIMING EXPERIMENTS of the fused EdgeTransition kernel, RESULTS ARE WRONG while any of them is set (tools/gpu_et_exp.sh):
 *           1 << 20 no weight bulk copies, 2 << 20 MMAs shrunk to N = 16, 4 << 20 no shared-memory stores in the epilogues,
 *           8 << 20 the timeline samples CTA 77, tiles 40-47 instead of CTA 0, tiles 0-7 (this one leaves the results intact) */
enum { FDPT_OPT_GEMM_TC = 0, FDPT_OPT_MN_SWAP = 1, FDPT_OPT_ET_TIMELINE = 2, FDPT_OPT_DEBUG_FLAGS = 3, FDPT_OPT_GRAPH = 4 /* 1 (default): replay one captured CUDA graph per timestep in fdpt_sample */,
       FDPT_OPT_ET_PAIR = 5 /* retired: the cta_group::2 EdgeTransition variant of round 1 was slower (1.3 vs 1.0 ms) and has been removed */,
       FDPT_OPT_ET_R2_TMEM = 10 /* retired switch: the fused EdgeTransition kernel always hands r2 to its third GEMM through tensor memory
                                   (tcgen05.st in place over D2, tcgen05.mma with the A operand in TMEM); 0 is rejected */,
       FDPT_OPT_LIN_WRES = 9 /* 1 (default): Linear layers whose CTAs own a single n-tile run the weight-resident kernel (lin_tcw.cuh: whole
                                weight panel prefetched under the predecessor's tail, activation streamed); 0: lin_tc.  A/B switch */,
       FDPT_OPT_TF_IMG = 8 /* 1 (default): the sequence transformer's attention GEMMs multiply operand images written by the in_proj
                              epilogue / the row softmax (gemm_img.cuh); 0: fp32 operands split on the fly (gemm_tc.cuh).  A/B switch */,
       FDPT_OPT_IPA_IMG = 7 /* 1 (default): the two batched attention GEMMs of the IPA multiply ready operand images (gemm_img.cuh)
                               written by the projection GEMM's own epilogue, and linear_out multiplies the concat row as an image
                               (ipa_opt_img_kernel); 3: same without the linear_out part; 2: images written by a separate prep kernel;
                               0: fp32 operands split on the fly (gemm_tc.cuh).  A/B switch */ };
int fdpt_set_option(fdpt_ctx* ctx, int option, int value);
/* clock64 timeline of CTA 0 of the last EdgeTransition kernel ([tile][48] stamps; profiling aid, needs FDPT_OPT_ET_TIMELINE) */
int fdpt_debug_read(fdpt_ctx* ctx, int64_t* out, int n);
/* same contract on the tcgen05 tensor cores (fp16 operands, fp32 accumulate): K multiple of 64 (<= 512), N multiple of 128.
 * Bring-up / unit entry of the building blocks the fused pair-side kernels use. */
int fdpt_tc_linear(fdpt_ctx* ctx, int M, int N, int K, const float* x, const float* w, const float* bias, int act,
                   float* y, void* stream);
/* bring-up unit of tcgen05.mma with the A operand in tensor memory (written by tcgen05.st): d[128,128] = fp16(a[128,64]) . fp16(b[128,64])^T,
 * fp32 accumulate; device pointers.  Pins the TMEM operand layout the fused EdgeTransition kernel uses for its GEMM3 partial products. */
int fdpt_tmem_a_selftest(fdpt_ctx* ctx, const float* a, const float* b, float* d, void* stream);
/* InvariantPointAttention.forward (ipa_pytorch.py:170-329) of block `blk` on given s [B,N,c_s], z [B,N,N,c_z],
 * frames (quats [B,N,4], trans in 0.1 A units [B,N,3]), mask [B,N] -> out [B,N,c_s] (linear_out applied, not masked) */
int fdpt_ipa(fdpt_ctx* ctx, int blk, int B, int N, const float* s, const float* z, const float* quats,
             const float* trans, const float* mask, float* out, void* stream);
/* EdgeTransition.forward (ipa_pytorch.py:84-102) of block `blk`, followed by *edge_mask: z_out may alias z_in */
int fdpt_edge_transition(fdpt_ctx* ctx, int blk, int B, int N, const float* node, const float* z_in,
                         const float* mask, float* z_out, void* stream);
/* The sequence-transformer sub-block of IpaScore.forward (ipa_pytorch.py:533-539) of block `blk`:
 *   x = cat[node, skip_embed(node0)] -> TransformerEncoder (2 post-norm layers) -> tfmr_out [B,N,320];
 *   node_out = node + post_tfmr(tfmr_out) [B,N,256].  Either output may be NULL. */
int fdpt_seq_tfmr(fdpt_ctx* ctx, int blk, int B, int N, const float* node, const float* node0, const float* mask,
                  float* tfmr_out, float* node_out, void* stream);
/* Embedder.forward (score_network.py:129-197) incl. the mask multiply of score_network.py:252-253 */
int fdpt_embed(fdpt_ctx* ctx, int B, int N, const fdpt_feats* in, float* node_out, float* edge_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FDPT_H */
