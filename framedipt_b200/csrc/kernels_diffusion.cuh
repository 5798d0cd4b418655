// kernels_diffusion.cuh — IGSO(3)/R^3 scores, reverse SDE step, backbone atoms.
#pragma once
#include "common.cuh"

namespace fdpt {

// Reference to a per-step slice of a buffer, resolved ON THE DEVICE, so that one captured CUDA graph of a timestep can be replayed
// for every step of every fdpt_sample call: the buffer base comes from a device-resident pointer slot (rewritten per call), the
// offset is (reversed ? T-1-step : step) * stride elements with step and T read from device memory.  Default: no indirection.
struct StepRef {
  const int* step = nullptr;
  const int* T = nullptr;
  void* const* base = nullptr;
  long long stride = 0;
  int reversed = 0;
};
template <typename P>
FDPT_DEVINL P* step_resolve(P* direct, const StepRef& r) {
  P* p = r.base ? reinterpret_cast<P*>(*r.base) : direct;
  if (r.step) {
    const int s = *r.step;
    p += (long long)(r.reversed ? (*r.T - 1 - s) : s) * r.stride;
  }
  return p;
}

// ------------------------------------------------------------------------------------------------
// Rotation score  (SE3Diffuser.calc_rot_score se3_diffuser.py:281-292; transforms.quat_to_rotvec
// transforms.py:53-69; SO3Diffuser.torch_score so3_diffuser.py:373-402; igso3_expansion 18-77; score 122-191).
// One warp per residue; the 1000-term series is split across lanes and reduced in float64.
// Mixed precision follows the reference: quaternion algebra / omega / sin,cos arguments in float32
// (torch fp32 tensors), coefficients exp(-l(l+1) sigma^2/2) and all sums in float64 (sigma is float64).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rot_score_kernel(int M, int N, const float* __restrict__ quats_t, int ldt,
                                                        const float* __restrict__ quats_0, int ld0,
                                                        const double* __restrict__ sigma_b, const float* __restrict__ mask,
                                                        double* __restrict__ out, const double* __restrict__ table = nullptr,
                                                        const double* __restrict__ bounds = nullptr, int num_omega = 0,
                                                        const int32_t* __restrict__ sigma_idx = nullptr) {
  // table != nullptr: so3.use_cached_score=True (so3_diffuser.py:389-396): the score norm is looked up in the precomputed
  // [num_sigma, num_omega] table (row t_to_idx(t), column torch.bucketize(omega, discrete_omega[:-1])) instead of the series
  __shared__ double coef_s[1000];
  if (!table) {
    const int bc = min(blockIdx.x * 8, M - 1) / N;
    const double sg = sigma_b[bc];
    const double h = 0.5 * sg * sg;
    for (int l = threadIdx.x; l < 1000; l += blockDim.x) coef_s[l] = (double)(2 * l + 1) * exp(-(double)l * (double)(l + 1) * h);
  }
  __syncthreads();
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  const int b = m / N;
  float q0[4], qt[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    q0[k] = quats_0[(long long)m * ld0 + k];
    qt[k] = quats_t[(long long)m * ldt + k];
  }
  // invert_quat (rigid_utils.py:282-286)
  const float n2 = q0[0] * q0[0] + q0[1] * q0[1] + q0[2] * q0[2] + q0[3] * q0[3];
  float qi[4] = {q0[0] / n2, -q0[1] / n2, -q0[2] / n2, -q0[3] / n2};
  float q[4];
  quat_mul(qi, qt, q);
  if (q[0] < 0.f) {
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = -q[k];
  }
  const float vn = sqrtf(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const float angle = 2.f * atan2f(vn, q[0]);
  const float a2 = angle * angle;
  const float scale = (angle <= 1e-3f) ? (2.f + a2 / 12.f + 7.f * a2 * a2 / 2880.f) : angle / sinf(angle / 2.f + 1e-6f);
  const float v[3] = {scale * q[1], scale * q[2], scale * q[3]};
  const float omega = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) + 1e-6f;
  if (table) {
    if (lane < 3) {
      // torch.bucketize(omega, bounds), right=False: number of boundaries strictly below omega (bounds has num_omega - 1 entries)
      int lo_i = 0, hi_i = num_omega - 1;
      const double om = (double)omega;
      while (lo_i < hi_i) {
        const int mid = (lo_i + hi_i) >> 1;
        if (bounds[mid] < om) lo_i = mid + 1; else hi_i = mid;
      }
      const double nrm = table[(long long)sigma_idx[b] * num_omega + lo_i];
      const double mk = mask ? (double)mask[m] : 1.0;
      out[(long long)m * 3 + lane] = nrm * (double)v[lane] / (double)omega * mk;
    }
    return;
  }
  const double sig = sigma_b[b];
  const double hs2 = 0.5 * sig * sig;
  const float lo = sinf(omega / 2.f);
  const float dlo = 0.5f * cosf(omega / 2.f);
  const float lo2 = lo * lo;
  // series coefficients (2l+1) exp(-l(l+1) sigma^2/2): they depend on the sample only (sigma is per sample), so the CTA evaluates
  // the 1000 float64 exponentials once for the sample of its first residue; a warp whose residue belongs to the next sample (CTA
  // straddling a sample boundary) evaluates its own
  const int b_cta = min(blockIdx.x * 8, M - 1) / N;
  const bool shared_coef = (b == b_cta);
  double f = 0.0, df = 0.0;
  // the two divisions by lo / lo^2 are taken out of the sums (same value up to float64 rounding of the last bit)
  for (int l = lane; l < 1000; l += 32) {
    const float lh = (float)l + 0.5f;
    const float arg = __fmul_rn(omega, lh);
    float hi, c;
    sincosf(arg, &hi, &c);
    const float dhi = lh * c;
    const double coef = shared_coef ? coef_s[l] : (double)(2 * l + 1) * exp(-(double)l * (double)(l + 1) * hs2);
    f += coef * (double)hi;
    const float num = __fsub_rn(__fmul_rn(lo, dhi), __fmul_rn(hi, dlo));
    df += coef * (double)num;
  }
  f /= (double)lo;
  df /= (double)lo2;
  f = warp_sum(f);
  df = warp_sum(df);
  if (lane < 3) {
    const double sc = df / (f + 1e-4);
    const double mk = mask ? (double)mask[m] : 1.0;
    out[(long long)m * 3 + lane] = sc * (double)v[lane] / (double)omega * mk;
  }
}

// ------------------------------------------------------------------------------------------------
// Translation score (R3Diffuser.score, r3_diffuser.py:410-440, scale=True, use_torch=True => float32):
//   score = -(0.1 x_t - exp(-beta/2) 0.1 x_0) / (1 - exp(-beta)),  beta = 0.1 t + 0.5 t^2 19.9
// ------------------------------------------------------------------------------------------------
__global__ void trans_score_kernel(int M, int N, const float* __restrict__ trans_t, int ldt, const float* __restrict__ trans_0,
                                   int ld0, float scale0 /* multiplies trans_0 before the 0.1 scaling (unscale) */,
                                   const float* __restrict__ t32, float min_b, float max_b, float cs, int do_scale,
                                   const float* __restrict__ mask, float* __restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float t = t32[m / N];
  const float mb = __fadd_rn(__fmul_rn(t, min_b), __fmul_rn(__fmul_rn(0.5f, __fmul_rn(t, t)), max_b - min_b));
  const float e = expf(-0.5f * mb);
  const float var = 1.f - expf(-mb);
  const float mk = mask ? mask[m] : 1.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float xt = trans_t[(long long)m * ldt + k];
    float x0 = trans_0[(long long)m * ld0 + k] * scale0;
    if (do_scale) {
      xt *= cs;
      x0 *= cs;
    }
    out[(long long)m * 3 + k] = -(xt - e * x0) / var * mk;
  }
}

// ------------------------------------------------------------------------------------------------
// helpers for the reverse step (float64, like the reference's numpy/scipy path)
// ------------------------------------------------------------------------------------------------
FDPT_DEVINL void rot_to_quat_d(const double R[9], double q[4]) {
  // Shepperd's method (what scipy's Rotation.from_matrix does for proper rotations), normalised; (w,x,y,z)
  const double tr = R[0] + R[4] + R[8];
  double w, x, y, z;
  if (tr >= R[0] && tr >= R[4] && tr >= R[8]) {
    w = 1.0 + tr;
    x = R[7] - R[5];
    y = R[2] - R[6];
    z = R[3] - R[1];
  } else if (R[0] >= R[4] && R[0] >= R[8]) {
    x = 1.0 - tr + 2.0 * R[0];
    y = R[3] + R[1];
    z = R[6] + R[2];
    w = R[7] - R[5];
  } else if (R[4] >= R[8]) {
    y = 1.0 - tr + 2.0 * R[4];
    z = R[7] + R[5];
    x = R[1] + R[3];
    w = R[2] - R[6];
  } else {
    z = 1.0 - tr + 2.0 * R[8];
    x = R[2] + R[6];
    y = R[5] + R[7];
    w = R[3] - R[1];
  }
  const double inv = 1.0 / sqrt(w * w + x * x + y * y + z * z);
  q[0] = w * inv;
  q[1] = x * inv;
  q[2] = y * inv;
  q[3] = z * inv;
}

FDPT_DEVINL void rotvec_to_quat_d(const double v[3], double q[4]) {
  // scipy Rotation.from_rotvec: small-angle Taylor below 1e-3
  const double a = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  double s;
  if (a <= 1e-3) {
    const double a2 = a * a;
    s = 0.5 - a2 / 48.0 + a2 * a2 / 3840.0;
  } else {
    s = sin(a / 2.0) / a;
  }
  q[0] = cos(a / 2.0);
  q[1] = s * v[0];
  q[2] = s * v[1];
  q[3] = s * v[2];
}

// ------------------------------------------------------------------------------------------------
// Reverse SDE step (SE3Diffuser.reverse se3_diffuser.py:346-401, SO3Diffuser.reverse so3_diffuser.py:569-602,
// R3Diffuser.reverse r3_diffuser.py:344-385, _assemble_rigid 26-36, to_tensor_7).  One CTA per sample
// (the centre-of-mass reduction is per sample: COM = sum_all x' / #diffused, sic).
//   sched: device row of FDPT_SCHED_COLS doubles.
// last_step (t == min_t): rigids_out = rigids_pred (experiments/utils.py:372-374).
// ------------------------------------------------------------------------------------------------
// Counter-based RNG for the throughput mode of the sampler (SURVEY §8b `philox_seed`): Philox4x32-10 (Salmon et al., SC'11; the
// generator behind curand / torch.cuda), keyed by the 64-bit seed, counter = (residue index lo, hi, timestep, draw index).  Parity
// runs never use it: they consume the reference's legacy numpy stream drawn on the host.
FDPT_DEVINL void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += W0; k1 += W1;
  }
}
// two independent N(0,1) doubles from one Philox block (Box-Muller on two 53-bit uniforms in (0,1))
FDPT_DEVINL void philox_normal2(unsigned long long seed, unsigned long long idx, uint32_t step, uint32_t draw, double& z0, double& z1) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), step, draw};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u0 = ((double)((((unsigned long long)c[0]) << 21) ^ (c[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  const double u1 = ((double)((((unsigned long long)c[2]) << 21) ^ (c[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  const double r = sqrt(-2.0 * log(u0));
  double sn, cs;
  sincospi(2.0 * u1, &sn, &cs);
  z0 = r * cs;
  z1 = r * sn;
}
// six normals of residue m at a timestep: z_rot[3], z_trans[3]
FDPT_DEVINL void philox_normal6(unsigned long long seed, unsigned long long m, uint32_t step, double zr[3], double zt[3]) {
  philox_normal2(seed, m, step, 0u, zr[0], zr[1]);
  philox_normal2(seed, m, step, 1u, zr[2], zt[0]);
  philox_normal2(seed, m, step, 2u, zt[1], zt[2]);
}

struct ReverseArgs {
  int N;
  const float* rigids_t;      // [B,N,7]
  const double* rot_score;    // [B,N,3]
  const float* trans_score;   // [B,N,3]
  const float* dmask;         // [B,N]
  const double* z_rot;        // [B,N,3]
  const double* z_trans;      // [B,N,3]
  const double* sched;        // [8]
  int center, diffuse_rot, diffuse_trans;
  float cs;                   // coordinate scaling 0.1
  float* rigids_out;          // [B,N,7]
  StepRef noise_ref, sched_ref;  // optional device-resolved per-step slices of the noise block and the schedule table
  long long noise_half = 0;      // elements between the rot and the trans half of a step's noise block (indirect mode)
  int use_philox = 0;            // 1: draw the normals on the device (z_rot / z_trans / noise_ref bases are ignored)
  unsigned long long philox_seed = 0;
  int philox_step = 0;           // timestep index when no device step counter is attached (noise_ref.step == nullptr)
};

__global__ void __launch_bounds__(256) reverse_kernel(ReverseArgs a) {
  const int b = blockIdx.x, N = a.N;
  uint32_t pstep = (uint32_t)a.philox_step;
  if (a.use_philox) {
    if (a.noise_ref.step) pstep = (uint32_t)*a.noise_ref.step;
    a.sched = step_resolve(a.sched, a.sched_ref);
  } else {
    const double* zr = step_resolve(a.z_rot, a.noise_ref);  // base of this step's [2][B,N,3] noise block when indirect
    a.z_trans = a.noise_ref.base ? zr + a.noise_half : step_resolve(a.z_trans, a.noise_ref);
    a.z_rot = zr;
    a.sched = step_resolve(a.sched, a.sched_ref);
  }
  const double g2dt = a.sched[2], gn = a.sched[3], b_t = a.sched[4], dt = a.sched[5], rn = a.sched[6];
  __shared__ double red[4][8];
  __shared__ double com[3];
  const double csd = (double)a.cs;
  // pass 1: translations (x' before centring) -> accumulate sums
  double sx = 0, sy = 0, sz = 0, sm = 0;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const long long m = (long long)b * N + n;
    const double mk = (double)a.dmask[m];
    sm += mk;
    double xp[3];
    double zr3[3], zt3[3];
    if (a.use_philox) philox_normal6(a.philox_seed, (unsigned long long)m, pstep, zr3, zt3);
    else { zt3[0] = a.z_trans[m * 3]; zt3[1] = a.z_trans[m * 3 + 1]; zt3[2] = a.z_trans[m * 3 + 2]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float x32 = __fmul_rn(a.rigids_t[m * 7 + 4 + k], a.cs);  // float32 array * python float
      const double x = (double)x32;
      const double f = -0.5 * b_t * x;
      const double perturb = ((f - b_t * (double)a.trans_score[m * 3 + k]) * dt + rn * zt3[k]) * mk;
      xp[k] = x - perturb;
    }
    sx += xp[0];
    sy += xp[1];
    sz += xp[2];
  }
  sx = warp_sum(sx);
  sy = warp_sum(sy);
  sz = warp_sum(sz);
  sm = warp_sum(sm);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][w] = sx;
    red[1][w] = sy;
    red[2][w] = sz;
    red[3][w] = sm;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[4] = {0, 0, 0, 0};
    for (int k = 0; k < 8; ++k) {
      t[0] += red[0][k];
      t[1] += red[1][k];
      t[2] += red[2][k];
      t[3] += red[3][k];
    }
    com[0] = t[0] / t[3];
    com[1] = t[1] / t[3];
    com[2] = t[2] / t[3];
  }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const long long m = (long long)b * N + n;
    const double mk = (double)a.dmask[m];
    float outv[7];
    double zr3[3], zt3[3];
    if (a.use_philox) philox_normal6(a.philox_seed, (unsigned long long)m, pstep, zr3, zt3);
    else {
#pragma unroll
      for (int k = 0; k < 3; ++k) { zr3[k] = a.z_rot[m * 3 + k]; zt3[k] = a.z_trans[m * 3 + k]; }
    }
    // ---- translation
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float xt = a.rigids_t[m * 7 + 4 + k];
      double res = (double)xt;
      if (a.diffuse_trans) {
        const double x = (double)__fmul_rn(xt, a.cs);
        const double f = -0.5 * b_t * x;
        const double perturb = ((f - b_t * (double)a.trans_score[m * 3 + k]) * dt + rn * zt3[k]) * mk;
        double xp = x - perturb;
        if (a.center) xp -= com[k];
        xp = xp / csd;
        res = mk * xp + (1.0 - mk) * (double)xt;
      }
      outv[4 + k] = (float)res;
    }
    // ---- rotation: R_t (fp32, from the fp32 quaternion like get_rot_mats) -> float64 quaternion algebra
    float qf[4] = {a.rigids_t[m * 7], a.rigids_t[m * 7 + 1], a.rigids_t[m * 7 + 2], a.rigids_t[m * 7 + 3]};
    float Rf[9];
    quat_to_rot(qf, Rf);
    double Rd[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rd[k] = (double)Rf[k];
    double qt[4];
    rot_to_quat_d(Rd, qt);
    double qn[4] = {qt[0], qt[1], qt[2], qt[3]};
    if (a.diffuse_rot) {
      double pv[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) pv[k] = (g2dt * a.rot_score[m * 3 + k] + gn * zr3[k]) * mk;
      double qp[4];
      rotvec_to_quat_d(pv, qp);
      quat_mul(qt, qp, qn);  // right multiply: R_t * exp(perturb)
    }
    // _assemble_rigid stores float32 rotation matrices; to_tensor_7 extracts the quaternion of that matrix
    double Rn[9];
    quat_to_rot(qn, Rn);
    double inv = 1.0 / (qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
#pragma unroll
    for (int k = 0; k < 9; ++k) Rn[k] = (double)(float)(Rn[k] * inv);
    double qo[4];
    rot_to_quat_d(Rn, qo);
#pragma unroll
    for (int k = 0; k < 4; ++k) outv[k] = (float)qo[k];
#pragma unroll
    for (int k = 0; k < 7; ++k) a.rigids_out[m * 7 + k] = outv[k];
  }
}

// ------------------------------------------------------------------------------------------------
// Backbone atoms (all_atom.compute_backbone all_atom.py:147-176; feats.torsion_angles_to_frames feats.py:165-228;
// frames_to_atom14_pos all_atom.py:108-144).  N, CA, C, CB = R ideal + t ; O = R (R_psi rot_x(psi) ideal_O + t_psi) + t.
// out [M,5,3] in atom37 slot order N, CA, C, CB, O.  float32 like the reference.
// tables: ideal [20,5,3] (atom14 order N,CA,C,O,CB), psi_frame [20,4,4], atom_mask [20,5]
// ------------------------------------------------------------------------------------------------
__global__ void backbone_kernel(int M, const float* __restrict__ rigids, const float* __restrict__ psi,
                                const int32_t* __restrict__ aatype, const float* __restrict__ ideal,
                                const float* __restrict__ psi_frame, const float* __restrict__ atom_mask,
                                float* __restrict__ out, StepRef out_ref = StepRef()) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  out = step_resolve(out, out_ref);
  int aa = aatype ? aatype[m] : 0;
  if (aa >= 20 || aa < 0) aa = 0;
  float q[4] = {rigids[m * 7], rigids[m * 7 + 1], rigids[m * 7 + 2], rigids[m * 7 + 3]};
  const float t[3] = {rigids[m * 7 + 4], rigids[m * 7 + 5], rigids[m * 7 + 6]};
  float R[9];
  quat_to_rot(q, R);
  const float* id = ideal + aa * 15;
  const float* am = atom_mask + aa * 5;
  const float* F = psi_frame + aa * 16;
  const float s = psi[m * 2], c = psi[m * 2 + 1];
  // R_psi = F_rot * rot_x ; rot_x = [[1,0,0],[0,c,-s],[0,s,c]]
  float Rp[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float f0 = F[r * 4 + 0], f1 = F[r * 4 + 1], f2 = F[r * 4 + 2];
    Rp[r * 3 + 0] = f0;
    Rp[r * 3 + 1] = f1 * c + f2 * s;
    Rp[r * 3 + 2] = -f1 * s + f2 * c;
  }
  // global psi frame: Rg = R * Rp ; tg = R * t_psi + t
  float Rg[9], tg[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) Rg[r * 3 + cc] = R[r * 3] * Rp[cc] + R[r * 3 + 1] * Rp[3 + cc] + R[r * 3 + 2] * Rp[6 + cc];
    tg[r] = R[r * 3] * F[3] + R[r * 3 + 1] * F[7] + R[r * 3 + 2] * F[11] + t[r];
  }
  const int a14_of_slot[5] = {0, 1, 2, 4, 3};  // atom37 slots N,CA,C,CB,O <- atom14 N,CA,C,O,CB
#pragma unroll
  for (int sl = 0; sl < 5; ++sl) {
    const int a14 = a14_of_slot[sl];
    const float* p = id + a14 * 3;
    const float mk = am[a14];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float v;
      if (a14 == 3)
        v = Rg[r * 3] * p[0] + Rg[r * 3 + 1] * p[1] + Rg[r * 3 + 2] * p[2] + tg[r];
      else
        v = R[r * 3] * p[0] + R[r * 3 + 1] * p[1] + R[r * 3 + 2] * p[2] + t[r];
      out[((long long)m * 5 + sl) * 3 + r] = v * mk;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// x_T for B samples of one structure (SE3Diffuser.sample_ref se3_diffuser.py:455-529; SO3Diffuser.sample so3_diffuser.py:325-357:
// unit(randn(n,3)) * omega with omega = np.interp(rand(n), cdf_row, discrete_omega); R3Diffuser.sample_stationary_distribution
// r3_diffuser.py:294-331: N(0,1) on the diffused residues in scaled coordinates; mask blend with the imputed (ground-truth) frames;
// _assemble_rigid: float32 rotation matrices, to_tensor_7: quaternion of that matrix).  float64 like the reference's numpy path.
// ------------------------------------------------------------------------------------------------
FDPT_DEVINL void quat_to_rotvec_d(const double qin[4], double v[3]) {  // scipy Rotation.as_rotvec
  double q[4] = {qin[0], qin[1], qin[2], qin[3]};
  if (q[0] < 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = -q[k];
  }
  const double n = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double angle = 2.0 * atan2(n, q[0]);
  const double a2 = angle * angle;
  const double sc = angle <= 1e-3 ? 2.0 + a2 / 12.0 + 7.0 * a2 * a2 / 2880.0 : angle / sin(angle / 2.0);
  v[0] = sc * q[1];
  v[1] = sc * q[2];
  v[2] = sc * q[3];
}

struct SampleRefArgs {
  int N;
  const float* impute;        // [N,7] or nullptr
  const float* dmask;         // [N] or nullptr
  const double* cdf;          // [num_omega]
  const double* omega_grid;   // [num_omega]
  int num_omega;
  const double* draws;        // [B][7N]: randn [N,3] | rand [N] | normal [N,3] (by residue), or nullptr (Philox)
  unsigned long long philox_seed;
  int diffuse_rot, diffuse_trans;
  float cs;
  float* out;                 // [B,N,7]
};

__global__ void __launch_bounds__(128) sample_ref_kernel(SampleRefArgs a) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y, N = a.N;
  if (n >= N) return;
  const long long m = (long long)b * N + n;
  double g[3], u, z[3];
  if (a.draws) {
    const double* d = a.draws + (long long)b * 7 * N;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      g[k] = d[n * 3 + k];
      z[k] = d[4 * N + n * 3 + k];
    }
    u = d[3 * N + n];
  } else {
    double t0, t1;
    philox_normal2(a.philox_seed, (unsigned long long)m, 0xFFFFFFFFu, 0u, g[0], g[1]);
    philox_normal2(a.philox_seed, (unsigned long long)m, 0xFFFFFFFFu, 1u, g[2], z[0]);
    philox_normal2(a.philox_seed, (unsigned long long)m, 0xFFFFFFFFu, 2u, z[1], z[2]);
    uint32_t c[4] = {(uint32_t)m, (uint32_t)((unsigned long long)m >> 32), 0xFFFFFFFFu, 3u};
    philox4x32_10(c, (uint32_t)a.philox_seed, (uint32_t)(a.philox_seed >> 32));
    u = (double)((((unsigned long long)c[0]) << 21) ^ (c[1] >> 11)) * (1.0 / 9007199254740992.0);  // [0, 1) like numpy.random.rand
    (void)t0; (void)t1;
  }
  const double dm = a.dmask ? (double)a.dmask[n] : 1.0;
  // imputed frame
  double rot_imp[3] = {0, 0, 0};
  float t_imp[3] = {0.f, 0.f, 0.f};
  if (a.impute) {
    float qf[4] = {a.impute[n * 7], a.impute[n * 7 + 1], a.impute[n * 7 + 2], a.impute[n * 7 + 3]};
    float Rf[9];
    quat_to_rot(qf, Rf);  // fp32 rotation matrix (Rigid.get_rots().get_rot_mats())
    double Rd[9], qd[4];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rd[k] = (double)Rf[k];
    rot_to_quat_d(Rd, qd);
    quat_to_rotvec_d(qd, rot_imp);
#pragma unroll
    for (int k = 0; k < 3; ++k) t_imp[k] = a.impute[n * 7 + 4 + k];
  }
  // rotation
  double rv[3] = {rot_imp[0], rot_imp[1], rot_imp[2]};
  if (a.diffuse_rot) {
    const double gn = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    // np.interp(u, cdf, omega)
    double ang;
    const int K = a.num_omega;
    if (u <= a.cdf[0]) ang = a.omega_grid[0];
    else if (u >= a.cdf[K - 1]) ang = a.omega_grid[K - 1];
    else {
      int lo = 0, hi = K - 1;  // invariant: cdf[lo] <= u < cdf[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.cdf[mid] <= u) lo = mid; else hi = mid;
      }
      const double slope = (a.omega_grid[lo + 1] - a.omega_grid[lo]) / (a.cdf[lo + 1] - a.cdf[lo]);
      ang = slope * (u - a.cdf[lo]) + a.omega_grid[lo];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double s = g[k] / gn * ang;
      rv[k] = a.dmask ? dm * s + (1.0 - dm) * rot_imp[k] : s;
    }
  }
  double qn[4], Rn[9], qo[4];
  rotvec_to_quat_d(rv, qn);
  quat_to_rot(qn, Rn);
#pragma unroll
  for (int k = 0; k < 9; ++k) Rn[k] = (double)(float)Rn[k];  // torch.Tensor(rotmat): float32 storage
  rot_to_quat_d(Rn, qo);
#pragma unroll
  for (int k = 0; k < 4; ++k) a.out[m * 7 + k] = (float)qo[k];
  // translation
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float x = t_imp[k];
    if (a.diffuse_trans) {
      float xs = __fmul_rn(t_imp[k], a.cs);
      if (dm != 0.0) xs = (float)z[k];
      x = __fdiv_rn(xs, a.cs);
    }
    a.out[m * 7 + 4 + k] = x;
  }
}

// trans_traj row (experiments/utils.py:379-384): diffuse_mask * pred_trans + fixed*res_mask * rigids_{t-1} trans
__global__ void trans0_kernel(int M, const float* __restrict__ rig_pred, const float* __restrict__ rig_next,
                              const float* __restrict__ res_mask, const float* __restrict__ fixed_mask,
                              float* __restrict__ out, StepRef out_ref = StepRef()) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  out = step_resolve(out, out_ref);
  const float dm = (1.f - fixed_mask[m]) * res_mask[m], fm = fixed_mask[m] * res_mask[m];
#pragma unroll
  for (int k = 0; k < 3; ++k) out[m * 3 + k] = dm * rig_pred[m * 7 + 4 + k] + fm * rig_next[m * 7 + 4 + k];
}

}  // namespace fdpt
