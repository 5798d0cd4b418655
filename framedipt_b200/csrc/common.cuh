// common.cuh — shared helpers for libfdpt.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#define FDPT_DEVINL __device__ __forceinline__

namespace fdpt {

constexpr int kWarp = 32;

// Model dims the kernels are specialised for (config/base.yaml:55-79 of the reference).
constexpr int C_S = 256, C_Z = 128, C_HID = 256, C_SKIP = 64, NH = 8, PQ = 8, PV = 12, NBLK = 4;
constexpr int EMB = 32, NBINS = 22, TF_D = C_S + C_SKIP /*320*/, TF_H = 4, TF_DH = TF_D / TF_H /*80*/, TF_LAYERS = 2;
constexpr int CAT = NH * (C_Z / 4 + C_HID + PV * 4);  // 2688
constexpr int CAT_OPT = NH * C_HID;                    // 2048: o_pt.x block start
constexpr int CAT_NRM = CAT_OPT + 3 * NH * PV;         // 2336
constexpr int CAT_PAIR = CAT_NRM + NH * PV;            // 2432
constexpr int ET_HID = C_Z + C_S;                      // 384

FDPT_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
FDPT_DEVINL double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
FDPT_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// quaternion (w,x,y,z) -> rotation matrix, same polynomial as openfold/utils/rigid_utils.py:161-205
template <typename T>
FDPT_DEVINL void quat_to_rot(const T q[4], T R[9]) {
  const T a = q[0], b = q[1], c = q[2], d = q[3];
  R[0] = a * a + b * b - c * c - d * d;
  R[1] = 2 * (b * c - a * d);
  R[2] = 2 * (b * d + a * c);
  R[3] = 2 * (b * c + a * d);
  R[4] = a * a - b * b + c * c - d * d;
  R[5] = 2 * (c * d - a * b);
  R[6] = 2 * (b * d - a * c);
  R[7] = 2 * (c * d + a * b);
  R[8] = a * a - b * b - c * c + d * d;
}

// Hamilton product, rigid_utils.py:229-263
template <typename T>
FDPT_DEVINL void quat_mul(const T p[4], const T q[4], T r[4]) {
  r[0] = p[0] * q[0] - p[1] * q[1] - p[2] * q[2] - p[3] * q[3];
  r[1] = p[0] * q[1] + p[1] * q[0] + p[2] * q[3] - p[3] * q[2];
  r[2] = p[0] * q[2] - p[1] * q[3] + p[2] * q[0] + p[3] * q[1];
  r[3] = p[0] * q[3] + p[1] * q[2] - p[2] * q[1] + p[3] * q[0];
}

}  // namespace fdpt
