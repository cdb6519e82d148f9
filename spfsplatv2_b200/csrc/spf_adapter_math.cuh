// Elementwise maps of the encoder's Gaussian head post-processing, shared by the stand-alone adapter kernels
// (adapter.cu) and the raw-head variants of the projection kernels (project_fwd.cu / project_bwd.cu), so that both give
// the same bits whatever the translation unit's -fmad setting is (explicit round-to-nearest intrinsics, no contraction):
//   scales    = clamp_max(0.001 * softplus(x), 0.3)                 gaussian_adapter.py:132-133
//   rotations = q / (|q| + eps)                                     gaussian_adapter.py:136
//   harmonics = raw * sh_mask  (1 for degree 0, 0.1 * 0.25^d above)  gaussian_adapter.py:42-48,139-140
//   opacity   = 0.5 * (1 - (1 - p)^e + p^(1/e)),  p = sigmoid(logit) encoder_spfsplatv2.py:146-159,258
#pragma once
#include <cuda_runtime.h>

namespace spf {

__device__ __forceinline__ float softplus_t(float x) { return x > 20.0f ? x : log1pf(expf(x)); }   // torch: beta 1, threshold 20

__device__ __forceinline__ float sh_mask_of(int k) {
  const int deg = (k >= 16) ? 4 : (k >= 9) ? 3 : (k >= 4) ? 2 : (k >= 1) ? 1 : 0;
  const float m[5] = {1.0f, 0.1f * 0.25f, 0.1f * 0.0625f, 0.1f * 0.015625f, 0.1f * 0.00390625f};
  return m[deg];
}

__device__ __forceinline__ float head_scale(float x) { return fminf(__fmul_rn(0.001f, softplus_t(x)), 0.3f); }
// d scale / d x  (clamp_max passes the gradient at equality, like torch)
__device__ __forceinline__ float head_scale_grad(float x) {
  const float sp = __fmul_rn(0.001f, softplus_t(x));
  const float sig = 1.0f / (1.0f + expf(-x));
  return (sp <= 0.3f) ? 0.001f * ((x > 20.0f) ? 1.0f : sig) : 0.0f;
}

__device__ __forceinline__ float quat_norm(const float q[4]) {
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(q[0], q[0]), __fmul_rn(q[1], q[1])),
                         __fadd_rn(__fmul_rn(q[2], q[2]), __fmul_rn(q[3], q[3]))));
}
__device__ __forceinline__ void head_quat(const float q[4], float eps, float out[4]) {
  const float d = __fadd_rn(quat_norm(q), eps);
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = __fdiv_rn(q[c], d);
}
// dL/dq_raw of  q / (|q| + eps)  given dL/dq_normalised
__device__ __forceinline__ void head_quat_grad(const float q[4], const float dq[4], float eps, float out[4]) {
  const float nrm = quat_norm(q);
  const float inv = 1.0f / (nrm + eps);
  const float dot = (q[0] * dq[0] + q[1] * dq[1]) + (q[2] * dq[2] + q[3] * dq[3]);
  const float k = (nrm > 0.0f) ? dot * inv * inv / nrm : 0.0f;
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = dq[c] * inv - q[c] * k;
}

// (exponent 1 -- the shipped schedule, opacity_mapping 0 / 0 -- skips the two powf: x^1 = x)
__device__ __forceinline__ float head_opacity(float logit, float exponent) {
  const float p = 1.0f / (1.0f + expf(-logit));
  if (exponent == 1.0f) return 0.5f * ((1.0f - (1.0f - p)) + p);
  return 0.5f * ((1.0f - powf(1.0f - p, exponent)) + powf(p, 1.0f / exponent));
}
// d opacity / d logit = 0.5 (e (1-p)^(e-1) + (1/e) p^(1/e-1)) p (1-p)
__device__ __forceinline__ float head_opacity_grad(float logit, float exponent) {
  const float p = 1.0f / (1.0f + expf(-logit));
  if (exponent == 1.0f) return p * (1.0f - p);
  const float ie = 1.0f / exponent;
  const float dy = 0.5f * (exponent * powf(1.0f - p, exponent - 1.0f) + ie * powf(p, ie - 1.0f));
  return dy * (p * (1.0f - p));
}

// v[c * K + k] *= mask(degree of k) for a [3][K] block of one Gaussian (no per-element index arithmetic)
__device__ __forceinline__ void sh_mask_rows(float* v, int K) {
  const float m[5] = {1.0f, 0.1f * 0.25f, 0.1f * 0.0625f, 0.1f * 0.015625f, 0.1f * 0.00390625f};
#pragma unroll
  for (int dgr = 1; dgr < 5; ++dgr) {
    const int k1 = min((dgr + 1) * (dgr + 1), K);
    for (int k = dgr * dgr; k < k1; ++k) {
      v[k] = __fmul_rn(v[k], m[dgr]); v[K + k] = __fmul_rn(v[K + k], m[dgr]); v[2 * K + k] = __fmul_rn(v[2 * K + k], m[dgr]);
    }
  }
}

}  // namespace spf
