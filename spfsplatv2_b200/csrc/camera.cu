// Camera setup fused into one tiny kernel (one thread per view), forward and backward.
//
// Replaces the ~60 small torch kernels the reference's host glue launches per render_cuda call:
// scale-invariance of the extrinsics (/root/reference/src/model/decoder/cuda_splatting.py:66-74),
// get_fov (src/geometry/projection.py:269-283), get_projection_matrix (cuda_splatting.py:15-42) and
// extrinsics.inverse() + the two transposes (cuda_splatting.py:88-90).
#include "spf_device.cuh"
#include "spf_kernels.h"

namespace spf {

__device__ __forceinline__ bool inv3(const float* m, float* o) {
  const float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  const float id = 1.0f / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return det != 0.0f;
}

// general 4x4 inverse (cofactor expansion), row-major
__device__ __forceinline__ void inv4(const float* m, float* inv) {
  float t[16];
  t[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  t[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  t[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  t[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  t[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  t[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  t[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  t[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  t[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  t[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  t[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  t[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  t[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  t[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  t[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  t[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  const float det = m[0] * t[0] + m[1] * t[4] + m[2] * t[8] + m[3] * t[12];
  const float id = 1.0f / det;
  for (int i = 0; i < 16; ++i) inv[i] = t[i] * id;
}

__global__ void camera_forward_kernel(int B, int scale_invariant, const float* __restrict__ ext,
                                      const float* __restrict__ intr, const float* __restrict__ near,
                                      const float* __restrict__ far, float* __restrict__ view,
                                      float* __restrict__ proj, float* __restrict__ tanfov,
                                      float* __restrict__ pre_scale) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float n = near[i], f = far[i];
  const float scale = scale_invariant ? 1.0f / n : 1.0f;
  float E[16];
  for (int k = 0; k < 16; ++k) E[k] = ext[i * 16 + k];
  if (scale_invariant) {
    E[3] *= scale; E[7] *= scale; E[11] *= scale;
    n = n * scale; f = f * scale;
  }
  float M[16];
  inv4(E, M);
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) view[i * 16 + 4 * r + c] = M[4 * c + r];   // transpose
  // field of view from the normalised intrinsics
  float Ki[9];
  inv3(intr + i * 9, Ki);
  const float e[4][3] = {{0.f, 0.5f, 1.f}, {1.f, 0.5f, 1.f}, {0.5f, 0.f, 1.f}, {0.5f, 1.f, 1.f}};
  float ray[4][3];
  for (int k = 0; k < 4; ++k) {
    float v[3];
    for (int r = 0; r < 3; ++r) v[r] = Ki[3 * r] * e[k][0] + Ki[3 * r + 1] * e[k][1] + Ki[3 * r + 2] * e[k][2];
    const float inv = 1.0f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int r = 0; r < 3; ++r) ray[k][r] = v[r] * inv;
  }
  const float fov_x = acosf(ray[0][0] * ray[1][0] + ray[0][1] * ray[1][1] + ray[0][2] * ray[1][2]);
  const float fov_y = acosf(ray[2][0] * ray[3][0] + ray[2][1] * ray[3][1] + ray[2][2] * ray[3][2]);
  const float tx = tanf(0.5f * fov_x), ty = tanf(0.5f * fov_y);
  tanfov[2 * i] = tx; tanfov[2 * i + 1] = ty;
  pre_scale[i] = scale;
  const float right = tx * n, top = ty * n;
  float P[16];
  for (int k = 0; k < 16; ++k) P[k] = 0.0f;
  P[0] = 2.0f * n / (right + right);
  P[5] = 2.0f * n / (top + top);
  P[14] = 1.0f;                         // P[3][2]
  P[10] = f / (f - n);                  // P[2][2]
  P[11] = -(f * n) / (f - n);           // P[2][3]
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) proj[i * 16 + 4 * r + c] = P[4 * c + r];   // transpose
}

// dL/dext from dL/dview:  view = (E'^-1)^T ;  dL/dE' = -(M^T) (dL/dM) (M^T), M = E'^-1 ; the
// translation column of E' is scale * that of E.
__global__ void camera_backward_kernel(int B, int scale_invariant, const float* __restrict__ near,
                                       const float* __restrict__ view, const float* __restrict__ d_view,
                                       float* __restrict__ d_ext) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float Mt[16], G[16];   // Mt = M^T = view ; G = dL/dM = (dL/dview)^T
  for (int k = 0; k < 16; ++k) Mt[k] = view[i * 16 + k];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) G[4 * r + c] = d_view[i * 16 + 4 * c + r];
  float A[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += Mt[4 * r + k] * G[4 * k + c];
      A[4 * r + c] = s;
    }
  const float scale = scale_invariant ? 1.0f / near[i] : 1.0f;
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += A[4 * r + k] * Mt[4 * k + c];
      float g = -s;
      if (c == 3 && r < 3) g *= scale;
      d_ext[i * 16 + 4 * r + c] = g;
    }
}

cudaError_t launch_camera_forward(int B, int scale_invariant, const float* ext, const float* intr,
                                  const float* near, const float* far, float* view, float* proj, float* tanfov,
                                  float* pre_scale, cudaStream_t s) {
  pdl_launch(camera_forward_kernel, (B + 63) / 64, 64, 0, s)(B, scale_invariant, ext, intr, near, far, view, proj, tanfov,
                                                     pre_scale);
  return cudaGetLastError();
}
cudaError_t launch_camera_backward(int B, int scale_invariant, const float* near, const float* view,
                                   const float* d_view, float* d_ext, cudaStream_t s) {
  pdl_launch(camera_backward_kernel, (B + 63) / 64, 64, 0, s)(B, scale_invariant, near, view, d_view, d_ext);
  return cudaGetLastError();
}

}  // namespace spf
