// PLY vertex rows of a Gaussian scene, packed on the device (the on-disk format next to the path).
//
// Replaces the numpy / scipy / einops body of export_ply (/root/reference/src/model/ply_export.py:76-141): per Gaussian
//   x y z          = R (mean - shift) / scale_factor                    (:87-126; R = viewer rotation o w2c rotation)
//   nx ny nz       = 0                                                  (:136)
//   f_dc_0..2      = harmonics[:, :, 0]                                 (:131, only the DC band is exported)
//   opacity        = opacities (as stored, no logit)                    (:138)
//   scale_0..2     = log(scales / scale_factor)                         (:93,139)
//   rot_0..3       = (w, x, y, z) of quat(R * matrix(q_xyzw))           (:129-133; scipy's from_quat normalises, its
//                    from_matrix -> as_quat picks the largest of {m00, m11, m22, trace} as the pivot)
// 17 floats = 68 bytes per row, written through shared memory so that the global stores are coalesced.
// `params` (device, 13 floats) = R row-major (9), shift (3), scale_factor (1): produced on the device by the caller
// (median / quantile / 3x3 inverse), so the export never waits for the GPU before the final copy to the host.
#include "spf_device.cuh"
#include "spf_kernels.h"

namespace spf {

constexpr int PLY_ROW = 17;
constexpr int PLY_THREADS = 128;

__global__ void __launch_bounds__(PLY_THREADS)
ply_pack_kernel(const float* __restrict__ means, const float* __restrict__ scales, const float* __restrict__ rots,
                const float* __restrict__ harmonics, const float* __restrict__ opac, const float* __restrict__ params,
                int64_t n, int sh_coeffs, float* __restrict__ out) {
  __shared__ float rows[PLY_THREADS * PLY_ROW];
  __shared__ float prm[13];
  const int tid = threadIdx.x;
  const int64_t g0 = (int64_t)blockIdx.x * PLY_THREADS;
  const int64_t g = g0 + tid;
  if (tid < 13) prm[tid] = params[tid];
  __syncthreads();
  if (g < n) {
    const float* R = prm;
    const float inv_s = 1.0f / prm[12];
    float* r = rows + tid * PLY_ROW;
    const float mx = (means[g * 3] - prm[9]) * inv_s, my = (means[g * 3 + 1] - prm[10]) * inv_s,
                mz = (means[g * 3 + 2] - prm[11]) * inv_s;
    r[0] = R[0] * mx + R[1] * my + R[2] * mz;
    r[1] = R[3] * mx + R[4] * my + R[5] * mz;
    r[2] = R[6] * mx + R[7] * my + R[8] * mz;
    r[3] = 0.0f; r[4] = 0.0f; r[5] = 0.0f;
    const float* h = harmonics + g * 3 * (int64_t)sh_coeffs;
    r[6] = h[0]; r[7] = h[sh_coeffs]; r[8] = h[2 * sh_coeffs];
    r[9] = opac[g];
    r[10] = logf(scales[g * 3] * inv_s); r[11] = logf(scales[g * 3 + 1] * inv_s); r[12] = logf(scales[g * 3 + 2] * inv_s);
    // rotation: normalised (x, y, z, w) -> matrix -> R * matrix -> quaternion
    float qx = rots[g * 4], qy = rots[g * 4 + 1], qz = rots[g * 4 + 2], qw = rots[g * 4 + 3];
    const float qn = rsqrtf(qx * qx + qy * qy + qz * qz + qw * qw);
    qx *= qn; qy *= qn; qz *= qn; qw *= qn;
    float A[9];
    A[0] = 1.0f - 2.0f * (qy * qy + qz * qz); A[1] = 2.0f * (qx * qy - qz * qw); A[2] = 2.0f * (qx * qz + qy * qw);
    A[3] = 2.0f * (qx * qy + qz * qw); A[4] = 1.0f - 2.0f * (qx * qx + qz * qz); A[5] = 2.0f * (qy * qz - qx * qw);
    A[6] = 2.0f * (qx * qz - qy * qw); A[7] = 2.0f * (qy * qz + qx * qw); A[8] = 1.0f - 2.0f * (qx * qx + qy * qy);
    float M[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) M[3 * i + j] = R[3 * i] * A[j] + R[3 * i + 1] * A[3 + j] + R[3 * i + 2] * A[6 + j];
    const float tr = M[0] + M[4] + M[8];
    float q[4];   // x, y, z, w
    const float dmax = fmaxf(fmaxf(M[0], M[4]), M[8]);
    if (tr >= dmax) {
      q[0] = M[7] - M[5]; q[1] = M[2] - M[6]; q[2] = M[3] - M[1]; q[3] = 1.0f + tr;
    } else if (M[0] >= M[4] && M[0] >= M[8]) {       // pivot x: i = 0, j = 1, k = 2
      q[0] = 1.0f - tr + 2.0f * M[0]; q[1] = M[3] + M[1]; q[2] = M[6] + M[2]; q[3] = M[7] - M[5];
    } else if (M[4] >= M[8]) {                       // pivot y: i = 1, j = 2, k = 0
      q[1] = 1.0f - tr + 2.0f * M[4]; q[2] = M[7] + M[5]; q[0] = M[1] + M[3]; q[3] = M[2] - M[6];
    } else {                                         // pivot z: i = 2, j = 0, k = 1
      q[2] = 1.0f - tr + 2.0f * M[8]; q[0] = M[2] + M[6]; q[1] = M[5] + M[7]; q[3] = M[3] - M[1];
    }
    const float n2 = rsqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    r[13] = q[3] * n2; r[14] = q[0] * n2; r[15] = q[1] * n2; r[16] = q[2] * n2;
  }
  __syncthreads();
  const int64_t valid = min((int64_t)PLY_THREADS, n - g0);
  float* dst = out + g0 * PLY_ROW;
  for (int i = tid; i < valid * PLY_ROW; i += PLY_THREADS) dst[i] = rows[i];
}

cudaError_t launch_ply_pack(const float* means, const float* scales, const float* rots, const float* harmonics,
                            const float* opac, const float* params, int64_t n, int sh_coeffs, float* out, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  ply_pack_kernel<<<(unsigned)((n + PLY_THREADS - 1) / PLY_THREADS), PLY_THREADS, 0, s>>>(means, scales, rots, harmonics, opac,
                                                                                          params, n, sh_coeffs, out);
  return cudaGetLastError();
}

}  // namespace spf
