// Fused UnifiedGaussianAdapter: raw head output [N, 7 + 3K] -> scales [N,3], rotations [N,4],
// harmonics [N,3,K] in ONE pass (and its backward), without materialising the [N,3,3] covariances the decoder never
// reads.  Replaces the ~10 elementwise torch kernels of
//   UnifiedGaussianAdapter.forward   /root/reference/src/model/encoder/common/gaussian_adapter.py:122-150
//     scales    = clamp_max(0.001 * softplus(raw[0:3]), 0.3)
//     rotations = raw[3:7] / (|raw[3:7]| + eps)
//     harmonics = raw[7:].view(3, K) * sh_mask          (sh_mask[0] = 1, degree d >= 1: 0.1 * 0.25^d, :42-48)
// and, when the row carries the density logit in front (the encoder head's 83-channel output,
// encoder_spfsplatv2.py:255-268), the opacity mapping of EncoderSPFSplatV2.map_pdf_to_opacity (:146-159):
//     p = sigmoid(raw[0]);  opacity = 0.5 * (1 - (1 - p)^e + p^(1/e)),  e = 2^x  (x from the warm-up schedule, host side)
// HBM-bound: 2 * (7 + 3K) * 4 bytes per Gaussian each way.  A block stages 128 raw rows in shared memory with coalesced
// 128-bit loads and writes the three outputs element-parallel (coalesced) -- no per-thread 328-byte strides.
#include "spf_adapter_math.cuh"
#include "spf_device.cuh"
#include "spf_kernels.h"

namespace spf {

constexpr int AD_ROWS = 128;
constexpr int AD_THREADS = 256;

// `dens` = 1: rows are [density logit, 7 + 3K parameters] and `opac` receives the mapped opacity; 0: rows are the 7 + 3K
// parameters only (the adapter's own contract).
__global__ void __launch_bounds__(AD_THREADS)
adapter_forward_kernel(const float* __restrict__ raw_all, int64_t n, int K, float eps, int dens, float exponent,
                       float* __restrict__ scales, float* __restrict__ rots, float* __restrict__ sh, float* __restrict__ opac) {
  extern __shared__ __align__(16) float tile_all[];
  const int R = dens + 7 + 3 * K;
  const int64_t g0 = (int64_t)blockIdx.x * AD_ROWS;
  const int rows = (int)min((int64_t)AD_ROWS, n - g0);
  block_copy_g2s(tile_all, raw_all + g0 * R, rows * R, R, R, threadIdx.x, AD_THREADS);
  __syncthreads();
  const float* tile = tile_all + dens;           // row g of the 7 + 3K parameters starts at tile + g * R
  if (dens && opac) {
    for (int g = threadIdx.x; g < rows; g += AD_THREADS) {
      opac[g0 + g] = head_opacity(tile_all[g * R], exponent);
    }
  }
  for (int i = threadIdx.x; i < rows * 3; i += AD_THREADS) {
    const int g = i / 3, c = i - g * 3;
    scales[g0 * 3 + i] = head_scale(tile[g * R + c]);
  }
  for (int i = threadIdx.x; i < rows * 4; i += AD_THREADS) {
    const int g = i >> 2, c = i & 3;
    const float* q = tile + g * R + 3;
    rots[g0 * 4 + i] = __fdiv_rn(q[c], __fadd_rn(quat_norm(q), eps));
  }
  const int W = 3 * K;
  for (int i = threadIdx.x; i < rows * W; i += AD_THREADS) {
    const int g = i / W, j = i - g * W;
    sh[g0 * W + i] = tile[g * R + 7 + j] * sh_mask_of(j % K);
  }
}

__global__ void __launch_bounds__(AD_THREADS)
adapter_backward_kernel(const float* __restrict__ raw, const float* __restrict__ d_scales, const float* __restrict__ d_rots,
                        const float* __restrict__ d_sh, const float* __restrict__ d_opac, int64_t n, int K, float eps, int dens,
                        float exponent, float* __restrict__ d_raw) {
  extern __shared__ __align__(16) float tile_all[];      // raw rows, overwritten in place by d_raw rows
  __shared__ float dq[AD_ROWS * 4];
  const int R = dens + 7 + 3 * K;
  const int64_t g0 = (int64_t)blockIdx.x * AD_ROWS;
  const int rows = (int)min((int64_t)AD_ROWS, n - g0);
  block_copy_g2s(tile_all, raw + g0 * R, rows * R, R, R, threadIdx.x, AD_THREADS);
  for (int i = threadIdx.x; i < rows * 4; i += AD_THREADS) dq[i] = d_rots ? d_rots[g0 * 4 + i] : 0.0f;
  __syncthreads();
  float* tile = tile_all + dens;
  if (dens) {
    // d opacity / d logit = 0.5 (e (1-p)^(e-1) + (1/e) p^(1/e-1)) p (1-p)
    for (int g = threadIdx.x; g < rows; g += AD_THREADS) {
      tile_all[g * R] = d_opac ? d_opac[g0 + g] * head_opacity_grad(tile_all[g * R], exponent) : 0.0f;
    }
  }
  // quaternion part first (needs all four raw components of a row before any is overwritten): one thread per row
  for (int g = threadIdx.x; g < rows; g += AD_THREADS) {
    float* q = tile + g * R + 3;
    const float qr[4] = {q[0], q[1], q[2], q[3]};
    float o4[4];
    // d/dq_raw [ q / (|q| + eps) ] = dq/(n+eps) - q (q . dq) / (n (n+eps)^2)
    head_quat_grad(qr, dq + g * 4, eps, o4);
    q[0] = o4[0]; q[1] = o4[1]; q[2] = o4[2]; q[3] = o4[3];
  }
  for (int i = threadIdx.x; i < rows * 3; i += AD_THREADS) {
    const int g = i / 3, c = i - g * 3;
    const float x = tile[g * R + c];
    const float gs = d_scales ? d_scales[g0 * 3 + i] : 0.0f;
    tile[g * R + c] = gs * head_scale_grad(x);
  }
  const int W = 3 * K;
  for (int i = threadIdx.x; i < rows * W; i += AD_THREADS) {
    const int g = i / W, j = i - g * W;
    tile[g * R + 7 + j] = d_sh ? d_sh[g0 * W + i] * sh_mask_of(j % K) : 0.0f;
  }
  __syncthreads();
  block_copy_s2g(d_raw + g0 * R, tile_all, rows * R, R, R, threadIdx.x, AD_THREADS);
}

cudaError_t launch_adapter_forward(const float* raw, int64_t n, int K, float eps, int dens, float exponent, float* scales,
                                   float* rots, float* sh, float* opac, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const size_t smem = (size_t)AD_ROWS * (dens + 7 + 3 * K) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(adapter_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  adapter_forward_kernel<<<(unsigned)((n + AD_ROWS - 1) / AD_ROWS), AD_THREADS, smem, s>>>(raw, n, K, eps, dens, exponent, scales,
                                                                                         rots, sh, opac);
  return cudaGetLastError();
}

cudaError_t launch_adapter_backward(const float* raw, const float* d_scales, const float* d_rots, const float* d_sh,
                                    const float* d_opac, int64_t n, int K, float eps, int dens, float exponent, float* d_raw,
                                    cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const size_t smem = (size_t)AD_ROWS * (dens + 7 + 3 * K) * sizeof(float);
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(adapter_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  adapter_backward_kernel<<<(unsigned)((n + AD_ROWS - 1) / AD_ROWS), AD_THREADS, smem, s>>>(raw, d_scales, d_rots, d_sh, d_opac,
                                                                                          n, K, eps, dens, exponent, d_raw);
  return cudaGetLastError();
}

}  // namespace spf
