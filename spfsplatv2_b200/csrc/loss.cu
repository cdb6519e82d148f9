// Fused image losses on the decoder's output: per-image mean squared error between the rendered
// colour and the target -- optionally after clipping both to [0,1] (the PSNR definition) -- and, in the same pass,
// dL/dcolor for the MSE training loss, so the image gradient is ready before autograd even asks for it.
//
// Replaces, for the renderer's output:
//   LossMse.forward   /root/reference/src/loss/loss_mse.py:36-51      weight * ((prediction - image) ** 2).mean()
//   compute_psnr      /root/reference/src/evaluation/metrics.py:12-19 -10 log10(mean((clip(gt) - clip(pred))^2)) per image
// and the ~6 elementwise / reduction kernels torch launches for them.  Deterministic: per-block partial sums in a fixed
// tree, summed in block order by a second tiny kernel (no float atomics).
#include "spf_device.cuh"
#include "spf_kernels.h"

namespace spf {

constexpr int LOSS_THREADS = 256;

__global__ void __launch_bounds__(LOSS_THREADS)
image_mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n_per_image, int clip,
                         float grad_scale, float* __restrict__ dL_dpred, float* __restrict__ partial) {
  pdl_enter();
  __shared__ float warp_part[LOSS_THREADS / 32];
  const int img = blockIdx.y;
  const int64_t base = (int64_t)img * n_per_image;
  const int64_t n4 = n_per_image >> 2;
  const bool vec = ((n_per_image & 3) == 0) && (((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target) |
                                                  reinterpret_cast<uintptr_t>(dL_dpred)) & 15) == 0);
  float acc = 0.0f;
  if (vec) {
    const float4* p4 = reinterpret_cast<const float4*>(pred + base);
    const float4* t4 = reinterpret_cast<const float4*>(target + base);
    float4* g4 = dL_dpred ? reinterpret_cast<float4*>(dL_dpred + base) : nullptr;
    for (int64_t i = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * LOSS_THREADS) {
      float4 p = p4[i], t = __ldg(t4 + i);
      if (clip) {
        p.x = __saturatef(p.x); p.y = __saturatef(p.y); p.z = __saturatef(p.z); p.w = __saturatef(p.w);
        t.x = __saturatef(t.x); t.y = __saturatef(t.y); t.z = __saturatef(t.z); t.w = __saturatef(t.w);
      }
      const float dx = p.x - t.x, dy = p.y - t.y, dz = p.z - t.z, dw = p.w - t.w;
      acc += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      if (g4) g4[i] = make_float4(grad_scale * dx, grad_scale * dy, grad_scale * dz, grad_scale * dw);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x; i < n_per_image;
         i += (int64_t)gridDim.x * LOSS_THREADS) {
      float p = pred[base + i], t = __ldg(target + base + i);
      if (clip) { p = __saturatef(p); t = __saturatef(t); }
      const float dd = p - t;
      acc += dd * dd;
      if (dL_dpred) dL_dpred[base + i] = grad_scale * dd;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) s += warp_part[w];
    partial[(size_t)img * gridDim.x + blockIdx.x] = s;
  }
}

__global__ void image_mse_final_kernel(const float* __restrict__ partial, int blocks_per_image, int n_images,
                                       float inv_n_per_image, float* __restrict__ mse_per_image,
                                       float* __restrict__ mean_all) {
  pdl_enter();
  // one warp per image, then lane 0 of warp 0 averages the images in order
  __shared__ float per_img[1024];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int img = wid; img < n_images; img += nw) {
    float s = 0.0f;
    for (int b = lane; b < blocks_per_image; b += 32) s += partial[(size_t)img * blocks_per_image + b];
    s = warp_sum(s) * inv_n_per_image;
    if (lane == 0) {
      if (mse_per_image) mse_per_image[img] = s;
      if (img < 1024) per_img[img] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && mean_all) {
    float t = 0.0f;
    for (int i = 0; i < n_images; ++i) t += (i < 1024) ? per_img[i] : mse_per_image[i];
    *mean_all = t / (float)n_images;
  }
}

cudaError_t launch_image_mse(const float* pred, const float* target, int n_images, int64_t n_per_image, int clip,
                             float grad_scale, float* dL_dpred, float* partial, int blocks_per_image,
                             float* mse_per_image, float* mean_all, cudaStream_t s) {
  dim3 grid(blocks_per_image, n_images);
  pdl_launch(image_mse_partial_kernel, grid, LOSS_THREADS, 0, s)(pred, target, n_per_image, clip, grad_scale, dL_dpred, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  pdl_launch(image_mse_final_kernel, 1, 256, 0, s)(partial, blocks_per_image, n_images, 1.0f / (float)n_per_image, mse_per_image,
                                           mean_all);
  return cudaGetLastError();
}

}  // namespace spf
