// Internal launcher declarations shared by capi.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <utility>

#include "../../include/spfsplat.h"

namespace spf {

constexpr int PROJ_THREADS = 128;   // Gaussians per projection block
constexpr int TILE_THREADS = 256;   // 16x16 pixels

// control buffer layout (int32 words)
struct ControlLayout {
  int64_t n_total, overflow, pair_max, need_fallback, tile_count, tile_cursor, tile_start, block_sum, block_off, pose_done, total;
};
inline ControlLayout control_layout(int B, int T, int NB) {
  ControlLayout c;
  c.n_total = 0;
  c.overflow = 1;
  c.pair_max = 2;        // largest per-warp pair-log count needed (blend forward)
  c.need_fallback = 3;   // some tile's pair log is unusable: the recomputing blend backward has work to do
  int64_t o = 4;
  c.tile_count = o;  o += (int64_t)B * T;
  c.tile_cursor = o; o += (int64_t)B * T;
  c.tile_start = o;  o += (int64_t)B * T + 1;
  c.block_sum = o;   o += (int64_t)B * NB;
  c.block_off = o;   o += (int64_t)B * NB + 1;
  c.pose_done = o;   o += (int64_t)B;          // per-view counters of finished projection-backward blocks
  c.total = (o + 3) & ~int64_t(3);
  return c;
}

struct Dims {
  int S, v, B, P, H, W, gx, gy, T, NB, K, deg;
  uint32_t flags;
  float mod;
  int64_t cap;
  int ticket;
  int pair_cap;   // pair-log records per warp (0 = no log)
};

// Programmatic dependent launch (PDL, sm_90+): every kernel of the raster path is launched with the
// programmatic-stream-serialization attribute and starts with `pdl_enter()` (spf_device.cuh: griddepcontrol.wait, then
// griddepcontrol.launch_dependents).  The next kernel's CTAs are scheduled into SM slots as they free up during the
// current kernel's last wave and park at the wait until the current kernel has completed and flushed, so the launch
// latency and the CTA ramp of each of the ~13 kernels of a step disappear from the critical path (also inside a
// captured CUDA graph, where the dependency becomes a programmatic edge).  Every kernel in the chain executes the wait,
// so completion stays transitive.  SPF_PDL=0 falls back to plain stream order.
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("SPF_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

// Number of SMs of the current device (148 on B200), read once per device; sizes the persistent / grid-stride launches.
inline int sm_count() {
  static thread_local int cached_dev = -1, cached_sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cached_sms;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached_sms = n;
    cached_dev = dev;
  }
  return cached_sms;
}

template <typename K>
struct PdlLaunch {
  K kernel;
  dim3 grid, block;
  size_t smem;
  cudaStream_t stream;
  template <typename... A>
  void operator()(A&&... args) const {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);   // callers read cudaGetLastError()
  }
};
template <typename K>
inline PdlLaunch<K> pdl_launch(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
  return PdlLaunch<K>{kernel, grid, block, smem, stream};
}

cudaError_t launch_project_forward(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st,
                                   const ControlLayout& cl, cudaStream_t s);
cudaError_t launch_scan(const Dims& d, const SpfRasterState& st, const ControlLayout& cl, cudaStream_t s);
cudaError_t launch_emit(const Dims& d, const SpfRasterState& st, const ControlLayout& cl, cudaStream_t s);
cudaError_t launch_tile_sort_pack(const Dims& d, const SpfRasterState& st, const ControlLayout& cl,
                                  cudaStream_t s);
cudaError_t launch_blend_forward(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st,
                                 const SpfRasterOut& out, cudaStream_t s);
cudaError_t launch_blend_backward(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st,
                                  const SpfRasterGradOut& gout, const SpfRasterGradIn& gin, cudaStream_t s);
cudaError_t launch_project_backward(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st,
                                    const SpfRasterGradIn& gin, cudaStream_t s);
cudaError_t launch_pose_reduce(const Dims& d, const SpfRasterIn& in, const SpfRasterGradIn& gin, cudaStream_t s);
cudaError_t launch_unpack_sorted(const Dims& d, const SpfRasterState& st, int64_t n, int32_t* point_list,
                                 uint64_t* keys, const ControlLayout& cl, cudaStream_t s);
cudaError_t launch_camera_forward(int B, int scale_invariant, const float* ext, const float* intr,
                                  const float* near, const float* far, float* view, float* proj, float* tanfov,
                                  float* pre_scale, cudaStream_t s);
cudaError_t launch_camera_backward(int B, int scale_invariant, const float* near, const float* view,
                                   const float* d_view, float* d_ext, cudaStream_t s);
cudaError_t launch_image_mse(const float* pred, const float* target, int n_images, int64_t n_per_image, int clip,
                             float grad_scale, float* dL_dpred, float* partial, int blocks_per_image,
                             float* mse_per_image, float* mean_all, cudaStream_t s);
cudaError_t launch_multimem_allreduce_f32(float* mc, int64_t numel, int rank, int world, int n_blocks, cudaStream_t s);
cudaError_t launch_multimem_allreduce_f32_fused(float* mc, int64_t numel, int rank, int world, int n_blocks,
                                                uint32_t* const* signal_pads, int pad_word_offset, cudaStream_t s);
cudaError_t launch_ply_pack(const float* means, const float* scales, const float* rots, const float* harmonics,
                            const float* opac, const float* params, int64_t n, int sh_coeffs, float* out, cudaStream_t s);
cudaError_t launch_adapter_forward(const float* raw, int64_t n, int K, float eps, int dens, float exponent, float* scales,
                                   float* rots, float* sh, float* opac, cudaStream_t s);
cudaError_t launch_adapter_backward(const float* raw, const float* d_scales, const float* d_rots, const float* d_sh,
                                    const float* d_opac, int64_t n, int K, float eps, int dens, float exponent, float* d_raw,
                                    cudaStream_t s);
cudaError_t launch_rope2d(void* tokens, void* tokens2, const int64_t* pos, int B, int N, int H, int D, int64_t sb,
                          int64_t sn, int dtype, float base, float fwd, cudaStream_t s);

}  // namespace spf
