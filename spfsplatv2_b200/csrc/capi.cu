// extern "C" entry points of libspfsplat.so (see include/spfsplat.h).  Argument validation,
// launch sequencing on the caller's stream, thread-local error text.  No allocation, no
// synchronisation, no global mutable state.
#include <stdarg.h>
#include <stdio.h>

#include <nvtx3/nvToolsExt.h>

#include "spf_kernels.h"
#include "spf_math.h"

namespace {
thread_local char g_err[512] = "";

// SPF_NVTX=1: one NVTX range per launch stage (spf/project_forward, spf/scan, ... spf/pose_reduce) around the host-side
// enqueue, so nsys / ncu --nvtx timelines show the stages by name; off by default (no overhead).
bool nvtx_on() {
  static const bool on = [] { const char* e = getenv("SPF_NVTX"); return e && e[0] == '1'; }();
  return on;
}
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) : on(nvtx_on()) { if (on) nvtxRangePushA(name); }
  ~NvtxRange() { if (on) nvtxRangePop(); }
};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(SPF_ERR_CUDA, "%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
}

bool make_dims(const SpfRasterDesc* desc, int sh_coeffs, bool use_sh, spf::Dims& d) {
  d.S = desc->n_scenes;
  d.v = desc->views_per_scene;
  d.B = d.S * d.v;
  d.P = desc->n_gaussians;
  d.H = desc->image_height;
  d.W = desc->image_width;
  d.gx = (d.W + 15) / 16;
  d.gy = (d.H + 15) / 16;
  d.T = d.gx * d.gy;
  d.NB = (d.P + spf::PROJ_THREADS - 1) / spf::PROJ_THREADS;
  d.deg = desc->sh_degree;
  d.K = (d.deg + 1) * (d.deg + 1);
  d.flags = desc->flags;
  d.mod = desc->scale_modifier;
  d.cap = desc->dup_capacity;
  d.ticket = desc->ticket;
  d.pair_cap = desc->pair_capacity;
  (void)sh_coeffs; (void)use_sh;
  return true;
}

int check_desc(const SpfRasterDesc* desc) {
  if (!desc) return fail(SPF_ERR_BAD_ARG, "desc is NULL");
  if (desc->n_scenes < 1 || desc->views_per_scene < 1 || desc->n_gaussians < 1)
    return fail(SPF_ERR_BAD_ARG, "n_scenes, views_per_scene and n_gaussians must be >= 1");
  if (desc->image_height < 1 || desc->image_width < 1) return fail(SPF_ERR_BAD_ARG, "bad image size");
  if (desc->sh_degree < 0 || desc->sh_degree > 4) return fail(SPF_ERR_BAD_ARG, "sh_degree must be in 0..4");
  if (desc->dup_capacity < 1 || desc->dup_capacity > 0x7fffffffLL)
    return fail(SPF_ERR_BAD_ARG, "dup_capacity must be in 1..2^31-1");
  if ((int64_t)desc->n_scenes * desc->views_per_scene > 65535)
    return fail(SPF_ERR_UNSUPPORTED, "more than 65535 views per call");
  if (desc->pair_capacity < 0 || (desc->pair_capacity & 31))
    return fail(SPF_ERR_BAD_ARG, "pair_capacity must be a non-negative multiple of 32");
  if (desc->flags & SPF_FLAG_DEPTH_NORMALIZED) return fail(SPF_ERR_UNSUPPORTED, "normalised depth not implemented");
  return 0;
}

int check_in(const SpfRasterDesc* desc, const SpfRasterIn* in) {
  if (!in) return fail(SPF_ERR_BAD_ARG, "in is NULL");
  if (in->raw_head) {
    // raw-head input: the adapter's tensors must not be given as well
    if (!in->means3D) return fail(SPF_ERR_BAD_ARG, "means3D must be provided");
    if (in->scales || in->rotations || in->shs || in->colors_precomp)
      return fail(SPF_ERR_BAD_ARG, "raw_head given: scales / rotations / shs / colors_precomp must be NULL");
    if ((in->opacities == nullptr) != (in->raw_has_density != 0))
      return fail(SPF_ERR_BAD_ARG, "raw_head: opacities must be NULL exactly when the rows carry a density logit");
    const int K = (desc->sh_degree + 1) * (desc->sh_degree + 1);
    if (in->sh_coeffs < K || in->sh_coeffs > spf::MAX_SH_COEFFS) return fail(SPF_ERR_BAD_ARG, "raw_head: bad sh_coeffs");
    if (in->raw_stride != (in->raw_has_density ? 1 : 0) + 7 + 3 * in->sh_coeffs)
      return fail(SPF_ERR_BAD_ARG, "raw_stride must be raw_has_density + 7 + 3 * sh_coeffs");
    if (!(in->opacity_exponent > 0.0f)) return fail(SPF_ERR_BAD_ARG, "opacity_exponent must be positive");
    if (desc->n_gaussians % 4 != 0)
      return fail(SPF_ERR_UNSUPPORTED, "raw_head needs n_gaussians % 4 == 0 (16-byte aligned rows for the bulk copies)");
    if ((reinterpret_cast<uintptr_t>(in->raw_head) | reinterpret_cast<uintptr_t>(in->means3D) |
         reinterpret_cast<uintptr_t>(in->opacities)) & 15)
      return fail(SPF_ERR_BAD_ARG, "raw_head / means3D / opacities must be 16-byte aligned");
    if (desc->flags & (SPF_FLAG_NO_TMA | SPF_FLAG_QUAT_XYZW))
      return fail(SPF_ERR_UNSUPPORTED, "raw_head does not combine with SPF_FLAG_NO_TMA / SPF_FLAG_QUAT_XYZW");
    if (!in->viewmatrix || !in->projmatrix || !in->tanfov || !in->bg)
      return fail(SPF_ERR_BAD_ARG, "viewmatrix / projmatrix / tanfov / bg must be provided");
    return 0;
  }
  if (!in->means3D || !in->scales || !in->rotations || !in->opacities)
    return fail(SPF_ERR_BAD_ARG, "means3D / scales / rotations / opacities must be provided");
  if ((in->shs != nullptr) == (in->colors_precomp != nullptr))
    return fail(SPF_ERR_BAD_ARG, "Please provide exactly one of either SHs or precomputed colors!");
  if (in->shs) {
    const int K = (desc->sh_degree + 1) * (desc->sh_degree + 1);
    if (in->sh_coeffs < K) return fail(SPF_ERR_BAD_ARG, "shs holds %d coefficients, degree %d needs %d",
                                       in->sh_coeffs, desc->sh_degree, K);
    if (in->sh_coeffs > spf::MAX_SH_COEFFS)
      return fail(SPF_ERR_UNSUPPORTED, "at most %d SH coefficients per channel", spf::MAX_SH_COEFFS);
  }
  if (!in->viewmatrix || !in->projmatrix || !in->tanfov || !in->bg)
    return fail(SPF_ERR_BAD_ARG, "viewmatrix / projmatrix / tanfov / bg must be provided");
  if (reinterpret_cast<uintptr_t>(in->rotations) & 15) return fail(SPF_ERR_BAD_ARG, "rotations must be 16-byte aligned");
  return 0;
}

int check_state(const SpfRasterState* st) {
  if (!st) return fail(SPF_ERR_BAD_ARG, "state is NULL");
  if (!st->xy || !st->depth || !st->conic_opacity || !st->rgb || !st->radii || !st->tiles_touched ||
      !st->dup_offset || !st->control || !st->bucket || !st->slab || !st->cullbox || !st->tile_ranges || !st->final_T ||
      !st->n_contrib || !st->accum)
    return fail(SPF_ERR_BAD_ARG, "every SpfRasterState buffer must be provided");
  if ((reinterpret_cast<uintptr_t>(st->slab) & 15) || (reinterpret_cast<uintptr_t>(st->accum) & 15) || (reinterpret_cast<uintptr_t>(st->cullbox) & 15) || (reinterpret_cast<uintptr_t>(st->conic_opacity) & 15) ||
      (reinterpret_cast<uintptr_t>(st->xy) & 7) || (reinterpret_cast<uintptr_t>(st->bucket) & 7))
    return fail(SPF_ERR_BAD_ARG, "state buffers are not sufficiently aligned (slab/conic 16 B, xy/bucket 8 B)");
  return 0;
}
}  // namespace

extern "C" {

int spf_version(void) { return 100; }

const char* spf_last_error(void) { return g_err; }

int64_t spf_raster_control_ints(const SpfRasterDesc* desc) {
  if (check_desc(desc)) return -1;
  spf::Dims d;
  make_dims(desc, 0, false, d);
  return spf::control_layout(d.B, d.T, d.NB).total;
}

int spf_raster_forward(const SpfRasterDesc* desc, const SpfRasterIn* in, SpfRasterState* st,
                       SpfRasterOut* out, void* stream) {
  return spf_raster_forward_stages(desc, in, st, out, 0xffffffffu, stream);
}

int spf_raster_forward_stages(const SpfRasterDesc* desc, const SpfRasterIn* in, SpfRasterState* st,
                              SpfRasterOut* out, uint32_t mask, void* stream) {
  int rc;
  if ((rc = check_desc(desc)) || (rc = check_in(desc, in)) || (rc = check_state(st))) return rc;
  if (!out || !out->color || !out->depth) return fail(SPF_ERR_BAD_ARG, "out.color / out.depth must be provided");
  spf::Dims d;
  make_dims(desc, in->sh_coeffs, in->shs != nullptr, d);
  const spf::ControlLayout cl = spf::control_layout(d.B, d.T, d.NB);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if ((mask & 1u) && (e = cudaMemsetAsync(st->control, 0, (size_t)cl.total * sizeof(int32_t), s)) != cudaSuccess)
    return cuda_fail(e, "memset(control)");
#define SPF_STAGE(bit, name, call)                                   \
  if (mask & (bit)) {                                                \
    NvtxRange r_("spf/" name);                                       \
    if ((e = (call)) != cudaSuccess) return cuda_fail(e, name);      \
  }
  SPF_STAGE(2u, "project_forward", spf::launch_project_forward(d, *in, *st, cl, s))
  SPF_STAGE(4u, "scan", spf::launch_scan(d, *st, cl, s))
  SPF_STAGE(8u, "emit", spf::launch_emit(d, *st, cl, s))
  SPF_STAGE(16u, "tile_sort_pack", spf::launch_tile_sort_pack(d, *st, cl, s))
  SPF_STAGE(32u, "blend_forward", spf::launch_blend_forward(d, *in, *st, *out, s))
  g_err[0] = 0;
  return SPF_OK;
}

int spf_raster_backward(const SpfRasterDesc* desc, const SpfRasterIn* in, const SpfRasterState* st,
                        const SpfRasterGradOut* gout, SpfRasterGradIn* gin, void* stream) {
  return spf_raster_backward_stages(desc, in, st, gout, gin, 0xffffffffu, stream);
}

int spf_raster_backward_stages(const SpfRasterDesc* desc, const SpfRasterIn* in, const SpfRasterState* st,
                               const SpfRasterGradOut* gout, SpfRasterGradIn* gin, uint32_t mask, void* stream) {
  int rc;
  if ((rc = check_desc(desc)) || (rc = check_in(desc, in)) || (rc = check_state(st))) return rc;
  if (!gout || !gin) return fail(SPF_ERR_BAD_ARG, "gradient structs must be provided");
  if (in->raw_head) {
    if (!gin->dup_grad || !gin->pose_partial || !gin->dL_dmeans3D || !gin->dL_dviewmatrix || !gin->dL_draw_head)
      return fail(SPF_ERR_BAD_ARG, "raw_head: dup_grad, pose_partial, dL_dmeans3D, dL_dviewmatrix and dL_draw_head are required");
    if (!in->raw_has_density && !gin->dL_dopacities) return fail(SPF_ERR_BAD_ARG, "raw_head without density: dL_dopacities is required");
    if ((reinterpret_cast<uintptr_t>(gin->dup_grad) | reinterpret_cast<uintptr_t>(gin->dL_draw_head)) & 15)
      return fail(SPF_ERR_BAD_ARG, "dup_grad / dL_draw_head must be 16-byte aligned");
  } else {
    if (!gin->dup_grad || !gin->pose_partial || !gin->dL_dmeans3D || !gin->dL_dscales || !gin->dL_drotations ||
        !gin->dL_dopacities || !gin->dL_dviewmatrix)
      return fail(SPF_ERR_BAD_ARG, "dup_grad, pose_partial and the dL_d{means3D,scales,rotations,opacities,viewmatrix} buffers are required");
    if (in->shs && !gin->dL_dshs) return fail(SPF_ERR_BAD_ARG, "dL_dshs is required when shs are given");
    if ((reinterpret_cast<uintptr_t>(gin->dup_grad) & 15) || (reinterpret_cast<uintptr_t>(gin->dL_drotations) & 15))
      return fail(SPF_ERR_BAD_ARG, "dup_grad / dL_drotations must be 16-byte aligned");
  }
  spf::Dims d;
  make_dims(desc, in->sh_coeffs, in->shs != nullptr, d);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  SPF_STAGE(1u, "blend_backward", spf::launch_blend_backward(d, *in, *st, *gout, *gin, s))
  SPF_STAGE(2u, "project_backward", spf::launch_project_backward(d, *in, *st, *gin, s))
  SPF_STAGE(4u, "pose_reduce", spf::launch_pose_reduce(d, *in, *gin, s))
  g_err[0] = 0;
  return SPF_OK;
}

int spf_raster_unpack_sorted(const SpfRasterDesc* desc, const SpfRasterState* st, int64_t n,
                             int32_t* point_list, uint64_t* keys, void* stream) {
  int rc;
  if ((rc = check_desc(desc)) || (rc = check_state(st))) return rc;
  if (n < 0 || n > desc->dup_capacity) return fail(SPF_ERR_BAD_ARG, "n out of range");
  spf::Dims d;
  make_dims(desc, 0, false, d);
  const spf::ControlLayout cl = spf::control_layout(d.B, d.T, d.NB);
  cudaError_t e = spf::launch_unpack_sorted(d, *st, n, point_list, keys, cl, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "unpack_sorted");
  return SPF_OK;
}

int spf_camera_forward(int32_t B, int32_t scale_invariant, const float* extrinsics, const float* intrinsics,
                       const float* near, const float* far, float* viewmatrix, float* projmatrix, float* tanfov,
                       float* pre_scale, void* stream) {
  if (B < 1) return fail(SPF_ERR_BAD_ARG, "B must be >= 1");
  if (!extrinsics || !intrinsics || !near || !far || !viewmatrix || !projmatrix || !tanfov || !pre_scale)
    return fail(SPF_ERR_BAD_ARG, "spf_camera_forward: NULL pointer");
  cudaError_t e = spf::launch_camera_forward(B, scale_invariant, extrinsics, intrinsics, near, far, viewmatrix,
                                             projmatrix, tanfov, pre_scale, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "camera_forward");
  return SPF_OK;
}

int spf_camera_backward(int32_t B, int32_t scale_invariant, const float* near, const float* viewmatrix,
                        const float* dL_dviewmatrix, float* dL_dextrinsics, void* stream) {
  if (B < 1) return fail(SPF_ERR_BAD_ARG, "B must be >= 1");
  if (!near || !viewmatrix || !dL_dviewmatrix || !dL_dextrinsics)
    return fail(SPF_ERR_BAD_ARG, "spf_camera_backward: NULL pointer");
  cudaError_t e = spf::launch_camera_backward(B, scale_invariant, near, viewmatrix, dL_dviewmatrix, dL_dextrinsics,
                                              static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "camera_backward");
  return SPF_OK;
}

int spf_image_mse_blocks(int64_t n_per_image) {
  int64_t b = (n_per_image / 4 + 4 * 256 - 1) / (4 * 256);      // ~4 float4 per thread
  if (b < 1) b = 1;
  if (b > 256) b = 256;
  return (int)b;
}

int spf_image_mse(const float* pred, const float* target, int32_t n_images, int64_t n_per_image, int32_t clip,
                  float grad_scale, float* dL_dpred, float* partial, float* mse_per_image, float* mean_all, void* stream) {
  if (!pred || !target || !partial) return fail(SPF_ERR_BAD_ARG, "spf_image_mse: pred / target / partial are NULL");
  if (n_images < 1 || n_images > 65535 || n_per_image < 1) return fail(SPF_ERR_BAD_ARG, "spf_image_mse: bad sizes");
  if (!mse_per_image && n_images > 1024) return fail(SPF_ERR_BAD_ARG, "spf_image_mse: mse_per_image is required above 1024 images");
  cudaError_t e = spf::launch_image_mse(pred, target, n_images, n_per_image, clip, grad_scale, dL_dpred, partial,
                                        spf_image_mse_blocks(n_per_image), mse_per_image, mean_all,
                                        static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "image_mse");
  return SPF_OK;
}

int spf_adapter_forward(const float* raw, int64_t n, int32_t sh_coeffs, float eps, float* scales, float* rotations,
                        float* harmonics, void* stream) {
  if (!raw || !scales || !rotations || !harmonics) return fail(SPF_ERR_BAD_ARG, "spf_adapter_forward: NULL pointer");
  if (n < 0 || sh_coeffs < 1 || sh_coeffs > spf::MAX_SH_COEFFS) return fail(SPF_ERR_BAD_ARG, "spf_adapter_forward: bad sizes");
  cudaError_t e = spf::launch_adapter_forward(raw, n, sh_coeffs, eps, 0, 1.0f, scales, rotations, harmonics, nullptr,
                                              static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "adapter_forward");
  return SPF_OK;
}

int spf_adapter_backward(const float* raw, const float* dL_dscales, const float* dL_drotations, const float* dL_dharmonics,
                         int64_t n, int32_t sh_coeffs, float eps, float* dL_draw, void* stream) {
  if (!raw || !dL_draw) return fail(SPF_ERR_BAD_ARG, "spf_adapter_backward: NULL pointer");
  if (n < 0 || sh_coeffs < 1 || sh_coeffs > spf::MAX_SH_COEFFS) return fail(SPF_ERR_BAD_ARG, "spf_adapter_backward: bad sizes");
  cudaError_t e = spf::launch_adapter_backward(raw, dL_dscales, dL_drotations, dL_dharmonics, nullptr, n, sh_coeffs, eps, 0, 1.0f,
                                               dL_draw, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "adapter_backward");
  return SPF_OK;
}

int spf_head_forward(const float* raw, int64_t n, int32_t sh_coeffs, float eps, float exponent, float* opacities, float* scales,
                     float* rotations, float* harmonics, void* stream) {
  if (!raw || !opacities || !scales || !rotations || !harmonics) return fail(SPF_ERR_BAD_ARG, "spf_head_forward: NULL pointer");
  if (n < 0 || sh_coeffs < 1 || sh_coeffs > spf::MAX_SH_COEFFS) return fail(SPF_ERR_BAD_ARG, "spf_head_forward: bad sizes");
  if (!(exponent > 0.0f)) return fail(SPF_ERR_BAD_ARG, "spf_head_forward: exponent must be positive");
  cudaError_t e = spf::launch_adapter_forward(raw, n, sh_coeffs, eps, 1, exponent, scales, rotations, harmonics, opacities,
                                              static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "head_forward");
  return SPF_OK;
}

int spf_head_backward(const float* raw, const float* dL_dopacities, const float* dL_dscales, const float* dL_drotations,
                      const float* dL_dharmonics, int64_t n, int32_t sh_coeffs, float eps, float exponent, float* dL_draw,
                      void* stream) {
  if (!raw || !dL_draw) return fail(SPF_ERR_BAD_ARG, "spf_head_backward: NULL pointer");
  if (n < 0 || sh_coeffs < 1 || sh_coeffs > spf::MAX_SH_COEFFS) return fail(SPF_ERR_BAD_ARG, "spf_head_backward: bad sizes");
  if (!(exponent > 0.0f)) return fail(SPF_ERR_BAD_ARG, "spf_head_backward: exponent must be positive");
  cudaError_t e = spf::launch_adapter_backward(raw, dL_dscales, dL_drotations, dL_dharmonics, dL_dopacities, n, sh_coeffs, eps, 1,
                                               exponent, dL_draw, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "head_backward");
  return SPF_OK;
}

int spf_ply_pack(const float* means, const float* scales, const float* rotations_xyzw, const float* harmonics,
                 const float* opacities, const float* params, int64_t n, int32_t sh_coeffs, float* rows, void* stream) {
  if (!means || !scales || !rotations_xyzw || !harmonics || !opacities || !params || !rows)
    return fail(SPF_ERR_BAD_ARG, "spf_ply_pack: NULL pointer");
  if (n < 0 || sh_coeffs < 1) return fail(SPF_ERR_BAD_ARG, "spf_ply_pack: bad sizes");
  cudaError_t e = spf::launch_ply_pack(means, scales, rotations_xyzw, harmonics, opacities, params, n, sh_coeffs, rows,
                                       static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "ply_pack");
  return SPF_OK;
}

int spf_multimem_allreduce_f32(float* multicast_bucket, int64_t numel, int32_t rank, int32_t world, int32_t n_blocks,
                               void* stream) {
  if (!multicast_bucket) return fail(SPF_ERR_BAD_ARG, "spf_multimem_allreduce_f32: multicast address is NULL");
  if ((reinterpret_cast<uintptr_t>(multicast_bucket) & 15) != 0 || numel < 0 || (numel & 3) != 0)
    return fail(SPF_ERR_BAD_ARG, "spf_multimem_allreduce_f32: bucket must be 16-byte aligned with numel % 4 == 0");
  if (world < 1 || rank < 0 || rank >= world) return fail(SPF_ERR_BAD_ARG, "spf_multimem_allreduce_f32: bad rank / world");
  if (n_blocks < 1 || n_blocks > 1024) return fail(SPF_ERR_BAD_ARG, "spf_multimem_allreduce_f32: n_blocks must be 1..1024");
  cudaError_t e = spf::launch_multimem_allreduce_f32(multicast_bucket, numel, rank, world, n_blocks,
                                                     static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "multimem_allreduce_f32");
  return SPF_OK;
}

int spf_multimem_allreduce_f32_fused(float* multicast_bucket, int64_t numel, int32_t rank, int32_t world, int32_t n_blocks,
                                     void* const* signal_pads, int32_t pad_word_offset, int32_t pad_words, void* stream) {
  if (!multicast_bucket || !signal_pads) return fail(SPF_ERR_BAD_ARG, "spf_multimem_allreduce_f32_fused: NULL bucket / signal pads");
  if (numel < 0 || (numel & 3) || (reinterpret_cast<uintptr_t>(multicast_bucket) & 15))
    return fail(SPF_ERR_BAD_ARG, "spf_multimem_allreduce_f32_fused: bucket must be 16-byte aligned with numel % 4 == 0");
  if (world < 1 || world > 1024 || rank < 0 || rank >= world) return fail(SPF_ERR_BAD_ARG, "spf_multimem_allreduce_f32_fused: bad rank / world");
  if (n_blocks < 1 || n_blocks > spf::sm_count()) return fail(SPF_ERR_BAD_ARG, "spf_multimem_allreduce_f32_fused: n_blocks must be 1..#SMs (the CTAs of all ranks must be co-resident)");
  if (pad_word_offset < 0 || (int64_t)pad_word_offset + 2LL * n_blocks * world > pad_words)
    return fail(SPF_ERR_WORKSPACE, "spf_multimem_allreduce_f32_fused: signal pad too small for 2 * n_blocks * world flags");
  cudaError_t e = spf::launch_multimem_allreduce_f32_fused(multicast_bucket, numel, rank, world, n_blocks,
                                                           reinterpret_cast<uint32_t* const*>(signal_pads), pad_word_offset,
                                                           static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "multimem_allreduce_f32_fused");
  return SPF_OK;
}

static int rope_common(void* q, void* k, const int64_t* positions, int32_t B, int32_t N, int32_t H, int32_t D,
                       int64_t stride_b, int64_t stride_n, int32_t dtype, float base, float fwd, void* stream) {
  if (!q || !positions) return fail(SPF_ERR_BAD_ARG, "tokens / positions are NULL");
  if (B < 0 || N < 0 || H < 0 || D < 0) return fail(SPF_ERR_BAD_ARG, "negative size");
  if (D % 4 != 0) return fail(SPF_ERR_BAD_ARG, "token dim must be multiple of 4");
  if (dtype < 0 || dtype > 3) return fail(SPF_ERR_BAD_ARG, "dtype must be 0 (fp32), 1 (fp16), 2 (bf16) or 3 (fp64)");
  cudaError_t e = spf::launch_rope2d(q, k, positions, B, N, H, D, stride_b, stride_n, dtype, base, fwd,
                                     static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "rope2d");
  return SPF_OK;
}

int spf_rope2d(void* tokens, const int64_t* positions, int32_t B, int32_t N, int32_t H, int32_t D,
               int64_t stride_b, int64_t stride_n, int32_t dtype, float base, float fwd, void* stream) {
  return rope_common(tokens, nullptr, positions, B, N, H, D, stride_b, stride_n, dtype, base, fwd, stream);
}

int spf_rope2d_qk(void* q, void* k, const int64_t* positions, int32_t B, int32_t N, int32_t H, int32_t D,
                  int64_t stride_b, int64_t stride_n, int32_t dtype, float base, float fwd, void* stream) {
  if (!k) return fail(SPF_ERR_BAD_ARG, "k is NULL");
  return rope_common(q, k, positions, B, N, H, D, stride_b, stride_n, dtype, base, fwd, stream);
}

}  // extern "C"
