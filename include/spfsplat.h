/* spfsplat.h -- C ABI of libspfsplat.so (sm_100a), the B200-native replacement for the two
 * native components on SPFSplatV2's decoder hot path:
 *
 *   (1) the external `diff_gauss_pose` rasterizer that
 *       /root/reference/src/model/decoder/cuda_splatting.py:5,105-138,218-249 constructs and calls
 *       once per view (GaussianRasterizationSettings + GaussianRasterizer.__call__), and
 *   (2) the in-tree `curope` extension, entry `rope_2d(tokens, positions, base, fwd)`
 *       (/root/reference/src/model/encoder/backbone/croco/curope/curope.cpp:49-69, kernels.cu:84-108).
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller (torch caching allocator).  The library
 *     never allocates, frees or retains memory, keeps no mutable global state (re-entrant; forward and
 *     backward may come from different host threads), launches only on the stream it is given and
 *     never synchronises the device.
 *   - Return value: 0 = ok, negative = SpfStatus.  spf_last_error() returns a thread-local message.
 *   - Matrices are row-major 4x4 in the ROW-VECTOR convention the reference passes (transposed
 *     world->camera and projection, cuda_splatting.py:88-90): p_view = [m,1] * viewmatrix.
 *   - A call renders B = n_scenes * views_per_scene views; view i reads the Gaussians of scene
 *     i / views_per_scene (this replaces the v-fold `repeat` copies of
 *     /root/reference/src/model/decoder/decoder_splatting_cuda.py:58-64).
 */
#ifndef SPFSPLAT_H_
#define SPFSPLAT_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SPF_API __attribute__((visibility("default")))
#else
#define SPF_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SPF_OK = 0,
  SPF_ERR_BAD_ARG = -1,        /* null pointer, bad size / degree / alignment */
  SPF_ERR_WORKSPACE = -2,      /* a caller-provided buffer is too small */
  SPF_ERR_CUDA = -3,           /* a CUDA runtime call or launch failed (message has the code) */
  SPF_ERR_UNSUPPORTED = -4
} SpfStatus;

enum {
  SPF_FLAG_SH_LAYOUT_CK = 1u << 0, /* sh is [S,P,3,K] (the encoder's native `harmonics` layout,
                                      /root/reference/src/model/types.py:13) instead of [S,P,K,3]
                                      (what cuda_splatting.py:79 materialises) */
  SPF_FLAG_NO_COV_GRAD  = 1u << 1, /* settings.enable_cov_grad == False */
  SPF_FLAG_NO_SH_GRAD   = 1u << 2, /* settings.enable_sh_grad == False */
  SPF_FLAG_NO_TMA       = 1u << 3, /* debugging: stage the slab with plain loads instead of cp.async.bulk */
  SPF_FLAG_QUAT_XYZW    = 1u << 4, /* rotations are (x,y,z,w); default (w,x,y,z) */
  SPF_FLAG_DEPTH_NORMALIZED = 1u << 5, /* reserved (depth / alpha); not implemented */
  SPF_FLAG_BWD_V1       = 1u << 6  /* debugging / cross-check: the first-generation blend backward (back to front,
                                      per-record warp butterfly) instead of the pair-compaction kernel */
};

/* Problem description; mirrors GaussianRasterizationSettings (cuda_splatting.py:105-120). */
typedef struct {
  int32_t n_scenes;          /* S */
  int32_t views_per_scene;   /* v ;  B = S*v */
  int32_t n_gaussians;       /* P per scene */
  int32_t image_height;      /* settings.image_height */
  int32_t image_width;       /* settings.image_width  */
  int32_t sh_degree;         /* settings.sh_degree, 0..4; ignored when colors_precomp is used */
  uint32_t flags;            /* SPF_FLAG_* */
  float   scale_modifier;    /* settings.scale_modifier */
  int64_t dup_capacity;      /* capacity (records) of bucket / slab / dup_grad buffers */
  int32_t ticket;            /* written next to N into state.host_counters (see there) */
  int32_t pair_capacity;     /* pair-log entries (8 B) per warp (state.pair_log); 0 = no pair log */
} SpfRasterDesc;

/* Inputs (forward kwargs of GaussianRasterizer.__call__, cuda_splatting.py:128-138, batched). */
typedef struct {
  const float* means3D;      /* [S,P,3] */
  const float* scales;       /* [S,P,3] */
  const float* rotations;    /* [S,P,4] */
  const float* opacities;    /* [S,P]   */
  const float* shs;          /* [S,P,K,3] or [S,P,3,K]; NULL if colors_precomp */
  const float* colors_precomp; /* [S,P,3] or NULL */
  int32_t      sh_coeffs;    /* K stored per channel in `shs` (>= (sh_degree+1)^2) */
  const float* viewmatrix;   /* [B,16] */
  const float* projmatrix;   /* [B,16]  settings.projmatrix */
  const float* tanfov;       /* [B,2]   settings.tanfovx, tanfovy */
  const float* bg;           /* [B,3]   settings.bg */
  const float* pre_scale;    /* [B] or NULL: means and scales are multiplied by this before use
                                (the 1/near scale-invariance step, cuda_splatting.py:66-74) */
  /* RAW-HEAD INPUT (optional; SURVEY.md 8f rank 2 fused into the projection kernels).  If raw_head is not NULL the
   * Gaussian parameters are taken straight from the encoder head's output rows and scales / rotations / shs (and, with a
   * density channel, opacities) above must be NULL: row = [density logit (if raw_has_density), 3 scale logits,
   * 4 quaternion components, 3 x sh_coeffs SH coefficients ([3][K], the encoder's layout)], mapped inside the kernels
   * exactly like UnifiedGaussianAdapter (gaussian_adapter.py:122-150) and EncoderSPFSplatV2.map_pdf_to_opacity
   * (encoder_spfsplatv2.py:146-159,255-268).  Neither the adapter's outputs nor their gradients ever exist in HBM.
   * Requires n_gaussians % 4 == 0 and 16-byte aligned tensors. */
  const float* raw_head;     /* [S,P,raw_stride] or NULL */
  int32_t      raw_stride;   /* floats per row = raw_has_density + 7 + 3*sh_coeffs */
  int32_t      raw_has_density;
  float        raw_eps;      /* quaternion normalisation eps (1e-8 in the reference) */
  float        opacity_exponent; /* e = 2^x of map_pdf_to_opacity; 1 for the shipped config */
} SpfRasterIn;

/* Caller-allocated intermediates.  Forward fills them; backward reads them.
 * T = ceil(W/16)*ceil(H/16) tiles per view, NB = ceil(P/128) projection blocks per view. */
typedef struct {
  float*    xy;              /* [B,P,2]  projected pixel-space means */
  float*    depth;           /* [B,P]    */
  float*    conic_opacity;   /* [B,P,4]  */
  float*    rgb;             /* [B,P,3]  */
  int32_t*  radii;           /* [B,P]    (also the `radii` return value) */
  int32_t*  tiles_touched;   /* [B,P]    */
  int32_t*  dup_offset;      /* [B,P]    exclusive prefix of tiles_touched (duplicate slots) */
  int32_t*  control;         /* [spf_raster_control_ints(desc)] zeroed by forward; layout private,
                                control[0] = total duplicates N, control[1] = overflow flag */
  uint64_t* bucket;          /* [dup_capacity] (depth_bits<<32 | gaussian) grouped by tile */
  float*    slab;            /* [dup_capacity,12] depth-sorted packed records per tile */
  float*    cullbox;         /* [dup_capacity,4]  per record: conservative pixel bounding box
                                (xmin,xmax,ymin,ymax) of the region where alpha >= 1/255 */
  int32_t*  tile_ranges;     /* [B*T,2] start,end into slab */
  float*    final_T;         /* [B,H,W] */
  int32_t*  n_contrib;       /* [B,H,W] */
  float*    accum;           /* [B,H,W,4] blended sums (rgb, depth) WITHOUT the background term; written by
                                forward, read by backward (16-byte aligned) */
  void*     pair_log;        /* NULL, or [B*T*8, pair_capacity, 8 B]: per-warp log of contributing (pixel,
                                Gaussian) pairs {record index | pixel lane << 25, exp(power)} written by the forward
                                blend, consumed by the backward.  NULL => the backward recomputes them (slower). */
  int32_t*  pair_count;      /* NULL, or [B*T*8]: pairs logged per warp, -1 = capacity exceeded (that tile falls
                                back to recomputation); control[2] = largest count needed by any warp */
  int32_t*  host_counters;   /* NULL, or 2 int32 of device-mapped PINNED HOST memory: the scan kernel
                                stores {N, desc.ticket} there (N first, system-scope fence, then the
                                ticket), so the host learns the duplicate count by polling for its ticket
                                while the remaining kernels are already queued -- no stream sync. */
} SpfRasterState;

typedef struct {
  float* color;              /* [B,3,H,W] */
  float* depth;              /* [B,1,H,W] */
  float* alpha;              /* [B,1,H,W] or NULL */
} SpfRasterOut;

typedef struct {
  const float* dL_dcolor;    /* [B,3,H,W] or NULL */
  const float* dL_ddepth;    /* [B,1,H,W] or NULL */
  const float* dL_dalpha;    /* [B,1,H,W] or NULL */
} SpfRasterGradOut;

typedef struct {
  float* dup_grad;           /* [dup_capacity,12] scratch: per-duplicate 2-D gradients */
  float* pose_partial;       /* [B, NB, 16] scratch */
  float* dL_dmeans3D;        /* [S,P,3] */
  float* dL_dscales;         /* [S,P,3] */
  float* dL_drotations;      /* [S,P,4] */
  float* dL_dopacities;      /* [S,P]   */
  float* dL_dshs;            /* same layout as shs, or NULL */
  float* dL_dcolors;         /* [S,P,3] or NULL */
  float* dL_dviewmatrix;     /* [B,16] */
  float* dL_dmeans2D;        /* [B,P,3] or NULL: screen-space (NDC) mean gradients, the side channel
                                of cuda_splatting.py:97-102 */
  float* dL_draw_head;       /* [S,P,raw_stride]: gradient of the raw head rows; required (and dL_dscales /
                                dL_drotations / dL_dshs ignored, may be NULL) when in.raw_head is given */
} SpfRasterGradIn;

SPF_API int         spf_version(void);
SPF_API const char* spf_last_error(void);

/* Number of int32 the `control` buffer needs. */
SPF_API int64_t spf_raster_control_ints(const SpfRasterDesc* desc);

/* Forward: projection + SH, tile binning, per-tile depth sort + slab pack, alpha blend.
 * If the duplicate count exceeds desc->dup_capacity the overflow flag control[1] is set, the
 * excess duplicates are dropped and the caller must re-run with a larger capacity
 * (control[0] holds the exact count needed). */
SPF_API int spf_raster_forward(const SpfRasterDesc* desc, const SpfRasterIn* in, SpfRasterState* st,
                       SpfRasterOut* out, void* stream);

/* Backward: blend backward (atomic-free, per-duplicate records) + projection/SH backward with
 * camera-pose gradient. */
SPF_API int spf_raster_backward(const SpfRasterDesc* desc, const SpfRasterIn* in, const SpfRasterState* st,
                        const SpfRasterGradOut* gout, SpfRasterGradIn* gin, void* stream);

/* Profiling entry points: run only the stages selected by `stage_mask` (bit i = stage i) so a caller
 * can bracket each kernel with its own CUDA events.  Stages, in order --
 * forward: 0 clear control, 1 project+SH, 2 scan, 3 emit, 4 tile sort+pack, 5 blend forward;
 * backward: 0 blend backward, 1 projection backward, 2 pose reduce.
 * spf_raster_forward / spf_raster_backward are these with every bit set. */
SPF_API int spf_raster_forward_stages(const SpfRasterDesc* desc, const SpfRasterIn* in, SpfRasterState* st,
                              SpfRasterOut* out, uint32_t stage_mask, void* stream);
SPF_API int spf_raster_backward_stages(const SpfRasterDesc* desc, const SpfRasterIn* in, const SpfRasterState* st,
                               const SpfRasterGradOut* gout, SpfRasterGradIn* gin, uint32_t stage_mask,
                               void* stream);

/* Debug / parity helper: unpack gaussian ids (point_list) and (tile<<32|depth_bits) keys of the
 * first n slab records into caller buffers (either may be NULL). */
SPF_API int spf_raster_unpack_sorted(const SpfRasterDesc* desc, const SpfRasterState* st, int64_t n,
                             int32_t* point_list, uint64_t* keys, void* stream);

/* Camera setup of render_cuda (cuda_splatting.py:66-74,84-90; get_fov projection.py:269-283;
 * get_projection_matrix cuda_splatting.py:15-42) for B views in one launch:
 *   extrinsics [B,16] camera-to-world (OpenCV), intrinsics [B,9] normalised, near/far [B]  ->
 *   viewmatrix [B,16] = transpose(inverse(E')), projmatrix [B,16] (transposed), tanfov [B,2],
 *   pre_scale [B] (= 1/near if scale_invariant else 1; E' has its translation scaled by it).
 * Backward maps dL/dviewmatrix to dL/dextrinsics (the camera-pose gradient path). */
SPF_API int spf_camera_forward(int32_t B, int32_t scale_invariant, const float* extrinsics, const float* intrinsics,
                       const float* near, const float* far, float* viewmatrix, float* projmatrix,
                       float* tanfov, float* pre_scale, void* stream);
SPF_API int spf_camera_backward(int32_t B, int32_t scale_invariant, const float* near, const float* viewmatrix,
                        const float* dL_dviewmatrix, float* dL_dextrinsics, void* stream);

/* Fused image losses on the rendered colour: per-image mean squared error of pred vs target, [n_images]
 * images of n_per_image floats each, optionally after clipping both to [0,1] (clip != 0: the PSNR definition,
 * src/evaluation/metrics.py:12-19).  If dL_dpred is not NULL it receives grad_scale * (pred - target) in the same pass
 * (the MSE training loss of src/loss/loss_mse.py:36-51 is weight * mean_all, so grad_scale = 2 * weight / numel).
 * partial: scratch of n_images * spf_image_mse_blocks(n_per_image) floats.  mse_per_image [n_images] and mean_all [1]
 * may each be NULL.  Deterministic (no float atomics). */
SPF_API int spf_image_mse_blocks(int64_t n_per_image);
SPF_API int spf_image_mse(const float* pred, const float* target, int32_t n_images, int64_t n_per_image, int32_t clip,
                  float grad_scale, float* dL_dpred, float* partial, float* mse_per_image, float* mean_all, void* stream);

/* Fused UnifiedGaussianAdapter (src/model/encoder/common/gaussian_adapter.py:122-150): raw [n, 7 + 3*sh_coeffs] ->
 * scales [n,3] = min(0.3, 0.001 softplus), rotations [n,4] = q / (|q| + eps), harmonics [n,3,sh_coeffs] = raw * sh_mask
 * (sh_mask per degree as gaussian_adapter.py:42-48).  The [n,3,3] covariances the reference also builds are never read
 * by the decoder and are not produced.  Backward: any of the three upstream gradients may be NULL (= zero). */
SPF_API int spf_adapter_forward(const float* raw, int64_t n, int32_t sh_coeffs, float eps, float* scales, float* rotations,
                        float* harmonics, void* stream);
SPF_API int spf_adapter_backward(const float* raw, const float* dL_dscales, const float* dL_drotations,
                         const float* dL_dharmonics, int64_t n, int32_t sh_coeffs, float eps, float* dL_draw, void* stream);

/* The encoder head's whole post-processing in one pass: rows [n, 1 + 7 + 3*sh_coeffs] = the 83-channel Gaussian-head output
 * (density logit first, /root/reference/src/model/encoder/encoder_spfsplatv2.py:255-268) -> opacities [n] by
 * EncoderSPFSplatV2.map_pdf_to_opacity (:146-159: p = sigmoid(logit), 0.5 * (1 - (1-p)^e + p^(1/e)); `exponent` e = 2^x with
 * x from the caller's warm-up schedule, 1 for the shipped config spfsplatv2.yaml:6-9) plus the adapter's scales / rotations /
 * harmonics as above.  Backward: any upstream gradient may be NULL (= zero). */
SPF_API int spf_head_forward(const float* raw, int64_t n, int32_t sh_coeffs, float eps, float exponent, float* opacities,
                     float* scales, float* rotations, float* harmonics, void* stream);
SPF_API int spf_head_backward(const float* raw, const float* dL_dopacities, const float* dL_dscales, const float* dL_drotations,
                      const float* dL_dharmonics, int64_t n, int32_t sh_coeffs, float eps, float exponent, float* dL_draw,
                      void* stream);

/* PLY vertex rows of a Gaussian scene, packed on the device.  Replaces the numpy / scipy body of export_ply
 * (/root/reference/src/model/ply_export.py:76-141): rows [n,17] = x y z (R (mean - shift) / scale_factor), nx ny nz (0),
 * f_dc_0..2 (the DC band of harmonics [n,3,sh_coeffs]), opacity (as stored), scale_0..2 (log(scale / scale_factor)),
 * rot_0..3 ((w,x,y,z) of quat(R * matrix(q)), q given scalar-last like the reference's rotations).
 * params (device, 13 floats) = R row-major (9), shift (3), scale_factor (1). */
SPF_API int spf_ply_pack(const float* means, const float* scales, const float* rotations_xyzw, const float* harmonics,
                 const float* opacities, const float* params, int64_t n, int32_t sh_coeffs, float* rows, void* stream);

/* 2-D RoPE, in place.  Replaces rope_2d (curope.cpp:49-65).  tokens: [B,N,H,D] view with
 * stride(3)==1, stride(2)==D (kernels.cu:91); positions int64 [B,N,2] contiguous.
 * dtype: 0 = fp32, 1 = fp16, 2 = bf16, 3 = fp64 (the floating types the reference dispatches, kernels.cu:101, plus bf16;
 * fp64 values are rotated in fp32 like the reference does through its float staging buffer, kernels.cu:30,66).
 * fwd = +F0 forward, -F0 backward. */
SPF_API int spf_rope2d(void* tokens, const int64_t* positions, int32_t B, int32_t N, int32_t H, int32_t D,
               int64_t stride_b, int64_t stride_n, int32_t dtype, float base, float fwd,
               void* stream);

/* The two rope_2d calls of a self-attention block (croco/blocks.py:102-104: q = rope(q, xpos); k = rope(k, xpos))
 * in ONE launch: q and k are same-shape, same-stride views (of the fused qkv tensor, blocks.py:97-98) sharing
 * `positions`; cos/sin are evaluated once per (token, frequency) and applied to both. */
SPF_API int spf_rope2d_qk(void* q, void* k, const int64_t* positions, int32_t B, int32_t N, int32_t H, int32_t D,
                  int64_t stride_b, int64_t stride_n, int32_t dtype, float base, float fwd,
                  void* stream);

/* In-switch (NVLS multicast) sum all-reduce of one fp32 gradient bucket, in place.  Replaces the NCCL all-reduce torch DDP
 * issues for the replicated parameters' gradients (src/main.py:141-145).  multicast_bucket: the MULTICAST address of a
 * symmetric-memory bucket of numel floats (numel % 4 == 0, 16-byte aligned), bound on all `world` ranks; this rank reduces
 * and re-broadcasts slice `rank` of `world`.  n_blocks: CTAs to spend.  The caller orders the launch between two
 * cross-rank barriers on the same stream (every bucket final before; every slice broadcast after). */
SPF_API int spf_multimem_allreduce_f32(float* multicast_bucket, int64_t numel, int32_t rank, int32_t world, int32_t n_blocks,
                               void* stream);

/* The same reduction with both cross-rank barriers INSIDE the kernel (no separate barrier launches): signal_pads is a
 * DEVICE array of `world` pointers, entry p = rank p's symmetric-memory signal pad (32-bit words, zero when idle) as mapped
 * into this rank's address space; flags [pad_word_offset, pad_word_offset + 2 * n_blocks * world) of every pad are used,
 * pad_words = words available per pad.  n_blocks must be the same on every rank. */
SPF_API int spf_multimem_allreduce_f32_fused(float* multicast_bucket, int64_t numel, int32_t rank, int32_t world, int32_t n_blocks,
                                     void* const* signal_pads, int32_t pad_word_offset, int32_t pad_words, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPFSPLAT_H_ */
