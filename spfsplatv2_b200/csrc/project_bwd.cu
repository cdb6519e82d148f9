// K8 / K9: projection + SH backward with camera-pose gradient.
//
// One thread per SCENE Gaussian; the thread loops over the views of its scene, sums that
// Gaussian's per-duplicate 2-D gradient records (contiguous slots written by blend-backward),
// chains them through conic -> cov2D -> (Sigma, J, W) -> (scale, quaternion, mean, pose) and through
// the SH colour (coefficients and view direction), and accumulates over views in registers /
// shared memory, so dL/d{means, scales, rotations, opacities, shs} are written exactly once with
// plain coalesced stores (no atomics).  The 15-float pose contribution (dA 9, dtau 3, dcampos 3)
// is reduced warp-shuffle -> block -> one partial per (view, block); K9 sums the partials in a
// fixed order and folds dcampos into dL/dviewmatrix.
//
// Replaces preprocess-backward of diff_gauss_pose incl. the viewmatrix gradient its `pose` branch
// adds (call site /root/reference/src/model/decoder/cuda_splatting.py:128-138).
#include "spf_adapter_math.cuh"
#include "spf_device.cuh"
#include "spf_kernels.h"
#include "spf_math.h"

namespace spf {

// MULTI = several views per scene: SH gradients accumulate over views in a second shared-memory buffer.
// !MULTI (training: one target view per scene): the SH gradient overwrites the staged SH input IN PLACE, so a
// block needs 38.4 KB instead of 76.8 KB of shared memory (5 resident blocks per SM instead of 2).
//
// Staging: the block's 128 SH rows are one contiguous 38.4 KB span in HBM (both [K,3] and [3,K] layouts), so
// when it is 16-B aligned one thread moves it with a single 1-D TMA bulk copy (cp.async.bulk + mbarrier) and
// the whole block meanwhile sums the duplicate records and runs the geometry backward; the SH gradient goes
// back the same way (bulk store).  Odd row length (75 floats at degree 4) makes the per-thread rows
// bank-conflict free without padding.  Unaligned / even-row cases use cooperative 128-bit copies.
template <bool MULTI>
__global__ void __launch_bounds__(PROJ_THREADS, MULTI ? 2 : 5)
project_backward_kernel(Dims d, SpfRasterIn in, SpfRasterState st, SpfRasterGradIn gin) {
  pdl_enter();
  extern __shared__ __align__(128) float smem[];
  __shared__ ViewConsts vc;
  __shared__ float pose_warp[PROJ_THREADS / 32][15];
  __shared__ __align__(8) uint64_t bar;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int scene = blockIdx.y;
  const int g0 = blockIdx.x * PROJ_THREADS;
  const int g = g0 + tid;
  const int nvalid = min(PROJ_THREADS, d.P - g0);
  const int row = 3 * in.sh_coeffs;
  const bool use_sh = in.shs != nullptr;
  const bool ck = (d.flags & SPF_FLAG_SH_LAYOUT_CK) != 0;
  const bool cov_grad = !(d.flags & SPF_FLAG_NO_COV_GRAD);
  const bool sh_grad = !(d.flags & SPF_FLAG_NO_SH_GRAD);
  const int sk = ck ? 1 : 3, sc = ck ? in.sh_coeffs : 1;

  // block-uniform: can the SH rows move by TMA?
  const size_t row_off = ((size_t)scene * d.P + g0) * row;
  const uint32_t bytes = (uint32_t)nvalid * row * 4u;
  const bool tma = use_sh && (row & 1) && ((bytes & 15u) == 0) &&
                   (((reinterpret_cast<uintptr_t>(in.shs) + row_off * 4) & 15) == 0) &&
                   (gin.dL_dshs == nullptr || ((reinterpret_cast<uintptr_t>(gin.dL_dshs) + row_off * 4) & 15) == 0);
  const int stride = ((row & 1) || tma) ? row : row + 1;
  float* sh_s = smem;                                            // input SH  [128][stride]
  float* dsh_s = MULTI ? smem + PROJ_THREADS * stride : smem;    // SH grads  [128][stride] (in place if !MULTI)

  if (use_sh) {
    if (tma) {
      if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, bytes);
        tma_load_1d(sh_s, in.shs + row_off, bytes, &bar);
      }
    } else {
      block_copy_g2s(sh_s, in.shs + row_off, nvalid * row, row, stride, tid, PROJ_THREADS);
    }
    if (MULTI)
      for (int i = tid; i < PROJ_THREADS * stride; i += PROJ_THREADS) dsh_s[i] = 0.0f;
  }

  const size_t sg = (size_t)scene * d.P + g;
  float m_in[3] = {0, 0, 0}, s_in[3] = {0, 0, 0}, q[4] = {1, 0, 0, 0};
  if (g < d.P) {
    for (int i = 0; i < 3; ++i) { m_in[i] = __ldg(in.means3D + sg * 3 + i); s_in[i] = __ldg(in.scales + sg * 3 + i); }
    const float4 qq = __ldg(reinterpret_cast<const float4*>(in.rotations) + sg);
    if (d.flags & SPF_FLAG_QUAT_XYZW) { q[0] = qq.w; q[1] = qq.x; q[2] = qq.y; q[3] = qq.z; }
    else { q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w; }
  }
  float dm[3] = {0, 0, 0}, ds[3] = {0, 0, 0}, dq[4] = {0, 0, 0, 0}, dop = 0.0f, dcol[3] = {0, 0, 0};

  for (int vi = 0; vi < d.v; ++vi) {
    const int view = scene * d.v + vi;
    // issue this view's per-Gaussian loads before the barrier so their latency overlaps the camera setup
    const size_t vg = (size_t)view * d.P + g;
    int tiles = 0, off = 0;
    float3 rgbv = make_float3(0.f, 0.f, 0.f);
    if (g < d.P) {
      tiles = st.tiles_touched[vg];
      off = st.dup_offset[vg];
      if (use_sh) rgbv = make_float3(st.rgb[vg * 3], st.rgb[vg * 3 + 1], st.rgb[vg * 3 + 2]);
    }
    __syncthreads();   // previous view's vc / pose_warp fully consumed; mbarrier init / plain SH staging visible
    if (tid == 0) {
      float V[16], Pm[16], bg[3];
      for (int i = 0; i < 16; ++i) { V[i] = in.viewmatrix[view * 16 + i]; Pm[i] = in.projmatrix[view * 16 + i]; }
      for (int i = 0; i < 3; ++i) bg[i] = in.bg[view * 3 + i];
      make_view_consts(vc, V, Pm, in.tanfov[view * 2], in.tanfov[view * 2 + 1], bg, d.mod, d.W, d.H);
    }
    // sum this Gaussian's duplicate records (contiguous slots) while thread 0 builds the view constants
    float a[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) a[k] = 0.0f;
    if (tiles > 0) {
      // a replayed CUDA graph whose duplicate count outgrew the frozen capacity: never read past the buffer (the
      // results of such a replay are invalid anyway and flagged through control[1])
      const int ndup = (int)max((int64_t)0, min((int64_t)tiles, d.cap - (int64_t)off));
      const float4* rec = reinterpret_cast<const float4*>(gin.dup_grad) + 3 * (size_t)off;
      for (int j = 0; j < ndup; ++j) {
        const float4 r0 = rec[3 * j], r1 = rec[3 * j + 1];
        const float2 r2 = *reinterpret_cast<const float2*>(rec + 3 * j + 2);
        a[0] += r0.x; a[1] += r0.y; a[2] += r0.z; a[3] += r0.w;
        a[4] += r1.x; a[5] += r1.y; a[6] += r1.z; a[7] += r1.w;
        a[8] += r2.x; a[9] += r2.y;
      }
    }
    __syncthreads();
    const float ps = in.pre_scale ? __ldg(in.pre_scale + view) : 1.0f;

    Grad3D o;
#pragma unroll
    for (int i = 0; i < 3; ++i) { o.dm[i] = 0.f; o.ds[i] = 0.f; o.dtau[i] = 0.f; o.dcam[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 4; ++i) o.dq[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) o.dA[i] = 0.f;

    if (tiles > 0) {
      Grad2D g2;
      g2.dpx = a[0]; g2.dpy = a[1]; g2.dconx = a[2]; g2.dcony = a[3]; g2.dconz = a[4];
      g2.dopacity = a[5]; g2.drgb[0] = a[6]; g2.drgb[1] = a[7]; g2.drgb[2] = a[8]; g2.ddepth = a[9];
      dop += g2.dopacity;
      if (gin.dL_dmeans2D) {
        float* o2 = gin.dL_dmeans2D + vg * 3;
        o2[0] = g2.dpx * 0.5f * vc.Wf; o2[1] = g2.dpy * 0.5f * vc.Hf; o2[2] = 0.0f;
      }
      const float m[3] = {m_in[0] * ps, m_in[1] * ps, m_in[2] * ps};
      const float s[3] = {s_in[0] * ps, s_in[1] * ps, s_in[2] * ps};
      project_backward_geom(vc, m, s, q, g2, cov_grad, o);
      if (use_sh) {
        const float dx = m[0] - vc.campos[0], dy = m[1] - vc.campos[1], dz = m[2] - vc.campos[2];
        const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        // clamp mask saved by the forward in the sign bit of rgb (-0.0f <=> SH colour was negative before the clamp)
        const float gm[3] = {(__float_as_uint(rgbv.x) >> 31) ? 0.0f : g2.drgb[0],
                             (__float_as_uint(rgbv.y) >> 31) ? 0.0f : g2.drgb[1],
                             (__float_as_uint(rgbv.z) >> 31) ? 0.0f : g2.drgb[2]};
        if (tma) mbar_wait(&bar, 0);
        float gdir[3];
        sh_backward_fused<MULTI>(d.deg, dx * inv, dy * inv, dz * inv, sh_s + tid * stride, dsh_s + tid * stride, sk, sc,
                                 gm, gdir[0], gdir[1], gdir[2]);
        if (sh_grad) view_dir_backward(vc, m, gdir, o);
      } else {
        dcol[0] += g2.drgb[0]; dcol[1] += g2.drgb[1]; dcol[2] += g2.drgb[2];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) { dm[i] += o.dm[i] * ps; ds[i] += o.ds[i] * ps; }
#pragma unroll
      for (int i = 0; i < 4; ++i) dq[i] += o.dq[i];
    } else {
      if (g < d.P && gin.dL_dmeans2D) {
        float* o2 = gin.dL_dmeans2D + vg * 3;
        o2[0] = 0.f; o2[1] = 0.f; o2[2] = 0.f;
      }
      if (use_sh && !MULTI) {
        // in-place mode: a Gaussian that contributes nothing must still hand back a zero SH gradient
        if (tma) mbar_wait(&bar, 0);
        float* mydsh = dsh_s + tid * stride;
        for (int i = 0; i < row; ++i) mydsh[i] = 0.0f;
      }
    }

    // pose contribution: 15 floats, warp shuffle -> block -> partial[view][block]
    float pv[15];
#pragma unroll
    for (int i = 0; i < 9; ++i) pv[i] = o.dA[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) { pv[9 + i] = o.dtau[i]; pv[12 + i] = o.dcam[i]; }
#pragma unroll
    for (int i = 0; i < 15; ++i) pv[i] = warp_sum(pv[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 15; ++i) pose_warp[wid][i] = pv[i];
    }
    __syncthreads();
    if (tid < 15) {
      float sum = 0.0f;
      for (int w = 0; w < PROJ_THREADS / 32; ++w) sum += pose_warp[w][tid];
      gin.pose_partial[((size_t)view * d.NB + blockIdx.x) * 16 + tid] = sum;
    }
  }

  if (g < d.P) {
    for (int i = 0; i < 3; ++i) { gin.dL_dmeans3D[sg * 3 + i] = dm[i]; gin.dL_dscales[sg * 3 + i] = ds[i]; }
    float4 qo;
    if (d.flags & SPF_FLAG_QUAT_XYZW) qo = make_float4(dq[1], dq[2], dq[3], dq[0]);
    else qo = make_float4(dq[0], dq[1], dq[2], dq[3]);
    reinterpret_cast<float4*>(gin.dL_drotations)[sg] = qo;
    gin.dL_dopacities[sg] = dop;
    if (!use_sh && gin.dL_dcolors)
      for (int i = 0; i < 3; ++i) gin.dL_dcolors[sg * 3 + i] = dcol[i];
  }
  // the bulk load must have landed before this CTA may exit or reuse the buffer, also when no thread consumed it
  // (every Gaussian of the block culled in every view)
  if (use_sh && tma && tid == 0) mbar_wait(&bar, 0);
  if (use_sh && gin.dL_dshs) {
    float* dst = gin.dL_dshs + row_off;
    if (tma) {
      fence_proxy_async();     // generic-proxy writes of dsh_s -> visible to the bulk-copy (async) proxy
      __syncthreads();
      if (tid == 0) {
        tma_store_1d(dst, dsh_s, bytes);
        tma_store_commit();
        tma_store_wait_all();  // shared memory must stay valid until the copy engine has read it
      }
    } else {
      __syncthreads();
      block_copy_s2g(dst, dsh_s, nvalid * row, row, stride, tid, PROJ_THREADS);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Raw-head variant (SpfRasterIn.raw_head): the block's 128 head-output rows [density logit?, 3 scale logits, 4 quaternion
// components, 3 x K SH coefficients] arrive by one bulk copy; scales / rotations / opacity are re-derived from them in
// registers (spf_adapter_math.cuh), the SH part is masked in place, and the gradient of the WHOLE row -- SH gradient
// times the mask, scale gradient through softplus / clamp, quaternion gradient through the normalisation, opacity
// gradient through the mapping and the sigmoid -- is assembled in shared memory (in place when a scene has one view)
// and leaves by one bulk store into dL_draw_head.  dL/d{scales, rotations, shs} never exist in HBM.
template <bool MULTI>
__global__ void __launch_bounds__(PROJ_THREADS, MULTI ? 2 : 4)
project_backward_raw_kernel(Dims d, SpfRasterIn in, SpfRasterState st, SpfRasterGradIn gin) {
  pdl_enter();
  extern __shared__ __align__(128) float smem[];
  __shared__ ViewConsts vc;
  __shared__ float pose_warp[PROJ_THREADS / 32][15];
  __shared__ __align__(8) uint64_t bar;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int scene = blockIdx.y;
  const int g0 = blockIdx.x * PROJ_THREADS;
  const int g = g0 + tid;
  const int nvalid = min(PROJ_THREADS, d.P - g0);
  const int R = in.raw_stride, dens = in.raw_has_density ? 1 : 0, K = in.sh_coeffs;
  const bool cov_grad = !(d.flags & SPF_FLAG_NO_COV_GRAD);
  const bool sh_grad = !(d.flags & SPF_FLAG_NO_SH_GRAD);
  const size_t row_off = ((size_t)scene * d.P + g0) * R;
  const uint32_t bytes = (uint32_t)nvalid * R * 4u;
  float* raw_s = smem;                                           // raw rows [128][R]
  float* out_s = MULTI ? smem + PROJ_THREADS * R : smem;         // gradient rows [128][R] (in place if !MULTI)
  float* myraw = raw_s + tid * R;
  float* myout = out_s + tid * R;

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
    mbar_expect_tx(&bar, bytes);
    tma_load_1d(raw_s, in.raw_head + row_off, bytes, &bar);
  }
  if (MULTI)
    for (int i = tid; i < PROJ_THREADS * R; i += PROJ_THREADS) out_s[i] = 0.0f;

  const size_t sg = (size_t)scene * d.P + g;
  float m_in[3] = {0, 0, 0};
  if (g < d.P)
    for (int i = 0; i < 3; ++i) m_in[i] = __ldg(in.means3D + sg * 3 + i);
  float xs[3] = {0, 0, 0}, qraw[4] = {1, 0, 0, 0}, logit = 0.0f;     // this Gaussian's raw parameters
  float s_in[3] = {0, 0, 0}, q[4] = {1, 0, 0, 0};
  bool have_row = false;
  float dm[3] = {0, 0, 0}, ds[3] = {0, 0, 0}, dq[4] = {0, 0, 0, 0}, dop = 0.0f;

  for (int vi = 0; vi < d.v; ++vi) {
    const int view = scene * d.v + vi;
    const size_t vg = (size_t)view * d.P + g;
    int tiles = 0, off = 0;
    float3 rgbv = make_float3(0.f, 0.f, 0.f);
    if (g < d.P) {
      tiles = st.tiles_touched[vg];
      off = st.dup_offset[vg];
      rgbv = make_float3(st.rgb[vg * 3], st.rgb[vg * 3 + 1], st.rgb[vg * 3 + 2]);
    }
    __syncthreads();   // previous view's vc / pose_warp fully consumed; mbarrier init visible
    if (tid == 0) {
      float V[16], Pm[16], bg[3];
      for (int i = 0; i < 16; ++i) { V[i] = in.viewmatrix[view * 16 + i]; Pm[i] = in.projmatrix[view * 16 + i]; }
      for (int i = 0; i < 3; ++i) bg[i] = in.bg[view * 3 + i];
      make_view_consts(vc, V, Pm, in.tanfov[view * 2], in.tanfov[view * 2 + 1], bg, d.mod, d.W, d.H);
    }
    // sum this Gaussian's duplicate records (contiguous slots) while thread 0 builds the view constants
    float a[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) a[k] = 0.0f;
    if (tiles > 0) {
      const int ndup = (int)max((int64_t)0, min((int64_t)tiles, d.cap - (int64_t)off));
      const float4* rec = reinterpret_cast<const float4*>(gin.dup_grad) + 3 * (size_t)off;
      for (int j = 0; j < ndup; ++j) {
        const float4 r0 = rec[3 * j], r1 = rec[3 * j + 1];
        const float2 r2 = *reinterpret_cast<const float2*>(rec + 3 * j + 2);
        a[0] += r0.x; a[1] += r0.y; a[2] += r0.z; a[3] += r0.w;
        a[4] += r1.x; a[5] += r1.y; a[6] += r1.z; a[7] += r1.w;
        a[8] += r2.x; a[9] += r2.y;
      }
    }
    __syncthreads();
    if (!have_row) {
      // the raw rows have landed: take this Gaussian's parameters into registers (the SH coefficients stay unmasked in
      // shared memory; sh_backward_fused<., MASKED> applies the mask as it reads them)
      mbar_wait(&bar, 0);
      if (g < d.P) {
        if (dens) logit = myraw[0];
#pragma unroll
        for (int c = 0; c < 3; ++c) { xs[c] = myraw[dens + c]; s_in[c] = head_scale(xs[c]); }
#pragma unroll
        for (int c = 0; c < 4; ++c) qraw[c] = myraw[dens + 3 + c];
        head_quat(qraw, in.raw_eps, q);
      }
      have_row = true;
    }
    const float ps = in.pre_scale ? __ldg(in.pre_scale + view) : 1.0f;

    Grad3D o;
#pragma unroll
    for (int i = 0; i < 3; ++i) { o.dm[i] = 0.f; o.ds[i] = 0.f; o.dtau[i] = 0.f; o.dcam[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 4; ++i) o.dq[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) o.dA[i] = 0.f;

    if (tiles > 0) {
      Grad2D g2;
      g2.dpx = a[0]; g2.dpy = a[1]; g2.dconx = a[2]; g2.dcony = a[3]; g2.dconz = a[4];
      g2.dopacity = a[5]; g2.drgb[0] = a[6]; g2.drgb[1] = a[7]; g2.drgb[2] = a[8]; g2.ddepth = a[9];
      dop += g2.dopacity;
      if (gin.dL_dmeans2D) {
        float* o2 = gin.dL_dmeans2D + vg * 3;
        o2[0] = g2.dpx * 0.5f * vc.Wf; o2[1] = g2.dpy * 0.5f * vc.Hf; o2[2] = 0.0f;
      }
      const float m[3] = {m_in[0] * ps, m_in[1] * ps, m_in[2] * ps};
      const float s[3] = {s_in[0] * ps, s_in[1] * ps, s_in[2] * ps};
      project_backward_geom(vc, m, s, q, g2, cov_grad, o);
      {
        const float dx = m[0] - vc.campos[0], dy = m[1] - vc.campos[1], dz = m[2] - vc.campos[2];
        const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        const float gm[3] = {(__float_as_uint(rgbv.x) >> 31) ? 0.0f : g2.drgb[0],
                             (__float_as_uint(rgbv.y) >> 31) ? 0.0f : g2.drgb[1],
                             (__float_as_uint(rgbv.z) >> 31) ? 0.0f : g2.drgb[2]};
        float gdir[3];
        sh_backward_fused<MULTI, true>(d.deg, dx * inv, dy * inv, dz * inv, myraw + dens + 7, myout + dens + 7, 1, K, gm,
                                       gdir[0], gdir[1], gdir[2]);
        if (sh_grad) view_dir_backward(vc, m, gdir, o);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) { dm[i] += o.dm[i] * ps; ds[i] += o.ds[i] * ps; }
#pragma unroll
      for (int i = 0; i < 4; ++i) dq[i] += o.dq[i];
    } else {
      if (g < d.P && gin.dL_dmeans2D) {
        float* o2 = gin.dL_dmeans2D + vg * 3;
        o2[0] = 0.f; o2[1] = 0.f; o2[2] = 0.f;
      }
      if (!MULTI && g < d.P) {
        // in-place mode: a Gaussian that contributes nothing must still hand back a zero SH gradient
        for (int i = 0; i < 3 * K; ++i) myout[dens + 7 + i] = 0.0f;
      }
    }

    // pose contribution: 15 floats, warp shuffle -> block -> partial[view][block]
    float pv[15];
#pragma unroll
    for (int i = 0; i < 9; ++i) pv[i] = o.dA[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) { pv[9 + i] = o.dtau[i]; pv[12 + i] = o.dcam[i]; }
#pragma unroll
    for (int i = 0; i < 15; ++i) pv[i] = warp_sum(pv[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 15; ++i) pose_warp[wid][i] = pv[i];
    }
    __syncthreads();
    if (tid < 15) {
      float sum = 0.0f;
      for (int w = 0; w < PROJ_THREADS / 32; ++w) sum += pose_warp[w][tid];
      gin.pose_partial[((size_t)view * d.NB + blockIdx.x) * 16 + tid] = sum;
    }
  }

  if (g < d.P) {
    for (int i = 0; i < 3; ++i) gin.dL_dmeans3D[sg * 3 + i] = dm[i];
    // the rest of the row: SH gradient x mask (already applied when a scene has one view), scale logits, raw quaternion,
    // density logit
    if (MULTI) sh_mask_rows(myout + dens + 7, K);
#pragma unroll
    for (int c = 0; c < 3; ++c) myout[dens + c] = ds[c] * head_scale_grad(xs[c]);
    float o4[4];
    head_quat_grad(qraw, dq, in.raw_eps, o4);
#pragma unroll
    for (int c = 0; c < 4; ++c) myout[dens + 3 + c] = o4[c];
    if (dens) myout[0] = (dop != 0.0f) ? dop * head_opacity_grad(logit, in.opacity_exponent) : 0.0f;
    else gin.dL_dopacities[sg] = dop;
  }
  fence_proxy_async();     // generic-proxy writes of out_s -> visible to the bulk-copy (async) proxy
  __syncthreads();
  if (tid == 0) {
    tma_store_1d(gin.dL_draw_head + row_off, out_s, bytes);
    tma_store_commit();
    tma_store_wait_all();  // shared memory must stay valid until the copy engine has read it
  }
}

// K9: dL/dviewmatrix[view] = fixed-order sum of block partials, campos gradient folded in.  One CTA per view:
// 15 components x 16 interleaved slices on 240 threads (independent, pipelined loads), then a 16-way fixed-order sum.
__global__ void __launch_bounds__(256)
pose_reduce_kernel(Dims d, const float* __restrict__ viewmatrix, const float* __restrict__ partial,
                   float* __restrict__ dV) {
  pdl_enter();
  __shared__ float red[16][15];
  const int view = blockIdx.x, tid = threadIdx.x;
  const int c = tid % 15, part = tid / 15;
  if (tid < 240) {
    float a0 = 0.0f, a1 = 0.0f;
    const float* p = partial + (size_t)view * d.NB * 16 + c;
    int b = part;
    for (; b + 16 < d.NB; b += 32) { a0 += __ldg(p + (size_t)b * 16); a1 += __ldg(p + (size_t)(b + 16) * 16); }
    if (b < d.NB) a0 += __ldg(p + (size_t)b * 16);
    red[part][c] = a0 + a1;
  }
  __syncthreads();
  if (tid == 0) {
    float tot[15];
    for (int i = 0; i < 15; ++i) {
      float t = 0.0f;
      for (int q = 0; q < 16; ++q) t += red[q][i];
      tot[i] = t;
    }
    float dA[9], dtau[3], dcam[3], V[16];
    for (int i = 0; i < 9; ++i) dA[i] = tot[i];
    for (int i = 0; i < 3; ++i) { dtau[i] = tot[9 + i]; dcam[i] = tot[12 + i]; }
    for (int i = 0; i < 16; ++i) V[i] = viewmatrix[view * 16 + i];
    fold_campos_grad(V, dcam, dA, dtau);
    float* o = dV + view * 16;
    for (int i = 0; i < 16; ++i) o[i] = 0.0f;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) o[4 * i + j] = dA[3 * i + j];
    for (int j = 0; j < 3; ++j) o[12 + j] = dtau[j];
  }
}

template <bool MULTI>
static cudaError_t launch_pb(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st, const SpfRasterGradIn& gin,
                             cudaStream_t s) {
  const int row = 3 * in.sh_coeffs;
  const int stride = (row & 1) ? row : row + 1;
  const size_t smem = in.shs ? (size_t)(MULTI ? 2 : 1) * PROJ_THREADS * stride * sizeof(float) : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(project_backward_kernel<MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid(d.NB, d.S);
  pdl_launch(project_backward_kernel<MULTI>, grid, PROJ_THREADS, smem, s)(d, in, st, gin);
  return cudaGetLastError();
}

template <bool MULTI>
static cudaError_t launch_pb_raw(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st, const SpfRasterGradIn& gin,
                                 cudaStream_t s) {
  const size_t smem = (size_t)(MULTI ? 2 : 1) * PROJ_THREADS * in.raw_stride * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(project_backward_raw_kernel<MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid(d.NB, d.S);
  pdl_launch(project_backward_raw_kernel<MULTI>, grid, PROJ_THREADS, smem, s)(d, in, st, gin);
  return cudaGetLastError();
}

cudaError_t launch_project_backward(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st,
                                    const SpfRasterGradIn& gin, cudaStream_t s) {
  if (in.raw_head) return d.v > 1 ? launch_pb_raw<true>(d, in, st, gin, s) : launch_pb_raw<false>(d, in, st, gin, s);
  return d.v > 1 ? launch_pb<true>(d, in, st, gin, s) : launch_pb<false>(d, in, st, gin, s);
}

cudaError_t launch_pose_reduce(const Dims& d, const SpfRasterIn& in, const SpfRasterGradIn& gin, cudaStream_t s) {
  pdl_launch(pose_reduce_kernel, d.B, 256, 0, s)(d, in.viewmatrix, gin.pose_partial, gin.dL_dviewmatrix);
  return cudaGetLastError();
}

}  // namespace spf
