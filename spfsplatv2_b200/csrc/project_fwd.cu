// K1: per-Gaussian 3-D -> 2-D EWA projection, 2-D covariance / conic / radius / tile rectangle and
// SH -> RGB, for all B views of a batched call.  One thread per (view, Gaussian); a block's 128 SH
// records (128 x 300 B at degree 4) are staged through shared memory with 128-bit coalesced loads
// and read back conflict-free (odd stride).
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: radii, rectangles and depth keys must be
// bit-identical to oracle/raster_oracle.py (see spf_math.h).
//
// Replaces the preprocess stage of diff_gauss_pose (call site
// /root/reference/src/model/decoder/cuda_splatting.py:128-138; algorithm SURVEY.md App. B 1-10).
#include <cstdlib>

#include "spf_adapter_math.cuh"
#include "spf_device.cuh"
#include "spf_kernels.h"
#include "spf_math.h"

namespace spf {

__global__ void __launch_bounds__(PROJ_THREADS)
project_forward_kernel(Dims d, SpfRasterIn in, SpfRasterState st, int* __restrict__ tile_count,
                       int* __restrict__ block_sum) {
  pdl_enter();
  extern __shared__ __align__(128) float sh_s[];
  __shared__ ViewConsts vc;
  __shared__ __align__(8) uint64_t bar;
  __shared__ int warp_tot[PROJ_THREADS / 32];

  const int tid = threadIdx.x;
  const int view = blockIdx.y;
  const int scene = view / d.v;
  const int g0 = blockIdx.x * PROJ_THREADS;
  const int g = g0 + tid;
  const int nvalid = min(PROJ_THREADS, d.P - g0);
  const float ps = in.pre_scale ? __ldg(in.pre_scale + view) : 1.0f;

  // SH staging: the block's rows are one contiguous span; when it is 16-B aligned and the row length is odd
  // (bank-conflict free without padding) ONE thread moves it with a 1-D TMA bulk copy and the projection math
  // below overlaps the transfer; otherwise cooperative 128-bit loads.
  const int row = 3 * in.sh_coeffs;
  const size_t row_off = ((size_t)scene * d.P + g0) * row;
  const uint32_t bytes = (uint32_t)nvalid * row * 4u;
  const bool tma = in.shs && (row & 1) && ((bytes & 15u) == 0) &&
                   (((reinterpret_cast<uintptr_t>(in.shs) + row_off * 4) & 15) == 0);
  const int stride = ((row & 1) || tma) ? row : row + 1;
  if (in.shs) {
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
  }
  // per-Gaussian loads are issued before the barrier so their latency overlaps thread 0's camera setup
  const size_t sg = (size_t)scene * d.P + min(g, d.P - 1);
  const size_t vg = (size_t)view * d.P + g;
  float m[3], s[3], q[4];
  for (int i = 0; i < 3; ++i) {
    m[i] = __ldg(in.means3D + sg * 3 + i) * ps;
    s[i] = __ldg(in.scales + sg * 3 + i) * ps;
  }
  {
    const float4 qq = __ldg(reinterpret_cast<const float4*>(in.rotations) + sg);
    if (d.flags & SPF_FLAG_QUAT_XYZW) { q[0] = qq.w; q[1] = qq.x; q[2] = qq.y; q[3] = qq.z; }
    else { q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w; }
  }
  const float opac = __ldg(in.opacities + sg);
  if (tid == 0) {
    float V[16], Pm[16], bg[3];
    for (int i = 0; i < 16; ++i) { V[i] = in.viewmatrix[view * 16 + i]; Pm[i] = in.projmatrix[view * 16 + i]; }
    for (int i = 0; i < 3; ++i) bg[i] = in.bg[view * 3 + i];
    make_view_consts(vc, V, Pm, in.tanfov[view * 2], in.tanfov[view * 2 + 1], bg, d.mod, d.W, d.H);
  }
  __syncthreads();

  int tiles = 0, cx0 = 0, cy0 = 0, cw = 1;
  if (g < d.P) {
    Projected o;
    const bool vis = project_forward(vc, m, s, q, o);
    tiles = o.tiles;

    float rgb[3];
    if (in.shs) {
      const float dx = m[0] - vc.campos[0], dy = m[1] - vc.campos[1], dz = m[2] - vc.campos[2];
      const float inv = 1.0f / sqrtf((dx * dx + dy * dy) + dz * dz);
      if (tma) mbar_wait(&bar, 0);
      const bool ck = (d.flags & SPF_FLAG_SH_LAYOUT_CK) != 0;
      float pre[3];
      sh_eval_fused(d.deg, dx * inv, dy * inv, dz * inv, sh_s + tid * stride, ck ? 1 : 3, ck ? in.sh_coeffs : 1, pre);
      // clamp at 0; a channel that was negative is stored as -0.0f: it blends like 0, and its sign bit is the
      // clamp mask projection-backward needs (so it never re-evaluates the SH sum)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float p5 = pre[c] + 0.5f;
        rgb[c] = (p5 < 0.0f) ? -0.0f : p5;
      }
    } else {
      for (int c = 0; c < 3; ++c) rgb[c] = __ldg(in.colors_precomp + sg * 3 + c);
    }

    reinterpret_cast<float2*>(st.xy)[vg] = make_float2(o.px, o.py);
    st.depth[vg] = o.depth;
    reinterpret_cast<float4*>(st.conic_opacity)[vg] =
        make_float4(o.conx, o.cony, o.conz, opac);
    st.rgb[vg * 3 + 0] = rgb[0]; st.rgb[vg * 3 + 1] = rgb[1]; st.rgb[vg * 3 + 2] = rgb[2];
    st.radii[vg] = o.radius;
    st.tiles_touched[vg] = o.tiles;
    if (vis) { cx0 = o.rx0; cy0 = o.ry0; cw = o.rx1 - o.rx0; }
  }
  // per-tile duplicate counts: warp-aggregated atomics over the cells of every lane's tile rectangle
  {
    int* tc = tile_count + (size_t)view * d.T;
    int maxc = tiles;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
    int x = 0, y = 0;
    for (int k = 0; k < maxc; ++k) {
      const bool has = k < tiles;
      warp_aggregated_add(has, tc, (cy0 + y) * d.gx + cx0 + x, tid & 31);
      if (++x == cw) { x = 0; ++y; }
    }
  }
  // block total of tiles_touched (for the duplicate-slot prefix sums)
  const int wsum = warp_sum_i(tiles);
  if ((tid & 31) == 0) warp_tot[tid >> 5] = wsum;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < PROJ_THREADS / 32; ++w) t += warp_tot[w];
    block_sum[(size_t)view * d.NB + blockIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------------------------------
// Streaming variant (the common case: SH colours, odd row length, 16-byte aligned tensors, P % 4 == 0):
// a PERSISTENT kernel, two CTAs per SM, each looping over (view, 128-Gaussian chunk) items with a two-stage
// shared-memory ring.  ALL inputs of an item (SH rows 38.4 KB, means, scales, rotations, opacities) are moved by 1-D
// TMA bulk copies onto one mbarrier, and the copies of item k+1 are issued before item k is computed: the memory
// system always has a full chunk per CTA in flight, while the one-shot kernel above only loads during the first
// part of every block's life (ncu: 50 % of HBM peak, long-scoreboard + barrier stalls).  Same arithmetic, same
// op order (same functions, same -fmad=false translation unit) => same bits.
// Per-view constants of item `view` (thread-serial: 32 cached loads and a handful of flops).
__device__ __forceinline__ void load_view_consts(ViewConsts& vc, const Dims& d, const SpfRasterIn& in, int view) {
  float V[16], Pm[16], bg[3];
  for (int i = 0; i < 16; ++i) { V[i] = in.viewmatrix[view * 16 + i]; Pm[i] = in.projmatrix[view * 16 + i]; }
  for (int i = 0; i < 3; ++i) bg[i] = in.bg[view * 3 + i];
  make_view_consts(vc, V, Pm, in.tanfov[view * 2], in.tanfov[view * 2 + 1], bg, d.mod, d.W, d.H);
}

constexpr int VC_MAX = 32;

struct __align__(128) PFStage {
  float sh[PROJ_THREADS * 75];
  float means[PROJ_THREADS * 3];
  float scales[PROJ_THREADS * 3];
  float4 rot[PROJ_THREADS];
  float opac[PROJ_THREADS];
};

template <int NSTAGE>
struct PFSmemT {
  PFStage stage[NSTAGE];
  ViewConsts vc[VC_MAX];
  uint64_t full[2];
  int warp_tot[2][PROJ_THREADS / 32];
};

// NSTAGE = 2: two CTAs per SM, item k+1 in flight while item k is computed.  NSTAGE = 1: FOUR CTAs per SM, every CTA
// loads, then computes -- the other three CTAs of the SM cover its load (twice the warps to hide the latency of the
// dependent sqrt / divide / shared-memory chains of the compute phase).
template <int NSTAGE>
__global__ void __launch_bounds__(2 * PROJ_THREADS, NSTAGE == 1 ? 4 : 2)
project_forward_stream_kernel(Dims d, SpfRasterIn in, SpfRasterState st, int* __restrict__ tile_count,
                              int* __restrict__ block_sum) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  PFSmemT<NSTAGE>& S = *reinterpret_cast<PFSmemT<NSTAGE>*>(pf_smem_raw);
  const int tid = threadIdx.x;
  const int row = 3 * in.sh_coeffs;
  const int total = d.B * d.NB;
  const bool ck = (d.flags & SPF_FLAG_SH_LAYOUT_CK) != 0;
  const int sk = ck ? 1 : 3, sc = ck ? in.sh_coeffs : 1;

  if (tid == 0) { mbar_init(&S.full[0], 1); mbar_init(&S.full[1], 1); mbar_fence_init(); }
  // Up to VC_MAX views: every view's constants are built once into a table.  More views (validation / video renders with
  // hundreds of views per scene): the table is a two-entry ring, the constants of item k+1 are built by the block's last
  // thread while item k is computed (published by item k's closing barrier).
  const bool per_item = d.B > VC_MAX;
  const int vc_thread = 2 * PROJ_THREADS - 1;
  if (!per_item) {
    if (tid < d.B) load_view_consts(S.vc[tid], d, in, tid);
  } else if (tid == vc_thread && (int)blockIdx.x < total) {
    load_view_consts(S.vc[0], d, in, (int)blockIdx.x / d.NB);
  }
  __syncthreads();

  auto issue = [&](int item, int sidx) {      // thread 0 only
    const int view = item / d.NB, chunk = item - view * d.NB;
    const int scene = view / d.v;
    const int g0 = chunk * PROJ_THREADS;
    const uint32_t nv = (uint32_t)min(PROJ_THREADS, d.P - g0);
    const size_t sg0 = (size_t)scene * d.P + g0;
    PFStage& T = S.stage[sidx];
    mbar_expect_tx(&S.full[sidx], nv * (uint32_t)(row * 4 + 12 + 12 + 16 + 4));
    tma_load_1d(T.sh, in.shs + sg0 * row, nv * (uint32_t)row * 4u, &S.full[sidx]);
    tma_load_1d(T.means, in.means3D + sg0 * 3, nv * 12u, &S.full[sidx]);
    tma_load_1d(T.scales, in.scales + sg0 * 3, nv * 12u, &S.full[sidx]);
    tma_load_1d(T.rot, in.rotations + sg0 * 4, nv * 16u, &S.full[sidx]);
    tma_load_1d(T.opac, in.opacities + sg0, nv * 4u, &S.full[sidx]);
  };

  int item = blockIdx.x;
  if (NSTAGE == 2 && tid == 0 && item < total) issue(item, 0);
  for (int k = 0; item < total; ++k, item += gridDim.x) {
    const int sidx = NSTAGE == 2 ? (k & 1) : 0;
    if (NSTAGE == 2) {
      if (tid == 0 && item + (int)gridDim.x < total) issue(item + gridDim.x, sidx ^ 1);
    } else if (tid == 0) {
      issue(item, 0);      // the previous item's closing barrier has released the stage
    }
    const int view = item / d.NB, chunk = item - view * d.NB;
    const int role = tid >> 7, gi = tid & (PROJ_THREADS - 1);   // warps 0-3: geometry, warps 4-7: SH colour
    const int g = chunk * PROJ_THREADS + gi;
    const float ps = in.pre_scale ? __ldg(in.pre_scale + view) : 1.0f;
    if (per_item && tid == vc_thread && item + (int)gridDim.x < total)
      load_view_consts(S.vc[(k + 1) & 1], d, in, (item + (int)gridDim.x) / d.NB);
    const ViewConsts& vc = S.vc[per_item ? (k & 1) : view];
    const PFStage& T = S.stage[sidx];
    mbar_wait(&S.full[sidx], (uint32_t)(NSTAGE == 2 ? ((k >> 1) & 1) : (k & 1)));

    int tiles = 0, cx0 = 0, cy0 = 0, cw = 1;
    const size_t vg = (size_t)view * d.P + g;
    if (role == 1) {
      if (g < d.P) {
        float m[3];
        for (int i = 0; i < 3; ++i) m[i] = T.means[gi * 3 + i] * ps;
        const float dx = m[0] - vc.campos[0], dy = m[1] - vc.campos[1], dz = m[2] - vc.campos[2];
        const float inv = 1.0f / sqrtf((dx * dx + dy * dy) + dz * dz);
        float pre[3];
        sh_eval_fused(d.deg, dx * inv, dy * inv, dz * inv, T.sh + gi * row, sk, sc, pre);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float p5 = pre[c] + 0.5f;
          st.rgb[vg * 3 + c] = (p5 < 0.0f) ? -0.0f : p5;     // sign bit = clamp mask for the backward
        }
      }
    } else if (g < d.P) {
      float m[3], s[3], q[4];
      for (int i = 0; i < 3; ++i) { m[i] = T.means[gi * 3 + i] * ps; s[i] = T.scales[gi * 3 + i] * ps; }
      {
        const float4 qq = T.rot[gi];
        if (d.flags & SPF_FLAG_QUAT_XYZW) { q[0] = qq.w; q[1] = qq.x; q[2] = qq.y; q[3] = qq.z; }
        else { q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w; }
      }
      Projected o;
      const bool vis = project_forward(vc, m, s, q, o);
      tiles = o.tiles;
      reinterpret_cast<float2*>(st.xy)[vg] = make_float2(o.px, o.py);
      st.depth[vg] = o.depth;
      reinterpret_cast<float4*>(st.conic_opacity)[vg] = make_float4(o.conx, o.cony, o.conz, T.opac[gi]);
      st.radii[vg] = o.radius;
      st.tiles_touched[vg] = o.tiles;
      if (vis) { cx0 = o.rx0; cy0 = o.ry0; cw = o.rx1 - o.rx0; }
    }
    if (role == 0) {
      int* tc = tile_count + (size_t)view * d.T;
      int maxc = tiles;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
      int x = 0, y = 0;
      for (int kk = 0; kk < maxc; ++kk) {
        const bool has = kk < tiles;
        warp_aggregated_add(has, tc, (cy0 + y) * d.gx + cx0 + x, tid & 31);
        if (++x == cw) { x = 0; ++y; }
      }
    }
    const int wsum = warp_sum_i(tiles);
    if (role == 0 && (tid & 31) == 0) S.warp_tot[k & 1][tid >> 5] = wsum;
    __syncthreads();     // stage sidx fully consumed (it is refilled two items later); warp totals visible
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < PROJ_THREADS / 32; ++w) t += S.warp_tot[k & 1][w];
      block_sum[(size_t)view * d.NB + chunk] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Raw-head variant of the streaming kernel (SpfRasterIn.raw_head): an item's 128 head-output rows
// [density logit?, 3 scale logits, 4 quaternion components, 3 x K SH coefficients] arrive by ONE bulk copy; the
// adapter's maps (softplus / clamp, quaternion normalisation, SH mask, density sigmoid + opacity mapping --
// spf_adapter_math.cuh, the same device functions the stand-alone adapter kernel uses) are applied in registers on the
// way into the projection, so scales / rotations / harmonics / opacities never exist in HBM.  Same role split, same
// ring, same arithmetic downstream as the streaming kernel above.
struct __align__(128) PFRawStage {
  float raw[PROJ_THREADS * 83];      // rows of R <= 83 floats, packed
  float means[PROJ_THREADS * 3];
  float opac[PROJ_THREADS];          // only when the rows carry no density logit
};

template <int NSTAGE>
struct PFRawSmemT {
  PFRawStage stage[NSTAGE];
  ViewConsts vc[VC_MAX];
  uint64_t full[2];
  int warp_tot[2][PROJ_THREADS / 32];
};

template <int NSTAGE>
__global__ void __launch_bounds__(2 * PROJ_THREADS, NSTAGE == 1 ? 4 : 2)
project_forward_raw_kernel(Dims d, SpfRasterIn in, SpfRasterState st, int* __restrict__ tile_count,
                           int* __restrict__ block_sum) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char pf_smem_raw[];
  PFRawSmemT<NSTAGE>& S = *reinterpret_cast<PFRawSmemT<NSTAGE>*>(pf_smem_raw);
  const int tid = threadIdx.x;
  const int R = in.raw_stride, dens = in.raw_has_density ? 1 : 0;
  const int total = d.B * d.NB;

  if (tid == 0) { mbar_init(&S.full[0], 1); mbar_init(&S.full[1], 1); mbar_fence_init(); }
  // Up to VC_MAX views: every view's constants are built once into a table.  More views (validation / video renders with
  // hundreds of views per scene): the table is a two-entry ring, the constants of item k+1 are built by the block's last
  // thread while item k is computed (published by item k's closing barrier).
  const bool per_item = d.B > VC_MAX;
  const int vc_thread = 2 * PROJ_THREADS - 1;
  if (!per_item) {
    if (tid < d.B) load_view_consts(S.vc[tid], d, in, tid);
  } else if (tid == vc_thread && (int)blockIdx.x < total) {
    load_view_consts(S.vc[0], d, in, (int)blockIdx.x / d.NB);
  }
  __syncthreads();

  auto issue = [&](int item, int sidx) {      // thread 0 only
    const int view = item / d.NB, chunk = item - view * d.NB;
    const int scene = view / d.v;
    const int g0 = chunk * PROJ_THREADS;
    const uint32_t nv = (uint32_t)min(PROJ_THREADS, d.P - g0);
    const size_t sg0 = (size_t)scene * d.P + g0;
    PFRawStage& T = S.stage[sidx];
    mbar_expect_tx(&S.full[sidx], nv * (uint32_t)(R * 4 + 12 + (dens ? 0 : 4)));
    tma_load_1d(T.raw, in.raw_head + sg0 * R, nv * (uint32_t)R * 4u, &S.full[sidx]);
    tma_load_1d(T.means, in.means3D + sg0 * 3, nv * 12u, &S.full[sidx]);
    if (!dens) tma_load_1d(T.opac, in.opacities + sg0, nv * 4u, &S.full[sidx]);
  };

  int item = blockIdx.x;
  if (NSTAGE == 2 && tid == 0 && item < total) issue(item, 0);
  for (int k = 0; item < total; ++k, item += gridDim.x) {
    const int sidx = NSTAGE == 2 ? (k & 1) : 0;
    if (NSTAGE == 2) {
      if (tid == 0 && item + (int)gridDim.x < total) issue(item + gridDim.x, sidx ^ 1);
    } else if (tid == 0) {
      issue(item, 0);      // the previous item's closing barrier has released the stage
    }
    const int view = item / d.NB, chunk = item - view * d.NB;
    const int role = tid >> 7, gi = tid & (PROJ_THREADS - 1);   // warps 0-3: geometry, warps 4-7: SH colour
    const int g = chunk * PROJ_THREADS + gi;
    const float ps = in.pre_scale ? __ldg(in.pre_scale + view) : 1.0f;
    if (per_item && tid == vc_thread && item + (int)gridDim.x < total)
      load_view_consts(S.vc[(k + 1) & 1], d, in, (item + (int)gridDim.x) / d.NB);
    const ViewConsts& vc = S.vc[per_item ? (k & 1) : view];
    const PFRawStage& T = S.stage[sidx];
    mbar_wait(&S.full[sidx], (uint32_t)(NSTAGE == 2 ? ((k >> 1) & 1) : (k & 1)));
    const float* rowp = T.raw + gi * R + dens;                  // [3 scale logits, 4 quaternion, 3 x K SH]

    int tiles = 0, cx0 = 0, cy0 = 0, cw = 1;
    const size_t vg = (size_t)view * d.P + g;
    if (role == 1) {
      if (g < d.P) {
        float m[3];
        for (int i = 0; i < 3; ++i) m[i] = T.means[gi * 3 + i] * ps;
        const float dx = m[0] - vc.campos[0], dy = m[1] - vc.campos[1], dz = m[2] - vc.campos[2];
        const float inv = 1.0f / sqrtf((dx * dx + dy * dy) + dz * dz);
        float pre[3];
        sh_eval_fused_t<true>(d.deg, dx * inv, dy * inv, dz * inv, rowp + 7, 1, in.sh_coeffs, pre);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float p5 = pre[c] + 0.5f;
          st.rgb[vg * 3 + c] = (p5 < 0.0f) ? -0.0f : p5;     // sign bit = clamp mask for the backward
        }
      }
    } else if (g < d.P) {
      float m[3], s[3], q[4];
      for (int i = 0; i < 3; ++i) { m[i] = T.means[gi * 3 + i] * ps; s[i] = head_scale(rowp[i]) * ps; }
      {
        const float qr[4] = {rowp[3], rowp[4], rowp[5], rowp[6]};
        head_quat(qr, in.raw_eps, q);
      }
      const float opac = dens ? head_opacity(T.raw[gi * R], in.opacity_exponent) : T.opac[gi];
      Projected o;
      const bool vis = project_forward(vc, m, s, q, o);
      tiles = o.tiles;
      reinterpret_cast<float2*>(st.xy)[vg] = make_float2(o.px, o.py);
      st.depth[vg] = o.depth;
      reinterpret_cast<float4*>(st.conic_opacity)[vg] = make_float4(o.conx, o.cony, o.conz, opac);
      st.radii[vg] = o.radius;
      st.tiles_touched[vg] = o.tiles;
      if (vis) { cx0 = o.rx0; cy0 = o.ry0; cw = o.rx1 - o.rx0; }
    }
    if (role == 0) {
      int* tc = tile_count + (size_t)view * d.T;
      int maxc = tiles;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
      int x = 0, y = 0;
      for (int kk = 0; kk < maxc; ++kk) {
        const bool has = kk < tiles;
        warp_aggregated_add(has, tc, (cy0 + y) * d.gx + cx0 + x, tid & 31);
        if (++x == cw) { x = 0; ++y; }
      }
    }
    const int wsum = warp_sum_i(tiles);
    if (role == 0 && (tid & 31) == 0) S.warp_tot[k & 1][tid >> 5] = wsum;
    __syncthreads();     // stage sidx fully consumed (it is refilled two items later); warp totals visible
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < PROJ_THREADS / 32; ++w) t += S.warp_tot[k & 1][w];
      block_sum[(size_t)view * d.NB + chunk] = t;
    }
  }
}

static bool stream_ok(const Dims& d, const SpfRasterIn& in) {
  if (!in.shs) return false;
  const int row = 3 * in.sh_coeffs;
  if (!(row & 1) || row > 75) return false;
  if (d.P % 4 != 0) return false;
  const uintptr_t a = reinterpret_cast<uintptr_t>(in.shs) | reinterpret_cast<uintptr_t>(in.means3D) |
                      reinterpret_cast<uintptr_t>(in.scales) | reinterpret_cast<uintptr_t>(in.rotations) |
                      reinterpret_cast<uintptr_t>(in.opacities);
  return (a & 15) == 0;
}

cudaError_t launch_project_forward(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st,
                                   const ControlLayout& cl, cudaStream_t s) {
  // SPF_PF_STAGES=2 selects the two-stage ring at two CTAs per SM (A/B switch; default: one stage, four CTAs per SM)
  static const bool two_stage = [] { const char* v = getenv("SPF_PF_STAGES"); return v && v[0] == '2'; }();
  if (in.raw_head) {      // (shape / alignment requirements were checked by the C entry point)
    if (two_stage) {
      cudaError_t e = cudaFuncSetAttribute(project_forward_raw_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(PFRawSmemT<2>));
      if (e != cudaSuccess) return e;
      pdl_launch(project_forward_raw_kernel<2>, min(d.B * d.NB, 2 * sm_count()), 2 * PROJ_THREADS, sizeof(PFRawSmemT<2>), s)(
          d, in, st, st.control + cl.tile_count, st.control + cl.block_sum);
      return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(project_forward_raw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(PFRawSmemT<1>));
    if (e != cudaSuccess) return e;
    pdl_launch(project_forward_raw_kernel<1>, min(d.B * d.NB, 4 * sm_count()), 2 * PROJ_THREADS, sizeof(PFRawSmemT<1>), s)(
        d, in, st, st.control + cl.tile_count, st.control + cl.block_sum);
    return cudaGetLastError();
  }
  const int row = 3 * in.sh_coeffs;
  const int stride = (row & 1) ? row : row + 1;
  const size_t smem = in.shs ? (size_t)PROJ_THREADS * stride * sizeof(float) : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(project_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (stream_ok(d, in) && !(d.flags & SPF_FLAG_NO_TMA)) {
    if (two_stage) {
      cudaError_t e = cudaFuncSetAttribute(project_forward_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(PFSmemT<2>));
      if (e != cudaSuccess) return e;
      pdl_launch(project_forward_stream_kernel<2>, min(d.B * d.NB, 2 * sm_count()), 2 * PROJ_THREADS, sizeof(PFSmemT<2>), s)(
          d, in, st, st.control + cl.tile_count, st.control + cl.block_sum);
      return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(project_forward_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(PFSmemT<1>));
    if (e != cudaSuccess) return e;
    pdl_launch(project_forward_stream_kernel<1>, min(d.B * d.NB, 4 * sm_count()), 2 * PROJ_THREADS, sizeof(PFSmemT<1>), s)(
        d, in, st, st.control + cl.tile_count, st.control + cl.block_sum);
    return cudaGetLastError();
  }
  dim3 grid(d.NB, d.B);
  pdl_launch(project_forward_kernel, grid, PROJ_THREADS, smem, s)(d, in, st, st.control + cl.tile_count,
                                                          st.control + cl.block_sum);
  return cudaGetLastError();
}

}  // namespace spf
