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
#include "spf_device.cuh"
#include "spf_kernels.h"
#include "spf_math.h"

namespace spf {

__global__ void __launch_bounds__(PROJ_THREADS)
project_backward_kernel(Dims d, SpfRasterIn in, SpfRasterState st, SpfRasterGradIn gin) {
  extern __shared__ __align__(16) float smem[];
  __shared__ ViewConsts vc;
  __shared__ float pose_warp[PROJ_THREADS / 32][15];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int scene = blockIdx.y;
  const int g0 = blockIdx.x * PROJ_THREADS;
  const int g = g0 + tid;
  const int nvalid = min(PROJ_THREADS, d.P - g0);
  const int row = 3 * in.sh_coeffs;
  const int stride = (row & 1) ? row : row + 1;
  float* sh_s = smem;                             // input SH  [128][stride]
  float* dsh_s = smem + PROJ_THREADS * stride;    // SH grads  [128][stride]
  const bool use_sh = in.shs != nullptr;
  const bool ck = (d.flags & SPF_FLAG_SH_LAYOUT_CK) != 0;
  const bool cov_grad = !(d.flags & SPF_FLAG_NO_COV_GRAD);
  const bool sh_grad = !(d.flags & SPF_FLAG_NO_SH_GRAD);

  if (use_sh) {
    const float* src = in.shs + ((size_t)scene * d.P + g0) * row;
    block_copy_g2s(sh_s, src, nvalid * row, row, stride, tid, PROJ_THREADS);
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
    __syncthreads();   // previous view's vc / pose_warp fully consumed; SH staging complete
    if (tid == 0) {
      float V[16], Pm[16], bg[3];
      for (int i = 0; i < 16; ++i) { V[i] = in.viewmatrix[view * 16 + i]; Pm[i] = in.projmatrix[view * 16 + i]; }
      for (int i = 0; i < 3; ++i) bg[i] = in.bg[view * 3 + i];
      make_view_consts(vc, V, Pm, in.tanfov[view * 2], in.tanfov[view * 2 + 1], bg, d.mod, d.W, d.H);
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

    const size_t vg = (size_t)view * d.P + g;
    const int tiles = (g < d.P) ? st.tiles_touched[vg] : 0;
    if (tiles > 0) {
      // sum this Gaussian's duplicate records (contiguous slots)
      float a[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) a[k] = 0.0f;
      const float4* rec = reinterpret_cast<const float4*>(gin.dup_grad) + 3 * (size_t)st.dup_offset[vg];
      for (int j = 0; j < tiles; ++j) {
        const float4 r0 = rec[3 * j], r1 = rec[3 * j + 1], r2 = rec[3 * j + 2];
        a[0] += r0.x; a[1] += r0.y; a[2] += r0.z; a[3] += r0.w;
        a[4] += r1.x; a[5] += r1.y; a[6] += r1.z; a[7] += r1.w;
        a[8] += r2.x; a[9] += r2.y;
      }
      Grad2D g2;
      g2.dpx = a[0]; g2.dpy = a[1]; g2.dconx = a[2]; g2.dcony = a[3]; g2.dconz = a[4];
      g2.dopacity = a[5]; g2.drgb[0] = a[6]; g2.drgb[1] = a[7]; g2.drgb[2] = a[8]; g2.ddepth = a[9];
      dop += g2.dopacity;
      if (gin.dL_dmeans2D) {
        float* o2 = gin.dL_dmeans2D + vg * 3;
        o2[0] = g2.dpx * 0.5f * vc.Wf; o2[1] = g2.dpy * 0.5f * vc.Hf; o2[2] = 0.0f;
      }
      float m[3] = {m_in[0] * ps, m_in[1] * ps, m_in[2] * ps};
      float s[3] = {s_in[0] * ps, s_in[1] * ps, s_in[2] * ps};
      float gdir[3] = {0.f, 0.f, 0.f};
      if (use_sh) {
        const float dx = m[0] - vc.campos[0], dy = m[1] - vc.campos[1], dz = m[2] - vc.campos[2];
        const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        const float x = dx * inv, y = dy * inv, z = dz * inv;
        float Bk[MAX_SH_COEFFS], vk[MAX_SH_COEFFS];
        sh_basis(d.deg, x, y, z, Bk);
        const float* mysh = sh_s + tid * stride;
        float* mydsh = dsh_s + tid * stride;
        float gm[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float acc = 0.0f;
          for (int k = 0; k < d.K; ++k) acc += Bk[k] * (ck ? mysh[c * in.sh_coeffs + k] : mysh[k * 3 + c]);
          gm[c] = (acc + 0.5f) < 0.0f ? 0.0f : g2.drgb[c];   // clamp mask
        }
        for (int k = 0; k < d.K; ++k) {
          float vv = 0.0f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int idx = ck ? (c * in.sh_coeffs + k) : (k * 3 + c);
            mydsh[idx] += Bk[k] * gm[c];
            vv += mysh[idx] * gm[c];
          }
          vk[k] = vv;
        }
        if (sh_grad) sh_basis_backward(d.deg, x, y, z, vk, gdir[0], gdir[1], gdir[2]);
      } else {
        dcol[0] += g2.drgb[0]; dcol[1] += g2.drgb[1]; dcol[2] += g2.drgb[2];
      }
      project_backward(vc, m, s, q, g2, gdir, cov_grad, o);
#pragma unroll
      for (int i = 0; i < 3; ++i) { dm[i] += o.dm[i] * ps; ds[i] += o.ds[i] * ps; }
#pragma unroll
      for (int i = 0; i < 4; ++i) dq[i] += o.dq[i];
    } else if (g < d.P && gin.dL_dmeans2D) {
      float* o2 = gin.dL_dmeans2D + vg * 3;
      o2[0] = 0.f; o2[1] = 0.f; o2[2] = 0.f;
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
  if (use_sh && gin.dL_dshs) {
    __syncthreads();
    float* dst = gin.dL_dshs + ((size_t)scene * d.P + g0) * row;
    block_copy_s2g(dst, dsh_s, nvalid * row, row, stride, tid, PROJ_THREADS);
  }
}

// K9: dL/dviewmatrix[view] = fixed-order sum of block partials, campos gradient folded in.
__global__ void __launch_bounds__(256)
pose_reduce_kernel(Dims d, const float* __restrict__ viewmatrix, const float* __restrict__ partial,
                   float* __restrict__ dV) {
  __shared__ float red[256][16];
  const int view = blockIdx.x, tid = threadIdx.x;
  float acc[15];
#pragma unroll
  for (int i = 0; i < 15; ++i) acc[i] = 0.0f;
  for (int b = tid; b < d.NB; b += 256) {
    const float* p = partial + ((size_t)view * d.NB + b) * 16;
#pragma unroll
    for (int i = 0; i < 15; ++i) acc[i] += p[i];
  }
#pragma unroll
  for (int i = 0; i < 15; ++i) red[tid][i] = acc[i];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s)
      for (int i = 0; i < 15; ++i) red[tid][i] += red[tid + s][i];
    __syncthreads();
  }
  if (tid == 0) {
    float dA[9], dtau[3], dcam[3], V[16];
    for (int i = 0; i < 9; ++i) dA[i] = red[0][i];
    for (int i = 0; i < 3; ++i) { dtau[i] = red[0][9 + i]; dcam[i] = red[0][12 + i]; }
    for (int i = 0; i < 16; ++i) V[i] = viewmatrix[view * 16 + i];
    fold_campos_grad(V, dcam, dA, dtau);
    float* o = dV + view * 16;
    for (int i = 0; i < 16; ++i) o[i] = 0.0f;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) o[4 * i + j] = dA[3 * i + j];
    for (int j = 0; j < 3; ++j) o[12 + j] = dtau[j];
  }
}

cudaError_t launch_project_backward(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st,
                                    const SpfRasterGradIn& gin, cudaStream_t s) {
  const int row = 3 * in.sh_coeffs;
  const int stride = (row & 1) ? row : row + 1;
  const size_t smem = in.shs ? (size_t)2 * PROJ_THREADS * stride * sizeof(float) : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(project_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid(d.NB, d.S);
  project_backward_kernel<<<grid, PROJ_THREADS, smem, s>>>(d, in, st, gin);
  return cudaGetLastError();
}

cudaError_t launch_pose_reduce(const Dims& d, const SpfRasterIn& in, const SpfRasterGradIn& gin, cudaStream_t s) {
  pose_reduce_kernel<<<d.B, 256, 0, s>>>(d, in.viewmatrix, gin.pose_partial, gin.dL_dviewmatrix);
  return cudaGetLastError();
}

}  // namespace spf
