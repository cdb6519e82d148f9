// Per-Gaussian math shared by the sm_100a kernels (nvcc) and the host-side unit-test
// harness (g++, tests/hostmath).  Everything here is a pure function of its
// arguments; no memory traffic, no CUDA intrinsics.
//
// Replaces the per-Gaussian preprocess / preprocess-backward of the external
// `diff_gauss_pose` rasterizer that /root/reference/src/model/decoder/cuda_splatting.py:124-138
// calls (algorithm: SURVEY.md Appendix B; op order pinned by oracle/raster_oracle.py).
//
// Bit-exactness contract: spf::project_forward is written so that, compiled
// WITHOUT fused multiply-add contraction (nvcc -fmad=false / g++ -ffp-contract=off),
// it reproduces oracle/raster_oracle.py:preprocess bit for bit on the
// index-affecting outputs (depth bits, radius, tile rectangle).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SPF_HD __host__ __device__ __forceinline__
#else
#define SPF_HD inline
#endif

namespace spf {

constexpr int   TILE          = 16;
constexpr float NEAR_CULL     = 0.2f;
constexpr float COV_DILATION  = 0.3f;
constexpr float FOV_CLAMP     = 1.3f;
constexpr float ALPHA_MAX     = 0.99f;
constexpr float ALPHA_MIN     = 1.0f / 255.0f;
constexpr float T_STOP        = 1e-4f;
constexpr float EIG_FLOOR     = 0.1f;
constexpr float W_EPS         = 1e-7f;
constexpr int   MAX_SH_COEFFS = 25;

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
constexpr float SH_C2_0 = 1.0925484305920792f, SH_C2_1 = -1.0925484305920792f,
                SH_C2_2 = 0.31539156525252005f, SH_C2_3 = -1.0925484305920792f,
                SH_C2_4 = 0.5462742152960396f;
constexpr float SH_C3_0 = -0.5900435899266435f, SH_C3_1 = 2.890611442640554f,
                SH_C3_2 = -0.4570457994644658f, SH_C3_3 = 0.3731763325901154f,
                SH_C3_4 = -0.4570457994644658f, SH_C3_5 = 1.445305721320277f,
                SH_C3_6 = -0.5900435899266435f;
constexpr float SH_C4_0 = 2.5033429417967046f, SH_C4_1 = -1.7701307697799304f,
                SH_C4_2 = 0.9461746957575601f, SH_C4_3 = -0.6690465435572892f,
                SH_C4_4 = 0.10578554691520431f, SH_C4_5 = -0.6690465435572892f,
                SH_C4_6 = 0.47308734787878004f, SH_C4_7 = -1.7701307697799304f,
                SH_C4_8 = 0.6258357354491761f;

// View constants derived once per view (device: by one warp into shared memory).
struct ViewConsts {
  float V[16];      // viewmatrix as passed (row-vector convention), row-major V[4*i+j]
  float P[16];      // projmatrix as passed
  float tanx, tany, fx, fy, limx, limy;
  float Wf, Hf;
  float campos[3];
  float bg[3];
  float mod;        // scale_modifier
  int   W, H, gx, gy;
};

SPF_HD void make_view_consts(ViewConsts& vc, const float* V, const float* P, float tanx,
                             float tany, const float* bg, float mod, int W, int H) {
  for (int i = 0; i < 16; ++i) { vc.V[i] = V[i]; vc.P[i] = P[i]; }
  vc.tanx = tanx; vc.tany = tany;
  vc.Wf = (float)W; vc.Hf = (float)H;
  vc.fx = vc.Wf / (2.0f * tanx);
  vc.fy = vc.Hf / (2.0f * tany);
  vc.limx = FOV_CLAMP * tanx;
  vc.limy = FOV_CLAMP * tany;
  // campos_i = -sum_j tau_j A_ij  (A = V[:3,:3], tau = V[3,:3])
  for (int i = 0; i < 3; ++i)
    vc.campos[i] = -((V[12] * V[4 * i + 0] + V[13] * V[4 * i + 1]) + V[14] * V[4 * i + 2]);
  vc.bg[0] = bg[0]; vc.bg[1] = bg[1]; vc.bg[2] = bg[2];
  vc.mod = mod;
  vc.W = W; vc.H = H;
  vc.gx = (W + TILE - 1) / TILE;
  vc.gy = (H + TILE - 1) / TILE;
}

struct Projected {
  float px, py, depth;
  float conx, cony, conz;
  int   radius;            // 0 if culled
  int   rx0, ry0, rx1, ry1;
  int   tiles;             // 0 if culled
};

SPF_HD void quat_to_rot(const float q[4], float R[9]) {
  const float r = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1.0f - 2.0f * (y * y + z * z); R[1] = 2.0f * (x * y - r * z); R[2] = 2.0f * (x * z + r * y);
  R[3] = 2.0f * (x * y + r * z); R[4] = 1.0f - 2.0f * (x * x + z * z); R[5] = 2.0f * (y * z - r * x);
  R[6] = 2.0f * (x * z - r * y); R[7] = 2.0f * (y * z + r * x); R[8] = 1.0f - 2.0f * (x * x + y * y);
}

// Sigma = (R diag(mod*s)) (R diag(mod*s))^T, full 3x3 (symmetric), row-major.
SPF_HD void cov3d(const float s[3], float mod, const float q[4], float Sg[9]) {
  float R[9];
  quat_to_rot(q, R);
  float L[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) L[3 * i + j] = R[3 * i + j] * (mod * s[j]);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      Sg[3 * i + j] = (L[3 * i] * L[3 * j] + L[3 * i + 1] * L[3 * j + 1]) + L[3 * i + 2] * L[3 * j + 2];
}

SPF_HD int tile_clamp(float q, int g) {
  float c = fminf(fmaxf(q, 0.0f), (float)g);
  return (int)c;    // truncation; c is in [0, g]
}

// Forward projection of one Gaussian.  Returns false (and radius = tiles = 0) if culled.
SPF_HD bool project_forward(const ViewConsts& vc, const float m[3], const float s[3],
                            const float q[4], Projected& o) {
  const float* V = vc.V;
  const float* P = vc.P;
  o.radius = 0; o.tiles = 0; o.rx0 = o.ry0 = o.rx1 = o.ry1 = 0;
  const float tx = ((V[0] * m[0] + V[4] * m[1]) + V[8] * m[2]) + V[12];
  const float ty = ((V[1] * m[0] + V[5] * m[1]) + V[9] * m[2]) + V[13];
  const float tz = ((V[2] * m[0] + V[6] * m[1]) + V[10] * m[2]) + V[14];
  o.depth = tz;
  const float hx = ((tx * P[0] + ty * P[4]) + tz * P[8]) + P[12];
  const float hy = ((tx * P[1] + ty * P[5]) + tz * P[9]) + P[13];
  const float hw = ((tx * P[3] + ty * P[7]) + tz * P[11]) + P[15];
  const float p_w = 1.0f / (hw + W_EPS);
  const float ndcx = hx * p_w, ndcy = hy * p_w;
  o.px = ((ndcx + 1.0f) * vc.Wf - 1.0f) * 0.5f;
  o.py = ((ndcy + 1.0f) * vc.Hf - 1.0f) * 0.5f;
  o.conx = o.cony = o.conz = 0.0f;
  if (!(tz > NEAR_CULL)) return false;

  float Sg[9];
  cov3d(s, vc.mod, q, Sg);
  const float txc = fminf(vc.limx, fmaxf(-vc.limx, tx / tz)) * tz;
  const float tyc = fminf(vc.limy, fmaxf(-vc.limy, ty / tz)) * tz;
  const float tz2 = tz * tz;
  const float J00 = vc.fx / tz;
  const float J02 = -(vc.fx * txc) / tz2;
  const float J11 = vc.fy / tz;
  const float J12 = -(vc.fy * tyc) / tz2;
  float T0[3], T1[3], U0[3], U1[3];
  for (int i = 0; i < 3; ++i) {
    T0[i] = J00 * V[4 * i + 0] + J02 * V[4 * i + 2];
    T1[i] = J11 * V[4 * i + 1] + J12 * V[4 * i + 2];
  }
  for (int j = 0; j < 3; ++j) {
    U0[j] = (T0[0] * Sg[j] + T0[1] * Sg[3 + j]) + T0[2] * Sg[6 + j];
    U1[j] = (T1[0] * Sg[j] + T1[1] * Sg[3 + j]) + T1[2] * Sg[6 + j];
  }
  const float a = ((U0[0] * T0[0] + U0[1] * T0[1]) + U0[2] * T0[2]) + COV_DILATION;
  const float b = (U0[0] * T1[0] + U0[1] * T1[1]) + U0[2] * T1[2];
  const float c = ((U1[0] * T1[0] + U1[1] * T1[1]) + U1[2] * T1[2]) + COV_DILATION;
  const float det = a * c - b * b;
  if (det == 0.0f) return false;
  const float det_inv = 1.0f / det;
  o.conx = c * det_inv; o.cony = -b * det_inv; o.conz = a * det_inv;
  const float mid = 0.5f * (a + c);
  const float root = sqrtf(fmaxf(mid * mid - det, EIG_FLOOR));
  const float lam = fmaxf(mid + root, mid - root);
  const float radius_f = ceilf(3.0f * sqrtf(lam));
  // Non-finite projections are culled (oracle does the same).
  if (!(fabsf(radius_f) <= 3.0e38f) || !(fabsf(o.px) <= 3.0e38f) || !(fabsf(o.py) <= 3.0e38f))
    return false;
  const int rx0 = tile_clamp((o.px - radius_f) / (float)TILE, vc.gx);
  const int ry0 = tile_clamp((o.py - radius_f) / (float)TILE, vc.gy);
  const int rx1 = tile_clamp(((o.px + radius_f) + (float)(TILE - 1)) / (float)TILE, vc.gx);
  const int ry1 = tile_clamp(((o.py + radius_f) + (float)(TILE - 1)) / (float)TILE, vc.gy);
  const int area = (rx1 - rx0) * (ry1 - ry0);
  if (area <= 0) return false;
  o.radius = (int)radius_f;
  o.rx0 = rx0; o.ry0 = ry0; o.rx1 = rx1; o.ry1 = ry1;
  o.tiles = area;
  return true;
}

// Real SH basis, degrees 0..4 (K = (deg+1)^2 values written).
SPF_HD void sh_basis(int deg, float x, float y, float z, float* B) {
  B[0] = SH_C0;
  if (deg < 1) return;
  B[1] = -SH_C1 * y; B[2] = SH_C1 * z; B[3] = -SH_C1 * x;
  if (deg < 2) return;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  B[4] = SH_C2_0 * xy; B[5] = SH_C2_1 * yz; B[6] = SH_C2_2 * (2.0f * zz - xx - yy);
  B[7] = SH_C2_3 * xz; B[8] = SH_C2_4 * (xx - yy);
  if (deg < 3) return;
  B[9]  = SH_C3_0 * y * (3.0f * xx - yy);
  B[10] = SH_C3_1 * xy * z;
  B[11] = SH_C3_2 * y * (4.0f * zz - xx - yy);
  B[12] = SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
  B[13] = SH_C3_4 * x * (4.0f * zz - xx - yy);
  B[14] = SH_C3_5 * z * (xx - yy);
  B[15] = SH_C3_6 * x * (xx - 3.0f * yy);
  if (deg < 4) return;
  B[16] = SH_C4_0 * xy * (xx - yy);
  B[17] = SH_C4_1 * yz * (3.0f * xx - yy);
  B[18] = SH_C4_2 * xy * (7.0f * zz - 1.0f);
  B[19] = SH_C4_3 * yz * (7.0f * zz - 3.0f);
  B[20] = SH_C4_4 * (zz * (35.0f * zz - 30.0f) + 3.0f);
  B[21] = SH_C4_5 * xz * (7.0f * zz - 3.0f);
  B[22] = SH_C4_6 * (xx - yy) * (7.0f * zz - 1.0f);
  B[23] = SH_C4_7 * xz * (xx - 3.0f * yy);
  B[24] = SH_C4_8 * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy));
}

// dL/d(x,y,z) given v_k = dL/dB_k (x,y,z treated as independent variables).
SPF_HD void sh_basis_backward(int deg, float x, float y, float z, const float* v,
                              float& gx, float& gy, float& gz) {
  gx = gy = gz = 0.0f;
  if (deg < 1) return;
  gy += -SH_C1 * v[1]; gz += SH_C1 * v[2]; gx += -SH_C1 * v[3];
  if (deg < 2) return;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  gx += SH_C2_0 * y * v[4];            gy += SH_C2_0 * x * v[4];
  gy += SH_C2_1 * z * v[5];            gz += SH_C2_1 * y * v[5];
  gx += SH_C2_2 * -2.0f * x * v[6];    gy += SH_C2_2 * -2.0f * y * v[6];  gz += SH_C2_2 * 4.0f * z * v[6];
  gx += SH_C2_3 * z * v[7];            gz += SH_C2_3 * x * v[7];
  gx += SH_C2_4 * 2.0f * x * v[8];     gy += SH_C2_4 * -2.0f * y * v[8];
  if (deg < 3) return;
  gx += SH_C3_0 * 6.0f * xy * v[9];    gy += SH_C3_0 * 3.0f * (xx - yy) * v[9];
  gx += SH_C3_1 * yz * v[10];          gy += SH_C3_1 * xz * v[10];        gz += SH_C3_1 * xy * v[10];
  gx += SH_C3_2 * -2.0f * xy * v[11];  gy += SH_C3_2 * (4.0f * zz - xx - 3.0f * yy) * v[11];
  gz += SH_C3_2 * 8.0f * yz * v[11];
  gx += SH_C3_3 * -6.0f * xz * v[12];  gy += SH_C3_3 * -6.0f * yz * v[12];
  gz += SH_C3_3 * (6.0f * zz - 3.0f * xx - 3.0f * yy) * v[12];
  gx += SH_C3_4 * (4.0f * zz - 3.0f * xx - yy) * v[13]; gy += SH_C3_4 * -2.0f * xy * v[13];
  gz += SH_C3_4 * 8.0f * xz * v[13];
  gx += SH_C3_5 * 2.0f * xz * v[14];   gy += SH_C3_5 * -2.0f * yz * v[14]; gz += SH_C3_5 * (xx - yy) * v[14];
  gx += SH_C3_6 * 3.0f * (xx - yy) * v[15]; gy += SH_C3_6 * -6.0f * xy * v[15];
  if (deg < 4) return;
  gx += SH_C4_0 * (3.0f * xx * y - yy * y) * v[16];  gy += SH_C4_0 * (xx * x - 3.0f * x * yy) * v[16];
  gx += SH_C4_1 * 6.0f * xy * z * v[17];  gy += SH_C4_1 * 3.0f * z * (xx - yy) * v[17];
  gz += SH_C4_1 * y * (3.0f * xx - yy) * v[17];
  gx += SH_C4_2 * y * (7.0f * zz - 1.0f) * v[18];  gy += SH_C4_2 * x * (7.0f * zz - 1.0f) * v[18];
  gz += SH_C4_2 * 14.0f * xy * z * v[18];
  gy += SH_C4_3 * z * (7.0f * zz - 3.0f) * v[19];  gz += SH_C4_3 * y * (21.0f * zz - 3.0f) * v[19];
  gz += SH_C4_4 * z * (140.0f * zz - 60.0f) * v[20];
  gx += SH_C4_5 * z * (7.0f * zz - 3.0f) * v[21];  gz += SH_C4_5 * x * (21.0f * zz - 3.0f) * v[21];
  gx += SH_C4_6 * 2.0f * x * (7.0f * zz - 1.0f) * v[22]; gy += SH_C4_6 * -2.0f * y * (7.0f * zz - 1.0f) * v[22];
  gz += SH_C4_6 * 14.0f * z * (xx - yy) * v[22];
  gx += SH_C4_7 * 3.0f * z * (xx - yy) * v[23];  gy += SH_C4_7 * -6.0f * xy * z * v[23];
  gz += SH_C4_7 * x * (xx - 3.0f * yy) * v[23];
  gx += SH_C4_8 * 4.0f * x * (xx - 3.0f * yy) * v[24];  gy += SH_C4_8 * 4.0f * y * (yy - 3.0f * xx) * v[24];
}

// SH colour of one Gaussian without a basis array: rgb[c] = sum_k B_k(x,y,z) * sh[k*sk + c*sc]  (explicit fmaf so the
// colour path keeps fused multiply-adds even in the -fmad=false translation unit; not index-affecting).
// MASKED: every coefficient is first multiplied by the encoder's per-degree SH mask (1 for degree 0, 0.1 * 0.25^d above,
// gaussian_adapter.py:42-48) -- the raw-head input path, where the masked coefficients never exist in memory; the
// product is rounded to fp32 before the fma, exactly like the stand-alone adapter stores it.
constexpr float SH_MASK_1 = 0.1f * 0.25f, SH_MASK_2 = 0.1f * 0.0625f, SH_MASK_3 = 0.1f * 0.015625f,
                SH_MASK_4 = 0.1f * 0.00390625f;

template <bool MASKED>
SPF_HD void sh_eval_fused_t(int deg, float x, float y, float z, const float* sh, int sk, int sc, float rgb[3]) {
  float r0 = 0.0f, r1 = 0.0f, r2 = 0.0f;
  float mk = 1.0f;
#define SPF_SH_EVAL(k, B)                                                                  \
  {                                                                                        \
    const float b_ = (B);                                                                  \
    const int i_ = (k) * sk;                                                               \
    const float s0_ = MASKED ? sh[i_] * mk : sh[i_];                                       \
    const float s1_ = MASKED ? sh[i_ + sc] * mk : sh[i_ + sc];                             \
    const float s2_ = MASKED ? sh[i_ + 2 * sc] * mk : sh[i_ + 2 * sc];                     \
    r0 = fmaf(b_, s0_, r0); r1 = fmaf(b_, s1_, r1); r2 = fmaf(b_, s2_, r2);                \
  }
  SPF_SH_EVAL(0, SH_C0)
  if (deg >= 1) {
    mk = SH_MASK_1;
    SPF_SH_EVAL(1, -SH_C1 * y) SPF_SH_EVAL(2, SH_C1 * z) SPF_SH_EVAL(3, -SH_C1 * x)
    if (deg >= 2) {
      mk = SH_MASK_2;
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      SPF_SH_EVAL(4, SH_C2_0 * xy) SPF_SH_EVAL(5, SH_C2_1 * yz) SPF_SH_EVAL(6, SH_C2_2 * (2.0f * zz - xx - yy))
      SPF_SH_EVAL(7, SH_C2_3 * xz) SPF_SH_EVAL(8, SH_C2_4 * (xx - yy))
      if (deg >= 3) {
        mk = SH_MASK_3;
        SPF_SH_EVAL(9, SH_C3_0 * y * (3.0f * xx - yy)) SPF_SH_EVAL(10, SH_C3_1 * xy * z)
        SPF_SH_EVAL(11, SH_C3_2 * y * (4.0f * zz - xx - yy))
        SPF_SH_EVAL(12, SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy))
        SPF_SH_EVAL(13, SH_C3_4 * x * (4.0f * zz - xx - yy)) SPF_SH_EVAL(14, SH_C3_5 * z * (xx - yy))
        SPF_SH_EVAL(15, SH_C3_6 * x * (xx - 3.0f * yy))
        if (deg >= 4) {
          mk = SH_MASK_4;
          SPF_SH_EVAL(16, SH_C4_0 * xy * (xx - yy)) SPF_SH_EVAL(17, SH_C4_1 * yz * (3.0f * xx - yy))
          SPF_SH_EVAL(18, SH_C4_2 * xy * (7.0f * zz - 1.0f)) SPF_SH_EVAL(19, SH_C4_3 * yz * (7.0f * zz - 3.0f))
          SPF_SH_EVAL(20, SH_C4_4 * (zz * (35.0f * zz - 30.0f) + 3.0f)) SPF_SH_EVAL(21, SH_C4_5 * xz * (7.0f * zz - 3.0f))
          SPF_SH_EVAL(22, SH_C4_6 * (xx - yy) * (7.0f * zz - 1.0f)) SPF_SH_EVAL(23, SH_C4_7 * xz * (xx - 3.0f * yy))
          SPF_SH_EVAL(24, SH_C4_8 * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy)))
        }
      }
    }
  }
#undef SPF_SH_EVAL
  rgb[0] = r0; rgb[1] = r1; rgb[2] = r2;
  (void)mk;
}

SPF_HD void sh_eval_fused(int deg, float x, float y, float z, const float* sh, int sk, int sc, float rgb[3]) {
  sh_eval_fused_t<false>(deg, x, y, z, sh, sk, sc, rgb);
}

// Fused single pass over the SH coefficients for the backward (no Bk[] / vk[] arrays => few live registers):
// for every k:  b = B_k(x,y,z);  v = sum_c sh[k,c]*gm[c];  dsh[k,c] (+)= b*gm[c];  (gx,gy,gz) += v * dB_k/d(x,y,z).
// Element (k,c) of sh / dsh lives at  k*sk + c*sc  (sk=3,sc=1 for [K,3]; sk=1,sc=Kstore for [3,K]).
// dsh may alias sh (each element is read before it is written).  gm = dL/drgb with the clamp mask applied.
// MASKED (raw-head input): sh holds the encoder's UNMASKED coefficients -- every read is multiplied by the degree's SH
// mask (product rounded to fp32 first, like the masked tensor the stand-alone adapter stores), and without ACCUM the
// gradient leaves already multiplied by the mask (d raw = d masked * mask); with ACCUM the caller applies the mask
// once after the last view.
#if defined(__CUDA_ARCH__)
#define SPF_MUL_RN(a, b) __fmul_rn((a), (b))
#else
#define SPF_MUL_RN(a, b) ((a) * (b))
#endif
template <bool ACCUM, bool MASKED = false>
SPF_HD void sh_backward_fused(int deg, float x, float y, float z, const float* sh, float* dsh, int sk, int sc,
                              const float gm[3], float& gx, float& gy, float& gz) {
  gx = gy = gz = 0.0f;
  float mk = 1.0f;
#define SPF_SH_TERM(k, B, DX, DY, DZ)                                                      \
  {                                                                                        \
    const float b_ = (B);                                                                  \
    const int i_ = (k) * sk;                                                               \
    const float s0_ = MASKED ? SPF_MUL_RN(sh[i_], mk) : sh[i_];                            \
    const float s1_ = MASKED ? SPF_MUL_RN(sh[i_ + sc], mk) : sh[i_ + sc];                  \
    const float s2_ = MASKED ? SPF_MUL_RN(sh[i_ + 2 * sc], mk) : sh[i_ + 2 * sc];          \
    const float v_ = (s0_ * gm[0] + s1_ * gm[1]) + s2_ * gm[2];                            \
    if (ACCUM) { dsh[i_] += b_ * gm[0]; dsh[i_ + sc] += b_ * gm[1]; dsh[i_ + 2 * sc] += b_ * gm[2]; } \
    else if (MASKED) {                                                                     \
      dsh[i_] = SPF_MUL_RN(SPF_MUL_RN(b_, gm[0]), mk); dsh[i_ + sc] = SPF_MUL_RN(SPF_MUL_RN(b_, gm[1]), mk); \
      dsh[i_ + 2 * sc] = SPF_MUL_RN(SPF_MUL_RN(b_, gm[2]), mk);                            \
    } else { dsh[i_] = b_ * gm[0]; dsh[i_ + sc] = b_ * gm[1]; dsh[i_ + 2 * sc] = b_ * gm[2]; } \
    gx += v_ * (DX); gy += v_ * (DY); gz += v_ * (DZ);                                     \
  }
  SPF_SH_TERM(0, SH_C0, 0.0f, 0.0f, 0.0f)
  if (deg < 1) return;
  mk = SH_MASK_1;
  SPF_SH_TERM(1, -SH_C1 * y, 0.0f, -SH_C1, 0.0f)
  SPF_SH_TERM(2, SH_C1 * z, 0.0f, 0.0f, SH_C1)
  SPF_SH_TERM(3, -SH_C1 * x, -SH_C1, 0.0f, 0.0f)
  if (deg < 2) return;
  mk = SH_MASK_2;
  const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  SPF_SH_TERM(4, SH_C2_0 * xy, SH_C2_0 * y, SH_C2_0 * x, 0.0f)
  SPF_SH_TERM(5, SH_C2_1 * yz, 0.0f, SH_C2_1 * z, SH_C2_1 * y)
  SPF_SH_TERM(6, SH_C2_2 * (2.0f * zz - xx - yy), SH_C2_2 * -2.0f * x, SH_C2_2 * -2.0f * y, SH_C2_2 * 4.0f * z)
  SPF_SH_TERM(7, SH_C2_3 * xz, SH_C2_3 * z, 0.0f, SH_C2_3 * x)
  SPF_SH_TERM(8, SH_C2_4 * (xx - yy), SH_C2_4 * 2.0f * x, SH_C2_4 * -2.0f * y, 0.0f)
  if (deg < 3) return;
  mk = SH_MASK_3;
  SPF_SH_TERM(9, SH_C3_0 * y * (3.0f * xx - yy), SH_C3_0 * 6.0f * xy, SH_C3_0 * 3.0f * (xx - yy), 0.0f)
  SPF_SH_TERM(10, SH_C3_1 * xy * z, SH_C3_1 * yz, SH_C3_1 * xz, SH_C3_1 * xy)
  SPF_SH_TERM(11, SH_C3_2 * y * (4.0f * zz - xx - yy), SH_C3_2 * -2.0f * xy, SH_C3_2 * (4.0f * zz - xx - 3.0f * yy),
              SH_C3_2 * 8.0f * yz)
  SPF_SH_TERM(12, SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy), SH_C3_3 * -6.0f * xz, SH_C3_3 * -6.0f * yz,
              SH_C3_3 * (6.0f * zz - 3.0f * xx - 3.0f * yy))
  SPF_SH_TERM(13, SH_C3_4 * x * (4.0f * zz - xx - yy), SH_C3_4 * (4.0f * zz - 3.0f * xx - yy), SH_C3_4 * -2.0f * xy,
              SH_C3_4 * 8.0f * xz)
  SPF_SH_TERM(14, SH_C3_5 * z * (xx - yy), SH_C3_5 * 2.0f * xz, SH_C3_5 * -2.0f * yz, SH_C3_5 * (xx - yy))
  SPF_SH_TERM(15, SH_C3_6 * x * (xx - 3.0f * yy), SH_C3_6 * 3.0f * (xx - yy), SH_C3_6 * -6.0f * xy, 0.0f)
  if (deg < 4) return;
  mk = SH_MASK_4;
  SPF_SH_TERM(16, SH_C4_0 * xy * (xx - yy), SH_C4_0 * (3.0f * xx * y - yy * y), SH_C4_0 * (xx * x - 3.0f * x * yy), 0.0f)
  SPF_SH_TERM(17, SH_C4_1 * yz * (3.0f * xx - yy), SH_C4_1 * 6.0f * xy * z, SH_C4_1 * 3.0f * z * (xx - yy),
              SH_C4_1 * y * (3.0f * xx - yy))
  SPF_SH_TERM(18, SH_C4_2 * xy * (7.0f * zz - 1.0f), SH_C4_2 * y * (7.0f * zz - 1.0f), SH_C4_2 * x * (7.0f * zz - 1.0f),
              SH_C4_2 * 14.0f * xy * z)
  SPF_SH_TERM(19, SH_C4_3 * yz * (7.0f * zz - 3.0f), 0.0f, SH_C4_3 * z * (7.0f * zz - 3.0f),
              SH_C4_3 * y * (21.0f * zz - 3.0f))
  SPF_SH_TERM(20, SH_C4_4 * (zz * (35.0f * zz - 30.0f) + 3.0f), 0.0f, 0.0f, SH_C4_4 * z * (140.0f * zz - 60.0f))
  SPF_SH_TERM(21, SH_C4_5 * xz * (7.0f * zz - 3.0f), SH_C4_5 * z * (7.0f * zz - 3.0f), 0.0f,
              SH_C4_5 * x * (21.0f * zz - 3.0f))
  SPF_SH_TERM(22, SH_C4_6 * (xx - yy) * (7.0f * zz - 1.0f), SH_C4_6 * 2.0f * x * (7.0f * zz - 1.0f),
              SH_C4_6 * -2.0f * y * (7.0f * zz - 1.0f), SH_C4_6 * 14.0f * z * (xx - yy))
  SPF_SH_TERM(23, SH_C4_7 * xz * (xx - 3.0f * yy), SH_C4_7 * 3.0f * z * (xx - yy), SH_C4_7 * -6.0f * xy * z,
              SH_C4_7 * x * (xx - 3.0f * yy))
  SPF_SH_TERM(24, SH_C4_8 * (xx * (xx - 3.0f * yy) - yy * (3.0f * xx - yy)), SH_C4_8 * 4.0f * x * (xx - 3.0f * yy),
              SH_C4_8 * 4.0f * y * (yy - 3.0f * xx), 0.0f)
#undef SPF_SH_TERM
  (void)mk;
}

// Upstream 2-D gradients of one Gaussian in one view (what blend-backward reduces).
struct Grad2D {
  float dpx, dpy;             // dL/d(pixel-space mean)
  float dconx, dcony, dconz;  // dL/dconic (cony = the single off-diagonal variable)
  float dopacity;
  float drgb[3];
  float ddepth;
};

// Per-Gaussian, per-view gradient wrt the 3-D parameters and the pose.
struct Grad3D {
  float dm[3];
  float ds[3];
  float dq[4];
  float dA[9];     // dL/dV[i][j], i,j<3 (row-major 3x3)
  float dtau[3];   // dL/dV[3][j]
  float dcam[3];   // dL/dcampos (folded into dA/dtau by fold_campos_grad)
};

// Backward of project_forward (+ the direction part of the SH colour).
//   g_dir : dL/d(unit direction) from sh_basis_backward (zero if colours are precomputed
//           or SH direction gradients are disabled)
//   cov_grad : include the covariance path's contribution to means and pose.
// All outputs are ACCUMULATED (+=) so one Gaussian can sum several views.
SPF_HD void project_backward_geom(const ViewConsts& vc, const float m[3], const float s[3],
                                  const float q[4], const Grad2D& g, bool cov_grad, Grad3D& out) {
  const float* V = vc.V;
  const float* P = vc.P;
  const float tx = V[0] * m[0] + V[4] * m[1] + V[8] * m[2] + V[12];
  const float ty = V[1] * m[0] + V[5] * m[1] + V[9] * m[2] + V[13];
  const float tz = V[2] * m[0] + V[6] * m[1] + V[10] * m[2] + V[14];

  float R[9];
  quat_to_rot(q, R);
  float sm[3] = {vc.mod * s[0], vc.mod * s[1], vc.mod * s[2]};
  float L[9], Sg[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) L[3 * i + j] = R[3 * i + j] * sm[j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      Sg[3 * i + j] = L[3 * i] * L[3 * j] + L[3 * i + 1] * L[3 * j + 1] + L[3 * i + 2] * L[3 * j + 2];

  const float itz = 1.0f / tz;
  const float rx = tx * itz, ry = ty * itz;
  const bool inx = (rx >= -vc.limx) && (rx <= vc.limx);
  const bool iny = (ry >= -vc.limy) && (ry <= vc.limy);
  const float ux = fminf(vc.limx, fmaxf(-vc.limx, rx));
  const float uy = fminf(vc.limy, fmaxf(-vc.limy, ry));
  const float J00 = vc.fx * itz, J11 = vc.fy * itz;
  const float J02 = -vc.fx * ux * itz, J12 = -vc.fy * uy * itz;
  float T0[3], T1[3], ST0[3], ST1[3];
  for (int i = 0; i < 3; ++i) {
    T0[i] = J00 * V[4 * i + 0] + J02 * V[4 * i + 2];
    T1[i] = J11 * V[4 * i + 1] + J12 * V[4 * i + 2];
  }
  for (int i = 0; i < 3; ++i) {
    ST0[i] = Sg[3 * i] * T0[0] + Sg[3 * i + 1] * T0[1] + Sg[3 * i + 2] * T0[2];
    ST1[i] = Sg[3 * i] * T1[0] + Sg[3 * i + 1] * T1[1] + Sg[3 * i + 2] * T1[2];
  }
  const float a = T0[0] * ST0[0] + T0[1] * ST0[1] + T0[2] * ST0[2] + COV_DILATION;
  const float b = T0[0] * ST1[0] + T0[1] * ST1[1] + T0[2] * ST1[2];
  const float c = T1[0] * ST1[0] + T1[1] * ST1[1] + T1[2] * ST1[2] + COV_DILATION;
  const float det = a * c - b * b;
  const float di = 1.0f / det, di2 = di * di;

  // conic -> (a, b, c)
  const float ga = di2 * (-c * c * g.dconx + b * c * g.dcony - b * b * g.dconz);
  const float gc = di2 * (-b * b * g.dconx + a * b * g.dcony - a * a * g.dconz);
  const float gb = di2 * (2.0f * b * c * g.dconx - (a * c + b * b) * g.dcony + 2.0f * a * b * g.dconz);

  // (a,b,c) -> Sigma (symmetrised gradient Gs = G + G^T) and T
  float Gs[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      Gs[3 * i + j] = 2.0f * ga * T0[i] * T0[j] + gb * (T0[i] * T1[j] + T1[i] * T0[j]) +
                      2.0f * gc * T1[i] * T1[j];
  // dL/dL = Gs L ; dL/ds_j = mod * sum_i H_ij R_ij ; dL/dR_ij = H_ij sm_j
  float gR[9];
  for (int j = 0; j < 3; ++j) {
    float acc = 0.0f;
    for (int i = 0; i < 3; ++i) {
      const float Hij = Gs[3 * i] * L[j] + Gs[3 * i + 1] * L[3 + j] + Gs[3 * i + 2] * L[6 + j];
      acc += Hij * R[3 * i + j];
      gR[3 * i + j] = Hij * sm[j];
    }
    out.ds[j] += vc.mod * acc;
  }
  {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    out.dq[0] += 2.0f * (-z * gR[1] + y * gR[2] + z * gR[3] - x * gR[5] - y * gR[6] + x * gR[7]);
    out.dq[1] += 2.0f * (y * gR[1] + z * gR[2] + y * gR[3] - 2.0f * x * gR[4] - r * gR[5] + z * gR[6] +
                         r * gR[7] - 2.0f * x * gR[8]);
    out.dq[2] += 2.0f * (-2.0f * y * gR[0] + x * gR[1] + r * gR[2] + x * gR[3] + z * gR[5] - r * gR[6] +
                         z * gR[7] - 2.0f * y * gR[8]);
    out.dq[3] += 2.0f * (-2.0f * z * gR[0] - r * gR[1] + x * gR[2] + r * gR[3] - 2.0f * z * gR[4] +
                         y * gR[5] + x * gR[6] + y * gR[7]);
  }

  float gt[3] = {0.0f, 0.0f, 0.0f};   // dL/dt
  if (cov_grad) {
    float gT0[3], gT1[3];
    for (int i = 0; i < 3; ++i) {
      gT0[i] = 2.0f * ga * ST0[i] + gb * ST1[i];
      gT1[i] = 2.0f * gc * ST1[i] + gb * ST0[i];
    }
    float gJ00 = 0.f, gJ02 = 0.f, gJ11 = 0.f, gJ12 = 0.f;
    for (int i = 0; i < 3; ++i) {
      gJ00 += gT0[i] * V[4 * i + 0];
      gJ02 += gT0[i] * V[4 * i + 2];
      gJ11 += gT1[i] * V[4 * i + 1];
      gJ12 += gT1[i] * V[4 * i + 2];
      out.dA[3 * i + 0] += gT0[i] * J00;
      out.dA[3 * i + 1] += gT1[i] * J11;
      out.dA[3 * i + 2] += gT0[i] * J02 + gT1[i] * J12;
    }
    // J00 = fx/tz, J11 = fy/tz, J02 = -fx*ux/tz, J12 = -fy*uy/tz, ux = clamp(tx/tz)
    const float itz2 = itz * itz;
    float gux = gJ02 * (-vc.fx * itz);
    float guy = gJ12 * (-vc.fy * itz);
    gt[2] += -vc.fx * itz2 * gJ00 - vc.fy * itz2 * gJ11 + vc.fx * ux * itz2 * gJ02 + vc.fy * uy * itz2 * gJ12;
    if (inx) { gt[0] += gux * itz; gt[2] += -gux * tx * itz2; }
    if (iny) { gt[1] += guy * itz; gt[2] += -guy * ty * itz2; }
  }

  // pixel mean -> ndc -> p_hom -> t
  {
    const float hx = tx * P[0] + ty * P[4] + tz * P[8] + P[12];
    const float hy = tx * P[1] + ty * P[5] + tz * P[9] + P[13];
    const float hw = tx * P[3] + ty * P[7] + tz * P[11] + P[15];
    const float p_w = 1.0f / (hw + W_EPS);
    const float gnx = g.dpx * 0.5f * vc.Wf, gny = g.dpy * 0.5f * vc.Hf;
    const float ghx = gnx * p_w, ghy = gny * p_w;
    const float ghw = -(gnx * hx + gny * hy) * p_w * p_w;
    for (int i = 0; i < 3; ++i) gt[i] += ghx * P[4 * i + 0] + ghy * P[4 * i + 1] + ghw * P[4 * i + 3];
  }
  gt[2] += g.ddepth;

  // t = m A + tau
  for (int i = 0; i < 3; ++i) {
    out.dm[i] += V[4 * i + 0] * gt[0] + V[4 * i + 1] * gt[1] + V[4 * i + 2] * gt[2];
    for (int j = 0; j < 3; ++j) out.dA[3 * i + j] += m[i] * gt[j];
  }
  for (int j = 0; j < 3; ++j) out.dtau[j] += gt[j];

}

// Backward of the SH view direction  dir = d/|d|, d = m - campos : g_dir = dL/d(dir) (from the SH basis).
SPF_HD void view_dir_backward(const ViewConsts& vc, const float m[3], const float g_dir[3], Grad3D& out) {
  const float dx = m[0] - vc.campos[0], dy = m[1] - vc.campos[1], dz = m[2] - vc.campos[2];
  const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
  const float nx = dx * inv, ny = dy * inv, nz = dz * inv;
  const float dot = nx * g_dir[0] + ny * g_dir[1] + nz * g_dir[2];
  const float gd0 = (g_dir[0] - nx * dot) * inv;
  const float gd1 = (g_dir[1] - ny * dot) * inv;
  const float gd2 = (g_dir[2] - nz * dot) * inv;
  out.dm[0] += gd0; out.dm[1] += gd1; out.dm[2] += gd2;
  out.dcam[0] -= gd0; out.dcam[1] -= gd1; out.dcam[2] -= gd2;
}

SPF_HD void project_backward(const ViewConsts& vc, const float m[3], const float s[3],
                             const float q[4], const Grad2D& g, const float g_dir[3],
                             bool cov_grad, Grad3D& out) {
  project_backward_geom(vc, m, s, q, g, cov_grad, out);
  view_dir_backward(vc, m, g_dir, out);
}

// campos_i = -sum_j tau_j A_ij  =>  fold dL/dcampos into dL/dA and dL/dtau.
SPF_HD void fold_campos_grad(const float* V, const float dcam[3], float dA[9], float dtau[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      dA[3 * i + j] += -dcam[i] * V[12 + j];
      dtau[j] += -dcam[i] * V[4 * i + j];
    }
}

}  // namespace spf
