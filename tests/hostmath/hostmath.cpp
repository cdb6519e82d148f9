// Host (g++) build of spfsplatv2_b200/csrc/spf_math.h for CPU unit tests of the
// per-Gaussian math.  TEST INFRASTRUCTURE ONLY: the product never loads this.
// Compile with -ffp-contract=off so that the index path is comparable bit-for-bit
// with oracle/raster_oracle.py (and with the -fmad=false CUDA build).
#include "../../spfsplatv2_b200/csrc/spf_math.h"

using namespace spf;

extern "C" {

// means[P,3] scales[P,3] quats[P,4] ; out_f [P,6] = px,py,depth,conx,cony,conz ; out_i [P,6] = radius,rx0,ry0,rx1,ry1,tiles
void hm_project_forward(int P, const float* means, const float* scales, const float* quats,
                        const float* V, const float* Pm, float tanx, float tany, float mod,
                        int W, int H, float* out_f, int* out_i) {
  ViewConsts vc;
  float bg[3] = {0, 0, 0};
  make_view_consts(vc, V, Pm, tanx, tany, bg, mod, W, H);
  for (int g = 0; g < P; ++g) {
    Projected o;
    project_forward(vc, means + 3 * g, scales + 3 * g, quats + 4 * g, o);
    float* f = out_f + 6 * g;
    f[0] = o.px; f[1] = o.py; f[2] = o.depth; f[3] = o.conx; f[4] = o.cony; f[5] = o.conz;
    int* ii = out_i + 6 * g;
    ii[0] = o.radius; ii[1] = o.rx0; ii[2] = o.ry0; ii[3] = o.rx1; ii[4] = o.ry1; ii[5] = o.tiles;
  }
}

// SH colour forward: shs [P,K,3]; rgb [P,3] (clamped at 0, +0.5)
void hm_sh_forward(int P, int deg, const float* means, const float* V, const float* shs, float* rgb) {
  ViewConsts vc;
  float bg[3] = {0, 0, 0};
  float Pm[16] = {0};
  make_view_consts(vc, V, Pm, 1.f, 1.f, bg, 1.f, 16, 16);
  const int K = (deg + 1) * (deg + 1);
  for (int g = 0; g < P; ++g) {
    const float* m = means + 3 * g;
    float dx = m[0] - vc.campos[0], dy = m[1] - vc.campos[1], dz = m[2] - vc.campos[2];
    float inv = 1.0f / sqrtf((dx * dx + dy * dy) + dz * dz);
    float pre[3];
    sh_eval_fused(deg, dx * inv, dy * inv, dz * inv, shs + (size_t)g * K * 3, 3, 1, pre);
    for (int c = 0; c < 3; ++c) {
      float acc = pre[c] + 0.5f;
      rgb[3 * g + c] = acc < 0.f ? 0.f : acc;
    }
  }
}

// Backward: g2d [P,10] = dpx,dpy,dconx,dcony,dconz,dopacity,drgb[3],ddepth.
// Outputs: dmeans[P,3] dscales[P,3] dquats[P,4] dshs[P,K,3] dV[16] (row-major, accumulated over P).
void hm_project_backward(int P, int deg, int use_sh, int cov_grad, int sh_grad,
                         const float* means, const float* scales, const float* quats,
                         const float* shs, const float* V, const float* Pm, float tanx, float tany,
                         float mod, int W, int H, const float* g2d, float* dmeans, float* dscales,
                         float* dquats, float* dshs, float* dV) {
  ViewConsts vc;
  float bg[3] = {0, 0, 0};
  make_view_consts(vc, V, Pm, tanx, tany, bg, mod, W, H);
  const int K = (deg + 1) * (deg + 1);
  double accA[9] = {0}, acct[3] = {0}, accc[3] = {0};
  for (int g = 0; g < P; ++g) {
    const float* m = means + 3 * g;
    Grad2D g2;
    const float* gg = g2d + 10 * g;
    g2.dpx = gg[0]; g2.dpy = gg[1]; g2.dconx = gg[2]; g2.dcony = gg[3]; g2.dconz = gg[4];
    g2.dopacity = gg[5]; g2.drgb[0] = gg[6]; g2.drgb[1] = gg[7]; g2.drgb[2] = gg[8]; g2.ddepth = gg[9];
    float gdir[3] = {0, 0, 0};
    if (use_sh) {
      float dx = m[0] - vc.campos[0], dy = m[1] - vc.campos[1], dz = m[2] - vc.campos[2];
      float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
      float x = dx * inv, y = dy * inv, z = dz * inv;
      float B[25], v[25];
      sh_basis(deg, x, y, z, B);
      float gm[3];
      for (int c = 0; c < 3; ++c) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc += B[k] * shs[(size_t)g * K * 3 + k * 3 + c];
        gm[c] = (acc + 0.5f) < 0.f ? 0.f : g2.drgb[c];
      }
      // the fused single-pass SH backward the CUDA kernel uses (dsh written, dL/ddir accumulated)
      float gx, gy, gz;
      sh_backward_fused<false>(deg, x, y, z, shs + (size_t)g * K * 3, dshs + (size_t)g * K * 3, 3, 1, gm, gx, gy, gz);
      if (sh_grad) { gdir[0] = gx; gdir[1] = gy; gdir[2] = gz; }
      (void)v;
    }
    Grad3D o = {};
    project_backward(vc, m, scales + 3 * g, quats + 4 * g, g2, gdir, cov_grad != 0, o);
    for (int i = 0; i < 3; ++i) { dmeans[3 * g + i] = o.dm[i]; dscales[3 * g + i] = o.ds[i]; }
    for (int i = 0; i < 4; ++i) dquats[4 * g + i] = o.dq[i];
    for (int i = 0; i < 9; ++i) accA[i] += o.dA[i];
    for (int i = 0; i < 3; ++i) { acct[i] += o.dtau[i]; accc[i] += o.dcam[i]; }
  }
  float dA[9], dtau[3], dcam[3];
  for (int i = 0; i < 9; ++i) dA[i] = (float)accA[i];
  for (int i = 0; i < 3; ++i) { dtau[i] = (float)acct[i]; dcam[i] = (float)accc[i]; }
  fold_campos_grad(V, dcam, dA, dtau);
  for (int i = 0; i < 16; ++i) dV[i] = 0.f;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) dV[4 * i + j] = dA[3 * i + j];
  for (int j = 0; j < 3; ++j) dV[12 + j] = dtau[j];
}

}  // extern "C"
