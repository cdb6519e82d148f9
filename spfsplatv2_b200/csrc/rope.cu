// 2-D rotary position embedding, in place, for sm_100a.
//
// Replaces rope_2d_cuda / rope_2d_cuda_kernel of the reference's curope extension
// (/root/reference/src/model/encoder/backbone/croco/curope/kernels.cu:17-108) with the arithmetic of
// its CPU path rope_2d_cpu (curope.cpp:11-47): angle = fwd * pos / powf(base, i / Q), Q = D/4,
// token = [u_Y(Q) v_Y(Q) u_X(Q) v_X(Q)],  u' = u cos - v sin,  v' = v cos + u sin.
//
// Layout: one thread per (token, half, VEC consecutive frequencies); it evaluates powf/sincosf once
// and sweeps all H heads with 128-bit (fp32) / 64-bit (fp16, bf16) loads and stores -- no shared
// memory, no per-head barrier, bf16 supported.  HBM-bound: 2 * B*N*H*D * sizeof(T) bytes.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "spf_kernels.h"

namespace spf {

template <typename T> struct Cvt;
template <> struct Cvt<float> {
  static __device__ __forceinline__ float ld(float v) { return v; }
  static __device__ __forceinline__ float st(float v) { return v; }
};
template <> struct Cvt<__half> {
  static __device__ __forceinline__ float ld(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half st(float v) { return __float2half_rn(v); }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 st(float v) { return __float2bfloat16_rn(v); }
};

template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
rope2d_kernel(T* __restrict__ tokens, const int64_t* __restrict__ pos, int64_t n_tokens, int N, int H, int D,
              int64_t sb, int64_t sn, float base, float fwd) {
  const int Q = D >> 2;
  const int per_half = Q / VEC;
  const int per_token = 2 * per_half;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_tokens * per_token) return;
  const int64_t tok = gid / per_token;
  const int r = (int)(gid - tok * per_token);
  const int X = r / per_half;            // 0: y half, 1: x half
  const int i0 = (r - X * per_half) * VEC;
  const int64_t b = tok / N, n = tok - b * N;
  const int p = (int)pos[tok * 2 + X];
  float cs[VEC], sn_[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    const float ang = fwd * p / powf(base, (float)(i0 + k) / (float)Q);
    sincosf(ang, &sn_[k], &cs[k]);
  }
  T* row = tokens + b * sb + n * sn + X * 2 * Q + i0;
  typedef Pack<T, VEC> PK;
  for (int h = 0; h < H; ++h, row += D) {
    PK u = *reinterpret_cast<const PK*>(row);
    PK v = *reinterpret_cast<const PK*>(row + Q);
    PK uo, vo;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float uf = Cvt<T>::ld(u.v[k]), vf = Cvt<T>::ld(v.v[k]);
      uo.v[k] = Cvt<T>::st(uf * cs[k] - vf * sn_[k]);
      vo.v[k] = Cvt<T>::st(vf * cs[k] + uf * sn_[k]);
    }
    *reinterpret_cast<PK*>(row) = uo;
    *reinterpret_cast<PK*>(row + Q) = vo;
  }
}

template <typename T>
static cudaError_t launch_t(void* tokens, const int64_t* pos, int B, int N, int H, int D, int64_t sb,
                            int64_t sn, float base, float fwd, cudaStream_t s) {
  const int Q = D / 4;
  const int64_t n_tokens = (int64_t)B * N;
  const size_t vb = sizeof(T) * 4;
  const bool vec4 = (Q % 4 == 0) && ((reinterpret_cast<uintptr_t>(tokens) % vb) == 0) &&
                    ((sb * sizeof(T)) % vb == 0) && ((sn * sizeof(T)) % vb == 0);
  if (n_tokens == 0 || H == 0) return cudaSuccess;
  if (vec4) {
    const int64_t threads = n_tokens * 2 * (Q / 4);
    rope2d_kernel<T, 4><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>((T*)tokens, pos, n_tokens, N, H, D, sb,
                                                                         sn, base, fwd);
  } else {
    const int64_t threads = n_tokens * 2 * Q;
    rope2d_kernel<T, 1><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>((T*)tokens, pos, n_tokens, N, H, D, sb,
                                                                         sn, base, fwd);
  }
  return cudaGetLastError();
}

cudaError_t launch_rope2d(void* tokens, const int64_t* pos, int B, int N, int H, int D, int64_t sb,
                          int64_t sn, int dtype, float base, float fwd, cudaStream_t s) {
  switch (dtype) {
    case 0: return launch_t<float>(tokens, pos, B, N, H, D, sb, sn, base, fwd, s);
    case 1: return launch_t<__half>(tokens, pos, B, N, H, D, sb, sn, base, fwd, s);
    case 2: return launch_t<__nv_bfloat16>(tokens, pos, B, N, H, D, sb, sn, base, fwd, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace spf
