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
//
// q and k of a self-attention block share their positions (blocks.py:97-104: both are views of one fused
// qkv tensor): spf_rope2d_qk rotates BOTH in one launch, the thread reusing its cos/sin for the second
// tensor -- half the launches and half the transcendental work of two rope_2d calls.
// fp64 (dispatched by the reference, kernels.cu:101) follows the reference's arithmetic: the value is
// rounded to fp32 (its shared-memory staging buffer is float, kernels.cu:30,66), rotated in fp32 and widened.
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

template <> struct Cvt<double> {
  static __device__ __forceinline__ float ld(double v) { return (float)v; }
  static __device__ __forceinline__ double st(float v) { return (double)v; }
};

template <typename T, int VEC> struct alignas(sizeof(T) * VEC <= 16 ? sizeof(T) * VEC : 16) Pack { T v[VEC]; };

// blockIdx.y selects a chunk of `hpt` heads: small problems (the decoder's 16 x 258 tokens) are split over the heads as
// well so that enough loads are in flight to cover the HBM latency; the per-thread cos/sin work is repeated per chunk
// (a few hundred instructions against kilobytes of traffic).  The loads of a head (u, v) for BOTH tensors are issued
// before anything is computed.  (Measured and dropped: issuing the first head's loads before the powf / sincosf chain
// plus a register prefetch of the next head -- 13.5 -> 17.8 us on the decoder shape, the extra live registers cost
// more occupancy than the overlap bought.)
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
rope2d_kernel(T* __restrict__ tokens, T* __restrict__ tokens2, const int64_t* __restrict__ pos, int64_t n_tokens, int N,
              int H, int D, int hpt, int64_t sb, int64_t sn, float base, float fwd) {
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
  typedef Pack<T, VEC> PK;
  const int h0 = blockIdx.y * hpt, h1 = min(H, h0 + hpt);
  const int64_t off = b * sb + n * sn + (int64_t)h0 * D + X * 2 * Q + i0;
  T* row = tokens + off;
  T* row2 = tokens2 ? tokens2 + off : nullptr;
  auto rot = [&](const PK& u, const PK& v, T* dst) {
    PK uo, vo;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float uf = Cvt<T>::ld(u.v[k]), vf = Cvt<T>::ld(v.v[k]);
      uo.v[k] = Cvt<T>::st(uf * cs[k] - vf * sn_[k]);
      vo.v[k] = Cvt<T>::st(vf * cs[k] + uf * sn_[k]);
    }
    *reinterpret_cast<PK*>(dst) = uo;
    *reinterpret_cast<PK*>(dst + Q) = vo;
  };
  if (row2) {
#pragma unroll 2
    for (int h = h0; h < h1; ++h, row += D, row2 += D) {
      const PK u = *reinterpret_cast<const PK*>(row), v = *reinterpret_cast<const PK*>(row + Q);
      const PK u2 = *reinterpret_cast<const PK*>(row2), v2 = *reinterpret_cast<const PK*>(row2 + Q);
      rot(u, v, row);
      rot(u2, v2, row2);
    }
  } else {
#pragma unroll 4
    for (int h = h0; h < h1; ++h, row += D) {
      const PK u = *reinterpret_cast<const PK*>(row), v = *reinterpret_cast<const PK*>(row + Q);
      rot(u, v, row);
    }
  }
}

template <typename T>
static cudaError_t launch_t(void* tokens, void* tokens2, const int64_t* pos, int B, int N, int H, int D, int64_t sb,
                            int64_t sn, float base, float fwd, cudaStream_t s) {
  const int Q = D / 4;
  const int64_t n_tokens = (int64_t)B * N;
  constexpr int VW = sizeof(T) >= 4 ? 4 : 8;                  // elements per 128-bit access (fp64: two accesses)
  const size_t vb = sizeof(T) * VW <= 16 ? sizeof(T) * VW : 16;
  auto aligned = [&](int vec) {
    const size_t bytes = sizeof(T) * vec <= 16 ? sizeof(T) * vec : 16;
    return (Q % vec == 0) && ((reinterpret_cast<uintptr_t>(tokens) % bytes) == 0) &&
           ((reinterpret_cast<uintptr_t>(tokens2) % bytes) == 0) && ((sb * sizeof(T)) % bytes == 0) &&
           ((sn * sizeof(T)) % bytes == 0) && ((Q * sizeof(T)) % bytes == 0) && ((D * sizeof(T)) % bytes == 0);
  };
  (void)vb;
  const int vec = aligned(VW) ? VW : (aligned(4) ? 4 : 1);
  if (n_tokens == 0 || H == 0) return cudaSuccess;
  const int64_t threads = n_tokens * 2 * (Q / vec);
  // heads per thread: all of them when the (token, frequency) threads alone fill the machine several times over,
  // otherwise fewer, down to 2, until about 600 k threads (148 SMs x 2048 x 2) are in flight
  int hpt = H;     // (1 head per thread measured slower than 2: 14.4 vs 13.1 us on the decoder shape)
  while (hpt > 2 && threads * ((H + hpt - 1) / hpt) < 600000) hpt = (hpt + 1) / 2;
  dim3 grid((unsigned)((threads + 255) / 256), (unsigned)((H + hpt - 1) / hpt));
  if (vec == 8)
    rope2d_kernel<T, (sizeof(T) >= 4 ? 4 : 8)><<<grid, 256, 0, s>>>((T*)tokens, (T*)tokens2, pos, n_tokens, N, H, D, hpt, sb, sn, base, fwd);
  else if (vec == 4)
    rope2d_kernel<T, 4><<<grid, 256, 0, s>>>((T*)tokens, (T*)tokens2, pos, n_tokens, N, H, D, hpt, sb, sn, base, fwd);
  else
    rope2d_kernel<T, 1><<<grid, 256, 0, s>>>((T*)tokens, (T*)tokens2, pos, n_tokens, N, H, D, hpt, sb, sn, base, fwd);
  return cudaGetLastError();
}

cudaError_t launch_rope2d(void* tokens, void* tokens2, const int64_t* pos, int B, int N, int H, int D, int64_t sb,
                          int64_t sn, int dtype, float base, float fwd, cudaStream_t s) {
  switch (dtype) {
    case 0: return launch_t<float>(tokens, tokens2, pos, B, N, H, D, sb, sn, base, fwd, s);
    case 1: return launch_t<__half>(tokens, tokens2, pos, B, N, H, D, sb, sn, base, fwd, s);
    case 2: return launch_t<__nv_bfloat16>(tokens, tokens2, pos, B, N, H, D, sb, sn, base, fwd, s);
    case 3: return launch_t<double>(tokens, tokens2, pos, B, N, H, D, sb, sn, base, fwd, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace spf
