// In-switch sum all-reduce of one fp32 gradient bucket over NVLink 5 / NVSwitch multicast (NVLS), SURVEY.md §8e.
//
// What it replaces: the NCCL all-reduce torch DDP issues for the replicated parameters' gradients
// (/root/reference/src/main.py:141-145, strategy "ddp_find_unused_parameters_true").  The bucket lives in symmetric memory
// (same virtual offset on every rank, bound to one multicast object); `mc` is the MULTICAST address of the bucket.
// Rank r owns the r-th 1/world slice: `multimem.ld_reduce` pulls the slice from all ranks and sums it inside the switch
// (one NVLink read of 1/world of the bucket per GPU instead of world-1 ring hops), `multimem.st` broadcasts the sum back
// to every rank.  HBM traffic per GPU: bucket/world read-side + bucket written once; SM cost: `n_blocks` CTAs, chosen
// small so the renderer kernels running beside it on the main stream keep their SMs.
//
// Cross-rank ordering is the caller's: a symmetric-memory barrier on the same stream BEFORE (every rank's bucket is
// final) and AFTER (every slice has been broadcast) the launch -- spfsplatv2_b200/dp.py.  fp32 sums of `world` addends
// in switch order: results are identical on all ranks (each element is reduced once, then broadcast).
#include "spf_device.cuh"
#include "spf_kernels.h"

namespace spf {

constexpr int AR_THREADS = 1024;
constexpr int AR_UNROLL = 4;

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}

__device__ __forceinline__ void multimem_st(float4* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(AR_THREADS)
multimem_allreduce_f32_kernel(float4* __restrict__ mc, int64_t begin4, int64_t end4) {
  const int64_t stride = (int64_t)gridDim.x * AR_THREADS;
  int64_t i = begin4 + (int64_t)blockIdx.x * AR_THREADS + threadIdx.x;
  // AR_UNROLL independent in-switch reductions in flight per thread before the first broadcast store
  for (; i + (AR_UNROLL - 1) * stride < end4; i += AR_UNROLL * stride) {
    float4 v[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) v[u] = multimem_ld_reduce_add(mc + i + u * stride);
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) multimem_st(mc + i + u * stride, v[u]);
  }
  for (; i < end4; i += stride) multimem_st(mc + i, multimem_ld_reduce_add(mc + i));
}

cudaError_t launch_multimem_allreduce_f32(float* mc, int64_t numel, int rank, int world, int n_blocks,
                                          cudaStream_t stream) {
  const int64_t total4 = numel >> 2;
  const int64_t begin4 = total4 * rank / world, end4 = total4 * (rank + 1) / world;
  if (end4 <= begin4) return cudaSuccess;
  const int64_t want = (end4 - begin4 + AR_THREADS - 1) / AR_THREADS;
  const int grid = (int)(want < n_blocks ? want : n_blocks);
  multimem_allreduce_f32_kernel<<<grid, AR_THREADS, 0, stream>>>(reinterpret_cast<float4*>(mc), begin4, end4);
  return cudaGetLastError();
}

}  // namespace spf
