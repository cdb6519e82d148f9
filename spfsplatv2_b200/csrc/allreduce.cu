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
// Cross-rank ordering.  Every rank's bucket must be final BEFORE the reduction and every slice broadcast AFTER it.
//   * spf_multimem_allreduce_f32: the caller brackets the launch with two symmetric-memory barriers on the same stream.
//   * spf_multimem_allreduce_f32_fused: both barriers are INSIDE the kernel, over the symmetric-memory signal pads
//     (one 32-bit flag per (channel, peer) in every rank's pad): CTA b of rank r raises flag (b, r) in every peer's pad
//     (release, system scope) and waits for the flags (b, peer) in its own pad (acquire), before its first
//     multimem.ld_reduce and again after its last multimem.st.  CTA b of every rank thus rendezvous with CTA b of every
//     other rank; a peer's CTA can only be running once the kernels before it on that peer's stream (the producers of
//     its bucket) have completed, and the kernel as a whole only completes once all its CTAs have seen all their peers
//     finish -- which is what the two separate barrier launches guaranteed, without the two launches.
// fp32 sums of `world` addends in switch order: results are identical on all ranks (each element is reduced once, then
// broadcast).
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

// ---- in-kernel cross-rank barrier over the signal pads ------------------------------------------------------------
__device__ __forceinline__ void signal_raise(uint32_t* flag) {     // release: my earlier writes are visible before the flag
  __threadfence_system();
  while (atomicCAS_system(flag, 0u, 1u) != 0u) {
  }
}
__device__ __forceinline__ void signal_take(uint32_t* flag) {      // acquire: the raiser's writes are visible after
  while (atomicCAS_system(flag, 1u, 0u) != 1u) {
  }
  __threadfence_system();
}
// flags of one barrier: word  base + (phase * gridDim.x + block) * world + source rank  of the DESTINATION rank's pad
__device__ __forceinline__ void block_barrier_across_ranks(uint32_t* const* pads, int base, int phase, int rank, int world) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int ch = (phase * (int)gridDim.x + (int)blockIdx.x) * world;
    signal_raise(pads[threadIdx.x] + base + ch + rank);
    signal_take(pads[rank] + base + ch + (int)threadIdx.x);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(AR_THREADS)
multimem_allreduce_f32_fused_kernel(float4* __restrict__ mc, int64_t begin4, int64_t end4, uint32_t* const* __restrict__ pads,
                                    int base, int rank, int world) {
  block_barrier_across_ranks(pads, base, 0, rank, world);          // every rank's bucket is final
  const int64_t stride = (int64_t)gridDim.x * AR_THREADS;
  int64_t i = begin4 + (int64_t)blockIdx.x * AR_THREADS + threadIdx.x;
  for (; i + (AR_UNROLL - 1) * stride < end4; i += AR_UNROLL * stride) {
    float4 v[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) v[u] = multimem_ld_reduce_add(mc + i + u * stride);
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) multimem_st(mc + i + u * stride, v[u]);
  }
  for (; i < end4; i += stride) multimem_st(mc + i, multimem_ld_reduce_add(mc + i));
  __threadfence_system();                                          // this thread's broadcast stores, before the flags
  block_barrier_across_ranks(pads, base, 1, rank, world);          // every slice has been broadcast to every rank
}

cudaError_t launch_multimem_allreduce_f32_fused(float* mc, int64_t numel, int rank, int world, int n_blocks,
                                                uint32_t* const* signal_pads, int pad_word_offset, cudaStream_t stream) {
  const int64_t total4 = numel >> 2;
  const int64_t begin4 = total4 * rank / world, end4 = total4 * (rank + 1) / world;
  // the grid must be the SAME on every rank (CTA b meets CTA b): size it from the largest slice
  const int64_t max_slice = (total4 + world - 1) / world;
  int64_t want = (max_slice + AR_THREADS - 1) / AR_THREADS;
  if (want < 1) want = 1;
  const int grid = (int)(want < n_blocks ? want : n_blocks);
  multimem_allreduce_f32_fused_kernel<<<grid, AR_THREADS, 0, stream>>>(reinterpret_cast<float4*>(mc), begin4, end4, signal_pads,
                                                                       pad_word_offset, rank, world);
  return cudaGetLastError();
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
