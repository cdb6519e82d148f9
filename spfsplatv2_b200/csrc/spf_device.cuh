// Device-side helpers: mbarrier + 1-D TMA bulk copies (cp.async.bulk, sm_90+/sm_100a),
// warp/block reductions, cooperative copies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spf {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- programmatic dependent launch ---------------------------------------------------------
// First statement of every raster-path kernel (see pdl_launch in spf_kernels.h): wait until the preceding kernel in the
// stream has completed and its writes are visible, then allow the following kernel's CTAs to be scheduled into free SM
// slots.  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  // make the init visible to the async (TMA) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA 1-D bulk copy global -> shared (bytes % 16 == 0, both addresses 16-B aligned) --------
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global bulk store (bulk async-group completion)
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// order generic-proxy smem writes before async-proxy reads (needed before a bulk store)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// ---- warp-aggregated atomics ------------------------------------------------------------------
// Every lane of the warp calls this (has = false for lanes with nothing to add).  Lanes that target the same counter
// are grouped with match.any; only the group leader issues the atomic.  Neighbouring Gaussians (one per context pixel,
// raster order) land in the same one or two tiles, so this removes ~10x of the same-address L2 atomics.
// Returns the value of the counter before the group's add, plus this lane's rank inside its group.
__device__ __forceinline__ int warp_aggregated_add(bool has, int* counters, int key, int lane) {
  const unsigned act = __ballot_sync(0xffffffffu, has);
  int pos = 0;
  if (has) {
    const unsigned peers = __match_any_sync(act, key);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counters + key, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    pos = base + __popc(peers & ((1u << lane) - 1u));
  }
  return pos;
}

// ---- reductions ------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// Cooperative copy of n floats global -> shared with optional row padding:
// dst[(i / row) * stride + i % row] = src[i].  Uses 128-bit loads when possible.
__device__ __forceinline__ void block_copy_g2s(float* dst, const float* __restrict__ src, int n, int row,
                                               int stride, int tid, int nthreads) {
  if (row == stride && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    const int n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = tid; i < n4; i += nthreads) d4[i] = __ldg(s4 + i);
    for (int i = (n4 << 2) + tid; i < n; i += nthreads) dst[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < n; i += nthreads) {
      const int r = i / row;
      dst[r * stride + (i - r * row)] = __ldg(src + i);
    }
  }
}
__device__ __forceinline__ void block_copy_s2g(float* __restrict__ dst, const float* src, int n, int row,
                                               int stride, int tid, int nthreads) {
  if (row == stride && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    const int n4 = n >> 2;
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int i = tid; i < n4; i += nthreads) d4[i] = s4[i];
    for (int i = (n4 << 2) + tid; i < n; i += nthreads) dst[i] = src[i];
  } else {
    for (int i = tid; i < n; i += nthreads) {
      const int r = i / row;
      dst[i] = src[r * stride + (i - r * row)];
    }
  }
}

}  // namespace spf
