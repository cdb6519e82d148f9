// K2-K5: duplicate-with-keys tile binning and on-device sort.
//
//   scan      : exclusive prefix sums of (a) per-block duplicate totals -> duplicate slots and the
//               total N, (b) per-tile duplicate counts -> tile bucket starts (tile ranges).
//   emit      : every (view, Gaussian) writes one 64-bit entry (depth_bits << 32 | gaussian) per
//               touched tile into that tile's bucket (MSD step of the radix sort: the tile digit is
//               resolved by counting, so the global 64-bit key (tile << 32 | depth_bits) never has to
//               be sorted as a whole).
//   sort+pack : one CTA per (view, tile) sorts its bucket by (depth_bits, gaussian) in shared memory
//               -- identical to a stable sort of (tile<<32|depth_bits) with ties in emission
//               (Gaussian-id) order, because a Gaussian appears at most once per tile -- and gathers
//               the projected attributes into a packed 48-byte slab record per duplicate, which is
//               what the blend kernels stream with TMA.
//
// Replaces the InclusiveSum / duplicateWithKeys / SortPairs / identifyTileRanges stages of
// diff_gauss_pose (SURVEY.md App. B "Binning"); bit-exact against oracle/raster_oracle.py:bin_and_sort.
#include "spf_device.cuh"
#include "spf_kernels.h"
#include "spf_math.h"

namespace spf {

// ---- K2: two independent exclusive scans in one launch (block 0 / block 1) ---------------------
__device__ void block_exclusive_scan(const int* __restrict__ in, int* __restrict__ out, int64_t n,
                                     int* total_out, volatile int* host_counters, int ticket) {
  __shared__ int warp_part[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += blockDim.x) {
    const int64_t i = base + tid;
    const int v = (i < n) ? in[i] : 0;
    int inc = warp_incl_scan_i(v, lane);
    if (lane == 31) warp_part[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int p = (lane < (blockDim.x >> 5)) ? warp_part[lane] : 0;
      int pi = warp_incl_scan_i(p, lane);
      warp_part[lane] = pi - p;  // exclusive warp offsets
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + warp_part[wid] + inc - v;
    if (i < n) out[i] = excl;
    __syncthreads();
    if (tid == blockDim.x - 1) carry_s = excl + v;
    __syncthreads();
  }
  if (tid == 0) {
    out[n] = carry_s;
    if (total_out) *total_out = carry_s;
    if (host_counters) {        // mapped pinned host memory: N, fence, ticket
      host_counters[0] = carry_s;
      __threadfence_system();
      host_counters[1] = ticket;
    }
  }
}

__global__ void __launch_bounds__(1024)
scan_kernel(const int* block_sum, int* block_off, int64_t n_blocks, const int* tile_count,
            int* tile_start, int64_t n_tiles, int* n_total, int* host_counters, int ticket) {
  if (blockIdx.x == 0) block_exclusive_scan(block_sum, block_off, n_blocks, n_total, host_counters, ticket);
  else block_exclusive_scan(tile_count, tile_start, n_tiles, nullptr, nullptr, 0);
}

cudaError_t launch_scan(const Dims& d, const SpfRasterState& st, const ControlLayout& cl, cudaStream_t s) {
  int* c = st.control;
  scan_kernel<<<2, 1024, 0, s>>>(c + cl.block_sum, c + cl.block_off, (int64_t)d.B * d.NB, c + cl.tile_count,
                                 c + cl.tile_start, (int64_t)d.B * d.T, c + cl.n_total, st.host_counters,
                                 d.ticket);
  return cudaGetLastError();
}

// Tile rectangle of a projected Gaussian, recomputed from (xy, radius).  Pure add / exact scaling
// by 1/16: identical bits to project_forward in any translation unit.
__device__ __forceinline__ void rect_of(float px, float py, int radius, int gx, int gy, int& rx0, int& ry0,
                                        int& rx1, int& ry1) {
  const float r = (float)radius;
  rx0 = tile_clamp((px - r) / (float)TILE, gx);
  ry0 = tile_clamp((py - r) / (float)TILE, gy);
  rx1 = tile_clamp(((px + r) + (float)(TILE - 1)) / (float)TILE, gx);
  ry1 = tile_clamp(((py + r) + (float)(TILE - 1)) / (float)TILE, gy);
}

// ---- K4: emit ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(PROJ_THREADS)
emit_kernel(Dims d, SpfRasterState st, const int* __restrict__ block_off, const int* __restrict__ tile_start,
            int* __restrict__ tile_cursor, int* __restrict__ overflow) {
  __shared__ int warp_part[PROJ_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int view = blockIdx.y;
  const int g = blockIdx.x * PROJ_THREADS + tid;
  const size_t vg = (size_t)view * d.P + g;
  const int tiles = (g < d.P) ? st.tiles_touched[vg] : 0;
  const int inc = warp_incl_scan_i(tiles, lane);
  if (lane == 31) warp_part[wid] = inc;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < wid; ++w) woff += warp_part[w];
  const int slot0 = block_off[(size_t)view * d.NB + blockIdx.x] + woff + inc - tiles;
  if (g >= d.P) return;
  st.dup_offset[vg] = slot0;
  if (tiles == 0) return;
  const float2 p = reinterpret_cast<const float2*>(st.xy)[vg];
  int rx0, ry0, rx1, ry1;
  rect_of(p.x, p.y, st.radii[vg], d.gx, d.gy, rx0, ry0, rx1, ry1);
  const uint64_t entry = ((uint64_t)__float_as_uint(st.depth[vg]) << 32) | (uint32_t)g;
  const size_t tbase = (size_t)view * d.T;
  for (int y = ry0; y < ry1; ++y)
    for (int x = rx0; x < rx1; ++x) {
      const size_t t = tbase + y * d.gx + x;
      const int64_t pos = (int64_t)tile_start[t] + atomicAdd(tile_cursor + t, 1);
      if (pos < d.cap) st.bucket[pos] = entry;
      else *overflow = 1;
    }
}

cudaError_t launch_emit(const Dims& d, const SpfRasterState& st, const ControlLayout& cl, cudaStream_t s) {
  int* c = st.control;
  dim3 grid(d.NB, d.B);
  emit_kernel<<<grid, PROJ_THREADS, 0, s>>>(d, st, c + cl.block_off, c + cl.tile_start, c + cl.tile_cursor,
                                            c + cl.overflow);
  return cudaGetLastError();
}

// Conservative pixel-space bounding box of {alpha >= 1/255}:  o*exp(-q/2) >= 1/255  <=>
// q = cx dx^2 + 2 cy dx dy + cz dy^2 <= 2 ln(255 o) =: tau.  Half extents of that ellipse are
// sqrt(tau*cz/det), sqrt(tau*cx/det) with det = cx*cz - cy^2.  Inflated (1e-4 relative + 0.01 px) so
// rounding can never reject a pixel the exact per-pixel test would accept; a Gaussian that can never
// reach 1/255 gets an empty box, a degenerate conic an infinite one.
__device__ __forceinline__ float4 alpha_bbox(float px, float py, float cx, float cy, float cz, float o) {
  const float inf = __int_as_float(0x7f800000);
  const float tau = 2.0f * logf(255.0f * o);
  if (!(tau >= 0.0f)) {
    if (tau < 0.0f) return make_float4(inf, -inf, inf, -inf);   // never visible
    return make_float4(-inf, inf, -inf, inf);                   // NaN: do not cull
  }
  const float det = cx * cz - cy * cy;
  if (!(det > 0.0f) || !(cx > 0.0f) || !(cz > 0.0f)) return make_float4(-inf, inf, -inf, inf);
  const float ex = sqrtf(tau * cz / det) * 1.0001f + 0.01f;
  const float ey = sqrtf(tau * cx / det) * 1.0001f + 0.01f;
  if (!(ex < inf) || !(ey < inf)) return make_float4(-inf, inf, -inf, inf);
  return make_float4(px - ex, px + ex, py - ey, py + ey);
}

// ---- K5: per-tile sort + slab pack ---------------------------------------------------------------
// Bitonic network with ascending-only comparators ("flip" then "shift" stages); comparators whose
// partner index is >= n are skipped, which equals padding with +inf, so any n works.
template <typename KeyPtr>
__device__ __forceinline__ void bitonic_sort(KeyPtr keys, int n, int tid, int nthreads) {
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  const int half = n2 >> 1;
  for (int k = 2; k <= n2; k <<= 1) {
    // flip stage
    {
      const int hk = k >> 1;
      for (int p = tid; p < half; p += nthreads) {
        const int blk = p / hk, off = p - blk * hk;
        const int i = blk * k + off;
        const int j = blk * k + k - 1 - off;
        if (j < n) {
          const uint64_t a = keys[i], b = keys[j];
          if (a > b) { keys[i] = b; keys[j] = a; }
        }
      }
      __syncthreads();
    }
    for (int jj = k >> 2; jj > 0; jj >>= 1) {
      for (int p = tid; p < half; p += nthreads) {
        const int blk = p / jj, off = p - blk * jj;
        const int i = blk * 2 * jj + off;
        const int j = i + jj;
        if (j < n) {
          const uint64_t a = keys[i], b = keys[j];
          if (a > b) { keys[i] = b; keys[j] = a; }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(TILE_THREADS)
tile_sort_pack_kernel(Dims d, SpfRasterState st, const int* __restrict__ tile_start) {
  __shared__ uint64_t skeys[SORT_SMEM_CAP];
  const int tid = threadIdx.x;
  const int t = blockIdx.x;              // view * T + tile
  const int view = t / d.T;
  const int tile = t - view * d.T;
  int64_t s64 = tile_start[t], e64 = tile_start[t + 1];
  if (s64 > d.cap) s64 = d.cap;
  if (e64 > d.cap) e64 = d.cap;
  const int n = (int)(e64 - s64);
  if (tid == 0) {
    st.tile_ranges[2 * (size_t)t] = n > 0 ? (int)s64 : 0;
    st.tile_ranges[2 * (size_t)t + 1] = n > 0 ? (int)e64 : 0;
  }
  if (n == 0) return;
  uint64_t* gk = st.bucket + s64;
  const uint64_t* sorted;
  if (n <= SORT_SMEM_CAP) {
    for (int i = tid; i < n; i += TILE_THREADS) skeys[i] = gk[i];
    __syncthreads();
    bitonic_sort(skeys, n, tid, TILE_THREADS);
    sorted = skeys;
  } else {
    __syncthreads();
    bitonic_sort(gk, n, tid, TILE_THREADS);   // rare: very long lists are sorted in place in HBM/L2
    sorted = gk;
  }
  const int tx = tile % d.gx, ty = tile / d.gx;
  float4* slab = reinterpret_cast<float4*>(st.slab) + 3 * s64;
  float4* cull = reinterpret_cast<float4*>(st.cullbox) + s64;
  for (int i = tid; i < n; i += TILE_THREADS) {
    const uint64_t key = sorted[i];
    const int g = (int)(uint32_t)(key & 0xffffffffu);
    const size_t vg = (size_t)view * d.P + g;
    const float2 p = reinterpret_cast<const float2*>(st.xy)[vg];
    const float4 co = reinterpret_cast<const float4*>(st.conic_opacity)[vg];
    const float r = st.rgb[vg * 3 + 0], gg = st.rgb[vg * 3 + 1], b = st.rgb[vg * 3 + 2];
    const float dep = __uint_as_float((uint32_t)(key >> 32));
    int rx0, ry0, rx1, ry1;
    rect_of(p.x, p.y, st.radii[vg], d.gx, d.gy, rx0, ry0, rx1, ry1);
    const int slot = st.dup_offset[vg] + (ty - ry0) * (rx1 - rx0) + (tx - rx0);
    slab[3 * i + 0] = make_float4(p.x, p.y, co.x, co.y);
    slab[3 * i + 1] = make_float4(co.z, co.w, r, gg);
    slab[3 * i + 2] = make_float4(b, dep, __int_as_float(slot), __int_as_float(g));
    cull[i] = alpha_bbox(p.x, p.y, co.x, co.y, co.z, co.w);
  }
}

cudaError_t launch_tile_sort_pack(const Dims& d, const SpfRasterState& st, const ControlLayout& cl,
                                  cudaStream_t s) {
  tile_sort_pack_kernel<<<d.B * d.T, TILE_THREADS, 0, s>>>(d, st, st.control + cl.tile_start);
  return cudaGetLastError();
}

// ---- parity helper: (point_list, keys) out of the slab -------------------------------------------
__global__ void unpack_sorted_kernel(Dims d, SpfRasterState st, const int* __restrict__ tile_start,
                                     int64_t n, int32_t* point_list, uint64_t* keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 c = reinterpret_cast<const float4*>(st.slab)[3 * i + 2];
  if (point_list) point_list[i] = __float_as_int(c.w);
  if (keys) {
    // find the tile whose range holds i (binary search over tile_start)
    int lo = 0, hi = d.B * d.T;   // tile_start has B*T+1 entries
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (tile_start[mid] <= i) lo = mid; else hi = mid;
    }
    const int tile = lo % d.T;
    keys[i] = ((uint64_t)tile << 32) | (uint64_t)__float_as_uint(c.y);
  }
}

cudaError_t launch_unpack_sorted(const Dims& d, const SpfRasterState& st, int64_t n, int32_t* point_list,
                                 uint64_t* keys, const ControlLayout& cl, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  unpack_sorted_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, s>>>(
      d, st, st.control + cl.tile_start, n, point_list, keys);
  return cudaGetLastError();
}

}  // namespace spf
