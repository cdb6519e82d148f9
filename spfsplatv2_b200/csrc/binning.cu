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
                                     int* total_out, volatile int* host_counters, int ticket, int64_t capacity = 0,
                                     int* overflow_out = nullptr) {
  __shared__ int warp_part[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  // 4 consecutive elements per thread per round: 4096 per round with 1024 threads (the whole job in 1-2 rounds at
  // the headline size instead of 8 barrier-separated ones)
  for (int64_t base = 0; base < n; base += 4 * (int64_t)blockDim.x) {
    const int64_t i0 = base + 4 * (int64_t)tid;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? in[i0 + k] : 0;
    const int tsum = (v[0] + v[1]) + (v[2] + v[3]);
    int inc = warp_incl_scan_i(tsum, lane);
    if (lane == 31) warp_part[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int p = (lane < (blockDim.x >> 5)) ? warp_part[lane] : 0;
      int pi = warp_incl_scan_i(p, lane);
      warp_part[lane] = pi - p;  // exclusive warp offsets
    }
    __syncthreads();
    const int carry = carry_s;
    int excl = carry + warp_part[wid] + inc - tsum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n) out[i0 + k] = excl;
      excl += v[k];
    }
    __syncthreads();
    if (tid == blockDim.x - 1) carry_s = excl;
    __syncthreads();
  }
  if (tid == 0) {
    out[n] = carry_s;
    if (total_out) {
      *total_out = carry_s;
      if ((int64_t)carry_s > capacity) *overflow_out = 1;     // read by the blend kernel, which poisons the image
    }
    if (host_counters) {        // mapped pinned host memory: N, fence, ticket
      host_counters[0] = carry_s;
      __threadfence_system();
      host_counters[1] = ticket;
    }
  }
}

__global__ void __launch_bounds__(1024)
scan_kernel(const int* block_sum, int* block_off, int64_t n_blocks, const int* tile_count,
            int* tile_start, int64_t n_tiles, int* n_total, int* host_counters, int ticket, int64_t capacity,
            int* overflow) {
  pdl_enter();
  if (blockIdx.x == 0) block_exclusive_scan(block_sum, block_off, n_blocks, n_total, host_counters, ticket, capacity, overflow);
  else block_exclusive_scan(tile_count, tile_start, n_tiles, nullptr, nullptr, 0);
}

cudaError_t launch_scan(const Dims& d, const SpfRasterState& st, const ControlLayout& cl, cudaStream_t s) {
  int* c = st.control;
  pdl_launch(scan_kernel, 2, 1024, 0, s)(c + cl.block_sum, c + cl.block_off, (int64_t)d.B * d.NB, c + cl.tile_count,
                                 c + cl.tile_start, (int64_t)d.B * d.T, c + cl.n_total, st.host_counters,
                                 d.ticket, d.cap, c + cl.overflow);
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
  pdl_enter();
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
  int rx0 = 0, ry0 = 0, rx1 = 1, ry1 = 0;
  uint64_t entry = 0;
  if (g < d.P) {
    st.dup_offset[vg] = slot0;
    if (tiles > 0) {
      const float2 p = reinterpret_cast<const float2*>(st.xy)[vg];
      rect_of(p.x, p.y, st.radii[vg], d.gx, d.gy, rx0, ry0, rx1, ry1);
      entry = ((uint64_t)__float_as_uint(st.depth[vg]) << 32) | (uint32_t)g;
    }
  }
  // one bucket slot per touched tile, claimed with warp-aggregated atomics on the tile cursors
  int* cursor = tile_cursor + (size_t)view * d.T;
  const int* tstart = tile_start + (size_t)view * d.T;
  int maxc = tiles;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
  const int cw = rx1 - rx0;
  int x = 0, y = 0;
  for (int k = 0; k < maxc; ++k) {
    const bool has = k < tiles;
    const int t = (ry0 + y) * d.gx + rx0 + x;
    const int rank = warp_aggregated_add(has, cursor, t, lane);
    if (has) {
      const int64_t pos = (int64_t)tstart[t] + rank;
      if (pos < d.cap) st.bucket[pos] = entry;
      else *overflow = 1;
    }
    if (++x == cw) { x = 0; ++y; }
  }
}

cudaError_t launch_emit(const Dims& d, const SpfRasterState& st, const ControlLayout& cl, cudaStream_t s) {
  int* c = st.control;
  dim3 grid(d.NB, d.B);
  pdl_launch(emit_kernel, grid, PROJ_THREADS, 0, s)(d, st, c + cl.block_off, c + cl.tile_start, c + cl.tile_cursor,
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

// Per-tile sort, fast path: a monotone BUCKET sort in shared memory.  Depth keys of one tile are spread over a
// range [mn, mx] of float bit patterns (monotone in depth for positive floats); bucket = floor((bits - mn) * NBK /
// (mx - mn + 1)) is monotone too, so a counting sort over the buckets orders everything except the few keys that
// share a bucket, and those are ranked exactly by comparing the full 64-bit (depth_bits, gaussian) keys inside
// their bucket.  O(n) work and 7 block barriers per tile instead of the ~log^2(n)/2 barrier-separated stages of
// a bitonic network.  Degenerate tiles (a bucket with more than MAX_BUCKET keys, e.g. many equal depths) and
// tiles longer than CAP fall back to the bitonic network.
constexpr int MAX_BUCKET = 48;

template <int CAP, int NBK>
struct SortSmem {
  uint64_t b[CAP];          // keys in bucket order
  int cur[NBK];             // histogram -> scatter cursors -> bucket ends
  unsigned red_min[8], red_max[8];
  int wtot[8];
  int bad;
};

__device__ __forceinline__ void pack_records(const Dims& d, const SpfRasterState& st, const uint64_t* sorted, int n,
                                             int64_t s64, int view, int tile, int tid) {
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

template <int CAP, int NBK>
__global__ void __launch_bounds__(TILE_THREADS)
tile_sort_pack_kernel(Dims d, SpfRasterState st, const int* __restrict__ tile_start) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char sort_smem_raw[];
  SortSmem<CAP, NBK>& S = *reinterpret_cast<SortSmem<CAP, NBK>*>(sort_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
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
  if (n > CAP) {                          // rare: very long list, sorted in place in HBM/L2
    bitonic_sort(gk, n, tid, TILE_THREADS);
    __syncthreads();
    pack_records(d, st, gk, n, s64, view, tile, tid);
    return;
  }
  // 1. the keys are read from global memory ONCE into registers (KPT per thread) and reused by the min/max pass,
  //    the histogram and the scatter: every extra pass would be another dependent trip to L2 per block
  constexpr int KPT = CAP / TILE_THREADS;
  uint64_t kreg[KPT];
  unsigned mn = 0xffffffffu, mx = 0u;
#pragma unroll
  for (int q = 0; q < KPT; ++q) {
    const int i = tid + q * TILE_THREADS;
    kreg[q] = (i < n) ? gk[i] : ~0ull;
    if (i < n) {
      const unsigned hi = (unsigned)(kreg[q] >> 32);
      mn = min(mn, hi); mx = max(mx, hi);
    }
  }
  for (int i = tid; i < NBK; i += TILE_THREADS) S.cur[i] = 0;
  if (tid == 0) S.bad = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) { S.red_min[wid] = mn; S.red_max[wid] = mx; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) { mn = min(mn, S.red_min[w]); mx = max(mx, S.red_max[w]); }
  const float scale = (float)NBK / ((float)(mx - mn) + 1.0f);
  auto bucket_of = [&](uint64_t k) -> int {
    const int b = (int)((float)((unsigned)(k >> 32) - mn) * scale);
    return min(b, NBK - 1);
  };
  // 2. histogram
#pragma unroll
  for (int q = 0; q < KPT; ++q)
    if (tid + q * TILE_THREADS < n) atomicAdd(&S.cur[bucket_of(kreg[q])], 1);
  __syncthreads();
  // 3. exclusive scan of the NBK counters, in place: cur[b] = first slot of bucket b
  {
    constexpr int PER = NBK / TILE_THREADS;
    int local[PER];
    int sum = 0, mxb = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { local[k] = S.cur[tid * PER + k]; sum += local[k]; mxb = max(mxb, local[k]); }
    if (mxb > MAX_BUCKET) S.bad = 1;
    const int inc = warp_incl_scan_i(sum, lane);
    if (lane == 31) S.wtot[wid] = inc;
    __syncthreads();
    int off = inc - sum;
    for (int w = 0; w < wid; ++w) off += S.wtot[w];
#pragma unroll
    for (int k = 0; k < PER; ++k) { S.cur[tid * PER + k] = off; off += local[k]; }
  }
  __syncthreads();
  if (S.bad) {                             // degenerate depth distribution: bitonic network
#pragma unroll
    for (int q = 0; q < KPT; ++q)
      if (tid + q * TILE_THREADS < n) S.b[tid + q * TILE_THREADS] = kreg[q];
    __syncthreads();
    bitonic_sort(S.b, n, tid, TILE_THREADS);
    pack_records(d, st, S.b, n, s64, view, tile, tid);
    return;
  }
  // 4. scatter into bucket order (arbitrary order inside a bucket); afterwards cur[b] = END of bucket b
#pragma unroll
  for (int q = 0; q < KPT; ++q)
    if (tid + q * TILE_THREADS < n) S.b[atomicAdd(&S.cur[bucket_of(kreg[q])], 1)] = kreg[q];
  __syncthreads();
  // 5. exact rank inside the bucket by full-key comparison (keys are unique: a Gaussian occurs once per tile);
  //    the sorted list goes back to the (fully consumed) global bucket, where the pack step streams it from
  for (int i = tid; i < n; i += TILE_THREADS) {
    const uint64_t k = S.b[i];
    const int bk = bucket_of(k);
    const int bs = bk ? S.cur[bk - 1] : 0, be = S.cur[bk];
    int r = 0;
    for (int q = bs; q < be; ++q) r += (S.b[q] < k) ? 1 : 0;
    gk[bs + r] = k;
  }
  __syncthreads();
  pack_records(d, st, gk, n, s64, view, tile, tid);
}

template <int CAP, int NBK>
static cudaError_t launch_tsp(const Dims& d, const SpfRasterState& st, const ControlLayout& cl, cudaStream_t s) {
  const size_t smem = sizeof(SortSmem<CAP, NBK>);
  cudaError_t e = cudaFuncSetAttribute(tile_sort_pack_kernel<CAP, NBK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  pdl_launch(tile_sort_pack_kernel<CAP, NBK>, d.B * d.T, TILE_THREADS, smem, s)(d, st, st.control + cl.tile_start);
  return cudaGetLastError();
}

cudaError_t launch_tile_sort_pack(const Dims& d, const SpfRasterState& st, const ControlLayout& cl,
                                  cudaStream_t s) {
  // shared-memory variant from the expected list length: dup_capacity is the caller's high-water mark of the
  // duplicate count (about 1.5 x N), so cap / tiles over-estimates the mean list; about twice the mean covers the
  // spread between tiles (longer lists still work: in-place global sort)
  const int64_t want = 4 * d.cap / (3 * (int64_t)d.B * d.T);   // cap ~ 1.5 x N  ->  ~2 x the mean list length
  if (want <= 1024) return launch_tsp<1024, 1024>(d, st, cl, s);     // 12 KB
  if (want <= 2048) return launch_tsp<2048, 2048>(d, st, cl, s);     // 24 KB
  if (want <= 4096) return launch_tsp<4096, 4096>(d, st, cl, s);     // 48 KB
  return launch_tsp<8192, 8192>(d, st, cl, s);                       // 96 KB
}

// ---- parity helper: (point_list, keys) out of the slab -------------------------------------------
__global__ void unpack_sorted_kernel(Dims d, SpfRasterState st, const int* __restrict__ tile_start,
                                     int64_t n, int32_t* point_list, uint64_t* keys) {
  pdl_enter();
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
  pdl_launch(unpack_sorted_kernel, (unsigned)((n + threads - 1) / threads), threads, 0, s)(
      d, st, st.control + cl.tile_start, n, point_list, keys);
  return cudaGetLastError();
}

}  // namespace spf
