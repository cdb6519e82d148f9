// K5 / K6: per-16x16-tile front-to-back alpha blend (forward) and its backward(s).
//
// One CTA (256 threads, one pixel each; every warp owns an 8x4 pixel block) per (view, tile).  The tile's
// depth-sorted slab -- contiguous 48-byte records {xy, conic, opacity, rgb, depth, slot, id} -- and the matching
// 16-byte alpha bounding boxes are streamed through a double-buffered shared-memory ring with 1-D TMA bulk
// copies (cp.async.bulk + mbarrier complete_tx).
//
// Culling: with the encoder's scale law most splats are ~2 px wide, so >90% of (pixel, Gaussian) pairs of a tile
// are rejections.  Each warp therefore first tests 32 records at a time, one record per LANE, against its 8x4
// pixel block (box vs box, conflict-free LDS.128), ballots, and then walks only the set bits with all 32 lanes
// evaluating their own pixel.  The boxes are conservative (binning.cu: alpha_bbox), the per-pixel test is
// unchanged, so results are bit-identical to the unculled loop.
//
// Backward, three generations (all atomic-free on floats, fixed summation order => bit-reproducible):
//   blend_backward_log_kernel  (default)  consumes the PAIR LOG the forward writes when a backward will follow:
//                              one contributing (pixel, Gaussian) pair per lane, closed-form dL/dalpha, segmented
//                              shuffle reduction per Gaussian, exchange slots for Gaussians overlapping several
//                              warp regions.
//   blend_backward_kernel      recomputing pair-compaction kernel: per-tile fallback when a warp's log overflowed,
//                              and the backward when no log was requested.
//   blend_backward_v1_kernel   first generation (back to front, 12-shuffle butterfly per hit), SPF_FLAG_BWD_V1,
//                              kept as an independent cross-check.
//
// Replaces renderCUDA forward/backward of diff_gauss_pose (SURVEY.md App. B "Blend forward/backward").
#include "spf_device.cuh"
#include "spf_kernels.h"
#include "spf_math.h"

namespace spf {

constexpr int CH_F = 256;  // records per forward chunk  (12 KB slab + 4 KB boxes)
constexpr int CH_B = 64;   // records per backward chunk

__device__ __forceinline__ void warp_block_of_thread(int tid, int tile, int gx, int& bx, int& by) {
  const int w = tid >> 5;
  const int tx = tile % gx, ty = tile / gx;
  bx = tx * TILE + (w & 1) * 8;
  by = ty * TILE + (w >> 1) * 4;
}

// alpha of one record at one pixel; shared by forward and backward so both make identical
// accept / reject decisions.
__device__ __forceinline__ bool eval_alpha(const float4& a, const float4& b, float pxf, float pyf, float& dx,
                                           float& dy, float& G, float& alpha) {
  dx = a.x - pxf;
  dy = a.y - pyf;
  const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
  G = expf(power);
  alpha = fminf(ALPHA_MAX, b.y * G);
  return (power <= 0.0f) && (alpha >= ALPHA_MIN);
}

template <bool TMA>
__device__ __forceinline__ void stage_chunk(float4* dst, float4* dst_box, const float4* __restrict__ src,
                                            const float4* __restrict__ src_box, int cnt, uint64_t* bar, int tid) {
  if (TMA) {
    if (tid == 0) {
      mbar_expect_tx(bar, (uint32_t)cnt * 64u);
      tma_load_1d(dst, src, (uint32_t)cnt * 48u, bar);
      tma_load_1d(dst_box, src_box, (uint32_t)cnt * 16u, bar);
    }
  } else {
    for (int i = tid; i < cnt * 3; i += TILE_THREADS) dst[i] = src[i];
    for (int i = tid; i < cnt; i += TILE_THREADS) dst_box[i] = src_box[i];
  }
}

__device__ __forceinline__ bool box_hits(const float4& bb, float x0, float x1, float y0, float y1) {
  return (bb.x <= x1) && (bb.y >= x0) && (bb.z <= y1) && (bb.w >= y0);
}

// Pair log (LOG = true, enabled when a backward pass will follow): every warp appends one 32-byte record per
// CONTRIBUTING (pixel, Gaussian) pair, in hit order (pairs of one Gaussian adjacent, in lane order):
//   word0 = record index in the tile list | lane << 25;  G;  T_before;  the blended sums (rgb, depth) AFTER this pair.
// Runs are padded so that none straddles a 32-record boundary of the log (PAIR_SKIP marker).
// With these the backward needs no per-pixel sequential pass at all (see blend_backward_log_kernel).  A warp's
// segment holds `pair_capacity` records; a warp that needs more stops writing and reports -1 (its tile is then
// handled by the recomputing v2 backward).  The largest per-warp count goes to control[2] for the host's sizing.
constexpr unsigned PAIR_J_MASK = (1u << 25) - 1u;
constexpr unsigned PAIR_SKIP = 0xffffffffu;   // word0 of a padding record: the rest of this 32-record block is unused

template <bool TMA, bool LOG>
__global__ void __launch_bounds__(TILE_THREADS)
blend_forward_kernel(Dims d, const float* __restrict__ bg_all, SpfRasterState st, SpfRasterOut out) {
  pdl_enter();
  __shared__ __align__(128) float4 buf[2][CH_F * 3];
  __shared__ __align__(128) float4 box[2][CH_F];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int blk_pairs;
  const int tid = threadIdx.x, lane = tid & 31;
  const int t = blockIdx.x;
  const int view = t / d.T, tile = t - view * d.T;
  const int s = st.tile_ranges[2 * (size_t)t], e = st.tile_ranges[2 * (size_t)t + 1];
  const int L = e - s;
  int bx, by;
  warp_block_of_thread(tid, tile, d.gx, bx, by);
  const int px = bx + (lane & 7), py = by + (lane >> 3);
  const bool inside = (px < d.W) && (py < d.H);
  const float pxf = (float)px, pyf = (float)py;
  const float wx0 = (float)bx, wx1 = (float)(bx + 7), wy0 = (float)by, wy1 = (float)(by + 3);
  if (TMA || LOG) {
    if (tid == 0) {
      if (TMA) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
      blk_pairs = 0;
    }
    __syncthreads();
  }
  const float4* slab = reinterpret_cast<const float4*>(st.slab) + 3 * (size_t)s;
  const float4* cull = reinterpret_cast<const float4*>(st.cullbox) + (size_t)s;
  const int nchunks = (L + CH_F - 1) / CH_F;
  const int Cw = d.pair_cap;
  uint4* plog = LOG ? reinterpret_cast<uint4*>(st.pair_log) + ((size_t)t * 8 + (tid >> 5)) * (size_t)Cw * 2 : nullptr;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned lane_bits = (unsigned)lane << 25;
  int room = Cw;   // pair-log records this warp may still write

  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
  int last = 0;
  bool done = !inside;
  int pending = -1;  // chunk whose TMA load is in flight beyond the current one

  // Lists of at most two chunks (the common case) live entirely in the two buffers: nothing is ever refilled, so the
  // warps need no block barrier at all -- each waits on the chunk's mbarrier and leaves as soon as ITS 32 pixels are
  // done.  Longer lists recycle the buffers and keep the block in step.
  const bool free_running = TMA && (nchunks <= 2);
  if (nchunks > 0) stage_chunk<TMA>(buf[0], box[0], slab, cull, min(CH_F, L), &bar[0], tid);
  for (int c = 0; c < nchunks; ++c) {
    const int cnt = min(CH_F, L - c * CH_F);
    if (c + 1 < nchunks) {
      stage_chunk<TMA>(buf[(c + 1) & 1], box[(c + 1) & 1], slab + 3 * (size_t)(c + 1) * CH_F,
                       cull + (size_t)(c + 1) * CH_F, min(CH_F, L - (c + 1) * CH_F), &bar[(c + 1) & 1], tid);
      pending = c + 1;
    } else {
      pending = -1;
    }
    if (TMA) mbar_wait(&bar[c & 1], (uint32_t)((c >> 1) & 1));
    else __syncthreads();
    if (!__all_sync(0xffffffffu, done)) {
      const float4* rec = buf[c & 1];
      const float4* bb = box[c & 1];
      const int base = c * CH_F;
      for (int g0 = 0; g0 < cnt; g0 += 32) {
        const int r = g0 + lane;
        bool hit = false;
        if (r < cnt) hit = box_hits(bb[r], wx0, wx1, wy0, wy1);
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int j = g0 + __ffs(mask) - 1;
          mask &= mask - 1;
          const float4 a = rec[3 * j], b = rec[3 * j + 1];
          float dx, dy, G, alpha;
          bool ok = eval_alpha(a, b, pxf, pyf, dx, dy, G, alpha) && !done;
          const float test_T = T * (1.0f - alpha);
          if (ok && test_T < T_STOP) { done = true; ok = false; }
          const float Tb = T;
          if (ok) {
            const float4 cc = rec[3 * j + 2];
            const float w = alpha * T;
            C0 += b.z * w; C1 += b.w * w; C2 += cc.x * w; D += cc.y * w;
            T = test_T;
            last = base + j + 1;
          }
          if (LOG) {
            const unsigned cb = __ballot_sync(0xffffffffu, ok);
            if (cb) {
              const int n = __popc(cb);
              // a run (the pairs of one record) never straddles a 32-record boundary of the log: skip to the next
              // boundary (one lane leaves a SKIP marker) so that the backward can walk the log in fixed 32-record
              // batches with every run whole and every batch address known in advance
              const int pos32 = (Cw - room) & 31;
              if (pos32 + n > 32) {
                const int pad = 32 - pos32;
                if (lane == 0 && room > 0) plog[0].x = PAIR_SKIP;
                plog += 2 * pad;
                room -= pad;
              }
              room -= n;             // keeps counting past the capacity: reports the size that would have been needed
              if (ok && room >= 0) {
                uint4* dst = plog + 2 * __popc(cb & lt_mask);
                dst[0] = make_uint4((unsigned)(base + j) | lane_bits, __float_as_uint(G), __float_as_uint(Tb), __float_as_uint(C0));
                dst[1] = make_uint4(__float_as_uint(C1), __float_as_uint(C2), __float_as_uint(D), 0u);
              }
              plog += 2 * n;
            }
          }
        }
      }
    }
    if (free_running) {
      if (__all_sync(0xffffffffu, done)) break;
    } else if (__syncthreads_and(done)) {
      break;
    }
  }
  if (TMA && pending >= 0 && tid == 0) mbar_wait(&bar[pending & 1], (uint32_t)((pending >> 1) & 1));
  if (LOG) {
    if (lane == 0) {
      const int npairs = Cw - room;
      const bool okw = (room >= 0 && L <= (int)PAIR_J_MASK);
      st.pair_count[(size_t)t * 8 + (tid >> 5)] = okw ? npairs : -1;
      if (!okw) st.control[3] = 1;
      atomicMax(&blk_pairs, npairs);
    }
    __syncthreads();
    if (tid == 0 && blk_pairs > 0) atomicMax(st.control + 2, blk_pairs);
  }

  if (inside) {
    const float* bg = bg_all + view * 3;
    const size_t hw = (size_t)d.H * d.W;
    const size_t pix = (size_t)py * d.W + px;
    float* col = out.color + (size_t)view * 3 * hw;
    col[pix] = C0 + T * bg[0];
    col[hw + pix] = C1 + T * bg[1];
    col[2 * hw + pix] = C2 + T * bg[2];
    out.depth[(size_t)view * hw + pix] = D;
    if (out.alpha) out.alpha[(size_t)view * hw + pix] = 1.0f - T;
    st.final_T[(size_t)view * hw + pix] = T;
    st.n_contrib[(size_t)view * hw + pix] = last;
    // private copy of the blended sums (no background term): backward derives every pixel's total
    // sum_j q_j w_j from it, independent of what the caller does with its output tensors afterwards
    reinterpret_cast<float4*>(st.accum)[(size_t)view * hw + pix] = make_float4(C0, C1, C2, D);
  }
}

cudaError_t launch_blend_forward(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st,
                                 const SpfRasterOut& out, cudaStream_t s) {
  const int grid = d.B * d.T;
  const bool log = st.pair_log != nullptr && st.pair_count != nullptr && d.pair_cap > 0;
  if (d.flags & SPF_FLAG_NO_TMA) {
    if (log) pdl_launch(blend_forward_kernel<false, true>, grid, TILE_THREADS, 0, s)(d, in.bg, st, out);
    else pdl_launch(blend_forward_kernel<false, false>, grid, TILE_THREADS, 0, s)(d, in.bg, st, out);
  } else {
    if (log) pdl_launch(blend_forward_kernel<true, true>, grid, TILE_THREADS, 0, s)(d, in.bg, st, out);
    else pdl_launch(blend_forward_kernel<true, false>, grid, TILE_THREADS, 0, s)(d, in.bg, st, out);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Warp reduction of 10 per-lane values in 12 shuffles (recursive halving).  On return lane L holds,
// in `out`, the warp total of component  comp_of_lane(L)  (or padding).
__device__ __forceinline__ float halving_step(float keep_lo, float keep_hi, bool hi, int xor_mask) {
  const float keep = hi ? keep_hi : keep_lo;
  const float send = hi ? keep_lo : keep_hi;
  return keep + __shfl_xor_sync(0xffffffffu, send, xor_mask);
}
__device__ __forceinline__ float warp_reduce10(const float v[10], int lane) {
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
  float u[6];
#pragma unroll
  for (int k = 0; k < 5; ++k) u[k] = halving_step(v[k], v[k + 5], h4, 16);
  u[5] = 0.0f;
  float w[4];
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = halving_step(u[k], u[k + 3], h3, 8);
  w[3] = 0.0f;
  float x[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) x[k] = halving_step(w[k], w[k + 2], h2, 4);
  float y = halving_step(x[0], x[1], h1, 2);
  y += __shfl_xor_sync(0xffffffffu, y, 1);
  return y;
}
// component held by a lane after warp_reduce10, or -1 (padding / duplicate)
__device__ __forceinline__ int comp_of_lane(int lane) {
  if (lane & 1) return -1;
  const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
  const int in3 = 2 * b2 + b1;          // index inside a group of 3 (+1 pad)
  if (in3 > 2) return -1;
  const int in5 = 3 * b3 + in3;         // index inside a group of 5 (+1 pad)
  if (in5 > 4) return -1;
  return 5 * b4 + in5;
}

template <bool TMA>
__global__ void __launch_bounds__(TILE_THREADS)
blend_backward_v1_kernel(Dims d, const float* __restrict__ bg_all, SpfRasterState st, SpfRasterGradOut go,
                      float* __restrict__ dup_grad) {
  pdl_enter();
  __shared__ __align__(128) float4 buf[2][CH_B * 3];
  __shared__ __align__(128) float4 box[2][CH_B];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ float part[8][CH_B][10];
  __shared__ unsigned long long hitmask[8];
  __shared__ int max_contrib_s;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t = blockIdx.x;
  const int view = t / d.T, tile = t - view * d.T;
  const int s = st.tile_ranges[2 * (size_t)t], e = st.tile_ranges[2 * (size_t)t + 1];
  const int L = e - s;
  if (L == 0) return;
  int bx, by;
  warp_block_of_thread(tid, tile, d.gx, bx, by);
  const int px = bx + (lane & 7), py = by + (lane >> 3);
  const bool inside = (px < d.W) && (py < d.H);
  const float pxf = (float)px, pyf = (float)py;
  const float wx0 = (float)bx, wx1 = (float)(bx + 7), wy0 = (float)by, wy1 = (float)(by + 3);
  const size_t hw = (size_t)d.H * d.W;
  const size_t pix = (size_t)py * d.W + px;
  const int mycomp = comp_of_lane(lane);

  if (tid == 0) {
    max_contrib_s = 0;
    if (TMA) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  }
  __syncthreads();

  float T_final = 1.0f, g0 = 0.f, g1 = 0.f, g2 = 0.f, gd = 0.f, ga = 0.f;
  int ncontrib = 0;
  if (inside) {
    T_final = st.final_T[(size_t)view * hw + pix];
    ncontrib = st.n_contrib[(size_t)view * hw + pix];
    if (go.dL_dcolor) {
      const float* gc = go.dL_dcolor + (size_t)view * 3 * hw;
      g0 = gc[pix]; g1 = gc[hw + pix]; g2 = gc[2 * hw + pix];
    }
    if (go.dL_ddepth) gd = go.dL_ddepth[(size_t)view * hw + pix];
    if (go.dL_dalpha) ga = go.dL_dalpha[(size_t)view * hw + pix];
  }
  int wmax = ncontrib;   // warp-level last contributor
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if (lane == 0) atomicMax(&max_contrib_s, wmax);
  __syncthreads();
  const int maxc = max_contrib_s;
  const float4* slab = reinterpret_cast<const float4*>(st.slab) + 3 * (size_t)s;
  const float4* cull = reinterpret_cast<const float4*>(st.cullbox) + (size_t)s;

  // duplicates nobody reached: zero gradient records
  for (int i = maxc * 10 + tid; i < L * 10; i += TILE_THREADS) {
    const int r = i / 10, k = i - r * 10;
    const int slot = __float_as_int(__ldg(reinterpret_cast<const float*>(slab + 3 * r + 2) + 2));
    dup_grad[(size_t)slot * 12 + k] = 0.0f;
  }
  if (maxc == 0) return;

  const float* bg = bg_all + view * 3;
  const float bgdot = bg[0] * g0 + bg[1] * g1 + bg[2] * g2 - ga;
  float T = T_final;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  float lv0 = 0.f, lv1 = 0.f, lv2 = 0.f, lv3 = 0.f, last_alpha = 0.f;

  const int ctop = (maxc - 1) / CH_B;
  stage_chunk<TMA>(buf[0], box[0], slab + 3 * (size_t)ctop * CH_B, cull + (size_t)ctop * CH_B,
                   min(CH_B, maxc - ctop * CH_B), &bar[0], tid);
  int it = 0;
  for (int c = ctop; c >= 0; --c, ++it) {
    const int cnt = min(CH_B, maxc - c * CH_B);
    if (c > 0)
      stage_chunk<TMA>(buf[(it + 1) & 1], box[(it + 1) & 1], slab + 3 * (size_t)(c - 1) * CH_B,
                       cull + (size_t)(c - 1) * CH_B, CH_B, &bar[(it + 1) & 1], tid);
    if (TMA) mbar_wait(&bar[it & 1], (uint32_t)((it >> 1) & 1));
    else __syncthreads();

    const float4* rec = buf[it & 1];
    const float4* bb = box[it & 1];
    unsigned long long mymask = 0ull;
    if (c * CH_B < wmax) {   // some pixel of this warp reaches into the chunk
      for (int g0i = ((cnt - 1) >> 5) << 5; g0i >= 0; g0i -= 32) {
        const int r = g0i + lane;
        bool hit = false;
        if (r < cnt && (c * CH_B + r) < wmax) hit = box_hits(bb[r], wx0, wx1, wy0, wy1);
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int bit = 31 - __clz(mask);
          mask &= ~(1u << bit);
          const int j = g0i + bit;
          const int idx = c * CH_B + j;
          const float4 a = rec[3 * j], b = rec[3 * j + 1];
          float dx, dy, G, alpha;
          const bool contrib = eval_alpha(a, b, pxf, pyf, dx, dy, G, alpha) && (idx < ncontrib);
          if (!__any_sync(0xffffffffu, contrib)) continue;
          float v[10];
#pragma unroll
          for (int k = 0; k < 10; ++k) v[k] = 0.0f;
          if (contrib) {
            const float4 cc = rec[3 * j + 2];
            T = T / (1.0f - alpha);
            const float w = alpha * T;
            float dL_dalpha = 0.0f;
            acc0 = last_alpha * lv0 + (1.0f - last_alpha) * acc0; lv0 = b.z;  dL_dalpha += (b.z - acc0) * g0;
            acc1 = last_alpha * lv1 + (1.0f - last_alpha) * acc1; lv1 = b.w;  dL_dalpha += (b.w - acc1) * g1;
            acc2 = last_alpha * lv2 + (1.0f - last_alpha) * acc2; lv2 = cc.x; dL_dalpha += (cc.x - acc2) * g2;
            acc3 = last_alpha * lv3 + (1.0f - last_alpha) * acc3; lv3 = cc.y; dL_dalpha += (cc.y - acc3) * gd;
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.0f - alpha)) * bgdot;
            const float dL_dG = b.y * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddx = -gdx * a.z - gdy * a.w;
            const float dG_ddy = -gdy * b.x - gdx * a.w;
            v[0] = dL_dG * dG_ddx;
            v[1] = dL_dG * dG_ddy;
            v[2] = -0.5f * gdx * dx * dL_dG;
            v[3] = -gdx * dy * dL_dG;
            v[4] = -0.5f * gdy * dy * dL_dG;
            v[5] = G * dL_dalpha;
            v[6] = w * g0; v[7] = w * g1; v[8] = w * g2; v[9] = w * gd;
          }
          const float tot = warp_reduce10(v, lane);
          if (mycomp >= 0) part[wid][j][mycomp] = tot;
          mymask |= (1ull << j);
        }
      }
    }
    if (lane == 0) hitmask[wid] = mymask;
    __syncthreads();
    // cross-warp reduction in fixed warp order, one plain store per (record, component)
    for (int i = tid; i < cnt * 10; i += TILE_THREADS) {
      const int j = i / 10, k = i - j * 10;
      float sum = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w)
        if ((hitmask[w] >> j) & 1ull) sum += part[w][j][k];
      const int slot = __float_as_int(reinterpret_cast<const float*>(rec + 3 * j + 2)[2]);
      dup_grad[(size_t)slot * 12 + k] = sum;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// Blend backward, v2: front-to-back traversal with per-warp PAIR COMPACTION.
//
// ncu on v1 (profiles/r1_summary_baseline.md): a warp walks ~57 box-hit records per tile, but on average
// only 4 of its 32 pixels actually receive a contribution from a hit record, and the 130-instruction
// gradient + butterfly body ran at 4/32 lane efficiency.  v2 splits the work:
//
//   test phase  (per hit record, all 32 lanes = pixels, same ~45-instruction alpha test as the forward):
//       contributing lanes update their running transmittance T and prefix P = sum_{j<=i} q_j w_j
//       (q_j = rgb_j . dL/dC + depth_j dL/dD, w_j = alpha_j T_j) and push one 16-byte pair record
//       {pixel, record, G, T_before, P_after} into the warp's ring queue in shared memory
//       (ballot + popc compaction; pairs of one record are adjacent, in lane order).
//   dense phase (whenever 32 pairs are queued): ONE PAIR PER LANE.  With the closed form
//       dL/dalpha_i = T_i q_i - (Qtot - P_i) / (1 - alpha_i),   Qtot = sum_j q_j w_j + T_final (bg.dL/dC - dL/dA)
//       every pair is independent, so all 32 lanes do useful gradient math; a segmented shuffle reduction
//       over runs of equal record (early exit at the longest run) leaves each record's 10 partial sums in
//       its run-head lane, which accumulates them into part[warp][record].
//   per chunk   : fixed-order cross-warp sum of part[] and ONE plain store per (Gaussian, tile) duplicate.
//
// Still no atomics and a fixed summation order: bit-reproducible run to run.  Front-to-back order also removes
// v1's T reconstruction by repeated division.
constexpr int CH_B2 = 128;   // records per backward chunk
constexpr int QCAP = 64;     // pair-queue ring capacity per warp (power of two, >= 63)

struct BwdSmem {
  float4 rec[2][CH_B2 * 3];
  float4 box[2][CH_B2];
  uint4 queue[8][QCAP];
  float part[8][CH_B2][10];
  float4 pg[TILE_THREADS];    // per pixel: dL/dC (3), dL/dD
  float pq[TILE_THREADS];     // per pixel: Qtot
  uint64_t bar[2];
  int max_contrib;
};

template <bool TMA>
__device__ __forceinline__ void blend_backward_tile(const Dims& d, const float* __restrict__ bg_all, const SpfRasterState& st,
                                                    const SpfRasterGradOut& go, float* __restrict__ dup_grad, int use_log,
                                                    BwdSmem& S, const int t) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int view = t / d.T, tile = t - view * d.T;
  const int s = st.tile_ranges[2 * (size_t)t], e = st.tile_ranges[2 * (size_t)t + 1];
  const int L = e - s;
  if (L == 0) return;
  if (use_log) {   // tiles whose 8 warp logs are complete were handled by blend_backward_log_kernel
    bool neg = false;
#pragma unroll
    for (int w = 0; w < 8; ++w) neg |= st.pair_count[(size_t)t * 8 + w] < 0;
    if (!neg) return;
  }
  int bx, by;
  warp_block_of_thread(tid, tile, d.gx, bx, by);
  const int px = bx + (lane & 7), py = by + (lane >> 3);
  const bool inside = (px < d.W) && (py < d.H);
  const float pxf = (float)px, pyf = (float)py;
  const float wx0 = (float)bx, wx1 = (float)(bx + 7), wy0 = (float)by, wy1 = (float)(by + 3);
  const size_t hw = (size_t)d.H * d.W;
  const size_t pix = (size_t)py * d.W + px;

  if (tid == 0) {
    S.max_contrib = 0;
    if (TMA) { mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1); mbar_fence_init(); }
  }
  __syncthreads();

  float g0 = 0.f, g1 = 0.f, g2 = 0.f, gd = 0.f, qtot = 0.f;
  int ncontrib = 0;
  if (inside) {
    ncontrib = st.n_contrib[(size_t)view * hw + pix];
    float ga = 0.f;
    if (go.dL_dcolor) {
      const float* gc = go.dL_dcolor + (size_t)view * 3 * hw;
      g0 = gc[pix]; g1 = gc[hw + pix]; g2 = gc[2 * hw + pix];
    }
    if (go.dL_ddepth) gd = go.dL_ddepth[(size_t)view * hw + pix];
    if (go.dL_dalpha) ga = go.dL_dalpha[(size_t)view * hw + pix];
    const float4 acc = reinterpret_cast<const float4*>(st.accum)[(size_t)view * hw + pix];
    const float* bg = bg_all + view * 3;
    const float bgdot = bg[0] * g0 + bg[1] * g1 + bg[2] * g2 - ga;
    qtot = (acc.x * g0 + acc.y * g1) + (acc.z * g2 + acc.w * gd) + st.final_T[(size_t)view * hw + pix] * bgdot;
  }
  S.pg[tid] = make_float4(g0, g1, g2, gd);
  S.pq[tid] = qtot;
  int wmax = ncontrib;   // warp-level last contributor
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if (lane == 0) atomicMax(&S.max_contrib, wmax);
  __syncthreads();
  const int maxc = S.max_contrib;
  const float4* slab = reinterpret_cast<const float4*>(st.slab) + 3 * (size_t)s;
  const float4* cull = reinterpret_cast<const float4*>(st.cullbox) + (size_t)s;

  // duplicates nobody reached: zero gradient records
  for (int i = maxc * 10 + tid; i < L * 10; i += TILE_THREADS) {
    const int r = i / 10, k = i - r * 10;
    const int slot = __float_as_int(__ldg(reinterpret_cast<const float*>(slab + 3 * r + 2) + 2));
    dup_grad[(size_t)slot * 12 + k] = 0.0f;
  }
  if (maxc == 0) return;

  float T = 1.0f, P = 0.0f;
  uint4* q = S.queue[wid];
  float (*part)[10] = S.part[wid];
  const float4* pgw = S.pg + wid * 32;
  const float* pqw = S.pq + wid * 32;
  const unsigned lt_mask = (1u << lane) - 1u;
  int qhead = 0, qcount = 0;

  const int nchunks = (maxc + CH_B2 - 1) / CH_B2;
  stage_chunk<TMA>(S.rec[0], S.box[0], slab, cull, min(CH_B2, maxc), &S.bar[0], tid);
  for (int c = 0; c < nchunks; ++c) {
    const int cnt = min(CH_B2, maxc - c * CH_B2);
    if (c + 1 < nchunks)
      stage_chunk<TMA>(S.rec[(c + 1) & 1], S.box[(c + 1) & 1], slab + 3 * (size_t)(c + 1) * CH_B2,
                       cull + (size_t)(c + 1) * CH_B2, min(CH_B2, maxc - (c + 1) * CH_B2), &S.bar[(c + 1) & 1], tid);
    // zero this warp's partial sums for the chunk's records (128-bit stores)
    {
      float4* p4 = reinterpret_cast<float4*>(&part[0][0]);
      const int n4 = (cnt * 10 + 3) >> 2;
      for (int i = lane; i < n4; i += 32) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (TMA) mbar_wait(&S.bar[c & 1], (uint32_t)((c >> 1) & 1));
    else __syncthreads();
    __syncwarp();

    const float4* rec = S.rec[c & 1];
    const float4* bb = S.box[c & 1];
    const int base = c * CH_B2;

    // dense phase over the first n queued pairs (n <= 32)
    auto dense = [&](int n) {
      int j = -1;
      float v[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) v[k] = 0.0f;
      if (lane < n) {
        const uint4 en = q[(qhead + lane) & (QCAP - 1)];
        j = (int)(en.x & 0xffffu);
        const int pl = (int)(en.x >> 16);
        const float G = __uint_as_float(en.y), Tb = __uint_as_float(en.z), Pa = __uint_as_float(en.w);
        const float4 a = rec[3 * j], b = rec[3 * j + 1], cc = rec[3 * j + 2];
        const float4 g = pgw[pl];
        const float dx = a.x - (float)(bx + (pl & 7)), dy = a.y - (float)(by + (pl >> 3));
        const float alpha = fminf(ALPHA_MAX, b.y * G);
        const float qv = (b.z * g.x + b.w * g.y) + (cc.x * g.z + cc.y * g.w);
        const float dL_dalpha = Tb * qv - __fdividef(pqw[pl] - Pa, 1.0f - alpha);
        const float dL_dG = b.y * dL_dalpha;
        const float gdx = G * dx, gdy = G * dy;
        v[0] = dL_dG * (-gdx * a.z - gdy * a.w);
        v[1] = dL_dG * (-gdy * b.x - gdx * a.w);
        v[2] = -0.5f * gdx * dx * dL_dG;
        v[3] = -gdx * dy * dL_dG;
        v[4] = -0.5f * gdy * dy * dL_dG;
        v[5] = G * dL_dalpha;
        const float w = alpha * Tb;
        v[6] = w * g.x; v[7] = w * g.y; v[8] = w * g.z; v[9] = w * g.w;
      }
      // segmented reduction: runs of equal record index are contiguous; sums end in the run-head lane
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int jo = __shfl_down_sync(0xffffffffu, j, off);
        const bool same = (lane + off < 32) && (jo == j) && (j >= 0);
        if (!__any_sync(0xffffffffu, same)) break;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const float vo = __shfl_down_sync(0xffffffffu, v[k], off);
          if (same) v[k] += vo;
        }
      }
      const int jprev = __shfl_up_sync(0xffffffffu, j, 1);
      if (j >= 0 && (lane == 0 || jprev != j)) {
#pragma unroll
        for (int k = 0; k < 10; ++k) part[j][k] += v[k];
      }
      __syncwarp();
      qhead = (qhead + n) & (QCAP - 1);
      qcount -= n;
    };

    if (base < wmax) {   // some pixel of this warp reaches into the chunk
      const int lim = min(cnt, wmax - base);
      for (int g0i = 0; g0i < lim; g0i += 32) {
        const int r = g0i + lane;
        bool hit = false;
        if (r < lim) hit = box_hits(bb[r], wx0, wx1, wy0, wy1);
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int j = g0i + __ffs(mask) - 1;
          mask &= mask - 1;
          const float4 a = rec[3 * j], b = rec[3 * j + 1];
          float dx, dy, G, alpha;
          const bool contrib = eval_alpha(a, b, pxf, pyf, dx, dy, G, alpha) && (base + j < ncontrib);
          const unsigned cb = __ballot_sync(0xffffffffu, contrib);
          if (cb == 0u) continue;
          if (contrib) {
            const float2 cc = *reinterpret_cast<const float2*>(rec + 3 * j + 2);
            const float qv = (b.z * g0 + b.w * g1) + (cc.x * g2 + cc.y * gd);
            P = P + qv * (alpha * T);
            const int pos = (qhead + qcount + __popc(cb & lt_mask)) & (QCAP - 1);
            q[pos] = make_uint4((unsigned)j | ((unsigned)lane << 16), __float_as_uint(G), __float_as_uint(T),
                                __float_as_uint(P));
            T = T * (1.0f - alpha);
          }
          qcount += __popc(cb);
          if (qcount >= 32) {
            __syncwarp();
            dense(32);
          }
        }
      }
      // the chunk's records leave shared memory after this chunk: drain the queue
      __syncwarp();
      if (qcount > 0) dense(qcount);
    }
    __syncthreads();
    // cross-warp reduction in fixed warp order, one plain store per (record, component)
    for (int i = tid; i < cnt * 10; i += TILE_THREADS) {
      const int j = i / 10, k = i - j * 10;
      float sum = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += S.part[w][j][k];
      const int slot = __float_as_int(reinterpret_cast<const float*>(rec + 3 * j + 2)[2]);
      dup_grad[(size_t)slot * 12 + k] = sum;
    }
    __syncthreads();
  }
}

// Grid-stride over (view, tile): with the pair log on, almost every tile is skipped after reading its 8 counters,
// so the launch is sized to the machine (resident CTAs) instead of one CTA per tile.
template <bool TMA>
__global__ void __launch_bounds__(TILE_THREADS, 3)
blend_backward_kernel(Dims d, const float* __restrict__ bg_all, SpfRasterState st, SpfRasterGradOut go,
                      float* __restrict__ dup_grad, int use_log) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdSmem& S = *reinterpret_cast<BwdSmem*>(smem_raw);
  const int n_tiles = d.B * d.T;
  if (use_log && st.control[3] == 0) return;     // every tile was handled from the pair log
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    blend_backward_tile<TMA>(d, bg_all, st, go, dup_grad, use_log, S, t);
    __syncthreads();     // shared memory (incl. the mbarriers, re-initialised per tile) is reused by the next tile
  }
}

// ---------------------------------------------------------------------------------------------------
// Blend backward, v3: consumes the forward's pair log -- no alpha tests, no per-pixel sequential state.
//
// For a logged pair i of pixel p:  dL/dalpha_i = T_i q_i - (Qtot_p - g_p . S_i) / (1 - alpha_i), with S_i the blended
// sums after the pair (logged), g_p = (dL/dC, dL/dD) and Qtot_p = g_p . S_final + T_final (bg . dL/dC - dL/dA).
// Every pair is independent: warp w walks the log of forward-warp w 32 pairs at a time, ONE PAIR PER LANE (whole
// runs only: a batch is cut at the last run boundary, so a Gaussian's pairs of this warp are always reduced in
// one segmented shuffle reduction).  A Gaussian whose alpha box overlaps a single 8x4 warp region (the common case
// with SPFSplatV2's ~2 px splats) is final after that reduction and its 10 sums are stored straight to its
// duplicate slot; one that overlaps several regions parks its per-warp sums in shared-memory exchange slots that
// are added in fixed region order after ONE block barrier.  No atomics on floats, fixed order: bit-reproducible.
// Long tile lists are processed in windows of LOG_W records (each warp's log is sorted by record).  Tiles with an
// incomplete log or too many multi-region records in a window are left to the recomputing kernel above (flagged
// through pair_count).
constexpr int LOG_W = 416;         // records per window of the tile list (staged in shared memory)
constexpr int LOG_ESLOTS = 640;    // exchange slots per window (one per (multi-region record, overlapped region))

struct LogSmem {
  float4 rec[LOG_W * 3];                 // the window's slab records
  float4 pg[TILE_THREADS];
  float pq[TILE_THREADS];
  uint32_t info[LOG_W];                  // region mask (8 bits) | first exchange slot << 8
  float exch[LOG_ESLOTS][10];
  uint32_t wrote[LOG_W / 32];
  int base;
};

__global__ void __launch_bounds__(TILE_THREADS, 4)
blend_backward_log_kernel(Dims d, const float* __restrict__ bg_all, SpfRasterState st, SpfRasterGradOut go,
                          float* __restrict__ dup_grad) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  LogSmem& S = *reinterpret_cast<LogSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t = blockIdx.x;
  const int view = t / d.T, tile = t - view * d.T;
  const int s = st.tile_ranges[2 * (size_t)t], e = st.tile_ranges[2 * (size_t)t + 1];
  const int L = e - s;
  if (L == 0) return;
  const int count = st.pair_count[(size_t)t * 8 + wid];
  if (__syncthreads_or(count < 0)) return;            // incomplete log: the recomputing kernel takes this tile
  int bx, by;
  warp_block_of_thread(tid, tile, d.gx, bx, by);
  const int px = bx + (lane & 7), py = by + (lane >> 3);
  const bool inside = (px < d.W) && (py < d.H);
  const size_t hw = (size_t)d.H * d.W;
  const size_t pix = (size_t)py * d.W + px;
  const int tx0 = (tile % d.gx) * TILE, ty0 = (tile / d.gx) * TILE;
  const float4* slab = reinterpret_cast<const float4*>(st.slab) + 3 * (size_t)s;
  const float4* cull = reinterpret_cast<const float4*>(st.cullbox) + (size_t)s;
  const uint4* lp = reinterpret_cast<const uint4*>(st.pair_log) + ((size_t)t * 8 + wid) * (size_t)d.pair_cap * 2;

  // the log is read in fixed 32-record batches (runs never straddle a batch: the forward pads), so batch addresses do
  // not depend on data: each batch is prefetched towards the SM three batches ahead and then loaded where it is
  // used (ptxas does not keep register loads in flight across the loop back-edge, so a register pipeline would
  // expose the full DRAM latency on every batch)
  const int nbatch = (count + 31) >> 5;
  for (int k = 0; k < 3; ++k)
    if ((k << 5) + lane < count) prefetch_l1(lp + 2 * (size_t)((k << 5) + lane));
  int bi = 0;

  // per-pixel upstream gradients and Qtot
  {
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, gd = 0.f, qtot = 0.f;
    if (inside) {
      float ga = 0.f;
      if (go.dL_dcolor) {
        const float* gc = go.dL_dcolor + (size_t)view * 3 * hw;
        g0 = gc[pix]; g1 = gc[hw + pix]; g2 = gc[2 * hw + pix];
      }
      if (go.dL_ddepth) gd = go.dL_ddepth[(size_t)view * hw + pix];
      if (go.dL_dalpha) ga = go.dL_dalpha[(size_t)view * hw + pix];
      const float4 acc = reinterpret_cast<const float4*>(st.accum)[(size_t)view * hw + pix];
      const float* bg = bg_all + view * 3;
      const float bgdot = bg[0] * g0 + bg[1] * g1 + bg[2] * g2 - ga;
      qtot = (acc.x * g0 + acc.y * g1) + (acc.z * g2 + acc.w * gd) + st.final_T[(size_t)view * hw + pix] * bgdot;
    }
    S.pg[tid] = make_float4(g0, g1, g2, gd);
    S.pq[tid] = qtot;
  }
  const float4* pgw = S.pg + wid * 32;
  const float* pqw = S.pq + wid * 32;

  for (int w0 = 0; w0 < L; w0 += LOG_W) {
    const int wn = min(LOG_W, L - w0), wend = w0 + wn;
    // ---- window prologue: stage the records, region masks (the same box_hits test the forward used), exchange slots
    if (tid == 0) S.base = 0;
    for (int i = tid; i < LOG_W / 32; i += TILE_THREADS) S.wrote[i] = 0u;
    {
      float4* z = reinterpret_cast<float4*>(&S.exch[0][0]);
      for (int i = tid; i < LOG_ESLOTS * 10 / 4; i += TILE_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* src = slab + 3 * (size_t)w0;
      for (int i = tid; i < wn * 3; i += TILE_THREADS) S.rec[i] = __ldg(src + i);
    }
    __syncthreads();
    for (int r0 = 0; r0 < wn; r0 += TILE_THREADS) {
      const int r = r0 + tid;
      unsigned mask = 0u;
      if (r < wn) {
        const float4 bb = __ldg(cull + w0 + r);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float x0 = (float)(tx0 + (w & 1) * 8), y0 = (float)(ty0 + (w >> 1) * 4);
          if (box_hits(bb, x0, x0 + 7.0f, y0, y0 + 3.0f)) mask |= 1u << w;
        }
      }
      const int nreg = __popc(mask);
      const int need = nreg > 1 ? nreg : 0;
      const int inc = warp_incl_scan_i(need, lane);
      const int tot = __shfl_sync(0xffffffffu, inc, 31);
      int wbase = 0;
      if (lane == 0 && tot > 0) wbase = atomicAdd(&S.base, tot);   // slot POSITIONS may vary run to run; sums do not
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (r < wn) S.info[r] = mask | ((unsigned)(wbase + inc - need) << 8);
    }
    __syncthreads();
    if (S.base > LOG_ESLOTS) {               // too many multi-region records: leave the tile to the recomputing kernel
      if (tid == 0) { st.pair_count[(size_t)t * 8] = -1; st.control[3] = 1; }
      return;
    }

    // ---- this warp's pairs whose record lies in the window
    while (bi < nbatch) {
      const int idx = (bi << 5) + lane;
      if (idx + 96 < count) prefetch_l1(lp + 2 * (size_t)(idx + 96));
      uint4 e0 = make_uint4(PAIR_SKIP, 0u, 0u, 0u), e1 = e0;
      if (idx < count) { e0 = lp[2 * (size_t)idx]; e1 = lp[2 * (size_t)idx + 1]; }
      // lanes at / after a SKIP marker (or past the end of the log) hold no pair
      const unsigned skipm = __ballot_sync(0xffffffffu, (idx >= count) || (e0.x == PAIR_SKIP));
      const int nv = skipm ? __ffs(skipm) - 1 : 32;
      const int jraw = (int)(e0.x & PAIR_J_MASK);
      const bool act = (lane < nv) && (jraw >= w0) && (jraw < wend);
      const bool beyond = (lane < nv) && (jraw >= wend);          // sorted by record: belongs to a later window
      const unsigned bym = __ballot_sync(0xffffffffu, beyond);
      int j = -1;
      float v[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) v[k] = 0.0f;
      int slot = 0;
      if (act) {
        j = jraw - w0;
        const int pl = (int)((e0.x >> 25) & 31u);
        const float G = __uint_as_float(e0.y), Tb = __uint_as_float(e0.z);
        const float4 a = S.rec[3 * j], b = S.rec[3 * j + 1], cc = S.rec[3 * j + 2];
        slot = __float_as_int(cc.z);
        const float4 g = pgw[pl];
        const float dx = a.x - (float)(bx + (pl & 7)), dy = a.y - (float)(by + (pl >> 3));
        const float alpha = fminf(ALPHA_MAX, b.y * G);
        const float qv = (b.z * g.x + b.w * g.y) + (cc.x * g.z + cc.y * g.w);
        const float sg = (__uint_as_float(e0.w) * g.x + __uint_as_float(e1.x) * g.y) +
                         (__uint_as_float(e1.y) * g.z + __uint_as_float(e1.z) * g.w);
        const float dL_dalpha = Tb * qv - __fdividef(pqw[pl] - sg, 1.0f - alpha);
        const float dL_dG = b.y * dL_dalpha;
        const float gdx = G * dx, gdy = G * dy;
        v[0] = dL_dG * (-gdx * a.z - gdy * a.w);
        v[1] = dL_dG * (-gdy * b.x - gdx * a.w);
        v[2] = -0.5f * gdx * dx * dL_dG;
        v[3] = -gdx * dy * dL_dG;
        v[4] = -0.5f * gdy * dy * dL_dG;
        v[5] = G * dL_dalpha;
        const float w = alpha * Tb;
        v[6] = w * g.x; v[7] = w * g.y; v[8] = w * g.z; v[9] = w * g.w;
      }
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int jo = __shfl_down_sync(0xffffffffu, j, off);
        const bool same = (lane + off < 32) && (jo == j) && (j >= 0);
        if (!__any_sync(0xffffffffu, same)) break;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const float vo = __shfl_down_sync(0xffffffffu, v[k], off);
          if (same) v[k] += vo;
        }
      }
      const int jprev = __shfl_up_sync(0xffffffffu, j, 1);
      if (act && (lane == 0 || jprev != j)) {          // run head: holds this warp's sums for record j
        const unsigned info = S.info[j];
        const unsigned mask = info & 0xffu;
        if (__popc(mask) <= 1) {
          float4* dst = reinterpret_cast<float4*>(dup_grad + (size_t)slot * 12);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
          *reinterpret_cast<float2*>(dst + 2) = make_float2(v[8], v[9]);
          atomicOr(&S.wrote[j >> 5], 1u << (j & 31));
        } else {
          float* ex = S.exch[(info >> 8) + __popc(mask & ((1u << wid) - 1u))];
#pragma unroll
          for (int k = 0; k < 10; ++k) ex[k] = v[k];
        }
      }
      if (bym) break;                      // the rest of this batch belongs to a later window: revisit it there
      ++bi;
    }
    __syncthreads();
    // ---- window epilogue: multi-region records: add the regions' sums in region order; untouched single-region
    // records: zeros
    for (int r = tid; r < wn; r += TILE_THREADS) {
      const unsigned info = S.info[r];
      const int nreg = __popc(info & 0xffu);
      float v[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) v[k] = 0.0f;
      if (nreg > 1) {
        const float* ex = S.exch[info >> 8];
        for (int o = 0; o < nreg; ++o)
#pragma unroll
          for (int k = 0; k < 10; ++k) v[k] += ex[o * 10 + k];
      } else if ((S.wrote[r >> 5] >> (r & 31)) & 1u) {
        continue;
      }
      const int slot = __float_as_int(reinterpret_cast<const float*>(&S.rec[3 * r + 2])[2]);
      float4* dst = reinterpret_cast<float4*>(dup_grad + (size_t)slot * 12);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      *reinterpret_cast<float2*>(dst + 2) = make_float2(v[8], v[9]);
    }
    if (w0 + LOG_W < L) __syncthreads();     // the next window re-initialises the shared tables
  }
}

cudaError_t launch_blend_backward(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st,
                                  const SpfRasterGradOut& gout, const SpfRasterGradIn& gin, cudaStream_t s) {
  const int grid = d.B * d.T;
  if (d.flags & SPF_FLAG_BWD_V1) {
    if (d.flags & SPF_FLAG_NO_TMA)
      pdl_launch(blend_backward_v1_kernel<false>, grid, TILE_THREADS, 0, s)(d, in.bg, st, gout, gin.dup_grad);
    else
      pdl_launch(blend_backward_v1_kernel<true>, grid, TILE_THREADS, 0, s)(d, in.bg, st, gout, gin.dup_grad);
    return cudaGetLastError();
  }
  const bool use_log = st.pair_log != nullptr && st.pair_count != nullptr && d.pair_cap > 0;
  if (use_log) {
    cudaError_t e0 = cudaFuncSetAttribute(blend_backward_log_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(LogSmem));
    if (e0 != cudaSuccess) return e0;
    pdl_launch(blend_backward_log_kernel, grid, TILE_THREADS, sizeof(LogSmem), s)(d, in.bg, st, gout, gin.dup_grad);
    e0 = cudaGetLastError();
    if (e0 != cudaSuccess) return e0;
  }
  const size_t smem = sizeof(BwdSmem);
  const int grid2 = use_log ? min(grid, 148 * 3) : grid;   // fallback-only launch: one wave of resident CTAs
  cudaError_t e;
  if (d.flags & SPF_FLAG_NO_TMA) {
    e = cudaFuncSetAttribute(blend_backward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pdl_launch(blend_backward_kernel<false>, grid2, TILE_THREADS, smem, s)(d, in.bg, st, gout, gin.dup_grad, use_log ? 1 : 0);
  } else {
    e = cudaFuncSetAttribute(blend_backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pdl_launch(blend_backward_kernel<true>, grid2, TILE_THREADS, smem, s)(d, in.bg, st, gout, gin.dup_grad, use_log ? 1 : 0);
  }
  return cudaGetLastError();
}

}  // namespace spf
