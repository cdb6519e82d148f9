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

// Integer pixel rectangle of a record's alpha box inside a warp region [x0,x1]x[y0,y1] (pixel centres are the integer
// coordinates; the region bounds are already clipped to the image).  The boxes are conservative, so a record whose
// rectangle is empty cannot contribute to any pixel of the region.  Forward and backward share this test: the
// backward's per-record region masks must cover every region whose warp logged a pair of the record.
__device__ __forceinline__ bool region_rect(const float4& bb, float x0, float x1, float y0, float y1, float& fx0,
                                            float& fx1, float& fy0, float& fy1) {
  fx0 = fmaxf(ceilf(bb.x), x0);
  fx1 = fminf(floorf(bb.y), x1);
  fy0 = fmaxf(ceilf(bb.z), y0);
  fy1 = fminf(floorf(bb.w), y1);
  return (fx0 <= fx1) && (fy0 <= fy1);
}

// ---------------------------------------------------------------------------------------------------
// Blend forward, v5: per-region HIT LISTS, PAIR-PARALLEL alpha evaluation, per-pixel queues for the compositing.
//
// With SPFSplatV2's splat sizes a warp region (8x4 pixels) is touched by ~57 records of a tile's ~370, and a touched
// record covers ~5.6 of the region's 32 pixel centres, ~4.3 of which pass the alpha test (scripts/pair_stats.py).  The
// first-generation kernel spent one 32-lane iteration per touched record at 4/32 useful lanes, and every warp culled
// the whole tile list by itself.  Here the tile list is processed in passes of PASS records (one TMA bulk copy):
//
//   cull (whole block, one record per thread): the integer pixel rectangle of the record's alpha box inside the tile
//       is intersected with the eight 8x4 warp regions; every non-empty intersection becomes one HIT entry
//       {record, first pixel lane, width-1, height-1} appended -- in list order (ballots + a prefix over the warps) --
//       to that region's hit list in shared memory.  Every box is read once per tile, straight from global memory.
//   phase A (per warp, lane = candidate (record, pixel) pair): 32 hits at a time live one per lane in registers; a
//       warp scan of their pixel counts and a bit mask of run heads (redux.or) let lane l of a batch find the hit
//       (one shuffle) and the pixel (a 256-byte lookup table) of candidate number 32*batch + l.  Each lane evaluates
//       ONE alpha; contributing pairs are appended, in (record, pixel) order, to the queue of their pixel (match.any
//       ranks same-pixel pairs of a batch) and, if a backward follows, to the warp's pair log.
//   phase B (lane = pixel): every pixel composites its queue front to back (T, colour, depth, early stop, last
//       contributor).  Queues are drained when one fills up and at the end of a pass.
//
// A group of hits that covers most of the region (large splats: mean candidates per hit >= DENSE_MIN) takes the DENSE
// path instead -- one record per iteration, every lane evaluates its own pixel and composites at once -- after the
// queues have been drained, so the per-pixel order is always the list order; so does a region whose hit list
// overflows.  The arithmetic per (pixel, record) is the same in all paths and the accept / reject decisions are the
// per-pixel tests of eval_alpha, so images, final_T and n_contrib do not depend on the path taken.
//
// Pair log (LOG = true): one 8-byte entry per contributing pair, dense, in (record, pixel-lane) order:
//   x = record index in the tile list | pixel lane << 25,  y = G = exp(power).
// Pairs logged for a pixel after it stopped (possible only between two drains) carry a record index >= that pixel's
// n_contrib; the backward ignores them.  A warp's segment holds `pair_capacity` entries; a warp that needs more stops
// writing and reports -1 (its tile is then handled by the recomputing backward).  control[2] = largest count needed.
constexpr unsigned PAIR_J_MASK = (1u << 25) - 1u;
constexpr int QD = 8;            // queue slots per pixel
constexpr int DENSE_MIN = 12;    // mean candidates per hit from which a group of hits takes the dense path
constexpr int PASS = 512;        // records per pass
constexpr int HCAP = 256;        // hit entries per region and pass
constexpr unsigned FULL = 0xffffffffu;

struct FwdSmem {
  float4 buf[PASS * 3];          // the pass's slab records
  uint32_t hit[8][HCAP];         // per region: record (9) | first pixel lane (5) << 9 | width-1 (3) << 14 | height-1 (2) << 17
  uint2 queue[8][QD * 32];       // per warp: [slot][pixel] = {alpha, record index in the tile list}
  int qcnt[8][32];               // per warp: queued pairs per pixel
  float2 pixf[8][32];            // per warp: pixel coordinates of the region's 32 pixel lanes
  int wcnt[8][8], woff[8][8];    // [region][warp]: hits found by a warp in the current cull round / their offsets
  unsigned bal[8][8];            // [region][warp]: the warp's hit ballot of the current cull round
  int hbase[8];                  // per region: hits so far in this pass
  uint8_t rowcol[8][32];         // [width-1][k] -> lane offset (row * 8 + column) of candidate k of a hit
  uint64_t bar;
  int blk_pairs;
};

template <bool TMA, bool LOG>
__global__ void __launch_bounds__(TILE_THREADS, 4)
blend_forward_kernel(Dims d, const float* __restrict__ bg_all, SpfRasterState st, SpfRasterOut out) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char fwd_smem_raw[];
  FwdSmem& S = *reinterpret_cast<FwdSmem*>(fwd_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t = blockIdx.x;
  const int view = t / d.T, tile = t - view * d.T;
  const int s = st.tile_ranges[2 * (size_t)t], e = st.tile_ranges[2 * (size_t)t + 1];
  const int L = e - s;
  int bx, by;
  warp_block_of_thread(tid, tile, d.gx, bx, by);
  const int px = bx + (lane & 7), py = by + (lane >> 3);
  const bool inside = (px < d.W) && (py < d.H);
  const float pxf = (float)px, pyf = (float)py;
  const int tx0 = (tile % d.gx) * TILE, ty0 = (tile / d.gx) * TILE;
  // the tile and the warp's region, clipped to the image
  const float tx0f = (float)tx0, tx1f = (float)min(tx0 + TILE - 1, d.W - 1), ty0f = (float)ty0, ty1f = (float)min(ty0 + TILE - 1, d.H - 1);
  const float wx0 = (float)bx, wx1 = (float)min(bx + 7, d.W - 1), wy0 = (float)by, wy1 = (float)min(by + 3, d.H - 1);
  uint2* q = S.queue[wid];
  int* qc = S.qcnt[wid];
  const uint32_t* hl = S.hit[wid];
  if (tid == 0) {
    if (TMA) { mbar_init(&S.bar, 1); mbar_fence_init(); }
    S.blk_pairs = 0;
  }
  qc[lane] = 0;
  S.pixf[wid][lane] = make_float2(pxf, pyf);
  {
    const int w1 = wid, k = lane, w = w1 + 1;          // 256 threads = 8 widths x 32 candidates
    S.rowcol[w1][k] = (uint8_t)(k < 4 * w ? (k / w) * 8 + (k % w) : 0);
  }
  __syncthreads();
  const float4* slab = reinterpret_cast<const float4*>(st.slab) + 3 * (size_t)s;
  const float4* cull = reinterpret_cast<const float4*>(st.cullbox) + (size_t)s;
  const int Cw = d.pair_cap;
  uint2* plog = LOG ? reinterpret_cast<uint2*>(st.pair_log) + ((size_t)t * 8 + wid) * (size_t)Cw : nullptr;
  const unsigned lt_mask = (1u << lane) - 1u;
  int npairs = 0;    // pairs this warp has produced for the log (keeps counting past the capacity)

  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
  int last = 0;
  bool done = !inside;
  unsigned donemask = __ballot_sync(FULL, done);
  bool queued = false;   // warp-uniform: some queue is non-empty
  int pass0 = 0;         // first record of the current pass (the queues hold tile-list indices)

  // phase B: every pixel composites its queued pairs in list order
  auto drain = [&]() {
    const int n = qc[lane];
    const int maxn = __reduce_max_sync(FULL, n);
    for (int i = 0; i < maxn; ++i) {
      const uint2 en = q[i * 32 + lane];
      const float alpha = __uint_as_float(en.x);
      const float test_T = T * (1.0f - alpha);
      const bool act = (i < n) && !done;
      if (act && test_T < T_STOP) done = true;
      if (act && !done) {
        const float4* r = S.buf + 3u * (en.y - (unsigned)pass0);
        const float2 c01 = *reinterpret_cast<const float2*>(&r[1].z);
        const float2 c2d = *reinterpret_cast<const float2*>(&r[2].x);
        const float w = alpha * T;
        C0 += c01.x * w; C1 += c01.y * w; C2 += c2d.x * w; D += c2d.y * w;
        T = test_T;
        last = (int)en.y + 1;
      }
    }
    qc[lane] = 0;
    queued = false;
    donemask = __ballot_sync(FULL, done);
    __syncwarp();
  };
  // dense step: every lane evaluates its own pixel against record jr of the pass and composites at once
  auto dense_step = [&](int jr, int p0) {
    const float4 a = S.buf[3 * jr], b = S.buf[3 * jr + 1];
    float dx, dy, G, alpha;
    bool ok = eval_alpha(a, b, pxf, pyf, dx, dy, G, alpha) && !done;
    const float test_T = T * (1.0f - alpha);
    if (ok && test_T < T_STOP) { done = true; ok = false; }
    if (ok) {
      const float4 cc = S.buf[3 * jr + 2];
      const float w = alpha * T;
      C0 += b.z * w; C1 += b.w * w; C2 += cc.x * w; D += cc.y * w;
      T = test_T;
      last = p0 + jr + 1;
    }
    if (LOG) {
      const unsigned cb = __ballot_sync(FULL, ok);
      if (cb) {
        const int off = npairs + __popc(cb & lt_mask);
        if (ok && off < Cw) plog[off] = make_uint2((unsigned)(p0 + jr) | ((unsigned)lane << 25), __float_as_uint(G));
        npairs += __popc(cb);
      }
    }
  };

  const int npass = (L + PASS - 1) / PASS;
  for (int p = 0; p < npass; ++p) {
    const int p0 = p * PASS, pc = min(PASS, L - p0);
    pass0 = p0;
    // ---- stage the pass's records (one bulk copy) while the block culls
    if (TMA) {
      if (tid == 0) {
        mbar_expect_tx(&S.bar, (uint32_t)pc * 48u);
        tma_load_1d(S.buf, slab + 3 * (size_t)p0, (uint32_t)pc * 48u, &S.bar);
      }
    } else {
      for (int i = tid; i < pc * 3; i += TILE_THREADS) S.buf[i] = slab[3 * (size_t)p0 + i];
    }
    // ---- cull: one record per thread, eight ordered per-region compactions (thread (w, wp) of the first 64 keeps
    //      region w's running hit count and hands warp wp its offset)
    int hrun = 0;
    for (int r0 = 0; r0 < pc; r0 += TILE_THREADS) {
      const int i = r0 + tid;
      int X0 = 1, X1 = 0, Y0 = 1, Y1 = 0;      // tile-local pixel rectangle of the record's alpha box (empty)
      if (i < pc) {
        const float4 bb = __ldg(cull + p0 + i);
        float f0, f1, f2, f3;
        if (region_rect(bb, tx0f, tx1f, ty0f, ty1f, f0, f1, f2, f3)) {
          X0 = (int)f0 - tx0; X1 = (int)f1 - tx0; Y0 = (int)f2 - ty0; Y1 = (int)f3 - ty0;
        }
      }
      // regions touched: columns X0>>3 .. X1>>3 (of 2), bands Y0>>2 .. Y1>>2 (of 4); region w = band * 2 + column
      unsigned mask = 0u;
      if (X0 <= X1) {
        const unsigned cols = (X0 < 8 ? 1u : 0u) | (X1 >= 8 ? 2u : 0u);
        const unsigned bands = ((2u << (Y1 >> 2)) - 1u) & ~((1u << (Y0 >> 2)) - 1u);
        mask = ((bands & 1u) ? cols : 0u) | ((bands & 2u) ? cols << 2 : 0u) | ((bands & 4u) ? cols << 4 : 0u) |
               ((bands & 8u) ? cols << 6 : 0u);
      }
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const unsigned bal = __ballot_sync(FULL, (mask >> w) & 1u);
        if (lane == w) { S.bal[w][wid] = bal; S.wcnt[w][wid] = __popc(bal); }
      }
      __syncthreads();
      if (tid < 64) {
        const int w = tid >> 3, wp = tid & 7;
        int off = hrun;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c = S.wcnt[w][k];
          if (k < wp) off += c;
          hrun += c;
        }
        S.woff[w][wp] = off;
      }
      __syncthreads();
      for (unsigned m = mask; m; m &= m - 1) {      // 1.2 regions per record on average
        const int w = __ffs(m) - 1;
        const int hx = (w & 1) * 8, hy = (w >> 1) * 4;
        const int xa = max(X0, hx), xb = min(X1, hx + 7), ya = max(Y0, hy), yb = min(Y1, hy + 3);
        const int pos = S.woff[w][wid] + __popc(S.bal[w][wid] & lt_mask);
        if (pos < HCAP)
          S.hit[w][pos] = (unsigned)i | ((unsigned)((ya - hy) * 8 + (xa - hx)) << 9) | ((unsigned)(xb - xa) << 14) |
                          ((unsigned)(yb - ya) << 17);
      }
      __syncwarp();       // every lane has read this round's ballots before lane w overwrites them in the next round
    }
    if (tid < 64 && (tid & 7) == 0) S.hbase[tid >> 3] = hrun;
    __syncthreads();               // hit lists and counts complete
    if (TMA) mbar_wait(&S.bar, (uint32_t)(p & 1));
    const int nh = S.hbase[wid];

    if (donemask != FULL && nh > HCAP) {
      // ---- the region's hit list overflowed (large splats): first-generation loop over the pass's records
      if (queued) drain();
      for (int g0 = 0; g0 < pc && donemask != FULL; g0 += 32) {
        const int r = g0 + lane;
        bool hit = false;
        if (r < pc) {
          float f0, f1, f2, f3;
          hit = region_rect(__ldg(cull + p0 + r), wx0, wx1, wy0, wy1, f0, f1, f2, f3);
        }
        unsigned m = __ballot_sync(FULL, hit);
        while (m) {
          const int jr = g0 + __ffs(m) - 1;
          m &= m - 1;
          dense_step(jr, p0);
        }
        donemask = __ballot_sync(FULL, done);
      }
    } else if (donemask != FULL) {
      for (int h0 = 0; h0 < nh && donemask != FULL; h0 += 32) {
        // ---- 32 hits, one per lane
        const int cnth = min(32, nh - h0);
        unsigned he = 0u;
        int n = 0;
        if (lane < cnth) {
          he = hl[h0 + lane];
          n = (int)(((he >> 14) & 7u) + 1u) * (int)(((he >> 17) & 3u) + 1u);
        }
        const int incl = warp_incl_scan_i(n, lane);
        const int total = __shfl_sync(FULL, incl, 31);
        const int excl = incl - n;
        if (total >= DENSE_MIN * cnth) {
          // ---- dense path: one hit per iteration, lane = its own pixel
          if (queued) drain();
          for (int b = 0; b < cnth; ++b) dense_step((int)(__shfl_sync(FULL, he, b) & 511u), p0);
          donemask = __ballot_sync(FULL, done);
          continue;
        }
        // ---- sparse path: lane = candidate pair
        he |= (unsigned)excl << 19;
        int nbefore = 0;   // hits that start before the current batch
        for (int b0 = 0; b0 < total; b0 += 32) {
          const unsigned rel = (unsigned)(excl - b0);
          const unsigned heads = __reduce_or_sync(FULL, (n > 0 && rel < 32u) ? (1u << rel) : 0u);
          const int qi = b0 + lane;
          const bool valid = qi < total;
          const int rr = nbefore + __popc(heads & (FULL >> (31 - lane))) - 1;
          nbefore += __popc(heads);
          const unsigned h = __shfl_sync(FULL, he, rr & 31);
          const int jr = (int)(h & 511u);
          const int k = (qi - (int)(h >> 19)) & 31;
          const int pl = ((int)((h >> 9) & 31u) + (int)S.rowcol[(h >> 14) & 7u][k]) & 31;   // lane that owns the pixel
          const float2 pf = S.pixf[wid][pl];
          const float4 a = S.buf[3 * jr], b = S.buf[3 * jr + 1];
          float dx, dy, G, alpha;
          const bool ok = eval_alpha(a, b, pf.x, pf.y, dx, dy, G, alpha) && valid && !((donemask >> pl) & 1u);
          const unsigned cb = __ballot_sync(FULL, ok);
          if (cb == 0u) continue;
          unsigned peers = 0u;
          int slot = 0, tot = 0;
          if (ok) {
            peers = __match_any_sync(cb, pl);
            const int cntp = qc[pl];
            slot = cntp + __popc(peers & lt_mask);
            tot = cntp + __popc(peers);
          }
          if (__ballot_sync(FULL, ok && slot >= QD)) {     // a queue would overflow: composite what is queued first
            drain();
            slot = __popc(peers & lt_mask);
            tot = __popc(peers);
          }
          for (;;) {
            if (ok && slot >= 0 && slot < QD) q[slot * 32 + pl] = make_uint2(__float_as_uint(alpha), (unsigned)(p0 + jr));
            if (ok && (peers >> lane) == 1u && tot > 0) qc[pl] = min(tot, QD);    // the pixel's last pair of the batch
            const unsigned more = __ballot_sync(FULL, ok && slot >= QD);
            __syncwarp();
            queued = true;
            if (more == 0u) break;
            drain();                                       // > QD pairs of one pixel in one batch: go round again
            slot -= QD;
            tot -= QD;
          }
          if (LOG) {
            const int off = npairs + __popc(cb & lt_mask);
            if (ok && off < Cw) plog[off] = make_uint2((unsigned)(p0 + jr) | ((unsigned)pl << 25), __float_as_uint(G));
            npairs += __popc(cb);
          }
        }
      }
    }
    if (queued) drain();           // the queues point into the staged records, which the next pass overwrites
    if (p + 1 < npass && __syncthreads_and(donemask == FULL)) break;
  }
  if (LOG) {
    if (lane == 0) {
      const bool okw = (npairs <= Cw && L <= (int)PAIR_J_MASK);
      st.pair_count[(size_t)t * 8 + wid] = okw ? npairs : -1;
      if (!okw) st.control[3] = 1;
      atomicMax(&S.blk_pairs, npairs);
    }
    __syncthreads();
    if (tid == 0 && S.blk_pairs > 0) atomicMax(st.control + 2, S.blk_pairs);
  }

  if (inside) {
    const float* bg = bg_all + view * 3;
    const size_t hw = (size_t)d.H * d.W;
    const size_t pix = (size_t)py * d.W + px;
    float* col = out.color + (size_t)view * 3 * hw;
    // Duplicate buffer overflow (control[1]: more (Gaussian, tile) duplicates than dup_capacity; the excess was dropped):
    // the image would silently miss Gaussians, so it is poisoned instead -- a loss computed from it is NaN, never a
    // plausible wrong number -- until the caller has re-run the forward with the capacity control[0] asks for.
    const float poison = st.control[1] != 0 ? __int_as_float(0x7fc00000) : 0.0f;
    col[pix] = C0 + T * bg[0] + poison;
    col[hw + pix] = C1 + T * bg[1] + poison;
    col[2 * hw + pix] = C2 + T * bg[2] + poison;
    out.depth[(size_t)view * hw + pix] = D + poison;
    if (out.alpha) out.alpha[(size_t)view * hw + pix] = 1.0f - T + poison;
    st.final_T[(size_t)view * hw + pix] = T;
    st.n_contrib[(size_t)view * hw + pix] = last;
    // private copy of the blended sums (no background term): backward derives every pixel's total
    // sum_j q_j w_j from it, independent of what the caller does with its output tensors afterwards
    reinterpret_cast<float4*>(st.accum)[(size_t)view * hw + pix] = make_float4(C0, C1, C2, D);
  }
}

template <bool TMA, bool LOG>
static cudaError_t launch_blend_forward_t(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st,
                                          const SpfRasterOut& out, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(blend_forward_kernel<TMA, LOG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(FwdSmem));
  if (e != cudaSuccess) return e;
  pdl_launch(blend_forward_kernel<TMA, LOG>, d.B * d.T, TILE_THREADS, sizeof(FwdSmem), s)(d, in.bg, st, out);
  return cudaGetLastError();
}

cudaError_t launch_blend_forward(const Dims& d, const SpfRasterIn& in, const SpfRasterState& st,
                                 const SpfRasterOut& out, cudaStream_t s) {
  const bool log = st.pair_log != nullptr && st.pair_count != nullptr && d.pair_cap > 0;
  if (d.flags & SPF_FLAG_NO_TMA)
    return log ? launch_blend_forward_t<false, true>(d, in, st, out, s) : launch_blend_forward_t<false, false>(d, in, st, out, s);
  return log ? launch_blend_forward_t<true, true>(d, in, st, out, s) : launch_blend_forward_t<true, false>(d, in, st, out, s);
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
    if ((int64_t)slot < d.cap) dup_grad[(size_t)slot * 12 + k] = 0.0f;
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
      if ((int64_t)slot < d.cap) dup_grad[(size_t)slot * 12 + k] = sum;
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
                                                    BwdSmem& S, const int t, uint32_t& ph0, uint32_t& ph1) {
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

  if (tid == 0) S.max_contrib = 0;
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
    if ((int64_t)slot < d.cap) dup_grad[(size_t)slot * 12 + k] = 0.0f;
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
    if (TMA) {   // the two mbarriers live for the whole (persistent) kernel: every thread tracks their phase parity
      uint32_t& ph = (c & 1) ? ph1 : ph0;
      mbar_wait(&S.bar[c & 1], ph & 1u);
      ++ph;
    } else {
      __syncthreads();
    }
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
      if ((int64_t)slot < d.cap) dup_grad[(size_t)slot * 12 + k] = sum;
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
  if (TMA && threadIdx.x == 0) { mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1); mbar_fence_init(); }
  uint32_t ph0 = 0u, ph1 = 0u;     // completed phases of the two mbarriers (initialised once, reused by every tile)
  __syncthreads();
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    blend_backward_tile<TMA>(d, bg_all, st, go, dup_grad, use_log, S, t, ph0, ph1);
    __syncthreads();     // shared memory is reused by the next tile
  }
}

// ---------------------------------------------------------------------------------------------------
// Blend backward, v4: consumes the forward's 8-byte pair log -- no alpha tests, no per-pixel sequential pass.
//
// For a logged pair i of pixel p:  dL/dalpha_i = T_i q_i - (Qtot_p - P_i) / (1 - alpha_i),  q_i = g_p . (rgb_i, depth_i),
// g_p = (dL/dC, dL/dD), P_i = sum_{j<=i} q_j alpha_j T_j and Qtot_p = g_p . S_final + T_final (bg . dL/dC - dL/dA).
// Warp w walks the log of forward-warp w 32 pairs at a time, ONE PAIR PER LANE.  T_i and P_i are not logged: they are
// rebuilt by a KEYED SCAN -- the running (T, P) of the warp's 32 pixels live in shared memory; within a batch the pairs
// of one pixel (match.any on the pixel lane; usually one, seldom more than three) are chained in log order, each taking
// (T, P) from its predecessor with two shuffles; the last one writes the pixel's state back.  T_i is recomputed with the
// forward's own operation (T <- T * (1 - alpha)) and is therefore bit-identical to the T the forward blended with.
//
// The pairs of a record (a "run": adjacent log entries, in pixel-lane order) are summed by a segmented shuffle
// reduction; a run cut by a batch boundary parks its first part in shared memory and the next batch's lane 0 adds it.
// A record whose alpha box overlaps a single 8x4 warp region (the common case with SPFSplatV2's ~2 px splats) is final
// after that and its 10 sums are stored straight to its duplicate slot; one that overlaps several regions parks its
// per-warp sums in exchange slots that are added in fixed region order after ONE block barrier.  No atomics on floats,
// fixed order: bit-reproducible.  Long tile lists are processed in windows of LOG_W records (each warp's log is sorted
// by record).  Tiles with an incomplete log or too many multi-region records in a window are left to the recomputing
// kernel above (flagged through pair_count).
constexpr int LOG_W = 256;         // records per window of the tile list (staged in shared memory)
constexpr int LOG_ESLOTS = 320;    // exchange slots per window (one per (multi-region record, overlapped region))

struct LogSmem {
  float4 rec[LOG_W * 3];                 // the window's slab records
  float4 pg[TILE_THREADS];               // per pixel: dL/dC (3), dL/dD
  float pq[TILE_THREADS];                // per pixel: Qtot
  float Ts[TILE_THREADS], Ps[TILE_THREADS];   // per pixel: running transmittance and prefix P of the keyed scan
  int nc[TILE_THREADS];                  // per pixel: n_contrib (pairs of later records are dead)
  uint32_t info[LOG_W];                  // region mask (8 bits) | first exchange slot << 8
  float exch[LOG_ESLOTS][10];
  float carry[8][12];                    // per warp: sums of a run cut by the batch boundary
  uint32_t wrote[LOG_W / 32];
  int base;
};

__global__ void __launch_bounds__(TILE_THREADS, 6)
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
  const uint2* lp = reinterpret_cast<const uint2*>(st.pair_log) + ((size_t)t * 8 + wid) * (size_t)d.pair_cap;
  const unsigned lt_mask = (1u << lane) - 1u;

  // the log is read in fixed 32-entry batches whose addresses do not depend on data: each batch is prefetched towards
  // the SM three batches ahead and then loaded where it is used
  const int nbatch = (count + 31) >> 5;
  for (int k = 0; k < 3; ++k)
    if ((k << 5) + lane < count) prefetch_l1(lp + (size_t)((k << 5) + lane));
  int bi = 0;
  int carry_j = -1;          // warp-uniform: record whose first part is parked in S.carry[wid]

  // per-pixel upstream gradients, Qtot and scan state
  {
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, gd = 0.f, qtot = 0.f;
    int ncontrib = 0;
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
      ncontrib = st.n_contrib[(size_t)view * hw + pix];
    }
    S.pg[tid] = make_float4(g0, g1, g2, gd);
    S.pq[tid] = qtot;
    S.Ts[tid] = 1.0f;
    S.Ps[tid] = 0.0f;
    S.nc[tid] = ncontrib;
  }
  const float4* pgw = S.pg + wid * 32;
  const float* pqw = S.pq + wid * 32;
  float* Tw = S.Ts + wid * 32;
  float* Pw = S.Ps + wid * 32;
  const int* ncw = S.nc + wid * 32;
  float* carry = S.carry[wid];

  for (int w0 = 0; w0 < L; w0 += LOG_W) {
    const int wn = min(LOG_W, L - w0), wend = w0 + wn;
    // ---- window prologue: stage the records, region masks (the same rectangle test the forward used), exchange slots
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
        float f0, f1, f2, f3;
        if (region_rect(bb, (float)tx0, (float)min(tx0 + TILE - 1, d.W - 1), (float)ty0, (float)min(ty0 + TILE - 1, d.H - 1), f0, f1,
                        f2, f3)) {
          // tile-local pixel rectangle, intersected with the eight regions in integers (as the forward's cull does)
          const int X0 = (int)f0 - tx0, X1 = (int)f1 - tx0, Y0 = (int)f2 - ty0, Y1 = (int)f3 - ty0;
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const int hx = (w & 1) * 8, hy = (w >> 1) * 4;
            if ((max(X0, hx) <= min(X1, hx + 7)) && (max(Y0, hy) <= min(Y1, hy + 3))) mask |= 1u << w;
          }
        }
      }
      const int nreg = __popc(mask);
      const int need = nreg > 1 ? nreg : 0;
      const int inc = warp_incl_scan_i(need, lane);
      const int tot = __shfl_sync(FULL, inc, 31);
      int wbase = 0;
      if (lane == 0 && tot > 0) wbase = atomicAdd(&S.base, tot);   // slot POSITIONS may vary run to run; sums do not
      wbase = __shfl_sync(FULL, wbase, 0);
      if (r < wn) S.info[r] = mask | ((unsigned)(wbase + inc - need) << 8);
    }
    __syncthreads();
    if (S.base > LOG_ESLOTS) {               // too many multi-region records: leave the tile to the recomputing kernel
      if (tid == 0) { st.pair_count[(size_t)t * 8] = -1; st.control[3] = 1; }
      return;
    }

    // a finished run (this warp's sums for window record jl): straight to the duplicate's slot if the record touches
    // no other region, else into this region's exchange slot
    auto emit_run = [&](int jl, const float* v) {
      const unsigned info = S.info[jl];
      const unsigned mask = info & 0xffu;
      if (__popc(mask) <= 1) {
        const int slot = __float_as_int(S.rec[3 * jl + 2].z);
        if ((int64_t)slot < d.cap) {
          float4* dst = reinterpret_cast<float4*>(dup_grad + (size_t)slot * 12);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
          *reinterpret_cast<float2*>(dst + 2) = make_float2(v[8], v[9]);
        }
        atomicOr(&S.wrote[jl >> 5], 1u << (jl & 31));
      } else {
        float* ex = S.exch[(info >> 8) + __popc(mask & ((1u << wid) - 1u))];
#pragma unroll
        for (int k = 0; k < 10; ++k) ex[k] = v[k];
      }
    };

    // ---- this warp's pairs whose record lies in the window
    while (bi < nbatch) {
      const int idx = (bi << 5) + lane;
      if (idx + 96 < count) prefetch_l1(lp + (size_t)(idx + 96));
      const bool have = idx < count;
      uint2 en = make_uint2(0u, 0u);
      if (have) en = lp[idx];
      const int jraw = have ? (int)(en.x & PAIR_J_MASK) : 0x7fffffff;
      const int pl = (int)((en.x >> 25) & 31u);
      const bool inwin = have && (jraw >= w0) && (jraw < wend);
      const unsigned bym = __ballot_sync(FULL, have && (jraw >= wend));   // sorted by record: belongs to a later window
      const int j = inwin ? jraw - w0 : -1;                               // run id (dead pairs stay part of their run)
      const bool act = inwin && (jraw < ncw[pl]);
      const unsigned am = __ballot_sync(FULL, act);
      float v[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) v[k] = 0.0f;
      if (am) {
        // keyed scan: (T_before, P_after) of every live pair
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, cc = a, g = a;
        float G = 0.f, alpha = 0.f, qv = 0.f, c = 0.f, Tb = 1.0f, Pb = 0.0f;
        unsigned peers = 0u;
        int rank = 0, pp = lane;
        if (act) {
          G = __uint_as_float(en.y);
          a = S.rec[3 * j]; b = S.rec[3 * j + 1]; cc = S.rec[3 * j + 2];
          g = pgw[pl];
          alpha = fminf(ALPHA_MAX, b.y * G);
          qv = (b.z * g.x + b.w * g.y) + (cc.x * g.z + cc.y * g.w);
          c = qv * alpha;
          peers = __match_any_sync(am, pl);
          const unsigned below = peers & lt_mask;
          rank = __popc(below);
          if (below) pp = 31 - __clz(below);
          Tb = Tw[pl];
          Pb = Pw[pl];
        }
        const int maxrank = __reduce_max_sync(FULL, rank);
        float Ta = Tb * (1.0f - alpha), Pa = Pb + c * Tb;
        for (int sidx = 1; sidx <= maxrank; ++sidx) {
          const float Tn = __shfl_sync(FULL, Ta, pp), Pn = __shfl_sync(FULL, Pa, pp);
          if (rank == sidx) {
            Tb = Tn; Pb = Pn;
            Ta = Tb * (1.0f - alpha); Pa = Pb + c * Tb;
          }
        }
        if (act && (peers >> lane) == 1u) { Tw[pl] = Ta; Pw[pl] = Pa; }   // the pixel's last pair of the batch
        __syncwarp();
        if (act) {
          const float dx = a.x - (float)(bx + (pl & 7)), dy = a.y - (float)(by + (pl >> 3));
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
      }
      // segmented reduction over runs of equal record
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int jo = __shfl_down_sync(FULL, j, off);
        const bool same = (lane + off < 32) && (jo == j) && (j >= 0);
        if (!__any_sync(FULL, same)) break;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const float vo = __shfl_down_sync(FULL, v[k], off);
          if (same) v[k] += vo;
        }
      }
      const int jprev = __shfl_up_sync(FULL, j, 1);
      const bool head = (j >= 0) && (lane == 0 || jprev != j);
      // The LAST run of a batch may continue in the next batch, so it is never emitted at once: its sums are parked
      // (carry) and either merged into lane 0's run of the next batch (same record) or emitted from there (no load of
      // the next entry is needed to find out which).
      const int j0 = __shfl_sync(FULL, j, 0);
      if (carry_j >= 0) {
        if (j0 == carry_j) {                   // lane 0's run is the continuation
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 10; ++k) v[k] = carry[k] + v[k];
          }
        } else if (lane == 0) {                // the parked run was complete
          float cv[10];
#pragma unroll
          for (int k = 0; k < 10; ++k) cv[k] = carry[k];
          emit_run(carry_j, cv);
        }
        carry_j = -1;
        __syncwarp();
      }
      const unsigned hm = __ballot_sync(FULL, head);
      const int hlast = hm ? 31 - __clz(hm) : -1;
      if (head) {
        if (lane == hlast) {
#pragma unroll
          for (int k = 0; k < 10; ++k) carry[k] = v[k];
        } else {
          emit_run(j, v);
        }
      }
      if (hm) carry_j = __shfl_sync(FULL, j, hlast);
      __syncwarp();
      if (bym) break;                      // the rest of this batch belongs to a later window: revisit it there
      ++bi;
    }
    if (carry_j >= 0) {                    // the window's (or the log's) last run
      if (lane == 0) {
        float cv[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) cv[k] = carry[k];
        emit_run(carry_j, cv);
      }
      carry_j = -1;
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
      if ((int64_t)slot >= d.cap) continue;
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
  const int grid2 = use_log ? min(grid, sm_count() * 3) : grid;   // fallback-only launch: one wave of resident CTAs
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
