// RoIAlign / RoIAlignAvg / RoIAlignMax forward + backward for sm_100a.
//
// Semantics follow the reference's legacy kernel exactly in geometry
// (lib/model/roi_align/src/roi_align_kernel.cu:27-68 fwd, :99-141 bwd) and fuse the 2x2
// stride-1 avg / max pooling that RoIAlignAvg / RoIAlignMax run afterwards
// (lib/model/roi_align/modules/roi_align.py:26-29, 39-42).
//
// B200 design (not the reference's one-thread-per-output gather):
//   1. k_roi_plan: one tiny launch turns every roi into a 128-byte "plan": the separable
//      sample grid (row index + row ratio per ph, column index + column ratio per pw),
//      computed once with the reference's exact fp32/fp64 expression order, plus per-image
//      roi lists.  The hot kernels never touch roi coordinates again.
//   2. forward (k_align8_fwd_walk): a CTA owns 4 consecutive channel planes of ONE image,
//      pulls them HBM -> shared memory once (async copies, interleaved per pixel) and serves
//      EVERY roi of that image from shared memory: 4 rois per warp, 4 channels per LDS.128,
//      per-roi orientation + tap order chosen so that the requests are bank-conflict-free,
//      2x2 pooling fused, each (roi, 4 channels) result leaves as ONE 784-byte bulk async
//      store (TMA engine).  The feature map is read from HBM exactly once and no 8x8
//      intermediate tensor exists.
//   3. backward: the transpose -- gradient planes accumulated in shared memory, grad_out tiles by
//      bulk async load, grad_in written once, coalesced; no global atomics, no memset.  The
//      shipped kernel is k_align8_bwd_own (roi_align_bwd.cu: whole-roi warps, merged taps, a
//      token ring between the warps of a CTA); k_align8_bwd_walk below (per-row spin locks) is
//      round 1's kernel, kept behind RLOD_BWD_V1=1 for A/B runs.
//   Generic kernels (any grid size / channel count / plane size, and RoIAlignMax backward)
//   cover everything the fast paths do not.
#include <cstdlib>
#include <type_traits>

#include "roi_lists.cuh"

#ifndef RLOD_ABL
#define RLOD_ABL 0  // ablation switch for profiling experiments only (tools/ablate.sh)
#endif

namespace rlod {

// ----------------------------------------------------------------------------------------
// plan kernel.  One thread per (roi, slot): slot < GH is a sample row, else a sample column.
// Expression order mirrors what nvcc emits for the reference kernel (roi_align_kernel.cu:
// 33-49): start = coord*scale (FMUL); size = max(fma(end,scale,-start) + 1, 0); bin =
// (float)((double)size / (G - 1.)); pos = fma(p, bin, start); idx = min(floor(pos), dim-2);
// ratio = pos - idx; a sample is zero when pos < 0 || pos >= dim.
// plan words per roi: [hs: GH int][hr: GH float][(ws, wr): GW pairs]; index -1 = invalid.
// ----------------------------------------------------------------------------------------
__global__ void k_roi_plan(const float *__restrict__ rois, int R, int B, int H, int W, int GH,
                           int GW, float scale, AlignWs ws) {
  const int slots = GH + GW;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)R * slots) return;
  const int r = (int)(t / slots), slot = (int)(t - (long long)r * slots);
  const float *roi = rois + (size_t)r * 5;
  const float bf = roi[0];
  const int bi = (int)bf;
  const bool bvalid = (bf >= 0.f) && (bi < B);
  const int b = bvalid ? bi : 0;
  const bool is_row = slot < GH;
  const int p = is_row ? slot : slot - GH;
  const int G = is_row ? GH : GW;
  const int dim = is_row ? H : W;
  const float c0 = is_row ? roi[2] : roi[1];
  const float c1 = is_row ? roi[4] : roi[3];
  const float start = __fmul_rn(c0, scale);
  const float size = fmaxf(__fadd_rn(__fmaf_rn(c1, scale, -start), 1.f), 0.f);
  const float bin = (float)((double)size / ((double)G - 1.));
  const float pos = __fmaf_rn((float)p, bin, start);
  const bool valid = bvalid && (pos >= 0.f) && (pos < (float)dim);
  int idx = -1;
  float ratio = 0.f;
  if (valid) {
    idx = (int)fminf(floorf(pos), (float)(dim - 2));
    ratio = __fsub_rn(pos, (float)idx);
  }
  int *pl = ws.plan + (size_t)r * (2 * GH + 2 * GW);
  if (is_row) {
    pl[p] = idx;
    pl[GH + p] = __float_as_int(ratio);
  } else {
    pl[2 * GH + 2 * p] = idx;
    pl[2 * GH + 2 * p + 1] = __float_as_int(ratio);
  }
  if (slot == 0) roi_list_mark(rois, r, R, B, b, ws);
}

// ----------------------------------------------------------------------------------------
// shared arithmetic: one bilinear sample from a plane, separable form
//   s = (1-hr) * ((1-wr)*p00 + wr*p01) + hr * ((1-wr)*p10 + wr*p11)
// (the reference evaluates the four products in fp64 and rounds once,
// roi_align_kernel.cu:64-67; fp32 FMA stays within a few ulp of that, tolerance 1e-5)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float lerp_row(const float *__restrict__ p, float wr) {
  return fmaf(wr, p[1], (1.f - wr) * p[0]);
}

__device__ __forceinline__ float sample_plane(const float *__restrict__ plane, int W, int hs,
                                              float hr, int ws, float wr) {
  if (hs < 0 || ws < 0) return 0.f;
  const float *p = plane + (size_t)hs * W + ws;
  const float t0 = lerp_row(p, wr), t1 = lerp_row(p + W, wr);
  return fmaf(hr, t1, (1.f - hr) * t0);
}

template <int POOL>
__device__ __forceinline__ float pool4(float a, float b, float c, float d) {
  if (POOL == RLOD_POOL_AVG) return (((a + b) + c) + d) * 0.25f;
  float m = a;
  if (b > m) m = b;
  if (c > m) m = c;
  if (d > m) m = d;
  return m;
}

// ----------------------------------------------------------------------------------------
// generic forward: one thread per output element, taps straight from global (L1/L2)
// ----------------------------------------------------------------------------------------
template <int POOL>
__global__ void __launch_bounds__(256)
    k_align_fwd_generic(const float *__restrict__ feat, const int *__restrict__ plan,
                        const int *__restrict__ roi_b, int C, int H, int W, int GH, int GW,
                        long long total, float *__restrict__ out, const int *__restrict__ only = nullptr,
                        const int *__restrict__ only_count = nullptr) {
  const int ah = POOL == RLOD_POOL_NONE ? GH : GH - 1, aw = POOL == RLOD_POOL_NONE ? GW : GW - 1;
  const int words = 2 * GH + 2 * GW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(idx % aw);
    const int oh = (int)((idx / aw) % ah);
    const int c = (int)((idx / ((long long)aw * ah)) % C);
    int r = (int)(idx / ((long long)aw * ah * C));
    int b;
    if (only) {
      // only the rois the tiled plane kernel left out: only[0 .. *only_count) = their ids, roi_b = -1 - image
      if (r >= *only_count) break;
      r = only[r];
      b = -1 - roi_b[r];
    } else {
      b = roi_b[r];
    }
    const int *pl = plan + (size_t)r * words;
    const float *plane = feat + ((size_t)b * C + c) * ((size_t)H * W);
    const int h0 = pl[oh], w0 = pl[2 * GH + 2 * ow];
    const float hr0 = __int_as_float(pl[GH + oh]), wr0 = __int_as_float(pl[2 * GH + 2 * ow + 1]);
    float v;
    if (POOL == RLOD_POOL_NONE) {
      v = sample_plane(plane, W, h0, hr0, w0, wr0);
    } else {
      const int h1 = pl[oh + 1], w1 = pl[2 * GH + 2 * ow + 2];
      const float hr1 = __int_as_float(pl[GH + oh + 1]);
      const float wr1 = __int_as_float(pl[2 * GH + 2 * ow + 3]);
      const float a = sample_plane(plane, W, h0, hr0, w0, wr0);
      const float b = sample_plane(plane, W, h0, hr0, w1, wr1);
      const float cc = sample_plane(plane, W, h1, hr1, w0, wr0);
      const float d = sample_plane(plane, W, h1, hr1, w1, wr1);
      v = pool4<POOL>(a, b, cc, d);
    }
    out[only ? ((long long)r * C + c) * (ah * aw) + oh * aw + ow : idx] = v;
  }
}

// ----------------------------------------------------------------------------------------
// fast forward: 8x8 sample grid, CTA = (image, 4 channel planes resident in shared memory)
// lane = (channel 0..3, sample column pw 0..7); a warp handles one roi per iteration.
// ----------------------------------------------------------------------------------------
struct Plan8 {
  int hs[8];
  float hr[8];
  int ws;
  float wr;
};

__device__ __forceinline__ void load_plan8(const int *__restrict__ plan, int r, int pw, Plan8 &p) {
  const int4 *q = reinterpret_cast<const int4 *>(plan + (size_t)r * 32);
  const int4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
  p.hs[0] = a.x, p.hs[1] = a.y, p.hs[2] = a.z, p.hs[3] = a.w;
  p.hs[4] = b.x, p.hs[5] = b.y, p.hs[6] = b.z, p.hs[7] = b.w;
  p.hr[0] = __int_as_float(c.x), p.hr[1] = __int_as_float(c.y);
  p.hr[2] = __int_as_float(c.z), p.hr[3] = __int_as_float(c.w);
  p.hr[4] = __int_as_float(d.x), p.hr[5] = __int_as_float(d.y);
  p.hr[6] = __int_as_float(d.z), p.hr[7] = __int_as_float(d.w);
  const int2 w = __ldg(reinterpret_cast<const int2 *>(plan + (size_t)r * 32 + 16) + pw);
  p.ws = w.x;
  p.wr = __int_as_float(w.y);
}

// ----------------------------------------------------------------------------------------
// fast forward ("walk" kernel): 8x8 sample grid, CTA = (image, 4 channel planes resident in
// shared memory), a warp serves FOUR rois per iteration (8 lanes each) and every lane carries
// 4 channels per LDS.128.
//
// Shared-memory plane layout: planes4[row * P + col] = float4(c0, c1, c2, c3).
//   * P (pixels per row) is the smallest odd number >= W + 1, so 8 consecutive rows of one
//     column fall into 8 distinct 16-byte bank groups, exactly like 8 consecutive columns of
//     one row;
//   * column W of every row and rows H, H+1 are zero: an out-of-range sample column / row is
//     redirected there with ratio 0, so validity never appears in the hot loop.
//
// A roi's 8x8 sample grid is separable.  The 8 lanes of a roi own the 8 positions of one axis
// (the "lane axis") and walk the 8 positions of the other ("walk axis"), keeping the two
// interpolated lines of the current walk position in registers (a line is re-used when
// consecutive walk positions share it or shift by one).  Every line costs two LDS.128
// requests per roi: each lane fetches its two neighbouring taps, one per request.  WHICH tap
// goes into the first request is free per lane (lerp(L, R, w) == lerp(R, L, 1 - w)), and the 16
// taps of a line cover each bank group at most twice as long as they span <= 16 pixels -- so
// an assignment exists that makes both requests conflict-free.  k_roi_plan8_walk searches,
// per roi, all 2^8 tap orders for both orientations
//   mode 0: lanes = sample columns, walk = sample rows
//   mode 1: lanes = sample rows,    walk = sample columns
// and keeps the cheapest (fewest shared-memory wavefronts = bank multiplicity x line loads);
// it emits one 32-word (128-byte) record per roi so that the kernel is the same
// straight-line code for every case:
//   [0..7]   walk position t: bit 0 "load line slot a", bit 1 "load line slot b"; bits 4-16 =
//            byte offset of slot a's line (a multiple of 16, < 128 KB); bits 17-29 = pixel offset
//            (bytes / 16) of slot b's line; bit 30 of word 0 = mode
//   [8..15]  r'[t] ratio from slot a to slot b (= 1 - r when the slots hold the pair swapped)
//   [16+2k]  lane k: low half = pixel offset of the tap it fetches first, high half (signed) =
//            pixel offset of the other tap relative to it;  [17+2k] ratio from first to second
// k_roi_order_by_key then partitions every image's roi list by mode so that the four rois a
// warp serves together stage their results with the same strides (bank-disjoint stores).
// ----------------------------------------------------------------------------------------
constexpr int kWalkWarps = 8;
constexpr int kWalkThreads = kWalkWarps * 32;


// Walk plan over the line pairs (L[t], L[t] + 1), t = 0..7.  Two register slots a, b hold
// interpolated lines; a position loads only the lines neither slot holds, and when the pair
// sits in the slots in swapped order (b = first line) the kernel's s = a + r' * (b - a) is made
// right by r' = 1 - r instead of moving registers.  Per position: bit t = load slot a, bit 8+t =
// load slot b, bit 16+t = swapped; la/lb = the line each slot holds after the position.
__device__ __forceinline__ int walk_slots(const int *L, int *la_out, int *lb_out, int &nloads) {
  int la = -5, lb = -5, flags = 0;
  nloads = 0;
  for (int t = 0; t < 8; ++t) {
    const int f = L[t], g = L[t] + 1;
    if (la == f && lb == g) {
    } else if (lb == f && la == g) {
      flags |= 1 << (16 + t);
    } else if (lb == f) {
      la = g;
      flags |= (1 << t) | (1 << (16 + t));
      nloads += 1;
    } else if (la == f) {
      lb = g;
      flags |= 1 << (8 + t);
      nloads += 1;
    } else {
      la = f, lb = g;
      flags |= (1 << t) | (1 << (8 + t));
      nloads += 2;
    }
    la_out[t] = la, lb_out[t] = lb;
  }
  return flags;
}

// LDS.128 wavefronts of one quarter-warp request = the largest number of DISTINCT 16-byte
// words in one of the 8 bank groups.  `occ` is the occupancy bitmap of the request's pixel
// indices along the lane axis (relative to the smallest one); along either axis the bank
// group of an index is a bijection of (index mod 8) because the row pitch is odd.
struct Occ {
  unsigned long long lo, hi;
};
template <bool WIDE>
__device__ __forceinline__ int bank_multiplicity(Occ o) {
  int worst = 1;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const unsigned long long m = 0x0101010101010101ull << g;
    int c = __popcll(o.lo & m);
    if (WIDE) c += __popcll(o.hi & m);
    worst = max(worst, c);
  }
  return worst;
}
template <bool WIDE>
__device__ __forceinline__ void occ_set(Occ &o, int i) {
  if (!WIDE || i < 64) o.lo |= 1ull << i;
  else o.hi |= 1ull << (i - 64);
}

// Best tap order for the lane-axis tap pairs (L[k], R[k]) (bit k set: lane k fetches R first).
// Swapping EVERY lane's order only exchanges the two requests, so tap 7 keeps its natural order and
// 128 orders remain: lane bits 0-4 fix the order of taps 0-4, the loop runs over the orders of taps
// 5-6.  Returns cost << 8 | bits of this lane's best.
template <bool WIDE>
__device__ __forceinline__ unsigned tap_order_search(const int *L, const int *R, int base, int lane) {
  Occ f = {0ull, 0ull}, s = {0ull, 0ull};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const bool sw = (lane >> k) & 1;
    occ_set<WIDE>(f, (sw ? R[k] : L[k]) - base);
    occ_set<WIDE>(s, (sw ? L[k] : R[k]) - base);
  }
  unsigned best = 0xffffffffu;
#pragma unroll 1
  for (int j = 0; j < 4; ++j) {
    Occ ff = f, ss = s;
#pragma unroll
    for (int k = 5; k < 8; ++k) {
      const bool sw = (j >> (k - 5)) & 1;
      occ_set<WIDE>(ff, (sw ? R[k] : L[k]) - base);
      occ_set<WIDE>(ss, (sw ? L[k] : R[k]) - base);
    }
    const unsigned c = (unsigned)(bank_multiplicity<WIDE>(ff) + bank_multiplicity<WIDE>(ss));
    best = min(best, (c << 8) | (unsigned)(lane | (j << 5)));
    if (__any_sync(0xffffffffu, c == 2u)) break;  // conflict-free: cannot be beaten
  }
  return best;
}

template <bool WIDE>
__device__ __forceinline__ int tap_order_natural(const int *L, const int *R, int base) {
  Occ f = {0ull, 0ull}, s = {0ull, 0ull};
#pragma unroll
  for (int k = 0; k < 8; ++k) occ_set<WIDE>(f, L[k] - base), occ_set<WIDE>(s, R[k] - base);
  return bank_multiplicity<WIDE>(f) + bank_multiplicity<WIDE>(s);
}

// one warp per roi
__global__ void __launch_bounds__(128, 6)
    k_roi_plan8_walk(const float *__restrict__ rois, int R, int B, int H, int W, int P, float scale,
                     int bwd, TileGrid tg, AlignWs ws) {
  pdl_trigger();  // k_roi_lists_finish may be set up behind this grid (it waits for our results itself)
  const int r = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  int *e = ws.ext + (size_t)r * 32;
  // the separable 8x8 sample grid (k_roi_plan's arithmetic): lanes 0-7 the sample rows, lanes
  // 8-15 the sample columns, then every lane gets all sixteen by shuffle
  int my_idx = -1, my_ratio = 0;
  {
    const float *roi = rois + (size_t)r * 5;
    const float bf = __ldg(roi);
    const int bi = (int)bf;
    const bool bvalid = (bf >= 0.f) && (bi < B);
    const bool is_row = lane < 8;
    const int p = lane & 7;
    const int dim = is_row ? H : W;
    const float c0 = is_row ? __ldg(roi + 2) : __ldg(roi + 1);
    const float c1 = is_row ? __ldg(roi + 4) : __ldg(roi + 3);
    const float start = __fmul_rn(c0, scale);
    const float size = fmaxf(__fadd_rn(__fmaf_rn(c1, scale, -start), 1.f), 0.f);
    const float bin = (float)((double)size / 7.);
    const float pos = __fmaf_rn((float)p, bin, start);
    if (lane < 16 && bvalid && (pos >= 0.f) && (pos < (float)dim)) {
      my_idx = (int)fminf(floorf(pos), (float)(dim - 2));
      my_ratio = __float_as_int(__fsub_rn(pos, (float)my_idx));
    }
    if (tg.ny * tg.nx == 1) {
      if (lane == 0) roi_list_mark(rois, r, R, B, bvalid ? bi : 0, ws);
    } else {
      // Tiled map: the roi goes to the tile that holds all of its taps (rows lo .. hi + 1, columns likewise);
      // tile origins are tg.sy / tg.sx apart, so the tile whose origin is the last one at or before the first
      // tap is the only candidate that can hold a roi of up to th - sy + 1 rows wherever it lies.  From here on
      // the indices are tile-local and the zero rows / column are the tile's (H, W become th, tw).  A roi
      // that fits no tile is planned as all-zero here and pooled by the generic kernel afterwards, from the
      // k_roi_plan record written below (its id is appended to a list in ws.own; ws.roi_b keeps its image).
      const unsigned full = 0xffffffffu;
      const bool v = my_idx >= 0;
      const int rlo = __reduce_min_sync(full, (v && lane < 8) ? my_idx : (1 << 30));
      const int rhi = __reduce_max_sync(full, (v && lane < 8) ? my_idx + 1 : -1);
      const int clo = __reduce_min_sync(full, (v && lane >= 8) ? my_idx : (1 << 30));
      const int chi = __reduce_max_sync(full, (v && lane >= 8) ? my_idx + 1 : -1);
      const int ty = rhi < 0 ? 0 : min(rlo / tg.sy, tg.ny - 1), tx = chi < 0 ? 0 : min(clo / tg.sx, tg.nx - 1);
      const int oy = ty * tg.sy, ox = tx * tg.sx;
      const bool fits = (rhi < 0 || rhi - oy < tg.th) && (chi < 0 || chi - ox < tg.tw);
      if (!fits && lane < 16) {
        int *pl = ws.plan + (size_t)r * 32;
        if (lane < 8) pl[lane] = my_idx, pl[8 + lane] = my_ratio;
        else pl[16 + 2 * (lane - 8)] = my_idx, pl[17 + 2 * (lane - 8)] = my_ratio;
      }
      if (lane == 0) {
        // the rois left to the generic kernel: ws.own[0 .. count) = their ids, count = ws.flag[2] (zeroed with
        // the flags before this launch); (a roi of an image out of range has no valid sample: it fits)
        if (!fits) ws.own[atomicAdd(ws.flag + 2, 1)] = r;
        // list key: the virtual image of the tile; a roi beyond the tiles carries -1 - image, which matches no
        // list (the plane kernel never sees it) and tells the generic kernel its image
        ws.roi_b[r] = !bvalid ? 0 : (fits ? (bi * tg.ny + ty) * tg.nx + tx : -1 - bi);
        ws.order[r] = r;
        if (r == 0) atomicOr(ws.flag, 1);  // lists by virtual image: always collected, never "grouped"
      }
      if (!fits) my_idx = -1, my_ratio = 0;
      else if (v) my_idx -= lane < 8 ? oy : ox;
      H = tg.th, W = tg.tw;
    }
  }
  int row[8], col[8];
  bool colv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int hs = __shfl_sync(0xffffffffu, my_idx, i), wsx = __shfl_sync(0xffffffffu, my_idx, 8 + i);
    row[i] = hs >= 0 ? hs : H;  // zero rows H, H+1
    colv[i] = wsx >= 0;
    col[i] = wsx >= 0 ? wsx : W;  // zero column
  }
  int nl[2], ra[8], rb[8], ca[8], cb[8];
  const int fl_rows = walk_slots(row, ra, rb, nl[0]), fl_cols = walk_slots(col, ca, cb, nl[1]);
  // lane-axis tap pairs in index space: mode 0 = columns, mode 1 = rows
  int L0[8], R0[8], L1[8], R1[8], base0 = 1 << 30, base1 = 1 << 30, top0 = 0, top1 = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    L0[k] = col[k], R0[k] = colv[k] ? col[k] + 1 : col[k];
    L1[k] = row[k], R1[k] = row[k] + 1;
    base0 = min(base0, L0[k]), top0 = max(top0, R0[k]);
    base1 = min(base1, L1[k]), top1 = max(top1, R1[k]);
  }
  const int span0 = top0 - base0, span1 = top1 - base1;
  const bool ok0 = span0 < 128, ok1 = span1 < 128;
  // natural order first: it is already optimal when it meets the other mode's lower bound
  // (2 wavefronts per line load); only otherwise search the 256 tap orders of a mode
  const int lb0 = 2 * nl[0], lb1 = 2 * nl[1];
  int c0 = 16 * nl[0], c1 = 16 * nl[1];
  if (ok0) c0 = (span0 < 64 ? tap_order_natural<false>(L0, R0, base0) : tap_order_natural<true>(L0, R0, base0)) * nl[0];
  if (ok1) c1 = (span1 < 64 ? tap_order_natural<false>(L1, R1, base1) : tap_order_natural<true>(L1, R1, base1)) * nl[1];
  unsigned best = c1 < c0 ? (((unsigned)c1 << 9) | 256u) : ((unsigned)c0 << 9);  // cost << 9 | mode << 8 | bits
  // the backward kernel scatters: it needs mode 0 (one row lock per line) and the natural tap
  // order (two lanes may not hit one pixel in the same request)
  if (bwd) best = 0u, c0 = c1 = 0;
  const int nat = min(c0, c1);
#ifdef RLOD_NO_SEARCH  // ablation (profiles/experiments/README.md): natural tap order only
  if (false) {
#else
  if (ok0 && lb0 < nat && c0 > lb0) {  // warp-uniform
#endif
    const unsigned q = span0 < 64 ? tap_order_search<false>(L0, R0, base0, lane) : tap_order_search<true>(L0, R0, base0, lane);
    best = min(best, (((q >> 8) * (unsigned)nl[0]) << 9) | (q & 255u));
  }
#ifdef RLOD_NO_SEARCH
  if (false) {
#else
  if (ok1 && lb1 < nat && c1 > lb1) {
#endif
    const unsigned q = span1 < 64 ? tap_order_search<false>(L1, R1, base1, lane) : tap_order_search<true>(L1, R1, base1, lane);
    best = min(best, (((q >> 8) * (unsigned)nl[1]) << 9) | 256u | (q & 255u));
  }
  best = __reduce_min_sync(0xffffffffu, best);
  const int mode = (best >> 8) & 1, bits = best & 255;
  // lane t < 8 computed sample row t itself; sample column t sits in lane 8 + t
  const int wsx = __shfl_sync(0xffffffffu, my_idx, 8 + (lane & 7)), wrt = __shfl_sync(0xffffffffu, my_ratio, 8 + (lane & 7));
  if (lane >= 8) return;
  // lane t writes walk position t and lane-axis entry t; pixel offsets (bytes / 16) stay
  // below 8192 (host-checked)
  const int t = lane;
  const int hs = my_idx, hrt = my_ratio;
  const int rowt = hs >= 0 ? hs : H, colt = wsx >= 0 ? wsx : W;
  const int colt1 = wsx >= 0 ? colt + 1 : colt;
  int sa = 0, sb = 0;  // the lines slots a / b hold at walk position t
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i == t) sa = mode ? ca[i] : ra[i], sb = mode ? cb[i] : rb[i];
  const int fl = mode ? fl_cols : fl_rows;
  // line -> pixel offset: rows scale by the pitch; the line after the zero column is itself
  const int A = mode ? min(sa, W) : sa * P;
  const int Bq = mode ? min(sb, W) : sb * P;
  const int f2 = ((fl >> t) & 1) | (((fl >> (8 + t)) & 1) << 1);
  e[t] = f2 | (A << 4) | (Bq << 17) | ((t == 0 && mode) ? (1 << 30) : 0);
  {
    const float rr = __int_as_float(mode ? wrt : hrt);
    e[8 + t] = __float_as_int(((fl >> (16 + t)) & 1) ? 1.f - rr : rr);
  }
  const int Lk = mode ? rowt * P : colt;
  const int Rk = mode ? (rowt + 1) * P : colt1;
  const float w = __int_as_float(mode ? hrt : wrt);
  const bool sw = (bits >> t) & 1;
  const int first = sw ? Rk : Lk, second = sw ? Lk : Rk;
  if (!bwd) {
    e[16 + 2 * t] = (first & 0xffff) | ((second - first) << 16);
  } else {
    // scatter: lanes of one roi that share a column take turns; rank = position inside the
    // run of equal columns, maxrank = the longest run - 1 (invalid columns land in the zero
    // column, whose content is never read: rank 0)
    int rank = 0, maxrank = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int run = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < i && colv[j] && colv[i] && col[j] == col[i]) ++run;
      if (i == t) rank = run;
      maxrank = max(maxrank, run);
    }
    e[16 + 2 * t] = (first & 0xffff) | ((second - first) << 16) | (rank << 17) | (maxrank << 20);
  }
  e[17 + 2 * t] = __float_as_int(sw ? 1.f - w : w);
}

struct WalkRec {
  int w[8];     // packed walk positions
  float rt[8];  // walk ratios
  int la;       // packed lane taps
  float wa;     // lane ratio
};

__device__ __forceinline__ void walk_rec_clear(WalkRec &p) {
#pragma unroll
  for (int i = 0; i < 8; ++i) p.w[i] = 0, p.rt[i] = 0.f;
  p.la = 0, p.wa = 0.f;
}

__device__ __forceinline__ void walk_rec_load(const int *__restrict__ ext, int r, int k, WalkRec &p) {
  const int4 *q = reinterpret_cast<const int4 *>(ext + (size_t)r * 32);
  const int4 a0 = __ldg(q), a1 = __ldg(q + 1), c0 = __ldg(q + 2), c1 = __ldg(q + 3);
  p.w[0] = a0.x, p.w[1] = a0.y, p.w[2] = a0.z, p.w[3] = a0.w;
  p.w[4] = a1.x, p.w[5] = a1.y, p.w[6] = a1.z, p.w[7] = a1.w;
  p.rt[0] = __int_as_float(c0.x), p.rt[1] = __int_as_float(c0.y);
  p.rt[2] = __int_as_float(c0.z), p.rt[3] = __int_as_float(c0.w);
  p.rt[4] = __int_as_float(c1.x), p.rt[5] = __int_as_float(c1.y);
  p.rt[6] = __int_as_float(c1.z), p.rt[7] = __int_as_float(c1.w);
  const int2 l = __ldg(reinterpret_cast<const int2 *>(ext + (size_t)r * 32 + 16) + k);
  p.la = l.x;
  p.wa = __int_as_float(l.y);
}

// One walk position for 4 channels, fully predicated on two flag bits of this lane's roi:
//   bit 0: fetch the line of slot a (t0);  bit 1: fetch the line of slot b (t1).
//   line value = x0 + wa * (x1 - x0) over the lane's two taps; s = t0 + r' * (t1 - t0).
// One PTX block so that the predicates guard the LDS.128 and the FFMAs directly: the four
// rois of a warp follow different flag patterns without a single divergent branch.
// The tap registers (x, y: first line, z, w: second line) are C++ variables handed in as
// read-write operands: a predicated load does not kill its destination, so block-local PTX
// temporaries would be live across the whole loop and cost 16 registers PER walk position.
struct WalkTaps {
  float4 x, y, z, w;
};

__device__ __forceinline__ void walk_step(float4 &t0, float4 &t1, float4 &s, WalkTaps &q,
                                          uint32_t pa, uint32_t pb, uint32_t pa2, uint32_t pb2,
                                          float wa, float r, int flags) {
  asm volatile(
      "{\n"
      ".reg .pred pr, pd;\n"
      ".reg .b32 tt;\n"
      ".reg .f32 d;\n"
      "and.b32 tt, %34, 1;\n"
      "setp.ne.b32 pr, tt, 0;\n"
      "and.b32 tt, %34, 2;\n"
      "setp.ne.b32 pd, tt, 0;\n"
      "@pr ld.shared.v4.f32 {%12, %13, %14, %15}, [%28];\n"
      "@pr ld.shared.v4.f32 {%16, %17, %18, %19}, [%29];\n"
      "@pd ld.shared.v4.f32 {%20, %21, %22, %23}, [%30];\n"
      "@pd ld.shared.v4.f32 {%24, %25, %26, %27}, [%31];\n"
      "@pr sub.f32 d, %16, %12;\n"
      "@pr fma.rn.f32 %0, %32, d, %12;\n"
      "@pr sub.f32 d, %17, %13;\n"
      "@pr fma.rn.f32 %1, %32, d, %13;\n"
      "@pr sub.f32 d, %18, %14;\n"
      "@pr fma.rn.f32 %2, %32, d, %14;\n"
      "@pr sub.f32 d, %19, %15;\n"
      "@pr fma.rn.f32 %3, %32, d, %15;\n"
      "@pd sub.f32 d, %24, %20;\n"
      "@pd fma.rn.f32 %4, %32, d, %20;\n"
      "@pd sub.f32 d, %25, %21;\n"
      "@pd fma.rn.f32 %5, %32, d, %21;\n"
      "@pd sub.f32 d, %26, %22;\n"
      "@pd fma.rn.f32 %6, %32, d, %22;\n"
      "@pd sub.f32 d, %27, %23;\n"
      "@pd fma.rn.f32 %7, %32, d, %23;\n"
      "sub.f32 d, %4, %0;\n"
      "fma.rn.f32 %8, %33, d, %0;\n"
      "sub.f32 d, %5, %1;\n"
      "fma.rn.f32 %9, %33, d, %1;\n"
      "sub.f32 d, %6, %2;\n"
      "fma.rn.f32 %10, %33, d, %2;\n"
      "sub.f32 d, %7, %3;\n"
      "fma.rn.f32 %11, %33, d, %3;\n"
      "}\n"
      : "+f"(t0.x), "+f"(t0.y), "+f"(t0.z), "+f"(t0.w), "+f"(t1.x), "+f"(t1.y), "+f"(t1.z),
        "+f"(t1.w), "=f"(s.x), "=f"(s.y), "=f"(s.z), "=f"(s.w), "+f"(q.x.x), "+f"(q.x.y),
        "+f"(q.x.z), "+f"(q.x.w), "+f"(q.y.x), "+f"(q.y.y), "+f"(q.y.z), "+f"(q.y.w),
        "+f"(q.z.x), "+f"(q.z.y), "+f"(q.z.z), "+f"(q.z.w), "+f"(q.w.x), "+f"(q.w.y),
        "+f"(q.w.z), "+f"(q.w.w)
      : "r"(pa), "r"(pb), "r"(pa2), "r"(pb2), "f"(wa), "f"(r), "r"(flags));
}

// four staged values of one lane (channels 0..3, OHW floats apart), stored only when p
__device__ __forceinline__ void sts4_if(float *q, int ohw_bytes, float4 v, bool p) {
  const uint32_t a = smem_u32(q);
  asm volatile(
      "{\n"
      ".reg .pred pp;\n"
      "setp.ne.b32 pp, %5, 0;\n"
      "@pp st.shared.f32 [%0], %1;\n"
      "@pp st.shared.f32 [%6], %2;\n"
      "@pp st.shared.f32 [%7], %3;\n"
      "@pp st.shared.f32 [%8], %4;\n"
      "}\n" ::"r"(a),
      "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)p), "r"(a + ohw_bytes), "r"(a + 2 * ohw_bytes),
      "r"(a + 3 * ohw_bytes)
      : "memory");
}

template <int POOL>
__device__ __forceinline__ float pool2(float a, float b) {
  return POOL == RLOD_POOL_AVG ? a + b : fmaxf(a, b);
}
template <int POOL>
__device__ __forceinline__ float4 pool2(float4 a, float4 b) {
  return make_float4(pool2<POOL>(a.x, b.x), pool2<POOL>(a.y, b.y), pool2<POOL>(a.z, b.z),
                     pool2<POOL>(a.w, b.w));
}

template <int POOL>
__global__ void __launch_bounds__(kWalkThreads, 2)
    k_align8_fwd_walk(const float *__restrict__ feat, const int *__restrict__ ext,
                      const int *__restrict__ order, const int *__restrict__ img_off, int C,
                      int H, int W, int P, int n_chunks, int tma_fill, float *__restrict__ out) {
  constexpr int OW = POOL == RLOD_POOL_NONE ? 8 : 7;
  constexpr int OHW = OW * OW;                            // 64 | 49
  constexpr int STG = 4 * OHW;                            // floats per (roi, 4 channels)
  constexpr int SLOT = POOL == RLOD_POOL_NONE ? 264 : 200;  // = 8 (mod 32): slots bank-disjoint
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *planes4 = reinterpret_cast<float4 *>(smem_raw);
  const int HW = H * W;
  float *stage = reinterpret_cast<float *>(planes4 + (H + 2) * P);

  const int b = blockIdx.x / n_chunks, chunk = blockIdx.x - b * n_chunks;
  const int r0 = img_off[b], r1 = img_off[b + 1];
  if (r0 >= r1) return;
  const float *src = feat + ((size_t)b * C + (size_t)chunk * 4) * HW;
  // ---- fill: the CTA's 4 planes, HBM -> shared memory, interleaved per pixel ---------------
  // Fast way (tma_fill): four bulk async copies (TMA engine: a plane is one contiguous run, no
  // per-sector request tracking in the LSU) land the planes PLANAR -- planes 1..3 in the output
  // staging area, which is idle until the first roi, plane 0 at the tail of the plane region
  // -- and the threads interleave them in place, batch by batch (the host checked that a batch
  // never overwrites plane-0 pixels that are still unread).  A plane starts on an 8-byte
  // boundary when H*W is even but not a multiple of 4: the copy starts 8 bytes early (skew).
  // Fallback: 4-byte async copies (LDGSTS), interleaved in flight.
  uint64_t *fill_bar = reinterpret_cast<uint64_t *>(stage + kWalkWarps * 2 * 4 * SLOT);
#if RLOD_ABL != 1 && RLOD_ABL != 4
  if (tma_fill) {
    const uint32_t plane_copy = (uint32_t)((HW * 4 + 8 + 15) & ~15);  // bytes per bulk copy (skew <= 8)
    unsigned char *buf[4];
    buf[0] = smem_raw + (size_t)(H + 2) * P * 16 - plane_copy;
    for (int c = 1; c < 4; ++c) buf[c] = reinterpret_cast<unsigned char *>(stage) + (size_t)(c - 1) * plane_copy;
    if (threadIdx.x == 0) {
      mbar_init(fill_bar, 1);
      uint32_t total = 0;
      for (int c = 0; c < 4; ++c) {
        const uint32_t skew = (uint32_t)((size_t)c * HW * 4) & 15u;
        total += (skew + (uint32_t)HW * 4 + 15u) & ~15u;
      }
      mbar_expect_tx(fill_bar, total);
      for (int c = 0; c < 4; ++c) {
        const uint32_t skew = (uint32_t)((size_t)c * HW * 4) & 15u;
        bulk_g2s(buf[c], reinterpret_cast<const char *>(src + (size_t)c * HW) - skew,
                 (skew + (uint32_t)HW * 4 + 15u) & ~15u, fill_bar);
      }
    }
    __syncthreads();  // barrier initialised before anyone polls it
    mbar_wait(fill_bar, 0);
    const float *pl[4];
    for (int c = 0; c < 4; ++c)
      pl[c] = reinterpret_cast<const float *>(buf[c] + ((uint32_t)((size_t)c * HW * 4) & 15u));
    constexpr int kPer = 4;  // pixels per thread and batch
    for (int base = 0; base < HW; base += kWalkThreads * kPer) {
      float4 v[kPer];
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const int p = base + q * kWalkThreads + threadIdx.x;
        if (p < HW) v[q] = make_float4(pl[0][p], pl[1][p], pl[2][p], pl[3][p]);
      }
      __syncthreads();  // plane-0 pixels of this batch are in registers before the region is written
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const int p = base + q * kWalkThreads + threadIdx.x;
        if (p < HW) {
          const int y = p / W, x = p - y * W;
          planes4[y * P + x] = v[q];
        }
      }
    }
    __syncthreads();  // staging and the plane tail are free again
  } else {
    fill_planes4_async<kWalkThreads>(planes4, src, H, W, P, HW);
  }
#endif
  // The CTA that will run on this SM slot one CTA-lifetime from now reads cold planes from HBM
  // while its warps can do nothing else: pull those planes into L2 now (same bytes, earlier).
  {
    const unsigned nb2 = blockIdx.x + 2u * (unsigned)kSmCount;
    if (nb2 < gridDim.x) {
      const unsigned b2 = nb2 / (unsigned)n_chunks, c2 = nb2 - b2 * (unsigned)n_chunks;
      const char *nx = reinterpret_cast<const char *>(feat + ((size_t)b2 * C + (size_t)c2 * 4) * HW);
      for (int i = threadIdx.x * 128; i < 4 * HW * 4; i += kWalkThreads * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + i));
    }
  }
  {
    const int padw = P - W;  // zero columns W .. P-1 of the data rows, then the two zero rows
    for (int p = threadIdx.x; p < H * padw; p += kWalkThreads) {
      const int y = p / padw, x = W + (p - y * padw);
      planes4[y * P + x] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int p = threadIdx.x; p < 2 * P; p += kWalkThreads) planes4[H * P + p] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = lane & 7, slot = lane >> 3;
  const uint32_t pbase = smem_u32(planes4);
  float *stg = stage + (warp * 2) * (4 * SLOT) + slot * SLOT;  // + (it & 1) * 4 * SLOT
  const int n_groups = (r1 - r0 + 3) >> 2;

  // the record of the NEXT group travels in registers while this one is computed (18 words),
  // and the roi id of the group after that one too (order -> record is a dependent pair of
  // loads: issued back to back it would stall the warp for an L2 round trip every iteration)
  auto roi_of = [&](int gg) {
    const int kk = r0 + 4 * gg + slot;
    return (gg < n_groups && kk < r1) ? __ldg(order + kk) : -1;
  };
  WalkRec cur;
  walk_rec_clear(cur);
  int r = roi_of(warp);
  if (r >= 0) walk_rec_load(ext, r, k, cur);
  int rn = roi_of(warp + kWalkWarps);
  WalkTaps taps[2];  // two sets: position T+1 loads while T computes
  taps[0].x = taps[0].y = taps[0].z = taps[0].w = make_float4(0.f, 0.f, 0.f, 0.f);
  taps[1] = taps[0];
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  for (int g = warp, it = 0; g < n_groups; g += kWalkWarps, ++it) {
    WalkRec nxt;
    walk_rec_clear(nxt);
    if (rn >= 0) walk_rec_load(ext, rn, k, nxt);
    const int rnn = roi_of(g + 2 * kWalkWarps);
    float *sbuf = stg + (it & 1) * (4 * SLOT);
    const bool mode1 = (cur.w[0] >> 30) & 1;
    const int st_t = mode1 ? 1 : OW, st_k = mode1 ? OW : 1;
    float *sp = sbuf + k * st_k;
    const uint32_t cbase = pbase + (((uint32_t)cur.la & 0xffffu) << 4);
    const uint32_t da = (uint32_t)((cur.la >> 16) << 4);
    float4 s[8];
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
#define RLOD_WALK(T)                                                                        \
  {                                                                                         \
    const uint32_t pa = cbase + ((uint32_t)cur.w[T] & 0x1fff0u);                            \
    const uint32_t pa2 = cbase + (((uint32_t)cur.w[T] >> 13) & 0x1fff0u);                   \
    walk_step(t0, t1, s[T], taps[(T) & 1], pa, pa + da, pa2, pa2 + da, cur.wa, cur.rt[T], (RLOD_ABL == 3 || RLOD_ABL == 4) ? (cur.w[T] & ~3) : cur.w[T]); \
  }
    // output of walk position T: NONE stores the sample itself; AVG / MAX pool 2x2 stride 1,
    // along the walk axis in registers and along the lane axis with one shuffle per value
#define RLOD_EMIT(T)                                                                        \
  if (POOL == RLOD_POOL_NONE) {                                                             \
    float *q = sp + (T) * st_t;                                                             \
    q[0] = s[T].x, q[OHW] = s[T].y, q[2 * OHW] = s[T].z, q[3 * OHW] = s[T].w;               \
  } else if ((T) < 7) {                                                                     \
    const float4 v = pool2<POOL>(s[T], s[(T) < 7 ? (T) + 1 : 7]);                           \
    float4 n;                                                                               \
    n.x = __shfl_down_sync(0xffffffffu, v.x, 1, 8);                                         \
    n.y = __shfl_down_sync(0xffffffffu, v.y, 1, 8);                                         \
    n.z = __shfl_down_sync(0xffffffffu, v.z, 1, 8);                                         \
    n.w = __shfl_down_sync(0xffffffffu, v.w, 1, 8);                                         \
    float4 o = pool2<POOL>(v, n);                                                           \
    if (POOL == RLOD_POOL_AVG) o = make_float4(o.x * 0.25f, o.y * 0.25f, o.z * 0.25f, o.w * 0.25f); \
    sts4_if(sp + (T) * st_t, OHW * 4, o, k < 7);                                            \
  }
    // software pipeline: the stores of position T-2 sit in the shadow of the loads of T
    RLOD_WALK(0) RLOD_WALK(1) RLOD_WALK(2)
    // the buffer about to be written was handed to the bulk-copy engine two iterations ago
    if (it >= 2) {
      if (k == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    RLOD_EMIT(0) RLOD_WALK(3) RLOD_EMIT(1)
    RLOD_WALK(4) RLOD_EMIT(2) RLOD_WALK(5) RLOD_EMIT(3) RLOD_WALK(6) RLOD_EMIT(4)
    RLOD_WALK(7) RLOD_EMIT(5) RLOD_EMIT(6) RLOD_EMIT(7)
#undef RLOD_WALK
#undef RLOD_EMIT
    // 4 channels x OHW floats are one contiguous, 16-byte aligned run of the (R,C,OH,OW)
    // output: each roi's staged block leaves as a single bulk store
    fence_async_smem();
    __syncwarp();
    if (k == 0) {
#if RLOD_ABL != 2
      if (r >= 0)
        bulk_s2g_nocommit(out + ((size_t)r * C + (size_t)chunk * 4) * OHW, sbuf, (uint32_t)(STG * sizeof(float)));
#endif
      bulk_commit();
    }
    cur = nxt;
    r = rn;
    rn = rnn;
  }
  if (k == 0) bulk_wait_read<0>();  // shared memory must outlive the engine's reads
}

// ----------------------------------------------------------------------------------------
// k_align8_fwd_walk2: the same walk with packed fp32 arithmetic.  sm_100 has two-wide fp32
// instructions (FADD2 / FMUL2 / FFMA2 on 64-bit register pairs), and LDS.128 delivers the four
// channels of a tap as exactly two such pairs: every lerp of the walk is one FADD2 + one FFMA2
// per PAIR of channels -- half the fp32 issue slots of the scalar walk (276 of its 646
// instructions per group of four rois were fp32; the kernel is bound by issue slots and the
// shared-memory pipe, not by HBM, profiles/r01_align_fwd.md).  Further differences:
//   * RoIAlignAvg: the 1/4 of the 2x2 average is applied to the planes once at fill time
//     (exact: a power of two), not to every pooled value;
//   * the staging stores are plain C++ under the loop-invariant predicate k < 7 (the asm
//     volatile block made ptxas wrap them in BSSY / BRA / BSYNC: 65 control instructions per
//     group);
//   * the loop is unrolled by two over two record register sets (no 18-register move per
//     iteration) and the staging buffer parity is a compile-time constant;
//   * the interleave of the fill walks (y, x) incrementally instead of dividing per pixel;
//   * rois are taken from a list partitioned by walk mode (k_roi_order_by_key), so the four
//     rois of a warp stage with the same strides and their stores stay bank-disjoint.
// ----------------------------------------------------------------------------------------
typedef unsigned long long u64;
struct F4 {
  u64 a, b;  // channels (0, 1) and (2, 3) as packed f32x2
};
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 unpack2(u64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ u64 add2(u64 x, u64 y) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
  return r;
}
__device__ __forceinline__ u64 max2(u64 x, u64 y) {
  const float2 a = unpack2(x), b = unpack2(y);
  return pack2(fmaxf(a.x, b.x), fmaxf(a.y, b.y));
}
__device__ __forceinline__ u64 shfl_down8(u64 v) { return __shfl_down_sync(0xffffffffu, v, 1, 8); }

struct WalkTaps2 {
  u64 x0, x1, y0, y1, z0, z1, w0, w1;  // line slot a: taps x, y; line slot b: taps z, w
};
__device__ __forceinline__ u64 lerp2(u64 p, u64 q, u64 w) {  // p + w * (q - p), two channels
  u64 d, r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(q), "l"(p));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(w), "l"(d), "l"(p));
  return r;
}

// The loads of one walk position: bit 0 of `flags` fetches the two taps of line slot a, bit 1 those
// of slot b; a slot that is not fetched keeps the taps it holds (predicated LDS.128 into read-write
// operands).  Only the loads are predicated: the lerps that follow are recomputed from the slot's
// taps whether or not they were just fetched (same inputs, same result), because a predicated
// packed fp32 instruction costs ptxas a pair of SELs on top.
__device__ __forceinline__ void walk_load2(WalkTaps2 &q, uint32_t pa, uint32_t pb, uint32_t pa2,
                                           uint32_t pb2, int flags) {
  asm volatile(
      "{\n"
      ".reg .pred pr, pd;\n"
      ".reg .b32 tt;\n"
      "and.b32 tt, %12, 1;\n"
      "setp.ne.b32 pr, tt, 0;\n"
      "and.b32 tt, %12, 2;\n"
      "setp.ne.b32 pd, tt, 0;\n"
      "@pr ld.shared.v2.b64 {%0, %1}, [%8];\n"
      "@pr ld.shared.v2.b64 {%2, %3}, [%9];\n"
      "@pd ld.shared.v2.b64 {%4, %5}, [%10];\n"
      "@pd ld.shared.v2.b64 {%6, %7}, [%11];\n"
      "}\n"
      : "+l"(q.x0), "+l"(q.x1), "+l"(q.y0), "+l"(q.y1), "+l"(q.z0), "+l"(q.z1), "+l"(q.w0), "+l"(q.w1)
      : "r"(pa), "r"(pb), "r"(pa2), "r"(pb2), "r"(flags));
}

// TILED: the lists are per tile of a larger map (TileGrid, roi_lists.cuh): H, W are the tile's size, Hf, Wf
// the map's; the planes are filled from the tile's window of the map (4-byte async copies, row stride Wf).
template <int POOL, bool TILED>
__global__ void __launch_bounds__(kWalkThreads, 2)
    k_align8_fwd_walk2(const float *__restrict__ feat, const int *__restrict__ ext,
                       const int *__restrict__ order, const int *__restrict__ img_off, int C,
                       int H, int W, int P, int n_chunks, int tma_fill, int split_from, int split,
                       TileGrid tg, int Hf, int Wf, float *__restrict__ out) {
  constexpr int OW = POOL == RLOD_POOL_NONE ? 8 : 7;
  constexpr int OHW = OW * OW;                            // 64 | 49
  constexpr int STG = 4 * OHW;                            // floats per (roi, 4 channels)
  constexpr int SLOT = POOL == RLOD_POOL_NONE ? 264 : 200;  // = 8 (mod 32): slots bank-disjoint
  constexpr float kPre = POOL == RLOD_POOL_AVG ? 0.25f : 1.f;  // plane prescale (exact)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *planes4 = reinterpret_cast<float4 *>(smem_raw);
  const int HW = H * W;
  float *stage = reinterpret_cast<float *>(planes4 + (H + 2) * P);

  // CTA = (image, 4 channels) for the first split_from blocks (whole waves of the grid); the items of the
  // last, partial wave are served by `split` CTAs each, every one with its own copy of the planes and a
  // contiguous share of the image's roi groups, so that the tail of the launch fills the SMs it would
  // leave idle (align_fwd_run picks split)
  unsigned item = blockIdx.x, part = 0, parts = 1;
  if ((int)blockIdx.x >= split_from) {
    const unsigned t = blockIdx.x - (unsigned)split_from;
    item = (unsigned)split_from + t / (unsigned)split, part = t % (unsigned)split, parts = (unsigned)split;
  }
  const int b = (int)item / n_chunks, chunk = (int)item - b * n_chunks;
  // The fill below reads the feature map only: it may run while the roi plan (k_roi_plan8_walk +
  // k_roi_lists_finish, launched just before on the stream) is still being written -- pdl_wait()
  // after the fill is where this grid meets the plan (rlod_roi_align_forward launches it as a
  // programmatic dependent of the list kernel).
  const float *src = feat + ((size_t)b * C + (size_t)chunk * 4) * HW;
  uint64_t *fill_bar = reinterpret_cast<uint64_t *>(stage + kWalkWarps * 2 * 4 * SLOT);
  if constexpr (TILED) {
    // most tiles of a large map hold no roi: look at the list before paying for the fill
    pdl_wait();
    if (img_off[b] >= img_off[b + 1]) return;
    const int tiles = tg.ny * tg.nx;
    const int bi = b / tiles, t = b - bi * tiles, ty = t / tg.nx, tx = t - ty * tg.nx;
    const int oy = ty * tg.sy, ox = tx * tg.sx;
    const int rows = min(H, Hf - oy), cols = min(W, Wf - ox);
    if (rows < H || cols < W) {  // a tile that hangs over the edge: no tap lands there, but keep it finite
      for (int p = threadIdx.x; p < H * P; p += kWalkThreads) planes4[p] = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncthreads();
    }
    const size_t HWf = (size_t)Hf * Wf;
    const float *win = feat + ((size_t)bi * C + (size_t)chunk * 4) * HWf + (size_t)oy * Wf + ox;
    if (tma_fill == 3) {  // 16-byte aligned window rows: 128-bit loads, prescale on the way
      fill_planes4_window_vec<kWalkThreads>(planes4, win, rows, cols, P, Wf, HWf, kPre);
    } else {
      fill_planes4_window_async<kWalkThreads>(planes4, win, rows, cols, P, Wf, HWf);
    }
    if (POOL == RLOD_POOL_AVG && tma_fill != 3) {  // prescale in place once the async copies have landed
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      for (int p = threadIdx.x; p < H * P; p += kWalkThreads) {
        float4 v = planes4[p];
        planes4[p] = make_float4(kPre * v.x, kPre * v.y, kPre * v.z, kPre * v.w);
      }
      __syncthreads();  // the pass touched the pad columns too: they are zeroed below, by other threads
    }
  } else
  // ---- fill (see k_align8_fwd_walk): bulk copies land the planes planar, threads interleave ----
  // tma_fill: 1 = NCHW by bulk copies, 0 = NCHW by 4-byte async copies, 2 = channels-last input
  if (tma_fill == 2) {
    fill_planes4_nhwc_async<kWalkThreads>(planes4, feat + (size_t)b * HW * C + (size_t)chunk * 4, H, W, P, C);
    if (POOL == RLOD_POOL_AVG) {  // prescale in place once the async copies have landed
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      for (int p = threadIdx.x; p < H * P; p += kWalkThreads) {
        float4 v = planes4[p];
        planes4[p] = make_float4(kPre * v.x, kPre * v.y, kPre * v.z, kPre * v.w);
      }
      __syncthreads();  // the pass touched the pad columns too: they are zeroed below, by other threads
    }
  } else if (tma_fill) {
    const uint32_t plane_copy = (uint32_t)((HW * 4 + 8 + 15) & ~15);
    unsigned char *buf[4];
    buf[0] = smem_raw + (size_t)(H + 2) * P * 16 - plane_copy;
    for (int c = 1; c < 4; ++c) buf[c] = reinterpret_cast<unsigned char *>(stage) + (size_t)(c - 1) * plane_copy;
    if (threadIdx.x == 0) {
      mbar_init(fill_bar, 1);
      uint32_t total = 0;
      for (int c = 0; c < 4; ++c) {
        const uint32_t skew = (uint32_t)((size_t)c * HW * 4) & 15u;
        total += (skew + (uint32_t)HW * 4 + 15u) & ~15u;
      }
      mbar_expect_tx(fill_bar, total);
      for (int c = 0; c < 4; ++c) {
        const uint32_t skew = (uint32_t)((size_t)c * HW * 4) & 15u;
        bulk_g2s_hint(buf[c], reinterpret_cast<const char *>(src + (size_t)c * HW) - skew,
                      (skew + (uint32_t)HW * 4 + 15u) & ~15u, fill_bar, l2_policy_evict_first());
      }
    }
    __syncthreads();  // barrier initialised before anyone polls it
    // next planes of this SM slot into L2 while the copies are in flight
    {
      const unsigned nb2 = blockIdx.x + 2u * (unsigned)kSmCount;
      if (nb2 < (unsigned)split_from) {
        const unsigned b2 = nb2 / (unsigned)n_chunks, c2 = nb2 - b2 * (unsigned)n_chunks;
        const char *nx = reinterpret_cast<const char *>(feat + ((size_t)b2 * C + (size_t)c2 * 4) * HW);
        for (int i = threadIdx.x * 128; i < 4 * HW * 4; i += kWalkThreads * 128)
          asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(nx + i));
      }
    }
    mbar_wait(fill_bar, 0);
    const float *pl[4];
    for (int c = 0; c < 4; ++c)
      pl[c] = reinterpret_cast<const float *>(buf[c] + ((uint32_t)((size_t)c * HW * 4) & 15u));
    constexpr int kPer = 4;  // pixels per thread and batch
    // pixel p = base + q * kWalkThreads + tid walks (y, x) by constant steps: no division per pixel
    const int dy = kWalkThreads / W, dx = kWalkThreads - dy * W;
    int y = (int)threadIdx.x / W, x = (int)threadIdx.x - y * W;
    for (int base = 0; base < HW; base += kWalkThreads * kPer) {
      float4 v[kPer];
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const int p = base + q * kWalkThreads + threadIdx.x;
        if (p < HW) v[q] = make_float4(kPre * pl[0][p], kPre * pl[1][p], kPre * pl[2][p], kPre * pl[3][p]);
      }
      __syncthreads();  // plane-0 pixels of this batch are in registers before the region is written
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        const int p = base + q * kWalkThreads + threadIdx.x;
        if (p < HW) planes4[y * P + x] = v[q];
        x += dx, y += dy;
        if (x >= W) x -= W, ++y;
      }
    }
    __syncthreads();  // staging and the plane tail are free again
  } else {
    fill_planes4_async<kWalkThreads>(planes4, src, H, W, P, HW);
    const unsigned nb2 = blockIdx.x + 2u * (unsigned)kSmCount;
    if (nb2 < (unsigned)split_from) {
      const unsigned b2 = nb2 / (unsigned)n_chunks, c2 = nb2 - b2 * (unsigned)n_chunks;
      const char *nx = reinterpret_cast<const char *>(feat + ((size_t)b2 * C + (size_t)c2 * 4) * HW);
      for (int i = threadIdx.x * 128; i < 4 * HW * 4; i += kWalkThreads * 128)
        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(nx + i));
    }
    if (POOL == RLOD_POOL_AVG) {  // prescale in place once the async copies have landed
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      for (int p = threadIdx.x; p < H * P; p += kWalkThreads) {
        float4 v = planes4[p];
        planes4[p] = make_float4(kPre * v.x, kPre * v.y, kPre * v.z, kPre * v.w);
      }
      __syncthreads();  // the pass touched the pad columns too: they are zeroed below, by other threads
    }
  }
  {
    const int padw = P - W;  // zero columns W .. P-1 of the data rows, then the two zero rows
    for (int p = threadIdx.x; p < H * padw; p += kWalkThreads) {
      const int yy = p / padw, xx = W + (p - yy * padw);
      planes4[yy * P + xx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int p = threadIdx.x; p < 2 * P; p += kWalkThreads) planes4[H * P + p] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  pdl_wait();  // the roi plan is complete and visible from here on
  const int r0 = img_off[b], r1 = img_off[b + 1];
  if (r0 >= r1) {  // an image without rois (uniform over the CTA)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = lane & 7, slot = lane >> 3;
  const uint32_t pbase = smem_u32(planes4);
  float *stg = stage + (warp * 2) * (4 * SLOT) + slot * SLOT;  // + parity * 4 * SLOT
  const int n_groups_all = (r1 - r0 + 3) >> 2;
  // this CTA's share of the image's groups: [g_lo, n_groups)
  const int g_lo = (int)(((long long)n_groups_all * part) / parts);
  const int n_groups = (int)(((long long)n_groups_all * (part + 1)) / parts);
  const bool kst = POOL == RLOD_POOL_NONE || k < 7;  // this lane stores (lane 7 only pools for lane 6)
  float *out_chunk = out + (size_t)chunk * 4 * OHW;
  const size_t roi_pitch = (size_t)C * OHW;
  const uint64_t pol_out = l2_policy_evict_first();

  auto roi_of = [&](int gg) {
    const int kk = r0 + 4 * gg + slot;
    return (gg < n_groups && kk < r1) ? __ldg(order + kk) : -1;
  };
  // two record register sets: one is computed while the other is loaded
  WalkRec recA, recB;
  walk_rec_clear(recA);
  int ra = roi_of(g_lo + warp);
  if (ra >= 0) walk_rec_load(ext, ra, k, recA);
  int rb = roi_of(g_lo + warp + kWalkWarps);
  WalkTaps2 taps;
  taps.x0 = taps.x1 = taps.y0 = taps.y1 = taps.z0 = taps.z1 = taps.w0 = taps.w1 = 0ull;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // one group of four rois; PAR = parity of the staging buffer (compile-time)
  auto body = [&](const WalkRec &cur, int r, WalkRec &nxt, int rn, int it, auto par) {
    constexpr int PAR = decltype(par)::value;
    walk_rec_clear(nxt);
    if (rn >= 0) walk_rec_load(ext, rn, k, nxt);
    float *sbuf = stg + PAR * (4 * SLOT);
    const bool mode1 = (cur.w[0] >> 30) & 1;
    const int st_t = mode1 ? 1 : OW, st_k = mode1 ? OW : 1;
    float *sp = sbuf + k * st_k;
    const uint32_t cbase = pbase + (((uint32_t)cur.la & 0xffffu) << 4);
    const uint32_t cbase2 = cbase + (uint32_t)((cur.la >> 16) << 4);
    const u64 wa2 = pack2(cur.wa, cur.wa);
    F4 s[8];
    // LOAD(T): the predicated tap fetches of walk position T; MATH(T): its three lerps per channel pair
#define RLOD_LOAD2(T)                                                                        \
  {                                                                                          \
    const uint32_t oa = (uint32_t)cur.w[T] & 0x1fff0u, ob = ((uint32_t)cur.w[T] >> 13) & 0x1fff0u; \
    walk_load2(taps, cbase + oa, cbase2 + oa, cbase + ob, cbase2 + ob, cur.w[T]);            \
  }
#define RLOD_MATH2(T)                                                                        \
  {                                                                                          \
    const u64 r2 = pack2(cur.rt[T], cur.rt[T]);                                              \
    s[T].a = lerp2(lerp2(taps.x0, taps.y0, wa2), lerp2(taps.z0, taps.w0, wa2), r2);          \
    s[T].b = lerp2(lerp2(taps.x1, taps.y1, wa2), lerp2(taps.z1, taps.w1, wa2), r2);          \
  }
#define RLOD_EMIT2(T)                                                                        \
  if (POOL == RLOD_POOL_NONE) {                                                              \
    float *q = sp + (T) * st_t;                                                              \
    const float2 lo = unpack2(s[T].a), hi = unpack2(s[T].b);                                 \
    q[0] = lo.x, q[OHW] = lo.y, q[2 * OHW] = hi.x, q[3 * OHW] = hi.y;                        \
  } else if ((T) < 7) {                                                                      \
    u64 va, vb;                                                                              \
    if (POOL == RLOD_POOL_AVG) {                                                             \
      va = add2(s[T].a, s[(T) < 7 ? (T) + 1 : 7].a), vb = add2(s[T].b, s[(T) < 7 ? (T) + 1 : 7].b); \
    } else {                                                                                 \
      va = max2(s[T].a, s[(T) < 7 ? (T) + 1 : 7].a), vb = max2(s[T].b, s[(T) < 7 ? (T) + 1 : 7].b); \
    }                                                                                        \
    const u64 na = shfl_down8(va), nb = shfl_down8(vb);                                      \
    const u64 oa = POOL == RLOD_POOL_AVG ? add2(va, na) : max2(va, na);                      \
    const u64 ob = POOL == RLOD_POOL_AVG ? add2(vb, nb) : max2(vb, nb);                      \
    if (kst) {                                                                               \
      float *q = sp + (T) * st_t;                                                            \
      const float2 lo = unpack2(oa), hi = unpack2(ob);                                       \
      q[0] = lo.x, q[OHW] = lo.y, q[2 * OHW] = hi.x, q[3 * OHW] = hi.y;                      \
    }                                                                                        \
  }
    // software pipeline: the loads of position T+1 are issued right after the math of T has read the taps, and
    // the pooling / stores of position T-1 sit in their shadow
    RLOD_LOAD2(0)
    RLOD_MATH2(0) RLOD_LOAD2(1)
    // the buffer about to be written was handed to the bulk-copy engine two iterations ago
    if (it >= 2) {
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    RLOD_MATH2(1) RLOD_LOAD2(2) RLOD_EMIT2(0)
    RLOD_MATH2(2) RLOD_LOAD2(3) RLOD_EMIT2(1)
    RLOD_MATH2(3) RLOD_LOAD2(4) RLOD_EMIT2(2)
    RLOD_MATH2(4) RLOD_LOAD2(5) RLOD_EMIT2(3)
    RLOD_MATH2(5) RLOD_LOAD2(6) RLOD_EMIT2(4)
    RLOD_MATH2(6) RLOD_LOAD2(7) RLOD_EMIT2(5)
    RLOD_MATH2(7) RLOD_EMIT2(6) RLOD_EMIT2(7)
#undef RLOD_LOAD2
#undef RLOD_MATH2
#undef RLOD_EMIT2
    fence_async_smem();
    __syncwarp();
    // lane 0 hands the four staged blocks to the bulk-copy engine (one issuing lane: the compiler serialises
    // divergent UBLKCP issuers through a vote loop, ~12 instructions per issuer), marked evict-first in L2
    {
      const int r1_ = __shfl_sync(0xffffffffu, r, 8), r2_ = __shfl_sync(0xffffffffu, r, 16), r3_ = __shfl_sync(0xffffffffu, r, 24);
      if (lane == 0) {
        const uint32_t sb = smem_u32(sbuf);
        const int rr[4] = {r, r1_, r2_, r3_};
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (rr[q] >= 0)
            bulk_s2g_hint_nocommit(out_chunk + (size_t)rr[q] * roi_pitch, sb + q * (SLOT * 4), (uint32_t)(STG * sizeof(float)), pol_out);
        bulk_commit();
      }
    }
  };
  using P0 = std::integral_constant<int, 0>;
  using P1 = std::integral_constant<int, 1>;
  for (int g = g_lo + warp, it = 0; g < n_groups;) {
    const int rc = roi_of(g + 2 * kWalkWarps);
    body(recA, ra, recB, rb, it, P0{});
    g += kWalkWarps, ++it;
    if (g >= n_groups) break;
    const int rd = roi_of(g + 2 * kWalkWarps);
    body(recB, rb, recA, rc, it, P1{});
    g += kWalkWarps, ++it;
    ra = rc, rb = rd;
  }
  if (lane == 0) bulk_wait_read<0>();  // shared memory must outlive the engine's reads
}

// ----------------------------------------------------------------------------------------
// generic backward (atomics): NONE / AVG one thread per sample point, MAX one thread per
// pooled output (argmax recomputed from feat: first maximum in row-major window order, the
// rule of ATen's max_pool2d backward).
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void scatter_sample(float *__restrict__ plane, int W, int hs, float hr,
                                               int ws, float wr, float g) {
  if (hs < 0 || ws < 0) return;
  float *p = plane + (size_t)hs * W + ws;
  const float g0 = (1.f - hr) * g, g1 = hr * g;
  atomicAdd(p, g0 * (1.f - wr));
  atomicAdd(p + 1, g0 * wr);
  atomicAdd(p + W, g1 * (1.f - wr));
  atomicAdd(p + W + 1, g1 * wr);
}

template <int POOL>
__global__ void __launch_bounds__(256)
    k_align_bwd_generic(const float *__restrict__ gout, const float *__restrict__ feat,
                        const int *__restrict__ plan, const int *__restrict__ roi_b, int C, int H,
                        int W, int GH, int GW, long long total, float *__restrict__ gin) {
  const int ah = POOL == RLOD_POOL_NONE ? GH : GH - 1, aw = POOL == RLOD_POOL_NONE ? GW : GW - 1;
  const int words = 2 * GH + 2 * GW;
  // iteration space: samples (GH x GW) for NONE/AVG, pooled outputs (ah x aw) for MAX
  const int IH = POOL == RLOD_POOL_MAX ? ah : GH, IW = POOL == RLOD_POOL_MAX ? aw : GW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % IW);
    const int i = (int)((idx / IW) % IH);
    const int c = (int)((idx / ((long long)IW * IH)) % C);
    const int r = (int)(idx / ((long long)IW * IH * C));
    const int *pl = plan + (size_t)r * words;
    const size_t poff = ((size_t)roi_b[r] * C + c) * ((size_t)H * W);
    const float *g = gout + ((size_t)r * C + c) * ((size_t)ah * aw);
    if (POOL == RLOD_POOL_MAX) {
      const float *plane = feat + poff;
      int hsv[2], wsv[2];
      float hrv[2], wrv[2];
      for (int d = 0; d < 2; ++d) {
        hsv[d] = pl[i + d];
        hrv[d] = __int_as_float(pl[GH + i + d]);
        wsv[d] = pl[2 * GH + 2 * (j + d)];
        wrv[d] = __int_as_float(pl[2 * GH + 2 * (j + d) + 1]);
      }
      int best = 0;
      float bv = sample_plane(plane, W, hsv[0], hrv[0], wsv[0], wrv[0]);
      for (int q = 1; q < 4; ++q) {
        const float v = sample_plane(plane, W, hsv[q >> 1], hrv[q >> 1], wsv[q & 1], wrv[q & 1]);
        if (v > bv) {
          bv = v;
          best = q;
        }
      }
      scatter_sample(gin + poff, W, hsv[best >> 1], hrv[best >> 1], wsv[best & 1], wrv[best & 1],
                     g[i * aw + j]);
    } else {
      float gs;
      if (POOL == RLOD_POOL_NONE) {
        gs = g[i * aw + j];
      } else {
        float acc = 0.f;
        for (int di = -1; di <= 0; ++di)
          for (int dj = -1; dj <= 0; ++dj) {
            const int oi = i + di, oj = j + dj;
            if (oi >= 0 && oi < ah && oj >= 0 && oj < aw) acc += g[oi * aw + oj] * 0.25f;
          }
        gs = acc;
      }
      scatter_sample(gin + poff, W, pl[i], __int_as_float(pl[GH + i]), pl[2 * GH + 2 * j],
                     __int_as_float(pl[2 * GH + 2 * j + 1]), gs);
    }
  }
}

// ----------------------------------------------------------------------------------------
// fast backward ("scatter walk"): the transpose of k_align8_fwd_walk.  CTA = (image, 4
// channels); the 4 gradient planes are accumulated in shared memory (same odd-pitch
// interleaved layout) and written to HBM once, coalesced -- no global atomics, no memset.
// A warp serves 4 rois per iteration (8 lanes each, 4 channels per lane):
//   1. the rois' grad_out tiles (4 channels x OHW floats = one contiguous 784-byte run each)
//      arrive by bulk async copy (TMA engine), double-buffered per warp behind mbarriers;
//   2. the gradient of every sample point is rebuilt in registers (AVG: each pooled gradient
//      / 4 goes to its 2x2 sample window);
//   3. the sample rows are walked with the forward's two line slots: a slot ACCUMULATES the
//      line's gradient while consecutive sample rows share it and is flushed when the forward
//      would have loaded a new line into it.  A flush spreads the 8 sample columns over their
//      two taps (LDS.128 / 4 FFMA / STS.128 per tap and lane) under a per-row spin lock in
//      shared memory, because other warps (and the other three rois of this warp) may hold a
//      line of the same row; lanes of one roi that share a column take turns (rank rounds).
// Rounding order across rois is therefore not fixed (the reference's atomics are not either).
// ----------------------------------------------------------------------------------------
// plane[p] += w * v for 4 channels, shared-space LDS.128 / FFMA / STS.128, only where `on`
__device__ __forceinline__ void smem_axpy4_if(uint32_t p, float w, float4 v, bool on) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      ".reg .f32 a, b, c, d;\n"
      "setp.ne.b32 q, %6, 0;\n"
      "@q ld.shared.v4.f32 {a, b, c, d}, [%0];\n"
      "@q fma.rn.f32 a, %1, %2, a;\n"
      "@q fma.rn.f32 b, %1, %3, b;\n"
      "@q fma.rn.f32 c, %1, %4, c;\n"
      "@q fma.rn.f32 d, %1, %5, d;\n"
      "@q st.shared.v4.f32 [%0], {a, b, c, d};\n"
      "}\n" ::"r"(p),
      "f"(w), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)on)
      : "memory");
}

// One line of one roi per 8-lane group: v (4 channels) of every lane goes to its two taps
// p0 / p1 with weights (1 - w1) / w1, under the spin lock of the line's row.  Deliberately
// NOT inlined: the walk has 18 flush points and the instruction cache matters more than the
// call.  Memory order inside a warp is program order, so a lane's second tap may be another
// lane's first tap; lanes that share a column within one request take turns by rank.
__device__ __noinline__ void row_flush(float4 v, int need, uint32_t p0, uint32_t p1, float w1, int rank,
                                       int maxrank_warp, uint32_t lock_addr, int k) {
  const unsigned full = 0xffffffffu;
  const float w0 = 1.f - w1;
  while (__any_sync(full, need)) {
    int got = 0;
    if (need && k == 0) {
      asm volatile("atom.acquire.cta.shared.cas.b32 %0, [%1], 0, 1;" : "=r"(got) : "r"(lock_addr) : "memory");
      got = got == 0;
    }
    got = __shfl_sync(full, got, 0, 8);
    const bool go = need && got;
    __syncwarp();  // the leader's acquire is ordered before every lane's plane accesses
    if (maxrank_warp == 0) {
      smem_axpy4_if(p0, w0, v, go);
      __syncwarp();  // a lane's second tap may be its neighbour's first
      smem_axpy4_if(p1, w1, v, go);
    } else {
      for (int rnd = 0; rnd <= maxrank_warp; ++rnd) {
        smem_axpy4_if(p0, w0, v, go && rank == rnd);
        __syncwarp();
        smem_axpy4_if(p1, w1, v, go && rank == rnd);
        __syncwarp();
      }
    }
    __syncwarp();  // every lane's updates happen before the leader's release
    if (go && k == 0) asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(lock_addr), "r"(0) : "memory");
    need = need && !got;
  }
}

template <int POOL>
__global__ void __launch_bounds__(kWalkThreads, 2)
    k_align8_bwd_walk(const float *__restrict__ gout, const int *__restrict__ ext,
                      const int *__restrict__ order, const int *__restrict__ img_off, int C,
                      int H, int W, int P, int n_chunks, int accumulate, int lock_mul,
                      float *__restrict__ gin) {
  constexpr int OW = POOL == RLOD_POOL_NONE ? 8 : 7;
  constexpr int OHW = OW * OW;
  constexpr int STG = 4 * OHW;
  constexpr int SLOT = POOL == RLOD_POOL_NONE ? 264 : 200;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *planes4 = reinterpret_cast<float4 *>(smem_raw);
  const int HW = H * W;
  float *stage = reinterpret_cast<float *>(planes4 + (H + 2) * P);
  int *locks = reinterpret_cast<int *>(stage + kWalkWarps * 2 * 4 * SLOT);
  uint64_t *bars = reinterpret_cast<uint64_t *>(locks + ((H + 2 + 3) & ~3));  // [kWalkWarps][2]
  int *s_next = reinterpret_cast<int *>(bars + 2 * kWalkWarps);

  const int b = blockIdx.x / n_chunks, chunk = blockIdx.x - b * n_chunks;
  const int r0 = img_off[b], r1 = img_off[b + 1];
  float *dst = gin + ((size_t)b * C + (size_t)chunk * 4) * HW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = lane & 7, slot = lane >> 3;

  // planes: zero, or the caller's gradient when accumulating
  for (int p = threadIdx.x; p < (H + 2) * P; p += kWalkThreads) planes4[p] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = threadIdx.x; p < H + 2; p += kWalkThreads) locks[p] = 0;
  if (threadIdx.x == 0) *s_next = kWalkWarps;
  if (lane == 0) {
    mbar_init(bars + 2 * warp, 1);
    mbar_init(bars + 2 * warp + 1, 1);
  }
  __syncthreads();
  if (accumulate) {
    for (int e = threadIdx.x; e < 4 * HW; e += kWalkThreads) {
      const int p = e >> 2, c = e & 3;
      const int y = p / W, x = p - y * W;
      reinterpret_cast<float *>(planes4 + y * P + x)[c] = __ldg(dst + (size_t)c * HW + p);
    }
    __syncthreads();
  }

  const uint32_t pbase = smem_u32(planes4);
  float *stg = stage + (warp * 2) * (4 * SLOT) + slot * SLOT;
  const int n_groups = (r1 - r0 + 3) >> 2;
  auto roi_of = [&](int gg) {
    const int kk = r0 + 4 * gg + slot;
    return (gg < n_groups && kk < r1) ? __ldg(order + kk) : -1;
  };
  // tile loads: lane 0 announces the byte count, every slot leader copies its roi's tile
  auto issue_tiles = [&](int rr, int buf) {
    const unsigned act = __ballot_sync(0xffffffffu, k == 0 && rr >= 0);
    if (lane == 0) mbar_expect_tx(bars + 2 * warp + buf, (uint32_t)(__popc(act) * STG * sizeof(float)));
    __syncwarp();
    if (k == 0 && rr >= 0)
      bulk_g2s(stg + buf * (4 * SLOT), gout + ((size_t)rr * C + (size_t)chunk * 4) * OHW,
               (uint32_t)(STG * sizeof(float)), bars + 2 * warp + buf);
  };

  // records are pulled into L1 one group ahead (the iteration is long and the flush calls want
  // the registers a register-held prefetch would take)
  auto prefetch_rec = [&](int rr) {
    if (rr >= 0 && k == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(ext + (size_t)rr * 32));
  };
  // groups are claimed from a shared counter (the cost of a group varies with lock retries and
  // rank rounds), two ahead: one whose tiles are in flight, one whose roi ids are being read
  auto claim = [&]() {
    int v = 0;
    if (lane == 0) v = atomicAdd(s_next, 1);
    return __shfl_sync(0xffffffffu, v, 0);
  };
  int g = warp, gn = claim();
  int r = roi_of(g);
  int rn = roi_of(gn);
  prefetch_rec(r);
  if (g < n_groups) issue_tiles(r, 0);

  for (int it = 0; g < n_groups; ++it) {
    WalkRec cur;
    walk_rec_clear(cur);
    if (r >= 0) walk_rec_load(ext, r, k, cur);
    prefetch_rec(rn);
    const int gnn = claim();
    const int rnn = roi_of(gnn);
    const int buf = it & 1;
    // next group's tiles go into the other buffer: its last reads (previous iteration) are
    // ordered before this point by the warp barriers of the flushes
    if (gn < n_groups) issue_tiles(rn, buf ^ 1);
    mbar_wait(bars + 2 * warp + buf, (uint32_t)((it >> 1) & 1));

    // ---- sample gradients of this lane's column, 4 channels, one sample row at a time --------
    // (AVG: pooled gradient / 4 to its 2x2 sample window = 0.25 * (c[t-1] + c[t]), c[i] = column
    // pair sum g[i][k-1] + g[i][k] of pooled row i)
    const float *tile = stg + buf * (4 * SLOT);
    float4 cprev = make_float4(0.f, 0.f, 0.f, 0.f);
    auto sample_grad = [&](int t) -> float4 {
      float4 G;
      if (POOL == RLOD_POOL_NONE) {
        const float *q = tile + t * 8 + k;
        G = make_float4(q[0], q[OHW], q[2 * OHW], q[3 * OHW]);
      } else {
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < 7) {
          const float *q = tile + t * 7 + k;
          if (k <= 6) c = make_float4(q[0], q[OHW], q[2 * OHW], q[3 * OHW]);
          if (k >= 1) c.x += q[-1], c.y += q[OHW - 1], c.z += q[2 * OHW - 1], c.w += q[3 * OHW - 1];
        }
        G = make_float4((cprev.x + c.x) * 0.25f, (cprev.y + c.y) * 0.25f, (cprev.z + c.z) * 0.25f,
                        (cprev.w + c.w) * 0.25f);
        cprev = c;
      }
      if (r < 0) G = make_float4(0.f, 0.f, 0.f, 0.f);
      return G;
    };

    // ---- scatter walk ---------------------------------------------------------------------
    const uint32_t lane_first = ((uint32_t)cur.la & 0xffffu) << 4;
    const uint32_t lane_da = (((uint32_t)cur.la >> 16) & 1u) << 4;
    const int rank = (cur.la >> 17) & 7;
    const int maxrank_warp = __reduce_max_sync(0xffffffffu, (cur.la >> 20) & 7);
    float4 da = make_float4(0.f, 0.f, 0.f, 0.f), db = da;
    const uint32_t lock_base = smem_u32(locks);
    const float wa = cur.wa;
    auto flush = [&](float4 v, bool need, uint32_t line_bytes) {
      const uint32_t row = ((line_bytes >> 4) * (uint32_t)lock_mul) >> 16;
      const uint32_t p0 = pbase + line_bytes + lane_first;
      row_flush(v, (int)need, p0, p0 + lane_da, wa, rank, maxrank_warp, lock_base + 4u * row, k);
    };
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int w = cur.w[t];
      if (t > 0) {
        const int wp = cur.w[t - 1];
        flush(da, (w & 1) && r >= 0, (uint32_t)wp & 0x1fff0u);
        flush(db, (w & 2) && r >= 0, ((uint32_t)wp >> 13) & 0x1fff0u);
      }
      if (w & 1) da = make_float4(0.f, 0.f, 0.f, 0.f);
      if (w & 2) db = make_float4(0.f, 0.f, 0.f, 0.f);
      const float rb = cur.rt[t], ra = 1.f - rb;
      const float4 G = sample_grad(t);
      da.x = fmaf(ra, G.x, da.x), da.y = fmaf(ra, G.y, da.y), da.z = fmaf(ra, G.z, da.z), da.w = fmaf(ra, G.w, da.w);
      db.x = fmaf(rb, G.x, db.x), db.y = fmaf(rb, G.y, db.y), db.z = fmaf(rb, G.z, db.z), db.w = fmaf(rb, G.w, db.w);
    }
    flush(da, r >= 0, (uint32_t)cur.w[7] & 0x1fff0u);
    flush(db, r >= 0, ((uint32_t)cur.w[7] >> 13) & 0x1fff0u);
    g = gn, gn = gnn;
    r = rn, rn = rnn;
  }
  __syncthreads();
  // planes -> HBM, coalesced per channel plane
  for (int cy = warp; cy < 4 * H; cy += kWalkWarps) {  // a warp takes (channel, row) runs: no division per element
    const int c = cy / H, y = cy - c * H;
    const float *src = reinterpret_cast<const float *>(planes4 + y * P) + c;
    float *d = dst + (size_t)c * HW + (size_t)y * W;
    for (int x = lane; x < W; x += 32) d[x] = src[4 * x];
  }
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
static int check_align_args(const float *rois, int B, int C, int H, int W, int R, int ah, int aw,
                            int pool_mode) {
  if (B < 0 || C < 0 || R < 0 || ah < 1 || aw < 1) return RLOD_EINVAL;
  if (pool_mode < RLOD_POOL_NONE || pool_mode > RLOD_POOL_MAX) return RLOD_EINVAL;
  if (H < 2 || W < 2) return RLOD_EINVAL;  // the reference reads row/col -1 there (:48-49)
  if (pool_mode == RLOD_POOL_NONE && (ah < 2 || aw < 2)) return RLOD_EINVAL;  // bin = x / 0
  if (R > 0 && rois == nullptr) return RLOD_EINVAL;
  if ((long long)B * C * H * W >= (1LL << 31)) return RLOD_EUNSUPPORTED;
  return RLOD_OK;
}

static int build_plan(const float *rois, int B, int H, int W, int R, int GH, int GW, float scale,
                      const AlignWs &ws, cudaStream_t st) {
  cudaMemsetAsync(ws.flag, 0, 4 * sizeof(int), st);
  const long long n = (long long)R * (GH + GW);
  RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st, k_roi_plan<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(rois, R, B, H, W, GH, GW, scale, ws));
  RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st, k_roi_group_fixup<<<1, 32, 0, st>>>(R, B, ws));
  return launch_status();
}

static size_t fwd_walk_smem(int H, int W, int pool_mode) {
  const int slot = pool_mode == RLOD_POOL_NONE ? 264 : 200;
  return (size_t)16 * (size_t)(H + 2) * walk_pitch(W) + (size_t)kWalkWarps * 2 * 4 * slot * 4;
}

// Can the forward kernel fill its planes with bulk copies + in-place interleave?  Needs 8-byte
// aligned planes (H*W even), three planar planes inside the staging area, and no batch of the
// interleave writing over plane-0 pixels that are still unread (see the kernel).
static bool fwd_tma_fill_ok(const float *feat, int H, int W, int pool_mode) {
  const int HW = H * W, P = walk_pitch(W);
  if ((HW & 1) || ((uintptr_t)feat % 16) != 0) return false;
  const size_t plane_copy = ((size_t)HW * 4 + 8 + 15) & ~(size_t)15;
  const int slot = pool_mode == RLOD_POOL_NONE ? 264 : 200;
  if (3 * plane_copy > (size_t)kWalkWarps * 2 * 4 * slot * 4) return false;
  const size_t region = (size_t)16 * (H + 2) * P;
  if (plane_copy > region) return false;
  const size_t tail = region - plane_copy;  // plane 0 (copy start; its pixel q is at >= tail + 4q)
  const int batch = kWalkThreads * 4;
  for (int base = 0; base < HW; base += batch) {
    const int last = (base + batch < HW ? base + batch : HW) - 1;  // last pixel written by this batch
    const size_t wr_end = (size_t)16 * ((size_t)(last / W) * P + last % W + 1);
    if (wr_end > tail + (size_t)4 * (last + 1)) return false;  // would hit an unread plane-0 pixel
  }
  return true;
}

}  // namespace rlod

using namespace rlod;

RLOD_API size_t rlod_roi_align_workspace_bytes(int B, int R, int ah, int aw, int pool_mode) {
  const int GH = pool_mode == RLOD_POOL_NONE ? ah : ah + 1;
  const int GW = pool_mode == RLOD_POOL_NONE ? aw : aw + 1;
  if (B < 0 || R < 0 || ah < 1 || aw < 1) return 0;
  return carve_align_ws(nullptr, B, R, GH, GW).bytes;
}

// The forward in two phases, so that a caller may plan the rois on another stream while the previous pooling is
// still running (rlod_roi_align_plan + rlod_roi_align_forward_planned); rlod_roi_align_forward is the two in a row.
// Which kernels run depends on the geometry only (pointer alignment is a precondition of the plane kernel that the
// planned entry point checks).
namespace rlod {

static bool align_fwd_geometry_fast(int B, int C, int H, int W, int R, int GH, int GW, int pool_mode) {
  const size_t smem = fwd_walk_smem(H, W, pool_mode) + 16;  // + the fill mbarrier
  return GH == 8 && GW == 8 && (C % 4) == 0 && smem <= (size_t)kMaxSmemPerCta && (H + 2) * walk_pitch(W) <= 8192 &&
         R >= 2 * B;
}

// A map too large for the plane kernel (FPN-sized levels, images beyond ~1400 pixels at stride 16) is covered by
// overlapping tiles of at most kTilePixels padded pixels (two CTAs per SM, like the 50 x 75 map of C4), origins
// half a tile apart: a roi of up to half a tile per side fits one tile wherever it lies, larger ones are left to
// the generic kernel.  Returns false when the map needs no tiles, cannot be tiled (more than kMaxTilesPerImage
// tiles), has too few rois per tile to pay for the tiles' planes, or the call is not the plane kernel's anyway.
constexpr int kTilePixels = 4030;
static bool align_fwd_tiles(int B, int C, int H, int W, int R, int GH, int GW, int pool_mode, int channels_last,
                            TileGrid *tg) {
  static const bool off = getenv("RLOD_NO_TILES") != nullptr;  // A/B switch: the generic kernel for large maps
  if (off || channels_last || GH != 8 || GW != 8 || (C % 4) != 0 || R < 2 * B) return false;
  if (align_fwd_geometry_fast(B, C, H, W, R, GH, GW, pool_mode)) return false;
  // as square as the map allows: an axis shorter than the square's side is not tiled
  int tw = W < 61 ? W : 61, th = kTilePixels / walk_pitch(tw) - 2;
  if (th >= H) {
    th = H;
    while (tw < W && (th + 2) * walk_pitch(tw + 1) <= kTilePixels) ++tw;
  }
  if (th < 8 || tw < 8) return false;
  tg->th = th, tg->tw = tw;
  // (a map whose width is a multiple of 4: column origins at multiples of 4 pixels, so that the window rows
  // start 16-byte aligned and are filled with 128-bit loads -- FPN P2, 8.5 rois per tile: 253 -> 220 us; other
  // widths keep the half-tile distance: rounding it down costs 2 x 1024 x 100 x 150 a fifth column of tiles, 216 ->
  // 245 us)
  tg->sy = th >= H ? th : th / 2, tg->sx = tw >= W ? tw : ((W % 4) == 0 ? ((tw / 2) & ~3) : tw / 2);
  tg->ny = th >= H ? 1 : (H - th + tg->sy - 1) / tg->sy + 1;
  tg->nx = tw >= W ? 1 : (W - tw + tg->sx - 1) / tg->sx + 1;
  if (tg->ny * tg->nx > kMaxTilesPerImage || tg->ny * tg->nx < 2) return false;
  // A (tile, 4 channels) CTA costs ~9.3 us + 0.06 us per roi (the planes arrive by 4-byte async copies), the
  // generic kernel ~0.72 us per (roi, 4 channels); measured with tools/time_op.py: an FPN P2 level (2 x 256 x 200 x
  // 304, 8.5 rois per tile) 253 us tiled against 172 us generic, 2 x 1024 x 100 x 150 with 25 rois per tile 216
  // against 371 us, with 166 per tile 398 against 2 353 us.  Tiles from 16 rois per tile on.
  static const bool force = getenv("RLOD_FORCE_TILES") != nullptr;
  if (!force && (long long)R < 16LL * B * tg->ny * tg->nx) return false;
  if ((long long)B * tg->ny * tg->nx * (C / 4) >= (1LL << 31)) return false;
  return fwd_walk_smem(th, tw, pool_mode) + 16 <= (size_t)kMaxSmemPerCta;
}

// A/B switches: RLOD_NO_PDL=1 (roi_lists.cuh) launches the list and pooling kernels of the forward strictly one
// after the other; RLOD_NO_SPLIT=1 keeps one CTA per (image, 4 channels) in the last wave of the pooling launch.

// How the pooling launch ends: the CTAs have equal cost, so the launch ends with a partial wave.  Each item of
// that wave can be served by S CTAs instead, every one with its own copy of the planes and a contiguous share of
// the image's roi groups.  Measured on the merged pooling call of the C4 step (tools/ab_split.sh; S = 1 / 2 / 3,
// a wave counted as one CTA per SM = 148 | as the 296 resident CTAs): 3 images per rank 130.0 / 121.8 / 119.8 |
// 130.0 / 123.9 / 123.9 us, 6 images 224.3 / 216.0 / 220.1 us (both), 12 images 405.5 / 406.5 / 406.5 us (both);
// C2 78.6 / 76.8 us.  Only the very end of the launch matters (the smaller wave is the better unit), and a wave
// model with the fill as a fixed cost does not predict the table -- an unsplit tail runs one CTA per SM, faster
// than two co-resident ones -- so the rule is the measured one: thirds up to 5 waves of 148, halves up to 12,
// while the parts keep ~16 roi groups.  Returns S (1 = no split) and the first split item.
static int fwd_tail_split(int items, int rois_per_image, int *split_from) {
  static const bool off = getenv("RLOD_NO_SPLIT") != nullptr;
  static const int force = getenv("RLOD_SPLIT") ? atoi(getenv("RLOD_SPLIT")) : 0;  // A/B: this S whenever rem > 0
  static const int env_slots = getenv("RLOD_SPLIT_SLOTS") ? atoi(getenv("RLOD_SPLIT_SLOTS")) : 0;  // A/B: the wave
  const int slots = env_slots > 0 ? env_slots : kSmCount;
  const int waves = items / slots, full = waves * slots, rem = items - full;
  *split_from = items;
  if (off || rem == 0) return 1;
  int S = waves <= 5 ? 3 : (waves <= 12 ? 2 : 1);
  while (S > 1 && rois_per_image / (4 * S) < 16) --S;
  if (force >= 1 && force <= 4) S = force;
  if (S > 1) *split_from = full;
  return S;
}

static const TileGrid kNoTiles = {1, 1, 0, 0, 0, 0};

// tiles: the map is tiled (align_fwd_tiles); fast then means "plane kernel over the tiles"
static int align_fwd_plan(const float *rois, int B, int H, int W, int R, int GH, int GW, float spatial_scale,
                          bool fast, const TileGrid *tiles, const AlignWs &ws, cudaStream_t st) {
  if (!fast) return build_plan(rois, B, H, W, R, GH, GW, spatial_scale, ws, st);
  // one launch: sample grid + orientation / tap-order search + walk record + roi lists
  cudaMemsetAsync(ws.flag, 0, 4 * sizeof(int), st);
  if (tiles) {
    // lists per (image, tile); rois that fit no tile get a k_roi_plan record for the generic kernel
    RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st,
                k_roi_plan8_walk<<<(unsigned)cdiv(R, 4), 128, 0, st>>>(rois, R, B, H, W, walk_pitch(tiles->tw), spatial_scale,
                                                                      0, *tiles, ws));
    const int V = B * tiles->ny * tiles->nx;
    RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st,
                launch_after(k_roi_lists_finish, dim3((unsigned)V), dim3(kOrderThreads), 0, st, pdl_enabled(),
                             (const int *)ws.ext, 0, 30, R, V, ws));
    return launch_status();
  }
  RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st,
              k_roi_plan8_walk<<<(unsigned)cdiv(R, 4), 128, 0, st>>>(rois, R, B, H, W, walk_pitch(W), spatial_scale, 0,
                                                                    kNoTiles, ws));
  static const bool v1 = getenv("RLOD_FWD_V1") != nullptr;  // A/B switch: the scalar walk of round 1
  if (v1) {
    RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st, k_roi_group_fixup<<<1, 32, 0, st>>>(R, B, ws));
  } else {
    // group fix-up + partition of every image's list by walk mode (the four rois of a warp then stage alike)
    // in one launch
    // (set up behind the plan kernel as its programmatic dependent: its CTAs are resident and past their
    // prologue when the plan completes)
    RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st,
                launch_after(k_roi_lists_finish, dim3((unsigned)B), dim3(kOrderThreads), 0, st, pdl_enabled(),
                             (const int *)ws.ext, 0, 30, R, B, ws));
  }
  return launch_status();
}

static int align_fwd_run(const float *feat, int B, int C, int H, int W, int R, int ah, int aw, int GH, int GW,
                         int pool_mode, int channels_last, bool fast, const TileGrid *tiles, float *out,
                         const AlignWs &ws, cudaStream_t st, bool behind_plan) {
  // behind_plan: the plan kernels are the launches just before this one on st (rlod_roi_align_forward); the
  // pooling kernel is then set up behind them and fills its planes while they run
  if (fast && tiles) {
    // plane kernel over the tiles (every (image, tile) is a virtual image with its own roi list), then the rois
    // that fit no tile through the generic kernel
    const int th = tiles->th, tw = tiles->tw, V = B * tiles->ny * tiles->nx, n_chunks = C / 4;
    const size_t smem = fwd_walk_smem(th, tw, pool_mode) + 16;
    const unsigned grid = (unsigned)(V * n_chunks);
    static const bool no_vec = getenv("RLOD_TILE_FILL_V1") != nullptr;  // A/B switch: 4-byte async copies
    const int vec_fill = (!no_vec && (W % 4) == 0 && ((uintptr_t)feat % 16) == 0 && (tiles->sx % 4) == 0) ? 3 : 0;
#define RLOD_LAUNCH_FWD_TILED(POOL)                                                                          \
  do {                                                                                                       \
    static bool attr_set = false;                                                                            \
    if (!attr_set) {                                                                                         \
      cudaFuncSetAttribute(k_align8_fwd_walk2<POOL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                           kMaxSmemPerCta);                                                                  \
      attr_set = true;                                                                                       \
    }                                                                                                        \
    ProfScope _ps(RLOD_KERNEL_ALIGN_FWD, st);                                                                \
    launch_after(k_align8_fwd_walk2<POOL, true>, dim3(grid), dim3(kWalkThreads), smem, st,                   \
                 behind_plan && pdl_enabled(), feat, (const int *)ws.ext, (const int *)ws.order2,            \
                 (const int *)ws.img_off, C, th, tw, walk_pitch(tw), n_chunks, vec_fill, (int)grid, 1, *tiles, H, W, out); \
  } while (0)
    if (pool_mode == RLOD_POOL_NONE)
      RLOD_LAUNCH_FWD_TILED(RLOD_POOL_NONE);
    else if (pool_mode == RLOD_POOL_AVG)
      RLOD_LAUNCH_FWD_TILED(RLOD_POOL_AVG);
    else
      RLOD_LAUNCH_FWD_TILED(RLOD_POOL_MAX);
#undef RLOD_LAUNCH_FWD_TILED
    // the rois beyond a tile (their number is only known on the device): a fixed grid strides over the list
    const long long total = (long long)R * C * ah * aw;
    const long long want = cdiv(total, 256);
    const unsigned ggrid = (unsigned)(want < 16LL * kSmCount ? want : 16LL * kSmCount);
    if (pool_mode == RLOD_POOL_NONE)
      RLOD_LAUNCH(RLOD_KERNEL_ALIGN_FWD_GENERIC, st, k_align_fwd_generic<RLOD_POOL_NONE>
          <<<ggrid, 256, 0, st>>>(feat, ws.plan, ws.roi_b, C, H, W, GH, GW, total, out, ws.own, ws.flag + 2));
    else if (pool_mode == RLOD_POOL_AVG)
      RLOD_LAUNCH(RLOD_KERNEL_ALIGN_FWD_GENERIC, st, k_align_fwd_generic<RLOD_POOL_AVG>
          <<<ggrid, 256, 0, st>>>(feat, ws.plan, ws.roi_b, C, H, W, GH, GW, total, out, ws.own, ws.flag + 2));
    else
      RLOD_LAUNCH(RLOD_KERNEL_ALIGN_FWD_GENERIC, st, k_align_fwd_generic<RLOD_POOL_MAX>
          <<<ggrid, 256, 0, st>>>(feat, ws.plan, ws.roi_b, C, H, W, GH, GW, total, out, ws.own, ws.flag + 2));
    return launch_status();
  }
  if (fast) {
    static const bool v1 = getenv("RLOD_FWD_V1") != nullptr;
    if (channels_last && v1) return RLOD_EUNSUPPORTED;
    const size_t smem = fwd_walk_smem(H, W, pool_mode) + 16;
    const int tma_fill = channels_last ? 2 : (fwd_tma_fill_ok(feat, H, W, pool_mode) ? 1 : 0);
    const int n_chunks = C / 4;
    const int P = walk_pitch(W);
    int split_from = B * n_chunks;
    const int split = v1 ? 1 : fwd_tail_split(B * n_chunks, R / B, &split_from);
    const unsigned grid = (unsigned)(split_from + (B * n_chunks - split_from) * split);
#define RLOD_LAUNCH_FWD(POOL)                                                                  \
  do {                                                                                         \
    static bool attr_set = false;                                                              \
    if (!attr_set) {                                                                           \
      cudaFuncSetAttribute(k_align8_fwd_walk<POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           kMaxSmemPerCta);                                                    \
      cudaFuncSetAttribute(k_align8_fwd_walk2<POOL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           kMaxSmemPerCta);                                                    \
      attr_set = true;                                                                         \
    }                                                                                          \
    ProfScope _ps(RLOD_KERNEL_ALIGN_FWD, st);                                                  \
    if (v1)                                                                                    \
      k_align8_fwd_walk<POOL><<<grid, kWalkThreads, smem, st>>>(feat, ws.ext, ws.order,        \
                                                                ws.img_off, C, H, W, P,        \
                                                                n_chunks, tma_fill, out);      \
    else                                                                                       \
      launch_after(k_align8_fwd_walk2<POOL, false>, dim3(grid), dim3(kWalkThreads), smem, st, \
                   behind_plan && pdl_enabled(), feat, (const int *)ws.ext,                    \
                   (const int *)ws.order2, (const int *)ws.img_off, C, H, W, P, n_chunks,      \
                   tma_fill, split_from, split, kNoTiles, H, W, out);                          \
  } while (0)
    if (pool_mode == RLOD_POOL_NONE)
      RLOD_LAUNCH_FWD(RLOD_POOL_NONE);
    else if (pool_mode == RLOD_POOL_AVG)
      RLOD_LAUNCH_FWD(RLOD_POOL_AVG);
    else
      RLOD_LAUNCH_FWD(RLOD_POOL_MAX);
#undef RLOD_LAUNCH_FWD
    return launch_status();
  }
  const long long total = (long long)R * C * ah * aw;
  const unsigned grid = (unsigned)(cdiv(total, 256) < (1LL << 30) ? cdiv(total, 256) : (1LL << 30));
  if (pool_mode == RLOD_POOL_NONE)
    RLOD_LAUNCH(RLOD_KERNEL_ALIGN_FWD_GENERIC, st, k_align_fwd_generic<RLOD_POOL_NONE>
        <<<grid, 256, 0, st>>>(feat, ws.plan, ws.roi_b, C, H, W, GH, GW, total, out));
  else if (pool_mode == RLOD_POOL_AVG)
    RLOD_LAUNCH(RLOD_KERNEL_ALIGN_FWD_GENERIC, st, k_align_fwd_generic<RLOD_POOL_AVG>
        <<<grid, 256, 0, st>>>(feat, ws.plan, ws.roi_b, C, H, W, GH, GW, total, out));
  else
    RLOD_LAUNCH(RLOD_KERNEL_ALIGN_FWD_GENERIC, st, k_align_fwd_generic<RLOD_POOL_MAX>
        <<<grid, 256, 0, st>>>(feat, ws.plan, ws.roi_b, C, H, W, GH, GW, total, out));
  return launch_status();
}

}  // namespace rlod

RLOD_API int rlod_roi_align_forward(const float *feat, const float *rois, int B, int C, int H,
                                    int W, int R, int ah, int aw, float spatial_scale,
                                    int pool_mode, int channels_last, float *out, void *workspace,
                                    size_t workspace_bytes, rlod_stream_t stream) {
  int rc = check_align_args(rois, B, C, H, W, R, ah, aw, pool_mode);
  if (rc != RLOD_OK) return rc;
  if (R == 0 || C == 0) return RLOD_OK;
  if (!feat || !out || !workspace || B < 1) return RLOD_EINVAL;
  const int GH = pool_mode == RLOD_POOL_NONE ? ah : ah + 1;
  const int GW = pool_mode == RLOD_POOL_NONE ? aw : aw + 1;
  AlignWs ws = carve_align_ws(workspace, B, R, GH, GW);
  if (workspace_bytes < ws.bytes) return RLOD_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;
  bool fast = align_fwd_geometry_fast(B, C, H, W, R, GH, GW, pool_mode) && ((uintptr_t)out % 16) == 0;
  // a channels-last map is only read by the plane kernel (its taps are contiguous float4s there); the
  // generic kernels index NCHW planes
  if (channels_last && (!fast || ((uintptr_t)feat % 16) != 0)) return RLOD_EUNSUPPORTED;
  TileGrid tg;
  const bool tiled = ((uintptr_t)out % 16) == 0 && align_fwd_tiles(B, C, H, W, R, GH, GW, pool_mode, channels_last, &tg);
  fast = fast || tiled;
  rc = align_fwd_plan(rois, B, H, W, R, GH, GW, spatial_scale, fast, tiled ? &tg : nullptr, ws, st);
  if (rc) return rc;
  return align_fwd_run(feat, B, C, H, W, R, ah, aw, GH, GW, pool_mode, channels_last, fast, tiled ? &tg : nullptr, out, ws,
                       st, true);
}

// Pure host arithmetic (no CUDA call): which kernels rlod_roi_align_forward would launch for this geometry.
RLOD_API int rlod_roi_align_forward_route(int B, int C, int H, int W, int R, int ah, int aw, int pool_mode,
                                          int channels_last, int *info) {
  if (!info) return RLOD_EINVAL;
  for (int i = 0; i < 8; ++i) info[i] = 0;
  if (check_align_args(reinterpret_cast<const float *>(info), B, C, H, W, R, ah, aw, pool_mode) != RLOD_OK || B < 1)
    return RLOD_EINVAL;
  const int GH = pool_mode == RLOD_POOL_NONE ? ah : ah + 1;
  const int GW = pool_mode == RLOD_POOL_NONE ? aw : aw + 1;
  TileGrid tg;
  if (align_fwd_geometry_fast(B, C, H, W, R, GH, GW, pool_mode)) {
    int split_from = B * (C / 4);
    info[0] = 1;
    info[7] = fwd_tail_split(B * (C / 4), R / B, &split_from);
    info[1] = info[2] = 1, info[3] = H, info[4] = W, info[5] = H, info[6] = W;
  } else if (align_fwd_tiles(B, C, H, W, R, GH, GW, pool_mode, channels_last, &tg)) {
    info[0] = 2, info[1] = tg.ny, info[2] = tg.nx, info[3] = tg.th, info[4] = tg.tw, info[5] = tg.sy, info[6] = tg.sx;
    info[7] = 1;
  }
  if (channels_last && info[0] != 1) return RLOD_EUNSUPPORTED;
  return RLOD_OK;
}

RLOD_API int rlod_roi_align_plan(const float *rois, int B, int C, int H, int W, int R, int ah, int aw,
                                 float spatial_scale, int pool_mode, void *workspace, size_t workspace_bytes,
                                 rlod_stream_t stream) {
  int rc = check_align_args(rois, B, C, H, W, R, ah, aw, pool_mode);
  if (rc != RLOD_OK) return rc;
  if (R == 0 || C == 0) return RLOD_OK;
  if (!workspace || B < 1) return RLOD_EINVAL;
  const int GH = pool_mode == RLOD_POOL_NONE ? ah : ah + 1;
  const int GW = pool_mode == RLOD_POOL_NONE ? aw : aw + 1;
  AlignWs ws = carve_align_ws(workspace, B, R, GH, GW);
  if (workspace_bytes < ws.bytes) return RLOD_ENOSPC;
  TileGrid tg;
  const bool tiled = align_fwd_tiles(B, C, H, W, R, GH, GW, pool_mode, 0, &tg);
  return align_fwd_plan(rois, B, H, W, R, GH, GW, spatial_scale,
                        tiled || align_fwd_geometry_fast(B, C, H, W, R, GH, GW, pool_mode), tiled ? &tg : nullptr, ws,
                        (cudaStream_t)stream);
}

RLOD_API int rlod_roi_align_forward_planned(const float *feat, int B, int C, int H, int W, int R, int ah, int aw,
                                            int pool_mode, int channels_last, float *out, void *workspace,
                                            size_t workspace_bytes, rlod_stream_t stream) {
  if (B < 0 || C < 0 || R < 0 || ah < 1 || aw < 1 || H < 2 || W < 2) return RLOD_EINVAL;
  if (pool_mode < RLOD_POOL_NONE || pool_mode > RLOD_POOL_MAX) return RLOD_EINVAL;
  if ((long long)B * C * H * W >= (1LL << 31)) return RLOD_EUNSUPPORTED;
  if (R == 0 || C == 0) return RLOD_OK;
  if (!feat || !out || !workspace || B < 1) return RLOD_EINVAL;
  const int GH = pool_mode == RLOD_POOL_NONE ? ah : ah + 1;
  const int GW = pool_mode == RLOD_POOL_NONE ? aw : aw + 1;
  AlignWs ws = carve_align_ws(workspace, B, R, GH, GW);
  if (workspace_bytes < ws.bytes) return RLOD_ENOSPC;
  bool fast = align_fwd_geometry_fast(B, C, H, W, R, GH, GW, pool_mode);
  if (channels_last && (!fast || ((uintptr_t)feat % 16) != 0)) return RLOD_EUNSUPPORTED;
  TileGrid tg;
  const bool tiled = align_fwd_tiles(B, C, H, W, R, GH, GW, pool_mode, 0, &tg);  // (the plan call saw no layout flag)
  fast = fast || tiled;
  // the plan in the workspace is the plane kernel's: it needs a 16-byte aligned output (and map, when channels-last)
  if (fast && ((uintptr_t)out % 16) != 0) return RLOD_EINVAL;
  if (tiled && channels_last) return RLOD_EUNSUPPORTED;
  return align_fwd_run(feat, B, C, H, W, R, ah, aw, GH, GW, pool_mode, channels_last, fast, tiled ? &tg : nullptr, out, ws,
                       (cudaStream_t)stream, false);
}

RLOD_API int rlod_roi_align_backward(const float *grad_out, const float *rois, const float *feat,
                                     int B, int C, int H, int W, int R, int ah, int aw,
                                     float spatial_scale, int pool_mode, int accumulate,
                                     float *grad_in, void *workspace, size_t workspace_bytes,
                                     rlod_stream_t stream) {
  int rc = check_align_args(rois, B, C, H, W, R, ah, aw, pool_mode);
  if (rc != RLOD_OK) return rc;
  if (B == 0 || C == 0) return RLOD_OK;
  if (!grad_in) return RLOD_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t gin_bytes = (size_t)B * C * H * W * sizeof(float);
  if (R == 0) {
    if (!accumulate) cudaMemsetAsync(grad_in, 0, gin_bytes, st);
    return launch_status();
  }
  if (!grad_out || !workspace) return RLOD_EINVAL;
  if (pool_mode == RLOD_POOL_MAX && !feat) return RLOD_EINVAL;
  const int GH = pool_mode == RLOD_POOL_NONE ? ah : ah + 1;
  const int GW = pool_mode == RLOD_POOL_NONE ? aw : aw + 1;
  AlignWs ws = carve_align_ws(workspace, B, R, GH, GW);
  if (workspace_bytes < ws.bytes) return RLOD_ENOSPC;
  // lock-free kernel (roi_align_bwd.cu); the lock kernel below stays behind RLOD_BWD_V1=1 for A/B runs
  static const bool bwd_v1 = getenv("RLOD_BWD_V1") != nullptr;
  if (!bwd_v1 && GH == 8 && GW == 8 && (C % 4) == 0 && pool_mode != RLOD_POOL_MAX &&
      ((uintptr_t)grad_out % 16) == 0 && bwd_own_supported(H, W, pool_mode))
    return launch_bwd_own(grad_out, rois, B, C, H, W, R, spatial_scale, pool_mode, accumulate, grad_in, ws, st);
  const int P = walk_pitch(W);
  const size_t smem = fwd_walk_smem(H, W, pool_mode) + (size_t)((H + 2 + 3) & ~3) * sizeof(int) +
                      (size_t)kWalkWarps * 2 * sizeof(uint64_t) + 16;
  const bool fast = GH == 8 && GW == 8 && (C % 4) == 0 && pool_mode != RLOD_POOL_MAX &&
                    ((uintptr_t)grad_out % 16) == 0 && smem <= (size_t)kMaxSmemPerCta &&
                    (H + 2) * P <= 8192 && H + 2 <= 64 * 4 && R >= 2 * B;
  if (fast) {
    cudaMemsetAsync(ws.flag, 0, 4 * sizeof(int), st);
    RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st,
                k_roi_plan8_walk<<<(unsigned)cdiv(R, 4), 128, 0, st>>>(rois, R, B, H, W, P, spatial_scale, 1, kNoTiles, ws));
    RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st, k_roi_group_fixup<<<1, 32, 0, st>>>(R, B, ws));
    RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st,
                k_roi_order_by_key<<<B, kOrderThreads, 0, st>>>(ws.ext, ws.order, ws.img_off, 16, 20, 8, ws.order2));
    const int n_chunks = C / 4;
    const unsigned grid = (unsigned)(B * n_chunks);
    const int lock_mul = (65536 + P - 1) / P;  // row = (pixel offset * lock_mul) >> 16, exact for rows < 2^16 / P
#define RLOD_LAUNCH_BWD(POOL)                                                                  \
  do {                                                                                         \
    cudaFuncSetAttribute(k_align8_bwd_walk<POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                         (int)smem);                                                           \
    ProfScope _ps(RLOD_KERNEL_ALIGN_BWD, st);                                                  \
    k_align8_bwd_walk<POOL><<<grid, kWalkThreads, smem, st>>>(grad_out, ws.ext, ws.order2,     \
                                                              ws.img_off, C, H, W, P, n_chunks, \
                                                              accumulate, lock_mul, grad_in);  \
  } while (0)
    if (pool_mode == RLOD_POOL_NONE)
      RLOD_LAUNCH_BWD(RLOD_POOL_NONE);
    else
      RLOD_LAUNCH_BWD(RLOD_POOL_AVG);
#undef RLOD_LAUNCH_BWD
    return launch_status();
  }
  rc = build_plan(rois, B, H, W, R, GH, GW, spatial_scale, ws, st);
  if (rc) return rc;
  if (!accumulate) cudaMemsetAsync(grad_in, 0, gin_bytes, st);
  const int IH = pool_mode == RLOD_POOL_MAX ? ah : GH, IW = pool_mode == RLOD_POOL_MAX ? aw : GW;
  const long long total = (long long)R * C * IH * IW;
  const unsigned grid = (unsigned)(cdiv(total, 256) < (1LL << 30) ? cdiv(total, 256) : (1LL << 30));
  if (pool_mode == RLOD_POOL_NONE)
    RLOD_LAUNCH(RLOD_KERNEL_ALIGN_BWD_GENERIC, st, k_align_bwd_generic<RLOD_POOL_NONE>
        <<<grid, 256, 0, st>>>(grad_out, feat, ws.plan, ws.roi_b, C, H, W, GH, GW, total, grad_in));
  else if (pool_mode == RLOD_POOL_AVG)
    RLOD_LAUNCH(RLOD_KERNEL_ALIGN_BWD_GENERIC, st, k_align_bwd_generic<RLOD_POOL_AVG>
        <<<grid, 256, 0, st>>>(grad_out, feat, ws.plan, ws.roi_b, C, H, W, GH, GW, total, grad_in));
  else
    RLOD_LAUNCH(RLOD_KERNEL_ALIGN_BWD_GENERIC, st, k_align_bwd_generic<RLOD_POOL_MAX>
        <<<grid, 256, 0, st>>>(grad_out, feat, ws.plan, ws.roi_b, C, H, W, GH, GW, total, grad_in));
  return launch_status();
}
