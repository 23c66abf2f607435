// RoIAlign / RoIAlignAvg backward, 8x8 sample grid: whole-roi warps and a flush token ring (round 2).
//
// Replaces the per-row spin locks of k_align8_bwd_walk (roi_align.cu; kept behind RLOD_BWD_V1=1 for
// A/B runs): there, four independent rois per warp read-modify-write the four gradient planes of
// (image, 4 channels) in shared memory, lanes of a roi that share a pixel column take turns (rank rounds),
// and ~40 % of the 2 580 instructions per group of four rois are lock traffic.  Here:
//
//   * A warp serves ONE roi at a time with all 32 lanes: lane = (column slot 0..15, channel pair).
//     The plan kernel (k_roi_plan_own) merges the 16 column taps of the roi's 8 sample columns into
//     its DISTINCT pixel columns (<= 16) and gives every column slot its weights over the pooled
//     columns of grad_out (2 for a wide roi, up to 7 when the whole roi sits inside two pixels; the
//     1/4 of the average pooling is folded in as 1/2 per axis).  Lanes of one instruction therefore
//     never meet in a pixel, whatever the roi's size -- no rank rounds, no intra-warp exclusion.
//   * All lanes walk the 8 sample rows together with the forward kernel's two line slots: a slot
//     accumulates while consecutive sample rows share the pixel row and is handed to one of 16 static
//     flush slots when the row is left, so a roi adds to every DISTINCT pixel it touches exactly once.
//     The slot logic is pure arithmetic on two 0 / 1 factors per sample row that the plan supplies (keep,
//     shift): five packed operations per sample row, no integer instruction, no select.  The flushes of a roi hit distinct rows by
//     construction, so they are issued as independent batches (8 LDS.64, 8 packed adds, 8 STS.64, twice);
//     slots that are not flushed point at a dump row behind the planes instead of being predicated.
//   * Exclusion between the warps of a CTA (4-12 warps share the planes) is a TOKEN RING instead of
//     locks: warp w flushes its k-th roi after warp w-1 flushed its k-th (named barriers in the
//     producer / consumer form: bar.arrive by the previous warp, bar.sync by this one -- no atomics,
//     no polling).  Everything but the 48-instruction flush runs outside the token, and as a by-product
//     every pixel's sum is formed in a fixed order: the result is bit-reproducible run to run (the
//     lock kernel and the reference's atomicAdd are not).  Spin locks with CAS or with acquire-load
//     polling were measured 4-8x slower: the waiting warps' shared-memory atomics starve the holder
//     (profiles/experiments/README.md section 11).
//   * Several small CTAs per SM (6 warps each, 3 at 50x75 and 4 at 38x63) rather than two big ones: one
//     flush at a time per CTA is the serial resource, so more plane sets per SM = more flushes in flight.
//   * grad_out tiles (784 bytes per roi and 4 channels) arrive through a per-warp ring of bulk async
//     copies (TMA engine) behind mbarriers, 2-8 rois ahead; the next roi's record is prefetched into L1.
//   * Arithmetic is packed fp32 (FFMA2 / FADD2): a lane's two channels are one 64-bit register.
//
// Semantics: /root/reference/lib/model/roi_align/src/roi_align_kernel.cu:94-143 (scatter of every
// sample gradient to its four taps) after the avg_pool2d(2, 1) backward of
// lib/model/roi_align/modules/roi_align.py:26-29.  Rounding order differs from the reference's
// atomicAdd order (which is not fixed); tolerance 1e-5 in the tests.  Zero-weight padding of the column
// windows and the 0 / 1 factors of the row walk multiply the roi's own grad_out values by 0: a non-finite
// grad_out value can therefore reach every tap of its roi as NaN, where the reference confines it to the taps
// of its samples.
#include "roi_lists.cuh"

namespace rlod {

typedef unsigned long long u64;

constexpr int kOwnWords = 192;  // record words per roi (768 bytes)
// record layout (32-bit words):
//   [0..127]   column slot n (0..15): 8 weights over grad_out columns a_n .. a_n + L_n - 1
//   [128..143] column slot n: x * 16 (bits 0-15) | a_n << 16 (3 bits) | L_n << 20 (4 bits; 0 = unused slot) | max L << 24
//   [144..159] flush slot k: byte offset of the pixel row it adds to (y * W * 16), or H * W * 16 = the
//              dump row behind the planes when the slot is not flushed.  Slot 2t-2 / 2t-1 (t = 1..7) holds the
//              upper / lower line slot as it is when the walk reaches sample row t, slots 14 / 15 the two at
//              the end
//   [160..191] sample row t: {weight of the upper tap row (1 - ratio), of the lower tap row (ratio) [* 1/2 for AVG],
//              mT, mS}: the walk's line slots as arithmetic -- upper' = mT * upper + mS * lower + wt * G,
//              lower' = mT * lower + wb * G, with (mT, mS) = (1, 0) when sample row t shares its tap rows with
//              t - 1, (0, 1) when its upper row is the previous lower row, (0, 0) when it starts afresh
//   column-slot word bits 24-27: max L_n over the slots (the same in every slot)

__device__ __forceinline__ u64 pack2f(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ void fma2_acc(u64 &acc, u64 a, u64 b) {  // acc += a * b, in place
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 addp2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// One axis of the roi's sample grid, the reference's expression order (roi_align_kernel.cu:106-121
// as nvcc compiles it; see k_roi_plan in roi_align.cu).  Invalid samples get index -4.
struct Axis8 {
  int idx[8];
  float r[8];
  unsigned valid;
};
__device__ __forceinline__ void axis8(float c0, float c1, float scale, int dim, bool bvalid, Axis8 &a) {
  const float start = __fmul_rn(c0, scale);
  const float size = fmaxf(__fadd_rn(__fmaf_rn(c1, scale, -start), 1.f), 0.f);
  const float bin = (float)((double)size / 7.0);
  a.valid = 0;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float pos = __fmaf_rn((float)p, bin, start);
    const bool v = bvalid && (pos >= 0.f) && (pos < (float)dim);
    int idx = -4;
    float ratio = 0.f;
    if (v) {
      idx = (int)fminf(floorf(pos), (float)(dim - 2));
      ratio = __fsub_rn(pos, (float)idx);
      a.valid |= 1u << p;
    }
    a.idx[p] = idx, a.r[p] = ratio;
  }
}

// ----------------------------------------------------------------------------------------
// plan: one warp per roi.  Lanes 0-15 build the column slots, lane 16 the row walk.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    k_roi_plan_own(const float *__restrict__ rois, int R, int B, int H, int W, float scale, int avg,
                   AlignWs ws) {
  pdl_trigger();  // the kernels behind this one are set up now; they wait for the plan themselves
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const float *roi = rois + (size_t)r * 5;
  const float bf = roi[0];
  const int bi = (int)bf;
  const bool bvalid = (bf >= 0.f) && (bi < B);
  int *rec = ws.own + (size_t)r * kOwnWords;
  if (lane == 0) roi_list_mark(rois, r, R, B, bvalid ? bi : 0, ws);

  // ---- columns ---------------------------------------------------------------------------
  Axis8 cx;
  axis8(roi[1], roi[3], scale, W, bvalid, cx);
  // tap k = (sample k >> 1, side k & 1): is it the first tap on its pixel column?  (sample indices
  // ascend, so an earlier tap on the same column belongs to the previous sample)
  const int kj = (lane >> 1) & 7, side = lane & 1;
  int my_idx = -4, prev_idx = -4;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    if (p == kj) my_idx = cx.idx[p];
    if (p + 1 == kj) prev_idx = cx.idx[p];
  }
  const bool tap_valid = lane < 16 && my_idx >= 0;
  bool dup = false;
  if (prev_idx >= 0) dup = side == 0 ? (prev_idx == my_idx || prev_idx + 1 == my_idx) : (prev_idx == my_idx);
  const unsigned firsts = __ballot_sync(full, tap_valid && !dup) & 0xffffu;
  const int ncols = __popc(firsts);
  const int tap_col = my_idx + side;
  // slot n takes the column of the n-th first tap
  const int src = lane < ncols ? (int)__fns(firsts, 0, lane + 1) : 0;
  const int X = __shfl_sync(full, tap_col, src);
  int Ln = 0, qa = 0, qb = -1;
  float wv[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) wv[q] = 0.f;
  if (lane < ncols) {
    float u[9];
    int ja = 8, jb = -1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool m0 = cx.idx[j] == X, m1 = cx.idx[j] + 1 == X;  // invalid samples: idx = -4, X >= 0
      u[j] = m0 ? 1.f - cx.r[j] : (m1 ? cx.r[j] : 0.f);
      if (m0 || m1) {
        ja = ja < j ? ja : j;
        jb = j;
      }
    }
    u[8] = 0.f;
    if (avg) {  // pooled column q feeds sample columns q and q + 1, half of the 1/4 on this axis
      qa = ja > 0 ? ja - 1 : 0, qb = jb < 6 ? jb : 6;
#pragma unroll
      for (int q = 0; q < 7; ++q) wv[q] = 0.5f * (u[q] + u[q + 1]);
    } else {
      qa = ja, qb = jb;
#pragma unroll
      for (int q = 0; q < 8; ++q) wv[q] = u[q];
    }
    Ln = qb - qa + 1;
  }
  // every slot reads Lmax consecutive grad_out columns (zero weights where it has fewer): its window
  // is shifted left where it would leave the row, so only the roi's own values are ever multiplied
  const int Lmax = __reduce_max_sync(full, Ln);
  // slot of every column: a flush request serves slots 0-7 and 8-15 as two half-warps, each conflict-free
  // when its eight pixel columns differ mod 8 (16-byte pixels, 128-byte bank window).  Greedy, in column
  // order: the first half-warp that does not hold the residue yet
  int dest = lane;
  {
    unsigned mA = 0u, mB = 0u;
    int cA = 0, cB = 0;
    for (int nn = 0; nn < ncols; ++nn) {  // warp-uniform
      const int rn = __shfl_sync(full, X, nn) & 7;
      int d;
      if (!((mA >> rn) & 1u) && cA < 8) d = cA++, mA |= 1u << rn;
      else if (!((mB >> rn) & 1u) && cB < 8) d = 8 + cB++, mB |= 1u << rn;
      else if (cA < 8) d = cA++;
      else d = 8 + cB++;
      if (nn == lane) dest = d;
    }
  }
  if (lane < 16) {  // every slot empty first, then the used ones
    int *wr = rec + lane * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) wr[i] = 0;
    rec[128 + lane] = Lmax << 24;
  }
  __syncwarp();
  if (lane < ncols && Ln > 0) {
    const int OWp = avg ? 7 : 8;
    const int a = qa + Lmax <= OWp ? qa : OWp - Lmax;
    int *wr = rec + dest * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (q >= qa && q <= qb) wr[q - a] = __float_as_int(wv[q]);
    rec[128 + dest] = (X * 16) | (a << 16) | (Ln << 20) | (Lmax << 24);
  }

  // ---- rows ------------------------------------------------------------------------------
  if (lane == 16) {
    Axis8 cy;
    axis8(roi[2], roi[4], scale, H, bvalid, cy);
    const float f = avg ? 0.5f : 1.f;
    const int dump = H * W * 16, rowb = W * 16;
    int cur = -1;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      int code;
      float wt = 0.f, wb = 0.f;
      const int prev = cur;
      if (cy.idx[t] >= 0) {
        const int y = cy.idx[t];
        if (cur < 0) code = 3;
        else code = y == cur ? 0 : (y == cur + 1 ? 1 : 2);
        cur = y;
        wb = f * cy.r[t];
        wt = f * (1.f - cy.r[t]);
      } else {
        code = cur >= 0 ? 2 : 3;
        cur = -1;
      }
      if (t > 0) {
        rec[144 + 2 * t - 2] = (code == 1 || code == 2) ? prev * rowb : dump;
        rec[144 + 2 * t - 1] = code == 2 ? (prev + 1) * rowb : dump;
      }
      rec[160 + 4 * t] = __float_as_int(wt);
      rec[161 + 4 * t] = __float_as_int(wb);
      rec[162 + 4 * t] = __float_as_int(code == 0 ? 1.f : 0.f);
      rec[163 + 4 * t] = __float_as_int(code == 1 ? 1.f : 0.f);
    }
    rec[144 + 14] = cur >= 0 ? cur * rowb : dump;
    rec[144 + 15] = cur >= 0 ? (cur + 1) * rowb : dump;
  }
}

// ----------------------------------------------------------------------------------------
// the kernel
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ u64 lds64(uint32_t a) {
  u64 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, u64 v) {
  asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}

// B16: all 16 read-modify-writes of a flush as ONE batch (one shared-memory latency under the token instead of
// two; needs 16 more registers, so only the instances that run three or fewer CTAs per SM use it)
template <int POOL, int NW, bool B16>
__global__ void __launch_bounds__(NW * 32, B16 ? (18 / NW > 0 ? 18 / NW : 1) : 24 / NW)
    k_align8_bwd_own(const float *__restrict__ gout, const int *__restrict__ rec,
                     const int *__restrict__ order, const int *__restrict__ img_off, int C, int H,
                     int W, int n_quads, int ns_log2, int accumulate, float *__restrict__ gin) {
  constexpr int OW = POOL == RLOD_POOL_NONE ? 8 : 7;
  constexpr int OHW = OW * OW;
  constexpr int STG = 4 * OHW;  // floats per tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NS = 1 << ns_log2;
  const int b = blockIdx.x / n_quads, quad = blockIdx.x - b * n_quads;
  const int HW = H * W;
  float4 *planes = reinterpret_cast<float4 *>(smem_raw);
  // planes [HW pixels] | dump row [W pixels] | tiles | tile barriers
  const size_t planes_bytes = (size_t)(HW + W) * 16;
  float *tiles = reinterpret_cast<float *>(smem_raw + planes_bytes) + (size_t)warp * NS * STG;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + planes_bytes + (size_t)NW * NS * STG * 4) + warp * NS;
  float *dst = gin + ((size_t)b * C + (size_t)quad * 4) * HW;
  const float *gsrc = gout + (size_t)quad * 4 * OHW;

  // planes: zero, or the caller's gradient when accumulating.  A pixel holds (c0, c2, c1, c3): a
  // lane's channel pair (h, h + 2) is one aligned 8-byte word, and the two lanes of a column read
  // grad_out channels h = 0 / 1 at the same time (49 floats apart: disjoint banks)
  if (accumulate) {
    for (int p = threadIdx.x; p < HW; p += NW * 32)
      planes[p] = make_float4(__ldg(dst + p), __ldg(dst + 2 * (size_t)HW + p), __ldg(dst + (size_t)HW + p),
                              __ldg(dst + 3 * (size_t)HW + p));
  } else {
    for (int p = threadIdx.x; p < HW; p += NW * 32) planes[p] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (lane == 0)
    for (int s = 0; s < NS; ++s) mbar_init(bars + s, 1);
  __syncthreads();
  // launched behind the plan kernels as their programmatic dependent: the planes above were set up while
  // those ran, the plan and the roi lists are complete and visible from here on
  pdl_wait();
  const int r0 = img_off[b], n = img_off[b + 1] - r0;

  // warp w serves rois w, w + NW, ... of the image (every roi costs the same: no work queue)
  auto roi_at = [&](int k) {  // k-th roi of this warp
    const int i = warp + k * NW;
    return i < n ? __ldg(order + r0 + i) : -1;
  };
  auto issue = [&](int k, int rr) {  // lane 0 only
    const int s = k & (NS - 1);
    mbar_expect_tx(bars + s, (uint32_t)(STG * 4));
    bulk_g2s(tiles + s * STG, gsrc + (size_t)rr * C * OHW, (uint32_t)(STG * 4), bars + s);
  };
  if (lane == 0)
    for (int k = 0; k < NS; ++k) {
      const int rr = roi_at(k);
      if (rr >= 0) issue(k, rr);
    }

  const int nslot = lane >> 1, h = lane & 1;
  const uint32_t pl = smem_u32(planes) + 8u * (uint32_t)h;
  // this warp's roi ids, 32 at a time: lane l holds the id of its (kbase + l)-th roi
  int ids = roi_at(lane), kbase = 0;
  auto roi_id = [&](int k) {  // kbase <= k < kbase + 32 (warp-uniform k)
    return __shfl_sync(0xffffffffu, ids, k - kbase);
  };
  int r = roi_id(0), r_next = roi_id(1);
  auto prefetch_rec = [&](int rr) {  // 768 bytes = 6 lines
    if (rr >= 0 && lane < 6) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + (size_t)rr * kOwnWords + lane * 32));
  };
  prefetch_rec(r);

  for (int k = 0; r >= 0; ++k) {
    const int *rp = rec + (size_t)r * kOwnWords;
    prefetch_rec(r_next);
    if (k + NS >= kbase + 32) {  // the ring holds ids k .. ; refill it so that k + NS is inside
      kbase = k;
      ids = roi_at(k + lane);
    }
    const int r_refill = roi_id(k + NS);
    const int r_nn = roi_id(k + 2);
    const int meta = __ldg(rp + 128 + nslot);
    const float4 w0 = __ldg(reinterpret_cast<const float4 *>(rp + nslot * 8));
    const int Lmax = (meta >> 24) & 15;
    float4 w1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (Lmax > 4) w1 = __ldg(reinterpret_cast<const float4 *>(rp + nslot * 8) + 1);
    const int an = (meta >> 16) & 7;
    const bool act = ((meta >> 20) & 15) != 0;
    const uint32_t pa = pl + ((uint32_t)meta & 0xffffu);
    const float wcol[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};

    // ---- column stage: hv[p] = sum_i w[i] * grad_out[pair][p][a_n + i], i < Lmax (warp-uniform) ----
    const int s = k & (NS - 1);
    mbar_wait(bars + s, (uint32_t)((k >> ns_log2) & 1));
    const float *tl = tiles + s * STG + h * OHW + an;
    u64 hv[OW];
#pragma unroll
    for (int p = 0; p < OW; ++p) hv[p] = mul2(pack2f(wcol[0], wcol[0]), pack2f(tl[p * OW], tl[2 * OHW + p * OW]));
#pragma unroll
    for (int ii = 1; ii < OW; ++ii) {
      if (ii >= Lmax) break;
#pragma unroll
      for (int p = 0; p < OW; ++p)
        fma2_acc(hv[p], pack2f(wcol[ii], wcol[ii]), pack2f(tl[p * OW + ii], tl[2 * OHW + p * OW + ii]));
    }

    // gradient of the lane's column at every sample row (AVG: the two pooled rows of a sample row)
    u64 G[8];
    if (POOL == RLOD_POOL_NONE) {
#pragma unroll
      for (int t = 0; t < 8; ++t) G[t] = hv[t % OW];
    } else {
      G[0] = hv[0];
#pragma unroll
      for (int t = 1; t < 7; ++t) G[t] = addp2(hv[t - 1], hv[t]);
      G[7] = hv[6];
    }

    // ---- row walk: two line slots, 16 static flush slots.  The slot logic is arithmetic on (mT, mS) from
    // the plan: no integer instruction, no select -- five packed operations per sample row
    u64 sv[16];
    u64 accT, accB;
    {
      const float4 wk = __ldg(reinterpret_cast<const float4 *>(rp + 160));
      accT = mul2(pack2f(wk.x, wk.x), G[0]), accB = mul2(pack2f(wk.y, wk.y), G[0]);
    }
#pragma unroll
    for (int t = 1; t < 8; ++t) {
      const float4 wk = __ldg(reinterpret_cast<const float4 *>(rp + 160 + 4 * t));
      sv[2 * t - 2] = accT, sv[2 * t - 1] = accB;
      const u64 mT = pack2f(wk.z, wk.z), mS = pack2f(wk.w, wk.w);
      const u64 x = fma2(mT, accT, mul2(mS, accB));
      accB = fma2(pack2f(wk.y, wk.y), G[t], mul2(mT, accB));
      accT = fma2(pack2f(wk.x, wk.x), G[t], x);
    }
    sv[14] = accT, sv[15] = accB;

    // flush addresses: slots that are not flushed point at the dump row behind the planes
    const int4 o0 = __ldg(reinterpret_cast<const int4 *>(rp + 144)), o1 = __ldg(reinterpret_cast<const int4 *>(rp + 148));
    const int4 o2 = __ldg(reinterpret_cast<const int4 *>(rp + 152)), o3 = __ldg(reinterpret_cast<const int4 *>(rp + 156));
    const uint32_t fa[16] = {pa + o0.x, pa + o0.y, pa + o0.z, pa + o0.w, pa + o1.x, pa + o1.y, pa + o1.z, pa + o1.w,
                             pa + o2.x, pa + o2.y, pa + o2.z, pa + o2.w, pa + o3.x, pa + o3.y, pa + o3.z, pa + o3.w};

    // the tile's values were consumed by the walk: refill its slot NS rois ahead
    if (lane == 0 && r_refill >= 0) issue(k + NS, r_refill);

    // ---- flush when the CTA's token arrives (warps take turns, in roi order: the sum of every pixel
    // is formed in a fixed order, run to run).  The slots hit distinct rows, the lanes distinct
    // columns / channel pairs, so the 16 read-modify-writes are independent.
    // (named barriers: the previous warp arrives on barrier 1 + warp, this warp syncs on it -- the
    // producer / consumer form of bar.arrive + bar.sync, which orders the producer's stores before the
    // consumer's loads)
    const int i_glob = warp + k * NW;
    asm volatile("" ::"r"(fa[0]), "r"(fa[1]), "r"(fa[2]), "r"(fa[3]), "r"(fa[4]), "r"(fa[5]), "r"(fa[6]), "r"(fa[7]),
                 "r"(fa[8]), "r"(fa[9]), "r"(fa[10]), "r"(fa[11]), "r"(fa[12]), "r"(fa[13]), "r"(fa[14]), "r"(fa[15]));
    if (i_glob > 0) asm volatile("bar.sync %0, 64;" ::"r"(1 + warp) : "memory");
    if (act) {
      if (B16) {
        u64 old[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) old[j] = lds64(fa[j]);
#pragma unroll
        for (int j = 0; j < 16; ++j) sts64(fa[j], addp2(old[j], sv[j]));
      } else {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          u64 old[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) old[j] = lds64(fa[half * 8 + j]);
#pragma unroll
          for (int j = 0; j < 8; ++j) sts64(fa[half * 8 + j], addp2(old[j], sv[half * 8 + j]));
        }
      }
    }
    if (i_glob + 1 < n) asm volatile("bar.arrive %0, 64;" ::"r"(warp + 1 == NW ? 1 : warp + 2) : "memory");
    r = r_next, r_next = r_nn;
  }
  __syncthreads();

  // planes -> HBM: a thread takes a pixel, four coalesced 128-byte stores per warp instruction
  for (int p = threadIdx.x; p < HW; p += NW * 32) {
    const float4 v = planes[p];
    dst[p] = v.x;
    dst[2 * (size_t)HW + p] = v.y;
    dst[(size_t)HW + p] = v.z;
    dst[3 * (size_t)HW + p] = v.w;
  }
}

// ----------------------------------------------------------------------------------------
// host side: called by rlod_roi_align_backward (roi_align.cu)
// ----------------------------------------------------------------------------------------
static size_t own_cta_bytes(int H, int W, int pool_mode, int ns, int nw) {
  const int ohw = pool_mode == RLOD_POOL_NONE ? 64 : 49;
  return align_up((size_t)(H * W + W) * 16 + (size_t)nw * ns * (ohw * 16 + 8), 128);
}

bool bwd_own_supported(int H, int W, int pool_mode) {
  return W * 16 < 65536 && own_cta_bytes(H, W, pool_mode, 2, 4) <= (size_t)kMaxSmemPerCta;
}

int launch_bwd_own(const float *grad_out, const float *rois, int B, int C, int H, int W, int R,
                   float spatial_scale, int pool_mode, int accumulate, float *grad_in, const AlignWs &ws,
                   cudaStream_t st) {
  cudaMemsetAsync(ws.flag, 0, 4 * sizeof(int), st);
  RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st,
              k_roi_plan_own<<<(unsigned)cdiv(R, 4), 128, 0, st>>>(rois, R, B, H, W, spatial_scale,
                                                                  pool_mode == RLOD_POOL_AVG ? 1 : 0, ws));
  // Programmatic dependent launch of the two kernels below (as in the forward) is wired but left off: measured
  // 220 against 207 us at C2 and 1349 against 1352 us at C4 (RLOD_BWD_PDL=1 turns it on) -- the early CTAs of the
  // backward land unevenly on the SMs the plan kernel still occupies, and the 1.7 waves of C2 end later for it.
  static const bool pdl = getenv("RLOD_BWD_PDL") != nullptr && pdl_enabled();
  RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st, launch_after(k_roi_group_fixup, dim3(1), dim3(32), 0, st, pdl, R, B, ws));
  // warps per CTA and tile ring depth per warp: 6-warp CTAs when at least two fit an SM (measured best at
  // C2 and C4: tools/time_op.py with RLOD_BWD_NW), else whatever puts the most warps on an SM; then the
  // deepest ring (8 / 4 / 2 tiles) that does not cost a resident CTA
  static const int env_ns = getenv("RLOD_BWD_NS") ? atoi(getenv("RLOD_BWD_NS")) : 0;
  static const int env_nw = getenv("RLOD_BWD_NW") ? atoi(getenv("RLOD_BWD_NW")) : 0;
  const size_t per_sm = (size_t)kMaxSmemPerCta - 1024;  // every resident CTA reserves 1 KB
  auto ctas = [&](int ns, int nw) {
    const int c = (int)(per_sm / (own_cta_bytes(H, W, pool_mode, ns, nw) + 1024));
    const int cap = 24 / nw;  // registers (__launch_bounds__ of the instance: 24 warps per SM)
    return c < cap ? c : cap;
  };
  int nw = 0;
  if (env_nw == 4 || env_nw == 6 || env_nw == 8 || env_nw == 12) {
    nw = env_nw;
  } else if (ctas(2, 6) >= 2) {
    nw = 6;
  } else {
    const int nws[4] = {12, 8, 6, 4};
    int best = 0;
    for (int wi = 0; wi < 4; ++wi)
      if (ctas(2, nws[wi]) * nws[wi] > best) best = ctas(2, nws[wi]) * nws[wi], nw = nws[wi];
    if (nw == 0) return RLOD_EUNSUPPORTED;
  }
  int ns_log2 = 1;
  for (int l2 = 2; l2 <= 3; ++l2)
    if (ctas(1 << l2, nw) >= ctas(2, nw) && ctas(1 << l2, nw) > 0) ns_log2 = l2;
  if (env_ns == 2 || env_ns == 4 || env_ns == 8) ns_log2 = env_ns == 2 ? 1 : (env_ns == 4 ? 2 : 3);
  const size_t smem = own_cta_bytes(H, W, pool_mode, 1 << ns_log2, nw);
  if (smem > (size_t)kMaxSmemPerCta) return RLOD_EUNSUPPORTED;
  const int n_quads = C / 4;
  const unsigned grid = (unsigned)(B * n_quads);
#define RLOD_LAUNCH_OWN(POOL, NWW, BB)                                                           \
  do {                                                                                           \
    static bool attr_set = false;                                                                \
    if (!attr_set) {                                                                             \
      cudaFuncSetAttribute(k_align8_bwd_own<POOL, NWW, BB>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           kMaxSmemPerCta);                                                      \
      attr_set = true;                                                                           \
    }                                                                                            \
    ProfScope _ps(RLOD_KERNEL_ALIGN_BWD, st);                                                    \
    launch_after(k_align8_bwd_own<POOL, NWW, BB>, dim3(grid), dim3(NWW * 32), smem, st, pdl,     \
                 grad_out, (const int *)ws.own, (const int *)ws.order, (const int *)ws.img_off,  \
                 C, H, W, n_quads, ns_log2, accumulate, grad_in);                                \
  } while (0)
#define RLOD_LAUNCH_OWN_NW(POOL, BB)                    \
  do {                                                  \
    if (nw == 4) RLOD_LAUNCH_OWN(POOL, 4, BB);          \
    else if (nw == 6) RLOD_LAUNCH_OWN(POOL, 6, BB);     \
    else if (nw == 8) RLOD_LAUNCH_OWN(POOL, 8, BB);     \
    else RLOD_LAUNCH_OWN(POOL, 12, BB);                 \
  } while (0)
  // one-batch flush where shared memory (not registers) limits the resident CTAs to 18 warps or fewer
  static const int env_b16 = getenv("RLOD_BWD_B16") ? atoi(getenv("RLOD_BWD_B16")) : -1;
  bool b16 = ctas(1 << ns_log2, nw) * nw <= 18;
  if (env_b16 == 0) b16 = false;
  if (env_b16 == 1 && ctas(1 << ns_log2, nw) * nw <= 18) b16 = true;
  if (pool_mode == RLOD_POOL_NONE) {
    if (b16) RLOD_LAUNCH_OWN_NW(RLOD_POOL_NONE, true);
    else RLOD_LAUNCH_OWN_NW(RLOD_POOL_NONE, false);
  } else {
    if (b16) RLOD_LAUNCH_OWN_NW(RLOD_POOL_AVG, true);
    else RLOD_LAUNCH_OWN_NW(RLOD_POOL_AVG, false);
  }
#undef RLOD_LAUNCH_OWN_NW
#undef RLOD_LAUNCH_OWN
  return launch_status();
}

}  // namespace rlod
