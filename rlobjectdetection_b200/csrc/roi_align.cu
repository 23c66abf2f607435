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
//   2. forward (k_align8_fwd_planes): a CTA owns 4 consecutive channel planes of ONE image,
//      pulls them HBM -> shared memory once (coalesced, interleaved per pixel so that lanes
//      of different channels never share a bank) and then serves EVERY roi of that image
//      from shared memory.  The feature map is read from HBM exactly once, all bilinear taps
//      are LDS, and each (roi, 4 channels) result is staged in shared memory and leaves as
//      ONE 784-byte bulk async store (TMA engine).  No 8x8 intermediate tensor exists.
//   3. backward (k_align8_bwd_bands): a warp owns a band of rows of 4 gradient planes in
//      shared memory exclusively, walks the image's rois, and accumulates with plain
//      LDS/FADD/STS (segmented warp-shuffle reduction resolves intra-roi collisions), so
//      there is not a single atomic and the result is deterministic.  grad_in is written
//      once, coalesced; no memset is needed.
//   Generic kernels (any grid size / channel count / plane size, and RoIAlignMax backward)
//   cover everything the fast paths do not.
#include "rlod_common.cuh"

namespace rlod {

// ----------------------------------------------------------------------------------------
// workspace layout
// ----------------------------------------------------------------------------------------
struct AlignWs {
  int *flag;     // [4]   flag[0] != 0: rois are not grouped by image
  int *img_off;  // [B+1] roi list offsets per image
  int *cursor;   // [B]
  int *order;    // [R]   roi ids grouped by image (stable)
  int *roi_b;    // [R]   batch index (0 when out of range: the plan is all-invalid then)
  int *plan;     // [R * words]
  int *ext;      // [R * 64] forward-kernel record (8x8 grids only)
  size_t bytes;
};

static AlignWs carve_align_ws(void *base, int B, int R, int GH, int GW) {
  AlignWs w;
  size_t off = 0;
  char *p = (char *)base;
  auto take = [&](size_t n) {
    char *q = p ? p + off : nullptr;
    off += align_up(n, 128);
    return q;
  };
  w.flag = (int *)take(4 * sizeof(int));
  w.img_off = (int *)take((size_t)(B + 1) * sizeof(int));
  w.cursor = (int *)take((size_t)(B > 0 ? B : 1) * sizeof(int));
  w.order = (int *)take((size_t)(R > 0 ? R : 1) * sizeof(int));
  w.roi_b = (int *)take((size_t)(R > 0 ? R : 1) * sizeof(int));
  w.plan = (int *)take((size_t)(R > 0 ? R : 1) * (size_t)(2 * GH + 2 * GW) * sizeof(int));
  w.ext = (int *)take((GH == 8 && GW == 8) ? (size_t)(R > 0 ? R : 1) * 64 * sizeof(int) : 0);
  w.bytes = off;
  return w;
}

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
  if (slot == 0) {
    // roi lists under the assumption that rois are grouped by image (what _ProposalLayer
    // emits); k_roi_group_fixup redoes them when the flag is raised.
    ws.roi_b[r] = b;
    ws.order[r] = r;
    int prev;
    if (r == 0) {
      prev = -1;
    } else {
      const float pf = rois[(size_t)(r - 1) * 5];
      const int pi = (int)pf;
      prev = ((pf >= 0.f) && (pi < B)) ? pi : 0;
    }
    if (b < prev) atomicOr(ws.flag, 1);
    for (int q = prev + 1; q <= b; ++q) ws.img_off[q] = r;
    if (r == R - 1)
      for (int q = b + 1; q <= B; ++q) ws.img_off[q] = R;
  }
}

// one warp: stable counting sort of roi ids by image, only when the rois were not grouped
__global__ void k_roi_group_fixup(int R, int B, AlignWs ws) {
  if (ws.flag[0] == 0) return;
  const int lane = threadIdx.x;
  const unsigned full = 0xffffffffu;
  for (int b = lane; b < B; b += 32) ws.cursor[b] = 0;
  __syncwarp();
  for (int base = 0; base < R; base += 32) {
    const int r = base + lane;
    const int b = r < R ? ws.roi_b[r] : -1 - lane;
    const unsigned peers = __match_any_sync(full, b);
    if (r < R && lane == __ffs(peers) - 1) ws.cursor[b] += __popc(peers);
    __syncwarp();
  }
  // exclusive scan over images, 32 at a time
  int carry = 0;
  for (int base = 0; base < B; base += 32) {
    const int b = base + lane;
    const int c = b < B ? ws.cursor[b] : 0;
    int inc = c;
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(full, inc, d);
      if (lane >= d) inc += t;
    }
    if (b < B) {
      ws.img_off[b] = carry + inc - c;
      ws.cursor[b] = carry + inc - c;
    }
    carry += __shfl_sync(full, inc, 31);
  }
  if (lane == 0) ws.img_off[B] = carry;
  __syncwarp();
  for (int base = 0; base < R; base += 32) {
    const int r = base + lane;
    const int b = r < R ? ws.roi_b[r] : -1 - lane;
    const unsigned peers = __match_any_sync(full, b);
    const int leader = __ffs(peers) - 1;
    int cur = 0;
    if (r < R && lane == leader) {
      cur = ws.cursor[b];
      ws.cursor[b] = cur + __popc(peers);
    }
    cur = __shfl_sync(full, cur, leader);
    if (r < R) ws.order[cur + __popc(peers & ((1u << lane) - 1u))] = r;
    __syncwarp();
  }
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
                        long long total, float *__restrict__ out) {
  const int ah = POOL == RLOD_POOL_NONE ? GH : GH - 1, aw = POOL == RLOD_POOL_NONE ? GW : GW - 1;
  const int words = 2 * GH + 2 * GW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(idx % aw);
    const int oh = (int)((idx / aw) % ah);
    const int c = (int)((idx / ((long long)aw * ah)) % C);
    const int r = (int)(idx / ((long long)aw * ah * C));
    const int *pl = plan + (size_t)r * words;
    const float *plane = feat + ((size_t)roi_b[r] * C + c) * ((size_t)H * W);
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
    out[idx] = v;
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

constexpr int kFwdThreads = 256;

// bulk async store shared -> global (TMA engine, SASS: UBLKCP), tracked by bulk groups
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t smem_addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_addr));
  return v;
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------
// Shared-memory layout of the forward kernel.  The CTA's 4 channel planes are interleaved per
// pixel, planes4[pix] = (c0, c1, c2, c3); lane (ch, pw) reads word 4*pix + ch, so lanes of
// different channels can never collide on a bank (bank = 4*(pix mod 8) + ch).  Inside a row the
// column is XOR-swizzled, col' = col ^ ((col >> 3) & 7) (identity in the last partial group of
// 8), so the 8 sample columns of a roi fall on distinct bank groups for strides 1, 2, 3, 4 ...
// as well.  Two extra all-zero rows (H, H+1) stand in for out-of-range sample rows and one
// zero pixel for out-of-range sample columns: validity never appears in the hot loop.
//
// k_roi_plan8_fwd turns the generic plan into the forward kernel's per-roi record (64 words):
//   [0..7]   byte offset of sample row ph inside planes4 (the zero row when invalid)
//   [8..15]  row ratio hr (0 when invalid)
//   [16]     flags: bit ph = both rows must be (re)loaded, bit 8+ph = shift or reload,
//            bit 16+ph = shift only (the lower row becomes the previous upper row)
//   [32+4pw..] per sample column: c0 = byte offset of the swizzled left tap, dc = offset of
//            the right tap relative to it, m = 1 (0: column invalid -> c0 is the zero pixel and
//            the row offset is multiplied away), wr = column ratio
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ int swz_col(int col, int W8) {
  return col < W8 ? (col ^ ((col >> 3) & 7)) : col;
}

__global__ void k_roi_plan8_fwd(const int *__restrict__ plan, int R, int H, int W,
                                int *__restrict__ ext) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int *pl = plan + (size_t)r * 32;
  int *e = ext + (size_t)r * 64;
  const int W8 = W & ~7;
  const int row_bytes = 16 * W;
  int prev = -5;  // row index of the previous sample row (zero rows count as row H)
  int flags = 0;
  for (int ph = 0; ph < 8; ++ph) {
    const int hs = pl[ph];
    const int row = hs >= 0 ? hs : H;
    e[ph] = row * row_bytes;
    e[8 + ph] = hs >= 0 ? pl[8 + ph] : 0;
    if (row == prev) {
      // keep both interpolated rows
    } else if (row == prev + 1) {
      flags |= (1 << (8 + ph)) | (1 << (16 + ph));
    } else {
      flags |= (1 << ph) | (1 << (8 + ph));
    }
    prev = row;
  }
  e[16] = flags;
  for (int pw = 0; pw < 8; ++pw) {
    const int ws = pl[16 + 2 * pw];
    int c0, dc, m;
    if (ws >= 0) {
      c0 = 16 * swz_col(ws, W8);
      dc = 16 * swz_col(ws + 1, W8) - c0;
      m = 1;
    } else {
      c0 = H * row_bytes;  // pixel 0 of the first zero row; its lower neighbour is zero too
      dc = 0;
      m = 0;
    }
    e[32 + 4 * pw + 0] = c0;
    e[32 + 4 * pw + 1] = dc;
    e[32 + 4 * pw + 2] = m;
    e[32 + 4 * pw + 3] = pl[16 + 2 * pw + 1];
  }
}

struct Plan8F {
  int roff[8];
  float hr[8];
  int flags;
  int c0, dc, m;
  float wr;
};

__device__ __forceinline__ void load_plan8f(const int *__restrict__ ext, int r, int pw, Plan8F &p) {
  const int4 *q = reinterpret_cast<const int4 *>(ext + (size_t)r * 64);
  const int4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
  p.roff[0] = a.x, p.roff[1] = a.y, p.roff[2] = a.z, p.roff[3] = a.w;
  p.roff[4] = b.x, p.roff[5] = b.y, p.roff[6] = b.z, p.roff[7] = b.w;
  p.hr[0] = __int_as_float(c.x), p.hr[1] = __int_as_float(c.y);
  p.hr[2] = __int_as_float(c.z), p.hr[3] = __int_as_float(c.w);
  p.hr[4] = __int_as_float(d.x), p.hr[5] = __int_as_float(d.y);
  p.hr[6] = __int_as_float(d.z), p.hr[7] = __int_as_float(d.w);
  p.flags = __ldg(ext + (size_t)r * 64 + 16);
  const int4 w = __ldg(q + 8 + pw);
  p.c0 = w.x, p.dc = w.y, p.m = w.z;
  p.wr = __int_as_float(w.w);
}

template <int POOL>
__global__ void __launch_bounds__(kFwdThreads, 3)
    k_align8_fwd_planes(const float *__restrict__ feat, const int *__restrict__ ext,
                        const int *__restrict__ order, const int *__restrict__ img_off, int C,
                        int H, int W, int n_chunks, float *__restrict__ out) {
  constexpr int OW = POOL == RLOD_POOL_NONE ? 8 : 7;
  constexpr int OHW = OW * OW;   // 64 | 49
  constexpr int STG = 4 * OHW;   // floats per (roi, 4 channels): 256 | 196
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *planes4 = reinterpret_cast<float4 *>(smem_raw);
  const int HW = H * W, W8 = W & ~7;
  float *stage = reinterpret_cast<float *>(planes4 + HW + 2 * W);

  const int b = blockIdx.x / n_chunks, chunk = blockIdx.x - b * n_chunks;
  const int r0 = img_off[b], r1 = img_off[b + 1];
  if (r0 >= r1) return;
  const float *src = feat + ((size_t)b * C + (size_t)chunk * 4) * HW;
#pragma unroll 4
  for (int p = threadIdx.x; p < HW; p += kFwdThreads) {
    const int y = p / W, x = p - y * W;
    planes4[y * W + swz_col(x, W8)] = make_float4(__ldg(src + p), __ldg(src + HW + p),
                                                  __ldg(src + 2 * HW + p), __ldg(src + 3 * HW + p));
  }
  for (int p = threadIdx.x; p < 2 * W; p += kFwdThreads) planes4[HW + p] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kFwdThreads >> 5;
  const int pw = lane & 7, ch = lane >> 3;
  const uint32_t pbase = smem_u32(planes4) + 4u * (uint32_t)ch;  // word 4*pix + ch
  const uint32_t w16 = 16u * (uint32_t)W;                           // bytes per plane row
  float *stg = stage + warp * (2 * STG);

  int k = r0 + warp;
  Plan8F cur;
  int r = 0;
  if (k < r1) {
    r = order[k];
    load_plan8f(ext, r, pw, cur);
  }
  __syncthreads();

  for (int it = 0; k < r1; k += nw, ++it) {
    // prefetch the next roi's plan while this one is computed
    Plan8F nxt;
    int rn = 0;
    if (k + nw < r1) {
      rn = order[k + nw];
      load_plan8f(ext, rn, pw, nxt);
    }
    const float wr = cur.wr;
    const uint32_t cbase = pbase + (uint32_t)cur.c0;
    const int fl = cur.flags;  // warp-uniform: the sample rows come from the roi alone
    float s[8];
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int ph = 0; ph < 8; ++ph) {
      // warp-uniform flag bits: which of the two interpolated rows must be refreshed
      const bool reload = (fl >> ph) & 1, adv = (fl >> (8 + ph)) & 1, shift = (fl >> (16 + ph)) & 1;
      const uint32_t pa = cbase + (uint32_t)(cur.roff[ph] * cur.m);
      const uint32_t pb = pa + (uint32_t)cur.dc;
      if (shift) t0 = t1;
      if (reload) {
        const float x0 = lds_f32(pa), x1 = lds_f32(pb);
        t0 = fmaf(wr, x1 - x0, x0);
      }
      if (adv) {
        const float y0 = lds_f32(pa + w16), y1 = lds_f32(pb + w16);
        t1 = fmaf(wr, y1 - y0, y0);
      }
      s[ph] = fmaf(cur.hr[ph], t1 - t0, t0);
    }
    // the buffer about to be written was handed to the bulk-copy engine two iterations ago
    if (it >= 2) {
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    float *sbuf = stg + (it & 1) * STG;
    float *sb = sbuf + ch * OHW;
    if (POOL == RLOD_POOL_NONE) {
#pragma unroll
      for (int ph = 0; ph < 8; ++ph) sb[ph * 8 + pw] = s[ph];
    } else {
      float sr[8];
#pragma unroll
      for (int ph = 0; ph < 8; ++ph) sr[ph] = __shfl_down_sync(0xffffffffu, s[ph], 1);
      if (pw < 7) {
#pragma unroll
        for (int i = 0; i < 7; ++i) sb[i * 7 + pw] = pool4<POOL>(s[i], sr[i], s[i + 1], sr[i + 1]);
      }
    }
    // 4 channels x OHW floats are one contiguous, 16-byte aligned run of the (R,C,OH,OW)
    // output: hand the staged block to the TMA engine as a single bulk store
    fence_async_smem();
    __syncwarp();
    if (lane == 0)
      bulk_s2g(out + ((size_t)r * C + (size_t)chunk * 4) * OHW, sbuf, (uint32_t)(STG * sizeof(float)));
    cur = nxt;
    r = rn;
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
// fast backward: 8x8 sample grid.  CTA = one warp = (image, 4 channels, band of rows); the
// band lives in shared memory and belongs to this warp alone.  lane = (channel, pw).
// ----------------------------------------------------------------------------------------
struct BandFlush {
  float *plane;  // this lane's channel band, row-major [rows][W]
  int lo, hi, W;
  int ws;
  float wr;
  bool wv, f1, f2, f4, tail;
};

__device__ __forceinline__ void band_flush(const BandFlush &f, int row, float v) {
  if (row < f.lo || row >= f.hi) return;  // warp-uniform
  float x = f.wv ? v * (1.f - f.wr) : 0.f;
  float y = f.wv ? v * f.wr : 0.f;
  // inclusive segmented scan over runs of equal ws inside each 8-lane channel group
  float tx = __shfl_up_sync(0xffffffffu, x, 1, 8), ty = __shfl_up_sync(0xffffffffu, y, 1, 8);
  if (f.f1) x += tx, y += ty;
  tx = __shfl_up_sync(0xffffffffu, x, 2, 8), ty = __shfl_up_sync(0xffffffffu, y, 2, 8);
  if (f.f2) x += tx, y += ty;
  tx = __shfl_up_sync(0xffffffffu, x, 4, 8), ty = __shfl_up_sync(0xffffffffu, y, 4, 8);
  if (f.f4) x += tx, y += ty;
  float *p = f.plane + (row - f.lo) * f.W + f.ws;
  if (f.tail) p[0] += x;  // run tails hold distinct ws: no two lanes share an address
  __syncwarp();
  if (f.tail) p[1] += y;
  __syncwarp();
}

template <int POOL>
__global__ void __launch_bounds__(32)
    k_align8_bwd_bands(const float *__restrict__ gout, const int *__restrict__ plan,
                       const int *__restrict__ order, const int *__restrict__ img_off, int C,
                       int H, int W, int n_chunks, int n_bands, int band_rows, int accumulate,
                       float *__restrict__ gin) {
  constexpr int OW = POOL == RLOD_POOL_NONE ? 8 : 7;
  constexpr int OHW = OW * OW;
  constexpr int STG = 4 * OHW;
  extern __shared__ __align__(16) float bsm[];
  float *stage = bsm;          // STG floats (16-byte aligned for the float4 staging writes)
  float *band = bsm + STG;     // [4][band_rows * W]
  const int lane = threadIdx.x, pw = lane & 7, ch = lane >> 3;
  int t = blockIdx.x;
  const int bandi = t % n_bands;
  t /= n_bands;
  const int chunk = t % n_chunks;
  const int b = t / n_chunks;
  const int lo = bandi * band_rows, hi = min(H, lo + band_rows);
  const int rows = hi - lo, bstride = band_rows * W;
  const size_t gbase = ((size_t)b * C + (size_t)chunk * 4) * ((size_t)H * W);
  for (int i = lane; i < 4 * rows * W; i += 32) {
    const int c = i / (rows * W), o = i - c * (rows * W);
    band[c * bstride + o] = accumulate ? gin[gbase + (size_t)c * H * W + (size_t)lo * W + o] : 0.f;
  }
  __syncwarp();
  const int r0 = img_off[b], r1 = img_off[b + 1];
  BandFlush f;
  f.plane = band + ch * bstride;
  f.lo = lo, f.hi = hi, f.W = W;
  for (int k = r0; k < r1; ++k) {
    const int r = order[k];
    Plan8 p;
    load_plan8(plan, r, pw, p);
    int hmin = 1 << 30, hmax = -1;
#pragma unroll
    for (int ph = 0; ph < 8; ++ph)
      if (p.hs[ph] >= 0) {
        hmin = min(hmin, p.hs[ph]);
        hmax = max(hmax, p.hs[ph] + 1);
      }
    if (hmax < lo || hmin >= hi) continue;  // warp-uniform: roi does not touch this band
    // grad_out chunk: 4 channels x OHW floats, one contiguous run
    const float *src = gout + ((size_t)r * C + (size_t)chunk * 4) * OHW;
#pragma unroll
    for (int q = lane; q < STG / 4; q += 32)
      *reinterpret_cast<float4 *>(stage + 4 * q) = ld_stream4(src + 4 * q);
    // per-roi run structure of the sample columns (ws is non-decreasing in pw)
    f.wv = p.ws >= 0;
    f.ws = f.wv ? p.ws : 0;
    f.wr = p.wr;
    const int key = f.wv ? p.ws : -100 - pw;
    const int k1 = __shfl_up_sync(0xffffffffu, key, 1, 8);
    const int k2 = __shfl_up_sync(0xffffffffu, key, 2, 8);
    const int k4 = __shfl_up_sync(0xffffffffu, key, 4, 8);
    const int kn = __shfl_down_sync(0xffffffffu, key, 1, 8);
    f.f1 = pw >= 1 && k1 == key;
    f.f2 = pw >= 2 && k2 == key;
    f.f4 = pw >= 4 && k4 == key;
    f.tail = f.wv && (pw == 7 || kn != key);
    __syncwarp();
    // gradient of each sample point of this lane's column: AVG spreads every pooled-bin
    // gradient /4 over its 2x2 sample window (autograd of avg_pool2d(2, stride 1))
    const float *g = stage + ch * OHW;
    float gs[8];
    if (POOL == RLOD_POOL_NONE) {
#pragma unroll
      for (int ph = 0; ph < 8; ++ph) gs[ph] = g[ph * 8 + pw];
    } else {
      float cs[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const float a = pw >= 1 ? g[i * 7 + pw - 1] : 0.f;
        const float bb = pw <= 6 ? g[i * 7 + pw] : 0.f;
        cs[i] = a + bb;
      }
      gs[0] = cs[0] * 0.25f;
#pragma unroll
      for (int ph = 1; ph < 7; ++ph) gs[ph] = (cs[ph - 1] + cs[ph]) * 0.25f;
      gs[7] = cs[6] * 0.25f;
    }
    // walk the sample rows; hs is warp-uniform and non-decreasing, so the two live rows
    // stay in registers and every touched row is flushed exactly once
    int row = -2;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int ph = 0; ph < 8; ++ph) {
      const int h0 = p.hs[ph];
      if (h0 < 0) continue;
      const float c0 = (1.f - p.hr[ph]) * gs[ph], c1 = p.hr[ph] * gs[ph];
      if (h0 == row) {
        a0 += c0;
        a1 += c1;
      } else if (h0 == row + 1) {
        band_flush(f, row, a0);
        a0 = a1 + c0;
        a1 = c1;
        row = h0;
      } else {
        if (row >= 0) {
          band_flush(f, row, a0);
          band_flush(f, row + 1, a1);
        }
        a0 = c0;
        a1 = c1;
        row = h0;
      }
    }
    if (row >= 0) {
      band_flush(f, row, a0);
      band_flush(f, row + 1, a1);
    }
    __syncwarp();
  }
  __syncwarp();
  for (int i = lane; i < 4 * rows * W; i += 32) {
    const int c = i / (rows * W), o = i - c * (rows * W);
    gin[gbase + (size_t)c * H * W + (size_t)lo * W + o] = band[c * bstride + o];
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

static size_t fwd_planes_smem(int H, int W, int pool_mode) {
  const int stg = 4 * (pool_mode == RLOD_POOL_NONE ? 64 : 49);
  return (size_t)16 * ((size_t)H * W + 2 * W) + (size_t)(kFwdThreads / 32) * 2 * stg * 4;
}

}  // namespace rlod

using namespace rlod;

RLOD_API size_t rlod_roi_align_workspace_bytes(int B, int R, int ah, int aw, int pool_mode) {
  const int GH = pool_mode == RLOD_POOL_NONE ? ah : ah + 1;
  const int GW = pool_mode == RLOD_POOL_NONE ? aw : aw + 1;
  if (B < 0 || R < 0 || ah < 1 || aw < 1) return 0;
  return carve_align_ws(nullptr, B, R, GH, GW).bytes;
}

RLOD_API int rlod_roi_align_forward(const float *feat, const float *rois, int B, int C, int H,
                                    int W, int R, int ah, int aw, float spatial_scale,
                                    int pool_mode, float *out, void *workspace,
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
  rc = build_plan(rois, B, H, W, R, GH, GW, spatial_scale, ws, st);
  if (rc) return rc;

  const size_t smem = fwd_planes_smem(H, W, pool_mode);
  const bool fast = GH == 8 && GW == 8 && (C % 4) == 0 && smem <= (size_t)kMaxSmemPerCta &&
                    ((uintptr_t)out % 16) == 0 && R >= 2 * B;
  if (fast) {
    const int n_chunks = C / 4;
    const unsigned grid = (unsigned)(B * n_chunks);
    RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st,
                k_roi_plan8_fwd<<<(unsigned)cdiv(R, 128), 128, 0, st>>>(ws.plan, R, H, W, ws.ext));
#define RLOD_LAUNCH_FWD(POOL)                                                                  \
  do {                                                                                         \
    cudaFuncSetAttribute(k_align8_fwd_planes<POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                         (int)smem);                                                           \
    ProfScope _ps(RLOD_KERNEL_ALIGN_FWD, st);                                                  \
    k_align8_fwd_planes<POOL><<<grid, kFwdThreads, smem, st>>>(feat, ws.ext, ws.order,         \
                                                               ws.img_off, C, H, W, n_chunks,  \
                                                               out);                           \
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
  rc = build_plan(rois, B, H, W, R, GH, GW, spatial_scale, ws, st);
  if (rc) return rc;

  const bool fast = GH == 8 && GW == 8 && (C % 4) == 0 && pool_mode != RLOD_POOL_MAX &&
                    ((uintptr_t)grad_out % 16) == 0 && W <= 2048;
  if (fast) {
    // band height: ~17 KB of shared memory per one-warp CTA -> ~12 resident warps per SM,
    // each owning its band exclusively
    int band_rows = (17 * 1024 / 16) / W;
    if (band_rows < 1) band_rows = 1;
    if (band_rows > H) band_rows = H;
    const int n_bands = (int)cdiv(H, band_rows);
    band_rows = (int)cdiv(H, n_bands);  // even out the bands
    const int n_chunks = C / 4;
    const int stg = 4 * (pool_mode == RLOD_POOL_NONE ? 64 : 49);
    const size_t smem = ((size_t)stg + (size_t)4 * band_rows * W) * sizeof(float);
    const unsigned grid = (unsigned)((long long)B * n_chunks * n_bands);
    if (pool_mode == RLOD_POOL_NONE) {
      cudaFuncSetAttribute(k_align8_bwd_bands<RLOD_POOL_NONE>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      RLOD_LAUNCH(RLOD_KERNEL_ALIGN_BWD, st, k_align8_bwd_bands<RLOD_POOL_NONE><<<grid, 32, smem, st>>>(
          grad_out, ws.plan, ws.order, ws.img_off, C, H, W, n_chunks, n_bands, band_rows,
          accumulate, grad_in));
    } else {
      cudaFuncSetAttribute(k_align8_bwd_bands<RLOD_POOL_AVG>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      RLOD_LAUNCH(RLOD_KERNEL_ALIGN_BWD, st, k_align8_bwd_bands<RLOD_POOL_AVG><<<grid, 32, smem, st>>>(
          grad_out, ws.plan, ws.order, ws.img_off, C, H, W, n_chunks, n_bands, band_rows,
          accumulate, grad_in));
    }
    return launch_status();
  }
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
