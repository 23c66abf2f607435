// RoICrop for sm_100a: the third POOLING_MODE of the reference ('crop', the cfg default,
// lib/model/utils/config.py:283) -- a spatial-transformer style bilinear sampler.
//
//   rlod_affine_grid      _affine_grid_gen (lib/model/utils/net_utils.py:143-165): theta from the
//                         roi ((x2-x1)/(W-1), 0, (x1+x2-W+1)/(W-1); 0, (y2-y1)/(H-1),
//                         (y1+y2-H+1)/(H-1), roi / 16) applied to F.affine_grid's base grid.
//                         The base grid is torch's linspace(-1, 1, g) [align_corners = 1: the
//                         behaviour of the PyTorch 0.x the reference targets] or
//                         (2j + 1) / g - 1 [align_corners = 0: what torch >= 1.3 does by default].
//   rlod_roi_crop_forward BilinearSamplerBHWD_updateOutput_cuda (lib/model/roi_crop/src/
//                         roi_crop_cuda_kernel.cu:47-118): grid (R,gh,gw,2) holds (y, x) in
//                         [-1,1]; xcoord = (x + 1) * (W - 1) / 2, top-left = floor, weight =
//                         1 - frac; taps outside the map contribute 0; roi r samples image
//                         r / (R / B) (the reference ignores the roi's batch column).
//   rlod_roi_crop_backward :120-199: gradient w.r.t. the feature map (the reference computes no
//                         grid gradient: its dot products are never stored).  One RED per tap.
// (grid.y = C / 32 must stay below 65536: C < 2^21.)
// Every fp32 operation rounds separately, in the reference's order (the plane-resident forward contracts the tap sum
// into FMAs, as nvcc does for the reference's kernel: see k_roi_crop_planes).
#include <cstdlib>

#include "rlod_common.cuh"

namespace rlod {

__device__ __forceinline__ void crop_top_left(float x, int width, int &point, float &weight) {
  // (x + 1) * (width - 1) / 2: halving is exact in binary floating point (the correctly rounded value of the same
  // real number), so the multiplication by 0.5 replaces the division sequence (FCHK + Newton step + slow-path call)
  const float xcoord = __fmul_rn(__fmul_rn(__fadd_rn(x, 1.f), (float)(width - 1)), 0.5f);
  const float fl = floorf(xcoord);
  point = (int)fl;
  weight = __fsub_rn(1.f, __fsub_rn(xcoord, fl));
}

__global__ void __launch_bounds__(256)
    k_affine_grid(const float *__restrict__ rois, int R, int H, int W, int g, int align_corners,
                  float *__restrict__ grid) {
  const long long total = (long long)R * g * g;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % g), i = (int)((idx / g) % g);
    const int r = (int)(idx / ((long long)g * g));
    const float *roi = rois + (size_t)r * 5;
    const float x1 = __fdiv_rn(__ldg(roi + 1), 16.f), y1 = __fdiv_rn(__ldg(roi + 2), 16.f);
    const float x2 = __fdiv_rn(__ldg(roi + 3), 16.f), y2 = __fdiv_rn(__ldg(roi + 4), 16.f);
    const float wm = (float)(W - 1), hm = (float)(H - 1);
    const float t00 = __fdiv_rn(__fsub_rn(x2, x1), wm);
    const float t02 = __fdiv_rn(__fadd_rn(__fsub_rn(__fadd_rn(x1, x2), (float)W), 1.f), wm);
    const float t11 = __fdiv_rn(__fsub_rn(y2, y1), hm);
    const float t12 = __fdiv_rn(__fadd_rn(__fsub_rn(__fadd_rn(y1, y2), (float)H), 1.f), hm);
    float xs, ys;
    if (align_corners) {  // torch.linspace(-1, 1, g)
      const float step = g > 1 ? __fdiv_rn(2.f, (float)(g - 1)) : 0.f;
      xs = g > 1 ? __fmaf_rn((float)j, step, -1.f) : -1.f;
      ys = g > 1 ? __fmaf_rn((float)i, step, -1.f) : -1.f;
      if (j == g - 1 && g > 1) xs = 1.f;
      if (i == g - 1 && g > 1) ys = 1.f;
    } else {  // linspace(-1, 1, g) * (g - 1) / g
      xs = __fsub_rn(__fdiv_rn((float)(2 * j + 1), (float)g), 1.f);
      ys = __fsub_rn(__fdiv_rn((float)(2 * i + 1), (float)g), 1.f);
    }
    // base_grid @ theta^T with theta[0][1] = theta[1][0] = 0
    grid[idx * 2 + 0] = __fmaf_rn(xs, t00, t02);
    grid[idx * 2 + 1] = __fmaf_rn(ys, t11, t12);
  }
}

// CTA = (roi, chunk of channels).  A thread owns sample points of the roi's gh x gw grid: the
// top-left tap offset, the four tap masks and the four weight products are computed ONCE and
// reused for every channel of the chunk, so the per-channel work is 4 gathers (the roi's patch
// of a plane: L1 / L2 hits, the image's planes stay L2-resident because its rois are launched
// together), 7 fp32 ops in the reference's rounding order and one coalesced store -- no index
// arithmetic per element.  Backward: the same walk with one RED per valid tap.
constexpr int kCropThreads = 224;  // 7 warps: a 14 x 14 grid is 196 sample points
constexpr int kCropChunk = 32;     // channels per CTA

template <bool BWD>
__global__ void __launch_bounds__(kCropThreads)
    k_roi_crop(const float *__restrict__ feat_or_gout, const float *__restrict__ grid, int B, int C, int H,
               int W, int R, int gh, int gw, int roi_per_image, float *__restrict__ out_or_gfeat) {
  const int r = blockIdx.x, c0 = blockIdx.y * kCropChunk;
  const int c1 = min(C, c0 + kCropChunk);
  const int b = r / roi_per_image;
  const int ghw = gh * gw;
  const size_t HW = (size_t)H * W;
  const bool valid_b = b >= 0 && b < B;
  for (int p = threadIdx.x; p < ghw; p += kCropThreads) {
    const float *gp = grid + ((size_t)r * ghw + p) * 2;
    const float yf = __ldg(gp), xf = __ldg(gp + 1);
    int x0, y0;
    float xw, yw;
    crop_top_left(xf, W, x0, xw);
    crop_top_left(yf, H, y0, yw);
    const bool xin0 = x0 >= 0 && x0 <= W - 1, xin1 = x0 + 1 >= 0 && x0 + 1 <= W - 1;
    const bool yin0 = y0 >= 0 && y0 <= H - 1, yin1 = y0 + 1 >= 0 && y0 + 1 <= H - 1;
    const bool m00 = valid_b && xin0 && yin0, m01 = valid_b && xin1 && yin0;
    const bool m10 = valid_b && xin0 && yin1, m11 = valid_b && xin1 && yin1;
    const float ixw = __fsub_rn(1.f, xw), iyw = __fsub_rn(1.f, yw);
    const float w00 = __fmul_rn(xw, yw), w01 = __fmul_rn(ixw, yw), w10 = __fmul_rn(xw, iyw), w11 = __fmul_rn(ixw, iyw);
    // masked-off taps are never dereferenced: point them at the plane's first pixel
    const long long tl = (long long)y0 * W + x0;
    const long long o00 = m00 ? tl : 0, o01 = m01 ? tl + 1 : 0, o10 = m10 ? tl + W : 0, o11 = m11 ? tl + W + 1 : 0;
    const size_t plane0 = ((size_t)(valid_b ? b : 0) * C + c0) * HW;
    const size_t io0 = ((size_t)r * C + c0) * ghw + p;  // out (fwd) / grad_out (bwd) element of channel c0
    if (!BWD) {
      const float *pl = feat_or_gout + plane0;
      float *o = out_or_gfeat + io0;
#pragma unroll 4
      for (int c = c0; c < c1; ++c, pl += HW, o += ghw) {
        const float a = m00 ? __ldg(pl + o00) : 0.f, bb = m01 ? __ldg(pl + o01) : 0.f;
        const float cc = m10 ? __ldg(pl + o10) : 0.f, d = m11 ? __ldg(pl + o11) : 0.f;
        float v = __fmul_rn(w00, a);
        v = __fadd_rn(v, __fmul_rn(w01, bb));
        v = __fadd_rn(v, __fmul_rn(w10, cc));
        v = __fadd_rn(v, __fmul_rn(w11, d));
        // no tap inside the map (or a bad image index): the reference writes 0
        __stcs(o, v);
      }
    } else if (valid_b) {
      float *pl = out_or_gfeat + plane0;
      const float *gq = feat_or_gout + io0;
#pragma unroll 4
      for (int c = c0; c < c1; ++c, pl += HW, gq += ghw) {
        const float go = __ldcs(gq);
        if (m00) atomicAdd(pl + o00, __fmul_rn(w00, go));
        if (m01) atomicAdd(pl + o01, __fmul_rn(w01, go));
        if (m10) atomicAdd(pl + o10, __fmul_rn(w10, go));
        if (m11) atomicAdd(pl + o11, __fmul_rn(w11, go));
      }
    }
  }
}

// ----------------------------------------------------------------------------------------
// backward as a GATHER when the roi's sampling grid is separable and monotone -- what
// _affine_grid_gen always produces (theta has no rotation, net_utils.py:143-165): x depends
// only on the sample column, y only on the sample row.  The scatter above issues 4 REDs per
// (sample, channel): 784 per (roi, channel) at 14 x 14, although the roi's taps cover only
// ~100 distinct pixels, and the kernel is RED-issue bound.  Here a thread owns one PIXEL of the
// roi's tap patch and sums the samples that touch it (a contiguous range of sample rows x a
// contiguous range of sample columns, weights tabulated once per roi in shared memory), then
// adds the sum with ONE RED: 6-8x fewer atomics.  grad_out tiles of 32 channels are staged in
// shared memory (coalesced); a warp owns pixels, its lanes the tile's channels.  Any other grid (not separable, not monotone, too large) takes the
// scatter walk inside the same launch, CTA by CTA.  Per-term rounding follows the reference
// ((wx * wy) * grad, roi_crop_cuda_kernel.cu:172-197); the order of the additions differs (as
// it does between two runs of the reference's atomics).
// ----------------------------------------------------------------------------------------
constexpr int kCropBwdThreads = 256;
constexpr int kCropTile = 32;      // channels staged per step: the lanes of a warp
constexpr int kCropBwdChunk = 256;  // channels per CTA (the per-roi tables are built once per CTA)
constexpr int kCropMaxG = 16;      // separable path: grid sides up to 16

__global__ void __launch_bounds__(kCropBwdThreads)
    k_roi_crop_bwd_gather(const float *__restrict__ gout, const float *__restrict__ grid, int B, int C, int H,
                          int W, int R, int gh, int gw, int roi_per_image, float *__restrict__ gfeat) {
  extern __shared__ __align__(16) unsigned char crop_smem[];
  const int r = blockIdx.x, c0 = blockIdx.y * kCropBwdChunk;
  const int c1 = min(C, c0 + kCropBwdChunk);
  const int b = r / roi_per_image;
  if (b < 0 || b >= B) return;  // the reference's scatter writes nothing for such a roi
  const int ghw = gh * gw, t = threadIdx.x;
  const size_t HW = (size_t)H * W;
  __shared__ int s_y0[kCropMaxG], s_x0[kCropMaxG];
  __shared__ float s_yw[kCropMaxG], s_xw[kCropMaxG];
  __shared__ int s_ok, s_box[4];
  if (t == 0) s_ok = (gh <= kCropMaxG && gw <= kCropMaxG) ? 1 : 0;
  __syncthreads();
  const float *gr = grid + (size_t)r * ghw * 2;
  // separable?  (bitwise: the generator computes a column's x once per sample row from the same
  // operands)
  if (s_ok) {
    bool ok = true;
    for (int p = t; p < ghw; p += kCropBwdThreads) {
      const int yo = p / gw, xo = p - yo * gw;
      ok = ok && __ldg(gr + 2 * p) == __ldg(gr + 2 * (yo * gw)) && __ldg(gr + 2 * p + 1) == __ldg(gr + 2 * xo + 1);
    }
    if (!ok) s_ok = 0;
  }
  __syncthreads();
  if (s_ok) {
    if (t < gh) crop_top_left(__ldg(gr + 2 * (t * gw)), H, s_y0[t], s_yw[t]);
    if (t >= 32 && t < 32 + gw) crop_top_left(__ldg(gr + 2 * (t - 32) + 1), W, s_x0[t - 32], s_xw[t - 32]);
  }
  __syncthreads();
  if (s_ok && t == 0) {
    bool mono = true;
    for (int i = 1; i < gh; ++i) mono = mono && s_y0[i] >= s_y0[i - 1];
    for (int i = 1; i < gw; ++i) mono = mono && s_x0[i] >= s_x0[i - 1];
    // |top-left| must stay far from the int range (a wild theta): the patch arithmetic below adds 1
    mono = mono && s_y0[0] > -(1 << 28) && s_y0[gh - 1] < (1 << 28) && s_x0[0] > -(1 << 28) && s_x0[gw - 1] < (1 << 28);
    if (!mono) s_ok = 0;
    // tap patch clipped to the map
    s_box[0] = max(s_y0[0], 0), s_box[1] = min(s_y0[gh - 1] + 1, H - 1);
    s_box[2] = max(s_x0[0], 0), s_box[3] = min(s_x0[gw - 1] + 1, W - 1);
  }
  __syncthreads();
  if (!s_ok) {
    // scatter walk (k_roi_crop<true>'s loop) for this CTA
    for (int p = t; p < ghw; p += kCropBwdThreads) {
      const float yf = __ldg(gr + 2 * p), xf = __ldg(gr + 2 * p + 1);
      int x0, y0;
      float xw, yw;
      crop_top_left(xf, W, x0, xw);
      crop_top_left(yf, H, y0, yw);
      const bool xin0 = x0 >= 0 && x0 <= W - 1, xin1 = x0 + 1 >= 0 && x0 + 1 <= W - 1;
      const bool yin0 = y0 >= 0 && y0 <= H - 1, yin1 = y0 + 1 >= 0 && y0 + 1 <= H - 1;
      const float ixw = __fsub_rn(1.f, xw), iyw = __fsub_rn(1.f, yw);
      const float w00 = __fmul_rn(xw, yw), w01 = __fmul_rn(ixw, yw), w10 = __fmul_rn(xw, iyw), w11 = __fmul_rn(ixw, iyw);
      const long long tl = (long long)y0 * W + x0;
      float *pl = gfeat + ((size_t)b * C + c0) * HW;
      const float *gq = gout + ((size_t)r * C + c0) * ghw + p;
      for (int c = c0; c < c1; ++c, pl += HW, gq += ghw) {
        const float go = __ldcs(gq);
        if (xin0 && yin0) atomicAdd(pl + tl, __fmul_rn(w00, go));
        if (xin1 && yin0) atomicAdd(pl + tl + 1, __fmul_rn(w01, go));
        if (xin0 && yin1) atomicAdd(pl + tl + W, __fmul_rn(w10, go));
        if (xin1 && yin1) atomicAdd(pl + tl + W + 1, __fmul_rn(w11, go));
      }
    }
    return;
  }
  const int py0 = s_box[0], py1 = s_box[1], px0 = s_box[2], px1 = s_box[3];
  const int nPy = py1 - py0 + 1, nPx = px1 - px0 + 1;
  if (nPy <= 0 || nPx <= 0) return;  // every tap outside the map
  // per patch row / column: first contributing sample, their count, and the weights
  // (pitch gh + 1 / gw + 1: odd for the 14-point grid, so neighbouring pixels' lists do not
  // share banks)
  const int pitch_y = gh + 1, pitch_x = gw + 1;
  float *wys = reinterpret_cast<float *>(crop_smem);             // [nPy][pitch_y]
  float *wxs = wys + (size_t)H * pitch_y;                         // [nPx][pitch_x]
  short *ya = reinterpret_cast<short *>(wxs + (size_t)W * pitch_x);  // [H] first sample, [H] count
  short *xa = ya + 2 * H;                                            // [W], [W]
  float *tile = reinterpret_cast<float *>(xa + 2 * W + ((2 * H + 2 * W) & 1));  // [kCropTile][ghw]
  for (int j = t; j < nPy + nPx; j += kCropBwdThreads) {
    const bool isy = j < nPy;
    const int pix = isy ? py0 + j : px0 + (j - nPy);
    const int G = isy ? gh : gw;
    const int *tl = isy ? s_y0 : s_x0;
    const float *tw = isy ? s_yw : s_xw;
    float *wrow = isy ? wys + (size_t)j * pitch_y : wxs + (size_t)(j - nPy) * pitch_x;
    int first = -1, n = 0;
    for (int q = 0; q < G; ++q) {
      const int d = pix - tl[q];  // 0: the sample's top-left tap, 1: its second tap
      if (d == 0 || d == 1) {
        if (first < 0) first = q;
        wrow[n++] = d == 0 ? tw[q] : __fsub_rn(1.f, tw[q]);
      }
    }
    short *lst = isy ? ya : xa;
    const int L = isy ? H : W, jj = isy ? j : j - nPy;
    lst[jj] = (short)(first < 0 ? 0 : first);
    lst[L + jj] = (short)n;
  }
  // a warp owns patch pixels, its lanes the 32 channels of the staged tile: no divergence (all
  // lanes run the same sample ranges), weights are broadcast reads, the tile is read with a
  // channel pitch of ghw + 1 floats (odd: 32 channels, 32 banks)
  const int nP = nPy * nPx, tp = ghw + 1;
  const int warp = t >> 5, lane = t & 31;
  for (int cb = c0; cb < c1; cb += kCropTile) {
    const int nc = min(kCropTile, c1 - cb);
    __syncthreads();  // tables ready / the previous tile is consumed
    {
      const float *src = gout + ((size_t)r * C + cb) * ghw;  // nc * ghw contiguous floats
      // element e of the run lands at e + e / ghw (channel pitch ghw + 1): the quotient advances
      // incrementally, a division per element was 16 % of the kernel's instructions
      const int dq = kCropBwdThreads / ghw, dr = kCropBwdThreads - dq * ghw;
      int c = t / ghw, rem = t - c * ghw;
      for (int e = t; e < nc * ghw; e += kCropBwdThreads) {
        tile[e + c] = __ldcs(src + e);
        c += dq, rem += dr;
        if (rem >= ghw) rem -= ghw, ++c;
      }
    }
    __syncthreads();
    if (lane < nc) {
      float *plane = gfeat + ((size_t)b * C + cb + lane) * HW;
      const float *gl = tile + lane * tp;
      const float inv_npx = 1.f / (float)nPx;
      for (int q = warp; q < nP; q += kCropBwdThreads / 32) {
        int j = (int)(((float)q + 0.5f) * inv_npx);  // q / nPx for q < 2^16, fixed up below
        int i = q - j * nPx;
        if (i < 0) --j, i += nPx;
        if (i >= nPx) ++j, i -= nPx;
        const int yf = ya[j], ny = ya[H + j], xf = xa[i], nx = xa[W + i];
        const float *g = gl + yf * gw + xf;
        const float *wy = wys + (size_t)j * pitch_y, *wx = wxs + (size_t)i * pitch_x;
        float acc = 0.f;
        // nx, ny are 1-5 for most pixels: plain loops (the compiler's 8/4/2/1 unrolling spent more
        // instructions on remainder dispatch than on terms)
#pragma unroll 1
        for (int u = 0; u < ny; ++u, g += gw) {
          const float wyu = wy[u];
#pragma unroll 1
          for (int v = 0; v < nx; ++v) acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(wx[v], wyu), g[v]));
        }
        if (ny > 0 && nx > 0) atomicAdd(plane + (size_t)(py0 + j) * W + (px0 + i), acc);
      }
    }
  }
}

static unsigned crop_grid(long long total) {
  const long long blocks = (total + 255) / 256;
  return (unsigned)(blocks < (1LL << 22) ? blocks : (1LL << 22));
}

// ----------------------------------------------------------------------------------------
// forward, plane-resident (round 2, second session): the RoIAlign forward's structure for the crop sampler.
// CTA = (image, 4 G channels), G = 1 or 2: the planes are staged in shared memory once, interleaved per pixel in
// groups of four (one LDS.128 = a tap of 4 channels), with a ZERO FRAME so that a tap outside the map needs no
// mask: padded pixel (y, x), y in [-1, H + 1], x in [-1, W - 1], sits at (y + 1) * P + (x + 1) with P = W + 1 --
// the pixel right of column W - 1 is the next row's (zero) left border.  A warp owns a contiguous range of the
// image's rois and walks their sample points as ONE flat list, 32 per iteration (no ragged last iteration per
// roi: 196 points are 6.125 warps), so the grid is read as one contiguous stream (8 bytes per point, shared by
// the image's CTAs through L2) and consecutive lanes store consecutive floats of an output plane.  A sample
// whose four taps all lie outside the map reads the zero rows below it.  The coordinate arithmetic of a point
// is done once for 4 G channels; the taps are summed in the reference's order, two channels per packed
// instruction: v = w00 a; v = fma(w01, b, v); v = fma(w10, c, v); v = fma(w11, d, v) -- the contraction nvcc
// applies to the reference's own expression (roi_crop_cuda_kernel.cu:103-106); ptxas 12.9 contracts
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2 regardless, so it is written out.  Against the gather kernel above
// (separate roundings) the results differ in the last bit; both sit within 1e-5 of the reference.
// Any grid is accepted (no separability assumed).  The gather kernel remains for C % 4 != 0, R % B != 0 and maps
// beyond shared memory.
// ----------------------------------------------------------------------------------------
constexpr int kCropPlWarps = 16;
constexpr int kCropPlThreads = kCropPlWarps * 32;

__device__ __forceinline__ unsigned long long crop_pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long crop_mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long crop_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// GHW = 196: the 14 x 14 grid of crop_pool (POOLING_SIZE * 2) as a compile-time constant -- the eight output planes
// of a point are then immediate offsets from one address; GHW = 0: any grid
template <int G, int GHW>
__global__ void __launch_bounds__(kCropPlThreads, 2)
    k_roi_crop_planes(const float *__restrict__ feat, const float *__restrict__ grid, int C, int H, int W,
                      int ghw_arg, int rpi, int n_chunks, float *__restrict__ out) {
  const int ghw = GHW ? GHW : ghw_arg;
  extern __shared__ __align__(128) unsigned char crop_raw[];
  float4 *planes4 = reinterpret_cast<float4 *>(crop_raw);  // [G][(H + 3) * P]
  const int P = W + 1, HW = H * W;
  const int n_pad = (H + 3) * P;
  const int b = blockIdx.x / n_chunks, chunk = blockIdx.x - b * n_chunks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < G * n_pad; i += kCropPlThreads) planes4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
#pragma unroll
  for (int g = 0; g < G; ++g)
    fill_planes4_async<kCropPlThreads>(planes4 + g * n_pad + P + 1,
                                       feat + ((size_t)b * C + (size_t)chunk * 4 * G + 4 * g) * HW, H, W, P, HW);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const uint32_t pbase = smem_u32(planes4);
  const uint32_t zero_tl = (uint32_t)((H + 1) * P) * 16u;  // a 2 x 2 block of zero pixels below the map
  const uint32_t gstride = (uint32_t)n_pad * 16u, rowb = (uint32_t)P * 16u;
  const int r_lo = (int)((long long)rpi * warp / kCropPlWarps), r_hi = (int)((long long)rpi * (warp + 1) / kCropPlWarps);
  const long long nq = (long long)(r_hi - r_lo) * ghw;
  const float2 *gp = reinterpret_cast<const float2 *>(grid) + ((size_t)b * rpi + r_lo) * ghw;
  float *ob = out + (((size_t)b * rpi + r_lo) * C + (size_t)chunk * 4 * G) * ghw;
  const size_t roi_stride = (size_t)C * ghw;
  int p = lane;
  float *orow = ob;
  while (p >= ghw) p -= ghw, orow += roi_stride;
  // the grid point of the NEXT iteration is fetched one iteration ahead: everything below depends on it, and an L2
  // round trip per iteration left the warps waiting on it (long-scoreboard stall 7.9 per issue before)
  float2 yx_next = lane < nq ? __ldg(gp + lane) : make_float2(0.f, 0.f);
#pragma unroll 2
  for (long long q = lane; q < nq; q += 32) {
    const float2 yx = yx_next;
    if (q + 32 < nq) yx_next = __ldg(gp + q + 32);
    int x0, y0;
    float xw, yw;
    crop_top_left(yx.y, W, x0, xw);
    crop_top_left(yx.x, H, y0, yw);
    const bool any = x0 >= -1 && x0 <= W - 1 && y0 >= -1 && y0 <= H - 1;
    const uint32_t tl = pbase + (any ? (uint32_t)((y0 + 1) * P + (x0 + 1)) * 16u : zero_tl);
    const float ixw = __fsub_rn(1.f, xw), iyw = __fsub_rn(1.f, yw);
    const float w00 = __fmul_rn(xw, yw), w01 = __fmul_rn(ixw, yw), w10 = __fmul_rn(xw, iyw), w11 = __fmul_rn(ixw, iyw);
    const unsigned long long q00 = crop_pack2(w00, w00), q01 = crop_pack2(w01, w01);
    const unsigned long long q10 = crop_pack2(w10, w10), q11 = crop_pack2(w11, w11);
    float *o = orow + p;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const uint32_t t = tl + (uint32_t)g * gstride;
      unsigned long long a0, a1, b0, b1, c0, c1, d0, d1;
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a0), "=l"(a1) : "r"(t));
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(b0), "=l"(b1) : "r"(t + 16u));
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(c0), "=l"(c1) : "r"(t + rowb));
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(d0), "=l"(d1) : "r"(t + rowb + 16u));
      unsigned long long v0 = crop_mul2(q00, a0), v1 = crop_mul2(q00, a1);
      v0 = crop_fma2(q01, b0, v0), v1 = crop_fma2(q01, b1, v1);
      v0 = crop_fma2(q10, c0, v0), v1 = crop_fma2(q10, c1, v1);
      v0 = crop_fma2(q11, d0, v0), v1 = crop_fma2(q11, d1, v1);
      float2 lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo.x), "=f"(lo.y) : "l"(v0));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(hi.x), "=f"(hi.y) : "l"(v1));
      float *og = o + (size_t)(4 * g) * ghw;
      __stcs(og, lo.x), __stcs(og + ghw, lo.y), __stcs(og + 2 * (size_t)ghw, hi.x), __stcs(og + 3 * (size_t)ghw, hi.y);
    }
    p += 32;
    while (p >= ghw) p -= ghw, orow += roi_stride;
  }
}

}  // namespace rlod

using namespace rlod;

RLOD_API int rlod_affine_grid(const float *rois, int R, int H, int W, int grid_size, int align_corners,
                              float *grid_xy, rlod_stream_t stream) {
  if (R < 0 || H < 2 || W < 2 || grid_size < 1) return RLOD_EINVAL;
  if (R == 0) return RLOD_OK;
  if (!rois || !grid_xy) return RLOD_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  RLOD_LAUNCH(RLOD_KERNEL_CROP, st,
              k_affine_grid<<<crop_grid((long long)R * grid_size * grid_size), 256, 0, st>>>(rois, R, H, W, grid_size,
                                                                                            align_corners, grid_xy));
  return launch_status();
}

RLOD_API int rlod_roi_crop_forward(const float *feat, const float *grid_yx, int B, int C, int H, int W,
                                   int R, int gh, int gw, float *out, rlod_stream_t stream) {
  if (B < 1 || C < 0 || H < 1 || W < 1 || R < 0 || gh < 1 || gw < 1) return RLOD_EINVAL;
  if (R == 0 || C == 0) return RLOD_OK;
  if (!feat || !grid_yx || !out) return RLOD_EINVAL;
  if (R < B) return RLOD_EINVAL;  // roiPerImage = R / B = 0 divides by zero in the reference
  cudaStream_t st = (cudaStream_t)stream;
  {
    // plane kernel: every image has the same number of rois and the zero-framed planes of 4 (G = 1) or 8 (G = 2)
    // channels fit one CTA's shared memory; G = 2 where two CTAs of it still share an SM
    const int ghw = gh * gw;
    const size_t plane = (size_t)(H + 3) * (W + 1) * 16;
    static const bool crop_v1 = getenv("RLOD_CROP_V1") != nullptr;  // A/B switch: the gather kernel
    static const int force_g = getenv("RLOD_CROP_G") ? atoi(getenv("RLOD_CROP_G")) : 0;
    if (!crop_v1 && (C % 4) == 0 && (R % B) == 0 && plane <= (size_t)kMaxSmemPerCta && (long long)B * (C / 4) < (1LL << 31) &&
        ((uintptr_t)grid_yx % 8) == 0) {
      int G = ((C % 8) == 0 && 2 * plane <= (size_t)kMaxSmemPerCta / 2) ? 2 : 1;
      if (force_g == 1 || (force_g == 2 && (C % 8) == 0 && 2 * plane <= (size_t)kMaxSmemPerCta)) G = force_g;
      static bool attr_set = false;
      if (!attr_set) {
        cudaFuncSetAttribute(k_roi_crop_planes<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta);
        cudaFuncSetAttribute(k_roi_crop_planes<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta);
        cudaFuncSetAttribute(k_roi_crop_planes<1, 196>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta);
        cudaFuncSetAttribute(k_roi_crop_planes<2, 196>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta);
        attr_set = true;
      }
      const int n_chunks = C / (4 * G);
      static const bool any_ghw = getenv("RLOD_CROP_GHW0") != nullptr;  // A/B switch: the run-time grid size
      const bool g196 = ghw == 196 && !any_ghw;
#define RLOD_CROP_PLANES(GG, HW_)                                                                              \
  RLOD_LAUNCH(RLOD_KERNEL_CROP, st,                                                                            \
              k_roi_crop_planes<GG, HW_><<<(unsigned)(B * n_chunks), kCropPlThreads, GG * plane, st>>>(        \
                  feat, grid_yx, C, H, W, ghw, R / B, n_chunks, out))
      if (G == 2 && g196) RLOD_CROP_PLANES(2, 196);
      else if (G == 2) RLOD_CROP_PLANES(2, 0);
      else if (g196) RLOD_CROP_PLANES(1, 196);
      else RLOD_CROP_PLANES(1, 0);
#undef RLOD_CROP_PLANES
      return launch_status();
    }
  }
  RLOD_LAUNCH(RLOD_KERNEL_CROP, st,
              k_roi_crop<false><<<dim3((unsigned)R, (unsigned)((C + kCropChunk - 1) / kCropChunk)), kCropThreads, 0, st>>>(
                  feat, grid_yx, B, C, H, W, R, gh, gw, R / B, out));
  return launch_status();
}

RLOD_API int rlod_roi_crop_backward(const float *grad_out, const float *grid_yx, int B, int C, int H, int W,
                                    int R, int gh, int gw, int accumulate, float *grad_feat,
                                    rlod_stream_t stream) {
  if (B < 1 || C < 0 || H < 1 || W < 1 || R < 0 || gh < 1 || gw < 1) return RLOD_EINVAL;
  if (C == 0) return RLOD_OK;
  if (!grad_feat) return RLOD_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) cudaMemsetAsync(grad_feat, 0, (size_t)B * C * H * W * sizeof(float), st);
  if (R == 0) return launch_status();
  if (!grad_out || !grid_yx || R < B) return RLOD_EINVAL;
  {
    // (H + W) rows of the weight tables, the per-row / per-column sample lists, one tile of
    // kCropTile channels; grids or maps too large for that take the scatter kernel
    const size_t smem = ((size_t)H * (gh + 1) + (size_t)W * (gw + 1)) * sizeof(float) +
                        (size_t)(2 * H + 2 * W + 1) * sizeof(short) + (size_t)kCropTile * (gh * gw + 1) * sizeof(float) + 16;
    const dim3 grid_dim((unsigned)R, (unsigned)((C + kCropChunk - 1) / kCropChunk));
    const dim3 grid_gather((unsigned)R, (unsigned)((C + kCropBwdChunk - 1) / kCropBwdChunk));
    if (gh <= kCropMaxG && gw <= kCropMaxG && H < 32768 && W < 32768 && smem <= 200 * 1024) {
      cudaFuncSetAttribute(k_roi_crop_bwd_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      RLOD_LAUNCH(RLOD_KERNEL_CROP, st,
                  k_roi_crop_bwd_gather<<<grid_gather, kCropBwdThreads, smem, st>>>(grad_out, grid_yx, B, C, H, W, R, gh, gw,
                                                                                R / B, grad_feat));
    } else {
      RLOD_LAUNCH(RLOD_KERNEL_CROP, st,
                  k_roi_crop<true><<<grid_dim, kCropThreads, 0, st>>>(grad_out, grid_yx, B, C, H, W, R, gh, gw, R / B,
                                                                     grad_feat));
    }
  }
  return launch_status();
}
