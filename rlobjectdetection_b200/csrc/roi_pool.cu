// RoIPool (max) forward + backward for sm_100a.
//
// Semantics: lib/model/roi_pooling/src/roi_pooling_kernel.cu:24-93 (forward: round() roi
// corners, max(...,1) sizes, floor/ceil bin edges clipped to the map, strict '>' scan h then
// w so the first maximum wins, argmax = flat index into the whole NCHW tensor, empty bin ->
// 0 / -1) and :128-203 (backward: each input element receives the top_diff of every bin
// whose argmax points at it).
//
// The reference's backward is a gather that makes every one of B*C*H*W threads loop over
// ALL R rois (O(B*C*H*W*R)); here the saved argmax turns it into a single pass over the
// R*C*ph*pw gradients with at most one fp32 RED each -- the only work that exists -- while
// keeping the gather's visiting rules, so malformed rois lose their gradient exactly as there.
#include <type_traits>

#include "roi_lists.cuh"

namespace rlod {

// ----------------------------------------------------------------------------------------
// fast forward for 7x7 bins: CTA = (image, 4 channel planes resident in shared memory), like
// the RoIAlign forward.  The planes are read from HBM once (async copies, 4 channels
// interleaved per pixel); a warp serves 4 rois per iteration, lane k of a roi owns output
// column k and scans its bins row by row (LDS.128 = 4 channels per pixel), keeping the first
// maximum in (h, w) scan order with strict '>' exactly like the reference
// (roi_pooling_kernel.cu:73-90).  out and argmax leave as one 784-byte bulk async store
// each per (roi, 4 channels).
// k_pool_plan: one thread per roi, the reference's bin arithmetic (:45-66) once per roi:
//   record[0..6] hstart, [7..13] hend, [14..20] wstart, [21..27] wend (already clipped),
//   [28] batch index or -1, [29] max rows of a bin, [30] max columns of a bin, [31] shape key.
// ----------------------------------------------------------------------------------------
constexpr int kPoolWarps = 8;
constexpr int kPoolThreads = kPoolWarps * 32;
constexpr int kPoolSlot = 200;  // floats per staged tile slot (196 used), = 8 (mod 32)

__global__ void k_pool_plan(const float *__restrict__ rois, int R, int B, int H, int W, float scale,
                            AlignWs ws) {
  pdl_trigger();  // the list kernels behind this one are set up now; they wait for the plan themselves
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float *roi = rois + (size_t)r * 5;
  const float bf = roi[0];
  const int bi = (int)bf;
  const bool bvalid = (bf >= 0.f) && (bi < B);
  const int roi_start_w = (int)roundf(__fmul_rn(roi[1], scale));
  const int roi_start_h = (int)roundf(__fmul_rn(roi[2], scale));
  const int roi_end_w = (int)roundf(__fmul_rn(roi[3], scale));
  const int roi_end_h = (int)roundf(__fmul_rn(roi[4], scale));
  const int roi_width = max(roi_end_w - roi_start_w + 1, 1);
  const int roi_height = max(roi_end_h - roi_start_h + 1, 1);
  const float bin_h = __fdiv_rn((float)roi_height, 7.f);
  const float bin_w = __fdiv_rn((float)roi_width, 7.f);
  int *e = ws.plan + (size_t)r * 32;
  int mr = 0, mc = 0;
  for (int p = 0; p < 7; ++p) {
    int hstart = (int)floorf(__fmul_rn((float)p, bin_h));
    int wstart = (int)floorf(__fmul_rn((float)p, bin_w));
    int hend = (int)ceilf(__fmul_rn((float)(p + 1), bin_h));
    int wend = (int)ceilf(__fmul_rn((float)(p + 1), bin_w));
    hstart = min(max(hstart + roi_start_h, 0), H);
    hend = min(max(hend + roi_start_h, 0), H);
    wstart = min(max(wstart + roi_start_w, 0), W);
    wend = min(max(wend + roi_start_w, 0), W);
    e[p] = hstart, e[7 + p] = hend, e[14 + p] = wstart, e[21 + p] = wend;
    mr = max(mr, hend - hstart), mc = max(mc, wend - wstart);
  }
  e[28] = bvalid ? bi : -1;
  e[29] = bvalid ? mr : 0;
  e[30] = bvalid ? mc : 0;
  // shape key (8 x 8 classes: bin rows, bin column steps of 4) so that the four rois a warp serves
  // together scan the same window: the kernel's loops run to the LONGEST bin of the four in each
  // direction, a tall-thin and a wide-short roi of equal area would cost their product
  // (columns: the widths the kernel's scan loop distinguishes -- 1, 2, up to 4, then steps of 4)
  const int kr = bvalid ? min(mr, 8) : 0;
  const int kc = !bvalid ? 0 : (mc <= 1 ? 0 : (mc == 2 ? 1 : min(((mc + 3) >> 2) + 1, 7)));
  e[31] = (kr > 0 ? kr - 1 : 0) * 8 + kc;
  roi_list_mark(rois, r, R, B, bvalid ? bi : 0, ws);
}

// AM = false: the caller passed no argmax buffer (inference: nothing will be back-propagated) -- the scan keeps
// only the running maximum (one FMNMX per pixel and channel instead of compare + two selects)
template <bool AM>
__global__ void __launch_bounds__(kPoolThreads, 2)
    k_roi_pool7_fwd_planes(const float *__restrict__ feat, const int *__restrict__ rec,
                           const int *__restrict__ order, const int *__restrict__ img_off, int C,
                           int H, int W, int P, int n_chunks, int channels_last,
                           float *__restrict__ out, int *__restrict__ argmax) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *planes4 = reinterpret_cast<float4 *>(smem_raw);
  const int HW = H * W;
  float *stage = reinterpret_cast<float *>(planes4 + H * P);  // [warps][4 slots][2 tiles][kPoolSlot]
  const int b = blockIdx.x / n_chunks, chunk = blockIdx.x - b * n_chunks;
  // the fill reads the feature map only and runs while the plan / list kernels launched just before are still
  // at work (programmatic dependent launch); pdl_wait() below is where this grid meets their results
  if (channels_last) {
    fill_planes4_nhwc_async<kPoolThreads>(planes4, feat + (size_t)b * HW * C + (size_t)chunk * 4, H, W, P, C);
  } else {
    const float *src = feat + ((size_t)b * C + (size_t)chunk * 4) * HW;
    fill_planes4_async<kPoolThreads>(planes4, src, H, W, P, HW);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = lane & 7, slot = lane >> 3;
  const uint32_t pbase = smem_u32(planes4);
  float *tile_v = stage + ((warp * 4 + slot) * 2) * kPoolSlot;
  int *tile_i = reinterpret_cast<int *>(tile_v + kPoolSlot);
  const int chan_base = (b * C + chunk * 4) * HW;  // flat NCHW index of this CTA's first plane
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  pdl_wait();
  const int r0 = img_off[b], r1 = img_off[b + 1];
  if (r0 >= r1) return;  // an image without rois (uniform over the CTA)
  const int n_groups = (r1 - r0 + 3) >> 2;
  __syncthreads();

  for (int g = warp, it = 0; g < n_groups; g += kPoolWarps, ++it) {
    const int kk = r0 + 4 * g + slot;
    const int r = kk < r1 ? __ldg(order + kk) : -1;
    // this lane's column bin; the roi's row bins are read from the record bin by bin (L1 hits: the loop over
    // the row bins below is NOT unrolled -- three scan widths x seven unrolled bins were 112 KB of code)
    int ws = 0, we = 0, mr = 0, mc = 0;
    const int *e = rec + (size_t)(r >= 0 ? r : 0) * 32;
    if (r >= 0) {
      if (k < 7) ws = __ldg(e + 14 + k), we = __ldg(e + 21 + k);
      if (__ldg(e + 28) >= 0) mr = __ldg(e + 29), mc = __ldg(e + 30);
      else we = ws;  // batch index out of range: every bin empty
    }
    const int mrw = __reduce_max_sync(0xffffffffu, mr), mcw = __reduce_max_sync(0xffffffffu, mc);
    // the staging tiles were handed to the bulk-copy engine one iteration ago
    if (it >= 1) {
      if (k == 0) bulk_wait_read<0>();
      __syncwarp();
    }
    // STEP pixels of a bin row per step (1, 2 or 4: the widest bin of the four rois decides, and the lists are
    // ordered by that class -- a masked pixel costs as much as a real one, and half of the rois of a detector
    // have bins of one or two columns): the loads are independent (predicated, a pixel outside the bin keeps
    // the -FLT_MAX sentinel and can never win a strict '>'), then the compare chain in scan order
    auto scan_bins = [&](auto stepc, auto loopc) {
    constexpr int STEP = decltype(stepc)::value;
    constexpr bool LOOP = decltype(loopc)::value;  // false: the widest bin fits one step
#pragma unroll 1
    for (int ph = 0; ph < 7; ++ph) {
      float4 mv = make_float4(-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f);
      int4 mi = make_int4(-1, -1, -1, -1);
      const int h0 = r >= 0 ? __ldg(e + ph) : 0, h1 = r >= 0 ? __ldg(e + 7 + ph) : 0;
      for (int dh = 0; dh < mrw; ++dh) {
        const int h = h0 + dh;
        const bool rowok = h < h1;
        const int rowpix = h * W + ws;
        const uint32_t rowaddr = pbase + 16u * (uint32_t)(h * P + ws);
        for (int dw = 0; LOOP ? dw < mcw : dw < 1; dw += STEP) {
          float4 v[STEP];
#pragma unroll
          for (int j = 0; j < STEP; ++j) {
            v[j] = make_float4(-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f);
            if (rowok && ws + dw + j < we) v[j] = lds128(rowaddr + 16u * (uint32_t)(dw + j));
          }
#pragma unroll
          for (int j = 0; j < STEP; ++j) {
            if (AM) {
              const int pix = rowpix + dw + j;
              if (v[j].x > mv.x) mv.x = v[j].x, mi.x = pix;
              if (v[j].y > mv.y) mv.y = v[j].y, mi.y = pix;
              if (v[j].z > mv.z) mv.z = v[j].z, mi.z = pix;
              if (v[j].w > mv.w) mv.w = v[j].w, mi.w = pix;
            } else {  // same values: a NaN never wins either form, the -FLT_MAX sentinel never beats a pixel
              mv.x = fmaxf(mv.x, v[j].x), mv.y = fmaxf(mv.y, v[j].y);
              mv.z = fmaxf(mv.z, v[j].z), mv.w = fmaxf(mv.w, v[j].w);
            }
          }
        }
      }
      if (k < 7) {
        // empty bin: 0 / -1 (roi_pooling_kernel.cu:67-72); else flat index into the NCHW tensor
        const bool empty = !(h1 > h0 && we > ws);
        float *qv = tile_v + ph * 7 + k;
        int *qi = tile_i + ph * 7 + k;
        qv[0] = empty ? 0.f : mv.x, qv[49] = empty ? 0.f : mv.y, qv[98] = empty ? 0.f : mv.z, qv[147] = empty ? 0.f : mv.w;
        if (AM) {
          qi[0] = mi.x < 0 ? -1 : chan_base + mi.x, qi[49] = mi.y < 0 ? -1 : chan_base + HW + mi.y;
          qi[98] = mi.z < 0 ? -1 : chan_base + 2 * HW + mi.z, qi[147] = mi.w < 0 ? -1 : chan_base + 3 * HW + mi.w;
        }
      }
    }
    };
    if (mcw <= 1) scan_bins(std::integral_constant<int, 1>{}, std::false_type{});  // warp-uniform
    else if (mcw == 2) scan_bins(std::integral_constant<int, 2>{}, std::false_type{});
    else if (mcw <= 4) scan_bins(std::integral_constant<int, 4>{}, std::false_type{});
    else scan_bins(std::integral_constant<int, 4>{}, std::true_type{});
    fence_async_smem();
    __syncwarp();
    if (k == 0) {
      if (r >= 0) {
        const size_t o = ((size_t)r * C + (size_t)chunk * 4) * 49;
        bulk_s2g_nocommit(out + o, tile_v, 196 * 4);
        if (AM) bulk_s2g_nocommit(argmax + o, tile_i, 196 * 4);
      }
      bulk_commit();
    }
  }
  if (k == 0) bulk_wait_read<0>();
}


__global__ void __launch_bounds__(256)
    k_roi_pool_fwd(const float *__restrict__ feat, const float *__restrict__ rois, int B, int C,
                   int H, int W, int PH, int PW, float scale, long long total,
                   float *__restrict__ out, int *__restrict__ argmax) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % PW);
    const int ph = (int)((idx / PW) % PH);
    const int c = (int)((idx / ((long long)PW * PH)) % C);
    const int n = (int)(idx / ((long long)PW * PH * C));
    const float *roi = rois + (size_t)n * 5;
    const int b = (int)roi[0];
    const int roi_start_w = (int)roundf(__fmul_rn(roi[1], scale));
    const int roi_start_h = (int)roundf(__fmul_rn(roi[2], scale));
    const int roi_end_w = (int)roundf(__fmul_rn(roi[3], scale));
    const int roi_end_h = (int)roundf(__fmul_rn(roi[4], scale));
    const int roi_width = max(roi_end_w - roi_start_w + 1, 1);
    const int roi_height = max(roi_end_h - roi_start_h + 1, 1);
    const float bin_h = __fdiv_rn((float)roi_height, (float)PH);
    const float bin_w = __fdiv_rn((float)roi_width, (float)PW);
    int hstart = (int)floorf(__fmul_rn((float)ph, bin_h));
    int wstart = (int)floorf(__fmul_rn((float)pw, bin_w));
    int hend = (int)ceilf(__fmul_rn((float)(ph + 1), bin_h));
    int wend = (int)ceilf(__fmul_rn((float)(pw + 1), bin_w));
    hstart = min(max(hstart + roi_start_h, 0), H);
    hend = min(max(hend + roi_start_h, 0), H);
    wstart = min(max(wstart + roi_start_w, 0), W);
    wend = min(max(wend + roi_start_w, 0), W);
    const bool empty = (hend <= hstart) || (wend <= wstart) || b < 0 || b >= B;
    float maxval = empty ? 0.f : -3.402823466e+38f;
    int maxidx = -1;
    if (!empty) {
      const int base = (b * C + c) * H * W;
      for (int h = hstart; h < hend; ++h)
        for (int w = wstart; w < wend; ++w) {
          const int i = base + h * W + w;
          const float v = __ldg(feat + i);
          if (v > maxval) {
            maxval = v;
            maxidx = i;
          }
        }
    }
    out[idx] = maxval;
    if (argmax) argmax[idx] = maxidx;
  }
}

// A pooled bin sends its gradient to its argmax element iff the reference's gather at that
// element would have visited the bin (roi_pooling_kernel.cu:161-186): the element lies inside
// the ROUNDED roi (never true for an inverted roi: the reference drops that gradient) and the
// bin is inside the feasible range derived back from the element.
__global__ void __launch_bounds__(256)
    k_roi_pool_bwd(const float *__restrict__ gout, const int *__restrict__ argmax,
                   const float *__restrict__ rois, int C, int H, int W, int PH, int PW, float scale,
                   long long total, long long n_bottom, float *__restrict__ gin) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int a = __ldg(argmax + idx);
    if (a < 0 || a >= n_bottom) continue;
    const int pw = (int)(idx % PW);
    const int ph = (int)((idx / PW) % PH);
    const int n = (int)(idx / ((long long)PW * PH * C));
    const float *roi = rois + (size_t)n * 5;
    const int roi_start_w = (int)roundf(__fmul_rn(roi[1], scale));
    const int roi_start_h = (int)roundf(__fmul_rn(roi[2], scale));
    const int roi_end_w = (int)roundf(__fmul_rn(roi[3], scale));
    const int roi_end_h = (int)roundf(__fmul_rn(roi[4], scale));
    const int w = a % W, h = (a / W) % H;
    if (!(w >= roi_start_w && w <= roi_end_w && h >= roi_start_h && h <= roi_end_h)) continue;
    const int roi_width = max(roi_end_w - roi_start_w + 1, 1);
    const int roi_height = max(roi_end_h - roi_start_h + 1, 1);
    const float bin_h = __fdiv_rn((float)roi_height, (float)PH);
    const float bin_w = __fdiv_rn((float)roi_width, (float)PW);
    int phstart = (int)floorf(__fdiv_rn((float)(h - roi_start_h), bin_h));
    int phend = (int)ceilf(__fdiv_rn((float)(h - roi_start_h + 1), bin_h));
    int pwstart = (int)floorf(__fdiv_rn((float)(w - roi_start_w), bin_w));
    int pwend = (int)ceilf(__fdiv_rn((float)(w - roi_start_w + 1), bin_w));
    phstart = min(max(phstart, 0), PH);
    phend = min(max(phend, 0), PH);
    pwstart = min(max(pwstart, 0), PW);
    pwend = min(max(pwend, 0), PW);
    if (ph < phstart || ph >= phend || pw < pwstart || pw >= pwend) continue;
    atomicAdd(gin + a, __ldg(gout + idx));
  }
}

// backward, CTA = (roi, chunk of channels): everything that depends only on the roi -- the
// rounded corners, the bin sizes and, per map row / column, the range of bins the reference's
// gather would visit from there (roi_pooling_kernel.cu:161-186, same fp32 divisions) -- is
// computed once per CTA into shared memory; an element then costs a few 32-bit operations, two
// table lookups and at most one RED (the flat kernel above redoes ~250 instructions of 64-bit
// index arithmetic and fp32 divisions per element).
constexpr int kPoolBwdThreads = 256;
constexpr int kPoolBwdChunk = 64;  // channels per CTA

__global__ void __launch_bounds__(kPoolBwdThreads)
    k_roi_pool_bwd_roi(const float *__restrict__ gout, const int *__restrict__ argmax,
                       const float *__restrict__ rois, int C, int H, int W, int PH, int PW, float scale,
                       long long n_bottom, float *__restrict__ gin) {
  extern __shared__ int s_tab[];  // [H] rows then [W] columns: lo | hi << 8 | valid << 16
  const int n = blockIdx.x, c0 = blockIdx.y * kPoolBwdChunk;
  const int c1 = min(C, c0 + kPoolBwdChunk);
  const float *roi = rois + (size_t)n * 5;
  const int roi_start_w = (int)roundf(__fmul_rn(__ldg(roi + 1), scale));
  const int roi_start_h = (int)roundf(__fmul_rn(__ldg(roi + 2), scale));
  const int roi_end_w = (int)roundf(__fmul_rn(__ldg(roi + 3), scale));
  const int roi_end_h = (int)roundf(__fmul_rn(__ldg(roi + 4), scale));
  const int roi_width = max(roi_end_w - roi_start_w + 1, 1);
  const int roi_height = max(roi_end_h - roi_start_h + 1, 1);
  const float bin_h = __fdiv_rn((float)roi_height, (float)PH);
  const float bin_w = __fdiv_rn((float)roi_width, (float)PW);
  for (int i = threadIdx.x; i < H + W; i += kPoolBwdThreads) {
    const bool row = i < H;
    const int v = row ? i : i - H;
    const int s0 = row ? roi_start_h : roi_start_w, s1 = row ? roi_end_h : roi_end_w;
    const float bin = row ? bin_h : bin_w;
    const int P = row ? PH : PW;
    int lo = (int)floorf(__fdiv_rn((float)(v - s0), bin));
    int hi = (int)ceilf(__fdiv_rn((float)(v - s0 + 1), bin));
    lo = min(max(lo, 0), P), hi = min(max(hi, 0), P);
    s_tab[i] = lo | (hi << 8) | ((v >= s0 && v <= s1) ? 1 << 16 : 0);
  }
  __syncthreads();
  const int PHW = PH * PW, HW = H * W;
  const float inv_w = 1.f / (float)W;
  const size_t base = ((size_t)n * C + c0) * PHW;
  const int count = (c1 - c0) * PHW;
  for (int e = threadIdx.x; e < count; e += kPoolBwdThreads) {
    const int a = __ldcs(argmax + base + e);
    if (a < 0 || a >= n_bottom) continue;
    const int p = e % PHW, ph = p / PW, pw = p - ph * PW;
    // w = a % W, h = (a / W) % H (the reference's decomposition of the flat index, :143-146)
    const int local = a % HW;
    int h = (int)((float)local * inv_w);
    int w = local - h * W;
    if (w < 0) --h, w += W;
    if (w >= W) ++h, w -= W;
    const int th = s_tab[h], tw = s_tab[H + w];
    if (!((th >> 16) & (tw >> 16) & 1)) continue;
    if (ph < (th & 255) || ph >= ((th >> 8) & 255) || pw < (tw & 255) || pw >= ((tw >> 8) & 255)) continue;
    atomicAdd(gin + a, __ldcs(gout + base + e));
  }
}

}  // namespace rlod

using namespace rlod;

RLOD_API size_t rlod_roi_pool_workspace_bytes(int B, int R) {
  if (B < 0 || R < 0) return 0;
  return carve_align_ws(nullptr, B, R, 8, 8).bytes;
}

RLOD_API int rlod_roi_pool_forward(const float *feat, const float *rois, int B, int C, int H,
                                   int W, int R, int ph, int pw, float spatial_scale,
                                   int channels_last, float *out, int *argmax, void *workspace,
                                   size_t workspace_bytes, rlod_stream_t stream) {
  if (B < 0 || C < 0 || H < 1 || W < 1 || R < 0 || ph < 1 || pw < 1) return RLOD_EINVAL;
  if ((long long)B * C * H * W >= (1LL << 31)) return RLOD_EUNSUPPORTED;  // int argmax
  if (R == 0 || C == 0) return RLOD_OK;
  if (!feat || !rois || !out) return RLOD_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  // plane kernel: 7x7 bins, 4-channel chunks, planes + staging fit one CTA, enough rois per image
  const int P = walk_pitch(W);
  const size_t smem = (size_t)16 * H * P + (size_t)kPoolWarps * 4 * 2 * kPoolSlot * sizeof(float);
  AlignWs ws = carve_align_ws(workspace, B, R, 8, 8);
  const bool fast = ph == 7 && pw == 7 && (C % 4) == 0 && B >= 1 && smem <= (size_t)kMaxSmemPerCta &&
                    workspace && workspace_bytes >= ws.bytes && ((uintptr_t)out % 16) == 0 &&
                    (!argmax || ((uintptr_t)argmax % 16) == 0) && R >= 2 * B;
  if (channels_last && (!fast || ((uintptr_t)feat % 16) != 0)) return RLOD_EUNSUPPORTED;  // plane kernel only
  if (fast) {
    cudaMemsetAsync(ws.flag, 0, 4 * sizeof(int), st);
    RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st, k_pool_plan<<<(unsigned)cdiv(R, 128), 128, 0, st>>>(rois, R, B, H, W, spatial_scale, ws));
    const bool pdl = pdl_enabled();
    RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st, launch_after(k_roi_group_fixup, dim3(1), dim3(32), 0, st, pdl, R, B, ws));
    RLOD_LAUNCH(RLOD_KERNEL_ROI_PLAN, st,
                launch_after(k_roi_order_by_key, dim3((unsigned)B), dim3(kOrderThreads), 0, st, pdl, (const int *)ws.plan,
                             (const int *)ws.order, (const int *)ws.img_off, 31, 0, 64, ws.order2));
    const int n_chunks = C / 4;
    if (argmax) {
      cudaFuncSetAttribute(k_roi_pool7_fwd_planes<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      RLOD_LAUNCH(RLOD_KERNEL_POOL_FWD, st,
                  launch_after(k_roi_pool7_fwd_planes<true>, dim3((unsigned)(B * n_chunks)), dim3(kPoolThreads), smem, st, pdl,
                               feat, (const int *)ws.plan, (const int *)ws.order2, (const int *)ws.img_off, C, H, W, P,
                               n_chunks, channels_last, out, argmax));
    } else {
      cudaFuncSetAttribute(k_roi_pool7_fwd_planes<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      RLOD_LAUNCH(RLOD_KERNEL_POOL_FWD, st,
                  launch_after(k_roi_pool7_fwd_planes<false>, dim3((unsigned)(B * n_chunks)), dim3(kPoolThreads), smem, st, pdl,
                               feat, (const int *)ws.plan, (const int *)ws.order2, (const int *)ws.img_off, C, H, W, P,
                               n_chunks, channels_last, out, argmax));
    }
    return launch_status();
  }
  const long long total = (long long)R * C * ph * pw;
  const long long blocks = cdiv(total, 256);
  const unsigned grid = (unsigned)(blocks < (1LL << 30) ? blocks : (1LL << 30));
  RLOD_LAUNCH(RLOD_KERNEL_POOL_FWD, st, k_roi_pool_fwd<<<grid, 256, 0, st>>>(feat, rois, B, C, H, W, ph, pw,
                                                         spatial_scale, total, out, argmax));
  return launch_status();
}

RLOD_API int rlod_roi_pool_backward(const float *grad_out, const int *argmax, const float *rois,
                                    int B, int C, int H, int W, int R, int ph, int pw,
                                    float spatial_scale, int accumulate, float *grad_in,
                                    rlod_stream_t stream) {
  if (B < 0 || C < 0 || H < 1 || W < 1 || R < 0 || ph < 1 || pw < 1) return RLOD_EINVAL;
  if ((long long)B * C * H * W >= (1LL << 31)) return RLOD_EUNSUPPORTED;
  if (B == 0 || C == 0) return RLOD_OK;
  if (!grad_in) return RLOD_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_bottom = (long long)B * C * H * W;
  if (!accumulate) cudaMemsetAsync(grad_in, 0, (size_t)n_bottom * sizeof(float), st);
  if (R == 0) return launch_status();
  if (!grad_out || !argmax || !rois) return RLOD_EINVAL;
  const long long total = (long long)R * C * ph * pw;
  const long long blocks = cdiv(total, 256);
  const unsigned grid = (unsigned)(blocks < (1LL << 30) ? blocks : (1LL << 30));
  const size_t tab = (size_t)(H + W) * sizeof(int);
  if (ph <= 255 && pw <= 255 && tab <= 96 * 1024 && (long long)H * W < (1LL << 24) && (C + kPoolBwdChunk - 1) / kPoolBwdChunk <= 65535) {
    if (tab > 48 * 1024) cudaFuncSetAttribute(k_roi_pool_bwd_roi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab);
    RLOD_LAUNCH(RLOD_KERNEL_POOL_BWD, st,
                k_roi_pool_bwd_roi<<<dim3((unsigned)R, (unsigned)((C + kPoolBwdChunk - 1) / kPoolBwdChunk)), kPoolBwdThreads,
                                     tab, st>>>(grad_out, argmax, rois, C, H, W, ph, pw, spatial_scale, n_bottom, grad_in));
    return launch_status();
  }
  RLOD_LAUNCH(RLOD_KERNEL_POOL_BWD, st,
              k_roi_pool_bwd<<<grid, 256, 0, st>>>(grad_out, argmax, rois, C, H, W, ph, pw,
                                                   spatial_scale, total, n_bottom, grad_in));
  return launch_status();
}
