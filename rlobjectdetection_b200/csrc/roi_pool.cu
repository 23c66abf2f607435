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
#include "rlod_common.cuh"

namespace rlod {

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

}  // namespace rlod

using namespace rlod;

RLOD_API int rlod_roi_pool_forward(const float *feat, const float *rois, int B, int C, int H,
                                   int W, int R, int ph, int pw, float spatial_scale, float *out,
                                   int *argmax, rlod_stream_t stream) {
  if (B < 0 || C < 0 || H < 1 || W < 1 || R < 0 || ph < 1 || pw < 1) return RLOD_EINVAL;
  if ((long long)B * C * H * W >= (1LL << 31)) return RLOD_EUNSUPPORTED;  // int argmax
  if (R == 0 || C == 0) return RLOD_OK;
  if (!feat || !rois || !out) return RLOD_EINVAL;
  const long long total = (long long)R * C * ph * pw;
  const long long blocks = cdiv(total, 256);
  const unsigned grid = (unsigned)(blocks < (1LL << 30) ? blocks : (1LL << 30));
  RLOD_LAUNCH(RLOD_KERNEL_POOL_FWD, (cudaStream_t)stream, k_roi_pool_fwd<<<grid, 256, 0, (cudaStream_t)stream>>>(feat, rois, B, C, H, W, ph, pw,
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
  RLOD_LAUNCH(RLOD_KERNEL_POOL_BWD, st,
              k_roi_pool_bwd<<<grid, 256, 0, st>>>(grad_out, argmax, rois, C, H, W, ph, pw,
                                                   spatial_scale, total, n_bottom, grad_in));
  return launch_status();
}
