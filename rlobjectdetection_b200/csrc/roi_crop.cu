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
// Every fp32 operation rounds separately, in the reference's order.
#include "rlod_common.cuh"

namespace rlod {

__device__ __forceinline__ void crop_top_left(float x, int width, int &point, float &weight) {
  const float xcoord = __fdiv_rn(__fmul_rn(__fadd_rn(x, 1.f), (float)(width - 1)), 2.f);
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

static unsigned crop_grid(long long total) {
  const long long blocks = (total + 255) / 256;
  return (unsigned)(blocks < (1LL << 22) ? blocks : (1LL << 22));
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
  RLOD_LAUNCH(RLOD_KERNEL_CROP, st,
              k_roi_crop<true><<<dim3((unsigned)R, (unsigned)((C + kCropChunk - 1) / kCropChunk)), kCropThreads, 0, st>>>(
                  grad_out, grid_yx, B, C, H, W, R, gh, gw, R / B, grad_feat));
  return launch_status();
}
