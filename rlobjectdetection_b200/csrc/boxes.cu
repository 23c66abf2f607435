// Box algebra of lib/model/rpn/bbox_transform.py as single launches for sm_100a.
// Every fp32 operation rounds separately (__f*_rn intrinsics, no FMA contraction) because
// the eager torch expressions these replace round after every elementwise op; results are
// therefore bit-identical except for exp(), where libdevice expf and torch's CPU
// vectorised exp may differ in the last ulp.
#include "rlod_common.cuh"

namespace rlod {

// bbox_transform_inv, bbox_transform.py:77-103.  One thread per (box, class) 4-tuple.
__global__ void __launch_bounds__(256)
    k_bbox_transform_inv(const float *__restrict__ boxes, const float *__restrict__ deltas,
                         long long BN, int k, float *__restrict__ out) {
  const long long total = BN * k;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / k;
    const float *bx = boxes + i * 4;
    const float *d = deltas + idx * 4;
    const float x1 = __ldg(bx), y1 = __ldg(bx + 1), x2 = __ldg(bx + 2), y2 = __ldg(bx + 3);
    const float w = __fadd_rn(__fsub_rn(x2, x1), 1.0f), h = __fadd_rn(__fsub_rn(y2, y1), 1.0f);
    const float cx = __fadd_rn(x1, __fmul_rn(0.5f, w)), cy = __fadd_rn(y1, __fmul_rn(0.5f, h));
    const float pcx = __fadd_rn(__fmul_rn(__ldg(d), w), cx);
    const float pcy = __fadd_rn(__fmul_rn(__ldg(d + 1), h), cy);
    const float pw = __fmul_rn(expf(__ldg(d + 2)), w), ph = __fmul_rn(expf(__ldg(d + 3)), h);
    float *o = out + idx * 4;
    o[0] = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
    o[1] = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
    o[2] = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
    o[3] = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
  }
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) {
  return (v != v) ? v : fminf(fmaxf(v, lo), hi);  // torch.clamp_: NaN propagates
}

// clip_boxes, bbox_transform.py:125-133 (4*B clamp_ launches + host reads of im_shape there)
__global__ void __launch_bounds__(256)
    k_clip_boxes(float *__restrict__ boxes, const float *__restrict__ im_info, int B, long long Nk) {
  const long long total = (long long)B * Nk;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / Nk);
    const float xmax = __fsub_rn(__ldg(im_info + b * 3 + 1), 1.f);
    const float ymax = __fsub_rn(__ldg(im_info + b * 3 + 0), 1.f);
    float *p = boxes + idx * 4;
    p[0] = clampf(p[0], 0.f, xmax);
    p[1] = clampf(p[1], 0.f, ymax);
    p[2] = clampf(p[2], 0.f, xmax);
    p[3] = clampf(p[3], 0.f, ymax);
  }
}

// IoU with the +1 convention, bbox_transform.py:136-166 / :168-257
__device__ __forceinline__ float overlap_rcnn(float a0, float a1, float a2, float a3, float aa,
                                              float g0, float g1, float g2, float g3, float ga) {
  float iw = __fadd_rn(__fsub_rn(fminf(a2, g2), fmaxf(a0, g0)), 1.f);
  if (iw < 0.f) iw = 0.f;
  float ih = __fadd_rn(__fsub_rn(fminf(a3, g3), fmaxf(a1, g1)), 1.f);
  if (ih < 0.f) ih = 0.f;
  const float inter = __fmul_rn(iw, ih);
  const float ua = __fsub_rn(__fadd_rn(aa, ga), inter);
  return __fdiv_rn(inter, ua);
}

// anchors (B,N,*) or shared (N,*); gt (B,K,*); out (B,N,K).  sentinels != 0: the degenerate
// box rules of bbox_overlaps_batch (:195-196, 212-213).
__global__ void __launch_bounds__(256)
    k_bbox_overlaps(const float *__restrict__ anchors, long long a_bstride, int a_rstride,
                    const float *__restrict__ gt, int g_rstride, int B, int N, int K, int sentinels,
                    float *__restrict__ out) {
  const long long total = (long long)B * N * K;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % K);
    const int n = (int)((idx / K) % N);
    const int b = (int)(idx / ((long long)K * N));
    const float *a = anchors + (size_t)b * a_bstride + (size_t)n * a_rstride;
    const float *g = gt + ((size_t)b * K + k) * g_rstride;
    const float a0 = __ldg(a), a1 = __ldg(a + 1), a2 = __ldg(a + 2), a3 = __ldg(a + 3);
    const float g0 = __ldg(g), g1 = __ldg(g + 1), g2 = __ldg(g + 2), g3 = __ldg(g + 3);
    const float ax = __fadd_rn(__fsub_rn(a2, a0), 1.f), ay = __fadd_rn(__fsub_rn(a3, a1), 1.f);
    const float gx = __fadd_rn(__fsub_rn(g2, g0), 1.f), gy = __fadd_rn(__fsub_rn(g3, g1), 1.f);
    float v = overlap_rcnn(a0, a1, a2, a3, __fmul_rn(ax, ay), g0, g1, g2, g3, __fmul_rn(gx, gy));
    if (sentinels) {
      if (gx == 1.f && gy == 1.f) v = 0.f;
      if (ax == 1.f && ay == 1.f) v = -1.f;
    }
    out[idx] = v;
  }
}

static unsigned grid_for(long long total) {
  const long long blocks = cdiv(total, 256);
  return (unsigned)(blocks < (1LL << 30) ? (blocks > 0 ? blocks : 1) : (1LL << 30));
}

}  // namespace rlod

using namespace rlod;

RLOD_API int rlod_bbox_transform_inv(const float *boxes, const float *deltas, long long BN, int k,
                                     float *out, rlod_stream_t stream) {
  if (BN < 0 || k < 1) return RLOD_EINVAL;
  if (BN == 0) return RLOD_OK;
  if (!boxes || !deltas || !out) return RLOD_EINVAL;
  RLOD_LAUNCH(RLOD_KERNEL_BOXES, (cudaStream_t)stream, k_bbox_transform_inv<<<grid_for(BN * k), 256, 0, (cudaStream_t)stream>>>(boxes, deltas, BN, k, out));
  return launch_status();
}

RLOD_API int rlod_clip_boxes(float *boxes, const float *im_info, int B, long long N, int k,
                             rlod_stream_t stream) {
  if (B < 0 || N < 0 || k < 1) return RLOD_EINVAL;
  if (B == 0 || N == 0) return RLOD_OK;
  if (!boxes || !im_info) return RLOD_EINVAL;
  RLOD_LAUNCH(RLOD_KERNEL_BOXES, (cudaStream_t)stream, k_clip_boxes<<<grid_for((long long)B * N * k), 256, 0, (cudaStream_t)stream>>>(boxes, im_info, B,
                                                                                N * k));
  return launch_status();
}

RLOD_API int rlod_bbox_overlaps(const float *anchors, const float *gt, int N, int K, float *out,
                                rlod_stream_t stream) {
  if (N < 0 || K < 0) return RLOD_EINVAL;
  if (N == 0 || K == 0) return RLOD_OK;
  if (!anchors || !gt || !out) return RLOD_EINVAL;
  RLOD_LAUNCH(RLOD_KERNEL_BOXES, (cudaStream_t)stream, k_bbox_overlaps<<<grid_for((long long)N * K), 256, 0, (cudaStream_t)stream>>>(
      anchors, 0, 4, gt, 4, 1, N, K, 0, out));
  return launch_status();
}

RLOD_API int rlod_bbox_overlaps_batch(const float *anchors, long long anchor_batch_stride,
                                      int anchor_row_stride, const float *gt, int gt_row_stride,
                                      int B, int N, int K, float *out, rlod_stream_t stream) {
  if (B < 0 || N < 0 || K < 0 || anchor_row_stride < 4 || gt_row_stride < 4 ||
      anchor_batch_stride < 0)
    return RLOD_EINVAL;
  if (B == 0 || N == 0 || K == 0) return RLOD_OK;
  if (!anchors || !gt || !out) return RLOD_EINVAL;
  RLOD_LAUNCH(RLOD_KERNEL_BOXES, (cudaStream_t)stream, k_bbox_overlaps<<<grid_for((long long)B * N * K), 256, 0, (cudaStream_t)stream>>>(
      anchors, anchor_batch_stride, anchor_row_stride, gt, gt_row_stride, B, N, K, 1, out));
  return launch_status();
}
