// Device-side pieces shared by nms.cu and proposal.cu.
#pragma once
#include "rlod_common.cuh"

namespace rlod {

// A set of independent NMS problems over one dets array.
struct NmsSegs {
  const float *dets;       // rows of `stride` floats, [x1,y1,x2,y2,...]
  int stride;
  int vec4;                // stride == 4 and dets 16-byte aligned: float4 loads
  const int *seg_offsets;  // nseg+1 (device) or NULL: uniform segments of uniform_n rows
  int uniform_n;
  int max_n;               // host-declared upper bound; longer segments are truncated to it

  __device__ __forceinline__ void get(int seg, int &off, int &n) const {
    if (seg_offsets) {
      off = seg_offsets[seg];
      n = seg_offsets[seg + 1] - off;
    } else {
      off = seg * uniform_n;
      n = uniform_n;
    }
    n = max(0, min(n, max_n));
  }
  __device__ __forceinline__ float4 load_box(int row) const {
    const float *p = dets + (size_t)row * stride;
    if (vec4) return __ldg(reinterpret_cast<const float4 *>(p));
    return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
  }
};

// Where the scan writes its result.
struct NmsOut {
  int *keep;     // one int per dets row (segment-local kept indices, -1 padded) or NULL
  int *num_out;  // [nseg] or NULL
  float *rois;   // proposal mode: (nseg, post, 5) [seg, x1, y1, x2, y2], zero padded, or NULL
  int post;

  __device__ __forceinline__ void emit(int seg, int off, int rank, int idx,
                                       const NmsSegs &segs) const {
    if (keep) keep[off + rank] = idx;
    if (rois && rank < post) {
      const float4 b = segs.load_box(off + idx);
      float *o = rois + ((size_t)seg * post + rank) * 5;
      o[0] = (float)seg;
      o[1] = b.x, o[2] = b.y, o[3] = b.z, o[4] = b.w;
    }
  }
  __device__ __forceinline__ void finish(int seg, int off, int n, int total, int t,
                                         int nthreads) const {
    if (keep)
      for (int i = total + t; i < n; i += nthreads) keep[off + i] = -1;
    if (num_out && t == 0) num_out[seg] = total;
    if (rois)
      for (int i = total + t; i < post; i += nthreads) {
        float *o = rois + ((size_t)seg * post + i) * 5;
        o[0] = (float)seg;
        o[1] = o[2] = o[3] = o[4] = 0.f;
      }
  }
};

// idx-th tile of the row-major upper triangle of an nb x nb grid -> (rb, cb), cb >= rb
__device__ __forceinline__ void tri_decode(int idx, int nb, int &rb, int &cb) {
  const float s = (float)(2 * nb + 1);
  int r = (int)((s - sqrtf(fmaxf(s * s - 8.f * (float)idx, 0.f))) * 0.5f);
  r = max(0, min(r, nb - 1));
  while (r > 0 && r * nb - (r * (r - 1)) / 2 > idx) --r;
  while ((r + 1) * nb - ((r + 1) * r) / 2 <= idx) ++r;
  rb = r;
  cb = r + (idx - (r * nb - (r * (r - 1)) / 2));
}

// IoU(a, b) > thresh with the reference's exact fp32 expression (nms_cuda_kernel.cu:31-39 as
// compiled: S = fma(w_b, h_b, Sa)).  a = the earlier (row) box, b = the later (column) box.
// zero_false: thresh >= 0, so an empty intersection can never exceed it (0/u is 0, -0 or
// NaN).  The interval test skips the IEEE division whenever inter is at least 2e-6
// (relative) away from thresh*u -- sixteen times the worst rounding error of the quotient,
// so the decision equals the divided one; near-ties take the exact division.
__device__ __forceinline__ bool iou_gt(const float4 &a, float Sa, const float4 &b,
                                       const float2 &bwh, float thresh, bool zero_false) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float w = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
  // most pairs do not even overlap in x: w == 0 makes inter == +-0 or NaN (0 * inf), never > thresh >= 0
  if (zero_false && w == 0.f) return false;
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float h = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
  const float inter = __fmul_rn(w, h);
  if (zero_false && inter == 0.f) return false;
  const float S = __fmaf_rn(bwh.x, bwh.y, Sa);
  const float u = __fsub_rn(S, inter);
  if (zero_false) {
    const float tq = __fmul_rn(thresh, u);
    if (tq > 1e-30f && tq < 1e30f && inter < 1e30f) {
      if (inter > __fmul_rn(tq, 1.000002f)) return true;
      if (inter < __fmul_rn(tq, 0.999998f)) return false;
    }
  }
  return __fdiv_rn(inter, u) > thresh;
}

size_t nms_mask_bytes(int nseg, int max_seg, int max_keep);
int nms_launch(const NmsSegs &segs, int nseg, int max_seg, float thresh, int max_keep,
               const NmsOut &out, void *workspace, size_t workspace_bytes, cudaStream_t st,
               int force_large);

}  // namespace rlod
