// Test-time post-processing of the detector head, the caller of per-class NMS
// (RCNN_bases/test_net.py:244-307, same loop in demo.py:305-334), as TWO launches for the whole
// batch instead of 80 Python iterations + 80 device->host copies per image:
//
//   k_detect_classes   one CTA per (class >= 1, image):
//       scores[:, j] > thresh (:278), class box = bbox_transform_inv(roi, delta * std + mean)
//       (:251-262), clip_boxes (:263), / im_scale (:268), stable descending sort by score
//       (:282, ties: lower roi index first), nms(cls_dets, TEST.NMS) (:291) -> the kept
//       [x1, y1, x2, y2, score] rows in order + their count.
//   k_detect_cap       one CTA per image: max_per_image over all classes (:299-307):
//       thresh = the max_per_image-th largest kept score, keep score >= thresh.  Every class
//       list is sorted by score, so the survivors are a prefix: only the counts shrink.
//
// Every fp32 operation rounds separately, in the order of the eager torch expressions it
// replaces; NMS arithmetic is nms_device.cuh's (bit-exact with the reference kernel).
#include "nms_device.cuh"

namespace rlod {

constexpr int kDetThreads = 256;
constexpr int kDetMaxN = 512;

__device__ __forceinline__ uint32_t det_desc_key(float f) {
  uint32_t u = __float_as_uint(f);
  if (f == 0.f) u = 0u;
  const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~asc;
}
__device__ __forceinline__ float det_key_score(uint32_t key) {
  const uint32_t asc = ~key;
  const uint32_t u = (asc & 0x80000000u) ? (asc & 0x7fffffffu) : ~asc;
  return __uint_as_float(u);
}

struct DetArgs {
  const float *rois;       // (B, N, 5) [img, x1, y1, x2, y2], image coordinates of the scaled input
  const float *cls_prob;   // (B, N, K)
  const float *bbox_pred;  // (B, N, 4K) or (B, N, 4) when class_agnostic; NULL: no regression
  const float *im_info;    // (B, 3) [h, w, scale]
  int B, N, K, class_agnostic, normalize;
  float std[4], mean[4];
  float score_thresh, nms_thresh;
  float *dets;   // (B, K, N, 5)
  int *counts;   // (B, K)
};

__global__ void __launch_bounds__(kDetThreads) k_detect_classes(DetArgs a, int mp) {
  extern __shared__ __align__(16) unsigned char det_raw[];
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(det_raw);  // [mp]
  float4 *box = reinterpret_cast<float4 *>(keys + mp);                          // [mp]
  float2 *wh = reinterpret_cast<float2 *>(box + mp);                            // [mp]
  unsigned long long *mk = reinterpret_cast<unsigned long long *>(wh + mp);     // [nblk][mp + 1]
  __shared__ int s_m, s_total;
  const int j = blockIdx.x + 1, b = blockIdx.y, t = threadIdx.x;
  const int N = a.N, K = a.K;
  if (t == 0) s_m = 0, s_total = 0;
  __syncthreads();
  // candidates of this class: composite key (desc(score) << 32 | roi index), others sort last
  int cnt = 0;
  for (int i = t; i < mp; i += kDetThreads) {
    unsigned long long key = ~0ull;
    if (i < N) {
      const float s = __ldg(a.cls_prob + ((size_t)b * N + i) * K + j);
      if (s > a.score_thresh) {
        key = ((unsigned long long)det_desc_key(s) << 32) | (unsigned)i;
        ++cnt;
      }
    }
    keys[i] = key;
  }
  if (cnt) atomicAdd(&s_m, cnt);
  __syncthreads();
  const int m = s_m;
  float *dout = a.dets + ((size_t)b * K + j) * (size_t)N * 5;
  if (m == 0) {
    if (t == 0) a.counts[b * K + j] = 0;
    return;
  }
  // bitonic sort, ascending composite key = descending score, ties by lower roi index
  for (int k = 2; k <= mp; k <<= 1)
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int p = t; p < (mp >> 1); p += kDetThreads) {
        const int i = ((p & ~(jj - 1)) << 1) | (p & (jj - 1));
        const int l = i + jj;
        const unsigned long long x = keys[i], y = keys[l];
        if ((x > y) == ((i & k) == 0)) keys[i] = y, keys[l] = x;
      }
      __syncthreads();
    }
  // class boxes of the sorted candidates
  const float imh = a.im_info[b * 3 + 0], imw = a.im_info[b * 3 + 1], ims = a.im_info[b * 3 + 2];
  const float xmax = __fsub_rn(imw, 1.f), ymax = __fsub_rn(imh, 1.f);
  const int nblk = (m + 63) >> 6, npad = nblk << 6;
  for (int i = t; i < npad; i += kDetThreads) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < m) {
      const int n = (int)(unsigned)(keys[i] & 0xffffffffull);
      const float *r = a.rois + ((size_t)b * N + n) * 5;
      const float x1 = __ldg(r + 1), y1 = __ldg(r + 2), x2 = __ldg(r + 3), y2 = __ldg(r + 4);
      if (a.bbox_pred) {
        const int dk = a.class_agnostic ? 1 : K;
        const float *d = a.bbox_pred + ((size_t)b * N + n) * 4 * dk + (a.class_agnostic ? 0 : 4 * j);
        float d0 = __ldg(d), d1 = __ldg(d + 1), d2 = __ldg(d + 2), d3 = __ldg(d + 3);
        if (a.normalize) {
          d0 = __fadd_rn(__fmul_rn(d0, a.std[0]), a.mean[0]);
          d1 = __fadd_rn(__fmul_rn(d1, a.std[1]), a.mean[1]);
          d2 = __fadd_rn(__fmul_rn(d2, a.std[2]), a.mean[2]);
          d3 = __fadd_rn(__fmul_rn(d3, a.std[3]), a.mean[3]);
        }
        const float w = __fadd_rn(__fsub_rn(x2, x1), 1.0f), h = __fadd_rn(__fsub_rn(y2, y1), 1.0f);
        const float cx = __fadd_rn(x1, __fmul_rn(0.5f, w)), cy = __fadd_rn(y1, __fmul_rn(0.5f, h));
        const float pcx = __fadd_rn(__fmul_rn(d0, w), cx), pcy = __fadd_rn(__fmul_rn(d1, h), cy);
        const float pw = __fmul_rn(expf(d2), w), ph = __fmul_rn(expf(d3), h);
        o.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
        o.y = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
        o.z = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
        o.w = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
        o.x = (o.x != o.x) ? o.x : fminf(fmaxf(o.x, 0.f), xmax);
        o.y = (o.y != o.y) ? o.y : fminf(fmaxf(o.y, 0.f), ymax);
        o.z = (o.z != o.z) ? o.z : fminf(fmaxf(o.z, 0.f), xmax);
        o.w = (o.w != o.w) ? o.w : fminf(fmaxf(o.w, 0.f), ymax);
      } else {
        o = make_float4(x1, y1, x2, y2);  // np.tile(boxes, ...) (:266)
      }
      o.x = __fdiv_rn(o.x, ims), o.y = __fdiv_rn(o.y, ims), o.z = __fdiv_rn(o.z, ims), o.w = __fdiv_rn(o.w, ims);
    }
    box[i] = o;
    wh[i] = make_float2(__fadd_rn(__fsub_rn(o.z, o.x), 1.f), __fadd_rn(__fsub_rn(o.w, o.y), 1.f));
  }
  __syncthreads();
  // bitmask NMS in shared memory (same structure as k_nms_small)
  const int mstride = npad + 1;
  const bool fast = a.nms_thresh >= 0.f && a.nms_thresh < 1e30f;
  for (int it = t; it < nblk * npad; it += kDetThreads) {
    const int cb = it / npad, i = it - cb * npad;
    const int rb = i >> 6;
    if (cb < rb) continue;
    unsigned long long bits = 0ull;
    if (i < m) {
      const float4 bi = box[i];
      const float Sa = __fmul_rn(wh[i].x, wh[i].y);
      const int jn = min(64, m - cb * 64);
      const int j0 = (rb == cb) ? (i & 63) + 1 : 0;
      for (int q = j0; q < jn; ++q)
        if (iou_gt(bi, Sa, box[cb * 64 + q], wh[cb * 64 + q], a.nms_thresh, fast)) bits |= 1ull << q;
    }
    mk[(size_t)cb * mstride + i] = bits;
  }
  __syncthreads();
  if (t < 32) {
    const int lane = t;
    unsigned long long rem = 0ull;  // lane w owns the removed-bitmap word of column block w
    int total = 0;
    for (int k = 0; k < nblk; ++k) {
      unsigned long long r = __shfl_sync(0xffffffffu, rem, k);
      const int valid = min(64, m - k * 64);
      if (valid < 64) r |= ~0ull << valid;
      unsigned long long kb = 0ull;
      const unsigned long long *dg = mk + (size_t)k * mstride + k * 64;
      for (int i = 0; i < valid; ++i)
        if (!((r >> i) & 1ull)) {
          kb |= 1ull << i;
          r |= dg[i];
        }
      for (int q = lane; q < 64; q += 32)
        if ((kb >> q) & 1ull) {
          const int rank = total + __popcll(kb & ((1ull << q) - 1ull));
          const int i = k * 64 + q;
          const float4 bx = box[i];
          float *o = dout + (size_t)rank * 5;
          o[0] = bx.x, o[1] = bx.y, o[2] = bx.z, o[3] = bx.w;
          o[4] = det_key_score((uint32_t)(keys[i] >> 32));
        }
      total += __popcll(kb);
      if (lane > k && lane < nblk) {
        const unsigned long long *row = mk + (size_t)lane * mstride + k * 64;
        unsigned long long kk = kb;
        while (kk) {
          const int i = __ffsll((long long)kk) - 1;
          kk &= kk - 1ull;
          rem |= row[i];
        }
      }
    }
    if (lane == 0) a.counts[b * K + j] = total;
  }
}

// max_per_image: one CTA per image.  The max_per_image-th largest kept score by bisection on
// the descending-orderable key; each class keeps its prefix with score >= that threshold.
__global__ void __launch_bounds__(kDetThreads)
    k_detect_cap(const float *__restrict__ dets, int *__restrict__ counts, int N, int K, int max_per_image) {
  __shared__ unsigned s_cnt[3];
  __shared__ int s_tot;
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31;
  int *cnt = counts + (size_t)b * K;
  const float *d = dets + (size_t)b * K * (size_t)N * 5;
  if (t < 3) s_cnt[t] = 0u;
  if (t == 0) {
    int tot = 0;
    for (int j = 1; j < K; ++j) tot += cnt[j];
    s_tot = tot;
    cnt[0] = 0;  // background is never reported (:277)
  }
  __syncthreads();
  if (s_tot <= max_per_image) return;
  int pass = 0;
  auto count_le = [&](uint32_t mid) -> unsigned {
    unsigned c = 0;
    for (int e = t; e < (K - 1) * N; e += kDetThreads) {
      const int j = 1 + e / N, i = e - (j - 1) * N;
      if (i < cnt[j]) c += det_desc_key(d[((size_t)j * N + i) * 5 + 4]) <= mid;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    const int slot = pass % 3;
    if (lane == 0 && c) atomicAdd(&s_cnt[slot], c);
    if (t == 0) s_cnt[(pass + 1) % 3] = 0u;
    __syncthreads();
    ++pass;
    return s_cnt[slot];
  };
  uint32_t lo = 0u, hi = 0xffffffffu;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (count_le(mid) >= (unsigned)max_per_image) hi = mid;
    else lo = mid + 1u;
  }
  __syncthreads();
  // np.where(scores >= image_thresh): key <= lo; lists are sorted, so this is a prefix
  for (int j = 1 + t; j < K; j += kDetThreads) {
    const int c = cnt[j];
    int keep = 0;
    while (keep < c && det_desc_key(d[((size_t)j * N + keep) * 5 + 4]) <= lo) ++keep;
    cnt[j] = keep;
  }
}

}  // namespace rlod

using namespace rlod;

RLOD_API int rlod_detect_postprocess(const float *rois, const float *cls_prob, const float *bbox_pred,
                                     const float *im_info, int B, int N, int K, int class_agnostic,
                                     const float *stds, const float *means, float score_thresh,
                                     float nms_thresh, int max_per_image, float *dets, int *counts,
                                     rlod_stream_t stream) {
  if (B < 0 || N < 0 || K < 1) return RLOD_EINVAL;
  if (B == 0 || K == 1) return RLOD_OK;
  if (!dets || !counts) return RLOD_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    cudaMemsetAsync(counts, 0, (size_t)B * K * sizeof(int), st);
    return launch_status();
  }
  if (!rois || !cls_prob || !im_info) return RLOD_EINVAL;
  if (N > kDetMaxN) return RLOD_EUNSUPPORTED;
  if (B > 65535) return RLOD_EUNSUPPORTED;
  DetArgs a;
  a.rois = rois, a.cls_prob = cls_prob, a.bbox_pred = bbox_pred, a.im_info = im_info;
  a.B = B, a.N = N, a.K = K, a.class_agnostic = class_agnostic;
  a.normalize = (stds && means) ? 1 : 0;
  for (int i = 0; i < 4; ++i) a.std[i] = stds ? stds[i] : 1.f, a.mean[i] = means ? means[i] : 0.f;
  a.score_thresh = score_thresh, a.nms_thresh = nms_thresh;
  a.dets = dets, a.counts = counts;
  int mp = 64;
  while (mp < N) mp <<= 1;
  const int nblk = mp / 64;
  const size_t smem = (size_t)mp * (sizeof(unsigned long long) + sizeof(float4) + sizeof(float2)) +
                      (size_t)nblk * (mp + 1) * sizeof(unsigned long long);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(k_detect_classes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  RLOD_LAUNCH(RLOD_KERNEL_DETECT, st, k_detect_classes<<<dim3(K - 1, B), kDetThreads, smem, st>>>(a, mp));
  // class 0 counts + the cap (the cap kernel also zeroes counts[:, 0])
  RLOD_LAUNCH(RLOD_KERNEL_DETECT, st,
              k_detect_cap<<<B, kDetThreads, 0, st>>>(dets, counts, N, K, max_per_image > 0 ? max_per_image : 0x7fffffff));
  return launch_status();
}
