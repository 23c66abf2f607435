/*
 * rlod_oracle.c -- CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity oracle for the sm_100a kernels in rlobjectdetection_b200/csrc.
 * It is NOT product code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product path never falls back to it.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference checkout, jbr97/RLObjectDetection).  Float arithmetic mirrors what the
 * reference's legacy CUDA kernels do when compiled UNCHANGED with nvcc 12.9 for sm_100a
 * (inspected with cuobjdump -sass): where nvcc contracts a*b+c into one FFMA the oracle
 * calls fmaf() explicitly; everywhere else the translation unit is built with
 * -ffp-contract=off so gcc never fuses on its own.
 *
 * Pinning: see oracle/README.md -- checked against (1) the reference's Python run in the
 * authoring container (tests/golden/make_golden.py), (2) the reference's legacy CUDA
 * kernels compiled unchanged into oracle/_ref/libref_legacy.so (GPU tests), (3) the
 * reference's vendored maskApi.c compiled into oracle/_ref/libmaskapi.so.
 *
 * Build: make -C oracle        (gcc -O2 -ffp-contract=off -fopenmp -shared)
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORC_API void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------
 * NMS.  lib/model/nms/src/nms_cuda_kernel.cu:31-39 (devIoU), :77-81 (strict '>' bit),
 * :123-144 (greedy host scan).  a = the earlier (row) box, b = the later (column) box.
 * SASS of the unchanged kernel (nvcc 12.9, sm_100a):
 *   Sa    = FMUL(a2-a0+1, a3-a1+1)
 *   S     = FFMA(b2-b0+1, b3-b1+1, Sa)          <- "Sa + Sb" is contracted
 *   inter = FMUL(max(r-l+1,0), max(b-t+1,0))
 *   iou   = inter / (S - inter)                  (IEEE division)
 * ---------------------------------------------------------------------------------- */
static inline float orc_dev_iou(const float *a, const float *b) {
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float width = fmaxf(right - left + 1.f, 0.f), height = fmaxf(bottom - top + 1.f, 0.f);
  float interS = width * height;
  float Sa = (a[2] - a[0] + 1.f) * (a[3] - a[1] + 1.f);
  float S = fmaf(b[2] - b[0] + 1.f, b[3] - b[1] + 1.f, Sa);
  return interS / (S - interS);
}

ORC_API float orc_nms_iou(const float *a, const float *b) { return orc_dev_iou(a, b); }

/* dets: n rows of `stride` floats, [x1,y1,x2,y2,...]; must be sorted by score descending
 * (the score column is never read, nms_cuda_kernel.cu:55-64).  keep gets ascending indices.
 * max_keep <= 0: unlimited (bare nms()); > 0: stop after that many keeps (the proposal
 * layer only uses keep[:post_nms_topN], proposal_layer.py:151-152).  Returns num_out. */
ORC_API int orc_nms(const float *dets, int n, int stride, float thresh, int max_keep, int *keep) {
  if (n <= 0) return 0;
  unsigned char *removed = (unsigned char *)calloc((size_t)n, 1);
  int num = 0;
  for (int i = 0; i < n; ++i) {
    if (removed[i]) continue;
    keep[num++] = i;
    if (max_keep > 0 && num >= max_keep) break;
    const float *a = dets + (size_t)i * stride;
    for (int j = i + 1; j < n; ++j) {
      if (removed[j]) continue;
      if (orc_dev_iou(a, dets + (size_t)j * stride) > thresh) removed[j] = 1;
    }
  }
  free(removed);
  return num;
}

/* segmented form: seg_offsets[nseg+1]; keep is written at the segment's own offset,
 * indices are segment-local; num_out[nseg]. (test_net.py:277-297 per-class loop) */
ORC_API void orc_nms_batched(const float *dets, int stride, const int *seg_offsets, int nseg,
                             float thresh, int max_keep, int *keep, int *num_out) {
#pragma omp parallel for schedule(dynamic, 8)
  for (int s = 0; s < nseg; ++s) {
    int o = seg_offsets[s], n = seg_offsets[s + 1] - o;
    num_out[s] = orc_nms(dets + (size_t)o * stride, n, stride, thresh, max_keep, keep + o);
  }
}

/* ------------------------------------------------------------------------------------
 * RoIAlign.  lib/model/roi_align/src/roi_align_kernel.cu:15-70 (fwd), :94-143 (bwd).
 * Geometry as compiled (SASS): start = FMUL(coord,scale); size = max(FFMA(end_coord,
 * scale,-start)+1, 0); bin = (float)((double)size/(double)(A-1)); h = FFMA(ph,bin,start).
 * Interpolation weights: mixed fp64/fp32 exactly as the C expression at :64-67 reads.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int batch, valid_any;
  float start_w, start_h, bin_w, bin_h;
} orc_roi_geom;

static inline orc_roi_geom orc_align_geom(const float *roi, float scale, int ah, int aw) {
  orc_roi_geom g;
  g.batch = 0;
  g.start_w = roi[1] * scale;
  g.start_h = roi[2] * scale;
  float roi_w = fmaxf(fmaf(roi[3], scale, -g.start_w) + 1.f, 0.f);
  float roi_h = fmaxf(fmaf(roi[4], scale, -g.start_h) + 1.f, 0.f);
  g.bin_h = (float)((double)roi_h / ((double)ah - 1.));
  g.bin_w = (float)((double)roi_w / ((double)aw - 1.));
  g.valid_any = 1;
  return g;
}

/* img_start = roi_batch_ind(float) * channels * height * width evaluated in fp32 then
 * truncated (roi_align_kernel.cu:51) */
static inline long orc_align_img_start(float batch_ind, int C, int H, int W) {
  float f = batch_ind * (float)C;
  f = f * (float)H;
  f = f * (float)W;
  return (long)(int)f;
}

ORC_API void orc_roi_align_fwd(const float *feat, const float *rois, int B, int C, int H, int W,
                               int R, int ah, int aw, float scale, float *out) {
  (void)B;
#pragma omp parallel for schedule(dynamic, 4)
  for (int n = 0; n < R; ++n) {
    const float *roi = rois + (size_t)n * 5;
    orc_roi_geom g = orc_align_geom(roi, scale, ah, aw);
    long img_start = orc_align_img_start(roi[0], C, H, W);
    for (int c = 0; c < C; ++c) {
      const float *plane = feat + img_start + (long)c * H * W;
      float *o = out + ((size_t)n * C + c) * ah * aw;
      for (int ph = 0; ph < ah; ++ph) {
        float h = fmaf((float)ph, g.bin_h, g.start_h);
        int hstart = (int)fminf(floorf(h), (float)(H - 2));
        for (int pw = 0; pw < aw; ++pw) {
          float w = fmaf((float)pw, g.bin_w, g.start_w);
          int wstart = (int)fminf(floorf(w), (float)(W - 2));
          if (h < 0 || h >= H || w < 0 || w >= W) {
            o[ph * aw + pw] = 0.f;
          } else {
            float h_ratio = h - (float)hstart, w_ratio = w - (float)wstart;
            const float *p = plane + (long)hstart * W + wstart;
            double v = (double)p[0] * (1. - h_ratio) * (1. - w_ratio) +
                       (double)p[1] * (1. - h_ratio) * w_ratio +
                       (double)(p[W] * h_ratio) * (1. - w_ratio) +
                       (double)(p[W + 1] * h_ratio * w_ratio);
            o[ph * aw + pw] = (float)v;
          }
        }
      }
    }
  }
}

/* 2x2 stride-1 pooling of (planes, ah, aw) -> (planes, ah-1, aw-1).
 * lib/model/roi_align/modules/roi_align.py:29 (F.avg_pool2d) and :42 (F.max_pool2d).
 * avg: fp32 sum in row-major window order then /4 (ATen avg_pool2d, float accumulate). */
ORC_API void orc_pool2x2(const float *x, long planes, int ah, int aw, int is_max, float *y) {
  int oh = ah - 1, ow = aw - 1;
#pragma omp parallel for schedule(static)
  for (long p = 0; p < planes; ++p) {
    const float *xi = x + p * ah * aw;
    float *yo = y + p * oh * ow;
    for (int i = 0; i < oh; ++i)
      for (int j = 0; j < ow; ++j) {
        float a = xi[i * aw + j], b = xi[i * aw + j + 1];
        float c = xi[(i + 1) * aw + j], d = xi[(i + 1) * aw + j + 1];
        if (is_max) {
          float m = a;
          if (b > m) m = b;
          if (c > m) m = c;
          if (d > m) m = d;
          yo[i * ow + j] = m;
        } else {
          yo[i * ow + j] = (((a + b) + c) + d) / 4.f;
        }
      }
  }
}

/* backward of orc_roi_align_fwd: scatter top_diff*weight into the four taps
 * (roi_align_kernel.cu:99-141).  Accumulates in fp64 (the legacy kernel's fp32 atomics have
 * no defined order; fp64 is the order-independent reference value).  grad_in is (B,C,H,W)
 * fp64 and must be zeroed by the caller (functions/roi_align.py:38-39). */
ORC_API void orc_roi_align_bwd(const float *top_diff, const float *rois, int B, int C, int H,
                               int W, int R, int ah, int aw, float scale, double *grad_in) {
  (void)B;
  /* parallel over channels: every (c) plane is private to one thread -> deterministic */
#pragma omp parallel for schedule(static)
  for (int c = 0; c < C; ++c) {
    for (int n = 0; n < R; ++n) {
      const float *roi = rois + (size_t)n * 5;
      orc_roi_geom g = orc_align_geom(roi, scale, ah, aw);
      long img_start = orc_align_img_start(roi[0], C, H, W);
      double *plane = grad_in + img_start + (long)c * H * W;
      const float *t = top_diff + ((size_t)n * C + c) * ah * aw;
      for (int ph = 0; ph < ah; ++ph) {
        float h = fmaf((float)ph, g.bin_h, g.start_h);
        int hstart = (int)fminf(floorf(h), (float)(H - 2));
        for (int pw = 0; pw < aw; ++pw) {
          float w = fmaf((float)pw, g.bin_w, g.start_w);
          int wstart = (int)fminf(floorf(w), (float)(W - 2));
          if (!(h < 0 || h >= H || w < 0 || w >= W)) {
            float h_ratio = h - (float)hstart, w_ratio = w - (float)wstart;
            double *p = plane + (long)hstart * W + wstart;
            float tf = t[ph * aw + pw];
            double d = tf;
            /* :137-140 -- "(1. - h_ratio)" is fp64, "(1 - w_ratio)" is fp32 */
            p[0] += (double)(float)(d * (1. - h_ratio) * (double)(1.f - w_ratio));
            p[1] += (double)(float)(d * (1. - h_ratio) * (double)w_ratio);
            p[W] += (double)(tf * h_ratio * (1.f - w_ratio));
            p[W + 1] += (double)(tf * h_ratio * w_ratio);
          }
        }
      }
    }
  }
}

/* backward of orc_pool2x2 (autograd of avg_pool2d / max_pool2d, stride 1, kernel 2).
 * avg: every window member receives g/4.  max: the first maximum in row-major window order
 * receives g (ATen max_pool2d argmax rule: strict '>' or NaN).  x is the pre-pool tensor
 * (needed for max only).  gx (planes, ah, aw) fp32 is overwritten. */
ORC_API void orc_pool2x2_bwd(const float *gy, const float *x, long planes, int ah, int aw,
                             int is_max, float *gx) {
  int oh = ah - 1, ow = aw - 1;
#pragma omp parallel for schedule(static)
  for (long p = 0; p < planes; ++p) {
    const float *g = gy + p * oh * ow;
    float *o = gx + p * ah * aw;
    double acc[64 * 64];
    for (int k = 0; k < ah * aw; ++k) acc[k] = 0.;
    for (int i = 0; i < oh; ++i)
      for (int j = 0; j < ow; ++j) {
        int idx[4] = {i * aw + j, i * aw + j + 1, (i + 1) * aw + j, (i + 1) * aw + j + 1};
        if (is_max) {
          const float *xi = x + p * ah * aw;
          int best = idx[0];
          for (int k = 1; k < 4; ++k)
            if (xi[idx[k]] > xi[best]) best = idx[k];
          acc[best] += g[i * ow + j];
        } else {
          for (int k = 0; k < 4; ++k) acc[idx[k]] += (double)(g[i * ow + j] / 4.f);
        }
      }
    for (int k = 0; k < ah * aw; ++k) o[k] = (float)acc[k];
  }
}

/* ------------------------------------------------------------------------------------
 * RoIPool (max).  Parity form: lib/model/roi_pooling/src/roi_pooling_kernel.cu:24-93 --
 * NCHW, argmax = flat index into the whole NCHW tensor, first maximum wins (strict '>'),
 * empty bin -> 0 / -1.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_roi_pool_fwd(const float *feat, const float *rois, int B, int C, int H, int W,
                              int R, int ph_n, int pw_n, float scale, float *out, int *argmax) {
  (void)B;
#pragma omp parallel for schedule(dynamic, 4)
  for (int n = 0; n < R; ++n) {
    const float *roi = rois + (size_t)n * 5;
    int roi_batch_ind = (int)roi[0];
    int roi_start_w = (int)roundf(roi[1] * scale);
    int roi_start_h = (int)roundf(roi[2] * scale);
    int roi_end_w = (int)roundf(roi[3] * scale);
    int roi_end_h = (int)roundf(roi[4] * scale);
    int roi_width = (int)fmaxf((float)(roi_end_w - roi_start_w + 1), 1.f);
    int roi_height = (int)fmaxf((float)(roi_end_h - roi_start_h + 1), 1.f);
    float bin_size_h = (float)roi_height / (float)ph_n;
    float bin_size_w = (float)roi_width / (float)pw_n;
    for (int c = 0; c < C; ++c) {
      int base = (roi_batch_ind * C + c) * H * W;
      for (int ph = 0; ph < ph_n; ++ph)
        for (int pw = 0; pw < pw_n; ++pw) {
          int hstart = (int)floorf((float)ph * bin_size_h);
          int wstart = (int)floorf((float)pw * bin_size_w);
          int hend = (int)ceilf((float)(ph + 1) * bin_size_h);
          int wend = (int)ceilf((float)(pw + 1) * bin_size_w);
          hstart = (int)fminf(fmaxf((float)(hstart + roi_start_h), 0.f), (float)H);
          hend = (int)fminf(fmaxf((float)(hend + roi_start_h), 0.f), (float)H);
          wstart = (int)fminf(fmaxf((float)(wstart + roi_start_w), 0.f), (float)W);
          wend = (int)fminf(fmaxf((float)(wend + roi_start_w), 0.f), (float)W);
          int is_empty = (hend <= hstart) || (wend <= wstart);
          float maxval = is_empty ? 0.f : -FLT_MAX;
          int maxidx = -1;
          for (int h = hstart; h < hend; ++h)
            for (int w = wstart; w < wend; ++w) {
              int idx = base + h * W + w;
              if (feat[idx] > maxval) {
                maxval = feat[idx];
                maxidx = idx;
              }
            }
          size_t oi = (((size_t)n * C + c) * ph_n + ph) * pw_n + pw;
          out[oi] = maxval;
          if (argmax) argmax[oi] = maxidx;
        }
    }
  }
}

/* RoIPool backward, roi_pooling_kernel.cu:128-203, restated as the scatter that produces the
 * same sums as the reference's gather: a pooled bin (n,c,ph,pw) whose argmax is the input
 * element (h,w) contributes its top_diff iff the gather at (h,w) would have visited it, i.e.
 * (h,w) lies inside the ROUNDED roi [start,end] (:161-165 -- false everywhere for an inverted
 * roi, whose gradient the reference therefore drops) and (ph,pw) is in the feasible bin range
 * the gather derives from (h,w) (:178-186).  fp64 accumulate (the reference adds in fp32,
 * roi-major); grad_in (B,C,H,W) fp64, zeroed by the caller (the reference overwrites, :201). */
ORC_API void orc_roi_pool_bwd(const float *top_diff, const int *argmax, const float *rois, int B,
                              int C, int H, int W, int R, int ph_n, int pw_n, float scale,
                              double *grad_in) {
  long n_bottom = (long)B * C * H * W;
  for (int n = 0; n < R; ++n) {
    const float *roi = rois + (size_t)n * 5;
    int roi_start_w = (int)roundf(roi[1] * scale);
    int roi_start_h = (int)roundf(roi[2] * scale);
    int roi_end_w = (int)roundf(roi[3] * scale);
    int roi_end_h = (int)roundf(roi[4] * scale);
    int roi_width = (int)fmaxf((float)(roi_end_w - roi_start_w + 1), 1.f);
    int roi_height = (int)fmaxf((float)(roi_end_h - roi_start_h + 1), 1.f);
    float bin_size_h = (float)roi_height / (float)ph_n;
    float bin_size_w = (float)roi_width / (float)pw_n;
    for (int c = 0; c < C; ++c)
      for (int ph = 0; ph < ph_n; ++ph)
        for (int pw = 0; pw < pw_n; ++pw) {
          size_t ti = (((size_t)n * C + c) * ph_n + ph) * pw_n + pw;
          int a = argmax[ti];
          if (a < 0 || a >= n_bottom) continue;
          int w = a % W, h = (a / W) % H;
          if (!(w >= roi_start_w && w <= roi_end_w && h >= roi_start_h && h <= roi_end_h)) continue;
          int phstart = (int)floorf((float)(h - roi_start_h) / bin_size_h);
          int phend = (int)ceilf((float)(h - roi_start_h + 1) / bin_size_h);
          int pwstart = (int)floorf((float)(w - roi_start_w) / bin_size_w);
          int pwend = (int)ceilf((float)(w - roi_start_w + 1) / bin_size_w);
          phstart = (int)fminf(fmaxf((float)phstart, 0.f), (float)ph_n);
          phend = (int)fminf(fmaxf((float)phend, 0.f), (float)ph_n);
          pwstart = (int)fminf(fmaxf((float)pwstart, 0.f), (float)pw_n);
          pwend = (int)fminf(fmaxf((float)pwend, 0.f), (float)pw_n);
          if (ph < phstart || ph >= phend || pw < pwstart || pw >= pwend) continue;
          grad_in[a] += (double)top_diff[ti];
        }
  }
}

/* The reference's only CPU pooling path, lib/model/roi_pooling/src/roi_pooling.c:4-104:
 * NHWC features, batch 1, output pre-filled with -1 (:30), no argmax.  Used as the timed
 * "reference CPU path" for RoIPool; single-threaded like the original.  Returns 0 when
 * batch != 1 exactly like :17-21. */
ORC_API int orc_roi_pool_fwd_cpu_nhwc(const float *feat_nhwc, const float *rois, int B, int H,
                                      int W, int C, int R, int ph_n, int pw_n, float scale,
                                      float *out) {
  if (B != 1) return 0;
  size_t total = (size_t)R * C * ph_n * pw_n;
  for (size_t i = 0; i < total; ++i) out[i] = -1.f;
  const int output_area = pw_n * ph_n;
  for (int n = 0; n < R; ++n) {
    const float *roi = rois + (size_t)n * 5;
    int roi_batch_ind = (int)roi[0];
    int roi_start_w = (int)round(roi[1] * scale);
    int roi_start_h = (int)round(roi[2] * scale);
    int roi_end_w = (int)round(roi[3] * scale);
    int roi_end_h = (int)round(roi[4] * scale);
    int roi_height = (int)fmaxf((float)(roi_end_h - roi_start_h + 1), 1.f);
    int roi_width = (int)fmaxf((float)(roi_end_w - roi_start_w + 1), 1.f);
    float bin_size_h = (float)roi_height / (float)ph_n;
    float bin_size_w = (float)roi_width / (float)pw_n;
    long index_data = (long)roi_batch_ind * H * W * C;
    float *o = out + (size_t)n * ph_n * pw_n * C;
    for (int ph = 0; ph < ph_n; ++ph)
      for (int pw = 0; pw < pw_n; ++pw) {
        int hstart = (int)floorf((float)ph * bin_size_h);
        int wstart = (int)floorf((float)pw * bin_size_w);
        int hend = (int)ceilf((float)(ph + 1) * bin_size_h);
        int wend = (int)ceilf((float)(pw + 1) * bin_size_w);
        hstart = (int)fminf(fmaxf((float)(hstart + roi_start_h), 0.f), (float)H);
        hend = (int)fminf(fmaxf((float)(hend + roi_start_h), 0.f), (float)H);
        wstart = (int)fminf(fmaxf((float)(wstart + roi_start_w), 0.f), (float)W);
        wend = (int)fminf(fmaxf((float)(wend + roi_start_w), 0.f), (float)W);
        const int pool_index = ph * pw_n + pw;
        if ((hend <= hstart) || (wend <= wstart)) {
          for (int c = 0; c < C; ++c) o[pool_index + c * output_area] = 0.f;
        } else {
          for (int h = hstart; h < hend; ++h)
            for (int w = wstart; w < wend; ++w) {
              const float *px = feat_nhwc + index_data + ((long)h * W + w) * C;
              for (int c = 0; c < C; ++c)
                if (px[c] > o[pool_index + c * output_area]) o[pool_index + c * output_area] = px[c];
            }
        }
      }
  }
  return 1;
}

/* ------------------------------------------------------------------------------------
 * Box algebra.  lib/model/rpn/bbox_transform.py:77-103 (bbox_transform_inv, torch fp32
 * elementwise: every op rounds separately), :125-133 (clip_boxes), :136-166 (bbox_overlaps).
 * ---------------------------------------------------------------------------------- */
static inline void orc_decode_one(const float *box, const float *d, float *o) {
  float w = box[2] - box[0] + 1.0f, h = box[3] - box[1] + 1.0f;
  float cx = box[0] + 0.5f * w, cy = box[1] + 0.5f * h;
  float pcx = d[0] * w + cx, pcy = d[1] * h + cy;
  float pw = expf(d[2]) * w, ph = expf(d[3]) * h;
  o[0] = pcx - 0.5f * pw;
  o[1] = pcy - 0.5f * ph;
  o[2] = pcx + 0.5f * pw;
  o[3] = pcy + 0.5f * ph;
}

static inline float orc_clampf(float v, float lo, float hi) {
  /* torch.clamp_: min(max(v, lo), hi); NaN propagates */
  if (v != v) return v;
  return fminf(fmaxf(v, lo), hi);
}

/* boxes (B,N,4), deltas (B,N,4k) -> out (B,N,4k) */
ORC_API void orc_bbox_transform_inv(const float *boxes, const float *deltas, long BN, int k,
                                    float *out) {
  for (long i = 0; i < BN; ++i)
    for (int j = 0; j < k; ++j)
      orc_decode_one(boxes + i * 4, deltas + (i * k + j) * 4, out + (i * k + j) * 4);
}

/* in place; im_info (B,3) = [h, w, scale] */
ORC_API void orc_clip_boxes(float *boxes, const float *im_info, int B, long N, int k) {
  for (int b = 0; b < B; ++b) {
    float xmax = im_info[b * 3 + 1] - 1.f, ymax = im_info[b * 3 + 0] - 1.f;
    float *p = boxes + (size_t)b * N * k * 4;
    for (long i = 0; i < N * k; ++i) {
      p[i * 4 + 0] = orc_clampf(p[i * 4 + 0], 0.f, xmax);
      p[i * 4 + 1] = orc_clampf(p[i * 4 + 1], 0.f, ymax);
      p[i * 4 + 2] = orc_clampf(p[i * 4 + 2], 0.f, xmax);
      p[i * 4 + 3] = orc_clampf(p[i * 4 + 3], 0.f, ymax);
    }
  }
}

/* anchors (N,4), gt (K,4) -> (N,K); +1 convention, fp32 (bbox_transform.py:136-166) */
ORC_API void orc_bbox_overlaps(const float *anchors, const float *gt, int N, int K, float *out) {
  for (int n = 0; n < N; ++n) {
    const float *a = anchors + (size_t)n * 4;
    float aa = (a[2] - a[0] + 1.f) * (a[3] - a[1] + 1.f);
    for (int k = 0; k < K; ++k) {
      const float *g = gt + (size_t)k * 4;
      float ga = (g[2] - g[0] + 1.f) * (g[3] - g[1] + 1.f);
      float iw = fminf(a[2], g[2]) - fmaxf(a[0], g[0]) + 1.f;
      if (iw < 0) iw = 0;
      float ih = fminf(a[3], g[3]) - fmaxf(a[1], g[1]) + 1.f;
      if (ih < 0) ih = 0;
      float ua = aa + ga - (iw * ih);
      out[(size_t)n * K + k] = iw * ih / ua;
    }
  }
}

/* anchors (B,N,4), gt (B,K,4) -> (B,N,K) with the degenerate-box sentinels of
 * bbox_overlaps_batch (bbox_transform.py:195-196, 212-213): gt w==1&&h==1 -> 0,
 * anchor w==1&&h==1 -> -1 (applied after, so it wins). */
ORC_API void orc_bbox_overlaps_batch(const float *anchors, const float *gt, int B, int N, int K,
                                     float *out) {
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < N; ++n) {
      const float *a = anchors + ((size_t)b * N + n) * 4;
      float ax = a[2] - a[0] + 1.f, ay = a[3] - a[1] + 1.f, aa = ax * ay;
      int a_zero = (ax == 1.f) && (ay == 1.f);
      for (int k = 0; k < K; ++k) {
        const float *g = gt + ((size_t)b * K + k) * 4;
        float gx = g[2] - g[0] + 1.f, gy = g[3] - g[1] + 1.f, ga = gx * gy;
        int g_zero = (gx == 1.f) && (gy == 1.f);
        float iw = fminf(a[2], g[2]) - fmaxf(a[0], g[0]) + 1.f;
        if (iw < 0) iw = 0;
        float ih = fminf(a[3], g[3]) - fmaxf(a[1], g[1]) + 1.f;
        if (ih < 0) ih = 0;
        float ua = aa + ga - (iw * ih);
        float v = iw * ih / ua;
        if (g_zero) v = 0.f;
        if (a_zero) v = -1.f;
        out[((size_t)b * N + n) * K + k] = v;
      }
    }
}

/* ------------------------------------------------------------------------------------
 * RPN proposal layer.  lib/model/rpn/proposal_layer.py:49-161.
 *   scores (B,2A,H,W) NCHW, fg = channels [A,2A) (:67); deltas (B,4A,H,W); im_info (B,3);
 *   anchors (A,4) fp32 from generate_anchors (:36-37); anchor index = (y*W+x)*A + a (:92-93).
 *   sort: descending by score, ties -> lower anchor index first (torch.sort's tie order is
 *   unspecified, :125; this is the documented rule).  No min-size filter (:113).
 *   out (B, post_nms_topN, 5) zero padded, column 0 = image index on every row (:127,158).
 * Optional taps for staged parity: order_out (B,pre) sorted anchor indices, props_out
 * (B,pre,4) decoded+clipped sorted boxes, keep_out (B,post) / nkeep_out (B).
 * boxes_override (B,pre,4): if non-NULL the NMS runs on these boxes instead of the CPU
 * decoded ones (lets a test feed the GPU's decoded boxes; expf differs by an ulp between
 * glibc and libdevice).
 * ---------------------------------------------------------------------------------- */
typedef struct {
  float s;
  int idx;
} orc_sc;

static int orc_sc_cmp(const void *pa, const void *pb) {
  const orc_sc *a = (const orc_sc *)pa, *b = (const orc_sc *)pb;
  if (a->s > b->s) return -1;
  if (a->s < b->s) return 1;
  return (a->idx > b->idx) - (a->idx < b->idx);
}

ORC_API void orc_proposal_layer(const float *scores, const float *deltas, const float *im_info,
                                const float *anchors, int B, int A, int H, int W,
                                int feat_stride, int pre_nms_topN, int post_nms_topN,
                                float nms_thresh, float *out, int *order_out, float *props_out,
                                int *keep_out, int *nkeep_out, const float *boxes_override) {
  const int KA = H * W * A;
  const int HW = H * W;
  int pre = (pre_nms_topN > 0 && pre_nms_topN < KA) ? pre_nms_topN : KA;
  memset(out, 0, sizeof(float) * (size_t)B * post_nms_topN * 5);
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    orc_sc *sc = (orc_sc *)malloc(sizeof(orc_sc) * (size_t)KA);
    float *props = (float *)malloc(sizeof(float) * (size_t)pre * 5);
    int *keep = (int *)malloc(sizeof(int) * (size_t)pre);
    const float *fg = scores + ((size_t)b * 2 * A + A) * HW;
    const float *dl = deltas + (size_t)b * 4 * A * HW;
    for (int pix = 0; pix < HW; ++pix)
      for (int a = 0; a < A; ++a) {
        sc[pix * A + a].s = fg[(size_t)a * HW + pix];
        sc[pix * A + a].idx = pix * A + a;
      }
    qsort(sc, (size_t)KA, sizeof(orc_sc), orc_sc_cmp);
    float xmax = im_info[b * 3 + 1] - 1.f, ymax = im_info[b * 3 + 0] - 1.f;
    for (int i = 0; i < pre; ++i) {
      int idx = sc[i].idx, a = idx % A, pix = idx / A;
      int y = pix / W, x = pix % W;
      float sx = (float)(x * feat_stride), sy = (float)(y * feat_stride);
      float box[4] = {anchors[a * 4 + 0] + sx, anchors[a * 4 + 1] + sy, anchors[a * 4 + 2] + sx,
                      anchors[a * 4 + 3] + sy};
      float d[4];
      for (int k = 0; k < 4; ++k) d[k] = dl[(size_t)(4 * a + k) * HW + pix];
      float p[4];
      orc_decode_one(box, d, p);
      p[0] = orc_clampf(p[0], 0.f, xmax);
      p[1] = orc_clampf(p[1], 0.f, ymax);
      p[2] = orc_clampf(p[2], 0.f, xmax);
      p[3] = orc_clampf(p[3], 0.f, ymax);
      if (boxes_override) memcpy(p, boxes_override + ((size_t)b * pre + i) * 4, sizeof(p));
      memcpy(props + (size_t)i * 5, p, sizeof(p));
      props[(size_t)i * 5 + 4] = sc[i].s;
      if (order_out) order_out[(size_t)b * pre + i] = idx;
      if (props_out) memcpy(props_out + ((size_t)b * pre + i) * 4, p, sizeof(p));
    }
    int nk = orc_nms(props, pre, 5, nms_thresh, post_nms_topN, keep);
    if (post_nms_topN > 0 && nk > post_nms_topN) nk = post_nms_topN;
    for (int i = 0; i < post_nms_topN; ++i) {
      float *o = out + ((size_t)b * post_nms_topN + i) * 5;
      o[0] = (float)b;
      if (i < nk) memcpy(o + 1, props + (size_t)keep[i] * 5, sizeof(float) * 4);
      if (keep_out) keep_out[(size_t)b * post_nms_topN + i] = (i < nk) ? keep[i] : -1;
    }
    if (nkeep_out) nkeep_out[b] = nk;
    free(sc);
    free(props);
    free(keep);
  }
}

/* ------------------------------------------------------------------------------------
 * RL reward / label arithmetic.
 *   bbIou: lib/pycocotools/maskApi.c:98-109 (xywh, no +1, crowd -> union = dt area, fp64).
 *   reward loop: lib/datasets/RL_coco_dataset.py:119-137.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_bbiou(const double *dt, const double *gt, long m, long n,
                       const unsigned char *iscrowd, double *o) {
  for (long g = 0; g < n; ++g) {
    const double *G = gt + g * 4;
    double ga = G[2] * G[3];
    int crowd = iscrowd != NULL && iscrowd[g];
    for (long d = 0; d < m; ++d) {
      const double *D = dt + d * 4;
      double da = D[2] * D[3];
      o[g * m + d] = 0;
      double w = fmin(D[2] + D[0], G[2] + G[0]) - fmax(D[0], G[0]);
      if (w <= 0) continue;
      double h = fmin(D[3] + D[1], G[3] + G[1]) - fmax(D[1], G[1]);
      if (h <= 0) continue;
      double i = w * h, u = crowd ? da : da + ga - i;
      o[g * m + d] = i / u;
    }
  }
}

/* mode 0 ("coco"): boxes/gt are xywh, fp64 IoU via bbIou.  The reference holds boxes as
 *   Python floats (fp64) and actDeltas as fp32 numpy; `bbox + act_delta*np.array([w,h,w,h])`
 *   (:124) promotes to fp64.  Inputs here are fp32 tensors, widened exactly.
 * mode 1 ("rcnn"): boxes/gt are x1y1x2y2, action applied on (x1,y1,w,h) then converted
 *   back, fp32 bbox_overlaps (+1 convention).
 * boxes (B,N,4) f32, gt (B,G,4) f32, crowd (B,G) u8 or NULL, ngt (B) int or NULL (valid gt
 * count per image; 0 -> one all-zero gt, :113-117), act (A,4) f32.
 * reward (B,N,A) f32 = max_g IoU(new) - max_g IoU(orig) (:126).  Optional label (+1/-1 by
 * reward > iou_thres, :128-134) and weight = wtrans(reward) * (pos|neg)_wratio.
 */
static double orc_max_bbiou(const double *dt, const float *gt, const unsigned char *crowd,
                            int ng) {
  double best = -INFINITY;
  if (ng <= 0) {
    double z[4] = {0, 0, 0, 0}, o;
    orc_bbiou(dt, z, 1, 1, NULL, &o);
    return o;
  }
  for (int g = 0; g < ng; ++g) {
    double G[4] = {gt[g * 4 + 0], gt[g * 4 + 1], gt[g * 4 + 2], gt[g * 4 + 3]}, o;
    unsigned char cr = crowd ? crowd[g] : 0;
    orc_bbiou(dt, G, 1, 1, &cr, &o);
    if (o > best) best = o;
  }
  return best;
}

static float orc_max_overlap_f32(const float *box, const float *gt, int ng) {
  float best = -INFINITY;
  float z[4] = {0, 0, 0, 0};
  if (ng <= 0) {
    float o;
    orc_bbox_overlaps(box, z, 1, 1, &o);
    return o;
  }
  for (int g = 0; g < ng; ++g) {
    float o;
    orc_bbox_overlaps(box, gt + g * 4, 1, 1, &o);
    if (o > best) best = o;
  }
  return best;
}

ORC_API void orc_action_reward(const float *boxes, const float *gt, const unsigned char *crowd,
                               const int *ngt, const float *act, int B, int N, int A, int G,
                               int mode, int wtrans, float iou_thres, float pos_wratio,
                               float neg_wratio, float *reward, float *label, float *weight) {
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b) {
    const float *gtb = gt + (size_t)b * G * 4;
    const unsigned char *crb = crowd ? crowd + (size_t)b * G : NULL;
    int ng = ngt ? ngt[b] : G;
    for (int n = 0; n < N; ++n) {
      const float *bx = boxes + ((size_t)b * N + n) * 4;
      for (int a = 0; a < A; ++a) {
        size_t oi = ((size_t)b * N + n) * A + a;
        double r;
        if (mode == 0) {
          double w = bx[2], h = bx[3];
          double dt[4] = {bx[0], bx[1], bx[2], bx[3]};
          double nb[4] = {dt[0] + (double)act[a * 4 + 0] * w, dt[1] + (double)act[a * 4 + 1] * h,
                          dt[2] + (double)act[a * 4 + 2] * w, dt[3] + (double)act[a * 4 + 3] * h};
          r = orc_max_bbiou(nb, gtb, crb, ng) - orc_max_bbiou(dt, gtb, crb, ng);
          reward[oi] = (float)r;
        } else {
          float w = bx[2] - bx[0] + 1.f, h = bx[3] - bx[1] + 1.f;
          float x = bx[0] + act[a * 4 + 0] * w, y = bx[1] + act[a * 4 + 1] * h;
          float nw = w + act[a * 4 + 2] * w, nh = h + act[a * 4 + 3] * h;
          float nb[4] = {x, y, x + nw - 1.f, y + nh - 1.f};
          float rf = orc_max_overlap_f32(nb, gtb, ng) - orc_max_overlap_f32(bx, gtb, ng);
          reward[oi] = rf;
          r = rf;
        }
        /* label / weight from the un-rounded reward (fp64 in coco mode, fp32 in rcnn mode) */
        int pos = r > (double)iou_thres;
        if (label) label[oi] = pos ? 1.f : -1.f;
        /* weight = wtrans(delta_iou) * ratio (:130-135); wtrans 0 = Identify (action.py:7-10),
         * 1 = exp(|x|) (config.py:48-51) */
        if (weight)
          weight[oi] = (float)((wtrans == 1 ? exp(fabs(r)) : r) * (double)(pos ? pos_wratio : neg_wratio));
      }
    }
  }
}

/* NCHW <-> NHWC helpers used by the reference-arm timing of roi_pooling.c */
ORC_API void orc_nchw_to_nhwc(const float *x, int B, int C, int H, int W, float *y) {
  long HW = (long)H * W;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b)
    for (long p = 0; p < HW; ++p)
      for (int c = 0; c < C; ++c) y[((long)b * HW + p) * C + c] = x[((long)b * C + c) * HW + p];
}
