// RL refinement step for sm_100a.
//
// k_action_reward: the per-box x per-action label loop of COCODataset.__getitem__
// (lib/datasets/RL_coco_dataset.py:119-137) -- in the reference N_boxes * num_acts separate
// Python -> Cython -> C calls of pycocotools bbIou (lib/pycocotools/maskApi.c:98-109) per
// image inside DataLoader workers -- as one launch: a thread per (image, box, action).
//   RLOD_IOU_COCO: xywh boxes, fp64, every operation rounded separately exactly like the C
//     source compiled without contraction -> bit-exact rewards.
//   RLOD_IOU_RCNN: x1y1x2y2 boxes, fp32 bbox_overlaps (lib/model/rpn/bbox_transform.py:
//     136-166), the action applied to (x1, y1, w, h).
// k_move_from_act: Action.move_from_act (lib/model/Reinforcement/action.py:25-59) on the
// device, one CTA per image: per-box best action, bitonic sort of the boxes by
// (pred descending, flat index DESCENDING among equal preds = np.flip(np.argsort(kind='stable'))),
// the first maxk boxes move if their target is 1.
#include "rlod_common.cuh"

namespace rlod {

__device__ __forceinline__ double bbiou_f64(const double *D, const double *G, bool crowd) {
  const double ga = __dmul_rn(G[2], G[3]), da = __dmul_rn(D[2], D[3]);
  const double w = __dsub_rn(fmin(__dadd_rn(D[2], D[0]), __dadd_rn(G[2], G[0])), fmax(D[0], G[0]));
  if (w <= 0) return 0.;
  const double h = __dsub_rn(fmin(__dadd_rn(D[3], D[1]), __dadd_rn(G[3], G[1])), fmax(D[1], G[1]));
  if (h <= 0) return 0.;
  const double i = __dmul_rn(w, h);
  const double u = crowd ? da : __dsub_rn(__dadd_rn(da, ga), i);
  return __ddiv_rn(i, u);
}

template <typename T>
__device__ __forceinline__ double max_bbiou(const double *D, const T *__restrict__ gt,
                                            const unsigned char *__restrict__ crowd, int ng) {
  if (ng <= 0) {
    const double z[4] = {0., 0., 0., 0.};
    return bbiou_f64(D, z, false);
  }
  double best = -INFINITY;
  for (int g = 0; g < ng; ++g) {
    const double G[4] = {(double)__ldg(gt + g * 4), (double)__ldg(gt + g * 4 + 1),
                         (double)__ldg(gt + g * 4 + 2), (double)__ldg(gt + g * 4 + 3)};
    const double o = bbiou_f64(D, G, crowd ? (crowd[g] != 0) : false);
    if (o > best) best = o;
  }
  return best;
}

// Action.wtrans (action.py:7-10 default Identify; config.py:48-51 exp(|x|)) applied to delta_iou,
// times the positive / negative ratio (RL_coco_dataset.py:128-135).  RLOD_WTRANS_RAW hands
// delta_iou itself back so that the caller can apply any other callable.
__device__ __forceinline__ float label_weight(double r, int wtrans, double ratio) {
  if (wtrans == RLOD_WTRANS_EXP_ABS) return (float)(exp(fabs(r)) * ratio);
  if (wtrans == RLOD_WTRANS_IDENTITY) return (float)(r * ratio);
  return (float)r;
}

__device__ __forceinline__ float overlap_f32(const float *a, const float *g) {
  const float aa = __fmul_rn(__fadd_rn(__fsub_rn(a[2], a[0]), 1.f), __fadd_rn(__fsub_rn(a[3], a[1]), 1.f));
  const float ga = __fmul_rn(__fadd_rn(__fsub_rn(g[2], g[0]), 1.f), __fadd_rn(__fsub_rn(g[3], g[1]), 1.f));
  float iw = __fadd_rn(__fsub_rn(fminf(a[2], g[2]), fmaxf(a[0], g[0])), 1.f);
  if (iw < 0.f) iw = 0.f;
  float ih = __fadd_rn(__fsub_rn(fminf(a[3], g[3]), fmaxf(a[1], g[1])), 1.f);
  if (ih < 0.f) ih = 0.f;
  const float inter = __fmul_rn(iw, ih);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ga), inter));
}

__device__ __forceinline__ float max_overlap_f32(const float *a, const float *__restrict__ gt, int ng) {
  if (ng <= 0) {
    const float z[4] = {0.f, 0.f, 0.f, 0.f};
    return overlap_f32(a, z);
  }
  float best = -INFINITY;
  for (int g = 0; g < ng; ++g) {
    const float G[4] = {__ldg(gt + g * 4), __ldg(gt + g * 4 + 1), __ldg(gt + g * 4 + 2),
                        __ldg(gt + g * 4 + 3)};
    const float o = overlap_f32(a, G);
    if (o > best) best = o;
  }
  return best;
}

// the same scan over ground truth held in shared memory (plain loads: __ldg is global-only)
__device__ __forceinline__ float max_overlap_f32_smem(const float *a, const float *gt, int ng) {
  if (ng <= 0) {
    const float z[4] = {0.f, 0.f, 0.f, 0.f};
    return overlap_f32(a, z);
  }
  float best = -INFINITY;
  for (int g = 0; g < ng; ++g) {
    const float o = overlap_f32(a, gt + g * 4);
    if (o > best) best = o;
  }
  return best;
}

template <typename T>
__global__ void __launch_bounds__(256)
    k_action_reward(const T *__restrict__ boxes, const T *__restrict__ gt,
                    const unsigned char *__restrict__ crowd, const int *__restrict__ ngt,
                    const float *__restrict__ act, int B, int N, int A, int G, int mode, int wtrans,
                    float iou_thres, float pos_wratio, float neg_wratio, float *__restrict__ reward,
                    float *__restrict__ label, float *__restrict__ weight) {
  const long long total = (long long)B * N * A;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(idx % A);
    const long long bn = idx / A;
    const int b = (int)(bn / N);
    const T *bx = boxes + bn * 4;
    const T *gtb = gt + (size_t)b * G * 4;
    const unsigned char *crb = crowd ? crowd + (size_t)b * G : nullptr;
    int ng = ngt ? __ldg(ngt + b) : G;
    if (ng > G) ng = G;
    const T x0 = __ldg(bx), x1 = __ldg(bx + 1), x2 = __ldg(bx + 2), x3 = __ldg(bx + 3);
    const float a0 = __ldg(act + a * 4), a1 = __ldg(act + a * 4 + 1);
    const float a2 = __ldg(act + a * 4 + 2), a3 = __ldg(act + a * 4 + 3);
    double r;
    float rf;
    if (mode == RLOD_IOU_COCO) {
      // bbox + act_delta * np.array([w, h, w, h])  (RL_coco_dataset.py:124): fp64
      const double w = x2, h = x3;
      const double dt[4] = {x0, x1, x2, x3};
      const double nb[4] = {__dadd_rn(dt[0], __dmul_rn((double)a0, w)),
                            __dadd_rn(dt[1], __dmul_rn((double)a1, h)),
                            __dadd_rn(dt[2], __dmul_rn((double)a2, w)),
                            __dadd_rn(dt[3], __dmul_rn((double)a3, h))};
      r = __dsub_rn(max_bbiou(nb, gtb, crb, ng), max_bbiou(dt, gtb, crb, ng));
      rf = (float)r;
    } else if constexpr (sizeof(T) == 4) {
      const float w = __fadd_rn(__fsub_rn(x2, x0), 1.f), h = __fadd_rn(__fsub_rn(x3, x1), 1.f);
      const float nx = __fadd_rn(x0, __fmul_rn(a0, w)), ny = __fadd_rn(x1, __fmul_rn(a1, h));
      const float nw = __fadd_rn(w, __fmul_rn(a2, w)), nh = __fadd_rn(h, __fmul_rn(a3, h));
      const float nb[4] = {nx, ny, __fsub_rn(__fadd_rn(nx, nw), 1.f), __fsub_rn(__fadd_rn(ny, nh), 1.f)};
      const float ob[4] = {x0, x1, x2, x3};
      rf = __fsub_rn(max_overlap_f32(nb, gtb, ng), max_overlap_f32(ob, gtb, ng));
      r = rf;
    } else {
      r = 0., rf = 0.f;  // fp64 boxes exist in COCO mode only (host-checked)
    }
    reward[idx] = rf;
    const bool pos = r > (double)iou_thres;
    if (label) label[idx] = pos ? 1.f : -1.f;
    if (weight) weight[idx] = label_weight(r, wtrans, (double)(pos ? pos_wratio : neg_wratio));
  }
}

// k_rl_labels: the label tensor of one collated RL batch, straight in the layout
// COCODataLoader._collate_fn hands to the model (lib/datasets/RL_coco_loader.py:19-76):
// labels (B, N, A, 3) = (act_id, label, weight), zero rows for padded boxes.  Per box only the
// ground truth of ITS category counts (RL_coco_dataset.py:108-117: gt_boxes[img_id, cat_id]);
// none -> the single all-zero gt.  IoU = pycocotools bbIou, fp64 (xywh boxes).
template <typename T>
__global__ void __launch_bounds__(256)
    k_rl_labels(const T *__restrict__ dets, int det_stride, const int *__restrict__ det_cat,
                const int *__restrict__ ndet, const T *__restrict__ gt, const int *__restrict__ gt_cat,
                const unsigned char *__restrict__ crowd, const int *__restrict__ ngt,
                const float *__restrict__ act, int B, int N, int A, int G, int wtrans, float iou_thres,
                float pos_wratio, float neg_wratio, float *__restrict__ labels) {
  const long long total = (long long)B * N * A;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(idx % A);
    const long long bn = idx / A;
    const int b = (int)(bn / N), n = (int)(bn - (long long)b * N);
    float *o = labels + idx * 3;
    if (ndet && n >= __ldg(ndet + b)) {
      o[0] = 0.f, o[1] = 0.f, o[2] = 0.f;  // collate pads with zeros (:70-72)
      continue;
    }
    const T *bx = dets + bn * det_stride;
    const int cat = det_cat ? __ldg(det_cat + bn) : 0;
    int ng = ngt ? __ldg(ngt + b) : G;
    if (ng > G) ng = G;
    const double dt[4] = {(double)__ldg(bx), (double)__ldg(bx + 1), (double)__ldg(bx + 2), (double)__ldg(bx + 3)};
    const double w = dt[2], h = dt[3];
    const double nb[4] = {__dadd_rn(dt[0], __dmul_rn((double)__ldg(act + a * 4), w)),
                          __dadd_rn(dt[1], __dmul_rn((double)__ldg(act + a * 4 + 1), h)),
                          __dadd_rn(dt[2], __dmul_rn((double)__ldg(act + a * 4 + 2), w)),
                          __dadd_rn(dt[3], __dmul_rn((double)__ldg(act + a * 4 + 3), h))};
    double bo = -INFINITY, bn_ = -INFINITY;
    bool any = false;
    for (int g = 0; g < ng; ++g) {
      if (gt_cat && __ldg(gt_cat + (size_t)b * G + g) != cat) continue;
      const T *gp = gt + ((size_t)b * G + g) * 4;
      const double Gb[4] = {(double)__ldg(gp), (double)__ldg(gp + 1), (double)__ldg(gp + 2), (double)__ldg(gp + 3)};
      const bool cr = crowd ? (crowd[(size_t)b * G + g] != 0) : false;
      const double o0 = bbiou_f64(dt, Gb, cr), o1 = bbiou_f64(nb, Gb, cr);
      if (o0 > bo) bo = o0;
      if (o1 > bn_) bn_ = o1;
      any = true;
    }
    if (!any) {
      const double z[4] = {0., 0., 0., 0.};
      bo = bbiou_f64(dt, z, false), bn_ = bbiou_f64(nb, z, false);
    }
    const double r = __dsub_rn(bn_, bo);
    const bool pos = r > (double)iou_thres;
    o[0] = (float)a;
    o[1] = pos ? 1.f : -1.f;
    o[2] = label_weight(r, wtrans, (double)(pos ? pos_wratio : neg_wratio));
  }
}

// float -> uint32 whose ascending order is the float's descending order
__device__ __forceinline__ uint32_t desc_key32(float f) {
  uint32_t u = __float_as_uint(f);
  if (f == 0.f) u = 0u;
  const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~asc;
}

constexpr int kMoveThreads = 256;
constexpr int kMoveMaxN = 4096;

__global__ void __launch_bounds__(kMoveThreads)
    k_move_from_act(float *__restrict__ boxes, int box_stride, int corners,
                    const float *__restrict__ preds, const float *__restrict__ targets,
                    const float *__restrict__ act, int N, int A, int maxk, int np2,
                    int *__restrict__ correct) {
  extern __shared__ unsigned long long mkeys[];  // [np2]
  const int b = blockIdx.x, t = threadIdx.x;
  const float *pr = preds + (size_t)b * N * A;
  for (int n = t; n < np2; n += kMoveThreads) {
    unsigned long long key = ~0ull;
    if (n < N) {
      // best action of the box: max pred; among equal preds the reference's
      // np.flip(np.argsort(...)) visits the HIGHER flat index first wherever the sort is stable
      // (action.py:44), so ties go to the highest action id, then to the highest box
      int best = 0;
      float bv = __ldg(pr + (size_t)n * A);
      for (int a = 1; a < A; ++a) {
        const float v = __ldg(pr + (size_t)n * A + a);
        if (v >= bv) {
          bv = v;
          best = a;
        }
      }
      key = ((unsigned long long)desc_key32(bv) << 32) | (0xffffffffu - (unsigned)(n * A + best));
    }
    mkeys[n] = key;
  }
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int p = t; p < (np2 >> 1); p += kMoveThreads) {
        const int i = ((p / j) * (j << 1)) + (p % j), l = i + j;
        const unsigned long long x = mkeys[i], y = mkeys[l];
        if ((x > y) == ((i & k) == 0)) {
          mkeys[i] = y;
          mkeys[l] = x;
        }
      }
      __syncthreads();
    }
  const int take = min(maxk, N);
  int local = 0;
  for (int i = t; i < take; i += kMoveThreads) {
    const unsigned flat = 0xffffffffu - (unsigned)(mkeys[i] & 0xffffffffull);
    const int n = (int)(flat / (unsigned)A), a = (int)(flat % (unsigned)A);
    if (__ldg(targets + ((size_t)b * N + n) * A + a) == 1.f) {
      ++local;
      float *bx = boxes + ((size_t)b * N + n) * box_stride;
      const float a0 = __ldg(act + a * 4 + 0), a1 = __ldg(act + a * 4 + 1);
      const float a2 = __ldg(act + a * 4 + 2), a3 = __ldg(act + a * 4 + 3);
      if (!corners) {
        // bboxes[bid][idx] += delta * np.array([w, h, w, h])   (xywh, fp32, action.py:55)
        const float w = bx[2], h = bx[3];
        bx[0] = __fadd_rn(bx[0], __fmul_rn(a0, w));
        bx[1] = __fadd_rn(bx[1], __fmul_rn(a1, h));
        bx[2] = __fadd_rn(bx[2], __fmul_rn(a2, w));
        bx[3] = __fadd_rn(bx[3], __fmul_rn(a3, h));
      } else {
        // x1y1x2y2 boxes (+1 convention): the same move applied to (x1, y1, w, h), identical
        // to the arithmetic of k_action_reward's RLOD_IOU_RCNN mode
        const float w = __fadd_rn(__fsub_rn(bx[2], bx[0]), 1.f);
        const float h = __fadd_rn(__fsub_rn(bx[3], bx[1]), 1.f);
        const float nx = __fadd_rn(bx[0], __fmul_rn(a0, w)), ny = __fadd_rn(bx[1], __fmul_rn(a1, h));
        const float nw = __fadd_rn(w, __fmul_rn(a2, w)), nh = __fadd_rn(h, __fmul_rn(a3, h));
        bx[0] = nx;
        bx[1] = ny;
        bx[2] = __fsub_rn(__fadd_rn(nx, nw), 1.f);
        bx[3] = __fsub_rn(__fadd_rn(ny, nh), 1.f);
      }
    }
  }
  // block reduce of the per-thread counts
  for (int d = 16; d > 0; d >>= 1) local += __shfl_down_sync(0xffffffffu, local, d);
  if ((t & 31) == 0 && local && correct) atomicAdd(correct, local);
}


// ----------------------------------------------------------------------------------------
// k_reward_refine: the hot path's whole RL-refine in ONE launch -- what the step used to do
// as rlod_action_reward (RCNN mode) + a copy of the rois + rlod_move_from_act(maxk = N, the
// rewards as predictions, the labels as targets) + concatenation + global image index:
//   reward[b,n,a] = max_g IoU(box moved by action a, gt_g) - max_g IoU(box, gt_g)
//   every box takes its best action (highest reward, ties -> highest action id, the visit
//   order of move_from_act) if that action's label is +1 (reward > iou_thres)
//   refined (B,N,5) = [b, moved box];  packed (B,N,5+A) = [b + first_image, moved box, rewards]
// One warp per box, lanes over the actions, the image's ground truth in shared memory; the
// unmoved box's max IoU is computed once per box (lanes over gt).  Arithmetic is the
// reward / move kernels' own, operation for operation, so the results are bit-identical to
// the unfused sequence (tests/test_gpu_parity.py::test_reward_refine_equals_unfused).
// ----------------------------------------------------------------------------------------
constexpr int kRefineWarps = 8;

__global__ void __launch_bounds__(kRefineWarps * 32)
    k_reward_refine(const float *__restrict__ rois, const float *__restrict__ gt,
                    const int *__restrict__ ngt, const float *__restrict__ act, int N, int A, int G,
                    int wtrans, float iou_thres, float pos_wratio, float neg_wratio,
                    float first_image, float *__restrict__ reward, float *__restrict__ label,
                    float *__restrict__ weight, float *__restrict__ refined,
                    float *__restrict__ packed, int *__restrict__ moved) {
  extern __shared__ float s_gt[];  // [G * 4]
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  for (int i = threadIdx.x; i < G * 4; i += blockDim.x) s_gt[i] = __ldg(gt + (size_t)b * G * 4 + i);
  __syncthreads();
  const int n = blockIdx.x * kRefineWarps + warp;
  if (n >= N) return;
  int ng = ngt ? __ldg(ngt + b) : G;
  if (ng > G) ng = G;
  const size_t bn = (size_t)b * N + n;
  const float *roi = rois + bn * 5;
  const float x0 = __ldg(roi + 1), x1 = __ldg(roi + 2), x2 = __ldg(roi + 3), x3 = __ldg(roi + 4);
  const float ob[4] = {x0, x1, x2, x3};
  // max IoU of the unmoved box: lanes over gt, warp max (max is exact, any order)
  float orig;
  if (ng <= 0) {
    const float z[4] = {0.f, 0.f, 0.f, 0.f};
    orig = overlap_f32(ob, z);
  } else {
    float best = -INFINITY;
    for (int g = lane; g < ng; g += 32) {
      const float o = overlap_f32(ob, s_gt + g * 4);
      if (o > best) best = o;
    }
    for (int d = 16; d > 0; d >>= 1) best = fmaxf(best, __shfl_xor_sync(full, best, d));
    orig = best;
  }
  const float w = __fadd_rn(__fsub_rn(x2, x0), 1.f), h = __fadd_rn(__fsub_rn(x3, x1), 1.f);
  // lanes over actions; per lane the best (reward, action) seen, later actions win ties
  float bv = 0.f;
  int ba = -1;
  bool first_nan = false;
  for (int a = lane; a < A; a += 32) {
    const float a0 = __ldg(act + a * 4), a1 = __ldg(act + a * 4 + 1);
    const float a2 = __ldg(act + a * 4 + 2), a3 = __ldg(act + a * 4 + 3);
    const float nx = __fadd_rn(x0, __fmul_rn(a0, w)), ny = __fadd_rn(x1, __fmul_rn(a1, h));
    const float nw = __fadd_rn(w, __fmul_rn(a2, w)), nh = __fadd_rn(h, __fmul_rn(a3, h));
    const float nb[4] = {nx, ny, __fsub_rn(__fadd_rn(nx, nw), 1.f), __fsub_rn(__fadd_rn(ny, nh), 1.f)};
    const float rf = __fsub_rn(max_overlap_f32_smem(nb, s_gt, ng), orig);
    const bool pos = (double)rf > (double)iou_thres;
    const size_t oi = bn * A + a;
    if (reward) reward[oi] = rf;
    if (label) label[oi] = pos ? 1.f : -1.f;
    if (weight) weight[oi] = label_weight((double)rf, wtrans, (double)(pos ? pos_wratio : neg_wratio));
    if (packed) packed[bn * (size_t)(5 + A) + 5 + a] = rf;
    if (a == 0 && rf != rf) first_nan = true;
    if (rf == rf && (ba < 0 || rf >= bv)) bv = rf, ba = a;  // a NaN never replaces the running best
  }
  // warp argmax with move_from_act's rule: the sequential scan "if (v >= best) take" over
  // a = 0..A-1 (a NaN in slot 0 is never replaced)
  for (int d = 16; d > 0; d >>= 1) {
    const float ov = __shfl_xor_sync(full, bv, d);
    const int oa = __shfl_xor_sync(full, ba, d);
    const bool take = oa >= 0 && (ba < 0 || (oa > ba ? ov >= bv : ov > bv));
    if (take) bv = ov, ba = oa;
  }
  first_nan = __any_sync(full, first_nan);
  if (lane != 0) return;
  if (first_nan || ba < 0) ba = 0, bv = __int_as_float(0x7fc00000);
  const bool go = (double)bv > (double)iou_thres;  // label of the best action is +1
  float o0 = x0, o1 = x1, o2 = x2, o3 = x3;
  if (go) {
    const float a0 = __ldg(act + ba * 4 + 0), a1 = __ldg(act + ba * 4 + 1);
    const float a2 = __ldg(act + ba * 4 + 2), a3 = __ldg(act + ba * 4 + 3);
    const float nx = __fadd_rn(x0, __fmul_rn(a0, w)), ny = __fadd_rn(x1, __fmul_rn(a1, h));
    const float nw = __fadd_rn(w, __fmul_rn(a2, w)), nh = __fadd_rn(h, __fmul_rn(a3, h));
    o0 = nx, o1 = ny, o2 = __fsub_rn(__fadd_rn(nx, nw), 1.f), o3 = __fsub_rn(__fadd_rn(ny, nh), 1.f);
    if (moved) atomicAdd(moved, 1);
  }
  const float bidx = __ldg(roi);
  if (refined) {
    float *q = refined + bn * 5;
    q[0] = bidx, q[1] = o0, q[2] = o1, q[3] = o2, q[4] = o3;
  }
  if (packed) {
    float *q = packed + bn * (size_t)(5 + A);
    q[0] = __fadd_rn(bidx, first_image), q[1] = o0, q[2] = o1, q[3] = o2, q[4] = o3;
  }
}

}  // namespace rlod

using namespace rlod;

RLOD_API int rlod_action_reward(const void *boxes, const void *gt, int f64_boxes,
                                const unsigned char *crowd, const int *ngt, const float *act, int B,
                                int N, int A, int G, int mode, int wtrans, float iou_thres,
                                float pos_wratio, float neg_wratio, float *reward, float *label,
                                float *weight, rlod_stream_t stream) {
  if (B < 0 || N < 0 || A < 0 || G < 0) return RLOD_EINVAL;
  if (mode != RLOD_IOU_COCO && mode != RLOD_IOU_RCNN) return RLOD_EINVAL;
  if (wtrans < RLOD_WTRANS_IDENTITY || wtrans > RLOD_WTRANS_RAW) return RLOD_EINVAL;
  if (f64_boxes && mode != RLOD_IOU_COCO) return RLOD_EINVAL;
  if (B == 0 || N == 0 || A == 0) return RLOD_OK;
  if (!boxes || !act || !reward) return RLOD_EINVAL;
  if (G > 0 && !gt) return RLOD_EINVAL;
  const long long total = (long long)B * N * A;
  const long long blocks = cdiv(total, 256);
  const unsigned grid = (unsigned)(blocks < (1LL << 30) ? blocks : (1LL << 30));
  cudaStream_t st = (cudaStream_t)stream;
  if (f64_boxes)
    RLOD_LAUNCH(RLOD_KERNEL_REWARD, st, k_action_reward<double><<<grid, 256, 0, st>>>(
        (const double *)boxes, (const double *)gt, crowd, ngt, act, B, N, A, G, mode, wtrans, iou_thres,
        pos_wratio, neg_wratio, reward, label, weight));
  else
    RLOD_LAUNCH(RLOD_KERNEL_REWARD, st, k_action_reward<float><<<grid, 256, 0, st>>>(
        (const float *)boxes, (const float *)gt, crowd, ngt, act, B, N, A, G, mode, wtrans, iou_thres,
        pos_wratio, neg_wratio, reward, label, weight));
  return launch_status();
}

RLOD_API int rlod_move_from_act(float *boxes, int box_stride, int corners, const float *preds,
                                const float *targets, const float *act, int B, int N, int A,
                                int maxk, int *correct, rlod_stream_t stream) {
  if (B < 0 || N < 0 || A < 1 || maxk < 0 || box_stride < 4) return RLOD_EINVAL;
  if (B == 0 || N == 0 || maxk == 0) return RLOD_OK;
  if (!boxes || !preds || !targets || !act) return RLOD_EINVAL;
  if (N > kMoveMaxN || (long long)N * A >= (1LL << 31)) return RLOD_EUNSUPPORTED;
  int np2 = 2;
  while (np2 < N) np2 <<= 1;
  RLOD_LAUNCH(RLOD_KERNEL_MOVE, (cudaStream_t)stream,
              k_move_from_act<<<B, kMoveThreads, (size_t)np2 * sizeof(unsigned long long),
                                (cudaStream_t)stream>>>(boxes, box_stride, corners, preds, targets, act, N, A,
                                                        maxk, np2, correct));
  return launch_status();
}

RLOD_API int rlod_rl_labels(const void *dets, int det_stride, int f64_boxes, const int *det_cat,
                            const int *ndet, const void *gt, const int *gt_cat,
                            const unsigned char *crowd, const int *ngt, const float *act, int B, int N,
                            int A, int G, int wtrans, float iou_thres, float pos_wratio,
                            float neg_wratio, float *labels, rlod_stream_t stream) {
  if (B < 0 || N < 0 || A < 0 || G < 0 || det_stride < 4) return RLOD_EINVAL;
  if (wtrans < RLOD_WTRANS_IDENTITY || wtrans > RLOD_WTRANS_RAW) return RLOD_EINVAL;
  if (B == 0 || N == 0 || A == 0) return RLOD_OK;
  if (!dets || !act || !labels || (G > 0 && !gt)) return RLOD_EINVAL;
  const long long total = (long long)B * N * A;
  const long long blocks = (total + 255) / 256;
  const unsigned grid = (unsigned)(blocks < (1LL << 20) ? blocks : (1LL << 20));
  cudaStream_t st = (cudaStream_t)stream;
  if (f64_boxes)
    RLOD_LAUNCH(RLOD_KERNEL_REWARD, st, k_rl_labels<double><<<grid, 256, 0, st>>>(
        (const double *)dets, det_stride, det_cat, ndet, (const double *)gt, gt_cat, crowd, ngt, act, B, N, A, G,
        wtrans, iou_thres, pos_wratio, neg_wratio, labels));
  else
    RLOD_LAUNCH(RLOD_KERNEL_REWARD, st, k_rl_labels<float><<<grid, 256, 0, st>>>(
        (const float *)dets, det_stride, det_cat, ndet, (const float *)gt, gt_cat, crowd, ngt, act, B, N, A, G,
        wtrans, iou_thres, pos_wratio, neg_wratio, labels));
  return launch_status();
}

RLOD_API int rlod_reward_refine(const float *rois, const float *gt, const int *ngt, const float *act,
                                int B, int N, int A, int G, int wtrans, float iou_thres,
                                float pos_wratio, float neg_wratio, int first_image, float *reward,
                                float *label, float *weight, float *refined, float *packed,
                                int *moved, rlod_stream_t stream) {
  if (B < 0 || N < 0 || A < 1 || G < 0) return RLOD_EINVAL;
  if (wtrans < RLOD_WTRANS_IDENTITY || wtrans > RLOD_WTRANS_RAW) return RLOD_EINVAL;
  if (B == 0 || N == 0) return RLOD_OK;
  if (!rois || !act || (G > 0 && !gt)) return RLOD_EINVAL;
  if (B > 65535 || (size_t)G * 16 > 40000) return RLOD_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)cdiv(N, kRefineWarps), (unsigned)B);
  RLOD_LAUNCH(RLOD_KERNEL_REWARD, st,
              k_reward_refine<<<grid, kRefineWarps * 32, (size_t)G * 16, st>>>(
                  rois, gt, ngt, act, N, A, G, wtrans, iou_thres, pos_wratio, neg_wratio,
                  (float)first_image, reward, label, weight, refined, packed, moved));
  return launch_status();
}
