// Training-target layers for sm_100a: the two layers that sit between the proposal
// layer and RoI pooling when the detector trains.
//
//   k_proposal_target   _ProposalTargetLayer.forward (lib/model/rpn/proposal_target_layer_cascade.py
//                       :33-213), one CTA per image: candidates = rois ++ gt boxes (:41-44),
//                       bbox_overlaps_batch incl. its 0 / -1 sentinels (bbox_transform.py:168-257),
//                       max / first argmax over gt (:134), fg = max >= FG_THRESH, bg = LO <= max < HI
//                       (:153-158), sampling (:160-199), labels / rois / bbox_transform_batch targets
//                       normalised by mean / std (:104-121), class-agnostic 4-wide targets and
//                       inside / outside weights (:70-102, 57).
//   k_anchor_target     _AnchorTargetLayer.forward (lib/model/rpn/anchor_target_layer.py:48-192), one
//                       CTA per image: inside-image anchors (:84-91), overlaps, per-anchor max /
//                       argmax and per-gt max (:101-102), label rules (:104-119), fg / bg
//                       subsampling (:126-147), targets (:151), weights (:154-166; the reference
//                       takes the example count of the LAST image for all images -- replicated),
//                       and the (B,1,A*H,W) / (B,4A,H,W) output layouts (:175-192).
//
// Randomness.  The reference draws np.random.permutation / np.random.rand inside the layer, with
// sizes that are only known after a device->host read.  Here the caller passes the random numbers
// as tensors and the rules are deterministic functions of them:
//   "random k of a set"  = the k members with the smallest key (ties: lower index), which is what
//                          fg_inds[np.random.permutation(n)[:k]] does when the permutation is the
//                          argsort of the keys;
//   "k draws with replacement" = set[floor(u_j * n)] with the caller's u_j, the reference's formula.
// tests/golden/make_golden_targets.py runs the reference's own layers with np.random patched to
// exactly these rules, which pins everything else bit for bit.
#include "rlod_common.cuh"

namespace rlod {

constexpr int kTgtThreads = 256;

__device__ __forceinline__ float tgt_overlap(float a0, float a1, float a2, float a3, float g0, float g1,
                                             float g2, float g3) {
  const float ax = __fadd_rn(__fsub_rn(a2, a0), 1.f), ay = __fadd_rn(__fsub_rn(a3, a1), 1.f);
  const float gx = __fadd_rn(__fsub_rn(g2, g0), 1.f), gy = __fadd_rn(__fsub_rn(g3, g1), 1.f);
  float iw = __fadd_rn(__fsub_rn(fminf(a2, g2), fmaxf(a0, g0)), 1.f);
  if (iw < 0.f) iw = 0.f;
  float ih = __fadd_rn(__fsub_rn(fminf(a3, g3), fmaxf(a1, g1)), 1.f);
  if (ih < 0.f) ih = 0.f;
  const float inter = __fmul_rn(iw, ih);
  float v = __fdiv_rn(inter, __fsub_rn(__fadd_rn(__fmul_rn(ax, ay), __fmul_rn(gx, gy)), inter));
  if (gx == 1.f && gy == 1.f) v = 0.f;   // zero-padded gt (:195-196)
  if (ax == 1.f && ay == 1.f) v = -1.f;  // degenerate anchor / roi (:212-213)
  return v;
}

// bbox_transform_batch (bbox_transform.py:44-75), one box
__device__ __forceinline__ float4 tgt_transform(float e0, float e1, float e2, float e3, float g0, float g1,
                                                float g2, float g3) {
  const float ew = __fadd_rn(__fsub_rn(e2, e0), 1.f), eh = __fadd_rn(__fsub_rn(e3, e1), 1.f);
  const float ecx = __fadd_rn(e0, __fmul_rn(0.5f, ew)), ecy = __fadd_rn(e1, __fmul_rn(0.5f, eh));
  const float gw = __fadd_rn(__fsub_rn(g2, g0), 1.f), gh = __fadd_rn(__fsub_rn(g3, g1), 1.f);
  const float gcx = __fadd_rn(g0, __fmul_rn(0.5f, gw)), gcy = __fadd_rn(g1, __fmul_rn(0.5f, gh));
  return make_float4(__fdiv_rn(__fsub_rn(gcx, ecx), ew), __fdiv_rn(__fsub_rn(gcy, ecy), eh),
                     logf(__fdiv_rn(gw, ew)), logf(__fdiv_rn(gh, eh)));
}

// ordered compaction of a predicate over i = 0..n-1 into list (ascending i); returns the count.
// All threads of the CTA call it; s_scan is kTgtThreads / 32 + 1 ints of shared scratch.
template <class Pred>
__device__ int tgt_compact(int n, Pred pred, int *list, int *s_scan) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int total = 0;
  for (int base = 0; base < n; base += kTgtThreads) {
    const int i = base + t;
    const bool p = i < n && pred(i);
    const unsigned m = __ballot_sync(0xffffffffu, p);
    if (lane == 0) s_scan[warp] = __popc(m);
    __syncthreads();
    int before = 0, chunk = 0;
    for (int w = 0; w < kTgtThreads / 32; ++w) {
      const int c = s_scan[w];
      if (w < warp) before += c;
      chunk += c;
    }
    if (p) list[total + before + __popc(m & ((1u << lane) - 1u))] = i;
    total += chunk;
    __syncthreads();
  }
  return total;
}

struct PtArgs {
  const float *rois;     // (B, N, 5)
  const float *gt;       // (B, G, 5) [x1,y1,x2,y2,cls], zero rows = padding
  const float *fg_keys;  // (B, N + G)
  const float *bg_u;     // (B, R)
  int B, N, G, R, fg_per_image;
  float fg_thresh, bg_hi, bg_lo;
  float mean[4], std[4], inside[4];
  int normalize;
  float *rois_out, *labels_out, *targets_out, *inside_out, *outside_out;  // (B,R,5) (B,R) (B,R,4) x3
  int *status;  // (B): 1 = neither fg nor bg candidates (the reference raises ValueError, :196-197)
};

__global__ void __launch_bounds__(kTgtThreads) k_proposal_target(PtArgs a, int mp) {
  extern __shared__ __align__(16) unsigned char pt_raw[];
  const int M = a.N + a.G;
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(pt_raw);  // [mp]
  float *maxov = reinterpret_cast<float *>(keys + mp);                        // [M]
  int *asg = reinterpret_cast<int *>(maxov + M);                              // [M]
  int *fg = asg + M, *bg = fg + M;                                            // [M] each
  __shared__ int s_scan[kTgtThreads / 32 + 1];
  const int b = blockIdx.x, t = threadIdx.x;
  const float *gtb = a.gt + (size_t)b * a.G * 5;
  auto cand = [&](int i, float &x1, float &y1, float &x2, float &y2) {
    if (i < a.N) {
      const float *r = a.rois + ((size_t)b * a.N + i) * 5;
      x1 = __ldg(r + 1), y1 = __ldg(r + 2), x2 = __ldg(r + 3), y2 = __ldg(r + 4);
    } else {  // gt_boxes_append[:, :, 1:5] = gt_boxes[:, :, :4] (:41-42)
      const float *g = gtb + (size_t)(i - a.N) * 5;
      x1 = __ldg(g), y1 = __ldg(g + 1), x2 = __ldg(g + 2), y2 = __ldg(g + 3);
    }
  };
  for (int i = t; i < M; i += kTgtThreads) {
    float x1, y1, x2, y2;
    cand(i, x1, y1, x2, y2);
    float best = -INFINITY;
    int bi = 0;
    for (int g = 0; g < a.G; ++g) {
      const float *q = gtb + (size_t)g * 5;
      const float v = tgt_overlap(x1, y1, x2, y2, __ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
      if (v > best) best = v, bi = g;  // first maximum
    }
    maxov[i] = best;
    asg[i] = bi;
  }
  __syncthreads();
  const int nfg = tgt_compact(M, [&](int i) { return maxov[i] >= a.fg_thresh; }, fg, s_scan);
  const int nbg = tgt_compact(M, [&](int i) { return maxov[i] < a.bg_hi && maxov[i] >= a.bg_lo; }, bg, s_scan);
  const int R = a.R;
  int fg_this = 0;
  if (nfg > 0 && nbg > 0) {
    fg_this = min(a.fg_per_image, nfg);
    // the fg_this candidates with the smallest keys, in key order: sort (key, position in fg)
    for (int i = t; i < mp; i += kTgtThreads) {
      unsigned long long k = ~0ull;
      if (i < nfg) {
        const float kf = __ldg(a.fg_keys + (size_t)b * M + fg[i]);
        uint32_t u = __float_as_uint(kf);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-orderable
        k = ((unsigned long long)u << 32) | (unsigned)i;
      }
      keys[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= mp; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int p = t; p < (mp >> 1); p += kTgtThreads) {
          const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1)), l = i + j;
          const unsigned long long x = keys[i], y = keys[l];
          if ((x > y) == ((i & k) == 0)) keys[i] = y, keys[l] = x;
        }
        __syncthreads();
      }
  } else if (nfg > 0) {
    fg_this = R;
  } else if (nbg == 0) {
    if (t == 0 && a.status) a.status[b] = 1;
  }
  __syncthreads();
  // one thread per output row
  for (int j = t; j < R; j += kTgtThreads) {
    int ci = -1;  // candidate index
    if (nfg > 0 && nbg > 0) {
      if (j < fg_this) {
        ci = fg[(int)(unsigned)(keys[j] & 0xffffffffull)];
      } else {
        const double u = (double)__ldg(a.bg_u + (size_t)b * R + (j - fg_this));
        ci = bg[min(nbg - 1, (int)floor(u * (double)nbg))];
      }
    } else if (nfg > 0) {
      const double u = (double)__ldg(a.bg_u + (size_t)b * R + j);
      ci = fg[min(nfg - 1, (int)floor(u * (double)nfg))];
    } else if (nbg > 0) {
      const double u = (double)__ldg(a.bg_u + (size_t)b * R + j);
      ci = bg[min(nbg - 1, (int)floor(u * (double)nbg))];
    }
    float *ro = a.rois_out + ((size_t)b * R + j) * 5;
    float4 tg = make_float4(0.f, 0.f, 0.f, 0.f);
    float label = 0.f;
    ro[0] = (float)b;
    if (ci < 0) {
      ro[1] = ro[2] = ro[3] = ro[4] = 0.f;
    } else {
      float x1, y1, x2, y2;
      cand(ci, x1, y1, x2, y2);
      ro[1] = x1, ro[2] = y1, ro[3] = x2, ro[4] = y2;
      const float *g = gtb + (size_t)asg[ci] * 5;
      label = j < fg_this ? __ldg(g + 4) : 0.f;  // bg rows are clamped to 0 (:207-208)
      tg = tgt_transform(x1, y1, x2, y2, __ldg(g), __ldg(g + 1), __ldg(g + 2), __ldg(g + 3));
      if (a.normalize) {
        tg.x = __fdiv_rn(__fsub_rn(tg.x, a.mean[0]), a.std[0]);
        tg.y = __fdiv_rn(__fsub_rn(tg.y, a.mean[1]), a.std[1]);
        tg.z = __fdiv_rn(__fsub_rn(tg.z, a.mean[2]), a.std[2]);
        tg.w = __fdiv_rn(__fsub_rn(tg.w, a.mean[3]), a.std[3]);
      }
    }
    a.labels_out[(size_t)b * R + j] = label;
    const bool pos = label > 0.f;
    float *to = a.targets_out + ((size_t)b * R + j) * 4, *io = a.inside_out + ((size_t)b * R + j) * 4;
    float *oo = a.outside_out + ((size_t)b * R + j) * 4;
    to[0] = pos ? tg.x : 0.f, to[1] = pos ? tg.y : 0.f, to[2] = pos ? tg.z : 0.f, to[3] = pos ? tg.w : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float w = pos ? a.inside[q] : 0.f;
      io[q] = w;
      oo[q] = w > 0.f ? 1.f : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// anchor target layer
// ---------------------------------------------------------------------------------------------
struct AtArgs {
  const float *gt;       // (B, G, 5)
  const float *im_info;  // (B, 3): image 0's h, w bound the inside test for the whole batch (:84-87)
  const float *anchors;  // (A, 4)
  const float *keys;     // (B, K*A) random keys
  int B, G, A, H, W, feat_stride;
  float pos_ov, neg_ov, inside_w, pos_weight;
  int clobber, num_fg, batchsize;
  float *labels;   // (B, 1, A*H, W)
  float *targets;  // (B, 4A, H, W)
  float *inside;   // (B, 4A, H, W)
  float *outside;  // (B, 4A, H, W)
  int *counts;     // (B, 4) scratch: examples (labels >= 0) per image, written by pass 1
  signed char *lab_ws;  // (B, K*A) labels after subsampling (scratch)
  int *arg_ws;          // (B, K*A) argmax gt (scratch)
};

__device__ __forceinline__ uint32_t asc_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// pass 1: labels incl. subsampling, per image.  smem: float gtmax[G]; then reductions in place.
__global__ void __launch_bounds__(kTgtThreads) k_anchor_target_labels(AtArgs a) {
  extern __shared__ __align__(16) unsigned char at_raw[];
  unsigned *gtmax = reinterpret_cast<unsigned *>(at_raw);  // [G] orderable bits of the per-gt max overlap
  __shared__ unsigned s_cnt[3];
  __shared__ int s_fg, s_bg;
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31;
  const int KA = a.H * a.W * a.A;
  const float *gtb = a.gt + (size_t)b * a.G * 5;
  const float imw = floorf(a.im_info[1]), imh = floorf(a.im_info[0]);  // long(im_info[0][1]), long(im_info[0][0])
  signed char *lab = a.lab_ws + (size_t)b * KA;
  int *arg = a.arg_ws + (size_t)b * KA;
  for (int g = t; g < a.G; g += kTgtThreads) gtmax[g] = 0u;  // below every real value's key
  if (t == 0) s_fg = 0, s_bg = 0;
  if (t < 3) s_cnt[t] = 0u;
  __syncthreads();
  auto anchor_of = [&](int i, float &x1, float &y1, float &x2, float &y2) -> bool {
    const int an = i % a.A, pix = i / a.A;
    const int y = pix / a.W, x = pix - y * a.W;
    const float sx = (float)(x * a.feat_stride), sy = (float)(y * a.feat_stride);
    const float4 base = __ldg(reinterpret_cast<const float4 *>(a.anchors) + an);
    x1 = __fadd_rn(base.x, sx), y1 = __fadd_rn(base.y, sy), x2 = __fadd_rn(base.z, sx), y2 = __fadd_rn(base.w, sy);
    return x1 >= 0.f && y1 >= 0.f && x2 < imw && y2 < imh;  // allowed_border = 0 (:84-87)
  };
  // per-gt maximum over the inside anchors (:102)
  for (int i = t; i < KA; i += kTgtThreads) {
    float x1, y1, x2, y2;
    if (!anchor_of(i, x1, y1, x2, y2)) continue;
    for (int g = 0; g < a.G; ++g) {
      const float *q = gtb + (size_t)g * 5;
      const float v = tgt_overlap(x1, y1, x2, y2, __ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
      atomicMax(&gtmax[g], asc_key(v));
    }
  }
  __syncthreads();
  // labels before subsampling (:104-119)
  int nfg = 0, nbg = 0;
  for (int i = t; i < KA; i += kTgtThreads) {
    float x1, y1, x2, y2;
    int l = -2;  // outside the image: not an example at all
    if (anchor_of(i, x1, y1, x2, y2)) {
      float best = -INFINITY;
      int bi = 0;
      bool is_gt_max = false;
      for (int g = 0; g < a.G; ++g) {
        const float *q = gtb + (size_t)g * 5;
        const float v = tgt_overlap(x1, y1, x2, y2, __ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
        if (v > best) best = v, bi = g;
        // gt_max_overlaps[gt_max_overlaps == 0] = 1e-5 (:107): a gt whose best overlap is 0 matches nobody
        float gm;
        {
          const uint32_t k = gtmax[g];
          const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
          gm = __uint_as_float(u);
        }
        if (gm == 0.f) gm = 1e-5f;
        is_gt_max |= (v == gm);
      }
      l = -1;
      if (!a.clobber && best < a.neg_ov) l = 0;
      if (is_gt_max) l = 1;
      if (best >= a.pos_ov) l = 1;
      if (a.clobber && best < a.neg_ov) l = 0;
      arg[i] = bi;
      nfg += l == 1, nbg += l == 0;
    }
    lab[i] = (signed char)l;
  }
  nfg = __reduce_add_sync(0xffffffffu, nfg), nbg = __reduce_add_sync(0xffffffffu, nbg);
  if (lane == 0) atomicAdd(&s_fg, nfg), atomicAdd(&s_bg, nbg);
  __syncthreads();
  const int sum_fg = s_fg, sum_bg = s_bg;
  // subsampling (:126-147): disable the (n - keep) members with the smallest keys
  int pass = 0;
  auto count_le = [&](int which, uint32_t mid, uint32_t tie_idx_max) -> unsigned {
    unsigned c = 0;
    for (int i = t; i < KA; i += kTgtThreads)
      if (lab[i] == which) {
        const uint32_t k = asc_key(__ldg(a.keys + (size_t)b * KA + i));
        c += (k < mid) || (k == mid && (uint32_t)i <= tie_idx_max);
      }
    c = __reduce_add_sync(0xffffffffu, c);
    const int slot = pass % 3;
    if (lane == 0 && c) atomicAdd(&s_cnt[slot], c);
    if (t == 0) s_cnt[(pass + 1) % 3] = 0u;
    __syncthreads();
    ++pass;
    return s_cnt[slot];
  };
  auto disable_smallest = [&](int which, int n_disable) {
    // threshold (key, index) of the n_disable-th smallest member
    uint32_t lo = 0u, hi = 0xffffffffu;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (count_le(which, mid, 0xffffffffu) >= (unsigned)n_disable) hi = mid;
      else lo = mid + 1u;
    }
    const uint32_t T = lo;
    uint32_t l2 = 0u, h2 = (uint32_t)KA - 1u;
    if (count_le(which, T, 0xffffffffu) > (unsigned)n_disable) {
      while (l2 < h2) {
        const uint32_t mid = l2 + ((h2 - l2) >> 1);
        if (count_le(which, T, mid) >= (unsigned)n_disable) h2 = mid;
        else l2 = mid + 1u;
      }
    } else {
      l2 = 0xffffffffu;
    }
    __syncthreads();
    for (int i = t; i < KA; i += kTgtThreads)
      if (lab[i] == which) {
        const uint32_t k = asc_key(__ldg(a.keys + (size_t)b * KA + i));
        if (k < T || (k == T && (uint32_t)i <= l2)) lab[i] = -1;
      }
    __syncthreads();
  };
  if (sum_fg > a.num_fg) disable_smallest(1, sum_fg - a.num_fg);
  const int num_bg = a.batchsize - sum_fg;  // uses sum_fg BEFORE subsampling, like the reference (:138)
  if (sum_bg > num_bg) disable_smallest(0, sum_bg - num_bg);
  // examples of this image (labels >= 0) for the weights of pass 2
  int ex = 0;
  for (int i = t; i < KA; i += kTgtThreads) ex += lab[i] >= 0;
  ex = __reduce_add_sync(0xffffffffu, ex);
  if (t == 0) s_fg = 0;
  __syncthreads();
  if (lane == 0) atomicAdd(&s_fg, ex);
  __syncthreads();
  if (t == 0) a.counts[b] = s_fg;
}

// pass 2: targets, weights and the output layouts; one thread per (image, anchor)
__global__ void __launch_bounds__(256) k_anchor_target_outputs(AtArgs a) {
  const int KA = a.H * a.W * a.A, HW = a.H * a.W;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)a.B * KA) return;
  const int b = (int)(gid / KA), i = (int)(gid - (long long)b * KA);
  const int an = i % a.A, pix = i / a.A;
  const int y = pix / a.W, x = pix - y * a.W;
  const int l = a.lab_ws[gid];
  // labels: (B, H, W, A) -> (B, A, H, W) viewed as (B, 1, A*H, W)
  a.labels[((size_t)b * a.A + an) * HW + pix] = l == -2 ? -1.f : (float)l;  // _unmap fill = -1
  float4 tg = make_float4(0.f, 0.f, 0.f, 0.f);
  float iw = 0.f, ow = 0.f;
  if (l != -2) {
    const float sx = (float)(x * a.feat_stride), sy = (float)(y * a.feat_stride);
    const float4 base = __ldg(reinterpret_cast<const float4 *>(a.anchors) + an);
    const float *g = a.gt + ((size_t)b * a.G + a.arg_ws[gid]) * 5;
    tg = tgt_transform(__fadd_rn(base.x, sx), __fadd_rn(base.y, sy), __fadd_rn(base.z, sx), __fadd_rn(base.w, sy),
                       __ldg(g), __ldg(g + 1), __ldg(g + 2), __ldg(g + 3));
    if (l == 1) iw = a.inside_w;
    // RPN_POSITIVE_WEIGHT < 0: uniform 1 / num_examples, with the count of the LAST image (:157-160)
    float pw, nw;
    if (a.pos_weight < 0.f) {
      pw = nw = __fdiv_rn(1.0f, (float)a.counts[a.B - 1]);
    } else {
      pw = a.pos_weight, nw = 1.f - a.pos_weight;  // not reachable in the reference (it only asserts)
    }
    if (l == 1) ow = pw;
    if (l == 0) ow = nw;
  }
  // (B, H, W, 4A) -> (B, 4A, H, W): channel 4*an + q
  const size_t o = ((size_t)b * 4 * a.A + 4 * an) * HW + pix;
  a.targets[o] = tg.x, a.targets[o + HW] = tg.y, a.targets[o + 2 * (size_t)HW] = tg.z, a.targets[o + 3 * (size_t)HW] = tg.w;
#pragma unroll
  for (int q = 0; q < 4; ++q) a.inside[o + q * (size_t)HW] = iw, a.outside[o + q * (size_t)HW] = ow;
}

}  // namespace rlod

using namespace rlod;

RLOD_API int rlod_proposal_target(const float *rois, const float *gt, const float *fg_keys, const float *bg_u,
                                  int B, int N, int G, int rois_per_image, int fg_rois_per_image,
                                  float fg_thresh, float bg_thresh_hi, float bg_thresh_lo, const float *means,
                                  const float *stds, const float *inside_weights, float *rois_out,
                                  float *labels_out, float *targets_out, float *inside_out, float *outside_out,
                                  int *status, rlod_stream_t stream) {
  if (B < 0 || N < 0 || G < 1 || rois_per_image < 1 || fg_rois_per_image < 0) return RLOD_EINVAL;
  if (B == 0) return RLOD_OK;
  if (!rois || !gt || !fg_keys || !bg_u || !rois_out || !labels_out || !targets_out || !inside_out || !outside_out)
    return RLOD_EINVAL;
  const int M = N + G;
  if (M > 16384) return RLOD_EUNSUPPORTED;
  PtArgs a;
  a.rois = rois, a.gt = gt, a.fg_keys = fg_keys, a.bg_u = bg_u;
  a.B = B, a.N = N, a.G = G, a.R = rois_per_image, a.fg_per_image = fg_rois_per_image;
  a.fg_thresh = fg_thresh, a.bg_hi = bg_thresh_hi, a.bg_lo = bg_thresh_lo;
  a.normalize = (means && stds) ? 1 : 0;
  for (int i = 0; i < 4; ++i) {
    a.mean[i] = means ? means[i] : 0.f, a.std[i] = stds ? stds[i] : 1.f;
    a.inside[i] = inside_weights ? inside_weights[i] : 1.f;
  }
  a.rois_out = rois_out, a.labels_out = labels_out, a.targets_out = targets_out;
  a.inside_out = inside_out, a.outside_out = outside_out, a.status = status;
  cudaStream_t st = (cudaStream_t)stream;
  if (status) cudaMemsetAsync(status, 0, (size_t)B * sizeof(int), st);
  int mp = 64;
  while (mp < M) mp <<= 1;
  const size_t smem = (size_t)mp * 8 + (size_t)M * 16;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_proposal_target, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  RLOD_LAUNCH(RLOD_KERNEL_TARGETS, st, k_proposal_target<<<B, kTgtThreads, smem, st>>>(a, mp));
  return launch_status();
}

RLOD_API size_t rlod_anchor_target_workspace_bytes(int B, int A, int H, int W) {
  if (B <= 0 || A <= 0 || H <= 0 || W <= 0) return 0;
  const size_t KA = (size_t)A * H * W;
  return align_up((size_t)B * 4 * sizeof(int), 256) + align_up((size_t)B * KA, 256) + (size_t)B * KA * sizeof(int);
}

RLOD_API int rlod_anchor_target(const float *gt, const float *im_info, const float *anchors, const float *keys,
                                int B, int G, int A, int H, int W, int feat_stride, float positive_overlap,
                                float negative_overlap, int clobber_positives, float fg_fraction, int batchsize,
                                float inside_weight, float positive_weight, float *labels, float *bbox_targets,
                                float *inside_weights, float *outside_weights, void *workspace,
                                size_t workspace_bytes, rlod_stream_t stream) {
  if (B < 0 || G < 1 || A < 1 || H < 1 || W < 1 || batchsize < 1) return RLOD_EINVAL;
  if (B == 0) return RLOD_OK;
  if (!gt || !im_info || !anchors || !keys || !labels || !bbox_targets || !inside_weights || !outside_weights ||
      !workspace)
    return RLOD_EINVAL;
  if (((uintptr_t)anchors % 16) != 0) return RLOD_EINVAL;
  if (workspace_bytes < rlod_anchor_target_workspace_bytes(B, A, H, W)) return RLOD_ENOSPC;
  if ((long long)A * H * W >= (1LL << 30)) return RLOD_EUNSUPPORTED;
  const size_t KA = (size_t)A * H * W;
  AtArgs a;
  a.gt = gt, a.im_info = im_info, a.anchors = anchors, a.keys = keys;
  a.B = B, a.G = G, a.A = A, a.H = H, a.W = W, a.feat_stride = feat_stride;
  a.pos_ov = positive_overlap, a.neg_ov = negative_overlap, a.inside_w = inside_weight, a.pos_weight = positive_weight;
  a.clobber = clobber_positives, a.num_fg = (int)(fg_fraction * batchsize), a.batchsize = batchsize;
  a.labels = labels, a.targets = bbox_targets, a.inside = inside_weights, a.outside = outside_weights;
  char *p = (char *)workspace;
  a.counts = (int *)p;
  p += align_up((size_t)B * 4 * sizeof(int), 256);
  a.lab_ws = (signed char *)p;
  p += align_up((size_t)B * KA, 256);
  a.arg_ws = (int *)p;
  cudaStream_t st = (cudaStream_t)stream;
  RLOD_LAUNCH(RLOD_KERNEL_TARGETS, st,
              k_anchor_target_labels<<<B, kTgtThreads, (size_t)G * sizeof(unsigned), st>>>(a));
  RLOD_LAUNCH(RLOD_KERNEL_TARGETS, st,
              k_anchor_target_outputs<<<(unsigned)cdiv((long long)B * KA, 256), 256, 0, st>>>(a));
  return launch_status();
}
