// RPN proposal layer for sm_100a: _ProposalLayer.forward (lib/model/rpn/proposal_layer.py:
// 49-161) as three launches for the whole batch and zero host synchronisation.
//
//   k_proposal_sort_decode  one CTA (1024 threads) per image:
//       1. select the pre_nms_topN best anchors on the 64-bit composite key
//          (descending-orderable score bits << 32 | anchor index) straight from the NCHW
//          score map, by bisection on the key value over the keys held in shared memory
//          -- the reference's two permute().contiguous() copies (proposal_layer.py:98,102)
//          and its full sort of all H*W*A scores (:125) vanish;
//       2. compact the selected keys and bitonic-sort them in shared memory, three network
//          levels per pass (all composite keys are distinct, so the order is exactly "score
//          descending, ties by lower anchor index" = a stable descending sort);
//   k_proposal_decode  one thread per selected anchor of the whole batch:
//       3. decode (bbox_transform_inv, bbox_transform.py:77-103) + clip (clip_boxes,
//          :125-133) only the selected anchors, anchors generated arithmetically from the
//          (A,4) base table (:80-93), deltas gathered from channel 4a+k; every fp32 op rounds
//          separately like the eager torch expression it replaces.
//   k_nms_mask + k_nms_scan (nms.cu): batched NMS; the scan stops at post_nms_topN keeps
//       and writes the zero-padded (B, post_nms_topN, 5) output itself (:151-159).
#include "nms_device.cuh"

namespace rlod {

constexpr int kSortThreads = 1024;
constexpr int kSortMax = 16384;  // composite keys held in shared memory (128 KB)

// float -> uint32 whose ASCENDING order is the float's DESCENDING order (NaN first, like
// torch.sort(descending=True); -0 == +0)
__device__ __forceinline__ uint32_t desc_key(float f) {
  uint32_t u = __float_as_uint(f);
  if (f == 0.f) u = 0u;
  const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~asc;
}

struct PropArgs {
  const float *scores, *deltas, *im_info, *anchors;
  int B, A, H, W, feat_stride, pre;  // pre = min(pre_nms_topN, H*W*A), > 0
  float4 *props;                     // (B, pre) decoded + clipped, sorted
  unsigned long long *sel;           // (B, mp) scratch: the selected composite keys, unordered
  int keys_in_smem;                  // all H*W*A score keys fit in shared memory
  int *order_out;                    // (B, pre) or NULL
  float *props_out;                  // (B, pre, 4) or NULL
};

// 0xffffffff when a <= b, else 0: one instruction, so `count -= le_mask(...)` is two
__device__ __forceinline__ unsigned le_mask(uint32_t a, uint32_t b) {
  unsigned r;
  asm("set.le.u32.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// compare-exchange of a bitonic network: ascending when up
__device__ __forceinline__ void cmpex(unsigned long long &x, unsigned long long &y, bool up) {
  if ((x > y) == up) {
    const unsigned long long t = x;
    x = y;
    y = t;
  }
}

// NL consecutive levels (strides jl << (NL-1) ... jl) of the bitonic merge of size k in ONE pass
// over shared memory: an item is 2^NL keys spaced jl apart, exchanged in registers.
// sort-buffer index: one pad word after every 8 keys, so that the 8-key items of the small
// strides do not all start in the same banks
__device__ __forceinline__ int sidx(int i) { return i + (i >> 3); }

template <int NL>
__device__ __forceinline__ void bitonic_pass(unsigned long long *keys, int mp, int k, int jl, int t) {
  constexpr int Q = 1 << NL;
  const int sh = 31 - __clz(jl);
  for (int p = t; p < (mp >> NL); p += kSortThreads) {
    const int base = ((p >> sh) << (sh + NL)) | (p & (jl - 1));
    const bool up = (base & k) == 0;
    unsigned long long x[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) x[q] = keys[sidx(base + q * jl)];
#pragma unroll
    for (int s = Q >> 1; s >= 1; s >>= 1)
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if ((q & s) == 0) cmpex(x[q], x[q | s], up);
#pragma unroll
    for (int q = 0; q < Q; ++q) keys[sidx(base + q * jl)] = x[q];
  }
}

__global__ void __launch_bounds__(kSortThreads) k_proposal_sort_decode(PropArgs a, int mp) {
  // dynamic smem, two lives: first the 32-bit score keys of ALL anchors (select phase, when
  // they fit), then the mp selected 64-bit composite keys (sort phase)
  extern __shared__ __align__(16) unsigned long long keys[];  // [mp], mp = pow2 >= pre
  uint32_t *skeys = reinterpret_cast<uint32_t *>(keys);       // [KA], indexed by anchor index
  __shared__ unsigned int s_cnt[3];
  __shared__ unsigned int s_count;
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31;
  const int HW = a.H * a.W, KA = HW * a.A, M = a.pre;
  const float *fg = a.scores + ((size_t)b * 2 * a.A + a.A) * HW;

  // One coalesced pass over the NCHW fg score map; keys land at their ANCHOR index
  // idx = pix*A + a (memory index e = a*HW + pix, proposal_layer.py:92-103), so every later
  // pass is a plain linear walk of shared memory and the composite key is (key << 32) | idx.
  const bool ks = a.keys_in_smem != 0;
  if (ks) {
#pragma unroll 8
    for (int e = t; e < KA; e += kSortThreads) {
      const int an = e / HW, pix = e - an * HW;
      skeys[pix * a.A + an] = desc_key(__ldg(fg + e));
    }
  }
  if (t < 3) s_cnt[t] = 0u;
  __syncthreads();
  auto key_at = [&](int i) -> uint32_t {
    if (ks) return skeys[i];
    const int pix = i / a.A, an = i - pix * a.A;
    return desc_key(__ldg(fg + (size_t)an * HW + pix));
  };

  // Selection threshold by bisection on the key value: count(key <= mid) is one compare and
  // add per key and one shared atomic per warp -- no histogram, no digit extraction.
  // T = the M-th smallest key; ties at T are cut by a second bisection on the anchor index,
  // so the selected set is exactly "the M smallest composite keys".
  uint32_t T = 0xffffffffu, I = 0xffffffffu;
  if (M < KA) {
    int pass = 0;
    auto finish_count = [&](unsigned c) -> unsigned {
      c = __reduce_add_sync(0xffffffffu, c);
      const int slot = pass % 3;
      if (lane == 0 && c) atomicAdd(&s_cnt[slot], c);
      if (t == 0) s_cnt[(pass + 1) % 3] = 0u;  // last read two passes ago
      __syncthreads();
      ++pass;
      return s_cnt[slot];
    };
    auto count_if = [&](auto pred) -> unsigned {
      unsigned c = 0;
      for (int e = t; e < KA; e += kSortThreads) c += pred(e) ? 1u : 0u;
      return finish_count(c);
    };
    // the hot one: count(key <= mid), four keys per 128-bit load, independent partial counts
    auto count_le = [&](uint32_t mid) -> unsigned {
      if (!ks) return count_if([&](int e) { return key_at(e) <= mid; });
      unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      const uint4 *k4 = reinterpret_cast<const uint4 *>(skeys);
      const int n4 = KA >> 2;
#pragma unroll 4
      for (int q = t; q < n4; q += kSortThreads) {
        const uint4 v = k4[q];
        c0 -= le_mask(v.x, mid), c1 -= le_mask(v.y, mid), c2 -= le_mask(v.z, mid), c3 -= le_mask(v.w, mid);
      }
      for (int e = (n4 << 2) + t; e < KA; e += kSortThreads) c0 += skeys[e] <= mid;
      return finish_count((c0 + c1) + (c2 + c3));
    };
    uint32_t lo = 0u, hi = 0xffffffffu;
    while (lo < hi) {  // CTA-uniform
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (count_le(mid) >= (unsigned)M) hi = mid;
      else lo = mid + 1u;
    }
    T = lo;
    const unsigned c_le = count_le(T);
    if (c_le > (unsigned)M) {
      const unsigned c_lt = T ? count_if([&](int e) { return key_at(e) < T; }) : 0u;
      const unsigned need = (unsigned)M - c_lt;
      uint32_t l2 = 0u, h2 = (uint32_t)KA - 1u;
      while (l2 < h2) {
        const uint32_t mid = l2 + ((h2 - l2) >> 1);
        if (count_if([&](int e) { return key_at(e) == T && (uint32_t)e <= mid; }) >= need) h2 = mid;
        else l2 = mid + 1u;
      }
      I = l2;
    }
  }

  // compaction (unordered; the sort below orders the distinct composite keys) through a
  // global scratch row, because the sort buffer reuses the shared memory the keys live in
  unsigned long long *sel = a.sel + (size_t)b * mp;
  if (t == 0) s_count = 0u;
  __syncthreads();
  for (int e0 = 0; e0 < KA; e0 += kSortThreads) {
    const int e = e0 + t;
    bool take = false;
    uint32_t kv = 0u;
    if (e < KA) {
      kv = key_at(e);
      take = kv < T || (kv == T && (uint32_t)e <= I);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    unsigned wbase = 0;
    if (lane == 0 && bal) wbase = atomicAdd(&s_count, (unsigned)__popc(bal));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (take) {
      const unsigned pos = wbase + __popc(bal & ((1u << lane) - 1u));
      if (pos < (unsigned)mp) sel[pos] = ((unsigned long long)kv << 32) | (unsigned)e;
    }
  }
  __syncthreads();  // also orders this CTA's global writes before its own reads below
  {
    const unsigned cnt = s_count;
    for (int i = t; i < mp; i += kSortThreads) keys[sidx(i)] = (unsigned)i < cnt ? sel[i] : ~0ull;
  }
  __syncthreads();

  // bitonic sort, ascending, mp a power of two; three levels of the network per pass over
  // shared memory (8 keys per thread exchanged in registers): 35 passes instead of 91 at
  // mp = 8192
  for (int k = 2; k <= mp; k <<= 1) {
    int j = k >> 1;
    while (j >= 1) {
      if (j >= 4) {
        bitonic_pass<3>(keys, mp, k, j >> 2, t);
        j >>= 3;
      } else if (j == 2) {
        bitonic_pass<2>(keys, mp, k, 1, t);
        j = 0;
      } else {
        bitonic_pass<1>(keys, mp, k, 1, t);
        j = 0;
      }
      __syncthreads();
    }
  }

  // the sorted keys go back to the global scratch row: decoding is embarrassingly parallel
  // and runs as its own launch over ALL images' boxes (this kernel has one CTA per image)
  for (int i = t; i < M; i += kSortThreads) sel[i] = keys[sidx(i)];
}

// decode + clip the selected anchors in sorted order: one thread per (image, rank)
__global__ void __launch_bounds__(256) k_proposal_decode(PropArgs a, int mp) {
  const int M = a.pre, HW = a.H * a.W;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)a.B * M) return;
  const int b = (int)(gid / M), i = (int)(gid - (long long)b * M);
  const float imh = a.im_info[b * 3 + 0], imw = a.im_info[b * 3 + 1];
  const float xmax = __fsub_rn(imw, 1.f), ymax = __fsub_rn(imh, 1.f);
  const float *dl = a.deltas + (size_t)b * 4 * a.A * HW;
  {
    const int idx = (int)(unsigned)(a.sel[(size_t)b * mp + i] & 0xffffffffull);
    const int an = idx % a.A, pix = idx / a.A;
    const int y = pix / a.W, x = pix - y * a.W;
    const float sx = (float)(x * a.feat_stride), sy = (float)(y * a.feat_stride);
    const float4 base = __ldg(reinterpret_cast<const float4 *>(a.anchors) + an);
    const float bx1 = __fadd_rn(base.x, sx), by1 = __fadd_rn(base.y, sy);
    const float bx2 = __fadd_rn(base.z, sx), by2 = __fadd_rn(base.w, sy);
    const float d0 = __ldg(dl + (size_t)(4 * an + 0) * HW + pix);
    const float d1 = __ldg(dl + (size_t)(4 * an + 1) * HW + pix);
    const float d2 = __ldg(dl + (size_t)(4 * an + 2) * HW + pix);
    const float d3 = __ldg(dl + (size_t)(4 * an + 3) * HW + pix);
    const float w = __fadd_rn(__fsub_rn(bx2, bx1), 1.0f), h = __fadd_rn(__fsub_rn(by2, by1), 1.0f);
    const float cx = __fadd_rn(bx1, __fmul_rn(0.5f, w)), cy = __fadd_rn(by1, __fmul_rn(0.5f, h));
    const float pcx = __fadd_rn(__fmul_rn(d0, w), cx), pcy = __fadd_rn(__fmul_rn(d1, h), cy);
    const float pw = __fmul_rn(expf(d2), w), ph = __fmul_rn(expf(d3), h);
    float4 o;
    o.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
    o.y = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
    o.z = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
    o.w = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
    // clamp_(0, max): min(max(v, 0), max); NaN propagates like torch
    o.x = (o.x != o.x) ? o.x : fminf(fmaxf(o.x, 0.f), xmax);
    o.y = (o.y != o.y) ? o.y : fminf(fmaxf(o.y, 0.f), ymax);
    o.z = (o.z != o.z) ? o.z : fminf(fmaxf(o.z, 0.f), xmax);
    o.w = (o.w != o.w) ? o.w : fminf(fmaxf(o.w, 0.f), ymax);
    a.props[(size_t)b * M + i] = o;
    if (a.order_out) a.order_out[(size_t)b * M + i] = idx;
    if (a.props_out) reinterpret_cast<float4 *>(a.props_out)[(size_t)b * M + i] = o;
  }
}

struct PropWs {
  float4 *props;
  unsigned long long *sel;
  void *mask;
  size_t mask_bytes, bytes;
};

static int pow2_at_least(int v) {
  int mp = 64;
  while (mp < v) mp <<= 1;
  return mp;
}

static PropWs carve_prop_ws(void *base, int B, int pre, int post) {
  PropWs w;
  char *p = (char *)base;
  size_t off = 0;
  w.props = (float4 *)(p ? p + off : nullptr);
  off += align_up((size_t)B * pre * sizeof(float4), 256);
  w.sel = (unsigned long long *)(p ? p + off : nullptr);
  off += align_up((size_t)B * pow2_at_least(pre) * sizeof(unsigned long long), 256);
  w.mask = p ? p + off : nullptr;
  w.mask_bytes = nms_mask_bytes(B, pre, post);
  off += align_up(w.mask_bytes, 256);
  w.bytes = off;
  return w;
}

static int eff_pre(int A, int H, int W, int pre_nms_topN) {
  const long long KA = (long long)A * H * W;
  if (pre_nms_topN > 0 && pre_nms_topN < KA) return pre_nms_topN;
  return (int)(KA < (1LL << 30) ? KA : (1LL << 30));
}

}  // namespace rlod

using namespace rlod;

RLOD_API size_t rlod_proposal_workspace_bytes(int B, int A, int H, int W, int pre_nms_topN,
                                              int post_nms_topN) {
  if (B <= 0 || A <= 0 || H <= 0 || W <= 0) return 0;
  return carve_prop_ws(nullptr, B, eff_pre(A, H, W, pre_nms_topN), post_nms_topN).bytes;
}

RLOD_API int rlod_proposal_forward(const float *scores, const float *deltas, const float *im_info,
                                   const float *anchors, int B, int A, int H, int W,
                                   int feat_stride, int pre_nms_topN, int post_nms_topN,
                                   float nms_thresh, float *rois_out, int *order_out,
                                   float *props_out, int *nkeep_out, void *workspace,
                                   size_t workspace_bytes, rlod_stream_t stream) {
  if (B < 0 || A <= 0 || H <= 0 || W <= 0 || post_nms_topN <= 0) return RLOD_EINVAL;
  if (B == 0) return RLOD_OK;
  if (!scores || !deltas || !im_info || !anchors || !rois_out || !workspace) return RLOD_EINVAL;
  if (((uintptr_t)anchors % 16) != 0) return RLOD_EINVAL;
  if (props_out && ((uintptr_t)props_out % 16) != 0) return RLOD_EINVAL;
  if ((long long)A * H * W >= (1LL << 30)) return RLOD_EUNSUPPORTED;
  const int pre = eff_pre(A, H, W, pre_nms_topN);
  if (pre > kSortMax) return RLOD_EUNSUPPORTED;
  PropWs ws = carve_prop_ws(workspace, B, pre, post_nms_topN);
  if (workspace_bytes < ws.bytes) return RLOD_ENOSPC;
  cudaStream_t st = (cudaStream_t)stream;

  const int mp = pow2_at_least(pre);
  const size_t key_bytes = (size_t)A * H * W * sizeof(uint32_t);
  PropArgs pa;
  pa.scores = scores, pa.deltas = deltas, pa.im_info = im_info, pa.anchors = anchors;
  pa.B = B, pa.A = A, pa.H = H, pa.W = W, pa.feat_stride = feat_stride, pa.pre = pre;
  pa.props = ws.props, pa.order_out = order_out, pa.props_out = props_out;
  pa.sel = ws.sel;
  pa.keys_in_smem = key_bytes <= (size_t)(kMaxSmemPerCta - 1024) ? 1 : 0;
  size_t smem = (size_t)(mp + (mp >> 3)) * sizeof(unsigned long long);
  if (pa.keys_in_smem && key_bytes > smem) smem = align_up(key_bytes, 16);
  cudaFuncSetAttribute(k_proposal_sort_decode, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)smem);
  RLOD_LAUNCH(RLOD_KERNEL_PROPOSAL_SORT, st, k_proposal_sort_decode<<<B, kSortThreads, smem, st>>>(pa, mp));
  RLOD_LAUNCH(RLOD_KERNEL_PROPOSAL_SORT, st,
              k_proposal_decode<<<(unsigned)cdiv((long long)B * pre, 256), 256, 0, st>>>(pa, mp));
  int rc = launch_status();
  if (rc) return rc;

  NmsSegs segs;
  segs.dets = reinterpret_cast<const float *>(ws.props);
  segs.stride = 4;
  segs.vec4 = 1;
  segs.seg_offsets = nullptr;
  segs.uniform_n = pre;
  segs.max_n = pre;
  NmsOut o = {};
  o.keep = nullptr;
  o.num_out = nkeep_out;
  o.rois = rois_out;
  o.post = post_nms_topN;
  // post_nms_topN <= 512: kept-list walk (k_nms_lazy); larger: tiled mask + device scan
  return nms_launch(segs, B, pre, nms_thresh, post_nms_topN, o, ws.mask, ws.mask_bytes, st, 0);
}
