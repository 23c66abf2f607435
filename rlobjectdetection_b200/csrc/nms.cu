// Bitmask NMS for sm_100a, fully on device.
//
// Arithmetic is bit-identical to the reference's devIoU + strict '>' + greedy scan
// (lib/model/nms/src/nms_cuda_kernel.cu:31-39, 77-81, 123-144) as nvcc 12.9 compiles it for
// sm_100a: Sa = w_a*h_a (FMUL), S = fma(w_b, h_b, Sa), inter = w*h (FMUL),
// iou = inter / (S - inter) (IEEE division), bit = iou > thresh.
//
// What differs is everything around it:
//   - k_nms_mask computes only the upper-triangular 64x64 tiles, for all segments (images /
//     (image,class) pairs) in one launch, and stores the mask column-block-major
//     ([col_block][row]) so both the tile write (512 B) and the scan's reads are coalesced.
//     A division-free interval test decides all but the razor-edge pairs; those fall through
//     to the exact IEEE division, so the bits never differ from the reference's.
//   - k_nms_scan resolves the serial greedy dependency ON THE DEVICE, one CTA per segment:
//     per 64-box block one thread walks the diagonal tile, then 8 warps OR the kept rows
//     into the running removed-bitmap with warp-wide OR reductions.  It stops as soon as
//     max_keep boxes are kept (the proposal layer only uses the first post_nms_topN) and
//     can emit the padded (B, post, 5) roi tensor directly.
//   - k_nms_small handles short segments (<= 512 boxes, e.g. per-class test-time NMS) in
//     one CTA each with the mask in shared memory: one launch for thousands of segments.
// No cudaMalloc, no mask D2H (18 MB per call at n=12000 in the reference), no host loop,
// no default-stream sync.
#include <cstdlib>

#include "nms_device.cuh"

namespace rlod {

// ----------------------------------------------------------------------------------------
// tile kernel: grid = (n_tiles_upper_triangular(max_blk), nseg), block = 64 threads.
// thread t owns row rb*64+t and tests it against the 64 boxes of column block cb.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
    k_nms_mask(NmsSegs segs, float thresh, int max_blk, unsigned long long *__restrict__ mask,
               size_t mask_seg_stride) {
  const int seg = blockIdx.y;
  int off, n;
  segs.get(seg, off, n);
  const int nblk = (n + 63) >> 6;
  // decode the upper-triangular tile index -> (rb, cb), cb >= rb, over max_blk blocks
  int rb, cb;
  tri_decode(blockIdx.x, max_blk, rb, cb);
  if (cb >= nblk) return;  // also covers rb >= nblk
  const int npad = nblk << 6;
  __shared__ float4 cbox[64];
  __shared__ float2 cwh[64];
  const int t = threadIdx.x;
  {
    const int j = cb * 64 + t;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < n) bx = segs.load_box(off + j);
    cbox[t] = bx;
    cwh[t] = make_float2(__fadd_rn(__fsub_rn(bx.z, bx.x), 1.f), __fadd_rn(__fsub_rn(bx.w, bx.y), 1.f));
  }
  __syncthreads();
  const int i = rb * 64 + t;
  unsigned long long bits = 0ull;
  if (i < n) {
    const float4 a = segs.load_box(off + i);
    const float Sa = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
    const int jn = min(64, n - cb * 64);
    const int j0 = (rb == cb) ? t + 1 : 0;
    const bool fast = thresh >= 0.f && thresh < 1e30f;
    for (int j = j0; j < jn; ++j)
      if (iou_gt(a, Sa, cbox[j], cwh[j], thresh, fast)) bits |= 1ull << j;
  }
  mask[(size_t)seg * mask_seg_stride + (size_t)cb * npad + i] = bits;
}

// ----------------------------------------------------------------------------------------
// scan kernel: one CTA (1024 threads) per segment.
// ----------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
    k_nms_scan(NmsSegs segs, int max_keep, const unsigned long long *__restrict__ mask,
               size_t mask_seg_stride, NmsOut o) {
  extern __shared__ unsigned long long remv[];  // [nblk]
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long s_kept;
  __shared__ int s_base, s_total;
  const int seg = blockIdx.x;
  int off, n;
  segs.get(seg, off, n);
  const int nblk = (n + 63) >> 6, npad = nblk << 6;
  const unsigned long long *m = mask + (size_t)seg * mask_seg_stride;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int w = t; w < nblk; w += kScanThreads) remv[w] = 0ull;
  if (t == 0) s_total = 0;
  __syncthreads();
  unsigned long long dnext = (t < 64 && nblk > 0) ? m[t] : 0ull;  // diagonal tile of block 0
  for (int k = 0; k < nblk; ++k) {
    if (t < 64) {
      diag[t] = dnext;
      // the next diagonal tile travels while this block is resolved and pushed
      if (k + 1 < nblk) dnext = m[(size_t)(k + 1) * npad + (k + 1) * 64 + t];
    }
    __syncthreads();
    if (t == 0) {
      unsigned long long r = remv[k];
      const int valid = min(64, n - k * 64);
      if (valid < 64) r |= ~0ull << valid;
      unsigned long long kb = 0ull;
#pragma unroll 8
      for (int i = 0; i < 64; ++i) {
        if (!((r >> i) & 1ull)) {
          kb |= 1ull << i;
          r |= diag[i];
        }
      }
      int total = s_total;
      int cnt = __popcll(kb);
      if (max_keep > 0 && total + cnt > max_keep) {
        // keep only the first (max_keep - total) set bits
        int need = max_keep - total;
        unsigned long long trimmed = 0ull, rest = kb;
        while (need-- > 0) {
          const unsigned long long low = rest & (~rest + 1ull);
          trimmed |= low;
          rest ^= low;
        }
        kb = trimmed;
        cnt = __popcll(kb);
      }
      s_kept = kb;
      s_base = total;
      s_total = total + cnt;
    }
    __syncthreads();
    const unsigned long long kb = s_kept;
    const int base = s_base, total = s_total;
    if (t < 64 && ((kb >> t) & 1ull)) {
      const int rank = base + __popcll(kb & ((1ull << t) - 1ull));
      o.emit(seg, off, rank, k * 64 + t, segs);
    }
    if (max_keep > 0 && total >= max_keep) break;  // CTA-uniform
    // OR the kept rows of this block into the removed-bitmap of all later blocks: 32 warps,
    // every warp issues the loads of up to four column blocks before it reduces any of them
    // (the scan is a chain of L2 round trips; this keeps one round trip per 128 column blocks)
    const bool k0 = (kb >> lane) & 1ull, k1 = (kb >> (lane + 32)) & 1ull;
    constexpr int kWarps = kScanThreads / 32;
    for (int w0 = k + 1 + warp; w0 < nblk; w0 += 4 * kWarps) {
      unsigned long long v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int w = w0 + u * kWarps;
        v[u] = 0ull;
        if (w < nblk) {
          const unsigned long long *row = m + (size_t)w * npad + k * 64;
          if (k0) v[u] = row[lane];
          if (k1) v[u] |= row[lane + 32];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int w = w0 + u * kWarps;
        if (w < nblk) {
          const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v[u]);
          const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v[u] >> 32));
          if (lane == 0) remv[w] |= ((unsigned long long)hi << 32) | lo;
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  o.finish(seg, off, n, s_total, t, kScanThreads);
}

// ----------------------------------------------------------------------------------------
// small segments: one CTA (128 threads) per segment, mask in shared memory.
// smem: float4 box[npad]; float2 wh[npad]; u64 mask[nblk][npad + 1]
// ----------------------------------------------------------------------------------------
constexpr int kSmallThreads = 128;
constexpr int kSmallMaxN = 512;

__global__ void __launch_bounds__(kSmallThreads)
    k_nms_small(NmsSegs segs, float thresh, int max_keep, int max_pad, NmsOut o) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int seg = blockIdx.x;
  int off, n;
  segs.get(seg, off, n);
  const int nblk = (n + 63) >> 6, npad = nblk << 6;
  float4 *box = reinterpret_cast<float4 *>(sm_raw);
  float2 *wh = reinterpret_cast<float2 *>(box + max_pad);
  unsigned long long *mk = reinterpret_cast<unsigned long long *>(wh + max_pad);
  const int mstride = npad + 1;
  __shared__ int s_total;
  const int t = threadIdx.x;
  for (int i = t; i < npad; i += kSmallThreads) {
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) bx = segs.load_box(off + i);
    box[i] = bx;
    wh[i] = make_float2(__fadd_rn(__fsub_rn(bx.z, bx.x), 1.f), __fadd_rn(__fsub_rn(bx.w, bx.y), 1.f));
  }
  if (t == 0) s_total = 0;
  __syncthreads();
  const bool fast = thresh >= 0.f && thresh < 1e30f;
  // items = (cb, row) with cb >= row/64
  for (int it = t; it < nblk * npad; it += kSmallThreads) {
    const int cb = it / npad, i = it - cb * npad;
    const int rb = i >> 6;
    if (cb < rb) continue;
    unsigned long long bits = 0ull;
    if (i < n) {
      const float4 a = box[i];
      const float Sa = __fmul_rn(wh[i].x, wh[i].y);
      const int jn = min(64, n - cb * 64);
      const int j0 = (rb == cb) ? (i & 63) + 1 : 0;
      for (int j = j0; j < jn; ++j)
        if (iou_gt(a, Sa, box[cb * 64 + j], wh[cb * 64 + j], thresh, fast)) bits |= 1ull << j;
    }
    mk[(size_t)cb * mstride + i] = bits;
  }
  __syncthreads();
  if (t < 32) {
    // warp 0: lane w owns the removed-bitmap word of column block w (nblk <= 8)
    const int lane = t;
    unsigned long long rem = 0ull;
    int total = 0;
    for (int k = 0; k < nblk; ++k) {
      unsigned long long r = __shfl_sync(0xffffffffu, rem, k);
      const int valid = min(64, n - k * 64);
      if (valid < 64) r |= ~0ull << valid;
      unsigned long long kb = 0ull;
      const unsigned long long *dg = mk + (size_t)k * mstride + k * 64;
      for (int i = 0; i < valid; ++i) {
        if (!((r >> i) & 1ull)) {
          kb |= 1ull << i;
          r |= dg[i];  // broadcast read, every lane runs the same chain
        }
      }
      int cnt = __popcll(kb);
      if (max_keep > 0 && total + cnt > max_keep) {
        int need = max_keep - total;
        unsigned long long trimmed = 0ull, rest = kb;
        while (need-- > 0) {
          const unsigned long long low = rest & (~rest + 1ull);
          trimmed |= low;
          rest ^= low;
        }
        kb = trimmed;
        cnt = __popcll(kb);
      }
      for (int q = lane; q < 64; q += 32)
        if ((kb >> q) & 1ull)
          o.emit(seg, off, total + __popcll(kb & ((1ull << q) - 1ull)), k * 64 + q, segs);
      total += cnt;
      if (max_keep > 0 && total >= max_keep) break;
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
    if (lane == 0) s_total = total;
  }
  __syncthreads();
  o.finish(seg, off, n, s_total, t, kSmallThreads);
}

// ----------------------------------------------------------------------------------------
// kept-list NMS for a bounded number of keeps (the proposal layer: post_nms_topN = 300).
// One CTA (512 threads) per segment walks the score-sorted boxes 64 at a time and never
// materialises the n x n mask: a chunk is tested against the boxes kept SO FAR (held in
// shared memory), then its 64 x 64 diagonal tile is resolved, survivors are appended.  Work
// is n_visited x n_kept pair tests instead of n^2 / 2, and the walk stops at max_keep keeps
// -- typically after a few hundred of the 6000 boxes.  Same greedy result as the mask scan:
// box j is suppressed iff some kept earlier box i has IoU(i, j) > thresh.
// smem: float4 kbox[max_keep]; float kSa[max_keep].
// ----------------------------------------------------------------------------------------
constexpr int kLazyThreads = 1024;  // 64 candidates x 16 slices of the kept list: the walk is a latency chain
constexpr int kLazyMaxKeep = 512;  // beyond, the all-SM mask + scan is as fast: measured at C1 (12000 -> 2000 keeps, heavy
                                   // suppression, ~180 chunks): cluster of 8 480 us, of 4 658 us, mask + scan 431 us
constexpr int kLazyMaxCluster = 8;

// The walk runs on a thread-block CLUSTER of cs CTAs per segment (cs = 1, 4 or 8, chosen by the launcher from
// max_keep): every CTA keeps the whole kept list and all 64 candidates of the chunk, but tests the candidates
// only against ITS share of the kept list (kept box kk belongs to CTA kk % cs); the cs partial 64-bit
// "suppressed" masks are exchanged through distributed shared memory (one st.shared::cluster per peer, one
// cluster barrier per chunk, double-buffered by chunk parity), after which every CTA resolves the chunk and
// appends the survivors itself -- identical lists everywhere, no second exchange.  Measured at C4 (24 images,
// 300 keeps from ~370 candidates): 1 CTA 43 us, cluster of 2 40 us, of 4 35 us, of 8 62 us (the barrier per chunk
// costs more than the shorter kept-list share saves).
__device__ __forceinline__ uint32_t lz_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t lz_cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void lz_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kLazyThreads)
    k_nms_lazy(NmsSegs segs, float thresh, int max_keep, NmsOut o) {
  extern __shared__ __align__(16) unsigned char lz_raw[];
  float4 *kbox = reinterpret_cast<float4 *>(lz_raw);
  float *kSa = reinterpret_cast<float *>(kbox + max_keep);
  __shared__ float4 cbox[64];
  __shared__ float2 cwh[64];
  __shared__ unsigned supw[kLazyThreads / 32];
  __shared__ unsigned diag32[64][2];
  __shared__ unsigned long long s_kept;
  __shared__ unsigned long long s_part[2][kLazyMaxCluster];  // partial suppressed masks of the peers, by chunk parity
  __shared__ int s_base, s_total;
  const uint32_t cs = lz_cluster_size(), cr = lz_cluster_ctarank();
  const int seg = blockIdx.x / cs;
  int off, n;
  segs.get(seg, off, n);
  const int nblk = (n + 63) >> 6;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int c = t & 63, slice = t >> 6;
  constexpr int kSlices = kLazyThreads / 64;
  const bool zf = thresh >= 0.f && thresh < 1e30f;
  if (t == 0) s_total = 0;
  float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < 64 && t < n) nxt = segs.load_box(off + t);
  __syncthreads();
  if (cs > 1) lz_cluster_sync();  // every CTA of the cluster runs before anyone writes into a peer
  for (int k = 0; k < nblk; ++k) {
    const int valid = min(64, n - k * 64);
    if (t < 64) {
      cbox[t] = nxt;
      cwh[t] = make_float2(__fadd_rn(__fsub_rn(nxt.z, nxt.x), 1.f), __fadd_rn(__fsub_rn(nxt.w, nxt.y), 1.f));
      const int j = (k + 1) * 64 + t;
      nxt = (j < n) ? segs.load_box(off + j) : make_float4(0.f, 0.f, 0.f, 0.f);  // prefetch
    }
    __syncthreads();
    const int K = s_total;
    // (a) candidate c against this CTA's share of the kept list, slice-strided; kept-box reads are broadcasts
    {
      const float4 cb = cbox[c];
      const float2 cw = cwh[c];
      bool sup = false;
      if (c < valid) {
        // two independent tests per trip: the test is a ~100-cycle dependent chain
        const int step = kSlices * (int)cs;
        int kk = (int)cr + slice * (int)cs;
        for (; kk + step < K && !sup; kk += 2 * step) {
          const bool s0 = iou_gt(kbox[kk], kSa[kk], cb, cw, thresh, zf);
          const bool s1 = iou_gt(kbox[kk + step], kSa[kk + step], cb, cw, thresh, zf);
          sup = s0 | s1;
        }
        if (!sup && kk < K) sup = iou_gt(kbox[kk], kSa[kk], cb, cw, thresh, zf);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, sup);
      if (lane == 0) supw[warp] = bal;
    }
    // (b) diagonal tile: row i vs the later boxes j of the same chunk (every CTA, it is 4 tests per thread)
#pragma unroll
    for (int q = 0; q < 4096 / kLazyThreads; ++q) {
      const int p = t + q * kLazyThreads;
      const int i = p >> 6, j = p & 63;
      bool bit = false;
      if (j > i && j < valid) {
        const float2 wi = cwh[i];
        bit = iou_gt(cbox[i], __fmul_rn(wi.x, wi.y), cbox[j], cwh[j], thresh, zf);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, bit);
      if (lane == 0) diag32[i][j >> 5] = bal;
    }
    __syncthreads();
    if (cs > 1) {
      // my partial mask goes to every CTA of the cluster (lane q of warp 0 writes peer q), then one barrier
      if (warp == 0) {
        static_assert(kLazyThreads / 32 == 32, "one supw word per lane");
        const unsigned v = supw[lane];
        const unsigned lo = __reduce_or_sync(0xffffffffu, (lane & 1) ? 0u : v);
        const unsigned hi = __reduce_or_sync(0xffffffffu, (lane & 1) ? v : 0u);
        if (lane < (int)cs) {
          const uint32_t local = smem_u32(&s_part[k & 1][cr]);
          uint32_t remote;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((uint32_t)lane));
          asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(remote), "l"(((unsigned long long)hi << 32) | lo) : "memory");
        }
      }
      lz_cluster_sync();
    }
    if (warp == 0) {
      // Greedy resolve of the chunk by ONE WARP: every lane keeps the removed mask r; lane l owns the
      // diagonal rows l and l + 32 and hands row i out by shuffle.  The 128 shuffles do not depend on r, so
      // they pipeline; the dependent chain per row is test / select / OR.
      unsigned long long r;
      if (cs > 1) {
        r = 0ull;
        for (uint32_t q = 0; q < cs; ++q) r |= s_part[k & 1][q];
      } else {
        const unsigned v = supw[lane];
        const unsigned lo = __reduce_or_sync(0xffffffffu, (lane & 1) ? 0u : v);
        const unsigned hi = __reduce_or_sync(0xffffffffu, (lane & 1) ? v : 0u);
        r = ((unsigned long long)hi << 32) | lo;
      }
      if (valid < 64) r |= ~0ull << valid;
      const unsigned long long rowA = ((unsigned long long)diag32[lane][1] << 32) | diag32[lane][0];
      const unsigned long long rowB = ((unsigned long long)diag32[lane + 32][1] << 32) | diag32[lane + 32][0];
      unsigned long long kb = 0ull;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const unsigned long long m = __shfl_sync(0xffffffffu, rowA, i);
        const bool alive = !((r >> i) & 1ull);
        kb |= alive ? (1ull << i) : 0ull;
        r |= alive ? m : 0ull;
      }
#pragma unroll
      for (int i = 32; i < 64; ++i) {
        const unsigned long long m = __shfl_sync(0xffffffffu, rowB, i - 32);
        const bool alive = !((r >> i) & 1ull);
        kb |= alive ? (1ull << i) : 0ull;
        r |= alive ? m : 0ull;
      }
      const int total = K;
      int cnt = __popcll(kb);
      if (total + cnt > max_keep) {
        int need = max_keep - total;
        unsigned long long trimmed = 0ull, rest = kb;
        while (need-- > 0) {
          const unsigned long long low = rest & (~rest + 1ull);
          trimmed |= low;
          rest ^= low;
        }
        kb = trimmed;
        cnt = __popcll(kb);
      }
      if (lane == 0) {
        s_kept = kb;
        s_base = total;
        s_total = total + cnt;
      }
    }
    __syncthreads();
    const unsigned long long kb = s_kept;
    const int total = s_total;
    if (t < 64 && ((kb >> t) & 1ull)) {
      const int rank = s_base + __popcll(kb & ((1ull << t) - 1ull));
      const float4 b = cbox[t];
      const float2 w = cwh[t];
      kbox[rank] = b;
      kSa[rank] = __fmul_rn(w.x, w.y);
      if (cr == 0) o.emit(seg, off, rank, k * 64 + t, segs);
    }
    if (total >= max_keep) break;  // uniform over the CTA and over the cluster (same data everywhere)
    __syncthreads();
  }
  __syncthreads();
  if (cr == 0) o.finish(seg, off, n, s_total, t, kLazyThreads);
  if (cs > 1) lz_cluster_sync();  // nobody leaves while a peer may still write its partial mask here
}

static size_t small_smem(int max_pad) {
  const int nblk = max_pad / 64;
  return (size_t)max_pad * (sizeof(float4) + sizeof(float2)) +
         (size_t)nblk * (max_pad + 1) * sizeof(unsigned long long);
}

size_t nms_mask_bytes(int nseg, int max_seg, int max_keep) {
  if (max_keep > 0 && max_keep <= kLazyMaxKeep) return 0;  // kept-list path: no mask
  if (nseg <= 0 || max_seg <= 0) return 0;
  const size_t nblk = (size_t)(max_seg + 63) / 64;
  return (size_t)nseg * nblk * nblk * 64 * sizeof(unsigned long long);
}

int nms_launch(const NmsSegs &segs, int nseg, int max_seg, float thresh, int max_keep,
               const NmsOut &out, void *workspace, size_t workspace_bytes, cudaStream_t st,
               int force_large) {
  if (nseg <= 0) return RLOD_OK;
  if (max_seg <= 0) {
    // every segment is empty: counts = 0
    RLOD_LAUNCH(RLOD_KERNEL_NMS_SMALL, st, k_nms_small<<<nseg, kSmallThreads, small_smem(64), st>>>(segs, thresh, max_keep, 64, out));
    return launch_status();
  }
  if (max_keep > 0 && max_keep <= kLazyMaxKeep && !force_large) {
    // bounded keeps (proposal layer): kept-list walk, no n x n mask, no workspace.  Cluster size by the length of
    // the kept list a chunk is tested against: 1 CTA up to 128 keeps, else 4
    const size_t smem = (size_t)max_keep * (sizeof(float4) + sizeof(float));
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(k_nms_lazy, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      attr_set = true;
    }
    static const int cs_env = getenv("RLOD_NMS_CLUSTER") ? atoi(getenv("RLOD_NMS_CLUSTER")) : 0;  // A/B switch
    int cs = max_keep <= 128 ? 1 : 4;
    if (cs_env == 1 || cs_env == 2 || cs_env == 4 || cs_env == 8) cs = cs_env;
    ProfScope _ps(RLOD_KERNEL_NMS_LAZY, st);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nseg * cs), 1, 1);
    cfg.blockDim = dim3(kLazyThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, k_nms_lazy, segs, thresh, max_keep, out);
    return e != cudaSuccess ? (int)e : launch_status();
  }
  if (max_seg <= kSmallMaxN && !force_large) {
    const int max_pad = (max_seg + 63) / 64 * 64;
    const size_t smem = small_smem(max_pad);
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(k_nms_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    RLOD_LAUNCH(RLOD_KERNEL_NMS_SMALL, st, k_nms_small<<<nseg, kSmallThreads, smem, st>>>(segs, thresh, max_keep, max_pad, out));
    return launch_status();
  }
  const int max_blk = (max_seg + 63) / 64;
  const size_t seg_stride = (size_t)max_blk * max_blk * 64;
  const size_t need = (size_t)nseg * seg_stride * sizeof(unsigned long long);
  if (!workspace || workspace_bytes < need) return RLOD_ENOSPC;
  if (nseg > 65535) return RLOD_EUNSUPPORTED;
  unsigned long long *mask = (unsigned long long *)workspace;
  const unsigned tiles = (unsigned)((long long)max_blk * (max_blk + 1) / 2);
  RLOD_LAUNCH(RLOD_KERNEL_NMS_MASK, st, k_nms_mask<<<dim3(tiles, nseg), 64, 0, st>>>(segs, thresh, max_blk, mask, seg_stride));
  RLOD_LAUNCH(RLOD_KERNEL_NMS_SCAN, st, k_nms_scan<<<nseg, kScanThreads, (size_t)max_blk * sizeof(unsigned long long), st>>>(
      segs, max_keep, mask, seg_stride, out));
  return launch_status();
}

}  // namespace rlod

using namespace rlod;

// debug knob of the tests (per calling thread: the library keeps no process-wide state)
static thread_local int g_force_large = 0;

RLOD_API int rlod_debug_nms_force_large(int on) {
  const int prev = g_force_large;
  g_force_large = on;
  return prev;
}

RLOD_API size_t rlod_nms_workspace_bytes(int nseg, int max_seg) {
  if (nseg <= 0 || max_seg <= 0) return 0;
  if (max_seg <= kSmallMaxN && !g_force_large) return 0;  // mask lives in shared memory
  const size_t nblk = (size_t)(max_seg + 63) / 64;
  return (size_t)nseg * nblk * nblk * 64 * sizeof(unsigned long long);
}

RLOD_API int rlod_nms(const float *dets, int n, int stride, float thresh, int max_keep, int *keep,
                      int *num_out, void *workspace, size_t workspace_bytes,
                      rlod_stream_t stream) {
  if (n < 0 || stride < 4 || !num_out) return RLOD_EINVAL;
  if (n > 0 && (!dets || !keep)) return RLOD_EINVAL;
  NmsSegs segs;
  segs.dets = dets;
  segs.stride = stride;
  segs.vec4 = (stride == 4 && ((uintptr_t)dets % 16) == 0) ? 1 : 0;
  segs.seg_offsets = nullptr;
  segs.uniform_n = n;
  segs.max_n = n;
  NmsOut o = {};
  o.keep = keep;
  o.num_out = num_out;
  return nms_launch(segs, 1, n, thresh, max_keep, o, workspace, workspace_bytes,
                    (cudaStream_t)stream, g_force_large);
}

RLOD_API int rlod_nms_batched(const float *dets, int stride, const int *seg_offsets, int nseg,
                              int max_seg, float thresh, int max_keep, int *keep, int *num_out,
                              void *workspace, size_t workspace_bytes, rlod_stream_t stream) {
  if (nseg < 0 || stride < 4 || max_seg < 0) return RLOD_EINVAL;
  if (nseg == 0) return RLOD_OK;
  if (!seg_offsets || !num_out) return RLOD_EINVAL;
  if (max_seg > 0 && (!dets || !keep)) return RLOD_EINVAL;
  NmsSegs segs;
  segs.dets = dets;
  segs.stride = stride;
  segs.vec4 = (stride == 4 && ((uintptr_t)dets % 16) == 0) ? 1 : 0;
  segs.seg_offsets = seg_offsets;
  segs.uniform_n = 0;
  segs.max_n = max_seg;
  NmsOut o = {};
  o.keep = keep;
  o.num_out = num_out;
  return nms_launch(segs, nseg, max_seg, thresh, max_keep, o, workspace, workspace_bytes,
                    (cudaStream_t)stream, g_force_large);
}
