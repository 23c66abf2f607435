// Bitmask NMS for sm_100a, fully on device.
//
// Arithmetic is bit-identical to the reference's devIoU + strict '>' + greedy scan
// (lib/model/nms/src/nms_cuda_kernel.cu:31-39, 77-81, 123-144) as nvcc 12.9 compiles it for
// sm_100a: Sa = w_a*h_a (FMUL), S = fma(w_b, h_b, Sa), inter = w*h (FMUL),
// iou = inter / (S - inter) (IEEE division), bit = iou > thresh.
//
// What differs is everything around it:
//   - k_nms_mask_rm computes the upper triangle of the n x n bit mask for all segments (images /
//     (image, class) pairs) in one launch and stores it ROW-major (a thread writes its four
//     words of a 64 x 256 unit as one 32-byte sector; a persistent grid walks the units in row
//     order).  Pairs are classified branch-free by a division-free interval test; only the
//     razor-edge ones take the exact IEEE division, so the bits never differ from the
//     reference's.
//   - k_nms_scan3 resolves the serial greedy dependency ON THE DEVICE, one CTA per segment, as
//     three decoupled groups of warps: a resolver warp that owns the serial chain (one thread
//     walks the 64 x 64 diagonal tile, two instructions per box), near warps that apply the
//     keep bits to speculatively fetched tiles of the next columns, and far warps that own one
//     column block per thread and OR the kept rows of every earlier row block into a register.
//     It stops as soon as max_keep boxes are kept (the proposal layer only uses the first
//     post_nms_topN) and can emit the padded (B, post, 5) roi tensor directly.
//   - k_nms_mask + k_nms_scan (round 1: column-major tiles, one CTA barrier and one L2 round
//     trip per 64-box block) remain for segments beyond 57 344 boxes.
//   - k_nms_lazy: bounded keeps (post_nms_topN <= 512) need no mask at all -- a kept-list walk
//     on a cluster of 4 CTAs per segment.
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

// One row box against the 64 boxes of a column block -> 64 mask bits, branch-free: every pair runs the same
// ~28 straight-line instructions of iou_gt's interval test (the divergent early exits of iou_gt leave the
// schedulers at ~1 instruction per cycle on overlapping boxes such as RPN proposals) and is classified as
// yes / no / undecided; only the undecided ones -- razor-edge ties and non-finite values -- take the exact
// division afterwards.  Same decisions as iou_gt, pair by pair (thresh >= 0 only; else the plain loop).
__device__ __forceinline__ void classify_pair(const float4 &a, float Sa, const float4 &b, const float2 &bwh,
                                              float thresh, bool &yes, bool &und) {
  const float w = fmaxf(__fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 1.f), 0.f);
  const float h = fmaxf(__fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 1.f), 0.f);
  const float inter = __fmul_rn(w, h);
  const float u = __fsub_rn(__fmaf_rn(bwh.x, bwh.y, Sa), inter);
  const float tq = __fmul_rn(thresh, u);
  const bool empty = (w == 0.f) || (inter == 0.f);
  const bool ranged = (tq > 1e-30f) && (tq < 1e30f) && (inter < 1e30f);
  yes = !empty && ranged && (inter > __fmul_rn(tq, 1.000002f));
  const bool no = empty || (ranged && (inter < __fmul_rn(tq, 0.999998f)));
  und = !yes && !no;
}

__device__ __forceinline__ unsigned long long row_tile_bits(const float4 &a, float Sa, const float4 *cbox,
                                                            const float2 *cwh, int j0, int jn, float thresh,
                                                            bool fast) {
  unsigned long long valid = ~0ull;
  if (j0 > 0) valid &= j0 < 64 ? ~0ull << j0 : 0ull;
  if (jn < 64) valid &= jn > 0 ? ~0ull >> (64 - jn) : 0ull;
  unsigned ylo = 0u, yhi = 0u, ulo = 0xffffffffu, uhi = 0xffffffffu;
  if (fast) {
    ulo = uhi = 0u;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      bool y, q;
      classify_pair(a, Sa, cbox[j], cwh[j], thresh, y, q);
      ylo |= y ? 1u << j : 0u, ulo |= q ? 1u << j : 0u;
    }
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      bool y, q;
      classify_pair(a, Sa, cbox[32 + j], cwh[32 + j], thresh, y, q);
      yhi |= y ? 1u << j : 0u, uhi |= q ? 1u << j : 0u;
    }
  }
  unsigned long long bits = (((unsigned long long)yhi << 32) | ylo) & valid;
  unsigned long long und = (((unsigned long long)uhi << 32) | ulo) & valid;
  while (und) {
    const int j = __ffsll((long long)und) - 1;
    und &= und - 1ull;
    if (iou_gt(a, Sa, cbox[j], cwh[j], thresh, fast)) bits |= 1ull << j;
  }
  return bits;
}

// ----------------------------------------------------------------------------------------
// tile kernel, row-major mask (for k_nms_scan3): mask[row][column block], row stride = nblk rounded up to 4
// words.  A CTA takes one row block and FOUR column blocks, thread t owns row rb*64+t and writes its four
// words as one full 32-byte sector.  The scan's far gather -- one kept row against all later column blocks --
// then reads contiguous words (lanes = column blocks) instead of one 32-byte sector per 8-byte word: at
// n = 12000 that gather, not the serial chain, bounded the scan (188 k sectors through one SM).
// Words of column blocks left of the diagonal are written as zero and never read.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
    k_nms_mask_rm(NmsSegs segs, float thresh, unsigned long long *__restrict__ mask, size_t mask_seg_stride,
                  int max_blk, int nseg) {
  // persistent grid (24 CTAs per SM) over the (row block, column group) units, unit = blockIdx.x + i * gridDim.x:
  // 131 -> 111 us at n = 12000 against one CTA per unit (half of those exit at once, the rest share the SMs 32 deep)
  __shared__ float4 cbox[256];
  __shared__ float2 cwh[256];
  const int t = threadIdx.x;
  const int G = (max_blk + 3) >> 2;
  // units of one segment, row blocks ascending: row block rb has the column groups rb/4 .. G-1
  const long long per_seg = 4LL * ((long long)(max_blk >> 2) * G - (long long)(max_blk >> 2) * ((max_blk >> 2) - 1) / 2) +
                            (long long)(max_blk & 3) * (G - (max_blk >> 2));
  const long long total = per_seg * nseg;
  for (long long u = blockIdx.x; u < total; u += gridDim.x) {
    const int seg = (int)(u / per_seg);
    long long v = u - (long long)seg * per_seg;
    // quad q = rb >> 2 holds 4 (G - q) units
    int q = 0;
    while (v >= 4LL * (G - q)) v -= 4LL * (G - q), ++q;
    const int rb = 4 * q + (int)(v / (G - q)), g = q + (int)(v % (G - q));
    int off, n;
    segs.get(seg, off, n);
    const int nblk = (n + 63) >> 6;
    if (rb >= nblk || g * 4 >= nblk) continue;  // a shorter segment: nothing to compute, nobody waits for it
    const int stride = (nblk + 3) & ~3;
    __syncthreads();  // the previous unit's column boxes are no longer read
    for (int c4 = t; c4 < 256; c4 += 64) {
      const int j = g * 256 + c4;
      float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < n) bx = segs.load_box(off + j);
      cbox[c4] = bx;
      cwh[c4] = make_float2(__fadd_rn(__fsub_rn(bx.z, bx.x), 1.f), __fadd_rn(__fsub_rn(bx.w, bx.y), 1.f));
    }
    __syncthreads();
    const int i = rb * 64 + t;
    unsigned long long bits[4] = {0ull, 0ull, 0ull, 0ull};
    if (i < n) {
      const float4 a = segs.load_box(off + i);
      const float Sa = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
      const bool fast = thresh >= 0.f && thresh < 1e30f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int cb = g * 4 + c;
        if (cb < rb || cb >= nblk) continue;
        const int jn = min(64, n - cb * 64);
        const int j0 = (rb == cb) ? t + 1 : 0;
        bits[c] = row_tile_bits(a, Sa, cbox + c * 64, cwh + c * 64, j0, jn, thresh, fast);
      }
    }
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(mask + (size_t)seg * mask_seg_stride + (size_t)i * stride + g * 4);
    dst[0] = make_ulonglong2(bits[0], bits[1]);
    dst[1] = make_ulonglong2(bits[2], bits[3]);
  }
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

// Greedy resolve of one 64 x 64 diagonal tile by ONE thread: box i of the block is kept unless a kept
// earlier box (or an earlier block: r) removed it; a kept box removes the boxes of its row.  Written on
// 32-bit halves with selects instead of branches: the chain is test -> OR, ~10 cycles per box (the
// branchy 64-bit form, one taken or not-taken branch per box, costs 2.5 us per tile of the scan's step).
__device__ __forceinline__ void resolve_step(unsigned &rtest, unsigned &rother, unsigned &k, unsigned bit,
                                             unsigned dtest, unsigned dother, bool both) {
  // one test that sets a predicate, then predicated ORs: the dependent chain through r is two instructions
  // per box (the select form ptxas derives from C++ is three)
  if (both) {
    asm("{\n"
        ".reg .pred p;\n"
        ".reg .b32 t;\n"
        "and.b32 t, %0, %3;\n"
        "setp.eq.u32 p, t, 0;\n"
        "@p or.b32 %0, %0, %4;\n"
        "@p or.b32 %1, %1, %5;\n"
        "@p or.b32 %2, %2, %3;\n"
        "}\n"
        : "+r"(rtest), "+r"(rother), "+r"(k)
        : "r"(bit), "r"(dtest), "r"(dother));
  } else {
    asm("{\n"
        ".reg .pred p;\n"
        ".reg .b32 t;\n"
        "and.b32 t, %0, %2;\n"
        "setp.eq.u32 p, t, 0;\n"
        "@p or.b32 %0, %0, %3;\n"
        "@p or.b32 %1, %1, %2;\n"
        "}\n"
        : "+r"(rtest), "+r"(k)
        : "r"(bit), "r"(dtest));
  }
}

__device__ __forceinline__ unsigned long long resolve_tile(unsigned long long r, const unsigned long long *dg) {
  unsigned rlo = (unsigned)r, rhi = (unsigned)(r >> 32), klo = 0u, khi = 0u;
  const uint2 *d2 = reinterpret_cast<const uint2 *>(dg);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const uint2 d = d2[i];
    resolve_step(rlo, rhi, klo, 1u << i, d.x, d.y, true);
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const unsigned dh = d2[32 + i].y;  // row 32+i only removes boxes > 32+i: the upper half
    unsigned dummy = 0u;
    resolve_step(rhi, dummy, khi, 1u << i, dh, 0u, false);
  }
  return ((unsigned long long)khi << 32) | klo;
}

// ----------------------------------------------------------------------------------------
// scan kernel, decoupled (round 2): the greedy scan as three groups of warps of ONE CTA that
// never meet at a CTA barrier inside the loop.
//   * warp 0, the RESOLVER, owns the serial chain: block k's removed word -> resolve_tile -> keep bits
//     -> publish (kept[k], resolved = k+1) -> emit -> its own share of the next column (d = 1: the tile
//     (k, k+1) was fetched speculatively, all 64 rows, two steps ahead, and is masked by the keep bits in
//     registers).  Its step is ~0.5 us and contains no global-memory latency;
//   * warps 1..ND, the NEAR warps, do the same for the columns k+2 .. k+ND+1 (tile (k, k+1+d) in
//     registers two steps ahead, masked, reduced, atomicOr into remv[]) and tell the resolver through a
//     named barrier pair; they trail the resolver by at most a step;
//   * the remaining warps are the FAR threads: thread w owns column block w and ORs, for every row block
//     j <= w-ND-2 and every KEPT row i of it, the word m[w][64 j + i] into a private register -- one load
//     and one OR per (kept row, column), no reduction, no atomics (the per-step warp reductions of
//     k_nms_scan are 2 350 warp instructions per step on this single SM) -- following `resolved` at their
//     own pace; when a column is complete it is handed to remv[] and flagged.
// Valid for nblk <= 32 * (32 - 1 - ND) column blocks; larger inputs take k_nms_scan.
// ----------------------------------------------------------------------------------------
constexpr int kScan3Near = 3;
constexpr int kScan3MaxBlk = 32 * (kScanThreads / 32 - 1 - kScan3Near);

// remv[] |= (hi:lo) as two native 32-bit shared-memory ORs (a 64-bit atomicOr in shared memory is a CAS loop)
__device__ __forceinline__ void or64_shared(unsigned long long *p, unsigned lo, unsigned hi) {
  unsigned *q = reinterpret_cast<unsigned *>(p);
  if (lo) atomicOr(q, lo);
  if (hi) atomicOr(q + 1, hi);
}

__device__ __forceinline__ int ld_volatile_s32(const int *p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}

__global__ void __launch_bounds__(kScanThreads, 1)
    k_nms_scan3(NmsSegs segs, int max_keep, const unsigned long long *__restrict__ mask,
                size_t mask_seg_stride, NmsOut o) {
  typedef unsigned long long u64;
  extern __shared__ u64 sm_scan[];  // ring[kRing][2 + ND][64] | remv[nblk] | kept[nblk] | far_done[nblk] (int)
  __shared__ int s_resolved, s_last, s_total_out;
  constexpr int kRing = 4;  // tiles travel kRing - 1 steps ahead (8-byte async copies, no registers held)
  constexpr int ND = kScan3Near;
  constexpr int kFarWarps = kScanThreads / 32 - 1 - ND;
  constexpr int kDone = 0x7fffffff;
  const int seg = blockIdx.x;
  int off, n;
  segs.get(seg, off, n);
  const int nblk = (n + 63) >> 6, stride = (nblk + 3) & ~3;
  constexpr int kTiles = 2 + kScan3Near;  // per step: diagonal, (k, k+1) [resolver], (k, k+2 ..) [near warps]
  u64 *ring = sm_scan;
  u64 *remv = sm_scan + kRing * kTiles * 64, *kept = remv + nblk;
  int *far_done = reinterpret_cast<int *>(kept + nblk);
  const u64 *m = mask + (size_t)seg * mask_seg_stride;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const unsigned full = 0xffffffffu;
  // a column's far row blocks are dealt to NP warps (row block j goes to part j % NP): each part then has NP
  // resolver steps per row block to fetch its rows
  const int groups = (nblk + 31) >> 5;
  const int NP = groups * 4 <= kFarWarps ? 4 : (groups * 2 <= kFarWarps ? 2 : 1);
  for (int w = t; w < nblk; w += kScanThreads) remv[w] = 0ull, far_done[w] = 0;
  if (t == 0) s_resolved = 0, s_last = nblk - 1, s_total_out = 0;
  __syncthreads();
  // this warp's copy of tile (rb, cb) into ring slot `slot`, tile index ti: rows lane and lane + 32
  auto fetch_tile = [&](int slot, int ti, int rb, int cb) {
    u64 *dst = ring + (slot * kTiles + ti) * 64;
    if (cb < nblk && rb < nblk) {
      const u64 *src = m + (size_t)(rb * 64 + lane) * stride + cb;  // row-major mask: rows are `stride` words apart
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst + lane)), "l"(src) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst + lane + 32)), "l"(src + (size_t)32 * stride) : "memory");
    } else {
      dst[lane] = 0ull, dst[lane + 32] = 0ull;
    }
  };
  auto commit = []() { asm volatile("cp.async.commit_group;" ::: "memory"); };
  auto wait_oldest = [&]() {  // the group committed kRing - 1 steps ago has landed; make it visible to the warp
    asm volatile("cp.async.wait_group %0;" ::"n"(kRing - 1) : "memory");
    __syncwarp();
  };

  if (warp == 0) {
    // ---- resolver ------------------------------------------------------------------------
    for (int p = 0; p < kRing - 1; ++p) {
      fetch_tile(p, 0, p, p);
      fetch_tile(p, 1, p, p + 1);
      commit();
    }
    u64 r_near = 0ull;  // what row block k-1 removes in column k
    int total = 0;
    for (int k = 0; k < nblk; ++k) {
      {
        const int kk = k + kRing - 1, slot = kk % kRing;  // the slot of step k-1: consumed
        fetch_tile(slot, 0, kk, kk);
        fetch_tile(slot, 1, kk, kk + 1);
        commit();
        wait_oldest();
      }
      const u64 *diag = ring + ((k % kRing) * kTiles + 0) * 64;
      const u64 *near1 = ring + ((k % kRing) * kTiles + 1) * 64;
      // the near warps' updates of remv[k] (row blocks k-ND-1 .. k-2) are complete with their step k-2
      if (k >= 2) asm volatile("bar.sync %0, %1;" ::"r"(1 + (k & 1)), "n"(32 * (1 + ND)) : "memory");
      u64 kb = 0ull;
      bool done = false;
      if (lane == 0) {
        while (ld_volatile_s32(far_done + k) < NP) {
        }
        __threadfence_block();
        u64 r = *reinterpret_cast<volatile u64 *>(remv + k) | r_near;
        const int valid = min(64, n - k * 64);
        if (valid < 64) r |= ~0ull << valid;
        kb = resolve_tile(r, diag);
        int cnt = __popcll(kb);
        if (max_keep > 0 && total + cnt > max_keep) {
          int need = max_keep - total;
          u64 trimmed = 0ull, rest = kb;
          while (need-- > 0) {
            const u64 low = rest & (~rest + 1ull);
            trimmed |= low;
            rest ^= low;
          }
          kb = trimmed;
          cnt = __popcll(kb);
        }
        total += cnt;
        done = (max_keep > 0 && total >= max_keep) || k + 1 == nblk;
        kept[k] = kb;
        if (done) s_last = k, s_total_out = total;
        __threadfence_block();
        *reinterpret_cast<volatile int *>(&s_resolved) = done ? kDone : k + 1;
      }
      __syncwarp();
      // hand the keep bits to the near warps (they sync on the same barrier)
      asm volatile("bar.arrive %0, %1;" ::"r"(3 + (k & 1)), "n"(32 * (1 + ND)) : "memory");
      kb = __shfl_sync(full, kb, 0);
      total = __shfl_sync(full, total, 0);
      if (__shfl_sync(full, (int)done, 0)) break;
      // column k+1, row block k: mask the speculative tile by the keep bits
      const bool k0 = (kb >> lane) & 1ull, k1 = (kb >> (lane + 32)) & 1ull;
      const u64 v = (k0 ? near1[lane] : 0ull) | (k1 ? near1[lane + 32] : 0ull);
      const unsigned lo = __reduce_or_sync(full, (unsigned)v), hi = __reduce_or_sync(full, (unsigned)(v >> 32));
      r_near = ((u64)hi << 32) | lo;
      __syncwarp();  // the slot is refilled two iterations from now; every lane is done with it
    }
  } else if (warp <= ND) {
    // ---- near warps: column k+d of row block k, d = warp + 1; warp 1 also writes the kept boxes out -------
    const int d = warp + 1;
    for (int p = 0; p < kRing - 1; ++p) {
      fetch_tile(p, d, p, p + d);
      commit();
    }
    int base = 0, pr0 = -1, pr1 = -1;
    float4 pb0 = make_float4(0.f, 0.f, 0.f, 0.f), pb1 = pb0;
    auto store_pending = [&]() {
      if (pr0 >= 0) {
        float *q = o.rois + ((size_t)seg * o.post + pr0) * 5;
        q[0] = (float)seg, q[1] = pb0.x, q[2] = pb0.y, q[3] = pb0.z, q[4] = pb0.w;
      }
      if (pr1 >= 0) {
        float *q = o.rois + ((size_t)seg * o.post + pr1) * 5;
        q[0] = (float)seg, q[1] = pb1.x, q[2] = pb1.y, q[3] = pb1.z, q[4] = pb1.w;
      }
      pr0 = pr1 = -1;
    };
    for (int k = 0; k < nblk; ++k) {
      {
        const int kk = k + kRing - 1;
        fetch_tile(kk % kRing, d, kk, kk + d);
        commit();
        wait_oldest();
      }
      const u64 *nt = ring + ((k % kRing) * kTiles + d) * 64;
      asm volatile("bar.sync %0, %1;" ::"r"(3 + (k & 1)), "n"(32 * (1 + ND)) : "memory");
      const bool last = ld_volatile_s32(&s_resolved) == kDone && ld_volatile_s32(&s_last) == k;
      const u64 kb = *reinterpret_cast<volatile u64 *>(kept + k);
      const bool k0 = (kb >> lane) & 1ull, k1 = (kb >> (lane + 32)) & 1ull;
      if (!last) {
        const u64 v = (k0 ? nt[lane] : 0ull) | (k1 ? nt[lane + 32] : 0ull);
        const unsigned lo = __reduce_or_sync(full, (unsigned)v), hi = __reduce_or_sync(full, (unsigned)(v >> 32));
        if (lane == 0 && k + d < nblk && (lo | hi)) or64_shared(&remv[k + d], lo, hi);
        __threadfence_block();
        // the resolver waits for this at its step k+2
        if (k + 2 < nblk) asm volatile("bar.arrive %0, %1;" ::"r"(1 + (k & 1)), "n"(32 * (1 + ND)) : "memory");
      }
      if (warp == 1) {
        // write-out, one step behind: the boxes fetched at the previous step are stored now, this step's are
        // fetched (an L2 round trip that must not sit between two barrier hand-overs)
        store_pending();
        if (k0) {
          const int rank = base + __popcll(kb & ((1ull << lane) - 1ull)), idx = k * 64 + lane;
          if (o.keep) o.keep[off + rank] = idx;
          if (o.rois && rank < o.post) pr0 = rank, pb0 = segs.load_box(off + idx);
        }
        if (k1) {
          const int rank = base + __popcll(kb & ((1ull << (lane + 32)) - 1ull)), idx = k * 64 + lane + 32;
          if (o.keep) o.keep[off + rank] = idx;
          if (o.rois && rank < o.post) pr1 = rank, pb1 = segs.load_box(off + idx);
        }
        base += __popcll(kb);
      }
      if (last) break;
      __syncwarp();
    }
    store_pending();
  } else {
    // ---- far warps: warp (group g, part q) owns columns 32 g + lane and their row blocks j = q (mod NP) ----
    const int fw = warp - ND - 1;
    const int q = fw % NP, g = fw / NP;
    const int w = g * 32 + lane;
    const bool mine = g < groups && w < nblk;
    const int last_j = mine ? w - ND - 2 : -1;  // row blocks 0 .. last_j are far for column w
    const int warp_last = __reduce_max_sync(full, last_j);
    const u64 *col = m + (mine ? w : 0);  // lanes = consecutive column blocks: a kept row is one contiguous read
    u64 r = 0ull;
    int seen = 0;
    bool open = mine;
    // a part is handed over as soon as its last row block is in (the resolver waits for NP parts per column)
    auto close_if_done = [&](int next_j) {
      if (open && next_j > last_j) {
        if (r) or64_shared(&remv[w], (unsigned)r, (unsigned)(r >> 32));
        __threadfence_block();
        atomicAdd(far_done + w, 1);
        open = false;
      }
    };
    close_if_done(q);
    for (int j = q; j <= warp_last; j += NP) {
      if (seen <= j) {
        if (lane == 0)
          while ((seen = ld_volatile_s32(&s_resolved)) <= j) __nanosleep(64);
        seen = __shfl_sync(full, seen, 0);
        __threadfence_block();
      }
      if (seen == kDone && j > ld_volatile_s32(&s_last)) break;  // the resolver stopped before row block j
      u64 kb = *reinterpret_cast<volatile u64 *>(kept + j);
      if (j <= last_j) {
        const u64 *row = col + (size_t)j * 64 * stride;
        while (kb) {  // 16 loads in flight
          u64 v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            v[e] = 0ull;
            if (kb) {
              const int i = __ffsll((long long)kb) - 1;
              kb &= kb - 1ull;
              v[e] = row[(size_t)i * stride];
            }
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) r |= v[e];
        }
      }
      close_if_done(j + NP);
    }
    close_if_done(0x7fffffff);  // the resolver stopped early: nobody reads the column any more
  }
  __syncthreads();
  o.finish(seg, off, n, s_total_out, t, kScanThreads);
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
  // launched as a programmatic dependent of whatever kernel precedes it on the stream (in the proposal layer:
  // k_proposal_decode): the boxes and the segment table are complete and visible after the wait
  pdl_wait();
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
  const size_t nblk = (size_t)(max_seg + 63) / 64, stride = (nblk + 3) / 4 * 4;  // row-major rows: 4-word sectors
  return (size_t)nseg * nblk * stride * 64 * sizeof(unsigned long long);
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
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = proposal_pdl_enabled() ? 2 : 1;
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
  const size_t seg_stride = (size_t)max_blk * ((max_blk + 3) / 4 * 4) * 64;
  const size_t need = (size_t)nseg * seg_stride * sizeof(unsigned long long);
  if (!workspace || workspace_bytes < need) return RLOD_ENOSPC;
  if (nseg > 65535) return RLOD_EUNSUPPORTED;
  unsigned long long *mask = (unsigned long long *)workspace;
  const unsigned tiles = (unsigned)((long long)max_blk * (max_blk + 1) / 2);
  static const bool scan_v1 = getenv("RLOD_NMS_SCAN_V1") != nullptr;  // A/B switch: the unpipelined scan
  if (max_blk <= kScan3MaxBlk && !scan_v1 && force_large != 2 && nseg <= 65535) {
    // row-major mask (rows of (max_blk rounded up to 4) words; fits the same workspace: 64 max_blk rows) + decoupled scan
    const int g4 = (max_blk + 3) / 4;
    const long long units = (long long)nseg * g4 * max_blk;  // an upper bound of the units (the kernel counts exactly)
    const long long cap = (long long)kSmCount * 24;
    const int mask_ctas = (int)(units < cap ? units : cap);
    RLOD_LAUNCH(RLOD_KERNEL_NMS_MASK, st,
                k_nms_mask_rm<<<mask_ctas, 64, 0, st>>>(segs, thresh, mask, seg_stride, max_blk, nseg));
    RLOD_LAUNCH(RLOD_KERNEL_NMS_SCAN, st, k_nms_scan3<<<nseg, kScanThreads, (size_t)max_blk * 20 + 4 * (2 + kScan3Near) * 64 * 8 + 16, st>>>(
        segs, max_keep, mask, seg_stride, out));
    return launch_status();
  }
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
  const size_t nblk = (size_t)(max_seg + 63) / 64, stride = (nblk + 3) / 4 * 4;  // row-major rows: 4-word sectors
  return (size_t)nseg * nblk * stride * 64 * sizeof(unsigned long long);
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
