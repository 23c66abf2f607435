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
#include <cstdlib>

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

// ----------------------------------------------------------------------------------------
// k_proposal_cluster: select + sort spread over a thread-block CLUSTER of 8 CTAs per image
// (distributed shared memory), instead of one 1024-thread CTA with 180 KB of shared memory
// per image (24 of 148 SMs at C4).  CTA c of the cluster owns the pixels [c * HWc, (c+1) * HWc)
// -- a contiguous range of anchor indices, since idx = pix * A + a -- and holds their 32-bit
// score keys in its own shared memory (22.5 KB at C4).
//   1. SELECT: the M-th smallest key by radix select, 8 bits per pass, 4 passes: local
//      256-bin histogram of the keys that match the digits fixed so far, summed into CTA 0's
//      shared memory by remote atomics (red.shared::cluster), read back by every CTA after a
//      cluster barrier.  Ties at the threshold are cut by anchor index (lower index first),
//      which is a prefix over the CTAs because they own index ranges in order.
//   2. BALANCE: every CTA writes its selected composite keys (key << 32 | idx) into a sort
//      buffer that is distributed over the cluster by global position (st.shared::cluster),
//      so each CTA sorts exactly ceil(M / 8) keys whatever the spatial distribution of scores.
//   3. SORT: bitonic sort of the local buffer.
//   4. MERGE: the final rank of a key = its local rank + the number of smaller keys in each of
//      the 7 other sorted lists (binary search through ld.shared::cluster); the key goes
//      straight to its place in the global `sel` row.
// Same result as k_proposal_sort_decode bit for bit (all composite keys are distinct).
// ----------------------------------------------------------------------------------------
constexpr int kClusterSize = 8;
constexpr int kClThreads = 512;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void dsmem_red_add(uint32_t addr, uint32_t v) {
  asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t dsmem_ld32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long dsmem_ld64(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void dsmem_st64(uint32_t addr, unsigned long long v) {
  asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void dsmem_st32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// bitonic_pass for a CTA of NT threads (same network as above)
template <int NL, int NT>
__device__ __forceinline__ void bitonic_pass_nt(unsigned long long *keys, int mp, int k, int jl, int t) {
  constexpr int Q = 1 << NL;
  const int sh = 31 - __clz(jl);
  for (int p = t; p < (mp >> NL); p += NT) {
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

// shared-memory layout of one CTA of the cluster (byte offsets from the dynamic base)
struct ClLayout {
  int hwc;      // pixels per CTA
  int kac;      // keys per CTA = hwc * A
  int cap;      // sort-buffer keys per CTA (power of two >= ceil(M / 8))
  size_t off_sort, off_hist, off_whist, off_cnt, bytes;
};
static __host__ __device__ inline ClLayout cl_layout(int A, int HW, int M) {
  ClLayout l;
  l.hwc = (HW + kClusterSize - 1) / kClusterSize;
  l.kac = l.hwc * A;
  int per = (M + kClusterSize - 1) / kClusterSize, cap = 64;
  while (cap < per) cap <<= 1;
  l.cap = cap;
  size_t off = ((size_t)l.kac * 4 + 15) & ~(size_t)15;
  l.off_sort = off;
  off += (size_t)(cap + (cap >> 3)) * 8;
  l.off_hist = off;          // [4 passes][256] cluster histograms (CTA 0's copy is the sum) + [256] local
  off += 5 * 256 * 4;        //   + [warps][256] private histograms
  l.off_whist = off;
  off += (size_t)(kClThreads / 32) * 256 * 4;
  l.off_cnt = off;           // [8] lt counts, [8] eq counts (CTA 0's copy is authoritative), scratch
  off += 32 * 4;
  l.bytes = off;
  return l;
}

__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kClThreads)
    k_proposal_cluster(PropArgs a, int mp) {
  pdl_trigger();  // k_proposal_decode is set up behind this grid (it waits for the selection itself)
  extern __shared__ __align__(16) unsigned char cl_raw[];
  const int HW = a.H * a.W, KA = HW * a.A, M = a.pre;
  const ClLayout L = cl_layout(a.A, HW, M);
  uint32_t *skeys = reinterpret_cast<uint32_t *>(cl_raw);
  unsigned long long *sortb = reinterpret_cast<unsigned long long *>(cl_raw + L.off_sort);
  uint32_t *ghist = reinterpret_cast<uint32_t *>(cl_raw + L.off_hist);   // [4][256]
  uint32_t *lhist = ghist + 4 * 256;                                      // [256]
  uint32_t *whist = reinterpret_cast<uint32_t *>(cl_raw + L.off_whist);   // [warps][256]
  uint32_t *cnts = reinterpret_cast<uint32_t *>(cl_raw + L.off_cnt);     // [0..7] lt, [8..15] eq, [16] local sel count
  __shared__ uint32_t s_digit, s_remaining, s_tie_index;
  const int t = threadIdx.x, lane = t & 31;
  const uint32_t c = cluster_ctarank();
  const int b = blockIdx.x / kClusterSize;
  const float *fg = a.scores + ((size_t)b * 2 * a.A + a.A) * HW;
  const int pix0 = (int)c * L.hwc, npix = max(0, min(L.hwc, HW - pix0));
  const int n_loc = npix * a.A, idx0 = pix0 * a.A;  // this CTA's anchor indices [idx0, idx0 + n_loc)

  // ---- keys of my pixels: one coalesced run per anchor plane --------------------------------
  for (int an = 0; an < a.A; ++an) {
    const float *src = fg + (size_t)an * HW + pix0;
    for (int p = t; p < npix; p += kClThreads) skeys[p * a.A + an] = desc_key(__ldg(src + p));
  }
  for (int i = t; i < 5 * 256; i += kClThreads) ghist[i] = 0u;
  if (t < 32) cnts[t] = 0u;
  for (int i = t; i < L.cap; i += kClThreads) sortb[sidx(i)] = ~0ull;
  __syncthreads();
  cluster_sync_all();  // every CTA's shared memory is initialised before anyone writes into it

  const uint32_t ghist0 = dsmem_addr(smem_u32(ghist), 0);
  const uint32_t cnts0 = dsmem_addr(smem_u32(cnts), 0);
  uint32_t T = 0xffffffffu, r_T = 0u;  // threshold key, number of keys == T to take (cluster-wide)
  const bool all = M >= KA;
  if (!all) {
    // ---- radix select of the M-th smallest key ----------------------------------------------
    uint32_t prefix = 0u, remaining = (uint32_t)M;
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = t; i < (kClThreads / 32) * 256; i += kClThreads) whist[i] = 0u;
      __syncthreads();
      // per-warp private histograms: the leading byte of a probability's key takes a handful of values, and
      // 5 600 same-address atomics on ONE shared histogram serialise across the whole CTA
      for (int i = t; i < n_loc; i += kClThreads) {
        const uint32_t kv = skeys[i];
        if (pass == 0 || (kv >> (shift + 8)) == prefix) atomicAdd(&whist[(t >> 5) * 256 + ((kv >> shift) & 255u)], 1u);
      }
      __syncthreads();
      if (t < 256) {
        uint32_t v = 0u;
#pragma unroll
        for (int w = 0; w < kClThreads / 32; ++w) v += whist[w * 256 + t];
        if (v) dsmem_red_add(ghist0 + (uint32_t)(pass * 256 + t) * 4u, v);
      }
      cluster_sync_all();
      // every CTA finds the digit itself from CTA 0's summed histogram
      if (t < 256) lhist[t] = dsmem_ld32(ghist0 + (uint32_t)(pass * 256 + t) * 4u);
      __syncthreads();
      if (t < 32) {
        // 8 bins per lane, warp scan of the lane sums, then the crossing bin
        uint32_t v[8], sum = 0u;
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = lhist[lane * 8 + q], sum += v[q];
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc += o;
        }
        uint32_t before = inc - sum;  // keys in bins below this lane's
        const bool mine = before < remaining && remaining <= inc;
        if (mine) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (remaining > before && remaining <= before + v[q]) {
              s_digit = (uint32_t)(lane * 8 + q);
              s_remaining = remaining - before;
            }
            before += v[q];
          }
        }
      }
      __syncthreads();
      prefix = (prefix << 8) | s_digit;
      remaining = s_remaining;
      __syncthreads();
    }
    T = prefix;
    r_T = remaining;
  }
  // ---- counts: keys < T and == T per CTA, published in CTA 0 --------------------------------
  {
    uint32_t lt = 0u, eq = 0u;
    for (int i = t; i < n_loc; i += kClThreads) {
      const uint32_t kv = skeys[i];
      lt += (all || kv < T) ? 1u : 0u;
      eq += (!all && kv == T) ? 1u : 0u;
    }
    lt = __reduce_add_sync(0xffffffffu, lt);
    eq = __reduce_add_sync(0xffffffffu, eq);
    if (lane == 0) {
      if (lt) dsmem_red_add(cnts0 + c * 4u, lt);
      if (eq) dsmem_red_add(cnts0 + (8u + c) * 4u, eq);
    }
  }
  cluster_sync_all();
  uint32_t lt_c[kClusterSize], eq_c[kClusterSize];
#pragma unroll
  for (int q = 0; q < kClusterSize; ++q) lt_c[q] = dsmem_ld32(cnts0 + q * 4u), eq_c[q] = dsmem_ld32(cnts0 + (8u + q) * 4u);
  // ties at T go to the lowest anchor indices: CTA q takes take_q = clamp(r_T - sum_{q' < q} eq_q', 0, eq_q)
  uint32_t base = 0u, my_take = 0u, eq_before = 0u;
#pragma unroll
  for (int q = 0; q < kClusterSize; ++q) {
    const uint32_t left = r_T > eq_before ? r_T - eq_before : 0u;
    const uint32_t take_q = min(left, eq_c[q]);
    if (q < (int)c) base += lt_c[q] + take_q;
    if (q == (int)c) my_take = take_q;
    eq_before += eq_c[q];
  }
  // partial ties inside this CTA: the index of its my_take-th tie (ordered walk by one warp; rare)
  uint32_t tie_hi = 0xffffffffu;  // take ties with local index <= tie_hi
  if (!all && my_take < eq_c[c]) {
    if (t < 32) {
      uint32_t seen = 0u, found = 0xffffffffu;
      if (my_take == 0u) found = 0u;  // none: "index <= found" must fail -> handled below by my_take == 0
      for (int i0 = 0; i0 < n_loc && found == 0xffffffffu; i0 += 32) {
        const int i = i0 + lane;
        const bool hit = i < n_loc && skeys[i] == T;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        const uint32_t cntb = (uint32_t)__popc(bal);
        if (seen + cntb >= my_take) {
          // the (my_take - seen)-th set bit of bal
          unsigned m = bal;
          for (uint32_t q = 1; q < my_take - seen; ++q) m &= m - 1u;
          found = (uint32_t)(i0 + __ffs(m) - 1);
        }
        seen += cntb;
      }
      if (lane == 0) s_tie_index = found;
    }
    __syncthreads();
    tie_hi = s_tie_index;
  }
  // ---- compaction into the cluster-distributed sort buffer ------------------------------------
  const uint32_t sort_local = smem_u32(sortb);
  for (int i0 = 0; i0 < n_loc; i0 += kClThreads) {
    const int i = i0 + t;
    bool take = false;
    uint32_t kv = 0u;
    if (i < n_loc) {
      kv = skeys[i];
      take = all || kv < T || (kv == T && my_take > 0u && (uint32_t)i <= tie_hi);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    unsigned wbase = 0;
    if (lane == 0 && bal) wbase = atomicAdd(&cnts[16], (unsigned)__popc(bal));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (take) {
      const uint32_t pos = base + wbase + (uint32_t)__popc(bal & ((1u << lane) - 1u));
      const uint32_t owner = pos / (uint32_t)L.cap, slot = pos - owner * (uint32_t)L.cap;
      if (owner < (uint32_t)kClusterSize)
        dsmem_st64(dsmem_addr(sort_local + (uint32_t)sidx((int)slot) * 8u, owner),
                   ((unsigned long long)kv << 32) | (unsigned)(idx0 + i));
    }
  }
  cluster_sync_all();
  // ---- local bitonic sort (cap keys, padded with ~0) ------------------------------------------
  for (int k = 2; k <= L.cap; k <<= 1) {
    int j = k >> 1;
    while (j >= 1) {
      if (j >= 4) {
        bitonic_pass_nt<3, kClThreads>(sortb, L.cap, k, j >> 2, t);
        j >>= 3;
      } else if (j == 2) {
        bitonic_pass_nt<2, kClThreads>(sortb, L.cap, k, 1, t);
        j = 0;
      } else {
        bitonic_pass_nt<1, kClThreads>(sortb, L.cap, k, 1, t);
        j = 0;
      }
      __syncthreads();
    }
  }
  cluster_sync_all();
  // ---- merge by rank: local rank + smaller keys in the other seven sorted lists ---------------
  unsigned long long *sel = a.sel + (size_t)b * mp;
  uint32_t rbase[kClusterSize];
#pragma unroll
  for (int q = 0; q < kClusterSize; ++q) rbase[q] = dsmem_addr(sort_local, (uint32_t)q);
  for (int i = t; i < L.cap; i += kClThreads) {
    const unsigned long long key = sortb[sidx(i)];
    if (key == ~0ull) continue;
    // lower bound in all eight lists at once (independent probes: their latencies overlap); in the key's own
    // list that is its local rank, because all composite keys are distinct
    int lo[kClusterSize];
#pragma unroll
    for (int q = 0; q < kClusterSize; ++q) lo[q] = 0;
    for (int half = L.cap >> 1; half >= 1; half >>= 1) {
#pragma unroll
      for (int q = 0; q < kClusterSize; ++q) {
        const unsigned long long v = dsmem_ld64(rbase[q] + (uint32_t)sidx(lo[q] + half - 1) * 8u);
        if (v < key) lo[q] += half;
      }
    }
    uint32_t rank = 0u;
#pragma unroll
    for (int q = 0; q < kClusterSize; ++q) {
      // lo[q] is in [0, cap - 1]: one more probe decides the last element
      const unsigned long long v = dsmem_ld64(rbase[q] + (uint32_t)sidx(lo[q]) * 8u);
      rank += (uint32_t)lo[q] + (v < key ? 1u : 0u);
    }
    if (rank < (uint32_t)M) sel[rank] = key;
  }
  cluster_sync_all();  // nobody leaves while its shared memory may still be read
}

// decode + clip the selected anchors in sorted order: one thread per (image, rank)
__global__ void __launch_bounds__(256) k_proposal_decode(PropArgs a, int mp) {
  // a programmatic dependent of the select / sort kernel, and the NMS kernel is one of this: set up early,
  // the sorted selection is complete and visible after the wait
  pdl_trigger();
  pdl_wait();
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
  // select + sort: a cluster of 8 CTAs per image when a slice of the keys and of the sort buffer fits one
  // CTA's shared memory (always at the reference's shapes), else one big CTA per image
  static const bool no_cluster = getenv("RLOD_PROPOSAL_V1") != nullptr;  // A/B switch: round 1's kernel
  const ClLayout cl = cl_layout(A, H * W, pre);
  if (!no_cluster && cl.bytes <= (size_t)(kMaxSmemPerCta - 1024)) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(k_proposal_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemPerCta - 1024);
      attr_set = true;
    }
    RLOD_LAUNCH(RLOD_KERNEL_PROPOSAL_SORT, st,
                k_proposal_cluster<<<B * kClusterSize, kClThreads, cl.bytes, st>>>(pa, mp));
  } else {
    size_t smem = (size_t)(mp + (mp >> 3)) * sizeof(unsigned long long);
    if (pa.keys_in_smem && key_bytes > smem) smem = align_up(key_bytes, 16);
    cudaFuncSetAttribute(k_proposal_sort_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    RLOD_LAUNCH(RLOD_KERNEL_PROPOSAL_SORT, st, k_proposal_sort_decode<<<B, kSortThreads, smem, st>>>(pa, mp));
  }
  RLOD_LAUNCH(RLOD_KERNEL_PROPOSAL_SORT, st,
              launch_after(k_proposal_decode, dim3((unsigned)cdiv((long long)B * pre, 256)), dim3(256), 0, st, proposal_pdl_enabled(), pa,
                           mp));
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
  // post_nms_topN <= 512: kept-list walk on a cluster of 4 CTAs per image (k_nms_lazy); larger: tiled mask + device scan
  return nms_launch(segs, B, pre, nms_thresh, post_nms_topN, o, ws.mask, ws.mask_bytes, st, 0);
}
