// Per-image roi lists shared by the RoIAlign and RoIPool plane kernels: workspace layout, the
// "rois are grouped by image" fast path with its fix-up, and the per-image counting sort by a
// small key.  Kernels are static: every .cu that includes this header gets its own copy.
#pragma once
#include "rlod_common.cuh"

namespace rlod {

// ----------------------------------------------------------------------------------------
// workspace layout
// ----------------------------------------------------------------------------------------
// Forward RoIAlign over a map that does not fit one CTA's shared memory: the map is covered by overlapping
// tiles, every tile is a virtual image of the plane kernel (image b, tile (ty, tx) -> list b * ny * nx +
// ty * nx + tx), and a roi goes to the tile that holds all of its taps (roi_align.cu).  ny = nx = 1: the
// whole map is the tile.
constexpr int kMaxTilesPerImage = 128;
struct TileGrid {
  int ny, nx;  // tiles per image
  int th, tw;  // tile size in pixels (rows, columns); the last tile of an axis may hang over the map's edge
  int sy, sx;  // distance between tile origins
};

struct AlignWs {
  int *flag;     // [4]   flag[0] != 0: rois are not grouped by image
  int *img_off;  // [B * kMaxTilesPerImage + 1] roi list offsets per (virtual) image
  int *cursor;   // [B * kMaxTilesPerImage]
  int *order;    // [R]   roi ids grouped by image (stable)
  int *roi_b;    // [R]   batch index (0 when out of range: the plan is all-invalid then)
  int *plan;     // [R * words]
  int *ext;      // [R * 32] forward-kernel record (8x8 grids only)
  int *order2;   // [R]   `order` with every image's list partitioned by walk mode (stable)
  int *own;      // [R * 192] record of the lock-free backward kernel (roi_align_bwd.cu, 8x8 grids only)
  size_t bytes;
};

static inline AlignWs carve_align_ws(void *base, int B, int R, int GH, int GW) {
  AlignWs w;
  size_t off = 0;
  char *p = (char *)base;
  auto take = [&](size_t n) {
    char *q = p ? p + off : nullptr;
    off += align_up(n, 128);
    return q;
  };
  w.flag = (int *)take(4 * sizeof(int));
  // (lists per virtual image when the forward tiles a large map)
  w.img_off = (int *)take(((size_t)(B > 0 ? B : 0) * kMaxTilesPerImage + 1) * sizeof(int));
  w.cursor = (int *)take((size_t)(B > 0 ? B : 1) * kMaxTilesPerImage * sizeof(int));
  w.order = (int *)take((size_t)(R > 0 ? R : 1) * sizeof(int));
  w.roi_b = (int *)take((size_t)(R > 0 ? R : 1) * sizeof(int));
  w.plan = (int *)take((size_t)(R > 0 ? R : 1) * (size_t)(2 * GH + 2 * GW) * sizeof(int));
  w.ext = (int *)take((GH == 8 && GW == 8) ? (size_t)(R > 0 ? R : 1) * 32 * sizeof(int) : 0);
  w.order2 = (int *)take((GH == 8 && GW == 8) ? (size_t)(R > 0 ? R : 1) * sizeof(int) : 0);
  w.own = (int *)take((GH == 8 && GW == 8) ? (size_t)(R > 0 ? R : 1) * 192 * sizeof(int) : 0);
  w.bytes = off;
  return w;
}


// Called by one thread per roi r (b = its validated batch index): roi lists under the
// assumption that rois are grouped by image (what _ProposalLayer emits); k_roi_group_fixup
// redoes them when the flag is raised.
__device__ __forceinline__ void roi_list_mark(const float *__restrict__ rois, int r, int R, int B, int b,
                                              const AlignWs &ws) {
  ws.roi_b[r] = b;
  ws.order[r] = r;
  int prev;
  if (r == 0) {
    prev = -1;
  } else {
    const float pf = rois[(size_t)(r - 1) * 5];
    const int pi = (int)pf;
    prev = ((pf >= 0.f) && (pi < B)) ? pi : 0;
  }
  if (b < prev) atomicOr(ws.flag, 1);
  for (int q = prev + 1; q <= b; ++q) ws.img_off[q] = r;
  if (r == R - 1)
    for (int q = b + 1; q <= B; ++q) ws.img_off[q] = R;
}

// one warp: stable counting sort of roi ids by image, only when the rois were not grouped
static __global__ void k_roi_group_fixup(int R, int B, AlignWs ws) {
  pdl_trigger();  // (see k_roi_lists_finish)
  pdl_wait();
  if (ws.flag[0] == 0) return;
  const int lane = threadIdx.x;
  const unsigned full = 0xffffffffu;
  for (int b = lane; b < B; b += 32) ws.cursor[b] = 0;
  __syncwarp();
  for (int base = 0; base < R; base += 32) {
    const int r = base + lane;
    const int b = r < R ? ws.roi_b[r] : -1 - lane;
    const unsigned peers = __match_any_sync(full, b);
    if (r < R && lane == __ffs(peers) - 1) ws.cursor[b] += __popc(peers);
    __syncwarp();
  }
  // exclusive scan over images, 32 at a time
  int carry = 0;
  for (int base = 0; base < B; base += 32) {
    const int b = base + lane;
    const int c = b < B ? ws.cursor[b] : 0;
    int inc = c;
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(full, inc, d);
      if (lane >= d) inc += t;
    }
    if (b < B) {
      ws.img_off[b] = carry + inc - c;
      ws.cursor[b] = carry + inc - c;
    }
    carry += __shfl_sync(full, inc, 31);
  }
  if (lane == 0) ws.img_off[B] = carry;
  __syncwarp();
  for (int base = 0; base < R; base += 32) {
    const int r = base + lane;
    const int b = r < R ? ws.roi_b[r] : -1 - lane;
    const unsigned peers = __match_any_sync(full, b);
    const int leader = __ffs(peers) - 1;
    int cur = 0;
    if (r < R && lane == leader) {
      cur = ws.cursor[b];
      ws.cursor[b] = cur + __popc(peers);
    }
    cur = __shfl_sync(full, cur, leader);
    if (r < R) ws.order[cur + __popc(peers & ((1u << lane) - 1u))] = r;
    __syncwarp();
  }
}

// one warp per image: stable counting sort of its roi list by a small key taken from the
// record -- forward: the walk mode (word 0 bit 30), so that the four rois a warp serves
// together stage with the same strides; backward: the longest run of lanes that share a
// column (word 16 bits 20-22), so that a narrow roi does not impose its extra scatter rounds
// on three wide ones.
constexpr int kOrderThreads = 256;
constexpr int kOrderMaxKeys = 64;    // nkeys: a power of two <= 64
constexpr int kOrderMaxRois = 2048;  // per image, held in shared memory; longer lists take the slow path
static __global__ void __launch_bounds__(kOrderThreads)
    k_roi_order_by_key(const int *__restrict__ ext, const int *__restrict__ order,
                       const int *__restrict__ img_off, int word, int shift, int nkeys,
                       int *__restrict__ order2) {
  // one CTA per image: all keys are fetched at once (one dependent pair of loads for the whole
  // list) into shared memory, a per-key count gives the key offsets, then every roi finds its
  // rank among the rois of its key that precede it (stable)
  __shared__ unsigned short s_key[kOrderMaxRois];
  __shared__ int s_tot[kOrderMaxKeys], s_base[kOrderMaxKeys];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31;
  pdl_trigger();  // (see k_roi_lists_finish)
  pdl_wait();
  const int r0 = img_off[b], r1 = img_off[b + 1], n = r1 - r0;
  const unsigned full = 0xffffffffu;
  if (n > kOrderMaxRois) {
    // slow path (one warp): stable counting sort, key by key
    if (t >= 32) return;
    const unsigned below = (1u << lane) - 1u;
    int start = r0;
    for (int key = 0; key < nkeys; ++key) {
      int c = start;
      for (int base = r0; base < r1; base += 32) {
        const int i = base + lane;
        const int r = i < r1 ? order[i] : 0;
        const bool hit = i < r1 && ((ext[(size_t)r * 32 + word] >> shift) & (nkeys - 1)) == key;
        const unsigned m = __ballot_sync(full, hit);
        if (hit) order2[c + __popc(m & below)] = r;
        c += __popc(m);
      }
      start = c;
    }
    return;
  }
  if (t < kOrderMaxKeys) s_tot[t] = 0;
  __syncthreads();
  for (int i = t; i < n; i += kOrderThreads) {
    const int key = (ext[(size_t)order[r0 + i] * 32 + word] >> shift) & (nkeys - 1);
    s_key[i] = (unsigned short)key;
    atomicAdd(&s_tot[key], 1);
  }
  __syncthreads();
  if (t == 0) {
    int acc = 0;
    for (int k = 0; k < nkeys; ++k) s_base[k] = acc, acc += s_tot[k];
  }
  __syncthreads();
  for (int i = t; i < n; i += kOrderThreads) {
    const int key = s_key[i];
    int rank = 0;
    for (int q = 0; q < i; ++q) rank += s_key[q] == key;  // n is a few hundred: a short scan
    order2[r0 + s_base[key] + rank] = order[r0 + i];
  }
}


// k_roi_lists_finish: group fix-up and 2-way partition in ONE launch, one CTA per image (the forward
// RoIAlign runs this between the plan and the walk kernel; two launches -- a one-warp fix-up that
// normally does nothing and a per-image sort -- cost ~11 us per call in the stream).
//   * rois grouped by image (flag clear): the image's list is the range the plan kernel marked;
//   * otherwise: the CTA collects its image's rois by scanning all of them (every CTA scans; the case
//     is rare and the scan is R loads), and publishes its own img_off entry;
//   * the list is then stably partitioned by one bit of the record (the walk mode), by ballots.
// An image with more than kOrderMaxRois rois gets its list collected but not partitioned (any order is
// correct for the walk kernel; the partition only keeps its staging stores bank-disjoint).
static __global__ void __launch_bounds__(kOrderThreads)
    k_roi_lists_finish(const int *__restrict__ ext, int word, int shift, int R, int B, AlignWs ws) {
  __shared__ int s_ids[kOrderMaxRois];
  __shared__ int s_wc[2][kOrderThreads / 32];
  __shared__ int s_n, s_before;
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const unsigned full = 0xffffffffu, below = (1u << lane) - 1u;
  constexpr int NW = kOrderThreads / 32;
  int n, first;
  // launched as a programmatic dependent of the plan kernel (the forward RoIAlign): the kernel after this
  // one may be set up now, and the plan is complete and visible after the wait (both do nothing otherwise)
  pdl_trigger();
  pdl_wait();
  const bool grouped = ws.flag[0] == 0;
  if (grouped) {
    first = ws.img_off[b];
    n = ws.img_off[b + 1] - first;
  } else {
    // how many rois belong to image b and to the images before it (every CTA scans all R batch indices)
    int mine = 0, before = 0;
    for (int r = t; r < R; r += kOrderThreads) {
      const int rb = ws.roi_b[r];
      mine += rb == b, before += rb < b;
    }
    mine = __reduce_add_sync(full, mine), before = __reduce_add_sync(full, before);
    if (lane == 0) s_wc[0][warp] = mine, s_wc[1][warp] = before;
    __syncthreads();
    n = 0, first = 0;
    for (int w2 = 0; w2 < NW; ++w2) n += s_wc[0][w2], first += s_wc[1][w2];
    __syncthreads();
    if (t == 0) {
      ws.img_off[b] = first;
      if (b == B - 1) ws.img_off[B] = first + n;
    }
  }
  const bool fits = n <= kOrderMaxRois;  // else: the list is only collected, not partitioned
  if (grouped) {
    for (int i = t; i < n; i += kOrderThreads) {
      const int id = ws.order[first + i];
      if (fits) s_ids[i] = id;
      else ws.order2[first + i] = id;
    }
  } else {
    // ordered collection of the rois of image b: every warp takes one contiguous range of roi ids, counts its
    // matches, and after one prefix over the warps writes them out in order (two passes over the batch indices,
    // no CTA barrier inside them -- the former chunk-by-chunk collection paid two barriers per 256 rois)
    const int per_warp = ((R + NW - 1) / NW + 31) & ~31;
    const int lo = warp * per_warp, hi = min(R, lo + per_warp);
    int mine = 0;
    for (int c0 = lo; c0 < hi; c0 += 32) {
      const int r = c0 + lane;
      mine += __popc(__ballot_sync(full, r < hi && ws.roi_b[r] == b));
    }
    if (lane == 0) s_wc[0][warp] = mine;
    __syncthreads();
    int pos = 0;
    for (int w2 = 0; w2 < warp; ++w2) pos += s_wc[0][w2];
    for (int c0 = lo; c0 < hi; c0 += 32) {
      const int r = c0 + lane;
      const bool hit = r < hi && ws.roi_b[r] == b;
      const unsigned m = __ballot_sync(full, hit);
      if (hit) {
        const int p = pos + __popc(m & below);
        if (fits) s_ids[p] = r;
        else ws.order2[first + p] = r;
      }
      pos += __popc(m);
    }
  }
  if (!fits) return;
  __syncthreads();
  // stable 2-way partition by the key bit: zeros first
  int run0 = 0, run1 = 0, n0 = 0;
  for (int c0 = 0; c0 < n; c0 += kOrderThreads) {  // count the zeros
    const int i = c0 + t;
    const bool z = i < n && ((ext[(size_t)s_ids[i] * 32 + word] >> shift) & 1) == 0;
    n0 += __popc(__ballot_sync(full, z));
  }
  if (lane == 0) s_wc[0][warp] = n0;
  __syncthreads();
  n0 = 0;
  for (int w2 = 0; w2 < NW; ++w2) n0 += s_wc[0][w2];
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += kOrderThreads) {
    const int i = c0 + t;
    const int id = i < n ? s_ids[i] : 0;
    const int key = i < n ? ((ext[(size_t)id * 32 + word] >> shift) & 1) : -1;
    const unsigned m0 = __ballot_sync(full, key == 0), m1 = __ballot_sync(full, key == 1);
    if (lane == 0) s_wc[0][warp] = __popc(m0), s_wc[1][warp] = __popc(m1);
    __syncthreads();
    int p0 = run0, p1 = run1, t0 = 0, t1 = 0;
    for (int w2 = 0; w2 < NW; ++w2) {
      if (w2 < warp) p0 += s_wc[0][w2], p1 += s_wc[1][w2];
      t0 += s_wc[0][w2], t1 += s_wc[1][w2];
    }
    if (key == 0) ws.order2[first + p0 + __popc(m0 & below)] = id;
    if (key == 1) ws.order2[first + n0 + p1 + __popc(m1 & below)] = id;
    run0 += t0, run1 += t1;
    __syncthreads();
  }
}

// roi_align_bwd.cu: the lock-free backward (one warp owns the planes of an image x 4 channels)
bool bwd_own_supported(int H, int W, int pool_mode);
int launch_bwd_own(const float *grad_out, const float *rois, int B, int C, int H, int W, int R,
                   float spatial_scale, int pool_mode, int accumulate, float *grad_in, const AlignWs &ws,
                   cudaStream_t st);

}  // namespace rlod
