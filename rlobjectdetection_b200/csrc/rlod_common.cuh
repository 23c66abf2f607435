// Shared helpers for the sm_100a kernels of librlod_sm100a.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>

#include "rlod.h"

#define RLOD_API extern "C" __attribute__((visibility("default")))

namespace rlod {

constexpr int kSmCount = 148;           // B200: 2 dies x 74 SMs
constexpr int kMaxSmemPerCta = 232448;  // 227 KB opt-in limit per CTA on sm_100a

static inline int launch_status() { return (int)cudaGetLastError(); }

// every kernel launch of the library goes through a ProfScope: it counts the launch
// (rlod_launch_count) and, when rlod_profile_enable(1) is set, brackets it with CUDA events on
// its own stream (rlod_profile_collect).
void note_launch(int n);
class ProfScope {
 public:
  ProfScope(int kernel_id, cudaStream_t st);
  ~ProfScope();

 private:
  int id_;
  cudaStream_t st_;
  bool on_;
  cudaEvent_t a_, b_;
};

#define RLOD_LAUNCH(ID, ST, ...)        \
  do {                                  \
    ::rlod::ProfScope _ps((ID), (ST));  \
    __VA_ARGS__;                        \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// ---- programmatic dependent launch (griddepcontrol) -------------------------------------
// A kernel launched through launch_after() with overlap = true may start while the kernel before it
// on the stream is still running, as soon as every CTA of that kernel has called pdl_trigger() (or
// exited); what it reads of the earlier kernel's results it reads after pdl_wait(), which returns
// once the earlier kernel has finished and its writes are visible.  Without the launch attribute
// both instructions do nothing, so kernels that carry them can be launched either way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// A/B switch: RLOD_NO_PDL=1 launches the kernels of a call strictly one after the other
static inline bool pdl_enabled() {
  static const bool on = getenv("RLOD_NO_PDL") == nullptr;
  return on;
}

// The proposal layer's decode / NMS kernels as dependents of the kernel before them: wired, off by default
// (RLOD_PROPOSAL_PDL=1).  A 3-image shard's layer gains 4 us (80.9 -> 76.8 us), but in the C4 step the early CTAs
// wait on SMs that the pooling kernel would otherwise use -- the step is bound by total SM time -- and the step
// loses 2 % (26 870 -> 26 330 images/s).
static inline bool proposal_pdl_enabled() {
  static const bool on = getenv("RLOD_PROPOSAL_PDL") != nullptr && pdl_enabled();
  return on;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_after(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                       cudaStream_t st, bool overlap, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = overlap ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}


static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP) --------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// bulk async store shared -> global (TMA engine, SASS: UBLKCP), tracked by bulk groups
__device__ __forceinline__ void bulk_s2g_nocommit(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
// the same with an L2 eviction policy (createpolicy): streamed outputs are marked evict-first so
// that they do not push the prefetched feature planes out of L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_s2g_hint_nocommit(void *dst_gmem, uint32_t src_smem, uint32_t bytes,
                                                       uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
               "r"(src_smem), "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                              uint64_t *bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t smem_addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_addr));
  return v;
}

// shared-memory row pitch (pixels) of the plane kernels: the smallest odd number >= W + 1
__host__ __device__ __forceinline__ int walk_pitch(int W) { return (W + 1) | 1; }

// HBM -> shared memory fill of the plane kernels: 4 channel planes (HW floats apart at src)
// land interleaved per pixel, planes4[y * P + x] = (c0, c1, c2, c3), with 4-byte async copies
// (LDGSTS): the interleave happens in flight, nothing is staged in registers and the whole read
// of the CTA is outstanding at once.  A warp takes whole rows; its lanes are (pixel 0..7,
// channel 0..3), so one instruction writes one contiguous 128-byte line of shared memory and
// reads four 32-byte runs, and the addresses advance by constants (no div / wrap per element).
// The caller commits nothing else to the async group and must cp.async.wait_group 0 +
// __syncthreads() before reading.
template <int THREADS>
__device__ __forceinline__ void fill_planes4_async(float4 *planes4, const float *__restrict__ src, int H,
                                                   int W, int P, int HW) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int px = lane >> 2, c = lane & 3;
  const float *sc = src + (size_t)c * HW + px;
  const uint32_t d0 = smem_u32(planes4) + 16u * (uint32_t)px + 4u * (uint32_t)c;
  for (int y = warp; y < H; y += THREADS / 32) {
    const float *s = sc + (size_t)y * W;
    uint32_t d = d0 + 16u * (uint32_t)(y * P);
    for (int x = px; x < W; x += 8, s += 8, d += 128u)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(s) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// The same fill from a window of a larger map: rows x cols pixels whose first is src, rows row_stride floats
// apart, channel planes chan_stride floats apart.
template <int THREADS>
__device__ __forceinline__ void fill_planes4_window_async(float4 *planes4, const float *__restrict__ src, int rows,
                                                          int cols, int P, int row_stride, size_t chan_stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int px = lane >> 2, c = lane & 3;
  const float *sc = src + (size_t)c * chan_stride + px;
  const uint32_t d0 = smem_u32(planes4) + 16u * (uint32_t)px + 4u * (uint32_t)c;
  for (int y = warp; y < rows; y += THREADS / 32) {
    const float *s = sc + (size_t)y * row_stride;
    uint32_t d = d0 + 16u * (uint32_t)(y * P);
    for (int x = px; x < cols; x += 8, s += 8, d += 128u)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(s) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// The window fill with 128-bit loads, for windows whose rows start 16-byte aligned in every channel plane (map
// width, window origin and the map's base all multiples of 4 floats): a warp takes blocks of 8 rows x 16 pixels,
// a lane 4 consecutive pixels of a row in all 4 channels (four coalesced LDG.128, 64 contiguous bytes per row and
// channel over 4 lanes), transposes them in registers and stores one float4 (4 channels) per pixel, scaled by
// `pre`.  4-byte async copies move one element per thread and instruction: ~9 us for a 61 x 61 x 4 window,
// against ~3.5 us for the bulk-copy fill of a whole plane set of that size.  Synchronous: the caller needs a
// __syncthreads() before reading, no cp.async wait.
template <int THREADS>
__device__ __forceinline__ void fill_planes4_window_vec(float4 *planes4, const float *__restrict__ src, int rows,
                                                        int cols, int P, int row_stride, size_t chan_stride,
                                                        float pre) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ly = lane >> 2, lx = (lane & 3) * 4;
  const int rblks = (rows + 7) >> 3, cblks = (cols + 15) >> 4;
  for (int blk = warp; blk < rblks * cblks; blk += THREADS / 32) {
    const int rb = blk / cblks, cb = blk - rb * cblks;
    const int y = rb * 8 + ly, x = cb * 16 + lx;
    if (y >= rows || x >= cols) continue;
    const float *s = src + (size_t)y * row_stride + x;
    float4 *d = planes4 + y * P + x;
    if (x + 3 < cols) {
      const float4 a = __ldg(reinterpret_cast<const float4 *>(s));
      const float4 b = __ldg(reinterpret_cast<const float4 *>(s + chan_stride));
      const float4 c = __ldg(reinterpret_cast<const float4 *>(s + 2 * chan_stride));
      const float4 e = __ldg(reinterpret_cast<const float4 *>(s + 3 * chan_stride));
      d[0] = make_float4(pre * a.x, pre * b.x, pre * c.x, pre * e.x);
      d[1] = make_float4(pre * a.y, pre * b.y, pre * c.y, pre * e.y);
      d[2] = make_float4(pre * a.z, pre * b.z, pre * c.z, pre * e.z);
      d[3] = make_float4(pre * a.w, pre * b.w, pre * c.w, pre * e.w);
    } else {
      for (int j = 0; x + j < cols; ++j)
        d[j] = make_float4(pre * __ldg(s + j), pre * __ldg(s + chan_stride + j), pre * __ldg(s + 2 * chan_stride + j),
                           pre * __ldg(s + 3 * chan_stride + j));
    }
  }
}

// The same fill from a channels-last (NHWC) feature map: the 4 channels of a pixel are already one
// contiguous, 16-byte aligned float4 (C % 4 == 0), C floats from the next pixel -- one 16-byte async
// copy per pixel, no interleave at all.  A 32-byte DRAM sector holds two chunks' worth, so the CTAs
// of neighbouring chunks (they run side by side) share every sector through L2.
template <int THREADS>
__device__ __forceinline__ void fill_planes4_nhwc_async(float4 *planes4, const float *__restrict__ src, int H,
                                                        int W, int P, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int y = warp; y < H; y += THREADS / 32) {
    const float *s = src + (size_t)y * W * C;
    const uint32_t d = smem_u32(planes4 + y * P);
    for (int x = lane; x < W; x += 32)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * (uint32_t)x), "l"(s + (size_t)x * C)
                   : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// streaming (evict-first) 128-bit store / load: outputs and one-shot inputs must not push
// the feature planes out of L2
__device__ __forceinline__ void st_stream4(float *p, float4 v) {
  __stcs(reinterpret_cast<float4 *>(p), v);
}
__device__ __forceinline__ float4 ld_stream4(const float *p) {
  return __ldcs(reinterpret_cast<const float4 *>(p));
}

}  // namespace rlod
