// Shared device/host helpers for libgatres_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/gatres_b200.h"

#ifndef __CUDA_ARCH__
#define GATRES_HOST_PASS 1
#endif
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgatres_b200 is written for sm_100a (B200); compile with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace gatres {

constexpr float kNegSlope = 0.2f;      // GATConv negative_slope default
constexpr float kSoftmaxEps = 1e-16f;  // torch_geometric.utils.softmax denominator epsilon
constexpr int kThreads = 256;          // CTA size of the row kernels
constexpr int kWarps = kThreads / 32;

void set_error(const char* fmt, ...);
int check_launch(const char* what);
int sm_count();

#define GATRES_REQUIRE(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      gatres::set_error(__VA_ARGS__);        \
      return GATRES_ERR_ARG;                 \
    }                                        \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (PDL) --------------------------------------
// Every kernel of the step is launched with the programmatic-stream-serialization
// attribute: it may start (and run its prologue: smem carve-up, weight / CSR staging,
// barrier init) while its predecessor is still draining, and blocks in pdl_wait()
// until the predecessor's memory is visible.  Only data that is constant within a
// step (parameters, CSR) may be touched before pdl_wait().
// A kernel releases its dependents early (pdl_launch_dependents) only when every one
// of its CTAs makes a single pass (small batches, where launch latency matters).  With
// multi-pass persistent grids an early release parks the next kernel's CTAs on whichever
// SMs drain first; its static work split then runs unbalanced (measured: dgrad -> wgrad
// chain 228 us -> 410 us at 2048 snapshots), so large launches keep plain serialization.
bool pdl_enabled();
void count_launch();                    // bumps the counter behind gatres_launch_count()
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline void launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  count_launch();
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);      // errors surface in check_launch()
}
#endif

// ------------------------------------------------------------------ device
__device__ __forceinline__ float lrelu(float z) { return z > 0.f ? z : kNegSlope * z; }
__device__ __forceinline__ float lrelu_slope(float z) { return z > 0.f ? 1.f : kNegSlope; }

// read-only 128-bit load (L1-allocating: neighbour rows are re-read by adjacent rows)
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming 128-bit load: data touched once by this kernel, keep it out of L1
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ void fma4(float4& acc, float s, float4 v) {
  acc.x = fmaf(s, v.x, acc.x);
  acc.y = fmaf(s, v.y, acc.y);
  acc.z = fmaf(s, v.z, acc.z);
  acc.w = fmaf(s, v.w, acc.w);
}
__device__ __forceinline__ void add4(float4& acc, float4 v) {
  acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
}

// butterfly sum over aligned groups of `width` lanes (width power of two <= 32);
// `mask` must name every lane of the groups that execute this call together.
template <int width>
__device__ __forceinline__ float group_sum(float v, unsigned mask) {
#pragma unroll
  for (int off = width / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off);
  return v;
}

// ---- row <-> lane mapping of the aggregation kernels -----------------------
// A feature row of F = H*C floats is F/4 float4 "chunks".  LPR lanes share one
// row (one chunk each, or V chunks each when F/4 > 32), so a warp works on
// RPW = 32/LPR rows at a time.  Chunk q covers floats [4q, 4q+4) and belongs to
// head 4q / C.
template <int H, int C, bool PACK = false>
struct RowMap {
  static constexpr int F = H * C;
  static constexpr int CHUNKS = F / 4;
  // PACK = false: one chunk per lane while the row fits a warp (two heads of 32 channels = 16 lanes per row).
  // PACK = true : lanes per row = lanes per head; a lane holds the same chunk of EVERY head (V = H chunks), so the
  //   index work of a row (row pointers, neighbour ids, their shuffles) is shared by the heads and a warp covers
  //   32 / LPR rows whatever H is.  Measured (B200, 2048 snapshots): forward snapshot-tile kernel 92.7 -> 86.6 us,
  //   resident kernels -27 % on the two-head aggregation; the backward tile kernel (64-register budget at 1024
  //   threads) and the gather kernels on the 100 000-node graph get slower, so they keep PACK = false.
  static constexpr int LPR = PACK ? ((C / 4) < 32 ? (C / 4) : 32) : (CHUNKS < 32 ? CHUNKS : 32);   // lanes per row
  static constexpr int V = CHUNKS / LPR;                  // chunks per lane
  static constexpr int RPW = 32 / LPR;                    // rows per warp
  static constexpr int LPH = (C / 4) < 32 ? (C / 4) : 32; // lanes holding one head of one row
  static_assert(F % 4 == 0 && CHUNKS % LPR == 0 && (LPR & (LPR - 1)) == 0, "unsupported row width");
  static_assert(V == 1 || V == H, "with several chunks per lane each chunk must be its own head");
  __device__ static __forceinline__ int chunk(int lig, int v) { return lig + v * LPR; }
  __device__ static __forceinline__ int head(int lig, int v) { return (4 * (lig + v * LPR)) / C; }
};

// Sum per-lane float4 accumulators over all lanes of a CTA that hold the same
// chunk (same lane-in-row, any row slot, any warp) and write them to
// dst[4*chunk ..].  red: shared scratch of kWarps*32*4 floats.  All threads call.
// Parameter-gradient emission: either a plain store into this CTA's row of the
// `partial` buffer (deterministic two-stage reduction) or an atomic add into the
// final gradient buffer (any grid size; summation order not fixed).
__device__ __forceinline__ void emit4(float* dst, float4 v, bool atomic) {
  if (atomic) atomicAdd(reinterpret_cast<float4*>(dst), v);      // sm_90+: 128-bit red.global.add
  else st4(dst, v);
}
__device__ __forceinline__ void emit1(float* dst, float v, bool atomic) {
  if (atomic) atomicAdd(dst, v);
  else *dst = v;
}

template <int LPR>
__device__ __forceinline__ void cta_chunk_sum_store(float4 acc, float* red, float* dst, int chunk_of_lig0_stride_v,
                                                    bool atomic) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  st4(red + (warp * 32 + lane) * 4, acc);
  __syncthreads();
  if (threadIdx.x < LPR) {
    float4 s = f4zero();
    for (int w = 0; w < kWarps; ++w)
#pragma unroll
      for (int sub = 0; sub < 32 / LPR; ++sub) add4(s, *reinterpret_cast<float4*>(red + (w * 32 + sub * LPR + threadIdx.x) * 4));
    emit4(dst + 4 * (threadIdx.x + chunk_of_lig0_stride_v), s, atomic);
  }
}

// grid of a row-parallel kernel: enough CTAs to cover the rows, capped at a few waves
static inline unsigned row_kernel_grid(unsigned M, unsigned rows_per_cta, unsigned ctas_per_sm) {
  unsigned grid = (M + rows_per_cta - 1) / rows_per_cta;
  const unsigned cap = (unsigned)sm_count() * ctas_per_sm;
  return grid > cap ? cap : (grid < 1 ? 1 : grid);
}

}  // namespace gatres
