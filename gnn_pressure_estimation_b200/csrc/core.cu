// Error reporting, device queries and the caller-side fused kernels
// (mask, masked MSE, Adam) of libgatres_b200.so.
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

namespace gatres {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return GATRES_OK;
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return GATRES_ERR_CUDA;
}

static long long g_launches = 0;
void count_launch() { ++g_launches; }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GATRES_PDL");
    v = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return 148;   // B200
  }
  cached = n;
  return n;
}

// x_masked = mask ? 0 : x   (train.py:174  data.x[batch_mask] = 0)
__global__ void __launch_bounds__(256)
apply_mask_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, float* __restrict__ out, size_t M) {
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < M; i += (size_t)gridDim.x * blockDim.x)
    out[i] = mask[i] ? 0.f : x[i];
}

// d_out = 2 (out - y) mask / count ; per-CTA partial of sum((out-y)^2 mask)
__global__ void __launch_bounds__(256)
masked_mse_kernel(const float* __restrict__ out, const float* __restrict__ y, const uint8_t* __restrict__ mask,
                  size_t M, float inv_count, float* __restrict__ d_out, float* __restrict__ partial_loss) {
  __shared__ float red[kWarps];
  pdl_wait();
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < M; i += (size_t)gridDim.x * blockDim.x) {
    const float d = mask[i] ? out[i] - y[i] : 0.f;
    s = fmaf(d, d, s);
    d_out[i] = 2.f * d * inv_count;
  }
  s = group_sum<32>(s, 0xffffffffu);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < kWarps; ++k) t += red[k];
    partial_loss[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256)
final_loss_kernel(const float* __restrict__ partial_loss, int n, float inv_count, float* __restrict__ loss) {
  __shared__ float red[kWarps];
  pdl_wait();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial_loss[i];
  s = group_sum<32>(s, 0xffffffffu);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < kWarps; ++k) t += red[k];
    loss[0] = t * inv_count;
  }
}

__global__ void bump_step_kernel(int* step) {
  pdl_wait();
  step[0] += 1;
}

// torch.optim.Adam, single-tensor formulation (no amsgrad, L2 weight decay)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            const int* __restrict__ step, size_t P, float lr, float b1, float b2, float eps, float wd,
            float grad_scale) {
  pdl_wait();
  const float t = (float)__ldg(step);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < P; i += (size_t)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i] * grad_scale);
    const float mi = fmaf(1.f - b1, gi - m[i], m[i]);            // lerp, as torch does
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
  }
}

// ---- gradient all-reduce fused into the Adam step, over NVLink peer memory -------------------------------------
// Data-parallel training needs exactly one collective per step: the SUM of the flat gradient buffer.
// Instead of all-reduce -> Adam (two passes over the gradients plus a collective launch), every rank's Adam kernel
// reads the gradient buffers of ALL ranks directly through peer pointers (symmetric memory over NVLink / NVSwitch),
// sums them in rank order (so every replica computes bit-identical updates) and applies the update: the collective
// costs one cross-GPU flag barrier plus (world - 1) x P floats of peer loads inside a kernel that had to run anyway.
//   flags[r][q] (uint32, one array per rank r, symmetric) = last epoch for which rank q announced "my backward is done".
// The gradient buffers are double-buffered by the caller (step parity), so one barrier per step also orders the
// re-zeroing of a buffer after every peer has finished reading it.
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys4(const float* p) {
  float4 r;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}

constexpr int kMaxPeers = 16;
struct PeerTable {
  const float* grads[kMaxPeers];
  unsigned* flags[kMaxPeers];
};

// Runs right after the last backward kernel of the step: bumps the Adam step counter and the barrier epoch and
// ANNOUNCES "this rank's gradients of epoch e are complete" in every peer's flag array — one launch gap earlier than
// the Adam kernel could, so the flags travel over NVLink while the Adam kernel is being launched and loads its state.
__global__ void __launch_bounds__(32)
bump_publish_kernel(int* step, int* epoch, const PeerTable peers, int rank, int world) {
  pdl_wait();                                                   // the backward kernels' writes are visible
  const unsigned e = (unsigned)epoch[0] + 1u;
  __syncwarp();
  if (threadIdx.x == 0) {
    step[0] += 1;
    epoch[0] = (int)e;
  }
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(peers.flags[threadIdx.x] + rank, e);
  }
}

// sum of the `world` ranks' gradient chunk i in rank order; the peer loads are issued together (eight in flight per
// thread) — a rank-by-rank loop would pay one NVLink round trip per peer
__device__ __forceinline__ float4 peer_sum4(const PeerTable& peers, size_t i, int world) {
  float4 acc = f4zero();
  for (int r0 = 0; r0 < world; r0 += 8) {
    float4 g[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (r0 + u < world) g[u] = ld_relaxed_sys4(peers.grads[r0 + u] + 4 * i);
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (r0 + u < world) add4(acc, g[u]);
  }
  return acc;
}

__global__ void __launch_bounds__(256)
adam_peer_kernel(float* __restrict__ p, const PeerTable peers, float* __restrict__ m, float* __restrict__ v,
                 const int* __restrict__ step, const int* __restrict__ epoch_dev, size_t P4, size_t P, float lr, float b1,
                 float b2, float eps, float wd, float grad_scale, int rank, int world) {
  pdl_wait();
  const unsigned epoch = (unsigned)__ldg(epoch_dev);
  // everything that does not depend on the peers is loaded BEFORE the flag wait: this thread's first chunk of the
  // parameters and moments, the step count and the bias corrections
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  float4 pv0 = f4zero(), mv0 = f4zero(), vv0 = f4zero();
  if (i0 < P4) {
    pv0 = *reinterpret_cast<const float4*>(p + 4 * i0);
    mv0 = *reinterpret_cast<const float4*>(m + 4 * i0);
    vv0 = *reinterpret_cast<const float4*>(v + 4 * i0);
  }
  const float t = (float)__ldg(step);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  if (threadIdx.x < world) {
    const unsigned* f = peers.flags[rank] + threadIdx.x;        // own (local) flag array; peers write into it
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < epoch)
      if (clock64() - t0 > (240ll << 30)) __trap();             // ~2 min at 2 GHz: a lost peer must not hang the GPU forever
  }
  __syncthreads();
  auto update = [&](float pi, float g, float& mi, float& vi) {
    const float gi = fmaf(wd, pi, g * grad_scale);
    mi = fmaf(1.f - b1, gi - mi, mi);
    vi = fmaf(b2, vi, (1.f - b2) * gi * gi);
    return pi - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
  };
  for (size_t i = i0; i < P4; i += stride) {
    const float4 g = peer_sum4(peers, i, world);                // rank order: same sum on every replica
    float4 pv, mv, vv;
    if (i == i0) { pv = pv0; mv = mv0; vv = vv0; }
    else {
      pv = *reinterpret_cast<float4*>(p + 4 * i); mv = *reinterpret_cast<float4*>(m + 4 * i); vv = *reinterpret_cast<float4*>(v + 4 * i);
    }
    pv.x = update(pv.x, g.x, mv.x, vv.x);
    pv.y = update(pv.y, g.y, mv.y, vv.y);
    pv.z = update(pv.z, g.z, mv.z, vv.z);
    pv.w = update(pv.w, g.w, mv.w, vv.w);
    st4(p + 4 * i, pv); st4(m + 4 * i, mv); st4(v + 4 * i, vv);
  }
  if (blockIdx.x == 0) {                                       // tail (P is not a multiple of 4)
    for (size_t i = 4 * P4 + threadIdx.x; i < P; i += blockDim.x) {
      float g = 0.f;
      for (int r = 0; r < world; ++r) {
        float x;
        asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(x) : "l"(peers.grads[r] + i) : "memory");
        g += x;
      }
      float mi = m[i], vi = v[i];
      p[i] = update(p[i], g, mi, vi);
      m[i] = mi;
      v[i] = vi;
    }
  }
}

static unsigned flat_grid(size_t n, unsigned cap_per_sm) {
  size_t g = (n + 255) / 256;
  const size_t cap = (size_t)sm_count() * cap_per_sm;
  return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace gatres

using namespace gatres;

extern "C" int gatres_abi_version(void) { return GATRES_ABI_VERSION; }
extern "C" const char* gatres_last_error(void) { return g_err; }
extern "C" int gatres_sm_count(void) { return sm_count(); }
extern "C" int64_t gatres_launch_count(void) { return g_launches; }

extern "C" int gatres_apply_mask(const float* x, const uint8_t* mask, float* x_masked, int64_t M, void* stream) {
  GATRES_REQUIRE(M >= 0, "apply_mask: bad M");
  if (M == 0) return GATRES_OK;
  launch_kernel(apply_mask_kernel, dim3(flat_grid((size_t)M, 8)), dim3(256), 0, as_stream(stream), x, mask, x_masked, (size_t)M);
  return check_launch("apply_mask");
}

extern "C" int gatres_masked_mse(const float* out, const float* y, const uint8_t* mask, int64_t M, int64_t count,
                                 float* d_out, float* loss_out, float* partial_loss, void* stream) {
  GATRES_REQUIRE(M > 0 && count > 0, "masked_mse: bad M=%lld count=%lld", (long long)M, (long long)count);
  unsigned grid = flat_grid((size_t)M, 4);
  if (grid > 1024) grid = 1024;
  const float inv = 1.f / (float)count;
  launch_kernel(masked_mse_kernel, dim3(grid), dim3(256), 0, as_stream(stream), out, y, mask, (size_t)M, inv, d_out, partial_loss);
  int rc = check_launch("masked_mse");
  if (rc) return rc;
  launch_kernel(final_loss_kernel, dim3(1), dim3(256), 0, as_stream(stream), partial_loss, (int)grid, inv, loss_out);
  return check_launch("masked_mse_final");
}

extern "C" int gatres_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                int32_t* step_count, int64_t P, float lr, float beta1, float beta2, float eps,
                                float weight_decay, float grad_scale, void* stream) {
  GATRES_REQUIRE(P > 0, "adam_step: bad P");
  launch_kernel(bump_step_kernel, dim3(1), dim3(1), 0, as_stream(stream), step_count);
  int rc = check_launch("adam_bump");
  if (rc) return rc;
  launch_kernel(adam_kernel, dim3(flat_grid((size_t)P, 4)), dim3(256), 0, as_stream(stream), params, grads, exp_avg, exp_avg_sq, step_count,
                                                                      (size_t)P, lr, beta1, beta2, eps, weight_decay,
                                                                      grad_scale);
  return check_launch("adam_step");
}

extern "C" int gatres_adam_step_peer(float* params, const float* const* peer_grads, uint32_t* const* peer_flags,
                                     int32_t rank, int32_t world, float* exp_avg, float* exp_avg_sq, int32_t* step_count,
                                     int32_t* epoch, int64_t P, float lr, float beta1, float beta2, float eps,
                                     float weight_decay, float grad_scale, void* stream) {
  GATRES_REQUIRE(P > 0 && world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world,
                 "adam_step_peer: bad P=%lld rank=%d world=%d (at most %d peers)", (long long)P, rank, world, kMaxPeers);
  GATRES_REQUIRE(peer_grads && peer_flags && epoch && step_count, "adam_step_peer: null table");
  PeerTable t = {};
  for (int r = 0; r < world; ++r) {
    GATRES_REQUIRE(peer_grads[r] != nullptr && peer_flags[r] != nullptr, "adam_step_peer: null pointer for rank %d", r);
    GATRES_REQUIRE((reinterpret_cast<uintptr_t>(peer_grads[r]) & 15) == 0, "adam_step_peer: gradient buffers must be 16-byte aligned");
    t.grads[r] = peer_grads[r];
    t.flags[r] = peer_flags[r];
  }
  launch_kernel(bump_publish_kernel, dim3(1), dim3(32), 0, as_stream(stream), step_count, epoch, t, (int)rank, (int)world);
  int rc = check_launch("adam_peer_bump");
  if (rc) return rc;
  const size_t P4 = (size_t)P / 4;
  launch_kernel(adam_peer_kernel, dim3(flat_grid(P4 ? P4 : 1, 2)), dim3(256), 0, as_stream(stream), params, t, exp_avg, exp_avg_sq,
                (const int*)step_count, (const int*)epoch, P4, (size_t)P, lr, beta1, beta2, eps, weight_decay, grad_scale,
                (int)rank, (int)world);
  return check_launch("adam_step_peer");
}
