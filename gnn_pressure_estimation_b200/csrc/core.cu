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
