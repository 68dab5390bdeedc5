// Caller-side kernels next to the hot path: the per-snapshot exact-count random mask
// and the seven training / evaluation metrics of the reference loop, on the device.
//
//   mask      /root/reference/gnn_pressure_estimation/utils/auxil.py:143-182 (mask_nodes / generate_batch_mask,
//             called per batch at train.py:171-172): exactly int(N * rate) nodes per snapshot, uniformly at
//             random without replacement.  The reference draws them on the host with the global NumPy RNG and
//             ships a bool mask to the GPU every step; here every node gets a counter-based 32-bit random key
//             and the `count` smallest keys of a snapshot are selected with a 4-pass radix select (ties by
//             node index), so the set is uniform, exact-count, reproducible from (seed, step) and never leaves
//             the device.  Nodes flagged in `required` (the sensors of evaluation.py:288-291) take key 0 and
//             are therefore always in the set, as mask_nodes(..., required_idx) guarantees.  (A host mask is still
//             accepted everywhere: NumPy-compatible mode.)
//   metrics   utils/auxil.py:101-140,185-203 applied as in train.py:177-198 / evaluation.py:326-338: relative
//             error, accuracy@threshold, correlation, R2, MAE, RMSE, NSE over the DESCALED predictions and
//             targets of the masked nodes.  Two passes (means first, then centred sums) with fp64 accumulators,
//             because descaled pressures have mean >> std and the reference centres before squaring.
#include "common.cuh"

namespace gatres {

// ------------------------------------------------------------------------------ random keys
__host__ __device__ __forceinline__ uint32_t mask_key(uint64_t seed, uint64_t step, uint64_t row) {
  uint64_t z = (seed ^ (step * 0xD1B54A32D192ED03ull)) + (row + 1) * 0x9E3779B97F4A7C15ull;   // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const uint32_t k = (uint32_t)(z >> 32);
  return k != 0 ? k : 1u;                 // key 0 is reserved for the nodes that must be masked
}

// One CTA per snapshot.  Keys are recomputed from the counter in every pass (cheaper than staging them, and it
// keeps the kernel independent of the graph size).
__global__ void __launch_bounds__(256)
generate_mask_kernel(uint64_t seed, uint64_t step0, const int* __restrict__ step_dev,
                     const uint8_t* __restrict__ required, unsigned N, unsigned count, uint8_t* __restrict__ mask) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_need, s_base;
  __shared__ unsigned warp_tot[8];
  pdl_wait();
  const uint64_t step = step0 + (step_dev != nullptr ? (uint64_t)(unsigned)__ldg(step_dev) : 0ull);
  const uint64_t row0 = (uint64_t)blockIdx.x * N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_prefix = 0; s_need = count; }
  // radix select, most significant byte first: after pass p the top 8(4-p) bits of the count-th smallest key are known
  for (int p = 3; p >= 0; --p) {
    hist[tid] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix, hi_mask = p == 3 ? 0u : (0xffffffffu << (8 * (p + 1)));
    for (unsigned i = tid; i < N; i += 256) {
      const unsigned k = (required != nullptr && required[i]) ? 0u : mask_key(seed, step, row0 + i);
      if ((k & hi_mask) == prefix) atomicAdd(&hist[(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    if (warp == 0) {
      // bin b* = first bin whose inclusive cumulative count reaches `need`
      unsigned need = s_need, run = 0, found = 0xffffffffu, before = 0;
      for (int c = 0; c < 8; ++c) {
        const unsigned v = hist[c * 32 + lane];
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const unsigned cum = run + inc;
        const unsigned hit = __ballot_sync(0xffffffffu, cum >= need);
        if (found == 0xffffffffu && hit != 0) {
          const int l = __ffs(hit) - 1;
          found = c * 32 + l;
          before = __shfl_sync(0xffffffffu, cum - v, l);
        }
        run = __shfl_sync(0xffffffffu, cum, 31);
      }
      if (lane == 0) {
        s_prefix = prefix | (found << (8 * p));
        s_need = need - before;
      }
    }
    __syncthreads();
  }
  // keys < K are in; of the keys == K the first `need` in node order are in (block-wide scan over 256-node tiles)
  const unsigned K = s_prefix, need = s_need;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (unsigned i0 = 0; i0 < N; i0 += 256) {
    const unsigned i = i0 + tid;
    const unsigned k = i < N ? ((required != nullptr && required[i]) ? 0u : mask_key(seed, step, row0 + i)) : 0xffffffffu;
    const bool tie = i < N && k == K;
    const unsigned bal = __ballot_sync(0xffffffffu, tie);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    unsigned rank = s_base + __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) rank += warp_tot[w];
    if (i < N) mask[row0 + i] = (k < K || (tie && rank < need)) ? 1 : 0;
    __syncthreads();
    if (tid == 0) {
      unsigned t = 0;
      for (int w = 0; w < 8; ++w) t += warp_tot[w];
      s_base += t;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ metrics
constexpr int kMetricSums = 8;   // pass B: sum|e|, sum e^2, sum rel, n_rel, n_acc, sum vx^2, sum vy^2, sum vx vy

// descale as the reference rounds it: one fp32 multiply, then one fp32 add (auxil.py:58-61), never fused
__device__ __forceinline__ float descale1(float v, float scale, float shift) { return __fadd_rn(__fmul_rn(v, scale), shift); }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int K>
__device__ __forceinline__ void cta_sum_store(double (&acc)[K], double* red /* [8][K] */, double* dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double s = warp_sum_d(acc[k]);
    if (lane == 0) red[warp * K + k] = s;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w * K + threadIdx.x];
    dst[threadIdx.x] = s;
  }
}

// pass A: n, sum p, sum t over the selected entries (descaled)
__global__ void __launch_bounds__(256)
metrics_pass_a_kernel(const float* __restrict__ out, const float* __restrict__ y, const uint8_t* __restrict__ mask,
                      size_t M, float scale, float shift, double* __restrict__ part_a) {
  __shared__ double red[8 * 3];
  pdl_wait();
  double acc[3] = {0.0, 0.0, 0.0};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < M; i += (size_t)gridDim.x * blockDim.x) {
    if (mask != nullptr && mask[i] == 0) continue;
    acc[0] += 1.0;
    acc[1] += (double)descale1(out[i], scale, shift);
    acc[2] += (double)descale1(y[i], scale, shift);
  }
  cta_sum_store<3>(acc, red, part_a + (size_t)blockIdx.x * 3);
}

// pass B: every CTA first folds the pass-A partials (a few hundred doubles) into the two means
__global__ void __launch_bounds__(256)
metrics_pass_b_kernel(const float* __restrict__ out, const float* __restrict__ y, const uint8_t* __restrict__ mask,
                      size_t M, float scale, float shift, float threshold, const double* __restrict__ part_a,
                      int blocks_a, double* __restrict__ part_b) {
  __shared__ double red[8 * kMetricSums];
  __shared__ double tot[3];
  pdl_wait();
  {
    double a[3] = {0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < blocks_a; b += blockDim.x) {
      a[0] += part_a[b * 3 + 0]; a[1] += part_a[b * 3 + 1]; a[2] += part_a[b * 3 + 2];
    }
    cta_sum_store<3>(a, red, tot);
    __syncthreads();
  }
  const double n = tot[0], mp = n > 0 ? tot[1] / n : 0.0, mt = n > 0 ? tot[2] / n : 0.0;
  double acc[kMetricSums];
#pragma unroll
  for (int k = 0; k < kMetricSums; ++k) acc[k] = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < M; i += (size_t)gridDim.x * blockDim.x) {
    if (mask != nullptr && mask[i] == 0) continue;
    const float p = descale1(out[i], scale, shift), t = descale1(y[i], scale, shift);
    const float e = fabsf(t - p);                       // auxil.py:115,122
    acc[0] += (double)e;
    acc[1] += (double)(p - t) * (double)(p - t);
    if (fabsf(t) > 0.01f) {                             // auxil.py:116-118
      acc[2] += (double)fabsf(e / t);
      acc[3] += 1.0;
    }
    if (e <= t * threshold) acc[4] += 1.0;              // auxil.py:123
    const double vx = (double)p - mp, vy = (double)t - mt;
    acc[5] += vx * vx;
    acc[6] += vy * vy;
    acc[7] += vx * vy;
  }
  __syncthreads();
  cta_sum_store<kMetricSums>(acc, red, part_b + (size_t)blockIdx.x * kMetricSums);
}

// final: [error, acc, corr, r2, mae, rmse, nse, count]  (order of get_metric_fn_collection, auxil.py:194-202)
__global__ void __launch_bounds__(256)
metrics_final_kernel(const double* __restrict__ part_a, int blocks_a, const double* __restrict__ part_b, int blocks_b,
                     float* __restrict__ metrics_out) {
  __shared__ double red[8 * kMetricSums];
  __shared__ double tot[kMetricSums + 1];
  pdl_wait();
  double n_acc[1] = {0.0};
  for (int b = threadIdx.x; b < blocks_a; b += blockDim.x) n_acc[0] += part_a[b * 3];
  cta_sum_store<1>(n_acc, red, tot + kMetricSums);
  __syncthreads();
  double acc[kMetricSums];
#pragma unroll
  for (int k = 0; k < kMetricSums; ++k) acc[k] = 0.0;
  for (int b = threadIdx.x; b < blocks_b; b += blockDim.x)
#pragma unroll
    for (int k = 0; k < kMetricSums; ++k) acc[k] += part_b[b * kMetricSums + k];
  cta_sum_store<kMetricSums>(acc, red, tot);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n = tot[kMetricSums];
    double corr = tot[7] / (sqrt(tot[5]) * sqrt(tot[6]));                 // auxil.py:131 (nan for constant inputs)
    corr = corr < -1.0 ? -1.0 : (corr > 1.0 ? 1.0 : corr);               // clamp keeps nan, as torch.clamp does
    metrics_out[0] = (float)(tot[2] / tot[3]);                            // mean over the kept entries (nan if none)
    metrics_out[1] = (float)(tot[4] / n);
    metrics_out[2] = (float)corr;
    metrics_out[3] = (float)(corr * corr);
    metrics_out[4] = (float)(tot[0] / n);
    metrics_out[5] = (float)sqrt(tot[1] / n);
    metrics_out[6] = (float)(1.0 - tot[1] / (tot[6] + 1e-12));           // auxil.py:104-107
    metrics_out[7] = (float)n;
  }
}

}  // namespace gatres

using namespace gatres;

extern "C" uint32_t gatres_mask_key(uint64_t seed, uint64_t step, uint64_t row) { return mask_key(seed, step, row); }

extern "C" int gatres_generate_mask(uint64_t seed, uint64_t step, const int32_t* step_dev, const uint8_t* required,
                                    int64_t B, int32_t N, int32_t count, uint8_t* mask, void* stream) {
  GATRES_REQUIRE(B >= 0 && N > 0 && count >= 0 && count <= N, "generate_mask: bad B=%lld N=%d count=%d", (long long)B, N, count);
  GATRES_REQUIRE(B < (1ll << 31) && mask != nullptr, "generate_mask: bad batch or null mask");
  if (B == 0) return GATRES_OK;
  if (count == 0) {
    if (cudaMemsetAsync(mask, 0, (size_t)B * N, as_stream(stream)) != cudaSuccess) return check_launch("generate_mask");
    return GATRES_OK;
  }
  launch_kernel(generate_mask_kernel, dim3((unsigned)B), dim3(256), 0, as_stream(stream), seed, step, step_dev, required,
                (unsigned)N, (unsigned)count, mask);
  return check_launch("generate_mask");
}

extern "C" int64_t gatres_metrics_scratch_doubles(void) { return 2048; }

extern "C" int gatres_masked_metrics(const float* out, const float* y, const uint8_t* mask, int64_t M, float scale,
                                     float shift, float threshold, double* scratch, float* metrics_out, void* stream) {
  GATRES_REQUIRE(M > 0 && out && y && scratch && metrics_out, "masked_metrics: bad M=%lld or null buffer", (long long)M);
  size_t want = ((size_t)M + 255) / 256;
  const unsigned blocks = (unsigned)(want > 128 ? 128 : want);           // 128 * (3 + 8) doubles <= 2048
  double* part_a = scratch;
  double* part_b = scratch + 128 * 3;
  cudaStream_t st = as_stream(stream);
  launch_kernel(metrics_pass_a_kernel, dim3(blocks), dim3(256), 0, st, out, y, mask, (size_t)M, scale, shift, part_a);
  int rc = check_launch("masked_metrics(a)");
  if (rc) return rc;
  launch_kernel(metrics_pass_b_kernel, dim3(blocks), dim3(256), 0, st, out, y, mask, (size_t)M, scale, shift, threshold,
                (const double*)part_a, (int)blocks, part_b);
  rc = check_launch("masked_metrics(b)");
  if (rc) return rc;
  launch_kernel(metrics_final_kernel, dim3(1), dim3(256), 0, st, (const double*)part_a, (int)blocks, (const double*)part_b,
                (int)blocks, metrics_out);
  return check_launch("masked_metrics(final)");
}
