// SimpleConv(aggr="mean") + residual + ReLU, forward and backward (sm_100a).
//
// Replaces `self.mean_conv(x, edge_index) + x_0` followed by `F.relu`
// (/root/reference/gnn_pressure_estimation/GraphModels.py:466-467; semantics in
// SURVEY.md §A.3).  Works on the shared CSR minus each row's trailing self-loop
// (gatres_csr_build puts it last), so in-degree = row length - 1.
#include "common.cuh"

namespace gatres {

template <int C>
__global__ void __launch_bounds__(kThreads)
mean_res_fwd_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                    const float* __restrict__ z, const float* __restrict__ x0, float* __restrict__ out,
                    unsigned M, unsigned N) {
  using RM = RowMap<1, C>;
  constexpr int LPR = RM::LPR, RPW = RM::RPW;
  static_assert(RM::V == 1, "row wider than one warp pass");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, lig = lane % LPR;
  constexpr unsigned rows_per_cta = kWarps * RPW;
  pdl_wait();
  if ((unsigned long long)gridDim.x * rows_per_cta >= M) pdl_launch_dependents();   // single pass: see common.cuh
  for (unsigned r0 = blockIdx.x * rows_per_cta; r0 < M; r0 += gridDim.x * rows_per_cta) {
    const unsigned r = r0 + warp * RPW + sub;
    if (r >= M) continue;
    const unsigned b = r / N, i = r - b * N;
    const size_t base = (size_t)b * N;
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1) - 1;     // drop the self-loop
    float4 acc = f4zero();
#pragma unroll 4
    for (int e = beg; e < end; ++e) add4(acc, ldg4(z + (base + __ldg(col + e)) * C + 4 * lig));
    const int deg = end - beg;
    const float inv = 1.f / (float)(deg > 1 ? deg : 1);
    const float4 xr = ldg4_stream(x0 + (size_t)r * C + 4 * lig);
    float4 o;
    o.x = fmaxf(fmaf(acc.x, inv, xr.x), 0.f);
    o.y = fmaxf(fmaf(acc.y, inv, xr.y), 0.f);
    o.z = fmaxf(fmaf(acc.z, inv, xr.z), 0.f);
    o.w = fmaxf(fmaf(acc.w, inv, xr.w), 0.f);
    st4(out + (size_t)r * C + 4 * lig, o);
  }
}

// dz[j] = sum over out-edges j->i (self-loop excluded) of gm[i] / max(indeg(i),1),
// gm = g_out * (out > 0) when `out` is given (then dres = gm is also written),
// gm = g_out when `out` is NULL (caller already applied the ReLU mask).
template <int C>
__global__ void __launch_bounds__(kThreads)
mean_res_bwd_kernel(const int* __restrict__ rowptr, const int* __restrict__ rowptr_t,
                    const int* __restrict__ col_t, const float* __restrict__ g_out,
                    const float* __restrict__ out, float* __restrict__ dz, float* __restrict__ dres,
                    unsigned M, unsigned N) {
  using RM = RowMap<1, C>;
  constexpr int LPR = RM::LPR, RPW = RM::RPW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, lig = lane % LPR;
  constexpr unsigned rows_per_cta = kWarps * RPW;
  pdl_wait();
  if ((unsigned long long)gridDim.x * rows_per_cta >= M) pdl_launch_dependents();   // single pass: see common.cuh
  for (unsigned r0 = blockIdx.x * rows_per_cta; r0 < M; r0 += gridDim.x * rows_per_cta) {
    // rows past the end are clamped (they redo the last row and store nothing): warp-uniform shuffles below
    const unsigned r_raw = r0 + warp * RPW + sub;
    const bool row_ok = r_raw < M;
    const unsigned r = row_ok ? r_raw : M - 1;
    const unsigned b = r / N, jn = r - b * N;
    const size_t base = (size_t)b * N;
    const int beg = __ldg(rowptr_t + jn), deg = __ldg(rowptr_t + jn + 1) - 1 - beg;      // out-edges minus the self-loop
    const int deg_max = __reduce_max_sync(0xffffffffu, deg);
    float4 acc = f4zero();
    for (int e0 = 0; e0 < deg_max; e0 += LPR) {
      // lane t of the row's group resolves edge e0 + t (target id and 1 / in-degree of the target) — the index chain
      // col_t -> rowptr -> rowptr is walked once per edge in parallel instead of once per edge in sequence
      const bool valid = e0 + lig < deg;
      const int it = valid ? __ldg(col_t + beg + e0 + lig) : 0;
      const int dg = valid ? __ldg(rowptr + it + 1) - __ldg(rowptr + it) - 1 : 1;
      const float inv_t = valid ? 1.f / (float)(dg > 1 ? dg : 1) : 0.f;
      const int cnt_max = min(LPR, deg_max - e0);
      for (int t = 0; t < cnt_max; ++t) {
        const int i = __shfl_sync(0xffffffffu, it, t, LPR);
        const float inv = __shfl_sync(0xffffffffu, inv_t, t, LPR);              // 0 beyond this row's edges
        float4 gv = ldg4(g_out + (base + i) * C + 4 * lig);
        if (out != nullptr) {
          const float4 ov = ldg4(out + (base + i) * C + 4 * lig);
          gv.x = ov.x > 0.f ? gv.x : 0.f; gv.y = ov.y > 0.f ? gv.y : 0.f;
          gv.z = ov.z > 0.f ? gv.z : 0.f; gv.w = ov.w > 0.f ? gv.w : 0.f;
        }
        fma4(acc, inv, gv);
      }
    }
    if (!row_ok) continue;
    st4(dz + (size_t)r * C + 4 * lig, acc);
    if (out != nullptr && dres != nullptr) {
      float4 gv = ldg4(g_out + (size_t)r * C + 4 * lig);
      const float4 ov = ldg4(out + (size_t)r * C + 4 * lig);
      gv.x = ov.x > 0.f ? gv.x : 0.f; gv.y = ov.y > 0.f ? gv.y : 0.f;
      gv.z = ov.z > 0.f ? gv.z : 0.f; gv.w = ov.w > 0.f ? gv.w : 0.f;
      st4(dres + (size_t)r * C + 4 * lig, gv);
    }
  }
}

template <int C>
static unsigned row_grid(unsigned M) {
  constexpr unsigned rows_per_cta = kWarps * RowMap<1, C>::RPW;
  unsigned grid = (M + rows_per_cta - 1) / rows_per_cta;
  const unsigned cap = (unsigned)sm_count() * 32u;
  return grid > cap ? cap : grid;
}

}  // namespace gatres

using namespace gatres;

extern "C" int gatres_mean_res_fwd(const int32_t* rowptr, const int32_t* col, const float* z, const float* x0,
                                   float* out, int64_t B, int32_t N, int32_t C, void* stream) {
  GATRES_REQUIRE(B >= 0 && N > 0, "mean_res_fwd: bad B=%lld N=%d", (long long)B, N);
  GATRES_REQUIRE(B * (int64_t)N < (1ll << 31), "mean_res_fwd: B*N must be < 2^31 rows");
  if (B == 0) return GATRES_OK;
  const unsigned M = (unsigned)(B * N);
  cudaStream_t st = as_stream(stream);
  switch (C) {
    case 32: launch_kernel(mean_res_fwd_kernel<32>, dim3(row_grid<32>(M)), dim3(kThreads), 0, st, rowptr, col, z, x0, out, M, N); break;
    case 64: launch_kernel(mean_res_fwd_kernel<64>, dim3(row_grid<64>(M)), dim3(kThreads), 0, st, rowptr, col, z, x0, out, M, N); break;
    case 128: launch_kernel(mean_res_fwd_kernel<128>, dim3(row_grid<128>(M)), dim3(kThreads), 0, st, rowptr, col, z, x0, out, M, N); break;
    default: set_error("mean_res_fwd: unsupported channels %d (32, 64, 128)", C); return GATRES_ERR_ARG;
  }
  return check_launch("mean_res_fwd");
}

extern "C" int gatres_mean_res_bwd(const int32_t* rowptr, const int32_t* rowptr_t, const int32_t* col_t,
                                   const float* g_out, const float* out, float* dz, float* dres,
                                   int64_t B, int32_t N, int32_t C, void* stream) {
  GATRES_REQUIRE(B >= 0 && N > 0, "mean_res_bwd: bad B=%lld N=%d", (long long)B, N);
  GATRES_REQUIRE(B * (int64_t)N < (1ll << 31), "mean_res_bwd: B*N must be < 2^31 rows");
  if (B == 0) return GATRES_OK;
  const unsigned M = (unsigned)(B * N);
  cudaStream_t st = as_stream(stream);
  switch (C) {
    case 32: launch_kernel(mean_res_bwd_kernel<32>, dim3(row_grid<32>(M)), dim3(kThreads), 0, st, rowptr, rowptr_t, col_t, g_out, out, dz, dres, M, N); break;
    case 64: launch_kernel(mean_res_bwd_kernel<64>, dim3(row_grid<64>(M)), dim3(kThreads), 0, st, rowptr, rowptr_t, col_t, g_out, out, dz, dres, M, N); break;
    case 128: launch_kernel(mean_res_bwd_kernel<128>, dim3(row_grid<128>(M)), dim3(kThreads), 0, st, rowptr, rowptr_t, col_t, g_out, out, dz, dres, M, N); break;
    default: set_error("mean_res_bwd: unsupported channels %d (32, 64, 128)", C); return GATRES_ERR_ARG;
  }
  return check_launch("mean_res_bwd");
}
