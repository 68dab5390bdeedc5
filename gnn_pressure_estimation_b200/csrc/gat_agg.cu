// Fused GAT aggregation, forward and recompute-based backward (sm_100a).
//
// Replaces GATConv.forward steps 2-6 and their autograd (call sites
// /root/reference/gnn_pressure_estimation/GraphModels.py:464-465; semantics in
// SURVEY.md §A.2/§A.4).  One group of LPR lanes owns one (snapshot, row); each
// lane owns one 128-bit chunk of the row, so every neighbour-row gather is a
// coalesced 16 B/lane load and no per-edge tensor is ever materialised.
#include <math_constants.h>
#include "common.cuh"

namespace gatres {

// ------------------------------------------------------------------ forward
template <int H, int C>
__global__ void __launch_bounds__(kThreads)
gat_agg_fwd_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                   const float* __restrict__ h, const float* __restrict__ s_src,
                   const float* __restrict__ s_dst, const float* __restrict__ bias,
                   float* __restrict__ out, float* __restrict__ m_out, float* __restrict__ l_out,
                   unsigned M, unsigned N, int relu) {
  using RM = RowMap<H, C>;
  constexpr int F = RM::F, V = RM::V, LPR = RM::LPR, RPW = RM::RPW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, lig = lane % LPR;
  constexpr unsigned rows_per_cta = kWarps * RPW;

  float4 bv[V];
#pragma unroll
  for (int v = 0; v < V; ++v) bv[v] = ldg4(bias + 4 * RM::chunk(lig, v));

  for (unsigned r0 = blockIdx.x * rows_per_cta; r0 < M; r0 += gridDim.x * rows_per_cta) {
    const unsigned r = r0 + warp * RPW + sub;
    if (r >= M) continue;                         // no cross-lane traffic in this kernel
    const unsigned b = r / N, i = r - b * N;
    const size_t base = (size_t)b * N;
    const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);

    float sd[V], mx[V], l[V];
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      sd[v] = __ldg(s_dst + (size_t)r * H + RM::head(lig, v));
      mx[v] = -CUDART_INF_F;
      l[v] = 0.f;
      acc[v] = f4zero();
    }
    // pass 1: row max of the LeakyReLU logits (scores only: 4 B per edge per head)
    for (int e = beg; e < end; ++e) {
      const size_t j = base + __ldg(col + e);
#pragma unroll
      for (int v = 0; v < V; ++v)
        mx[v] = fmaxf(mx[v], lrelu(__ldg(s_src + j * H + RM::head(lig, v)) + sd[v]));
    }
    // pass 2: exp, running sum, weighted accumulation of neighbour rows
#pragma unroll 4
    for (int e = beg; e < end; ++e) {
      const size_t j = base + __ldg(col + e);
      const float* hj = h + j * F;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 x = ldg4(hj + 4 * RM::chunk(lig, v));
        const float p = __expf(lrelu(__ldg(s_src + j * H + RM::head(lig, v)) + sd[v]) - mx[v]);
        l[v] += p;
        fma4(acc[v], p, x);
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float inv = 1.f / (l[v] + kSoftmaxEps);
      float4 o;
      o.x = fmaf(acc[v].x, inv, bv[v].x);
      o.y = fmaf(acc[v].y, inv, bv[v].y);
      o.z = fmaf(acc[v].z, inv, bv[v].z);
      o.w = fmaf(acc[v].w, inv, bv[v].w);
      if (relu) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
      }
      st4(out + (size_t)r * F + 4 * RM::chunk(lig, v), o);
      if (m_out != nullptr && (RM::chunk(lig, v) % (C / 4)) == 0) {
        m_out[(size_t)r * H + RM::head(lig, v)] = mx[v];
        l_out[(size_t)r * H + RM::head(lig, v)] = l[v];
      }
    }
  }
}

// --------------------------------------------------------- backward, pass 1
// Per target row i: D_i = sum_e alpha_e dalpha_e and ds_dst[i] = sum_e dz_e, with
// dalpha_e = <g[i], h[j]>.  Emits rec[i,h] = {s_dst, m, 1/(l+eps), D} so pass 2
// needs one 16 B load per (edge, head) for all target-side scalars.
// Also accumulates the bias gradient (column sums of g).
template <int H, int C>
__global__ void __launch_bounds__(kThreads)
gat_agg_bwd_p1_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                      const float* __restrict__ g, const float* __restrict__ h,
                      const float* __restrict__ s_src, const float* __restrict__ s_dst,
                      const float* __restrict__ m, const float* __restrict__ l,
                      float* __restrict__ rec, float* __restrict__ ds_dst,
                      float* __restrict__ partial, long long P, long long off_bias,
                      unsigned M, unsigned N, int atomic) {
  using RM = RowMap<H, C>;
  constexpr int F = RM::F, V = RM::V, LPR = RM::LPR, RPW = RM::RPW, LPH = RM::LPH;
  __shared__ float red[kWarps * 32 * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, lig = lane % LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  constexpr unsigned rows_per_cta = kWarps * RPW;

  float4 bacc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) bacc[v] = f4zero();

  for (unsigned r0 = blockIdx.x * rows_per_cta; r0 < M; r0 += gridDim.x * rows_per_cta) {
    const unsigned r = r0 + warp * RPW + sub;
    if (r < M) {                                  // uniform across the LPR lanes of a row
      const unsigned b = r / N, i = r - b * N;
      const size_t base = (size_t)b * N;
      const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
      float4 gv[V];
      float sd[V], mi[V], il[V], S1[V], S2[V], S3[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int hd = RM::head(lig, v);
        gv[v] = ldg4_stream(g + (size_t)r * F + 4 * RM::chunk(lig, v));
        add4(bacc[v], gv[v]);
        sd[v] = __ldg(s_dst + (size_t)r * H + hd);
        mi[v] = __ldg(m + (size_t)r * H + hd);
        il[v] = 1.f / (__ldg(l + (size_t)r * H + hd) + kSoftmaxEps);
        S1[v] = S2[v] = S3[v] = 0.f;
      }
#pragma unroll 2
      for (int e = beg; e < end; ++e) {
        const size_t j = base + __ldg(col + e);
        const float* hj = h + j * F;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float da = group_sum<LPH>(dot4(gv[v], ldg4(hj + 4 * RM::chunk(lig, v))), gmask);
          const float z = __ldg(s_src + j * H + RM::head(lig, v)) + sd[v];
          const float alpha = __expf(lrelu(z) - mi[v]) * il[v];
          const float sl = lrelu_slope(z);
          S1[v] = fmaf(alpha, da, S1[v]);
          S2[v] = fmaf(alpha * sl, da, S2[v]);
          S3[v] = fmaf(alpha, sl, S3[v]);
        }
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        if ((RM::chunk(lig, v) % (C / 4)) == 0) {
          const size_t o = (size_t)r * H + RM::head(lig, v);
          st4(rec + o * 4, make_float4(sd[v], mi[v], il[v], S1[v]));
          ds_dst[o] = S2[v] - S1[v] * S3[v];
        }
      }
    }
  }
  float* dst = partial + (atomic ? 0 : (size_t)blockIdx.x * P) + off_bias;
#pragma unroll
  for (int v = 0; v < V; ++v) cta_chunk_sum_store<LPR>(bacc[v], red, dst, v * LPR, atomic);
}

// --------------------------------------------------------- backward, pass 2
// Per source row j (out-edge CSR): dh[j] = sum_{e: j->i} alpha_e g[i]
//   + ds_src[j] att_src + ds_dst[j] att_dst,  ds_src[j] = sum_e dz_e,
// dz_e = alpha_e (dalpha_e - D_i) lrelu'(z_e).  Accumulates datt_src / datt_dst.
template <int H, int C>
__global__ void __launch_bounds__(kThreads)
gat_agg_bwd_p2_kernel(const int* __restrict__ rowptr_t, const int* __restrict__ col_t,
                      const float* __restrict__ g, const float* __restrict__ h,
                      const float* __restrict__ s_src, const float* __restrict__ rec,
                      const float* __restrict__ ds_dst,
                      const float* __restrict__ att_src, const float* __restrict__ att_dst,
                      float* __restrict__ dh,
                      float* __restrict__ partial, long long P, long long off_att_src, long long off_att_dst,
                      unsigned M, unsigned N, int atomic) {
  using RM = RowMap<H, C>;
  constexpr int F = RM::F, V = RM::V, LPR = RM::LPR, RPW = RM::RPW, LPH = RM::LPH;
  __shared__ float red[kWarps * 32 * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, lig = lane % LPR;
  const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  constexpr unsigned rows_per_cta = kWarps * RPW;

  float4 as[V], ad[V], accs[V], accd[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    as[v] = ldg4(att_src + 4 * RM::chunk(lig, v));
    ad[v] = ldg4(att_dst + 4 * RM::chunk(lig, v));
    accs[v] = f4zero();
    accd[v] = f4zero();
  }

  for (unsigned r0 = blockIdx.x * rows_per_cta; r0 < M; r0 += gridDim.x * rows_per_cta) {
    const unsigned r = r0 + warp * RPW + sub;
    if (r < M) {
      const unsigned b = r / N, jn = r - b * N;
      const size_t base = (size_t)b * N;
      const int beg = __ldg(rowptr_t + jn), end = __ldg(rowptr_t + jn + 1);
      float4 hv[V], dacc[V];
      float ss[V], dsrc[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        hv[v] = ldg4_stream(h + (size_t)r * F + 4 * RM::chunk(lig, v));
        ss[v] = __ldg(s_src + (size_t)r * H + RM::head(lig, v));
        dacc[v] = f4zero();
        dsrc[v] = 0.f;
      }
#pragma unroll 2
      for (int e = beg; e < end; ++e) {
        const size_t i = base + __ldg(col_t + e);
        const float* gi = g + i * F;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float4 gv = ldg4(gi + 4 * RM::chunk(lig, v));
          const float4 t = ldg4(rec + (i * H + RM::head(lig, v)) * 4);   // {s_dst, m, 1/l, D}
          const float da = group_sum<LPH>(dot4(gv, hv[v]), gmask);
          const float z = ss[v] + t.x;
          const float alpha = __expf(lrelu(z) - t.y) * t.z;
          dsrc[v] = fmaf(alpha * (da - t.w), lrelu_slope(z), dsrc[v]);
          fma4(dacc[v], alpha, gv);
        }
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float dd = __ldg(ds_dst + (size_t)r * H + RM::head(lig, v));
        fma4(dacc[v], dsrc[v], as[v]);
        fma4(dacc[v], dd, ad[v]);
        st4(dh + (size_t)r * F + 4 * RM::chunk(lig, v), dacc[v]);
        fma4(accs[v], dsrc[v], hv[v]);
        fma4(accd[v], dd, hv[v]);
      }
    }
  }
  float* row = partial + (atomic ? 0 : (size_t)blockIdx.x * P);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    cta_chunk_sum_store<LPR>(accs[v], red, row + off_att_src, v * LPR, atomic);
    cta_chunk_sum_store<LPR>(accd[v], red, row + off_att_dst, v * LPR, atomic);
  }
}

// ------------------------------------------------------------------ launch
template <int H, int C>
static int launch_fwd(const int* rowptr, const int* col, const float* h, const float* s_src, const float* s_dst,
                      const float* bias, float* out, float* m, float* l, unsigned M, unsigned N, int relu,
                      cudaStream_t st) {
  constexpr unsigned rows_per_cta = kWarps * RowMap<H, C>::RPW;
  unsigned grid = (M + rows_per_cta - 1) / rows_per_cta;
  const unsigned cap = (unsigned)sm_count() * 32u;
  if (grid > cap) grid = cap;
  gat_agg_fwd_kernel<H, C><<<grid, kThreads, 0, st>>>(rowptr, col, h, s_src, s_dst, bias, out, m, l, M, N, relu);
  return check_launch("gat_agg_fwd");
}

template <int H, int C>
static int launch_bwd(const int* rowptr, const int* col, const int* rowptr_t, const int* col_t, const float* g,
                      const float* h, const float* s_src, const float* s_dst, const float* m, const float* l,
                      const float* att_src, const float* att_dst, float* rec, float* ds_dst, float* dh,
                      float* partial, long long P, int slots, long long off_as, long long off_ad, long long off_b,
                      unsigned M, unsigned N, cudaStream_t st) {
  const int atomic = slots <= 0;
  const unsigned grid = atomic ? row_kernel_grid(M, kWarps * RowMap<H, C>::RPW, 16) : (unsigned)slots;
  gat_agg_bwd_p1_kernel<H, C><<<grid, kThreads, 0, st>>>(rowptr, col, g, h, s_src, s_dst, m, l, rec, ds_dst,
                                                         partial, P, off_b, M, N, atomic);
  int rc = check_launch("gat_agg_bwd_p1");
  if (rc) return rc;
  gat_agg_bwd_p2_kernel<H, C><<<grid, kThreads, 0, st>>>(rowptr_t, col_t, g, h, s_src, rec, ds_dst, att_src,
                                                         att_dst, dh, partial, P, off_as, off_ad, M, N, atomic);
  return check_launch("gat_agg_bwd_p2");
}

}  // namespace gatres

using namespace gatres;

#define GATRES_DISPATCH_HC(H, C, CALL)                                          \
  do {                                                                          \
    if ((H) == 1 && (C) == 32) return CALL(1, 32);                              \
    if ((H) == 2 && (C) == 32) return CALL(2, 32);                              \
    if ((H) == 1 && (C) == 64) return CALL(1, 64);                              \
    if ((H) == 2 && (C) == 64) return CALL(2, 64);                              \
    if ((H) == 1 && (C) == 128) return CALL(1, 128);                            \
    if ((H) == 2 && (C) == 128) return CALL(2, 128);                            \
    set_error("unsupported (heads=%d, channels=%d): H in {1,2}, C in {32,64,128}", (int)(H), (int)(C)); \
    return GATRES_ERR_ARG;                                                      \
  } while (0)

extern "C" int gatres_gat_agg_fwd(const int32_t* rowptr, const int32_t* col, const float* h, const float* s_src,
                                  const float* s_dst, const float* bias, float* out, float* m, float* l,
                                  int64_t B, int32_t N, int32_t H, int32_t C, int32_t relu, void* stream) {
  GATRES_REQUIRE(B >= 0 && N > 0, "gat_agg_fwd: bad B=%lld N=%d", (long long)B, N);
  GATRES_REQUIRE(B * (int64_t)N < (1ll << 31), "gat_agg_fwd: B*N must be < 2^31 rows");
  GATRES_REQUIRE((m == nullptr) == (l == nullptr), "gat_agg_fwd: m and l must both be given or both NULL");
  if (B == 0) return GATRES_OK;
  const unsigned M = (unsigned)(B * N);
#define CALL(HH, CC) launch_fwd<HH, CC>(rowptr, col, h, s_src, s_dst, bias, out, m, l, M, (unsigned)N, relu, as_stream(stream))
  GATRES_DISPATCH_HC(H, C, CALL);
#undef CALL
}

extern "C" int gatres_gat_agg_bwd(const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t,
                                  const int32_t* col_t, const float* g, const float* h, const float* s_src,
                                  const float* s_dst, const float* m, const float* l, const float* att_src,
                                  const float* att_dst, float* rec, float* ds_dst, float* dh, float* partial,
                                  int64_t P, int32_t slots, int64_t off_att_src, int64_t off_att_dst,
                                  int64_t off_bias, int64_t B, int32_t N, int32_t H, int32_t C, void* stream) {
  GATRES_REQUIRE(B > 0 && N > 0, "gat_agg_bwd: bad B=%lld N=%d", (long long)B, N);
  GATRES_REQUIRE(B * (int64_t)N < (1ll << 31), "gat_agg_bwd: B*N must be < 2^31 rows");
  GATRES_REQUIRE(off_att_src % 4 == 0 && off_att_dst % 4 == 0 && off_bias % 4 == 0 && P % 4 == 0,
                 "gat_agg_bwd: partial row stride and parameter offsets must be multiples of 4 floats");
  const unsigned M = (unsigned)(B * N);
#define CALL(HH, CC)                                                                                              \
  launch_bwd<HH, CC>(rowptr, col, rowptr_t, col_t, g, h, s_src, s_dst, m, l, att_src, att_dst, rec, ds_dst, dh,   \
                     partial, (long long)P, slots, (long long)off_att_src, (long long)off_att_dst,                \
                     (long long)off_bias, M, (unsigned)N, as_stream(stream))
  GATRES_DISPATCH_HC(H, C, CALL);
#undef CALL
}
