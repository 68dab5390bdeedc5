// Fused GAT aggregation, forward and recompute-based backward (sm_100a).
//
// Replaces GATConv.forward steps 2-6 and their autograd (call sites
// /root/reference/gnn_pressure_estimation/GraphModels.py:464-465; semantics in
// SURVEY.md §A.2/§A.4).  One group of LPR lanes owns one (snapshot, row); each
// lane owns one 128-bit chunk of the row, so every neighbour-row gather is a
// coalesced 16 B/lane load and no per-edge tensor is ever materialised.
//
// The per-edge scalar work (logit, LeakyReLU, exp, softmax weight) is done
// COOPERATIVELY: within the LPH lanes that hold one head of a row, lane t
// evaluates edge t of that row, the row max / sum are butterflies over those
// lanes, and the accumulation loop only broadcasts (neighbour id, weight) with
// two shuffles per edge.  (Round-1 profile: the one-lane-does-everything version
// was issue-bound at ~330 instructions per warp pass.)  Rows with more than LPH
// in-edges are handled in LPH-sized chunks with an online-softmax rescale.
#include <math_constants.h>
#include <stdlib.h>
#include "common.cuh"

namespace gatres {

template <int width>
__device__ __forceinline__ float group_max(float v, unsigned mask) {
#pragma unroll
  for (int off = width / 2; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, off));
  return v;
}

// r / n for r < 2^32 with magic = floor(2^32 / n) (n >= 2) or 0xffffffff (n == 1)
__device__ __forceinline__ void divmod(unsigned r, unsigned n, unsigned magic, unsigned& q, unsigned& rem) {
  q = __umulhi(r, magic);
  rem = r - q * n;
  if (rem >= n) { ++q; rem -= n; }
}
static inline unsigned div_magic(unsigned n) { return n <= 1 ? 0xffffffffu : (unsigned)((1ull << 32) / n); }

// ------------------------------------------------------------------ forward
template <int H, int C>
__global__ void __launch_bounds__(kThreads)
gat_agg_fwd_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                   const float* __restrict__ h, const float* __restrict__ s_src,
                   const float* __restrict__ s_dst, const float* __restrict__ bias,
                   float* __restrict__ out, float* __restrict__ m_out, float* __restrict__ l_out,
                   unsigned M, unsigned N, unsigned magic, int relu) {
  using RM = RowMap<H, C>;
  constexpr int F = RM::F, V = RM::V, LPR = RM::LPR, RPW = RM::RPW, LPH = RM::LPH;
  constexpr int PRE = 4;                          // gathers issued before the softmax math (mean WDN in-degree+1 = 3.2)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, lig = lane % LPR, slot = lig % LPH;
  constexpr unsigned gmask = 0xffffffffu;          // control flow below is warp-uniform: full-mask shuffles
  constexpr unsigned rows_per_cta = kWarps * RPW;

  float4 bv[V];
#pragma unroll
  for (int v = 0; v < V; ++v) bv[v] = ldg4(bias + 4 * RM::chunk(lig, v));
  pdl_wait();                                     // activations of the previous kernel become visible here

  if ((unsigned long long)gridDim.x * rows_per_cta >= M) pdl_launch_dependents();   // single pass: see common.cuh
  for (unsigned r0 = blockIdx.x * rows_per_cta; r0 < M; r0 += gridDim.x * rows_per_cta) {
    // rows past the end are clamped (their lanes redo the last row and skip the stores) so that every
    // lane of the warp runs the same instruction stream: plain full-mask SHFLs, no convergence barriers
    const unsigned r_raw = r0 + warp * RPW + sub;
    const bool row_ok = r_raw < M;
    const unsigned r = row_ok ? r_raw : M - 1;
    unsigned b, i;
    divmod(r, N, magic, b, i);
    const float* hb = h + (size_t)b * N * F + 4 * lig;            // this lane's chunk column of snapshot b
    const float* ssb = s_src + (size_t)b * N * H;
    const int beg = __ldg(rowptr + i), deg = __ldg(rowptr + i + 1) - beg;
    const int deg_max = __reduce_max_sync(gmask, deg);

    float sd[V], mrun[V], lrun[V];
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      sd[v] = __ldg(s_dst + (size_t)r * H + RM::head(lig, v));
      mrun[v] = -CUDART_INF_F;
      lrun[v] = 0.f;
      acc[v] = f4zero();
    }
    for (int e0 = 0; e0 < deg_max; e0 += LPH) {
      // lane `slot` of each head group evaluates edge e0 + slot
      const bool valid = e0 + slot < deg;
      const int j = valid ? __ldg(col + beg + e0 + slot) : 0;
      const int cnt = min(LPH, deg - e0);
      const int cnt_max = min(LPH, deg_max - e0);
      // issue the first PRE neighbour-row gathers now: they only depend on `col`, so they fly while the
      // scores are fetched and the softmax is computed (the kernel is latency-bound, not issue-bound)
      float4 x[PRE][V];
#pragma unroll
      for (int u = 0; u < PRE; ++u) {
        const int ju = __shfl_sync(gmask, j, u, LPH);
#pragma unroll
        for (int v = 0; v < V; ++v) x[u][v] = u < cnt ? ldg4(hb + (unsigned)ju * F + 4 * v * LPR) : f4zero();
      }
      float p[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float a = valid ? lrelu(__ldg(ssb + (unsigned)j * H + RM::head(lig, v)) + sd[v]) : -CUDART_INF_F;
        const float nm = fmaxf(mrun[v], group_max<LPH>(a, gmask));
        if (e0 > 0) {                              // online-softmax rescale (never taken for WDN degrees)
          const float sc = __expf(mrun[v] - nm);
          lrun[v] *= sc;
          acc[v].x *= sc; acc[v].y *= sc; acc[v].z *= sc; acc[v].w *= sc;
        }
        p[v] = __expf(a - nm);                     // exp(-inf) = 0 for the idle slots
        lrun[v] += group_sum<LPH>(p[v], gmask);
        mrun[v] = nm;
      }
#pragma unroll
      for (int u = 0; u < PRE; ++u)
#pragma unroll
        for (int v = 0; v < V; ++v) fma4(acc[v], __shfl_sync(gmask, p[v], u, LPH), x[u][v]);   // p = 0 beyond cnt
      for (int t = PRE; t < cnt_max; t += 2) {     // rows with more than PRE in-edges: two gathers in flight
        const int j0 = __shfl_sync(gmask, j, t, LPH), j1 = __shfl_sync(gmask, j, t + 1, LPH);
        float4 x0[V], x1[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          x0[v] = t < cnt ? ldg4(hb + (unsigned)j0 * F + 4 * v * LPR) : f4zero();
          x1[v] = t + 1 < cnt ? ldg4(hb + (unsigned)j1 * F + 4 * v * LPR) : f4zero();
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {               // p is 0 in the idle slots (t >= cnt)
          const float p0 = __shfl_sync(gmask, p[v], t, LPH), p1 = __shfl_sync(gmask, p[v], t + 1, LPH);
          fma4(acc[v], p0, x0[v]);
          fma4(acc[v], t + 1 < LPH ? p1 : 0.f, x1[v]);
        }
      }
    }
    if (!row_ok) continue;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float inv = 1.f / (lrun[v] + kSoftmaxEps);
      float4 o;
      o.x = fmaf(acc[v].x, inv, bv[v].x);
      o.y = fmaf(acc[v].y, inv, bv[v].y);
      o.z = fmaf(acc[v].z, inv, bv[v].z);
      o.w = fmaf(acc[v].w, inv, bv[v].w);
      if (relu) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
      }
      st4(out + (size_t)r * F + 4 * RM::chunk(lig, v), o);
      if (m_out != nullptr && slot == 0) {
        m_out[(size_t)r * H + RM::head(lig, v)] = mrun[v];
        l_out[(size_t)r * H + RM::head(lig, v)] = lrun[v];
      }
    }
  }
}

// --------------------------------------------------------- backward, pass 1
// Per target row i: D_i = sum_e alpha_e dalpha_e and ds_dst[i] = sum_e dz_e, with
// dalpha_e = <g[i], h[j]>.  Emits rec[i,h] = {s_dst, m, 1/(l+eps), D} so pass 2
// needs one 16 B load per (edge, head) for all target-side scalars.
// Also accumulates the bias gradient (column sums of g).
// Lane `slot` keeps (alpha, slope, dalpha) of edge `slot`; the three row sums are
// butterflies at the end.
template <int H, int C>
__global__ void __launch_bounds__(kThreads)
gat_agg_bwd_p1_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                      const float* __restrict__ g, const float* __restrict__ h,
                      const float* __restrict__ s_src, const float* __restrict__ s_dst,
                      const float* __restrict__ m, const float* __restrict__ l,
                      float* __restrict__ rec, float* __restrict__ ds_dst,
                      float* __restrict__ partial, long long P, long long off_bias,
                      unsigned M, unsigned N, unsigned magic, int atomic) {
  using RM = RowMap<H, C>;
  constexpr int F = RM::F, V = RM::V, LPR = RM::LPR, RPW = RM::RPW, LPH = RM::LPH;
  constexpr int PRE = 4;                          // gathers issued ahead of the per-edge scalar math
  __shared__ float red[kWarps * 32 * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, lig = lane % LPR, slot = lig % LPH;
  constexpr unsigned gmask = 0xffffffffu;          // warp-uniform control flow, full-mask shuffles
  constexpr unsigned rows_per_cta = kWarps * RPW;

  float4 bacc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) bacc[v] = f4zero();
  pdl_wait();

  if ((unsigned long long)gridDim.x * rows_per_cta >= M) pdl_launch_dependents();   // single pass: see common.cuh
  for (unsigned r0 = blockIdx.x * rows_per_cta; r0 < M; r0 += gridDim.x * rows_per_cta) {
    const unsigned r_raw = r0 + warp * RPW + sub;
    const bool row_ok = r_raw < M;                // clamped rows recompute the last row and emit nothing
    const unsigned r = row_ok ? r_raw : M - 1;
    {
      unsigned b, i;
      divmod(r, N, magic, b, i);
      const float* hb = h + (size_t)b * N * F + 4 * lig;
      const float* ssb = s_src + (size_t)b * N * H;
      const int beg = __ldg(rowptr + i), deg = __ldg(rowptr + i + 1) - beg;
      const int deg_max = __reduce_max_sync(gmask, deg);
      float4 gv[V];
      float sd[V], mi[V], il[V], S1[V], S2[V], S3[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int hd = RM::head(lig, v);
        gv[v] = ldg4_stream(g + (size_t)r * F + 4 * RM::chunk(lig, v));
        if (row_ok) add4(bacc[v], gv[v]);
        sd[v] = __ldg(s_dst + (size_t)r * H + hd);
        mi[v] = __ldg(m + (size_t)r * H + hd);
        il[v] = 1.f / (__ldg(l + (size_t)r * H + hd) + kSoftmaxEps);
        S1[v] = S2[v] = S3[v] = 0.f;
      }
      for (int e0 = 0; e0 < deg_max; e0 += LPH) {
        const bool valid = e0 + slot < deg;
        const int j = valid ? __ldg(col + beg + e0 + slot) : 0;
        const int cnt = min(LPH, deg - e0);
        const int cnt_max = min(LPH, deg_max - e0);
        float4 x[PRE][V];                          // first PRE neighbour rows, in flight during the alpha math
#pragma unroll
        for (int u = 0; u < PRE; ++u) {
          const int ju = __shfl_sync(gmask, j, u, LPH);
#pragma unroll
          for (int v = 0; v < V; ++v) x[u][v] = u < cnt ? ldg4(hb + (unsigned)ju * F + 4 * v * LPR) : f4zero();
        }
        float alpha[V], sl[V], da[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float z = __ldg(ssb + (unsigned)j * H + RM::head(lig, v)) + sd[v];
          alpha[v] = valid ? __expf(lrelu(z) - mi[v]) * il[v] : 0.f;
          sl[v] = lrelu_slope(z);
          da[v] = 0.f;
        }
#pragma unroll
        for (int u = 0; u < PRE; ++u)
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float d = group_sum<LPH>(dot4(gv[v], x[u][v]), gmask);
            da[v] = slot == u ? d : da[v];
          }
        for (int t = PRE; t < cnt_max; t += 2) {
          const int j0 = __shfl_sync(gmask, j, t, LPH), j1 = __shfl_sync(gmask, j, t + 1, LPH);
          float4 x0[V], x1[V];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            x0[v] = t < cnt ? ldg4(hb + (unsigned)j0 * F + 4 * v * LPR) : f4zero();
            x1[v] = t + 1 < cnt ? ldg4(hb + (unsigned)j1 * F + 4 * v * LPR) : f4zero();
          }
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float d0 = group_sum<LPH>(dot4(gv[v], x0[v]), gmask);
            const float d1 = group_sum<LPH>(dot4(gv[v], x1[v]), gmask);
            da[v] = slot == t ? d0 : (slot == t + 1 ? d1 : da[v]);
          }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {             // idle slots have alpha = 0
          S1[v] = fmaf(alpha[v], da[v], S1[v]);
          S2[v] = fmaf(alpha[v] * sl[v], da[v], S2[v]);
          S3[v] = fmaf(alpha[v], sl[v], S3[v]);
        }
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float D = group_sum<LPH>(S1[v], gmask);
        const float T2 = group_sum<LPH>(S2[v], gmask), T3 = group_sum<LPH>(S3[v], gmask);
        if (slot == 0 && row_ok) {
          const size_t o = (size_t)r * H + RM::head(lig, v);
          st4(rec + o * 4, make_float4(sd[v], mi[v], il[v], D));
          ds_dst[o] = T2 - D * T3;
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
                      unsigned M, unsigned N, unsigned magic, int atomic) {
  using RM = RowMap<H, C>;
  constexpr int F = RM::F, V = RM::V, LPR = RM::LPR, RPW = RM::RPW, LPH = RM::LPH;
  constexpr int PRE = 4;                          // gathers issued ahead of the per-edge scalar math
  __shared__ float red[kWarps * 32 * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPR, lig = lane % LPR, slot = lig % LPH;
  constexpr unsigned gmask = 0xffffffffu;          // warp-uniform control flow, full-mask shuffles
  constexpr unsigned rows_per_cta = kWarps * RPW;

  float4 as[V], ad[V], accs[V], accd[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    as[v] = ldg4(att_src + 4 * RM::chunk(lig, v));
    ad[v] = ldg4(att_dst + 4 * RM::chunk(lig, v));
    accs[v] = f4zero();
    accd[v] = f4zero();
  }
  pdl_wait();

  if ((unsigned long long)gridDim.x * rows_per_cta >= M) pdl_launch_dependents();   // single pass: see common.cuh
  for (unsigned r0 = blockIdx.x * rows_per_cta; r0 < M; r0 += gridDim.x * rows_per_cta) {
    const unsigned r_raw = r0 + warp * RPW + sub;
    const bool row_ok = r_raw < M;
    const unsigned r = row_ok ? r_raw : M - 1;
    {
      unsigned b, jn;
      divmod(r, N, magic, b, jn);
      const float* gb = g + (size_t)b * N * F + 4 * lig;
      const float* recb = rec + (size_t)b * N * H * 4;
      const int beg = __ldg(rowptr_t + jn), deg = __ldg(rowptr_t + jn + 1) - beg;
      const int deg_max = __reduce_max_sync(gmask, deg);
      float4 hv[V], dacc[V];
      float ss[V], dsrc[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        hv[v] = ldg4_stream(h + (size_t)r * F + 4 * RM::chunk(lig, v));
        ss[v] = __ldg(s_src + (size_t)r * H + RM::head(lig, v));
        dacc[v] = f4zero();
        dsrc[v] = 0.f;
      }
      for (int e0 = 0; e0 < deg_max; e0 += LPH) {
        const bool valid = e0 + slot < deg;
        const int i = valid ? __ldg(col_t + beg + e0 + slot) : 0;
        const int cnt = min(LPH, deg - e0);
        const int cnt_max = min(LPH, deg_max - e0);
        float4 gx[PRE][V];                         // first PRE target-gradient rows, in flight during the alpha math
#pragma unroll
        for (int u = 0; u < PRE; ++u) {
          const int iu = __shfl_sync(gmask, i, u, LPH);
#pragma unroll
          for (int v = 0; v < V; ++v) gx[u][v] = u < cnt ? ldg4(gb + (unsigned)iu * F + 4 * v * LPR) : f4zero();
        }
        float alpha[V], k2[V], Dt[V], da[V];      // k2 = alpha * lrelu'(z); Dt = D of the edge's target
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float4 t4 = ldg4(recb + ((unsigned)i * H + RM::head(lig, v)) * 4);   // {s_dst, m, 1/l, D}
          const float z = ss[v] + t4.x;
          alpha[v] = valid ? __expf(lrelu(z) - t4.y) * t4.z : 0.f;
          k2[v] = alpha[v] * lrelu_slope(z);
          Dt[v] = t4.w;
          da[v] = 0.f;
        }
#pragma unroll
        for (int u = 0; u < PRE; ++u)
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float d = group_sum<LPH>(dot4(gx[u][v], hv[v]), gmask);
            da[v] = slot == u ? d : da[v];
            fma4(dacc[v], __shfl_sync(gmask, alpha[v], u, LPH), gx[u][v]);      // alpha = 0 beyond cnt
          }
        for (int t = PRE; t < cnt_max; t += 2) {
          const int i0 = __shfl_sync(gmask, i, t, LPH), i1 = __shfl_sync(gmask, i, t + 1, LPH);
          float4 g0[V], g1[V];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            g0[v] = t < cnt ? ldg4(gb + (unsigned)i0 * F + 4 * v * LPR) : f4zero();
            g1[v] = t + 1 < cnt ? ldg4(gb + (unsigned)i1 * F + 4 * v * LPR) : f4zero();
          }
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float d0 = group_sum<LPH>(dot4(g0[v], hv[v]), gmask);
            const float d1 = group_sum<LPH>(dot4(g1[v], hv[v]), gmask);
            da[v] = slot == t ? d0 : (slot == t + 1 ? d1 : da[v]);
            const float a0 = __shfl_sync(gmask, alpha[v], t, LPH), a1 = __shfl_sync(gmask, alpha[v], t + 1, LPH);
            fma4(dacc[v], a0, g0[v]);               // alpha is 0 in the idle slots
            fma4(dacc[v], t + 1 < LPH ? a1 : 0.f, g1[v]);
          }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) dsrc[v] = fmaf(k2[v], da[v] - Dt[v], dsrc[v]);
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float ds = group_sum<LPH>(dsrc[v], gmask);
        const float dd = __ldg(ds_dst + (size_t)r * H + RM::head(lig, v));
        fma4(dacc[v], ds, as[v]);
        fma4(dacc[v], dd, ad[v]);
        if (row_ok) {
          st4(dh + (size_t)r * F + 4 * RM::chunk(lig, v), dacc[v]);
          fma4(accs[v], ds, hv[v]);
          fma4(accd[v], dd, hv[v]);
        }
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
  const unsigned grid = row_kernel_grid(M, kWarps * RowMap<H, C>::RPW, 32);
  launch_kernel(gat_agg_fwd_kernel<H, C>, dim3(grid), dim3(kThreads), 0, st, rowptr, col, h, s_src, s_dst, bias, out, m, l, M, N,
                                                      div_magic(N), relu);
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
  launch_kernel(gat_agg_bwd_p1_kernel<H, C>, dim3(grid), dim3(kThreads), 0, st, rowptr, col, g, h, s_src, s_dst, m, l, rec, ds_dst,
                                                         partial, P, off_b, M, N, div_magic(N), atomic);
  int rc = check_launch("gat_agg_bwd_p1");
  if (rc) return rc;
  launch_kernel(gat_agg_bwd_p2_kernel<H, C>, dim3(grid), dim3(kThreads), 0, st, rowptr_t, col_t, g, h, s_src, rec, ds_dst, att_src,
                                                         att_dst, dh, partial, P, off_as, off_ad, M, N,
                                                         div_magic(N), atomic);
  return check_launch("gat_agg_bwd_p2");
}

// snapshot-tile (TMA-staged) variants, gat_agg_tile.cu
bool bwd_tile_eligible(unsigned N, unsigned H, unsigned C, unsigned E1);
int gat_agg_bwd_tile(const int* rowptr, const int* col, const int* rowptr_t, const int* col_t, unsigned E1,
                     const float* g, const float* h, const float* s_src, const float* s_dst, const float* m,
                     const float* l, const float* att_src, const float* att_dst, float* dh, float* grads,
                     long long off_as, long long off_ad, long long off_b, unsigned B, unsigned N, int H, int C,
                     cudaStream_t st);
bool fwd_tile_eligible(unsigned N, unsigned H, unsigned C, unsigned E1);
int gat_agg_fwd_tile(const int* rowptr, const int* col, unsigned E1, const float* h, const float* s_src,
                     const float* s_dst, const float* bias, float* out, float* m, float* l, unsigned B, unsigned N,
                     int H, int C, int relu, cudaStream_t st);

// Batches below this use the gather kernels (a snapshot-per-CTA grid would leave SMs idle).
// GATRES_TILE_MIN_B overrides (set it huge to disable the tile kernels).
static long long g_tile_min_batch = -1;
static long long tile_min_batch() {
  if (g_tile_min_batch < 0) {
    const char* e = getenv("GATRES_TILE_MIN_B");
    g_tile_min_batch = e ? atoll(e) : 64;
  }
  return g_tile_min_batch;
}

// conv2 aggregation fused with SimpleConv(mean) + residual + ReLU (gat_agg_tile.cu); used by gatres_forward
bool fwd_tile_mean_eligible(unsigned N, unsigned C, unsigned E1);
int gat_agg_mean_res_fwd_tile(const int* rowptr, const int* col, unsigned E1, const float* h, const float* s_src,
                              const float* s_dst, const float* bias, float* m, float* l, const float* x0, float* xout,
                              unsigned B, unsigned N, int C, cudaStream_t st);

static int fused_mean_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GATRES_FUSE_MEAN");
    v = e ? atoi(e) : 1;
  }
  return v;
}

// -> 1 if the fused kernel ran, 0 if the shapes / batch do not take the tile path (caller runs the two kernels), < 0 on error
int gat_agg_mean_res_fwd(const int* rowptr, const int* col, unsigned E1, const float* h, const float* s_src,
                         const float* s_dst, const float* bias, float* m, float* l, const float* x0, float* xout,
                         long long B, int N, int C, cudaStream_t st) {
  if (!fused_mean_enabled() || E1 == 0 || B < tile_min_batch() || (C != 32 && C != 64 && C != 128) ||
      !fwd_tile_mean_eligible((unsigned)N, (unsigned)C, E1))
    return 0;
  const int rc = gat_agg_mean_res_fwd_tile(rowptr, col, E1, h, s_src, s_dst, bias, m, l, x0, xout, (unsigned)B, (unsigned)N, C, st);
  return rc ? rc : 1;
}

}  // namespace gatres

using namespace gatres;

namespace gatres {
// SimpleConv(mean) backward of the model's layer path (ReLU mask already applied): snapshot-tile kernel for large
// batches of graphs whose slab fits shared memory, else the gather kernel.  1 = done, 0 = not applicable, < 0 = error.
bool mean_bwd_tile_eligible(unsigned N, unsigned C, unsigned E1);
int mean_res_bwd_tile(const int* rowptr, const int* rowptr_t, const int* col_t, unsigned E1, const float* g, float* dz,
                      unsigned B, unsigned N, int C, cudaStream_t st);
int mean_res_bwd_model(const int* rowptr, const int* rowptr_t, const int* col_t, unsigned E1, const float* g, float* dz,
                       long long B, int N, int C, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("GATRES_MEAN_BWD_TILE");
    enabled = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  if (!enabled || E1 == 0 || B < tile_min_batch() || B * (long long)N >= (1ll << 31) ||
      !mean_bwd_tile_eligible((unsigned)N, (unsigned)C, E1))
    return 0;
  const int rc = mean_res_bwd_tile(rowptr, rowptr_t, col_t, E1, g, dz, (unsigned)B, (unsigned)N, C, st);
  return rc == GATRES_OK ? 1 : rc;
}
}  // namespace gatres

extern "C" int gatres_mean_res_bwd(const int32_t* rowptr, const int32_t* rowptr_t, const int32_t* col_t,
                                   const float* g_out, const float* out, float* dz, float* dres,
                                   int64_t B, int32_t N, int32_t C, void* stream);
extern "C" int gatres_mean_res_bwd_e1(const int32_t* rowptr, const int32_t* rowptr_t, const int32_t* col_t, int32_t E1,
                                      const float* g_masked, float* dz, int64_t B, int32_t N, int32_t C, void* stream) {
  GATRES_REQUIRE(B >= 0 && N > 0 && E1 >= 0, "mean_res_bwd_e1: bad B=%lld N=%d E1=%d", (long long)B, N, E1);
  if (B == 0) return GATRES_OK;
  const int tiled = mean_res_bwd_model(rowptr, rowptr_t, col_t, (unsigned)E1, g_masked, dz, B, N, C, as_stream(stream));
  if (tiled != 0) return tiled < 0 ? tiled : GATRES_OK;
  return gatres_mean_res_bwd(rowptr, rowptr_t, col_t, g_masked, nullptr, dz, nullptr, B, N, C, stream);
}

extern "C" int64_t gatres_set_tile_min_batch(int64_t min_batch) {
  const long long prev = tile_min_batch();
  if (min_batch >= 0) g_tile_min_batch = min_batch;
  return prev;
}

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
                                  int64_t B, int32_t N, int32_t E1, int32_t H, int32_t C, int32_t relu,
                                  void* stream) {
  GATRES_REQUIRE(B >= 0 && N > 0, "gat_agg_fwd: bad B=%lld N=%d", (long long)B, N);
  GATRES_REQUIRE(B * (int64_t)N < (1ll << 31), "gat_agg_fwd: B*N must be < 2^31 rows");
  GATRES_REQUIRE((int64_t)N * H * C < (1ll << 31), "gat_agg_fwd: one snapshot must be < 2^31 floats");
  GATRES_REQUIRE((m == nullptr) == (l == nullptr), "gat_agg_fwd: m and l must both be given or both NULL");
  if (B == 0) return GATRES_OK;
  const unsigned M = (unsigned)(B * N);
  if (E1 > 0 && (H == 1 || H == 2) && (C == 32 || C == 64 || C == 128) && B >= tile_min_batch() &&
      fwd_tile_eligible((unsigned)N, (unsigned)H, (unsigned)C, (unsigned)E1))
    return gat_agg_fwd_tile(rowptr, col, (unsigned)E1, h, s_src, s_dst, bias, out, m, l, (unsigned)B, (unsigned)N, H, C,
                            relu, as_stream(stream));
#define CALL(HH, CC) launch_fwd<HH, CC>(rowptr, col, h, s_src, s_dst, bias, out, m, l, M, (unsigned)N, relu, as_stream(stream))
  GATRES_DISPATCH_HC(H, C, CALL);
#undef CALL
}

extern "C" int gatres_gat_agg_bwd(const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t,
                                  const int32_t* col_t, const float* g, const float* h, const float* s_src,
                                  const float* s_dst, const float* m, const float* l, const float* att_src,
                                  const float* att_dst, float* rec, float* ds_dst, float* dh, float* partial,
                                  int64_t P, int32_t slots, int64_t off_att_src, int64_t off_att_dst,
                                  int64_t off_bias, int64_t B, int32_t N, int32_t E1, int32_t H, int32_t C,
                                  void* stream) {
  GATRES_REQUIRE(B > 0 && N > 0, "gat_agg_bwd: bad B=%lld N=%d", (long long)B, N);
  GATRES_REQUIRE(B * (int64_t)N < (1ll << 31), "gat_agg_bwd: B*N must be < 2^31 rows");
  GATRES_REQUIRE((int64_t)N * H * C < (1ll << 31), "gat_agg_bwd: one snapshot must be < 2^31 floats");
  GATRES_REQUIRE(off_att_src % 4 == 0 && off_att_dst % 4 == 0 && off_bias % 4 == 0 && P % 4 == 0,
                 "gat_agg_bwd: partial row stride and parameter offsets must be multiples of 4 floats");
  const unsigned M = (unsigned)(B * N);
  // fused snapshot-tile backward: atomic-accumulation mode only (it adds straight into the gradient buffer)
  if (slots <= 0 && E1 > 0 && (H == 1 || H == 2) && (C == 32 || C == 64 || C == 128) && B >= tile_min_batch() &&
      bwd_tile_eligible((unsigned)N, (unsigned)H, (unsigned)C, (unsigned)E1))
    return gat_agg_bwd_tile(rowptr, col, rowptr_t, col_t, (unsigned)E1, g, h, s_src, s_dst, m, l, att_src, att_dst, dh,
                            partial, (long long)off_att_src, (long long)off_att_dst, (long long)off_bias, (unsigned)B,
                            (unsigned)N, H, C, as_stream(stream));
#define CALL(HH, CC)                                                                                              \
  launch_bwd<HH, CC>(rowptr, col, rowptr_t, col_t, g, h, s_src, s_dst, m, l, att_src, att_dst, rec, ds_dst, dh,   \
                     partial, (long long)P, slots, (long long)off_att_src, (long long)off_att_dst,                \
                     (long long)off_bias, M, (unsigned)N, as_stream(stream))
  GATRES_DISPATCH_HC(H, C, CALL);
#undef CALL
}
