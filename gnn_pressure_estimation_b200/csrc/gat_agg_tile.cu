// Snapshot-tile variants of the GAT aggregation kernels (sm_100a): for graphs whose
// per-snapshot feature slab fits in shared memory (C-Town: 388 x 64 fp32 = 97 KB) a
// persistent CTA owns one snapshot at a time.  The slab h[b] ([N,F], contiguous) and
// the source scores are staged by the TMA engine (1-D `cp.async.bulk` + mbarrier,
// double-buffered, so the copy of snapshot k+1 flies under the math of snapshot k);
// the shared CSR is staged once per CTA.  Every neighbour gather then hits shared
// memory, DRAM traffic is exactly the algorithmic bytes, and DRAM latency is off the
// warps' critical path (round-1 ncu: the gather version is long-scoreboard bound).
//
// Same arithmetic and lane mapping as gat_agg.cu (cooperative per-row softmax).
#include <math_constants.h>
#include "common.cuh"
#include "tma.cuh"

namespace gatres {

// CTA size: 1024 threads when only one CTA fits per SM (slab > ~110 KB double-buffered), 512 when two fit.

template <int width>
__device__ __forceinline__ float tile_group_max(float v, unsigned mask) {
#pragma unroll
  for (int off = width / 2; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, off));
  return v;
}

// shared-memory plan of the forward tile kernel (host and device agree through these)
struct FwdTilePlan {
  uint32_t h_bytes, ss_bytes, stage_bytes, rp_off, col_off, bar_off, total;
  __host__ __device__ FwdTilePlan(unsigned N, unsigned F, unsigned H, unsigned E1) {
    h_bytes = N * F * 4u;
    ss_bytes = N * H * 4u;
    stage_bytes = h_bytes + 2u * ss_bytes;          // slab + source scores + target scores
    rp_off = 2u * stage_bytes;
    col_off = rp_off + (((N + 1u) * 4u + 15u) & ~15u);
    bar_off = col_off + ((E1 * 4u + 15u) & ~15u);
    total = bar_off + 32u;
  }
};

template <int H, int C, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
gat_agg_fwd_tile_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, unsigned E1,
                        const float* __restrict__ h, const float* __restrict__ s_src,
                        const float* __restrict__ s_dst, const float* __restrict__ bias,
                        float* __restrict__ out, float* __restrict__ m_out, float* __restrict__ l_out,
                        unsigned B, unsigned N, int relu) {
  using RM = RowMap<H, C>;
  constexpr int F = RM::F, V = RM::V, LPR = RM::LPR, RPW = RM::RPW, LPH = RM::LPH;
  constexpr int kTileThreads = THREADS, kTileWarps = THREADS / 32;
  extern __shared__ __align__(128) unsigned char smem[];
  const FwdTilePlan plan(N, F, H, E1);
  int* rp_s = reinterpret_cast<int*>(smem + plan.rp_off);
  int* col_s = reinterpret_cast<int*>(smem + plan.col_off);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + plan.bar_off);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane / LPR, lig = lane % LPR, slot = lig % LPH;
  constexpr unsigned gmask = 0xffffffffu;          // control flow below is warp-uniform: full-mask shuffles

  auto issue = [&](int stage, unsigned bb) {       // one elected thread
    unsigned char* dst = smem + (size_t)stage * plan.stage_bytes;
    mbar_arrive_expect_tx(&full[stage], plan.stage_bytes);
    bulk_g2s(dst, h + (size_t)bb * N * F, plan.h_bytes, &full[stage]);
    bulk_g2s(dst + plan.h_bytes, s_src + (size_t)bb * N * H, plan.ss_bytes, &full[stage]);
    bulk_g2s(dst + plan.h_bytes + plan.ss_bytes, s_dst + (size_t)bb * N * H, plan.ss_bytes, &full[stage]);
  };

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
  }
  for (unsigned k = tid; k <= N; k += kTileThreads) rp_s[k] = __ldg(rowptr + k);
  for (unsigned k = tid; k < E1; k += kTileThreads) col_s[k] = __ldg(col + k);
  __syncthreads();
  if (tid == 0) {
    if (blockIdx.x < B) issue(0, blockIdx.x);
    if (blockIdx.x + gridDim.x < B) issue(1, blockIdx.x + gridDim.x);
  }

  float4 bv[V];
#pragma unroll
  for (int v = 0; v < V; ++v) bv[v] = ldg4(bias + 4 * RM::chunk(lig, v));

  unsigned k = 0;
  for (unsigned b = blockIdx.x; b < B; b += gridDim.x, ++k) {
    const int stage = k & 1;
    const float* hs = reinterpret_cast<const float*>(smem + (size_t)stage * plan.stage_bytes) + 4 * lig;
    const float* sss = reinterpret_cast<const float*>(smem + (size_t)stage * plan.stage_bytes + plan.h_bytes);
    const float* sds = sss + N * H;
    mbar_wait(&full[stage], (k >> 1) & 1);

    for (unsigned i0 = warp * RPW; i0 < N; i0 += kTileWarps * RPW) {
      // rows past the end are clamped (recompute the last row, store nothing): every lane of the warp
      // runs the same instruction stream, so shuffles are plain full-mask SHFLs
      const bool row_ok = i0 + sub < N;
      const unsigned i = row_ok ? i0 + sub : N - 1;
      const size_t r = (size_t)b * N + i;
      const int beg = rp_s[i], deg = rp_s[i + 1] - beg;
      const int deg_max = __reduce_max_sync(gmask, deg);
      float sd[V], mrun[V], lrun[V];
      float4 acc[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        sd[v] = sds[i * H + RM::head(lig, v)];
        mrun[v] = -CUDART_INF_F;
        lrun[v] = 0.f;
        acc[v] = f4zero();
      }
      for (int e0 = 0; e0 < deg_max; e0 += LPH) {
        const bool valid = e0 + slot < deg;
        const int j = valid ? col_s[beg + e0 + slot] : 0;
        float p[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float a = valid ? lrelu(sss[j * H + RM::head(lig, v)] + sd[v]) : -CUDART_INF_F;
          const float nm = fmaxf(mrun[v], tile_group_max<LPH>(a, gmask));
          if (e0 > 0) {
            const float sc = __expf(mrun[v] - nm);
            lrun[v] *= sc;
            acc[v].x *= sc; acc[v].y *= sc; acc[v].z *= sc; acc[v].w *= sc;
          }
          p[v] = __expf(a - nm);
          lrun[v] += group_sum<LPH>(p[v], gmask);
          mrun[v] = nm;
        }
        const int cnt_max = min(LPH, deg_max - e0);
        for (int t = 0; t < cnt_max; ++t) {          // idle slots carry j = 0, p = 0
          const int jt = __shfl_sync(gmask, j, t, LPH);
#pragma unroll
          for (int v = 0; v < V; ++v)
            fma4(acc[v], __shfl_sync(gmask, p[v], t, LPH),
                 *reinterpret_cast<const float4*>(hs + jt * F + 4 * v * LPR));
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
        st4(out + r * F + 4 * RM::chunk(lig, v), o);
        if (m_out != nullptr && slot == 0) {
          m_out[r * H + RM::head(lig, v)] = mrun[v];
          l_out[r * H + RM::head(lig, v)] = lrun[v];
        }
      }
    }
    __syncthreads();                               // every warp is done with this stage
    if (tid == 0) {
      const unsigned nb = b + 2u * gridDim.x;
      if (nb < B) {
        fence_proxy_async();
        issue(stage, nb);
      }
    }
  }
}

template <int H, int C>
static int launch_fwd_tile(const int* rowptr, const int* col, unsigned E1, const float* h, const float* s_src,
                           const float* s_dst, const float* bias, float* out, float* m, float* l, unsigned B,
                           unsigned N, int relu, cudaStream_t st) {
  const FwdTilePlan plan(N, H * C, H, E1);
  unsigned per_sm = (unsigned)((227u * 1024u) / (plan.total + 1024u));
  per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
  unsigned grid = (unsigned)sm_count() * per_sm;
  if (grid > B) grid = B;
  if (per_sm >= 2) {
    auto kern = gat_agg_fwd_tile_kernel<H, C, 512>;
    static uint32_t configured = 0;
    if (configured < plan.total) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total) != cudaSuccess)
        return check_launch("gat_agg_fwd_tile: smem attribute");
      configured = plan.total;
    }
    kern<<<grid, 512, plan.total, st>>>(rowptr, col, E1, h, s_src, s_dst, bias, out, m, l, B, N, relu);
  } else {
    auto kern = gat_agg_fwd_tile_kernel<H, C, 1024>;
    static uint32_t configured = 0;
    if (configured < plan.total) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total) != cudaSuccess)
        return check_launch("gat_agg_fwd_tile: smem attribute");
      configured = plan.total;
    }
    kern<<<grid, 1024, plan.total, st>>>(rowptr, col, E1, h, s_src, s_dst, bias, out, m, l, B, N, relu);
  }
  return check_launch("gat_agg_fwd_tile");
}

// Eligibility: slab + scores double-buffered + CSR must fit, and the per-snapshot byte
// counts must be 16 B multiples (bulk-copy granularity).
bool fwd_tile_eligible(unsigned N, unsigned H, unsigned C, unsigned E1) {
  const FwdTilePlan plan(N, H * C, H, E1);
  return (N * H) % 4u == 0 && plan.total <= 227u * 1024u && (H * C) <= 128;
}

int gat_agg_fwd_tile(const int* rowptr, const int* col, unsigned E1, const float* h, const float* s_src,
                     const float* s_dst, const float* bias, float* out, float* m, float* l, unsigned B, unsigned N,
                     int H, int C, int relu, cudaStream_t st) {
#define T(HH, CC) \
  if (H == HH && C == CC) return launch_fwd_tile<HH, CC>(rowptr, col, E1, h, s_src, s_dst, bias, out, m, l, B, N, relu, st)
  T(1, 32);
  T(2, 32);
  T(1, 64);
  T(2, 64);
  T(1, 128);
#undef T
  set_error("gat_agg_fwd_tile: unsupported (H=%d, C=%d)", H, C);
  return GATRES_ERR_ARG;
}

}  // namespace gatres
