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
#include <cuda.h>
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tensormap.cuh"
#include "tma.cuh"

namespace gatres {

// global -> shared 2-D tiled bulk load (TMA): the box of `map` at element coordinates (x = column, y = row) lands densely
// ([box rows][box columns]) at smem_dst; completion is signalled on `bar` as the box's bytes
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}

// CTA size: 1024 threads when only one CTA fits per SM (slab > ~110 KB double-buffered), 512 when two fit.

template <int width>
__device__ __forceinline__ float tile_group_max(float v, unsigned mask) {
#pragma unroll
  for (int off = width / 2; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, off));
  return v;
}

// Rows in ascending degree (counting sort in shared memory, degrees capped at 31).  A warp runs the edge loops of its 4-8
// rows to the LARGEST degree among them (warp-uniform control flow); handing it rows of equal degree removes the padding
// iterations (C-Town-shaped network: 1.23x the edge visits with 4 rows per warp in node order, 1.03x sorted).  The order
// inside a degree class is whatever the atomics give: rows are independent, only the row -> warp assignment changes.
// rp: staged row pointers (visible: call after a CTA barrier).  Ends with a CTA barrier.
template <int THREADS>
__device__ __forceinline__ void degree_order(const int* rp, unsigned N, unsigned short* ord, int* hist) {
  const int tid = threadIdx.x;
  if (tid < 32) hist[tid] = 0;
  __syncthreads();
  for (unsigned k = tid; k < N; k += THREADS) atomicAdd(&hist[min(rp[k + 1] - rp[k], 31)], 1);
  __syncthreads();
  if (tid < 32) {                                   // exclusive scan: first slot of every degree class
    const int c = hist[tid];
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += up;
    }
    hist[tid] = incl - c;
  }
  __syncthreads();
  for (unsigned k = tid; k < N; k += THREADS) ord[atomicAdd(&hist[min(rp[k + 1] - rp[k], 31)], 1)] = (unsigned short)k;
  __syncthreads();
}

// shared-memory plan of the forward tile kernel (host and device agree through these)
struct FwdTilePlan {
  uint32_t h_bytes, ss_bytes, tx_bytes, stage_bytes, zs_off, xs_off, rp_off, col_off, ord_off, hist_off, bar_off, total;
  // fuse_mean: two more slabs, the layer output z of the snapshot (never written to HBM) and the block input x0
  // lean (fused mean in 64-channel slices): ONE stage and no x0 slab — the next unit's slab is requested when the aggregation
  // phase is done with the stage and lands under the mean phase, x0 rows are read from global memory
  __host__ __device__ FwdTilePlan(unsigned N, unsigned F, unsigned H, unsigned E1, bool fuse_mean = false, bool lean = false) {
    h_bytes = N * F * 4u;
    ss_bytes = N * H * 4u;
    tx_bytes = h_bytes + 2u * ss_bytes;                       // slab + source scores + target scores (what the TMA delivers)
    stage_bytes = (tx_bytes + 127u) & ~127u;                  // stride between the two stages
    zs_off = (lean ? 1u : 2u) * stage_bytes;
    xs_off = zs_off + (fuse_mean ? h_bytes : 0u);
    rp_off = xs_off + ((fuse_mean && !lean) ? h_bytes : 0u);
    col_off = rp_off + (((N + 1u) * 4u + 15u) & ~15u);
    ord_off = col_off + ((E1 * 4u + 15u) & ~15u);                  // degree order of the rows (uint16) + its histogram
    hist_off = ord_off + ((N * 2u + 15u) & ~15u);
    bar_off = hist_off + 128u;
    total = bar_off + 32u;
  }
};

// FUSE_MEAN (one head, concat = False: conv2 of a GATRes block): the layer output z stays in shared memory and the
// SimpleConv(mean) + residual + ReLU that always follows (GraphModels.py:466-467) runs in the same kernel:
// xout[i] = relu(mean_{j in N(i)} z[j] + x0[i]).  z is neither written to nor re-read from HBM and one launch
// disappears: 272 + 384 algorithmic B/node become 400.
//
// SLICED (rows wider than a slab that fits: nc = 128, two heads of nc = 64): the aggregation is independent per channel
// once the attention coefficients are known, so a work unit is (snapshot, slice of C channels of ONE head) instead of a
// snapshot: the slab is the [N x C] column block of h[b] (row stride ld floats), fetched by 2-D TMA tensor loads
// (hmap / xmap: boxes of C columns x box_rows rows, dense in shared memory, i.e. the layout of the contiguous case);
// every slice of a head recomputes that head's softmax from the (tiny) score vectors, the first slice of a head
// writes (m, l).  H = 1 in this mode; Hs = heads of the score tensors, sph = slices per head.
template <int H, int C, int THREADS, bool FUSE_MEAN, bool SLICED>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : 1)
gat_agg_fwd_tile_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, unsigned E1,
                        const float* __restrict__ h, const float* __restrict__ s_src,
                        const float* __restrict__ s_dst, const float* __restrict__ bias,
                        float* __restrict__ out, float* __restrict__ m_out, float* __restrict__ l_out,
                        const float* __restrict__ x0, float* __restrict__ xout,
                        unsigned B, unsigned N, int relu, unsigned Hs, unsigned sph, unsigned box_rows, int lean,
                        const __grid_constant__ CUtensorMap hmap, const __grid_constant__ CUtensorMap xmap) {
  using RM = RowMap<H, C, true>;
  constexpr int F = RM::F, V = RM::V, LPR = RM::LPR, RPW = RM::RPW, LPH = RM::LPH;
  constexpr int kTileThreads = THREADS, kTileWarps = THREADS / 32;
  static_assert(!FUSE_MEAN || SLICED || (H == 1 && V == 1), "the fused mean follows the one-head layer");
  // SLICED: every chunk of a lane belongs to the slice's head (H = 2 only selects the packed lane map: 8 lanes x 2 chunks
  // per row, 4 rows per warp, as in the two-head nc = 32 kernel whose slab has the same shape)
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned HS = SLICED ? Hs : (unsigned)H;            // heads of the score tensors
  const unsigned nsl = SLICED ? Hs * sph : 1u;              // slices per row
  const unsigned ld = nsl * F;                              // row stride of h / out / x0 / xout
  const unsigned U = B * nsl;                               // work units
  const FwdTilePlan plan(N, F, HS, E1, FUSE_MEAN, FUSE_MEAN && lean);
  int* rp_s = reinterpret_cast<int*>(smem + plan.rp_off);
  int* col_s = reinterpret_cast<int*>(smem + plan.col_off);
  unsigned short* ord = reinterpret_cast<unsigned short*>(smem + plan.ord_off);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + plan.bar_off);          // [0], [1]: stages; [2]: x0 slab
  float* ZS = reinterpret_cast<float*>(smem + plan.zs_off);
  const float* XS = reinterpret_cast<const float*>(smem + plan.xs_off);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane / LPR, lig = lane % LPR, slot = lig % LPH;
  constexpr unsigned gmask = 0xffffffffu;          // control flow below is warp-uniform: full-mask shuffles

  auto issue = [&](int stage, unsigned u) {        // one elected thread; u = work unit (snapshot, or snapshot * nsl + slice)
    unsigned char* dst = smem + (size_t)stage * plan.stage_bytes;
    const unsigned bb = SLICED ? u / nsl : u, q = SLICED ? u % nsl : 0u;
    mbar_arrive_expect_tx(&full[stage], plan.tx_bytes);
    if (SLICED) {
      for (unsigned r0 = 0; r0 < N; r0 += box_rows)
        tma_load_2d(dst + (size_t)r0 * F * 4, &hmap, (int)(q * F), (int)(bb * N + r0), &full[stage]);
    } else {
      bulk_g2s(dst, h + (size_t)bb * N * F, plan.h_bytes, &full[stage]);
    }
    bulk_g2s(dst + plan.h_bytes, s_src + (size_t)bb * N * HS, plan.ss_bytes, &full[stage]);
    bulk_g2s(dst + plan.h_bytes + plan.ss_bytes, s_dst + (size_t)bb * N * HS, plan.ss_bytes, &full[stage]);
  };

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&full[2], 1);
    mbar_fence_init();
  }
  for (unsigned k = tid; k <= N; k += kTileThreads) rp_s[k] = __ldg(rowptr + k);
  for (unsigned k = tid; k < E1; k += kTileThreads) col_s[k] = __ldg(col + k);
  __syncthreads();
  degree_order<THREADS>(rp_s, N, ord, reinterpret_cast<int*>(smem + plan.hist_off));
  pdl_wait();                                     // CSR staging above overlapped the previous kernel's tail
  if (tid == 0) {
    if (blockIdx.x < U) issue(0, blockIdx.x);
    if (!lean && blockIdx.x + gridDim.x < U) issue(1, blockIdx.x + gridDim.x);
  }

  float4 bv[V];
  if (!SLICED) {
#pragma unroll
    for (int v = 0; v < V; ++v) bv[v] = ldg4(bias + 4 * RM::chunk(lig, v));
  }

  unsigned k = 0;
  if (gridDim.x >= U) pdl_launch_dependents();
  for (unsigned u = blockIdx.x; u < U; u += gridDim.x, ++k) {
    const unsigned b = SLICED ? u / nsl : u, q = SLICED ? u % nsl : 0u;
    const unsigned hq = SLICED ? q / sph : 0u;     // head of this slice
    const bool write_ml = !SLICED || q % sph == 0;
    const int stage = lean ? 0 : (int)(k & 1);
    const float* hs = reinterpret_cast<const float*>(smem + (size_t)stage * plan.stage_bytes) + 4 * lig;
    const float* sss = reinterpret_cast<const float*>(smem + (size_t)stage * plan.stage_bytes + plan.h_bytes);
    const float* sds = sss + N * HS;
    if (SLICED) {
#pragma unroll
      for (int v = 0; v < V; ++v) bv[v] = ldg4(bias + q * F + 4 * RM::chunk(lig, v));
    }
    if (FUSE_MEAN && !lean && tid == 0) {          // the previous snapshot's mean phase is done with the x0 slab
      fence_proxy_async();
      mbar_arrive_expect_tx(&full[2], plan.h_bytes);
      if (SLICED) {
        for (unsigned r0 = 0; r0 < N; r0 += box_rows)
          tma_load_2d(smem + plan.xs_off + (size_t)r0 * F * 4, &xmap, (int)(q * F), (int)(b * N + r0), &full[2]);
      } else {
        bulk_g2s(smem + plan.xs_off, x0 + (size_t)b * N * F, plan.h_bytes, &full[2]);
      }
    }
    mbar_wait(&full[stage], lean ? (k & 1) : ((k >> 1) & 1));

    for (unsigned i0 = warp * RPW; i0 < N; i0 += kTileWarps * RPW) {
      // rows past the end are clamped (recompute the last row, store nothing): every lane of the warp
      // runs the same instruction stream, so shuffles are plain full-mask SHFLs
      const bool row_ok = i0 + sub < N;
      const unsigned i = ord[row_ok ? i0 + sub : N - 1];             // rows in ascending degree
      const size_t r = (size_t)b * N + i;
      const int beg = rp_s[i], deg = rp_s[i + 1] - beg;
      const int deg_max = __reduce_max_sync(gmask, deg);
      float sd[V], mrun[V], lrun[V];
      float4 acc[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        sd[v] = sds[i * HS + (SLICED ? hq : (unsigned)RM::head(lig, v))];
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
          if (SLICED && v > 0) {                     // same head as chunk 0: same softmax
            if (e0 > 0) {
              const float sc = __expf(mrun[v] - mrun[0]);
              acc[v].x *= sc; acc[v].y *= sc; acc[v].z *= sc; acc[v].w *= sc;
            }
            p[v] = p[0]; lrun[v] = lrun[0]; mrun[v] = mrun[0];
            continue;
          }
          const float a = valid ? lrelu(sss[j * HS + (SLICED ? hq : (unsigned)RM::head(lig, v))] + sd[v]) : -CUDART_INF_F;
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
        if (FUSE_MEAN) st4(ZS + i * F + 4 * RM::chunk(lig, v), o);
        else st4(out + r * ld + q * F + 4 * RM::chunk(lig, v), o);
        if (m_out != nullptr && slot == 0 && write_ml && (!SLICED || v == 0)) {
          m_out[r * HS + (SLICED ? hq : (unsigned)RM::head(lig, v))] = mrun[v];
          l_out[r * HS + (SLICED ? hq : (unsigned)RM::head(lig, v))] = lrun[v];
        }
      }
    }
    __syncthreads();                               // every warp is done with this stage (and z is complete)
    if (tid == 0) {
      const unsigned nu = u + (lean ? 1u : 2u) * gridDim.x;
      if (nu < U) {
        fence_proxy_async();
        issue(stage, nu);
      }
    }
    if (FUSE_MEAN) {
      if (!lean) mbar_wait(&full[2], k & 1);
      for (unsigned io = warp * RPW + sub; io < N; io += kTileWarps * RPW) {
        const unsigned i = ord[io];
        const int beg = rp_s[i], end = rp_s[i + 1] - 1;                    // drop the self-loop
        float4 xg[V], acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          acc[v] = f4zero();
          xg[v] = lean ? ldg4_stream(x0 + ((size_t)b * N + i) * ld + q * F + 4 * RM::chunk(lig, v))   // flies under the neighbour sum
                       : *reinterpret_cast<const float4*>(XS + i * F + 4 * RM::chunk(lig, v));
        }
#pragma unroll 4
        for (int e = beg; e < end; ++e) {
          const int j = col_s[e];
#pragma unroll
          for (int v = 0; v < V; ++v) add4(acc[v], *reinterpret_cast<const float4*>(ZS + j * F + 4 * RM::chunk(lig, v)));
        }
        const int deg = end - beg;
        const float inv = 1.f / (float)(deg > 1 ? deg : 1);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float4 o;
          o.x = fmaxf(fmaf(acc[v].x, inv, xg[v].x), 0.f);
          o.y = fmaxf(fmaf(acc[v].y, inv, xg[v].y), 0.f);
          o.z = fmaxf(fmaf(acc[v].z, inv, xg[v].z), 0.f);
          o.w = fmaxf(fmaf(acc[v].w, inv, xg[v].w), 0.f);
          st4(xout + ((size_t)b * N + i) * ld + q * F + 4 * RM::chunk(lig, v), o);
        }
      }
      __syncthreads();                             // z and x0 slabs are free for the next snapshot
    }
  }
}

template <int H, int C, bool FUSE>
static int launch_fwd_tile(const int* rowptr, const int* col, unsigned E1, const float* h, const float* s_src,
                           const float* s_dst, const float* bias, float* out, float* m, float* l, const float* x0,
                           float* xout, unsigned B, unsigned N, int relu, cudaStream_t st) {
  const FwdTilePlan plan(N, H * C, H, E1, FUSE);
  unsigned per_sm = (unsigned)((227u * 1024u) / (plan.total + 1024u));
  per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
  unsigned grid = (unsigned)sm_count() * per_sm;
  if (grid > B) grid = B;
  CUtensorMap none;
  memset(&none, 0, sizeof(none));
#define LAUNCH(THR)                                                                                               \
  do {                                                                                                            \
    auto kern = gat_agg_fwd_tile_kernel<H, C, THR, FUSE, false>;                                                  \
    static uint32_t configured = 0;                                                                               \
    if (configured < plan.total) {                                                                                \
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total) != cudaSuccess) \
        return check_launch("gat_agg_fwd_tile: smem attribute");                                                  \
      configured = plan.total;                                                                                    \
    }                                                                                                             \
    launch_kernel(kern, dim3(grid), dim3(THR), plan.total, st, rowptr, col, E1, h, s_src, s_dst, bias, out, m, l, x0, xout, B, N, relu, \
                  (unsigned)H, 1u, 0u, 0, none, none);                                                            \
  } while (0)
  if (per_sm >= 2) LAUNCH(512); else LAUNCH(1024);
#undef LAUNCH
  return check_launch("gat_agg_fwd_tile");
}

// ---- channel-sliced work units (rows wider than a slab that fits; see the kernel's header) ----
static bool sliced_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GATRES_TILE_SLICED");
    v = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}
// rows per TMA box: the snapshot's N rows in equal boxes of at most 256 rows (0 = no such split)
static unsigned slice_box_rows(unsigned N) {
  const unsigned nbox = (N + 255u) / 256u;
  return N % nbox == 0 ? N / nbox : 0u;
}
// slice width (channels) for (H heads x C channels): 64 for the plain aggregation (the [N x 64] slab of the two-head nc = 32
// kernel), 32 with the fused mean (four slabs); 0 = the sliced form does not apply
static unsigned slice_width(unsigned N, unsigned H, unsigned C, unsigned E1, bool fuse_mean) {
  const unsigned CS = fuse_mean ? 32u : 64u;
  if (!sliced_enabled() || C % CS != 0 || C <= CS || slice_box_rows(N) == 0 || (N * H) % 4u != 0) return 0;
  if (encode_tiled_fn() == nullptr) return 0;
  const FwdTilePlan plan(N, CS, H, E1, fuse_mean);
  return plan.total <= 227u * 1024u ? CS : 0u;
}

template <int CS, bool FUSE, bool LEAN = false>
static int launch_fwd_tile_sliced(const int* rowptr, const int* col, unsigned E1, const float* h, const float* s_src,
                                  const float* s_dst, const float* bias, float* out, float* m, float* l, const float* x0,
                                  float* xout, unsigned B, unsigned N, unsigned H, unsigned C, int relu, cudaStream_t st) {
  const FwdTilePlan plan(N, CS, H, E1, FUSE, LEAN);
  const unsigned sph = C / CS, nsl = H * sph, box_rows = slice_box_rows(N);
  CUtensorMap hmap, xmap;
  memset(&xmap, 0, sizeof(xmap));
  if (!make_map_2d(&hmap, h, (unsigned long long)B * N, H * C, CS, box_rows, false) ||
      (FUSE && !make_map_2d(&xmap, x0, (unsigned long long)B * N, H * C, CS, box_rows, false))) {
    set_error("gat_agg_fwd_tile: cuTensorMapEncodeTiled failed");
    return GATRES_ERR_CUDA;
  }
  unsigned per_sm = (unsigned)((227u * 1024u) / (plan.total + 1024u));
  per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
  unsigned grid = (unsigned)sm_count() * per_sm;
  if (grid > B * nsl) grid = B * nsl;
#define LAUNCH(THR)                                                                                               \
  do {                                                                                                            \
    auto kern = gat_agg_fwd_tile_kernel<(CS == 64 ? 2 : 1), (CS == 64 ? 32 : CS), THR, FUSE, true>;                                                  \
    static uint32_t configured = 0;                                                                               \
    if (configured < plan.total) {                                                                                \
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total) != cudaSuccess) \
        return check_launch("gat_agg_fwd_tile (sliced): smem attribute");                                         \
      configured = plan.total;                                                                                    \
    }                                                                                                             \
    launch_kernel(kern, dim3(grid), dim3(THR), plan.total, st, rowptr, col, E1, h, s_src, s_dst, bias, out, m, l, x0, xout, B, N, relu, \
                  H, sph, box_rows, (int)LEAN, hmap, xmap);                                                       \
  } while (0)
  if (per_sm >= 2) LAUNCH(512); else LAUNCH(1024);
#undef LAUNCH
  return check_launch("gat_agg_fwd_tile (sliced)");
}

// Eligibility: slab + scores double-buffered + CSR must fit, and the per-snapshot byte
// counts must be 16 B multiples (bulk-copy granularity).
// conv2 aggregation + SimpleConv(mean) + residual + ReLU in one launch (one head; nc = 32 / 64 whole rows, wider rows in
// 32-channel slices)
static bool fwd_tile_mean_plain(unsigned N, unsigned C, unsigned E1) {
  const FwdTilePlan plan(N, C, 1, E1, true);
  return N % 4u == 0 && plan.total <= 227u * 1024u && (C == 32 || C == 64);
}
bool fwd_tile_mean_eligible(unsigned N, unsigned C, unsigned E1) {
  return fwd_tile_mean_plain(N, C, E1) || slice_width(N, 1, C, E1, true) != 0;
}

int gat_agg_mean_res_fwd_tile(const int* rowptr, const int* col, unsigned E1, const float* h, const float* s_src,
                              const float* s_dst, const float* bias, float* m, float* l, const float* x0, float* xout,
                              unsigned B, unsigned N, int C, cudaStream_t st) {
  if (fwd_tile_mean_plain(N, (unsigned)C, E1)) {
    if (C == 32) return launch_fwd_tile<1, 32, true>(rowptr, col, E1, h, s_src, s_dst, bias, nullptr, m, l, x0, xout, B, N, 0, st);
    if (C == 64) return launch_fwd_tile<1, 64, true>(rowptr, col, E1, h, s_src, s_dst, bias, nullptr, m, l, x0, xout, B, N, 0, st);
  }
  static int lean = -1;
  if (lean < 0) {
    const char* e = getenv("GATRES_TILE_LEAN");
    lean = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  if (lean && slice_width(N, 1, (unsigned)C, E1, true) != 0 && C % 64 == 0 && C > 64 &&
      FwdTilePlan(N, 64, 1, E1, true, true).total <= 227u * 1024u)
    return launch_fwd_tile_sliced<64, true, true>(rowptr, col, E1, h, s_src, s_dst, bias, nullptr, m, l, x0, xout, B, N, 1u, (unsigned)C, 0, st);
  if (slice_width(N, 1, (unsigned)C, E1, true) == 32)
    return launch_fwd_tile_sliced<32, true>(rowptr, col, E1, h, s_src, s_dst, bias, nullptr, m, l, x0, xout, B, N, 1u, (unsigned)C, 0, st);
  set_error("gat_agg_mean_res_fwd_tile: unsupported channels %d", C);
  return GATRES_ERR_ARG;
}

static bool fwd_tile_plain(unsigned N, unsigned H, unsigned C, unsigned E1) {
  const FwdTilePlan plan(N, H * C, H, E1);
  return (N * H) % 4u == 0 && plan.total <= 227u * 1024u && (H * C) <= 128;
}
bool fwd_tile_eligible(unsigned N, unsigned H, unsigned C, unsigned E1) {
  return fwd_tile_plain(N, H, C, E1) || slice_width(N, H, C, E1, false) != 0;
}

int gat_agg_fwd_tile(const int* rowptr, const int* col, unsigned E1, const float* h, const float* s_src,
                     const float* s_dst, const float* bias, float* out, float* m, float* l, unsigned B, unsigned N,
                     int H, int C, int relu, cudaStream_t st) {
  if (fwd_tile_plain(N, (unsigned)H, (unsigned)C, E1)) {
#define T(HH, CC) \
  if (H == HH && C == CC) return launch_fwd_tile<HH, CC, false>(rowptr, col, E1, h, s_src, s_dst, bias, out, m, l, nullptr, nullptr, B, N, relu, st)
    T(1, 32);
    T(2, 32);
    T(1, 64);
    T(2, 64);
    T(1, 128);
#undef T
  }
  if (slice_width(N, (unsigned)H, (unsigned)C, E1, false) == 64)
    return launch_fwd_tile_sliced<64, false>(rowptr, col, E1, h, s_src, s_dst, bias, out, m, l, nullptr, nullptr, B, N,
                                             (unsigned)H, (unsigned)C, relu, st);
  set_error("gat_agg_fwd_tile: unsupported (H=%d, C=%d)", H, C);
  return GATRES_ERR_ARG;
}

}  // namespace gatres

// =============================================================================
// Fused backward (pass 1 + pass 2 of gat_agg.cu) for one snapshot per CTA.
// Both feature slabs of the snapshot (h and g = dL/d(out)) and its per-node scalars
// are staged by TMA; pass 1 leaves ds_dst in shared memory and scatters every edge's
// ds_src term to its source with a shared-memory atomic, so the `rec` round trip
// through HBM and one kernel launch disappear, h / g are read from DRAM once instead
// of twice (12*S + 28*H algorithmic bytes per node instead of 20*S + 52*H), and pass 2
// is a pure alpha-weighted gather of g rows (no second evaluation of <g_i, h_j>).
// =============================================================================
namespace gatres {

struct BwdTilePlan {
  // [inputs of a snapshot: h slab, g slab, s_src, s_dst, m, l] x nbuf, then D, ds_dst, the two CSRs, barriers
  uint32_t slab, sc, in_bytes, nbuf, hs_off, gs_off, ss_off, sd_off, mm_off, ll_off, dd_off, dsd_off, rpi_off, ci_off,
      rpo_off, co_off, oi_off, oo_off, hist_off, bar_off, total, tx_bytes;
  __host__ __device__ BwdTilePlan(unsigned N, unsigned F, unsigned H, unsigned E1, unsigned nbuf_ = 1) {
    slab = N * F * 4u;
    sc = N * H * 4u;
    nbuf = nbuf_;
    hs_off = 0;
    gs_off = slab;
    ss_off = 2u * slab;
    sd_off = ss_off + sc;
    mm_off = sd_off + sc;
    ll_off = mm_off + sc;
    in_bytes = (ll_off + sc + 127u) & ~127u;              // stride between the input buffers
    dd_off = nbuf * in_bytes;
    dsd_off = dd_off + sc;
    rpi_off = dsd_off + sc;
    const uint32_t rp = ((N + 1u) * 4u + 15u) & ~15u, cl = (E1 * 2u + 15u) & ~15u;
    ci_off = rpi_off + rp;
    rpo_off = ci_off + cl;
    co_off = rpo_off + rp;
    oi_off = co_off + cl;                                 // degree orders of the rows: in-degree (pass 1), out-degree (pass 2)
    oo_off = oi_off + ((N * 2u + 15u) & ~15u);
    hist_off = oo_off + ((N * 2u + 15u) & ~15u);
    bar_off = hist_off + 128u;
    total = bar_off + 16u;
    tx_bytes = 2u * slab + 4u * sc;
  }
};

// PACK (see RowMap): used when one CTA per SM runs 512 threads with a 128-register budget (two heads, nc = 32:
// the two slabs fill shared memory); the 2-CTA / 1024-thread shapes have 64 registers per thread and keep PACK off.
// NBUF = 2 (one head, nc = 32: two input sets fit one SM): the next snapshot's slabs are loaded while the two passes
// of the current one run; with NBUF = 1 the load of a snapshot starts when the passes of the previous one are done.
template <int H, int C, int THREADS, bool PACK, int NBUF>
__global__ void __launch_bounds__(THREADS, (THREADS == 512 && !PACK && NBUF == 1) ? 2 : 1)
gat_agg_bwd_tile_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                        const int* __restrict__ rowptr_t, const int* __restrict__ col_t, unsigned E1,
                        const float* __restrict__ g, const float* __restrict__ h,
                        const float* __restrict__ s_src, const float* __restrict__ s_dst,
                        const float* __restrict__ m, const float* __restrict__ l,
                        const float* __restrict__ att_src, const float* __restrict__ att_dst,
                        float* __restrict__ dh, float* __restrict__ grads,
                        long long off_att_src, long long off_att_dst, long long off_bias,
                        unsigned B, unsigned N) {
  using RM = RowMap<H, C, PACK>;
  constexpr int F = RM::F, V = RM::V, LPR = RM::LPR, RPW = RM::RPW, LPH = RM::LPH;
  constexpr int kWarpsT = THREADS / 32;
  constexpr unsigned gmask = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem[];
  const BwdTilePlan plan(N, F, H, E1, NBUF);
  const float *HS, *GS, *SS, *SD, *MM, *LL;                 // inputs of the current snapshot (set per iteration)
  float* DD = reinterpret_cast<float*>(smem + plan.dd_off);
  float* DSD = reinterpret_cast<float*>(smem + plan.dsd_off);
  int* rpi = reinterpret_cast<int*>(smem + plan.rpi_off);
  unsigned short* ci = reinterpret_cast<unsigned short*>(smem + plan.ci_off);
  int* rpo = reinterpret_cast<int*>(smem + plan.rpo_off);
  unsigned short* co = reinterpret_cast<unsigned short*>(smem + plan.co_off);
  unsigned short* ord_in = reinterpret_cast<unsigned short*>(smem + plan.oi_off);
  unsigned short* ord_out = reinterpret_cast<unsigned short*>(smem + plan.oo_off);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + plan.bar_off);        // [NBUF]
  float* red = reinterpret_cast<float*>(smem);     // reused for the final CTA reduction (slabs are dead by then)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane / LPR, lig = lane % LPR, slot = lig % LPH;

  auto issue = [&](unsigned bb, unsigned buf) {
    const size_t ro = (size_t)bb * N;
    unsigned char* base = smem + buf * plan.in_bytes;
    mbar_arrive_expect_tx(full + buf, plan.tx_bytes);
    bulk_g2s(base + plan.hs_off, h + ro * F, plan.slab, full + buf);
    bulk_g2s(base + plan.gs_off, g + ro * F, plan.slab, full + buf);
    bulk_g2s(base + plan.ss_off, s_src + ro * H, plan.sc, full + buf);
    bulk_g2s(base + plan.sd_off, s_dst + ro * H, plan.sc, full + buf);
    bulk_g2s(base + plan.mm_off, m + ro * H, plan.sc, full + buf);
    bulk_g2s(base + plan.ll_off, l + ro * H, plan.sc, full + buf);
  };

  if (tid == 0) {
    for (int q = 0; q < NBUF; ++q) mbar_init(full + q, 1);
    mbar_fence_init();
  }
  for (unsigned k = tid; k <= N; k += THREADS) { rpi[k] = __ldg(rowptr + k); rpo[k] = __ldg(rowptr_t + k); }
  for (unsigned k = tid; k < E1; k += THREADS) {
    ci[k] = (unsigned short)__ldg(col + k);
    co[k] = (unsigned short)__ldg(col_t + k);
  }
  float4 as[V], ad[V], accs[V], accd[V], bacc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    as[v] = ldg4(att_src + 4 * RM::chunk(lig, v));
    ad[v] = ldg4(att_dst + 4 * RM::chunk(lig, v));
    accs[v] = accd[v] = bacc[v] = f4zero();
  }
  // DD accumulates ds_src[j] = sum over the out-edges of j of dz_e: pass 1 (which owns dalpha_e and D_i of every edge)
  // adds each edge's term with a shared-memory atomic, pass 2 reads the sum and clears it for the next snapshot —
  // so pass 2 needs neither the per-edge dot products <g_i, h_j> again nor D of the edges' targets
  for (unsigned k = tid; k < N * H; k += THREADS) DD[k] = 0.f;
  __syncthreads();
  degree_order<THREADS>(rpi, N, ord_in, reinterpret_cast<int*>(smem + plan.hist_off));
  degree_order<THREADS>(rpo, N, ord_out, reinterpret_cast<int*>(smem + plan.hist_off));
  pdl_wait();
  if (tid == 0 && blockIdx.x < B) issue(blockIdx.x, 0);

  unsigned it = 0;
  if (gridDim.x >= B) pdl_launch_dependents();
  for (unsigned b = blockIdx.x; b < B; b += gridDim.x, ++it) {
    const unsigned cur = NBUF == 2 ? (it & 1u) : 0u;
    if (NBUF == 2 && tid == 0 && b + gridDim.x < B) {       // that buffer's snapshot finished last iteration
      fence_proxy_async();
      issue(b + gridDim.x, cur ^ 1u);
    }
    {
      const unsigned char* base = smem + cur * plan.in_bytes;
      HS = reinterpret_cast<const float*>(base + plan.hs_off);
      GS = reinterpret_cast<const float*>(base + plan.gs_off);
      SS = reinterpret_cast<const float*>(base + plan.ss_off);
      SD = reinterpret_cast<const float*>(base + plan.sd_off);
      MM = reinterpret_cast<const float*>(base + plan.mm_off);
      LL = reinterpret_cast<const float*>(base + plan.ll_off);
    }
    mbar_wait(full + cur, NBUF == 2 ? ((it >> 1) & 1u) : (it & 1u));

    // ---------------- pass 1: per target row, D and ds_dst into shared memory
    for (unsigned i0 = warp * RPW; i0 < N; i0 += kWarpsT * RPW) {
      const bool row_ok = i0 + sub < N;
      const unsigned i = ord_in[row_ok ? i0 + sub : N - 1];          // rows in ascending in-degree
      const int beg = rpi[i], deg = rpi[i + 1] - beg;
      const int deg_max = __reduce_max_sync(gmask, deg);
      float4 gv[V];
      float sd[V], mi[V], il[V], S1[V], S2[V], S3[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int hd = RM::head(lig, v);
        gv[v] = *reinterpret_cast<const float4*>(GS + i * F + 4 * RM::chunk(lig, v));
        if (row_ok) add4(bacc[v], gv[v]);
        sd[v] = SD[i * H + hd];
        mi[v] = MM[i * H + hd];
        il[v] = 1.f / (LL[i * H + hd] + kSoftmaxEps);
        S1[v] = S2[v] = S3[v] = 0.f;
      }
      // one chunk of up to LPH in-edges: lane `slot` owns edge e0 + slot (alpha, LeakyReLU slope, dalpha = <g_i, h_j>)
      auto chunk = [&](int e0, int& j, float (&alpha)[V], float (&sl)[V], float (&da)[V]) {
        const bool valid = e0 + slot < deg;
        j = valid ? (int)ci[beg + e0 + slot] : 0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float z = SS[j * H + RM::head(lig, v)] + sd[v];
          alpha[v] = valid ? __expf(lrelu(z) - mi[v]) * il[v] : 0.f;
          sl[v] = lrelu_slope(z);
          da[v] = 0.f;
        }
        const int cnt_max = min(LPH, deg_max - e0);
        for (int t = 0; t < cnt_max; ++t) {
          const int jt = __shfl_sync(gmask, j, t, LPH);
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float4 x = *reinterpret_cast<const float4*>(HS + jt * F + 4 * RM::chunk(lig, v));
            const float d = group_sum<LPH>(dot4(gv[v], x), gmask);
            da[v] = slot == t ? d : da[v];
          }
        }
      };
      int j;
      float alpha[V], sl[V], da[V];
      for (int e0 = 0; e0 < deg_max; e0 += LPH) {
        chunk(e0, j, alpha, sl, da);
#pragma unroll
        for (int v = 0; v < V; ++v) {             // idle slots have alpha = 0
          S1[v] = fmaf(alpha[v], da[v], S1[v]);
          S2[v] = fmaf(alpha[v] * sl[v], da[v], S2[v]);
          S3[v] = fmaf(alpha[v], sl[v], S3[v]);
        }
      }
      float Dv[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        Dv[v] = group_sum<LPH>(S1[v], gmask);
        const float T2 = group_sum<LPH>(S2[v], gmask), T3 = group_sum<LPH>(S3[v], gmask);
        if (slot == 0 && row_ok) DSD[i * H + RM::head(lig, v)] = T2 - Dv[v] * T3;
      }
      // ds_src contribution of every in-edge of this row: dz_e = alpha_e (dalpha_e - D_i) slope_e, added to its SOURCE
      for (int e0 = 0; e0 < deg_max; e0 += LPH) {
        if (deg_max > LPH) chunk(e0, j, alpha, sl, da);        // rows with more in-edges than lanes: recompute the chunk
        if (row_ok && e0 + slot < deg) {
#pragma unroll
          for (int v = 0; v < V; ++v)                          // every (edge, head) is held by exactly one lane
            atomicAdd(DD + j * H + RM::head(lig, v), alpha[v] * sl[v] * (da[v] - Dv[v]));
        }
      }
    }
    __syncthreads();

    // ---------------- pass 2: per source row, dh and the attention-vector gradients
    for (unsigned j0 = warp * RPW; j0 < N; j0 += kWarpsT * RPW) {
      const bool row_ok = j0 + sub < N;
      const unsigned jn = ord_out[row_ok ? j0 + sub : N - 1];        // rows in ascending out-degree
      const int beg = rpo[jn], deg = rpo[jn + 1] - beg;
      const int deg_max = __reduce_max_sync(gmask, deg);
      float4 hv[V], dacc[V];
      float ss[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        hv[v] = *reinterpret_cast<const float4*>(HS + jn * F + 4 * RM::chunk(lig, v));
        ss[v] = SS[jn * H + RM::head(lig, v)];
        dacc[v] = f4zero();
      }
      for (int e0 = 0; e0 < deg_max; e0 += LPH) {
        const bool valid = e0 + slot < deg;
        const int i = valid ? (int)co[beg + e0 + slot] : 0;
        float alpha[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const int o = i * H + RM::head(lig, v);
          const float z = ss[v] + SD[o];
          alpha[v] = valid ? __fdividef(__expf(lrelu(z) - MM[o]), LL[o] + kSoftmaxEps) : 0.f;
        }
        const int cnt_max = min(LPH, deg_max - e0);
        for (int t = 0; t < cnt_max; ++t) {
          const int itg = __shfl_sync(gmask, i, t, LPH);
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float4 gx = *reinterpret_cast<const float4*>(GS + itg * F + 4 * RM::chunk(lig, v));
            fma4(dacc[v], __shfl_sync(gmask, alpha[v], t, LPH), gx);
          }
        }
      }
      float dsv[V];
#pragma unroll
      for (int v = 0; v < V; ++v) dsv[v] = DD[jn * H + RM::head(lig, v)];      // ds_src[j], accumulated by pass 1
      __syncwarp();
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (slot == 0 && row_ok) DD[jn * H + RM::head(lig, v)] = 0.f;          // cleared for the next snapshot
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float ds = dsv[v];
        const float dd = DSD[jn * H + RM::head(lig, v)];
        fma4(dacc[v], ds, as[v]);
        fma4(dacc[v], dd, ad[v]);
        if (row_ok) {
          st4(dh + ((size_t)b * N + jn) * F + 4 * RM::chunk(lig, v), dacc[v]);
          fma4(accs[v], ds, hv[v]);
          fma4(accd[v], dd, hv[v]);
        }
      }
    }
    __syncthreads();                               // both slabs are free again
    if (NBUF == 1 && tid == 0) {
      const unsigned nb = b + gridDim.x;
      if (nb < B) {
        fence_proxy_async();
        issue(nb, 0);
      }
    }
  }

  // ---------------- parameter gradients: CTA reduction, then one atomic add per column chunk
#pragma unroll
  for (int v = 0; v < V; ++v) {
    float4 accv[3] = {accs[v], accd[v], bacc[v]};
    const long long offs[3] = {off_att_src, off_att_dst, off_bias};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      __syncthreads();
      st4(red + tid * 4, accv[q]);
      __syncthreads();
      if (tid < LPR) {
        float4 s = f4zero();
        for (int w = 0; w < kWarpsT; ++w)
#pragma unroll
          for (int sb = 0; sb < 32 / LPR; ++sb) add4(s, *reinterpret_cast<float4*>(red + (w * 32 + sb * LPR + tid) * 4));
        atomicAdd(reinterpret_cast<float4*>(grads + offs[q] + 4 * (tid + v * LPR)), s);
      }
    }
  }
}

// -----------------------------------------------------------------------------
// Pipelined form of the fused backward for shapes whose two slabs fill shared memory (two heads x 32 channels, one head
// x 64 channels on the C-Town-sized graphs): with one input set per CTA the kernel above loads a snapshot (213 KB, ~7 us
// at one SM's share of HBM) and only then runs its two passes (~10 us) — the copy engine and the warps take turns.  Here
// the two halves of the input set travel separately and each flies under the OTHER pass:
//   stage H = {h slab, s_src}        requested when pass 1 of the previous snapshot is done (pass 2 no longer reads the
//                                    h slab: the row's own h — needed only at the end of its iteration, for the
//                                    attention-vector gradients — is a global load issued at the top of the iteration
//                                    and served by L2, where the TMA copy put the slab microseconds earlier; s_src is
//                                    double-buffered, 3 KB);
//   stage G = {g slab, s_dst, m, l}  requested when pass 2 of the previous snapshot is done, in row chunks of one
//                                    warp-iteration each with one mbarrier per chunk: pass 1 walks the chunks in order
//                                    (degree-sorted inside a chunk) and only waits for the chunk it is about to read.
// Same arithmetic, same results as the kernel above (the tests run both).
// -----------------------------------------------------------------------------
constexpr int kPipeMaxChunks = 8;

struct BwdPipePlan {
  uint32_t slab, sc, hs_off, gs_off, ss_off, sd_off, mm_off, ll_off, dd_off, dsd_off, rpi_off, ci_off, rpo_off, co_off,
      oi_off, oo_off, hist_off, bar_off, total;
  __host__ __device__ BwdPipePlan(unsigned N, unsigned F, unsigned H, unsigned E1) {
    slab = N * F * 4u;
    sc = N * H * 4u;
    hs_off = 0;
    gs_off = slab;
    ss_off = 2u * slab;                                   // two s_src buffers
    sd_off = ss_off + 2u * sc;
    mm_off = sd_off + sc;
    ll_off = mm_off + sc;
    dd_off = (ll_off + sc + 15u) & ~15u;
    dsd_off = dd_off + sc;
    rpi_off = dsd_off + sc;
    const uint32_t rp = ((N + 1u) * 4u + 15u) & ~15u, cl = (E1 * 2u + 15u) & ~15u;
    ci_off = rpi_off + rp;
    rpo_off = ci_off + cl;
    co_off = rpo_off + rp;
    oi_off = co_off + cl;
    oo_off = oi_off + ((N * 2u + 15u) & ~15u);
    hist_off = oo_off + ((N * 2u + 15u) & ~15u);
    bar_off = hist_off + 128u;
    total = bar_off + 8u * (1u + kPipeMaxChunks);
  }
};

template <int H, int C, int THREADS, bool PACK>
__global__ void __launch_bounds__(THREADS, 1)
gat_agg_bwd_tile_pipe_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                             const int* __restrict__ rowptr_t, const int* __restrict__ col_t, unsigned E1,
                             const float* __restrict__ g, const float* __restrict__ h,
                             const float* __restrict__ s_src, const float* __restrict__ s_dst,
                             const float* __restrict__ m, const float* __restrict__ l,
                             const float* __restrict__ att_src, const float* __restrict__ att_dst,
                             float* __restrict__ dh, float* __restrict__ grads,
                             long long off_att_src, long long off_att_dst, long long off_bias,
                             unsigned B, unsigned N) {
  using RM = RowMap<H, C, PACK>;
  constexpr int F = RM::F, V = RM::V, LPR = RM::LPR, RPW = RM::RPW, LPH = RM::LPH;
  constexpr int kWarpsT = THREADS / 32;
  constexpr unsigned RC = kWarpsT * RPW;                  // rows per chunk = rows of one warp-iteration of the CTA
  constexpr unsigned gmask = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem[];
  const BwdPipePlan plan(N, F, H, E1);
  const float* HS = reinterpret_cast<const float*>(smem + plan.hs_off);
  const float* GS = reinterpret_cast<const float*>(smem + plan.gs_off);
  const float* SD = reinterpret_cast<const float*>(smem + plan.sd_off);
  const float* MM = reinterpret_cast<const float*>(smem + plan.mm_off);
  const float* LL = reinterpret_cast<const float*>(smem + plan.ll_off);
  float* DD = reinterpret_cast<float*>(smem + plan.dd_off);
  float* DSD = reinterpret_cast<float*>(smem + plan.dsd_off);
  int* rpi = reinterpret_cast<int*>(smem + plan.rpi_off);
  unsigned short* ci = reinterpret_cast<unsigned short*>(smem + plan.ci_off);
  int* rpo = reinterpret_cast<int*>(smem + plan.rpo_off);
  unsigned short* co = reinterpret_cast<unsigned short*>(smem + plan.co_off);
  unsigned short* ord_in = reinterpret_cast<unsigned short*>(smem + plan.oi_off);
  unsigned short* ord_out = reinterpret_cast<unsigned short*>(smem + plan.oo_off);
  uint64_t* bar_h = reinterpret_cast<uint64_t*>(smem + plan.bar_off);
  uint64_t* bar_g = bar_h + 1;                            // [nch]
  float* red = reinterpret_cast<float*>(smem);            // final CTA reduction (slabs are dead by then)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane / LPR, lig = lane % LPR, slot = lig % LPH;
  const unsigned nch = (N + RC - 1) / RC;

  auto issue_h = [&](unsigned bb, unsigned buf) {         // one thread
    const size_t ro = (size_t)bb * N;
    mbar_arrive_expect_tx(bar_h, plan.slab + plan.sc);
    bulk_g2s(smem + plan.hs_off, h + ro * F, plan.slab, bar_h);
    bulk_g2s(smem + plan.ss_off + buf * plan.sc, s_src + ro * H, plan.sc, bar_h);
  };
  auto issue_g = [&](unsigned bb) {                       // one thread; the small arrays ride with the first chunk
    const size_t ro = (size_t)bb * N;
    for (unsigned p = 0; p < nch; ++p) {
      const unsigned r0 = p * RC, rows = min(RC, N - r0), bytes = rows * F * 4u;
      mbar_arrive_expect_tx(bar_g + p, bytes + (p == 0 ? 3u * plan.sc : 0u));
      if (p == 0) {
        bulk_g2s(smem + plan.sd_off, s_dst + ro * H, plan.sc, bar_g);
        bulk_g2s(smem + plan.mm_off, m + ro * H, plan.sc, bar_g);
        bulk_g2s(smem + plan.ll_off, l + ro * H, plan.sc, bar_g);
      }
      bulk_g2s(smem + plan.gs_off + (size_t)r0 * F * 4u, g + (ro + r0) * F, bytes, bar_g + p);
    }
  };
  constexpr unsigned kBarCount = 1;
  const bool issuer = tid == 0;

  if (tid == 0) {
    mbar_init(bar_h, kBarCount);
    for (unsigned p = 0; p < nch; ++p) mbar_init(bar_g + p, kBarCount);
    mbar_fence_init();
  }
  for (unsigned k = tid; k <= N; k += THREADS) { rpi[k] = __ldg(rowptr + k); rpo[k] = __ldg(rowptr_t + k); }
  for (unsigned k = tid; k < E1; k += THREADS) {
    ci[k] = (unsigned short)__ldg(col + k);
    co[k] = (unsigned short)__ldg(col_t + k);
  }
  float4 as[V], ad[V], accs[V], accd[V], bacc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    as[v] = ldg4(att_src + 4 * RM::chunk(lig, v));
    ad[v] = ldg4(att_dst + 4 * RM::chunk(lig, v));
    accs[v] = accd[v] = bacc[v] = f4zero();
  }
  for (unsigned k = tid; k < N * H; k += THREADS) DD[k] = 0.f;
  __syncthreads();
  // pass 1 walks the row chunks in order: ascending in-degree inside every chunk; pass 2: ascending out-degree overall
  for (unsigned p = 0; p < nch; ++p)
    degree_order<THREADS>(rpi + p * RC, min(RC, N - p * RC), ord_in + p * RC, reinterpret_cast<int*>(smem + plan.hist_off));
  degree_order<THREADS>(rpo, N, ord_out, reinterpret_cast<int*>(smem + plan.hist_off));
  pdl_wait();
  if (issuer && blockIdx.x < B) {
    issue_h(blockIdx.x, 0);
    issue_g(blockIdx.x);
  }

  unsigned it = 0;
  if (gridDim.x >= B) pdl_launch_dependents();
  for (unsigned b = blockIdx.x; b < B; b += gridDim.x, ++it) {
    const unsigned cur = it & 1u, par = it & 1u;
    const float* SS = reinterpret_cast<const float*>(smem + plan.ss_off + cur * plan.sc);
    mbar_wait(bar_h, par);

    // ---------------- pass 1: per target row, D and ds_dst into shared memory, ds_src scattered to the sources
    for (unsigned p = 0; p < nch; ++p) {
      mbar_wait(bar_g + p, par);
      const unsigned kk = p * RC + warp * RPW + sub;
      const bool row_ok = kk < N;
      const unsigned i = p * RC + ord_in[row_ok ? kk : p * RC];        // (the order is local to the chunk)
      const int beg = rpi[i], deg = rpi[i + 1] - beg;
      const int deg_max = __reduce_max_sync(gmask, deg);
      float4 gv[V];
      float sd[V], mi[V], il[V], S1[V], S2[V], S3[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int hd = RM::head(lig, v);
        gv[v] = *reinterpret_cast<const float4*>(GS + i * F + 4 * RM::chunk(lig, v));
        if (row_ok) add4(bacc[v], gv[v]);
        sd[v] = SD[i * H + hd];
        mi[v] = MM[i * H + hd];
        il[v] = 1.f / (LL[i * H + hd] + kSoftmaxEps);
        S1[v] = S2[v] = S3[v] = 0.f;
      }
      auto chunk = [&](int e0, int& j, float (&alpha)[V], float (&sl)[V], float (&da)[V]) {
        const bool valid = e0 + slot < deg;
        j = valid ? (int)ci[beg + e0 + slot] : 0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float z = SS[j * H + RM::head(lig, v)] + sd[v];
          alpha[v] = valid ? __expf(lrelu(z) - mi[v]) * il[v] : 0.f;
          sl[v] = lrelu_slope(z);
          da[v] = 0.f;
        }
        const int cnt_max = min(LPH, deg_max - e0);
        for (int t = 0; t < cnt_max; ++t) {
          const int jt = __shfl_sync(gmask, j, t, LPH);
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float4 x = *reinterpret_cast<const float4*>(HS + jt * F + 4 * RM::chunk(lig, v));
            const float d = group_sum<LPH>(dot4(gv[v], x), gmask);
            da[v] = slot == t ? d : da[v];
          }
        }
      };
      int j;
      float alpha[V], sl[V], da[V];
      for (int e0 = 0; e0 < deg_max; e0 += LPH) {
        chunk(e0, j, alpha, sl, da);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          S1[v] = fmaf(alpha[v], da[v], S1[v]);
          S2[v] = fmaf(alpha[v] * sl[v], da[v], S2[v]);
          S3[v] = fmaf(alpha[v], sl[v], S3[v]);
        }
      }
      float Dv[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        Dv[v] = group_sum<LPH>(S1[v], gmask);
        const float T2 = group_sum<LPH>(S2[v], gmask), T3 = group_sum<LPH>(S3[v], gmask);
        if (slot == 0 && row_ok) DSD[i * H + RM::head(lig, v)] = T2 - Dv[v] * T3;
      }
      for (int e0 = 0; e0 < deg_max; e0 += LPH) {
        if (deg_max > LPH) chunk(e0, j, alpha, sl, da);
        if (row_ok && e0 + slot < deg) {
#pragma unroll
          for (int v = 0; v < V; ++v)
            atomicAdd(DD + j * H + RM::head(lig, v), alpha[v] * sl[v] * (da[v] - Dv[v]));
        }
      }
    }
    __syncthreads();                               // the h slab and this s_src buffer's readers of pass 1 are done
    if (issuer && b + gridDim.x < B) {             // stage H of the next snapshot flies under pass 2
      fence_proxy_async();
      issue_h(b + gridDim.x, cur ^ 1u);
    }

    // ---------------- pass 2: per source row, dh and the attention-vector gradients
    for (unsigned j0 = warp * RPW; j0 < N; j0 += kWarpsT * RPW) {
      const bool row_ok = j0 + sub < N;
      const unsigned jn = ord_out[row_ok ? j0 + sub : N - 1];
      const int beg = rpo[jn], deg = rpo[jn + 1] - beg;
      const int deg_max = __reduce_max_sync(gmask, deg);
      float4 hv[V], dacc[V];
      float ss[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        // own row of h: needed only at the end of the iteration -> global load (L2), the slab belongs to the next snapshot
        hv[v] = ldg4(h + ((size_t)b * N + jn) * F + 4 * RM::chunk(lig, v));
        ss[v] = SS[jn * H + RM::head(lig, v)];
        dacc[v] = f4zero();
      }
      for (int e0 = 0; e0 < deg_max; e0 += LPH) {
        const bool valid = e0 + slot < deg;
        const int i = valid ? (int)co[beg + e0 + slot] : 0;
        float alpha[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const int o = i * H + RM::head(lig, v);
          const float z = ss[v] + SD[o];
          alpha[v] = valid ? __fdividef(__expf(lrelu(z) - MM[o]), LL[o] + kSoftmaxEps) : 0.f;
        }
        const int cnt_max = min(LPH, deg_max - e0);
        for (int t = 0; t < cnt_max; ++t) {
          const int itg = __shfl_sync(gmask, i, t, LPH);
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float4 gx = *reinterpret_cast<const float4*>(GS + itg * F + 4 * RM::chunk(lig, v));
            fma4(dacc[v], __shfl_sync(gmask, alpha[v], t, LPH), gx);
          }
        }
      }
      float dsv[V];
#pragma unroll
      for (int v = 0; v < V; ++v) dsv[v] = DD[jn * H + RM::head(lig, v)];
      __syncwarp();
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (slot == 0 && row_ok) DD[jn * H + RM::head(lig, v)] = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float ds = dsv[v];
        const float dd = DSD[jn * H + RM::head(lig, v)];
        fma4(dacc[v], ds, as[v]);
        fma4(dacc[v], dd, ad[v]);
        if (row_ok) {
          st4(dh + ((size_t)b * N + jn) * F + 4 * RM::chunk(lig, v), dacc[v]);
          fma4(accs[v], ds, hv[v]);
          fma4(accd[v], dd, hv[v]);
        }
      }
    }
    __syncthreads();                               // the g slab and s_dst / m / l are free again
    if (issuer && b + gridDim.x < B) {             // stage G of the next snapshot: its chunks land under pass 1
      fence_proxy_async();
      issue_g(b + gridDim.x);
    }
  }

  // ---------------- parameter gradients: CTA reduction, then one atomic add per column chunk
#pragma unroll
  for (int v = 0; v < V; ++v) {
    float4 accv[3] = {accs[v], accd[v], bacc[v]};
    const long long offs[3] = {off_att_src, off_att_dst, off_bias};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      __syncthreads();
      st4(red + tid * 4, accv[q]);
      __syncthreads();
      if (tid < LPR) {
        float4 s = f4zero();
        for (int w = 0; w < kWarpsT; ++w)
#pragma unroll
          for (int sb = 0; sb < 32 / LPR; ++sb) add4(s, *reinterpret_cast<float4*>(red + (w * 32 + sb * LPR + tid) * 4));
        atomicAdd(reinterpret_cast<float4*>(grads + offs[q] + 4 * (tid + v * LPR)), s);
      }
    }
  }
}

static bool bwd_tile_pipe() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GATRES_BWD_TILE_PIPE");      // 0 = load a snapshot, then run its passes (kernel above)
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

static int bwd_tile_pack() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GATRES_BWD_TILE_PACK");     // 0 = 1024 threads, one chunk per lane; 1 = packed, 512 threads;
    v = e ? atoi(e) : 2;                                // 2 = packed, 640 threads (5 passes of 80 rows for 388 nodes).
                                                        // measured at 2048 snapshots (two heads, nc = 32): 277 / 261 / 251 us
  }
  return v;
}

static bool bwd_tile_double_buffer() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GATRES_BWD_TILE_DB");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

template <int H, int C>
static int launch_bwd_tile(const int* rowptr, const int* col, const int* rowptr_t, const int* col_t, unsigned E1,
                           const float* g, const float* h, const float* s_src, const float* s_dst, const float* m,
                           const float* l, const float* att_src, const float* att_dst, float* dh, float* grads,
                           long long off_as, long long off_ad, long long off_b, unsigned B, unsigned N,
                           cudaStream_t st) {
  BwdTilePlan plan(N, H * C, H, E1);
  unsigned per_sm = (unsigned)((227u * 1024u) / (plan.total + 1024u));
  per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
  // one CTA per SM but room for a second input set: double-buffer the snapshot loads instead
  const BwdTilePlan plan2(N, H * C, H, E1, 2);
  const bool dbl = per_sm == 1 && plan2.total <= 227u * 1024u && bwd_tile_double_buffer();
  if (dbl) plan = plan2;
  unsigned grid = (unsigned)sm_count() * per_sm;
  if (grid > B) grid = B;
#define LAUNCH(THR, PK, NB)                                                                                       \
  do {                                                                                                            \
    auto kern = gat_agg_bwd_tile_kernel<H, C, THR, PK, NB>;                                                            \
    static uint32_t configured = 0;                                                                               \
    if (configured < plan.total) {                                                                                \
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total) != cudaSuccess) \
        return check_launch("gat_agg_bwd_tile: smem attribute");                                                  \
      configured = plan.total;                                                                                    \
    }                                                                                                             \
    launch_kernel(kern, dim3(grid), dim3(THR), plan.total, st, rowptr, col, rowptr_t, col_t, E1, g, h, s_src, s_dst, m, l, att_src,      \
                                        att_dst, dh, grads, off_as, off_ad, off_b, B, N);                         \
  } while (0)
  if (per_sm < 2 && !dbl && bwd_tile_pipe()) {
    // one input set per SM: the pipelined kernel (halves of the set under the other pass)
    const BwdPipePlan pp(N, H * C, H, E1);
    const bool pack = H == 2 && C == 32 && bwd_tile_pack() != 0;
    const unsigned rc = pack ? (640u / 32u) * 4u : 32u * (32u / (unsigned)(H * C / 4));      // rows per chunk of the variant below
    if (pp.total <= 227u * 1024u && (N + rc - 1) / rc <= (unsigned)kPipeMaxChunks && (rc * H * C * 4u) % 16u == 0) {
#define LAUNCH_PIPE(THR, PK)                                                                                         \
  do {                                                                                                               \
    auto kern = gat_agg_bwd_tile_pipe_kernel<H, C, THR, PK>;                                                         \
    static uint32_t configured = 0;                                                                                  \
    if (configured < pp.total) {                                                                                     \
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp.total) != cudaSuccess)     \
        return check_launch("gat_agg_bwd_tile_pipe: smem attribute");                                                \
      configured = pp.total;                                                                                         \
    }                                                                                                                \
    launch_kernel(kern, dim3(grid), dim3(THR), pp.total, st, rowptr, col, rowptr_t, col_t, E1, g, h, s_src, s_dst, m, l, \
                  att_src, att_dst, dh, grads, off_as, off_ad, off_b, B, N);                                         \
  } while (0)
      if (pack) LAUNCH_PIPE(640, true); else LAUNCH_PIPE(1024, false);
#undef LAUNCH_PIPE
      return check_launch("gat_agg_bwd_tile_pipe");
    }
  }
  if (per_sm >= 2) LAUNCH(512, false, 1);
  else if (dbl) LAUNCH(1024, false, 2);
  else if (H == 2 && C == 32 && bwd_tile_pack() == 2) LAUNCH(640, true, 1);
  else if (H == 2 && C == 32 && bwd_tile_pack()) LAUNCH(512, true, 1);
  else LAUNCH(1024, false, 1);
#undef LAUNCH
  return check_launch("gat_agg_bwd_tile");
}

bool bwd_tile_eligible(unsigned N, unsigned H, unsigned C, unsigned E1) {
  const BwdTilePlan plan(N, H * C, H, E1);
  return (N * H) % 4u == 0 && N < 65536u && plan.total <= 227u * 1024u && plan.slab >= 1024u * 16u && (H * C) <= 128;
}

int gat_agg_bwd_tile(const int* rowptr, const int* col, const int* rowptr_t, const int* col_t, unsigned E1,
                     const float* g, const float* h, const float* s_src, const float* s_dst, const float* m,
                     const float* l, const float* att_src, const float* att_dst, float* dh, float* grads,
                     long long off_as, long long off_ad, long long off_b, unsigned B, unsigned N, int H, int C,
                     cudaStream_t st) {
#define T(HH, CC)                                                                                                 \
  if (H == HH && C == CC)                                                                                         \
  return launch_bwd_tile<HH, CC>(rowptr, col, rowptr_t, col_t, E1, g, h, s_src, s_dst, m, l, att_src, att_dst, dh, \
                                 grads, off_as, off_ad, off_b, B, N, st)
  T(1, 32);
  T(2, 32);
  T(1, 64);
  T(2, 64);
  T(1, 128);
#undef T
  set_error("gat_agg_bwd_tile: unsupported (H=%d, C=%d)", H, C);
  return GATRES_ERR_ARG;
}

}  // namespace gatres

// =============================================================================
// SimpleConv(mean) backward for one snapshot per CTA iteration: dz[j] = sum over the out-edges j -> i (self-loop
// excluded) of g[i] / max(indeg(i), 1)  (mean_res.cu: mean_res_bwd_kernel with the ReLU mask already applied, the form
// the model's backward uses).  The gather version walks col_t -> rowptr -> g through L1 / L2 per edge (44 % of the
// copy bandwidth); here the snapshot's gradient slab is staged by TMA (double-buffered), the out-edge CSR and the
// per-edge weights 1 / max(indeg, 1) sit in shared memory, rows are handed out in ascending out-degree.
// =============================================================================
namespace gatres {

struct MeanBwdPlan {
  uint32_t slab, stage, rp_off, co_off, wt_off, ord_off, hist_off, bar_off, total;
  __host__ __device__ MeanBwdPlan(unsigned N, unsigned C, unsigned E1) {
    slab = N * C * 4u;
    stage = (slab + 127u) & ~127u;
    rp_off = 2u * stage;
    co_off = rp_off + (((N + 1u) * 4u + 15u) & ~15u);
    wt_off = co_off + ((E1 * 2u + 15u) & ~15u);
    ord_off = wt_off + ((E1 * 4u + 15u) & ~15u);
    hist_off = ord_off + ((N * 2u + 15u) & ~15u);
    bar_off = hist_off + 128u;
    total = bar_off + 16u;
  }
};

template <int C, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : 1)
mean_res_bwd_tile_kernel(const int* __restrict__ rowptr, const int* __restrict__ rowptr_t, const int* __restrict__ col_t,
                         unsigned E1, const float* __restrict__ g, float* __restrict__ dz, unsigned B, unsigned N) {
  constexpr int LPR = C / 4, RPW = 32 / LPR, kWarpsT = THREADS / 32;
  static_assert(LPR <= 32, "row wider than one warp pass");
  extern __shared__ __align__(128) unsigned char smem[];
  const MeanBwdPlan plan(N, C, E1);
  int* rpo = reinterpret_cast<int*>(smem + plan.rp_off);
  unsigned short* co = reinterpret_cast<unsigned short*>(smem + plan.co_off);
  float* wt = reinterpret_cast<float*>(smem + plan.wt_off);
  unsigned short* ord = reinterpret_cast<unsigned short*>(smem + plan.ord_off);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + plan.bar_off);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane / LPR, lig = lane % LPR;

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
  }
  for (unsigned k = tid; k <= N; k += THREADS) rpo[k] = __ldg(rowptr_t + k);
  for (unsigned k = tid; k < E1; k += THREADS) {
    const int t = __ldg(col_t + k);
    co[k] = (unsigned short)t;
    const int dg = __ldg(rowptr + t + 1) - __ldg(rowptr + t) - 1;        // in-degree of the target without its self-loop
    wt[k] = 1.f / (float)(dg > 1 ? dg : 1);
  }
  __syncthreads();
  degree_order<THREADS>(rpo, N, ord, reinterpret_cast<int*>(smem + plan.hist_off));
  pdl_wait();
  auto issue = [&](int stage, unsigned bb) {
    mbar_arrive_expect_tx(&full[stage], plan.slab);
    bulk_g2s(smem + (size_t)stage * plan.stage, g + (size_t)bb * N * C, plan.slab, &full[stage]);
  };
  if (tid == 0) {
    if (blockIdx.x < B) issue(0, blockIdx.x);
    if (blockIdx.x + gridDim.x < B) issue(1, blockIdx.x + gridDim.x);
  }
  unsigned k = 0;
  if (gridDim.x >= B) pdl_launch_dependents();
  for (unsigned b = blockIdx.x; b < B; b += gridDim.x, ++k) {
    const int stage = k & 1;
    const float* gs = reinterpret_cast<const float*>(smem + (size_t)stage * plan.stage) + 4 * lig;
    mbar_wait(&full[stage], (k >> 1) & 1);
    for (unsigned j0 = warp * RPW + sub; j0 < N; j0 += kWarpsT * RPW) {
      const unsigned j = ord[j0];
      const int beg = rpo[j], end = rpo[j + 1] - 1;                      // the self-loop is the last out-edge of a row
      float4 acc = f4zero();
#pragma unroll 4
      for (int e = beg; e < end; ++e) fma4(acc, wt[e], *reinterpret_cast<const float4*>(gs + co[e] * C));
      st4(dz + ((size_t)b * N + j) * C + 4 * lig, acc);
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

bool mean_bwd_tile_eligible(unsigned N, unsigned C, unsigned E1) {
  const MeanBwdPlan plan(N, C, E1);
  return (N * C) % 4u == 0 && N < 65536u && plan.total <= 227u * 1024u && (C == 32 || C == 64 || C == 128);
}

template <int C>
static int launch_mean_bwd_tile(const int* rowptr, const int* rowptr_t, const int* col_t, unsigned E1, const float* g,
                                float* dz, unsigned B, unsigned N, cudaStream_t st) {
  const MeanBwdPlan plan(N, C, E1);
  unsigned per_sm = (unsigned)((227u * 1024u) / (plan.total + 1024u));
  per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
  unsigned grid = (unsigned)sm_count() * per_sm;
  if (grid > B) grid = B;
#define LAUNCH(THR)                                                                                                \
  do {                                                                                                             \
    auto kern = mean_res_bwd_tile_kernel<C, THR>;                                                                  \
    static uint32_t configured = 0;                                                                                \
    if (configured < plan.total) {                                                                                 \
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total) != cudaSuccess) \
        return check_launch("mean_res_bwd_tile: smem attribute");                                                  \
      configured = plan.total;                                                                                     \
    }                                                                                                              \
    launch_kernel(kern, dim3(grid), dim3(THR), plan.total, st, rowptr, rowptr_t, col_t, E1, g, dz, B, N);          \
  } while (0)
  if (per_sm >= 2) LAUNCH(512); else LAUNCH(1024);
#undef LAUNCH
  return check_launch("mean_res_bwd_tile");
}

int mean_res_bwd_tile(const int* rowptr, const int* rowptr_t, const int* col_t, unsigned E1, const float* g, float* dz,
                      unsigned B, unsigned N, int C, cudaStream_t st) {
  if (C == 32) return launch_mean_bwd_tile<32>(rowptr, rowptr_t, col_t, E1, g, dz, B, N, st);
  if (C == 64) return launch_mean_bwd_tile<64>(rowptr, rowptr_t, col_t, E1, g, dz, B, N, st);
  if (C == 128) return launch_mean_bwd_tile<128>(rowptr, rowptr_t, col_t, E1, g, dz, B, N, st);
  set_error("mean_res_bwd_tile: unsupported channels %d", C);
  return GATRES_ERR_ARG;
}

}  // namespace gatres
