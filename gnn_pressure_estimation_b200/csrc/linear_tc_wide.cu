// Wide projections (nc = 64 / 128: K up to 256, N up to 256) on tcgen05, warp-specialised.
//
//   h = x W^T + attention-score epilogue     (GATConv.forward step 1 of the large GATRes,
//   /root/reference/gnn_pressure_estimation/GraphModels.py:464-465 with ConfigModels.py:33-42: 25 blocks x 128 channels)
//
// The first wide kernel (linear_tc.cu::gemm_tc_wide_kernel) let all 256 threads walk one serial chain per 32-column
// K chunk — wait for A, split W and A into 3xTF32 parts, barrier, issue, wait for the previous chunk's MMAs, refill —
// and drained the single accumulator with the tensor pipe idle: 5.3 k cycles per chunk against 1.5 k of MMA time
// (profiles/r1_configs.md).  Here every stage has its own warps and they only meet at mbarriers:
//
//   warps 0-3  A producers : the A chunk [128 rows x 32 columns] of step q + 2 is requested with 128-bit global loads into
//                            registers (two register sets alternate), the chunk of step q is written to shared memory as
//                            the K-major SWIZZLE_128B tile the tensor core reads (hi = the raw fp32 words, the MMA
//                            truncates them to TF32) together with its low part rna(x - trunc x); no raw staging ring,
//                            no cp.async groups: 32 KB of shared-memory writes per chunk instead of 48 KB + a read.
//   warp 9     W loader    : the weights are split ONCE per launch by wide_w_image_kernel into a pre-swizzled image
//                            [chunk][hi | lo][N x 32] in global memory (L2 resident, <= 256 KB); one lane streams it through a
//                            ring of stages with 1-D bulk copies (cp.async.bulk + mbarrier complete_tx): no thread ever
//                            touches a weight.
//   warp 8     MMA issuer  : waits for the A and W stages, one elected lane issues the 12 tcgen05.mma kind::tf32 of the
//                            chunk (4 K-steps x {hi hi, lo hi, hi lo}) and commits them to the "stage free" barriers; the
//                            last chunk of a tile also commits to "accumulator full".
//   warps 4-7  epilogue    : TWO accumulators in TMEM (2 x N columns: all 512 for N = 256), so the drain of tile t
//                            (tcgen05.ld, thread = row: the attention scores are in-thread dot products, rows leave
//                            through a 4 KB per-warp transposing buffer as full 128-byte lines) runs under the MMAs of
//                            tile t + 1.
//
// One CTA per SM (224 KB of shared memory), persistent over 128-row tiles.
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"
#include "tensormap.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace gatres {

// shared -> global 2-D tiled bulk store (TMA): the box described by `map` at element coordinates (x = column, y = row);
// rows beyond the tensor's extent are clipped by the engine
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x), "r"(y),
               "r"(smem_src)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int KK, int NN>
struct WideShape {
  static constexpr int BM = 128, KC = 32, NCHUNK = KK / KC;
  static constexpr uint32_t A_CH = BM * KC * 4;                 // one A part (hi or lo) of a chunk
  static constexpr uint32_t B_CH = NN * KC * 4;                 // one W part of a chunk
  static constexpr int SA = 2;                                  // A stages (hi + lo each)
  static constexpr int SB = NN == 256 ? 2 : 4;                  // W stages (hi + lo each): 128 KB either way
  static constexpr uint32_t EPI_WARP = 2 * 4096;               // two [32 rows x 128 B] transposing buffers per epilogue warp
  static constexpr uint32_t OFF_A = 0, OFF_B = OFF_A + SA * 2 * A_CH, OFF_EPI = OFF_B + SB * 2 * B_CH,
                            OFF_ATT = OFF_EPI + 4 * EPI_WARP, OFF_BAR = OFF_ATT + 2 * NN * 4;
  static constexpr int NBAR = 2 * SA + 2 * SB + 4;
  static constexpr uint32_t TOTAL = OFF_BAR + NBAR * 8 + 16;
  static_assert(KK % KC == 0 && (NN == 128 || NN == 256), "unsupported wide tensor-core shape");
  static_assert(TOTAL <= 232448, "shared memory budget");
};

// Weight image of one projection: image[ch][part][swz(n, k % 32)] with part 0 = W (its TF32 truncation is the hi part)
// and part 1 = rna(w - trunc w); every (chunk, part) block is a ready-to-read K-major SWIZZLE_128B B tile.
template <int KK, int NN>
__device__ __align__(1024) float g_wide_image[2 * KK * NN];

template <int KK, int NN>
__global__ void __launch_bounds__(256) wide_w_image_kernel(const float* __restrict__ W) {
  using S = WideShape<KK, NN>;
  pdl_wait();                                  // the previous projection of this shape may still be reading the image
  unsigned char* img = reinterpret_cast<unsigned char*>(g_wide_image<KK, NN>);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < NN * (KK / 4); idx += gridDim.x * blockDim.x) {
    const int n = idx / (KK / 4), k = 4 * (idx % (KK / 4));
    const float4 w = ldg4(W + (size_t)n * KK + k);
    float4 lo;
    lo.x = lo_tf32(w.x); lo.y = lo_tf32(w.y); lo.z = lo_tf32(w.z); lo.w = lo_tf32(w.w);
    const uint32_t off = (uint32_t)(k / S::KC) * 2u * S::B_CH + swz_off((uint32_t)n, (uint32_t)(k % S::KC), NN);
    *reinterpret_cast<float4*>(img + off) = w;
    *reinterpret_cast<float4*>(img + off + S::B_CH) = lo;
  }
}

// TS = rows leave through TMA tensor stores (out_map describes Cout as [M][NN] with 32 x 32 boxes, SWIZZLE_128B);
// otherwise the epilogue warps read their transposing buffer back and store 128-byte lines themselves.
template <int KK, int NN, int H, bool TS>
__global__ void __launch_bounds__(320, 1)
gemm_tc_wide2_kernel(const float* __restrict__ A, const float* __restrict__ att_src, const float* __restrict__ att_dst,
                     float* __restrict__ Cout, float* __restrict__ s0, float* __restrict__ s1, unsigned M,
                     const __grid_constant__ CUtensorMap out_map) {
  using S = WideShape<KK, NN>;
  constexpr int BM = S::BM, KC = S::KC, NCHUNK = S::NCHUNK, SA = S::SA, SB = S::SB;
  constexpr uint32_t A_CH = S::A_CH, B_CH = S::B_CH;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  static_assert(H == 1 || H == 2, "heads");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();
  unsigned char* sm = smem_raw;
  float* att = reinterpret_cast<float*>(sm + S::OFF_ATT);                       // [2][NN]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);              // [SA] 128 producer arrivals
  uint64_t* a_empty = a_full + SA;                                              // [SA] tcgen05.commit
  uint64_t* w_full = a_empty + SA;                                              // [SB] bulk-copy transaction bytes
  uint64_t* w_empty = w_full + SB;                                              // [SB] tcgen05.commit
  uint64_t* acc_full = w_empty + SB;                                            // [2]  tcgen05.commit
  uint64_t* acc_empty = acc_full + 2;                                           // [2]  128 epilogue arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned ntiles = (M + BM - 1) / BM;
  const unsigned my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const unsigned total = my_tiles * NCHUNK;                                     // flattened (tile, chunk) sequence

  if (warp == 0) tmem_alloc(tmem_slot, 2 * NN);
  if (tid == 32) {
    for (int i = 0; i < SA; ++i) { mbar_init(a_full + i, 128); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(w_full + i, 1); mbar_init(w_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 128); }
    mbar_fence_init();
  }
  for (int idx = tid; idx < NN; idx += blockDim.x) {
    att[idx] = __ldg(att_src + idx);
    att[NN + idx] = __ldg(att_dst + idx);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp < 4) {
    // ------------------------------------------------------------------ A producers
    // element idx = it * 128 + tid of a chunk: row idx / 8, 16-byte column idx % 8 (8 lanes cover one 128-byte row segment)
    float4 r0[8], r1[8];
    auto request = [&](unsigned q, float4 (&r)[8]) {
      const unsigned tile = blockIdx.x + (q / NCHUNK) * gridDim.x, ch = q % NCHUNK;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * 128 + tid;
        const unsigned grow = tile * BM + (unsigned)(idx >> 3);
        r[it] = grow < M ? ldg4_stream(A + (size_t)grow * KK + ch * KC + 4 * (idx & 7)) : f4zero();
      }
    };
    auto publish = [&](unsigned q, const float4 (&r)[8]) {
      const unsigned s = q % SA, n = q / SA;
      if (n > 0) mbar_wait(a_empty + s, (n - 1) & 1u);                          // MMAs of step q - SA have read the stage
      unsigned char* hi = sm + S::OFF_A + s * 2 * A_CH;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * 128 + tid;
        const uint32_t off = swz_off((uint32_t)(idx >> 3), 4u * (uint32_t)(idx & 7), BM);
        const float4 x = r[it];
        float4 lo;
        lo.x = lo_tf32(x.x); lo.y = lo_tf32(x.y); lo.z = lo_tf32(x.z); lo.w = lo_tf32(x.w);
        *reinterpret_cast<float4*>(hi + off) = x;
        *reinterpret_cast<float4*>(hi + A_CH + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(a_full + s);
    };
    if (total > 0) request(0, r0);
    if (total > 1) request(1, r1);
    for (unsigned q = 0; q < total; q += 2) {
      publish(q, r0);
      if (q + 2 < total) request(q + 2, r0);
      if (q + 1 < total) {
        publish(q + 1, r1);
        if (q + 3 < total) request(q + 3, r1);
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ epilogue: warp w drains TMEM lanes 32 (w % 4) ..
    const int quarter = warp & 3;
    unsigned char* stg0 = sm + S::OFF_EPI + quarter * S::EPI_WARP;               // 2 x [32 rows][128 B], 16-byte chunks XOR row % 8
    for (unsigned t = 0; t < my_tiles; ++t) {
      const unsigned tile = blockIdx.x + t * gridDim.x, acc = t & 1u;
      const unsigned row0 = tile * BM + quarter * 32;
      mbar_wait(acc_full + acc, (t >> 1) & 1u);
      tc_fence_after();
      float ps[H], pd[H];
#pragma unroll
      for (int h = 0; h < H; ++h) ps[h] = pd[h] = 0.f;
#pragma unroll
      for (int cb = 0; cb < NN / 32; ++cb) {
        const int col0 = cb * 32, h = (H == 2 && cb >= NN / 64) ? 1 : 0;
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + acc * NN + col0, v);
        if (cb == NN / 32 - 1) {                                                  // accumulator drained: the MMAs of tile t + 2 may start
          tc_fence_before();
          mbar_arrive(acc_empty + acc);
        }
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          a = fmaf(v[k], att[col0 + k], a);
          b = fmaf(v[k], att[NN + col0 + k], b);
        }
        if (H == 2) {
          if (h == 0) { ps[0] += a; pd[0] += b; } else { ps[H - 1] += a; pd[H - 1] += b; }
        } else {
          ps[0] += a; pd[0] += b;
        }
        unsigned char* stg = stg0 + (cb & 1) * 4096;
        if (TS) {                                                                 // the store issued two column blocks ago has read this buffer
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4*>(stg + lane * 128 + ((k ^ (lane & 7)) << 4)) =
              make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        if (TS) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && row0 < M) tma_store_2d(&out_map, smem_u32(stg), col0, (int)row0);
        } else {
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = j * 4 + (lane >> 3), c = lane & 7;
            const float4 o = *reinterpret_cast<const float4*>(stg + r * 128 + ((c ^ (r & 7)) << 4));
            if (row0 + r < M) st4(Cout + (size_t)(row0 + r) * NN + col0 + 4 * c, o);
          }
          __syncwarp();
        }
      }
      if (row0 + lane < M) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          s0[(size_t)(row0 + lane) * H + h] = ps[h];
          s1[(size_t)(row0 + lane) * H + h] = pd[h];
        }
      }
    }
    if (TS && lane == 0) tma_store_wait_all();
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    for (unsigned q = 0; q < total; ++q) {
      const unsigned t = q / NCHUNK, ch = q % NCHUNK, acc = t & 1u;
      const unsigned sa = q % SA, sb = q % SB;
      if (ch == 0 && t >= 2) mbar_wait(acc_empty + acc, ((t >> 1) - 1) & 1u);    // epilogue of tile t - 2 has drained it
      mbar_wait(w_full + sb, (q / SB) & 1u);
      mbar_wait(a_full + sa, (q / SA) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_hi = base + S::OFF_A + sa * 2 * A_CH, a_lo = a_hi + A_CH;
        const uint32_t b_hi = base + S::OFF_B + sb * 2 * B_CH, b_lo = b_hi + B_CH;
        const uint32_t d = tmem + acc * NN;
#pragma unroll
        for (int ks = 0; ks < KC / 8; ++ks) {
          const uint32_t ko = (uint32_t)ks * 32u;
          umma_tf32(d, umma_desc_k128(a_hi + ko), umma_desc_k128(b_hi + ko), IDESC, (ch > 0 || ks > 0) ? 1u : 0u);
          umma_tf32(d, umma_desc_k128(a_lo + ko), umma_desc_k128(b_hi + ko), IDESC, 1);
          umma_tf32(d, umma_desc_k128(a_hi + ko), umma_desc_k128(b_lo + ko), IDESC, 1);
        }
        umma_commit(a_empty + sa);
        umma_commit(w_empty + sb);
        if (ch == NCHUNK - 1) umma_commit(acc_full + acc);
      }
      __syncwarp();
    }
  } else if (lane == 0) {
    // ------------------------------------------------------------------ W loader (one lane of warp 9)
    const unsigned char* img = reinterpret_cast<const unsigned char*>(g_wide_image<KK, NN>);
    for (unsigned q = 0; q < total; ++q) {
      const unsigned sb = q % SB, n = q / SB, ch = q % NCHUNK;
      if (n > 0) mbar_wait(w_empty + sb, (n - 1) & 1u);
      mbar_arrive_expect_tx(w_full + sb, 2 * B_CH);
      bulk_g2s(sm + S::OFF_B + sb * 2 * B_CH, img + (size_t)ch * 2 * B_CH, 2 * B_CH, w_full + sb);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem, 2 * NN);
}

template <int KK, int NN, int H>
static int launch_tc_wide2(const float* A, const float* W, const float* e0, const float* e1, float* Cout, float* s0,
                           float* s1, unsigned M, cudaStream_t st) {
  using S = WideShape<KK, NN>;
  static int tma_store = -1;
  if (tma_store < 0) {
    const char* e = getenv("GATRES_TC_WIDE2_TMA_STORE");
    tma_store = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  CUtensorMap map;
  const bool ts = tma_store == 1 && make_map_2d(&map, Cout, M, NN, 32, 32, true);
  if (!ts) memset(&map, 0, sizeof(map));
  auto kern = ts ? gemm_tc_wide2_kernel<KK, NN, H, true> : gemm_tc_wide2_kernel<KK, NN, H, false>;
  static bool configured[2] = {false, false};
  if (!configured[ts]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL) != cudaSuccess)
      return check_launch("gemm_tc_wide2: smem attribute");
    configured[ts] = true;
  }
  const unsigned ntiles = (M + 127) / 128;
  unsigned grid = (unsigned)sm_count();
  if (grid > ntiles) grid = ntiles;
  launch_kernel(wide_w_image_kernel<KK, NN>, dim3(NN * (KK / 4) / 256), dim3(256), (size_t)0, st, W);
  launch_kernel(kern, dim3(grid), dim3(320), (size_t)S::TOTAL, st, A, e0, e1, Cout, s0, s1, M, map);
  return check_launch("gemm_tc_wide2");
}

// -> 1 if handled, 0 if the shape is not covered or the kernel is switched off (GATRES_TC_WIDE2=0), < 0 on error
int gemm_tc_wide2_dispatch(int H, int KK, int NN, const float* A, const float* W, const float* e0, const float* e1,
                           float* Cout, float* s0, float* s1, unsigned M, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("GATRES_TC_WIDE2");
    enabled = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  if (!enabled) return 0;
  int rc;
  if (KK == 128 && NN == 256 && H == 2) rc = launch_tc_wide2<128, 256, 2>(A, W, e0, e1, Cout, s0, s1, M, st);
  else if (KK == 256 && NN == 128 && H == 1) rc = launch_tc_wide2<256, 128, 1>(A, W, e0, e1, Cout, s0, s1, M, st);
  else if (KK == 64 && NN == 128 && H == 2) rc = launch_tc_wide2<64, 128, 2>(A, W, e0, e1, Cout, s0, s1, M, st);
  else return 0;
  return rc == GATRES_OK ? 1 : rc;
}

}  // namespace gatres
