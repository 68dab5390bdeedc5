// Projections on tcgen05, warp-specialised: wide shapes (nc = 64 / 128: K up to 256, N up to 256) and, for large launches, nc = 32.
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
//                            (tcgen05.ld, thread = row: the attention scores are in-thread dot products) runs under the
//                            MMAs of tile t + 1; rows leave as 32 x 32 boxes through TMA tensor stores
//                            (cp.async.bulk.tensor.2d from a SWIZZLE_128B staging buffer, two buffers per warp) — the
//                            register-store epilogue (read the buffer back, st.global) stays as the fallback when the
//                            driver entry point for cuTensorMapEncodeTiled is missing.
//
// For the nc = 32 shapes (hi + lo of the whole W = 16 KB) W is loaded once per CTA and stays resident, the ring's shared memory
// goes to five A stages and there are eight producer warps (the other roles move up by four warps); the same file holds the
// weight gradient and the fused projection backward of those layers on tcgen05 (MN-major operands, below).
//
// One CTA per SM (226 KB of shared memory), persistent over 128-row tiles; the next tile's A block (contiguous) is
// prefetched into L2 with one cp.async.bulk.prefetch.  Measured (B200, 794 624 rows, profiles/r2_wide_and_sliced.md):
// K = 128 -> N = 256: 473 -> 280 us (40 -> 67 % of the copy bandwidth, tensor pipe 27 -> 70 %), K = 256 -> N = 128: 480 -> 316 us,
// K = 64 -> N = 128: 177 -> 111 us (86 %).  Bound by the traffic of the shared-memory data array (SS-mode operand reads + W
// bulk copies + staging) and, in sustained runs, by the board power cap.  A CTA-pair form (cta_group::2, half the W bytes
// and 8 instead of 12 KB of operand reads per CTA and instruction) is kept opt-in below: parity-green, same speed.
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

// ask the L2 for a contiguous block ahead of its use (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void l2_prefetch_bulk(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
// The A tile of a persistent CTA's NEXT tile (128 rows x K floats, contiguous) is prefetched into L2 as one sequential
// block when the current tile starts: the chunk loads read a 128-byte slice of each of 128 rows (row stride K floats), and
// issued straight at DRAM they open a page per 128 bytes (K = 256: 56 % of the copy bandwidth whatever the pipeline
// around them looked like — single CTA, CTA pair, four or eight producer warps).
__device__ __forceinline__ void prefetch_a_tile(const float* A, unsigned tile, unsigned M, unsigned KK) {
  const unsigned row0 = tile * 128u;
  if (row0 >= M) return;
  const unsigned rows = M - row0 < 128u ? M - row0 : 128u;
  l2_prefetch_bulk(A + (size_t)row0 * KK, rows * KK * 4u);
}

#ifndef GATRES_WIDE_L2PF
#define GATRES_WIDE_L2PF 1
#endif
constexpr bool L2PF = GATRES_WIDE_L2PF != 0;

template <int KK, int NN>
struct WideShape {
  static constexpr int BM = 128, KC = 32, NCHUNK = KK / KC;
  static constexpr uint32_t A_CH = BM * KC * 4;                 // one A part (hi or lo) of a chunk
  static constexpr uint32_t B_CH = NN * KC * 4;                 // one W part of a chunk
  // RESIDENT (nc = 32 shapes: hi + lo of the whole W <= 32 KB): every chunk of W is loaded once per CTA and stays; the
  // shared memory that the W ring would take goes to A stages, and eight producer warps keep up with tiles whose MMAs
  // take ~400 cycles
  static constexpr bool RESIDENT = 2 * KK * NN * 4 <= 32768;
  static constexpr int PW = RESIDENT ? 8 : 4;                   // producer warps; + 4 epilogue warps + MMA issuer + W loader
  static constexpr int THREADS = (PW + 6) * 32;
  static constexpr int SA = RESIDENT ? 5 : 2;                   // A stages (hi + lo each)
  static constexpr int SB = RESIDENT ? NCHUNK : (NN == 256 ? 2 : 4);   // W stages (hi + lo each): 128 KB in the streamed form
  static constexpr uint32_t EPI_WARP = 2 * 4096;               // two [32 rows x 128 B] transposing buffers per epilogue warp
  static constexpr uint32_t OFF_A = 0, OFF_B = OFF_A + SA * 2 * A_CH, OFF_EPI = OFF_B + SB * 2 * B_CH,
                            OFF_ATT = OFF_EPI + 4 * EPI_WARP, OFF_BAR = OFF_ATT + 2 * NN * 4;
  static constexpr int NBAR = 2 * SA + 2 * SB + 4;
  static constexpr uint32_t TOTAL = OFF_BAR + NBAR * 8 + 16;
  static_assert(KK % KC == 0 && (NN == 32 || NN == 64 || NN == 128 || NN == 256), "unsupported tensor-core shape");
  static_assert(TOTAL <= 232448, "shared memory budget");
};

// Weight image of one projection: image[ch][part][swz(n, k % 32)] with part 0 = W (its TF32 truncation is the hi part)
// and part 1 = rna(w - trunc w); every (chunk, part) block is a ready-to-read K-major SWIZZLE_128B B tile.
// One image per shape and process, rewritten before every launch in stream order (the writer waits for the kernels
// ahead of it in the stream): projections of ONE shape must not run concurrently on two streams of one process — the
// library's callers (model stack, TrainStep, evaluation) enqueue them on a single stream.
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
__global__ void __launch_bounds__((WideShape<KK, NN>::THREADS), 1)
gemm_tc_wide2_kernel(const float* __restrict__ A, const float* __restrict__ att_src, const float* __restrict__ att_dst,
                     float* __restrict__ Cout, float* __restrict__ s0, float* __restrict__ s1, unsigned M,
                     const __grid_constant__ CUtensorMap out_map) {
  using S = WideShape<KK, NN>;
  constexpr int BM = S::BM, KC = S::KC, NCHUNK = S::NCHUNK, SA = S::SA, SB = S::SB, PW = S::PW;
  constexpr bool RESIDENT = S::RESIDENT;
  constexpr int PT = PW * 32, PI = 1024 / PT;                                   // producer threads, float4 per thread and chunk
  constexpr uint32_t A_CH = S::A_CH, B_CH = S::B_CH;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  static_assert(H == 1 || H == 2, "heads");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();
  unsigned char* sm = smem_raw;
  float* att = reinterpret_cast<float*>(sm + S::OFF_ATT);                       // [2][NN]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);              // [SA] producer-thread arrivals
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
    for (int i = 0; i < SA; ++i) { mbar_init(a_full + i, PT); mbar_init(a_empty + i, 1); }
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

  if (warp < PW) {
    // ------------------------------------------------------------------ A producers
    // element idx = it * PT + tid of a chunk: row idx / 8, 16-byte column idx % 8 (8 lanes cover one 128-byte row segment)
    float4 r0[PI], r1[PI];
    auto request = [&](unsigned q, float4 (&r)[PI]) {
      const unsigned tile = blockIdx.x + (q / NCHUNK) * gridDim.x, ch = q % NCHUNK;
      if (L2PF && ch == 0 && tid == 0) prefetch_a_tile(A, tile + gridDim.x, M, KK);
#pragma unroll
      for (int it = 0; it < PI; ++it) {
        const int idx = it * PT + tid;
        const unsigned grow = tile * BM + (unsigned)(idx >> 3);
        r[it] = grow < M ? ldg4_stream(A + (size_t)grow * KK + ch * KC + 4 * (idx & 7)) : f4zero();
      }
    };
    auto publish = [&](unsigned q, const float4 (&r)[PI]) {
      const unsigned s = q % SA, n = q / SA;
      if (n > 0) mbar_wait(a_empty + s, (n - 1) & 1u);                          // MMAs of step q - SA have read the stage
      unsigned char* hi = sm + S::OFF_A + s * 2 * A_CH;
#pragma unroll
      for (int it = 0; it < PI; ++it) {
        const int idx = it * PT + tid;
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
  } else if (warp < PW + 4) {
    // ------------------------------------------------------------------ epilogue: warp w drains TMEM lanes 32 (w % 4) ..
    const int quarter = warp & 3;
    unsigned char* stg0 = sm + S::OFF_EPI + quarter * S::EPI_WARP;               // 2 x [32 rows][128 B], 16-byte chunks XOR row % 8
    unsigned nstage = 0;
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
        unsigned char* stg = stg0 + (nstage++ & 1u) * 4096;          // alternate per store (N = 32: one store per tile)
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
  } else if (warp == PW + 4) {
    // ------------------------------------------------------------------ MMA issuer
    for (unsigned q = 0; q < total; ++q) {
      const unsigned t = q / NCHUNK, ch = q % NCHUNK, acc = t & 1u;
      const unsigned sa = q % SA, sb = RESIDENT ? ch : q % SB;
      if (ch == 0 && t >= 2) mbar_wait(acc_empty + acc, ((t >> 1) - 1) & 1u);    // epilogue of tile t - 2 has drained it
      mbar_wait(w_full + sb, RESIDENT ? 0u : ((q / SB) & 1u));                   // resident chunks complete once
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
        if (!RESIDENT) umma_commit(w_empty + sb);
        if (ch == NCHUNK - 1) umma_commit(acc_full + acc);
      }
      __syncwarp();
    }
  } else if (lane == 0) {
    // ------------------------------------------------------------------ W loader (one lane of the last warp)
    const unsigned char* img = reinterpret_cast<const unsigned char*>(g_wide_image<KK, NN>);
    for (unsigned q = 0; q < (RESIDENT ? (total > 0 ? (unsigned)NCHUNK : 0u) : total); ++q) {
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

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair form (tcgen05 cta_group::2): two CTAs of a cluster (one TPC) share every MMA.  CTA r owns the 128-row tile
// 2 p + r of pair p and HALF of the weight chunk (rows r N/2 .. of W); one instruction M = 256 x N x 8 issued by CTA 0 reads
// A from both CTAs' shared memory and each half of B once per pair: 8 KB of operand reads per CTA and MMA instead of 12
// (N = 256) and half the bulk-copy writes of W per CTA — the single-CTA form is bound by exactly that traffic (ncu: LSU +
// tensor-core + TMA wavefronts of the shared-memory data array ~ 90 % of the tile time, tensor pipe 54 %).
//   * barriers live at the same offsets in both CTAs; "full" barriers that the issuing CTA waits on count the arrivals of
//     both CTAs (the peer's producer / epilogue threads arrive remotely through mapa), the peer's bulk copies complete on
//     its own barrier and one relay lane forwards that completion; tcgen05.commit multicasts "stage free" / "accumulator
//     full" to both CTAs.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned wide_cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned wide_cluster_id() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned wide_nclusters() { unsigned r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void wide_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, unsigned rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {      // acquire at cluster scope
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int KK, int NN>
struct PairShape {
  static constexpr int BM = 128, KC = 32, NCHUNK = KK / KC, NH = NN / 2;
  static constexpr uint32_t A_CH = BM * KC * 4;                 // one A part (hi or lo) of a chunk
  static constexpr uint32_t B_CH = NH * KC * 4;                 // one part of this CTA's half of a W chunk
  static constexpr int SA = NN == 256 ? 3 : 4;
  static constexpr int SB = NN == 256 ? 3 : 4;
  static constexpr uint32_t EPI_WARP = 2 * 4096;
  static constexpr uint32_t OFF_A = 0, OFF_B = OFF_A + SA * 2 * A_CH, OFF_EPI = OFF_B + SB * 2 * B_CH,
                            OFF_ATT = OFF_EPI + 4 * EPI_WARP, OFF_BAR = OFF_ATT + 2 * NN * 4;
  static constexpr int NBAR = 3 * SA + 3 * SB + 6;
  static constexpr uint32_t TOTAL = OFF_BAR + NBAR * 8 + 16;
  static_assert(KK % KC == 0 && (NN == 128 || NN == 256), "unsupported wide tensor-core shape");
  static_assert(TOTAL <= 232448, "shared memory budget");
};

// image[ch][rank][part][swz(n - rank N/2, k % 32)]: every (chunk, rank) block = the hi and lo tiles of one CTA's half
template <int KK, int NN>
__device__ __align__(1024) float g_pair_image[2 * KK * NN];

template <int KK, int NN>
__global__ void __launch_bounds__(256) pair_w_image_kernel(const float* __restrict__ W) {
  using S = PairShape<KK, NN>;
  pdl_wait();
  unsigned char* img = reinterpret_cast<unsigned char*>(g_pair_image<KK, NN>);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < NN * (KK / 4); idx += gridDim.x * blockDim.x) {
    const int n = idx / (KK / 4), k = 4 * (idx % (KK / 4));
    const float4 w = ldg4(W + (size_t)n * KK + k);
    float4 lo;
    lo.x = lo_tf32(w.x); lo.y = lo_tf32(w.y); lo.z = lo_tf32(w.z); lo.w = lo_tf32(w.w);
    const uint32_t off = ((uint32_t)(k / S::KC) * 2u + (uint32_t)(n / S::NH)) * 2u * S::B_CH +
                         swz_off((uint32_t)(n % S::NH), (uint32_t)(k % S::KC), S::NH);
    *reinterpret_cast<float4*>(img + off) = w;
    *reinterpret_cast<float4*>(img + off + S::B_CH) = lo;
  }
}

#ifndef GATRES_WIDE_L2PF
#define GATRES_WIDE_L2PF 1
#endif
constexpr int PAIR_PRODUCER_WARPS = 8;                      // + 4 epilogue warps, MMA issuer / W relay, W loader, A relay
constexpr int PAIR_THREADS = (PAIR_PRODUCER_WARPS + 7) * 32;

template <int KK, int NN, int H>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
gemm_tc_pair_kernel(const float* __restrict__ A, const float* __restrict__ att_src, const float* __restrict__ att_dst,
                    float* __restrict__ s0, float* __restrict__ s1, unsigned M, const __grid_constant__ CUtensorMap out_map) {
  using S = PairShape<KK, NN>;
  constexpr int BM = S::BM, KC = S::KC, NCHUNK = S::NCHUNK, SA = S::SA, SB = S::SB;
  constexpr uint32_t A_CH = S::A_CH, B_CH = S::B_CH;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
  static_assert(H == 1 || H == 2, "heads");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();
  unsigned char* sm = smem_raw;
  float* att = reinterpret_cast<float*>(sm + S::OFF_ATT);                       // [2][NN]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);              // [SA] own producer warps
  uint64_t* a_empty = a_full + SA;                                              // [SA] tcgen05.commit (multicast)
  uint64_t* w_full = a_empty + SA;                                              // [SB] own bulk copies
  uint64_t* w_peer = w_full + SB;                                               // [SB] CTA 0: the peer's half has landed (relay)
  uint64_t* w_empty = w_peer + SB;                                              // [SB] tcgen05.commit (multicast)
  uint64_t* acc_full = w_empty + SB;                                            // [2]  tcgen05.commit (multicast)
  uint64_t* acc_empty = acc_full + 2;                                           // [2]  own epilogue warps
  uint64_t* a_peer = acc_empty + 2;                                             // [SA] CTA 0: the peer's A stage is published (relay)
  uint64_t* acc_peer = a_peer + SA;                                             // [2]  CTA 0: the peer's accumulator is drained (relay)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_peer + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned rank = wide_cluster_rank();
  const unsigned ntiles = (M + BM - 1) / BM, npairs = (ntiles + 1) / 2;
  const unsigned cid = wide_cluster_id(), ncl = wide_nclusters();
  const unsigned my_pairs = cid < npairs ? (npairs - cid + ncl - 1) / ncl : 0;
  const unsigned total = my_pairs * NCHUNK;                                     // flattened (pair, chunk) sequence

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * NN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int i = 0; i < SA; ++i) { mbar_init(a_full + i, PAIR_PRODUCER_WARPS); mbar_init(a_empty + i, 1); mbar_init(a_peer + i, 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(w_full + i, 1); mbar_init(w_peer + i, 1); mbar_init(w_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); mbar_init(acc_peer + i, 1); }
    mbar_fence_init();
  }
  for (int idx = tid; idx < NN; idx += blockDim.x) {
    att[idx] = __ldg(att_src + idx);
    att[NN + idx] = __ldg(att_dst + idx);
  }
  tc_fence_before();
  __syncthreads();
  wide_cluster_sync();                                                          // both CTAs' barriers exist before any remote arrival
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp < PAIR_PRODUCER_WARPS) {
    // ------------------------------------------------------------------ A producers (as in the single-CTA form, eight warps)
    constexpr int PT = PAIR_PRODUCER_WARPS * 32, PI = 1024 / PT;          // producer threads, float4 per thread and chunk
    float4 r0[PI], r1[PI];
    auto request = [&](unsigned q, float4 (&r)[PI]) {
      const unsigned tile = 2 * (cid + (q / NCHUNK) * ncl) + rank, ch = q % NCHUNK;
      if (L2PF && ch == 0 && tid == 0) prefetch_a_tile(A, tile + 2 * ncl, M, KK);
#pragma unroll
      for (int it = 0; it < PI; ++it) {
        const int idx = it * PT + tid;
        const unsigned grow = tile * BM + (unsigned)(idx >> 3);
        r[it] = grow < M ? ldg4_stream(A + (size_t)grow * KK + ch * KC + 4 * (idx & 7)) : f4zero();
      }
    };
#ifdef GATRES_TC_PROF
    long long qt[3] = {0, 0, 0}, ql = clock64();
#define QQ(i) do { const long long now_ = clock64(); qt[i] += now_ - ql; ql = now_; } while (0)
#else
#define QQ(i)
#endif
    auto publish = [&](unsigned q, const float4 (&r)[PI]) {
      const unsigned s = q % SA, n = q / SA;
      QQ(2);
      if (n > 0) mbar_wait(a_empty + s, (n - 1) & 1u);
      QQ(0);
      unsigned char* hi = sm + S::OFF_A + s * 2 * A_CH;
#pragma unroll
      for (int it = 0; it < PI; ++it) {
        const int idx = it * PT + tid;
        const uint32_t off = swz_off((uint32_t)(idx >> 3), 4u * (uint32_t)(idx & 7), BM);
        const float4 x = r[it];
        float4 lo;
        lo.x = lo_tf32(x.x); lo.y = lo_tf32(x.y); lo.z = lo_tf32(x.z); lo.w = lo_tf32(x.w);
        *reinterpret_cast<float4*>(hi + off) = x;
        *reinterpret_cast<float4*>(hi + A_CH + off) = lo;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + s);                                   // one (local) arrival per producer warp
      QQ(1);
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
#ifdef GATRES_TC_PROF
    if (cid == 0 && tid == 0)
      printf("tc_pair<%d,%d> rank %u producer: wait_empty %lld publish %lld request %lld (cycles per chunk)\n", KK, NN, rank,
             qt[0] / total, qt[1] / total, qt[2] / total);
#endif
  } else if (warp < PAIR_PRODUCER_WARPS + 4) {
    // ------------------------------------------------------------------ epilogue
    const int quarter = warp & 3;
    unsigned char* stg0 = sm + S::OFF_EPI + quarter * S::EPI_WARP;
    unsigned nstage = 0;
    for (unsigned t = 0; t < my_pairs; ++t) {
      const unsigned tile = 2 * (cid + t * ncl) + rank, acc = t & 1u;
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
        if (cb == NN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty + acc);                          // one (local) arrival per epilogue warp
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
        unsigned char* stg = stg0 + (nstage++ & 1u) * 4096;          // alternate per store (N = 32: one store per tile)
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4*>(stg + lane * 128 + ((k ^ (lane & 7)) << 4)) =
              make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && row0 < M) tma_store_2d(&out_map, smem_u32(stg), col0, (int)row0);
      }
      if (row0 + lane < M) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          s0[(size_t)(row0 + lane) * H + h] = ps[h];
          s1[(size_t)(row0 + lane) * H + h] = pd[h];
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  } else if (warp == PAIR_PRODUCER_WARPS + 4) {
    if (rank == 0) {
      // ---------------------------------------------------------------- MMA issuer (CTA 0 of the pair)
#ifdef GATRES_TC_PROF
      long long pt[5] = {0, 0, 0, 0, 0}, pl = clock64();
#define PP(i) do { const long long now_ = clock64(); pt[i] += now_ - pl; pl = now_; } while (0)
#else
#define PP(i)
#endif
      for (unsigned q = 0; q < total; ++q) {
        const unsigned t = q / NCHUNK, ch = q % NCHUNK, acc = t & 1u;
        const unsigned sa = q % SA, sb = q % SB;
        if (ch == 0 && t >= 2) {
          mbar_wait(acc_empty + acc, ((t >> 1) - 1) & 1u);
          mbar_wait_cluster(acc_peer + acc, ((t >> 1) - 1) & 1u);
        }
        PP(0);
        mbar_wait(w_full + sb, (q / SB) & 1u);
        PP(1);
        mbar_wait_cluster(w_peer + sb, (q / SB) & 1u);
        PP(2);
        mbar_wait(a_full + sa, (q / SA) & 1u);
        mbar_wait_cluster(a_peer + sa, (q / SA) & 1u);
        PP(3);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_hi = base + S::OFF_A + sa * 2 * A_CH, a_lo = a_hi + A_CH;
          const uint32_t b_hi = base + S::OFF_B + sb * 2 * B_CH, b_lo = b_hi + B_CH;
          const uint32_t d = tmem + acc * NN;
#pragma unroll
          for (int ks = 0; ks < KC / 8; ++ks) {
            const uint32_t ko = (uint32_t)ks * 32u;
            umma_tf32_2cta(d, umma_desc_k128(a_hi + ko), umma_desc_k128(b_hi + ko), IDESC, (ch > 0 || ks > 0) ? 1u : 0u);
            umma_tf32_2cta(d, umma_desc_k128(a_lo + ko), umma_desc_k128(b_hi + ko), IDESC, 1);
            umma_tf32_2cta(d, umma_desc_k128(a_hi + ko), umma_desc_k128(b_lo + ko), IDESC, 1);
          }
          umma_commit_2cta(a_empty + sa);
          umma_commit_2cta(w_empty + sb);
          if (ch == NCHUNK - 1) umma_commit_2cta(acc_full + acc);
        }
        __syncwarp();
        PP(4);
      }
#ifdef GATRES_TC_PROF
      if (cid == 0 && lane == 0)
        printf("tc_pair<%d,%d> chunks %u: acc_empty %lld w_full %lld w_peer %lld a_full %lld issue %lld (cycles per chunk)\n", KK, NN,
               total, pt[0] / total, pt[1] / total, pt[2] / total, pt[3] / total, pt[4] / total);
#endif
    } else if (lane == 0) {
      // ---------------------------------------------------------------- relay (CTA 1): own half of W landed -> tell CTA 0
      for (unsigned q = 0; q < total; ++q) {
        const unsigned sb = q % SB;
        mbar_wait(w_full + sb, (q / SB) & 1u);
        mbar_arrive_cluster(w_peer + sb, 0);
      }
    }
  } else if (warp == PAIR_PRODUCER_WARPS + 6) {
    // ------------------------------------------------------------------ relay (CTA 1): published A stages and drained
    // accumulators are forwarded to CTA 0 by a lane that has no memory operations of its own in flight: the cluster-scope
    // release of a remote arrival waits for the issuing thread's outstanding loads, which stalled the producers on their
    // own prefetches when they arrived remotely themselves (2.0 k of 2.8 k cycles per chunk)
    if (rank == 1 && lane == 0) {
      for (unsigned q = 0; q < total; ++q) {
        const unsigned t = q / NCHUNK, ch = q % NCHUNK, acc = t & 1u, sa = q % SA;
        if (ch == 0 && t >= 2) {
          mbar_wait(acc_empty + acc, ((t >> 1) - 1) & 1u);
          mbar_arrive_cluster(acc_peer + acc, 0);
        }
        mbar_wait(a_full + sa, (q / SA) & 1u);
        mbar_arrive_cluster(a_peer + sa, 0);
      }
    }
  } else if (lane == 0) {
    // ------------------------------------------------------------------ W loader: this CTA's half of every chunk
    const unsigned char* img = reinterpret_cast<const unsigned char*>(g_pair_image<KK, NN>);
    for (unsigned q = 0; q < total; ++q) {
      const unsigned sb = q % SB, n = q / SB, ch = q % NCHUNK;
      if (n > 0) mbar_wait(w_empty + sb, (n - 1) & 1u);
      mbar_arrive_expect_tx(w_full + sb, 2 * B_CH);
      bulk_g2s(sm + S::OFF_B + sb * 2 * B_CH, img + ((size_t)ch * 2 + rank) * 2 * B_CH, 2 * B_CH, w_full + sb);
    }
  }
  tc_fence_before();
  __syncthreads();
  wide_cluster_sync();                                                          // the peer's MMAs / remote arrivals are done with this CTA
  tc_fence_after();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * NN) : "memory");
}

template <int KK, int NN, int H>
static int launch_tc_pair(const float* A, const float* W, const float* e0, const float* e1, float* Cout, float* s0,
                          float* s1, unsigned M, cudaStream_t st) {
  using S = PairShape<KK, NN>;
  CUtensorMap map;
  if (!make_map_2d(&map, Cout, M, NN, 32, 32, true)) return 0;                  // no driver entry point: single-CTA form
  auto kern = gemm_tc_pair_kernel<KK, NN, H>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL) != cudaSuccess)
      return check_launch("gemm_tc_pair: smem attribute");
    configured = true;
  }
  const unsigned npairs = ((M + 127) / 128 + 1) / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(PAIR_THREADS);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static int resident_pairs = -1;                 // CTA pairs the device runs at once (the work split is static)
  if (resident_pairs < 0) {
    cfg.gridDim = dim3((unsigned)sm_count());
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = sm_count() / 2;
    }
    resident_pairs = n;
  }
  unsigned ncl = (unsigned)resident_pairs;
  if (ncl > npairs) ncl = npairs;
  cfg.gridDim = dim3(2 * ncl);
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  launch_kernel(pair_w_image_kernel<KK, NN>, dim3(NN * (KK / 4) / 256), dim3(256), (size_t)0, st, W);
  count_launch();
  cudaLaunchKernelEx(&cfg, kern, A, e0, e1, s0, s1, M, map);
  const int rc = check_launch("gemm_tc_pair");
  return rc == GATRES_OK ? 1 : rc;
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
  launch_kernel(wide_w_image_kernel<KK, NN>, dim3((NN * (KK / 4) + 255) / 256), dim3(256), (size_t)0, st, W);
  launch_kernel(kern, dim3(grid), dim3(S::THREADS), (size_t)S::TOTAL, st, A, e0, e1, Cout, s0, s1, M, map);
  return check_launch("gemm_tc_wide2");
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight gradient dW = dh^T x of the nc = 32 layers on tcgen05 (atomic accumulation mode, large launches).
//
// The reduction runs over ROWS, the dimension both operands are strided in, so both are MN-major for the tensor core:
// a tile is stored as the rows come from memory — [128 rows][32 floats = 128 B] per 32-column group, the 32-byte chunk of a
// row XORed with row % 4 (descriptor layout SWIZZLE_128B_BASE32B, the only layout tcgen05 reads an MN-major tf32 operand
// from; LBO = distance between 32-column groups, SBO = 512 B = 4 rows) — and no transpose is needed.  P = the 64-column
// operand is A (M = 64), Q = the 32-column operand is B (N = 32): D[64][32] = P^T Q accumulates in 32 TMEM columns over ALL
// row tiles of a persistent CTA (3xTF32: hi hi + lo hi + hi lo, 48 MMAs per 128 rows).  The tensor core's fp32 accumulation
// truncates: one accumulator carried through the ~2000 MMAs of a CTA ended 3.4e-5 off the fp64 product (mma.sync with register
// accumulators: 2.3e-6), so the accumulator is drained every two tiles (96 MMAs, the chain length of the K = 256 forward
// projection) by four epilogue warps that keep the running sum in registers (round-to-nearest adds); two accumulators
// alternate so that the drain runs under the next group's MMAs.  One flush with atomics at the end.
//   conv1 (dh1 [M, 64], x [M, 32]):  P = dh, Q = x,  D = dW1 [64][32]
//   conv2 (dh2 [M, 32], y1 [M, 64]): P = y1, Q = dh, D = dW2^T
// Eight producer warps (register prefetch two chunks ahead, hi = raw words, lo = rna(x - trunc x)), four epilogue warps, one MMA-issuer warp; the
// legacy form (wgrad_mma_kernel, mma.sync m16n8k8) issues 768 instructions per 128 rows at ~20 cycles per scheduler.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mn32_off(uint32_t row, uint32_t j) {        // byte offset of 16-byte chunk j (0..7) of a row
  return row * 128u + ((((j >> 1) ^ (row & 3u))) << 5) + ((j & 1u) << 4);
}

struct WgradShape {
  static constexpr int BM = 128, ST = 2, PW = 8, THREADS = (PW + 5) * 32, GROUP = 2;   // GROUP: tiles per accumulator drain
  static constexpr uint32_t G = BM * 128;                       // one 32-column group of a tile (hi or lo)
  static constexpr uint32_t STAGE = 6 * G;                      // P hi (2 groups), P lo (2), Q hi, Q lo
  static constexpr uint32_t OFF_BAR = ST * STAGE, TOTAL = OFF_BAR + (2 * ST + 4) * 8 + 16;
};

template <bool P_IS_DH>
__global__ void __launch_bounds__(WgradShape::THREADS, 1)
wgrad_tc_kernel(const float* __restrict__ dh, const float* __restrict__ x, float* __restrict__ grads, long long off_W, unsigned M) {
  using S = WgradShape;
  constexpr int BM = S::BM, ST = S::ST, PT = S::PW * 32;
  constexpr uint32_t G = S::G;
  // M = 64, N = 32, both operands MN-major (bits 15, 16), tf32 inputs, fp32 accumulation
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) |
                             ((uint32_t)(64 >> 4) << 24);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();
  unsigned char* sm = smem_raw;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);                // [ST] producer arrivals
  uint64_t* empty = full + ST;                                                  // [ST] tcgen05.commit
  uint64_t* acc_full = empty + ST;                                              // [2] tcgen05.commit at the end of a group
  uint64_t* acc_empty = acc_full + 2;                                           // [2] 128 epilogue arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const float* Pm = P_IS_DH ? dh : x;                                           // [M][64]
  const float* Qm = P_IS_DH ? x : dh;                                           // [M][32]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned ntiles = (M + BM - 1) / BM;
  const unsigned total = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  constexpr unsigned GROUP = S::GROUP;
  const unsigned ngroups = (total + GROUP - 1) / GROUP;
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  if (tid == 32) {
    for (int i = 0; i < ST; ++i) { mbar_init(full + i, PT); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 128); }
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp < S::PW) {
    // P: 128 rows x 16 chunks (8 per thread), Q: 128 rows x 8 chunks (4 per thread); 16 / 8 lanes cover one row
    float4 p0[8], q0[4], p1[8], q1[4];
    auto request = [&](unsigned t, float4 (&pr)[8], float4 (&qr)[4]) {
      const unsigned row0 = (blockIdx.x + t * gridDim.x) * BM;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * PT + tid;
        const unsigned r = row0 + (unsigned)(idx >> 4);
        pr[it] = r < M ? ldg4_stream(Pm + (size_t)r * 64 + 4 * (idx & 15)) : f4zero();
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = it * PT + tid;
        const unsigned r = row0 + (unsigned)(idx >> 3);
        qr[it] = r < M ? ldg4_stream(Qm + (size_t)r * 32 + 4 * (idx & 7)) : f4zero();
      }
    };
    auto publish = [&](unsigned t, const float4 (&pr)[8], const float4 (&qr)[4]) {
      const unsigned s = t % ST, n = t / ST;
      if (n > 0) mbar_wait(empty + s, (n - 1) & 1u);
      unsigned char* st = sm + s * S::STAGE;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * PT + tid;
        const uint32_t j = (uint32_t)(idx & 15), off = (j >> 3) * G + mn32_off((uint32_t)(idx >> 4), j & 7u);
        const float4 v = pr[it];
        float4 lo;
        lo.x = lo_tf32(v.x); lo.y = lo_tf32(v.y); lo.z = lo_tf32(v.z); lo.w = lo_tf32(v.w);
        *reinterpret_cast<float4*>(st + off) = v;
        *reinterpret_cast<float4*>(st + 2 * G + off) = lo;
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = it * PT + tid;
        const uint32_t off = mn32_off((uint32_t)(idx >> 3), (uint32_t)(idx & 7));
        const float4 v = qr[it];
        float4 lo;
        lo.x = lo_tf32(v.x); lo.y = lo_tf32(v.y); lo.z = lo_tf32(v.z); lo.w = lo_tf32(v.w);
        *reinterpret_cast<float4*>(st + 4 * G + off) = v;
        *reinterpret_cast<float4*>(st + 5 * G + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(full + s);
    };
    if (total > 0) request(0, p0, q0);
    if (total > 1) request(1, p1, q1);
    for (unsigned t = 0; t < total; t += 2) {
      publish(t, p0, q0);
      if (t + 2 < total) request(t + 2, p0, q0);
      if (t + 1 < total) {
        publish(t + 1, p1, q1);
        if (t + 3 < total) request(t + 3, p1, q1);
      }
    }
  } else if (warp < S::PW + 4) {
    // ------------------------------------------------------------------ epilogue: running sum in registers
    // accumulator row r (of 64) sits in TMEM lane 32 (r / 16) + r % 16; warp w owns lane quarter w % 4
    const int quarter = warp & 3;
    float sum[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) sum[c] = 0.f;
    for (unsigned g = 0; g < ngroups; ++g) {
      const unsigned acc = g & 1u;
      mbar_wait(acc_full + acc, (g >> 1) & 1u);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + acc * 32, v);
      tc_fence_before();
      mbar_arrive(acc_empty + acc);
#pragma unroll
      for (int c = 0; c < 32; ++c) sum[c] += v[c];
    }
    if (lane < 16 && total > 0) {
      const int r = quarter * 16 + lane;                                        // P column
      float* out = grads + off_W;
#pragma unroll
      for (int c = 0; c < 32; ++c) {                                            // Q column
        if (P_IS_DH) atomicAdd(out + (size_t)r * 32 + c, sum[c]);               // dW [64][32]: row = dh column, col = x column
        else atomicAdd(out + (size_t)c * 64 + r, sum[c]);                       // dW [32][64]: row = dh column (Q), col = y1 column (P)
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    for (unsigned t = 0; t < total; ++t) {
      const unsigned s = t % ST, g = t / GROUP, acc = g & 1u;
      const bool first = t % GROUP == 0, last = (t % GROUP == GROUP - 1) || t == total - 1;
      if (first && g >= 2) mbar_wait(acc_empty + acc, ((g >> 1) - 1) & 1u);     // the drain of group g - 2 is done
      mbar_wait(full + s, (t / ST) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p_hi = base + s * S::STAGE, p_lo = p_hi + 2 * G, q_hi = p_hi + 4 * G, q_lo = p_hi + 5 * G;
        const uint32_t d = tmem + acc * 32;
#pragma unroll
        for (int ks = 0; ks < BM / 8; ++ks) {                                   // 8 rows = 1024 B per K step
          const uint32_t ko = (uint32_t)ks * 1024u;
          umma_tf32(d, umma_desc_mn32(p_hi + ko, G), umma_desc_mn32(q_hi + ko, G), IDESC, (!first || ks > 0) ? 1u : 0u);
          umma_tf32(d, umma_desc_mn32(p_lo + ko, G), umma_desc_mn32(q_hi + ko, G), IDESC, 1);
          umma_tf32(d, umma_desc_mn32(p_hi + ko, G), umma_desc_mn32(q_lo + ko, G), IDESC, 1);
        }
        umma_commit(empty + s);
        if (last) umma_commit(acc_full + acc);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// 1 = done, 0 = shape not covered / switched off (GATRES_WGRAD_TC=0), < 0 = error
int wgrad_tc_dispatch(int NO, int KI, const float* dh, const float* x, float* grads, long long off_W, unsigned M, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("GATRES_WGRAD_TC");
    enabled = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  if (!enabled || !((NO == 64 && KI == 32) || (NO == 32 && KI == 64))) return 0;
  using S = WgradShape;
  const bool p_is_dh = NO == 64;
  auto kern = p_is_dh ? wgrad_tc_kernel<true> : wgrad_tc_kernel<false>;
  static bool configured[2] = {false, false};
  if (!configured[p_is_dh]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL) != cudaSuccess)
      return check_launch("wgrad_tc: smem attribute");
    configured[p_is_dh] = true;
  }
  const unsigned ntiles = (M + 127) / 128;
  unsigned grid = (unsigned)sm_count();
  if (grid > ntiles) grid = ntiles;
  launch_kernel(kern, dim3(grid), dim3(S::THREADS), (size_t)S::TOTAL, st, dh, x, grads, off_W, M);
  const int rc = check_launch("wgrad_tc");
  return rc == GATRES_OK ? 1 : rc;
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of the conv1 projection of the nc = 32 layers in ONE pass over dh, everything on tcgen05:
//   dx = dh W (+ residual gradient) (* ReLU mask)      dh [M, 64] K-major SWIZZLE_128B tile  x  W as B[n = ki][k = no]
//   dW += dh^T x                                        dh and x as MN-major BASE32B tiles (as in wgrad_tc_kernel)
// The first fused form (linear_tc.cu::linear_bwd_fused_kernel) computed dW with mma.sync from the same tiles and was held
// by the legacy tensor path (768 instructions per 128 rows).  Here the producers write dh twice — the two layouts differ
// only in the swizzle inside the 128-byte line — which costs 160 KB per 128-row stage, so there is ONE stage: producers
// and MMAs of a tile alternate (~2 k cycles each) under the ~4.4 k cycles the tile's 96 KB take at the SM's share of the
// HBM bandwidth, while the epilogue of the previous tile runs on its own accumulator: four epilogue warps request the
// residual / ReLU-reference rows BEFORE they wait for the accumulator, park the rows in a swizzled staging buffer and
// store 128-byte lines; the same warps drain the weight-gradient accumulator every two tiles into registers.
// ---------------------------------------------------------------------------------------------------------------------
template <int NO, int KI>
struct Fused2Shape {
  static_assert((NO == 64 && KI == 32) || (NO == 32 && KI == 64), "nc = 32 layers");
  static constexpr int BM = 128, PW = 8, THREADS = (PW + 5) * 32, GROUP = 2;
  static constexpr int GD = NO / 32, GX = KI / 32;              // 32-column groups (= K-major atoms) of dh and x
  static constexpr uint32_t G = BM * 128;                       // [128 rows][128 B]
  static constexpr uint32_t OFF_DHK = 0;                        // dh K-major: hi atoms | lo atoms
  static constexpr uint32_t OFF_DHM = 2 * GD * G;               // dh MN-major: hi groups | lo groups
  static constexpr uint32_t OFF_XM = 4 * GD * G;                // x MN-major: hi groups | lo groups
  static constexpr uint32_t W_PART = NO * KI * 4;               // W^T K-major [KI rows x NO]: hi | lo
  static constexpr uint32_t OFF_W = OFF_XM + 2 * GX * G;
  static constexpr uint32_t OFF_EPI = OFF_W + 2 * W_PART;       // 4 warps x 4 KB
  static constexpr uint32_t OFF_BAR = OFF_EPI + 4 * 4096;
  static constexpr uint32_t TOTAL = OFF_BAR + 10 * 8 + 16;
  static constexpr uint32_t TMEM_COLS = KI == 32 ? 128 : 256;   // dx accumulators: columns 0 / KI, dW: 2 KI / 2 KI + 32
  static_assert(TOTAL <= 232448, "shared memory budget");
};

template <int NO, int KI>
__global__ void __launch_bounds__((Fused2Shape<NO, KI>::THREADS), 1)
linear_bwd_fused2_kernel(const float* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ W,
                         const float* __restrict__ add, const float* __restrict__ relu_ref, float* __restrict__ dx,
                         float* __restrict__ grads, long long off_W, unsigned M) {
  using S = Fused2Shape<NO, KI>;
  constexpr int BM = S::BM, PT = S::PW * 32, GD = S::GD, GX = S::GX;
  constexpr bool P_IS_DH = NO == 64;                            // the 64-column operand is A (M = 64) of the weight gradient
  constexpr unsigned GROUP = S::GROUP;
  constexpr uint32_t G = S::G;
  constexpr uint32_t IDESC_D = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KI >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  constexpr uint32_t IDESC_W = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) |
                               ((uint32_t)(64 >> 4) << 24);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();
  unsigned char* sm = smem_raw;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);                // producers -> MMA
  uint64_t* empty = full + 1;                                                   // tcgen05.commit: the stage may be rewritten
  uint64_t* dacc_full = empty + 1;                                              // [2]
  uint64_t* dacc_empty = dacc_full + 2;                                         // [2] 128 epilogue arrivals
  uint64_t* wacc_full = dacc_empty + 2;                                         // [2]
  uint64_t* wacc_empty = wacc_full + 2;                                         // [2] 128 epilogue arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wacc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned ntiles = (M + BM - 1) / BM;
  const unsigned total = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0) tmem_alloc(tmem_slot, S::TMEM_COLS);
  if (tid == 32) {
    mbar_init(full, PT);
    mbar_init(empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(dacc_full + i, 1); mbar_init(dacc_empty + i, 128);
      mbar_init(wacc_full + i, 1); mbar_init(wacc_empty + i, 128);
    }
    mbar_fence_init();
  }
  for (int idx = tid; idx < NO * KI; idx += blockDim.x) {                       // B[n = ki][k = no] = W[no][ki], K-major, hi / lo
    const int n = idx % KI, k = idx / KI;
    const float w = __ldg(W + (size_t)k * KI + n);
    const uint32_t off = swz_off((uint32_t)n, (uint32_t)k, KI);
    *reinterpret_cast<float*>(sm + S::OFF_W + off) = w;
    *reinterpret_cast<float*>(sm + S::OFF_W + S::W_PART + off) = lo_tf32(w);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp < S::PW) {
    // ------------------------------------------------------------------ producers: one tile ahead in registers
    constexpr int ND = BM * (NO / 4) / PT, NX = BM * (KI / 4) / PT;             // float4 per thread and tile
    float4 pr[ND], qr[NX];
    auto request = [&](unsigned t) {
      const unsigned row0 = (blockIdx.x + t * gridDim.x) * BM;
#pragma unroll
      for (int it = 0; it < ND; ++it) {
        const int idx = it * PT + tid;
        const unsigned r = row0 + (unsigned)(idx / (NO / 4));
        pr[it] = r < M ? ldg4_stream(dh + (size_t)r * NO + 4 * (idx % (NO / 4))) : f4zero();
      }
#pragma unroll
      for (int it = 0; it < NX; ++it) {
        const int idx = it * PT + tid;
        const unsigned r = row0 + (unsigned)(idx / (KI / 4));
        qr[it] = r < M ? ldg4_stream(x + (size_t)r * KI + 4 * (idx % (KI / 4))) : f4zero();
      }
    };
    if (total > 0) request(0);
    for (unsigned t = 0; t < total; ++t) {
      if (t > 0) mbar_wait(empty, (t - 1) & 1u);                                // the MMAs of tile t - 1 have read the stage
#pragma unroll
      for (int it = 0; it < ND; ++it) {
        const int idx = it * PT + tid;
        const uint32_t row = (uint32_t)(idx / (NO / 4)), j = (uint32_t)(idx % (NO / 4));
        const uint32_t ok = swz_off(row, 4u * j, BM);                          // K-major (dx)
        const uint32_t om = (j >> 3) * G + mn32_off(row, j & 7u);               // MN-major (dW)
        const float4 v = pr[it];
        float4 lo;
        lo.x = lo_tf32(v.x); lo.y = lo_tf32(v.y); lo.z = lo_tf32(v.z); lo.w = lo_tf32(v.w);
        *reinterpret_cast<float4*>(sm + S::OFF_DHK + ok) = v;
        *reinterpret_cast<float4*>(sm + S::OFF_DHK + GD * G + ok) = lo;
        *reinterpret_cast<float4*>(sm + S::OFF_DHM + om) = v;
        *reinterpret_cast<float4*>(sm + S::OFF_DHM + GD * G + om) = lo;
      }
#pragma unroll
      for (int it = 0; it < NX; ++it) {
        const int idx = it * PT + tid;
        const uint32_t row = (uint32_t)(idx / (KI / 4)), j = (uint32_t)(idx % (KI / 4));
        const uint32_t off = (j >> 3) * G + mn32_off(row, j & 7u);
        const float4 v = qr[it];
        float4 lo;
        lo.x = lo_tf32(v.x); lo.y = lo_tf32(v.y); lo.z = lo_tf32(v.z); lo.w = lo_tf32(v.w);
        *reinterpret_cast<float4*>(sm + S::OFF_XM + off) = v;
        *reinterpret_cast<float4*>(sm + S::OFF_XM + GX * G + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(full);
      if (t + 1 < total) request(t + 1);
    }
  } else if (warp < S::PW + 4) {
    // ------------------------------------------------------------------ epilogue: dx rows + running dW sum
    const int quarter = warp & 3;
    unsigned char* stg = sm + S::OFF_EPI + quarter * 4096;
    float sum[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) sum[c] = 0.f;
    for (unsigned t = 0; t < total; ++t) {
      const unsigned acc = t & 1u, row0 = (blockIdx.x + t * gridDim.x) * BM + quarter * 32;
      float v[32];
#pragma unroll
      for (int cb = 0; cb < KI / 32; ++cb) {
        float4 addv[8], refv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {                                           // requested before the wait for the accumulator
          const int r = j * 4 + (lane >> 3), c = lane & 7;
          const size_t off = (size_t)(row0 + r) * KI + cb * 32 + 4 * c;
          const bool ok = row0 + r < M;
          addv[j] = (ok && add != nullptr) ? ldg4_stream(add + off) : f4zero();
          refv[j] = (ok && relu_ref != nullptr) ? ldg4_stream(relu_ref + off) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
        if (cb == 0) {
          mbar_wait(dacc_full + acc, (t >> 1) & 1u);
          tc_fence_after();
        }
        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + acc * KI + cb * 32, v);
        if (cb == KI / 32 - 1) {
          tc_fence_before();
          mbar_arrive(dacc_empty + acc);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4*>(stg + lane * 128 + ((k ^ (lane & 7)) << 4)) =
              make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = j * 4 + (lane >> 3), c = lane & 7;
          float4 o = *reinterpret_cast<const float4*>(stg + r * 128 + ((c ^ (r & 7)) << 4));
          add4(o, addv[j]);
          o.x = refv[j].x > 0.f ? o.x : 0.f; o.y = refv[j].y > 0.f ? o.y : 0.f;
          o.z = refv[j].z > 0.f ? o.z : 0.f; o.w = refv[j].w > 0.f ? o.w : 0.f;
          if (row0 + r < M) st4(dx + (size_t)(row0 + r) * KI + cb * 32 + 4 * c, o);
        }
        __syncwarp();
      }
      if (t % GROUP == GROUP - 1 || t == total - 1) {                            // the group's weight-gradient accumulator is complete
        const unsigned g = t / GROUP, wa = g & 1u;
        mbar_wait(wacc_full + wa, (g >> 1) & 1u);
        tc_fence_after();
        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + 2 * KI + wa * 32, v);
        tc_fence_before();
        mbar_arrive(wacc_empty + wa);
#pragma unroll
        for (int c = 0; c < 32; ++c) sum[c] += v[c];
      }
    }
    if (lane < 16 && total > 0) {                                               // accumulator row r of 64 = TMEM lane 32 (r / 16) + r % 16
      const int r = quarter * 16 + lane;                                        // column of the 64-column operand
      float* out = grads + off_W;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (P_IS_DH) atomicAdd(out + (size_t)r * 32 + c, sum[c]);               // dW [64][32]: row = dh column, col = x column
        else atomicAdd(out + (size_t)c * 64 + r, sum[c]);                       // dW [32][64]: row = dh column, col = x column
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    for (unsigned t = 0; t < total; ++t) {
      const unsigned acc = t & 1u, g = t / GROUP, wa = g & 1u;
      const bool first = t % GROUP == 0, last = (t % GROUP == GROUP - 1) || t == total - 1;
      if (t >= 2) mbar_wait(dacc_empty + acc, ((t >> 1) - 1) & 1u);
      if (first && g >= 2) mbar_wait(wacc_empty + wa, ((g >> 1) - 1) & 1u);
      mbar_wait(full, t & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ak_hi = base + S::OFF_DHK, ak_lo = ak_hi + GD * G;
        const uint32_t w_hi = base + S::OFF_W, w_lo = w_hi + S::W_PART;
        const uint32_t dm_hi = base + S::OFF_DHM, dm_lo = dm_hi + GD * G, xm_hi = base + S::OFF_XM, xm_lo = xm_hi + GX * G;
        const uint32_t pm_hi = P_IS_DH ? dm_hi : xm_hi, pm_lo = P_IS_DH ? dm_lo : xm_lo;
        const uint32_t qm_hi = P_IS_DH ? xm_hi : dm_hi, qm_lo = P_IS_DH ? xm_lo : dm_lo;
        const uint32_t dd = tmem + acc * KI, dw = tmem + 2 * KI + wa * 32;
#pragma unroll
        for (int ks = 0; ks < NO / 8; ++ks) {                                   // dx: K = the dh columns
          const uint32_t ka = (uint32_t)(ks >> 2) * G + (uint32_t)(ks & 3) * 32u;
          const uint32_t kb = (uint32_t)(ks >> 2) * (KI * 128u) + (uint32_t)(ks & 3) * 32u;
          umma_tf32(dd, umma_desc_k128(ak_hi + ka), umma_desc_k128(w_hi + kb), IDESC_D, ks > 0 ? 1u : 0u);
          umma_tf32(dd, umma_desc_k128(ak_lo + ka), umma_desc_k128(w_hi + kb), IDESC_D, 1);
          umma_tf32(dd, umma_desc_k128(ak_hi + ka), umma_desc_k128(w_lo + kb), IDESC_D, 1);
        }
        umma_commit(dacc_full + acc);
#pragma unroll
        for (int ks = 0; ks < BM / 8; ++ks) {                                   // dW: K = 128 rows
          const uint32_t ko = (uint32_t)ks * 1024u;
          umma_tf32(dw, umma_desc_mn32(pm_hi + ko, G), umma_desc_mn32(qm_hi + ko, G), IDESC_W, (!first || ks > 0) ? 1u : 0u);
          umma_tf32(dw, umma_desc_mn32(pm_lo + ko, G), umma_desc_mn32(qm_hi + ko, G), IDESC_W, 1);
          umma_tf32(dw, umma_desc_mn32(pm_hi + ko, G), umma_desc_mn32(qm_lo + ko, G), IDESC_W, 1);
        }
        umma_commit(empty);
        if (last) umma_commit(wacc_full + wa);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, S::TMEM_COLS);
}

template <int NO, int KI>
static int launch_fused2(const float* dh, const float* x, const float* W, const float* add, const float* relu_ref, float* dx,
                         float* grads, long long off_W, unsigned M, cudaStream_t st) {
  using S = Fused2Shape<NO, KI>;
  auto kern = linear_bwd_fused2_kernel<NO, KI>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::TOTAL) != cudaSuccess)
      return check_launch("linear_bwd_fused2: smem attribute");
    configured = true;
  }
  const unsigned ntiles = (M + 127) / 128;
  unsigned grid = (unsigned)sm_count();
  if (grid > ntiles) grid = ntiles;
  launch_kernel(kern, dim3(grid), dim3(S::THREADS), (size_t)S::TOTAL, st, dh, x, W, add, relu_ref, dx, grads, off_W, M);
  return check_launch("linear_bwd_fused2");
}

// 1 = done, 0 = not covered / switched off (GATRES_LINEAR_BWD_FUSED2=0; =1: conv1 only, default 2: conv1 and conv2), < 0 = error
int linear_bwd_fused2_dispatch(int NO, int KI, const float* dh, const float* x, const float* W, const float* add,
                               const float* relu_ref, float* dx, float* grads, long long off_W, unsigned M, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("GATRES_LINEAR_BWD_FUSED2");
    enabled = e == nullptr ? 2 : atoi(e);
  }
  int rc;
  if (enabled >= 1 && NO == 64 && KI == 32) rc = launch_fused2<64, 32>(dh, x, W, add, relu_ref, dx, grads, off_W, M, st);
  else if (enabled >= 2 && NO == 32 && KI == 64) rc = launch_fused2<32, 64>(dh, x, W, add, relu_ref, dx, grads, off_W, M, st);
  else return 0;
  return rc == GATRES_OK ? 1 : rc;
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
  static int pair = -1;
  if (pair < 0) {
    const char* e = getenv("GATRES_TC_PAIR");             // opt-in: measured equal to the single-CTA form (see the header)
    pair = (e != nullptr && atoi(e) != 0) ? 1 : 0;
  }
  int rc;
  if (pair) {
    rc = 0;
    if (KK == 128 && NN == 256 && H == 2) rc = launch_tc_pair<128, 256, 2>(A, W, e0, e1, Cout, s0, s1, M, st);
    else if (KK == 256 && NN == 128 && H == 1) rc = launch_tc_pair<256, 128, 1>(A, W, e0, e1, Cout, s0, s1, M, st);
    else if (KK == 64 && NN == 128 && H == 2) rc = launch_tc_pair<64, 128, 2>(A, W, e0, e1, Cout, s0, s1, M, st);
    if (rc != 0) return rc;
  }
  static int narrow = -1;
  if (narrow < 0) {
    const char* e = getenv("GATRES_TC_NARROW2");
    narrow = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  if (KK == 32 && NN == 64 && H == 2) {
    if (!narrow) return 0;
    rc = launch_tc_wide2<32, 64, 2>(A, W, e0, e1, Cout, s0, s1, M, st);
  } else if (KK == 64 && NN == 32 && H == 1) {
    if (!narrow) return 0;
    rc = launch_tc_wide2<64, 32, 1>(A, W, e0, e1, Cout, s0, s1, M, st);
  } else if (KK == 128 && NN == 256 && H == 2) rc = launch_tc_wide2<128, 256, 2>(A, W, e0, e1, Cout, s0, s1, M, st);
  else if (KK == 256 && NN == 128 && H == 1) rc = launch_tc_wide2<256, 128, 1>(A, W, e0, e1, Cout, s0, s1, M, st);
  else if (KK == 64 && NN == 128 && H == 2) rc = launch_tc_wide2<64, 128, 2>(A, W, e0, e1, Cout, s0, s1, M, st);
  else if (KK == 128 && NN == 64 && H == 1) rc = launch_tc_wide2<128, 64, 1>(A, W, e0, e1, Cout, s0, s1, M, st);   // conv2, nc = 64
  else return 0;
  return rc == GATRES_OK ? 1 : rc;
}

}  // namespace gatres
