// Projection GEMMs on the 5th-generation tensor cores (sm_100a): tcgen05.mma
// kind::tf32 with the accumulator in TMEM, error-compensated to fp32 accuracy
// ("3xTF32":  A B ~= A_hi B_hi + A_lo B_hi + A_hi B_lo, A = A_hi + A_lo with both
// parts rounded to TF32 by the staging threads, fp32 accumulation in TMEM), because
// the parity contract is 1e-4 relative through 30 chained projections and a single
// TF32 pass (2^-11 per operand) does not meet it.
//
//   MODE 0  h = x W^T  + attention-score epilogue   (GATConv.forward step 1,
//           /root/reference/gnn_pressure_estimation/GraphModels.py:464-465)
//   MODE 1  dx = dh W (+ residual gradient) (* ReLU mask)   (its data gradient)
//
// One CTA = 128 threads = one 128-row tile at a time (UMMA M = 128, N = 32 or 64,
// K = 8 per instruction).  Operands live in shared memory in the canonical K-major
// SWIZZLE_128B layout (rows of 128 B, 16-byte chunk index XOR row%8, 1024-byte
// 8-row atoms); thread 0 issues the MMAs, completion is signalled through
// tcgen05.commit on an mbarrier, and each of the four warps drains its 32 TMEM
// lanes (= 32 rows) with tcgen05.ld: one thread owns one output row, so the
// attention scores are in-thread dot products.  Several CTAs share an SM
// (48-80 KB of shared memory, 32-64 TMEM columns each), which overlaps one CTA's
// loads with another's MMAs and epilogue.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace gatres {


template <int KK, int NN, int MODE, int H>
__global__ void __launch_bounds__(128)
gemm_tc_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ e0,
               const float* __restrict__ e1, float* __restrict__ Cout, float* __restrict__ s0,
               float* __restrict__ s1, unsigned M) {
  constexpr int BM = 128;
  constexpr uint32_t A_BYTES = BM * KK * 4, B_BYTES = NN * KK * 4;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  static_assert(KK % 32 == 0 && NN % 16 == 0 && NN >= 32 && NN <= 64 && NN <= 2 * KK, "unsupported tensor-core shape");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t base = smem_u32(smem_raw);                    // SWIZZLE_128B atoms must be 1024 B aligned
  if ((base & 1023u) != 0) __trap();                           // (no slack is allocated: two CTAs of the K=64 shape fill an SM)
  unsigned char* sm = smem_raw;
  const uint32_t a_raw0 = base, a_lo = base + A_BYTES, a_raw1 = base + 2 * A_BYTES, b_hi = base + 3 * A_BYTES,
                 b_lo = b_hi + B_BYTES;
  float* att = reinterpret_cast<float*>(sm + 3 * A_BYTES + 2 * B_BYTES);  // [2][NN] (MODE 0)
  uint64_t* bar = reinterpret_cast<uint64_t*>(att + 2 * NN);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const unsigned ntiles = (M + BM - 1) / BM;

  // ---- prologue (constant data only: runs under the previous kernel's tail) ----
  if (warp == 0) tmem_alloc(tmem_slot, NN);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  for (int idx = tid; idx < NN * KK; idx += 128) {            // B operand [NN rows x KK], K-major, split hi/lo
    const int n = idx / KK, k = idx % KK;
    const float w = MODE == 0 ? __ldg(W + (size_t)n * KK + k) : __ldg(W + (size_t)k * NN + n);
    const float hi = to_tf32(w), lo = to_tf32(w - hi);
    const uint32_t off = swz_off(n, k, NN);
    *reinterpret_cast<float*>(sm + 3 * A_BYTES + off) = hi;
    *reinterpret_cast<float*>(sm + 3 * A_BYTES + B_BYTES + off) = lo;
  }
  if (MODE == 0)
    for (int idx = tid; idx < NN; idx += 128) {
      att[idx] = __ldg(e0 + idx);
      att[NN + idx] = __ldg(e1 + idx);
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  // ---- software pipeline over this CTA's tiles ----
  // buffers: [raw0][lo][raw1].  raw_c holds the fp32 tile as loaded (the tensor core reads its TF32 truncation,
  // hi = trunc(x)), lo = rna(x - trunc(x)) is written by the staging threads, and the NEXT tile's cp.async flies
  // into the other raw buffer during this tile's MMAs and epilogue.
  constexpr int CHUNKS = BM * KK / 4;                         // 16-byte chunks per tile
  auto issue_load = [&](unsigned t, uint32_t raw_addr) {
#pragma unroll
    for (int q = 0; q < CHUNKS / 128; ++q) {
      const int idx = q * 128 + tid;
      const uint32_t row = idx / (KK / 4), c = idx % (KK / 4);
      const unsigned grow = t * BM + row;
      const bool ok = grow < M;
      tc_cp_async16(raw_addr + swz_off(row, 4 * c, BM), A + (size_t)(ok ? grow : 0) * KK + 4 * c, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (gridDim.x >= ntiles) pdl_launch_dependents();
  uint32_t phase = 0, it = 0;
  if (blockIdx.x < ntiles) issue_load(blockIdx.x, a_raw0);
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const uint32_t cur = it & 1u;
    const uint32_t raw_c = cur ? a_raw1 : a_raw0, raw_n = cur ? a_raw0 : a_raw1;
    const uint32_t stg = cur ? A_BYTES : 0u;                  // epilogue staging: raw0+lo (even) / lo+raw1 (odd)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (tile + gridDim.x < ntiles) issue_load(tile + gridDim.x, raw_n);
    // ---- lo part of every chunk this thread loaded ----
#pragma unroll
    for (int q = 0; q < CHUNKS / 128; ++q) {
      const int idx = q * 128 + tid;
      const uint32_t row = idx / (KK / 4), c = idx % (KK / 4);
      const uint32_t off = swz_off(row, 4 * c, BM);
      const float4 x = *reinterpret_cast<const float4*>(sm + (raw_c - base) + off);
      float4 lo;
      lo.x = lo_tf32(x.x); lo.y = lo_tf32(x.y); lo.z = lo_tf32(x.z); lo.w = lo_tf32(x.w);
      *reinterpret_cast<float4*>(sm + A_BYTES + off) = lo;
    }
    fence_proxy_async();                                      // generic-proxy smem writes -> tensor-core (async) proxy
    __syncthreads();

    // ---- one thread issues the 3 x KK/8 MMAs and commits them to the mbarrier ----
    if (tid == 0) {                                           // (elect.sync issue measured slower here: 73 vs 67 us at K = 32)
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < KK / 8; ++s) {
        const uint32_t ka = (uint32_t)(s >> 2) * BM * 128u + (uint32_t)(s & 3) * 32u;     // A: atom block + 32 B per K slice
        const uint32_t kb = (uint32_t)(s >> 2) * NN * 128u + (uint32_t)(s & 3) * 32u;
        umma_tf32(tmem, umma_desc_k128(raw_c + ka), umma_desc_k128(b_hi + kb), IDESC, s > 0);
        umma_tf32(tmem, umma_desc_k128(a_lo + ka), umma_desc_k128(b_hi + kb), IDESC, 1);
        umma_tf32(tmem, umma_desc_k128(raw_c + ka), umma_desc_k128(b_lo + kb), IDESC, 1);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();

    // ---- epilogue: thread = output row; TMEM lanes [32*warp, 32*warp+32) ----
    // The row is parked in shared memory (the operand buffers are dead once the MMAs have completed; 16-byte
    // chunk index XOR row keeps the column-strided stores at the 4-wavefront minimum) so that the global
    // stores below are fully coalesced instead of 32 rows x 16 B per instruction.
    const unsigned row = tile * BM + tid;
    float ps[H], pd[H];
#pragma unroll
    for (int hh = 0; hh < H; ++hh) ps[hh] = pd[hh] = 0.f;
    constexpr int NCH = NN / 4;                                // 16-byte chunks per output row
#pragma unroll
    for (int cb = 0; cb < NN / 32; ++cb) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
      if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int n = cb * 32 + k, hh = n / (NN / H);
          ps[hh] = fmaf(v[k], att[n], ps[hh]);
          pd[hh] = fmaf(v[k], att[NN + n], pd[hh]);
        }
      }
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        const int c = cb * 8 + k / 4;
        *reinterpret_cast<float4*>(sm + stg + (size_t)tid * (NN * 4) + (((c ^ tid) & (NCH - 1)) << 4)) =
            make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
      }
    }
    if (MODE == 0 && row < M) {
#pragma unroll
      for (int hh = 0; hh < H; ++hh) {
        s0[(size_t)row * H + hh] = ps[hh];
        s1[(size_t)row * H + hh] = pd[hh];
      }
    }
    tc_fence_before();
    __syncthreads();                                          // staging complete; TMEM accumulator free
    // coalesced write-out in batches of EPB chunks per thread: the residual-gradient / ReLU-reference loads of a whole
    // batch are issued before any of them is consumed (they were latency-exposed one by one: the data-gradient kernels
    // ran at 44 % of HBM peak against 55 % for the same GEMM forward)
    constexpr int EPB = 8, EPI = BM * NCH / 128;
    static_assert(EPI % EPB == 0, "epilogue batching");
#pragma unroll 1
    for (int q0 = 0; q0 < EPI; q0 += EPB) {
      float4 addv[EPB], refv[EPB];
      size_t off[EPB];
      bool okv[EPB];
#pragma unroll
      for (int u = 0; u < EPB; ++u) {
        const int idx = (q0 + u) * 128 + tid;
        const int r = idx / NCH, c = idx % NCH;
        const unsigned grow = tile * BM + r;
        okv[u] = grow < M;
        off[u] = (size_t)grow * NN + 4 * c;
        if (MODE == 1) {
          addv[u] = (okv[u] && e0 != nullptr) ? ldg4_stream(e0 + off[u]) : f4zero();
          refv[u] = (okv[u] && e1 != nullptr) ? ldg4_stream(e1 + off[u]) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
      }
#pragma unroll
      for (int u = 0; u < EPB; ++u) {
        const int idx = (q0 + u) * 128 + tid;
        const int r = idx / NCH, c = idx % NCH;
        if (okv[u]) {
          float4 t = *reinterpret_cast<const float4*>(sm + stg + (size_t)r * (NN * 4) + (((c ^ r) & (NCH - 1)) << 4));
          if (MODE == 1) {
            add4(t, addv[u]);
            t.x = refv[u].x > 0.f ? t.x : 0.f; t.y = refv[u].y > 0.f ? t.y : 0.f;
            t.z = refv[u].z > 0.f ? t.z : 0.f; t.w = refv[u].w > 0.f ? t.w : 0.f;
          }
          st4(Cout + off[u], t);
        }
      }
    }
    __syncthreads();                                          // staging (= operand buffers) free for the next tile
  }
  if (warp == 0) tmem_dealloc(tmem, NN);
}

template <int KK, int NN, int MODE, int H>
static int launch_tc(const float* A, const float* W, const float* e0, const float* e1, float* Cout, float* s0,
                     float* s1, unsigned M, cudaStream_t st, const char* what) {
  constexpr size_t smem = 3 * (size_t)128 * KK * 4 + 2 * (size_t)NN * KK * 4 + 2 * NN * 4 + 32;
  auto kern = gemm_tc_kernel<KK, NN, MODE, H>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch(what);
    configured = true;
  }
  const unsigned ntiles = (M + 127) / 128;
  unsigned per_sm = (unsigned)((228u * 1024u) / (smem + 1024u));       // +1 KB the system reserves per CTA
  per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
  unsigned grid = (unsigned)sm_count() * per_sm;
  if (grid > ntiles) grid = ntiles;
  launch_kernel(kern, dim3(grid), dim3(128), smem, st, A, W, e0, e1, Cout, s0, s1, M);
  return check_launch(what);
}


// ---------------------------------------------------------------------------------------------------------
// Fused backward of a projection (conv1 of the nc = 32 model: dh [rows, 64], x [rows, 32]): dx = dh W (+ residual)
// (* ReLU mask) AND dW += dh^T x in ONE pass over dh.  The two-kernel form reads dh twice (data gradient on tcgen05,
// weight gradient on mma.sync in a second launch that re-reads dh and x: 305 MB for an 8 KB result).  Here a persistent
// CTA (256 threads, two per SM) stages a 128-row tile of dh (SWIZZLE_128B K-major, the tcgen05 A operand) and of x
// once; one elected thread issues the data gradient's 3 x NO/8 tcgen05.mma into TMEM, and WHILE they run all eight
// warps contract the same shared tiles for the weight gradient with mma.sync 3xTF32 (fragments of dh^T straight from
// the swizzled tile: row steps are multiples of 8, so the swizzled offsets are loop invariants), accumulating in
// registers across all tiles of the CTA (atomics at the end).  Then the epilogue of gemm_tc MODE 1 (accumulator rows
// parked in shared memory, coalesced write-out with the residual / ReLU-reference loads batched).  The phases of a tile
// are serial inside a CTA, so the tile buffers are single and the SECOND CTA of the SM fills the gaps.
// Measured at 2048 snapshots (two kernels: 149.6 us for dh [., 64], 145.1 us for dh [., 32]): this form 134.4 / 149.0 us;
// one double-buffered CTA per SM 137.0 / 141.9; warp-specialised (four loader / data-gradient warps, four weight-gradient
// warps meeting at mbarriers) 129.6 / 150.6; weight gradient on the FP32 pipe instead (4 x 8 register patches, exact
// fp32) 150.2 / 196.3.  All forms are held by the weight gradient: 768 mma.sync per tile at ~20 cycles per scheduler on
// the legacy tensor path, which the tcgen05 MMAs also contend with; a tcgen05 weight gradient needs dh in a second
// shared layout (MN-major tf32 operands must be SWIZZLE_128B_BASE32B) and does not fit next to the data gradient's.
// Used for the dh [., 64] shape only.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned fb_lo_bits(float x) {
  const float r = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  return (__float_as_uint(r) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ void fb_mma(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NO, int KI>
struct FusedBwdPlan {                       // bytes; the epilogue staging [128][KI] lies over the low-part tile / the x tile
  static constexpr int BM = 128, LDX = KI + 4;
  static constexpr uint32_t DH = BM * NO * 4, XT = BM * LDX * 4, BW = KI * NO * 4, STG = BM * KI * 4;
  static constexpr uint32_t dh0 = 0, dhl = DH, x0 = 2 * DH, bh = (x0 + XT + 1023u) & ~1023u, bl = bh + BW, bar = bl + BW,
                            total = bar + 32;
  static constexpr uint32_t stg = dhl;      // dead once the MMAs have completed (DH + XT >= STG for both shapes)
  static_assert(DH + XT >= STG, "staging over the low-part and x tiles");
};

template <int NO, int KI>
__global__ void __launch_bounds__(256, 2)
linear_bwd_fused_kernel(const float* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ W,
                        const float* __restrict__ add, const float* __restrict__ relu_ref, float* __restrict__ dx,
                        float* __restrict__ grads, long long off_W, unsigned M) {
  using P = FusedBwdPlan<NO, KI>;
  constexpr int BM = P::BM, T = 256, LDX = P::LDX;
  constexpr int NT = KI / 8, TILES = (NO / 16) * NT, TPW = TILES / 8;
  static_assert(TILES % 8 == 0 && NT % TPW == 0, "a warp's weight-gradient tiles share one 16-row block of dh^T");
  static_assert(NO % 32 == 0 && (KI == 32 || KI == 64), "shapes of the nc = 32 model");
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KI >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  extern __shared__ __align__(1024) unsigned char sm[];
  const uint32_t base = smem_u32(sm);
  if ((base & 1023u) != 0) __trap();
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + P::bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const unsigned ntiles = (M + BM - 1) / BM;

  // ---- prologue: TMEM, barrier, W as the data gradient's B operand ([KI rows n][NO k], K-major, hi / lo) ----
  if (warp == 0) tmem_alloc(tmem_slot, KI);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  for (int idx = tid; idx < KI * NO; idx += T) {
    const int n = idx % KI, k = idx / KI;                      // W[k][n] as stored ([NO][KI]): coalesced reads
    const float w = __ldg(W + (size_t)k * KI + n);
    const float hi = to_tf32(w), lo = to_tf32(w - hi);
    const uint32_t off = swz_off(n, k, KI);
    *reinterpret_cast<float*>(sm + P::bh + off) = hi;
    *reinterpret_cast<float*>(sm + P::bl + off) = lo;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  auto issue_load = [&](unsigned tl) {
#pragma unroll
    for (int q = 0; q < BM * (NO / 4) / T; ++q) {
      const int idx = q * T + tid;
      const uint32_t row = idx / (NO / 4), c = idx % (NO / 4);
      const unsigned grow = tl * BM + row;
      const bool ok = grow < M;
      tc_cp_async16(base + P::dh0 + swz_off(row, 4 * c, BM), dh + (size_t)(ok ? grow : 0) * NO + 4 * c, ok ? 16 : 0);
    }
#pragma unroll
    for (int q = 0; q < BM * (KI / 4) / T; ++q) {
      const int idx = q * T + tid;
      const uint32_t row = idx / (KI / 4), c = idx % (KI / 4);
      const unsigned grow = tl * BM + row;
      const bool ok = grow < M;
      tc_cp_async16(base + P::x0 + (row * LDX + 4 * c) * 4u, x + (size_t)(ok ? grow : 0) * KI + 4 * c, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // weight-gradient tiles of this warp: output rows no0 .. no0 + 15 (columns of dh), columns ki0 + 8 q
  const int tile0 = warp * TPW, no0 = (tile0 / NT) * 16, ki0 = (tile0 % NT) * 8;
  int aoff[4];                                                 // float offsets of the four dh^T fragment elements, minus 32 * m0
  {
    const int c0 = no0 + g, kb = (c0 >> 5) * (BM * 32), ch = (c0 & 31) >> 2, lw = c0 & 3;
    aoff[0] = kb + t * 32 + ((ch ^ t) << 2) + lw;
    aoff[1] = kb + t * 32 + (((ch + 2) ^ t) << 2) + lw;
    aoff[2] = kb + (t + 4) * 32 + ((ch ^ (t + 4)) << 2) + lw;
    aoff[3] = kb + (t + 4) * 32 + (((ch + 2) ^ (t + 4)) << 2) + lw;
  }
  float acc[TPW][4], acl[TPW][4], acm[TPW][4];
#pragma unroll
  for (int q = 0; q < TPW; ++q)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[q][i] = acl[q][i] = acm[q][i] = 0.f;

  if (gridDim.x >= ntiles) pdl_launch_dependents();
  uint32_t phase = 0;
  if (blockIdx.x < ntiles) issue_load(blockIdx.x);
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // ---- low parts of the dh chunks this thread copied ----
#pragma unroll
    for (int q = 0; q < BM * (NO / 4) / T; ++q) {
      const int idx = q * T + tid;
      const uint32_t off = swz_off(idx / (NO / 4), 4 * (idx % (NO / 4)), BM);
      const float4 v = *reinterpret_cast<const float4*>(sm + P::dh0 + off);
      *reinterpret_cast<float4*>(sm + P::dhl + off) = make_float4(lo_tf32(v.x), lo_tf32(v.y), lo_tf32(v.z), lo_tf32(v.w));
    }
    fence_proxy_async();
    __syncthreads();                                           // tile visible CTA-wide
    // ---- data gradient: one thread issues 3 x NO / 8 MMAs ----
    if (warp == 0 && elect_one()) {
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < NO / 8; ++s) {
        const uint32_t ka = (uint32_t)(s >> 2) * BM * 128u + (uint32_t)(s & 3) * 32u;
        const uint32_t kb = (uint32_t)(s >> 2) * KI * 128u + (uint32_t)(s & 3) * 32u;
        umma_tf32(tmem, umma_desc_k128(base + P::dh0 + ka), umma_desc_k128(base + P::bh + kb), IDESC, s > 0);
        umma_tf32(tmem, umma_desc_k128(base + P::dhl + ka), umma_desc_k128(base + P::bh + kb), IDESC, 1);
        umma_tf32(tmem, umma_desc_k128(base + P::dh0 + ka), umma_desc_k128(base + P::bl + kb), IDESC, 1);
      }
      umma_commit(bar);
    }
    // ---- weight gradient of the tile on mma.sync while the MMAs above run ----
    {
      const float* hs = reinterpret_cast<const float*>(sm + P::dh0);
      const float* xs = reinterpret_cast<const float*>(sm + P::x0);
#pragma unroll 2
      for (int m0 = 0; m0 < BM; m0 += 8) {                     // rows past M were zero-filled
        const float* hr = hs + m0 * 32;
        const float av[4] = {hr[aoff[0]], hr[aoff[1]], hr[aoff[2]], hr[aoff[3]]};
        unsigned a[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = __float_as_uint(av[i]); al[i] = fb_lo_bits(av[i]); }
#pragma unroll
        for (int q = 0; q < TPW; ++q) {
          const float b0 = xs[(m0 + t) * LDX + ki0 + 8 * q + g], b1 = xs[(m0 + t + 4) * LDX + ki0 + 8 * q + g];
          const unsigned b[2] = {__float_as_uint(b0), __float_as_uint(b1)}, bl[2] = {fb_lo_bits(b0), fb_lo_bits(b1)};
          fb_mma(acl[q], al, b);
          fb_mma(acm[q], a, bl);
          fb_mma(acc[q], a, b);
        }
      }
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    __syncthreads();                                           // every warp is done with the tiles: staging may overwrite them
    // ---- epilogue: thread = (accumulator row, column half) -> shared staging -> coalesced write-out ----
    constexpr int NCH = KI / 4, HALF = KI / 2;
    {
      const int lane_row = tid & 127, half = tid >> 7;
      float v[HALF];
      if constexpr (HALF == 16) tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + half * HALF, v);
      else tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + half * HALF, v);
#pragma unroll
      for (int k = 0; k < HALF; k += 4) {
        const int c = (half * HALF + k) / 4;
        *reinterpret_cast<float4*>(sm + P::stg + (size_t)lane_row * (KI * 4) + (((c ^ lane_row) & (NCH - 1)) << 4)) =
            make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
      }
    }
    tc_fence_before();
    __syncthreads();                                           // staging complete; TMEM accumulator free
    const unsigned next = tile + gridDim.x;
    constexpr int EPB = 4, EPI = BM * NCH / T;
    static_assert(EPI % EPB == 0, "epilogue batching");
#pragma unroll 1
    for (int q0 = 0; q0 < EPI; q0 += EPB) {
      float4 addv[EPB], refv[EPB];
      size_t off[EPB];
      bool okv[EPB];
#pragma unroll
      for (int u = 0; u < EPB; ++u) {
        const int idx = (q0 + u) * T + tid;
        const int r = idx / NCH, c = idx % NCH;
        const unsigned grow = tile * BM + r;
        okv[u] = grow < M;
        off[u] = (size_t)grow * KI + 4 * c;
        addv[u] = (okv[u] && add != nullptr) ? ldg4_stream(add + off[u]) : f4zero();
        refv[u] = (okv[u] && relu_ref != nullptr) ? ldg4_stream(relu_ref + off[u]) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
#pragma unroll
      for (int u = 0; u < EPB; ++u) {
        const int idx = (q0 + u) * T + tid;
        const int r = idx / NCH, c = idx % NCH;
        if (okv[u]) {
          float4 o = *reinterpret_cast<const float4*>(sm + P::stg + (size_t)r * (KI * 4) + (((c ^ r) & (NCH - 1)) << 4));
          add4(o, addv[u]);
          o.x = refv[u].x > 0.f ? o.x : 0.f; o.y = refv[u].y > 0.f ? o.y : 0.f;
          o.z = refv[u].z > 0.f ? o.z : 0.f; o.w = refv[u].w > 0.f ? o.w : 0.f;
          st4(dx + off[u], o);
        }
      }
    }
    __syncthreads();                                           // staging read: its buffers may be refilled
    if (next < ntiles) issue_load(next);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  float* out = grads + off_W;
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    const int k = ki0 + 8 * q + 2 * t;
    atomicAdd(reinterpret_cast<float2*>(out + (size_t)(no0 + g) * KI + k),
              make_float2(acc[q][0] + acl[q][0] + acm[q][0], acc[q][1] + acl[q][1] + acm[q][1]));
    atomicAdd(reinterpret_cast<float2*>(out + (size_t)(no0 + g + 8) * KI + k),
              make_float2(acc[q][2] + acl[q][2] + acm[q][2], acc[q][3] + acl[q][3] + acm[q][3]));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, KI);
}

template <int NO, int KI>
static int launch_linear_bwd_fused(const float* dh, const float* x, const float* W, const float* add, const float* relu_ref,
                                   float* dx, float* grads, long long off_W, unsigned M, cudaStream_t st) {
  using P = FusedBwdPlan<NO, KI>;
  auto kern = linear_bwd_fused_kernel<NO, KI>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::total) != cudaSuccess)
      return check_launch("linear_bwd_fused: smem attribute");
    configured = true;
  }
  const unsigned ntiles = (M + 127) / 128;
  unsigned per_sm = (unsigned)((228u * 1024u) / (P::total + 1024u));
  per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
  unsigned grid = (unsigned)sm_count() * per_sm;
  if (grid > ntiles) grid = ntiles;
  launch_kernel(kern, dim3(grid), dim3(256), (size_t)P::total, st, dh, x, W, add, relu_ref, dx, grads, off_W, M);
  return check_launch("linear_bwd_fused");
}

int linear_bwd_fused2_dispatch(int NO, int KI, const float* dh, const float* x, const float* W, const float* add,
                               const float* relu_ref, float* dx, float* grads, long long off_W, unsigned M, cudaStream_t st);

// 1 = done, 0 = shape not covered (caller runs the two-kernel form), < 0 = error
int linear_bwd_fused_dispatch(int NO, int KI, const float* dh, const float* x, const float* W, const float* add,
                              const float* relu_ref, float* dx, float* grads, long long off_W, unsigned M, cudaStream_t st) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("GATRES_LINEAR_BWD_FUSED");
    enabled = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  if (!enabled) return 0;
  int rc = linear_bwd_fused2_dispatch(NO, KI, dh, x, W, add, relu_ref, dx, grads, off_W, M, st);   // all-tcgen05 form (linear_tc_wide.cu)
  if (rc != 0) return rc;
  if (NO == 64 && KI == 32) rc = launch_linear_bwd_fused<64, 32>(dh, x, W, add, relu_ref, dx, grads, off_W, M, st);
  else return 0;                              // (dh [., 32] / x [., 64]: the two-kernel form measured faster)
  return rc == GATRES_OK ? 1 : rc;
}

// ---------------------------------------------------------------------------------------------------------
// Wide shapes (nc = 64 / 128: K up to 256, N up to 256): the operands of a whole tile no longer fit shared
// memory, so K is walked in 32-column chunks (one SWIZZLE_128B atom column) through a software pipeline:
//   A chunk [128 x 32] fp32 : cp.async into a 3-deep ring two chunks ahead (the tensor core reads its TF32
//                             truncation as A_hi), A_lo = rna(x - trunc x) written by the staging threads;
//   B chunk [NN x 32]       : the matching columns of W, read through L2 (the weights of one projection are
//                             <= 128 KB and shared by every CTA), split into B_hi / B_lo by the staging threads;
//   12 MMAs per chunk (4 K-steps x {hi*hi, lo*hi, hi*lo}) accumulate into NN TMEM columns; the chunk's commit
//   arrives on the mbarrier of its stage, so the MMAs of chunk q run under the staging of chunk q + 1.
// 256 threads: all stage; in the epilogue thread t drains TMEM lane t % 128 (its warp's lane quarter) for the
// column half t / 128.  FP32 FFMA peaks at ~75 TFLOP/s on B200 and the nc = 128 projections are 4 x 65 kFLOP
// per node and block, which made the large model FFMA bound (profiles/r1_configs.md); 3xTF32 on tcgen05 puts
// them back under the HBM roofline.  MODE 0 only (projection + score epilogue).
// ---------------------------------------------------------------------------------------------------------
#ifdef GATRES_TC_PROF
#define TCP_DECL long long tcp_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tcp_last = clock64()
#define TCP(i) do { const long long now_ = clock64(); tcp_t[i] += now_ - tcp_last; tcp_last = now_; } while (0)
#else
#define TCP_DECL
#define TCP(i)
#endif

template <int KK, int NN, int H>
__global__ void __launch_bounds__(256)
gemm_tc_wide_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ att_src,
                    const float* __restrict__ att_dst, float* __restrict__ Cout, float* __restrict__ s0,
                    float* __restrict__ s1, unsigned M) {
  constexpr int BM = 128, KC = 32, NCHUNK = KK / KC, T = 256;
  constexpr uint32_t A_CH = BM * KC * 4, B_CH = NN * KC * 4;                 // bytes of one chunk buffer
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  static_assert(KK % KC == 0 && (NN == 128 || NN == 256) && (H == 1 || H == 2), "unsupported wide tensor-core shape");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();
  unsigned char* sm = smem_raw;
  // [A_raw x3][A_lo x2][B_hi x2][B_lo x2][att 2*NN][score partials 2*T][barriers]
  constexpr uint32_t OFF_ALO = 3 * A_CH, OFF_BHI = OFF_ALO + 2 * A_CH, OFF_BLO = OFF_BHI + 2 * B_CH,
                     OFF_END = OFF_BLO + 2 * B_CH;
  static_assert(OFF_END - OFF_ALO >= (uint32_t)BM * NN * 4, "epilogue staging must fit the A_lo/B buffers");
  float* att = reinterpret_cast<float*>(sm + OFF_END);
  float* part = att + 2 * NN;                                               // [2][T] score partials
  uint64_t* bar = reinterpret_cast<uint64_t*>(part + 2 * T);                // [2] one per stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const unsigned ntiles = (M + BM - 1) / BM;
  if (warp == 0) tmem_alloc(tmem_slot, NN);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    mbar_fence_init();
  }
  for (int idx = tid; idx < NN; idx += T) {
    att[idx] = __ldg(att_src + idx);
    att[NN + idx] = __ldg(att_dst + idx);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  const unsigned my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const unsigned total = my_tiles * NCHUNK;                                  // flattened (tile, chunk) sequence
  auto issue_a = [&](unsigned q) {                                           // A chunk of step q -> ring slot q % 3
    const unsigned tile = blockIdx.x + (q / NCHUNK) * gridDim.x, ch = q % NCHUNK;
    const uint32_t dst = base + (q % 3) * A_CH;
#pragma unroll
    for (int it = 0; it < BM * (KC / 4) / T; ++it) {
      const int idx = it * T + tid;
      const uint32_t row = idx / (KC / 4), c = idx % (KC / 4);
      const unsigned grow = tile * BM + row;
      const bool ok = grow < M;
      tc_cp_async16(dst + swz_off(row, 4 * c, BM), A + (size_t)(ok ? grow : 0) * KK + ch * KC + 4 * c, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // W chunk of the NEXT step lives in registers: its L2 latency hides under this step's MMAs instead of stalling
  // all eight warps at the top of every step (ncu: long-scoreboard bound, tensor pipe 23 % busy without it)
  constexpr int WPT = NN * (KC / 4) / T;
  float4 wr[WPT];
  auto load_w = [&](unsigned q) {
    const unsigned ch = q % NCHUNK;
#pragma unroll
    for (int it = 0; it < WPT; ++it) {
      const int idx = it * T + tid;
      wr[it] = ldg4(W + (size_t)(idx / (KC / 4)) * KK + ch * KC + 4 * (idx % (KC / 4)));
    }
  };
  if (total > 0) {
    issue_a(0);
    load_w(0);
  }
  if (total > 1) issue_a(1);
  if (gridDim.x >= ntiles) pdl_launch_dependents();

  TCP_DECL;
  uint32_t ph[2] = {0, 0};                                                   // parity of the next completion of bar[s]
  bool pending[2] = {false, false};                                          // MMAs committed to bar[s] and not yet waited for
  auto ensure = [&](uint32_t x) {
    if (pending[x]) {
      mbar_wait(bar + x, ph[x]);
      ph[x] ^= 1u;
      pending[x] = false;
    }
  };
  for (unsigned q = 0; q < total; ++q) {
    const unsigned tile = blockIdx.x + (q / NCHUNK) * gridDim.x, ch = q % NCHUNK, s = q & 1u;
    const uint32_t a_raw = base + (q % 3) * A_CH, a_lo = base + OFF_ALO + s * A_CH, b_hi = base + OFF_BHI + s * B_CH,
                   b_lo = base + OFF_BLO + s * B_CH;
    // stage s was last read by the MMAs of step q - 2: they were waited for at step q - 1 (below)
    if (q + 1 < total) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    TCP(0);
    // ---- B chunk: W[n][ch*32 .. +32) -> hi / lo, K-major SWIZZLE_128B (values prefetched one step ahead) ----
#pragma unroll
    for (int it = 0; it < WPT; ++it) {
      const int idx = it * T + tid;
      const uint32_t n = idx / (KC / 4), c = idx % (KC / 4);
      const float4 w = wr[it];
      float4 lo;                                       // B_hi = the raw values (the tensor core truncates them to TF32)
      lo.x = lo_tf32(w.x); lo.y = lo_tf32(w.y); lo.z = lo_tf32(w.z); lo.w = lo_tf32(w.w);
      const uint32_t off = swz_off(n, 4 * c, NN);
      *reinterpret_cast<float4*>(sm + OFF_BHI + s * B_CH + off) = w;
      *reinterpret_cast<float4*>(sm + OFF_BLO + s * B_CH + off) = lo;
    }
    // ---- A_lo of the chunks this thread loaded ----
#pragma unroll
    for (int it = 0; it < BM * (KC / 4) / T; ++it) {
      const int idx = it * T + tid;
      const uint32_t row = idx / (KC / 4), c = idx % (KC / 4);
      const uint32_t off = swz_off(row, 4 * c, BM);
      const float4 x = *reinterpret_cast<const float4*>(sm + (q % 3) * A_CH + off);
      float4 lo;
      lo.x = lo_tf32(x.x); lo.y = lo_tf32(x.y); lo.z = lo_tf32(x.z); lo.w = lo_tf32(x.w);
      *reinterpret_cast<float4*>(sm + OFF_ALO + s * A_CH + off) = lo;
    }
    TCP(1);
    fence_proxy_async();
    __syncthreads();
    TCP(2);
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < KC / 8; ++ks) {
        const uint32_t ko = (uint32_t)ks * 32u;
        umma_tf32(tmem, umma_desc_k128(a_raw + ko), umma_desc_k128(b_hi + ko), IDESC, (ch > 0 || ks > 0) ? 1u : 0u);
        umma_tf32(tmem, umma_desc_k128(a_lo + ko), umma_desc_k128(b_hi + ko), IDESC, 1);
        umma_tf32(tmem, umma_desc_k128(a_raw + ko), umma_desc_k128(b_lo + ko), IDESC, 1);
      }
      umma_commit(bar + s);
    }
    pending[s] = true;
    if (q + 1 < total) load_w(q + 1);
    TCP(3);
    // the MMAs of step q - 1 (other stage, ring slot (q-1) % 3 = (q+2) % 3) must be done before that ring slot is refilled
    ensure(s ^ 1u);
    TCP(4);
    if (q + 2 < total) issue_a(q + 2);
    TCP(5);

    if (ch == NCHUNK - 1) {
      // ---- tile complete: wait for this step's MMAs too, then the epilogue ----
      ensure(s);
      TCP(6);
      tc_fence_after();
      const int lane_row = tid & 127, half = tid >> 7;                      // TMEM lane (= row of the tile), column half
      const unsigned row = tile * BM + lane_row;
      constexpr int NCH = NN / 4;
      unsigned char* stg = sm + OFF_ALO;                                     // [128][NN] floats over A_lo / B buffers
      float ps = 0.f, pd = 0.f;
#pragma unroll
      for (int cb = 0; cb < NN / 64; ++cb) {
        const int col0 = half * (NN / 2) + cb * 32;
        float v[32];
        tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + col0, v);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          ps = fmaf(v[k], att[col0 + k], ps);
          pd = fmaf(v[k], att[NN + col0 + k], pd);
        }
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const int c = col0 / 4 + k / 4;
          *reinterpret_cast<float4*>(stg + (size_t)lane_row * (NN * 4) + (((c ^ lane_row) & (NCH - 1)) << 4)) =
              make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
        }
      }
      part[tid] = ps;
      part[T + tid] = pd;
      tc_fence_before();
      __syncthreads();
      if (tid < BM && row < M) {
        if (H == 2) {                                                         // column half == head
          s0[(size_t)row * 2 + 0] = part[tid];       s0[(size_t)row * 2 + 1] = part[tid + 128];
          s1[(size_t)row * 2 + 0] = part[T + tid];   s1[(size_t)row * 2 + 1] = part[T + tid + 128];
        } else {
          s0[row] = part[tid] + part[tid + 128];
          s1[row] = part[T + tid] + part[T + tid + 128];
        }
      }
#pragma unroll 4
      for (int it = 0; it < BM * NCH / T; ++it) {
        const int idx = it * T + tid;
        const int r = idx / NCH, c = idx % NCH;
        const unsigned grow = tile * BM + r;
        if (grow < M)
          st4(Cout + (size_t)grow * NN + 4 * c,
              *reinterpret_cast<const float4*>(stg + (size_t)r * (NN * 4) + (((c ^ r) & (NCH - 1)) << 4)));
      }
      __syncthreads();                                                       // staging area (= stage buffers) free again
      TCP(7);
    }
  }
#ifdef GATRES_TC_PROF
  if (blockIdx.x == 0 && (tid == 0 || tid == 200))
    printf("tc_wide<%d,%d> tid %d steps %u: wait_a %lld stage %lld sync %lld mma_issue+loadw %lld ensure_prev %lld issue_a %lld ensure_tile %lld epilogue %lld (cycles per step)\n",
           KK, NN, tid, total, tcp_t[0] / total, tcp_t[1] / total, tcp_t[2] / total, tcp_t[3] / total, tcp_t[4] / total, tcp_t[5] / total,
           tcp_t[6] / total, tcp_t[7] / total);
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, NN);
}

template <int KK, int NN, int H>
static int launch_tc_wide(const float* A, const float* W, const float* e0, const float* e1, float* Cout, float* s0,
                          float* s1, unsigned M, cudaStream_t st) {
  constexpr size_t smem = 5 * (size_t)128 * 32 * 4 + 4 * (size_t)NN * 32 * 4 + 2 * NN * 4 + 2 * 256 * 4 + 32;
  auto kern = gemm_tc_wide_kernel<KK, NN, H>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("gemm_tc_wide");
    configured = true;
  }
  const unsigned ntiles = (M + 127) / 128;
  unsigned per_sm = (unsigned)((228u * 1024u) / (smem + 1024u));
  per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
  if (per_sm * NN > 512) per_sm = 512 / NN;                                 // TMEM columns per SM
  unsigned grid = (unsigned)sm_count() * per_sm;
  if (grid > ntiles) grid = ntiles;
  launch_kernel(kern, dim3(grid), dim3(256), smem, st, A, W, e0, e1, Cout, s0, s1, M);
  return check_launch("gemm_tc_wide");
}

// warp-specialised form of the wide projections (linear_tc_wide.cu); 0 = not covered / switched off
int gemm_tc_wide2_dispatch(int H, int KK, int NN, const float* A, const float* W, const float* e0, const float* e1,
                           float* Cout, float* s0, float* s1, unsigned M, cudaStream_t st);

// -> 1 if handled, 0 if this shape has no tensor-core path (caller falls back to the FFMA kernel), <0 on error
int gemm_tc_dispatch(int mode, int H, int KK, int NN, const float* A, const float* W, const float* e0,
                     const float* e1, float* Cout, float* s0, float* s1, unsigned M, cudaStream_t st) {
  int rc = 1;
  if (mode == 0) {                              // warp-specialised form (linear_tc_wide.cu): wide shapes and, by default, nc = 32
    rc = gemm_tc_wide2_dispatch(H, KK, NN, A, W, e0, e1, Cout, s0, s1, M, st);
    if (rc != 0) return rc;
  }
#define TC(KKv, NNv, MODEv, Hv)                                                                           \
  if (mode == MODEv && KK == KKv && NN == NNv && (MODEv == 1 || H == Hv)) {                               \
    rc = launch_tc<KKv, NNv, MODEv, Hv>(A, W, e0, e1, Cout, s0, s1, M, st, "gemm_tc");                    \
    return rc ? rc : 1;                                                                                   \
  }
  TC(32, 64, 0, 2)      // conv1 projection, nc = 32
  TC(64, 32, 0, 1)      // conv2 projection, nc = 32
  TC(32, 64, 1, 1)      // conv2 data gradient (dh2 [M,32] -> dy1 [M,64])
  TC(64, 32, 1, 1)      // conv1 data gradient (dh1 [M,64] -> dx0 [M,32])
#undef TC
#define TCW(KKv, NNv, Hv)                                                                                  \
  if (mode == 0 && KK == KKv && NN == NNv && H == Hv) {                                                    \
    rc = launch_tc_wide<KKv, NNv, Hv>(A, W, e0, e1, Cout, s0, s1, M, st);                                   \
    return rc ? rc : 1;                                                                                   \
  }
  TCW(64, 128, 2)       // conv1 projection, nc = 64 (its conv2, K = 128 -> N = 64, stays on the FFMA kernel)
  TCW(128, 256, 2)      // conv1 projection, nc = 128
  TCW(256, 128, 1)      // conv2 projection, nc = 128
#undef TCW
  return 0;
}

}  // namespace gatres
