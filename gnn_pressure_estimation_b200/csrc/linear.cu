// Node-wise dense contractions of the GATRes path (sm_100a): the GATConv
// projection with the attention-score epilogue, its data/weight gradients, and
// the Linear(1,nc) encoder / Linear(nc,1) decoder
// (/root/reference/gnn_pressure_estimation/GraphModels.py:458-459, :477, :484;
// SURVEY.md §A.2 step 1, §A.4 last two lines).
//
// fp32 FFMA on purpose: the parity contract is 1e-4 relative through 30 chained
// projections, and these GEMMs are skinny (K <= 256, ~10 flop/B) and
// HBM-bound, so the kernels are organised around the memory system instead:
// persistent CTAs keep the whole weight matrix resident in shared memory and
// stream row tiles through a cp.async double buffer.
#include <stdlib.h>
#include "common.cuh"

namespace gatres {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float f4get(const float4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

// ---------------------------------------------------------------------------
// C[M,NN] = A[M,KK] * Bm[KK,NN]  over persistent CTAs.
//   MODE 0 (projection fwd): Bm[k][n] = W[n][k] (W is [NN,KK]); epilogue also
//          emits s_src/s_dst[m,h] = <C[m,h,:], att[h,:]>.
//   MODE 1 (data gradient) : Bm = W as stored ([KK,NN]); epilogue adds `e0`
//          (residual gradient) and masks by `e1 > 0` (ReLU of the layer input).
// Thread (tx,ty) owns rows {ty + i*TY} and, per 4-wide column group jv, columns
// jv*(NN/NV) + 4*tx .. +3, so every shared/global access of a warp is contiguous.
// ---------------------------------------------------------------------------
template <int KK, int NN, int BM, int TM, int TN, int STAGES, int MODE, int H>
__global__ void __launch_bounds__(256)
gemm_rows_kernel(const float* __restrict__ A, const float* __restrict__ W,
                 const float* __restrict__ e0, const float* __restrict__ e1,
                 float* __restrict__ Cout, float* __restrict__ s0, float* __restrict__ s1, unsigned M) {
  constexpr int TX = NN / TN, TY = BM / TM, NV = TN / 4, LDA = KK + 4, LDB = NN + 4, CG = NN / NV;
  static_assert(TX * TY == 256 && TN % 4 == 0 && TX <= 32 && (32 % TX) == 0, "bad tiling");
  static_assert(MODE == 1 || (H == 1 || NV == 1 || NV == 2), "score epilogue layout");
  extern __shared__ __align__(16) float smem[];
  float* Bs = smem;                       // [KK][LDB]
  float* As = smem + KK * LDB;            // [STAGES][BM][LDA]
  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
  const unsigned ntiles = (M + BM - 1) / BM;

  auto load_tile = [&](unsigned tile, int stage) {
    float* dst = As + stage * BM * LDA;
    for (int idx = tid; idx < BM * (KK / 4); idx += 256) {
      const int m = idx / (KK / 4), kv = idx % (KK / 4);
      const unsigned row = tile * BM + m;
      const bool ok = row < M;
      cp_async16(dst + m * LDA + 4 * kv, A + (size_t)(ok ? row : 0) * KK + 4 * kv, ok ? 16 : 0);
    }
  };

  // weights / attention vectors are constant within a step: stage them before the dependency wait
  if (MODE == 0) {
    for (int idx = tid; idx < NN * (KK / 4); idx += 256) {
      const int n = idx / (KK / 4), kv = idx % (KK / 4);
      const float4 w = ldg4(W + (size_t)n * KK + 4 * kv);
      Bs[(4 * kv + 0) * LDB + n] = w.x;
      Bs[(4 * kv + 1) * LDB + n] = w.y;
      Bs[(4 * kv + 2) * LDB + n] = w.z;
      Bs[(4 * kv + 3) * LDB + n] = w.w;
    }
  } else {
    for (int idx = tid; idx < KK * (NN / 4); idx += 256) {
      const int k = idx / (NN / 4), nv = idx % (NN / 4);
      st4(Bs + k * LDB + 4 * nv, ldg4(W + (size_t)k * NN + 4 * nv));
    }
  }
  float att_s[TN], att_d[TN];
  if (MODE == 0) {
#pragma unroll
    for (int jv = 0; jv < NV; ++jv) {
      const float4 a = ldg4(e0 + jv * CG + 4 * tx), d = ldg4(e1 + jv * CG + 4 * tx);
      att_s[jv * 4 + 0] = a.x; att_s[jv * 4 + 1] = a.y; att_s[jv * 4 + 2] = a.z; att_s[jv * 4 + 3] = a.w;
      att_d[jv * 4 + 0] = d.x; att_d[jv * 4 + 1] = d.y; att_d[jv * 4 + 2] = d.z; att_d[jv * 4 + 3] = d.w;
    }
  }
  pdl_wait();
  unsigned tile = blockIdx.x;
  if (tile < ntiles) load_tile(tile, 0);
  cp_async_commit();

  int stage = 0;
  if (gridDim.x >= ntiles) pdl_launch_dependents();
  for (; tile < ntiles; tile += gridDim.x) {
    const unsigned next = tile + gridDim.x;
    if (STAGES == 2) {
      if (next < ntiles) load_tile(next, stage ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    const float* as = As + stage * BM * LDA;
#pragma unroll 2
    for (int k4 = 0; k4 < KK / 4; ++k4) {
      float4 a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(as + (ty + i * TY) * LDA + 4 * k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float b[TN];
#pragma unroll
        for (int jv = 0; jv < NV; ++jv) {
          const float4 t = *reinterpret_cast<const float4*>(Bs + (4 * k4 + kk) * LDB + jv * CG + 4 * tx);
          b[jv * 4 + 0] = t.x; b[jv * 4 + 1] = t.y; b[jv * 4 + 2] = t.z; b[jv * 4 + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const float ai = f4get(a[i], kk);
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ai, b[j], acc[i][j]);
        }
      }
    }

    // ---- epilogue (registers -> global) ----
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const unsigned row = tile * BM + ty + i * TY;
      const bool ok = row < M;
      if (MODE == 0) {
        float ps[NV], pd[NV];
#pragma unroll
        for (int jv = 0; jv < NV; ++jv) {
          ps[jv] = pd[jv] = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            ps[jv] = fmaf(acc[i][jv * 4 + c], att_s[jv * 4 + c], ps[jv]);
            pd[jv] = fmaf(acc[i][jv * 4 + c], att_d[jv * 4 + c], pd[jv]);
          }
        }
        if (H == 1) {
          float a = ps[0], d = pd[0];
#pragma unroll
          for (int jv = 1; jv < NV; ++jv) { a += ps[jv]; d += pd[jv]; }
          a = group_sum<TX>(a, 0xffffffffu);
          d = group_sum<TX>(d, 0xffffffffu);
          if (ok && tx == 0) { s0[row] = a; s1[row] = d; }
        } else if (NV == 2) {          // column group jv == head jv
#pragma unroll
          for (int jv = 0; jv < NV; ++jv) {
            const float a = group_sum<TX>(ps[jv], 0xffffffffu), d = group_sum<TX>(pd[jv], 0xffffffffu);
            if (ok && tx == 0) { s0[(size_t)row * 2 + jv] = a; s1[(size_t)row * 2 + jv] = d; }
          }
        } else {                       // NV == 1: heads split the tx range in halves
          const float a = group_sum<TX / 2>(ps[0], 0xffffffffu), d = group_sum<TX / 2>(pd[0], 0xffffffffu);
          if (ok && (tx % (TX / 2)) == 0) {
            s0[(size_t)row * 2 + tx / (TX / 2)] = a;
            s1[(size_t)row * 2 + tx / (TX / 2)] = d;
          }
        }
      }
      if (ok) {
#pragma unroll
        for (int jv = 0; jv < NV; ++jv) {
          const size_t o = (size_t)row * NN + jv * CG + 4 * tx;
          float4 v = make_float4(acc[i][jv * 4 + 0], acc[i][jv * 4 + 1], acc[i][jv * 4 + 2], acc[i][jv * 4 + 3]);
          if (MODE == 1) {
            if (e0 != nullptr) add4(v, ldg4_stream(e0 + o));
            if (e1 != nullptr) {
              const float4 r = ldg4_stream(e1 + o);
              v.x = r.x > 0.f ? v.x : 0.f; v.y = r.y > 0.f ? v.y : 0.f;
              v.z = r.z > 0.f ? v.z : 0.f; v.w = r.w > 0.f ? v.w : 0.f;
            }
          }
          st4(Cout + o, v);
        }
      }
    }
    __syncthreads();                    // everyone is done with As[stage]
    if (STAGES == 1) {
      if (next < ntiles) load_tile(next, 0);
      cp_async_commit();
    } else {
      stage ^= 1;
    }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------
// Weight gradient dW[NO,KI] = dh^T x, reduced over this CTA's row tiles and
// emitted once (row of `partial`, or atomic add into the gradient buffer).
// The 256 threads form G row groups; within a group thread (tk,tn) owns a
// TNn x TKk block of dW and group g takes rows g, g+G, ... of every tile, so a
// thread does TNn*TKk FFMAs per (TNn+TKk)/4 shared-memory loads (32 : 3 for
// nc = 32) instead of 8 : 1.5; the G partial copies are summed through shared
// memory at the end.
// ---------------------------------------------------------------------------
template <int NO, int KI, int BM, int TNn, int TKk, int G>
__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ dh, const float* __restrict__ x, float* __restrict__ partial,
             long long P, long long off_W, unsigned M, int atomic) {
  constexpr int TKT = KI / TKk, TNT = NO / TNn, GT = 256 / G, LDH = NO + 4, LDX = KI + 4;
  constexpr int NVn = TNn / 4, CGn = NO / NVn, NVk = TKk / 4, CGk = KI / NVk;
  static_assert(TKT * TNT == GT && TNn % 4 == 0 && TKk % 4 == 0 && BM % G == 0, "bad wgrad tiling");
  extern __shared__ __align__(16) float smem[];
  float* Hs = smem;                         // [2][BM][LDH]
  float* Xs = smem + 2 * BM * LDH;          // [2][BM][LDX]
  const int tid = threadIdx.x, grp = tid / GT, t = tid % GT, tk = t % TKT, tn = t / TKT;
  const unsigned ntiles = (M + BM - 1) / BM;

  auto load_tile = [&](unsigned tile, int stage) {
    float* hd = Hs + stage * BM * LDH;
    float* xd = Xs + stage * BM * LDX;
    for (int idx = tid; idx < BM * (NO / 4); idx += 256) {
      const int m = idx / (NO / 4), v = idx % (NO / 4);
      const unsigned row = tile * BM + m;
      const bool ok = row < M;
      cp_async16(hd + m * LDH + 4 * v, dh + (size_t)(ok ? row : 0) * NO + 4 * v, ok ? 16 : 0);
    }
    for (int idx = tid; idx < BM * (KI / 4); idx += 256) {
      const int m = idx / (KI / 4), v = idx % (KI / 4);
      const unsigned row = tile * BM + m;
      const bool ok = row < M;
      cp_async16(xd + m * LDX + 4 * v, x + (size_t)(ok ? row : 0) * KI + 4 * v, ok ? 16 : 0);
    }
  };

  float acc[TNn][TKk];
#pragma unroll
  for (int i = 0; i < TNn; ++i)
#pragma unroll
    for (int j = 0; j < TKk; ++j) acc[i][j] = 0.f;

  pdl_wait();
  unsigned tile = blockIdx.x;
  if (tile < ntiles) load_tile(tile, 0);
  cp_async_commit();
  int stage = 0;
  if (gridDim.x >= ntiles) pdl_launch_dependents();
  for (; tile < ntiles; tile += gridDim.x) {
    const unsigned next = tile + gridDim.x;
    if (next < ntiles) load_tile(next, stage ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* hs = Hs + stage * BM * LDH;
    const float* xs = Xs + stage * BM * LDX;
#pragma unroll 2
    for (int m = grp; m < BM; m += G) {       // rows past M were zero-filled
      float a[TNn], b[TKk];
#pragma unroll
      for (int jv = 0; jv < NVn; ++jv) {
        const float4 q = *reinterpret_cast<const float4*>(hs + m * LDH + jv * CGn + 4 * tn);
        a[jv * 4 + 0] = q.x; a[jv * 4 + 1] = q.y; a[jv * 4 + 2] = q.z; a[jv * 4 + 3] = q.w;
      }
#pragma unroll
      for (int jv = 0; jv < NVk; ++jv) {
        const float4 q = *reinterpret_cast<const float4*>(xs + m * LDX + jv * CGk + 4 * tk);
        b[jv * 4 + 0] = q.x; b[jv * 4 + 1] = q.y; b[jv * 4 + 2] = q.z; b[jv * 4 + 3] = q.w;
      }
#pragma unroll
      for (int i = 0; i < TNn; ++i)
#pragma unroll
        for (int j = 0; j < TKk; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    stage ^= 1;
  }
  cp_async_wait<0>();

  // ---- sum the G row-group copies through shared memory, then emit ----
  float* red = smem;                         // [G-1][GT][TNn*TKk], aliases the (now idle) tile buffers
  if (G > 1) {
    __syncthreads();
    if (grp > 0) {
      float* dst = red + ((size_t)(grp - 1) * GT + t) * (TNn * TKk);
#pragma unroll
      for (int i = 0; i < TNn; ++i)
#pragma unroll
        for (int j = 0; j < TKk; j += 4)
          st4(dst + i * TKk + j, make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]));
    }
    __syncthreads();
  }
  if (grp == 0) {
    float* out = partial + (atomic ? 0 : (size_t)blockIdx.x * P) + off_W;
#pragma unroll
    for (int i = 0; i < TNn; ++i) {
      const int n = (i / 4) * CGn + 4 * tn + (i % 4);
#pragma unroll
      for (int jv = 0; jv < NVk; ++jv) {
        float4 v = make_float4(acc[i][jv * 4 + 0], acc[i][jv * 4 + 1], acc[i][jv * 4 + 2], acc[i][jv * 4 + 3]);
#pragma unroll
        for (int g2 = 1; g2 < G; ++g2)
          add4(v, *reinterpret_cast<const float4*>(red + ((size_t)(g2 - 1) * GT + t) * (TNn * TKk) + i * TKk + jv * 4));
        emit4(out + (size_t)n * KI + jv * CGk + 4 * tk, v, atomic);
      }
    }
  }
}


// ---------------------------------------------------------------------------
// Tensor-core weight gradient for the nc = 32 shapes (atomic accumulation mode): the same persistent, cp.async
// double-buffered tiles as wgrad_kernel, contracted with mma.sync.m16n8k8 TF32 in the error-compensated 3xTF32 form
// (hi = the tensor core's truncation of the fp32 operand, lo = rna_tf32(x - trunc x); three independent accumulator
// chains).  dW[NO][KI] is NO/16 x KI/8 output tiles of 16 x 8; the 8 warps own TILES/8 tiles each, a warp's tiles
// share one A fragment (dh^T) per 8-row step.  The FFMA kernel above is issue bound (2048 FMA per row = 64 warp
// instructions + loads, ncu: issue 70 %, DRAM 38 %); this form needs ~38 per row.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned wg_lo_bits(float x) {
  const float r = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  return (__float_as_uint(r) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ void wg_mma(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NO, int KI, int BM>
__global__ void __launch_bounds__(256)
wgrad_mma_kernel(const float* __restrict__ dh, const float* __restrict__ x, float* __restrict__ grads, long long off_W,
                 unsigned M) {
  constexpr int LDH = NO + 4, LDX = KI + 4, NT = KI / 8, TILES = (NO / 16) * NT, TPW = TILES / 8;
  static_assert(TILES % 8 == 0 && NT % TPW == 0 && BM % 8 == 0, "wgrad_mma tiling: a warp's tiles share one 16-row A block");
  extern __shared__ __align__(16) float smem[];
  float* Hs = smem;                         // [2][BM][LDH]
  float* Xs = smem + 2 * BM * LDH;          // [2][BM][LDX]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int tile0 = warp * TPW, no0 = (tile0 / NT) * 16, ki0 = (tile0 % NT) * 8;
  const unsigned ntiles = (M + BM - 1) / BM;

  auto load_tile = [&](unsigned tile, int stage) {
    float* hd = Hs + stage * BM * LDH;
    float* xd = Xs + stage * BM * LDX;
    for (int idx = tid; idx < BM * (NO / 4); idx += 256) {
      const int m = idx / (NO / 4), v = idx % (NO / 4);
      const unsigned row = tile * BM + m;
      const bool ok = row < M;
      cp_async16(hd + m * LDH + 4 * v, dh + (size_t)(ok ? row : 0) * NO + 4 * v, ok ? 16 : 0);
    }
    for (int idx = tid; idx < BM * (KI / 4); idx += 256) {
      const int m = idx / (KI / 4), v = idx % (KI / 4);
      const unsigned row = tile * BM + m;
      const bool ok = row < M;
      cp_async16(xd + m * LDX + 4 * v, x + (size_t)(ok ? row : 0) * KI + 4 * v, ok ? 16 : 0);
    }
  };

  float acc[TPW][4], acl[TPW][4], acm[TPW][4];
#pragma unroll
  for (int q = 0; q < TPW; ++q)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[q][i] = acl[q][i] = acm[q][i] = 0.f;

  pdl_wait();
  unsigned tile = blockIdx.x;
  if (tile < ntiles) load_tile(tile, 0);
  cp_async_commit();
  int stage = 0;
  if (gridDim.x >= ntiles) pdl_launch_dependents();
  for (; tile < ntiles; tile += gridDim.x) {
    const unsigned next = tile + gridDim.x;
    if (next < ntiles) load_tile(next, stage ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* hs = Hs + stage * BM * LDH;
    const float* xs = Xs + stage * BM * LDX;
#pragma unroll
    for (int m0 = 0; m0 < BM; m0 += 8) {      // rows past M were zero-filled
      // A = dh^T block [16 output rows x 8 data rows]: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4)
      const float* h0 = hs + (m0 + t) * LDH + no0 + g;
      const float* h1 = hs + (m0 + t + 4) * LDH + no0 + g;
      const float av[4] = {h0[0], h0[8], h1[0], h1[8]};
      unsigned a[4], al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = __float_as_uint(av[i]); al[i] = wg_lo_bits(av[i]); }
#pragma unroll
      for (int q = 0; q < TPW; ++q) {
        const float b0 = xs[(m0 + t) * LDX + ki0 + 8 * q + g], b1 = xs[(m0 + t + 4) * LDX + ki0 + 8 * q + g];
        const unsigned b[2] = {__float_as_uint(b0), __float_as_uint(b1)}, bl[2] = {wg_lo_bits(b0), wg_lo_bits(b1)};
        wg_mma(acl[q], al, b);
        wg_mma(acm[q], a, bl);
        wg_mma(acc[q], a, b);
      }
    }
    __syncthreads();
    stage ^= 1;
  }
  cp_async_wait<0>();
  float* out = grads + off_W;
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    const int k = ki0 + 8 * q + 2 * t;
    atomicAdd(reinterpret_cast<float2*>(out + (size_t)(no0 + g) * KI + k),
              make_float2(acc[q][0] + acl[q][0] + acm[q][0], acc[q][1] + acl[q][1] + acm[q][1]));
    atomicAdd(reinterpret_cast<float2*>(out + (size_t)(no0 + g + 8) * KI + k),
              make_float2(acc[q][2] + acl[q][2] + acm[q][2], acc[q][3] + acl[q][3] + acm[q][3]));
  }
}

static unsigned wgrad_mma_ctas_per_sm() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GATRES_WGRAD_CTAS");
    v = e ? atoi(e) : 6;
    if (v < 1 || v > 8) v = 6;
  }
  return (unsigned)v;
}

template <int NO, int KI>
static int launch_wgrad_mma(const float* dh, const float* x, float* grads, long long off_W, unsigned M, cudaStream_t st) {
  constexpr int BM = 32;
  constexpr size_t smem = (size_t)(2 * BM * (NO + 4) + 2 * BM * (KI + 4)) * sizeof(float);
  auto kern = wgrad_mma_kernel<NO, KI, BM>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("wgrad_mma");
    configured = true;
  }
  const unsigned ntiles = (M + BM - 1) / BM;
  unsigned grid = (unsigned)sm_count() * wgrad_mma_ctas_per_sm();
  if (grid > ntiles) grid = ntiles;
  launch_kernel(kern, dim3(grid), dim3(256), smem, st, dh, x, grads, off_W, M);
  return check_launch("wgrad_mma");
}

// ------------------------------------------------------------ host launchers
template <int KK, int NN, int BM, int TM, int TN, int STAGES, int MODE, int H>
static int launch_gemm(const float* A, const float* W, const float* e0, const float* e1, float* Cout, float* s0,
                       float* s1, unsigned M, cudaStream_t st, const char* what) {
  constexpr size_t smem = (size_t)(KK * (NN + 4) + STAGES * BM * (KK + 4)) * sizeof(float);
  auto kern = gemm_rows_kernel<KK, NN, BM, TM, TN, STAGES, MODE, H>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch(what);
    configured = true;
  }
  const unsigned ntiles = (M + BM - 1) / BM;
  unsigned per_sm = (unsigned)((200 * 1024) / smem);
  per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
  unsigned grid = (unsigned)sm_count() * per_sm;
  if (grid > ntiles) grid = ntiles;
  launch_kernel(kern, dim3(grid), dim3(256), smem, st, A, W, e0, e1, Cout, s0, s1, M);
  return check_launch(what);
}

template <int NO, int KI, int TNn, int TKk, int G>
static int launch_wgrad(const float* dh, const float* x, float* partial, long long P, int slots, long long off_W,
                        unsigned M, cudaStream_t st) {
  constexpr int BM = 32;
  constexpr size_t tiles = (size_t)(2 * BM * (NO + 4) + 2 * BM * (KI + 4)) * sizeof(float);
  constexpr size_t red = (size_t)(G - 1) * NO * KI * sizeof(float);
  constexpr size_t smem = tiles > red ? tiles : red;
  auto kern = wgrad_kernel<NO, KI, BM, TNn, TKk, G>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch("wgrad");
    configured = true;
  }
  const int atomic = slots <= 0;
  const unsigned ntiles = (M + BM - 1) / BM;
  unsigned grid = (unsigned)slots;
  if (atomic) {
    grid = (unsigned)sm_count() * 2u;
    if (grid > ntiles) grid = ntiles;
  }
  launch_kernel(kern, dim3(grid), dim3(256), smem, st, dh, x, partial, P, off_W, M, atomic);
  return check_launch("wgrad");
}

// (K, H*C) -> tiling; the same six shapes serve forward (KK=K, NN=H*C) and the
// data gradient (KK=H*C, NN=K).
template <int MODE, int H>
static int dispatch_gemm(int KK, int NN, const float* A, const float* W, const float* e0, const float* e1,
                         float* Cout, float* s0, float* s1, unsigned M, cudaStream_t st, const char* what) {
#define G(KKv, NNv, BM, TM, TN, ST) \
  if (KK == KKv && NN == NNv) return launch_gemm<KKv, NNv, BM, TM, TN, ST, MODE, H>(A, W, e0, e1, Cout, s0, s1, M, st, what)
  G(32, 64, 64, 4, 4, 2);
  G(64, 32, 128, 4, 4, 2);
  G(64, 64, 64, 4, 4, 2);       // sibling GAT model (GraphModels.py:210-230): GATConv(2*32 -> 2 heads x 32)
  G(64, 128, 64, 4, 8, 2);
  G(128, 64, 64, 4, 4, 2);
  G(128, 256, 64, 8, 8, 2);
  G(256, 128, 64, 4, 8, 1);
#undef G
  set_error("%s: unsupported contraction [M,%d] x [%d,%d]", what, KK, KK, NN);
  return GATRES_ERR_ARG;
}

// ------------------------------------------------- encoder / decoder kernels
__global__ void __launch_bounds__(256)
encoder_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                   float* __restrict__ out, size_t total4, int nc4) {
  pdl_wait();
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total4; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t m = idx / nc4;
    const int c = (int)(idx - m * nc4);
    const float xv = __ldg(x + m);
    const float4 wv = ldg4(w + 4 * c), bv = ldg4(b + 4 * c);
    st4(out + idx * 4, make_float4(fmaf(xv, wv.x, bv.x), fmaf(xv, wv.y, bv.y), fmaf(xv, wv.z, bv.z), fmaf(xv, wv.w, bv.w)));
  }
}

// column sums over rows of g (db) and of g * x[m] (dw); thread owns chunk tid % nc4
__global__ void __launch_bounds__(256)
encoder_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, float* __restrict__ partial,
                   long long P, long long off_w, long long off_b, unsigned M, int nc4, int atomic) {
  __shared__ float red[2 * 256 * 4];
  const int c = threadIdx.x % nc4, rl = threadIdx.x / nc4, rows_per_pass = 256 / nc4;
  float4 aw = f4zero(), ab = f4zero();
  pdl_wait();
  for (size_t m = (size_t)blockIdx.x * rows_per_pass + rl; m < M; m += (size_t)gridDim.x * rows_per_pass) {
    const float4 gv = ldg4_stream(g + m * (size_t)(nc4 * 4) + 4 * c);
    fma4(aw, __ldg(x + m), gv);
    add4(ab, gv);
  }
  st4(red + threadIdx.x * 4, aw);
  st4(red + (256 + threadIdx.x) * 4, ab);
  __syncthreads();
  if (threadIdx.x < nc4) {
    float4 sw = f4zero(), sb = f4zero();
    for (int k = 0; k < rows_per_pass; ++k) {
      add4(sw, *reinterpret_cast<float4*>(red + (k * nc4 + threadIdx.x) * 4));
      add4(sb, *reinterpret_cast<float4*>(red + (256 + k * nc4 + threadIdx.x) * 4));
    }
    float* row = partial + (atomic ? 0 : (size_t)blockIdx.x * P);
    emit4(row + off_w + 4 * threadIdx.x, sw, atomic);
    emit4(row + off_b + 4 * threadIdx.x, sb, atomic);
  }
}

template <int NC>
__global__ void __launch_bounds__(256)
decoder_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                   float* __restrict__ out, const int* __restrict__ poison, unsigned M) {
  constexpr int LPR = NC / 4, RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / LPR, lig = lane % LPR;
  const float4 wv = ldg4(w + 4 * lig);
  const float bias = __ldg(b);
  pdl_wait();
  const bool bad = poison != nullptr && __ldg(poison) != 0;
  constexpr unsigned rows_per_cta = kWarps * RPW;
  if ((unsigned long long)gridDim.x * rows_per_cta >= M) pdl_launch_dependents();   // single pass: see common.cuh
  for (unsigned r0 = blockIdx.x * rows_per_cta; r0 < M; r0 += gridDim.x * rows_per_cta) {
    const unsigned r = r0 + warp * RPW + sub;
    float p = 0.f;
    if (r < M) p = dot4(ldg4_stream(x + (size_t)r * NC + 4 * lig), wv);
    p = group_sum<LPR>(p, 0xffffffffu);
    if (r < M && lig == 0) out[r] = bad ? __int_as_float(0x7fc00000) : p + bias;
  }
}

// dx[m,c] = g[m] w[c] (masked by x>0 when mask_relu), dw[c] = sum_m g[m] x[m,c], db = sum_m g[m]
template <int NC>
__global__ void __launch_bounds__(256)
decoder_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ w,
                   float* __restrict__ dx, float* __restrict__ partial, long long P, long long off_w,
                   long long off_b, unsigned M, int mask_relu, int atomic) {
  constexpr int LPR = NC / 4, RPW = 32 / LPR;
  __shared__ float red[kWarps * 32 * 4];
  __shared__ float redb[kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / LPR, lig = lane % LPR;
  const float4 wv = ldg4(w + 4 * lig);
  float4 aw = f4zero();
  float ab = 0.f;
  pdl_wait();

  constexpr unsigned rows_per_cta = kWarps * RPW;
  if ((unsigned long long)gridDim.x * rows_per_cta >= M) pdl_launch_dependents();   // single pass: see common.cuh
  for (unsigned r0 = blockIdx.x * rows_per_cta; r0 < M; r0 += gridDim.x * rows_per_cta) {
    const unsigned r = r0 + warp * RPW + sub;
    if (r >= M) continue;
    const float gv = __ldg(g + r);
    const float4 xv = ldg4_stream(x + (size_t)r * NC + 4 * lig);
    float4 d = make_float4(gv * wv.x, gv * wv.y, gv * wv.z, gv * wv.w);
    if (mask_relu) {
      d.x = xv.x > 0.f ? d.x : 0.f; d.y = xv.y > 0.f ? d.y : 0.f;
      d.z = xv.z > 0.f ? d.z : 0.f; d.w = xv.w > 0.f ? d.w : 0.f;
    }
    st4(dx + (size_t)r * NC + 4 * lig, d);
    fma4(aw, gv, xv);
    if (lig == 0) ab += gv;
  }
  float* row = partial + (atomic ? 0 : (size_t)blockIdx.x * P);
  cta_chunk_sum_store<LPR>(aw, red, row + off_w, 0, atomic);
  ab = group_sum<32>(ab, 0xffffffffu);
  if (lane == 0) redb[warp] = ab;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < kWarps; ++k) s += redb[k];
    emit1(row + off_b, s, atomic);
  }
}

__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, long long P, int slots, long long p_begin,
                       long long p_end, float* __restrict__ grads) {
  pdl_wait();
  for (long long p = p_begin + blockIdx.x * (long long)blockDim.x + threadIdx.x; p < p_end;
       p += (long long)gridDim.x * blockDim.x) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int s = 0;
    for (; s + 4 <= slots; s += 4) {
      s0 += __ldg(partial + (size_t)(s + 0) * P + p);
      s1 += __ldg(partial + (size_t)(s + 1) * P + p);
      s2 += __ldg(partial + (size_t)(s + 2) * P + p);
      s3 += __ldg(partial + (size_t)(s + 3) * P + p);
    }
    for (; s < slots; ++s) s0 += __ldg(partial + (size_t)s * P + p);
    grads[p] = (s0 + s1) + (s2 + s3);
  }
}

// tensor-core (tcgen05 3xTF32) variants, linear_tc.cu
int linear_bwd_fused_dispatch(int NO, int KI, const float* dh, const float* x, const float* W, const float* add,
                              const float* relu_ref, float* dx, float* grads, long long off_W, unsigned M, cudaStream_t st);
int gemm_tc_dispatch(int mode, int H, int KK, int NN, const float* A, const float* W, const float* e0,
                     const float* e1, float* Cout, float* s0, float* s1, unsigned M, cudaStream_t st);
int wgrad_tc_dispatch(int NO, int KI, const float* dh, const float* x, float* grads, long long off_W, unsigned M, cudaStream_t st);

// 0 = never, 1 = auto (launches of at least kTcMinRows rows: below that the 128-row UMMA tiles leave
// most SMs idle and the FFMA kernel with 64-row tiles is faster), 2 = always
static int g_tensor_core = -1;
constexpr long long kTcMinRows = 32768;
static int tensor_core_mode() {
  if (g_tensor_core < 0) {
    const char* e = getenv("GATRES_TC");
    g_tensor_core = e == nullptr ? 1 : atoi(e);
    if (g_tensor_core < 0 || g_tensor_core > 2) g_tensor_core = 1;
  }
  return g_tensor_core;
}
static bool tensor_core_enabled(long long rows) {
  const int m = tensor_core_mode();
  return m == 2 || (m == 1 && rows >= kTcMinRows);
}

}  // namespace gatres

using namespace gatres;

extern "C" int gatres_set_tensor_core(int mode) {
  const int prev = tensor_core_mode();
  if (mode >= 0 && mode <= 2) g_tensor_core = mode;
  return prev;
}

extern "C" int gatres_linear_att_fwd(const float* x, const float* W, const float* att_src, const float* att_dst,
                                     float* h, float* s_src, float* s_dst, int64_t M, int32_t K, int32_t H,
                                     int32_t C, void* stream) {
  GATRES_REQUIRE(M >= 0 && M < (1ll << 31), "linear_att_fwd: bad M=%lld", (long long)M);
  if (M == 0) return GATRES_OK;
  if (tensor_core_enabled(M) && (H == 1 || H == 2)) {
    const int rc = gemm_tc_dispatch(0, H, K, H * C, x, W, att_src, att_dst, h, s_src, s_dst, (unsigned)M, as_stream(stream));
    if (rc != 0) return rc < 0 ? rc : GATRES_OK;
  }
  if (H == 1) return dispatch_gemm<0, 1>(K, H * C, x, W, att_src, att_dst, h, s_src, s_dst, (unsigned)M, as_stream(stream), "linear_att_fwd");
  if (H == 2) return dispatch_gemm<0, 2>(K, H * C, x, W, att_src, att_dst, h, s_src, s_dst, (unsigned)M, as_stream(stream), "linear_att_fwd");
  set_error("linear_att_fwd: heads must be 1 or 2, got %d", H);
  return GATRES_ERR_ARG;
}

extern "C" int gatres_linear_bwd(const float* dh, const float* x, const float* W, const float* add,
                                 const float* relu_ref, float* dx, float* partial, int64_t P, int32_t slots,
                                 int64_t off_W, int64_t M, int32_t K, int32_t H, int32_t C, void* stream) {
  GATRES_REQUIRE(M > 0 && M < (1ll << 31), "linear_bwd: bad M=%lld", (long long)M);
  GATRES_REQUIRE(P % 4 == 0 && off_W % 4 == 0, "linear_bwd: bad P/off_W");
  cudaStream_t st = as_stream(stream);
  const int NO = H * C;
  if (dx != nullptr && slots <= 0 && tensor_core_enabled(M)) {
    // atomic accumulation mode, large launch, nc = 32 shapes: dx and dW in one pass over dh (linear_tc.cu)
    const int rc = linear_bwd_fused_dispatch(NO, K, dh, x, W, add, relu_ref, dx, partial, off_W, (unsigned)M, st);
    if (rc != 0) return rc < 0 ? rc : GATRES_OK;
  }
  if (dx != nullptr) {
    int rc = tensor_core_enabled(M)
                 ? gemm_tc_dispatch(1, 1, NO, K, dh, W, add, relu_ref, dx, nullptr, nullptr, (unsigned)M, st)
                 : 0;
    if (rc < 0) return rc;
    rc = rc == 1 ? 0 : dispatch_gemm<1, 1>(NO, K, dh, W, add, relu_ref, dx, nullptr, nullptr, (unsigned)M, st, "linear_bwd_dx");
    if (rc) return rc;
  }
  if (slots <= 0 && tensor_core_enabled(M)) {          // atomic accumulation mode, large launch: tensor-core form
    const int tc = wgrad_tc_dispatch(NO, K, dh, x, partial, off_W, (unsigned)M, st);      // tcgen05 (linear_tc_wide.cu)
    if (tc != 0) return tc < 0 ? tc : GATRES_OK;
    if (NO == 64 && K == 32) return launch_wgrad_mma<64, 32>(dh, x, partial, off_W, (unsigned)M, st);
    if (NO == 32 && K == 64) return launch_wgrad_mma<32, 64>(dh, x, partial, off_W, (unsigned)M, st);
  }
#define WG(NOv, KIv, TNn, TKk, G) \
  if (NO == NOv && K == KIv) return launch_wgrad<NOv, KIv, TNn, TKk, G>(dh, x, partial, P, slots, off_W, (unsigned)M, st)
  WG(64, 32, 8, 4, 4);
  WG(32, 64, 4, 8, 4);
  WG(64, 64, 8, 8, 4);
  WG(128, 64, 8, 8, 2);
  WG(64, 128, 8, 8, 2);
  WG(256, 128, 16, 8, 1);
  WG(128, 256, 8, 16, 1);
#undef WG
  set_error("linear_bwd: unsupported weight shape [%d,%d]", NO, K);
  return GATRES_ERR_ARG;
}

static unsigned flat_grid(size_t work_items) {
  size_t g = (work_items + 255) / 256;
  const size_t cap = (size_t)sm_count() * 16;
  return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

extern "C" int gatres_encoder_fwd(const float* x, const float* w, const float* b, float* out, int64_t M,
                                  int32_t nc, void* stream) {
  GATRES_REQUIRE(M >= 0 && nc > 0 && nc % 4 == 0, "encoder_fwd: bad M=%lld nc=%d", (long long)M, nc);
  if (M == 0) return GATRES_OK;
  const size_t total4 = (size_t)M * (nc / 4);
  launch_kernel(encoder_fwd_kernel, dim3(flat_grid(total4)), dim3(256), 0, as_stream(stream), x, w, b, out, total4, nc / 4);
  return check_launch("encoder_fwd");
}

extern "C" int gatres_encoder_bwd(const float* g, const float* x, float* partial, int64_t P, int32_t slots,
                                  int64_t off_w, int64_t off_b, int64_t M, int32_t nc, void* stream) {
  GATRES_REQUIRE(M > 0 && M < (1ll << 31), "encoder_bwd: bad M");
  GATRES_REQUIRE(nc % 4 == 0 && 256 % (nc / 4) == 0 && nc / 4 <= 256, "encoder_bwd: unsupported nc=%d", nc);
  GATRES_REQUIRE(P % 4 == 0 && off_w % 4 == 0 && off_b % 4 == 0, "encoder_bwd: misaligned offsets");
  const int atomic = slots <= 0;
  const unsigned grid = atomic ? row_kernel_grid((unsigned)M, 256 / (nc / 4), 4) : (unsigned)slots;
  launch_kernel(encoder_bwd_kernel, dim3(grid), dim3(256), 0, as_stream(stream), g, x, partial, P, off_w, off_b, (unsigned)M, nc / 4, atomic);
  return check_launch("encoder_bwd");
}

extern "C" int gatres_decoder_fwd(const float* x, const float* w, const float* b, float* out,
                                  const int32_t* poison, int64_t M, int32_t nc, void* stream) {
  GATRES_REQUIRE(M >= 0 && M < (1ll << 31), "decoder_fwd: bad M=%lld", (long long)M);
  if (M == 0) return GATRES_OK;
  cudaStream_t st = as_stream(stream);
  const unsigned Mu = (unsigned)M;
  switch (nc) {
    case 32: launch_kernel(decoder_fwd_kernel<32>, dim3(flat_grid((size_t)M * 8)), dim3(256), 0, st, x, w, b, out, poison, Mu); break;
    case 64: launch_kernel(decoder_fwd_kernel<64>, dim3(flat_grid((size_t)M * 16)), dim3(256), 0, st, x, w, b, out, poison, Mu); break;
    case 128: launch_kernel(decoder_fwd_kernel<128>, dim3(flat_grid((size_t)M * 32)), dim3(256), 0, st, x, w, b, out, poison, Mu); break;
    default: set_error("decoder_fwd: unsupported nc=%d", nc); return GATRES_ERR_ARG;
  }
  return check_launch("decoder_fwd");
}

extern "C" int gatres_decoder_bwd(const float* g_out, const float* x, const float* w, float* dx, float* partial,
                                  int64_t P, int32_t slots, int64_t off_w, int64_t off_b, int64_t M, int32_t nc,
                                  int32_t mask_relu, void* stream) {
  GATRES_REQUIRE(M > 0 && M < (1ll << 31), "decoder_bwd: bad M");
  GATRES_REQUIRE(P % 4 == 0 && off_w % 4 == 0, "decoder_bwd: misaligned offsets");
  cudaStream_t st = as_stream(stream);
  const unsigned Mu = (unsigned)M;
  const int atomic = slots <= 0;
  const unsigned grid = atomic ? row_kernel_grid(Mu, kWarps * (32 / (nc / 4)), 8) : (unsigned)slots;
  switch (nc) {
    case 32: launch_kernel(decoder_bwd_kernel<32>, dim3(grid), dim3(256), 0, st, g_out, x, w, dx, partial, P, off_w, off_b, Mu, mask_relu, atomic); break;
    case 64: launch_kernel(decoder_bwd_kernel<64>, dim3(grid), dim3(256), 0, st, g_out, x, w, dx, partial, P, off_w, off_b, Mu, mask_relu, atomic); break;
    case 128: launch_kernel(decoder_bwd_kernel<128>, dim3(grid), dim3(256), 0, st, g_out, x, w, dx, partial, P, off_w, off_b, Mu, mask_relu, atomic); break;
    default: set_error("decoder_bwd: unsupported nc=%d", nc); return GATRES_ERR_ARG;
  }
  return check_launch("decoder_bwd");
}

extern "C" int gatres_reduce_partials(const float* partial, int64_t P, int32_t slots, int64_t p_begin,
                                      int64_t p_end, float* grads, void* stream) {
  GATRES_REQUIRE(slots > 0 && p_begin >= 0 && p_end >= p_begin && p_end <= P, "reduce_partials: bad range");
  if (p_end == p_begin) return GATRES_OK;
  launch_kernel(reduce_partials_kernel, dim3(flat_grid((size_t)(p_end - p_begin))), dim3(256), 0, as_stream(stream), 
      partial, P, slots, p_begin, p_end, grads);
  return check_launch("reduce_partials");
}
