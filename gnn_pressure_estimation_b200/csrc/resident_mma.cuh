// Tensor-core versions of the row-local contractions of the snapshot-resident kernels (included inside the
// RES_NS namespace of resident_impl.cuh; uses its T / lds helpers).
//
// The FFMA versions (project_rows / dgrad_rows / wgrad_rows) are bound by shared-memory bandwidth: a 4x4 register
// tile issues 8 LDS.128 per 64 FMA (profiles/r1_resident.md).  Here one warp owns a 16-row tile and issues
// mma.sync.m16n8k8 TF32 instructions whose fragments are distinct 4-byte shared loads per lane (no broadcast
// traffic): ~5x fewer shared-memory wavefronts and ~2.4x fewer instructions per output.  Operands stay fp32 in
// shared memory and are split on the fly into hi = the tensor core's own truncation of x and
// lo = rna_tf32(x - trunc x); a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi with fp32 accumulation ("3xTF32"), the same
// error-compensated scheme as the tcgen05 kernels of linear_tc.cu, so results stay within ~1e-6 of the FFMA path.
// (mma.sync, not tcgen05: the tiles are 16-64 rows per CTA, far below the 128-row UMMA atom, and the operands
// are already in shared memory for the gather phases.)
#pragma once

__device__ __forceinline__ unsigned tf32_lo_bits(float x) {
  const float r = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  return (__float_as_uint(r) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// c + cl += a * b with fp32-grade accuracy; a / b hold fp32 bit patterns (the tensor core ignores their low 13 bits).
// The two correction terms go to their own accumulators (cl, cm): three independent MMA chains per output tile
// instead of one chain three times as long (the phases are latency bound: 16-64 rows per CTA).
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], float (&cl)[4], float (&cm)[4], const unsigned (&a)[4],
                                           const unsigned (&al)[4], const unsigned (&b)[2], const unsigned (&bl)[2]) {
  mma_tf32(cl, al, b);
  mma_tf32(cm, a, bl);
  mma_tf32(c, a, b);
}
__device__ __forceinline__ void fold3(float (&c)[4], const float (&cl)[4], const float (&cm)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] += cl[i] + cm[i];
}
template <int NF>
__device__ __forceinline__ void split_frag(const float (&x)[NF], unsigned (&hi)[NF], unsigned (&lo)[NF]) {
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    hi[i] = __float_as_uint(x[i]);
    lo[i] = tf32_lo_bits(x[i]);
  }
}

// out[m][n] = sum_k A[m][k] W[n][k] (+ attention scores).  Warp w: 16-row tile (w % 4) of every 64-row group,
// column half w / 4 (for H = 2 the half is the head).  scr: 4 * n floats of shared scratch (H = 1 only).
template <int K, int NOUT, int H, int LDA, int LDW>
__device__ __forceinline__ void project_rows_mma(const float* As, const float* Ws, const float* att_s, const float* att_d,
                                                 float* h_own, float* ss_own, float* sd_own, float* scr, int n) {
  static_assert(T == 256, "warp tiling assumes 8 warps");
  constexpr int NTW = NOUT / 16;                 // 8-column tiles per warp (half of the columns)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int half = warp >> 2, nbase = half * (NOUT / 2);
  for (int m0 = (warp & 3) * 16; m0 < n; m0 += 64) {
    const int r0 = min(m0 + g, n - 1), r1 = min(m0 + g + 8, n - 1);
    float acc[NTW][4], acl[NTW][4], acm[NTW][4];
#pragma unroll
    for (int j = 0; j < NTW; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = acl[j][i] = acm[j][i] = 0.f;
#pragma unroll 2
    for (int k0 = 0; k0 < K; k0 += 8) {
      const float av[4] = {As[r0 * LDA + k0 + t], As[r1 * LDA + k0 + t], As[r0 * LDA + k0 + t + 4], As[r1 * LDA + k0 + t + 4]};
      unsigned a[4], al[4];
      split_frag<4>(av, a, al);
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        const float* wp = Ws + (nbase + 8 * j + g) * LDW + k0 + t;
        const float bv[2] = {wp[0], wp[4]};
        unsigned b[2], bl[2];
        split_frag<2>(bv, b, bl);
        mma_3xtf32(acc[j], acl[j], acm[j], a, al, b, bl);
      }
    }
#pragma unroll
    for (int j = 0; j < NTW; ++j) fold3(acc[j], acl[j], acm[j]);
    // fragment (g, 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1) of every 8-column tile
    float ps[2] = {0.f, 0.f}, pd[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      const int c = nbase + 8 * j + 2 * t;
      const float s0 = att_s[c], s1 = att_s[c + 1], d0 = att_d[c], d1 = att_d[c + 1];
      ps[0] = fmaf(acc[j][0], s0, fmaf(acc[j][1], s1, ps[0]));
      ps[1] = fmaf(acc[j][2], s0, fmaf(acc[j][3], s1, ps[1]));
      pd[0] = fmaf(acc[j][0], d0, fmaf(acc[j][1], d1, pd[0]));
      pd[1] = fmaf(acc[j][2], d0, fmaf(acc[j][3], d1, pd[1]));
      if (m0 + g < n) *reinterpret_cast<float2*>(h_own + (size_t)(m0 + g) * NOUT + c) = make_float2(acc[j][0], acc[j][1]);
      if (m0 + g + 8 < n) *reinterpret_cast<float2*>(h_own + (size_t)(m0 + g + 8) * NOUT + c) = make_float2(acc[j][2], acc[j][3]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ps[i] += __shfl_xor_sync(FULL, ps[i], 1); ps[i] += __shfl_xor_sync(FULL, ps[i], 2);
      pd[i] += __shfl_xor_sync(FULL, pd[i], 1); pd[i] += __shfl_xor_sync(FULL, pd[i], 2);
    }
    if (t == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int m = m0 + g + 8 * i;
        if (m < n) {
          if (H == 2) {                          // column half == head
            ss_own[m * 2 + half] = ps[i];
            sd_own[m * 2 + half] = pd[i];
          } else {                               // the two halves of the single head meet in shared memory
            scr[(half * 2 + 0) * n + m] = ps[i];
            scr[(half * 2 + 1) * n + m] = pd[i];
          }
        }
      }
    }
  }
  if (H == 1) {
    __syncthreads();
    for (int m = threadIdx.x; m < n; m += T) {
      ss_own[m] = scr[0 * n + m] + scr[2 * n + m];
      sd_own[m] = scr[1 * n + m] + scr[3 * n + m];
    }
  }
}

// out[m][c] = sum_r G[m][r] W[r][c]  (data gradient; W [NRED][NOUT] as stored).  pre(m, c) -> float2 is evaluated for
// every output fragment BEFORE the MMAs (global loads the epilogue needs fly under them); fin(m, c, value, pre value)
// gets columns c, c+1.
template <int NRED, int NOUT, int LDG, int LDW, typename Pre, typename Fin>
__device__ __forceinline__ void dgrad_rows_mma(const float* Gs, const float* Ws, int n, Pre pre, Fin fin) {
  constexpr int NTW = NOUT / 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int nbase = (warp >> 2) * (NOUT / 2);
  for (int m0 = (warp & 3) * 16; m0 < n; m0 += 64) {
    const int r0 = min(m0 + g, n - 1), r1 = min(m0 + g + 8, n - 1);
    float acc[NTW][4], acl[NTW][4], acm[NTW][4];
#pragma unroll
    for (int j = 0; j < NTW; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = acl[j][i] = acm[j][i] = 0.f;
    float2 pr[NTW][2];
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      const int c = nbase + 8 * j + 2 * t;
      pr[j][0] = pre(r0, c);
      pr[j][1] = pre(r1, c);
    }
#pragma unroll 2
    for (int k0 = 0; k0 < NRED; k0 += 8) {
      const float av[4] = {Gs[r0 * LDG + k0 + t], Gs[r1 * LDG + k0 + t], Gs[r0 * LDG + k0 + t + 4], Gs[r1 * LDG + k0 + t + 4]};
      unsigned a[4], al[4];
      split_frag<4>(av, a, al);
#pragma unroll
      for (int j = 0; j < NTW; ++j) {
        const float* wp = Ws + (k0 + t) * LDW + nbase + 8 * j + g;
        const float bv[2] = {wp[0], wp[4 * LDW]};
        unsigned b[2], bl[2];
        split_frag<2>(bv, b, bl);
        mma_3xtf32(acc[j], acl[j], acm[j], a, al, b, bl);
      }
    }
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      fold3(acc[j], acl[j], acm[j]);
      const int c = nbase + 8 * j + 2 * t;
      if (m0 + g < n) fin(m0 + g, c, make_float2(acc[j][0], acc[j][1]), pr[j][0]);
      if (m0 + g + 8 < n) fin(m0 + g + 8, c, make_float2(acc[j][2], acc[j][3]), pr[j][1]);
    }
  }
}

// dW[no][ki] += sum_m G[m][no] X[m][ki]: 16 x 8 output tiles, the reduction runs over this CTA's rows (zero padded)
template <int NO, int KI, int LDG, int LDXX>
__device__ __forceinline__ void wgrad_rows_mma(const float* Gs, const float* Xs, int n, float* dW) {
  constexpr int TILES = (NO / 16) * (KI / 8), TPW = TILES / (T / 32);
  static_assert(TILES % (T / 32) == 0, "wgrad tiles per warp");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  if (n <= 0) return;
  float acc[TPW][4], acl[TPW][4], acm[TPW][4];
  int no0[TPW], ki0[TPW];
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    const int tile = warp * TPW + q;
    no0[q] = (tile / (KI / 8)) * 16;
    ki0[q] = (tile % (KI / 8)) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[q][i] = acl[q][i] = acm[q][i] = 0.f;
  }
#pragma unroll 2
  for (int m0 = 0; m0 < n; m0 += 8) {
    const bool k0ok = m0 + t < n, k1ok = m0 + t + 4 < n;
#pragma unroll
    for (int q = 0; q < TPW; ++q) {
      const float* g0 = Gs + (m0 + t) * LDG + no0[q] + g;
      const float* g1 = Gs + (m0 + t + 4) * LDG + no0[q] + g;
      const float av[4] = {k0ok ? g0[0] : 0.f, k0ok ? g0[8] : 0.f, k1ok ? g1[0] : 0.f, k1ok ? g1[8] : 0.f};
      const float bv[2] = {k0ok ? Xs[(m0 + t) * LDXX + ki0[q] + g] : 0.f, k1ok ? Xs[(m0 + t + 4) * LDXX + ki0[q] + g] : 0.f};
      unsigned a[4], al[4], b[2], bl[2];
      split_frag<4>(av, a, al);
      split_frag<2>(bv, b, bl);
      mma_3xtf32(acc[q], acl[q], acm[q], a, al, b, bl);
    }
  }
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    fold3(acc[q], acl[q], acm[q]);
    atomicAdd(reinterpret_cast<float2*>(dW + (size_t)(no0[q] + g) * KI + ki0[q] + 2 * t), make_float2(acc[q][0], acc[q][1]));
    atomicAdd(reinterpret_cast<float2*>(dW + (size_t)(no0[q] + g + 8) * KI + ki0[q] + 2 * t), make_float2(acc[q][2], acc[q][3]));
  }
}
