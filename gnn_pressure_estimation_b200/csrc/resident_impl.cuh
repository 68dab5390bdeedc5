// Implementation body of the snapshot-resident kernels; included by resident.cu once per CTA size
// (RES_NS = namespace, RES_T = threads per CTA, RES_MIN_CTAS = CTAs per SM the register budget must allow).
namespace gatres {
namespace RES_NS {

constexpr int NC = 32;             // channels (gatres_small); other widths use the layer kernels
constexpr int T = RES_T;           // threads per CTA
constexpr int LDX = NC + 4;        // padded row of a [*, 32] shared tile
constexpr int LDY = 2 * NC + 4;    // padded row of a [*, 64] shared tile
constexpr int W1F = 2 * NC * LDX;  // conv1 weight [64][32] padded
constexpr int W2F = NC * LDY;      // conv2 weight [32][64] padded
constexpr int VECF = 9 * NC;       // as1 ad1 b1 (64 each) as2 ad2 b2 (32 each)
constexpr unsigned FULL = 0xffffffffu;

struct Args {
  const int* rowptr;
  const int* col;
  const int* rowptr_t;
  const int* col_t;
  const float* params;
  const float* x;        // [M] model input (masked nodes already zeroed)
  float* out;            // fwd: [M]
  float* saved;          // fwd: training activations or NULL; bwd: the same buffer
  float* scratch;
  const float* d_out;    // bwd: [M]
  float* grads;          // bwd: flat gradient buffer (atomic accumulation)
  const int* poison;
  long long M;
  int N, nb, R, B, E1;
  int k_hi, k_lo, head, tail;
  long long* prof;       // optional phase-timestamp buffer [CTA][kProfSlots] (tools/resident_probe.py), else NULL
  int prof_slots;
};

// ---- cluster / memory primitives ------------------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
// all threads of all CTAs of the cluster; global writes before it are visible cluster-wide after it
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// phase timestamps for the profiling tool: thread 0 of every CTA records the global timer
struct Stamper {
  long long* p;
  int left;
  __device__ __forceinline__ Stamper(const Args& a) {
    p = (a.prof != nullptr && threadIdx.x == 0) ? a.prof + (size_t)blockIdx.x * a.prof_slots : nullptr;
    left = a.prof_slots;
  }
  __device__ __forceinline__ void operator()() {
    if (p != nullptr && left > 0) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      *p++ = t;
      --left;
    }
  }
};
// coherent loads: tensors produced inside this kernel by other CTAs of the cluster (never ld.global.nc)
__device__ __forceinline__ float4 ldc4(const float* p) {
  float4 r;
  asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldc1(const float* p) {
  float r;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }

__device__ __forceinline__ void cp16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__device__ __forceinline__ float gmax8(float v) {
  v = fmaxf(v, __shfl_xor_sync(FULL, v, 4));
  v = fmaxf(v, __shfl_xor_sync(FULL, v, 2));
  return fmaxf(v, __shfl_xor_sync(FULL, v, 1));
}
__device__ __forceinline__ float4 relu4(float4 v) {
  return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}
__device__ __forceinline__ float4 mask4(float4 v, float4 ref) {
  return make_float4(ref.x > 0.f ? v.x : 0.f, ref.y > 0.f ? v.y : 0.f, ref.z > 0.f ? v.z : 0.f, ref.w > 0.f ? v.w : 0.f);
}

// one block's parameters (contiguous W1 as1 ad1 b1 W2 as2 ad2 b2) -> padded shared tiles, 16 B per cp.async
__device__ __forceinline__ void stage_block_params(const float* blk, float* W1s, float* W2s, float* vec) {
  constexpr int C_W1 = 2 * NC * NC / 4, C_V1 = 6 * NC / 4, C_W2 = 2 * NC * NC / 4, C_V2 = 3 * NC / 4;
  for (int c = threadIdx.x; c < C_W1 + C_V1 + C_W2 + C_V2; c += T) {
    float* dst;
    if (c < C_W1) dst = W1s + (c / (NC / 4)) * LDX + 4 * (c % (NC / 4));
    else if (c < C_W1 + C_V1) dst = vec + 4 * (c - C_W1);
    else if (c < C_W1 + C_V1 + C_W2) {
      const int q = c - C_W1 - C_V1;
      dst = W2s + (q / (2 * NC / 4)) * LDY + 4 * (q % (2 * NC / 4));
    } else dst = vec + 6 * NC + 4 * (c - C_W1 - C_V1 - C_W2);
    cp16(dst, blk + 4 * c);
  }
}

// own rows [0, n) of a row-major global tensor -> padded shared tile (cp.async, 16 B chunks)
template <int F, int LD>
__device__ __forceinline__ void stage_rows(const float* g, float* s, int n) {
  for (int c = threadIdx.x; c < n * (F / 4); c += T) cp16(s + (c / (F / 4)) * LD + 4 * (c % (F / 4)), g + 4 * c);
}

// ---- row-local projections (A operand resident in shared memory) ---------------------
// out[m][n] = sum_k A[m][k] W[n][k]  (W [NOUT][K] as stored by PyG Linear), + attention scores.
// Thread (tx, ty): output columns {tx + TX j}, rows {m0 + ty + TY i}; k ascending like linear.cu.
template <int K, int NOUT, int H, int LDA, int LDW>
__device__ __forceinline__ void project_rows(const float* As, const float* Ws, const float* att_s, const float* att_d,
                                             float* h_own, float* ss_own, float* sd_own, int n) {
  constexpr int TN = 4, TX = NOUT / TN, TY = T / TX, TM = 64 / TY;
  static_assert(TM * TY == 64 && (TX == 8 || TX == 16), "tiling");
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  float as_[TN], ad_[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) { as_[j] = att_s[tx + TX * j]; ad_[j] = att_d[tx + TX * j]; }
  for (int m0 = 0; m0 < n; m0 += 64) {
    float acc[TM][TN];
    int row[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      row[i] = min(m0 + ty + TY * i, n - 1);
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    }
#pragma unroll 2
    for (int k4 = 0; k4 < K / 4; ++k4) {
      float4 a[TM], w[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = lds4(As + row[i] * LDA + 4 * k4);
#pragma unroll
      for (int j = 0; j < TN; ++j) w[j] = lds4(Ws + (tx + TX * j) * LDW + 4 * k4);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty + TY * i;
      const bool ok = m < n;
      // column tx + TX j belongs to head (tx + TX j) / 32: H = 2 -> j / 2 (TX = 16); H = 1 -> 0
      float ps[H], pd[H];
#pragma unroll
      for (int hh = 0; hh < H; ++hh) ps[hh] = pd[hh] = 0.f;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int hh = H == 1 ? 0 : j / 2;
        ps[hh] = fmaf(acc[i][j], as_[j], ps[hh]);
        pd[hh] = fmaf(acc[i][j], ad_[j], pd[hh]);
      }
#pragma unroll
      for (int hh = 0; hh < H; ++hh) {
        ps[hh] = group_sum<TX>(ps[hh], FULL);
        pd[hh] = group_sum<TX>(pd[hh], FULL);
      }
      if (ok) {
#pragma unroll
        for (int j = 0; j < TN; ++j) h_own[(size_t)m * NOUT + tx + TX * j] = acc[i][j];
        if (tx == 0) {
#pragma unroll
          for (int hh = 0; hh < H; ++hh) { ss_own[m * H + hh] = ps[hh]; sd_own[m * H + hh] = pd[hh]; }
        }
      }
    }
  }
}

// dx[m][k] = sum_n G[m][n] W[n][k]   (data gradient; W [NRED][NOUT] as stored).  Thread owns 4 consecutive
// output columns.  Epilogue supplied by the caller through `fin(m, col4, value)`.
template <int NRED, int NOUT, int LDG, int LDW, typename Fin>
__device__ __forceinline__ void dgrad_rows(const float* Gs, const float* Ws, int n, Fin fin) {
  constexpr int TX = NOUT / 4, TY = T / TX, TM = 64 / TY;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  for (int m0 = 0; m0 < n; m0 += 64) {
    float4 acc[TM];
    int row[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) { row[i] = min(m0 + ty + TY * i, n - 1); acc[i] = f4zero(); }
#pragma unroll 2
    for (int n4 = 0; n4 < NRED / 4; ++n4) {
      float4 a[TM], w[4];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = lds4(Gs + row[i] * LDG + 4 * n4);
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = lds4(Ws + (4 * n4 + q) * LDW + 4 * tx);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        fma4(acc[i], a[i].x, w[0]);
        fma4(acc[i], a[i].y, w[1]);
        fma4(acc[i], a[i].z, w[2]);
        fma4(acc[i], a[i].w, w[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty + TY * i;
      if (m < n) fin(m, 4 * tx, acc[i]);
    }
  }
}

// dW[no][ki] += sum_m G[m][no] X[m][ki]: thread owns a 4 x KPT block, atomics at the end.
template <int NO, int KI, int LDG, int LDXX>
__device__ __forceinline__ void wgrad_rows(const float* Gs, const float* Xs, int n, float* dW) {
  constexpr int TNN = NO / 4;                 // threads along the output rows
  constexpr int KPT = NO * KI / (4 * T);      // k columns per thread (2 at 256 threads, 1 at 512)
  static_assert(KPT == 1 || KPT == 2, "wgrad tiling");
  const int tn = threadIdx.x % TNN, tk = threadIdx.x / TNN;
  float acc[4][KPT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int q = 0; q < KPT; ++q) acc[i][q] = 0.f;
#pragma unroll 4
  for (int m = 0; m < n; ++m) {
    const float4 g = lds4(Gs + m * LDG + 4 * tn);
    float xv[KPT];
    if (KPT == 2) {
      const float2 t = lds2(Xs + m * LDXX + 2 * tk);
      xv[0] = t.x; xv[KPT - 1] = t.y;
    } else {
      xv[0] = Xs[m * LDXX + tk];
    }
#pragma unroll
    for (int q = 0; q < KPT; ++q) {
      acc[0][q] = fmaf(g.x, xv[q], acc[0][q]);
      acc[1][q] = fmaf(g.y, xv[q], acc[1][q]);
      acc[2][q] = fmaf(g.z, xv[q], acc[2][q]);
      acc[3][q] = fmaf(g.w, xv[q], acc[3][q]);
    }
  }
  if (n > 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* dst = dW + (size_t)(4 * tn + i) * KI + KPT * tk;
      if (KPT == 2) atomicAdd(reinterpret_cast<float2*>(dst), make_float2(acc[i][0], acc[i][KPT - 1]));
      else atomicAdd(dst, acc[i][0]);
    }
  }
}

#if RES_USE_MMA
#include "resident_mma.cuh"
#endif

// sum per-lane float4 accumulators over every lane of the CTA that holds the same chunk (lane % LPR) and add
// the LPR chunk sums into dst (atomic).  red: T*4 floats of shared scratch.  All threads call.
template <int LPR>
__device__ __forceinline__ void cta_chunk_sum_atomic(float4 acc, float* red, float* dst) {
  __syncthreads();
  st4(red + threadIdx.x * 4, acc);
  __syncthreads();
  if (threadIdx.x < LPR) {
    float4 s = f4zero();
    for (int t = threadIdx.x; t < T; t += LPR) add4(s, lds4(red + t * 4));
    atomicAdd(reinterpret_cast<float4*>(dst + 4 * threadIdx.x), s);
  }
}

// ---- fused GAT aggregation over this CTA's rows (C = 32) -------------------------------------------------
// Eight lanes own a row: lane `slot` holds the float4 chunk `slot` of EVERY head of the row (H chunks per lane) and
// evaluates edge `slot` of the row for every head, so a warp covers four rows per pass whatever H is, and the
// index work (row pointers, neighbour ids, their shuffles) is shared by the heads.  (The first version gave each
// head its own eight lanes: twice the passes and twice the index instructions for conv1 — these phases are issue
// bound, profiles/r1_resident.md.)
template <int H, bool TRAIN>
__device__ __forceinline__ void agg_fwd_rows(const int* __restrict__ rowptr, const int* __restrict__ col,
                                             const float* hsnap, const float* sssnap, const float* sd_own,
                                             const float* bias_s, float* out_g, float* out_s, int lds_out,
                                             float* m_own, float* l_own, int lo, int n, bool relu) {
  constexpr int F = 32 * H, RPW = 4, PRE = 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, slot = lane & 7;
  float4 bv[H];
#pragma unroll
  for (int v = 0; v < H; ++v) bv[v] = lds4(bias_s + 32 * v + 4 * slot);
  for (int i0 = 0; i0 < n; i0 += (T / 32) * RPW) {
    const int il_raw = i0 + warp * RPW + sub;
    const bool ok = il_raw < n;
    const int il = ok ? il_raw : n - 1, i = lo + il;
    const int beg = rowptr[i], deg = rowptr[i + 1] - beg;
    const int deg_max = __reduce_max_sync(FULL, deg);
    float sd[H], mrun[H], lrun[H];
    float4 acc[H];
#pragma unroll
    for (int v = 0; v < H; ++v) {
      sd[v] = ldc1(sd_own + il * H + v);
      mrun[v] = -CUDART_INF_F;
      lrun[v] = 0.f;
      acc[v] = f4zero();
    }
    for (int e0 = 0; e0 < deg_max; e0 += 8) {
      const bool valid = e0 + slot < deg;
      const int j = valid ? col[beg + e0 + slot] : 0;
      const int cnt = min(8, deg - e0), cnt_max = min(8, deg_max - e0);
      float4 x[PRE][H];
#pragma unroll
      for (int u = 0; u < PRE; ++u) {
        const int ju = __shfl_sync(FULL, j, u, 8);
#pragma unroll
        for (int v = 0; v < H; ++v) x[u][v] = u < cnt ? ldc4(hsnap + (size_t)ju * F + 32 * v + 4 * slot) : f4zero();
      }
      float p[H];
#pragma unroll
      for (int v = 0; v < H; ++v) {
        const float a = valid ? lrelu(ldc1(sssnap + j * H + v) + sd[v]) : -CUDART_INF_F;
        const float nm = fmaxf(mrun[v], gmax8(a));
        if (e0 > 0) {
          const float sc = __expf(mrun[v] - nm);
          lrun[v] *= sc;
          acc[v].x *= sc; acc[v].y *= sc; acc[v].z *= sc; acc[v].w *= sc;
        }
        p[v] = __expf(a - nm);
        lrun[v] += group_sum<8>(p[v], FULL);
        mrun[v] = nm;
      }
#pragma unroll
      for (int u = 0; u < PRE; ++u)
#pragma unroll
        for (int v = 0; v < H; ++v) fma4(acc[v], __shfl_sync(FULL, p[v], u, 8), x[u][v]);
      for (int t = PRE; t < cnt_max; t += 2) {
        const int j0 = __shfl_sync(FULL, j, t, 8), j1 = __shfl_sync(FULL, j, t + 1, 8);
#pragma unroll
        for (int v = 0; v < H; ++v) {
          const float4 x0 = t < cnt ? ldc4(hsnap + (size_t)j0 * F + 32 * v + 4 * slot) : f4zero();
          const float4 x1 = t + 1 < cnt ? ldc4(hsnap + (size_t)j1 * F + 32 * v + 4 * slot) : f4zero();
          const float p0 = __shfl_sync(FULL, p[v], t, 8), p1 = __shfl_sync(FULL, p[v], t + 1, 8);
          fma4(acc[v], p0, x0);
          fma4(acc[v], t + 1 < 8 ? p1 : 0.f, x1);
        }
      }
    }
    if (!ok) continue;
#pragma unroll
    for (int v = 0; v < H; ++v) {
      const float inv = 1.f / (lrun[v] + kSoftmaxEps);
      float4 o = make_float4(fmaf(acc[v].x, inv, bv[v].x), fmaf(acc[v].y, inv, bv[v].y), fmaf(acc[v].z, inv, bv[v].z),
                             fmaf(acc[v].w, inv, bv[v].w));
      if (relu) o = relu4(o);
      if (out_g != nullptr) st4(out_g + (size_t)il * F + 32 * v + 4 * slot, o);
      if (out_s != nullptr) st4(out_s + il * lds_out + 32 * v + 4 * slot, o);
      if (TRAIN && slot == 0) { m_own[il * H + v] = mrun[v]; l_own[il * H + v] = lrun[v]; }
    }
  }
}

// sum a per-lane float4 over the four row slots of a warp and park it in this warp's row of `vred`
__device__ __forceinline__ void warp_chunk_park(float4 v, float* dst) {
#pragma unroll
  for (int o = 8; o < 32; o <<= 1) {
    v.x += __shfl_xor_sync(FULL, v.x, o); v.y += __shfl_xor_sync(FULL, v.y, o);
    v.z += __shfl_xor_sync(FULL, v.z, o); v.w += __shfl_xor_sync(FULL, v.w, o);
  }
  if ((threadIdx.x & 31) < 8) st4(dst + 4 * (threadIdx.x & 31), v);
}

// Backward pass 1 over own target rows: D_i, ds_dst[i]; rec = {s_dst, m, 1/(l+eps), D}.
// MEAN = true (conv2, H = 1): the incoming gradient is produced on the fly as the SimpleConv(mean) backward
// of gA (dz[j] = sum_{j->i} gA[i] / max(indeg(i), 1)) and also written to dz_own for pass 2's gathers;
// MEAN = false (conv1): it is read from the shared tile g_s.
template <int H, bool MEAN>
__device__ __forceinline__ void bwd_p1_rows(const int* __restrict__ rowptr, const int* __restrict__ col,
                                            const int* __restrict__ rowptr_t, const int* __restrict__ col_t,
                                            const float* gA_snap, float* dz_own, const float* g_s, int ldg_s,
                                            const float* __restrict__ hsnap, const float* __restrict__ sssnap,
                                            const float* __restrict__ sd_own, const float* __restrict__ m_own,
                                            const float* __restrict__ l_own, float* rec_own, float* dsd_own,
                                            float* vred_bias, int lo, int n) {
  static_assert(!MEAN || H == 1, "the mean backward feeds conv2 (one head)");
  constexpr int F = 32 * H, RPW = 4, PRE = 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, slot = lane & 7;
  float4 bacc[H];
#pragma unroll
  for (int v = 0; v < H; ++v) bacc[v] = f4zero();
  for (int i0 = 0; i0 < n; i0 += (T / 32) * RPW) {
    const int il_raw = i0 + warp * RPW + sub;
    const bool ok = il_raw < n;
    const int il = ok ? il_raw : n - 1, i = lo + il;
    const int beg = rowptr[i], deg = rowptr[i + 1] - beg;
    const int deg_max = __reduce_max_sync(FULL, deg);
    // Everything that does not depend on the incoming gradient is requested first: the first chunk's neighbour rows
    // and scores (saved activations: L2 or HBM) fly while the mean backward below waits for its own gathers of gA.
    const bool valid0 = slot < deg;
    const int j0 = valid0 ? col[beg + slot] : 0;
    const int cnt0 = min(8, deg);
    float4 x0[PRE][H];
    float ssv0[H], sd[H], mi[H], il_[H], S1[H], S2[H], S3[H];
#pragma unroll
    for (int u = 0; u < PRE; ++u) {
      const int ju = __shfl_sync(FULL, j0, u, 8);
#pragma unroll
      for (int v = 0; v < H; ++v) x0[u][v] = u < cnt0 ? ldg4(hsnap + (size_t)ju * F + 32 * v + 4 * slot) : f4zero();
    }
#pragma unroll
    for (int v = 0; v < H; ++v) {
      ssv0[v] = __ldg(sssnap + j0 * H + v);
      sd[v] = __ldg(sd_own + il * H + v);
      mi[v] = __ldg(m_own + il * H + v);
      il_[v] = 1.f / (__ldg(l_own + il * H + v) + kSoftmaxEps);
      S1[v] = S2[v] = S3[v] = 0.f;
    }
    float4 gv[H];
    if (MEAN) {
      const int tb = rowptr_t[i], te = rowptr_t[i + 1] - 1;     // out-edges minus the self-loop
      gv[0] = f4zero();
#pragma unroll 4
      for (int e = tb; e < te; ++e) {
        const int t = col_t[e];
        const int dg = rowptr[t + 1] - rowptr[t] - 1;
        fma4(gv[0], 1.f / (float)(dg > 1 ? dg : 1), ldc4(gA_snap + (size_t)t * F + 4 * slot));
      }
      if (ok) st4(dz_own + (size_t)il * F + 4 * slot, gv[0]);
    } else {
#pragma unroll
      for (int v = 0; v < H; ++v) gv[v] = lds4(g_s + il * ldg_s + 32 * v + 4 * slot);
    }
#pragma unroll
    for (int v = 0; v < H; ++v)
      if (ok) add4(bacc[v], gv[v]);
    for (int e0 = 0; e0 < deg_max; e0 += 8) {
      const bool valid = e0 + slot < deg;
      const int j = e0 == 0 ? j0 : (valid ? col[beg + e0 + slot] : 0);
      const int cnt = min(8, deg - e0), cnt_max = min(8, deg_max - e0);
      float4 x[PRE][H];
      float alpha[H], sl[H], da[H];
      if (e0 == 0) {
#pragma unroll
        for (int u = 0; u < PRE; ++u)
#pragma unroll
          for (int v = 0; v < H; ++v) x[u][v] = x0[u][v];
      } else {
#pragma unroll
        for (int u = 0; u < PRE; ++u) {
          const int ju = __shfl_sync(FULL, j, u, 8);
#pragma unroll
          for (int v = 0; v < H; ++v) x[u][v] = u < cnt ? ldg4(hsnap + (size_t)ju * F + 32 * v + 4 * slot) : f4zero();
        }
      }
#pragma unroll
      for (int v = 0; v < H; ++v) {
        const float z = (e0 == 0 ? ssv0[v] : __ldg(sssnap + j * H + v)) + sd[v];
        alpha[v] = valid ? __expf(lrelu(z) - mi[v]) * il_[v] : 0.f;
        sl[v] = lrelu_slope(z);
        da[v] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < PRE; ++u)
#pragma unroll
        for (int v = 0; v < H; ++v) {
          const float d = group_sum<8>(dot4(gv[v], x[u][v]), FULL);
          da[v] = slot == u ? d : da[v];
        }
      for (int t = PRE; t < cnt_max; t += 2) {
        const int j0_ = __shfl_sync(FULL, j, t, 8), j1_ = __shfl_sync(FULL, j, t + 1, 8);
#pragma unroll
        for (int v = 0; v < H; ++v) {
          const float4 xa = t < cnt ? ldg4(hsnap + (size_t)j0_ * F + 32 * v + 4 * slot) : f4zero();
          const float4 xb = t + 1 < cnt ? ldg4(hsnap + (size_t)j1_ * F + 32 * v + 4 * slot) : f4zero();
          const float d0 = group_sum<8>(dot4(gv[v], xa), FULL), d1 = group_sum<8>(dot4(gv[v], xb), FULL);
          da[v] = slot == t ? d0 : (slot == t + 1 ? d1 : da[v]);
        }
      }
#pragma unroll
      for (int v = 0; v < H; ++v) {
        S1[v] = fmaf(alpha[v], da[v], S1[v]);
        S2[v] = fmaf(alpha[v] * sl[v], da[v], S2[v]);
        S3[v] = fmaf(alpha[v], sl[v], S3[v]);
      }
    }
#pragma unroll
    for (int v = 0; v < H; ++v) {
      const float D = group_sum<8>(S1[v], FULL), T2 = group_sum<8>(S2[v], FULL), T3 = group_sum<8>(S3[v], FULL);
      if (slot == 0 && ok) {
        st4(rec_own + (size_t)(il * H + v) * 4, make_float4(sd[v], mi[v], il_[v], D));
        dsd_own[il * H + v] = T2 - D * T3;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < H; ++v) warp_chunk_park(bacc[v], vred_bias + 32 * v);
}

// Backward pass 2 over own source rows: dh[j] (-> shared tile), ds_src, datt_src / datt_dst partials.
template <int H>
__device__ __forceinline__ void bwd_p2_rows(const int* __restrict__ rowptr_t, const int* __restrict__ col_t,
                                            const float* gsnap, const float* recsnap, const float* dsd_own,
                                            const float* __restrict__ h_own, const float* __restrict__ ss_own,
                                            const float* att_s, const float* att_d, float* dh_s, int ld_dh,
                                            float* vred_as, float* vred_ad, int lo, int n) {
  constexpr int F = 32 * H, RPW = 4, PRE = H == 1 ? 4 : 2;     // two heads per lane: fewer gathers in flight (registers)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, slot = lane & 7;
  float4 as[H], ad[H], accs[H], accd[H];
#pragma unroll
  for (int v = 0; v < H; ++v) {
    as[v] = lds4(att_s + 32 * v + 4 * slot);
    ad[v] = lds4(att_d + 32 * v + 4 * slot);
    accs[v] = f4zero();
    accd[v] = f4zero();
  }
  for (int i0 = 0; i0 < n; i0 += (T / 32) * RPW) {
    const int il_raw = i0 + warp * RPW + sub;
    const bool ok = il_raw < n;
    const int il = ok ? il_raw : n - 1, jn = lo + il;
    const int beg = rowptr_t[jn], deg = rowptr_t[jn + 1] - beg;
    const int deg_max = __reduce_max_sync(FULL, deg);
    float4 hv[H], dacc[H];
    float ss[H], dsrc[H];
#pragma unroll
    for (int v = 0; v < H; ++v) {
      hv[v] = ldg4(h_own + (size_t)il * F + 32 * v + 4 * slot);
      ss[v] = __ldg(ss_own + il * H + v);
      dacc[v] = f4zero();
      dsrc[v] = 0.f;
    }
    for (int e0 = 0; e0 < deg_max; e0 += 8) {
      const bool valid = e0 + slot < deg;
      const int i = valid ? col_t[beg + e0 + slot] : 0;
      const int cnt = min(8, deg - e0), cnt_max = min(8, deg_max - e0);
      float4 gx[PRE][H];
#pragma unroll
      for (int u = 0; u < PRE; ++u) {
        const int iu = __shfl_sync(FULL, i, u, 8);
#pragma unroll
        for (int v = 0; v < H; ++v) gx[u][v] = u < cnt ? ldc4(gsnap + (size_t)iu * F + 32 * v + 4 * slot) : f4zero();
      }
      float alpha[H], k2[H], Dt[H], da[H];
#pragma unroll
      for (int v = 0; v < H; ++v) {
        const float4 t4 = ldc4(recsnap + (size_t)(i * H + v) * 4);      // {s_dst, m, 1/l, D} of the edge's target
        const float z = ss[v] + t4.x;
        alpha[v] = valid ? __expf(lrelu(z) - t4.y) * t4.z : 0.f;
        k2[v] = alpha[v] * lrelu_slope(z);
        Dt[v] = t4.w;
        da[v] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < PRE; ++u)
#pragma unroll
        for (int v = 0; v < H; ++v) {
          const float d = group_sum<8>(dot4(gx[u][v], hv[v]), FULL);
          da[v] = slot == u ? d : da[v];
          fma4(dacc[v], __shfl_sync(FULL, alpha[v], u, 8), gx[u][v]);
        }
      for (int t = PRE; t < cnt_max; t += 2) {
        const int i0_ = __shfl_sync(FULL, i, t, 8), i1_ = __shfl_sync(FULL, i, t + 1, 8);
#pragma unroll
        for (int v = 0; v < H; ++v) {
          const float4 g0 = t < cnt ? ldc4(gsnap + (size_t)i0_ * F + 32 * v + 4 * slot) : f4zero();
          const float4 g1 = t + 1 < cnt ? ldc4(gsnap + (size_t)i1_ * F + 32 * v + 4 * slot) : f4zero();
          const float d0 = group_sum<8>(dot4(g0, hv[v]), FULL), d1 = group_sum<8>(dot4(g1, hv[v]), FULL);
          da[v] = slot == t ? d0 : (slot == t + 1 ? d1 : da[v]);
          const float a0 = __shfl_sync(FULL, alpha[v], t, 8), a1 = __shfl_sync(FULL, alpha[v], t + 1, 8);
          fma4(dacc[v], a0, g0);
          fma4(dacc[v], t + 1 < 8 ? a1 : 0.f, g1);
        }
      }
#pragma unroll
      for (int v = 0; v < H; ++v) dsrc[v] = fmaf(k2[v], da[v] - Dt[v], dsrc[v]);
    }
#pragma unroll
    for (int v = 0; v < H; ++v) {
      const float ds = group_sum<8>(dsrc[v], FULL);
      const float dd = ldc1(dsd_own + il * H + v);
      fma4(dacc[v], ds, as[v]);
      fma4(dacc[v], dd, ad[v]);
      if (ok) {
        st4(dh_s + il * ld_dh + 32 * v + 4 * slot, dacc[v]);
        fma4(accs[v], ds, hv[v]);
        fma4(accd[v], dd, hv[v]);
      }
    }
  }
#pragma unroll
  for (int v = 0; v < H; ++v) {
    warp_chunk_park(accs[v], vred_as + 32 * v);
    warp_chunk_park(accd[v], vred_ad + 32 * v);
  }
}

// =============================================================================== forward
template <bool TRAIN>
__global__ void __launch_bounds__(T, RES_MIN_CTAS)
resident_fwd_kernel(const Args a) {
  extern __shared__ __align__(16) float smem[];
  const int R = a.R, N = a.N;
  float* xs = smem;                        // [R][LDX] block input / output (own rows)
  float* ys = xs + R * LDX;                // [R][LDY] conv1 output (own rows)
  float* W1s = ys + R * LDY;               // [2][64][LDX]
  float* W2s = W1s + 2 * W1F;              // [2][32][LDY]
  float* vec = W2s + 2 * W2F;              // [2][288]
  float* scr = vec + 2 * VECF;             // [4][R] score partials of the tensor-core conv2 projection
  int* rp_s = reinterpret_cast<int*>(scr + a4(4 * R));    // shared CSR of the water network: rowptr [N+1]
  int* col_s = rp_s + a4(N + 1);                          //                                   col [E1]
  const int rank = (int)cluster_ctarank();
  const long long b = cluster_id_x();
  const int lo = rank * R, n = max(0, min(R, N - lo));
  const long long M = a.M, rb = b * N, ro = rb + lo;    // snapshot base row, first own row
  const ParamLayout pl(a.nb, NC);
  const SavedLayout sl(M, NC);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // the topology is constant across steps: stage it before the dependency wait (L1 is invalidated at every
  // cluster barrier, so reading it from global would cost two extra L2 round trips per row pass)
  for (int c = threadIdx.x; c <= N; c += T) rp_s[c] = __ldg(a.rowptr + c);
  for (int c = threadIdx.x; c < a.E1; c += T) col_s[c] = __ldg(a.col + c);
  pdl_wait();
  Stamper stamp(a);
  stamp();
  if (a.nb > 0) stage_block_params(a.params + pl.block(0), W1s, W2s, vec);
  cp_commit();

  // encoder Linear(1, nc): x0[i][c] = x[i] w[c] + b[c]
  {
    const int lig = lane & 7, sub = lane >> 3;
    const float4 wv = ldg4(a.params + pl.lin0_w() + 4 * lig), bv = ldg4(a.params + pl.lin0_b() + 4 * lig);
    for (int il = warp * 4 + sub; il < n; il += T / 8) {
      const float xv = __ldg(a.x + ro + il);
      const float4 o = make_float4(fmaf(xv, wv.x, bv.x), fmaf(xv, wv.y, bv.y), fmaf(xv, wv.z, bv.z), fmaf(xv, wv.w, bv.w));
      st4(xs + il * LDX + 4 * lig, o);
      if (TRAIN) st4(a.saved + sl.x_enc() + (ro + il) * NC + 4 * lig, o);
    }
  }

  // inference: rolling buffers carved from scratch (layout of gatres_scratch_floats, training = 0)
  float* sc = a.scratch;
  for (int k = 0; k < a.nb; ++k) {
    const int buf = k & 1;
    float* W1 = W1s + buf * W1F;
    float* W2 = W2s + buf * W2F;
    float* vc = vec + buf * VECF;
    float* h1 = TRAIN ? a.saved + sl.h1(k) : sc + 2 * M * NC;
    float* ss1 = TRAIN ? a.saved + sl.ss1(k) : sc + 8 * M * NC;
    float* sd1 = TRAIN ? a.saved + sl.sd1(k) : sc + 8 * M * NC + a4(2 * M);
    float* h2 = TRAIN ? a.saved + sl.h2(k) : sc + 6 * M * NC;
    float* ss2 = TRAIN ? a.saved + sl.ss2(k) : sc;
    float* sd2 = TRAIN ? a.saved + sl.sd2(k) : sc + a4(M);
    float* z = TRAIN ? sc : sc + 7 * M * NC;

    cp_wait_all();
    __syncthreads();                         // parameters of block k and xs are in place
    if (k + 1 < a.nb) stage_block_params(a.params + pl.block(k + 1), W1s + (buf ^ 1) * W1F, W2s + (buf ^ 1) * W2F, vec + (buf ^ 1) * VECF);
    cp_commit();

    // conv1 projection + scores  (GraphModels.py:464)
    stamp();
#if RES_USE_MMA
    project_rows_mma<NC, 2 * NC, 2, LDX, LDX>(xs, W1, vc, vc + 2 * NC, h1 + ro * 2 * NC, ss1 + ro * 2, sd1 + ro * 2, scr, n);
#else
    project_rows<NC, 2 * NC, 2, LDX, LDX>(xs, W1, vc, vc + 2 * NC, h1 + ro * 2 * NC, ss1 + ro * 2, sd1 + ro * 2, n);
#endif
    stamp();
    cluster_sync();
    stamp();
    // conv1 aggregation + bias + ReLU -> y1
    agg_fwd_rows<2, TRAIN>(rp_s, col_s, h1 + rb * 2 * NC, ss1 + rb * 2, sd1 + ro * 2, vc + 4 * NC,
                           TRAIN ? a.saved + sl.y1(k) + ro * 2 * NC : nullptr, ys, LDY,
                           TRAIN ? a.saved + sl.m1(k) + ro * 2 : nullptr, TRAIN ? a.saved + sl.l1(k) + ro * 2 : nullptr,
                           lo, n, true);
    __syncthreads();
    stamp();
    // conv2 projection + scores  (:465)
#if RES_USE_MMA
    project_rows_mma<2 * NC, NC, 1, LDY, LDY>(ys, W2, vc + 6 * NC, vc + 7 * NC, h2 + ro * NC, ss2 + ro, sd2 + ro, scr, n);
#else
    project_rows<2 * NC, NC, 1, LDY, LDY>(ys, W2, vc + 6 * NC, vc + 7 * NC, h2 + ro * NC, ss2 + ro, sd2 + ro, n);
#endif
    stamp();
    cluster_sync();
    stamp();
    // conv2 aggregation + bias -> z (neighbours read it in the mean)
    agg_fwd_rows<1, TRAIN>(rp_s, col_s, h2 + rb * NC, ss2 + rb, sd2 + ro, vc + 8 * NC, z + ro * NC, nullptr, 0,
                           TRAIN ? a.saved + sl.m2(k) + ro : nullptr, TRAIN ? a.saved + sl.l2(k) + ro : nullptr, lo, n,
                           false);
    stamp();
    cluster_sync();
    stamp();
    // SimpleConv(mean) + residual + ReLU  (:466-467)
    {
      const int lig = lane & 7, sub = lane >> 3;
      const float* zs = z + rb * NC;
      for (int il = warp * 4 + sub; il < n; il += T / 8) {
        const int i = lo + il;
        const int beg = rp_s[i], end = rp_s[i + 1] - 1;
        float4 acc = f4zero();
#pragma unroll 4
        for (int e = beg; e < end; ++e) add4(acc, ldc4(zs + (size_t)col_s[e] * NC + 4 * lig));
        const int deg = end - beg;
        const float inv = 1.f / (float)(deg > 1 ? deg : 1);
        const float4 xr = lds4(xs + il * LDX + 4 * lig);
        const float4 o = relu4(make_float4(fmaf(acc.x, inv, xr.x), fmaf(acc.y, inv, xr.y), fmaf(acc.z, inv, xr.z), fmaf(acc.w, inv, xr.w)));
        st4(xs + il * LDX + 4 * lig, o);
        if (TRAIN) st4(a.saved + sl.xout(k) + (ro + il) * NC + 4 * lig, o);
      }
    }
  }
  cp_wait_all();
  __syncthreads();
  // decoder Linear(nc, 1)  (:492)
  {
    const int lig = lane & 7, sub = lane >> 3;
    const float4 wv = ldg4(a.params + pl.lin1_w() + 4 * lig);
    const float bias = __ldg(a.params + pl.lin1_b());
    const bool bad = a.poison != nullptr && __ldg(a.poison) != 0;
    for (int i0 = 0; i0 < n; i0 += T / 8) {
      const int il = i0 + warp * 4 + sub;
      float p = il < n ? dot4(lds4(xs + il * LDX + 4 * lig), wv) : 0.f;
      p = group_sum<8>(p, FULL);
      if (il < n && lig == 0) a.out[ro + il] = bad ? __int_as_float(0x7fc00000) : p + bias;
    }
  }
}

// =============================================================================== backward
__global__ void __launch_bounds__(T, RES_MIN_CTAS)
resident_bwd_kernel(const Args a) {
  extern __shared__ __align__(16) float smem[];
  const int R = a.R, N = a.N, nb = a.nb;
  float* xs = smem;                        // [R][LDX] x0 of the block (own rows)
  float* ys = xs + R * LDX;                // [R][LDY] y1 -> dy1 -> dh1 (own rows)
  float* d2s = ys + R * LDY;               // [R][LDX] dh2 (own rows)
  float* W1s = d2s + R * LDX;              // [2][64][LDX]
  float* W2s = W1s + 2 * W1F;              // [2][32][LDY]
  float* vec = W2s + 2 * W2F;              // [2][288]
  float* vred = vec + 2 * VECF;            // [8 warps][288] parameter-vector gradient partials
  float* red = vred + (T / 32) * VECF;     // [T*4] scratch of cta_chunk_sum_atomic
  int* rp_s = reinterpret_cast<int*>(red + T * 4);        // shared CSR pair of the water network
  int* col_s = rp_s + a4(N + 1);
  int* rpt_s = col_s + a4(a.E1);
  int* colt_s = rpt_s + a4(N + 1);
  const int rank = (int)cluster_ctarank();
  const long long b = cluster_id_x();
  const int lo = rank * R, n = max(0, min(R, N - lo));
  const long long M = a.M, rb = b * N, ro = rb + lo;
  const ParamLayout pl(nb, NC);
  const SavedLayout sl(M, NC);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // scratch carve-up (gatres_scratch_floats, training = 1): gbuf[2] dz (dh2) dy1 (dh1) rec dsd
  float* gbuf[2] = {a.scratch, a.scratch + M * NC};
  float* dz = a.scratch + 2 * M * NC;
  float* rec2 = a.scratch + 3 * M * NC;            // dh2 region: rec of conv2 [M][4] + ds_dst of conv2 [M]
  float* dsd2 = rec2 + 4 * M;
  float* dy1 = a.scratch + 4 * M * NC;
  float* rec1 = a.scratch + 8 * M * NC;            // [M][2][4]
  float* dsd1 = rec1 + 8 * M;                      // [M][2]

  for (int c = threadIdx.x; c <= N; c += T) {
    rp_s[c] = __ldg(a.rowptr + c);
    rpt_s[c] = __ldg(a.rowptr_t + c);
  }
  for (int c = threadIdx.x; c < a.E1; c += T) {
    col_s[c] = __ldg(a.col + c);
    colt_s[c] = __ldg(a.col_t + c);
  }
  pdl_wait();
  Stamper stamp(a);
  stamp();
  const int k_first = nb > 0 ? a.k_hi : -1;
  // cp.async groups, oldest first: [parameters of block k] [saved y1 / x0 rows of block k] [parameters of block k - 1]
  if (nb > 0) stage_block_params(a.params + pl.block(k_first), W1s + (k_first & 1) * W1F, W2s + (k_first & 1) * W2F, vec + (k_first & 1) * VECF);
  cp_commit();
  if (nb > 0) {
    stage_rows<2 * NC, LDY>(a.saved + sl.y1(k_first) + ro * 2 * NC, ys, n);
    stage_rows<NC, LDX>(a.saved + (k_first > 0 ? sl.xout(k_first - 1) : sl.x_enc()) + ro * NC, xs, n);
  }
  cp_commit();

  if (a.head) {
    // decoder backward: g[i][c] = d_out[i] w[c] (masked by the last ReLU), dw = sum d_out x, db = sum d_out
    const int lig = lane & 7, sub = lane >> 3;
    const float* xl = a.saved + (nb > 0 ? sl.xout(nb - 1) : sl.x_enc());
    const float4 wv = ldg4(a.params + pl.lin1_w() + 4 * lig);
    float4 aw = f4zero();
    float ab = 0.f;
    for (int il = warp * 4 + sub; il < n; il += T / 8) {
      const float gv = __ldg(a.d_out + ro + il);
      const float4 xv = ldg4(xl + (ro + il) * NC + 4 * lig);
      float4 d = make_float4(gv * wv.x, gv * wv.y, gv * wv.z, gv * wv.w);
      if (nb > 0) d = mask4(d, xv);
      st4(gbuf[0] + (ro + il) * NC + 4 * lig, d);
      fma4(aw, gv, xv);
      if (lig == 0) ab += gv;
    }
    cta_chunk_sum_atomic<8>(aw, red, a.grads + pl.lin1_w());
    ab = group_sum<32>(ab, FULL);
    if (lane == 0 && ab != 0.f) atomicAdd(a.grads + pl.lin1_b(), ab);
  }
  cluster_sync();                            // gA of the first block is visible cluster-wide

  for (int k = k_first; k >= a.k_lo && k >= 0; --k) {
    const int buf = k & 1;
    const float* W1 = W1s + buf * W1F;
    const float* W2 = W2s + buf * W2F;
    const float* vc = vec + buf * VECF;
    float* gA = gbuf[(nb - 1 - k) & 1];
    float* gB = gbuf[(nb - k) & 1];
    const float* sv = a.saved;

    stamp();
    // (1) SimpleConv(mean) backward fused with conv2 pass 1 (incoming gradient dz stays in registers)
    bwd_p1_rows<1, true>(rp_s, col_s, rpt_s, colt_s, gA + rb * NC, dz + ro * NC, nullptr, 0,
                         sv + sl.h2(k) + rb * NC, sv + sl.ss2(k) + rb, sv + sl.sd2(k) + ro, sv + sl.m2(k) + ro,
                         sv + sl.l2(k) + ro, rec2 + ro * 4, dsd2 + ro, vred + warp * VECF + 8 * NC, lo, n);
    cp_wait_but_one();                       // this block's parameters have landed (its y1 / x0 rows may still fly) ...
    stamp();
    cluster_sync();                          // ... and are visible CTA-wide; dz / rec2 / dsd2 cluster-wide
    stamp();
    if (k - 1 >= a.k_lo && k - 1 >= 0)
      stage_block_params(a.params + pl.block(k - 1), W1s + (buf ^ 1) * W1F, W2s + (buf ^ 1) * W2F, vec + (buf ^ 1) * VECF);
    cp_commit();
    // (2) conv2 pass 2 -> dh2 (shared)
    bwd_p2_rows<1>(rpt_s, colt_s, dz + rb * NC, rec2 + rb * 4, dsd2 + ro, sv + sl.h2(k) + ro * NC,
                   sv + sl.ss2(k) + ro, vc + 6 * NC, vc + 7 * NC, d2s, LDX, vred + warp * VECF + 6 * NC,
                   vred + warp * VECF + 7 * NC, lo, n);
    if (k - 1 >= a.k_lo && k - 1 >= 0) cp_wait_but_one();      // y1 / x0 rows of this block (the newest group is the
    else cp_wait_all();                                        //  next block's parameters, if there is a next block)
    __syncthreads();                         // dh2, y1 and x0 are in shared memory
    stamp();
    // (3) conv2 projection backward: dW2 = dh2^T y1 ; dy1 = (dh2 W2) masked by y1 > 0 (in place over y1)
#if RES_USE_MMA
    wgrad_rows_mma<NC, 2 * NC, LDX, LDY>(d2s, ys, n, a.grads + pl.c2_W(k));
    __syncthreads();
    {
      float* dy1o = dy1 + ro * 2 * NC;
      dgrad_rows_mma<NC, 2 * NC, LDX, LDY>(d2s, W2, n, [](int, int) { return make_float2(0.f, 0.f); }, [&](int m, int c, float2 v, float2) {
        float2* p = reinterpret_cast<float2*>(ys + m * LDY + c);
        const float2 y = *p;
        v = make_float2(y.x > 0.f ? v.x : 0.f, y.y > 0.f ? v.y : 0.f);
        *p = v;
        *reinterpret_cast<float2*>(dy1o + (size_t)m * 2 * NC + c) = v;
      });
    }
#else
    wgrad_rows<NC, 2 * NC, LDX, LDY>(d2s, ys, n, a.grads + pl.c2_W(k));
    __syncthreads();
    {
      float* dy1o = dy1 + ro * 2 * NC;
      dgrad_rows<NC, 2 * NC, LDX, LDY>(d2s, W2, n, [&](int m, int c, float4 v) {
        float* p = ys + m * LDY + c;
        v = mask4(v, lds4(p));
        st4(p, v);
        st4(dy1o + (size_t)m * 2 * NC + c, v);
      });
    }
#endif
    __syncthreads();
    stamp();
    // (4) conv1 pass 1 (incoming gradient = dy1 from shared memory)
    bwd_p1_rows<2, false>(rp_s, col_s, rpt_s, colt_s, nullptr, nullptr, ys, LDY, sv + sl.h1(k) + rb * 2 * NC,
                          sv + sl.ss1(k) + rb * 2, sv + sl.sd1(k) + ro * 2, sv + sl.m1(k) + ro * 2,
                          sv + sl.l1(k) + ro * 2, rec1 + ro * 8, dsd1 + ro * 2, vred + warp * VECF + 4 * NC, lo, n);
    stamp();
    cluster_sync();
    stamp();
    // (5) conv1 pass 2 -> dh1 (shared, over the dy1 tile)
    bwd_p2_rows<2>(rpt_s, colt_s, dy1 + rb * 2 * NC, rec1 + rb * 8, dsd1 + ro * 2, sv + sl.h1(k) + ro * 2 * NC,
                   sv + sl.ss1(k) + ro * 2, vc, vc + 2 * NC, ys, LDY, vred + warp * VECF, vred + warp * VECF + 2 * NC,
                   lo, n);
    __syncthreads();
    stamp();
    // (6) conv1 projection backward: dW1 = dh1^T x0 ; gB = dh1 W1 + gA (residual), masked by x0 > 0 for k > 0
#if RES_USE_MMA
    wgrad_rows_mma<2 * NC, NC, LDY, LDX>(ys, xs, n, a.grads + pl.c1_W(k));
    {
      const float* gAo = gA + ro * NC;
      float* gBo = gB + ro * NC;
      const bool mask = k > 0;
      dgrad_rows_mma<2 * NC, NC, LDY, LDX>(
          ys, W1, n,
          [&](int m, int c) {                 // residual gradient of this fragment: requested before the MMAs
            float2 r;
            asm volatile("ld.global.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(gAo + (size_t)m * NC + c));
            return r;
          },
          [&](int m, int c, float2 v, float2 r) {
            v.x += r.x;
            v.y += r.y;
            if (mask) {
              const float2 x0 = lds2(xs + m * LDX + c);
              v = make_float2(x0.x > 0.f ? v.x : 0.f, x0.y > 0.f ? v.y : 0.f);
            }
            *reinterpret_cast<float2*>(gBo + (size_t)m * NC + c) = v;
          });
    }
#else
    wgrad_rows<2 * NC, NC, LDY, LDX>(ys, xs, n, a.grads + pl.c1_W(k));
    {
      const float* gAo = gA + ro * NC;
      float* gBo = gB + ro * NC;
      const bool mask = k > 0;
      dgrad_rows<2 * NC, NC, LDY, LDX>(ys, W1, n, [&](int m, int c, float4 v) {
        add4(v, ldc4(gAo + (size_t)m * NC + c));
        if (mask) v = mask4(v, lds4(xs + m * LDX + c));
        st4(gBo + (size_t)m * NC + c, v);
      });
    }
#endif
    // parameter-vector gradients of the block: sum the 8 warp rows, one 128-bit red per chunk
    for (int c = threadIdx.x; c < VECF / 4; c += T) {
      float4 s = f4zero();
#pragma unroll
      for (int w = 0; w < T / 32; ++w) add4(s, lds4(vred + w * VECF + 4 * c));
      const long long off = c < 6 * NC / 4 ? pl.c1_as(k) + 4 * c : pl.c2_as(k) + 4 * (c - 6 * NC / 4);
      if (n > 0) atomicAdd(reinterpret_cast<float4*>(a.grads + off), s);
    }
    __syncthreads();                         // xs / ys / vred are free again
    if (k - 1 >= a.k_lo && k - 1 >= 0) {
      stage_rows<2 * NC, LDY>(sv + sl.y1(k - 1) + ro * 2 * NC, ys, n);
      stage_rows<NC, LDX>(sv + (k - 1 > 0 ? sl.xout(k - 2) : sl.x_enc()) + ro * NC, xs, n);
    }
    cp_commit();
    stamp();
    cluster_sync();                          // gB is visible to the next block's gathers
  }
  cp_wait_all();

  if (a.tail) {
    // encoder backward: dw[c] = sum_i g[i][c] x[i], db[c] = sum_i g[i][c]
    const int lig = lane & 7, sub = lane >> 3;
    const float* g = gbuf[nb & 1];
    float4 aw = f4zero(), ab = f4zero();
    for (int il = warp * 4 + sub; il < n; il += T / 8) {
      const float4 gv = ldc4(g + (ro + il) * NC + 4 * lig);
      fma4(aw, __ldg(a.x + ro + il), gv);
      add4(ab, gv);
    }
    cta_chunk_sum_atomic<8>(aw, red, a.grads + pl.lin0_w());
    cta_chunk_sum_atomic<8>(ab, red, a.grads + pl.lin0_b());
  }
}

// ------------------------------------------------------------------------- host side
static size_t csr_ints(int N, int E1) { return (size_t)a4(N + 1) + (size_t)a4(E1); }
static size_t fwd_smem(int R, int N, int E1) {
  return sizeof(float) * ((size_t)R * (LDX + LDY) + 2 * (W1F + W2F + VECF) + a4(4 * R) + csr_ints(N, E1));
}
static size_t bwd_smem(int R, int N, int E1) {
  return sizeof(float) * ((size_t)R * (2 * LDX + LDY) + 2 * (W1F + W2F + VECF) + (T / 32) * VECF + T * 4 + 2 * csr_ints(N, E1));
}

// cluster size: as many CTAs per snapshot as keeps the whole batch co-resident (RES_MIN_CTAS CTAs per SM),
// power of two <= 8; grown again if the per-CTA row slice would not fit shared memory
static int pick_cluster(long long B, int N, int E1, size_t (*smem_of)(int, int, int)) {
  int cs = 8;
  if (res::forced_cluster() > 0) cs = res::forced_cluster();
  else
    while (cs > 1 && B * cs > (long long)RES_MIN_CTAS * sm_count()) cs >>= 1;
  while (cs < 8 && smem_of((N + cs - 1) / cs, N, E1) > res::kMaxSmem) cs <<= 1;
  return cs;
}

template <void (*kern)(const Args)>
static int launch_cluster(const char* what, int cs, long long B, size_t smem, cudaStream_t st, const Args& a) {
  static size_t configured = 0;             // one instance per kernel (the kernel is a template argument)
  if (smem > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch(what);
    configured = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * cs));
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  count_launch();
  cudaLaunchKernelEx(&cfg, kern, a);
  return check_launch(what);
}


static int forward(const gatres_model_desc* d, const float* params, const float* x, float* out, float* saved,
                     float* scratch, cudaStream_t st) {
  Args a = {};
  a.rowptr = d->rowptr; a.col = d->col; a.rowptr_t = d->rowptr_t; a.col_t = d->col_t;
  a.params = params; a.x = x; a.out = out; a.saved = saved; a.scratch = scratch; a.poison = d->poison;
  a.M = d->B * (long long)d->N; a.N = d->N; a.nb = d->num_blocks; a.B = (int)d->B; a.E1 = d->E1; a.prof = res::g_prof; a.prof_slots = res::g_prof_slots;
  const int cs = pick_cluster(d->B, d->N, d->E1, fwd_smem);
  a.R = (d->N + cs - 1) / cs;
  const size_t smem = fwd_smem(a.R, d->N, d->E1);
  return saved != nullptr ? launch_cluster<resident_fwd_kernel<true>>("resident_forward(train)", cs, d->B, smem, st, a)
                          : launch_cluster<resident_fwd_kernel<false>>("resident_forward", cs, d->B, smem, st, a);
}

static bool fits(int N, int E1, bool bwd) {
  const int R = (N + 7) / 8;
  return (bwd ? bwd_smem(R, N, E1) : fwd_smem(R, N, E1)) <= res::kMaxSmem;
}

static int backward(const gatres_model_desc* d, const float* params, const float* x, const float* saved,
                      const float* d_out, float* grads, float* scratch, int k_hi, int k_lo, bool head, bool tail,
                      cudaStream_t st) {
  Args a = {};
  a.rowptr = d->rowptr; a.col = d->col; a.rowptr_t = d->rowptr_t; a.col_t = d->col_t;
  a.params = params; a.x = x; a.saved = const_cast<float*>(saved); a.scratch = scratch; a.d_out = d_out; a.grads = grads;
  a.M = d->B * (long long)d->N; a.N = d->N; a.nb = d->num_blocks; a.B = (int)d->B; a.E1 = d->E1; a.prof = res::g_prof; a.prof_slots = res::g_prof_slots;
  a.k_hi = k_hi; a.k_lo = k_lo; a.head = head; a.tail = tail;
  const int cs = pick_cluster(d->B, d->N, d->E1, bwd_smem);
  a.R = (d->N + cs - 1) / cs;
  return launch_cluster<resident_bwd_kernel>("resident_backward", cs, d->B, bwd_smem(a.R, d->N, d->E1), st, a);
}


}  // namespace RES_NS
}  // namespace gatres
