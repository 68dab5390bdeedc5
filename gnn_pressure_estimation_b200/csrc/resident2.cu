// Snapshot-resident GATRes stacks, second generation: the tensors neighbours must see live in DISTRIBUTED SHARED
// MEMORY (sm_100a thread-block clusters), not in L2.
//
// Same decomposition as resident.cu — one cluster per snapshot carries all blocks of GATResMeanConv.forward
// (/root/reference/gnn_pressure_estimation/GraphModels.py:486-494; block :462-468) and of its autograd backward
// (train.py:185), each CTA owning a contiguous slice of the rows — but:
//   * the rows are numbered by a locality order (recursive spectral bisection of the water network, computed once
//     per template on the host: graph.py) so that ~85-90 % of a row's neighbours belong to the same CTA;
//   * h1 / h2 / z (forward) and the gradient rows and per-row softmax records (backward) are written to the
//     owner's shared memory only; a neighbour row is read with ld.shared::cluster through mapa — a plain shared
//     load (~40 cycles) when the owner is this CTA, a DSMEM load (~215 cycles) otherwise — instead of an L2 round
//     trip under load after every cluster barrier invalidated L1;
//   * saved activations for the backward are streamed to HBM AFTER the barrier arrival that follows their
//     production, so the release fence of a barrier never waits for them.
// The first generation measured 1.0 us per cluster barrier (MEMBAR.ALL.GPU waiting on the global exchange stores +
// CCTL.IVALL) and ~1 us per row pass of L2 gathers (profiles/r1_resident.md).
//
// Arithmetic is identical to resident.cu / the layer kernels: 3xTF32 mma.sync projections, cooperative per-row
// softmax with the heads packed per lane, in-row summation in the reference's edge order (ascending original source
// id, self-loop last — the permuted CSR keeps each row's entry order), recompute backward, atomic parameter gradients.
// The saved-activation buffer uses the locality row order, so the two kernels of this file are always used as a pair.
#include <math_constants.h>
#include <stdlib.h>
#include "common.cuh"
#include "layout.cuh"
#include "umma.cuh"

namespace gatres {
namespace res2 {

constexpr int NC = 32;
constexpr int T = 256;             // threads per CTA (8 warps: the mma tiling below assumes it)
constexpr int LDX = NC + 4;        // padded row of a [*, 32] shared tile
constexpr int LDY = 2 * NC + 4;    // padded row of a [*, 64] shared tile
constexpr int W1F = 2 * NC * LDX;  // conv1 weight [64][32] padded
constexpr int W2F = NC * LDY;      // conv2 weight [32][64] padded
constexpr int VECF = 9 * NC;       // as1 ad1 b1 (64 each) as2 ad2 b2 (32 each)
constexpr unsigned FULL = 0xffffffffu;
constexpr size_t kMaxSmem = 113 * 1024;     // two CTAs per SM (228 KB per SM, 1 KB reserved per CTA)

struct Args {
  const int* rowptr;     // in-edge CSR in LOCALITY numbering (self-loop last in every row)
  const int* col;
  const int* rowptr_t;   // out-edge CSR, same numbering
  const int* col_t;
  const int* perm;       // locality row -> original node id
  const float* params;
  const float* x;        // [M] model input, original row order
  float* out;            // fwd: [M], original row order
  float* saved;          // training activations (locality row order) or NULL
  const float* d_out;    // bwd: [M], original row order
  float* grads;
  float* scratch;        // bwd: running gradient between the ranges of a split backward ([M][32], locality order)
  const int* poison;
  long long M;
  int N, nb, R, ecap;    // ecap: shared-memory capacity (ints) of one CTA's slice of a CSR (max over the cluster)
  int k_hi, k_lo, head, tail;
  int barrier;           // ClusterBarrier flavour
  int cs;                // CTAs per snapshot (cluster size)
  int images;            // saved activations in the CTA-image layout (ImageLayout) instead of SavedLayout
  long long* prof;       // optional phase-timestamp buffer [CTA][prof_slots] (tools/resident_probe.py), else NULL
  int prof_slots;
};

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

// ---- cluster / distributed shared memory primitives ---------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
// Cluster barrier flavours (GATRES_RES2_BARRIER / gatres_set_resident_barrier):
//   0: every thread arrives with .release — the formally fenced variant.  On sm_100a the cluster-scope release is a
//      MEMBAR.ALL.GPU per thread: a round trip to L2 (~0.4 us measured with nothing outstanding) that also waits for
//      the thread's outstanding GLOBAL stores / atomics / cp.async, none of which the exchange needs.
//   1: __syncthreads, then warp 0 arrives with .release (cumulative over what the CTA barrier ordered before it) and
//      the other warps with .relaxed.  Measured no faster than 0 (the fence drains the SM's queue, not the warp's).
//   2 (default): membar.cta + __syncthreads + .relaxed arrival of every thread.  What crosses CTAs here is SHARED
//      memory only: the owner's stores are performed in its SM's shared memory once the CTA-scope fence and the CTA
//      barrier have drained them (B300_MICROARCH.md: BAR.SYNC drains pending STS), and a remote ld.shared::cluster
//      reads that same SRAM through the SM-to-SM network after the cluster barrier completed — there is no cache in
//      between that a gpu-scope fence would have to flush.  The PTX memory model does not name this case (a relaxed
//      arrival does not synchronise formally); flavour 0 stays available and both are covered by the parity tests.
//      Measured at 32 snapshots: forward barriers 0.85 -> 0.43 us each, training step 537 -> 516 us.
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
struct ClusterBarrier {
  int flavour;
  __device__ __forceinline__ void arrive() const {
    if (flavour == 0) { cluster_arrive_release(); return; }
    if (flavour == 2) __threadfence_block();
    __syncthreads();
    if (flavour == 2 || threadIdx.x >= 32) cluster_arrive_relaxed();
    else cluster_arrive_release();
  }
};
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// address of `local_addr` (a shared-window address of THIS CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ unsigned mapa(unsigned local_addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ldsc4(unsigned a) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 ldsc2(unsigned a) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float ldsc1(unsigned a) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
// a CSR entry in shared memory: owner CTA in the high half, row inside the owner's slice in the low half
__device__ __forceinline__ unsigned row_addr(unsigned base, int packed, int ld_bytes) {
  return mapa(base + (unsigned)(packed & 0xffff) * (unsigned)ld_bytes, (unsigned)packed >> 16);
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void cp16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp4(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void st4_stream(float* p, float4 v) {        // saved activations: written once, read by another kernel
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

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
__device__ __forceinline__ void stage_scalars(const float* g, float* s, int cnt) {      // g 16-byte aligned, cnt floats
  for (int c = threadIdx.x; c < (cnt + 3) / 4; c += T) cp16(s + 4 * c, g + 4 * c);
}
// padded shared tile -> own rows of a row-major global tensor (streaming stores)
template <int F, int LD>
__device__ __forceinline__ void store_rows(const float* s, float* g, int n) {
  for (int c = threadIdx.x; c < n * (F / 4); c += T) st4_stream(g + 4 * c, lds4(s + (c / (F / 4)) * LD + 4 * (c % (F / 4))));
}
__device__ __forceinline__ void store_scalars(const float* s, float* g, int cnt) {
  for (int c = threadIdx.x; c < cnt; c += T) g[c] = s[c];
}

// this CTA's slice of a CSR -> shared memory: row pointers relative to the slice, entries packed (owner << 16 | row)
__device__ __forceinline__ void stage_csr_slice(const int* __restrict__ rowptr, const int* __restrict__ col, int lo, int n,
                                                int R, int* rp_s, int* col_s, int ecap) {
  const int e_lo = __ldg(rowptr + lo), cnt = min(__ldg(rowptr + lo + n) - e_lo, ecap);
  for (int c = threadIdx.x; c <= n; c += T) rp_s[c] = __ldg(rowptr + lo + c) - e_lo;
  for (int c = threadIdx.x; c < cnt; c += T) {
    const int j = __ldg(col + e_lo + c);
    const int owner = j / R;
    col_s[c] = (owner << 16) | (j - owner * R);
  }
}

// ---- tensor-core contractions (mma.sync m16n8k8 TF32, 3xTF32 error compensated; see resident_mma.cuh) --------------
__device__ __forceinline__ unsigned tf32_lo_bits(float x) {
  const float r = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  return (__float_as_uint(r) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
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

// h[m][n] = sum_k A[m][k] W[n][k] into a padded SHARED tile (row stride LDO) + attention scores into shared arrays.
// Warp w: 16-row tile (w % 4) of every 64-row group, column half w / 4 (for H = 2 the half is the head).
template <int K, int NOUT, int H, int LDA, int LDW, int LDO>
__device__ __forceinline__ void project_mma(const float* As, const float* Ws, const float* att_s, const float* att_d,
                                            float* h_s, float* ss_s, float* sd_s, float* scr, int n) {
  constexpr int NTW = NOUT / 16;
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
    float ps[2] = {0.f, 0.f}, pd[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
      const int c = nbase + 8 * j + 2 * t;
      const float s0 = att_s[c], s1 = att_s[c + 1], d0 = att_d[c], d1 = att_d[c + 1];
      ps[0] = fmaf(acc[j][0], s0, fmaf(acc[j][1], s1, ps[0]));
      ps[1] = fmaf(acc[j][2], s0, fmaf(acc[j][3], s1, ps[1]));
      pd[0] = fmaf(acc[j][0], d0, fmaf(acc[j][1], d1, pd[0]));
      pd[1] = fmaf(acc[j][2], d0, fmaf(acc[j][3], d1, pd[1]));
      if (m0 + g < n) *reinterpret_cast<float2*>(h_s + (m0 + g) * LDO + c) = make_float2(acc[j][0], acc[j][1]);
      if (m0 + g + 8 < n) *reinterpret_cast<float2*>(h_s + (m0 + g + 8) * LDO + c) = make_float2(acc[j][2], acc[j][3]);
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
            ss_s[m * 2 + half] = ps[i];
            sd_s[m * 2 + half] = pd[i];
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
      ss_s[m] = scr[0 * n + m] + scr[2 * n + m];
      sd_s[m] = scr[1 * n + m] + scr[3 * n + m];
    }
  }
}

// ---- fused GAT aggregation over this CTA's rows (C = 32), neighbour rows through distributed shared memory ---------
// FOUR lanes own a row, so a warp covers eight rows and the 49-row slice of a CTA is ONE pass of its eight warps (the
// phases are bound by the length of a warp's dependent instruction chain, not by issue slots: one chain instead of
// two).  Lane `slot` holds the float4 chunks {slot, slot + 4} of EVERY head of the row and evaluates the edges
// {slot, slot + 4} of each group of eight in-edges for every head (logit, LeakyReLU, exp); row max / sum are two-level
// butterflies over the four lanes; the accumulation loop broadcasts (neighbour, weight) with two shuffles per edge.
// h_base / ss_base: shared-window addresses of the h tile (row stride ldh bytes) and of the source-score array
// ([row][H]) — the same offsets in every CTA of the cluster.
__device__ __forceinline__ float gmax4(float v) {
  v = fmaxf(v, __shfl_xor_sync(FULL, v, 2));
  return fmaxf(v, __shfl_xor_sync(FULL, v, 1));
}
template <int H>
__device__ __forceinline__ void load_scores(unsigned ss_base, int packed, float (&s)[H]) {
  const unsigned a = row_addr(ss_base, packed, 4 * H);
  if (H == 2) {
    const float2 t2 = ldsc2(a);
    s[0] = t2.x; s[H - 1] = t2.y;
  } else s[0] = ldsc1(a);
}

// float offset of the 16-byte chunk q (0..7) of row r in a [rows x 32] SWIZZLE_128B operand tile (rows of 128 B, chunk
// index XOR row % 8: the canonical K-major UMMA layout, and — read with rows as K — the MN-major one)
__device__ __forceinline__ int sw_chunk(int r, int q) { return r * 32 + ((q ^ (r & 7)) << 2); }
__device__ __forceinline__ float4 lo4(float4 x) { return make_float4(lo_tf32(x.x), lo_tf32(x.y), lo_tf32(x.z), lo_tf32(x.w)); }

// SW = false: out_s is a padded row-major tile (row stride ld_out floats).  SW = true: out_s is a tensor-core operand,
// one [rows x 32] SWIZZLE_128B tile per head (tile stride ld_out floats), and out_lo receives the 3xTF32 low parts.
template <int H, bool SW = false>
__device__ __forceinline__ void agg_fwd(const int* rp_s, const int* col_s, unsigned h_base, int ldh, unsigned ss_base,
                                        const float* sd_s, const float* bias_s, float* out_s, int ld_out,
                                        float* m_dst, float* l_dst, int n, int self_owner, bool relu,
                                        float* out_lo = nullptr) {
  constexpr int RPW = 8, PRE = H == 1 ? 4 : 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 2, slot = lane & 3;
  float4 bv[H][2];
#pragma unroll
  for (int v = 0; v < H; ++v)
#pragma unroll
    for (int c = 0; c < 2; ++c) bv[v][c] = lds4(bias_s + 32 * v + 16 * c + 4 * slot);
  for (int i0 = 0; i0 < n; i0 += (T / 32) * RPW) {
    const int il_raw = i0 + warp * RPW + sub;
    const bool ok = il_raw < n;
    const int il = ok ? il_raw : n - 1;
    const int beg = rp_s[il], deg = rp_s[il + 1] - beg;
    const int deg_max = __reduce_max_sync(FULL, deg);
    const int self = (self_owner << 16) | il;
    float sd[H], mrun[H], lrun[H];
    float4 acc[H][2];
#pragma unroll
    for (int v = 0; v < H; ++v) {
      sd[v] = sd_s[il * H + v];
      mrun[v] = -CUDART_INF_F;
      lrun[v] = 0.f;
      acc[v][0] = acc[v][1] = f4zero();
    }
    for (int e0 = 0; e0 < deg_max; e0 += 8) {
      const bool valid0 = e0 + slot < deg, valid1 = e0 + slot + 4 < deg;
      const int j0 = valid0 ? col_s[beg + e0 + slot] : self;
      const int j1 = valid1 ? col_s[beg + e0 + slot + 4] : self;
      const int cnt = min(8, deg - e0), cnt_max = min(8, deg_max - e0);
      float4 x[PRE][H][2];
#pragma unroll
      for (int u = 0; u < PRE; ++u) {
        const int ju = __shfl_sync(FULL, j0, u, 4);
        const unsigned au = row_addr(h_base, ju, ldh) + 16u * slot;
#pragma unroll
        for (int v = 0; v < H; ++v)
#pragma unroll
          for (int c = 0; c < 2; ++c) x[u][v][c] = u < cnt ? ldsc4(au + 128u * v + 64u * c) : f4zero();
      }
      float s0[H], s1[H], p0[H], p1[H];
      load_scores<H>(ss_base, j0, s0);
      if (cnt_max > 4) load_scores<H>(ss_base, j1, s1);
      else {
#pragma unroll
        for (int v = 0; v < H; ++v) s1[v] = 0.f;
      }
#pragma unroll
      for (int v = 0; v < H; ++v) {
        const float a0 = valid0 ? lrelu(s0[v] + sd[v]) : -CUDART_INF_F;
        const float a1 = valid1 ? lrelu(s1[v] + sd[v]) : -CUDART_INF_F;
        const float nm = fmaxf(mrun[v], gmax4(fmaxf(a0, a1)));
        if (e0 > 0) {
          const float sc = __expf(mrun[v] - nm);
          lrun[v] *= sc;
#pragma unroll
          for (int c = 0; c < 2; ++c) { acc[v][c].x *= sc; acc[v][c].y *= sc; acc[v][c].z *= sc; acc[v][c].w *= sc; }
        }
        p0[v] = __expf(a0 - nm);
        p1[v] = __expf(a1 - nm);
        lrun[v] += group_sum<4>(p0[v] + p1[v], FULL);
        mrun[v] = nm;
      }
#pragma unroll
      for (int u = 0; u < PRE; ++u)
#pragma unroll
        for (int v = 0; v < H; ++v) {
          const float pu = __shfl_sync(FULL, p0[v], u, 4);
          fma4(acc[v][0], pu, x[u][v][0]);
          fma4(acc[v][1], pu, x[u][v][1]);
        }
      for (int t = PRE; t < cnt_max; ++t) {
        const int jt = __shfl_sync(FULL, t < 4 ? j0 : j1, t & 3, 4);
        const unsigned at = row_addr(h_base, jt, ldh) + 16u * slot;
        const bool on = t < cnt;
#pragma unroll
        for (int v = 0; v < H; ++v) {
          const float4 xa = on ? ldsc4(at + 128u * v) : f4zero();
          const float4 xb = on ? ldsc4(at + 128u * v + 64u) : f4zero();
          const float pt = __shfl_sync(FULL, t < 4 ? p0[v] : p1[v], t & 3, 4);
          fma4(acc[v][0], pt, xa);
          fma4(acc[v][1], pt, xb);
        }
      }
    }
    if (!ok) continue;
#pragma unroll
    for (int v = 0; v < H; ++v) {
      const float inv = 1.f / (lrun[v] + kSoftmaxEps);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float4 o = make_float4(fmaf(acc[v][c].x, inv, bv[v][c].x), fmaf(acc[v][c].y, inv, bv[v][c].y),
                               fmaf(acc[v][c].z, inv, bv[v][c].z), fmaf(acc[v][c].w, inv, bv[v][c].w));
        if (relu) o = relu4(o);
        if (SW) {
          const int off = v * ld_out + sw_chunk(il, 4 * c + slot);
          st4(out_s + off, o);
          st4(out_lo + off, lo4(o));
        } else st4(out_s + il * ld_out + 32 * v + 16 * c + 4 * slot, o);
      }
      if (m_dst != nullptr && slot == 0) { m_dst[il * H + v] = mrun[v]; l_dst[il * H + v] = lrun[v]; }
    }
  }
}

// =============================================================================== forward
// Shared memory (floats): xs[R][LDX] h1s[R][LDY] ys[R][LDY] h2s[R][LDX] zs[R][LDX] ss1[2R] sd1[2R] ss2[R] sd2[R] ml1[4R] ml2[2R]
//                         W1s[2][W1F] W2s[2][W2F] vec[2][VECF] scr[4R] | ints: rp_s[R+1] col_s[ecap]
struct FwdSmem {
  int xs, h1s, ys, h2s, zs, ss1, sd1, ss2, sd2, ml1, ml2, W1s, W2s, vec, scr, rp, col, total;
  __host__ __device__ FwdSmem(int R, int ecap) {
    int o = 0;
    auto take = [&](int nfl) { const int at = o; o += (int)a4(nfl); return at; };
    xs = take(R * LDX); h1s = take(R * LDY); ys = take(R * LDY); h2s = take(R * LDX); zs = take(R * LDX);
    ss1 = take(2 * R); sd1 = take(2 * R); ss2 = take(R); sd2 = take(R); ml1 = take(4 * R); ml2 = take(2 * R);
    W1s = take(2 * W1F); W2s = take(2 * W2F); vec = take(2 * VECF); scr = take(4 * R);
    rp = take(R + 1); col = take(ecap);
    total = o;
  }
};

template <bool TRAIN>
__global__ void __launch_bounds__(T, 2)
fwd_kernel(const Args a) {
  extern __shared__ __align__(16) float smem[];
  const int R = a.R, N = a.N;
  const FwdSmem L(R, a.ecap);
  float* xs = smem + L.xs;
  float* h1s = smem + L.h1s;
  float* ys = smem + L.ys;
  float* h2s = smem + L.h2s;
  float* zs = smem + L.zs;
  float* ss1 = smem + L.ss1;
  float* sd1 = smem + L.sd1;
  float* ss2 = smem + L.ss2;
  float* sd2 = smem + L.sd2;
  float* ml1 = smem + L.ml1;
  float* ml2 = smem + L.ml2;
  float* W1s = smem + L.W1s;
  float* W2s = smem + L.W2s;
  float* vec = smem + L.vec;
  float* scr = smem + L.scr;
  int* rp_s = reinterpret_cast<int*>(smem + L.rp);
  int* col_s = reinterpret_cast<int*>(smem + L.col);
  const int rank = (int)cluster_ctarank();
  const ClusterBarrier cb = {a.barrier};
  const long long b = cluster_id_x();
  const int lo = rank * R, n = max(0, min(R, N - lo));
  const long long M = a.M, rb = b * N, ro = rb + lo;    // snapshot base row, first own row (locality order)
  const ParamLayout pl(a.nb, NC);
  const SavedLayout sl(M, NC);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // the topology is constant across steps: stage this CTA's slice before the dependency wait
  if (n > 0) stage_csr_slice(a.rowptr, a.col, lo, n, R, rp_s, col_s, a.ecap);
  pdl_wait();
  Stamper stamp(a);
  stamp();
  if (a.nb > 0) stage_block_params(a.params + pl.block(0), W1s, W2s, vec);
  cp_commit();

  // encoder Linear(1, nc): x0[i][c] = x[i] w[c] + b[c]   (GraphModels.py:487)
  {
    const int lig = lane & 7, sub = lane >> 3;
    const float4 wv = ldg4(a.params + pl.lin0_w() + 4 * lig), bv = ldg4(a.params + pl.lin0_b() + 4 * lig);
    for (int il = warp * 4 + sub; il < n; il += T / 8) {
      const float xv = __ldg(a.x + rb + __ldg(a.perm + lo + il));
      st4(xs + il * LDX + 4 * lig, make_float4(fmaf(xv, wv.x, bv.x), fmaf(xv, wv.y, bv.y), fmaf(xv, wv.z, bv.z), fmaf(xv, wv.w, bv.w)));
    }
  }
  const unsigned h1_base = smem_u32(h1s), h2_base = smem_u32(h2s), z_base = smem_u32(zs);
  const unsigned ss1_base = smem_u32(ss1), ss2_base = smem_u32(ss2);

  for (int k = 0; k < a.nb; ++k) {
    const int buf = k & 1;
    float* W1 = W1s + buf * W1F;
    float* W2 = W2s + buf * W2F;
    float* vc = vec + buf * VECF;
    cp_wait_all();
    __syncthreads();                         // parameters of block k and xs are in place

    // conv1 projection + scores  (GraphModels.py:464) -> own shared tiles
    stamp();
    project_mma<NC, 2 * NC, 2, LDX, LDX, LDY>(xs, W1, vc, vc + 2 * NC, h1s, ss1, sd1, scr, n);
    stamp();
    cb.arrive();
    // Saved activations go to HBM only in the shadow of a barrier wait, right AFTER an arrival: the release fence of
    // the next arrival then finds them long complete.  Here: the block input (encoder output / previous block's output).
    if (TRAIN) store_rows<NC, LDX>(xs, a.saved + (k > 0 ? sl.xout(k - 1) : sl.x_enc()) + ro * NC, n);
    cluster_wait();
    stamp();
    // conv1 aggregation + bias + ReLU -> y1 (row-local from here on); (m, l) wait in shared memory
    agg_fwd<2>(rp_s, col_s, h1_base, LDY * 4, ss1_base, sd1, vc + 4 * NC, ys, LDY, TRAIN ? ml1 : nullptr,
               TRAIN ? ml1 + 2 * R : nullptr, n, rank, true);
    __syncthreads();
    stamp();
    // conv2 projection + scores  (:465)
    project_mma<2 * NC, NC, 1, LDY, LDY, LDX>(ys, W2, vc + 6 * NC, vc + 7 * NC, h2s, ss2, sd2, scr, n);
    stamp();
    cb.arrive();
    // the next block's parameters are requested in the shadow of this barrier (their buffer was last read a block ago)
    if (k + 1 < a.nb) stage_block_params(a.params + pl.block(k + 1), W1s + (buf ^ 1) * W1F, W2s + (buf ^ 1) * W2F, vec + (buf ^ 1) * VECF);
    cp_commit();
    if (TRAIN) {                             // conv1's tensors: all final, none rewritten before the next block
      store_rows<2 * NC, LDY>(h1s, a.saved + sl.h1(k) + ro * 2 * NC, n);
      store_rows<2 * NC, LDY>(ys, a.saved + sl.y1(k) + ro * 2 * NC, n);
      store_scalars(ss1, a.saved + sl.ss1(k) + ro * 2, 2 * n);
      store_scalars(sd1, a.saved + sl.sd1(k) + ro * 2, 2 * n);
      store_scalars(ml1, a.saved + sl.m1(k) + ro * 2, 2 * n);
      store_scalars(ml1 + 2 * R, a.saved + sl.l1(k) + ro * 2, 2 * n);
    }
    cluster_wait();
    stamp();
    // conv2 aggregation + bias -> z (neighbours read it in the mean)
    agg_fwd<1>(rp_s, col_s, h2_base, LDX * 4, ss2_base, sd2, vc + 8 * NC, zs, LDX, TRAIN ? ml2 : nullptr,
               TRAIN ? ml2 + R : nullptr, n, rank, false);
    stamp();
    cb.arrive();
    if (TRAIN) {
      store_rows<NC, LDX>(h2s, a.saved + sl.h2(k) + ro * NC, n);
      store_scalars(ss2, a.saved + sl.ss2(k) + ro, n);
      store_scalars(sd2, a.saved + sl.sd2(k) + ro, n);
      __syncthreads();                       // (m, l) of conv2 were written by other warps just before the arrival
      store_scalars(ml2, a.saved + sl.m2(k) + ro, n);
      store_scalars(ml2 + R, a.saved + sl.l2(k) + ro, n);
    }
    cluster_wait();
    stamp();
    // SimpleConv(mean) + residual + ReLU  (:466-467): in-neighbours minus the trailing self-loop; four lanes per row
    {
      const int slot = lane & 3, sub = lane >> 2;
      for (int il = warp * 8 + sub; il < n; il += T / 4) {
        const int beg = rp_s[il], end = rp_s[il + 1] - 1;
        float4 acc0 = f4zero(), acc1 = f4zero();
#pragma unroll 4
        for (int e = beg; e < end; ++e) {
          const unsigned ar = row_addr(z_base, col_s[e], LDX * 4) + 16u * slot;
          add4(acc0, ldsc4(ar));
          add4(acc1, ldsc4(ar + 64u));
        }
        const int deg = end - beg;
        const float inv = 1.f / (float)(deg > 1 ? deg : 1);
        float* xr = xs + il * LDX + 4 * slot;
        const float4 r0 = lds4(xr), r1 = lds4(xr + 16);
        st4(xr, relu4(make_float4(fmaf(acc0.x, inv, r0.x), fmaf(acc0.y, inv, r0.y), fmaf(acc0.z, inv, r0.z), fmaf(acc0.w, inv, r0.w))));
        st4(xr + 16, relu4(make_float4(fmaf(acc1.x, inv, r1.x), fmaf(acc1.y, inv, r1.y), fmaf(acc1.z, inv, r1.z), fmaf(acc1.w, inv, r1.w))));
      }
    }
  }
  cp_wait_all();
  __syncthreads();
  // no CTA may leave while a neighbour still reads its shared memory
  cb.arrive();
  if (TRAIN && a.nb > 0) store_rows<NC, LDX>(xs, a.saved + sl.xout(a.nb - 1) + ro * NC, n);
  if (TRAIN && a.nb == 0) store_rows<NC, LDX>(xs, a.saved + sl.x_enc() + ro * NC, n);
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
      if (il < n && lig == 0) a.out[rb + __ldg(a.perm + lo + il)] = bad ? __int_as_float(0x7fc00000) : p + bias;
    }
  }
  cluster_wait();
}

// =============================================================================== saved activations as CTA images
// The tensor-core forward / backward pair keeps its saved activations as straight IMAGES of the forward's shared-memory
// arrays, one record per (block, snapshot, CTA): every array is contiguous in shared memory AND in the record, so the
// forward writes a block's activations with seven bulk copies (cp.async.bulk shared -> global, issued by one thread,
// carried out by the TMA engine) instead of ~3000 LDS + STG per CTA and block — those stores cost the training forward
// 2.3 us per block in the shadow of its barrier waits (153 us against 120 us for the same stack without them), and one
// bulk copy per ROW was worse still (per-request overhead: 205 us).  Record (floats, R rows per CTA):
//   x   [R][32]      block input, SWIZZLE_128B operand rows (16-byte chunk index XOR row % 8)
//   h1  [R][LDY]     conv1 projection, padded rows as the neighbours gather them
//   y1  2 x [R][32]  conv1 output, one SWIZZLE_128B operand tile per head
//   s1  4 x a4(2R)   s_src, s_dst, m, l of conv1 ([row][head])
//   h2  [R][LDX]     conv2 projection, padded rows
//   s2  4 x a4(R)    s_src, s_dst, m, l of conv2
// followed, after the last block, by one x image per CTA (the stack's output = the decoder's input).  ~5 % larger than
// SavedLayout (padding); gatres_saved_floats covers both.  Private to this file: the backward's loaders undo the
// swizzle / padding while they copy.
struct ImageLayout {
  long long B;
  int R, cs, off_h1, off_y, off_s1, off_h2, off_s2, rec;
  __host__ __device__ ImageLayout(int R_, long long B_, int cs_) : B(B_), R(R_), cs(cs_) {
    off_h1 = R * 32;
    off_y = off_h1 + R * LDY;
    off_s1 = off_y + 2 * R * 32;
    off_h2 = off_s1 + 4 * (int)a4(2 * R);
    off_s2 = off_h2 + R * LDX;
    rec = off_s2 + 4 * (int)a4(R);
  }
  __host__ __device__ long long record(int k, long long b, int rank) const { return ((k * B + b) * cs + rank) * (long long)rec; }
  __host__ __device__ long long tail(int nb, long long b, int rank) const { return nb * B * cs * (long long)rec + (b * cs + rank) * (long long)(R * 32); }
  __host__ __device__ long long total(int nb) const { return nb * B * cs * (long long)rec + B * cs * (long long)(R * 32); }
};
__device__ __forceinline__ void bulk_s2g(float* gdst, const float* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// =============================================================================== forward, tcgen05 projections
// Same stack as fwd_kernel, with the two projections of a block on the 5th-generation tensor cores: the mma.sync
// 3xTF32 contractions were the largest single item of a block (3.0 of 9.7 us: ~15 cycles per MMA per scheduler).
//   * the block input x and the conv1 output y1 — the A operands — live in shared memory as SWIZZLE_128B K-major
//     tiles (rows of 128 B, one tile per 32 columns) together with their 3xTF32 low parts, both written by the phase
//     that produces them (encoder / mean / conv1 aggregation);
//   * W1 / W2 are staged by cp.async straight into K-major SWIZZLE_128B tiles (single-buffered: the next block's
//     copy of a matrix is requested as soon as its MMAs have completed) and every thread derives the low parts of
//     the chunks it copied;
//   * one thread issues the 3 x K/8 tcgen05.mma (M = 128: the CTA's <= 64 rows are the first rows of the tile, the rest
//     of the datapath reads whatever follows in shared memory and its accumulator rows are never read), commits to
//     an mbarrier, and warps 0/1 (+ 4/5 for the second column half) drain the accumulator with tcgen05.ld — thread
//     = row, so the attention scores are in-thread dot products — into the padded tiles the neighbours gather from.
// Saved activations go to HBM as images of the shared-memory arrays (ImageLayout above): seven bulk copies per block.
constexpr int TC_COLS = 64;                  // TMEM columns per CTA (conv1: N = 64, conv2: N = 32)
#ifndef GATRES_TC_M
#define GATRES_TC_M 64
#endif
constexpr int TC_M = GATRES_TC_M;            // UMMA M: 64 (accumulator row r in TMEM lane 32 (r / 16) + r % 16: the CTA's <= 64 rows
                                             // spread over the four lane quarters, all eight warps drain, half the A reads) or 128

struct FwdTcSmem {
  int xa, xl, ya, yl, w1h, w1l, w2h, w2l, h1s, h2s, s1, s2, vec, rp, col, bar, total, RP;
  __host__ __device__ FwdTcSmem(int R, int ecap) {
    RP = (R + 7) & ~7;
    int o = 0;
    auto take = [&](int nfl) { const int at = o; o += (int)a4(nfl); return at; };
    xa = take(RP * 32); xl = take(RP * 32); ya = take(2 * RP * 32); yl = take(2 * RP * 32);       // 1024-byte multiples
    w1h = take(2 * NC * NC); w1l = take(2 * NC * NC); w2h = take(2 * NC * NC); w2l = take(2 * NC * NC);
    h1s = take(R * LDY); h2s = take(R * LDX);
    s1 = take(4 * (int)a4(2 * R)); s2 = take(4 * (int)a4(R));        // (s_src, s_dst, m, l) blocks: images of the saved records
    vec = take(2 * VECF); rp = take(R + 1); col = take(ecap); bar = take(4);
    total = o;
  }
};

// one block's W1 -> [64 x 32] K-major SWIZZLE_128B tile; W2 -> two [32 x 32] tiles (one per 32 input columns)
__device__ __forceinline__ void stage_w1_tc(const float* blk, float* W1H) {
  for (int c = threadIdx.x; c < 2 * NC * NC / 4; c += T) cp16(W1H + sw_chunk(c >> 3, c & 7), blk + 4 * c);
}
__device__ __forceinline__ void stage_w2_tc(const float* blk, float* W2H) {
  const float* w2 = blk + 2 * NC * NC + 6 * NC;
  for (int c = threadIdx.x; c < 2 * NC * NC / 4; c += T)
    cp16(W2H + ((c >> 3) & 1) * (NC * 32) + sw_chunk(c >> 4, c & 7), w2 + 4 * c);
}
__device__ __forceinline__ void stage_vec_tc(const float* blk, float* vec) {
  for (int c = threadIdx.x; c < 9 * NC / 4; c += T)
    cp16(vec + 4 * c, blk + (c < 6 * NC / 4 ? 2 * NC * NC + 4 * c : 4 * NC * NC + 4 * c));
}
// 3xTF32 low parts of the weight chunks THIS thread copied (same chunk -> thread map as the staging loops)
__device__ __forceinline__ void lo_own_weights(const float* W1H, float* W1L, const float* W2H, float* W2L) {
  for (int c = threadIdx.x; c < 2 * NC * NC / 4; c += T) {
    const int o1 = sw_chunk(c >> 3, c & 7), o2 = ((c >> 3) & 1) * (NC * 32) + sw_chunk(c >> 4, c & 7);
    st4(W1L + o1, lo4(lds4(W1H + o1)));
    st4(W2L + o2, lo4(lds4(W2H + o2)));
  }
}
// D[128 x NOUT] (TMEM) = A[128 x K] B[NOUT x K]^T, 3xTF32; A / B: SWIZZLE_128B K-major tiles, one per 32 columns of K
template <int K, int NOUT>
__device__ __forceinline__ void issue_proj_tc(uint32_t tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_tile_bytes,
                                              uint32_t b_hi, uint32_t b_lo, uint32_t b_tile_bytes) {
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
#pragma unroll
  for (int s = 0; s < K / 8; ++s) {
    const uint32_t ka = (uint32_t)(s >> 2) * a_tile_bytes + (uint32_t)(s & 3) * 32u;
    const uint32_t kb = (uint32_t)(s >> 2) * b_tile_bytes + (uint32_t)(s & 3) * 32u;
    umma_tf32(tmem, umma_desc_k128(a_hi + ka), umma_desc_k128(b_hi + kb), IDESC, s > 0);
    umma_tf32(tmem, umma_desc_k128(a_lo + ka), umma_desc_k128(b_hi + kb), IDESC, 1);
    umma_tf32(tmem, umma_desc_k128(a_hi + ka), umma_desc_k128(b_lo + kb), IDESC, 1);
  }
}
// accumulator row (thread = row) -> padded shared tile + the two attention scores of the row's 32 columns
__device__ __forceinline__ void drain_row_tc(uint32_t taddr, const float* att_s, const float* att_d, float* h_row,
                                             float* ss_dst, float* sd_dst, bool ok) {
  float v[32];
  tmem_ld32(taddr, v);
  if (!ok) return;
  float ps = 0.f, pd = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 s4 = lds4(att_s + 4 * q), d4 = lds4(att_d + 4 * q);
    ps = fmaf(v[4 * q], s4.x, fmaf(v[4 * q + 1], s4.y, fmaf(v[4 * q + 2], s4.z, fmaf(v[4 * q + 3], s4.w, ps))));
    pd = fmaf(v[4 * q], d4.x, fmaf(v[4 * q + 1], d4.y, fmaf(v[4 * q + 2], d4.z, fmaf(v[4 * q + 3], d4.w, pd))));
    st4(h_row + 4 * q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
  }
  *ss_dst = ps;
  *sd_dst = pd;
}
template <bool TRAIN>
__global__ void __launch_bounds__(T, 2)
fwd_tc_kernel(const Args a) {
  extern __shared__ __align__(1024) float smem_tc[];
  float* const smem = smem_tc;
  const int R = a.R, N = a.N;
  const FwdTcSmem L(R, a.ecap);
  const int TF = L.RP * 32;                  // floats per operand tile
  float* XA = smem + L.xa;
  float* XL = smem + L.xl;
  float* YA = smem + L.ya;
  float* YL = smem + L.yl;
  float* zs = YL;                            // conv2 output z [R][LDX] over the dead low part of y1
  float* W1H = smem + L.w1h;
  float* W1L = smem + L.w1l;
  float* W2H = smem + L.w2h;
  float* W2L = smem + L.w2l;
  float* h1s = smem + L.h1s;
  float* h2s = smem + L.h2s;
  float* ss1 = smem + L.s1;
  float* sd1 = ss1 + a4(2 * R);
  float* m1 = sd1 + a4(2 * R);
  float* l1 = m1 + a4(2 * R);
  float* ss2 = smem + L.s2;
  float* sd2 = ss2 + a4(R);
  float* m2 = sd2 + a4(R);
  float* l2 = m2 + a4(R);
  float* vec = smem + L.vec;
  int* rp_s = reinterpret_cast<int*>(smem + L.rp);
  int* col_s = reinterpret_cast<int*>(smem + L.col);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar + 2);
  if ((smem_u32(smem) & 1023u) != 0) __trap();            // SWIZZLE_128B atoms are 1024-byte aligned
  const int rank = (int)cluster_ctarank();
  const ClusterBarrier cb = {a.barrier};
  const long long b = cluster_id_x();
  const int lo = rank * R, n = max(0, min(R, N - lo));
  const long long M = a.M, rb = b * N;
  const ParamLayout pl(a.nb, NC);
  const ImageLayout im(R, M / N, a.cs);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // saved activations: thread 32 issues the bulk copies of a group after the barrier arrival that follows the last
  // write of its arrays (every writer has fenced them towards the async proxy before arriving)
  const bool storer = TRAIN && threadIdx.x == 32;
  auto after_arrival = [&]() {
    if (TRAIN && cb.flavour == 0) __syncthreads();          // flavour 0 arrives without a CTA barrier
  };

  if (warp == 0) tmem_alloc(tmem_slot, TC_COLS);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (n > 0) stage_csr_slice(a.rowptr, a.col, lo, n, R, rp_s, col_s, a.ecap);
  pdl_wait();
  Stamper stamp(a);
  stamp();
  if (a.nb > 0) {
    const float* blk = a.params + pl.block(0);
    stage_w1_tc(blk, W1H);
    stage_w2_tc(blk, W2H);
    stage_vec_tc(blk, vec);
  }
  cp_commit();

  // encoder Linear(1, nc)  (GraphModels.py:487) -> operand tile + low part
  {
    const int lig = lane & 7, sub = lane >> 3;
    const float4 wv = ldg4(a.params + pl.lin0_w() + 4 * lig), bv = ldg4(a.params + pl.lin0_b() + 4 * lig);
    for (int il = warp * 4 + sub; il < n; il += T / 8) {
      const float xv = __ldg(a.x + rb + __ldg(a.perm + lo + il));
      const float4 o = make_float4(fmaf(xv, wv.x, bv.x), fmaf(xv, wv.y, bv.y), fmaf(xv, wv.z, bv.z), fmaf(xv, wv.w, bv.w));
      const int off = sw_chunk(il, lig);
      st4(XA + off, o);
      st4(XL + off, lo4(o));
    }
  }
  cp_wait_all();
  if (a.nb > 0) lo_own_weights(W1H, W1L, W2H, W2L);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const unsigned h1_base = smem_u32(h1s), h2_base = smem_u32(h2s), z_base = smem_u32(zs);
  const unsigned ss1_base = smem_u32(ss1), ss2_base = smem_u32(ss2);
  const uint32_t xa_u = smem_u32(XA), xl_u = smem_u32(XL), ya_u = smem_u32(YA), yl_u = smem_u32(YL);
  const uint32_t w1h_u = smem_u32(W1H), w1l_u = smem_u32(W1L), w2h_u = smem_u32(W2H), w2l_u = smem_u32(W2L);
  uint32_t phase = 0;

  for (int k = 0; k < a.nb; ++k) {
    const int buf = k & 1;
    const float* vc = vec + buf * VECF;
    if (storer) bulk_wait_read<0>();         // the previous block's h1 / y1 / h2 / score images have left shared memory
    fence_proxy_async();                     // x / low parts / weights written through the generic proxy -> tensor core
    __syncthreads();

    // conv1 projection + scores  (GraphModels.py:464)
    stamp();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      issue_proj_tc<NC, 2 * NC>(tmem, xa_u, xl_u, 0u, w1h_u, w1l_u, 0u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);                   // every thread: W1 / x are free for whoever writes them next
    phase ^= 1u;
    if (TC_M == 64 || (warp & 3) < 2) {
      tc_fence_after();
      const int row = TC_M == 64 ? 16 * (warp & 3) + (lane & 15) : 32 * (warp & 3) + lane, half = warp >> 2;
      drain_row_tc(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + 32u * half, vc + 32 * half, vc + 2 * NC + 32 * half,
                   h1s + row * LDY + 32 * half, ss1 + row * 2 + half, sd1 + row * 2 + half,
                   row < n && (TC_M == 128 || lane < 16));
      tc_fence_before();
    }
    stamp();
    cb.arrive();
    if (k + 1 < a.nb) stage_w1_tc(a.params + pl.block(k + 1), W1H);
    cp_commit();
    after_arrival();
    if (storer) {                            // the block input
      bulk_s2g(a.saved + im.record(k, b, rank), XA, (unsigned)R * 32u * 4u);
      bulk_commit();
    }
    cluster_wait();
    stamp();
    // conv1 aggregation + bias + ReLU -> y1 operand tiles (one per head) + low parts
    agg_fwd<2, true>(rp_s, col_s, h1_base, LDY * 4, ss1_base, sd1, vc + 4 * NC, YA, TF, TRAIN ? m1 : nullptr,
                     TRAIN ? l1 : nullptr, n, rank, true, YL);
    fence_proxy_async();
    __syncthreads();
    stamp();
    // conv2 projection + scores  (:465)
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      issue_proj_tc<2 * NC, NC>(tmem, ya_u, yl_u, (uint32_t)TF * 4u, w2h_u, w2l_u, NC * 32 * 4u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    if (warp < (TC_M == 64 ? 4 : 2)) {
      tc_fence_after();
      const int row = TC_M == 64 ? 16 * warp + (lane & 15) : 32 * warp + lane;
      drain_row_tc(tmem + ((uint32_t)(32 * warp) << 16), vc + 6 * NC, vc + 7 * NC, h2s + row * LDX, ss2 + row, sd2 + row,
                   row < n && (TC_M == 128 || lane < 16));
      tc_fence_before();
    }
    stamp();
    if (TRAIN) fence_proxy_async();          // h1 / scores / (m, l) of conv1 -> async proxy (the bulk copies below)
    cb.arrive();
    if (k + 1 < a.nb) {
      stage_w2_tc(a.params + pl.block(k + 1), W2H);
      stage_vec_tc(a.params + pl.block(k + 1), vec + (buf ^ 1) * VECF);
    }
    cp_commit();
    after_arrival();
    if (storer) {                            // conv1's tensors: all final, none rewritten before the next block
      float* rec = a.saved + im.record(k, b, rank);
      bulk_s2g(rec + im.off_h1, h1s, (unsigned)R * LDY * 4u);
      bulk_s2g(rec + im.off_y, YA, (unsigned)R * 32u * 4u);
      bulk_s2g(rec + im.off_y + R * 32, YA + TF, (unsigned)R * 32u * 4u);
      bulk_s2g(rec + im.off_s1, ss1, 4u * (unsigned)a4(2 * R) * 4u);
      bulk_commit();
    }
    cluster_wait();
    stamp();
    // conv2 aggregation + bias -> z (neighbours read it in the mean)
    agg_fwd<1>(rp_s, col_s, h2_base, LDX * 4, ss2_base, sd2, vc + 8 * NC, zs, LDX, TRAIN ? m2 : nullptr,
               TRAIN ? l2 : nullptr, n, rank, false);
    stamp();
    if (storer) bulk_wait_read<1>();         // the block input has left the x tile (the mean below rewrites it)
    if (TRAIN) fence_proxy_async();          // h2 / scores / (m, l) of conv2 -> async proxy
    cb.arrive();
    after_arrival();
    if (storer) {
      float* rec = a.saved + im.record(k, b, rank);
      bulk_s2g(rec + im.off_h2, h2s, (unsigned)R * LDX * 4u);
      bulk_s2g(rec + im.off_s2, ss2, 4u * (unsigned)a4(R) * 4u);
      bulk_commit();
    }
    cp_wait_all();                           // the next block's W1 / W2 / vectors have landed
    if (k + 1 < a.nb) lo_own_weights(W1H, W1L, W2H, W2L);
    cluster_wait();
    stamp();
    // SimpleConv(mean) + residual + ReLU  (:466-467) -> next block's x operand + low part
    {
      const int slot = lane & 3, sub = lane >> 2;
      for (int il = warp * 8 + sub; il < n; il += T / 4) {
        const int beg = rp_s[il], end = rp_s[il + 1] - 1;
        float4 acc0 = f4zero(), acc1 = f4zero();
#pragma unroll 4
        for (int e = beg; e < end; ++e) {
          const unsigned ar = row_addr(z_base, col_s[e], LDX * 4) + 16u * slot;
          add4(acc0, ldsc4(ar));
          add4(acc1, ldsc4(ar + 64u));
        }
        const int deg = end - beg;
        const float inv = 1.f / (float)(deg > 1 ? deg : 1);
        const int o0 = sw_chunk(il, slot), o1 = sw_chunk(il, slot + 4);
        const float4 r0 = lds4(XA + o0), r1 = lds4(XA + o1);
        const float4 n0 = relu4(make_float4(fmaf(acc0.x, inv, r0.x), fmaf(acc0.y, inv, r0.y), fmaf(acc0.z, inv, r0.z), fmaf(acc0.w, inv, r0.w)));
        const float4 n1 = relu4(make_float4(fmaf(acc1.x, inv, r1.x), fmaf(acc1.y, inv, r1.y), fmaf(acc1.z, inv, r1.z), fmaf(acc1.w, inv, r1.w)));
        st4(XA + o0, n0);
        st4(XA + o1, n1);
        st4(XL + o0, lo4(n0));
        st4(XL + o1, lo4(n1));
      }
    }
  }
  cp_wait_all();
  if (TRAIN) fence_proxy_async();
  __syncthreads();
  // no CTA may leave while a neighbour still reads its shared memory
  cb.arrive();
  if (storer) {                              // the stack's output (decoder input)
    bulk_s2g(a.saved + im.tail(a.nb, b, rank), XA, (unsigned)R * 32u * 4u);
    bulk_commit();
  }
  // decoder Linear(nc, 1)  (:492)
  {
    const int lig = lane & 7, sub = lane >> 3;
    const float4 wv = ldg4(a.params + pl.lin1_w() + 4 * lig);
    const float bias = __ldg(a.params + pl.lin1_b());
    const bool bad = a.poison != nullptr && __ldg(a.poison) != 0;
    for (int i0 = 0; i0 < n; i0 += T / 8) {
      const int il = i0 + warp * 4 + sub;
      float p = il < n ? dot4(lds4(XA + sw_chunk(il, lig)), wv) : 0.f;
      p = group_sum<8>(p, FULL);
      if (il < n && lig == 0) a.out[rb + __ldg(a.perm + lo + il)] = bad ? __int_as_float(0x7fc00000) : p + bias;
    }
  }
  if (storer) bulk_wait_all();
  cluster_wait();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TC_COLS);
}

// =============================================================================== backward
// dx[m][c] = sum_r G[m][r] W[r][c]  (data gradient; W [NRED][NOUT] as stored).  pre(m, c) -> float2 is evaluated for
// every output fragment BEFORE the MMAs; fin(m, c, value, pre value) gets columns c, c+1.
template <int NRED, int NOUT, int LDG, int LDW, typename Pre, typename Fin>
__device__ __forceinline__ void dgrad_mma(const float* Gs, const float* Ws, int n, Pre pre, Fin fin) {
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

// dW[no][ki] += sum_m G[m][no] X[m][ki]: 16 x 8 output tiles, the reduction runs over this CTA's rows (zero padded).
// Split in two so that the atomics can be issued later than the MMAs (behind a barrier arrival).
template <int NO, int KI>
struct WgradAcc {
  static constexpr int TILES = (NO / 16) * (KI / 8), TPW = TILES / (T / 32);
  static_assert(TILES % (T / 32) == 0, "wgrad tiles per warp");
  float acc[TPW][4];
};
template <int NO, int KI, int LDG, int LDXX>
__device__ __forceinline__ void wgrad_mma_compute(const float* Gs, const float* Xs, int n, WgradAcc<NO, KI>& out) {
  constexpr int TPW = WgradAcc<NO, KI>::TPW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  float acl[TPW][4], acm[TPW][4];
  int no0[TPW], ki0[TPW];
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    const int tile = warp * TPW + q;
    no0[q] = (tile / (KI / 8)) * 16;
    ki0[q] = (tile % (KI / 8)) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) out.acc[q][i] = acl[q][i] = acm[q][i] = 0.f;
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
      mma_3xtf32(out.acc[q], acl[q], acm[q], a, al, b, bl);
    }
  }
#pragma unroll
  for (int q = 0; q < TPW; ++q) fold3(out.acc[q], acl[q], acm[q]);
}
template <int NO, int KI>
__device__ __forceinline__ void wgrad_mma_flush(const WgradAcc<NO, KI>& in, int n, float* dW) {
  constexpr int TPW = WgradAcc<NO, KI>::TPW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  if (n <= 0) return;
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    const int tile = warp * TPW + q;
    const int no0 = (tile / (KI / 8)) * 16, ki0 = (tile % (KI / 8)) * 8;
    atomicAdd(reinterpret_cast<float2*>(dW + (size_t)(no0 + g) * KI + ki0 + 2 * t), make_float2(in.acc[q][0], in.acc[q][1]));
    atomicAdd(reinterpret_cast<float2*>(dW + (size_t)(no0 + g + 8) * KI + ki0 + 2 * t), make_float2(in.acc[q][2], in.acc[q][3]));
  }
}
template <int NO, int KI, int LDG, int LDXX>
__device__ __forceinline__ void wgrad_mma(const float* Gs, const float* Xs, int n, float* dW) {
  WgradAcc<NO, KI> acc;
  wgrad_mma_compute<NO, KI, LDG, LDXX>(Gs, Xs, n, acc);
  wgrad_mma_flush<NO, KI>(acc, n, dW);
}

// sum per-lane float4 accumulators over every lane of the CTA that holds the same chunk (lane % LPR) and add the LPR
// chunk sums into dst (atomic).  red: T*4 floats of shared scratch.  All threads call.
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
// sum a per-lane float4 over the four row slots of a warp and park it in this warp's row of `vred`
__device__ __forceinline__ void warp_chunk_park(float4 v, float* dst) {
#pragma unroll
  for (int o = 8; o < 32; o <<= 1) {
    v.x += __shfl_xor_sync(FULL, v.x, o); v.y += __shfl_xor_sync(FULL, v.y, o);
    v.z += __shfl_xor_sync(FULL, v.z, o); v.w += __shfl_xor_sync(FULL, v.w, o);
  }
  if ((threadIdx.x & 31) < 8) st4(dst + 4 * (threadIdx.x & 31), v);
}

// sum a per-lane float4 over the eight row slots of a warp (four lanes per row) and park it in this warp's row of vred
__device__ __forceinline__ void warp_chunk_park4(float4 v, float* dst) {
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    v.x += __shfl_xor_sync(FULL, v.x, o); v.y += __shfl_xor_sync(FULL, v.y, o);
    v.z += __shfl_xor_sync(FULL, v.z, o); v.w += __shfl_xor_sync(FULL, v.w, o);
  }
  if ((threadIdx.x & 31) < 4) st4(dst + 4 * (threadIdx.x & 31), v);
}

// Backward pass 1 of ONE head over own target rows: D_i, ds_dst[i]; rec = {s_dst, m, 1/(l+eps), D}
// (read by the neighbours' pass 2 through DSMEM).  Four lanes own a row — a warp covers eight rows, the CTA's slice
// is one pass — and lane `slot` holds the chunks {slot, slot + 4} of the head and evaluates the edges {slot, slot + 4}
// of every group of eight in-edges.  Two-head layers call it once per head (the chain of a head is half as long as
// the packed two-head chain of the first generation, and there is one row pass instead of two).
// MEAN (conv2, H = 1): the incoming gradient is produced on the fly as the SimpleConv(mean) backward of the running
// gradient g (dz[j] = sum_{j->i} g[i] / max(indeg(i), 1); wt_s holds the weight of every out-edge) and is also written
// to dz_s for pass 2's gathers; otherwise it is read from the own tile g_s.
template <int H, bool MEAN>
__device__ __forceinline__ void bwd_p1_head(const int v, const int* rp_s, const int* col_s, const int* rpt_s,
                                            const int* colt_s, const float* wt_s, unsigned g_base, float* dz_s,
                                            const float* g_s, int ldg_s, unsigned h_base, int ldh, unsigned ss_base,
                                            const float* sd_s, const float* m_s, const float* l_s, float* erec_s,
                                            float* dsd_s, float* vred_bias, int n, int self_owner) {
  static_assert(!MEAN || H == 1, "the mean backward feeds conv2 (one head)");
  constexpr int RPW = 8, PRE = 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 2, slot = lane & 3;
  const unsigned hoff = 128u * v + 16u * slot;
  float4 bacc0 = f4zero(), bacc1 = f4zero();
  for (int i0 = 0; i0 < n; i0 += (T / 32) * RPW) {
    const int il_raw = i0 + warp * RPW + sub;
    const bool ok = il_raw < n;
    const int il = ok ? il_raw : n - 1;
    const int beg = rp_s[il], deg = rp_s[il + 1] - beg;
    const int deg_max = __reduce_max_sync(FULL, deg);
    const int self = (self_owner << 16) | il;
    const float sd = sd_s[il * H + v], mi = m_s[il * H + v], il_ = 1.f / (l_s[il * H + v] + kSoftmaxEps);
    float S1 = 0.f, S2 = 0.f, S3 = 0.f;
    float la0 = 0.f, la1 = 0.f, lk0 = 0.f, lk1 = 0.f, ld0 = 0.f, ld1 = 0.f;      // the last chunk's per-edge values
    float4 gv0, gv1;
    if (MEAN) {
      const int tb = rpt_s[il], te = rpt_s[il + 1] - 1;     // out-edges minus the self-loop
      gv0 = gv1 = f4zero();
#pragma unroll 4
      for (int e = tb; e < te; ++e) {
        const unsigned ar = row_addr(g_base, colt_s[e], LDX * 4) + 16u * slot;
        const float w = wt_s[e];
        fma4(gv0, w, ldsc4(ar));
        fma4(gv1, w, ldsc4(ar + 64u));
      }
      if (ok) {
        st4(dz_s + il * LDX + 4 * slot, gv0);
        st4(dz_s + il * LDX + 16 + 4 * slot, gv1);
      }
    } else {
      gv0 = lds4(g_s + il * ldg_s + 32 * v + 4 * slot);
      gv1 = lds4(g_s + il * ldg_s + 32 * v + 16 + 4 * slot);
    }
    if (ok) { add4(bacc0, gv0); add4(bacc1, gv1); }
    for (int e0 = 0; e0 < deg_max; e0 += 8) {
      const bool valid0 = e0 + slot < deg, valid1 = e0 + slot + 4 < deg;
      const int j0 = valid0 ? col_s[beg + e0 + slot] : self;
      const int j1 = valid1 ? col_s[beg + e0 + slot + 4] : self;
      const int cnt = min(8, deg - e0), cnt_max = min(8, deg_max - e0);
      float4 x[PRE][2];
#pragma unroll
      for (int u = 0; u < PRE; ++u) {
        const int ju = __shfl_sync(FULL, j0, u, 4);
        const unsigned au = row_addr(h_base, ju, ldh) + hoff;
        x[u][0] = u < cnt ? ldsc4(au) : f4zero();
        x[u][1] = u < cnt ? ldsc4(au + 64u) : f4zero();
      }
      const float z0 = ldsc1(row_addr(ss_base, j0, 4 * H) + 4u * v) + sd;
      const float z1 = cnt_max > 4 ? ldsc1(row_addr(ss_base, j1, 4 * H) + 4u * v) + sd : 0.f;
      const float alpha0 = valid0 ? __expf(lrelu(z0) - mi) * il_ : 0.f;
      const float alpha1 = valid1 ? __expf(lrelu(z1) - mi) * il_ : 0.f;
      float da0 = 0.f, da1 = 0.f;
#pragma unroll
      for (int u = 0; u < PRE; ++u) {
        const float d = group_sum<4>(dot4(gv0, x[u][0]) + dot4(gv1, x[u][1]), FULL);
        da0 = slot == u ? d : da0;
      }
      for (int t = PRE; t < cnt_max; ++t) {
        const int jt = __shfl_sync(FULL, j1, t & 3, 4);
        const unsigned at = row_addr(h_base, jt, ldh) + hoff;
        const bool on = t < cnt;
        const float4 xa = on ? ldsc4(at) : f4zero(), xb = on ? ldsc4(at + 64u) : f4zero();
        const float d = group_sum<4>(dot4(gv0, xa) + dot4(gv1, xb), FULL);
        da1 = slot == (t & 3) ? d : da1;
      }
      const float sl0 = lrelu_slope(z0), sl1 = lrelu_slope(z1);
      S1 = fmaf(alpha0, da0, fmaf(alpha1, da1, S1));
      S2 = fmaf(alpha0 * sl0, da0, fmaf(alpha1 * sl1, da1, S2));
      S3 = fmaf(alpha0, sl0, fmaf(alpha1, sl1, S3));
      la0 = alpha0; la1 = alpha1; lk0 = alpha0 * sl0; lk1 = alpha1 * sl1; ld0 = da0; ld1 = da1;
    }
    const float D = group_sum<4>(S1, FULL), T2 = group_sum<4>(S2, FULL), T3 = group_sum<4>(S3, FULL);
    if (slot == 0 && ok) dsd_s[il * H + v] = T2 - D * T3;
    // edge records {alpha_e, dz_e = alpha_e slope_e (dalpha_e - D_i)} at the in-edge's position: pass 2 of the edge's SOURCE
    // reads them through distributed shared memory and needs neither the softmax record of the target nor <g_i, h_j> again
    if (deg_max <= 8) {
      if (ok && slot < deg) *reinterpret_cast<float2*>(erec_s + (beg + slot) * 2 * H + 2 * v) = make_float2(la0, lk0 * (ld0 - D));
      if (ok && slot + 4 < deg) *reinterpret_cast<float2*>(erec_s + (beg + slot + 4) * 2 * H + 2 * v) = make_float2(la1, lk1 * (ld1 - D));
    } else {                                   // rows with more than eight in-edges: second sweep (values recomputed)
      for (int e0 = 0; e0 < deg_max; e0 += 8) {
        const bool valid0 = e0 + slot < deg, valid1 = e0 + slot + 4 < deg;
        const int j0 = valid0 ? col_s[beg + e0 + slot] : self;
        const int j1 = valid1 ? col_s[beg + e0 + slot + 4] : self;
        const int cnt = min(8, deg - e0), cnt_max = min(8, deg_max - e0);
        const float z0 = ldsc1(row_addr(ss_base, j0, 4 * H) + 4u * v) + sd;
        const float z1 = ldsc1(row_addr(ss_base, j1, 4 * H) + 4u * v) + sd;
        const float alpha0 = valid0 ? __expf(lrelu(z0) - mi) * il_ : 0.f;
        const float alpha1 = valid1 ? __expf(lrelu(z1) - mi) * il_ : 0.f;
        float da0 = 0.f, da1 = 0.f;
        for (int t = 0; t < cnt_max; ++t) {
          const int jt = __shfl_sync(FULL, t < 4 ? j0 : j1, t & 3, 4);
          const unsigned at = row_addr(h_base, jt, ldh) + hoff;
          const bool on = t < cnt;
          const float4 xa = on ? ldsc4(at) : f4zero(), xb = on ? ldsc4(at + 64u) : f4zero();
          const float d = group_sum<4>(dot4(gv0, xa) + dot4(gv1, xb), FULL);
          if (t < 4) da0 = slot == t ? d : da0;
          else da1 = slot == (t & 3) ? d : da1;
        }
        if (ok && valid0)
          *reinterpret_cast<float2*>(erec_s + (beg + e0 + slot) * 2 * H + 2 * v) = make_float2(alpha0, alpha0 * lrelu_slope(z0) * (da0 - D));
        if (ok && valid1)
          *reinterpret_cast<float2*>(erec_s + (beg + e0 + slot + 4) * 2 * H + 2 * v) = make_float2(alpha1, alpha1 * lrelu_slope(z1) * (da1 - D));
      }
    }
  }
  warp_chunk_park4(bacc0, vred_bias + 32 * v);
  warp_chunk_park4(bacc1, vred_bias + 32 * v + 16);
}

// Packed form of pass 1 (eight lanes per row, the heads of a row packed per lane): used for the two-head layer.
// Backward pass 1 over own target rows: D_i, ds_dst[i]; rec = {s_dst, m, 1/(l+eps), D} (read by the
// neighbours' pass 2 through DSMEM).  MEAN (conv2, H = 1): the incoming gradient is produced on the fly as the
// SimpleConv(mean) backward of the running gradient g (dz[j] = sum_{j->i} g[i] / max(indeg(i), 1); wt_s holds the
// weight of every out-edge) and is also written to dz_s for pass 2's gathers; otherwise it is read from the tile g_s.
template <int H, bool MEAN>
__device__ __forceinline__ void bwd_p1_packed(const int* rp_s, const int* col_s, const int* rpt_s, const int* colt_s,
                                       const float* wt_s, unsigned g_base, float* dz_s, const float* g_s, int ldg_s,
                                       unsigned h_base, int ldh, unsigned ss_base, const float* sd_s, const float* m_s,
                                       const float* l_s, float* erec_s, float* dsd_s, float* vred_bias, int n,
                                       int self_owner) {
  static_assert(!MEAN || H == 1, "the mean backward feeds conv2 (one head)");
  constexpr int RPW = 4, PRE = 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >> 3, slot = lane & 7;
  float4 bacc[H];
#pragma unroll
  for (int v = 0; v < H; ++v) bacc[v] = f4zero();
  for (int i0 = 0; i0 < n; i0 += (T / 32) * RPW) {
    const int il_raw = i0 + warp * RPW + sub;
    const bool ok = il_raw < n;
    const int il = ok ? il_raw : n - 1;
    const int beg = rp_s[il], deg = rp_s[il + 1] - beg;
    const int deg_max = __reduce_max_sync(FULL, deg);
    const int self = (self_owner << 16) | il;
    // neighbour rows and scores of the first chunk are requested before the mean backward's own gathers
    const bool valid0 = slot < deg;
    const int j0 = valid0 ? col_s[beg + slot] : self;
    const int cnt0 = min(8, deg);
    float4 x0[PRE][H];
    float ssv0[H], sd[H], mi[H], il_[H], S1[H], S2[H], S3[H];
#pragma unroll
    for (int u = 0; u < PRE; ++u) {
      const int ju = __shfl_sync(FULL, j0, u, 8);
      const unsigned au = row_addr(h_base, ju, ldh) + 16u * slot;
#pragma unroll
      for (int v = 0; v < H; ++v) x0[u][v] = u < cnt0 ? ldsc4(au + 128u * v) : f4zero();
    }
    {
      const unsigned as_ = row_addr(ss_base, j0, 4 * H);
      if (H == 2) {
        const float2 t2 = ldsc2(as_);
        ssv0[0] = t2.x; ssv0[H - 1] = t2.y;
      } else ssv0[0] = ldsc1(as_);
    }
#pragma unroll
    for (int v = 0; v < H; ++v) {
      sd[v] = sd_s[il * H + v];
      mi[v] = m_s[il * H + v];
      il_[v] = 1.f / (l_s[il * H + v] + kSoftmaxEps);
      S1[v] = S2[v] = S3[v] = 0.f;
    }
    float la[H], lk[H], ld[H];                 // the last chunk's per-edge values
#pragma unroll
    for (int v = 0; v < H; ++v) la[v] = lk[v] = ld[v] = 0.f;
    float4 gv[H];
    if (MEAN) {
      const int tb = rpt_s[il], te = rpt_s[il + 1] - 1;     // out-edges minus the self-loop
      gv[0] = f4zero();
#pragma unroll 4
      for (int e = tb; e < te; ++e) fma4(gv[0], wt_s[e], ldsc4(row_addr(g_base, colt_s[e], LDX * 4) + 16u * slot));
      if (ok) st4(dz_s + il * LDX + 4 * slot, gv[0]);
    } else {
#pragma unroll
      for (int v = 0; v < H; ++v) gv[v] = lds4(g_s + il * ldg_s + 32 * v + 4 * slot);
    }
#pragma unroll
    for (int v = 0; v < H; ++v)
      if (ok) add4(bacc[v], gv[v]);
    for (int e0 = 0; e0 < deg_max; e0 += 8) {
      const bool valid = e0 + slot < deg;
      const int j = e0 == 0 ? j0 : (valid ? col_s[beg + e0 + slot] : self);
      const int cnt = min(8, deg - e0), cnt_max = min(8, deg_max - e0);
      float4 x[PRE][H];
      float alpha[H], sl[H], da[H], ssv[H];
      if (e0 == 0) {
#pragma unroll
        for (int u = 0; u < PRE; ++u)
#pragma unroll
          for (int v = 0; v < H; ++v) x[u][v] = x0[u][v];
#pragma unroll
        for (int v = 0; v < H; ++v) ssv[v] = ssv0[v];
      } else {
#pragma unroll
        for (int u = 0; u < PRE; ++u) {
          const int ju = __shfl_sync(FULL, j, u, 8);
          const unsigned au = row_addr(h_base, ju, ldh) + 16u * slot;
#pragma unroll
          for (int v = 0; v < H; ++v) x[u][v] = u < cnt ? ldsc4(au + 128u * v) : f4zero();
        }
        const unsigned as_ = row_addr(ss_base, j, 4 * H);
        if (H == 2) {
          const float2 t2 = ldsc2(as_);
          ssv[0] = t2.x; ssv[H - 1] = t2.y;
        } else ssv[0] = ldsc1(as_);
      }
#pragma unroll
      for (int v = 0; v < H; ++v) {
        const float z = ssv[v] + sd[v];
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
        const unsigned a0 = row_addr(h_base, j0_, ldh) + 16u * slot, a1 = row_addr(h_base, j1_, ldh) + 16u * slot;
#pragma unroll
        for (int v = 0; v < H; ++v) {
          const float4 xa = t < cnt ? ldsc4(a0 + 128u * v) : f4zero();
          const float4 xb = t + 1 < cnt ? ldsc4(a1 + 128u * v) : f4zero();
          const float d0 = group_sum<8>(dot4(gv[v], xa), FULL), d1 = group_sum<8>(dot4(gv[v], xb), FULL);
          da[v] = slot == t ? d0 : (slot == t + 1 ? d1 : da[v]);
        }
      }
#pragma unroll
      for (int v = 0; v < H; ++v) {
        S1[v] = fmaf(alpha[v], da[v], S1[v]);
        S2[v] = fmaf(alpha[v] * sl[v], da[v], S2[v]);
        S3[v] = fmaf(alpha[v], sl[v], S3[v]);
        la[v] = alpha[v]; lk[v] = alpha[v] * sl[v]; ld[v] = da[v];
      }
    }
    float Dv[H];
#pragma unroll
    for (int v = 0; v < H; ++v) {
      const float D = group_sum<8>(S1[v], FULL), T2 = group_sum<8>(S2[v], FULL), T3 = group_sum<8>(S3[v], FULL);
      if (slot == 0 && ok) dsd_s[il * H + v] = T2 - D * T3;
      Dv[v] = D;
    }
    // edge records {alpha_e, dz_e} of every head at the in-edge's position (see bwd_p1_head)
    auto put = [&](int pos, const float (&al)[H], const float (&dz)[H]) {
      if (H == 2) st4(erec_s + pos * 4, make_float4(al[0], dz[0], al[H - 1], dz[H - 1]));
      else *reinterpret_cast<float2*>(erec_s + pos * 2) = make_float2(al[0], dz[0]);
    };
    if (deg_max <= 8) {
      if (ok && slot < deg) {
        float dz[H];
#pragma unroll
        for (int v = 0; v < H; ++v) dz[v] = lk[v] * (ld[v] - Dv[v]);
        put(beg + slot, la, dz);
      }
    } else {                                   // rows with more than eight in-edges: second sweep (values recomputed)
      for (int e0 = 0; e0 < deg_max; e0 += 8) {
        const bool valid = e0 + slot < deg;
        const int j = valid ? col_s[beg + e0 + slot] : self;
        const int cnt = min(8, deg - e0), cnt_max = min(8, deg_max - e0);
        float ssv[H], alpha[H], dz[H], da[H];
        load_scores<H>(ss_base, j, ssv);
#pragma unroll
        for (int v = 0; v < H; ++v) da[v] = 0.f;
        for (int t = 0; t < cnt_max; ++t) {
          const int jt = __shfl_sync(FULL, j, t, 8);
          const unsigned at = row_addr(h_base, jt, ldh) + 16u * slot;
#pragma unroll
          for (int v = 0; v < H; ++v) {
            const float4 xa = t < cnt ? ldsc4(at + 128u * v) : f4zero();
            const float d = group_sum<8>(dot4(gv[v], xa), FULL);
            da[v] = slot == t ? d : da[v];
          }
        }
#pragma unroll
        for (int v = 0; v < H; ++v) {
          const float z = ssv[v] + sd[v];
          alpha[v] = valid ? __expf(lrelu(z) - mi[v]) * il_[v] : 0.f;
          dz[v] = alpha[v] * lrelu_slope(z) * (da[v] - Dv[v]);
        }
        if (ok && valid) put(beg + e0 + slot, alpha, dz);
      }
    }
  }
#pragma unroll
  for (int v = 0; v < H; ++v) warp_chunk_park(bacc[v], vred_bias + 32 * v);
}

template <int H, bool MEAN>
__device__ __forceinline__ void bwd_p1(const int* rp_s, const int* col_s, const int* rpt_s, const int* colt_s,
                                       const float* wt_s, unsigned g_base, float* dz_s, const float* g_s, int ldg_s,
                                       unsigned h_base, int ldh, unsigned ss_base, const float* sd_s, const float* m_s,
                                       const float* l_s, float* erec_s, float* dsd_s, float* vred_bias, int n,
                                       int self_owner) {
  if (H == 1)
    bwd_p1_head<H, MEAN>(0, rp_s, col_s, rpt_s, colt_s, wt_s, g_base, dz_s, g_s, ldg_s, h_base, ldh, ss_base, sd_s, m_s, l_s,
                         erec_s, dsd_s, vred_bias, n, self_owner);
  else
    bwd_p1_packed<H, MEAN>(rp_s, col_s, rpt_s, colt_s, wt_s, g_base, dz_s, g_s, ldg_s, h_base, ldh, ss_base, sd_s, m_s, l_s,
                           erec_s, dsd_s, vred_bias, n, self_owner);
}
// Backward pass 2 over own source rows: dh[j] (-> own shared tile), ds_src, datt_src / datt_dst partials.  Eight lanes
// per row with the heads packed per lane (measured faster than the four-lane / one-head form for this pass: its chain
// is dominated by the dependent dot -> softmax-gradient -> accumulate sequence, not by the number of row passes).  The
// gradient rows g[i] and the records of the edges' targets come through DSMEM; h, s_src, ds_dst of the row are own.
template <int H>
__device__ __forceinline__ void bwd_p2(const int* rpt_s, const int* colt_s, const int* emap_s, unsigned g_base, int ldg,
                                       unsigned erec_base, const float* dsd_s, const float* h_s, int ldh_own,
                                       const float* att_s, const float* att_d, float* dh_s, int ld_dh, float* vred_as,
                                       float* vred_ad, int n, int self_owner) {
  constexpr int RPW = 4, PRE = H == 1 ? 4 : 2;     // two heads per lane: fewer gathers in flight (registers)
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
    const int il = ok ? il_raw : n - 1;
    const int beg = rpt_s[il], deg = rpt_s[il + 1] - beg;
    const int deg_max = __reduce_max_sync(FULL, deg);
    const int self = (self_owner << 16) | il;
    float4 hv[H], dacc[H];
    float dsrc[H];
#pragma unroll
    for (int v = 0; v < H; ++v) {
      hv[v] = lds4(h_s + il * ldh_own + 32 * v + 4 * slot);
      dacc[v] = f4zero();
      dsrc[v] = 0.f;
    }
    for (int e0 = 0; e0 < deg_max; e0 += 8) {
      const bool valid = e0 + slot < deg;
      const int i = valid ? colt_s[beg + e0 + slot] : self;
      const int em = valid ? emap_s[beg + e0 + slot] : 0;
      const int cnt = min(8, deg - e0), cnt_max = min(8, deg_max - e0);
      float4 gx[PRE][H];
#pragma unroll
      for (int u = 0; u < PRE; ++u) {
        const int iu = __shfl_sync(FULL, i, u, 8);
        const unsigned au = row_addr(g_base, iu, ldg) + 16u * slot;
#pragma unroll
        for (int v = 0; v < H; ++v) gx[u][v] = u < cnt ? ldsc4(au + 128u * v) : f4zero();
      }
      // {alpha_e, dz_e} of this lane's out-edge, written by pass 1 of the edge's TARGET at the edge's in-position
      float alpha[H];
      {
        const unsigned ar = row_addr(erec_base, em, 8 * H);
        if (H == 2) {
          const float4 t4 = valid ? ldsc4(ar) : f4zero();
          alpha[0] = t4.x; alpha[H - 1] = t4.z;
          dsrc[0] += t4.y; dsrc[H - 1] += t4.w;
        } else {
          const float2 t2 = valid ? ldsc2(ar) : make_float2(0.f, 0.f);
          alpha[0] = t2.x;
          dsrc[0] += t2.y;
        }
      }
#pragma unroll
      for (int u = 0; u < PRE; ++u)
#pragma unroll
        for (int v = 0; v < H; ++v) fma4(dacc[v], __shfl_sync(FULL, alpha[v], u, 8), gx[u][v]);
      for (int t = PRE; t < cnt_max; t += 2) {
        const int i0_ = __shfl_sync(FULL, i, t, 8), i1_ = __shfl_sync(FULL, i, t + 1, 8);
        const unsigned a0 = row_addr(g_base, i0_, ldg) + 16u * slot, a1 = row_addr(g_base, i1_, ldg) + 16u * slot;
#pragma unroll
        for (int v = 0; v < H; ++v) {
          const float4 g0 = t < cnt ? ldsc4(a0 + 128u * v) : f4zero();
          const float4 g1 = t + 1 < cnt ? ldsc4(a1 + 128u * v) : f4zero();
          const float a0v = __shfl_sync(FULL, alpha[v], t, 8), a1v = __shfl_sync(FULL, alpha[v], t + 1, 8);
          fma4(dacc[v], a0v, g0);
          fma4(dacc[v], t + 1 < 8 ? a1v : 0.f, g1);
        }
      }
    }
#pragma unroll
    for (int v = 0; v < H; ++v) {
      const float ds = group_sum<8>(dsrc[v], FULL);
      const float dd = dsd_s[il * H + v];
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

// Shared memory (floats): g[R][LDX] h2s[R][LDX] dzd[R][2 LDX] (dz | dh2, later dh1) h1s[R][LDY] ys[R][LDY] xs[R][LDX]
//   ss2[R] sc2[3R] (sd2 m2 l2) erec2[2 ecap] dsd2[R] ss1[2R] sc1[6R] (sd1 m1 l1) erec1[4 ecap] dsd1[2R] emap[ecap]
//   W1s[W1F] W2s[W2F] vec[VECF] vred[8][VECF] wt[ecap] | ints: rp[R+1] col[ecap] rpt[R+1] colt[ecap]
struct BwdSmem {
  int g, h2s, dzd, h1s, ys, xs, ss2, sc2, erec2, dsd2, ss1, sc1, erec1, dsd1, emap, W1s, W2s, vec, vred, wt, rp, col, rpt, colt, total;
  __host__ __device__ BwdSmem(int R, int ecap) {
    int o = 0;
    auto take = [&](int nfl) { const int at = o; o += (int)a4(nfl); return at; };
    g = take(R * LDX); h2s = take(R * LDX); dzd = take(R * 2 * LDX > T * 4 ? R * 2 * LDX : T * 4); h1s = take(R * LDY); ys = take(R * LDY); xs = take(R * LDX);
    ss2 = take(R); sc2 = take(3 * (int)a4(R)); erec2 = take(2 * ecap); dsd2 = take(R);
    ss1 = take(2 * R); sc1 = take(3 * (int)a4(2 * R)); erec1 = take(4 * ecap); dsd1 = take(2 * R); emap = take(ecap);
    W1s = take(W1F); W2s = take(W2F); vec = take(VECF); vred = take((T / 32) * VECF); wt = take(ecap);
    rp = take(R + 1); col = take(ecap); rpt = take(R + 1); colt = take(ecap);
    total = o;
  }
};

__global__ void __launch_bounds__(T, 2)
bwd_kernel(const Args a) {
  extern __shared__ __align__(16) float smem[];
  const int R = a.R, N = a.N, nb = a.nb;
  const BwdSmem L(R, a.ecap);
  float* gs = smem + L.g;                  // running gradient w.r.t. the block output (own rows; neighbours read it)
  float* h2s = smem + L.h2s;
  float* dz = smem + L.dzd;                // [R][LDX] mean-backward output (neighbours read it in conv2's pass 2)
  float* d2s = dz + R * LDX;               // [R][LDX] dh2 (own)
  float* dh1 = dz;                         // [R][LDY] dh1 (own) over dz | dh2 once both are dead
  float* h1s = smem + L.h1s;
  float* ys = smem + L.ys;                 // y1 -> dy1 (own rows; neighbours read dy1 in conv1's pass 2)
  float* xs = smem + L.xs;                 // x0 of the block
  float* ss2 = smem + L.ss2;
  float* sd2 = smem + L.sc2;
  float* m2 = sd2 + a4(R);
  float* l2 = m2 + a4(R);
  float* erec2 = smem + L.erec2;           // per in-edge {alpha, dz} of conv2 (neighbours' pass 2 reads them)
  float* dsd2 = smem + L.dsd2;
  float* ss1 = smem + L.ss1;
  float* sd1 = smem + L.sc1;
  float* m1 = sd1 + a4(2 * R);
  float* l1 = m1 + a4(2 * R);
  float* erec1 = smem + L.erec1;           // per in-edge {alpha, dz} x 2 heads of conv1
  int* emap_s = reinterpret_cast<int*>(smem + L.emap);      // out-edge slot -> (owner CTA << 16 | in-edge position there)
  float* dsd1 = smem + L.dsd1;
  float* W1 = smem + L.W1s;
  float* W2 = smem + L.W2s;
  float* vc = smem + L.vec;
  float* vred = smem + L.vred;
  float* wt = smem + L.wt;
  float* red = dz;                         // T*4 floats of reduction scratch for the decoder / encoder gradients
  int* rp_s = reinterpret_cast<int*>(smem + L.rp);
  int* col_s = reinterpret_cast<int*>(smem + L.col);
  int* rpt_s = reinterpret_cast<int*>(smem + L.rpt);
  int* colt_s = reinterpret_cast<int*>(smem + L.colt);
  const int rank = (int)cluster_ctarank();
  const ClusterBarrier cb = {a.barrier};
  const long long b = cluster_id_x();
  const int lo = rank * R, n = max(0, min(R, N - lo));
  const long long M = a.M, rb = b * N, ro = rb + lo;
  const ParamLayout pl(nb, NC);
  const SavedLayout sl(M, NC);
  const ImageLayout im(R, M / N, a.cs);    // a.images: the tensor-core forward's records (loaders undo swizzle / padding)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* sv = a.saved;

  if (n > 0) {
    stage_csr_slice(a.rowptr, a.col, lo, n, R, rp_s, col_s, a.ecap);
    // out-edge slice + the SimpleConv(mean) weight of every out-edge: 1 / max(in-degree of its target, 1)
    const int e_lo = __ldg(a.rowptr_t + lo), cnt = min(__ldg(a.rowptr_t + lo + n) - e_lo, a.ecap);
    for (int c = threadIdx.x; c <= n; c += T) rpt_s[c] = __ldg(a.rowptr_t + lo + c) - e_lo;
    for (int c = threadIdx.x; c < cnt; c += T) {
      const int t = __ldg(a.col_t + e_lo + c);
      const int owner = t / R;
      colt_s[c] = (owner << 16) | (t - owner * R);
      const int dg = __ldg(a.rowptr + t + 1) - __ldg(a.rowptr + t) - 1;
      wt[c] = 1.f / (float)(dg > 1 ? dg : 1);
    }
    // where pass 1 of an out-edge's TARGET leaves the edge's record: the position of this edge inside the target's in-edge
    // list, relative to the in-edge slice of the CTA that owns the target.  Both CSRs keep parallel edges in edge-list
    // order, so the m-th occurrence of the target in this row's out-list is the m-th occurrence of this row in its in-list.
    // (one thread per out-edge; the edge's row by bisection of the global row pointers of this slice)
    for (int e = threadIdx.x; e < cnt; e += T) {
      int lo_r = 0, hi_r = n;                  // largest row with rowptr_t[lo + row] - e_lo <= e
      while (hi_r - lo_r > 1) {
        const int mid = (lo_r + hi_r) >> 1;
        if (__ldg(a.rowptr_t + lo + mid) - e_lo <= e) lo_r = mid; else hi_r = mid;
      }
      const int jg = lo + lo_r, eb = __ldg(a.rowptr_t + jg) - e_lo;
      const int t = __ldg(a.col_t + e_lo + e);
      int m = 0;
      for (int q = eb; q < e; ++q) m += __ldg(a.col_t + e_lo + q) == t ? 1 : 0;
      const int tb = __ldg(a.rowptr + t), te = __ldg(a.rowptr + t + 1), owner = t / R;
      int pos = tb;
      for (int q = tb; q < te; ++q)
        if (__ldg(a.col + q) == jg && m-- == 0) { pos = q; break; }
      emap_s[e] = (owner << 16) | (pos - __ldg(a.rowptr + owner * R));
    }
  }
  pdl_wait();
  Stamper stamp(a);
  stamp();
  const int k_first = nb > 0 ? a.k_hi : -1;
  const bool any = nb > 0 && k_first >= a.k_lo && k_first >= 0;

  // cp.async groups in flight (oldest first) while block k runs:
  //   [A: h2 / s2 scalars of k-1, after phase 2] [B: h1 / s1 scalars of k-1, after phase 5] [C: parameters of k-1, after
  //   phase 6] | cluster barrier | [D: y1 / x0 of k-1]
  auto load_conv2_side = [&](int k) {
    if (a.images) {                          // padded h2 rows and the (s_src, s_dst, m, l) block: straight copies
      const float* rec = sv + im.record(k, b, rank);
      for (int c = threadIdx.x; c < n * (LDX / 4); c += T) cp16(h2s + 4 * c, rec + im.off_h2 + 4 * c);
      for (int c = threadIdx.x; c < (int)a4(R); c += T) cp16(ss2 + 4 * c, rec + im.off_s2 + 4 * c);
      return;
    }
    stage_rows<NC, LDX>(sv + sl.h2(k) + ro * NC, h2s, n);
    // scalar arrays start at ro (not 16-byte aligned in general): 4-byte copies
    for (int c = threadIdx.x; c < n; c += T) {
      cp4(ss2 + c, sv + sl.ss2(k) + ro + c);
      cp4(sd2 + c, sv + sl.sd2(k) + ro + c);
      cp4(m2 + c, sv + sl.m2(k) + ro + c);
      cp4(l2 + c, sv + sl.l2(k) + ro + c);
    }
  };
  auto load_conv1_side = [&](int k) {
    if (a.images) {
      const float* rec = sv + im.record(k, b, rank);
      for (int c = threadIdx.x; c < n * (LDY / 4); c += T) cp16(h1s + 4 * c, rec + im.off_h1 + 4 * c);
      for (int c = threadIdx.x; c < (int)a4(2 * R); c += T) cp16(ss1 + 4 * c, rec + im.off_s1 + 4 * c);
      return;
    }
    stage_rows<2 * NC, LDY>(sv + sl.h1(k) + ro * 2 * NC, h1s, n);
    for (int c = threadIdx.x; c < 2 * n; c += T) {
      cp4(ss1 + c, sv + sl.ss1(k) + ro * 2 + c);
      cp4(sd1 + c, sv + sl.sd1(k) + ro * 2 + c);
      cp4(m1 + c, sv + sl.m1(k) + ro * 2 + c);
      cp4(l1 + c, sv + sl.l1(k) + ro * 2 + c);
    }
  };
  auto load_row_local = [&](int k) {
    if (a.images) {                          // SWIZZLE_128B operand rows -> padded row-major tiles
      const float* rec = sv + im.record(k, b, rank);
      for (int c = threadIdx.x; c < n * 16; c += T) {
        const int row = c >> 4, v = (c >> 3) & 1, q = c & 7;
        cp16(ys + row * LDY + 32 * v + 4 * q, rec + im.off_y + v * (R * 32) + sw_chunk(row, q));
      }
      for (int c = threadIdx.x; c < n * 8; c += T) cp16(xs + (c >> 3) * LDX + 4 * (c & 7), rec + sw_chunk(c >> 3, c & 7));
      return;
    }
    stage_rows<2 * NC, LDY>(sv + sl.y1(k) + ro * 2 * NC, ys, n);
    stage_rows<NC, LDX>(sv + (k > 0 ? sl.xout(k - 1) : sl.x_enc()) + ro * NC, xs, n);
  };
  if (any) {
    stage_block_params(a.params + pl.block(k_first), W1, W2, vc);
    load_conv2_side(k_first);
    load_conv1_side(k_first);
  }
  cp_commit();

  if (a.head) {
    // decoder backward: g[i][c] = d_out[i] w[c] (masked by the last ReLU), dw = sum d_out x, db = sum d_out
    const int lig = lane & 7, sub = lane >> 3;
    const float* xl = sv + (nb > 0 ? sl.xout(nb - 1) : sl.x_enc());
    const float* xi = sv + im.tail(nb, b, rank);
    const float4 wv = ldg4(a.params + pl.lin1_w() + 4 * lig);
    float4 aw = f4zero();
    float ab = 0.f;
    for (int il = warp * 4 + sub; il < n; il += T / 8) {
      const float gv = __ldg(a.d_out + rb + __ldg(a.perm + lo + il));
      const float4 xv = a.images ? ldg4(xi + sw_chunk(il, lig)) : ldg4(xl + (ro + il) * NC + 4 * lig);
      float4 d = make_float4(gv * wv.x, gv * wv.y, gv * wv.z, gv * wv.w);
      if (nb > 0) d = mask4(d, xv);
      st4(gs + il * LDX + 4 * lig, d);
      fma4(aw, gv, xv);
      if (lig == 0) ab += gv;
    }
    cta_chunk_sum_atomic<8>(aw, red, a.grads + pl.lin1_w());
    ab = group_sum<32>(ab, FULL);
    if (lane == 0 && ab != 0.f) atomicAdd(a.grads + pl.lin1_b(), ab);
  } else {
    // a later range of a split backward: the running gradient comes back from the scratch buffer (locality order)
    for (int c = threadIdx.x; c < n * (NC / 4); c += T)
      st4(gs + (c / (NC / 4)) * LDX + 4 * (c % (NC / 4)), ldg4_stream(a.scratch + (ro * NC) + 4 * c));
  }
  cp_wait_all();
  cb.arrive();
  cluster_wait();                            // g, h2 / h1 and their source scores are visible cluster-wide

  const unsigned g_base = smem_u32(gs), h2_base = smem_u32(h2s), h1_base = smem_u32(h1s), dz_base = smem_u32(dz);
  const unsigned y_base = smem_u32(ys), ss2_base = smem_u32(ss2), ss1_base = smem_u32(ss1);
  const unsigned erec2_base = smem_u32(erec2), erec1_base = smem_u32(erec1);

  for (int k = k_first; k >= a.k_lo && k >= 0; --k) {
    const bool more = k - 1 >= a.k_lo && k - 1 >= 0;
    load_row_local(k);                       // group D of this block: y1 / x0 (own rows only), needed from phase 3 on
    cp_commit();
    stamp();
    // (1) SimpleConv(mean) backward fused with conv2 pass 1 (incoming gradient dz stays in registers)
    bwd_p1<1, true>(rp_s, col_s, rpt_s, colt_s, wt, g_base, dz, nullptr, 0, h2_base, LDX * 4, ss2_base, sd2, m2, l2, erec2,
                    dsd2, vred + warp * VECF + 8 * NC, n, rank);
    cp_wait_but_one();                       // this block's parameters have landed (group D may still fly)
    stamp();
    cb.arrive();
    cluster_wait();                          // dz / edge records of conv2 cluster-wide, parameters CTA-wide
    stamp();
    // (2) conv2 pass 2 -> dh2 (own)
    bwd_p2<1>(rpt_s, colt_s, emap_s, dz_base, LDX * 4, erec2_base, dsd2, h2s, LDX, vc + 6 * NC, vc + 7 * NC, d2s, LDX,
              vred + warp * VECF + 6 * NC, vred + warp * VECF + 7 * NC, n, rank);
    cp_wait_all();                           // y1 / x0 rows of this block
    __syncthreads();                         // dh2, y1 and x0 are in shared memory; h2 / s2 buffers are free
    if (more) load_conv2_side(k - 1);        // group A
    cp_commit();
    stamp();
    // (3) conv2 projection backward: dW2 = dh2^T y1 ; dy1 = (dh2 W2) masked by y1 > 0 (in place over y1)
    wgrad_mma<NC, 2 * NC, LDX, LDY>(d2s, ys, n, a.grads + pl.c2_W(k));
    __syncthreads();
    dgrad_mma<NC, 2 * NC, LDX, LDY>(d2s, W2, n, [](int, int) { return make_float2(0.f, 0.f); }, [&](int m, int c, float2 v, float2) {
      float2* p = reinterpret_cast<float2*>(ys + m * LDY + c);
      const float2 y = *p;
      *p = make_float2(y.x > 0.f ? v.x : 0.f, y.y > 0.f ? v.y : 0.f);
    });
    __syncthreads();
    stamp();
    // (4) conv1 pass 1 (incoming gradient = dy1 from the own tile)
    bwd_p1<2, false>(rp_s, col_s, rpt_s, colt_s, wt, 0u, nullptr, ys, LDY, h1_base, LDY * 4, ss1_base, sd1, m1, l1, erec1,
                     dsd1, vred + warp * VECF + 4 * NC, n, rank);
    stamp();
    cb.arrive();
    cluster_wait();                          // dy1 / edge records of conv1 cluster-wide
    stamp();
    // (5) conv1 pass 2 -> dh1 (own, over dz | dh2: every CTA is past its pass 2 of conv2)
    bwd_p2<2>(rpt_s, colt_s, emap_s, y_base, LDY * 4, erec1_base, dsd1, h1s, LDY, vc, vc + 2 * NC, dh1, LDY,
              vred + warp * VECF, vred + warp * VECF + 2 * NC, n, rank);
    __syncthreads();                         // dh1 complete; h1 / s1 buffers are free
    if (more) load_conv1_side(k - 1);        // group B
    cp_commit();
    stamp();
    // (6) conv1 projection backward: dW1 = dh1^T x0 ; g = dh1 W1 + g (residual), masked by x0 > 0 for k > 0 (in place)
    WgradAcc<2 * NC, NC> dw1;
    wgrad_mma_compute<2 * NC, NC, LDY, LDX>(dh1, xs, n, dw1);
    {
      const bool mask = k > 0;
      dgrad_mma<2 * NC, NC, LDY, LDX>(
          dh1, W1, n, [&](int m, int c) { return lds2(gs + m * LDX + c); },
          [&](int m, int c, float2 v, float2 r) {
            v.x += r.x;
            v.y += r.y;
            if (mask) {
              const float2 x0 = lds2(xs + m * LDX + c);
              v = make_float2(x0.x > 0.f ? v.x : 0.f, x0.y > 0.f ? v.y : 0.f);
            }
            *reinterpret_cast<float2*>(gs + m * LDX + c) = v;
          });
    }
    // parameter-vector gradients of the block: sum the 8 warp rows (one float4 per thread, flushed behind the arrival)
    float4 vsum = f4zero();
    if (threadIdx.x < VECF / 4) {
#pragma unroll
      for (int w = 0; w < T / 32; ++w) add4(vsum, lds4(vred + w * VECF + 4 * threadIdx.x));
    }
    __syncthreads();                         // W1 / W2 / vec / vred are free again
    cp_wait_all();                           // groups A and B (what neighbours read next) have landed
    stamp();
    cb.arrive();
    // behind the arrival: nothing below is covered by its release fence, and everything has a whole phase to drain
    // before the next one — the weight-gradient atomics, the vector-gradient atomics, the next block's parameters
    wgrad_mma_flush<2 * NC, NC>(dw1, n, a.grads + pl.c1_W(k));
    if (threadIdx.x < VECF / 4 && n > 0) {
      const int c = threadIdx.x;
      const long long off = c < 6 * NC / 4 ? pl.c1_as(k) + 4 * c : pl.c2_as(k) + 4 * (c - 6 * NC / 4);
      atomicAdd(reinterpret_cast<float4*>(a.grads + off), vsum);
    }
    if (more) stage_block_params(a.params + pl.block(k - 1), W1, W2, vc);      // group C
    cp_commit();
    cluster_wait();                          // the new g and the next block's h2 / h1 / scores are visible
  }
  cp_wait_all();

  if (a.tail) {
    // encoder backward: dw[c] = sum_i g[i][c] x[i], db[c] = sum_i g[i][c]
    const int lig = lane & 7, sub = lane >> 3;
    float4 aw = f4zero(), ab = f4zero();
    for (int il = warp * 4 + sub; il < n; il += T / 8) {
      const float4 gv = lds4(gs + il * LDX + 4 * lig);
      fma4(aw, __ldg(a.x + rb + __ldg(a.perm + lo + il)), gv);
      add4(ab, gv);
    }
    cta_chunk_sum_atomic<8>(aw, red, a.grads + pl.lin0_w());
    cta_chunk_sum_atomic<8>(ab, red, a.grads + pl.lin0_b());
  } else {
    store_rows<NC, LDX>(gs, a.scratch + ro * NC, n);      // the next range of a split backward continues from here
  }
}

// ------------------------------------------------------------------------- host side
static int g_enabled = -1;
static bool enabled() {
  if (g_enabled < 0) {
    const char* e = getenv("GATRES_RESIDENT_DSM");
    g_enabled = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return g_enabled == 1;
}

static int g_barrier = -1;
static int barrier_flavour() {
  if (g_barrier < 0) {
    const char* e = getenv("GATRES_RES2_BARRIER");
    g_barrier = e ? atoi(e) : 2;
    if (g_barrier < 0 || g_barrier > 2) g_barrier = 2;
  }
  return g_barrier;
}

static int g_tc = -1;
static bool tc_enabled() {                  // tcgen05 projections in the forward stack (GATRES_RES2_TC=0: mma.sync form)
  if (g_tc < 0) {
    const char* e = getenv("GATRES_RES2_TC");
    g_tc = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return g_tc == 1;
}
static size_t fwd_tc_smem(int R, int ecap) { return sizeof(float) * (size_t)FwdTcSmem(R, ecap).total; }
static size_t fwd_smem(int R, int ecap) { return sizeof(float) * (size_t)FwdSmem(R, ecap).total; }
static size_t bwd_smem(int R, int ecap) { return sizeof(float) * (size_t)BwdSmem(R, ecap).total; }

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

}  // namespace res2

int resident_forced_cluster();      // resident.cu (gatres_set_resident_cluster)
bool resident_max_batch_is_default();
void resident_profile(long long** buf, int* slots);      // resident.cu (gatres_set_resident_profile)

// Cluster size: eight CTAs per snapshot, whatever the batch.  Clusters are independent (one snapshot each), so a batch
// that does not fit the GPU at two CTAs per SM (37 clusters) simply runs in waves; measured (tools/resident_probe.py,
// training step, us): 48 snapshots 853 against 1279 with 4 CTAs per snapshot in one wave + the layer backward, 64:
// 911 / 1503, 96: 1343 / 1591 (layer kernels), 128: 1774 / 1839, 192: 2630 / 2789, 256: 3456 / 3066 — the layer kernels
// take over between 192 and 256 snapshots (inference: between 256 and ~400).  Third session of round 2 (warp-specialised
// projections, all-tcgen05 projection backward in the layer path; bench.py --batch B, snapshots/s resident pair / layer
// kernels): 96: 70.9 k / 57.4 k, 128: 71.6 k / 71.1 k, 192: 72.6 k / 74.8 k, 256: - / 89.8 k — the training crossover moved to
// between 128 and 192 snapshots.
constexpr long long kRes2MaxTrain = 160, kRes2MaxInfer = 320;
static int res2_cluster(long long B) {
  (void)B;
  return resident_forced_cluster() > 0 ? resident_forced_cluster() : 8;
}

// the tensor-core forward (and with it the CTA-image layout of the saved activations) applies
static bool res2_tc_pair(int R, int ecap) {
  return res2::tc_enabled() && R <= 64 && res2::fwd_tc_smem(R, ecap) <= res2::kMaxSmem;
}

static int res2_ecap(const gatres_model_desc* d, int cs) { return d->p_ecap[cs == 8 ? 3 : (cs == 4 ? 2 : (cs == 2 ? 1 : 0))]; }

// Applicable when the descriptor carries a locality plan, nc = 32, atomics mode, and the slices fit shared memory.
// `training` = the forward / backward PAIR (same predicate on both sides: the saved buffer is in locality order).
bool resident2_eligible(const gatres_model_desc* d, bool training, long long max_batch) {
  if (!res2::enabled() || d->perm == nullptr || d->p_rowptr == nullptr || d->p_col == nullptr) return false;
  if (d->nc != 32 || d->E1 <= 0 || d->slots > 0) return false;
  // the caller's limit when GATRES_RESIDENT_MAX_B / gatres_set_resident_max_batch chose one, else the measured crossovers
  if (d->B > (resident_max_batch_is_default() ? (training ? kRes2MaxTrain : kRes2MaxInfer) : max_batch)) return false;
  const int cs = res2_cluster(d->B);
  const int R = (d->N + cs - 1) / cs, ecap = res2_ecap(d, cs);
  if (R >= 65536 || ecap <= 0) return false;
  // two CTAs per SM while the batch needs them, one otherwise
  const size_t budget = d->B * cs > (long long)sm_count() ? res2::kMaxSmem : 200 * 1024;
  if (res2::fwd_smem(R, ecap) > budget) return false;
  return !training || res2::bwd_smem(R, ecap) <= budget;
}

// floats the saved-activation buffer needs when the pair uses the CTA-image layout (upper bound over the cluster
// sizes the knobs can select); 0 when the kernels cannot apply
long long resident2_saved_floats(const gatres_model_desc* d) {
  if (d->nc != 32) return 0;               // (whether a locality plan is attached may change after the buffer is sized)
  long long worst = 0;
  for (int cs = 1; cs <= 8; cs <<= 1) {
    const int R = (d->N + cs - 1) / cs;
    if (R > 64) continue;
    const long long t = res2::ImageLayout(R, d->B, cs).total(d->num_blocks);
    worst = t > worst ? t : worst;
  }
  return worst;
}

int resident2_forward(const gatres_model_desc* d, const float* params, const float* x, float* out, float* saved,
                      cudaStream_t st) {
  res2::Args a = {};
  a.rowptr = d->p_rowptr; a.col = d->p_col; a.rowptr_t = d->p_rowptr_t; a.col_t = d->p_col_t; a.perm = d->perm;
  a.params = params; a.x = x; a.out = out; a.saved = saved; a.poison = d->poison;
  a.M = d->B * (long long)d->N; a.N = d->N; a.nb = d->num_blocks;
  resident_profile(&a.prof, &a.prof_slots);
  a.barrier = res2::barrier_flavour();
  const int cs = res2_cluster(d->B);
  a.R = (d->N + cs - 1) / cs;
  a.ecap = res2_ecap(d, cs);
  a.cs = cs;
  // tensor-core projections when the operand tiles fit next to a second CTA (M = 128 datapath: at most 64 rows per CTA
  // are drained) — otherwise the mma.sync form
  const size_t tc_smem = res2::fwd_tc_smem(a.R, a.ecap);
  a.images = res2_tc_pair(a.R, a.ecap) ? 1 : 0;
  if (a.images)
    return saved != nullptr ? res2::launch_cluster<res2::fwd_tc_kernel<true>>("resident2_forward_tc(train)", cs, d->B, tc_smem, st, a)
                            : res2::launch_cluster<res2::fwd_tc_kernel<false>>("resident2_forward_tc", cs, d->B, tc_smem, st, a);
  const size_t smem = res2::fwd_smem(a.R, a.ecap);
  return saved != nullptr ? res2::launch_cluster<res2::fwd_kernel<true>>("resident2_forward(train)", cs, d->B, smem, st, a)
                          : res2::launch_cluster<res2::fwd_kernel<false>>("resident2_forward", cs, d->B, smem, st, a);
}

int resident2_backward(const gatres_model_desc* d, const float* params, const float* x, const float* saved,
                       const float* d_out, float* grads, float* scratch, int k_hi, int k_lo, bool head, bool tail,
                       cudaStream_t st) {
  res2::Args a = {};
  a.rowptr = d->p_rowptr; a.col = d->p_col; a.rowptr_t = d->p_rowptr_t; a.col_t = d->p_col_t; a.perm = d->perm;
  a.params = params; a.x = x; a.saved = const_cast<float*>(saved); a.d_out = d_out; a.grads = grads; a.scratch = scratch;
  a.M = d->B * (long long)d->N; a.N = d->N; a.nb = d->num_blocks;
  a.k_hi = k_hi; a.k_lo = k_lo; a.head = head; a.tail = tail;
  resident_profile(&a.prof, &a.prof_slots);
  a.barrier = res2::barrier_flavour();
  const int cs = res2_cluster(d->B);
  a.R = (d->N + cs - 1) / cs;
  a.ecap = res2_ecap(d, cs);
  a.cs = cs;
  a.images = res2_tc_pair(a.R, a.ecap) ? 1 : 0;          // same predicate as the forward: the two agree on the layout
  return res2::launch_cluster<res2::bwd_kernel>("resident2_backward", cs, d->B, res2::bwd_smem(a.R, a.ecap), st, a);
}

}  // namespace gatres

extern "C" int gatres_set_resident_barrier(int flavour) {
  const int prev = gatres::res2::barrier_flavour();
  if (flavour >= 0 && flavour <= 2) gatres::res2::g_barrier = flavour;
  return prev;
}

extern "C" int gatres_set_resident_tc(int on) {
  const int prev = gatres::res2::tc_enabled() ? 1 : 0;
  if (on == 0 || on == 1) gatres::res2::g_tc = on;
  return prev;
}

extern "C" int gatres_set_resident_dsm(int on) {
  const int prev = gatres::res2::enabled() ? 1 : 0;
  if (on == 0 || on == 1) gatres::res2::g_enabled = on;
  return prev;
}
