// Shared-topology indexing (sm_100a): one destination-sorted CSR (and its
// transpose) per water-network template, built on the device once per .inp,
// bit-exact with a stable sort of the GATConv-rewritten edge list
// (replaces the per-call remove_self_loops /
// add_self_loops / scatter indexing inside PyG's GATConv and SimpleConv, call
// sites /root/reference/gnn_pressure_estimation/GraphModels.py:464-466).
//
// Stable order without a sort of the whole list: count -> exclusive scan ->
// unordered atomic fill of EDGE IDS -> each row sorts its few ids ascending
// (WDN degrees are <= ~10) -> ids are replaced by the neighbour they name and
// the self-loop is appended last.  Deterministic regardless of atomic order.
#include "common.cuh"

namespace gatres {

__global__ void csr_zero_kernel(int* a, size_t na, int* b, size_t nb, int* c, size_t nc, int* info) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = i; k < na; k += stride) a[k] = 0;
  for (size_t k = i; k < nb; k += stride) b[k] = 0;
  for (size_t k = i; k < nc; k += stride) c[k] = 0;
  if (i < 4) info[i] = 0;
}

// keep_loops = 0: GATConv's view (existing self loops dropped); 1: SimpleConv's view (they stay ordinary edges)
__global__ void csr_count_kernel(const long long* __restrict__ ei, long long E, int N, int* cnt_in, int* cnt_out,
                                 int* info, int keep_loops) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    const long long s = ei[e], d = ei[E + e];
    if (s < 0 || s >= N || d < 0 || d >= N) { atomicAdd(info + 2, 1); continue; }
    if (s == d) {
      atomicAdd(info + 0, 1);
      if (!keep_loops) continue;
    }
    atomicAdd(cnt_in + d, 1);
    atomicAdd(cnt_out + s, 1);
  }
}

// in-place exclusive scan of (count + 1) over N entries; a[N] = total. One CTA.
__global__ void __launch_bounds__(1024) csr_scan_kernel(int* a, int* b, int N, int* info) {
  __shared__ int sums[1024];
  for (int which = 0; which < 2; ++which) {
    int* p = which ? b : a;
    const int per = (N + 1023) / 1024;
    const int lo = min(N, (int)threadIdx.x * per), hi = min(N, lo + per);
    int s = 0;
    for (int k = lo; k < hi; ++k) s += p[k] + 1;
    sums[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {                    // inclusive Hillis-Steele
      const int t = threadIdx.x >= off ? sums[threadIdx.x - off] : 0;
      __syncthreads();
      sums[threadIdx.x] += t;
      __syncthreads();
    }
    int run = threadIdx.x ? sums[threadIdx.x - 1] : 0;
    for (int k = lo; k < hi; ++k) {
      const int c = p[k] + 1;
      p[k] = run;
      run += c;
    }
    if (threadIdx.x == 1023) {
      p[N] = sums[1023];
      if (which == 0) info[1] = sums[1023];
    }
    __syncthreads();
  }
}

__global__ void csr_fill_kernel(const long long* __restrict__ ei, long long E, int N,
                                const int* __restrict__ rowptr, const int* __restrict__ rowptr_t, int* cur_in,
                                int* cur_out, int* col, int* col_t, int keep_loops) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    const long long s = ei[e], d = ei[E + e];
    if (s < 0 || s >= N || d < 0 || d >= N || (s == d && !keep_loops)) continue;
    col[rowptr[d] + atomicAdd(cur_in + d, 1)] = (int)e;
    col_t[rowptr_t[s] + atomicAdd(cur_out + s, 1)] = (int)e;
  }
}

__global__ void csr_finish_kernel(const long long* __restrict__ ei, long long E, int N,
                                  const int* __restrict__ rowptr, const int* __restrict__ rowptr_t, int* col,
                                  int* col_t) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    for (int which = 0; which < 2; ++which) {
      int* c = which ? col_t : col;
      const int beg = (which ? rowptr_t : rowptr)[i], end = (which ? rowptr_t : rowptr)[i + 1] - 1;
      for (int a = beg + 1; a < end; ++a) {                       // insertion sort of edge ids
        const int key = c[a];
        int k = a - 1;
        while (k >= beg && c[k] > key) { c[k + 1] = c[k]; --k; }
        c[k + 1] = key;
      }
      const long long* nb = which ? ei + E : ei;                  // in-edges name sources, out-edges name targets
      for (int a = beg; a < end; ++a) c[a] = (int)nb[c[a]];
      c[end] = i;                                                 // the appended self-loop
    }
  }
}

__global__ void check_replicated_kernel(const long long* __restrict__ eb, const long long* __restrict__ et,
                                        long long B, long long E, long long N, int* mismatch) {
  const long long total = B * E;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / E, e = idx - b * E;
    if (eb[idx] != et[e] + b * N || eb[total + idx] != et[E + e] + b * N) atomicAdd(mismatch, 1);
  }
}

}  // namespace gatres

using namespace gatres;

extern "C" size_t gatres_csr_scratch_bytes(int64_t E, int32_t N) {
  (void)E;
  return (size_t)(2 * (int64_t)N + 8) * sizeof(int32_t);
}

static int csr_build_impl(const int64_t* edge_index, int64_t E, int32_t N, int32_t* rowptr, int32_t* col,
                          int32_t* rowptr_t, int32_t* col_t, int32_t* info, void* scratch, size_t scratch_bytes,
                          void* stream, int keep_loops) {
  GATRES_REQUIRE(N > 0 && E >= 0 && E + N < (1ll << 31), "csr_build: bad N=%d E=%lld", N, (long long)E);
  GATRES_REQUIRE(scratch_bytes >= gatres_csr_scratch_bytes(E, N), "csr_build: scratch too small");
  cudaStream_t st = as_stream(stream);
  int* cur_in = static_cast<int*>(scratch);
  int* cur_out = cur_in + N;
  const long long* ei = reinterpret_cast<const long long*>(edge_index);
  const unsigned ge = (unsigned)((E + 255) / 256 > 0 ? ((E + 255) / 256 < 4096 ? (E + 255) / 256 : 4096) : 1);
  const unsigned gn = (unsigned)(((long long)N + 255) / 256 < 4096 ? ((long long)N + 255) / 256 : 4096);
  csr_zero_kernel<<<gn, 256, 0, st>>>(rowptr, (size_t)N + 1, rowptr_t, (size_t)N + 1, cur_in, (size_t)2 * N, info);
  csr_count_kernel<<<ge, 256, 0, st>>>(ei, E, N, rowptr, rowptr_t, info, keep_loops);
  csr_scan_kernel<<<1, 1024, 0, st>>>(rowptr, rowptr_t, N, info);
  csr_fill_kernel<<<ge, 256, 0, st>>>(ei, E, N, rowptr, rowptr_t, cur_in, cur_out, col, col_t, keep_loops);
  csr_finish_kernel<<<gn, 256, 0, st>>>(ei, E, N, rowptr, rowptr_t, col, col_t);
  return check_launch("csr_build");
}

extern "C" int gatres_csr_build(const int64_t* edge_index, int64_t E, int32_t N, int32_t* rowptr, int32_t* col,
                                int32_t* rowptr_t, int32_t* col_t, int32_t* info, void* scratch,
                                size_t scratch_bytes, void* stream) {
  return csr_build_impl(edge_index, E, N, rowptr, col, rowptr_t, col_t, info, scratch, scratch_bytes, stream, 0);
}

extern "C" int gatres_csr_build_mean(const int64_t* edge_index, int64_t E, int32_t N, int32_t* rowptr, int32_t* col,
                                     int32_t* rowptr_t, int32_t* col_t, int32_t* info, void* scratch,
                                     size_t scratch_bytes, void* stream) {
  return csr_build_impl(edge_index, E, N, rowptr, col, rowptr_t, col_t, info, scratch, scratch_bytes, stream, 1);
}

extern "C" int gatres_check_replicated(const int64_t* edge_index_batch, const int64_t* edge_index_tmpl, int64_t B,
                                       int64_t E, int32_t N, int32_t* mismatch, void* stream) {
  GATRES_REQUIRE(B > 0 && E >= 0 && N > 0, "check_replicated: bad B/E/N");
  if (E == 0) return GATRES_OK;
  const long long total = B * E;
  const unsigned grid = (unsigned)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192);
  check_replicated_kernel<<<grid, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const long long*>(edge_index_batch), reinterpret_cast<const long long*>(edge_index_tmpl), B, E,
      N, mismatch);
  return check_launch("check_replicated");
}
