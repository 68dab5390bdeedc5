// Snapshot-resident GATRes forward / backward for small batches (sm_100a).
//
// Every snapshot of a batch is an independent graph (block-diagonal collation,
// /root/reference/gnn_pressure_estimation/train.py:302), so the whole dependency
// chain of GATResMeanConv.forward (GraphModels.py:486-494) and of its autograd
// backward (train.py:185) is local to one snapshot.  At the headline batch
// (32 snapshots x 388 nodes) every layer is a few microseconds of work and the
// layer-by-layer path is bound by ~220 kernel boundaries per step.  Here ONE
// thread-block cluster owns one snapshot for the entire stack: each CTA of the
// cluster keeps its slice of the rows, the layer inputs that are row-local
// (block input x0, conv1 output y1, the gradients feeding the two projections)
// stay in shared memory, the tensors neighbours must see (h1, h2, z, and the
// gradient rows in the backward) go through L2, and layers are separated by
// cluster barriers (release/acquire) instead of kernel launches.  The per-block
// parameters (4 384 floats for nc = 32) are double-buffered in shared memory with
// cp.async one block ahead.
//
// Same arithmetic as the layer kernels (linear.cu, gat_agg.cu, mean_res.cu): fp32
// FFMA projections with k ascending, cooperative per-row softmax, saved (m, l) for
// the recompute backward, atomic accumulation of parameter gradients.  The saved
// activation layout is the one of model.cu, so forward and backward may be mixed
// freely with the layer-by-layer path.
#include <math_constants.h>
#include <stdlib.h>
#include "common.cuh"
#include "layout.cuh"

namespace gatres {
namespace res {
// Largest batch that takes the resident path.  Measured on B200 (tools/resident_probe.py, C-Town-shaped graph):
// the cluster kernels win while a batch fits one wave of 4-CTA clusters at two CTAs per SM (32 snapshots: 0.70 ms
// per training step against 1.19 ms layer by layer; 64: 1.30 against 1.66 ms); from 128 snapshots on the
// layer-by-layer kernels are faster (2.23 against 2.59 ms).
static long long g_max_batch = -1;
static long long max_batch() {
  if (g_max_batch < 0) {
    const char* e = getenv("GATRES_RESIDENT_MAX_B");
    g_max_batch = e ? atoll(e) : (2ll * sm_count()) / 4;
  }
  return g_max_batch;
}

// cluster size: as many CTAs per snapshot as keeps the whole batch co-resident (2 CTAs per SM), power of two <= 8
constexpr size_t kMaxSmem = 200 * 1024;    // per-CTA shared-memory budget of the resident kernels
static int g_forced_cluster = -2;
static int forced_cluster() {
  if (g_forced_cluster == -2) {
    const char* e = getenv("GATRES_RESIDENT_CLUSTER");
    g_forced_cluster = e ? atoi(e) : 0;
    if (g_forced_cluster != 1 && g_forced_cluster != 2 && g_forced_cluster != 4 && g_forced_cluster != 8) g_forced_cluster = 0;
  }
  return g_forced_cluster;
}
static long long* g_prof = nullptr;       // phase-timestamp buffer of the profiling tool (NULL = off)
static int g_prof_slots = 0;
static int g_threads = -1;
static int threads() {
  if (g_threads < 0) {
    const char* e = getenv("GATRES_RESIDENT_THREADS");
    g_threads = e ? atoi(e) : 256;
    if (g_threads != 256 && g_threads != 255) g_threads = 256;
  }
  return g_threads;
}
}  // namespace res
}  // namespace gatres

#define RES_NS res256
#define RES_T 256
#define RES_MIN_CTAS 2
#define RES_USE_MMA 1            // projections / data / weight gradients on mma.sync 3xTF32 (resident_mma.cuh)
#include "resident_impl.cuh"
#undef RES_NS
#undef RES_USE_MMA

#define RES_NS res256f
#define RES_USE_MMA 0            // the same kernels with fp32 FFMA contractions (A/B comparison, gatres_set_resident_threads(255))
#include "resident_impl.cuh"
#undef RES_NS
#undef RES_T
#undef RES_MIN_CTAS
#undef RES_USE_MMA

namespace gatres {

// Is the snapshot-resident path applicable?  nc = 32, the network has its self-loop CSR (E1 > 0), a batch small
// enough that clusters cover it in about one wave, and a graph whose CSR + per-CTA row slice fit shared memory.
bool resident_eligible(const gatres_model_desc* d, bool backward) {
  if (d->nc != 32 || d->E1 <= 0 || d->B > res::max_batch() || d->B * 8ll >= (1ll << 31)) return false;
  if (backward && d->slots > 0) return false;                       // deterministic two-stage reduction: layer path
  // the backward stack only wins with 8 CTAs per snapshot (49 rows per CTA); with 4 (batches of 38..74) its row passes
  // double and the layer-by-layer backward is faster, while the forward stack still wins (measured at 40 / 48 / 64
  // snapshots: forward 312-319 us against 369-430, backward 1053-1082 us against 921-1145; profiles/r1_resident.md)
  if (backward && res::forced_cluster() == 0 && d->B * 8ll > 2ll * sm_count()) return false;
  return res256::fits(d->N, d->E1, backward);                       // res256f needs less
}

int resident_forced_cluster() { return res::forced_cluster(); }
long long resident_max_batch() { return res::max_batch(); }
// the knob still holds its built-in value (a caller that restores the value it read keeps the defaults): the second
// generation then applies its own measured crossovers (resident2.cu)
bool resident_max_batch_is_default() { return res::max_batch() == (2ll * sm_count()) / 4; }
void resident_profile(long long** buf, int* slots) { *buf = res::g_prof; *slots = res::g_prof_slots; }

int resident_forward(const gatres_model_desc* d, const float* params, const float* x, float* out, float* saved,
                     float* scratch, cudaStream_t st) {
  if (res::threads() == 255) return res256f::forward(d, params, x, out, saved, scratch, st);
  return res256::forward(d, params, x, out, saved, scratch, st);
}

int resident_backward(const gatres_model_desc* d, const float* params, const float* x, const float* saved,
                      const float* d_out, float* grads, float* scratch, int k_hi, int k_lo, bool head, bool tail,
                      cudaStream_t st) {
  if (res::threads() == 255)
    return res256f::backward(d, params, x, saved, d_out, grads, scratch, k_hi, k_lo, head, tail, st);
  return res256::backward(d, params, x, saved, d_out, grads, scratch, k_hi, k_lo, head, tail, st);
}

}  // namespace gatres

extern "C" int32_t gatres_set_resident_cluster(int32_t ctas) {
  const int prev = gatres::res::forced_cluster();
  if (ctas == 0 || ctas == 1 || ctas == 2 || ctas == 4 || ctas == 8) gatres::res::g_forced_cluster = ctas;
  return prev;
}

extern "C" void gatres_set_resident_profile(int64_t* device_buf, int32_t slots_per_cta) {
  gatres::res::g_prof = reinterpret_cast<long long*>(device_buf);
  gatres::res::g_prof_slots = device_buf != nullptr ? slots_per_cta : 0;
}

extern "C" int32_t gatres_set_resident_threads(int32_t threads) {
  const int prev = gatres::res::threads();
  if (threads == 256 || threads == 255) gatres::res::g_threads = threads;
  return prev;
}

extern "C" int64_t gatres_set_resident_max_batch(int64_t max_batch) {
  const long long prev = gatres::res::max_batch();
  if (max_batch >= 0) gatres::res::g_max_batch = max_batch;
  return prev;
}
