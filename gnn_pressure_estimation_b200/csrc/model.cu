// Whole-model orchestration: GATResMeanConv.forward and its backward as one C
// call each (/root/reference/gnn_pressure_estimation/GraphModels.py:486-494 for
// the model, :462-468 for a block; backward = what autograd does at
// train.py:185).  One call enqueues every kernel of the stack on the caller's
// stream, so the Python host pays two FFI crossings per training step and the
// whole step can be captured into a CUDA graph.
#include "common.cuh"
#include "layout.cuh"

namespace gatres {

// snapshot-resident cluster kernels for small batches (resident.cu)
bool resident_eligible(const gatres_model_desc* d, bool backward);
int resident_forward(const gatres_model_desc* d, const float* params, const float* x, float* out, float* saved,
                     float* scratch, cudaStream_t st);
int resident_backward(const gatres_model_desc* d, const float* params, const float* x, const float* saved,
                      const float* d_out, float* grads, float* scratch, int k_hi, int k_lo, bool head, bool tail,
                      cudaStream_t st);

// second generation (resident2.cu): exchange tensors in distributed shared memory, rows in a locality order
bool resident2_eligible(const gatres_model_desc* d, bool training, long long max_batch);
long long resident2_saved_floats(const gatres_model_desc* d);
long long resident_max_batch();
int resident2_forward(const gatres_model_desc* d, const float* params, const float* x, float* out, float* saved,
                      cudaStream_t st);
int resident2_backward(const gatres_model_desc* d, const float* params, const float* x, const float* saved,
                       const float* d_out, float* grads, float* scratch, int k_hi, int k_lo, bool head, bool tail,
                       cudaStream_t st);

// conv2 aggregation + SimpleConv(mean) + residual + ReLU in one launch when the snapshot tile path applies (gat_agg.cu)
int gat_agg_mean_res_fwd(const int* rowptr, const int* col, unsigned E1, const float* h, const float* s_src,
                         const float* s_dst, const float* bias, float* m, float* l, const float* x0, float* xout,
                         long long B, int N, int C, cudaStream_t st);

static int validate(const gatres_model_desc* d, const char* who) {
  GATRES_REQUIRE(d != nullptr, "%s: null model descriptor", who);
  GATRES_REQUIRE(d->num_blocks >= 0 && (d->nc == 32 || d->nc == 64 || d->nc == 128),
                 "%s: unsupported model (num_blocks=%d, nc=%d; nc in {32,64,128})", who, d->num_blocks, d->nc);
  GATRES_REQUIRE(d->N > 0 && d->B > 0 && d->B * (int64_t)d->N < (1ll << 31), "%s: bad B=%lld N=%d", who,
                 (long long)d->B, d->N);
  GATRES_REQUIRE(d->rowptr && d->col && d->rowptr_t && d->col_t, "%s: CSR pointers missing", who);
  return GATRES_OK;
}

}  // namespace gatres

using namespace gatres;

#define TRY(call)            \
  do {                       \
    const int rc_ = (call);  \
    if (rc_) return rc_;     \
  } while (0)

extern "C" size_t gatres_model_desc_bytes(void) { return sizeof(gatres_model_desc); }

extern "C" int64_t gatres_param_count(int32_t num_blocks, int32_t nc) { return ParamLayout(num_blocks, nc).count(); }

extern "C" int64_t gatres_saved_floats(const gatres_model_desc* d) {
  if (validate(d, "saved_floats")) return -1;
  // the snapshot-resident tensor-core pair keeps CTA images of its shared-memory arrays (~5 % larger: resident2.cu)
  const int64_t layer = SavedLayout(d->B * (int64_t)d->N, d->nc).total(d->num_blocks), images = resident2_saved_floats(d);
  return layer > images ? layer : images;
}

extern "C" int64_t gatres_scratch_floats(const gatres_model_desc* d, int32_t training) {
  if (validate(d, "scratch_floats")) return -1;
  const int64_t M = d->B * (int64_t)d->N, nc = d->nc;
  const int64_t fwd_infer = 8 * M * nc + 2 * a4(2 * M);                 // xa xb h1 y1 h2 z + ss sd
  const int64_t fwd_train = M * nc;                                     // z
  const int64_t bwd = 8 * M * nc + 8 * M + a4(2 * M);                   // gA gB dz dh2 dy1 dh1 rec dsd
  return training ? (fwd_train > bwd ? fwd_train : bwd) : fwd_infer;
}

extern "C" int gatres_forward(const gatres_model_desc* d, const float* params, const float* x, float* out,
                              float* saved, float* scratch, void* stream) {
  TRY(validate(d, "forward"));
  GATRES_REQUIRE(params && x && out && scratch, "forward: null buffer");
  const int64_t M = d->B * (int64_t)d->N, nc = d->nc;
  const int32_t N = d->N, C = d->nc;
  const ParamLayout pl(d->num_blocks, d->nc);
  const SavedLayout sl(M, nc);
  const bool train = saved != nullptr;
  if (resident2_eligible(d, train, resident_max_batch())) return resident2_forward(d, params, x, out, saved, as_stream(stream));
  if (resident_eligible(d, false)) return resident_forward(d, params, x, out, saved, scratch, as_stream(stream));

  // inference: rolling buffers carved from scratch
  float* xa = scratch;
  float* xb = xa + M * nc;
  float* h1_i = xb + M * nc;
  float* y1_i = h1_i + 2 * M * nc;
  float* h2_i = y1_i + 2 * M * nc;
  float* z_i = h2_i + M * nc;
  float* ss_i = z_i + M * nc;
  float* sd_i = ss_i + a4(2 * M);

  float* x0 = train ? saved + sl.x_enc() : xa;
  TRY(gatres_encoder_fwd(x, params + pl.lin0_w(), params + pl.lin0_b(), x0, M, C, stream));
  for (int k = 0; k < d->num_blocks; ++k) {
    float* h1 = train ? saved + sl.h1(k) : h1_i;
    float* ss1 = train ? saved + sl.ss1(k) : ss_i;
    float* sd1 = train ? saved + sl.sd1(k) : sd_i;
    float* m1 = train ? saved + sl.m1(k) : nullptr;
    float* l1 = train ? saved + sl.l1(k) : nullptr;
    float* y1 = train ? saved + sl.y1(k) : y1_i;
    float* h2 = train ? saved + sl.h2(k) : h2_i;
    float* ss2 = train ? saved + sl.ss2(k) : ss_i;
    float* sd2 = train ? saved + sl.sd2(k) : sd_i;
    float* m2 = train ? saved + sl.m2(k) : nullptr;
    float* l2 = train ? saved + sl.l2(k) : nullptr;
    float* z = train ? scratch : z_i;
    float* xo = train ? saved + sl.xout(k) : (x0 == xa ? xb : xa);
    TRY(gatres_linear_att_fwd(x0, params + pl.c1_W(k), params + pl.c1_as(k), params + pl.c1_ad(k), h1, ss1, sd1, M,
                              C, 2, C, stream));
    TRY(gatres_gat_agg_fwd(d->rowptr, d->col, h1, ss1, sd1, params + pl.c1_b(k), y1, m1, l1, d->B, N, d->E1, 2, C, 1, stream));
    TRY(gatres_linear_att_fwd(y1, params + pl.c2_W(k), params + pl.c2_as(k), params + pl.c2_ad(k), h2, ss2, sd2, M,
                              2 * C, 1, C, stream));
    const int fused = gat_agg_mean_res_fwd(d->rowptr, d->col, (unsigned)d->E1, h2, ss2, sd2, params + pl.c2_b(k), m2, l2, x0, xo,
                                           d->B, N, C, as_stream(stream));
    if (fused < 0) return fused;
    if (fused == 0) {
      TRY(gatres_gat_agg_fwd(d->rowptr, d->col, h2, ss2, sd2, params + pl.c2_b(k), z, m2, l2, d->B, N, d->E1, 1, C, 0, stream));
      TRY(gatres_mean_res_fwd(d->rowptr, d->col, z, x0, xo, d->B, N, C, stream));
    }
    x0 = xo;
  }
  return gatres_decoder_fwd(x0, params + pl.lin1_w(), params + pl.lin1_b(), out, d->poison, M, C, stream);
}

// Backward of blocks k_hi .. k_lo (descending).  The first range (k_hi == num_blocks-1) also runs the
// decoder backward and zeroes the gradient buffer (atomic mode); the last one (k_lo == 0) also runs the
// encoder backward and, in deterministic mode, the final reduction.  The gradient w.r.t. a block's output
// lives in scratch slot (num_blocks-1-k) & 1, so consecutive ranges chain without extra state.
extern "C" int gatres_backward_range(const gatres_model_desc* d, const float* params, const float* x,
                                     const float* saved, const float* d_out, float* partial, float* grads,
                                     float* scratch, int32_t k_hi, int32_t k_lo, void* stream) {
  TRY(validate(d, "backward"));
  GATRES_REQUIRE(params && x && saved && d_out && grads && scratch, "backward: null buffer");
  GATRES_REQUIRE(d->slots <= 0 || partial != nullptr, "backward: deterministic mode (slots > 0) needs `partial`");
  const int64_t M = d->B * (int64_t)d->N, nc = d->nc;
  const int32_t N = d->N, C = d->nc, S = d->slots, nb = d->num_blocks;
  GATRES_REQUIRE(nb == 0 || (k_hi < nb && k_lo >= 0 && k_lo <= k_hi), "backward_range: bad block range [%d, %d]", k_lo,
                 k_hi);
  const bool head = nb == 0 || k_hi == nb - 1, tail = nb == 0 || k_lo == 0;
  const ParamLayout pl(nb, d->nc);
  const SavedLayout sl(M, nc);
  const int64_t P = a4(pl.count());                      // row stride of `partial`
  if (S <= 0) {
    // atomic mode: every kernel adds its contribution straight into `grads`
    if (head && cudaMemsetAsync(grads, 0, (size_t)pl.count() * sizeof(float), as_stream(stream)) != cudaSuccess)
      return check_launch("backward: zero grads");
    partial = grads;
  }
  if (resident2_eligible(d, true, resident_max_batch()))      // the same predicate gatres_forward(training) evaluated
    return resident2_backward(d, params, x, saved, d_out, grads, scratch, nb > 0 ? k_hi : -1, nb > 0 ? k_lo : 0, head,
                              tail, as_stream(stream));
  if (resident_eligible(d, true))
    return resident_backward(d, params, x, saved, d_out, grads, scratch, nb > 0 ? k_hi : -1, nb > 0 ? k_lo : 0, head,
                             tail, as_stream(stream));

  float* gbuf[2] = {scratch, scratch + M * nc};
  float* dz = scratch + 2 * M * nc;
  float* dh2 = dz + M * nc;
  float* dy1 = dh2 + M * nc;
  float* dh1 = dy1 + 2 * M * nc;
  float* rec = dh1 + 2 * M * nc;
  float* dsd = rec + 8 * M;

  if (head) {
    const float* x_last = nb > 0 ? saved + sl.xout(nb - 1) : saved + sl.x_enc();
    TRY(gatres_decoder_bwd(d_out, x_last, params + pl.lin1_w(), gbuf[0], partial, P, S, pl.lin1_w(), pl.lin1_b(), M, C,
                           nb > 0 ? 1 : 0, stream));
  }
  for (int k = nb > 0 ? k_hi : -1; k >= k_lo && k >= 0; --k) {
    float* gA = gbuf[(nb - 1 - k) & 1];
    float* gB = gbuf[(nb - k) & 1];
    const float* x0 = k > 0 ? saved + sl.xout(k - 1) : saved + sl.x_enc();
    // SimpleConv(mean) + residual: gA already carries the ReLU mask of this block's output
    TRY(gatres_mean_res_bwd_e1(d->rowptr, d->rowptr_t, d->col_t, d->E1, gA, dz, d->B, N, C, stream));
    // conv2 (heads=1, mean over one head)
    TRY(gatres_gat_agg_bwd(d->rowptr, d->col, d->rowptr_t, d->col_t, dz, saved + sl.h2(k), saved + sl.ss2(k),
                           saved + sl.sd2(k), saved + sl.m2(k), saved + sl.l2(k), params + pl.c2_as(k),
                           params + pl.c2_ad(k), rec, dsd, dh2, partial, P, S, pl.c2_as(k), pl.c2_ad(k), pl.c2_b(k),
                           d->B, N, d->E1, 1, C, stream));
    TRY(gatres_linear_bwd(dh2, saved + sl.y1(k), params + pl.c2_W(k), nullptr, saved + sl.y1(k), dy1, partial, P, S,
                          pl.c2_W(k), M, 2 * C, 1, C, stream));
    // conv1 (heads=2, concat) — dy1 already carries the ReLU mask of y1
    TRY(gatres_gat_agg_bwd(d->rowptr, d->col, d->rowptr_t, d->col_t, dy1, saved + sl.h1(k), saved + sl.ss1(k),
                           saved + sl.sd1(k), saved + sl.m1(k), saved + sl.l1(k), params + pl.c1_as(k),
                           params + pl.c1_ad(k), rec, dsd, dh1, partial, P, S, pl.c1_as(k), pl.c1_ad(k), pl.c1_b(k),
                           d->B, N, d->E1, 2, C, stream));
    // dx0 = dh1 W1 + (residual branch gA), masked by the previous block's ReLU (none before block 0)
    TRY(gatres_linear_bwd(dh1, x0, params + pl.c1_W(k), gA, k > 0 ? x0 : nullptr, gB, partial, P, S, pl.c1_W(k), M, C,
                          2, C, stream));
  }
  if (!tail) return GATRES_OK;
  TRY(gatres_encoder_bwd(gbuf[nb & 1], x, partial, P, S, pl.lin0_w(), pl.lin0_b(), M, C, stream));
  if (S <= 0) return GATRES_OK;
  return gatres_reduce_partials(partial, P, S, 0, pl.count(), grads, stream);
}

extern "C" int64_t gatres_param_offset_of_block(int32_t num_blocks, int32_t nc, int32_t k) {
  const ParamLayout pl(num_blocks, nc);
  return k >= num_blocks ? pl.lin1_w() : (k < 0 ? 0 : pl.block(k));
}

extern "C" int gatres_backward(const gatres_model_desc* d, const float* params, const float* x, const float* saved,
                               const float* d_out, float* partial, float* grads, float* scratch, void* stream) {
  TRY(validate(d, "backward"));
  return gatres_backward_range(d, params, x, saved, d_out, partial, grads, scratch, d->num_blocks - 1, 0, stream);
}
