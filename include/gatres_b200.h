/*
 * gatres_b200.h — C ABI of libgatres_b200.so: the GATRes message-passing hot
 * path (stacked GATConv-with-residual blocks, forward + backward) as hand
 * written CUDA for sm_100a.
 *
 * The reference has no FFI of its own: its operator boundary is the set of
 * PyTorch-Geometric modules it instantiates (all paths relative to
 * /root/reference/gnn_pressure_estimation/):
 *
 *   GATConv(in, hc, 2, concat=True) / GATConv(2hc, out, 1, concat=False)
 *                                     GraphModels.py:458-459, called :464-465
 *   SimpleConv(aggr="mean") + x_0, relu     GraphModels.py:460, :466-467
 *   Linear(1, nc) / Linear(nc, 1)           GraphModels.py:477/:487, :484/:492
 *   GATResMeanConv.forward                  GraphModels.py:486-494
 *   autograd backward of all of the above   train.py:185
 *   PyG collation (replicated edge_index)   train.py:302, utils/DataLoader.py:28-37
 *   mask + MSE over masked nodes, Adam      train.py:171-188, :348
 *
 * Each entry point below names the interface it replaces.  INTEGRATION.md
 * shows the ctypes binding (the Python host side in
 * gnn_pressure_estimation_b200/_lib.py is exactly that binding).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - all buffers are caller-allocated (torch owns the memory); the library
 *     never allocates, frees or synchronises, keeps no global state except a
 *     thread-local error string, and enqueues all work on `stream`
 *     (a cudaStream_t passed as void*), so every call is CUDA-graph capturable;
 *   - return value 0 = ok, negative = error (gatres_last_error() has the text);
 *   - features are fp32, row-major [B*N, F] (snapshot-major: row = b*N + node),
 *     heads concatenated along F ([B, N, H, C] contiguous);
 *   - graph structure is int32 CSR, shared by all B snapshots of a batch.
 *
 * Flat parameter layout (P = gatres_param_count(num_blocks, nc) floats; grads
 * use the same layout).  Matches the state_dict order of SURVEY.md §A.1:
 *     lin0.weight[nc] lin0.bias[nc]
 *     per block k: conv1.W[2nc,nc] conv1.att_src[2nc] conv1.att_dst[2nc] conv1.bias[2nc]
 *                  conv2.W[nc,2nc] conv2.att_src[nc]  conv2.att_dst[nc]  conv2.bias[nc]
 *     lin1.weight[nc] lin1.bias[1]
 */
#ifndef GATRES_B200_H_
#define GATRES_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GATRES_OK 0
#define GATRES_ERR_ARG (-1)       /* bad argument / unsupported shape */
#define GATRES_ERR_CUDA (-2)      /* a CUDA runtime call failed        */

#define GATRES_ABI_VERSION 3

int gatres_abi_version(void);
const char* gatres_last_error(void);
/* Number of SMs of the current device (grid sizing); <0 on error. */
int gatres_sm_count(void);
/* Kernels this library has launched so far in this process (step/forward kernels; the one-off CSR build is
 * not counted).  Benchmarks difference it around one step to report launches per step. */
int64_t gatres_launch_count(void);

/* ------------------------------------------------------------------ graph */

/* Bytes of device scratch gatres_csr_build needs. */
size_t gatres_csr_scratch_bytes(int64_t E, int32_t N);

/*
 * Build the two CSR structures all kernels share from a template edge_index
 * (int64 [2,E], row 0 = source j, row 1 = target i; the layout PyG's
 * from_networkx emits, utils/DataLoader.py:28-37).  Replaces the per-call
 * remove_self_loops/add_self_loops + scatter indexing inside GATConv.forward.
 * Existing self loops are dropped, one loop per node is appended LAST, and the
 * list is stably sorted by target (rowptr/col = in-edges: sources of each
 * target, in edge-list order, self-loop last) and by source
 * (rowptr_t/col_t = out-edges: targets of each source, self-loop last).
 * rowptr* have N+1 entries, col* have capacity E+N; E' = rowptr[N].
 * info (device int32[4]): [0] dropped self loops, [1] E', [2] columns with an
 * out-of-range node id (ignored), [3] reserved.  scratch: gatres_csr_scratch_bytes.
 */
int gatres_csr_build(const int64_t* edge_index, int64_t E, int32_t N,
                     int32_t* rowptr, int32_t* col, int32_t* rowptr_t, int32_t* col_t,
                     int32_t* info, void* scratch, size_t scratch_bytes, void* stream);

/*
 * SimpleConv's view of the same template (GraphModels.py:466): SimpleConv(aggr="mean") aggregates
 * over the ORIGINAL edge_index, so a self loop (i,i) of the template is an ordinary in-edge of i there, while
 * GATConv drops it and adds its own.  Same construction and outputs as gatres_csr_build, except that existing
 * self loops are kept as ordinary entries (info[0] still counts them); the appended trailing entry per row is
 * what gatres_mean_res_fwd / _bwd skip, so these arrays are passed to them unchanged.  Only needed when
 * info[0] > 0 (WDN templates come from simple graphs and have none; then both views coincide).
 */
int gatres_csr_build_mean(const int64_t* edge_index, int64_t E, int32_t N,
                          int32_t* rowptr, int32_t* col, int32_t* rowptr_t, int32_t* col_t,
                          int32_t* info, void* scratch, size_t scratch_bytes, void* stream);

/*
 * Check that a collated edge_index (int64 [2, B*E]) is B shifted copies of the
 * template (PyG Batch collation, train.py:302): column b*E+e == template[:,e] + b*N.
 * *mismatch (device int32) is incremented once per violating column.
 */
int gatres_check_replicated(const int64_t* edge_index_batch, const int64_t* edge_index_tmpl,
                            int64_t B, int64_t E, int32_t N, int32_t* mismatch, void* stream);

/*
 * Kernel-selection knob: batches of at least `min_batch` snapshots use the TMA-staged
 * snapshot-tile aggregation kernels when the graph fits (default 64, or the
 * GATRES_TILE_MIN_B environment variable); smaller batches use the gather kernels.
 * Pass a negative value to only query.  Returns the previous value.
 */
int64_t gatres_set_tile_min_batch(int64_t min_batch);

/*
 * Kernel-selection knob for gatres_forward / gatres_backward[_range]: batches of at most `max_batch`
 * snapshots (nc = 32, gradient mode slots = 0 for the backward, graph slice fits shared memory) run the
 * snapshot-resident cluster kernels — one thread-block cluster per snapshot carries the whole stack, layers
 * separated by cluster barriers instead of kernel launches; larger batches run layer by layer.  While the knob holds
 * its built-in value (SM count / 2 = 74 on B200; GATRES_RESIDENT_MAX_B presets it) the second-generation kernels
 * (locality plan attached) apply their own measured crossovers: 8 CTAs per snapshot whatever the batch — clusters
 * that do not fit the GPU at once run in waves — up to 208 snapshots for training and 320 for inference; the first
 * generation (no plan) keeps 74.  Any other value is the limit for both generations; 0 disables.  Negative = query
 * only.  Returns the previous value.
 */
int64_t gatres_set_resident_max_batch(int64_t max_batch);
/* CTAs per snapshot cluster of the resident kernels: 1, 2, 4 or 8; 0 = automatic (as many as keeps the batch
 * co-resident at two CTAs per SM; GATRES_RESIDENT_CLUSTER presets it).  Other values only query.  Returns
 * the previous setting. */
int32_t gatres_set_resident_cluster(int32_t ctas);
/* Variant of the resident kernels (256-thread CTAs): 256 = tensor-core (mma.sync 3xTF32) contractions (default);
 * 255 = the same kernels with fp32 FFMA contractions, kept for A/B measurements (GATRES_RESIDENT_THREADS presets
 * it).  Other values only query.  Returns the previous setting.  (A 512-thread / 64-register variant was measured
 * and dropped: it halves the row passes but not the time, profiles/r1_resident.md.) */
int32_t gatres_set_resident_threads(int32_t threads);
/* Profiling aid (tools/resident_probe.py): when device_buf is not NULL, thread 0 of CTA c of every resident kernel
 * writes %globaltimer at its phase boundaries to device_buf[c * slots_per_cta ...] (at most slots_per_cta stamps).
 * NULL switches it off (default). */
void gatres_set_resident_profile(int64_t* device_buf, int32_t slots_per_cta);

/* Second generation of the resident kernels (exchange tensors in distributed shared memory; needs the locality plan
 * of gatres_model_desc): 1 = use it when applicable (default; GATRES_RESIDENT_DSM presets it), 0 = never.  Other
 * values only query.  Returns the previous setting. */
int gatres_set_resident_dsm(int on);
/* Projections of the second-generation forward stack on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
 * 3xTF32, accumulator in TMEM; operands in SWIZZLE_128B shared tiles): 1 = when the tiles fit (default;
 * GATRES_RES2_TC presets it), 0 = the mma.sync form.  Other values only query.  Returns the previous setting. */
int gatres_set_resident_tc(int on);
/* Cluster-barrier flavour of those kernels: 0 = every thread arrives with .release (gpu-scope fence per barrier),
 * 1 = one releasing warp, 2 = CTA-scope fence + CTA barrier + relaxed arrival (default: the exchange is shared memory
 * only; rationale and measurements in csrc/resident2.cu).  GATRES_RES2_BARRIER presets it.  Other values only query.
 * Returns the previous setting. */
int gatres_set_resident_barrier(int flavour);

/*
 * Kernel-selection knob for the projections and their data gradients: 0 = fp32 FFMA kernels
 * everywhere; 1 (default) = tensor cores (tcgen05.mma kind::tf32, 3xTF32 error-compensated,
 * accumulator in TMEM) for launches of >= 32768 rows and the shapes that have such a kernel
 * (nc = 32); 2 = tensor cores whenever the shape allows.  GATRES_TC environment variable sets
 * the initial mode.  Negative = query only.  Returns the previous mode.
 */
int gatres_set_tensor_core(int mode);

/* --------------------------------------------------------------- operators */

/*
 * GATConv projection + attention scores (GATConv.forward step 1; call sites
 * GraphModels.py:464-465):  h = x W^T  ([M,K] x [H*C,K]^T -> [M,H*C]),
 * s_src[m,h] = <h[m,h,:], att_src[h,:]>, s_dst likewise.
 * Supported (K, H, C): C in {32,64,128}, H in {1,2}, K in {C, 2C} as the model uses.
 */
int gatres_linear_att_fwd(const float* x, const float* W, const float* att_src, const float* att_dst,
                          float* h, float* s_src, float* s_dst,
                          int64_t M, int32_t K, int32_t H, int32_t C, void* stream);

/*
 * Fused GAT aggregation (GATConv.forward steps 2-6): per target row, LeakyReLU(0.2)
 * logits over in-edges, segment softmax, alpha-weighted sum of neighbour rows,
 * + bias, optional ReLU (the `.relu()` at GraphModels.py:464).
 * m/l ([M,H], may be NULL) receive the softmax row max and exp-sum for the
 * recompute-based backward.  concat=0 requires H=1 (mean over one head).
 * E1 = number of CSR entries (rowptr[N], known to the host from gatres_csr_build's
 * info[1]); when E1 > 0 and one snapshot's [N, H*C] slab fits in shared memory the
 * TMA-staged snapshot-tile kernel is used, otherwise (or with E1 = 0) the gather kernel.
 */
int gatres_gat_agg_fwd(const int32_t* rowptr, const int32_t* col,
                       const float* h, const float* s_src, const float* s_dst, const float* bias,
                       float* out, float* m, float* l,
                       int64_t B, int32_t N, int32_t E1, int32_t H, int32_t C, int32_t relu, void* stream);

/*
 * Backward of gatres_gat_agg_fwd w.r.t. h, s (folded into dh), att_src, att_dst,
 * bias — two gather passes, no atomics, softmax recomputed from (m,l)
 * (SURVEY.md §A.4).  g = dL/d(out) AFTER any ReLU mask has been applied.
 *   pass 1 (per target, in-edge CSR):  rec = {s_dst, m, 1/l, D}, ds_dst
 *   pass 2 (per source, out-edge CSR): dh [M,H*C]
 *   rec [M,H,4] and ds_dst [M,H] are caller-allocated scratch.
 * Parameter-gradient partial sums go to `partial` rows [slots][P] (P = row
 * stride in floats, a multiple of 4, >= the parameter count) at offsets
 * off_att_src / off_att_dst / off_bias (floats, multiples of 4); every one of
 * the `slots` CTAs writes its row, gatres_reduce_partials sums them.
 */
int gatres_gat_agg_bwd(const int32_t* rowptr, const int32_t* col,
                       const int32_t* rowptr_t, const int32_t* col_t,
                       const float* g, const float* h, const float* s_src, const float* s_dst,
                       const float* m, const float* l, const float* att_src, const float* att_dst,
                       float* rec, float* ds_dst, float* dh,
                       float* partial, int64_t P, int32_t slots,
                       int64_t off_att_src, int64_t off_att_dst, int64_t off_bias,
                       int64_t B, int32_t N, int32_t E1, int32_t H, int32_t C, void* stream);

/*
 * SimpleConv(aggr="mean")(z) + x0, ReLU  (GraphModels.py:466-467).  Uses the
 * in-edge CSR minus each row's trailing self-loop; isolated nodes get 0 + x0.
 */
int gatres_mean_res_fwd(const int32_t* rowptr, const int32_t* col,
                        const float* z, const float* x0, float* out,
                        int64_t B, int32_t N, int32_t C, void* stream);

/*
 * Backward of gatres_mean_res_fwd: with gm = g_out * (out > 0),
 *   dz[j] = sum_{i : j->i} gm[i] / max(indeg(i),1)   (out-edge CSR minus self-loop)
 *   dres  = gm                                         (gradient of the +x0 branch)
 * `out` may be NULL when the caller already applied the ReLU mask to g_out
 * (then gm = g_out and dres is not written).
 */
int gatres_mean_res_bwd(const int32_t* rowptr, const int32_t* rowptr_t, const int32_t* col_t,
                        const float* g_out, const float* out, float* dz, float* dres,
                        int64_t B, int32_t N, int32_t C, void* stream);
/*
 * The form the model's backward uses (ReLU mask already applied, no residual output: out = dres = NULL above) with
 * the E1 hint of the aggregation kernels: for batches >= the tile threshold of graphs whose gradient slab [N, C] fits
 * shared memory twice, the snapshot's slab is staged by TMA (double-buffered), the out-edge CSR and the per-edge
 * weights 1 / max(indeg, 1) sit in shared memory and rows are handed out in ascending out-degree; otherwise (or with
 * E1 = 0) the gather kernel of gatres_mean_res_bwd runs.  Same results.
 */
int gatres_mean_res_bwd_e1(const int32_t* rowptr, const int32_t* rowptr_t, const int32_t* col_t, int32_t E1,
                           const float* g_masked, float* dz, int64_t B, int32_t N, int32_t C, void* stream);

/*
 * Backward of the projection: dx = dh W (+ add, if non-NULL) (* (relu_ref > 0), if non-NULL),
 * and dW partial sums (dh^T x) into `partial` at off_W.
 */
int gatres_linear_bwd(const float* dh, const float* x, const float* W,
                      const float* add, const float* relu_ref, float* dx,
                      float* partial, int64_t P, int32_t slots, int64_t off_W,
                      int64_t M, int32_t K, int32_t H, int32_t C, void* stream);

/* lin0 = Linear(1,nc) (GraphModels.py:487): out[m,c] = x[m]*w[c] + b[c]. */
int gatres_encoder_fwd(const float* x, const float* w, const float* b, float* out,
                       int64_t M, int32_t nc, void* stream);
int gatres_encoder_bwd(const float* g, const float* x, float* partial, int64_t P, int32_t slots,
                       int64_t off_w, int64_t off_b, int64_t M, int32_t nc, void* stream);
/* lin1 = Linear(nc,1) (GraphModels.py:492): out[m] = <x[m,:], w> + b. `poison`
 * (device int32, may be NULL): if nonzero the output is NaN (a failed topology check). */
int gatres_decoder_fwd(const float* x, const float* w, const float* b, float* out,
                       const int32_t* poison, int64_t M, int32_t nc, void* stream);
/* dx[m,c] = g_out[m] w[c], times (x[m,c] > 0) when mask_relu (x is a ReLU output). */
int gatres_decoder_bwd(const float* g_out, const float* x, const float* w, float* dx,
                       float* partial, int64_t P, int32_t slots, int64_t off_w, int64_t off_b,
                       int64_t M, int32_t nc, int32_t mask_relu, void* stream);

/* grads[p] = sum_s partial[s][p] for p in [p_begin, p_end). Deterministic. */
int gatres_reduce_partials(const float* partial, int64_t P, int32_t slots,
                           int64_t p_begin, int64_t p_end, float* grads, void* stream);

/* ------------------------------------------------------------ whole model */

typedef struct gatres_model_desc {
  int32_t num_blocks;            /* GATResMeanConv(num_blocks, nc), GraphModels.py:472 */
  int32_t nc;
  int32_t N;                     /* template nodes */
  int32_t slots;                 /* > 0: deterministic gradient reduction over `slots` partial rows; 0: atomics */
  int32_t E1;                    /* CSR entries (rowptr[N] = edges + N self loops) */
  int32_t reserved;
  int64_t B;                     /* snapshots in this batch */
  const int32_t* rowptr;         /* in-edge CSR incl. self loops (gatres_csr_build) */
  const int32_t* col;
  const int32_t* rowptr_t;       /* out-edge CSR incl. self loops */
  const int32_t* col_t;
  const int32_t* poison;         /* optional device flag: nonzero -> NaN output */
  /* Optional locality plan of the template (all NULL / 0 = none; the snapshot-resident kernels of small batches then
   * exchange rows through L2 instead of distributed shared memory).  perm[r] = original node id of locality row r
   * (a permutation of 0..N-1 that puts neighbours close together, e.g. recursive bisection of the network); p_* = the
   * same two CSR structures as above in locality numbering (row r = node perm[r], entries = locality ids of the
   * neighbours IN THE SAME ORDER as the original row, self-loop last); p_ecap[q] = largest number of CSR entries
   * owned by one CTA when the rows are cut into 2^q equal slices of ceil(N / 2^q) rows (q = 0..3), taken over both
   * CSRs.  Model inputs / outputs / gradients keep the original row order; only the saved-activation buffer between
   * gatres_forward(training) and gatres_backward is in locality order when the plan is used. */
  const int32_t* perm;
  const int32_t* p_rowptr;
  const int32_t* p_col;
  const int32_t* p_rowptr_t;
  const int32_t* p_col_t;
  int32_t p_ecap[4];
} gatres_model_desc;

/* sizeof(gatres_model_desc) as this library was built (a binding checks its own struct against it). */
size_t gatres_model_desc_bytes(void);
int64_t gatres_param_count(int32_t num_blocks, int32_t nc);
/* floats of activation storage forward(training) hands to backward */
int64_t gatres_saved_floats(const gatres_model_desc* d);
/* floats of scratch for forward (training=0) or forward+backward (training=1) */
int64_t gatres_scratch_floats(const gatres_model_desc* d, int32_t training);

/* GATResMeanConv.forward (GraphModels.py:486-494). x [M] -> out [M]; saved may be NULL (inference). */
int gatres_forward(const gatres_model_desc* d, const float* params, const float* x,
                   float* out, float* saved, float* scratch, void* stream);
/* Backward of gatres_forward for d_out [M]: grads [param_count] (overwritten, all params).
 * partial: [slots][align4(param_count)] floats of scratch for the two-stage reductions. */
int gatres_backward(const gatres_model_desc* d, const float* params, const float* x,
                    const float* saved, const float* d_out, float* partial, float* grads,
                    float* scratch, void* stream);

/*
 * The same backward in pieces, for overlapping the data-parallel gradient all-reduce with the rest of
 * the backward pass: blocks k_hi .. k_lo (descending).  The range that starts at num_blocks-1 also runs
 * the decoder backward (and zeroes `grads` in atomic mode); the range that ends at 0 also runs the
 * encoder backward (and the final reduction in deterministic mode).  Ranges must be issued in
 * descending order on one stream.  In atomic mode the gradients of blocks [k_lo, k_hi] (and lin1 / lin0
 * for the first / last range) are final when the call's kernels complete:
 * grads[gatres_param_offset_of_block(k_lo) .. gatres_param_offset_of_block(k_hi + 1)).
 */
int gatres_backward_range(const gatres_model_desc* d, const float* params, const float* x,
                          const float* saved, const float* d_out, float* partial, float* grads,
                          float* scratch, int32_t k_hi, int32_t k_lo, void* stream);
/* Offset (floats) of block k's parameters in the flat layout; k >= num_blocks -> lin1.weight, k < 0 -> 0. */
int64_t gatres_param_offset_of_block(int32_t num_blocks, int32_t nc, int32_t k);

/* ------------------------------------------------------- caller-side fusions */

/*
 * Masked-MSE (train.py:177-183): loss = mean_{mask}( (out-y)^2 ), d_out = 2 (out-y) mask / count.
 * mask: uint8 [M]; count = number of masked nodes (host-known: B*int(N*rate)).
 * loss_out: device float[1] (overwritten); partial_loss: device float[blocks_used] scratch (>= 1024 floats).
 */
int gatres_masked_mse(const float* out, const float* y, const uint8_t* mask, int64_t M, int64_t count,
                      float* d_out, float* loss_out, float* partial_loss, void* stream);
/* x_masked[m] = mask[m] ? 0 : x[m]  (train.py:174) */
int gatres_apply_mask(const float* x, const uint8_t* mask, float* x_masked, int64_t M, void* stream);

/*
 * Per-snapshot exact-count random mask on the device, replacing the host-side
 * generate_batch_mask(num_nodes, mask_rate, required_idx=[]) of utils/auxil.py:143-182 (called per batch at
 * train.py:171-172): for each of the B snapshots exactly `count` (= int(N * mask_rate)) of its N nodes get
 * mask = 1, uniformly at random without replacement.  Node (b, i) draws the 32-bit key
 * gatres_mask_key(seed, step, b*N + i); the `count` smallest keys of a snapshot are selected (ties by node
 * index), so the result is a pure function of (seed, step) - reproducible, replayable from a CUDA graph.
 * step_dev: optional device int32[1] added to `step` at run time (pass the Adam step counter so that a
 * captured graph draws a fresh mask every replay); NULL = use `step` alone.
 * required: optional device uint8[N] flags of template nodes that must be masked in every snapshot
 * (`required_idx`, the sensor nodes of evaluation.py:288-291); they take key 0 (hashed keys are >= 1), so they
 * are always selected while count >= their number; NULL = none.  mask: uint8 [B*N] (overwritten).
 */
int gatres_generate_mask(uint64_t seed, uint64_t step, const int32_t* step_dev, const uint8_t* required,
                         int64_t B, int32_t N, int32_t count, uint8_t* mask, void* stream);
/* The key function above, on the host (splitmix64 of the counter; for tests and host-side replication). */
uint32_t gatres_mask_key(uint64_t seed, uint64_t step, uint64_t row);

/*
 * The seven metrics of get_metric_fn_collection (utils/auxil.py:185-203) as train.py:177-198 and
 * evaluation.py:326-338 apply them: over the entries with mask != 0 (mask NULL = all M entries) of the DESCALED
 * prediction / target, descaled = v * scale + shift (znorm: scale = std, shift = mean; minmax: scale = max - min,
 * shift = min; none: 1, 0 - utils/auxil.py:42-64).
 * metrics_out: device float[8] = { rel. error (|t| > 0.01 only), accuracy(|e| <= t * threshold), correlation
 * (clamped to [-1, 1]), r2 = corr^2, MAE, RMSE, NSE, number of entries }.  Two passes with fp64 accumulators.
 * scratch: gatres_metrics_scratch_doubles() doubles of device memory.
 */
int64_t gatres_metrics_scratch_doubles(void);
int gatres_masked_metrics(const float* out, const float* y, const uint8_t* mask, int64_t M, float scale,
                          float shift, float threshold, double* scratch, float* metrics_out, void* stream);

/*
 * torch.optim.Adam step (train.py:348: lr 5e-4, weight_decay 6e-6 as L2 added to the
 * gradient) on the flat buffers.  step_count: device int32[1], incremented on device
 * so the call can be replayed from a CUDA graph.  grad_scale multiplies grads first
 * (1/world_size after an all-reduce sum).
 */
int gatres_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                     int32_t* step_count, int64_t P, float lr, float beta1, float beta2,
                     float eps, float weight_decay, float grad_scale, void* stream);

/*
 * Data-parallel variant: gradient all-reduce FUSED into the Adam step over NVLink peer memory (no reference
 * counterpart: the reference is single-process, train.py:306-324; contract = SURVEY 8e, "N-rank step == 1-rank
 * step on the concatenated batch").  peer_grads[r] / peer_flags[r] (HOST arrays of `world` device pointers) are
 * rank r's flat gradient buffer (16-byte aligned) and uint32[world] flag array, mapped into this process (CUDA
 * symmetric memory / IPC).  The kernel announces "rank's backward of this step is done" in every peer's flag array,
 * waits for all peers, then reads every rank's gradients through the peer pointers, sums them in rank order (all
 * replicas compute bit-identical updates) and applies Adam with grad_scale (1/world).  epoch: device int32[1],
 * incremented here (never reset by the caller, unlike step_count).  The caller alternates between two gradient
 * buffers per step, so this one barrier also protects a buffer from being re-zeroed while a peer still reads it.
 * world == 1 degenerates to gatres_adam_step.
 */
int gatres_adam_step_peer(float* params, const float* const* peer_grads, uint32_t* const* peer_flags,
                          int32_t rank, int32_t world, float* exp_avg, float* exp_avg_sq, int32_t* step_count,
                          int32_t* epoch, int64_t P, float lr, float beta1, float beta2, float eps,
                          float weight_decay, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GATRES_B200_H_ */
