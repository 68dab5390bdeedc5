// Flat parameter layout and saved-activation layout of the GATRes stack, shared by the
// host orchestration (model.cu) and the snapshot-resident kernels (resident.cu).
// The parameter order is the one documented in include/gatres_b200.h.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define GATRES_HD __host__ __device__
#else
#define GATRES_HD
#endif

namespace gatres {

GATRES_HD static inline int64_t a4(int64_t n) { return (n + 3) & ~(int64_t)3; }

struct ParamLayout {
  int64_t nc, nb;
  GATRES_HD explicit ParamLayout(int32_t num_blocks, int32_t nc_) : nc(nc_), nb(num_blocks) {}
  GATRES_HD int64_t lin0_w() const { return 0; }
  GATRES_HD int64_t lin0_b() const { return nc; }
  GATRES_HD int64_t block_size() const { return 4 * nc * nc + 9 * nc; }
  GATRES_HD int64_t block(int64_t k) const { return 2 * nc + k * block_size(); }
  GATRES_HD int64_t c1_W(int64_t k) const { return block(k); }
  GATRES_HD int64_t c1_as(int64_t k) const { return block(k) + 2 * nc * nc; }
  GATRES_HD int64_t c1_ad(int64_t k) const { return c1_as(k) + 2 * nc; }
  GATRES_HD int64_t c1_b(int64_t k) const { return c1_ad(k) + 2 * nc; }
  GATRES_HD int64_t c2_W(int64_t k) const { return c1_b(k) + 2 * nc; }
  GATRES_HD int64_t c2_as(int64_t k) const { return c2_W(k) + 2 * nc * nc; }
  GATRES_HD int64_t c2_ad(int64_t k) const { return c2_as(k) + nc; }
  GATRES_HD int64_t c2_b(int64_t k) const { return c2_ad(k) + nc; }
  GATRES_HD int64_t lin1_w() const { return block(nb); }
  GATRES_HD int64_t lin1_b() const { return lin1_w() + nc; }
  GATRES_HD int64_t count() const { return lin1_b() + 1; }
};

// activations forward(training) keeps for backward
struct SavedLayout {
  int64_t M, nc;
  GATRES_HD SavedLayout(int64_t M_, int64_t nc_) : M(M_), nc(nc_) {}
  GATRES_HD int64_t x_enc() const { return 0; }
  GATRES_HD int64_t block_size() const { return 6 * M * nc + 4 * a4(2 * M) + 4 * a4(M); }
  GATRES_HD int64_t block(int64_t k) const { return M * nc + k * block_size(); }
  GATRES_HD int64_t h1(int64_t k) const { return block(k); }
  GATRES_HD int64_t ss1(int64_t k) const { return h1(k) + 2 * M * nc; }
  GATRES_HD int64_t sd1(int64_t k) const { return ss1(k) + a4(2 * M); }
  GATRES_HD int64_t m1(int64_t k) const { return sd1(k) + a4(2 * M); }
  GATRES_HD int64_t l1(int64_t k) const { return m1(k) + a4(2 * M); }
  GATRES_HD int64_t y1(int64_t k) const { return l1(k) + a4(2 * M); }
  GATRES_HD int64_t h2(int64_t k) const { return y1(k) + 2 * M * nc; }
  GATRES_HD int64_t ss2(int64_t k) const { return h2(k) + M * nc; }
  GATRES_HD int64_t sd2(int64_t k) const { return ss2(k) + a4(M); }
  GATRES_HD int64_t m2(int64_t k) const { return sd2(k) + a4(M); }
  GATRES_HD int64_t l2(int64_t k) const { return m2(k) + a4(M); }
  GATRES_HD int64_t xout(int64_t k) const { return l2(k) + a4(M); }
  GATRES_HD int64_t total(int64_t nb) const { return block(nb); }
};

}  // namespace gatres
