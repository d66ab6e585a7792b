// CPU harness of fenapack_b200/csrc/galerkin_plan.hpp (host logic of the library, no CUDA):
// built by tests/test_host_logic.py with g++ and called through ctypes.
#include <cstring>

#include "../../fenapack_b200/csrc/galerkin_plan.hpp"

namespace {
struct Csr {
  std::vector<int32_t> rowptr, col;
  std::vector<double> val;
};
Csr make(int64_t nrows, const int32_t *rp, const int32_t *ci, const double *va) {
  Csr m;
  m.rowptr.assign(rp, rp + nrows + 1);
  m.col.assign(ci, ci + rp[nrows]);
  m.val.assign(va, va + rp[nrows]);
  return m;
}
fnp::GalerkinPlan g_plan;
}  // namespace

extern "C" {
// returns the number of terms (or -1 when the coarse pattern misses a product entry); the plan is
// kept for plan_copy
int64_t plan_build(int64_t n, int64_t nc, const int32_t *a_rp, const int32_t *a_ci, const double *a_va, const int32_t *p_rp,
                   const int32_t *p_ci, const double *p_va, const int32_t *r_rp, const int32_t *r_ci, const double *r_va,
                   const int32_t *c_rp, const int32_t *c_ci, const double *c_va) {
  try {
    fnp::build_galerkin_plan(make(n, a_rp, a_ci, a_va), make(n, p_rp, p_ci, p_va), make(nc, r_rp, r_ci, r_va),
                             make(nc, c_rp, c_ci, c_va), g_plan);
  } catch (const std::exception &) {
    return -1;
  }
  return g_plan.terms();
}
void plan_copy(int64_t *ptr, int32_t *src, double *coef) {
  std::memcpy(ptr, g_plan.ptr.data(), g_plan.ptr.size() * sizeof(int64_t));
  std::memcpy(src, g_plan.src.data(), g_plan.src.size() * sizeof(int32_t));
  std::memcpy(coef, g_plan.coef.data(), g_plan.coef.size() * sizeof(double));
}
}
