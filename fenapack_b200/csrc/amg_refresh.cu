// Device-side numeric refresh of an AMG hierarchy with frozen prolongators.
//
// With P frozen, A_c = P^T A P is linear in the values of A: A_c.val = W * A.val with the plan
// matrix W of galerkin_plan.hpp.  W is built once per hierarchy on the host, re-indexed to the
// positions of the device value arrays (CSR order or SELL-padded order) and uploaded like any other
// operator, so that refreshing a level is one spmv_store on device-resident values -- deterministic,
// no atomics -- followed by a Jacobi-diagonal kernel.  The smoother bounds rho are kept (measured
// irrelevant in the oracle prototype: DESIGN.md section 8 item 4), the dense coarsest-level inverse
// is recomputed on the host from the few hundred rows it has.
//
// Option pc_amg_refresh galerkin (the default on single-rank contexts); covered on the GPU by
// tests/test_gpu_parity.py::test_device_side_galerkin_refresh and the refresh tests of
// tests/test_api_dropin.py, host plan logic by tests/test_host_logic.py.
#include <algorithm>

#include "fnp_internal.cuh"
#include "galerkin_plan.hpp"

namespace fnp {


double comm_allreduce(Ctx &c, double v, bool max_op);

static double *dev_values(DevCsr &A) { return A.sell ? A.sl_val.p : A.val.p; }
static int64_t dev_nvalues(const DevCsr &A) { return A.sell ? A.sell_entries : A.nnz; }
static int64_t dev_position(const DevCsr &A, int64_t k) { return A.sell ? (int64_t)A.sell_pos[(size_t)k] : k; }

__global__ void refresh_dinv_kernel(int64_t n, const int32_t *__restrict__ diagpos, const double *__restrict__ val,
                                    double *__restrict__ dinv, int bs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = diagpos[i] >= 0 ? val[diagpos[i]] : 0.0;
  const double r = d != 0.0 ? 1.0 / d : 0.0;
  for (int b = 0; b < bs; ++b) dinv[i * bs + b] = r;
}

static void build_plans(Ctx &c, DevHierarchy &H) {
  const size_t L = H.levels.size();
  H.refresh_W.clear();
  H.refresh_row0.clear();
  H.refresh_diag.clear();
  H.refresh_W.resize(L - 1);
  H.refresh_row0.resize(L - 1);
  H.refresh_diag.resize(L - 1);
  for (size_t l = 0; l + 1 < L; ++l) {
    const HostLevel &hf = H.host.levels[l], &hc = H.host.levels[l + 1];
    GalerkinPlan plan;
    try {
      // several ranks: P with the rows of the ghost dofs, columns in the coarse level's local numbering
      // (amg_setup.cpp); the plan then involves this rank's values only -- no communication at refresh
      build_galerkin_plan(hf.A, hf.Pext.nrows > 0 ? hf.Pext : hf.P, hf.R, hc.A, plan);
    } catch (const std::exception &e) {
      throw Error(FNP_ERR_STATE, e.what());
    }
    DevCsr &Af = H.levels[l].A(), &Ac = H.levels[l + 1].A();
    // W re-indexed: row = device position of the coarse entry, column = device position of the fine
    // entry.  Stored in row chunks of at most ~2^30 terms (32-bit entry offsets on the device): level 0
    // of the 128^3 cavity has 1.7e9 terms.
    const int64_t wrows = dev_nvalues(Ac), nq = hc.A.nnz();
    std::vector<int64_t> cnt((size_t)wrows + 1, 0);            // terms per W row, then offsets
    for (int64_t q = 0; q < nq; ++q) cnt[(size_t)dev_position(Ac, q) + 1] = plan.ptr[q + 1] - plan.ptr[q];
    for (int64_t r = 0; r < wrows; ++r) cnt[(size_t)r + 1] += cnt[(size_t)r];
    std::vector<int64_t> row_of_q((size_t)nq);
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < nq; ++q) row_of_q[(size_t)q] = dev_position(Ac, q);
    const int64_t LIMIT = c.refresh_chunk_terms;
    std::vector<int64_t> cuts{0};
    while (cuts.back() < wrows) {
      const int64_t r0 = cuts.back();
      // largest r1 with terms(r0, r1) <= LIMIT (at least one row)
      int64_t r1 = std::upper_bound(cnt.begin() + r0 + 1, cnt.end(), cnt[(size_t)r0] + LIMIT) - cnt.begin() - 1;
      r1 = std::max(r1, r0 + 1);
      cuts.push_back(std::min(r1, wrows));
    }
    const size_t nchunk = cuts.size() - 1;
    H.refresh_W[l].resize(nchunk);
    H.refresh_row0[l].assign(cuts.begin(), cuts.end() - 1);
    for (size_t ch = 0; ch < nchunk; ++ch) {
      const int64_t r0 = cuts[ch], r1 = cuts[ch + 1], t0 = cnt[(size_t)r0];
      HostCsr W;
      W.nrows = r1 - r0;
      W.ncols = dev_nvalues(Af);
      W.rowptr.resize((size_t)W.nrows + 1);
#pragma omp parallel for schedule(static)
      for (int64_t r = r0; r <= r1; ++r) W.rowptr[(size_t)(r - r0)] = (int32_t)(cnt[(size_t)r] - t0);
      const int64_t terms = cnt[(size_t)r1] - t0;
      FNP_REQUIRE(terms < (int64_t)INT32_MAX, FNP_ERR_ARG, "Galerkin refresh plan: one coarse entry with more than 2^31 terms");
      W.col.resize((size_t)terms);
      W.val.resize((size_t)terms);
#pragma omp parallel for schedule(static)
      for (int64_t q = 0; q < nq; ++q) {
        const int64_t r = row_of_q[(size_t)q];
        if (r < r0 || r >= r1) continue;
        int64_t dst = cnt[(size_t)r] - t0;
        for (int64_t t = plan.ptr[q]; t < plan.ptr[q + 1]; ++t, ++dst) {
          W.col[(size_t)dst] = (int32_t)dev_position(Af, plan.src[(size_t)t]);
          W.val[(size_t)dst] = plan.coef[(size_t)t];
        }
      }
      DevCsr &Wd = H.refresh_W[l][ch];
      csr_upload_pattern(c, Wd, W, "refresh/W" + std::to_string(l));
      csr_set_values(c, Wd, W, W.val.data(), false);
      Wd.sell_pos.clear();
      Wd.sell_pos.shrink_to_fit();                  // the plan is never refreshed itself
    }
    std::vector<int32_t> dp((size_t)hc.A.nrows, -1);
    for (int64_t i = 0; i < hc.A.nrows; ++i)
      for (int32_t k = hc.A.rowptr[i]; k < hc.A.rowptr[i + 1]; ++k)
        if (hc.A.col[k] == i) dp[(size_t)i] = (int32_t)dev_position(Ac, k);
    H.refresh_diag[l].upload(dp.data(), dp.size(), c.stream);
  }
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  H.refresh_built = true;
}

void amg_refresh_device(Ctx &c, DevHierarchy &H, int bs) {
  FNP_REQUIRE(H.built && !H.levels.empty(), FNP_ERR_STATE, "AMG hierarchy not built");
  FNP_REQUIRE(!H.tail, FNP_ERR_OPTION, "pc_amg_refresh galerkin does not cover a replicated coarse tail (pc_amg_replicate_size)");
  FNP_REQUIRE(H.params.coarse_drop == 0.0, FNP_ERR_OPTION, "pc_amg_refresh galerkin needs pc_amg_coarse_drop 0");
  const size_t L = H.levels.size();
  StageTimer t(c, "FENaPack: AMG Galerkin refresh");
  if (L > 1 && !H.refresh_built) {
    // host work without collectives: a failure on one rank must not leave the others waiting in the
    // gather of the coarsest level below
    std::string err;
    try {
      build_plans(c, H);
    } catch (const std::exception &e) {
      err = e.what();
    }
    const double failed = comm_allreduce(c, err.empty() ? 0.0 : 1.0, true);
    if (failed > 0.5) {
      H.refresh_W.clear();
      H.refresh_row0.clear();
      H.refresh_diag.clear();
      H.refresh_built = false;
      throw Error(FNP_ERR_STATE, err.empty() ? std::string("Galerkin refresh plan failed on another rank") : err);
    }
  }
  for (size_t l = 0; l + 1 < L; ++l) {
    DevCsr &Af = H.levels[l].A(), &Ac = H.levels[l + 1].A();
    for (size_t ch = 0; ch < H.refresh_W[l].size(); ++ch)
      spmv_store(c, H.refresh_W[l][ch], dev_values(Af), dev_values(Ac) + H.refresh_row0[l][ch]);
    const int64_t n = Ac.nrows;
    refresh_dinv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(n, H.refresh_diag[l].p, dev_values(Ac), Ac.dinv.p, bs);
    c.launches++;
    FNP_CUDA(cudaPeekAtLastError());
  }
  // host copies of the level values are fetched lazily (amg_sync_host: introspection, later
  // rebuilds); only the coarsest level comes back now, to be inverted again
  H.host_vals_stale = true;
  if (L > 1) {
    DevCsr &A = H.levels.back().A();
    HostCsr &h = H.host.levels.back().A;
    std::vector<double> buf((size_t)dev_nvalues(A));
    FNP_CUDA(cudaMemcpyAsync(buf.data(), dev_values(A), buf.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    FNP_CUDA(cudaStreamSynchronize(c.stream));
    for (int64_t k = 0; k < h.nnz(); ++k) h.val[(size_t)k] = buf[(size_t)dev_position(A, k)];
  }
  if (L > 1) {
    // collective on several ranks: the coarsest rows are gathered and inverted redundantly
    amg_coarse_inverse_host(c, H.host);
    H.coarse_inv.upload(H.host.coarse_inv.data(), H.host.coarse_inv.size(), c.stream);
    FNP_CUDA(cudaStreamSynchronize(c.stream));
  }
}

}  // namespace fnp

namespace fnp {

// Host mirror of the level operators (values and Jacobi diagonals) after device-side refreshes.
void amg_sync_host(Ctx &c, DevHierarchy &H) {
  if (!H.host_vals_stale) return;
  std::vector<double> buf;
  for (size_t l = 0; l < H.levels.size(); ++l) {
    DevCsr &A = H.levels[l].A();
    HostCsr &h = H.host.levels[l].A;
    buf.resize((size_t)dev_nvalues(A));
    if (!buf.empty())
      FNP_CUDA(cudaMemcpyAsync(buf.data(), dev_values(A), buf.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    FNP_CUDA(cudaStreamSynchronize(c.stream));
    h.val.resize((size_t)h.nnz());
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < h.nnz(); ++k) h.val[(size_t)k] = buf[(size_t)dev_position(A, k)];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < h.nrows; ++i) {
      double d = 0.0;
      for (int32_t k = h.rowptr[i]; k < h.rowptr[i + 1]; ++k)
        if (h.col[k] == i) d = h.val[k];
      H.host.levels[l].dinv[(size_t)i] = d != 0.0 ? 1.0 / d : 0.0;
    }
  }
  H.host_vals_stale = false;
}

}  // namespace fnp
