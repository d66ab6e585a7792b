// Plan of the numeric Galerkin refresh with frozen prolongators (host logic, no CUDA dependency;
// compiled into libfenapack_cuda and, alone, into the CPU test harness tests/host/plan_harness.cpp).
//
// With P frozen the coarse operator A_c = P^T A P is linear in the values of A:
//     A_c.val[q] = sum_{t in [ptr[q], ptr[q+1])} coef[t] * A.val[src[t]]
// entry q = (I, J) of A_c collecting p_iI * a_ij * p_jJ over the fine entries e = (i, j).  The plan is
// the CSR matrix W = (ptr, src, coef) with one row per stored entry of A_c and one column per stored
// entry of A, so a value refresh of a level is ONE SpMV-class pass -- deterministic, no atomics.
// Prototype and measurements: oracle/amg.py:galerkin_plan, DESIGN.md section 8 item 4.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <utility>
#include <vector>

namespace fnp {

struct GalerkinPlan {
  std::vector<int64_t> ptr;    // size nnz(A_c) + 1
  std::vector<int32_t> src;    // entry index into A's CSR value array
  std::vector<double> coef;    // p_iI * p_jJ
  int64_t terms() const { return ptr.empty() ? 0 : ptr.back(); }
};

// A: n x n fine operator, P: n x nc prolongator, R = P^T (rows sorted), Ac: nc x nc with the
// structural pattern of R A P (rows sorted by column).  Csr needs rowptr / col / val vectors.
template <class Csr>
void build_galerkin_plan(const Csr &A, const Csr &P, const Csr &R, const Csr &Ac, GalerkinPlan &W) {
  const int64_t nc = (int64_t)Ac.rowptr.size() - 1;
  const int64_t nq = Ac.rowptr.empty() ? 0 : Ac.rowptr.back();
  W.ptr.assign(nq + 1, 0);
  // column-sorted view of every coarse row (on several ranks the rows of Ac are in the local numbering
  // [owned | ghost], which is not ascending along a row): scol = columns sorted, sq = entry index
  std::vector<int32_t> scol(Ac.col.begin(), Ac.col.begin() + nq);
  std::vector<int64_t> sq((size_t)nq);
#pragma omp parallel
  {
    std::vector<std::pair<int32_t, int64_t>> row;
#pragma omp for schedule(static)
    for (int64_t I = 0; I < nc; ++I) {
      const int64_t b = Ac.rowptr[I], e = Ac.rowptr[I + 1];
      bool sorted = true;
      for (int64_t k = b + 1; k < e; ++k)
        if (Ac.col[k - 1] > Ac.col[k]) { sorted = false; break; }
      if (sorted) {
        for (int64_t k = b; k < e; ++k) sq[(size_t)k] = k;
        continue;
      }
      row.clear();
      for (int64_t k = b; k < e; ++k) row.push_back({Ac.col[k], k});
      std::sort(row.begin(), row.end());
      for (int64_t k = b; k < e; ++k) {
        scol[(size_t)k] = row[(size_t)(k - b)].first;
        sq[(size_t)k] = row[(size_t)(k - b)].second;
      }
    }
  }
  auto find = [&](int64_t I, int32_t J) -> int64_t {
    const int32_t *b = scol.data() + Ac.rowptr[I], *e = scol.data() + Ac.rowptr[I + 1];
    const int32_t *it = std::lower_bound(b, e, J);
    return (it != e && *it == J) ? sq[(size_t)(it - scol.data())] : -1;
  };
  bool missing = false;
  // pass 1: terms per coarse entry (row I owns the entries [Ac.rowptr[I], Ac.rowptr[I+1]))
#pragma omp parallel for schedule(dynamic, 256) reduction(|| : missing)
  for (int64_t I = 0; I < nc; ++I) {
    for (int32_t kr = R.rowptr[I]; kr < R.rowptr[I + 1]; ++kr) {
      const int32_t i = R.col[kr];
      for (int32_t e = A.rowptr[i]; e < A.rowptr[i + 1]; ++e) {
        const int32_t j = A.col[e];
        for (int32_t kp = P.rowptr[j]; kp < P.rowptr[j + 1]; ++kp) {
          const int64_t q = find(I, P.col[kp]);
          if (q < 0) { missing = true; continue; }
          ++W.ptr[q + 1];
        }
      }
    }
  }
  if (missing) throw std::runtime_error("galerkin plan: the coarse pattern does not contain the structural product");
  for (int64_t q = 0; q < nq; ++q) W.ptr[q + 1] += W.ptr[q];
  W.src.resize((size_t)W.ptr[nq]);
  W.coef.resize((size_t)W.ptr[nq]);
  // pass 2: fill, same traversal order (fine row, fine entry, coarse column ascending)
#pragma omp parallel
  {
    std::vector<int64_t> cursor;
#pragma omp for schedule(dynamic, 256)
    for (int64_t I = 0; I < nc; ++I) {
      const int64_t q0 = Ac.rowptr[I], q1 = Ac.rowptr[I + 1];
      cursor.assign((size_t)(q1 - q0), 0);
      for (int32_t kr = R.rowptr[I]; kr < R.rowptr[I + 1]; ++kr) {
        const int32_t i = R.col[kr];
        const double pi = R.val[kr];
        for (int32_t e = A.rowptr[i]; e < A.rowptr[i + 1]; ++e) {
          const int32_t j = A.col[e];
          for (int32_t kp = P.rowptr[j]; kp < P.rowptr[j + 1]; ++kp) {
            const int64_t q = find(I, P.col[kp]);
            const int64_t t = W.ptr[q] + cursor[(size_t)(q - q0)]++;
            W.src[(size_t)t] = e;
            W.coef[(size_t)t] = pi * P.val[kp];
          }
        }
      }
    }
  }
}

}  // namespace fnp
