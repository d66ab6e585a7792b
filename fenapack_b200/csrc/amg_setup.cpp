// Host-side set-up of the smoothed-aggregation hierarchy (set-up time only; the
// V-cycle itself runs on the device, amg.cu).  Takes over the role of
// hypre BoomerAMG's set-up phase, which the reference triggers through
// ksp.setUp() (fenapack/field_split_backend.py:250-255, field_split.py:103-106).
//
// Algorithm (mirrored by oracle/amg.py so the two can be compared to rounding):
//   strength   |a_ij| >= theta_l sqrt(|a_ii||a_jj|), theta_l = theta / 2^l
//   aggregates greedy three-phase; rows without strong neighbours (Dirichlet
//              rows) stay out of every aggregate
//   tentative  T[i, agg(i)] = 1/sqrt(|agg|)
//   P          T - (omega_scale/rho) D^-1 A T,  rho = 1.1 * power-iteration(20)
//   A_c        P^T A P  (row-wise Gustavson products, OpenMP over rows)
//   coarsest   dense inverse by LU with partial pivoting
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "fnp_internal.cuh"

namespace fnp {

static void csr_diag_inv(const HostCsr &A, std::vector<double> &dinv) {
  dinv.assign(A.nrows, 0.0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < A.nrows; ++i) {
    double d = 0.0;
    for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
      if (A.col[k] == i) d = A.val[k];
    dinv[i] = d != 0.0 ? 1.0 / d : 0.0;
  }
}

static double estimate_rho(const HostCsr &A, const std::vector<double> &dinv, int steps = 20, double safety = 1.1) {
  const int64_t n = A.nrows;
  std::vector<double> v(n), w(n);
  double nrm = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    uint64_t h = ((uint64_t)i * 2654435761ull + 12345ull) & 0xFFFFFFFFull;
    v[i] = 0.5 + (double)h / 4294967296.0;
    nrm += v[i] * v[i];
  }
  nrm = std::sqrt(nrm);
  for (int64_t i = 0; i < n; ++i) v[i] /= nrm;
  double rho = 0.0;
  // fixed-size chunks summed in a fixed order: the estimate (and with it the whole
  // hierarchy) is bit-reproducible whatever the number of host threads
  const int64_t CH = 4096, nch = (n + CH - 1) / CH;
  std::vector<double> part(nch);
  for (int s = 0; s < steps; ++s) {
#pragma omp parallel for schedule(static)
    for (int64_t cidx = 0; cidx < nch; ++cidx) {
      double a = 0.0;
      const int64_t e = std::min(n, (cidx + 1) * CH);
      for (int64_t i = cidx * CH; i < e; ++i) {
        double t = 0.0;
        for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) t += A.val[k] * v[A.col[k]];
        t *= dinv[i];
        w[i] = t;
        a += t * t;
      }
      part[cidx] = a;
    }
    double acc = 0.0;
    for (int64_t cidx = 0; cidx < nch; ++cidx) acc += part[cidx];
    rho = std::sqrt(acc);
    if (rho == 0.0) return 1.0;
    const double inv = 1.0 / rho;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) v[i] = w[i] * inv;
  }
  return safety * rho;
}

// strength graph as CSR (indices + |a_ij|)
static void strength(const HostCsr &A, double theta, std::vector<int32_t> &sp, std::vector<int32_t> &sc,
                     std::vector<double> &sv) {
  const int64_t n = A.nrows;
  std::vector<double> d(n, 0.0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
      if (A.col[k] == i) d[i] = std::fabs(A.val[k]);
  sp.assign(n + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int cnt = 0;
    for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
      const int32_t j = A.col[k];
      const double a = A.val[k];
      if (j != i && a != 0.0 && std::fabs(a) >= theta * std::sqrt(d[i] * d[j])) ++cnt;
    }
    sp[i + 1] = cnt;
  }
  for (int64_t i = 0; i < n; ++i) sp[i + 1] += sp[i];
  sc.resize(sp[n]);
  sv.resize(sp[n]);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int32_t o = sp[i];
    for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
      const int32_t j = A.col[k];
      const double a = A.val[k];
      if (j != i && a != 0.0 && std::fabs(a) >= theta * std::sqrt(d[i] * d[j])) {
        sc[o] = j;
        sv[o] = std::fabs(a);
        ++o;
      }
    }
  }
}

static int64_t aggregate_greedy(int64_t n, const std::vector<int32_t> &sp, const std::vector<int32_t> &sc,
                                const std::vector<double> &sv, std::vector<int32_t> &agg) {
  agg.assign(n, -1);
  int32_t nagg = 0;
  // phase 1
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] != -1 || sp[i + 1] == sp[i]) continue;
    bool free_nb = true;
    for (int32_t k = sp[i]; k < sp[i + 1]; ++k)
      if (agg[sc[k]] != -1) { free_nb = false; break; }
    if (!free_nb) continue;
    agg[i] = nagg;
    for (int32_t k = sp[i]; k < sp[i + 1]; ++k) agg[sc[k]] = nagg;
    ++nagg;
  }
  // phase 2 (decisions based on the phase-1 state only)
  std::vector<int32_t> agg1(agg);
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] != -1 || sp[i + 1] == sp[i]) continue;
    int32_t best = -1;
    double bestv = -1.0;
    for (int32_t k = sp[i]; k < sp[i + 1]; ++k) {
      const int32_t j = sc[k];
      if (agg1[j] != -1 && sv[k] > bestv) { best = agg1[j]; bestv = sv[k]; }
    }
    if (best != -1) agg[i] = best;
  }
  // phase 3
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] != -1 || sp[i + 1] == sp[i]) continue;
    agg[i] = nagg;
    for (int32_t k = sp[i]; k < sp[i + 1]; ++k) {
      const int32_t j = sc[k];
      if (agg[j] == -1 && sp[j + 1] != sp[j]) agg[j] = nagg;
    }
    ++nagg;
  }
  return nagg;
}

// C = A * B (row-wise Gustavson, sorted output rows)
static void spgemm(const HostCsr &A, const HostCsr &B, HostCsr &C) {
  FNP_REQUIRE(A.ncols == B.nrows, FNP_ERR_ARG, "spgemm: dimension mismatch");
  const int64_t n = A.nrows, m = B.ncols;
  C.nrows = n;
  C.ncols = m;
  C.rowptr.assign(n + 1, 0);
  // symbolic
#pragma omp parallel
  {
    std::vector<int32_t> mark(m, -1);
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; ++i) {
      int32_t cnt = 0;
      for (int32_t ka = A.rowptr[i]; ka < A.rowptr[i + 1]; ++ka) {
        const int32_t k = A.col[ka];
        for (int32_t kb = B.rowptr[k]; kb < B.rowptr[k + 1]; ++kb) {
          const int32_t j = B.col[kb];
          if (mark[j] != (int32_t)i) { mark[j] = (int32_t)i; ++cnt; }
        }
      }
      C.rowptr[i + 1] = cnt;
    }
  }
  int64_t total = 0;
  for (int64_t i = 0; i < n; ++i) {
    total += C.rowptr[i + 1];
    FNP_REQUIRE(total < (int64_t)INT32_MAX, FNP_ERR_ARG, "spgemm: product exceeds 2^31 non-zeros");
    C.rowptr[i + 1] = (int32_t)total;
  }
  C.col.resize(total);
  C.val.resize(total);
  // numeric
#pragma omp parallel
  {
    std::vector<int32_t> pos(m, -1);
    std::vector<std::pair<int32_t, double>> rowbuf;
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; ++i) {
      const int32_t beg = C.rowptr[i];
      int32_t cnt = 0;
      for (int32_t ka = A.rowptr[i]; ka < A.rowptr[i + 1]; ++ka) {
        const int32_t k = A.col[ka];
        const double a = A.val[ka];
        for (int32_t kb = B.rowptr[k]; kb < B.rowptr[k + 1]; ++kb) {
          const int32_t j = B.col[kb];
          if (pos[j] < beg) {
            pos[j] = beg + cnt;
            C.col[beg + cnt] = j;
            C.val[beg + cnt] = a * B.val[kb];
            ++cnt;
          } else {
            C.val[pos[j]] += a * B.val[kb];
          }
        }
      }
      // sort the row by column
      rowbuf.resize(cnt);
      for (int32_t t = 0; t < cnt; ++t) rowbuf[t] = {C.col[beg + t], C.val[beg + t]};
      std::sort(rowbuf.begin(), rowbuf.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
      for (int32_t t = 0; t < cnt; ++t) {
        C.col[beg + t] = rowbuf[t].first;
        C.val[beg + t] = rowbuf[t].second;
        pos[rowbuf[t].first] = -1;
      }
    }
  }
}

static void transpose(const HostCsr &A, HostCsr &T) {
  T.nrows = A.ncols;
  T.ncols = A.nrows;
  T.rowptr.assign(T.nrows + 1, 0);
  const int64_t nnz = A.nnz();
  for (int64_t k = 0; k < nnz; ++k) T.rowptr[A.col[k] + 1]++;
  for (int64_t i = 0; i < T.nrows; ++i) T.rowptr[i + 1] += T.rowptr[i];
  T.col.resize(nnz);
  T.val.resize(nnz);
  std::vector<int32_t> cur(T.rowptr.begin(), T.rowptr.end() - 1);
  for (int64_t i = 0; i < A.nrows; ++i)
    for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
      const int32_t o = cur[A.col[k]]++;
      T.col[o] = (int32_t)i;
      T.val[o] = A.val[k];
    }
}

// P = T - omega * D^-1 (A T), T given by agg / counts
static void smoothed_prolongator(const HostCsr &A, const std::vector<double> &dinv, const std::vector<int32_t> &agg,
                                 int64_t nagg, double omega, HostCsr &P) {
  const int64_t n = A.nrows;
  std::vector<double> tval(nagg, 0.0);
  {
    std::vector<int64_t> cnt(nagg, 0);
    for (int64_t i = 0; i < n; ++i)
      if (agg[i] >= 0) cnt[agg[i]]++;
    for (int64_t a = 0; a < nagg; ++a) tval[a] = 1.0 / std::sqrt((double)cnt[a]);
  }
  HostCsr T;
  T.nrows = n;
  T.ncols = nagg;
  T.rowptr.resize(n + 1);
  T.rowptr[0] = 0;
  for (int64_t i = 0; i < n; ++i) T.rowptr[i + 1] = T.rowptr[i] + (agg[i] >= 0 ? 1 : 0);
  T.col.resize(T.rowptr[n]);
  T.val.resize(T.rowptr[n]);
  for (int64_t i = 0; i < n; ++i)
    if (agg[i] >= 0) {
      T.col[T.rowptr[i]] = agg[i];
      T.val[T.rowptr[i]] = tval[agg[i]];
    }
  HostCsr AT;
  spgemm(A, T, AT);
  // merge: P_ij = T_ij - omega * (dinv_i * AT_ij); pattern = union
  P.nrows = n;
  P.ncols = nagg;
  P.rowptr.assign(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    int32_t cnt = AT.rowptr[i + 1] - AT.rowptr[i];
    if (agg[i] >= 0) {
      bool found = false;
      for (int32_t k = AT.rowptr[i]; k < AT.rowptr[i + 1]; ++k)
        if (AT.col[k] == agg[i]) { found = true; break; }
      if (!found) ++cnt;
    }
    P.rowptr[i + 1] = P.rowptr[i] + cnt;
  }
  P.col.resize(P.rowptr[n]);
  P.val.resize(P.rowptr[n]);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int32_t o = P.rowptr[i];
    const int32_t a = agg[i];
    bool placed = a < 0;
    for (int32_t k = AT.rowptr[i]; k < AT.rowptr[i + 1]; ++k) {
      const int32_t j = AT.col[k];
      const double s = omega * (dinv[i] * AT.val[k]);
      if (!placed && a < j) {
        P.col[o] = a; P.val[o] = tval[a]; ++o;
        placed = true;
      }
      if (j == a) {
        P.col[o] = j; P.val[o] = tval[a] - s; ++o;
        placed = true;
      } else {
        P.col[o] = j; P.val[o] = 0.0 - s; ++o;
      }
    }
    if (!placed) { P.col[o] = a; P.val[o] = tval[a]; ++o; }
  }
}

static void dense_inverse(const HostCsr &A, std::vector<double> &inv) {
  const int64_t n = A.nrows;
  std::vector<double> M((size_t)n * n, 0.0);
  for (int64_t i = 0; i < n; ++i)
    for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) M[(size_t)i * n + A.col[k]] = A.val[k];
  inv.assign((size_t)n * n, 0.0);
  for (int64_t i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int64_t c = 0; c < n; ++c) {
    int64_t piv = c;
    double best = std::fabs(M[(size_t)c * n + c]);
    for (int64_t r = c + 1; r < n; ++r)
      if (std::fabs(M[(size_t)r * n + c]) > best) { best = std::fabs(M[(size_t)r * n + c]); piv = r; }
    FNP_REQUIRE(best > 0.0, FNP_ERR_NUMERIC, "AMG coarsest-level matrix is singular");
    if (piv != c)
      for (int64_t j = 0; j < n; ++j) {
        std::swap(M[(size_t)c * n + j], M[(size_t)piv * n + j]);
        std::swap(inv[(size_t)c * n + j], inv[(size_t)piv * n + j]);
      }
    const double d = 1.0 / M[(size_t)c * n + c];
    for (int64_t j = 0; j < n; ++j) { M[(size_t)c * n + j] *= d; inv[(size_t)c * n + j] *= d; }
#pragma omp parallel for schedule(static) if (n > 256)
    for (int64_t r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = M[(size_t)r * n + c];
      if (f == 0.0) continue;
      for (int64_t j = 0; j < n; ++j) {
        M[(size_t)r * n + j] -= f * M[(size_t)c * n + j];
        inv[(size_t)r * n + j] -= f * inv[(size_t)c * n + j];
      }
    }
  }
}

void amg_build_host(const HostCsr &A0, const AmgParams &p, HostHierarchy &H) {
  H.levels.clear();
  H.coarse_inv.clear();
  H.levels.emplace_back();
  H.levels.back().A = A0;
  while (true) {
    HostLevel &lvl = H.levels.back();
    const HostCsr &A = lvl.A;
    csr_diag_inv(A, lvl.dinv);
    lvl.rho = estimate_rho(A, lvl.dinv);
    if (A.nrows <= p.coarse_size || (int)H.levels.size() >= p.max_levels) break;
    std::vector<int32_t> sp, sc, agg;
    std::vector<double> sv;
    strength(A, p.theta * std::pow(0.5, (double)(H.levels.size() - 1)), sp, sc, sv);
    const int64_t nagg = aggregate_greedy(A.nrows, sp, sc, sv, agg);
    if (nagg == 0 || nagg >= A.nrows) break;
    smoothed_prolongator(A, lvl.dinv, agg, nagg, p.omega_scale / lvl.rho, lvl.P);
    transpose(lvl.P, lvl.R);
    HostCsr AP, Ac;
    spgemm(A, lvl.P, AP);
    spgemm(lvl.R, AP, Ac);
    H.levels.emplace_back();
    H.levels.back().A = std::move(Ac);
  }
  dense_inverse(H.levels.back().A, H.coarse_inv);
}

}  // namespace fnp
