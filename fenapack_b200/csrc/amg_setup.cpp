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

// power iteration on D^-1 A_dd (A_dd = diagonal block: columns owned by this rank)
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
        for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
          if (A.col[k] < n) t += A.val[k] * v[A.col[k]];
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

// strength graph of the diagonal block as CSR (indices + |a_ij|); ghost columns (>= n) are ignored,
// so aggregates never cross rank boundaries
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
      if (j != i && j < n && a != 0.0 && std::fabs(a) >= theta * std::sqrt(d[i] * d[j])) ++cnt;
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
      if (j != i && j < n && a != 0.0 && std::fabs(a) >= theta * std::sqrt(d[i] * d[j])) {
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
                                 int64_t nagg, double omega, double trunc, HostCsr &P) {
  const int64_t n = A.nrows;
  std::vector<double> tval(nagg, 0.0);
  {
    std::vector<int64_t> cnt(nagg, 0);
    for (int64_t i = 0; i < n; ++i)
      if (agg[i] >= 0) cnt[agg[i]]++;
    for (int64_t a = 0; a < nagg; ++a) tval[a] = 1.0 / std::sqrt((double)cnt[a]);
  }
  // T has one (empty) row per ghost column too: the prolongator smoothing is block local
  HostCsr T;
  T.nrows = A.ncols;
  T.ncols = nagg;
  T.rowptr.resize(A.ncols + 1);
  T.rowptr[0] = 0;
  for (int64_t i = 0; i < A.ncols; ++i) T.rowptr[i + 1] = T.rowptr[i] + ((i < n && agg[i] >= 0) ? 1 : 0);
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
  // truncation: entries below trunc * max|row| are dropped and the row is rescaled to its
  // former row sum (constants stay in the range of P).  Keeps the Galerkin structure while
  // cutting the coarse stencils of P2 operators in 3D by ~4x.
  if (trunc > 0.0) {
    std::vector<int32_t> rp(n + 1, 0);
    int64_t o = 0;
    for (int64_t i = 0; i < n; ++i) {
      double rmax = 0.0, rs0 = 0.0, rs1 = 0.0;
      for (int32_t k = P.rowptr[i]; k < P.rowptr[i + 1]; ++k) {
        rmax = std::max(rmax, std::fabs(P.val[k]));
        rs0 += P.val[k];
      }
      const int64_t start = o;
      for (int32_t k = P.rowptr[i]; k < P.rowptr[i + 1]; ++k)
        if (std::fabs(P.val[k]) >= trunc * rmax) {
          P.col[o] = P.col[k];
          P.val[o] = P.val[k];
          rs1 += P.val[k];
          ++o;
        }
      const double scale = rs1 != 0.0 ? rs0 / rs1 : 1.0;
      for (int64_t k = start; k < o; ++k) P.val[k] = scale * P.val[k];
      rp[i + 1] = (int32_t)o;
    }
    P.rowptr.swap(rp);
    P.col.resize(o);
    P.val.resize(o);
  }
}

// Sparsify a coarse operator in place: off-diagonal entries below
// drop*sqrt(|a_ii||a_jj|) are removed and added to the diagonal of their row.
// `A` has local column numbering [owned | ghost]; d_ghost holds |a_jj| of the ghost
// columns; `gcol` (optional) is the same pattern with global column ids and is
// compacted in lock step.
static void filter_lumped(HostCsr &A, double drop, const std::vector<double> &d_ghost, std::vector<int32_t> *gcol) {
  if (drop <= 0.0) return;
  const int64_t n = A.nrows;
  std::vector<double> d(A.ncols, 0.0);
  for (int64_t i = 0; i < n; ++i)
    for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
      if (A.col[k] == i) d[i] = std::fabs(A.val[k]);
  for (size_t g = 0; g < d_ghost.size(); ++g) d[n + g] = d_ghost[g];
  std::vector<int32_t> rp(n + 1, 0);
  int64_t o = 0;
  for (int64_t i = 0; i < n; ++i) {
    double lump = 0.0;
    int64_t diag_pos = -1;
    for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
      const int32_t j = A.col[k];
      const double a = A.val[k];
      if (j != i && std::fabs(a) < drop * std::sqrt(d[i] * d[j])) {
        lump += a;
        continue;
      }
      if (j == i) diag_pos = o;
      A.col[o] = j;          // compaction in place: o <= k always
      A.val[o] = a;
      if (gcol) (*gcol)[o] = (*gcol)[k];
      ++o;
    }
    if (diag_pos >= 0) A.val[diag_pos] += lump;
    rp[i + 1] = (int32_t)o;
  }
  A.rowptr.swap(rp);
  A.col.resize(o);
  A.val.resize(o);
  if (gcol) gcol->resize(o);
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

void host_spgemm(const HostCsr &A, const HostCsr &B, HostCsr &C) { spgemm(A, B, C); }
void host_transpose(const HostCsr &A, HostCsr &T) { transpose(A, T); }
void host_dense_inverse(const HostCsr &A, std::vector<double> &inv) { dense_inverse(A, inv); }

// helpers from dist.cu
std::vector<int64_t> comm_ranges(Ctx &c, int64_t n_local);
double comm_allreduce(Ctx &c, double v, bool max_op);
std::vector<double> comm_allgather_padded(Ctx &c, const double *v, int64_t count, int64_t maxcount);
std::shared_ptr<HaloPlan> build_halo(Ctx &c, HostCsr &h, const std::vector<int64_t> &begins,
                                     std::vector<int64_t> *ghost_global_out);
std::vector<double> halo_exchange_host(Ctx &c, HaloPlan &plan, const std::vector<double> &x_own);

// Rows of P for the ghost dofs of A (their owners computed them): returns P extended to
// n_own + n_ghost rows with GLOBAL coarse column ids.
static void extend_prolongator(Ctx &c, const HostCsr &P, int64_t coarse_begin, HaloPlan *plan, int64_t nghost,
                               HostCsr &Pext) {
  const int64_t n = P.nrows;
  Pext.nrows = n + nghost;
  Pext.ncols = 0;   // global ids; set by the caller
  Pext.rowptr.assign(n + nghost + 1, 0);
  for (int64_t i = 0; i < n; ++i) Pext.rowptr[i + 1] = P.rowptr[i + 1];
  int64_t W = 0;
  for (int64_t i = 0; i < n; ++i) W = std::max<int64_t>(W, P.rowptr[i + 1] - P.rowptr[i]);
  W = (int64_t)comm_allreduce(c, (double)W, true);
  std::vector<std::vector<double>> gc, gv;
  std::vector<double> glen;
  if (plan && nghost > 0) {
    std::vector<double> tmp(n);
    for (int64_t i = 0; i < n; ++i) tmp[i] = (double)(P.rowptr[i + 1] - P.rowptr[i]);
    glen = halo_exchange_host(c, *plan, tmp);
  } else if (plan) {
    std::vector<double> tmp(n, 0.0);
    halo_exchange_host(c, *plan, tmp);      // collective: every rank takes part
  }
  for (int64_t k = 0; k < W; ++k) {
    std::vector<double> cc(n, -1.0), vv(n, 0.0);
    for (int64_t i = 0; i < n; ++i)
      if (P.rowptr[i] + k < P.rowptr[i + 1]) {
        cc[i] = (double)(coarse_begin + P.col[P.rowptr[i] + k]);
        vv[i] = P.val[P.rowptr[i] + k];
      }
    if (plan) {
      gc.push_back(halo_exchange_host(c, *plan, cc));
      gv.push_back(halo_exchange_host(c, *plan, vv));
    }
  }
  for (int64_t g = 0; g < nghost; ++g) Pext.rowptr[n + g + 1] = Pext.rowptr[n + g] + (int32_t)glen[g];
  Pext.col.resize(Pext.rowptr[n + nghost]);
  Pext.val.resize(Pext.rowptr[n + nghost]);
  for (int64_t i = 0; i < n; ++i)
    for (int32_t k = P.rowptr[i]; k < P.rowptr[i + 1]; ++k) {
      Pext.col[k] = (int32_t)(coarse_begin + P.col[k]);
      Pext.val[k] = P.val[k];
    }
  for (int64_t g = 0; g < nghost; ++g)
    for (int32_t k = 0; k < (int32_t)glen[g]; ++k) {
      Pext.col[Pext.rowptr[n + g] + k] = (int32_t)gc[k][g];
      Pext.val[Pext.rowptr[n + g] + k] = gv[k][g];
    }
}

// Set-up of the hierarchy of one operator.  `A0` holds this rank's rows with GLOBAL
// column ids; `begins` are the ownership offsets of all ranks.  Aggregation and
// prolongator smoothing are local to the rank (aggregates never cross a rank
// boundary, P and R are block diagonal over ranks); the Galerkin product couples
// neighbouring ranks through the ghost rows of P.  With one rank this is the plain
// serial algorithm.
void amg_build_host(Ctx &c, const HostCsr &A0, std::vector<int64_t> begins, const AmgParams &p, HostHierarchy &H,
                    int level0) {
  H.levels.clear();
  H.coarse_inv.clear();
  H.tail.reset();
  const int me = c.rank, R = c.nranks;
  HostCsr Ag = A0;                 // current level, global column ids
  while (true) {
    H.levels.emplace_back();
    HostLevel &lvl = H.levels.back();
    lvl.A = Ag;
    std::vector<int64_t> ghosts;
    lvl.halo = build_halo(c, lvl.A, begins, &ghosts);     // relabels lvl.A to [owned | ghost]
    lvl.n_own = begins[me + 1] - begins[me];
    lvl.begins = begins;
    const HostCsr &A = lvl.A;
    csr_diag_inv(A, lvl.dinv);
    lvl.rho = comm_allreduce(c, estimate_rho(A, lvl.dinv), true);
    const int64_t n_global = begins[R];
    if (n_global <= p.coarse_size || (int)H.levels.size() + level0 >= p.max_levels) break;
    std::vector<int32_t> sp, sc, agg;
    std::vector<double> sv;
    strength(A, p.theta * std::pow(0.5, (double)(H.levels.size() - 1 + level0)), sp, sc, sv);
    const int64_t nagg = aggregate_greedy(A.nrows, sp, sc, sv, agg);
    std::vector<int64_t> cbegins = comm_ranges(c, nagg);
    if (cbegins[R] == 0 || cbegins[R] >= n_global) break;
    smoothed_prolongator(A, lvl.dinv, agg, nagg, p.omega_scale / lvl.rho, p.p_trunc, lvl.P);
    transpose(lvl.P, lvl.R);
    // Galerkin product with the ghost rows of P fetched from their owners
    HostCsr Pext;
    extend_prolongator(c, lvl.P, cbegins[me], lvl.halo.get(), (int64_t)ghosts.size(), Pext);
    // compact numbering of the coarse columns that occur: [my aggregates | foreign ones, sorted]
    std::vector<int64_t> foreign;
    for (int64_t k = 0; k < Pext.nnz(); ++k) {
      const int64_t g = Pext.col[k];
      if (g < cbegins[me] || g >= cbegins[me + 1]) foreign.push_back(g);
    }
    std::sort(foreign.begin(), foreign.end());
    foreign.erase(std::unique(foreign.begin(), foreign.end()), foreign.end());
    for (int64_t k = 0; k < Pext.nnz(); ++k) {
      const int64_t g = Pext.col[k];
      if (g >= cbegins[me] && g < cbegins[me + 1]) Pext.col[k] = (int32_t)(g - cbegins[me]);
      else Pext.col[k] = (int32_t)(nagg + (std::lower_bound(foreign.begin(), foreign.end(), g) - foreign.begin()));
    }
    Pext.ncols = nagg + (int64_t)foreign.size();
    HostCsr AP, Ac;
    spgemm(A, Pext, AP);
    spgemm(lvl.R, AP, Ac);
    // back to global coarse ids, rows sorted by global column
    std::vector<int32_t> gcol(Ac.col.size());
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < Ac.nrows; ++i) {
      std::vector<std::pair<int32_t, double>> row;
      for (int32_t k = Ac.rowptr[i]; k < Ac.rowptr[i + 1]; ++k) {
        const int32_t l = Ac.col[k];
        const int64_t g = l < nagg ? cbegins[me] + l : foreign[l - nagg];
        row.push_back({(int32_t)g, Ac.val[k]});
      }
      std::sort(row.begin(), row.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
      for (int32_t k = Ac.rowptr[i]; k < Ac.rowptr[i + 1]; ++k) {
        gcol[k] = row[k - Ac.rowptr[i]].first;
        Ac.val[k] = row[k - Ac.rowptr[i]].second;
      }
    }
    // sparsify: needs |a_jj| of the ghost columns
    HostCsr Aloc = Ac;
    Aloc.col = gcol;
    Aloc.ncols = cbegins[R];
    std::vector<int64_t> cghosts;
    std::shared_ptr<HaloPlan> cplan = build_halo(c, Aloc, cbegins, &cghosts);
    std::vector<double> dg;
    if (cplan) {
      std::vector<double> down(Aloc.nrows, 0.0);
      for (int64_t i = 0; i < Aloc.nrows; ++i)
        for (int32_t k = Aloc.rowptr[i]; k < Aloc.rowptr[i + 1]; ++k)
          if (Aloc.col[k] == i) down[i] = std::fabs(Aloc.val[k]);
      dg = halo_exchange_host(c, *cplan, down);
    }
    filter_lumped(Aloc, p.coarse_drop, dg, &gcol);
    if (R > 1) {
      // frozen-P Galerkin refresh (amg_refresh.cu) on several ranks: keep P with the ghost rows, columns
      // translated from the compact numbering [my aggregates | foreign] to the next level's local
      // numbering [owned | ghost]; foreign columns that never meet a local row are dropped
      HostLevel &cur = H.levels.back();
      cur.Pext = Pext;
      HostCsr &Q = cur.Pext;
      Q.ncols = nagg + (int64_t)cghosts.size();
      std::vector<int32_t> rp(Q.nrows + 1, 0);
      int64_t o = 0;
      for (int64_t i = 0; i < Q.nrows; ++i) {
        for (int32_t k = Pext.rowptr[i]; k < Pext.rowptr[i + 1]; ++k) {
          const int32_t l = Pext.col[k];
          int32_t loc = l;
          if (l >= nagg) {
            const int64_t g = foreign[l - nagg];
            const auto it = std::lower_bound(cghosts.begin(), cghosts.end(), g);
            loc = (it != cghosts.end() && *it == g) ? (int32_t)(nagg + (it - cghosts.begin())) : -1;
          }
          if (loc < 0) continue;
          Q.col[o] = loc;
          Q.val[o] = Pext.val[k];
          ++o;
        }
        rp[i + 1] = (int32_t)o;
      }
      Q.rowptr.swap(rp);
      Q.col.resize(o);
      Q.val.resize(o);
    }
    Ag = std::move(Aloc);
    Ag.col = std::move(gcol);
    Ag.ncols = cbegins[R];
    begins = cbegins;
    if (R > 1 && begins[R] <= p.replicate_size) {
      // gather the level on every rank and continue serially (identical on all ranks)
      const int64_t nloc = Ag.nrows, nnzloc = Ag.nnz();
      const int64_t maxrows = (int64_t)comm_allreduce(c, (double)nloc, true);
      const int64_t maxnnz = (int64_t)comm_allreduce(c, (double)nnzloc, true);
      std::vector<double> len(nloc), cold(nnzloc);
      for (int64_t i = 0; i < nloc; ++i) len[i] = (double)(Ag.rowptr[i + 1] - Ag.rowptr[i]);
      for (int64_t k = 0; k < nnzloc; ++k) cold[k] = (double)Ag.col[k];
      std::vector<double> all_len = comm_allgather_padded(c, len.data(), nloc, std::max<int64_t>(maxrows, 1));
      std::vector<double> all_col = comm_allgather_padded(c, cold.data(), nnzloc, std::max<int64_t>(maxnnz, 1));
      std::vector<double> all_val = comm_allgather_padded(c, Ag.val.data(), nnzloc, std::max<int64_t>(maxnnz, 1));
      HostCsr Af;
      Af.nrows = Af.ncols = begins[R];
      Af.rowptr.assign(begins[R] + 1, 0);
      for (int q = 0; q < R; ++q)
        for (int64_t i = 0; i < begins[q + 1] - begins[q]; ++i)
          Af.rowptr[begins[q] + i + 1] = (int32_t)all_len[(size_t)q * std::max<int64_t>(maxrows, 1) + i];
      for (int64_t i = 0; i < begins[R]; ++i) Af.rowptr[i + 1] += Af.rowptr[i];
      Af.col.resize(Af.rowptr[begins[R]]);
      Af.val.resize(Af.rowptr[begins[R]]);
      for (int q = 0; q < R; ++q) {
        const int64_t o = Af.rowptr[begins[q]], cnt = Af.rowptr[begins[q + 1]] - o;
        for (int64_t k = 0; k < cnt; ++k) {
          Af.col[o + k] = (int32_t)all_col[(size_t)q * std::max<int64_t>(maxnnz, 1) + k];
          Af.val[o + k] = all_val[(size_t)q * std::max<int64_t>(maxnnz, 1) + k];
        }
      }
      H.tail = std::make_shared<HostHierarchy>();
      H.tail_begins = begins;
      const int saved_rank = c.rank, saved_n = c.nranks;
      c.rank = 0;
      c.nranks = 1;                       // the helpers below short-circuit every collective
      try {
        amg_build_host(c, Af, {0, begins[R]}, p, *H.tail, level0 + (int)H.levels.size());
      } catch (...) {
        c.rank = saved_rank;
        c.nranks = saved_n;
        throw;
      }
      c.rank = saved_rank;
      c.nranks = saved_n;
      return;
    }
  }
  // coarsest level: every rank inverts the (small) global matrix and keeps its own rows,
  // columns laid out as the padded all-gather of the right-hand side delivers them
  H.coarse_gcol = Ag.col;
  H.coarse_begins = begins;
  amg_coarse_inverse_host(c, H);
}

void amg_coarse_inverse_host(Ctx &c, HostHierarchy &H) {
  const int me = c.rank, R = c.nranks;
  const std::vector<int64_t> &begins = H.coarse_begins;
  HostLevel &lvl = H.levels.back();
  const HostCsr &A = lvl.A;              // values current; global columns in H.coarse_gcol
  const int64_t n = lvl.n_own, ng = begins[R];
  int64_t maxloc = 0;
  for (int q = 0; q < R; ++q) maxloc = std::max(maxloc, begins[q + 1] - begins[q]);
  FNP_REQUIRE(ng <= 8192, FNP_ERR_NUMERIC, "AMG coarsening stalled: coarsest level has " + std::to_string(ng) + " rows");
  std::vector<double> rows((size_t)n * ng, 0.0);
  for (int64_t i = 0; i < n; ++i)
    for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) rows[(size_t)i * ng + H.coarse_gcol[k]] = A.val[k];
  std::vector<double> all = comm_allgather_padded(c, rows.data(), n * ng, maxloc * ng);
  HostCsr dense;      // reuse dense_inverse through a CSR view of the dense matrix
  dense.nrows = dense.ncols = ng;
  dense.rowptr.resize(ng + 1);
  dense.col.resize((size_t)ng * ng);
  dense.val.resize((size_t)ng * ng);
  for (int64_t gi = 0; gi <= ng; ++gi) dense.rowptr[gi] = (int32_t)(gi * ng);
  for (int q = 0; q < R; ++q)
    for (int64_t i = 0; i < begins[q + 1] - begins[q]; ++i)
      for (int64_t j = 0; j < ng; ++j) {
        const int64_t gi = begins[q] + i;
        dense.col[(size_t)gi * ng + j] = (int32_t)j;
        dense.val[(size_t)gi * ng + j] = all[(size_t)q * maxloc * ng + (size_t)i * ng + j];
      }
  std::vector<double> inv;
  dense_inverse(dense, inv);
  H.coarse_cols = R * maxloc;
  H.coarse_maxloc = maxloc;
  H.coarse_inv.assign((size_t)n * H.coarse_cols, 0.0);
  for (int64_t i = 0; i < n; ++i)
    for (int q = 0; q < R; ++q)
      for (int64_t t = 0; t < begins[q + 1] - begins[q]; ++t)
        H.coarse_inv[(size_t)i * H.coarse_cols + q * maxloc + t] = inv[(size_t)(begins[me] + i) * ng + begins[q] + t];
}

}  // namespace fnp
