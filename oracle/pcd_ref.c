/* C / OpenMP restatement of the PCD-preconditioned FGMRES path -- TEST
 * INFRASTRUCTURE and the CPU arm timed by bench.py (cpu_baseline, --impl
 * reference).  PARITY UNPINNED for the PETSc-owned algorithms (see oracle/__init__.py; the
 * Schur applies are pinned through the numpy restatement): the reference only wires
 * PETSc / hypre together; this file restates the same algorithm chain as
 * oracle/petsc_algos.py + oracle/amg.py, function for function, so that it can be
 * (a) checked against the numpy restatement and (b) timed on all host cores,
 * which is what `mpirun -n <cores>` PETSc would use.
 *
 *   ref_spmv            Mat.mult                    fenapack/preconditioners.py:131,164
 *   ref_cheb_jacobi     KSPCHEBYSHEV + PCJACOBI     demo_navier-stokes-pcd.py:161-165
 *   ref_vcycle          one AMG V-cycle             demo_navier-stokes-pcd.py:153-160 (BoomerAMG slot)
 *   ref_schur_apply     PCDPC_BRM1/2.apply          fenapack/preconditioners.py:124-135,158-169
 *   ref_pc_apply        PCFIELDSPLIT SCHUR/UPPER    fenapack/field_split.py:54-57
 *   ref_fgmres          KSPGMRES right PC, CGS      fenapack/field_split.py:52-53
 *   ref_aggregate_greedy  the aggregation loop of oracle/amg.py (speed only)
 *
 * Build: make -C oracle   ->  oracle/_ref/libpcd_ref.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int n, m;
  const int32_t *rp, *ci;
  const double *v;
} csr_t;

void ref_spmv(int n, const int32_t *rp, const int32_t *ci, const double *v, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int k = rp[i]; k < rp[i + 1]; ++k) s += v[k] * x[ci[k]];
    y[i] = s;
  }
}

static void spmv(const csr_t *A, const double *x, double *y) { ref_spmv(A->n, A->rp, A->ci, A->v, x, y); }

/* y = a*A x + b*z */
static void spmv_axpby(const csr_t *A, const double *x, double a, double b, const double *z, double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < A->n; ++i) {
    double s = 0.0;
    for (int k = A->rp[i]; k < A->rp[i + 1]; ++k) s += A->v[k] * x[A->ci[k]];
    y[i] = a * s + b * z[i];
  }
}

/* KSPCHEBYSHEV + PCJACOBI, zero initial guess, `steps` Jacobi applications:
 * out = (add ? add : 0) + scale * p_last.  Same recurrence as oracle/petsc_algos.py. */
void ref_cheb_jacobi(int n, const int32_t *rp, const int32_t *ci, const double *v, const double *dinv,
                     const double *b, double emin, double emax, int steps, double scale, const double *add,
                     double *out, double *w0, double *w1) {
  const double s = 2.0 / (emax + emin), alpha = 1.0 - s * emin, mu = 1.0 / alpha, omegaprod = 2.0 / alpha;
  if (steps == 1) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) out[i] = (add ? add[i] : 0.0) + scale * s * dinv[i] * b[i];
    return;
  }
  double *p0 = NULL, *p1 = w0, *spare = w1;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) p1[i] = s * dinv[i] * b[i];
  double c0 = 1.0, c1 = mu;
  for (int pass = 1; pass < steps; ++pass) {
    const double c2 = 2.0 * mu * c1 - c0, omega = omegaprod * c1 / c2;
    const int last = pass == steps - 1;
    double *p2 = last ? out : (p0 ? p0 : spare);
    const double sc = last ? scale : 1.0;
    const double k0 = sc * (1.0 - omega), k1 = sc * omega, k2 = sc * omega * s;
    const double *ad = last ? add : NULL;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      double t = 0.0;
      for (int k = rp[i]; k < rp[i + 1]; ++k) t += v[k] * p1[ci[k]];
      double r = k1 * p1[i] + k2 * dinv[i] * (b[i] - t);
      if (p0) r += k0 * p0[i];
      if (ad) r += ad[i];
      p2[i] = r;
    }
    p0 = p1;
    p1 = p2;
    c0 = c1;
    c1 = c2;
  }
}

/* ---- AMG hierarchy ------------------------------------------------------ */
typedef struct {
  csr_t A, P, R;
  const double *dinv;
  double rho;
  double *x, *b, *r, *w0, *w1;
} level_t;

typedef struct {
  int nlev, smooth_steps;
  double eig_ratio;
  level_t *lv;
  const double *coarse_inv;
  int coarse_n;
} hier_t;

hier_t *ref_hier_create(int nlev, int smooth_steps, double eig_ratio) {
  hier_t *h = (hier_t *)calloc(1, sizeof(hier_t));
  h->nlev = nlev;
  h->smooth_steps = smooth_steps;
  h->eig_ratio = eig_ratio;
  h->lv = (level_t *)calloc((size_t)nlev, sizeof(level_t));
  return h;
}

static csr_t mk(int n, int m, const int32_t *rp, const int32_t *ci, const double *v) {
  csr_t a = {n, m, rp, ci, v};
  return a;
}

void ref_hier_set_level(hier_t *h, int l, int n, int nc, const int32_t *arp, const int32_t *aci, const double *av,
                        const double *dinv, double rho, const int32_t *prp, const int32_t *pci, const double *pv,
                        const int32_t *rrp, const int32_t *rci, const double *rv) {
  level_t *L = &h->lv[l];
  L->A = mk(n, n, arp, aci, av);
  L->dinv = dinv;
  L->rho = rho;
  if (prp) {
    L->P = mk(n, nc, prp, pci, pv);
    L->R = mk(nc, n, rrp, rci, rv);
  }
  L->x = (double *)malloc(sizeof(double) * (size_t)n);
  L->b = (double *)malloc(sizeof(double) * (size_t)n);
  L->r = (double *)malloc(sizeof(double) * (size_t)n);
  L->w0 = (double *)malloc(sizeof(double) * (size_t)n);
  L->w1 = (double *)malloc(sizeof(double) * (size_t)n);
}

void ref_hier_set_coarse(hier_t *h, int n, const double *inv) {
  h->coarse_n = n;
  h->coarse_inv = inv;
}

void ref_hier_free(hier_t *h) {
  if (!h) return;
  for (int l = 0; l < h->nlev; ++l) {
    free(h->lv[l].x); free(h->lv[l].b); free(h->lv[l].r); free(h->lv[l].w0); free(h->lv[l].w1);
  }
  free(h->lv);
  free(h);
}

static void vcycle_level(hier_t *h, int l, const double *b, double *x) {
  level_t *L = &h->lv[l];
  if (l == h->nlev - 1) {
    const int n = h->coarse_n;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += h->coarse_inv[(size_t)i * n + k] * b[k];
      x[i] = s;
    }
    return;
  }
  const double emax = L->rho, emin = L->rho / h->eig_ratio;
  level_t *C = &h->lv[l + 1];
  ref_cheb_jacobi(L->A.n, L->A.rp, L->A.ci, L->A.v, L->dinv, b, emin, emax, h->smooth_steps, 1.0, NULL, x, L->w0, L->w1);
  spmv_axpby(&L->A, x, -1.0, 1.0, b, L->r);
  spmv(&L->R, L->r, C->b);
  vcycle_level(h, l + 1, C->b, C->x);
  spmv_axpby(&L->P, C->x, 1.0, 1.0, x, x);
  spmv_axpby(&L->A, x, -1.0, 1.0, b, L->r);
  ref_cheb_jacobi(L->A.n, L->A.rp, L->A.ci, L->A.v, L->dinv, L->r, emin, emax, h->smooth_steps, 1.0, x, x, L->w0, L->w1);
}

void ref_vcycle(hier_t *h, const double *b, double *x) { vcycle_level(h, 0, b, x); }

/* ---- the preconditioner ---------------------------------------------------- */
typedef struct {
  int n_u, n_p, variant; /* 1 = BRM1, 2 = BRM2 */
  csr_t A00, A01, A10, Ap, Mp, Kp, P00;
  const double *mp_dinv;
  const int32_t *bc_idx;
  const double *bc_val;
  int nbc;
  hier_t *hu, *hp;
  int cheb_steps, ap_its, u_its;
  double emin, emax;
  double *pw[6], *uw[3];
} pcd_t;

pcd_t *ref_pcd_create(int n_u, int n_p, int variant) {
  pcd_t *p = (pcd_t *)calloc(1, sizeof(pcd_t));
  p->n_u = n_u; p->n_p = n_p; p->variant = variant;
  p->cheb_steps = 5; p->ap_its = 2; p->u_its = 1; p->emin = 0.5; p->emax = 2.0;
  for (int i = 0; i < 6; ++i) p->pw[i] = (double *)malloc(sizeof(double) * (size_t)n_p);
  for (int i = 0; i < 3; ++i) p->uw[i] = (double *)malloc(sizeof(double) * (size_t)n_u);
  return p;
}

void ref_pcd_free(pcd_t *p) {
  if (!p) return;
  for (int i = 0; i < 6; ++i) free(p->pw[i]);
  for (int i = 0; i < 3; ++i) free(p->uw[i]);
  free(p);
}

/* which: 0 A00, 1 A01, 2 A10, 3 Ap, 4 Mp, 5 Kp, 6 P00 (same ids as the C ABI) */
void ref_pcd_set_matrix(pcd_t *p, int which, int n, int m, const int32_t *rp, const int32_t *ci, const double *v) {
  csr_t a = mk(n, m, rp, ci, v);
  switch (which) {
    case 0: p->A00 = a; break;
    case 1: p->A01 = a; break;
    case 2: p->A10 = a; break;
    case 3: p->Ap = a; break;
    case 4: p->Mp = a; break;
    case 5: p->Kp = a; break;
    default: p->P00 = a; break;
  }
}

void ref_pcd_set_inner(pcd_t *p, const double *mp_dinv, double emin, double emax, int cheb_steps, int ap_its,
                       int u_its, hier_t *hu, hier_t *hp, const int32_t *bc_idx, const double *bc_val, int nbc) {
  p->mp_dinv = mp_dinv; p->emin = emin; p->emax = emax; p->cheb_steps = cheb_steps;
  p->ap_its = ap_its; p->u_its = u_its; p->hu = hu; p->hp = hp;
  p->bc_idx = bc_idx; p->bc_val = bc_val; p->nbc = nbc;
}

/* KSPRICHARDSON + one V-cycle as PC: x1 = B b ; x_{k+1} = x_k + B (b - A x_k) */
static void richardson_amg(const csr_t *A, hier_t *h, const double *b, double *x, int its, double *w0, double *w1) {
  ref_vcycle(h, b, x);
  for (int k = 1; k < its; ++k) {
    spmv_axpby(A, x, -1.0, 1.0, b, w0);
    ref_vcycle(h, w0, w1);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < A->n; ++i) x[i] += w1[i];
  }
}

static void mp_solve(pcd_t *p, const double *b, double scale, double *x) {
  ref_cheb_jacobi(p->n_p, p->Mp.rp, p->Mp.ci, p->Mp.v, p->mp_dinv, b, p->emin, p->emax, p->cheb_steps, scale, NULL,
                  x, p->pw[3], p->pw[4]);
}

void ref_schur_apply(pcd_t *p, const double *x, double *y) {
  const int n = p->n_p;
  double *z = p->pw[0];
  if (p->variant == 1) {
    memcpy(z, x, sizeof(double) * (size_t)n);
    for (int i = 0; i < p->nbc; ++i) z[p->bc_idx[i]] = p->bc_val[i];
    richardson_amg(&p->Ap, p->hp, z, y, p->ap_its, p->pw[1], p->pw[2]);
    spmv_axpby(&p->Kp, y, 1.0, 1.0, x, z);
    mp_solve(p, z, -1.0, y);
  } else {
    double *z0 = p->pw[5];
    mp_solve(p, x, 1.0, y);
    spmv(&p->Kp, y, z);
    for (int i = 0; i < p->nbc; ++i) z[p->bc_idx[i]] = p->bc_val[i];
    richardson_amg(&p->Ap, p->hp, z, z0, p->ap_its, p->pw[1], p->pw[2]);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) y[i] = -(y[i] + z0[i]);
  }
}

void ref_pc_apply(pcd_t *p, const double *xu, const double *xp, double *yu, double *yp) {
  ref_schur_apply(p, xp, yp);
  spmv_axpby(&p->A01, yp, -1.0, 1.0, xu, p->uw[0]);
  const csr_t *P00 = p->P00.rp ? &p->P00 : &p->A00;
  richardson_amg(P00, p->hu, p->uw[0], yu, p->u_its, p->uw[1], p->uw[2]);
}

void ref_system_matvec(pcd_t *p, const double *x, double *y) {
  const double *xu = x, *xp = x + p->n_u;
  spmv(&p->A00, xu, y);
  spmv_axpby(&p->A01, xp, 1.0, 1.0, y, y);
  spmv(&p->A10, xu, y + p->n_u);
}

static double dotp(int64_t n, const double *a, const double *b) {
  double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* Right-preconditioned FGMRES(restart), classical Gram-Schmidt, zero initial
 * guess, stop on |g_{j+1}| <= max(rtol*||b||, atol) or after max_it iterations.
 * hist (capacity max_it+1) receives the residual-norm estimates. Returns the
 * iteration count; *napply counts preconditioner applications. */
int ref_fgmres(pcd_t *p, const double *b, double *x, double rtol, double atol, int restart, int max_it, double *hist,
               int *napply) {
  const int64_t n = (int64_t)p->n_u + p->n_p;
  const int m = restart < max_it ? restart : (max_it > 0 ? max_it : 1); /* never allocate more basis than can be used */
  double *V = (double *)malloc(sizeof(double) * (size_t)n * (size_t)(m + 1));
  double *Z = (double *)malloc(sizeof(double) * (size_t)n * (size_t)m);
  double *w = (double *)malloc(sizeof(double) * (size_t)n);
  double *H = (double *)calloc((size_t)(m + 1) * m, sizeof(double));
  double *cs = (double *)calloc((size_t)m, sizeof(double)), *sn = (double *)calloc((size_t)m, sizeof(double));
  double *g = (double *)calloc((size_t)m + 1, sizeof(double)), *yv = (double *)calloc((size_t)m, sizeof(double));
#define HH(i, j) H[(size_t)(j) * (m + 1) + (i)]
  memset(x, 0, sizeof(double) * (size_t)n);
  const double bnorm = sqrt(dotp(n, b, b));
  const double tol = fmax(rtol * bnorm, atol);
  int its = 0, nap = 0;
  hist[0] = bnorm;
  double beta = bnorm;
  if (bnorm > tol) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) V[i] = b[i] / beta;
    int done = 0;
    while (!done) {
      memset(g, 0, sizeof(double) * (size_t)(m + 1));
      g[0] = beta;
      int jdone = 0, converged = 0;
      for (int j = 0; j < m; ++j) {
        double *vj = V + (size_t)j * n, *zj = Z + (size_t)j * n, *vn = V + (size_t)(j + 1) * n;
        ref_pc_apply(p, vj, vj + p->n_u, zj, zj + p->n_u);
        ++nap;
        ref_system_matvec(p, zj, w);
        for (int i = 0; i <= j; ++i) HH(i, j) = dotp(n, V + (size_t)i * n, w);
        for (int i = 0; i <= j; ++i) {
          const double h = HH(i, j);
          const double *vi = V + (size_t)i * n;
#pragma omp parallel for schedule(static)
          for (int64_t k = 0; k < n; ++k) w[k] -= h * vi[k];
        }
        const double hn = sqrt(dotp(n, w, w));
        HH(j + 1, j) = hn;
        if (hn != 0.0) {
#pragma omp parallel for schedule(static)
          for (int64_t k = 0; k < n; ++k) vn[k] = w[k] / hn;
        }
        for (int i = 0; i < j; ++i) {
          const double a = HH(i, j), bb = HH(i + 1, j);
          HH(i, j) = cs[i] * a + sn[i] * bb;
          HH(i + 1, j) = -sn[i] * a + cs[i] * bb;
        }
        const double a = HH(j, j), bb = HH(j + 1, j), rho = hypot(a, bb);
        if (rho == 0.0) { cs[j] = 1.0; sn[j] = 0.0; } else { cs[j] = a / rho; sn[j] = bb / rho; }
        HH(j, j) = rho;
        HH(j + 1, j) = 0.0;
        g[j + 1] = -sn[j] * g[j];
        g[j] = cs[j] * g[j];
        ++its;
        jdone = j + 1;
        hist[its] = fabs(g[j + 1]);
        if (hist[its] <= tol || its >= max_it) { converged = hist[its] <= tol; break; }
      }
      for (int i = jdone - 1; i >= 0; --i) {
        double s = g[i];
        for (int k = i + 1; k < jdone; ++k) s -= HH(i, k) * yv[k];
        yv[i] = s / HH(i, i);
      }
      for (int i = 0; i < jdone; ++i) {
        const double c = yv[i];
        const double *zi = Z + (size_t)i * n;
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < n; ++k) x[k] += c * zi[k];
      }
      if (converged || its >= max_it) break;
      ref_system_matvec(p, x, w);
#pragma omp parallel for schedule(static)
      for (int64_t k = 0; k < n; ++k) w[k] = b[k] - w[k];
      beta = sqrt(dotp(n, w, w));
      if (beta <= tol) break;
#pragma omp parallel for schedule(static)
      for (int64_t k = 0; k < n; ++k) V[k] = w[k] / beta;
    }
  }
#undef HH
  free(V); free(Z); free(w); free(H); free(cs); free(sn); free(g); free(yv);
  *napply = nap;
  return its;
}

/* ---- aggregation loop of oracle/amg.py:aggregate_greedy -------------------- */
int ref_aggregate_greedy(int n, const int32_t *sp, const int32_t *sc, const double *sv, int64_t *agg) {
  int nagg = 0;
  for (int i = 0; i < n; ++i) agg[i] = -1;
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -1 || sp[i + 1] == sp[i]) continue;
    int ok = 1;
    for (int k = sp[i]; k < sp[i + 1]; ++k)
      if (agg[sc[k]] != -1) { ok = 0; break; }
    if (!ok) continue;
    agg[i] = nagg;
    for (int k = sp[i]; k < sp[i + 1]; ++k) agg[sc[k]] = nagg;
    ++nagg;
  }
  int64_t *agg1 = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
  memcpy(agg1, agg, sizeof(int64_t) * (size_t)n);
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -1 || sp[i + 1] == sp[i]) continue;
    int64_t best = -1;
    double bestv = -1.0;
    for (int k = sp[i]; k < sp[i + 1]; ++k)
      if (agg1[sc[k]] != -1 && sv[k] > bestv) { best = agg1[sc[k]]; bestv = sv[k]; }
    if (best != -1) agg[i] = best;
  }
  free(agg1);
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -1 || sp[i + 1] == sp[i]) continue;
    agg[i] = nagg;
    for (int k = sp[i]; k < sp[i + 1]; ++k) {
      const int j = sc[k];
      if (agg[j] == -1 && sp[j + 1] != sp[j]) agg[j] = nagg;
    }
    ++nagg;
  }
  return nagg;
}

int ref_num_threads(void) {
  int n = 1;
#pragma omp parallel
  {
#pragma omp single
    n =
#ifdef _OPENMP
        omp_get_num_threads();
#else
        1;
#endif
  }
  return n;
}
