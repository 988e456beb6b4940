/*
 * kt_oracle.c -- C/OpenMP port of the reference's Krylov solvers.  TEST / BASELINE INFRASTRUCTURE ONLY
 * (see oracle/__init__.py): used by tests (cross-check of the NumPy oracle), by bench.py's cpu_baseline leg and by
 * `bench.py --impl reference`.  Never linked into or called by the product.
 *
 * It keeps the reference's kernel granularity one to one -- every @ti.kernel of cgsolver.py / bicgsolver.py is one
 * `#pragma omp parallel for` loop over the outermost index (Taichi's CPU backend parallelises the outermost loop
 * over a thread pool in the same way), every reduce() is one parallel reduction returned to the (serial) driver,
 * including the identity copies update_phat / update_shat / copy(Ashat, t) the reference performs
 * (bicgsolver.py:104,111,114): 14 parallel loops and 5 reductions per BiCGSTAB iteration, 6 + 2 per CG iteration.
 *
 * Layout = the reference's: coef[i][j][k] (k fastest; 0 ap, 1 aw (i-1), 2 ae (i+1), 3 an (j+1), 4 as (j-1)),
 * vectors [i][j], all double.  A neighbour outside the array contributes 0.
 * Compile with -ffp-contract=off so element-wise results are bit-identical to the NumPy oracle.
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define IDX(i, j) ((size_t)(i) * NY + (j))

/* number of OpenMP threads of every later loop (1: strictly sequential sums in index order) */
#ifdef _OPENMP
#include <omp.h>
void ko_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
#else
void ko_set_threads(int n) { (void)n; }
#endif

/* cgsolver.py:49-54 */
double ko_reduce(const double* p, const double* q, long n) {
  double sum = 0.0;
#pragma omp parallel for reduction(+ : sum) schedule(static)
  for (long k = 0; k < n; ++k) sum += p[k] * q[k];
  return sum;
}

/* cgsolver.py:56-60 / bicgsolver.py:65-76 */
void ko_spmv(const double* coef, const double* x, double* y, int NX, int NY) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < NX; ++i) {
    for (int j = 0; j < NY; ++j) {
      const double* c = coef + IDX(i, j) * 5;
      const double xw = i > 0 ? x[IDX(i - 1, j)] : 0.0;
      const double xe = i < NX - 1 ? x[IDX(i + 1, j)] : 0.0;
      const double xn = j < NY - 1 ? x[IDX(i, j + 1)] : 0.0;
      const double xs = j > 0 ? x[IDX(i, j - 1)] : 0.0;
      double acc = c[0] * x[IDX(i, j)];
      acc = acc + c[1] * xw;
      acc = acc + c[2] * xe;
      acc = acc + c[3] * xn;
      acc = acc + c[4] * xs;
      y[IDX(i, j)] = acc;
    }
  }
}

static void k_copy(const double* src, double* dst, long n) {
#pragma omp parallel for schedule(static)
  for (long k = 0; k < n; ++k) dst[k] = src[k];
}

/* status codes = include/karman_b200.h KT_STATUS_* */
enum { CONVERGED = 0, MAX_ITER = 1, RHO_ZERO = 2, ALPHA_DIV_ZERO = 3, OMEGA_DIV_ZERO = 4 };

/* CGSolver.solve, cgsolver.py:83-114.  work = 3 planes (r, p, Ap). */
int ko_cg(const double* coef, const double* b, double* x, int NX, int NY, double eps, long max_iter, double* work,
          long* iters_out, double* rel_res) {
  const long n = (long)NX * NY;
  double *r = work, *p = work + n, *Ap = work + 2 * n;
  /* init: cgsolver.py:31-39 */
#pragma omp parallel for schedule(static)
  for (long k = 0; k < n; ++k) { r[k] = b[k]; Ap[k] = 0.0; p[k] = 0.0; x[k] = 0.0; }
  const double initial_rTr = ko_reduce(r, r, n);
  double old_rTr = initial_rTr, alpha, beta = 0.0, new_rTr = initial_rTr;
#pragma omp parallel for schedule(static)
  for (long k = 0; k < n; ++k) p[k] = r[k] + beta * p[k];               /* :90 */
  long it = 0;
  int status = MAX_ITER;
  for (long i = 0; i < max_iter; ++i) {
    it = i + 1;
    ko_spmv(coef, p, Ap, NX, NY);                                        /* :94 */
    const double pAp = ko_reduce(p, Ap, n);                              /* :95 */
    if (pAp == 0.0) { status = ALPHA_DIV_ZERO; break; }
    alpha = old_rTr / pAp;                                               /* :96 */
#pragma omp parallel for schedule(static)
    for (long k = 0; k < n; ++k) x[k] += alpha * p[k];                   /* update_x :68-71 */
#pragma omp parallel for schedule(static)
    for (long k = 0; k < n; ++k) r[k] -= alpha * Ap[k];                  /* update_r :73-76 */
    new_rTr = ko_reduce(r, r, n);                                        /* :101 */
    if (sqrt(new_rTr) < sqrt(initial_rTr) * eps) { status = CONVERGED; break; }
    beta = new_rTr / old_rTr;                                            /* :107 */
#pragma omp parallel for schedule(static)
    for (long k = 0; k < n; ++k) p[k] = r[k] + beta * p[k];              /* update_p :78-81 */
    old_rTr = new_rTr;
  }
  *iters_out = it;
  *rel_res = sqrt(new_rTr / initial_rTr);
  return status;
}

/* BICGSolver.solve, bicgsolver.py:88-135.  work = 9 planes (r, r_tld, p, p_hat, Ap, s, s_hat, Ashat, t). */
int ko_bicgstab(const double* coef, const double* b, double* x, int NX, int NY, double eps, long max_iter, double* work,
                long* iters_out, double* rel_res) {
  const long n = (long)NX * NY;
  double *r = work, *r_tld = work + n, *p = work + 2 * n, *p_hat = work + 3 * n, *Ap = work + 4 * n, *s = work + 5 * n,
         *s_hat = work + 6 * n, *Ashat = work + 7 * n, *t = work + 8 * n;
  /* init: bicgsolver.py:21-36 */
#pragma omp parallel for schedule(static)
  for (long k = 0; k < n; ++k) { r[k] = b[k]; r_tld[k] = b[k]; Ap[k] = 0.0; Ashat[k] = 0.0; p[k] = 0.0; x[k] = 0.0; }
  double omega = 1.0, alpha = 1.0, beta = 1.0, rho_1 = 1.0, rho = 0.0;
  const double initial_rTr = ko_reduce(r, r, n);                         /* :90 */
  double rTr = initial_rTr;
  long it = 0;
  int status = MAX_ITER;
  for (long i = 0; i < max_iter; ++i) {
    rho = ko_reduce(r, r_tld, n);                                        /* :95 */
    if (rho == 0.0) { status = RHO_ZERO; break; }                        /* :96-98 */
    if (i == 0) {
      k_copy(r, p, n);                                                   /* :100 */
    } else {
      if (omega == 0.0) { status = OMEGA_DIV_ZERO; break; }
      beta = (rho / rho_1) * (alpha / omega);                            /* :102 */
#pragma omp parallel for schedule(static)
      for (long k = 0; k < n; ++k) p[k] = r[k] + beta * (p[k] - omega * Ap[k]);   /* update_p :43-46 */
    }
    k_copy(p, p_hat, n);                                                 /* update_phat :48-52 */
    ko_spmv(coef, p_hat, Ap, NX, NY);                                    /* :106 */
    const double alpha_lower = ko_reduce(r_tld, Ap, n);                  /* :107 */
    if (alpha_lower == 0.0) { status = ALPHA_DIV_ZERO; break; }
    alpha = rho / alpha_lower;                                           /* :108 */
#pragma omp parallel for schedule(static)
    for (long k = 0; k < n; ++k) s[k] = r[k] - alpha * Ap[k];            /* update_s :60-63 */
    k_copy(s, s_hat, n);                                                 /* update_shat :54-58 */
    ko_spmv(coef, s_hat, Ashat, NX, NY);                                 /* :113 */
    k_copy(Ashat, t, n);                                                 /* :114 */
    const double omega_upper = ko_reduce(t, s, n);                       /* :116 */
    const double omega_lower = ko_reduce(t, t, n);                       /* :117 */
    if (omega_lower == 0.0) { status = OMEGA_DIV_ZERO; break; }
    omega = omega_upper / omega_lower;                                   /* :118 */
#pragma omp parallel for schedule(static)
    for (long k = 0; k < n; ++k) x[k] += alpha * p_hat[k] + omega * s_hat[k];     /* update_x :78-81 */
#pragma omp parallel for schedule(static)
    for (long k = 0; k < n; ++k) r[k] = s[k] - omega * t[k];             /* update_r :83-86 */
    rTr = ko_reduce(r, r, n);                                            /* :123 */
    it = i + 1;
    if (sqrt(rTr / initial_rTr) < eps) { status = CONVERGED; break; }    /* :130 */
    rho_1 = rho;                                                         /* :135 */
  }
  *iters_out = it;
  *rel_res = sqrt(rTr / initial_rTr);
  return status;
}
