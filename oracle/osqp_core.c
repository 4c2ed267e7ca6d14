/*
 * oracle/osqp_core.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the ADMM QP solver the reference calls at
 * sco_py/sco_osqp/osqp_utils.py:195-216 (osqp.OSQP().setup(...).solve()).
 * The arithmetic lives in the third-party package osqp==0.6.2.post5
 * (poetry.lock:101-102) with qdldl==0.1.5.post2 (poetry.lock:231-232); neither
 * is vendored under /root/reference nor installable in this image, so this
 * file restates the *published* OSQP algorithm (Stellato et al., "OSQP: an
 * operator splitting solver for quadratic programs", and the 0.6.x solver
 * structure: Ruiz equilibration + cost scaling, per-constraint rho, quasi-
 * definite KKT LDL^T, over-relaxed ADMM, termination / infeasibility tests
 * every check_termination iterations, unscaling).  SURVEY.md Appendix B is the
 * spec followed.  Parity is pinned on the reference's own known-answer tests
 * (tests/sco_osqp/*.py run unmodified on top of this core through the
 * oracle/shims/osqp module -- see tests/test_oracle.py) and on golden QPs captured
 * at the reference's own call site (tests/golden/, oracle/gen_golden.py).
 * ADMM-iterate parity with upstream OSQP is NOT pinned.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this library.  The product path (sco_py_b200/csrc) never does.
 *
 * minimise 0.5 x'Px + q'x   s.t.  l <= Ax <= u
 * P: n x n upper-triangular CSC, A: m x n CSC.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_EQ_OVER_RHO_INEQ 1e3
#define RHO_TOL 1e-4

enum {
  ST_SOLVED = 1,
  ST_SOLVED_INACCURATE = 2,
  ST_PRIMAL_INFEASIBLE_INACCURATE = 3,
  ST_DUAL_INFEASIBLE_INACCURATE = 4,
  ST_MAX_ITER_REACHED = -2,
  ST_PRIMAL_INFEASIBLE = -3,
  ST_DUAL_INFEASIBLE = -4,
  ST_NON_CVX = -7,
  ST_UNSOLVED = -10
};

typedef struct {
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  double adaptive_rho_tolerance;
  int32_t max_iter, scaling, check_termination, adaptive_rho;
  int32_t adaptive_rho_interval, scaled_termination;
} osqp_oracle_settings;

typedef struct {
  int32_t status_val, iter, rho_updates, setup_error;
  double obj_val, pri_res, dua_res, rho_estimate;
} osqp_oracle_info;

/* ------------------------------------------------------------------ */
/* small vector helpers                                                */
static double norm_inf(const double *v, int n) {
  double m = 0.0;
  for (int i = 0; i < n; i++) {
    double a = fabs(v[i]);
    if (a > m) m = a;
  }
  return m;
}
static double scaled_norm_inf(const double *s, const double *v, int n) {
  double m = 0.0;
  for (int i = 0; i < n; i++) {
    double a = fabs(s[i] * v[i]);
    if (a > m) m = a;
  }
  return m;
}
static double limit_scaling(double v) {
  v = v < MIN_SCALING ? 1.0 : v;
  v = v > MAX_SCALING ? MAX_SCALING : v;
  return v;
}

/* ------------------------------------------------------------------ */
/* workspace                                                           */
typedef struct {
  int n, m;
  /* scaled copies of the data */
  int *Pp, *Pi;
  double *Px;
  int *Ap, *Ai;
  double *Ax;
  double *q, *l, *u;
  double *D, *E, *Dinv, *Einv;
  double c, cinv;
  /* rho */
  double rho;
  double *rho_vec, *rho_inv_vec;
  int *constr_type;
  /* KKT (upper-tri CSC of the permuted matrix) + factor */
  int N;
  int *perm, *iperm;  /* perm[new] = old */
  int *Kp, *Ki;
  double *Kx;
  int *Kdiag_rho; /* position in Kx of the -1/rho_i diagonal for constraint i */
  int *K_from;    /* for numeric refresh: not needed, values recomputed */
  int *Lp, *Li, *Parent, *Lnz, *Flag, *Pattern;
  double *Lx, *Dd, *Y, *bp;
  /* iterates */
  double *x, *z, *y, *x_prev, *z_prev, *xz_tilde, *delta_x, *delta_y;
  double *Ax_, *Px_, *Aty, *Pdx, *Adx, *Atdy;
  double *tmpn, *tmpm;
} work_t;

static void mat_vec_A(const work_t *w, const double *x, double *out) {
  memset(out, 0, sizeof(double) * w->m);
  for (int j = 0; j < w->n; j++)
    for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) out[w->Ai[p]] += w->Ax[p] * x[j];
}
static void mat_tvec_A(const work_t *w, const double *y, double *out) {
  for (int j = 0; j < w->n; j++) {
    double s = 0.0;
    for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) s += w->Ax[p] * y[w->Ai[p]];
    out[j] = s;
  }
}
/* symmetric product with the upper-triangular P */
static void mat_vec_P(const work_t *w, const double *x, double *out) {
  memset(out, 0, sizeof(double) * w->n);
  for (int j = 0; j < w->n; j++)
    for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
      int i = w->Pi[p];
      out[i] += w->Px[p] * x[j];
      if (i != j) out[j] += w->Px[p] * x[i];
    }
}

/* ------------------------------------------------------------------ */
/* Ruiz equilibration + cost scaling (Appendix B step 1)               */
static void scale_data(work_t *w, int scaling) {
  int n = w->n, m = w->m;
  double *Dt = w->tmpn, *Et = w->tmpm;
  double *colP = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
  for (int i = 0; i < n; i++) w->D[i] = w->Dinv[i] = 1.0;
  for (int i = 0; i < m; i++) w->E[i] = w->Einv[i] = 1.0;
  w->c = 1.0;
  for (int it = 0; it < scaling; it++) {
    /* column inf-norms of [P A'; A 0] */
    for (int j = 0; j < n; j++) Dt[j] = 0.0;
    for (int i = 0; i < m; i++) Et[i] = 0.0;
    for (int j = 0; j < n; j++)
      for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
        int i = w->Pi[p];
        double a = fabs(w->Px[p]);
        if (a > Dt[j]) Dt[j] = a;
        if (i != j && a > Dt[i]) Dt[i] = a;
      }
    for (int j = 0; j < n; j++)
      for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) {
        double a = fabs(w->Ax[p]);
        if (a > Dt[j]) Dt[j] = a;
        if (a > Et[w->Ai[p]]) Et[w->Ai[p]] = a;
      }
    for (int j = 0; j < n; j++) Dt[j] = 1.0 / sqrt(limit_scaling(Dt[j]));
    for (int i = 0; i < m; i++) Et[i] = 1.0 / sqrt(limit_scaling(Et[i]));
    /* P <- Dt P Dt ; A <- Et A Dt ; q <- Dt q */
    for (int j = 0; j < n; j++)
      for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) w->Px[p] *= Dt[w->Pi[p]] * Dt[j];
    for (int j = 0; j < n; j++)
      for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) w->Ax[p] *= Et[w->Ai[p]] * Dt[j];
    for (int j = 0; j < n; j++) w->q[j] *= Dt[j];
    for (int j = 0; j < n; j++) w->D[j] *= Dt[j];
    for (int i = 0; i < m; i++) w->E[i] *= Et[i];
    /* cost normalisation */
    for (int j = 0; j < n; j++) colP[j] = 0.0;
    for (int j = 0; j < n; j++)
      for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
        int i = w->Pi[p];
        double a = fabs(w->Px[p]);
        if (a > colP[j]) colP[j] = a;
        if (i != j && a > colP[i]) colP[i] = a;
      }
    double mean = 0.0;
    for (int j = 0; j < n; j++) mean += colP[j];
    mean = n > 0 ? mean / n : 0.0;
    double nq = limit_scaling(norm_inf(w->q, n));
    double ct = mean > nq ? mean : nq;
    ct = 1.0 / limit_scaling(ct);
    for (int p = 0; p < w->Pp[n]; p++) w->Px[p] *= ct;
    for (int j = 0; j < n; j++) w->q[j] *= ct;
    w->c *= ct;
  }
  w->cinv = 1.0 / w->c;
  for (int j = 0; j < n; j++) w->Dinv[j] = 1.0 / w->D[j];
  for (int i = 0; i < m; i++) w->Einv[i] = 1.0 / w->E[i];
  for (int i = 0; i < m; i++) {
    w->l[i] *= w->E[i];
    w->u[i] *= w->E[i];
  }
  free(colP);
}

/* ------------------------------------------------------------------ */
/* rho vector (Appendix B step 2)                                      */
static void set_rho_vec(work_t *w, double rho) {
  rho = rho < RHO_MIN ? RHO_MIN : rho;
  rho = rho > RHO_MAX ? RHO_MAX : rho;
  w->rho = rho;
  for (int i = 0; i < w->m; i++) {
    if (w->l[i] < -OSQP_INFTY * MIN_SCALING && w->u[i] > OSQP_INFTY * MIN_SCALING) {
      w->constr_type[i] = -1;
      w->rho_vec[i] = RHO_MIN;
    } else if (w->u[i] - w->l[i] < RHO_TOL) {
      w->constr_type[i] = 1;
      w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * rho;
    } else {
      w->constr_type[i] = 0;
      w->rho_vec[i] = rho;
    }
    w->rho_inv_vec[i] = 1.0 / w->rho_vec[i];
  }
}

/* ------------------------------------------------------------------ */
/* fill-reducing ordering: greedy minimum degree on the KKT graph      */
static int popcnt_row(const uint64_t *r, int W) {
  int c = 0;
  for (int k = 0; k < W; k++) c += __builtin_popcountll(r[k]);
  return c;
}
static void min_degree_order(int N, const int *Kp, const int *Ki, int *perm) {
  int W = (N + 63) / 64;
  uint64_t *adj = (uint64_t *)calloc((size_t)N * W, sizeof(uint64_t));
  char *done = (char *)calloc(N, 1);
  int *deg = (int *)malloc(sizeof(int) * N);
  for (int j = 0; j < N; j++)
    for (int p = Kp[j]; p < Kp[j + 1]; p++) {
      int i = Ki[p];
      if (i != j) {
        adj[(size_t)i * W + j / 64] |= 1ull << (j % 64);
        adj[(size_t)j * W + i / 64] |= 1ull << (i % 64);
      }
    }
  for (int i = 0; i < N; i++) deg[i] = popcnt_row(adj + (size_t)i * W, W);
  for (int k = 0; k < N; k++) {
    int best = -1, bd = 1 << 30;
    for (int i = 0; i < N; i++)
      if (!done[i] && deg[i] < bd) {
        bd = deg[i];
        best = i;
      }
    perm[k] = best;
    done[best] = 1;
    uint64_t *rb = adj + (size_t)best * W;
    for (int wd = 0; wd < W; wd++) {
      uint64_t bits = rb[wd];
      while (bits) {
        int b = __builtin_ctzll(bits);
        bits &= bits - 1;
        int u = wd * 64 + b;
        uint64_t *ru = adj + (size_t)u * W;
        for (int k2 = 0; k2 < W; k2++) ru[k2] |= rb[k2];
        ru[best / 64] &= ~(1ull << (best % 64));
        ru[u / 64] &= ~(1ull << (u % 64));
        deg[u] = popcnt_row(ru, W);
      }
    }
  }
  free(adj);
  free(done);
  free(deg);
}

/* ------------------------------------------------------------------ */
/* KKT = [P + sigma I, A'; A, -diag(1/rho)], stored permuted, upper CSC */
static int build_kkt(work_t *w, double sigma) {
  int n = w->n, m = w->m, N = n + m;
  w->N = N;
  int nnzP = w->Pp[n], nnzA = w->Ap[n];
  int cap = nnzP + n + nnzA + m;
  /* triplets of the unpermuted upper triangle */
  int *ti = (int *)malloc(sizeof(int) * cap), *tj = (int *)malloc(sizeof(int) * cap);
  double *tv = (double *)malloc(sizeof(double) * cap);
  int *trho = (int *)malloc(sizeof(int) * cap); /* constraint index if -1/rho diag */
  int nt = 0;
  for (int j = 0; j < n; j++) {
    int has_diag = 0;
    for (int p = w->Pp[j]; p < w->Pp[j + 1]; p++) {
      int i = w->Pi[p];
      ti[nt] = i; tj[nt] = j; tv[nt] = w->Px[p] + (i == j ? sigma : 0.0); trho[nt] = -1;
      if (i == j) has_diag = 1;
      nt++;
    }
    if (!has_diag) { ti[nt] = j; tj[nt] = j; tv[nt] = sigma; trho[nt] = -1; nt++; }
  }
  for (int j = 0; j < n; j++)
    for (int p = w->Ap[j]; p < w->Ap[j + 1]; p++) {
      ti[nt] = j; tj[nt] = n + w->Ai[p]; tv[nt] = w->Ax[p]; trho[nt] = -1; nt++;
    }
  for (int i = 0; i < m; i++) {
    ti[nt] = n + i; tj[nt] = n + i; tv[nt] = -w->rho_inv_vec[i]; trho[nt] = i; nt++;
  }
  /* temporary unpermuted CSC pattern for the ordering */
  int *Tp = (int *)calloc(N + 1, sizeof(int)), *Ti = (int *)malloc(sizeof(int) * nt);
  for (int k = 0; k < nt; k++) Tp[tj[k] + 1]++;
  for (int j = 0; j < N; j++) Tp[j + 1] += Tp[j];
  int *pos = (int *)malloc(sizeof(int) * (N + 1));
  memcpy(pos, Tp, sizeof(int) * (N + 1));
  for (int k = 0; k < nt; k++) Ti[pos[tj[k]]++] = ti[k];
  w->perm = (int *)malloc(sizeof(int) * N);
  w->iperm = (int *)malloc(sizeof(int) * N);
  min_degree_order(N, Tp, Ti, w->perm);
  for (int k = 0; k < N; k++) w->iperm[w->perm[k]] = k;
  /* permuted upper-tri CSC */
  w->Kp = (int *)calloc(N + 1, sizeof(int));
  w->Ki = (int *)malloc(sizeof(int) * nt);
  w->Kx = (double *)malloc(sizeof(double) * nt);
  w->Kdiag_rho = (int *)malloc(sizeof(int) * (m > 0 ? m : 1));
  for (int k = 0; k < nt; k++) {
    int a = w->iperm[ti[k]], b = w->iperm[tj[k]];
    int col = a > b ? a : b;
    w->Kp[col + 1]++;
  }
  for (int j = 0; j < N; j++) w->Kp[j + 1] += w->Kp[j];
  memcpy(pos, w->Kp, sizeof(int) * (N + 1));
  for (int k = 0; k < nt; k++) {
    int a = w->iperm[ti[k]], b = w->iperm[tj[k]];
    int col = a > b ? a : b, row = a > b ? b : a;
    int p = pos[col]++;
    w->Ki[p] = row;
    w->Kx[p] = tv[k];
    if (trho[k] >= 0) w->Kdiag_rho[trho[k]] = p;
  }
  free(ti); free(tj); free(tv); free(trho); free(Tp); free(Ti); free(pos);
  /* symbolic (up-looking LDL', elimination-tree based) */
  w->Lp = (int *)malloc(sizeof(int) * (N + 1));
  w->Parent = (int *)malloc(sizeof(int) * N);
  w->Lnz = (int *)malloc(sizeof(int) * N);
  w->Flag = (int *)malloc(sizeof(int) * N);
  w->Pattern = (int *)malloc(sizeof(int) * N);
  for (int k = 0; k < N; k++) {
    w->Parent[k] = -1; w->Flag[k] = k; w->Lnz[k] = 0;
    for (int p = w->Kp[k]; p < w->Kp[k + 1]; p++) {
      int i = w->Ki[p];
      if (i < k)
        for (; w->Flag[i] != k; i = w->Parent[i]) {
          if (w->Parent[i] == -1) w->Parent[i] = k;
          w->Lnz[i]++;
          w->Flag[i] = k;
        }
    }
  }
  w->Lp[0] = 0;
  for (int k = 0; k < N; k++) w->Lp[k + 1] = w->Lp[k] + w->Lnz[k];
  int lnz = w->Lp[N];
  w->Li = (int *)malloc(sizeof(int) * (lnz > 0 ? lnz : 1));
  w->Lx = (double *)malloc(sizeof(double) * (lnz > 0 ? lnz : 1));
  w->Dd = (double *)malloc(sizeof(double) * N);
  w->Y = (double *)malloc(sizeof(double) * N);
  w->bp = (double *)malloc(sizeof(double) * N);
  return 0;
}

/* numeric LDL'; returns number of positive pivots, or -1 on a zero pivot */
static int factor_kkt(work_t *w) {
  int N = w->N, npos = 0;
  for (int k = 0; k < N; k++) {
    int top = N;
    w->Y[k] = 0.0; w->Flag[k] = k; w->Lnz[k] = 0;
    for (int p = w->Kp[k]; p < w->Kp[k + 1]; p++) {
      int i = w->Ki[p];
      w->Y[i] += w->Kx[p];
      int len = 0;
      for (; w->Flag[i] != k; i = w->Parent[i]) {
        w->Pattern[len++] = i;
        w->Flag[i] = k;
      }
      while (len > 0) w->Pattern[--top] = w->Pattern[--len];
    }
    w->Dd[k] = w->Y[k];
    w->Y[k] = 0.0;
    for (; top < N; top++) {
      int i = w->Pattern[top];
      double yi = w->Y[i];
      w->Y[i] = 0.0;
      int p2 = w->Lp[i] + w->Lnz[i];
      for (int p = w->Lp[i]; p < p2; p++) w->Y[w->Li[p]] -= w->Lx[p] * yi;
      double lki = yi / w->Dd[i];
      w->Dd[k] -= lki * yi;
      w->Li[p2] = k;
      w->Lx[p2] = lki;
      w->Lnz[i]++;
    }
    if (w->Dd[k] == 0.0) return -1;
    if (w->Dd[k] > 0.0) npos++;
  }
  return npos;
}

static void solve_kkt(work_t *w, double *b) {
  int N = w->N;
  double *x = w->bp;
  for (int k = 0; k < N; k++) x[k] = b[w->perm[k]];
  for (int j = 0; j < N; j++) {
    double xj = x[j];
    for (int p = w->Lp[j]; p < w->Lp[j + 1]; p++) x[w->Li[p]] -= w->Lx[p] * xj;
  }
  for (int j = 0; j < N; j++) x[j] /= w->Dd[j];
  for (int j = N - 1; j >= 0; j--) {
    double s = x[j];
    for (int p = w->Lp[j]; p < w->Lp[j + 1]; p++) s -= w->Lx[p] * x[w->Li[p]];
    x[j] = s;
  }
  for (int k = 0; k < N; k++) b[w->perm[k]] = x[k];
}

/* ------------------------------------------------------------------ */
/* residuals / termination (Appendix B steps 5-6)                      */
typedef struct { double pri_res, dua_res; } resid_t;

static double compute_pri_res(work_t *w, int scaled_term) {
  mat_vec_A(w, w->x, w->Ax_);
  for (int i = 0; i < w->m; i++) w->z_prev[i] = w->Ax_[i] - w->z[i];
  if (!scaled_term) return scaled_norm_inf(w->Einv, w->z_prev, w->m);
  return norm_inf(w->z_prev, w->m);
}
static double compute_pri_tol(work_t *w, double ea, double er, int scaled_term) {
  double a, b;
  if (!scaled_term) {
    a = scaled_norm_inf(w->Einv, w->z, w->m);
    b = scaled_norm_inf(w->Einv, w->Ax_, w->m);
  } else {
    a = norm_inf(w->z, w->m);
    b = norm_inf(w->Ax_, w->m);
  }
  return ea + er * (a > b ? a : b);
}
static double compute_dua_res(work_t *w, int scaled_term) {
  mat_vec_P(w, w->x, w->Px_);
  mat_tvec_A(w, w->y, w->Aty);
  for (int j = 0; j < w->n; j++) w->x_prev[j] = w->q[j] + w->Px_[j] + w->Aty[j];
  if (!scaled_term) return w->cinv * scaled_norm_inf(w->Dinv, w->x_prev, w->n);
  return norm_inf(w->x_prev, w->n);
}
static double compute_dua_tol(work_t *w, double ea, double er, int scaled_term) {
  double a, b, c3, mx;
  if (!scaled_term) {
    a = scaled_norm_inf(w->Dinv, w->q, w->n);
    b = scaled_norm_inf(w->Dinv, w->Aty, w->n);
    c3 = scaled_norm_inf(w->Dinv, w->Px_, w->n);
    mx = a > b ? a : b; mx = mx > c3 ? mx : c3;
    mx *= w->cinv;
  } else {
    a = norm_inf(w->q, w->n); b = norm_inf(w->Aty, w->n); c3 = norm_inf(w->Px_, w->n);
    mx = a > b ? a : b; mx = mx > c3 ? mx : c3;
  }
  return ea + er * mx;
}

static int is_primal_infeasible(work_t *w, double eps, int scaled_term) {
  int m = w->m, n = w->n;
  double *dy = w->delta_y;
  for (int i = 0; i < m; i++) {
    if (w->u[i] > OSQP_INFTY * MIN_SCALING) {
      if (w->l[i] < -OSQP_INFTY * MIN_SCALING) dy[i] = 0.0;
      else dy[i] = dy[i] < 0.0 ? dy[i] : 0.0;
    } else if (w->l[i] < -OSQP_INFTY * MIN_SCALING) {
      dy[i] = dy[i] > 0.0 ? dy[i] : 0.0;
    }
  }
  double nrm = scaled_term ? norm_inf(dy, m) : scaled_norm_inf(w->E, dy, m);
  if (nrm > eps) {
    double lhs = 0.0;
    for (int i = 0; i < m; i++)
      lhs += w->u[i] * (dy[i] > 0.0 ? dy[i] : 0.0) + w->l[i] * (dy[i] < 0.0 ? dy[i] : 0.0);
    if (lhs < -eps * nrm) {
      mat_tvec_A(w, dy, w->Atdy);
      double r = scaled_term ? norm_inf(w->Atdy, n) : scaled_norm_inf(w->Dinv, w->Atdy, n);
      return r < eps * nrm;
    }
  }
  return 0;
}

static int is_dual_infeasible(work_t *w, double eps, int scaled_term) {
  int m = w->m, n = w->n;
  double nrm, cost_scaling;
  if (!scaled_term) { nrm = scaled_norm_inf(w->D, w->delta_x, n); cost_scaling = w->c; }
  else { nrm = norm_inf(w->delta_x, n); cost_scaling = 1.0; }
  if (nrm > eps) {
    double qdx = 0.0;
    for (int j = 0; j < n; j++) qdx += w->q[j] * w->delta_x[j];
    if (qdx < -cost_scaling * eps * nrm) {
      mat_vec_P(w, w->delta_x, w->Pdx);
      if (!scaled_term) for (int j = 0; j < n; j++) w->Pdx[j] *= w->Dinv[j];
      if (norm_inf(w->Pdx, n) < cost_scaling * eps * nrm) {
        mat_vec_A(w, w->delta_x, w->Adx);
        if (!scaled_term) for (int i = 0; i < m; i++) w->Adx[i] *= w->Einv[i];
        for (int i = 0; i < m; i++) {
          if ((w->u[i] < OSQP_INFTY * MIN_SCALING && w->Adx[i] > eps * nrm) ||
              (w->l[i] > -OSQP_INFTY * MIN_SCALING && w->Adx[i] < -eps * nrm))
            return 0;
        }
        return 1;
      }
    }
  }
  return 0;
}

static double compute_obj(work_t *w) {
  double o = 0.0;
  mat_vec_P(w, w->x, w->tmpn);
  for (int j = 0; j < w->n; j++) o += 0.5 * w->x[j] * w->tmpn[j] + w->q[j] * w->x[j];
  return o * w->cinv;
}

/* returns 1 if a terminal status was set */
static int check_termination(work_t *w, const osqp_oracle_settings *s, osqp_oracle_info *info,
                             int approximate) {
  double ea = s->eps_abs, er = s->eps_rel, epi = s->eps_prim_inf, edi = s->eps_dual_inf;
  int prim_ok = 0, dual_ok = 0, pinf = 0, dinf = 0;
  if (info->pri_res > OSQP_INFTY || info->dua_res > OSQP_INFTY) {
    info->status_val = ST_NON_CVX;
    info->obj_val = NAN;
    return 1;
  }
  if (approximate) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
  if (w->m == 0) prim_ok = 1;
  else {
    double ep = compute_pri_tol(w, ea, er, s->scaled_termination);
    if (info->pri_res < ep) prim_ok = 1;
    else pinf = is_primal_infeasible(w, epi, s->scaled_termination);
  }
  double ed = compute_dua_tol(w, ea, er, s->scaled_termination);
  if (info->dua_res < ed) dual_ok = 1;
  else dinf = is_dual_infeasible(w, edi, s->scaled_termination);
  if (prim_ok && dual_ok) {
    info->status_val = approximate ? ST_SOLVED_INACCURATE : ST_SOLVED;
    return 1;
  }
  if (pinf) {
    info->status_val = approximate ? ST_PRIMAL_INFEASIBLE_INACCURATE : ST_PRIMAL_INFEASIBLE;
    info->obj_val = OSQP_INFTY;
    return 1;
  }
  if (dinf) {
    info->status_val = approximate ? ST_DUAL_INFEASIBLE_INACCURATE : ST_DUAL_INFEASIBLE;
    info->obj_val = -OSQP_INFTY;
    return 1;
  }
  return 0;
}

static void update_info(work_t *w, const osqp_oracle_settings *s, osqp_oracle_info *info, int iter) {
  info->iter = iter;
  info->obj_val = compute_obj(w);
  info->pri_res = w->m == 0 ? 0.0 : compute_pri_res(w, s->scaled_termination);
  info->dua_res = compute_dua_res(w, s->scaled_termination);
}

static double compute_rho_estimate(work_t *w) {
  /* z_prev holds Ax - z and x_prev holds q + Px + A'y from the last update_info */
  double pri = norm_inf(w->z_prev, w->m), dua = norm_inf(w->x_prev, w->n);
  double a = norm_inf(w->z, w->m), b = norm_inf(w->Ax_, w->m);
  pri /= ((a > b ? a : b) + 1e-10);
  double d1 = norm_inf(w->q, w->n), d2 = norm_inf(w->Aty, w->n), d3 = norm_inf(w->Px_, w->n);
  double mx = d1 > d2 ? d1 : d2; mx = mx > d3 ? mx : d3;
  dua /= (mx + 1e-10);
  double r = w->rho * sqrt(pri / (dua + 1e-10));
  r = r < RHO_MIN ? RHO_MIN : r;
  r = r > RHO_MAX ? RHO_MAX : r;
  return r;
}

#define FREE(p) do { if (p) free(p); } while (0)
static void free_work(work_t *w) {
  FREE(w->Pp); FREE(w->Pi); FREE(w->Px); FREE(w->Ap); FREE(w->Ai); FREE(w->Ax);
  FREE(w->q); FREE(w->l); FREE(w->u); FREE(w->D); FREE(w->E); FREE(w->Dinv); FREE(w->Einv);
  FREE(w->rho_vec); FREE(w->rho_inv_vec); FREE(w->constr_type);
  FREE(w->perm); FREE(w->iperm); FREE(w->Kp); FREE(w->Ki); FREE(w->Kx); FREE(w->Kdiag_rho);
  FREE(w->Lp); FREE(w->Li); FREE(w->Parent); FREE(w->Lnz); FREE(w->Flag); FREE(w->Pattern);
  FREE(w->Lx); FREE(w->Dd); FREE(w->Y); FREE(w->bp);
  FREE(w->x); FREE(w->z); FREE(w->y); FREE(w->x_prev); FREE(w->z_prev); FREE(w->xz_tilde);
  FREE(w->delta_x); FREE(w->delta_y); FREE(w->Ax_); FREE(w->Px_); FREE(w->Aty);
  FREE(w->Pdx); FREE(w->Adx); FREE(w->Atdy); FREE(w->tmpn); FREE(w->tmpm);
}

static double *dalloc(int n) { return (double *)calloc(n > 0 ? n : 1, sizeof(double)); }
static int *icopy(const int *s, int n) {
  int *d = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
  memcpy(d, s, sizeof(int) * n);
  return d;
}
static double *dcopy(const double *s, int n) {
  double *d = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
  memcpy(d, s, sizeof(double) * n);
  return d;
}

void osqp_oracle_default_settings(osqp_oracle_settings *s) {
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6; s->eps_abs = 1e-3; s->eps_rel = 1e-3;
  s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4; s->adaptive_rho_tolerance = 5.0;
  s->max_iter = 4000; s->scaling = 10; s->check_termination = 25; s->adaptive_rho = 1;
  s->adaptive_rho_interval = 0; s->scaled_termination = 0;
}

/*
 * Returns 0 on success (info->status_val holds the OSQP status), a positive
 * setup error otherwise: 1 bad data (l > u), 4 non-convex / singular KKT.
 * x (n), y (m) receive the unscaled solution (NaN for infeasible statuses).
 * Optional outputs (may be NULL): D_out (n), E_out (m), c_out (1) -- the
 * scaling actually used, exposed for the parity tests of the CUDA path.
 */
int osqp_oracle_solve(int n, int m, const int *Pp, const int *Pi, const double *Px,
                      const double *q, const int *Ap, const int *Ai, const double *Ax,
                      const double *l, const double *u, const osqp_oracle_settings *s,
                      double *x_out, double *y_out, osqp_oracle_info *info,
                      double *D_out, double *E_out, double *c_out) {
  work_t W;
  work_t *w = &W;
  memset(w, 0, sizeof(W));
  memset(info, 0, sizeof(*info));
  info->status_val = ST_UNSOLVED;
  w->n = n; w->m = m;
  for (int i = 0; i < m; i++)
    if (l[i] > u[i]) { info->setup_error = 1; return 1; }
  w->Pp = icopy(Pp, n + 1); w->Pi = icopy(Pi, Pp[n]); w->Px = dcopy(Px, Pp[n]);
  w->Ap = icopy(Ap, n + 1); w->Ai = icopy(Ai, Ap[n]); w->Ax = dcopy(Ax, Ap[n]);
  w->q = dcopy(q, n); w->l = dcopy(l, m); w->u = dcopy(u, m);
  for (int i = 0; i < m; i++) {
    if (w->l[i] < -OSQP_INFTY) w->l[i] = -OSQP_INFTY;
    if (w->u[i] > OSQP_INFTY) w->u[i] = OSQP_INFTY;
  }
  w->D = dalloc(n); w->Dinv = dalloc(n); w->E = dalloc(m); w->Einv = dalloc(m);
  w->tmpn = dalloc(n); w->tmpm = dalloc(m);
  w->rho_vec = dalloc(m); w->rho_inv_vec = dalloc(m);
  w->constr_type = (int *)calloc(m > 0 ? m : 1, sizeof(int));
  w->x = dalloc(n); w->z = dalloc(m); w->y = dalloc(m); w->x_prev = dalloc(n);
  w->z_prev = dalloc(m); w->xz_tilde = dalloc(n + m); w->delta_x = dalloc(n);
  w->delta_y = dalloc(m); w->Ax_ = dalloc(m); w->Px_ = dalloc(n); w->Aty = dalloc(n);
  w->Pdx = dalloc(n); w->Adx = dalloc(m); w->Atdy = dalloc(n);

  if (s->scaling > 0) scale_data(w, s->scaling);
  else {
    for (int i = 0; i < n; i++) w->D[i] = w->Dinv[i] = 1.0;
    for (int i = 0; i < m; i++) w->E[i] = w->Einv[i] = 1.0;
    w->c = w->cinv = 1.0;
  }
  if (D_out) memcpy(D_out, w->D, sizeof(double) * n);
  if (E_out) memcpy(E_out, w->E, sizeof(double) * m);
  if (c_out) *c_out = w->c;
  set_rho_vec(w, s->rho);
  build_kkt(w, s->sigma);
  if (factor_kkt(w) != n) { info->setup_error = 4; free_work(w); return 4; }

  const double alpha = s->alpha, sigma = s->sigma;
  int interval = s->adaptive_rho_interval;
  if (s->adaptive_rho && interval == 0) {
    /* upstream derives the interval from wall-clock setup time; without
       profiling it falls back to 4 * check_termination (or 100). */
    interval = s->check_termination ? 4 * s->check_termination : 100;
  }
  int iter, can_check = 0, terminated = 0;
  for (iter = 1; iter <= s->max_iter; iter++) {
    double *t;
    t = w->x; w->x = w->x_prev; w->x_prev = t;
    t = w->z; w->z = w->z_prev; w->z_prev = t;
    /* x_tilde, z_tilde */
    for (int j = 0; j < n; j++) w->xz_tilde[j] = sigma * w->x_prev[j] - w->q[j];
    for (int i = 0; i < m; i++) w->xz_tilde[n + i] = w->z_prev[i] - w->rho_inv_vec[i] * w->y[i];
    solve_kkt(w, w->xz_tilde);
    for (int i = 0; i < m; i++)
      w->xz_tilde[n + i] = w->z_prev[i] + w->rho_inv_vec[i] * (w->xz_tilde[n + i] - w->y[i]);
    /* x */
    for (int j = 0; j < n; j++) {
      w->x[j] = alpha * w->xz_tilde[j] + (1.0 - alpha) * w->x_prev[j];
      w->delta_x[j] = w->x[j] - w->x_prev[j];
    }
    /* z */
    for (int i = 0; i < m; i++) {
      double v = alpha * w->xz_tilde[n + i] + (1.0 - alpha) * w->z_prev[i] + w->rho_inv_vec[i] * w->y[i];
      v = v < w->l[i] ? w->l[i] : v;
      v = v > w->u[i] ? w->u[i] : v;
      w->z[i] = v;
    }
    /* y */
    for (int i = 0; i < m; i++) {
      w->delta_y[i] = w->rho_vec[i] *
                      (alpha * w->xz_tilde[n + i] + (1.0 - alpha) * w->z_prev[i] - w->z[i]);
      w->y[i] += w->delta_y[i];
    }
    can_check = s->check_termination && (iter % s->check_termination == 0);
    if (can_check) {
      update_info(w, s, info, iter);
      if (check_termination(w, s, info, 0)) { terminated = 1; break; }
    }
    if (s->adaptive_rho && interval && (iter % interval == 0)) {
      if (!can_check) update_info(w, s, info, iter);
      double rn = compute_rho_estimate(w);
      info->rho_estimate = rn;
      if (rn > w->rho * s->adaptive_rho_tolerance || rn < w->rho / s->adaptive_rho_tolerance) {
        set_rho_vec(w, rn);
        for (int i = 0; i < m; i++) w->Kx[w->Kdiag_rho[i]] = -w->rho_inv_vec[i];
        if (factor_kkt(w) != n) { info->setup_error = 4; free_work(w); return 4; }
        info->rho_updates++;
      }
    }
  }
  if (!terminated) {
    iter = s->max_iter;
    if (!can_check) {
      update_info(w, s, info, iter);
      terminated = check_termination(w, s, info, 0);
    }
    info->iter = iter;
    if (!terminated && info->status_val == ST_UNSOLVED) {
      if (!check_termination(w, s, info, 1)) info->status_val = ST_MAX_ITER_REACHED;
    }
  }
  info->rho_estimate = compute_rho_estimate(w);
  int st = info->status_val;
  if (st != ST_PRIMAL_INFEASIBLE && st != ST_PRIMAL_INFEASIBLE_INACCURATE &&
      st != ST_DUAL_INFEASIBLE && st != ST_DUAL_INFEASIBLE_INACCURATE && st != ST_NON_CVX) {
    for (int j = 0; j < n; j++) x_out[j] = w->D[j] * w->x[j];
    for (int i = 0; i < m; i++) y_out[i] = w->E[i] * w->y[i] * w->cinv;
  } else {
    for (int j = 0; j < n; j++) x_out[j] = NAN;
    for (int i = 0; i < m; i++) y_out[i] = NAN;
  }
  free_work(w);
  return 0;
}
