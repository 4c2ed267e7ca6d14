// sco_qp.cuh -- the in-house batched OSQP-style ADMM QP solve, one QP per thread team,
// working set resident in shared memory.
//
// Replaces osqp.OSQP().setup()/solve() as driven by sco_py/sco_osqp/osqp_utils.py:195-216 for
// the penalty QP of SURVEY.md Appendix A:
//
//   min  0.5 x' Psym x + q'x + pi * 1's
//   s.t. A_lin x in [l_lin, u_lin]
//        (J.*M) x - s1 (+ s2) {<=, =} -b          (each row present `kd` times, quirk C-3)
//        lbx <= x <= ubx ,  s >= 0
//
// The iteration is the one of OSQP 0.6.2 (Ruiz equilibration + cost scaling, per-row rho,
// alpha-relaxed ADMM from x=z=y=0, termination tests every check_termination iterations on the
// unscaled residuals) so that iterates agree with the CPU oracle to
// rounding.  What differs is the linear algebra: instead of an LDL' of the (n_q+m_q) KKT matrix
// the team eliminates the constraint block (nu = R(A xt - z) + y), then the slack variables
// (their diagonal blocks are 1x1 / 2x2), and keeps the explicit inverse of the remaining
// n x n SPD matrix  S = Psym^ + sigma I + A_x' R A_x - (slack Schur terms)  in shared memory, so
// one ADMM iteration is three small mat-vecs with no sequential triangular solve.
// Duplicate penalty rows are never materialised: they share z, y, E and rho by symmetry, so they
// enter as the weight `kd` wherever A' multiplies a row-space vector.
#pragma once
#include "sco_device.cuh"

struct QPArgs {
  const double *prm;  // problem parameter block
  const double *Jg;   // unscaled Jacobian entries, global layout (may be null when !use_pen)
  double pi;          // slack cost (compounded penalty weight)
  double kd;          // copies of every penalty row
  int use_pen;        // 0: no penalty rows / slacks (closest feasible point)
  int closest;        // objective |x - xs|^2
  int has_hq;         // add the degree-2 model of a non-quadratic objective term (w.Hq, w.gq)
};

struct QPResult {
  int status, iters;
  double pri_res, dua_res;
#ifdef SCO_TIMING
  long long cyc_loop, cyc_check, cyc_setup;  // clock64() ticks: ADMM loop (incl. checks), checks, setup
  long long cyc_c[5];                        // stages of the termination test
#endif
};

// rho for a row with scaled bounds [l, u] (OSQP set_rho_vec)
__device__ __forceinline__ double rho_of(double l, double u, double rho) {
  if (l < -OSQP_INFTY * OSQP_MIN_SCALING && u > OSQP_INFTY * OSQP_MIN_SCALING) return OSQP_RHO_MIN;
  if (u - l < OSQP_RHO_TOL) return OSQP_RHO_EQ_OVER_RHO_INEQ * rho;
  return rho;
}

__device__ __forceinline__ double clampd(double v, double lo, double hi) {
  v = v < lo ? lo : v;
  v = v > hi ? hi : v;
  return v;
}

// project dy on the polar of the recession cone of [l,u] (OSQP is_primal_infeasible)
__device__ __forceinline__ double proj_dy(double dy, double l, double u) {
  if (u > OSQP_INFTY * OSQP_MIN_SCALING) {
    if (l < -OSQP_INFTY * OSQP_MIN_SCALING) return 0.0;
    return fmin(dy, 0.0);
  } else if (l < -OSQP_INFTY * OSQP_MIN_SCALING) {
    return fmax(dy, 0.0);
  }
  return dy;
}

#include "sco_dense.cuh"

template <int TEAM, int DK>
struct QPSolver {
  const DevStruct &S;
  const DevSettings &st;
  QPW &w;
  const QPArgs &a;
  const int tid;
  const int n, m_lin, m_nl, ms;
  const double *Qg;  // global Q (may be null = zeros)
  double c;          // cost scaling
  double rho;        // current scalar rho

  __device__ QPSolver(const DevStruct &S_, const DevSettings &st_, QPW &w_, const QPArgs &a_)
      : S(S_), st(st_), w(w_), a(a_), tid(threadIdx.x), n(S_.n), m_lin(S_.m_lin),
        m_nl(a_.use_pen ? S_.m_nl : 0), ms(S_.m_nl), Qg(field_ptr(S_, S_.Q, a_.prm)), c(1.0),
        rho(st_.rho) {}

  __device__ __forceinline__ void sync() { Team<TEAM>::sync(); }

  // Every method starts with SCO_QP_LOCALS: by-value copies that shadow the members of the same name
  // (see DevIdx), followed by the small helpers as lambdas over those locals.
  //   psym(i, j)          entry of the symmetrised objective matrix (osqp_utils.py:153-163)
  //   gatherAT(j, vl, vp) A' * (row-space vector) for user variable j: linear rows (vl) and penalty
  //                       rows (vp; the caller folds the multiplicity kd into vp); bound row by the caller
  //   lin_row_dot / pen_row_dot   one row of A times a variable-space vector
  //   pcol_norm(j)        column j of c D |Psym| D
#define SCO_QP_LOCALS                                                                                       \
  const DevStruct &SS = this->S;                                                                            \
  const QPW w = this->w;                                                                                    \
  const DevIdx S(this->S);                                                                                  \
  const DevSettings st = this->st;                                                                          \
  const QPArgs a = this->a;                                                                                 \
  const int tid = this->tid, n = this->n, m_lin = this->m_lin, m_nl = this->m_nl, ms = this->ms;            \
  const double *const Qg = this->Qg;                                                                        \
  double c = this->c, rho = this->rho;                                                                      \
  (void)SS; (void)ms; (void)rho;                                                                            \
  auto psym = [&](int i, int j) -> double {                                                                 \
    if (a.closest) return i == j ? 2.0 : 0.0;                                                               \
    double v = Qg ? 0.5 * (Qg[i * n + j] + Qg[j * n + i]) : 0.0;                                            \
    if (a.has_hq) v += 0.5 * (w.Hq[i * n + j] + w.Hq[j * n + i]); /* prob.py:348-367 */                     \
    return v;                                                                                               \
  };                                                                                                        \
  auto gatherAT = [&](int j, Sh vl, Sh vp) -> double {                                                      \
    double acc = 0.0;                                                                                       \
    if (m_lin)                                                                                              \
      for (int p = S.lin_cptr[j]; p < S.lin_cptr[j + 1]; p++) acc += w.Als[S.lin_centry[p]] * vl[S.lin_crow[p]]; \
    if (m_nl) {                                                                                             \
      double accp = 0.0;                                                                                    \
      for (int p = S.pc_ptr[j]; p < S.pc_ptr[j + 1]; p++) accp += w.Js[S.pc_e[p]] * vp[S.pc_r[p]];          \
      acc += accp;                                                                                          \
    }                                                                                                       \
    return acc;                                                                                             \
  };                                                                                                        \
  auto lin_row_dot = [&](int r, Sh v) -> double {                                                           \
    double acc = 0.0;                                                                                       \
    for (int p = S.lin_rowptr[r]; p < S.lin_rowptr[r + 1]; p++) acc += w.Als[p] * v[S.lin_col[p]];          \
    return acc;                                                                                             \
  };                                                                                                        \
  auto pen_row_dot = [&](int i, Sh v) -> double {                                                           \
    const int so = S.row_soff[i], go = S.row_goff[i], wd = S.row_w[i];                                      \
    double acc = 0.0;                                                                                       \
    for (int k = 0; k < wd; k++) acc += w.Js[so + k] * v[S.jcol_g[go + k]];                                 \
    return acc;                                                                                             \
  };                                                                                                        \
  auto pcol_norm = [&](int j) -> double {                                                                   \
    double cp = 0.0;                                                                                        \
    for (int i = 0; i < n; i++) cp = fmax(cp, w.D[i] * fabs(w.Sm[i * n + j]));                              \
    return cp * c * w.D[j];                                                                                 \
  };                                                                                                        \
  (void)psym; (void)gatherAT; (void)lin_row_dot; (void)pen_row_dot; (void)pcol_norm;

  // ================================================================== setup
  // expects (unscaled): w.lb/w.ub bounds on x, w.bb = b, w.msk, w.xs (closest point target)
  __device__ __noinline__ void load_and_scale() {
    SCO_QP_LOCALS
    const double *qg = field_ptr(SS, SS.q, a.prm);
    const double *llg = field_ptr(SS, SS.lin_l, a.prm), *ulg = field_ptr(SS, SS.lin_u, a.prm);
    for (int e = tid; e < n * n; e += TEAM) w.Sm[e] = psym(e / n, e % n);
    for (int j = tid; j < n; j += TEAM) {
      w.qh[j] = a.closest ? -2.0 * w.xs[j] : (qg ? qg[j] : 0.0) + (a.has_hq ? w.gq[j] : 0.0);
      w.D[j] = 1.0;
      w.bx[j] = 1.0;
      w.Eb[j] = 1.0;
    }
    for (int e = tid; e < S.nnz_lin; e += TEAM) w.Als[e] = S.lin_val[e];
    for (int r = tid; r < m_lin; r += TEAM) w.El[r] = 1.0;
    for (int i = tid; i < m_nl; i += TEAM) {
      const int so = S.row_soff[i], go = S.row_goff[i], wd = S.row_w[i];
      const uint32_t mk = w.msk[i];
      for (int k = 0; k < wd; k++) w.Js[so + k] = ((mk >> k) & 1u) ? a.Jg[go + k] : 0.0;
      w.Ep[i] = 1.0;
      w.sl[i] = -1.0; w.bs[i] = 1.0; w.Ds[i] = 1.0; w.Es[i] = 1.0;
      if (S.row_eq[i]) {
        w.sl[ms + i] = 1.0; w.bs[ms + i] = 1.0; w.Ds[ms + i] = 1.0; w.Es[ms + i] = 1.0;
      }
    }
    c = 1.0;
    sync();
    const int nq = n + (a.use_pen ? S.n_slack : 0);
    // Ruiz passes.  temporaries: Dt -> xt, Etb -> xt2, Etl -> wl, Etp -> wp, Dts -> gs, Ets -> dys
    for (int it = 0; it < st.scaling; it++) {
      for (int j = tid; j < n; j += TEAM) {
        const double cp = pcol_norm(j);
        double ca = fabs(w.bx[j]);
        if (m_lin)
          for (int p = S.lin_cptr[j]; p < S.lin_cptr[j + 1]; p++)
            ca = fmax(ca, fabs(w.Als[S.lin_centry[p]]));
        if (m_nl)
          for (int p = S.pc_ptr[j]; p < S.pc_ptr[j + 1]; p++) ca = fmax(ca, fabs(w.Js[S.pc_e[p]]));
        w.xt[j] = 1.0 / sqrt(limit_scaling(fmax(cp, ca)));
        w.xt2[j] = 1.0 / sqrt(limit_scaling(fabs(w.bx[j])));
      }
      for (int r = tid; r < m_lin; r += TEAM) {
        double rn = 0.0;
        for (int p = S.lin_rowptr[r]; p < S.lin_rowptr[r + 1]; p++) rn = fmax(rn, fabs(w.Als[p]));
        w.wl[r] = 1.0 / sqrt(limit_scaling(rn));
      }
      for (int i = tid; i < m_nl; i += TEAM) {
        const int so = S.row_soff[i], wd = S.row_w[i], eq = S.row_eq[i];
        double rn = fabs(w.sl[i]);
        if (eq) rn = fmax(rn, fabs(w.sl[ms + i]));
        for (int k = 0; k < wd; k++) rn = fmax(rn, fabs(w.Js[so + k]));
        w.wp[i] = 1.0 / sqrt(limit_scaling(rn));
        for (int k2 = 0; k2 <= eq; k2++) {
          const int si = k2 * ms + i;
          w.gs[si] = 1.0 / sqrt(limit_scaling(fmax(fabs(w.sl[si]), fabs(w.bs[si]))));
          w.dys[si] = 1.0 / sqrt(limit_scaling(fabs(w.bs[si])));
        }
      }
      sync();
      for (int r = tid; r < m_lin; r += TEAM) {
        const double er = w.wl[r];
        for (int p = S.lin_rowptr[r]; p < S.lin_rowptr[r + 1]; p++)
          w.Als[p] *= er * w.xt[S.lin_col[p]];
        w.El[r] *= er;
      }
      for (int i = tid; i < m_nl; i += TEAM) {
        const int so = S.row_soff[i], go = S.row_goff[i], wd = S.row_w[i], eq = S.row_eq[i];
        const double er = w.wp[i];
        for (int k = 0; k < wd; k++) w.Js[so + k] *= er * w.xt[S.jcol_g[go + k]];
        w.Ep[i] *= er;
        for (int k2 = 0; k2 <= eq; k2++) {
          const int si = k2 * ms + i;
          w.sl[si] *= er * w.gs[si];
          w.bs[si] *= w.dys[si] * w.gs[si];
          w.Ds[si] *= w.gs[si];
          w.Es[si] *= w.dys[si];
        }
      }
      sync();  // column data of Js / Als final before D changes (pcol_norm below reads D only)
      for (int j = tid; j < n; j += TEAM) {
        w.bx[j] *= w.xt2[j] * w.xt[j];
        w.Eb[j] *= w.xt2[j];
        w.qh[j] *= w.xt[j];
        w.D[j] *= w.xt[j];
      }
      sync();
      // cost normalisation
      double vs[1] = {0.0}, vm[1] = {0.0};
      for (int j = tid; j < n; j += TEAM) {
        vs[0] += pcol_norm(j);
        vm[0] = fmax(vm[0], fabs(w.qh[j]));
      }
      if (m_nl) {
        const double cq = fabs(c * a.pi);
        for (int i = tid; i < m_nl; i += TEAM)
          for (int k2 = 0; k2 <= S.row_eq[i]; k2++) vm[0] = fmax(vm[0], cq * w.Ds[k2 * ms + i]);
      }
      Team<TEAM>::reduce_sum(vs, w.red);
      Team<TEAM>::reduce_max(vm, w.red);
      const double mean = vs[0] / (double)nq;
      const double ct = 1.0 / limit_scaling(fmax(mean, limit_scaling(vm[0])));
      for (int j = tid; j < n; j += TEAM) w.qh[j] *= ct;
      c *= ct;
      sync();
    }
    // scaled bounds
    for (int j = tid; j < n; j += TEAM) {
      const double eb = w.Eb[j];
      w.lb[j] = eb * fmax(w.lb[j], -OSQP_INFTY);
      w.ub[j] = eb * fmin(w.ub[j], OSQP_INFTY);
    }
    for (int r = tid; r < m_lin; r += TEAM) {
      w.ll[r] = w.El[r] * fmax(llg ? llg[r] : 0.0, -OSQP_INFTY);
      w.ul[r] = w.El[r] * fmin(ulg ? ulg[r] : 0.0, OSQP_INFTY);
    }
    for (int i = tid; i < m_nl; i += TEAM) {
      const double hi = clampd(-w.bb[i], -OSQP_INFTY, OSQP_INFTY);
      w.up[i] = w.Ep[i] * hi;
      w.lp[i] = S.row_eq[i] ? w.Ep[i] * hi : -OSQP_INFTY * w.Ep[i];
    }
    this->c = c;
    sync();
  }

  __device__ void set_rho() {
    SCO_QP_LOCALS
    rho = fmin(fmax(rho, OSQP_RHO_MIN), OSQP_RHO_MAX);
    this->rho = rho;
    for (int j = tid; j < n; j += TEAM) w.rb[j] = rho_of(w.lb[j], w.ub[j], rho);
    for (int r = tid; r < m_lin; r += TEAM) w.rl[r] = rho_of(w.ll[r], w.ul[r], rho);
    for (int i = tid; i < m_nl; i += TEAM) {
      w.rp[i] = rho_of(w.lp[i], w.up[i], rho);
      for (int k2 = 0; k2 <= S.row_eq[i]; k2++)
        w.rs[k2 * ms + i] = rho_of(0.0, OSQP_INFTY * w.Es[k2 * ms + i], rho);
    }
    sync();
  }

  // S = Psym^ + sigma I + A_x' R A_x - slack Schur terms, then S <- S^-1 (in place).
  // `reload`: Sm does not hold the unscaled Psym any more (rho update) -> fetch it again.
  __device__ __noinline__ void assemble_and_invert(bool reload) {
    SCO_QP_LOCALS
    const double sigma = st.sigma;
    for (int i = tid; i < m_nl; i += TEAM) {
      const double kr = a.kd * w.rp[i];
      const double s1 = w.sl[i], b1 = w.bs[i];
      const double m11 = sigma + kr * s1 * s1 + w.rs[i] * b1 * b1;
      double coef;
      if (S.row_eq[i]) {
        const double s2 = w.sl[ms + i], b2 = w.bs[ms + i];
        const double m22 = sigma + kr * s2 * s2 + w.rs[ms + i] * b2 * b2;
        const double m12 = kr * s1 * s2;
        const double det = m11 * m22 - m12 * m12;
        const double i11 = m22 / det, i22 = m11 / det, i12 = -m12 / det;
        w.Minv[3 * i] = i11; w.Minv[3 * i + 1] = i12; w.Minv[3 * i + 2] = i22;
        const double h1 = kr * (i11 * s1 + i12 * s2), h2 = kr * (i12 * s1 + i22 * s2);
        w.hs[i] = h1; w.hs[ms + i] = h2;
        coef = kr - kr * (s1 * h1 + s2 * h2);
      } else {
        const double i11 = 1.0 / m11;
        w.Minv[3 * i] = i11; w.Minv[3 * i + 1] = 0.0; w.Minv[3 * i + 2] = 0.0;
        const double h1 = kr * i11 * s1;
        w.hs[i] = h1;
        coef = kr - kr * s1 * h1;
      }
      w.wp[i] = coef;
    }
    for (int e = tid; e < n * n; e += TEAM) {
      const int i = e / n, j = e % n;
      const double pv = reload ? psym(i, j) : w.Sm[e];
      double v = c * w.D[i] * pv * w.D[j];
      if (i == j) v += sigma + w.rb[j] * w.bx[j] * w.bx[j];
      w.Sm[e] = v;
    }
    sync();
    for (int j = tid; j < n; j += TEAM) {
      if (m_lin) {
        for (int p = S.lin_cptr[j]; p < S.lin_cptr[j + 1]; p++) {
          const int r = S.lin_crow[p];
          const double f = w.rl[r] * w.Als[S.lin_centry[p]];
          for (int q2 = S.lin_rowptr[r]; q2 < S.lin_rowptr[r + 1]; q2++)
            w.Sm[S.lin_col[q2] * n + j] += f * w.Als[q2];
        }
      }
      if (m_nl) {
        for (int p = S.pc_ptr[j]; p < S.pc_ptr[j + 1]; p++) {
          const int r = S.pc_r[p];
          const double f = w.wp[r] * w.Js[S.pc_e[p]];
          const int so = S.row_soff[r], go = S.row_goff[r], wd = S.row_w[r];
          for (int k = 0; k < wd; k++) w.Sm[S.jcol_g[go + k] * n + j] += f * w.Js[so + k];
        }
      }
    }
    sync();
    // in-place Gauss-Jordan inverse (S is SPD: no pivoting).  xt = pivot row, xt2 = pivot column
    for (int k = 0; k < n; k++) {
      const double d = 1.0 / w.Sm[k * n + k];
      for (int j = tid; j < n; j += TEAM) {
        w.xt[j] = w.Sm[k * n + j] * d;
        w.xt2[j] = w.Sm[j * n + k];
      }
      sync();
      for (int e = tid; e < n * n; e += TEAM) {
        const int i = e / n, j = e % n;
        double v;
        if (i == k) v = (j == k) ? d : w.xt[j];
        else if (j == k) v = -w.xt2[i] * d;
        else v = w.Sm[e] - w.xt2[i] * w.xt[j];
        w.Sm[e] = v;
      }
      sync();
    }
  }

  // ================================================================== termination
  // Returns a terminal status or 0.  Scratch: xt (D.*x), wp (kd*yp).  sc[] receives the scaled
  // norms needed by the rho estimate: {|Ax-z|, |z|, |Ax|, |Px+q+A'y|, |q|, |A'y|, |Px|}.
  __device__ __noinline__ int check(int approximate, double &pri_res_out, double &dua_res_out, double *sc) {
    SCO_QP_LOCALS
    double ea = st.eps_abs, er = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
    if (approximate) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
    const double cinv = 1.0 / c;
    // unscaled: v[0]=pri_res v[1]=|z/E| v[2]=|Ax/E| v[3]=dua_res*c v[4]=|q/D| v[5]=|A'y/D| v[6]=|Px/D|
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    double u[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int r = tid; r < m_lin; r += TEAM) {
      const double ax = lin_row_dot(r, w.x), ei = 1.0 / w.El[r], z = w.zl[r];
      v[0] = fmax(v[0], fabs((ax - z) * ei)); v[1] = fmax(v[1], fabs(z * ei)); v[2] = fmax(v[2], fabs(ax * ei));
      u[0] = fmax(u[0], fabs(ax - z)); u[1] = fmax(u[1], fabs(z)); u[2] = fmax(u[2], fabs(ax));
    }
    for (int i = tid; i < m_nl; i += TEAM) {
      const int eq = S.row_eq[i];
      double ax = pen_row_dot(i, w.x) + w.sl[i] * w.s[i];
      if (eq) ax += w.sl[ms + i] * w.s[ms + i];
      double ei = 1.0 / w.Ep[i], z = w.zp[i];
      v[0] = fmax(v[0], fabs((ax - z) * ei)); v[1] = fmax(v[1], fabs(z * ei)); v[2] = fmax(v[2], fabs(ax * ei));
      u[0] = fmax(u[0], fabs(ax - z)); u[1] = fmax(u[1], fabs(z)); u[2] = fmax(u[2], fabs(ax));
      for (int k2 = 0; k2 <= eq; k2++) {
        const int si = k2 * ms + i;
        const double axs = w.bs[si] * w.s[si];
        ei = 1.0 / w.Es[si]; z = w.zs[si];
        v[0] = fmax(v[0], fabs((axs - z) * ei)); v[1] = fmax(v[1], fabs(z * ei)); v[2] = fmax(v[2], fabs(axs * ei));
        u[0] = fmax(u[0], fabs(axs - z)); u[1] = fmax(u[1], fabs(z)); u[2] = fmax(u[2], fabs(axs));
        // slack variable: q^ + A'y (its P block is zero)
        const double di = 1.0 / w.Ds[si];
        const double qs = c * a.pi * w.Ds[si];
        const double aty = a.kd * w.sl[si] * w.yp[i] + w.bs[si] * w.ys[si];
        v[3] = fmax(v[3], fabs((qs + aty) * di)); v[4] = fmax(v[4], fabs(qs * di)); v[5] = fmax(v[5], fabs(aty * di));
        u[3] = fmax(u[3], fabs(qs + aty)); u[4] = fmax(u[4], fabs(qs)); u[5] = fmax(u[5], fabs(aty));
      }
    }
    for (int i = tid; i < m_nl; i += TEAM) w.wp[i] = a.kd * w.yp[i];
    for (int j = tid; j < n; j += TEAM) w.xt[j] = w.D[j] * w.x[j];
    sync();
    for (int j = tid; j < n; j += TEAM) {
      const double ax = w.bx[j] * w.x[j], ei = 1.0 / w.Eb[j], z = w.zb[j];
      v[0] = fmax(v[0], fabs((ax - z) * ei)); v[1] = fmax(v[1], fabs(z * ei)); v[2] = fmax(v[2], fabs(ax * ei));
      u[0] = fmax(u[0], fabs(ax - z)); u[1] = fmax(u[1], fabs(z)); u[2] = fmax(u[2], fabs(ax));
      double px = 0.0;
      if (a.closest) px = 2.0 * w.xt[j];
      else if (Qg || a.has_hq)
        for (int k = 0; k < n; k++) px += psym(k, j) * w.xt[k];
      px *= c * w.D[j];
      const double aty = gatherAT(j, w.yl, w.wp) + w.bx[j] * w.yb[j];
      const double di = 1.0 / w.D[j], q = w.qh[j];
      v[3] = fmax(v[3], fabs((q + px + aty) * di)); v[4] = fmax(v[4], fabs(q * di));
      v[5] = fmax(v[5], fabs(aty * di)); v[6] = fmax(v[6], fabs(px * di));
      u[3] = fmax(u[3], fabs(q + px + aty)); u[4] = fmax(u[4], fabs(q)); u[5] = fmax(u[5], fabs(aty));
      u[6] = fmax(u[6], fabs(px));
    }
    Team<TEAM>::reduce_max(v, w.red);
    if (sc) {
      Team<TEAM>::reduce_max(u, w.red);
      for (int k = 0; k < 7; k++) sc[k] = u[k];
    }
    const double pri_res = v[0], dua_res = cinv * v[3];
    pri_res_out = pri_res;
    dua_res_out = dua_res;
    if (pri_res > OSQP_INFTY || dua_res > OSQP_INFTY) return -7;
    const double eps_p = ea + er * fmax(v[1], v[2]);
    const double eps_d = ea + er * cinv * fmax(v[4], fmax(v[5], v[6]));
    const bool prim_ok = pri_res < eps_p, dual_ok = dua_res < eps_d;
    if (prim_ok && dual_ok) return approximate ? 2 : 1;
    bool pinf = false, dinf = false;
    if (!prim_ok) pinf = primal_infeasible(epi);
    if (!dual_ok) dinf = dual_infeasible(edi);
    if (pinf) return approximate ? 3 : -3;
    if (dinf) return approximate ? 4 : -4;
    return 0;
  }

  // delta_y of the last iteration: dyl (lin), dyp (pen), dyb (x bounds), dys (slack bounds)
  __device__ __noinline__ bool primal_infeasible(double eps) {
    SCO_QP_LOCALS
    double nv[1] = {0.0}, lhs[1] = {0.0};
    for (int r = tid; r < m_lin; r += TEAM) {
      const double dy = proj_dy(w.dyl[r], w.ll[r], w.ul[r]);
      w.dyl[r] = dy;
      nv[0] = fmax(nv[0], fabs(w.El[r] * dy));
      lhs[0] += w.ul[r] * fmax(dy, 0.0) + w.ll[r] * fmin(dy, 0.0);
    }
    for (int i = tid; i < m_nl; i += TEAM) {
      const double dy = proj_dy(w.dyp[i], w.lp[i], w.up[i]);
      w.dyp[i] = dy;
      w.wp[i] = a.kd * dy;
      nv[0] = fmax(nv[0], fabs(w.Ep[i] * dy));
      lhs[0] += a.kd * (w.up[i] * fmax(dy, 0.0) + w.lp[i] * fmin(dy, 0.0));
      for (int k2 = 0; k2 <= S.row_eq[i]; k2++) {
        const int si = k2 * ms + i;
        const double us = OSQP_INFTY * w.Es[si];
        const double dys = proj_dy(w.dys[si], 0.0, us);
        w.dys[si] = dys;
        nv[0] = fmax(nv[0], fabs(w.Es[si] * dys));
        lhs[0] += us * fmax(dys, 0.0);
      }
    }
    for (int j = tid; j < n; j += TEAM) {
      const double dy = proj_dy(w.dyb[j], w.lb[j], w.ub[j]);
      w.dyb[j] = dy;
      nv[0] = fmax(nv[0], fabs(w.Eb[j] * dy));
      lhs[0] += w.ub[j] * fmax(dy, 0.0) + w.lb[j] * fmin(dy, 0.0);
    }
    Team<TEAM>::reduce_max(nv, w.red);
    Team<TEAM>::reduce_sum(lhs, w.red);
    sync();
    bool res = false;
    if (nv[0] > eps && lhs[0] < -eps * nv[0]) {
      double mv[1] = {0.0};
      for (int j = tid; j < n; j += TEAM) {
        const double aty = gatherAT(j, w.dyl, w.wp) + w.bx[j] * w.dyb[j];
        mv[0] = fmax(mv[0], fabs(aty / w.D[j]));
      }
      for (int i = tid; i < m_nl; i += TEAM)
        for (int k2 = 0; k2 <= S.row_eq[i]; k2++) {
          const int si = k2 * ms + i;
          const double aty = w.sl[si] * w.wp[i] + w.bs[si] * w.dys[si];
          mv[0] = fmax(mv[0], fabs(aty / w.Ds[si]));
        }
      Team<TEAM>::reduce_max(mv, w.red);
      res = mv[0] < eps * nv[0];
    }
    sync();
    return res;
  }

  // delta_x of the last iteration: dxv (user variables), dss (slacks)
  __device__ __noinline__ bool dual_infeasible(double eps) {
    SCO_QP_LOCALS
    double nv[1] = {0.0}, qd[1] = {0.0};
    for (int j = tid; j < n; j += TEAM) {
      nv[0] = fmax(nv[0], fabs(w.D[j] * w.dxv[j]));
      qd[0] += w.qh[j] * w.dxv[j];
      w.xt[j] = w.D[j] * w.dxv[j];
    }
    for (int i = tid; i < m_nl; i += TEAM)
      for (int k2 = 0; k2 <= S.row_eq[i]; k2++) {
        const int si = k2 * ms + i;
        nv[0] = fmax(nv[0], fabs(w.Ds[si] * w.dss[si]));
        qd[0] += c * a.pi * w.Ds[si] * w.dss[si];
      }
    Team<TEAM>::reduce_max(nv, w.red);
    Team<TEAM>::reduce_sum(qd, w.red);
    sync();
    bool res = false;
    const double thr = c * eps * nv[0];
    if (nv[0] > eps && qd[0] < -thr) {
      double pv[1] = {0.0};
      for (int j = tid; j < n; j += TEAM) {
        double px = 0.0;
        if (a.closest) px = 2.0 * w.xt[j];
        else if (Qg || a.has_hq)
          for (int k = 0; k < n; k++) px += psym(k, j) * w.xt[k];
        pv[0] = fmax(pv[0], fabs(c * px));  // Dinv .* (c D Psym D dx) = c * Psym (D dx)
      }
      Team<TEAM>::reduce_max(pv, w.red);
      if (pv[0] < thr) {
        const double lim = eps * nv[0];
        double bad[1] = {0.0};
        for (int r = tid; r < m_lin; r += TEAM) {
          const double adx = lin_row_dot(r, w.dxv) / w.El[r];
          if ((w.ul[r] < OSQP_INFTY * OSQP_MIN_SCALING && adx > lim) ||
              (w.ll[r] > -OSQP_INFTY * OSQP_MIN_SCALING && adx < -lim)) bad[0] = 1.0;
        }
        for (int i = tid; i < m_nl; i += TEAM) {
          double adx = pen_row_dot(i, w.dxv) + w.sl[i] * w.dss[i];
          if (S.row_eq[i]) adx += w.sl[ms + i] * w.dss[ms + i];
          adx /= w.Ep[i];
          if ((w.up[i] < OSQP_INFTY * OSQP_MIN_SCALING && adx > lim) ||
              (w.lp[i] > -OSQP_INFTY * OSQP_MIN_SCALING && adx < -lim)) bad[0] = 1.0;
          for (int k2 = 0; k2 <= S.row_eq[i]; k2++) {
            const int si = k2 * ms + i;
            const double ads = w.bs[si] * w.dss[si] / w.Es[si];
            if ((OSQP_INFTY * w.Es[si] < OSQP_INFTY * OSQP_MIN_SCALING && ads > lim) || ads < -lim)
              bad[0] = 1.0;
          }
        }
        for (int j = tid; j < n; j += TEAM) {
          const double adx = w.bx[j] * w.dxv[j] / w.Eb[j];
          if ((w.ub[j] < OSQP_INFTY * OSQP_MIN_SCALING && adx > lim) ||
              (w.lb[j] > -OSQP_INFTY * OSQP_MIN_SCALING && adx < -lim)) bad[0] = 1.0;
        }
        Team<TEAM>::reduce_max(bad, w.red);
        res = bad[0] == 0.0;
      }
    }
    sync();
    return res;
  }

  // ================================================================== the ADMM loop
  // On return w.x (user variables) and w.s (slacks) hold the UNSCALED solution.
  // generic shared-memory ADMM loop; returns the status (0 = max_iter reached without a verdict)
  __device__ __noinline__ int generic_loop(int &iter_out, bool &checked_out, QPResult &res) {
    SCO_QP_LOCALS
    const double sigma = st.sigma, alpha = st.alpha, oma = 1.0 - st.alpha;
    for (int j = tid; j < n; j += TEAM) { w.x[j] = 0.0; w.zb[j] = 0.0; w.yb[j] = 0.0; }
    for (int r = tid; r < m_lin; r += TEAM) { w.zl[r] = 0.0; w.yl[r] = 0.0; }
    for (int i = tid; i < m_nl; i += TEAM) {
      w.zp[i] = 0.0; w.yp[i] = 0.0;
      for (int k2 = 0; k2 <= S.row_eq[i]; k2++) {
        const int si = k2 * ms + i;
        w.s[si] = 0.0; w.zs[si] = 0.0; w.ys[si] = 0.0;
      }
    }
    sync();
    const double cpi = c * a.pi;
    int interval = st.adaptive_rho_interval;
    if (st.adaptive_rho && interval == 0) interval = st.check_termination ? 4 * st.check_termination : 100;
    int iter, status = 0;
    bool checked = false;
    for (iter = 1; iter <= st.max_iter; iter++) {
      const bool can_check = st.check_termination && (iter % st.check_termination == 0);
      const bool do_rho = st.adaptive_rho && interval && (iter % interval == 0);
      const bool want_delta = can_check || do_rho;
      // ---- P1: row weights  w = rho z - y, slack elimination
      for (int r = tid; r < m_lin; r += TEAM) w.wl[r] = w.rl[r] * w.zl[r] - w.yl[r];
      for (int i = tid; i < m_nl; i += TEAM) {
        const double wpen = w.rp[i] * w.zp[i] - w.yp[i];
        const double kr = a.kd * w.rp[i];
        const double r1 = sigma * w.s[i] - cpi * w.Ds[i] + a.kd * w.sl[i] * wpen +
                          w.bs[i] * (w.rs[i] * w.zs[i] - w.ys[i]);
        if (S.row_eq[i]) {
          const int i2 = ms + i;
          const double r2 = sigma * w.s[i2] - cpi * w.Ds[i2] + a.kd * w.sl[i2] * wpen +
                            w.bs[i2] * (w.rs[i2] * w.zs[i2] - w.ys[i2]);
          const double g1 = w.Minv[3 * i] * r1 + w.Minv[3 * i + 1] * r2;
          const double g2 = w.Minv[3 * i + 1] * r1 + w.Minv[3 * i + 2] * r2;
          w.gs[i] = g1; w.gs[i2] = g2;
          w.wp[i] = a.kd * wpen - kr * (w.sl[i] * g1 + w.sl[i2] * g2);
        } else {
          const double g1 = w.Minv[3 * i] * r1;
          w.gs[i] = g1;
          w.wp[i] = a.kd * wpen - kr * w.sl[i] * g1;
        }
      }
      sync();
      // ---- P2: reduced right-hand side
      for (int j = tid; j < n; j += TEAM)
        w.xt[j] = sigma * w.x[j] - w.qh[j] + w.bx[j] * (w.rb[j] * w.zb[j] - w.yb[j]) +
                  gatherAT(j, w.wl, w.wp);
      sync();
      // ---- P3: x~ = S^-1 rhs ; x, bound rows
      for (int j = tid; j < n; j += TEAM) {
        double acc0 = 0.0, acc1 = 0.0;
        int k = 0;
        for (; k + 1 < n; k += 2) {
          acc0 += w.Sm[k * n + j] * w.xt[k];
          acc1 += w.Sm[(k + 1) * n + j] * w.xt[k + 1];
        }
        if (k < n) acc0 += w.Sm[k * n + j] * w.xt[k];
        const double xtil = acc0 + acc1;
        w.xt2[j] = xtil;
        const double xo = w.x[j];
        const double xn = alpha * xtil + oma * xo;
        w.x[j] = xn;
        const double zt = w.bx[j] * xtil;
        const double vv = alpha * zt + oma * w.zb[j];
        const double zn = clampd(vv + w.yb[j] / w.rb[j], w.lb[j], w.ub[j]);
        const double dy = w.rb[j] * (vv - zn);
        w.yb[j] += dy;
        w.zb[j] = zn;
        if (want_delta) { w.dxv[j] = xn - xo; w.dyb[j] = dy; }
      }
      sync();
      // ---- P4: rows
      for (int r = tid; r < m_lin; r += TEAM) {
        const double zt = lin_row_dot(r, w.xt2);
        const double vv = alpha * zt + oma * w.zl[r];
        const double zn = clampd(vv + w.yl[r] / w.rl[r], w.ll[r], w.ul[r]);
        const double dy = w.rl[r] * (vv - zn);
        w.yl[r] += dy;
        w.zl[r] = zn;
        if (want_delta) w.dyl[r] = dy;
      }
      for (int i = tid; i < m_nl; i += TEAM) {
        const double t = pen_row_dot(i, w.xt2);
        const int eq = S.row_eq[i];
        double zt = t;
        for (int k2 = 0; k2 <= eq; k2++) {
          const int si = k2 * ms + i;
          const double stil = w.gs[si] - w.hs[si] * t;
          zt += w.sl[si] * stil;
          const double so = w.s[si];
          const double sn = alpha * stil + oma * so;
          w.s[si] = sn;
          const double zts = w.bs[si] * stil;
          const double vs = alpha * zts + oma * w.zs[si];
          const double zns = clampd(vs + w.ys[si] / w.rs[si], 0.0, OSQP_INFTY * w.Es[si]);
          const double dys = w.rs[si] * (vs - zns);
          w.ys[si] += dys;
          w.zs[si] = zns;
          if (want_delta) { w.dss[si] = sn - so; w.dys[si] = dys; }
        }
        const double vv = alpha * zt + oma * w.zp[i];
        const double zn = clampd(vv + w.yp[i] / w.rp[i], w.lp[i], w.up[i]);
        const double dy = w.rp[i] * (vv - zn);
        w.yp[i] += dy;
        w.zp[i] = zn;
        if (want_delta) w.dyp[i] = dy;
      }
      checked = false;
      if (want_delta) {
        sync();
        double sc[7];
        status = check(0, res.pri_res, res.dua_res, do_rho ? sc : nullptr);
        checked = can_check;
        if (can_check && status != 0) break;
        status = 0;
        if (do_rho) {
          double pri = sc[0] / (fmax(sc[1], sc[2]) + 1e-10);
          double dua = sc[3] / (fmax(sc[4], fmax(sc[5], sc[6])) + 1e-10);
          double rn = rho * sqrt(pri / (dua + 1e-10));
          rn = fmin(fmax(rn, OSQP_RHO_MIN), OSQP_RHO_MAX);
          if (rn > rho * 5.0 || rn < rho / 5.0) {
            this->rho = rn;
            set_rho();
            rho = this->rho;
            assemble_and_invert(true);
          }
        }
        sync();
      }
    }
    iter_out = iter;
    checked_out = checked;
    return status;
  }

  __device__ __noinline__ QPResult solve() {
    // dense hinge-only structures: the specialised two-warp solve (same iteration, sco_dense.cuh).
    // The size table is kept in sync with sco_create (sco_abi.cu).
    if constexpr (TEAM == 64 && DK != 0) {
      if (!st.force_generic && !st.adaptive_rho && a.use_pen) {
        if constexpr (DK == 1) return dense_qp_solve<8, 6>(S, st, w, a);
        else if constexpr (DK == 2) return dense_qp_solve<12, 16>(S, st, w, a);
        else if constexpr (DK == 3) return dense_qp_solve<20, 30>(S, st, w, a);
        else return dense_qp_solve<32, 32>(S, st, w, a);
      }
    }
#ifdef SCO_TIMING
    const long long t_begin = clock64();
#endif
    load_and_scale();
    rho = st.rho;
    set_rho();
    assemble_and_invert(false);
    QPResult res;
    res.status = 0; res.iters = 0; res.pri_res = 0.0; res.dua_res = 0.0;
#ifdef SCO_TIMING
    res.cyc_check = 0;
    for (int k = 0; k < 5; k++) res.cyc_c[k] = 0;
    res.cyc_setup = clock64() - t_begin;
    const long long t_loop = clock64();
#endif
    int iter = 0, status = 0;
    bool checked = false;
    status = generic_loop(iter, checked, res);
#ifdef SCO_TIMING
    res.cyc_loop = clock64() - t_loop;
#endif
    if (status == 0) {
      iter = st.max_iter;
      if (!checked) {
        sync();
        status = check(0, res.pri_res, res.dua_res, nullptr);
      }
      if (status == 0) {
        sync();
        status = check(1, res.pri_res, res.dua_res, nullptr);
        if (status == 0) status = -2;
      }
    }
    sync();
    // unscale
    for (int j = tid; j < n; j += TEAM) w.x[j] *= w.D[j];
    for (int i = tid; i < m_nl; i += TEAM)
      for (int k2 = 0; k2 <= S.row_eq[i]; k2++) w.s[k2 * ms + i] *= w.Ds[k2 * ms + i];
    sync();
    res.status = status;
    res.iters = iter;
    return res;
  }
};
