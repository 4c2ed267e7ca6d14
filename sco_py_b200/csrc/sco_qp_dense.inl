// sco_qp_dense.inl -- member template of QPSolver<64> (included inside the struct body).
//
// Register-resident ADMM loop for the dense hinge-only structure (one QUADFORM block, no linear
// rows; BASELINE.json configs[3], the headline workload).  A team of TWO warps owns the QP:
// warp 0, lane i  = penalty row i and its slack      (i < m)
// warp 1, lane j  = user variable j and its box row   (j < n)
// With the lane-local quantities
//     c_j  = sigma x_j - q^_j + bx_j (rho_j zb_j - yb_j)
//     wp_i = kd (rho_i zp_i - yp_i) - kd rho_i sl_i g_i          (slack already eliminated)
// one iteration of the generic loop in sco_qp.cuh is the linear map
//     x~ = S^-1 (c + J' wp)        = [ S^-1      K   ] [ c  ]      K = S^-1 J'
//     t  = J x~                    = [ K'        G   ] [ wp ]      G = J S^-1 J'
// followed by lane-local updates of (x, zb, yb) and (s, zs, ys, zp, yp).  Each lane keeps ITS ROW
// of that (n+m) x (n+m) matrix in registers for the whole QP (the FP64 pipe consumes 64 matrix
// words per clock per SM, four times what shared memory can deliver), so an iteration is: one
// (n+m)-term dot product against the vector (c, wp) read with broadcast 128-bit shared loads,
// the lane-local update, one store of the lane's new c_j / wp_i, ONE barrier.  (c, wp) is
// double-buffered so readers of iteration k never race writers of iteration k+1.
// NP >= n and MP >= m are the (even) compile-time paddings; padded entries are zero.
template <int NP, int MP>
__device__ __forceinline__ int dense_loop(int &iter_out, bool &checked_out, QPResult &res) {
  constexpr int NV = NP + MP;
  const int lane = tid & 31;
  const bool rowwarp = tid < 32;
  const bool isv = !rowwarp && lane < n, isr = rowwarp && lane < m_nl;
  const double sigma = st.sigma, alpha = st.alpha, oma = 1.0 - st.alpha;
  const double cpi = c * a.pi, kd = a.kd;
  const int ldj = (n & 1) ? n : n + 1;
  const Sh vb = w.fbuf;                       // 2 x VB_LD: (c[NP], wp[MP]) double-buffered
  const Sh xS = w.fbuf + 2 * FBUF_LD, yS = w.fbuf + 2 * FBUF_LD + 32;
  const Sh prev = w.fbuf + 2 * FBUF_LD + 64;  // iterates of the iteration before a check: 4 x 64
  const Sh Ks = w.Kd;                         // K[j*m + i], n x m

  // ---- K = S^-1 J' into shared memory
  for (int e = tid; e < n * m_nl; e += 64) {
    const int j = e / m_nl, i = e - j * m_nl;
    double acc = 0.0;
    for (int k = 0; k < n; k++) acc = fma(w.Sm[k * n + j], w.Js[i * ldj + k], acc);
    Ks[e] = acc;
  }
  __syncthreads();
  // ---- this lane's row of [S^-1 K ; K' G] -> registers
  double Mr[NV];
  if (rowwarp) {
#pragma unroll
    for (int k = 0; k < NP; k++) Mr[k] = (isr && k < n) ? Ks[k * m_nl + lane] : 0.0;
#pragma unroll
    for (int l = 0; l < MP; l++) {
      double acc = 0.0;
      if (isr && l < m_nl)
        for (int k = 0; k < n; k++) acc = fma(w.Js[lane * ldj + k], Ks[k * m_nl + l], acc);
      Mr[NP + l] = acc;
    }
  } else {
#pragma unroll
    for (int k = 0; k < NP; k++) Mr[k] = (isv && k < n) ? w.Sm[k * n + lane] : 0.0;
#pragma unroll
    for (int i = 0; i < MP; i++) Mr[NP + i] = (isv && i < m_nl) ? Ks[lane * m_nl + i] : 0.0;
  }
  // ---- lane-local constants.  One set of names for both roles (keeps the register count down):
  //   variable lane: u0=q^      u1=bx  u2=rho_j u3=1/rho_j lo/hi = box            (rho_j varies:
  //                  a trust region narrower than RHO_TOL makes the box row an equality)
  //   row lane     : u0=cpi*Ds  u1=sl  u2=bs    u3=hs      lo=kd*sl hi=up  Mi=Minv
  // Hinge rows and slack-bound rows always carry the scalar rho (their bounds are one-sided),
  // verified by dense_eligible(); their never-active clamps (-1e30 Ep, +1e30 Es) are dropped.
  const double rho_u = rho, rhoi_u = 1.0 / rho;
  double u0, u1, u2, u3, lo, hi, Mi = 0.0;
  if (rowwarp) {
    u0 = isr ? cpi * w.Ds[lane] : 0.0;
    u1 = isr ? w.sl[lane] : 0.0;
    u2 = isr ? w.bs[lane] : 0.0;
    u3 = isr ? w.hs[lane] : 0.0;
    lo = kd * u1;
    hi = isr ? w.up[lane] : 0.0;
    Mi = isr ? w.Minv[3 * lane] : 0.0;
  } else {
    u0 = isv ? w.qh[lane] : 0.0;
    u1 = isv ? w.bx[lane] : 0.0;
    u2 = isv ? w.rb[lane] : 1.0;
    u3 = 1.0 / u2;
    lo = isv ? w.lb[lane] : 0.0;
    hi = isv ? w.ub[lane] : 0.0;
  }
  // ---- iterates: (p0, z0, y0) = (x, zb, yb) or (-, zp, yp); row lanes also (s, zs, ys, g)
  double p0 = 0.0, z0 = 0.0, y0 = 0.0, s = 0.0, zs = 0.0, ys = 0.0;
  double g = Mi * (-u0);
  const double wp0 = -(rho_u * lo) * g;
  const int slot = rowwarp ? NP + lane : lane;           // this lane's entry of (c, wp)
  const bool publishes = rowwarp ? lane < MP : lane < NP;
  for (int e = tid; e < 2 * FBUF_LD + 64 + 256; e += 64) vb[e] = 0.0;
  __syncthreads();
  if (publishes) vb[slot] = rowwarp ? wp0 : -u0;
  int p = 0, iter, status = 0;
  bool checked = false;
  const int max_iter = st.max_iter, chk = st.check_termination;
  int next_check = chk ? chk : max_iter + 1;
  __syncthreads();

  for (iter = 1; iter <= max_iter; iter++) {
    const bool want = iter == next_check;
    if (want || iter == max_iter) {
      // the infeasibility certificates need delta_x / delta_y of the checked iteration: remember
      // the iterates it starts from (kept out of the update code below, which runs every iteration)
      prev[tid] = p0; prev[64 + tid] = y0; prev[128 + tid] = s; prev[192 + tid] = ys;
    }
    const Sh vcur = vb + p * FBUF_LD;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int k = 0; k < NV / 4; k++) {
      const double2 va = vcur.v2(2 * k), vc = vcur.v2(2 * k + 1);
      a0 = fma(Mr[4 * k], va.x, a0);
      a1 = fma(Mr[4 * k + 1], va.y, a1);
      a2 = fma(Mr[4 * k + 2], vc.x, a2);
      a3 = fma(Mr[4 * k + 3], vc.y, a3);
    }
    if (NV % 4) {
      const double2 va = vcur.v2(NV / 2 - 1);
      a0 = fma(Mr[NV - 2], va.x, a0);
      a1 = fma(Mr[NV - 1], va.y, a1);
    }
    const double acc = (a0 + a1) + (a2 + a3);
    double out;
    if (rowwarp) {
      // t = acc ; slack and penalty-row updates, then wp for the next iteration
      const double stil = g - u3 * acc;
      const double zt = acc + u1 * stil;
      const double sn = alpha * stil + oma * s;
      const double vs = alpha * (u2 * stil) + oma * zs;
      const double zns = fmax(vs + ys * rhoi_u, 0.0);
      const double dys = rho_u * (vs - zns);
      ys += dys;
      const double vv = alpha * zt + oma * z0;
      const double zn = fmin(vv + y0 * rhoi_u, hi);
      const double dyp = rho_u * (vv - zn);
      y0 += dyp;
      s = sn;
      zs = zns;
      z0 = zn;
      const double wpen = rho_u * zn - y0;
      const double r1 = sigma * sn - u0 + lo * wpen + u2 * (rho_u * zns - ys);
      g = Mi * r1;
      out = kd * wpen - rho_u * (lo * g);
    } else {
      // x~ = acc ; x and box-row updates, then c for the next iteration
      const double xn = alpha * acc + oma * p0;
      const double vv = alpha * (u1 * acc) + oma * z0;
      const double zn = clampd(vv + y0 * u3, lo, hi);
      const double dyb = u2 * (vv - zn);
      y0 += dyb;
      p0 = xn;
      z0 = zn;
      out = sigma * xn - u0 + u1 * (u2 * zn - y0);
    }
    p ^= 1;
    if (publishes) vb[p * FBUF_LD + slot] = out;
    checked = false;
    if (want) {
#ifdef SCO_TIMING
      const long long tc0 = clock64();
#endif
      next_check += chk;
      checked = true;
      // ================= termination test (OSQP check_termination, unscaled residuals) =========
      if (rowwarp) yS[lane] = kd * y0;
      else xS[lane] = p0;
      if (isr) { w.dss[lane] = s - prev[128 + tid]; w.dys[lane] = ys - prev[192 + tid]; w.dyp[lane] = y0 - prev[64 + tid]; }
      if (isv) { w.dxv[lane] = p0 - prev[tid]; w.dyb[lane] = y0 - prev[64 + tid]; }
      __syncthreads();
      double v[7] = {0, 0, 0, 0, 0, 0, 0};
      if (isr) {
        double ax = u1 * s;
        for (int k = 0; k < n; k++) ax = fma(w.Js[lane * ldj + k], xS[k], ax);
        double ei = 1.0 / w.Ep[lane];
        v[0] = fabs((ax - z0) * ei); v[1] = fabs(z0 * ei); v[2] = fabs(ax * ei);
        const double axs = u2 * s;
        ei = 1.0 / w.Es[lane];
        v[0] = fmax(v[0], fabs((axs - zs) * ei)); v[1] = fmax(v[1], fabs(zs * ei));
        v[2] = fmax(v[2], fabs(axs * ei));
        const double di = 1.0 / w.Ds[lane];
        const double aty = lo * y0 + u2 * ys;
        v[3] = fabs((u0 + aty) * di); v[4] = fabs(u0 * di); v[5] = fabs(aty * di);
      }
      if (isv) {
        const double ax = u1 * p0, ei = 1.0 / w.Eb[lane];
        v[0] = fabs((ax - z0) * ei); v[1] = fabs(z0 * ei); v[2] = fabs(ax * ei);
        double px = 0.0, aty = 0.0;
        for (int k = 0; k < n; k++) px = fma(w.Ph[k * n + lane], xS[k], px);
        for (int i = 0; i < m_nl; i++) aty = fma(w.Js[i * ldj + lane], yS[i], aty);
        aty += u1 * y0;
        const double di = 1.0 / w.D[lane];
        v[3] = fabs((u0 + px + aty) * di); v[4] = fabs(u0 * di);
        v[5] = fabs(aty * di); v[6] = fabs(px * di);
      }
      Team<64>::reduce_max(v, w.red);
      const double cinv = 1.0 / c;
      const double pri_res = v[0], dua_res = cinv * v[3];
      res.pri_res = pri_res;
      res.dua_res = dua_res;
      if (pri_res > OSQP_INFTY || dua_res > OSQP_INFTY) { status = -7; break; }
      const double eps_p = st.eps_abs + st.eps_rel * fmax(v[1], v[2]);
      const double eps_d = st.eps_abs + st.eps_rel * cinv * fmax(v[4], fmax(v[5], v[6]));
      const bool prim_ok = pri_res < eps_p, dual_ok = dua_res < eps_d;
      if (prim_ok && dual_ok) { status = 1; break; }
      // first stage of the infeasibility certificates (lane-local); the second stage is rare and
      // runs through the generic shared-memory code after spilling the iterates
      bool stage2 = false;
      if (!prim_ok) {
        double nv[1] = {0.0}, lhs[1] = {0.0};
        if (isr) {
          const double lpl = w.lp[lane], usmax = OSQP_INFTY * w.Es[lane];
          const double d1 = proj_dy(w.dyp[lane], lpl, hi), d2 = proj_dy(w.dys[lane], 0.0, usmax);
          nv[0] = fmax(fabs(w.Ep[lane] * d1), fabs(w.Es[lane] * d2));
          lhs[0] = kd * (hi * fmax(d1, 0.0) + lpl * fmin(d1, 0.0)) + usmax * fmax(d2, 0.0);
        }
        if (isv) {
          const double d3 = proj_dy(w.dyb[lane], lo, hi);
          nv[0] = fabs(w.Eb[lane] * d3);
          lhs[0] = hi * fmax(d3, 0.0) + lo * fmin(d3, 0.0);
        }
        Team<64>::reduce_max(nv, w.red);
        Team<64>::reduce_sum(lhs, w.red);
        if (nv[0] > st.eps_prim_inf && lhs[0] < -st.eps_prim_inf * nv[0]) stage2 = true;
      }
      if (!dual_ok && !stage2) {
        double nv[1] = {0.0}, qd[1] = {0.0};
        if (isv) { nv[0] = fabs(w.D[lane] * w.dxv[lane]); qd[0] = u0 * w.dxv[lane]; }
        if (isr) { nv[0] = fabs(w.Ds[lane] * w.dss[lane]); qd[0] = u0 * w.dss[lane]; }
        Team<64>::reduce_max(nv, w.red);
        Team<64>::reduce_sum(qd, w.red);
        if (nv[0] > st.eps_dual_inf && qd[0] < -c * st.eps_dual_inf * nv[0]) stage2 = true;
      }
      if (stage2) {
        if (isv) { w.x[lane] = p0; w.zb[lane] = z0; w.yb[lane] = y0; }
        if (isr) { w.s[lane] = s; w.zs[lane] = zs; w.ys[lane] = ys; w.zp[lane] = z0; w.yp[lane] = y0; }
        __syncthreads();
        bool pinf = false, dinf = false;
        if (!prim_ok) pinf = primal_infeasible(st.eps_prim_inf);
        if (!dual_ok) dinf = dual_infeasible(st.eps_dual_inf);
        if (pinf) { status = -3; break; }
        if (dinf) { status = -4; break; }
      }
#ifdef SCO_TIMING
      res.cyc_check += clock64() - tc0;
#endif
    }
    __syncthreads();
  }
  if (iter > max_iter) iter = max_iter;
  // ---- hand the iterates back to the shared-memory arrays of the generic epilogue
  __syncthreads();
  if (isv) { w.x[lane] = p0; w.zb[lane] = z0; w.yb[lane] = y0; }
  if (isr) { w.s[lane] = s; w.zs[lane] = zs; w.ys[lane] = ys; w.zp[lane] = z0; w.yp[lane] = y0; }
  if (!checked) {
    if (isr) { w.dss[lane] = s - prev[128 + tid]; w.dys[lane] = ys - prev[192 + tid]; w.dyp[lane] = y0 - prev[64 + tid]; }
    if (isv) { w.dxv[lane] = p0 - prev[tid]; w.dyb[lane] = y0 - prev[64 + tid]; }
  }
  __syncthreads();
  iter_out = iter;
  checked_out = checked;
  return status;
}
