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
  // Reciprocals of the scaling vectors for the termination test (r0, r1, r2):
  //   variable lane: 1/Eb, 1/D, -        row lane: 1/Ep, 1/Es, 1/Ds       (0 on inactive lanes)
  const double rho_u = rho, rhoi_u = 1.0 / rho, cinv = 1.0 / c;
  double u0, u1, u2, u3, lo, hi, Mi = 0.0, r0, r1, r2 = 0.0;
  if (rowwarp) {
    u0 = isr ? cpi * w.Ds[lane] : 0.0;
    u1 = isr ? w.sl[lane] : 0.0;
    u2 = isr ? w.bs[lane] : 0.0;
    u3 = isr ? w.hs[lane] : 0.0;
    lo = kd * u1;
    hi = isr ? w.up[lane] : 0.0;
    Mi = isr ? w.Minv[3 * lane] : 0.0;
    r0 = isr ? 1.0 / w.Ep[lane] : 0.0;
    r1 = isr ? 1.0 / w.Es[lane] : 0.0;
    r2 = isr ? 1.0 / w.Ds[lane] : 0.0;
  } else {
    u0 = isv ? w.qh[lane] : 0.0;
    u1 = isv ? w.bx[lane] : 0.0;
    u2 = isv ? w.rb[lane] : 1.0;
    u3 = 1.0 / u2;
    lo = isv ? w.lb[lane] : 0.0;
    hi = isv ? w.ub[lane] : 0.0;
    r0 = isv ? 1.0 / w.Eb[lane] : 0.0;
    r1 = isv ? 1.0 / w.D[lane] : 0.0;
  }
  // ---- iterates: (p0, z0, y0) = (x, zb, yb) or (-, zp, yp); row lanes also (s, zs, ys, g)
  double p0 = 0.0, z0 = 0.0, y0 = 0.0, s = 0.0, zs = 0.0, ys = 0.0;
  double g = Mi * (-u0);
  const double wp0 = -(rho_u * lo) * g;
  const int slot = rowwarp ? NP + lane : lane;           // this lane's entry of (c, wp)
  const bool publishes = rowwarp ? lane < MP : lane < NP;
  for (int e = tid; e < 2 * FBUF_LD + 64; e += 64) vb[e] = 0.0;
  __syncthreads();
  if (publishes) vb[slot] = rowwarp ? wp0 : -u0;
  // 32-bit shared-window addresses: the loop below addresses shared memory through explicit
  // ld.shared / st.shared so that nothing but the loads, the FMAs and the update is in its body
  const uint32_t vb0 = (uint32_t)__cvta_generic_to_shared(vb.ptr());
  const uint32_t vb1 = vb0 + 8u * FBUF_LD;
  const uint32_t my_out = 8u * (uint32_t)slot;
  const uint32_t my_chk = vb0 + 8u * (2 * FBUF_LD + (rowwarp ? 32 : 0) + lane);  // xS[lane] / yS[lane]
  double pp0 = 0.0, py0 = 0.0, ps = 0.0, pys = 0.0;  // iterates before the last (checked) iteration
  int iter = 0, status = 0;
  bool checked = false;
  const int max_iter = st.max_iter, chk = st.check_termination;
  int next_check = chk ? chk : max_iter + 1;
  __syncthreads();

  // one ADMM iteration: reads (c, wp) at `vcur`, publishes this lane's new entry at `vnext`
  auto iterate = [&](const uint32_t vcur, const uint32_t vnext) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int k = 0; k < NV / 4; k++) {
      const double2 va = lds_v2(vcur + 32u * k), vc = lds_v2(vcur + 32u * k + 16u);
      a0 = fma(Mr[4 * k], va.x, a0);
      a1 = fma(Mr[4 * k + 1], va.y, a1);
      a2 = fma(Mr[4 * k + 2], vc.x, a2);
      a3 = fma(Mr[4 * k + 3], vc.y, a3);
    }
    if (NV % 4) {
      const double2 va = lds_v2(vcur + 8u * (NV - 2));
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
      const double ts = vs + ys * rhoi_u;
      const double zns = ts > 0.0 ? ts : 0.0;
      ys += rho_u * (vs - zns);
      const double vv = alpha * zt + oma * z0;
      const double tz = vv + y0 * rhoi_u;
      const double zn = tz < hi ? tz : hi;
      y0 += rho_u * (vv - zn);
      s = sn;
      zs = zns;
      z0 = zn;
      const double wpen = rho_u * zn - y0;
      const double rr = sigma * sn - u0 + lo * wpen + u2 * (rho_u * zns - ys);
      g = Mi * rr;
      out = kd * wpen - rho_u * (lo * g);
    } else {
      // x~ = acc ; x and box-row updates, then c for the next iteration
      const double xn = alpha * acc + oma * p0;
      const double vv = alpha * (u1 * acc) + oma * z0;
      const double zn = clampd(vv + y0 * u3, lo, hi);
      y0 += u2 * (vv - zn);
      p0 = xn;
      z0 = zn;
      out = sigma * xn - u0 + u1 * (u2 * zn - y0);
    }
    if (publishes) sts_f64(vnext + my_out, out);
  };

  int p = 0;
  while (iter < max_iter) {
    // ---- plain iterations up to (not including) the next checked / last one
    const int seg_end = next_check < max_iter ? next_check : max_iter;
    for (int i = iter + 1; i < seg_end; i++) {
      iterate(p ? vb1 : vb0, p ? vb0 : vb1);
      p ^= 1;
      __syncthreads();
    }
    // ---- the checked (or last) iteration: the infeasibility certificates need delta_x / delta_y
    pp0 = p0; py0 = y0; ps = s; pys = ys;
    iterate(p ? vb1 : vb0, p ? vb0 : vb1);
    p ^= 1;
    iter = seg_end;
    checked = false;
    if (iter != next_check) { __syncthreads(); break; }
    // ================= termination test (OSQP check_termination, unscaled residuals) ===========
#ifdef SCO_TIMING
    const long long tc0 = clock64();
    long long tcs = tc0;
#endif
    next_check += chk;
    checked = true;
    sts_f64(my_chk, rowwarp ? kd * y0 : p0);
    __syncthreads();
#ifdef SCO_TIMING
    { const long long tq = clock64(); res.cyc_c[0] += tq - tcs; tcs = tq; }
#endif
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // 7 residual norms, |dy| and |dx| of the certificates
    double sm2[2] = {0.0, 0.0};                  // their linear terms (sums)
    if (rowwarp) {
      // (J x)_i: compile-time trip count so the loads are issued back to back (a run-time loop makes
      // every FMA wait for its own shared-memory round trip)
      const Sh Jrow = w.Js + (isr ? lane : 0) * ldj;
      double ax0 = 0.0, ax1 = 0.0;
#pragma unroll
      for (int k = 0; k < NP; k += 2) {
        const double2 xk = xS.v2(k >> 1);
        if (k < n) ax0 = fma(Jrow[k], xk.x, ax0);
        if (k + 1 < n) ax1 = fma(Jrow[k + 1], xk.y, ax1);
      }
      const double ax = (ax0 + ax1) + u1 * s;
      v[0] = fabs((ax - z0) * r0); v[1] = fabs(z0 * r0); v[2] = fabs(ax * r0);
      const double axs = u2 * s;
      v[0] = fmax(v[0], fabs((axs - zs) * r1)); v[1] = fmax(v[1], fabs(zs * r1));
      v[2] = fmax(v[2], fabs(axs * r1));
      const double aty = lo * y0 + u2 * ys;
      v[3] = fabs((u0 + aty) * r2); v[4] = fabs(u0 * r2); v[5] = fabs(aty * r2);
      // first stage of the certificates (lane-local part)
      const double Epl = isr ? w.Ep[lane] : 0.0, Esl = isr ? w.Es[lane] : 0.0, Dsl = isr ? w.Ds[lane] : 0.0;
      const double lpl = -OSQP_INFTY * Epl, usmax = OSQP_INFTY * Esl;
      const double d1 = isr ? proj_dy(y0 - py0, lpl, hi) : 0.0, d2 = isr ? proj_dy(ys - pys, 0.0, usmax) : 0.0;
      v[7] = fmax(fabs(Epl * d1), fabs(Esl * d2));
      sm2[0] = kd * (hi * fmax(d1, 0.0) + lpl * fmin(d1, 0.0)) + usmax * fmax(d2, 0.0);
      const double dss = s - ps;
      v[8] = fabs(Dsl * dss);
      sm2[1] = u0 * dss;
    } else {
      const int col = isv ? lane : 0;
      const double ax = u1 * p0;
      v[0] = fabs((ax - z0) * r0); v[1] = fabs(z0 * r0); v[2] = fabs(ax * r0);
      const Sh Pcol = w.Ph + col, Jcol = w.Js + col;
      double px0 = 0.0, px1 = 0.0, at0 = 0.0, at1 = 0.0;
#pragma unroll
      for (int k = 0; k < NP; k += 2) {
        const double2 xk = xS.v2(k >> 1);
        if (k < n) px0 = fma(Pcol[k * n], xk.x, px0);
        if (k + 1 < n) px1 = fma(Pcol[(k + 1) * n], xk.y, px1);
      }
#pragma unroll
      for (int i = 0; i < MP; i += 2) {
        const double2 yi = yS.v2(i >> 1);
        if (i < m_nl) at0 = fma(Jcol[i * ldj], yi.x, at0);
        if (i + 1 < m_nl) at1 = fma(Jcol[(i + 1) * ldj], yi.y, at1);
      }
      const double px = px0 + px1, aty = (at0 + at1) + u1 * y0;
      v[3] = fabs((u0 + px + aty) * r1); v[4] = fabs(u0 * r1);
      v[5] = fabs(aty * r1); v[6] = fabs(px * r1);
      const double Ebl = isv ? w.Eb[lane] : 0.0, Dl = isv ? w.D[lane] : 0.0;
      const double d3 = isv ? proj_dy(y0 - py0, lo, hi) : 0.0;
      v[7] = fabs(Ebl * d3);
      sm2[0] = hi * fmax(d3, 0.0) + lo * fmin(d3, 0.0);
      const double dx = p0 - pp0;
      v[8] = fabs(Dl * dx);
      sm2[1] = u0 * dx;
    }
#ifdef SCO_TIMING
    { const long long tq = clock64(); res.cyc_c[1] += tq - tcs; tcs = tq; }
#endif
    // one combined team reduction: maxima through redux.sync on the (order-preserving) bit patterns
#pragma unroll
    for (int k = 0; k < 9; k++) v[k] = warp_max_nonneg(v[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sm2[0] += __shfl_xor_sync(0xffffffffu, sm2[0], o);
      sm2[1] += __shfl_xor_sync(0xffffffffu, sm2[1], o);
    }
#ifdef SCO_TIMING
    { const long long tq = clock64(); res.cyc_c[2] += tq - tcs; tcs = tq; }
#endif
    {
      const int wi = rowwarp ? 0 : 1;
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; k++) w.red[wi * 16 + k] = v[k];
        w.red[wi * 16 + 9] = sm2[0];
        w.red[wi * 16 + 10] = sm2[1];
      }
      __syncthreads();
      const int wo = wi ^ 1;
#pragma unroll
      for (int k = 0; k < 9; k++) v[k] = fmax(v[k], w.red[wo * 16 + k]);
      // fixed summation order (warp 0 + warp 1) so that both warps take the same decision
      sm2[0] = w.red[9] + w.red[16 + 9];
      sm2[1] = w.red[10] + w.red[16 + 10];
    }
#ifdef SCO_TIMING
    { const long long tq = clock64(); res.cyc_c[3] += tq - tcs; tcs = tq; }
#endif
    const double pri_res = v[0], dua_res = cinv * v[3];
    res.pri_res = pri_res;
    res.dua_res = dua_res;
    if (pri_res > OSQP_INFTY || dua_res > OSQP_INFTY) { status = -7; break; }
    const double eps_p = st.eps_abs + st.eps_rel * fmax(v[1], v[2]);
    const double eps_d = st.eps_abs + st.eps_rel * cinv * fmax(v[4], fmax(v[5], v[6]));
    const bool prim_ok = pri_res < eps_p, dual_ok = dua_res < eps_d;
    if (prim_ok && dual_ok) { status = 1; break; }
    // the second stage of the infeasibility certificates is rare and runs through the generic
    // shared-memory code after spilling the iterates
    bool stage2 = false;
    if (!prim_ok && v[7] > st.eps_prim_inf && sm2[0] < -st.eps_prim_inf * v[7]) stage2 = true;
    if (!dual_ok && !stage2 && v[8] > st.eps_dual_inf && sm2[1] < -c * st.eps_dual_inf * v[8]) stage2 = true;
    if (stage2) {
      if (isv) { w.x[lane] = p0; w.zb[lane] = z0; w.yb[lane] = y0; w.dxv[lane] = p0 - pp0; w.dyb[lane] = y0 - py0; }
      if (isr) {
        w.s[lane] = s; w.zs[lane] = zs; w.ys[lane] = ys; w.zp[lane] = z0; w.yp[lane] = y0;
        w.dss[lane] = s - ps; w.dys[lane] = ys - pys; w.dyp[lane] = y0 - py0;
      }
      __syncthreads();
      bool pinf = false, dinf = false;
      if (!prim_ok) pinf = primal_infeasible(st.eps_prim_inf);
      if (!dual_ok) dinf = dual_infeasible(st.eps_dual_inf);
      if (pinf) { status = -3; break; }
      if (dinf) { status = -4; break; }
    }
#ifdef SCO_TIMING
    { const long long tq = clock64(); res.cyc_c[4] += tq - tcs; res.cyc_check += tq - tc0; }
#endif
  }
  // ---- hand the iterates (and the deltas of the last iteration) back to the shared-memory arrays
  // of the generic epilogue
  __syncthreads();
  if (isv) { w.x[lane] = p0; w.zb[lane] = z0; w.yb[lane] = y0; w.dxv[lane] = p0 - pp0; w.dyb[lane] = y0 - py0; }
  if (isr) {
    w.s[lane] = s; w.zs[lane] = zs; w.ys[lane] = ys; w.zp[lane] = z0; w.yp[lane] = y0;
    w.dss[lane] = s - ps; w.dys[lane] = ys - pys; w.dyp[lane] = y0 - py0;
  }
  __syncthreads();
  iter_out = iter;
  checked_out = checked;
  return status;
}
