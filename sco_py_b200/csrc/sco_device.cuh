// sco_device.cuh -- device-side data model shared by every kernel of the engine.
//
// One thread TEAM (a warp, or a CTA of 2..8 warps) owns one problem at a time and keeps the
// whole working set of its current penalty QP in shared memory for the duration of the ADMM
// solve ("resident regime").  All sizes are run-time values taken from
// the structure; the shared-memory carve-up (Layout) is computed once on the host.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SCO_DEV_MAX_BLOCKS 8
#define SCO_DEV_MAX_GROUPS 8

#define OSQP_INFTY 1e30
#define OSQP_MIN_SCALING 1e-4
#define OSQP_MAX_SCALING 1e4
#define OSQP_RHO_MIN 1e-6
#define OSQP_RHO_MAX 1e6
#define OSQP_RHO_EQ_OVER_RHO_INEQ 1e3
#define OSQP_RHO_TOL 1e-4

struct DevField {
  long long off;
  int shared;
  int pad_;
};

struct DevBlock {
  int family, cnt_type, m, group_mask;
  int jw;    // stored Jacobian entries per row
  int row0;  // first penalty row of the block
  int joff;  // offset of the block inside the per-problem J array (global layout, unpadded)
  int pad_;
  int ipar[8];
  DevField par, val;
};

// shared-memory carve-up, offsets in doubles
struct Layout {
  int Js, S, Als;
  // per user variable (n)
  int x, xt, xt2, qh, D, bx, rb, lb, ub, zb, yb, Eb, dxv, dyb, xs;
  // per linear row (m_lin)
  int El, rl, ll, ul, zl, yl, wl, dyl;
  // per penalty row (m_nl)
  int Ep, rp, lp, up, zp, yp, wp, bb, fv, dyp;
  // per slack (nsl*m_nl, slot a*m_nl + i)
  int s, Ds, sl, bs, zs, ys, Es, gs, hs, rs, dss, dys;
  int Minv;   // 3*m_nl
  int red;    // reduction scratch: 8 warps * 8 values
  int msk;    // m_nl uint32 (counted in doubles, rounded up)
  int stage;  // per-warp staging buffers for family evaluation
  int total;  // doubles
};

struct DevStruct {
  int n, m_lin, nnz_lin, n_blocks, n_groups, m_nl, n_slack, jnnz, n_q, nsl;
  int sjnnz;           // padded Jacobian entries in shared memory
  int stage_per_warp;  // doubles
  long long stride;
  DevField Q, q, c, lin_l, lin_u;
  // linear rows: CSR + CSC (entry index into lin_val / Als)
  const int *lin_rowptr, *lin_col, *lin_cptr, *lin_centry, *lin_crow;
  const double *lin_val;
  // penalty rows
  const int *row_goff;   // m_nl: offset of the row in the global J layout
  const int *row_soff;   // m_nl: offset of the row in the shared-memory (padded) layout
  const int *row_w;      // m_nl: entries in the row
  const int *row_eq;     // m_nl: 1 = equality (abs penalty, two slacks)
  const int *row_gmask;  // m_nl: constraint-group membership
  const int *jcol_g;     // jnnz: column of every stored entry (global layout)
  const int *pc_ptr, *pc_e, *pc_r;  // CSC over user variables: shared-memory entry index, row
  const double *shared;
  int overlap[SCO_DEV_MAX_GROUPS];  // bit g2 of overlap[g] <=> groups overlap (prob.py:139-142)
  DevBlock blocks[SCO_DEV_MAX_BLOCKS];
  Layout L;
};

struct DevSettings {
  double improve_ratio_threshold, min_trust_region_size, min_approx_improve;
  double trust_shrink_ratio, trust_expand_ratio, cnt_tolerance, merit_coeff_increase_ratio;
  double initial_trust_region_size, initial_penalty_coeff;
  int max_merit_coeff_increases, max_sqp_iters;
  double eps_abs, eps_rel, rho, sigma, alpha, eps_prim_inf, eps_dual_inf;
  int max_iter, scaling, check_termination, adaptive_rho, adaptive_rho_interval;
  int compound_penalty, freeze_sparsity, duplicate_rows;
};

__device__ __forceinline__ const double *field_ptr(const DevStruct &S, const DevField &f,
                                                   const double *prm) {
  return f.off < 0 ? nullptr : ((f.shared ? S.shared : prm) + f.off);
}

// ------------------------------------------------------------------------------------
// team primitives
template <int TEAM>
struct Team {
  static __device__ __forceinline__ void sync() {
    if (TEAM == 32) __syncwarp();
    else __syncthreads();
  }

  template <int K, bool IS_MAX>
  static __device__ __forceinline__ void reduce(double (&v)[K], double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        double t = __shfl_xor_sync(0xffffffffu, v[k], o);
        v[k] = IS_MAX ? fmax(v[k], t) : v[k] + t;
      }
    }
    if (TEAM > 32) {
      const int wi = threadIdx.x >> 5, l = threadIdx.x & 31;
      __syncthreads();
      if (l == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) red[wi * 16 + k] = v[k];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < K; k++) {
        double acc = red[k];
        for (int ww = 1; ww < TEAM / 32; ww++)
          acc = IS_MAX ? fmax(acc, red[ww * 16 + k]) : acc + red[ww * 16 + k];
        v[k] = acc;
      }
    }
  }
  template <int K>
  static __device__ __forceinline__ void reduce_max(double (&v)[K], double *red) {
    reduce<K, true>(v, red);
  }
  template <int K>
  static __device__ __forceinline__ void reduce_sum(double (&v)[K], double *red) {
    reduce<K, false>(v, red);
  }
};

__device__ __forceinline__ double limit_scaling(double v) {
  v = v < OSQP_MIN_SCALING ? 1.0 : v;
  v = v > OSQP_MAX_SCALING ? OSQP_MAX_SCALING : v;
  return v;
}

// shared-memory views
struct QPW {
  double *Js, *Sm, *Als;
  double *x, *xt, *xt2, *qh, *D, *bx, *rb, *lb, *ub, *zb, *yb, *Eb, *dxv, *dyb, *xs;
  double *El, *rl, *ll, *ul, *zl, *yl, *wl, *dyl;
  double *Ep, *rp, *lp, *up, *zp, *yp, *wp, *bb, *fv, *dyp;
  double *s, *Ds, *sl, *bs, *zs, *ys, *Es, *gs, *hs, *rs, *dss, *dys, *Minv;
  double *red;
  uint32_t *msk;
  double *stage;

  __device__ __forceinline__ void bind(double *sm, const Layout &L) {
    Js = sm + L.Js; Sm = sm + L.S; Als = sm + L.Als;
    x = sm + L.x; xt = sm + L.xt; xt2 = sm + L.xt2; qh = sm + L.qh; D = sm + L.D; bx = sm + L.bx;
    rb = sm + L.rb; lb = sm + L.lb; ub = sm + L.ub; zb = sm + L.zb; yb = sm + L.yb; Eb = sm + L.Eb;
    dxv = sm + L.dxv; dyb = sm + L.dyb; xs = sm + L.xs;
    El = sm + L.El; rl = sm + L.rl; ll = sm + L.ll; ul = sm + L.ul; zl = sm + L.zl; yl = sm + L.yl;
    wl = sm + L.wl; dyl = sm + L.dyl;
    Ep = sm + L.Ep; rp = sm + L.rp; lp = sm + L.lp; up = sm + L.up; zp = sm + L.zp; yp = sm + L.yp;
    wp = sm + L.wp; bb = sm + L.bb; fv = sm + L.fv; dyp = sm + L.dyp;
    s = sm + L.s; Ds = sm + L.Ds; sl = sm + L.sl; bs = sm + L.bs; zs = sm + L.zs; ys = sm + L.ys;
    Es = sm + L.Es; gs = sm + L.gs; hs = sm + L.hs; rs = sm + L.rs; dss = sm + L.dss; dys = sm + L.dys;
    Minv = sm + L.Minv;
    red = sm + L.red;
    msk = reinterpret_cast<uint32_t *>(sm + L.msk);
    stage = sm + L.stage;
  }
};
