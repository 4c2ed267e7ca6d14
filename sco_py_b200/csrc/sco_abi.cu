// sco_abi.cu -- host side of the C ABI declared in include/sco_b200.h: the structure compiler
// (index arrays, shared-memory layout, team size) and the launch wrappers.  The kernels live in
// sco_kernels.cuh and are instantiated per team size by sco_team.cu.
// Build: python -m sco_py_b200.build  (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "sco_b200.h"
#include "sco_device.cuh"
#include "sco_launch.h"

// ------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return fail(SCO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                                    \
  } while (0)

extern "C" const char *sco_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------ handle
struct sco_handle {
  int device = 0;
  int team = 32;
  const TeamOps *ops = nullptr;    // solve / qp (dense variant when the structure qualifies)
  const TeamOps *gen_ops = nullptr; // convexify / merit (generic team kernels)
  int carve_ops = 100, carve_gen = 100;  // shared-memory carve-out of this handle (kernel attributes are shared)
  int sm_count = 0;
  int occupancy = 1;
  size_t smem_bytes = 0;
  DevStruct S;
  std::vector<void *> dev_allocs;
  // Launch slots: every launch owns a work-queue counter and a Jacobian scratch area until it ends,
  // so launches of one handle may be in flight on several streams at once (batches pipelined over
  // streams hide the tail of a launch, where a few long problems keep a handful of SMs busy).
  static const int NSLOT = SCO_LAUNCH_SLOTS;
  double *Jscr[NSLOT] = {};
  size_t Jscr_ctas = 0;
  unsigned long long *counter = nullptr;  // NSLOT counters
  int *order_err = nullptr;               // NSLOT flags (sco_solve_batch_ordered)
  unsigned *seen[NSLOT] = {};             // bitmap of the order check, grow-only
  long long seen_cap[NSLOT] = {};
  cudaEvent_t slot_done[NSLOT] = {};
  unsigned next_slot = 0;
  // host-entry staging buffers, one set per staging slot (grow-only)
  static const int NSTG = SCO_STAGING_SETS;
  struct Staging {
    double *params = nullptr, *x0 = nullptr, *x = nullptr, *merit = nullptr, *obj = nullptr, *vio = nullptr;
    int *verdict = nullptr, *stats = nullptr, *nonconv = nullptr;
    long long cap_B = 0;
    cudaEvent_t done = nullptr;
  } stg[NSTG];
  unsigned next_stg = 0;
};

template <typename T>
static int upload(sco_handle *h, const std::vector<T> &v, const T **out) {
  T *d = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  CUDA_TRY(cudaMalloc(&d, bytes));
  h->dev_allocs.push_back(d);
  if (!v.empty()) CUDA_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}

static DevField cvt(const sco_field &f) {
  DevField d;
  d.off = f.off;
  d.shared = f.shared;
  d.pad_ = 0;
  return d;
}

static void build_layout(DevStruct &S, int team) {
  Layout &L = S.L;
  // dense kinds: the two-warp solve owns a compile-time region at the start of dynamic shared memory
  static const int dense_words[5] = {0, DenseL<8, 6>::total, DenseL<12, 16>::total, DenseL<20, 30>::total,
                                     DenseL<32, 32>::total};
  int off = dense_words[S.dense_kind];
  auto take = [&](int cnt) {
    int o = off;
    off += (cnt + 1) & ~1;
    return o;
  };
  const int n = S.n, ml = S.m_lin, mp = S.m_nl, sl = S.nsl * S.m_nl;
  L.Js = take(S.sjnnz); L.S = take(S.sws ? 0 : n * n); L.Als = take(S.nnz_lin);
  L.x = take(n); L.xt = take(n); L.xt2 = take(n + 4); L.qh = take(n); L.D = take(n); L.bx = take(n);  // xt2, wp: + 4 = zero padding
  L.rb = take(n); L.lb = take(n); L.ub = take(n); L.zb = take(n); L.yb = take(n); L.Eb = take(n);
  L.dxv = take(n); L.dyb = take(n); L.xs = take(n);
  L.El = take(ml); L.rl = take(ml); L.ll = take(ml); L.ul = take(ml); L.zl = take(ml); L.yl = take(ml);
  L.wl = take(ml); L.dyl = take(ml);
  L.Ep = take(mp); L.rp = take(mp); L.lp = take(mp); L.up = take(mp); L.zp = take(mp); L.yp = take(mp);
  L.wp = take(mp + 4); L.bb = take(mp); L.fv = take(mp); L.dyp = take(mp);
  L.s = take(sl); L.Ds = take(sl); L.sl = take(sl); L.bs = take(sl); L.zs = take(sl); L.ys = take(sl);
  L.Es = take(sl); L.gs = take(sl); L.hs = take(sl); L.rs = take(sl); L.dss = take(sl); L.dys = take(sl);
  L.Minv = take(3 * mp);
  L.red = take((std::max(team / 32, 8) + 1) * 16);  // one 16-double slot per warp + one for the combined results
  L.msk = take((mp * S.mw + 1) / 2);
  L.stage = take(std::max((team / 32) * S.stage_per_warp, 100));
  L.Hq = take(S.obj_len ? n * n + 2 : 0); L.gq = take(S.obj_len ? n : 0);
  L.ps = take(4 * n);
  L.total = off;
}

extern "C" void sco_default_settings(sco_settings *s) {
  memset(s, 0, sizeof(*s));
  s->improve_ratio_threshold = 0.25;  // solver.py:17-28
  s->min_trust_region_size = 1e-4;
  s->min_approx_improve = 1e-8;
  s->trust_shrink_ratio = 0.1;
  s->trust_expand_ratio = 1.5;
  s->cnt_tolerance = 1e-4;
  s->merit_coeff_increase_ratio = 10.0;
  s->initial_trust_region_size = 1.0;
  s->initial_penalty_coeff = 1e3;
  s->max_merit_coeff_increases = 1;
  s->max_sqp_iters = 100000;
  s->osqp_eps_abs = 1e-6;  // osqp_utils.py:10-15
  s->osqp_eps_rel = 1e-9;
  s->osqp_rho = 0.1;
  s->osqp_sigma = 5e-10;
  s->osqp_alpha = 1.6;  // OSQP defaults, SURVEY.md Appendix B
  s->osqp_eps_prim_inf = 1e-4;
  s->osqp_eps_dual_inf = 1e-4;
  s->osqp_max_iter = 100000;
  s->osqp_scaling = 10;
  s->osqp_check_termination = 25;
  s->osqp_adaptive_rho = 0;
  s->osqp_adaptive_rho_interval = 0;
  s->compound_penalty = 1;
  s->freeze_sparsity = 1;
  s->duplicate_rows = 1;
  s->threads_per_problem = 0;
  s->force_generic = 0;
  s->aff_obj_quirk = 1;
  s->warm_start = 0;
}

static DevSettings to_dev(const sco_settings *s) {
  DevSettings d;
  d.improve_ratio_threshold = s->improve_ratio_threshold;
  d.min_trust_region_size = s->min_trust_region_size;
  d.min_approx_improve = s->min_approx_improve;
  d.trust_shrink_ratio = s->trust_shrink_ratio;
  d.trust_expand_ratio = s->trust_expand_ratio;
  d.cnt_tolerance = s->cnt_tolerance;
  d.merit_coeff_increase_ratio = s->merit_coeff_increase_ratio;
  d.initial_trust_region_size = s->initial_trust_region_size;
  d.initial_penalty_coeff = s->initial_penalty_coeff;
  d.max_merit_coeff_increases = s->max_merit_coeff_increases;
  d.max_sqp_iters = s->max_sqp_iters;
  d.eps_abs = s->osqp_eps_abs;
  d.eps_rel = s->osqp_eps_rel;
  d.rho = s->osqp_rho;
  d.sigma = s->osqp_sigma;
  d.alpha = s->osqp_alpha;
  d.eps_prim_inf = s->osqp_eps_prim_inf;
  d.eps_dual_inf = s->osqp_eps_dual_inf;
  d.max_iter = s->osqp_max_iter;
  d.scaling = s->osqp_scaling;
  d.check_termination = s->osqp_check_termination;
  d.adaptive_rho = s->osqp_adaptive_rho;
  d.adaptive_rho_interval = s->osqp_adaptive_rho_interval;
  d.compound_penalty = s->compound_penalty;
  d.freeze_sparsity = s->freeze_sparsity;
  d.duplicate_rows = s->duplicate_rows;
  d.force_generic = s->force_generic;
  d.aff_obj_quirk = s->aff_obj_quirk;
  d.warm_start = s->warm_start;
  return d;
}

extern "C" int sco_create(const sco_structure_desc *desc, int device, sco_handle **out) {
  if (!desc || !out) return fail(SCO_ERR_ARG, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(SCO_ERR_CUDA, "no CUDA device: the engine has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(SCO_ERR_ARG, "device %d out of range", device);
  if (desc->n <= 0 || desc->n_blocks < 0 || desc->n_blocks > SCO_MAX_BLOCKS)
    return fail(SCO_ERR_ARG, "bad n / n_blocks");
  // n_groups = 0: no constraint belongs to a group -- the reference then skips the per-group test (solver.py:209)
  if (desc->n_groups < 0 || desc->n_groups > SCO_MAX_GROUPS) return fail(SCO_ERR_ARG, "bad n_groups");
  if (desc->m_lin < 0 || desc->stride <= 0 || desc->shared_len < 0) return fail(SCO_ERR_ARG, "negative m_lin / stride / shared_len");
  if (desc->shared_len > 0 && !desc->shared) return fail(SCO_ERR_ARG, "shared_len > 0 but shared is null");
  {
    // every field must lie inside the block it lives in (out-of-range offsets would be device reads out of bounds)
    const long long n_ = desc->n;
    auto field_ok = [&](const sco_field &f, long long len) {
      if (f.off < 0) return true;
      const long long lim = f.shared ? (long long)desc->shared_len : (long long)desc->stride;
      return f.off + len <= lim;
    };
    bool okf = field_ok(desc->Q, n_ * n_) && field_ok(desc->q, n_) && field_ok(desc->c, 1) &&
               field_ok(desc->lin_l, desc->m_lin) && field_ok(desc->lin_u, desc->m_lin) && field_ok(desc->qa, n_) &&
               field_ok(desc->lb0, n_) && field_ok(desc->ub0, n_) &&
               (desc->obj_prog_len <= 0 || field_ok(desc->obj_prog, 1 + 2LL * desc->obj_prog_len));
    for (int bi = 0; bi < desc->n_blocks && okf; bi++) {
      const sco_block_desc &b = desc->blocks[bi];
      if (b.m <= 0) return fail(SCO_ERR_ARG, "block %d: m must be positive", bi);
      if (desc->n_groups < 32 && (b.group_mask >> desc->n_groups) != 0)
        return fail(SCO_ERR_ARG, "block %d: group_mask names a group >= n_groups", bi);
      long long plen = 0;
      if (b.family == SCO_FAM_QUADFORM) plen = (long long)b.m * (n_ * (n_ + 1) / 2 + n_);
      else if (b.family == SCO_FAM_CIRCLE2D) plen = 3LL * b.ipar[1];
      else if (b.family == SCO_FAM_FK7) plen = 30;
      else if (b.family == SCO_FAM_VM) plen = b.m + 2LL * b.ipar[2];
      if (b.par.off < 0) return fail(SCO_ERR_ARG, "block %d: parameters are missing", bi);
      okf = field_ok(b.par, plen) && field_ok(b.val, b.m);
    }
    if (!okf) return fail(SCO_ERR_ARG, "a field lies outside its parameter block (offset + length > stride / shared_len)");
    if (desc->m_lin > 0 && (desc->lin_l.off < 0 || desc->lin_u.off < 0)) return fail(SCO_ERR_ARG, "m_lin > 0 needs lin_l and lin_u");
  }
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(SCO_ERR_CUDA, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);

  sco_handle *h = new sco_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  DevStruct &S = h->S;
  memset(&S, 0, sizeof(S));
  const int n = desc->n;
  S.n = n;
  S.m_lin = desc->m_lin;
  S.n_blocks = desc->n_blocks;
  S.n_groups = desc->n_groups;
  S.stride = desc->stride;
  S.Q = cvt(desc->Q); S.q = cvt(desc->q); S.c = cvt(desc->c);
  S.lin_l = cvt(desc->lin_l); S.lin_u = cvt(desc->lin_u);
  S.qa = cvt(desc->qa); S.lb0 = cvt(desc->lb0); S.ub0 = cvt(desc->ub0);
  S.objp = cvt(desc->obj_prog); S.obj_len = desc->obj_prog_len > 0 && desc->obj_prog.off >= 0 ? desc->obj_prog_len : 0;
  S.obj_flags = desc->obj_prog_flags;
  if (S.obj_len && n > 64) { delete h; return fail(SCO_ERR_UNSUPPORTED, "non-quadratic objectives are limited to 64 variables"); }
  for (int g = 0; g < desc->n_groups; g++) {
    int bits = 0;
    if (desc->group_overlap)
      for (int g2 = 0; g2 < desc->n_groups; g2++)
        if (desc->group_overlap[g * desc->n_groups + g2]) bits |= 1 << g2;
    S.overlap[g] = bits;
  }
  // ---- penalty rows
  std::vector<int> row_goff, row_soff, row_w, row_eq, row_gmask, jcol;
  int m_nl = 0, jnnz = 0, sj = 0, n_slack = 0, any_eq = 0, max_stage = 0;
  int max_jw = 1;
  for (int bi = 0; bi < desc->n_blocks; bi++) {
    const sco_block_desc &b = desc->blocks[bi];
    DevBlock &d = S.blocks[bi];
    d.family = b.family; d.cnt_type = b.cnt_type; d.m = b.m; d.group_mask = b.group_mask;
    d.row0 = m_nl; d.joff = jnnz;
    memcpy(d.ipar, b.ipar, sizeof(d.ipar));
    d.par = cvt(b.par); d.val = cvt(b.val);
    int jw = 0;
    if (b.family == SCO_FAM_QUADFORM) { jw = n; max_stage = std::max(max_stage, n * (n + 1) / 2); }
    else if (b.family == SCO_FAM_CIRCLE2D) {
      jw = 2;
      if (b.ipar[0] * b.ipar[1] != b.m || 2 * b.ipar[0] > n) { delete h; return fail(SCO_ERR_ARG, "CIRCLE2D: m != T*K or 2T > n"); }
    } else if (b.family == SCO_FAM_FK7) {
      jw = 7;
      if (b.m != 3 || n < 7) { delete h; return fail(SCO_ERR_ARG, "FK7: m must be 3 and n >= 7"); }
    } else if (b.family == SCO_FAM_VM) {
      jw = n;
      if (b.ipar[0] != n || b.ipar[1] != b.m || b.ipar[2] <= 0) { delete h; return fail(SCO_ERR_ARG, "VM: ipar must be {n, m, instructions}"); }
    } else { delete h; return fail(SCO_ERR_UNSUPPORTED, "unknown constraint family %d", b.family); }
    max_jw = std::max(max_jw, jw);
    d.jw = jw;
    const int ldj = (jw % 2 == 0) ? jw + 1 : jw;
    for (int r = 0; r < b.m; r++) {
      row_goff.push_back(jnnz + r * jw);
      row_soff.push_back(sj + r * ldj);
      row_w.push_back(jw);
      row_eq.push_back(b.cnt_type == SCO_CNT_EQ ? 1 : 0);
      row_gmask.push_back(b.group_mask);
      for (int k = 0; k < jw; k++) {
        int col = 0;
        if (b.family == SCO_FAM_QUADFORM || b.family == SCO_FAM_VM) col = k;
        else if (b.family == SCO_FAM_CIRCLE2D) col = 2 * (r / b.ipar[1]) + k;
        else col = n - 7 + k;
        jcol.push_back(col);
      }
    }
    if (b.cnt_type == SCO_CNT_EQ) any_eq = 1;
    n_slack += b.m * (b.cnt_type == SCO_CNT_EQ ? 2 : 1);
    m_nl += b.m;
    jnnz += b.m * jw;
    sj += b.m * ldj;
  }
  S.m_nl = m_nl; S.jnnz = jnnz; S.sjnnz = sj; S.n_slack = n_slack; S.n_q = n + n_slack;
  S.nsl = any_eq ? 2 : 1;
  S.stage_per_warp = (max_stage + 1) & ~1;
  S.mw = (max_jw + 31) / 32;
  // CSC over user variables of the stored Jacobian pattern
  std::vector<int> pc_ptr(n + 1, 0), pc_e, pc_r;
  {
    std::vector<std::vector<std::pair<int, int>>> cols(n);
    for (int i = 0; i < m_nl; i++)
      for (int k = 0; k < row_w[i]; k++) cols[jcol[row_goff[i] + k]].push_back({row_soff[i] + k, i});
    for (int j = 0; j < n; j++) {
      pc_ptr[j + 1] = pc_ptr[j] + (int)cols[j].size();
      for (auto &pr : cols[j]) { pc_e.push_back(pr.first); pc_r.push_back(pr.second); }
    }
  }
  // linear rows
  std::vector<int> lrp(desc->m_lin + 1, 0), lcol, lcptr(n + 1, 0), lcentry, lcrow;
  std::vector<double> lval;
  if (desc->m_lin > 0) {
    if (!desc->lin_rowptr || !desc->lin_col || !desc->lin_val) { delete h; return fail(SCO_ERR_ARG, "m_lin > 0 but CSR arrays are null"); }
    lrp.assign(desc->lin_rowptr, desc->lin_rowptr + desc->m_lin + 1);
    const int nnz = lrp[desc->m_lin];
    lcol.assign(desc->lin_col, desc->lin_col + nnz);
    lval.assign(desc->lin_val, desc->lin_val + nnz);
    std::vector<std::vector<std::pair<int, int>>> cols(n);
    for (int r = 0; r < desc->m_lin; r++)
      for (int p = lrp[r]; p < lrp[r + 1]; p++) {
        if (lcol[p] < 0 || lcol[p] >= n) { delete h; return fail(SCO_ERR_ARG, "linear row column out of range"); }
        cols[lcol[p]].push_back({p, r});
      }
    for (int j = 0; j < n; j++) {
      lcptr[j + 1] = lcptr[j] + (int)cols[j].size();
      for (auto &pr : cols[j]) { lcentry.push_back(pr.first); lcrow.push_back(pr.second); }
    }
    S.nnz_lin = nnz;
  }
  std::vector<double> shared_v;
  if (desc->shared_len > 0 && desc->shared) shared_v.assign(desc->shared, desc->shared + desc->shared_len);
  // pattern of Psym = (Q + Q') / 2 by column: exact for a shared Q, dense for a per-problem Q, empty without Q
  std::vector<int> P_cptr(n + 1, 0), P_row;
  if (desc->Q.off >= 0) {
    for (int j = 0; j < n; j++) {
      for (int i = 0; i < n; i++) {
        bool nz = true;
        if (desc->Q.shared) nz = shared_v[desc->Q.off + (size_t)i * n + j] != 0.0 || shared_v[desc->Q.off + (size_t)j * n + i] != 0.0;
        if (nz) P_row.push_back(i);
      }
      P_cptr[j + 1] = (int)P_row.size();
    }
  }
  int rc = 0;
  rc |= upload(h, row_goff, &S.row_goff); rc |= upload(h, row_soff, &S.row_soff);
  rc |= upload(h, row_w, &S.row_w); rc |= upload(h, row_eq, &S.row_eq);
  rc |= upload(h, row_gmask, &S.row_gmask); rc |= upload(h, jcol, &S.jcol_g);
  rc |= upload(h, pc_ptr, &S.pc_ptr); rc |= upload(h, pc_e, &S.pc_e); rc |= upload(h, pc_r, &S.pc_r);
  rc |= upload(h, lrp, &S.lin_rowptr); rc |= upload(h, lcol, &S.lin_col); rc |= upload(h, lval, &S.lin_val);
  rc |= upload(h, lcptr, &S.lin_cptr); rc |= upload(h, lcentry, &S.lin_centry); rc |= upload(h, lcrow, &S.lin_crow);
  rc |= upload(h, shared_v, &S.shared);
  rc |= upload(h, P_cptr, &S.P_cptr); rc |= upload(h, P_row, &S.P_row);
  {
    int bw = 0;
    for (int j = 0; j < n; j++)
      for (int p = P_cptr[j]; p < P_cptr[j + 1]; p++) bw = std::max(bw, std::abs(P_row[p] - j));
    for (int r = 0; r < desc->m_lin; r++)
      if (lrp[r + 1] > lrp[r]) {
        int lo = n, hi = -1;
        for (int p = lrp[r]; p < lrp[r + 1]; p++) { lo = std::min(lo, lcol[p]); hi = std::max(hi, lcol[p]); }
        bw = std::max(bw, hi - lo);
      }
    for (int i = 0; i < m_nl; i++)
      if (row_w[i] > 0) {
        int lo = n, hi = -1;
        for (int k = 0; k < row_w[i]; k++) { lo = std::min(lo, jcol[row_goff[i] + k]); hi = std::max(hi, jcol[row_goff[i] + k]); }
        bw = std::max(bw, hi - lo);
      }
    if (S.obj_len) bw = n - 1;  // the degree-2 model of a non-quadratic objective is dense
    S.s_bw = std::min(bw, n - 1);
    // thread-per-entity ADMM loop (sco_qp.cuh: fast_role keeps SCO_EN = 8 row entries / SCO_EH = 4 + 4 column entries
    // in registers and has no path for more)
    bool ok = true;
    for (int r = 0; r < desc->m_lin; r++) ok = ok && lrp[r + 1] - lrp[r] <= 8;
    for (int i = 0; i < m_nl; i++) ok = ok && row_w[i] <= 8;
    for (int j = 0; j < n; j++) ok = ok && lcptr[j + 1] - lcptr[j] <= 4 && pc_ptr[j + 1] - pc_ptr[j] <= 4;
    S.fast_ok = ok ? 1 : 0;
    bool narrow = true;
    for (int j = 0; j < n; j++) narrow = narrow && P_cptr[j + 1] - P_cptr[j] <= 4;
    S.p_narrow = narrow ? 1 : 0;
    bool dense = desc->m_lin == 0 && n <= 32 && m_nl > 0 && m_nl <= 48;
    for (int i = 0; i < m_nl; i++) dense = dense && row_w[i] == n;
    S.fast_dense = dense && !ok ? 1 : 0;
  }
  if (rc) { sco_destroy(h); return SCO_ERR_CUDA; }
  // ---- dense fast path: one dense hinge block, no linear rows, sizes within the instantiated table
  S.dense_kind = 0;
  if (desc->m_lin == 0 && S.obj_len == 0 && desc->n_blocks == 1 && desc->blocks[0].family == SCO_FAM_QUADFORM &&
      desc->blocks[0].cnt_type == SCO_CNT_LEQ) {
    static const int table[][2] = {{8, 6}, {12, 16}, {20, 30}, {32, 32}};  // keep in sync with QPSolver::solve in sco_qp.cuh
    for (int k = 0; k < 4; k++)
      if (n <= table[k][0] && m_nl <= table[k][1]) { S.dense_kind = k + 1; break; }
  }
  // ---- team size and shared-memory layout
  const int work = std::max(std::max(n, m_nl), desc->m_lin);
  int team = work <= 40 ? 32 : work <= 96 ? 64 : work <= 192 ? 128 : 256;
  {
    // one thread per entity (variable / linear row / penalty row) is what the fast ADMM loop wants (sco_qp.cuh:
    // fast_loop); structures with more than 512 entities stay on the strided shared-memory loop
    // (every role starts on a warp boundary)
    const int entities = ((n + 31) & ~31) + ((desc->m_lin + 31) & ~31) + ((m_nl + 31) & ~31);
    if (entities <= 512) team = entities <= 32 ? 32 : entities <= 64 ? 64 : entities <= 128 ? 128 : entities <= 256 ? 256 : 512;
  }
  if (S.dense_kind) team = 64;  // two warps per problem: rows | variables (sco_dense.cuh)
  for (;;) {
    build_layout(S, team);
    h->smem_bytes = (size_t)(S.L.total + ((n + 1) & ~1)) * sizeof(double);
    if (h->smem_bytes <= (size_t)prop.sharedMemPerBlockOptin) break;
    if (!S.sws && !S.dense_kind && !S.obj_len) {
      // the n x n matrix S (and its inverse) moves to a global workspace per resident team: the ADMM loop streams it
      // from L2 / HBM in every iteration (strided loop only); everything else stays in shared memory
      S.sws = n * n;
      continue;
    }
    sco_destroy(h);
    return fail(SCO_ERR_UNSUPPORTED, "problem working set (%zu B) exceeds shared memory per block (%zu B)",
                h->smem_bytes, (size_t)prop.sharedMemPerBlockOptin);
  }
  h->team = team;
  h->gen_ops = team == 32 ? sco_team_ops_32() : team == 64 ? sco_team_ops_64() : team == 128 ? sco_team_ops_128()
               : team == 256 ? sco_team_ops_256() : sco_team_ops_512();
  h->ops = h->gen_ops;
  if (S.dense_kind)
    h->ops = S.dense_kind == 1 ? sco_dense_ops_1() : S.dense_kind == 2 ? sco_dense_ops_2() : S.dense_kind == 3 ? sco_dense_ops_3() : sco_dense_ops_4();
  {
    int occ = 0;
    cudaError_t ce = h->gen_ops->configure(h->smem_bytes, &occ, &h->carve_gen);
    h->carve_ops = h->carve_gen;
    if (ce == cudaSuccess && h->ops != h->gen_ops) ce = h->ops->configure(h->smem_bytes, &occ, &h->carve_ops);
    if (ce != cudaSuccess) {
      sco_destroy(h);
      return fail(SCO_ERR_CUDA, "kernel configuration failed: %s", cudaGetErrorString(ce));
    }
    h->occupancy = std::max(occ, 1);
  }
  h->Jscr_ctas = (size_t)h->sm_count * h->occupancy;
  bool ok = cudaMalloc(&h->counter, sco_handle::NSLOT * sizeof(unsigned long long)) == cudaSuccess &&
            cudaMalloc(&h->order_err, sco_handle::NSLOT * sizeof(int)) == cudaSuccess;
  for (int k = 0; k < sco_handle::NSLOT && ok; k++)
    ok = cudaMalloc(&h->Jscr[k], std::max<size_t>(h->Jscr_ctas * (size_t)std::max(jnnz + S.sws, 1), 1) * sizeof(double)) == cudaSuccess &&
         cudaEventCreateWithFlags(&h->slot_done[k], cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; k < sco_handle::NSTG && ok; k++)
    ok = cudaEventCreateWithFlags(&h->stg[k].done, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    sco_destroy(h);
    return fail(SCO_ERR_CUDA, "cudaMalloc of scratch failed");
  }
  *out = h;
  return SCO_OK;
}

extern "C" int sco_destroy(sco_handle *h) {
  if (!h) return SCO_OK;
  cudaSetDevice(h->device);
  for (void *p : h->dev_allocs) cudaFree(p);
  cudaFree(h->counter);
  cudaFree(h->order_err);
  for (int k = 0; k < sco_handle::NSLOT; k++) {
    cudaFree(h->Jscr[k]);
    cudaFree(h->seen[k]);
    if (h->slot_done[k]) cudaEventDestroy(h->slot_done[k]);
  }
  for (int k = 0; k < sco_handle::NSTG; k++) {
    sco_handle::Staging &g = h->stg[k];
    cudaFree(g.params); cudaFree(g.x0); cudaFree(g.x); cudaFree(g.merit); cudaFree(g.obj); cudaFree(g.vio);
    cudaFree(g.verdict); cudaFree(g.stats); cudaFree(g.nonconv);
    if (g.done) cudaEventDestroy(g.done);
  }
  delete h;
  return SCO_OK;
}

extern "C" int sco_query(sco_handle *h, int64_t *out8) {
  if (!h || !out8) return fail(SCO_ERR_ARG, "null argument");
  out8[0] = h->S.n; out8[1] = h->S.m_nl; out8[2] = h->S.n_slack; out8[3] = h->S.jnnz;
  out8[4] = h->S.n_q; out8[5] = (int64_t)h->smem_bytes; out8[6] = h->team; out8[7] = h->occupancy;
  return SCO_OK;
}

// d_order must be a permutation of 0..B-1: every entry in range and seen once.  Violations are counted in *err;
// k_solve reads it and refuses the batch (verdict -3 everywhere) -- stream-ordered, no host synchronisation.
__global__ void k_check_order(const int *__restrict__ order, long long B, unsigned *seen, int *err) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) {
    const int v = order[i];
    if (v < 0 || v >= B) { atomicAdd(err, 1); continue; }
    const unsigned bit = 1u << (v & 31);
    if (atomicOr(seen + (v >> 5), bit) & bit) atomicAdd(err, 1);
  }
}

extern "C" int sco_solve_batch(sco_handle *h, int64_t B, const double *d_params, const double *d_x0,
                               const sco_settings *s, double *d_x_out, int32_t *d_verdict,
                               double *d_merit, double *d_objective, double *d_max_vio,
                               int32_t *d_stats, void *stream) {
  return sco_solve_batch_ordered(h, B, d_params, d_x0, s, d_x_out, d_verdict, d_merit, d_objective, d_max_vio,
                                 d_stats, nullptr, stream);
}

extern "C" int sco_solve_batch_ordered(sco_handle *h, int64_t B, const double *d_params, const double *d_x0,
                                       const sco_settings *s, double *d_x_out, int32_t *d_verdict,
                                       double *d_merit, double *d_objective, double *d_max_vio,
                                       int32_t *d_stats, const int32_t *d_order, void *stream) {
  sco_batch_io io;
  memset(&io, 0, sizeof(io));
  io.d_params = d_params; io.d_x0 = d_x0; io.d_x_out = d_x_out; io.d_verdict = d_verdict; io.d_merit = d_merit;
  io.d_objective = d_objective; io.d_max_vio = d_max_vio; io.d_stats = d_stats; io.d_order = d_order;
  return sco_solve_batch_io(h, B, &io, s, stream);
}

extern "C" int sco_solve_batch_io(sco_handle *h, int64_t B, const sco_batch_io *io, const sco_settings *s, void *stream) {
  if (!h || !s || !io) return fail(SCO_ERR_ARG, "null argument");
  if (B <= 0) return SCO_OK;  // an empty batch is valid and touches nothing
  if (!io->d_params || !io->d_x0 || !io->d_x_out || !io->d_verdict) return fail(SCO_ERR_ARG, "null argument");
  if (io->d_x_warm || io->d_y_warm) return fail(SCO_ERR_UNSUPPORTED, "x_warm / y_warm are reserved");
  if (B > 0x7fffffffLL) return fail(SCO_ERR_ARG, "batch too large");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  DevSettings d = to_dev(s);
  const int slot = (int)(h->next_slot++ % sco_handle::NSLOT);
  CUDA_TRY(cudaStreamWaitEvent(st, h->slot_done[slot], 0));  // the slot's previous launch (any stream)
  CUDA_TRY(cudaMemsetAsync(h->counter + slot, 0, sizeof(unsigned long long), st));
  const int *order_err = nullptr;
  if (io->d_order) {
    const long long words = (B + 31) / 32;
    if (words > h->seen_cap[slot]) {
      CUDA_TRY(cudaEventSynchronize(h->slot_done[slot]));
      if (h->seen[slot]) cudaFree(h->seen[slot]);
      h->seen[slot] = nullptr;
      h->seen_cap[slot] = 0;
      CUDA_TRY(cudaMalloc(&h->seen[slot], (size_t)words * sizeof(unsigned)));
      h->seen_cap[slot] = words;
    }
    CUDA_TRY(cudaMemsetAsync(h->seen[slot], 0, (size_t)words * sizeof(unsigned), st));
    CUDA_TRY(cudaMemsetAsync(h->order_err + slot, 0, sizeof(int), st));
    const int blocks = (int)std::min<long long>((B + 255) / 256, 4096);
    k_check_order<<<blocks, 256, 0, st>>>(io->d_order, (long long)B, h->seen[slot], h->order_err + slot);
    CUDA_TRY(cudaGetLastError());
    order_err = h->order_err + slot;
  }
  const long long grid = std::min<long long>(B, (long long)h->Jscr_ctas);
  SolveArgs a = {(long long)B, io->d_params, io->d_x0, io->d_x_out, io->d_verdict, io->d_merit, io->d_objective,
                 io->d_max_vio, io->d_stats, h->Jscr[slot], h->counter + slot, io->d_order, io->d_nonconverged, order_err};
  CUDA_TRY(h->ops->prepare(h->carve_ops));
  h->ops->solve((unsigned)grid, h->smem_bytes, st, h->S, d, a);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(h->slot_done[slot], st));
  return SCO_OK;
}

template <typename T>
static int grow(T **p, size_t count) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  CUDA_TRY(cudaMalloc(p, std::max<size_t>(count, 1) * sizeof(T)));
  return 0;
}

extern "C" int sco_solve_batch_host_async(sco_handle *h, int64_t B, const double *params, const double *x0,
                                          const sco_settings *s, double *x_out, int32_t *verdict,
                                          double *merit, double *objective, double *max_vio,
                                          int32_t *stats, void *stream) {
  return sco_solve_batch_host_groups(h, B, params, x0, s, x_out, verdict, merit, objective, max_vio, stats, nullptr,
                                     stream);
}

extern "C" int sco_solve_batch_host_groups(sco_handle *h, int64_t B, const double *params, const double *x0,
                                           const sco_settings *s, double *x_out, int32_t *verdict,
                                           double *merit, double *objective, double *max_vio,
                                           int32_t *stats, int32_t *nonconverged, void *stream) {
  if (!h || !s) return fail(SCO_ERR_ARG, "null argument");
  if (B <= 0) return SCO_OK;
  if (!params || !x0 || !x_out || !verdict) return fail(SCO_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  const int n = h->S.n;
  cudaStream_t st = (cudaStream_t)stream;
  sco_handle::Staging &g = h->stg[h->next_stg++ % sco_handle::NSTG];
  if (B > g.cap_B) {
    CUDA_TRY(cudaEventSynchronize(g.done));  // nothing in flight may still use the old buffers
    int rc = 0;
    rc |= grow(&g.params, (size_t)B * h->S.stride); rc |= grow(&g.x0, (size_t)B * n);
    rc |= grow(&g.x, (size_t)B * n); rc |= grow(&g.merit, (size_t)B); rc |= grow(&g.obj, (size_t)B);
    rc |= grow(&g.vio, (size_t)B); rc |= grow(&g.verdict, (size_t)B); rc |= grow(&g.stats, (size_t)4 * B);
    rc |= grow(&g.nonconv, (size_t)B);
    if (rc) { g.cap_B = 0; return SCO_ERR_CUDA; }
    g.cap_B = B;
  }
  CUDA_TRY(cudaStreamWaitEvent(st, g.done, 0));
  CUDA_TRY(cudaMemcpyAsync(g.params, params, (size_t)B * h->S.stride * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(g.x0, x0, (size_t)B * n * sizeof(double), cudaMemcpyHostToDevice, st));
  sco_batch_io io;
  memset(&io, 0, sizeof(io));
  io.d_params = g.params; io.d_x0 = g.x0; io.d_x_out = g.x; io.d_verdict = g.verdict; io.d_merit = g.merit;
  io.d_objective = g.obj; io.d_max_vio = g.vio; io.d_stats = g.stats; io.d_nonconverged = nonconverged ? g.nonconv : nullptr;
  int rc = sco_solve_batch_io(h, B, &io, s, st);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(x_out, g.x, (size_t)B * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(verdict, g.verdict, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (merit) CUDA_TRY(cudaMemcpyAsync(merit, g.merit, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (objective) CUDA_TRY(cudaMemcpyAsync(objective, g.obj, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (max_vio) CUDA_TRY(cudaMemcpyAsync(max_vio, g.vio, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (stats) CUDA_TRY(cudaMemcpyAsync(stats, g.stats, (size_t)4 * B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (nonconverged) CUDA_TRY(cudaMemcpyAsync(nonconverged, g.nonconv, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(g.done, st));
  return SCO_OK;
}

extern "C" int sco_solve_batch_host(sco_handle *h, int64_t B, const double *params, const double *x0,
                                    const sco_settings *s, double *x_out, int32_t *verdict,
                                    double *merit, double *objective, double *max_vio,
                                    int32_t *stats) {
  int rc = sco_solve_batch_host_async(h, B, params, x0, s, x_out, verdict, merit, objective, max_vio, stats, nullptr);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(0));
  return SCO_OK;
}

static unsigned stage_grid(sco_handle *h, int64_t B) {
  return (unsigned)std::min<long long>(B, (long long)h->Jscr_ctas);
}

extern "C" int sco_convexify(sco_handle *h, int64_t B, const double *d_params, const double *d_x,
                             double *d_f, double *d_J, double *d_b, double *d_obj, void *stream) {
  return sco_convexify_model(h, B, d_params, d_x, d_f, d_J, d_b, d_obj, nullptr, nullptr, nullptr, stream);
}

extern "C" int sco_convexify_model(sco_handle *h, int64_t B, const double *d_params, const double *d_x,
                                   double *d_f, double *d_J, double *d_b, double *d_obj, double *d_H, double *d_g,
                                   double *d_c, void *stream) {
  if (!h || !d_params || !d_x) return fail(SCO_ERR_ARG, "null argument");
  if ((d_H || d_g || d_c) && !h->S.obj_len) return fail(SCO_ERR_ARG, "the structure has no non-quadratic objective term");
  if (B <= 0) return SCO_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  // without d_J the kernel parks the Jacobian in a launch slot's scratch: take the slot like a solve does
  double *scratch = nullptr;
  int slot = -1;
  if (!d_J) {
    slot = (int)(h->next_slot++ % sco_handle::NSLOT);
    CUDA_TRY(cudaStreamWaitEvent(st, h->slot_done[slot], 0));
    scratch = h->Jscr[slot];
  }
  ConvexifyArgs a = {(long long)B, d_params, d_x, d_f, d_J, d_b, d_obj, scratch, d_H, d_g, d_c};
  CUDA_TRY(h->gen_ops->prepare(h->carve_gen));
  h->gen_ops->convexify(stage_grid(h, B), h->smem_bytes, st, h->S, a);
  CUDA_TRY(cudaGetLastError());
  if (slot >= 0) CUDA_TRY(cudaEventRecord(h->slot_done[slot], st));
  return SCO_OK;
}

extern "C" int sco_qp_solve(sco_handle *h, int64_t B, const double *d_params, const double *d_J,
                            const double *d_b, const uint32_t *d_mask, const double *d_lbx,
                            const double *d_ubx, const double *d_pi, const int32_t *d_kdup,
                            const double *d_xref, int use_penalty, int closest_point,
                            const sco_settings *s, double *d_xq, int32_t *d_status, int32_t *d_iters,
                            void *stream) {
  return sco_qp_solve_w(h, B, d_params, d_J, d_b, d_mask, d_lbx, d_ubx, d_pi, d_kdup, nullptr, d_xref, use_penalty,
                        closest_point, s, d_xq, d_status, d_iters, stream);
}

extern "C" int sco_qp_solve_w(sco_handle *h, int64_t B, const double *d_params, const double *d_J,
                              const double *d_b, const uint32_t *d_mask, const double *d_lbx,
                              const double *d_ubx, const double *d_pi, const int32_t *d_kdup,
                              const double *d_wa, const double *d_xref, int use_penalty, int closest_point,
                              const sco_settings *s, double *d_xq, int32_t *d_status, int32_t *d_iters,
                              void *stream) {
  if (!h || !s || !d_params || !d_xq || !d_status || !d_iters) return fail(SCO_ERR_ARG, "null argument");
  if (use_penalty && h->S.m_nl > 0 && (!d_J || !d_b)) return fail(SCO_ERR_ARG, "penalty rows need J and b");
  if (closest_point && !d_xref) return fail(SCO_ERR_ARG, "closest_point needs xref");
  if (B <= 0) return SCO_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  DevSettings d = to_dev(s);
  QpStageArgs a = {(long long)B, d_params, d_J, d_b, d_mask, d_lbx, d_ubx, d_pi, d_kdup, d_wa, d_xref,
                   use_penalty, closest_point, d_xq, d_status, d_iters, nullptr};
  int slot = -1;
  if (h->S.sws) {  // S lives in a launch slot's scratch: take the slot like a solve does
    slot = (int)(h->next_slot++ % sco_handle::NSLOT);
    CUDA_TRY(cudaStreamWaitEvent(st, h->slot_done[slot], 0));
    a.scr = h->Jscr[slot];
  }
  CUDA_TRY(h->ops->prepare(h->carve_ops));
  h->ops->qp(stage_grid(h, B), h->smem_bytes, st, h->S, d, a);
  CUDA_TRY(cudaGetLastError());
  if (slot >= 0) CUDA_TRY(cudaEventRecord(h->slot_done[slot], st));
  return SCO_OK;
}

extern "C" int sco_merit(sco_handle *h, int64_t B, const double *d_params, const double *d_x,
                         const double *d_J, const double *d_b, const double *d_mu, double *d_merit,
                         double *d_model, double *d_max_vio, double *d_gv, double *d_gm, void *stream) {
  if (!h || !d_params || !d_x) return fail(SCO_ERR_ARG, "null argument");
  if (B <= 0) return SCO_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  MeritArgs a = {(long long)B, d_params, d_x, d_J, d_b, d_mu, d_merit, d_model, d_max_vio, d_gv, d_gm};
  CUDA_TRY(h->gen_ops->prepare(h->carve_gen));
  h->gen_ops->merit(stage_grid(h, B), h->smem_bytes, st, h->S, a);
  CUDA_TRY(cudaGetLastError());
  return SCO_OK;
}
