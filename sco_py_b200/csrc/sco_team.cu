// sco_team.cu -- instantiates every kernel for ONE team size (-DSCO_TEAM=32|64|128|256).
#include "sco_kernels.cuh"
#include "sco_launch.h"

#ifndef SCO_TEAM
#error "compile with -DSCO_TEAM=<threads per problem> [-DSCO_DK=<dense kind>]"
#endif
#ifndef SCO_DK
#define SCO_DK 0
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

namespace {
constexpr int T = SCO_TEAM;
constexpr int DK = SCO_DK;

// The dynamic-shared-memory limit and the carve-out are attributes of the KERNEL, shared by every
// handle of this team size.  The limit therefore only ever grows (a later handle with a smaller
// working set must not lower it under an earlier one), and the carve-out a handle wants is re-applied
// by `prepare` right before each of its launches.
constexpr int MAX_DEV = 64;  // attributes are per device (context)
size_t g_max_bytes_dev[MAX_DEV] = {};
int g_carve_dev[MAX_DEV] = {};
bool g_carve_set[MAX_DEV] = {};

int cur_dev() {
  int d = 0;
  cudaGetDevice(&d);
  return d < 0 || d >= MAX_DEV ? 0 : d;
}

cudaError_t raise_limit(size_t bytes) {
  size_t &g_max_bytes = g_max_bytes_dev[cur_dev()];
  if (bytes <= g_max_bytes) return cudaSuccess;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k_solve<T, DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e;
  if ((e = cudaFuncSetAttribute(k_qp<T, DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e;
#if SCO_DK == 0
  if ((e = cudaFuncSetAttribute(k_convexify<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e;
  if ((e = cudaFuncSetAttribute(k_merit<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e;
#endif
  g_max_bytes = bytes;
  return cudaSuccess;
}

cudaError_t prepare(int carve) {
  const int dev = cur_dev();
  int &g_carve = g_carve_dev[dev];
  if (g_carve_set[dev] && carve == g_carve) return cudaSuccess;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k_solve<T, DK>, cudaFuncAttributePreferredSharedMemoryCarveout, carve))) return e;
  if ((e = cudaFuncSetAttribute(k_qp<T, DK>, cudaFuncAttributePreferredSharedMemoryCarveout, carve))) return e;
#if SCO_DK == 0
  if ((e = cudaFuncSetAttribute(k_convexify<T>, cudaFuncAttributePreferredSharedMemoryCarveout, carve))) return e;
  if ((e = cudaFuncSetAttribute(k_merit<T>, cudaFuncAttributePreferredSharedMemoryCarveout, carve))) return e;
#endif
  g_carve = carve;
  g_carve_set[dev] = true;
  return cudaSuccess;
}

// -> resident CTAs per SM for a working set of `bytes`, and the carve-out (percent, or
// cudaSharedmemCarveoutMaxShared) this handle passes to `prepare` before its launches.
cudaError_t configure(size_t bytes, int *occ, int *carve) {
  cudaError_t e;
  if ((e = raise_limit(bytes))) return e;
  // Shared-memory carve-out: first find the occupancy with the largest carve-out, then ask only for
  // what that many CTAs need.  The rest of the 256 KB stays L1 cache: the kernels keep their stack
  // frames and the structure's index arrays there, and with the maximal carve-out (28 KB of L1 left
  // for 8+ warps) every such access was an L2 round trip.
  if ((e = prepare(cudaSharedmemCarveoutMaxShared))) return e;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_solve<T, DK>, T, bytes))) return e;
  const size_t need = (size_t)(*occ > 0 ? *occ : 1) * (bytes + 1024);  // 1 KB per CTA is reserved by the driver
  int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
  if (pct > 100) pct = 100;
  if ((e = prepare(pct))) return e;
  int occ2 = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_solve<T, DK>, T, bytes))) return e;
  *carve = occ2 < *occ ? (int)cudaSharedmemCarveoutMaxShared : pct;  // the hint cost occupancy: maximal carve-out
  return prepare(*carve);
}
void solve(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const DevSettings &d,
           const SolveArgs &a) {
  k_solve<T, DK><<<grid, T, smem, st>>>(S, d, a.B, a.params, a.x0, a.x_out, a.verdict, a.merit, a.objective,
                                    a.max_vio, a.stats, a.Jscr, a.counter, a.order, a.nonconv, a.order_err);
}
#if SCO_DK == 0
void convexify(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const ConvexifyArgs &a) {
  k_convexify<T><<<grid, T, smem, st>>>(S, a.B, a.params, a.x, a.f, a.J, a.b, a.obj, a.Jscr, a.Hq, a.gq, a.cq);
}
#endif
void qp(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const DevSettings &d,
        const QpStageArgs &a) {
  k_qp<T, DK><<<grid, T, smem, st>>>(S, d, a.B, a.params, a.J, a.b, a.mask, a.lbx, a.ubx, a.pi, a.kdup, a.wa, a.xref,
                                 a.use_pen, a.closest, a.xq, a.status, a.iters, a.scr);
}
#if SCO_DK == 0
void merit(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const MeritArgs &a) {
  k_merit<T><<<grid, T, smem, st>>>(S, a.B, a.params, a.x, a.J, a.b, a.mu, a.merit, a.model, a.max_vio, a.gv,
                                    a.gm);
}
const TeamOps ops = {T, configure, prepare, solve, convexify, qp, merit};
#else
const TeamOps ops = {T, configure, prepare, solve, nullptr, qp, nullptr};
#endif
}  // namespace

#if SCO_DK == 0
const TeamOps *CAT(sco_team_ops_, SCO_TEAM)() { return &ops; }
#else
const TeamOps *CAT(sco_dense_ops_, SCO_DK)() { return &ops; }
#endif
