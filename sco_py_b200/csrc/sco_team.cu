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

cudaError_t configure(size_t bytes, int *occ) {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(k_solve<T, DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e;
  if ((e = cudaFuncSetAttribute(k_qp<T, DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e;
#if SCO_DK == 0
  if ((e = cudaFuncSetAttribute(k_convexify<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e;
  if ((e = cudaFuncSetAttribute(k_merit<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes))) return e;
#endif
  // Shared-memory carve-out: first find the occupancy with the largest carve-out, then ask only for
  // what that many CTAs need.  The rest of the 256 KB stays L1 cache: the kernels keep their stack
  // frames and the structure's index arrays there, and with the maximal carve-out (28 KB of L1 left
  // for 8+ warps) every such access was an L2 round trip.
  if ((e = cudaFuncSetAttribute(k_solve<T, DK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_solve<T, DK>, T, bytes))) return e;
  {
    const size_t need = (size_t)(*occ > 0 ? *occ : 1) * (bytes + 1024);  // 1 KB per CTA is reserved by the driver
    int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    if ((e = cudaFuncSetAttribute(k_solve<T, DK>, cudaFuncAttributePreferredSharedMemoryCarveout, pct))) return e;
    if ((e = cudaFuncSetAttribute(k_qp<T, DK>, cudaFuncAttributePreferredSharedMemoryCarveout, pct))) return e;
#if SCO_DK == 0
    if ((e = cudaFuncSetAttribute(k_convexify<T>, cudaFuncAttributePreferredSharedMemoryCarveout, pct))) return e;
    if ((e = cudaFuncSetAttribute(k_merit<T>, cudaFuncAttributePreferredSharedMemoryCarveout, pct))) return e;
#endif
    int occ2 = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_solve<T, DK>, T, bytes))) return e;
    if (occ2 < *occ) {  // the hint cost occupancy: go back to the maximal carve-out
      if ((e = cudaFuncSetAttribute(k_solve<T, DK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
      if ((e = cudaFuncSetAttribute(k_qp<T, DK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
    }
  }
  return cudaSuccess;
}
void solve(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const DevSettings &d,
           const SolveArgs &a) {
  k_solve<T, DK><<<grid, T, smem, st>>>(S, d, a.B, a.params, a.x0, a.x_out, a.verdict, a.merit, a.objective,
                                    a.max_vio, a.stats, a.Jscr, a.counter, a.order);
}
#if SCO_DK == 0
void convexify(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const ConvexifyArgs &a) {
  k_convexify<T><<<grid, T, smem, st>>>(S, a.B, a.params, a.x, a.f, a.J, a.b, a.obj, a.Jscr);
}
#endif
void qp(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const DevSettings &d,
        const QpStageArgs &a) {
  k_qp<T, DK><<<grid, T, smem, st>>>(S, d, a.B, a.params, a.J, a.b, a.mask, a.lbx, a.ubx, a.pi, a.kdup, a.xref,
                                 a.use_pen, a.closest, a.xq, a.status, a.iters);
}
#if SCO_DK == 0
void merit(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const MeritArgs &a) {
  k_merit<T><<<grid, T, smem, st>>>(S, a.B, a.params, a.x, a.J, a.b, a.mu, a.merit, a.model, a.max_vio, a.gv,
                                    a.gm);
}
const TeamOps ops = {T, configure, solve, convexify, qp, merit};
#else
const TeamOps ops = {T, configure, solve, nullptr, qp, nullptr};
#endif
}  // namespace

#if SCO_DK == 0
const TeamOps *CAT(sco_team_ops_, SCO_TEAM)() { return &ops; }
#else
const TeamOps *CAT(sco_dense_ops_, SCO_DK)() { return &ops; }
#endif
