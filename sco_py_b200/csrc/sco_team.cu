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
  // the working sets are sized against the full 227 KB of shared memory per SM
  if ((e = cudaFuncSetAttribute(k_solve<T, DK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
  if ((e = cudaFuncSetAttribute(k_qp<T, DK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_solve<T, DK>, T, bytes);
}
void solve(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const DevSettings &d,
           const SolveArgs &a) {
  k_solve<T, DK><<<grid, T, smem, st>>>(S, d, a.B, a.params, a.x0, a.x_out, a.verdict, a.merit, a.objective,
                                    a.max_vio, a.stats, a.Jscr, a.counter);
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
