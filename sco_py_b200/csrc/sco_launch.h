// sco_launch.h -- host-side launch table: one TeamOps per team size, defined by sco_team.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sco_device.cuh"

struct SolveArgs {
  long long B;
  const double *params, *x0;
  double *x_out;
  int *verdict;
  double *merit, *objective, *max_vio;
  int *stats;
  double *Jscr;
  unsigned long long *counter;
  const int *order;
  int *nonconv;           // [B] or null
  const int *order_err;   // != 0 on the device: `order` is not a permutation -> every verdict = -3, nothing solved
};
struct ConvexifyArgs {
  long long B;
  const double *params, *x;
  double *f, *J, *b, *obj, *Jscr;
  double *Hq, *gq, *cq;  // degree-2 model of a non-quadratic objective term: H+ [B,n,n], linear term [B,n], constant [B]
};
struct QpStageArgs {
  long long B;
  const double *params, *J, *b;
  const uint32_t *mask;
  const double *lbx, *ubx, *pi;
  const int *kdup;
  const double *wa;
  const double *xref;
  int use_pen, closest;
  double *xq;
  int *status, *iters;
  double *scr;  // a launch slot's scratch (DevStruct::sws: S of every team lives there), else null
};
struct MeritArgs {
  long long B;
  const double *params, *x, *J, *b, *mu;
  double *merit, *model, *max_vio, *gv, *gm;
};

struct TeamOps {
  int team;
  cudaError_t (*configure)(size_t smem_bytes, int *occupancy, int *carveout);
  cudaError_t (*prepare)(int carveout);  // re-applies the handle's carve-out before a launch (kernel attributes are shared)
  void (*solve)(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const DevSettings &d,
                const SolveArgs &a);
  void (*convexify)(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const ConvexifyArgs &a);
  void (*qp)(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const DevSettings &d,
             const QpStageArgs &a);
  void (*merit)(unsigned grid, size_t smem, cudaStream_t st, const DevStruct &S, const MeritArgs &a);
};

const TeamOps *sco_team_ops_32();
const TeamOps *sco_team_ops_64();
const TeamOps *sco_team_ops_128();
const TeamOps *sco_team_ops_256();
const TeamOps *sco_team_ops_512();
// dense kinds (two-warp register-resident ADMM loop, sco_dense.cuh); convexify / merit are null
const TeamOps *sco_dense_ops_1();
const TeamOps *sco_dense_ops_2();
const TeamOps *sco_dense_ops_3();
const TeamOps *sco_dense_ops_4();
