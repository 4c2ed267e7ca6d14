// admm_iter.cu -- design-space microbenchmark for the register-resident dense ADMM iteration
// (n = 20 variables, m = 30 hinge rows: the (n+m) x (n+m) iteration matrix of sco_qp_dense.inl).
// Variant 0: 2 warps / problem, one full matrix row (50 doubles) per lane          (round-1 design)
// Variant 1: 4 warps / problem, half a row (26 doubles) per lane, pair-combined with one shuffle
// Variant 2: as 1, but only the even lane of a pair runs the update (odd lanes idle in the update)
// Reports cycles / iteration of CTA 0 (latency) and iterations / s over the whole chip (throughput).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o admm_iter admm_iter.cu && ./admm_iter
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NP = 20, MP = 30, NV = NP + MP;
__device__ __forceinline__ double clampd(double v, double lo, double hi) { v = v < lo ? lo : v; return v > hi ? hi : v; }

struct Consts { double sigma, alpha, rho, kd; };

// ---------------------------------------------------------------- variant 0
__device__ __forceinline__ double relu_bits(double v) {  // max(v, 0) through the sign bit
  const int hi = __double2hiint(v), m = ~(hi >> 31);
  return __hiloint2double(hi & m, __double2loint(v) & m);
}
template <int FAST>
__device__ __forceinline__ void v0_body(const double *M, double *out, long long *cyc, int iters, Consts cs);
__global__ void __maxnreg__(255) k_v0(const double *M, double *out, long long *cyc, int iters, Consts cs) { v0_body<0>(M, out, cyc, iters, cs); }
__global__ void __maxnreg__(168) k_v0_168(const double *M, double *out, long long *cyc, int iters, Consts cs) { v0_body<0>(M, out, cyc, iters, cs); }
__global__ void __maxnreg__(255) k_v0f(const double *M, double *out, long long *cyc, int iters, Consts cs) { v0_body<1>(M, out, cyc, iters, cs); }
__global__ void __maxnreg__(168) k_v0f_168(const double *M, double *out, long long *cyc, int iters, Consts cs) { v0_body<1>(M, out, cyc, iters, cs); }
template <int FAST>
__device__ __forceinline__ void v0_body(const double *M, double *out, long long *cyc, int iters, Consts cs) {
  __shared__ __align__(16) double vb[2][NV + 2];
  const int tid = threadIdx.x, lane = tid & 31;
  const bool rowwarp = tid < 32;
  const bool act = rowwarp ? lane < MP : lane < NP;
  const int slot = rowwarp ? NP + lane : lane;
  double Mr[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) Mr[k] = act ? M[(size_t)slot * NV + k] : 0.0;
  const double sigma = cs.sigma, alpha = cs.alpha, oma = 1.0 - cs.alpha, rho = cs.rho, rhoi = 1.0 / cs.rho, kd = cs.kd;
  double u0 = 0.01 * (lane + 1), u1 = rowwarp ? -1.0 : 1.0, u2 = 1.0, u3 = rowwarp ? 0.3 : 1.0 / rho;
  double lo = rowwarp ? kd * u1 : -0.5, hi = 0.5, Mi = 0.7;
  double p0 = 0, z0 = 0, y0 = 0, s = 0, zs = 0, ys = 0, g = Mi * (-u0);
  if (tid < NV + 2) { vb[0][tid] = 0.0; vb[1][tid] = 0.0; }
  __syncthreads();
  if (act) vb[0][slot] = -u0;
  __syncthreads();
  int p = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const double2 *vc = reinterpret_cast<const double2 *>(vb[p]);
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int k = 0; k < NV / 4; k++) {
      const double2 va = vc[2 * k], vd = vc[2 * k + 1];
      a0 = fma(Mr[4 * k], va.x, a0); a1 = fma(Mr[4 * k + 1], va.y, a1);
      a2 = fma(Mr[4 * k + 2], vd.x, a2); a3 = fma(Mr[4 * k + 3], vd.y, a3);
    }
    { const double2 va = vc[NV / 2 - 1]; a0 = fma(Mr[NV - 2], va.x, a0); a1 = fma(Mr[NV - 1], va.y, a1); }
    const double acc = (a0 + a1) + (a2 + a3);
    double o;
    if (FAST && rowwarp) {
      // re-associated: y' = rho (t - z'), rho z' - y' = rho (2 z' - t); exact-arithmetic identities
      const double stil = g - u3 * acc;
      const double sn = alpha * stil + oma * s;
      const double ts = (alpha * u2) * stil + (oma * zs + ys * rhoi);
      const double zns = relu_bits(ts);
      ys = rho * (ts - zns);
      const double zt = acc + u1 * stil;
      const double tz = alpha * zt + (oma * z0 + y0 * rhoi);
      const double zn = tz < hi ? tz : hi;
      y0 = rho * (tz - zn);
      s = sn; zs = zns; z0 = zn;
      const double wpen = rho * (2.0 * zn - tz);
      const double r1 = (sigma * sn - u0) + u2 * (rho * (2.0 * zns - ts)) + lo * wpen;
      g = Mi * r1;
      o = kd * wpen - (rho * lo) * g;
    } else if (FAST) {
      const double xn = alpha * acc + oma * p0;
      const double tz = (alpha * u1) * acc + (oma * z0 + y0 * u3);
      const double zn = clampd(tz, lo, hi);
      y0 = rho * (tz - zn);
      p0 = xn; z0 = zn;
      o = (sigma * xn - u0) + (u1 * rho) * (2.0 * zn - tz);
    } else if (rowwarp) {
      const double stil = g - u3 * acc, zt = acc + u1 * stil, sn = alpha * stil + oma * s;
      const double vs = alpha * (u2 * stil) + oma * zs, zns = fmax(vs + ys * rhoi, 0.0);
      ys += rho * (vs - zns);
      const double vv = alpha * zt + oma * z0, zn = fmin(vv + y0 * rhoi, hi);
      y0 += rho * (vv - zn);
      s = sn; zs = zns; z0 = zn;
      const double wpen = rho * zn - y0;
      const double r1 = sigma * sn - u0 + lo * wpen + u2 * (rho * zns - ys);
      g = Mi * r1;
      o = kd * wpen - rho * (lo * g);
    } else {
      const double xn = alpha * acc + oma * p0, vv = alpha * (u1 * acc) + oma * z0;
      const double zn = clampd(vv + y0 * u3, lo, hi);
      y0 += rho * (vv - zn);
      p0 = xn; z0 = zn;
      o = sigma * xn - u0 + u1 * (rho * zn - y0);
    }
    p ^= 1;
    if (act) vb[p][slot] = o;
    __syncthreads();
  }
  const long long t1 = clock64();
  out[blockIdx.x * 64 + tid] = p0 + z0 + y0 + s + zs + ys + g;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---------------------------------------------------------------- variants 1 / 2
// 128 threads: warps 0,1 hold the rows (15 rows x 2 halves each), warps 2,3 the variables (10 x 2 halves).
constexpr int NH = 26;  // half-row length (even), v padded to 2*NH = 52
template <int EVEN_ONLY>
__global__ void __maxnreg__(128)
k_v1(const double *M, double *out, long long *cyc, int iters, Consts cs) {
  __shared__ __align__(16) double vb[2][2 * NH];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool rowwarp = warp < 2;
  const int h = lane & 1, li = lane >> 1;
  const int idx = rowwarp ? warp * 15 + li : (warp - 2) * 10 + li;       // row / variable index
  const bool act = rowwarp ? li < 15 : li < 10;
  const int slot = rowwarp ? NP + idx : idx;
  double Mr[NH];
#pragma unroll
  for (int k = 0; k < NH; k++) {
    const int col = h * NH + k;
    Mr[k] = (act && col < NV) ? M[(size_t)slot * NV + col] : 0.0;
  }
  const double sigma = cs.sigma, alpha = cs.alpha, oma = 1.0 - cs.alpha, rho = cs.rho, rhoi = 1.0 / cs.rho, kd = cs.kd;
  double u0 = 0.01 * (idx + 1), u1 = rowwarp ? -1.0 : 1.0, u2 = 1.0, u3 = rowwarp ? 0.3 : 1.0 / rho;
  double lo = rowwarp ? kd * u1 : -0.5, hi = 0.5, Mi = 0.7;
  double p0 = 0, z0 = 0, y0 = 0, s = 0, zs = 0, ys = 0, g = Mi * (-u0);
  if (tid < 2 * NH) { vb[0][tid] = 0.0; vb[1][tid] = 0.0; }
  __syncthreads();
  if (act && h == 0) vb[0][slot] = -u0;
  __syncthreads();
  int p = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const double2 *vc = reinterpret_cast<const double2 *>(vb[p] + h * NH);
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int k = 0; k < NH / 4; k++) {
      const double2 va = vc[2 * k], vd = vc[2 * k + 1];
      a0 = fma(Mr[4 * k], va.x, a0); a1 = fma(Mr[4 * k + 1], va.y, a1);
      a2 = fma(Mr[4 * k + 2], vd.x, a2); a3 = fma(Mr[4 * k + 3], vd.y, a3);
    }
    { const double2 va = vc[NH / 2 - 1]; a0 = fma(Mr[NH - 2], va.x, a0); a1 = fma(Mr[NH - 1], va.y, a1); }
    double acc = (a0 + a1) + (a2 + a3);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    double o = 0.0;
    if (!EVEN_ONLY || h == 0) {
      if (rowwarp) {
        const double stil = g - u3 * acc, zt = acc + u1 * stil, sn = alpha * stil + oma * s;
        const double vs = alpha * (u2 * stil) + oma * zs, zns = fmax(vs + ys * rhoi, 0.0);
        ys += rho * (vs - zns);
        const double vv = alpha * zt + oma * z0, zn = fmin(vv + y0 * rhoi, hi);
        y0 += rho * (vv - zn);
        s = sn; zs = zns; z0 = zn;
        const double wpen = rho * zn - y0;
        const double r1 = sigma * sn - u0 + lo * wpen + u2 * (rho * zns - ys);
        g = Mi * r1;
        o = kd * wpen - rho * (lo * g);
      } else {
        const double xn = alpha * acc + oma * p0, vv = alpha * (u1 * acc) + oma * z0;
        const double zn = clampd(vv + y0 * u3, lo, hi);
        y0 += rho * (vv - zn);
        p0 = xn; z0 = zn;
        o = sigma * xn - u0 + u1 * (rho * zn - y0);
      }
    }
    p ^= 1;
    if (act && h == 0) vb[p][slot] = o;
    __syncthreads();
  }
  const long long t1 = clock64();
  out[blockIdx.x * 128 + tid] = p0 + z0 + y0 + s + zs + ys + g;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K>
void run(const char *name, K kern, int threads, const double *dM, double *out, long long *cyc, Consts cs) {
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0);
  const int iters = 20000;
  for (int per_sm = 1; per_sm <= occ; per_sm++) {
    const int grid = 148 * per_sm;
    kern<<<grid, threads>>>(dM, out, cyc, 100, cs);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<grid, threads>>>(dM, out, cyc, iters, cs);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s CTAs/SM %d (max %d): %7.1f cycles/iter (CTA 0)  | chip %.3f G problem-iters/s | %s\n", name, per_sm, occ,
           (double)h / iters, (double)grid * iters / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  double *hM = new double[NV * NV];
  unsigned s = 12345;
  for (int i = 0; i < NV * NV; i++) { s = s * 1664525u + 1013904223u; hM[i] = 0.02 * ((double)(s >> 8) / (1 << 24) - 0.5); }
  double *dM, *out; long long *cyc;
  cudaMalloc(&dM, NV * NV * 8); cudaMalloc(&out, 148 * 16 * 128 * 8); cudaMalloc(&cyc, 148 * 16 * 8);
  cudaMemcpy(dM, hM, NV * NV * 8, cudaMemcpyHostToDevice);
  Consts cs = {5e-10, 1.6, 0.1, 3.0};
  run("v0 2 warps full rows", k_v0, 64, dM, out, cyc, cs);
  run("v0 maxnreg 168", k_v0_168, 64, dM, out, cyc, cs);
  run("v0 fast update", k_v0f, 64, dM, out, cyc, cs);
  run("v0 fast update maxnreg 168", k_v0f_168, 64, dM, out, cyc, cs);
  return 0;
}
