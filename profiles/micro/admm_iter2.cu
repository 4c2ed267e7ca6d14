// admm_iter2.cu -- round-2 design-space microbenchmark of the dense ADMM iteration (n = 20, m = 30).
//
// v0  (round 1, for reference on the same box): ONE fused mat-vec with the (n+m) x (n+m) matrix
//     [S^-1 K; K' G]: 2 warps, a full 50-entry row per lane (100 registers), 25 broadcast LDS.128 per lane,
//     255 registers -> 4 CTAs / SM.
// v3  two-phase split mat-vec: x~ = [S^-1 | K] (c, wp)  (20 x 50, three lanes per row, 18-entry slices),
//     then t = J x~ (30 x 20, two lanes per row, 10-entry slices); partial sums meet in shared memory, the
//     owner lanes (warp 0 = rows, warp 1 = variables) run the updates.  1,600 instead of 2,500 multiply-adds,
//     14 instead of 25 vector loads per lane, 28 instead of 50 matrix entries per lane -> 128 registers,
//     8 CTAs / SM; the price is four barriers per iteration instead of one.
// Reports cycles / iteration of CTA 0 (latency) and problem-iterations / s over the whole chip (throughput).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o admm_iter2 admm_iter2.cu && ./admm_iter2
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NP = 20, MP = 30, NV = NP + MP;
__device__ __forceinline__ double clampd(double v, double lo, double hi) { v = v < lo ? lo : v; return v > hi ? hi : v; }
__device__ __forceinline__ double relu_bits(double v) {
  const int hi = __double2hiint(v), m = ~(hi >> 31);
  return __hiloint2double(hi & m, __double2loint(v) & m);
}
struct Consts { double sigma, alpha, rho, kd; };

__device__ __forceinline__ double2 lds_v2(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds_f64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// ---------------------------------------------------------------- v0 (round 1)
__global__ void __maxnreg__(255) k_v0(const double *M, double *out, long long *cyc, int iters, Consts cs) {
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
    if (rowwarp) {
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
    } else {
      const double xn = alpha * acc + oma * p0;
      const double tz = (alpha * u1) * acc + (oma * z0 + y0 * u3);
      const double zn = clampd(tz, lo, hi);
      y0 = rho * (tz - zn);
      p0 = xn; z0 = zn;
      o = (sigma * xn - u0) + (u1 * rho) * (2.0 * zn - tz);
    }
    p ^= 1;
    if (act) vb[p][slot] = o;
    __syncthreads();
  }
  const long long t1 = clock64();
  out[blockIdx.x * 64 + tid] = p0 + z0 + y0 + s + zs + ys + g;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---------------------------------------------------------------- v3: two-phase split mat-vec
// shared memory (doubles): V[2][VLD] exchange vector (c | wp), XT[NP] x~, PA[64], PB[64]
constexpr int LA = 3, CA = 18;   // lanes per x~ row, columns per lane (LA * CA = 54 >= NV, zero padded)
constexpr int LB = 2, CB = 10;   // lanes per t row, columns per lane (LB * CB = NP)
constexpr int VLD = LA * CA + 2;
template <int VAR_PARALLEL>
__device__ __forceinline__ void v3_body(const double *M, const double *J, double *out, long long *cyc, int iters, Consts cs) {
  __shared__ __align__(16) double sm[2 * VLD + 32 + 64 + 64];
  const unsigned sb = (unsigned)__cvta_generic_to_shared(sm);
  const unsigned V0 = sb, V1 = sb + 8u * VLD, XT = sb + 8u * (2 * VLD), PA = XT + 8u * 32, PB = PA + 8u * 64;
  const int tid = threadIdx.x, lane = tid & 31;
  const bool rowwarp = tid < 32;
  const bool act = rowwarp ? lane < MP : lane < NP;  // owner lanes
  // phase A role: row ja, segment sa
  const int ja = tid / LA, sa = tid % LA;
  const bool actA = ja < NP;
  double Wa[CA];
#pragma unroll
  for (int k = 0; k < CA; k++) {
    const int col = sa * CA + k;
    Wa[k] = (actA && col < NV) ? M[(size_t)ja * NV + col] : 0.0;
  }
  // phase B role: row ib, segment sb2
  const int ib = tid / LB, sb2 = tid % LB;
  const bool actB = ib < MP;
  double Jb[CB];
#pragma unroll
  for (int k = 0; k < CB; k++) Jb[k] = actB ? J[(size_t)ib * NP + sb2 * CB + k] : 0.0;
  const double sigma = cs.sigma, alpha = cs.alpha, oma = 1.0 - cs.alpha, rho = cs.rho, rhoi = 1.0 / cs.rho, kd = cs.kd;
  double u0 = 0.01 * (lane + 1), u1 = rowwarp ? -1.0 : 1.0, u2 = 1.0, u3 = rowwarp ? 0.3 : 1.0 / rho;
  double lo = rowwarp ? kd * u1 : -0.5, hi = 0.5, Mi = 0.7;
  double p0 = 0, z0 = 0, y0 = 0, s = 0, zs = 0, ys = 0, g = Mi * (-u0);
  for (int e = tid; e < 2 * VLD + 32 + 64 + 64; e += 64) sm[e] = 0.0;
  __syncthreads();
  const unsigned my_v = 8u * (unsigned)(rowwarp ? NP + lane : lane);
  if (act) sts_f64(V0 + my_v, -u0);
  __syncthreads();
  int p = 0;
  const unsigned offA = 8u * (unsigned)(sa * CA), offB = 8u * (unsigned)(sb2 * CB);
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const unsigned vcur = p ? V1 : V0, vnext = p ? V0 : V1;
    // ---- phase A: partial of x~ row ja over the lane's column slice
    {
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
      for (int k = 0; k < CA / 4; k++) {
        const double2 va = lds_v2(vcur + offA + 32u * k), vd = lds_v2(vcur + offA + 32u * k + 16u);
        a0 = fma(Wa[4 * k], va.x, a0); a1 = fma(Wa[4 * k + 1], va.y, a1);
        a2 = fma(Wa[4 * k + 2], vd.x, a2); a3 = fma(Wa[4 * k + 3], vd.y, a3);
      }
      if (CA % 4) { const double2 va = lds_v2(vcur + offA + 8u * (CA - 2)); a0 = fma(Wa[CA - 2], va.x, a0); a1 = fma(Wa[CA - 1], va.y, a1); }
      sts_f64(PA + 8u * tid, (a0 + a1) + (a2 + a3));
    }
    __syncthreads();
    // ---- x~ (variable owners), optionally with the variable update right here
    double xt = 0.0;
    if (!rowwarp) {
      const unsigned pa = PA + 8u * (unsigned)(LA * (lane < NP ? lane : 0));
      xt = (lds_f64(pa) + lds_f64(pa + 8u)) + lds_f64(pa + 16u);
      if (lane < NP) sts_f64(XT + 8u * lane, xt);
      if (!VAR_PARALLEL) {
        const double xn = alpha * xt + oma * p0;
        const double tz = (alpha * u1) * xt + (oma * z0 + y0 * u3);
        const double zn = clampd(tz, lo, hi);
        y0 = rho * (tz - zn);
        p0 = xn; z0 = zn;
        if (act) sts_f64(vnext + my_v, (sigma * xn - u0) + (u1 * rho) * (2.0 * zn - tz));
      }
    }
    __syncthreads();
    // ---- phase B: partial of t row ib
    {
      double a0 = 0, a1 = 0;
#pragma unroll
      for (int k = 0; k < CB / 2; k++) {
        const double2 va = lds_v2(XT + offB + 16u * k);
        a0 = fma(Jb[2 * k], va.x, a0); a1 = fma(Jb[2 * k + 1], va.y, a1);
      }
      sts_f64(PB + 8u * tid, a0 + a1);
    }
    __syncthreads();
    // ---- updates: rows (warp 0) and, in the parallel variant, variables (warp 1)
    if (rowwarp) {
      const double2 pb = lds_v2(PB + 16u * (unsigned)(lane < MP ? lane : 0));
      const double acc = pb.x + pb.y;
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
      if (act) sts_f64(vnext + my_v, kd * wpen - (rho * lo) * g);
    } else if (VAR_PARALLEL) {
      const double xn = alpha * xt + oma * p0;
      const double tz = (alpha * u1) * xt + (oma * z0 + y0 * u3);
      const double zn = clampd(tz, lo, hi);
      y0 = rho * (tz - zn);
      p0 = xn; z0 = zn;
      if (act) sts_f64(vnext + my_v, (sigma * xn - u0) + (u1 * rho) * (2.0 * zn - tz));
    }
    p ^= 1;
    __syncthreads();
  }
  const long long t1 = clock64();
  out[blockIdx.x * 64 + tid] = p0 + z0 + y0 + s + zs + ys + g;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void __maxnreg__(128) k_v3(const double *M, const double *J, double *out, long long *cyc, int iters, Consts cs) { v3_body<0>(M, J, out, cyc, iters, cs); }
__global__ void __maxnreg__(128) k_v3p(const double *M, const double *J, double *out, long long *cyc, int iters, Consts cs) { v3_body<1>(M, J, out, cyc, iters, cs); }
__global__ void __maxnreg__(96) k_v3p96(const double *M, const double *J, double *out, long long *cyc, int iters, Consts cs) { v3_body<1>(M, J, out, cyc, iters, cs); }
__global__ void __maxnreg__(80) k_v3p80(const double *M, const double *J, double *out, long long *cyc, int iters, Consts cs) { v3_body<1>(M, J, out, cyc, iters, cs); }


// ---------------------------------------------------------------- v5: fused mat-vec, REGISTER-TILED R rows x C columns per lane
// Round-1 v0 delivers every vector entry to every lane: 64 lanes x 50 doubles = 25.6 KB from shared memory to
// the register file per iteration (~100 wavefronts), and every variant so far saturates near 2.0 G
// problem-iterations/s whatever its FP64 count.  Here a group of R lanes handles R rows: each lane multiplies a
// C = ceil(50 / R)-column slice of those R rows (R x C matrix entries in registers, C vector entries loaded) and the
// R partial sums are exchanged inside the group with log2(R) butterfly rounds, after which lane k of the group
// holds the full dot product of "its" row -- so lane <-> row ownership, the update code and the single barrier per
// iteration stay exactly as in v0.  Vector traffic drops by R.
template <int R>
__device__ __forceinline__ void v5_body(const double *M, double *out, long long *cyc, int iters, Consts cs) {
  constexpr int C = ((NV + R - 1) / R + 1) & ~1;  // even: 128-bit loads
  constexpr int VL = R * C + 2;
  __shared__ __align__(16) double vb[2][VL];
  const int tid = threadIdx.x, lane = tid & 31;
  const bool rowwarp = tid < 32;
  const bool act = rowwarp ? lane < MP : lane < NP;
  const int slot = rowwarp ? NP + lane : lane;   // my row of the iteration matrix
  const int grp = lane & ~(R - 1), sub = lane & (R - 1);
  double Mt[R][C];  // rows grp .. grp+R-1 of my warp's block, columns sub*C .. sub*C+C-1
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int row_lane = grp + r;
    const bool ra = rowwarp ? row_lane < MP : row_lane < NP;
    const int rs = rowwarp ? NP + row_lane : row_lane;
#pragma unroll
    for (int k = 0; k < C; k++) {
      const int col = sub * C + k;
      Mt[r][k] = (ra && col < NV) ? M[(size_t)rs * NV + col] : 0.0;
    }
  }
  const double sigma = cs.sigma, alpha = cs.alpha, oma = 1.0 - cs.alpha, rho = cs.rho, rhoi = 1.0 / cs.rho, kd = cs.kd;
  double u0 = 0.01 * (lane + 1), u1 = rowwarp ? -1.0 : 1.0, u2 = 1.0, u3 = rowwarp ? 0.3 : 1.0 / rho;
  double lo = rowwarp ? kd * u1 : -0.5, hi = 0.5, Mi = 0.7;
  double p0 = 0, z0 = 0, y0 = 0, s = 0, zs = 0, ys = 0, g = Mi * (-u0);
  for (int e = tid; e < 2 * VL; e += 64) (&vb[0][0])[e] = 0.0;
  __syncthreads();
  if (act) vb[0][slot] = -u0;
  __syncthreads();
  const unsigned sb = (unsigned)__cvta_generic_to_shared(&vb[0][0]);
  const unsigned off = 8u * (unsigned)(sub * C);
  int p = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const unsigned vcur = sb + (p ? 8u * VL : 0u) + off;
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0.0;
    double acc2[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc2[r] = 0.0;
#pragma unroll
    for (int k = 0; k < C; k += 2) {
      const double2 va = lds_v2(vcur + 8u * k);
#pragma unroll
      for (int r = 0; r < R; r++) {
        acc[r] = fma(Mt[r][k], va.x, acc[r]);
        acc2[r] = fma(Mt[r][k + 1], va.y, acc2[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] += acc2[r];
    // butterfly: after the round with distance d a lane keeps the rows whose bit d matches its own
#pragma unroll
    for (int d = R / 2; d >= 1; d >>= 1) {
      const bool upper = (sub & d) != 0;
#pragma unroll
      for (int r = 0; r < d; r++) {
        // rows [0, d) stay with the lower lanes, rows [d, 2d) go to the upper lanes
        const double send = upper ? acc[r] : acc[r + d];
        const double keep = upper ? acc[r + d] : acc[r];
        acc[r] = keep + __shfl_xor_sync(0xffffffffu, send, d);
      }
    }
    const double a = acc[0];
    double o;
    if (rowwarp) {
      const double stil = g - u3 * a;
      const double sn = alpha * stil + oma * s;
      const double ts = (alpha * u2) * stil + (oma * zs + ys * rhoi);
      const double zns = relu_bits(ts);
      ys = rho * (ts - zns);
      const double zt = a + u1 * stil;
      const double tz = alpha * zt + (oma * z0 + y0 * rhoi);
      const double zn = tz < hi ? tz : hi;
      y0 = rho * (tz - zn);
      s = sn; zs = zns; z0 = zn;
      const double wpen = rho * (2.0 * zn - tz);
      const double r1 = (sigma * sn - u0) + u2 * (rho * (2.0 * zns - ts)) + lo * wpen;
      g = Mi * r1;
      o = kd * wpen - (rho * lo) * g;
    } else {
      const double xn = alpha * a + oma * p0;
      const double tz = (alpha * u1) * a + (oma * z0 + y0 * u3);
      const double zn = clampd(tz, lo, hi);
      y0 = rho * (tz - zn);
      p0 = xn; z0 = zn;
      o = (sigma * xn - u0) + (u1 * rho) * (2.0 * zn - tz);
    }
    p ^= 1;
    if (act) vb[p][slot] = o;
    __syncthreads();
  }
  const long long t1 = clock64();
  out[blockIdx.x * 64 + tid] = p0 + z0 + y0 + s + zs + ys + g;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void __maxnreg__(255) k_v5r2(const double *M, double *out, long long *cyc, int iters, Consts cs) { v5_body<2>(M, out, cyc, iters, cs); }
__global__ void __maxnreg__(255) k_v5r4(const double *M, double *out, long long *cyc, int iters, Consts cs) { v5_body<4>(M, out, cyc, iters, cs); }
__global__ void __maxnreg__(255) k_v5r8(const double *M, double *out, long long *cyc, int iters, Consts cs) { v5_body<8>(M, out, cyc, iters, cs); }
__global__ void __maxnreg__(168) k_v5r4_168(const double *M, double *out, long long *cyc, int iters, Consts cs) { v5_body<4>(M, out, cyc, iters, cs); }

template <typename K, typename... A>
void run(const char *name, K kern, int threads, double *out, long long *cyc, A... args) {
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0);
  const int iters = 20000;
  for (int per_sm = 1; per_sm <= occ; per_sm++) {
    if (per_sm > 4 && per_sm != occ && per_sm % 2) continue;
    const int grid = 148 * per_sm;
    kern<<<grid, threads>>>(args..., out, cyc, 100, Consts{5e-10, 1.6, 0.1, 3.0});
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<grid, threads>>>(args..., out, cyc, iters, Consts{5e-10, 1.6, 0.1, 3.0});
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-30s CTAs/SM %d (max %d): %7.1f cycles/iter (CTA 0)  | chip %.3f G problem-iters/s | %s\n", name, per_sm, occ,
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
  run("v0 fused, 255 regs", k_v0, 64, out, cyc, (const double *)dM);
  run("v3 split, var || row update", k_v3p, 64, out, cyc, (const double *)dM, (const double *)dM);
  run("v5 fused, tiled R=2", k_v5r2, 64, out, cyc, (const double *)dM);
  run("v5 fused, tiled R=4", k_v5r4, 64, out, cyc, (const double *)dM);
  run("v5 fused, tiled R=8", k_v5r8, 64, out, cyc, (const double *)dM);
  run("v5 fused, tiled R=4, 168 regs", k_v5r4_168, 64, out, cyc, (const double *)dM);
  return 0;
}
