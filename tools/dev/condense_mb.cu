// Development microbenchmark (not part of the product): variants of the k >= 3 condensation kernel on random affine
// cells, timed with CUDA events and compared with the register kernel k_condense<K>.  Built by tools/dev/build_mb.sh.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "hdg_poisson_s.cuh"

template <int K, int BD, int E, int E2, bool DIAG, int M0, int MC>
__device__ __forceinline__ void condense_block_x(const Geo& g, const double (&nu)[3][2], double tau, const int (&fl)[3],
                                                 const double (&v)[MC][Dims<K>::NP], int nc, int cell, bool active,
                                                 double* __restrict__ SK) {
  using T = RefTables<K>;
  constexpr int NP = Dims<K>::NP, NL1 = Dims<K>::NL1, NL = Dims<K>::NL;
  const double nn = (g.n[E][0] * g.n[E2][0] + g.n[E][1] * g.n[E2][1]) * g.le[E] * g.le[E2] * g.idetJ;
  double nub[3][2];
  HDG_UNROLL
  for (int e = 0; e < 3; ++e) nub[e][0] = nub[e][1] = 0.0;
  nub[E2][0] = opaque(nu[E2][0]);
  nub[E2][1] = opaque(nu[E2][1]);
  double taub = opaque(tau);
  HDG_UNROLL
  for (int m2 = 0; m2 < NL1; ++m2) {
    if (DIAG && m2 < M0) continue;
    double s[MC];
    HDG_UNROLL
    for (int mm = 0; mm < MC; ++mm) s[mm] = 0.0;
    HDG_UNROLL
    for (int a = 0; a < NP; ++a)
      if (W_nonzero<K>(E2, m2, a)) {
        const double w = W_entry<K>(g, nub, taub, E2, m2, a);
        HDG_UNROLL
        for (int mm = 0; mm < MC; ++mm)
          if (!DIAG || M0 + mm <= m2) s[mm] = fma(v[mm][a], w, s[mm]);
      }
    const double sg2 = flip_sign(fl[E2], m2);
    HDG_UNROLL
    for (int mm = 0; mm < MC; ++mm) {
      const int m = M0 + mm;
      if (!DIAG || m <= m2) {
        double t = s[mm];
        if (T::NN(E, E2, m, m2) != 0.0) t = fma(-nn, T::NN(E, E2, m, m2), t);
        if (DIAG && m == m2) t -= tau * g.le[E];
        t *= flip_sign(fl[E], m) * sg2;
        const int r = E * NL1 + m, c = E2 * NL1 + m2;
        if (active) {
          SK[(size_t)(r * NL + c) * nc + cell] = t;
          if (c != r) SK[(size_t)(c * NL + r) * nc + cell] = t;
        }
      }
    }
    opaque_after(nub[E2][0], nub[E2][1], taub, s[0]);
  }
}

template <int K, int BD, bool SYNC, int E, int M0, int MC>
__device__ __forceinline__ void condense_rows_x(const Geo& g, const double (&nu)[3][2], double tau, const int (&fl)[3],
                                                const LsCol<BD>& L, int nc, int cell, bool active,
                                                double* __restrict__ SK) {
  constexpr int NP = Dims<K>::NP;
  double v[MC][NP];
  HDG_UNROLL
  for (int mm = 0; mm < MC; ++mm)
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) v[mm][a] = W_entry<K>(g, nu, tau, E, M0 + mm, a);
  chol_solve_s<NP, MC, BD>(L, v);
  if (SYNC) __syncthreads();
  condense_block_x<K, BD, E, E, true, M0, MC>(g, nu, tau, fl, v, nc, cell, active, SK);
  if (SYNC) __syncthreads();
  condense_block_x<K, BD, E, (E + 1) % 3, false, M0, MC>(g, nu, tau, fl, v, nc, cell, active, SK);
  if (SYNC) __syncthreads();
}

template <int K, int BD, bool SYNC, int MCH, int E>
__device__ __forceinline__ void condense_facet_x(const Geo& g, const double (&nu)[3][2], double tau, const int (&fl)[3],
                                                 const LsCol<BD>& L, int nc, int cell, bool active,
                                                 double* __restrict__ SK) {
  constexpr int NL1 = Dims<K>::NL1;
  if constexpr (MCH >= NL1) {
    condense_rows_x<K, BD, SYNC, E, 0, NL1>(g, nu, tau, fl, L, nc, cell, active, SK);
  } else {
    condense_rows_x<K, BD, SYNC, E, 0, MCH>(g, nu, tau, fl, L, nc, cell, active, SK);
    condense_rows_x<K, BD, SYNC, E, MCH, NL1 - MCH>(g, nu, tau, fl, L, nc, cell, active, SK);
  }
}

// W warps per block that pass the phases of a cell batch together (barriers between the phases keep them within one
// instruction-cache window of each other); MCH = rows of a facet solved together
template <int K, int W, int MCH, bool SYNC>
__global__ void __launch_bounds__(32 * W) k_condense_x(const double* __restrict__ xy, const int* __restrict__ flip,
                                                       int nc, double tau, double* __restrict__ SK) {
  using D = Dims<K>;
  constexpr int BD = 32 * W;
  extern __shared__ double Lsh[];
  const LsCol<BD> L{Lsh + threadIdx.x};
  for (int base = blockIdx.x * BD; base < nc; base += gridDim.x * BD) {
    const bool active = base + (int)threadIdx.x < nc;
    const int cell = active ? base + (int)threadIdx.x : nc - 1;
    Geo g = make_geo(xy, nc, cell);
    build_H_s<K, BD>(g, tau, L);
    if (SYNC) __syncthreads();
    cholesky_s<D::NP, BD>(L);
    if (SYNC) __syncthreads();
    double nu[3][2];
    int fl[3];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      fl[e] = flip[(size_t)e * nc + cell];
      nu[e][0] = g.Ji[0][0] * g.n[e][0] + g.Ji[0][1] * g.n[e][1];
      nu[e][1] = g.Ji[1][0] * g.n[e][0] + g.Ji[1][1] * g.n[e][1];
    }
    condense_facet_x<K, BD, SYNC, MCH, 0>(g, nu, tau, fl, L, nc, cell, active, SK);
    condense_facet_x<K, BD, SYNC, MCH, 1>(g, nu, tau, fl, L, nc, cell, active, SK);
    condense_facet_x<K, BD, SYNC, MCH, 2>(g, nu, tau, fl, L, nc, cell, active, SK);
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <class F>
static float time_ms(F launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return ms / reps;
}

static double maxrel(const std::vector<double>& a, const std::vector<double>& b) {
  double d = 0, m = 0;
  for (size_t i = 0; i < a.size(); ++i) { d = fmax(d, fabs(a[i] - b[i])); m = fmax(m, fabs(b[i])); }
  return d / m;
}

template <int K, int W, int MCH, bool SYNC>
static void run_x(const double* xy, const int* flip, int nc, double* SK, const std::vector<double>& ref, int reps) {
  constexpr int BD = 32 * W;
  const size_t smem = (size_t)Dims<K>::NH * BD * sizeof(double);
  CK(cudaFuncSetAttribute(k_condense_x<K, W, MCH, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaMemset(SK, 0, sizeof(double) * ref.size()));
  float ms = time_ms([&] { k_condense_x<K, W, MCH, SYNC><<<(nc + BD - 1) / BD, BD, smem>>>(xy, flip, nc, 1.0, SK); }, reps);
  std::vector<double> out(ref.size());
  CK(cudaMemcpy(out.data(), SK, sizeof(double) * ref.size(), cudaMemcpyDeviceToHost));
  int nblk = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, k_condense_x<K, W, MCH, SYNC>, BD, smem));
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, k_condense_x<K, W, MCH, SYNC>));
  printf("{\"k\": %d, \"variant\": \"x\", \"warps_per_block\": %d, \"rows_per_solve\": %d, \"sync\": %d, \"ms\": %.4f, \"maxrel_vs_register_kernel\": %.2e, \"blocks_per_sm\": %d, \"regs\": %d, \"local_bytes\": %zu}\n",
         K, W, MCH, (int)SYNC, ms, maxrel(out, ref), nblk, fa.numRegs, fa.localSizeBytes);
  fflush(stdout);
}

template <int K>
static void run_k(int nc, int reps) {
  constexpr int NL = Dims<K>::NL;
  std::vector<double> xy(6 * (size_t)nc);
  std::vector<int> flip(3 * (size_t)nc);
  srand(1);
  auto rnd = [] { return rand() / (double)RAND_MAX; };
  for (int c = 0; c < nc; ++c) {  // random affine images of the reference triangle, positive orientation
    double x0 = rnd(), y0 = rnd(), a = 0.5 + rnd(), b = 0.5 + rnd(), th = 6.283 * rnd(), sh = 0.6 * (rnd() - 0.5);
    double J00 = a * cos(th), J10 = a * sin(th), J01 = b * (-sin(th) + sh * cos(th)), J11 = b * (cos(th) + sh * sin(th));
    xy[0 * (size_t)nc + c] = x0; xy[1 * (size_t)nc + c] = y0;
    xy[2 * (size_t)nc + c] = x0 + J00; xy[3 * (size_t)nc + c] = y0 + J10;
    xy[4 * (size_t)nc + c] = x0 + J01; xy[5 * (size_t)nc + c] = y0 + J11;
    for (int e = 0; e < 3; ++e) flip[e * (size_t)nc + c] = rand() & 1;
  }
  double *dxy, *SK; int* dflip;
  const size_t nsk = (size_t)NL * NL * nc;
  CK(cudaMalloc(&dxy, sizeof(double) * xy.size())); CK(cudaMalloc(&dflip, sizeof(int) * flip.size()));
  CK(cudaMalloc(&SK, sizeof(double) * nsk));
  CK(cudaMemcpy(dxy, xy.data(), sizeof(double) * xy.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dflip, flip.data(), sizeof(int) * flip.size(), cudaMemcpyHostToDevice));
  std::vector<double> ref(nsk), out(nsk);
  float ms = time_ms([&] { k_condense<K><<<(nc + 127) / 128, 128>>>(dxy, dflip, nc, 1.0, SK); }, reps);
  CK(cudaMemcpy(ref.data(), SK, sizeof(double) * nsk, cudaMemcpyDeviceToHost));
  printf("{\"k\": %d, \"nc\": %d, \"variant\": \"register kernel k_condense\", \"ms\": %.4f}\n", K, nc, ms);
  CK(cudaMemset(SK, 0, sizeof(double) * nsk));
  constexpr int BDs = LsBlock<K>::BD;
  ms = time_ms([&] { k_condense_b<K><<<(nc + BDs - 1) / BDs, BDs>>>(dxy, dflip, nc, 1.0, SK); }, reps);
  CK(cudaMemcpy(out.data(), SK, sizeof(double) * nsk, cudaMemcpyDeviceToHost));
  printf("{\"k\": %d, \"variant\": \"k_condense_b\", \"ms\": %.4f, \"maxrel_vs_register_kernel\": %.2e}\n", K, ms, maxrel(out, ref));
  fflush(stdout);
  constexpr int NL1 = Dims<K>::NL1;
  constexpr int HALF = (NL1 + 1) / 2;
  run_x<K, 1, NL1, false>(dxy, dflip, nc, SK, ref, reps);
  run_x<K, 1, HALF, false>(dxy, dflip, nc, SK, ref, reps);
  run_x<K, 4, NL1, true>(dxy, dflip, nc, SK, ref, reps);
  run_x<K, 4, HALF, true>(dxy, dflip, nc, SK, ref, reps);
  run_x<K, 4, HALF, false>(dxy, dflip, nc, SK, ref, reps);
  if (K == 4) {
    run_x<K, 7, NL1, true>(dxy, dflip, nc, SK, ref, reps);
    run_x<K, 7, HALF, true>(dxy, dflip, nc, SK, ref, reps);
    run_x<K, 6, HALF, true>(dxy, dflip, nc, SK, ref, reps);
  } else {
    run_x<K, 8, NL1, true>(dxy, dflip, nc, SK, ref, reps);
    run_x<K, 8, HALF, true>(dxy, dflip, nc, SK, ref, reps);
    run_x<K, 16, HALF, true>(dxy, dflip, nc, SK, ref, reps);
  }
  cudaFree(dxy); cudaFree(dflip); cudaFree(SK);
}

int main(int argc, char** argv) {
  int nc = argc > 1 ? atoi(argv[1]) : 1000000, reps = argc > 2 ? atoi(argv[2]) : 3;
  run_k<4>(nc, reps);
  run_k<3>(nc, reps);
  return 0;
}
