// TEST INFRASTRUCTURE -- not part of the product.  Lets g++ compile the engine's device headers so that the
// arithmetic of thread-per-entity kernels can be executed on the CPU and compared with the numpy oracle when no
// GPU is around.  A kernel is called as a plain function with a 1 x 1 launch geometry: its grid-stride loop then
// visits every entity serially.  Only kernels without shared memory, barriers or warp shuffles are meaningful
// this way.  Compile with
//   -D__device__= -D__host__= -D__global__= -D__forceinline__=inline '-D__launch_bounds__(...)='
// (tests/host_kernels/build.py).  The engine itself has no CPU path and never sees this file.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
using std::fabs;
using std::fma;
using std::sqrt;
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
struct HostDim3 {
  unsigned x, y, z;
};
static const HostDim3 gridDim = {1, 1, 1}, blockDim = {1, 1, 1}, blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0};
// stand-ins that only have to compile (kernels that reduce across a block are not run on the host)
#define __shared__ static
static inline void __syncthreads() {}
static inline double __shfl_down_sync(unsigned, double v, int) { return 0.0 * v; }
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline void __threadfence() {}
static inline int atomicAdd(int* p, int v) {
  int old = *p;
  *p += v;
  return old;
}
