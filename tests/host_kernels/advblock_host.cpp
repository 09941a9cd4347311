// TEST INFRASTRUCTURE -- not part of the product.  Compiles the per-cell bodies of csrc/hdg_advblock.cuh with
// g++ (CUDA qualifiers defined away on the command line, see cuda_shim.h and build.py) so that the arithmetic
// of the device code can be checked against the numpy oracle without a GPU.  The engine never loads this.
#include "cuda_shim.h"
#include "hdg_advblock.cuh"

template <int K>
static void run(int upwind, int nc, const double* xy, const int* nbr, const double* Qstar, double adt, double* blk,
                float* blk32, const double* X, double* Y) {
  for (int cell = 0; cell < nc; ++cell) {
    if (upwind)
      advblock_build_cell<K, true>(xy, nbr, nc, cell, Qstar, adt, blk);
    else
      advblock_build_cell<K, false>(xy, nbr, nc, cell, Qstar, adt, blk);
    if (blk32) advblock_invert_cell<Dims<K>::NQ1>(nc, cell, blk, blk32);
    if (blk32 && X && Y) advblock_apply_cell<K>(nc, cell, blk32, X, Y);
  }
}

// all arrays in the engine's SoA layout: xy [6][nc], nbr [3][nc], Qstar / X / Y [2 NQ1][nc], blk / blk32 [NQ1*NQ1][nc];
// blk32 == NULL: build the blocks only; otherwise invert them (FP64 in place + FP32 copy) and, given X and Y, apply
extern "C" int advblock_host(int k, int upwind, int nc, const double* xy, const int* nbr, const double* Qstar,
                             double adt, double* blk, float* blk32, const double* X, double* Y) {
  switch (k) {
    case 1: run<1>(upwind, nc, xy, nbr, Qstar, adt, blk, blk32, X, Y); return 0;
    case 2: run<2>(upwind, nc, xy, nbr, Qstar, adt, blk, blk32, X, Y); return 0;
    case 3: run<3>(upwind, nc, xy, nbr, Qstar, adt, blk, blk32, X, Y); return 0;
    case 4: run<4>(upwind, nc, xy, nbr, Qstar, adt, blk, blk32, X, Y); return 0;
  }
  return 1;
}
