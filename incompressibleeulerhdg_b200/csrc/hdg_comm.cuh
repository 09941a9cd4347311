// Multi-GPU plumbing of the engine: one process per GPU, NCCL over NVLink 5 /
// NVSwitch.  Replaces what Firedrake/PETSc do implicitly for the reference under mpiexec: PyOP2 /
// PetscSF halo exchanges inside every assemble and MatMult, and the MPI_Allreduce of every Krylov
// dot product (hdg_imex.py:110 is the only place the reference even names a communicator).
//
//   * halo exchange of an SoA field [ndof][n_local]: one pack kernel (owned entries a peer holds as
//     ghosts -> entity-major send buffer), one grouped ncclSend/ncclRecv per peer, one unpack kernel
//     (ghost entries are numbered contiguously, grouped by owner, so the receive buffer maps 1:1);
//   * reductions: the engine's two-stage deterministic partial sums are finished per slot, summed
//     over ranks with one ncclAllReduce of <= 8 doubles and spread back into the partial arrays, so
//     the consumer kernels are identical in the single- and multi-GPU paths;
//   * all-gather of the first replicated multigrid level.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, normally the copy torch already loaded), so the
// library has no link-time dependency on a particular NCCL build.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>
#include <vector>

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

static NcclApi g_nccl;

static bool nccl_load() {
  if (g_nccl.lib) return true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    g_nccl.err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    return false;
  }
#define HDG_NCCL_SYM(field, name)                                           \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name)); \
  if (!g_nccl.field) {                                                      \
    g_nccl.err = std::string("dlsym failed: ") + name;                      \
    return false;                                                           \
  }
  HDG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  HDG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  HDG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  HDG_NCCL_SYM(Send, "ncclSend")
  HDG_NCCL_SYM(Recv, "ncclRecv")
  HDG_NCCL_SYM(GroupStart, "ncclGroupStart")
  HDG_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  HDG_NCCL_SYM(AllReduce, "ncclAllReduce")
  HDG_NCCL_SYM(AllGather, "ncclAllGather")
  HDG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef HDG_NCCL_SYM
  g_nccl.lib = lib;
  return true;
}

struct HaloPlanDev {
  bool set = false;
  int n_owned = 0, n_local = 0;
  std::vector<int> peers, send_ptr, recv_off, recv_cnt;
  int* send_idx = nullptr;  // device [total_send]
  int total_send = 0, total_recv = 0;
};

constexpr int HDG_MAX_PLANS = 2 + 16;  // cells, facets, P1 levels

struct Comm {
  ncclComm_t nccl = nullptr;
  int rank = 0, nranks = 1;
  HaloPlanDev plans[HDG_MAX_PLANS];
  double *sendbuf = nullptr, *recvbuf = nullptr;  // device staging, grown on demand
  size_t send_cap = 0, recv_cap = 0;
  double* red = nullptr;  // device [16] reduction scratch
  int64_t exchanges = 0, allreduces = 0;
};

// sendbuf[i*ndof + d] = field[d*n_local + send_idx[i]]
__global__ void k_halo_pack(int total, int ndof, int n_local, const int* __restrict__ send_idx,
                            const double* __restrict__ field, double* __restrict__ buf) {
  size_t n = (size_t)total * ndof;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int d = (int)(t / total);
    int i = (int)(t - (size_t)d * total);
    buf[(size_t)i * ndof + d] = field[(size_t)d * n_local + send_idx[i]];
  }
}
// field[d*n_local + n_owned + g] = recvbuf[g*ndof + d]
__global__ void k_halo_unpack(int total, int ndof, int n_local, int n_owned, const double* __restrict__ buf,
                              double* __restrict__ field) {
  size_t n = (size_t)total * ndof;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int d = (int)(t / total);
    int g = (int)(t - (size_t)d * total);
    field[(size_t)d * n_local + n_owned + g] = buf[(size_t)g * ndof + d];
  }
}

// red[slot] = sum of the G partials of slot `slot` (one block per slot, fixed tree)
__global__ void k_part_finish(const double* __restrict__ part, int G, double* __restrict__ red) {
  const double* p = part + (size_t)blockIdx.x * G;
  double v = 0.0;
  for (int i = threadIdx.x; i < G; i += blockDim.x) v += p[i];
  v = block_reduce(v);
  if (threadIdx.x == 0) red[blockIdx.x] = v;
}
// part[slot][0] = red[slot], part[slot][1..G) = 0: the consumers' re-reduction then yields red[slot]
__global__ void k_part_spread(double* __restrict__ part, int G, const double* __restrict__ red) {
  double* p = part + (size_t)blockIdx.x * G;
  for (int i = threadIdx.x; i < G; i += blockDim.x) p[i] = (i == 0) ? red[blockIdx.x] : 0.0;
}
// replicated-level gather: out[gid[q*0 + ...]]: buf is [nranks][maxcnt]; entry (q, i) goes to out[gid[ptr[q]+i]]
__global__ void k_gather_scatter(int nranks, int maxcnt, const int* __restrict__ ptr, const int* __restrict__ gid,
                                 const double* __restrict__ buf, double* __restrict__ out) {
  int total = ptr[nranks];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    int q = 0;
    while (t >= ptr[q + 1]) ++q;
    out[gid[t]] = buf[(size_t)q * maxcnt + (t - ptr[q])];
  }
}
