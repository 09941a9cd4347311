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

constexpr int HDG_MAX_PLANS = 2 + 16 + 1;  // cells, facets, P1 levels, CG dofs of the tracer path
constexpr int HDG_MAX_RANKS = 8;       // GPUs of one box (NVSwitch domain)
constexpr int HDG_RED_MAX = 8;         // doubles per all-reduce

// ---- peer-memory transport (NVLink P2P through CUDA IPC; one process per GPU) -------------------------
// Every rank owns one device buffer laid out as P2PLayout and maps the buffers of all peers.  A halo
// exchange is a *push*: the pack kernel stores the owned entries a peer needs straight into that
// peer's mailbox slab (remote stores over NVLink), then raises a flag in the peer's memory; the
// receiver's unpack kernel spins on its local flags and copies the slabs into the ghost entries.  An
// all-reduce is one single-CTA kernel: every rank stores its partial result into all peers' reduction
// boxes, waits for theirs and sums in rank order (bitwise identical on all ranks).  Slabs and boxes are
// double-buffered by a per-pair / per-reduction counter, which is sufficient because peer sets are
// symmetric and each side can only be one exchange ahead of the other (its next push needs the
// peer's previous one).  Spins are bounded (spin_cycles, two minutes by default, HDG_P2P_TIMEOUT_S): a timeout is
// FATAL -- the kernel sets the sticky `error`, does not unpack and does not advance its counters, every later
// transport kernel returns at once, and the host turns the flag into HDG_ECOMM at the next scalar readback of a
// Krylov loop and at the end of every C-ABI call that communicated (p2p_poll in hdg_engine.cu).
struct P2PHeader {
  unsigned long long halo_flag[HDG_MAX_RANKS];  // halo_flag[q]: number of pushes received from rank q
  unsigned long long red_flag[HDG_MAX_RANKS];   // red_flag[q]: number of reductions rank q contributed to
  double red_box[2][HDG_MAX_RANKS][HDG_RED_MAX];
  // exchange / reduction counters live on the device (not in kernel arguments) so that a captured
  // CUDA graph of a Krylov iteration can be replayed: every kernel derives "this exchange" from them
  unsigned long long pair_cnt[HDG_MAX_RANKS];   // completed exchanges with rank q
  unsigned long long red_cnt;                   // completed all-reduces
  int error;                                    // sticky: set by a kernel whose bounded spin ran out
  unsigned int ticket;                          // last-block detection of the push kernel
  unsigned int ticket2;                         // ... and of the wait/unpack kernel
  int pad0;
  long long spin_cycles;                        // bound of one spin in SM clock cycles
  int pad[10];
};
static_assert(sizeof(P2PHeader) % 8 == 0, "the mailbox slabs behind the header hold doubles");

struct P2P {
  bool enabled = false;
  size_t slab = 0;                          // doubles per (parity, sender) slab
  char* base = nullptr;                     // own buffer: P2PHeader, then mbox[2][nranks][slab]
  char* peer_base[HDG_MAX_RANKS] = {};      // mapped peer buffers (own entry = base)
};

__host__ __device__ inline P2PHeader* p2p_header(char* base) { return reinterpret_cast<P2PHeader*>(base); }
__host__ __device__ inline double* p2p_slab(char* base, size_t slab, int nranks, int parity, int sender) {
  return reinterpret_cast<double*>(base + sizeof(P2PHeader)) + ((size_t)parity * nranks + sender) * slab;
}

static_assert(HDG_MAX_RANKS <= 256, "k_p2p_wait_unpack / k_p2p_allreduce index peers by threadIdx.x");
struct P2PPeers {  // kernel argument: the peers of one exchange (at most HDG_MAX_RANKS - 1, checked by hdg_set_halo_plan)
  int npeers;
  int rank[HDG_MAX_RANKS];
  int send_ptr[HDG_MAX_RANKS + 1];
  int recv_ptr[HDG_MAX_RANKS + 1];          // ghost blocks relative to n_owned
  char* peer_base[HDG_MAX_RANKS];
};

constexpr long long HDG_SPIN_CYCLES = 230000000000ll;  // default bound: ~2 minutes at 1.9 GHz

__device__ __forceinline__ bool p2p_failed(const P2PHeader* own) { return *(volatile const int*)&own->error != 0; }

__device__ __forceinline__ void p2p_wait(volatile unsigned long long* flag, unsigned long long want, P2PHeader* own) {
  const long long t0 = clock64();
  const long long bound = own->spin_cycles > 0 ? own->spin_cycles : HDG_SPIN_CYCLES;
  while (*flag < want) {
    if (p2p_failed(own)) break;  // another waiter already gave up
    if (clock64() - t0 > bound) {
      *(volatile int*)&own->error = 1;
      break;
    }
  }
}

struct Comm {
  ncclComm_t nccl = nullptr;
  int rank = 0, nranks = 1;
  HaloPlanDev plans[HDG_MAX_PLANS];
  double *sendbuf = nullptr, *recvbuf = nullptr;  // device staging, grown on demand
  size_t send_cap = 0, recv_cap = 0;
  double* red = nullptr;  // device [16] reduction scratch
  int64_t exchanges = 0, allreduces = 0;
  P2P p2p;
  int* p2p_err_host = nullptr;  // pinned copy of P2PHeader::error (p2p_poll)
};

// sendbuf[i*ndof + d] = field[d*n_local + send_idx[i]]
// (T = element type of the field: double, or float for the vectors of the mixed-precision tentative solver)
template <typename T>
__global__ void k_halo_pack(int total, int ndof, int n_local, const int* __restrict__ send_idx,
                            const T* __restrict__ field, T* __restrict__ buf) {
  size_t n = (size_t)total * ndof;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int d = (int)(t / total);
    int i = (int)(t - (size_t)d * total);
    buf[(size_t)i * ndof + d] = field[(size_t)d * n_local + send_idx[i]];
  }
}
// field[d*n_local + n_owned + g] = recvbuf[g*ndof + d]
template <typename T>
__global__ void k_halo_unpack(int total, int ndof, int n_local, int n_owned, const T* __restrict__ buf,
                              T* __restrict__ field) {
  size_t n = (size_t)total * ndof;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int d = (int)(t / total);
    int g = (int)(t - (size_t)d * total);
    field[(size_t)d * n_local + n_owned + g] = buf[(size_t)g * ndof + d];
  }
}

// push: slab[(i - send_ptr[j]) * ndof + d] on peer j = field[d*n_local + send_idx[i]], then signal
template <typename T>
__global__ void __launch_bounds__(256) k_p2p_push(P2PPeers pp, int myrank, int nranks, size_t slab, int ndof,
                                                  int n_local, const int* __restrict__ send_idx,
                                                  const T* __restrict__ field, char* own_base) {
  if (p2p_failed(p2p_header(own_base))) return;  // fatal state: the host reports HDG_ECOMM
  const int total = pp.send_ptr[pp.npeers];
  const size_t n = (size_t)total * ndof;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int d = (int)(t / total);
    int i = (int)(t - (size_t)d * total);
    int j = 0;
    while (i >= pp.send_ptr[j + 1]) ++j;
    const unsigned long long cnt = p2p_header(own_base)->pair_cnt[pp.rank[j]] + 1ull;  // this exchange
    T* dst = reinterpret_cast<T*>(p2p_slab(pp.peer_base[j], slab, nranks, (int)(cnt & 1ull), myrank));
    dst[(size_t)(i - pp.send_ptr[j]) * ndof + d] = field[(size_t)d * n_local + send_idx[i]];
  }
  // the last CTA to finish publishes the data: all remote stores of this grid precede the flags
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    P2PHeader* own = p2p_header(own_base);
    unsigned int tk = atomicAdd(&own->ticket, 1u);
    if (tk == gridDim.x - 1) {
      own->ticket = 0;
      __threadfence_system();
      for (int j = 0; j < pp.npeers; ++j) {
        volatile unsigned long long* f = &p2p_header(pp.peer_base[j])->halo_flag[myrank];
        *f = own->pair_cnt[pp.rank[j]] + 1ull;
      }
      __threadfence_system();
    }
  }
}

// wait for every peer's push of this exchange, then field[d*n_local + n_owned + g] = slab_q[(g - recv_ptr[q])*ndof + d]
template <typename T>
__global__ void __launch_bounds__(256) k_p2p_wait_unpack(P2PPeers pp, int nranks, size_t slab, int ndof, int n_local,
                                                         int n_owned, char* own_base, T* __restrict__ field) {
  P2PHeader* own = p2p_header(own_base);
  __shared__ unsigned long long cnt[HDG_MAX_RANKS];
  if (p2p_failed(own)) return;
  if (threadIdx.x < pp.npeers) {
    cnt[threadIdx.x] = own->pair_cnt[pp.rank[threadIdx.x]] + 1ull;  // this exchange
    p2p_wait(&own->halo_flag[pp.rank[threadIdx.x]], cnt[threadIdx.x], own);
  }
  __syncthreads();
  if (p2p_failed(own)) return;  // timed out: no stale slab is unpacked, pair_cnt does not advance
  __threadfence_system();
  const int total = pp.recv_ptr[pp.npeers];
  const size_t n = (size_t)total * ndof;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int d = (int)(t / total);
    int g = (int)(t - (size_t)d * total);
    int j = 0;
    while (g >= pp.recv_ptr[j + 1]) ++j;
    const T* src = reinterpret_cast<const T*>(p2p_slab(own_base, slab, nranks, (int)(cnt[j] & 1ull), pp.rank[j]));
    field[(size_t)d * n_local + n_owned + g] = __ldcv(&src[(size_t)(g - pp.recv_ptr[j]) * ndof + d]);
  }
  // the last CTA to finish closes the exchange: every CTA has read the counters by then
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int tk = atomicAdd(&own->ticket2, 1u);
    if (tk == gridDim.x - 1) {
      own->ticket2 = 0;
      for (int j = 0; j < pp.npeers; ++j) own->pair_cnt[pp.rank[j]] = cnt[j];
      __threadfence();
    }
  }
}

// push + wait + unpack of one halo exchange in ONE kernel (half the launches of the pair above and no stream-order gap
// between them): every CTA first stores its share of the owned entries into the peers' slabs; the last CTA to finish
// raises the peers' flags; then every CTA waits for this rank's flags and copies its share of the slabs into the ghost
// entries; the last CTA to finish advances the exchange counters.  No CTA waits before it has pushed, and the flags of
// a rank depend on its own pushes only, so two ranks cannot wait for each other -- provided all CTAs of the grid are
// resident (the launch caps the grid at the number of SMs).
template <typename T>
__global__ void __launch_bounds__(256) k_p2p_exchange(P2PPeers pp, int myrank, int nranks, size_t slab, int ndof,
                                                      int n_local, int n_owned, const int* __restrict__ send_idx,
                                                      T* __restrict__ field, char* own_base) {
  P2PHeader* own = p2p_header(own_base);
  __shared__ unsigned long long cnt[HDG_MAX_RANKS];
  if (p2p_failed(own)) return;
  if (threadIdx.x < pp.npeers) cnt[threadIdx.x] = own->pair_cnt[pp.rank[threadIdx.x]] + 1ull;  // this exchange
  __syncthreads();
  {
    const int total = pp.send_ptr[pp.npeers];
    const size_t n = (size_t)total * ndof;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
      int d = (int)(t / total);
      int i = (int)(t - (size_t)d * total);
      int j = 0;
      while (i >= pp.send_ptr[j + 1]) ++j;
      T* dst = reinterpret_cast<T*>(p2p_slab(pp.peer_base[j], slab, nranks, (int)(cnt[j] & 1ull), myrank));
      dst[(size_t)(i - pp.send_ptr[j]) * ndof + d] = field[(size_t)d * n_local + send_idx[i]];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int tk = atomicAdd(&own->ticket, 1u);
    if (tk == gridDim.x - 1) {
      own->ticket = 0;
      __threadfence_system();
      for (int j = 0; j < pp.npeers; ++j) {
        volatile unsigned long long* f = &p2p_header(pp.peer_base[j])->halo_flag[myrank];
        *f = cnt[j];
      }
      __threadfence_system();
    }
  }
  if (threadIdx.x < pp.npeers) p2p_wait(&own->halo_flag[pp.rank[threadIdx.x]], cnt[threadIdx.x], own);
  __syncthreads();
  if (p2p_failed(own)) return;
  __threadfence_system();
  {
    const int total = pp.recv_ptr[pp.npeers];
    const size_t n = (size_t)total * ndof;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
      int d = (int)(t / total);
      int g = (int)(t - (size_t)d * total);
      int j = 0;
      while (g >= pp.recv_ptr[j + 1]) ++j;
      const T* src = reinterpret_cast<const T*>(p2p_slab(own_base, slab, nranks, (int)(cnt[j] & 1ull), pp.rank[j]));
      field[(size_t)d * n_local + n_owned + g] = __ldcv(&src[(size_t)(g - pp.recv_ptr[j]) * ndof + d]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int tk = atomicAdd(&own->ticket2, 1u);
    if (tk == gridDim.x - 1) {
      own->ticket2 = 0;
      for (int j = 0; j < pp.npeers; ++j) own->pair_cnt[pp.rank[j]] = cnt[j];
      __threadfence();
    }
  }
}

struct P2PAll {  // kernel argument of the all-reduce: every rank of the communicator
  char* base[HDG_MAX_RANKS];
};

// in-place all-reduce of the partial-sum slots part[0..nslots)[G] over all ranks, one CTA:
// finish the slots, store them into every peer's box, wait, sum in rank order, spread back
__global__ void __launch_bounds__(256) k_p2p_allreduce(P2PAll all, int myrank, int nranks, double* __restrict__ part,
                                                       int G, int nslots) {
  __shared__ double mine[HDG_RED_MAX];
  __shared__ double total[HDG_RED_MAX];
  if (p2p_failed(p2p_header(all.base[myrank]))) return;
  const unsigned long long count = p2p_header(all.base[myrank])->red_cnt + 1ull;  // this reduction
  for (int s = 0; s < nslots; ++s) {
    const double* p = part + (size_t)s * G;
    double v = 0.0;
    for (int i = threadIdx.x; i < G; i += blockDim.x) v += p[i];
    v = block_reduce(v);
    if (threadIdx.x == 0) mine[s] = v;
    __syncthreads();
  }
  const int parity = (int)(count & 1ull);
  if (threadIdx.x < nranks) {
    const int q = threadIdx.x;
    double* box = p2p_header(all.base[q])->red_box[parity][myrank];
    for (int s = 0; s < nslots; ++s) box[s] = mine[s];
    __threadfence_system();
    volatile unsigned long long* f = &p2p_header(all.base[q])->red_flag[myrank];
    *f = count;
  }
  __syncthreads();
  P2PHeader* own = p2p_header(all.base[myrank]);
  if (threadIdx.x < nranks) p2p_wait(&own->red_flag[threadIdx.x], count, own);
  __syncthreads();
  if (p2p_failed(own)) return;  // timed out: the partial sums stay local, red_cnt does not advance
  __threadfence_system();
  if (threadIdx.x < nslots) {
    double v = 0.0;
    for (int q = 0; q < nranks; ++q) v += __ldcv(&own->red_box[parity][q][threadIdx.x]);
    total[threadIdx.x] = v;
  }
  __syncthreads();
  for (int s = 0; s < nslots; ++s) {
    double* p = part + (size_t)s * G;
    for (int i = threadIdx.x; i < G; i += blockDim.x) p[i] = (i == 0) ? total[s] : 0.0;
  }
  if (threadIdx.x == 0) own->red_cnt = count;
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
