// p2p.cu -- vector exchange between the GPUs of one box over NVLink peer memory (sm_100a).
//
// Every H.v under row sharding needs the whole vector on every GPU (the reference: zero-padded n-long MPI_ALLREDUCE,
// more_tools.f90:2647,2772) and, for the distributed-slice entry points, the result back at the determinant's owner
// (the reference: MPI_REDUCE_SCATTER, mpi_routines.f90:1592).  Both are done here with plain stores into the peers'
// memory instead of an NCCL collective:
//   * each rank owns one "slab" (cudaMalloc) whose cudaIpcMemHandle is exchanged once per (re)allocation; peers map it
//     with cudaIpcOpenMemHandle, so a kernel on rank r can store into the buffers of every other rank through NVSwitch;
//   * gather ("push"): the producer of a vector block writes it into the same place of ALL ranks' x buffer with
//     coalesced stores, fences, and the last CTA publishes an epoch number into a flag word of every peer;
//     the consumer (a one-warp kernel in front of the H.v kernel) spins until all flags have reached the epoch.
//     x buffers are double buffered by epoch parity, which makes the write-after-read hazard impossible without
//     a second handshake;
//   * owner exchange: every rank scatters its result rows straight into the owners' y buffers (8-byte remote stores,
//     each element goes to exactly one peer), bracketed by a flag barrier.
// Spins are bounded (about 20 s) and raise a sticky error instead of hanging the GPU.
// NCCL stays for set-up (handle exchange), the small Davidson reductions and as the fallback when peer mapping fails
// (SQMC_P2P=0 forces it).
#include <cstdlib>
#include <cstring>

#include "handle.h"

namespace sqmc {

static const int kFlagWords = 64;  // [0,16) x epochs, [16,32) y epochs, [32,48) barrier epochs, 48 CTA counter, 49 timeout

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct PushArgs {
  double *dst[kMaxRanks];               // destination of element 0 on every rank (already offset)
  unsigned long long *flag[kMaxRanks];  // the word of every rank's flag array that belongs to the calling rank
  unsigned long long *counter;          // local CTA counter
  int nranks, rank;
};

// dst[p][i] = src[i] for every rank p; the last CTA publishes `epoch`
__global__ void __launch_bounds__(256) push_contig_kernel(const double *__restrict__ src, int64_t count, PushArgs A, unsigned long long epoch,
                                                          int misalign /* elements to the previous 128-byte boundary of dst */) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x - misalign; i < count; i += stride) {
    if (i < 0) continue;
    const double v = src[i];
#pragma unroll 1
    for (int j = 0; j < A.nranks; j++) {
      int p = A.rank + j;
      if (p >= A.nranks) p -= A.nranks;
      A.dst[p][i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long prev = atomicAdd(A.counter, 1ull);
    if (prev == gridDim.x - 1) {
      *A.counter = 0ull;
      __threadfence_system();
      for (int p = 0; p < A.nranks; p++) st_release_sys(A.flag[p], epoch);
    }
  }
}

struct ScatterArgs {
  double *dst[kMaxRanks];  // y buffer of every rank
  unsigned long long *flag[kMaxRanks];
  unsigned long long *counter;
  int nranks, rank;
};
// dst[owner[q]][pos[q]] = src[q]: every result row goes to the rank that owns the determinant
__global__ void __launch_bounds__(256) push_scatter_kernel(const double *__restrict__ src, const int32_t *__restrict__ owner, const int32_t *__restrict__ pos,
                                                           int64_t count, ScatterArgs A, unsigned long long epoch) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < count; q += stride) A.dst[owner[q]][pos[q]] = src[q];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long prev = atomicAdd(A.counter, 1ull);
    if (prev == gridDim.x - 1) {
      *A.counter = 0ull;
      __threadfence_system();
      for (int p = 0; p < A.nranks; p++) st_release_sys(A.flag[p], epoch);
    }
  }
}

// one warp: lane r waits until flags[r] >= epoch
__global__ void wait_flags_kernel(const unsigned long long *flags, int nranks, unsigned long long epoch, unsigned long long *timeout_word) {
  const int r = threadIdx.x;
  if (r < nranks) {
    unsigned long long spins = 0;
    while (ld_acquire_sys(flags + r) < epoch) {
      __nanosleep(64);
      if (++spins > (1ull << 28)) {  // ~20 s: a peer died or the ranks disagree about the call sequence
        *timeout_word = epoch;
        break;
      }
    }
  }
  __syncwarp();
  __threadfence_system();
}

// barrier = publish + wait in one single-warp kernel (no data)
struct BarrierArgs {
  unsigned long long *flag[kMaxRanks];
  int nranks;
};
__global__ void barrier_kernel(BarrierArgs A, const unsigned long long *flags, unsigned long long epoch, unsigned long long *timeout_word) {
  const int r = threadIdx.x;
  __threadfence_system();
  if (r < A.nranks) st_release_sys(A.flag[r], epoch);
  if (r < A.nranks) {
    unsigned long long spins = 0;
    while (ld_acquire_sys(flags + r) < epoch) {
      __nanosleep(64);
      if (++spins > (1ull << 28)) {
        *timeout_word = epoch;
        break;
      }
    }
  }
  __syncwarp();
  __threadfence_system();
}

static int nccl_barrier(cudaStream_t s) {
  DevBuf<int> t;
  SQ_CHECK(t.alloc(1));
  SQ_CUDA(cudaMemsetAsync(t.p, 0, sizeof(int), s));
  ncclResult_t r = ncclAllReduce(t.p, t.p, 1, ncclInt32, ncclSum, G.comm, s);
  if (r != ncclSuccess) { set_error("p2p: ncclAllReduce barrier failed: %s", ncclGetErrorString(r)); return 3; }
  SQ_CUDA(cudaStreamSynchronize(s));
  return 0;
}

void p2p_release(sqmc_b200_handle *h) {
  P2P &P = h->p2p;
  if (!P.slab) return;
  cudaStream_t s = G.stream;
  cudaStreamSynchronize(s);
  for (int r = 0; r < G.nranks; r++)
    if (r != G.rank && P.peer[r]) cudaIpcCloseMemHandle(P.peer[r]);
  if (G.nranks > 1 && G.comm) nccl_barrier(s);  // nobody frees while a peer still maps the slab
  cudaFree(P.slab);
  P = P2P();
}

static bool p2p_wanted() {
  const char *e = getenv("SQMC_P2P");
  return !(e && atoi(e) == 0);
}

// (re)allocate the slab for vectors of n doubles and map the peers' slabs.  Collective.  On any failure to map
// peer memory every rank falls back to NCCL together (the decision is all-reduced).
int p2p_setup(sqmc_b200_handle *h, int64_t n) {
  P2P &P = h->p2p;
  if (G.nranks == 1) return 0;
  if (G.nranks > kMaxRanks) { P.on = false; return 0; }
  cudaStream_t s = G.stream;
  if (P.slab && P.cap >= n) return 0;  // every rank sees the same n, so they all take the same branch
  p2p_release(h);
  int ok = p2p_wanted() ? 1 : 0;
  const int64_t cap = (n + 63) / 64 * 64;  // every buffer starts on a 512-byte boundary (linear textures over x)
  // layout (doubles): flags | x[0] (2 cap) | x[1] (2 cap) | xs[0] | xs[1] | y
  const size_t flag_bytes = 4096;
  const size_t bytes = flag_bytes + (size_t)cap * 7 * sizeof(double);
  if (ok && cudaMalloc(&P.slab, bytes) != cudaSuccess) { cudaGetLastError(); P.slab = nullptr; ok = 0; }
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof mine);
  if (ok) {
    cudaMemsetAsync(P.slab, 0, flag_bytes, s);
    if (cudaIpcGetMemHandle(&mine, P.slab) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  }
  // exchange handles (+ the ok bit) through NCCL
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  const int rec = 64 + 8;
  DevBuf<unsigned char> xbuf;
  SQ_CHECK(xbuf.alloc((int64_t)rec * G.nranks));
  std::vector<unsigned char> hostrec((size_t)rec * G.nranks, 0);
  memcpy(&hostrec[(size_t)rec * G.rank], &mine, 64);
  hostrec[(size_t)rec * G.rank + 64] = (unsigned char)ok;
  SQ_CUDA(cudaMemcpyAsync(xbuf.p + (size_t)rec * G.rank, &hostrec[(size_t)rec * G.rank], rec, cudaMemcpyHostToDevice, s));
  ncclResult_t nr = ncclAllGather(xbuf.p + (size_t)rec * G.rank, xbuf.p, rec, ncclUint8, G.comm, s);
  if (nr != ncclSuccess) { set_error("p2p: ncclAllGather(handles) failed: %s", ncclGetErrorString(nr)); return 3; }
  SQ_CUDA(cudaMemcpyAsync(hostrec.data(), xbuf.p, hostrec.size(), cudaMemcpyDeviceToHost, s));
  SQ_CUDA(cudaStreamSynchronize(s));
  for (int r = 0; r < G.nranks; r++) ok = ok && hostrec[(size_t)rec * r + 64];
  if (ok) {
    for (int r = 0; r < G.nranks && ok; r++) {
      if (r == G.rank) { P.peer[r] = P.slab; continue; }
      cudaIpcMemHandle_t hd;
      memcpy(&hd, &hostrec[(size_t)rec * r], 64);
      if (cudaIpcOpenMemHandle(&P.peer[r], hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        P.peer[r] = nullptr;
        ok = 0;
      }
    }
  }
  // all ranks must agree
  {
    DevBuf<int> t;
    SQ_CHECK(t.alloc(1));
    SQ_CUDA(cudaMemcpyAsync(t.p, &ok, sizeof(int), cudaMemcpyHostToDevice, s));
    nr = ncclAllReduce(t.p, t.p, 1, ncclInt32, ncclMin, G.comm, s);
    if (nr != ncclSuccess) { set_error("p2p: ncclAllReduce failed: %s", ncclGetErrorString(nr)); return 3; }
    SQ_CUDA(cudaMemcpyAsync(&ok, t.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    SQ_CUDA(cudaStreamSynchronize(s));
  }
  if (!ok) {
    for (int r = 0; r < G.nranks; r++)
      if (r != G.rank && P.peer[r]) cudaIpcCloseMemHandle(P.peer[r]);
    if (P.slab) cudaFree(P.slab);
    P = P2P();
    static bool said = false;
    if (!said && p2p_wanted() && G.rank == 0) fprintf(stderr, "[sqmc_b200] peer memory mapping unavailable: vector exchange falls back to NCCL\n");
    said = true;
    return 0;
  }
  P.on = true;
  P.cap = cap;
  P.bytes = bytes;
  P.epoch_x = P.epoch_y = P.epoch_bar = 0;
  P.last_stream = nullptr;
  return 0;
}

static inline unsigned long long *flag_base(void *slab) { return reinterpret_cast<unsigned long long *>(slab); }
static inline double *dbl_base(void *slab) { return reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(slab) + 4096); }
double *p2p_x(sqmc_b200_handle *h, int r, int b) { return dbl_base(h->p2p.peer[r]) + (size_t)b * 2 * h->p2p.cap; }
double *p2p_xs(sqmc_b200_handle *h, int r, int b) { return dbl_base(h->p2p.peer[r]) + (size_t)(4 + b) * h->p2p.cap; }
double *p2p_y(sqmc_b200_handle *h, int r) { return dbl_base(h->p2p.peer[r]) + (size_t)6 * h->p2p.cap; }

// work on one handle normally stays on one stream; when the caller switches streams the old one is drained first so
// that the epoch protocol's "stream order" argument keeps holding
static int p2p_stream(sqmc_b200_handle *h, cudaStream_t s) {
  P2P &P = h->p2p;
  if (P.last_stream && P.last_stream != s) SQ_CUDA(cudaStreamSynchronize(P.last_stream));
  P.last_stream = s;
  return 0;
}

int p2p_check(sqmc_b200_handle *h, cudaStream_t s) {
  if (!h->p2p.on) return 0;
  unsigned long long t = 0;
  SQ_CUDA(cudaMemcpyAsync(&t, flag_base(h->p2p.slab) + 49, sizeof t, cudaMemcpyDeviceToHost, s));
  SQ_CUDA(cudaStreamSynchronize(s));
  if (t) { set_error("p2p: timed out waiting for a peer GPU at epoch %llu (a rank died or the ranks disagree about the call sequence)", t); return 5; }
  return 0;
}

static unsigned push_grid(int64_t count) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>(div_up(count + 16, 256), (int64_t)G.sm_count * 4)); }

// all ranks: `count` doubles from src go to element `off` of x buffer (which = 0) or the staging buffer (which = 1)
// of EVERY rank; returns after queueing the wait, i.e. kernels queued on s afterwards see the complete vector.
// *full_out = this rank's buffer.
int p2p_gather(sqmc_b200_handle *h, const double *src, int64_t count, int64_t off, int which, cudaStream_t s, double **full_out) {
  P2P &P = h->p2p;
  SQ_CHECK(p2p_stream(h, s));
  const unsigned long long e = ++P.epoch_x;
  const int b = (int)(e & 1);
  PushArgs A;
  A.nranks = G.nranks;
  A.rank = G.rank;
  A.counter = flag_base(P.slab) + 48;
  for (int r = 0; r < G.nranks; r++) {
    A.dst[r] = (which == 0 ? p2p_x(h, r, b) : p2p_xs(h, r, b)) + off;
    A.flag[r] = flag_base(P.peer[r]) + G.rank;
  }
  const int mis = (int)(((uintptr_t)A.dst[G.rank] >> 3) & 15);
  push_contig_kernel<<<push_grid(count), 256, 0, s>>>(src, count, A, e, mis);
  SQ_LAUNCH_CHECK();
  wait_flags_kernel<<<1, 32, 0, s>>>(flag_base(P.slab), G.nranks, e, flag_base(P.slab) + 49);
  SQ_LAUNCH_CHECK();
  *full_out = which == 0 ? p2p_x(h, G.rank, b) : p2p_xs(h, G.rank, b);
  return 0;
}

int p2p_barrier(sqmc_b200_handle *h, cudaStream_t s) {
  P2P &P = h->p2p;
  SQ_CHECK(p2p_stream(h, s));
  const unsigned long long e = ++P.epoch_bar;
  BarrierArgs A;
  A.nranks = G.nranks;
  for (int r = 0; r < G.nranks; r++) A.flag[r] = flag_base(P.peer[r]) + 32 + G.rank;
  barrier_kernel<<<1, 32, 0, s>>>(A, flag_base(P.slab) + 32, e, flag_base(P.slab) + 49);
  SQ_LAUNCH_CHECK();
  return 0;
}

// all ranks: result rows src[0..count) go to (owner[q], pos[q]) in the owners' y buffers; on return (stream order)
// this rank's y buffer holds every row it owns.  *y_out = this rank's y buffer.
int p2p_to_owners(sqmc_b200_handle *h, const double *src, const int32_t *owner, const int32_t *pos, int64_t count, cudaStream_t s, double **y_out) {
  P2P &P = h->p2p;
  SQ_CHECK(p2p_barrier(h, s));  // every rank has finished reading its y buffer of the previous exchange
  const unsigned long long e = ++P.epoch_y;
  ScatterArgs A;
  A.nranks = G.nranks;
  A.rank = G.rank;
  A.counter = flag_base(P.slab) + 48;
  for (int r = 0; r < G.nranks; r++) {
    A.dst[r] = p2p_y(h, r);
    A.flag[r] = flag_base(P.peer[r]) + 16 + G.rank;
  }
  push_scatter_kernel<<<push_grid(count), 256, 0, s>>>(src, owner, pos, count, A, e);
  SQ_LAUNCH_CHECK();
  wait_flags_kernel<<<1, 32, 0, s>>>(flag_base(P.slab) + 16, G.nranks, e, flag_base(P.slab) + 49);
  SQ_LAUNCH_CHECK();
  *y_out = p2p_y(h, G.rank);
  return 0;
}

// ---- owner exchange fused into the producing kernel (bundle_hv_kernel<..., SCAT = true>)
int p2p_owner_begin(sqmc_b200_handle *h, OwnerScatter &O, cudaStream_t s) {
  P2P &P = h->p2p;
  SQ_CHECK(p2p_barrier(h, s));  // every rank has finished reading its result buffer of the previous exchange
  O.epoch = ++P.epoch_y;
  O.nranks = G.nranks;
  O.counter = flag_base(P.slab) + 48;
  for (int r = 0; r < kMaxRanks; r++) {
    O.dst[r] = r < G.nranks ? p2p_y(h, r) : nullptr;
    O.flag[r] = r < G.nranks ? flag_base(P.peer[r]) + 16 + G.rank : nullptr;
  }
  return 0;
}
int p2p_owner_wait(sqmc_b200_handle *h, const OwnerScatter &O, cudaStream_t s, double **y_out) {
  P2P &P = h->p2p;
  wait_flags_kernel<<<1, 32, 0, s>>>(flag_base(P.slab) + 16, G.nranks, O.epoch, flag_base(P.slab) + 49);
  SQ_LAUNCH_CHECK();
  *y_out = p2p_y(h, G.rank);
  return 0;
}

}  // namespace sqmc
