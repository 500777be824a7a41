// One-shot sum-allreduce of a small bucket over NVLink / NVSwitch peer memory (sm_100a), one process per GPU.
//
// What it replaces: the gradient / moment exchanges of the rollout-sharded update (the
// reference itself is single-process -- the reduction over the global batch it computes is the mean inside
// harl/algorithms/actors/happo.py:85-91 and harl/algorithms/critics/v_critic.py:116-133).  The buckets are tiny
// (<= 100 KB of gradients, or 3-4 doubles of loss normalisers), there are ~30 of them per iteration and each one
// sits between a backward kernel and its Adam step, so the exchange is pure latency: a ring / tree collective
// (~100 us through torch.distributed at 8 ranks) costs more than the update kernels it separates.
//
// Protocol (every rank runs the same kernel, `seq` = number of exchanges on this communicator so far + 1):
//   1. each CTA copies its part of the local bucket into this rank's slot `seq & 1` of a cudaMalloc'ed region that
//      every peer has mapped through CUDA IPC; the last CTA to finish (device counter) publishes `seq` into flag
//      [rank] of EVERY peer's flag array with system-scope release stores, one thread per peer (in parallel);
//   2. every CTA spins (system-scope acquire loads, bounded by a clock64 budget) until its own flag array shows
//      `seq` for all ranks, then reads all `world` slots straight out of peer memory (the loads of an element go out four
//      ranks at a time: one NVLink round trip per four ranks) and adds them IN RANK ORDER --
//      every rank computes the bit-identical sum, so the replicas never drift;
//   3. no trailing barrier: slots are double-buffered by `seq & 1`, and a rank can only overwrite slot parity p again
//      at seq + 2, which it reaches only after every peer has published seq + 1, i.e. finished reading seq.
// Cost: one launch, a flag round trip plus world / 4 data round trips over NVLink, world x bucket bytes of peer reads per GPU.
#include <stdlib.h>

#include "common.cuh"

namespace hb {

constexpr int COMM_MAX_WORLD = 16;
// Small CTAs on purpose: an exchange kernel spins while it waits for its peers, and the OTHER stream's persistent update
// kernel (148 CTAs x 320 threads x 168 registers, 222 KB of shared memory) must still find room on every SM -- a 128-thread,
// <= 64-register CTA co-resides with it; a 512-thread one did not, and a persistent kernel that starts 16 CTAs late runs
// up to twice as long (measured: +2.5 ms per iteration at 2 GPUs with the critic update on a side stream).
constexpr int COMM_GRID = 32;        // CTAs for a gradient-sized bucket (small buckets use fewer)
constexpr int COMM_THREADS = 128;
constexpr size_t COMM_HEADER = 4096;  // flags[COMM_MAX_WORLD] (u32), arrival counter (u64), error word

struct Comm {
  int rank, world, device;
  size_t slot_bytes;
  unsigned char* local;                      // cudaMalloc: [header][slot 0][slot 1]
  unsigned char* peer[COMM_MAX_WORLD];       // peer[r]: rank r's region mapped into this process (peer[rank] = local)
  uint32_t seq;
  int* host_err;                             // pinned, mapped: set by a kernel whose wait ran out
  int* dev_err;
};

struct CommPeers { unsigned char* p[COMM_MAX_WORLD]; };

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// V = float4 / double2 for 16-byte aligned buckets (tail elements as scalars), V = T otherwise
template <typename V>
__device__ __forceinline__ V ld_peer(const V* p) { return __ldcv(p); }
__device__ __forceinline__ void acc_add(float4& a, const float4& x) { a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w; }
__device__ __forceinline__ void acc_add(double2& a, const double2& x) { a.x += x.x; a.y += x.y; }
__device__ __forceinline__ void acc_add(float& a, const float& x) { a += x; }
__device__ __forceinline__ void acc_add(double& a, const double& x) { a += x; }

template <typename T, typename V, int VN>
__global__ void __launch_bounds__(COMM_THREADS) allreduce_oneshot_kernel(const __grid_constant__ CommPeers peers, T* __restrict__ buf, int64_t n, int rank,
                                                                         int world, uint32_t seq, size_t slot_bytes,
                                                                         long long spin_budget, int* err) {
  unsigned char* mine = peers.p[rank];
  uint32_t* flags = reinterpret_cast<uint32_t*>(mine);
  unsigned int* arrive = reinterpret_cast<unsigned int*>(mine + 256);
  const size_t slot_off = COMM_HEADER + (size_t)(seq & 1u) * slot_bytes;
  T* my_slot = reinterpret_cast<T*>(mine + slot_off);
  const int64_t nv = n / VN;
  const int64_t tid = (int64_t)blockIdx.x * COMM_THREADS + threadIdx.x, stride = (int64_t)gridDim.x * COMM_THREADS;
  __shared__ int s_last, s_ok;
  // ---- 1. publish the local bucket
  for (int64_t i = tid; i < nv; i += stride) reinterpret_cast<V*>(my_slot)[i] = reinterpret_cast<const V*>(buf)[i];
  for (int64_t i = nv * VN + tid; i < n; i += stride) my_slot[i] = buf[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    s_ok = 1;
    const unsigned int got = atomicAdd(arrive, 1u) + 1u;
    s_last = got == gridDim.x;
    if (s_last) {                       // every CTA of this exchange has arrived; the next exchange is a later kernel on the stream
      *arrive = 0;
      __threadfence_system();           // acquire side of the other CTAs' fence + arrive: their slot writes precede our flag stores
    }
  }
  __syncthreads();
  if (s_last && threadIdx.x < world) {  // last CTA: all slot writes are fenced; the flags of all peers go out in parallel
    st_release_sys(reinterpret_cast<uint32_t*>(peers.p[threadIdx.x]) + rank, seq);
  }
  // ---- 2. wait for every rank's flag, then sum the slots in rank order
  if (threadIdx.x < world) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - seq) < 0) {
      if (clock64() - t0 > spin_budget) {
        s_ok = 0;
        *err = 1 + (int)threadIdx.x;     // which rank never arrived
        break;
      }
    }
  }
  __syncthreads();
  if (!s_ok) return;                     // the host sees *err; the bucket is left unreduced
  // the peer loads of an element are issued four at a time before the adds (one NVLink round trip per four ranks, rank order kept)
  for (int64_t i = tid; i < nv; i += stride) {
    V acc;
    for (int r0 = 0; r0 < world; r0 += 4) {
      V x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (r0 + q < world) x[q] = ld_peer(reinterpret_cast<const V*>(peers.p[r0 + q] + slot_off) + i);
      if (r0 == 0) acc = x[0]; else acc_add(acc, x[0]);
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (r0 + q < world) acc_add(acc, x[q]);
    }
    reinterpret_cast<V*>(buf)[i] = acc;
  }
  for (int64_t i = nv * VN + tid; i < n; i += stride) {
    T acc = (T)0;
    for (int r0 = 0; r0 < world; r0 += 4) {
      T x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (r0 + q < world) x[q] = __ldcv(reinterpret_cast<const T*>(peers.p[r0 + q] + slot_off) + i);
      if (r0 == 0) acc = x[0]; else acc += x[0];
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (r0 + q < world) acc += x[q];
    }
    buf[i] = acc;
  }
}

}  // namespace hb

extern "C" {

int hb_comm_create(int32_t rank, int32_t world, size_t slot_bytes, void** comm_out, void* ipc_handle_out64) {
  using namespace hb;
  HB_CHECK_ARG(comm_out && ipc_handle_out64, "NULL output");
  HB_CHECK_ARG(world >= 1 && world <= COMM_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world size");
  HB_CHECK_ARG(slot_bytes > 0 && slot_bytes % 16 == 0, "slot_bytes must be a positive multiple of 16");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  Comm* c = (Comm*)calloc(1, sizeof(Comm));
  c->rank = rank;
  c->world = world;
  c->slot_bytes = slot_bytes;
  cudaError_t e = cudaGetDevice(&c->device);
  const size_t total = COMM_HEADER + 2 * slot_bytes;
  if (e == cudaSuccess) e = cudaMalloc(&c->local, total);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, total);
  if (e == cudaSuccess) e = cudaHostAlloc(&c->host_err, sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) { *c->host_err = 0; e = cudaHostGetDevicePointer(&c->dev_err, c->host_err, 0); }
  if (e == cudaSuccess) e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)ipc_handle_out64, c->local);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { free(c); return cuda_fail(e, "hb_comm_create"); }
  c->peer[rank] = c->local;
  *comm_out = c;
  return HB_OK;
}

int hb_comm_open_peers(void* comm, const void* all_handles) {
  using namespace hb;
  HB_CHECK_ARG(comm && all_handles, "NULL argument");
  Comm* c = (Comm*)comm;
  const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)all_handles;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail(e, "hb_comm_open_peers(cudaIpcOpenMemHandle)");
    c->peer[r] = (unsigned char*)p;
  }
  return HB_OK;
}

int hb_allreduce_bucket(void* comm, void* buf, int64_t n, int32_t dtype, void* stream) {
  using namespace hb;
  HB_CHECK_ARG(comm && buf && n > 0, "bad argument");
  HB_CHECK_ARG(dtype == 0 || dtype == 1, "dtype: 0 = float32, 1 = float64");
  Comm* c = (Comm*)comm;
  const size_t bytes = (size_t)n * (dtype == 0 ? 4 : 8);
  HB_CHECK_ARG(bytes <= c->slot_bytes, "bucket larger than the communicator's slot");
  HB_CHECK_ARG(((uintptr_t)buf & (dtype == 0 ? 3 : 7)) == 0, "bucket must be aligned to its element size");
  const bool vec = ((uintptr_t)buf & 15) == 0;
  for (int r = 0; r < c->world; ++r) HB_CHECK_ARG(c->peer[r] != nullptr, "hb_comm_open_peers has not run");
  if (*c->host_err) { set_error("hb_allreduce_bucket: rank %d never arrived in an earlier exchange", *c->host_err - 1); return HB_ERR_CUDA; }
  CommPeers peers;
  for (int r = 0; r < COMM_MAX_WORLD; ++r) peers.p[r] = c->peer[r];
  const uint32_t seq = ++c->seq;
  static const long long budget = (long long)(getenv("HB_COMM_TIMEOUT_S") ? atof(getenv("HB_COMM_TIMEOUT_S")) : 20.0) * 1900000000ll;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t per_cta = (int64_t)COMM_THREADS * 2 * (16 / (dtype == 0 ? 4 : 8));   // two 16-byte vectors per thread
  int grid = (int)((n + per_cta - 1) / per_cta);
  grid = grid < 1 ? 1 : (grid > COMM_GRID ? COMM_GRID : grid);
#define HB_AR(T, V, VN) allreduce_oneshot_kernel<T, V, VN><<<grid, COMM_THREADS, 0, st>>>(peers, (T*)buf, n, c->rank, c->world, seq, \
                                                                                              c->slot_bytes, budget, c->dev_err)
  if (dtype == 0) { if (vec) HB_AR(float, float4, 4); else HB_AR(float, float, 1); }
  else            { if (vec) HB_AR(double, double2, 2); else HB_AR(double, double, 1); }
#undef HB_AR
  HB_LAUNCH_DONE(st, "hb_allreduce_bucket");
  return HB_OK;
}

int hb_comm_status(void* comm) {
  if (!comm) return HB_ERR_INVALID;
  return *((hb::Comm*)comm)->host_err;   // 0 = healthy, 1 + r = rank r did not arrive within the spin budget
}

int hb_comm_destroy(void* comm) {
  using namespace hb;
  if (!comm) return HB_OK;
  Comm* c = (Comm*)comm;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  cudaFree(c->local);
  cudaFreeHost(c->host_err);
  free(c);
  return HB_OK;
}
}
