// One-shot all-reduce over NVLink peer memory for the two small exchanges of a data-parallel step (the 8 fp64 loss sums
// between forward and backward, the ~2 MB fp32 gradient buffer after backward).  Every rank's input lives in a symmetric
// allocation (torch.distributed._symmetric_memory: the peers' device pointers are exchanged once); a launch
//   1. publishes "my input of round `seq` is complete" in every peer's flag array (system-scope release),
//   2. waits until all peers have published the same round, then reads all `world` inputs over NVLink and adds them in rank
//      order (identical bits on every rank),
//   3. publishes "I have finished reading" and waits for the same from all peers, so that when the launch completes nobody is
//      still reading this rank's input and the next step may overwrite it.
// NCCL's ring / tree needs ~20-40 us for these sizes on 8 GPUs; this is one NVLink round trip plus 2 MB x world of peer reads.
// Spins are bounded (minutes): a missing peer traps (CUDA error at the next call) instead of hanging the GPU for good.
#include "common.cuh"

namespace marf {

constexpr int kMaxPeers = 8;
struct PeerArgs {
  const void* in[kMaxPeers];       // every rank's input (device pointers valid on this device)
  uint32_t* flags[kMaxPeers];      // every rank's flag array: [0, world) ready, [world, 2 world) done, [2 world] block counter
  int rank, world;
  uint32_t seq;                    // round number, strictly increasing over the calls of a group
  long long n;
  void* out;
  int in_place;                    // out == in[rank]: one block, n <= blockDim.x, the sum waits in registers for the done barrier
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_all(const uint32_t* my_flags, int world, uint32_t seq) {
  if ((int)threadIdx.x < world) {
    uint32_t spins = 0;
    // (bounded, but generously: ranks of a training job drift by seconds around checkpoints / visualisation; ~1-2 minutes)
    while ((int32_t)(ld_acquire_sys(my_flags + threadIdx.x) - seq) < 0) {
      __nanosleep(spins < 1024 ? 32 : 256);
      if (++spins > (1u << 28)) __trap();
    }
  }
  __syncthreads();
}

template <typename T>
static __global__ void __launch_bounds__(256) k_peer_allreduce(const PeerArgs a) {
  pdl_wait();                                            // this rank's input is complete (stream order)
  uint32_t* mine = a.flags[a.rank];
  if (blockIdx.x == 0 && (int)threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + a.rank, a.seq);
  }
  wait_all(mine, a.world, a.seq);
  T keep = T(0);
  if (a.in_place) {
    if ((long long)threadIdx.x < a.n)
      for (int r = 0; r < a.world; ++r) keep += reinterpret_cast<const T*>(a.in[r])[threadIdx.x];
  } else {
    // 16-byte vectors where the element count allows
    constexpr int V = 16 / sizeof(T);
    const long long nv = a.n / V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
      T acc[V];
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] = T(0);
      for (int r = 0; r < a.world; ++r) {
        const uint4 q = reinterpret_cast<const uint4*>(a.in[r])[i];
        const T* t = reinterpret_cast<const T*>(&q);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] += t[e];
      }
      reinterpret_cast<uint4*>(a.out)[i] = *reinterpret_cast<const uint4*>(acc);
    }
    for (long long i = nv * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
      T acc = T(0);
      for (int r = 0; r < a.world; ++r) acc += reinterpret_cast<const T*>(a.in[r])[i];
      reinterpret_cast<T*>(a.out)[i] = acc;
    }
  }
  // ---- done barrier: the last block of this rank publishes, then waits for every peer
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(mine + 2 * a.world, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) mine[2 * a.world] = 0u;          // (counter ready for the next round)
  if ((int)threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + a.world + a.rank, a.seq);
  }
  wait_all(mine + a.world, a.world, a.seq);
  if (a.in_place && (long long)threadIdx.x < a.n) reinterpret_cast<T*>(a.out)[threadIdx.x] = keep;
}

}  // namespace marf

extern "C" int marf_peer_allreduce(int device, int dtype, const void* const* peer_in, uint32_t* const* peer_flags, int rank, int world,
                                   void* out, long long n, uint32_t seq, void* stream) {
  using namespace marf;
  if (!peer_in || !peer_flags || !out || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || n <= 0 || (dtype != 0 && dtype != 1))
    return MARF_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return MARF_ERR_CUDA;
  PeerArgs a{};
  for (int r = 0; r < world; ++r) { a.in[r] = peer_in[r]; a.flags[r] = peer_flags[r]; }
  a.rank = rank; a.world = world; a.seq = seq; a.n = n; a.out = out;
  a.in_place = out == peer_in[rank];
  if (a.in_place && n > 256) return MARF_ERR_INVALID;
  const size_t esz = dtype == 0 ? 4 : 8;
  const int grid = a.in_place ? 1 : (int)std::max<long long>(1, std::min<long long>(64, (n * (long long)esz / 16 + 255) / 256));
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 0) launch_k(k_peer_allreduce<float>, grid, 256, 0, st, a);
  else launch_k(k_peer_allreduce<double>, grid, 256, 0, st, a);
  return cudaGetLastError() == cudaSuccess ? MARF_OK : MARF_ERR_CUDA;
}
