// Embedding exchange over NVLink peer memory: all-gather and its backward (reduce-scatter) as ONE kernel each that
// reads the peers' buffers directly (P2P loads through NVSwitch) and synchronises with signal flags in peer memory — no NCCL
// launch, no staging copy.  The reference does list-all_gather + concat (moco.py:198-210, distributed/nn/functional.py:100-127).
//
// Protocol (one "slot" = data buffer + flag row per rank, two slots used alternately by the host wrapper):
//   1. the producer of this rank's shard has written it to the local slot buffer (earlier kernel on the same stream);
//   2. block 0: __threadfence_system(), then store `epoch` into flag[my_rank] of EVERY rank's flag row (remote st.release.sys);
//   3. every block: for each peer q spin on the LOCAL flag row until flag[q] >= epoch (ld.acquire.sys), then copy (gather) or
//      accumulate (reduce-scatter) q's shard with 16-byte P2P loads.
// A rank can run at most one epoch ahead of the slowest peer before it blocks in step 3, so two slots are enough: the buffer
// written for epoch e+2 was last read by peers in epoch e, and every peer finished its epoch-e kernel before it signalled e+1.
#include <string.h>
#include "common.cuh"
#include "host_utils.h"
#include "../../include/passl_b200.h"

namespace pb {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
  const void* data[kMaxPeers];   // peer-mapped base of each rank's data buffer for this slot
  unsigned* flags[kMaxPeers];    // peer-mapped base of each rank's flag row [world] for this slot
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void peer_signal_and_wait(const PeerPtrs& pp, int rank, int world, unsigned epoch) {
  if (blockIdx.x == 0 && threadIdx.x < (unsigned)world) {
    __threadfence_system();
    st_release_sys(pp.flags[threadIdx.x] + rank, epoch);        // my shard is ready: tell rank threadIdx.x
  }
  if (threadIdx.x < (unsigned)world) {
    const unsigned* f = pp.flags[rank] + threadIdx.x;
    long long spins = 0;
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {
      if (++spins > (1LL << 31)) { printf("passl_b200 peer exchange: rank %d timed out waiting for rank %d\n", rank, (int)threadIdx.x); __trap(); }
    }
  }
  __syncthreads();
}

// out[q*n16 + i] = data_q[i]   (uint4 units)
__global__ void peer_allgather_kernel(PeerPtrs pp, uint4* __restrict__ out, long long n16, int rank, int world, unsigned epoch) {
  peer_signal_and_wait(pp, rank, world, epoch);
  const long long total = n16 * world;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i / n16);
    const long long j = i - (long long)q * n16;
    out[i] = __ldcv(reinterpret_cast<const uint4*>(pp.data[q]) + j);     // volatile: peer lines must not be served from a stale L1
  }
}

// out[i] = sum_q data_q[rank*n4 + i]   (float4 units; each rank's buffer holds the full [world * n] gradient)
__global__ void peer_reduce_scatter_kernel(PeerPtrs pp, float4* __restrict__ out, long long n4, int rank, int world, unsigned epoch) {
  peer_signal_and_wait(pp, rank, world, epoch);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < world; ++q) {
      const float4 v = __ldcv(reinterpret_cast<const float4*>(pp.data[q]) + (long long)rank * n4 + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[i] = acc;
  }
}

// this rank's keys fp32 [n*D] -> bf16 in its own slot buffer; then (system-scope fence) flag[rank] = epoch in every rank's row
__global__ void peer_publish_keys_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n4, PeerPtrs pp,
                                         int rank, int world, unsigned epoch, unsigned* __restrict__ done) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = u;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicInc(done, gridDim.x - 1) == gridDim.x - 1);      // wraps to 0: reusable
  __syncthreads();
  if (last && threadIdx.x < (unsigned)world) {
    __threadfence_system();
    st_release_sys(pp.flags[threadIdx.x] + rank, epoch);
  }
}

}  // namespace pb

using namespace pb;

static int fill_ptrs(PeerPtrs& pp, const void* const* data_ptrs, void* const* flag_ptrs, int world) {
  if (world < 1 || world > kMaxPeers) return PB_ERR_BAD_ARG;
  for (int q = 0; q < world; ++q) {
    if (!data_ptrs[q] || !flag_ptrs[q]) return PB_ERR_BAD_ARG;
    pp.data[q] = data_ptrs[q];
    pp.flags[q] = reinterpret_cast<unsigned*>(flag_ptrs[q]);
  }
  return PB_OK;
}

// data_ptrs / flag_ptrs: HOST arrays of `world` device pointers (this process's mappings of every rank's slot, own rank included).
extern "C" int passl_b200_peer_allgather(const void* shard, const void* const* data_ptrs, void* const* flag_ptrs, void* out,
                                         long long shard_bytes, int rank, int world, unsigned epoch, void* stream) {
  if (!shard || shard_bytes <= 0 || shard_bytes % 16 || rank < 0 || rank >= world) return PB_ERR_BAD_ARG;
  PeerPtrs pp;
  int rc = fill_ptrs(pp, data_ptrs, flag_ptrs, world);
  if (rc) return rc;
  // publish this rank's shard in its own slot (stream ordered before the kernel's system-scope fence + flag stores)
  PB_CUDA_CHECK(cudaMemcpyAsync(const_cast<void*>(pp.data[rank]), shard, (size_t)shard_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  const long long n16 = shard_bytes / 16;
  long long blocks = (n16 * world + 255) / 256;
  if (blocks > num_sms()) blocks = num_sms();      // all blocks must be co-resident: every block spins on the flags
  peer_allgather_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(pp, reinterpret_cast<uint4*>(out), n16, rank, world, epoch);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_peer_reduce_scatter_f32(const float* grad_all, const void* const* data_ptrs, void* const* flag_ptrs, float* out,
                                                  long long shard_elems, int rank, int world, unsigned epoch, void* stream) {
  if (!grad_all || shard_elems <= 0 || shard_elems % 4 || rank < 0 || rank >= world) return PB_ERR_BAD_ARG;
  PeerPtrs pp;
  int rc = fill_ptrs(pp, data_ptrs, flag_ptrs, world);
  if (rc) return rc;
  PB_CUDA_CHECK(cudaMemcpyAsync(const_cast<void*>(pp.data[rank]), grad_all, (size_t)shard_elems * world * sizeof(float),
                                cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  const long long n4 = shard_elems / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > num_sms()) blocks = num_sms();
  peer_reduce_scatter_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(pp, reinterpret_cast<float4*>(out), n4, rank, world, epoch);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

// Publish this rank's key shard for the peer-sharded InfoNCE (passl_b200_infonce_tc_fwd_peer): cast keys fp32 [n, D] to bf16 into
// this rank's own data buffer (data_ptrs[rank]) and raise flag[rank] = epoch in every rank's flag row.  `done`: uint32 zeroed once.
extern "C" int passl_b200_peer_publish_keys_bf16(const float* keys, int n, int D, const void* const* data_ptrs, void* const* flag_ptrs,
                                                 int rank, int world, unsigned epoch, void* done, void* stream) {
  if (!keys || n <= 0 || D <= 0 || ((long long)n * D) % 4 || rank < 0 || rank >= world || !done) return PB_ERR_BAD_ARG;
  PeerPtrs pp;
  int rc = fill_ptrs(pp, data_ptrs, flag_ptrs, world);
  if (rc) return rc;
  const long long n4 = (long long)n * D / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 64) blocks = 64;
  peer_publish_keys_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(keys, reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(pp.data[rank])),
                                                                          n4, pp, rank, world, epoch, reinterpret_cast<unsigned*>(done));
  PB_LAUNCH_CHECK();
  return PB_OK;
}

// Exchange buffers.  They are the one place where the library allocates device memory itself: a CUDA IPC handle refers to the
// base of a cudaMalloc allocation, which a framework caching allocator does not expose.  create: cudaMalloc + zero + export handle
// (64 bytes); open: map a peer's buffer into this process with the CURRENT device as the accessor (lazy peer access, the way NCCL
// opens P2P buffers); close / destroy release them.
extern "C" int passl_b200_peer_buffer_create(long long bytes, void** base, unsigned char* handle64) {
  if (bytes <= 0 || !base || !handle64) return PB_ERR_BAD_ARG;
  PB_CUDA_CHECK(cudaMalloc(base, (size_t)bytes));
  PB_CUDA_CHECK(cudaMemset(*base, 0, (size_t)bytes));
  PB_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  PB_CUDA_CHECK(cudaIpcGetMemHandle(&h, *base));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return PB_OK;
}
extern "C" int passl_b200_peer_buffer_open(const unsigned char* handle64, void** mapped) {
  if (!handle64 || !mapped) return PB_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  PB_CUDA_CHECK(cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess));
  return PB_OK;
}
extern "C" int passl_b200_peer_buffer_close(void* mapped) {
  PB_CUDA_CHECK(cudaIpcCloseMemHandle(mapped));
  return PB_OK;
}
extern "C" int passl_b200_peer_buffer_destroy(void* base) {
  PB_CUDA_CHECK(cudaFree(base));
  return PB_OK;
}
