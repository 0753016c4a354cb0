// Shared pieces of the fused similarity-softmax-CE kernels (fp32 SIMT variant and tcgen05 bf16 variant).
#pragma once
#include <string.h>
#include "common.cuh"
#include "host_utils.h"

namespace pb {

// Combine split partials -> lse, per-row loss, scalar outputs.  One CTA.
//   out[0] = loss_scale * mean_i(loss_i), out[1] = acc1 (%), out[2] = acc5 (%)
// scratch layout (floats): [nblk][3] per-CTA partial sums, then one unsigned ticket counter
static __global__ void __launch_bounds__(256) simce_finalize_kernel(const float* part_m, const float* part_l, const int* part_cnt,
                                      const float* tgt, int N, int splits, int extra_col, float loss_scale, float* lse_out,
                                      float* loss_rows, float* out, float* scratch) {
  // programmatic dependent launch: when launched with the PDL attribute this grid may start while the producer of the partials
  // is still running; the wait returns once that grid has completed and flushed (no-op for ordinary launches)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // one warp per row (8 rows per CTA, all rows of the batch in flight at once), lanes over the split partials
  __shared__ float red[3][8];
  __shared__ bool is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  float s_loss = 0.f, s_a1 = 0.f, s_a5 = 0.f;
  if (row < N) {
    const float t = tgt[row];
    float m = extra_col ? t : -INFINITY;
    int cnt = 0;
    float pm[8], pl[8];                       // up to 256 splits
    int ns = 0;
    for (int s = lane; s < splits; s += 32, ++ns) {
      pm[ns] = part_m[(size_t)row * splits + s];
      pl[ns] = part_l[(size_t)row * splits + s];
      cnt += part_cnt[(size_t)row * splits + s];
      m = fmaxf(m, pm[ns]);
    }
    m = warp_max(m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    float l = 0.f;
    for (int i = 0; i < ns; ++i)
      if (pm[i] > -INFINITY) l += pl[i] * __expf(pm[i] - m);
    l = warp_sum(l);
    if (extra_col) l += __expf(t - m);
    const float lse = m + __logf(l);
    const float li = lse - t;
    if (lane == 0) {
      lse_out[row] = lse;
      if (loss_rows) loss_rows[row] = li;
      s_loss = li;
      s_a1 = (cnt == 0) ? 1.f : 0.f;
      s_a5 = (cnt < 5) ? 1.f : 0.f;
    }
  }
  if (lane == 0) { red[0][warp] = s_loss; red[1][warp] = s_a1; red[2][warp] = s_a5; }
  __syncthreads();
  unsigned* ticket = reinterpret_cast<unsigned*>(scratch + (size_t)gridDim.x * 3);
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; }
    scratch[blockIdx.x * 3 + 0] = a; scratch[blockIdx.x * 3 + 1] = b; scratch[blockIdx.x * 3 + 2] = c;
    __threadfence();
    is_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);   // wraps back to 0: no re-initialisation needed
  }
  __syncthreads();
  if (is_last && warp == 0) {
    __threadfence();
    float a = 0.f, b = 0.f, c = 0.f;
    for (int k = lane; k < (int)gridDim.x; k += 32) {
      a += __ldcg(scratch + k * 3 + 0); b += __ldcg(scratch + k * 3 + 1); c += __ldcg(scratch + k * 3 + 2);
    }
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) {
      out[0] = loss_scale * a / N;
      out[1] = 100.f * b / N;
      out[2] = 100.f * c / N;
    }
  }
}

static inline long long simce_finalize_scratch_bytes(int N) { return ((long long)((N + 7) / 8) * 3 + 4) * 4; }

// scratch must hold simce_finalize_scratch_bytes(N); its ticket word is zeroed here once per call (stream ordered)
// pdl: the ticket word was already zeroed by the caller BEFORE the producer kernel and this launch carries the programmatic
// stream serialization attribute (it must directly follow the producer in the stream).
static inline cudaError_t launch_simce_finalize(const float* part_m, const float* part_l, const int* part_cnt, const float* tgt,
                                                int N, int splits, int extra_col, float loss_scale, float* lse_out,
                                                float* loss_rows, float* out, float* scratch, cudaStream_t st,
                                                bool pdl = false) {
  const int nblk = (N + 7) / 8;
  if (!pdl) {
    cudaError_t e = cudaMemsetAsync(scratch + (size_t)nblk * 3, 0, 4, st);
    if (e != cudaSuccess) return e;
    simce_finalize_kernel<<<nblk, 256, 0, st>>>(part_m, part_l, part_cnt, tgt, N, splits, extra_col, loss_scale, lse_out,
                                                loss_rows, out, scratch);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(nblk); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, simce_finalize_kernel, part_m, part_l, part_cnt, tgt, N, splits, extra_col, loss_scale, lse_out,
                            loss_rows, out, scratch);
}

// per-row gradient scale: grow[i] = dloss * factor
static __global__ void fill_rowgrad_kernel(float* grow, const float* dloss, float factor, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) grow[i] = (dloss ? dloss[0] : 1.f) * factor;
}

}  // namespace pb
