// Shared pieces of the fused similarity-softmax-CE kernels (fp32 SIMT variant and tcgen05 bf16 variant).
#pragma once
#include "common.cuh"

namespace pb {

// Combine split partials -> lse, per-row loss, scalar outputs.  One CTA.
//   out[0] = loss_scale * mean_i(loss_i), out[1] = acc1 (%), out[2] = acc5 (%)
static __global__ void simce_finalize_kernel(const float* part_m, const float* part_l, const int* part_cnt, const float* tgt,
                                      int N, int splits, int extra_col, float loss_scale, float* lse_out,
                                      float* loss_rows, float* out) {
  // one warp per row, lanes over the split partials (coalesced); 32 warps per CTA
  __shared__ float red[3][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float s_loss = 0.f, s_a1 = 0.f, s_a5 = 0.f;
  for (int row = warp; row < N; row += nw) {
    const float t = tgt[row];
    float m = extra_col ? t : -INFINITY;
    int cnt = 0;
    for (int s = lane; s < splits; s += 32) {
      m = fmaxf(m, part_m[(size_t)row * splits + s]);
      cnt += part_cnt[(size_t)row * splits + s];
    }
    m = warp_max(m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    float l = 0.f;
    for (int s = lane; s < splits; s += 32) {
      const float pm = part_m[(size_t)row * splits + s];
      if (pm > -INFINITY) l += part_l[(size_t)row * splits + s] * __expf(pm - m);
    }
    l = warp_sum(l);
    if (extra_col) l += __expf(t - m);
    const float lse = m + __logf(l);
    const float li = lse - t;
    if (lane == 0) {
      lse_out[row] = lse;
      if (loss_rows) loss_rows[row] = li;
      s_loss += li;
      s_a1 += (cnt == 0) ? 1.f : 0.f;
      s_a5 += (cnt < 5) ? 1.f : 0.f;
    }
  }
  if (lane == 0) { red[0][warp] = s_loss; red[1][warp] = s_a1; red[2][warp] = s_a5; }
  __syncthreads();
  if (warp == 0) {
    float a = lane < nw ? red[0][lane] : 0.f, b = lane < nw ? red[1][lane] : 0.f, c = lane < nw ? red[2][lane] : 0.f;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) {
      out[0] = loss_scale * a / N;
      out[1] = 100.f * b / N;
      out[2] = 100.f * c / N;
    }
  }
}

// per-row gradient scale: grow[i] = dloss * factor
static __global__ void fill_rowgrad_kernel(float* grow, const float* dloss, float factor, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) grow[i] = (dloss ? dloss[0] : 1.f) * factor;
}

}  // namespace pb
