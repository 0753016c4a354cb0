// Shared pieces of the fused similarity-softmax-CE kernels (fp32 SIMT variant and tcgen05 bf16 variant).
#pragma once
#include "common.cuh"

namespace pb {

// Combine split partials -> lse, per-row loss, scalar outputs.  One CTA.
//   out[0] = loss_scale * mean_i(loss_i), out[1] = acc1 (%), out[2] = acc5 (%)
static __global__ void simce_finalize_kernel(const float* part_m, const float* part_l, const int* part_cnt, const float* tgt,
                                      int N, int splits, int extra_col, float loss_scale, float* lse_out,
                                      float* loss_rows, float* out) {
  __shared__ float red[3][32];
  float s_loss = 0.f, s_a1 = 0.f, s_a5 = 0.f;
  for (int row = threadIdx.x; row < N; row += blockDim.x) {
    float m = extra_col ? tgt[row] : -INFINITY;
    int cnt = 0;
    for (int s = 0; s < splits; ++s) {
      m = fmaxf(m, part_m[(size_t)row * splits + s]);
      cnt += part_cnt[(size_t)row * splits + s];
    }
    float l = extra_col ? expf(tgt[row] - m) : 0.f;
    for (int s = 0; s < splits; ++s) {
      float pm = part_m[(size_t)row * splits + s];
      if (pm > -INFINITY) l += part_l[(size_t)row * splits + s] * expf(pm - m);
    }
    float lse = m + logf(l);
    float li = lse - tgt[row];
    lse_out[row] = lse;
    if (loss_rows) loss_rows[row] = li;
    s_loss += li;
    s_a1 += (cnt == 0) ? 1.f : 0.f;
    s_a5 += (cnt < 5) ? 1.f : 0.f;
  }
  s_loss = warp_sum(s_loss); s_a1 = warp_sum(s_a1); s_a5 = warp_sum(s_a5);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s_loss; red[1][warp] = s_a1; red[2][warp] = s_a5; }
  __syncthreads();
  if (warp == 0) {
    int nw = blockDim.x >> 5;
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
