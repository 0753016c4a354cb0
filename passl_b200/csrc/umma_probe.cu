// Developer probe (tests/test_umma_probe_gpu.py): semantics of tcgen05 shared-memory descriptors for ROW-SHIFTED views of a
// SWIZZLE_128B tile — the building block of halo-tile 3x3 convolutions (one TMA halo load, nine shifted operand views).
//   A  : [256, 64] bf16 loaded by ONE TMA box {64, 256} (rows of 128 B, 8-row / 1024 B swizzle atoms)
//   B  : [64, 64]  bf16 (K-major)
//   D[m][n] = sum_k A[row(m)][k] * B[n][k],   row(m) = shift + (m / 8) * (sbo_bytes / 128) + m % 8
// The A descriptor starts at  a_smem + shift * 128  (not 1024-aligned when shift % 8 != 0); `base_offset` goes to descriptor
// bits [49, 52).  The host test finds which setting reproduces the expected product.
#include "common.cuh"
#include "host_utils.h"
#include "../../include/passl_b200.h"

#include <string.h>

namespace pb {

struct ProbeParams {
  CUtensorMap a_map, b_map;
  float* out;   // [128, 64]
  int shift_rows, sbo_bytes, base_offset, a_mn;   // a_mn: 1 = treat A tile as MN-major operand (K = rows)
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;                 // 256 * 128 = 32 KB
  uint8_t* b_smem = smem + 32768;         // 64 * 128  = 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint64_t* done = bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  const uint32_t warp = warp_id(), lane = lane_id();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar, 32768 + 8192);
      tma_load_2d(a_smem, &p.a_map, bar, 0, 0);
      tma_load_2d(b_smem, &p.b_map, bar, 0, 0);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    if (elect_one()) {
      if (!p.a_mn) {
        // K-major A: M rows = tile rows; 4 k-steps of 16 inside the 128 B row
        const uint32_t idesc = make_idesc_bf16(128, 64, false, false);
        uint64_t da = make_smem_desc_sw128(smem_u32(a_smem) + p.shift_rows * 128, 16, (uint32_t)p.sbo_bytes);
        da |= (uint64_t)(p.base_offset & 7) << 49;
        const uint64_t db = make_smem_desc_sw128(smem_u32(b_smem), 16, 1024);
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da + (uint64_t)((k * 32) >> 4), db + (uint64_t)((k * 32) >> 4), idesc, k > 0);
      } else {
        // MN-major A: K = tile rows (pixels), M = the 64 channels of a row (+ a second 64-wide atom LBO bytes further = rows +128)
        // D[m][n] = sum_{k<64} A[shift + k][m % 64 (+ second atom: rows shifted by 128)] * B[n][k]
        const uint32_t idesc = make_idesc_bf16(128, 64, true, false);
        for (int k = 0; k < 4; ++k) {
          uint64_t da = make_smem_desc_sw128(smem_u32(a_smem) + (p.shift_rows + k * 16) * 128, 128 * 128, 1024);
          da |= (uint64_t)((p.base_offset >= 0 ? p.base_offset : ((p.shift_rows + k * 16) & 7)) & 7) << 49;
          const uint64_t db = make_smem_desc_sw128(smem_u32(b_smem), 16, 1024);
          umma_bf16(tmem_base, da, db + (uint64_t)((k * 32) >> 4), idesc, k > 0);
        }
      }
      umma_commit(done);
    }
    __syncwarp();
  }
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t v[32];
  const int row = warp * 32 + lane;
  for (int c = 0; c < 2; ++c) {
    tmem_ld_32x32(tmem_base + ((warp * 32u) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.out[row * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

}  // namespace pb

using namespace pb;

extern "C" int passl_b200_umma_probe(const void* A, const void* B, float* out, int shift_rows, int sbo_bytes, int base_offset,
                                     int a_mn, void* stream) {
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  uint64_t ad[2] = {64, 256}, as[1] = {128};
  uint32_t abx[2] = {64, 256};
  int rc = make_tmap_bf16(&p.a_map, A, 2, ad, as, abx);
  if (rc) return rc;
  uint64_t bd[2] = {64, 64};
  uint32_t bbx[2] = {64, 64};
  rc = make_tmap_bf16(&p.b_map, B, 2, bd, as, bbx);
  if (rc) return rc;
  p.out = out; p.shift_rows = shift_rows; p.sbo_bytes = sbo_bytes; p.base_offset = base_offset; p.a_mn = a_mn;
  static bool attr = false;
  if (!attr) {
    PB_CUDA_CHECK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    attr = true;
  }
  umma_probe_kernel<<<1, 128, 32768 + 8192 + 64 + 1024, (cudaStream_t)stream>>>(p);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
