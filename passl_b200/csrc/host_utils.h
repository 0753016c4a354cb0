// Host-side helpers: error codes, TMA tensor-map encoding through the driver entry point
// (resolved at run time so the library links without libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace pb {

// C-ABI error codes (0 = ok, >0 = cudaError_t, <0 = passl_b200 argument / contract error).
enum : int {
  PB_OK = 0,
  PB_ERR_BAD_ARG = -1,
  PB_ERR_UNSUPPORTED = -2,
  PB_ERR_TMAP = -3,
  PB_ERR_WORKSPACE = -4,
};

#define PB_CUDA_CHECK(expr)                     \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

// every kernel launch site is followed by PB_LAUNCH_CHECK(): it also feeds the launch counter bench.py reports
extern "C" long long passl_b200_launch_counter_add(long long n);
#define PB_LAUNCH_CHECK()                       \
  do {                                          \
    passl_b200_launch_counter_add(1);           \
    PB_CUDA_CHECK(cudaGetLastError());          \
  } while (0)

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// bf16 tensor map, SWIZZLE_128B, zero OOB fill.  dims/box are innermost-first; strides_bytes has
// rank-1 entries (stride of dims 1..rank-1).  Returns 0 on success.
inline int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box,
                          CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B,
                          CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) return PB_ERR_TMAP;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr,
            "passl_b200: cuTensorMapEncodeTiled failed (%d): rank=%d base=%p dims=[%llu,%llu,%llu,%llu] "
            "box=[%u,%u,%u,%u] stride1=%llu\n",
            (int)r, rank, base, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
            (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
            rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
            (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    return PB_ERR_TMAP;
  }
  return PB_OK;
}

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace pb
