// Image input stage on the device: the per-view image ops that sit right before the backbone.
//
// Reference (per view, CPU workers): PIL image -> RandomResizedCrop = crop + Image.resize(224, bilinear | bicubic)
// (configs/simclr/simclr_r50_IM.yaml:35-39, configs/moco/moco_v2_r50.yaml, paddle.vision.transforms over Pillow) ->
// [ColorJitter, GaussianBlur: not built yet] -> RandomGrayscale (transforms.py:150-170) -> RandomHorizontalFlip -> Transpose ->
// NormalizeImage (transforms.py:462-467).  Here: decoded uint8 HWC images already in HBM (ragged sizes, one byte offset per
// image), crop boxes / grayscale / flip decisions drawn on the host, and three integer kernels + one finalize kernel:
//
//   resample_coeffs   per (item, axis, output index): window start, tap count and fixed-point taps — Pillow's precompute_coeffs +
//                     normalize_coeffs_8bpc (libImaging/Resample.c) in double precision with explicitly rounded operations
//                     (no FMA contraction), so the taps are the integers Pillow computes
//   resample_h / _v   Pillow's two 8-bit passes: horizontal into a uint8 intermediate, then vertical; accumulators start at
//                     1 << 21, result = clip8(acc >> 22)
//   views_finalize    grayscale (L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16, Convert.c), horizontal flip, HWC -> CHW and
//                     (x * scale - mean) / std through a 3 x 256 table evaluated in double, stored as float32
//
// All of it is byte / integer work bound by HBM traffic: source crop read once, S x crop_h x 3 intermediate, S x S x 3 output.
#include "common.cuh"
#include "host_utils.h"
#include "input_stage_core.h"
#include "../../include/passl_b200.h"

namespace pb {

using namespace istage;

// tables: bounds [items][2 axes][S][2] (start, count), taps [items][2][S][kmax]; status[0] |= 1 on a bad box, |= 2 when kmax is short
__global__ void resample_coeffs_kernel(const long long* __restrict__ src_off, const int* __restrict__ src_h,
                                       const int* __restrict__ src_w, const int* __restrict__ item_img,
                                       const int* __restrict__ item_box, int* __restrict__ bounds, int* __restrict__ taps,
                                       int* __restrict__ status, int items, int S, int kmax, int bicubic) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)items * 2 * S) return;
  const int xx = (int)(idx % S);
  const int axis = (int)((idx / S) % 2);                          // 0 horizontal, 1 vertical
  const int m = (int)(idx / (2 * S));
  const ItemGeom g = item_geom(src_off, src_h, src_w, item_img, item_box, m);
  int* b = bounds + window_index(m, axis, S, xx) * 2;
  int* k = taps + window_index(m, axis, S, xx) * kmax;
  if (!geom_ok(g)) {
    b[0] = 0; b[1] = 0;
    atomicOr(status, 1);
    return;
  }
  const int bad = resample_window(axis == 0 ? g.cw : g.ch, S, xx, bicubic, kmax, b, k);
  if (bad) atomicOr(status, bad);
}

// tmp[m][y][xx][c] for y < crop_h: one thread per (m, y, xx)
__global__ void resample_h_kernel(const unsigned char* __restrict__ src, const long long* __restrict__ src_off,
                                  const int* __restrict__ src_h, const int* __restrict__ src_w, const int* __restrict__ item_img,
                                  const int* __restrict__ item_box, const int* __restrict__ bounds, const int* __restrict__ taps,
                                  unsigned char* __restrict__ tmp, int items, int S, int kmax, int max_crop_h) {
  const long long total = (long long)items * max_crop_h * S;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(idx % S);
    const int y = (int)((idx / S) % max_crop_h);
    const int m = (int)(idx / ((long long)S * max_crop_h));
    const ItemGeom g = item_geom(src_off, src_h, src_w, item_img, item_box, m);
    if (!geom_ok(g) || y >= g.ch) continue;
    h_pass_pixel(src, g, bounds, taps, tmp, m, y, xx, S, kmax, max_crop_h);
  }
}

// dst[m][yy][x][c]: one thread per (m, yy, x)
__global__ void resample_v_kernel(const unsigned char* __restrict__ tmp, const long long* __restrict__ src_off,
                                  const int* __restrict__ src_h, const int* __restrict__ src_w, const int* __restrict__ item_img,
                                  const int* __restrict__ item_box, const int* __restrict__ bounds, const int* __restrict__ taps,
                                  unsigned char* __restrict__ dst, int items, int S, int kmax, int max_crop_h) {
  const long long total = (long long)items * S * S;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % S);
    const int yy = (int)((idx / S) % S);
    const int m = (int)(idx / ((long long)S * S));
    const ItemGeom g = item_geom(src_off, src_h, src_w, item_img, item_box, m);
    const int* b = bounds + window_index(m, 1, S, yy) * 2;
    if (!geom_ok(g) || b[1] == 0) {                               // bad box or short kmax: defined output, flagged in status
      unsigned char* o = dst + idx * 3;
      o[0] = o[1] = o[2] = 0;
      continue;
    }
    v_pass_pixel(tmp, bounds, taps, dst, m, yy, x, S, kmax, max_crop_h);
  }
}

// out[m][c][y][x] fp32 from img[m][y][x'][c] uint8, x' = flip ? S-1-x : x
__global__ void views_finalize_kernel(const unsigned char* __restrict__ img, const int* __restrict__ gray, const int* __restrict__ flip,
                                      float* __restrict__ out, int items, int S, double scale, float m0, float m1, float m2,
                                      float s0, float s1, float s2) {
  __shared__ float lut[3 * 256];
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
    const int c = i >> 8;
    lut[i] = normalize_entry(i & 255, scale, c == 0 ? m0 : (c == 1 ? m1 : m2), c == 0 ? s0 : (c == 1 ? s1 : s2));
  }
  __syncthreads();
  const long long total = (long long)items * S * S;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % S);
    const int y = (int)((idx / S) % S);
    const int m = (int)(idx / ((long long)S * S));
    finalize_pixel(img, lut, out, m, y, x, S, gray[m], flip[m]);
  }
}

// ---- ColorJitter: up to four ops per view in a host-drawn order; the contrast op needs the mean luma of the view as it stands
// when the op runs, so every position is (optional) exact integer luma sum -> in-place per-pixel op --------------------------------
__global__ void luma_sum_kernel(const unsigned char* __restrict__ img, const int* __restrict__ ops,
                                unsigned long long* __restrict__ sums, int S, int pos) {
  const int m = blockIdx.y;
  if (ops[4 * m + pos] != JIT_CONTRAST) return;
  const long long n = (long long)S * S;
  const unsigned char* base = img + (long long)m * n * 3;
  unsigned long long acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += (unsigned long long)luma_byte(base[3 * i], base[3 * i + 1], base[3 * i + 2]);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ unsigned long long part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    atomicAdd(sums + m, t);
  }
}

__global__ void jitter_apply_kernel(unsigned char* __restrict__ img, const int* __restrict__ ops, const float* __restrict__ factors,
                                    const unsigned long long* __restrict__ sums, int S, int pos) {
  const int m = blockIdx.y;
  const int op = ops[4 * m + pos];
  if (op == JIT_NONE) return;
  const long long n = (long long)S * S;
  const float factor = factors[4 * m + pos];
  const int mean = op == JIT_CONTRAST ? contrast_mean(sums[m], n) : 0;
  unsigned char* base = img + (long long)m * n * 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    jitter_pixel(base + 3 * i, op, factor, mean);
}

// ---- GaussianBlur: per view a host-computed 8.8 tap row (or apply[m] == 0); two passes with a 16-bit intermediate ---------------
__global__ void blur_h_kernel(const unsigned char* __restrict__ img, const int* __restrict__ taps, const int* __restrict__ apply,
                              unsigned short* __restrict__ tmp16, int S, int ksize) {
  const int m = blockIdx.y;
  if (!apply[m]) return;
  const long long n = (long long)S * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    blur_h_pixel(img, taps + (long long)m * ksize, tmp16, m, (int)(i / S), (int)(i % S), S, ksize);
}

__global__ void blur_v_kernel(const unsigned short* __restrict__ tmp16, const int* __restrict__ taps, const int* __restrict__ apply,
                              unsigned char* __restrict__ img, int S, int ksize) {
  const int m = blockIdx.y;
  if (!apply[m]) return;
  const long long n = (long long)S * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    blur_v_pixel(tmp16, taps + (long long)m * ksize, img, m, (int)(i / S), (int)(i % S), S, ksize);
}

static inline int grid_for(long long total, int block) {
  long long blocks = (total + block - 1) / block;
  const long long cap = 148LL * 32;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

struct StageLayout {
  size_t status, bounds, taps, tmp, total;
};

static inline StageLayout stage_layout(int items, int S, int max_crop_h, int kmax) {
  StageLayout L;
  auto align = [](size_t v) { return (v + 255) / 256 * 256; };
  L.status = 0;
  L.bounds = 256;
  L.taps = align(L.bounds + (size_t)items * 2 * S * 2 * sizeof(int));
  L.tmp = align(L.taps + (size_t)items * 2 * S * kmax * sizeof(int));
  L.total = align(L.tmp + (size_t)items * max_crop_h * S * 3);
  return L;
}

}  // namespace pb

using namespace pb;

extern "C" int passl_b200_resample_kmax(int max_crop, int out_size, int interpolation) {
  if (max_crop <= 0 || out_size <= 0) return PB_ERR_BAD_ARG;
  double scale = (double)max_crop / out_size;
  if (scale < 1.0) scale = 1.0;
  return (int)ceil((interpolation ? 2.0 : 1.0) * scale) * 2 + 1;
}

extern "C" long long passl_b200_resized_crop_workspace_bytes(int items, int out_size, int max_crop_h, int kmax) {
  if (items <= 0 || out_size <= 0 || max_crop_h <= 0 || kmax <= 0) return 0;
  return (long long)stage_layout(items, out_size, max_crop_h, kmax).total;
}

extern "C" int passl_b200_resized_crop_u8(const void* src, const long long* src_off, const int* src_h, const int* src_w,
                                          const int* item_img, const int* item_box, void* dst, void* workspace,
                                          long long workspace_bytes, int items, int out_size, int max_crop_h, int kmax,
                                          int interpolation, void* stream) {
  if (items <= 0 || out_size <= 0 || max_crop_h <= 0 || kmax < 3 || (interpolation != 0 && interpolation != 1)) return PB_ERR_BAD_ARG;
  if (!src || !src_off || !src_h || !src_w || !item_img || !item_box || !dst || !workspace) return PB_ERR_BAD_ARG;
  const StageLayout L = stage_layout(items, out_size, max_crop_h, kmax);
  if (workspace_bytes < (long long)L.total) return PB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = reinterpret_cast<char*>(workspace);
  int* status = reinterpret_cast<int*>(ws + L.status);
  int* bounds = reinterpret_cast<int*>(ws + L.bounds);
  int* taps = reinterpret_cast<int*>(ws + L.taps);
  unsigned char* tmp = reinterpret_cast<unsigned char*>(ws + L.tmp);
  PB_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int), st));
  const long long n_coef = (long long)items * 2 * out_size;
  resample_coeffs_kernel<<<(unsigned)((n_coef + 127) / 128), 128, 0, st>>>(src_off, src_h, src_w, item_img, item_box, bounds, taps,
                                                                           status, items, out_size, kmax, interpolation);
  PB_LAUNCH_CHECK();
  resample_h_kernel<<<grid_for((long long)items * max_crop_h * out_size, 256), 256, 0, st>>>(
      reinterpret_cast<const unsigned char*>(src), src_off, src_h, src_w, item_img, item_box, bounds, taps, tmp, items, out_size,
      kmax, max_crop_h);
  PB_LAUNCH_CHECK();
  resample_v_kernel<<<grid_for((long long)items * out_size * out_size, 256), 256, 0, st>>>(
      tmp, src_off, src_h, src_w, item_img, item_box, bounds, taps, reinterpret_cast<unsigned char*>(dst), items, out_size, kmax,
      max_crop_h);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_views_finalize_f32(const void* img, const int* gray, const int* flip, float* out, int items, int size,
                                             double scale, const float* mean3, const float* std3, void* stream) {
  if (items <= 0 || size <= 0 || !img || !gray || !flip || !out || !mean3 || !std3) return PB_ERR_BAD_ARG;
  if (std3[0] == 0.f || std3[1] == 0.f || std3[2] == 0.f) return PB_ERR_BAD_ARG;
  views_finalize_kernel<<<grid_for((long long)items * size * size, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const unsigned char*>(img), gray, flip, out, items, size, scale, mean3[0], mean3[1], mean3[2], std3[0],
      std3[1], std3[2]);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int passl_b200_color_jitter_u8(void* img, const int* ops, const float* factors, void* workspace, long long workspace_bytes,
                                          int items, int size, int contrast_positions, void* stream) {
  if (items <= 0 || size <= 0 || !img || !ops || !factors || !workspace) return PB_ERR_BAD_ARG;
  if (workspace_bytes < (long long)items * (long long)sizeof(unsigned long long)) return PB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* sums = reinterpret_cast<unsigned long long*>(workspace);
  const long long n = (long long)size * size;
  int bx = (int)((n + 255) / 256);
  if (bx > 64) bx = 64;
  const dim3 grid(bx, items);
  for (int pos = 0; pos < 4; ++pos) {
    if (contrast_positions & (1 << pos)) {
      PB_CUDA_CHECK(cudaMemsetAsync(sums, 0, (size_t)items * sizeof(unsigned long long), st));
      luma_sum_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(img), ops, sums, size, pos);
      PB_LAUNCH_CHECK();
    }
    jitter_apply_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<unsigned char*>(img), ops, factors, sums, size, pos);
    PB_LAUNCH_CHECK();
  }
  return PB_OK;
}

extern "C" long long passl_b200_gaussian_blur_workspace_bytes(int items, int size) {
  if (items <= 0 || size <= 0) return 0;
  return (long long)items * size * size * 3 * (long long)sizeof(unsigned short);
}

extern "C" int passl_b200_gaussian_blur_u8(void* img, const int* taps, const int* apply, void* workspace, long long workspace_bytes,
                                           int items, int size, int ksize, void* stream) {
  if (items <= 0 || size <= 0 || ksize < 1 || !(ksize & 1) || !img || !taps || !apply || !workspace) return PB_ERR_BAD_ARG;
  if (workspace_bytes < passl_b200_gaussian_blur_workspace_bytes(items, size)) return PB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)size * size;
  int bx = (int)((n + 255) / 256);
  if (bx > 64) bx = 64;
  const dim3 grid(bx, items);
  unsigned short* tmp16 = reinterpret_cast<unsigned short*>(workspace);
  blur_h_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(img), taps, apply, tmp16, size, ksize);
  PB_LAUNCH_CHECK();
  blur_v_kernel<<<grid, 256, 0, st>>>(tmp16, taps, apply, reinterpret_cast<unsigned char*>(img), size, ksize);
  PB_LAUNCH_CHECK();
  return PB_OK;
}
