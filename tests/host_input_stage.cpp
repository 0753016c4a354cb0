// Host build of the image input stage's per-element bodies (passl_b200/csrc/input_stage_core.h) — TEST INFRASTRUCTURE.
// The loops below stand where the CUDA kernels' thread indexing stands (input_stage.cu); everything inside them is the shared
// source.  Built by tests/test_input_stage_host_cpu.py with g++ -O2 -ffp-contract=off and compared with Pillow / the oracle.
#include <vector>

#include "../passl_b200/csrc/input_stage_core.h"

using namespace pb::istage;

extern "C" int host_resized_crop_u8(const unsigned char* src, const long long* src_off, const int* src_h, const int* src_w,
                                    const int* item_img, const int* item_box, unsigned char* dst, int items, int S, int max_crop_h,
                                    int kmax, int bicubic) {
  std::vector<int> bounds((size_t)items * 2 * S * 2), taps((size_t)items * 2 * S * kmax);
  std::vector<unsigned char> tmp((size_t)items * max_crop_h * S * 3);
  int status = 0;
  for (int m = 0; m < items; ++m) {
    const ItemGeom g = item_geom(src_off, src_h, src_w, item_img, item_box, m);
    for (int axis = 0; axis < 2; ++axis)
      for (int xx = 0; xx < S; ++xx) {
        int* b = bounds.data() + window_index(m, axis, S, xx) * 2;
        int* k = taps.data() + window_index(m, axis, S, xx) * kmax;
        if (!geom_ok(g)) { b[0] = b[1] = 0; status |= 1; continue; }
        status |= resample_window(axis == 0 ? g.cw : g.ch, S, xx, bicubic, kmax, b, k);
      }
  }
  for (int m = 0; m < items; ++m) {
    const ItemGeom g = item_geom(src_off, src_h, src_w, item_img, item_box, m);
    for (int y = 0; y < max_crop_h; ++y)
      for (int xx = 0; xx < S; ++xx) {
        if (!geom_ok(g) || y >= g.ch) continue;
        h_pass_pixel(src, g, bounds.data(), taps.data(), tmp.data(), m, y, xx, S, kmax, max_crop_h);
      }
  }
  for (int m = 0; m < items; ++m) {
    const ItemGeom g = item_geom(src_off, src_h, src_w, item_img, item_box, m);
    for (int yy = 0; yy < S; ++yy)
      for (int x = 0; x < S; ++x) {
        const int* b = bounds.data() + window_index(m, 1, S, yy) * 2;
        if (!geom_ok(g) || b[1] == 0) {
          unsigned char* o = dst + (((long long)m * S + yy) * S + x) * 3;
          o[0] = o[1] = o[2] = 0;
          continue;
        }
        v_pass_pixel(tmp.data(), bounds.data(), taps.data(), dst, m, yy, x, S, kmax, max_crop_h);
      }
  }
  return status;
}

extern "C" void host_views_finalize_f32(const unsigned char* img, const int* gray, const int* flip, float* out, int items, int S,
                                        double scale, const float* mean3, const float* std3) {
  float lut[3 * 256];
  for (int i = 0; i < 3 * 256; ++i) lut[i] = normalize_entry(i & 255, scale, mean3[i >> 8], std3[i >> 8]);
  for (int m = 0; m < items; ++m)
    for (int y = 0; y < S; ++y)
      for (int x = 0; x < S; ++x) finalize_pixel(img, lut, out, m, y, x, S, gray[m], flip[m]);
}

// colour jitter: ops / factors [items][4]; per position a luma sum over the view (only used by the contrast op), then the op
extern "C" void host_color_jitter_u8(unsigned char* img, const int* ops, const float* factors, int items, int S) {
  for (int pos = 0; pos < 4; ++pos)
    for (int m = 0; m < items; ++m) {
      const int op = ops[4 * m + pos];
      if (op == JIT_NONE) continue;
      unsigned char* base = img + (long long)m * S * S * 3;
      unsigned long long sum = 0;
      for (long long i = 0; i < (long long)S * S; ++i) sum += (unsigned long long)luma_byte(base[3 * i], base[3 * i + 1], base[3 * i + 2]);
      const int mean = contrast_mean(sum, (long long)S * S);
      for (long long i = 0; i < (long long)S * S; ++i) jitter_pixel(base + 3 * i, op, factors[4 * m + pos], mean);
    }
}

extern "C" void host_rgb_to_hsv(const unsigned char* in, unsigned char* out, long long n) {
  for (long long i = 0; i < n; ++i) {
    int h, s, v;
    rgb_to_hsv(in[3 * i], in[3 * i + 1], in[3 * i + 2], &h, &s, &v);
    out[3 * i] = (unsigned char)h; out[3 * i + 1] = (unsigned char)s; out[3 * i + 2] = (unsigned char)v;
  }
}

extern "C" void host_hsv_to_rgb(const unsigned char* in, unsigned char* out, long long n) {
  for (long long i = 0; i < n; ++i) {
    int r, g, b;
    hsv_to_rgb(in[3 * i], in[3 * i + 1], in[3 * i + 2], &r, &g, &b);
    out[3 * i] = (unsigned char)r; out[3 * i + 1] = (unsigned char)g; out[3 * i + 2] = (unsigned char)b;
  }
}

extern "C" void host_gaussian_blur_u8(unsigned char* img, const int* taps, const int* apply, int items, int S, int ksize) {
  std::vector<unsigned short> tmp16((size_t)items * S * S * 3);
  for (int m = 0; m < items; ++m) {
    if (!apply[m]) continue;
    for (int y = 0; y < S; ++y)
      for (int x = 0; x < S; ++x) blur_h_pixel(img, taps + (long long)m * ksize, tmp16.data(), m, y, x, S, ksize);
  }
  for (int m = 0; m < items; ++m) {
    if (!apply[m]) continue;
    for (int y = 0; y < S; ++y)
      for (int x = 0; x < S; ++x) blur_v_pixel(tmp16.data(), taps + (long long)m * ksize, img, m, y, x, S, ksize);
  }
}

// ---- the same entry points as the CUDA library, by their ABI names and signatures (checked against include/passl_b200.h by the
// compiler), running the shared bodies on host memory: tests/test_input_stage_host_cpu.py points the Python wrappers of
// passl_b200/data at this library to exercise their argument marshalling and the stage's orchestration without a GPU ------------
#include <math.h>
#include <string.h>

#include "../include/passl_b200.h"

extern "C" int passl_b200_resample_kmax(int max_crop, int out_size, int interpolation) {
  if (max_crop <= 0 || out_size <= 0) return -1;
  double scale = (double)max_crop / out_size;
  if (scale < 1.0) scale = 1.0;
  return (int)ceil((interpolation ? 2.0 : 1.0) * scale) * 2 + 1;
}

extern "C" long long passl_b200_resized_crop_workspace_bytes(int items, int out_size, int max_crop_h, int kmax) {
  return (items <= 0 || out_size <= 0 || max_crop_h <= 0 || kmax <= 0) ? 0 : 256;
}

extern "C" int passl_b200_resized_crop_u8(const void* src, const long long* src_off, const int* src_h, const int* src_w,
                                          const int* item_img, const int* item_box, void* dst, void* workspace,
                                          long long workspace_bytes, int items, int out_size, int max_crop_h, int kmax,
                                          int interpolation, void* stream) {
  (void)stream;
  if (items <= 0 || out_size <= 0 || max_crop_h <= 0 || kmax < 3 || (interpolation != 0 && interpolation != 1)) return -1;
  if (workspace_bytes < 256) return -4;
  const int status = host_resized_crop_u8((const unsigned char*)src, src_off, src_h, src_w, item_img, item_box, (unsigned char*)dst,
                                          items, out_size, max_crop_h, kmax, interpolation);
  memcpy(workspace, &status, sizeof(int));
  return 0;
}

extern "C" int passl_b200_views_finalize_f32(const void* img, const int* gray, const int* flip, float* out, int items, int size,
                                             double scale, const float* mean3, const float* std3, void* stream) {
  (void)stream;
  if (items <= 0 || size <= 0) return -1;
  host_views_finalize_f32((const unsigned char*)img, gray, flip, out, items, size, scale, mean3, std3);
  return 0;
}

extern "C" int passl_b200_color_jitter_u8(void* img, const int* ops, const float* factors, void* workspace, long long workspace_bytes,
                                          int items, int size, int contrast_positions, void* stream) {
  (void)stream; (void)workspace;
  if (items <= 0 || size <= 0) return -1;
  if (workspace_bytes < 8LL * items) return -4;
  for (int m = 0; m < items; ++m)                              // the host must have announced every contrast position
    for (int pos = 0; pos < 4; ++pos)
      if (ops[4 * m + pos] == JIT_CONTRAST && !(contrast_positions & (1 << pos))) return -1;
  host_color_jitter_u8((unsigned char*)img, ops, factors, items, size);
  return 0;
}

extern "C" long long passl_b200_gaussian_blur_workspace_bytes(int items, int size) {
  return (items <= 0 || size <= 0) ? 0 : (long long)items * size * size * 3 * 2;
}

extern "C" int passl_b200_gaussian_blur_u8(void* img, const int* taps, const int* apply, void* workspace, long long workspace_bytes,
                                           int items, int size, int ksize, void* stream) {
  (void)stream; (void)workspace;
  if (items <= 0 || size <= 0 || ksize < 1 || !(ksize & 1)) return -1;
  if (workspace_bytes < passl_b200_gaussian_blur_workspace_bytes(items, size)) return -4;
  host_gaussian_blur_u8((unsigned char*)img, taps, apply, items, size, ksize);
  return 0;
}
