// Per-element arithmetic of the image input stage (input_stage.cu), written so that the SAME source also compiles as plain host C++:
// tests/test_input_stage_host_cpu.py builds it with g++ (-ffp-contract=off) and checks it against Pillow, which validates the
// arithmetic and indexing of the kernels' bodies without a GPU.  On the device every double operation is an explicitly rounded
// intrinsic (never contracted into FMA), so both builds produce the integers Pillow's libImaging/Resample.c produces.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline
#endif

namespace pb {
namespace istage {

constexpr int kPrecisionBits = 32 - 8 - 2;

#if defined(__CUDA_ARCH__)
PB_HD double rn_add(double a, double b) { return __dadd_rn(a, b); }
PB_HD double rn_sub(double a, double b) { return __dsub_rn(a, b); }
PB_HD double rn_mul(double a, double b) { return __dmul_rn(a, b); }
PB_HD double rn_div(double a, double b) { return __ddiv_rn(a, b); }
PB_HD int trunc_int(double a) { return __double2int_rz(a); }
PB_HD float to_float(double a) { return __double2float_rn(a); }
#else
PB_HD double rn_add(double a, double b) { return a + b; }
PB_HD double rn_sub(double a, double b) { return a - b; }
PB_HD double rn_mul(double a, double b) { return a * b; }
PB_HD double rn_div(double a, double b) { return a / b; }
PB_HD int trunc_int(double a) { return (int)a; }
PB_HD float to_float(double a) { return (float)a; }
#endif

struct ItemGeom {
  int top, left, ch, cw;      // crop box inside the source image
  int src_h, src_w;
  long long src_off;          // byte offset of the image in the packed source buffer
};

PB_HD ItemGeom item_geom(const long long* src_off, const int* src_h, const int* src_w, const int* item_img, const int* item_box,
                         int m) {
  ItemGeom g;
  const int n = item_img[m];
  g.src_off = src_off[n];
  g.src_h = src_h[n];
  g.src_w = src_w[n];
  g.top = item_box[4 * m + 0];
  g.left = item_box[4 * m + 1];
  g.ch = item_box[4 * m + 2];
  g.cw = item_box[4 * m + 3];
  return g;
}

PB_HD bool geom_ok(const ItemGeom& g) {
  return g.ch > 0 && g.cw > 0 && g.top >= 0 && g.left >= 0 && g.top + g.ch <= g.src_h && g.left + g.cw <= g.src_w;
}

// filter(x): bilinear triangle (support 1) or Pillow's bicubic with a = -0.5 (support 2); every operation rounded on its own
PB_HD double resample_filter(double x, int bicubic) {
  if (x < 0.0) x = -x;
  if (!bicubic) return x < 1.0 ? rn_sub(1.0, x) : 0.0;
  if (x < 1.0) {                                                  // ((a + 2) x - (a + 3)) x x + 1
    double t = rn_sub(rn_mul(1.5, x), 2.5);
    t = rn_mul(rn_mul(t, x), x);
    return rn_add(t, 1.0);
  }
  if (x < 2.0) {                                                  // (((x - 5) x + 8) x - 4) a
    double t = rn_add(rn_mul(rn_sub(x, 5.0), x), 8.0);
    t = rn_sub(rn_mul(t, x), 4.0);
    return rn_mul(t, -0.5);
  }
  return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for ONE output index xx of a box (0, in_size) -> S.
// b[0] = first source index, b[1] = tap count, k[0..kmax) fixed-point taps.  Returns 0, or 2 when kmax is too small.
PB_HD int resample_window(int in_size, int S, int xx, int bicubic, int kmax, int* b, int* k) {
  const double scale = rn_div((double)in_size, (double)S);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = rn_mul(bicubic ? 2.0 : 1.0, filterscale);
  const int ksize = (int)ceil(support) * 2 + 1;
  if (ksize > kmax) {
    b[0] = 0;
    b[1] = 0;
    return 2;
  }
  const double center = rn_mul((double)xx + 0.5, scale);           // in0 = 0
  const double ss = rn_div(1.0, filterscale);
  int xmin = trunc_int(rn_add(rn_sub(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = trunc_int(rn_add(rn_add(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    const double arg = rn_mul(rn_add(rn_sub((double)(x + xmin), center), 0.5), ss);
    ww = rn_add(ww, resample_filter(arg, bicubic));
  }
  for (int x = 0; x < xmax; ++x) {                                 // recomputed instead of stored: no double scratch needed
    const double arg = rn_mul(rn_add(rn_sub((double)(x + xmin), center), 0.5), ss);
    double w = resample_filter(arg, bicubic);
    if (ww != 0.0) w = rn_div(w, ww);
    const double scaled = rn_mul(w, (double)(1 << kPrecisionBits));
    k[x] = trunc_int(w < 0.0 ? rn_add(-0.5, scaled) : rn_add(0.5, scaled));
  }
  for (int x = xmax; x < kmax; ++x) k[x] = 0;
  b[0] = xmin;
  b[1] = xmax;
  return 0;
}

PB_HD unsigned char clip8(int acc) {
  const int v = acc >> kPrecisionBits;
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// table layout shared by the three resample kernels: bounds [items][2 axes][S][2], taps [items][2][S][kmax]
PB_HD long long window_index(int m, int axis, int S, int xx) { return (long long)(m * 2 + axis) * S + xx; }

// horizontal pass, one output pixel: tmp[m][y][xx][0..3) from row (top + y) of the crop
PB_HD void h_pass_pixel(const unsigned char* src, const ItemGeom& g, const int* bounds, const int* taps, unsigned char* tmp, int m,
                        int y, int xx, int S, int kmax, int max_crop_h) {
  const int* b = bounds + window_index(m, 0, S, xx) * 2;
  const int* k = taps + window_index(m, 0, S, xx) * kmax;
  const int x0 = b[0], cnt = b[1];
  const unsigned char* row = src + g.src_off + ((long long)(g.top + y) * g.src_w + g.left + x0) * 3;
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  for (int x = 0; x < cnt; ++x) {
    const int w = k[x];
    a0 += (int)row[3 * x + 0] * w;
    a1 += (int)row[3 * x + 1] * w;
    a2 += (int)row[3 * x + 2] * w;
  }
  unsigned char* o = tmp + (((long long)m * max_crop_h + y) * S + xx) * 3;
  o[0] = clip8(a0);
  o[1] = clip8(a1);
  o[2] = clip8(a2);
}

// vertical pass, one output pixel: dst[m][yy][x][0..3) from column x of the intermediate
PB_HD void v_pass_pixel(const unsigned char* tmp, const int* bounds, const int* taps, unsigned char* dst, int m, int yy, int x, int S,
                        int kmax, int max_crop_h) {
  const int* b = bounds + window_index(m, 1, S, yy) * 2;
  const int* k = taps + window_index(m, 1, S, yy) * kmax;
  const int y0 = b[0], cnt = b[1];
  const unsigned char* col = tmp + (((long long)m * max_crop_h + y0) * S + x) * 3;
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  for (int y = 0; y < cnt; ++y) {
    const int w = k[y];
    const unsigned char* p = col + (long long)y * S * 3;
    a0 += (int)p[0] * w;
    a1 += (int)p[1] * w;
    a2 += (int)p[2] * w;
  }
  unsigned char* o = dst + (((long long)m * S + yy) * S + x) * 3;
  o[0] = clip8(a0);
  o[1] = clip8(a1);
  o[2] = clip8(a2);
}

// NormalizeImage table entry: float32((v * scale - mean) / std) with the arithmetic in double (transforms.py:462-467)
PB_HD float normalize_entry(int v, double scale, float mean, float stdv) {
  return to_float(rn_div(rn_sub(rn_mul((double)v, scale), (double)mean), (double)stdv));
}

// finalize, one output pixel: grayscale (Convert.c rgb2l) where gray, mirrored source column where flip, CHW planes from `lut`
PB_HD void finalize_pixel(const unsigned char* img, const float* lut, float* out, int m, int y, int x, int S, int gray, int flip) {
  const int xs = flip ? S - 1 - x : x;
  const unsigned char* p = img + (((long long)m * S + y) * S + xs) * 3;
  int r = p[0], g = p[1], b = p[2];
  if (gray) r = g = b = (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16;
  float* o = out + (long long)m * 3 * S * S + (long long)y * S + x;
  o[0] = lut[r];
  o[(long long)S * S] = lut[256 + g];
  o[2LL * S * S] = lut[512 + b];
}

// ---- ColorJitter (configs/simclr/simclr_r50_IM.yaml:41-48): Pillow's ImageEnhance blends and HSV round trip --------------------
#if defined(__CUDA_ARCH__)
PB_HD float rnf_add(float a, float b) { return __fadd_rn(a, b); }
PB_HD float rnf_mul(float a, float b) { return __fmul_rn(a, b); }
PB_HD float rnf_div(float a, float b) { return __fdiv_rn(a, b); }
#else
PB_HD float rnf_add(float a, float b) { return a + b; }
PB_HD float rnf_mul(float a, float b) { return a * b; }
PB_HD float rnf_div(float a, float b) { return a / b; }
#endif

enum JitterOp : int { JIT_NONE = 0, JIT_BRIGHTNESS = 1, JIT_CONTRAST = 2, JIT_SATURATION = 3, JIT_HUE = 4, JIT_GRAY = 5 };

// Blend.c: out = a + alpha (b - a) in float32, truncated; clipped when alpha lies outside [0, 1] (a = degenerate, b = image)
PB_HD unsigned char blend_byte(int a, int b, float alpha) {
  const float t = rnf_add((float)a, rnf_mul(alpha, (float)(b - a)));
  if (alpha >= 0.f && alpha <= 1.f) return (unsigned char)(int)t;
  if (t <= 0.f) return 0;
  if (t >= 255.f) return 255;
  return (unsigned char)(int)t;
}

PB_HD int luma_byte(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// ImageStat mean of the L plane rounded half up: the grey level ImageEnhance.Contrast blends towards
PB_HD int contrast_mean(unsigned long long luma_sum, long long count) {
  return (int)(rn_add(rn_div((double)luma_sum, (double)count), 0.5));
}

// Convert.c rgb2hsv: ratios in float32, hue assembled in double, rounded to float32, scaled by 255.0 in double
PB_HD void rgb_to_hsv(int r, int g, int b, int* h, int* s, int* v) {
  const int maxc = r > g ? (r > b ? r : b) : (g > b ? g : b);
  const int minc = r < g ? (r < b ? r : b) : (g < b ? g : b);
  *v = maxc;
  if (minc == maxc) {
    *h = 0;
    *s = 0;
    return;
  }
  const float cr = (float)(maxc - minc);
  const float sf = rnf_div(cr, (float)maxc);
  const float rc = rnf_div((float)(maxc - r), cr), gc = rnf_div((float)(maxc - g), cr), bc = rnf_div((float)(maxc - b), cr);
  double hd;
  if (r == maxc) hd = rn_sub((double)bc, (double)gc);
  else if (g == maxc) hd = rn_sub(rn_add(2.0, (double)rc), (double)bc);
  else hd = rn_sub(rn_add(4.0, (double)gc), (double)rc);
  const float hf = to_float(hd);
  const double turn = rn_add(rn_div((double)hf, 6.0), 1.0);        // in [5/6, 11/6]: fmod(x, 1.0) = x - floor(x), exact
  const float hfrac = to_float(rn_sub(turn, floor(turn)));
  int uh = trunc_int(rn_mul((double)hfrac, 255.0));
  int us = trunc_int(rn_mul((double)sf, 255.0));
  *h = uh < 0 ? 0 : (uh > 255 ? 255 : uh);
  *s = us < 0 ? 0 : (us > 255 ? 255 : us);
}

PB_HD int round_clip8(double x) {                                  // C round() of a non-negative value, then CLIP8
  const int v = (int)floor(rn_add(x, 0.5));
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Convert.c hsv2rgb (colorsys sextants)
PB_HD void hsv_to_rgb(int h, int s, int v, int* r, int* g, int* b) {
  if (s == 0) {
    *r = *g = *b = v;
    return;
  }
  const double hd = rn_div(rn_mul((double)h, 6.0), 255.0);
  const int i = (int)floor(hd);
  const double f = (double)to_float(rn_sub(hd, (double)i));
  const double fs = (double)to_float(rn_div((double)s, 255.0));
  const int p = round_clip8(rn_mul((double)v, rn_sub(1.0, fs)));
  const int q = round_clip8(rn_mul((double)v, rn_sub(1.0, rn_mul(fs, f))));
  const int t = round_clip8(rn_mul((double)v, rn_sub(1.0, rn_mul(fs, rn_sub(1.0, f)))));
  switch (i % 6) {
    case 0: *r = v; *g = t; *b = p; break;
    case 1: *r = q; *g = v; *b = p; break;
    case 2: *r = p; *g = v; *b = t; break;
    case 3: *r = p; *g = q; *b = v; break;
    case 4: *r = t; *g = p; *b = v; break;
    default: *r = v; *g = p; *b = q; break;
  }
}

// one pixel of one jitter op, in place.  `factor` is the blend factor (a C float, as Pillow receives it); for JIT_HUE it carries the
// H-plane shift paddle's adjust_hue computes on the host, np.uint8(hue_factor * 255) in 0..255 (double product, truncation toward
// zero, wrap).  `mean` = contrast_mean of the view as it stands before this op.
PB_HD void jitter_pixel(unsigned char* p, int op, float factor, int mean) {
  int r = p[0], g = p[1], b = p[2];
  if (op == JIT_BRIGHTNESS) {
    p[0] = blend_byte(0, r, factor); p[1] = blend_byte(0, g, factor); p[2] = blend_byte(0, b, factor);
  } else if (op == JIT_CONTRAST) {
    p[0] = blend_byte(mean, r, factor); p[1] = blend_byte(mean, g, factor); p[2] = blend_byte(mean, b, factor);
  } else if (op == JIT_SATURATION) {
    const int L = luma_byte(r, g, b);
    p[0] = blend_byte(L, r, factor); p[1] = blend_byte(L, g, factor); p[2] = blend_byte(L, b, factor);
  } else if (op == JIT_HUE) {
    int h, s, v;
    rgb_to_hsv(r, g, b, &h, &s, &v);
    h = (h + (int)factor) & 255;
    hsv_to_rgb(h, s, v, &r, &g, &b);
    p[0] = (unsigned char)r; p[1] = (unsigned char)g; p[2] = (unsigned char)b;
  } else if (op == JIT_GRAY) {                                    // RandomGrayscale when it has to run before the blur
    p[0] = p[1] = p[2] = (unsigned char)luma_byte(r, g, b);
  }
}

// ---- GaussianBlur (transforms.py:173-191 -> cv2.GaussianBlur on uint8): OpenCV's fixed-point filter.  Taps are 8.8 fixed point
// (sum exactly 256, computed on the host), the horizontal pass keeps unrounded 8.8 values (<= 255 * 256 fits 16 bits), the vertical
// pass rounds once: (sum + 2^15) >> 16.  Borders: BORDER_REFLECT_101 (cv::borderInterpolate).
PB_HD int reflect101(int p, int len) {
  if (len == 1) return 0;
  while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * len - 2 - p;
  return p;
}

// horizontal pass, one pixel: tmp16[m][y][x][c] = sum_i taps[i] * img[m][y][reflect(x + i - r)][c]
PB_HD void blur_h_pixel(const unsigned char* img, const int* taps, unsigned short* tmp16, int m, int y, int x, int S, int ksize) {
  const int r = ksize / 2;
  const unsigned char* row = img + ((long long)m * S + y) * S * 3;
  int a0 = 0, a1 = 0, a2 = 0;
  for (int i = 0; i < ksize; ++i) {
    const unsigned char* p = row + 3 * reflect101(x + i - r, S);
    const int w = taps[i];
    a0 += w * p[0];
    a1 += w * p[1];
    a2 += w * p[2];
  }
  unsigned short* o = tmp16 + (((long long)m * S + y) * S + x) * 3;
  o[0] = (unsigned short)a0;
  o[1] = (unsigned short)a1;
  o[2] = (unsigned short)a2;
}

// vertical pass, one pixel: img[m][y][x][c] = (sum_j taps[j] * tmp16[m][reflect(y + j - r)][x][c] + 2^15) >> 16
PB_HD void blur_v_pixel(const unsigned short* tmp16, const int* taps, unsigned char* img, int m, int y, int x, int S, int ksize) {
  const int r = ksize / 2;
  const unsigned short* base = tmp16 + (long long)m * S * S * 3 + (long long)x * 3;
  unsigned int a0 = 0, a1 = 0, a2 = 0;
  for (int j = 0; j < ksize; ++j) {
    const unsigned short* p = base + (long long)reflect101(y + j - r, S) * S * 3;
    const unsigned int w = (unsigned int)taps[j];
    a0 += w * p[0];
    a1 += w * p[1];
    a2 += w * p[2];
  }
  unsigned char* o = img + (((long long)m * S + y) * S + x) * 3;
  o[0] = (unsigned char)((a0 + (1u << 15)) >> 16);
  o[1] = (unsigned char)((a1 + (1u << 15)) >> 16);
  o[2] = (unsigned char)((a2 + (1u << 15)) >> 16);
}

}  // namespace istage
}  // namespace pb
