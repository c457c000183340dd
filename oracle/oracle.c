/*
 * oracle.c -- CPU restatement of go-webp's pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (go_webp_b200/,
 * include/) may link, import or call this file.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * use it, and only as the checker / the reported CPU baseline.
 *
 * Parity status: PINNED against the libwebp 1.6.0 binary the reference
 * transliterates (the reference's own tests hold no vectors for this path:
 * pkg/generics/sizeof_test.go is its only test).  Parts 1 and 2: the binary's
 * import / upsample / decode entry points (oracle/libwebp_ref.py,
 * tests/golden/make_golden.py).  Part 3: residual images and transform data cut
 * out of lossless streams the binary wrote, checked against what the binary
 * decodes them to (tests/golden/make_golden_vp8l.py, vp8l_libwebp160.npz:
 * predictor with all 14 modes, cross-colour, subtract-green, colour indexing at
 * 1/2/4/8 bits); the forward transforms are pinned as exact inverses on the same
 * vectors.  See DESIGN.md "Oracle".
 *
 * Every function cites the reference file:line (relative to the go-webp tree)
 * whose arithmetic it restates.  Where the Go text is ill-typed the intended C
 * semantics are implemented.
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* pkg/color/yuv/constants.go:3-9                                            */
enum {
  YUV_FIX = 16,
  YUV_HALF = 1 << (YUV_FIX - 1),
  YUV_FIX2 = 6,
  YUV_MASK2 = (256 << YUV_FIX2) - 1
};

/* pkg/color/yuv/conversion.go:41-48 */
static inline int MultHi(int v, int coeff) { return (v * coeff) >> 8; }
static inline int Clip8(int v) {
  return ((v & ~YUV_MASK2) == 0) ? (v >> YUV_FIX2) : (v < 0) ? 0 : 255;
}
/* pkg/color/yuv/conversion.go:28-38 */
int orc_YUVToR(int y, int v) {
  return Clip8(MultHi(y, 19077) + MultHi(v, 26149) - 14234);
}
int orc_YUVToG(int y, int u, int v) {
  return Clip8(MultHi(y, 19077) - MultHi(u, 6419) - MultHi(v, 13320) + 8708);
}
int orc_YUVToB(int y, int u) {
  return Clip8(MultHi(y, 19077) + MultHi(u, 33050) - 17685);
}
/* pkg/color/yuv/conversion.go:51-69 */
int orc_ClipUV(int uv, int rounding) {
  uv = (uv + rounding + (128 << (YUV_FIX + 2))) >> (YUV_FIX + 2);
  return ((uv & ~0xff) == 0) ? uv : (uv < 0) ? 0 : 255;
}
int orc_RGBToY(int r, int g, int b, int rounding) {
  const int luma = 16839 * r + 33059 * g + 6420 * b;
  return (luma + rounding + (16 << YUV_FIX)) >> YUV_FIX;
}
int orc_RGBToU(int r, int g, int b, int rounding) {
  return orc_ClipUV(-9719 * r - 19081 * g + 28800 * b, rounding);
}
int orc_RGBToV(int r, int g, int b, int rounding) {
  return orc_ClipUV(28800 * r - 24116 * g - 4684 * b, rounding);
}

/* ------------------------------------------------------------------------- */
/* Gamma tables: pkg/libwebp/dsp/yuv.go:155-181 (C-intended float division,   */
/* SURVEY Appendix B-1).                                                      */
enum {
  GAMMA_FIX = 12,
  GAMMA_TAB_FIX = 7,
  GAMMA_TAB_SIZE = 1 << (GAMMA_FIX - GAMMA_TAB_FIX),
  kGammaScale = (1 << GAMMA_FIX) - 1,
  kGammaTabScale = 1 << GAMMA_TAB_FIX,
  kGammaTabRounder = (1 << GAMMA_TAB_FIX) >> 1
};
static const double kGamma = 0.80;
static int kLinearToGammaTab[GAMMA_TAB_SIZE + 1];
static uint16_t kGammaToLinearTab[256];
/* libwebp-binary divide mode only (USE_INVERSE_ALPHA_TABLE). */
enum { kAlphaFix = 19 };
static uint32_t kInvAlpha[4 * 0xff + 1];
static volatile int kGammaTablesOk = 0;
static pthread_mutex_t kInitLock = PTHREAD_MUTEX_INITIALIZER;

void orc_init(void) {
  pthread_mutex_lock(&kInitLock);
  if (!kGammaTablesOk) {
    int v;
    const double scale = (double)(1 << GAMMA_TAB_FIX) / kGammaScale;
    const double norm = 1. / 255.;
    for (v = 0; v <= 255; ++v) {
      kGammaToLinearTab[v] =
          (uint16_t)(pow(norm * v, kGamma) * kGammaScale + .5);
    }
    for (v = 0; v <= GAMMA_TAB_SIZE; ++v) {
      kLinearToGammaTab[v] = (int)(255. * pow(scale * v, 1. / kGamma) + .5);
    }
    kInvAlpha[0] = 0;
    for (v = 1; v <= 4 * 0xff; ++v) kInvAlpha[v] = (1u << kAlphaFix) / v;
    kGammaTablesOk = 1;
  }
  pthread_mutex_unlock(&kInitLock);
}

void orc_gamma_tables(uint16_t* g2l /*[256]*/, int32_t* l2g /*[33]*/) {
  int i;
  orc_init();
  for (i = 0; i < 256; ++i) g2l[i] = kGammaToLinearTab[i];
  for (i = 0; i <= GAMMA_TAB_SIZE; ++i) l2g[i] = kLinearToGammaTab[i];
}

/* yuv.go:183-202 */
static inline uint32_t GammaToLinear(uint8_t v) { return kGammaToLinearTab[v]; }
static inline int Interpolate(int v) {
  const int tab_pos = v >> (GAMMA_TAB_FIX + 2);
  const int x = v & ((kGammaTabScale << 2) - 1);
  const int v0 = kLinearToGammaTab[tab_pos];
  const int v1 = kLinearToGammaTab[tab_pos + 1];
  return v1 * x + v0 * ((kGammaTabScale << 2) - x);
}
static inline int LinearToGamma(uint32_t base_value, int shift) {
  const int y = Interpolate((int)(base_value << shift));
  return (y + kGammaTabRounder) >> GAMMA_TAB_FIX;
}
int orc_LinearToGamma(uint32_t base_value, int shift) {
  orc_init();
  return LinearToGamma(base_value, shift);
}

/* Divide modes for DIVIDE_BY_ALPHA (yuv.go:231): 0 = exact integer division   */
/* as the reference states it (graded); 1 = libwebp-binary inverse table.      */
enum { ORC_DIV_EXACT = 0, ORC_DIV_TABLE = 1 };
static inline uint32_t DivideByAlpha(uint32_t sum, uint32_t a, int mode) {
  if (mode == ORC_DIV_TABLE) return (sum * kInvAlpha[a]) >> (kAlphaFix - 2);
  return 4u * sum / a;
}

/* yuv.go:204-241.  'ptr' is one channel inside interleaved pixels. */
static inline int Sum4(const uint8_t* ptr, int step, int rgb_stride) {
  return LinearToGamma(GammaToLinear(ptr[0]) + GammaToLinear(ptr[step]) +
                           GammaToLinear(ptr[rgb_stride]) +
                           GammaToLinear(ptr[rgb_stride + step]),
                       0);
}
static inline int Sum2(const uint8_t* ptr, int rgb_stride) {
  return LinearToGamma(GammaToLinear(ptr[0]) + GammaToLinear(ptr[rgb_stride]),
                       1);
}
static inline uint32_t Sum2Alpha(const uint8_t* ptr, int rgb_stride) {
  return (uint32_t)ptr[0] + ptr[rgb_stride];
}
static inline uint32_t Sum4Alpha(const uint8_t* ptr, int rgb_stride) {
  return Sum2Alpha(ptr, rgb_stride) + Sum2Alpha(ptr + 4, rgb_stride);
}
static inline int LinearToGammaWeighted(const uint8_t* src,
                                        const uint8_t* a_ptr, uint32_t total_a,
                                        int step, int rgb_stride, int mode) {
  const uint32_t sum =
      a_ptr[0] * GammaToLinear(src[0]) +
      a_ptr[step] * GammaToLinear(src[step]) +
      a_ptr[rgb_stride] * GammaToLinear(src[rgb_stride]) +
      a_ptr[rgb_stride + step] * GammaToLinear(src[rgb_stride + step]);
  return LinearToGamma(DivideByAlpha(sum, total_a, mode), 0);
}

/* yuv.go:124-140 */
static void ConvertRGBToY(const uint8_t* r, const uint8_t* g, const uint8_t* b,
                          uint8_t* y, int width, int step) {
  int i;
  for (i = 0; i < width; ++i, r += step, g += step, b += step) {
    y[i] = (uint8_t)orc_RGBToY(r[0], g[0], b[0], YUV_HALF);
  }
}
/* yuv.go:142-153 */
static void ConvertRGBA32ToUV(const uint16_t* rgb, uint8_t* u, uint8_t* v,
                              int width) {
  int i;
  for (i = 0; i < width; ++i, rgb += 4) {
    const int r = rgb[0], g = rgb[1], b = rgb[2];
    u[i] = (uint8_t)orc_RGBToU(r, g, b, YUV_HALF << 2);
    v[i] = (uint8_t)orc_RGBToV(r, g, b, YUV_HALF << 2);
  }
}
/* yuv.go:243-284 */
static void AccumulateRGBA(const uint8_t* r_ptr, const uint8_t* g_ptr,
                           const uint8_t* b_ptr, const uint8_t* a_ptr,
                           int rgb_stride, uint16_t* dst, int width, int mode) {
  int i, j;
  for (i = 0, j = 0; i < (width >> 1); i += 1, j += 2 * 4, dst += 4) {
    const uint32_t a = Sum4Alpha(a_ptr + j, rgb_stride);
    int r, g, b;
    if (a == 4 * 0xff || a == 0) {
      r = Sum4(r_ptr + j, 4, rgb_stride);
      g = Sum4(g_ptr + j, 4, rgb_stride);
      b = Sum4(b_ptr + j, 4, rgb_stride);
    } else {
      r = LinearToGammaWeighted(r_ptr + j, a_ptr + j, a, 4, rgb_stride, mode);
      g = LinearToGammaWeighted(g_ptr + j, a_ptr + j, a, 4, rgb_stride, mode);
      b = LinearToGammaWeighted(b_ptr + j, a_ptr + j, a, 4, rgb_stride, mode);
    }
    dst[0] = (uint16_t)r; dst[1] = (uint16_t)g; dst[2] = (uint16_t)b;
    dst[3] = (uint16_t)a;
  }
  if (width & 1) {
    const uint32_t a = 2u * Sum2Alpha(a_ptr + j, rgb_stride);
    int r, g, b;
    if (a == 4 * 0xff || a == 0) {
      r = Sum2(r_ptr + j, rgb_stride);
      g = Sum2(g_ptr + j, rgb_stride);
      b = Sum2(b_ptr + j, rgb_stride);
    } else {
      r = LinearToGammaWeighted(r_ptr + j, a_ptr + j, a, 0, rgb_stride, mode);
      g = LinearToGammaWeighted(g_ptr + j, a_ptr + j, a, 0, rgb_stride, mode);
      b = LinearToGammaWeighted(b_ptr + j, a_ptr + j, a, 0, rgb_stride, mode);
    }
    dst[0] = (uint16_t)r; dst[1] = (uint16_t)g; dst[2] = (uint16_t)b;
    dst[3] = (uint16_t)a;
  }
}
/* yuv.go:286-302 */
static void AccumulateRGB(const uint8_t* r_ptr, const uint8_t* g_ptr,
                          const uint8_t* b_ptr, int step, int rgb_stride,
                          uint16_t* dst, int width) {
  int i, j;
  for (i = 0, j = 0; i < (width >> 1); i += 1, j += 2 * step, dst += 4) {
    dst[0] = (uint16_t)Sum4(r_ptr + j, step, rgb_stride);
    dst[1] = (uint16_t)Sum4(g_ptr + j, step, rgb_stride);
    dst[2] = (uint16_t)Sum4(b_ptr + j, step, rgb_stride);
  }
  if (width & 1) {
    dst[0] = (uint16_t)Sum2(r_ptr + j, rgb_stride);
    dst[1] = (uint16_t)Sum2(g_ptr + j, rgb_stride);
    dst[2] = (uint16_t)Sum2(b_ptr + j, rgb_stride);
  }
}
/* pkg/libwebp/dsp/alpha_processing.go:193-207 */
static int ExtractAlpha(const uint8_t* argb, int argb_stride, int width,
                        int height, uint8_t* alpha, int alpha_stride) {
  uint8_t alpha_mask = 0xff;
  int i, j;
  for (j = 0; j < height; ++j) {
    for (i = 0; i < width; ++i) {
      const uint8_t alpha_value = argb[4 * i];
      alpha[i] = alpha_value;
      alpha_mask &= alpha_value;
    }
    argb += argb_stride;
    alpha += alpha_stride;
  }
  return (alpha_mask == 0xff);
}
/* upstream CheckNonOpaque(a,w,h,step,stride), the intent of                   */
/* pkg/picture/conversions.go:56 / alpha/opague.go:16-24 (Appendix B-4).       */
static int CheckNonOpaque(const uint8_t* alpha, int width, int height,
                          int x_step, int y_step) {
  if (alpha == NULL) return 0;
  for (; height-- > 0; alpha += y_step) {
    int x;
    for (x = 0; x < width * x_step; x += x_step) {
      if (alpha[x] != 0xff) return 1;
    }
  }
  return 0;
}

/* yuv.go:304-355 */
static void ImportYUVAFromRGBA_rows(const uint8_t* r_ptr, const uint8_t* g_ptr,
                                    const uint8_t* b_ptr, const uint8_t* a_ptr,
                                    int step, int rgb_stride, int has_alpha,
                                    int width, int height, uint16_t* tmp_rgb,
                                    int y_stride, int uv_stride, int a_stride,
                                    uint8_t* dst_y, uint8_t* dst_u,
                                    uint8_t* dst_v, uint8_t* dst_a, int mode) {
  int y;
  const int uv_width = (width + 1) >> 1;
  has_alpha = has_alpha && (dst_a != NULL);
  for (y = 0; y < (height >> 1); ++y) {
    int rows_have_alpha = has_alpha;
    ConvertRGBToY(r_ptr, g_ptr, b_ptr, dst_y, width, step);
    ConvertRGBToY(r_ptr + rgb_stride, g_ptr + rgb_stride, b_ptr + rgb_stride,
                  dst_y + y_stride, width, step);
    dst_y += 2 * y_stride;
    if (has_alpha) {
      rows_have_alpha &=
          !ExtractAlpha(a_ptr, rgb_stride, width, 2, dst_a, a_stride);
      dst_a += 2 * a_stride;
    } else if (dst_a != NULL) {
      int i;
      for (i = 0; i < 2; ++i, dst_a += a_stride) memset(dst_a, 0xff, width);
    }
    if (!rows_have_alpha) {
      AccumulateRGB(r_ptr, g_ptr, b_ptr, step, rgb_stride, tmp_rgb, width);
    } else {
      AccumulateRGBA(r_ptr, g_ptr, b_ptr, a_ptr, rgb_stride, tmp_rgb, width,
                     mode);
    }
    ConvertRGBA32ToUV(tmp_rgb, dst_u, dst_v, uv_width);
    dst_u += uv_stride;
    dst_v += uv_stride;
    r_ptr += 2 * rgb_stride;
    b_ptr += 2 * rgb_stride;
    g_ptr += 2 * rgb_stride;
    if (has_alpha) a_ptr += 2 * rgb_stride;
  }
}
/* yuv.go:357-382 */
static void ImportYUVAFromRGBALastLine(const uint8_t* r_ptr,
                                       const uint8_t* g_ptr,
                                       const uint8_t* b_ptr,
                                       const uint8_t* a_ptr, int step,
                                       int has_alpha, int width,
                                       uint16_t* tmp_rgb, uint8_t* dst_y,
                                       uint8_t* dst_u, uint8_t* dst_v,
                                       uint8_t* dst_a, int mode) {
  const int uv_width = (width + 1) >> 1;
  int row_has_alpha = has_alpha && dst_a != NULL;
  ConvertRGBToY(r_ptr, g_ptr, b_ptr, dst_y, width, step);
  if (row_has_alpha) {
    row_has_alpha &= !ExtractAlpha(a_ptr, 0, width, 1, dst_a, 0);
  } else if (dst_a != NULL) {
    memset(dst_a, 0xff, width);
  }
  if (!row_has_alpha) {
    AccumulateRGB(r_ptr, g_ptr, b_ptr, step, /*rgb_stride=*/0, tmp_rgb, width);
  } else {
    AccumulateRGBA(r_ptr, g_ptr, b_ptr, a_ptr, /*rgb_stride=*/0, tmp_rgb, width,
                   mode);
  }
  ConvertRGBA32ToUV(tmp_rgb, dst_u, dst_v, uv_width);
}

/*
 * pkg/picture/conversions.go:50-121 (no-dither, non-iterative branch) with the
 * channel selection of pkg/libwebp/enc/picture_csp_enc.c.go:123-137.
 * r_off/g_off/b_off/a_off: byte offsets of the channels inside a pixel
 * (a_off < 0: no alpha imported).  dst_a may be NULL.  When the image turns
 * out opaque and dst_a is given it is filled with 0xff (yuv.go:329-336).
 * Returns has_alpha (conversions.go:56-58 -> WEBP_YUV420A vs WEBP_YUV420),
 * or -1 on bad arguments.
 */
int orc_import_yuva(const uint8_t* rgb, int r_off, int g_off, int b_off,
                    int a_off, int step, int rgb_stride, int width, int height,
                    uint8_t* dst_y, int y_stride, uint8_t* dst_u,
                    uint8_t* dst_v, int uv_stride, uint8_t* dst_a,
                    int a_stride, int divide_mode) {
  const uint8_t* r_ptr = rgb + r_off;
  const uint8_t* g_ptr = rgb + g_off;
  const uint8_t* b_ptr = rgb + b_off;
  const uint8_t* a_ptr = (a_off >= 0) ? rgb + a_off : NULL;
  int has_alpha;
  uint16_t* tmp_rgb;
  if (width <= 0 || height <= 0) return -1;
  if (abs(rgb_stride) < ((a_off >= 0) ? 4 : 3) * width) return -1;
  if (a_off >= 0 && step != 4) return -1;
  orc_init();
  has_alpha = CheckNonOpaque(a_ptr, width, height, step, rgb_stride);
  tmp_rgb = (uint16_t*)malloc(sizeof(uint16_t) * 4 * ((width + 1) >> 1));
  if (tmp_rgb == NULL) return -1;
  ImportYUVAFromRGBA_rows(r_ptr, g_ptr, b_ptr, a_ptr, step, rgb_stride,
                          has_alpha, width, height, tmp_rgb, y_stride,
                          uv_stride, a_stride, dst_y, dst_u, dst_v, dst_a,
                          divide_mode);
  if (height & 1) {
    const ptrdiff_t ro = (ptrdiff_t)(height - 1) * rgb_stride;
    ImportYUVAFromRGBALastLine(
        r_ptr + ro, g_ptr + ro, b_ptr + ro, has_alpha ? a_ptr + ro : a_ptr,
        step, has_alpha, width, tmp_rgb,
        dst_y + (ptrdiff_t)(height - 1) * y_stride,
        dst_u + (ptrdiff_t)(height >> 1) * uv_stride,
        dst_v + (ptrdiff_t)(height >> 1) * uv_stride,
        dst_a ? dst_a + (ptrdiff_t)(height - 1) * a_stride : NULL, divide_mode);
  }
  free(tmp_rgb);
  return has_alpha;
}

/* ---- dithered import ---------------------------------------- */
/* pkg/vp8/random.go:13-95 (VP8Random; D. Knuth's difference-based generator).  The Go
 * text increments the indices after testing them against the table size; the intended C
 * semantics (upstream random_utils.h: `if (++index == VP8_RANDOM_TABLE_SIZE) index = 0`)
 * are implemented, and pinned by the binary's WebPPictureARGBToYUVADithered. */
#define VP8_RANDOM_DITHER_FIX 8
#define VP8_RANDOM_TABLE_SIZE 55
static const uint32_t kRandomTable[VP8_RANDOM_TABLE_SIZE] = {
    0x0de15230, 0x03b31886, 0x775faccb, 0x1c88626a, 0x68385c55, 0x14b3b828, 0x4a85fef8, 0x49ddb84b,
    0x64fcf397, 0x5c550289, 0x4a290000, 0x0d7ec1da, 0x5940b7ab, 0x5492577d, 0x4e19ca72, 0x38d38c69,
    0x0c01ee65, 0x32a1755f, 0x5437f652, 0x5abb2c32, 0x0faa57b1, 0x73f533e7, 0x685feeda, 0x7563cce2,
    0x6e990e83, 0x4730a7ed, 0x4fc0d9c6, 0x496b153c, 0x4f1403fa, 0x541afb0c, 0x73990b32, 0x26d7cb1c,
    0x6fcc3706, 0x2cbb77d8, 0x75762f2a, 0x6425ccdd, 0x24b35461, 0x0a7d8715, 0x220414a8, 0x141ebf67,
    0x56b41583, 0x73e502e3, 0x44cab16f, 0x28264d42, 0x73baaefb, 0x0a50ebed, 0x1d6ab6fb, 0x0d3ad40b,
    0x35db3b68, 0x2b081e83, 0x77ce6b95, 0x5181e5f0, 0x78853bbc, 0x009f9494, 0x27e5ed3c};
typedef struct {
  int index1, index2;
  uint32_t tab[VP8_RANDOM_TABLE_SIZE];
  int amp;
} VP8Random;
/* random.go:49-63 */
static void VP8InitRandom(VP8Random* rg, float dithering) {
  memcpy(rg->tab, kRandomTable, sizeof(rg->tab));
  rg->index1 = 0;
  rg->index2 = 31;
  rg->amp = (dithering < 0.f) ? 0
            : (dithering > 1.f) ? (1 << VP8_RANDOM_DITHER_FIX)
                                : (int)((1 << VP8_RANDOM_DITHER_FIX) * dithering);
}
/* random.go:68-95 */
static int VP8RandomBits(VP8Random* rg, int num_bits) {
  int diff = (int)rg->tab[rg->index1] - (int)rg->tab[rg->index2];
  if (diff < 0) diff += (1u << 31);
  rg->tab[rg->index1] = (uint32_t)diff;
  if (++rg->index1 == VP8_RANDOM_TABLE_SIZE) rg->index1 = 0;
  if (++rg->index2 == VP8_RANDOM_TABLE_SIZE) rg->index2 = 0;
  diff = (int)((uint32_t)diff << 1) >> (32 - num_bits); /* sign-extend, 0-center */
  diff = (diff * rg->amp) >> VP8_RANDOM_DITHER_FIX;     /* restrict range */
  diff += 1 << (num_bits - 1);                          /* shift back to 0.5-center */
  return diff;
}
/* the amp and the first n raw values of the generator, for tests of the skip-ahead */
int orc_vp8_random(float dithering, int num_bits, int n, int32_t* out) {
  VP8Random rg;
  VP8InitRandom(&rg, dithering);
  for (int i = 0; i < n; ++i) out[i] = VP8RandomBits(&rg, num_bits);
  return rg.amp;
}
/* pkg/libwebp/enc/picture_csp_enc.c.go:31-39 */
static void ConvertRowToY(const uint8_t* r_ptr, const uint8_t* g_ptr, const uint8_t* b_ptr, int step, uint8_t* dst_y,
                          int width, VP8Random* rg) {
  for (int i = 0, j = 0; i < width; i += 1, j += step) {
    dst_y[i] = (uint8_t)orc_RGBToY(r_ptr[j], g_ptr[j], b_ptr[j], VP8RandomBits(rg, YUV_FIX));
  }
}
/* picture_csp_enc.c.go:41-48 */
static void ConvertRowsToUV(const uint16_t* rgb, uint8_t* dst_u, uint8_t* dst_v, int width, VP8Random* rg) {
  for (int i = 0; i < width; i += 1, rgb += 4) {
    const int r = rgb[0], g = rgb[1], b = rgb[2];
    dst_u[i] = (uint8_t)orc_RGBToU(r, g, b, VP8RandomBits(rg, YUV_FIX + 2));
    dst_v[i] = (uint8_t)orc_RGBToV(r, g, b, VP8RandomBits(rg, YUV_FIX + 2));
  }
}
/*
 * pkg/picture/conversions.go:122-165: the dithered copy of the import (dithering > 0,
 * no iterative conversion).  Same arguments as orc_import_yuva; dst_a of an opaque image
 * is filled with 0xff as there.  One generator runs through the whole image: two luma
 * rows, then the chroma row of every row pair (u then v per sample).
 */
int orc_import_yuva_dithered(const uint8_t* rgb, int r_off, int g_off, int b_off, int a_off, int step, int rgb_stride,
                             int width, int height, uint8_t* dst_y, int y_stride, uint8_t* dst_u, uint8_t* dst_v,
                             int uv_stride, uint8_t* dst_a, int a_stride, int divide_mode, float dithering) {
  const uint8_t* r_ptr = rgb + r_off;
  const uint8_t* g_ptr = rgb + g_off;
  const uint8_t* b_ptr = rgb + b_off;
  const uint8_t* a_ptr = (a_off >= 0) ? rgb + a_off : NULL;
  const int uv_width = (width + 1) >> 1;
  int has_alpha, use_alpha, y;
  uint16_t* tmp_rgb;
  VP8Random rg;
  if (width <= 0 || height <= 0) return -1;
  if (abs(rgb_stride) < ((a_off >= 0) ? 4 : 3) * width) return -1;
  if (a_off >= 0 && step != 4) return -1;
  orc_init();
  has_alpha = CheckNonOpaque(a_ptr, width, height, step, rgb_stride);
  use_alpha = has_alpha && dst_a != NULL;
  tmp_rgb = (uint16_t*)malloc(sizeof(uint16_t) * 4 * uv_width);
  if (tmp_rgb == NULL) return -1;
  VP8InitRandom(&rg, dithering);
  for (y = 0; y < (height >> 1); ++y) {
    int rows_have_alpha = use_alpha;
    ConvertRowToY(r_ptr, g_ptr, b_ptr, step, dst_y, width, &rg);
    ConvertRowToY(r_ptr + rgb_stride, g_ptr + rgb_stride, b_ptr + rgb_stride, step, dst_y + y_stride, width, &rg);
    dst_y += 2 * y_stride;
    if (use_alpha) {
      rows_have_alpha &= !ExtractAlpha(a_ptr, rgb_stride, width, 2, dst_a, a_stride);
      dst_a += 2 * a_stride;
    } else if (dst_a != NULL) {
      for (int i = 0; i < 2; ++i, dst_a += a_stride) memset(dst_a, 0xff, width);
    }
    if (!rows_have_alpha) {
      AccumulateRGB(r_ptr, g_ptr, b_ptr, step, rgb_stride, tmp_rgb, width);
    } else {
      AccumulateRGBA(r_ptr, g_ptr, b_ptr, a_ptr, rgb_stride, tmp_rgb, width, divide_mode);
    }
    ConvertRowsToUV(tmp_rgb, dst_u, dst_v, uv_width, &rg);
    dst_u += uv_stride;
    dst_v += uv_stride;
    r_ptr += 2 * rgb_stride;
    b_ptr += 2 * rgb_stride;
    g_ptr += 2 * rgb_stride;
    if (use_alpha) a_ptr += 2 * rgb_stride;
  }
  if (height & 1) { /* extra last row */
    int row_has_alpha = use_alpha;
    ConvertRowToY(r_ptr, g_ptr, b_ptr, step, dst_y, width, &rg);
    if (row_has_alpha) {
      row_has_alpha &= !ExtractAlpha(a_ptr, 0, width, 1, dst_a, 0);
    } else if (dst_a != NULL) {
      memset(dst_a, 0xff, width);
    }
    if (!row_has_alpha) {
      AccumulateRGB(r_ptr, g_ptr, b_ptr, step, /*rgb_stride=*/0, tmp_rgb, width);
    } else {
      AccumulateRGBA(r_ptr, g_ptr, b_ptr, a_ptr, /*rgb_stride=*/0, tmp_rgb, width, divide_mode);
    }
    ConvertRowsToUV(tmp_rgb, dst_u, dst_v, uv_width, &rg);
  }
  free(tmp_rgb);
  return has_alpha;
}

/* yuv.go:74-120 (non-gamma ARGB -> Y / UV rows; row a14). */
void orc_ConvertARGBToY(const uint32_t* argb, uint8_t* y, int width) {
  int i;
  for (i = 0; i < width; ++i) {
    const uint32_t p = argb[i];
    y[i] = (uint8_t)orc_RGBToY((p >> 16) & 0xff, (p >> 8) & 0xff, p & 0xff,
                               YUV_HALF);
  }
}
void orc_ConvertARGBToUV(const uint32_t* argb, uint8_t* u, uint8_t* v,
                         int src_width, int do_store) {
  const int uv_width = src_width >> 1;
  int i;
  for (i = 0; i < uv_width; ++i) {
    const uint32_t v0 = argb[2 * i + 0];
    const uint32_t v1 = argb[2 * i + 1];
    const int r = ((v0 >> 15) & 0x1fe) + ((v1 >> 15) & 0x1fe);
    const int g = ((v0 >> 7) & 0x1fe) + ((v1 >> 7) & 0x1fe);
    const int b = ((v0 << 1) & 0x1fe) + ((v1 << 1) & 0x1fe);
    const int tmp_u = orc_RGBToU(r, g, b, YUV_HALF << 2);
    const int tmp_v = orc_RGBToV(r, g, b, YUV_HALF << 2);
    if (do_store) {
      u[i] = (uint8_t)tmp_u; v[i] = (uint8_t)tmp_v;
    } else {
      u[i] = (uint8_t)((u[i] + tmp_u + 1) >> 1);
      v[i] = (uint8_t)((v[i] + tmp_v + 1) >> 1);
    }
  }
  if (src_width & 1) {
    const uint32_t v0 = argb[2 * i + 0];
    const int r = (v0 >> 14) & 0x3fc;
    const int g = (v0 >> 6) & 0x3fc;
    const int b = (v0 << 2) & 0x3fc;
    const int tmp_u = orc_RGBToU(r, g, b, YUV_HALF << 2);
    const int tmp_v = orc_RGBToV(r, g, b, YUV_HALF << 2);
    if (do_store) {
      u[i] = (uint8_t)tmp_u; v[i] = (uint8_t)tmp_v;
    } else {
      u[i] = (uint8_t)((u[i] + tmp_u + 1) >> 1);
      v[i] = (uint8_t)((v[i] + tmp_v + 1) >> 1);
    }
  }
}

/* ------------------------------------------------------------------------- */
/* Part 2: fancy upsampler.                                                  */
/* Output modes: pkg/libwebp/webp/enums.go:45-64.                            */
enum {
  MODE_RGB = 0, MODE_RGBA = 1, MODE_BGR = 2, MODE_BGRA = 3, MODE_ARGB = 4,
  MODE_RGBA_4444 = 5, MODE_RGB_565 = 6, MODE_rgbA = 7, MODE_bgrA = 8,
  MODE_Argb = 9, MODE_rgbA_4444 = 10
};
/* kModeBpp: pkg/libwebp/decoder/buffer_dec.c.go:28-32 */
int orc_mode_bpp(int mode) {
  static const int bpp[11] = {3, 4, 3, 4, 4, 2, 2, 4, 4, 4, 2};
  return (mode >= 0 && mode < 11) ? bpp[mode] : 0;
}
/* Per-pixel writers (referenced by upsampling.c.go:107-114, undefined in the  */
/* Go tree; upstream yuv.h semantics, pinned by the libwebp binary via         */
/* WebPDecodeRGBA/BGRA/ARGB -- SURVEY Appendix D).                             */
static inline void YuvToPixel(int mode, int y, int u, int v, uint8_t* dst) {
  const int r = orc_YUVToR(y, v), g = orc_YUVToG(y, u, v), b = orc_YUVToB(y, u);
  switch (mode) {
    case MODE_RGB: dst[0] = r; dst[1] = g; dst[2] = b; break;
    case MODE_BGR: dst[0] = b; dst[1] = g; dst[2] = r; break;
    case MODE_RGBA: case MODE_rgbA:
      dst[0] = r; dst[1] = g; dst[2] = b; dst[3] = 0xff; break;
    case MODE_BGRA: case MODE_bgrA:
      dst[0] = b; dst[1] = g; dst[2] = r; dst[3] = 0xff; break;
    case MODE_ARGB: case MODE_Argb:
      dst[0] = 0xff; dst[1] = r; dst[2] = g; dst[3] = b; break;
    case MODE_RGBA_4444: case MODE_rgbA_4444: {
      /* byte order of pkg/libwebp/dsp/lossless.go:575-585 (Appendix B-7 is  */
      /* about that file); upstream default yuv.h order: rg then ba.         */
      const int rg = (r & 0xf0) | (g >> 4);
      const int ba = (b & 0xf0) | 0x0f;
      dst[0] = rg; dst[1] = ba;
      break;
    }
    case MODE_RGB_565: {
      const int rg = (r & 0xf8) | (g >> 5);
      const int gb = ((g << 3) & 0xe0) | (b >> 3);
      dst[0] = rg; dst[1] = gb;
      break;
    }
    default: break;
  }
}

/* pkg/libwebp/dsp/upsampling.c.go:41-103 (UPSAMPLE_FUNC), SWAR as written. */
#define LOAD_UV(u, v) ((uint32_t)(u) | ((uint32_t)(v) << 16))
void orc_upsample_line_pair(int mode, const uint8_t* top_y,
                            const uint8_t* bottom_y, const uint8_t* top_u,
                            const uint8_t* top_v, const uint8_t* cur_u,
                            const uint8_t* cur_v, uint8_t* top_dst,
                            uint8_t* bottom_dst, int len) {
  const int XSTEP = orc_mode_bpp(mode);
  int x;
  const int last_pixel_pair = (len - 1) >> 1;
  uint32_t tl_uv = LOAD_UV(top_u[0], top_v[0]);
  uint32_t l_uv = LOAD_UV(cur_u[0], cur_v[0]);
  {
    const uint32_t uv0 = (3 * tl_uv + l_uv + 0x00020002u) >> 2;
    YuvToPixel(mode, top_y[0], uv0 & 0xff, (uv0 >> 16) & 0xff, top_dst);
  }
  if (bottom_y != NULL) {
    const uint32_t uv0 = (3 * l_uv + tl_uv + 0x00020002u) >> 2;
    YuvToPixel(mode, bottom_y[0], uv0 & 0xff, (uv0 >> 16) & 0xff, bottom_dst);
  }
  for (x = 1; x <= last_pixel_pair; ++x) {
    const uint32_t t_uv = LOAD_UV(top_u[x], top_v[x]);
    const uint32_t uv = LOAD_UV(cur_u[x], cur_v[x]);
    const uint32_t avg = tl_uv + t_uv + l_uv + uv + 0x00080008u;
    const uint32_t diag_12 = (avg + 2 * (t_uv + l_uv)) >> 3;
    const uint32_t diag_03 = (avg + 2 * (tl_uv + uv)) >> 3;
    {
      const uint32_t uv0 = (diag_12 + tl_uv) >> 1;
      const uint32_t uv1 = (diag_03 + t_uv) >> 1;
      YuvToPixel(mode, top_y[2 * x - 1], uv0 & 0xff, (uv0 >> 16) & 0xff,
                 top_dst + (2 * x - 1) * XSTEP);
      YuvToPixel(mode, top_y[2 * x - 0], uv1 & 0xff, (uv1 >> 16) & 0xff,
                 top_dst + (2 * x - 0) * XSTEP);
    }
    if (bottom_y != NULL) {
      const uint32_t uv0 = (diag_03 + l_uv) >> 1;
      const uint32_t uv1 = (diag_12 + uv) >> 1;
      YuvToPixel(mode, bottom_y[2 * x - 1], uv0 & 0xff, (uv0 >> 16) & 0xff,
                 bottom_dst + (2 * x - 1) * XSTEP);
      YuvToPixel(mode, bottom_y[2 * x + 0], uv1 & 0xff, (uv1 >> 16) & 0xff,
                 bottom_dst + (2 * x + 0) * XSTEP);
    }
    tl_uv = t_uv;
    l_uv = uv;
  }
  if (!(len & 1)) {
    {
      const uint32_t uv0 = (3 * tl_uv + l_uv + 0x00020002u) >> 2;
      YuvToPixel(mode, top_y[len - 1], uv0 & 0xff, (uv0 >> 16) & 0xff,
                 top_dst + (len - 1) * XSTEP);
    }
    if (bottom_y != NULL) {
      const uint32_t uv0 = (3 * l_uv + tl_uv + 0x00020002u) >> 2;
      YuvToPixel(mode, bottom_y[len - 1], uv0 & 0xff, (uv0 >> 16) & 0xff,
                 bottom_dst + (len - 1) * XSTEP);
    }
  }
}

/* pkg/libwebp/dsp/alpha_processing.go:165-180 (DispatchAlpha_C). */
static int DispatchAlpha(const uint8_t* alpha, int alpha_stride, int width,
                         int height, uint8_t* dst, int dst_stride) {
  uint32_t alpha_mask = 0xff;
  int i, j;
  for (j = 0; j < height; ++j) {
    for (i = 0; i < width; ++i) {
      const uint32_t alpha_value = alpha[i];
      dst[4 * i] = (uint8_t)alpha_value;
      alpha_mask &= alpha_value;
    }
    alpha += alpha_stride;
    dst += dst_stride;
  }
  return (alpha_mask != 0xff);
}

/*
 * Whole-image schedule of pkg/libwebp/enc/picture_csp_enc.c.go:62-118
 * (WebPPictureYUVAToARGB), generalised over the output mode of
 * WebPUpsamplers[] (upsampling.c.go:275-289); alpha insertion per
 * picture_csp_enc.c.go:106-115 / io_dec.c.go:175-195 (offset 0 for ARGB,
 * else 3; 4-byte modes only).  a may be NULL.  Returns 1, or 0 on bad args.
 */
int orc_yuva_to_rgba(int mode, const uint8_t* y, int y_stride,
                     const uint8_t* u, const uint8_t* v, int uv_stride,
                     const uint8_t* a, int a_stride, int width, int height,
                     uint8_t* dst, int dst_stride) {
  int row;
  const uint8_t *cur_u = u, *cur_v = v, *cur_y = y;
  uint8_t* d = dst;
  if (width <= 0 || height <= 0 || orc_mode_bpp(mode) == 0) return 0;
  orc_upsample_line_pair(mode, cur_y, NULL, cur_u, cur_v, cur_u, cur_v, d, NULL,
                         width);
  cur_y += y_stride;
  d += dst_stride;
  for (row = 1; row + 1 < height; row += 2) {
    const uint8_t* const top_u = cur_u;
    const uint8_t* const top_v = cur_v;
    cur_u += uv_stride;
    cur_v += uv_stride;
    orc_upsample_line_pair(mode, cur_y, cur_y + y_stride, top_u, top_v, cur_u,
                           cur_v, d, d + dst_stride, width);
    cur_y += 2 * y_stride;
    d += 2 * dst_stride;
  }
  if (height > 1 && !(height & 1)) {
    orc_upsample_line_pair(mode, cur_y, NULL, cur_u, cur_v, cur_u, cur_v, d,
                           NULL, width);
  }
  if (a != NULL && orc_mode_bpp(mode) == 4) {
    const int alpha_first = (mode == MODE_ARGB || mode == MODE_Argb);
    DispatchAlpha(a, a_stride, width, height, dst + (alpha_first ? 0 : 3),
                  dst_stride);
  } else if (a != NULL && (mode == MODE_RGBA_4444 || mode == MODE_rgbA_4444)) {
    /* EmitAlphaRGBA4444, pkg/libwebp/decoder/io_dec.c.go:197-226 (without the
     * premultiply, like the 4-byte modes here): the alpha nibble of the pixel takes
     * alpha >> 4.  The 4444 writers of this path keep the upstream default byte order
     * (rg byte first, ba byte second: what the libwebp binary pins), so the nibble is in
     * byte 1; the Go text shows the WEBP_SWAP_16BIT_CSP build, where it is byte 0. */
    int i, j;
    for (j = 0; j < height; ++j) {
      uint8_t* const alpha_dst = dst + (ptrdiff_t)j * dst_stride + 1;
      for (i = 0; i < width; ++i) {
        const uint32_t alpha_value = a[(ptrdiff_t)j * a_stride + i] >> 4;
        alpha_dst[2 * i] = (uint8_t)((alpha_dst[2 * i] & 0xf0) | alpha_value);
      }
    }
  }
  return 1;
}

/*
 * Streaming schedule of pkg/libwebp/decoder/io_dec.c.go:65-115 (EmitFancyRGB):
 * the image is delivered in batches of rows [mb_y, mb_y+mb_h) (mb_y even
 * except for nothing; batches as the lossy decoder emits them) and one
 * unfinished row is carried in tmp_y/u/v.  Used by the tests to show the
 * batched whole-image kernel equals the decoder's streaming order.
 * 'splits' holds n_splits ascending batch end rows, the last == height.
 */
int orc_emit_fancy_rgb_stream(int mode, const uint8_t* y, int y_stride,
                              const uint8_t* u, const uint8_t* v,
                              int uv_stride, int width, int height,
                              const int* splits, int n_splits, uint8_t* rgba,
                              int stride) {
  const int uv_w = (width + 1) / 2;
  uint8_t* tmp_y = (uint8_t*)malloc((size_t)width);
  uint8_t* tmp_u = (uint8_t*)malloc((size_t)uv_w);
  uint8_t* tmp_v = (uint8_t*)malloc((size_t)uv_w);
  int mb_y = 0, s, total = 0;
  if (!tmp_y || !tmp_u || !tmp_v) return -1;
  for (s = 0; s < n_splits; ++s) {
    const int y_end = splits[s];
    const int mb_h = y_end - mb_y;
    int num_lines_out = mb_h;
    uint8_t* dst = rgba + (ptrdiff_t)mb_y * stride;
    const uint8_t* cur_y = y + (ptrdiff_t)mb_y * y_stride;
    const uint8_t* cur_u = u + (ptrdiff_t)(mb_y >> 1) * uv_stride;
    const uint8_t* cur_v = v + (ptrdiff_t)(mb_y >> 1) * uv_stride;
    const uint8_t* top_u = tmp_u;
    const uint8_t* top_v = tmp_v;
    int yy = mb_y;
    if (yy == 0) {
      orc_upsample_line_pair(mode, cur_y, NULL, cur_u, cur_v, cur_u, cur_v, dst,
                             NULL, width);
    } else {
      orc_upsample_line_pair(mode, tmp_y, cur_y, top_u, top_v, cur_u, cur_v,
                             dst - stride, dst, width);
      ++num_lines_out;
    }
    for (; yy + 2 < y_end; yy += 2) {
      top_u = cur_u;
      top_v = cur_v;
      cur_u += uv_stride;
      cur_v += uv_stride;
      dst += 2 * stride;
      cur_y += 2 * y_stride;
      orc_upsample_line_pair(mode, cur_y - y_stride, cur_y, top_u, top_v, cur_u,
                             cur_v, dst - stride, dst, width);
    }
    cur_y += y_stride;
    if (y_end < height) {
      memcpy(tmp_y, cur_y, (size_t)width);
      memcpy(tmp_u, cur_u, (size_t)uv_w);
      memcpy(tmp_v, cur_v, (size_t)uv_w);
      --num_lines_out;
    } else if (!(y_end & 1)) {
      orc_upsample_line_pair(mode, cur_y, NULL, cur_u, cur_v, cur_u, cur_v,
                             dst + stride, NULL, width);
    }
    total += num_lines_out;
    mb_y = y_end;
  }
  free(tmp_y); free(tmp_u); free(tmp_v);
  return total;
}

/* ------------------------------------------------------------------------- */
/* Part 3: VP8L per-pixel transforms.                                        */

/* pkg/libwebp/dsp/lossless.go:337-347 */
void orc_add_green(const uint32_t* src, int num_pixels, uint32_t* dst) {
  int i;
  for (i = 0; i < num_pixels; ++i) {
    const uint32_t argb = src[i];
    const uint32_t green = (argb >> 8) & 0xff;
    uint32_t red_blue = argb & 0x00ff00ffu;
    red_blue += (green << 16) | green;
    red_blue &= 0x00ff00ffu;
    dst[i] = (argb & 0xff00ff00u) | red_blue;
  }
}
/* pkg/libwebp/dsp/lossless_enc.go:207-216 */
void orc_sub_green(uint32_t* argb_data, int num_pixels) {
  int i;
  for (i = 0; i < num_pixels; ++i) {
    const int argb = (int)argb_data[i];
    const int green = (argb >> 8) & 0xff;
    const uint32_t new_r = (((argb >> 16) & 0xff) - green) & 0xff;
    const uint32_t new_b = (((argb >> 0) & 0xff) - green) & 0xff;
    argb_data[i] = ((uint32_t)argb & 0xff00ff00u) | (new_r << 16) | new_b;
  }
}
/* lossless.go:349-351 with both operands promoted to int (Appendix B-3). */
int orc_ColorTransformDelta(int8_t color_pred, int8_t color) {
  return ((int)color_pred * (int)color) >> 5;
}
typedef struct { uint8_t green_to_red, green_to_blue, red_to_blue; } Multipliers;
/* lossless.go:353-357 */
static inline void ColorCodeToMultipliers(uint32_t color_code, Multipliers* m) {
  m->green_to_red = (color_code >> 0) & 0xff;
  m->green_to_blue = (color_code >> 8) & 0xff;
  m->red_to_blue = (color_code >> 16) & 0xff;
}
/* lossless.go:359-374 */
static void TransformColorInverse(const Multipliers* m, const uint32_t* src,
                                  int num_pixels, uint32_t* dst) {
  int i;
  for (i = 0; i < num_pixels; ++i) {
    const uint32_t argb = src[i];
    const int8_t green = (int8_t)(argb >> 8);
    const uint32_t red = argb >> 16;
    int new_red = red & 0xff;
    int new_blue = argb & 0xff;
    new_red += orc_ColorTransformDelta((int8_t)m->green_to_red, green);
    new_red &= 0xff;
    new_blue += orc_ColorTransformDelta((int8_t)m->green_to_blue, green);
    new_blue += orc_ColorTransformDelta((int8_t)m->red_to_blue, (int8_t)new_red);
    new_blue &= 0xff;
    dst[i] = (argb & 0xff00ff00u) | ((uint32_t)new_red << 16) | (uint32_t)new_blue;
  }
}
void orc_transform_color_inverse(uint32_t color_code, const uint32_t* src,
                                 int num_pixels, uint32_t* dst) {
  Multipliers m;
  ColorCodeToMultipliers(color_code, &m);
  TransformColorInverse(&m, src, num_pixels, dst);
}
/* pkg/libwebp/dsp/lossless_common.go:34-36 */
static inline uint32_t SubSampleSize(uint32_t size, uint32_t sampling_bits) {
  return (size + (1u << sampling_bits) - 1) >> sampling_bits;
}
uint32_t orc_VP8LSubSampleSize(uint32_t size, uint32_t bits) {
  return SubSampleSize(size, bits);
}
/* lossless.go:377-413; src/dst point at row y_start. */
void orc_color_space_inverse(int bits, int xsize, const uint32_t* data,
                             int y_start, int y_end, const uint32_t* src,
                             uint32_t* dst) {
  const int width = xsize;
  const int tile_width = 1 << bits;
  const int mask = tile_width - 1;
  const int safe_width = width & ~mask;
  const int remaining_width = width - safe_width;
  const int tiles_per_row = (int)SubSampleSize(width, bits);
  int y = y_start;
  const uint32_t* pred_row = data + (y >> bits) * tiles_per_row;
  while (y < y_end) {
    const uint32_t* pred = pred_row;
    Multipliers m = {0, 0, 0};
    const uint32_t* const src_safe_end = src + safe_width;
    const uint32_t* const src_end = src + width;
    while (src < src_safe_end) {
      ColorCodeToMultipliers(*pred++, &m);
      TransformColorInverse(&m, src, tile_width, dst);
      src += tile_width;
      dst += tile_width;
    }
    if (src < src_end) {
      ColorCodeToMultipliers(*pred++, &m);
      TransformColorInverse(&m, src, remaining_width, dst);
      src += remaining_width;
      dst += remaining_width;
    }
    ++y;
    if ((y & mask) == 0) pred_row += tiles_per_row;
  }
}
/* pkg/libwebp/dsp/lossless_enc.go:222-239 */
static void TransformColor(const Multipliers* m, uint32_t* data,
                           int num_pixels) {
  int i;
  for (i = 0; i < num_pixels; ++i) {
    const uint32_t argb = data[i];
    const int8_t green = (int8_t)((argb >> 8) & 0xff);
    const int8_t red = (int8_t)((argb >> 16) & 0xff);
    int new_red = red & 0xff;
    int new_blue = argb & 0xff;
    new_red -= orc_ColorTransformDelta((int8_t)m->green_to_red, green);
    new_red &= 0xff;
    new_blue -= orc_ColorTransformDelta((int8_t)m->green_to_blue, green);
    new_blue -= orc_ColorTransformDelta((int8_t)m->red_to_blue, red);
    new_blue &= 0xff;
    data[i] = (argb & 0xff00ff00u) | ((uint32_t)new_red << 16) | (uint32_t)new_blue;
  }
}
void orc_transform_color(uint32_t color_code, uint32_t* data, int num_pixels) {
  Multipliers m;
  ColorCodeToMultipliers(color_code, &m);
  TransformColor(&m, data, num_pixels);
}
/* Forward cross-colour over an image with GIVEN tile codes: the tile walk of  */
/* pkg/libwebp/enc/predictor_enc.go:879-889 (CopyTileWithColorTransform) for   */
/* every tile of :913-925; the multiplier search (:741-877) is out of scope.   */
void orc_color_space_forward(int bits, int width, int height,
                             const uint32_t* data, uint32_t* argb) {
  const int max_tile_size = 1 << bits;
  const int tile_xsize = (int)SubSampleSize(width, bits);
  const int tile_ysize = (int)SubSampleSize(height, bits);
  int tile_x, tile_y;
  for (tile_y = 0; tile_y < tile_ysize; ++tile_y) {
    for (tile_x = 0; tile_x < tile_xsize; ++tile_x) {
      Multipliers m;
      const int xscan0 = max_tile_size < width - tile_x * max_tile_size
                             ? max_tile_size : width - tile_x * max_tile_size;
      int yscan = max_tile_size < height - tile_y * max_tile_size
                      ? max_tile_size : height - tile_y * max_tile_size;
      uint32_t* p = argb + (size_t)tile_y * max_tile_size * width +
                    (size_t)tile_x * max_tile_size;
      ColorCodeToMultipliers(data[tile_y * tile_xsize + tile_x], &m);
      while (yscan-- > 0) {
        TransformColor(&m, p, xscan0);
        p += width;
      }
    }
  }
}
/* lossless.go:415-426 (MapARGB_C), lossless_common.go:25-31 */
void orc_map_argb(const uint32_t* src, const uint32_t* color_map, uint32_t* dst,
                  int y_start, int y_end, int width) {
  int y;
  for (y = y_start; y < y_end; ++y) {
    int x;
    for (x = 0; x < width; ++x) *dst++ = color_map[(*src++ >> 8) & 0xff];
  }
}
/* lossless.go:428-459 */
void orc_color_index_inverse(int bits, int xsize, const uint32_t* color_map,
                             int y_start, int y_end, const uint32_t* src,
                             uint32_t* dst) {
  int y;
  const int bits_per_pixel = 8 >> bits;
  const int width = xsize;
  if (bits_per_pixel < 8) {
    const int pixels_per_byte = 1 << bits;
    const int count_mask = pixels_per_byte - 1;
    const uint32_t bit_mask = (1u << bits_per_pixel) - 1;
    for (y = y_start; y < y_end; ++y) {
      uint32_t packed_pixels = 0;
      int x;
      for (x = 0; x < width; ++x) {
        if ((x & count_mask) == 0) packed_pixels = (*src++ >> 8) & 0xff;
        *dst++ = color_map[packed_pixels & bit_mask];
        packed_pixels >>= bits_per_pixel;
      }
    }
  } else {
    orc_map_argb(src, color_map, dst, y_start, y_end, width);
  }
}
/* lossless.go:461-505 (alpha-plane variant: u8 source, green of the entry). */
void orc_color_index_inverse_alpha(int bits, int xsize,
                                   const uint32_t* color_map, int y_start,
                                   int y_end, const uint8_t* src,
                                   uint8_t* dst) {
  int y;
  const int bits_per_pixel = 8 >> bits;
  const int width = xsize;
  if (bits_per_pixel < 8) {
    const int pixels_per_byte = 1 << bits;
    const int count_mask = pixels_per_byte - 1;
    const uint32_t bit_mask = (1u << bits_per_pixel) - 1;
    for (y = y_start; y < y_end; ++y) {
      uint32_t packed_pixels = 0;
      int x;
      for (x = 0; x < width; ++x) {
        if ((x & count_mask) == 0) packed_pixels = *src++;
        *dst++ = (uint8_t)((color_map[packed_pixels & bit_mask] >> 8) & 0xff);
        packed_pixels >>= bits_per_pixel;
      }
    }
  } else {
    for (y = y_start; y < y_end; ++y) {
      int x;
      for (x = 0; x < width; ++x) {
        *dst++ = (uint8_t)((color_map[*src++] >> 8) & 0xff);
      }
    }
  }
}
/* pkg/vp8/vp8l_dec.c.go:1196-1219 (ExpandColorMap); out has 1<<(8>>bits). */
void orc_expand_color_map(int num_colors, int bits, const uint32_t* data,
                          uint32_t* new_color_map) {
  int i;
  const int final_num_colors = 1 << (8 >> bits);
  const uint8_t* d = (const uint8_t*)data;
  uint8_t* nd = (uint8_t*)new_color_map;
  new_color_map[0] = data[0];
  for (i = 4; i < 4 * num_colors; ++i) nd[i] = (d[i] + nd[i - 4]) & 0xff;
  for (; i < 4 * final_num_colors; ++i) nd[i] = 0;
}
/* lossless_enc.go:293-312 */
void orc_bundle_color_map(const uint8_t* row, int width, int xbits,
                          uint32_t* dst) {
  int x;
  if (xbits > 0) {
    const int bit_depth = 1 << (3 - xbits);
    const int mask = (1 << xbits) - 1;
    uint32_t code = 0xff000000u;
    for (x = 0; x < width; ++x) {
      const int xsub = x & mask;
      if (xsub == 0) code = 0xff000000u;
      code |= (uint32_t)row[x] << (8 + bit_depth * xsub);
      dst[x >> xbits] = code;
    }
  } else {
    for (x = 0; x < width; ++x) dst[x] = 0xff000000u | ((uint32_t)row[x] << 8);
  }
}

/* ---- ApplyPalette: pkg/vp8/vp8l_encoder.go:1093-1212, the colour -> index half of
 * row a33 (forward counterpart of colour indexing).  Strategy order as the reference:
 * greedy compare below 4 colours; else the first of three hashes that maps the palette
 * without collision into a 2048-entry table; else binary search in the sorted palette
 * (pkg/libwebp/utils/palette.go:74-119).  The hashes are the intended C ones: the product
 * is truncated to 32 bits before the shift (the Go text drops the cast).
 * The reference "assumes that all src[] values have a corresponding entry in the palette";
 * for the others it is only partly defined: greedy answers 3, a table answers what the
 * slot holds (0xffff truncated to a byte when empty), the binary search never ends -- here
 * it stops after 9 halvings and answers the entry it stands on. */
#define ORC_PALETTE_INV_SIZE_BITS 11
#define ORC_PALETTE_INV_SIZE (1 << ORC_PALETTE_INV_SIZE_BITS)
static uint32_t ApplyPaletteHash0(uint32_t color) { return (color >> 8) & 0xff; }           /* :1111-1114 */
static uint32_t ApplyPaletteHash1(uint32_t color) {                                         /* :1119-1122 */
  return ((uint32_t)((color & 0x00ffffffu) * 4222244071ull)) >> (32 - ORC_PALETTE_INV_SIZE_BITS);
}
static uint32_t ApplyPaletteHash2(uint32_t color) {                                         /* :1124-1127 */
  return ((uint32_t)((color & 0x00ffffffu) * ((1ull << 31) - 1))) >> (32 - ORC_PALETTE_INV_SIZE_BITS);
}
static uint32_t palette_hash(int which, uint32_t c) {
  return which == 0 ? ApplyPaletteHash0(c) : which == 1 ? ApplyPaletteHash1(c) : ApplyPaletteHash2(c);
}
static int SearchColorGreedy(const uint32_t* palette, uint32_t color) {                     /* :1095-1109 */
  if (color == palette[0]) return 0;
  if (color == palette[1]) return 1;
  if (color == palette[2]) return 2;
  return 3;
}
static int SearchColorNoIdxBounded(const uint32_t* sorted, uint32_t color, int num_colors) { /* palette.go:74-96 */
  int low = 0, hi = num_colors, it;
  if (sorted[low] == color) return low;
  for (it = 0; it < 9; ++it) {
    const int mid = (low + hi) >> 1;
    if (sorted[mid] == color) return mid;
    if (sorted[mid] < color) low = mid; else hi = mid;
  }
  return low;
}
static int cmp_u32(const void* a, const void* b) {
  const uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
  return x < y ? -1 : x > y;
}
/* Returns the strategy used: 0 greedy, 1..3 hash 0..2, 4 sorted search. */
int orc_apply_palette(const uint32_t* src, int src_stride, uint32_t* dst, int dst_stride,
                      const uint32_t* palette, int palette_size, int width, int height, int xbits) {
  uint8_t* tmp_row = (uint8_t*)malloc((size_t)width);
  uint16_t buffer[ORC_PALETTE_INV_SIZE];
  uint32_t sorted[256], idx_map[256], pal3[3] = {0, 0, 0};
  int strategy = 0, i, j, x, y;
  if (tmp_row == NULL) return -1;
  if (palette_size < 4) {                                                                   /* :1154 */
    for (i = 0; i < 3 && i < palette_size; ++i) pal3[i] = palette[i];
    /* entries past palette_size: the reference reads palette[] (MAX_PALETTE_SIZE words, zero beyond) */
  } else {
    for (i = 0; i < 3; ++i) {                                                               /* :1168-1184 */
      int use_LUT = 1;
      memset(buffer, 0xff, sizeof(buffer));
      for (j = 0; j < palette_size; ++j) {
        const uint32_t ind = palette_hash(i, palette[j]);
        if (buffer[ind] != 0xffffu) { use_LUT = 0; break; }
        buffer[ind] = (uint16_t)j;
      }
      if (use_LUT) break;
    }
    strategy = i + 1;
    if (i == 3) {                                                                           /* :1193-1199, palette.go:99-106 */
      memcpy(sorted, palette, sizeof(uint32_t) * (size_t)palette_size);
      qsort(sorted, (size_t)palette_size, sizeof(uint32_t), cmp_u32);
      for (j = 0; j < palette_size; ++j) idx_map[SearchColorNoIdxBounded(sorted, palette[j], palette_size)] = (uint32_t)j;
    }
  }
  for (y = 0; y < height; ++y) {                                                            /* APPLY_PALETTE_FOR :1130-1147 */
    /* The reference keeps a one-pixel cache (prev_pix / prev_idx, seeded with palette[0] -> 0) across
     * the whole image; the index is a pure function of the pixel, so the cache never changes a result. */
    uint32_t prev_pix = palette[0];
    int prev_idx = 0;
    const uint32_t* row = src + (size_t)y * src_stride;
    for (x = 0; x < width; ++x) {
      const uint32_t pix = row[x];
      if (pix != prev_pix) {
        if (strategy == 0) prev_idx = SearchColorGreedy(pal3, pix);
        else if (strategy <= 3) prev_idx = buffer[palette_hash(strategy - 1, pix)];
        else prev_idx = (int)idx_map[SearchColorNoIdxBounded(sorted, pix, palette_size)];
        prev_pix = pix;
      }
      tmp_row[x] = (uint8_t)prev_idx;
    }
    orc_bundle_color_map(tmp_row, width, xbits, dst + (size_t)y * dst_stride);
  }
  free(tmp_row);
  return strategy;
}

/* ---- predictors: pkg/libwebp/dsp/lossless.go:31-333 (row a-next, 8f-1) --- */
static inline uint32_t AddPixels(uint32_t a, uint32_t b) { /* common.go:110 */
  const uint32_t ag = (a & 0xff00ff00u) + (b & 0xff00ff00u);
  const uint32_t rb = (a & 0x00ff00ffu) + (b & 0x00ff00ffu);
  return (ag & 0xff00ff00u) | (rb & 0x00ff00ffu);
}
static inline uint32_t SubPixels(uint32_t a, uint32_t b) { /* common.go:117 */
  const uint32_t ag = 0x00ff00ffu + (a & 0xff00ff00u) - (b & 0xff00ff00u);
  const uint32_t rb = 0xff00ff00u + (a & 0x00ff00ffu) - (b & 0x00ff00ffu);
  return (ag & 0xff00ff00u) | (rb & 0x00ff00ffu);
}
static inline uint32_t Average2(uint32_t a0, uint32_t a1) {
  return (((a0 ^ a1) & 0xfefefefeu) >> 1) + (a0 & a1);
}
static inline uint32_t Average3(uint32_t a0, uint32_t a1, uint32_t a2) {
  return Average2(Average2(a0, a2), a1);
}
static inline uint32_t Average4(uint32_t a0, uint32_t a1, uint32_t a2,
                                uint32_t a3) {
  return Average2(Average2(a0, a1), Average2(a2, a3));
}
static inline uint32_t Clip255(uint32_t a) {
  if (a < 256) return a;
  return ~a >> 24; /* 0 when a is negative (wrapped), 255 when positive */
}
static inline int AddSubtractComponentFull(int a, int b, int c) {
  return (int)Clip255((uint32_t)(a + b - c));
}
static inline uint32_t ClampedAddSubtractFull(uint32_t c0, uint32_t c1,
                                              uint32_t c2) {
  const int a = AddSubtractComponentFull(c0 >> 24, c1 >> 24, c2 >> 24);
  const int r = AddSubtractComponentFull((c0 >> 16) & 0xff, (c1 >> 16) & 0xff,
                                         (c2 >> 16) & 0xff);
  const int g = AddSubtractComponentFull((c0 >> 8) & 0xff, (c1 >> 8) & 0xff,
                                         (c2 >> 8) & 0xff);
  const int b = AddSubtractComponentFull(c0 & 0xff, c1 & 0xff, c2 & 0xff);
  return ((uint32_t)a << 24) | (r << 16) | (g << 8) | b;
}
static inline int AddSubtractComponentHalf(int a, int b) {
  return (int)Clip255((uint32_t)(a + (a - b) / 2));
}
static inline uint32_t ClampedAddSubtractHalf(uint32_t c0, uint32_t c1,
                                              uint32_t c2) {
  const uint32_t ave = Average2(c0, c1);
  const int a = AddSubtractComponentHalf(ave >> 24, c2 >> 24);
  const int r = AddSubtractComponentHalf((ave >> 16) & 0xff, (c2 >> 16) & 0xff);
  const int g = AddSubtractComponentHalf((ave >> 8) & 0xff, (c2 >> 8) & 0xff);
  const int b = AddSubtractComponentHalf((ave >> 0) & 0xff, (c2 >> 0) & 0xff);
  return ((uint32_t)a << 24) | (r << 16) | (g << 8) | b;
}
static inline int Sub3(int a, int b, int c) {
  const int pb = b - c;
  const int pa = a - c;
  return abs(pb) - abs(pa);
}
static inline uint32_t Select(uint32_t a, uint32_t b, uint32_t c) {
  const int pa_minus_pb =
      Sub3((a >> 24), (b >> 24), (c >> 24)) +
      Sub3((a >> 16) & 0xff, (b >> 16) & 0xff, (c >> 16) & 0xff) +
      Sub3((a >> 8) & 0xff, (b >> 8) & 0xff, (c >> 8) & 0xff) +
      Sub3((a) & 0xff, (b) & 0xff, (c) & 0xff);
  return (pa_minus_pb <= 0) ? a : b;
}
/* VP8LPredictor0..13_C: left = out[x-1], top = upper + x (T=top[0],          */
/* TR=top[1], TL=top[-1]).                                                    */
static inline uint32_t Predict(int mode, uint32_t left, const uint32_t* top) {
  switch (mode) {
    case 0: return 0xff000000u;
    case 1: return left;
    case 2: return top[0];
    case 3: return top[1];
    case 4: return top[-1];
    case 5: return Average3(left, top[0], top[1]);
    case 6: return Average2(left, top[-1]);
    case 7: return Average2(left, top[0]);
    case 8: return Average2(top[-1], top[0]);
    case 9: return Average2(top[0], top[1]);
    case 10: return Average4(left, top[-1], top[0], top[1]);
    case 11: return Select(top[0], left, top[-1]);
    case 12: return ClampedAddSubtractFull(left, top[0], top[-1]);
    case 13: return ClampedAddSubtractHalf(left, top[0], top[-1]);
    default: return 0xff000000u; /* modes 14,15 alias mode 0 (lossless.go) */
  }
}
uint32_t orc_predict(int mode, uint32_t left, uint32_t tl, uint32_t t,
                     uint32_t tr) {
  const uint32_t top[3] = {tl, t, tr};
  return Predict(mode, left, top + 1);
}
/*
 * lossless.go:290-333 (PredictorInverseTransform_C) over a whole image
 * (y_start = 0 .. height); 'out' must not alias 'in'.
 */
void orc_predictor_inverse(int bits, int width, int height,
                           const uint32_t* data, const uint32_t* in,
                           uint32_t* out) {
  const int tiles_per_row = (int)SubSampleSize(width, bits);
  int x, y;
  /* first row: pixel 0 mode 0 (black), rest mode 1 (L) */
  out[0] = AddPixels(in[0], 0xff000000u);
  for (x = 1; x < width; ++x) out[x] = AddPixels(in[x], out[x - 1]);
  for (y = 1; y < height; ++y) {
    const uint32_t* irow = in + (size_t)y * width;
    uint32_t* orow = out + (size_t)y * width;
    const uint32_t* upper = orow - width;
    const uint32_t* modes = data + (y >> bits) * tiles_per_row;
    orow[0] = AddPixels(irow[0], upper[0]); /* mode 2 (T) */
    for (x = 1; x < width; ++x) {
      const int mode = (modes[x >> bits] >> 8) & 0xf;
      orow[x] = AddPixels(irow[x], Predict(mode, orow[x - 1], upper + x));
    }
  }
}
/* Forward residuals with GIVEN modes: lossless_enc.go:339-557 (PredictorSubN) */
/* driven as predictor_enc.go:666-712 does once modes are chosen.             */
void orc_predictor_forward(int bits, int width, int height,
                           const uint32_t* data, const uint32_t* argb,
                           uint32_t* out) {
  const int tiles_per_row = (int)SubSampleSize(width, bits);
  int x, y;
  out[0] = SubPixels(argb[0], 0xff000000u);
  for (x = 1; x < width; ++x) out[x] = SubPixels(argb[x], argb[x - 1]);
  for (y = 1; y < height; ++y) {
    const uint32_t* cur = argb + (size_t)y * width;
    const uint32_t* upper = cur - width;
    uint32_t* orow = out + (size_t)y * width;
    const uint32_t* modes = data + (y >> bits) * tiles_per_row;
    orow[0] = SubPixels(cur[0], upper[0]);
    for (x = 1; x < width; ++x) {
      const int mode = (modes[x >> bits] >> 8) & 0xf;
      orow[x] = SubPixels(cur[x], Predict(mode, cur[x - 1], upper + x));
    }
  }
}

/*
 * lossless.go:511-547 (VP8LInverseTransform) for one transform over rows
 * [row_start,row_end); out-of-place.  vtype: pkg/vp8/enums.go:33-40.
 * 'in'/'out' point at row_start.  The predictor case needs out[-width..-1]
 * readable when row_start > 0, exactly as the reference does.
 */
enum { PREDICTOR_TRANSFORM = 0, CROSS_COLOR_TRANSFORM = 1,
       SUBTRACT_GREEN_TRANSFORM = 2, COLOR_INDEXING_TRANSFORM = 3 };
int orc_inverse_transform(int vtype, int bits, int xsize, int ysize,
                          const uint32_t* data, int row_start, int row_end,
                          const uint32_t* in, uint32_t* out) {
  if (!(row_start < row_end) || row_end > ysize) return 0;
  switch (vtype) {
    case SUBTRACT_GREEN_TRANSFORM:
      orc_add_green(in, (row_end - row_start) * xsize, out);
      return 1;
    case CROSS_COLOR_TRANSFORM:
      orc_color_space_inverse(bits, xsize, data, row_start, row_end, in, out);
      return 1;
    case COLOR_INDEXING_TRANSFORM:
      orc_color_index_inverse(bits, xsize, data, row_start, row_end, in, out);
      return 1;
    case PREDICTOR_TRANSFORM:
      if (row_start != 0) return 0; /* whole-image use only in this oracle */
      orc_predictor_inverse(bits, xsize, row_end, data, in, out);
      return 1;
    default:
      return 0;
  }
}

/* ------------------------------------------------------------------------- */
/* Synthetic input generator: identical on host and device.      */
static inline uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
uint32_t orc_synth_word(uint32_t seed, uint32_t img, uint32_t plane,
                        uint32_t y, uint32_t x) {
  return lowbias32(seed ^ (img * 0x9E3779B1u) ^ (plane * 0x85EBCA77u) ^
                   (y * 0xC2B2AE3Du) ^ (x * 0x27D4EB2Fu));
}
/* Shapes the top byte (alpha of an LE RGBA pixel); modes as B2_SYNTH_* in    */
/* include/b2webp.h: 0 raw, 1 opaque, 2 ~30% translucent, 3 binary.           */
uint32_t orc_synth_alpha(uint32_t word, int alpha_mode) {
  uint32_t h2;
  if (alpha_mode == 0) return word;
  if (alpha_mode == 1) return word | 0xff000000u;
  h2 = lowbias32(word ^ 0x9E3779B9u);
  if (alpha_mode == 2) {
    if ((h2 & 0xff) >= 77) return word | 0xff000000u;
    return (word & 0x00ffffffu) | (h2 & 0xff000000u);
  }
  return (word & 0x00ffffffu) | ((h2 & 1) ? 0xff000000u : 0u);
}
/* One u32 per (y, x) position; rows of 'words_per_row' u32 at 'stride' bytes. */
void orc_synth_fill_u32(uint32_t seed, uint32_t img, uint32_t plane,
                        uint32_t* dst, int words_per_row, int rows,
                        int stride_bytes, int alpha_mode) {
  int y, x;
  for (y = 0; y < rows; ++y) {
    uint32_t* row = (uint32_t*)((uint8_t*)dst + (ptrdiff_t)y * stride_bytes);
    for (x = 0; x < words_per_row; ++x) {
      row[x] = orc_synth_alpha(
          orc_synth_word(seed, img, plane, (uint32_t)y, (uint32_t)x),
          alpha_mode);
    }
  }
}

/* ------------------------------------------------------------------------- */
/* Multi-threaded batch drivers: the CPU baseline timed by bench.py.          */
typedef struct {
  int kind;           /* 0 import, 1 upsample, 2 vp8l chain A, 3 index */
  int first, last;    /* image range */
  const void* args;
} BatchSlice;

typedef struct {
  const uint8_t* rgba; size_t rgba_img_bytes; int rgba_stride;
  int width, height;
  uint8_t* y; uint8_t* u; uint8_t* v; uint8_t* a;
  size_t y_img_bytes, uv_img_bytes;
  int divide_mode;
  int* has_alpha;
} ImportBatch;

static void* import_worker(void* p) {
  const BatchSlice* s = (const BatchSlice*)p;
  const ImportBatch* b = (const ImportBatch*)s->args;
  int i;
  const int uvw = (b->width + 1) >> 1;
  for (i = s->first; i < s->last; ++i) {
    b->has_alpha[i] = orc_import_yuva(
        b->rgba + (size_t)i * b->rgba_img_bytes, 0, 1, 2, 3, 4, b->rgba_stride,
        b->width, b->height, b->y + (size_t)i * b->y_img_bytes, b->width,
        b->u + (size_t)i * b->uv_img_bytes, b->v + (size_t)i * b->uv_img_bytes,
        uvw, b->a ? b->a + (size_t)i * b->y_img_bytes : NULL, b->width,
        b->divide_mode);
  }
  return NULL;
}

typedef struct {
  int mode;
  const uint8_t* y; const uint8_t* u; const uint8_t* v; const uint8_t* a;
  size_t y_img_bytes, uv_img_bytes;
  int width, height;
  uint8_t* dst; size_t dst_img_bytes; int dst_stride;
} UpsampleBatch;

static void* upsample_worker(void* p) {
  const BatchSlice* s = (const BatchSlice*)p;
  const UpsampleBatch* b = (const UpsampleBatch*)s->args;
  int i;
  const int uvw = (b->width + 1) >> 1;
  for (i = s->first; i < s->last; ++i) {
    orc_yuva_to_rgba(b->mode, b->y + (size_t)i * b->y_img_bytes, b->width,
                     b->u + (size_t)i * b->uv_img_bytes,
                     b->v + (size_t)i * b->uv_img_bytes, uvw,
                     b->a ? b->a + (size_t)i * b->y_img_bytes : NULL, b->width,
                     b->width, b->height,
                     b->dst + (size_t)i * b->dst_img_bytes, b->dst_stride);
  }
  return NULL;
}

static int run_slices(void* (*fn)(void*), int kind, const void* args, int n,
                      int n_threads) {
  pthread_t* th;
  BatchSlice* sl;
  int t, started = 0;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n) n_threads = n > 0 ? n : 1;
  th = (pthread_t*)malloc(sizeof(*th) * n_threads);
  sl = (BatchSlice*)malloc(sizeof(*sl) * n_threads);
  if (!th || !sl) { free(th); free(sl); return -1; }
  for (t = 0; t < n_threads; ++t) {
    sl[t].kind = kind;
    sl[t].first = (int)((long long)n * t / n_threads);
    sl[t].last = (int)((long long)n * (t + 1) / n_threads);
    sl[t].args = args;
    if (pthread_create(&th[t], NULL, fn, &sl[t]) != 0) break;
    ++started;
  }
  for (t = 0; t < started; ++t) pthread_join(th[t], NULL);
  free(th); free(sl);
  return started == n_threads ? 0 : -1;
}

/* n tightly packed RGBA images -> tightly packed Y/U/V/A planes per image. */
int orc_import_batch_mt(const uint8_t* rgba, int n, int width, int height,
                        uint8_t* y, uint8_t* u, uint8_t* v, uint8_t* a,
                        int* has_alpha, int divide_mode, int n_threads) {
  ImportBatch b;
  orc_init();
  b.rgba = rgba; b.rgba_stride = 4 * width;
  b.rgba_img_bytes = (size_t)4 * width * height;
  b.width = width; b.height = height;
  b.y = y; b.u = u; b.v = v; b.a = a;
  b.y_img_bytes = (size_t)width * height;
  b.uv_img_bytes = (size_t)((width + 1) >> 1) * ((height + 1) >> 1);
  b.divide_mode = divide_mode; b.has_alpha = has_alpha;
  return run_slices(import_worker, 0, &b, n, n_threads);
}
int orc_upsample_batch_mt(int mode, const uint8_t* y, const uint8_t* u,
                          const uint8_t* v, const uint8_t* a, int n, int width,
                          int height, uint8_t* dst, int n_threads) {
  UpsampleBatch b;
  b.mode = mode; b.y = y; b.u = u; b.v = v; b.a = a;
  b.width = width; b.height = height;
  b.y_img_bytes = (size_t)width * height;
  b.uv_img_bytes = (size_t)((width + 1) >> 1) * ((height + 1) >> 1);
  b.dst = dst; b.dst_stride = orc_mode_bpp(mode) * width;
  b.dst_img_bytes = (size_t)b.dst_stride * height;
  return run_slices(upsample_worker, 1, &b, n, n_threads);
}

typedef struct {
  int op; /* 0: crosscolor^-1 then add-green; 1: color-index^-1 */
  int bits, width, height;
  const uint32_t* src; size_t src_img_px;
  const uint32_t* data; size_t data_img_px;
  uint32_t* dst; size_t dst_img_px;
  uint32_t* tmp;
} Vp8lBatch;

static void* vp8l_worker(void* p) {
  const BatchSlice* s = (const BatchSlice*)p;
  const Vp8lBatch* b = (const Vp8lBatch*)s->args;
  int i;
  for (i = s->first; i < s->last; ++i) {
    const uint32_t* src = b->src + (size_t)i * b->src_img_px;
    const uint32_t* data = b->data + (size_t)i * b->data_img_px;
    uint32_t* dst = b->dst + (size_t)i * b->dst_img_px;
    if (b->op == 0) {
      /* ApplyInverseTransforms order (vp8l_dec.c.go:710-729): last read     */
      /* first => cross-colour^-1, then add-green; second step in place.     */
      orc_color_space_inverse(b->bits, b->width, data, 0, b->height, src, dst);
      orc_add_green(dst, b->width * b->height, dst);
    } else {
      orc_color_index_inverse(b->bits, b->width, data, 0, b->height, src, dst);
    }
  }
  return NULL;
}
int orc_vp8l_batch_mt(int op, int bits, int width, int height,
                      const uint32_t* src, size_t src_img_px,
                      const uint32_t* data, size_t data_img_px, uint32_t* dst,
                      int n, int n_threads) {
  Vp8lBatch b;
  b.op = op; b.bits = bits; b.width = width; b.height = height;
  b.src = src; b.src_img_px = src_img_px;
  b.data = data; b.data_img_px = data_img_px;
  b.dst = dst; b.dst_img_px = (size_t)width * height; b.tmp = NULL;
  return run_slices(vp8l_worker, 2, &b, n, n_threads);
}
