/*
 * oracle_tools.c -- CPU restatement of the picture pre-processing that follows the
 * import in the encoder: WebPCleanupTransparentArea,
 * WebPReplaceTransparentPixels, WebPBlendAlpha.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.c).  Pinned against the libwebp 1.6.0 binary,
 * which exports WebPCleanupTransparentArea and WebPBlendAlpha
 * (tests/golden/make_golden_tools.py).  Later sections carry their own pin in their
 * headers (animation: WebPAnimDecoder; alpha filters: ALPH chunks of real files; rescaler:
 * WebPPictureRescale; level smoothing: WebPDecode with alpha_dithering_strength, stock
 * DFIX = 0 build).  File:line citations are relative to the go-webp tree.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* pkg/libwebp/enc/picture_tools_enc.c.go:29-30 */
#define SIZE 8
#define SIZE2 (SIZE / 2)

/* picture_tools_enc.c.go:31-42 */
static int IsTransparentARGBArea(const uint32_t* ptr, int stride, int size) {
  for (int y = 0; y < size; ++y) {
    for (int x = 0; x < size; ++x) {
      if (ptr[x] & 0xff000000u) return 0;
    }
    ptr += stride;
  }
  return 1;
}
/* picture_tools_enc.c.go:44-50 */
static void Flatten(uint8_t* ptr, int v, int stride, int size) {
  for (int y = 0; y < size; ++y) {
    memset(ptr, v, (size_t)size);
    ptr += stride;
  }
}
/* picture_tools_enc.c.go:52-58 */
static void FlattenARGB(uint32_t* ptr, uint32_t v, int stride, int size) {
  for (int y = 0; y < size; ++y) {
    for (int x = 0; x < size; ++x) ptr[x] = v;
    ptr += stride;
  }
}
/* picture_tools_enc.c.go:62-91 */
static int SmoothenBlock(const uint8_t* a_ptr, int a_stride, uint8_t* y_ptr, int y_stride, int width, int height) {
  int sum = 0, count = 0;
  const uint8_t* alpha_ptr = a_ptr;
  uint8_t* luma_ptr = y_ptr;
  for (int y = 0; y < height; ++y) {
    for (int x = 0; x < width; ++x) {
      if (alpha_ptr[x] != 0) {
        ++count;
        sum += luma_ptr[x];
      }
    }
    alpha_ptr += a_stride;
    luma_ptr += y_stride;
  }
  if (count > 0 && count < width * height) {
    const uint8_t avg_u8 = (uint8_t)(sum / count);
    alpha_ptr = a_ptr;
    luma_ptr = y_ptr;
    for (int y = 0; y < height; ++y) {
      for (int x = 0; x < width; ++x) {
        if (alpha_ptr[x] == 0) luma_ptr[x] = avg_u8;
      }
      alpha_ptr += a_stride;
      luma_ptr += y_stride;
    }
  }
  return (count == 0);
}

/* WebPCleanupTransparentArea, use_argb branch: picture_tools_enc.c.go:116-133.
 * argb_stride in pixels. */
void orc_cleanup_transparent_argb(uint32_t* argb, int argb_stride, int width, int height) {
  const int w = width / SIZE, h = height / SIZE;
  uint32_t argb_value = 0;
  for (int y = 0; y < h; ++y) {
    int need_reset = 1;
    for (int x = 0; x < w; ++x) {
      const int off = (y * argb_stride + x) * SIZE;
      if (IsTransparentARGBArea(argb + off, argb_stride, SIZE)) {
        if (need_reset) {
          argb_value = argb[off];
          need_reset = 0;
        }
        FlattenARGB(argb + off, argb_value, argb_stride, SIZE);
      } else {
        need_reset = 1;
      }
    }
  }
}

/* WebPCleanupTransparentArea, YUVA branch: picture_tools_enc.c.go:134-183 */
void orc_cleanup_transparent_yuva(uint8_t* y_ptr, uint8_t* u_ptr, uint8_t* v_ptr, const uint8_t* a_ptr,
                                  int y_stride, int uv_stride, int a_stride, int width, int height) {
  int values[3] = {0, 0, 0};
  int x, y;
  if (a_ptr == NULL || y_ptr == NULL || u_ptr == NULL || v_ptr == NULL) return;
  for (y = 0; y + SIZE <= height; y += SIZE) {
    int need_reset = 1;
    for (x = 0; x + SIZE <= width; x += SIZE) {
      if (SmoothenBlock(a_ptr + x, a_stride, y_ptr + x, y_stride, SIZE, SIZE)) {
        if (need_reset) {
          values[0] = y_ptr[x];
          values[1] = u_ptr[x >> 1];
          values[2] = v_ptr[x >> 1];
          need_reset = 0;
        }
        Flatten(y_ptr + x, values[0], y_stride, SIZE);
        Flatten(u_ptr + (x >> 1), values[1], uv_stride, SIZE2);
        Flatten(v_ptr + (x >> 1), values[2], uv_stride, SIZE2);
      } else {
        need_reset = 1;
      }
    }
    if (x < width) SmoothenBlock(a_ptr + x, a_stride, y_ptr + x, y_stride, width - x, SIZE);
    a_ptr += SIZE * a_stride;
    y_ptr += SIZE * y_stride;
    u_ptr += SIZE2 * uv_stride;
    v_ptr += SIZE2 * uv_stride;
  }
  if (y < height) {
    const int sub_height = height - y;
    for (x = 0; x + SIZE <= width; x += SIZE) {
      SmoothenBlock(a_ptr + x, a_stride, y_ptr + x, y_stride, SIZE, sub_height);
    }
    if (x < width) SmoothenBlock(a_ptr + x, a_stride, y_ptr + x, y_stride, width - x, sub_height);
  }
}

/* WebPReplaceTransparentPixels (picture_tools_enc.c.go:95-108) over WebPAlphaReplace
 * (pkg/color/colorspace/alpha/opague.go:47-53; upstream AlphaReplace_C: a pixel whose
 * alpha byte is 0 becomes 'color' with alpha forced to 0). */
void orc_replace_transparent(uint32_t* argb, int argb_stride, int width, int height, uint32_t color) {
  color &= 0xffffffu;
  for (int y = 0; y < height; ++y) {
    for (int x = 0; x < width; ++x) {
      if ((argb[x] >> 24) == 0) argb[x] = color;
    }
    argb += argb_stride;
  }
}

/* picture_tools_enc.c.go:191-194 */
#define BLEND(V0, V1, ALPHA) ((((V0) * (255 - (ALPHA)) + (V1) * (ALPHA)) * 0x101 + 256) >> 16)
#define BLEND_10BIT(V0, V1, ALPHA) ((((V0) * (1020 - (ALPHA)) + (V1) * (ALPHA)) * 0x101 + 1024) >> 18)

/* pkg/color/yuv/conversion.go:51-69 */
enum { T_YUV_FIX = 16, T_YUV_HALF = 1 << 15 };
static int t_ClipUV(int uv, int rounding) {
  uv = (uv + rounding + (128 << (T_YUV_FIX + 2))) >> (T_YUV_FIX + 2);
  return ((uv & ~0xff) == 0) ? uv : (uv < 0) ? 0 : 255;
}
static int t_RGBToY(int r, int g, int b, int rounding) {
  return (16839 * r + 33059 * g + 6420 * b + rounding + (16 << T_YUV_FIX)) >> T_YUV_FIX;
}
static int t_RGBToU(int r, int g, int b, int rounding) { return t_ClipUV(-9719 * r - 19081 * g + 28800 * b, rounding); }
static int t_RGBToV(int r, int g, int b, int rounding) { return t_ClipUV(28800 * r - 24116 * g - 4684 * b, rounding); }

/* WebPBlendAlpha, YUVA branch: picture_tools_enc.c.go:208-252.  The caller checks
 * has_alpha / a != NULL (:218-219). */
void orc_blend_alpha_yuva(uint8_t* y_ptr, uint8_t* u_ptr, uint8_t* v_ptr, uint8_t* a_ptr, int y_stride, int uv_stride,
                          int a_stride, int width, int height, uint32_t background_rgb) {
  const int red = (background_rgb >> 16) & 0xff, green = (background_rgb >> 8) & 0xff, blue = background_rgb & 0xff;
  const int uv_width = width >> 1;   /* omit last pixel during u/v loop */
  const int Y0 = t_RGBToY(red, green, blue, T_YUV_HALF);
  /* VP8RGBToU/V expects the u/v values summed over four pixels */
  const int U0 = t_RGBToU(4 * red, 4 * green, 4 * blue, 4 * T_YUV_HALF);
  const int V0 = t_RGBToV(4 * red, 4 * green, 4 * blue, 4 * T_YUV_HALF);
  int x, y;
  if (a_ptr == NULL) return;
  for (y = 0; y < height; ++y) {
    for (x = 0; x < width; ++x) {
      const uint8_t alpha = a_ptr[x];
      if (alpha < 0xff) y_ptr[x] = (uint8_t)BLEND(Y0, y_ptr[x], alpha);
    }
    if ((y & 1) == 0) {
      const uint8_t* const a_ptr2 = (y + 1 == height) ? a_ptr : a_ptr + a_stride;
      for (x = 0; x < uv_width; ++x) {
        const uint32_t alpha = a_ptr[2 * x + 0] + a_ptr[2 * x + 1] + a_ptr2[2 * x + 0] + a_ptr2[2 * x + 1];
        u_ptr[x] = (uint8_t)BLEND_10BIT(U0, u_ptr[x], alpha);
        v_ptr[x] = (uint8_t)BLEND_10BIT(V0, v_ptr[x], alpha);
      }
      if (width & 1) {   /* rightmost pixel */
        const uint32_t alpha = 2 * (a_ptr[2 * x + 0] + a_ptr2[2 * x + 0]);
        u_ptr[x] = (uint8_t)BLEND_10BIT(U0, u_ptr[x], alpha);
        v_ptr[x] = (uint8_t)BLEND_10BIT(V0, v_ptr[x], alpha);
      }
    } else {
      u_ptr += uv_stride;
      v_ptr += uv_stride;
    }
    memset(a_ptr, 0xff, (size_t)width);   /* reset alpha value to opaque */
    a_ptr += a_stride;
    y_ptr += y_stride;
  }
}

/* WebPBlendAlpha, ARGB branch: picture_tools_enc.c.go:253-276 */
void orc_blend_alpha_argb(uint32_t* argb, int argb_stride, int width, int height, uint32_t background_rgb) {
  const int red = (background_rgb >> 16) & 0xff, green = (background_rgb >> 8) & 0xff, blue = background_rgb & 0xff;
  const uint32_t background = 0xff000000u | ((uint32_t)red << 16) | ((uint32_t)green << 8) | (uint32_t)blue;
  for (int y = 0; y < height; ++y) {
    for (int x = 0; x < width; ++x) {
      const int alpha = (argb[x] >> 24) & 0xff;
      if (alpha != 0xff) {
        if (alpha > 0) {
          int r = (argb[x] >> 16) & 0xff, g = (argb[x] >> 8) & 0xff, b = argb[x] & 0xff;
          r = BLEND(red, r, alpha);
          g = BLEND(green, g, alpha);
          b = BLEND(blue, b, alpha);
          argb[x] = 0xff000000u | ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b;
        } else {
          argb[x] = background;
        }
      }
    }
    argb += argb_stride;
  }
}

/* Import() into an ARGB picture: pkg/libwebp/enc/picture_csp_enc.c.go:138-181.
 * import_alpha: BGRA rows are copied (:149-154), RGBA rows go through
 * VP8LConvertBGRAToRGBA (:165-166, pkg/libwebp/dsp/lossless.go:549-560: swaps bytes 0 and
 * 2); otherwise WebPPackRGB / PackRGB_C (pkg/libwebp/dsp/alpha_processing.go:227-233).
 * Every branch yields A << 24 | R << 16 | G << 8 | B with A = 0xff when not imported. */
int orc_import_to_argb(const uint8_t* rgb, int rgb_stride, int step, int swap_rb, int import_alpha, int width, int height,
                       uint32_t* dst, int dst_stride) {
  const int a_stride = rgb_stride < 0 ? -rgb_stride : rgb_stride;
  if (a_stride < (import_alpha ? 4 : 3) * width) return 0;   /* :132 */
  for (int y = 0; y < height; ++y) {
    const uint8_t* r_ptr = rgb + (swap_rb ? 2 : 0);
    const uint8_t* g_ptr = rgb + 1;
    const uint8_t* b_ptr = rgb + (swap_rb ? 0 : 2);
    if (import_alpha) {
      for (int i = 0; i < width; ++i) {
        dst[i] = ((uint32_t)rgb[4 * i + 3] << 24) | ((uint32_t)r_ptr[4 * i] << 16) | ((uint32_t)g_ptr[4 * i] << 8) | b_ptr[4 * i];
      }
    } else {
      int offset = 0;
      for (int i = 0; i < width; ++i) {
        dst[i] = 0xff000000u | ((uint32_t)r_ptr[offset] << 16) | ((uint32_t)g_ptr[offset] << 8) | b_ptr[offset];
        offset += step;
      }
    }
    rgb += rgb_stride;
    dst += dst_stride;
  }
  return 1;
}

/* ---- animation: canvas blending and frame comparison ------------------ */
/* pkg/libwebp/demux/anim_decode.go:195-203.  Pixels are uint32 with alpha in bits 24-31
 * (CHANNEL_SHIFT(3) on little endian); the three colour channels are treated alike. */
static uint8_t BlendChannelNonPremult(uint32_t src, uint8_t src_a, uint32_t dst, uint8_t dst_a, uint32_t scale, int shift) {
  const uint8_t src_channel = (src >> shift) & 0xff;
  const uint8_t dst_channel = (dst >> shift) & 0xff;
  const uint32_t blend_unscaled = src_channel * src_a + dst_channel * dst_a;
  return (uint8_t)((blend_unscaled * scale) >> 24);
}
/* anim_decode.go:205-232 */
static uint32_t BlendPixelNonPremult(uint32_t src, uint32_t dst) {
  const uint8_t src_a = (src >> 24) & 0xff;
  if (src_a == 0) {
    return dst;
  } else {
    const uint8_t dst_a = (dst >> 24) & 0xff;
    const uint8_t dst_factor_a = (uint8_t)((dst_a * (256 - src_a)) >> 8);
    const uint8_t blend_a = (uint8_t)(src_a + dst_factor_a);
    const uint32_t scale = (1u << 24) / blend_a;
    const uint8_t blend_r = BlendChannelNonPremult(src, src_a, dst, dst_factor_a, scale, 0);
    const uint8_t blend_g = BlendChannelNonPremult(src, src_a, dst, dst_factor_a, scale, 8);
    const uint8_t blend_b = BlendChannelNonPremult(src, src_a, dst, dst_factor_a, scale, 16);
    return ((uint32_t)blend_r << 0) | ((uint32_t)blend_g << 8) | ((uint32_t)blend_b << 16) | ((uint32_t)blend_a << 24);
  }
}
/* anim_decode.go:245-257 */
static uint32_t ChannelwiseMultiply(uint32_t pix, uint32_t scale) {
  const uint32_t mask = 0x00FF00FF;
  const uint32_t rb = ((pix & mask) * scale) >> 8;
  const uint32_t ag = ((pix >> 8) & mask) * scale;
  return (rb & mask) | (ag & ~mask);
}
static uint32_t BlendPixelPremult(uint32_t src, uint32_t dst) {
  const uint8_t src_a = (src >> 24) & 0xff;
  return src + ChannelwiseMultiply(dst, 256 - src_a);
}
/* BlendPixelRowNonPremult / BlendPixelRowPremult (anim_decode.go:234-243, 259-270) over
 * `rows` rows; strides in pixels.  src is blended in place over dst. */
void orc_anim_blend_rows(uint32_t* src, int src_stride, const uint32_t* dst, int dst_stride, int width, int rows, int premult) {
  for (int y = 0; y < rows; ++y) {
    for (int i = 0; i < width; ++i) {
      const uint8_t src_alpha = (src[i] >> 24) & 0xff;
      if (src_alpha != 0xff) src[i] = premult ? BlendPixelPremult(src[i], dst[i]) : BlendPixelNonPremult(src[i], dst[i]);
    }
    src += src_stride;
    dst += dst_stride;
  }
}
/* pkg/libwebp/mux/anim_encode.go:318-332 */
int orc_compare_pixels_lossless(const uint32_t* src, int src_step, const uint32_t* dst, int dst_step, int length) {
  while (length-- > 0) {
    if (*src != *dst) return 0;
    src += src_step;
    dst += dst_step;
  }
  return 1;
}
/* anim_encode.go:336-351 */
static int PixelsAreSimilar(uint32_t src, uint32_t dst, int max_allowed_diff) {
  const int src_a = (src >> 24) & 0xff, src_r = (src >> 16) & 0xff, src_g = (src >> 8) & 0xff, src_b = src & 0xff;
  const int dst_a = (dst >> 24) & 0xff, dst_r = (dst >> 16) & 0xff, dst_g = (dst >> 8) & 0xff, dst_b = dst & 0xff;
  return (src_a == dst_a) && (abs(src_r - dst_r) * dst_a <= (max_allowed_diff * 255)) &&
         (abs(src_g - dst_g) * dst_a <= (max_allowed_diff * 255)) && (abs(src_b - dst_b) * dst_a <= (max_allowed_diff * 255));
}
/* anim_encode.go:355-367 */
int orc_compare_pixels_lossy(const uint32_t* src, int src_step, const uint32_t* dst, int dst_step, int length,
                             int max_allowed_diff) {
  while (length-- > 0) {
    if (!PixelsAreSimilar(*src, *dst, max_allowed_diff)) return 0;
    src += src_step;
    dst += dst_step;
  }
  return 1;
}

/* ---- alpha-plane filters ---------------------------------------------- */
/* GradientPredictor_C, pkg/libwebp/dsp/filters.go (clip(a + b - c) to [0, 255]) */
static int GradientPredictor(uint8_t a, uint8_t b, uint8_t c) {
  const int g = a + b - c;
  return ((g & ~0xff) == 0) ? g : (g < 0) ? 0 : 255;
}
/* filters.go:138-145 */
static void HorizontalUnfilter(const uint8_t* prev, const uint8_t* in, uint8_t* out, int width) {
  uint8_t pred = (prev == NULL) ? 0 : prev[0];
  for (int i = 0; i < width; ++i) {
    out[i] = (uint8_t)(pred + in[i]);
    pred = out[i];
  }
}
/* filters.go:147-155 */
static void VerticalUnfilter(const uint8_t* prev, const uint8_t* in, uint8_t* out, int width) {
  if (prev == NULL) {
    HorizontalUnfilter(NULL, in, out, width);
  } else {
    for (int i = 0; i < width; ++i) out[i] = (uint8_t)(prev[i] + in[i]);
  }
}
/* filters.go:157-171 */
static void GradientUnfilter(const uint8_t* prev, const uint8_t* in, uint8_t* out, int width) {
  if (prev == NULL) {
    HorizontalUnfilter(NULL, in, out, width);
  } else {
    uint8_t top = prev[0], top_left = top, left = top;
    for (int i = 0; i < width; ++i) {
      top = prev[i]; /* need to read this first, in case prev==out */
      left = (uint8_t)(in[i] + GradientPredictor(left, top, top_left));
      top_left = top;
      out[i] = left;
    }
  }
}
/* The row loop of ALPHDecode / AlphaApplyFilter (pkg/vp8/vp8l_dec.c.go:802-814): every row
 * is unfiltered against the previous OUTPUT row.  filter: 0 none, 1 horizontal, 2 vertical,
 * 3 gradient (WEBP_FILTER_TYPE). */
void orc_alpha_unfilter(int filter, const uint8_t* in, int in_stride, uint8_t* out, int out_stride, int width, int height) {
  const uint8_t* prev = NULL;
  for (int y = 0; y < height; ++y) {
    const uint8_t* src = in + (ptrdiff_t)y * in_stride;
    uint8_t* dst = out + (ptrdiff_t)y * out_stride;
    switch (filter) {
      case 1: HorizontalUnfilter(prev, src, dst, width); break;
      case 2: VerticalUnfilter(prev, src, dst, width); break;
      case 3: GradientUnfilter(prev, src, dst, width); break;
      default: memcpy(dst, src, (size_t)width); break;
    }
    prev = dst;
  }
}
/* Forward filters: DoHorizontalFilter_C / DoVerticalFilter_C / DoGradientFilter_C
 * (filters.go:24-112): the first row is predicted from the left (first pixel as is), the
 * leftmost pixel of the other rows from above. */
void orc_alpha_filter(int filter, const uint8_t* in, int in_stride, uint8_t* out, int out_stride, int width, int height) {
  for (int y = 0; y < height; ++y) {
    const uint8_t* cur = in + (ptrdiff_t)y * in_stride;
    const uint8_t* up = cur - in_stride;
    uint8_t* dst = out + (ptrdiff_t)y * out_stride;
    if (filter == 0) { memcpy(dst, cur, (size_t)width); continue; }
    dst[0] = (y == 0) ? cur[0] : (uint8_t)(cur[0] - up[0]);
    for (int x = 1; x < width; ++x) {
      int pred;
      if (y == 0 || filter == 1) pred = cur[x - 1];
      else if (filter == 2) pred = up[x];
      else pred = GradientPredictor(cur[x - 1], up[x], up[x - 1]);
      dst[x] = (uint8_t)(cur[x] - pred);
    }
  }
}

/* ---- rescaler ---------------------------------------------------------- */
/* pkg/libwebp/utils/rescaler_utils.h.go (WebPRescaler, WEBP_RESCALER_RFIX = 32) */
typedef uint32_t rescaler_t;
#define RS_RFIX 32
#define RS_ONE (1ull << RS_RFIX)
#define RS_FRAC(x, y) ((uint32_t)(((uint64_t)(x) << RS_RFIX) / (y)))
#define RS_ROUNDER (RS_ONE >> 1)
#define RS_MULT_FIX(x, y) (((uint64_t)(x) * (y) + RS_ROUNDER) >> RS_RFIX)
#define RS_MULT_FIX_FLOOR(x, y) (((uint64_t)(x) * (y)) >> RS_RFIX)
typedef struct {
  int x_expand, y_expand, num_channels;
  uint32_t fx_scale, fy_scale, fxy_scale;
  int y_accum, y_add, y_sub, x_add, x_sub;
  int src_width, src_height, dst_width, dst_height, src_y, dst_y;
  uint8_t* dst;
  int dst_stride;
  rescaler_t *irow, *frow;
} Rescaler;
/* pkg/libwebp/utils/rescaler_utils.c.go:31-86 */
static void RescalerInit(Rescaler* r, int src_width, int src_height, uint8_t* dst, int dst_width, int dst_height,
                         int dst_stride, int num_channels, rescaler_t* work) {
  const int x_add = src_width, x_sub = dst_width, y_add = src_height, y_sub = dst_height;
  r->x_expand = (src_width < dst_width);
  r->y_expand = (src_height < dst_height);
  r->src_width = src_width; r->src_height = src_height; r->dst_width = dst_width; r->dst_height = dst_height;
  r->src_y = 0; r->dst_y = 0; r->dst = dst; r->dst_stride = dst_stride; r->num_channels = num_channels;
  r->irow = work;
  r->frow = work + num_channels * dst_width;
  memset(work, 0, sizeof(*work) * 2 * (size_t)dst_width * num_channels);
  r->x_add = r->x_expand ? (x_sub - 1) : x_add;
  r->x_sub = r->x_expand ? (x_add - 1) : x_sub;
  r->fx_scale = 0;
  if (!r->x_expand) r->fx_scale = RS_FRAC(1, r->x_sub);
  r->y_add = r->y_expand ? (y_add - 1) : y_add;
  r->y_sub = r->y_expand ? (y_sub - 1) : y_sub;
  r->y_accum = r->y_expand ? r->y_sub : r->y_add;
  r->fxy_scale = 0;
  if (!r->y_expand) {
    const uint64_t num = (uint64_t)dst_height * RS_ONE;
    const uint64_t den = (uint64_t)r->x_add * r->y_add;
    const uint64_t ratio = num / den;
    r->fxy_scale = (ratio != (uint32_t)ratio) ? 0 : (uint32_t)ratio;
    r->fy_scale = RS_FRAC(1, r->y_sub);
  } else {
    r->fy_scale = RS_FRAC(1, r->x_add);
  }
}
/* pkg/libwebp/dsp/rescaler.c.go:39-69 */
static void ImportRowExpand(Rescaler* wrk, const uint8_t* src) {
  const int x_stride = wrk->num_channels, x_out_max = wrk->dst_width * wrk->num_channels;
  for (int channel = 0; channel < x_stride; ++channel) {
    int x_in = channel, x_out = channel;
    int accum = wrk->x_add;
    rescaler_t left = (rescaler_t)src[x_in];
    rescaler_t right = (wrk->src_width > 1) ? (rescaler_t)src[x_in + x_stride] : left;
    x_in += x_stride;
    while (1) {
      wrk->frow[x_out] = right * wrk->x_add + (left - right) * accum;
      x_out += x_stride;
      if (x_out >= x_out_max) break;
      accum -= wrk->x_sub;
      if (accum < 0) {
        left = right;
        x_in += x_stride;
        right = (rescaler_t)src[x_in];
        accum += wrk->x_add;
      }
    }
  }
}
/* rescaler.c.go:71-105 */
static void ImportRowShrink(Rescaler* wrk, const uint8_t* src) {
  const int x_stride = wrk->num_channels, x_out_max = wrk->dst_width * wrk->num_channels;
  for (int channel = 0; channel < x_stride; ++channel) {
    int x_in = channel, x_out = channel;
    uint32_t sum = 0;
    int accum = 0;
    while (x_out < x_out_max) {
      uint32_t base = 0;
      accum += wrk->x_add;
      while (accum > 0) {
        accum -= wrk->x_sub;
        base = src[x_in];
        sum += base;
        x_in += x_stride;
      }
      {
        const rescaler_t frac = base * (-accum);
        wrk->frow[x_out] = sum * wrk->x_sub - frac;
        sum = (int)RS_MULT_FIX(frac, wrk->fx_scale);
      }
      x_out += x_stride;
    }
  }
}
/* rescaler.c.go:110-136 */
static void ExportRowExpand(Rescaler* wrk) {
  uint8_t* dst = wrk->dst;
  const int x_out_max = wrk->dst_width * wrk->num_channels;
  if (wrk->y_accum == 0) {
    for (int x_out = 0; x_out < x_out_max; ++x_out) {
      const uint32_t J = wrk->frow[x_out];
      const int v = (int)RS_MULT_FIX(J, wrk->fy_scale);
      dst[x_out] = (v > 255) ? 255u : (uint8_t)v;
    }
  } else {
    const uint32_t B = RS_FRAC(-wrk->y_accum, wrk->y_sub);
    const uint32_t A = (uint32_t)(RS_ONE - B);
    for (int x_out = 0; x_out < x_out_max; ++x_out) {
      const uint64_t I = (uint64_t)A * wrk->frow[x_out] + (uint64_t)B * wrk->irow[x_out];
      const uint32_t J = (uint32_t)((I + RS_ROUNDER) >> RS_RFIX);
      const int v = (int)RS_MULT_FIX(J, wrk->fy_scale);
      dst[x_out] = (v > 255) ? 255u : (uint8_t)v;
    }
  }
}
/* rescaler.c.go:138-163 */
static void ExportRowShrink(Rescaler* wrk) {
  uint8_t* dst = wrk->dst;
  const int x_out_max = wrk->dst_width * wrk->num_channels;
  const uint32_t yscale = wrk->fy_scale * (-wrk->y_accum);
  if (yscale) {
    for (int x_out = 0; x_out < x_out_max; ++x_out) {
      const uint32_t frac = (uint32_t)RS_MULT_FIX_FLOOR(wrk->frow[x_out], yscale);
      const int v = (int)RS_MULT_FIX(wrk->irow[x_out] - frac, wrk->fxy_scale);
      dst[x_out] = (v > 255) ? 255u : (uint8_t)v;
      wrk->irow[x_out] = frac;
    }
  } else {
    for (int x_out = 0; x_out < x_out_max; ++x_out) {
      const int v = (int)RS_MULT_FIX(wrk->irow[x_out], wrk->fxy_scale);
      dst[x_out] = (v > 255) ? 255u : (uint8_t)v;
      wrk->irow[x_out] = 0;
    }
  }
}
/* rescaler.c.go:178-197 + rescaler_utils.h (HasPendingOutput = !OutputDone && y_accum <= 0) */
static int HasPendingOutput(const Rescaler* r) { return r->dst_y < r->dst_height && r->y_accum <= 0; }
static void ExportRow(Rescaler* wrk) {
  if (wrk->y_accum <= 0) {
    if (wrk->y_expand) {
      ExportRowExpand(wrk);
    } else if (wrk->fxy_scale) {
      ExportRowShrink(wrk);
    } else { /* special case */
      for (int i = 0; i < wrk->num_channels * wrk->dst_width; ++i) {
        wrk->dst[i] = (uint8_t)wrk->irow[i];
        wrk->irow[i] = 0;
      }
    }
    wrk->y_accum += wrk->y_add;
    wrk->dst += wrk->dst_stride;
    ++wrk->dst_y;
  }
}
/* RescalePlane (pkg/libwebp/enc/picture_rescale_enc.c.go:125-136) over WebPRescalerImport /
 * WebPRescalerExport (rescaler_utils.c.go:119-158). */
int orc_rescale_plane(const uint8_t* src, int src_width, int src_height, int src_stride, uint8_t* dst, int dst_width,
                      int dst_height, int dst_stride, int num_channels) {
  Rescaler r;
  rescaler_t* work = (rescaler_t*)malloc(sizeof(rescaler_t) * 2 * (size_t)dst_width * num_channels);
  int y = 0;
  if (work == NULL) return 0;
  RescalerInit(&r, src_width, src_height, dst, dst_width, dst_height, dst_stride, num_channels, work);
  while (y < src_height) {
    int num_lines = src_height - y, total_imported = 0;
    const uint8_t* s = src + (ptrdiff_t)y * src_stride;
    while (total_imported < num_lines && !HasPendingOutput(&r)) {
      if (r.y_expand) {
        rescaler_t* const tmp = r.irow;
        r.irow = r.frow;
        r.frow = tmp;
      }
      if (!r.x_expand) ImportRowShrink(&r, s); else ImportRowExpand(&r, s);
      if (!r.y_expand) {
        for (int x = 0; x < r.num_channels * r.dst_width; ++x) r.irow[x] += r.frow[x];
      }
      ++r.src_y;
      s += src_stride;
      ++total_imported;
      r.y_accum -= r.y_sub;
    }
    y += total_imported;
    while (HasPendingOutput(&r)) ExportRow(&r);
  }
  free(work);
  return 1;
}

/* ---- alpha level smoothing --------------------------------------------- */
/* WebPDequantizeLevels, pkg/libwebp/utils/quant_levels_dec_utils.c.go:30-286: a sliding
 * (2r+1)^2 box average (vertical sums with replicated rows, horizontal sums on cumulative
 * columns with mirroring), a correction look-up that pulls each level towards the local
 * average by at most the smallest gap between two used levels, ordered dithering. */
#define DQ_FIX 16
#define DQ_LFIX 2
#define DQ_LUT_SIZE ((1 << (8 + DQ_LFIX)) - 1)
#define DQ_DSIZE 4
/* The Go file keeps the ordered-dither branch of the C source (DFIX = 4, :36-45); the stock
 * libwebp build leaves USE_DITHERING undefined (DFIX = 0, no dither).  `dfix` selects: 4 is
 * the reference as written, 0 is what the libwebp 1.6.0 binary runs (the pinned variant);
 * the two differ in this constant and the dither add only.  PINNED (dfix = 0) against
 * WebPDecode of files with quantised alpha levels: tests/golden/make_golden_dequant.py. */
static const uint8_t kOrderedDither[DQ_DSIZE][DQ_DSIZE] = {{0, 8, 2, 10}, {12, 4, 14, 6}, {3, 11, 1, 9}, {15, 7, 13, 5}};
typedef struct {
  int width, height, stride, row;
  uint8_t* src;
  uint8_t* dst;
  int radius, scale;
  uint16_t *start, *cur, *end, *top, *average;
  int num_levels, min, max, min_level_dist;
  int16_t* correction;
  int dfix;
} SmoothParams;
static uint8_t dq_clip_8b(int v, int dfix) {
  const int mask = (int)(~0u << (8 + dfix));
  return (!(v & mask)) ? (uint8_t)(v >> dfix) : (v < 0) ? 0u : 255u;
}
/* quant_levels_dec_utils.c.go:86-111 */
static void DqVFilter(SmoothParams* p) {
  const uint8_t* src = p->src;
  const int w = p->width;
  uint16_t* const cur = p->cur;
  const uint16_t* const top = p->top;
  uint16_t* const out = p->end;
  uint16_t sum = 0; /* all arithmetic is modulo 16bit */
  for (int x = 0; x < w; ++x) {
    uint16_t new_value;
    sum += src[x];
    new_value = top[x] + sum;
    out[x] = new_value - cur[x]; /* vertical sum of 'r' pixels. */
    cur[x] = new_value;
  }
  p->top = p->cur;
  p->cur += w;
  if (p->cur == p->end) p->cur = p->start; /* roll-over */
  if (p->row >= 0 && p->row < p->height - 1) p->src += p->stride;
}
/* :115-136 */
static void DqHFilter(SmoothParams* p) {
  const uint16_t* const in = p->end;
  uint16_t* const out = p->average;
  const uint32_t scale = p->scale;
  const int w = p->width, r = p->radius;
  int x;
  for (x = 0; x <= r; ++x) { /* left mirroring */
    const uint16_t delta = in[x + r - 1] + in[r - x];
    out[x] = (delta * scale) >> DQ_FIX;
  }
  for (; x < w - r; ++x) { /* bulk middle run */
    const uint16_t delta = in[x + r] - in[x - r - 1];
    out[x] = (delta * scale) >> DQ_FIX;
  }
  for (; x < w; ++x) { /* right mirroring */
    const uint16_t delta = 2 * in[w - 1] - in[2 * w - 2 - r - x] - in[x - r - 1];
    out[x] = (delta * scale) >> DQ_FIX;
  }
}
/* :139-157 */
static void DqApplyFilter(SmoothParams* p) {
  const uint16_t* const average = p->average;
  const int w = p->width;
  const int16_t* const correction = p->correction + DQ_LUT_SIZE;
  const uint8_t* const dither = kOrderedDither[p->row % DQ_DSIZE];
  uint8_t* const dst = p->dst;
  for (int x = 0; x < w; ++x) {
    const int v = dst[x];
    if (v < p->max && v > p->min) {
      const int c = (v << p->dfix) + correction[average[x] - (v << DQ_LFIX)];
      dst[x] = dq_clip_8b(c + (p->dfix ? dither[x % DQ_DSIZE] : 0), p->dfix);
    }
  }
  p->dst += p->stride;
}
/* :162-184 */
static void DqInitCorrectionLUT(int16_t* lut_ptr, int min_dist, int dfix) {
  const int threshold1 = min_dist << DQ_LFIX;
  const int threshold2 = (3 * threshold1) >> 2;
  const int max_threshold = threshold2 << dfix;
  const int delta = threshold1 - threshold2;
  int16_t* const lut = lut_ptr + DQ_LUT_SIZE;
  for (int i = 1; i <= DQ_LUT_SIZE; ++i) {
    int c = (i <= threshold2) ? (i << dfix) : (i < threshold1) ? max_threshold * (threshold1 - i) / delta : 0;
    c >>= DQ_LFIX;
    lut[+i] = (int16_t)+c;
    lut[-i] = (int16_t)-c;
  }
  lut[0] = 0;
}
/* :186-219 */
static void DqCountLevels(SmoothParams* p) {
  uint8_t used_levels[256] = {0};
  const uint8_t* data = p->src;
  p->min = 255;
  p->max = 0;
  for (int j = 0; j < p->height; ++j) {
    for (int i = 0; i < p->width; ++i) {
      const int v = data[i];
      if (v < p->min) p->min = v;
      if (v > p->max) p->max = v;
      used_levels[v] = 1;
    }
    data += p->stride;
  }
  p->min_level_dist = p->max - p->min;
  int last_level = -1;
  for (int i = 0; i < 256; ++i) {
    if (used_levels[i]) {
      ++p->num_levels;
      if (last_level >= 0) {
        const int level_dist = i - last_level;
        if (level_dist < p->min_level_dist) p->min_level_dist = level_dist;
      }
      last_level = i;
    }
  }
}
/* :222-286 (InitParams + WebPDequantizeLevels), in place on data */
int orc_dequantize_levels(uint8_t* data, int width, int height, int stride, int strength, int ordered_dither) {
  int radius = 4 * strength / 100;
  if (strength < 0 || strength > 100) return 0;
  if (data == NULL || width <= 0 || height <= 0) return 0;
  if (2 * radius + 1 > width) radius = (width - 1) >> 1;
  if (2 * radius + 1 > height) radius = (height - 1) >> 1;
  if (radius > 0) {
    SmoothParams p;
    const int R = 2 * radius + 1;
    memset(&p, 0, sizeof(p));
    uint16_t* scratch = (uint16_t*)calloc((size_t)(R + 1) * width + width, sizeof(uint16_t));
    int16_t* lut = (int16_t*)calloc(1 + 2 * DQ_LUT_SIZE, sizeof(int16_t));
    if (scratch == NULL || lut == NULL) { free(scratch); free(lut); return 0; }
    p.start = scratch;
    p.cur = p.start;
    p.end = p.start + R * width;
    p.top = p.end - width;
    memset(p.top, 0, width * sizeof(*p.top));
    p.width = width;
    p.average = scratch + (size_t)(R + 1) * width;
    p.height = height;
    p.stride = stride;
    p.src = data;
    p.dst = data;
    p.radius = radius;
    p.scale = (1 << (DQ_FIX + DQ_LFIX)) / (R * R);
    p.row = -radius;
    p.dfix = ordered_dither ? 4 : 0;
    DqCountLevels(&p);
    p.correction = lut;
    DqInitCorrectionLUT(p.correction, p.min_level_dist, p.dfix);
    if (p.num_levels > 2) {
      for (; p.row < p.height; ++p.row) {
        DqVFilter(&p);
        if (p.row >= p.radius) {
          DqHFilter(&p);
          DqApplyFilter(&p);
        }
      }
    }
    free(scratch);
    free(lut);
  }
  return 1;
}
