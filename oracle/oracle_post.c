/*
 * oracle_post.c -- CPU restatement of the output conversions that follow parts 2
 * and 3 in the decoder.  TEST INFRASTRUCTURE ONLY, same rules as
 * oracle.c: only tests/, smoke() and bench.py's CPU legs may use it.
 *
 * Parity status: unpinned by any reference test or binary vector (libwebp exports none
 * of these row functions); checked by identities in tests/test_oracle_properties.py.
 *
 * Every function cites the reference file:line (relative to the go-webp tree).
 * 16-bit modes follow the Go text literally: VP8LConvertBGRAToRGBA4444/RGB565 write
 * the ba / gb byte first and ApplyAlphaMultiply_16b uses rg_byte_pos 1, i.e. upstream
 * built with WEBP_SWAP_16BIT_CSP; swap16 == 0 selects that,
 * swap16 == 1 the upstream default order.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

int orc_YUVToR(int y, int v);
int orc_YUVToG(int y, int u, int v);
int orc_YUVToB(int y, int u);
int orc_mode_bpp(int mode);

enum {
  MODE_RGB = 0, MODE_RGBA = 1, MODE_BGR = 2, MODE_BGRA = 3, MODE_ARGB = 4,
  MODE_RGBA_4444 = 5, MODE_RGB_565 = 6, MODE_rgbA = 7, MODE_bgrA = 8,
  MODE_Argb = 9, MODE_rgbA_4444 = 10
};

/* ---- pkg/libwebp/dsp/alpha_processing.go:100-163 (premultiplied modes) ---- */
static inline uint32_t MULTIPLIER1(uint32_t a) { return a * 32897u; }
static inline uint8_t PREMULTIPLY(uint32_t x, uint32_t m) { return (uint8_t)((x * m) >> 23); }
/* ApplyAlphaMultiply_C, alpha_processing.go:110-126 */
void orc_apply_alpha_multiply(uint8_t* rgba, int alpha_first, int w, int h, int stride) {
  while (h-- > 0) {
    uint8_t* const rgb = rgba + (alpha_first ? 1 : 0);
    const uint8_t* const alpha = rgba + (alpha_first ? 0 : 3);
    int i;
    for (i = 0; i < w; ++i) {
      const uint32_t a = alpha[4 * i];
      if (a != 0xff) {
        const uint32_t mult = MULTIPLIER1(a);
        rgb[4 * i + 0] = PREMULTIPLY(rgb[4 * i + 0], mult);
        rgb[4 * i + 1] = PREMULTIPLY(rgb[4 * i + 1], mult);
        rgb[4 * i + 2] = PREMULTIPLY(rgb[4 * i + 2], mult);
      }
    }
    rgba += stride;
  }
}
/* alpha_processing.go:128-163 */
static inline uint8_t dither_hi(uint8_t x) { return (uint8_t)((x & 0xf0) | (x >> 4)); }
static inline uint8_t dither_lo(uint8_t x) { return (uint8_t)((x & 0x0f) | (x << 4)); }
static inline uint8_t multiply4(uint8_t x, uint32_t m) { return (uint8_t)((x * m) >> 16); }
void orc_apply_alpha_multiply_4444(uint8_t* rgba4444, int w, int h, int stride, int rg_byte_pos) {
  while (h-- > 0) {
    int i;
    for (i = 0; i < w; ++i) {
      const uint8_t rg = rgba4444[2 * i + rg_byte_pos];
      const uint8_t ba = rgba4444[2 * i + (rg_byte_pos ^ 1)];
      const uint8_t a = ba & 0x0f;
      const uint32_t mult = (uint32_t)a * 0x1111u;
      const uint8_t r = multiply4(dither_hi(rg), mult);
      const uint8_t g = multiply4(dither_lo(rg), mult);
      const uint8_t b = multiply4(dither_hi(ba), mult);
      rgba4444[2 * i + rg_byte_pos] = (uint8_t)((r & 0xf0) | ((g >> 4) & 0x0f));
      rgba4444[2 * i + (rg_byte_pos ^ 1)] = (uint8_t)((b & 0xf0) | a);
    }
    rgba4444 += stride;
  }
}

/* ---- pkg/libwebp/dsp/alpha_processing.go:17-89 (WebPMultARGBRow / WebPMultRow) ---- */
#define MFIX 24
#define HALF ((1u << MFIX) >> 1)
#define KINV_255 ((1u << MFIX) / 255u)
static inline uint32_t Mult(uint8_t x, uint32_t mult) { return (x * mult + HALF) >> MFIX; }
static inline uint32_t GetScale(uint32_t a, int inverse) { return inverse ? (255u << MFIX) / a : a * KINV_255; }
void orc_mult_argb_row(uint32_t* ptr, int width, int inverse) {
  int x;
  for (x = 0; x < width; ++x) {
    const uint32_t argb = ptr[x];
    if (argb < 0xff000000u) {
      if (argb <= 0x00ffffffu) {
        ptr[x] = 0;
      } else {
        const uint32_t alpha = (argb >> 24) & 0xff;
        const uint32_t scale = GetScale(alpha, inverse);
        uint32_t out = argb & 0xff000000u;
        out |= Mult((uint8_t)(argb >> 0), scale) << 0;
        out |= Mult((uint8_t)(argb >> 8), scale) << 8;
        out |= Mult((uint8_t)(argb >> 16), scale) << 16;
        ptr[x] = out;
      }
    }
  }
}
void orc_mult_row(uint8_t* ptr, const uint8_t* alpha, int width, int inverse) {
  int x;
  for (x = 0; x < width; ++x) {
    const uint32_t a = alpha[x];
    if (a != 255) {
      if (a == 0) ptr[x] = 0;
      else ptr[x] = (uint8_t)Mult(ptr[x], GetScale(a, inverse));
    }
  }
}

/* ---- pkg/libwebp/dsp/lossless.go:549-666 (VP8LConvertFromBGRA and its rows) ---- */
void orc_convert_from_bgra(const uint32_t* src, int num_pixels, int mode, int swap16, uint8_t* dst) {
  int i;
  switch (mode) {
    case MODE_RGB:
      for (i = 0; i < num_pixels; ++i) { const uint32_t p = src[i]; *dst++ = p >> 16; *dst++ = p >> 8; *dst++ = p; }
      break;
    case MODE_BGR:
      for (i = 0; i < num_pixels; ++i) { const uint32_t p = src[i]; *dst++ = p; *dst++ = p >> 8; *dst++ = p >> 16; }
      break;
    case MODE_RGBA: case MODE_rgbA:
      for (i = 0; i < num_pixels; ++i) { const uint32_t p = src[i]; dst[4 * i] = p >> 16; dst[4 * i + 1] = p >> 8; dst[4 * i + 2] = p; dst[4 * i + 3] = p >> 24; }
      if (mode == MODE_rgbA) orc_apply_alpha_multiply(dst, 0, num_pixels, 1, 0);
      break;
    case MODE_BGRA: case MODE_bgrA:  /* CopyOrSwap(..., 1): little-endian memcpy */
      memcpy(dst, src, (size_t)num_pixels * 4);
      if (mode == MODE_bgrA) orc_apply_alpha_multiply(dst, 0, num_pixels, 1, 0);
      break;
    case MODE_ARGB: case MODE_Argb:  /* CopyOrSwap(..., 0): byte-swapped words */
      for (i = 0; i < num_pixels; ++i) { const uint32_t p = src[i]; dst[4 * i] = p >> 24; dst[4 * i + 1] = p >> 16; dst[4 * i + 2] = p >> 8; dst[4 * i + 3] = p; }
      if (mode == MODE_Argb) orc_apply_alpha_multiply(dst, 1, num_pixels, 1, 0);
      break;
    case MODE_RGBA_4444: case MODE_rgbA_4444:
      for (i = 0; i < num_pixels; ++i) {
        const uint32_t p = src[i];
        const uint8_t rg = (uint8_t)(((p >> 16) & 0xf0) | ((p >> 12) & 0xf));
        const uint8_t ba = (uint8_t)(((p >> 0) & 0xf0) | ((p >> 28) & 0xf));
        dst[2 * i + (swap16 ? 0 : 1)] = rg;   /* Go text: ba first (lossless.go:581-584) */
        dst[2 * i + (swap16 ? 1 : 0)] = ba;
      }
      if (mode == MODE_rgbA_4444) orc_apply_alpha_multiply_4444(dst, num_pixels, 1, 0, swap16 ? 0 : 1);
      break;
    case MODE_RGB_565:
      for (i = 0; i < num_pixels; ++i) {
        const uint32_t p = src[i];
        const uint8_t rg = (uint8_t)(((p >> 16) & 0xf8) | ((p >> 13) & 0x7));
        const uint8_t gb = (uint8_t)(((p >> 5) & 0xe0) | ((p >> 3) & 0x1f));
        dst[2 * i + (swap16 ? 0 : 1)] = rg;   /* Go text: gb first (lossless.go:593-596) */
        dst[2 * i + (swap16 ? 1 : 0)] = gb;
      }
      break;
    default: break;
  }
}

/* ---- point sampling and 4:4:4 (pkg/libwebp/dsp/yuv.go:19-56, upsampling.c.go:191-245) ---- */
/* per-pixel writers: upstream yuv.h semantics as in oracle.c (the Go tree references but never defines them) */
static void yuv_pixel(int mode, int y, int u, int v, uint8_t* dst) {
  const int r = orc_YUVToR(y, v), g = orc_YUVToG(y, u, v), b = orc_YUVToB(y, u);
  switch (mode) {
    case MODE_RGB: dst[0] = r; dst[1] = g; dst[2] = b; break;
    case MODE_BGR: dst[0] = b; dst[1] = g; dst[2] = r; break;
    case MODE_RGBA: case MODE_rgbA: dst[0] = r; dst[1] = g; dst[2] = b; dst[3] = 0xff; break;
    case MODE_BGRA: case MODE_bgrA: dst[0] = b; dst[1] = g; dst[2] = r; dst[3] = 0xff; break;
    case MODE_ARGB: case MODE_Argb: dst[0] = 0xff; dst[1] = r; dst[2] = g; dst[3] = b; break;
    case MODE_RGBA_4444: case MODE_rgbA_4444: dst[0] = (r & 0xf0) | (g >> 4); dst[1] = (b & 0xf0) | 0x0f; break;
    case MODE_RGB_565: dst[0] = (r & 0xf8) | (g >> 5); dst[1] = ((g << 3) & 0xe0) | (b >> 3); break;
    default: break;
  }
}
/* row_func (yuv.go:19-36) driven by WebPSamplerProcessPlane (yuv.go:47-58) == EmitSampledRGB (io_dec.c.go:53-59) */
void orc_sampler_process_plane(int mode, const uint8_t* y, int y_stride, const uint8_t* u, const uint8_t* v,
                               int uv_stride, uint8_t* dst, int dst_stride, int width, int height) {
  const int bpp = orc_mode_bpp(mode);
  int j, x;
  for (j = 0; j < height; ++j) {
    for (x = 0; x < width; ++x) yuv_pixel(mode, y[x], u[x >> 1], v[x >> 1], dst + x * bpp);
    y += y_stride;
    if (j & 1) { u += uv_stride; v += uv_stride; }
    dst += dst_stride;
  }
}
/* YUV444_FUNC (upsampling.c.go:191-203) */
void orc_yuv444_row(int mode, const uint8_t* y, const uint8_t* u, const uint8_t* v, uint8_t* dst, int len) {
  const int bpp = orc_mode_bpp(mode);
  int i;
  for (i = 0; i < len; ++i) yuv_pixel(mode, y[i], u[i], v[i], dst + i * bpp);
}
