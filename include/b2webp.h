/*
 * b2webp.h -- C ABI of the B200 (sm_100a) pixel hot path for go-webp.
 *
 * This is the drop-in boundary: the only thing go-webp's Go packages bind
 * (through the cgo shim go/pkg/cuda).  Plain pointers and
 * sizes, no C++/torch types.  The library owns every device and pinned-host
 * allocation; no caller pointer is retained after a call returns (cgo rule).
 *
 * All `file:line` citations are relative to the go-webp tree.
 *
 * Threading: a b2_ctx is bound to one device and is not thread-safe; use one
 * context per host thread / per GPU.  All work is enqueued on the context's
 * streams and is asynchronous until b2_sync().
 *
 * Status codes map onto go-webp's sentinels (pkg/picture/errors.go:8-21,
 * pkg/vp8/enums.go:20-31): see b2_status below.
 */
#ifndef B2WEBP_H_
#define B2WEBP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_ABI_VERSION 0x0105

typedef enum b2_status {
  B2_OK = 0,
  B2_ERR_OOM = -1,            /* ENC_ERROR_OUT_OF_MEMORY / VP8_STATUS_OUT_OF_MEMORY */
  B2_ERR_BAD_DIM = -2,        /* ENC_ERROR_BAD_DIMENSION   */
  B2_ERR_NULL = -3,           /* ENC_ERROR_nil_PARAMETER   */
  B2_ERR_INVALID_CONFIG = -4, /* ENC_ERROR_INVALID_CONFIGURATION / VP8_STATUS_INVALID_PARAM */
  B2_ERR_CUDA = -5            /* device/runtime failure; b2_last_error() has the text */
} b2_status;

/* WEBP_MAX_DIMENSION, pkg/libwebp/webp/encode.h.go:95 */
#define B2_MAX_DIMENSION 16383

typedef struct b2_ctx b2_ctx;

/* ---- library / context ------------------------------------------------- */
int b2_abi_version(void);
const char* b2_strerror(int status);
int b2_device_count(int* count);
/* One context per GPU.  n_streams in [1,16].  A context owns, per stream, a descriptor ring, a grow-only
 * scratch area and a few work-counter words (the import and the fancy upsampler hand their row pairs out
 * by ticket; the last warp of a launch clears the counters it used).  Consequence for callers: launches on
 * ONE stream of ONE context must come from one host thread at a time (they are stream-ordered on the
 * device); different streams and different contexts are independent.
 * b2_device_count asks the driver again for up to 3 s when a box has only just come up. */
int b2_ctx_create(int device, int n_streams, b2_ctx** out);
void b2_ctx_destroy(b2_ctx* ctx);
/* Byte order of the 16-bit output modes (MODE_RGBA_4444, MODE_rgbA_4444, MODE_RGB_565), ONE setting for
 * every producer and consumer of such pixels (fancy upsampler, point samplers, YUV444 converters,
 * VP8LConvertFromBGRA, the alpha nibble of EmitAlphaRGBA4444, ApplyAlphaMultiply4444's default
 * rg_byte_pos in the host mirror).  swap = 1 (default): the order the go-webp text has everywhere --
 * libwebp built with WEBP_SWAP_16BIT_CSP, the ba / gb byte first (pkg/libwebp/dsp/lossless.go:575-600,
 * pkg/libwebp/decoder/io_dec.c.go:197-226 writes the alpha nibble into byte 2i,
 * ApplyAlphaMultiply_16b_C is bound with rg_byte_pos = 1).  swap = 0: the stock libwebp build (rg byte
 * first), which is what the libwebp 1.6.0 binary behind the golden vectors produces. */
int b2_ctx_set_swap_16bit_csp(b2_ctx* ctx, int swap);
int b2_ctx_swap_16bit_csp(const b2_ctx* ctx);
int b2_ctx_device(const b2_ctx* ctx);
int b2_ctx_sm_count(const b2_ctx* ctx);
const char* b2_last_error(const b2_ctx* ctx);
/* Kernels launched by this context so far (bench.py's gpu_launches claim). */
uint64_t b2_launch_count(const b2_ctx* ctx);

/* ---- memory, copies, sync, timing --------------------------------------- */
int b2_pinned_alloc(b2_ctx* ctx, size_t bytes, void** host_ptr);
int b2_pinned_free(b2_ctx* ctx, void* host_ptr);
/* Page-lock caller memory for async copies (Go []byte must be pinned by the
 * runtime.Pinner for the duration; the shim does that). */
int b2_host_register(b2_ctx* ctx, void* host_ptr, size_t bytes);
int b2_host_unregister(b2_ctx* ctx, void* host_ptr);
int b2_dev_alloc(b2_ctx* ctx, size_t bytes, void** dev_ptr);
int b2_dev_free(b2_ctx* ctx, void* dev_ptr);
int b2_h2d_async(b2_ctx* ctx, int stream, void* dst_dev, const void* src_host, size_t bytes);
int b2_d2h_async(b2_ctx* ctx, int stream, void* dst_host, const void* src_dev, size_t bytes);
/* Pitched copies: the reference's planes have arbitrary (signed) strides;
 * pitches here are the positive row distances of both sides. */
int b2_h2d_2d_async(b2_ctx* ctx, int stream, void* dst_dev, size_t dst_pitch,
                    const void* src_host, size_t src_pitch, size_t width_bytes, size_t rows);
int b2_d2h_2d_async(b2_ctx* ctx, int stream, void* dst_host, size_t dst_pitch,
                    const void* src_dev, size_t src_pitch, size_t width_bytes, size_t rows);
int b2_memset_async(b2_ctx* ctx, int stream, void* dst_dev, int value, size_t bytes);
/* stream < 0 waits for every stream of the context. */
int b2_sync(b2_ctx* ctx, int stream);
/* CUDA-event stopwatch on a context stream (slots 0..31). */
int b2_event_record(b2_ctx* ctx, int stream, int slot);
int b2_event_elapsed_ms(b2_ctx* ctx, int slot_start, int slot_stop, float* ms);
/* Blocks the host until the event in 'slot' has happened. */
int b2_event_sync(b2_ctx* ctx, int slot);
/* Copies issued through this context so far: out = {H2D calls, H2D bytes, D2H calls,
 * D2H bytes}.  What a caller uses to check that chained calls on device-resident
 * pictures do not touch the host in between. */
int b2_copy_stats(const b2_ctx* ctx, uint64_t out[4]);

/* ---- part 1: RGBA -> YUV420(+A) import ----------------------------------
 * Replaces the row loop of dsp.WebPImportYUVAFromRGBA /
 * WebPImportYUVAFromRGBALastLine (pkg/libwebp/dsp/yuv.go:304-382) and the
 * has-alpha scan + driver of picture.ImportYUVAFromRGBA
 * (pkg/picture/conversions.go:50-121) for a whole batch of images.
 * All pointers are DEVICE pointers.  src addresses pixel (0,0); strides are in
 * bytes and may be negative (flipped buffers, buffer_dec.c.go:161-184).
 */
typedef struct b2_import_job {
  const uint8_t* src;
  uint8_t* y;
  uint8_t* u;
  uint8_t* v;
  uint8_t* a;         /* may be NULL: no alpha plane is written */
  int32_t src_stride; /* |src_stride| >= bpp*width (picture_csp_enc.c.go:132) */
  int32_t y_stride, uv_stride, a_stride;
  int32_t width, height;
} b2_import_job;

/* Byte order of one source pixel (Import(): picture_csp_enc.c.go:123-137;
 * ARGB-as-uint32 little-endian of PictureARGBToYUVA, conversions.go:22-26, is
 * B,G,R,A in memory = B2_ORDER_BGRA). */
typedef enum b2_pixel_order {
  B2_ORDER_RGBA = 0, /* step 4: WebPPictureImportRGBA / RGBX */
  B2_ORDER_BGRA = 1, /* step 4: WebPPictureImportBGRA / BGRX / ARGB32 */
  B2_ORDER_RGB = 2,  /* step 3: WebPPictureImportRGB */
  B2_ORDER_BGR = 3   /* step 3: WebPPictureImportBGR */
} b2_pixel_order;

#define B2_IMPORT_ALPHA 0x1     /* import_alpha=1: use byte 3 as alpha (step 4 only) */
#define B2_IMPORT_DIV_TABLE 0x2 /* libwebp-binary kInvAlpha divide instead of the
                                   reference's exact 4*sum/a (yuv.go:231) */

/* has_alpha_dev: DEVICE int32[n_jobs]; entry i becomes 1 iff image i has any
 * imported alpha != 0xff (conversions.go:56-58 -> WEBP_YUV420A), else 0.  May
 * be NULL.  jobs is a HOST array (copied before return). */
int b2_rgba_to_yuv420a(b2_ctx* ctx, int stream, const b2_import_job* jobs, int n_jobs,
                       int order, int flags, int32_t* has_alpha_dev);

/* The dithered copy of the import: pkg/picture/conversions.go:122-165
 * with ConvertRowToY / ConvertRowsToUV (pkg/libwebp/enc/picture_csp_enc.c.go:31-48) and
 * VP8Random (pkg/vp8/random.go:13-95) -- the rounding constants of RGBToY / RGBToU /
 * RGBToV become draws of ONE generator that runs through the whole image (two luma rows,
 * then u, v per chroma sample, per row pair).  dithering in (0, 1] as
 * PictureARGBToYUVA's argument (amp = dithering * 256; 0 or an amp of 0 = the plain
 * import).  The generator is linear, so every row pair gets its state by skip-ahead and
 * the image is still processed in parallel.  Preconditions of the dithered path: 4-byte
 * pixels, source rows and Y / A planes 4-byte aligned, U / V planes 2-byte aligned
 * (B2_ERR_INVALID_CONFIG otherwise); PictureARGBToYUVA's ARGB pictures always meet them. */
int b2_rgba_to_yuv420a_dithered(b2_ctx* ctx, int stream, const b2_import_job* jobs, int n_jobs,
                                int order, int flags, int32_t* has_alpha_dev, float dithering);

/* ---- part 2: YUV420(+A) -> RGBA fancy upsampling -------------------------
 * Replaces WebPUpsamplers[mode] line pairs (pkg/libwebp/dsp/upsampling.c.go:
 * 43-114,275-289) driven by the whole-image schedule of WebPPictureYUVAToARGB
 * (pkg/libwebp/enc/picture_csp_enc.c.go:62-118) == EmitFancyRGB's streaming
 * schedule (pkg/libwebp/decoder/io_dec.c.go:65-115), with the alpha dispatch
 * of EmitAlphaRGB (io_dec.c.go:175-195) fused into the store.
 * mode: WEBP_CSP_MODE numbering (pkg/libwebp/webp/enums.go:45-64); premultiplied
 * modes alias the plain ones exactly as upsampling.c.go:280-288 does.
 */
typedef enum b2_csp_mode {
  B2_MODE_RGB = 0, B2_MODE_RGBA = 1, B2_MODE_BGR = 2, B2_MODE_BGRA = 3,
  B2_MODE_ARGB = 4, B2_MODE_RGBA_4444 = 5, B2_MODE_RGB_565 = 6,
  B2_MODE_rgbA = 7, B2_MODE_bgrA = 8, B2_MODE_Argb = 9, B2_MODE_rgbA_4444 = 10
} b2_csp_mode;

typedef struct b2_upsample_job {
  const uint8_t* y;
  const uint8_t* u;
  const uint8_t* v;
  const uint8_t* a; /* may be NULL: alpha byte = 0xff */
  uint8_t* dst;
  int32_t y_stride, uv_stride, a_stride, dst_stride; /* bytes, signed */
  int32_t width, height;
  /* Output rows to produce, [row_begin, row_end); both 0 = the whole image.  This is
   * EmitFancyRGB's streaming schedule (pkg/libwebp/decoder/io_dec.c.go:65-115): for a
   * macroblock-row batch [mb_y, mb_y + mb_h) it emits rows max(mb_y - 1, 0) ..
   * mb_y + mb_h - 2 (the last row of a batch waits for the chroma row below it) and, on
   * the last batch, the final row as well.  Pointers always address row 0 of the
   * image; rows outside the range are neither read for output nor written. */
  int32_t row_begin, row_end;
} b2_upsample_job;

int b2_yuv420a_to_rgba_fancy(b2_ctx* ctx, int stream, const b2_upsample_job* jobs,
                             int n_jobs, int mode);

/* ---- part 3: VP8L per-pixel transforms -----------------------------------
 * Pixels are uint32 0xAARRGGBB, images tightly packed (row stride == width,
 * as dec.pixels / argb_cache are: pkg/vp8/vp8l_dec.c.go:710-760).
 */
typedef struct b2_argb_job {
  const uint32_t* src;
  uint32_t* dst;        /* may equal src for the same-size transforms */
  const uint32_t* data; /* tile codes (cross-colour), palette (colour index), or NULL */
  int32_t width, height;
  int32_t bits;         /* VP8LTransform.bits (pkg/vp8/models.go:126-132) */
  int32_t reserved;
} b2_argb_job;

/* VP8LAddGreenToBlueAndRed_C, pkg/libwebp/dsp/lossless.go:337-347 */
int b2_vp8l_add_green(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs);
/* VP8LSubtractGreenFromBlueAndRed_C, pkg/libwebp/dsp/lossless_enc.go:207-216 */
int b2_vp8l_sub_green(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs);
/* ColorSpaceInverseTransform_C + VP8LTransformColorInverse_C, lossless.go:359-413;
 * data = ceil(w/2^bits) x ceil(h/2^bits) colour codes, bits in [2,9]. */
int b2_vp8l_color_inverse(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs);
/* VP8LTransformColor_C over given tile codes, lossless_enc.go:224-239 +
 * pkg/libwebp/enc/predictor_enc.go:879-889. */
int b2_vp8l_color_forward(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs);
/* Fused cross-colour^-1 then add-green: the ApplyInverseTransforms order
 * (pkg/vp8/vp8l_dec.c.go:710-729) for a chain without predictor, in one pass. */
int b2_vp8l_inverse_chain(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs);
/* Forward counterpart: sub-green then cross-colour forward in one pass
 * (pkg/vp8/vp8l_encoder.go:926-961). */
int b2_vp8l_forward_chain(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs);
/* ColorIndexInverseTransform_C / MapARGB_C, lossless.go:415-459.  src rows are
 * ceil(width/2^bits) pixels; data = palette of 1<<(8>>bits) entries (already
 * expanded, vp8l_dec.c.go:1196-1219); bits in [0,3].  dst must not alias src. */
int b2_vp8l_color_index_inverse(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs);
/* VP8LBundleColorMap_C, lossless_enc.go:293-312: src = uint8 indices (one per
 * pixel, row stride == width, passed through the src pointer), dst rows are
 * ceil(width/2^bits) uint32. */
int b2_vp8l_bundle_color_map(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs);

/* ApplyPalette (pkg/vp8/vp8l_encoder.go:1093-1212; caller MapImageFromPalette :1220-1244): every
 * ARGB pixel of src becomes its index in 'palette' -- greedy compares below 4 colours, else the first
 * of the three ApplyPaletteHash functions that is collision-free on the palette, else an exact search
 * (SearchColorNoIdx, pkg/libwebp/utils/palette.go:74-119) -- and the indices are bundled exactly as
 * VP8LBundleColorMap does (lossless_enc.go:293-312): dst rows hold ceil(width / 2^xbits) words
 * 0xff000000 | indices << 8.  The reference goes through a u8 row buffer; here lookup and bundle are
 * one pass.  palette: DEVICE array of palette_size (1..256) colours; strides in pixels / words; xbits
 * in [0,3] (MapImageFromPalette: <= 2 colours 3, <= 4: 2, <= 16: 1, else 0).  dst may equal src only
 * with xbits == 0 and equal strides.  Pixels that are not palette colours give what the reference
 * gives where it is defined (greedy, hash tables); its binary search does not terminate for them. */
typedef struct b2_palette_job {
  const uint32_t* src;
  uint32_t* dst;
  const uint32_t* palette;
  int32_t src_stride, dst_stride;
  int32_t width, height;
  int32_t palette_size;
  int32_t xbits;
} b2_palette_job;
int b2_vp8l_apply_palette(b2_ctx* ctx, int stream, const b2_palette_job* jobs, int n_jobs);

/* ---- next row of the table: VP8L predictor transform -----------
 * data = mode image, ceil(w/2^bits) x ceil(h/2^bits) words, mode in bits 8..11
 * (lossless.go:313); bits in [2,9]; dst must not alias src.
 * Inverse: PredictorInverseTransform_C + PredictorAdd0..13_C over whole images
 * (pkg/libwebp/dsp/lossless.go:31-333): src = residuals, dst = pixels.  Runs as ONE
 * persistent launch: warps take 32-row bands from a ticket counter and walk them as a
 * skewed wavefront, waiting on the band above (per-pixel tagged messages, or a progress word
 * for unaligned images / 4-pixel tiles: see b2_predict.cu; a waited-on band always holds a
 * lower ticket, so it is running: no deadlock).  Uses the stream's scratch: 8 bytes per pixel
 * of every 32nd row, cleared by the call.
 * Forward: PredictorSub0..13_C with the given modes (lossless_enc.go:339-557, driven as
 * pkg/libwebp/enc/predictor_enc.go:666-712): src = pixels, dst = residuals. */
int b2_vp8l_predictor_inverse(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs);
int b2_vp8l_predictor_forward(b2_ctx* ctx, int stream, const b2_argb_job* jobs, int n_jobs);

/* ApplyInverseTransforms (pkg/vp8/vp8l_dec.c.go:710-729) for the chain libwebp's encoder writes --
 * cross-colour inverse, then predictor inverse, then add-green -- and the 4-byte output conversion of
 * EmitRows -> VP8LConvertFromBGRA (vp8l_dec.c.go:771, pkg/libwebp/dsp/lossless.go:598-666) in ONE pass
 * over the pixels: the wavefront of the predictor inverse applies VP8LTransformColorInverse to each
 * residual before the prediction is added and VP8LAddGreenToBlueAndRed + the byte order to each pixel on
 * its way out (8 B/px instead of 24 + 8 for four launches).  color_data NULL: no cross-colour step;
 * add_green 0: no add-green step; out_mode: B2_MODE_BGRA (uint32 0xAARRGGBB as it is), B2_MODE_RGBA or
 * B2_MODE_ARGB.  dst must not alias src; rows tightly packed. */
typedef struct b2_vp8l_chain_job {
  const uint32_t* src;
  uint32_t* dst;
  const uint32_t* predictor_data; /* mode image, ceil(w/2^predictor_bits) x ceil(h/2^predictor_bits) */
  const uint32_t* color_data;     /* cross-colour codes, ceil(w/2^color_bits) x ceil(h/2^color_bits), or NULL */
  int32_t width, height;
  int32_t predictor_bits, color_bits;
  int32_t add_green;
  int32_t reserved;
} b2_vp8l_chain_job;
int b2_vp8l_predictor_inverse_chain(b2_ctx* ctx, int stream, const b2_vp8l_chain_job* jobs, int n_jobs, int out_mode);

/* ---- next rows: output conversions after parts 2 and 3 ---------
 * One job layout for the three entry points (DEVICE pointers, signed strides). */
typedef struct b2_convert_job {
  const void* src;     /* uint32 0xAARRGGBB pixels (convert) or the Y plane (point samplers); unused by alpha multiply */
  const uint8_t* u;    /* U plane (samplers) or the alpha plane (B2_ALPHA_MULT_ROW, B2_ALPHA_DISPATCH) */
  const uint8_t* v;    /* V plane (samplers) */
  uint8_t* dst;        /* output bytes; the in-place buffer for alpha multiply */
  int32_t src_stride;  /* pixels for uint32 sources, bytes for planes */
  int32_t uv_stride;   /* bytes */
  int32_t dst_stride;  /* bytes */
  int32_t width, height;
  int32_t reserved;
} b2_convert_job;

/* VP8LConvertFromBGRA (pkg/libwebp/dsp/lossless.go:616-666) with its row converters
 * (:549-614) and, for the premultiplied modes, WebPApplyAlphaMultiply{,4444}
 * (pkg/libwebp/dsp/alpha_processing.go:100-163).  mode: WEBP_CSP_MODE.  The 16-bit
 * modes follow the context's order (b2_ctx_set_swap_16bit_csp; default: ba / gb byte first as
 * the Go text does, lossless.go:581-584,593-596); B2_CONVERT_UPSTREAM_16BIT in flags forces
 * the stock libwebp order (rg first) for this call whatever the context says. */
#define B2_CONVERT_UPSTREAM_16BIT 0x1
int b2_vp8l_convert_from_bgra(b2_ctx* ctx, int stream, const b2_convert_job* jobs, int n_jobs, int mode, int flags);

/* In-place alpha multiplies on dst (pkg/libwebp/dsp/alpha_processing.go):
 *   B2_ALPHA_APPLY       ApplyAlphaMultiply_C :110-126, 4 bytes/pixel, flag = alpha_first
 *   B2_ALPHA_APPLY_4444  ApplyAlphaMultiply4444_C :141-163, 2 bytes/pixel, flag = rg_byte_pos
 *   B2_ALPHA_MULT_ARGB   WebPMultARGBRow_C / WebPMultARGBRows :33-52,73-79, uint32 pixels, flag = inverse
 *   B2_ALPHA_MULT_ROW    WebPMultRow_C / WebPMultRows :54-67,81-89, plane dst with alpha plane u, flag = inverse
 *   B2_ALPHA_DISPATCH    WebPDispatchAlpha = DispatchAlpha_C :165-180 as ExportAlpha uses it
 *                        (pkg/libwebp/decoder/io_dec.c.go:405-430): alpha plane u into the alpha byte of 4-byte
 *                        pixels, flag = alpha_first.  Its "non opaque" return value only gates the premultiply
 *                        that follows, which leaves opaque pixels alone: not reported */
enum { B2_ALPHA_APPLY = 0, B2_ALPHA_APPLY_4444 = 1, B2_ALPHA_MULT_ARGB = 2, B2_ALPHA_MULT_ROW = 3, B2_ALPHA_DISPATCH = 4 };
int b2_alpha_multiply(b2_ctx* ctx, int stream, const b2_convert_job* jobs, int n_jobs, int kind, int flag);

/* Point-sampled YUV420 -> RGB: WebPSamplers[mode] driven by WebPSamplerProcessPlane
 * (pkg/libwebp/dsp/yuv.go:19-72) == EmitSampledRGB (pkg/libwebp/decoder/io_dec.c.go:
 * 53-59); with yuv444 != 0 the WebPYUV444Converters rows (pkg/libwebp/dsp/
 * upsampling.c.go:191-245) over full-resolution chroma planes. */
int b2_yuv_to_rgb_point(b2_ctx* ctx, int stream, const b2_convert_job* jobs, int n_jobs, int mode, int yuv444);

/* ---- remaining rows of the hot-path table (a14, a28, a29) ----- */

/* Non-gamma ARGB -> Y, U, V (+A): WebPConvertARGBToY / WebPConvertARGBToUV
 * (pkg/libwebp/dsp/yuv.go:74-120) applied row by row exactly as ConvertToYUVA does
 * (pkg/vp8/vp8l_dec.c.go:543-562): even rows store U/V from 2-pixel sums, odd rows
 * average into them with (u + tmp + 1) >> 1; A = byte 3 of every pixel when a != NULL
 * (WebPExtractAlpha, alpha_processing.go:193-207).  Pixels are uint32 0xAARRGGBB;
 * argb_stride is in pixels, the plane strides in bytes.  DEVICE pointers.
 * The job covers rows y_pos .. y_pos + height - 1 of the output (ConvertToYUVA's y_pos):
 * argb, y and a point at row y_pos, u and v at chroma row y_pos >> 1.  When y_pos is odd
 * the first row averages into the U/V values an earlier call stored, exactly like the
 * streamed reference. */
typedef struct b2_plain_job {
  const uint32_t* argb;
  uint8_t* y;
  uint8_t* u;
  uint8_t* v;
  uint8_t* a; /* may be NULL */
  int32_t argb_stride;
  int32_t y_stride, uv_stride, a_stride;
  int32_t width, height;
  int32_t y_pos;
  int32_t reserved;
} b2_plain_job;
int b2_argb_to_yuva_plain(b2_ctx* ctx, int stream, const b2_plain_job* jobs, int n_jobs);

/* VP8LColorIndexInverseTransformAlpha / MapAlpha_C (pkg/libwebp/dsp/lossless.go:
 * 461-505; caller pkg/vp8/vp8l_dec.c.go:816-838): 8-bit source plane whose bytes pack
 * 1 << bits indices, 8-bit destination = green byte of the palette entry.  src rows
 * are ceil(width / 2^bits) bytes, dst rows width bytes, both tightly packed; palette
 * holds 1 << (8 >> bits) entries; bits in [0,3]; dst must not alias src. */
typedef struct b2_alpha_index_job {
  const uint8_t* src;
  uint8_t* dst;
  const uint32_t* palette;
  int32_t width, height;
  int32_t bits;
  int32_t reserved;
} b2_alpha_index_job;
int b2_vp8l_color_index_inverse_alpha(b2_ctx* ctx, int stream, const b2_alpha_index_job* jobs, int n_jobs);

/* ExpandColorMap (pkg/vp8/vp8l_dec.c.go:1196-1219): src = num_colors delta-coded
 * entries, dst = 1 << (8 >> bits) entries: byte-wise running sum (VP8LAddPixels,
 * lossless_common.go:110-114) then zero padding.  dst may equal src. */
typedef struct b2_expand_job {
  const uint32_t* src;
  uint32_t* dst;
  int32_t num_colors;
  int32_t bits;
} b2_expand_job;
int b2_vp8l_expand_color_map(b2_ctx* ctx, int stream, const b2_expand_job* jobs, int n_jobs);

/* Import() into an ARGB picture, the use_argb branch of WebPPictureImportRGB/RGBA/RGBX/
 * BGR/BGRA/BGRX (pkg/libwebp/enc/picture_csp_enc.c.go:138-181,189-225): WebPPackRGB /
 * PackRGB_C (pkg/libwebp/dsp/alpha_processing.go:227-233) for the orders without alpha,
 * the row copy of BGRA and VP8LConvertBGRAToRGBA of RGBA with it.  src = interleaved
 * bytes, src_stride in BYTES (signed), dst = uint32 0xAARRGGBB rows, dst_stride in bytes
 * (4 * Picture.ARGBStride).  import_alpha needs a 4-byte order. */
int b2_import_to_argb(b2_ctx* ctx, int stream, const b2_convert_job* jobs, int n_jobs, int order, int import_alpha);

/* ---- next rows: picture pre-processing after the import ---------
 * In-place operations on a picture the library holds on the device.  YUVA form: y, u, v,
 * a planes with byte strides (pkg/picture/picture.go:33-105 Y/U/V/A + strides); ARGB
 * form: uint32 0xAARRGGBB rows, argb_stride in PIXELS (Picture.ARGBStride). */
typedef struct b2_picture_job {
  uint8_t* y;
  uint8_t* u;
  uint8_t* v;
  uint8_t* a;
  uint32_t* argb;
  int32_t y_stride, uv_stride, a_stride; /* bytes, >= the plane's width */
  int32_t argb_stride;                   /* pixels, >= width */
  int32_t width, height;
} b2_picture_job;

/* WebPCleanupTransparentArea (pkg/libwebp/enc/picture_tools_enc.c.go:110-183): fully
 * transparent 8x8 blocks are flattened to the first value of their run along the block
 * row (need_reset carry), partially transparent blocks get their transparent luma
 * replaced by the block average (SmoothenBlock :62-91), left-overs on the right / bottom
 * are smoothened only (YUVA) or ignored (ARGB).  use_argb selects the form, as
 * Picture.UseARGB does. */
int b2_picture_cleanup_transparent(b2_ctx* ctx, int stream, const b2_picture_job* jobs, int n_jobs, int use_argb);
/* WebPBlendAlpha (picture_tools_enc.c.go:199-276): blends with background_rgb (24-bit RGB)
 * and resets alpha to 0xff; YUVA form: luma per pixel, chroma with the summed alpha of
 * its 2x2 block (BLEND / BLEND_10BIT :191-194).  The caller keeps the reference's
 * "nothing to do without an alpha plane" rule (:218-219). */
int b2_picture_blend_alpha(b2_ctx* ctx, int stream, const b2_picture_job* jobs, int n_jobs, int use_argb,
                           uint32_t background_rgb);
/* WebPReplaceTransparentPixels (picture_tools_enc.c.go:95-108; WebPAlphaReplace,
 * pkg/color/colorspace/alpha/opague.go:47-53): ARGB pixels with alpha 0 become
 * color & 0xffffff. */
int b2_picture_replace_transparent(b2_ctx* ctx, int stream, const b2_picture_job* jobs, int n_jobs, uint32_t color);

/* ---- next rows: animation pixel work --------------------------------
 * Canvas blending of the animation decoder: BlendPixelRowNonPremult / BlendPixelRowPremult
 * (pkg/libwebp/demux/anim_decode.go:195-270) over a rectangle: src (the frame just
 * decoded into the canvas) is blended IN PLACE over dst (the previous canvas) wherever its
 * alpha is not 0xff.  Strides in pixels.  Which rectangle to blend is the caller's
 * decision (FindBlendRangeAtRow, anim_decode.go:272-300). */
typedef struct b2_anim_blend_job {
  uint32_t* src;
  const uint32_t* dst;
  int32_t src_stride, dst_stride;
  int32_t width, height;
} b2_anim_blend_job;
int b2_anim_blend_rows(b2_ctx* ctx, int stream, const b2_anim_blend_job* jobs, int n_jobs, int premultiplied);

/* Frame comparison of the animation encoder: ComparePixelsLossless / ComparePixelsLossy
 * (pkg/libwebp/mux/anim_encode.go:318-367): result_dev[i] = 1 iff all `length` pixel pairs
 * of job i, `src_step` / `dst_step` pixels apart, are equal (lossless) or similar within
 * max_allowed_diff (PixelsAreSimilar :336-351).  length > 0. */
typedef struct b2_anim_compare_job {
  const uint32_t* src;
  const uint32_t* dst;
  int32_t src_step, dst_step;
  int32_t length;
  int32_t reserved;
} b2_anim_compare_job;
int b2_anim_compare_pixels(b2_ctx* ctx, int stream, const b2_anim_compare_job* jobs, int n_jobs, int lossless,
                           int max_allowed_diff, int32_t* result_dev);

/* ---- next rows: alpha-plane filters ----------------------------------
 * WEBP_FILTER_TYPE filter: 0 none, 1 horizontal, 2 vertical, 3 gradient.  Planes of bytes,
 * strides in bytes, dst != src.  b2_alpha_unfilter = WebPUnfilters[filter] row after row
 * against the previous output row (HorizontalUnfilter_C / VerticalUnfilter_C /
 * GradientUnfilter_C, pkg/libwebp/dsp/filters.go:130-171, driven as ALPHDecode /
 * AlphaApplyFilter do, pkg/vp8/vp8l_dec.c.go:802-814); b2_alpha_filter = WebPFilters[filter]
 * on the whole plane (DoHorizontalFilter_C / DoVerticalFilter_C / DoGradientFilter_C,
 * filters.go:24-112).  Jobs of one batch may use different filters. */
typedef struct b2_alpha_filter_job {
  const uint8_t* src;
  uint8_t* dst;
  int32_t src_stride, dst_stride;
  int32_t width, height;
  int32_t filter;
  int32_t reserved;
} b2_alpha_filter_job;
int b2_alpha_unfilter(b2_ctx* ctx, int stream, const b2_alpha_filter_job* jobs, int n_jobs);
int b2_alpha_filter(b2_ctx* ctx, int stream, const b2_alpha_filter_job* jobs, int n_jobs);

/* Alpha level smoothing: WebPDequantizeLevels
 * (pkg/libwebp/utils/quant_levels_dec_utils.c.go:259-286) as ALPHDecode calls it on the
 * decoded plane (pkg/libwebp/decoder/alpha_dec.go:198-203): radius = 4 * strength / 100
 * clamped to the plane, nothing happens with fewer than 3 used levels; strength outside
 * [0, 100] is B2_ERR_INVALID_CONFIG (the reference returns false).  The reference works in
 * place; here dst != src (the result is the same, see b2_dequant.cu).  ordered_dither = 1
 * is the Go file as written (DFIX = 4 and the 4x4 ordered dither, :36-45, :147-148);
 * 0 is the stock libwebp build it was transliterated from (USE_DITHERING undefined).
 * At most 65535 planes per call (B2_ERR_INVALID_CONFIG beyond). */
typedef struct b2_alpha_dequant_job {
  const uint8_t* src;
  uint8_t* dst;
  int32_t src_stride, dst_stride;
  int32_t width, height;
  int32_t strength;
  int32_t reserved;
} b2_alpha_dequant_job;
int b2_alpha_dequantize_levels(b2_ctx* ctx, int stream, const b2_alpha_dequant_job* jobs, int n_jobs, int ordered_dither);

/* The fixed-point rescaler: WebPRescalerInit / Import / Export
 * (pkg/libwebp/utils/rescaler_utils.c.go:31-158) with WebPRescalerImportRow{Expand,Shrink}
 * and WebPRescalerExportRow{Expand,Shrink} (pkg/libwebp/dsp/rescaler.c.go:39-197) over a
 * whole plane of num_channels interleaved byte channels, as RescalePlane drives them
 * (pkg/libwebp/enc/picture_rescale_enc.c.go:125-136).  Strides in bytes, dst != src. */
typedef struct b2_rescale_job {
  const uint8_t* src;
  uint8_t* dst;
  int32_t src_width, src_height, src_stride;
  int32_t dst_width, dst_height, dst_stride;
  int32_t num_channels; /* 1..4 */
  int32_t reserved;
} b2_rescale_job;
int b2_rescale_planes(b2_ctx* ctx, int stream, const b2_rescale_job* jobs, int n_jobs);

/* ---- host-staged batch entry points -------------------------------------------------
 * What the added batched Go entry points bind (picture.ImportRGBABatch,
 * picture.YUVAToRGBABatch, dsp.VP8LInverseTransformBatch): the whole
 * host -> device -> kernel -> host pipeline of a batch in one synchronous call.  'jobs'
 * is the same DEVICE-side job array the resident entry points take (the device planes of
 * pkg/picture's device-resident pictures, or library scratch), 'host' the matching HOST
 * buffers, entry by entry.  The batch is cut into chunks of 'chunk' jobs that go round the
 * context's streams, so the H2D copy of one chunk, the kernel of another and the D2H copy of
 * a third overlap.  A NULL host pointer skips that copy: inputs already resident are not
 * sent again, outputs that stay on the device for the next call are not brought back.
 * Host buffers should be pinned (b2_pinned_alloc / b2_host_register); pageable memory
 * works but is staged by the driver.  Host planes are top-down (positive strides, in
 * bytes); contiguous ranges are merged into single copies. */
typedef struct b2_import_host {
  const uint8_t* src; /* pixel (0,0) of the host image; rows job.src_stride apart (signed) */
  uint8_t* y;
  uint8_t* u;
  uint8_t* v;
  uint8_t* a;         /* copied back only when the picture turns out to have alpha (conversions.go:56-58) */
  int32_t y_stride, uv_stride, a_stride;
  int32_t reserved;
} b2_import_host;
/* Per image enc.Import -> picture.ImportYUVAFromRGBA (pkg/libwebp/enc/picture_csp_enc.c.go:123-137,
 * pkg/picture/conversions.go:50-165).  has_alpha_host: int32[n_jobs], may be NULL.
 * dithering > 0 selects the dithered import. */
int b2_import_batch_staged(b2_ctx* ctx, const b2_import_job* jobs, const b2_import_host* host, int n_jobs,
                           int order, int flags, int chunk, float dithering, int32_t* has_alpha_host);

typedef struct b2_upsample_host {
  const uint8_t* y;
  const uint8_t* u;
  const uint8_t* v;
  const uint8_t* a;
  uint8_t* dst;
  int32_t y_stride, uv_stride, a_stride, dst_stride;
} b2_upsample_host;
/* Per image WebPPictureYUVAToARGB / EmitFancyRGB (picture_csp_enc.c.go:62-118, io_dec.c.go:65-115);
 * only the rows [row_begin, row_end) a job produces are copied back. */
int b2_upsample_batch_staged(b2_ctx* ctx, const b2_upsample_job* jobs, const b2_upsample_host* host, int n_jobs,
                             int mode, int chunk);

typedef struct b2_argb_host {
  const void* src;  /* src_bytes of pixels (or indices for the bundle) */
  void* dst;
  const void* data; /* tile codes / palette / mode image */
  size_t src_bytes, dst_bytes, data_bytes;
} b2_argb_host;
typedef enum b2_vp8l_op {
  B2_VP8L_ADD_GREEN = 0, B2_VP8L_SUB_GREEN = 1, B2_VP8L_COLOR_INVERSE = 2, B2_VP8L_COLOR_FORWARD = 3,
  B2_VP8L_INVERSE_CHAIN = 4, B2_VP8L_FORWARD_CHAIN = 5, B2_VP8L_COLOR_INDEX_INVERSE = 6,
  B2_VP8L_BUNDLE_COLOR_MAP = 7, B2_VP8L_PREDICTOR_INVERSE = 8, B2_VP8L_PREDICTOR_FORWARD = 9
} b2_vp8l_op;
/* One step of VP8LInverseTransform / ApplyInverseTransforms (lossless.go:511-547,
 * pkg/vp8/vp8l_dec.c.go:710-729) or its forward counterpart over a batch of images. */
int b2_vp8l_batch_staged(b2_ctx* ctx, int op, const b2_argb_job* jobs, const b2_argb_host* host, int n_jobs,
                         int chunk);

typedef struct b2_vp8l_chain_host {
  const void* src;            /* 4 * width * height bytes of residuals, or NULL (already resident) */
  void* dst;                  /* 4 * width * height bytes, or NULL (stays on the device) */
  const void* predictor_data;
  const void* color_data;
  size_t predictor_bytes, color_bytes;
} b2_vp8l_chain_host;
/* b2_vp8l_predictor_inverse_chain over host buffers: the whole ApplyInverseTransforms chain of a batch,
 * pixels up once and down once. */
int b2_vp8l_chain_batch_staged(b2_ctx* ctx, const b2_vp8l_chain_job* jobs, const b2_vp8l_chain_host* host, int n_jobs,
                               int out_mode, int chunk);

/* ---- synthetic inputs (SURVEY.md 8d generator, identical to the oracle's) -
 * One uint32 per (row, word) position of n images laid out 'image_bytes'
 * apart, image ids first_img, first_img+1, ...  alpha_mode shapes the top byte
 * (the alpha of a little-endian RGBA/BGRA pixel): */
#define B2_SYNTH_RAW 0     /* uniform bytes everywhere                         */
#define B2_SYNTH_OPAQUE 1  /* top byte forced to 0xff                          */
#define B2_SYNTH_SPARSE 2  /* ~30% of pixels get a uniform alpha, rest 0xff     */
#define B2_SYNTH_BINARY 3  /* alpha is 0x00 or 0xff, 50/50                      */
int b2_synth_fill_batch_u32(b2_ctx* ctx, int stream, void* dst_dev, int words_per_row,
                            int rows, int stride_bytes, size_t image_bytes, int n_images,
                            uint32_t seed, uint32_t first_img, uint32_t plane,
                            int alpha_mode);

#ifdef __cplusplus
}
#endif
#endif /* B2WEBP_H_ */
