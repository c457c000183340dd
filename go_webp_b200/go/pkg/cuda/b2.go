// Package cuda is the cgo shim between go-webp and libb2webp.so (the B200 pixel
// path).  It is the only package that imports "C"; pkg/picture and
// pkg/libwebp/dsp call it through plain Go types.
//
// NOTE: written without a Go toolchain at hand (the build image has none and
// the go-webp tree does not compile); kept gofmt/go vet clean by inspection.
package cuda

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../lib -lb2webp -Wl,-rpath,${SRCDIR}/../../../lib
#include <stdlib.h>
#include "b2webp.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"unsafe"
)

// Sentinels mirror pkg/picture/errors.go:8-21; pkg/picture maps them onto its
// own variables so existing errors.Is checks keep working.
var (
	ErrOutOfMemory   = errors.New("memory error allocating objects")
	ErrBadDimension  = errors.New("picture has invalid width/height")
	ErrNilParameter  = errors.New("a pointer parameter is nil")
	ErrInvalidConfig = errors.New("configuration is invalid")
)

// Pixel orders and flags (include/b2webp.h).
const (
	OrderRGBA = int(C.B2_ORDER_RGBA)
	OrderBGRA = int(C.B2_ORDER_BGRA)
	OrderRGB  = int(C.B2_ORDER_RGB)
	OrderBGR  = int(C.B2_ORDER_BGR)

	ImportAlpha    = int(C.B2_IMPORT_ALPHA)
	ImportDivTable = int(C.B2_IMPORT_DIV_TABLE)
)

// Context owns one GPU: streams, descriptor rings, gamma tables.  It is not
// safe for concurrent use; create one per device and per OS thread.
type Context struct {
	h *C.b2_ctx
}

// DevicePtr is an address in device memory returned by Alloc.
type DevicePtr uintptr

func status(ctx *Context, st C.int) error {
	switch st {
	case C.B2_OK:
		return nil
	case C.B2_ERR_OOM:
		return ErrOutOfMemory
	case C.B2_ERR_BAD_DIM:
		return ErrBadDimension
	case C.B2_ERR_NULL:
		return ErrNilParameter
	case C.B2_ERR_INVALID_CONFIG:
		return ErrInvalidConfig
	default:
		msg := "cuda error"
		if ctx != nil && ctx.h != nil {
			msg = C.GoString(C.b2_last_error(ctx.h))
		}
		return fmt.Errorf("%w: %s", ErrOutOfMemory, msg)
	}
}

// DeviceCount reports the number of CUDA devices (0 when there is no driver).
func DeviceCount() int {
	var n C.int
	if C.b2_device_count(&n) != C.B2_OK {
		return 0
	}
	return int(n)
}

// NewContext binds a context to one device with the given number of streams.
func NewContext(device, streams int) (*Context, error) {
	var h *C.b2_ctx
	if st := C.b2_ctx_create(C.int(device), C.int(streams), &h); st != C.B2_OK {
		return nil, status(nil, st)
	}
	c := &Context{h: h}
	runtime.SetFinalizer(c, (*Context).Close)
	return c, nil
}

// Close releases every stream, event and buffer the context owns.
func (c *Context) Close() {
	if c.h != nil {
		C.b2_ctx_destroy(c.h)
		c.h = nil
	}
}

// Pinned returns page-locked host memory as a Go slice for zero-copy staging.
// The slice is owned by the library: release it with FreePinned.
func (c *Context) Pinned(n int) ([]byte, error) {
	var p unsafe.Pointer
	if st := C.b2_pinned_alloc(c.h, C.size_t(n), &p); st != C.B2_OK {
		return nil, status(c, st)
	}
	return unsafe.Slice((*byte)(p), n), nil
}

// FreePinned releases a slice obtained from Pinned.
func (c *Context) FreePinned(b []byte) error {
	if len(b) == 0 {
		return nil
	}
	return status(c, C.b2_pinned_free(c.h, unsafe.Pointer(&b[0])))
}

// Alloc reserves device memory.
func (c *Context) Alloc(n int) (DevicePtr, error) {
	var p unsafe.Pointer
	if st := C.b2_dev_alloc(c.h, C.size_t(n), &p); st != C.B2_OK {
		return 0, status(c, st)
	}
	return DevicePtr(uintptr(p)), nil
}

// Free releases device memory.
func (c *Context) Free(p DevicePtr) error {
	return status(c, C.b2_dev_free(c.h, unsafe.Pointer(uintptr(p))))
}

// H2D enqueues a host-to-device copy; src must stay alive until Sync (use
// pinned memory for true asynchrony).
func (c *Context) H2D(stream int, dst DevicePtr, src []byte) error {
	if len(src) == 0 {
		return nil
	}
	return status(c, C.b2_h2d_async(c.h, C.int(stream), unsafe.Pointer(uintptr(dst)),
		unsafe.Pointer(&src[0]), C.size_t(len(src))))
}

// D2H enqueues a device-to-host copy.
func (c *Context) D2H(stream int, dst []byte, src DevicePtr) error {
	if len(dst) == 0 {
		return nil
	}
	return status(c, C.b2_d2h_async(c.h, C.int(stream), unsafe.Pointer(&dst[0]),
		unsafe.Pointer(uintptr(src)), C.size_t(len(dst))))
}

// Sync waits for one stream, or for all when stream < 0.
func (c *Context) Sync(stream int) error {
	return status(c, C.b2_sync(c.h, C.int(stream)))
}

// ImportJob is one image of a batched RGBA->YUV420A import (device pointers).
type ImportJob struct {
	Src                            DevicePtr
	Y, U, V, A                     DevicePtr
	SrcStride                      int
	YStride, UVStride, AStride     int
	Width, Height                  int
}

func importJobs(jobs []ImportJob) []C.b2_import_job {
	cj := make([]C.b2_import_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_import_job{
			src: (*C.uint8_t)(unsafe.Pointer(uintptr(j.Src))),
			y:   (*C.uint8_t)(unsafe.Pointer(uintptr(j.Y))),
			u:   (*C.uint8_t)(unsafe.Pointer(uintptr(j.U))),
			v:   (*C.uint8_t)(unsafe.Pointer(uintptr(j.V))),
			a:   (*C.uint8_t)(unsafe.Pointer(uintptr(j.A))),
			src_stride: C.int32_t(j.SrcStride), y_stride: C.int32_t(j.YStride),
			uv_stride: C.int32_t(j.UVStride), a_stride: C.int32_t(j.AStride),
			width: C.int32_t(j.Width), height: C.int32_t(j.Height),
		}
	}
	return cj
}

// ImportRGBA replaces dsp.WebPImportYUVAFromRGBA (pkg/libwebp/dsp/yuv.go:304-382)
// and the has-alpha scan of picture.ImportYUVAFromRGBA (conversions.go:56) for a
// batch.  hasAlpha addresses int32[len(jobs)] in device memory (or 0).
func (c *Context) ImportRGBA(stream int, jobs []ImportJob, order, flags int, hasAlpha DevicePtr) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := importJobs(jobs)
	st := C.b2_rgba_to_yuv420a(c.h, C.int(stream), &cj[0], C.int(len(cj)), C.int(order),
		C.int(flags), (*C.int32_t)(unsafe.Pointer(uintptr(hasAlpha))))
	return status(c, st)
}

// UpsampleJob is one image of a batched fancy upsampling (device pointers).
type UpsampleJob struct {
	Y, U, V, A, Dst                          DevicePtr
	YStride, UVStride, AStride, DstStride    int
	Width, Height                            int
	RowBegin, RowEnd                         int // output rows [RowBegin, RowEnd); 0, 0 = whole image (EmitFancyRGB batches)
}

// UpsampleFancy replaces WebPUpsamplers[mode] driven by WebPPictureYUVAToARGB /
// EmitFancyRGB (upsampling.c.go:43-114, picture_csp_enc.c.go:62-118,
// io_dec.c.go:65-115,175-195) for a batch.  mode is a webp.WEBP_CSP_MODE.
func (c *Context) UpsampleFancy(stream int, jobs []UpsampleJob, mode int) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := make([]C.b2_upsample_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_upsample_job{
			y:   (*C.uint8_t)(unsafe.Pointer(uintptr(j.Y))),
			u:   (*C.uint8_t)(unsafe.Pointer(uintptr(j.U))),
			v:   (*C.uint8_t)(unsafe.Pointer(uintptr(j.V))),
			a:   (*C.uint8_t)(unsafe.Pointer(uintptr(j.A))),
			dst: (*C.uint8_t)(unsafe.Pointer(uintptr(j.Dst))),
			y_stride: C.int32_t(j.YStride), uv_stride: C.int32_t(j.UVStride),
			a_stride: C.int32_t(j.AStride), dst_stride: C.int32_t(j.DstStride),
			width: C.int32_t(j.Width), height: C.int32_t(j.Height),
		}
	}
	return status(c, C.b2_yuv420a_to_rgba_fancy(c.h, C.int(stream), &cj[0], C.int(len(cj)), C.int(mode)))
}

// ArgbJob is one image of a VP8L per-pixel transform (device pointers).
type ArgbJob struct {
	Src, Dst, Data DevicePtr
	Width, Height  int
	Bits           int
}

// VP8LOp selects the transform kernel.
type VP8LOp int

const (
	AddGreen VP8LOp = iota
	SubGreen
	ColorInverse
	ColorForward
	InverseChain
	ForwardChain
	ColorIndexInverse
	BundleColorMap
	PredictorInverse
	PredictorForward
)

// VP8L runs one batched VP8L transform (pkg/libwebp/dsp/lossless.go:337-459,
// lossless_enc.go:207-312).
func (c *Context) VP8L(stream int, op VP8LOp, jobs []ArgbJob) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := make([]C.b2_argb_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_argb_job{
			src:  (*C.uint32_t)(unsafe.Pointer(uintptr(j.Src))),
			dst:  (*C.uint32_t)(unsafe.Pointer(uintptr(j.Dst))),
			data: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Data))),
			width: C.int32_t(j.Width), height: C.int32_t(j.Height), bits: C.int32_t(j.Bits),
		}
	}
	n, s := C.int(len(cj)), C.int(stream)
	var st C.int
	switch op {
	case AddGreen:
		st = C.b2_vp8l_add_green(c.h, s, &cj[0], n)
	case SubGreen:
		st = C.b2_vp8l_sub_green(c.h, s, &cj[0], n)
	case ColorInverse:
		st = C.b2_vp8l_color_inverse(c.h, s, &cj[0], n)
	case ColorForward:
		st = C.b2_vp8l_color_forward(c.h, s, &cj[0], n)
	case InverseChain:
		st = C.b2_vp8l_inverse_chain(c.h, s, &cj[0], n)
	case ForwardChain:
		st = C.b2_vp8l_forward_chain(c.h, s, &cj[0], n)
	case ColorIndexInverse:
		st = C.b2_vp8l_color_index_inverse(c.h, s, &cj[0], n)
	case BundleColorMap:
		st = C.b2_vp8l_bundle_color_map(c.h, s, &cj[0], n)
	case PredictorInverse:
		st = C.b2_vp8l_predictor_inverse(c.h, s, &cj[0], n)
	case PredictorForward:
		st = C.b2_vp8l_predictor_forward(c.h, s, &cj[0], n)
	default:
		return ErrInvalidConfig
	}
	return status(c, st)
}

// PlainJob is one chunk of rows of the non-gamma ARGB -> YUVA conversion (device
// pointers; Argb/Y/A at row YPos, U/V at chroma row YPos>>1).
type PlainJob struct {
	Argb                          DevicePtr
	Y, U, V, A                    DevicePtr
	ArgbStride                    int // pixels
	YStride, UVStride, AStride    int
	Width, Height, YPos           int
}

// ArgbToYUVAPlain replaces WebPConvertARGBToY / WebPConvertARGBToUV
// (pkg/libwebp/dsp/yuv.go:74-120) as driven by ConvertToYUVA
// (pkg/vp8/vp8l_dec.c.go:543-562) for a batch.
func (c *Context) ArgbToYUVAPlain(stream int, jobs []PlainJob) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := make([]C.b2_plain_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_plain_job{
			argb: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Argb))),
			y:    (*C.uint8_t)(unsafe.Pointer(uintptr(j.Y))),
			u:    (*C.uint8_t)(unsafe.Pointer(uintptr(j.U))),
			v:    (*C.uint8_t)(unsafe.Pointer(uintptr(j.V))),
			a:    (*C.uint8_t)(unsafe.Pointer(uintptr(j.A))),
			argb_stride: C.int32_t(j.ArgbStride), y_stride: C.int32_t(j.YStride),
			uv_stride: C.int32_t(j.UVStride), a_stride: C.int32_t(j.AStride),
			width: C.int32_t(j.Width), height: C.int32_t(j.Height), y_pos: C.int32_t(j.YPos),
		}
	}
	return status(c, C.b2_argb_to_yuva_plain(c.h, C.int(stream), &cj[0], C.int(len(cj))))
}

// AlphaIndexJob is one 8-bit plane of VP8LColorIndexInverseTransformAlpha.
type AlphaIndexJob struct {
	Src, Dst, Palette      DevicePtr
	Width, Height, Bits    int
}

// ColorIndexInverseAlpha replaces VP8LColorIndexInverseTransformAlpha / MapAlpha_C
// (pkg/libwebp/dsp/lossless.go:461-505) for a batch.
func (c *Context) ColorIndexInverseAlpha(stream int, jobs []AlphaIndexJob) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := make([]C.b2_alpha_index_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_alpha_index_job{
			src:     (*C.uint8_t)(unsafe.Pointer(uintptr(j.Src))),
			dst:     (*C.uint8_t)(unsafe.Pointer(uintptr(j.Dst))),
			palette: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Palette))),
			width:   C.int32_t(j.Width), height: C.int32_t(j.Height), bits: C.int32_t(j.Bits),
		}
	}
	return status(c, C.b2_vp8l_color_index_inverse_alpha(c.h, C.int(stream), &cj[0], C.int(len(cj))))
}

// ExpandJob is one palette of ExpandColorMap.
type ExpandJob struct {
	Src, Dst           DevicePtr
	NumColors, Bits    int
}

// ExpandColorMap replaces ExpandColorMap (pkg/vp8/vp8l_dec.c.go:1196-1219) for a batch.
func (c *Context) ExpandColorMap(stream int, jobs []ExpandJob) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := make([]C.b2_expand_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_expand_job{
			src: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Src))),
			dst: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Dst))),
			num_colors: C.int32_t(j.NumColors), bits: C.int32_t(j.Bits),
		}
	}
	return status(c, C.b2_vp8l_expand_color_map(c.h, C.int(stream), &cj[0], C.int(len(cj))))
}

// PredictorInverse / PredictorForward run the VP8L predictor transform on whole images
// (pkg/libwebp/dsp/lossless.go:31-333, lossless_enc.go:339-557); jobs[i].Data is the
// mode image, Bits the transform's size bits.
func (c *Context) PredictorInverse(stream int, jobs []ArgbJob) error { return c.predictor(stream, jobs, true) }
func (c *Context) PredictorForward(stream int, jobs []ArgbJob) error { return c.predictor(stream, jobs, false) }

func (c *Context) predictor(stream int, jobs []ArgbJob, inverse bool) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := make([]C.b2_argb_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_argb_job{
			src:  (*C.uint32_t)(unsafe.Pointer(uintptr(j.Src))),
			dst:  (*C.uint32_t)(unsafe.Pointer(uintptr(j.Dst))),
			data: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Data))),
			width: C.int32_t(j.Width), height: C.int32_t(j.Height), bits: C.int32_t(j.Bits),
		}
	}
	if inverse {
		return status(c, C.b2_vp8l_predictor_inverse(c.h, C.int(stream), &cj[0], C.int(len(cj))))
	}
	return status(c, C.b2_vp8l_predictor_forward(c.h, C.int(stream), &cj[0], C.int(len(cj))))
}

// ConvertJob is one image of the output conversions (include/b2webp.h b2_convert_job).
type ConvertJob struct {
	Src, U, V, Dst                    DevicePtr
	SrcStride, UVStride, DstStride    int
	Width, Height                     int
}

func convertJobs(jobs []ConvertJob) []C.b2_convert_job {
	cj := make([]C.b2_convert_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_convert_job{
			src: unsafe.Pointer(uintptr(j.Src)),
			u:   (*C.uint8_t)(unsafe.Pointer(uintptr(j.U))),
			v:   (*C.uint8_t)(unsafe.Pointer(uintptr(j.V))),
			dst: (*C.uint8_t)(unsafe.Pointer(uintptr(j.Dst))),
			src_stride: C.int32_t(j.SrcStride), uv_stride: C.int32_t(j.UVStride), dst_stride: C.int32_t(j.DstStride),
			width: C.int32_t(j.Width), height: C.int32_t(j.Height),
		}
	}
	return cj
}

// ConvertFromBGRA replaces VP8LConvertFromBGRA (pkg/libwebp/dsp/lossless.go:616-666).
func (c *Context) ConvertFromBGRA(stream int, jobs []ConvertJob, mode, flags int) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := convertJobs(jobs)
	return status(c, C.b2_vp8l_convert_from_bgra(c.h, C.int(stream), &cj[0], C.int(len(cj)), C.int(mode), C.int(flags)))
}

// AlphaMultiply replaces ApplyAlphaMultiply_C / ..4444_C / WebPMultARGBRow_C / WebPMultRow_C
// (pkg/libwebp/dsp/alpha_processing.go:17-163) and, with kind = C.B2_ALPHA_DISPATCH,
// DispatchAlpha_C (:165-180); kind is one of the B2_ALPHA_* values.
func (c *Context) AlphaMultiply(stream int, jobs []ConvertJob, kind, flag int) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := convertJobs(jobs)
	return status(c, C.b2_alpha_multiply(c.h, C.int(stream), &cj[0], C.int(len(cj)), C.int(kind), C.int(flag)))
}

// YUVToRGBPoint replaces WebPSamplers[]/WebPSamplerProcessPlane (pkg/libwebp/dsp/yuv.go:19-72)
// and, with yuv444, WebPYUV444Converters[] (pkg/libwebp/dsp/upsampling.c.go:191-245).
func (c *Context) YUVToRGBPoint(stream int, jobs []ConvertJob, mode int, yuv444 bool) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := convertJobs(jobs)
	f := 0
	if yuv444 {
		f = 1
	}
	return status(c, C.b2_yuv_to_rgb_point(c.h, C.int(stream), &cj[0], C.int(len(cj)), C.int(mode), C.int(f)))
}

// PictureJob is one device-resident picture of the pre-processing entry points
// (include/b2webp.h b2_picture_job): Y/U/V/A planes with byte strides, or ARGB rows with
// a stride in pixels (pkg/picture/picture.go:33-105).
type PictureJob struct {
	Y, U, V, A, ARGB                       DevicePtr
	YStride, UVStride, AStride, ARGBStride int
	Width, Height                          int
}

func pictureJobs(jobs []PictureJob) []C.b2_picture_job {
	cj := make([]C.b2_picture_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_picture_job{
			y: (*C.uint8_t)(unsafe.Pointer(uintptr(j.Y))), u: (*C.uint8_t)(unsafe.Pointer(uintptr(j.U))),
			v: (*C.uint8_t)(unsafe.Pointer(uintptr(j.V))), a: (*C.uint8_t)(unsafe.Pointer(uintptr(j.A))),
			argb:     (*C.uint32_t)(unsafe.Pointer(uintptr(j.ARGB))),
			y_stride: C.int32_t(j.YStride), uv_stride: C.int32_t(j.UVStride), a_stride: C.int32_t(j.AStride),
			argb_stride: C.int32_t(j.ARGBStride), width: C.int32_t(j.Width), height: C.int32_t(j.Height),
		}
	}
	return cj
}

func boolInt(b bool) C.int {
	if b {
		return 1
	}
	return 0
}

// CleanupTransparentArea replaces WebPCleanupTransparentArea
// (pkg/libwebp/enc/picture_tools_enc.c.go:110-183), called by the encoder right after the
// import when !config.Exact (pkg/libwebp/enc/webp_encoding.go:49-72).
func (c *Context) CleanupTransparentArea(stream int, jobs []PictureJob, useARGB bool) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := pictureJobs(jobs)
	return status(c, C.b2_picture_cleanup_transparent(c.h, C.int(stream), &cj[0], C.int(len(cj)), boolInt(useARGB)))
}

// BlendAlpha replaces WebPBlendAlpha (pkg/libwebp/enc/picture_tools_enc.c.go:199-276).
func (c *Context) BlendAlpha(stream int, jobs []PictureJob, useARGB bool, backgroundRGB uint32) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := pictureJobs(jobs)
	return status(c, C.b2_picture_blend_alpha(c.h, C.int(stream), &cj[0], C.int(len(cj)), boolInt(useARGB), C.uint32_t(backgroundRGB)))
}

// ReplaceTransparentPixels replaces WebPReplaceTransparentPixels
// (pkg/libwebp/enc/picture_tools_enc.c.go:95-108).
func (c *Context) ReplaceTransparentPixels(stream int, jobs []PictureJob, color uint32) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := pictureJobs(jobs)
	return status(c, C.b2_picture_replace_transparent(c.h, C.int(stream), &cj[0], C.int(len(cj)), C.uint32_t(color)))
}

// ImportToARGB replaces the use_argb branch of Import()
// (pkg/libwebp/enc/picture_csp_enc.c.go:138-181): WebPPackRGB, the BGRA row copy and
// VP8LConvertBGRAToRGBA.  SrcStride is the caller's rgb_stride in bytes (may be negative),
// DstStride = 4 * Picture.ARGBStride.
func (c *Context) ImportToARGB(stream int, jobs []ConvertJob, order int, importAlpha bool) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := convertJobs(jobs)
	return status(c, C.b2_import_to_argb(c.h, C.int(stream), &cj[0], C.int(len(cj)), C.int(order), boolInt(importAlpha)))
}

// ImportRGBADithered replaces the dithered copy of the import (pkg/picture/conversions.go:
// 122-165; ConvertRowToY / ConvertRowsToUV, pkg/libwebp/enc/picture_csp_enc.c.go:31-48;
// VP8Random, pkg/vp8/random.go).  dithering as PictureARGBToYUVA's argument.
func (c *Context) ImportRGBADithered(stream int, jobs []ImportJob, order, flags int, hasAlpha DevicePtr, dithering float32) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := importJobs(jobs)
	return status(c, C.b2_rgba_to_yuv420a_dithered(c.h, C.int(stream), &cj[0], C.int(len(cj)), C.int(order), C.int(flags),
		(*C.int32_t)(unsafe.Pointer(uintptr(hasAlpha))), C.float(dithering)))
}

// AnimBlendJob is one rectangle of the animation decoder's canvas blend
// (include/b2webp.h b2_anim_blend_job); strides in pixels.
type AnimBlendJob struct {
	Src, Dst             DevicePtr
	SrcStride, DstStride int
	Width, Height        int
}

// AnimBlendRows replaces BlendPixelRowNonPremult / BlendPixelRowPremult
// (pkg/libwebp/demux/anim_decode.go:234-270) over whole rectangles.
func (c *Context) AnimBlendRows(stream int, jobs []AnimBlendJob, premultiplied bool) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := make([]C.b2_anim_blend_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_anim_blend_job{
			src: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Src))), dst: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Dst))),
			src_stride: C.int32_t(j.SrcStride), dst_stride: C.int32_t(j.DstStride),
			width: C.int32_t(j.Width), height: C.int32_t(j.Height),
		}
	}
	return status(c, C.b2_anim_blend_rows(c.h, C.int(stream), &cj[0], C.int(len(cj)), boolInt(premultiplied)))
}

// AnimCompareJob is one row or column scan of the animation encoder's frame comparison.
type AnimCompareJob struct {
	Src, Dst         DevicePtr
	SrcStep, DstStep int
	Length           int
}

// AnimComparePixels replaces ComparePixelsLossless / ComparePixelsLossy
// (pkg/libwebp/mux/anim_encode.go:318-367); result addresses int32[len(jobs)] on the device.
func (c *Context) AnimComparePixels(stream int, jobs []AnimCompareJob, lossless bool, maxAllowedDiff int, result DevicePtr) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := make([]C.b2_anim_compare_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_anim_compare_job{
			src: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Src))), dst: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Dst))),
			src_step: C.int32_t(j.SrcStep), dst_step: C.int32_t(j.DstStep), length: C.int32_t(j.Length),
		}
	}
	return status(c, C.b2_anim_compare_pixels(c.h, C.int(stream), &cj[0], C.int(len(cj)), boolInt(lossless),
		C.int(maxAllowedDiff), (*C.int32_t)(unsafe.Pointer(uintptr(result)))))
}

// AlphaFilterJob is one alpha plane of the filter / unfilter entry points
// (include/b2webp.h b2_alpha_filter_job); Filter is a WEBP_FILTER_TYPE.
type AlphaFilterJob struct {
	Src, Dst             DevicePtr
	SrcStride, DstStride int
	Width, Height        int
	Filter               int
}

func alphaFilterJobs(jobs []AlphaFilterJob) []C.b2_alpha_filter_job {
	cj := make([]C.b2_alpha_filter_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_alpha_filter_job{
			src: (*C.uint8_t)(unsafe.Pointer(uintptr(j.Src))), dst: (*C.uint8_t)(unsafe.Pointer(uintptr(j.Dst))),
			src_stride: C.int32_t(j.SrcStride), dst_stride: C.int32_t(j.DstStride),
			width: C.int32_t(j.Width), height: C.int32_t(j.Height), filter: C.int32_t(j.Filter),
		}
	}
	return cj
}

// AlphaUnfilter replaces WebPUnfilters[filter] applied row after row by the alpha decoder
// (pkg/libwebp/dsp/filters.go:130-171, pkg/vp8/vp8l_dec.c.go:802-814).
func (c *Context) AlphaUnfilter(stream int, jobs []AlphaFilterJob) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := alphaFilterJobs(jobs)
	return status(c, C.b2_alpha_unfilter(c.h, C.int(stream), &cj[0], C.int(len(cj))))
}

// AlphaFilter replaces WebPFilters[filter] (pkg/libwebp/dsp/filters.go:24-124).
func (c *Context) AlphaFilter(stream int, jobs []AlphaFilterJob) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := alphaFilterJobs(jobs)
	return status(c, C.b2_alpha_filter(c.h, C.int(stream), &cj[0], C.int(len(cj))))
}

// RescaleJob is one plane of the rescaler (include/b2webp.h b2_rescale_job); strides in bytes.
type RescaleJob struct {
	Src, Dst                         DevicePtr
	SrcWidth, SrcHeight, SrcStride   int
	DstWidth, DstHeight, DstStride   int
	NumChannels                      int
}

// RescalePlanes replaces RescalePlane (pkg/libwebp/enc/picture_rescale_enc.c.go:125-136):
// WebPRescalerInit / Import / Export with the import / export row functions of
// pkg/libwebp/dsp/rescaler.c.go.
func (c *Context) RescalePlanes(stream int, jobs []RescaleJob) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := make([]C.b2_rescale_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_rescale_job{
			src: (*C.uint8_t)(unsafe.Pointer(uintptr(j.Src))), dst: (*C.uint8_t)(unsafe.Pointer(uintptr(j.Dst))),
			src_width: C.int32_t(j.SrcWidth), src_height: C.int32_t(j.SrcHeight), src_stride: C.int32_t(j.SrcStride),
			dst_width: C.int32_t(j.DstWidth), dst_height: C.int32_t(j.DstHeight), dst_stride: C.int32_t(j.DstStride),
			num_channels: C.int32_t(j.NumChannels),
		}
	}
	return status(c, C.b2_rescale_planes(c.h, C.int(stream), &cj[0], C.int(len(cj))))
}

// AlphaDequantJob is one alpha plane of a DequantizeLevels batch (device pointers, strides in
// bytes, Dst != Src).
type AlphaDequantJob struct {
	Src, Dst             DevicePtr
	SrcStride, DstStride int
	Width, Height        int
	Strength             int // options.alpha_dithering_strength, 0..100
}

// DequantizeLevels replaces WebPDequantizeLevels
// (pkg/libwebp/utils/quant_levels_dec_utils.c.go:259-286) as ALPHDecode calls it on the decoded
// alpha plane (pkg/libwebp/decoder/alpha_dec.go:198-203).  orderedDither = true is this
// tree's text (DFIX = 4 with the 4x4 dither table); false is the stock libwebp build.
func (c *Context) DequantizeLevels(stream int, jobs []AlphaDequantJob, orderedDither bool) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := make([]C.b2_alpha_dequant_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_alpha_dequant_job{
			src: (*C.uint8_t)(unsafe.Pointer(uintptr(j.Src))), dst: (*C.uint8_t)(unsafe.Pointer(uintptr(j.Dst))),
			src_stride: C.int32_t(j.SrcStride), dst_stride: C.int32_t(j.DstStride),
			width: C.int32_t(j.Width), height: C.int32_t(j.Height), strength: C.int32_t(j.Strength),
		}
	}
	return status(c, C.b2_alpha_dequantize_levels(c.h, C.int(stream), &cj[0], C.int(len(cj)), boolInt(orderedDither)))
}
