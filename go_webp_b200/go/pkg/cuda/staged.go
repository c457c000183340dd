// Host-staged batch entry points and device-resident picture memory (include/b2webp.h,
// "host-staged batch entry points").  One cgo call moves a whole batch host -> device ->
// kernel -> host, chunk-pipelined over the context's streams inside the library, so the Go
// side never loops over images for copies.
//
// NOTE: written without a Go toolchain at hand (see b2.go); the tested twin is
// go_webp_b200/picture.py + dsp.py over the same C entry points.
package cuda

/*
#include <stdlib.h>
#include "b2webp.h"
*/
import "C"

import "unsafe"

// Block is one pinned-host block and one device block of the same size and layout: the
// memory behind the device-resident pictures of a batch (Picture.memory_ / memory_argb_,
// pkg/picture/picture.go:100-103).  Host is the pinned mirror as a Go slice.
type Block struct {
	ctx  *Context
	Dev  DevicePtr
	Host []byte
}

// NewBlock allocates n bytes of device memory and n bytes of pinned host memory.
func (c *Context) NewBlock(n int) (*Block, error) {
	d, err := c.Alloc(n)
	if err != nil {
		return nil, err
	}
	h, err := c.Pinned(n)
	if err != nil {
		c.Free(d)
		return nil, err
	}
	return &Block{ctx: c, Dev: d, Host: h}, nil
}

// Free releases both halves (WebPPictureFree of the last picture of the batch).
func (b *Block) Free() {
	if b == nil || b.ctx == nil {
		return
	}
	b.ctx.Free(b.Dev)
	b.ctx.FreePinned(b.Host)
	b.ctx = nil
}

// CopyStats reports the copies issued through the context: calls and bytes per direction.
func (c *Context) CopyStats() (h2dCalls, h2dBytes, d2hCalls, d2hBytes uint64) {
	var out [4]C.uint64_t
	C.b2_copy_stats(c.h, &out[0])
	return uint64(out[0]), uint64(out[1]), uint64(out[2]), uint64(out[3])
}

// SetSwap16BitCSP selects the byte order of the 16-bit output modes library-wide: true (the
// default) is the order of the go-webp text (WEBP_SWAP_16BIT_CSP), false the stock libwebp one.
func (c *Context) SetSwap16BitCSP(swap bool) error {
	v := C.int(0)
	if swap {
		v = 1
	}
	return status(c, C.b2_ctx_set_swap_16bit_csp(c.h, v))
}

func hostPtr(b []byte) unsafe.Pointer {
	if len(b) == 0 {
		return nil
	}
	return unsafe.Pointer(&b[0])
}

// ImportHost is the host side of one image of ImportBatchStaged: Src addresses pixel (0,0)
// (rows ImportJob.SrcStride apart); Y, U, V, A are the pinned planes the result is copied
// into (nil: that plane stays on the device).  A comes back only when the image has alpha.
type ImportHost struct {
	Src                        []byte
	Y, U, V, A                 []byte
	YStride, UVStride, AStride int
}

// ImportBatchStaged is enc.Import -> picture.ImportYUVAFromRGBA for a batch in one call
// (b2_import_batch_staged).  Returns the per-image has-alpha flags.
func (c *Context) ImportBatchStaged(jobs []ImportJob, host []ImportHost, order, flags, chunk int, dithering float32) ([]bool, error) {
	if len(jobs) == 0 {
		return nil, nil
	}
	if len(jobs) != len(host) {
		return nil, ErrInvalidConfig
	}
	cj := importJobs(jobs)
	ch := make([]C.b2_import_host, len(host))
	for i, h := range host {
		ch[i] = C.b2_import_host{
			src: (*C.uint8_t)(hostPtr(h.Src)), y: (*C.uint8_t)(hostPtr(h.Y)), u: (*C.uint8_t)(hostPtr(h.U)),
			v: (*C.uint8_t)(hostPtr(h.V)), a: (*C.uint8_t)(hostPtr(h.A)),
			y_stride: C.int32_t(h.YStride), uv_stride: C.int32_t(h.UVStride), a_stride: C.int32_t(h.AStride),
		}
	}
	flagsOut := make([]C.int32_t, len(jobs))
	// the Go slices are reachable from cj / ch only through C pointers: keep them alive over the call
	st := C.b2_import_batch_staged(c.h, &cj[0], &ch[0], C.int(len(cj)), C.int(order), C.int(flags), C.int(chunk),
		C.float(dithering), &flagsOut[0])
	if st != C.B2_OK {
		return nil, status(c, st)
	}
	out := make([]bool, len(jobs))
	for i, f := range flagsOut {
		out[i] = f != 0
	}
	return out, nil
}

// UpsampleHost is the host side of one image of UpsampleBatchStaged (nil = no copy).
type UpsampleHost struct {
	Y, U, V, A, Dst                       []byte
	YStride, UVStride, AStride, DstStride int
}

// UpsampleBatchStaged is WebPPictureYUVAToARGB / EmitFancyRGB for a batch in one call.
func (c *Context) UpsampleBatchStaged(jobs []UpsampleJob, host []UpsampleHost, mode, chunk int) error {
	if len(jobs) == 0 {
		return nil
	}
	if len(jobs) != len(host) {
		return ErrInvalidConfig
	}
	cj := make([]C.b2_upsample_job, len(jobs))
	ch := make([]C.b2_upsample_host, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_upsample_job{
			y: (*C.uint8_t)(unsafe.Pointer(uintptr(j.Y))), u: (*C.uint8_t)(unsafe.Pointer(uintptr(j.U))),
			v: (*C.uint8_t)(unsafe.Pointer(uintptr(j.V))), a: (*C.uint8_t)(unsafe.Pointer(uintptr(j.A))),
			dst: (*C.uint8_t)(unsafe.Pointer(uintptr(j.Dst))),
			y_stride: C.int32_t(j.YStride), uv_stride: C.int32_t(j.UVStride), a_stride: C.int32_t(j.AStride),
			dst_stride: C.int32_t(j.DstStride), width: C.int32_t(j.Width), height: C.int32_t(j.Height),
			row_begin: C.int32_t(j.RowBegin), row_end: C.int32_t(j.RowEnd),
		}
		h := host[i]
		ch[i] = C.b2_upsample_host{
			y: (*C.uint8_t)(hostPtr(h.Y)), u: (*C.uint8_t)(hostPtr(h.U)), v: (*C.uint8_t)(hostPtr(h.V)),
			a: (*C.uint8_t)(hostPtr(h.A)), dst: (*C.uint8_t)(hostPtr(h.Dst)),
			y_stride: C.int32_t(h.YStride), uv_stride: C.int32_t(h.UVStride), a_stride: C.int32_t(h.AStride),
			dst_stride: C.int32_t(h.DstStride),
		}
	}
	return status(c, C.b2_upsample_batch_staged(c.h, &cj[0], &ch[0], C.int(len(cj)), C.int(mode), C.int(chunk)))
}

// ChainJob is one image of the fused VP8L inverse chain (device pointers): cross-colour
// inverse (ColorData != 0), predictor inverse, add-green (AddGreen) in one pass.
type ChainJob struct {
	Src, Dst, PredictorData, ColorData DevicePtr
	Width, Height                      int
	PredictorBits, ColorBits           int
	AddGreen                           bool
}

// ChainHost is its host side: residuals in, pixels out, the two sub-resolution images.
type ChainHost struct {
	Src, Dst                 []uint32
	PredictorData, ColorData []uint32
}

func u32Ptr(b []uint32) unsafe.Pointer {
	if len(b) == 0 {
		return nil
	}
	return unsafe.Pointer(&b[0])
}

// InverseChainBatchStaged is ApplyInverseTransforms (pkg/vp8/vp8l_dec.c.go:710-729) + the 4-byte
// VP8LConvertFromBGRA of EmitRows (:771) for a batch of images in one call.
func (c *Context) InverseChainBatchStaged(jobs []ChainJob, host []ChainHost, outMode, chunk int) error {
	if len(jobs) == 0 {
		return nil
	}
	if len(jobs) != len(host) {
		return ErrInvalidConfig
	}
	cj := make([]C.b2_vp8l_chain_job, len(jobs))
	ch := make([]C.b2_vp8l_chain_host, len(jobs))
	for i, j := range jobs {
		g := C.int32_t(0)
		if j.AddGreen {
			g = 1
		}
		cj[i] = C.b2_vp8l_chain_job{
			src: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Src))), dst: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Dst))),
			predictor_data: (*C.uint32_t)(unsafe.Pointer(uintptr(j.PredictorData))),
			color_data:     (*C.uint32_t)(unsafe.Pointer(uintptr(j.ColorData))),
			width: C.int32_t(j.Width), height: C.int32_t(j.Height),
			predictor_bits: C.int32_t(j.PredictorBits), color_bits: C.int32_t(j.ColorBits), add_green: g,
		}
		h := host[i]
		ch[i] = C.b2_vp8l_chain_host{
			src: u32Ptr(h.Src), dst: u32Ptr(h.Dst), predictor_data: u32Ptr(h.PredictorData), color_data: u32Ptr(h.ColorData),
			predictor_bytes: C.size_t(4 * len(h.PredictorData)), color_bytes: C.size_t(4 * len(h.ColorData)),
		}
	}
	return status(c, C.b2_vp8l_chain_batch_staged(c.h, &cj[0], &ch[0], C.int(len(cj)), C.int(outMode), C.int(chunk)))
}

// ArgbHost is the host side of one image of VP8LBatchStaged.
type ArgbHost struct {
	Src, Dst, Data []uint32
}

// VP8LBatchStaged runs one transform step over a batch with its copies (b2_vp8l_batch_staged).
func (c *Context) VP8LBatchStaged(op VP8LOp, jobs []ArgbJob, host []ArgbHost, chunk int) error {
	if len(jobs) == 0 {
		return nil
	}
	if len(jobs) != len(host) {
		return ErrInvalidConfig
	}
	cj := make([]C.b2_argb_job, len(jobs))
	ch := make([]C.b2_argb_host, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_argb_job{
			src: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Src))), dst: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Dst))),
			data: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Data))),
			width: C.int32_t(j.Width), height: C.int32_t(j.Height), bits: C.int32_t(j.Bits),
		}
		h := host[i]
		ch[i] = C.b2_argb_host{src: u32Ptr(h.Src), dst: u32Ptr(h.Dst), data: u32Ptr(h.Data),
			src_bytes: C.size_t(4 * len(h.Src)), dst_bytes: C.size_t(4 * len(h.Dst)), data_bytes: C.size_t(4 * len(h.Data))}
	}
	return status(c, C.b2_vp8l_batch_staged(c.h, C.int(op), &cj[0], &ch[0], C.int(len(cj)), C.int(chunk)))
}

// PaletteJob is one image of ApplyPalette (device pointers; strides in pixels / words).
type PaletteJob struct {
	Src, Dst, Palette      DevicePtr
	SrcStride, DstStride   int
	Width, Height          int
	PaletteSize, XBits     int
}

// ApplyPalette replaces ApplyPalette + VP8LBundleColorMap (pkg/vp8/vp8l_encoder.go:1130-1212,
// pkg/libwebp/dsp/lossless_enc.go:293-312) for a batch.
func (c *Context) ApplyPalette(stream int, jobs []PaletteJob) error {
	if len(jobs) == 0 {
		return nil
	}
	cj := make([]C.b2_palette_job, len(jobs))
	for i, j := range jobs {
		cj[i] = C.b2_palette_job{
			src: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Src))), dst: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Dst))),
			palette: (*C.uint32_t)(unsafe.Pointer(uintptr(j.Palette))),
			src_stride: C.int32_t(j.SrcStride), dst_stride: C.int32_t(j.DstStride),
			width: C.int32_t(j.Width), height: C.int32_t(j.Height),
			palette_size: C.int32_t(j.PaletteSize), xbits: C.int32_t(j.XBits),
		}
	}
	return status(c, C.b2_vp8l_apply_palette(c.h, C.int(stream), &cj[0], C.int(len(cj))))
}
