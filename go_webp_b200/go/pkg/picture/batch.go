// Batched, GPU-backed entry points added to pkg/picture (SURVEY.md 8b).  The
// existing signatures (ImportYUVAFromRGBA, PictureARGBToYUVA, ...) keep their
// parameter lists and forward here with a batch of one.
//
// NOTE: untested Go (no toolchain in the build image); the tested twin of this
// file is go_webp_b200/picture.py.
package picture

import (
	"github.com/daanv2/go-webp/pkg/color/colorspace"
	"github.com/daanv2/go-webp/pkg/cuda"
)

// ImportOptions selects the source layout of ImportRGBABatch.
type ImportOptions struct {
	Order       int  // cuda.OrderRGBA, OrderBGRA, OrderRGB, OrderBGR
	ImportAlpha bool // picture_csp_enc.c.go:123 import_alpha
	Chunk       int  // images per pipelined chunk (default 16)
}

// ImportRGBABatch imports srcs[i] (row stride strides[i]) into pics[i] as
// YUV420(A): per image exactly picture.ImportYUVAFromRGBA with dithering = 0 and
// use_iterative_conversion = 0 (conversions.go:50-121).  Chunks of the batch are
// pipelined over the context's streams so H2D, kernel and D2H overlap.
func ImportRGBABatch(ctx *cuda.Context, pics []*Picture, srcs [][]uint8, strides []int, opt ImportOptions) error {
	if len(pics) != len(srcs) || len(pics) != len(strides) {
		return ENC_ERROR_INVALID_CONFIGURATION
	}
	chunk := opt.Chunk
	if chunk <= 0 {
		chunk = 16
	}
	step := 4
	if opt.Order == cuda.OrderRGB || opt.Order == cuda.OrderBGR {
		step = 3
	}
	flags := 0
	if opt.ImportAlpha {
		flags |= cuda.ImportAlpha
	}
	type devImg struct {
		src, y, u, v, a cuda.DevicePtr
		ys, uvs         int
	}
	imgs := make([]devImg, len(pics))
	jobs := make([]cuda.ImportJob, len(pics))
	hasAlphaDev, err := ctx.Alloc(4 * len(pics))
	if err != nil {
		return err
	}
	defer ctx.Free(hasAlphaDev)
	for i, p := range pics {
		if p == nil || srcs[i] == nil {
			return ENC_ERROR_nil_PARAMETER
		}
		if err := WebPValidatePicture(p); err != nil {
			return p.SetEncodingError(err)
		}
		w, h := p.Width, p.Height
		uvw, uvh := (w+1)>>1, (h+1)>>1
		d := devImg{ys: (w + 15) &^ 15, uvs: (uvw + 15) &^ 15}
		abs := strides[i]
		if abs < 0 {
			abs = -abs
		}
		if d.src, err = ctx.Alloc(abs * h); err != nil {
			return p.SetEncodingError(ENC_ERROR_OUT_OF_MEMORY)
		}
		if d.y, err = ctx.Alloc(d.ys * h); err != nil {
			return p.SetEncodingError(ENC_ERROR_OUT_OF_MEMORY)
		}
		if d.u, err = ctx.Alloc(d.uvs * uvh); err != nil {
			return p.SetEncodingError(ENC_ERROR_OUT_OF_MEMORY)
		}
		if d.v, err = ctx.Alloc(d.uvs * uvh); err != nil {
			return p.SetEncodingError(ENC_ERROR_OUT_OF_MEMORY)
		}
		if d.a, err = ctx.Alloc(d.ys * h); err != nil {
			return p.SetEncodingError(ENC_ERROR_OUT_OF_MEMORY)
		}
		imgs[i] = d
		src0 := d.src
		if strides[i] < 0 {
			src0 += cuda.DevicePtr(abs * (h - 1))
		}
		jobs[i] = cuda.ImportJob{Src: src0, Y: d.y, U: d.u, V: d.v, A: d.a,
			SrcStride: strides[i], YStride: d.ys, UVStride: d.uvs, AStride: d.ys, Width: w, Height: h}
		p.ColorSpace = colorspace.WEBP_YUV420A
		p.UseARGB = false
		if err := WebPPictureAllocYUVA(p); err != nil {
			return err
		}
	}
	hasAlpha := make([]byte, 4*len(pics))
	for c0, k := 0, 0; c0 < len(pics); c0, k = c0+chunk, k+1 {
		c1 := min(len(pics), c0+chunk)
		s := k % 4
		for i := c0; i < c1; i++ {
			n := abs(strides[i])*(pics[i].Height-1) + step*pics[i].Width
			if err := ctx.H2D(s, imgs[i].src, srcs[i][:n]); err != nil {
				return err
			}
		}
		if err := ctx.ImportRGBA(s, jobs[c0:c1], opt.Order, flags, hasAlphaDev+cuda.DevicePtr(4*c0)); err != nil {
			return err
		}
		for i := c0; i < c1; i++ {
			p, d := pics[i], imgs[i]
			if err := copyPlane(ctx, s, p.Y, p.YStride, d.y, d.ys, p.Width, p.Height); err != nil {
				return err
			}
			uvw, uvh := (p.Width+1)>>1, (p.Height+1)>>1
			if err := copyPlane(ctx, s, p.U, p.UVStride, d.u, d.uvs, uvw, uvh); err != nil {
				return err
			}
			if err := copyPlane(ctx, s, p.V, p.UVStride, d.v, d.uvs, uvw, uvh); err != nil {
				return err
			}
			if err := copyPlane(ctx, s, p.A, p.AStride, d.a, d.ys, p.Width, p.Height); err != nil {
				return err
			}
		}
		if err := ctx.D2H(s, hasAlpha[4*c0:4*c1], hasAlphaDev+cuda.DevicePtr(4*c0)); err != nil {
			return err
		}
	}
	if err := ctx.Sync(-1); err != nil {
		return err
	}
	for i, p := range pics {
		if hasAlpha[4*i] == 0 { // conversions.go:56-58: opaque image -> no A plane
			p.ColorSpace = colorspace.WEBP_YUV420
			p.A, p.AStride = nil, 0
		}
		d := imgs[i]
		for _, q := range []cuda.DevicePtr{d.src, d.y, d.u, d.v, d.a} {
			ctx.Free(q)
		}
	}
	return nil
}

func abs(v int) int {
	if v < 0 {
		return -v
	}
	return v
}

// copyPlane brings one pitched device plane back into a tightly strided host
// plane row by row (cuda.Context.D2H2D wraps b2_d2h_2d_async).
func copyPlane(ctx *cuda.Context, stream int, dst []uint8, dstStride int, src cuda.DevicePtr, srcPitch, w, h int) error {
	for r := 0; r < h; r++ {
		if err := ctx.D2H(stream, dst[r*dstStride:r*dstStride+w], src+cuda.DevicePtr(r*srcPitch)); err != nil {
			return err
		}
	}
	return nil
}
