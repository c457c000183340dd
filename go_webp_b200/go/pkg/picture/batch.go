// Batched, GPU-backed entry points added to pkg/picture (SURVEY.md 8b), and the
// device-resident form of Picture.  The existing signatures (ImportYUVAFromRGBA,
// PictureARGBToYUVA, ...) keep their parameter lists and forward here with a batch of one.
//
// NOTE: untested Go (no toolchain in the build image); the tested twin of this
// file is go_webp_b200/picture.py.
package picture

import (
	"github.com/daanv2/go-webp/pkg/color/colorspace"
	"github.com/daanv2/go-webp/pkg/cuda"
)

// Residency is what a Picture gains (picture.go:100-103: memory_ / memory_argb_ become these):
// the planes live in a device block with a pinned-host mirror; Y, U, V, A are slices of the
// mirror.  HostValid / DevValid say which side holds the current data, so chained batched
// calls (import -> WebPCleanupTransparentArea -> ...) never touch the host in between.
type Residency struct {
	Block               *cuda.Block
	Off                 int // offset of the picture inside the block
	YOff, UOff, VOff    int // plane offsets (same on both sides)
	AOff                int
	YStride, UVStride   int // 16-byte aligned (allocations.go:81: "we could align the y/u/v planes and adjust stride")
	Bytes               int
	HostValid, DevValid bool
}

func al16(n int) int { return (n + 15) &^ 15 }

// yuvaBytes is the size of one resident YUVA picture: Y | U | V | A, strides rounded up to 16.
func yuvaBytes(w, h int) int {
	return 2*al16(w)*h + 2*al16((w+1)>>1)*((h+1)>>1)
}

// AllocResidentBatch is WebPPictureAllocYUVA for a batch: ONE device block and ONE pinned block
// for all pictures that do not already own memory of the right geometry.
func AllocResidentBatch(ctx *cuda.Context, pics []*Picture) error {
	need := 0
	var todo []*Picture
	for _, p := range pics {
		if err := WebPValidatePicture(p); err != nil {
			return p.SetEncodingError(err)
		}
		if r := p.Resident; r != nil && r.Bytes == yuvaBytes(p.Width, p.Height) && r.YStride == al16(p.Width) {
			continue
		}
		todo = append(todo, p)
		need += yuvaBytes(p.Width, p.Height)
	}
	if len(todo) == 0 {
		return nil
	}
	block, err := ctx.NewBlock(need)
	if err != nil {
		for _, p := range todo {
			p.SetEncodingError(ENC_ERROR_OUT_OF_MEMORY)
		}
		return ENC_ERROR_OUT_OF_MEMORY
	}
	off := 0
	for _, p := range todo {
		w, h := p.Width, p.Height
		uvh := (h + 1) >> 1
		r := &Residency{Block: block, Off: off, YStride: al16(w), UVStride: al16((w + 1) >> 1), Bytes: yuvaBytes(w, h)}
		r.YOff = off
		r.UOff = r.YOff + r.YStride*h
		r.VOff = r.UOff + r.UVStride*uvh
		r.AOff = r.VOff + r.UVStride*uvh
		p.Resident = r
		p.Y = block.Host[r.YOff : r.YOff+r.YStride*h]
		p.U = block.Host[r.UOff : r.UOff+r.UVStride*uvh]
		p.V = block.Host[r.VOff : r.VOff+r.UVStride*uvh]
		p.A = block.Host[r.AOff : r.AOff+r.YStride*h]
		p.YStride, p.UVStride, p.AStride = r.YStride, r.UVStride, r.YStride
		off += r.Bytes
	}
	return nil
}

// ImportOptions selects the source layout of ImportRGBABatch.
type ImportOptions struct {
	Order       int     // cuda.OrderRGBA, OrderBGRA, OrderRGB, OrderBGR
	ImportAlpha bool    // picture_csp_enc.c.go:123 import_alpha
	Chunk       int     // images per pipelined chunk (default 16)
	Dithering   float32 // conversions.go:122-165; 0 = none
	DivTable    bool    // libwebp-binary inverse-alpha table instead of the exact division of yuv.go:231
	KeepOnGPU   bool    // leave the planes on the device for the next batched call (no D2H)
}

// ImportRGBABatch imports srcs[i] (row stride strides[i]; pinned memory from cuda.Context.Pinned
// copies asynchronously) into pics[i] as YUV420(A): per image exactly enc.Import ->
// picture.ImportYUVAFromRGBA (picture_csp_enc.c.go:123-137, conversions.go:50-165).  The pictures
// become device-resident; the whole batch is ONE cgo call (b2_import_batch_staged) that pipelines
// chunks over the context's streams and brings the alpha plane back only for pictures that have
// alpha (conversions.go:56-58).
func ImportRGBABatch(ctx *cuda.Context, pics []*Picture, srcs [][]uint8, strides []int, opt ImportOptions) error {
	if len(pics) != len(srcs) || len(pics) != len(strides) {
		return ENC_ERROR_INVALID_CONFIGURATION
	}
	if len(pics) == 0 {
		return nil
	}
	step, minStep := 4, 3
	if opt.Order == cuda.OrderRGB || opt.Order == cuda.OrderBGR {
		step = 3
	}
	if opt.ImportAlpha {
		minStep = 4
	}
	flags := 0
	if opt.ImportAlpha {
		flags |= cuda.ImportAlpha
	}
	if opt.DivTable {
		flags |= cuda.ImportDivTable
	}
	srcBytes := 0
	for i, p := range pics {
		if p == nil || srcs[i] == nil {
			return ENC_ERROR_nil_PARAMETER
		}
		if err := WebPValidatePicture(p); err != nil {
			return p.SetEncodingError(err)
		}
		if abs(strides[i]) < minStep*p.Width || abs(strides[i]) < step*p.Width {
			return ENC_ERROR_INVALID_CONFIGURATION // picture_csp_enc.c.go:132
		}
		srcBytes += al16(abs(strides[i])*(p.Height-1) + step*p.Width)
	}
	if err := AllocResidentBatch(ctx, pics); err != nil {
		return err
	}
	// device slots for the sources: one allocation per call (the Python twin keeps a grow-only arena)
	srcDev, err := ctx.Alloc(srcBytes)
	if err != nil {
		return ENC_ERROR_OUT_OF_MEMORY
	}
	defer ctx.Free(srcDev)
	jobs := make([]cuda.ImportJob, len(pics))
	host := make([]cuda.ImportHost, len(pics))
	off := 0
	for i, p := range pics {
		r := p.Resident
		span := abs(strides[i])*(p.Height-1) + step*p.Width
		back := 0
		if strides[i] < 0 { // pixel (0,0) is in the last row of memory (buffer_dec.c.go:161-184)
			back = abs(strides[i]) * (p.Height - 1)
		}
		d := r.Block.Dev
		jobs[i] = cuda.ImportJob{Src: srcDev + cuda.DevicePtr(off+back),
			Y: d + cuda.DevicePtr(r.YOff), U: d + cuda.DevicePtr(r.UOff), V: d + cuda.DevicePtr(r.VOff), A: d + cuda.DevicePtr(r.AOff),
			SrcStride: strides[i], YStride: r.YStride, UVStride: r.UVStride, AStride: r.YStride, Width: p.Width, Height: p.Height}
		host[i] = cuda.ImportHost{Src: srcs[i][back:span]}
		if !opt.KeepOnGPU {
			host[i].Y, host[i].U, host[i].V, host[i].A = p.Y, p.U, p.V, r.Block.Host[r.AOff:r.AOff+r.YStride*p.Height]
			host[i].YStride, host[i].UVStride, host[i].AStride = r.YStride, r.UVStride, r.YStride
		}
		off += al16(span)
	}
	hasAlpha, err := ctx.ImportBatchStaged(jobs, host, opt.Order, flags, opt.Chunk, opt.Dithering)
	if err != nil {
		for _, p := range pics {
			p.SetEncodingError(ENC_ERROR_OUT_OF_MEMORY)
		}
		return err
	}
	for i, p := range pics {
		r := p.Resident
		r.DevValid, r.HostValid = true, !opt.KeepOnGPU
		p.UseARGB = false // conversions.go:59
		if hasAlpha[i] {  // conversions.go:56-58
			p.ColorSpace = colorspace.WEBP_YUV420A
			p.A, p.AStride = r.Block.Host[r.AOff:r.AOff+r.YStride*p.Height], r.YStride
		} else {
			p.ColorSpace = colorspace.WEBP_YUV420
			p.A, p.AStride = nil, 0
		}
	}
	return nil
}

// Download brings a resident picture whose host mirror is stale back to the host (one copy).
func Download(ctx *cuda.Context, p *Picture) error {
	r := p.Resident
	if r == nil || r.HostValid || !r.DevValid {
		return nil
	}
	if err := ctx.D2H(0, r.Block.Host[r.Off:r.Off+r.Bytes], r.Block.Dev+cuda.DevicePtr(r.Off)); err != nil {
		return err
	}
	if err := ctx.Sync(0); err != nil {
		return err
	}
	r.HostValid = true
	return nil
}

func abs(v int) int {
	if v < 0 {
		return -v
	}
	return v
}
