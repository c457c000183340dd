// Rebinding of the reference's dsp function variables to the CUDA library: the drop-in
// boundary of SURVEY.md 8b.  The reference binds the encoder-side VP8L functions at
// pkg/libwebp/dsp/lossless_enc.go:488-519, the upsampler table at upsampling.c.go:275-289,
// and calls VP8LDspInit() (pkg/vp8/vp8l_dec.c.go:1301, pkg/libwebp/enc/picture_csp_enc.c.go:140)
// without defining it; this file supplies it.  There is no CPU fallback: if the library or a
// device is missing, Init returns the error and the function variables stay unbound.
//
// The per-row adapters keep the reference's signatures (lossless.h.go:33-45, dsp.h.go:144);
// they move one row through the GPU and are slow by construction -- throughput comes from the
// batched entry points (picture.ImportRGBABatch, picture.YUVAToRGBABatch,
// dsp.VP8LInverseTransformBatch, dsp.ApplyPaletteBatch).
//
// NOTE: untested Go (no toolchain in the build image); the tested twin is go_webp_b200/dsp.py.
package dsp

import (
	"sync"

	"github.com/daanv2/go-webp/pkg/cuda"
	"github.com/daanv2/go-webp/pkg/libwebp/webp"
)

var (
	cudaOnce sync.Once
	cudaCtx  *cuda.Context
	cudaErr  error
)

// Context returns the process-wide context of device 0 the row adapters use (batched entry
// points take an explicit context: one per GPU and OS thread).
func Context() (*cuda.Context, error) {
	cudaOnce.Do(func() { cudaCtx, cudaErr = cuda.NewContext(0, 4) })
	return cudaCtx, cudaErr
}

// rowOp runs one flat per-pixel transform on n pixels through the staged entry point.
func rowOp(op cuda.VP8LOp, src []uint32, n int, dst []uint32, data []uint32, bits int) {
	ctx, err := Context()
	if err != nil {
		panic(err) // the reference's row functions have no error return; no CPU fallback by design
	}
	buf, err := ctx.Alloc(8*n + 4*len(data) + 512)
	if err != nil {
		panic(err)
	}
	defer ctx.Free(buf)
	j := cuda.ArgbJob{Src: buf, Dst: buf + cuda.DevicePtr((4*n+255)&^255), Width: n, Height: 1, Bits: bits}
	if len(data) > 0 {
		j.Data = j.Dst + cuda.DevicePtr((4*n+255)&^255)
	}
	h := cuda.ArgbHost{Src: src[:n], Dst: dst[:n], Data: data}
	if err := ctx.VP8LBatchStaged(op, []cuda.ArgbJob{j}, []cuda.ArgbHost{h}, 1); err != nil {
		panic(err)
	}
}

// VP8LDspInit binds the decoder- and encoder-side VP8L function variables.
func VP8LDspInit() {
	// VP8LProcessDecBlueAndRedFunc (lossless.h.go:35; used at lossless.go:517)
	VP8LAddGreenToBlueAndRed = func(src []uint32, numPixels int, dst []uint32) {
		rowOp(cuda.AddGreen, src, numPixels, dst, nil, 0)
	}
	// VP8LTransformColorInverseFunc (lossless.h.go:36; used at lossless.go:396): one tile segment
	VP8LTransformColorInverse = func(m *VP8LMultipliers, src []uint32, numPixels int, dst []uint32) {
		code := []uint32{0xff000000 | uint32(m.red_to_blue)<<16 | uint32(m.green_to_blue)<<8 | uint32(m.green_to_red)}
		for x := 0; x < numPixels; x += 512 { // bits = 9: one tile covers 512 pixels of a one-row image
			n := min(512, numPixels-x)
			rowOp(cuda.ColorInverse, src[x:], n, dst[x:], code, 9)
		}
	}
	// VP8LMapARGBFunc (lossless.h.go:40; used at lossless.go:457)
	VP8LMapColor32b = func(src []uint32, colorMap []uint32, dst []uint32, yStart, yEnd, width int) {
		pal := make([]uint32, 256)
		copy(pal, colorMap)
		rowOp(cuda.ColorIndexInverse, src, (yEnd-yStart)*width, dst, pal, 0)
	}
	// encoder side, lossless_enc.go:488-519
	VP8LSubtractGreenFromBlueAndRed = func(argb []uint32, numPixels int) {
		rowOp(cuda.SubGreen, argb, numPixels, argb, nil, 0)
	}
	VP8LTransformColor = func(m *VP8LMultipliers, data []uint32, numPixels int) {
		code := []uint32{0xff000000 | uint32(m.red_to_blue)<<16 | uint32(m.green_to_blue)<<8 | uint32(m.green_to_red)}
		for x := 0; x < numPixels; x += 512 {
			n := min(512, numPixels-x)
			rowOp(cuda.ColorForward, data[x:], n, data[x:], code, 9)
		}
	}
}

// WebPInitUpsamplers binds WebPUpsamplers[mode] (upsampling.c.go:275-289): every line-pair call
// becomes a two-row upsample job; premultiplied modes alias the plain ones as the reference does.
func WebPInitUpsamplers() {
	for m := webp.MODE_RGB; m < webp.MODE_YUV; m++ {
		mode := m
		WebPUpsamplers[mode] = func(topY, bottomY, topU, topV, curU, curV, topDst, bottomDst []uint8, length int) {
			linePair(int(mode), topY, bottomY, topU, topV, curU, curV, topDst, bottomDst, length)
		}
	}
}

// linePair: rows (1, 2) of a 3-row picture whose chroma rows are (top, cur) -- or a 1-row picture for
// the mirrored first / last row (bottomY == nil) -- through UpsampleBatchStaged.
func linePair(mode int, topY, bottomY, topU, topV, curU, curV, topDst, bottomDst []uint8, n int) {
	ctx, err := Context()
	if err != nil {
		panic(err)
	}
	uw := (n + 1) >> 1
	bpp := [...]int{3, 4, 3, 4, 4, 2, 2, 4, 4, 4, 2}[mode]
	rows := 3
	if bottomY == nil {
		rows = 1
	}
	ys, uvs, ds := (n+15)&^15, (uw+15)&^15, (n*bpp+15)&^15
	host := make([]byte, rows*n+4*uw+rows*n*bpp)
	y, u, v := host[:rows*n], host[rows*n:rows*n+2*uw], host[rows*n+2*uw:rows*n+4*uw]
	out := host[rows*n+4*uw:]
	if rows == 1 {
		copy(y, topY[:n])
		copy(u, topU[:uw])
		copy(v, topV[:uw])
	} else {
		copy(y[n:], topY[:n])
		copy(y[2*n:], bottomY[:n])
		copy(u, topU[:uw])
		copy(u[uw:], curU[:uw])
		copy(v, topV[:uw])
		copy(v[uw:], curV[:uw])
	}
	dev, err := ctx.Alloc(rows*ys + 4*uvs + rows*ds + 1024)
	if err != nil {
		panic(err)
	}
	defer ctx.Free(dev)
	j := cuda.UpsampleJob{Y: dev, YStride: ys, UVStride: uvs, DstStride: ds, Width: n, Height: rows}
	j.U = j.Y + cuda.DevicePtr((rows*ys+255)&^255)
	j.V = j.U + cuda.DevicePtr((2*uvs+255)&^255)
	j.Dst = j.V + cuda.DevicePtr((2*uvs+255)&^255)
	h := cuda.UpsampleHost{Y: y, U: u, V: v, Dst: out, YStride: n, UVStride: uw, DstStride: n * bpp}
	if err := ctx.UpsampleBatchStaged([]cuda.UpsampleJob{j}, []cuda.UpsampleHost{h}, mode, 1); err != nil {
		panic(err)
	}
	if rows == 1 {
		copy(topDst, out[:n*bpp])
		return
	}
	copy(topDst, out[n*bpp:2*n*bpp])
	copy(bottomDst, out[2*n*bpp:3*n*bpp])
}
