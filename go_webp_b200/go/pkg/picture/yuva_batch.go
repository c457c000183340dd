// YUVAToRGBABatch: the added batched entry point of part 2 (SURVEY.md 8b).
//
// NOTE: untested Go (no toolchain in the build image); the tested twin is
// go_webp_b200/picture.py (YUVAToRGBABatch, YUVAToARGBBatch).
package picture

import (
	"github.com/daanv2/go-webp/pkg/color/colorspace"
	"github.com/daanv2/go-webp/pkg/cuda"
	"github.com/daanv2/go-webp/pkg/libwebp/webp"
)

var modeBpp = [...]int{3, 4, 3, 4, 4, 2, 2, 4, 4, 4, 2} // kModeBpp, pkg/libwebp/decoder/buffer_dec.c.go:28-32

// YUVAToRGBABatch fancy-upsamples pics[i].Y/U/V(/A) into dsts[i] (row stride dstStrides[i]) in
// WEBP_CSP_MODE mode: per image the schedule of WebPPictureYUVAToARGB
// (pkg/libwebp/enc/picture_csp_enc.c.go:62-118) with WebPUpsamplers[mode]
// (pkg/libwebp/dsp/upsampling.c.go:275-289) and EmitAlphaRGB fused.  Device-resident pictures
// (ImportRGBABatch, the picture tools) are read where they lie: only the output crosses PCIe.
// dsts should be pinned (cuda.Context.Pinned).
func YUVAToRGBABatch(ctx *cuda.Context, pics []*Picture, dsts [][]uint8, dstStrides []int, mode webp.WEBP_CSP_MODE, chunk int) error {
	if len(pics) != len(dsts) || len(pics) != len(dstStrides) {
		return ENC_ERROR_INVALID_CONFIGURATION
	}
	if len(pics) == 0 {
		return nil
	}
	bpp := modeBpp[int(mode)]
	need := 0
	for i, p := range pics {
		if p == nil || p.Y == nil || p.U == nil || p.V == nil || dsts[i] == nil {
			return ENC_ERROR_nil_PARAMETER
		}
		if p.ColorSpace&colorspace.WEBP_CSP_UV_MASK != colorspace.WEBP_YUV420 {
			return p.SetEncodingError(ENC_ERROR_INVALID_CONFIGURATION)
		}
		need += al16(al16(p.Width*bpp) * p.Height)
		if p.Resident == nil || !p.Resident.DevValid {
			need += yuvaBytes(p.Width, p.Height)
		}
	}
	scratch, err := ctx.Alloc(need)
	if err != nil {
		return ENC_ERROR_OUT_OF_MEMORY
	}
	defer ctx.Free(scratch)
	jobs := make([]cuda.UpsampleJob, len(pics))
	host := make([]cuda.UpsampleHost, len(pics))
	off := 0
	for i, p := range pics {
		w, h := p.Width, p.Height
		uvh := (h + 1) >> 1
		hasA := p.ColorSpace&colorspace.WEBP_CSP_ALPHA_BIT != 0 && p.A != nil && (bpp == 4 || mode == webp.MODE_RGBA_4444 || mode == webp.MODE_rgbA_4444)
		j := cuda.UpsampleJob{Width: w, Height: h, DstStride: al16(w * bpp)}
		if r := p.Resident; r != nil && r.DevValid { // planes already on the device: no upload
			d := r.Block.Dev
			j.Y, j.U, j.V = d+cuda.DevicePtr(r.YOff), d+cuda.DevicePtr(r.UOff), d+cuda.DevicePtr(r.VOff)
			j.YStride, j.UVStride, j.AStride = r.YStride, r.UVStride, r.YStride
			if hasA {
				j.A = d + cuda.DevicePtr(r.AOff)
			}
		} else { // user-owned planes: staged into scratch by the same call
			ys, uvs := al16(w), al16((w+1)>>1)
			j.Y = scratch + cuda.DevicePtr(off)
			j.U = j.Y + cuda.DevicePtr(ys*h)
			j.V = j.U + cuda.DevicePtr(uvs*uvh)
			j.YStride, j.UVStride, j.AStride = ys, uvs, ys
			host[i].Y, host[i].U, host[i].V = p.Y, p.U, p.V
			host[i].YStride, host[i].UVStride = p.YStride, p.UVStride
			if hasA {
				j.A = j.V + cuda.DevicePtr(uvs*uvh)
				host[i].A, host[i].AStride = p.A, p.AStride
			}
			off += yuvaBytes(w, h)
		}
		j.Dst = scratch + cuda.DevicePtr(off)
		off += al16(j.DstStride * h)
		host[i].Dst, host[i].DstStride = dsts[i], dstStrides[i]
		jobs[i] = j
	}
	if err := ctx.UpsampleBatchStaged(jobs, host, int(mode), chunk); err != nil {
		for _, p := range pics {
			p.SetEncodingError(ENC_ERROR_OUT_OF_MEMORY)
		}
		return err
	}
	return nil
}
