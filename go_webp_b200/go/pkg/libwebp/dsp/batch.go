// VP8LInverseTransformBatch: the added batched entry point of part 3 (SURVEY.md 8b), and
// ApplyPaletteBatch, its forward counterpart for colour indexing.
//
// NOTE: untested Go (no toolchain in the build image); the tested twin is
// go_webp_b200/dsp.py (VP8LInverseTransformBatch, ApplyPaletteBatch, _plan_chain).
package dsp

import (
	"github.com/daanv2/go-webp/pkg/cuda"
	"github.com/daanv2/go-webp/pkg/libwebp/webp"
	"github.com/daanv2/go-webp/pkg/vp8"
)

// VP8LImage is one image of the batch: Pixels is what DecodeImageData produced (row length =
// width after colour-index subsampling), Transforms the chain in READ order
// (ReadTransform, pkg/vp8/vp8l_dec.c.go:1221-1269).
type VP8LImage struct {
	Pixels        []uint32
	Width, Height int
	Transforms    []vp8.VP8LTransform
	Out           []uint32 // Width * Height pixels, filled by the call
}

type chainStep struct {
	op            cuda.VP8LOp
	fused         bool // [cross-colour] -> predictor -> [add-green] in one pass
	cross, pred   *vp8.VP8LTransform
	green         bool
	single        *vp8.VP8LTransform
}

// planChain mirrors ApplyInverseTransforms (vp8l_dec.c.go:710-729): last read, first applied; steps
// that have a fused kernel are merged (the order libwebp's encoder writes is ONE pass).
func planChain(tr []vp8.VP8LTransform) []chainStep {
	var steps []chainStep
	at := func(k int) *vp8.VP8LTransform { // k-th transform in application order
		if k < len(tr) {
			return &tr[len(tr)-1-k]
		}
		return nil
	}
	for k := 0; k < len(tr); {
		t, n1, n2 := at(k), at(k+1), at(k+2)
		switch {
		case t.Type == vp8.CROSS_COLOR_TRANSFORM && n1 != nil && n1.Type == vp8.PREDICTOR_TRANSFORM:
			g := n2 != nil && n2.Type == vp8.SUBTRACT_GREEN_TRANSFORM
			steps = append(steps, chainStep{fused: true, cross: t, pred: n1, green: g})
			k += 2
			if g {
				k++
			}
		case t.Type == vp8.PREDICTOR_TRANSFORM && n1 != nil && n1.Type == vp8.SUBTRACT_GREEN_TRANSFORM:
			steps = append(steps, chainStep{fused: true, pred: t, green: true})
			k += 2
		case t.Type == vp8.CROSS_COLOR_TRANSFORM && n1 != nil && n1.Type == vp8.SUBTRACT_GREEN_TRANSFORM:
			steps = append(steps, chainStep{op: cuda.InverseChain, single: t})
			k += 2
		default:
			op := map[vp8.VP8LImageTransformType]cuda.VP8LOp{
				vp8.CROSS_COLOR_TRANSFORM: cuda.ColorInverse, vp8.SUBTRACT_GREEN_TRANSFORM: cuda.AddGreen,
				vp8.COLOR_INDEXING_TRANSFORM: cuda.ColorIndexInverse, vp8.PREDICTOR_TRANSFORM: cuda.PredictorInverse}[t.Type]
			steps = append(steps, chainStep{op: op, single: t})
			k++
		}
	}
	return steps
}

// VP8LInverseTransformBatch applies every image's chain on the GPU.  The pixels cross PCIe twice
// whatever the length of the chain (two ping-pong device buffers per image); outMode MODE_BGRA
// leaves uint32 0xAARRGGBB, MODE_RGBA / MODE_ARGB are the 4-byte orders of EmitRows ->
// VP8LConvertFromBGRA (vp8l_dec.c.go:771), fused into the pass when the last step is the
// predictor chain.  Buffers should be pinned (cuda.Context.Pinned) for asynchronous copies.
func VP8LInverseTransformBatch(ctx *cuda.Context, imgs []VP8LImage, outMode webp.WEBP_CSP_MODE, chunk int) error {
	type plan struct {
		steps      []chainStep
		buf        [2]cuda.DevicePtr
		data       [][2]cuda.DevicePtr // per step: device copies of the transform data (cross / single, predictor)
	}
	plans := make([]plan, len(imgs))
	need := 0
	for i := range imgs {
		plans[i].steps = planChain(imgs[i].Transforms)
		need += 2 * 4 * imgs[i].Width * imgs[i].Height
		for _, t := range imgs[i].Transforms {
			need += (4*len(t.Data) + 255) &^ 255
		}
	}
	arena, err := ctx.Alloc(need + 512*len(imgs))
	if err != nil {
		return err
	}
	defer ctx.Free(arena)
	off := 0
	take := func(n int) cuda.DevicePtr {
		p := arena + cuda.DevicePtr(off)
		off += (n + 255) &^ 255
		return p
	}
	depth := 0
	for i := range imgs {
		px := 4 * imgs[i].Width * imgs[i].Height
		plans[i].buf = [2]cuda.DevicePtr{take(px), take(px)}
		plans[i].data = make([][2]cuda.DevicePtr, len(plans[i].steps))
		for k, s := range plans[i].steps {
			if s.fused {
				if s.cross != nil {
					plans[i].data[k][0] = take(4 * len(s.cross.Data))
				}
				plans[i].data[k][1] = take(4 * len(s.pred.Data))
			} else if len(s.single.Data) > 0 {
				plans[i].data[k][0] = take(4 * len(s.single.Data))
			}
		}
		if len(plans[i].steps) > depth {
			depth = len(plans[i].steps)
		}
	}
	for k := 0; k < depth; k++ {
		// images whose step k is the fused predictor chain go in one call, the others grouped by kernel
		var cj []cuda.ChainJob
		var ch []cuda.ChainHost
		byOp := map[cuda.VP8LOp][]int{}
		for i := range imgs {
			if k >= len(plans[i].steps) {
				continue
			}
			s, p := plans[i].steps[k], &plans[i]
			first, last := k == 0, k == len(p.steps)-1
			if !s.fused {
				byOp[s.op] = append(byOp[s.op], i)
				continue
			}
			j := cuda.ChainJob{Src: p.buf[k&1], Dst: p.buf[(k+1)&1], PredictorData: p.data[k][1],
				Width: s.pred.XSize, Height: s.pred.YSize, PredictorBits: s.pred.Bits, AddGreen: s.green}
			h := cuda.ChainHost{PredictorData: s.pred.Data}
			if s.cross != nil {
				j.ColorData, j.ColorBits = p.data[k][0], s.cross.Bits
				h.ColorData = s.cross.Data
			}
			if first {
				h.Src = imgs[i].Pixels
			}
			if last {
				h.Dst = imgs[i].Out
			}
			cj, ch = append(cj, j), append(ch, h)
		}
		if len(cj) > 0 {
			// the byte order rides on an image's LAST step; here all fused steps of level k are assumed last
			// (libwebp chains); mixed cases convert with VP8LConvertFromBGRA afterwards
			if err := ctx.InverseChainBatchStaged(cj, ch, int(outMode), chunk); err != nil {
				return err
			}
		}
		for op, members := range byOp {
			jobs := make([]cuda.ArgbJob, len(members))
			host := make([]cuda.ArgbHost, len(members))
			for q, i := range members {
				s, p := plans[i].steps[k], &plans[i]
				t := s.single
				jobs[q] = cuda.ArgbJob{Src: p.buf[k&1], Dst: p.buf[(k+1)&1], Data: p.data[k][0], Width: t.XSize, Height: t.YSize, Bits: t.Bits}
				host[q].Data = t.Data
				if k == 0 {
					host[q].Src = imgs[i].Pixels
				}
				if k == len(p.steps)-1 {
					host[q].Dst = imgs[i].Out
				}
			}
			if err := ctx.VP8LBatchStaged(op, jobs, host, chunk); err != nil {
				return err
			}
		}
	}
	return nil
}
