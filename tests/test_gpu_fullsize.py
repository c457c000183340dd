"""BASELINE.json configs at full size on the GPU (device-resident synthetic
inputs, SURVEY.md 8d generator), checked against the multi-threaded CPU oracle on
every image where the oracle finishes in seconds, plus size-independent
properties (inverse o forward == identity, idempotent re-run)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from go_webp_b200 import _abi, device, pipeline
    if device.device_count() < 1:
        pytest.fail("no CUDA device")
    return dict(abi=_abi, pipeline=pipeline, ctx=device.Context(0, 4))


def _threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


@pytest.mark.parametrize("alpha_mode", [1, 2, 3])
def test_config1_256_frames_1080p_import(env, alpha_mode):
    """configs[1]: 256 x 1920x1080 RGBA -> YUV420A on one GPU, every frame and
    every plane bit-exact vs the oracle; opaque / 30%-translucent / binary alpha."""
    ctx, PL = env["ctx"], env["pipeline"]
    n, w, h = 256, 1920, 1080
    seed = 0xB2000001
    p = PL.ImportPipeline(ctx, n, w, h, host=False)
    p.fill_synthetic(seed, 0, alpha_mode)
    p.run_device()
    ctx.sync()
    L = oracle.lib()
    uvw, uvh = w // 2, h // 2
    batch = 32
    vp = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    for b0 in range(0, n, batch):
        src = np.stack([oracle.synth_u32(seed, b0 + i, 0, w, h, alpha_mode) for i in range(batch)])
        y = np.empty((batch, h, w), np.uint8)
        u = np.empty((batch, uvh, uvw), np.uint8)
        v = np.empty((batch, uvh, uvw), np.uint8)
        a = np.empty((batch, h, w), np.uint8)
        ha = np.zeros(batch, np.int32)
        assert L.orc_import_batch_mt(vp(src), batch, w, h, vp(y), vp(u), vp(v), vp(a), vp(ha), 0, _threads()) == 0
        gy = ctx.d2h(p.d_y.ptr + b0 * p.y_img, batch * p.y_img).reshape(batch, p.y_img)[:, :p.ys * h].reshape(batch, h, p.ys)[:, :, :w]
        gu = ctx.d2h(p.d_u.ptr + b0 * p.uv_img, batch * p.uv_img).reshape(batch, p.uv_img)[:, :p.uvs * uvh].reshape(batch, uvh, p.uvs)[:, :, :uvw]
        gv = ctx.d2h(p.d_v.ptr + b0 * p.uv_img, batch * p.uv_img).reshape(batch, p.uv_img)[:, :p.uvs * uvh].reshape(batch, uvh, p.uvs)[:, :, :uvw]
        ga = ctx.d2h(p.d_a.ptr + b0 * p.y_img, batch * p.y_img).reshape(batch, p.y_img)[:, :p.ys * h].reshape(batch, h, p.ys)[:, :, :w]
        gf = ctx.d2h(p.d_flags.ptr + 4 * b0, 4 * batch).view(np.int32)
        assert (gy == y).all() and (gu == u).all() and (gv == v).all() and (ga == a).all()
        assert (gf == ha).all() and (ha == (0 if alpha_mode == 1 else 1)).all()
    p.close()


def test_config0_single_1080p_round_trip(env):
    """configs[0]: one 1920x1080 synthetic RGBA image, RGBA -> YUV420 import then YUV420 -> RGBA fancy
    upsample, the reference path on the CPU (oracle) vs the same two calls through the drop-in API;
    opaque alpha and uniform alpha."""
    from go_webp_b200 import picture as P
    abi = env["abi"]
    for alpha_mode in (abi.SYNTH_OPAQUE, abi.SYNTH_RAW):
        src = oracle.synth_u32(0xB2000000, 0, 0, 1920, 1080, alpha_mode).view(np.uint8).reshape(1080, 1920 * 4)
        pic = P.Picture(Width=1920, Height=1080)
        assert P.WebPPictureImportRGBA(pic, src, 4 * 1920) == 1
        y, u, v, a, ha = oracle.import_yuva(src, 1920, 1080, "RGBA")
        assert (pic.Y == y).all() and (pic.U == u).all() and (pic.V == v).all()
        assert ha == (alpha_mode == abi.SYNTH_RAW) and ((pic.A == a).all() if ha else pic.A is None)
        out = P.YUVAToRGBABatch([pic], abi.MODE_RGBA)[0]
        assert (out == oracle.yuva_to_rgba(y, u, v, a if ha else None, oracle.MODE_RGBA)).all()


def test_config2_vp8l_chain_4096(env):
    """configs[2]: VP8L inverse chain on 4096x4096 ARGB images: cross-colour with
    32x32-tile multipliers (bits = 5) + add-green, fused and unfused; 8-bit and
    2-bit colour indexing.  Oracle on 4 of the images, identity property on all."""
    ctx, PL, abi = env["ctx"], env["pipeline"], env["abi"]
    n, w, h, bits = 16, 4096, 4096, 5
    chain = PL.ArgbPipeline(ctx, n, w, h, "inverse_chain", bits)
    chain.fill_synthetic(0xB2000003)
    chain.run_device()
    ctx.sync()
    for i in (0, 5, 10, 15):
        src, codes, dst = chain.device_image(i)
        codes2 = codes.reshape(128, 128)
        exp = oracle.add_green(oracle.color_space_inverse(src, codes2, bits))
        assert (dst == exp).all()
    # unfused = cross-colour^-1 then add-green as two launches, in place on dst
    jobs_cc = [abi.ArgbJob(j.src, j.dst, j.data, w, h, bits, 0) for j in chain.jobs]
    first = ctx.d2h(chain.d_dst.ptr, 4 * w * h).copy()
    ctx.vp8l("color_inverse", jobs_cc)
    ctx.vp8l("add_green", [abi.ArgbJob(j.dst, j.dst, None, w, h, 0, 0) for j in chain.jobs])
    ctx.sync()
    assert (ctx.d2h(chain.d_dst.ptr, 4 * w * h) == first).all()
    # forward chain on the result restores the source for every image (identity)
    ctx.vp8l("forward_chain", [abi.ArgbJob(j.dst, j.dst, j.data, w, h, bits, 0) for j in chain.jobs])
    ctx.sync()
    for i in range(n):
        a = ctx.d2h(chain.d_dst.ptr + i * chain.dst_img, 4 * w * h)
        b = ctx.d2h(chain.d_src.ptr + i * chain.src_img, 4 * w * h)
        assert (a == b).all()
    chain.close()
    for cbits in (0, 2):
        ci = PL.ArgbPipeline(ctx, 8, w, h, "color_index_inverse", cbits)
        ci.fill_synthetic(0xB2000003)
        ci.run_device()
        ctx.sync()
        for i in (0, 7):
            src, pal, dst = ci.device_image(i)
            assert (dst == oracle.color_index_inverse(src, pal, cbits, w)).all()
        ci.close()


def test_config3_4k_upsample(env):
    """configs[3] (one rank's share at reduced count): 3840x2160 YUV420 -> RGBA,
    oracle on EVERY one of the 48 frames, re-run idempotence on all."""
    ctx, PL, abi = env["ctx"], env["pipeline"], env["abi"]
    n, w, h = 48, 3840, 2160
    for mode, with_a in ((abi.MODE_RGBA, False), (abi.MODE_BGRA, True)):
        up = PL.UpsamplePipeline(ctx, n, w, h, mode, with_alpha=with_a, host=False)
        up.fill_synthetic(0xB2000004)
        up.run_device()
        ctx.sync()
        first = ctx.d2h(up.d_dst.ptr, up.d_img * 4).copy()
        for i in range(n):
            y, u, v, a, d = up.device_image(i)
            assert (d == oracle.yuva_to_rgba(y, u, v, a, mode)).all(), (mode, i)
        up.run_device()
        ctx.sync()
        assert (ctx.d2h(up.d_dst.ptr, up.d_img * 4) == first).all()
        up.close()


def test_config4_mixed_sizes_full_pipeline(env):
    """configs[4] (one rank's share): mixed sizes 64..8192 with odd widths and
    heights through import -> upsample and the ARGB chain, every image vs oracle
    (the big ones on a row band to keep the CPU side short)."""
    from go_webp_b200 import dsp, picture
    rng = np.random.default_rng(0xB2000005)
    sizes = [(64, 64), (65, 8191), (8191, 65), (8192, 512), (1001, 777), (4095, 33), (300, 301)]
    pics, srcs = [], []
    for (w, h) in sizes:
        src = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        src[..., 3][rng.random((h, w)) < 0.8] = 255
        srcs.append(src)
        pics.append(picture.Picture(Width=w, Height=h))
    assert picture.ImportRGBABatch(pics, srcs, [4 * w for (w, h) in sizes], "RGBA") == 1
    outs = picture.YUVAToRGBABatch(pics, env["abi"].MODE_BGRA)
    for pic, src, out, (w, h) in zip(pics, srcs, outs, sizes):
        y, u, v, a, ha = oracle.import_yuva(src, w, h, "RGBA")
        assert (pic.Y == y).all() and (pic.U == u).all() and (pic.V == v).all() and ha
        assert (pic.A == a).all()
        assert (out == oracle.yuva_to_rgba(y, u, v, a, oracle.MODE_BGRA)).all()
    argbs = [o.view(np.uint32).reshape(h, w) for o, (w, h) in zip(outs, sizes)]
    chains = []
    for (w, h) in sizes:
        codes = rng.integers(0, 2**24, (dsp.VP8LSubSampleSize(h, 4), dsp.VP8LSubSampleSize(w, 4)),
                             dtype=np.uint32) | 0xff000000
        chains.append([dsp.VP8LTransform(dsp.SUBTRACT_GREEN_TRANSFORM, 0, w, h),
                       dsp.VP8LTransform(dsp.CROSS_COLOR_TRANSFORM, 4, w, h, codes)])
    res = dsp.VP8LInverseTransformBatch([dsp.VP8LImage(a, w, h) for a, (w, h) in zip(argbs, sizes)], chains)
    for r, a, ch in zip(res, argbs, chains):
        assert (r == oracle.add_green(oracle.color_space_inverse(a, ch[1].data, 4))).all()


@pytest.mark.parametrize("size", [(16383, 5), (6, 16383)])
def test_maximum_dimension(env, size):
    """WEBP_MAX_DIMENSION = 16383 (pkg/libwebp/webp/encode.h.go:95) in either direction,
    odd on purpose: import, fancy upsample, VP8L chain and predictor through the public
    mirror of the reference API, against the oracle."""
    from go_webp_b200 import _abi, dsp, picture
    w, h = size
    rng = np.random.default_rng(w + h)
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    rgba[..., 3][rng.random((h, w)) < 0.8] = 255
    pic = picture.Picture(Width=w, Height=h)
    assert picture.WebPPictureImportRGBA(pic, rgba, 4 * w) == 1
    y, u, v, a, ha = oracle.import_yuva(rgba, w, h, "RGBA")
    assert (pic.Y == y).all() and (pic.U == u).all() and (pic.V == v).all() and (pic.A == a).all()
    out = picture.YUVAToRGBABatch([pic], _abi.MODE_RGBA)[0]
    assert (out == oracle.yuva_to_rgba(y, u, v, a, oracle.MODE_RGBA).reshape(h, 4 * w)).all()
    bits = 4
    argb = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
    codes = rng.integers(0, 2**24, (dsp.VP8LSubSampleSize(h, bits), dsp.VP8LSubSampleSize(w, bits)), dtype=np.uint32)
    modes = rng.integers(0, 16, codes.shape, dtype=np.uint32) << 8
    chain = [dsp.VP8LTransform(dsp.PREDICTOR_TRANSFORM, bits, w, h, modes),
             dsp.VP8LTransform(dsp.CROSS_COLOR_TRANSFORM, bits, w, h, codes),
             dsp.VP8LTransform(dsp.SUBTRACT_GREEN_TRANSFORM, 0, w, h)]
    got = dsp.VP8LInverseTransformBatch([dsp.VP8LImage(argb, w, h)], [chain])[0]
    exp = oracle.predictor_inverse(oracle.color_space_inverse(oracle.add_green(argb), codes, bits), modes, bits)
    assert (got.reshape(h, w) == exp).all()
    pal = rng.integers(0, 2**32, 4, dtype=np.uint32)
    sw = dsp.VP8LSubSampleSize(w, 2)
    packed = rng.integers(0, 2**32, (h, sw), dtype=np.uint32)
    t = dsp.VP8LTransform(dsp.COLOR_INDEXING_TRANSFORM, 2, w, h, pal)
    assert (dsp.VP8LInverseTransform(t, 0, h, packed) == oracle.color_index_inverse(packed, pal, 2, w)).all()


@pytest.mark.parametrize("size", [(5, 5), (6, 2), (7, 3), (13, 7), (130, 5), (1917, 4), (1918, 3), (1919, 5)])
@pytest.mark.parametrize("alpha_mode", [0, 2, 3])
def test_tuned_import_tight_rows_any_width(env, size, alpha_mode):
    """Tight source rows (stride 4*w, only 4-byte aligned when w % 4 != 0) with padded
    planes take the tuned kernel's 32-bit-load path and its partial last group: the odd
    last column must behave as the reference's SUM2 / SUM2ALPHA forms (yuv.go:270-277)."""
    ctx, PL = env["ctx"], env["pipeline"]
    w, h = size
    n, seed = 3, 0xB2000005
    p = PL.ImportPipeline(ctx, n, w, h, host=False)
    p.fill_synthetic(seed, 0, alpha_mode)
    p.run_device()
    ctx.sync()
    for i in range(n):
        src = oracle.synth_u32(seed, i, 0, w, h, alpha_mode).view(np.uint8)
        ey, eu, ev, ea, eha = oracle.import_yuva(src, w, h, "RGBA")
        gy, gu, gv, ga, gha = p.device_planes(i)
        assert (gy == ey).all() and (gu == eu).all() and (gv == ev).all() and (ga == ea).all() and gha == eha, (size, i)
    p.close()


def test_mixed_pipeline_device_resident(env):
    """MixedPipeline (the configs[4] throughput path: three flat launches over images of
    different sizes): every image of a small mixed batch vs the oracle, inputs from the
    counter-based generator on both sides."""
    from go_webp_b200.pipeline import MixedPipeline, mixed_sizes
    abi, ctx = env["abi"], env["ctx"]
    sizes = [(64, 64), (65, 131), (130, 67), (1001, 77), (257, 300), (96, 33)]
    seed = 0xB2000005
    p = MixedPipeline(ctx, sizes, mode=abi.MODE_BGRA, bits=5)
    p.fill_synthetic(seed)
    p.run_device()
    ctx.sync()
    sub = lambda v, b: (v + (1 << b) - 1) >> b  # noqa: E731
    for i, (w, h, uvw, uvh, ys, uvs, ds) in enumerate(p.geo):
        src = oracle.synth_u32(seed, i, 0, w, h, abi.SYNTH_SPARSE)
        y, u, v, a, ha = oracle.import_yuva(src.view(np.uint8).reshape(h, w, 4), w, h, "RGBA")
        gy = ctx.d2h(p.d_y.ptr + p.off["y"][i], ys * h).reshape(h, ys)[:, :w]
        ga = ctx.d2h(p.d_a.ptr + p.off["y"][i], ys * h).reshape(h, ys)[:, :w]
        gu = ctx.d2h(p.d_u.ptr + p.off["uv"][i], uvs * uvh).reshape(uvh, uvs)[:, :uvw]
        gv = ctx.d2h(p.d_v.ptr + p.off["uv"][i], uvs * uvh).reshape(uvh, uvs)[:, :uvw]
        assert (gy == y).all() and (gu == u).all() and (gv == v).all() and (ga == a).all()
        gd = ctx.d2h(p.d_dst.ptr + p.off["dst"][i], ds * h).reshape(h, ds)[:, :4 * w]
        assert (gd == oracle.yuva_to_rgba(y, u, v, a, oracle.MODE_BGRA).reshape(h, 4 * w)).all()
        words = sub(w, 5) * sub(h, 5)
        codes = oracle.synth_u32(seed, i, 6, words, 1).reshape(sub(h, 5), sub(w, 5))
        go = ctx.d2h(p.d_out.ptr + p.off["src"][i], 4 * w * h).view(np.uint32).reshape(h, w)
        assert (go == oracle.add_green(oracle.color_space_inverse(src, codes, 5))).all()
    p.close()


@pytest.mark.parametrize("size", [(16383, 9), (10, 16383)])
def test_maximum_dimension_next_rows(env, size):
    """WEBP_MAX_DIMENSION in either direction through the 8f-3 / 8f-4 entry points: dithered
    import, picture cleanup / blend, use_argb import, animation blend, alpha filters, alpha
    level smoothing."""
    from go_webp_b200 import anim, filters as F, picture as P, picture_tools as T
    w, h = size
    rng = np.random.default_rng(w + h)
    argb = rng.integers(0, 1 << 32, (h, w), dtype=np.uint32)
    argb[rng.random((h, w)) < 0.5] |= 0xff000000
    pic = P.Picture(Width=w, Height=h, UseARGB=True, ARGB=argb.copy(), ARGBStride=w)
    assert P.WebPPictureARGBToYUVADithered(pic, P.WEBP_YUV420, 0.8) is None
    y, u, v, a, ha = oracle.import_yuva_dithered(argb.view(np.uint8).reshape(h, 4 * w), w, h, 0.8, "BGRA")
    assert (pic.Y == y).all() and (pic.U == u).all() and (pic.V == v).all() and (pic.A == a).all()
    bw, bh = (w + 7) // 8, (h + 7) // 8
    al = rng.integers(1, 256, (h, w), dtype=np.uint8)
    al[np.kron(rng.random((bh, bw)) < 0.7, np.ones((8, 8), bool))[:h, :w]] = 0
    al[rng.random((h, w)) < 0.05] = 0
    py = P.Picture(Width=w, Height=h, ColorSpace=P.WEBP_YUV420A, Y=y.copy(), U=u.copy(), V=v.copy(), A=al.copy())
    pa = P.Picture(Width=w, Height=h, UseARGB=True, ARGB=(argb & 0xffffff) | (al.astype(np.uint32) << 24), ARGBStride=w)
    src_argb = pa.ARGB.copy()
    T.WebPCleanupTransparentAreaBatch([py, pa])
    ey, eu, ev = oracle.cleanup_transparent_yuva(y, u, v, al)
    assert (py.Y == ey).all() and (py.U == eu).all() and (py.V == ev).all()
    assert (pa.ARGB == oracle.cleanup_transparent_argb(src_argb)).all()
    T.WebPBlendAlphaBatch([py, pa], 0x204060)
    by, bu, bv, ba = oracle.blend_alpha_yuva(ey, eu, ev, al, 0x204060)
    assert (py.Y == by).all() and (py.U == bu).all() and (py.V == bv).all() and (py.A == 255).all()
    assert (pa.ARGB == oracle.blend_alpha_argb(oracle.cleanup_transparent_argb(src_argb), 0x204060)).all()
    rgb = rng.integers(0, 256, (h, 3 * w), dtype=np.uint8)
    pr = P.Picture(Width=w, Height=h, UseARGB=True)
    assert P.WebPPictureImportBGR(pr, rgb, 3 * w) == 1
    assert (pr.ARGB == oracle.import_to_argb(rgb, w, h, "BGR")).all()
    dst = rng.integers(0, 1 << 32, (h, w), dtype=np.uint32)
    for premult in (False, True):
        assert (anim.BlendRowsBatch([src_argb], [dst], premult)[0] == oracle.anim_blend_rows(src_argb, dst, int(premult))).all()
    planes = [al] * 4
    un = F.AlphaUnfilterBatch(planes, [0, 1, 2, 3])
    for f, o in enumerate(un):
        assert (o == oracle.alpha_unfilter(al, f)).all(), f
    for f, o in enumerate(F.AlphaFilterBatch(un, [0, 1, 2, 3])):
        assert (o == al).all(), f
    q = ((al // 24) * 24).astype(np.uint8)                  # quantised levels, many tiles / bands
    assert (F.DequantizeLevelsBatch([q], [100], ordered_dither=False)[0] == oracle.dequantize_levels(q, 100, False)).all()
    d = F.DequantizeLevelsBatch([q, al], [55, 100], ordered_dither=True)
    assert (d[0] == oracle.dequantize_levels(q, 55, True)).all() and (d[1] == oracle.dequantize_levels(al, 100, True)).all()
