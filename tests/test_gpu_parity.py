"""GPU parity tests: the CUDA path, called through the C ABI (via the host
mirror of the reference API), against the CPU oracle and the committed libwebp
1.6.0 fixtures.  Bar: bit-exact (all of this is integer fixed point)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

SIZES = [(1, 1), (1, 2), (2, 1), (2, 2), (3, 3), (4, 4), (5, 7), (6, 8), (7, 10), (16, 33),
         (33, 16), (64, 64), (65, 31), (127, 129), (128, 2), (129, 3), (255, 9), (256, 8),
         (257, 5), (260, 4), (515, 7), (1024, 6)]
ALPHA = ["opaque", "rand", "sparse", "binary", "transp"]


@pytest.fixture(scope="module")
def gpu():
    from go_webp_b200 import _abi, device, dsp, picture
    if device.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return dict(abi=_abi, device=device, dsp=dsp, picture=picture, ctx=picture.default_context())


def make_rgba(rng, w, h, kind):
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    if kind == "opaque":
        rgba[..., 3] = 255
    elif kind == "sparse":
        rgba[..., 3][rng.random((h, w)) < 0.7] = 255
    elif kind == "binary":
        rgba[..., 3] = np.where(rng.random((h, w)) < 0.5, 0, 255)
    elif kind == "transp":
        rgba[..., 3] = 0
    return rgba


def check_import(pic, exp, ctxt):
    y, u, v, a, ha = exp
    P = ctxt["picture"]
    assert pic.ColorSpace == (P.WEBP_YUV420A if ha else P.WEBP_YUV420)
    assert (pic.Y == y).all() and (pic.U == u).all() and (pic.V == v).all()
    if ha:
        assert (pic.A == a).all()
    else:
        assert pic.A is None


# ---- part 1 ---------------------------------------------------------------------
@pytest.mark.parametrize("kind", ALPHA)
def test_import_size_sweep_batched(gpu, kind):
    """One mixed-size batch per alpha mix (odd widths/heights, 1x1 .. 1024x6)."""
    P = gpu["picture"]
    rng = np.random.default_rng(ALPHA.index(kind))
    srcs = [make_rgba(rng, w, h, kind) for (w, h) in SIZES]
    for table in (False, True):
        pics = [P.Picture(Width=w, Height=h) for (w, h) in SIZES]
        assert P.ImportRGBABatch(pics, srcs, [4 * w for (w, h) in SIZES], "RGBA", divide_table=table, chunk=5) == 1
        for pic, src, (w, h) in zip(pics, srcs, SIZES):
            exp = oracle.import_yuva(src, w, h, "RGBA",
                                     divide_mode=oracle.DIV_TABLE if table else oracle.DIV_EXACT)
            check_import(pic, exp, gpu)


FAST_SIZES = [(4, 1), (4, 2), (8, 3), (64, 64), (128, 2), (132, 5), (256, 8), (260, 4), (1024, 6),
              (1920, 8), (124, 7), (388, 2)]


@pytest.mark.parametrize("order", ["RGBA", "BGRA", "RGBX", "BGRX"])
@pytest.mark.parametrize("kind", ALPHA)
def test_import_tuned_kernel(gpu, order, kind):
    """Batches whose every image meets the tuned kernel's preconditions (width % 4
    == 0, 16-byte aligned rows: b2_import_fast.cu), all alpha mixes, both channel
    orders, with and without imported alpha, exact and table divide."""
    P = gpu["picture"]
    rng = np.random.default_rng(ALPHA.index(kind) * 7 + len(order))
    srcs = [make_rgba(rng, w, h, kind) for (w, h) in FAST_SIZES]
    for table in (False, True):
        pics = [P.Picture(Width=w, Height=h) for (w, h) in FAST_SIZES]
        assert P.ImportRGBABatch(pics, srcs, [4 * w for (w, h) in FAST_SIZES], order, divide_table=table) == 1
        for pic, src, (w, h) in zip(pics, srcs, FAST_SIZES):
            exp = oracle.import_yuva(src, w, h, order,
                                     divide_mode=oracle.DIV_TABLE if table else oracle.DIV_EXACT)
            check_import(pic, exp, gpu)


@pytest.mark.parametrize("order", ["BGRA", "RGBX", "BGRX", "RGB", "BGR"])
def test_import_orders_strides(gpu, order):
    P = gpu["picture"]
    rng = np.random.default_rng(7)
    step = 3 if len(order) == 3 else 4
    for (w, h, pad) in [(37, 23, 12), (64, 16, 0), (5, 5, 3), (130, 9, 16)]:
        stride = w * step + pad
        src = rng.integers(0, 256, (h, stride), dtype=np.uint8)
        pic = P.Picture(Width=w, Height=h)
        fn = getattr(P, "WebPPictureImport" + order)
        assert fn(pic, src, stride) == 1
        check_import(pic, oracle.import_yuva(src, w, h, order, stride=stride), gpu)
        # negative stride = vertically flipped buffer (buffer_dec.c.go:161-184)
        pic2 = P.Picture(Width=w, Height=h)
        assert P.ImportRGBABatch([pic2], [src], [-stride], order) == 1
        check_import(pic2, oracle.import_yuva(src, w, h, order, stride=-stride), gpu)


@pytest.mark.parametrize("order", ["RGB", "BGR"])
def test_import_three_byte_pixels_fast_kernel(gpu, order):
    """WebPPictureImportRGB / BGR through the tuned kernel (b2_import_fast.cu, STEP3): rows of every byte
    alignment (stride = 3 w + pad), widths around the 128-pixel strip and the 16-byte chunk, one and many
    rows, flipped buffers, several images per call."""
    P = gpu["picture"]
    rng = np.random.default_rng(33 + len(order))
    cases = [(1, 1, 0), (2, 3, 1), (3, 2, 0), (5, 5, 2), (16, 4, 0), (21, 7, 5), (127, 3, 0), (128, 2, 0), (129, 5, 1),
             (131, 4, 3), (256, 6, 0), (260, 3, 7), (515, 9, 0), (1030, 2, 2), (1920, 4, 0), (2050, 3, 1)]
    srcs, pics, strides = [], [], []
    for (w, h, pad) in cases:
        stride = 3 * w + pad
        srcs.append(rng.integers(0, 256, (h, stride), dtype=np.uint8))
        pics.append(P.Picture(Width=w, Height=h))
        strides.append(stride)
    assert P.ImportRGBABatch(pics, srcs, strides, order) == 1
    for pic, src, (w, h, pad), stride in zip(pics, srcs, cases, strides):
        check_import(pic, oracle.import_yuva(src, w, h, order, stride=stride), gpu)
    pics2 = [P.Picture(Width=w, Height=h) for (w, h, pad) in cases]
    assert P.ImportRGBABatch(pics2, srcs, [-s for s in strides], order) == 1
    for pic, src, (w, h, pad), stride in zip(pics2, srcs, cases, strides):
        check_import(pic, oracle.import_yuva(src, w, h, order, stride=-stride), gpu)


def test_import_golden_libwebp(gpu, golden_dir):
    """Against the libwebp 1.6.0 binary's outputs (table-divide mode)."""
    P = gpu["picture"]
    g = np.load(os.path.join(golden_dir, "import_libwebp160.npz"))
    n = int(g["n"][0])
    pics, srcs, strides = [], [], []
    for i in range(n):
        w, h, ha = (int(x) for x in g[f"c{i}_meta"])
        pics.append(P.Picture(Width=w, Height=h))
        srcs.append(g[f"c{i}_src"])
        strides.append(4 * w)
    assert P.ImportRGBABatch(pics, srcs, strides, "RGBA", divide_table=True) == 1
    for i, pic in enumerate(pics):
        w, h, ha = (int(x) for x in g[f"c{i}_meta"])
        assert (pic.Y == g[f"c{i}_y"]).all() and (pic.U == g[f"c{i}_u"]).all() and (pic.V == g[f"c{i}_v"]).all()
        assert (pic.A is not None) == bool(ha)
        if ha:
            assert (pic.A == g[f"c{i}_a"]).all()


def test_argb_to_yuva_route(gpu):
    """PictureARGBToYUVA (conversions.go:12-30): ARGB u32 -> same planes as the
    byte import of the equivalent RGBA image."""
    P = gpu["picture"]
    rng = np.random.default_rng(3)
    w, h = 45, 19
    rgba = make_rgba(rng, w, h, "sparse")
    argb = (rgba[..., 3].astype(np.uint32) << 24) | (rgba[..., 0].astype(np.uint32) << 16) | \
           (rgba[..., 1].astype(np.uint32) << 8) | rgba[..., 2]
    pic = P.Picture(Width=w, Height=h, UseARGB=True, ARGB=argb, ARGBStride=w)
    assert P.WebPPictureARGBToYUVA(pic, P.WEBP_YUV420) is None
    assert pic.UseARGB is False
    check_import(pic, oracle.import_yuva(rgba, w, h, "RGBA"), gpu)


def test_import_error_behaviour(gpu):
    P, abi, ctx = gpu["picture"], gpu["abi"], gpu["ctx"]
    pic = P.Picture(Width=0, Height=4)
    assert P.WebPPictureImportRGBA(pic, np.zeros(16, np.uint8), 4) == 0
    assert pic.ErrorCode == P.ENC_ERROR_BAD_DIMENSION
    pic = P.Picture(Width=8, Height=2)
    assert P.WebPPictureImportRGBA(pic, np.zeros(64, np.uint8), 16) == 0  # |stride| < 4*w
    assert P.WebPPictureImportRGBA(None, np.zeros(4, np.uint8), 4) == 0
    # raw ABI: null plane, oversize dimension, alpha with a 3-byte order
    job = abi.ImportJob(None, None, None, None, None, 16, 4, 2, 4, 4, 4)
    arr = (abi.ImportJob * 1)(job)
    assert ctx.lib.b2_rgba_to_yuv420a(ctx.h, 0, arr, 1, 0, 1, None) == abi.B2_ERR_NULL
    buf = ctx.alloc(1 << 16)
    job = abi.ImportJob(buf.ptr, buf.ptr, buf.ptr, buf.ptr, None, 4 * 16384, 16384, 8192, 0, 16384, 1)
    arr = (abi.ImportJob * 1)(job)
    assert ctx.lib.b2_rgba_to_yuv420a(ctx.h, 0, arr, 1, 0, 1, None) == abi.B2_ERR_BAD_DIM
    job = abi.ImportJob(buf.ptr, buf.ptr, buf.ptr, buf.ptr, None, 12, 4, 2, 0, 4, 4)
    arr = (abi.ImportJob * 1)(job)
    assert ctx.lib.b2_rgba_to_yuv420a(ctx.h, 0, arr, 1, abi.ORDER_RGB, abi.IMPORT_ALPHA, None) == abi.B2_ERR_INVALID_CONFIG
    assert ctx.lib.b2_rgba_to_yuv420a(ctx.h, 0, arr, 0, 0, 1, None) == abi.B2_OK  # empty batch
    assert ctx.lib.b2_rgba_to_yuv420a(ctx.h, 99, arr, 1, 0, 1, None) == abi.B2_ERR_INVALID_CONFIG


# ---- part 2 ---------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
def test_upsample_size_sweep_batched(gpu, mode):
    """Every WEBP_CSP_MODE; the 16-bit ones in the context's default order (the Go text's, ba / gb first)."""
    P = gpu["picture"]
    assert gpu["ctx"].swap_16bit_csp
    rng = np.random.default_rng(100 + mode)
    for with_a in (False, True):
        pics, exps = [], []
        for (w, h) in SIZES:
            uw, uh = (w + 1) // 2, (h + 1) // 2
            y = rng.integers(0, 256, (h, w), dtype=np.uint8)
            u = rng.integers(0, 256, (uh, uw), dtype=np.uint8)
            v = rng.integers(0, 256, (uh, uw), dtype=np.uint8)
            a = rng.integers(0, 256, (h, w), dtype=np.uint8) if with_a else None
            pics.append(P.Picture(Width=w, Height=h, Y=y, U=u, V=v, A=a,
                                  ColorSpace=P.WEBP_YUV420A if with_a else P.WEBP_YUV420))
            omode = {7: 1, 8: 3, 9: 4, 10: 5}.get(mode, mode)  # premultiplied modes alias (upsampling.c.go:280-288)
            exps.append(oracle.yuva_to_rgba(y, u, v, a, omode, swap16=True))
        outs = P.YUVAToRGBABatch(pics, mode, chunk=7)
        assert outs is not None
        for o, e, s in zip(outs, exps, SIZES):
            assert (o == e).all(), (s, mode, with_a)


def test_upsample_golden_libwebp(gpu, golden_dir):
    P = gpu["picture"]
    g = np.load(os.path.join(golden_dir, "upsample_libwebp160.npz"))
    for i in range(int(g["n"][0])):
        y, u, v, a = g[f"c{i}_y"], g[f"c{i}_u"], g[f"c{i}_v"], g[f"c{i}_a"]
        h, w = y.shape
        pic = P.Picture(Width=w, Height=h, Y=y, U=u, V=v, A=a if a.size else None,
                        ColorSpace=P.WEBP_YUV420A if a.size else P.WEBP_YUV420)
        assert P.WebPPictureYUVAToARGB(pic) == 1 and pic.UseARGB
        assert (pic.ARGB == g[f"c{i}_argb"]).all(), (w, h)


def test_decoder_golden_all_modes(gpu, golden_dir):
    """Real lossy files decoded by libwebp 1.6.0: planes in, every RGB mode out.  The binary is the stock
    build (rg byte first in the 16-bit modes): the context is switched to that order for the comparison."""
    P = gpu["picture"]
    g = np.load(os.path.join(golden_dir, "decode_libwebp160.npz"))
    gpu["ctx"].swap_16bit_csp = False
    try:
        for i in range(int(g["n"][0])):
            y, u, v, a = g[f"c{i}_y"], g[f"c{i}_u"], g[f"c{i}_v"], g[f"c{i}_a"]
            h, w = y.shape
            for mode in range(7):
                pic = P.Picture(Width=w, Height=h, Y=y, U=u, V=v, A=a if a.size else None,
                                ColorSpace=P.WEBP_YUV420A if a.size else P.WEBP_YUV420)
                got = P.YUVAToRGBABatch([pic], mode)[0]
                exp = g[f"c{i}_m{mode}"]
                assert (got == exp).all(), (i, mode)
    finally:
        gpu["ctx"].swap_16bit_csp = True


def test_upsample_line_pair_table(gpu):
    """WebPUpsamplers[mode](top_y, bottom_y, top_u, top_v, cur_u, cur_v, len) with
    the reference's argument meaning (dsp.h.go:144)."""
    D = gpu["dsp"]
    rng = np.random.default_rng(11)
    for n in (1, 2, 7, 8, 33, 300):
        uw = (n + 1) // 2
        ty, by = rng.integers(0, 256, n, dtype=np.uint8), rng.integers(0, 256, n, dtype=np.uint8)
        tu, tv, cu, cv = (rng.integers(0, 256, uw, dtype=np.uint8) for _ in range(4))
        for mode in (1, 3, 4, 0):
            top, bot = D.WebPUpsamplers[mode](ty, by, tu, tv, cu, cv, n)
            etop, ebot = oracle.upsample_line_pair(mode, ty, by, tu, tv, cu, cv)
            assert (top == etop).all() and (bot == ebot).all()
            top, bot = D.WebPUpsamplers[mode](ty, None, tu, tv, tu, tv, n)
            etop, _ = oracle.upsample_line_pair(mode, ty, None, tu, tv, tu, tv)
            assert (top == etop).all() and bot is None


def test_upsample_error_behaviour(gpu):
    P, abi, ctx = gpu["picture"], gpu["abi"], gpu["ctx"]
    pic = P.Picture(Width=4, Height=4)
    assert P.WebPPictureYUVAToARGB(pic) == 0 and pic.ErrorCode == P.ENC_ERROR_nil_PARAMETER
    y = np.zeros((4, 4), np.uint8)
    uv = np.zeros((2, 2), np.uint8)
    pic = P.Picture(Width=4, Height=4, Y=y, U=uv, V=uv, ColorSpace=P.WEBP_YUV420A)
    assert P.WebPPictureYUVAToARGB(pic) == 0 and pic.ErrorCode == P.ENC_ERROR_nil_PARAMETER
    pic = P.Picture(Width=4, Height=4, Y=y, U=uv, V=uv, ColorSpace=1)
    assert P.WebPPictureYUVAToARGB(pic) == 0 and pic.ErrorCode == P.ENC_ERROR_INVALID_CONFIGURATION
    buf = ctx.alloc(4096)
    job = abi.UpsampleJob(buf.ptr, buf.ptr, buf.ptr, None, buf.ptr, 8, 4, 0, 8, 8, 2)  # dst_stride < 4*w
    arr = (abi.UpsampleJob * 1)(job)
    assert ctx.lib.b2_yuv420a_to_rgba_fancy(ctx.h, 0, arr, 1, abi.MODE_RGBA) == abi.B2_ERR_BAD_DIM
    assert ctx.lib.b2_yuv420a_to_rgba_fancy(ctx.h, 0, arr, 1, 11) == abi.B2_ERR_INVALID_CONFIG


def test_import_then_upsample_roundtrip_config0(gpu):
    """BASELINE configs[0]: one 1920x1080 image, import then fancy upsample, GPU
    vs the CPU path end to end (opaque and random alpha)."""
    P = gpu["picture"]
    w, h = 1920, 1080
    for alpha_mode in (1, 0):
        rgba = oracle.synth_u32(0xB2000001, 0, 0, w, h, alpha_mode).view(np.uint8).reshape(h, w, 4)
        pic = P.Picture(Width=w, Height=h)
        assert P.WebPPictureImportRGBA(pic, rgba, 4 * w) == 1
        y, u, v, a, ha = oracle.import_yuva(rgba, w, h, "RGBA")
        check_import(pic, (y, u, v, a, ha), gpu)
        assert P.WebPPictureYUVAToARGB(pic) == 1
        exp = oracle.yuva_to_rgba(y, u, v, a if ha else None, oracle.MODE_BGRA).view(np.uint32).reshape(h, w)
        assert (pic.ARGB == exp).all()


# ---- part 3 ---------------------------------------------------------------------
VP8L_SIZES = [(1, 1), (3, 1), (4, 4), (5, 7), (33, 17), (100, 70), (128, 64), (129, 5), (256, 4), (1000, 3)]


@pytest.mark.parametrize("bits", [2, 3, 5, 9])
def test_vp8l_flat_transforms(gpu, bits):
    D, ctx, abi = gpu["dsp"], gpu["ctx"], gpu["abi"]
    rng = np.random.default_rng(bits)
    for (w, h) in VP8L_SIZES:
        x = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
        tw, th = D.VP8LSubSampleSize(w, bits), D.VP8LSubSampleSize(h, bits)
        codes = rng.integers(0, 2**24, (th, tw), dtype=np.uint32) | 0xff000000
        r = {op: D._flat(op, x, codes, w, h, bits) for op in
             ("add_green", "sub_green", "color_inverse", "color_forward", "inverse_chain", "forward_chain")}
        assert (r["add_green"] == oracle.add_green(x)).all()
        assert (r["sub_green"] == oracle.sub_green(x)).all()
        assert (r["color_inverse"] == oracle.color_space_inverse(x, codes, bits)).all()
        assert (r["color_forward"] == oracle.color_space_forward(x, codes, bits)).all()
        assert (r["inverse_chain"] == oracle.add_green(oracle.color_space_inverse(x, codes, bits))).all()
        fwd = oracle.color_space_forward(oracle.sub_green(x), codes, bits)
        assert (r["forward_chain"] == fwd).all()
        # inverse o forward == identity on the device path
        assert (D._flat("inverse_chain", r["forward_chain"], codes, w, h, bits) == x).all()


def test_vp8l_in_place(gpu):
    """dst may alias src for the same-size transforms (the reference's
    VP8LSubtractGreenFromBlueAndRed / VP8LTransformColor work in place)."""
    ctx, abi = gpu["ctx"], gpu["abi"]
    rng = np.random.default_rng(5)
    w, h = 260, 37
    x = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
    buf = ctx.upload(x)
    ctx.vp8l("sub_green", [abi.ArgbJob(buf.ptr, buf.ptr, None, w, h, 0, 0)])
    got = ctx.d2h(buf.ptr, x.nbytes).view(np.uint32).reshape(h, w)
    assert (got == oracle.sub_green(x)).all()


@pytest.mark.parametrize("bits", [0, 1, 2, 3])
def test_vp8l_color_index_and_bundle(gpu, bits):
    D = gpu["dsp"]
    rng = np.random.default_rng(40 + bits)
    ncol = 1 << (8 >> bits)
    for w in list(range(1, 41)) + [127, 128, 129, 255, 256, 257, 1000]:
        h = 3 if w < 100 else 5
        pal = rng.integers(0, 2**32, ncol, dtype=np.uint32)
        idx = rng.integers(0, ncol, (h, w), dtype=np.uint8)
        packed = D.VP8LBundleColorMap(idx, w, bits)
        assert (packed == oracle.bundle_color_map(idx, bits)).all(), w
        t = D.VP8LTransform(D.COLOR_INDEXING_TRANSFORM, bits, w, h, pal)
        out = D.VP8LInverseTransform(t, 0, h, packed)
        assert (out == oracle.color_index_inverse(packed, pal, bits, w)).all(), w
        assert (out == pal[idx]).all()


def test_vp8l_named_row_functions(gpu):
    D = gpu["dsp"]
    rng = np.random.default_rng(77)
    src = rng.integers(0, 2**32, 777, dtype=np.uint32)
    assert (D.VP8LAddGreenToBlueAndRed(src, 777) == oracle.add_green(src)).all()
    assert (D.VP8LSubtractGreenFromBlueAndRed(src, 700) == oracle.sub_green(src[:700])).all()
    code = D.MultipliersToColorCode(0x9c, 0x31, 0xe0)
    exp = np.empty_like(src)
    L = oracle.lib()
    L.orc_transform_color_inverse(C.c_uint32(code), src.ctypes.data_as(C.c_void_p), src.size,
                                  exp.ctypes.data_as(C.c_void_p))
    assert (D.VP8LTransformColorInverse(code, src) == exp).all()
    fw = src.copy()
    L.orc_transform_color(C.c_uint32(code), fw.ctypes.data_as(C.c_void_p), fw.size)
    assert (D.VP8LTransformColor(code, src) == fw).all()
    pal = rng.integers(0, 2**32, 200, dtype=np.uint32)
    img = (rng.integers(0, 200, (6, 50), dtype=np.uint32) << 8) | 0xff000000
    assert (D.VP8LMapColor32b(img, pal, 0, 6, 50) == pal[(img >> 8) & 0xff]).all()


def test_vp8l_inverse_transform_rows_and_batch(gpu):
    """VP8LInverseTransform on row windows (as ProcessRows feeds <=16 rows,
    vp8l_dec.c.go:733-760) and the batched chain entry point."""
    D = gpu["dsp"]
    rng = np.random.default_rng(88)
    w, h, bits = 70, 50, 3
    x = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
    codes = rng.integers(0, 2**24, (D.VP8LSubSampleSize(h, bits), D.VP8LSubSampleSize(w, bits)),
                         dtype=np.uint32) | 0xff000000
    t_cc = D.VP8LTransform(D.CROSS_COLOR_TRANSFORM, bits, w, h, codes)
    t_sg = D.VP8LTransform(D.SUBTRACT_GREEN_TRANSFORM, 0, w, h)
    full = oracle.color_space_inverse(x, codes, bits)
    for (r0, r1) in [(0, 16), (16, 32), (32, 48), (48, 50), (5, 21)]:
        assert (D.VP8LInverseTransform(t_cc, r0, r1, x[r0:r1]) == full[r0:r1]).all()
        assert (D.VP8LInverseTransform(t_sg, r0, r1, x[r0:r1]) == oracle.add_green(x[r0:r1])).all()
    # batch: image 0 = [sub-green, cross-colour] (fused), image 1 = colour index 2-bit,
    # image 2 = cross-colour only
    pal = rng.integers(0, 2**32, 4, dtype=np.uint32)
    idx = rng.integers(0, 4, (h, w), dtype=np.uint8)
    packed = oracle.bundle_color_map(idx, 2)
    t_ci = D.VP8LTransform(D.COLOR_INDEXING_TRANSFORM, 2, w, h, pal)
    outs = D.VP8LInverseTransformBatch(
        [D.VP8LImage(x, w, h), D.VP8LImage(packed, w, h), D.VP8LImage(x, w, h)],
        [[t_sg, t_cc], [t_ci], [t_cc]])
    assert (outs[0] == oracle.add_green(full)).all()
    assert (outs[1] == pal[idx]).all()
    assert (outs[2] == full).all()


def test_vp8l_error_behaviour(gpu):
    ctx, abi = gpu["ctx"], gpu["abi"]
    buf = ctx.alloc(4096)
    bad_bits = (abi.ArgbJob * 1)(abi.ArgbJob(buf.ptr, buf.ptr, buf.ptr, 8, 8, 1, 0))
    assert ctx.lib.b2_vp8l_color_inverse(ctx.h, 0, bad_bits, 1) == abi.B2_ERR_INVALID_CONFIG
    no_codes = (abi.ArgbJob * 1)(abi.ArgbJob(buf.ptr, buf.ptr, None, 8, 8, 3, 0))
    assert ctx.lib.b2_vp8l_inverse_chain(ctx.h, 0, no_codes, 1) == abi.B2_ERR_NULL
    alias = (abi.ArgbJob * 1)(abi.ArgbJob(buf.ptr, buf.ptr, buf.ptr, 8, 8, 2, 0))
    assert ctx.lib.b2_vp8l_color_index_inverse(ctx.h, 0, alias, 1) == abi.B2_ERR_INVALID_CONFIG
    bad_dim = (abi.ArgbJob * 1)(abi.ArgbJob(buf.ptr, buf.ptr, None, 0, 8, 0, 0))
    assert ctx.lib.b2_vp8l_add_green(ctx.h, 0, bad_dim, 1) == abi.B2_ERR_BAD_DIM


# ---- device-resident synthetic inputs ----------------------------------------------
def test_synth_generator_matches_oracle(gpu):
    ctx, abi = gpu["ctx"], gpu["abi"]
    w, h, n = 52, 9, 3
    buf = ctx.alloc(4 * w * h * n)
    for mode in range(4):
        ctx.synth_fill(buf.ptr, w, h, 4 * w, 4 * w * h, n, 0xB2000002, 5, 1, mode)
        got = ctx.d2h(buf.ptr, 4 * w * h * n).view(np.uint32).reshape(n, h, w)
        for i in range(n):
            assert (got[i] == oracle.synth_u32(0xB2000002, 5 + i, 1, w, h, mode)).all()


# ---- rows a14, a28, a29 of the hot-path table --------------------------------------
def _oracle_to_yuva(argb):
    """ConvertToYUVA (vp8l_dec.c.go:543-562) row by row with the oracle's row functions."""
    h, w = argb.shape
    uw = (w + 1) // 2
    y = np.stack([oracle.argb_to_y(argb[r]) for r in range(h)])
    u = np.zeros(((h + 1) // 2, uw), np.uint8)
    v = np.zeros_like(u)
    for r in range(h):
        u[r >> 1], v[r >> 1] = oracle.argb_to_uv(argb[r], u[r >> 1], v[r >> 1], do_store=(r & 1) == 0)
    return y, u, v, (argb >> 24).astype(np.uint8)


@pytest.mark.parametrize("size", [(1, 1), (2, 2), (3, 3), (5, 4), (7, 9), (64, 32), (129, 7), (260, 5), (1024, 3)])
def test_argb_to_yuva_plain(gpu, size):
    D = gpu["dsp"]
    w, h = size
    rng = np.random.default_rng(w * 31 + h)
    argb = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
    ey, eu, ev, ea = _oracle_to_yuva(argb)
    y, u, v, a = D.ConvertToYUVA(argb, with_alpha=True)
    assert (y == ey).all() and (u == eu).all() and (v == ev).all() and (a == ea).all()
    # streamed the way ExportYUVA feeds it: arbitrary row chunks, odd y_pos averaging
    # into the chroma stored by the previous call (yuv.go:100-102)
    u2 = np.zeros_like(eu)
    v2 = np.zeros_like(ev)
    pos = 0
    for n in [1, 2, 3, 1, 5, 1 << 20]:
        if pos >= h:
            break
        n = min(n, h - pos)
        cy, cu, cv, _ = D.ConvertToYUVA(argb[pos:pos + n], y_pos=pos, u=u2[pos >> 1], v=v2[pos >> 1])
        assert (cy == ey[pos:pos + n]).all()
        u2[pos >> 1:(pos >> 1) + cu.shape[0]] = cu
        v2[pos >> 1:(pos >> 1) + cv.shape[0]] = cv
        pos += n
    assert (u2 == eu).all() and (v2 == ev).all()
    # the row functions under the reference's names
    assert (D.WebPConvertARGBToY(argb[0], w) == ey[0]).all()
    ru, rv = D.WebPConvertARGBToUV(argb[h - 1], eu[0], ev[0], w, False)
    xu, xv = oracle.argb_to_uv(argb[h - 1], eu[0], ev[0], do_store=False)
    assert (ru == xu).all() and (rv == xv).all()


@pytest.mark.parametrize("bits", [0, 1, 2, 3])
def test_color_index_inverse_alpha(gpu, bits):
    D = gpu["dsp"]
    rng = np.random.default_rng(40 + bits)
    ncol = 1 << (8 >> bits)
    for (w, h) in [(1, 1), (3, 2), (7, 3), (8, 1), (33, 5), (128, 4), (257, 3), (4100, 2)]:
        pal = rng.integers(0, 2**32, ncol, dtype=np.uint32)
        sw = D.VP8LSubSampleSize(w, bits)
        src = rng.integers(0, 256, (h, sw), dtype=np.uint8)
        t = D.VP8LTransform(D.COLOR_INDEXING_TRANSFORM, bits, w, h, pal)
        got = D.VP8LColorIndexInverseTransformAlpha(t, 0, h, src)
        assert (got == oracle.color_index_inverse_alpha(src, pal, bits, w)).all(), (w, h, bits)
    pal = rng.integers(0, 2**32, 256, dtype=np.uint32)
    src = rng.integers(0, 256, (3, 50), dtype=np.uint8)
    assert (D.VP8LMapColor8b(src, pal, 0, 3, 50) == oracle.color_index_inverse_alpha(src, pal, 0, 50)).all()


def test_expand_color_map(gpu):
    D = gpu["dsp"]
    rng = np.random.default_rng(5)
    # bits from the colour count exactly as vp8l_dec.c.go:1248-1254
    for ncol in [1, 2, 3, 4, 5, 15, 16, 17, 100, 255, 256]:
        bits = 0 if ncol > 16 else 1 if ncol > 4 else 2 if ncol > 2 else 3
        data = rng.integers(0, 2**32, ncol, dtype=np.uint32)
        t = D.VP8LTransform(D.COLOR_INDEXING_TRANSFORM, bits, 10, 10, data)
        assert (D.ExpandColorMap(ncol, t) == oracle.expand_color_map(data, bits)).all(), ncol


def test_new_entry_points_reject_bad_arguments(gpu):
    A = gpu["abi"]
    ctx = gpu["ctx"]
    lib = ctx.lib
    buf = ctx.alloc(4096)
    j = A.AlphaIndexJob(buf.ptr, buf.ptr + 1024, buf.ptr + 2048, 8, 8, 4, 0)   # bits out of range
    assert lib.b2_vp8l_color_index_inverse_alpha(ctx.h, 0, (A.AlphaIndexJob * 1)(j), 1) == A.B2_ERR_INVALID_CONFIG
    j = A.AlphaIndexJob(buf.ptr, buf.ptr, buf.ptr + 2048, 8, 8, 0, 0)          # in place
    assert lib.b2_vp8l_color_index_inverse_alpha(ctx.h, 0, (A.AlphaIndexJob * 1)(j), 1) == A.B2_ERR_INVALID_CONFIG
    e = A.ExpandJob(buf.ptr, buf.ptr, 17, 1, )                                # 17 colours need bits == 0
    assert lib.b2_vp8l_expand_color_map(ctx.h, 0, (A.ExpandJob * 1)(e), 1) == A.B2_ERR_INVALID_CONFIG
    p = A.PlainJob(buf.ptr, buf.ptr, buf.ptr, buf.ptr, None, 4, 8, 4, 8, 8, 8, 0, 0)  # stride < width
    assert lib.b2_argb_to_yuva_plain(ctx.h, 0, (A.PlainJob * 1)(p), 1) == A.B2_ERR_BAD_DIM
    assert lib.b2_argb_to_yuva_plain(ctx.h, 0, None, 1) == A.B2_ERR_NULL
    buf.free()


# ---- predictor transform ----------------------------------------------
def _rand_modes(rng, w, h, bits, modes=None):
    tw, th = (w + (1 << bits) - 1) >> bits, (h + (1 << bits) - 1) >> bits
    m = rng.integers(0, 16, (th, tw), dtype=np.uint32) if modes is None else rng.choice(modes, (th, tw)).astype(np.uint32)
    return (m << 8) | 0xff000000 | rng.integers(0, 256, (th, tw), dtype=np.uint32)  # other bits are ignored


@pytest.mark.parametrize("bits", [2, 3, 5])
def test_predictor_inverse_and_forward(gpu, bits):
    """All 14 predictors (+ the two aliases of mode 0), ragged tiles, images narrower and
    wider than one 128-column wavefront tile, shorter and taller than one 32-row band."""
    D = gpu["dsp"]
    rng = np.random.default_rng(100 + bits)
    for (w, h) in [(1, 1), (1, 5), (2, 2), (5, 1), (7, 9), (33, 34), (64, 65), (127, 3), (129, 40),
                   (200, 33), (300, 70), (513, 97), (2, 40), (6, 33), (66, 70), (130, 64), (1026, 45), (2050, 35)]:   # the last six: even widths that are not multiples of 4 (pair-wise fast band)
        modes = _rand_modes(rng, w, h, bits)
        img = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
        t = D.VP8LTransform(D.PREDICTOR_TRANSFORM, bits, w, h, modes)
        got = D.VP8LInverseTransform(t, 0, h, img)
        assert (got == oracle.predictor_inverse(img, modes, bits)).all(), (w, h, bits)
        fwd = D.VP8LPredictorForward(img, modes, bits)
        assert (fwd == oracle.predictor_forward(img, modes, bits)).all(), (w, h, bits)
        # inverse(forward(x)) == x, the identity the codec relies on
        assert (D.VP8LInverseTransform(t, 0, h, fwd) == img).all(), (w, h, bits)


@pytest.mark.parametrize("mode", list(range(14)))
def test_predictor_each_mode_alone(gpu, mode):
    D = gpu["dsp"]
    rng = np.random.default_rng(mode)
    w, h, bits = 150, 45, 4
    modes = _rand_modes(rng, w, h, bits, [mode])
    # smooth-ish content plus extremes so the clamped / select predictors hit both branches
    img = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
    img[::3] &= 0x0f0f0f0f
    img[1::5] |= 0xf0f0f0f0
    t = D.VP8LTransform(D.PREDICTOR_TRANSFORM, bits, w, h, modes)
    assert (D.VP8LInverseTransform(t, 0, h, img) == oracle.predictor_inverse(img, modes, bits)).all()
    assert (D.VP8LPredictorForward(img, modes, bits) == oracle.predictor_forward(img, modes, bits)).all()


def test_predictor_in_batched_chain(gpu):
    """Full decoder order (vp8l_dec.c.go:710-729): predictor, cross-colour and subtract-green
    read in that order are undone last-read-first."""
    D = gpu["dsp"]
    rng = np.random.default_rng(9)
    imgs, chains, exp = [], [], []
    for (w, h) in [(96, 40), (257, 66)]:
        bits = 3
        modes = _rand_modes(rng, w, h, bits)
        codes = rng.integers(0, 2**24, (D.VP8LSubSampleSize(h, bits), D.VP8LSubSampleSize(w, bits)), dtype=np.uint32)
        px = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
        chains.append([D.VP8LTransform(D.PREDICTOR_TRANSFORM, bits, w, h, modes),
                       D.VP8LTransform(D.CROSS_COLOR_TRANSFORM, bits, w, h, codes),
                       D.VP8LTransform(D.SUBTRACT_GREEN_TRANSFORM, 0, w, h)])
        imgs.append(D.VP8LImage(px, w, h))
        e = oracle.add_green(px)
        e = oracle.color_space_inverse(e, codes, bits)
        exp.append(oracle.predictor_inverse(e, modes, bits))
    outs = D.VP8LInverseTransformBatch(imgs, chains)
    for o, e in zip(outs, exp):
        assert (o.reshape(e.shape) == e).all()


def test_emit_fancy_rgb_streaming(gpu):
    """EmitFancyRGB's macroblock-row batches (io_dec.c.go:65-115) through the row-range
    form of the kernel: every batch writes exactly the rows the reference would, and
    the assembled image equals the whole-image conversion (aligned fast kernel and
    ragged generic kernel)."""
    P, A = gpu["picture"], gpu["abi"]
    rng = np.random.default_rng(77)
    for (w, h, mode, alpha) in [(256, 64, A.MODE_RGBA, False), (520, 49, A.MODE_BGRA, True), (37, 35, A.MODE_RGB, False),
                                (1024, 34, A.MODE_ARGB, False), (64, 16, A.MODE_RGB_565, False)]:
        uw, uh = (w + 1) // 2, (h + 1) // 2
        y = rng.integers(0, 256, (h, w), dtype=np.uint8)
        u = rng.integers(0, 256, (uh, uw), dtype=np.uint8)
        v = rng.integers(0, 256, (uh, uw), dtype=np.uint8)
        a = rng.integers(0, 256, (h, w), dtype=np.uint8) if alpha else None
        exp = oracle.yuva_to_rgba(y, u, v, a, mode, swap16=True)
        pic = P.Picture(Width=w, Height=h, ColorSpace=P.WEBP_YUV420A if alpha else P.WEBP_YUV420, Y=y, U=u, V=v, A=a)
        bpp = A.MODE_BPP[mode]
        dst = np.full((h, w * bpp), 0xA5, np.uint8)
        total, mb_y = 0, 0
        while mb_y < h:
            mb_h = min(16, h - mb_y)
            before = dst.copy()
            n = P.EmitFancyRGB(pic, mode, mb_y, mb_h, dst)
            lo = max(mb_y - 1, 0)
            hi = h if mb_y + mb_h >= h else mb_y + mb_h - 1
            assert n == hi - lo
            assert (dst[:lo] == before[:lo]).all() and (dst[hi:] == before[hi:]).all()  # nothing else touched
            total += n
            mb_y += mb_h
        assert total == h
        assert (dst == exp.reshape(h, w * bpp)).all(), (w, h, mode)


# ---- output conversions -----------------------------------------------
@pytest.mark.parametrize("mode", list(range(11)))
def test_convert_from_bgra(gpu, mode):
    D = gpu["dsp"]
    rng = np.random.default_rng(200 + mode)
    for (w, h) in [(1, 1), (3, 2), (4, 4), (7, 5), (128, 3), (131, 4), (516, 2), (1024, 3), (1543, 2)]:
        argb = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
        argb[0, 0] |= 0xff000000            # opaque, translucent and transparent pixels all present
        if w > 2:
            argb[0, 1] &= 0x00ffffff
        for up16 in (False, True):
            got = D.VP8LConvertFromBGRA(argb, None, mode, upstream_16bit_order=up16)
            assert (got == oracle.convert_from_bgra(argb, mode, int(up16))).all(), (w, h, mode, up16)


def test_alpha_multiplies(gpu):
    D = gpu["dsp"]
    rng = np.random.default_rng(31)
    for (w, h) in [(1, 1), (5, 3), (64, 4), (130, 2), (512, 3), (1040, 2), (1536, 2), (1922, 3), (515, 5)]:   # (512..1536): whole 512-sample items; the last two: packed rows of odd alignment
        rgba = rng.integers(0, 256, (h, 4 * w + 8), dtype=np.uint8)   # stride > 4 w
        rgba[:, 3::4][rng.random((h, rgba[:, 3::4].shape[1])) < 0.3] = 255
        for af in (0, 1):
            exp = rgba.copy()
            exp[:, :4 * w] = oracle.apply_alpha_multiply(np.ascontiguousarray(rgba[:, :4 * w]), af)
            assert (D.WebPApplyAlphaMultiply(rgba, af, w, h, rgba.shape[1]) == exp).all()
        # tightly packed rows (stride = 4 w): widths that are not multiples of 4 are handed to the kernel as one row
        tight = np.ascontiguousarray(rgba[:, :4 * w])
        for af in (0, 1):
            assert (D.WebPApplyAlphaMultiply(tight, af, w, h, 4 * w) == oracle.apply_alpha_multiply(tight, af)).all(), (w, h, af)
        p16 = rng.integers(0, 256, (h, 2 * w), dtype=np.uint8)
        for pos in (0, 1):
            assert (D.WebPApplyAlphaMultiply4444(p16, w, h, 2 * w, pos) == oracle.apply_alpha_multiply_4444(p16, pos)).all()
        argb = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
        argb[rng.random((h, w)) < 0.2] |= 0xff000000
        argb[rng.random((h, w)) < 0.2] &= 0x00ffffff
        plane = rng.integers(0, 256, (h, w), dtype=np.uint8)
        alpha = rng.integers(0, 256, (h, w), dtype=np.uint8)
        alpha[rng.random((h, w)) < 0.2] = 255
        alpha[rng.random((h, w)) < 0.2] = 0
        for inv in (0, 1):
            if inv:   # the inverse scale needs premultiplied input (value <= alpha), as the reference asserts
                a8 = (argb >> 24)
                m = np.minimum
                src = (argb & 0xff000000) | m(argb & 0xff, a8) | (m((argb >> 8) & 0xff, a8) << 8) | (m((argb >> 16) & 0xff, a8) << 16)
                pl = np.minimum(plane, alpha)
            else:
                src, pl = argb, plane
            assert (D.WebPMultARGBRows(src, w, h, inv) == oracle.mult_argb_rows(src, inv)).all()
            assert (D.WebPMultRows(pl, alpha, w, h, inv) == oracle.mult_rows(pl, alpha, inv)).all()


@pytest.mark.parametrize("mode", list(range(11)))
def test_point_sampler_and_yuv444(gpu, mode):
    D = gpu["dsp"]
    rng = np.random.default_rng(300 + mode)
    for (w, h) in [(1, 1), (2, 3), (5, 4), (127, 3), (260, 5), (512, 2), (1030, 5), (1536, 3), (640, 3), (896, 2), (1920, 2)]:   # from 512 on: whole 512-pixel items, then rows whose last item is 1 - 3 whole 128-pixel groups
        y = rng.integers(0, 256, (h, w), dtype=np.uint8)
        u = rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8)
        v = rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8)
        assert (D.WebPSamplerProcessPlane(y, u, v, mode) == oracle.sampler_process_plane(y, u, v, mode, swap16=True)).all(), (w, h, mode)
        u4 = rng.integers(0, 256, (h, w), dtype=np.uint8)
        v4 = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert (D.WebPYUV444ToRGB(y, u4, v4, mode) == oracle.yuv444_to_rgb(y, u4, v4, mode, swap16=True)).all(), (w, h, mode)


def test_vp8l_golden_libwebp_streams(gpu, golden_dir):
    """Part 3 against the upstream binary: residual images and transform data cut out of
    libwebp 1.6.0 lossless streams (tests/golden/make_golden_vp8l.py) through
    VP8LInverseTransformBatch must give what WebPDecodeBGRA gave -- predictor with the
    encoder's own mode mix, cross-colour, subtract-green (fused and alone), colour
    indexing at 1/2/4/8 bits with ExpandColorMap on the device."""
    from test_oracle_golden import vp8l_vectors
    dsp = gpu["dsp"]
    imgs, chains, exps, names = [], [], [], []
    for name, res, trs, expected in vp8l_vectors(golden_dir):
        chain = []
        for (t, b, w, h, data) in trs:
            tr = dsp.VP8LTransform(t, b, w, h, data)
            if t == dsp.COLOR_INDEXING_TRANSFORM:
                tr.data = dsp.ExpandColorMap(int(data.size), tr)
            chain.append(tr)
        imgs.append(dsp.VP8LImage(res, res.shape[1], res.shape[0]))
        chains.append(chain)
        exps.append(expected)
        names.append(name)
    outs = dsp.VP8LInverseTransformBatch(imgs, chains)
    for name, o, e in zip(names, outs, exps):
        assert o.shape == e.shape and (o == e).all(), name


# ---- picture pre-processing ----------------------------------------------
def _yuva_pic(P, c):
    h, w = c["y"].shape
    return P.Picture(Width=w, Height=h, ColorSpace=P.WEBP_YUV420A, Y=c["y"].copy(), U=c["u"].copy(), V=c["v"].copy(),
                     A=c["a"].copy(), YStride=w, UVStride=(w + 1) >> 1, AStride=w)


def _argb_pic(P, argb):
    h, w = argb.shape
    return P.Picture(Width=w, Height=h, UseARGB=True, ARGB=argb.copy(), ARGBStride=w)


def test_picture_tools_golden_libwebp(gpu, golden_dir):
    """WebPCleanupTransparentArea / WebPBlendAlpha through the batched entry points on the
    vectors the libwebp 1.6.0 binary produced (YUVA and ARGB pictures in one batch)."""
    from test_oracle_golden import tools_vectors
    from go_webp_b200 import picture_tools as T
    P = gpu["picture"]
    cases = [c for _, c in tools_vectors(golden_dir)]
    pics = [_yuva_pic(P, c) for c in cases] + [_argb_pic(P, c["argb"]) for c in cases]
    T.WebPCleanupTransparentAreaBatch(pics)
    for c, py, pa in zip(cases, pics[:len(cases)], pics[len(cases):]):
        assert (py.Y == c["clean_y"]).all() and (py.U == c["clean_u"]).all() and (py.V == c["clean_v"]).all()
        assert (py.A == c["a"]).all()
        assert (pa.ARGB == c["clean_argb"]).all()
    for c in cases:   # one background colour per call
        py, pa = _yuva_pic(P, c), _argb_pic(P, c["argb"])
        T.WebPBlendAlphaBatch([py, pa], int(c["bg"]))
        assert (py.Y == c["blend_y"]).all() and (py.U == c["blend_u"]).all() and (py.V == c["blend_v"]).all()
        assert (py.A == 255).all() and (pa.ARGB == c["blend_argb"]).all()


@pytest.mark.parametrize("size", [(8, 8), (7, 7), (40, 9), (264, 16), (263, 23), (1000, 33), (2056, 64), (521, 130)])
def test_picture_tools_vs_oracle(gpu, size):
    """Same functions on seeded pictures against the oracle: long runs of transparent
    blocks across the 32-block chunks of a warp, ragged right / bottom edges, all-
    transparent and opaque pictures; WebPReplaceTransparentPixels."""
    from go_webp_b200 import picture_tools as T
    P = gpu["picture"]
    w, h = size
    rng = np.random.default_rng(w * 131 + h)
    for kind in ("runs", "speckle", "transp", "opaque"):
        a = rng.integers(1, 256, (h, w), dtype=np.uint8)
        if kind == "runs":
            bw, bh = (w + 7) // 8, (h + 7) // 8
            m = rng.random((bh, bw)) < 0.8
            a[np.kron(m, np.ones((8, 8), bool))[:h, :w]] = 0
        elif kind == "speckle":
            a[rng.random((h, w)) < 0.5] = 0
        elif kind == "transp":
            a[:] = 0
        else:
            a[:] = 255
        c = dict(y=rng.integers(0, 256, (h, w), dtype=np.uint8), u=rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8),
                 v=rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8), a=a)
        argb = rng.integers(0, 1 << 24, (h, w), dtype=np.uint32) | (a.astype(np.uint32) << 24)
        bg = int(rng.integers(0, 1 << 24))
        py, pa = _yuva_pic(P, c), _argb_pic(P, argb)
        T.WebPCleanupTransparentAreaBatch([py, pa])
        ey, eu, ev = oracle.cleanup_transparent_yuva(c["y"], c["u"], c["v"], a)
        assert (py.Y == ey).all() and (py.U == eu).all() and (py.V == ev).all(), kind
        assert (pa.ARGB == oracle.cleanup_transparent_argb(argb)).all(), kind
        py, pa = _yuva_pic(P, c), _argb_pic(P, argb)
        T.WebPBlendAlpha(py, bg)
        T.WebPBlendAlpha(pa, bg)
        ey, eu, ev, ea = oracle.blend_alpha_yuva(c["y"], c["u"], c["v"], a, bg)
        assert (py.Y == ey).all() and (py.U == eu).all() and (py.V == ev).all() and (py.A == ea).all(), kind
        assert (pa.ARGB == oracle.blend_alpha_argb(argb, bg)).all(), kind
        pa = _argb_pic(P, argb)
        T.WebPReplaceTransparentPixels(pa, 0xff000000 | bg)
        assert (pa.ARGB == oracle.replace_transparent(argb, bg)).all(), kind
    # reference "nothing to do" rules: YUV picture without alpha, non-ARGB picture for replace
    p = P.Picture(Width=w, Height=h, ColorSpace=P.WEBP_YUV420, Y=c["y"].copy(), U=c["u"].copy(), V=c["v"].copy())
    T.WebPBlendAlpha(p, 0x123456)
    T.WebPCleanupTransparentArea(p)
    T.WebPReplaceTransparentPixels(p, 0)
    assert (p.Y == c["y"]).all() and p.A is None


def test_import_to_argb(gpu, golden_dir):
    """use_argb import (Import(), picture_csp_enc.c.go:138-181): libwebp vectors for all six
    entry points in one batch, then seeded sizes / flipped and padded strides vs the oracle."""
    P = gpu["picture"]
    g = np.load(os.path.join(golden_dir, "import_argb_libwebp160.npz"))
    by_order = {}
    for k in g["names"]:
        by_order.setdefault(str(g[f"{k}/order"]), []).append((g[f"{k}/src"], g[f"{k}/argb"]))
    for order, cases in by_order.items():
        step = 3 if order in ("RGB", "BGR") else 4
        pics = [P.Picture(Width=e.shape[1], Height=e.shape[0], UseARGB=True) for _, e in cases]
        assert P.ImportRGBABatch(pics, [s for s, _ in cases], [step * e.shape[1] for _, e in cases], order) == 1
        for pic, (_, e) in zip(pics, cases):
            assert pic.UseARGB and (pic.ARGB == e).all(), order
    rng = np.random.default_rng(77)
    for (w, h) in [(1, 1), (3, 2), (129, 7), (1000, 3), (1921, 5)]:
        for order in ("RGBA", "BGRX", "RGB", "BGR"):
            step = 3 if order in ("RGB", "BGR") else 4
            stride = step * w + 8
            buf = rng.integers(0, 256, (h, stride), dtype=np.uint8)
            pic = P.Picture(Width=w, Height=h, UseARGB=True)
            fn = getattr(P, "WebPPictureImport" + order)
            assert fn(pic, buf, stride) == 1
            assert (pic.ARGB == oracle.import_to_argb(buf, w, h, order, stride)).all(), (w, h, order)
            pic = P.Picture(Width=w, Height=h, UseARGB=True)      # bottom-up source
            assert fn(pic, buf, -stride) == 1
            assert (pic.ARGB == oracle.import_to_argb(buf, w, h, order, -stride)).all(), (w, h, order, "flipped")
    pic = P.Picture(Width=8, Height=2, UseARGB=True)
    assert P.WebPPictureImportRGBA(pic, np.zeros((2, 16), np.uint8), 16) == 0   # stride < 4 * width (:132)


def test_picture_tools_and_argb_import_error_behaviour(gpu):
    """Raw ABI of the 8f-3 entry points: null planes, bad dimensions / strides, bad order,
    empty batch, bad stream -- never a launch, always a status."""
    abi, ctx = gpu["abi"], gpu["ctx"]
    lib, h = ctx.lib, ctx.h
    buf = ctx.alloc(1 << 16)
    one = lambda j: (type(j) * 1)(j)  # noqa: E731
    yuva = abi.PictureJob(buf.ptr, buf.ptr, buf.ptr, None, None, 16, 8, 16, 0, 16, 8)       # a == NULL
    assert lib.b2_picture_cleanup_transparent(h, 0, one(yuva), 1, 0) == abi.B2_ERR_NULL
    assert lib.b2_picture_blend_alpha(h, 0, one(yuva), 1, 0, 0) == abi.B2_ERR_NULL
    argb = abi.PictureJob(None, None, None, None, None, 0, 0, 0, 16, 16, 8)                  # argb == NULL
    assert lib.b2_picture_cleanup_transparent(h, 0, one(argb), 1, 1) == abi.B2_ERR_NULL
    assert lib.b2_picture_replace_transparent(h, 0, one(argb), 1, 0) == abi.B2_ERR_NULL
    bad = abi.PictureJob(None, None, None, None, buf.ptr, 0, 0, 0, 8, 16, 8)                 # argb_stride < width
    assert lib.b2_picture_blend_alpha(h, 0, one(bad), 1, 1, 0) == abi.B2_ERR_BAD_DIM
    bad = abi.PictureJob(buf.ptr, buf.ptr, buf.ptr, buf.ptr, None, 16, 4, 16, 0, 16, 8)      # uv_stride < (w + 1) / 2
    assert lib.b2_picture_cleanup_transparent(h, 0, one(bad), 1, 0) == abi.B2_ERR_BAD_DIM
    bad = abi.PictureJob(None, None, None, None, buf.ptr, 0, 0, 0, 20000, 16384, 1)           # > WEBP_MAX_DIMENSION
    assert lib.b2_picture_replace_transparent(h, 0, one(bad), 1, 0) == abi.B2_ERR_BAD_DIM
    ok = abi.PictureJob(None, None, None, None, buf.ptr, 0, 0, 0, 16, 16, 4)                  # lower than a block row: nothing to do
    assert lib.b2_picture_cleanup_transparent(h, 0, one(ok), 1, 1) == abi.B2_OK
    assert lib.b2_picture_cleanup_transparent(h, 0, one(ok), 0, 1) == abi.B2_OK               # empty batch
    assert lib.b2_picture_cleanup_transparent(h, 99, one(ok), 1, 1) == abi.B2_ERR_INVALID_CONFIG
    cj = abi.ConvertJob(None, None, None, buf.ptr, 64, 0, 64, 16, 2, 0)                       # src == NULL
    assert lib.b2_import_to_argb(h, 0, one(cj), 1, abi.ORDER_RGBA, 1) == abi.B2_ERR_NULL
    cj = abi.ConvertJob(buf.ptr, None, None, buf.ptr, 48, 0, 64, 16, 2, 0)
    assert lib.b2_import_to_argb(h, 0, one(cj), 1, abi.ORDER_RGB, 1) == abi.B2_ERR_INVALID_CONFIG   # alpha needs 4 bytes
    assert lib.b2_import_to_argb(h, 0, one(cj), 1, abi.ORDER_RGBA, 1) == abi.B2_ERR_BAD_DIM         # |stride| < 4 * width
    assert lib.b2_import_to_argb(h, 0, one(cj), 1, 7, 0) == abi.B2_ERR_INVALID_CONFIG
    cj = abi.ConvertJob(buf.ptr, None, None, buf.ptr, 64, 0, 32, 16, 2, 0)                    # dst_stride < 4 * width
    assert lib.b2_import_to_argb(h, 0, one(cj), 1, abi.ORDER_RGBA, 1) == abi.B2_ERR_BAD_DIM
    ctx.sync()
    buf.free()


def test_dithered_import(gpu, golden_dir):
    """PictureARGBToYUVA with dithering (conversions.go:122-165): the libwebp vectors
    (binary divide), then seeded pictures vs the oracle in the exact-divide mode: widths
    over several 128-pixel strips, odd sizes, a last odd row, amp 0, and a mixed-width batch
    in ONE call (every image runs its own generator, reached by skip-ahead per row pair)."""
    P = gpu["picture"]
    g = np.load(os.path.join(golden_dir, "import_dithered_libwebp160.npz"))
    for k in g["names"]:
        argb, d = g[f"{k}/argb"], float(g[f"{k}/d"])
        h, w = argb.shape
        pic = P.Picture(Width=w, Height=h, UseARGB=True, ARGB=argb.copy(), ARGBStride=w)
        assert P.PictureARGBToYUVA(pic, P.WEBP_YUV420, d, 0, divide_table=True) is None
        ha = bool(g[f"{k}/ha"])
        assert (pic.Y == g[f"{k}/y"]).all() and (pic.U == g[f"{k}/u"]).all() and (pic.V == g[f"{k}/v"]).all(), k
        assert pic.ColorSpace == (P.WEBP_YUV420A if ha else P.WEBP_YUV420)
        if ha:
            assert (pic.A == g[f"{k}/a"]).all(), k
    rng = np.random.default_rng(4242)
    sizes = [(1, 1), (3, 5), (128, 2), (129, 3), (257, 8), (640, 7), (1000, 33), (1921, 6), (64, 301)]
    pics, srcs, exps = [], [], []
    for i, (w, h) in enumerate(sizes):
        argb = rng.integers(0, 1 << 32, (h, w), dtype=np.uint32)
        if i % 3 == 0:
            argb |= 0xff000000
        elif i % 3 == 1:
            argb[rng.random((h, w)) < 0.6] |= 0xff000000
        d = [0.3, 1.0, 0.004][i % 3]
        pic = P.Picture(Width=w, Height=h, UseARGB=True, ARGB=argb.copy(), ARGBStride=w)
        assert P.WebPPictureARGBToYUVADithered(pic, P.WEBP_YUV420, d) is None
        y, u, v, a, ha = oracle.import_yuva_dithered(argb.view(np.uint8).reshape(h, 4 * w), w, h, d, "BGRA")
        assert (pic.Y == y).all(), (w, h, d, np.argwhere(pic.Y != y)[:4])
        assert (pic.U == u).all() and (pic.V == v).all(), (w, h, d)
        if ha:
            assert (pic.A == a).all()
        pics.append(P.Picture(Width=w, Height=h))
        srcs.append(argb.view(np.uint8).reshape(h, 4 * w))
        exps.append(oracle.import_yuva_dithered(srcs[-1], w, h, 0.7, "BGRA"))
    assert P.ImportRGBABatch(pics, srcs, [4 * w for (w, h) in sizes], "BGRA", dithering=0.7) == 1
    for pic, (y, u, v, a, ha) in zip(pics, exps):
        assert (pic.Y == y).all() and (pic.U == u).all() and (pic.V == v).all()
    # 3-byte pixels are outside the dithered path's preconditions
    pic = P.Picture(Width=8, Height=2)
    assert P.ImportRGBABatch([pic], [np.zeros((2, 24), np.uint8)], [24], "RGB", dithering=0.5) == 0


def test_anim_blend_and_compare(gpu, golden_dir):
    """Animation pixel work: the WebPAnimDecoder vectors through
    BlendRowsBatch (both blends in one batch each), seeded canvases vs the oracle, the row
    functions, and ComparePixels{Lossless,Lossy} over rows and columns vs the oracle."""
    from go_webp_b200 import anim
    g = np.load(os.path.join(golden_dir, "anim_libwebp160.npz"))
    for premult in (0, 1):
        ks = [str(k) for k in g["names"] if int(g[f"{k}/premult"]) == premult]
        outs = anim.BlendRowsBatch([g[f"{k}/src"] for k in ks], [g[f"{k}/dst"] for k in ks], bool(premult))
        for k, o in zip(ks, outs):
            assert (o == g[f"{k}/exp"]).all(), k
    rng = np.random.default_rng(909)
    for (w, h) in [(1, 1), (5, 3), (130, 9), (1027, 4), (2500, 3)]:
        src = rng.integers(0, 1 << 32, (h, w), dtype=np.uint32)
        dst = rng.integers(0, 1 << 32, (h, w), dtype=np.uint32)
        src[rng.random((h, w)) < 0.3] |= 0xff000000
        src[rng.random((h, w)) < 0.2] &= 0x00ffffff
        for premult in (False, True):
            o = anim.BlendRowsBatch([src], [dst], premult)[0]
            assert (o == oracle.anim_blend_rows(src, dst, int(premult))).all(), (w, h, premult)
    row, prev = src[0].copy(), dst[0].copy()
    anim.BlendPixelRowNonPremult(row, prev, w - 1)
    exp = oracle.anim_blend_rows(src[:1, :w - 1], dst[:1, :w - 1], 0)[0]
    assert (row[:w - 1] == exp).all() and row[w - 1] == src[0, w - 1]
    # frame comparison: rows and columns of a 64 x 48 pair of frames that differ in a rectangle
    W, H = 64, 48
    a = rng.integers(0, 1 << 32, (H, W), dtype=np.uint32)
    b = a.copy()
    b[10:20, 30:40] ^= rng.integers(0, 8, (10, 10), dtype=np.uint32) << 8     # small green differences
    runs = [(y * W, 1, y * W, 1, W) for y in range(H)] + [(x, W, x, W, H) for x in range(W)]
    for lossless, diff in ((True, 0), (False, 0), (False, 3), (False, 255)):
        got = anim.ComparePixelsBatch(a, b, runs, lossless, diff)
        exp = [oracle.compare_pixels(a.reshape(-1)[so:], ss, b.reshape(-1)[do:], ds, ln, lossless, diff) for (so, ss, do, ds, ln) in runs]
        assert got == exp, (lossless, diff)
    assert anim.ComparePixelsLossless(a[0], 1, a[0].copy(), 1, W) == 1
    assert anim.ComparePixelsLossy(a[12], 1, b[12], 1, W, 0) == oracle.compare_pixels(a[12], 1, b[12], 1, W, False, 0)


def test_alpha_filters(gpu, golden_dir):
    """Alpha-plane filters: the libwebp vectors (all four filters in one
    batch), then seeded planes vs the oracle: widths over several 128-column segments, more
    than 32 rows (several wavefront bands for the gradient), odd sizes, and
    filter(unfilter(x)) == x."""
    from go_webp_b200 import filters as F
    g = np.load(os.path.join(golden_dir, "alpha_unfilter_libwebp160.npz"))
    ks = [str(k) for k in g["names"]]
    outs = F.AlphaUnfilterBatch([g[f"{k}/in"] for k in ks], [int(g[f"{k}/filter"]) for k in ks])
    for k, o in zip(ks, outs):
        assert (o == g[f"{k}/out"]).all(), (k, int(g[f"{k}/filter"]))
    fwd = F.AlphaFilterBatch([g[f"{k}/out"] for k in ks], [int(g[f"{k}/filter"]) for k in ks])
    for k, o in zip(ks, fwd):
        assert (o == g[f"{k}/in"]).all(), k
    rng = np.random.default_rng(31337)
    planes, filts = [], []
    for (w, h) in [(1, 1), (1, 40), (5, 3), (129, 33), (300, 70), (1000, 9), (257, 100)]:
        for f in range(4):
            planes.append(rng.integers(0, 256, (h, w), dtype=np.uint8))
            filts.append(f)
    un = F.AlphaUnfilterBatch(planes, filts)
    for p, f, o in zip(planes, filts, un):
        assert (o == oracle.alpha_unfilter(p, f)).all(), (p.shape, f)
    back = F.AlphaFilterBatch(un, filts)
    for p, f, o in zip(planes, filts, back):
        assert (o == p).all() and (o == oracle.alpha_filter(oracle.alpha_unfilter(p, f), f)).all(), (p.shape, f)
    # rows and pointers that are NOT word aligned (the byte-wise gradient wavefront with its
    # packed / done & ~3 / ragged-tail logic, the non-aligned horizontal and vertical unfilters, the scalar
    # forward filter) and planes taller than 512 rows (a warp goes through more than 16 bands)
    planes, filts = [], []
    for (w, h) in [(1, 3), (3, 70), (37, 41), (129, 33), (255, 130), (301, 35), (130, 700), (64, 1030)]:
        for f in range(4):
            planes.append(rng.integers(0, 256, (h, w), dtype=np.uint8))
            filts.append(f)
    for (row_align, ptr_offset) in [(1, 0), (1, 1), (4, 2), (16, 3), (2, 5)]:
        un = F.AlphaUnfilterBatch(planes, filts, row_align=row_align, ptr_offset=ptr_offset)
        for p, f, o in zip(planes, filts, un):
            assert (o == oracle.alpha_unfilter(p, f)).all(), (p.shape, f, row_align, ptr_offset)
        back = F.AlphaFilterBatch(un, filts, row_align=row_align, ptr_offset=ptr_offset)
        for p, f, o in zip(planes, filts, back):
            assert (o == p).all(), (p.shape, f, row_align, ptr_offset)
    # wide planes: the gradient wavefront's steady state (whole 8-word blocks of every row of a band in flight),
    # rows ending inside a word and inside a block, a short last band
    planes = [rng.integers(0, 256, (h, w), dtype=np.uint8) for (w, h) in [(1920, 300), (1923, 100), (2048, 64), (997, 161), (4099, 37)]]
    for row_align in (16, 4):
        un = F.AlphaUnfilterBatch(planes, [3] * len(planes), row_align=row_align)
        for p, o in zip(planes, un):
            assert (o == oracle.alpha_unfilter(p, 3)).all(), (p.shape, row_align)
    # forward filters on 16-byte aligned rows whose width is a multiple of 16 (the templated whole-unit path): every
    # filter, rows shorter than / equal to / longer than a 512-pixel unit, groups of 4 rows with a short last group
    planes = [rng.integers(0, 256, (h, w), dtype=np.uint8) for (w, h) in [(1920, 37), (2048, 9), (528, 70), (512, 4), (16, 5), (1040, 3)]]
    for f in (1, 2, 3):
        fw = F.AlphaFilterBatch(planes, [f] * len(planes), row_align=16)
        for p, o in zip(planes, fw):
            assert (o == oracle.alpha_filter(p, f)).all(), (p.shape, f)
    # more gradient planes than resident CTAs (2 per SM): a CTA goes through several planes, its shared progress
    # words, ring and tiles are reused from plane to plane
    planes = [rng.integers(0, 256, (33 + (i % 5) * 16, 60 + (i % 7) * 8), dtype=np.uint8) for i in range(700)]
    un = F.AlphaUnfilterBatch(planes, [3] * len(planes))
    for i, (p, o) in enumerate(zip(planes, un)):
        assert (o == oracle.alpha_unfilter(p, 3)).all(), (i, p.shape)
    data = rng.integers(0, 256, 40 * 7, dtype=np.uint8)
    out = np.zeros_like(data)
    F.GradientFilter_C(data, 33, 7, 40, out)
    assert (out.reshape(7, 40)[:, :33] == oracle.alpha_filter(data.reshape(7, 40)[:, :33], 3)).all()


def test_rescaled_decoder_output(gpu, golden_dir):
    """EmitRescaledRGB / EmitRescaledYUV (io_dec.c.go:232-515) on the planes the libwebp decoder
    emitted, vs what the same decoder returned with options.use_scaling: eight RGB-family
    modes (alpha dispatched, premultiplied variants) and MODE_YUVA; plus WebPDispatchAlpha on
    a seeded buffer with the alpha byte first and last."""
    from go_webp_b200 import _abi, dsp as D, rescaler as R
    g = np.load(os.path.join(golden_dir, "rescaled_decode_libwebp160.npz"))
    for k in [str(k) for k in g["names"]]:
        y, u, v, a = (g[f"{k}/{nm}"] for nm in "yuva")
        dw, dh = (int(t) for t in g[f"{k}/size"])
        for mode in (_abi.MODE_RGB, _abi.MODE_RGBA, _abi.MODE_BGR, _abi.MODE_BGRA, _abi.MODE_ARGB, _abi.MODE_rgbA, _abi.MODE_bgrA, _abi.MODE_Argb):
            got = R.EmitRescaledRGB(y, u, v, a, dw, dh, mode)
            assert (got == g[f"{k}/rgb{mode}"]).all(), (k, mode)
        for got, nm in zip(R.EmitRescaledYUV(y, u, v, a, dw, dh), "yuva"):
            assert (got == g[f"{k}/out_{nm}"]).all(), (k, nm)
    rng = np.random.default_rng(77)
    buf = rng.integers(0, 256, (37, 4 * 131 + 8), dtype=np.uint8)
    al = rng.integers(0, 256, (37, 131), dtype=np.uint8)
    for first in (0, 1):
        exp = buf.copy()
        exp[:, (0 if first else 3):4 * 131:4] = al
        assert (D.WebPDispatchAlpha(al, 131, 37, buf, 4 * 131 + 8, first) == exp).all()


def test_alpha_level_smoothing(gpu, golden_dir):
    """WebPDequantizeLevels: the libwebp decoder vectors in one batch (stock
    build, no dither), the same planes with the Go file's ordered dither vs the oracle, then
    seeded quantised planes: several 120-column tiles and 32-row bands, widths around the tile
    edge, radius clamped by a thin plane, unaligned rows, two-level planes (copied), strength
    0 (copied), mixed strengths in one batch."""
    from go_webp_b200 import filters as F
    g = np.load(os.path.join(golden_dir, "alpha_dequant_libwebp160.npz"))
    ks = [str(k) for k in g["names"]]
    ins, strengths = [g[f"{k}/in"] for k in ks], [int(g[f"{k}/strength"]) for k in ks]
    outs = F.DequantizeLevelsBatch(ins, strengths, ordered_dither=False)
    for k, o in zip(ks, outs):
        assert (o == g[f"{k}/out"]).all(), k
    outs = F.DequantizeLevelsBatch(ins, strengths, ordered_dither=True)
    for k, p, s, o in zip(ks, ins, strengths, outs):
        assert (o == oracle.dequantize_levels(p, s, True)).all(), k
    rng = np.random.default_rng(4242)
    planes, strengths = [], []
    for (w, h) in [(1, 1), (2, 9), (3, 3), (9, 2), (119, 33), (120, 64), (121, 65), (124, 31), (240, 40), (245, 97), (1000, 70), (333, 200)]:
        yy, xx = np.mgrid[0:h, 0:w]
        base = (128 + 90 * np.sin(xx / 23.0 + 1) * np.cos(yy / 17.0) + rng.integers(-6, 7, (h, w))).clip(0, 255)
        for q, s in [(17, 100), (40, 60), (8, 30), (64, 100), (17, 0), (255, 100)]:
            planes.append(((base // q) * q).astype(np.uint8))
            strengths.append(s)
    for dither in (False, True):
        for align in (16, 1):
            outs = F.DequantizeLevelsBatch(planes, strengths, ordered_dither=dither, row_align=align)
            for p, s, o in zip(planes, strengths, outs):
                assert (o == oracle.dequantize_levels(p, s, dither)).all(), (p.shape, s, dither, align)
    data = np.ascontiguousarray(planes[40]).copy()
    h, w = data.shape
    flat = data.reshape(-1)
    assert F.WebPDequantizeLevels(flat, w - 3, h, w, 100)
    exp = data.copy()
    exp[:, :w - 3] = oracle.dequantize_levels(planes[40][:, :w - 3], 100, True)
    assert (flat.reshape(h, w) == exp).all()
    assert not F.WebPDequantizeLevels(flat, w, h, w, 101)
    with pytest.raises(Exception):
        F.DequantizeLevelsBatch([planes[0]], [-1])


def test_rescaler(gpu, golden_dir):
    """The rescaler: the WebPPictureRescale vectors in one batch (all mixes of
    horizontal / vertical expansion and shrinking, identity, 1-pixel sources), seeded planes
    with 1, 3 and 4 interleaved channels vs the oracle, WebPPictureRescale on a picture."""
    from go_webp_b200 import rescaler as R
    P = gpu["picture"]
    g = np.load(os.path.join(golden_dir, "rescale_libwebp160.npz"))
    ks = [str(k) for k in g["names"]]
    outs = R.RescalePlanesBatch([g[f"{k}/src"] for k in ks], [g[f"{k}/dst"].shape[::-1] for k in ks])
    for k, o in zip(ks, outs):
        assert (o == g[f"{k}/dst"]).all(), k
    rng = np.random.default_rng(2718)
    for ch in (1, 3, 4):
        for (w, h, dw, dh) in [(640, 360, 213, 120), (100, 80, 333, 201), (257, 3, 64, 9), (31, 200, 90, 41), (500, 20, 500, 7)]:
            src = rng.integers(0, 256, (h, w * ch), dtype=np.uint8)
            o = R.RescalePlane(src, dw, dh, ch)
            e = oracle.rescale_plane(src, dw, dh, ch)
            assert (o == e).all(), (ch, w, h, dw, dh, np.argwhere(o != e)[:3])
    # one-channel planes around the limits of the loop-free kernels: shrink ratios up to and past 4 (windows of 4, 5,
    # 6 samples: the last falls back), several 512-column units with a ragged last one, heights off the 32-row
    # groups, near-identity ratios, expansion over several units, a batch of several geometries of one direction
    shrink = [(1000, 64, 250, 16), (1001, 70, 250, 19), (1100, 33, 250, 31), (3000, 100, 1500, 50),
              (2051, 67, 1027, 33), (640, 640, 639, 161), (1536, 45, 1535, 44)]
    wide = [(1300, 40, 250, 9), (1000, 64, 250, 16), (900, 600, 300, 100)]    # a 6-sample window sends the whole batch to the streamed kernel
    expand = [(700, 300, 1500, 650), (300, 200, 301, 201), (513, 40, 2049, 161), (64, 9, 1100, 70)]
    for group in (shrink, wide, expand):
        planes = [rng.integers(0, 256, (h, w), dtype=np.uint8) for (w, h, dw, dh) in group]
        outs = R.RescalePlanesBatch(planes, [(dw, dh) for (w, h, dw, dh) in group])
        for (w, h, dw, dh), src, o in zip(group, planes, outs):
            e = oracle.rescale_plane(src, dw, dh)
            assert (o == e).all(), (w, h, dw, dh, np.argwhere(o != e)[:3])
    y = rng.integers(0, 256, (90, 160), dtype=np.uint8)
    u = rng.integers(0, 256, (45, 80), dtype=np.uint8)
    v = rng.integers(0, 256, (45, 80), dtype=np.uint8)
    pic = P.Picture(Width=160, Height=90, Y=y, U=u, V=v)
    assert R.WebPPictureRescale(pic, 64, 0) == 1          # height follows the ratio
    assert (pic.Width, pic.Height) == (64, 36)
    assert (pic.Y == oracle.rescale_plane(y, 64, 36)).all() and (pic.U == oracle.rescale_plane(u, 32, 18)).all()
    assert (pic.V == oracle.rescale_plane(v, 32, 18)).all()
    assert R.WebPPictureRescale(P.Picture(Width=4, Height=4, Y=y[:4, :4], U=u[:2, :2], V=v[:2, :2]), 0, 0) == 0
    # the premultiplying branches (YUV420A, ARGB) against what the libwebp binary returned
    g = np.load(os.path.join(golden_dir, "rescale_alpha_libwebp160.npz"))
    for k in [str(k) for k in g["names"]]:
        h, w = g[f"{k}/y"].shape
        dh, dw = g[f"{k}/out_y"].shape
        pic = P.Picture(Width=w, Height=h, Y=g[f"{k}/y"], U=g[f"{k}/u"], V=g[f"{k}/v"], A=g[f"{k}/a"], ColorSpace=P.WEBP_YUV420A)
        assert R.WebPPictureRescale(pic, dw, dh) == 1
        for got, nm in ((pic.Y, "y"), (pic.U, "u"), (pic.V, "v"), (pic.A, "a")):
            assert (got == g[f"{k}/out_{nm}"]).all(), (k, nm)
        pic = P.Picture(Width=w, Height=h, UseARGB=True, ARGB=g[f"{k}/argb"], ARGBStride=w)
        assert R.WebPPictureRescale(pic, dw, dh) == 1
        assert (pic.ARGB == g[f"{k}/out_argb"]).all(), (k, "argb")
        assert (pic.Width, pic.Height, pic.ARGBStride) == (dw, dh, dw)


# ---- ApplyPalette: colour -> index, the forward half of colour indexing (row a33) -------------
def test_apply_palette_every_strategy_sizes_and_alignments(gpu):
    """b2_vp8l_apply_palette vs the oracle's restatement of ApplyPalette + VP8LBundleColorMap
    (vp8l_encoder.go:1093-1212, lossless_enc.go:293-312): palette sizes 2..256 (all four packings), every
    strategy of the reference, widths around the 4-pixel / word / segment boundaries, rows that miss the
    vector alignment, one mixed batch per launch."""
    from _palettes import palette_cases
    dsp = gpu["dsp"]
    rng = np.random.default_rng(21)
    widths = [1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 64, 127, 128, 129, 515, 2048, 2049, 4100]
    items, expect, strategies = [], [], set()
    for name, pal in palette_cases(rng):
        xb = oracle.palette_xbits(pal.size)
        for w in widths:
            h = 3 if w < 1000 else 2
            idx = rng.integers(0, pal.size, (h, w), dtype=np.uint8)
            img = pal[idx]
            out, st = oracle.apply_palette(img, pal, xb)
            strategies.add(st)
            items.append((img, pal, xb))
            expect.append(out)
    assert strategies == {0, 1, 2, 3, 4}
    got = dsp.ApplyPaletteBatch(items)
    for (img, pal, xb), e, g in zip(items, expect, got):
        assert g.shape == e.shape and (g == e).all(), (pal.size, xb, img.shape)


def test_apply_palette_pixels_outside_the_palette(gpu):
    """Where the reference defines an answer for a colour that is not in the palette (greedy: 3; hash
    tables: whatever the slot holds, 0xffff cut to a byte when empty) the device gives the same."""
    from _palettes import hash2_palette
    dsp = gpu["dsp"]
    rng = np.random.default_rng(22)
    pals = [rng.choice(1 << 32, 3, replace=False).astype(np.uint32),       # greedy
            ((np.arange(64, dtype=np.uint64) * 4) * 0x010101 | 0xff000000).astype(np.uint32),   # hash 0
            rng.choice(1 << 32, 12, replace=False).astype(np.uint32),      # hash 1 (checked below)
            hash2_palette()]                                               # hash 2
    items, expect = [], []
    for pal in pals:
        xb = oracle.palette_xbits(pal.size)
        img = rng.integers(0, 2**32, (5, 131), dtype=np.uint32)
        img[:, ::3] = pal[rng.integers(0, pal.size, img[:, ::3].shape)]
        out, st = oracle.apply_palette(img, pal, xb)
        assert st <= 3
        items.append((img, pal, xb))
        expect.append(out)
    for e, g in zip(expect, dsp.ApplyPaletteBatch(items)):
        assert (g == e).all()


def test_apply_palette_reference_signature_and_libwebp_streams(gpu, golden_dir):
    """dsp.ApplyPalette with the reference's argument list (padded src_stride), and the pin: on the
    colour-indexed streams the libwebp binary encoded, ApplyPalette of the decoded pixels is the packed
    index image that sits in the stream (tests/golden/vp8l_libwebp160.npz)."""
    dsp = gpu["dsp"]
    rng = np.random.default_rng(23)
    pal = rng.choice(1 << 32, 9, replace=False).astype(np.uint32)
    w, h, stride = 45, 6, 52
    buf = rng.integers(0, 2**32, (h, stride), dtype=np.uint32)
    buf[:, :w] = pal[rng.integers(0, 9, (h, w))]
    got = dsp.ApplyPalette(buf, stride, pal, 9, w, h, dsp.PaletteXBits(9))
    assert (got == oracle.apply_palette(buf[:, :w], pal, 1)[0]).all()
    from test_oracle_golden import vp8l_vectors
    sub = lambda v, b: (v + (1 << b) - 1) >> b  # noqa: E731
    n = 0
    for name, img, trs, expected in vp8l_vectors(golden_dir):
        cur = img
        for (t, b, tw, th, data) in reversed(trs):
            if t == 0:
                cur = oracle.predictor_inverse(cur, data.reshape(sub(th, b), sub(tw, b)), b)
            elif t == 1:
                cur = oracle.color_space_inverse(cur, data.reshape(sub(th, b), sub(tw, b)), b)
            elif t == 2:
                cur = oracle.add_green(cur)
            else:
                palette = oracle.expand_color_map(data, b)[:data.size]
                assert b == dsp.PaletteXBits(data.size)
                packed = dsp.ApplyPaletteBatch([(expected, palette, b)])[0]
                assert (packed == cur).all(), name
                n += 1
                cur = expected
    assert n >= 4


def test_apply_palette_full_size_round_trip(gpu):
    """8 x 2048x1024 per packing: synthetic packed indices -> colour-index inverse -> ApplyPalette gives the
    indices back (green bytes), bit for bit on every image."""
    from go_webp_b200.pipeline import PalettePipeline
    ctx = gpu["ctx"]
    for psize in (2, 4, 16, 256):
        p = PalettePipeline(ctx, 8, 2048, 1024, psize)
        p.fill_synthetic(0xB2000003)
        p.run_device()
        ctx.sync()
        idx = ctx.d2h(p.d_idx.ptr, p.pk * p.n).view(np.uint32)
        out = ctx.d2h(p.d_out.ptr, p.pk * p.n).view(np.uint32)
        assert ((out >> 24) == 0xff).all() and ((out & 0x00ff00ff) == 0).all()
        assert (((out ^ idx) >> 8) & 0xff == 0).all(), psize
        img, o = p.device_image(3)
        assert (o == oracle.apply_palette(img, p.palettes[3], p.xbits)[0]).all()
        p.close()


def test_16bit_order_is_library_wide(gpu):
    """one byte order for the 16-bit modes across a decode path.  The reference flow for
    MODE_rgbA_4444 is upsample (VP8YuvToRgba4444) -> EmitAlphaRGBA4444 (alpha nibble into byte 2i,
    io_dec.c.go:197-226) -> WebPApplyAlphaMultiply4444 (rg_byte_pos = 1 in the Go text): with the default
    context every step uses the Go text's order, with swap_16bit_csp = False every step uses the stock
    libwebp order, and lossy and lossless outputs of one colourspace share a layout."""
    P, D, A, ctx = gpu["picture"], gpu["dsp"], gpu["abi"], gpu["ctx"]
    rng = np.random.default_rng(41)
    w, h = 37, 9
    uw, uh = (w + 1) // 2, (h + 1) // 2
    y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    u = rng.integers(0, 256, (uh, uw), dtype=np.uint8)
    v = rng.integers(0, 256, (uh, uw), dtype=np.uint8)
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    argb = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
    try:
        for swap in (True, False):
            ctx.swap_16bit_csp = swap
            pic = P.Picture(Width=w, Height=h, ColorSpace=P.WEBP_YUV420A, Y=y, U=u, V=v, A=a)
            up = P.YUVAToRGBABatch([pic], A.MODE_rgbA_4444)[0]
            e_up = oracle.yuva_to_rgba(y, u, v, a, A.MODE_RGBA_4444, swap16=swap)
            assert (up == e_up).all()
            got = D.WebPApplyAlphaMultiply4444(up, w, h, 2 * w)       # default rg_byte_pos follows the context
            assert (got == oracle.apply_alpha_multiply_4444(e_up, 1 if swap else 0)).all()
            # the alpha nibble really is in the ba byte: byte 2i with the Go order, 2i + 1 with the stock one
            assert ((up[:, (0 if swap else 1)::2] & 0x0f) == (a >> 4)).all()
            # lossless output of the same colourspace, premultiplied in one go: same layout
            ll = D.VP8LConvertFromBGRA(argb, None, A.MODE_rgbA_4444)
            assert (ll == oracle.convert_from_bgra(argb, A.MODE_rgbA_4444, 0 if swap else 1)).all()
            for mode in (A.MODE_RGB_565, A.MODE_RGBA_4444):
                s1 = D.WebPSamplerProcessPlane(y, u, v, mode)
                assert (s1 == oracle.sampler_process_plane(y, u, v, mode, swap16=swap)).all()
    finally:
        ctx.swap_16bit_csp = True


def test_dithered_import_descriptor_ring_wrap(gpu):
    """the dithered import uploads its job table and then 660 B of jump polynomials per
    image before the kernels that read both run.  With ~5700-6350 images the pair crosses the end of
    the 4 MB descriptor ring of a fresh stream; the second upload must not land on the first."""
    from go_webp_b200.device import Context
    P = gpu["picture"]
    ctx = Context(0, n_streams=1)          # fresh ring: offset 0
    rng = np.random.default_rng(61)
    n, w, h = 6000, 4, 2
    srcs = [rng.integers(0, 256, (h, 4 * w), dtype=np.uint8) for _ in range(n)]
    pics = [P.Picture(Width=w, Height=h) for _ in range(n)]
    # one call with 6000 images (sliced inside the library), then 3000-image calls that walk the ring
    # over its end: the reserve wraps BEFORE the first of a call's two uploads
    for lo, hi in [(0, n), (0, 3000), (3000, 6000), (1000, 4000)]:
        assert P.ImportRGBABatch(pics[lo:hi], srcs[lo:hi], [4 * w] * (hi - lo), "BGRA", ctx=ctx, dithering=0.6, chunk=n) == 1
        for i in list(range(lo, hi, 397)) + [hi - 1]:
            y, u, v, a, ha = oracle.import_yuva_dithered(srcs[i], w, h, 0.6, "BGRA")
            assert (pics[i].Y == y).all() and (pics[i].U == u).all() and (pics[i].V == v).all(), (lo, hi, i)
        for p in pics[lo:hi:7]:
            p.memory_.host_valid = False      # next round must really come from the device again
    for p in pics:
        P.WebPPictureFree(p)
    ctx.close()


def test_predictor_inverse_chain_fused(gpu):
    """b2_vp8l_predictor_inverse_chain: [cross-colour inverse] -> predictor inverse -> [add-green] -> 4-byte
    output order in one pass of the wavefront (vp8l_dec.c.go:710-729, :771) vs the oracle applying the
    transforms one after the other; every combination, tile sizes of the two transforms that differ, widths
    that leave the vector path, several bands, images that end inside a band."""
    D, A = gpu["dsp"], gpu["abi"]
    rng = np.random.default_rng(88)
    sub = lambda v, b: (v + (1 << b) - 1) >> b  # noqa: E731
    perm = {A.MODE_BGRA: None, A.MODE_RGBA: [2, 1, 0, 3], A.MODE_ARGB: [3, 2, 1, 0]}
    imgs, chains, exps = [], [], []
    for (w, h, pb, cb) in [(64, 64, 3, 3), (150, 70, 3, 5), (131, 33, 2, 4), (260, 100, 5, 2), (1, 1, 2, 2), (7, 40, 4, 3),
                           (1024, 96, 5, 5), (36, 513, 3, 2)]:
        for use_cc in (True, False):
            for use_green in (True, False):
                px = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
                modes = (rng.integers(0, 16, (sub(h, pb), sub(w, pb)), dtype=np.uint32) << 8) | 0xff000000
                codes = rng.integers(0, 2**24, (sub(h, cb), sub(w, cb)), dtype=np.uint32) | 0xff000000
                chain = ([D.VP8LTransform(D.SUBTRACT_GREEN_TRANSFORM, 0, w, h)] if use_green else []) + \
                    [D.VP8LTransform(D.PREDICTOR_TRANSFORM, pb, w, h, modes)] + \
                    ([D.VP8LTransform(D.CROSS_COLOR_TRANSFORM, cb, w, h, codes)] if use_cc else [])
                e = oracle.color_space_inverse(px, codes, cb) if use_cc else px
                e = oracle.predictor_inverse(e, modes, pb)
                e = oracle.add_green(e) if use_green else e
                imgs.append(D.VP8LImage(px, w, h))
                chains.append(chain)
                exps.append(e)
    plans = [D._plan_chain(c) for c in chains]
    assert all(len(p) == 1 for p in plans), "each of these chains is one fused pass (or the plain predictor)"
    for mode in (A.MODE_BGRA, A.MODE_RGBA, A.MODE_ARGB):
        outs = D.VP8LInverseTransformBatch(imgs, chains, out_mode=mode)
        for o, e, im in zip(outs, exps, imgs):
            if perm[mode] is not None:
                e = np.ascontiguousarray(e.view(np.uint8).reshape(e.shape[0], -1, 4)[:, :, perm[mode]]).view(np.uint32).reshape(e.shape)
            assert (o == e).all(), (im.width, im.height, mode)


def test_import_unaligned_source_rows_raw_abi(gpu):
    """Source rows that are only 4-byte aligned (b2_import_fast.cu, UNAL kernels): every misalignment of
    the pointer (0, 4, 8, 12 bytes) and of the stride, widths around the strip and chunk boundaries,
    1-pixel images, with and without the alpha plane -- through the raw C ABI on two streams, several
    launches in a row (the row pairs are handed out by a ticket counter that the last warp of a launch
    clears for the next one)."""
    from go_webp_b200.device import Context
    A = gpu["abi"]
    ctx = Context(0, n_streams=2)
    rng = np.random.default_rng(0x0a11)
    widths = [1, 2, 3, 5, 31, 33, 124, 127, 128, 129, 131, 255, 256, 257, 260, 515, 1021, 1028]
    cases = []
    for k, w in enumerate(widths):
        h = 1 + (k * 5) % 9
        cases.append((w, h, 4 * (k % 4), 4 * ((k // 4) % 4)))       # pointer offset, stride padding
    cases += [(128, 3, 4, 0), (256, 2, 12, 0), (132, 4, 8, 4), (4, 1, 12, 12)]
    al = lambda v, a=256: (v + a - 1) // a * a  # noqa: E731
    lay, tot = [], dict(src=0, y=0, uv=0)
    for (w, h, off, pad) in cases:
        stride = 4 * w + pad
        ys, uvs = al(w, 4), al((w + 1) // 2, 2)
        lay.append((tot["src"] + off, stride, tot["y"], ys, tot["uv"], uvs))
        tot["src"] += al(off + stride * h + 16)
        tot["y"] += al(ys * h)
        tot["uv"] += al(uvs * ((h + 1) // 2))
    host_src = np.zeros(tot["src"], np.uint8)
    srcs = []
    for (w, h, off, pad), (so, stride, *_r) in zip(cases, lay):
        img = rng.integers(0, 256, (h, stride), dtype=np.uint8)
        img[:, 3:4 * w:4][rng.random((h, w)) < 0.6] = 255
        host_src[so:so + stride * h] = img.reshape(-1)
        srcs.append(img)
    d_src, d_y, d_a = ctx.alloc(tot["src"]), ctx.alloc(tot["y"]), ctx.alloc(tot["y"])
    d_u, d_v, d_f = ctx.alloc(tot["uv"]), ctx.alloc(tot["uv"]), ctx.alloc(al(4 * len(cases)))
    ctx.h2d(d_src.ptr, host_src)
    jobs = [A.ImportJob(d_src.ptr + so, d_y.ptr + yo, d_u.ptr + uo, d_v.ptr + uo, d_a.ptr + yo, stride, ys, uvs, ys, w, h)
            for (w, h, off, pad), (so, stride, yo, ys, uo, uvs) in zip(cases, lay)]
    arr = (A.ImportJob * len(jobs))(*jobs)
    for rep in range(4):
        stream = rep % 2
        order, flags = (A.ORDER_RGBA, A.IMPORT_ALPHA) if rep < 3 else (A.ORDER_BGRA, 0)
        for buf, nb in ((d_y, tot["y"]), (d_a, tot["y"]), (d_u, tot["uv"]), (d_v, tot["uv"])):
            ctx.h2d(buf.ptr, np.full(nb, 0x5a, np.uint8))
        ctx._chk(ctx.lib.b2_rgba_to_yuv420a(ctx.h, stream, arr, len(jobs), order, flags, d_f.ptr))
        ctx.sync()
        Y, U, V, Aa = (ctx.d2h(b.ptr, nb) for b, nb in ((d_y, tot["y"]), (d_u, tot["uv"]), (d_v, tot["uv"]), (d_a, tot["y"])))
        fl = ctx.d2h(d_f.ptr, 4 * len(cases)).view(np.int32)
        for i, ((w, h, off, pad), (so, stride, yo, ys, uo, uvs), src) in enumerate(zip(cases, lay, srcs)):
            name = "RGBA" if rep < 3 else "BGRX"
            y, u, v, a, ha = oracle.import_yuva(src, w, h, name, stride=stride)
            uh, uw = (h + 1) // 2, (w + 1) // 2
            got_y = Y[yo:yo + ys * h].reshape(h, ys)
            assert (got_y[:, :w] == y).all(), (rep, i, w, h, off, pad)
            assert (got_y[:, w:] == 0x5a).all(), (rep, i, "wrote past the end of a Y row")
            assert (U[uo:uo + uvs * uh].reshape(uh, uvs)[:, :uw] == u).all(), (rep, i, w, h, off, pad)
            assert (V[uo:uo + uvs * uh].reshape(uh, uvs)[:, :uw] == v).all(), (rep, i, w, h, off, pad)
            if rep < 3:
                assert bool(fl[i]) == bool(ha), (rep, i)
                assert (Aa[yo:yo + ys * h].reshape(h, ys)[:, :w] == src[:, 3:4 * w:4]).all(), (rep, i)
    for b in (d_src, d_y, d_a, d_u, d_v, d_f):
        b.free()
    ctx.close()


def test_vp8l_flat_items_odd_widths_and_tile_edges(gpu):
    """Flat VP8L kernels (2048-pixel items walked by additions): widths that are not multiples of 4, narrower
    than a 128-pixel step, images of less than one item and of several, every tile size, in one batch."""
    D = gpu["dsp"]
    rng = np.random.default_rng(0xf1a7)
    sizes = [(1, 1), (3, 700), (5, 5), (63, 40), (127, 33), (129, 17), (130, 64), (257, 9), (1021, 7), (2049, 3), (4099, 2), (333, 333)]
    for bits in (2, 3, 5, 9):
        imgs, chains = [], []
        for (w, h) in sizes:
            imgs.append(rng.integers(0, 2**32, (h, w), dtype=np.uint32))
            codes = rng.integers(0, 2**24, (D.VP8LSubSampleSize(h, bits), D.VP8LSubSampleSize(w, bits)), dtype=np.uint32) | 0xff000000
            chains.append([D.VP8LTransform(D.SUBTRACT_GREEN_TRANSFORM, 0, w, h), D.VP8LTransform(D.CROSS_COLOR_TRANSFORM, bits, w, h, codes)])
        res = D.VP8LInverseTransformBatch([D.VP8LImage(a, w, h) for a, (w, h) in zip(imgs, sizes)], chains)
        for r, a, ch, sz in zip(res, imgs, chains, sizes):
            assert (r == oracle.add_green(oracle.color_space_inverse(a, ch[1].data, bits))).all(), (bits, sz)
