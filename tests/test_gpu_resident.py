"""Device-resident pictures and the host-staged batch entry points (north_star:
"pkg/picture gains device-resident Y/U/V/A and ARGB planes with pinned-host staging"):
chained calls must not touch the host in between, results stay bit-exact vs the oracle."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

SIZES = [(64, 64), (65, 31), (127, 129), (1, 1), (2, 3), (33, 16), (256, 8), (515, 7), (1920, 24)]


@pytest.fixture(scope="module")
def gpu():
    from go_webp_b200 import _abi, device, dsp, picture, picture_tools
    if device.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return dict(abi=_abi, dsp=dsp, P=picture, T=picture_tools, ctx=picture.default_context())


def _rgba(rng, w, h, kind):
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    if kind == "opaque":
        rgba[..., 3] = 255
    elif kind == "sparse":
        rgba[..., 3][rng.random((h, w)) < 0.7] = 255
    elif kind == "binary":
        rgba[..., 3] = np.where(rng.random((h, w)) < 0.5, 0, 255)
    return rgba


def test_import_makes_resident_pictures_and_reuses_memory(gpu):
    P, ctx = gpu["P"], gpu["ctx"]
    rng = np.random.default_rng(1)
    kinds = ["opaque", "sparse", "binary"]
    srcs = [_rgba(rng, w, h, kinds[i % 3]) for i, (w, h) in enumerate(SIZES)]
    pics = [P.Picture(Width=w, Height=h) for (w, h) in SIZES]
    assert P.ImportRGBABatch(pics, srcs, [4 * w for (w, h) in SIZES], "RGBA", chunk=2) == 1
    blocks = {id(p.memory_.block) for p in pics}
    assert len(blocks) == 1, "one device + one pinned block for the whole batch"
    for pic, src, (w, h) in zip(pics, srcs, SIZES):
        y, u, v, a, ha = oracle.import_yuva(src, w, h, "RGBA")
        assert pic.resident() and pic.memory_.host_valid
        assert pic.ColorSpace == (P.WEBP_YUV420A if ha else P.WEBP_YUV420)
        assert pic.YStride % 16 == 0 and pic.YStride >= w
        assert (pic.Y == y).all() and (pic.U == u).all() and (pic.V == v).all()
        assert (pic.A == a).all() if ha else pic.A is None
        assert not pic.Y.flags.writeable
    # a second import into the same pictures keeps the block (no reallocation)
    mem = [p.memory_ for p in pics]
    srcs2 = [_rgba(rng, w, h, "sparse") for (w, h) in SIZES]
    assert P.ImportRGBABatch(pics, srcs2, [4 * w for (w, h) in SIZES], "RGBA") == 1
    assert all(p.memory_ is m for p, m in zip(pics, mem))
    for pic, src, (w, h) in zip(pics, srcs2, SIZES):
        y, u, v, a, ha = oracle.import_yuva(src, w, h, "RGBA")
        assert (pic.Y == y).all() and (pic.U == u).all() and (pic.V == v).all()
        assert (pic.A == a).all() if ha else pic.A is None
    # freeing every picture releases the block
    import weakref
    ref = weakref.ref(pics[0].memory_.block)
    del mem
    for p in pics:
        P.WebPPictureFree(p)
    import gc
    gc.collect()
    assert ref() is None and pics[0].Y is None


def test_alpha_plane_comes_back_only_with_alpha(gpu):
    """conversions.go:56-58: opaque pictures have no alpha plane -> its bytes never cross PCIe."""
    P, ctx = gpu["P"], gpu["ctx"]
    rng = np.random.default_rng(2)
    w, h, n = 320, 64, 8
    srcs = [_rgba(rng, w, h, "opaque" if i % 2 else "sparse") for i in range(n)]
    pics = [P.Picture(Width=w, Height=h) for _ in range(n)]
    P.AllocResidentBatch(pics, "yuva", ctx)
    s0 = ctx.copy_stats()
    assert P.ImportRGBABatch(pics, srcs, [4 * w] * n, "RGBA", chunk=3) == 1
    s1 = ctx.copy_stats()
    yuv = w * h + 2 * (w // 2) * (h // 2)
    assert s1["h2d_bytes"] - s0["h2d_bytes"] == n * 4 * w * h
    assert s1["d2h_bytes"] - s0["d2h_bytes"] == n * yuv + (n // 2) * w * h + 4 * n   # A for the 4 pictures with alpha only, + flags
    for i, pic in enumerate(pics):
        assert (pic.A is None) == bool(i % 2)


def test_three_chained_calls_one_upload_one_download(gpu):
    """import -> WebPCleanupTransparentArea -> fancy upsample on resident pictures
    (webp_encoding.go:49-72): the sources go up once, the final pixels come down once."""
    P, T, ctx, abi = gpu["P"], gpu["T"], gpu["ctx"], gpu["abi"]
    rng = np.random.default_rng(3)
    sizes = [(64, 64), (72, 40), (130, 66), (16, 16)]
    srcs = [_rgba(rng, w, h, "binary") for (w, h) in sizes]
    for s in srcs:
        s[:16, :16, 3] = 0          # one fully transparent 8x8 block run at least
    pics = [P.Picture(Width=w, Height=h) for (w, h) in sizes]
    s0, l0 = ctx.copy_stats(), ctx.launches
    assert P.ImportRGBABatch(pics, srcs, [4 * w for (w, h) in sizes], "RGBA", download=False) == 1
    T.WebPCleanupTransparentAreaBatch(pics, download=False)
    outs = P.YUVAToRGBABatch(pics, abi.MODE_RGBA)
    s1, l1 = ctx.copy_stats(), ctx.launches
    assert outs is not None
    assert s1["h2d_bytes"] - s0["h2d_bytes"] == sum(4 * w * h for (w, h) in sizes), "only the sources go up"
    assert s1["d2h_bytes"] - s0["d2h_bytes"] == sum(4 * w * h for (w, h) in sizes) + 4 * len(sizes), \
        "only the final pixels (+ the has-alpha flags) come down"
    assert l1 - l0 >= 3
    for pic, src, out, (w, h) in zip(pics, srcs, outs, sizes):
        y, u, v, a, ha = oracle.import_yuva(src, w, h, "RGBA")
        assert ha
        y, u, v = oracle.cleanup_transparent_yuva(y, u, v, a)
        exp = oracle.yuva_to_rgba(y, u, v, a, oracle.MODE_RGBA)
        assert (out == exp).all()
        # reading a plane of a picture whose host mirror is stale fetches it
        assert not pic.memory_.host_valid
        assert (pic.Y == y).all() and pic.memory_.host_valid and (pic.A == a).all()


def test_assigning_a_plane_makes_the_picture_user_owned(gpu):
    P, abi = gpu["P"], gpu["abi"]
    rng = np.random.default_rng(4)
    w, h = 96, 34
    src = _rgba(rng, w, h, "opaque")
    pic = P.Picture(Width=w, Height=h)
    assert P.ImportRGBABatch([pic], [src], [4 * w], "RGBA", download=False) == 1
    y2 = rng.integers(0, 256, (h, w), dtype=np.uint8)
    pic.Y = y2                      # user data: residency is dropped, U and V were fetched first
    assert pic.memory_ is None
    _, u, v, _, _ = oracle.import_yuva(src, w, h, "RGBA")
    assert (pic.U == u).all() and (pic.V == v).all()
    out = P.YUVAToRGBABatch([pic], abi.MODE_BGRA)[0]
    assert (out == oracle.yuva_to_rgba(y2, u, v, None, oracle.MODE_BGRA)).all()


def test_yuva_to_argb_resident_then_argb_tools(gpu):
    P, T = gpu["P"], gpu["T"]
    rng = np.random.default_rng(5)
    w, h = 200, 50
    src = _rgba(rng, w, h, "binary")
    pic = P.Picture(Width=w, Height=h)
    assert P.ImportRGBABatch([pic], [src], [4 * w], "RGBA", download=False) == 1
    assert P.WebPPictureYUVAToARGB(pic, download=False) == 1
    assert pic.UseARGB and pic.resident("argb") and not pic.memory_argb_.host_valid
    T.WebPReplaceTransparentPixelsBatch([pic], 0x123456, download=False)
    y, u, v, a, ha = oracle.import_yuva(src, w, h, "RGBA")
    argb = oracle.yuva_to_rgba(y, u, v, a, oracle.MODE_BGRA).view(np.uint32).reshape(h, w)
    exp = np.where((argb >> 24) == 0, np.uint32(0x123456), argb)
    assert (pic.ARGB == exp).all()


def test_staged_vp8l_chain_crosses_pcie_twice(gpu):
    dsp, ctx = gpu["dsp"], gpu["ctx"]
    rng = np.random.default_rng(6)
    w, h, bits = 150, 70, 3
    px = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
    tw, th = dsp.VP8LSubSampleSize(w, bits), dsp.VP8LSubSampleSize(h, bits)
    codes = rng.integers(0, 2**24, (th, tw), dtype=np.uint32) | 0xff000000
    modes = (rng.integers(0, 14, (th, tw), dtype=np.uint32) << 8) | 0xff000000
    chain = [dsp.VP8LTransform(dsp.PREDICTOR_TRANSFORM, bits, w, h, modes),
             dsp.VP8LTransform(dsp.CROSS_COLOR_TRANSFORM, bits, w, h, codes),
             dsp.VP8LTransform(dsp.SUBTRACT_GREEN_TRANSFORM, 0, w, h)]
    s0 = ctx.copy_stats()
    out = dsp.VP8LInverseTransformBatch([dsp.VP8LImage(px, w, h)] * 3, [chain, chain[1:], chain[:1]])
    s1 = ctx.copy_stats()
    e_full = oracle.predictor_inverse(oracle.color_space_inverse(oracle.add_green(px), codes, bits), modes, bits)
    assert (out[0] == e_full).all()
    assert (out[1] == oracle.color_space_inverse(oracle.add_green(px), codes, bits)).all()
    assert (out[2] == oracle.predictor_inverse(px, modes, bits)).all()
    assert s1["d2h_bytes"] - s0["d2h_bytes"] == 3 * px.nbytes
    assert s1["h2d_bytes"] - s0["h2d_bytes"] == 3 * px.nbytes + 2 * modes.nbytes + 2 * codes.nbytes


def test_pipeline_and_drop_in_entry_point_agree(gpu):
    """bench.py times picture.ImportRGBABatch; ImportPipeline.run is the bench-shaped twin."""
    from go_webp_b200.pipeline import ImportPipeline
    P, ctx, abi = gpu["P"], gpu["ctx"], gpu["abi"]
    n, w, h = 6, 1920, 30
    pipe = ImportPipeline(ctx, n, w, h, "RGBA", chunk=4, host=True)
    pipe.fill_synthetic(0xB2000001, 0, abi.SYNTH_SPARSE, to_host=True)
    pipe.run()
    srcs = [pipe.src_view(i) for i in range(n)]
    pics = [P.Picture(Width=w, Height=h) for _ in range(n)]
    assert P.ImportRGBABatch(pics, srcs, [4 * w] * n, "RGBA", chunk=4) == 1
    for i, pic in enumerate(pics):
        y, u, v, a, ha = pipe.planes(i)
        assert ha and (pic.Y == y).all() and (pic.U == u).all() and (pic.V == v).all() and (pic.A == a).all()
        e = oracle.import_yuva(np.ascontiguousarray(srcs[i]), w, h, "RGBA")
        assert (pic.Y == e[0]).all() and (pic.A == e[3]).all()
    pipe.close()
