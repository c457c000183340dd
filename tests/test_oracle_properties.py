"""Algebraic / exhaustive checks of the oracle (CPU only): the pins available
for part 3, which neither the reference's tests nor the libwebp binary expose
(SURVEY.md 4, 8c)."""
import ctypes as C

import numpy as np
import pytest

import oracle


def test_yuv_to_rgb_exhaustive_against_formula():
    # pkg/color/yuv/conversion.go:28-48 over the whole 256^3 domain via the
    # formula evaluated in numpy; the oracle is sampled on a dense lattice.
    L = oracle.lib()
    y, u, v = np.meshgrid(np.arange(0, 256, 5), np.arange(0, 256, 3),
                          np.arange(0, 256, 7), indexing="ij")

    def clip8(x):
        return np.where((x & ~16383) == 0, x >> 6, np.where(x < 0, 0, 255))
    r = clip8(((y * 19077) >> 8) + ((v * 26149) >> 8) - 14234)
    g = clip8(((y * 19077) >> 8) - ((u * 6419) >> 8) - ((v * 13320) >> 8) + 8708)
    b = clip8(((y * 19077) >> 8) + ((u * 33050) >> 8) - 17685)
    for yy, uu, vv, rr, gg, bb in zip(y.ravel()[::97], u.ravel()[::97], v.ravel()[::97],
                                      r.ravel()[::97], g.ravel()[::97], b.ravel()[::97]):
        assert L.orc_YUVToR(int(yy), int(vv)) == rr
        assert L.orc_YUVToG(int(yy), int(uu), int(vv)) == gg
        assert L.orc_YUVToB(int(yy), int(uu)) == bb


def test_uv_never_clips_for_gamma_domain_inputs():
    # LinearToGamma returns <= 1020 (l2g[32]=255 -> (255*512+64)>>7), so the
    # ClipUV branches of conversion.go:51-54 are dead for the import path; the
    # CUDA kernel relies on this to drop the clamp.
    L = oracle.lib()
    assert L.orc_LinearToGamma(4 * 4095, 0) == 1020
    assert L.orc_LinearToGamma(2 * 4095, 1) == 1020
    for r in (0, 1020):
        for g in (0, 1020):
            for b in (0, 1020):
                for f in (L.orc_RGBToU, L.orc_RGBToV):
                    raw = f(r, g, b, 1 << 17)
                    assert 16 <= raw <= 240


def test_color_transform_delta_exhaustive():
    L = oracle.lib()
    for m in range(-128, 128):
        for c in range(-128, 128, 3):
            assert L.orc_ColorTransformDelta(m, c) == (m * c) >> 5


@pytest.mark.parametrize("bits", [2, 3, 5, 9])
@pytest.mark.parametrize("wh", [(100, 70), (33, 17), (4, 4), (1, 1), (129, 5)])
def test_subgreen_crosscolor_roundtrip(bits, wh):
    w, h = wh
    rng = np.random.default_rng(bits * 1000 + w)
    x = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
    tw, th = (w + (1 << bits) - 1) >> bits, (h + (1 << bits) - 1) >> bits
    codes = rng.integers(0, 2**24, (th, tw), dtype=np.uint32) | 0xff000000
    fwd = oracle.color_space_forward(oracle.sub_green(x), codes, bits)
    back = oracle.add_green(oracle.color_space_inverse(fwd, codes, bits))
    assert (back == x).all()
    # numpy restatement of SURVEY Appendix A.3 (independent of the C loops)
    yy, xx = np.mgrid[0:h, 0:w]
    code = codes[yy >> bits, xx >> bits]

    def s8(a):
        return ((a & 0xff) ^ 0x80).astype(np.int32) - 0x80
    src = fwd.astype(np.int64)
    g = s8(src >> 8)
    r = (src >> 16) & 0xff
    b = src & 0xff
    nr = (r + ((s8(code) * g) >> 5)) & 0xff
    nb = (b + ((s8(code >> 8) * g) >> 5) + ((s8(code >> 16) * s8(nr)) >> 5)) & 0xff
    exp = (src & 0xff00ff00) | (nr << 16) | nb
    assert (oracle.color_space_inverse(fwd, codes, bits) == exp.astype(np.uint32)).all()


@pytest.mark.parametrize("bits", [0, 1, 2, 3])
def test_bundle_then_index_inverse(bits):
    rng = np.random.default_rng(bits)
    ncol = 1 << (8 >> bits)
    pal = rng.integers(0, 2**32, ncol, dtype=np.uint32)
    for w in list(range(1, 41)) + [255, 256, 257]:
        h = 3
        idx = rng.integers(0, ncol, (h, w), dtype=np.uint8)
        packed = oracle.bundle_color_map(idx, bits)
        assert packed.shape[1] == (w + (1 << bits) - 1) >> bits
        out = oracle.color_index_inverse(packed, pal, bits, w)
        assert (out == pal[idx]).all()
        # alpha-plane variant (lossless.go:461-505): u8 source = the green bytes
        a_src = ((packed >> 8) & 0xff).astype(np.uint8)
        a_out = oracle.color_index_inverse_alpha(a_src, pal, bits, w)
        assert (a_out == ((pal[idx] >> 8) & 0xff)).all()


def test_expand_color_map():
    rng = np.random.default_rng(9)
    for ncol, bits in [(2, 3), (3, 2), (4, 2), (5, 1), (16, 1), (17, 0), (256, 0)]:
        deltas = rng.integers(0, 2**32, ncol, dtype=np.uint32)
        out = oracle.expand_color_map(deltas, bits)
        assert out.size == 1 << (8 >> bits)
        by = deltas.view(np.uint8).reshape(ncol, 4).astype(np.uint32)
        exp = (np.cumsum(by, axis=0) & 0xff).astype(np.uint8).view(np.uint32).ravel()
        assert (out[:ncol] == exp).all() and (out[ncol:] == 0).all()


@pytest.mark.parametrize("bits", [2, 4])
def test_predictor_roundtrip(bits):
    rng = np.random.default_rng(bits)
    for (w, h) in [(1, 1), (5, 1), (1, 5), (37, 29), (64, 64)]:
        x = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
        tw, th = (w + (1 << bits) - 1) >> bits, (h + (1 << bits) - 1) >> bits
        modes = (rng.integers(0, 16, (th, tw), dtype=np.uint32) << 8) | 0xff000000
        res = oracle.predictor_forward(x, modes, bits)
        assert (oracle.predictor_inverse(res, modes, bits) == x).all()


def test_synth_generator_matches_spec():
    # SURVEY 8d generator, restated in numpy
    def lowbias32(x):
        x = x.astype(np.uint64)
        x ^= x >> 16; x = (x * 0x7feb352d) & 0xffffffff
        x ^= x >> 15; x = (x * 0x846ca68b) & 0xffffffff
        x ^= x >> 16
        return x.astype(np.uint32)
    seed, img, plane = 0xB2000001, 7, 2
    got = oracle.synth_u32(seed, img, plane, 13, 5)
    yy, xx = np.mgrid[0:5, 0:13].astype(np.uint64)
    key = (np.uint64(seed) ^ np.uint64((img * 0x9E3779B1) & 0xffffffff)
           ^ np.uint64((plane * 0x85EBCA77) & 0xffffffff)
           ^ ((yy * 0xC2B2AE3D) & 0xffffffff) ^ ((xx * 0x27D4EB2F) & 0xffffffff))
    assert (got == lowbias32(key)).all()


# ---- oracle_post.c: identities that pin the restatement -----------------
def test_post_conversions_identities():
    rng = np.random.default_rng(17)
    argb = rng.integers(0, 2**32, (6, 37), dtype=np.uint32)
    a, r, g, b = (argb >> 24) & 0xff, (argb >> 16) & 0xff, (argb >> 8) & 0xff, argb & 0xff
    bgra = oracle.convert_from_bgra(argb, oracle.MODE_BGRA)
    assert (bgra.view(np.uint32) == argb).all()                                  # CopyOrSwap(…, 1) on little endian
    rgba = oracle.convert_from_bgra(argb, oracle.MODE_RGBA)
    assert (rgba[:, 0::4] == r).all() and (rgba[:, 3::4] == a).all()
    rgb = oracle.convert_from_bgra(argb, oracle.MODE_RGB)
    assert (rgb[:, 0::3] == r).all() and (rgb[:, 1::3] == g).all() and (rgb[:, 2::3] == b).all()
    # premultiply == (int)(x * a / 255.) for every 8-bit pair (alpha_processing.go:96-99)
    x, al = np.meshgrid(np.arange(256, dtype=np.uint32), np.arange(256, dtype=np.uint32))
    px = np.zeros((256, 256 * 4), np.uint8)
    px[:, 0::4] = x; px[:, 1::4] = x; px[:, 2::4] = x; px[:, 3::4] = al
    pm = oracle.apply_alpha_multiply(px, 0)
    want = np.where(al == 255, x, (x * al) // 255)
    assert (pm[:, 0::4] == want).all() and (pm[:, 3::4] == al).all()
    # rgbA == RGBA then premultiply; 16-bit orders are byte swaps of each other
    assert (oracle.convert_from_bgra(argb, oracle.MODE_rgbA) == oracle.apply_alpha_multiply(rgba, 0)).all()
    for m in (oracle.MODE_RGBA_4444, oracle.MODE_RGB_565, oracle.MODE_rgbA_4444):
        s0, s1 = oracle.convert_from_bgra(argb, m, 0), oracle.convert_from_bgra(argb, m, 1)
        assert (s0[:, 0::2] == s1[:, 1::2]).all() and (s0[:, 1::2] == s1[:, 0::2]).all()
    # point sampling == 4:4:4 conversion of the replicated chroma; WebPMultRow(255 alpha) is the identity
    y = rng.integers(0, 256, (5, 9), dtype=np.uint8)
    u = rng.integers(0, 256, (3, 5), dtype=np.uint8)
    v = rng.integers(0, 256, (3, 5), dtype=np.uint8)
    rep = lambda c: np.repeat(np.repeat(c, 2, axis=0), 2, axis=1)[:5, :9]  # noqa: E731
    for m in range(11):
        assert (oracle.sampler_process_plane(y, u, v, m) == oracle.yuv444_to_rgb(y, rep(u), rep(v), m)).all()
    assert (oracle.mult_rows(y, np.full_like(y, 255), 0) == y).all()
    assert (oracle.mult_rows(y, np.zeros_like(y), 0) == 0).all()
    # multiply then inverse-multiply recovers premultiplied values up to the reference's own rounding
    m1 = oracle.mult_argb_rows(argb | 0x80000000, 0)
    assert ((m1 >> 24) == ((argb | 0x80000000) >> 24)).all()


def test_apply_palette_strategies_and_round_trip():
    """ApplyPalette (vp8l_encoder.go:1093-1212): every strategy the reference can pick gives the index
    of the colour, bundled as VP8LBundleColorMap does; colour-index inverse undoes it."""
    from _palettes import palette_cases
    rng = np.random.default_rng(11)
    seen = set()
    for name, pal in palette_cases(rng):
        P = pal.size
        xb = oracle.palette_xbits(P)
        for (w, h) in [(1, 1), (7, 3), (37, 13), (64, 2), (129, 5)]:
            idx = rng.integers(0, P, (h, w), dtype=np.uint8)
            img = pal[idx]
            out, st = oracle.apply_palette(img, pal, xb)
            seen.add(st)
            assert (out == oracle.bundle_color_map(idx, xb)).all(), (name, w, h)
            full = np.zeros(1 << (8 >> xb), np.uint32)
            full[:P] = pal
            assert (oracle.color_index_inverse(out, full, xb, w) == img).all()
    assert seen == {0, 1, 2, 3, 4}, seen
