"""A second witness for the graded import mode.

The translucent 2x2 blocks of the RGBA -> YUV420A import divide exactly in the go-webp text
(`DIVIDE_BY_ALPHA(sum, a) = 4 * sum / a`, pkg/libwebp/dsp/yuv.go:231) while the libwebp binary
behind the golden vectors uses an inverse-alpha table, so that branch was pinned only by the
oracle's own restatement.  This file restates yuv.go:155-302 + pkg/color/yuv/conversion.go:51-69
a second time, in numpy int64, written from the Go text and NOT from oracle/oracle.c, and checks
the oracle against it on > 10^6 random translucent blocks, every one of the 1019 possible
divisors, odd widths (the 2-pixel last column, step = 0) and odd heights (the duplicated last
row, rgb_stride = 0).  CPU only."""
import numpy as np

import oracle

# yuv.go:156-162
GAMMA_FIX, GAMMA_TAB_FIX = 12, 7
GAMMA_TAB_SIZE = 1 << (GAMMA_FIX - GAMMA_TAB_FIX)
K_GAMMA = 0.80
K_GAMMA_SCALE = (1 << GAMMA_FIX) - 1
K_GAMMA_TAB_SCALE = 1 << GAMMA_TAB_FIX
K_GAMMA_TAB_ROUNDER = (1 << GAMMA_TAB_FIX) >> 1
# pkg/color/yuv/constants.go
YUV_FIX = 16
YUV_HALF = 1 << (YUV_FIX - 1)


def gamma_tables():
    """yuv.go:168-181 (scale and norm are floating-point quotients: SURVEY Appendix B-1)."""
    scale = float(1 << GAMMA_TAB_FIX) / K_GAMMA_SCALE
    norm = 1.0 / 255.0
    g2l = np.array([int(pow(norm * v, K_GAMMA) * K_GAMMA_SCALE + 0.5) for v in range(256)], np.int64)
    l2g = np.array([int(255.0 * pow(scale * v, 1.0 / K_GAMMA) + 0.5) for v in range(GAMMA_TAB_SIZE + 1)], np.int64)
    return g2l, l2g


G2L, L2G = gamma_tables()


def linear_to_gamma(base_value, shift):
    """Interpolate + LinearToGamma, yuv.go:187-202."""
    v = base_value << shift
    tab_pos = v >> (GAMMA_TAB_FIX + 2)
    x = v & ((K_GAMMA_TAB_SCALE << 2) - 1)
    v0, v1 = L2G[tab_pos], L2G[tab_pos + 1]
    y = v1 * x + v0 * ((K_GAMMA_TAB_SCALE << 2) - x)
    return (y + K_GAMMA_TAB_ROUNDER) >> GAMMA_TAB_FIX


def clip_uv(uv, rounding):
    """conversion.go:51-54."""
    uv = (uv + rounding + (128 << (YUV_FIX + 2))) >> (YUV_FIX + 2)
    return np.clip(uv, 0, 255)


def accumulate_block(px):
    """WebPAccumulateRGBA for 2x2 blocks (yuv.go:243-268): px int64 [n, 4 (pixels: top-left, top-right,
    bottom-left, bottom-right), 4 (r, g, b, a)] -> r, g, b, a sums as RGBToU/V take them."""
    a = px[:, :, 3].sum(axis=1)                                        # SUM4ALPHA
    lin = G2L[px[:, :, :3]]                                            # GammaToLinear of every sample
    plain = linear_to_gamma(lin.sum(axis=1), 0)                        # SUM4: opaque or fully transparent
    den = np.where(a == 0, 1, a)
    wsum = (px[:, :, 3:4] * lin).sum(axis=1)                           # LinearToGammaWeighted, :233-241
    weighted = linear_to_gamma((4 * wsum) // den[:, None], 0)          # DIVIDE_BY_ALPHA: exact integer division
    use_plain = ((a == 4 * 0xff) | (a == 0))[:, None]
    return np.where(use_plain, plain, weighted), a


def accumulate_pair(px):
    """The odd last column (yuv.go:269-284): two pixels, top and bottom; px [n, 2, 4].  a = 2 * SUM2ALPHA; the
    weighted form runs with step = 0, i.e. every pixel of the column counted twice."""
    a = 2 * px[:, :, 3].sum(axis=1)
    lin = G2L[px[:, :, :3]]
    plain = linear_to_gamma(lin.sum(axis=1), 1)                        # SUM2
    den = np.where(a == 0, 1, a)
    wsum = 2 * (px[:, :, 3:4] * lin).sum(axis=1)
    weighted = linear_to_gamma((4 * wsum) // den[:, None], 0)
    use_plain = ((a == 4 * 0xff) | (a == 0))[:, None]
    return np.where(use_plain, plain, weighted), a


def import_uv(rgba):
    """U and V planes of picture.ImportYUVAFromRGBA's non-dithered branch for an RGBA image [h, w, 4]
    (conversions.go:106-121 -> yuv.go:304-382): row pairs, the last row of an odd height doubled."""
    h, w, _ = rgba.shape
    img = rgba.astype(np.int64)
    if h & 1:
        img = np.concatenate([img, img[-1:]], axis=0)                  # WebPImportYUVAFromRGBALastLine: rgb_stride = 0
    hh = img.shape[0] // 2
    we = w & ~1
    blocks = img[:, :we].reshape(hh, 2, we // 2, 2, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 4, 4)
    rgb, _ = accumulate_block(blocks)
    rgb = rgb.reshape(hh, we // 2, 3)
    if w & 1:
        pairs = img[:, w - 1].reshape(hh, 2, 4)
        last, _ = accumulate_pair(pairs)
        rgb = np.concatenate([rgb, last.reshape(hh, 1, 3)], axis=1)
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    u = clip_uv(-9719 * r - 19081 * g + 28800 * b, YUV_HALF << 2)      # WebPConvertRGBA32ToUV, yuv.go:142-153
    v = clip_uv(28800 * r - 24116 * g - 4684 * b, YUV_HALF << 2)
    return u.astype(np.uint8), v.astype(np.uint8)


def import_y(rgba):
    img = rgba.astype(np.int64)
    luma = 16839 * img[..., 0] + 33059 * img[..., 1] + 6420 * img[..., 2]
    return ((luma + YUV_HALF + (16 << YUV_FIX)) >> YUV_FIX).astype(np.uint8)   # RGBToY, conversion.go:56-59


def _check(rgba):
    h, w, _ = rgba.shape
    y, u, v, a, ha = oracle.import_yuva(rgba, w, h, "RGBA")            # exact divide is the oracle's default
    eu, ev = import_uv(rgba)
    assert (y == import_y(rgba)).all()
    assert (u == eu).all() and (v == ev).all(), (w, h)
    assert (a == rgba[..., 3]).all() and ha == bool((rgba[..., 3] != 255).any())


def test_tables_match_the_oracle():
    g2l, l2g = oracle.gamma_tables()
    assert (np.asarray(g2l, np.int64) == G2L).all() and (np.asarray(l2g, np.int64) == L2G).all()


def test_a_million_random_translucent_blocks():
    rng = np.random.default_rng(2026)
    rgba = rng.integers(0, 256, (1024, 2048, 4), dtype=np.uint8)       # 2^20 blocks, alpha uniform: all translucent
    _check(rgba)
    sums = rgba[..., 3].astype(np.int64).reshape(512, 2, 1024, 2).sum(axis=(1, 3))
    assert ((sums > 0) & (sums < 1020)).mean() > 0.999


def test_every_divisor():
    """Total alpha 1 .. 1019: every divisor the reference can meet, 64 random colour blocks each, with the
    alpha spread over the four pixels in several ways (one pixel carries it all, even split, random)."""
    rng = np.random.default_rng(7)
    rows = []
    for total in range(1, 1020):
        for k in range(64):
            if k % 3 == 0:
                parts = rng.multinomial(total, [0.25] * 4)
            elif k % 3 == 1:
                parts = np.array([min(total, 255), 0, 0, 0])
                parts[1] = min(total - parts[0], 255)
                parts[2] = min(total - parts[0] - parts[1], 255)
                parts[3] = total - parts[:3].sum()
            else:
                parts = np.array([total // 4 + (i < total % 4) for i in range(4)])
            while parts.max() > 255:                                    # push the excess on
                i = int(parts.argmax())
                j = int(parts.argmin())
                d = parts[i] - 255
                parts[i] -= d
                parts[j] += d
            assert parts.sum() == total and parts.max() <= 255
            rows.append(rng.permutation(parts))
    alpha = np.array(rows, np.uint8)                                    # [n, 4]
    n = alpha.shape[0]
    px = rng.integers(0, 256, (n, 4, 4), dtype=np.uint8)
    px[:, :, 3] = alpha
    img = px.reshape(n, 2, 2, 4).transpose(1, 0, 2, 3).reshape(2, 2 * n, 4)   # one row pair of n blocks
    _check(np.ascontiguousarray(img))


def test_odd_widths_and_heights():
    rng = np.random.default_rng(99)
    for (w, h) in [(1, 1), (1, 2), (2, 1), (3, 3), (5, 7), (33, 17), (255, 129), (1001, 333)]:
        rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        _check(rgba)
        sparse = rgba.copy()
        sparse[..., 3][rng.random((h, w)) < 0.6] = 255
        sparse[..., 3][rng.random((h, w)) < 0.2] = 0
        _check(sparse)
