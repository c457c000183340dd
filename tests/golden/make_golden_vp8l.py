"""Generate tests/golden/vp8l_libwebp160.npz: VP8L inverse-transform vectors cut out of
streams written and decoded by the upstream libwebp 1.6.0 binary.

    python tests/golden/make_golden_vp8l.py

For every synthetic image: WebPEncodeLosslessRGBA (the binary's own encoder picks the
transforms: predictor with per-tile modes, cross-colour with per-tile multipliers,
subtract-green, colour indexing with 1/2/4/8-bit packing), then
  * vp8l_bitstream.parse_vp8l entropy-decodes the stream up to the input of
    ApplyInverseTransforms (pkg/vp8/vp8l_dec.c.go:710-729): transform list + residual;
  * WebPDecodeBGRA of the same stream gives the image those transforms must produce.
The script asserts that the oracle's inverse chain maps the first onto the second --
that is what pins oracle part 3 to the upstream binary -- and stores both so that the
oracle test and the GPU parity test run without libwebp or the bitstream reader.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import oracle  # noqa: E402
from oracle import libwebp_ref as ref  # noqa: E402
import vp8l_bitstream as vb  # noqa: E402


def encode_lossless(rgba: np.ndarray) -> bytes:
    L = ref.lib()
    h, w, _ = rgba.shape
    rgba = np.ascontiguousarray(rgba)
    out = C.c_void_p()
    L.WebPEncodeLosslessRGBA.restype = C.c_size_t
    L.WebPEncodeLosslessRGBA.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    n = L.WebPEncodeLosslessRGBA(rgba.ctypes.data, w, h, 4 * w, C.byref(out))
    assert n > 0
    data = C.string_at(out.value, n)
    L.WebPFree.argtypes = [C.c_void_p]
    L.WebPFree(out)
    return data


def decode_bgra(webp: bytes):
    L = ref.lib()
    w, h = C.c_int(), C.c_int()
    L.WebPDecodeBGRA.restype = C.c_void_p
    L.WebPDecodeBGRA.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    p = L.WebPDecodeBGRA(webp, len(webp), C.byref(w), C.byref(h))
    assert p
    a = np.frombuffer(C.string_at(p, 4 * w.value * h.value), np.uint8).copy()
    L.WebPFree.argtypes = [C.c_void_p]
    L.WebPFree(p)
    return a.view(np.uint32).reshape(h.value, w.value)   # little endian B,G,R,A bytes = 0xAARRGGBB


def apply_inverse(parsed):
    """ApplyInverseTransforms (vp8l_dec.c.go:710-729): last transform read is undone first."""
    img = parsed["residual"]
    for tr in reversed(parsed["transforms"]):
        w, h, t = tr["xsize"], tr["ysize"], tr["type"]
        if t == vb.PREDICTOR:
            b = tr["bits"]
            img = oracle.predictor_inverse(img, tr["data"].reshape(vb.sub_sample_size(h, b), vb.sub_sample_size(w, b)), b)
        elif t == vb.CROSS_COLOR:
            b = tr["bits"]
            img = oracle.color_space_inverse(img, tr["data"].reshape(vb.sub_sample_size(h, b), vb.sub_sample_size(w, b)), b)
        elif t == vb.SUBTRACT_GREEN:
            img = oracle.add_green(img)
        else:
            cmap = oracle.expand_color_map(tr["data"], tr["bits"])
            img = oracle.color_index_inverse(img, cmap, tr["bits"], w)
    return img


def images():
    rng = np.random.default_rng(0xB2003)
    out = []

    def photo(w, h, noise, alpha):
        yy, xx = np.mgrid[0:h, 0:w]
        r = (xx * 255 // max(w - 1, 1) + 40 * np.sin(yy / 7.0)).astype(np.int64)
        g = (yy * 255 // max(h - 1, 1) + 30 * np.cos(xx / 5.0)).astype(np.int64)
        b = ((xx + yy) * 255 // max(w + h - 2, 1)).astype(np.int64) + (r >> 2)
        img = np.stack([r, g, b, np.full_like(r, 255)], -1) + rng.integers(-noise, noise + 1, (h, w, 4))
        img = np.clip(img, 0, 255).astype(np.uint8)
        if alpha == "opaque":
            img[..., 3] = 255
        elif alpha == "grad":
            img[..., 3] = np.clip(255 - xx * 200 // max(w - 1, 1), 1, 255).astype(np.uint8)
        return img

    out.append(("photo_96x80", photo(96, 80, 6, "opaque")))
    out.append(("photo_131x67_odd", photo(131, 67, 3, "opaque")))
    out.append(("photo_alpha_70x90", photo(70, 90, 10, "grad")))
    out.append(("photo_noisy_64x64", photo(64, 64, 40, "opaque")))
    out.append(("photo_wide_300x33", photo(300, 33, 2, "grad")))
    out.append(("photo_flat_50x50", photo(50, 50, 0, "opaque")))
    out.append(("photo_large_400x300", photo(400, 300, 4, "opaque")))
    # grey-ish images: R and B follow G, which is what makes the encoder pick subtract-green
    for nm, (w, h), noise in (("grey_80x60", (80, 60), 2), ("grey_noisy_120x45", (120, 45), 12), ("grey_257x129", (257, 129), 5)):
        yy, xx = np.mgrid[0:h, 0:w]
        g = (128 + 100 * np.sin(xx / 9.0) * np.cos(yy / 6.0)).astype(np.int64)
        base = np.stack([g, g, g, np.full_like(g, 255)], -1) + rng.integers(-noise, noise + 1, (h, w, 4))
        img = np.clip(base, 0, 255).astype(np.uint8)
        img[..., 3] = 255
        out.append((nm, img))
    out.append(("random_40x40", rng.integers(0, 256, (40, 40, 4), dtype=np.uint8)))
    for n_col, (w, h) in ((2, (67, 41)), (3, (64, 40)), (4, (33, 70)), (9, (90, 31)), (16, (48, 48)), (40, (77, 55)), (200, (96, 64))):
        pal = rng.integers(0, 256, (n_col, 4), dtype=np.uint8)
        pal[:, 3] = 255 if n_col % 2 else rng.integers(1, 256, n_col)
        yy, xx = np.mgrid[0:h, 0:w]
        idx = ((xx // 3 + yy // 2) + rng.integers(0, 2, (h, w))) % n_col
        out.append((f"palette{n_col}_{w}x{h}", pal[idx]))
    return out


def main():
    store = {}
    names = []
    seen_types = set()
    for name, rgba in images():
        webp = encode_lossless(rgba)
        expected = decode_bgra(webp)
        parsed = vb.parse_vp8l(vb.vp8l_chunk(webp))
        got = apply_inverse(parsed)
        assert got.shape == expected.shape and (got == expected).all(), f"{name}: oracle chain != libwebp decode"
        kinds = [(t["type"], t.get("bits", 0)) for t in parsed["transforms"]]
        seen_types.update(kinds)
        print(f"{name:24s} {len(webp):6d} bytes  transforms (stream order): {kinds}")
        names.append(name)
        store[f"{name}/residual"] = parsed["residual"]
        store[f"{name}/expected"] = expected
        store[f"{name}/types"] = np.array([t["type"] for t in parsed["transforms"]], np.int32)
        store[f"{name}/bits"] = np.array([t.get("bits", 0) for t in parsed["transforms"]], np.int32)
        store[f"{name}/xsize"] = np.array([t["xsize"] for t in parsed["transforms"]], np.int32)
        store[f"{name}/ysize"] = np.array([t["ysize"] for t in parsed["transforms"]], np.int32)
        for k, t in enumerate(parsed["transforms"]):
            if "data" in t:
                store[f"{name}/data{k}"] = t["data"]
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "vp8l_libwebp160.npz"), **store)
    print("transform kinds covered (type, bits):", sorted(seen_types))
    modes = np.zeros(16, np.int64)
    for name in names:
        for k, t in enumerate(store[f"{name}/types"]):
            if t == vb.PREDICTOR:
                modes += np.bincount((store[f"{name}/data{k}"] >> 8) & 0xf, minlength=16)
    print("predictor modes used (tiles per mode 0..15):", modes.tolist())


if __name__ == "__main__":
    main()
