"""Palettes that make ApplyPalette (pkg/vp8/vp8l_encoder.go:1149-1212) take each of its strategies."""
import numpy as np


def _h1(c):
    return (((c & 0xffffff) * 4222244071) & 0xffffffff) >> 21


def _h2(c):
    return (((c & 0xffffff) * 0x7fffffff) & 0xffffffff) >> 21


def hash2_palette():
    """Four colours for which hashes 0 and 1 collide and hash 2 does not.  Hash 2 keeps only bit 0 and bits
    21..23 of the colour (c * (2^31 - 1) mod 2^32), so it is perfect only for a few colours: take 4 that share
    a green byte (hash 0 collides) and, two of them, a hash-1 bucket, but differ in those 4 bits."""
    first, pair = {}, None
    for rb in range(1 << 16):
        c = ((rb >> 8) << 16) | 0x5500 | (rb & 0xff)
        k = _h1(c)
        if k in first and _h2(first[k]) != _h2(c):
            pair = (first[k], c)
            break
        first.setdefault(k, c)
    assert pair is not None
    rest = [c for c in (0x205501, 0x405500, 0x605501, 0x805500, 0xa05501, 0xc05500)
            if _h2(c) not in (_h2(pair[0]), _h2(pair[1]))][:2]
    assert len({_h2(c) for c in rest}) == 2
    return np.array([pair[0], rest[0], pair[1], rest[1]], np.uint32) | np.uint32(0xff000000)


def palette_cases(rng):
    """(name, palette) pairs: sizes 2..256 of random colours (greedy, hash 1, sorted search), gray ramps
    (hash 0), colours that differ in alpha only (hashes 1 and 2 forget alpha: sorted search), and the
    constructed hash-2 case."""
    for P in (2, 3, 4, 5, 16, 17, 40, 200, 256):
        pal = rng.choice(1 << 32, P, replace=False).astype(np.uint32)
        yield f"random{P}", pal
    for P in (64, 256):
        pal = ((np.arange(P, dtype=np.uint64) * (256 // P)) * 0x010101 | 0xff000000).astype(np.uint32)
        rng.shuffle(pal)
        yield f"gray{P}", pal
    pal = ((np.arange(32, dtype=np.uint64) << 24) | 0x00123456).astype(np.uint32)
    rng.shuffle(pal)
    yield "alpha_only32", pal
    yield "hash2_4", hash2_palette()
