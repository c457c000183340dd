"""Minimal VP8L bitstream reader -- TEST INFRASTRUCTURE, used only by
make_golden_vp8l.py to cut a libwebp-encoded lossless stream at the point where the
hot path starts: it entropy-decodes the stream into the transform list and the
residual ARGB image that `ApplyInverseTransforms` receives, WITHOUT applying any
transform.  (The transforms themselves are what the oracle and the CUDA path are
checked on; see tests/test_oracle_golden.py.)

Follows the WebP lossless format as the reference restates it (go-webp tree):
  header                      pkg/vp8/vp8l_dec.c.go:112-122
  ReadTransform               pkg/vp8/vp8l_dec.c.go:1221-1270 (palette delta coding :1196-1218)
  DecodeImageStream           pkg/vp8/vp8l_dec.c.go:1330-1413
  ReadHuffmanCodes(+Helper)   pkg/vp8/vp8l_dec.c.go:344-528, ReadHuffmanCode :301-342,
                              ReadHuffmanCodeLengths :235-299
  DecodeImageData             pkg/vp8/vp8l_dec.c.go:1038-1194
  GetCopyDistance / PlaneCodeToDistance  pkg/vp8/vp8l_dec.c.go:143-172, table :76-77
  colour cache                pkg/libwebp/utils/color_cache_utils.go (hash 0x1e35a7bd)
Pure Python: meant for images of a few ten thousand pixels.
"""
from __future__ import annotations

import numpy as np

PREDICTOR, CROSS_COLOR, SUBTRACT_GREEN, COLOR_INDEXING = 0, 1, 2, 3
NUM_LITERAL, NUM_LENGTH, NUM_DISTANCE = 256, 24, 40
CODE_LENGTH_ORDER = [17, 18, 0, 1, 2, 3, 4, 5, 16, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]
CODE_TO_PLANE = [
    0x18, 0x07, 0x17, 0x19, 0x28, 0x06, 0x27, 0x29, 0x16, 0x1a, 0x26, 0x2a, 0x38, 0x05, 0x37, 0x39, 0x15, 0x1b, 0x36, 0x3a,
    0x25, 0x2b, 0x48, 0x04, 0x47, 0x49, 0x14, 0x1c, 0x35, 0x3b, 0x46, 0x4a, 0x24, 0x2c, 0x58, 0x45, 0x4b, 0x34, 0x3c, 0x03,
    0x57, 0x59, 0x13, 0x1d, 0x56, 0x5a, 0x23, 0x2d, 0x44, 0x4c, 0x55, 0x5b, 0x33, 0x3d, 0x68, 0x02, 0x67, 0x69, 0x12, 0x1e,
    0x66, 0x6a, 0x22, 0x2e, 0x54, 0x5c, 0x43, 0x4d, 0x65, 0x6b, 0x32, 0x3e, 0x78, 0x01, 0x77, 0x79, 0x53, 0x5d, 0x11, 0x1f,
    0x64, 0x6c, 0x42, 0x4e, 0x76, 0x7a, 0x21, 0x2f, 0x75, 0x7b, 0x31, 0x3f, 0x63, 0x6d, 0x52, 0x5e, 0x00, 0x74, 0x7c, 0x41,
    0x4f, 0x10, 0x20, 0x62, 0x6e, 0x30, 0x73, 0x7d, 0x51, 0x5f, 0x40, 0x72, 0x7e, 0x61, 0x6f, 0x50, 0x71, 0x7f, 0x60, 0x70]


def sub_sample_size(size: int, bits: int) -> int:
    return (size + (1 << bits) - 1) >> bits


class BitReader:
    """LSB-first bit reader (VP8LReadBits)."""

    def __init__(self, data: bytes):
        self.v = int.from_bytes(data, "little")
        self.pos = 0
        self.nbits = 8 * len(data)

    def read(self, n: int) -> int:
        if n == 0:
            return 0
        r = (self.v >> self.pos) & ((1 << n) - 1)
        self.pos += n
        if self.pos > self.nbits:
            raise ValueError("VP8L stream truncated")
        return r


class Huffman:
    """Canonical prefix code, codes read LSB first (bit-reversed canonical codes)."""

    def __init__(self, lengths):
        self.single = None
        syms = [(l, s) for s, l in enumerate(lengths) if l > 0]
        if len(syms) == 0:
            self.single = 0
            return
        if len(syms) == 1:
            self.single = syms[0][1]   # zero-length code: no bits are read
            return
        syms.sort()
        self.table = {}
        code, prev_len = 0, syms[0][0]
        for l, s in syms:
            code <<= (l - prev_len)
            prev_len = l
            rev = int(format(code, "0%db" % l)[::-1], 2)
            self.table[(l, rev)] = s
            code += 1
        self.max_len = syms[-1][0]

    def read(self, br: BitReader) -> int:
        if self.single is not None:
            return self.single
        acc = 0
        for l in range(1, self.max_len + 1):
            acc |= br.read(1) << (l - 1)
            s = self.table.get((l, acc))
            if s is not None:
                return s
        raise ValueError("bad prefix code")


def read_code_lengths(br: BitReader, cl_lengths, num_symbols: int):
    """ReadHuffmanCodeLengths (vp8l_dec.c.go:235-299)."""
    h = Huffman(cl_lengths)
    if br.read(1):
        length_nbits = 2 + 2 * br.read(3)
        max_symbol = 2 + br.read(length_nbits)
        if max_symbol > num_symbols:
            raise ValueError("max_symbol")
    else:
        max_symbol = num_symbols
    lengths = [0] * num_symbols
    prev, sym = 8, 0
    while sym < num_symbols:
        if max_symbol == 0:
            break
        max_symbol -= 1
        c = h.read(br)
        if c < 16:
            lengths[sym] = c
            sym += 1
            if c != 0:
                prev = c
        else:
            slot = c - 16
            extra = (2, 3, 7)[slot]
            offset = (3, 3, 11)[slot]
            rep = br.read(extra) + offset
            if sym + rep > num_symbols:
                raise ValueError("repeat overflow")
            val = prev if c == 16 else 0
            for _ in range(rep):
                lengths[sym] = val
                sym += 1
    return lengths


def read_huffman_code(br: BitReader, alphabet_size: int) -> Huffman:
    """ReadHuffmanCode (vp8l_dec.c.go:301-342)."""
    if br.read(1):  # simple code
        lengths = [0] * alphabet_size
        num = br.read(1) + 1
        first_8 = br.read(1)
        lengths[br.read(8 if first_8 else 1)] = 1
        if num == 2:
            lengths[br.read(8)] = 1
        return Huffman(lengths)
    n = br.read(4) + 4
    cl = [0] * 19
    for i in range(n):
        cl[CODE_LENGTH_ORDER[i]] = br.read(3)
    return Huffman(read_code_lengths(br, cl, alphabet_size))


def prefix_value(sym: int, br: BitReader) -> int:
    """GetCopyDistance / GetCopyLength (vp8l_dec.c.go:143-156)."""
    if sym < 4:
        return sym + 1
    extra = (sym - 2) >> 1
    offset = (2 + (sym & 1)) << extra
    return offset + br.read(extra) + 1


def plane_code_to_distance(xsize: int, code: int) -> int:
    if code > 120:
        return code - 120
    d = CODE_TO_PLANE[code - 1]
    dist = (d >> 4) * xsize + (8 - (d & 0xf))
    return dist if dist >= 1 else 1


def decode_image_stream(br: BitReader, xsize: int, ysize: int, level0: bool):
    """DecodeImageStream (vp8l_dec.c.go:1330-1413) + DecodeImageData (:1038-1194).
    Returns (transforms, pixels, final_xsize): transforms in stream order."""
    transforms = []
    if level0:
        seen = 0
        while br.read(1):
            t = br.read(2)
            if seen & (1 << t):
                raise ValueError("transform repeated")
            seen |= 1 << t
            tr = {"type": t, "xsize": xsize, "ysize": ysize}
            if t in (PREDICTOR, CROSS_COLOR):
                tr["bits"] = 2 + br.read(3)
                _, data, _ = decode_image_stream(br, sub_sample_size(xsize, tr["bits"]),
                                                 sub_sample_size(ysize, tr["bits"]), False)
                tr["data"] = data
            elif t == COLOR_INDEXING:
                n = br.read(8) + 1
                bits = 0 if n > 16 else 1 if n > 4 else 2 if n > 2 else 3
                tr["bits"] = bits
                tr["num_colors"] = n
                _, data, _ = decode_image_stream(br, n, 1, False)
                tr["data"] = data      # delta coded, as read (ExpandColorMap is part of the hot path)
                xsize = sub_sample_size(xsize, bits)
            transforms.append(tr)
    cache_bits = 0
    if br.read(1):
        cache_bits = br.read(4)
        if not 1 <= cache_bits <= 11:
            raise ValueError("cache bits")
    # meta prefix codes
    meta_bits, meta = 0, None
    num_groups = 1
    if level0 and br.read(1):
        meta_bits = br.read(3) + 2
        mw, mh = sub_sample_size(xsize, meta_bits), sub_sample_size(ysize, meta_bits)
        _, mdata, _ = decode_image_stream(br, mw, mh, False)
        meta = [(int(p) >> 8) & 0xffff for p in mdata]
        num_groups = max(meta) + 1
        meta_w = mw
    groups = []
    for _ in range(num_groups):
        sizes = [NUM_LITERAL + NUM_LENGTH + ((1 << cache_bits) if cache_bits else 0), 256, 256, 256, NUM_DISTANCE]
        groups.append([read_huffman_code(br, s) for s in sizes])
    n = xsize * ysize
    out = [0] * n
    cache = [0] * (1 << cache_bits) if cache_bits else None
    shift = 32 - cache_bits
    pos = 0
    last_cached = 0
    g = groups[0]

    def cache_insert_upto(p):
        nonlocal last_cached
        if cache is not None:
            while last_cached < p:
                v = out[last_cached]
                cache[((v * 0x1e35a7bd) & 0xffffffff) >> shift] = v
                last_cached += 1

    while pos < n:
        if meta is not None:
            x, y = pos % xsize, pos // xsize
            g = groups[meta[(y >> meta_bits) * meta_w + (x >> meta_bits)]]
        s = g[0].read(br)
        if s < NUM_LITERAL:
            r = g[1].read(br)
            b = g[2].read(br)
            a = g[3].read(br)
            out[pos] = (a << 24) | (r << 16) | (s << 8) | b
            pos += 1
        elif s < NUM_LITERAL + NUM_LENGTH:
            length = prefix_value(s - NUM_LITERAL, br)
            dist = plane_code_to_distance(xsize, prefix_value(g[4].read(br), br))
            if pos < dist or pos + length > n:
                raise ValueError("bad backward reference")
            for k in range(length):
                out[pos + k] = out[pos + k - dist]
            pos += length
        else:
            cache_insert_upto(pos)
            out[pos] = cache[s - NUM_LITERAL - NUM_LENGTH]
            pos += 1
    return transforms, np.array(out, dtype=np.uint32), xsize


def parse_vp8l(payload: bytes):
    """payload = the VP8L chunk data.  Returns dict(width, height, transforms, residual)
    where residual is the (height, final_xsize) uint32 image the inverse transforms start from."""
    if payload[0] != 0x2f:
        raise ValueError("not a VP8L stream")
    br = BitReader(payload[1:])
    w = br.read(14) + 1
    h = br.read(14) + 1
    br.read(1)          # alpha_is_used hint
    if br.read(3) != 0:
        raise ValueError("VP8L version")
    transforms, pix, fx = decode_image_stream(br, w, h, True)
    return {"width": w, "height": h, "transforms": transforms, "residual": pix.reshape(h, fx)}


def vp8l_chunk(webp: bytes) -> bytes:
    """Extracts the VP8L chunk payload from a RIFF/WEBP file."""
    assert webp[:4] == b"RIFF" and webp[8:12] == b"WEBP"
    p = 12
    while p + 8 <= len(webp):
        tag = webp[p:p + 4]
        size = int.from_bytes(webp[p + 4:p + 8], "little")
        if tag == b"VP8L":
            return webp[p + 8:p + 8 + size]
        p += 8 + size + (size & 1)
    raise ValueError("no VP8L chunk")
