"""Host-side mirror of go-webp's picture API for the pixel hot path.

Names, argument meaning and error behaviour follow the reference:
  Picture, WebPPictureAllocYUVA/AllocARGB      pkg/picture/picture.go:33-105,
                                               pkg/picture/allocations.go:32-122
  WebPPictureImportRGBA/RGBX/BGRA/BGRX/RGB/BGR pkg/libwebp/enc/picture_csp_enc.c.go:123-225
  PictureARGBToYUVA / WebPPictureARGBToYUVA    pkg/picture/conversions.go:12-48
  WebPPictureYUVAToARGB                        pkg/libwebp/enc/picture_csp_enc.c.go:62-118
plus the added batched entry points ImportRGBABatch and YUVAToRGBABatch
(SURVEY.md 8b).  Every conversion dispatches to the CUDA library through the C
ABI; there is no CPU path here (the Go version of this file is go/pkg/picture,
see INTEGRATION.md).
"""
from __future__ import annotations

import numpy as np

from . import _abi
from ._abi import B2Error
from .device import Context

# pkg/color/colorspace/colorspace.go:4-16
WEBP_YUV420, WEBP_YUV420A, WEBP_CSP_UV_MASK, WEBP_CSP_ALPHA_BIT = 0, 4, 3, 4
WEBP_MAX_DIMENSION = 16383  # pkg/libwebp/webp/encode.h.go:95

# pkg/picture/errors.go:8-21 (sentinels; ENC_OK is None like Go's nil)
ENC_OK = None
ENC_ERROR_OUT_OF_MEMORY = "memory error allocating objects"
ENC_ERROR_nil_PARAMETER = "a pointer parameter is nil"
ENC_ERROR_INVALID_CONFIGURATION = "configuration is invalid"
ENC_ERROR_BAD_DIMENSION = "picture has invalid width/height"

_STATUS_TO_ENC = {
    _abi.B2_ERR_OOM: ENC_ERROR_OUT_OF_MEMORY,
    _abi.B2_ERR_BAD_DIM: ENC_ERROR_BAD_DIMENSION,
    _abi.B2_ERR_NULL: ENC_ERROR_nil_PARAMETER,
    _abi.B2_ERR_INVALID_CONFIG: ENC_ERROR_INVALID_CONFIGURATION,
    _abi.B2_ERR_CUDA: ENC_ERROR_OUT_OF_MEMORY,
}

_CTX: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _CTX:
        _CTX[device] = Context(device, n_streams=4)
    return _CTX[device]


class PlaneBlock:
    """One pinned-host block and one device block with the same layout, shared by the
    pictures of a batch: the role of Picture.memory_ / memory_argb_ (picture.go:100-103)
    in the device-resident form north_star gives pkg/picture.  Freed when the last
    picture (and the last numpy view handed out) lets go of it."""

    def __init__(self, ctx: Context, nbytes: int):
        self.ctx, self.nbytes = ctx, int(nbytes)
        self.dev = ctx.alloc(self.nbytes)
        self.pin = ctx.pinned(self.nbytes)

    def host(self, off: int, shape, strides=None, dtype=np.uint8) -> np.ndarray:
        """numpy view of the pinned block; the view keeps the block alive."""
        holder = _HostView(self, off, shape, strides, dtype)
        return np.asarray(holder)


class _HostView:
    def __init__(self, block: PlaneBlock, off: int, shape, strides, dtype):
        self.block = block
        self.__array_interface__ = {"shape": tuple(shape), "typestr": np.dtype(dtype).str,
                                    "data": (block.pin.ptr + off, False), "strides": strides, "version": 3}


class PictureMemory:
    """Where a picture's planes live: offsets into a PlaneBlock (same on the host and on
    the device) + which side holds the current data.  kind 'yuva': Y | U | V | A with
    16-byte aligned strides (the TODO of allocations.go:81 'we could align the y/u/v planes
    and adjust stride'); kind 'argb': uint32 rows, stride a multiple of 4 pixels."""

    def __init__(self, block: PlaneBlock, off: int, kind: str, w: int, h: int):
        self.block, self.off, self.kind, self.w, self.h = block, off, kind, w, h
        self.host_valid = self.dev_valid = False
        if kind == "yuva":
            self.uvw, self.uvh = (w + 1) >> 1, (h + 1) >> 1
            self.ys, self.uvs = _al(w, 16), _al(self.uvw, 16)
            self.o_y = off
            self.o_u = self.o_y + self.ys * h
            self.o_v = self.o_u + self.uvs * self.uvh
            self.o_a = self.o_v + self.uvs * self.uvh
            self.nbytes = self.o_a + self.ys * h - off
            d, p = block.dev.ptr, block.pin.ptr
            # one row of the batch tables the staged entry points take (device planes, pinned planes, geometry)
            self.row = (d + self.o_y, d + self.o_u, d + self.o_v, d + self.o_a, p + self.o_y, p + self.o_u,
                        p + self.o_v, p + self.o_a, self.ys, self.uvs, w, h)
            self.bound_alpha = None
        else:
            self.stride = _al(w, 4)
            self.nbytes = 4 * self.stride * h

    @staticmethod
    def size(kind: str, w: int, h: int) -> int:
        if kind == "yuva":
            return 2 * _al(w, 16) * h + 2 * _al((w + 1) >> 1, 16) * ((h + 1) >> 1)
        return 4 * _al(w, 4) * h

    def dev(self, key: str) -> int:
        return self.block.dev.ptr + (self.off if key == "argb" else getattr(self, "o_" + key))

    def pin(self, key: str) -> int:
        return self.block.pin.ptr + (self.off if key == "argb" else getattr(self, "o_" + key))

    def views(self):
        b = self.block
        if self.kind == "argb":
            return b.host(self.off, (self.h, self.w), (4 * self.stride, 4), np.uint32)
        return (b.host(self.o_y, (self.h, self.w), (self.ys, 1)), b.host(self.o_u, (self.uvh, self.uvw), (self.uvs, 1)),
                b.host(self.o_v, (self.uvh, self.uvw), (self.uvs, 1)), b.host(self.o_a, (self.h, self.w), (self.ys, 1)))

    def fits(self, kind: str, w: int, h: int, ctx: Context) -> bool:
        return self.kind == kind and self.w == w and self.h == h and self.block.ctx is ctx

    def download(self, with_alpha: bool = True):
        """device -> pinned host for this picture alone (the batched calls do it pipelined)."""
        c = self.block.ctx
        n = self.nbytes if (with_alpha or self.kind == "argb") else self.o_a - self.off
        c._chk(c.lib.b2_d2h_async(c.h, 0, self.block.pin.ptr + self.off, self.block.dev.ptr + self.off, n))
        c.sync(0)
        self.host_valid = True

    def upload(self):
        c = self.block.ctx
        c._chk(c.lib.b2_h2d_async(c.h, 0, self.block.dev.ptr + self.off, self.block.pin.ptr + self.off, self.nbytes))
        c.sync(0)
        self.dev_valid = True


class Picture:
    """pkg/picture/picture.go:33-105 with the fields the code actually uses (SURVEY
    Appendix B-10): planes as numpy arrays, strides in bytes (ARGBStride in pixels).

    Device residency (north_star: "pkg/picture gains device-resident Y/U/V/A and ARGB
    planes with pinned-host staging"): the batched entry points give a picture a
    PictureMemory (memory_ for YUVA, memory_argb_ for ARGB, picture.go:100-103) -- planes in
    a device block with a pinned-host mirror.  Y/U/V/A/ARGB then are numpy views of the
    pinned mirror; reading them after a call that left the result on the device
    (download=False) brings the picture back first.  Assigning a plane makes the picture
    user-owned again (picture.go:28-32), so the next call uploads what was assigned.  The
    views of a resident picture are read-only: write through an assignment, or call
    MarkHostModified() after writing into a writable copy you assigned."""

    _YUVA = ("Y", "U", "V", "A")

    def __init__(self, Width: int = 0, Height: int = 0, UseARGB: bool = False, ColorSpace: int = WEBP_YUV420,
                 Y=None, U=None, V=None, A=None, YStride: int = 0, UVStride: int = 0, AStride: int = 0,
                 ARGB=None, ARGBStride: int = 0, ErrorCode=None, UserData=None):
        self.Width, self.Height, self.UseARGB, self.ColorSpace = Width, Height, UseARGB, ColorSpace
        self._Y, self._U, self._V, self._A, self._ARGB = Y, U, V, A, ARGB
        self.YStride, self.UVStride, self.AStride, self.ARGBStride = YStride, UVStride, AStride, ARGBStride
        self.ErrorCode, self.UserData = ErrorCode, UserData
        self.memory_: PictureMemory | None = None
        self.memory_argb_: PictureMemory | None = None

    def __repr__(self):
        return (f"Picture({self.Width}x{self.Height}, UseARGB={self.UseARGB}, ColorSpace={self.ColorSpace}, "
                f"resident={'yuva' if self.memory_ else ''}{'+argb' if self.memory_argb_ else ''})")

    def _get(self, name):
        m = self.memory_argb_ if name == "ARGB" else self.memory_
        if m is not None and not m.host_valid and m.dev_valid:
            m.download()
        return getattr(self, "_" + name)

    def _set(self, name, value):
        if name == "ARGB":
            if self.memory_argb_ is not None:
                self.memory_argb_ = None
        elif self.memory_ is not None:
            m = self.memory_
            if not m.host_valid and m.dev_valid:
                m.download()          # the planes that are NOT being replaced stay valid host arrays
            self.memory_ = None
        setattr(self, "_" + name, value)

    Y = property(lambda self: self._get("Y"), lambda self, v: self._set("Y", v))
    U = property(lambda self: self._get("U"), lambda self, v: self._set("U", v))
    V = property(lambda self: self._get("V"), lambda self, v: self._set("V", v))
    A = property(lambda self: self._get("A"), lambda self, v: self._set("A", v))
    ARGB = property(lambda self: self._get("ARGB"), lambda self, v: self._set("ARGB", v))

    def SetEncodingError(self, err):  # picture.go:220-229: the oldest error wins
        if self.ErrorCode is None:
            self.ErrorCode = err
        return 0

    # -- residency helpers ------------------------------------------------------------
    def resident(self, kind: str = "yuva") -> bool:
        m = self.memory_argb_ if kind == "argb" else self.memory_
        return m is not None and m.dev_valid

    def _bind(self, mem: PictureMemory, has_alpha: bool = True):
        """Points the plane fields at a PictureMemory (read-only views of its pinned mirror)."""
        if mem.kind == "argb":
            v = mem.views()
            v.flags.writeable = False
            self.memory_argb_, self._ARGB, self.ARGBStride = mem, v, mem.stride
            return
        if self.memory_ is mem and mem.bound_alpha is not None:   # same planes: only the alpha plane comes and goes
            if mem.bound_alpha != has_alpha:
                self._A = mem.a_view if has_alpha else None
                self.AStride = mem.ys if has_alpha else 0
                mem.bound_alpha = has_alpha
            return
        y, u, v, a = mem.views()
        for x in (y, u, v, a):
            x.flags.writeable = False
        self.memory_ = mem
        mem.a_view, mem.bound_alpha = a, has_alpha
        self._Y, self._U, self._V = y, u, v
        self._A = a if has_alpha else None
        self.YStride, self.UVStride, self.AStride = mem.ys, mem.uvs, (mem.ys if has_alpha else 0)


def WebPPictureFree(pic: Picture) -> None:
    """picture.go:135-139 -> WebPPictureResetBuffers: drops the planes; the device and pinned
    blocks go when the last picture of the batch that shares them is freed."""
    if pic is None:
        return
    pic.memory_ = pic.memory_argb_ = None
    pic._Y = pic._U = pic._V = pic._A = pic._ARGB = None
    pic.YStride = pic.UVStride = pic.AStride = pic.ARGBStride = 0


def WebPValidatePicture(pic: Picture):
    if pic is None:
        return ENC_ERROR_nil_PARAMETER
    if pic.Width <= 0 or pic.Height <= 0 or pic.Width > WEBP_MAX_DIMENSION or pic.Height > WEBP_MAX_DIMENSION:
        return ENC_ERROR_BAD_DIMENSION
    return None


def WebPPictureAllocYUVA(pic: Picture) -> int:
    """allocations.go:32-93: y_stride = width, uv_stride = ceil(w/2), A iff the
    colourspace has the alpha bit.  Plain host planes (user-owned picture); the batched
    entry points allocate resident pictures with AllocResidentBatch."""
    err = WebPValidatePicture(pic)
    if err:
        return pic.SetEncodingError(err)
    w, h = pic.Width, pic.Height
    uvw, uvh = (w + 1) >> 1, (h + 1) >> 1
    has_alpha = bool(pic.ColorSpace & WEBP_CSP_ALPHA_BIT)
    pic.Y = np.empty((h, w), np.uint8)
    pic.U = np.empty((uvh, uvw), np.uint8)
    pic.V = np.empty((uvh, uvw), np.uint8)
    pic.A = np.empty((h, w), np.uint8) if has_alpha else None
    pic.YStride, pic.UVStride, pic.AStride = w, uvw, (w if has_alpha else 0)
    return 1


def WebPPictureAllocARGB(pic: Picture) -> int:
    err = WebPValidatePicture(pic)
    if err:
        return pic.SetEncodingError(err)
    pic.ARGB = np.empty((pic.Height, pic.Width), np.uint32)
    pic.ARGBStride = pic.Width
    return 1


def AllocResidentBatch(pics, kind: str = "yuva", ctx: Context | None = None) -> int:
    """WebPPictureAllocYUVA / AllocARGB for a batch of device-resident pictures: ONE device
    block and ONE pinned block for all of them (a pinned allocation per picture would cost
    more than the conversion), pictures that already own memory of the right geometry keep
    it.  Returns 1, or 0 after recording the error."""
    ctx = ctx or default_context()
    todo, need = [], 0
    for pic in pics:
        err = WebPValidatePicture(pic)
        if err:
            return pic.SetEncodingError(err)
        m = pic.memory_argb_ if kind == "argb" else pic.memory_
        if m is not None and m.fits(kind, pic.Width, pic.Height, ctx):
            continue
        todo.append((pic, need))
        need += _al(PictureMemory.size(kind, pic.Width, pic.Height), 16)
    if not todo:
        return 1
    try:
        block = PlaneBlock(ctx, need)
    except B2Error as e:
        for pic, _ in todo:
            pic.SetEncodingError(_STATUS_TO_ENC.get(e.status, ENC_ERROR_OUT_OF_MEMORY))
        return 0
    for pic, off in todo:
        pic._bind(PictureMemory(block, off, kind, pic.Width, pic.Height))
    return 1


class PinnedBytes:
    """Pinned host memory for the caller's source / destination buffers (what the cgo shim
    hands to Go as a []byte over b2_pinned_alloc): .array is a uint8 numpy view."""

    def __init__(self, nbytes: int, ctx: Context | None = None):
        self.buf = (ctx or default_context()).pinned(nbytes)
        self.array = self.buf.array


class _Arena:
    """Grow-only device arena reused across batched calls of one context."""

    def __init__(self, ctx: Context):
        self.ctx, self.buf, self.cap, self.off = ctx, None, 0, 0

    def reset(self, need: int):
        if need > self.cap:
            if self.buf is not None:
                self.ctx.sync()
                self.buf.free()
            self.cap = int(need * 1.25) + (1 << 20)
            self.buf = self.ctx.alloc(self.cap)
        self.off = 0

    def take(self, nbytes: int) -> int:
        p = self.buf.ptr + self.off
        self.off += (nbytes + 255) & ~255
        assert self.off <= self.cap
        return p


_ARENAS: dict[int, _Arena] = {}


def _arena(ctx: Context) -> _Arena:
    a = _ARENAS.get(id(ctx))
    if a is None:
        a = _ARENAS[id(ctx)] = _Arena(ctx)
    return a


def _al(n: int, a: int = 256) -> int:
    return (n + a - 1) & ~(a - 1)


def _addr(arr: np.ndarray) -> int:
    return arr.__array_interface__["data"][0]


_IMPORT_JOB_DT = np.dtype(_abi.ImportJob)
_IMPORT_HOST_DT = np.dtype(_abi.ImportHost)

_ORDERS = {  # name -> (b2 order, step, import_alpha)   picture_csp_enc.c.go:189-225
    "RGBA": (_abi.ORDER_RGBA, 4, 1), "RGBX": (_abi.ORDER_RGBA, 4, 0),
    "BGRA": (_abi.ORDER_BGRA, 4, 1), "BGRX": (_abi.ORDER_BGRA, 4, 0),
    "RGB": (_abi.ORDER_RGB, 3, 0), "BGR": (_abi.ORDER_BGR, 3, 0),
}


def _fail(pics, e: B2Error) -> int:
    for pic in pics:
        pic.SetEncodingError(_STATUS_TO_ENC.get(e.status, ENC_ERROR_OUT_OF_MEMORY))
    return 0


def _stage_sources(srcs, strides, heights, row_bytes, ctx):
    """Device slots for the callers' source buffers (grow-only arena of the context) and the
    host address of pixel (0,0) of each.  Sources stay where they are: pinned memory
    (PinnedBytes / b2_pinned_alloc) is copied asynchronously, pageable memory is staged by
    the driver."""
    ar = _arena(ctx)
    sizes = [_al(abs(st) * (h - 1) + rb, 16) for st, h, rb in zip(strides, heights, row_bytes)]
    ar.reset(sum(sizes) + 256)
    base = ar.take(0)
    dev, host, keep, off = [], [], [], 0
    for src, st, h, sz in zip(srcs, strides, heights, sizes):
        a = src if type(src) is np.ndarray and src.flags.c_contiguous else np.ascontiguousarray(src)
        keep.append(a)
        back = -st * (h - 1) if st < 0 else 0   # negative stride: pixel (0,0) is in the LAST row of memory
        dev.append(base + off + back)
        host.append(a.__array_interface__["data"][0] + back)
        off += sz
    return dev, host, keep


def _import_argb_batch(pics, srcs, strides, b2order, step, import_alpha, ctx, download=True) -> int:
    """Import() into ARGB pictures: WebPPictureAlloc + WebPPackRGB / row copy /
    VP8LConvertBGRAToRGBA (picture_csp_enc.c.go:138-181) as one b2_import_to_argb launch;
    the pictures become device-resident ARGB pictures."""
    lib, hnd = ctx.lib, ctx.h
    if not AllocResidentBatch(pics, "argb", ctx):
        return 0
    try:
        dev, host, keep = _stage_sources(srcs, strides, [p.Height for p in pics], [step * p.Width for p in pics], ctx)
        jobs = []
        for pic, stride, d, hp in zip(pics, strides, dev, host):
            w, h = pic.Width, pic.Height
            span = abs(stride) * (h - 1) + step * w
            back = abs(stride) * (h - 1) if stride < 0 else 0
            ctx._chk(lib.b2_h2d_async(hnd, 0, d - back, hp - back, span))
            m = pic.memory_argb_
            jobs.append(_abi.ConvertJob(d, None, None, m.dev("argb"), stride, 0, 4 * m.stride, w, h, 0))
        arr = (_abi.ConvertJob * len(jobs))(*jobs)
        ctx._chk(lib.b2_import_to_argb(hnd, 0, arr, len(jobs), b2order, int(import_alpha)))
        for pic in pics:
            m = pic.memory_argb_
            m.dev_valid, m.host_valid = True, False
            if download:
                ctx._chk(lib.b2_d2h_async(hnd, 0, m.pin("argb"), m.dev("argb"), m.nbytes))
                m.host_valid = True
        ctx.sync()
    except B2Error as e:
        return _fail(pics, e)
    return 1


def ImportRGBABatch(pics, srcs, strides, order: str = "RGBA", ctx: Context | None = None,
                    divide_table: bool = False, chunk: int = 16, dithering: float = 0.0,
                    download: bool = True) -> int:
    """Added batched entry point (one per part): imports srcs[i]
    (interleaved bytes, row stride strides[i]) into pics[i] as YUV420(A).

    Per image this is enc.Import(...) -> picture.ImportYUVAFromRGBA(...) with
    use_iterative_conversion=0 (picture_csp_enc.c.go:123-137, conversions.go:50-165).
    The pictures become device-resident (one device + one pinned block for the batch,
    kept across calls while the geometry stays); the whole batch runs as ONE
    b2_import_batch_staged call: chunks of 'chunk' pictures go round the context's streams
    so H2D, kernel and D2H of different chunks overlap, and the alpha plane comes back only
    for pictures that have alpha (conversions.go:56-58).  download=False leaves the planes
    on the device for the next batched call (WebPCleanupTransparentAreaBatch,
    YUVAToRGBABatch, ...); reading pic.Y etc. fetches them on demand.
    Returns 1 on success, 0 after recording the error on the offending picture.
    """
    ctx = ctx or default_context()
    b2order, step, import_alpha = _ORDERS[order]
    n = len(pics)
    if n == 0:
        return 1
    for pic, src, stride in zip(pics, srcs, strides):
        if pic is None or src is None:
            return 0
        err = WebPValidatePicture(pic)
        if err:
            return pic.SetEncodingError(err)
        if abs(stride) < (4 if import_alpha else 3) * pic.Width or abs(stride) < step * pic.Width:
            return 0  # picture_csp_enc.c.go:132
    argb_idx = [i for i, pic in enumerate(pics) if pic.UseARGB]
    if argb_idx:  # use_argb branch of Import() (picture_csp_enc.c.go:138-181)
        if not _import_argb_batch([pics[i] for i in argb_idx], [srcs[i] for i in argb_idx],
                                  [strides[i] for i in argb_idx], b2order, step, import_alpha, ctx, download):
            return 0
        rest = [i for i in range(n) if not pics[i].UseARGB]
        if not rest:
            return 1
        return ImportRGBABatch([pics[i] for i in rest], [srcs[i] for i in rest], [strides[i] for i in rest],
                               order, ctx, divide_table, chunk, dithering, download)
    if not AllocResidentBatch(pics, "yuva", ctx):
        return 0
    flags = (_abi.IMPORT_ALPHA if import_alpha else 0) | (_abi.IMPORT_DIV_TABLE if divide_table else 0)
    has_alpha = (_abi.C.c_int32 * n)()
    try:
        mems = [p.memory_ for p in pics]
        t = np.array([m.row for m in mems], np.int64)       # [n, 12]: filled column-wise, not picture by picture
        dev, host, keep = _stage_sources(srcs, strides, t[:, 11].tolist(), (step * t[:, 10]).tolist(), ctx)
        jobs = (_abi.ImportJob * n)()
        hosts = (_abi.ImportHost * n)()
        ja = np.frombuffer(jobs, _IMPORT_JOB_DT)
        ha_ = np.frombuffer(hosts, _IMPORT_HOST_DT)
        ja["src"], ja["src_stride"] = dev, strides
        ja["y"], ja["u"], ja["v"], ja["a"] = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
        ja["y_stride"], ja["uv_stride"], ja["a_stride"], ja["width"], ja["height"] = t[:, 8], t[:, 9], t[:, 8], t[:, 10], t[:, 11]
        ha_["src"] = host
        if download:
            ha_["y"], ha_["u"], ha_["v"], ha_["a"] = t[:, 4], t[:, 5], t[:, 6], t[:, 7]
            ha_["y_stride"], ha_["uv_stride"], ha_["a_stride"] = t[:, 8], t[:, 9], t[:, 8]
        ctx._chk(ctx.lib.b2_import_batch_staged(ctx.h, jobs, hosts, n, b2order, flags, chunk, float(max(dithering, 0.0)),
                                                has_alpha))
    except B2Error as e:
        return _fail(pics, e)
    download = bool(download)
    for pic, m, ha in zip(pics, mems, has_alpha):
        m.dev_valid, m.host_valid = True, download
        pic.UseARGB = False                      # conversions.go:59
        if m.bound_alpha != bool(ha):
            pic._bind(m, has_alpha=bool(ha))     # conversions.go:56-58: the alpha plane exists only with alpha
        pic.ColorSpace = WEBP_YUV420A if ha else WEBP_YUV420
    return 1


def _import_one(pic, buf, stride, order):
    if pic is None or buf is None:
        return 0
    return ImportRGBABatch([pic], [buf], [stride], order)


# picture_csp_enc.c.go:189-225
def WebPPictureImportRGB(pic, rgb, rgb_stride): return _import_one(pic, rgb, rgb_stride, "RGB")
def WebPPictureImportRGBA(pic, rgba, rgba_stride): return _import_one(pic, rgba, rgba_stride, "RGBA")
def WebPPictureImportRGBX(pic, rgbx, rgbx_stride): return _import_one(pic, rgbx, rgbx_stride, "RGBX")
def WebPPictureImportBGR(pic, bgr, bgr_stride): return _import_one(pic, bgr, bgr_stride, "BGR")
def WebPPictureImportBGRA(pic, bgra, bgra_stride): return _import_one(pic, bgra, bgra_stride, "BGRA")
def WebPPictureImportBGRX(pic, bgrx, bgrx_stride): return _import_one(pic, bgrx, bgrx_stride, "BGRX")


def PictureARGBToYUVA(pic: Picture, csp: int, dithering: float = 0.0, use_iterative_conversion: int = 0,
                      divide_table: bool = False):
    """conversions.go:12-30: ARGB (uint32 0xAARRGGBB) -> YUV420(A).  Returns an
    error sentinel or None, like the Go function."""
    if pic is None:
        return None
    if pic._ARGB is None:
        pic.SetEncodingError(ENC_ERROR_nil_PARAMETER)
        return ENC_ERROR_nil_PARAMETER
    if (csp & WEBP_CSP_UV_MASK) != WEBP_YUV420:
        pic.SetEncodingError(ENC_ERROR_INVALID_CONFIGURATION)
        return ENC_ERROR_INVALID_CONFIGURATION
    if use_iterative_conversion:
        raise NotImplementedError("sharp-YUV import is not on the hot path (SURVEY 5; the sharpyuv package is absent)")
    argb = np.ascontiguousarray(pic.ARGB, np.uint32)
    stride = 4 * argb.shape[1]
    pic.UseARGB = False  # ImportYUVAFromRGBA sets use_argb = 0 (conversions.go:59)
    ok = ImportRGBABatch([pic], [argb.view(np.uint8)], [stride], "BGRA", divide_table=divide_table,
                         dithering=max(float(dithering), 0.0))
    return None if ok else (pic.ErrorCode or ENC_ERROR_OUT_OF_MEMORY)


def WebPPictureARGBToYUVA(pic: Picture, csp: int = WEBP_YUV420):
    return PictureARGBToYUVA(pic, csp, 0.0, 0)


def WebPPictureARGBToYUVADithered(pic: Picture, csp: int, dithering: float):
    """conversions.go:36-38."""
    return PictureARGBToYUVA(pic, csp, dithering, 0)


def _mode_takes_alpha(mode: int) -> bool:
    return _abi.MODE_BPP[mode] == 4 or mode in (_abi.MODE_RGBA_4444, _abi.MODE_rgbA_4444)


def _yuva_device_planes(pics, ctx, want_alpha, extra_bytes: int = 0):
    """Device planes of a batch of YUV(A) pictures: resident pictures are used where they
    lie (no copy); user-owned planes get slots in the context's arena and host addresses for
    the staged upload.  Returns per picture (dev dict, host dict or None, strides, has_a)."""
    need, plan = 0, []
    for pic, wa in zip(pics, want_alpha):
        m = pic.memory_
        if m is not None and m.dev_valid and m.block.ctx is ctx:
            plan.append(None)
            continue
        w, h = pic.Width, pic.Height
        uvw, uvh = (w + 1) >> 1, (h + 1) >> 1
        ys, uvs = _al(w, 16), _al(uvw, 16)
        sz = (ys * h, uvs * uvh, uvs * uvh, ys * h if wa else 0)
        plan.append((ys, uvs, sz))
        need += sum(sz)
    ar = _arena(ctx)
    base = 0
    if need + extra_bytes:
        ar.reset(need + extra_bytes + 256)
        base = ar.take(0)
    out, keep, off = [], [], 0
    for pic, wa, pl in zip(pics, want_alpha, plan):
        if pl is None:
            m = pic.memory_
            out.append((dict(y=m.dev("y"), u=m.dev("u"), v=m.dev("v"), a=m.dev("a") if wa else None), None, (m.ys, m.uvs)))
            continue
        ys, uvs, sz = pl
        d, hst = {}, {}
        for key, nbytes, plane in zip("yuva", sz, (pic.Y, pic.U, pic.V, pic.A if wa else None)):
            if not nbytes:
                d[key] = None
                hst[key] = (None, 0)
                continue
            arr = np.asarray(plane, np.uint8)
            if arr.strides[-1] != 1 or arr.strides[0] < arr.shape[1]:
                arr = np.ascontiguousarray(arr)
            keep.append(arr)
            d[key] = base + off
            hst[key] = (_addr(arr), arr.strides[0])
            off += nbytes
        out.append((d, hst, (ys, uvs)))
    if extra_bytes:
        return out, keep, base + _al(off, 256)
    return out, keep


def YUVAToRGBABatch(pics, mode: int = _abi.MODE_BGRA, dsts=None, ctx: Context | None = None,
                    chunk: int = 8, download: bool = True):
    """Added batched entry point for part 2: fancy-upsamples pics[i].Y/U/V(/A)
    into dsts[i] (uint8 [h, w*bpp]; allocated when None) in WEBP_CSP_MODE 'mode'.
    Per image this is the schedule of WebPPictureYUVAToARGB
    (picture_csp_enc.c.go:62-118) with WebPUpsamplers[mode].  Device-resident pictures
    (ImportRGBABatch, the picture tools) are read where they lie: only the output crosses
    PCIe.  Returns the list of outputs, or None after recording the error."""
    ctx = ctx or default_context()
    n = len(pics)
    bpp = _abi.MODE_BPP[mode]
    if dsts is None:
        dsts = [None] * n
    want_alpha = []
    for pic in pics:
        if pic is None:
            return None
        if pic._Y is None or pic._U is None or pic._V is None:
            pic.SetEncodingError(ENC_ERROR_nil_PARAMETER)
            return None
        if (pic.ColorSpace & WEBP_CSP_ALPHA_BIT) and pic._A is None:
            pic.SetEncodingError(ENC_ERROR_nil_PARAMETER)
            return None
        if (pic.ColorSpace & WEBP_CSP_UV_MASK) != WEBP_YUV420:
            pic.SetEncodingError(ENC_ERROR_INVALID_CONFIGURATION)
            return None
        want_alpha.append(bool(pic.ColorSpace & WEBP_CSP_ALPHA_BIT) and _mode_takes_alpha(mode))
    if n == 0:
        return []
    try:
        # outputs: caller buffers are written in place (device side: the context's arena); missing ones
        # come from one pinned block (+ its device twin), so they are resident for a following step
        out, need = [], 0
        for pic, dst in zip(pics, dsts):
            need += 0 if dst is not None else _al(_al(pic.Width * bpp, 16) * pic.Height, 16)
        scratch = sum(_al(_al(p.Width * bpp, 16) * p.Height, 16) for p, d in zip(pics, dsts) if d is not None)
        planes, keep, tmp_base = _yuva_device_planes(pics, ctx, want_alpha, scratch + 256)
        block = PlaneBlock(ctx, need) if need else None
        jobs = (_abi.UpsampleJob * n)()
        hosts = (_abi.UpsampleHost * n)()
        boff = toff = 0
        for i, (pic, dst) in enumerate(zip(pics, dsts)):
            w, h = pic.Width, pic.Height
            ds = _al(w * bpp, 16)
            d, hst, (ys, uvs) = planes[i]
            j, hh = jobs[i], hosts[i]
            j.y, j.u, j.v, j.a = d["y"], d["u"], d["v"], d["a"]
            j.y_stride, j.uv_stride, j.a_stride, j.dst_stride, j.width, j.height = ys, uvs, ys, ds, w, h
            if hst is not None:
                (hh.y, hh.y_stride), (hh.u, hh.uv_stride), (hh.v, _), (hh.a, hh.a_stride) = hst["y"], hst["u"], hst["v"], hst["a"]
            if dst is None:
                j.dst = block.dev.ptr + boff
                o = block.host(boff, (h, w * bpp), (ds, 1))
                if download:
                    hh.dst, hh.dst_stride = block.pin.ptr + boff, ds
                boff += _al(ds * h, 16)
            else:
                o = dst
                j.dst = tmp_base + toff
                toff += _al(ds * h, 16)
                hh.dst, hh.dst_stride = _addr(o), o.strides[0]
            out.append(o)
        ctx._chk(ctx.lib.b2_upsample_batch_staged(ctx.h, jobs, hosts, n, mode, chunk))
    except B2Error as e:
        _fail(pics, e)
        return None
    return out


def EmitFancyRGB(pic: Picture, mode: int, mb_y: int, mb_h: int, dst: np.ndarray, ctx: Context | None = None) -> int:
    """EmitFancyRGB (pkg/libwebp/decoder/io_dec.c.go:65-115) for the macroblock-row batch
    [mb_y, mb_y + mb_h) of a decoded picture: writes the output rows that batch completes
    into dst (uint8 [h, w*bpp]) and returns their number.  First batch: row 0 is
    mirrored; every batch leaves its last row for the next call (it needs the chroma
    row below), the last batch also emits the final row.  The reference carries the
    unfinished row in tmp_y/u/v; here the planes are whole, so the carried row is simply
    read again, and EmitAlphaRGB (io_dec.c.go:175-195) is fused when pic has alpha."""
    ctx = ctx or default_context()
    w, h = pic.Width, pic.Height
    bpp = _abi.MODE_BPP[mode]
    y_end = min(mb_y + mb_h, h)
    row_begin = max(mb_y - 1, 0)
    row_end = h if y_end >= h else y_end - 1
    if row_end <= row_begin:
        return 0
    has_a = bool(pic.ColorSpace & WEBP_CSP_ALPHA_BIT) and _mode_takes_alpha(mode) and pic._A is not None
    try:
        ds = _al(w * bpp, 16)
        planes, keep, tmp_ptr = _yuva_device_planes([pic], ctx, [has_a], ds * h + 256)
        d, hst, (ys, uvs) = planes[0]
        job = _abi.UpsampleJob(d["y"], d["u"], d["v"], d["a"], tmp_ptr, ys, uvs, ys, ds, w, h, row_begin, row_end)
        hh = _abi.UpsampleHost()
        if hst is not None:
            (hh.y, hh.y_stride), (hh.u, hh.uv_stride), (hh.v, _), (hh.a, hh.a_stride) = hst["y"], hst["u"], hst["v"], hst["a"]
        hh.dst, hh.dst_stride = _addr(dst), dst.strides[0]
        ctx._chk(ctx.lib.b2_upsample_batch_staged(ctx.h, (_abi.UpsampleJob * 1)(job), (_abi.UpsampleHost * 1)(hh), 1, mode, 1))
    except B2Error as e:
        _fail([pic], e)
        return 0
    return row_end - row_begin


def WebPPictureYUVAToARGB(pic: Picture, download: bool = True) -> int:
    """picture_csp_enc.c.go:62-118: fills pic.ARGB (uint32 0xAARRGGBB, i.e. the
    MODE_BGRA byte order on little-endian, :85) and sets UseARGB."""
    return YUVAToARGBBatch([pic], download=download)


def YUVAToARGBBatch(pics, ctx: Context | None = None, chunk: int = 8, download: bool = True) -> int:
    """WebPPictureYUVAToARGB for a batch: the ARGB planes become device-resident
    (memory_argb_), the YUVA planes are read where they lie."""
    ctx = ctx or default_context()
    for pic in pics:
        if pic is None:
            return 0
        if pic._Y is None or pic._U is None or pic._V is None:
            return pic.SetEncodingError(ENC_ERROR_nil_PARAMETER)
        if (pic.ColorSpace & WEBP_CSP_ALPHA_BIT) and pic._A is None:
            return pic.SetEncodingError(ENC_ERROR_nil_PARAMETER)
        if (pic.ColorSpace & WEBP_CSP_UV_MASK) != WEBP_YUV420:
            return pic.SetEncodingError(ENC_ERROR_INVALID_CONFIGURATION)
    if not AllocResidentBatch(pics, "argb", ctx):
        return 0
    n = len(pics)
    try:
        want_alpha = [bool(p.ColorSpace & WEBP_CSP_ALPHA_BIT) for p in pics]
        planes, keep = _yuva_device_planes(pics, ctx, want_alpha)
        jobs = (_abi.UpsampleJob * n)()
        hosts = (_abi.UpsampleHost * n)()
        for i, pic in enumerate(pics):
            d, hst, (ys, uvs) = planes[i]
            m = pic.memory_argb_
            j, hh = jobs[i], hosts[i]
            j.y, j.u, j.v, j.a, j.dst = d["y"], d["u"], d["v"], d["a"], m.dev("argb")
            j.y_stride, j.uv_stride, j.a_stride, j.dst_stride, j.width, j.height = ys, uvs, ys, 4 * m.stride, m.w, m.h
            if hst is not None:
                (hh.y, hh.y_stride), (hh.u, hh.uv_stride), (hh.v, _), (hh.a, hh.a_stride) = hst["y"], hst["u"], hst["v"], hst["a"]
            if download:
                hh.dst, hh.dst_stride = m.pin("argb"), 4 * m.stride
        ctx._chk(ctx.lib.b2_upsample_batch_staged(ctx.h, jobs, hosts, n, _abi.MODE_BGRA, chunk))
    except B2Error as e:
        return _fail(pics, e)
    for pic in pics:
        m = pic.memory_argb_
        m.dev_valid, m.host_valid = True, bool(download)
        pic.UseARGB = True
    return 1
