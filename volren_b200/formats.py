"""Host-side data formats either side of the hot path (Python mirror; the C++ host in host/ has its own):

* `.brick` / `.dense` files: cereal PortableBinary archives written by voldata
  (reference: submodules/voldata/src/serialization.cpp:16-43,66-80; layout decoded in SURVEY.md App. A)
* Radiance `.hdr` (RGBE, new-style RLE), loaded bottom-up like cppgl's image_load
  (reference: submodules/cppgl/src/image_load_store.cpp:14-41, stb_image.h hdr path)
* LUT text files + the monotone-alpha CDF rewrite (reference: src/transferfunc.cpp:33-58,79-93)
"""
from __future__ import annotations

import struct

import numpy as np


class BrickGridData:
    """Plain numpy holder of one voldata::BrickGrid (grid_brick.h:27-33); arrays are [z][y][x]."""

    def __init__(self, n_bricks, atlas_dim, brick_count, indirection, range_, atlas, mips, min_maj=(0.0, 0.0),
                 transform=None):
        self.n_bricks = tuple(int(v) for v in n_bricks)
        self.atlas_dim = tuple(int(v) for v in atlas_dim)
        self.brick_count = int(brick_count)
        self.indirection = np.ascontiguousarray(indirection, dtype=np.uint32)
        self.range = np.ascontiguousarray(range_, dtype=np.uint32)
        self.atlas = np.ascontiguousarray(atlas, dtype=np.uint8)
        self.mips = [np.ascontiguousarray(m, dtype=np.uint32) for m in mips]
        self.min_maj = (float(min_maj[0]), float(min_maj[1]))
        # 4x4, stored so that transform[c] is glm column c (i.e. the transpose of the math matrix)
        self.transform = np.eye(4, dtype=np.float32) if transform is None else np.asarray(transform, np.float32).reshape(4, 4)

    def index_extent(self):
        return tuple(8 * n for n in self.n_bricks)

    def matrix(self):
        """Index->world transform as a conventional row-major math matrix."""
        return self.transform.T.copy()

    def decode_ptr(self):
        d = self.indirection
        return (d >> 22) & 1023, (d >> 12) & 1023, (d >> 2) & 1023

    def decode_all(self):
        """BrickGrid::lookup for every voxel of the padded extent (grid_brick.cpp:148-154), vectorised."""
        nbx, nby, nbz = self.n_bricks
        px, py, pz = self.decode_ptr()
        lo = (self.range & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float32)
        hi = (self.range >> 16).astype(np.uint16).view(np.float16).astype(np.float32)
        out = np.empty((nbz * 8, nby * 8, nbx * 8), np.float32)
        if self.atlas.size == 0:
            out[:] = np.repeat(np.repeat(np.repeat(lo, 8, 0), 8, 1), 8, 2)
            return out
        az, ay, ax = self.atlas.shape
        blocks = self.atlas.reshape(az // 8, 8, ay // 8, 8, ax // 8, 8).transpose(0, 2, 4, 1, 3, 5)
        # slots outside the (pruned) atlas only occur for empty bricks (hi == lo); clamp for the gather
        vox = blocks[np.minimum(pz, az // 8 - 1), np.minimum(py, ay // 8 - 1), np.minimum(px, ax // 8 - 1)].astype(np.float32)
        scale = np.float32(1.0) / np.float32(255.0)
        val = lo[..., None, None, None] + vox * scale * (hi - lo)[..., None, None, None]
        out[:] = val.transpose(0, 3, 1, 4, 2, 5).reshape(out.shape)
        return out


class DenseGridData:
    """voldata::DenseGrid (grid_dense.h): u8 voxels [z][y][x] + global (min, max)."""

    def __init__(self, voxels, vmin, vmax, transform=None):
        self.voxels = np.ascontiguousarray(voxels, np.uint8)
        self.min_value, self.max_value = float(vmin), float(vmax)
        self.transform = np.eye(4, dtype=np.float32) if transform is None else np.asarray(transform, np.float32).reshape(4, 4)

    def index_extent(self):
        d, h, w = self.voxels.shape
        return (w, h, d)

    def matrix(self):
        return self.transform.T.copy()


# ---------------------------------------------------------------------------------------------------
# cereal PortableBinary (little-endian) readers / writers

class _Reader:
    def __init__(self, data: bytes):
        self.b = memoryview(data)
        self.o = 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v

    def array(self, dtype, count):
        a = np.frombuffer(self.b, dtype=dtype, count=count, offset=self.o).copy()
        self.o += a.nbytes
        return a

    def buf3d(self, dtype):
        sx, sy, sz = self.take("3I")
        (n,) = self.take("Q")
        if n != sx * sy * sz:
            raise ValueError("corrupt Buf3D: count != stride product")
        return self.array(dtype, n).reshape(sz, sy, sx)


def load_brick(path) -> BrickGridData:
    with open(path, "rb") as f:
        r = _Reader(f.read())
    (endian,) = r.take("B")
    if endian != 1:
        raise ValueError("big-endian cereal archives are not supported")
    transform = np.array(r.take("16f"), np.float32).reshape(4, 4)
    n_bricks = r.take("3I")
    min_maj = r.take("2f")
    (count,) = r.take("Q")
    ind = r.buf3d(np.uint32)
    rng = r.buf3d(np.uint32)
    atlas = r.buf3d(np.uint8)
    (n_mips,) = r.take("Q")
    mips = [r.buf3d(np.uint32) for _ in range(n_mips)]
    if n_mips != 3:
        raise ValueError(f"expected 3 range mipmaps, found {n_mips}")
    ad = (atlas.shape[2], atlas.shape[1], atlas.shape[0])
    g = BrickGridData(n_bricks, ad, count, ind, rng, atlas, mips, min_maj, transform)
    check_brick_layout(g)
    return g


def check_brick_layout(g) -> None:
    """The buffer shapes BrickGrid's constructor produces (grid_brick.cpp:60-141) and the upload relies on: it copies
    n_bricks-sized blocks out of these arrays, so a corrupt file must be rejected before it gets there."""
    nb = tuple(int(v) for v in g.n_bricks)
    if any(n == 0 or n >= 1024 or n % 8 for n in nb):
        raise ValueError(f"n_bricks {nb} must be multiples of 8 in [8, 1016]")
    want = (nb[2], nb[1], nb[0])
    for name in ("indirection", "range"):
        a = np.asarray(getattr(g, name))
        if a.shape != want:
            raise ValueError(f"{name} has shape {a.shape}, n_bricks needs {want}")
    atlas = np.asarray(g.atlas)
    if atlas.ndim != 3 or any(s % 8 for s in atlas.shape) or tuple(int(v) for v in g.atlas_dim) != (atlas.shape[2], atlas.shape[1], atlas.shape[0]):
        raise ValueError(f"atlas shape {atlas.shape} does not match atlas_dim {tuple(g.atlas_dim)} / multiples of 8")
    if len(g.mips) != 3:
        raise ValueError(f"expected 3 range mipmaps, found {len(g.mips)}")
    for i, m in enumerate(g.mips):
        if np.asarray(m).shape != tuple(n >> (i + 1) for n in want):
            raise ValueError(f"range mipmap {i} has shape {np.asarray(m).shape}, n_bricks needs {tuple(n >> (i + 1) for n in want)}")


def _w_buf3d(out, a):
    sz, sy, sx = a.shape
    out.append(struct.pack("<3IQ", sx, sy, sz, a.size))
    out.append(np.ascontiguousarray(a).tobytes())


def save_brick(path, g: BrickGridData):
    out = [struct.pack("<B", 1), np.asarray(g.transform, np.float32).tobytes(), struct.pack("<3I", *g.n_bricks),
           struct.pack("<2f", *g.min_maj), struct.pack("<Q", g.brick_count)]
    _w_buf3d(out, g.indirection)
    _w_buf3d(out, g.range)
    _w_buf3d(out, g.atlas)
    out.append(struct.pack("<Q", len(g.mips)))
    for m in g.mips:
        _w_buf3d(out, m)
    with open(path, "wb") as f:
        f.write(b"".join(out))


def load_dense(path) -> DenseGridData:
    with open(path, "rb") as f:
        r = _Reader(f.read())
    (endian,) = r.take("B")
    if endian != 1:
        raise ValueError("big-endian cereal archives are not supported")
    transform = np.array(r.take("16f"), np.float32).reshape(4, 4)
    w, h, d = r.take("3I")
    vmin, vmax = r.take("2f")
    (n,) = r.take("Q")
    vox = r.array(np.uint8, n).reshape(d, h, w)
    return DenseGridData(vox, vmin, vmax, transform)


def save_dense(path, g: DenseGridData):
    d, h, w = g.voxels.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<B", 1) + np.asarray(g.transform, np.float32).tobytes() + struct.pack("<3I2fQ", w, h, d, g.min_value, g.max_value, g.voxels.size))
        f.write(g.voxels.tobytes())


# ---------------------------------------------------------------------------------------------------
# Radiance HDR

def load_hdr(path, flip=True) -> np.ndarray:
    """Returns float32 (h, w, 3). flip=True gives the bottom-up order cppgl uploads (image_load_store.cpp:15)."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0

    def line():
        nonlocal pos
        e = data.index(b"\n", pos)
        s = data[pos:e]
        pos = e + 1
        return s

    head = line()
    if head not in (b"#?RADIANCE", b"#?RGBE"):
        raise ValueError("not a Radiance HDR file")
    fmt_ok = False
    while True:
        s = line()
        if s == b"":
            break
        if s.startswith(b"FORMAT=32-bit_rle_rgbe"):
            fmt_ok = True
    if not fmt_ok:
        raise ValueError("unsupported HDR format")
    res = line().split()
    if len(res) != 4 or res[0] != b"-Y" or res[2] != b"+X":
        raise ValueError("unsupported HDR orientation")
    h, w = int(res[1]), int(res[3])
    buf = np.frombuffer(data, np.uint8, offset=pos)
    rgbe = np.empty((h, w, 4), np.uint8)
    p = 0
    if w < 8 or w >= 32768:
        rgbe[:] = buf[: h * w * 4].reshape(h, w, 4)
    else:
        for y in range(h):
            if not (buf[p] == 2 and buf[p + 1] == 2 and not (buf[p + 2] & 0x80)):
                if y != 0:
                    raise ValueError("mixed flat/RLE HDR scanlines")
                rgbe[:] = buf[: h * w * 4].reshape(h, w, 4)  # flat file
                break
            if (int(buf[p + 2]) << 8 | int(buf[p + 3])) != w:
                raise ValueError("corrupt HDR scanline width")
            p += 4
            for c in range(4):
                x = 0
                while x < w:
                    n = int(buf[p]); p += 1
                    if n > 128:
                        n -= 128
                        rgbe[y, x:x + n, c] = buf[p]; p += 1
                    else:
                        rgbe[y, x:x + n, c] = buf[p:p + n]; p += n
                    x += n
    e = rgbe[..., 3].astype(np.int32)
    # stb_image: f = ldexp(1.0f, e - (128 + 8)); rgb = byte * f (exact in fp32); e == 0 -> 0
    scale = np.where(e != 0, np.ldexp(np.float32(1.0), e - 136), np.float32(0)).astype(np.float32)
    out = rgbe[..., :3].astype(np.float32) * scale[..., None]
    return np.ascontiguousarray(out[::-1] if flip else out)


# ---------------------------------------------------------------------------------------------------
# transfer functions

def load_lut_txt(path) -> np.ndarray:
    """`%f, %f, %f, %f` per line (transferfunc.cpp:79-93)."""
    rows = []
    with open(path) as f:
        for ln in f.read().splitlines():
            vals = ln.replace(",", " ").split()[:4]
            if len(vals) == 4:
                rows.append([np.float32(v) for v in vals])
    return np.array(rows, np.float32).reshape(-1, 4)


def lut_for_upload(lut) -> np.ndarray:
    """TransferFunction::upload_gpu (transferfunc.cpp:45-58): CDF rewrite iff alpha is not monotone."""
    lut = np.ascontiguousarray(lut, np.float32).reshape(-1, 4).copy()
    a = lut[:, 3]
    if not np.any(a[:-1] > a[1:]):
        return lut
    acc = np.float32(0)
    cdf = np.empty_like(a)
    for i in range(len(a)):  # sequential fp32 prefix sum, as compute_lut_cdf (:33-43)
        acc = np.float32(acc + a[i]) if i else a[0]
        cdf[i] = acc
    integral = cdf[-1]
    if integral <= 0:
        cdf = (np.arange(len(a), dtype=np.float32) + np.float32(1)) / np.float32(len(a))
    else:
        cdf = cdf / integral
    lut[:, 3] = cdf.astype(np.float32)
    return lut
