"""ctypes binding of the CPU oracle (oracle/vr_oracle.c) and, when present, of the compiled
reference voldata sources (oracle/_ref/libvoldata_ref.so).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from volren_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libvr_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libvoldata_ref.so")

IMP_DIM = 512
IMP_LEVELS = 10


def build(ref: bool = True) -> None:
    """(Re)build the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "all" if ref else os.path.join(HERE, "_build", "libvr_oracle.so")],
                   check=True)


from volren_b200._capi import BrickView, Counters, Params  # shared PODs (include/vrb200.h)
from volren_b200.formats import BrickGridData as _BGD


class _Grid(C.Structure):
    _fields_ = [
        ("n_bricks", C.c_uint32 * 3), ("atlas_dim", C.c_uint32 * 3),
        ("indirection", C.c_void_p), ("range", C.c_void_p), ("atlas", C.c_void_p), ("range_mips", C.c_void_p * 3),
    ]


class _Scene(C.Structure):
    _fields_ = [
        ("density", _Grid), ("emission", _Grid),
        ("env_rgb", C.c_void_p), ("env_w", C.c_int32), ("env_h", C.c_int32),
        ("impmap", C.c_void_p), ("tf_lut", C.c_void_p), ("tf_size", C.c_uint32),
    ]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def mip_dims(n_bricks, level):
    return tuple(int(n) >> (level + 1) for n in n_bricks)


class BrickGridData(_BGD):
    """formats.BrickGridData + the ctypes views the oracle needs."""

    def view(self) -> BrickView:
        v = BrickView()
        v.n_bricks[:] = self.n_bricks
        v.atlas_dim[:] = self.atlas_dim
        v.brick_count = self.brick_count
        v.indirection = _ptr(self.indirection)
        v.range = _ptr(self.range)
        v.atlas = _ptr(self.atlas)
        for i in range(3):
            v.range_mips[i] = _ptr(self.mips[i])
        return v

    @staticmethod
    def grid_of(g) -> "_Grid":
        out = _Grid()
        out.n_bricks[:] = g.n_bricks
        out.atlas_dim[:] = g.atlas_dim
        out.indirection = _ptr(g.indirection)
        out.range = _ptr(g.range)
        out.atlas = _ptr(g.atlas)
        for i in range(3):
            out.range_mips[i] = _ptr(g.mips[i])
        return out


def alloc_brick_arrays(n_bricks, atlas_dim=None):
    nbx, nby, nbz = (int(v) for v in n_bricks)
    ad = atlas_dim if atlas_dim is not None else (nbx * 8, nby * 8, nbz * 8)
    ind = np.zeros((nbz, nby, nbx), np.uint32)
    rng = np.zeros((nbz, nby, nbx), np.uint32)
    atlas = np.zeros((int(ad[2]), int(ad[1]), int(ad[0])), np.uint8)
    mips = [np.zeros((nbz >> (i + 1), nby >> (i + 1), nbx >> (i + 1)), np.uint32) for i in range(3)]
    return ind, rng, atlas, mips


class Oracle:
    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.vro_tea.restype = C.c_uint32
        L.vro_tea.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.vro_rng_stream.argtypes = [C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.vro_to_half.restype = C.c_uint16
        L.vro_to_half.argtypes = [C.c_float]
        L.vro_from_half.restype = C.c_float
        L.vro_from_half.argtypes = [C.c_uint16]
        L.vro_dense_from_float.argtypes = [C.c_void_p, C.c_uint32 * 3, C.c_void_p, C.c_float * 2]
        L.vro_brick_dims.argtypes = [C.c_uint32 * 3, C.c_uint32 * 3]
        L.vro_brick_build.argtypes = [C.c_void_p, C.c_uint32 * 3, C.c_float, C.c_float, C.POINTER(BrickView)]
        L.vro_brick_build_values.argtypes = [C.c_void_p, C.c_uint32 * 3, C.POINTER(BrickView)]
        L.vro_brick_lookup.restype = C.c_float
        L.vro_lut_upload.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.vro_env_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.vro_env_pyramid_floats.restype = C.c_size_t
        L.vro_trace.argtypes = [C.POINTER(_Scene), C.POINTER(Params), C.c_int, C.c_int, C.c_void_p, C.c_int,
                                C.c_void_p, C.POINTER(Counters), C.c_int]
        L.vro_trace_deterministic.argtypes = [C.POINTER(_Scene), C.POINTER(Params), C.c_void_p, C.c_int]
        L.vro_tonemap_inplace.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]
        L.vro_draw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p]
        L.vro_color_to_ldr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]

    # --- integer / bit-exact layer ---
    def tea(self, v0, v1, n=32):
        return int(self.lib.vro_tea(v0 & 0xFFFFFFFF, v1 & 0xFFFFFFFF, n))

    def rng_stream(self, seed, n):
        out = np.empty(n, np.float32)
        states = np.empty(n, np.uint32)
        self.lib.vro_rng_stream(seed & 0xFFFFFFFF, n, _ptr(out), _ptr(states))
        return out, states

    def to_half(self, f):
        return int(self.lib.vro_to_half(float(f)))

    def from_half(self, h):
        return float(self.lib.vro_from_half(int(h)))

    def dense_from_float(self, data):
        data = np.ascontiguousarray(data, np.float32)
        d, h, w = data.shape
        out = np.empty(data.shape, np.uint8)
        mm = (C.c_float * 2)()
        self.lib.vro_dense_from_float(_ptr(data), (C.c_uint32 * 3)(w, h, d), _ptr(out), mm)
        return out, (float(mm[0]), float(mm[1]))

    def brick_dims(self, dim):
        nb = (C.c_uint32 * 3)()
        st = self.lib.vro_brick_dims((C.c_uint32 * 3)(*dim), nb)
        return st, tuple(nb)

    def brick_build(self, vox_u8, vmin, vmax) -> BrickGridData:
        vox = np.ascontiguousarray(vox_u8, np.uint8)
        d, h, w = vox.shape
        st, nb = self.brick_dims((w, h, d))
        if st:
            raise RuntimeError("exceeded max brick count of 1024")
        ind, rng, atlas, mips = alloc_brick_arrays(nb)
        view = BrickView()
        view.indirection, view.range, view.atlas = _ptr(ind), _ptr(rng), _ptr(atlas)
        for i in range(3):
            view.range_mips[i] = _ptr(mips[i])
        st = self.lib.vro_brick_build(_ptr(vox), (C.c_uint32 * 3)(w, h, d), vmin, vmax, C.byref(view))
        assert st == 0
        ad = tuple(view.atlas_dim)
        atlas = np.ascontiguousarray(atlas[: ad[2]])
        return BrickGridData(nb, ad, view.brick_count, ind, rng, atlas, mips, (vmin, vmax))

    def brick_build_values(self, padded_values, extent_whd, min_maj=(0.0, 0.0)) -> BrickGridData:
        """BrickGrid(const Grid&) for any Grid: lookup() values on the padded lattice [-2, 8 nb + 2)^3, array [z][y][x]."""
        val = np.ascontiguousarray(padded_values, np.float32)
        st, nb = self.brick_dims(tuple(extent_whd))
        if st:
            raise RuntimeError("exceeded max brick count of 1024")
        assert val.shape == (nb[2] * 8 + 4, nb[1] * 8 + 4, nb[0] * 8 + 4), val.shape
        ind, rng, atlas, mips = alloc_brick_arrays(nb)
        view = BrickView()
        view.indirection, view.range, view.atlas = _ptr(ind), _ptr(rng), _ptr(atlas)
        for i in range(3):
            view.range_mips[i] = _ptr(mips[i])
        st = self.lib.vro_brick_build_values(_ptr(val), (C.c_uint32 * 3)(*extent_whd), C.byref(view))
        assert st == 0
        ad = tuple(view.atlas_dim)
        return BrickGridData(nb, ad, view.brick_count, ind, rng, np.ascontiguousarray(atlas[: ad[2]]), mips, min_maj)

    def lut_upload(self, rgba):
        rgba = np.ascontiguousarray(rgba, np.float32).reshape(-1, 4)
        out = np.empty_like(rgba)
        changed = self.lib.vro_lut_upload(_ptr(rgba), rgba.shape[0], _ptr(out))
        return out, bool(changed)

    # --- shader layer ---
    def env_build(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.float32)
        h, w, _ = rgb.shape
        pyr = np.empty(self.lib.vro_env_pyramid_floats(), np.float32)
        self.lib.vro_env_build(_ptr(rgb), w, h, _ptr(pyr))
        return pyr

    @staticmethod
    def pyramid_level(pyr, level):
        off = sum((IMP_DIM >> l) ** 2 for l in range(level))
        d = IMP_DIM >> level
        return pyr[off: off + d * d].reshape(d, d)

    def make_scene(self, density: BrickGridData, env_rgb, impmap, lut=None, emission: BrickGridData | None = None):
        sc = _Scene()
        sc.density = BrickGridData.grid_of(density)
        if emission is not None:
            sc.emission = BrickGridData.grid_of(emission)
        env_rgb = np.ascontiguousarray(env_rgb, np.float32)
        sc.env_rgb = _ptr(env_rgb)
        sc.env_h, sc.env_w = env_rgb.shape[0], env_rgb.shape[1]
        sc.impmap = _ptr(impmap)
        keep = [density, emission, env_rgb, impmap]
        if lut is not None:
            lut = np.ascontiguousarray(lut, np.float32).reshape(-1, 4)
            sc.tf_lut = _ptr(lut)
            sc.tf_size = lut.shape[0]
            keep.append(lut)
        sc._keep = keep
        return sc

    def trace(self, scene, params: Params, first_sample, n_samples, color=None, tile=None, accum_mode=0,
              n_threads=0):
        W, H = params.resolution[0], params.resolution[1]
        if color is None:
            color = np.zeros((H, W, 4), np.float32)
        cnt = Counters()
        t = None if tile is None else (C.c_int * 4)(*tile)
        self.lib.vro_trace(C.byref(scene), C.byref(params), first_sample, n_samples, t, accum_mode, _ptr(color),
                           C.byref(cnt), n_threads)
        return color, cnt

    def trace_deterministic(self, scene, params: Params, n_threads=0):
        W, H = params.resolution[0], params.resolution[1]
        color = np.zeros((H, W, 4), np.float32)
        self.lib.vro_trace_deterministic(C.byref(scene), C.byref(params), _ptr(color), n_threads)
        return color

    def tonemap_inplace(self, color, exposure, gamma):
        color = np.ascontiguousarray(color, np.float32).copy()
        self.lib.vro_tonemap_inplace(_ptr(color), color.shape[1], color.shape[0], exposure, gamma)
        return color

    def draw(self, color, exposure, gamma, tonemapping=True):
        color = np.ascontiguousarray(color, np.float32)
        out = np.empty(color.shape[:2] + (4,), np.uint8)
        self.lib.vro_draw(_ptr(color), color.shape[1], color.shape[0], exposure, gamma, int(tonemapping), _ptr(out))
        return out

    def color_to_ldr(self, color):
        color = np.ascontiguousarray(color, np.float32)
        out = np.empty(color.shape[:2] + (4,), np.uint8)
        self.lib.vro_color_to_ldr(_ptr(color), color.shape[1], color.shape[0], _ptr(out))
        return out

    def max_threads(self):
        return int(self.lib.vro_max_threads())


class VoldataRef:
    """The compiled, unmodified reference voldata (oracle/_ref). None-able: check available()."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        L = self.lib = C.CDLL(REF_SO)
        L.ref_to_half.restype = C.c_uint16
        L.ref_to_half.argtypes = [C.c_float]
        L.ref_from_half.restype = C.c_float
        L.ref_from_half.argtypes = [C.c_uint16]
        for name in ("ref_dense_from_float", "ref_dense_from_u8"):
            getattr(L, name).restype = C.c_void_p
            getattr(L, name).argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.ref_dense_set_range.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.ref_dense_info.argtypes = [C.c_void_p, C.c_uint32 * 3, C.c_float * 2]
        L.ref_dense_voxels.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_dense_free.argtypes = [C.c_void_p]
        L.ref_dense_load.restype = C.c_void_p
        L.ref_dense_load.argtypes = [C.c_char_p]
        L.ref_dense_write.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_brick_from_dense.restype = C.c_void_p
        L.ref_brick_from_dense.argtypes = [C.c_void_p]
        L.ref_brick_load.restype = C.c_void_p
        L.ref_brick_load.argtypes = [C.c_char_p]
        L.ref_brick_write.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_brick_info.argtypes = [C.c_void_p, C.c_uint32 * 3, C.c_uint32 * 3, C.c_float * 2,
                                     C.POINTER(C.c_uint64), C.c_float * 16, C.POINTER(C.c_uint32)]
        L.ref_brick_copy.argtypes = [C.c_void_p] * 7
        L.ref_brick_decode_all.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_brick_free.argtypes = [C.c_void_p]
        # NanoVDB adapter of the reference (grid_nvdb.cpp) + generic Grid* access
        L.ref_nvdb_build.restype = C.c_void_p
        L.ref_nvdb_build.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_char_p, C.c_float, C.c_double, C.c_double * 3]
        L.ref_nvdb_write.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_char_p]
        L.ref_nvdb_fog_sphere.restype = C.c_void_p
        L.ref_nvdb_fog_sphere.argtypes = [C.c_double, C.c_double * 3, C.c_double, C.c_double, C.c_char_p]
        L.ref_nvdb_load.restype = C.c_void_p
        L.ref_nvdb_load.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_nvdb_ibb_min.argtypes = [C.c_void_p, C.c_int32 * 3]
        L.ref_grid_info.argtypes = [C.c_void_p, C.c_uint32 * 3, C.c_float * 2, C.c_float * 16]
        L.ref_grid_lookup_padded.argtypes = [C.c_void_p, C.c_uint32 * 3, C.c_void_p]
        L.ref_brick_from_grid.restype = C.c_void_p
        L.ref_brick_from_grid.argtypes = [C.c_void_p]
        L.ref_grid_free.argtypes = [C.c_void_p]
        L.ref_brick_from_values.restype = C.c_void_p
        L.ref_brick_from_values.argtypes = [C.c_void_p, C.c_uint32 * 3, C.c_uint32 * 3]
        L.ref_colormap_lut.argtypes = [C.c_int, C.c_uint32, C.c_void_p]

    def colormap_lut(self, type_index, n_bins=256):
        """TransferFunction::colormap(type, n_bins) (src/transferfunc.cpp:69-77) with the reference's own tinycolormap."""
        out = np.empty((n_bins, 4), np.float32)
        self.lib.ref_colormap_lut(int(type_index), int(n_bins), _ptr(out))
        return out

    def to_half(self, f):
        return int(self.lib.ref_to_half(float(f)))

    def from_half(self, h):
        return float(self.lib.ref_from_half(int(h)))

    def dense_from_float(self, data):
        data = np.ascontiguousarray(data, np.float32)
        d, h, w = data.shape
        g = self.lib.ref_dense_from_float(w, h, d, _ptr(data))
        out = np.empty(data.shape, np.uint8)
        dim = (C.c_uint32 * 3)()
        mm = (C.c_float * 2)()
        self.lib.ref_dense_info(g, dim, mm)
        self.lib.ref_dense_voxels(g, _ptr(out))
        self.lib.ref_dense_free(g)
        return out, (float(mm[0]), float(mm[1]))

    # --- NanoVDB (grid_nvdb.cpp) ---
    def nvdb_write(self, path, grids):
        """grids: list of (name, ijk int32 [n,3], values float32 [n], background, voxel_size, origin xyz) -> one .nvdb file,
        written by the reference's NanoVDB headers (tools::createNanoGrid + io::writeGrids, uncompressed)."""
        handles = (C.c_void_p * len(grids))()
        for i, (name, ijk, values, background, voxel_size, origin) in enumerate(grids):
            ijk = np.ascontiguousarray(ijk, np.int32).reshape(-1, 3)
            values = np.ascontiguousarray(values, np.float32)
            handles[i] = self.lib.ref_nvdb_build(_ptr(ijk), _ptr(values), len(values), name.encode(), background, voxel_size, (C.c_double * 3)(*origin))
        assert self.lib.ref_nvdb_write(handles, len(grids), path.encode()) == 0

    def nvdb_write_fog_sphere(self, path, radius, center, voxel_size=1.0, half_width=3.0, name="density"):
        """nanovdb::tools::createFogVolumeSphere -> one-grid .nvdb file (interior stored as active constant tiles)."""
        h = (C.c_void_p * 1)(self.lib.ref_nvdb_fog_sphere(radius, (C.c_double * 3)(*center), voxel_size, half_width, name.encode()))
        assert self.lib.ref_nvdb_write(h, 1, path.encode()) == 0

    def nvdb_load(self, path, gridname="density"):
        """voldata::NanoVDBGrid(path, gridname) -> dict with extent, ibb_min, min_maj, transform (rows = glm columns),
        the lookup() values on the padded brick lattice and the reference's BrickGrid built from it; None if it throws."""
        g = self.lib.ref_nvdb_load(path.encode(), gridname.encode())
        if not g:
            return None
        ext, mm, tr, ibb = (C.c_uint32 * 3)(), (C.c_float * 2)(), (C.c_float * 16)(), (C.c_int32 * 3)()
        self.lib.ref_grid_info(g, ext, mm, tr)
        self.lib.ref_nvdb_ibb_min(g, ibb)
        nb = tuple(((int(np.ceil(np.float32(np.ceil(np.float32(e) / np.float32(8))) / np.float32(8)))) * 1) << 3 for e in ext)
        padded = np.empty((nb[2] * 8 + 4, nb[1] * 8 + 4, nb[0] * 8 + 4), np.float32)
        self.lib.ref_grid_lookup_padded(g, (C.c_uint32 * 3)(*nb), _ptr(padded))
        b = self.lib.ref_brick_from_grid(g)
        brick = self._brick_to_data(b) if b else None
        if b:
            self.lib.ref_brick_free(b)
        self.lib.ref_grid_free(g)
        return dict(extent=tuple(ext), ibb_min=tuple(ibb), min_maj=(float(mm[0]), float(mm[1])),
                    transform=np.array(list(tr), np.float32).reshape(4, 4), n_bricks=nb, padded=padded, brick=brick)

    def brick_build_values(self, padded_values, extent_whd, n_bricks):
        """The unmodified BrickGrid(const Grid&) on a table-backed Grid (lookup() values on the padded lattice, [z][y][x])."""
        val = np.ascontiguousarray(padded_values, np.float32)
        assert val.shape == (n_bricks[2] * 8 + 4, n_bricks[1] * 8 + 4, n_bricks[0] * 8 + 4)
        b = self.lib.ref_brick_from_values(_ptr(val), (C.c_uint32 * 3)(*extent_whd), (C.c_uint32 * 3)(*n_bricks))
        if not b:
            return None
        out = self._brick_to_data(b)
        self.lib.ref_brick_free(b)
        return out

    def _brick_to_data(self, b) -> BrickGridData:
        nb = (C.c_uint32 * 3)()
        ad = (C.c_uint32 * 3)()
        mm = (C.c_float * 2)()
        cnt = C.c_uint64()
        tr = (C.c_float * 16)()
        nm = C.c_uint32()
        self.lib.ref_brick_info(b, nb, ad, mm, C.byref(cnt), tr, C.byref(nm))
        ind, rng, atlas, mips = alloc_brick_arrays(tuple(nb), tuple(ad))
        self.lib.ref_brick_copy(b, _ptr(ind), _ptr(rng), _ptr(atlas), _ptr(mips[0]), _ptr(mips[1]), _ptr(mips[2]))
        transform = np.array(list(tr), np.float32).reshape(4, 4)  # rows of this array = glm columns
        return BrickGridData(tuple(nb), tuple(ad), cnt.value, ind, rng, atlas, mips, (mm[0], mm[1]), transform)

    def brick_build(self, vox_u8, vmin, vmax, decode=False):
        vox = np.ascontiguousarray(vox_u8, np.uint8)
        d, h, w = vox.shape
        g = self.lib.ref_dense_from_u8(w, h, d, _ptr(vox))
        self.lib.ref_dense_set_range(g, vmin, vmax)
        b = self.lib.ref_brick_from_dense(g)
        self.lib.ref_dense_free(g)
        if not b:
            raise RuntimeError("exceeded max brick count of 1024")
        data = self._brick_to_data(b)
        dec = None
        if decode:
            e = data.index_extent()
            dec = np.empty((e[2], e[1], e[0]), np.float32)
            self.lib.ref_brick_decode_all(b, _ptr(dec))
        self.lib.ref_brick_free(b)
        return (data, dec) if decode else data

    def brick_load(self, path) -> BrickGridData:
        b = self.lib.ref_brick_load(path.encode())
        if not b:
            raise RuntimeError("cannot load " + path)
        data = self._brick_to_data(b)
        self.lib.ref_brick_free(b)
        return data

    def brick_roundtrip_write(self, vox_u8, vmin, vmax, path):
        vox = np.ascontiguousarray(vox_u8, np.uint8)
        d, h, w = vox.shape
        g = self.lib.ref_dense_from_u8(w, h, d, _ptr(vox))
        self.lib.ref_dense_set_range(g, vmin, vmax)
        b = self.lib.ref_brick_from_dense(g)
        self.lib.ref_brick_write(b, path.encode())
        self.lib.ref_brick_free(b)
        self.lib.ref_dense_free(g)

    def dense_write(self, vox_u8, vmin, vmax, path):
        vox = np.ascontiguousarray(vox_u8, np.uint8)
        d, h, w = vox.shape
        g = self.lib.ref_dense_from_u8(w, h, d, _ptr(vox))
        self.lib.ref_dense_set_range(g, vmin, vmax)
        self.lib.ref_dense_write(g, path.encode())
        self.lib.ref_dense_free(g)


GLSL_REF_SO = os.path.join(HERE, "_ref", "libglsl_ref.so")


class GlslRef:
    """The reference's UNMODIFIED GLSL compute shaders compiled as C++ (oracle/glsl_ref/: glsl2cpp.py + glsl_shim.h over the
    reference's own glm) -- "the reference compiled here" for the shader layer. Only built where /root/reference exists;
    the prebuilt .so travels to the GPU box. Scenes and parameters are the oracle's (vro_scene / vrb_params)."""

    @staticmethod
    def available():
        return os.path.exists(GLSL_REF_SO)

    def __init__(self):
        L = self.lib = C.CDLL(GLSL_REF_SO)
        L.glslref_trace.argtypes = [C.POINTER(_Scene), C.POINTER(Params), C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.glslref_env_setup.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.glslref_tonemap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]
        L.glslref_tea.restype = C.c_uint32
        L.glslref_tea.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.glslref_rng.restype = C.c_float
        L.glslref_rng.argtypes = [C.POINTER(C.c_uint32)]

    def trace(self, scene, params: Params, first_sample, n_samples, color=None, n_threads=0):
        W, H = params.resolution[0], params.resolution[1]
        if color is None:
            color = np.zeros((H, W, 4), np.float32)
        self.lib.glslref_trace(C.byref(scene), C.byref(params), first_sample, n_samples, _ptr(color),
                               n_threads or (os.cpu_count() or 1))
        return color

    def env_setup(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.float32)
        h, w, _ = rgb.shape
        out = np.empty((IMP_DIM, IMP_DIM), np.float32)
        self.lib.glslref_env_setup(_ptr(rgb), w, h, _ptr(out))
        return out

    def tonemap(self, color, exposure, gamma):
        color = np.ascontiguousarray(color, np.float32).copy()
        self.lib.glslref_tonemap(_ptr(color), color.shape[1], color.shape[0], exposure, gamma)
        return color

    def tea(self, v0, v1, n=32):
        return int(self.lib.glslref_tea(v0 & 0xFFFFFFFF, v1 & 0xFFFFFFFF, n))

    def rng(self, state):
        s = C.c_uint32(state & 0xFFFFFFFF)
        v = self.lib.glslref_rng(C.byref(s))
        return float(v), int(s.value)
