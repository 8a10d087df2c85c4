"""ctypes binding of libvrb200.so (include/vrb200.h). Fails loudly when the CUDA library is missing:
there is no CPU fallback in the product."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VRB200_LIB") or os.path.join(PKG, "libvrb200.so")   # VRB200_LIB: tuning builds (tools/sweep.py)

VRB_OK = 0
VRB_ERR_INVALID, VRB_ERR_NO_DEVICE, VRB_ERR_CUDA, VRB_ERR_OOM, VRB_ERR_TOO_MANY_BRICKS, VRB_ERR_STATE = -1, -2, -3, -4, -5, -6
SLOT_DENSITY, SLOT_EMISSION = 0, 1
ACCUM_MEAN, ACCUM_SUM = 0, 1

# every symbol include/vrb200.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "vrb_create", "vrb_destroy", "vrb_last_error", "vrb_status_string", "vrb_abi_version", "vrb_set_stream", "vrb_sync",
    "vrb_resize", "vrb_grid_clear", "vrb_grid_free", "vrb_grid_upload_brick", "vrb_grid_build_from_dense",
    "vrb_grid_build_from_dense_device", "vrb_grid_build_from_float_device", "vrb_brick_lattice", "vrb_grid_build_from_values", "vrb_nvdb_open", "vrb_nvdb_lookup", "vrb_grid_build_from_nvdb", "vrb_grid_info", "vrb_grid_download", "vrb_debug_sample_density", "vrb_dense_from_float",
    "vrb_env_upload", "vrb_env_download_impmap", "vrb_tf_upload", "vrb_trace", "vrb_trace_deterministic", "vrb_set_kernel", "vrb_set_option", "vrb_get_stat", "vrb_probe_bandwidth", "vrb_scale",
    "vrb_clear", "vrb_set_counting", "vrb_get_counters", "vrb_tonemap", "vrb_download_color", "vrb_download_color_ldr",
    "vrb_download_framebuffer", "vrb_upload_color", "vrb_color_device_ptr", "vrb_bind_color", "vrb_reduce", "vrb_copy_rows",
]


class VrbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"vrb200 error {status}: {message}")
        self.status = status


class Params(C.Structure):
    """vrb_params: the uniform block of RendererOpenGL::trace (src/renderer.cpp:88-139)."""
    _fields_ = [
        ("bounces", C.c_int32), ("seed", C.c_int32), ("show_environment", C.c_int32), ("frame", C.c_int32),
        ("cam_pos", C.c_float * 3), ("cam_fov", C.c_float), ("cam_transform", C.c_float * 9),
        ("vol_bb_min", C.c_float * 3), ("vol_bb_max", C.c_float * 3),
        ("vol_minorant", C.c_float), ("vol_majorant", C.c_float), ("vol_inv_majorant", C.c_float),
        ("vol_albedo", C.c_float * 3), ("vol_phase_g", C.c_float), ("vol_density_scale", C.c_float),
        ("vol_emission_scale", C.c_float), ("vol_emission_norm", C.c_float),
        ("vol_density_transform", C.c_float * 16), ("vol_density_inv_transform", C.c_float * 16),
        ("has_emission", C.c_int32),
        ("vol_emission_transform", C.c_float * 16), ("vol_emission_inv_transform", C.c_float * 16),
        ("use_transferfunc", C.c_int32), ("tf_window_left", C.c_float), ("tf_window_width", C.c_float),
        ("env_transform", C.c_float * 9), ("env_inv_transform", C.c_float * 9), ("env_strength", C.c_float),
        ("resolution", C.c_int32 * 2),
    ]

    def copy(self):
        p = Params()
        C.memmove(C.byref(p), C.byref(self), C.sizeof(Params))
        return p


class BrickView(C.Structure):
    """vrb_brick_view (voldata/src/grid_brick.h:27-33)."""
    _fields_ = [
        ("n_bricks", C.c_uint32 * 3), ("atlas_dim", C.c_uint32 * 3), ("brick_count", C.c_uint64),
        ("indirection", C.c_void_p), ("range", C.c_void_p), ("atlas", C.c_void_p),
        ("range_mips", C.c_void_p * 3),
    ]


class NvdbInfo(C.Structure):
    """vrb_nvdb_info: what voldata::NanoVDBGrid(path, gridname) derives from a grid (grid_nvdb.cpp:8-28)."""
    _fields_ = [("grid_offset", C.c_uint64), ("grid_size", C.c_uint64), ("active_voxels", C.c_uint64),
                ("extent", C.c_uint32 * 3), ("ibb_min", C.c_int32 * 3), ("minorant", C.c_float), ("majorant", C.c_float),
                ("transform", C.c_float * 16)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_samples", "n_maj", "n_dens", "n_emis", "n_nee", "n_env", "n_real")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen libvrb200.so. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -m volren_b200.build` (nvcc, sm_100a). "
            "volren_b200 has no CPU fallback.")
    L = C.CDLL(path)
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    L.vrb_create.argtypes = [ci, C.POINTER(vp)]
    L.vrb_destroy.argtypes = [vp]
    L.vrb_destroy.restype = None
    L.vrb_last_error.argtypes = [vp]
    L.vrb_last_error.restype = C.c_char_p
    L.vrb_status_string.argtypes = [ci]
    L.vrb_status_string.restype = C.c_char_p
    L.vrb_set_stream.argtypes = [vp, vp, ci]
    L.vrb_sync.argtypes = [vp]
    L.vrb_resize.argtypes = [vp, ci, ci]
    L.vrb_grid_clear.argtypes = [vp]
    L.vrb_grid_free.argtypes = [vp, ci, ci]
    L.vrb_grid_upload_brick.argtypes = [vp, ci, ci, C.POINTER(BrickView)]
    L.vrb_grid_build_from_dense.argtypes = [vp, ci, ci, vp, C.c_uint32 * 3, cf, cf]
    L.vrb_grid_build_from_dense_device.argtypes = [vp, ci, ci, vp, C.c_uint32 * 3, cf, cf]
    L.vrb_grid_build_from_float_device.argtypes = [vp, ci, ci, vp, C.c_uint32 * 3, C.c_float * 2]
    L.vrb_brick_lattice.argtypes = [C.c_uint32 * 3, C.c_uint32 * 3, C.c_uint32 * 3]
    L.vrb_grid_build_from_values.argtypes = [vp, ci, ci, vp, C.c_uint32 * 3]
    L.vrb_nvdb_open.argtypes = [vp, C.c_size_t, C.c_char_p, C.POINTER(NvdbInfo), C.c_char_p, C.c_size_t]
    L.vrb_nvdb_lookup.argtypes = [vp, C.c_int32 * 3, vp, C.c_size_t, vp]
    L.vrb_grid_build_from_nvdb.argtypes = [vp, ci, ci, vp, C.POINTER(NvdbInfo)]
    L.vrb_grid_info.argtypes = [vp, ci, ci, C.POINTER(BrickView)]
    L.vrb_grid_download.argtypes = [vp, ci, ci, C.POINTER(BrickView)]
    L.vrb_debug_sample_density.argtypes = [vp, ci, ci, vp, C.c_size_t, ci, vp]
    L.vrb_dense_from_float.argtypes = [vp, vp, C.c_uint32 * 3, vp, C.c_float * 2]
    L.vrb_env_upload.argtypes = [vp, vp, ci, ci]
    L.vrb_env_download_impmap.argtypes = [vp, ci, vp]
    L.vrb_tf_upload.argtypes = [vp, vp, C.c_uint32]
    L.vrb_trace.argtypes = [vp, C.POINTER(Params), ci, ci, vp, ci]
    L.vrb_trace_deterministic.argtypes = [vp, C.POINTER(Params)]
    L.vrb_set_kernel.argtypes = [vp, ci]
    L.vrb_set_option.argtypes = [vp, C.c_char_p, ci]
    L.vrb_get_stat.argtypes = [vp, C.c_char_p, C.POINTER(C.c_uint64)]
    L.vrb_probe_bandwidth.argtypes = [vp, C.c_size_t, ci, C.POINTER(C.c_double)]
    L.vrb_scale.argtypes = [vp, cf]
    L.vrb_clear.argtypes = [vp]
    L.vrb_set_counting.argtypes = [vp, ci]
    L.vrb_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.vrb_tonemap.argtypes = [vp, cf, cf, ci, ci]
    L.vrb_download_color.argtypes = [vp, vp, ci]
    L.vrb_download_color_ldr.argtypes = [vp, vp]
    L.vrb_download_framebuffer.argtypes = [vp, vp]
    L.vrb_upload_color.argtypes = [vp, vp]
    L.vrb_color_device_ptr.argtypes = [vp]
    L.vrb_color_device_ptr.restype = vp
    L.vrb_bind_color.argtypes = [vp, vp]
    L.vrb_reduce.argtypes = [C.POINTER(vp), ci, ci]
    L.vrb_copy_rows.argtypes = [vp, vp, ci, ci]
    _lib = L
    return L


def brick_lattice(extent_whd):
    """(n_bricks, padded_dim) of a grid with the given index extent: roundup8(ceil(extent / 8)) (grid_brick.cpp:62), 8 nb + 4."""
    nb, pd = (C.c_uint32 * 3)(), (C.c_uint32 * 3)()
    st = load_library().vrb_brick_lattice((C.c_uint32 * 3)(*extent_whd), nb, pd)
    if st:
        raise VrbError(st, "exceeded max brick count of 1024")
    return tuple(nb), tuple(pd)


class NanoVDBGridData:
    """voldata::NanoVDBGrid (voldata/src/grid_nvdb.h:13-39) over the bytes of a .nvdb file: the grid named `gridname`
    located and validated by vrb_nvdb_open; lookup() is the host accessor of the C ABI. Raises like the reference throws."""

    def __init__(self, file_bytes, gridname="density"):
        raw = np.frombuffer(file_bytes, np.uint8) if not isinstance(file_bytes, np.ndarray) else file_bytes.view(np.uint8).reshape(-1)
        self._file = np.ascontiguousarray(raw)
        self.info = NvdbInfo()
        err = C.create_string_buffer(512)
        st = load_library().vrb_nvdb_open(_ptr(self._file), self._file.size, gridname.encode(), C.byref(self.info), err, len(err))
        if st:
            raise VrbError(st, err.value.decode(errors="replace"))
        i = self.info
        self.grid = self._file[i.grid_offset:i.grid_offset + i.grid_size]
        self.extent, self.ibb_min = tuple(i.extent), tuple(i.ibb_min)
        self.min_maj = (float(i.minorant), float(i.majorant))
        self.num_voxels, self.size_bytes = int(i.active_voxels), int(i.grid_size)
        self.transform = np.array(list(i.transform), np.float32).reshape(4, 4)     # rows = glm columns

    def index_extent(self):
        return self.extent

    def matrix(self):
        return self.transform.T.copy()

    def lookup(self, ipos_xyz):
        """NanoVDBGrid::lookup (grid_nvdb.cpp:64-67) for an [n, 3] array of uint32 index positions (wrapped negatives allowed)."""
        ipos = np.ascontiguousarray(np.asarray(ipos_xyz).astype(np.int64) & 0xffffffff, np.uint32).reshape(-1, 3)
        out = np.empty(len(ipos), np.float32)
        st = load_library().vrb_nvdb_lookup(_ptr(self.grid), (C.c_int32 * 3)(*self.ibb_min), _ptr(ipos), len(ipos), _ptr(out))
        if st:
            raise VrbError(st, "vrb_nvdb_lookup")
        return out

    def padded_lattice(self):
        """lookup() on [-2, 8 nb + 2)^3, array [z][y][x]: the input of Context.grid_build_from_values."""
        nb, pd = brick_lattice(self.extent)
        z, y, x = np.meshgrid(np.arange(-2, pd[2] - 2), np.arange(-2, pd[1] - 2), np.arange(-2, pd[0] - 2), indexing="ij")
        return self.lookup(np.stack([x.ravel(), y.ravel(), z.ravel()], -1)).reshape(pd[2], pd[1], pd[0])


class Context:
    """One vrb_ctx (one CUDA device, one stream). Thin, exception-raising wrapper of the C ABI."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.vrb_create(device, C.byref(h))
        if st != VRB_OK:
            raise VrbError(st, self.lib.vrb_status_string(st).decode() + " (vrb_create; no CPU fallback exists)")
        self.handle = h
        self.device = device
        self.w = self.h = 0

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vrb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != VRB_OK:
            raise VrbError(st, self.lib.vrb_last_error(self.handle).decode())

    # --- plumbing ---
    def set_stream(self, cuda_stream_ptr, external=True):
        self._ck(self.lib.vrb_set_stream(self.handle, cuda_stream_ptr, int(external)))

    def sync(self):
        self._ck(self.lib.vrb_sync(self.handle))

    def resize(self, w, h):
        self._ck(self.lib.vrb_resize(self.handle, w, h))
        self.w, self.h = w, h

    # --- volume ---
    def grid_clear(self):
        self._ck(self.lib.vrb_grid_clear(self.handle))

    def grid_free(self, slot=SLOT_DENSITY, frame=0):
        self._ck(self.lib.vrb_grid_free(self.handle, slot, frame))

    def grid_upload_brick(self, grid, slot=SLOT_DENSITY, frame=0):
        """grid: any object with n_bricks, atlas_dim, brick_count, indirection, range, atlas, mips (numpy)."""
        from .formats import check_brick_layout
        check_brick_layout(grid)          # the library copies n_bricks-sized blocks out of these arrays
        v = BrickView()
        v.n_bricks[:] = grid.n_bricks
        v.atlas_dim[:] = grid.atlas_dim
        v.brick_count = grid.brick_count
        keep = [np.ascontiguousarray(grid.indirection, np.uint32), np.ascontiguousarray(grid.range, np.uint32),
                np.ascontiguousarray(grid.atlas, np.uint8)] + [np.ascontiguousarray(m, np.uint32) for m in grid.mips]
        v.indirection, v.range, v.atlas = _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2])
        for i in range(3):
            v.range_mips[i] = _ptr(keep[3 + i])
        self._ck(self.lib.vrb_grid_upload_brick(self.handle, slot, frame, C.byref(v)))

    def grid_build_from_dense(self, vox_u8, vmin, vmax, slot=SLOT_DENSITY, frame=0):
        vox = np.ascontiguousarray(vox_u8, np.uint8)
        d, h, w = vox.shape
        self._ck(self.lib.vrb_grid_build_from_dense(self.handle, slot, frame, _ptr(vox), (C.c_uint32 * 3)(w, h, d), vmin, vmax))

    def grid_build_from_dense_device(self, dev_ptr, dim_whd, vmin, vmax, slot=SLOT_DENSITY, frame=0):
        self._ck(self.lib.vrb_grid_build_from_dense_device(self.handle, slot, frame, dev_ptr, (C.c_uint32 * 3)(*dim_whd), vmin, vmax))

    def grid_build_from_float_device(self, dev_ptr, dim_whd, slot=SLOT_DENSITY, frame=0):
        """DenseGrid(float*) + BrickGrid for fp32 voxels in device memory -> (min_value, max_value)"""
        mm = (C.c_float * 2)()
        self._ck(self.lib.vrb_grid_build_from_float_device(self.handle, slot, frame, dev_ptr, (C.c_uint32 * 3)(*dim_whd), mm))
        return float(mm[0]), float(mm[1])

    def grid_build_from_values(self, padded_values, extent_whd, slot=SLOT_DENSITY, frame=0):
        """BrickGrid(const Grid&) for any source: lookup() values on the padded lattice [-2, 8 nb + 2)^3, array [z][y][x]."""
        val = np.ascontiguousarray(padded_values, np.float32)
        nb, pd = brick_lattice(extent_whd)
        if val.shape != (pd[2], pd[1], pd[0]):
            raise ValueError(f"padded lattice must have shape {(pd[2], pd[1], pd[0])}, got {val.shape}")
        self._ck(self.lib.vrb_grid_build_from_values(self.handle, slot, frame, _ptr(val), (C.c_uint32 * 3)(*extent_whd)))

    def grid_build_from_nvdb(self, nvdb: "NanoVDBGridData", slot=SLOT_DENSITY, frame=0):
        """BrickGrid(NanoVDBGrid) on the device: grid buffer upload + device accessor + any-Grid brick build."""
        self._ck(self.lib.vrb_grid_build_from_nvdb(self.handle, slot, frame, _ptr(nvdb.grid), C.byref(nvdb.info)))

    def grid_info(self, slot=SLOT_DENSITY, frame=0):
        v = BrickView()
        self._ck(self.lib.vrb_grid_info(self.handle, slot, frame, C.byref(v)))
        return tuple(v.n_bricks), tuple(v.atlas_dim), int(v.brick_count)

    def grid_download(self, slot=SLOT_DENSITY, frame=0, atlas=True):
        from .formats import BrickGridData
        nb, ad, cnt = self.grid_info(slot, frame)
        ind = np.empty((nb[2], nb[1], nb[0]), np.uint32)
        rng = np.empty_like(ind)
        atl = np.empty((ad[2], ad[1], ad[0]), np.uint8) if atlas else None
        mips = [np.empty((nb[2] >> (i + 1), nb[1] >> (i + 1), nb[0] >> (i + 1)), np.uint32) for i in range(3)]
        v = BrickView()
        v.indirection, v.range, v.atlas = _ptr(ind), _ptr(rng), _ptr(atl)
        for i in range(3):
            v.range_mips[i] = _ptr(mips[i])
        self._ck(self.lib.vrb_grid_download(self.handle, slot, frame, C.byref(v)))
        return BrickGridData(nb, ad, cnt, ind, rng, atl if atlas else np.zeros((0, ad[1], ad[0]), np.uint8), mips)

    def sample_density(self, ipos_xyz, mode=0, slot=SLOT_DENSITY, frame=0):
        """Density at index-space points through the tracer's fetch functions: mode 0 trilinear (records + u8 atlas),
        1 trilinear (decoded apron blocks, the production path), 2 nearest voxel at floor(p)."""
        pts = np.ascontiguousarray(ipos_xyz, np.float32).reshape(-1, 3)
        out = np.empty(len(pts), np.float32)
        self._ck(self.lib.vrb_debug_sample_density(self.handle, slot, frame, _ptr(pts), len(pts), mode, _ptr(out)))
        return out

    def dense_from_float(self, data):
        data = np.ascontiguousarray(data, np.float32)
        d, h, w = data.shape
        out = np.empty(data.shape, np.uint8)
        mm = (C.c_float * 2)()
        self._ck(self.lib.vrb_dense_from_float(self.handle, _ptr(data), (C.c_uint32 * 3)(w, h, d), _ptr(out), mm))
        return out, (float(mm[0]), float(mm[1]))

    # --- environment / transfer function ---
    def env_upload(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.float32)
        assert rgb.ndim == 3 and rgb.shape[2] == 3
        self._ck(self.lib.vrb_env_upload(self.handle, _ptr(rgb), rgb.shape[1], rgb.shape[0]))

    def env_download_impmap(self, level=0):
        d = 512 >> level
        out = np.empty((d, d), np.float32)
        self._ck(self.lib.vrb_env_download_impmap(self.handle, level, _ptr(out)))
        return out

    def tf_upload(self, rgba):
        rgba = np.ascontiguousarray(rgba, np.float32).reshape(-1, 4)
        self._ck(self.lib.vrb_tf_upload(self.handle, _ptr(rgba), rgba.shape[0]))

    # --- rendering ---
    def trace(self, params: Params, first_sample=1, n_samples=1, tile=None, accum_mode=ACCUM_MEAN):
        t = None if tile is None else (C.c_int * 4)(*tile)
        self._ck(self.lib.vrb_trace(self.handle, C.byref(params), first_sample, n_samples, t, accum_mode))

    def trace_deterministic(self, params: Params):
        self._ck(self.lib.vrb_trace_deterministic(self.handle, C.byref(params)))

    def probe_bandwidth(self, nbytes, mode=0):
        """GB/s of streaming (mode 0) or random 32-B sector gather (mode 1) reads over a working set of nbytes on this device."""
        out = C.c_double()
        self._ck(self.lib.vrb_probe_bandwidth(self.handle, C.c_size_t(nbytes), int(mode), C.byref(out)))
        return float(out.value)

    def set_kernel(self, kind):
        self._ck(self.lib.vrb_set_kernel(self.handle, kind))

    def set_option(self, name, value):
        self._ck(self.lib.vrb_set_option(self.handle, name.encode(), int(value)))

    def get_stat(self, name):
        v = C.c_uint64()
        self._ck(self.lib.vrb_get_stat(self.handle, name.encode(), C.byref(v)))
        return int(v.value)

    def scale(self, s):
        self._ck(self.lib.vrb_scale(self.handle, s))

    def clear(self):
        self._ck(self.lib.vrb_clear(self.handle))

    def set_counting(self, enable=True):
        self._ck(self.lib.vrb_set_counting(self.handle, int(enable)))

    def get_counters(self) -> Counters:
        c = Counters()
        self._ck(self.lib.vrb_get_counters(self.handle, C.byref(c)))
        return c

    def tonemap(self, exposure, gamma, in_place=True, tonemapping=True):
        self._ck(self.lib.vrb_tonemap(self.handle, exposure, gamma, int(in_place), int(tonemapping)))

    def download_color(self, channels=4):
        out = np.empty((self.h, self.w, channels), np.float32)
        self._ck(self.lib.vrb_download_color(self.handle, _ptr(out), channels))
        return out

    def download_color_ldr(self):
        out = np.empty((self.h, self.w, 4), np.uint8)
        self._ck(self.lib.vrb_download_color_ldr(self.handle, _ptr(out)))
        return out

    def download_framebuffer(self):
        out = np.empty((self.h, self.w, 4), np.uint8)
        self._ck(self.lib.vrb_download_framebuffer(self.handle, _ptr(out)))
        return out

    def upload_color(self, rgba):
        rgba = np.ascontiguousarray(rgba, np.float32)
        assert rgba.shape == (self.h, self.w, 4)
        self._ck(self.lib.vrb_upload_color(self.handle, _ptr(rgba)))

    def color_device_ptr(self):
        return self.lib.vrb_color_device_ptr(self.handle)

    def bind_color(self, dev_ptr):
        self._ck(self.lib.vrb_bind_color(self.handle, dev_ptr))


def reduce_contexts(ctxs, root=0):
    lib = load_library()
    arr = (C.c_void_p * len(ctxs))(*[c.handle for c in ctxs])
    st = lib.vrb_reduce(arr, len(ctxs), root)
    if st != VRB_OK:
        raise VrbError(st, lib.vrb_last_error(ctxs[root].handle).decode())
