"""Python mirror of the host's colour maps (volren_b200/host/transferfunc.cpp `colormap::GetColor`, i.e. tinycolormap::GetColor
as reference src/transferfunc.cpp:69-77 calls it). Reads the same generated tables (host/colormap_tables.inc)."""
from __future__ import annotations

import math
import os
import re

import numpy as np

TYPES = ["Parula", "Heat", "Jet", "Turbo", "Hot", "Gray", "Magma", "Inferno", "Plasma", "Viridis", "Cividis", "Github", "Cubehelix", "HSV"]
_TABLES: dict[str, np.ndarray] = {}


def _tables():
    if not _TABLES:
        inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "colormap_tables.inc")
        for m in re.finditer(r"colormap_(\w+)\[\d+\] = \{ ([\d,]+) \};", open(inc).read()):
            _TABLES[m.group(1)] = np.array([int(v) for v in m.group(2).split(",")], np.float64).reshape(-1, 3) / 1e6
    return _TABLES


def get_color(x: float, name: str):
    """tinycolormap::GetColor(x, type) in double precision -> (r, g, b)."""
    xc = 0.0 if x < 0.0 else 1.0 if x > 1.0 else float(x)
    if name == "Hot":
        if xc < 0.4:
            return (xc / 0.4, 0.0, 0.0)
        if xc < 0.8:
            return (1.0, (xc - 0.4) / (0.8 - 0.4), 0.0)
        return (1.0, 1.0, (xc - 0.8) / (1.0 - 0.8))
    if name == "Gray":
        return (1.0 - xc,) * 3
    data = _tables()[name.lower()]
    a = xc * (len(data) - 1)
    i = math.floor(a)
    t = a - i
    c0, c1 = data[int(i)], data[int(math.ceil(a))]
    return tuple((1.0 - t) * c0 + t * c1)


def colormap_lut(name: str, n_bins: int = 256) -> np.ndarray:
    """TransferFunction::colormap(type, n_bins) (transferfunc.cpp:69-77): rgb = GetColor(float(i) / n_bins), alpha = that float."""
    out = np.zeros((n_bins, 4), np.float32)
    for i in range(n_bins):
        f = np.float32(i) / np.float32(n_bins)
        out[i, :3] = np.array(get_color(float(f), name), np.float64).astype(np.float32)
        out[i, 3] = f
    return out
