"""Generates tests/golden/nvdb_golden.npz: a two-grid .nvdb file written by the reference's own NanoVDB headers
(submodules/voldata/submodules/openvdb/nanovdb, v32.7) and what the UNMODIFIED reference adapter voldata::NanoVDBGrid
(voldata/src/grid_nvdb.cpp) + voldata::BrickGrid(const Grid&) (grid_brick.cpp:60-142) make of it -- through
oracle/_ref/libvoldata_ref.so. Run in the build container only (needs /root/reference):

    PYTHONPATH=. python tests/golden/make_nvdb_golden.py
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.binding import VoldataRef  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fog_points(rng, lo, hi, keep, scale, negative_zero=False):
    """Active voxels of a sparse fog volume inside the box [lo, hi): a smooth blob with holes, values in (0, scale]."""
    x, y, z = np.meshgrid(np.arange(lo[0], hi[0]), np.arange(lo[1], hi[1]), np.arange(lo[2], hi[2]), indexing="ij")
    c = (np.array(lo) + np.array(hi)) / 2.0
    r = np.sqrt(((x - c[0]) / (hi[0] - lo[0])) ** 2 + ((y - c[1]) / (hi[1] - lo[1])) ** 2 + ((z - c[2]) / (hi[2] - lo[2])) ** 2)
    v = np.clip(0.55 - r, 0, None) * (0.5 + rng.random(x.shape))
    mask = (v > 0) & (rng.random(x.shape) < keep)
    ijk = np.stack([x[mask], y[mask], z[mask]], -1).astype(np.int32)
    val = (v[mask] / v.max() * scale).astype(np.float32)
    if negative_zero:
        val[::97] = -0.0          # active voxels holding -0: the first-seen order of std::min decides the sign of the range
    return ijk, val


def main():
    ref = VoldataRef()
    rng = np.random.default_rng(2024)
    # density: spans negative index coordinates, several leaf / lower nodes, extent not a multiple of 8
    d_ijk, d_val = fog_points(rng, (-21, 3, -7), (38, 45, 30), 0.7, 2.5, negative_zero=True)
    # temperature: a second grid in the same file (second segment), other box / voxel size
    t_ijk, t_val = fog_points(rng, (130, 120, 4090), (150, 141, 4110), 0.9, 900.0)   # straddles an upper-node boundary in z (4096)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "fixture.nvdb")
        ref.nvdb_write(path, [("density", d_ijk, d_val, 0.0, 0.25, (1.0, -2.0, 0.5)),
                              ("temperature", t_ijk, t_val, 0.0, 0.5, (0.0, 0.0, 0.0))])
        raw = np.fromfile(path, np.uint8)
        out = {"nvdb_file": raw}
        for name in ("density", "temperature"):
            g = ref.nvdb_load(path, name)
            assert g is not None and g["brick"] is not None, name
            b = g["brick"]
            out[name + ".extent"] = np.array(g["extent"], np.uint32)
            out[name + ".ibb_min"] = np.array(g["ibb_min"], np.int32)
            out[name + ".min_maj"] = np.array(g["min_maj"], np.float32)
            out[name + ".transform"] = g["transform"]
            out[name + ".padded"] = g["padded"]
            out[name + ".n_bricks"] = np.array(b.n_bricks, np.uint32)
            out[name + ".atlas_dim"] = np.array(b.atlas_dim, np.uint32)
            out[name + ".brick_count"] = np.array([b.brick_count], np.uint64)
            out[name + ".indirection"] = b.indirection
            out[name + ".range"] = b.range
            out[name + ".atlas"] = b.atlas
            for i in range(3):
                out[name + f".mip{i}"] = b.mips[i]
            print(name, "extent", g["extent"], "ibb_min", g["ibb_min"], "min_maj", g["min_maj"], "bricks", b.n_bricks, b.brick_count,
                  "active", len(d_val if name == "density" else t_val))
        assert ref.nvdb_load(path, "nope") is None          # unknown grid name throws in the reference
        # NanoVDB's own fog-volume sphere: the interior is stored as ACTIVE CONSTANT TILES of the lower internal nodes
        spath = os.path.join(tmp, "sphere.nvdb")
        ref.nvdb_write_fog_sphere(spath, 22.0, (3.0, -5.0, 60.0), voxel_size=1.0, half_width=3.0, name="density")
        out["sphere_file"] = np.fromfile(spath, np.uint8)
        g = ref.nvdb_load(spath, "density")
        b = g["brick"]
        name = "sphere"
        out[name + ".extent"] = np.array(g["extent"], np.uint32)
        out[name + ".ibb_min"] = np.array(g["ibb_min"], np.int32)
        out[name + ".min_maj"] = np.array(g["min_maj"], np.float32)
        out[name + ".transform"] = g["transform"]
        out[name + ".padded"] = g["padded"]
        out[name + ".n_bricks"] = np.array(b.n_bricks, np.uint32)
        out[name + ".atlas_dim"] = np.array(b.atlas_dim, np.uint32)
        out[name + ".brick_count"] = np.array([b.brick_count], np.uint64)
        out[name + ".indirection"] = b.indirection
        out[name + ".range"] = b.range
        out[name + ".atlas"] = b.atlas
        for i in range(3):
            out[name + f".mip{i}"] = b.mips[i]
        import struct
        tiles = struct.unpack_from("<3I", out["sphere_file"].tobytes(), 16 + 176 + 8 + 672 + 44)
        print("sphere extent", g["extent"], "ibb_min", g["ibb_min"], "min_maj", g["min_maj"], "bricks", b.n_bricks, b.brick_count, "active tiles (lower, upper, root)", tiles)
        assert tiles[0] > 0
    np.savez_compressed(os.path.join(HERE, "nvdb_golden.npz"), **out)
    print("file bytes", raw.size, "npz bytes", os.path.getsize(os.path.join(HERE, "nvdb_golden.npz")))


if __name__ == "__main__":
    main()
