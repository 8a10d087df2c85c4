"""TEST INFRASTRUCTURE ONLY (see oracle/binding.py): numpy restatement of what voldata::NanoVDBGrid (voldata/src/grid_nvdb.cpp)
reads out of a .nvdb file, written independently of the product's accessor (volren_b200/csrc/vr_nvdb.cuh): instead of a
top-down getValue per voxel, every node of the tree is PAINTED into a dense box (root tiles -> upper -> lower -> leaves).

Algorithm source: NanoVDB ABI 32 as pinned by the reference (submodules/voldata/submodules/openvdb/nanovdb, v32.7):
  io/IO.h:386-431 Segment::read, :349-356 FileGridMetaData::read, :568-594 readGrid(is, gridName)
  NanoVDB.h:1810-1830 GridData, :2262-2267 TreeData, :2512-2553 RootData/Tile (NANOVDB_USE_SINGLE_ROOT_KEY, :151),
  :3184-3202 InternalData, :3568-3573 CoordToOffset, :3749-3758 LeafData<float>
Pinned against the unmodified reference through tests/golden/nvdb_golden.npz (make_nvdb_golden.py)."""
import struct

import numpy as np

MAGIC_NUMB, MAGIC_GRID, MAGIC_FILE = 0x304244566f6e614e, 0x314244566f6e614e, 0x324244566f6e614e


def find_grid(data: bytes, gridname: str):
    """(offset, size) of the first grid called `gridname` in a segment file; KeyError if absent."""
    at = 0
    while at + 16 <= len(data):
        magic, version, count, codec = struct.unpack_from("<QIHH", data, at)
        if magic not in (MAGIC_NUMB, MAGIC_FILE):
            raise ValueError("not a NanoVDB file")
        at += 16
        metas = []
        for _ in range(count):
            grid_size, file_size, name_key, voxels = struct.unpack_from("<4Q", data, at)
            (name_size,) = struct.unpack_from("<I", data, at + 136)
            name = data[at + 176:at + 176 + name_size].split(b"\0")[0].decode()
            metas.append((name, grid_size, file_size))
            at += 176 + name_size
        for name, grid_size, file_size in metas:
            if name == gridname:
                if codec != 0:
                    raise ValueError("compressed file")
                return at, grid_size
            at += file_size
    raise KeyError(gridname)


class Grid:
    def __init__(self, data: bytes, gridname="density"):
        off, size = find_grid(bytes(data), gridname)
        self.buf = g = bytes(data)[off:off + size]
        self.grid_class, self.grid_type = struct.unpack_from("<II", g, 632)
        self.matf = np.frombuffer(g, np.float32, 9, 296)
        self.vecf = np.frombuffer(g, np.float32, 3, 368)
        tree = 672
        self.node_offset = struct.unpack_from("<4q", g, tree)
        (self.voxel_count,) = struct.unpack_from("<Q", g, tree + 56)
        self.root = tree + self.node_offset[3]
        self.bbox = np.array(struct.unpack_from("<6i", g, self.root), np.int64).reshape(2, 3)
        self.n_tiles, self.background, self.minimum, self.maximum = struct.unpack_from("<Ifff", g, self.root + 24)

    # what the constructor derives (grid_nvdb.cpp:13-27)
    def derived(self):
        empty = self.n_tiles == 0
        fmin = np.zeros(3, np.float32) if empty else self.bbox[0].astype(np.float32)
        ibb_min = fmin.astype(np.int32)
        extent = np.zeros(3, np.uint32) if empty else (self.bbox[1] - self.bbox[0] + 1).astype(np.float32).astype(np.uint32)
        T = np.zeros((4, 4), np.float32)              # T[c] = glm column c
        T[:3, :3] = self.matf.reshape(3, 3)
        T[3, :3] = self.vecf
        T[3, 3] = 1
        f = np.float32
        add = (T[0] * f(fmin[0]) + T[1] * f(fmin[1])) + (T[2] * f(fmin[2]) + T[3] * f(0))    # glm mat4 * vec4 grouping
        T[3] = T[3] + add
        return dict(extent=tuple(int(v) for v in extent), ibb_min=tuple(int(v) for v in ibb_min), min_maj=(self.minimum, self.maximum), transform=T)

    def paint(self, lo, hi):
        """getValue on the box [lo, hi) of tree coordinates -> float32 array [x][y][z] (hi - lo)."""
        lo, hi = np.asarray(lo, np.int64), np.asarray(hi, np.int64)
        out = np.full(tuple(hi - lo), np.float32(self.background), np.float32)
        g = self.buf

        def clip(origin, dim):
            a, b = np.maximum(origin, lo), np.minimum(origin + dim, hi)
            return (a, b) if np.all(a < b) else (None, None)

        def fill(origin, dim, value):
            a, b = clip(origin, dim)
            if a is not None:
                out[a[0] - lo[0]:b[0] - lo[0], a[1] - lo[1]:b[1] - lo[1], a[2] - lo[2]:b[2] - lo[2]] = value

        def internal(node, origin, log2dim, child_dim, mask_off, table_off, leaf_level):
            n = 1 << log2dim
            if clip(origin, n * child_dim)[0] is None:
                return
            words = np.frombuffer(g, np.uint64, n ** 3 // 64, node + mask_off)
            table = np.frombuffer(g, np.uint8, n ** 3 * 8, node + table_off).reshape(-1, 8)
            a, b = clip(origin, n * child_dim)
            i0, i1 = (a - origin) // child_dim, (b - 1 - origin) // child_dim + 1
            for i in range(i0[0], i1[0]):
                for j in range(i0[1], i1[1]):
                    for k in range(i0[2], i1[2]):
                        idx = (i << (2 * log2dim)) | (j << log2dim) | k
                        o = origin + np.array([i, j, k]) * child_dim
                        if (int(words[idx >> 6]) >> (idx & 63)) & 1:
                            child = node + int(table[idx].view(np.int64)[0])
                            if leaf_level:
                                vals = np.frombuffer(g, np.float32, 512, child + 96).reshape(8, 8, 8)      # [x][y][z]
                                a2, b2 = clip(o, 8)
                                if a2 is not None:
                                    s = tuple(slice(a2[d] - o[d], b2[d] - o[d]) for d in range(3))
                                    out[a2[0] - lo[0]:b2[0] - lo[0], a2[1] - lo[1]:b2[1] - lo[1], a2[2] - lo[2]:b2[2] - lo[2]] = vals[s]
                            else:
                                internal(child, o, 4, 8, 32 + 512, 1088, True)
                        else:
                            fill(o, child_dim, table[idx, :4].view(np.float32)[0])

        for t in range(self.n_tiles):
            key, child, state, value = struct.unpack_from("<QqIf", g, self.root + 64 + 32 * t)
            # KeyToCoord: 21 bits per axis, sign-extended by the uint32 << 12 wrap
            o = np.array([((key >> 42) & 0x1fffff), ((key >> 21) & 0x1fffff), key & 0x1fffff], np.int64) << 12
            o = ((o + 2 ** 31) % 2 ** 32) - 2 ** 31
            if child == 0:
                fill(o, 4096, np.float32(value))
            else:
                internal(self.root + child, o, 5, 128, 32 + 4096, 8256, False)
        return out

    def padded_lattice(self, n_bricks):
        """NanoVDBGrid::lookup on [-2, 8 nb + 2)^3 as BrickGrid(const Grid&) addresses it -> array [z][y][x]."""
        d = self.derived()
        lo = np.array(d["ibb_min"], np.int64) - 2
        hi = np.array(d["ibb_min"], np.int64) + 8 * np.array(n_bricks, np.int64) + 2
        return np.ascontiguousarray(self.paint(lo, hi).transpose(2, 1, 0))
