// NanoVDB float grids as a brick-build source: voldata::NanoVDBGrid (voldata/src/grid_nvdb.cpp:8-28,
// :64-67) without the NanoVDB headers. The reference pins NanoVDB ABI 32 (submodules/voldata/submodules/openvdb, v32.7);
// the byte layout below is that ABI's serialized grid buffer (nanovdb/NanoVDB.h: GridData :1810-1830, TreeData :2262-2267,
// RootData + Tile :2512-2553 with NANOVDB_USE_SINGLE_ROOT_KEY :151, InternalData :3184-3202, LeafData<float> :3749-3758),
// every node 32-byte aligned. A grid buffer is position independent (all links are byte offsets), so the same bytes are
// walked by the host accessor (Grid::lookup of the host model) and, after one H2D copy, by the device accessor that
// tabulates lookup() on the padded brick lattice for the brick builder (k_nvdb_tabulate).
#pragma once

#include "vr_common.cuh"

#include <cstring>

namespace vr {
namespace nvdb {

constexpr uint64_t MAGIC_NUMB = 0x304244566f6e614eull;   // "NanoVDB0"  (NanoVDB.h:134)
constexpr uint64_t MAGIC_GRID = 0x314244566f6e614eull;   // "NanoVDB1"
constexpr uint64_t MAGIC_FILE = 0x324244566f6e614eull;   // "NanoVDB2"
constexpr uint32_t ABI_MAJOR = 32;                        // Version::isCompatible (NanoVDB.h:144)
constexpr uint32_t GRID_TYPE_FLOAT = 1, GRID_TYPE_END = 27, GRID_CLASS_FOG = 2, GRID_CLASS_END = 10;

// GridData (672 bytes)
constexpr size_t GRID_SIZE = 672, G_MAGIC = 0, G_VERSION = 16, G_INDEX = 24, G_COUNT = 28, G_BYTES = 32, G_NAME = 40, G_NAME_LEN = 256,
                 G_MAP_MATF = 296, G_MAP_VECF = 368, G_CLASS = 632, G_TYPE = 636, G_DATA2 = 664;
// TreeData (64 bytes, directly behind GridData): int64 mNodeOffset[4] (leaf, lower, upper, root; relative to the tree),
// u32 mNodeCount[3], u32 mTileCount[3], u64 mVoxelCount
constexpr size_t TREE = GRID_SIZE, TREE_SIZE = 64, T_NODE_OFFSET = 0, T_NODE_COUNT = 32, T_VOXEL_COUNT = 56;
// RootData<float> (64 bytes): CoordBBox (2 x 3 x i32), u32 table size, background, minimum, maximum, average, std dev
constexpr size_t ROOT_SIZE = 64, R_BBOX = 0, R_TABLE_SIZE = 24, R_BACKGROUND = 28, R_MINIMUM = 32, R_MAXIMUM = 36;
// Root tile (32 bytes): u64 key, i64 child (byte offset from the ROOT; 0 = constant tile), u32 state, float value
constexpr size_t TILE_SIZE = 32, RT_KEY = 0, RT_CHILD = 8, RT_VALUE = 20;
// InternalData<LOG2DIM 5> (upper, 4096^3 voxels) and <LOG2DIM 4> (lower, 128^3): bbox 24, flags 8, value mask, child mask,
// min/max/avg/dev, table of 8-byte unions { float value; int64 child (offset from THIS node) } at the next 32-byte boundary
constexpr size_t UPPER_CHILD_MASK = 32 + 4096, UPPER_TABLE = 8256, UPPER_SIZE = UPPER_TABLE + 32768 * 8;
constexpr size_t LOWER_CHILD_MASK = 32 + 512, LOWER_TABLE = 1088, LOWER_SIZE = LOWER_TABLE + 4096 * 8;
// LeafData<float> (8^3): bbox min 12, dif 3, flags 1, value mask 64, min/max/avg/dev, 512 floats at byte 96
constexpr size_t LEAF_VALUES = 96, LEAF_SIZE = LEAF_VALUES + 512 * 4;

// every field is naturally aligned relative to the grid start; device copies start at an allocation boundary, a grid
// inside a host file image may start at any byte (variable-length names precede it), hence the memcpy on the host
template <typename T> VR_HD T rd(const uint8_t* p) {
#ifdef __CUDA_ARCH__
    return *reinterpret_cast<const T*>(p);
#else
    T v;
    memcpy(&v, p, sizeof(T));
    return v;
#endif
}

// RootData::CoordToKey (NanoVDB.h:2492-2499): 21 bits per axis of the upper-node coordinate, x in the top bits
VR_HD uint64_t root_key(int32_t x, int32_t y, int32_t z) {
    return uint64_t(uint32_t(z) >> 12) | (uint64_t(uint32_t(y) >> 12) << 21) | (uint64_t(uint32_t(x) >> 12) << 42);
}

// ReadAccessor::getValue without the cache == RootNode::getValue -> InternalNode::getValue (NanoVDB.h:3528-3532) ->
// LeafNode::getValue: linear probe of the root table by key, background when absent, tile/table values regardless of
// their active state, x-major node offsets (InternalNode::CoordToOffset :3568-3573).
VR_HD float get_value(const uint8_t* grid, int32_t x, int32_t y, int32_t z) {
    const uint8_t* tree = grid + TREE;
    const uint8_t* root = tree + rd<int64_t>(tree + T_NODE_OFFSET + 24);
    const uint32_t n_tiles = rd<uint32_t>(root + R_TABLE_SIZE);
    const uint64_t key = root_key(x, y, z);
    const uint8_t* tile = root + ROOT_SIZE;
    uint32_t t = 0;
    for (; t < n_tiles; ++t, tile += TILE_SIZE)
        if (rd<uint64_t>(tile + RT_KEY) == key) break;
    if (t == n_tiles) return rd<float>(root + R_BACKGROUND);
    const int64_t up = rd<int64_t>(tile + RT_CHILD);
    if (up == 0) return rd<float>(tile + RT_VALUE);
    const uint8_t* upper = root + up;
    const uint32_t nu = (((uint32_t(x) & 4095u) >> 7) << 10) | (((uint32_t(y) & 4095u) >> 7) << 5) | ((uint32_t(z) & 4095u) >> 7);
    const uint8_t* ue = upper + UPPER_TABLE + size_t(nu) * 8;
    if (!((rd<uint64_t>(upper + UPPER_CHILD_MASK + (nu >> 6) * 8) >> (nu & 63u)) & 1ull)) return rd<float>(ue);
    const uint8_t* lower = upper + rd<int64_t>(ue);
    const uint32_t nl = (((uint32_t(x) & 127u) >> 3) << 8) | (((uint32_t(y) & 127u) >> 3) << 4) | ((uint32_t(z) & 127u) >> 3);
    const uint8_t* le = lower + LOWER_TABLE + size_t(nl) * 8;
    if (!((rd<uint64_t>(lower + LOWER_CHILD_MASK + (nl >> 6) * 8) >> (nl & 63u)) & 1ull)) return rd<float>(le);
    const uint8_t* leaf = lower + rd<int64_t>(le);
    return rd<float>(leaf + LEAF_VALUES + size_t(((uint32_t(x) & 7u) << 6) | ((uint32_t(y) & 7u) << 3) | (uint32_t(z) & 7u)) * 4);
}

// NanoVDBGrid::lookup(uvec3 ipos) = getValue(Coord(ipos + ibb_min)) (grid_nvdb.cpp:64-67) for every voxel of the padded
// lattice [-2, 8 nb + 2)^3 the brick constructor addresses (grid_brick.cpp:87; the uint32 wrap of negative window
// coordinates and the int32 wrap of `ipos + ibb_min` cancel). One thread per voxel, x fastest: a warp covers 32 voxels
// of one x-row (several 8-voxel leaves) while leaf values run along z, so neighbouring rows re-hit the same 32-byte
// sectors in L1/L2; the tree descent (root tile, two masks, two offsets) is shared by all lanes in a leaf.
__global__ void __launch_bounds__(256) k_nvdb_tabulate(const uint8_t* __restrict__ grid, int3 ibb_min, uint3 pd, float* __restrict__ out) {
    const size_t n = size_t(pd.x) * pd.y * pd.z;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const int32_t x = int32_t(i % pd.x) - 2, y = int32_t((i / pd.x) % pd.y) - 2, z = int32_t(i / (size_t(pd.x) * pd.y)) - 2;
        out[i] = get_value(grid, int32_t(uint32_t(x) + uint32_t(ibb_min.x)), int32_t(uint32_t(y) + uint32_t(ibb_min.y)), int32_t(uint32_t(z) + uint32_t(ibb_min.z)));
    }
}

}  // namespace nvdb
}  // namespace vr
