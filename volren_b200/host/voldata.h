// voldata.h -- host-side volume data model of the VolRen host: the part of the reference's `voldata` library that
// the renderer, the CLI and the Python module expose (Grid / DenseGrid / BrickGrid / Volume, .brick/.dense/.dat
// loaders), rebuilt around the B200 back end:
//   * DenseGrid(w,h,d,const float*)  -> min/max reduction + 8-bit quantisation on the GPU (vrb_dense_from_float),
//     bit-exact w.r.t. voldata/src/grid_dense.cpp:57-95
//   * BrickGrid(const Grid&)         -> brick build on the GPU, bit-exact w.r.t. the serial voldata/src/grid_brick.cpp:60-142:
//     vrb_grid_build_from_dense for a DenseGrid, vrb_grid_build_from_nvdb for a NanoVDBGrid (device accessor), and
//     vrb_grid_build_from_values (lookup() tabulated on the padded lattice) for any other Grid
//   * NanoVDBGrid(path, gridname)    -> .nvdb reader + accessor without the NanoVDB headers (vrb_nvdb_open / _lookup),
//     voldata/src/grid_nvdb.cpp:8-28,64-67
//   * file formats                   -> cereal PortableBinary layout of voldata/src/serialization.cpp:16-43,66-80
//     re-implemented without cereal
// Interface names and semantics follow voldata/src/{grid,grid_dense,grid_brick,volume,buf3d}.h.
#pragma once

#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/vrb200.h"
#include "vmath.h"

namespace voldata {

using namespace vmath;

// x-fastest 3-D buffer (voldata/src/buf3d.h)
template <typename T> class Buf3D {
public:
    Buf3D(const uvec3& stride = uvec3(0)) : stride(stride), data(size_t(stride.x) * stride.y * stride.z) {}
    T& operator[](const uvec3& at) { return data[to_idx(at)]; }
    const T& operator[](const uvec3& at) const { return data[to_idx(at)]; }
    uvec3 size() const { return stride; }
    void resize(const uvec3& s) { stride = s; data.resize(size_t(s.x) * s.y * s.z); }
    size_t to_idx(const uvec3& c) const { return size_t(c.z) * stride.x * stride.y + size_t(c.y) * stride.x + c.x; }
    uvec3 to_coord(size_t idx) const { return uvec3(uint32_t(idx % stride.x), uint32_t((idx / stride.x) % stride.y), uint32_t(idx / (size_t(stride.x) * stride.y))); }
    uvec3 stride;
    std::vector<T> data;
};

// voldata/src/grid.h
class Grid {
public:
    Grid() : transform(1.f) {}
    virtual ~Grid() {}
    virtual float lookup(const uvec3& ipos) const = 0;
    virtual std::pair<float, float> minorant_majorant() const = 0;
    virtual uvec3 index_extent() const = 0;
    virtual size_t num_voxels() const = 0;
    virtual size_t size_bytes() const = 0;
    virtual std::string to_string(const std::string& indent = "") const;
    float operator[](const uvec3& ipos) const { return lookup(ipos); }
    mat4 transform;
};

// voldata/src/grid_dense.h
class DenseGrid : public Grid {
public:
    DenseGrid();
    DenseGrid(const Grid& grid);                                   // re-quantise any grid (host loop over lookup())
    DenseGrid(const std::shared_ptr<Grid>& grid);
    DenseGrid(size_t w, size_t h, size_t d, const uint8_t* data);  // range fixed to [0, 1]
    DenseGrid(size_t w, size_t h, size_t d, const float* data);    // GPU: global min/max + quantise
    float lookup(const uvec3& ipos) const override;
    std::pair<float, float> minorant_majorant() const override;
    uvec3 index_extent() const override;
    size_t num_voxels() const override;
    size_t size_bytes() const override;
    uvec3 n_voxels;
    float min_value, max_value;
    std::vector<uint8_t> voxel_data;
};

// voldata/src/grid_brick.h
class BrickGrid : public Grid {
public:
    BrickGrid();
    BrickGrid(const Grid& grid);                                   // GPU brick build (DenseGrid source; others via DenseGrid)
    BrickGrid(const std::shared_ptr<Grid>& grid);
    float lookup(const uvec3& ipos) const override;
    std::pair<float, float> minorant_majorant() const override;
    uvec3 index_extent() const override;
    size_t num_voxels() const override;
    size_t size_bytes() const override;
    std::string to_string(const std::string& indent = "") const override;
    std::string check_layout() const;            // "" when the buffers have the shapes the constructor produces, else what is wrong
    uvec3 n_bricks;
    std::pair<float, float> min_maj;
    size_t brick_counter;
    Buf3D<uint32_t> indirection;                 // 10/10/10/2-bit atlas pointer
    Buf3D<uint32_t> range;                       // 2 x fp16 (minorant, majorant)
    Buf3D<uint8_t> atlas;                        // 8^3 unorm8 voxels per allocated brick
    std::vector<Buf3D<uint32_t>> range_mipmaps;  // 3 levels of 2x2x2 min/max
};

// voldata/src/grid_nvdb.h (file constructor; converting other grids TO NanoVDB is not part of this build)
class NanoVDBGrid : public Grid {
public:
    NanoVDBGrid(const std::string& path, const std::string& gridname = "density");
    float lookup(const uvec3& ipos) const override;
    std::pair<float, float> minorant_majorant() const override;
    uvec3 index_extent() const override;
    size_t num_voxels() const override;
    size_t size_bytes() const override;
    const uint8_t* grid_data() const { return file.data() + info.grid_offset; }   // the serialized grid buffer
    std::vector<uint8_t> file;                   // file image (handle)
    vrb_nvdb_info info;
    ivec3 ibb_min;
    uvec3 extent;
    float minorant, majorant;
};

// voldata/src/serialization.h
void write_grid(const std::shared_ptr<Grid>& grid, const std::string& path);
std::shared_ptr<DenseGrid> load_dense_grid(const std::string& path);
std::shared_ptr<BrickGrid> load_brick_grid(const std::string& path);

// voldata/src/volume.h
class Volume {
public:
    using GridPtr = std::shared_ptr<Grid>;
    using DenseGridPtr = std::shared_ptr<DenseGrid>;
    using BrickGridPtr = std::shared_ptr<BrickGrid>;
    using GridFrame = std::map<std::string, GridPtr>;
    using VolumePtr = std::shared_ptr<Volume>;

    Volume();
    Volume(const GridPtr& grid, const std::string& gridname = "density");
    Volume(const std::string& filename, const std::string& gridname = "density");
    Volume(size_t w, size_t h, size_t d, const uint8_t* data, const std::string& gridname = "density");
    Volume(size_t w, size_t h, size_t d, const float* data, const std::string& gridname = "density");
    virtual ~Volume() {}

    void clear();
    void add_grid_frame(const GridFrame& frame = GridFrame());
    void update_grid_frame(const size_t i, const GridPtr& grid, const std::string& gridname = "density");
    bool has_grid(const size_t i, const std::string& gridname) const;
    size_t n_grid_frames() const;

    GridFrame current_grid_frame() const;
    GridPtr current_grid(const std::string& gridname = "density") const;
    DenseGridPtr current_grid_dense(const std::string& gridname = "density") const;
    BrickGridPtr current_grid_brick(const std::string& gridname = "density") const;

    mat4 get_transform(const std::string& gridname = "density") const;
    vec4 to_world(const vec4& index, const std::string& gridname = "density") const;
    vec4 to_index(const vec4& world, const std::string& gridname = "density") const;
    std::pair<vec3, vec3> AABB(const std::string& gridname = "density") const;
    std::pair<float, float> minorant_majorant(const std::string& gridname = "density") const;
    std::string to_string(const std::string& indent = "") const;

    static GridPtr load_grid(const std::string& filename, const std::string& gridname = "density");
    static DenseGridPtr to_dense_grid(const GridPtr& grid);
    static BrickGridPtr to_brick_grid(const GridPtr& grid);
    static VolumePtr load_folder(const std::string& path, std::vector<std::string> gridnames = { "density" });

    size_t grid_frame_counter;
    std::vector<GridFrame> grids;
    mat4 transform;
};

}  // namespace voldata
