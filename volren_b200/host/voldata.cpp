// voldata.cpp -- see voldata.h. Heavy lifting (quantisation, brick build) runs on the GPU through the C ABI;
// this file holds the containers, the CPU decode (`lookup`, the definition of a voxel's value), the file formats
// and the Volume frame bookkeeping.
#include "voldata.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

#include "context.h"

namespace fs = std::filesystem;

namespace voldata {

// ------------------------------------------------------------------------------------------------
// number formats (grid_brick.cpp:24-52; glm type_half.inl toFloat32)

static float half_to_float(uint16_t h) {
    const uint32_t s = (h >> 15) & 1u, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    uint32_t bits;
    if (e == 0) {
        if (m == 0) bits = s << 31;
        else {  // subnormal: renormalise
            int ee = -1;
            uint32_t mm = m;
            do { ++ee; mm <<= 1; } while (!(mm & 0x400u));
            bits = (s << 31) | (uint32_t(127 - 15 - ee) << 23) | ((mm & 0x3ffu) << 13);
        }
    } else if (e == 31) bits = (s << 31) | 0x7f800000u | (m << 13);
    else bits = (s << 31) | ((e + 112u) << 23) | (m << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static const uint32_t BRICK_SIZE = 8, VOXELS_PER_BRICK = 512, NUM_MIPMAPS = 3, MAX_BRICKS = 1 << 10;

// ------------------------------------------------------------------------------------------------
// Grid

std::string Grid::to_string(const std::string& indent) const {
    std::stringstream out;
    const uvec3 ibb_max = index_extent();
    out << indent << "AABB (index-space): " << vmath::to_string(uvec3(0)) << " / " << vmath::to_string(ibb_max) << std::endl;
    const auto [mn, mj] = minorant_majorant();
    out << indent << "minorant: " << mn << ", majorant: " << mj << std::endl;
    const size_t active = num_voxels(), dense = size_t(ibb_max.x) * ibb_max.y * ibb_max.z;
    out << indent << "active voxels: " << active / 1000 << "k / " << dense / 1000 << "k (" << uint32_t(std::round(100 * active / float(dense))) << "%)" << std::endl;
    out << indent << "transform: " << std::endl;
    for (int i = 0; i < 4; ++i)
        out << indent << "    " << std::fixed << transform[0][i] << ", " << transform[1][i] << ", " << transform[2][i] << ", " << transform[3][i] << std::endl;
    out << indent << "memory: " << (size_bytes() / 100000) / 10.f << " MB";
    return out.str();
}

// ------------------------------------------------------------------------------------------------
// DenseGrid

DenseGrid::DenseGrid() : Grid(), n_voxels(0), min_value(0), max_value(0) {}

DenseGrid::DenseGrid(const Grid& grid) : Grid(grid), n_voxels(grid.index_extent()), min_value(grid.minorant_majorant().first), max_value(grid.minorant_majorant().second) {
    // grid_dense.cpp:12-32: re-quantise through the virtual lookup (host loop; not on the hot path)
    voxel_data.resize(size_t(n_voxels.x) * n_voxels.y * n_voxels.z);
    for (uint32_t z = 0; z < n_voxels.z; ++z)
        for (uint32_t y = 0; y < n_voxels.y; ++y)
            for (uint32_t x = 0; x < n_voxels.x; ++x) {
                const float value = grid.lookup(uvec3(x, y, z));
                voxel_data[size_t(z) * n_voxels.x * n_voxels.y + size_t(y) * n_voxels.x + x] = uint8_t(std::round(255 * (value - min_value) / (max_value - min_value)));
            }
}

DenseGrid::DenseGrid(const std::shared_ptr<Grid>& grid) : DenseGrid(*grid) {}

DenseGrid::DenseGrid(size_t w, size_t h, size_t d, const uint8_t* data) : Grid(), n_voxels(uint32_t(w), uint32_t(h), uint32_t(d)), min_value(0), max_value(1) {
    voxel_data.assign(data, data + w * h * d);
}

DenseGrid::DenseGrid(size_t w, size_t h, size_t d, const float* data) : Grid(), n_voxels(uint32_t(w), uint32_t(h), uint32_t(d)), min_value(FLT_MAX), max_value(FLT_MIN) {
    voxel_data.resize(w * h * d);
    if (voxel_data.empty()) return;
    const uint32_t dim[3] = { n_voxels.x, n_voxels.y, n_voxels.z };
    float mm[2];
    vrb_ctx* ctx = volren::Context::device();
    volren::check(ctx, vrb_dense_from_float(ctx, data, dim, voxel_data.data(), mm), "DenseGrid(float*)");
    min_value = mm[0];
    max_value = mm[1];
}

float DenseGrid::lookup(const uvec3& ipos) const {
    if (ipos.x >= n_voxels.x || ipos.y >= n_voxels.y || ipos.z >= n_voxels.z) return 0.f;
    const size_t idx = size_t(ipos.z) * n_voxels.x * n_voxels.y + size_t(ipos.y) * n_voxels.x + ipos.x;
    return min_value + (voxel_data[idx] / 255.f) * (max_value - min_value);
}
std::pair<float, float> DenseGrid::minorant_majorant() const { return { min_value, max_value }; }
uvec3 DenseGrid::index_extent() const { return n_voxels; }
size_t DenseGrid::num_voxels() const { return size_t(n_voxels.x) * n_voxels.y * n_voxels.z; }
size_t DenseGrid::size_bytes() const { return size_t(n_voxels.x) * n_voxels.y * n_voxels.z; }

// ------------------------------------------------------------------------------------------------
// BrickGrid

BrickGrid::BrickGrid() : Grid(), n_bricks(0), min_maj({ 0, 0 }), brick_counter(0) {}

static const int SCRATCH_FRAME = 0x7ffffff0;   // device-side staging slot of BrickGrid(const Grid&)

BrickGrid::BrickGrid(const Grid& grid) : Grid(grid), n_bricks(0), min_maj(grid.minorant_majorant()), brick_counter(0) {
    vrb_ctx* ctx = volren::Context::device();
    int st;
    if (const DenseGrid* dense = dynamic_cast<const DenseGrid*>(&grid)) {
        const uint32_t dim[3] = { dense->n_voxels.x, dense->n_voxels.y, dense->n_voxels.z };
        st = vrb_grid_build_from_dense(ctx, VRB_SLOT_DENSITY, SCRATCH_FRAME, dense->voxel_data.data(), dim, dense->min_value, dense->max_value);
    } else if (const NanoVDBGrid* nvdb = dynamic_cast<const NanoVDBGrid*>(&grid)) {
        st = vrb_grid_build_from_nvdb(ctx, VRB_SLOT_DENSITY, SCRATCH_FRAME, nvdb->grid_data(), &nvdb->info);
    } else {
        // any other source: grid.lookup() once per voxel of the padded lattice the constructor addresses (grid_brick.cpp:87)
        const uvec3 e = grid.index_extent();
        const uint32_t extent[3] = { e.x, e.y, e.z };
        uint32_t nb[3], pd[3];
        st = vrb_brick_lattice(extent, nb, pd);
        if (st == VRB_OK) {
            std::vector<float> values(size_t(pd[0]) * pd[1] * pd[2]);
            size_t i = 0;
            for (int z = -2; z < int(pd[2]) - 2; ++z)
                for (int y = -2; y < int(pd[1]) - 2; ++y)
                    for (int x = -2; x < int(pd[0]) - 2; ++x) values[i++] = grid.lookup(uvec3(uint32_t(x), uint32_t(y), uint32_t(z)));
            st = vrb_grid_build_from_values(ctx, VRB_SLOT_DENSITY, SCRATCH_FRAME, values.data(), extent);
        }
    }
    if (st == VRB_ERR_TOO_MANY_BRICKS) throw std::runtime_error(std::string("exceeded max brick count of ") + std::to_string(MAX_BRICKS));
    volren::check(ctx, st, "BrickGrid(const Grid&)");
    vrb_brick_view v;
    memset(&v, 0, sizeof v);
    volren::check(ctx, vrb_grid_info(ctx, VRB_SLOT_DENSITY, SCRATCH_FRAME, &v), "vrb_grid_info");
    n_bricks = uvec3(v.n_bricks[0], v.n_bricks[1], v.n_bricks[2]);
    brick_counter = size_t(v.brick_count);
    indirection.resize(n_bricks);
    range.resize(n_bricks);
    atlas.resize(uvec3(v.atlas_dim[0], v.atlas_dim[1], v.atlas_dim[2]));
    range_mipmaps.resize(NUM_MIPMAPS);
    for (uint32_t i = 0; i < NUM_MIPMAPS; ++i) range_mipmaps[i].resize(uvec3(n_bricks.x >> (i + 1), n_bricks.y >> (i + 1), n_bricks.z >> (i + 1)));
    v.indirection = indirection.data.data();
    v.range = range.data.data();
    v.atlas = atlas.data.empty() ? nullptr : atlas.data.data();
    for (uint32_t i = 0; i < NUM_MIPMAPS; ++i) v.range_mips[i] = range_mipmaps[i].data.data();
    st = vrb_grid_download(ctx, VRB_SLOT_DENSITY, SCRATCH_FRAME, &v);
    vrb_grid_free(ctx, VRB_SLOT_DENSITY, SCRATCH_FRAME);
    volren::check(ctx, st, "vrb_grid_download");
}

BrickGrid::BrickGrid(const std::shared_ptr<Grid>& grid) : BrickGrid(*grid) {}

float BrickGrid::lookup(const uvec3& ipos) const {   // grid_brick.cpp:148-154
    const uvec3 brick(ipos.x >> 3, ipos.y >> 3, ipos.z >> 3);
    const uint32_t p = indirection[brick], r = range[brick];
    const uvec3 ptr((p >> 22) & 1023u, (p >> 12) & 1023u, (p >> 2) & 1023u);
    const float lo = half_to_float(uint16_t(r & 0xffffu)), hi = half_to_float(uint16_t(r >> 16));
    const uvec3 voxel((ptr.x << 3) + (ipos.x & 7u), (ptr.y << 3) + (ipos.y & 7u), (ptr.z << 3) + (ipos.z & 7u));
    return lo + atlas[voxel] * (1.f / 255.f) * (hi - lo);
}
std::pair<float, float> BrickGrid::minorant_majorant() const { return min_maj; }
uvec3 BrickGrid::index_extent() const { return n_bricks * BRICK_SIZE; }
size_t BrickGrid::num_voxels() const { return brick_counter * VOXELS_PER_BRICK; }
size_t BrickGrid::size_bytes() const {
    const size_t dense_bricks = size_t(n_bricks.x) * n_bricks.y * n_bricks.z;
    size_t size_mipmaps = 0;
    for (const auto& mip : range_mipmaps) size_mipmaps += sizeof(uint32_t) * mip.data.size();
    return 2 * sizeof(uint32_t) * dense_bricks + brick_counter * VOXELS_PER_BRICK + size_mipmaps;
}
std::string BrickGrid::to_string(const std::string& indent) const {
    std::stringstream out;
    out << Grid::to_string(indent) << std::endl;
    out << indent << "voxel dim: " << vmath::to_string(index_extent()) << std::endl;
    out << indent << "brick dim: " << vmath::to_string(n_bricks) << std::endl;
    const size_t allocd = brick_counter, capacity = atlas.data.size() / VOXELS_PER_BRICK;
    out << indent << "bricks in atlas: " << allocd << " / " << capacity << " (" << uint32_t(std::round(100 * allocd / float(capacity))) << "%)" << std::endl;
    out << indent << "atlas dim: " << vmath::to_string(atlas.size()) << std::endl;
    return out.str();
}

// ------------------------------------------------------------------------------------------------
// NanoVDBGrid (grid_nvdb.cpp:8-28, :64-88): the file image is the handle; location, validation and the derived
// members come from vrb_nvdb_open, lookup() from the host accessor of the C ABI

NanoVDBGrid::NanoVDBGrid(const std::string& path, const std::string& gridname) : Grid() {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in.is_open()) throw std::ios_base::failure("Unable to open file named \"" + path + "\" for input");
    file.resize(size_t(in.tellg()));
    in.seekg(0);
    in.read(reinterpret_cast<char*>(file.data()), std::streamsize(file.size()));
    char err[512];
    if (vrb_nvdb_open(file.data(), file.size(), gridname.c_str(), &info, err, sizeof err) != VRB_OK) throw std::runtime_error(err);
    ibb_min = ivec3(info.ibb_min[0], info.ibb_min[1], info.ibb_min[2]);
    extent = uvec3(info.extent[0], info.extent[1], info.extent[2]);
    minorant = info.minorant;
    majorant = info.majorant;
    memcpy(&transform[0][0], info.transform, sizeof info.transform);
}
float NanoVDBGrid::lookup(const uvec3& ipos) const {
    float v = 0.f;
    vrb_nvdb_lookup(grid_data(), info.ibb_min, &ipos.x, 1, &v);
    return v;
}
std::pair<float, float> NanoVDBGrid::minorant_majorant() const { return { minorant, majorant }; }
uvec3 NanoVDBGrid::index_extent() const { return extent; }
size_t NanoVDBGrid::num_voxels() const { return size_t(info.active_voxels); }
size_t NanoVDBGrid::size_bytes() const { return size_t(info.grid_size); }

// ------------------------------------------------------------------------------------------------
// cereal PortableBinary (little-endian) archives, field order of serialization.cpp:36-43

namespace {

struct Reader {
    std::ifstream in;
    std::string path;
    explicit Reader(const std::string& p) : in(p, std::ios::binary), path(p) {
        if (!in.is_open()) throw std::runtime_error("Unable to read file: " + p);
        uint8_t little = 1;
        raw(&little, 1);
        if (little != 1) throw std::runtime_error("big-endian archive not supported: " + p);
    }
    void raw(void* dst, size_t n) {
        in.read(static_cast<char*>(dst), std::streamsize(n));
        if (size_t(in.gcount()) != n) throw std::runtime_error("unexpected end of file: " + path);
    }
    template <typename T> T pod() { T v; raw(&v, sizeof v); return v; }
    mat4 matrix() { mat4 m; raw(&m[0].x, 64); return m; }
    uvec3 uv3() { uvec3 v; raw(&v.x, 12); return v; }
    template <typename T> void vec(std::vector<T>& out) {
        const uint64_t n = pod<uint64_t>();
        out.resize(n);
        if (n) raw(out.data(), n * sizeof(T));
    }
    template <typename T> void buf3d(Buf3D<T>& b) {
        b.stride = uv3();
        vec(b.data);
        if (b.data.size() != size_t(b.stride.x) * b.stride.y * b.stride.z) throw std::runtime_error("corrupt Buf3D in " + path);
    }
};

struct Writer {
    std::ofstream out;
    explicit Writer(const std::string& p) : out(p, std::ios::binary) {
        if (!out.is_open()) throw std::runtime_error("Unable to write file: " + p);
        const uint8_t little = 1;
        raw(&little, 1);
    }
    void raw(const void* src, size_t n) { out.write(static_cast<const char*>(src), std::streamsize(n)); }
    template <typename T> void pod(const T& v) { raw(&v, sizeof v); }
    template <typename T> void vec(const std::vector<T>& v) {
        pod<uint64_t>(v.size());
        if (!v.empty()) raw(v.data(), v.size() * sizeof(T));
    }
    template <typename T> void buf3d(const Buf3D<T>& b) { raw(&b.stride.x, 12); vec(b.data); }
};

}  // namespace

std::shared_ptr<DenseGrid> load_dense_grid(const std::string& path) {
    Reader r(path);
    auto g = std::make_shared<DenseGrid>();
    g->transform = r.matrix();
    g->n_voxels = r.uv3();
    g->min_value = r.pod<float>();
    g->max_value = r.pod<float>();
    r.vec(g->voxel_data);
    if (g->voxel_data.size() != g->num_voxels()) throw std::runtime_error("corrupt dense grid: " + path);
    return g;
}

std::shared_ptr<BrickGrid> load_brick_grid(const std::string& path) {
    Reader r(path);
    auto g = std::make_shared<BrickGrid>();
    g->transform = r.matrix();
    g->n_bricks = r.uv3();
    g->min_maj.first = r.pod<float>();
    g->min_maj.second = r.pod<float>();
    g->brick_counter = size_t(r.pod<uint64_t>());
    r.buf3d(g->indirection);
    r.buf3d(g->range);
    r.buf3d(g->atlas);
    const uint64_t n_mips = r.pod<uint64_t>();
    if (n_mips > 16) throw std::runtime_error("corrupt brick grid: " + path);
    g->range_mipmaps.resize(n_mips);
    for (auto& m : g->range_mipmaps) r.buf3d(m);
    const std::string why = g->check_layout();
    if (!why.empty()) throw std::runtime_error("corrupt brick grid (" + why + "): " + path);
    return g;
}

// The invariants of a BrickGrid the constructor establishes (grid_brick.cpp:60-141) and the uploads rely on: a file or a
// hand-assembled grid that breaks them would make the upload read past its host buffers.
std::string BrickGrid::check_layout() const {
    for (int a = 0; a < 3; ++a)
        if (n_bricks[a] == 0 || n_bricks[a] >= 1024u || (n_bricks[a] & 7u)) return "n_bricks must be multiples of 8 in [8, 1016]";
    const size_t n = size_t(n_bricks.x) * n_bricks.y * n_bricks.z;
    if (!(indirection.stride == n_bricks) || indirection.data.size() != n) return "indirection does not match n_bricks";
    if (!(range.stride == n_bricks) || range.data.size() != n) return "range does not match n_bricks";
    if ((atlas.stride.x & 7u) || (atlas.stride.y & 7u) || (atlas.stride.z & 7u) ||
        atlas.data.size() != size_t(atlas.stride.x) * atlas.stride.y * atlas.stride.z) return "atlas stride / size mismatch";
    if (range_mipmaps.size() != 3) return "expected 3 range mipmaps";
    for (uint32_t i = 0; i < 3; ++i) {
        const uvec3 d(n_bricks.x >> (i + 1), n_bricks.y >> (i + 1), n_bricks.z >> (i + 1));
        if (!(range_mipmaps[i].stride == d) || range_mipmaps[i].data.size() != size_t(d.x) * d.y * d.z) return "range mipmap " + std::to_string(i) + " does not match n_bricks";
    }
    return std::string();
}

void write_grid(const std::shared_ptr<Grid>& grid, const std::string& path) {
    if (auto* d = dynamic_cast<DenseGrid*>(grid.get())) {
        Writer w(path);
        w.raw(&d->transform[0].x, 64);
        w.raw(&d->n_voxels.x, 12);
        w.pod(d->min_value);
        w.pod(d->max_value);
        w.vec(d->voxel_data);
    } else if (auto* b = dynamic_cast<BrickGrid*>(grid.get())) {
        Writer w(path);
        w.raw(&b->transform[0].x, 64);
        w.raw(&b->n_bricks.x, 12);
        w.pod(b->min_maj.first);
        w.pod(b->min_maj.second);
        w.pod<uint64_t>(b->brick_counter);
        w.buf3d(b->indirection);
        w.buf3d(b->range);
        w.buf3d(b->atlas);
        w.pod<uint64_t>(b->range_mipmaps.size());
        for (const auto& m : b->range_mipmaps) w.buf3d(m);
    } else
        throw std::runtime_error("Unsupported grid type!");
    std::cout << fs::path(path) << " written." << std::endl;
}

// ------------------------------------------------------------------------------------------------
// Volume

Volume::Volume() : grid_frame_counter(0), transform(1.f) {}
Volume::Volume(const GridPtr& grid, const std::string& gridname) : Volume() {
    GridFrame frame;
    frame[gridname] = grid;
    add_grid_frame(frame);
}
Volume::Volume(const std::string& filename, const std::string& gridname) : Volume() {
    GridFrame frame;
    frame[gridname] = load_grid(filename, gridname);
    add_grid_frame(frame);
}
Volume::Volume(size_t w, size_t h, size_t d, const uint8_t* data, const std::string& gridname) : Volume() {
    GridFrame frame;
    frame[gridname] = std::make_shared<DenseGrid>(w, h, d, data);
    add_grid_frame(frame);
}
Volume::Volume(size_t w, size_t h, size_t d, const float* data, const std::string& gridname) : Volume() {
    GridFrame frame;
    frame[gridname] = std::make_shared<DenseGrid>(w, h, d, data);
    add_grid_frame(frame);
}

void Volume::clear() { grids.clear(); }
void Volume::add_grid_frame(const GridFrame& frame) { grids.push_back(frame); }
void Volume::update_grid_frame(const size_t i, const GridPtr& grid, const std::string& gridname) { grids.at(i)[gridname] = grid; }
bool Volume::has_grid(const size_t i, const std::string& gridname) const { return grids.at(i).find(gridname) != grids.at(i).end(); }
size_t Volume::n_grid_frames() const { return grids.size(); }
Volume::GridFrame Volume::current_grid_frame() const { return grids.at(grid_frame_counter); }
Volume::GridPtr Volume::current_grid(const std::string& gridname) const { return grids.at(grid_frame_counter).at(gridname); }
Volume::DenseGridPtr Volume::current_grid_dense(const std::string& gridname) const { return to_dense_grid(current_grid(gridname)); }
Volume::BrickGridPtr Volume::current_grid_brick(const std::string& gridname) const { return to_brick_grid(current_grid(gridname)); }

mat4 Volume::get_transform(const std::string& gridname) const {
    if (grids.size() <= grid_frame_counter) return transform;
    return transform * current_grid(gridname)->transform;
}
vec4 Volume::to_world(const vec4& index, const std::string& gridname) const { return get_transform(gridname) * index; }
vec4 Volume::to_index(const vec4& world, const std::string& gridname) const { return inverse(get_transform(gridname)) * world; }

std::pair<vec3, vec3> Volume::AABB(const std::string& gridname) const {
    if (grids.size() <= grid_frame_counter) return { vec3(0), vec3(0) };
    // volume.cpp:102-107: the extent is always the *density* grid's (current_grid() with the default name)
    const vec3 wbb_min = vec3(to_world(vec4(0, 0, 0, 1), gridname));
    const vec3 wbb_max = vec3(to_world(vec4(vec3(current_grid()->index_extent()), 1), gridname));
    return { wbb_min, wbb_max };
}

std::pair<float, float> Volume::minorant_majorant(const std::string& gridname) const {
    if (grids.size() <= grid_frame_counter) return { 0.f, 0.f };
    return current_grid(gridname)->minorant_majorant();
}

std::string Volume::to_string(const std::string& indent) const {
    std::stringstream out;
    const auto [bb_min, bb_max] = AABB();
    out << indent << "AABB: " << vmath::to_string(bb_min) << " / " << vmath::to_string(bb_max) << std::endl;
    out << indent << "modelmatrix: " << std::endl;
    for (int i = 0; i < 4; ++i)
        out << indent << "    " << std::fixed << transform[0][i] << ", " << transform[1][i] << ", " << transform[2][i] << ", " << transform[3][i] << std::endl;
    out << indent << "current grid frame: " << grid_frame_counter << " / " << grids.size() << std::endl;
    out << indent << "current grid: " << std::endl;
    if (grids.size() > grid_frame_counter) out << current_grid()->to_string(indent + "    ") << std::endl;
    return out.str();
}

static std::string lower_ext(const fs::path& p) {
    std::string e = p.extension().string();
    std::transform(e.begin(), e.end(), e.begin(), ::tolower);
    return e;
}

// .dat descriptor + raw file (volume.cpp:132-192)
static Volume::GridPtr load_dat(const fs::path& path) {
    std::ifstream dat_file(path);
    if (!dat_file.is_open()) throw std::runtime_error("Unable to read file: " + path.string());
    std::string raw_name, format, key;
    ivec3 dim(0);
    vec3 slice_thickness(1.f);
    int bits = 0;
    while (dat_file >> key) {
        if (key == "ObjectFileName:") dat_file >> raw_name;
        else if (key == "Resolution:") dat_file >> dim.x >> dim.y >> dim.z;
        else if (key == "SliceThickness:") dat_file >> slice_thickness.x >> slice_thickness.y >> slice_thickness.z;
        else if (key == "Format:") dat_file >> format;
        else if (key == "BitsUsed:") dat_file >> bits;
        else std::cout << "Skipping key: " << key << "..." << std::endl;
    }
    (void)bits;
    if (raw_name.size() >= 2 && raw_name.front() == '"' && raw_name.back() == '"') raw_name = raw_name.substr(1, raw_name.size() - 2);
    const fs::path raw_path = path.parent_path() / raw_name;
    std::ifstream raw_file(raw_path, std::ios::binary);
    if (!raw_file.is_open()) throw std::runtime_error("Unable to read file: " + raw_path.string());
    const std::vector<uint8_t> data((std::istreambuf_iterator<char>(raw_file)), std::istreambuf_iterator<char>());
    if (dim.x <= 0 || dim.y <= 0 || dim.z <= 0) throw std::runtime_error("Bad resolution in .dat file: " + path.string());
    const size_t n = size_t(dim.x) * dim.y * dim.z;
    std::cout << "data size bytes: " << data.size() << " / " << n << std::endl;
    Volume::GridPtr grid;
    if (format == "UCHAR") {
        if (data.size() < n) throw std::runtime_error("raw file too small: " + raw_path.string());
        grid = std::make_shared<DenseGrid>(dim.x, dim.y, dim.z, data.data());
    } else if (format == "USHORT") {
        // the reference's conversion appends round(v / 65535) behind n zeros, i.e. the grid it builds is all zero
        // (volume.cpp:177-182); normalised 16-bit data is what was meant and what is loaded here
        if (data.size() < 2 * n) throw std::runtime_error("raw file too small: " + raw_path.string());
        std::vector<float> f(n);
        const uint16_t* p16 = reinterpret_cast<const uint16_t*>(data.data());
        for (size_t i = 0; i < n; ++i) f[i] = p16[i] / 65535.f;
        grid = std::make_shared<DenseGrid>(dim.x, dim.y, dim.z, f.data());
    } else if (format == "FLOAT") {
        if (data.size() < 4 * n) throw std::runtime_error("raw file too small: " + raw_path.string());
        grid = std::make_shared<DenseGrid>(dim.x, dim.y, dim.z, reinterpret_cast<const float*>(data.data()));
    } else
        throw std::runtime_error("Unsupported data format for .dat file: " + format);
    // scale and map from z up to y up
    grid->transform = scale(rotate(mat4(1.f), float(1.5 * M_PI), vec3(1, 0, 0)), slice_thickness);
    return grid;
}

Volume::GridPtr Volume::load_grid(const std::string& filename, const std::string& gridname) {
    const fs::path path = filename;
    const std::string extension = lower_ext(path);
    if (extension == ".dat") return load_dat(path);
    if (extension == ".dense") return load_dense_grid(path.string());
    if (extension == ".brick") return load_brick_grid(path.string());
    if (extension == ".nvdb") return std::make_shared<NanoVDBGrid>(path.string(), gridname);   // volume.cpp:198-200
    if (extension == ".vdb" || extension == ".dcm")
        throw std::runtime_error("Unable to load file extension: " + extension + " (OpenVDB/DICOM adapters are not part of the B200 build; convert to .nvdb/.brick/.dense)");
    throw std::runtime_error("Unable to load file extension: " + extension);
}

Volume::DenseGridPtr Volume::to_dense_grid(const GridPtr& grid) {
    auto dense = std::dynamic_pointer_cast<DenseGrid>(grid);
    if (!dense) dense = std::make_shared<DenseGrid>(grid);
    return dense;
}

Volume::BrickGridPtr Volume::to_brick_grid(const GridPtr& grid) {
    auto brick = std::dynamic_pointer_cast<BrickGrid>(grid);
    if (!brick) brick = std::make_shared<BrickGrid>(grid);
    return brick;
}

Volume::VolumePtr Volume::load_folder(const std::string& path, std::vector<std::string> gridnames) {
    VolumePtr result = std::make_shared<Volume>();
    std::cout << "Loading grid files from " << path << "..." << std::endl;
    std::vector<fs::path> files;
    for (auto& p : fs::directory_iterator(fs::path(path))) files.push_back(p);
    // shorter names first, then lexicographic (frame_2 before frame_10)
    std::sort(files.begin(), files.end(), [](const fs::path& lhs, const fs::path& rhs) {
        return lhs.string().size() == rhs.string().size() ? lhs.string() < rhs.string() : lhs.string().size() < rhs.string().size();
    });
    result->grids.resize(files.size());
    // volume.cpp:283-291 calls load_grid(file, name) for EVERY requested name; the single-grid formats handled here
    // ignore the name, so one file fills every requested slot of its frame (a .brick sequence loaded by the CLI with
    // { density, temperature, flame, flames } therefore also acts as its own emission grid). Same observable result,
    // one read per file.
    // A .nvdb container selects by name: one load per requested name, absent names are skipped (volume.cpp:286-290).
    for (size_t i = 0; i < files.size(); ++i) {
        if (lower_ext(files[i]) == ".nvdb") {
            for (const auto& gridname : gridnames) {
                try {
                    result->update_grid_frame(i, load_grid(files[i].string(), gridname), gridname);
                } catch (std::runtime_error&) {}
            }
            continue;
        }
        try {
            const GridPtr grid = load_grid(files[i].string(), gridnames.empty() ? "density" : gridnames[0]);
            for (const auto& gridname : gridnames) result->update_grid_frame(i, grid, gridname);
        } catch (std::runtime_error&) {}
    }
    return result;
}

}  // namespace voldata
