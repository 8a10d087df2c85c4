// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// C-callable harness around the *unmodified* reference voldata sources
// (/root/reference/submodules/voldata/src/{grid,grid_dense,grid_brick}.cpp), which are compiled
// where they lie by oracle/Makefile into oracle/_ref/libvoldata_ref.so. It is used to
//   (1) pin oracle/vr_oracle.c (the CPU restatement) bit-exactly, and
//   (2) generate the golden vectors under tests/golden/ (tests/golden/make_golden.py).
// It cannot travel to the GPU box as source (it needs /root/reference at compile time); the
// built .so does travel, but the -m gpu tests only rely on vr_oracle.c + committed goldens.
//
// The cereal archive layout is the one declared at voldata/src/serialization.cpp:16-43 (that file
// itself cannot be compiled here because it drags in OpenVDB/NanoVDB headers), so the field order
// is re-declared below against the reference's own cereal headers.

#include <cstdint>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>

#include "grid_dense.h"
#include "grid_brick.h"
#include "grid_nvdb.h"          // the reference's NanoVDB adapter (voldata/src/grid_nvdb.cpp), header-only NanoVDB 32.7
#include <nanovdb/io/IO.h>
#include <nanovdb/tools/GridBuilder.h>
#include <nanovdb/tools/CreateNanoGrid.h>
#include <nanovdb/tools/CreatePrimitives.h>

#include <cereal/types/utility.hpp>
#include <cereal/types/atomic.hpp>
#include <cereal/types/vector.hpp>
#include <cereal/archives/portable_binary.hpp>
#include <glm/detail/type_half.hpp>
#include <tinycolormap.hpp>    // the reference's colour-map submodule (src/transferfunc.cpp:69-77 samples it)

namespace cereal {
    template <class Archive> void serialize(Archive& ar, glm::uvec3& v) { ar(v.x, v.y, v.z); }
    template <class Archive> void serialize(Archive& ar, glm::vec4& v) { ar(v.x, v.y, v.z, v.w); }
    template <class Archive> void serialize(Archive& ar, glm::mat4& m) { ar(m[0], m[1], m[2], m[3]); }
}
namespace voldata {
    template <class Archive, typename T> void serialize(Archive& ar, Buf3D<T>& buf) { ar(buf.stride, buf.data); }
    template <class Archive> void serialize(Archive& ar, DenseGrid& g) {
        ar(g.transform, g.n_voxels, g.min_value, g.max_value, g.voxel_data);
    }
    template <class Archive> void serialize(Archive& ar, BrickGrid& g) {
        ar(g.transform, g.n_bricks, g.min_maj, g.brick_counter, g.indirection, g.range, g.atlas, g.range_mipmaps);
    }
}

using namespace voldata;

extern "C" {

// ---- TransferFunction::colormap (src/transferfunc.cpp:69-77): the loop restated around the REAL tinycolormap::GetColor ----
// (transferfunc.cpp itself needs the GL SSBO wrapper); out: n_bins x RGBA float, before upload_gpu
void ref_colormap_lut(int type, uint32_t n_bins, float* out) {
    for (uint32_t i = 0; i < n_bins; ++i) {
        const float f = float(i) / n_bins;
        const tinycolormap::Color color = tinycolormap::GetColor(f, tinycolormap::ColormapType(type));
        const glm::vec4 v(color.r(), color.g(), color.b(), f);
        out[4 * i] = v.x; out[4 * i + 1] = v.y; out[4 * i + 2] = v.z; out[4 * i + 3] = v.w;
    }
}

// ---- glm half conversion (glm/detail/type_half.inl) ----
uint16_t ref_to_half(float f) { return uint16_t(glm::detail::toFloat16(f)); }
float ref_from_half(uint16_t h) { return glm::detail::toFloat32(glm::detail::hdata(h)); }

// ---- DenseGrid ----
void* ref_dense_from_float(uint32_t w, uint32_t h, uint32_t d, const float* data) { return new DenseGrid(w, h, d, data); }
void* ref_dense_from_u8(uint32_t w, uint32_t h, uint32_t d, const uint8_t* data) { return new DenseGrid(w, h, d, data); }
void ref_dense_set_range(void* g, float lo, float hi) { ((DenseGrid*)g)->min_value = lo; ((DenseGrid*)g)->max_value = hi; }
void ref_dense_info(void* g, uint32_t dim[3], float minmax[2]) {
    auto* gr = (DenseGrid*)g;
    dim[0] = gr->n_voxels.x; dim[1] = gr->n_voxels.y; dim[2] = gr->n_voxels.z;
    minmax[0] = gr->min_value; minmax[1] = gr->max_value;
}
void ref_dense_voxels(void* g, uint8_t* out) { auto* gr = (DenseGrid*)g; memcpy(out, gr->voxel_data.data(), gr->voxel_data.size()); }
float ref_dense_lookup(void* g, uint32_t x, uint32_t y, uint32_t z) { return ((DenseGrid*)g)->lookup(glm::uvec3(x, y, z)); }
void ref_dense_free(void* g) { delete (DenseGrid*)g; }
void* ref_dense_load(const char* path) {
    std::ifstream file(path, std::ios::binary);
    if (!file.is_open()) return nullptr;
    cereal::PortableBinaryInputArchive archive(file);
    auto* g = new DenseGrid();
    archive(*g);
    return g;
}
int ref_dense_write(void* g, const char* path) {
    std::ofstream file(path, std::ios::binary);
    if (!file.is_open()) return -1;
    cereal::PortableBinaryOutputArchive archive(file);
    archive(*(DenseGrid*)g);
    return 0;
}

// ---- BrickGrid ----
void* ref_brick_from_dense(void* dense) {
    try { return new BrickGrid(*(DenseGrid*)dense); } catch (std::runtime_error&) { return nullptr; }
}
void* ref_brick_load(const char* path) {
    std::ifstream file(path, std::ios::binary);
    if (!file.is_open()) return nullptr;
    cereal::PortableBinaryInputArchive archive(file);
    auto* g = new BrickGrid();
    archive(*g);
    return g;
}
int ref_brick_write(void* g, const char* path) {
    std::ofstream file(path, std::ios::binary);
    if (!file.is_open()) return -1;
    cereal::PortableBinaryOutputArchive archive(file);
    archive(*(BrickGrid*)g);
    return 0;
}
void ref_brick_info(void* g, uint32_t n_bricks[3], uint32_t atlas_dim[3], float min_maj[2], uint64_t* counter,
                    float transform[16], uint32_t* n_mips) {
    auto* b = (BrickGrid*)g;
    n_bricks[0] = b->n_bricks.x; n_bricks[1] = b->n_bricks.y; n_bricks[2] = b->n_bricks.z;
    atlas_dim[0] = b->atlas.stride.x; atlas_dim[1] = b->atlas.stride.y; atlas_dim[2] = b->atlas.stride.z;
    min_maj[0] = b->min_maj.first; min_maj[1] = b->min_maj.second;
    *counter = b->brick_counter;
    memcpy(transform, &b->transform[0][0], 64);
    *n_mips = uint32_t(b->range_mipmaps.size());
}
void ref_brick_copy(void* g, uint32_t* indirection, uint32_t* range, uint8_t* atlas, uint32_t* mip0, uint32_t* mip1, uint32_t* mip2) {
    auto* b = (BrickGrid*)g;
    if (indirection) memcpy(indirection, b->indirection.data.data(), b->indirection.data.size() * 4);
    if (range) memcpy(range, b->range.data.data(), b->range.data.size() * 4);
    if (atlas) memcpy(atlas, b->atlas.data.data(), b->atlas.data.size());
    uint32_t* mips[3] = { mip0, mip1, mip2 };
    for (size_t i = 0; i < 3 && i < b->range_mipmaps.size(); ++i)
        if (mips[i]) memcpy(mips[i], b->range_mipmaps[i].data.data(), b->range_mipmaps[i].data.size() * 4);
}
float ref_brick_lookup(void* g, uint32_t x, uint32_t y, uint32_t z) { return ((BrickGrid*)g)->lookup(glm::uvec3(x, y, z)); }
// decode every voxel of the padded index extent through BrickGrid::lookup (grid_brick.cpp:148-154)
void ref_brick_decode_all(void* g, float* out) {
    auto* b = (BrickGrid*)g;
    const glm::uvec3 e = b->index_extent();
    size_t i = 0;
    for (uint32_t z = 0; z < e.z; ++z) for (uint32_t y = 0; y < e.y; ++y) for (uint32_t x = 0; x < e.x; ++x)
        out[i++] = b->lookup(glm::uvec3(x, y, z));
}
void ref_brick_free(void* g) { delete (BrickGrid*)g; }

// ---- NanoVDB (voldata/src/grid_nvdb.cpp; fixtures are written with the reference's own NanoVDB headers) ----
// Sparse float fog volume from (coord, value) pairs; returns a heap GridHandle.
void* ref_nvdb_build(const int32_t* ijk, const float* values, size_t n, const char* name, float background, double voxel_size, const double origin[3]) {
    nanovdb::tools::build::Grid<float> builder(background, name, nanovdb::GridClass::FogVolume);
    builder.setTransform(voxel_size, nanovdb::Vec3d(origin[0], origin[1], origin[2]));
    auto acc = builder.getAccessor();
    for (size_t i = 0; i < n; ++i) acc.setValue(nanovdb::Coord(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]), values[i]);
    auto* h = new nanovdb::GridHandle<nanovdb::HostBuffer>(nanovdb::tools::createNanoGrid(builder));
    return h;
}
// NanoVDB's own fog-volume sphere (tools/CreatePrimitives.h): narrow-band ramp at the surface, the interior as ACTIVE
// CONSTANT TILES of the internal nodes -- exercises the tile branches of the accessor with a grid the library built itself
void* ref_nvdb_fog_sphere(double radius, const double center[3], double voxel_size, double half_width, const char* name) {
    auto* h = new nanovdb::GridHandle<nanovdb::HostBuffer>(nanovdb::tools::createFogVolumeSphere<float>(
        radius, nanovdb::Vec3d(center[0], center[1], center[2]), voxel_size, half_width, nanovdb::Vec3d(0.0), name));
    return h;
}
// writes the handles as consecutive file segments (io::writeGrids, uncompressed)
int ref_nvdb_write(void* const* handles, int n, const char* path) {
    std::vector<nanovdb::GridHandle<nanovdb::HostBuffer>> v;
    for (int i = 0; i < n; ++i) v.push_back(std::move(*(nanovdb::GridHandle<nanovdb::HostBuffer>*)handles[i]));
    try { nanovdb::io::writeGrids<nanovdb::HostBuffer, std::vector>(path, v); } catch (...) { return -1; }
    for (int i = 0; i < n; ++i) delete (nanovdb::GridHandle<nanovdb::HostBuffer>*)handles[i];
    return 0;
}
// voldata::NanoVDBGrid(path, gridname) (grid_nvdb.cpp:8-28); returned as Grid*
void* ref_nvdb_load(const char* path, const char* gridname) {
    try { return static_cast<Grid*>(new NanoVDBGrid(path, gridname)); } catch (...) { return nullptr; }
}
void ref_nvdb_ibb_min(void* g, int32_t out[3]) {
    auto* n = static_cast<NanoVDBGrid*>((Grid*)g);
    out[0] = n->ibb_min.x; out[1] = n->ibb_min.y; out[2] = n->ibb_min.z;
}
// ---- generic Grid* ----
void ref_grid_info(void* g, uint32_t extent[3], float min_maj[2], float transform[16]) {
    auto* gr = (Grid*)g;
    const glm::uvec3 e = gr->index_extent();
    extent[0] = e.x; extent[1] = e.y; extent[2] = e.z;
    const auto mm = gr->minorant_majorant();
    min_maj[0] = mm.first; min_maj[1] = mm.second;
    memcpy(transform, &gr->transform[0][0], 64);
}
// Grid::lookup on the padded lattice [-2, 8 nb + 2)^3 exactly as BrickGrid(const Grid&) addresses it (grid_brick.cpp:87)
void ref_grid_lookup_padded(void* g, const uint32_t nb[3], float* out) {
    auto* gr = (Grid*)g;
    size_t i = 0;
    for (int z = -2; z < int(nb[2] * 8) + 2; ++z) for (int y = -2; y < int(nb[1] * 8) + 2; ++y) for (int x = -2; x < int(nb[0] * 8) + 2; ++x)
        out[i++] = gr->lookup(glm::uvec3(glm::ivec3(x, y, z)));
}
// A Grid whose lookup() reads a table on the padded lattice [-2, 8 nb + 2)^3 (x fastest, origin (-2, -2, -2)): lets the
// UNMODIFIED BrickGrid(const Grid&) run on arbitrary float values (negative, -0, denormal, > fp16 range, NaN).
struct TableGrid : public Grid {
    const float* val; uint32_t px, py; glm::uvec3 extent;
    float lookup(const glm::uvec3& ipos) const override {
        return val[(size_t(int32_t(ipos.z) + 2) * py + size_t(int32_t(ipos.y) + 2)) * px + size_t(int32_t(ipos.x) + 2)];
    }
    std::pair<float, float> minorant_majorant() const override { return { 0.f, 0.f }; }
    glm::uvec3 index_extent() const override { return extent; }
    size_t num_voxels() const override { return size_t(extent.x) * extent.y * extent.z; }
    size_t size_bytes() const override { return 0; }
};
void* ref_brick_from_values(const float* padded, const uint32_t extent[3], const uint32_t nb[3]) {
    TableGrid t;
    t.val = padded; t.px = nb[0] * 8 + 4; t.py = nb[1] * 8 + 4; t.extent = glm::uvec3(extent[0], extent[1], extent[2]);
    try { return new BrickGrid(t); } catch (std::runtime_error&) { return nullptr; }
}
void* ref_brick_from_grid(void* g) {
    try { return new BrickGrid(*(Grid*)g); } catch (std::runtime_error&) { return nullptr; }
}
void ref_grid_free(void* g) { delete (Grid*)g; }

}
