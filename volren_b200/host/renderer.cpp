// renderer.cpp -- RendererOpenGL on top of the C ABI (include/vrb200.h). Follows reference src/renderer.cpp:
//   init :29-50, resize :52-54, commit :56-76, trace :78-145 (uniform block -> vrb_params), draw :147-153,
//   reset :155-157, scale_and_move_to_unit_cube :227-242.
// Multi-GPU (new): the scene is replicated on every device of the Context; `--partition spp` gives each
// device a contiguous slice of every sample batch (sum buffers, merged by vrb_reduce), `--partition tile` gives each
// device a band of rows (no reduction, rows are gathered).
#include "renderer.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <iostream>

#include "image_io.h"

using namespace volren;
using namespace vmath;

namespace {

void copy3(float* dst, const vec3& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }
void copy9(float* dst, const mat3& m) { memcpy(dst, &m[0].x, 36); }
void copy16(float* dst, const mat4& m) { memcpy(dst, &m[0].x, 64); }

// upload one grid of one frame to one device: BrickGrids verbatim, DenseGrids and NanoVDBGrids through the device brick builder
void upload_grid(vrb_ctx* ctx, int slot, int frame, const voldata::Volume::GridPtr& grid) {
    if (auto dense = std::dynamic_pointer_cast<voldata::DenseGrid>(grid)) {
        const uint32_t dim[3] = { dense->n_voxels.x, dense->n_voxels.y, dense->n_voxels.z };
        const int st = vrb_grid_build_from_dense(ctx, slot, frame, dense->voxel_data.data(), dim, dense->min_value, dense->max_value);
        if (st == VRB_ERR_TOO_MANY_BRICKS) throw std::runtime_error("exceeded max brick count of 1024");
        check(ctx, st, "vrb_grid_build_from_dense");
        return;
    }
    if (auto nvdb = std::dynamic_pointer_cast<voldata::NanoVDBGrid>(grid)) {     // device accessor + brick build, no host round trip
        const int st = vrb_grid_build_from_nvdb(ctx, slot, frame, nvdb->grid_data(), &nvdb->info);
        if (st == VRB_ERR_TOO_MANY_BRICKS) throw std::runtime_error("exceeded max brick count of 1024");
        check(ctx, st, "vrb_grid_build_from_nvdb");
        return;
    }
    const auto brick = voldata::Volume::to_brick_grid(grid);
    const std::string why = brick->check_layout();     // the upload copies n_bricks-sized blocks out of these vectors
    if (!why.empty()) throw std::runtime_error("malformed brick grid: " + why);
    vrb_brick_view v;
    memset(&v, 0, sizeof v);
    for (int a = 0; a < 3; ++a) {
        v.n_bricks[a] = brick->n_bricks[a];
        v.atlas_dim[a] = brick->atlas.stride[a];
    }
    v.brick_count = brick->brick_counter;
    v.indirection = brick->indirection.data.data();
    v.range = brick->range.data.data();
    v.atlas = brick->atlas.data.empty() ? nullptr : brick->atlas.data.data();
    for (int i = 0; i < 3; ++i) v.range_mips[i] = brick->range_mipmaps[i].data.data();
    check(ctx, vrb_grid_upload_brick(ctx, slot, frame, &v), "vrb_grid_upload_brick");
}

}  // namespace

void ColorBuffer::save_ldr(const std::string& path, bool flip, bool) const {
    if (!owner) return;
    owner->sync_image();
    vrb_ctx* ctx = Context::device();
    std::vector<uint8_t> px(size_t(w) * h * 4);
    check(ctx, vrb_download_color_ldr(ctx, px.data()), "vrb_download_color_ldr");
    store_ldr(path, px.data(), int(w), int(h), 4, flip);
}

void RendererOpenGL::init() {
    if (!volume) volume = std::make_shared<voldata::Volume>();
    if (!environment) {   // default environment map: 1x1 white
        const float white[3] = { 1.f, 1.f, 1.f };
        environment = std::make_shared<Environment>(1, 1, white);
    }
    if (!color) {
        const ivec2 res = Context::resolution();
        color.w = uint32_t(res.x);
        color.h = uint32_t(res.y);
        color.owner = this;
    }
}

void RendererOpenGL::resize(uint32_t w, uint32_t h) {
    if (!color) return;
    Context::resize(w, h);
    color.w = w;
    color.h = h;
    sample = 0;
    partial = false;
    root_is_mean = true;
}

void RendererOpenGL::commit() {
    density_grids.clear();
    emission_grids.clear();
    majorant_emission = 0.f;
    generation = Context::generation();
    std::cout << "Preparing brick grids for B200..." << std::endl;
    for (int d = 0; d < Context::n_devices(); ++d) check(Context::device(d), vrb_grid_clear(Context::device(d)), "vrb_grid_clear");
    int frame_idx = 0;
    for (const auto& frame : volume->grids) {
        const voldata::Volume::GridPtr density_grid = frame.at("density");
        for (int d = 0; d < Context::n_devices(); ++d) upload_grid(Context::device(d), VRB_SLOT_DENSITY, frame_idx, density_grid);
        density_grids.push_back(density_grid->transform);
        voldata::Volume::GridPtr emission_grid;
        for (const auto& name : { "flame", "flames", "temperature" }) {
            if (frame.find(name) != frame.end()) {
                emission_grid = frame.at(name);
                break;
            }
        }
        if (emission_grid) {
            // emission grids are indexed by their position in this vector (renderer.cpp:117-118), like the reference
            const int eidx = int(emission_grids.size());
            for (int d = 0; d < Context::n_devices(); ++d) upload_grid(Context::device(d), VRB_SLOT_EMISSION, eidx, emission_grid);
            emission_grids.push_back(emission_grid->transform);
            majorant_emission = std::max(majorant_emission, emission_grid->minorant_majorant().second);
        }
        ++frame_idx;
    }
}

void RendererOpenGL::fill_params(vrb_params& p) const {
    memset(&p, 0, sizeof p);
    const Camera cam = current_camera();
    p.bounces = bounces;
    p.seed = seed;
    p.show_environment = show_environment ? 1 : 0;
    p.frame = int(volume->grid_frame_counter);
    // camera
    copy3(p.cam_pos, cam->pos);
    p.cam_fov = cam->fov_degree;
    copy9(p.cam_transform, inverse(to_mat3(cam->view)));
    // volume
    const auto [bb_min, bb_max] = volume->AABB();
    const auto [mn, maj] = volume->minorant_majorant();
    copy3(p.vol_bb_min, bb_min + vol_clip_min * (bb_max - bb_min));
    copy3(p.vol_bb_max, bb_min + vol_clip_max * (bb_max - bb_min));
    p.vol_minorant = mn * density_scale;
    p.vol_majorant = maj * density_scale;
    p.vol_inv_majorant = 1.f / (maj * density_scale);
    copy3(p.vol_albedo, albedo);
    p.vol_phase_g = phase;
    p.vol_density_scale = density_scale;
    p.vol_emission_scale = emission_scale;
    p.vol_emission_norm = majorant_emission > 0.f ? 1.f / fmaxf(majorant_emission, 1e-4f) : 1.f;
    const mat4 density = volume->transform * density_grids.at(volume->grid_frame_counter);
    copy16(p.vol_density_transform, density);
    copy16(p.vol_density_inv_transform, inverse(density));
    if (volume->grid_frame_counter < emission_grids.size()) {
        const mat4 emission = volume->transform * emission_grids[volume->grid_frame_counter];
        p.has_emission = 1;
        copy16(p.vol_emission_transform, emission);
        copy16(p.vol_emission_inv_transform, inverse(emission));
    }
    // transfer function
    if (transferfunc) {
        p.use_transferfunc = 1;
        p.tf_window_left = transferfunc->window_left;
        p.tf_window_width = transferfunc->window_width;
    }
    // environment
    copy9(p.env_transform, environment->transform);
    copy9(p.env_inv_transform, inverse(environment->transform));
    p.env_strength = environment->strength;
    const ivec2 res = Context::resolution();
    p.resolution[0] = res.x;
    p.resolution[1] = res.y;
}

void RendererOpenGL::push_scene() {
    // the context was re-created: nothing of ours is on the devices any more
    if (generation != Context::generation() && !density_grids.empty()) commit();
    if (scene_generation != Context::generation()) {
        env_uploaded = tf_uploaded_id = tf_uploaded_version = 0;
        scene_generation = Context::generation();
    }
    if (environment->id != env_uploaded) {
        for (int d = 0; d < Context::n_devices(); ++d)
            check(Context::device(d), vrb_env_upload(Context::device(d), environment->pixels.data(), environment->width, environment->height), "vrb_env_upload");
        env_uploaded = environment->id;
    }
    if (transferfunc && (transferfunc->id != tf_uploaded_id || transferfunc->version != tf_uploaded_version)) {
        if (transferfunc->lut_gpu.empty()) throw std::runtime_error("transfer function has an empty LUT");
        for (int d = 0; d < Context::n_devices(); ++d)
            check(Context::device(d), vrb_tf_upload(Context::device(d), &transferfunc->lut_gpu[0].x, uint32_t(transferfunc->lut_gpu.size())), "vrb_tf_upload");
        tf_uploaded_id = transferfunc->id;
        tf_uploaded_version = transferfunc->version;
    }
}

void RendererOpenGL::trace() { trace(1); }

void RendererOpenGL::trace(int n_samples) {
    if (n_samples <= 0) return;
    {   // the colour buffer always has the context's resolution (Context::resolution() drives the dispatch size)
        const ivec2 res = Context::resolution();
        if (!color) color.owner = this;
        color.w = uint32_t(res.x);
        color.h = uint32_t(res.y);
    }
    if (!volume || volume->grids.empty() || density_grids.size() <= volume->grid_frame_counter)
        throw std::runtime_error("RendererOpenGL::trace: no committed density grid for the current frame (call commit())");
    push_scene();
    vrb_params p;
    fill_params(p);
    const int first = sample + 1;   // `++sample` before upload: current_sample is 1-based (renderer.cpp:138)
    const int G = Context::n_devices();
    if (G == 1) {
        check(Context::device(), vrb_trace(Context::device(), &p, first, n_samples, nullptr, VRB_ACCUM_MEAN), "vrb_trace");
    } else if (Context::partition() == "tile") {
        // row bands, 4-row aligned (the tracer walks 8x4 pixel tiles); every device keeps the reference running mean
        const int H = int(color.h), W = int(color.w), bands = (H + 3) / 4;
        for (int d = 0; d < G; ++d) {
            const int y0 = std::min(H, (bands * d / G) * 4), y1 = std::min(H, (bands * (d + 1) / G) * 4);
            if (y0 >= y1) continue;
            const int tile[4] = { 0, y0, W, y1 };
            check(Context::device(d), vrb_trace(Context::device(d), &p, first, n_samples, tile, VRB_ACCUM_MEAN), "vrb_trace");
        }
        partial = true;
    } else {
        // spp slices: device d traces a contiguous slice of this batch into a SUM buffer
        if (sample == 0) {
            for (int d = 0; d < G; ++d) check(Context::device(d), vrb_clear(Context::device(d)), "vrb_clear");
            samples_merged = 0;
        } else if (root_is_mean) {
            check(Context::device(0), vrb_scale(Context::device(0), float(samples_merged)), "vrb_scale");
        }
        root_is_mean = false;
        int s = first;
        for (int d = 0; d < G; ++d) {
            const int n = n_samples * (d + 1) / G - n_samples * d / G;
            if (n > 0) check(Context::device(d), vrb_trace(Context::device(d), &p, s, n, nullptr, VRB_ACCUM_SUM), "vrb_trace");
            s += n;
        }
        partial = true;
    }
    sample += n_samples;
}

void RendererOpenGL::sync_image() {
    const int G = Context::n_devices();
    if (G == 1 || !partial) return;
    std::vector<vrb_ctx*> ctxs;
    for (int d = 0; d < G; ++d) ctxs.push_back(Context::device(d));
    if (Context::partition() == "tile") {
        const int H = int(color.h), bands = (H + 3) / 4;
        for (int d = 1; d < G; ++d) {
            const int y0 = std::min(H, (bands * d / G) * 4), y1 = std::min(H, (bands * (d + 1) / G) * 4);
            if (y0 < y1) check(ctxs[0], vrb_copy_rows(ctxs[0], ctxs[d], y0, y1), "vrb_copy_rows");
        }
    } else {
        check(ctxs[0], vrb_reduce(ctxs.data(), G, 0), "vrb_reduce");
        for (int d = 1; d < G; ++d) check(ctxs[d], vrb_clear(ctxs[d]), "vrb_clear");
        check(ctxs[0], vrb_scale(ctxs[0], 1.f / float(sample)), "vrb_scale");
        samples_merged = sample;
        root_is_mean = true;
    }
    partial = false;
}

void RendererOpenGL::draw() {
    if (!color) return;
    sync_image();
    vrb_ctx* ctx = Context::device();
    check(ctx, vrb_tonemap(ctx, tonemap_exposure, tonemap_gamma, 0, tonemapping ? 1 : 0), "vrb_tonemap");
}

void RendererOpenGL::tonemap_in_place() {
    sync_image();
    vrb_ctx* ctx = Context::device();
    check(ctx, vrb_tonemap(ctx, tonemap_exposure, tonemap_gamma, 1, 1), "vrb_tonemap");
}

void RendererOpenGL::reset() { sample = 0; }

std::vector<float> RendererOpenGL::read_color(int channels) {
    sync_image();
    vrb_ctx* ctx = Context::device();
    std::vector<float> out(size_t(color.w) * color.h * channels);
    check(ctx, vrb_download_color(ctx, out.data(), channels), "vrb_download_color");
    return out;
}

std::vector<uint8_t> RendererOpenGL::read_framebuffer() {
    vrb_ctx* ctx = Context::device();
    std::vector<uint8_t> out(size_t(color.w) * color.h * 4);
    check(ctx, vrb_download_framebuffer(ctx, out.data()), "vrb_download_framebuffer");
    return out;
}

void RendererOpenGL::scale_and_move_to_unit_cube() {
    // max AABB over the whole volume (animation), from two corners per frame like the reference
    vec3 bb_min = vec3(FLT_MAX), bb_max = vec3(FLT_MIN);
    for (const auto& frame : volume->grids) {
        const auto grid = frame.at("density");
        bb_min = min(bb_min, vec3(grid->transform * vec4(0, 0, 0, 1)));
        bb_max = max(bb_max, vec3(grid->transform * vec4(vec3(grid->index_extent()), 1)));
    }
    const vec3 extent = bb_max - bb_min;
    const float size = fmaxf(extent.x, fmaxf(extent.y, extent.z));
    if (size != 1.f) {
        volume->transform = translate(scale(mat4(1.f), vec3(1.f / size)), -bb_min - 0.5f * extent);
        density_scale *= size;
    }
}
