// renderer.h -- the VolRen renderer host API (reference src/renderer.h:16-63), kept name for name: six methods
// (init/resize/commit/trace/draw/reset), the unit-cube helper and the public data members that the CLI
// (main.cpp:360-435) and the Python module (bindings.cpp:167-185) mutate directly. The "OpenGL data" block is
// replaced by device bookkeeping: grids, environment and LUT live in the vrb_ctx of every GPU (include/vrb200.h).
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "camera.h"
#include "context.h"
#include "environment.h"
#include "transferfunc.h"
#include "voldata.h"

// stands in for the RGBA32F `cppgl::Texture2D color` (renderer.h:48): a handle on the device colour buffer
struct ColorBuffer {
    uint32_t w = 0, h = 0;
    explicit operator bool() const { return w > 0 && h > 0; }
    // Texture2DImpl::save_ldr (cppgl texture.cpp:107-113): float -> unorm8 readback, flipped on write by default
    void save_ldr(const std::string& path, bool flip = true, bool async = false) const;
    struct RendererOpenGL* owner = nullptr;
};

struct RendererOpenGL {
    // Renderer interface (renderer.h:18-23)
    void init();
    void resize(uint32_t w, uint32_t h);
    void commit();
    void trace();      // exactly one more sample per pixel; ++sample
    void draw();       // tonemap.fs / blit.fs into the RGBA8 framebuffer
    void reset();

    // scale and move volume to fit into [-0.5, 0.5] unit cube (renderer.h:28)
    void scale_and_move_to_unit_cube();

    // ---- additions of the B200 host ----
    // n more samples in ONE launch per GPU; the image equals n successive trace() calls (the running mean is
    // evaluated in sample order in registers). render(spp) and the offline loop use this.
    void trace(int n_samples);
    // shader/tonemap.glsl in place on `color` (the offline CLI path, main.cpp:540-550)
    void tonemap_in_place();
    // make `color` on device 0 the finished mean image (multi-GPU: reduce / gather partial results)
    void sync_image();
    // readbacks (bindings.cpp:141-166)
    std::vector<float> read_color(int channels = 3);
    std::vector<uint8_t> read_framebuffer();
    // the uniform block trace() would upload right now
    vrb_params debug_params() const { vrb_params p; fill_params(p); return p; }

    // General settings (renderer.h:31-38)
    int sample = 0;
    int sppx = 1024;
    int seed = 42;
    int bounces = 100;
    float tonemap_exposure = 5.f;
    float tonemap_gamma = 2.2f;
    bool tonemapping = true;
    bool show_environment = true;

    // Volume settings (:41-44)
    glm::vec3 albedo = glm::vec3(0.9f);
    float phase = 0.f;
    float density_scale = 1.f;
    float emission_scale = 100.f;

    // device data (replaces :47-51)
    ColorBuffer color;
    std::vector<glm::mat4> density_grids;    // per frame: the grid transform captured at commit() (BrickGridGL::transform)
    std::vector<glm::mat4> emission_grids;
    float majorant_emission = 0.f;

    // Volume data (:54)
    std::shared_ptr<voldata::Volume> volume;

    // Volume clip planes (:57-58)
    glm::vec3 vol_clip_min = glm::vec3(0.f);
    glm::vec3 vol_clip_max = glm::vec3(1.f);

    // Scene data (:61-62)
    std::shared_ptr<Environment> environment;
    std::shared_ptr<TransferFunction> transferfunc;

private:
    void fill_params(vrb_params& p) const;
    void push_scene();                       // (re)upload environment / LUT when the bound objects changed
    uint64_t env_uploaded = 0, tf_uploaded_id = 0, tf_uploaded_version = 0;
    bool partial = false;                    // multi-GPU: device buffers hold un-merged partial results
    bool root_is_mean = true;                // spp partition: device 0 holds the mean (true) or a running sum (false)
    int samples_merged = 0;
    uint64_t generation = 0, scene_generation = 0;   // Context generation the device-side grids / env+LUT belong to
};

using Renderer = RendererOpenGL;
