/*
 * glsl_ref.cpp -- TEST INFRASTRUCTURE ONLY: the host side of oracle/_ref/libglsl_ref.so.
 *
 * Plays the role of RendererOpenGL::trace (src/renderer.cpp:78-145), Environment::Environment (src/environment.cpp:11-33)
 * and the offline tonemap dispatch (src/main.cpp:540-550) for the reference's own shader text, compiled as C++ through
 * glsl2cpp.py + glsl_shim.h: assigns the "uniforms" field for field from vrb_params (which mirrors that uniform block),
 * binds the brick textures / environment / LUT, and runs main() once per (pixel, sample) -- one glDispatchCompute per
 * sample, exactly the loop of src/main.cpp:533-537.
 *
 * Used only by tests/ to pin oracle/vr_oracle.c: the restatement must equal this library bit for bit.
 */
#include "glsl_shim.h"

#include <cstring>

#include "../vr_oracle.h"

namespace glsl {
thread_local invocation_id gl_GlobalInvocationID;

namespace pt {
#include "pathtracer_brick.inc"
}
namespace pt_tf {
#include "pathtracer_brick_tf.inc"
}
namespace env_setup {
#include "env_setup.inc"
}
namespace tonemap {
#include "tonemap.inc"
}

static mat3 m3(const float* m) { return glm::mat3(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]); }
static mat4 m4(const float* m) {
    return glm::mat4(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15]);
}
static vec3 v3(const float* v) { return vec3(v[0], v[1], v[2]); }

static void bind_grid(const vro_grid& g, usampler3D& ind, sampler3D& range, sampler3D& atlas) {
    const ivec3 nb(int(g.n_bricks[0]), int(g.n_bricks[1]), int(g.n_bricks[2]));
    ind.data = g.indirection;                       /* renderer.cpp:161-168 */
    ind.dim = nb;
    range = sampler3D();                            /* renderer.cpp:177-205: level 0 + 3 manually uploaded mips */
    range.rg16f[0] = g.range;
    for (int i = 0; i < 3; ++i) range.rg16f[i + 1] = g.range_mips[i];
    range.dim = nb;
    range.levels = 4;
    atlas = sampler3D();                            /* renderer.cpp:208-215 */
    atlas.r8 = g.atlas;
    atlas.dim = ivec3(int(g.atlas_dim[0]), int(g.atlas_dim[1]), int(g.atlas_dim[2]));
    atlas.levels = 1;
}

/* the uniform block of renderer.cpp:88-139; a macro because the two programs own separate copies of every uniform */
#define GLSL_SET_UNIFORMS(NS)                                                                                        \
    do {                                                                                                             \
        NS::bounces = p->bounces;                                                                                    \
        NS::u_seed = p->seed;                                                                                        \
        NS::show_environment = p->show_environment;                                                                  \
        NS::cam_pos = v3(p->cam_pos);                                                                                \
        NS::cam_fov = p->cam_fov;                                                                                    \
        NS::cam_transform = m3(p->cam_transform);                                                                    \
        NS::vol_bb_min = v3(p->vol_bb_min);                                                                          \
        NS::vol_bb_max = v3(p->vol_bb_max);                                                                          \
        NS::vol_minorant = p->vol_minorant;                                                                          \
        NS::vol_majorant = p->vol_majorant;                                                                          \
        NS::vol_inv_majorant = p->vol_inv_majorant;                                                                  \
        NS::vol_albedo = v3(p->vol_albedo);                                                                          \
        NS::vol_phase_g = p->vol_phase_g;                                                                            \
        NS::vol_density_scale = p->vol_density_scale;                                                                \
        NS::vol_emission_scale = p->vol_emission_scale;                                                              \
        NS::vol_emission_norm = p->vol_emission_norm;                                                                \
        NS::vol_density_transform = m4(p->vol_density_transform);                                                    \
        NS::vol_density_inv_transform = m4(p->vol_density_inv_transform);                                            \
        bind_grid(scene->density, NS::vol_density_indirection, NS::vol_density_range, NS::vol_density_atlas);        \
        if (p->has_emission) {                                                                                       \
            NS::vol_emission_transform = m4(p->vol_emission_transform);                                              \
            NS::vol_emission_inv_transform = m4(p->vol_emission_inv_transform);                                      \
            bind_grid(scene->emission, NS::vol_emission_indirection, NS::vol_emission_range, NS::vol_emission_atlas); \
        } else { /* never set: GL default-initialises uniforms to 0 and unbound samplers fetch 0 */                  \
            NS::vol_emission_transform = mat4(0.f);                                                                  \
            NS::vol_emission_inv_transform = mat4(0.f);                                                              \
            NS::vol_emission_indirection = usampler3D();                                                             \
            NS::vol_emission_range = sampler3D();                                                                    \
            NS::vol_emission_atlas = sampler3D();                                                                    \
        }                                                                                                            \
        NS::tf_lut = reinterpret_cast<const vec4*>(scene->tf_lut);                                                   \
        NS::tf_size = scene->tf_size;                                                                                \
        NS::tf_window_left = p->tf_window_left;                                                                      \
        NS::tf_window_width = p->tf_window_width;                                                                    \
        NS::env_transform = m3(p->env_transform);                                                                    \
        NS::env_inv_transform = m3(p->env_inv_transform);                                                            \
        NS::env_strength = p->env_strength;                                                                          \
        NS::env_imp_inv_dim = vec2(1.f / 512.f);                                                                     \
        NS::env_imp_base_mip = 9;                                                                                    \
        NS::env_envmap = envmap;                                                                                     \
        NS::env_impmap = impmap;                                                                                     \
        NS::resolution = ivec2(p->resolution[0], p->resolution[1]);                                                  \
        NS::color = img;                                                                                             \
    } while (0)

}  // namespace glsl

using namespace glsl;

extern "C" {

/* RendererOpenGL::trace() called n_samples times with sample = first_sample - 1 (renderer.cpp:138: ++sample before upload).
 * color: W*H*4 floats, bottom-up, read and written (running mean). impmap pyramid as in vro_scene. */
void glslref_trace(const vro_scene* scene, const vrb_params* p, int first_sample, int n_samples, float* color, int n_threads) {
    sampler2D envmap;
    envmap.level[0] = scene->env_rgb;
    envmap.w = scene->env_w;
    envmap.h = scene->env_h;
    envmap.channels = 3;
    envmap.levels = 1;
    sampler2D impmap;
    {
        size_t off = 0;
        for (int l = 0; l < 10; ++l) {
            impmap.level[l] = scene->impmap + off;
            off += size_t(512 >> l) * size_t(512 >> l);
        }
        impmap.w = impmap.h = 512;
        impmap.channels = 1;
        impmap.levels = 10;
    }
    image2D img;
    img.data = color;
    img.w = p->resolution[0];
    img.h = p->resolution[1];
    img.channels = 4;
    const bool tf = p->use_transferfunc != 0;
    if (tf) GLSL_SET_UNIFORMS(pt_tf); else GLSL_SET_UNIFORMS(pt);
    const int W = p->resolution[0], H = p->resolution[1];
    /* cppgl Shader::dispatch_compute: ceil(w/16) x ceil(h/16) groups of 16x16 (cppgl/shader.cpp:299-305) */
    const int GW = (W + 15) / 16 * 16, GH = (H + 15) / 16 * 16;
    (void)n_threads;
    for (int s = first_sample; s < first_sample + n_samples; ++s) {
        if (tf) pt_tf::current_sample = s; else pt::current_sample = s;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads > 0 ? n_threads : 1)
        for (int y = 0; y < GH; ++y)
            for (int x = 0; x < GW; ++x) {
                gl_GlobalInvocationID.xy = uvec2(uint(x), uint(y));
                if (tf) pt_tf::main(); else pt::main();
            }
    }
}

/* Environment::Environment (environment.cpp:19-27): level 0 of the importance map, 512^2 floats */
void glslref_env_setup(const float* rgb, int w, int h, float* impmap_level0) {
    sampler2D envmap;
    envmap.level[0] = rgb;
    envmap.w = w;
    envmap.h = h;
    envmap.channels = 3;
    envmap.levels = 1;
    const uint32_t DIMENSION = 512, SAMPLES = 64;
    const uint32_t n_samples = (uint32_t)std::sqrt(SAMPLES);
    image2D out;
    out.data = impmap_level0;
    out.w = out.h = int(DIMENSION);
    out.channels = 1;
    env_setup::impmap = out;
    env_setup::envmap = envmap;
    env_setup::output_size = ivec2(DIMENSION);
    env_setup::output_size_samples = ivec2(DIMENSION * n_samples, DIMENSION * n_samples);
    env_setup::num_samples = ivec2(n_samples, n_samples);
    env_setup::inv_samples = 1.f / (n_samples * n_samples);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < int(DIMENSION); ++y)
        for (int x = 0; x < int(DIMENSION); ++x) {
            gl_GlobalInvocationID.xy = uvec2(uint(x), uint(y));
            env_setup::main();
        }
}

/* the in-place tonemap dispatch of the offline loop (main.cpp:540-550) */
void glslref_tonemap(float* color, int w, int h, float exposure, float gamma) {
    image2D img;
    img.data = color;
    img.w = w;
    img.h = h;
    img.channels = 4;
    tonemap::color = img;
    tonemap::exposure = exposure;
    tonemap::gamma = gamma;
    tonemap::resolution = ivec2(w, h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            gl_GlobalInvocationID.xy = uvec2(uint(x), uint(y));
            tonemap::main();
        }
}

/* helpers the tests compare one by one */
uint32_t glslref_tea(uint32_t v0, uint32_t v1, uint32_t n) { return pt::tea(v0, v1, n); }
float glslref_rng(uint32_t* state) { return pt::rng(*state); }

}  // extern "C"
