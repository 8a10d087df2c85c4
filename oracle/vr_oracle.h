/*
 * vr_oracle.h -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference algorithms).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library. The product (libvrb200.so, the C++ host, volpy) never links or calls it.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - brick / dense / half / mip layer: PINNED bit-exactly against the compiled reference voldata
 *     sources (oracle/_ref) and data/smoke.brick invariants.
 *   - LUT CDF: pinned against data/lut.txt semantics of transferfunc.cpp:33-58 (arithmetic restated).
 *   - shader layer (tracking, NEE, env importance map, tonemap): PARITY UNPINNED -- the GLSL reference
 *     cannot execute in this environment (no GL/EGL/OSMesa); the restatement follows common.glsl
 *     line by line and is only loosely anchored on imgs/example.jpg.
 */
#ifndef VR_ORACLE_H
#define VR_ORACLE_H

#include <stdint.h>
#include <stddef.h>
#include "../include/vrb200.h" /* vrb_params, vrb_brick_view, vrb_counters: shared PODs only */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vro_grid {
    uint32_t n_bricks[3];
    uint32_t atlas_dim[3];
    const uint32_t* indirection;
    const uint32_t* range;
    const uint8_t* atlas;
    const uint32_t* range_mips[3];
} vro_grid;

typedef struct vro_scene {
    vro_grid density;
    vro_grid emission;            /* only read when params->has_emission */
    const float* env_rgb;         /* w*h*3, bottom-up */
    int32_t env_w, env_h;
    const float* impmap;          /* pyramid, levels 0..9 concatenated: 512^2, 256^2, ..., 1 */
    const float* tf_lut;          /* tf_size * 4 */
    uint32_t tf_size;
} vro_scene;

uint32_t vro_tea(uint32_t v0, uint32_t v1, uint32_t n);
float vro_rng(uint32_t* state);
void vro_rng_stream(uint32_t seed, int n, float* out, uint32_t* states_out);

uint16_t vro_to_half(float f);
float vro_from_half(uint16_t h);
uint32_t vro_encode_range(float lo, float hi);

void vro_dense_from_float(const float* data, const uint32_t dim[3], uint8_t* out_u8, float out_minmax[2]);
float vro_dense_lookup(const uint8_t* vox, const uint32_t dim[3], float vmin, float vmax, uint32_t x, uint32_t y, uint32_t z);

int vro_brick_dims(const uint32_t dim[3], uint32_t n_bricks[3]);
int vro_brick_build(const uint8_t* vox, const uint32_t dim[3], float vmin, float vmax, vrb_brick_view* out);
/* the same constructor for any other Grid: lookup() values tabulated on the padded lattice [-2, 8 nb + 2)^3 */
int vro_brick_build_values(const float* padded_values, const uint32_t extent[3], vrb_brick_view* out);
float vro_brick_lookup(const vro_grid* g, uint32_t x, uint32_t y, uint32_t z);

int vro_lut_upload(const float* rgba, uint32_t n, float* out);

void vro_env_build(const float* rgb, int w, int h, float* pyramid_out);
size_t vro_env_pyramid_floats(void);

void vro_trace(const vro_scene* scene, const vrb_params* p, int first_sample, int n_samples,
               const int tile[4], int accum_mode, float* color, vrb_counters* counters, int n_threads);
void vro_trace_deterministic(const vro_scene* scene, const vrb_params* p, float* color, int n_threads);

void vro_tonemap_inplace(float* color, int w, int h, float exposure, float gamma);
void vro_draw(const float* color, int w, int h, float exposure, float gamma, int tonemapping, uint8_t* rgba8);
void vro_color_to_ldr(const float* color, int w, int h, uint8_t* rgba8);

int vro_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
