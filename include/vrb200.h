/*
 * vrb200.h -- C ABI of libvrb200.so, the B200 (sm_100a) back end of the VolRen volume path tracer.
 *
 * This is the drop-in boundary for the reference's GPU seam: everything nihofm/volren does through
 * OpenGL (texture/SSBO uploads, uniform marshalling, glDispatchCompute, glGetTexImage) is replaced
 * by the calls below. Each entry point cites the reference interface it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - every function returns VRB_OK (0) or a negative vrb_status; nothing throws across the ABI;
 *     a human-readable message for the last failure of a context is vrb_last_error(ctx).
 *   - a vrb_ctx is bound to ONE CUDA device and is NOT thread-safe (one host thread per ctx).
 *   - all work is enqueued on the context's stream; only vrb_sync, vrb_*download* and
 *     vrb_get_counters block the host.
 *   - host pointers are borrowed for the duration of the call; device memory is owned by the ctx.
 *   - matrices are column-major (glm layout), images are bottom-up (row 0 = bottom, GL layout),
 *     3-D buffers are x-fastest (voldata/src/buf3d.h:27-29).
 *   - there is NO CPU fallback: without a CUDA device vrb_create fails with VRB_ERR_NO_DEVICE.
 */
#ifndef VRB200_H
#define VRB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRB_ABI_VERSION 1

typedef struct vrb_ctx vrb_ctx;

typedef enum vrb_status {
    VRB_OK = 0,
    VRB_ERR_INVALID = -1,      /* bad argument (null pointer, bad slot/frame, size 0, ...) */
    VRB_ERR_NO_DEVICE = -2,    /* no usable CUDA device / wrong architecture */
    VRB_ERR_CUDA = -3,         /* a CUDA runtime call or kernel failed (see vrb_last_error) */
    VRB_ERR_OOM = -4,          /* device allocation failed */
    VRB_ERR_TOO_MANY_BRICKS = -5, /* mirrors "exceeded max brick count of 1024" (grid_brick.cpp:66-67) */
    VRB_ERR_STATE = -6         /* call order problem: no grid / env / resolution set */
} vrb_status;

/* grid slots of one frame (renderer.cpp:61-74: "density" and the first of flame|flames|temperature) */
enum { VRB_SLOT_DENSITY = 0, VRB_SLOT_EMISSION = 1 };

/* accumulation modes of vrb_trace */
enum {
    VRB_ACCUM_MEAN = 0, /* reference semantics: color = mix(color, L, 1/current_sample) (pathtracer_brick.glsl:36) */
    VRB_ACCUM_SUM = 1   /* color += L  (spp-sliced multi-GPU: sum buffers are reduced, then vrb_scale) */
};

/*
 * The uniform block that RendererOpenGL::trace uploads (src/renderer.cpp:88-139), field for field.
 * tf_size, env_imp_inv_dim (1/512) and env_imp_base_mip (9) are owned by the context
 * (they follow from vrb_tf_upload / vrb_env_upload, as in transferfunc.cpp:28 and renderer.cpp:130-131).
 */
typedef struct vrb_params {
    int32_t bounces;                 /* renderer.cpp:88 */
    int32_t seed;                    /* :89 (signed: the shader multiplies it as int, pathtracer_brick.glsl:28) */
    int32_t show_environment;        /* :90 */
    int32_t frame;                   /* volume->grid_frame_counter (:110) */
    float cam_pos[3];                /* :93 */
    float cam_fov;                   /* :94 degrees */
    float cam_transform[9];          /* :95 inverse(mat3(view)), column-major */
    float vol_bb_min[3];             /* :99 */
    float vol_bb_max[3];             /* :100 */
    float vol_minorant;              /* :101 */
    float vol_majorant;              /* :102 */
    float vol_inv_majorant;          /* :103 */
    float vol_albedo[3];             /* :104 */
    float vol_phase_g;               /* :105 */
    float vol_density_scale;         /* :106 */
    float vol_emission_scale;        /* :107 */
    float vol_emission_norm;         /* :108 */
    float vol_density_transform[16];     /* :111 volume.transform * grid.transform */
    float vol_density_inv_transform[16]; /* :112 */
    int32_t has_emission;            /* :117 frame < emission_grids.size() */
    float vol_emission_transform[16];     /* :119 */
    float vol_emission_inv_transform[16]; /* :120 (all-zero when has_emission == 0: unset GL uniform) */
    int32_t use_transferfunc;        /* :80 selects pathtracer_brick_tf.glsl */
    float tf_window_left;            /* transferfunc.cpp:29 */
    float tf_window_width;           /* transferfunc.cpp:30 */
    float env_transform[9];          /* renderer.cpp:127 */
    float env_inv_transform[9];      /* :128 */
    float env_strength;              /* :129 */
    int32_t resolution[2];           /* :137,139 */
} vrb_params;

/* Host view of one brick grid (voldata/src/grid_brick.h:27-33). Pointers may be NULL to skip a copy. */
typedef struct vrb_brick_view {
    uint32_t n_bricks[3];            /* grid_brick.h:27 */
    uint32_t atlas_dim[3];           /* atlas.stride in voxels (pruned in z, grid_brick.cpp:112) */
    uint64_t brick_count;            /* brick_counter */
    uint32_t* indirection;           /* n_bricks.x*y*z words, encode_ptr packing (grid_brick.cpp:32-37) */
    uint32_t* range;                 /* n_bricks.x*y*z words, 2 x fp16 (grid_brick.cpp:24-26) */
    uint8_t* atlas;                  /* atlas_dim.x*y*z bytes */
    uint32_t* range_mips[3];         /* level i: (n_bricks >> (i+1)) words (grid_brick.cpp:114-141) */
} vrb_brick_view;

/* Event counters of the tracking kernels: they define the algorithmic bytes per sample. */
typedef struct vrb_counters {
    uint64_t n_samples;   /* path samples (pixel, spp) traced */
    uint64_t n_maj;       /* majorant DDA steps, camera + shadow rays (common.glsl:425/427, 472/474) */
    uint64_t n_dens;      /* tentative collisions = density lookups (common.glsl:437-440, 484-487) */
    uint64_t n_emis;      /* emission-grid lookups actually fetched (common.glsl:489) */
    uint64_t n_nee;       /* sample_environment calls (common.glsl:616) */
    uint64_t n_env;       /* escape lookups lookup_environment (+pdf) (common.glsl:646-647) */
    uint64_t n_real;      /* real collisions = path vertices */
} vrb_counters;

/* ---- context ---------------------------------------------------------------------------------- */
/* replaces cppgl Context::init + RendererOpenGL::init (src/renderer.cpp:29-50) */
int vrb_create(int device, vrb_ctx** out);
void vrb_destroy(vrb_ctx* ctx);
const char* vrb_last_error(vrb_ctx* ctx);
const char* vrb_status_string(int status);
int vrb_abi_version(void);
/* external != 0: run all work of this context on the given cudaStream_t (NULL = the legacy default stream, which is
 * what torch.cuda.current_stream() is unless changed); external == 0: back to the context's own non-blocking stream */
int vrb_set_stream(vrb_ctx* ctx, void* cuda_stream, int external);
int vrb_sync(vrb_ctx* ctx);
/* replaces RendererOpenGL::resize / the RGBA32F `color` texture (src/renderer.cpp:46-54); zero-fills */
int vrb_resize(vrb_ctx* ctx, int w, int h);

/* ---- volume data (src/renderer.cpp:56-76 commit, :159-225 brick_grid_to_textures) --------------- */
/* commit() starts with density_grids.clear(); emission_grids.clear() (renderer.cpp:57-58) */
int vrb_grid_clear(vrb_ctx* ctx);
/* release one grid (frees its device memory); no-op if absent */
int vrb_grid_free(vrb_ctx* ctx, int slot, int frame);
/* upload an existing BrickGrid verbatim (a loaded .brick file) */
int vrb_grid_upload_brick(vrb_ctx* ctx, int slot, int frame, const vrb_brick_view* grid);
/* voldata::BrickGrid::BrickGrid(const Grid&) for a DenseGrid source (voldata/src/grid_brick.cpp:60-142,
 * grid_dense.cpp:99-103), run on the GPU, bit-exact w.r.t. the serial reference (raster allocation order).
 * voxels_u8: dim[0]*dim[1]*dim[2] bytes, HOST memory; (vmin, vmax) = DenseGrid::min_value/max_value. */
int vrb_grid_build_from_dense(vrb_ctx* ctx, int slot, int frame, const uint8_t* voxels_u8,
                              const uint32_t dim[3], float vmin, float vmax);
/* same, voxels already resident in DEVICE memory of ctx's device. The build runs on the context's stream (vrb_set_stream): the
 * voxels must be complete with respect to THAT stream -- written on it, or the writer synchronised -- like any CUDA consumer. */
int vrb_grid_build_from_dense_device(vrb_ctx* ctx, int slot, int frame, const void* d_voxels_u8,
                                     const uint32_t dim[3], float vmin, float vmax);
/* voldata::DenseGrid(w,h,d,const float*) (grid_dense.cpp:57-95) followed by BrickGrid(const Grid&) for FLOAT voxels resident in
 * device memory (the C3 pipeline: a synthetic fp32 field -> DenseGrid -> bricks without leaving the GPU): global min/max with
 * the reference's initial values, 8-bit quantisation, brick build. out_minmax (may be NULL) receives DenseGrid::min_value /
 * max_value. Same stream rule as vrb_grid_build_from_dense_device. */
int vrb_grid_build_from_float_device(vrb_ctx* ctx, int slot, int frame, const void* d_voxels_f32,
                                     const uint32_t dim[3], float out_minmax[2]);
/* voldata::BrickGrid::BrickGrid(const Grid&) for ANY Grid source (grid_brick.cpp:60-142 with the virtual Grid::lookup,
 * e.g. NanoVDBGrid, grid_nvdb.cpp:64-67): the caller evaluates grid.lookup(uvec3(x, y, z)) once for every voxel of the
 * padded lattice x in [-2, 8 n_bricks.x + 2) (likewise y, z; negative coordinates wrap to uint32 as in the reference's
 * dilated windows, :87) into padded_values[(z + 2) * py * px + (y + 2) * px + (x + 2)]; vrb_brick_lattice returns
 * n_bricks = roundup8(ceil(extent / 8)) (:62) and the padded dimensions (px, py, pz) = 8 n_bricks + 4.
 * Bit-exact with the serial reference including its first-seen std::min/std::max order. Host memory, borrowed. */
int vrb_brick_lattice(const uint32_t extent[3], uint32_t n_bricks[3], uint32_t padded_dim[3]);
int vrb_grid_build_from_values(vrb_ctx* ctx, int slot, int frame, const float* padded_values, const uint32_t extent[3]);

/* ---- NanoVDB sources (voldata/src/grid_nvdb.cpp; NanoVDB ABI 32 as pinned by the reference's openvdb submodule) ---- */
/* What voldata::NanoVDBGrid::NanoVDBGrid(path, gridname) (grid_nvdb.cpp:8-28) derives from a grid. */
typedef struct vrb_nvdb_info {
    uint64_t grid_offset;            /* byte offset of the serialized grid buffer inside the file image */
    uint64_t grid_size;              /* GridData::mGridSize */
    uint64_t active_voxels;          /* NanoVDBGrid::num_voxels (grid_nvdb.cpp:78-80) */
    uint32_t extent[3];              /* index_extent: ibb.max - ibb.min + 1, 0 for an empty grid (:15) */
    int32_t ibb_min[3];              /* :14 */
    float minorant, majorant;        /* root minimum / maximum (:16-17) */
    float transform[16];             /* :19-27 map matrix + translation, shifted by ibb_min; column-major */
} vrb_nvdb_info;
/* nanovdb::io::readGrid(path, gridname) + the constructor's checks on the bytes of a .nvdb file (segment files and raw
 * grid buffers, codec NONE -- the reference builds NanoVDB without ZIP/BLOSC): locates the first grid of that name,
 * verifies that it is a valid float fog volume whose node offsets stay inside the buffer, fills *out. Pure host code,
 * no context needed. On failure returns VRB_ERR_INVALID with the reference's exception text in err[err_len]. */
int vrb_nvdb_open(const void* file, size_t bytes, const char* gridname, vrb_nvdb_info* out, char* err, size_t err_len);
/* NanoVDBGrid::lookup (grid_nvdb.cpp:64-67) for n index positions ipos[3 n] (uint32, as Grid::lookup takes them) of a
 * grid buffer accepted by vrb_nvdb_open (grid = file + grid_offset). Host code: the Grid interface of the host model. */
int vrb_nvdb_lookup(const void* grid, const int32_t ibb_min[3], const uint32_t* ipos, size_t n, float* out);
/* voldata::BrickGrid::BrickGrid(const Grid&) for a NanoVDBGrid source (what Volume::current_grid_brick does with a
 * loaded .nvdb, volume.cpp:89-91 -> grid_brick.cpp:60-142): one H2D copy of the grid buffer, lookup() tabulated on the
 * padded brick lattice by a device accessor, then the any-Grid brick build. Bit-exact w.r.t. the serial reference. */
int vrb_grid_build_from_nvdb(vrb_ctx* ctx, int slot, int frame, const void* grid, const vrb_nvdb_info* info);
/* sizes of an uploaded/built grid; then a second call with buffers allocated copies it back */
int vrb_grid_info(vrb_ctx* ctx, int slot, int frame, vrb_brick_view* sizes_out);
int vrb_grid_download(vrb_ctx* ctx, int slot, int frame, vrb_brick_view* out);
/* test hook: the density the tracer itself fetches at n index-space points (x, y, z triples, host memory).
 * mode 0 = lookup_density_trilinear (common.glsl:289-297, without density_scale) through the records + u8 atlas,
 * mode 1 = the same through the decoded apron blocks the production kernel reads (must equal mode 0 bit for bit),
 * mode 2 = lookup_density_brick at floor(p) (common.glsl:268-275 == BrickGrid::lookup, grid_brick.cpp:148-154; 0 outside) */
int vrb_debug_sample_density(vrb_ctx* ctx, int slot, int frame, const float* ipos_xyz, size_t n, int mode, float* out);
/* voldata::DenseGrid(w,h,d,const float*) (voldata/src/grid_dense.cpp:57-95): global min/max + 8-bit quantise.
 * out_u8 (host, n bytes) and out_minmax[2]; if d_out_u8 != NULL the quantised grid is also left on the device. */
int vrb_dense_from_float(vrb_ctx* ctx, const float* data, const uint32_t dim[3], uint8_t* out_u8,
                         float out_minmax[2]);

/* ---- environment (src/environment.cpp:11-33 + shader/env_setup.glsl) ---------------------------- */
/* rgb: w*h*3 floats, bottom-up (already flipped like cppgl image_load). Builds the 512^2 importance
 * map and its 9 box-filter mips (glGenerateMipmap). */
int vrb_env_upload(vrb_ctx* ctx, const float* rgb, int w, int h);
/* read back importance-map level `level` (0..9): (512>>level)^2 floats */
int vrb_env_download_impmap(vrb_ctx* ctx, int level, float* out);

/* ---- transfer function (src/transferfunc.cpp:26-58) -------------------------------------------- */
/* rgba: n*4 floats, already CDF-corrected by the host (compute_lut_cdf runs above the ABI) */
int vrb_tf_upload(vrb_ctx* ctx, const float* rgba, uint32_t n);

/* ---- rendering (src/renderer.cpp:78-145 trace, shader/pathtracer_brick{,_tf}.glsl) -------------- */
/* Traces samples first_sample .. first_sample+n_samples-1 (1-based, = `current_sample`) for every pixel
 * of `tile` (x0,y0,x1,y1 half-open; NULL = whole image) and folds them into the colour buffer. */
int vrb_trace(vrb_ctx* ctx, const vrb_params* params, int first_sample, int n_samples,
              const int tile[4], int accum_mode);
/* deterministic transmittance-only mode (defined by this build, DESIGN.md "T1"): centre ray per pixel,
 * exact voxel DDA, color = (Tr * Le_env(dir), 1 - Tr) */
int vrb_trace_deterministic(vrb_ctx* ctx, const vrb_params* params);
/* kernel selection: 0 = ray-pool persistent kernel with MUFU fast math (default, production: every warp schedules a pool
 * of 64 path states in shared memory, csrc/vr_trace_pool.cuh);
 * 1 = the straightforward one-thread-per-pixel kernel with IEEE math and no FMA contraction; 2 = the lane-resident
 * persistent kernel with the same IEEE math. 1 and 2 are cross-checks compiled in their own translation unit
 * (csrc/vrb200_strict.cu, -fmad=false): they replay the CPU oracle -- and through it the reference's GLSL -- path for path
 * up to the last bit of libm, and 2 must reproduce 1 exactly (same paths, same counters);
 * 3 = the lane-resident persistent kernel with fast math (round 1's production schedule): bit-identical images to 0.
 * Environment variable VRB200_KERNEL sets the default. */
int vrb_set_kernel(vrb_ctx* ctx, int kind);
/* scheduling options of the production kernel (none changes the image): "lpt" (heaviest tiles first, default 1),
 * "cull" (hidden environment only: pixels outside the screen rectangle of the volume's box and 8x4 tiles onto which no
 * brick with a positive majorant projects are exactly zero and are not traced, default 1), "pass" (samples per pixel and
 * internal pass, default 32), "count_culled" (default 0: the counting build traces every sample, i.e. counts the events of
 * the reference algorithm; 1: it keeps the culling and counts the events the production launch executes).
 * Environment variables VRB200_LPT / VRB200_CULL / VRB200_PASS set the defaults.
 * "async_upload" (default 0) changes the ownership rule of vrb_grid_upload_brick / vrb_env_upload / vrb_tf_upload: with 1
 * they only enqueue (cudaMemcpyAsync semantics) -- the host buffers must stay valid and unchanged until the next vrb_sync
 * or download on this context; pass pinned memory so that the copies really are asynchronous. A pipelined caller can
 * then enqueue the uploads and the trace of the next frame while the previous one is still running.
 * "overlap" (default 1): the passes of one vrb_trace call alternate between two streams so that pass k + 1 starts in the tail
 * of pass k. "l2_persist" (default 0, VRB200_L2_PERSIST): MiB of L2 set aside for persisting lines, with an access-policy
 * window over the grid's brick records + majorant tables (what every DDA step fetches first); measured flat on B200. */
int vrb_set_option(vrb_ctx* ctx, const char* name, int value);
/* measurement aid (new; SURVEY 8(d): the L2 roofline "must be micro-benchmarked on the box"): read bandwidth of this device over
 * a working set of `bytes` (1 MiB ... 8 GiB; below the 126 MB L2 it measures L2, 1 GiB measures HBM), best of 5 launches timed
 * with CUDA events on the context's stream. mode 0 = streaming 16-byte loads, mode 1 = random 32-byte sector gathers
 * (4-byte load per sector, the tracer's access pattern; GB/s counted at 32 B per sector). */
int vrb_probe_bandwidth(vrb_ctx* ctx, size_t bytes, int mode, double* gbytes_per_s);
/* counters of the context: "trace_launches" = hand-written kernels vrb_trace has launched so far on the production path
 * (majorant tables, brick mask, tile keys, tracking kernel, fold; library sorts and memsets are not counted) */
int vrb_get_stat(vrb_ctx* ctx, const char* name, uint64_t* out);
/* color *= s (finalise VRB_ACCUM_SUM buffers) */
int vrb_scale(vrb_ctx* ctx, float s);
/* zero the colour buffer */
int vrb_clear(vrb_ctx* ctx);
/* enable (1) / disable (0) the counting build of the kernels and reset the counters */
int vrb_set_counting(vrb_ctx* ctx, int enable);
int vrb_get_counters(vrb_ctx* ctx, vrb_counters* out);

/* ---- tonemap + readback (shader/tonemap.glsl, tonemap.fs, blit.fs; bindings.cpp:141-166) --------- */
/* in_place != 0: shader/tonemap.glsl on `color` (src/main.cpp:540-550);
 * in_place == 0: RendererOpenGL::draw() into the RGBA8 framebuffer (renderer.cpp:147-153):
 *                tonemapping != 0 -> tonemap.fs, else blit.fs */
int vrb_tonemap(vrb_ctx* ctx, float exposure, float gamma, int in_place, int tonemapping);
/* glGetTexImage(color, GL_RGB/GL_RGBA, GL_FLOAT) (bindings.cpp:141-148); channels = 3 or 4 */
int vrb_download_color(vrb_ctx* ctx, float* out, int channels);
/* Texture2DImpl::save_ldr readback: color -> unorm8 RGBA (cppgl texture.cpp:107-113) */
int vrb_download_color_ldr(vrb_ctx* ctx, uint8_t* rgba8);
/* glReadPixels of the framebuffer written by the last draw (bindings.cpp:149-166) */
int vrb_download_framebuffer(vrb_ctx* ctx, uint8_t* rgba8);
/* overwrite the colour buffer from host memory (resume / tile assembly); rgba: w*h*4 floats */
int vrb_upload_color(vrb_ctx* ctx, const float* rgba);
/* device pointer of the float4 colour buffer (for NCCL reductions issued by the host runtime) */
void* vrb_color_device_ptr(vrb_ctx* ctx);
/* render into a caller-owned device buffer (w*h float4, e.g. a torch tensor that NCCL reduces) instead of
 * the context's own; the caller keeps ownership and must outlive the binding (undone by vrb_resize) */
int vrb_bind_color(vrb_ctx* ctx, void* device_rgba32f);

/* ---- multi-GPU (new) --------------------------------------------------------------- */
/* single-process helper: sum the colour buffers of n contexts (one per device) into ctxs[root]
 * over peer-to-peer NVLink copies + an add kernel. One-process-per-GPU runs use NCCL on
 * vrb_color_device_ptr instead (bench.py). */
int vrb_reduce(vrb_ctx* const* ctxs, int n, int root);
/* tile-partitioned rendering: copy image rows [y0, y1) of src's colour buffer into dst's (peer-to-peer when the
 * contexts live on different devices); both must have the same resolution */
int vrb_copy_rows(vrb_ctx* dst, vrb_ctx* src, int y0, int y1);

#ifdef __cplusplus
}
#endif
#endif /* VRB200_H */
