// libvrb200.so -- C ABI (include/vrb200.h) over the sm_100a kernels. No CPU fallback: every entry
// point either runs on the context's CUDA device or returns an error status.
#include "../../include/vrb200.h"

#include "vr_common.cuh"
#include "vr_brick.cuh"
#include "vr_nvdb.cuh"
#include "vr_env.cuh"
#include "vr_trace.cuh"
#include "vr_trace2.cuh"
#include "vr_trace_pool.cuh"
#include "vr_probe.cuh"

#include <nvtx3/nvToolsExt.h>      // header-only NVTX 3: ranges show up in Nsight Systems / ncu --nvtx, no-ops otherwise
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace vr;

namespace vr {   // vrb200_strict.cu (-fmad=false): the IEEE cross-check kernels
cudaError_t launch_trace_pixels(const TraceArgs& a, bool tf, bool count, dim3 grid, cudaStream_t stream);
const void* strict_persistent_kernel(bool tf, bool count);
cudaError_t launch_majorant_table_strict(const TraceArgs& a, bool tf, int level, float* out, size_t n, int blocks, cudaStream_t stream);
}

// ------------------------------------------------------------------------------------------------
// context

namespace {

// NVTX range around a host-side phase (upload, brick build, trace pass, fold, reduce, read-back): stands in for the
// reference's TimerQueryGL("trace") (src/main.cpp:479,509-511)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

struct DeviceGrid {
    bool valid = false;
    uint3 nb = { 0, 0, 0 };
    uint3 atlas_dim = { 0, 0, 0 };
    uint64_t brick_count = 0;
    // canonical voldata buffers (download / bit-exact contract)
    uint32_t* indirection = nullptr;
    uint32_t* range = nullptr;
    uint8_t* atlas = nullptr;
    uint32_t* mips[3] = { nullptr, nullptr, nullptr };
    // tracer layout. `hot` is ONE allocation holding the structures every DDA step / tentative collision fetches first -- the records
    // and the four majorant tables -- so that a single L2 access-policy window (option "l2_persist") can cover them
    void* hot = nullptr;
    size_t hot_bytes = 0;
    uint2* rec = nullptr;           // = hot
    uint2* recp = nullptr;       // padded (nb + 2)^3 records for the trilinear fetch (ensure_padded_records)
    bool recp_valid = false;
    uint8_t* atlas_lin = nullptr;   // n_slots bricks + one all-zero brick
    size_t n_slots = 0;
    bool decoded_valid = false;     // cslot / datlas describe the current contents
    uint32_t* cslot = nullptr;      // (nb + 1)^3 cells -> decoded apron block (0 = shared zero block)
    float* datlas = nullptr;        // (n_dblocks + 1) x 729 decoded voxels
    size_t n_dblocks = 0;
    // majorant tables of the persistent kernel (inside `hot`, behind the records), valid for maj_key
    float* maj[4] = { nullptr, nullptr, nullptr, nullptr };
    uint64_t maj_key = 0;
    uint64_t version = 0;           // bumped whenever the contents change (keys the cached tile order / brick mask)
};

struct Frame {
    DeviceGrid slot[2];
};

}  // namespace

struct vrb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    int w = 0, h = 0;
    float4* color = nullptr;
    bool color_external = false;
    uchar4* fb = nullptr;
    uchar4* ldr = nullptr;
    std::map<int, Frame> frames;
    float4* env_rgb = nullptr;
    float* env_stage = nullptr;      // RGB staging of vrb_env_upload (kept while the size stays the same)
    int env_w = 0, env_h = 0;
    float* impmap = nullptr;
    float4* env_split = nullptr;   // split tables of sample_environment (k_env_split), rebuilt with the pyramid
    float4* lut = nullptr;
    uint32_t tf_size = 0;
    unsigned long long* counters = nullptr;
    unsigned int* job_counter = nullptr;     // two block tickets: consecutive passes of a vrb_trace call alternate between two lanes
    float4* lbuf = nullptr;      // per-launch sample buffer of the persistent kernel: lbuf_samples x (w * h) float4
    float4* lbuf2 = nullptr;     // the second lane's (allocated when a call has more than one pass)
    int lbuf_samples = 0, lbuf2_samples = 0;
    cudaStream_t pass_stream = nullptr;      // second lane: pass k + 1 starts in the tail of pass k (persistent kernels fill the SMs one wave deep)
    cudaEvent_t ev_entry = nullptr, ev_fold[2] = { nullptr, nullptr }, ev_order = nullptr;   // ev_order: the latest rebuild of the brick mask / tile order
    int overlap = 1;             // VRB200_OVERLAP / option "overlap": 0 = every pass on the context's stream
    // L2 access-policy window over the grid's `hot` allocation (records + majorant tables): option "l2_persist" = MiB of L2 set
    // aside for persisting lines (0 = off, the default: measured flat on C3 / C4, profiles/r02_l2_persist.txt)
    int l2_persist_mb = 0;
    const void* l2_window_ptr = nullptr;      // what the streams' attribute currently covers
    size_t l2_window_bytes = 0;
    int pass_samples = 32;       // samples per pixel and pass (VRB200_PASS); bounds lbuf (also capped at 1 GiB = 32 samples at 1080p).
                                 // B200, configs[1]: 16 -> 33.6, 32 -> 36.9, 64 -> 36.7, 128 -> 36.5 Gsamples/s (fixed costs + tail per pass)
    // heaviest-tiles-first scheduling (vr_trace2.cuh): per-tile cost of the last launch, its view key, the sorted order
    unsigned int* tile_cost = nullptr;
    unsigned int* tile_cost_sorted = nullptr;
    uint32_t* tile_iota = nullptr;
    uint32_t* tile_order = nullptr;
    unsigned int* tile_live = nullptr;   // k_tile_mask: 1 = some non-empty brick projects onto the tile
    unsigned int* tile_key = nullptr;    // sort keys (k_tile_keys)
    unsigned int* live_info = nullptr;   // {number of live tiles, mask unusable}
    bool lut_monotone = false;           // alpha of the uploaded LUT is non-decreasing (the TF majorant bounds the density then)
    void* sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    int tile_capacity = 0;
    uint64_t cost_key = 0;       // view the costs in tile_cost belong to (0 = none)
    uint64_t order_key = 0;      // (view, grid contents, mask / cost mode) tile_order + tile_live + live_info were built for (0 = none)
    int order_age = 0;           // passes traced with that order; it is rebuilt from the latest costs every ORDER_REUSE passes
    bool lpt = true;             // VRB200_LPT=0 disables
    bool cull = true;            // VRB200_CULL=0 disables the screen-space box culling
    uint64_t trace_launches = 0; // hand-written kernels launched by vrb_trace so far (vrb_get_stat "trace_launches")
    bool async_upload = false;   // option "async_upload": upload calls return without waiting; the caller keeps the host buffers alive until vrb_sync / a download
    bool count_culled = false;   // option "count_culled": the counting build keeps the culling (events of the production launch, not of the reference algorithm)
    int tile_coords_tx = 0;      // tiles_x the packed coordinates in tile_iota were made for
    bool counting = false;
    int kernel = 0;            // 0 = persistent FastMath (production), 1 = simple strict cross-check, 2 = persistent StrictMath
    int trace_blocks[16] = { 0 };
    uint64_t lut_version = 0;
    std::string err;
};

namespace {

int fail(vrb_ctx* c, int status, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return status;
}

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? VRB_ERR_OOM : VRB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                          \
    } while (0)

#define CK_LAUNCH() CK(cudaGetLastError())

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Grid storage and build scratch come from the device's stream-ordered pool (release threshold raised in vrb_create):
// after the first build of a given size an allocation or free costs microseconds instead of the 0.1-1 ms of a
// cudaMalloc / cudaFree of hundreds of MiB (the brick build of a 1024^3 grid spent 26 of its 29 ms there).
template <typename T> cudaError_t pool_alloc(T** p, size_t bytes, cudaStream_t s) { return cudaMallocAsync(reinterpret_cast<void**>(p), bytes ? bytes : 16, s); }
inline void pool_free(void* p, cudaStream_t s) { if (p) cudaFreeAsync(p, s); }

void free_grid(DeviceGrid& g, cudaStream_t s) {
    pool_free(g.indirection, s); pool_free(g.range, s); pool_free(g.atlas, s);
    for (auto& m : g.mips) pool_free(m, s);
    pool_free(g.hot, s); pool_free(g.recp, s); pool_free(g.atlas_lin, s);
    pool_free(g.cslot, s); pool_free(g.datlas, s);
    g = DeviceGrid();
}

inline int grid_for(size_t n, int block, int sm_count, int per_sm = 16) {
    const size_t need = (n + block - 1) / block;
    const size_t cap = size_t(sm_count) * per_sm;
    return int(need < 1 ? 1 : (need < cap ? need : cap));
}

size_t mip_words(const uint3& nb, int level) { return size_t(nb.x >> (level + 1)) * (nb.y >> (level + 1)) * (nb.z >> (level + 1)); }

// records + majorant tables (level 0 with one extra entry for the out-of-bounds majorant) in one allocation
cudaError_t alloc_hot(DeviceGrid& g, cudaStream_t s) {
    const size_t n = size_t(g.nb.x) * g.nb.y * g.nb.z;
    size_t words[4] = { n + 1, mip_words(g.nb, 0), mip_words(g.nb, 1), mip_words(g.nb, 2) };
    size_t bytes = n * sizeof(uint2);
    for (size_t w : words) bytes += ((w * 4 + 255) / 256) * 256;
    const cudaError_t e = pool_alloc(&g.hot, bytes, s);
    if (e != cudaSuccess) return e;
    g.hot_bytes = bytes;
    g.rec = static_cast<uint2*>(g.hot);
    char* p = static_cast<char*>(g.hot) + n * sizeof(uint2);
    for (int l = 0; l < 4; ++l) { g.maj[l] = reinterpret_cast<float*>(p); p += ((words[l] * 4 + 255) / 256) * 256; }
    g.maj_key = 0;
    return cudaSuccess;
}

// builds the tracer layout (records + brick-linear atlas) from the canonical buffers
int finalize_grid(vrb_ctx* ctx, DeviceGrid& g, bool reuse = false, bool lin_done = false) {     // lin_done: the builder wrote rec and atlas_lin itself
    NvtxRange nvtx_("vrb:finalize_grid (records, linear atlas, decoded blocks)");
    const size_t n = size_t(g.nb.x) * g.nb.y * g.nb.z;
    const uint3 ab = make_uint3(g.atlas_dim.x >> 3, g.atlas_dim.y >> 3, g.atlas_dim.z >> 3);
    g.n_slots = size_t(ab.x) * ab.y * ab.z;
    if (g.n_slots >= 0xffffffffull) return fail(ctx, VRB_ERR_INVALID, "atlas too large");
    if (!reuse) {
        CK(alloc_hot(g, ctx->stream));
        CK(pool_alloc(&g.atlas_lin, (g.n_slots + 1) * 512, ctx->stream));
    }
    g.maj_key = 0;   // the majorant tables (if any) belong to the previous contents
    static uint64_t grid_versions = 0;
    g.version = ++grid_versions;
    if (!lin_done) CK(cudaMemsetAsync(g.atlas_lin + g.n_slots * 512, 0, 512, ctx->stream));   // the all-zero brick
    if (!lin_done) {
        k_make_records<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(g.indirection, g.range, n, ab, g.rec);
        CK_LAUNCH();
    }
    g.recp_valid = false;        // padded records (8-tap fetch through the u8 atlas: IEEE cross-check kernels, debug sampler): built on first use
    if (g.n_slots && !lin_done) {
        k_linearize_atlas<<<grid_for(g.n_slots * 64, 256, ctx->sm_count), 256, 0, ctx->stream>>>(g.atlas, g.atlas_dim, g.atlas_lin, g.n_slots);
        CK_LAUNCH();
    }
    g.decoded_valid = false;     // the decoded apron blocks (TF kernels only) are rebuilt on first use: ensure_decoded_blocks
    g.valid = true;
    return VRB_OK;
}

// Decoded apron blocks for the production trilinear fetch (vr_trace.cuh density_trilinear_decoded): only the TF kernels read
// them, and they are 11x the atlas (1024^3 fBm: 0.87 GB written, 0.93 ms), so they are built on the first TF trace of a
// grid (and by the debug sampler) instead of on every upload / build.
int ensure_padded_records(vrb_ctx* ctx, DeviceGrid& g) {
    if (g.recp_valid) return VRB_OK;
    const size_t np = size_t(g.nb.x + 2) * (g.nb.y + 2) * (g.nb.z + 2);
    if (!g.recp) CK(pool_alloc(&g.recp, np * sizeof(uint2), ctx->stream));
    k_make_records_padded<<<grid_for(np, 256, ctx->sm_count), 256, 0, ctx->stream>>>(g.rec, g.nb, uint32_t(g.n_slots), g.recp);
    CK_LAUNCH();
    g.recp_valid = true;
    return VRB_OK;
}

int ensure_decoded_blocks(vrb_ctx* ctx, DeviceGrid& g) {
    if (g.decoded_valid) return VRB_OK;
    NvtxRange nvtx_("vrb:decoded_blocks");
    const size_t nc = size_t(g.nb.x + 1) * (g.nb.y + 1) * (g.nb.z + 1);
    uint32_t *flags = nullptr, *excl = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    if (!g.cslot) CK(pool_alloc(&g.cslot, nc * 4, ctx->stream));
    CK(pool_alloc(&flags, nc * 4, ctx->stream));
    CK(pool_alloc(&excl, (nc + 1) * 4, ctx->stream));
    k_cell_flags<<<grid_for(nc, 256, ctx->sm_count), 256, 0, ctx->stream>>>(g.rec, g.nb, flags);
    CK_LAUNCH();
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flags, excl, nc, ctx->stream));
    CK(pool_alloc(&tmp, tmp_bytes, ctx->stream));
    CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flags, excl, nc, ctx->stream));
    k_cell_slots<<<grid_for(nc, 256, ctx->sm_count), 256, 0, ctx->stream>>>(flags, excl, nc, g.cslot);
    CK_LAUNCH();
    size_t n_blocks = nc;           // small lattices (<= 64 MiB of blocks): worst case, no read-back and no host sync
    if (nc * DBRICK * 4 > (size_t(64) << 20)) {
        uint32_t last[2] = { 0, 0 };    // exclusive sum and flag of the last cell -> number of blocks
        CK(cudaMemcpyAsync(&last[0], excl + (nc - 1), 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(&last[1], flags + (nc - 1), 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        n_blocks = size_t(last[0]) + last[1];
    }
    if (!g.datlas || n_blocks > g.n_dblocks) {
        pool_free(g.datlas, ctx->stream);
        g.datlas = nullptr;
        CK(pool_alloc(&g.datlas, (n_blocks + 1) * DBRICK * 4, ctx->stream));
        g.n_dblocks = n_blocks;
    }
    CK(cudaMemsetAsync(g.datlas, 0, DBRICK * 4, ctx->stream));      // block 0: zeros
    if (n_blocks) {
        GridView v;
        memset(&v, 0, sizeof v);
        v.nb = g.nb; v.rec = g.rec; v.atlas_lin = g.atlas_lin;
        k_decode_cells<<<grid_for(nc * 32, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(v, g.cslot, g.datlas);
        CK_LAUNCH();
    }
    pool_free(flags, ctx->stream); pool_free(excl, ctx->stream); pool_free(tmp, ctx->stream);
    g.decoded_valid = true;
    return VRB_OK;
}

int check_slot_frame(vrb_ctx* ctx, int slot, int frame) {
    if (!ctx) return VRB_ERR_INVALID;
    if (slot != VRB_SLOT_DENSITY && slot != VRB_SLOT_EMISSION) return fail(ctx, VRB_ERR_INVALID, "bad grid slot %d", slot);
    if (frame < 0) return fail(ctx, VRB_ERR_INVALID, "bad frame %d", frame);
    return VRB_OK;
}

int compute_n_bricks(const uint32_t dim[3], uint3& nb) {
    uint32_t out[3];
    for (int a = 0; a < 3; ++a) {
        // div_round_up goes through float (grid_brick.cpp:54-56); "* 1u << 3" parses as (x * 1u) << 3 (:62)
        const uint32_t b = uint32_t(std::ceil(float(dim[a]) / 8.f));
        const uint32_t c = uint32_t(std::ceil(float(b) / 8.f));
        out[a] = (c * 1u) << 3;
        if (out[a] >= 1024u) return VRB_ERR_TOO_MANY_BRICKS;
    }
    nb = make_uint3(out[0], out[1], out[2]);
    return VRB_OK;
}

// d_values != nullptr: any-Grid source, lookup() values on the padded lattice [-2, 8 nb + 2)^3 (vr_brick.cuh); else the
// u8 voxels of a DenseGrid
int build_from_device_voxels(vrb_ctx* ctx, int slot, int frame, const uint8_t* d_vox, const uint32_t dim[3], float vmin, float vmax,
                             const float* d_values = nullptr) {
    NvtxRange nvtx_("vrb:brick_build");
    uint3 nb;
    if (compute_n_bricks(dim, nb) != VRB_OK)
        return fail(ctx, VRB_ERR_TOO_MANY_BRICKS, "exceeded max brick count of 1024");
    DeviceGrid& g = ctx->frames[frame].slot[slot];
    free_grid(g, ctx->stream);
    g.nb = nb;
    const size_t n = size_t(nb.x) * nb.y * nb.z;
    const uint3 vdim = make_uint3(dim[0], dim[1], dim[2]);
    uint32_t *flags = nullptr, *brick_id = nullptr, *block_sums = nullptr, *brick_of_id = nullptr;
    unsigned long long* d_total = nullptr;
    const int n_blocks = int((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    CK(pool_alloc(&g.indirection, n * 4, ctx->stream));
    CK(pool_alloc(&g.range, n * 4, ctx->stream));
    CK(pool_alloc(&flags, n * 4, ctx->stream));
    CK(pool_alloc(&brick_id, n * 4, ctx->stream));
    CK(pool_alloc(&brick_of_id, n * 4, ctx->stream));
    CK(pool_alloc(&block_sums, size_t(n_blocks) * 4, ctx->stream));
    CK(pool_alloc(&d_total, 8, ctx->stream));
    // A: ranges
    bool counted = false;          // block_sums already filled by the range pass
    if (d_values) {
        k_brick_range_values<<<grid_for(n * 32, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(d_values, nb, g.range, flags);
        CK_LAUNCH();
    } else if ((dim[0] & 7u) == 0 && (reinterpret_cast<uintptr_t>(d_vox) & 7u) == 0) {
        // separable reduction over the 12^3 windows: x and y fused in one streaming pass over the voxels, then z
        uint16_t* m2 = nullptr;
        const size_t n2 = size_t(dim[2]) * nb.y * nb.x;
        CK(pool_alloc(&m2, n2 * 2, ctx->stream));
        const int vec16 = ((dim[0] & 15u) == 0 && (reinterpret_cast<uintptr_t>(d_vox) & 15u) == 0) ? 1 : 0;
        const size_t items = size_t(dim[2]) * (nb.x >> 1);       // (z-slice, pair of brick columns): < 2^31 / RANGE_XY_THREADS blocks for any grid below the 1024-brick limit
        k_range_xy<<<dim3(unsigned((items + RANGE_XY_THREADS - 1) / RANGE_XY_THREADS), (nb.y + RANGE_BAND_BY - 1) / RANGE_BAND_BY), RANGE_XY_THREADS, 0, ctx->stream>>>(d_vox, vdim, nb, m2, vec16);
        CK_LAUNCH();
        k_range_z_count<<<n_blocks, SCAN_BLOCK / 2, 0, ctx->stream>>>(m2, vdim, vmin, vmax, nb, g.range, flags, block_sums);
        CK_LAUNCH();
        pool_free(m2, ctx->stream);
        counted = true;
    } else {
        k_brick_range<<<grid_for(n * 32, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(d_vox, vdim, vmin, vmax, nb, g.range, flags);
        CK_LAUNCH();
    }
    // D: range mips (they depend on the ranges only: issued here so that they run while the host waits for the brick count below)
    for (int i = 0; i < 3; ++i) CK(pool_alloc(&g.mips[i], mip_words(nb, i) * 4, ctx->stream));
    k_range_mips3<<<unsigned(mip_words(nb, 2)), 64, 0, ctx->stream>>>(g.range, nb, g.mips[0], g.mips[1], g.mips[2]);     // one CTA per 8^3-brick region
    CK_LAUNCH();
    // B: ordered allocation
    if (!counted) {
        k_scan_block_sums<<<n_blocks, SCAN_BLOCK, 0, ctx->stream>>>(flags, n, block_sums);
        CK_LAUNCH();
    }
    k_scan_sums<<<1, 1024, 0, ctx->stream>>>(block_sums, n_blocks, d_total);
    CK_LAUNCH();
    const bool aligned8 = !d_values && (dim[0] & 7u) == 0 && (reinterpret_cast<uintptr_t>(d_vox) & 7u) == 0;
    if (aligned8) CK(alloc_hot(g, ctx->stream));      // fast path: the scan writes the tracer's records as well
    k_scan_assign<<<n_blocks, SCAN_BLOCK, 0, ctx->stream>>>(flags, n, block_sums, nb, g.indirection, brick_id, brick_of_id, g.range, d_total, aligned8 ? g.rec : nullptr);
    CK_LAUNCH();
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    g.brick_count = total;
    // atlas pruned in z: 8 * round(ceil(count / float(nbx * nby)))  (grid_brick.cpp:112; float arithmetic)
    const uint32_t az = 8u * uint32_t(std::round(std::ceil(float(total) / float(nb.x * nb.y))));
    g.atlas_dim = make_uint3(nb.x * 8, nb.y * 8, az);
    const size_t atlas_bytes = size_t(g.atlas_dim.x) * g.atlas_dim.y * g.atlas_dim.z;
    CK(pool_alloc(&g.atlas, atlas_bytes, ctx->stream));
    const bool enc_lut = aligned8;                 // table-driven encode that also writes the brick-linear tracer atlas
    if (enc_lut) {
        // every allocated brick is written whole: only the unused tail of the last brick layer needs zeros
        const size_t layer = size_t(g.atlas_dim.x) * g.atlas_dim.y * 8;
        if (atlas_bytes) CK(cudaMemsetAsync(g.atlas + atlas_bytes - layer, 0, layer, ctx->stream));
        const size_t n_slots = size_t(nb.x) * nb.y * (az >> 3);
        CK(pool_alloc(&g.atlas_lin, (n_slots + 1) * 512, ctx->stream));
        CK(cudaMemsetAsync(g.atlas_lin + size_t(total) * 512, 0, (n_slots + 1 - size_t(total)) * 512, ctx->stream));   // unused slots + the all-zero brick
        if (total) {
            k_brick_encode_lut<<<grid_for(size_t(total) * 32, ENC_WARPS * 32, ctx->sm_count, 8), ENC_WARPS * 32, 0, ctx->stream>>>(
                d_vox, vdim, vmin, vmax, nb, g.range, flags, g.indirection, brick_of_id, uint32_t(total), g.atlas, g.atlas_dim, g.atlas_lin);
            CK_LAUNCH();
        }
    } else if (atlas_bytes) CK(cudaMemsetAsync(g.atlas, 0, atlas_bytes, ctx->stream));
    // C: encode
    if (enc_lut) {
    } else if (total && d_values) {
        k_brick_encode_values<<<grid_for(n * 32, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(d_values, nb, g.range, brick_id, g.atlas, g.atlas_dim);
        CK_LAUNCH();
    } else if (total) {
        k_brick_encode<<<grid_for(n * 32, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(d_vox, vdim, vmin, vmax, nb, g.range, brick_id, g.atlas, g.atlas_dim,
                                                                                                 ((dim[0] & 7u) == 0 && (reinterpret_cast<uintptr_t>(d_vox) & 7u) == 0) ? 1 : 0);
        CK_LAUNCH();
    }

    const int st = finalize_grid(ctx, g, enc_lut, enc_lut);
    CK(cudaStreamSynchronize(ctx->stream));
    pool_free(flags, ctx->stream); pool_free(brick_id, ctx->stream); pool_free(brick_of_id, ctx->stream); pool_free(block_sums, ctx->stream); pool_free(d_total, ctx->stream);
    return st;
}

GridView make_view(const DeviceGrid& g) {
    GridView v;
    v.nb = g.nb;
    v.rec = g.rec;
    for (int i = 0; i < 3; ++i) v.mips[i] = g.mips[i];
    v.atlas_lin = g.atlas_lin;
    v.recp = g.recp;
    v.psx = g.nb.x + 2;
    v.psxy = (g.nb.x + 2) * (g.nb.y + 2);
    v.cslot = g.cslot;
    v.datlas = g.datlas;
    return v;
}

void mat4_mul(const float* a, const float* b, float* out) {
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r)
            out[c * 4 + r] = a[0 * 4 + r] * b[c * 4 + 0] + a[1 * 4 + r] * b[c * 4 + 1] + a[2 * 4 + r] * b[c * 4 + 2] + a[3 * 4 + r] * b[c * 4 + 3];
}

int fill_trace_args(vrb_ctx* ctx, const vrb_params* p, TraceArgs& a) {
    if (!p) return fail(ctx, VRB_ERR_INVALID, "params is NULL");
    if (!ctx->color || ctx->w <= 0) return fail(ctx, VRB_ERR_STATE, "no colour buffer: call vrb_resize first");
    if (p->resolution[0] != ctx->w || p->resolution[1] != ctx->h)
        return fail(ctx, VRB_ERR_INVALID, "params.resolution %dx%d != colour buffer %dx%d", p->resolution[0], p->resolution[1], ctx->w, ctx->h);
    auto it = ctx->frames.find(p->frame);
    if (it == ctx->frames.end() || !it->second.slot[VRB_SLOT_DENSITY].valid)
        return fail(ctx, VRB_ERR_STATE, "no density grid uploaded for frame %d", p->frame);
    if (!ctx->env_rgb) return fail(ctx, VRB_ERR_STATE, "no environment uploaded");
    if (p->use_transferfunc && (!ctx->lut || ctx->tf_size == 0)) return fail(ctx, VRB_ERR_STATE, "use_transferfunc set but no LUT uploaded");
    memset(&a, 0, sizeof(a));
    a.p = *p;
    if (p->use_transferfunc && (ctx->kernel == 0 || ctx->kernel == 3)) {      // the fast-math TF kernels sample the decoded apron blocks
        const int st = ensure_decoded_blocks(ctx, it->second.slot[VRB_SLOT_DENSITY]);
        if (st) return st;
    }
    if (p->use_transferfunc) {      // every TF kernel is compiled with the 8-tap fetch through the padded records (the IEEE cross-checks use it)
        const int st = ensure_padded_records(ctx, it->second.slot[VRB_SLOT_DENSITY]);
        if (st) return st;
    }
    a.density = make_view(it->second.slot[VRB_SLOT_DENSITY]);
    if (p->has_emission) {
        if (!it->second.slot[VRB_SLOT_EMISSION].valid) return fail(ctx, VRB_ERR_STATE, "has_emission set but no emission grid for frame %d", p->frame);
        a.emission = make_view(it->second.slot[VRB_SLOT_EMISSION]);
        mat4_mul(p->vol_emission_inv_transform, p->vol_density_transform, a.emis_from_density.m);
    }
    a.env.rgb = ctx->env_rgb;
    a.env.w = ctx->env_w;
    a.env.h = ctx->env_h;
    a.env.impmap = ctx->impmap;
    a.env.split = ctx->env_split;
    a.lut = ctx->lut;
    a.tf_size = ctx->tf_size;
    a.color = ctx->color;
    a.counters = ctx->counters;
    a.cam_z = -.5f / std::tan(.5f * PI_F * p->cam_fov / 180.f);   // common.glsl:78, fp32 like the shader
    return VRB_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI

extern "C" {

int vrb_abi_version(void) { return VRB_ABI_VERSION; }

const char* vrb_status_string(int status) {
    switch (status) {
        case VRB_OK: return "ok";
        case VRB_ERR_INVALID: return "invalid argument";
        case VRB_ERR_NO_DEVICE: return "no usable CUDA device";
        case VRB_ERR_CUDA: return "CUDA error";
        case VRB_ERR_OOM: return "out of device memory";
        case VRB_ERR_TOO_MANY_BRICKS: return "exceeded max brick count of 1024";
        case VRB_ERR_STATE: return "invalid state";
        default: return "unknown status";
    }
}

int vrb_create(int device, vrb_ctx** out) {
    if (!out) return VRB_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return VRB_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VRB_ERR_NO_DEVICE;
    if (prop.major != 10) {  // the library only carries sm_100a SASS
        fprintf(stderr, "vrb200: device %d is sm_%d%d, this build targets sm_100a only\n", device, prop.major, prop.minor);
        return VRB_ERR_NO_DEVICE;
    }
    vrb_ctx* ctx = new vrb_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    DeviceGuard guard(device);
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&ctx->counters, 7 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->pass_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_entry, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fold[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fold[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming) != cudaSuccess ||
        cudaMalloc(&ctx->job_counter, 2 * sizeof(unsigned int)) != cudaSuccess) {
        delete ctx;
        return VRB_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    {   // keep freed blocks in the stream-ordered pool instead of returning them to the driver at every sync
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t threshold = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
        }
    }
    if (const char* e = getenv("VRB200_LPT")) ctx->lpt = atoi(e) != 0;
    if (const char* e = getenv("VRB200_CULL")) ctx->cull = atoi(e) != 0;
    if (const char* e = getenv("VRB200_PASS")) ctx->pass_samples = std::max(1, atoi(e));
    if (const char* e = getenv("VRB200_OVERLAP")) ctx->overlap = atoi(e) != 0;
    if (const char* e = getenv("VRB200_L2_PERSIST")) ctx->l2_persist_mb = std::max(0, atoi(e));
    if (const char* e = getenv("VRB200_KERNEL")) ctx->kernel = std::min(3, std::max(0, atoi(e)));
    cudaMemsetAsync(ctx->counters, 0, 7 * sizeof(unsigned long long), ctx->stream);
    *out = ctx;
    return VRB_OK;
}

void vrb_destroy(vrb_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& f : ctx->frames) { free_grid(f.second.slot[0], ctx->stream); free_grid(f.second.slot[1], ctx->stream); }
    if (!ctx->color_external) cudaFree(ctx->color);
    cudaFree(ctx->tile_cost); cudaFree(ctx->tile_cost_sorted); cudaFree(ctx->tile_iota); cudaFree(ctx->tile_order); cudaFree(ctx->sort_tmp);
    cudaFree(ctx->tile_live); cudaFree(ctx->tile_key); cudaFree(ctx->live_info);
    cudaFree(ctx->lbuf); cudaFree(ctx->lbuf2); cudaFree(ctx->env_stage);
    if (ctx->pass_stream) cudaStreamDestroy(ctx->pass_stream);
    if (ctx->ev_entry) cudaEventDestroy(ctx->ev_entry);
    for (int i = 0; i < 2; ++i) if (ctx->ev_fold[i]) cudaEventDestroy(ctx->ev_fold[i]);
    if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
    cudaFree(ctx->fb); cudaFree(ctx->ldr); cudaFree(ctx->env_rgb); cudaFree(ctx->impmap); cudaFree(ctx->env_split); cudaFree(ctx->lut); cudaFree(ctx->counters); cudaFree(ctx->job_counter);
    cudaStreamSynchronize(ctx->stream);
    {   // hand the pooled grid memory back to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* vrb_last_error(vrb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int vrb_set_stream(vrb_ctx* ctx, void* cuda_stream, int external) {
    if (!ctx) return VRB_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = external ? cudaStream_t(cuda_stream) : ctx->own_stream;   // external + NULL = the legacy default stream
    ctx->l2_window_ptr = nullptr; ctx->l2_window_bytes = 0;                 // the L2 access-policy window is a per-stream attribute: re-applied by the next trace
    return VRB_OK;
}

int vrb_sync(vrb_ctx* ctx) {
    if (!ctx) return VRB_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    return VRB_OK;
}

int vrb_resize(vrb_ctx* ctx, int w, int h) {
    if (!ctx) return VRB_ERR_INVALID;
    if (w <= 0 || h <= 0) return fail(ctx, VRB_ERR_INVALID, "bad resolution %dx%d", w, h);
    DeviceGuard guard(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    if (!ctx->color_external) cudaFree(ctx->color);
    cudaFree(ctx->fb); cudaFree(ctx->ldr); cudaFree(ctx->lbuf); cudaFree(ctx->lbuf2);
    ctx->color = nullptr; ctx->fb = nullptr; ctx->ldr = nullptr; ctx->lbuf = nullptr; ctx->lbuf2 = nullptr; ctx->lbuf_samples = ctx->lbuf2_samples = 0; ctx->color_external = false;
    ctx->w = w; ctx->h = h;
    const size_t n = size_t(w) * h;
    CK(cudaMalloc(&ctx->color, n * sizeof(float4)));
    CK(cudaMalloc(&ctx->fb, n * sizeof(uchar4)));
    CK(cudaMalloc(&ctx->ldr, n * sizeof(uchar4)));
    CK(cudaMemsetAsync(ctx->color, 0, n * sizeof(float4), ctx->stream));
    CK(cudaMemsetAsync(ctx->fb, 0, n * sizeof(uchar4), ctx->stream));
    return VRB_OK;
}

int vrb_bind_color(vrb_ctx* ctx, void* device_rgba32f) {
    if (!ctx || ctx->w <= 0) return ctx ? fail(ctx, VRB_ERR_STATE, "vrb_resize first") : VRB_ERR_INVALID;
    if (!device_rgba32f) return fail(ctx, VRB_ERR_INVALID, "NULL colour buffer");
    DeviceGuard guard(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    if (!ctx->color_external) cudaFree(ctx->color);
    ctx->color = reinterpret_cast<float4*>(device_rgba32f);
    ctx->color_external = true;
    return VRB_OK;
}

int vrb_grid_clear(vrb_ctx* ctx) {
    if (!ctx) return VRB_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto& f : ctx->frames) { free_grid(f.second.slot[0], ctx->stream); free_grid(f.second.slot[1], ctx->stream); }
    ctx->frames.clear();
    return VRB_OK;
}

int vrb_grid_free(vrb_ctx* ctx, int slot, int frame) {
    int st = check_slot_frame(ctx, slot, frame);
    if (st) return st;
    auto it = ctx->frames.find(frame);
    if (it == ctx->frames.end()) return VRB_OK;
    DeviceGuard guard(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    free_grid(it->second.slot[slot], ctx->stream);
    if (!it->second.slot[0].valid && !it->second.slot[1].valid) ctx->frames.erase(it);
    return VRB_OK;
}

int vrb_grid_upload_brick(vrb_ctx* ctx, int slot, int frame, const vrb_brick_view* v) {
    NvtxRange nvtx_("vrb:grid_upload_brick");
    int st = check_slot_frame(ctx, slot, frame);
    if (st) return st;
    if (!v || !v->indirection || !v->range || !v->range_mips[0] || !v->range_mips[1] || !v->range_mips[2])
        return fail(ctx, VRB_ERR_INVALID, "brick view has NULL buffers");
    for (int a = 0; a < 3; ++a) {
        if (v->n_bricks[a] == 0 || v->n_bricks[a] >= 1024u || (v->n_bricks[a] & 7u)) return fail(ctx, VRB_ERR_INVALID, "n_bricks[%d] = %u must be a multiple of 8 in [8, 1016]", a, v->n_bricks[a]);
        if (v->atlas_dim[a] & 7u) return fail(ctx, VRB_ERR_INVALID, "atlas_dim[%d] = %u must be a multiple of 8", a, v->atlas_dim[a]);
    }
    const size_t atlas_bytes = size_t(v->atlas_dim[0]) * v->atlas_dim[1] * v->atlas_dim[2];
    if (atlas_bytes && !v->atlas) return fail(ctx, VRB_ERR_INVALID, "atlas is NULL");
    DeviceGuard guard(ctx->device);
    DeviceGrid& g = ctx->frames[frame].slot[slot];
    // same shape as what the slot already holds (re-upload of an animation frame, progressive edits): keep the allocations
    const bool reuse = g.valid && g.nb.x == v->n_bricks[0] && g.nb.y == v->n_bricks[1] && g.nb.z == v->n_bricks[2] &&
                       g.atlas_dim.x == v->atlas_dim[0] && g.atlas_dim.y == v->atlas_dim[1] && g.atlas_dim.z == v->atlas_dim[2];
    if (!reuse) {
        CK(cudaStreamSynchronize(ctx->stream));
        free_grid(g, ctx->stream);
    }
    g.nb = make_uint3(v->n_bricks[0], v->n_bricks[1], v->n_bricks[2]);
    g.atlas_dim = make_uint3(v->atlas_dim[0], v->atlas_dim[1], v->atlas_dim[2]);
    g.brick_count = v->brick_count;
    const size_t n = size_t(g.nb.x) * g.nb.y * g.nb.z;
    if (!reuse) {
        CK(pool_alloc(&g.indirection, n * 4, ctx->stream));
        CK(pool_alloc(&g.range, n * 4, ctx->stream));
        CK(pool_alloc(&g.atlas, atlas_bytes, ctx->stream));
        for (int i = 0; i < 3; ++i) CK(pool_alloc(&g.mips[i], mip_words(g.nb, i) * 4, ctx->stream));
    }
    CK(cudaMemcpyAsync(g.indirection, v->indirection, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(g.range, v->range, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (atlas_bytes) CK(cudaMemcpyAsync(g.atlas, v->atlas, atlas_bytes, cudaMemcpyHostToDevice, ctx->stream));
    for (int i = 0; i < 3; ++i) CK(cudaMemcpyAsync(g.mips[i], v->range_mips[i], mip_words(g.nb, i) * 4, cudaMemcpyHostToDevice, ctx->stream));
    st = finalize_grid(ctx, g, reuse);
    if (!ctx->async_upload) CK(cudaStreamSynchronize(ctx->stream));   // host buffers are only borrowed for the duration of the call
    return st;
}

int vrb_grid_build_from_dense_device(vrb_ctx* ctx, int slot, int frame, const void* d_voxels_u8, const uint32_t dim[3], float vmin, float vmax) {
    int st = check_slot_frame(ctx, slot, frame);
    if (st) return st;
    if (!d_voxels_u8 || !dim || !dim[0] || !dim[1] || !dim[2]) return fail(ctx, VRB_ERR_INVALID, "bad dense grid");
    DeviceGuard guard(ctx->device);
    return build_from_device_voxels(ctx, slot, frame, static_cast<const uint8_t*>(d_voxels_u8), dim, vmin, vmax);
}

int vrb_grid_build_from_dense(vrb_ctx* ctx, int slot, int frame, const uint8_t* voxels_u8, const uint32_t dim[3], float vmin, float vmax) {
    int st = check_slot_frame(ctx, slot, frame);
    if (st) return st;
    if (!voxels_u8 || !dim || !dim[0] || !dim[1] || !dim[2]) return fail(ctx, VRB_ERR_INVALID, "bad dense grid");
    uint3 nb;
    if (compute_n_bricks(dim, nb) != VRB_OK) return fail(ctx, VRB_ERR_TOO_MANY_BRICKS, "exceeded max brick count of 1024");
    DeviceGuard guard(ctx->device);
    const size_t n = size_t(dim[0]) * dim[1] * dim[2];
    uint8_t* d_vox = nullptr;
    CK(pool_alloc(&d_vox, n, ctx->stream));
    cudaError_t e = cudaMemcpyAsync(d_vox, voxels_u8, n, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { pool_free(d_vox, ctx->stream); return fail(ctx, VRB_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e)); }
    st = build_from_device_voxels(ctx, slot, frame, d_vox, dim, vmin, vmax);
    pool_free(d_vox, ctx->stream);
    return st;
}

int vrb_brick_lattice(const uint32_t extent[3], uint32_t n_bricks[3], uint32_t padded_dim[3]) {
    if (!extent || !n_bricks) return VRB_ERR_INVALID;
    uint3 nb;
    if (compute_n_bricks(extent, nb) != VRB_OK) return VRB_ERR_TOO_MANY_BRICKS;
    n_bricks[0] = nb.x; n_bricks[1] = nb.y; n_bricks[2] = nb.z;
    if (padded_dim) { padded_dim[0] = nb.x * 8 + 4; padded_dim[1] = nb.y * 8 + 4; padded_dim[2] = nb.z * 8 + 4; }
    return VRB_OK;
}

int vrb_grid_build_from_values(vrb_ctx* ctx, int slot, int frame, const float* padded_values, const uint32_t extent[3]) {
    int st = check_slot_frame(ctx, slot, frame);
    if (st) return st;
    if (!padded_values || !extent) return fail(ctx, VRB_ERR_INVALID, "bad value lattice");
    if (!extent[0] || !extent[1] || !extent[2]) return fail(ctx, VRB_ERR_INVALID, "empty grid (index extent 0): nothing to build");
    uint3 nb;
    if (compute_n_bricks(extent, nb) != VRB_OK) return fail(ctx, VRB_ERR_TOO_MANY_BRICKS, "exceeded max brick count of 1024");
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t(nb.x) * 8 + 4) * (size_t(nb.y) * 8 + 4) * (size_t(nb.z) * 8 + 4);
    float* d_val = nullptr;
    CK(pool_alloc(&d_val, n * 4, ctx->stream));
    cudaError_t e = cudaMemcpyAsync(d_val, padded_values, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { pool_free(d_val, ctx->stream); return fail(ctx, VRB_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e)); }
    st = build_from_device_voxels(ctx, slot, frame, nullptr, extent, 0.f, 0.f, d_val);
    pool_free(d_val, ctx->stream);
    return st;
}

// ---- NanoVDB sources ------------------------------------------------------------------------------
namespace {

int nvdb_fail(char* err, size_t err_len, const char* fmt, ...) {
    if (err && err_len) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, err_len, fmt, ap);
        va_end(ap);
    }
    return VRB_ERR_INVALID;
}

// GridData::isValid (NanoVDB.h:1864-1875) on the first 672 bytes at p
bool nvdb_grid_header_valid(const uint8_t* p) {
    using namespace nvdb;
    const uint64_t magic = rd<uint64_t>(p + G_MAGIC);
    if (magic == MAGIC_GRID || rd<uint64_t>(p + G_DATA2) == MAGIC_GRID) return true;
    return magic == MAGIC_NUMB && (rd<uint32_t>(p + G_VERSION) >> 21) == ABI_MAJOR && rd<uint32_t>(p + G_COUNT) > 0u &&
           rd<uint32_t>(p + G_INDEX) < rd<uint32_t>(p + G_COUNT) && rd<uint32_t>(p + G_CLASS) < GRID_CLASS_END && rd<uint32_t>(p + G_TYPE) < GRID_TYPE_END;
}

uint64_t nvdb_string_hash(const char* s) {        // io::stringHash (io/IO.h:710-721)
    uint64_t hash = 0;
    for (const unsigned char* c = reinterpret_cast<const unsigned char*>(s); *c; ++c) {
        const uint64_t overflow = hash >> (64 - 8);
        hash *= 67;
        hash += *c + overflow;
    }
    return hash;
}

// Every link the accessor follows must stay inside the buffer (the reference trusts the file; a truncated or corrupt
// file must not crash the host or the device accessor here). O(nodes).
bool nvdb_links_valid(const uint8_t* g, uint64_t size) {
    using namespace nvdb;
    if (size < GRID_SIZE + TREE_SIZE + ROOT_SIZE) return false;
    const uint8_t* tree = g + TREE;
    const int64_t root_off = rd<int64_t>(tree + T_NODE_OFFSET + 24);
    if (root_off < int64_t(TREE_SIZE) || (root_off & 31) || uint64_t(root_off) + TREE + ROOT_SIZE > size) return false;
    const uint64_t root = TREE + uint64_t(root_off);
    const uint64_t n_tiles = rd<uint32_t>(g + root + R_TABLE_SIZE);
    if (root + ROOT_SIZE + n_tiles * TILE_SIZE > size) return false;
    auto inside = [&](uint64_t node, int64_t link, uint64_t node_size, uint64_t* child) {
        const int64_t at = int64_t(node) + link;
        if (link == 0 || at < 0 || (at & 31) || uint64_t(at) + node_size > size) return false;
        *child = uint64_t(at);
        return true;
    };
    for (uint64_t t = 0; t < n_tiles; ++t) {
        const int64_t up = rd<int64_t>(g + root + ROOT_SIZE + t * TILE_SIZE + RT_CHILD);
        if (up == 0) continue;
        uint64_t upper, lower, leaf;
        if (!inside(root, up, UPPER_SIZE, &upper)) return false;
        for (uint32_t w = 0; w < 512; ++w) {
            uint64_t bits = rd<uint64_t>(g + upper + UPPER_CHILD_MASK + w * 8);
            for (; bits; bits &= bits - 1) {
                const uint32_t nu = w * 64 + uint32_t(__builtin_ctzll(bits));
                if (!inside(upper, rd<int64_t>(g + upper + UPPER_TABLE + size_t(nu) * 8), LOWER_SIZE, &lower)) return false;
                for (uint32_t v = 0; v < 64; ++v) {
                    uint64_t lbits = rd<uint64_t>(g + lower + LOWER_CHILD_MASK + v * 8);
                    for (; lbits; lbits &= lbits - 1) {
                        const uint32_t nl = v * 64 + uint32_t(__builtin_ctzll(lbits));
                        if (!inside(lower, rd<int64_t>(g + lower + LOWER_TABLE + size_t(nl) * 8), LEAF_SIZE, &leaf)) return false;
                    }
                }
            }
        }
    }
    return true;
}

}  // namespace

int vrb_nvdb_open(const void* file, size_t bytes, const char* gridname, vrb_nvdb_info* out, char* err, size_t err_len) {
    using namespace nvdb;
    if (err && err_len) err[0] = 0;
    if (!file || !gridname || !out) return nvdb_fail(err, err_len, "bad argument");
    const uint8_t* f = static_cast<const uint8_t*>(file);
    uint64_t at = 0, found = UINT64_MAX;
    if (bytes >= GRID_SIZE && nvdb_grid_header_valid(f)) {
        // a raw grid buffer, possibly several grids back to back (GridHandle::read(is, gridName), GridHandle.h:405-426)
        uint32_t n = 0;
        const uint32_t count = rd<uint32_t>(f + G_COUNT);
        // all offsets are compared in the overflow-safe form `x > bytes - at` (at <= bytes holds throughout): sizes come from the file
        while (true) {
            if (at > bytes || GRID_SIZE > bytes - at) break;
            if (strncmp(reinterpret_cast<const char*>(f + at + G_NAME), gridname, G_NAME_LEN) == 0) { found = at; break; }
            if (n++ >= count) break;
            const uint64_t step = rd<uint64_t>(f + at + G_BYTES);
            if (step < GRID_SIZE || step > bytes - at) break;         // a grid is at least its header; never walk backwards / in place
            at += step;
        }
        if (found == UINT64_MAX) return nvdb_fail(err, err_len, "No raw grid named \"%s\"", gridname);
    } else {
        // segments: FileHeader (16) + gridCount x (FileMetaData (176) + name) + the grids (io/IO.h:386-431, :568-594)
        const uint64_t key = nvdb_string_hash(gridname);
        while (found == UINT64_MAX && at <= bytes && 16 <= bytes - at) {
            const uint64_t magic = rd<uint64_t>(f + at);
            if (magic != MAGIC_NUMB && magic != MAGIC_FILE)
                return nvdb_fail(err, err_len, "Expected a NanoVDB file, but read a file of unknown type!");
            if ((rd<uint32_t>(f + at + 8) >> 21) != ABI_MAJOR)
                return nvdb_fail(err, err_len, "An unrecoverable error in nanovdb::Segment::read:\n\tIncompatible file format: NanoVDB major version %u, expected %u",
                                 rd<uint32_t>(f + at + 8) >> 21, ABI_MAJOR);
            const uint32_t grid_count = rd<uint16_t>(f + at + 12), codec = rd<uint16_t>(f + at + 14);
            at += 16;
            uint64_t seek = 0, hit = UINT64_MAX;
            for (uint32_t i = 0; i < grid_count; ++i) {
                if (at > bytes || 176 > bytes - at) return nvdb_fail(err, err_len, "Failed reading FileGridMetaData");
                const uint64_t file_size = rd<uint64_t>(f + at + 8), name_key = rd<uint64_t>(f + at + 16);
                const uint32_t name_size = rd<uint32_t>(f + at + 136);
                if (name_size > bytes - at - 176) return nvdb_fail(err, err_len, "Failed reading FileGridMetaData");
                const std::string name(reinterpret_cast<const char*>(f + at + 176), strnlen(reinterpret_cast<const char*>(f + at + 176), name_size));
                if (hit == UINT64_MAX) {
                    if ((name_key == 0u || name_key == key) && name == gridname) hit = seek;
                    else {
                        if (file_size > bytes || seek > bytes - file_size) return nvdb_fail(err, err_len, "Failed reading FileGridMetaData");    // the skipped grids must fit in the file
                        seek += file_size;
                    }
                }
                at += 176 + name_size;
            }
            if (hit != UINT64_MAX) {
                if (codec == 1) return nvdb_fail(err, err_len, "ZIP compression codec was disabled during build");
                if (codec == 2) return nvdb_fail(err, err_len, "BLOSC compression codec was disabled during build");
                if (hit > bytes - at) return nvdb_fail(err, err_len, "Failed to read Tree from file");
                found = at + hit;
            } else {
                if (seek > bytes - at) break;
                at += seek;
            }
        }
        if (found == UINT64_MAX) return nvdb_fail(err, err_len, "Grid name '%s' not found in file", gridname);
    }
    if (found > bytes || GRID_SIZE + TREE_SIZE > bytes - found) return nvdb_fail(err, err_len, "Failed to read Tree from file");
    const uint8_t* g = f + found;
    const uint64_t grid_size = rd<uint64_t>(g + G_BYTES);
    if (grid_size > bytes - found || grid_size < GRID_SIZE + TREE_SIZE) return nvdb_fail(err, err_len, "Failed to read Tree from file");
    // handle.grid<float>() is null for other value types; !isValid(); !isFogVolume() (grid_nvdb.cpp:10-12)
    if (rd<uint32_t>(g + G_TYPE) != GRID_TYPE_FLOAT || !nvdb_grid_header_valid(g) || rd<uint32_t>(g + G_CLASS) != GRID_CLASS_FOG)
        return nvdb_fail(err, err_len, "Empty or invalid NanoVDB grid!");
    if (!nvdb_links_valid(g, grid_size)) return nvdb_fail(err, err_len, "Empty or invalid NanoVDB grid! (node offsets leave the buffer)");
    memset(out, 0, sizeof *out);
    out->grid_offset = found;
    out->grid_size = grid_size;
    out->active_voxels = rd<uint64_t>(g + TREE + T_VOXEL_COUNT);
    const uint8_t* root = g + TREE + rd<int64_t>(g + TREE + T_NODE_OFFSET + 24);
    const bool empty = rd<uint32_t>(root + R_TABLE_SIZE) == 0u;                  // GridData::isEmpty (NanoVDB.h:1992)
    float fmin[3] = { 0.f, 0.f, 0.f };
    for (int i = 0; i < 3; ++i) {
        const int32_t lo = rd<int32_t>(root + R_BBOX + 4 * i), hi = rd<int32_t>(root + R_BBOX + 12 + 4 * i);
        // both go through glm::vec3 (float) before the integer members take them (:14-15)
        fmin[i] = empty ? 0.f : float(lo);
        out->ibb_min[i] = int32_t(fmin[i]);
        out->extent[i] = empty ? 0u : uint32_t(float(hi - lo + 1));
    }
    out->minorant = rd<float>(root + R_MINIMUM);
    out->majorant = rd<float>(root + R_MAXIMUM);
    float* T = out->transform;                                                     // T[4 c + r]: glm column c, row r
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[4 * i + j] = rd<float>(g + G_MAP_MATF + 4 * (i * 3 + j));
        T[12 + i] = rd<float>(g + G_MAP_VECF + 4 * i);
    }
    T[15] = 1.f;
    // transform[3] += transform * vec4(ibb_min, 0), glm's mat4 * vec4: (m0 x + m1 y) + (m2 z + m3 w)  (type_mat4x4.inl:561-572)
    float add[4];
    for (int r = 0; r < 4; ++r) {
        const volatile float a = T[r] * fmin[0], b = T[4 + r] * fmin[1], c = T[8 + r] * fmin[2], d = T[12 + r] * 0.f;
        const volatile float ab = a + b, cd = c + d;
        add[r] = ab + cd;
    }
    for (int r = 0; r < 4; ++r) { const volatile float s = T[12 + r] + add[r]; T[12 + r] = s; }
    return VRB_OK;
}

int vrb_nvdb_lookup(const void* grid, const int32_t ibb_min[3], const uint32_t* ipos, size_t n, float* out) {
    if (!grid || !ibb_min || (n && (!ipos || !out))) return VRB_ERR_INVALID;
    const uint8_t* g = static_cast<const uint8_t*>(grid);
    for (size_t i = 0; i < n; ++i)
        out[i] = nvdb::get_value(g, int32_t(ipos[3 * i] + uint32_t(ibb_min[0])), int32_t(ipos[3 * i + 1] + uint32_t(ibb_min[1])), int32_t(ipos[3 * i + 2] + uint32_t(ibb_min[2])));
    return VRB_OK;
}

int vrb_grid_build_from_nvdb(vrb_ctx* ctx, int slot, int frame, const void* grid, const vrb_nvdb_info* info) {
    int st = check_slot_frame(ctx, slot, frame);
    if (st) return st;
    if (!grid || !info || info->grid_size < nvdb::GRID_SIZE + nvdb::TREE_SIZE + nvdb::ROOT_SIZE) return fail(ctx, VRB_ERR_INVALID, "bad NanoVDB grid");
    // an empty NanoVDB grid has extent 0; the reference's constructor divides by n_bricks.x * n_bricks.y there (grid_brick.cpp:112)
    if (!info->extent[0] || !info->extent[1] || !info->extent[2]) return fail(ctx, VRB_ERR_INVALID, "empty grid (index extent 0): nothing to build");
    uint3 nb;
    if (compute_n_bricks(info->extent, nb) != VRB_OK) return fail(ctx, VRB_ERR_TOO_MANY_BRICKS, "exceeded max brick count of 1024");
    DeviceGuard guard(ctx->device);
    const uint3 pd = make_uint3(nb.x * 8 + 4, nb.y * 8 + 4, nb.z * 8 + 4);
    const size_t n = size_t(pd.x) * pd.y * pd.z;
    uint8_t* d_grid = nullptr;
    float* d_val = nullptr;
    CK(pool_alloc(&d_grid, info->grid_size, ctx->stream));
    cudaError_t e = pool_alloc(&d_val, n * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_grid, grid, info->grid_size, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
        pool_free(d_grid, ctx->stream); pool_free(d_val, ctx->stream);
        return fail(ctx, e == cudaErrorMemoryAllocation ? VRB_ERR_OOM : VRB_ERR_CUDA, "NanoVDB upload failed: %s", cudaGetErrorString(e));
    }
    nvdb::k_nvdb_tabulate<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(d_grid, make_int3(info->ibb_min[0], info->ibb_min[1], info->ibb_min[2]), pd, d_val);
    e = cudaGetLastError();
    if (e == cudaSuccess) st = build_from_device_voxels(ctx, slot, frame, nullptr, info->extent, 0.f, 0.f, d_val);
    else st = fail(ctx, VRB_ERR_CUDA, "k_nvdb_tabulate launch failed: %s", cudaGetErrorString(e));
    pool_free(d_grid, ctx->stream);
    pool_free(d_val, ctx->stream);
    return st;
}

int vrb_grid_info(vrb_ctx* ctx, int slot, int frame, vrb_brick_view* out) {
    int st = check_slot_frame(ctx, slot, frame);
    if (st) return st;
    auto it = ctx->frames.find(frame);
    if (!out || it == ctx->frames.end() || !it->second.slot[slot].valid) return fail(ctx, VRB_ERR_STATE, "no grid in slot %d frame %d", slot, frame);
    const DeviceGrid& g = it->second.slot[slot];
    out->n_bricks[0] = g.nb.x; out->n_bricks[1] = g.nb.y; out->n_bricks[2] = g.nb.z;
    out->atlas_dim[0] = g.atlas_dim.x; out->atlas_dim[1] = g.atlas_dim.y; out->atlas_dim[2] = g.atlas_dim.z;
    out->brick_count = g.brick_count;
    return VRB_OK;
}

// density at index-space points through the tracer's own fetch functions (parity tests):
// mode 0 = lookup_density_trilinear via records + u8 atlas, 1 = the same via the decoded apron blocks, 2 = nearest voxel at floor(p)
__global__ void k_sample_density(const GridView g, const float* __restrict__ pts, size_t n, int mode, float* __restrict__ out) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const float3 p = f3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
        out[i] = mode == 0 ? density_trilinear(g, p) : mode == 1 ? density_trilinear_decoded(g, p) : brick_value(g, int(floorf(p.x)), int(floorf(p.y)), int(floorf(p.z)));
    }
}

int vrb_debug_sample_density(vrb_ctx* ctx, int slot, int frame, const float* ipos_xyz, size_t n, int mode, float* out) {
    int st = check_slot_frame(ctx, slot, frame);
    if (st) return st;
    auto it = ctx->frames.find(frame);
    if (it == ctx->frames.end() || !it->second.slot[slot].valid) return fail(ctx, VRB_ERR_STATE, "no grid in slot %d frame %d", slot, frame);
    if (!ipos_xyz || !out || mode < 0 || mode > 2) return fail(ctx, VRB_ERR_INVALID, "bad arguments");
    if (n == 0) return VRB_OK;
    DeviceGuard guard(ctx->device);
    if (mode == 1) { st = ensure_decoded_blocks(ctx, it->second.slot[slot]); if (st) return st; }
    if (mode == 0) { st = ensure_padded_records(ctx, it->second.slot[slot]); if (st) return st; }
    float *d_p = nullptr, *d_o = nullptr;
    CK(pool_alloc(&d_p, n * 12, ctx->stream));
    CK(pool_alloc(&d_o, n * 4, ctx->stream));
    CK(cudaMemcpyAsync(d_p, ipos_xyz, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    k_sample_density<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(make_view(it->second.slot[slot]), d_p, n, mode, d_o);
    CK_LAUNCH();
    CK(cudaMemcpyAsync(out, d_o, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    pool_free(d_p, ctx->stream); pool_free(d_o, ctx->stream);
    return VRB_OK;
}

int vrb_grid_download(vrb_ctx* ctx, int slot, int frame, vrb_brick_view* out) {
    int st = vrb_grid_info(ctx, slot, frame, out);
    if (st) return st;
    DeviceGuard guard(ctx->device);
    const DeviceGrid& g = ctx->frames[frame].slot[slot];
    const size_t n = size_t(g.nb.x) * g.nb.y * g.nb.z;
    const size_t atlas_bytes = size_t(g.atlas_dim.x) * g.atlas_dim.y * g.atlas_dim.z;
    if (out->indirection) CK(cudaMemcpyAsync(out->indirection, g.indirection, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->range) CK(cudaMemcpyAsync(out->range, g.range, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->atlas && atlas_bytes) CK(cudaMemcpyAsync(out->atlas, g.atlas, atlas_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    for (int i = 0; i < 3; ++i)
        if (out->range_mips[i]) CK(cudaMemcpyAsync(out->range_mips[i], g.mips[i], mip_words(g.nb, i) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VRB_OK;
}

static int dense_quantize_device(vrb_ctx* ctx, const float* d_data, size_t n, uint8_t* d_out, float* d_mm);

int vrb_dense_from_float(vrb_ctx* ctx, const float* data, const uint32_t dim[3], uint8_t* out_u8, float out_minmax[2]) {
    if (!ctx) return VRB_ERR_INVALID;
    if (!data || !dim || !out_u8 || !out_minmax || !dim[0] || !dim[1] || !dim[2]) return fail(ctx, VRB_ERR_INVALID, "bad arguments");
    DeviceGuard guard(ctx->device);
    const size_t n = size_t(dim[0]) * dim[1] * dim[2];
    float *d_data = nullptr, *d_mm = nullptr;
    uint8_t* d_out = nullptr;
    CK(cudaMalloc(&d_data, n * 4));
    CK(cudaMalloc(&d_out, n));
    CK(cudaMalloc(&d_mm, 8));
    CK(cudaMemcpyAsync(d_data, data, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    const int st = dense_quantize_device(ctx, d_data, n, d_out, d_mm);
    if (st == VRB_OK) {
        CK(cudaMemcpyAsync(out_u8, d_out, n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(out_minmax, d_mm, 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_data); cudaFree(d_out); cudaFree(d_mm);
    return st;
}

// min/max + quantisation of float voxels that are already on the device (both on ctx->stream); d_mm: two floats on the device
static int dense_quantize_device(vrb_ctx* ctx, const float* d_data, size_t n, uint8_t* d_out, float* d_mm) {
    float *d_bmin = nullptr, *d_bmax = nullptr;
    const bool vec4 = (n & 3u) == 0 && (reinterpret_cast<uintptr_t>(d_data) & 15u) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 3u) == 0;
    const int blocks = grid_for(vec4 ? n / 4 : n, 256, ctx->sm_count, 8);
    CK(pool_alloc(&d_bmin, size_t(blocks) * 4, ctx->stream));
    CK(pool_alloc(&d_bmax, size_t(blocks) * 4, ctx->stream));
    if (vec4) k_dense_minmax4<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const float4*>(d_data), n / 4, d_bmin, d_bmax);
    else k_dense_minmax<<<blocks, 256, 0, ctx->stream>>>(d_data, n, d_bmin, d_bmax);
    CK_LAUNCH();
    k_dense_minmax_final<<<1, 32, 0, ctx->stream>>>(d_bmin, d_bmax, blocks, d_mm);
    CK_LAUNCH();
    if (vec4) k_dense_quantize4<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const float4*>(d_data), n / 4, d_mm, reinterpret_cast<uint32_t*>(d_out));
    else k_dense_quantize<<<blocks, 256, 0, ctx->stream>>>(d_data, n, d_mm, d_out);
    CK_LAUNCH();
    pool_free(d_bmin, ctx->stream); pool_free(d_bmax, ctx->stream);
    return VRB_OK;
}

int vrb_grid_build_from_float_device(vrb_ctx* ctx, int slot, int frame, const void* d_voxels_f32, const uint32_t dim[3], float out_minmax[2]) {
    NvtxRange nvtx_("vrb:grid_build_from_float_device (DenseGrid(float*) + BrickGrid)");
    int st = check_slot_frame(ctx, slot, frame);
    if (st) return st;
    if (!d_voxels_f32 || !dim || !dim[0] || !dim[1] || !dim[2]) return fail(ctx, VRB_ERR_INVALID, "bad arguments");
    DeviceGuard guard(ctx->device);
    const size_t n = size_t(dim[0]) * dim[1] * dim[2];
    uint8_t* d_u8 = nullptr;
    float* d_mm = nullptr;
    CK(pool_alloc(&d_u8, n, ctx->stream));
    CK(pool_alloc(&d_mm, 8, ctx->stream));
    st = dense_quantize_device(ctx, static_cast<const float*>(d_voxels_f32), n, d_u8, d_mm);
    float mm[2] = { 0.f, 0.f };
    if (st == VRB_OK) {
        CK(cudaMemcpyAsync(mm, d_mm, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));      // DenseGrid::min_value / max_value parameterise the brick build
        if (out_minmax) { out_minmax[0] = mm[0]; out_minmax[1] = mm[1]; }
        st = build_from_device_voxels(ctx, slot, frame, d_u8, dim, mm[0], mm[1]);
    }
    pool_free(d_u8, ctx->stream); pool_free(d_mm, ctx->stream);
    return st;
}

int vrb_env_upload(vrb_ctx* ctx, const float* rgb, int w, int h) {
    NvtxRange nvtx_("vrb:env_upload (importance map + pyramid)");
    if (!ctx) return VRB_ERR_INVALID;
    if (!rgb || w <= 0 || h <= 0) return fail(ctx, VRB_ERR_INVALID, "bad environment map");
    DeviceGuard guard(ctx->device);
    if (!ctx->impmap) CK(cudaMalloc(&ctx->impmap, size_t(imp_offset(IMP_LEVELS)) * 4));
    if (!ctx->env_split) CK(cudaMalloc(&ctx->env_split, size_t(SPLIT_QUADS) * 3 * sizeof(float4)));
    const size_t n = size_t(w) * h;
    if (!ctx->env_rgb || ctx->env_w != w || ctx->env_h != h) {     // same size as before: keep the allocations
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->env_rgb); cudaFree(ctx->env_stage);
        ctx->env_rgb = nullptr; ctx->env_stage = nullptr;
        CK(cudaMalloc(&ctx->env_stage, n * 12));
        CK(cudaMalloc(&ctx->env_rgb, n * sizeof(float4)));
    }
    float* d_rgb = ctx->env_stage;
    CK(cudaMemcpyAsync(d_rgb, rgb, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    k_env_pad<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(d_rgb, ctx->env_rgb, n);
    CK_LAUNCH();
    ctx->env_w = w; ctx->env_h = h;
    EnvView e { ctx->env_rgb, w, h, ctx->impmap, nullptr };
    k_env_impmap<<<dim3(IMP_DIM / 16, IMP_DIM / 16), 256, 0, ctx->stream>>>(e, ctx->impmap);
    CK_LAUNCH();
    for (int l = 1; l < IMP_LEVELS; ++l) {
        const int d = IMP_DIM >> l;
        const dim3 block(16, 16), grid((d + 15) / 16, (d + 15) / 16);
        k_env_mip<<<grid, block, 0, ctx->stream>>>(ctx->impmap + imp_offset(l - 1), d * 2, ctx->impmap + imp_offset(l), d);
        CK_LAUNCH();
    }
    k_env_split<<<grid_for(SPLIT_QUADS, 256, ctx->sm_count), 256, 0, ctx->stream>>>(ctx->impmap, ctx->env_split);
    CK_LAUNCH();
    if (!ctx->async_upload) CK(cudaStreamSynchronize(ctx->stream));   // host buffer is only borrowed for the duration of the call
    return VRB_OK;
}

int vrb_env_download_impmap(vrb_ctx* ctx, int level, float* out) {
    if (!ctx) return VRB_ERR_INVALID;
    if (!ctx->impmap || !ctx->env_rgb) return fail(ctx, VRB_ERR_STATE, "no environment uploaded");
    if (level < 0 || level >= IMP_LEVELS || !out) return fail(ctx, VRB_ERR_INVALID, "bad level %d", level);
    DeviceGuard guard(ctx->device);
    const int d = IMP_DIM >> level;
    CK(cudaMemcpyAsync(out, ctx->impmap + imp_offset(level), size_t(d) * d * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VRB_OK;
}

int vrb_tf_upload(vrb_ctx* ctx, const float* rgba, uint32_t n) {
    NvtxRange nvtx_("vrb:tf_upload");
    if (!ctx) return VRB_ERR_INVALID;
    if (!rgba || n == 0) return fail(ctx, VRB_ERR_INVALID, "empty LUT");
    DeviceGuard guard(ctx->device);
    if (!ctx->lut || ctx->tf_size != n) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->lut); ctx->lut = nullptr;
        CK(cudaMalloc(&ctx->lut, size_t(n) * 16));
    }
    CK(cudaMemcpyAsync(ctx->lut, rgba, size_t(n) * 16, cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->async_upload) CK(cudaStreamSynchronize(ctx->stream));
    ctx->tf_size = n;
    ctx->lut_version++;
    // k_tile_mask relies on tf(hi).a bounding tf(d).a for d <= hi: true for the CDF-corrected LUTs of transferfunc.cpp:33-58,
    // checked here because the ABI accepts any table
    ctx->lut_monotone = true;
    for (uint32_t i = 0; i < n; ++i)
        if (!(rgba[4 * i + 3] >= (i ? rgba[4 * (i - 1) + 3] : 0.f))) { ctx->lut_monotone = false; break; }
    return VRB_OK;
}

int vrb_trace(vrb_ctx* ctx, const vrb_params* params, int first_sample, int n_samples, const int tile[4], int accum_mode) {
    NvtxRange nvtx_("vrb:trace");
    if (!ctx) return VRB_ERR_INVALID;
    if (first_sample < 1 || n_samples < 0) return fail(ctx, VRB_ERR_INVALID, "first_sample must be >= 1 (1-based current_sample)");
    if (accum_mode != VRB_ACCUM_MEAN && accum_mode != VRB_ACCUM_SUM) return fail(ctx, VRB_ERR_INVALID, "bad accum_mode");
    if (n_samples == 0) return VRB_OK;
    TraceArgs a;
    const int st = fill_trace_args(ctx, params, a);
    if (st) return st;
    a.x0 = tile ? tile[0] : 0; a.y0 = tile ? tile[1] : 0; a.x1 = tile ? tile[2] : ctx->w; a.y1 = tile ? tile[3] : ctx->h;
    if (a.x0 < 0 || a.y0 < 0 || a.x1 > ctx->w || a.y1 > ctx->h || a.x0 >= a.x1 || a.y0 >= a.y1) return fail(ctx, VRB_ERR_INVALID, "bad tile");
    a.first_sample = first_sample; a.n_samples = n_samples; a.accum_mode = accum_mode;
    DeviceGuard guard(ctx->device);
    const bool tf = params->use_transferfunc != 0;
    if (ctx->kernel == 1) {   // cross-check kernel: one thread per pixel, no regeneration
        const dim3 grid((a.x1 - a.x0 + 15) / 16, (a.y1 - a.y0 + 15) / 16);
        CK(launch_trace_pixels(a, tf, ctx->counting, grid, ctx->stream));
        return VRB_OK;
    }
    // ---- production path: persistent kernel ----
    DeviceGrid& g = ctx->frames[params->frame].slot[VRB_SLOT_DENSITY];
    // majorant tables depend on (grid, density_scale, global majorant, TF + window + LUT): rebuild when the key changes
    uint64_t key = 1469598103934665603ull;
    auto mix_key = [&key](const void* p, size_t n) { const unsigned char* b = static_cast<const unsigned char*>(p); for (size_t i = 0; i < n; ++i) { key ^= b[i]; key *= 1099511628211ull; } };
    mix_key(&params->vol_density_scale, 4); mix_key(&params->vol_majorant, 4); mix_key(&params->vol_inv_majorant, 4);
    mix_key(&params->use_transferfunc, 4);
    const int strict_tables = ctx->kernel == 2 ? 1 : 0;      // the IEEE cross-check kernel reads tables filled without FMA contraction
    mix_key(&strict_tables, 4);
    if (tf) { mix_key(&params->tf_window_left, 4); mix_key(&params->tf_window_width, 4); mix_key(&ctx->lut_version, 8); }
    const size_t n0 = size_t(g.nb.x) * g.nb.y * g.nb.z;
    if (!g.maj[0]) return fail(ctx, VRB_ERR_STATE, "grid without tracer layout");      // allocated with the records (alloc_hot)
    for (int l = 0; l < 4; ++l) a.maj[l] = g.maj[l];
    a.maj_oob = g.maj[0] + n0;
    for (int l = 0; l < 4; ++l) a.maj_off[l] = uint32_t(g.maj[l] - g.maj[0]);       // one allocation (alloc_hot): < 2^31 floats
    a.maj_off_oob = uint32_t(n0);
    if (g.maj_key != key) {
        for (int l = 0; l < 4; ++l) {
            const size_t n = l == 0 ? n0 : mip_words(g.nb, l - 1);
            const int blocks = grid_for(n, 256, ctx->sm_count);
            if (strict_tables) CK(launch_majorant_table_strict(a, tf, l, g.maj[l], n, blocks, ctx->stream));
            else if (tf) k_majorant_table<true><<<blocks, 256, 0, ctx->stream>>>(a, l, g.maj[l], n);
            else k_majorant_table<false><<<blocks, 256, 0, ctx->stream>>>(a, l, g.maj[l], n);
            CK_LAUNCH();
            ++ctx->trace_launches;
        }
        g.maj_key = key;
    }
    a.job_counter = ctx->job_counter;
    {
        // L2 policy: the records and majorant tables (g.hot) are the first, dependent fetch of every DDA step and tentative
        // collision; on the HBM-resident configs the atlas sectors streaming through L2 compete with them. One window per stream.
        const void* want_ptr = ctx->l2_persist_mb > 0 ? g.hot : nullptr;
        const size_t want_bytes = want_ptr ? g.hot_bytes : 0;
        if (want_ptr != ctx->l2_window_ptr || want_bytes != ctx->l2_window_bytes) {
            int max_window = 0, max_persist = 0;
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
            const size_t carve = std::min<size_t>(size_t(ctx->l2_persist_mb) << 20, size_t(max_persist));
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof attr);
            attr.accessPolicyWindow.base_ptr = const_cast<void*>(want_ptr);
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(want_bytes, size_t(max_window));
            attr.accessPolicyWindow.hitRatio = want_bytes ? float(std::min(1.0, double(carve) / double(std::max<size_t>(want_bytes, 1)))) : 0.f;
            attr.accessPolicyWindow.hitProp = want_ptr ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
            attr.accessPolicyWindow.missProp = want_ptr ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
            CK(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
            if (ctx->pass_stream) CK(cudaStreamSetAttribute(ctx->pass_stream, cudaStreamAttributeAccessPolicyWindow, &attr));
            if (!want_ptr) cudaCtxResetPersistingL2Cache();
            ctx->l2_window_ptr = want_ptr; ctx->l2_window_bytes = want_bytes;
        }
    }
    // ---- per-launch sample buffer: passes of at most `pass` samples per pixel (16 B per sample and pixel) ----
    const size_t n_px = size_t(ctx->w) * ctx->h;
    const int pass = int(std::max<size_t>(1, std::min<size_t>(size_t(ctx->pass_samples), (size_t(1) << 30) / (n_px * sizeof(float4)))));
    const int want_pass = std::min(pass, n_samples);
    if (ctx->lbuf_samples < want_pass) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->lbuf);
        ctx->lbuf = nullptr; ctx->lbuf_samples = 0;
        CK(cudaMalloc(&ctx->lbuf, n_px * sizeof(float4) * size_t(want_pass)));
        ctx->lbuf_samples = want_pass;
    }
    // two lanes: a call of several passes alternates them between the context's stream and a second one, each with its own
    // sample buffer and block ticket, so that pass k + 1 starts in the tail of pass k (a persistent kernel fills the SMs exactly one
    // wave deep: while its last CTAs finish, the next kernel's CTAs take the freed SMs). The folds stay in sample order.
    const bool two_lanes = ctx->overlap && !ctx->counting && n_samples > pass;
    if (two_lanes && ctx->lbuf2_samples < pass) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(ctx->pass_stream));
        cudaFree(ctx->lbuf2);
        ctx->lbuf2 = nullptr; ctx->lbuf2_samples = 0;
        CK(cudaMalloc(&ctx->lbuf2, n_px * sizeof(float4) * size_t(pass)));
        ctx->lbuf2_samples = pass;
    }
    a.lbuf = ctx->lbuf;
    a.lbuf_stride = n_px;
    // ---- screen-space culling of the volume's box (hidden environment only; never in the counting build) ----
    const int fold_x0 = a.x0, fold_y0 = a.y0, fold_x1 = a.x1, fold_y1 = a.y1;       // the caller's region
    bool box_cull_ok = false;
    if (!params->show_environment && (!ctx->counting || ctx->count_culled) && ctx->cull) {
        // view_dir (common.glsl:76-80): dir ~ cam_transform * (px, py, z), px = (x + jitter - w/2) / h, z = -0.5 / tan(fov/2).
        // A box corner c maps to v = cam_transform^-1 (c - cam_pos); in front of the camera (v.z < 0) it projects to
        // px = v.x * z / v.z. The bounding rectangle of the 8 projections (+ 1 pixel) contains every pixel that can hit.
        const float* T = params->cam_transform;   // column-major, orthonormal up to rounding: inverse = transpose
        const double z = -0.5 / std::tan(0.5 * M_PI * double(params->cam_fov) / 180.0), w = ctx->w, h = ctx->h;
        double xmin = 1e30, xmax = -1e30, ymin = 1e30, ymax = -1e30;
        bool ok = std::isfinite(z);
        for (int k = 0; k < 8 && ok; ++k) {
            const double c[3] = { double(k & 1 ? params->vol_bb_max[0] : params->vol_bb_min[0]) - params->cam_pos[0],
                                  double(k & 2 ? params->vol_bb_max[1] : params->vol_bb_min[1]) - params->cam_pos[1],
                                  double(k & 4 ? params->vol_bb_max[2] : params->vol_bb_min[2]) - params->cam_pos[2] };
            const double vx = T[0] * c[0] + T[1] * c[1] + T[2] * c[2], vy = T[3] * c[0] + T[4] * c[1] + T[5] * c[2], vz = T[6] * c[0] + T[7] * c[1] + T[8] * c[2];
            if (!(vz < -1e-4)) { ok = false; break; }       // a corner beside / behind the camera: no culling
            const double sx = vx * z / vz * h + 0.5 * w, sy = vy * z / vz * h + 0.5 * h;
            xmin = std::min(xmin, sx); xmax = std::max(xmax, sx); ymin = std::min(ymin, sy); ymax = std::max(ymax, sy);
        }
        // orthonormality check of cam_transform (the transpose is only its inverse then)
        const double d01 = T[0] * T[3] + T[1] * T[4] + T[2] * T[5], d00 = T[0] * T[0] + T[1] * T[1] + T[2] * T[2], d22 = T[6] * T[6] + T[7] * T[7] + T[8] * T[8];
        if (std::fabs(d01) > 1e-4 || std::fabs(d00 - 1) > 1e-4 || std::fabs(d22 - 1) > 1e-4) ok = false;
        if (ok && std::isfinite(xmin + xmax + ymin + ymax)) {
            box_cull_ok = true;
            // pixels outside [cx0, cx1] x [cy0, cy1] are exactly (0, 0, 0, 0) for every sample: they get no tickets at all
            // (the traced rectangle shrinks) and k_fold folds zeros for them without reading the sample buffer
            const int cx0 = int(std::floor(std::max(-1e9, xmin))) - 2, cx1 = int(std::ceil(std::min(1e9, xmax))) + 1;
            const int cy0 = int(std::floor(std::max(-1e9, ymin))) - 2, cy1 = int(std::ceil(std::min(1e9, ymax))) + 1;
            a.x0 = std::max(a.x0, cx0); a.x1 = std::min(a.x1, cx1 + 1);
            a.y0 = std::max(a.y0, cy0); a.y1 = std::min(a.y1, cy1 + 1);
        }
    }
    const bool nothing_visible = a.x0 >= a.x1 || a.y0 >= a.y1;     // the whole region is culled: only the fold runs
    if (nothing_visible) { a.x0 = a.x1 = fold_x0; a.y0 = a.y1 = fold_y0; }
    // brick mask (k_tile_mask): needs a usable projection and, with a transfer function, a majorant that bounds the density
    const bool mask = box_cull_ok && !nothing_visible && params->vol_density_scale > 0.f && params->vol_majorant > 0.f &&
                      (!tf || (ctx->lut_monotone && params->tf_window_width > 0.f));
    a.tiles_x = std::max(1, (a.x1 - a.x0 + 7) / 8);
    const int n_tiles = std::max(1, a.tiles_x * ((a.y1 - a.y0 + 3) / 4));
    // ---- heaviest tiles first: order the blocks by the per-tile cost the previous launch of this view measured ----
    const bool lpt = ctx->lpt && !ctx->counting && (ctx->kernel == 0 || ctx->kernel == 3);    // the fast-math production schedules
    uint64_t vkey = key;
    if (a.tiles_x >= 65536 || n_tiles / a.tiles_x >= 65536 || ctx->w > 65535 || ctx->h > 65535) return fail(ctx, VRB_ERR_INVALID, "image too large");   // (the ray pool packs a pixel as y << 16 | x)
    {
        // the cost landscape depends on everything but the seed and the sample range
        auto mix2 = [&vkey](const void* p, size_t n) { const unsigned char* b = static_cast<const unsigned char*>(p); for (size_t i = 0; i < n; ++i) { vkey ^= b[i]; vkey *= 1099511628211ull; } };
        vrb_params pk = *params;
        pk.seed = 0;
        mix2(&pk, sizeof pk);
        mix2(&a.x0, 4 * sizeof(int));
        mix2(&ctx->env_w, sizeof(int));
        if (vkey == 0) vkey = 1;
        if (n_tiles > ctx->tile_capacity || a.tiles_x != ctx->tile_coords_tx) {
            cudaFree(ctx->tile_cost); cudaFree(ctx->tile_cost_sorted); cudaFree(ctx->tile_iota); cudaFree(ctx->tile_order); cudaFree(ctx->sort_tmp);
            cudaFree(ctx->tile_live); cudaFree(ctx->tile_key); cudaFree(ctx->live_info);
            ctx->tile_cost = ctx->tile_cost_sorted = nullptr; ctx->tile_iota = ctx->tile_order = nullptr; ctx->sort_tmp = nullptr;
            ctx->tile_live = ctx->tile_key = ctx->live_info = nullptr;
            ctx->tile_capacity = 0; ctx->cost_key = 0; ctx->order_key = 0;
            CK(cudaMalloc(&ctx->tile_live, size_t(n_tiles) * 4));
            CK(cudaMalloc(&ctx->tile_key, size_t(n_tiles) * 4));
            CK(cudaMalloc(&ctx->live_info, 8));
            CK(cudaMalloc(&ctx->tile_cost, size_t(n_tiles) * 4));
            CK(cudaMalloc(&ctx->tile_cost_sorted, size_t(n_tiles) * 4));
            CK(cudaMalloc(&ctx->tile_iota, size_t(n_tiles) * 4));
            CK(cudaMalloc(&ctx->tile_order, size_t(n_tiles) * 4));
            ctx->sort_tmp_bytes = 0;
            CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, ctx->sort_tmp_bytes, ctx->tile_cost, ctx->tile_cost_sorted, ctx->tile_iota, ctx->tile_order, n_tiles, 0, 32, ctx->stream));
            CK(cudaMalloc(&ctx->sort_tmp, ctx->sort_tmp_bytes ? ctx->sort_tmp_bytes : 16));
            k_tile_coords<<<grid_for(size_t(n_tiles), 256, ctx->sm_count), 256, 0, ctx->stream>>>(ctx->tile_iota, size_t(n_tiles), uint32_t(a.tiles_x));
            CK_LAUNCH();
            ctx->tile_capacity = n_tiles;
            ctx->tile_coords_tx = a.tiles_x;
        }
    }
    // kernel kinds on this path: 0 = ray-pool kernel (production), 2 = lane-resident kernel with IEEE math (cross-check,
    // vrb200_strict.cu), 3 = lane-resident kernel with fast math (round 1's production schedule, kept for A/B runs)
    const bool pool = ctx->kernel == 0;
    const int variant = (tf ? 1 : 0) | (ctx->counting ? 2 : 0) | (ctx->kernel == 2 ? 4 : 0) | (pool ? 8 : 0);
    const void* fn = nullptr;
    switch (variant) {
        case 8: fn = (const void*)k_trace_pool<false, false, FastMath>; break;
        case 9: fn = (const void*)k_trace_pool<true, false, FastMath>; break;
        case 10: fn = (const void*)k_trace_pool<false, true, FastMath>; break;
        case 11: fn = (const void*)k_trace_pool<true, true, FastMath>; break;
        case 0: fn = (const void*)k_trace_persistent<false, false, FastMath>; break;
        case 1: fn = (const void*)k_trace_persistent<true, false, FastMath>; break;
        case 2: fn = (const void*)k_trace_persistent<false, true, FastMath>; break;
        case 3: fn = (const void*)k_trace_persistent<true, true, FastMath>; break;
        default: fn = strict_persistent_kernel(tf, ctx->counting); break;
    }
    const int block_threads = pool ? VR_POOL_WARPS * 32 : VR_TRACE_BLOCK;
    const size_t dyn_smem = pool ? pool_smem_bytes() : 0;
    if (!ctx->trace_blocks[variant]) {
        if (pool) {
            CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dyn_smem)));
            CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block_threads, dyn_smem));
        ctx->trace_blocks[variant] = ctx->sm_count * (per_sm > 0 ? per_sm : 1);   // one resident wave: grid = 148 x blocks/SM
    }
    const int first_end = first_sample + n_samples;
    cudaStream_t lane_stream[2] = { ctx->stream, two_lanes ? ctx->pass_stream : ctx->stream };
    float4* lane_lbuf[2] = { ctx->lbuf, two_lanes ? ctx->lbuf2 : ctx->lbuf };
    if (two_lanes) {        // the second lane starts behind everything the caller enqueued on the context's stream so far
        CK(cudaEventRecord(ctx->ev_entry, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->pass_stream, ctx->ev_entry, 0));
    }
    int n_pass = 0;
    int order_lane = -1;                          // lane whose stream rebuilt the mask / order arrays last in this call (ev_order)
    bool fold_pending[2] = { false, false };      // ev_fold[lane] marks the end of that lane's latest fold
    for (int s0 = first_sample; s0 < first_end; s0 += pass, ++n_pass) {
        const int lane = two_lanes ? (n_pass & 1) : 0;
        cudaStream_t st_ = lane_stream[lane];
        a.first_sample = s0;
        a.n_samples = std::min(pass, first_end - s0);
        a.lbuf = lane_lbuf[lane];
        a.job_counter = ctx->job_counter + lane;
        // 32-bit block counter: tiles * samples blocks per pass (a pass holds at most 2^30 / 16 sample-pixels)
        a.sample_bits = 0;
        while ((1 << a.sample_bits) < a.n_samples) ++a.sample_bits;
        a.n_jobs = n_tiles << a.sample_bits;
        a.tile_order = ctx->tile_iota;       // natural order
        a.tile_cost = nullptr;
        a.n_live = nullptr;
        // the fold of pass k comes after the fold of pass k - 1 (other lane): the running mean is sequential in the sample index
        auto wait_other_fold = [&]() -> cudaError_t {
            if (two_lanes && fold_pending[lane ^ 1]) return cudaStreamWaitEvent(st_, ctx->ev_fold[lane ^ 1], 0);
            return cudaSuccess;
        };
        if (nothing_visible) {
            CK(wait_other_fold());
            k_fold<<<dim3((fold_x1 - fold_x0 + 63) / 64, (fold_y1 - fold_y0 + 3) / 4), 256, 0, st_>>>(
                ctx->color, a.lbuf, n_px, ctx->w, fold_x0, fold_y0, fold_x1, fold_y1, a.x0, a.y0, a.x1, a.y1, s0, a.n_samples, accum_mode, nullptr, nullptr, 0);
            CK_LAUNCH();
            if (two_lanes) { CK(cudaEventRecord(ctx->ev_fold[lane], st_)); fold_pending[lane] = true; }
            continue;
        }
        CK(cudaMemsetAsync(a.job_counter, 0, sizeof(unsigned int), st_));
        const bool cost_valid = lpt && ctx->cost_key == vkey;      // the previous pass / launch measured this view
        // The brick mask and the tile order are functions of (view, grid contents, mode): a static view re-uses them and
        // re-sorts from the latest measured costs every ORDER_REUSE passes (mask + keys + 4 radix passes are ~50 us per pass)
        constexpr int ORDER_REUSE = 8;
        uint64_t okey = vkey ^ (g.version * 0x9e3779b97f4a7c15ull) ^ (mask ? 0x5bd1e995ull : 0ull) ^ (cost_valid ? 0xc2b2ae35ull << 8 : 0ull);
        if (okey == 0) okey = 1;
        if ((mask || cost_valid) && ctx->order_key == okey && ctx->order_age < ORDER_REUSE) {
            ++ctx->order_age;
            a.tile_order = ctx->tile_order;
            if (mask) a.n_live = ctx->live_info;
            // the arrays may have been (re)built on the OTHER lane's stream by the previous pass: read them behind that build
            if (two_lanes && order_lane >= 0 && order_lane != lane) CK(cudaStreamWaitEvent(st_, ctx->ev_order, 0));
        } else if (mask || cost_valid) {
            // the order / mask arrays are rewritten: nothing of the other lane may still read them (its tracking kernel reads the
            // order and the live count, its fold the mask)
            CK(wait_other_fold());
            ctx->order_key = okey; ctx->order_age = 1;
            if (mask) {
                CK(cudaMemsetAsync(ctx->tile_live, 0, size_t(n_tiles) * 4, st_));
                CK(cudaMemsetAsync(ctx->live_info, 0, 8, st_));
                k_tile_mask<<<grid_for(n0, 128, ctx->sm_count), 128, 0, st_>>>(a, g.maj[0], ctx->tile_live, ctx->live_info, (a.y1 - a.y0 + 3) / 4);
                CK_LAUNCH();
                k_tile_keys<<<grid_for(size_t(n_tiles), 256, ctx->sm_count), 256, 0, st_>>>(ctx->tile_live, cost_valid ? ctx->tile_cost : nullptr, ctx->tile_key, ctx->live_info, n_tiles);
                CK_LAUNCH();
                ctx->trace_launches += 2;      // k_tile_mask + k_tile_keys
                size_t tmp_bytes = ctx->sort_tmp_bytes;
                CK(cub::DeviceRadixSort::SortPairsDescending(ctx->sort_tmp, tmp_bytes, ctx->tile_key, ctx->tile_cost_sorted, ctx->tile_iota, ctx->tile_order, n_tiles, 0, 32, st_));
                a.n_live = ctx->live_info;
            } else {
                size_t tmp_bytes = ctx->sort_tmp_bytes;
                CK(cudaStreamSynchronize(st_) == cudaSuccess ? cudaSuccess : cudaGetLastError());
                CK(cub::DeviceRadixSort::SortPairsDescending(ctx->sort_tmp, tmp_bytes, ctx->tile_cost, ctx->tile_cost_sorted, ctx->tile_iota, ctx->tile_order, n_tiles, 0, 32, st_));
            }
            a.tile_order = ctx->tile_order;
            if (lpt) CK(cudaMemsetAsync(ctx->tile_cost, 0, size_t(n_tiles) * 4, st_));     // costs accumulate until the next re-sort
            if (two_lanes) { CK(cudaEventRecord(ctx->ev_order, st_)); order_lane = lane; }
        }
        if (lpt) {
            if (ctx->cost_key != vkey) CK(cudaMemsetAsync(ctx->tile_cost, 0, size_t(n_tiles) * 4, st_));   // first pass of a new view
            ctx->cost_key = vkey;
            a.tile_cost = ctx->tile_cost;
        }
        const int per_block = pool ? VR_POOL_WARPS * ((VR_POOL_SLOTS + 31) / 32) : VR_TRACE_BLOCK / 32;     // blocks of 32 samples one CTA holds at a time
        const int needed = (n_tiles * a.n_samples + per_block - 1) / per_block;
        const int blocks = needed < ctx->trace_blocks[variant] ? (needed > 0 ? needed : 1) : ctx->trace_blocks[variant];
        void* kargs[] = { (void*)&a };
        nvtxRangePushA("vrb:trace pass (tracking kernel)");
        CK(cudaLaunchKernel(fn, dim3(blocks), dim3(block_threads), kargs, dyn_smem, st_));
        nvtxRangePop();
        NvtxRange nvtx_fold("vrb:trace pass (fold)");
        ctx->trace_launches += 2;          // the tracking kernel + k_fold below
        CK(wait_other_fold());
        k_fold<<<dim3((fold_x1 - fold_x0 + 63) / 64, (fold_y1 - fold_y0 + 3) / 4), 256, 0, st_>>>(
            ctx->color, a.lbuf, n_px, ctx->w, fold_x0, fold_y0, fold_x1, fold_y1, a.x0, a.y0, a.x1, a.y1, s0, a.n_samples, accum_mode,
            mask ? ctx->tile_live : nullptr, mask ? ctx->live_info : nullptr, a.tiles_x);
        CK_LAUNCH();
        if (two_lanes) { CK(cudaEventRecord(ctx->ev_fold[lane], st_)); fold_pending[lane] = true; }
    }
    // the caller's stream continues behind the second lane's last fold
    if (two_lanes && fold_pending[1]) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_fold[1], 0));
    return VRB_OK;
}

int vrb_probe_bandwidth(vrb_ctx* ctx, size_t bytes, int mode, double* gbytes_per_s) {
    if (!ctx) return VRB_ERR_INVALID;
    if (!gbytes_per_s || bytes < (size_t(1) << 20) || bytes > (size_t(8) << 30) || (mode != 0 && mode != 1)) return fail(ctx, VRB_ERR_INVALID, "bad probe arguments");
    DeviceGuard guard(ctx->device);
    uint4* buf = nullptr;
    uint32_t* sink = nullptr;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemsetAsync(buf, 1, bytes, ctx->stream));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    double moved = 0;
    if (mode == 0) {
        const size_t n16 = bytes / 16;
        const int passes = int(std::max<size_t>(1, (size_t(8) << 30) / bytes));        // ~8 GiB of traffic per timed launch
        k_probe_stream<<<ctx->sm_count * 4, 512, 0, ctx->stream>>>(buf, n16, 1, sink);
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0, ctx->stream));
            k_probe_stream<<<ctx->sm_count * 4, 512, 0, ctx->stream>>>(buf, n16, passes, sink);
            CK(cudaEventRecord(e1, ctx->stream));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
        }
        moved = double(n16) * 16.0 * passes;
    } else {
        const uint32_t n_sectors = uint32_t(bytes / 32);
        const int iters = 256, grid = ctx->sm_count * 16;
        k_probe_gather<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(buf), n_sectors, 8, sink);
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0, ctx->stream));
            k_probe_gather<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(buf), n_sectors, iters, sink);
            CK(cudaEventRecord(e1, ctx->stream));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
        }
        moved = double(grid) * 256.0 * iters * 4.0 * 32.0;      // sectors x 32 B
    }
    CK_LAUNCH();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(sink);
    *gbytes_per_s = moved / (double(best) * 1e-3) / 1e9;
    return VRB_OK;
}

int vrb_set_option(vrb_ctx* ctx, const char* name, int value) {
    if (!ctx) return VRB_ERR_INVALID;
    if (!name) return fail(ctx, VRB_ERR_INVALID, "NULL option name");
    if (!strcmp(name, "lpt")) ctx->lpt = value != 0;
    else if (!strcmp(name, "cull")) ctx->cull = value != 0;
    else if (!strcmp(name, "count_culled")) ctx->count_culled = value != 0;
    else if (!strcmp(name, "async_upload")) ctx->async_upload = value != 0;
    else if (!strcmp(name, "overlap")) ctx->overlap = value != 0;
    else if (!strcmp(name, "l2_persist")) { if (value < 0) return fail(ctx, VRB_ERR_INVALID, "l2_persist must be >= 0 (MiB)"); ctx->l2_persist_mb = value; }
    else if (!strcmp(name, "pass")) { if (value < 1) return fail(ctx, VRB_ERR_INVALID, "pass must be >= 1"); ctx->pass_samples = value; }
    else return fail(ctx, VRB_ERR_INVALID, "unknown option '%s'", name);
    return VRB_OK;
}

int vrb_get_stat(vrb_ctx* ctx, const char* name, uint64_t* out) {
    if (!ctx) return VRB_ERR_INVALID;
    if (!name || !out) return fail(ctx, VRB_ERR_INVALID, "NULL argument");
    if (!strcmp(name, "trace_launches")) *out = ctx->trace_launches;
    else return fail(ctx, VRB_ERR_INVALID, "unknown statistic '%s'", name);
    return VRB_OK;
}

int vrb_set_kernel(vrb_ctx* ctx, int kind) {
    if (!ctx) return VRB_ERR_INVALID;
    if (kind < 0 || kind > 3) return fail(ctx, VRB_ERR_INVALID, "kernel kind must be 0 (ray pool, fast math), 1 (one thread per pixel, IEEE math), 2 (lane-resident persistent, IEEE math) or 3 (lane-resident persistent, fast math)");
    ctx->kernel = kind;
    return VRB_OK;
}

int vrb_trace_deterministic(vrb_ctx* ctx, const vrb_params* params) {
    if (!ctx) return VRB_ERR_INVALID;
    TraceArgs a;
    const int st = fill_trace_args(ctx, params, a);
    if (st) return st;
    DeviceGuard guard(ctx->device);
    const dim3 grid((ctx->w + 15) / 16, (ctx->h + 15) / 16);
    k_trace_deterministic<<<grid, 256, 0, ctx->stream>>>(a);
    CK_LAUNCH();
    return VRB_OK;
}

int vrb_scale(vrb_ctx* ctx, float s) {
    if (!ctx) return VRB_ERR_INVALID;
    if (!ctx->color) return fail(ctx, VRB_ERR_STATE, "no colour buffer");
    DeviceGuard guard(ctx->device);
    const size_t n = size_t(ctx->w) * ctx->h;
    k_scale<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(ctx->color, n, s);
    CK_LAUNCH();
    return VRB_OK;
}

int vrb_clear(vrb_ctx* ctx) {
    if (!ctx) return VRB_ERR_INVALID;
    if (!ctx->color) return fail(ctx, VRB_ERR_STATE, "no colour buffer");
    DeviceGuard guard(ctx->device);
    CK(cudaMemsetAsync(ctx->color, 0, size_t(ctx->w) * ctx->h * sizeof(float4), ctx->stream));
    return VRB_OK;
}

int vrb_set_counting(vrb_ctx* ctx, int enable) {
    if (!ctx) return VRB_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    ctx->counting = enable != 0;
    CK(cudaMemsetAsync(ctx->counters, 0, 7 * sizeof(unsigned long long), ctx->stream));
    return VRB_OK;
}

int vrb_get_counters(vrb_ctx* ctx, vrb_counters* out) {
    if (!ctx || !out) return VRB_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    unsigned long long v[7];
    CK(cudaMemcpyAsync(v, ctx->counters, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    out->n_samples = v[0]; out->n_maj = v[1]; out->n_dens = v[2]; out->n_emis = v[3]; out->n_nee = v[4]; out->n_env = v[5]; out->n_real = v[6];
    return VRB_OK;
}

int vrb_tonemap(vrb_ctx* ctx, float exposure, float gamma, int in_place, int tonemapping) {
    NvtxRange nvtx_("vrb:tonemap");
    if (!ctx) return VRB_ERR_INVALID;
    if (!ctx->color) return fail(ctx, VRB_ERR_STATE, "no colour buffer");
    DeviceGuard guard(ctx->device);
    const size_t n = size_t(ctx->w) * ctx->h;
    const int blocks = grid_for(n, 256, ctx->sm_count);
    if (in_place) k_tonemap_inplace<<<blocks, 256, 0, ctx->stream>>>(ctx->color, n, exposure, 1.f / gamma);
    else k_draw<<<blocks, 256, 0, ctx->stream>>>(ctx->color, ctx->fb, n, exposure, 1.f / gamma, tonemapping);
    CK_LAUNCH();
    return VRB_OK;
}

int vrb_download_color(vrb_ctx* ctx, float* out, int channels) {
    NvtxRange nvtx_("vrb:download_color");
    if (!ctx) return VRB_ERR_INVALID;
    if (!ctx->color) return fail(ctx, VRB_ERR_STATE, "no colour buffer");
    if (!out || (channels != 3 && channels != 4)) return fail(ctx, VRB_ERR_INVALID, "channels must be 3 or 4");
    DeviceGuard guard(ctx->device);
    const size_t n = size_t(ctx->w) * ctx->h;
    if (channels == 4) {
        CK(cudaMemcpyAsync(out, ctx->color, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else {
        CK(cudaMemcpy2DAsync(out, 12, ctx->color, 16, 12, n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return VRB_OK;
}

int vrb_download_color_ldr(vrb_ctx* ctx, uint8_t* rgba8) {
    if (!ctx) return VRB_ERR_INVALID;
    if (!ctx->color) return fail(ctx, VRB_ERR_STATE, "no colour buffer");
    if (!rgba8) return fail(ctx, VRB_ERR_INVALID, "NULL output");
    DeviceGuard guard(ctx->device);
    const size_t n = size_t(ctx->w) * ctx->h;
    k_color_to_ldr<<<grid_for(n, 256, ctx->sm_count), 256, 0, ctx->stream>>>(ctx->color, ctx->ldr, n);
    CK_LAUNCH();
    CK(cudaMemcpyAsync(rgba8, ctx->ldr, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VRB_OK;
}

int vrb_download_framebuffer(vrb_ctx* ctx, uint8_t* rgba8) {
    if (!ctx) return VRB_ERR_INVALID;
    if (!ctx->fb) return fail(ctx, VRB_ERR_STATE, "no framebuffer");
    if (!rgba8) return fail(ctx, VRB_ERR_INVALID, "NULL output");
    DeviceGuard guard(ctx->device);
    CK(cudaMemcpyAsync(rgba8, ctx->fb, size_t(ctx->w) * ctx->h * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VRB_OK;
}

int vrb_upload_color(vrb_ctx* ctx, const float* rgba) {
    if (!ctx) return VRB_ERR_INVALID;
    if (!ctx->color) return fail(ctx, VRB_ERR_STATE, "no colour buffer");
    if (!rgba) return fail(ctx, VRB_ERR_INVALID, "NULL input");
    DeviceGuard guard(ctx->device);
    CK(cudaMemcpyAsync(ctx->color, rgba, size_t(ctx->w) * ctx->h * 16, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VRB_OK;
}

void* vrb_color_device_ptr(vrb_ctx* ctx) { return ctx ? ctx->color : nullptr; }

int vrb_reduce(vrb_ctx* const* ctxs, int n, int root) {
    NvtxRange nvtx_("vrb:reduce");
    if (!ctxs || n <= 0 || root < 0 || root >= n) return VRB_ERR_INVALID;
    vrb_ctx* ctx = ctxs[root];
    if (!ctx || !ctx->color) return VRB_ERR_INVALID;
    const size_t px = size_t(ctx->w) * ctx->h;
    for (int i = 0; i < n; ++i)
        if (!ctxs[i] || !ctxs[i]->color || ctxs[i]->w != ctx->w || ctxs[i]->h != ctx->h) return fail(ctx, VRB_ERR_INVALID, "context %d does not match the root resolution", i);
    DeviceGuard guard(ctx->device);
    float4* staging = nullptr;
    CK(cudaMalloc(&staging, px * sizeof(float4)));
    for (int i = 0; i < n; ++i) {
        if (i == root) continue;
        { DeviceGuard g2(ctxs[i]->device); cudaStreamSynchronize(ctxs[i]->stream); }
        CK(cudaMemcpyPeerAsync(staging, ctx->device, ctxs[i]->color, ctxs[i]->device, px * sizeof(float4), ctx->stream));
        k_add<<<grid_for(px, 256, ctx->sm_count), 256, 0, ctx->stream>>>(ctx->color, staging, px);
        CK_LAUNCH();
    }
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(staging);
    return VRB_OK;
}

int vrb_copy_rows(vrb_ctx* dst, vrb_ctx* src, int y0, int y1) {
    NvtxRange nvtx_("vrb:copy_rows");
    if (!dst || !src) return VRB_ERR_INVALID;
    vrb_ctx* ctx = dst;
    if (!dst->color || !src->color || dst->w != src->w || dst->h != src->h) return fail(ctx, VRB_ERR_INVALID, "contexts do not share a resolution");
    if (y0 < 0 || y1 > dst->h || y0 >= y1) return fail(ctx, VRB_ERR_INVALID, "bad row range [%d, %d)", y0, y1);
    { DeviceGuard g2(src->device); cudaStreamSynchronize(src->stream); }
    DeviceGuard guard(dst->device);
    const size_t off = size_t(y0) * dst->w, bytes = size_t(y1 - y0) * dst->w * sizeof(float4);
    if (dst->device == src->device) CK(cudaMemcpyAsync(dst->color + off, src->color + off, bytes, cudaMemcpyDeviceToDevice, dst->stream));
    else CK(cudaMemcpyPeerAsync(dst->color + off, dst->device, src->color + off, src->device, bytes, dst->stream));
    CK(cudaStreamSynchronize(dst->stream));
    return VRB_OK;
}

}  // extern "C"
