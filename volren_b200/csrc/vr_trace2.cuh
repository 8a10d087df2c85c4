// Persistent-thread tracking kernel (the production path of vrb_trace).
//
// Same per-path algorithm and random-number order as the straightforward kernel in vr_trace.cuh
// (kept as the in-library cross-check, vrb_set_kernel(ctx, 1)), restructured for the SIMT machine:
//   * the unit of work is ONE PATH SAMPLE (pixel, sample index). Warps draw blocks of 32 samples = (8x4 pixel tile,
//     sample index) from a global counter (the next block is prefetched one switch ahead, so the atomic's round trip is
//     never on the critical path) and hand them to whichever lane is free. A sample's radiance goes to a per-launch
//     buffer lbuf[sample][pixel]; k_fold then folds the samples of each pixel into `color` IN SAMPLE ORDER, so the image
//     is bit-identical to successive dispatches of pathtracer_brick.glsl (:36 running mean). History: with whole-pixel
//     tickets (v3-v5) the heaviest pixel's samples ran back to back on one lane and a third of the SM time was tail
//     (profiles/r01_v4_*, r01_v5_*); handing chunks of a pixel to different lanes needed a release/acquire per hand-off
//     that cost more than the tail (profiles/r01_v4_chunk_sweep_*.txt). The buffer costs 32 B of traffic per sample;
//   * when a warp switches to a new block ALL 32 lanes prepare it together -- TEA seed (32 rounds, the largest single
//     cost of a short path), pixel jitter and view direction of the block's 32 samples go to shared memory -- instead of
//     each lane seeding its own sample whenever it happens to become free (17 of 32 lanes active before);
//   * with a hidden environment, tiles whose pixels cannot see the volume's bounding box (host-side conservative
//     projection of the box, one pixel of margin) are written as zeros without seeding or tracing anything;
//   * blocks are issued HEAVIEST TILE FIRST when the previous launch of the same view left per-tile costs (cycles a
//     sample occupied its lane), so the last blocks of a launch are the cheap ones;
//   * camera segments and shadow rays share ONE brick-DDA loop body (common.glsl:412-501 differ only in what
//     happens at a collision), so a warp's lanes step convergently whatever kind of ray they are on;
//   * path events are scheduled wavefront-style INSIDE the warp: a lane whose ray hit a tentative collision or ended
//     parks in one of four queues (COLLIDE / NEE / SCATTER / FINISH); a queue's stage runs when enough lanes wait
//     in it (or nothing else can make progress), so the heavy stages (8-tap trilinear + LUT collision test,
//     importance-pyramid warp, phase sampling, escape lookup) execute with many active lanes
//     instead of a few; every new ray (camera, shadow, scattered) is started at ONE shared site after the stages;
//   * per-level majorants are read from float tables precomputed per (grid, params) with the identical
//     expression (the TF variant otherwise evaluates a LUT lerp and a divide on every DDA step);
//   * MT = FastMath in production (MUFU rcp/rsqrt/lg2/sin/cos): 3x less SASS than the IEEE sequences, which
//     matters because the kernel was instruction-fetch bound (profiles/r01_v1_*: stall_no_instruction 50 %).
#pragma once

#include "vr_trace.cuh"

namespace vr {

// lane stages
enum : int { SG_STEP = 0, SG_COLLIDE = 1, SG_NEE = 2, SG_SCATTER = 3, SG_FINISH = 4, SG_IDLE = 5 };

// fills one level of the majorant table: exactly majorant_at() of vr_trace.cuh, hoisted out of the DDA loop
template <bool TF>
VR_GLOBAL void k_majorant_table      // (internal linkage in the strict unit: the two units instantiate DIFFERENT arithmetic under one name)
    (const __grid_constant__ TraceArgs a, int level, float* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const uint32_t w = level == 0 ? a.density.rec[i].y : a.density.mips[level - 1][i];
        const float m = a.p.vol_density_scale * range_hi(w);
        out[i] = TF ? a.p.vol_majorant * tf_lookup_alpha(a, m * a.p.vol_inv_majorant) : m;
    }
    if (level == 0 && blockIdx.x == 0 && threadIdx.x == 0) {   // out-of-bounds texelFetch returns 0 (robust access)
        const float m = a.p.vol_density_scale * 0.f;
        out[n] = TF ? a.p.vol_majorant * tf_lookup_alpha(a, m * a.p.vol_inv_majorant) : m;
    }
}

// The same lookup with one base pointer: the four tables and the out-of-bounds entry live in one allocation (alloc_hot), so the
// level selects a 32-bit element offset instead of a 64-bit pointer and the out-of-bounds case selects an INDEX instead of a
// second load site -- one IMAD.WIDE + one LDG instead of LDC.64 + four address instructions + two predicated loads.
VR_DEV float table_majorant_idx(const TraceArgs& a, float3 ipos, int mip) {
    const int bx = int(floorf(ipos.x)) >> (3 + mip), by = int(floorf(ipos.y)) >> (3 + mip), bz = int(floorf(ipos.z)) >> (3 + mip);
    const uint32_t nx = a.density.nb.x >> mip, ny = a.density.nb.y >> mip, nz = a.density.nb.z >> mip;
    const bool in = unsigned(bx) < nx && unsigned(by) < ny && unsigned(bz) < nz;
    const uint32_t idx = in ? a.maj_off[mip] + (uint32_t(bz) * ny + uint32_t(by)) * nx + uint32_t(bx) : a.maj_off_oob;
    return __ldg(a.maj[0] + idx);
}

VR_DEV float table_majorant(const TraceArgs& a, float3 ipos, int mip) {
    const int bx = int(floorf(ipos.x)) >> (3 + mip), by = int(floorf(ipos.y)) >> (3 + mip), bz = int(floorf(ipos.z)) >> (3 + mip);
    const uint32_t nx = a.density.nb.x >> mip, ny = a.density.nb.y >> mip, nz = a.density.nb.z >> mip;
    if (unsigned(bx) >= nx || unsigned(by) >= ny || unsigned(bz) >= nz) return __ldg(a.maj_oob);
    return __ldg(a.maj[mip] + (uint32_t(bz) * ny + uint32_t(by)) * nx + uint32_t(bx));   // < 2^30 entries: 32-bit index math
}

#ifndef VR_TRACE_BLOCK
#define VR_TRACE_BLOCK 128
#endif
#ifndef VR_TRACE_MIN_BLOCKS
#define VR_TRACE_MIN_BLOCKS 7     // CTAs per SM: 7 -> 72 registers without spills (28 warps/SM). B200, TF / non-TF Gsamples/s:
                                  // 6 (80 regs) 40.8 / 4.36, 7 (72) 44.1 / 4.41, 8 (64, spills) 41.9 / 4.32 (profiles/r01_v10_duo_and_steps_sweeps.txt)
#endif
// Queue thresholds, tuned on B200 (tools/sweep.py, profiles/r01_sweep*.txt): a queue's stage runs once this many lanes
// wait in it. The TF variant has cheap events and an expensive 8-tap collision; the non-TF variant a cheap 1-tap
// collision and relatively expensive events.
#ifndef VR_K_EVENT_TF
#define VR_K_EVENT_TF 4
#endif
#ifndef VR_K_EVENT
#define VR_K_EVENT 8
#endif
#ifndef VR_K_FINISH_TF
#define VR_K_FINISH_TF VR_K_EVENT_TF
#endif
#ifndef VR_K_FINISH
#define VR_K_FINISH VR_K_EVENT
#endif
#ifndef VR_K_COLLIDE_TF
#define VR_K_COLLIDE_TF 8
#endif
#ifndef VR_K_COLLIDE
#define VR_K_COLLIDE 4
#endif
#ifndef VR_MIN_STEP_TF
#define VR_MIN_STEP_TF 8
#endif
#ifndef VR_MIN_STEP
#define VR_MIN_STEP 24    // fewer stepping lanes than this: drain the fullest queue even below its threshold
#endif
#ifndef VR_STEPS_PER_PASS
#define VR_STEPS_PER_PASS 3       // DDA steps per scheduler pass: the five ballots + queue logic are 7 % of the issued instructions at 32 lanes;
                                  // B200 (TF / non-TF Gsamples/s): 1 -> 37.0 / 4.01, 2 -> 39.0 / 4.19, 3 -> 38.8 / 4.26, 4 -> 38.2 / 4.16, 6 -> 36.6 / 3.88
#endif
#ifndef VR_LBUF_STREAM
#define VR_LBUF_STREAM 1          // sample-buffer stores are streaming (evict-first): written once, read once by k_fold
#endif
#if VR_LBUF_STREAM
#define VR_LBUF_STORE(ptr, v) __stcs(ptr, v)
#else
#define VR_LBUF_STORE(ptr, v) (*(ptr) = (v))
#endif
constexpr int MAX_RAY_STEPS = 1 << 20;  // hang guard only: no finite ray takes this many DDA steps

template <bool TF, bool COUNT, class MT>
VR_GLOBAL void __launch_bounds__(VR_TRACE_BLOCK, VR_TRACE_MIN_BLOCKS) k_trace_persistent(const __grid_constant__ TraceArgs a) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int W = a.p.resolution[0];
    Cnt<COUNT> cnt;

    // ---- lane state ----
    int stage = SG_FINISH;         // everybody starts by asking for a sample
    bool shadow = false;           // kind of the ray being stepped
    bool escaped = false;          // FINISH entered because the camera segment left the volume (-> environment)
    bool have_item = false;
    int px = 0, py = 0, sj = 0, steps = 0;   // the lane's sample: pixel and sample index relative to first_sample
    unsigned t_item = 0;           // clock at which this lane took its sample
    uint32_t seed = 0, n_paths = 0;
    float3 pos = f3(0.f), dir = f3(0.f, 0.f, -1.f), thr = f3(1.f);
    float3 L = f3(0.f), pend = f3(0.f);    // radiance so far; NEE term waiting for its shadow ray
    float3 ipos = f3(0.f), idir = f3(1.f), ri = f3(1.f);
    float t = 0.f, tfar = -1.f, tau = 0.f, mip = 3.f, f_p = 0.f, Tr = 1.f, majorant = 0.f;
    // ---- warp state: the current block of 32 samples (one tile, one sample index), prepared in shared memory ----
    __shared__ float4 s_prep[VR_TRACE_BLOCK / 32][32];     // {view dir, seed after the two jitter draws}
    float4* prep = s_prep[threadIdx.x >> 5];
    int blk_x0 = 0, blk_y0 = 0, blk_sj = 0;
    int blk_next = 32;             // next sample of the block to hand out (>= 32: block used up)
    bool blk_done = false;         // the counter is exhausted
    unsigned nxt_block = 0;        // lane 0: id of the prefetched next block
    if (lane == 0) nxt_block = atomicAdd(a.job_counter, 1u);
    const unsigned n_jobs = a.n_live ? (__ldg(a.n_live) << a.sample_bits) : unsigned(a.n_jobs);

    while (true) {
        // ================= STEP: one brick-DDA step (common.glsl:423-435 / 470-482) =================
#pragma unroll
        for (int rep = 0; rep < VR_STEPS_PER_PASS; ++rep) {
        if (stage == SG_STEP) {
            if (t < tfar) {
                const float3 curr = ipos + t * idir;
                const int m = round_mip(mip);
                cnt.maj();
                majorant = table_majorant(a, curr, m);
                const float dt = step_dda(curr, ri, m);
                t += dt;
                tau -= majorant * dt;
                mip = fminf(mip + 0.25f, 3.f);
                if (!(tau > 0.f)) {
                    t += MT::div(tau, majorant);
                    if (!(t >= tfar)) stage = SG_COLLIDE;   // the reference tests `if (t >= far) break;` (a NaN t goes on to the lookup)
                }
                if (++steps > MAX_RAY_STEPS) { t = INFINITY; stage = SG_STEP; }
            }
            if (stage == SG_STEP && !(t < tfar)) {   // the ray left the volume
                if (shadow) {
                    if (Tr != 0.f) L = L + pend * Tr;     // (x * 0) stays 0 even if pend overflowed, as in the reference's order
                    stage = SG_SCATTER;
                } else {
                    escaped = true;
                    stage = SG_FINISH;
                }
            }
        }
        }

        // ================= scheduler =================
        const unsigned m_step = __ballot_sync(FULL, stage == SG_STEP);
        const unsigned m_col = __ballot_sync(FULL, stage == SG_COLLIDE);
        const unsigned m_nee = __ballot_sync(FULL, stage == SG_NEE);
        const unsigned m_scat = __ballot_sync(FULL, stage == SG_SCATTER);
        const unsigned m_fin = __ballot_sync(FULL, stage == SG_FINISH);
        if ((m_step | m_col | m_nee | m_scat | m_fin) == 0u) break;     // every lane idle
        const int n_step = __popc(m_step), n_col = __popc(m_col), n_nee = __popc(m_nee), n_scat = __popc(m_scat), n_fin = __popc(m_fin);
        constexpr int KF = TF ? VR_K_FINISH_TF : VR_K_FINISH;
        constexpr int K = TF ? VR_K_EVENT_TF : VR_K_EVENT, KC = TF ? VR_K_COLLIDE_TF : VR_K_COLLIDE, MIN_STEP = TF ? VR_MIN_STEP_TF : VR_MIN_STEP;
        bool run_col = n_col >= KC;
        bool run_nee = n_nee >= K, run_scat = n_scat >= K, run_fin = n_fin >= KF;
        if (!(run_col | run_nee | run_scat | run_fin)) {
            if (n_step >= MIN_STEP) continue;                  // keep stepping
            // too few lanes can step: drain the fullest queue
            if (n_col >= n_nee && n_col >= n_scat && n_col >= n_fin) run_col = n_col > 0;
            else if (n_nee >= n_scat && n_nee >= n_fin) run_nee = n_nee > 0;
            else if (n_scat >= n_fin) run_scat = n_scat > 0;
            else run_fin = n_fin > 0;
        }

        // ================= COLLIDE: tentative collision (common.glsl:436-452 / 483-498) =================
        if (run_col && stage == SG_COLLIDE) {
            cnt.dens();
            stage = SG_STEP;
            const float3 at = ipos + t * idir;
            float d;
            float3 tf_rgb = f3(1.f);
            if (TF) {
                const float4 rgba = tf_lookup<MT>(a, a.p.vol_density_scale * (MT::decoded ? density_trilinear_decoded(a.density, at) : density_trilinear(a.density, at)) * a.p.vol_inv_majorant);
                d = a.p.vol_majorant * rgba.w;
                tf_rgb = f3(rgba.x, rgba.y, rgba.z);
            } else {
                const int3 tap = stochastic_tricubic_filter<MT>(at, seed);
                d = a.p.vol_density_scale * ((MT::decoded && VR_DECODED_TAP) ? decoded_value(a.density, tap.x, tap.y, tap.z) : brick_value(a.density, tap.x, tap.y, tap.z));
            }
            bool parked = false;
            if (!shadow) {
                bool fetched;
                const float3 em = lookup_emission<MT>(a, at, seed, fetched);
                if (fetched) {
                    cnt.emis();
                    const float3 albedo = f3(a.p.vol_albedo[0], a.p.vol_albedo[1], a.p.vol_albedo[2]);
                    L = L + thr * (f3(1.f) - albedo) * em * d * a.p.vol_inv_majorant;
                }
                if (rng(seed) * majorant < d) {          // real collision: the segment ends here (common.glsl:490-496)
                    thr = thr * f3(a.p.vol_albedo[0], a.p.vol_albedo[1], a.p.vol_albedo[2]);
                    if (TF) thr = thr * tf_rgb;
                    stage = SG_NEE;
                    parked = true;
                }
            } else {
                if (rng(seed) * majorant < d) {          // common.glsl:442-450
                    Tr *= fmaxf(0.f, 1.f - MT::div(a.p.vol_majorant, majorant));
                    if (Tr < .1f) {
                        const float prob = 1 - Tr;
                        if (rng(seed) < prob) { Tr = 0.f; stage = SG_SCATTER; parked = true; }   // absorbed: `return 0.f`, nothing is added to L
                        else Tr = MT::div(Tr, 1 - prob);
                    }
                }
            }
            if (!parked) {
                tau = -MT::log(1.f - rng(seed));
                mip = fmaxf(0.f, mip - 2.f);
            }
        }

        bool start = false;     // this lane starts a new ray from (pos, rd) below
        float3 rd = dir;

        // ================= NEE: real collision -> next-event estimation (common.glsl:611-626) =================
        if (run_nee && stage == SG_NEE) {
            cnt.real();
            pos = pos + t * dir;
            float3 w_i;
            const float r0 = rng(seed), r1 = rng(seed);
            cnt.nee();
            const float4 Le_pdf = sample_environment<MT>(a, r0, r1, w_i);
            if (Le_pdf.w > 0) {
                f_p = phase_hg<MT>(dot(-dir, w_i), a.p.vol_phase_g);
                const float mis_weight = a.p.show_environment > 0 ? MT::div(sqr(Le_pdf.w), sqr(Le_pdf.w) + sqr(f_p)) : 1.f;
                // L += throughput * mis_weight * f_p * Tr * Le / pdf, with Tr applied when the shadow ray is done
                const float3 c = thr * mis_weight * f_p * f3(Le_pdf.x, Le_pdf.y, Le_pdf.z);
                pend = f3(MT::div(c.x, Le_pdf.w), MT::div(c.y, Le_pdf.w), MT::div(c.z, Le_pdf.w));
                Tr = 1.f;
                shadow = true;
                rd = w_i; start = true;
                stage = SG_STEP;
            } else {
                stage = SG_SCATTER;
            }
        }

        // ================= SCATTER: bounce limit, Russian roulette, phase sampling (common.glsl:628-641) =================
        if (run_scat && stage == SG_SCATTER) {
            bool end = false;
            if (++n_paths >= uint32_t(a.p.bounces)) end = true;
            else {
                const float rr_val = luma(thr);
                if (rr_val < .1f) {
                    const float prob = 1 - rr_val;
                    if (rng(seed) < prob) end = true;
                    else { const float k = 1 - prob; thr = f3(MT::div(thr.x, k), MT::div(thr.y, k), MT::div(thr.z, k)); }
                }
            }
            if (end) {
                escaped = false;
                stage = SG_FINISH;
            } else {
                const float s0 = rng(seed), s1 = rng(seed);
                const float3 scatter_dir = sample_phase_hg<MT>(dir, a.p.vol_phase_g, s0, s1);
                f_p = phase_hg<MT>(dot(-dir, scatter_dir), a.p.vol_phase_g);
                dir = scatter_dir;
                shadow = false;
                rd = dir; start = true;
                stage = SG_STEP;
            }
        }

        // ================= FINISH: environment on escape, store the sample, take the next one =================
        if (run_fin) {     // warp-uniform: the block switch below is a warp-wide cooperative step
            const bool mine = stage == SG_FINISH;
            if (mine && have_item) {
                float3 Lf = L;
                if (escaped && a.p.show_environment > 0) {      // common.glsl:644-649
                    cnt.env();
                    const float3 Le = lookup_environment(a, dir);
                    const float pe = pdf_environment<MT>(a, Le);
                    const float mis_weight = n_paths > 0 ? MT::div(sqr(f_p), sqr(f_p) + sqr(pe)) : 1.f;
                    Lf = Lf + thr * mis_weight * Le;
                }
                cnt.samp();                                     // pathtracer_brick.glsl:36: sanitize(L), folded by k_fold
                VR_LBUF_STORE(a.lbuf + size_t(sj) * a.lbuf_stride + size_t(py) * W + px,
                              make_float4(sanitize(Lf.x), sanitize(Lf.y), sanitize(Lf.z), sanitize(fminf(float(n_paths), 1.f))));
                have_item = false;
                if (a.tile_cost && ((px ^ py ^ sj) & 3) == 0) {   // a dithered quarter of the samples is enough to rank tiles
                    unsigned now;
                    asm volatile("mov.u32 %0, %%clock;" : "=r"(now));
                    atomicAdd(a.tile_cost + ((py - a.y0) >> 2) * a.tiles_x + ((px - a.x0) >> 3), (now - t_item) >> 8);
                }
            }
            bool want = mine;
            while (true) {
                const unsigned m_want = __ballot_sync(FULL, want);
                if (m_want == 0u) break;
                if (blk_next >= 32) {                              // warp-uniform: switch to the prefetched block
                    unsigned b = 0xffffffffu;
                    if (!blk_done) {
                        b = __shfl_sync(FULL, nxt_block, 0);
                        if (lane == 0) nxt_block = atomicAdd(a.job_counter, 1u);
                    }
                    if (b >= n_jobs) {                             // no blocks left
                        blk_done = true;
                        if (want) { want = false; stage = SG_IDLE; }
                        break;
                    }
                    // tile-major: the blocks of a tile slot are consecutive (sample index in the low bits); the slot's tile
                    // comes from the order array (natural or heaviest-first), packed as (tile y << 16 | tile x)
                    blk_sj = int(b & ((1u << a.sample_bits) - 1u));
                    if (blk_sj >= a.n_samples) continue;           // padding of a non-power-of-two sample count
                    const unsigned txy = __ldg(a.tile_order + (b >> a.sample_bits));
                    blk_x0 = a.x0 + int(txy & 0xffffu) * 8;
                    blk_y0 = a.y0 + int(txy >> 16) * 4;
                    const int ix = blk_x0 + (lane & 7), iy = blk_y0 + (lane >> 3);
                    // all 32 lanes prepare the block: lane i seeds sample (pixel i of the tile, sample blk_sj)
                    // (pathtracer_brick.glsl:28-30: TEA seed, two jitter draws, view direction)
                    uint32_t sd = tea32(uint32_t(a.p.seed) * uint32_t(iy * W + ix), uint32_t(a.first_sample + blk_sj));
                    const float jx = rng(sd), jy = rng(sd);
                    const float3 vd = view_dir<MT>(a, ix, iy, jx, jy);
                    __syncwarp();
                    prep[lane] = make_float4(vd.x, vd.y, vd.z, __uint_as_float(sd));
                    __syncwarp();
                    blk_next = 0;
                }
                // the r-th wanting lane takes the r-th remaining sample of the block (samples of a border tile that fall
                // outside the image are dropped and the lane asks again)
                const int i = blk_next + __popc(m_want & ((1u << lane) - 1u));
                if (want && i < 32) {
                    px = blk_x0 + (i & 7);
                    py = blk_y0 + (i >> 3);
                    if (px < a.x1 && py < a.y1) {
                        const float4 pr = prep[i];
                        sj = blk_sj;
                        seed = __float_as_uint(pr.w);
                        dir = f3(pr.x, pr.y, pr.z);
                        have_item = true;
                        want = false;
                    }
                }
                blk_next += __popc(m_want);
            }
            // new sample: camera ray of the prepared sample
            if (mine && have_item) {
                asm volatile("mov.u32 %0, %%clock;" : "=r"(t_item));
                pos = f3(a.p.cam_pos[0], a.p.cam_pos[1], a.p.cam_pos[2]);
                thr = f3(1.f); L = f3(0.f); n_paths = 0; f_p = 0.f;
                shadow = false; escaped = false;
                rd = dir; start = true;
                stage = SG_STEP;
            }
        }

        // ================= start the new rays: clip + world->index + first free-flight draw (common.glsl:459-468 / 413-421) =================
        if (start) {
            float tn, tf;
            steps = 0;
            if (intersect_box<MT>(pos, rd, a.p.vol_bb_min, a.p.vol_bb_max, tn, tf)) {
                const Mat4& M = *reinterpret_cast<const Mat4*>(a.p.vol_density_inv_transform);
                ipos = mul_point(M, pos);
                idir = mul_dir(M, rd);
                ri = f3(MT::rcp(idir.x), MT::rcp(idir.y), MT::rcp(idir.z));
                t = tn + 1e-6f;
                tfar = tf;
                tau = -MT::log(1.f - rng(seed));
                mip = 3.f;
            } else {
                t = 0.f; tfar = -1.f;   // missed the box: the ray "ends" at once (Tr = 1 / escape)
            }
        }
    }
    flush_counters(a, cnt);
}

// Folds the samples of one launch into the colour buffer in sample order (pathtracer_brick.glsl:36):
// color = mix(color, L_s, 1 / s) for s = first_sample ... (VRB_ACCUM_MEAN) or color += L_s (VRB_ACCUM_SUM).
// Screen-space brick mask (hidden environment only). With the environment hidden a sample is non-zero only if its camera
// ray has a REAL collision, and a real collision at x needs density(x) > 0, hence a level-0 majorant > 0 in the brick
// that contains x (the majorants are conservative: the ranges are dilated by the filter footprint, grid_brick.cpp:83-92).
// x lies on the camera ray, so it projects into the sample's own pixel square: a tile onto which no brick with a positive
// majorant projects yields exactly (0, 0, 0, 0) for every sample, whatever the seed. One thread per brick marks the
// tiles under the screen bounding rectangle of its 8 corners (+ 1 pixel); a brick with a corner beside or behind the
// camera sets `info[1]` (mask unusable: every tile is live).
VR_GLOBAL void k_tile_mask(const __grid_constant__ TraceArgs a, const float* __restrict__ maj0, unsigned int* __restrict__ tile_live,
                            unsigned int* __restrict__ info, int tiles_y) {
    const uint32_t nbx = a.density.nb.x, nby = a.density.nb.y, nbz = a.density.nb.z;
    const size_t n = size_t(nbx) * nby * nbz;
    const float w = float(a.p.resolution[0]), h = float(a.p.resolution[1]);
    const Mat4& M = *reinterpret_cast<const Mat4*>(a.p.vol_density_transform);
    const float* T = a.p.cam_transform;      // column-major, orthonormal (checked by the host): inverse = transpose
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        if (maj0[i] <= 0.f) continue;        // (a NaN majorant counts as live)
        const uint32_t bx = uint32_t(i % nbx), by = uint32_t((i / nbx) % nby), bz = uint32_t(i / (size_t(nbx) * nby));
        float xmin = 1e30f, xmax = -1e30f, ymin = 1e30f, ymax = -1e30f;
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float3 c = mul_point(M, f3(float(8u * (bx + (k & 1))), float(8u * (by + ((k >> 1) & 1))), float(8u * (bz + (k >> 2))))) -
                             f3(a.p.cam_pos[0], a.p.cam_pos[1], a.p.cam_pos[2]);
            const float vx = T[0] * c.x + T[1] * c.y + T[2] * c.z, vy = T[3] * c.x + T[4] * c.y + T[5] * c.z, vz = T[6] * c.x + T[7] * c.y + T[8] * c.z;
            if (!(vz < -1e-4f)) { ok = false; break; }
            const float sx = vx * a.cam_z / vz * h + 0.5f * w, sy = vy * a.cam_z / vz * h + 0.5f * h;     // inverse of view_dir (common.glsl:76-80)
            xmin = fminf(xmin, sx); xmax = fmaxf(xmax, sx); ymin = fminf(ymin, sy); ymax = fmaxf(ymax, sy);
        }
        if (!ok || !(xmin <= xmax) || !(ymin <= ymax)) { info[1] = 1u; continue; }
        // pixel p covers [p, p + 1): pixels floor(min) - 1 ... floor(max) + 1, clipped to the traced rectangle
        const int px0 = max(a.x0, int(floorf(fmaxf(xmin, -1e9f))) - 1), px1 = min(a.x1 - 1, int(floorf(fminf(xmax, 1e9f))) + 1);
        const int py0 = max(a.y0, int(floorf(fmaxf(ymin, -1e9f))) - 1), py1 = min(a.y1 - 1, int(floorf(fminf(ymax, 1e9f))) + 1);
        if (px0 > px1 || py0 > py1) continue;
        for (int ty = (py0 - a.y0) >> 2; ty <= (py1 - a.y0) >> 2; ++ty)
            for (int tx = (px0 - a.x0) >> 3; tx <= (px1 - a.x0) >> 3; ++tx) tile_live[ty * a.tiles_x + tx] = 1u;
    }
}

// sort keys of the tile slots: dead tiles 0 (they sort behind every live tile), live tiles max(cost, 1) (cost = 0 everywhere
// when the view has not been measured yet: the stable sort keeps raster order); info[0] = number of live tiles
VR_GLOBAL void k_tile_keys(const unsigned int* __restrict__ tile_live, const unsigned int* __restrict__ cost, unsigned int* __restrict__ key,
                            unsigned int* __restrict__ info, int n_tiles) {
    const bool all_live = info[1] != 0u;
    unsigned int mine = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_tiles; i += gridDim.x * blockDim.x) {
        const bool live = all_live || tile_live[i] != 0u;
        key[i] = live ? max(cost ? cost[i] : 0u, 1u) : 0u;
        mine += live ? 1u : 0u;
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(info, mine);
}

// Pixels of the region [x0, x1) x [y0, y1) outside the traced rectangle [tx0, tx1) x [ty0, ty1) were culled on the host (no
// ray of theirs can reach the volume's box and the environment is hidden), pixels of dead tiles by k_tile_mask: each of
// their samples is exactly (0, 0, 0, 0).
VR_GLOBAL void __launch_bounds__(256) k_fold(float4* __restrict__ color, const float4* __restrict__ lbuf, size_t lbuf_stride, int W, int x0, int y0, int x1, int y1,
                                              int tx0, int ty0, int tx1, int ty1, int first_sample, int n_samples, int accum_mode,
                                              const unsigned int* __restrict__ tile_live, const unsigned int* __restrict__ info, int tiles_x) {
    // block = 64 x 4 pixels (grid covers the region): coalesced 1 KiB rows, no integer division
    const int x = x0 + blockIdx.x * 64 + (threadIdx.x & 63), y = y0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= x1 || y >= y1) return;
    const size_t p = size_t(y) * W + x;
    bool traced = x >= tx0 && x < tx1 && y >= ty0 && y < ty1;
    if (traced && tile_live && info[1] == 0u) traced = tile_live[((y - ty0) >> 2) * tiles_x + ((x - tx0) >> 3)] != 0u;    // k_tile_mask
    float4 acc = color[p];
    if (!traced) {
        // every sample is (0, 0, 0, 0): the sum does not change; the running mean of an all-zero pixel stays zero
        if (accum_mode != VRB_ACCUM_MEAN || (acc.x == 0.f && acc.y == 0.f && acc.z == 0.f && acc.w == 0.f)) return;
        for (int j = 0; j < n_samples; ++j) {
            const float w = 1.f / float(first_sample + j);
            acc.x = mix_rn(acc.x, 0.f, w); acc.y = mix_rn(acc.y, 0.f, w); acc.z = mix_rn(acc.z, 0.f, w); acc.w = mix_rn(acc.w, 0.f, w);
        }
        color[p] = acc;
        return;
    }
    const float4* src = lbuf + p;
    int j = 0;
    for (; j + 4 <= n_samples; j += 4) {            // four independent loads in flight per thread
        float4 L[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) L[k] = __ldcs(src + size_t(j + k) * lbuf_stride);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (accum_mode == VRB_ACCUM_MEAN) {
                const float w = 1.f / float(first_sample + j + k);
                acc.x = mix_rn(acc.x, L[k].x, w); acc.y = mix_rn(acc.y, L[k].y, w); acc.z = mix_rn(acc.z, L[k].z, w); acc.w = mix_rn(acc.w, L[k].w, w);
            } else {
                acc.x += L[k].x; acc.y += L[k].y; acc.z += L[k].z; acc.w += L[k].w;
            }
        }
    }
    for (; j < n_samples; ++j) {
        const float4 L = __ldcs(src + size_t(j) * lbuf_stride);
        if (accum_mode == VRB_ACCUM_MEAN) {
            const float w = 1.f / float(first_sample + j);
            acc.x = mix_rn(acc.x, L.x, w); acc.y = mix_rn(acc.y, L.y, w); acc.z = mix_rn(acc.z, L.z, w); acc.w = mix_rn(acc.w, L.w, w);
        } else {
            acc.x += L.x; acc.y += L.y; acc.z += L.z; acc.w += L.w;
        }
    }
    color[p] = acc;
}

}  // namespace vr
