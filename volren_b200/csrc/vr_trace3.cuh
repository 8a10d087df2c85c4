// Two rays per lane (vrb_set_kernel(ctx, 3); 4 = the same schedule with IEEE math for the path-for-path cross-check).
//
// The persistent kernel of vr_trace2.cuh keeps ONE ray per lane: a lane whose ray waits in an event queue (tentative
// collision, NEE, scatter, finish) idles until the queue runs, and the queues run with the few lanes that wait
// (COLLIDE 13.5, NEE 5.6 of 32 lanes on configs[1]; profiles/r01_v9_*). Every threshold sweep came out flat: the idle
// lanes are set by the asynchrony of 32 rays, not by the policy. Here a lane owns TWO rays. The state the DDA step needs
// (index-space ray, t, tau, mip, majorant, seed, stage: 17 registers per ray) stays in registers for both; the state only
// the path events touch (world position / direction, throughput, radiance, pending NEE term, pixel, counters: 21 words)
// lives in shared memory, one conflict-free column per (lane, ray). A lane steps whichever of its rays can step, so an
// event queue can be left to fill up (thresholds 2-3x higher) without idling its lanes, and the event stages run with
// most lanes active. Same per-path algorithm, draws and sums as k_trace_persistent (and hence as the reference shaders).
#pragma once

#include "vr_trace2.cuh"

namespace vr {

#ifndef VR_DUO_MIN_BLOCKS
#define VR_DUO_MIN_BLOCKS 6      // 80 registers: B200 sweep, 4 / 5 / 6 CTAs per SM = 31.3 / 34.5 / 35.8 Gsamples/s (TF)
#endif
#ifndef VR_DUO_SELECT
#define VR_DUO_SELECT 0          // 1: one step instance per repeat on a selected ray instead of one instance per ray
#endif
#ifndef VR_DUO_STEPS
#define VR_DUO_STEPS 1           // DDA steps of each ray per scheduler pass
#endif
#ifndef VR_DUO_K_COLLIDE_TF
#define VR_DUO_K_COLLIDE_TF 12
#endif
#ifndef VR_DUO_K_COLLIDE
#define VR_DUO_K_COLLIDE 8
#endif
#ifndef VR_DUO_K_EVENT_TF
#define VR_DUO_K_EVENT_TF 16
#endif
#ifndef VR_DUO_K_EVENT
#define VR_DUO_K_EVENT 8
#endif
#ifndef VR_DUO_K_FINISH_TF
#define VR_DUO_K_FINISH_TF 16
#endif
#ifndef VR_DUO_K_FINISH
#define VR_DUO_K_FINISH 8
#endif
#ifndef VR_DUO_MIN_STEP_TF
#define VR_DUO_MIN_STEP_TF 24
#endif
#ifndef VR_DUO_MIN_STEP
#define VR_DUO_MIN_STEP 24
#endif

// register-resident part of a ray
struct HotRay {
    float3 ipos, idir, ri;
    float t, tfar, tau, mip, maj;
    uint32_t seed;
    int stage;
    bool shadow;
};

// shared-memory part of a ray: s_cold[field][ray * VR_TRACE_BLOCK + threadIdx.x]
enum : int { C_POSX, C_POSY, C_POSZ, C_DIRX, C_DIRY, C_DIRZ, C_THRX, C_THRY, C_THRZ, C_LX, C_LY, C_LZ, C_PENDX, C_PENDY, C_PENDZ,
             C_FP, C_TR, C_PIX, C_SJ, C_NPATHS, C_TITEM, C_FLAGS, C_COUNT };
enum : uint32_t { FL_ITEM = 1u, FL_ESCAPED = 2u };

template <bool TF, bool COUNT, class MT>
__global__ void __launch_bounds__(VR_TRACE_BLOCK, VR_DUO_MIN_BLOCKS) k_trace_duo(const __grid_constant__ TraceArgs a) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int W = a.p.resolution[0];
    Cnt<COUNT> cnt;

    __shared__ float s_cold[C_COUNT][2 * VR_TRACE_BLOCK];
    __shared__ float4 s_prep[VR_TRACE_BLOCK / 32][32];     // {view dir, seed after the two jitter draws} of the current block
    float4* prep = s_prep[threadIdx.x >> 5];
#define COLD(f, slot) s_cold[f][slot]
#define COLD3(f, slot) f3(s_cold[f][slot], s_cold[(f) + 1][slot], s_cold[(f) + 2][slot])
#define SET3(f, slot, v) do { const float3 v_ = (v); s_cold[f][slot] = v_.x; s_cold[(f) + 1][slot] = v_.y; s_cold[(f) + 2][slot] = v_.z; } while (0)
#define COLDU(f, slot) __float_as_uint(s_cold[f][slot])
#define SETU(f, slot, v) s_cold[f][slot] = __uint_as_float(v)

    HotRay h0, h1;
    h0.ipos = f3(0.f); h0.idir = f3(1.f); h0.ri = f3(1.f);
    h0.t = 0.f; h0.tfar = -1.f; h0.tau = 0.f; h0.mip = 3.f; h0.maj = 0.f; h0.seed = 0; h0.stage = SG_FINISH; h0.shadow = false;
    h1 = h0;
    SETU(C_FLAGS, threadIdx.x, 0u);
    SETU(C_FLAGS, VR_TRACE_BLOCK + threadIdx.x, 0u);

    // ---- warp state: the current block of 32 samples (one tile, one sample index) ----
    int blk_x0 = 0, blk_y0 = 0, blk_sj = 0;
    int blk_next = 32;
    bool blk_done = false;
    unsigned nxt_block = 0;
    if (lane == 0) nxt_block = atomicAdd(a.job_counter, 1u);
    const unsigned n_jobs = a.n_live ? (__ldg(a.n_live) << a.sample_bits) : unsigned(a.n_jobs);

    // one brick-DDA step of ray h (common.glsl:423-435 / 470-482) and what happens when the ray leaves the volume
    auto step = [&](HotRay& h, const int slot) {
        if (h.stage != SG_STEP) return;
        if (h.t < h.tfar) {
            const float3 curr = h.ipos + h.t * h.idir;
            const int m = round_mip(h.mip);
            cnt.maj();
            h.maj = table_majorant(a, curr, m);
            const float dt = step_dda(curr, h.ri, m);
            h.t += dt;
            h.tau -= h.maj * dt;
            h.mip = fminf(h.mip + 0.25f, 3.f);
            if (!(h.tau > 0.f)) {
                h.t += MT::div(h.tau, h.maj);
                if (!(h.t >= h.tfar)) h.stage = SG_COLLIDE;   // `if (t >= far) break;` (a NaN t goes on to the lookup)
            }
        }
        if (h.stage == SG_STEP && !(h.t < h.tfar)) {
            if (h.shadow) {
                const float Tr = COLD(C_TR, slot);
                if (Tr != 0.f) SET3(C_LX, slot, COLD3(C_LX, slot) + COLD3(C_PENDX, slot) * Tr);
                h.stage = SG_SCATTER;
            } else {
                SETU(C_FLAGS, slot, COLDU(C_FLAGS, slot) | FL_ESCAPED);
                h.stage = SG_FINISH;
            }
        }
    };

    [[maybe_unused]] unsigned iter = 0;
    while (true) {
        // ================= STEP: both rays (independent chains: the two majorant fetches overlap) =================
#if VR_DUO_SELECT
        // ONE step instance per repeat: the lane picks whichever of its rays can step (alternating when both can), so the
        // instance runs with every lane that has a stepping ray instead of two instances with half the lanes each
#pragma unroll
        for (int rep = 0; rep < VR_DUO_STEPS; ++rep) {
            const bool s0 = h0.stage == SG_STEP, s1 = h1.stage == SG_STEP;
            if (s0 | s1) {
                const bool pick1 = s1 && (!s0 || (((iter + rep) & 1u) != 0u));
                HotRay c = pick1 ? h1 : h0;
                step(c, (pick1 ? VR_TRACE_BLOCK : 0) + threadIdx.x);
                if (pick1) { h1.t = c.t; h1.tau = c.tau; h1.mip = c.mip; h1.maj = c.maj; h1.stage = c.stage; }
                else { h0.t = c.t; h0.tau = c.tau; h0.mip = c.mip; h0.maj = c.maj; h0.stage = c.stage; }
            }
        }
        ++iter;
#else
#pragma unroll
        for (int rep = 0; rep < VR_DUO_STEPS; ++rep) {
            step(h0, threadIdx.x);
            step(h1, VR_TRACE_BLOCK + threadIdx.x);
        }
#endif

        // ================= scheduler: lanes that hold a ray in each stage =================
        const unsigned m_step = __ballot_sync(FULL, h0.stage == SG_STEP || h1.stage == SG_STEP);
        const unsigned m_col = __ballot_sync(FULL, h0.stage == SG_COLLIDE || h1.stage == SG_COLLIDE);
        const unsigned m_nee = __ballot_sync(FULL, h0.stage == SG_NEE || h1.stage == SG_NEE);
        const unsigned m_scat = __ballot_sync(FULL, h0.stage == SG_SCATTER || h1.stage == SG_SCATTER);
        const unsigned m_fin = __ballot_sync(FULL, h0.stage == SG_FINISH || h1.stage == SG_FINISH);
        if ((m_step | m_col | m_nee | m_scat | m_fin) == 0u) break;     // every ray idle
        const int n_step = __popc(m_step), n_col = __popc(m_col), n_nee = __popc(m_nee), n_scat = __popc(m_scat), n_fin = __popc(m_fin);
        constexpr int KF = TF ? VR_DUO_K_FINISH_TF : VR_DUO_K_FINISH, K = TF ? VR_DUO_K_EVENT_TF : VR_DUO_K_EVENT;
        constexpr int KC = TF ? VR_DUO_K_COLLIDE_TF : VR_DUO_K_COLLIDE, MIN_STEP = TF ? VR_DUO_MIN_STEP_TF : VR_DUO_MIN_STEP;
        bool run_col = n_col >= KC, run_nee = n_nee >= K, run_scat = n_scat >= K, run_fin = n_fin >= KF;
        if (!(run_col | run_nee | run_scat | run_fin)) {
            if (n_step >= MIN_STEP) continue;                  // keep stepping
            if (n_col >= n_nee && n_col >= n_scat && n_col >= n_fin) run_col = n_col > 0;      // too few lanes can step: drain the fullest queue
            else if (n_nee >= n_scat && n_nee >= n_fin) run_nee = n_nee > 0;
            else if (n_scat >= n_fin) run_scat = n_scat > 0;
            else run_fin = n_fin > 0;
        }
        // at most one of the ray-starting stages per iteration (they share the start code below)
        if (run_nee) { run_scat = false; run_fin = false; }
        else if (run_scat) run_fin = false;

        // ================= COLLIDE: tentative collision (common.glsl:436-452 / 483-498) =================
        if (run_col) {
            const int r = h0.stage == SG_COLLIDE ? 0 : (h1.stage == SG_COLLIDE ? 1 : -1);
            if (r >= 0) {
                HotRay c = r ? h1 : h0;
                const int slot = r * VR_TRACE_BLOCK + threadIdx.x;
                cnt.dens();
                c.stage = SG_STEP;
                const float3 at = c.ipos + c.t * c.idir;
                float d;
                float3 tf_rgb = f3(1.f);
                if (TF) {
                    const float4 rgba = tf_lookup<MT>(a, a.p.vol_density_scale * (MT::decoded ? density_trilinear_decoded(a.density, at) : density_trilinear(a.density, at)) * a.p.vol_inv_majorant);
                    d = a.p.vol_majorant * rgba.w;
                    tf_rgb = f3(rgba.x, rgba.y, rgba.z);
                } else {
                    const int3 tap = stochastic_tricubic_filter<MT>(at, c.seed);
                    d = a.p.vol_density_scale * brick_value(a.density, tap.x, tap.y, tap.z);
                }
                bool parked = false;
                if (!c.shadow) {
                    bool fetched;
                    const float3 em = lookup_emission<MT>(a, at, c.seed, fetched);
                    if (fetched) {
                        cnt.emis();
                        const float3 albedo = f3(a.p.vol_albedo[0], a.p.vol_albedo[1], a.p.vol_albedo[2]);
                        SET3(C_LX, slot, COLD3(C_LX, slot) + COLD3(C_THRX, slot) * (f3(1.f) - albedo) * em * d * a.p.vol_inv_majorant);
                    }
                    if (rng(c.seed) * c.maj < d) {          // real collision: the segment ends here (common.glsl:490-496)
                        float3 thr = COLD3(C_THRX, slot) * f3(a.p.vol_albedo[0], a.p.vol_albedo[1], a.p.vol_albedo[2]);
                        if (TF) thr = thr * tf_rgb;
                        SET3(C_THRX, slot, thr);
                        c.stage = SG_NEE;
                        parked = true;
                    }
                } else {
                    if (rng(c.seed) * c.maj < d) {          // common.glsl:442-450
                        float Tr = COLD(C_TR, slot) * fmaxf(0.f, 1.f - MT::div(a.p.vol_majorant, c.maj));
                        if (Tr < .1f) {
                            const float prob = 1 - Tr;
                            if (rng(c.seed) < prob) { Tr = 0.f; c.stage = SG_SCATTER; parked = true; }   // absorbed: nothing is added to L
                            else Tr = MT::div(Tr, 1 - prob);
                        }
                        COLD(C_TR, slot) = Tr;
                    }
                }
                if (!parked) {
                    c.tau = -MT::log(1.f - rng(c.seed));
                    c.mip = fmaxf(0.f, c.mip - 2.f);
                }
                if (r) { h1.seed = c.seed; h1.stage = c.stage; h1.tau = c.tau; h1.mip = c.mip; }
                else { h0.seed = c.seed; h0.stage = c.stage; h0.tau = c.tau; h0.mip = c.mip; }
            }
        }

        // ================= the ray-starting stages: NEE | SCATTER | FINISH =================
        const int X = run_nee ? SG_NEE : (run_scat ? SG_SCATTER : (run_fin ? SG_FINISH : -1));
        if (X < 0) continue;
        const int r = h0.stage == X ? 0 : (h1.stage == X ? 1 : -1);
        const int slot = (r > 0 ? VR_TRACE_BLOCK : 0) + threadIdx.x;
        HotRay c = r > 0 ? h1 : h0;
        bool start = false;     // the lane starts a new ray from (pos, rd) below
        float3 pos = f3(0.f), rd = f3(0.f, 0.f, -1.f);

        // ---- NEE: real collision -> next-event estimation (common.glsl:611-626) ----
        if (X == SG_NEE && r >= 0) {
            cnt.real();
            const float3 dir = COLD3(C_DIRX, slot);
            pos = COLD3(C_POSX, slot) + c.t * dir;
            SET3(C_POSX, slot, pos);
            float3 w_i;
            const float r0 = rng(c.seed), r1 = rng(c.seed);
            cnt.nee();
            const float4 Le_pdf = sample_environment<MT>(a, r0, r1, w_i);
            if (Le_pdf.w > 0) {
                const float f_p = phase_hg<MT>(dot(-dir, w_i), a.p.vol_phase_g);
                const float mis_weight = a.p.show_environment > 0 ? MT::div(sqr(Le_pdf.w), sqr(Le_pdf.w) + sqr(f_p)) : 1.f;
                const float3 cc = COLD3(C_THRX, slot) * mis_weight * f_p * f3(Le_pdf.x, Le_pdf.y, Le_pdf.z);
                SET3(C_PENDX, slot, f3(MT::div(cc.x, Le_pdf.w), MT::div(cc.y, Le_pdf.w), MT::div(cc.z, Le_pdf.w)));
                COLD(C_FP, slot) = f_p;
                COLD(C_TR, slot) = 1.f;
                c.shadow = true;
                rd = w_i; start = true;
                c.stage = SG_STEP;
            } else {
                c.stage = SG_SCATTER;
            }
        }

        // ---- SCATTER: bounce limit, Russian roulette, phase sampling (common.glsl:628-641) ----
        if (X == SG_SCATTER && r >= 0) {
            bool end = false;
            const uint32_t n_paths = COLDU(C_NPATHS, slot) + 1u;
            SETU(C_NPATHS, slot, n_paths);
            if (n_paths >= uint32_t(a.p.bounces)) end = true;
            else {
                float3 thr = COLD3(C_THRX, slot);
                const float rr_val = luma(thr);
                if (rr_val < .1f) {
                    const float prob = 1 - rr_val;
                    if (rng(c.seed) < prob) end = true;
                    else { const float k = 1 - prob; thr = f3(MT::div(thr.x, k), MT::div(thr.y, k), MT::div(thr.z, k)); SET3(C_THRX, slot, thr); }
                }
            }
            if (end) {
                SETU(C_FLAGS, slot, COLDU(C_FLAGS, slot) & ~FL_ESCAPED);
                c.stage = SG_FINISH;
            } else {
                const float3 dir = COLD3(C_DIRX, slot);
                const float s0 = rng(c.seed), s1 = rng(c.seed);
                const float3 scatter_dir = sample_phase_hg<MT>(dir, a.p.vol_phase_g, s0, s1);
                COLD(C_FP, slot) = phase_hg<MT>(dot(-dir, scatter_dir), a.p.vol_phase_g);
                SET3(C_DIRX, slot, scatter_dir);
                c.shadow = false;
                pos = COLD3(C_POSX, slot);
                rd = scatter_dir; start = true;
                c.stage = SG_STEP;
            }
        }

        // ---- FINISH: environment on escape, store the sample, take the next one (warp-uniform: cooperative block switch) ----
        if (X == SG_FINISH) {
            const bool mine = r >= 0;
            bool have_item = false;
            if (mine && (COLDU(C_FLAGS, slot) & FL_ITEM)) {
                const uint32_t flags = COLDU(C_FLAGS, slot);
                float3 L = COLD3(C_LX, slot);
                const uint32_t n_paths = COLDU(C_NPATHS, slot);
                if ((flags & FL_ESCAPED) && a.p.show_environment > 0) {      // common.glsl:644-649
                    cnt.env();
                    const float3 Le = lookup_environment(a, COLD3(C_DIRX, slot));
                    const float pe = pdf_environment<MT>(a, Le);
                    const float f_p = COLD(C_FP, slot);
                    const float mis_weight = n_paths > 0 ? MT::div(sqr(f_p), sqr(f_p) + sqr(pe)) : 1.f;
                    L = L + COLD3(C_THRX, slot) * mis_weight * Le;
                }
                cnt.samp();                                     // pathtracer_brick.glsl:36: sanitize(L), folded by k_fold
                const uint32_t pix = COLDU(C_PIX, slot);
                const int px = int(pix & 0xffffu), py = int(pix >> 16), sj = int(COLDU(C_SJ, slot));
                VR_LBUF_STORE(a.lbuf + size_t(sj) * a.lbuf_stride + size_t(py) * W + px,
                              make_float4(sanitize(L.x), sanitize(L.y), sanitize(L.z), sanitize(fminf(float(n_paths), 1.f))));
                SETU(C_FLAGS, slot, 0u);
                if (a.tile_cost && ((px ^ py ^ sj) & 3) == 0) {   // a dithered quarter of the samples is enough to rank tiles
                    unsigned now;
                    asm volatile("mov.u32 %0, %%clock;" : "=r"(now));
                    atomicAdd(a.tile_cost + ((py - a.y0) >> 2) * a.tiles_x + ((px - a.x0) >> 3), (now - COLDU(C_TITEM, slot)) >> 8);
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
                        if (want) { want = false; c.stage = SG_IDLE; }
                        break;
                    }
                    blk_sj = int(b & ((1u << a.sample_bits) - 1u));
                    if (blk_sj >= a.n_samples) continue;           // padding of a non-power-of-two sample count
                    const unsigned txy = __ldg(a.tile_order + (b >> a.sample_bits));
                    blk_x0 = a.x0 + int(txy & 0xffffu) * 8;
                    blk_y0 = a.y0 + int(txy >> 16) * 4;
                    const int ix = blk_x0 + (lane & 7), iy = blk_y0 + (lane >> 3);
                    // all 32 lanes prepare the block (pathtracer_brick.glsl:28-30: TEA seed, two jitter draws, view direction)
                    uint32_t sd = tea32(uint32_t(a.p.seed) * uint32_t(iy * W + ix), uint32_t(a.first_sample + blk_sj));
                    const float jx = rng(sd), jy = rng(sd);
                    const float3 vd = view_dir<MT>(a, ix, iy, jx, jy);
                    __syncwarp();
                    prep[lane] = make_float4(vd.x, vd.y, vd.z, __uint_as_float(sd));
                    __syncwarp();
                    blk_next = 0;
                }
                const int i = blk_next + __popc(m_want & ((1u << lane) - 1u));
                if (want && i < 32) {
                    const int px = blk_x0 + (i & 7), py = blk_y0 + (i >> 3);
                    if (px < a.x1 && py < a.y1) {
                        const float4 pr = prep[i];
                        SETU(C_PIX, slot, uint32_t(px) | (uint32_t(py) << 16));
                        SETU(C_SJ, slot, uint32_t(blk_sj));
                        c.seed = __float_as_uint(pr.w);
                        rd = f3(pr.x, pr.y, pr.z);
                        have_item = true;
                        want = false;
                    }
                }
                blk_next += __popc(m_want);
            }
            if (mine && have_item) {   // new sample: camera ray of the prepared sample
                unsigned now;
                asm volatile("mov.u32 %0, %%clock;" : "=r"(now));
                SETU(C_TITEM, slot, now);
                pos = f3(a.p.cam_pos[0], a.p.cam_pos[1], a.p.cam_pos[2]);
                SET3(C_POSX, slot, pos);
                SET3(C_DIRX, slot, rd);
                SET3(C_THRX, slot, f3(1.f));
                SET3(C_LX, slot, f3(0.f));
                SETU(C_NPATHS, slot, 0u);
                COLD(C_FP, slot) = 0.f;
                SETU(C_FLAGS, slot, FL_ITEM);
                c.shadow = false;
                start = true;
                c.stage = SG_STEP;
            }
        }

        // ---- start the new ray: clip + world->index + first free-flight draw (common.glsl:459-468 / 413-421) ----
        if (start) {
            float tn, tf;
            if (intersect_box<MT>(pos, rd, a.p.vol_bb_min, a.p.vol_bb_max, tn, tf)) {
                const Mat4& M = *reinterpret_cast<const Mat4*>(a.p.vol_density_inv_transform);
                c.ipos = mul_point(M, pos);
                c.idir = mul_dir(M, rd);
                c.ri = f3(MT::rcp(c.idir.x), MT::rcp(c.idir.y), MT::rcp(c.idir.z));
                c.t = tn + 1e-6f;
                c.tfar = tf;
                c.tau = -MT::log(1.f - rng(c.seed));
                c.mip = 3.f;
            } else {
                c.t = 0.f; c.tfar = -1.f;   // missed the box: the ray "ends" at once (Tr = 1 / escape)
            }
        }
        if (r > 0) h1 = c; else if (r == 0) h0 = c;
    }
    flush_counters(a, cnt);
#undef COLD
#undef COLD3
#undef SET3
#undef COLDU
#undef SETU
}

}  // namespace vr
