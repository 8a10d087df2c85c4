// Ray-pool tracking kernel (production path of vrb_trace since round 2).
//
// Per path it executes exactly the arithmetic and the random-number order of k_trace_persistent (vr_trace2.cuh) -- the images
// are bit-identical -- but the lanes of a warp no longer OWN a path. The lane-resident kernel ran its stages at 12-17 of 32
// active lanes (profiles/r01_v11_trace_*_summary.txt): a lane whose path waited for a rare event (NEE, scatter, finish +
// regeneration) was lost to the hot brick-DDA loop, and every threshold / occupancy / two-rays-per-lane variant ended within
// a few percent, because 32 paths simply are asynchronous. Here every warp keeps a POOL of VR_POOL_SLOTS (48) path states in
// shared memory, struct-of-arrays so that 32 lanes touching 32 different slots are (at worst two-way) bank-conflict free, and
// a scheduler iteration
//   1. picks the stage class with the most waiting slots -- SEGMENT (a ray that steps or sits at a tentative collision), NEE,
//      SCATTER or FINISH -- from per-class counts that every stage keeps up to date with a few ballots,
//   2. hands the first 32 slots of that class to the lanes (rank by popcount -> slot list in shared memory),
//   3. runs that ONE stage with up to 32 lanes: the lanes load the fields the stage reads, run the code of the lane-resident
//      kernel, and store what changed.
// SEGMENT is the hot loop and stays in REGISTERS: the 32 lanes step their rays (brick DDA) and resolve tentative collisions
// in batches (a lane at a collision waits until VR_SEG_K_COLLIDE lanes are, like the in-warp queues of round 1) until fewer
// than VR_SEG_EXIT lanes still have a live ray -- lanes only drop out at a REAL collision or at the end of a ray, ~10x rarer
// than a step or a null collision -- and only then is the state written back and the next class scheduled. So the pool pays
// its ~25 instructions of scheduling + ~20 loads/stores once per many DDA events, the event stages run with a full warp, and
// the hot loop starts every visit with 32 live rays. History: v1 scheduled every DDA visit through the pool and recounted
// the stages with ten ballots per iteration (29 % of the issued instructions, 27 active lanes but no faster:
// profiles/r02_pool_v1_*); v2 kept the counts incrementally (-6 %); v3 is the register-resident SEGMENT loop.
// FINISH also regenerates: finished slots take the next samples of the warp's current block of 32 (tile, sample index)
// exactly like the lane-resident kernel (all 32 lanes prepare a block together: TEA seed, jitter, view direction).
#pragma once

#include "vr_trace2.cuh"

namespace vr {

#ifndef VR_POOL_SLOTS
#define VR_POOL_SLOTS 48          // path states per warp (multiple of 16). B200 sweep (profiles/r02_pool_v4_sweep.txt): 32 slots lose 25 % (the
                                  // event stages starve), 64 slots x 6 CTAs/SM and 48 x 8 are within 2 %, 48 x 8 with VR_SEG_EXIT 16 is best on all four configs
#endif
#ifndef VR_POOL_WARPS
#define VR_POOL_WARPS 4           // warps per CTA
#endif
#ifndef VR_POOL_MIN_BLOCKS
#define VR_POOL_MIN_BLOCKS 8      // CTAs per SM the register allocation must allow: 64 registers (shared memory: 8 x 26.5 KiB at 48 slots)
#endif
#ifndef VR_SEG_FULL
#define VR_SEG_FULL 32            // SEGMENT runs at once when this many slots wait for it
#endif
#ifndef VR_SEG_EXIT
#define VR_SEG_EXIT 16            // leave the SEGMENT loop when fewer of its lanes still have a live ray (12 ... 28 are within 4 %)
#endif
#ifndef VR_SEG_MAX_ITERS
#define VR_SEG_MAX_ITERS 64       // ... or after this many loop iterations (bounds the time a waiting event stage is starved)
#endif
#ifndef VR_SEG_K_COLLIDE
#define VR_SEG_K_COLLIDE 8        // resolve tentative collisions once this many lanes wait at one (non-TF: cheap 1-tap collision)
#endif
#ifndef VR_SEG_K_COLLIDE_TF
#define VR_SEG_K_COLLIDE_TF 8     // TF: 8-tap trilinear + LUT collision
#endif
#ifndef VR_SEG_MIN_STEP
#define VR_SEG_MIN_STEP 12        // fewer stepping lanes than this: resolve the waiting collisions even below the threshold
#endif

// fields of a slot (one 32-bit word each; PF_COUNT words per slot)
enum : int {
    PF_STAGE = 0,   // stage | flags | STEP visits of the current ray
    PF_PIX /* y << 16 | x */, PF_SJ, PF_SEED, PF_NPATHS,
    PF_POS, PF_DIR = PF_POS + 3, PF_THR = PF_DIR + 3, PF_L = PF_THR + 3, PF_PEND = PF_L + 3,
    PF_FP = PF_PEND + 3, PF_TR,
    PF_IPOS, PF_IDIR = PF_IPOS + 3, PF_T = PF_IDIR + 3, PF_TFAR,
    PF_TAU,         // optical depth left to the next tentative collision; while the ray SITS at one (stage COLLIDE): its majorant
    PF_MIP,
    PF_COUNT
};
// flags in the stage word
enum : uint32_t { PL_STAGE = 7u, PL_SHADOW = 8u, PL_ESCAPED = 16u, PL_ITEM = 32u, PL_LIT = 64u,
                  PL_FLAGS = 255u, PL_VISIT = 256u, PL_VISIT_MAX = 0xfff00000u };     // bits 8..31: SEGMENT visits of the current ray (hang guard)

constexpr size_t pool_warp_words() { return size_t(PF_COUNT) * VR_POOL_SLOTS + 32 /* slot list */ + 128 /* prepared block: 32 x float4 */; }
constexpr size_t pool_smem_bytes() { return pool_warp_words() * 4 * VR_POOL_WARPS; }

template <bool TF, bool COUNT, class MT>
__global__ void __launch_bounds__(VR_POOL_WARPS * 32, VR_POOL_MIN_BLOCKS) k_trace_pool(const __grid_constant__ TraceArgs a) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int SL = VR_POOL_SLOTS, NS = (VR_POOL_SLOTS + 31) / 32;
    static_assert(VR_POOL_SLOTS % 16 == 0 && VR_POOL_SLOTS >= 32, "the last row of 32 slots may be half full (48, 80 ... slots)");
    static_assert((pool_warp_words() * 4) % 16 == 0, "float4 alignment of the prepared block");
    extern __shared__ float4 pool_smem[];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int W = a.p.resolution[0];
    Cnt<COUNT> cnt;

    float* sf = reinterpret_cast<float*>(pool_smem) + size_t(threadIdx.x >> 5) * pool_warp_words();
    uint32_t* su = reinterpret_cast<uint32_t*>(sf);
    uint32_t* sel = su + size_t(PF_COUNT) * SL;
    float4* prep = reinterpret_cast<float4*>(sf + size_t(PF_COUNT) * SL + 32);
#define PF(f, s) sf[(f) * SL + (s)]
#define PU(f, s) su[(f) * SL + (s)]
#define PLOAD3(f, s) f3(PF((f), s), PF((f) + 1, s), PF((f) + 2, s))
#define PSTORE3(f, s, v) do { const float3 v_ = (v); PF((f), s) = v_.x; PF((f) + 1, s) = v_.y; PF((f) + 2, s) = v_.z; } while (0)

#pragma unroll
    for (int k = 0; k < NS; ++k) if (k * 32 + lane < SL) PU(PF_STAGE, k * 32 + lane) = SG_FINISH;      // every slot starts by asking for a sample

    // ---- warp state: the current block of 32 samples (one tile, one sample index), prepared in shared memory ----
    int blk_x0 = 0, blk_y0 = 0, blk_sj = 0;
    int blk_next = 32;             // next sample of the block to hand out (>= 32: block used up)
    bool blk_done = false;         // the counter is exhausted
    unsigned nxt_block = 0;        // lane 0: id of the prefetched next block
    if (lane == 0) nxt_block = atomicAdd(a.job_counter, 1u);
    const unsigned n_jobs = a.n_live ? (__ldg(a.n_live) << a.sample_bits) : unsigned(a.n_jobs);
    // slots per stage class (warp-uniform, kept up to date by every stage from the transitions of the slots it processed);
    // n_seg counts the slots in stage STEP or COLLIDE
    int n_seg = 0, n_nee = 0, n_scat = 0, n_fin = SL;

    while (true) {
        __syncwarp();
        // ================= scheduler: pick the class with the most waiting slots, hand its slots to the lanes =================
        if ((n_seg | n_nee | n_scat | n_fin) == 0) break;      // every slot idle
        int S = SG_STEP, best = n_seg;
        if (n_seg < VR_SEG_FULL) {
            if (n_nee > best) { best = n_nee; S = SG_NEE; }
            if (n_scat > best) { best = n_scat; S = SG_SCATTER; }
            if (n_fin > best) { best = n_fin; S = SG_FINISH; }
        }
        int base = 0;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const bool have = (k + 1) * 32 <= SL || k * 32 + lane < SL;       // compile-time true for full rows
            const uint32_t st = have ? (PU(PF_STAGE, k * 32 + lane) & PL_STAGE) : uint32_t(SG_IDLE);
            const bool in = S == SG_STEP ? st <= uint32_t(SG_COLLIDE) : st == uint32_t(S);      // SG_STEP = 0, SG_COLLIDE = 1
            const unsigned b = __ballot_sync(FULL, in);
            const int r = base + __popc(b & lt);
            if (in && r < 32) sel[r] = uint32_t(k * 32 + lane);
            base += __popc(b);
        }
        __syncwarp();
        const int nsel = min(base, 32);
        const bool act = lane < nsel;
        const int slot = act ? int(sel[lane]) : 0;

        bool start = false;            // this lane starts a new ray from (spos, rd) with the stream `seed` below
        float3 spos = f3(0.f), rd = f3(0.f, 0.f, -1.f);
        uint32_t seed = 0, fl = 0;

        if (S == SG_STEP) {
            // ================= SEGMENT: brick-DDA steps + tentative collisions of up to 32 rays, in registers =================
            // (common.glsl:423-452 / 470-498; one loop body for camera segments and shadow rays)
            float3 ipos = f3(0.f), idir = f3(1.f), ri = f3(1.f);
            float t = 0.f, tfar = -1.f, tau = 0.f, mip = 3.f, majorant = 0.f;
            int sub = SG_IDLE;         // SG_STEP: stepping, SG_COLLIDE: waits at a tentative collision, other: left the loop (next stage)
            if (act) {
                fl = PU(PF_STAGE, slot);
                sub = int(fl & PL_STAGE);
                ipos = PLOAD3(PF_IPOS, slot); idir = PLOAD3(PF_IDIR, slot);
                t = PF(PF_T, slot); tfar = PF(PF_TFAR, slot); mip = PF(PF_MIP, slot);
                seed = PU(PF_SEED, slot);
                const float tm = PF(PF_TAU, slot);
                if (sub == SG_COLLIDE) majorant = tm; else tau = tm;
                fl += PL_VISIT;                                  // hang guard: SEGMENT visits of this ray (>= 1 DDA event each)
                if (fl >= PL_VISIT_MAX) { t = INFINITY; sub = SG_STEP; }      // no finite ray gets here: end it
                ri = f3(MT::rcp(idir.x), MT::rcp(idir.y), MT::rcp(idir.z));
            }
            const bool shadow = (fl & PL_SHADOW) != 0u;
            const int sub_left = shadow ? SG_SCATTER : SG_FINISH;      // where a ray goes when it leaves the volume
            constexpr int KC = TF ? VR_SEG_K_COLLIDE_TF : VR_SEG_K_COLLIDE;
            const int waiting = (n_seg - nsel) + n_nee + n_scat + n_fin;      // slots of the pool this visit does not work on
#pragma unroll 1
            for (int iter = 0; iter < VR_SEG_MAX_ITERS; ++iter) {
                // ---- STEP ----
                // two steps whose majorant loads are in flight TOGETHER: the second step's position t + dt and level do not depend on
                // the first step's majorant (dt is geometry), only on whether the first step ends in a tentative collision -- so its
                // table address is computed and its load issued before the first majorant is consumed; a collision / exit in the
                // first step just drops it. Same operations per ray in the same order (bit-identical); the fetch latency of a
                // loop iteration halves: +0.5 ... +1.5 % on the four configs against two plain steps (profiles/r02_pool_v4_sweep.txt).
                if (sub == SG_STEP) {
                    if (t < tfar) {
                        const float3 curr_a = ipos + t * idir;
                        const int m_a = round_mip(mip);
                        const float maj_a = table_majorant_idx(a, curr_a, m_a);
                        const float dt_a = step_dda(curr_a, ri, m_a);
                        const float t_a = t + dt_a, mip_a = fminf(mip + 0.25f, 3.f);
                        const float3 curr_b = ipos + t_a * idir;
                        const int m_b = round_mip(mip_a);
                        const float maj_b = table_majorant_idx(a, curr_b, m_b);          // speculative: issued before maj_a is used
                        cnt.maj();
                        majorant = maj_a;
                        t = t_a;
                        tau -= maj_a * dt_a;
                        mip = mip_a;
                        if (!(tau > 0.f)) {
                            t += MT::div(tau, maj_a);
                            if (!(t >= tfar)) sub = SG_COLLIDE;   // the reference tests `if (t >= far) break;` (a NaN t goes on to the lookup)
                        } else if (t < tfar) {
                            const float dt_b = step_dda(curr_b, ri, m_b);
                            cnt.maj();
                            majorant = maj_b;
                            t += dt_b;
                            tau -= maj_b * dt_b;
                            mip = fminf(mip + 0.25f, 3.f);
                            if (!(tau > 0.f)) {
                                t += MT::div(tau, maj_b);
                                if (!(t >= tfar)) sub = SG_COLLIDE;
                            }
                        } else {
                            sub = sub_left;
                        }
                    } else {
                        sub = sub_left;      // `while (t < far)` ended: the ray left the volume
                    }
                }
                // ---- COLLIDE (batched) ----
                const int n_col = __popc(__ballot_sync(FULL, sub == SG_COLLIDE)), n_stp = __popc(__ballot_sync(FULL, sub == SG_STEP));
                if (n_col >= KC || (n_col > 0 && n_stp < VR_SEG_MIN_STEP)) {
                    if (sub == SG_COLLIDE) {
                        cnt.dens();
                        sub = SG_STEP;
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
                        if (!shadow) {
                            bool fetched;
                            const float3 em = lookup_emission<MT>(a, at, seed, fetched);
                            if (fetched) {
                                cnt.emis();
                                const float3 albedo = f3(a.p.vol_albedo[0], a.p.vol_albedo[1], a.p.vol_albedo[2]);
                                PSTORE3(PF_L, slot, PLOAD3(PF_L, slot) + PLOAD3(PF_THR, slot) * (f3(1.f) - albedo) * em * d * a.p.vol_inv_majorant);
                            }
                            if (rng(seed) * majorant < d) {          // real collision: the segment ends here (common.glsl:490-496)
                                float3 thr = PLOAD3(PF_THR, slot) * f3(a.p.vol_albedo[0], a.p.vol_albedo[1], a.p.vol_albedo[2]);
                                if (TF) thr = thr * tf_rgb;
                                PSTORE3(PF_THR, slot, thr);
                                sub = SG_NEE;
                            }
                        } else {
                            if (rng(seed) * majorant < d) {          // common.glsl:442-450
                                float Tr = PF(PF_TR, slot);
                                Tr *= fmaxf(0.f, 1.f - MT::div(a.p.vol_majorant, majorant));
                                if (Tr < .1f) {
                                    const float prob = 1 - Tr;
                                    if (rng(seed) < prob) { Tr = 0.f; sub = SG_SCATTER + 8; }   // absorbed: `return 0.f`, nothing is added to L
                                    else Tr = MT::div(Tr, 1 - prob);
                                }
                                PF(PF_TR, slot) = Tr;
                            }
                        }
                        if (sub == SG_STEP) {
                            tau = -MT::log(1.f - rng(seed));
                            mip = fmaxf(0.f, mip - 2.f);
                        }
                    }
                }
                // leave once the lanes freed by finished rays can be refilled from the pool (at the tail of a launch the pool
                // runs dry and the loop simply keeps its rays)
                const int live = __popc(__ballot_sync(FULL, sub <= SG_COLLIDE));
                if (live == 0 || (live < VR_SEG_EXIT && waiting + (nsel - live) >= 32 - live)) break;
            }
            // ---- write back, count the transitions ----
            const bool absorbed = sub == SG_SCATTER + 8;
            if (absorbed) sub = SG_SCATTER;
            if (act) {
                PF(PF_T, slot) = t; PF(PF_MIP, slot) = mip;
                PF(PF_TAU, slot) = sub == SG_COLLIDE ? majorant : tau;
                PU(PF_SEED, slot) = seed;
                uint32_t f = (fl & ~(PL_STAGE | PL_LIT | PL_ESCAPED)) | uint32_t(sub);
                if (sub == SG_SCATTER && !absorbed) f |= PL_LIT;    // shadow ray got through: SCATTER adds pend * Tr first
                if (sub == SG_FINISH) f |= PL_ESCAPED;              // camera segment left the volume: environment on escape
                PU(PF_STAGE, slot) = f;
            }
            {
                const int c_nee = __popc(__ballot_sync(FULL, act && sub == SG_NEE)), c_scat = __popc(__ballot_sync(FULL, act && sub == SG_SCATTER)),
                          c_fin = __popc(__ballot_sync(FULL, act && sub == SG_FINISH));
                n_seg -= c_nee + c_scat + c_fin; n_nee += c_nee; n_scat += c_scat; n_fin += c_fin;
            }
        } else if (S == SG_NEE) {
            // ================= NEE: real collision -> next-event estimation (common.glsl:611-626) =================
            if (act) {
                cnt.real();
                fl = PU(PF_STAGE, slot);
                seed = PU(PF_SEED, slot);
                const float3 dir = PLOAD3(PF_DIR, slot);
                spos = PLOAD3(PF_POS, slot) + PF(PF_T, slot) * dir;
                PSTORE3(PF_POS, slot, spos);
                float3 w_i;
                const float r0 = rng(seed), r1 = rng(seed);
                cnt.nee();
                const float4 Le_pdf = sample_environment<MT>(a, r0, r1, w_i);
                if (Le_pdf.w > 0) {
                    const float f_p = phase_hg<MT>(dot(-dir, w_i), a.p.vol_phase_g);
                    const float mis_weight = a.p.show_environment > 0 ? MT::div(sqr(Le_pdf.w), sqr(Le_pdf.w) + sqr(f_p)) : 1.f;
                    // L += throughput * mis_weight * f_p * Tr * Le / pdf, with Tr applied when the shadow ray is done
                    const float3 c = PLOAD3(PF_THR, slot) * mis_weight * f_p * f3(Le_pdf.x, Le_pdf.y, Le_pdf.z);
                    PSTORE3(PF_PEND, slot, f3(MT::div(c.x, Le_pdf.w), MT::div(c.y, Le_pdf.w), MT::div(c.z, Le_pdf.w)));
                    PF(PF_FP, slot) = f_p;
                    PF(PF_TR, slot) = 1.f;
                    fl = (fl & PL_FLAGS & ~(PL_STAGE | PL_LIT)) | PL_SHADOW | uint32_t(SG_STEP);     // new ray: visit counter back to 0
                    rd = w_i; start = true;
                } else {
                    PU(PF_SEED, slot) = seed;
                    fl = (fl & ~(PL_STAGE | PL_LIT)) | uint32_t(SG_SCATTER);
                }
                PU(PF_STAGE, slot) = fl;
            }
            {
                const int c_seg = __popc(__ballot_sync(FULL, start));
                n_nee -= nsel; n_seg += c_seg; n_scat += nsel - c_seg;
            }
        } else if (S == SG_SCATTER) {
            // ================= SCATTER: bounce limit, Russian roulette, phase sampling (common.glsl:628-641) =================
            if (act) {
                fl = PU(PF_STAGE, slot);
                if (fl & PL_LIT) {                               // the shadow ray reached the environment (deferred from SEGMENT)
                    const float Tr = PF(PF_TR, slot);
                    if (Tr != 0.f) PSTORE3(PF_L, slot, PLOAD3(PF_L, slot) + PLOAD3(PF_PEND, slot) * Tr);     // (x * 0) stays 0 even if pend overflowed, as in the reference's order
                }
                seed = PU(PF_SEED, slot);
                uint32_t n_paths = PU(PF_NPATHS, slot);
                bool end = false;
                if (++n_paths >= uint32_t(a.p.bounces)) end = true;
                else {
                    float3 thr = PLOAD3(PF_THR, slot);
                    const float rr_val = luma(thr);
                    if (rr_val < .1f) {
                        const float prob = 1 - rr_val;
                        if (rng(seed) < prob) end = true;
                        else { const float k = 1 - prob; PSTORE3(PF_THR, slot, f3(MT::div(thr.x, k), MT::div(thr.y, k), MT::div(thr.z, k))); }
                    }
                }
                PU(PF_NPATHS, slot) = n_paths;
                if (end) {
                    PU(PF_SEED, slot) = seed;
                    fl = (fl & ~(PL_STAGE | PL_LIT | PL_ESCAPED | PL_SHADOW)) | uint32_t(SG_FINISH);
                } else {
                    const float3 dir = PLOAD3(PF_DIR, slot);
                    const float s0 = rng(seed), s1 = rng(seed);
                    const float3 scatter_dir = sample_phase_hg<MT>(dir, a.p.vol_phase_g, s0, s1);
                    PF(PF_FP, slot) = phase_hg<MT>(dot(-dir, scatter_dir), a.p.vol_phase_g);
                    PSTORE3(PF_DIR, slot, scatter_dir);
                    fl = (fl & PL_FLAGS & ~(PL_STAGE | PL_LIT | PL_SHADOW)) | uint32_t(SG_STEP);
                    spos = PLOAD3(PF_POS, slot);
                    rd = scatter_dir; start = true;
                }
                PU(PF_STAGE, slot) = fl;
            }
            {
                const int c_seg = __popc(__ballot_sync(FULL, start));
                n_scat -= nsel; n_seg += c_seg; n_fin += nsel - c_seg;
            }
        } else {
            // ================= FINISH: environment on escape, store the sample, take the next one =================
            bool want = act;
            if (act) {
                fl = PU(PF_STAGE, slot);
                if (fl & PL_ITEM) {
                    float3 Lf = PLOAD3(PF_L, slot);
                    const uint32_t n_paths = PU(PF_NPATHS, slot);
                    if ((fl & PL_ESCAPED) && a.p.show_environment > 0) {      // common.glsl:644-649
                        cnt.env();
                        const float3 Le = lookup_environment(a, PLOAD3(PF_DIR, slot));
                        const float pe = pdf_environment<MT>(a, Le);
                        const float f_p = PF(PF_FP, slot);
                        const float mis_weight = n_paths > 0 ? MT::div(sqr(f_p), sqr(f_p) + sqr(pe)) : 1.f;
                        Lf = Lf + PLOAD3(PF_THR, slot) * mis_weight * Le;
                    }
                    cnt.samp();                                     // pathtracer_brick.glsl:36: sanitize(L), folded by k_fold
                    const uint32_t sj = PU(PF_SJ, slot), pxy = PU(PF_PIX, slot), px = pxy & 0xffffu, py = pxy >> 16;     // (no division by W)
                    VR_LBUF_STORE(a.lbuf + size_t(sj) * a.lbuf_stride + (py * uint32_t(W) + px),
                                  make_float4(sanitize(Lf.x), sanitize(Lf.y), sanitize(Lf.z), sanitize(fminf(float(n_paths), 1.f))));
                    if (a.tile_cost) {     // what ranks the tiles for the next launch of this view: path vertices of a dithered quarter of the samples
                        if (((px ^ py ^ sj) & 3u) == 0u)
                            atomicAdd(a.tile_cost + ((int(py) - a.y0) >> 2) * a.tiles_x + ((int(px) - a.x0) >> 3), 1u + min(n_paths, 4095u));
                    }
                }
            }
            int px = 0, py = 0, sj = 0;
            bool have_item = false;
            float3 dir = f3(0.f, 0.f, -1.f);
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
                        want = false;
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
                const int i = blk_next + __popc(m_want & lt);
                if (want && i < 32) {
                    px = blk_x0 + (i & 7);
                    py = blk_y0 + (i >> 3);
                    if (px < a.x1 && py < a.y1) {
                        const float4 pr = prep[i];
                        sj = blk_sj;                               // (the loop may switch to the next block for the other lanes)
                        seed = __float_as_uint(pr.w);
                        dir = f3(pr.x, pr.y, pr.z);
                        have_item = true;
                        want = false;
                    }
                }
                blk_next += __popc(m_want);
            }
            if (act) {
                if (have_item) {                                   // new sample: camera ray of the prepared sample
                    PU(PF_PIX, slot) = (uint32_t(py) << 16) | uint32_t(px);      // image sides < 65536 (checked by the host)
                    PU(PF_SJ, slot) = uint32_t(sj);
                    PU(PF_NPATHS, slot) = 0u;
                    spos = f3(a.p.cam_pos[0], a.p.cam_pos[1], a.p.cam_pos[2]);
                    PSTORE3(PF_POS, slot, spos);
                    PSTORE3(PF_DIR, slot, dir);
                    PSTORE3(PF_THR, slot, f3(1.f));
                    PSTORE3(PF_L, slot, f3(0.f));
                    PF(PF_FP, slot) = 0.f;
                    PU(PF_STAGE, slot) = PL_ITEM | uint32_t(SG_STEP);
                    rd = dir; start = true;
                } else {
                    PU(PF_STAGE, slot) = uint32_t(SG_IDLE);        // the counter is exhausted
                }
            }
            {
                const int c_seg = __popc(__ballot_sync(FULL, start));
                n_fin -= nsel; n_seg += c_seg;
            }
        }

        // ================= start the new rays: clip + world->index + first free-flight draw (common.glsl:459-468 / 413-421) =================
        if (start) {
            float tn, tf;
            if (intersect_box<MT>(spos, rd, a.p.vol_bb_min, a.p.vol_bb_max, tn, tf)) {
                const Mat4& M = *reinterpret_cast<const Mat4*>(a.p.vol_density_inv_transform);
                PSTORE3(PF_IPOS, slot, mul_point(M, spos));
                PSTORE3(PF_IDIR, slot, mul_dir(M, rd));
                PF(PF_T, slot) = tn + 1e-6f;
                PF(PF_TFAR, slot) = tf;
                PF(PF_TAU, slot) = -MT::log(1.f - rng(seed));
                PF(PF_MIP, slot) = 3.f;
            } else {
                PF(PF_T, slot) = 0.f; PF(PF_TFAR, slot) = -1.f;   // missed the box: the ray "ends" at once (Tr = 1 / escape)
            }
            PU(PF_SEED, slot) = seed;
        }
    }
    flush_counters(a, cnt);
#undef PF
#undef PU
#undef PLOAD3
#undef PSTORE3
}

}  // namespace vr
