// CPU check (TEST INFRASTRUCTURE) of encode_code_fast (volren_b200/csrc/vr_brick.cuh): the same IEEE operations in C, against the plain
// expression round(255 * clamp((v - lo) / span)) of the reference (grid_brick.cpp:45-48), random fp16 ranges incl. collapsed ones.
// usage: encode_fast_host [iterations]
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static float half_to_float(uint16_t h) { uint32_t sgn = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023; float v; if (e == 0) v = ldexpf((float)m, -24); else if (e == 31) v = m ? NAN : INFINITY; else v = ldexpf((float)(m | 1024), (int)e - 25); return sgn ? -v : v; }
static uint32_t enc_ref(float v, float lo, float span) {
    float vn = (v - lo) / span;
    vn = vn < 0.f ? 0.f : vn; vn = 1.f < vn ? 1.f : vn;
    float q = roundf(255.f * vn);
    return isnan(q) ? 0u : (uint32_t)(int)q;
}
static uint32_t enc_fast(float v, float lo, float span, float r, int span_ok) {
    float a = v - lo, vn;
    float aa = fabsf(a);
    if (span_ok && (a == 0.f || (aa > 0x1p-60f && aa < 0x1p60f))) {
        float q = a * r; float e = fmaf(-span, q, a); q = fmaf(e, r, q); e = fmaf(-span, q, a); vn = fmaf(e, r, q);
    } else vn = a / span;
    vn = vn < 0.f ? 0.f : vn; vn = 1.f < vn ? 1.f : vn;
    float x = 255.f * vn, t = truncf(x);
    float q = (x - t >= 0.5f) ? t + 1.f : t;
    return isnan(q) ? 0u : (uint32_t)(int)q;
}
int main(int argc, char** argv) {
    long bad = 0, n = 0;
    const long iters = argc > 1 ? atol(argv[1]) : 300000000L;
    for (long it = 0; it < iters; ++it) {
        uint16_t hl = rnd() & 0xffff, hh = rnd() & 0xffff;
        float lo = half_to_float(hl), hi = half_to_float(hh);
        float vmin = 0.f, vmax = 1.f;
        int mode = it % 5;
        if (mode == 1) { vmin = lo * 0.5f; vmax = hi * 2.f; }
        if (mode == 2) { vmin = -3.f; vmax = 7.5f; }
        if (mode == 3) { vmin = 1e-30f; vmax = 3e-30f; }
        if (mode == 4) { hl = hh; lo = hi; }       /* collapsed range: span = 0 */
        uint32_t c = rnd() & 255;
        float v = (rnd() & 15) == 0 ? 0.f : vmin + ((float)c / 255.f) * (vmax - vmin);
        float span = hi - lo;
        float r = 1.0f / span;
        int span_ok = fabsf(span) > 0x1p-60f && fabsf(span) < 0x1p60f;
        uint32_t a = enc_ref(v, lo, span), b = enc_fast(v, lo, span, r, span_ok);
        n++;
        if (a != b && bad++ < 10) printf("bad v=%a lo=%a span=%a ref=%u fast=%u\n", v, lo, span, a, b);
    }
    printf("checked %ld bad %ld\n", n, bad);
    return bad != 0;
}
