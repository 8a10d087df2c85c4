// CPU harness for volren_b200/csrc/vr_brick_range.cuh (TEST INFRASTRUCTURE): the kernel uses no shuffles and no shared
// memory, so with a dozen shims for the CUDA built-ins it compiles as plain C++ and every "thread" can be run in a loop.
// Compared against a brute-force min/max over the in-grid voxels of each (z, by, bx) window. Usage: range_xy_host [seed]
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct uint2 { uint32_t x, y; };
struct uint3 { uint32_t x, y, z; };
struct uint4 { uint32_t x, y, z, w; };
static uint3 blockIdx, threadIdx;
using std::max;
using std::min;
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t v = (uint64_t(y) << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= uint32_t((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
}
static inline uint32_t __vminu2(uint32_t a, uint32_t b) { return min(a & 0xffffu, b & 0xffffu) | (min(a >> 16, b >> 16) << 16); }
static inline uint32_t __vmaxu2(uint32_t a, uint32_t b) { return max(a & 0xffffu, b & 0xffffu) | (max(a >> 16, b >> 16) << 16); }
#define VR_RANGE_HOST_HARNESS 1
#define __restrict__
#include "../../volren_b200/csrc/vr_brick_range.cuh"

static uint32_t n_bricks_of(uint32_t d) { const uint32_t b = (d + 7) / 8; return ((b + 7) / 8) * 8; }

static int run_case(uint32_t dx, uint32_t dy, uint32_t dz, unsigned seed, int sparse) {
    srand(seed);
    std::vector<uint8_t> buf(size_t(dx) * dy * dz + 64);
    uint8_t* vox = buf.data();
    while ((reinterpret_cast<uintptr_t>(vox) & 15u) != 0) ++vox;
    for (size_t i = 0; i < size_t(dx) * dy * dz; ++i) vox[i] = sparse ? ((rand() % 97) == 0 ? uint8_t(rand()) : 0) : uint8_t(rand());
    const uint3 dim = { dx, dy, dz }, nb = { n_bricks_of(dx), n_bricks_of(dy), n_bricks_of(dz) };
    std::vector<uint16_t> m2(size_t(dz) * nb.y * nb.x, 0xabcd);
    const uint32_t chunks_x = nb.x / 2;
    const size_t items = size_t(dz) * chunks_x;
    const int vec16 = (dx % 16 == 0) ? 1 : 0;
    for (uint32_t gy = 0; gy < (nb.y + vr::RANGE_BAND_BY - 1) / vr::RANGE_BAND_BY; ++gy)
        for (uint32_t gx = 0; gx < (items + vr::RANGE_XY_THREADS - 1) / vr::RANGE_XY_THREADS; ++gx)
            for (uint32_t t = 0; t < uint32_t(vr::RANGE_XY_THREADS); ++t) {
                blockIdx = { gx, gy, 0 };
                threadIdx = { t, 0, 0 };
                vr::k_range_xy(vox, dim, nb, m2.data(), vec16);
            }
    int bad = 0;
    for (uint32_t z = 0; z < dz; ++z)
        for (uint32_t by = 0; by < nb.y; ++by)
            for (uint32_t bx = 0; bx < nb.x; ++bx) {
                uint32_t mn = 255, mx = 0;
                for (int y = int(by) * 8 - 2; y <= int(by) * 8 + 9; ++y)
                    for (int x = int(bx) * 8 - 2; x <= int(bx) * 8 + 9; ++x)
                        if (x >= 0 && y >= 0 && x < int(dx) && y < int(dy)) {
                            const uint32_t u = vox[(size_t(z) * dy + y) * dx + x];
                            mn = min(mn, u); mx = max(mx, u);
                        }
                const uint16_t want = uint16_t(mn | (mx << 8)), got = m2[(size_t(z) * nb.y + by) * nb.x + bx];
                if (want != got && bad++ < 5) printf("  mismatch dims %ux%ux%u z %u by %u bx %u: got %04x want %04x\n", dx, dy, dz, z, by, bx, got, want);
            }
    printf("%ux%ux%u sparse=%d: %s (%d mismatches)\n", dx, dy, dz, sparse, bad ? "FAIL" : "ok", bad);
    return bad;
}

int main(int argc, char** argv) {
    const unsigned seed = argc > 1 ? unsigned(atoi(argv[1])) : 1u;
    int bad = 0;
    const uint32_t shapes[][3] = { { 8, 8, 3 }, { 16, 8, 2 }, { 24, 17, 2 }, { 64, 64, 2 }, { 72, 33, 3 }, { 128, 130, 2 }, { 136, 300, 1 }, { 512, 70, 1 }, { 520, 12, 2 }, { 1024, 9, 1 }, { 40, 129, 2 } };
    for (const auto& s : shapes)
        for (int sparse = 0; sparse < 2; ++sparse) bad += run_case(s[0], s[1], s[2], seed + s[0] + 7 * s[1], sparse);
    return bad ? 1 : 0;
}
