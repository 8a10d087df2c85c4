// Brick build, pass A for u8 DenseGrid sources with dim.x % 8 == 0: x- and y-part of the min/max over the 12^3 dilated window
// of every brick (reference: submodules/voldata/src/grid_brick.cpp:80-95), one streaming pass over the voxels.
//
// Round-2 history (1024^3 on B200): round 1 ran x and y as two kernels (645 + 168 us, 1.75x the voxels through DRAM); the first
// fused version kept four bytes per word and reduced them with __vminu4 / __vmaxu4 -- which sm_100a EMULATES with ~6 LOP3/SHF/PRMT
// each (there is no packed-byte min/max instruction), 170 instructions per 16 voxels, issue bound at 529 us = 2.1 TB/s
// (history in profiles/r02_brick_build_launches_v3.txt). sm_90+ does have a native packed 16-bit min/max (VIMNMX.U16x2 = __vminu2 /
// __vmaxu2), so this version widens the bytes of a word to two u16x2 values (even bytes: one LOP3, odd bytes: one PRMT) and
// keeps every running minimum / maximum as u16x2:
//   * a THREAD owns 16 consecutive voxels of a row (two brick columns A, B) in one z-slice and walks a band of RANGE_BAND_BY
//     brick rows; per row it loads its 16 bytes (one 16-byte load) plus the 4-byte words left and right of them (the two halo
//     voxels on either side; L1 hits, the neighbouring threads load the same lines) -- no shuffles, no shared memory, so the
//     kernel also compiles as plain C++ (tests/cpu_harness/range_xy_host.cpp checks it against a brute-force loop on the CPU);
//   * rows are reduced in groups of four, aligned so that group 2b is rows 8b-2 ... 8b+1, 2b+1 is 8b+2 ... 8b+5 and 2b+2 is
//     8b+6 ... 8b+9: the y-window of brick row b is groups 2b, 2b+1, 2b+2, and group 2b+2 is also the first group of row b+1,
//     so every voxel row enters exactly ONE accumulation (20 VIMNMX per row) and the 12-voxel x-window is folded once per
//     group (24 instructions per four rows), not once per row;
//   * out-of-grid voxels never enter a minimum / maximum (DenseGrid::lookup returns a literal 0.f there; whether a window leaves
//     the grid is pure geometry and is added by k_range_z): missing rows are skipped, a missing halo is replaced by a duplicate
//     of an in-window voxel; {255, 0} (min > max) marks "no in-grid voxel". Threads whose 16 voxels are not all inside the grid
//     (the padding columns of n_bricks, rounded up to a multiple of 8) take a scalar per-voxel path.
// Output: m2[z][by][bx] = min | max << 8 over x in [8bx-2, 8bx+9], y in [8by-2, 8by+9] of slice z. k_range_z finishes the window.
// Bytes: voxels x (1 + 4 / (8 RANGE_BAND_BY)) read, 2 B per (z-slice, brick column) written.
#pragma once

#include <stdint.h>

namespace vr {

constexpr int RANGE_BAND_BY = 16;
constexpr int RANGE_XY_THREADS = 128;

#ifdef VR_RANGE_HOST_HARNESS
#define VR_RXY_DEV static inline
#else
#define VR_RXY_DEV __device__ __forceinline__
#endif

VR_RXY_DEV uint32_t rxy_even(uint32_t w) { return w & 0x00ff00ffu; }                 // bytes 0, 2 as u16x2
VR_RXY_DEV uint32_t rxy_odd(uint32_t w) { return __byte_perm(w, 0u, 0x4341u); }      // bytes 1, 3 as u16x2
VR_RXY_DEV uint32_t rxy_fold_min(uint32_t v) { return min(v & 0xffffu, v >> 16); }
VR_RXY_DEV uint32_t rxy_fold_max(uint32_t v) { return max(v & 0xffffu, v >> 16); }

// scalar path of one (brick row, brick column): min | max << 8 over the in-grid voxels of the 12 x 12 window
VR_RXY_DEV uint32_t rxy_window_scalar(const uint8_t* __restrict__ slice, uint32_t dim_x, uint32_t dim_y, int bx, int by) {
    uint32_t mn = 255u, mx = 0u;
    for (int y = by * 8 - 2; y <= by * 8 + 9; ++y) {
        if (y < 0 || y >= int(dim_y)) continue;
        for (int x = bx * 8 - 2; x <= bx * 8 + 9; ++x) {
            if (x < 0 || x >= int(dim_x)) continue;
            const uint32_t u = slice[size_t(y) * dim_x + uint32_t(x)];
            mn = min(mn, u);
            mx = max(mx, u);
        }
    }
    return mn | (mx << 8);
}

#ifndef VR_RANGE_HOST_HARNESS
__global__ void __launch_bounds__(RANGE_XY_THREADS)
#else
static void
#endif
k_range_xy(const uint8_t* __restrict__ vox, uint3 dim, uint3 nb, uint16_t* __restrict__ m2, int vec16) {
    const uint32_t chunks_x = nb.x >> 1;                              // n_bricks is a multiple of 8: whole pairs of brick columns
    const size_t item = blockIdx.x * size_t(RANGE_XY_THREADS) + threadIdx.x;
    const uint32_t z = uint32_t(item / chunks_x), chunk = uint32_t(item % chunks_x);
    if (z >= dim.z) return;
    const uint32_t bxA = chunk * 2u, x0 = bxA * 8u;
    const uint32_t by0 = blockIdx.y * uint32_t(RANGE_BAND_BY), by1 = min(by0 + uint32_t(RANGE_BAND_BY), nb.y);
    const uint8_t* slice = vox + size_t(z) * dim.y * dim.x;
    uint32_t* out = reinterpret_cast<uint32_t*>(m2 + size_t(z) * nb.y * nb.x + bxA);       // brick row b: out[b * (nb.x / 2)]
    const uint32_t out_stride = nb.x >> 1;

    if (x0 + 16u > dim.x) {
        // padding columns: at most one of the two columns holds grid voxels (or only sees them through its halo)
        for (uint32_t b = by0; b < by1; ++b)
            out[size_t(b) * out_stride] = rxy_window_scalar(slice, dim.x, dim.y, int(bxA), int(b)) | (rxy_window_scalar(slice, dim.x, dim.y, int(bxA) + 1, int(b)) << 16);
        return;
    }
    const bool valid_l = x0 > 0u, valid_r = x0 + 16u < dim.x;
    constexpr uint32_t MN0 = 0x00ff00ffu;                              // neutral elements of the u16x2 min / max over byte values
    uint32_t hm_mn_a = MN0, hm_mx_a = 0u, hm_mn_b = MN0, hm_mx_b = 0u;   // groups 2b (and 2b + 1) of the brick row being assembled
    for (uint32_t g = 2u * by0; g <= 2u * by1; ++g) {
        // values per row: [0..7] = even / odd bytes of the four own words, [8] = left halo pair, [9] = right halo pair
        uint32_t mn[10], mx[10];
        const int yb = int(4u * g) - 2;
        if (yb >= 0 && yb + 3 < int(dim.y)) {
            // interior group (all but the first / last of a slice): branch-free, the twelve loads of its four rows go out together
            uint32_t w[4][4], hl[4], hr[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint8_t* row = slice + size_t(yb + r) * dim.x + x0;
                if (vec16) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(row));
                    w[r][0] = v.x; w[r][1] = v.y; w[r][2] = v.z; w[r][3] = v.w;
                } else {
                    const uint2 v0 = __ldg(reinterpret_cast<const uint2*>(row)), v1 = __ldg(reinterpret_cast<const uint2*>(row + 8));
                    w[r][0] = v0.x; w[r][1] = v0.y; w[r][2] = v1.x; w[r][3] = v1.y;
                }
                hl[r] = valid_l ? __ldg(reinterpret_cast<const uint32_t*>(row - 4)) : 0u;
                hr[r] = valid_r ? __ldg(reinterpret_cast<const uint32_t*>(row + 16)) : 0u;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                uint32_t v[10];
#pragma unroll
                for (int k = 0; k < 4; ++k) { v[2 * k] = rxy_even(w[r][k]); v[2 * k + 1] = rxy_odd(w[r][k]); }
                v[8] = valid_l ? __byte_perm(hl[r], 0u, 0x4342u) : v[0];
                v[9] = valid_r ? __byte_perm(hr[r], 0u, 0x4140u) : v[7];
#pragma unroll
                for (int i = 0; i < 10; ++i) {
                    if (r == 0) { mn[i] = v[i]; mx[i] = v[i]; }
                    else { mn[i] = __vminu2(mn[i], v[i]); mx[i] = __vmaxu2(mx[i], v[i]); }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 10; ++i) { mn[i] = MN0; mx[i] = 0u; }
            for (int r = 0; r < 4; ++r) {
                const int y = yb + r;
                if (y < 0 || y >= int(dim.y)) continue;                // rows outside the grid never enter a minimum / maximum
                const uint8_t* row = slice + size_t(y) * dim.x + x0;
                uint32_t w[4];
                if (vec16) {
                    const uint4 q = __ldg(reinterpret_cast<const uint4*>(row));
                    w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
                } else {
                    const uint2 v0 = __ldg(reinterpret_cast<const uint2*>(row)), v1 = __ldg(reinterpret_cast<const uint2*>(row + 8));
                    w[0] = v0.x; w[1] = v0.y; w[2] = v1.x; w[3] = v1.y;
                }
                uint32_t v[10];
#pragma unroll
                for (int k = 0; k < 4; ++k) { v[2 * k] = rxy_even(w[k]); v[2 * k + 1] = rxy_odd(w[k]); }
                v[8] = valid_l ? __byte_perm(__ldg(reinterpret_cast<const uint32_t*>(row - 4)), 0u, 0x4342u) : v[0];    // voxels x0 - 2, x0 - 1 (bytes 2, 3 of the word on the left); else a duplicate of column A
                v[9] = valid_r ? __byte_perm(__ldg(reinterpret_cast<const uint32_t*>(row + 16)), 0u, 0x4140u) : v[7];   // voxels x0 + 16, x0 + 17; else a duplicate of column B
#pragma unroll
                for (int i = 0; i < 10; ++i) { mn[i] = __vminu2(mn[i], v[i]); mx[i] = __vmaxu2(mx[i], v[i]); }
            }
        }
        // x-windows of the group. Column A: x0 - 2 ... x0 + 9 = left pair, words 0 and 1, bytes 0, 1 of word 2 (the low halves of
        // v[4], v[5]); column B: x0 + 6 ... x0 + 17 = bytes 2, 3 of word 1 (the high halves of v[2], v[3]), words 2 and 3, right pair
        const uint32_t a_mn = __vminu2(__vminu2(__vminu2(mn[0], mn[1]), __vminu2(mn[2], mn[3])), __vminu2(mn[8], __byte_perm(mn[4], mn[5], 0x5410u)));
        const uint32_t a_mx = __vmaxu2(__vmaxu2(__vmaxu2(mx[0], mx[1]), __vmaxu2(mx[2], mx[3])), __vmaxu2(mx[8], __byte_perm(mx[4], mx[5], 0x5410u)));
        const uint32_t b_mn = __vminu2(__vminu2(__vminu2(mn[4], mn[5]), __vminu2(mn[6], mn[7])), __vminu2(mn[9], __byte_perm(mn[2], mn[3], 0x7632u)));
        const uint32_t b_mx = __vmaxu2(__vmaxu2(__vmaxu2(mx[4], mx[5]), __vmaxu2(mx[6], mx[7])), __vmaxu2(mx[9], __byte_perm(mx[2], mx[3], 0x7632u)));
        if ((g & 1u) == 0u) {
            // an even group closes brick row b = g / 2 - 1 (its rows 8b + 6 ... 8b + 9) and opens row g / 2 (its rows 8b' - 2 ... 8b' + 1)
            if (g > 2u * by0) {
                const uint32_t b = (g >> 1) - 1u;
                const uint32_t lo_a = rxy_fold_min(__vminu2(hm_mn_a, a_mn)), hi_a = rxy_fold_max(__vmaxu2(hm_mx_a, a_mx));
                const uint32_t lo_b = rxy_fold_min(__vminu2(hm_mn_b, b_mn)), hi_b = rxy_fold_max(__vmaxu2(hm_mx_b, b_mx));
                out[size_t(b) * out_stride] = (lo_a | (hi_a << 8)) | ((lo_b | (hi_b << 8)) << 16);
            }
            hm_mn_a = a_mn; hm_mx_a = a_mx; hm_mn_b = b_mn; hm_mx_b = b_mx;
        } else {
            hm_mn_a = __vminu2(hm_mn_a, a_mn); hm_mx_a = __vmaxu2(hm_mx_a, a_mx);
            hm_mn_b = __vminu2(hm_mn_b, b_mn); hm_mx_b = __vmaxu2(hm_mx_b, b_mx);
        }
    }
}

}  // namespace vr
