// Bandwidth probes behind vrb_probe_bandwidth: the ceilings bench.py quotes next to the tracking kernel, measured in the
// same process on the same device. mode 0: every CTA streams the working set with 16-byte loads (grid-stride, CTA-
// dependent start so that CTAs do not march over the same lines); mode 1: random 32-byte sector gathers, one 4-byte load
// per sector with four independent loads in flight per thread -- the access pattern of a tracer lane. A working set below
// the 126 MB L2 measures L2, 1 GiB measures HBM.
#pragma once

#include "vr_common.cuh"

namespace vr {

__global__ void __launch_bounds__(512) k_probe_stream(const uint4* __restrict__ buf, size_t n16, int passes, uint32_t* sink) {
    uint32_t acc = 0;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (int p = 0; p < passes; ++p) {
        size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x + size_t(p) * 7919u * blockDim.x) % n16;
        for (size_t k = 0; k < n16; k += stride) {
            const uint4 v = __ldg(buf + i);
            acc += v.x ^ v.y ^ v.z ^ v.w;
            i += stride;
            if (i >= n16) i -= n16;
        }
    }
    if (acc == 0x12345678u) *sink = acc;      // never true for the all-ones fill: keeps the loads alive
}

__global__ void __launch_bounds__(256) k_probe_gather(const uint32_t* __restrict__ buf, uint32_t n_sectors, int iters, uint32_t* sink) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t idx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { s = s * LCG_A + LCG_C; idx[j] = uint32_t((uint64_t(s) * n_sectors) >> 32); }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += __ldg(buf + size_t(idx[j]) * 8u);
    }
    if (acc == 0x12345678u) *sink = acc;
}

}  // namespace vr
