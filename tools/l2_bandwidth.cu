// L2 / HBM read-bandwidth micro-benchmark: the denominators DESIGN.md quotes for L2-resident grids (C1/C2/C5) next to
// the driver-measured HBM copy bandwidth of MEASURED_PEAKS.json.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/l2_bandwidth tools/l2_bandwidth.cu
//   tools/_build/l2_bandwidth            # prints one JSON line per working-set size
// Each CTA streams the whole working set with 16-byte loads (grid-stride, every pass starts at a CTA-dependent offset so
// CTAs do not march in lock-step over the same lines); the sum is kept live through a never-true store.
// A second mode gathers random 32-byte sectors (one 4-byte load per sector, the tracer's access pattern) and reports
// sectors/s * 32 B: the ceiling of a dependent-gather kernel whose data sits in L2.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(512) k_stream(const uint4* __restrict__ buf, size_t n16, int passes, uint32_t* sink) {
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
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void __launch_bounds__(256) k_gather(const uint32_t* __restrict__ buf, uint32_t n_sectors, int iters, uint32_t* sink) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0;
    // 4 independent loads in flight per thread per iteration (a tracer lane has 1-8)
    for (int it = 0; it < iters; ++it) {
        uint32_t idx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { s = s * 1664525u + 1013904223u; idx[j] = uint32_t((uint64_t(s) * n_sectors) >> 32); }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += __ldg(buf + size_t(idx[j]) * 8u);
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint32_t* sink;
    CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t sizes_mb[] = { 2, 8, 16, 32, 64, 96, 256, 1024 };
    for (size_t mb : sizes_mb) {
        const size_t bytes = mb << 20, n16 = bytes / 16;
        uint4* buf;
        CK(cudaMalloc(&buf, bytes));
        CK(cudaMemset(buf, 1, bytes));
        const size_t target = size_t(8) << 30;                      // ~8 GiB of traffic per timed launch
        const int grid = sms * 4, block = 512;
        int passes = int(target / bytes);
        if (passes < 1) passes = 1;
        // every CTA walks the whole set `passes_cta` times; total traffic = grid-stride covers the set once per pass
        k_stream<<<grid, block>>>(buf, n16, 1, sink);               // warm
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0));
            k_stream<<<grid, block>>>(buf, n16, passes, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const double gbs = double(bytes) * passes / (best * 1e-3) / 1e9;
        // random sector gather
        const uint32_t n_sectors = uint32_t(bytes / 32);
        const int iters = 256;
        const int ggrid = sms * 16, gblock = 256;
        k_gather<<<ggrid, gblock>>>(reinterpret_cast<const uint32_t*>(buf), n_sectors, 8, sink);
        CK(cudaDeviceSynchronize());
        float gbest = 1e30f;
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0));
            k_gather<<<ggrid, gblock>>>(reinterpret_cast<const uint32_t*>(buf), n_sectors, iters, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < gbest) gbest = ms;
        }
        const double sectors = double(ggrid) * gblock * iters * 4;
        printf("{\"working_set_mib\": %zu, \"stream_read_gbs\": %.1f, \"gather_gsectors_s\": %.2f, \"gather_gbs_at_32B\": %.1f}\n",
               mb, gbs, sectors / (gbest * 1e-3) / 1e9, sectors * 32 / (gbest * 1e-3) / 1e9);
        CK(cudaFree(buf));
    }
    return 0;
}
