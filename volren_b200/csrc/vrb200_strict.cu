// The IEEE cross-check kernels of libvrb200.so (vrb_set_kernel 1 and 2), in a translation unit of their own that is
// compiled with -fmad=false: no FMA contraction, IEEE division / sqrt, accurate log / sincos -- operation for operation
// the arithmetic of the CPU restatement the tests check against, which in turn equals the reference's GLSL compiled as
// C++ bit for bit (tests/test_glsl_ref.py). What is left between these kernels and the oracle is the last bit of the libm
// functions (CUDA's logf / sincosf / atan2f / acosf vs glibc's), so a same-seed replay follows the oracle path for path
// except where such a bit flips a comparison (tests/test_gpu_render.py::test_ieee_kernels_replay_the_oracle).
// Test infrastructure inside the library: the production path (vrb200.cu, FastMath kernels) never calls into this file
// unless vrb_set_kernel(ctx, 1 | 2) was requested.
#define VR_STRICT_TU 1
#define VR_TRACE_MIN_BLOCKS 4      // cross-check kernels: registers before occupancy (the IEEE sequences spill at 72)
#include "vr_common.cuh"
#include "vr_env.cuh"
#include "vr_trace.cuh"
#include "vr_trace2.cuh"

namespace vr {

cudaError_t launch_trace_pixels(const TraceArgs& a, bool tf, bool count, dim3 grid, cudaStream_t stream) {
    if (count) {
        if (tf) k_trace_pixels<true, true><<<grid, 256, 0, stream>>>(a);
        else k_trace_pixels<false, true><<<grid, 256, 0, stream>>>(a);
    } else {
        if (tf) k_trace_pixels<true, false><<<grid, 256, 0, stream>>>(a);
        else k_trace_pixels<false, false><<<grid, 256, 0, stream>>>(a);
    }
    return cudaGetLastError();
}

// the per-level majorant tables the persistent kernel reads, filled with THIS unit's arithmetic (kind 1 evaluates the same
// expression inline, majorant_at): one rounding difference in a majorant moves every later collision point of a ray
cudaError_t launch_majorant_table_strict(const TraceArgs& a, bool tf, int level, float* out, size_t n, int blocks, cudaStream_t stream) {
    if (tf) k_majorant_table<true><<<blocks, 256, 0, stream>>>(a, level, out, n);
    else k_majorant_table<false><<<blocks, 256, 0, stream>>>(a, level, out, n);
    return cudaGetLastError();
}

// the lane-resident persistent kernel with StrictMath (kind 2): must reproduce kind 1 path for path
const void* strict_persistent_kernel(bool tf, bool count) {
    if (count) return tf ? (const void*)k_trace_persistent<true, true, StrictMath> : (const void*)k_trace_persistent<false, true, StrictMath>;
    return tf ? (const void*)k_trace_persistent<true, false, StrictMath> : (const void*)k_trace_persistent<false, false, StrictMath>;
}

}  // namespace vr
