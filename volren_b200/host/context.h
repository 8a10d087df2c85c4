// context.h -- what is left of cppgl's `Context` singleton once there is no OpenGL: the output resolution
// (Context::resolution() == framebuffer size == `-w/-h`, cppgl/src/context.cpp:252-256), the per-process device
// contexts of the B200 back end (one vrb_ctx per GPU; include/vrb200.h) and the per-sample sync that
// `Context::swap_buffers()` provided in the offline loop (main.cpp:536, bindings.cpp:130).
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vrb200.h"
#include "vmath.h"

namespace volren {

struct ContextParameters {      // cppgl ContextParameters (context.h:12-14): only what survives without a window
    uint32_t width = 1280;
    uint32_t height = 720;
    int n_gpus = 1;             // new: --gpus N
    int first_device = 0;       // new: VOLREN_DEVICE / --device
    std::string partition = "spp";   // new: --partition spp|tile (multi-GPU split)
};

class Context {
public:
    static void init(const ContextParameters& params = ContextParameters());
    static bool initialized();
    static Context& instance();              // lazily initialises with defaults
    static vmath::ivec2 resolution();
    static void resize(uint32_t w, uint32_t h);
    static void swap_buffers();              // stream sync on every device
    static vrb_ctx* device(int i = 0);       // primary context = device(0)
    static int n_devices();
    static const std::string& partition();
    static void shutdown();
    static uint64_t generation();            // bumped by every init(): device-side state of older generations is gone

    ContextParameters params;
    std::vector<vrb_ctx*> ctxs;
};

// C-ABI status -> C++ exception (the reference's error convention above the ABI is std::runtime_error)
inline void check(vrb_ctx* ctx, int status, const char* what) {
    if (status != VRB_OK)
        throw std::runtime_error(std::string(what) + ": " + vrb_status_string(status) + (ctx ? std::string(" (") + vrb_last_error(ctx) + ")" : std::string()));
}

}  // namespace volren
