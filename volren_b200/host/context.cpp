#include "context.h"

#include <cstdlib>
#include <iostream>

namespace volren {

static Context* g_ctx = nullptr;
static uint64_t g_generation = 0;

uint64_t Context::generation() { return g_generation; }

void Context::init(const ContextParameters& params) {
    if (g_ctx) shutdown();
    g_ctx = new Context();
    g_ctx->params = params;
    ++g_generation;
    if (const char* dev = std::getenv("VOLREN_DEVICE")) g_ctx->params.first_device = std::atoi(dev);
    const int n = params.n_gpus > 0 ? params.n_gpus : 1;
    for (int i = 0; i < n; ++i) {
        vrb_ctx* c = nullptr;
        const int st = vrb_create(g_ctx->params.first_device + i, &c);
        if (st != VRB_OK) {
            const std::string msg = std::string("Failed to create context on CUDA device ") + std::to_string(g_ctx->params.first_device + i) + ": " + vrb_status_string(st) +
                                    " (libvrb200 has no CPU fallback)";
            shutdown();
            throw std::runtime_error(msg);
        }
        g_ctx->ctxs.push_back(c);
        check(c, vrb_resize(c, int(params.width), int(params.height)), "vrb_resize");
    }
}

bool Context::initialized() { return g_ctx != nullptr; }

Context& Context::instance() {
    if (!g_ctx) init(ContextParameters());
    return *g_ctx;
}

vmath::ivec2 Context::resolution() {
    const Context& c = instance();
    return vmath::ivec2(int(c.params.width), int(c.params.height));
}

void Context::resize(uint32_t w, uint32_t h) {
    Context& c = instance();
    c.params.width = w;
    c.params.height = h;
    for (vrb_ctx* ctx : c.ctxs) check(ctx, vrb_resize(ctx, int(w), int(h)), "vrb_resize");
}

void Context::swap_buffers() {
    for (vrb_ctx* ctx : instance().ctxs) check(ctx, vrb_sync(ctx), "vrb_sync");
}

vrb_ctx* Context::device(int i) { return instance().ctxs.at(size_t(i)); }
int Context::n_devices() { return int(instance().ctxs.size()); }
const std::string& Context::partition() { return instance().params.partition; }

void Context::shutdown() {
    if (!g_ctx) return;
    for (vrb_ctx* ctx : g_ctx->ctxs) vrb_destroy(ctx);
    delete g_ctx;
    g_ctx = nullptr;
}

}  // namespace volren
