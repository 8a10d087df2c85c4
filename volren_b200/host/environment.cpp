#include "environment.h"

#include <atomic>
#include <stdexcept>

#include "image_io.h"

static std::atomic<uint64_t> next_env_id{ 1 };

Environment::Environment(const std::string& path) : transform(1.f), strength(1.f), width(0), height(0), id(next_env_id++) {
    volren::ImageF img = volren::load_environment_image(path);     // .hdr, or an LDR .png sampled as unorm (cppgl texture.cpp:27-60)
    width = img.w;
    height = img.h;
    pixels = std::move(img.data);
}

Environment::Environment(int w, int h, const float* rgb) : transform(1.f), strength(1.f), width(w), height(h), pixels(rgb, rgb + size_t(w) * h * 3), id(next_env_id++) {
    if (w <= 0 || h <= 0) throw std::runtime_error("Environment: bad image size");
}
