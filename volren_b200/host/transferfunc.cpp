#include "transferfunc.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <stdexcept>

using vmath::vec3;
using vmath::vec4;

namespace colormap {

// Turbo: A. Mikhailov's polynomial fit (Google, Apache-2.0); Viridis: M. Zucker's degree-6 fit (public domain).
// The reference samples tinycolormap's 256-entry tables; these fits stay within ~1e-2 of them in the interior.
vec3 GetColor(float x, ColormapType type) {
    x = std::fmin(std::fmax(x, 0.f), 1.f);
    auto sat = [](float v) { return std::fmin(std::fmax(v, 0.f), 1.f); };
    switch (type) {
        case ColormapType::Turbo: {
            const float x2 = x * x, x3 = x2 * x, x4 = x2 * x2, x5 = x4 * x;
            return vec3(sat(0.13572138f + 4.61539260f * x - 42.66032258f * x2 + 132.13108234f * x3 - 152.94239396f * x4 + 59.28637943f * x5),
                        sat(0.09140261f + 2.19418839f * x + 4.84296658f * x2 - 14.18503333f * x3 + 4.27729857f * x4 + 2.82956604f * x5),
                        sat(0.10667330f + 12.64194608f * x - 60.58204836f * x2 + 110.36276771f * x3 - 89.90310912f * x4 + 27.34824973f * x5));
        }
        case ColormapType::Viridis: {
            const vec3 c0(0.2777273272234177f, 0.005407344544966578f, 0.3340998053353061f), c1(0.1050930431085774f, 1.404613529898575f, 1.384590162594685f),
                c2(-0.3308618287255563f, 0.214847559468213f, 0.09509516302823659f), c3(-4.634230498983486f, -5.799100973351585f, -19.33244095627987f),
                c4(6.228269936347081f, 14.17993336680509f, 56.69055260068105f), c5(4.776384997670288f, -13.74514537774601f, -65.35303263337234f),
                c6(-5.435455855934631f, 4.645852612178535f, 26.3124352495832f);
            const vec3 c = c0 + x * (c1 + x * (c2 + x * (c3 + x * (c4 + x * (c5 + x * c6)))));
            return vec3(sat(c.x), sat(c.y), sat(c.z));
        }
        case ColormapType::Heat: return vec3(sat(3.f * x), sat(3.f * x - 1.f), sat(3.f * x - 2.f));
        default: return vec3(x);
    }
}

}  // namespace colormap

static std::atomic<uint64_t> next_tf_id{ 1 };

TransferFunction::TransferFunction() : window_left(0), window_width(1), id(next_tf_id++), version(0) { randomize(); }
TransferFunction::TransferFunction(const std::string& path) : TransferFunction() { load_from_file(path); }
TransferFunction::TransferFunction(colormap::ColormapType type) : TransferFunction() { colormap(type); }
TransferFunction::TransferFunction(const std::vector<vec4>& l) : TransferFunction() {
    lut = l;
    upload_gpu();
}

// transferfunc.cpp:33-43: sequential fp32 prefix sum of alpha, normalised by the total (uniform ramp if it is <= 0)
std::vector<vec4> TransferFunction::compute_lut_cdf(const std::vector<vec4>& lut) {
    auto cdf = lut;
    if (cdf.empty()) return cdf;
    for (size_t i = 1; i < cdf.size(); ++i) cdf[i].w += cdf[i - 1].w;
    const float integral = cdf[cdf.size() - 1].w;
    for (size_t i = 0; i < cdf.size(); ++i) cdf[i].w = integral <= 0.f ? (i + 1) / float(cdf.size()) : cdf[i].w / integral;
    return cdf;
}

void TransferFunction::upload_gpu() {
    bool needs_cdf = false;   // monotone alpha is a hard requirement of the majorant mapping (common.glsl:472)
    for (size_t i = 1; i < lut.size(); ++i)
        if (lut[i - 1].w > lut[i].w) { needs_cdf = true; break; }
    lut_gpu = needs_cdf ? compute_lut_cdf(lut) : lut;
    ++version;
}

static inline float randf() { return rand() / (RAND_MAX + 1.f); }

void TransferFunction::randomize(size_t n_bins) {
    lut.clear();
    for (size_t i = 0; i < n_bins; ++i) {
        if (i == 0) { lut.push_back(vec4(0.f)); continue; }
        const float r = randf(), g = randf(), b = randf(), a = randf();   // argument evaluation order pinned left to right
        lut.push_back(vec4(r, g, b, a));
    }
    upload_gpu();
}

void TransferFunction::colormap(colormap::ColormapType type, size_t n_bins) {
    lut.clear();
    for (size_t i = 0; i < n_bins; ++i) {
        const float f = float(i) / n_bins;
        const vec3 c = colormap::GetColor(f, type);
        lut.push_back(vec4(c.x, c.y, c.z, f));
    }
    upload_gpu();
}

void TransferFunction::load_from_file(const std::string& path) {
    std::ifstream lut_file(path);
    if (!lut_file.is_open()) throw std::runtime_error("Unable to read file: " + path);
    lut.clear();
    std::cout << "Loading LUT: " << std::filesystem::path(path) << std::endl;
    char tmp[256];
    float r = 0, g = 0, b = 0, a = 0;   // a malformed line repeats the previous values, as sscanf leaves them untouched
    while (lut_file.getline(tmp, 256)) {
        sscanf(tmp, "%f, %f, %f, %f", &r, &g, &b, &a);
        lut.emplace_back(r, g, b, a);
    }
    upload_gpu();
}

void TransferFunction::write_to_file(const std::string& filename) {
    std::filesystem::path filepath = filename;
    filepath.replace_extension(".txt");
    std::ofstream file(filepath);
    if (!file.is_open()) return;
    char tmp[256];
    for (const auto& rgba : lut) {
        snprintf(tmp, 256, "%f, %f, %f, %f", rgba.x, rgba.y, rgba.z, rgba.w);
        file << tmp << std::endl;
    }
}
