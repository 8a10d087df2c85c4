#include "transferfunc.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <stdexcept>

using vmath::vec3;
using vmath::vec4;

namespace colormap {

#include "colormap_tables.inc"

// tinycolormap::GetColor (tinycolormap.hpp:200-237) in double precision, as TransferFunction::colormap calls it
// (reference src/transferfunc.cpp:69-77). Table maps: internal::CalcLerp (:160-170) -- a = clamp01(x) * (N - 1),
// (1 - t) * data[floor(a)] + t * data[ceil(a)]; Hot and Gray are closed forms (:806-834).
static void lerp_table(double x, const unsigned int* data, size_t n, double out[3]) {
    const double xc = (x < 0.0) ? 0.0 : (x > 1.0) ? 1.0 : x;
    const double a = xc * double(n - 1);
    const double i = std::floor(a);
    const double t = a - i;
    const unsigned int* c0 = data + 3 * static_cast<size_t>(i);
    const unsigned int* c1 = data + 3 * static_cast<size_t>(std::ceil(a));
    for (int k = 0; k < 3; ++k) out[k] = (1.0 - t) * (double(c0[k]) / 1e6) + t * (double(c1[k]) / 1e6);
}

void GetColor(double x, ColormapType type, double out[3]) {
#define VR_TABLE(name) lerp_table(x, colormap_##name, sizeof(colormap_##name) / sizeof(unsigned int) / 3, out); return
    switch (type) {
        case ColormapType::Parula: VR_TABLE(parula);
        case ColormapType::Heat: VR_TABLE(heat);
        case ColormapType::Jet: VR_TABLE(jet);
        case ColormapType::Turbo: VR_TABLE(turbo);
        case ColormapType::Magma: VR_TABLE(magma);
        case ColormapType::Inferno: VR_TABLE(inferno);
        case ColormapType::Plasma: VR_TABLE(plasma);
        case ColormapType::Viridis: VR_TABLE(viridis);
        case ColormapType::Cividis: VR_TABLE(cividis);
        case ColormapType::Github: VR_TABLE(github);
        case ColormapType::Cubehelix: VR_TABLE(cubehelix);
        case ColormapType::HSV: VR_TABLE(hsv);
        case ColormapType::Hot: {
            const double xc = (x < 0.0) ? 0.0 : (x > 1.0) ? 1.0 : x;
            if (xc < 0.4) { out[0] = xc / 0.4 * 1.0; out[1] = xc / 0.4 * 0.0; out[2] = xc / 0.4 * 0.0; }
            else if (xc < 0.8) { const double t = (xc - 0.4) / (0.8 - 0.4); out[0] = 1.0 + t * 0.0; out[1] = 0.0 + t * 1.0; out[2] = 0.0 + t * 0.0; }
            else { const double t = (xc - 0.8) / (1.0 - 0.8); out[0] = 1.0 + 0.0 + t * 0.0; out[1] = 0.0 + 1.0 + t * 0.0; out[2] = 0.0 + 0.0 + t * 1.0; }
            return;
        }
        case ColormapType::Gray: {
            const double xc = (x < 0.0) ? 0.0 : (x > 1.0) ? 1.0 : x;
            out[0] = out[1] = out[2] = 1.0 - xc;
            return;
        }
    }
#undef VR_TABLE
    throw std::invalid_argument("unknown ColormapType");
}

vec3 GetColor(float x, ColormapType type) {
    double c[3];
    GetColor(double(x), type, c);
    return vec3(float(c[0]), float(c[1]), float(c[2]));
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
