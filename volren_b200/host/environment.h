// environment.h -- HDR environment map of the VolRen host (reference src/environment.{h,cpp}). The texture pair
// (envmap RGB32F + 512^2 importance map with its mip pyramid) lives on the device: the renderer uploads `pixels` with
// vrb_env_upload, which runs the env_setup / pyramid kernels (include/vrb200.h), whenever the bound Environment changes.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "vmath.h"

class Environment {
public:
    explicit Environment(const std::string& path);                 // environment.cpp:9 (HDR file, flipped on load)
    Environment(int w, int h, const float* rgb);                   // environment.cpp:11 (from an existing RGB image)
    virtual ~Environment() {}

    explicit operator bool() const { return !pixels.empty(); }
    uint32_t num_mip_levels() const { return 10; }                 // 1 + floor(log2(512))
    uint32_t dimension() const { return 512; }

    // data
    vmath::mat3 transform;
    float strength;
    int width, height;
    std::vector<float> pixels;     // RGB, bottom-up
    uint64_t id;                   // unique per object: lets the renderer see that a different map was bound
};
