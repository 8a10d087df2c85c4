// image_io.h -- HDR environment-map loading and LDR image writing for the VolRen host (see image_io.cpp).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace volren {

struct ImageF {
    int w = 0, h = 0, channels = 0;
    std::vector<float> data;   // row-major, `channels` interleaved
};

// Radiance .hdr -> RGB float; flip = bottom-up rows (cppgl image_load: stbi_set_flip_vertically_on_load(1))
ImageF load_hdr(const std::string& path, bool flip = true);

// 8-bit image writer (.png, .ppm); flip = write the last row first (cppgl image_store_ldr default)
void store_ldr(const std::string& path, const uint8_t* pixels, int w, int h, int channels, bool flip = true);

}  // namespace volren
