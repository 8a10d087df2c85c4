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

// 8/16-bit non-interlaced PNG -> floats u8 / 255 (what sampling a GL_R8 / RG8 / RGB8 / RGBA8 texture returns; cppgl
// texture.cpp:27-60 uploads LDR files as unsigned-byte formats, no gamma), expanded to RGB the way `.rgb` of such a
// texture reads: gray -> (g, 0, 0), gray+alpha -> (g, a, 0), RGBA -> RGB. 16-bit samples keep their high byte (stb_image).
ImageF load_png_rgb(const std::string& path, bool flip = true);
// environment maps: .hdr (Radiance) or .png by extension; anything else is an error
ImageF load_environment_image(const std::string& path);

// 8-bit image writer (.png, .ppm); flip = write the last row first (cppgl image_store_ldr default)
void store_ldr(const std::string& path, const uint8_t* pixels, int w, int h, int channels, bool flip = true);

}  // namespace volren
