// image_io.cpp -- image files either side of the renderer, without stb:
//   load_hdr   : Radiance RGBE (flat or new-style RLE), decoded like stbi_loadf (ldexp(1, e - 136) * byte, e == 0 -> 0)
//                and flipped to bottom-up rows like cppgl's image_load (image_load_store.cpp:14-41)
//   store_ldr  : 8-bit PNG (zlib deflate, filter 0) / binary PPM, flipped on write by default like image_store_ldr
//                (image_load_store.cpp:46-60)
#include "image_io.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <stdexcept>

namespace volren {

namespace {

std::string read_line(const std::vector<uint8_t>& d, size_t& pos) {
    size_t e = pos;
    while (e < d.size() && d[e] != '\n') ++e;
    if (e >= d.size()) throw std::runtime_error("truncated HDR header");
    std::string s(d.begin() + pos, d.begin() + e);
    pos = e + 1;
    return s;
}

}  // namespace

ImageF load_hdr(const std::string& path, bool flip) {
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("Failed to load image file: " + path);
    const std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t pos = 0;
    const std::string magic = read_line(d, pos);
    if (magic != "#?RADIANCE" && magic != "#?RGBE") throw std::runtime_error("Failed to load image file (not a Radiance HDR): " + path);
    bool fmt_ok = false;
    while (true) {
        const std::string s = read_line(d, pos);
        if (s.empty()) break;
        if (s.rfind("FORMAT=32-bit_rle_rgbe", 0) == 0) fmt_ok = true;
    }
    if (!fmt_ok) throw std::runtime_error("Unsupported HDR format: " + path);
    int h = 0, w = 0;
    const std::string res = read_line(d, pos);
    if (sscanf(res.c_str(), "-Y %d +X %d", &h, &w) != 2 || h <= 0 || w <= 0) throw std::runtime_error("Unsupported HDR orientation: " + path);
    std::vector<uint8_t> rgbe(size_t(w) * h * 4);
    auto need = [&](size_t n) { if (pos + n > d.size()) throw std::runtime_error("truncated HDR data: " + path); };
    bool flat = w < 8 || w >= 32768;
    if (!flat) {
        for (int y = 0; y < h && !flat; ++y) {
            need(4);
            if (!(d[pos] == 2 && d[pos + 1] == 2 && !(d[pos + 2] & 0x80))) {
                if (y != 0) throw std::runtime_error("mixed flat/RLE HDR scanlines: " + path);
                flat = true;
                break;
            }
            if (((int(d[pos + 2]) << 8) | int(d[pos + 3])) != w) throw std::runtime_error("corrupt HDR scanline width: " + path);
            pos += 4;
            for (int c = 0; c < 4; ++c) {
                int x = 0;
                while (x < w) {
                    need(1);
                    int n = d[pos++];
                    if (n > 128) {   // run
                        n -= 128;
                        need(1);
                        const uint8_t v = d[pos++];
                        if (x + n > w) throw std::runtime_error("corrupt HDR run: " + path);
                        for (int i = 0; i < n; ++i) rgbe[(size_t(y) * w + x + i) * 4 + c] = v;
                    } else {         // literal
                        need(size_t(n));
                        if (n == 0 || x + n > w) throw std::runtime_error("corrupt HDR literal: " + path);
                        for (int i = 0; i < n; ++i) rgbe[(size_t(y) * w + x + i) * 4 + c] = d[pos + i];
                        pos += size_t(n);
                    }
                    x += n;
                }
            }
        }
    }
    if (flat) {
        need(rgbe.size());
        memcpy(rgbe.data(), d.data() + pos, rgbe.size());
    }
    ImageF img;
    img.w = w; img.h = h; img.channels = 3;
    img.data.resize(size_t(w) * h * 3);
    for (int y = 0; y < h; ++y) {
        const int yo = flip ? h - 1 - y : y;
        for (int x = 0; x < w; ++x) {
            const uint8_t* p = &rgbe[(size_t(y) * w + x) * 4];
            float* o = &img.data[(size_t(yo) * w + x) * 3];
            if (p[3] != 0) {
                const float s = std::ldexp(1.0f, int(p[3]) - (128 + 8));
                o[0] = p[0] * s; o[1] = p[1] * s; o[2] = p[2] * s;
            } else
                o[0] = o[1] = o[2] = 0.f;
        }
    }
    return img;
}

namespace {

void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x)); }

void png_chunk(std::ofstream& out, const char tag[4], const std::vector<uint8_t>& payload) {
    std::vector<uint8_t> buf;
    put_be32(buf, uint32_t(payload.size()));
    out.write(reinterpret_cast<const char*>(buf.data()), 4);
    std::vector<uint8_t> body(tag, tag + 4);
    body.insert(body.end(), payload.begin(), payload.end());
    out.write(reinterpret_cast<const char*>(body.data()), std::streamsize(body.size()));
    buf.clear();
    put_be32(buf, uint32_t(crc32(0L, body.data(), uInt(body.size()))));
    out.write(reinterpret_cast<const char*>(buf.data()), 4);
}

}  // namespace

void store_ldr(const std::string& path, const uint8_t* pixels, int w, int h, int channels, bool flip) {
    const std::string ext = std::filesystem::path(path).extension().string();
    if (channels < 1 || channels > 4) throw std::runtime_error("save_image_ldr: bad channel count");
    std::ofstream out(path, std::ios::binary);
    if (!out.is_open()) throw std::runtime_error("save_image_ldr: cannot open " + path);
    if (ext == ".png") {
        static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
        out.write(reinterpret_cast<const char*>(sig), 8);
        std::vector<uint8_t> ihdr;
        put_be32(ihdr, uint32_t(w));
        put_be32(ihdr, uint32_t(h));
        static const uint8_t color_type[5] = { 0, 0, 4, 2, 6 };   // grey, grey+alpha, RGB, RGBA
        ihdr.push_back(8); ihdr.push_back(color_type[channels]); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
        png_chunk(out, "IHDR", ihdr);
        const size_t row = size_t(w) * channels;
        std::vector<uint8_t> raw((row + 1) * h);
        for (int y = 0; y < h; ++y) {
            const int ys = flip ? h - 1 - y : y;
            raw[(row + 1) * y] = 0;   // filter: none
            memcpy(&raw[(row + 1) * y + 1], pixels + row * ys, row);
        }
        uLongf clen = compressBound(uLong(raw.size()));
        std::vector<uint8_t> comp(clen);
        if (compress2(comp.data(), &clen, raw.data(), uLong(raw.size()), 6) != Z_OK) throw std::runtime_error("save_image_ldr: deflate failed");
        comp.resize(clen);
        png_chunk(out, "IDAT", comp);
        png_chunk(out, "IEND", {});
    } else if (ext == ".ppm" && channels >= 3) {
        out << "P6\n" << w << " " << h << "\n255\n";
        for (int y = 0; y < h; ++y) {
            const int ys = flip ? h - 1 - y : y;
            for (int x = 0; x < w; ++x) out.write(reinterpret_cast<const char*>(pixels + (size_t(ys) * w + x) * channels), 3);
        }
    } else
        throw std::runtime_error("save_image_ldr: unsupported image format: " + ext);
}

}  // namespace volren
