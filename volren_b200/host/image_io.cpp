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

constexpr uint64_t MAX_IMAGE_PIXELS = uint64_t(1) << 28;     // 16k x 16k: far above any environment map, far below size_t overflow

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
    if (uint64_t(w) * uint64_t(h) > MAX_IMAGE_PIXELS) throw std::runtime_error("HDR image too large (" + std::to_string(w) + " x " + std::to_string(h) + "): " + path);
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

uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace

ImageF load_png_rgb(const std::string& path, bool flip) {
    std::ifstream in(path, std::ios::binary);
    if (!in.is_open()) throw std::runtime_error("Failed to load image file: " + path);
    const std::vector<uint8_t> d((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    if (d.size() < 33 || memcmp(d.data(), sig, 8) != 0) throw std::runtime_error("Failed to load image file (not a PNG): " + path);
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    for (size_t pos = 8; pos + 12 <= d.size();) {
        const uint32_t len = be32(&d[pos]);
        if (pos + 12 + size_t(len) > d.size()) throw std::runtime_error("truncated PNG: " + path);
        const uint8_t* c = &d[pos + 8];
        if (!memcmp(&d[pos + 4], "IHDR", 4) && len >= 13) { w = be32(c); h = be32(c + 4); depth = c[8]; ctype = c[9]; interlace = c[12]; }
        else if (!memcmp(&d[pos + 4], "PLTE", 4)) plte.assign(c, c + len);
        else if (!memcmp(&d[pos + 4], "IDAT", 4)) idat.insert(idat.end(), c, c + len);
        else if (!memcmp(&d[pos + 4], "IEND", 4)) break;
        pos += 12 + size_t(len);
    }
    const int comps = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!w || !h || !comps || (depth != 8 && depth != 16) || (ctype == 3 && depth != 8) || interlace)
        throw std::runtime_error("unsupported PNG (need 8/16-bit, non-interlaced gray / RGB / palette / alpha variants): " + path);
    // sizes come from the file: cap them before any product is formed (a crafted IHDR would otherwise wrap size_t and leave
    // undersized buffers for the unfilter loop)
    if (uint64_t(w) * uint64_t(h) > MAX_IMAGE_PIXELS) throw std::runtime_error("PNG too large (" + std::to_string(w) + " x " + std::to_string(h) + "): " + path);
    const size_t bpp = size_t(comps) * (depth / 8), stride = size_t(w) * bpp;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf raw_len = uLongf(raw.size());
    if (uncompress(raw.data(), &raw_len, idat.data(), uLong(idat.size())) != Z_OK || raw_len != raw.size())
        throw std::runtime_error("corrupt PNG data: " + path);
    std::vector<uint8_t> px(stride * h);
    for (uint32_t y = 0; y < h; ++y) {       // undo the scanline filters (PNG spec 9.2)
        const uint8_t* src = &raw[(stride + 1) * y];
        uint8_t* dst = &px[stride * y];
        const uint8_t* up = y ? &px[stride * (y - 1)] : nullptr;
        const int filter = src[0];
        if (filter > 4) throw std::runtime_error("corrupt PNG filter: " + path);
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? dst[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            const int pred = filter == 0 ? 0 : filter == 1 ? a : filter == 2 ? b : filter == 3 ? ((a + b) >> 1) : paeth(a, b, c);
            dst[i] = uint8_t(src[1 + i] + pred);
        }
    }
    ImageF img;
    img.w = int(w); img.h = int(h); img.channels = 3;
    img.data.assign(size_t(w) * h * 3, 0.f);
    for (uint32_t y = 0; y < h; ++y) {
        const uint32_t oy = flip ? h - 1 - y : y;
        for (uint32_t x = 0; x < w; ++x) {
            const uint8_t* s = &px[stride * y + size_t(x) * bpp];
            uint8_t v[4] = { 0, 0, 0, 255 };
            for (int k = 0; k < comps; ++k) v[k] = s[size_t(k) * (depth / 8)];      // 16-bit: the high (first) byte
            float* o = &img.data[(size_t(oy) * w + x) * 3];
            if (ctype == 3) {
                if (size_t(v[0]) * 3 + 2 >= plte.size()) throw std::runtime_error("PNG palette index out of range: " + path);
                for (int k = 0; k < 3; ++k) o[k] = plte[size_t(v[0]) * 3 + k] / 255.f;
            } else {
                const int n = std::min(comps, 3);
                for (int k = 0; k < n; ++k) o[k] = v[k] / 255.f;
            }
        }
    }
    return img;
}

ImageF load_environment_image(const std::string& path) {
    std::string ext = std::filesystem::path(path).extension().string();
    std::transform(ext.begin(), ext.end(), ext.begin(), ::tolower);
    if (ext == ".hdr") return load_hdr(path, true);
    if (ext == ".png") return load_png_rgb(path, true);
    throw std::runtime_error("Failed to load image file: " + path + " (environment maps are read from .hdr and .png)");
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
