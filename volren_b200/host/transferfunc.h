// transferfunc.h -- RGBA look-up table + window of the VolRen host (reference src/transferfunc.{h,cpp}).
// `upload_gpu()` keeps its name: it freezes the LUT that the device will see (`lut_gpu`, CDF-corrected when the alpha
// channel is not monotone, transferfunc.cpp:45-58) and bumps `version`; the renderer pushes it with vrb_tf_upload.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "vmath.h"

namespace colormap {
// tinycolormap::ColormapType, same names and order (tinycolormap.hpp:78-81)
enum class ColormapType { Parula, Heat, Jet, Turbo, Hot, Gray, Magma, Inferno, Plasma, Viridis, Cividis, Github, Cubehelix, HSV };
// tinycolormap::GetColor: exact tables / closed forms, double precision (transferfunc.cpp; tables in colormap_tables.inc)
void GetColor(double x, ColormapType type, double out[3]);
vmath::vec3 GetColor(float x, ColormapType type);
}  // namespace colormap

class TransferFunction {
public:
    TransferFunction();
    explicit TransferFunction(const std::string& path);
    explicit TransferFunction(colormap::ColormapType type);
    explicit TransferFunction(const std::vector<vmath::vec4>& lut);
    virtual ~TransferFunction() {}

    static std::vector<vmath::vec4> compute_lut_cdf(const std::vector<vmath::vec4>& lut);
    void upload_gpu();
    void randomize(size_t n_bins = 8);
    void colormap(colormap::ColormapType type, size_t n_bins = 256);
    void load_from_file(const std::string& path);
    void write_to_file(const std::string& filename);

    // data
    float window_left, window_width;
    std::vector<vmath::vec4> lut;
    std::vector<vmath::vec4> lut_gpu;   // what the SSBO held in the reference
    uint64_t id, version;
};
