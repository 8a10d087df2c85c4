// bindings.cpp -- the `volpy` Python module of the VolRen host: the surface of reference src/bindings.cpp:64-417
// (ImageDataFloat, Volume, Environment, TransferFunction, Renderer with its static camera properties and colmap
// helpers, glm vec/ivec/uvec/mat/quat with arithmetic and the buffer protocol) over the B200 renderer.
// Built twice from this file:
//   -DVOLPY_EMBEDDED : PYBIND11_EMBEDDED_MODULE, linked into the `volren` executable (as in the reference)
//   otherwise        : a regular extension module `volpy*.so`, importable from any Python process
// Additions (not in the reference): numpy-array Volume constructors, `Grid`, and `volpy.create_context(...)`.
#include <pybind11/numpy.h>
#include <pybind11/operators.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#ifdef VOLPY_EMBEDDED
#include <pybind11/embed.h>
#endif

#include <cmath>
#include <filesystem>
#include <iostream>

#include "camera.h"
#include "context.h"
#include "environment.h"
#include "image_io.h"
#include "renderer.h"
#include "transferfunc.h"
#include "voldata.h"

namespace py = pybind11;
using namespace volren;

namespace {

template <typename VecT, typename ScalarT> py::class_<VecT>& register_vector_operators(py::class_<VecT>& c) {
    return c.def(py::self + py::self).def(py::self + ScalarT()).def(ScalarT() + py::self).def(py::self += py::self).def(py::self += ScalarT())
        .def(py::self - py::self).def(py::self - ScalarT()).def(ScalarT() - py::self).def(py::self -= py::self).def(py::self -= ScalarT())
        .def(py::self * py::self).def(py::self * ScalarT()).def(ScalarT() * py::self).def(py::self *= py::self).def(py::self *= ScalarT())
        .def(py::self / py::self).def(py::self / ScalarT()).def(ScalarT() / py::self).def(py::self /= py::self).def(py::self /= ScalarT())
        .def(-py::self);
}

template <typename MatT, typename ScalarT> py::class_<MatT>& register_matrix_operators(py::class_<MatT>& c) {
    return c.def(py::self + py::self).def(py::self += py::self).def(py::self - py::self).def(py::self -= py::self).def(py::self * py::self)
        .def(py::self * ScalarT()).def(ScalarT() * py::self).def(py::self *= py::self).def(py::self *= ScalarT()).def(-py::self);
}

template <typename V> py::buffer_info float_buffer_1d(V& v, py::ssize_t n) {
    return py::buffer_info(&v[0], sizeof(float), py::format_descriptor<float>::format(), 1, { n }, { py::ssize_t(sizeof(float)) });
}

int render_batch() {
    if (const char* b = std::getenv("VOLREN_BATCH")) return std::max(1, std::atoi(b));
    return 256;
}

void bind_volpy(py::module_& m) {
    using voldata::Buf3D;
    using voldata::Grid;
    using voldata::Volume;

    m.doc() = "volpy: VolRen renderer module (B200 back end)";

    // new: explicit context creation for stand-alone use (the embedded module gets its context from the CLI flags)
    m.def("create_context", [](uint32_t width, uint32_t height, int gpus, const std::string& partition, int device) {
        ContextParameters p;
        p.width = width; p.height = height; p.n_gpus = gpus; p.partition = partition; p.first_device = device;
        Context::init(p);
    }, py::arg("width") = 1280, py::arg("height") = 720, py::arg("gpus") = 1, py::arg("partition") = "spp", py::arg("device") = 0);

    // new: the host's own image codecs, exposed for tests and scripts
    m.def("load_hdr", [](const std::string& path, bool flip) {
        ImageF img = load_hdr(path, flip);
        py::array_t<float> out({ py::ssize_t(img.h), py::ssize_t(img.w), py::ssize_t(3) });
        std::copy(img.data.begin(), img.data.end(), out.mutable_data());
        return out;
    }, py::arg("path"), py::arg("flip") = true);
    m.def("load_environment_image", [](const std::string& path) {      // what Environment(path) reads: .hdr or LDR .png, bottom-up RGB
        ImageF img = load_environment_image(path);
        py::array_t<float> out({ py::ssize_t(img.h), py::ssize_t(img.w), py::ssize_t(3) });
        std::copy(img.data.begin(), img.data.end(), out.mutable_data());
        return out;
    }, py::arg("path"));
    m.def("save_ldr", [](const std::string& path, py::array_t<uint8_t, py::array::c_style | py::array::forcecast> px, bool flip) {
        if (px.ndim() != 3) throw std::runtime_error("save_ldr: expected an (h, w, c) uint8 array");
        store_ldr(path, px.data(), int(px.shape(1)), int(px.shape(0)), int(px.shape(2)), flip);
    }, py::arg("path"), py::arg("pixels"), py::arg("flip") = true);

    // ---- voldata::Buf3D<float> (bindings.cpp:69-77): shape (x, y, z), strides (4*z*y, 4*z, 4) ----
    py::class_<Buf3D<float>, std::shared_ptr<Buf3D<float>>>(m, "ImageDataFloat", py::buffer_protocol()).def_buffer([](Buf3D<float>& buf) -> py::buffer_info {
        return py::buffer_info(buf.data.data(), sizeof(float), py::format_descriptor<float>::format(), 3,
                               { py::ssize_t(buf.stride.x), py::ssize_t(buf.stride.y), py::ssize_t(buf.stride.z) },
                               { py::ssize_t(sizeof(float) * buf.stride.z * buf.stride.y), py::ssize_t(sizeof(float) * buf.stride.z), py::ssize_t(sizeof(float)) });
    });

    // ---- grids (new: lets load_grid / add_grid_frame / update_grid_frame be used from Python) ----
    py::class_<Grid, std::shared_ptr<Grid>>(m, "Grid")
        .def("minorant_majorant", &Grid::minorant_majorant)
        .def("index_extent", &Grid::index_extent)
        .def("num_voxels", &Grid::num_voxels)
        .def("size_bytes", &Grid::size_bytes)
        .def("lookup", &Grid::lookup)
        .def_readwrite("transform", &Grid::transform)
        .def("__repr__", [](const Grid& g) { return g.to_string(""); });

    // ---- voldata::Volume (bindings.cpp:82-95) ----
    py::class_<Volume, std::shared_ptr<Volume>>(m, "Volume")
        .def(py::init<>())
        .def(py::init<std::string>())
        .def(py::init([](size_t w, size_t h, size_t d, py::array_t<uint8_t, py::array::c_style | py::array::forcecast> a) {
            if (size_t(a.size()) != w * h * d) throw std::runtime_error("Volume: array size does not match w*h*d");
            return std::make_shared<Volume>(w, h, d, a.data());
        }))
        .def(py::init([](size_t w, size_t h, size_t d, py::array_t<float, py::array::c_style | py::array::forcecast> a) {
            if (size_t(a.size()) != w * h * d) throw std::runtime_error("Volume: array size does not match w*h*d");
            return std::make_shared<Volume>(w, h, d, a.data());
        }))
        .def_static("load_grid", &Volume::load_grid, py::arg("filename"), py::arg("gridname") = "density")
        .def_static("load_folder", &Volume::load_folder, py::arg("path"), py::arg("gridnames") = std::vector<std::string>{ "density" })
        .def("clear", &Volume::clear)
        .def("add_grid_frame", &Volume::add_grid_frame, py::arg("frame") = Volume::GridFrame())
        .def("update_grid_frame", &Volume::update_grid_frame, py::arg("i"), py::arg("grid"), py::arg("gridname") = "density")
        .def("n_grid_frames", &Volume::n_grid_frames)
        .def("AABB", &Volume::AABB)   // no default argument, as in the reference
        .def_readwrite("grid_frame_counter", &Volume::grid_frame_counter)
        .def_readwrite("transform", &Volume::transform)
        .def("minorant_majorant", &Volume::minorant_majorant, py::arg("gridname") = "density")
        .def("__repr__", &Volume::to_string, py::arg("indent") = "");

    // ---- environment (bindings.cpp:100-102) ----
    py::class_<Environment, std::shared_ptr<Environment>>(m, "Environment")
        .def(py::init<std::string>())
        .def(py::init([](py::array_t<float, py::array::c_style | py::array::forcecast> rgb) {
            if (rgb.ndim() != 3 || rgb.shape(2) != 3) throw std::runtime_error("Environment: expected an (h, w, 3) float array (bottom-up rows)");
            return std::make_shared<Environment>(int(rgb.shape(1)), int(rgb.shape(0)), rgb.data());
        }))
        .def_readwrite("strength", &Environment::strength)
        .def_readwrite("transform", &Environment::transform);

    // ---- transfer function (bindings.cpp:107-113) ----
    py::class_<TransferFunction, std::shared_ptr<TransferFunction>>(m, "TransferFunction")
        .def(py::init<>())
        .def(py::init<const std::string&>())
        .def(py::init<const std::vector<glm::vec4>&>())
        .def("randomize", &TransferFunction::randomize, py::arg("n_bins") = 8)
        .def_readwrite("window_left", &TransferFunction::window_left)
        .def_readwrite("window_width", &TransferFunction::window_width)
        .def_readonly("lut", &TransferFunction::lut);
    // test hook (not part of the reference surface): the LUT of TransferFunction::colormap(type, n_bins) -- what `--turbo` /
    // `--viridis` upload (main.cpp:385-392); type = index into tinycolormap::ColormapType
    m.def("_colormap_lut", [](int type, size_t n_bins) {
        TransferFunction tf;
        tf.colormap(static_cast<colormap::ColormapType>(type), n_bins);
        return tf.lut;
    }, py::arg("type"), py::arg("n_bins") = 256);

    // ---- renderer (bindings.cpp:118-209) ----
    py::class_<RendererOpenGL, std::shared_ptr<RendererOpenGL>>(m, "Renderer")
        .def(py::init<>())
        .def("init", &RendererOpenGL::init)
        .def("commit", &RendererOpenGL::commit)
        .def("trace", static_cast<void (RendererOpenGL::*)()>(&RendererOpenGL::trace))
        .def("reset", &RendererOpenGL::reset)
        .def("scale_and_move_to_unit_cube", &RendererOpenGL::scale_and_move_to_unit_cube)
        .def("render", [](const std::shared_ptr<RendererOpenGL>& renderer, int spp) {
            current_camera()->update();
            renderer->sample = 0;
            const int batch = render_batch();
            while (renderer->sample < spp) {
                renderer->trace(std::min(batch, spp - renderer->sample));
                Context::swap_buffers();
            }
        })
        .def("draw", [](const std::shared_ptr<RendererOpenGL>& renderer) {
            renderer->draw();
            Context::swap_buffers();
        })
        .def_static("resolution", []() { return Context::resolution(); })
        .def("fbo_data", [](const std::shared_ptr<RendererOpenGL>& renderer) {
            auto buf = std::make_shared<Buf3D<float>>(glm::uvec3(renderer->color.w, renderer->color.h, 3));
            buf->data = renderer->read_color(3);
            return buf;
        })
        .def("save", [](const std::shared_ptr<RendererOpenGL>& renderer, const std::string& filename) {
            const glm::ivec2 size = Context::resolution();
            const std::vector<uint8_t> rgba = renderer->read_framebuffer();
            std::vector<uint8_t> pixels(size_t(size.x) * size.y * 3);
            for (size_t i = 0; i < size_t(size.x) * size.y; ++i) { pixels[3 * i] = rgba[4 * i]; pixels[3 * i + 1] = rgba[4 * i + 1]; pixels[3 * i + 2] = rgba[4 * i + 2]; }
            const std::filesystem::path outfile = filename;
            store_ldr(outfile.string(), pixels.data(), size.x, size.y, 3);
            std::cout << outfile << " written." << std::endl;
        }, py::arg("filename") = "out.png")
        .def("save_with_alpha", [](const std::shared_ptr<RendererOpenGL>& renderer, const std::string& filename) {
            const glm::ivec2 size = Context::resolution();
            const std::vector<uint8_t> pixels = renderer->read_framebuffer();
            const std::filesystem::path outfile = std::filesystem::path(filename).replace_extension(".png");
            store_ldr(outfile.string(), pixels.data(), size.x, size.y, 4);
            std::cout << outfile << " written." << std::endl;
        }, py::arg("filename") = "out.png")
        // new: n samples in one launch, and the uniform block trace() would upload (raw bytes of vrb_params) for tests
        .def("trace_samples", static_cast<void (RendererOpenGL::*)(int)>(&RendererOpenGL::trace))
        .def("tonemap_in_place", &RendererOpenGL::tonemap_in_place)
        .def("_params", [](const std::shared_ptr<RendererOpenGL>& renderer) {
            current_camera()->update();
            const vrb_params p = renderer->debug_params();
            return py::bytes(reinterpret_cast<const char*>(&p), sizeof p);
        })
        // members
        .def_readwrite("volume", &RendererOpenGL::volume)
        .def_readwrite("environment", &RendererOpenGL::environment)
        .def_readwrite("transferfunc", &RendererOpenGL::transferfunc)
        .def_readwrite("sample", &RendererOpenGL::sample)
        .def_readwrite("sppx", &RendererOpenGL::sppx)
        .def_readwrite("bounces", &RendererOpenGL::bounces)
        .def_readwrite("seed", &RendererOpenGL::seed)
        .def_readwrite("tonemap_exposure", &RendererOpenGL::tonemap_exposure)
        .def_readwrite("tonemap_gamma", &RendererOpenGL::tonemap_gamma)
        .def_readwrite("tonemapping", &RendererOpenGL::tonemapping)
        .def_readwrite("show_environment", &RendererOpenGL::show_environment)
        .def_readwrite("albedo", &RendererOpenGL::albedo)
        .def_readwrite("phase", &RendererOpenGL::phase)
        .def_readwrite("density_scale", &RendererOpenGL::density_scale)
        .def_readwrite("emission_scale", &RendererOpenGL::emission_scale)
        .def_readwrite("vol_clip_min", &RendererOpenGL::vol_clip_min)
        .def_readwrite("vol_clip_max", &RendererOpenGL::vol_clip_max)
        // camera: bound by ADDRESS of the global camera's members, like the reference
        .def_readwrite_static("cam_pos", &current_camera()->pos)
        .def_readwrite_static("cam_dir", &current_camera()->dir)
        .def_readwrite_static("cam_up", &current_camera()->up)
        .def_readwrite_static("cam_fov", &current_camera()->fov_degree)
        .def_readwrite_static("cam_near", &current_camera()->near)
        .def_readwrite_static("cam_far", &current_camera()->far)
        .def_readwrite_static("view_matrix", &current_camera()->view)
        .def_readwrite_static("proj_matrix", &current_camera()->proj)
        .def_static("cam_aspect", &CameraImpl::aspect_ratio)
        // addition: the camera matrices are otherwise refreshed by render() / the main loops (CameraImpl::update)
        .def_static("update_camera", []() { current_camera()->update(); })
        // colmap
        .def_static("colmap_view_trans", []() {
            const glm::mat4 GL_TO_COLMAP = glm::inverse(glm::mat4(1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1));
            return glm::vec3((GL_TO_COLMAP * current_camera()->view)[3]);
        })
        .def_static("colmap_view_rot", []() {
            const glm::mat4 GL_TO_COLMAP = glm::inverse(glm::mat4(1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1));
            return glm::normalize(glm::toQuat(GL_TO_COLMAP * current_camera()->view));
        })
        .def_static("colmap_focal_length", []() { return Context::resolution().y / (2 * std::tan(0.5 * double(glm::radians(current_camera()->fov_degree)))); })
        .def_static("shutdown", []() {
            Context::shutdown();
#ifdef VOLPY_EMBEDDED
            std::cout << std::flush;
            exit(0);   // the reference ends the process here (bindings.cpp:207-209)
#endif
        });

    // ---- glm vectors (bindings.cpp:214-337) ----
    {
        auto c = py::class_<glm::vec2>(m, "vec2").def(py::init<>()).def(py::init<float>()).def(py::init<float, float>())
            .def_readwrite("x", &glm::vec2::x).def_readwrite("y", &glm::vec2::y)
            .def("normalize", [](const glm::vec2& v) { return glm::normalize(v); }).def("length", [](const glm::vec2& v) { return glm::length(v); })
            .def("__repr__", [](const glm::vec2& v) { return glm::to_string(v); });
        register_vector_operators<glm::vec2, float>(c);
    }
    {
        auto c = py::class_<glm::vec3>(m, "vec3", py::buffer_protocol()).def(py::init<>()).def(py::init<float>()).def(py::init<float, float, float>())
            .def_readwrite("x", &glm::vec3::x).def_readwrite("y", &glm::vec3::y).def_readwrite("z", &glm::vec3::z)
            .def("normalize", [](const glm::vec3& v) { return glm::normalize(v); }).def("length", [](const glm::vec3& v) { return glm::length(v); })
            .def_buffer([](glm::vec3& v) { return float_buffer_1d(v, 3); })
            .def("__repr__", [](const glm::vec3& v) { return glm::to_string(v); });
        register_vector_operators<glm::vec3, float>(c);
    }
    {
        auto c = py::class_<glm::vec4>(m, "vec4", py::buffer_protocol()).def(py::init<>()).def(py::init<float>()).def(py::init<float, float, float, float>())
            .def_readwrite("x", &glm::vec4::x).def_readwrite("y", &glm::vec4::y).def_readwrite("z", &glm::vec4::z).def_readwrite("w", &glm::vec4::w)
            .def("normalize", [](const glm::vec4& v) { return glm::normalize(v); }).def("length", [](const glm::vec4& v) { return glm::length(v); })
            .def_buffer([](glm::vec4& v) { return float_buffer_1d(v, 4); })
            .def("__repr__", [](const glm::vec4& v) { return glm::to_string(v); });
        register_vector_operators<glm::vec4, float>(c);
    }
#define VOLPY_INT_VEC(T, S, NAME)                                                                                                     \
    {                                                                                                                                 \
        auto c2 = py::class_<glm::T##2>(m, NAME "2").def(py::init<>()).def(py::init<S>()).def(py::init<S, S>())                        \
            .def_readwrite("x", &glm::T##2::x).def_readwrite("y", &glm::T##2::y)                                                      \
            .def("__repr__", [](const glm::T##2& v) { return glm::to_string(v); });                                                   \
        register_vector_operators<glm::T##2, S>(c2);                                                                                  \
        auto c3 = py::class_<glm::T##3>(m, NAME "3").def(py::init<>()).def(py::init<S>()).def(py::init<S, S, S>())                     \
            .def_readwrite("x", &glm::T##3::x).def_readwrite("y", &glm::T##3::y).def_readwrite("z", &glm::T##3::z)                    \
            .def("__repr__", [](const glm::T##3& v) { return glm::to_string(v); });                                                   \
        register_vector_operators<glm::T##3, S>(c3);                                                                                  \
        auto c4 = py::class_<glm::T##4>(m, NAME "4").def(py::init<>()).def(py::init<S>()).def(py::init<S, S, S, S>())                  \
            .def_readwrite("x", &glm::T##4::x).def_readwrite("y", &glm::T##4::y).def_readwrite("z", &glm::T##4::z).def_readwrite("w", &glm::T##4::w) \
            .def("__repr__", [](const glm::T##4& v) { return glm::to_string(v); });                                                   \
        register_vector_operators<glm::T##4, S>(c4);                                                                                  \
    }
    VOLPY_INT_VEC(ivec, int32_t, "ivec")
    VOLPY_INT_VEC(uvec, uint32_t, "uvec")
#undef VOLPY_INT_VEC

    // ---- glm matrices (bindings.cpp:342-389): the buffer exposes COLUMNS as numpy rows ----
    {
        auto c = py::class_<glm::mat3>(m, "mat3", py::buffer_protocol()).def(py::init<>()).def(py::init<float>()).def(py::init<glm::vec3, glm::vec3, glm::vec3>())
            .def("column", [](const glm::mat3& mm, uint32_t i) { return mm[int(i)]; })
            .def("value", [](const glm::mat3& mm, uint32_t i, uint32_t j) { return mm[int(i)][int(j)]; })
            .def_buffer([](glm::mat3& mm) {
                return py::buffer_info(&mm[0].x, sizeof(float), py::format_descriptor<float>::format(), 2, { 3, 3 }, { py::ssize_t(sizeof(float) * 3), py::ssize_t(sizeof(float)) });
            })
            .def("__repr__", [](const glm::mat3& mm) { return glm::to_string(mm); });
        register_matrix_operators<glm::mat3, float>(c);
    }
    {
        auto c = py::class_<glm::mat4>(m, "mat4", py::buffer_protocol()).def(py::init<>()).def(py::init<float>()).def(py::init<glm::vec4, glm::vec4, glm::vec4, glm::vec4>())
            .def("column", [](const glm::mat4& mm, uint32_t i) { return mm[int(i)]; })
            .def("value", [](const glm::mat4& mm, uint32_t i, uint32_t j) { return mm[int(i)][int(j)]; })
            .def_buffer([](glm::mat4& mm) {
                return py::buffer_info(&mm[0].x, sizeof(float), py::format_descriptor<float>::format(), 2, { 4, 4 }, { py::ssize_t(sizeof(float) * 4), py::ssize_t(sizeof(float)) });
            })
            .def("__repr__", [](const glm::mat4& mm) { return glm::to_string(mm); });
        register_matrix_operators<glm::mat4, float>(c);
    }
    // ---- glm quaternion (bindings.cpp:394-416): buffer order x, y, z, w ----
    {
        auto c = py::class_<glm::quat>(m, "quat", py::buffer_protocol()).def(py::init<>())
            .def(py::init([](const glm::vec3& euler) { return glm::quat_from_euler(euler); }))
            .def(py::init([](const glm::mat3& mm) { return glm::toQuat(mm); }))
            .def(py::init([](const glm::mat4& mm) { return glm::toQuat(mm); }))
            .def_readwrite("x", &glm::quat::x).def_readwrite("y", &glm::quat::y).def_readwrite("z", &glm::quat::z).def_readwrite("w", &glm::quat::w)
            .def_buffer([](glm::quat& q) { return float_buffer_1d(q, 4); })
            .def("__repr__", [](const glm::quat& q) { return glm::to_string(q); });
        register_matrix_operators<glm::quat, float>(c);
    }
}

}  // namespace

#ifdef VOLPY_EMBEDDED
PYBIND11_EMBEDDED_MODULE(volpy, m) { bind_volpy(m); }
#else
PYBIND11_MODULE(volpy, m) { bind_volpy(m); }
#endif
