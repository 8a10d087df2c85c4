// main.cpp -- the `volren` command line of the B200 host. Same arguments, same order-dependent semantics and the
// same offline loop as reference src/main.cpp (init_opengl_from_args :311-357, parse_cmd :360-435, handle_path
// :93-102, offline loop :524-558). There is no window: the interactive loop (:477-523 -- progressive trace(), draw(),
// camera input, save at sppx) is stood in for by `--preview FILE`: the same loop with the tonemapped framebuffer written
// to FILE (atomically, every --preview-interval seconds) for any image viewer that reloads on change, and "input" read
// from FILE.cmd (one line of ordinary command-line flags, e.g. `--cam_pos 0 1 2 --density 50`, or `quit`). Without
// --preview the executable renders offline (`--render` is accepted and implied).
// New flags: --gpus N, --partition spp|tile, --device D, --batch N (samples per launch, progress granularity),
// --preview FILE, --preview-interval SEC, --preview-keep (stay alive at sppx and wait for commands).
#include <pybind11/embed.h>
#include <pybind11/eval.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "image_io.h"

#include "camera.h"
#include "context.h"
#include "renderer.h"

namespace fs = std::filesystem;
using namespace volren;

// ------------------------------------------
// settings

static bool interactive = true;   // kept for flag parity; without a window both modes run the offline loop
static std::string out_filename = "output.png";
static int batch_spp = 64;
static std::string preview_file;          // --preview FILE: progressive preview (stand-in for the interactive window)
static double preview_interval = 0.5;     // seconds between two writes of FILE
static bool preview_keep = false;         // keep running at sppx and wait for FILE.cmd (the window would stay open)

static std::shared_ptr<RendererOpenGL> renderer;

// ------------------------------------------
// helper funcs (main.cpp:37-102)

static void load_volume(const std::string& path) {
    try {
        std::cout << "load volume: " << path << std::endl;
        if (fs::is_directory(path))
            renderer->volume = voldata::Volume::load_folder(path, { "density", "temperature", "flame", "flames" });
        else
            renderer->volume = std::make_shared<voldata::Volume>(path);
        renderer->density_scale = 1.f;
        renderer->scale_and_move_to_unit_cube();
        renderer->commit();
        renderer->sample = 0;
    } catch (std::runtime_error& e) {
        std::cerr << "Unable to load volume from " << path << ": " << e.what() << std::endl;
    }
}

static void load_envmap(const std::string& path) {
    try {
        renderer->environment = std::make_shared<Environment>(path);
        renderer->sample = 0;
    } catch (std::runtime_error& e) {
        std::cerr << "Unable to load envmap from " << path << ": " << e.what() << std::endl;
    }
}

static void load_transferfunc(const std::string& path) {
    try {
        renderer->transferfunc = std::make_shared<TransferFunction>(path);
        renderer->show_environment = false;
        renderer->sample = 0;
    } catch (std::runtime_error& e) {
        std::cerr << "Unable to load transferfunc from " << path << ": " << e.what() << std::endl;
    }
}

static void run_script(const std::string& path) {
    try {
        pybind11::scoped_interpreter guard{};
        {   // the embedded interpreter starts from the system prefix: add the site-packages of the active / build-time venv
            pybind11::module_ site = pybind11::module_::import("site");
            if (const char* venv = std::getenv("VIRTUAL_ENV")) {
                for (auto& e : fs::directory_iterator(fs::path(venv) / "lib"))
                    if (fs::is_directory(e.path() / "site-packages")) site.attr("addsitedir")((e.path() / "site-packages").string());
            }
#ifdef VOLREN_SITE_PACKAGES
            if (fs::is_directory(VOLREN_SITE_PACKAGES)) site.attr("addsitedir")(std::string(VOLREN_SITE_PACKAGES));
#endif
        }
        pybind11::eval_file(path);
        renderer->sample = 0;
    } catch (pybind11::error_already_set& e) {
        std::cerr << "Error executing python script " << path << ": " << e.what() << std::endl;
    }
}

static void handle_path(const std::string& path) {
    const std::string ext = fs::path(path).extension().string();
    if (ext == ".py") run_script(path);
    else if (ext == ".hdr") load_envmap(path);
    else if (ext == ".txt") load_transferfunc(path);
    else load_volume(path);
}

// ------------------------------------------
// command line options

// first pass: what used to create the GL context (main.cpp:311-357). Window flags are accepted and ignored,
// with their operand counts, so existing command lines keep working.
static void init_context_from_args(int argc, char** argv) {
    ContextParameters params;
    for (int i = 1; i < argc; ++i) {
        const std::string arg = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) throw std::runtime_error("missing operand for " + arg); return argv[++i]; };
        if (arg == "-w") params.width = uint32_t(std::stoi(next()));
        else if (arg == "-h") params.height = uint32_t(std::stoi(next()));
        else if (arg == "--title" || arg == "--major" || arg == "--minor" || arg == "--swap" || arg == "--font" || arg == "--fontsize") next();
        else if (arg == "--gpus") params.n_gpus = std::stoi(next());
        else if (arg == "--partition") params.partition = next();
        else if (arg == "--device") params.first_device = std::stoi(next());
        // --no-resize --hidden --render --no-decoration --floating --maximised ---debug: no operands
    }
    if (params.partition != "spp" && params.partition != "tile") throw std::runtime_error("--partition must be spp or tile");
    Context::init(params);
}

// second pass: renderer / camera flags, processed IN ORDER (main.cpp:360-435)
static void parse_cmd(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
        const std::string arg = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) throw std::runtime_error("missing operand for " + arg); return argv[++i]; };
        auto tf = [&]() { if (!renderer->transferfunc) renderer->transferfunc = std::make_shared<TransferFunction>(); };
        if (arg == "--render") interactive = false;
        else if (arg == "--output") out_filename = next();
        else if (arg == "--samples" || arg == "--spp" || arg == "--sppx") renderer->sppx = std::stoi(next());
        else if (arg == "--bounces") renderer->bounces = std::stoi(next());
        else if (arg == "--albedo") renderer->albedo = glm::vec3(std::stof(next()));
        else if (arg == "--density") renderer->density_scale = std::stof(next());
        else if (arg == "--emission") renderer->emission_scale = std::stof(next());
        else if (arg == "--phase") renderer->phase = std::stof(next());
        else if (arg == "--env_strength") renderer->environment->strength = std::stof(next());
        else if (arg == "--env_rot") renderer->environment->transform = glm::to_mat3(glm::rotate(glm::mat4(1.f), glm::radians(std::stof(next())), glm::vec3(0, 1, 0)));
        else if (arg == "--env_hide") renderer->show_environment = false;
        else if (arg == "--turbo") { tf(); renderer->transferfunc->colormap(colormap::ColormapType::Turbo); }
        else if (arg == "--viridis") { tf(); renderer->transferfunc->colormap(colormap::ColormapType::Viridis); }
        else if (arg == "--fau") {
            renderer->transferfunc = std::make_shared<TransferFunction>(std::vector<glm::vec4>({ glm::vec4(0.f), glm::vec4(4 / 255.f, 49 / 255.f, 106 / 255.f, 0.33f),
                                                                                                 glm::vec4(38 / 255.f, 97 / 255.f, 65 / 255.f, 0.66f), glm::vec4(151 / 255.f, 27 / 255.f, 47 / 255.f, 1.f) }));
        } else if (arg == "--tf_left") { if (renderer->transferfunc) renderer->transferfunc->window_left = std::stof(next()); }
        else if (arg == "--tf_width") { if (renderer->transferfunc) renderer->transferfunc->window_width = std::stof(next()); }
        else if (arg == "--cam_pos") { auto& p = current_camera()->pos; p.x = std::stof(next()); p.y = std::stof(next()); p.z = std::stof(next()); }
        else if (arg == "--cam_dir") { auto& d = current_camera()->dir; d.x = std::stof(next()); d.y = std::stof(next()); d.z = std::stof(next()); }
        else if (arg == "--cam_fov") current_camera()->fov_degree = std::stof(next());
        else if (arg == "--exposure") renderer->tonemap_exposure = std::stof(next());
        else if (arg == "--gamma") renderer->tonemap_gamma = std::stof(next());
        // --vol_rot_*: the result is truncated to a mat3 (drops the translation), exactly as main.cpp:417-422
        else if (arg == "--vol_rot_x") renderer->volume->transform = glm::mat4(glm::to_mat3(glm::rotate(renderer->volume->transform, glm::radians(std::stof(next())), glm::vec3(1, 0, 0))));
        else if (arg == "--vol_rot_y") renderer->volume->transform = glm::mat4(glm::to_mat3(glm::rotate(renderer->volume->transform, glm::radians(std::stof(next())), glm::vec3(0, 1, 0))));
        else if (arg == "--vol_rot_z") renderer->volume->transform = glm::mat4(glm::to_mat3(glm::rotate(renderer->volume->transform, glm::radians(std::stof(next())), glm::vec3(0, 0, 1))));
        else if (arg == "--vol_crop_min") { auto& v = renderer->vol_clip_min; v.x = std::stof(next()); v.y = std::stof(next()); v.z = std::stof(next()); }
        else if (arg == "--vol_crop_max") { auto& v = renderer->vol_clip_max; v.x = std::stof(next()); v.y = std::stof(next()); v.z = std::stof(next()); }
        // operands of the context pass: skip them here so that e.g. `-w 1024` is not taken for a path
        else if (arg == "-w" || arg == "-h" || arg == "--title" || arg == "--major" || arg == "--minor" || arg == "--swap" || arg == "--font" || arg == "--fontsize" ||
                 arg == "--gpus" || arg == "--partition" || arg == "--device") ++i;
        else if (arg == "--batch") batch_spp = std::max(1, std::stoi(next()));
        else if (arg == "--preview") preview_file = next();
        else if (arg == "--preview-interval") preview_interval = std::stod(next());
        else if (arg == "--preview-keep") preview_keep = true;
        else if (fs::is_regular_file(argv[i]) || fs::is_directory(argv[i])) handle_path(argv[i]);
    }
}

// ------------------------------------------
// progressive preview (main.cpp:477-523 without a window)

static void write_preview() {
    renderer->draw();                                         // tonemap.fs / blit.fs -> RGBA8 framebuffer (:516)
    const std::vector<uint8_t> fb = renderer->read_framebuffer();
    const fs::path target(preview_file);
    const fs::path tmp = target.parent_path() / (target.stem().string() + ".tmp" + target.extension().string());
    store_ldr(tmp.string(), fb.data(), int(renderer->color.w), int(renderer->color.h), 4, true);
    fs::rename(tmp, target);                                  // viewers never see a half-written file
}

// "input": one line of command-line flags in FILE.cmd, consumed (deleted) once applied. Returns false on `quit`.
static bool poll_preview_commands(bool& changed) {
    const std::string cmd_file = preview_file + ".cmd";
    std::ifstream in(cmd_file);
    if (!in.is_open()) return true;
    std::stringstream text;
    text << in.rdbuf();
    in.close();
    fs::remove(cmd_file);
    std::vector<std::string> tok = { "volren" };
    std::string t;
    while (text >> t) tok.push_back(t);
    if (tok.size() == 1) return true;
    if (tok[1] == "quit") return false;
    std::vector<char*> argv;
    for (auto& s : tok) argv.push_back(s.data());
    parse_cmd(int(argv.size()), argv.data());
    changed = true;
    return true;
}

static void preview_loop() {
    using clock = std::chrono::steady_clock;
    current_camera()->update();
    std::cout << "preview: " << preview_file << " (commands: " << preview_file << ".cmd)" << std::endl;
    auto last_write = clock::now() - std::chrono::hours(1), t0 = clock::now();
    long long traced = 0;
    bool saved = false;
    while (true) {
        bool changed = false;
        if (!poll_preview_commands(changed)) break;
        if (changed) {                                        // CameraImpl::default_input_handler(...) -> renderer->reset() (:481-483)
            current_camera()->update();
            renderer->reset();
            saved = false;
        }
        if (renderer->sample < renderer->sppx) {
            const int n = std::min(batch_spp, renderer->sppx - renderer->sample);
            renderer->trace(n);
            Context::swap_buffers();
            traced += n;
            if (renderer->sample == renderer->sppx && !saved) {   // :511-512 (linear colour, flipped)
                renderer->color.save_ldr(out_filename, true, true);
                saved = true;
                write_preview();
                last_write = clock::now();
                const double sec = std::chrono::duration<double>(clock::now() - t0).count();
                std::cout << renderer->sample << " / " << renderer->sppx << " spp, " << double(traced) * renderer->color.w * renderer->color.h / sec / 1e9
                          << " Gsamples/s incl. preview writes; " << out_filename << " written." << std::endl;
                if (!preview_keep) break;
            }
        } else
            std::this_thread::sleep_for(std::chrono::milliseconds(100));   // glfwWaitEventsTimeout(1.f / 10): 10 fps idle (:514)
        if (std::chrono::duration<double>(clock::now() - last_write).count() >= preview_interval) {
            write_preview();
            last_write = clock::now();
            std::cout << renderer->sample << " / " << renderer->sppx << "\r" << std::flush;
        }
    }
}

// ------------------------------------------
// main

int main(int argc, char** argv) {
    try {
        init_context_from_args(argc, argv);
    } catch (std::exception& e) {
        std::cerr << "Failed to create context: " << e.what() << std::endl;
        return 1;
    }

    renderer = std::make_shared<RendererOpenGL>();
    renderer->init();

    // default cam pos (main.cpp:458-459)
    current_camera()->pos = glm::vec3(1, 0, 1);
    current_camera()->dir = glm::normalize(-current_camera()->pos);

    try {
        parse_cmd(argc, argv);

        // debug box if no volume has been loaded (main.cpp:465-474)
        if (renderer->volume->grids.empty()) {
            const float scale = 1.f;
            float values[4] = { 1, 2.5, 5, 10 };
            auto box = std::make_shared<voldata::DenseGrid>(1, 1, 4, values);
            box->transform = glm::translate(glm::scale(glm::mat4(1.f), glm::vec3(scale)), 2 * scale * current_camera()->dir + glm::vec3(0, -0.5f, -2));
            renderer->volume = std::make_shared<voldata::Volume>(box);
            renderer->commit();
        }
        renderer->reset();

        if (!preview_file.empty()) {
            preview_loop();
            Context::shutdown();
            return 0;
        }

        // offline loop (main.cpp:524-558)
        current_camera()->update();
        std::cout << "rendering..." << std::endl;
        for (size_t i = 0; i < renderer->volume->n_grid_frames(); ++i) {
            renderer->reset();
            renderer->volume->grid_frame_counter = i;
            while (renderer->sample < renderer->sppx) {
                renderer->trace(std::min(batch_spp, renderer->sppx - renderer->sample));
                std::cout << renderer->sample << " / " << renderer->sppx << "\r" << std::flush;
                Context::swap_buffers();
            }
            renderer->tonemap_in_place();
            const size_t n_zero = 6;
            const std::string idx = std::to_string(i);
            const std::string out_fn = fs::path(out_filename).stem().string() + "_" + std::string(n_zero - std::min(n_zero, idx.length()), '0') + idx + ".png";
            renderer->color.save_ldr(out_fn);
            std::cout << out_fn << " written." << std::endl;
        }
    } catch (std::exception& e) {
        std::cerr << "volren: " << e.what() << std::endl;
        Context::shutdown();
        return 1;
    }
    Context::shutdown();
    return 0;
}
